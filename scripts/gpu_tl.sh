#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
N=${NGPU:-2}; O=gpurun_out/${TAG:-r2t}; mkdir -p $O
for w in wide default; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/timeline_dp.py $O/timeline_${w}_${N}gpu.txt $w graph > $O/tl_$w.log 2>&1; echo "$w rc=$?"
tail -4 $O/timeline_${w}_${N}gpu.txt
done
grep -c nccl $O/timeline_wide_${N}gpu.txt
