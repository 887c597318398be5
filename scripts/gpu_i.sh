#!/bin/bash
# 2-GPU: NCCL tests (incl. the bf16 wire format) + bench A/B fp32 vs bf16 gradient communication
cd "$GRAFT_REPO_ROOT" || exit 1
N=${NGPU:-2}; O=gpurun_out/${TAG:-r2v}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -s > $O/pytest_nccl.log 2>&1; echo "pytest_nccl rc=$?" >> $O/rc.txt
for mode in fp32 bf16; do
PLUME_GRAD_COMM=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29525 \
  bench.py --gpus $N --steps 15 --warmup 5 --configs 4 --no-cpu-baseline > $O/bench_$mode.json 2> $O/bench_$mode.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$mode.json").read().strip().splitlines()[-1])
    print("$mode", "cfg1 ms/step", round(d["ms_per_step"],3), {k:round(v.get("ms_per_step",-1),3) for k,v in d.get("other_configs",{}).items()})
except Exception as e: print("$mode ERR", e)
PY
done
cat $O/rc.txt; tail -6 $O/pytest_nccl.log
