#!/bin/bash
# A/B under `gpurun --gpus N`: Adam per gradient bucket as its all-reduce completes (PLUME_ADAM_PER_BUCKET=1, default)
# vs. wait for every all-reduce, then one Adam launch (=0); configs[1] and the wide configs[4]; NCCL parity tests first.
cd "$GRAFT_REPO_ROOT" || exit 1
N=${NGPU:-2}
O=gpurun_out/${TAG:-r2ad}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -s > $O/pytest_nccl.log 2>&1; echo "pytest_nccl rc=$?"; tail -3 $O/pytest_nccl.log
run() {  # name, PLUME_ADAM_PER_BUCKET
  PLUME_ADAM_PER_BUCKET=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 20 --warmup 5 --configs "${CONFIGS:-4}" --no-cpu-baseline > $O/$1.json 2> $O/$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/$1.json").read().strip().splitlines()[-1])
    o=d.get("other_configs",{})
    print("$1", "cfg1 ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v.get("ms_per_step",-1),3) for k,v in o.items()})
except Exception as e: print("$1 ERR", e)
PY
}
run all_0 0
run bucket_0 1
run all_1 0
run bucket_1 1
