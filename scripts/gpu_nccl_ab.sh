#!/bin/bash
# A/B under `gpurun --gpus N`: NCCL CTA budget x SM margin of the persistent GEMM kernels, configs[1] and configs[4]
cd "$GRAFT_REPO_ROOT" || exit 1
N=${NGPU:-2}
O=gpurun_out/${TAG:-r2n}; mkdir -p $O
run() {  # name, NCCL_MAX_CTAS (or -), PLUME_SM_MARGIN
  local envs="PLUME_SM_MARGIN=$3"
  [ "$2" != "-" ] && envs="$envs NCCL_MAX_CTAS=$2"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 15 --warmup 5 --configs "${CONFIGS:-4}" --no-cpu-baseline > $O/$1.json 2> $O/$1.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/$1.json").read().strip().splitlines()[-1])
    o=d.get("other_configs",{})
    print("$1", "cfg1 ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v.get("ms_per_step",-1),3) for k,v in o.items()})
except Exception as e: print("$1 ERR", e)
PY
}
run base - 0
run c8_m8 8 8
run c16_m16 16 16
run c4_m4 4 4
run c8_m0 8 0
run cdef_m16 - 16
