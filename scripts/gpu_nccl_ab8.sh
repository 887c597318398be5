#!/bin/bash
# A/B under `gpurun --gpus N`: NCCL stream priority x CTA budget x SM margin, configs[1] and configs[4]
cd "$GRAFT_REPO_ROOT" || exit 1
N=${NGPU:-8}
O=gpurun_out/${TAG:-r2p}; mkdir -p $O
run() {  # name, env assignments...
  local name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus $N --steps 15 --warmup 5 --configs "${CONFIGS:-4}" --no-cpu-baseline > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/$name.json").read().strip().splitlines()[-1])
    o=d.get("other_configs",{})
    print("$name", "cfg1 ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v.get("ms_per_step",-1),3) for k,v in o.items()})
except Exception as e: print("$name ERR", e)
PY
}
run base PLUME_SM_MARGIN=0 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,GRAPH
grep -i -m5 "nvls" $O/base.err | cut -c1-200; grep -i -m3 "channels\|algo" $O/base.err | cut -c1-200
run hp_m0 PLUME_NCCL_HIGH_PRIORITY=1 PLUME_SM_MARGIN=0
run hp_c16_m16 PLUME_NCCL_HIGH_PRIORITY=1 NCCL_MAX_CTAS=16 PLUME_SM_MARGIN=16
run hp_c8_m8 PLUME_NCCL_HIGH_PRIORITY=1 NCCL_MAX_CTAS=8 PLUME_SM_MARGIN=8
run hp_m32 PLUME_NCCL_HIGH_PRIORITY=1 PLUME_SM_MARGIN=32
run hp_c32_m24 PLUME_NCCL_HIGH_PRIORITY=1 NCCL_MAX_CTAS=32 PLUME_SM_MARGIN=24
