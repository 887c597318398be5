#!/bin/bash
# multi-GPU pass (run under `gpurun --gpus N`): the NCCL parity tests (N >= 2) and the bench at N GPUs with every config
cd "$GRAFT_REPO_ROOT" || exit 1
N=${NGPU:-2}
O=gpurun_out/${TAG:-r2m$N}; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -s > $O/pytest_nccl.log 2>&1; echo "pytest_nccl rc=$?" >> $O/rc.txt
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 --configs "${CONFIGS:-2,3,4}" > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_nccl.log
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    for k,v in d.get("other_configs",{}).items(): print(k, {a:b for a,b in v.items() if a in ("ms_per_step","tiles_per_s","ms_per_scene","error","tflops_per_gpu","n_gpus")})
except Exception as e: print("ERR", e)
PY
tail -5 $O/bench_${N}gpu.err
