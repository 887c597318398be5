#!/bin/bash
# A/B of the fused BatchNorm-backward reduction (pool / head backward) on one box + the tests that cover it + mode timings
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2j}; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_determinism.py tests/test_gpu_precise.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
for rep in 1 2; do
PLUME_NO_FUSED_BN_REDUCE=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_unfused_$rep.json 2>> $O/bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_fused_$rep.json 2>> $O/bench.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs 'm' > $O/bench_modes.json 2>> $O/bench.err; echo "modes rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -4 $O/pytest.log
python - <<'PY'
import json,os,glob
O=os.path.join("gpurun_out",os.environ.get("TAG","r2j"))
for f in sorted(glob.glob(O+"/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],3), json.dumps(d.get("other_configs",{}).get("modes"))[:600])
    except Exception as e: print(f, "ERR", e)
PY
