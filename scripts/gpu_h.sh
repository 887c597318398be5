#!/bin/bash
# A/B of the resident-weights / half-staging conv3 configuration + the op tests that cover it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2u}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_determinism.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
for rep in 1 2; do
PLUME_CONV3_RESIDENT_HALF=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_off_$rep.json 2>> $O/bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_on_$rep.json 2>> $O/bench.err
done
cat $O/rc.txt; tail -3 $O/pytest.log
python - <<'PY'
import json,os,glob
O=os.path.join("gpurun_out",os.environ.get("TAG","r2u"))
for f in sorted(glob.glob(O+"/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pl={(r["layer"],r["pass"]):r["us"] for r in d["per_layer"]}
        print(os.path.basename(f),"ms/step",round(d["ms_per_step"],3),"fwd TF/s",round(d["roofline"]["achieved"]), {k:round(pl[k],1) for k in (("enc1.conv1","fwd"),("dec0.conv1","dgrad"),("dec0.conv1","fwd"),("enc1.conv1","dgrad"))})
    except Exception as e: print(f,"ERR",e)
PY
