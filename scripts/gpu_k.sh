#!/bin/bash
# wgrad3 CTA pairs: tests + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2aa}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "wgrad" > $O/pytest_ops.log 2>&1; rc=$?; echo "ops rc=$rc" >> $O/rc.txt
tail -8 $O/pytest_ops.log
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_determinism.py tests/test_gpu_precise.py -m gpu -q -x > $O/pytest_net.log 2>&1; echo "net rc=$?" >> $O/rc.txt
tail -4 $O/pytest_net.log
for rep in 1 2; do
PLUME_WGRAD3_PAIR=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_off_$rep.json 2>> $O/bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_on_$rep.json 2>> $O/bench.err
done
python - <<'PY'
import json,os,glob
O=os.path.join("gpurun_out",os.environ.get("TAG","r2aa"))
for f in sorted(glob.glob(O+"/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pl={(r["layer"],r["pass"]):r["us"] for r in d["per_layer"]}
        print(os.path.basename(f),"ms/step",round(d["ms_per_step"],3),"wgrad TF/s",round(d["roofline_wgrad"]["achieved"]), {k:round(pl[k],1) for k in (("dec1.conv1","wgrad"),("dec2.conv2","wgrad"),("dec2.conv1","wgrad"),("dec3.conv1","wgrad"),("bottleneck.conv2","wgrad"),("enc3.conv2","wgrad"))})
    except Exception as e: print(f,"ERR",e)
PY
fi
cat $O/rc.txt
