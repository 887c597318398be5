#!/bin/bash
# CTA-pair conv3: op tests, then A/B bench (PLUME_CONV3_PAIR=0 vs default) on one box
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2x}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv3" > $O/pytest_ops.log 2>&1; rc=$?; echo "ops rc=$rc" >> $O/rc.txt
tail -15 $O/pytest_ops.log
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_determinism.py -m gpu -q -x > $O/pytest_net.log 2>&1; echo "net rc=$?" >> $O/rc.txt
tail -4 $O/pytest_net.log
for rep in 1 2; do
PLUME_CONV3_PAIR=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_off_$rep.json 2>> $O/bench.err
PLUME_CONV3_PAIR=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_p256_$rep.json 2>> $O/bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_on_$rep.json 2>> $O/bench.err
done
python - <<'PY'
import json,os,glob
O=os.path.join("gpurun_out",os.environ.get("TAG","r2x"))
for f in sorted(glob.glob(O+"/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pl={(r["layer"],r["pass"]):r["us"] for r in d["per_layer"]}
        print(os.path.basename(f),"ms/step",round(d["ms_per_step"],3),"fwd TF/s",round(d["roofline"]["achieved"]), {k:round(pl[k],1) for k in (("enc0.conv2","fwd"),("enc1.conv1","fwd"),("enc1.conv2","fwd"),("dec0.conv1","fwd"),("dec1.conv1","fwd"),("dec0.conv1","dgrad"),("enc2.conv2","fwd"))})
    except Exception as e: print(f,"ERR",e)
PY
fi
cat $O/rc.txt
