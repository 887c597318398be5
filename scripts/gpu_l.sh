#!/bin/bash
# A/B on one box: one vs two MMA issuer threads in the CTA-pair conv3 kernels (two libraries built beforehand)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2ab}; mkdir -p $O
L=kcl_ltss_bioatm_b200
cp $L/libplume_b200.so /tmp/lib_iss1.so
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv3" > $O/pytest_ops.log 2>&1; rc=$?; echo "ops(iss1) rc=$rc" >> $O/rc.txt
tail -3 $O/pytest_ops.log
if [ $rc -eq 0 ]; then
for rep in 1 2; do
cp $L/libplume_b200_iss2.so $L/libplume_b200.so
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_iss2_$rep.json 2>> $O/bench.err
cp /tmp/lib_iss1.so $L/libplume_b200.so
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_iss1_$rep.json 2>> $O/bench.err
done
python - <<'PY'
import json,os,glob
O=os.path.join("gpurun_out",os.environ.get("TAG","r2ab"))
for f in sorted(glob.glob(O+"/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pl={(r["layer"],r["pass"]):r["us"] for r in d["per_layer"]}
        print(os.path.basename(f),"ms/step",round(d["ms_per_step"],3),"fwd TF/s",round(d["roofline"]["achieved"]), {k:round(pl[k],1) for k in (("enc0.conv2","fwd"),("enc0.conv2","dgrad"),("enc1.conv1","fwd"),("enc1.conv2","fwd"),("dec0.conv1","fwd"),("dec1.conv1","fwd"),("dec0.conv1","dgrad"))})
    except Exception as e: print(f,"ERR",e)
PY
fi
cat $O/rc.txt
