#!/bin/bash
# full GPU suite on the pair build + A/B: 128->128 layers resident (half staging, 3 halo slots) vs ring (6 halo slots) in pair mode
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2z}; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -s --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -4 $O/pytest_gpu.log
for rep in 1 2; do
PLUME_CONV3_RESIDENT_HALF=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_ring_$rep.json 2>> $O/bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --configs '' > $O/bench_res_$rep.json 2>> $O/bench.err
done
python - <<'PY'
import json,os,glob
O=os.path.join("gpurun_out",os.environ.get("TAG","r2z"))
for f in sorted(glob.glob(O+"/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pl={(r["layer"],r["pass"]):r["us"] for r in d["per_layer"]}
        print(os.path.basename(f),"ms/step",round(d["ms_per_step"],3),"fwd TF/s",round(d["roofline"]["achieved"]), {k:round(pl[k],1) for k in (("enc1.conv2","fwd"),("dec1.conv2","fwd"),("enc1.conv2","dgrad"),("dec1.conv2","dgrad"))})
    except Exception as e: print(f,"ERR",e)
PY
cat $O/rc.txt
