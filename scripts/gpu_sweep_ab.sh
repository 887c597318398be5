#!/bin/bash
# GPU box: sweep tests on the default build, then the sweep bench under every PLUME_SWEEP_VARIANT of interest
# (bit 0 band-local phase, bit 1 path halving, bit 2 skip known stretches, bit 3 one-strip mask kernel), then ncu.
# The variant switch existed in the build measured in profiles/r2_sweep_variants.txt (commit d2128d5^ .. d2128d5); the
# kept variant is now the only code path, and VARIANTS=0 simply runs the bench once.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2ab}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
for v in ${VARIANTS:-0 4 2 6 1 5 7 13}; do
  PLUME_SWEEP_VARIANT=$v timeout 300 python scripts/bench_sweep.py > $O/bench_v$v.json 2> $O/err_v$v.txt || tail -5 $O/err_v$v.txt
  python - $v $O/bench_v$v.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); print(f"variant {sys.argv[1]:>2}: one call {d['ms_per_timestamp']*1e3:7.1f} us  per sweep {d['ms_per_timestamp_one_call_per_sweep']*1e3:7.1f} us  dense {d['ms_per_timestamp_dense_planes']*1e3:7.1f} us  parity {d['parity_on_sample']}  e2e {1e3/d['e2e']['value']:.2f} ms")
PY
done
timeout 600 ncu -k regex:bits --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 190 -c 10 --csv --log-file $O/launches.csv python scripts/bench_sweep.py > $O/ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv,os,collections,re
O=os.path.join("gpurun_out",os.environ.get("TAG","r2ab"))
rows=list(csv.DictReader(l for l in open(O+"/launches.csv") if not l.startswith("==")))
agg=collections.OrderedDict()
for r in rows:
    k=re.sub(r"\(.*","",r["Kernel Name"]).replace("plume::","")
    a=agg.setdefault(k,{"n":0,"us":0,"rd":0,"wr":0})
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time"): a["us"]+= v/1e3 if u.startswith("n") else v; a["n"]+=1
    elif "read" in r["Metric Name"]: a["rd"]+= v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]
    else: a["wr"]+= v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]
for k,a in agg.items(): print(f"{k:32s} n={a['n']:3d} {a['us']/max(a['n'],1):8.1f} us/launch  rd {a['rd']/max(a['n'],1)/1e6:8.1f} MB wr {a['wr']/max(a['n'],1)/1e6:8.1f} MB")
PY
