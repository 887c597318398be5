#!/bin/bash
# GPU box: sweep tests, the sweep bench line, and an ncu launch list of two one-call timestamps (bit-plane kernels).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2sw}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 300 python scripts/bench_sweep.py > $O/bench_sweep.json 2> $O/err.txt || { tail -20 $O/err.txt; exit 1; }
timeout 600 ncu -k regex:bits --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 190 -c 10 --csv --log-file $O/launches.csv python scripts/bench_sweep.py > $O/ncu.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv,os,collections,re
O=os.path.join("gpurun_out",os.environ.get("TAG","r2sw"))
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
cut -c1-700 $O/bench_sweep.json
