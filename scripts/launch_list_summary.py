"""Turns an ncu launch list of bench.py (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv) into the committed artefacts: a compact per-launch CSV, a per-kernel summary and a traffic JSON (the
per-family DRAM traffic bench.py reports as roofline.traffic).
Usage: python scripts/launch_list_summary.py gpurun_out/launches.csv profiles/r2_launches_final [profiles/r2_traffic.json] ["what was profiled"]"""
import collections
import csv
import json
import re
import sys

src, stem = sys.argv[1], sys.argv[2]
traffic_path = sys.argv[3] if len(sys.argv) > 3 else "profiles/r2_traffic.json"
what = sys.argv[4] if len(sys.argv) > 4 else "python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
by = collections.OrderedDict()
for r in rows:
    d = by.setdefault(int(r["ID"]), {
        "id": int(r["ID"]),
        "kernel": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("plume::", "").replace("at::native::", "")[:70],
        "grid": r["Grid Size"], "block": r["Block Size"], "stream": r["Stream"]})
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time"):
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    elif "read" in r["Metric Name"]:
        d["dram_read_bytes"] = v * scale[u]
    else:
        d["dram_write_bytes"] = v * scale[u]
L = list(by.values())
with open(stem + ".csv", "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=list(L[0].keys()))
    w.writeheader()
    w.writerows(L)
tot = sum(d["us"] for d in L)
agg = collections.OrderedDict()
for d in L:
    a = agg.setdefault(d["kernel"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["us"]
    a[2] += d["dram_read_bytes"] + d["dram_write_bytes"]
out = ["ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none " + what,
       f"{len(L)} launches, {tot:.1f} us total; serialised, cold cache: compare SHARES",
       f"{'us total':>12} {'share':>6} {'n':>5} {'us/launch':>10} {'DRAM MB/launch':>15}  kernel"]
for n, (c, v, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v:12.1f} {100 * v / tot:5.1f}% {c:5d} {v / c:10.1f} {b / c / 1e6:15.2f}  {n[:90]}")
fam = {"fwd_kernel": ("igemm_conv3_kernel", "igemm_fwd_kernel"), "wgrad_kernel": ("igemm_wgrad3_kernel", "igemm_wgrad_kernel")}
traffic = {}
for f_, keys in fam.items():
    sel = [d for d in L if d["kernel"].startswith(keys)]
    traffic[f_] = {"launches": len(sel),
                   "dram_bytes_per_launch": sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in sel) / len(sel),
                   "us_per_launch": sum(d["us"] for d in sel) / len(sel),
                   "share_of_step": sum(d["us"] for d in sel) / tot,
                   "source": stem + ".csv (ncu, dram__bytes_read.sum + dram__bytes_write.sum per launch)"}
    out.append(f"family {f_}: {len(sel)} launches, {traffic[f_]['share_of_step'] * 100:.1f}% of the time, "
               f"{traffic[f_]['dram_bytes_per_launch'] / 1e6:.1f} MB DRAM traffic per launch")
open(stem + ".txt", "w").write("\n".join(out) + "\n")
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("\n".join(out[:14] + out[-2:]))
