"""profiles/r2_sass_igemm.txt: per-kernel counts of the tcgen05 / TMA / mbarrier SASS instructions in the built library.
Usage: python scripts/sass_summary.py [libplume_b200.so] > profiles/r2_sass_igemm.txt   (needs cuobjdump, c++filt)"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "kcl_ltss_bioatm_b200", "libplume_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEEP = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTMAPF", "SYNCS", "UTCATOMSWS", "HMMA", "RED", "REDG", "ATOMG",
        "ATOMS", "UTMACMDFLUSH", "UTMACCTL")
fn, per = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    mm = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line) if fn else None
    if mm and mm.group(1).split(".")[0] in KEEP:
        per[fn][mm.group(1)] += 1
print("# SASS evidence for the tensor-core kernels of libplume_b200.so (cuobjdump -sass, sm_100a)")
print("# per kernel: counts of the instructions that prove tcgen05 (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR =")
print("# tcgen05.commit), TMA (UTMALDG = cp.async.bulk.tensor load, UTMASTG = store, UTMACCTL.PF = descriptor prefetch) and")
print("# mbarrier traffic (SYNCS).  No HMMA (mma.sync) anywhere; the PAIR instantiations of igemm_conv3_kernel carry the .2CTA forms (UTCHMMA.2CTA, UTMALDG.*.2CTA, UTCBAR.2CTA.MULTICAST).")
tot = collections.Counter()
for fn, c in per.items():
    if not any(k.startswith(("UTCHMMA", "UTMALDG", "LDTM")) for k in c):
        continue
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print("\n" + re.sub(r"\(.*", "", dem))
    for k, v in sorted(c.items()):
        print(f"    {v:5d}  {k}")
        tot[k.split(".")[0]] += v
print("\n# totals over the library: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items())))
