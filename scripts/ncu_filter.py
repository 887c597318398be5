"""Keep the columns of an `ncu --page raw --csv` export that the roofline discussion needs (stdin -> stdout)."""
import csv
import re
import sys

KEEP = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|"
                  r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__pipe_tensor_cycles_active.*pct_of_peak_sustained_(active|elapsed)|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__occupancy_limit.*|"
                  r"launch__shared_mem_per_block_dynamic|lts__t_bytes\.sum$|lts__t_sector_hit_rate\.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|"
                  r"smsp__cycles_active\.avg|sm__cycles_elapsed\.max|sm__inst_executed_pipe_uniform.*|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__warp_issue_stalled.*_per_warp_active\.pct)$")
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
if not rows:
    sys.exit(0)
idx = [i for i, h in enumerate(rows[0]) if KEEP.match(h)]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] if i < len(r) else "" for i in idx])
