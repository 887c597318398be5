#!/bin/bash
# GPU box: sweep tests with both mask-kernel variants, A/B of the bench (PLUME_SWEEP_RANK=0 ballots, 1 rank + transpose;
# the switch existed in the build measured in profiles/r2_sweep_variants.txt, the rank variant is now the only one),
# ncu launch list and one --set full capture of the mask and merge kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2rk}; mkdir -p $O
for r in 0 1; do
  PLUME_SWEEP_RANK=$r timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu > $O/pytest_rank$r.log 2>&1; echo "pytest rank=$r rc=$?"; tail -3 $O/pytest_rank$r.log
done
for r in 0 1 0 1; do
  PLUME_SWEEP_RANK=$r timeout 300 python scripts/bench_sweep.py > $O/bench_rank$r.json 2> $O/err_rank$r.txt || tail -5 $O/err_rank$r.txt
  python - $r $O/bench_rank$r.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); print(f"rank {sys.argv[1]}: one call {d['ms_per_timestamp']*1e3:7.1f} us  per sweep {d['ms_per_timestamp_one_call_per_sweep']*1e3:7.1f} us  parity {d['parity_on_sample']}  e2e {1e3/d['e2e']['value']:.2f} ms  fill {d['nearest_fill']}")
PY
done
timeout 600 ncu -k regex:bits --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 190 -c 8 --csv --log-file $O/launches.csv python scripts/bench_sweep.py > $O/ncu.log 2>&1
echo "ncu rc=$?"; grep -v "^==" $O/launches.csv | grep "gpu__time" | cut -d, -f5,12- | head -12
timeout 900 ncu -k regex:"mask_bits|bits_merge|bits_flatten" --set full --clock-control none --import-source on -s 36 -c 3 -o $O/sweep_full python scripts/bench_sweep.py > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i $O/sweep_full.ncu-rep --page raw --csv > $O/sweep_full_raw.csv 2>/dev/null; wc -c $O/sweep_full_raw.csv
