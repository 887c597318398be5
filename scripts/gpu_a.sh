#!/bin/bash
# round-2 GPU pass A: full GPU test suite, bench (all configs), sanitizer logs, ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 --per-layer-out $O/per_layer.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?" >> $O/rc.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_run.py > $O/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/rc.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_run.py train > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs '' > $O/ncu_bench.log 2>&1; echo "ncu rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -5 $O/pytest_gpu.log
