#!/bin/bash
# sweep A/B + whole GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2s}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_golden.py -m gpu -q > $O/pytest_sweep.log 2>&1; echo "sweep tests rc=$?" >> $O/rc.txt
timeout 600 python scripts/bench_sweep.py > $O/bench_sweep.json 2> $O/bench_sweep.err; echo "bench_sweep rc=$?" >> $O/rc.txt
timeout 2400 python -m pytest tests -m gpu -q -s --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/pytest_sweep.log; cat $O/bench_sweep.json | cut -c1-400; tail -4 $O/pytest_gpu.log
