#!/bin/bash
# round-2 GPU pass B: full GPU test suite without -x (every failure listed), then a short bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2b}; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -s --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs '' > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -15 $O/pytest_gpu.log
