#!/bin/bash
# targeted GPU pass: the tests named in $TESTS (default: the bf16x3 suite), every failure listed
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2x}; mkdir -p $O
timeout ${TMO:-900} python -m pytest ${TESTS:-tests/test_gpu_precise.py} -m gpu -q -s --durations=8 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -40 $O/pytest.log
