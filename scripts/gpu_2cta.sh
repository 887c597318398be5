#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2w}; mkdir -p $O
timeout 120 ./scripts/test_2cta > $O/test_2cta.log 2>&1; echo "rc=$?" >> $O/test_2cta.log
cat $O/test_2cta.log
