#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2y}; mkdir -p $O
timeout ${TMO:-600} python ${SCRIPT} > $O/out.log 2>&1; echo "rc=$?" >> $O/out.log
tail -60 $O/out.log
