#!/bin/bash
# A/B of the generic forward kernel's epilogue warpgroups + the tests that exercise it + a short bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2g}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_precise.py tests/test_gpu_unet.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
PLUME_FWD_ONE_EPILOGUE_WG=1 timeout 300 python scripts/time_convT.py > $O/convT_one_wg.txt 2>&1
timeout 300 python scripts/time_convT.py > $O/convT_two_wg.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs '' > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest.log; cat $O/convT_one_wg.txt $O/convT_two_wg.txt
python - <<'PY'
import json,os
d=json.loads(open(os.path.join("gpurun_out",os.environ.get("TAG","r2g"),"bench.json")).read().strip().splitlines()[-1])
print("ms/step",d["ms_per_step"],"value",d["value"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"])
PY
