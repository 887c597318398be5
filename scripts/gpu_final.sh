#!/bin/bash
# round-2 final single-GPU pass: whole GPU suite, the bench line (all configs) + reference arm, ncu launch list and
# --set full captures of the GEMM and bandwidth kernel families (one eager training step, scripts/profile_step.py)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${TAG:-r2final}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [ "${PYTEST:-1}" = "1" ]; then
timeout 2400 python -m pytest tests -m gpu -q -s --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
fi
timeout 900 python bench.py --steps 20 --warmup 5 --per-layer-out $O/per_layer.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?" >> $O/rc.txt
if [ "${NCU:-1}" = "1" ]; then
timeout 300 python scripts/profile_step.py > $O/plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/rc.txt
cap() {  # name, regex, count
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -c $3 -f -o $O/$1 python scripts/profile_step.py > $O/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?" >> $O/rc.txt
  ncu -i $O/$1.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_filter.py > $O/$1.raw.csv
  ls -la $O/$1.ncu-rep 2>/dev/null | awk '{print $5}' >> $O/rc.txt
  rm -f $O/$1.ncu-rep      # gpurun merges at most 64 MiB back; the filtered CSV is what profiles/ keeps
}
cap conv3 igemm_conv3 12
cap wgrad igemm_wgrad 10
cap fwdgen igemm_fwd 8
cap bw 'scale_shift|bn_bwd|act_pool|maxpool|head_|adam|pack_batch|pad_channels|channel_sum|bn_finalize' 45
rm -f $O/*.ncu-rep.tmp
du -sh $O
fi
cat $O/rc.txt; tail -6 $O/pytest_gpu.log
