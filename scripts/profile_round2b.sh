#!/bin/bash
# Refresh of the round-2 ncu evidence on the FINAL build (wide kernel launched beside the main kernel): launch list of
# the bench command and the DRAM / cache / pipe metrics of the real 1M-query launches.  bash scripts/profile_round2b.sh
O=gpurun_out; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
M=$M,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2b_launches_bench_1M.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_bench_r2b.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"c2a_(solve|wide)_kernel" --launch-skip 2 -c 3 --csv --log-file $O/r2b_real_launch_1M_metrics.csv \
    python scripts/one_launch.py --batch 1000000 > $O/ncu_real_r2b.log 2>&1
tail -2 $O/ncu_bench_r2b.log $O/ncu_real_r2b.log
grep -c c2a_ $O/r2b_launches_bench_1M.csv $O/r2b_real_launch_1M_metrics.csv
