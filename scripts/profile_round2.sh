#!/bin/bash
# One gpurun call that produces the ncu evidence of round 2 under gpurun_out/ (copied to profiles/ afterwards):
#   bash scripts/profile_round2.sh <tag>
# launch list of the bench command, DRAM / cache / pipe metrics of the REAL 1M-query launch of both kernels of a step,
# `--set full` captures of a steady-state slice of c2a_solve_kernel and of c2a_wide_kernel, and the same metrics on a
# bunny batch.  Numbers printed under ncu are never bench values.
TAG=${1:-r2}; O=gpurun_out; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
M=$M,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --batch 65536 --steps 2 --warmup 1 --no-cpu > $O/ncu_bench.log 2>&1
ncu --metrics $M --clock-control none -k regex:"c2a_(solve|wide)_kernel" --launch-skip 2 -c 2 --csv --log-file $O/${TAG}_real_launch_1M_metrics.csv \
    python scripts/one_launch.py --batch 1000000 > $O/ncu_real.log 2>&1
ncu --metrics $M --clock-control none -k regex:"c2a_(solve|wide)_kernel" --launch-skip 2 -c 2 --csv --log-file $O/${TAG}_bunny_10k_metrics.csv \
    python scripts/one_launch.py --batch 10000 --bunny > $O/ncu_bunny.log 2>&1
C2A_B200_NO_WIDE=1 C2A_B200_MAX_BLOCKS=16 ncu --set full --clock-control none --import-source on -k regex:c2a_solve --launch-skip 1 -c 1 \
    -o $O/${TAG}_solve_full python scripts/prof_run.py --batch 12288 > $O/ncu_full_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:c2a_wide --launch-skip 1 -c 1 \
    -o $O/${TAG}_wide_full python scripts/wide_stats.py 4096 > $O/ncu_full_wide.log 2>&1
tail -2 $O/ncu_real.log $O/ncu_bunny.log $O/ncu_full_solve.log $O/ncu_full_wide.log
ls -la $O | tail -12
