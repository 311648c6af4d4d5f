#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round:  bash scripts/profile_round.sh <tag>
# (bench JSON lines of both arms, the ncu launch list of the bench command, ncu metrics of the real 1 M-query
# launch, and one `--set full` capture of a steady-state slice).  Numbers printed under ncu are never bench values.
TAG=${1:-rX}; O=gpurun_out; mkdir -p $O
python bench.py --impl reference > $O/bench_${TAG}_n1_reference.json 2> $O/bench_ref.err
python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench.err
tail -c 600 $O/bench_${TAG}_n1_reference.json; echo; tail -c 3000 $O/bench_${TAG}_n1.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --batch 65536 --steps 2 --warmup 1 --no-cpu > $O/ncu_bench.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
M=$M,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
ncu --metrics $M --clock-control none -k regex:c2a_solve --launch-skip 1 -c 1 --csv --log-file $O/${TAG}_real_launch_1M_metrics.csv \
    python scripts/one_launch.py --batch 1000000 > $O/ncu_real.log 2>&1
C2A_B200_MAX_BLOCKS=16 ncu --set full --clock-control none --import-source on -k regex:c2a_solve --launch-skip 1 -c 1 \
    -o $O/${TAG}_full python scripts/prof_run.py --batch 12288 > $O/ncu_full.log 2>&1
ls -la $O
