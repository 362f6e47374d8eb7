#!/bin/bash
# Round profile pass on a B200 box (run through gpurun from the repo root):
#   bash tools/profile_round.sh            -> everything lands in gpurun_out/, summaries are copied to profiles/ by hand
# Every ncu command runs only after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
set -x
O=gpurun_out
python bench.py > $O/bench_c3_1gpu.json 2> $O/bench_c3_1gpu.err || exit 1
python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err || exit 1
B="python bench.py --n 512 --steps 2 --warmup 3 --no-cpu --no-e2e"
$B > $O/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 30 --csv \
    --log-file $O/launches.csv $B > $O/ncu_launches.log 2>&1
if [ "$1" != "nofull" ]; then
ncu --set full --clock-control none --import-source on -k regex:stage_tiled -s 10 -c 3 -f -o $O/prof_c3 $B > $O/ncu_full.log 2>&1
fi
# one RK step of every BASELINE configuration under ncu (sections, not --set full): every specialised kernel once
K="python tools/bench_configs.py --steps 1 --warmup 0 --reps 1 --skip C1"
$K > $O/cfg1.log 2>&1 && ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section SchedulerStats \
    --section WarpStateStats --section LaunchStats --section Occupancy --clock-control none -k regex:stage_tiled -c 30 --csv --page raw \
    --log-file $O/ncu_kernels.csv $K > $O/ncu_kernels.log 2>&1
