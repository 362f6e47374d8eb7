#!/bin/bash
# Round profile pass on a B200 box (run through gpurun from the repo root):
#   bash tools/profile_round.sh [nofull]     -> everything lands in gpurun_out/, summaries are copied to profiles/ by hand
# Every ncu command runs only after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
set -x
O=gpurun_out
R=${ROUND:-r02}
timeout 600 python bench.py > $O/bench_c3_1gpu_$R.json 2> $O/bench_c3_1gpu_$R.err || exit 1
timeout 900 python tools/bench_configs.py > $O/configs_$R.jsonl 2> $O/configs_$R.err || exit 1
B="python bench.py --n 512 --steps 2 --warmup 3 --no-cpu --no-e2e"
timeout 300 $B > $O/plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
    --log-file $O/launches_$R.csv $B > $O/ncu_launches.log 2>&1
if [ "$1" != "nofull" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair3d -s 6 -c 3 -f -o $O/prof_c3_$R $B > $O/ncu_full.log 2>&1
fi
# one RK step of every BASELINE configuration under ncu (sections, not --set full): every specialised kernel once
K="python tools/bench_configs.py --steps 1 --warmup 0 --reps 1 --skip C1"
timeout 600 $K > $O/cfg1.log 2>&1 && timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section SchedulerStats \
    --section WarpStateStats --section LaunchStats --section Occupancy --clock-control none -k "regex:stage_tiled|pair3d" -c 30 --csv --page raw \
    --log-file $O/ncu_kernels_$R.csv $K > $O/ncu_kernels.log 2>&1
# the resident cluster kernel of small 2-D grids (C1): timing, then one full capture of the 200-step launch
timeout 120 python tools/bench_configs.py --only C1 --reps 5 > $O/c1_resident.jsonl 2>&1
K1="python tools/bench_configs.py --only C1 --steps 20 --warmup 0 --reps 1"
timeout 120 $K1 > $O/c1_plain.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:resident2d -c 1 -f \
    -o $O/prof_resident_$R $K1 > $O/ncu_resident.log 2>&1
# the register-only FP64 floor of the WENO5 stage (build: see the header of tools/weno_floor.cu)
[ -x tools/weno_floor ] && timeout 60 tools/weno_floor > $O/weno_floor_$R.txt
# read here:  ncu -i X.ncu-rep --page raw --csv ;  ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_opmix.py src.csv NODE_STAGES
