#!/bin/bash
# The ncu evidence kept under profiles/ (B200_PROFILING.md recipe), one GPU, bench workload (configs[1]):
#   1. launch list: per-launch device time of one step          -> <tag>_ncu_launches.csv
#   2. SpeedOfLight / memory / launch / occupancy sections       -> <tag>_ncu_step_sections.csv (via ncu_summarise.py)
#   3. --set full + source of one 512->512 3x3 head conv         -> <tag>_ncu_conv_head_full_raw.csv, _source_top.csv
tag=${1:-r02}
only=${2:-all}
o=gpurun_out
mkdir -p $o
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $o/${tag}_ncu_launches.csv python tools/prof_step.py 2> $o/ops.txt
ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second \
    --clock-control none -f -o $o/step_sections python tools/prof_step.py 2> $o/ops2.txt > /dev/null
ncu -i $o/step_sections.ncu-rep --page raw --csv > $o/step_sections_raw.csv
python tools/ncu_summarise.py $o/step_sections_raw.csv $o/ops.txt $o/${tag}_ncu_step_sections.csv
# stamp: the hash of the CUDA sources this capture was taken from (bench.py refuses a capture whose stamp is stale)
python - "$o/${tag}_ncu_step_sections.meta.json" <<'P'
import json, sys
sys.path.insert(0, ".")
import bench
json.dump({"csrc_sha": bench.csrc_sha(), "command": "tools/profile_round.sh", "workload": "bench.py default (configs[1])"}, open(sys.argv[1], "w"))
P
rm -f $o/step_sections.ncu-rep
[ "$only" = sections ] && exit 0
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_igemm -s 83 -c 1 \
    -f -o $o/conv_head_full python tools/prof_step.py 2> /dev/null > /dev/null
ncu -i $o/conv_head_full.ncu-rep --page raw --csv > $o/${tag}_ncu_conv_head_full_raw.csv
ncu -i $o/conv_head_full.ncu-rep --page source --csv > $o/conv_head_source.csv 2>/dev/null
rm -f $o/step_sections.ncu-rep
ls -la $o | tail -12
