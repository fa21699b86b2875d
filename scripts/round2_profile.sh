#!/bin/bash
# ncu --set full captures of the stage kernel for one configuration, exported to the CSV pages the summary scripts read.
#   gpurun --timeout 900 -- 'bash scripts/round2_profile.sh <tag> [bench.py flags...]'
# then here:  python scripts/ncu_raw_summary.py gpurun_out/<tag>_raw.csv <tag> ; python scripts/ncu_src_summary.py gpurun_out/<tag>_src.csv 524288 0|2 ; python scripts/ncu_src_segments.py ...
set -u
tag=$1; shift
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mhd_stage_xy -s 6 -c 2 -f -o gpurun_out/${tag} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv 2>/dev/null
rm -f gpurun_out/${tag}.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2>&1
ls -la gpurun_out/${tag}* | awk '{print $5, $9}'
