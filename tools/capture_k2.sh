#!/bin/bash
# Refresh of the tail kernel's evidence only (the other kernels: tools/capture_profiles.sh): ncu --set full of one K2 launch
# in the bench workload, the launch list of a bench run, and the bench line itself (not under ncu).
P=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-parity --no-strong --no-wideband"
KERNELS='decim1|tail_kernel|fft_afc|stats_snap|init_cfg|mid_stage|ssdv|nco_mix|carry_kernel'
python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 400 --csv --log-file gpurun_out/${P}_launches_step.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity --no-strong --no-wideband > gpurun_out/${P}_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tail_kernel -s 6 -c 1 -o gpurun_out/${P}_k2 -f $B > /dev/null 2>&1
python tools/ncu_lines.py gpurun_out/${P}_k2.ncu-rep 25 > gpurun_out/${P}_k2_summary.txt 2>&1
ncu -i gpurun_out/${P}_k2.ncu-rep --page raw --csv > gpurun_out/${P}_k2_ncu_full_raw.csv 2>/dev/null
tail -c 600 gpurun_out/${P}_bench_n1.json | head -c 600; echo
head -12 gpurun_out/${P}_k2_summary.txt
