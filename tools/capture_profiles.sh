#!/bin/bash
# One gpurun session that refreshes the evidence under gpurun_out/ (copied into profiles/ by hand afterwards):
# launch list of a bench step, ncu --set full of K1 / K2 / K4 / K1-NCO / K4-16384 with per-line summaries.
P=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-parity --no-strong --no-wideband"
KERNELS='decim1|tail_kernel|fft_afc|stats_snap|init_cfg|mid_stage|ssdv|nco_mix|carry_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 400 --csv --log-file gpurun_out/${P}_launches_step.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity --no-strong --no-wideband > gpurun_out/${P}_ncu_bench.log 2>&1
for k in decim1:k1 tail_kernel:k2; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:${k%%:*} -s 6 -c 1 -o gpurun_out/${P}_${k##*:} -f $B > /dev/null 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fft_afc4096 -s 0 -c 1 -o gpurun_out/${P}_k4 -f python bench.py --steps 3 --warmup 20 --no-e2e --no-cpu-baseline --no-parity --no-strong --no-wideband > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decim1 -s 6 -c 1 -o gpurun_out/${P}_k1nco -f python tools/bench_wideband.py --channels 4096 --steps 4 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fft_afc_kernel -c 1 -o gpurun_out/${P}_k4_16384 -f python -m pytest tests/test_gpu_parity.py -q -k "test_fft_and_afc and 16384 and 2500000" > /dev/null 2>&1
for r in k1 k2 k4 k1nco k4_16384; do
  [ -f gpurun_out/${P}_$r.ncu-rep ] && python tools/ncu_lines.py gpurun_out/${P}_$r.ncu-rep 25 > gpurun_out/${P}_${r}_summary.txt 2>&1
  [ -f gpurun_out/${P}_$r.ncu-rep ] && ncu -i gpurun_out/${P}_$r.ncu-rep --page raw --csv > gpurun_out/${P}_${r}_ncu_full_raw.csv 2>/dev/null
done
ls -la gpurun_out/${P}_*
