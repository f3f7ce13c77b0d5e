#!/bin/bash
# one gpurun session: parity tests, then bench with 1/2/4 channel groups, then an ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for g in ${GROUPS_LIST:-1 2 4}; do
  HBD_GROUPS=$g timeout 600 python bench.py --steps ${STEPS:-40} --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python tools/summarize_bench.py "groups=$g"
done
if [ -n "$NCU_LIST" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decim1|tail_kernel|carry_kernel|fft_afc|slicer|mark_kernel|pack_raw|init_cfg" -c 48 --csv --log-file gpurun_out/$NCU_LIST python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
fi
