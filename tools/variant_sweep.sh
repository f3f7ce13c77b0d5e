#!/bin/bash
# bench every tuning variant of the library (habdec_b200/variant_*.so) with 1 and 2 channel groups
for v in habdec_b200/libhabdec_b200.so habdec_b200/variant_*.so; do
  for g in ${GROUPS_LIST:-1 2}; do
    HBD_LIB=$PWD/$v HBD_GROUPS=$g timeout 300 python bench.py --steps ${STEPS:-60} --warmup 3 --chunk ${CHUNK:-65536} --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python tools/summarize_bench.py "$(basename $v) groups=$g" | cut -c1-230
  done
done
