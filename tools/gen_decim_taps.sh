#!/bin/bash
# Regenerates habdec_b200/csrc/decim_taps.inc from the reference tap tables.
set -e
REF=${HABDEC_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
g++ -O0 -I "$REF/code/Decoder" "$HERE/tools/gen_decim_taps.cpp" -o /tmp/hbd_gen_taps
/tmp/hbd_gen_taps > "$HERE/habdec_b200/csrc/decim_taps.inc"
echo "wrote $HERE/habdec_b200/csrc/decim_taps.inc"
