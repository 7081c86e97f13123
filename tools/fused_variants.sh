#!/bin/bash
# Times the variants of the fused tile kernel (CVB_FUSED = old | 0 | 1 | 2) on resident 1080p frames and checks each
# against the oracle on a few shapes.  usage: tools/fused_variants.sh [frames] [variants...]
n=${1:-32}; shift
vars=${@:-old 0 1 2}
mkdir -p gpurun_out
for v in $vars; do
  echo "=== CVB_FUSED=$v"
  CVB_FUSED=$v CVB_CHECK=1 timeout 600 python tools/prof_run.py $n 3 2>&1 | grep -E "k_fused|total|parity|Error|error|Traceback" 
done
