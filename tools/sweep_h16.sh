#!/bin/bash
# tile / accumulator / CTA-pair sweep of the kind::f16 convolution kernels (tools/bench_conv.py); knobs are read once
# per process by the library, so every configuration is its own run.  Output: gpurun_out/sweep_h16.txt
out=gpurun_out/sweep_h16.txt
: > $out
run() { echo "=== $*" >> $out; env "$@" timeout 120 python tools/bench_conv.py $MODE $WHAT 2>&1 | grep -v "^precision" >> $out; }
MODE=f16x3; WHAT=fwd
run PXL_TC_BN_MAX_F16X3=256 PXL_TC_NACC_F16X3=1
run PXL_TC_PAIR=1 PXL_TC_BN_MAX_F16X3=256 PXL_TC_NACC_F16X3=1
run PXL_TC_PAIR=1 PXL_TC_BN_MAX_F16X3=128 PXL_TC_NACC_F16X3=2
run PXL_TC_PERSIST=0 PXL_TC_SMEM_KB=100 PXL_TC_NACC_F16X3=2
WHAT=wgrad
run PXL_WG_BN_MAX_F16X3=256 PXL_WG_ROWS_F16X3=32
run PXL_WG_BN_MAX_F16X3=256 PXL_WG_ROWS_F16X3=64
run PXL_WG_BN_MAX_F16X3=128 PXL_WG_ROWS_F16X3=32
run PXL_WG_BN_MAX_F16X3=128 PXL_WG_ROWS_F16X3=96
MODE=f16; WHAT=fwd
run PXL_TC_PAIR=1
run PXL_TC_PERSIST=1 PXL_TC_BN_MAX_F16=128
WHAT=wgrad
run PXL_WG_ROWS_F16=128
run PXL_WG_ROWS_F16=32
run PXL_WG_BN_MAX_F16=128 PXL_WG_ROWS_F16=128
