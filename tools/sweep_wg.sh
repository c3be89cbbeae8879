#!/bin/bash
out=gpurun_out/sweep_wg.txt
: > $out
for b in 0 2 6 16; do echo "=== PXL_WG_WIDE_BIAS=$b" >> $out; PXL_WG_WIDE_BIAS=$b timeout 120 python tools/bench_conv.py f16x3 wgrad 2>&1 | grep -E "3x3|weighted" >> $out; done
