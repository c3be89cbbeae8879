#!/bin/bash
# sweep of the conv_tc launch configuration (shared memory per CTA, N tile, accumulators)
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 4 --warmup 3 --precision $PREC --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['dtype'], round(d['ms_per_step'],2),'ms', round(d['value'],1),'img/s')"; }
PREC=tf32
run PXL_TC_SMEM_KB=200
run PXL_TC_SMEM_KB=100
run PXL_TC_SMEM_KB=100 PXL_TC_BN_MAX_TF32=128
PREC=tf32x3
run PXL_TC_SMEM_KB=200
run PXL_TC_SMEM_KB=100 PXL_TC_BN_MAX_TF32X3=64 PXL_TC_NACC_TF32X3=4
run PXL_TC_SMEM_KB=100 PXL_TC_BN_MAX_TF32X3=64 PXL_TC_NACC_TF32X3=2
run PXL_TC_SMEM_KB=200 PXL_TC_BN_MAX_TF32X3=64 PXL_TC_NACC_TF32X3=4
run PXL_TC_SMEM_KB=200 PXL_TC_NACC_TF32X3=2
