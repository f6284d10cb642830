#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu > $O/aa_gemm.log 2>&1; echo "gemm rc=$?"; tail -2 $O/aa_gemm.log
echo "--- KW_STORE_TMA=1"; timeout 200 python tools/bench_kernels.py gemm 2>&1 | tee $O/aa_kern1.log
echo "--- KW_STORE_TMA=0"; KW_STORE_TMA=0 timeout 200 python tools/bench_kernels.py gemm 2>&1 | tee $O/aa_kern0.log
