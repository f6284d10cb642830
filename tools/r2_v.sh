#!/bin/bash
O=gpurun_out
for v in "" atp1 atp1p; do
  echo -n "attention variant '${v:-base}': " | tee -a $O/v_attn.log
  KW_LIB_VARIANT=$v timeout 200 python tools/bench_kernels.py attn 2>&1 | tail -1 | tee -a $O/v_attn.log
done
for v in "" atp1p ""; do
  echo -n "attention variant '${v:-base}': " | tee -a $O/v_attn.log
  KW_LIB_VARIANT=$v timeout 200 python tools/bench_kernels.py attn 2>&1 | tail -1 | tee -a $O/v_attn.log
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_reduce tools/micro/tma_reduce_bench.cu && /tmp/tma_reduce 2>&1 | tee $O/v_tma_reduce_bench.log
timeout 900 compute-sanitizer --tool memcheck python tools/prof_step.py --batch 2 --max-length 6 --enc-layers 1 > $O/v_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -8 $O/v_sanitizer.log
