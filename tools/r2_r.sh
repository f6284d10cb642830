#!/bin/bash
# TMA-reduce residual epilogue: correctness, A/B per kernel, A/B on the stream step
O=gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu > $O/r_gemm.log 2>&1; echo "gemm rc=$?"; tail -2 $O/r_gemm.log
echo "--- KW_RESID_TMA=1"; timeout 200 python tools/bench_kernels.py gemm 2>&1 | tee $O/r_kern_tma1.log
echo "--- KW_RESID_TMA=0"; KW_RESID_TMA=0 timeout 200 python tools/bench_kernels.py gemm 2>&1 | tee $O/r_kern_tma0.log
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "kotoba_fp32 or tiny_bf16 or bf16_tensor" > $O/r_model.log 2>&1; echo "model rc=$?"; tail -2 $O/r_model.log
for v in 1 0; do
  KW_RESID_TMA=$v timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-parity > $O/r_bench_tma$v.json 2> $O/r_bench_tma$v.err; echo "bench tma=$v rc=$?"
  python -c "
import json
d=json.loads(open('$O/r_bench_tma$v.json').read().strip().splitlines()[-1])
print('tma=$v value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'plain', round(d['config']['ms_per_step_batch_by_batch'],2), 'e2e', round(d['e2e']['value']), 'gemm TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])
"
done
