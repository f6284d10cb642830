#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_gemm_tc_gpu.py tests/test_attention_tc_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > $O/w_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/w_tests.log
for v in 0 1 0 1; do
  KW_L2_ORDER=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/w_bench.json 2> $O/w_bench.err
  python -c "
import json
d=json.loads(open('$O/w_bench.json').read().strip().splitlines()[-1])
x=[e for e in d['roofline_extra'] if e['kernel'].startswith('encoder self')][0]
print('l2_order=$v value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'plain', round(d['config']['ms_per_step_batch_by_batch'],2), 'stream1', round(d['config']['ms_per_step_stream_coalesce1'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), 'attn TF', round(x['achieved']), 'clk', d['clocks']['sm_mhz'])
" | tee -a $O/w_ab.log
done
tail -3 $O/w_bench.err
