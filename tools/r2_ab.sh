#!/bin/bash
O=gpurun_out
for v in 0 1 0 1; do
  KW_STORE_TMA=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/ab_bench.json 2> $O/ab_bench.err
  python -c "
import json
d=json.loads(open('$O/ab_bench.json').read().strip().splitlines()[-1])
print('store_tma=$v value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'plain', round(d['config']['ms_per_step_batch_by_batch'],2), 'stream1', round(d['config']['ms_per_step_stream_coalesce1'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])
" | tee -a $O/ab_store_tma.log
done
