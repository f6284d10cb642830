#!/bin/bash
# A/B of KW_RESID_TMA on the bench line, alternating, 20 steps each
O=gpurun_out
for v in 0 1 0 1; do
  KW_RESID_TMA=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/s_bench.json 2> $O/s_bench.err
  python -c "
import json
d=json.loads(open('$O/s_bench.json').read().strip().splitlines()[-1])
print('tma=$v value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'plain', round(d['config']['ms_per_step_batch_by_batch'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])
" | tee -a $O/s_ab.log
done
