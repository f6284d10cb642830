#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_surface_gpu.py -x -q -m gpu -k "stream or coalesce" > $O/u_surface.log 2>&1; echo "surface rc=$?"; tail -3 $O/u_surface.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for v in 0 1 0 1; do
  KW_STREAM_DEFER=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/u_bench.json 2> $O/u_bench.err
  python -c "
import json
d=json.loads(open('$O/u_bench.json').read().strip().splitlines()[-1])
print('defer=$v value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'plain', round(d['config']['ms_per_step_batch_by_batch'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])
" | tee -a $O/u_ab.log
done
tail -3 $O/u_bench.err
for gp in 4 16; do
  KW_GRAPH_POS=$gp timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/u_bench.json 2> $O/u_bench.err
  python -c "
import json
d=json.loads(open('$O/u_bench.json').read().strip().splitlines()[-1])
print('graph_pos=$gp value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'plain', round(d['config']['ms_per_step_batch_by_batch'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'gemm TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])
" | tee -a $O/u_ab.log
done
