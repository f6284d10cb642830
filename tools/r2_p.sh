#!/bin/bash
# round-2 session 5, call 1: 128-row decode GEMMs + coalesced stream: correctness, pass timing, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu > gpurun_out/p_gemm.log 2>&1; echo "gemm rc=$?" 
tail -3 gpurun_out/p_gemm.log
timeout 600 python -m pytest tests/test_surface_gpu.py -x -q -m gpu -k "coalesce or stream" > gpurun_out/p_surface.log 2>&1; echo "surface rc=$?"
tail -3 gpurun_out/p_surface.log
timeout 300 python tools/time_decode.py 64 8 2>&1 | tail -1 | tee gpurun_out/p_decode64.log
timeout 300 python tools/time_decode.py 128 8 2>&1 | tail -1 | tee gpurun_out/p_decode128.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/p_bench_co2.json 2> gpurun_out/p_bench_co2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/p_bench_co2.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'plain', d['config']['ms_per_step_batch_by_batch'], 'e2e', d['e2e']['value'])
    for e in d['roofline_extra']:
        print(' ', e['kernel'][:60], e.get('frac'), e.get('ms_per_position'))
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -5 gpurun_out/p_bench_co2.err
