#!/bin/bash
O=gpurun_out/r2k.log
: > $O
python -m pytest tests/test_surface_gpu.py -x -q -k "stream" >> $O 2>&1
echo "== attention q-tail trim" >> $O; python tools/bench_kernels.py attn >> $O 2>&1
echo "== attention no trim" >> $O; KW_LIB_VARIANT=notrim python tools/bench_kernels.py attn >> $O 2>&1
python -m pytest tests/test_attention_tc_gpu.py -x -q >> $O 2>&1
echo "== stream vs plain" >> $O; python tools/time_stream.py 6 >> $O 2>&1
grep -v Warn $O
