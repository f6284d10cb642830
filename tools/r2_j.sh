#!/bin/bash
O=gpurun_out/r2j.log
: > $O
echo "== attention, trim build" >> $O; python tools/bench_kernels.py attn >> $O 2>&1
echo "== attention, no-trim build" >> $O; KW_LIB_VARIANT=notrim python tools/bench_kernels.py attn >> $O 2>&1
echo "== attention, trim build" >> $O; python tools/bench_kernels.py attn >> $O 2>&1
python -m pytest tests/test_attention_tc_gpu.py -x -q >> $O 2>&1
echo "== decode hints default" >> $O; python tools/time_decode.py 64 8 >> $O 2>&1
echo "== decode hints off" >> $O; KW_XA_HINT=0 KW_W_HINT=0 KW_VOCAB_HINT=0 python tools/time_decode.py 64 8 >> $O 2>&1
echo "== interleave probe" >> $O; python tools/interleave_probe.py >> $O 2>&1
cat $O | grep -v Warn
