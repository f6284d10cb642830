#!/bin/bash
O=gpurun_out/r2l.log
: > $O
python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?" >> $O
for g in 8 4 16; do echo "== KW_GRAPH_POS=$g" >> $O; KW_GRAPH_POS=$g python tools/time_stream.py 6 >> $O 2>&1; done
grep -v Warn $O; cut -c1-1500 gpurun_out/r2l_bench.json; tail -5 gpurun_out/r2l_bench.err
