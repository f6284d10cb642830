#!/bin/bash
O=gpurun_out/r2m.log
: > $O
for g in 8 4 2; do echo "== KW_GRAPH_POS=$g pass alone" >> $O; KW_GRAPH_POS=$g python tools/time_decode.py 64 6 >> $O 2>&1; echo "== KW_GRAPH_POS=$g stream" >> $O; KW_GRAPH_POS=$g python tools/time_stream.py 6 >> $O 2>&1; done
grep -v Warn $O
