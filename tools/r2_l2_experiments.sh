#!/bin/bash
# A/B of the decode step's L2 management knobs (api.cu decode_hidden): greedy pass alone, B = 64, same box.
O=gpurun_out/r2i_l2.log
: > $O
run() { echo "== $*" >> $O; env "$@" python tools/time_decode.py 64 8 >> $O 2>&1; }
run KW_XA_HINT=0
run KW_XA_HINT=1
run KW_XA_HINT=1 KW_W_HINT=1
run KW_XA_HINT=1 KW_W_HINT=1 KW_VOCAB_HINT=2
run KW_XA_HINT=2 KW_XA_ROWS=64
run KW_XA_HINT=2 KW_XA_ROWS=128
run KW_XA_HINT=2 KW_XA_ROWS=192
run KW_XA_HINT=3 KW_XA_ROWS=192 KW_L2PF_MASK=63
run KW_XA_HINT=3 KW_XA_ROWS=288 KW_L2PF_MASK=63
run KW_XA_HINT=3 KW_XA_ROWS=192 KW_L2PF_MASK=7
run KW_XA_HINT=3 KW_XA_ROWS=192 KW_L2PF_MASK=56
run KW_XA_HINT=3 KW_XA_ROWS=192 KW_L2PF_MASK=63 KW_VOCAB_HINT=2
run KW_XA_HINT=0 KW_XA_ROWS=192 KW_L2PF_MASK=63
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu tools/micro/mufu_x2_bench.cu && /tmp/mufu > gpurun_out/r2i_mufu_x2.log 2>&1
python - >> $O 2>&1 <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persisting max", getattr(p, "persisting_l2_cache_max_size", None))
PY
cat $O
