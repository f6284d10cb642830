#!/bin/bash
# PDL mask at 128 rows (default 3 = LayerNorm + self-attention), and split-K / GEMM PDL knobs
O=gpurun_out
for m in 3 1 0 7 19 17; do
  echo -n "KW_PDL_MASK=$m  " | tee -a $O/y_pdl128.log
  KW_PDL_MASK=$m timeout 200 python tools/time_decode.py 128 6 2>&1 | tail -1 | tee -a $O/y_pdl128.log
done
echo -n "KW_SPLITK=2  " | tee -a $O/y_pdl128.log
KW_SPLITK=2 timeout 200 python tools/time_decode.py 128 6 2>&1 | tail -1 | tee -a $O/y_pdl128.log
