#!/bin/bash
# decode attribution at 128 rows: leave one kernel kind out of the chain (KW_DECODE_SKIP), greedy pass alone
O=gpurun_out
for s in 0 1 2 4 8 16 32 64 128 256 512; do
  echo -n "skip=$s  " | tee -a $O/q_attr128.log
  KW_DECODE_SKIP=$s timeout 200 python tools/time_decode.py 128 6 2>&1 | tail -1 | tee -a $O/q_attr128.log
done
