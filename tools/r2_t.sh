#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_surface_gpu.py tests/test_decode_fused_gpu.py -x -q -m gpu > $O/t_model.log 2>&1; echo "model+surface rc=$?"; tail -3 $O/t_model.log
timeout 300 python tools/time_decode.py 64 8 2>&1 | tail -1 | tee $O/t_decode64.log
timeout 300 python tools/time_decode.py 128 8 2>&1 | tail -1 | tee $O/t_decode128.log
