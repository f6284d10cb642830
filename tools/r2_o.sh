#!/bin/bash
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64 tools/micro/fp64_bench.cu && /tmp/fp64 > $O/r2o_fp64.log 2>&1
python -m pytest tests/test_logmel_gpu.py tests/test_pipeline_gpu.py -x -q > $O/r2o_logmel_tests.log 2>&1
python tools/bench_kernels.py logmel > $O/r2o_logmel_bench.log 2>&1
python tools/bench_configs.py cfg5 > $O/r2o_cfg5.json 2>> $O/r2o_logmel_bench.log
cat $O/r2o_fp64.log; tail -3 $O/r2o_logmel_tests.log; cat $O/r2o_logmel_bench.log
bash tools/r2_measure.sh r2n
