#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): GPU tests, bench lines (ours + reference arm), the other
# configs, the bf16 parity artifact, launch list.  Everything lands in gpurun_out/ and is copied to profiles/ by hand.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log
python bench.py --steps 5 --warmup 3 > $O/r2h_bench.json 2> $O/r2h_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2h_bench_reference.json 2> $O/r2h_bench_reference.err
python bench.py --impl reference --steps 1 --warmup 1 --ref-batch 8 > $O/r2h_bench_reference_b8.json 2>> $O/r2h_bench_reference.err
python tools/bench_configs.py > $O/r2h_configs.json 2> $O/r2h_configs.err
python tools/bf16_parity.py 128 $O/r2h_bf16_parity.json > $O/r2h_parity.log 2>&1
python tools/prof_step.py > $O/r2h_prof_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file $O/r2h_launches.csv python tools/prof_step.py > $O/r2h_ncu.log 2>&1
tail -3 $O/r2h_pytest.log; cat $O/r2h_bench.json | cut -c1-600
