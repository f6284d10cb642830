#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): [GPU tests, smoke,] bench line, configs, launch lists,
# --set full capture (summarised on the box: the report itself does not fit the 64 MiB return channel).
# usage: bash tools/r2_measure.sh TAG [tests]
set -x
O=gpurun_out
T=${1:-r2n}
if [ "$2" = "tests" ]; then
  python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
  python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log
fi
python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
python tools/bench_configs.py cfg3 > $O/${T}_configs_cfg3.json 2> $O/${T}_configs.err
python tools/prof_step.py --stream > $O/${T}_prof_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file $O/${T}_launches_stream.csv python tools/prof_step.py --stream > $O/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"skinny|dec_cross|sample_combine|dec_self|layernorm" -c 60 -o /tmp/${T}_decode_full python tools/prof_step.py --max-length 8 --enc-layers 2 > $O/${T}_ncu2.log 2>&1
python tools/ncu_summary.py /tmp/${T}_decode_full.ncu-rep > $O/${T}_ncu_decode_kernels.txt 2>&1
ls -la /tmp/${T}_decode_full.ncu-rep >> $O/${T}_ncu2.log
du -sh $O
