#!/bin/bash
# Round-2 final measurement pass on one B200 (run under gpurun): GPU tests, smoke, bench line, launch list of the
# coalesced stream schedule, --set full capture of the 128-row decode kernels and the TMA-reduce GEMM epilogue,
# DRAM traffic of the wide GEMMs.  usage: bash tools/r2_final.sh TAG [tests]
set -x
O=gpurun_out
T=${1:-r2z}
if [ "$2" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
  timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log
  tail -3 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 600 python tools/prof_step.py --stream --coalesce 2 > $O/${T}_prof_plain.log 2>&1 && \
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/${T}_launches_stream_co2.csv python tools/prof_step.py --stream --coalesce 2 > $O/${T}_ncu1.log 2>&1
python tools/summarize_launches.py $O/${T}_launches_stream_co2.csv > $O/${T}_launches_stream_co2.txt 2>&1
gzip -f $O/${T}_launches_stream_co2.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"skinny|dec_cross|sample_combine|dec_self|gemm_tc2" -c 70 -o /tmp/${T}_full python tools/prof_step.py --batch 128 --max-length 8 --enc-layers 1 > $O/${T}_ncu2.log 2>&1
python tools/ncu_summary.py /tmp/${T}_full.ncu-rep > $O/${T}_ncu_kernels_b128.txt 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"gemm_tc2_kernel|gemm_tc_kernel" --csv --log-file $O/${T}_traffic.csv python tools/prof_step.py --enc-layers 4 --max-length 5 > $O/${T}_ncu3.log 2>&1
python tools/ncu_traffic.py $O/${T}_traffic.csv 4 64 $O/${T}_traffic.json > $O/${T}_traffic.log 2>&1
gzip -f $O/${T}_traffic.csv
du -sh $O
