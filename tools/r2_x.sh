#!/bin/bash
# final state of the session: GPU tests, smoke, full bench line
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/x_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/x_pytest.log; tail -3 $O/x_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/x_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/x_smoke.log; tail -2 $O/x_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/x_bench.json 2> $O/x_bench.err; echo "bench rc=$?"
head -c 400 $O/x_bench.json
