#!/bin/bash
# Final round-2 validation + captures on the GPU box: gpurun -- 'bash profiles/final_r2.sh'.  Everything lands in gpurun_out/.
O=gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > $O/y_tests.log 2>&1; echo tests rc=$?; tail -3 $O/y_tests.log
timeout 200 bash profiles/ncu_r2.sh > $O/y_ncu.log 2>&1; echo ncu rc=$?
timeout 100 python bench.py --dtype f64 --steps 2 --warmup 3 > $O/y_bench_f64.json 2> $O/y_bench_f64.err; echo f64 rc=$?
tail -c 600 $O/y_bench_f64.json
