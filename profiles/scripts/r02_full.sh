#!/bin/bash
# full GPU check: the -m gpu suite with the default dispatch, then the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
rm -f gpurun_out/parity_report.jsonl
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err
timeout 300 python profiles/scripts/r02_small.py > gpurun_out/small_latency.log 2>&1; tail -5 gpurun_out/t_all.log; cat gpurun_out/small_latency.log; head -c 1500 gpurun_out/bench_now.json; tail -3 gpurun_out/bench_now.err
