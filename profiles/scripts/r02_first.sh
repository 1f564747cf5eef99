#!/bin/bash
# round 2, first GPU call: parity of both forward kernel families on BASELINE's shapes, then A/B timing
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests/test_baseline_shapes_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/t_baseline.log
timeout 600 python profiles/scripts/flow_sweep.py 1 > gpurun_out/sweep_pipe.log 2>&1
timeout 600 python profiles/scripts/flow_sweep.py 2 > gpurun_out/sweep_owner.log 2>&1
timeout 600 python profiles/scripts/flow_sweep.py 2 $((17*1024*1024)) > gpurun_out/sweep_owner_g1.log 2>&1
timeout 600 python profiles/scripts/flow_sweep.py 2 $((100*1024*1024)) > gpurun_out/sweep_owner_g6.log 2>&1
DCB_TEST_FWD_PATH=2 timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/t_all_owner.log
tail -5 gpurun_out/*.log
