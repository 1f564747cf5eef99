#!/bin/bash
# iteration loop for the owner kernels: parity of the owner-specific tests, then the roughness sweep on both kernel families
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_baseline_shapes_gpu.py -x -q -k "owner" 2>&1 | tail -12 > gpurun_out/t_owner.log
timeout 300 python profiles/scripts/flow_sweep.py 2 > gpurun_out/sweep_owner.log 2>&1
cat gpurun_out/t_owner.log gpurun_out/sweep_owner.log
