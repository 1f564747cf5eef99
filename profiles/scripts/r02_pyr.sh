#!/bin/bash
# pyramid path: its tests, then the timing script
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_conditioning_gpu.py tests/test_abi_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/t_pyr.log
timeout 300 python profiles/scripts/run_pyramid_fused.py > gpurun_out/pyr_fused.log 2>&1
cat gpurun_out/t_pyr.log; cat gpurun_out/pyr_fused.log
