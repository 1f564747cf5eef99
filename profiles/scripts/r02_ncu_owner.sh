#!/bin/bash
# ncu --set full of the target-tile-owner kernels on 4 frames of the bench workload
cd "$GRAFT_REPO_ROOT" || exit 1
export DCB_FWD_PATH=2
python profiles/scripts/run_fwd.py 4 soft 3 > gpurun_out/plain_owner.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_splat_owner|k_strip_box" -s 4 -c 4 -o gpurun_out/owner_${1:-v1} -f python profiles/scripts/run_fwd.py 4 soft 3 > gpurun_out/ncu_owner.log 2>&1
tail -3 gpurun_out/plain_owner.log gpurun_out/ncu_owner.log
