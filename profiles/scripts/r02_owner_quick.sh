#!/bin/bash
# quick look at the owner kernels: parity of the owner tests, roughness sweep per library variant,
# instruction count / duration per launch (light ncu pass on the in-tree build)
# args: library paths to compare ("default" = the in-tree build)
cd "$GRAFT_REPO_ROOT" || exit 1
export DCB_FWD_PATH=2
for lib in "$@"; do
  export DCB_LIB_PATH="$lib"
  [ "$lib" = default ] && unset DCB_LIB_PATH
  echo "=== $lib"
  timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py -x -q -k "owner and not recipe and not backward" 2>&1 | tail -2
  timeout 300 python profiles/scripts/flow_sweep.py 2 2>&1 | grep -E "noise cell +(32|256|0) "
done
unset DCB_LIB_PATH
python profiles/scripts/run_fwd.py 4 soft 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_splat_owner|k_strip_box" -s 4 -c 4 --csv python profiles/scripts/run_fwd.py 4 soft 3 2>/dev/null | grep -E "k_splat|k_strip" | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-200
