#!/bin/bash
# A/B of accumulator-pipeline builds: parity of the BASELINE-shape tests (pipe family), roughness sweep, light ncu pass
# args: library paths ("default" = in-tree build)
cd "$GRAFT_REPO_ROOT" || exit 1
for lib in "$@"; do
  export DCB_LIB_PATH="$lib"
  [ "$lib" = default ] && unset DCB_LIB_PATH
  echo "=== $lib"
  [ -z "$NOTEST" ] && timeout 900 python -m pytest tests/test_baseline_shapes_gpu.py -x -q -k "pipe or ring_groups" 2>&1 | tail -2
  timeout 300 python profiles/scripts/flow_sweep.py 1 2>&1 | grep -E "noise cell +(32|256|0) "
done
unset DCB_LIB_PATH
python profiles/scripts/run_fwd.py 4 soft 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_splat_step" -s 6 -c 3 --csv python profiles/scripts/run_fwd.py 4 soft 3 2>/dev/null | grep -E "k_splat" | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-200
