#!/bin/bash
# steady-state DRAM traffic per launch (ncu --cache-control none) of the two-slot ring vs one slot with alternating launches
cd "$GRAFT_REPO_ROOT" || exit 1
for slots in 2 1; do
  export DCB_PIPE_RING_SLOTS=$slots
  echo "=== ring slots $slots"
  python profiles/scripts/run_fwd.py 8 soft 3 > /dev/null 2>&1 &&
  ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_splat_step" -s 20 -c 4 --csv python profiles/scripts/run_fwd.py 8 soft 3 2>/dev/null | grep -E "k_splat" | awk -F'","' '{print $(NF-2), $(NF)}' | tr '\n' ' ' | sed 's/gpu__time/\ngpu__time/g'
  echo
done
