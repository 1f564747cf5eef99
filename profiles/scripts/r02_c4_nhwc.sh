#!/bin/bash
# per-kernel numbers of the list path at C4 (8 x 64 x 256 x 256): NCHW gather vs channels-last gather
cd "$GRAFT_REPO_ROOT" || exit 1
python profiles/scripts/run_c4_nhwc.py > gpurun_out/c4_nhwc_plain.log 2>&1 || { cat gpurun_out/c4_nhwc_plain.log; exit 1; }
cat gpurun_out/c4_nhwc_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio --clock-control none -k regex:"gather_nhwc" -s 30 -c 2 --csv --log-file gpurun_out/c4_nhwc_ncu.csv python profiles/scripts/run_c4_nhwc.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/c4_nhwc_ncu.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = r[vi]
for i, m in per.items():
    print(i, m["k"][:48], " ".join(f"{k.split('.')[0].split('__')[1][:22]}={v}" for k, v in m.items() if k != "k"))
PY
