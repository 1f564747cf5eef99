import csv, sys, subprocess
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_red.sum','lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print("kernel:", r[hdr.index('Kernel Name')][:60])
    for i,h in enumerate(hdr):
        if h in keys: print("  ",h, rows[1][i], r[i])
    for i,h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(r[i] or 0)>0.3: print("   stall",h[34:-23], r[i])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]
ia=hdr.index("Address"); isrc=hdr.index("Source"); iall=hdr.index("Warp Stall Sampling (All Samples)"); iex=hdr.index("Instructions Executed")
data=[]
for r in rows[2:]:
    try: data.append((int(r[iall] or 0), int(r[iex] or 0), r[isrc]))
    except Exception: pass
tot=sum(d[0] for d in data)
print("samples",tot)
for d in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 16]:
    print(f"{100*d[0]/tot:5.1f}% ex={d[1]:9d} {d[2][:100]}")
