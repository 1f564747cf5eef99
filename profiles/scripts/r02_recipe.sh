#!/bin/bash
# per-launch numbers of the two recipe passes (pass A: mask scatter + mask epilogue; pass B: rider scatter + recipe epilogue)
cd "$GRAFT_REPO_ROOT" || exit 1
python profiles/scripts/run_recipe.py 8 3 > gpurun_out/recipe_plain.log 2>&1 || { cat gpurun_out/recipe_plain.log; exit 1; }
cat gpurun_out/recipe_plain.log
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_splat_step" -s 32 -c 32 --csv --log-file gpurun_out/recipe_ncu.csv python profiles/scripts/run_recipe.py 8 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/recipe_ncu.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = r[vi]
agg = collections.OrderedDict()
for i, m in per.items():
    kind = ("B" if "1>(" in m["k"] else "A") + (" scatter" if float(m["lts__t_sectors_srcunit_tex_op_red.sum"].replace(",", "")) > 0 else " epilogue")
    a = agg.setdefault(kind, collections.Counter())
    a["n"] += 1
    for k, v in m.items():
        if k != "k": a[k] += float(v.replace(",", ""))
for kind, a in agg.items():
    n = a.pop("n")
    print(kind, f"x{n:.0f}", " ".join(f"{k.split('.')[0].split('__')[1]}={v / n:.4g}" for k, v in a.items()))
PY
