"""Per-source-line shares of executed instructions and stall samples from an ncu report (needs -lineinfo and
--import-source on). usage: ncu_lines.py report.ncu-rep kernel-regex [min-percent] [launch-id]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, out = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ex = int(r[hdr.index("Instructions Executed")] or 0)
        smp = int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0)
        key = (cur, int(r[0]))
        e = out.setdefault(key, [0, 0, r[1]])
        e[0] += ex; e[1] += smp
tot = sum(v[0] for v in out.values()) or 1; ts = sum(v[1] for v in out.values()) or 1
print("instructions executed", tot, "stall samples", ts)
for (f, ln), (ex, smp, src) in sorted(out.items()):
    if 100 * ex / tot >= minp or 100 * smp / ts >= minp:
        print(f"{f[:16]:16s} {ln:4d}  inst {100 * ex / tot:5.1f}%  stall {100 * smp / ts:5.1f}%  {src.strip()[:100]}")
