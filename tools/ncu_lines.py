"""Per-source-line instruction counts / stall samples of one kernel from an .ncu-rep (needs -lineinfo).
usage: python tools/ncu_lines.py <rep> <kernel-regex> [top-N]"""
import csv, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find the first header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
iex = hdr.index("Instructions Executed"); smp = hdr.index("# Samples")
line_inst = collections.Counter(); line_smp = collections.Counter(); src = {}
cur = None; fname = "?"
for r in rows:
    if not r or r[0] in ("Line No", "Function Name", "Kernel Name", "File Name"):
        if r and r[0] == "File Name":
            fname = r[1].split("/")[-1]
        continue
    if r[0] != "":
        cur = fname + ":" + r[0]; src[cur] = r[1]
    if len(r) > iex and r[2] != "" and cur is not None:
        try:
            line_inst[cur] += int(r[iex].replace(",", "")); line_smp[cur] += int(r[smp].replace(",", "") or 0)
        except ValueError:
            pass
tot = sum(line_inst.values()); ts = sum(line_smp.values())
print(f"total warp-instructions {tot:,}  samples {ts:,}")
for ln, n in line_inst.most_common(top):
    print(f"{ln:>22s} {100*n/tot:5.1f}% inst {100*line_smp[ln]/max(ts,1):5.1f}% smp  {src[ln].strip()[:110]}")
