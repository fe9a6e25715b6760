"""Top SASS instructions by warp-stall samples from an ncu report (source page), per kernel.
usage: python profiles/tools/hot_sass.py rep.ncu-rep [kernel-regex] [topN]"""
import csv, subprocess, sys, re
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else "."; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if not row: continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
for b in blocks:
    if not re.search(rx, b["name"]): continue
    h = b["hdr"]; si = h.index("Source"); ci = h.index("Warp Stall Sampling (All Samples)"); ei = h.index("Instructions Executed")
    tot = sum(int(r[ci] or 0) for r in b["rows"])
    print(f"== {b['name'][:120]}  total samples {tot}")
    idx = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][ci] or 0))[:top]
    for i in sorted(idx):
        r = b["rows"][i]
        print(f"{i:6d} {int(r[ci] or 0):7d} {100*int(r[ci] or 0)/max(tot,1):5.1f}%  exec={r[ei]:>8s}  {r[si].strip()[:110]}")
