#!/usr/bin/env python
"""Per-source-line instruction counts / lane efficiency / stall samples from an ncu --set full capture
(compiled with -lineinfo).  usage: python tools/ncu_lines.py prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; agg = {}; kernel_seen = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[2] != "-":   # sass rows (address set) are children; the cuda line row has "-" address
        continue
    d = dict(zip(hdr, r))
    # duplicate "Source" column names: csv dict keeps the last; use indices
    line, src = r[0], r[1]
    try:
        inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"]); samp = int(d["# Samples"])
    except ValueError:
        continue
    key = (cur_file, int(line))
    a = agg.setdefault(key, [0, 0, 0, src.strip()])
    a[0] += inst; a[1] += thr; a[2] += samp
tot_i = sum(a[0] for a in agg.values()); tot_t = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot_i:,}  thread-inst {tot_t:,}  avg lanes {tot_t / max(tot_i,1):.2f}  samples {tot_s}")
print(f"{'file:line':24s} {'warp-inst%':>10s} {'lanes':>6s} {'samples%':>9s}  source")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f + ':' + str(l):24s} {100 * a[0] / tot_i:9.2f}% {a[1] / max(a[0],1):6.1f} {100 * a[2] / max(tot_s,1):8.2f}%  {a[3][:110]}")
