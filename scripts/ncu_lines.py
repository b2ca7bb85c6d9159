"""Per-source-line instruction and stall-sample shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py src.csv [kernel-substring] [top=40]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
agg = {}     # kernel -> {(file,line,src): [inst, samples]}
f = fn = hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1].split("(")[0][-70:]; continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or r[0] == "" or len(r) <= ii: continue      # SASS rows have an empty line number
    try: inst = int(r[ii]); smp = int(r[si])
    except ValueError: continue
    d = agg.setdefault(fn, {})
    k = (f, int(r[0]), r[1].strip()[:110])
    v = d.setdefault(k, [0, 0]); v[0] += inst; v[1] += smp
for fn, d in agg.items():
    if want not in fn: continue
    ti = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values())
    print("=====", fn, "warp-instructions", ti, "samples", ts)
    for (f, ln, src), (i, s) in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f[:14]:>14}:{ln:<4} {100*i/ti:5.1f}% inst {100*s/max(ts,1):5.1f}% smp  {src}")
