"""Top stalled SASS instructions of a kernel in an .ncu-rep (source page). usage: ncu_hot.py rep kernel_regex [n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
I = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > I["# Samples"]]
tot = sum(int(r[I["# Samples"]]) for r in body)
texec = sum(int(r[I["Instructions Executed"]]) for r in body)
print("total samples", tot, "warp instructions executed", texec, "static instrs", len(body))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body_sorted = sorted(enumerate(body), key=lambda t: -int(t[1][I["# Samples"]]))[:n]
for idx, r in sorted(body_sorted):
    s = int(r[I["# Samples"]])
    top = sorted(((int(r[I[h]]), h) for h in stalls), reverse=True)[:2]
    print(f"{idx:5d} {100*s/tot:5.1f}%  exec {int(r[I['Instructions Executed']]):>10d}  {r[I['Source']].strip()[:70]:70s} {top}")
