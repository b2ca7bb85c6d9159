"""Per-opcode and per-region stall samples / executed instructions of one kernel from an .ncu-rep.
usage: ncu_source.py rep kernel_regex [launch_skip]"""
import csv, subprocess, sys, io
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
data = []
for r in rows[rows.index(hdr) + 1:]:
    if len(r) != len(hdr) or not r[hdr.index("# Samples")].isdigit():
        if data: break
        continue
    data.append(r)
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot_s = sum(int(r[iS]) for r in data); tot_i = sum(int(r[iI]) for r in data)
print("static instr", len(data), "samples", tot_s, "warp-instr executed", tot_i)
cs = Counter(); ci = Counter()
for r in data:
    toks = [t for t in r[iSrc].split() if not t.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    cs[op] += int(r[iS]); ci[op] += int(r[iI])
for op, s in cs.most_common(20):
    print(f"{op:10s} samples {s:7d} ({100 * s / max(tot_s, 1):5.1f}%)  instr {ci[op]:11d} ({100 * ci[op] / max(tot_i, 1):5.1f}%)")
step = max(len(data) // 24, 1)
print("regions (static index, samples, executed, first instr):")
for b in range(0, len(data), step):
    blk = data[b:b + step]
    print(f"{b:6d} {sum(int(r[iS]) for r in blk):7d} {sum(int(r[iI]) for r in blk):11d}  {blk[0][iSrc].strip()[:60]}")
