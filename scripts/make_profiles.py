"""Turns what scripts/gpu_profile_round.sh brought back in gpurun_out/ into the tracked evidence under profiles/:
   python scripts/make_profiles.py r02
  profiles/<tag>_bench_line.json, _bench_reference.json, _configs.json     copies
  profiles/<tag>_launches.csv + _launch_list_summary.md                     per-kernel launches / time / share
  profiles/<tag>_ncu_full_<what>.txt                                        key metrics of the ncu --set full captures
  profiles/traffic.json                                                     dram bytes per launch of kernel 1 / kernel 2 / fused (+ git sha)"""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()

for name in ("bench_line.json", "bench_reference.json", "configs.json", "launches.csv", "gpu.txt"):
    src = os.path.join(G, f"{tag}_{name}")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{tag}_{name}"))

# ---- launch list
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp, errors="ignore") if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = {}
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += ms; a[2].append(ms)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launch_list_summary.md"), "w") as f:
        f.write(f"# {tag} -- ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs` (commit {sha})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (first 400 launches: synth, the device-resident steps of 256 pages\n"
                "on the two-kernel path and on the default fused path, then the 8-page chunks of the e2e pass).  Times under ncu are cold-cache and\n"
                f"serialised: compare SHARES, not absolutes.  Raw CSV: `{tag}_launches.csv`.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:110]}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % |\n")
        big = {k: sorted(a[2])[-4:] for k, a in agg.items() if k.startswith(("integral_sq_kernel", "threshold_tma_kernel"))}
        if len(big) == 2:
            avg = {k: sum(v) / len(v) for k, v in big.items()}
            s = sum(avg.values())
            f.write("\nDevice-resident two-kernel steps only (the 4 longest launches of each kernel = the 256-page launches):\n\n"
                    "| kernel | avg ms under ncu | share of K1+K2 |\n|---|---:|---:|\n")
            for k, v in avg.items():
                f.write(f"| `{k[:80]}` | {v:.3f} | {100 * v / s:.1f} % |\n")
            try:
                line = json.loads(open(os.path.join(G, f"{tag}_bench_line.json")).read().strip().splitlines()[-1])
                kk = line["roofline"]["kernels"]
                a, b = kk["integral"]["avg_ms"], kk["threshold"]["avg_ms"]
                f.write(f"\nbench.py's own CUDA-event figures for the same kernels ({tag}_bench_line.json): integral {a:.2f} ms, threshold {b:.2f} ms "
                        f"per launch -> shares {100 * a / (a + b):.1f} % / {100 * b / (a + b):.1f} %.\n")
            except Exception as ex:
                f.write(f"\n(bench line not available: {ex})\n")

# ---- ncu --set full summaries + traffic.json
traffic = {}
for what in ("k1k2", "fused", "otsu"):
    rep = os.path.join(G, f"{tag}_{what}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{tag}_ncu_full_{what}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, commit {sha}; scripts/gpu_profile_round.sh {tag}; read with scripts/ncu_summary.py\n" + out)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units = rr[0], rr[1]
    def col(name):
        return hdr.index(name)
    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    for r in rr[2:]:
        name = r[col("Kernel Name")]
        fam = "integral" if "integral_sq" in name else "threshold" if "threshold_tma" in name else "fused" if "local_fused" in name else None
        if not fam or fam in traffic:
            continue
        rd = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
        wr = to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
        t = float(r[col("gpu__time_duration.sum")].replace(",", ""))
        tu = units[col("gpu__time_duration.sum")]
        traffic[fam] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                        "gpu_time_ms_under_ncu": t / 1e6 if tu.startswith("ns") else t / 1e3 if tu.startswith("us") else t,
                        "pages_per_launch": 256, "kernel": name.split("(")[0][-60:],
                        "source": f"profiles/{tag}_ncu_full_{what}.txt (ncu --set full --clock-control none, scripts/prof_step.py pages=256, commit {sha})"}
if traffic:
    traffic["commit"] = sha
    with open(os.path.join(P, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
print("profiles written for", tag, "at", sha, "->", sorted(os.listdir(P))[-12:])
