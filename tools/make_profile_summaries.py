#!/usr/bin/env python
"""Turns the raw outputs of tools/final_profiles.sh (gpurun_out/r2m_*) into the tracked summaries under profiles/:
the ncu --set full capture of the persistent U-family epoch kernel, the launch list with per-kernel shares, the bench lines."""
import collections
import csv
import json
import os
import re
import shutil
import statistics as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

rows = list(csv.reader(open(os.path.join(G, "r2m_umma_raw.csv"))))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
dur = float(m["gpu__time_duration.sum"])
rd, wr = float(m["dram__bytes_read.sum"]), float(m["dram__bytes_write.sum"])
u = dict(zip(hdr, units))
with open(os.path.join(P, "r2_train_umma_persistent_ncu.txt"), "w") as f:
    f.write("# r2 (final) — ncu --set full capture of ONE launch of umma::train_umma_kernel<18,18,1> (persistent epoch kernel: 32 minibatches of 8192 samples, C3)\n")
    f.write("# command: ncu --set full --clock-control none --import-source on -k regex:train_umma -s 20 -c 1 -o gpurun_out/r2m_umma python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-microbench   (tools/final_profiles.sh, tools/make_profile_summaries.py)\n")
    f.write(f"# per minibatch: {dur:.1f} {u['gpu__time_duration.sum']} / 32 = {dur / 32:.1f} under the profiler; DRAM {rd:.2f} {u['dram__bytes_read.sum']} read + {wr:.2f} {u['dram__bytes_write.sum']} written per launch (algorithmic 8192 x 160 B = 1.31 MB per minibatch)\n")
    f.write(f"# tensor pipe active {float(m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']):.2f} % (104 tcgen05.mma per tile; a latency chain: tile ~6.6 us + gradient step ~12 us per minibatch), warps active {float(m['sm__warps_active.avg.pct_of_peak_sustained_active']):.1f} %\n\n")
    for h, un, v in zip(hdr, units, vals):
        f.write(f"{h:100s} {v} {un}\n")

shutil.copy(os.path.join(G, "r2m_c3_launches.csv"), os.path.join(P, "r2_c3_launches.csv"))
rows = [r for r in csv.reader(open(os.path.join(G, "r2m_c3_launches.csv"))) if len(r) > 5]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    a = agg.setdefault(re.sub(r"\(.*", "", r[ki]), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, "r2_c3_launches_summary.txt"), "w") as f:
    f.write("# r2 (final) — ncu launch list summary, C3 (4096 envs, [64,64])\n")
    f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-microbench   (tools/final_profiles.sh)\n")
    f.write("# (cold-cache, serialised times: compare SHARES, not absolutes; the kernels of the CUDA graphs are listed as they replay;\n#  train_umma_kernel<..,1> = one persistent launch per epoch of 32 minibatches; the trailing train_umma_kernel<..,0> launches are\n#  bench.py's stand-alone timing of one minibatch for the roofline object)\n\n")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{name[:60]:62s} launches {n:5d} total {t:10.1f} us avg {t / n:9.2f} us share {100 * t / tot:5.1f}%\n")

for w in ("c3", "c1", "c1x4096", "c4"):
    shutil.copy(os.path.join(G, f"r2m_{w}.json"), os.path.join(P, f"r2_bench_{w}.json"))
shutil.copy(os.path.join(G, "r2m_ref.json"), os.path.join(P, "r2_bench_c3_reference_arm.json"))

# per-CTA timeline summary (appended to the phase-clock file by hand-kept history below it)
lines = open(os.path.join(G, "r2m_uprof.txt")).read().splitlines()
ctas = []
for l in lines:
    mm = re.match(r"cta\s+(\d+):\s+(.*)", l)
    if mm:
        ctas.append((int(mm.group(1)), [int(x) for x in mm.group(2).split()]))
names = ["start", "weights staged", "tiles done", "flushed", "barrier 1", "reduce + Adam", "barrier 3"]
out = [l for l in lines if l.startswith("update") or l.startswith("umma phases") or "reduce phases" in l or l.startswith("train_fwdbwd")]
out.append("")
out.append("# per-CTA timeline of minibatch 2 (ns since the earliest start; min / median / max over the 64 CTAs of a tower, then median phase durations)")
for name, sel in (("tower 0 (pi)", [r for r in ctas if r[0] < 64]), ("tower 1 (V)", [r for r in ctas if r[0] >= 64])):
    out.append(name)
    cols = list(zip(*[r[1] for r in sel]))
    med = []
    for n, c in zip(names, cols):
        out.append(f"  {n:16s} min {min(c):6d}  med {int(st.median(c)):6d}  max {max(c):6d}")
        med.append(int(st.median(c)))
    d = [med[i + 1] - med[i] for i in range(6)]
    out.append(f"  median durations (ns): weights {d[0]} | tiles {d[1]} | flush {d[2]} | barrier 1 {d[3]} | reduce + Adam {d[4]} | barrier 3 {d[5]}")
print("\n".join(out))
