#!/usr/bin/env python
"""Copy one scripts/gpu_final.sh evidence set from gpurun_out/ (scratch) into profiles/ (tracked) and derive the
summaries the docs cite: launch shares, ncu text summaries, DRAM traffic per launch, clock summary.

    python scripts/publish_profiles.py r01c r01      # gpurun tag -> profiles prefix
"""
import csv
import io
import json
import os
import shutil
import statistics
import subprocess
import sys

tag, pre = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)


def cp(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
        return True
    print("missing", src)
    return False


for m in ("hifigan", "basis-melgan", "multiband-hifigan", "melgan"):
    cp(f"bench_{m}_{tag}.json", f"{pre}_bench_{m}.json")
    cp(f"profile_{m}_{tag}.json", f"{pre}_layers_{m}.json")
cp(f"bench_hifigan_fp32path_{tag}.json", f"{pre}_bench_hifigan_fp32path.json")
cp(f"bench_reference_{tag}.json", f"{pre}_bench_reference.json")
cp(f"launches_{tag}.csv", f"{pre}_launches.csv")
for n in ("fused_unit", "tc2_c128k11", "tc2_basis"):
    cp(f"ncu_{n}_{tag}.ncu-rep", f"{pre}_ncu_{n}.ncu-rep")

# ---- launch shares
rows = [l for l in open(os.path.join(P, f"{pre}_launches.csv")) if l.startswith('"')]
r = list(csv.reader(rows))
hdr, data = r[0], r[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for d in data:
    name = d[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(d[vi])
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{pre}_launch_shares.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 1 --warmup 3 "
            "(cold-cache, serialised launches: compare SHARES)\n")
    f.write(f"launches {len(data)}\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-52s n=%4d  share %5.1f%%  avg %8.1f us\n" % (k[:52], a[0], 100 * a[1] / tot, a[1] / a[0] / 1e3))
print(open(os.path.join(P, f"{pre}_launch_shares.txt")).read())

# ---- ncu text summaries + traffic
traffic = {"_doc": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures "
                   f"(profiles/{pre}_ncu_*.ncu-rep), taken at --batch 8; bench.py scales by batch (traffic is "
                   "proportional to positions)."}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    return rr[0], rr[1], rr[2:]


for n in ("fused_unit", "tc2_c128k11", "tc2_basis"):
    rep = os.path.join(P, f"{pre}_ncu_{n}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run([sys.executable, "scripts/ncu_summary.py", rep], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{pre}_ncu_{n}.txt"), "w").write(txt)
    hdr, units, data = raw(rep)
    d = data[0]

    def val(k):
        i = hdr.index(k)
        v = float(d[i])
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    print(n, "dram bytes/launch %.1f MB" % (b / 1e6), "duration", d[hdr.index("gpu__time_duration.sum")], "us",
          "tensor active %", d[hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")],
          "issue active %", d[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")])
    if n == "fused_unit":
        traffic.setdefault("hifigan", {})["tcgen05-fused-unit"] = {
            "bytes_per_launch": b, "batch": 8, "launch": "fused ResBlock1 unit C=32 k=3 d=1 (L=120000), the first stage-3 unit (tc3 launch index 9 of a forward)",
            "algorithmic_bytes": 2 * 8 * 32 * 120000 * 4}
    elif n == "tc2_c128k11":
        traffic.setdefault("hifigan", {})["tcgen05"] = {
            "bytes_per_launch": b, "batch": 8,
            "launch": "conv C=128 k=11 d=3 (L=8000); output still L2-resident at kernel end",
            "algorithmic_bytes": 2 * 8 * 128 * 8000 * 4}
    else:
        traffic.setdefault("basis-melgan", {})["tcgen05"] = {
            "bytes_per_launch": b, "batch": 8,
            "launch": "fused pair 1x1 (Cin=512 -> 256, L=16000, 9 utterances incl. the zero pass); inputs partly L2-resident",
            "algorithmic_bytes": 9 * 16000 * (512 + 256) * 4}
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)

# ---- clocks
cpath = os.path.join(G, f"clocks_{tag}.csv")
if os.path.exists(cpath):
    rr = list(csv.reader(open(cpath)))
    h = [x.strip() for x in rr[0]]
    rows = [x for x in rr[1:] if len(x) == len(h)]
    sm = [float(x[h.index("clocks.current.sm [MHz]")].split()[0]) for x in rows]
    pw = [float(x[h.index("power.draw [W]")].split()[0]) for x in rows]

    def cnt(col):
        i = [k for k, n in enumerate(h) if col in n][0]
        return sum(1 for x in rows if "Active" in x[i] and "Not" not in x[i])
    busy = [s for s, p in zip(sm, pw) if p > 400]
    with open(os.path.join(P, f"{pre}_clocks_summary.txt"), "w") as f:
        f.write(f"nvidia-smi -lms 200 during scripts/gpu_final.sh (tests + benches), {len(rows)} samples\n")
        f.write("clocks.sm MHz: median %.0f  p10 %.0f  min %.0f  max %.0f (clocks.max.sm 1965); under load (>400 W, %d samples): "
                "median %.0f min %.0f\n" % (statistics.median(sm), sorted(sm)[len(sm) // 10], min(sm), max(sm), len(busy),
                                           statistics.median(busy) if busy else 0, min(busy) if busy else 0))
        f.write("power.draw W: median %.0f max %.0f\n" % (statistics.median(pw), max(pw)))
        f.write("samples with hw_slowdown Active: %d, hw_thermal_slowdown: %d, sw_thermal_slowdown: %d, sw_power_cap: %d\n" % (
            cnt("hw_slowdown"), cnt("hw_thermal_slowdown"), cnt("sw_thermal_slowdown"), cnt("sw_power_cap")))
    print(open(os.path.join(P, f"{pre}_clocks_summary.txt")).read())
