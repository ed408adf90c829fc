"""Turn the ncu artefacts of `profiles/run_ncu.sh <tag>` into the tracked summaries under profiles/.

    python profiles/summarize.py <tag>

reads  gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list of one bench step window)
       gpurun_out/prof_<tag>.ncu-rep   (ncu --set full capture of the tensor-core kernels)
writes profiles/<tag>_launches.md      per-kernel launch count / total time / share of the captured window
       profiles/<tag>_launches.csv     the launch list itself (kernel, grid, block, ns)
       profiles/<tag>_kernels.md       per captured kernel: duration, pipe utilisations, DRAM traffic, occupancy, stalls
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
OUT = os.path.join(ROOT, "profiles")


def short(name):
    m = re.search(r"(\w+)(<[^>]*>)?\(", name)
    base = m.group(1) + (m.group(2) or "") if m else name
    return base.replace("ca::<unnamed>::", "")


def launches():
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    total = 0.0
    with open(os.path.join(OUT, f"{tag}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "ns"])
        for r in rows:
            k, ns = short(r[4]), float(r[14].replace(",", ""))
            w.writerow([r[0], k, r[8], r[7], int(ns)])
            c, t = agg.get(k, (0, 0.0))
            agg[k] = (c + 1, t + ns)
            total += ns
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list `{tag}` — {len(rows)} launches, {total / 1e6:.3f} ms summed "
                "(`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: use the SHARES)\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t / 1e6:.3f} | {100 * t / total:.1f} % | {t / c / 1e3:.1f} |\n")


KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("smsp__warps_active.avg.per_cycle_active", "active warps / SMSP"),
]


def kernels():
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(h)}
    with open(os.path.join(OUT, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# ncu --set full capture `{tag}` (per launch; `--clock-control none`, values under the profiler — "
                "timings here are NOT bench numbers)\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[idx['Kernel Name']])}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n")
            f.write("| metric | value |\n|---|---:|\n")
            for key, label in KEYS:
                if key in idx and r[idx[key]] != "":
                    f.write(f"| {label} | {r[idx[key]]} {u[idx[key]]} |\n")
            stalls = []
            for n, i in idx.items():
                m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", n)
                if m and not n.endswith("_not_issued") and r[i] not in ("", "0"):
                    stalls.append((float(r[i].replace(",", "")), m.group(1)))
            tot = sum(s for s, _ in stalls) or 1.0
            top = ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:6])
            f.write(f"| top warp-stall samples | {top} |\n\n")


def traffic():
    """Mean DRAM bytes (read + write) per launch of the dense GEMM kernel over the captured launches -> the
    `roofline.traffic` figure bench.py reports next to the algorithmic work."""
    import json
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(h)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows[2:]:
        k = short(r[idx["Kernel Name"]])
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]].replace(",", "")) * scale[u[idx[m]]]
        out.setdefault(k, []).append(tot)
    dense = [v for k, vs in out.items() if k.startswith("gemm_tcgen05_kernel<256") or k.startswith("gemm2_tcgen05_kernel") for v in vs]
    res = {"tag": tag, "source": f"ncu --set full capture prof_{tag}.ncu-rep (dram__bytes_read.sum + dram__bytes_write.sum)",
           "per_kernel_mean_bytes": {k: sum(v) / len(v) for k, v in out.items()},
           "dense_gemm_mean_bytes_per_launch": sum(dense) / len(dense) if dense else None,
           "dense_gemm_launches_captured": len(dense)}
    json.dump(res, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)


launches()
kernels()
traffic()
print("wrote", [p for p in os.listdir(OUT) if p.startswith(tag)])
