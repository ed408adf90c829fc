"""ncu --set full captures of the bandwidth / latency kernels (gpurun_out/prof_<tag>_bw*.ncu-rep) -> profiles/<tag>_bw_kernels.md
    python profiles/summarize_bw.py r01e"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
agg = collections.OrderedDict()
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{tag}_bw*.ncu-rep"))):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]

    def find(name, exact=True):
        for i, c in enumerate(h):
            if (c == name) if exact else c.endswith(name):
                return i
        return None

    idx = {"dur": find("gpu__time_duration.sum"), "rd": find("dram__bytes_read.sum"), "wr": find("dram__bytes_write.sum"),
           "grid": find("launch__grid_size"), "block": find("launch__block_size"),
           "l2": find("lts__t_sector_hit_rate.pct")}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[find("Kernel Name")]).replace("ca::<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")

        def num(k):
            i = idx[k]
            try:
                v = float(r[i].replace(",", ""))
            except (TypeError, ValueError):
                return 0.0
            return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "ms": 1e3}.get(units[i], 1)

        agg.setdefault(name, []).append({k: num(k) for k in idx})
out = [f"# ncu --set full captures `{tag}_bw*` of the bandwidth / latency kernels (per launch; under the profiler, "
       "`--clock-control none`; launches are serialised, so L2 is colder than in the step)", "",
       "| kernel | launches | grid x block | duration us | DRAM read MB | DRAM write MB | DRAM GB/s | of 6540 GB/s | L2 hit % |",
       "|---|---:|---|---:|---:|---:|---:|---:|---:|"]
for name, es in agg.items():
    n = len(es)
    m = {k: sum(e[k] for e in es) / n for k in es[0]}
    gbs = (m["rd"] + m["wr"]) / m["dur"] / 1e3
    out.append(f"| `{name}` | {n} | {int(m['grid'])} x {int(m['block'])} | {m['dur']:.1f} | {m['rd'] / 1e6:.1f} | "
               f"{m['wr'] / 1e6:.1f} | {gbs:.0f} | {gbs / 6540 * 100:.0f} % | {m['l2']:.0f} |")
open(os.path.join(ROOT, "profiles", f"{tag}_bw_kernels.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
