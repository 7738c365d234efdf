"""Turn the raw ncu exports under gpurun_out/ into the small, committed summaries in profiles/.

    python profiles/summarize.py <round-tag>      e.g. r1b

Reads (all produced by the commands recorded in profiles/README.md):
  gpurun_out/prof_<tag>_step_c2.ncu-rep   ncu --set full of the step kernels, configs[2]
  gpurun_out/launches_<tag>_c2.csv        ncu launch list (gpu__time_duration.sum), configs[2]
  gpurun_out/traffic_<tag>_c3.csv         dram bytes + duration of the step kernels, configs[3]
Writes profiles/<tag>_*.{csv,json,md} and refreshes profiles/step_kernel_traffic.json (the
per-launch DRAM traffic that bench.py reports in roofline.traffic).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__cycles_elapsed.avg",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * mult


def full_report(tag):
    rep = os.path.join(G, f"prof_{tag}_step_c2.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    cols = [c for c in KEEP if c in hdr] + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled")
                                             and h.endswith("per_issue_active.ratio")]
    out = []
    for r in data:
        d = {"kernel": r[name_i]}
        for c in cols:
            i = hdr.index(c)
            d[c] = {"value": r[i], "unit": units[i]}
        out.append(d)
    with open(os.path.join(OUT, f"{tag}_step_c2_ncu_full_summary.json"), "w") as fh:
        json.dump({"source": f"gpurun_out/prof_{tag}_step_c2.ncu-rep (ncu --set full --clock-control none)",
                   "workload": "configs[2] 4096x2048 NACA 4412 alpha=10", "launches": out}, fh, indent=1)
    return out


def launch_list(tag):
    path = os.path.join(G, f"launches_{tag}_c2.csv")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r["Kernel Name"]
        per[k][0] += 1
        per[k][1] += float(r["Metric Value"]) / 1e3       # ns -> us
    total = sum(v[1] for v in per.values())
    with open(os.path.join(OUT, f"{tag}_launches_c2.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_gpu_time"])
        for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{t:.1f}", f"{t / n:.2f}", f"{t / total:.4f}"])
    return per, total


def traffic(tag):
    path = os.path.join(G, f"traffic_{tag}_c3.csv")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    by_id = defaultdict(dict)
    for r in rows:
        by_id[(r["ID"], r["Kernel Name"])][r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    out = []
    for (kid, name), m in by_id.items():
        d = {"id": int(kid), "kernel": name}
        for k, (v, u) in m.items():
            d[k] = to_bytes(v, u) if "bytes" in k else float(v)
            if "bytes" not in k:
                d[k + ".unit"] = u
        out.append(d)
    with open(os.path.join(OUT, f"{tag}_traffic_c3.json"), "w") as fh:
        json.dump({"source": f"gpurun_out/traffic_{tag}_c3.csv", "workload": "configs[3] 32768x16384 NACA 2412 alpha=5",
                   "launches": out}, fh, indent=1)
    return out


def main():
    tag = sys.argv[1]
    full = full_report(tag)
    ll = launch_list(tag)
    tr = traffic(tag)
    traffic_json = {}
    tpath = os.path.join(OUT, "step_kernel_traffic.json")
    if os.path.exists(tpath):
        traffic_json = json.load(open(tpath))
    if full:
        fast = [d for d in full if re.search(r"step_kernel<0, 0(, 0)?>", d["kernel"])]
        gen = [d for d in full if re.search(r"step_kernel<0, 1(, 0)?>", d["kernel"])]

        def dram(d):
            return (to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"])
                    + to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"]))
        if fast:
            t = sum(dram(d) for d in fast) / len(fast) + (sum(dram(d) for d in gen) / len(gen) if gen else 0)
            traffic_json["configs[2]"] = t
    if tr:
        fast = [d for d in tr if re.search(r"step_kernel<0, 0(, 0)?>", d["kernel"])]
        gen = [d for d in tr if re.search(r"step_kernel<0, 1(, 0)?>", d["kernel"])]
        s = lambda L: sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in L) / max(1, len(L))
        traffic_json["configs[3]"] = s(fast) + s(gen)
    traffic_json["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per step (fast + general kernel), "
                             f"from the {tag} ncu captures; algorithmic bytes are 72 B x cells")
    json.dump(traffic_json, open(tpath, "w"), indent=1)
    print(json.dumps(traffic_json, indent=1))
    if ll:
        per, total = ll
        for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:6]:
            print(f"{t / total:6.1%} {n:4d} x {t / n:9.2f} us  {k[:90]}")


if __name__ == "__main__":
    main()
