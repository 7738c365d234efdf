"""Summaries of the configs[3] captures of the DEFAULT bench command (double steps).

    python profiles/summarize_c3.py <tag>        e.g. r1f

Reads   gpurun_out/prof_<tag>_step_c3.ncu-rep   ncu --set full of the step kernels
        gpurun_out/launches_<tag>_c3.csv         ncu launch list (gpu__time_duration.sum)
Writes  profiles/<tag>_step_c3_ncu_full_summary.json, profiles/<tag>_launches_c3.csv (+ _raw.csv),
        and the "configs[3]:step2" entry of profiles/step_kernel_traffic.json (DRAM bytes of one
        step2_kernel launch = two steps of the deep lattice; bench.py reports it as roofline.traffic).
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize import G, KEEP, OUT, to_bytes  # noqa: E402


def main():
    tag = sys.argv[1]
    rep = os.path.join(G, f"prof_{tag}_step_c3.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    cols = [c for c in KEEP if c in hdr] + ["launch__shared_mem_per_block_dynamic"] + \
           [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    cols = [c for c in cols if c in hdr]
    out = []
    for r in data:
        d = {"kernel": r[name_i]}
        for c in cols:
            i = hdr.index(c)
            d[c] = {"value": r[i], "unit": units[i]}
        out.append(d)
    with open(os.path.join(OUT, f"{tag}_step_c3_ncu_full_summary.json"), "w") as fh:
        json.dump({"source": f"gpurun_out/prof_{tag}_step_c3.ncu-rep (ncu --set full --clock-control none --import-source on "
                             "-k regex:step2_kernel|step_kernel|copy_tasks -s 12 -c 12 (r1f) / -c 6 (r1g); python bench.py --steps 20 --warmup 3 "
                             "--no-cpu-baseline --no-e2e)",
                   "workload": "configs[3] 32768x16384 NACA 2412 alpha=5, double steps", "launches": out}, fh, indent=1)

    def dram(d):
        return (to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"])
                + to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"]))
    s2 = [d for d in out if "step2_kernel" in d["kernel"]]
    others = [d for d in out if "step2_kernel" not in d["kernel"] and "copy_tasks" not in d["kernel"]]
    tpath = os.path.join(OUT, "step_kernel_traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if s2:
        tj["configs[3]:step2"] = sum(dram(d) for d in s2) / len(s2)
        # the list-driven passes that accompany one step2 launch (captured launches / double steps captured)
        tj["configs[3]:step2_passes"] = sum(dram(d) for d in others) / len(s2)
        tj["_note_step2"] = (f"{tag}: DRAM bytes of ONE step2_kernel launch (two steps; algorithmic 2 x 72 B x cells = "
                             "77.3e9) and of the list-driven passes that accompany it")
    json.dump(tj, open(tpath, "w"), indent=1)
    print(json.dumps(tj, indent=1))

    path = os.path.join(G, f"launches_{tag}_c3.csv")
    if os.path.exists(path):
        shutil.copy(path, os.path.join(OUT, f"{tag}_launches_c3_raw.csv"))
        lines = [l for l in open(path) if l.startswith('"')]
        lrows = list(csv.DictReader(io.StringIO("".join(lines))))
        per = defaultdict(lambda: [0, 0.0])
        for r in lrows:
            per[r["Kernel Name"]][0] += 1
            per[r["Kernel Name"]][1] += float(r["Metric Value"].replace(",", "")) / 1e3
        total = sum(v[1] for v in per.values())
        with open(os.path.join(OUT, f"{tag}_launches_c3.csv"), "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_gpu_time"])
            for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, n, f"{t:.1f}", f"{t / n:.2f}", f"{t / total:.4f}"])
                print(f"{t / total:6.1%} {n:4d} x {t / n:10.2f} us  {k[:100]}")


if __name__ == "__main__":
    main()
