"""Summarise one `ncu --set full` capture (round 2) into profiles/<tag>_ncu_full_summary.json.

    python profiles/summarize_r2.py gpurun_out/r2b_march_c3.ncu-rep r2b_march_c3 "what was captured"

Keeps the metrics the design is steered by (duration, DRAM bytes and throughput, issue slots,
occupancy, registers, instruction counts per pipe, L2 hit rate, shared-memory wavefronts), the
stall breakdown (average warps stalled per issue-active cycle) and, from the source page, the
instructions that collect the most stall samples.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_xu.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'sm__cycles_elapsed.avg.per_second']


def main():
    rep, tag, what = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    here = os.path.dirname(os.path.abspath(__file__))
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index('Kernel Name')]}
        for k in KEEP:
            if k in hdr:
                d[k] = {"value": r[hdr.index(k)], "unit": units[hdr.index(k)]}
        d["stalls_per_issue_active"] = {
            h.split('issue_stalled_')[1].split('_per_')[0]: float(r[i]) for i, h in enumerate(hdr)
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')}
        launches.append(d)
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    top = []
    if len(srows) > 2:
        sh = srows[1]
        ix = {h: i for i, h in enumerate(sh)}
        data = [r for r in srows[2:] if len(r) == len(sh)]
        total = sum(int(r[ix['# Samples']] or 0) for r in data)
        cols = [h for h in sh if h.startswith('stall_') and 'Not Issued' not in h]
        sums = {c: sum(int(r[ix[c]] or 0) for r in data) for c in cols}
        for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:25]:
            n = int(r[ix['# Samples']] or 0)
            why = max(cols, key=lambda c: int(r[ix[c]] or 0))
            top.append({"sass": r[ix['Source']].strip(), "samples": n, "share": round(n / max(total, 1), 4), "main_stall": why})
        sampling = {"total_samples": total, "by_reason": sums, "top_instructions": top}
    else:
        sampling = None
    out = {"source": rep, "what": what, "launches": launches, "warp_state_sampling": sampling}
    path = os.path.join(here, f"{tag}_ncu_full_summary.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(path)


if __name__ == "__main__":
    main()
