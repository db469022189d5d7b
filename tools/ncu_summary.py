"""Summarise an ncu report (+ optional launch list) into profiles/<name>.md.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [gpurun_out/launches.csv] profiles/r01_assembly
"""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    args = sys.argv[1:]
    rep, dst = args[0], args[-1]
    launches = args[1] if len(args) == 3 else None
    hdr, units, rows = raw_rows(rep)
    L = [f"# ncu summary: `{rep}`", ""]
    for r in rows:
        L.append(f"## {r[hdr.index('Kernel Name')]} (launch id {r[hdr.index('ID')]})")
        L.append("")
        L.append("| metric | value | unit |")
        L.append("|---|---|---|")
        for m in METRICS:
            if m in hdr:
                L.append(f"| `{m}` | {r[hdr.index(m)]} | {units[hdr.index(m)]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
        L.append("")
        L.append("Top stall reasons (warps stalled per issue-active cycle):")
        L.append("")
        for v, h in sorted(stalls, reverse=True)[:6]:
            name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            L.append(f"* {name}: {v:.2f}")
        L.append("")
    if launches:
        rows = [r for r in csv.reader(open(launches)) if r and r[0].isdigit()]
        agg = OrderedDict()
        for r in rows:
            name, unit, val = r[4], r[-2], float(r[-1].replace(",", ""))
            val *= {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0}.get(unit, 1.0)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += val
        tot = sum(a[1] for a in agg.values())
        L.append(f"## Launch list `{launches}` (cold-cache, serialised: compare shares)")
        L.append("")
        L.append("| kernel | launches | total | share |")
        L.append("|---|---|---|---|")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            L.append(f"| `{name[:90]}` | {n} | {t / 1e6:.3f} ms | {t / tot * 100:.1f}% |")
        L.append("")
    open(dst + ".md", "w").write("\n".join(L))
    print("\n".join(L))


if __name__ == "__main__":
    main()
