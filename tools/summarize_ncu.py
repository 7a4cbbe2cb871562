#!/usr/bin/env python
"""ncu report -> compact per-launch summary table (CSV on stdout):  python tools/summarize_ncu.py X.ncu-rep > profiles/X_summary.csv
   launch list (ncu --metrics gpu__time_duration.sum --csv log) -> per-kernel totals:  python tools/summarize_ncu.py --launches X.csv"""
import csv, io, subprocess, sys, collections

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_tmem_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tma.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
           "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
           "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
           "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_sample_count"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        tot[name] += float(r[iv].replace(",", "")); cnt[name] += 1
    unit = rows[1][hdr.index("Metric Unit")]
    scale = 1e-3 if unit in ("ns", "nsecond") else 1.0
    all_t = sum(tot.values()) * scale
    print("kernel,launches,total_us,avg_us,share")
    for k, v in tot.most_common():
        print(f"{k},{cnt[k]},{v * scale:.1f},{v * scale / cnt[k]:.2f},{v * scale / all_t:.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    data = rows[2:]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [f"{r[ik].split('(')[0].replace('b2rl::', '')}#{i}" for i, r in enumerate(data)])
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            w.writerow([m, units[i]] + [r[i] for r in data])
    for i, m in enumerate(hdr):  # every tensor-pipe counter the report holds, whatever this ncu version calls them
        if "pipe_tensor" in m and m not in METRICS:
            w.writerow([m, units[i]] + [r[i] for r in data])


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[1])
