"""Summarise gpurun_out ncu artefacts into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_prof.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "sm__cycles_elapsed.avg"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum launch list ({src}); cold-cache, serialised: compare shares\n")
        f.write(f"# total {T:.3f} ms over {sum(cnt.values())} launches\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"{k[:80]:80s} n={cnt[k]:5d} total_ms={v:10.3f} share={v / T:6.3f} avg_ms={v / cnt[k]:.4f}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({src}), one column per captured launch\n")
        ki = hdr.index("Kernel Name")
        f.write("kernel: " + " | ".join(r[ki][:70] for r in data) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k} [{units[i]}]: " + " | ".join(r[i] for r in data) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
