"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py cite.
usage: python profiles/summarize_ncu.py gpurun_out/prof_X.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print(f"## {rep}: {r[hdr.index('Kernel Name')]}")
            vals = {}
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    vals[w] = (r[i], units[i])
                    print(f"{w:82s} {r[i]:>18s} {units[i]}")
            try:
                def f(k):
                    v, u = vals[k]
                    x = float(v.replace(",", ""))
                    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
                tr = f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
                t = f("gpu__time_duration.sum")
                print(f"{'traffic (dram read+write) bytes':82s} {tr:18.0f}")
                print(f"{'traffic / duration GB/s':82s} {tr / t / 1e9:18.1f}")
            except Exception as e:
                print("derived values unavailable:", e)
            print()


if __name__ == "__main__":
    main()
