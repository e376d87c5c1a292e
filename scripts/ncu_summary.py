"""Summarise an .ncu-rep of the blind-rotation kernel: headline metrics + stall samples per kernel phase."""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
            "launch__shared_mem_per_block_dynamic"]
    print("== metrics ==")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h:80s} {v} {u}")
    for h, u, v in zip(hdr, units, vals):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.02:
            print(f"{h:80s} {v}")
    rows = page(rep, "source")
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[idx["Source"]]]
    keys = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_math", "stall_short_sb", "stall_not_selected",
            "stall_selected", "stall_dispatch", "stall_mio", "stall_lg", "stall_no_inst"]
    print("== stall samples per region (regions split at BAR.SYNC: prologue | forward | pointwise | inverse) ==")
    edges = [0] + bars + [len(data)]
    for a, b in zip(edges[:-1], edges[1:]):
        s = sum(int(r[idx["# Samples"]]) for r in data[a:b])
        ex = sum(int(r[idx["Instructions Executed"]]) for r in data[a:b])
        st = {k: sum(int(r[idx[k]]) for r in data[a:b]) for k in keys}
        st = {k.replace("stall_", ""): round(v / max(s, 1), 3) for k, v in st.items() if v > 0.03 * s}
        print(f"[{a:5d},{b:5d}) samples {s / tot:6.3f}  warp-instrs {ex / 1e9:7.3f} G  {st}")
    print("== top instructions by samples ==")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:12]:
        print(r[idx["# Samples"]], r[idx["Source"]].strip()[:60], "long_sb", r[idx["stall_long_sb"]], "wait", r[idx["stall_wait"]],
              "bar", r[idx["stall_barrier"]], "math", r[idx["stall_math"]], "short_sb", r[idx["stall_short_sb"]])


if __name__ == "__main__":
    main(sys.argv[1])
