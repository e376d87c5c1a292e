"""Copy the evidence of the latest GPU calls from gpurun_out/ (scratch) into profiles/ (tracked)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r01"


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def main():
    PROF.mkdir(exist_ok=True)
    # 1. full captures -> markdown summaries + traffic json
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    for rep in sorted(OUT.glob("prof_br_*.ncu-rep")):
        tag = rep.stem.replace("prof_br_", "")
        txt = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(rep)],
                             capture_output=True, text=True).stdout
        (PROF / f"{ROUND}_br_kernel_{tag}.md").write_text(
            f"# ncu --set full --clock-control none, blind-rotation kernel ({tag}), {ROUND}\n\n"
            f"Source: `{rep.name}` (scripts/gpu_round.sh / gpu_ncu.sh); summarised by scripts/ncu_summary.py\n\n```\n{txt}```\n")
        if tag == "auto":  # the dominant launch of the default bench.py step
            m, u = raw_metrics(rep)
            rd = float(m["dram__bytes_read.sum"]) * unit[u["dram__bytes_read.sum"]]
            wr = float(m["dram__bytes_write.sum"]) * unit[u["dram__bytes_write.sum"]]
            grid, block = int(float(m["launch__grid_size"])), int(float(m["launch__block_size"]))
            name, per_cta = {384: ("br3_kernel<6>", 6), 256: ("br3_kernel<4>", 4)}.get(block, (f"block{block}", 1))
            if block == 384 and float(m["launch__registers_per_thread"]) < 100:
                name, per_cta = "br4_kernel", 1
            jobs = int(os.environ.get("DOMINANT_JOBS", grid * per_cta))
            json.dump({"kernel": name, "jobs_in_captured_launch": jobs, "dram_bytes_per_launch": rd + wr,
                       "dram_bytes_per_job": (rd + wr) / jobs, "ms_under_ncu": float(m["gpu__time_duration.sum"]),
                       "source": rep.name, "round": ROUND},
                      open(PROF / "br_kernel_traffic.json", "w"), indent=1)
    # 2. launch list
    lc = OUT / "launches.csv"
    if lc.exists():
        rows = [r for r in csv.reader(l for l in open(lc) if not l.startswith("==")) if r]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        d = defaultdict(list)
        for r in rows[1:]:
            if len(r) > vi:
                d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        tot = sum(sum(v) for v in d.values())
        lines = [f"# ncu launch list (gpu__time_duration.sum, --clock-control none), {ROUND}", "",
                 "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv python bench.py --steps 1 --warmup 3` (scripts/gpu_round.sh)", "",
                 "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| `{k}` | {len(v)} | {sum(v) / 1e6:.3f} | {sum(v) / tot:.3f} |")
        (PROF / f"{ROUND}_launch_list.md").write_text("\n".join(lines) + "\n")
        (PROF / f"{ROUND}_launches.csv").write_text(open(lc).read())
    for name in ("bench.log", "bench_ref.log", "pytest_gpu.log", "smoke.log", "microbench_pipes.txt", "latency_table.json",
                 "net_mux-ram-8-16-16_n1.log", "net_cahp-pearl-mux_n1.log", "net_cahp-ruby-mux_n1.log",
                 "net_mux-ram-8-16-16_n2.log", "net_cahp-pearl-mux_n2.log", "net_cahp-ruby-mux_n2.log",
                 "net_mux-ram-8-16-16_n4.log", "net_cahp-pearl-mux_n4.log", "net_cahp-ruby-mux_n4.log",
                 "net_mux-ram-8-16-16_n8.log", "net_cahp-pearl-mux_n8.log", "net_cahp-ruby-mux_n8.log",
                 "bench_n1.log", "bench_n2.log", "bench_n4.log", "bench_n8.log"):
        if (OUT / name).exists():
            txt = "\n".join(l for l in open(OUT / name).read().splitlines() if "Warning" not in l and "warn" not in l)
            (PROF / f"{ROUND}_{name}").write_text(txt + "\n")


if __name__ == "__main__":
    main()
