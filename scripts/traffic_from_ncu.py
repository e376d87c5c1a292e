"""DRAM bytes of the bench's own rotation launch: picks, from an `ncu --csv --metrics dram__bytes_read.sum,
dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size -k regex:br7_kernel` log of one bench step, the launch with
the largest grid (the 8192-job launch of the step) and writes profiles/br_kernel_traffic.json, which bench.py reports as
roofline.traffic when kernel name and job count match."""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    d = launches.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
def to_bytes(v):
    x, u = v
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
best = max(launches.values(), key=lambda d: d.get("launch__grid_size", (0, ""))[0])
grid = int(best["launch__grid_size"][0])
dram = to_bytes(best["dram__bytes_read.sum"]) + to_bytes(best["dram__bytes_write.sum"])
ms = best["gpu__time_duration.sum"][0] * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3}.get(best["gpu__time_duration.sum"][1], 1)
jobs = int(sys.argv[2]) if len(sys.argv) > 2 else grid * 8
out = {"kernel": "br7_kernel<8>", "jobs_in_captured_launch": jobs, "grid": grid, "dram_bytes_per_launch": dram,
       "dram_bytes_per_job": dram / jobs, "ms_under_ncu": ms, "launches_seen": len(launches),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on `bench.py --steps 1 --warmup 0 --no-netlist --no-cpu-baseline` (scripts/gpu_r2_final2.sh)",
       "round": "r02"}
json.dump(out, open(sys.argv[3] if len(sys.argv) > 3 else "profiles/br_kernel_traffic.json", "w"), indent=1)
print(out)
