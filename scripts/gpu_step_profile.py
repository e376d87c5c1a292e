"""Where a clock of a netlist goes, step by step: `iyokan-b200 --dump-time-csv-prefix` (b200net_profile_run: the clock
replayed step by step between CUDA events) on the reference's blueprints; prints one line per schedule step."""
import collections, csv, datetime, json, os, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench_netlist as BN

names = sys.argv[1:] or ["cahp-pearl-mux"]
work = Path(tempfile.mkdtemp(prefix="b200fhe_prof_"))
tools = BN.RefTools(work)
assets = BN._assets(work / "test")
tools.genkeys()
out = {}
for name in names:
    bp_rel, req_rel = BN.CASES[name]
    req, enc = tools.encrypt_request(assets / req_rel, name)
    r = subprocess.run([str(tools.O.IYOKAN_B200), "tfhe", "--blueprint", str(assets / bp_rel), "--evalkey", str(tools.ek), "-i", str(enc),
                        "-o", str(work / "res.enc"), "-c", "3", "--dump-time-csv-prefix", str(work / name)],
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr[-800:]
    rows = list(csv.reader(open(work / f"{name}-2.csv")))   # third cycle: warm
    ts = lambda s: datetime.datetime.strptime(s, "%Y-%m-%d %H:%M:%S.%f")
    steps = collections.OrderedDict()
    for start, end, index, nid, kind, desc in rows:
        if not desc.startswith("step "):
            continue
        k = int(desc.split()[1])
        d = steps.setdefault(k, {"gates": 0, "bootstraps": 0, "ms": (ts(end) - ts(start)).total_seconds() * 1e3})
        d["gates"] += 1
        d["bootstraps"] += 0 if kind in ("NOT", "CONSTONE", "CONSTZERO") else (2 if kind == "MUX" else 1)
    tot = sum(d["ms"] for d in steps.values())
    print(f"{name}: {len(steps)} steps, {sum(d['bootstraps'] for d in steps.values())} bootstraps, {tot:.1f} ms per clock when replayed step by step "
          f"(file time stamps have 1 ms resolution)")
    hist = collections.Counter()
    for k, d in steps.items():
        b = d["bootstraps"]
        bucket = "<=74" if b <= 74 else "<=148" if b <= 148 else "<=592" if b <= 592 else "<=1184" if b <= 1184 else "<=2368" if b <= 2368 else ">2368"
        hist[bucket] += 1
        hist[bucket + " ms"] += d["ms"]
        hist[bucket + " bootstraps"] += b
    for bucket in ("<=74", "<=148", "<=592", "<=1184", "<=2368", ">2368"):
        if hist[bucket]:
            print(f"  steps with {bucket:7s} bootstraps: {hist[bucket]:3d} steps, {hist[bucket + ' bootstraps']:6d} bootstraps, {hist[bucket + ' ms']:7.1f} ms")
    out[name] = {"steps": [{"step": k, **d} for k, d in steps.items()], "ms_per_clock_step_by_step": tot}
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "step_profile.json", "w"), indent=1)
