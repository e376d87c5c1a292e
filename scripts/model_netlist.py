"""Predict the clock-cycle time of a netlist on N GPUs from the launch-plan cost model (no GPU needed).

    python scripts/model_netlist.py tests/golden/netlists/cahp-pearl-mux.npz [--gpus 1 2 4 8]

Per dependency level: blind-rotation time = b200fhe_plan_ms(jobs) (the measured per-wave times of the kernel shapes,
iyokan_b200/csrc/b200fhe.cu plan_rotation), split across ranks when the sharding policy of iyokan_b200/shard.py would
split it (+ one collective), plus the key switch (0.07 ms for <= 148 gates on the split kernel, 0.71 us per gate above).
The model reproduces the measured clocks of profiles/r01_summary.md within a few percent; what it leaves out is
host-side launch latency and the unary / tick kernels.
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from iyokan_b200.lib import plan_ms  # noqa: E402
from iyokan_b200.netlist import NetEngine, Netlist  # noqa: E402

KS_SPLIT_MS, KS_PER_GATE_MS, COLLECTIVE_MS = 0.07, 5.8 / 8192, 0.1


def model(nl: Netlist, world: int = 1):
    eng = NetEngine(nl)
    jobs, widths = eng.level_bootstraps, eng.level_widths
    total, narrow, ncoll = 0.0, 0, 0
    for j, w in zip(jobs, widths):
        if j == 0:
            continue
        share_j, share_w = -(-j // world), -(-w // world)
        split = world > 1 and plan_ms(share_j) + COLLECTIVE_MS < plan_ms(j)
        jj, ww = (share_j, share_w) if split else (j, w)
        total += plan_ms(jj) + (KS_SPLIT_MS if ww <= 148 else ww * KS_PER_GATE_MS) + (COLLECTIVE_MS if split else 0.0)
        ncoll += split
        narrow += jj <= 148
    eng.close()
    return {"ms_per_cycle": total, "levels": len(jobs), "latency_bound_levels": narrow, "collectives": ncoll,
            "bootstraps_per_cycle": sum(jobs), "bootstraps_per_s": sum(jobs) / (total / 1e3)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("netlist")
    ap.add_argument("--gpus", type=int, nargs="*", default=[1, 2, 4, 8])
    a = ap.parse_args()
    nl = Netlist.load(a.netlist)
    for n in a.gpus:
        m = model(nl, n)
        print(f"{n} GPU(s): {m['ms_per_cycle'] / 1e3:.3f} s/cycle, {m['bootstraps_per_s'] / 1e3:.1f} k bootstraps/s, "
              f"{m['latency_bound_levels']}/{m['levels']} levels latency bound, {m['collectives']} collectives")
