"""Derive netlist fixtures + expected plaintext results from the reference's test assets.

    python tests/golden/make_netlists.py        (in the build container, /root/reference present)

For each case: blueprint (test/config-toml/*.toml) -> flat netlist (tests/tools/netlist_tools.py), request
packet (test/in/*.in) and the reference's golden result packet (test/out/*.out), the pairs registered in
test.rb:387-548.  Stored as tests/golden/netlists/<case>.npz (+ cases.json)."""
import json
import sys
import tomllib
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "tools"))
import netlist_tools as T  # noqa: E402

REF = Path("/root/reference/test")
CASES = [  # (name, blueprint, in, out, cycles)
    ("counter-4bit", "counter-4bit.toml", "test13.in", "test13.out", 3),
    ("addr-4bit", "addr-4bit.toml", "test04.in", "test04.out", 1),
    ("div-8bit", "div-8bit.toml", "test05.in", "test05.out", 1),
    ("register-init-4bit", "register-init-4bit.toml", "test17.in", "test17.out", 1),
    ("mux-ram-8-16-16", "mux-ram-8-16-16.toml", "test08.in", "test08.out", 8),
    ("cahp-pearl-mux", "cahp-pearl-mux.toml", "test09.in", "test09-pearl.out", 3),
    ("cahp-ruby-mux", "cahp-ruby-mux.toml", "test09.in", "test09-ruby.out", 7),
]


def packet(path):
    t = tomllib.load(open(path, "rb"))
    return {"cycles": t.get("cycles"), "bits": {e["name"]: {"size": e["size"], "bytes": e["bytes"]} for e in t.get("bits", [])},
            "ram": {e["name"]: {"size": e["size"], "bytes": e["bytes"]} for e in t.get("ram", [])},
            "rom": {e["name"]: {"size": e["size"], "bytes": e["bytes"]} for e in t.get("rom", [])}}


def main():
    out = ROOT / "tests" / "golden" / "netlists"
    out.mkdir(parents=True, exist_ok=True)
    meta = {}
    for name, bp, fin, fout, cycles in CASES:
        try:
            nl = T.read_blueprint(REF / "config-toml" / bp)
        except Exception as e:  # noqa: BLE001
            print("skip", name, e)
            continue
        nl.save(out / f"{name}.npz")
        meta[name] = {"cycles": cycles, "request": packet(REF / "in" / fin), "expected": packet(REF / "out" / fout),
                      "source": {"blueprint": f"test/config-toml/{bp}", "in": f"test/in/{fin}", "out": f"test/out/{fout}"}}
        print(name, nl.n, "nodes", (out / f"{name}.npz").stat().st_size, "bytes")
    json.dump(meta, open(out / "cases.json", "w"), indent=1)


if __name__ == "__main__":
    main()
