"""Copies the reference's TEST ASSETS (blueprints, synthesised netlists, request / golden packets in TOML form) that the
north-star workloads need into tests/golden/ref_assets/, so that bench.py and the GPU tests can run the reference's own
`iyokan` / `iyokan-packet` binaries (oracle/_ref/) and this repo's `iyokan-b200` on them where /root/reference does not
exist (the GPU box).  Data only - no reference source code.  Netlists are gzip-compressed.

    python tests/golden/make_ref_assets.py        # needs /root/reference (this container)
"""
import gzip
import shutil
from pathlib import Path

REF = Path("/root/reference/test")
OUT = Path(__file__).resolve().parent / "ref_assets"
ASSETS = [
    "config-toml/cahp-pearl-mux.toml", "config-toml/cahp-ruby-mux.toml", "config-toml/mux-ram-8-16-16.toml",
    "config-toml/addr-4bit.toml", "config-toml/counter-4bit.toml",
    "yosys-json/cahp-pearl-core-yosys.json", "yosys-json/cahp-ruby-core-yosys.json",
    "yosys-json/addr-4bit-yosys.json", "yosys-json/counter-4bit-yosys.json",
    # Iyokan-L1 netlists the reference's templated unit tests read (src/test0.cpp:157-432)
    "iyokanl1-json/pass-4bit-iyokanl1.json", "iyokanl1-json/and-4bit-iyokanl1.json", "iyokanl1-json/and-4_2bit-iyokanl1.json",
    "iyokanl1-json/mux-4bit-iyokanl1.json", "iyokanl1-json/addr-4bit-iyokanl1.json", "iyokanl1-json/register-4bit-iyokanl1.json",
    "iyokanl1-json/counter-4bit-iyokanl1.json",
    "in/test09.in", "in/test08.in", "in/test04.in", "in/test13.in",
    "out/test09-pearl.out", "out/test09-ruby.out", "out/test08.out", "out/test04.out", "out/test13.out",
]


def main():
    for rel in ASSETS:
        src, dst = REF / rel, OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        if src.suffix == ".json":
            with open(src, "rb") as f, gzip.GzipFile(str(dst) + ".gz", "wb", mtime=0) as g:
                shutil.copyfileobj(f, g)
        else:
            shutil.copyfile(src, dst)
    print("wrote", len(ASSETS), "assets to", OUT)


def materialise(dst: Path) -> Path:
    """Unpacks ref_assets into `dst` with the reference's directory layout (blueprints use ../yosys-json/ paths)."""
    for p in OUT.rglob("*"):
        if p.is_dir():
            continue
        out = dst / p.relative_to(OUT)
        out.parent.mkdir(parents=True, exist_ok=True)
        if p.suffix == ".gz":
            with gzip.open(p, "rb") as g, open(out.with_suffix(""), "wb") as f:
                shutil.copyfileobj(g, f)
        else:
            shutil.copyfile(p, out)
    return dst


if __name__ == "__main__":
    main()
