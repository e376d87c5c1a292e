"""Iyokan blueprints and netlist files -> one flat Netlist (the loader in front of the hot path).

Reads what `iyokan plain|tfhe --blueprint X.toml` reads:
  * TOML blueprint: [[file]] (type = "yosys-json" | "iyokanl1-json", path relative to the blueprint),
    [[builtin]] (type = "mux-rom" | "mux-ram"), [connect] edges with `@port[i]` external ports
    (NetworkBlueprint, src/iyokan.hpp:1731-1895);
  * Yosys JSON netlists (YosysJSONReader::read, src/iyokan.hpp:2130-2351): $_AND_ ... $_MUX_, $_NOT_, $_DFF_P_,
    constant drivers, ports named `clock` dropped;
  * IyokanL1 JSON netlists (IyokanL1JSONReader::read, :2354-2482).
and performs what the front end's constructor does before the DAG reaches the scheduler
(TFHEppFrontend, src/iyokan_tfhepp.cpp:312-458): every sub-network is instantiated into ONE node space and
every [connect] edge turns the consumer's INPUT wire into an alias of its producer.  MUX ROM / RAM builtins are
generated (netlist.mux_rom / netlist.mux_ram: the circuits of iyokan.hpp:2538-2593, 2646-2762); CMUX-memory
builtins (type "rom" / "ram") are functionally identical and are evaluated as MUX memories as well.
"""
from __future__ import annotations

import json
import re
import tomllib
from pathlib import Path

from .netlist import NetBuilder, mux_ram, mux_rom

YOSYS_GATES = {"$_AND_": "AND", "$_NAND_": "NAND", "$_ANDNOT_": "ANDNOT", "$_OR_": "OR", "$_NOR_": "NOR",
               "$_ORNOT_": "ORNOT", "$_XOR_": "XOR", "$_XNOR_": "XNOR"}


class SubNet:
    """Ports of one sub-network inside the shared builder: name -> node id per bit."""

    def __init__(self):
        self.inputs: dict[str, dict[int, int]] = {}
        self.outputs: dict[str, dict[int, int]] = {}


def read_yosys(b: NetBuilder, path) -> SubNet:
    mod = next(iter(json.load(open(path))["modules"].values()))
    sn = SubNet()
    driver: dict[int, int] = {}
    for pname, port in mod["ports"].items():
        if port["direction"] != "input" or pname == "clock":
            continue
        for i, bit in enumerate(port["bits"]):
            n = b._add(32)  # INPUT
            sn.inputs.setdefault(pname, {})[i] = n
            driver[bit] = n
    cells = list(mod["cells"].values())
    ids = []
    for c in cells:
        t, con = c["type"], c["connections"]
        if t in YOSYS_GATES:
            n = b.gate(YOSYS_GATES[t], -1, -1)
        elif t == "$_NOT_":
            n = b.gate("NOT", -1)
        elif t == "$_MUX_":
            n = b.gate("MUX", -1, -1, -1)
        elif t == "$_DFF_P_":
            n = b.dff(-1)
        else:
            raise ValueError(f"unsupported cell type {t}")
        ids.append(n)
        driver[con["Q" if t == "$_DFF_P_" else "Y"][0]] = n
    const = {}

    def src(bit, port=False):
        if isinstance(bit, str):  # constant driver "0" / "1": allowed on output ports only (iyokan.hpp:2170-2193)
            if not port:
                raise ValueError("Connection of cells to a constant driver is not implemented.")  # iyokan.hpp:2124
            if bit not in const:
                const[bit] = b.gate("CONST1" if bit == "1" else "CONST0")
            return const[bit]
        if bit not in driver:
            raise ValueError(f"Invalid JSON of network: signal {bit} has no driver")
        return driver[bit]

    for c, n in zip(cells, ids):
        t, con = c["type"], c["connections"]
        if t == "$_DFF_P_":
            b.ins[n] = [src(con["D"][0]), -1, -1]
        elif t == "$_NOT_":
            b.ins[n] = [src(con["A"][0]), -1, -1]
        elif t == "$_MUX_":
            b.ins[n] = [src(con["A"][0]), src(con["B"][0]), src(con["S"][0])]
        else:
            b.ins[n] = [src(con["A"][0]), src(con["B"][0]), -1]
    for pname, port in mod["ports"].items():
        if port["direction"] != "output":
            continue
        for i, bit in enumerate(port["bits"]):
            sn.outputs.setdefault(pname, {})[i] = b._add(34, src(bit, port=True))  # OUTPUT wire
    return sn


def read_iyokanl1(b: NetBuilder, path, mem_name="ram", width=None) -> SubNet:
    obj = json.load(open(path))
    sn = SubNet()
    node = {}
    for p in obj["ports"]:
        if p["type"] == "input":
            node[p["id"]] = b._add(32)
            sn.inputs.setdefault(p["portName"], {})[p["portBit"]] = node[p["id"]]
    kinds = {"DFFP": None, "RAM": None}
    for c in obj["cells"]:
        t = c["type"]
        if t in ("DFFP", "RAM"):
            n = b.dff(-1)
            if t == "RAM":
                lst = b.mem.setdefault(mem_name, [])
                idx = c["ramAddress"] * width + c["ramBit"]
                lst.extend([-1] * (idx + 1 - len(lst)))
                lst[idx] = n
        elif t == "MUX":
            n = b.gate("MUX", -1, -1, -1)
        elif t == "NOT":
            n = b.gate("NOT", -1)
        else:
            n = b.gate(t, -1, -1)
        node[c["id"]] = n
    del kinds
    for c in obj["cells"]:
        t, i, n = c["type"], c["input"], node[c["id"]]
        if t in ("DFFP", "RAM"):
            b.ins[n] = [node[i["D"]], -1, -1]
        elif t == "NOT":
            b.ins[n] = [node[i["A"]], -1, -1]
        elif t == "MUX":
            b.ins[n] = [node[i["A"]], node[i["B"]], node[i["S"]]]
        else:
            b.ins[n] = [node[i["A"]], node[i["B"]], -1]
    for p in obj["ports"]:
        if p["type"] == "output":
            sn.outputs.setdefault(p["portName"], {})[p["portBit"]] = b._add(34, node[p["bits"][0]])
    return sn


PORT_RE = re.compile(r"^(@?)([^/\[\]]+)(?:/([^\[\]]+))?(?:\[(\d+)(?::(\d+))?\])?$")


def parse_ref(s):
    m = PORT_RE.match(s.strip())
    if m is None:
        raise ValueError(f"Invalid port string: {s}")   # NetworkBlueprint::parsePortString, iyokan.hpp:1697
    ext, a, bname, lo, hi = m.groups()
    lo = int(lo) if lo is not None else 0
    hi = int(hi) if hi is not None else lo
    if ext:
        return ("@", a, list(range(lo, hi + 1)))
    return (a, bname, list(range(lo, hi + 1)))


def read_blueprint(toml_path, mux_ram_json_dir=None):
    """Flatten a blueprint into one Netlist (what TFHEppFrontend's constructor does, iyokan_tfhepp.cpp:312-458).

    mux_ram_json_dir: directory holding `mux-ram-A-W-R.min.json` netlists to use for mux-ram builtins instead of
    the generated circuit (the reference embeds such pre-synthesised files; same function, different gate count)."""
    toml_path = Path(toml_path)
    bp = tomllib.load(open(toml_path, "rb"))
    b = NetBuilder()
    subs: dict[str, SubNet] = {}
    for f in bp.get("file", []):
        path = (toml_path.parent / f["path"]).resolve()
        if f["type"] == "yosys-json":
            subs[f["name"]] = read_yosys(b, path)
        elif f["type"] == "iyokanl1-json":
            subs[f["name"]] = read_iyokanl1(b, path)
        else:
            raise ValueError(f["type"])
    for bi in bp.get("builtin", []):
        name = bi["name"]
        sn = SubNet()
        # CMUX memories (type "rom" / "ram": TRLWE-packed contents read with CMUX trees after circuit bootstrapping,
        # src/iyokan_tfhepp.hpp:194-889) have the same ports and the same function as the MUX memories; this back-end
        # evaluates them AS MUX memories (more gate bootstraps, no lvl02 kernels), using the per-bit TLWE form of the
        # contents that `iyokan-packet enc` always writes next to the TRLWE form (romInTLWE / ramInTLWE).
        if bi["type"] in ("mux-rom", "rom"):
            addr = [b._add(32) for _ in range(bi["in_addr_width"])]
            sn.inputs["addr"] = dict(enumerate(addr))
            outs = mux_rom(b, addr, bi["out_rdata_width"], name=name)
            sn.outputs["rdata"] = {i: b._add(34, o) for i, o in enumerate(outs)}
        elif bi["type"] in ("mux-ram", "ram"):
            if bi["type"] == "ram":
                b.write_through.append(name)
            a, w, r = bi["in_addr_width"], bi["in_wdata_width"], bi["out_rdata_width"]
            assert w == r
            pre = Path(mux_ram_json_dir) / f"mux-ram-{a}-{w}-{r}.min.json" if mux_ram_json_dir else None
            if pre is not None and pre.exists():
                sn = read_iyokanl1(b, pre, mem_name=name, width=w)
            else:
                addr = [b._add(32) for _ in range(a)]
                wren = b._add(32)
                wdata = [b._add(32) for _ in range(w)]
                sn.inputs["addr"], sn.inputs["wren"] = dict(enumerate(addr)), {0: wren}
                sn.inputs["wdata"] = dict(enumerate(wdata))
                outs = mux_ram(b, addr, wren, wdata, name=name)
                sn.outputs["rdata"] = {i: b._add(34, o) for i, o in enumerate(outs)}
        else:
            raise ValueError(f"unknown builtin type {bi['type']!r}")
        subs[name] = sn
    def port(net, kind, name, bit):
        """Node of net/name[bit]; the reference dies with these messages in TaskNetwork::get (iyokan.hpp:955-980)."""
        if net not in subs:
            raise ValueError(f"Invalid network name: {net}")
        table = subs[net].inputs if kind == "input" else subs[net].outputs
        if name not in table or bit not in table[name]:
            raise ValueError(f"Invalid {kind} port: {net}/{name}[{bit}]")
        return table[name][bit]

    connect = dict(bp.get("connect", {}))
    togrnd = connect.pop("TOGND", [])
    for dst, src in connect.items():
        if not isinstance(src, str) or not dst or not src or (dst[0] == "@" and src[0] == "@"):
            raise ValueError(f"Invalid connect: {dst} = {src}")
        d, s = parse_ref(dst), parse_ref(src)
        if len(d[2]) != len(s[2]):
            raise ValueError(f"Invalid connect: {dst} = {src}")
        for db, sb in zip(d[2], s[2]):
            if d[0] == "@":      # external output  "@out[i]" = "net/port[j]"
                lst = b.out_ports.setdefault(d[1], [])
                lst.extend([-1] * (db + 1 - len(lst)))
                lst[db] = port(s[0], "output", s[1], sb)
            elif s[0] == "@":    # external input   "net/port[i]" = "@in[j]"
                lst = b.in_ports.setdefault(s[1], [])
                lst.extend([-1] * (sb + 1 - len(lst)))
                node = port(d[0], "input", d[1], db)
                if lst[sb] == -1:
                    lst[sb] = node
                else:            # one external bit feeding several inputs: alias the later ones
                    b.alias(node, lst[sb])
            else:                # internal edge: the consumer's INPUT wire becomes an alias of the producer
                b.alias(port(d[0], "input", d[1], db), port(s[0], "output", s[1], sb))
    # TOGND = ["@port[n:m]", ...]: bits that belong to an external port (they count for its width, hence for the
    # stride of its bit stream) but are connected to nothing (iyokan.hpp:1809-1825)
    for ref in togrnd:
        if not isinstance(ref, str) or not ref.startswith("@"):
            raise ValueError(f"Invalid port name for TOGND: {ref}")
        _, name, bits = parse_ref(ref)
        lst = b.out_ports[name] if name in b.out_ports else b.in_ports.setdefault(name, [])
        lst.extend([-1] * (max(bits) + 1 - len(lst)))
    return b.build()
