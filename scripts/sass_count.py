"""Multiply-pipe instruction count of the CMUX-step loop of a rotation kernel, from the SASS of the built library.

usage: python scripts/sass_count.py [lib.so] [mangled-name substring, default br7_kernelILi8ELi8]
The step loop is the widest backward branch of the function; every instruction inside it is counted once (the inner loops
of the kernels are fully unrolled or run once per step).  Pipe cycles per warp-instruction (B300_MICROARCH.md, measured
with scripts/microbench/pipes.cu): IMAD 2, IMAD.HI / IMAD.WIDE 4, IMAD.IADD / IMAD.SHL / IMAD.MOV 2."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "iyokan_b200/csrc/libb200fhe.so"
want = sys.argv[2] if len(sys.argv) > 2 else "br7_kernelILi8ELi8"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
body = next(f for f in funcs if f.startswith("_Z") and want in f.split("\n")[0])
ins = []
for ln in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
best = (0, 0, 0)
for addr, text in ins:
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr and addr - tgt > best[0]:
            best = (addr - tgt, tgt, addr)
_, lo, hi = best
cnt = collections.Counter()
for addr, text in ins:
    if lo <= addr <= hi:
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
        cnt[op] += 1
total = sum(cnt.values())
def grp(pred): return sum(v for k, v in cnt.items() if pred(k))
imad_hi = grp(lambda k: k.startswith("IMAD.HI"))
imad_wide = grp(lambda k: k.startswith("IMAD.WIDE"))
imad_misc = grp(lambda k: k.startswith(("IMAD.IADD", "IMAD.SHL", "IMAD.MOV")))
imad = grp(lambda k: k.startswith("IMAD")) - imad_hi - imad_wide - imad_misc
cycles = 2 * imad + 4 * imad_hi + 4 * imad_wide + 2 * imad_misc
print(f"{want}: loop {lo:#x}..{hi:#x}, {total} instructions per thread and step")
print(f"  IMAD {imad}  IMAD.HI {imad_hi}  IMAD.WIDE {imad_wide}  IMAD.IADD/SHL/MOV {imad_misc}  -> {cycles} multiply-pipe cycles per warp and step")
print("  LDS", grp(lambda k: k.startswith("LDS")), " STS", grp(lambda k: k.startswith("STS")), " LDG", grp(lambda k: k.startswith("LDG")),
      " LDL", grp(lambda k: k.startswith("LDL")), " STL", grp(lambda k: k.startswith("STL")), " BAR", grp(lambda k: k.startswith("BAR")))
