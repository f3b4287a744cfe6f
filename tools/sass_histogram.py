"""Per-kernel SASS opcode histogram of the built library (runs on the CPU box):

    python tools/sass_histogram.py [diffbir_b200/libdiffbir_b200.so] > profiles/r02_sass_histogram.txt

Lists, for every kernel in the cubin, the instruction count and the Blackwell-specific opcodes that prove
the code path: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA
tensor loads / stores), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus legacy HMMA (must be 0)."""
import collections
import re
import subprocess
import sys

KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "FFMA2", "MUFU",
       "LDG", "STG", "LDS", "STS", "SHFL", "ATOM", "RED", "BAR", "ELECT", "UCGABAR"]


def main(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    kern, counts, order = None, {}, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            order.append(kern)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and kern:
            counts[kern][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
    print(f"# cuobjdump -sass {path}: {len(order)} kernels")
    print(f"# {'instr':>6} " + " ".join(f"{k:>7}" for k in KEY) + "  kernel")
    tot = collections.Counter()
    for k, name in zip(order, demangle or order):
        c = counts[k]
        tot.update(c)
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        name = re.sub(r"\(.*", "", name)
        print(f"  {sum(c.values()):6d} " + " ".join(f"{c.get(x, 0):7d}" for x in KEY) + f"  {name[:90]}")
    print(f"# total {sum(tot.values())} instructions; " + ", ".join(f"{x} {tot.get(x, 0)}" for x in KEY if tot.get(x, 0)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "diffbir_b200/libdiffbir_b200.so")
