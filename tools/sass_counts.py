"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass, sm_100a): the evidence that the conv kernels are
tcgen05/TMA code (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk,
HMMA = mma.sync, LDSM = ldmatrix, SYNCS = mbarrier ops, MUFU.EX2 = ex2).   python tools/sass_counts.py [lib.so] > profiles/rNN_sass_counts.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "leastereo_b200", "_C", "libleastereo_b200.so")
COLS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "HMMA", "LDSM", "SYNCS", "MUFU.EX2", "ATOMG", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return out[:len(names)]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = {"name": m.group(1), "n": 0, **{c: 0 for c in COLS}}
            kernels.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["n"] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                cur[c] += 1
    names = demangle([k["name"] for k in kernels])
    print("SASS instruction counts per kernel of %s (cuobjdump -sass, sm_100a; tools/sass_counts.py)." % os.path.relpath(LIB, ROOT))
    print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, HMMA = mma.sync,")
    print("LDSM = ldmatrix, SYNCS = mbarrier ops, MUFU.EX2 = ex2.  The ConvBR forward / data-gradient kernels are tcgen05-only; the only")
    print("kernels with HMMA are the weight-gradient kernels (lea_wgrad_mma_kernel, see lea_wgrad_mma.cu for why).  UTMASTG = 0: the")
    print("TMA-store epilogue was measured and rejected (profiles/r02_tma_store_epilogue_ab.txt).")
    print()
    print("%-110s %7s " % ("kernel", "instr") + " ".join("%8s" % c for c in COLS))
    for k, nm in sorted(zip(kernels, names), key=lambda t: t[1]):
        nm = re.sub(r"\(anonymous namespace\)::", "", nm)
        nm = re.sub(r"\(.*", "", nm)
        print("%-110s %7d " % (nm[:110], k["n"]) + " ".join("%8d" % k[c] for c in COLS))


if __name__ == "__main__":
    main()
