"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses, from `cuobjdump -sass` of the
in-tree library (runs without a GPU):  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor loads, UBLKCP = bulk (non-tensor) TMA
copies, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, LDGSTS = cp.async.
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import re
import subprocess
import sys
from collections import OrderedDict

LIB = sys.argv[1] if len(sys.argv) > 1 else "procyon_b200/libprocyon_b200.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "LDSM", "MUFU.EX2",
        "ATOM", "RED", "BAR.SYNC", "ACQBULK", "SYNCS"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in KEYS}
            kernels[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["instructions"] += 1
        for k in KEYS:
            if k == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    kernels[cur][k] += 1
            elif op == k or op.startswith(k + "."):
                kernels[cur][k] += 1
    names = demangle(list(kernels))
    print(f"# {LIB}: {len(kernels)} kernels (cuobjdump -sass, sm_100a).  Columns: instruction counts in the kernel's SASS.")
    cols = ["instructions"] + KEYS
    print(f"{'kernel':78s} " + " ".join(f"{c[:9]:>9s}" for c in cols))
    for k, v in kernels.items():
        n = names.get(k, k)
        n = n.replace("pcy::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("pcy::", "")
        n = re.sub(r"\(.*", "", n.replace("void ", ""))
        if len(n) > 77:
            n = n[:74] + "..."
        print(f"{n:78s} " + " ".join(f"{v[c]:9d}" for c in cols))


if __name__ == "__main__":
    main()
