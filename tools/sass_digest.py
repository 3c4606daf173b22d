#!/usr/bin/env python
"""Per-kernel SASS digest of the shipped library (no GPU needed): which instructions prove what the kernels are.

    python tools/sass_digest.py [path/to/libvis_b200.so] > profiles/rNN_sass_digest.txt

For every kernel in the cubin: registers, shared memory, instruction count, and the counts of the mnemonics that matter
for this path — UBLKCP (cp.async.bulk: the 1-D bulk-copy engine that stages row segments), SYNCS (mbarrier), STG.E.128 /
LDG.E.128 (128-bit global accesses), IMAD / IDP (the integer MACs), PRMT (byte unpack), LDS / STS, NANOSLEEP (mbarrier
poll back-off), IMMA / I2IP (the integer tensor path and the saturating pack of the 9..33-tap kernel k_fused_mma: the ONLY
kernel that may hold IMMA — its passes are banded u8 x byte-limb products) — and, as negative evidence,
the other tensor-core / tensor-map mnemonics (HMMA, UTC*MMA, UTMALDG): nothing else on this path is a contraction.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = ["UBLKCP", "SYNCS", "STG.E.128", "LDG.E.128", "STG", "LDG", "IMAD", "IDP", "PRMT", "LDS", "STS", "NANOSLEEP", "BAR",
        "SHFL", "VIMNMX", "ATOM", "RED", "IMMA", "I2IP"]
FORBIDDEN = ["HMMA", "DMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "HGMMA", "LDTM", "STTM"]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except Exception:
        return names


def main():
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "vision-inspection-system_b200" / "libvis_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(lib)], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            reg = re.search(r"REG:(\d+)", line)
            sh = re.search(r"SHARED:(\d+)", line)
            usage[cur] = (int(reg.group(1)) if reg else -1, int(sh.group(1)) if sh else 0)
            cur = None
    kernels, name = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            c = kernels[name]
            c["_total"] += 1
            c[op.split(".")[0]] += 1
            if op.startswith("STG.E.128") or op.startswith("LDG.E.128"):
                c[op[:9]] += 1
    names = list(kernels)
    pretty = demangle(names)
    print(f"# SASS digest of {lib.name}: {len(names)} kernels, architectures in the fatbin: {sorted(arch)}")
    print(f"# columns: regs smem(static) instrs | " + " ".join(WANT))
    bad_total = collections.Counter()
    order = sorted(range(len(names)), key=lambda i: pretty[i])
    for i in order:
        c = kernels[names[i]]
        reg, sh = usage.get(names[i], (-1, 0))
        short = pretty[i].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        short = re.sub(r"\((int|bool|unsigned int|long|char)\)", "", short)     # template value casts
        short = re.sub(r"\(.*", "", short)                                        # parameter list
        print(f"{short:70s} {reg:4d} {sh:6d} {c['_total']:6d} | " + " ".join(f"{c[w]:5d}" for w in WANT))
        for f in FORBIDDEN:
            if c[f]:
                bad_total[f] += c[f]
        if c["IMMA"] and "k_fused_mma" not in short:
            bad_total["IMMA outside k_fused_mma"] += c["IMMA"]
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# library totals: " + ", ".join(f"{w} {tot[w]}" for w in WANT))
    print("# floating-point tensor-core / tensor-map mnemonics, and IMMA outside k_fused_mma (must be absent): " +
          (", ".join(f"{k} {v}" for k, v in bad_total.items()) if bad_total else "none"))


if __name__ == "__main__":
    main()
