#!/usr/bin/env python
"""Sample / instruction density of an .ncu-rep's SASS in windows of N instructions (which role of a warp-specialised
kernel is busy, which one waits):  python tools/ncu_regions.py gpurun_out/x.ncu-rep [window]"""
import collections, csv, io, subprocess, sys

def main(rep, win=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    src = list(csv.reader(io.StringIO(out)))
    SH = src[1]
    data = [r for r in src[2:] if len(r) >= len(SH)]
    isamp, isrc, iex = SH.index("# Samples"), SH.index("Source"), SH.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(SH) if h.startswith("stall_") and "Not Issued" not in h]
    S = [int(r[isamp] or 0) for r in data]
    EX = [int(r[iex] or 0) for r in data]
    T = sum(S) or 1
    for a in range(0, len(data), win):
        b = min(a + win, len(data))
        if sum(S[a:b]) < 0.003 * T and sum(EX[a:b]) < 0.003 * sum(EX):
            continue
        ops, st = collections.Counter(), collections.Counter()
        for i in range(a, b):
            toks = data[i][isrc].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            ops[op.split(".")[0]] += EX[i]
            for c in stall_cols:
                st[SH[c][6:]] += int(data[i][c] or 0)
        print(f"[{a:4d},{b:4d}) samples {100 * sum(S[a:b]) / T:5.1f}%  warp-instr {sum(EX[a:b]) / 1e6:7.1f} M  "
              + ", ".join(f"{k} {v / 1e6:.1f}" for k, v in ops.most_common(5)) + "  | "
              + ", ".join(f"{k} {100 * v / T:.1f}" for k, v in st.most_common(3)))

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
