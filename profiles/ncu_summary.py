#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, stall mix, per-phase instruction mix.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rNN_name.txt]
"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = page(rep, "raw")
    H, U, V = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "launch__shared_mem_per_block_dynamic"]
    print("== kernel:", V[H.index("Kernel Name")] if "Kernel Name" in H else "?")
    for w in want:
        if w in H:
            i = H.index(w)
            print(f"{w:85s} {U[i]:16s} {V[i]}")
    print("\n== warp stall reasons (per issue-active ratio)")
    for i, h in enumerate(H):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(V[i])
            except ValueError:
                continue
            if v >= 0.05:
                print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")

    src = page(rep, "source")
    if len(src) < 3:
        return
    SH = src[1]
    data = [r for r in src[2:] if len(r) >= len(SH)]
    isamp, isrc, iex = SH.index("# Samples"), SH.index("Source"), SH.index("Instructions Executed")
    S = [int(r[isamp] or 0) for r in data]
    EX = [int(r[iex] or 0) for r in data]
    T = sum(S) or 1
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[isrc]]
    print(f"\n== SASS: {len(data)} instructions, {sum(EX) / 1e6:.1f} M warp-instructions executed, {T} samples")
    print("   regions split at BAR.SYNC:", bars)
    cuts = [0] + bars + [len(data)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b - a < 8:
            continue
        ops = collections.Counter()
        for i in range(a, b):
            toks = data[i][isrc].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            ops[op.split(".")[0]] += EX[i]
        s = sum(S[a:b])
        print(f"   [{a:4d},{b:4d}) samples {100 * s / T:5.1f}%  warp-instr {sum(EX[a:b]) / 1e6:7.1f} M  "
              + ", ".join(f"{k} {v / 1e6:.1f}" for k, v in ops.most_common(10)))
    stall_cols = [i for i, h in enumerate(SH) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        for i in stall_cols:
            tot[SH[i]] += int(r[i] or 0)
    print("\n== sampled stall mix: " + ", ".join(f"{k[6:]} {100 * v / T:.1f}%" for k, v in tot.most_common(8)))
    print("\n== top 25 SASS instructions by samples")
    for idx in sorted(range(len(data)), key=lambda i: -S[i])[:25]:
        st = sorted(((SH[i], int(data[idx][i] or 0)) for i in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"   {idx:5d} {100 * S[idx] / T:4.1f}% ex={EX[idx]:9d} {data[idx][isrc][:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1])
