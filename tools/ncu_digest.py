"""Digest of an ncu report (--set full): key utilisation metrics + SASS opcode mix + top stall lines.
usage: python tools/ncu_digest.py report.ncu-rep [kernel-index]"""
import collections
import csv
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(rep, idx=0):
    rows = list(csv.reader(run([rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2 + idx]
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
    print("kernel:", r[hdr.index("Kernel Name")][:100])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k} [{units[i]}] = {r[i]}")
    src = list(csv.reader(run([rep, "--page", "source", "--csv"]).splitlines()))
    h = src[1]
    ie, ci = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    ops, tot, lines, agg = collections.Counter(), 0, [], collections.Counter()
    for row in src[2:]:
        if len(row) <= ie or row[0] == "Address":
            break
        try:
            n, s = int(row[ie]), int(row[ci])
        except ValueError:
            continue
        tok = row[1].strip().split()
        if not tok:
            continue
        op = (tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]).split(".")[0]
        ops[op] += n
        tot += n
        st = {h[i]: int(row[i]) for i in stall_cols if len(row) > i and row[i].isdigit() and int(row[i]) > 0}
        for a, b in st.items():
            agg[a] += b
        lines.append((s, row[1].strip()[:80], sorted(st.items(), key=lambda x: -x[1])[:2]))
    print(f"  warp instructions executed: {tot}")
    print("  opcode mix:", ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(14)))
    print("  stall reasons (samples):", ", ".join(f"{a} {b}" for a, b in agg.most_common(7)))
    lines.sort(reverse=True)
    for s, txt, st in lines[:12]:
        print(f"    {s:6d}  {txt}  {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
