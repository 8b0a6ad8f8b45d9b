#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed): headline metrics, stall reasons and the
SASS opcode mix of the first kernel in the report.  Usage: python profiles/ncu_summary.py rep [out.md]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_shared_st.sum", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = []
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out.append(f"# ncu summary of `{rep}`\n")
    out.append(f"kernel: `{data[0][hdr.index('Kernel Name')]}`  ({len(data)} launches captured)\n")
    out.append("| metric | unit | per launch |\n|---|---|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {units[i]} | {', '.join(r[i] for r in data)} |")
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"hdr": None, "rows": []}
            secs.append(cur)
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = r
            continue
        cur["rows"].append(r)
    s = secs[0]
    idx = {n: i for i, n in enumerate(s["hdr"])}
    stalls = [n for n in s["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    byop, byop_t = collections.Counter(), collections.Counter()
    for r in s["rows"]:
        for n in stalls:
            try:
                tot[n] += float(r[idx[n]])
            except ValueError:
                pass
        op = [x for x in r[idx["Source"]].split() if not x.startswith("@")]
        o = op[0].split(".")[0] if op else "?"
        byop[o] += float(r[idx["Instructions Executed"]])
        byop_t[o] += float(r[idx["Thread Instructions Executed"]])
    S = sum(tot.values())
    out.append(f"\nSASS instructions in the kernel: {len(s['rows'])} ({len(s['rows']) * 16 / 1024:.0f} KiB)\n")
    out.append("| warp stall (sampled) | share |\n|---|---|")
    for n, v in tot.most_common(10):
        out.append(f"| {n} | {100 * v / S:.1f}% |")
    T = sum(byop.values())
    out.append(f"\nwarp-level instructions executed: {T:.0f}, thread-level: {sum(byop_t.values()):.0f}\n")
    out.append("| opcode | share of warp instr | avg active threads |\n|---|---|---|")
    for o, v in byop.most_common(16):
        out.append(f"| {o} | {100 * v / T:.1f}% | {byop_t[o] / max(v, 1):.1f} |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
