#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full --import-source on` capture) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_coarse.ncu-rep > profiles/rNN_ncu_<kernel>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = (
    "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum,", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_tmem_ldt.sum ", "smsp__sass_inst_executed_op_tmem_stt.sum ",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
)


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = dict(zip(hdr, row)).get("Kernel Name", "?")
        print(f"== kernel: {name}")
        for h, u, v in zip(hdr, units, row):
            if any(h == k.strip().rstrip(",") for k in KEYS) or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                print(f"{h} [{u}] = {v}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    start = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[start]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in src[start + 1:] if len(r) == len(h)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    print(f"== source page: {len(data)} SASS instructions, {tot} warp-stall samples")
    agg = {n: sum(int(r[ix[n]] or 0) for r in data) for n in stall_cols}
    print("stall totals:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    ops = {}
    for r in data:
        op = r[ix["Source"]].replace("@!P0", "").replace("@P0", "").split()[0] if r[ix["Source"]].split() else "?"
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
    print("warp-instructions by opcode:", dict(sorted(ops.items(), key=lambda kv: -kv[1])[:24]))
    print("== top 30 SASS instructions by samples")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:30]:
        n = int(r[ix["# Samples"]] or 0)
        st = {c[6:]: int(r[ix[c]] or 0) for c in stall_cols if int(r[ix[c]] or 0) > 0.1 * max(n, 1)}
        print(f"{r[ix['Address']][-6:]} samples={n} ({100.0 * n / max(tot, 1):.1f}%) exec={r[ix['Instructions Executed']]} {r[ix['Source']][:80]} {st}")


if __name__ == "__main__":
    main()
