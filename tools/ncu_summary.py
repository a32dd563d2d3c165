"""Summarise an ncu report (raw + source pages) into text: key metrics and the hottest instructions.

    python tools/ncu_summary.py gpurun_out/prof_local.ncu-rep [top_n]
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

KEYS = r"""Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|launch__registers_per_thread$|launch__grid_size|
launch__block_size|launch__occupancy_limit_(registers|shared_mem|warps)|sm__warps_active.avg.pct_of_peak_sustained_active|
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on.sum.per_cycle_elapsed$|sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained$|sm__inst_executed_pipe_(xu|alu|lsu|fma).avg.pct_of_peak_sustained_active|
smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum$|
smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|
dram__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum$|smsp__cycles_active.avg$|
local_(load|store)|smsp__inst_executed_op_local""".replace("\n", "")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, top=25):
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        for k in sorted(d):
            if re.search(KEYS, k) and d[k][0] not in ("", "0"):
                if "stalled" in k and float(d[k][0]) < 0.05:
                    continue
                print(f"{k:90s} {d[k][0]} {d[k][1]}")
        print("-" * 60)
    src = page(rep, "source")
    # one block per kernel: header line 'Kernel Name', then column header, then rows
    i = 0
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            print("KERNEL", src[i][1][:120])
            cols = src[i + 1]
            ix = {h: n for n, h in enumerate(cols)}
            j = i + 2
            rows = []
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                if len(src[j]) == len(cols):
                    rows.append(src[j])
                j += 1
            tot = sum(int(r[ix["# Samples"]]) for r in rows) or 1
            stall_cols = [c for c in cols if c.startswith("stall_") and "Not Issued" not in c]
            agg = Counter()
            for r in rows:
                for c in stall_cols:
                    agg[c] += int(r[ix[c]])
            print("stall samples:", [(k, f"{100 * v / tot:.1f}%") for k, v in agg.most_common(8)])
            ops = Counter()
            for r in rows:
                t = r[ix["Source"]].split()
                ops[t[1] if t[0].startswith("@") else t[0]] += int(r[ix["# Samples"]])
            print("by opcode:", [(k, f"{100 * v / tot:.1f}%") for k, v in ops.most_common(10)])
            for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
                s = int(r[ix["# Samples"]])
                why = sorted(((c, int(r[ix[c]])) for c in stall_cols), key=lambda kv: -kv[1])[:2]
                print(f"{100 * s / tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {why}")
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
