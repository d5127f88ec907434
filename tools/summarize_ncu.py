"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py full  gpurun_out/r01_s1_full.ncu-rep  profiles/r01_s1_kernels  "128x128 C64 G4"
    python tools/summarize_ncu.py list  gpurun_out/r01_launches.csv     profiles/r01_launch_shares
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/smem data pipe % busy"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed_op_shared_atom.sum", "ATOMS instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "ATOMS wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (must be 0: gather op)"),
])
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full(rep, dst, shape, dtype="f32"):
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    md = [f"# ncu --set full --clock-control none, {shape}, {dtype}, batch 16 (source: {rep})", ""]
    traffic = {}
    for r in rows:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("dcnv3::", "")
        md += [f"## {name}", "", "| metric | value |", "|---|---|"]
        for m, label in METRICS.items():
            if m in idx:
                md.append(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        stalls = {h[len(STALLS):]: float(r[idx[h]]) for h in hdr if h.startswith(STALLS) and not h.endswith("not_issued")}
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        md.append("| warp stall samples (top) | " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top) + " |")
        md.append("")
        def num(m):
            v, u = float(r[idx[m]]), units[idx[m]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        short = name.split("<")[0].replace("_kernel", "")
        traffic[f"{dtype}:{short} {shape}"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    open(dst + ".md", "w").write("\n".join(md) + "\n")
    return traffic


def launch_list(path, dst):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    # columns: ID, PID, Process, Host, Kernel Name, Context, Stream, Block, Grid, Device, CC, Section, Metric, Unit, Value
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("dcnv3::", "")
        v = float(r[-1].replace(",", ""))
        unit = r[-2]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    md = [f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares) -- {path}",
          "", "| kernel | launches | total us | share |", "|---|---|---|---|"]
    ours = {k: v for k, v in agg.items() if k.startswith(("fwd_", "bwd_", "merge_", "redo_", "amax_", "fixed_"))}
    tot_ours = sum(v[1] for v in ours.values()) or 1.0
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| {name} | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
    md += ["", "The `at::` / softmax / fill kernels are torch generating the synthetic inputs, outside the timed region.",
           "Shares among the library's own kernels (the ones inside the timed step):", "",
           "| kernel | share of the step's kernels |", "|---|---|"]
    for name, (n, us) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| {name} | {100 * us / tot_ours:.1f}% |")
    open(dst + ".md", "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        t = full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "f32")
        tp = "profiles/traffic.json"
        try:
            cur = json.load(open(tp))
        except OSError:
            cur = {}
        cur.update(t)
        json.dump(cur, open(tp, "w"), indent=1, sort_keys=True)
    else:
        launch_list(sys.argv[2], sys.argv[3])
