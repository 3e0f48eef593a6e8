"""
Turns the raw outputs of tools/campaign_n1.sh (bench JSON, ncu launch list, one `ncu --set full` report) into the
tracked summaries under profiles/:  r01_launches_cfg2.csv (copied), r01_ncu_full_selected_metrics.csv and the tables
of profiles/README.md (printed to stdout as markdown).  Needs `ncu` on PATH to read the .ncu-rep.
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(name):
    return re.sub(r"^void ", "", re.sub(r"\(.*", "", name))


def launch_table():
    lines = [l for l in open(os.path.join(OUT, "r01_launches_cfg2.csv")) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | mean us | total us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, v[0], v[1] / v[0], v[1], 100 * v[1] / tot))
    shutil.copy(os.path.join(OUT, "r01_launches_cfg2.csv"), os.path.join(PROF, "r01_launches_cfg2.csv"))
    return "\n".join(out)


def full_table():
    raw = subprocess.run(["ncu", "-i", os.path.join(OUT, "r01_full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    col = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    sel = [("time us", "gpu__time_duration.sum"), ("dram rd MB", "dram__bytes_read.sum"), ("dram wr MB", "dram__bytes_write.sum"),
           ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
           ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
           ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           ("ipc/sm", "sm__inst_executed.avg.per_cycle_active"), ("warp instr", "smsp__inst_executed.sum"),
           ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct")]
    seen, table = set(), []
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        vals = [r[col[m]] if m in col else "" for _, m in sel]
        st = sorted(((float(r[col[s]] or 0), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls),
                    reverse=True)[:4]
        table.append((name, vals, ", ".join("%s %.1f" % (n, v) for v, n in st)))
    with open(os.path.join(PROF, "r01_ncu_full_selected_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [a for a, _ in sel] + ["top stalls (warp cycles per issue)"])
        for name, vals, st in table:
            w.writerow([name] + vals + [st])
    out = ["| kernel | " + " | ".join(a for a, _ in sel) + " | top stalls (cycles per issue) |", "|---|" + "---:|" * len(sel) + "---|"]
    for name, vals, st in table:
        fmt = []
        for v in vals:
            try:
                x = float(v)
                fmt.append("%.4g" % x)
            except ValueError:
                fmt.append(v)
        out.append("| `%s` | " % name + " | ".join(fmt) + " | " + st + " |")
    return "\n".join(out)


def full_table_r02(rep, out_csv):
    """Round 2: selected metrics of one `ncu --set full --clock-control none` capture per kernel -> profiles/<out_csv>
    (read by bench.py for `roofline.traffic`, `fp64_pipe_frac` and the counted FP64 work) and a markdown table."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    col = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    sel = [("time us", "gpu__time_duration.sum"), ("dram rd MB", "dram__bytes_read.sum"), ("dram wr MB", "dram__bytes_write.sum"),
           ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
           ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warp instr", "smsp__inst_executed.sum"),
           ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct")]
    # thread-level FP64 instructions per elapsed cycle (summed over the SMSPs) x elapsed cycles = executed FP64 operations
    fl = ["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
          "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"]
    seen, table = set(), []
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        vals = [r[col[m]] if m in col else "" for _, m in sel]
        try:
            gflop = (2 * float(r[col[fl[0]]]) + float(r[col[fl[1]]]) + float(r[col[fl[2]]])) * float(r[col["smsp__cycles_elapsed.avg"]]) / 1e9
        except (KeyError, ValueError):
            gflop = ""
        st = sorted(((float(r[col[s]] or 0), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls),
                    reverse=True)[:4]
        table.append((name, vals + [gflop], ", ".join("%s %.1f" % (n, v) for v, n in st)))
    names = [a for a, _ in sel] + ["fp64 Gflop"]
    with open(os.path.join(PROF, out_csv), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + names + ["top stalls (warp cycles per issue)"])
        for name, vals, st in table:
            w.writerow([name] + vals + [st])
    out = ["| kernel | " + " | ".join(names) + " | top stalls (cycles per issue) |", "|---|" + "---:|" * len(names) + "---|"]
    for name, vals, st in table:
        fmt = []
        for v in vals:
            try:
                fmt.append("%.4g" % float(v))
            except ValueError:
                fmt.append(str(v))
        out.append("| `%s` | " % name + " | ".join(fmt) + " | " + st + " |")
    return "\n".join(out)


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 2 and sys.argv[1] == "r02":
        print(full_table_r02(sys.argv[2], sys.argv[3]))
        sys.exit(0)
    d = json.load(open(os.path.join(OUT, "r01_bench_n1.json")))
    shutil.copy(os.path.join(OUT, "r01_bench_n1.json"), os.path.join(PROF, "r01_bench_n1.json"))
    print("### launch list\n")
    print(launch_table())
    print("\n### full-set capture\n")
    print(full_table())
    print("\n### bench\n")
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "lm_iters_per_s", "jacobian_pass_ms")}))
