"""Summarise an .ncu-rep: headline metrics + hottest CUDA source lines (development aid)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        try:
            if float(v) < 0.05 and h.startswith("smsp__average"): continue
        except ValueError: pass
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; out = []; ts = ti = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 12 or r[0] in ("Line No", ""): continue
    try: ln = int(r[0]); s = int(r[4]); i = int(r[7]); t = int(r[8])
    except ValueError: continue
    out.append((s, i, t, cur, ln, r[1].strip()[:95])); ts += s; ti += i
out.sort(reverse=True)
print(f"-- total samples {ts}, warp instructions {ti}")
for s, i, t, f, l, code in out[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% inst eff={t/max(i,1)/32:4.2f} {f}:{l} {code}")
