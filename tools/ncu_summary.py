"""Summarise an .ncu-rep (read here, no GPU): key throughput numbers, stall reasons, hottest source lines.
usage: python tools/ncu_summary.py rep.ncu-rep [--lines N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name", "?")[:80], " grid", d.get("Grid Size"), " block", d.get("Block Size"))
    for k in WANT:
        if k in d: print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
    stalls = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("_per_warp_active.pct") and d.get(h) not in ("", "n/a", None):
            try: stalls.append((float(d[h].replace(',', '')), h))
            except ValueError: pass
    for v, h in sorted(stalls, reverse=True)[:10]: print(f"  stall {h.split('stalled_')[-1][:50]:52s} {v:10.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
def f(s):
    try: return float((s or "0").replace(",", ""))
    except ValueError: return 0.0
cur_file, hdr2, lines = "?", None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr2 = r; continue
    if hdr2 is None or r[0] in ("", "Function Name") or len(r) != len(hdr2) or not r[0].isdigit(): continue
    csamp = hdr2.index("# Samples"); cinst = hdr2.index("Instructions Executed")
    lines.append((f(r[csamp]), f(r[cinst]), cur_file, r[0], r[1].strip()))
tot = sum(l[0] for l in lines) or 1; toti = sum(l[1] for l in lines) or 1
print(f"-- hottest CUDA source lines ({tot:.0f} samples, {toti:.4g} warp-instructions)")
for smp, ins, fl, ln, txt in sorted(lines, key=lambda l: -l[0])[:nlines]:
    print(f"  {100*smp/tot:5.1f}% smp {100*ins/toti:5.1f}% inst  {fl}:{ln:>4s}  {txt[:110]}")
