"""Summarise an .ncu-rep (run here, no GPU): key raw metrics + hottest SASS lines + opcode mix per kernel."""
import csv, subprocess, sys, io, re
from collections import Counter
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "."
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_tex_throttle",
        "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_sleeping"]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    if not re.search(pat, d["Kernel Name"]):
        continue
    print("==", d["Kernel Name"][:100])
    for k in KEYS:
        if k in d:
            print(f"  {k:85s} {d[k]:>16s} {u[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections = []  # (kernel name, header index map, rows)
name = "?"
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1] if len(r) > 1 else "?"
    elif r and r[0] == "Address":
        sections.append((name, {x: i for i, x in enumerate(r)}, []))
    elif sections and r and r[0].startswith("0x"):
        sections[-1][2].append(r)
for name, h, data in sections:
    if not re.search(pat, name) or not data:
        continue
    S, E, SRC = h["# Samples"], h["Instructions Executed"], h["Source"]
    tot = sum(int(r[S]) for r in data); ti = sum(int(r[E]) for r in data)
    print(f"-- source page of {name[:80]}: {len(data)} SASS lines, {tot} samples, {ti} warp instructions executed")
    for r in sorted(data, key=lambda r: -int(r[S]))[:25]:
        print(f"  {int(r[S]):8d} {100*int(r[S])/max(tot,1):5.1f}%  exec {int(r[E]):10d}  {r[SRC][:90]}")
    c = Counter(); cs = Counter()
    for r in data:
        op = [o for o in r[SRC].split() if not o.startswith("@")][0].split(".")[0]
        c[op] += int(r[E]); cs[op] += int(r[S])
    print("  executed by opcode:", [(k, round(v / 1e6, 1)) for k, v in c.most_common(22)])
    print("  samples by opcode :", cs.most_common(14))
