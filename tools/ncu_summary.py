"""Summarise an .ncu-rep (one kernel): headline metrics, SASS opcode mix, top stall lines.  Usage: ncu_summary.py rep [n_lines]"""
import csv, collections, re, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__warps_active.avg.per_cycle_active"]
for h, u, v in zip(hdr, units, vals):
    if h in want or re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h):
        print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
iS, iI, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
cnt, smp = collections.Counter(), collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) <= iI: continue
    op = r[iS].split()
    if not op: continue
    o = (op[0] if not op[0].startswith("@") else op[1]).rstrip(";")
    try:
        cnt[o] += int(r[iI]); smp[o] += int(r[iN])
    except ValueError:
        continue
    st = {h[i][6:]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
    lines.append((int(r[iN]), r[0], r[iS][:70], st))
tot, tots = sum(cnt.values()), sum(smp.values())
print(f"\ninstructions executed: {tot}   samples: {tots}")
for o, c in cnt.most_common(22):
    print(f"  {o:32s} {c:12d} {100*c/tot:5.1f}%   samples {smp[o]:7d} {100*smp[o]/max(tots,1):5.1f}%")
print("\ntop sampled instructions")
for n, addr, s, st in sorted(lines, reverse=True)[:topn]:
    top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"  {n:6d} {100*n/max(tots,1):4.1f}%  {s:70s} {top}")
