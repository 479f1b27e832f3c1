"""Summarise an .ncu-rep: key raw metrics, stall reasons, top source lines.  Usage: ncu_summary.py rep [out_prefix]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
def page(args):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--csv"] + args, capture_output=True, text=True).stdout)))
raw = page(["--page", "raw"])
hdr, units, vals = raw[0], raw[1], raw[2]
keep = ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','smsp__inst_executed.sum',
 'smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
out = []
for k in keep:
    if k in hdr:
        i = hdr.index(k); out.append((k, vals[i], units[i]))
for k,v,u in out: print(f"{k} = {v} {u}")
rows = page(["--page", "source", "--print-source", "cuda,sass"])
hi = next(i for i,r in enumerate(rows[:6]) if '# Samples' in r)
h = rows[hi]; si = h.index('# Samples'); ii = h.index('Instructions Executed')
stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
lines = []; tot = 0; st = collections.Counter()
for r in rows[hi+1:]:
    if len(r) <= si: continue
    if r[0] and r[0].isdigit():
        try: ns = int(r[si]); ni = int(r[ii]) if r[ii] not in ('-','') else 0
        except: continue
        lines.append((int(r[0]), r[1].strip()[:100], ns, ni)); tot += ns
    elif r[2] not in ('', '-', '...'):
        for c in stalls:
            try: st[c] += int(r[h.index(c)])
            except: pass
print("\nstall reasons (% of samples):")
ts = sum(st.values())
for c,v in st.most_common(9): print(f"  {c:26s} {100*v/max(ts,1):5.1f}%")
print("\ntop source lines (% samples, instructions executed):")
for ln, src, ns, ni in sorted(lines, key=lambda x: -x[2])[:32]:
    print(f"  {ln:5d} {100*ns/tot:5.1f}% {ni:12d}  {src}")
if len(sys.argv) > 2:
    with open(sys.argv[2] + "_metrics.csv", "w") as f:
        w = csv.writer(f); w.writerow(["metric","value","unit"]); w.writerows(out)
