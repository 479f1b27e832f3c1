"""Summarise an .ncu-rep: key raw metrics, stall reasons, top source lines.

    ncu_summary.py rep [out_prefix [capture_key [accepted_steps]]]

With `capture_key` ("workload:trajectories", e.g. "rober:1048576") the numbers bench.py cannot measure itself -- DRAM bytes of the
kernel, FP64-pipe utilisation, executed FP64 operations -- are stored under that key in profiles/ncu_captures.json, from where
bench.py reports them for a run of the same workload at the same size."""
import csv, collections, json, os, subprocess, sys, io
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

if len(sys.argv) > 3:
    def val(name):
        return float(vals[hdr.index(name)].replace(",", "")) if name in hdr and vals[hdr.index(name)] not in ("", "no data") else None
    def unit(name):
        return units[hdr.index(name)]
    def to_bytes(name):
        v, u = val(name), unit(name)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    dur_ms = val("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit("gpu__time_duration.sum")]
    cyc = val("sm__cycles_elapsed.max")
    per = {op: val(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") or 0.0 for op in ("dfma", "dmul", "dadd")}
    flops = (2 * per["dfma"] + per["dmul"] + per["dadd"]) * cyc
    entry = {
        "source": os.path.basename(rep), "kernel": vals[hdr.index("Kernel Name")], "duration_ms": dur_ms,
        "dram_bytes": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
        "executed": {"fp64_flops": flops, "fp64_tflops": flops / (dur_ms * 1e-3) / 1e12,
                     "fp64_pipe_active_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                     "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "local_load_sectors": val("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
                     "local_store_sectors": val("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
                     "note": "ncu --set full capture (serialised, cold cache): shares and counts, not a timing"}}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_captures.json")
    tab = json.load(open(path)) if os.path.exists(path) else {}
    if len(sys.argv) > 4:   # accepted steps of the captured launch (from the run's own bench line): executed FLOPs per step
        entry["accepted_steps"] = float(sys.argv[4])
        entry["executed"]["fp64_flops_per_accepted_step"] = flops / float(sys.argv[4])
    tab[sys.argv[3]] = entry
    json.dump(tab, open(path, "w"), indent=1, sort_keys=True)
    print("\nstored", sys.argv[3], "->", path)
