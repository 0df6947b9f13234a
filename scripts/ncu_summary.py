"""Summarise an .ncu-rep: key raw metrics + the hottest SASS lines.  ncu_summary.py file.ncu-rep [ntop]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pat = re.compile(r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum|lts__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__throughput.avg.pct_of_peak_sustained_elapsed|sm__issue_active.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__pipe_tensor_subpipe.*pct.*|sm__inst_executed_pipe_(lsu|alu|fma|fmaheavy|xu|uniform|tensor.*).sum|smsp__inst_executed.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|sm__cycles_elapsed.max|smsp__cycles_active.avg|lts__t_sector_hit_rate.pct|sm__pipe_tensor_op.*)$")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("KERNEL", d.get("Kernel Name", "")[:110])
    for k in hdr:
        if pat.match(k):
            print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, ins, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[2:]:
    if len(r) < len(h):
        continue
    try:
        data.append((int(r[ins] or 0), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
agg = {}
for n, r in data:
    for i in stall_cols:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8], "samples", tot)
for n, r in sorted(data, key=lambda x: -x[0])[:ntop]:
    st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:7d} {100*n/tot:5.1f}% ex={r[iex]:>9s} {r[ia][-5:]} {r[isrc][:64]:64s} {st}")
