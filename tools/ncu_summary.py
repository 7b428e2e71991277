"""Summarise an .ncu-rep (raw metrics + hottest source lines) -> text.  usage: ncu_summary.py rep [out.txt]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct', 'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency',
        'sass__inst_executed_local', 'lts__t_sector_hit_rate', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__icc_request_hit_rate', 'launch__registers_per_thread ',
        'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'Kernel Name']
out = []
for h, u, v in zip(hdr, units, vals):
    if any(w in h + ' ' for w in want):
        out.append(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hd = None; data = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"): cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hd = r; continue
    if hd and len(r) == len(hd):
        d = {}
        for k, v in zip(hd, r):
            if k not in d: d[k] = v
        try: ln = int(d["Line No"])
        except Exception: continue
        data.append((cur, ln, int(d.get("Instructions Executed") or 0), int(d.get("# Samples") or 0),
                     int(d.get("Thread Instructions Executed") or 0), d["Source"]))
ti = sum(x[2] for x in data) or 1; ts = sum(x[3] for x in data) or 1
agg = collections.defaultdict(lambda: [0, 0, 0, ''])
for f, ln, ins, smp, thr, s in data:
    a = agg[(f, ln)]; a[0] += ins; a[1] += smp; a[2] += thr; a[3] = s
out.append("\nhottest source lines by stall samples (inst% / samples% / avg active lanes)")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    out.append(f"{f}:{ln:4d} inst {v[0]/ti*100:5.2f}% samp {v[1]/ts*100:5.2f}% lanes {v[2]/max(v[0],1):4.1f} | {v[3].strip()[:100]}")
out.append("\nhottest source lines by executed instructions")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    out.append(f"{f}:{ln:4d} inst {v[0]/ti*100:5.2f}% samp {v[1]/ts*100:5.2f}% lanes {v[2]/max(v[0],1):4.1f} | {v[3].strip()[:100]}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"summary of {rep} (ncu --set full --clock-control none --import-source on)\n" + txt + "\n")
