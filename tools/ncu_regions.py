"""Aggregate an .ncu-rep's source-level samples / instructions by enclosing function of solver_core.cuh.
usage: ncu_regions.py rep"""
import csv, collections, subprocess, sys, io, re, os
rep = sys.argv[1]
root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
src_lines = open(os.path.join(root, "igt_mpc_int_b200/csrc/solver_core.cuh")).read().split("\n")
marks = []
for i, l in enumerate(src_lines, 1):
    m = re.match(r'\s*IGT_HDN? .*?(\w+)\(', l) or re.match(r'__device__ __forceinline__ .*?(\w+)\(', l)
    if m and not l.strip().endswith(';'):
        marks.append((i, m.group(1)))
def region(f, ln):
    if f != "solver_core.cuh":
        return f
    r = "?"
    for i, n in marks:
        if i <= ln: r = n
        else: break
    return r
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hd = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"): cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hd = r; continue
    if hd and len(r) == len(hd):
        d = {}
        for k, v in zip(hd, r):
            if k not in d: d[k] = v
        try: ln = int(d["Line No"])
        except Exception: continue
        a = agg[region(cur, ln)]
        a[0] += int(d.get("Instructions Executed") or 0); a[1] += int(d.get("# Samples") or 0)
        a[2] += int(d.get("Thread Instructions Executed") or 0)
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print("region                     inst%  samples%  lanes")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k:26s} {v[0]/ti*100:6.2f} {v[1]/ts*100:8.2f} {v[2]/max(v[0],1):6.1f}")
