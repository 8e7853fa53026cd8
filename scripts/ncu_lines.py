"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source sass,cuda) per
source line and per engine phase: instructions executed and stall samples.  usage: ncu_lines.py REP [top [kernel-name-substring]]"""
import csv, subprocess, sys, collections, re, io
def I(x):
    try: return int(x)
    except ValueError: return 0
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fn_filter = sys.argv[3] if len(sys.argv) > 3 else None   # substring of the kernel name to keep
import os
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] + os.environ.get("NCU_SELECT", "").split(), capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(out)))
cur_file = None; cols = None
agg = collections.defaultdict(lambda: [0, 0, 0.0, ""])   # (file,line) -> inst, samples, thread_inst, text
kernel = 0
first_fn = None
for r in lines:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if fn_filter is not None: kernel = 1 if fn_filter in r[1] else 2
        else:
            if first_fn is None: first_fn = r[1]
            kernel = 1 if r[1] == first_fn else 2
        continue
    if r[0] == "Line No": cols = {c: i for i, c in enumerate(r)}; continue
    if cols is None or kernel > 1: continue
    if r[0] != "":   # a source line row (aggregated over its SASS)
        key = (cur_file, int(r[0]))
        a = agg[key]
        a[0] += I(r[cols["Instructions Executed"]])
        a[1] += I(r[cols["# Samples"]])
        a[2] += I(r[cols["Thread Instructions Executed"]])
        a[3] = r[1].strip()
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total inst %d samples %d" % (tot_i, tot_s))
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]
for (f, ln), a in rows:
    print("%5.1f%% inst %5.1f%% samp  thr/inst %4.1f  %s:%d  %s" % (100 * a[0] / tot_i, 100 * a[1] / max(tot_s, 1), a[2] / max(a[0], 1), f, ln, a[3][:90]))
# per phase of lux_step_core.cuh by line ranges given as markers
core = [l for l in open("luxpythonenvgym_b200/csrc/lux_step_core.cuh")]
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(core) if l.strip().startswith("// ----") or l.strip().startswith("// W1") or l.strip().startswith("// C:") or l.strip().startswith("// W2")]
marks.append((len(core) + 1, "end"))
print("---- per phase (lux_step_core.cuh)")
for (a0, name), (a1, _) in zip(marks, marks[1:]):
    i = sum(v[0] for (f, ln), v in agg.items() if f == "lux_step_core.cuh" and a0 <= ln < a1)
    s = sum(v[1] for (f, ln), v in agg.items() if f == "lux_step_core.cuh" and a0 <= ln < a1)
    if i: print("%5.1f%% inst %5.1f%% samp  %s" % (100 * i / tot_i, 100 * s / max(tot_s, 1), name))
for fn in sorted(set(f for f, _ in agg)):
    i = sum(v[0] for (f, ln), v in agg.items() if f == fn); s = sum(v[1] for (f, ln), v in agg.items() if f == fn)
    print("file %-28s %5.1f%% inst %5.1f%% samp" % (fn, 100 * i / tot_i, 100 * s / max(tot_s, 1)))
