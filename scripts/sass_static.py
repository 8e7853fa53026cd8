"""Static SASS size of one kernel per source line / engine phase (no GPU needed).
  python scripts/sass_static.py [kernel-substring]       default: the 32x32 small-class plain kernel
Extracts the cubins of luxpythonenvgym_b200/_lux_b200.so and reads nvdisasm -g line annotations."""
import collections, os, re, subprocess, sys, tempfile
kern = sys.argv[1] if len(sys.argv) > 1 else "lux_step_kernelILb1ELi32ELi1ELi32ELi1E"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "luxpythonenvgym_b200", "_lux_b200.so")], cwd=tmp, capture_output=True)
text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, "lux_step.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(text) if l.startswith(".text.") and kern in l)
end = next((i for i in range(start + 1, len(text)) if text[i].startswith("//--------------------- .text")), len(text))
cur = ("?", 0)
per = collections.Counter()
n = 0
stack = []
for l in text[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f = os.path.basename(m.group(1)); ln = int(m.group(2))
        # "inlined at" chains: attribute to the outermost lux_* file line if present
        cur = (f, ln)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        per[cur] += 1; n += 1
print("kernel %s: %d SASS instructions, %.1f KB" % (kern, n, n * 16 / 1024))
core = open(os.path.join(root, "luxpythonenvgym_b200/csrc/lux_step_core.cuh")).read().splitlines()
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(core) if l.strip().startswith("// ----") or l.strip().startswith("// W1") or l.strip().startswith("// C:") or l.strip().startswith("// W2")]
marks.append((len(core) + 1, "end"))
for (a0, name), (a1, _) in zip(marks, marks[1:]):
    k = sum(v for (f, ln), v in per.items() if f == "lux_step_core.cuh" and a0 <= ln < a1)
    if k: print("%5d instr %5.1f KB  %s" % (k, k * 16 / 1024, name))
byfile = collections.Counter()
for (f, ln), v in per.items(): byfile[f] += v
for f, v in byfile.most_common(): print("file %-28s %5d instr %5.1f KB" % (f, v, v * 16 / 1024))
print("top lines:")
for (f, ln), v in per.most_common(25):
    src = core[ln - 1].strip()[:90] if f == "lux_step_core.cuh" and ln <= len(core) else ""
    print("%5d  %s:%d  %s" % (v, f, ln, src))
