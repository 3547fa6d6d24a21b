"""Static estimate of the hot instruction footprint of k_trace in a NEW build, using per-source-line execution counts
from an ncu capture of an OLD build: a SASS instruction is 'hot' when its (file, line) executed on >= `thr` of the RHS
evaluations in the capture. Prints the number of 128-byte instruction lines holding at least one hot instruction and
the span they cover.
usage: hot_footprint.py old_sass.csv old.so new.so kernel_substr [thr]"""
import collections, csv, os, re, subprocess, sys, tempfile

def disasm(so, kern):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    out, cur, on = [], ("?", 0), False
    for ln in dis:
        if ln.startswith(".text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            out.append((int(m.group(1), 16), cur, ln.strip()))
    return out

csvf, old_so, new_so, kern = sys.argv[1:5]
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
old = disasm(old_so, kern)
rows = list(csv.reader(open(csvf)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr, data = rows[hi], rows[hi + 1:]
iE = hdr.index("Instructions Executed")
assert len(data) == len(old), (len(data), len(old))
# per source line: the max execution count of any of its instructions
line_exec = collections.Counter()
for (addr, loc, _), r in zip(old, data):
    line_exec[loc] = max(line_exec[loc], int(r[iE]))
rhs = max(line_exec.values())  # proxy: hottest non-loop line... refine: use the median of the top stencil lines
rhs = sorted(line_exec.values())[-200]
new = disasm(new_so, kern)
for name, prog in (("old", old), ("new", new)):
    hot = [a for a, loc, _ in prog if line_exec.get(loc, 0) >= thr * rhs]
    lines = sorted({a // 128 for a in hot})
    print(f"{name}: {len(prog)} instrs, hot instrs {len(hot)}, hot 128B lines {len(lines)} ({len(lines) * 128 / 1024:.1f} KB), "
          f"span {(lines[-1] - lines[0] + 1) * 128 / 1024:.1f} KB")
