"""Join an ncu SASS-level source page with nvdisasm line info: samples / executed instructions per CUDA source line.
usage: ncu_by_line.py report.ncu-rep libtorj_cuda.so kernel_mangled_substr [top]"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel's .text section in order, each tagged with the last seen (file, line) [inlined-at chain ignored]
insts, cur, on = [], ("?", 0), False
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
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        insts.append(cur)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr, data = rows[hi], rows[hi + 1:]
iS, iE, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
print(f"sass lines in report {len(data)}, in disassembly {len(insts)}")
n = min(len(data), len(insts))
samp, exe = collections.Counter(), collections.Counter()
for k in range(n):
    samp[insts[k]] += int(data[k][iS]); exe[insts[k]] += int(data[k][iE])
ts, te = sum(samp.values()), sum(exe.values())
src_cache = {}
def src(f, l):
    for d in ("torj_jl_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:100] if 0 < l <= len(src_cache[p]) else ""
    return ""
print(f"total samples {ts}  executed warp-instructions {te}")
for (f, l), v in sorted(exe.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v/te*100:5.2f}% inst {samp[(f,l)]/ts*100:5.2f}% samp  {f}:{l:<4d} {src(f,l)}")
