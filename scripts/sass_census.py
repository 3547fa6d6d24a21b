"""SASS census of a kernel: static instruction counts by opcode class from cuobjdump and, when an ncu report with the source
page is given, the executed (dynamic) counts of the same classes. For this non-contraction FP64 kernel the evidence that
matters is DFMA/DMUL/DADD, MUFU.*64H (the rcp/rsqrt seeds), LDG.E.128 (stencil), LDS/STS (stage derivatives) and LDL/STL
(spills) — not tcgen05/TMA.  usage: sass_census.py lib.so mangled_kernel_name [report.ncu-rep]"""
import collections, csv, re, subprocess, sys

so, kern = sys.argv[1:3]
rep = sys.argv[3] if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", "-fun", kern, so], capture_output=True, text=True).stdout
ops = []
for ln in txt.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", ln)
    if m:
        t = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
        ops.append(t.split()[0])


def cls(op):
    b = op.split(".")[0]
    if b in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): return b
    if op.startswith("MUFU"): return "MUFU." + ".".join(op.split(".")[1:])
    if b in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "LDCU", "ATOMS", "ATOMG", "RED", "ATOM"):
        return b + (".128" if ".128" in op else (".64" if ".64" in op else ""))
    if b in ("SHFL", "VOTE", "BRA", "CALL", "RET", "BSSY", "BSYNC", "BAR", "EXIT", "WARPSYNC"): return b
    if b in ("FFMA", "FMUL", "FADD", "HFMA2", "HMMA", "IMMA", "UTCHMMA", "UTMALDG", "TCGEN05"): return b
    if b in ("IMAD", "IADD3", "LOP3", "SHF", "ISETP", "SEL", "FSEL", "MOV", "UMOV", "VIADD", "LEA", "PRMT", "I2F", "F2I", "F2F", "I2FP", "R2UR", "CS2R", "S2R", "UIADD3", "ULOP3", "USHF", "UISETP", "PLOP3", "FSETP", "IABS", "IMNMX", "VIMNMX", "UIMAD", "ULEA", "USEL", "UPLOP3"):
        return "int/move/select"
    return "other:" + b


static = collections.Counter(cls(o) for o in ops)
dyn = None
if rep:
    t2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(t2.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    ie = hdr.index("Instructions Executed")
    if len(data) == len(ops):
        dyn = collections.Counter()
        for o, r in zip(ops, data):
            dyn[cls(o)] += int(r[ie] or 0)
print(f"kernel {kern}: {len(ops)} SASS instructions")
tot = sum(dyn.values()) if dyn else 0
print(f"{'class':18s} {'static':>7s}" + (f" {'executed (warp-instr)':>22s} {'share':>7s}" if dyn else ""))
for k, v in sorted(static.items(), key=lambda kv: -(dyn[kv[0]] if dyn else kv[1])):
    print(f"{k:18s} {v:7d}" + (f" {dyn[k]:22d} {dyn[k] / tot * 100:6.2f}%" if dyn else ""))
for must_be_zero in ("HMMA", "IMMA", "UTCHMMA", "TCGEN05", "UTMALDG"):
    assert static[must_be_zero] == 0
print("tensor-core / TMA instructions: none (nothing on this path is a dense contraction; tables are gathered through L1 with LDG.E.128)")
