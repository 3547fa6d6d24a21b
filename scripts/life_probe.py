"""Life-ordered rounds: time of sweep shards and the headline beam against the weight of in-layer segments in the predicted
cost (TORJ_LIFE_HARM_COST, read by the library at every trace).  python scripts/life_probe.py out.json"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
from torj_jl_b200.distributed import shard_block_cyclic
import bench

L = tj.lib()
ctx = _lib.context()
tj.abs_Al_init(24)
PSI = np.linspace(0.0, 1.0, 1000)
pl = tj.Plasma(*tj.solovev_arrays(257, 257).values())
out = {}


def run(pos, dirs, w, reps=1, **kw):
    opt = tj.default_options(lanes_per_ray=1, **kw)
    ms = []
    for _ in range(reps + 1):
        r = tj.trace_bundle(pl, pos, dirs, w, 95e9, 1, 1.0, PSI, options=opt)
        t = C.c_double()
        _lib.check(L.torj_ctx_last_trace_ms(ctx, C.byref(t)))
        ms.append(t.value)
    return min(ms[1:]), r


pa, da, wa = bench.sweep_bundle()
pb, db, wb = bench.beam_bundle("beam64k")
gammas = [float(g) for g in os.environ.get("GAMMAS", "0,0.5,1,2,4").split(",")]
for g in gammas:
    os.environ["TORJ_LIFE_HARM_COST"] = str(g)
    row = {}
    for N in (16, 8):
        idx = shard_block_cyclic(len(wa), 1025, 0, N)
        row[f"sweep_1/{N}"], r = run(pa[idx], da[idx], wa[idx])
    row["beam64k"], r = run(pb, db, wb, reps=2)
    out[str(g)] = row
    print(g, json.dumps(row), flush=True)
os.environ["TORJ_LIFE_HARM_COST"] = os.environ.get("BEST", "1")
row = {}
for N in (4, 1):
    idx = shard_block_cyclic(len(wa), 1025, 0, N)
    row[f"sweep_1/{N}"], r = run(pa[idx], da[idx], wa[idx])
out["best_" + os.environ["TORJ_LIFE_HARM_COST"]] = row
print(json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
