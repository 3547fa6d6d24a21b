"""T(n) of the 1 M-ray sweep on ONE GPU: the block-cyclic shard a rank would get at N = 1, 2, 4, 8, 16 GPUs, hand-off schedule
with (0) and without (3) the tail stages. The strong-scaling curve of the kernel itself, without any communication."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torj_jl_b200 as tj
from torj_jl_b200 import _lib
from torj_jl_b200.distributed import shard_block_cyclic
import bench

L = tj.lib(); ctx = _lib.context(); tj.abs_Al_init(24)
PSI = np.linspace(0.0, 1.0, 1000)
pl = tj.Plasma(*tj.solovev_arrays(257, 257).values())
pa, da, wa = bench.sweep_bundle()
out = {}
for world in (16, 8, 4, 2, 1):
    idx = shard_block_cyclic(len(wa), 1025, 0, world)
    for sch in (3, 0):  # 3 plain hand-off, 0 life-ordered rounds
        ms = []
        for rep in range(3 if world > 1 else 2):
            r = tj.trace_bundle(pl, pa[idx], da[idx], wa[idx], 95e9, 1, 1.0, PSI, options=tj.default_options(schedule=sch, lanes_per_ray=1))
            t = C.c_double(); _lib.check(L.torj_ctx_last_trace_ms(ctx, C.byref(t))); ms.append(t.value)
        out[f"N{world}_sch{sch}"] = dict(ms=min(ms[1:]), rays=len(idx), steps=r["counters"]["n_acc"])
        print(world, sch, out[f"N{world}_sch{sch}"], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
