"""A/B timing of kernel variants on one GPU. Each variant is a separately built library (TORJ_CUDA_LIB) timed in its own
process: device time of the trace stage(s) from torj_ctx_last_trace_ms, best of `reps` after one warm-up.
  python scripts/ab_time.py out.json [lib ...]        # libs default to torj_jl_b200/libtorj_cuda.so
  python scripts/ab_time.py --one                     # (internal) time the library named by TORJ_CUDA_LIB
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torj_jl_b200 as tj
    from torj_jl_b200 import _lib
    from torj_jl_b200.distributed import shard_block_cyclic
    import bench
    L = tj.lib()
    ctx = _lib.context()
    tj.abs_Al_init(24)
    PSI = np.linspace(0.0, 1.0, 1000)
    res = {}

    def run(pl, pos, dirs, w, f, reps=2, beam_id=None, n_beams=1, **kw):
        opt = tj.default_options(**kw)
        ms, r = [], None
        for _ in range(reps + 1):
            r = tj.trace_bundle(pl, pos, dirs, w, f, 1, 1.0, PSI, options=opt, beam_id=beam_id, n_beams=n_beams)
            t = C.c_double()
            _lib.check(L.torj_ctx_last_trace_ms(ctx, C.byref(t)))
            ms.append(t.value)
        return min(ms[1:] or ms), r

    pl = tj.Plasma(*tj.solovev_arrays(257, 257).values())
    pos, dirs, w = bench.beam_bundle("beam64k")
    for sch in (3, 0):
        ms, r = run(pl, pos, dirs, w, 95e9, reps=3, schedule=sch, lanes_per_ray=1)
        res[f"beam64k_sch{sch}"] = ms
        res["beam64k_dep"] = r["deposited_power"]
    ms, r = run(pl, pos, dirs, w, 95e9, reps=1, schedule=3, lanes_per_ray=1, alpha_floor=0.0)
    res["beam64k_exact"] = ms
    ms, r = run(pl, pos, dirs, w, 95e9, reps=2, lanes_per_ray=1, beam_id=np.zeros(len(w), dtype=np.int32), n_beams=2)
    res["beam64k_global_atomics"] = ms  # n_beams > 1 books straight into global memory: all 65 543 rays into ONE row
    pa, da, wa = bench.sweep_bundle()
    idx = shard_block_cyclic(len(wa), 1025, 0, 8)
    for sch in (3, 0):
        ms, r = run(pl, pa[idx], da[idx], wa[idx], 95e9, reps=2, schedule=sch, lanes_per_ray=1)
        res[f"sweep8_sch{sch}"] = ms
    bid = (np.arange(len(idx)) // 1025).astype(np.int32)
    ms, r = run(pl, pa[idx], da[idx], wa[idx], 95e9, reps=1, beam_id=bid, n_beams=int(bid.max()) + 1, lanes_per_ray=1)
    res["sweep8_per_beam"] = ms
    ps, ds, ws = bench.beam_bundle("small")
    for lanes in (1, 8, 32):
        ms, r = run(pl, ps, ds, ws, 95e9, lanes_per_ray=lanes)
        res[f"small1025_lanes{lanes}"] = ms
    plh = tj.Plasma(*tj.solovev_arrays(257, 257, Te0=10e3).values())
    P, D, W, F, B = [], [], [], [], []
    b = 0
    for f in (110e9, 170e9):
        for z0, pol in ((0.4, 30.0), (-0.4, -30.0)):
            p, d, ww = tj.launch_peripheral_rays(np.array([2.5, 0.0, z0]), tj.pol_tor_angles_2_vector(np.deg2rad(pol), 0.1), 0.0174,
                                                 1 / 3.99, f, N_rings=7, min_azimuthal_points=20)
            P.append(p); D.append(d); W.append(ww); F.append(np.full(len(ww), f)); B.append(np.full(len(ww), b, dtype=np.int32)); b += 1
    P, D, W, F, B = map(np.concatenate, (P, D, W, F, B))
    for lanes in (1, 8, 32):
        ms, r = run(plh, P, D, W, F, reps=1, lanes_per_ray=lanes, beam_id=B, n_beams=4)
        res[f"scan4100_Te10_lanes{lanes}"] = ms
    ms, r = run(plh, P[:1025], D[:1025], W[:1025], F[:1025], reps=1, lanes_per_ray=32)
    res["hot1025_lanes32"] = ms
    print("AB_RESULT " + json.dumps(res), flush=True)


def main():
    out = sys.argv[1]
    libs = sys.argv[2:] or [os.path.join(ROOT, "torj_jl_b200", "libtorj_cuda.so")]
    table = {}
    for lib in libs:
        env = dict(os.environ, TORJ_CUDA_LIB=os.path.abspath(lib))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True, timeout=900)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("AB_RESULT ")]
        table[os.path.basename(lib)] = json.loads(line[-1][10:]) if line else {"error": (p.stderr or p.stdout)[-2000:]}
        print(os.path.basename(lib), json.dumps(table[os.path.basename(lib)]), flush=True)
    with open(out, "w") as fh:
        json.dump(table, fh, indent=1)


if __name__ == "__main__":
    one() if sys.argv[1:] == ["--one"] else main()
