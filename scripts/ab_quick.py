"""Quick A/B of library builds: headline beam and the 1/8 sweep shard only (see ab_time.py for the full table).
  python scripts/ab_quick.py out.json lib [lib ...]"""
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

    def run(pl, pos, dirs, w, f, reps=2, **kw):
        opt = tj.default_options(**kw)
        ms = []
        for _ in range(reps + 1):
            r = tj.trace_bundle(pl, pos, dirs, w, f, 1, 1.0, PSI, options=opt)
            t = C.c_double()
            _lib.check(L.torj_ctx_last_trace_ms(ctx, C.byref(t)))
            ms.append(t.value)
        return min(ms[1:]), r

    pl = tj.Plasma(*tj.solovev_arrays(257, 257).values())
    pos, dirs, w = bench.beam_bundle("beam64k")
    res["beam64k"], r = run(pl, pos, dirs, w, 95e9, reps=3, lanes_per_ray=1)
    res["beam64k_dep"] = r["deposited_power"]
    res["beam64k_exact"], r = run(pl, pos, dirs, w, 95e9, reps=1, lanes_per_ray=1, alpha_floor=0.0)
    pa, da, wa = bench.sweep_bundle()
    idx = shard_block_cyclic(len(wa), 1025, 0, 8)
    res["sweep8"], r = run(pl, pa[idx], da[idx], wa[idx], 95e9, reps=2, lanes_per_ray=1)
    ps, ds, ws = bench.beam_bundle("small")
    res["small1025"], r = run(pl, ps, ds, ws, 95e9)
    print("AB_RESULT " + json.dumps(res), flush=True)


def main():
    out = sys.argv[1]
    table = {}
    for lib in sys.argv[2:]:
        env = dict(os.environ, TORJ_CUDA_LIB=os.path.abspath(lib))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True, timeout=600)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("AB_RESULT ")]
        table[os.path.basename(lib)] = json.loads(line[-1][10:]) if line else {"error": (p.stderr or p.stdout)[-2000:]}
        print(os.path.basename(lib), json.dumps(table[os.path.basename(lib)]), flush=True)
    json.dump(table, open(out, "w"), indent=1)


if __name__ == "__main__":
    one() if sys.argv[1:] == ["--one"] else main()
