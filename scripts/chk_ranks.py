import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torj_jl_b200 as tj
import bench
tj.abs_Al_init(24)
pl = tj.Plasma(*tj.solovev_arrays().values())
psi = np.linspace(0,1,1000)
pos,dirs,w = bench.bundle_for_rank(7)
r = tj.trace_bundle(pl,pos,dirs,w,95e9,1,1.0,psi)
bad = np.nonzero(r['status']!=0)[0]
print("bad", bad, r['status'][bad], "reason", r['P_deposited_ray'][bad])
for i in bad: print(i, repr(pos[i]), repr(dirs[i]))
