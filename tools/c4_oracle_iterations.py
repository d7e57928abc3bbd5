import sys, time, json, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import pyoracle as O
from pfemfort_b200 import driver as D, mesh as M, solver as S
t=time.time()
m = M.gen_tetra(-0.5, 0.5, 50, 0.0, 6.0, 300, -0.5, 0.5, 50, dbc="clamp_y0", ndof=3)
kind = S.ELASTICITY_TETRA
num = D.number(m, kind)
rp, col = O.pattern(num.elemDof, num.size_global)
print('pattern', col.size, time.time()-t, flush=True)
val, rhs, nbad = O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA, rp, col)
print('assembled', time.time()-t, flush=True)
res={}
for th in (1, 4):
    x, its, reason, rn = O.cg_jacobi(rp, col, val, rhs, rtol=1e-10, max_it=200000, threads=th)
    res[th]=(its, reason)
    print('threads', th, 'its', its, 'reason', reason, time.time()-t, flush=True)
json.dump(res, open('/tmp/c4_oracle.json','w'))
