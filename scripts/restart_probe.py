import sys; sys.path.insert(0,'/root/repo')
import numpy as np, cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import synthetic as syn
ctx=pkg.Context(0)
M=8; n=M**3
rp,c,v=syn.convdiff3d_csr(M); op=pkg.DeviceOperator.from_csr(ctx,rp,c,v)
x0=syn.start_vector(n,seed=7); exact=syn.convdiff3d_eigenvalues(M,count=3)
es=pkg.ArnoldiEigenSolver(np.float64)
es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(20).setMaxIterations(20).setMaxEigenvalues(1)
for cyc in (1,2,4,6,10,20):
    es.setInitialVector(x0)
    es.computeWithRestarts(cyc)
    print(cyc, es.eigenvalues()[0], abs(es.eigenvalues()[0]-exact[0]), es.ritzResiduals())
print(exact)
