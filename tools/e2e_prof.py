import sys,time,os
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from sparse_gslam_b200 import SparseOptimizerB200, capi, graphgen as gg
g=gg.make_c5()
opt=SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC, device=0)
for i in range(3):
    t0=time.perf_counter(); opt.initialize_optimization(g); t1=time.perf_counter()
    n,_=opt.optimize(15); t2=time.perf_counter()
    opt.estimates(); t3=time.perf_counter()
    print("set_graph %.3f optimize %.3f estimates %.3f"%(t1-t0,t2-t1,t3-t2), file=sys.stderr)
