import time, numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcppml_b200 as rb
from rcppml_b200 import synth
m,n,k=1_000_000,100_000,64
eng=rb.Engine(0)
eng.set_matrix_synthetic_sharded(m,n,1e-3,synth.SEED_A)
eng.init_factors(k,42,0)
p,i,x=eng.get_matrix(); W0,H0,_=eng.get_factors(); eng.close()
x64=x.astype(np.float64); W64=W0.astype(np.float64); H64=H0.astype(np.float64)
for rep in range(2):
    e=rb.Engine(0)
    t=time.perf_counter(); e.set_matrix(m,n,p,i,x64); t1=time.perf_counter()
    import ctypes as C
    from rcppml_b200 import _lib
    _lib.check(e._lib.rcppml_b200_set_factors_f64(e._h,k,W64.ctypes.data_as(C.POINTER(C.c_double)),H64.ctypes.data_as(C.POINTER(C.c_double))),"x"); e.k=k
    t2=time.perf_counter()
    r=e.fit(rb.make_config(k,max_iter=10,tol=0,solver_mode=1)); t3=time.perf_counter()
    W,H,d=e.get_factors(); t4=time.perf_counter()
    print(f"set_matrix {t1-t:.3f} set_factors {t2-t1:.3f} fit {t3-t2:.3f} (loop {r.loop_ms:.1f} ms) get {t4-t3:.3f}")
    e.close()
from rcppml_b200 import bridge
for rep in range(2):
    W=W64.copy(); H=H64.copy()
    call=bridge.PackedCall(p,i,x64,m,n,k,W,H,max_iter=10,tol=0.0,solver_mode=1,verbose=True)
    t=time.perf_counter(); call(); print('bridge call', time.perf_counter()-t, call.status)
