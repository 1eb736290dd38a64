import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I
text=I.queens(16)
for it in range(6):
    t0=time.perf_counter(); m=cb.Model(text); t1=time.perf_counter(); p=cb.GpuProblem(m); t2=time.perf_counter(); r=p.solve(); t3=time.perf_counter(); p.close(); m.close(); t4=time.perf_counter()
    print("iter %d parse %.2f load %.2f solve %.2f (device %.2f) close %.2f ms"%(it,(t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,r.kernel_ms+r.expand_ms,(t4-t3)*1e3), flush=True)
