import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from directdemod_b200 import filters
torch.cuda.set_device(0)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[0], ts[len(ts)//2]
for n in (20000000, 60000000):
    x=torch.empty(n,dtype=torch.complex64,device="cuda"); torch.view_as_real(x).normal_(0,40)
    f=filters.butter(2400000,100000,n=8)
    c=filters.cascade([filters.butter(2400000,100000,n=8)],max_taps=1025)
    h=np.asarray(c.getB)
    p=filters.filter(h,[1.0]); ps=filters.filter(h,[1.0],storeState=False)
    rec=filters.butter(2400000,100000,n=8); rec._no_fir_form=True
    print(n, "iir via fir form", timeit(lambda: f._apply_dev(x)), "cascade obj", timeit(lambda: c._apply_dev(x)),
          "plain fir stateful", timeit(lambda: p._apply_dev(x)), "stateless", timeit(lambda: ps._apply_dev(x)),
          "recursion", timeit(lambda: rec._apply_dev(x)), flush=True)
