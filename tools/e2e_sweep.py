#!/usr/bin/env python
"""e2e_sweep.py -- raw PCIe copy rates (H2D, D2H, both at once) and hostpipe.HostSuite over slot count x workspace size, one GPU."""
import sys, os, time
ROOT=os.getcwd(); sys.path[:0]=[os.path.join(ROOT,'earthkit-meteo_b200')]
import numpy as np, torch
from ek_thermo import hostpipe
n=6599680*16
rng=np.random.default_rng(0)
tp,qp,pp=(hostpipe.pinned_empty(n) for _ in range(3))
tp[:]=rng.uniform(200,310,n); pp[:]=rng.uniform(1e3,1.05e5,n); qp[:]=rng.uniform(1e-6,0.02,n)
outs={k:hostpipe.pinned_empty(n) for k in ("theta","es","rh","td","tv")}
# raw copy rates
d=torch.empty(n,dtype=torch.float64,device='cuda'); h=torch.from_numpy(tp)
for name,fn in (("H2D",lambda: d.copy_(h,non_blocking=True)),("D2H",lambda: h.copy_(d,non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, 5*n*8/(time.perf_counter()-t0)/1e9,'GB/s')
s1,s2=torch.cuda.Stream(),torch.cuda.Stream(); h2=torch.from_numpy(qp); d2=torch.empty_like(d)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0; print('bidirectional each', 5*n*8/dt/1e9,'GB/s')
for slots in (2,3,4,6):
    for ws in (256,768,1536,3072):
        hs=hostpipe.HostSuite("cuda:0",workspace_bytes=ws<<20,n_slots=slots)
        hs.suite_tqp(tp,qp,pp,outputs=tuple(outs),out=outs); torch.cuda.synchronize()
        t0=time.perf_counter()
        for _ in range(3): hs.suite_tqp(tp,qp,pp,outputs=tuple(outs),out=outs)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
        print(f"slots={slots} workspace={ws}MB  {n/dt/1e9:.3f} Gpt/s  {64*n/dt/1e9:.1f} GB/s")
        del hs
