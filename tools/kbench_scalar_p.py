import sys, os, json
sys.path[:0]=[os.path.join(os.getcwd(),'earthkit-meteo_b200')]
import torch
from ek_thermo import fused, thermo
n=6599680*24; dev='cuda:0'
g=torch.Generator(device=dev).manual_seed(0)
t=torch.empty(n,device=dev,dtype=torch.float64).uniform_(230,300,generator=g)
q=torch.empty(n,device=dev,dtype=torch.float64).uniform_(1e-6,5e-3,generator=g)
out={k:torch.empty_like(t) for k in fused.ALL7_TQP}
def timed(fn, bpp, name):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/10
    print(f"{name:44s} {ms:7.3f} ms frac={bpp*n/ms/1e6/6551:.3f}")
timed(lambda: fused.suite_tqp(t,q,85000.0,out=out), 8*7, "suite5, scalar p (pressure level)")
timed(lambda: fused.suite_tqp(t,q,85000.0,outputs=fused.ALL7_TQP,out=out), 8*9, "suite7, scalar p")
timed(lambda: thermo.potential_temperature(t,85000.0), 16, "theta, scalar p")
timed(lambda: thermo.relative_humidity_from_specific_humidity(t,q,85000.0), 24, "rh, scalar p")
timed(lambda: thermo.ept_from_specific_humidity(t,q,85000.0), 24, "ept, scalar p")
