import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, ctypes as C
from scipy.special import polygamma
from pydeseq2_b200 import _lib
ctx=_lib.Context(0); L=ctx.lib; cd=_lib.c_dptr
for n in (20000, 160000, 1000000):
    rng=np.random.default_rng(0)
    means=np.exp(rng.normal(4,2,n)*np.log(2)); gw=(4/means+0.1)*np.exp(rng.normal(0,0.5,n))
    dm,dg,dout,dfit=ctx.malloc(n*8),ctx.malloc(n*8),ctx.malloc(128),ctx.malloc(n*8)
    ctx.h2d(dm,means); ctx.h2d(dg,gw); ctx.sync()
    out=np.zeros(16)
    for rep in range(3):
        ctx.record(0); ctx.check(L.pdq_trend_fit_dev(ctx.h,cd(dm),cd(dg),n,1e-8,200.0,float(polygamma(1,99)),cd(dout),cd(dfit))); ctx.record(1); ctx.sync()
        ms=ctx.elapsed_ms(0,1)
    ctx.d2h(out,dout); ctx.sync()
    print(n, f"{ms:.3f} ms", "c", out[:2], "rounds", out[3], "passes", out[5], "sq", out[8])
