"""Offline numerics check (numpy float16 emulation) of the packed-half GELU used by the GEMM epilogues
(wavjepa_b200/csrc/ptx.cuh: gelu_h2 / gelu_h2_save): tanh-form fitted to the erf form, evaluated in fp16 pairs, against
the exact erf GELU (nn.GELU(approximate='none'), wavjepa/types/wavjepa_configs.py:37) and against the autocast-faithful
baseline (exact GELU of the bf16-rounded pre-activation, rounded to bf16).  Prints the fitted constants and rel-L2 errors.
    python scripts/check_gelu_h2.py"""
import numpy as np
from scipy.special import erf
rng=np.random.default_rng(0)
h=np.float16
def bf16(x):
    x=np.asarray(x,dtype=np.float32); u=x.view(np.uint32).astype(np.uint64)
    r=((u+0x7fff+((u>>16)&1))>>16<<16).astype(np.uint32); return r.view(np.float32)
def gelu_exact(x): return 0.5*x*(1+erf(x/np.sqrt(2)))
def dgelu_exact(x): return 0.5*(1+erf(x/np.sqrt(2)))+x*np.exp(-0.5*x*x)/np.sqrt(2*np.pi)
def tanh_f16(u):  # tanh.approx.f16x2: model as exact tanh rounded to f16 (+ up to ~2^-10.987 abs err)
    return h(np.tanh(u.astype(np.float64)))
def pipeline(a, c1, c3, c5=None):
    x=h(a)
    x2=h(x*x)
    if c5 is None:
        p=h(x2*h(c3)+h(c1))
    else:
        p=h(h(x2*h(c5)+h(c3))*x2+h(c1))
    u=h(x*p)
    t=tanh_f16(u)
    hp=h(t*h(0.5)+h(0.5))
    g=h(x*hp)
    if c5 is None:
        q=h(x2*h(3*c3)+h(c1))
    else:
        q=h(h(x2*h(5*c5)+h(3*c3))*x2+h(c1))
    s=h(h(1)-h(t*t))
    r=h(h(x*s)*q)
    d=h(r*h(0.5)+hp)
    return g.astype(np.float32), d.astype(np.float32)
def rel(a,b): return np.linalg.norm(a-b)/np.linalg.norm(b)
c1=np.sqrt(2/np.pi); c3=c1*0.044715
# fit better coefficients for erf-form: minimize |0.5(1+tanh(x(c1+c3x^2+c5x^4))) - Phi(x)|
from scipy.optimize import least_squares
xs=np.linspace(-6,6,4001)
Phi=0.5*(1+erf(xs/np.sqrt(2)))
def res3(c): return 0.5*(1+np.tanh(xs*(c[0]+c[1]*xs**2)))-Phi
f3=least_squares(res3,[c1,c3]).x
def res5(c): return 0.5*(1+np.tanh(xs*(c[0]+c[1]*xs**2+c[2]*xs**4)))-Phi
f5=least_squares(res5,[c1,c3,0.0]).x
print('fit3',f3,np.abs(res3(f3)).max(),'std tanh',np.abs(res3([c1,c3])).max()); print('fit5',f5,np.abs(res5(f5)).max())
for scale in (0.3,1.0,2.0):
    a=(rng.standard_normal(2_000_000)*scale).astype(np.float32)
    ab=bf16(a)
    ge=gelu_exact(ab.astype(np.float64)); de=dgelu_exact(ab.astype(np.float64))
    base_g=bf16(ge.astype(np.float32)); base_d=bf16(de.astype(np.float32))
    print(f'scale {scale}: baseline(bf16 rounding only) g {rel(base_g,ge):.2e} d {rel(base_d,de):.2e}')
    for name,cs in (('tanh-std',(c1,c3)),('fit3',tuple(f3)),('fit5',tuple(f5))):
        g,d=pipeline(a,*cs)
        # compare against exact gelu of the UNROUNDED a too (the fp32 oracle never rounds)
        ge2=gelu_exact(a.astype(np.float64)); de2=dgelu_exact(a.astype(np.float64))
        print(f'   {name}: g {rel(bf16(g),ge):.2e} d {rel(bf16(d),de):.2e} | vs unrounded-input exact: g {rel(bf16(g),ge2):.2e} (baseline {rel(base_g,ge2):.2e}) d {rel(bf16(d),de2):.2e} (baseline {rel(base_d,de2):.2e})')
