"""Independent numpy restatement of the SABC algorithm for the 1-D Gaussian config (C1): numpy RNG, np.interp ECDF, scipy
brentq for eps.  Used only statistically (tests/test_oracle.py) as a second opinion on the C oracle; follows
src/SimulatedAnnealingABC.jl:151-227,251-402 and src/proposals.jl:101-114,137-148."""
import numpy as np
from scipy.optimize import brentq
def run(N=10000, nsim=2_000_000, v=1.0, delta=0.1, seed=0, prop="de"):
    rng=np.random.default_rng(seed)
    yobs=1.0; sd=1/np.sqrt(10)
    f=lambda th: np.abs(th+sd*rng.standard_normal(th.shape)-yobs)
    lp=lambda th: -0.5*th**2
    th=rng.standard_normal(N); rho=f(th)
    K=np.concatenate([[0],np.sort(rho[rho>0]),[rho.max()*1.5]]); y=np.linspace(0,1,len(K))
    G=lambda r: np.interp(r,K,y)
    u=G(rho)
    def resample(th,u):
        w=np.exp(-u*delta/u.mean()); idx=rng.choice(N,N,p=w/w.sum()); return th[idx],u[idx]
    th,u=resample(th,u)
    epsf=lambda ub: brentq(lambda e: e*e+v*e**1.5-ub*ub,0,ub)
    eps=epsf(u.mean()); nacc=0; nres=1
    h=N//2
    for it in range(nsim//N-1):
        for (a,b) in ((slice(0,h),slice(h,N)),(slice(h,N),slice(0,h))):
            A=th[a]; P=th[b]; n=len(A); M=len(P)
            if prop=="de":
                i1=rng.integers(0,M,n); i2=rng.integers(0,M-1,n); i2=i2+(i2>=i1)
                g=2.38/np.sqrt(2)*(1+1e-5*rng.standard_normal(n)); tp=A+g*(P[i1]-P[i2])
            else:
                i=rng.integers(0,M,n); z=((2-1)*rng.random(n)+1)**2/2; tp=P[i]+z*(A-P[i])
            rp=f(tp); up=G(rp)
            L=lp(tp)-lp(A)+(u[a]-up)/eps
            acc=np.log(rng.random(n))<L
            A=np.where(acc,tp,A); th[a]=A; ua=u[a]; ua=np.where(acc,up,ua); u[a]=ua; nacc+=acc.sum()
        if nacc>=(nres+1)*2*N: th,u=resample(th,u); nres+=1
        eps=epsf(u.mean())
    return th.mean(), th.var(), eps, u.mean(), nacc, nres
if __name__ == "__main__":
    for prop in ("de", "stretch"):
        for seed in (1, 2):
            print(prop, seed, run(seed=seed, prop=prop))
