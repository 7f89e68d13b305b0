#!/usr/bin/env python3
"""Shared-memory wavefront count of the 12 kHz monitor FFT (monitor.cu: kiss_fft stage order on float2 elements in 16 banks of 8 bytes, a warp
served as two half-warps) for candidate paddings i + i // d and, per stage, k-fastest or g-fastest thread order.  Pure Python, no GPU: this is how
StaticPlan<1920>::pad and ::gfast were chosen.  usage: tools/fft_bank_sim.py"""
import numpy as np, itertools, sys
def kf_perm(n, radix_rec, rem):
    nf=len(radix_rec); stride=[1]*nf; acc=1
    for l in range(nf): stride[l]=acc; acc*=radix_rec[l]
    perm=np.zeros(n,int)
    for o in range(n):
        r=o; src=0
        for l in range(nf):
            d=r//rem[l]; r-=d*rem[l]; src+=d*stride[l]
        perm[o]=src
    inv=np.zeros(n,int); inv[perm]=np.arange(n)
    return inv
def wavefronts(addrs):
    """addrs: list of 32 (or fewer) 8-byte-unit addresses (None = inactive). two half-warps; per half: max multiplicity of distinct addresses per bank (16 banks of 8 B)"""
    tot=0
    for h in range(0,len(addrs),16):
        half=[a for a in addrs[h:h+16] if a is not None]
        if not half: continue
        banks={}
        for a in set(half): banks.setdefault(a%16,0); banks[a%16]+=1
        tot+=max(banks.values())
    return tot
def sim(n, exec_radix, exec_m, inv, phi, T=256):
    res={}
    # load: src=t.. store z[phi(inv[src])]
    w=0; ideal=0
    for base in range(0,n,32):
        a=[phi(inv[s]) for s in range(base,min(base+32,n))]
        w+=wavefronts(a); ideal+=2 if len(a)>16 else 1
    res['load']=(w,ideal)
    for s,(p,m) in enumerate(zip(exec_radix,exec_m)):
        nbf=n//p; w=0; ideal=0
        for b0 in range(0,nbf,32):
            bs=range(b0,min(b0+32,nbf))
            for q in range(p):
                a=[phi((b//m)*p*m+(b%m)+q*m) for b in bs]
                ww=wavefronts(a); w+=2*ww; ideal+=2*(2 if len(a)>16 else 1)  # load+store
        res['stage%d(p=%d,m=%d)'%(s,p,m)]=(w,ideal)
    # split: z[k], z[n-k]
    w=0; ideal=0
    for k0 in range(0,n//2+1,32):
        ks=range(k0,min(k0+32,n//2+1))
        a=[phi(k) for k in ks]; b=[phi((n-k)%n) for k in ks]
        w+=wavefronts(a)+wavefronts(b); ideal+=2*(2 if len(a)>16 else 1)
    res['split']=(w,ideal)
    return res
n=1920
# recursion order radices: 4,4,4,2,3,5 ; rem: 480,120,30,15,5,1
rec=[4,4,4,2,3,5]; rem=[480,120,30,15,5,1]
inv=kf_perm(n,rec,rem)
ex_r=[5,3,2,4,4,4]; ex_m=[1,5,15,30,120,480]
for name,phi in (("identity",lambda i:i),("i+i//16",lambda i:i+i//16),("i+i//32",lambda i:i+i//32),("i+i//15",lambda i:i+i//15),("i+i//30",lambda i:i+i//30),("i+i//5",lambda i:i+i//5), ("i+i//120", lambda i:i+i//120)):
    r=sim(n,ex_r,ex_m,inv,phi)
    tw=sum(v[0] for v in r.values()); ti=sum(v[1] for v in r.values())
    print(name, tw, ti, {k:v for k,v in r.items()})

print("---- search")
def stage_cost(n,p,m,phi,mode):
    nbf=n//p; G=nbf//m; w=0
    for b0 in range(0,nbf,32):
        bs=range(b0,min(b0+32,nbf))
        for q in range(p):
            if mode=='k': a=[phi((b//m)*p*m+(b%m)+q*m) for b in bs]
            else: a=[phi((b%G)*p*m+(b//G)+q*m) for b in bs]
            w+=2*wavefronts(a)
    return w
def load_cost(n,inv,phi):
    w=0
    for base in range(0,n,32):
        w+=wavefronts([phi(inv[s]) for s in range(base,min(base+32,n))])
    return w
def split_cost(n,phi):
    w=0
    for k0 in range(0,n//2+1,32):
        ks=range(k0,min(k0+32,n//2+1))
        w+=wavefronts([phi(k) for k in ks])+wavefronts([phi((n-k)%n) for k in ks])
    return w
pads={"id":lambda i:i}
for d in (8,15,16,24,30,32,40,48,60,64,96,120,128,240,480):
    pads["i+i//%d"%d]=(lambda d:(lambda i:i+i//d))(d)
best=[]
for name,phi in pads.items():
    tot=load_cost(n,inv,phi)+split_cost(n,phi); choice=[]
    for p,m in zip(ex_r,ex_m):
        ck=stage_cost(n,p,m,phi,'k'); cg=stage_cost(n,p,m,phi,'g')
        choice.append(('k',ck) if ck<=cg else ('g',cg)); tot+=min(ck,cg)
    best.append((tot,name,load_cost(n,inv,phi),split_cost(n,phi),choice))
for b in sorted(best)[:6]: print(b)
# FT4: n=576: rec 4,4,4,3,3 rem 144,36,9,3,1
n2=576; inv2=kf_perm(n2,[4,4,4,3,3],[144,36,9,3,1]); r2=[3,3,4,4,4]; m2=[1,3,9,36,144]
best=[]
for name,phi in pads.items():
    tot=load_cost(n2,inv2,phi)+split_cost(n2,phi); choice=[]
    for p,m in zip(r2,m2):
        ck=stage_cost(n2,p,m,phi,'k'); cg=stage_cost(n2,p,m,phi,'g')
        choice.append(('k',ck) if ck<=cg else ('g',cg)); tot+=min(ck,cg)
    best.append((tot,name,load_cost(n2,inv2,phi),split_cost(n2,phi),choice))
print("FT4")
for b in sorted(best)[:5]: print(b)
print([b for b in best if b[1]=="id"])
