"""Shared-memory wavefront model of csrc/fft.cu::fft_tile_kernel (CPU only, numpy).

Counts the wavefronts of every phase of one tile -- first-sweep stores, the in-place sweeps, their twiddle reads, the
store phase -- for a plan (n, radix order), a tile layout ("pad": pitch W + 1 as shipped, "xor": the swizzled layout of
the round-2 experiment, "dense") and a lane mapping ("old": butterflies fastest after the first sweep, "new": columns
fastest everywhere).  64-bit accesses are modelled per half-warp: the wavefronts of a request are the largest number
of distinct 8-byte words that fall on one of the 16 bank pairs.  Validation: n = 900 as shipped gives 5572 wavefronts
per tile against 5762 measured by ncu (profiles/r01_ncu_fft.txt); the swizzled layout predicted 3605 and measured
4046 (profiles/r02_fft_experiments.txt).

    python scripts/sim_fft_banks.py
"""
import numpy as np
NT=384
def wavefronts(idx, active=None):
    """idx: (nwarps,32) float2 indices; returns total wavefronts (64-bit accesses, half-warp granularity)"""
    idx=np.asarray(idx)
    if active is None: active=np.ones_like(idx,bool)
    tot=0
    for h in range(2):
        a=idx[:,16*h:16*h+16]; m=active[:,16*h:16*h+16]
        for row,mk in zip(a,m):
            v=np.unique(row[mk])
            if v.size==0: continue
            tot+=np.bincount(v%16).max()
    return tot
def warps(items):
    """pad item array to multiple of NT iterations -> list of (nwarps,32) arrays + active masks per loop iteration"""
    n=len(items)
    out=[]
    for k0 in range(0,n,NT):
        w=np.arange(k0,k0+NT); act=w<n
        out.append((w.reshape(-1,32),act.reshape(-1,32)))
    return out
def pairs(radices, allowed):
    res=[];s=0
    while s<len(radices):
        ra=radices[s]; rb=1
        if s+1<len(radices) and (ra,radices[s+1]) in allowed: rb=radices[s+1]
        res.append((ra,rb)); s+=2 if rb>1 else 1
    return res
OLD_ALLOWED={(4,4),(4,2),(4,3),(4,5),(2,3),(2,5),(3,3),(3,5),(5,5)}
def sim(n, radices, W=8, layout='pad', mapping='old', transposed=True, keep=None, allowed=OLD_ALLOWED, packed_tw=False, verbose=True):
    WP = W+1 if layout=='pad' else W
    def phys(row,c):
        if layout=='pad': return row*WP+c
        if layout=='dense': return row*W+c
        if layout=='xor':
            rpl=16//W
            return row*W + (c ^ ((row//rpl)%W))
    twbase = ((n*WP+15)//16)*16
    ps=pairs(radices,allowed)
    rep={}
    # first sweep writes
    ra,rb=ps[0]; R=ra*rb; nsb=n//R
    tot=0
    for w,act in warps(np.arange(nsb*W)):
        c=w%W; g=w//W
        for k in range(R):
            tot+=wavefronts(phys(g*R+k,c),act)
    rep['s0_write']=tot
    L=R
    for s,(ra,rb) in enumerate(ps[1:],1):
        lp=L; Lt=lp*ra; R=ra*rb; nsb=n//R
        tws_a=n//(lp*ra); tws_b=n//(lp*ra*rb)
        col_fast = (lp==1 or nsb==1) if mapping=='old' else True
        tot=0; ttw=0
        for w,act in warps(np.arange(nsb*W)):
            if col_fast: c=w%W; b=w//W
            else: c=w//nsb; b=w-c*nsb
            g=b//lp; j=b-g*lp
            base=g*Lt*rb+j
            for qb in range(rb):
                for qa in range(ra):
                    tot+=2*wavefronts(phys(base+qa*lp+qb*Lt,c),act)
            for qa in range(1,ra):
                if packed_tw: ti=twbase+(qa-1)*lp+j
                else: ti=twbase+j*qa*tws_a
                ttw+=wavefronts(ti,act&(j!=0))
            if rb>1:
                for pa in range(ra):
                    jp=j+pa*lp
                    for qb in range(1,rb):
                        if packed_tw: ti=twbase+(ra-1)*lp+(qb-1)*Lt+jp
                        else: ti=twbase+jp*qb*tws_b
                        ttw+=wavefronts(ti,act&(jp!=0))
        rep[f's{s}_tile(lp={lp},R={ra}x{rb})']=tot; rep[f's{s}_tw']=ttw
        L=Lt*rb
    tot=0
    if transposed:
        for w,act in warps(np.arange(n)):
            for c in range(W): tot+=wavefronts(phys(w,c),act)
    else:
        for w,act in warps(np.arange(n*W)):
            q=w//W; c=w%W
            a=act if keep is None else act&((q<=keep[0])|(q>=keep[1]))
            tot+=wavefronts(phys(q,c),a)
    rep['store']=tot
    rep['total']=sum(rep.values())
    rep['ideal_tile']=n*W*8*(2*len(ps))//128
    if verbose:
        print(n,radices,layout,mapping,ps)
        for k,v in rep.items(): print('   ',k,v)
    return rep
if __name__=='__main__':
    # current plans at C2
    sim(900,[4,3,3,5,5],transposed=True)
    sim(1000,[4,2,5,5,5],transposed=False,keep=(720000//2//900, (900000-720000//2)//900))
    sim(720,[4,4,3,3,5],transposed=True)
    sim(1000,[4,2,5,5,5],transposed=False)
    print("==== option A: xor layout, col-fast, odd radices first")
    ALL={(a,b) for a in (2,3,4,5) for b in (2,3,4,5) if a*b<=25}
    for lay in ('xor','dense'):
        sim(900,[3,3,5,5,4],layout=lay,mapping='new',allowed=ALL)
        sim(1000,[5,5,5,4,2],layout=lay,mapping='new',allowed=ALL,transposed=False)
        sim(720,[3,3,5,4,4],layout=lay,mapping='new',allowed=ALL)
    print("==== option B: packed twiddles, old mapping")
    sim(900,[4,3,3,5,5],packed_tw=True)
    sim(1000,[4,2,5,5,5],packed_tw=True,transposed=False)
    sim(720,[4,4,3,3,5],packed_tw=True)
