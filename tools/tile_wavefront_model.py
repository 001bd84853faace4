"""Model of roi_tile_tma_kernel's shared-memory wavefronts for config-2-like RoIs (uniform centres, w, h ~ U(1.5, 40), 32 x 24
tiles): per (piece, bin column) unit 4 wavefronts per 16-byte chunk and row + 2 for the row weights (`cur`, 71.5 M x 8 channel
groups / 8 = what ncu counts: 82 M loads), against one pass per piece (`merged`), the hybrid for pieces spanning <= 2 / <= 3 chunks
(`hyb2`, `hyb3`) and the floor of one wavefront per pixel column and row (`min_wf`).  CPU only:  python tools/tile_wavefront_model.py"""
import numpy as np, math
rng=np.random.default_rng(1)
H,W=272,480
N=11324
cx=rng.uniform(0,W,N); cy=rng.uniform(0,H,N)
w=rng.uniform(1.5,40,N); h=rng.uniform(1.5,40,N)
x1=np.clip(cx-w/2,0,None); y1=np.clip(cy-h/2,0,None); x2=x1+w; y2=y1+h
def axis(a,b,size):
    ln=max(b-a,1.0); bn=ln/3; grid=int(math.ceil(ln/3))
    lo=[size]*3; hi=[-1]*3
    for p in range(3):
        for i in range(grid):
            v=a+p*bn+(i+.5)*bn/grid
            if v<-1 or v>size: continue
            v=max(v,0); l=int(v)
            if l>=size-1: l=hh=size-1
            else: hh=l+1
            lo[p]=min(lo[p],l); hi[p]=max(hi[p],hh)
    return lo,hi
TW,TH=32,24
tot=dict(cur=0,merged=0,hyb2=0,hyb3=0,rows=0,pieces=0,units=0,min_wf=0)
nq_hist={}
import collections
piece_span=collections.Counter()
for n in range(N):
    xl,xh=axis(x1[n],x2[n],W); yl,yh=axis(y1[n],y2[n],H)
    if min(xl)>max(xh) or min(yl)>max(yh): continue
    X0,X1=min(xl),max(xh); Y0,Y1=min(yl),max(yh)
    for ty in range(Y0//TH,Y1//TH+1):
        r0=max(Y0,ty*TH); r1=min(Y1,ty*TH+TH-1); nr=r1-r0+1
        for tx in range(X0//TW,X1//TW+1):
            px0=tx*TW
            tot['pieces']+=1
            qs=[]
            cur=0
            for p in range(3):
                c0=max(xl[p],px0); c1=min(xh[p],px0+TW-1)
                if c1<c0: 
                    tot['units']+=1; continue
                q0=(c0-px0)>>2; q1=(c1-px0)>>2
                qs.append((q0,q1))
                cur+=nr*(4*(q1-q0+1)+1)
                tot['units']+=1
                tot['min_wf']+=nr*(c1-c0+1)
            if not qs: continue
            span=max(q for _,q in qs)-min(q for q,_ in qs)+1
            piece_span[span]+=nr
            mer=nr*(4*span+1)
            tot['cur']+=cur; tot['merged']+=mer
            tot['hyb2']+=mer if span<=2 else cur
            tot['hyb3']+=mer if span<=3 else cur
            tot['rows']+=nr
print(tot)
print({k:v/1e6*8 for k,v in tot.items()})  # x8 channel groups, in M wavefronts
print(sorted(piece_span.items()))
