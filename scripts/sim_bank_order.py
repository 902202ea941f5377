"""Dev tool (CPU only): L1 data-stage model of the 16-byte neighbour gathers (scripts/ubench/gather_banks.cu: a warp-gather costs
max over the 8 chunk offsets (j mod 8) of the number of DISTINCT addresses with that offset, >= ceil(active lanes / 8)),
evaluated on an evolved state for the list orders:
   ascending   the frozen lists in ascending index order (shipped until round 2)
   rr8         round-robin over the 8 classes (j - i) mod 8, ascending inside a class, empty classes skipped
Only index DIFFERENCES enter the order, so it is the same on a slab and on one GPU."""
import numpy as np, sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import helpers as H
from scipy.spatial import cKDTree
nx,ny,nz=48,48,48
steps=int(sys.argv[1]) if len(sys.argv)>1 else 6
pos,vel=H.lattice_block(nx,ny,nz,origin=(0.1,0.1,0.1),spacing=0.1,v0=(0,-1,0),jitter=0.001,seed=1234)
prm=H.default_params(rest_density=700.0,box_min=(0,0,0),box_max=(12.0,8.0,4.9),y_light=8.0,z_front=4.9,xsph_mode=H.XSPH_JACOBI)
o=H.Oracle(prm,32,H.COLLIDE_BOX,H.SEARCH_GRID); o.upload(pos,vel); o.step(steps)
P,V,R=o.download()
h=0.3; cell=float(np.float32(h)*(1+1/256))
tree=cKDTree(P); nb0=tree.query_ball_point(P,h)
n=len(P); ids=np.arange(n)
cx=np.floor(P[:,0]/cell).astype(int); cy=np.floor(P[:,1]/cell).astype(int); zt=np.floor(P[:,2]/(cell/8)).astype(int)
order=np.lexsort((ids,zt,cy,cx)); inv=np.empty(n,dtype=np.int64); inv[order]=np.arange(n)
def rr(i,js,M):
    b=[list(js[(js-i)%M==c]) for c in range(M)]
    out=[]; k=0
    while len(out)<len(js):
        for c in range(M):
            if k<len(b[c]): out.append(b[c][k])
        k+=1
    return np.array(out,dtype=np.int64)
def rr_keep(i,js,M):      # rounds stay 8 slots long: an empty class borrows from the fullest one
    b=[list(js[(js-i)%M==c]) for c in range(M)]
    out=[]
    while len(out)<len(js):
        for c in range(M):
            if len(out)==len(js): break
            src=c if b[c] else max(range(M),key=lambda q:(len(b[q]),-q))
            out.append(b[src].pop(0))
    return np.array(out,dtype=np.int64)
def group_greedy(i,js,M,G=8):   # groups of G consecutive neighbours; inside a group an element goes to the slot of its class if that slot is free
    out=[]
    for g0 in range(0,len(js),G):
        grp=list(js[g0:g0+G]); L=len(grp)
        slots=[None]*G; rest=[]
        for j in grp:
            c=int((j-i)%M)%G
            if c<L and slots[c] is None: slots[c]=j
            else: rest.append(j)
        for q in range(L):
            if slots[q] is None: slots[q]=rest.pop(0)
        out.extend(slots[:L])
    return np.array(out,dtype=np.int64)
rng=np.random.default_rng(0); nw=n//32
ws=rng.choice(np.arange(nw//4,3*nw//4),size=120,replace=False)
res={}
for name,fn in (("ascending",lambda i,js:js),("rr8",lambda i,js:rr(i,js,8)),("rr8 keep rounds",lambda i,js:rr_keep(i,js,8)),("rr4 (32-byte records)",lambda i,js:rr(i,js,4)),("groups of 8, greedy",lambda i,js:group_greedy(i,js,8,8)),("groups of 16, greedy",lambda i,js:group_greedy(i,js,8,16))):
    conf=0; conf32=0; lines=0; g=0; floor_=0
    for w in ws:
        L=[]
        for l in range(32):
            i=w*32+l
            js=np.sort(inv[[j for j in nb0[order[i]] if j!=order[i]]])
            L.append(fn(i,js))
        m=max(len(x) for x in L)
        for s in range(m):
            a=np.unique([L[l][s] for l in range(32) if s<len(L[l])])
            conf+=np.bincount(a%8,minlength=8).max(); conf32+=np.bincount(a%4,minlength=4).max(); lines+=len(np.unique(a>>3)); g+=1
            floor_+=-(-len(a)//8)
    print(f"{name:24s} 16-B gathers: bank wavefronts {conf/g:5.2f} (floor {floor_/g:4.2f})   32-B records: {conf32/g:5.2f}   lines {lines/g:5.2f}   gathers/warp {g/len(ws):6.1f}")
