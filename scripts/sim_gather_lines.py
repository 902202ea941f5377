"""Dev tool (CPU only): distinct 128-B lines / 32-B sectors per warp gather step for alternative particle orders, on an
evolved 110k-particle state from the fp32 oracle.  Used to pre-screen layout ideas before spending GPU time
(results in profiles/r01_cache_policy_ab.txt)."""
import numpy as np, sys, time
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import helpers as H
from scipy.spatial import cKDTree
nx,ny,nz=48,48,48
pos,vel=H.lattice_block(nx,ny,nz,origin=(0.1,0.1,0.1),spacing=0.1,v0=(0,-1,0),jitter=0.001,seed=1234)
prm=H.default_params(rest_density=700.0,box_min=(0,0,0),box_max=(12.0,8.0,4.9),y_light=8.0,z_front=4.9,xsph_mode=H.XSPH_JACOBI)
o=H.Oracle(prm,32,H.COLLIDE_BOX,H.SEARCH_GRID); o.upload(pos,vel); o.step(6)
P,V,R=o.download()
h=0.3; cell=float(np.float32(h)*(1+1/256))
tree=cKDTree(P); nb0=tree.query_ball_point(P,h)
n=len(P); ids=np.arange(n)
def evaluate(name,key_cols):
    order=np.lexsort(tuple(reversed(key_cols)))   # first key most significant
    inv=np.empty(n,dtype=np.int64); inv[order]=np.arange(n)
    # interior warps only
    tot=0; steps=0; sect=0
    rng=np.random.default_rng(0)
    nw=n//32
    ws=rng.choice(np.arange(nw//4,3*nw//4),size=80,replace=False)
    for w in ws:
        L=[]
        for l in range(32):
            i=order[w*32+l]
            js=np.sort(inv[[j for j in nb0[i] if j!=i]])
            L.append(js)
        m=max(len(x) for x in L)
        for s in range(m):
            for q in range(4):
                js=[L[l][s] for l in range(q*8,q*8+8) if s<len(L[l])]
                if js:
                    tot+=len(set(j>>3 for j in js)); sect+=len(set(j>>1 for j in js))
        steps+=m
    print(f"{name:34s} lines/warp-step {tot/steps:6.2f}  sectors/warp-step {sect/steps:6.2f}  steps/warp {steps/len(ws):6.1f}")
cx=np.floor(P[:,0]/cell).astype(int); cy=np.floor(P[:,1]/cell).astype(int)
def zt(k): return np.floor(P[:,2]/(cell/k)).astype(int)
evaluate("col, z/1, id",[cx,cy,zt(1),ids])
evaluate("col, z/8, id  (shipped)",[cx,cy,zt(8),ids])
evaluate("col, exact z",[cx,cy,P[:,2]])
hx=np.floor(P[:,0]/(cell/2)).astype(int)&1; hy=np.floor(P[:,1]/(cell/2)).astype(int)&1
evaluate("col, z/4, xy-quadrant, id",[cx,cy,zt(4),hx*2+hy,ids])
evaluate("col, z/2, xy-quadrant, z",[cx,cy,zt(2),hx*2+hy,P[:,2]])
evaluate("col, xy-quadrant, exact z",[cx,cy,hx*2+hy,P[:,2]])
evaluate("col, z/3, y-half, z",[cx,cy,zt(3),hy,P[:,2]])
