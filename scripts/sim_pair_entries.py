import numpy as np, sys, time
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import helpers as H
from scipy.spatial import cKDTree
nx,ny,nz=48,48,48
pos,vel=H.lattice_block(nx,ny,nz,origin=(0.1,0.1,0.1),spacing=0.1,v0=(0,-1,0),jitter=0.001,seed=1234)
prm=H.default_params(rest_density=700.0,box_min=(0,0,0),box_max=(12.0,8.0,4.9),y_light=8.0,z_front=4.9,xsph_mode=H.XSPH_JACOBI)
o=H.Oracle(prm,32,H.COLLIDE_BOX,H.SEARCH_GRID); o.upload(pos,vel); o.step(int(sys.argv[1]) if len(sys.argv)>1 else 6)
P,V,R=o.download()
h=0.3; cell=float(np.float32(h)*(1+1/256))
tree=cKDTree(P); nb0=tree.query_ball_point(P,h)
n=len(P); ids=np.arange(n)
cx=np.floor(P[:,0]/cell).astype(int); cy=np.floor(P[:,1]/cell).astype(int); zt=np.floor(P[:,2]/(cell/8)).astype(int)
order=np.lexsort((ids,zt,cy,cx)); inv=np.empty(n,dtype=np.int64); inv[order]=np.arange(n)
rng=np.random.default_rng(0); nw=n//32
ws=rng.choice(np.arange(nw//4,3*nw//4),size=120,replace=False)
tot_n=0; tot_e=0; steps1=0; steps2=0; w1=0; w2=0; w2g4=0
for w in ws:
    L=[]; E=[]
    for l in range(32):
        i=order[w*32+l]
        js=np.sort(inv[[j for j in nb0[i] if j!=i]]); L.append(js)
        E.append(np.unique(js>>1))
    tot_n+=sum(len(x) for x in L); tot_e+=sum(len(x) for x in E)
    m1=max(len(x) for x in L); m2=max(len(x) for x in E); steps1+=m1; steps2+=m2
    for s in range(m1):
        for q in range(4):
            js=[L[l][s] for l in range(q*8,q*8+8) if s<len(L[l])]
            if js: w1+=len(set(j>>3 for j in js))
    for s in range(m2):
        for q in range(4):
            ps=[E[l][s] for l in range(q*8,q*8+8) if s<len(E[l])]
            if ps: w2+=len(set(p>>2 for p in ps))
        for q in range(8):
            ps=[E[l][s] for l in range(q*4,q*4+4) if s<len(E[l])]
            if ps: w2g4+=len(set(p>>2 for p in ps))
print(f"neighbours/particle {tot_n/len(ws)/32:.1f}  pair entries/particle {tot_e/len(ws)/32:.1f}  ratio {tot_e/tot_n:.3f}")
print(f"steps/warp: single {steps1/len(ws):.1f}  pairs {steps2/len(ws):.1f}")
print(f"line-wavefronts/warp: single(16B, per 8 lanes) {w1/len(ws):.0f}   pairs per-8-lane lines {w2/len(ws):.0f}   pairs per-4-lane lines {w2g4/len(ws):.0f}")
