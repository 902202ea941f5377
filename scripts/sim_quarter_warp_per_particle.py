import numpy as np, sys
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
cx=np.floor(P[:,0]/cell).astype(int); cy=np.floor(P[:,1]/cell).astype(int); zt=np.floor(P[:,2]/(cell/8)).astype(int)
order=np.lexsort((ids,zt,cy,cx)); inv=np.empty(n,dtype=np.int64); inv[order]=np.arange(n)
rng=np.random.default_rng(0); nw=n//32
ws=rng.choice(np.arange(nw//4,3*nw//4),size=120,replace=False)
cur_lines=0; cur_steps=0; q_lines=0; q_steps=0; q_slots=0; tot=0
for w in ws:
    L=[np.sort(inv[[j for j in nb0[order[w*32+l]] if j!=order[w*32+l]]]) for l in range(32)]
    tot+=sum(len(x) for x in L)
    m=max(len(x) for x in L); cur_steps+=m
    for s in range(m):
        for q in range(4):
            js=[L[l][s] for l in range(q*8,q*8+8) if s<len(L[l])]
            if js: cur_lines+=len(set(j>>3 for j in js))
    # quarter-warp per particle: a warp handles 4 particles at a time, 8 groups per 32 particles
    for g in range(8):
        Ls=L[4*g:4*g+4]
        mm=max((len(x)+7)//8 for x in Ls); q_steps+=mm
        for s in range(mm):
            for x in Ls:
                js=x[8*s:8*s+8]
                if len(js): q_lines+=len(set(j>>3 for j in js))
                q_slots+=8
print(f"neighbours per 32 particles {tot/len(ws):.0f}")
print(f"lane per particle : warp-steps per 32 particles {cur_steps/len(ws):.1f}, line-wavefronts {cur_lines/len(ws):.0f}")
print(f"8 lanes / particle: warp-steps per 32 particles {q_steps/len(ws):.1f}, line-wavefronts {q_lines/len(ws):.0f}, lane slots {q_slots/len(ws):.0f}")
