"""Long-running fuzz of the HOST build of the exact-skip marcher (rt_march_candidates_host) against the oracle's literal
loop: random surfaces / parameters / transforms / steps / depths, primary-like rays and secondary rays starting on the
surface.  usage: python tools/fuzz_march_host.py <seed> <configurations>   (tests/test_march_host.py runs a short version)"""
import sys, json, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import rs_pathtracing_b200 as rt
from oracle import pyoracle as po
from test_march_host import host_march, CAMERA, GREY
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad=0; total=0
for trial in range(int(sys.argv[2]) if len(sys.argv)>2 else 60):
    kind=rng.choice(["Heart","Sine","Star","DupinCyclide","HuntsSurface","Cushion"])
    if kind=="Heart": surf={"type":"Heart"}; R0=1.5
    elif kind=="Sine": r=rng.uniform(0.8,3.0); surf={"type":"Sine","a":rng.uniform(0.2,1.5),"sphere_radius":r}; R0=r
    elif kind=="Star": r=rng.uniform(0.8,3.0); surf={"type":"Star","a":rng.uniform(0.3,3.0),"sphere_radius":r}; R0=r
    elif kind=="DupinCyclide": r=rng.uniform(1.5,3.5); surf={"type":"DupinCyclide","a":rng.uniform(0.8,1.5),"b":rng.uniform(0.5,1.2),"c":rng.uniform(0.1,0.8),"d":rng.uniform(0.05,0.5),"sphere_radius":r}; R0=r
    elif kind=="HuntsSurface": r=rng.uniform(3.0,6.0); surf={"type":"HuntsSurface","sphere_radius":r}; R0=r
    else: r=rng.uniform(0.8,2.5); surf={"type":"Cushion","sphere_radius":r}; R0=r
    scale=[float(10**rng.uniform(-0.5,1.5))*float(rng.uniform(0.5,2.0)) for _ in range(3)] if rng.random()<0.5 else [float(10**rng.uniform(-0.5,2.0))]*3
    rot=[float(x) for x in rng.uniform(-180,180,3)]
    centre=rng.uniform(-5,5,3)
    step=float(10**rng.uniform(-3,-1.3)); depth=int(rng.integers(1,6))
    text=json.dumps({"camera":CAMERA,"background":[0,0,0],"materials":{"M":GREY},"shapes":[{"type":"BruteForsableShape","shape":surf,"step":step,"depth":depth,"material":"M","transform":{"translate":list(map(float,centre)),"rotate":rot,"scale":scale}}]})
    sc=rt.Scene.from_json(text,add_random_spheres=False)
    R=R0*max(scale)
    n=300
    def ball(m,r):
        v=rng.normal(size=(m,3)); v/=np.linalg.norm(v,axis=1,keepdims=True); return v*r*rng.uniform(0,1,(m,1))**(1/3)
    o=centre+ball(n,4*R); tgt=centre+ball(n,0.9*R)
    rays=rt.make_rays(o,tgt-o)
    # limit oracle cost: chord steps = 2R/step
    if 2*R/step > 3e5: continue
    want=po.OracleScene(sc.desc()).intersect_batch(rays)
    hit,t,ev=host_march(sc,rays)
    wh=want["index"]==0
    mism=(hit!=wh).sum()+ (t[hit&wh].view(np.uint64)!=want["t"][hit&wh].view(np.uint64)).sum()
    total+=n
    # secondary rays: from the hit points (on the surface, to rounding), in random directions; and nudged copies
    pts=want["point"][wh]
    if len(pts):
        dirs=rng.normal(size=(len(pts),3)); dirs/=np.linalg.norm(dirs,axis=1,keepdims=True)
        rays2=np.concatenate([pts,dirs],axis=1)
        nudged=np.concatenate([pts+dirs*rng.uniform(-1e-9,1e-9,(len(pts),1))*R,dirs],axis=1)
        rays2=np.ascontiguousarray(np.concatenate([rays2,nudged]))
        want2=po.OracleScene(sc.desc()).intersect_batch(rays2)
        hit2,t2,_=host_march(sc,rays2)
        wh2=want2["index"]==0
        m2=(hit2!=wh2).sum()+(t2[hit2&wh2].view(np.uint64)!=want2["t"][hit2&wh2].view(np.uint64)).sum()
        total+=len(rays2)
        if m2:
            bad+=1
            print("MISMATCH secondary",trial,kind,surf,scale,rot,step,depth,"rays",int(m2),"of",len(rays2))
    if mism:
        bad+=1
        print("MISMATCH",trial,kind,surf,scale,rot,step,depth,"rays",int(mism),"hits",int(wh.sum()))
print("trials done, rays",total,"bad configs",bad)
