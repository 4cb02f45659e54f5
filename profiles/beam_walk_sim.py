"""CPU model of the beam walk of the hierarchy kernels (csrc/rt3_kernels.cuh, beam_for_chunk_bvh) on BASELINE C5 (10^6 random spheres), written
BEFORE the kernel to size its two capacities and to check the inequality: for random chunks of PIX adjacent pixels it builds the chunk's beam,
walks the tree level by level with the kernel's box test (bounding sphere of the widened box against the beam), and checks that the candidate
list holds the closest hit of every ray of the chunk (pixel corners and centres, brute force over all spheres). Also models an occluder filter
on the list (a sphere every ray of the beam must hit bounds how far the others can matter), which was not built.
Runs here, no GPU; tree as profiles/c5_visits_sim.py. Prints one JSON line: candidates / visits / level widths, how many chunks exceed 48 / 64.
Result (80 chunks of 4 pixels, profiles/r02_beam_walk_sim.jsonl): 56 candidates on average, 152 at most; 785 node records per walk; levels of 103 nodes
on average, 153 at most; with the occluder filter 44 candidates; no hit missed. Hence lists of 192 and levels of 192 in the kernel (the GPU counts 58.9).
Usage: python profiles/beam_walk_sim.py [n_chunks] [pixels per chunk]"""
import json, math, os, sys
import numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rt3_b200
from rt3_b200 import scenes
f32=np.float32
n_chunks=int(sys.argv[1]) if len(sys.argv)>1 else 100
PIX=int(sys.argv[2]) if len(sys.argv)>2 else 4
scene, cam = scenes.random_spheres(1000000)
sp = scene.spheres.astype(np.float64)
c, r = sp[:, :3], sp[:, 3]
N=len(sp)
def spread21(v):
    x = v & 0x1fffff
    x = (x | x << 32) & 0x1f00000000ffff
    x = (x | x << 16) & 0x1f0000ff0000ff
    x = (x | x << 8) & 0x100f00f00f00f00f
    x = (x | x << 4) & 0x10c30c30c30c30c3
    x = (x | x << 2) & 0x1249249249249249
    return x
mn = c.min(0)
q = np.minimum(((c - mn) / (c.max(0) - mn) * 2097152).astype(np.uint64), 2097151)
key = spread21(q[:, 0]) | (spread21(q[:, 1]) << 1) | (spread21(q[:, 2]) << 2)
order = np.argsort(key, kind="stable")
keys, cs, rs = key[order], c[order], r[order]
u = 2.0 ** -24
R = np.sqrt(rs*rs + 64*u*((cs*cs).sum(1)+rs*rs))
lo, hi = cs - R[:,None], cs + R[:,None]
splits={}; cache={}
def split(a,b):
    if (a,b) not in splits:
        f,l=int(keys[a]),int(keys[b-1])
        if f==l: splits[(a,b)]=(a+b)//2
        else:
            bit=(f^l).bit_length()-1
            splits[(a,b)]=a+int(np.searchsorted(keys[a:b], np.uint64(((f>>bit)|1)<<bit),"left"))
    return splits[(a,b)]
def box(a,b):
    if (a,b) not in cache: cache[(a,b)]=(lo[a:b].min(0), hi[a:b].max(0))
    return cache[(a,b)]
W,H=1920,1080
hor,ver,llc,org=(np.array(list(v)[:3],float) for v in (cam.horizontal,cam.vertical,cam.lower_left_corner,cam.origin))
lens_radius=float(cam.lens_radius); lens_u=np.array(list(cam.lens_u)[:3],float); lens_v=np.array(list(cam.lens_v)[:3],float)
print("origin",org,"lens",lens_radius, file=sys.stderr)
ray_margin = 2.0**-9
def beam_geom(xa,xb,y):
    iw,ih=1/(W-1),1/(H-1)
    u0,u1=xa*iw,(xb+1)*iw; v0,v1=(H-1-y)*ih,(H-1-y+1)*ih
    um,vm=0.5*(u0+u1),0.5*(v0+v1)
    D=llc+um*hor+vm*ver-org
    len_hor,len_ver,len_o=np.linalg.norm(hor),np.linalg.norm(ver),np.linalg.norm(org)
    tiny=2.0**-21*(len_hor+len_ver+np.linalg.norm(llc)+len_o)
    delta=(0.5*(u1-u0)*len_hor+0.5*(v1-v0)*len_ver)*1.001+tiny
    lens=lens_radius*(np.linalg.norm(lens_u)+np.linalg.norm(lens_v))*1.001+tiny if lens_radius>0 else tiny
    k=lens+delta; dd=D@D; len_d=math.sqrt(dd)
    return dict(D=D,k=k,lens=lens,len_d=len_d,inv_dd=1/dd,inv_reach=1/(len_d-k),o_max=len_o+lens,tiny=tiny)
def meets(g,blo,bhi,grow):
    ctr=0.5*blo+0.5*bhi; half=0.5*bhi-0.5*blo; rho=math.sqrt(half@half)
    re=(rho+grow)*1.0001+2.0**-21*(np.abs(ctr).sum()+rho)
    co=ctr-org; proj=co@g['D']; co2=co@co; s_star=proj*g['inv_dd']; dist2=co2-s_star*proj
    s_hi=(max(s_star,0)*g['len_d']+re+g['lens'])*g['inv_reach']; reach=(re+g['lens']+s_hi*g['k'])*1.0001
    return not (dist2>reach*reach+2.0**-19*co2)
def walk(g):
    grow=1.7320508*ray_margin*g['o_max']*1.0001+g['tiny']
    level=[(0,N)]; cands=[]; visits=0; maxw=1
    while level:
        nxt=[]
        for (a,b) in level:
            visits+=1; m=split(a,b)
            for (x,y2) in ((a,m),(m,b)):
                bl,bh=box(x,y2)
                if meets(g,bl,bh,grow):
                    if y2-x==1: cands.append(x)
                    else: nxt.append((x,y2))
        level=nxt; maxw=max(maxw,len(level))
    return cands,visits,maxw
SLACK=64*2.0**-24
def occl_filter(g,cands):
    tb=math.inf
    info=[]
    for j in cands:
        cc=cs[j]; rr=rs[j]; co=cc-org; co2=co@co; lco=math.sqrt(co2)
        proj=co@g['D']; s_star=proj*g['inv_dd']; dist=math.sqrt(max(co2-s_star*proj,0.0))
        E=SLACK*(cc@cc+rr*rr+g['o_max']**2); sE=math.sqrt(E)
        cover=(g['lens']+max(s_star,0)*g['k']+dist)*1.001
        is_occ = s_star>0 and cover<=0.9*rr and 0.19*rr*rr>2*E and rr>0.01 and lco-g['lens']-rr>0.01
        if is_occ: tb=min(tb,(lco+g['lens']+rr+sE)*1.0001)
        info.append((j,(lco-g['lens']-(rr+sE)*1.001)))
    return [j for j,tmin in info if not (tmin>tb)]
def brute(o,d):
    oc=o-cs; h=oc@d; cc=(oc*oc).sum(1)-rs*rs; disc=h*h-cc
    ok=disc>=0; sq=np.sqrt(np.where(ok,disc,1)); t1=-h-sq; t2=-h+sq
    t=np.where(t1>=0.001,t1,t2); ok&=t>=0.001
    if not ok.any(): return -1
    t=np.where(ok,t,np.inf); return int(np.argmin(t))
rng=np.random.default_rng(1)
stats=[]; miss=0
for _ in range(n_chunks):
    xa=int(rng.integers(0,W-PIX)); y=int(rng.integers(0,H)); xb=xa+PIX-1
    g=beam_geom(xa,xb,y)
    cands,visits,maxw=walk(g)
    f=occl_filter(g,cands)
    stats.append((len(cands),visits,maxw,len(f)))
    cs_set=set(f)
    for x in range(xa,xb+1):
        for (jx,jy) in ((0,0),(0.999,0.999),(0.5,0.5),(0,0.999),(0.999,0)):
            uu=(x+jx)/(W-1); vv=(H-1-y+jy)/(H-1)
            d=llc+uu*hor+vv*ver-org; d/=np.linalg.norm(d)
            hit=brute(org,d)
            if hit>=0 and hit not in cs_set: miss+=1; print("MISS",xa,y,x,jx,jy,hit,file=sys.stderr)
st=np.array(stats)
print(json.dumps({"chunks":n_chunks,"pixels":PIX,"cands_mean":st[:,0].mean(),"cands_max":int(st[:,0].max()),"cands_gt48":int((st[:,0]>48).sum()),
 "visits_mean":st[:,1].mean(),"visits_max":int(st[:,1].max()),"width_mean":st[:,2].mean(),"width_max":int(st[:,2].max()),"width_gt64":int((st[:,2]>64).sum()),"final_mean":st[:,3].mean(),"final_max":int(st[:,3].max()),"final_gt48":int((st[:,3]>48).sum()),"final_hist":np.percentile(st[:,3],[10,50,90]).tolist(),"missed":miss}))
