import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from test_gpu_surface import _volumes
from followmyhold_b200.guidance.surface import SurfaceExtractor
from oracle import surface_oracle as SO
D=17; sdf=_volumes(1,D,seed=D)
ex=SurfaceExtractor(1,D,index_base=7); ex.extract(sdf.cuda()); torch.cuda.synchronize()
v,f,e=ex.meshes()[0]
ov,of,oe=SO.extract(sdf[0].double())
es=torch.sort(e,1).values
oset=set(map(tuple,oe.tolist()))
bad=[(i,tuple(p)) for i,p in enumerate(es.tolist()) if tuple(p) not in oset]
print("unique", torch.unique(es,dim=0).shape, "oracle", oe.shape, "faces equal", torch.equal(f,of))
print("n",len(es),"bad",len(bad), bad[:10])
print("min/max", es.min().item(), es.max().item(), "neg", (es<0).sum().item())
# which list
nA = 568
print("bad in A:", sum(1 for i,_ in bad if i<nA), "bad in B:", sum(1 for i,_ in bad if i>=nA))
cov=ex.cube_of_vert[:v.shape[0]].cpu()
n=D-1
for i,p in bad[:6]:
    c0,c1=cov[p[0]].item(),cov[p[1]].item()
    print(i,p,[(c0//(n*n),(c0//n)%n,c0%n)],[(c1//(n*n),(c1//n)%n,c1%n)])
from collections import defaultdict
pos=defaultdict(list)
for i,p in enumerate(es.tolist()): pos[tuple(p)].append(i)
dups=[(p,ix) for p,ix in pos.items() if len(ix)>1]
print("dups",len(dups), dups[:8])
missing=[p for p in oset if p not in pos]
print("missing",len(missing), missing[:8])
for p,ix in dups[:4]:
    c0,c1=cov[p[0]].item(),cov[p[1]].item()
    print(p,ix,(c0//(n*n),(c0//n)%n,c0%n),(c1//(n*n),(c1//n)%n,c1%n))
for p in missing[:4]:
    c0,c1=cov[p[0]].item(),cov[p[1]].item()
    print("missing",p,(c0//(n*n),(c0//n)%n,c0%n),(c1//(n*n),(c1//n)%n,c1%n))
