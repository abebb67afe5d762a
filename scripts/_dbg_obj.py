import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from test_gpu_surface import _hoi_setup
from followmyhold_b200 import _lib
loop, samples, st, raw_faces, fovs, tg_hoi, _ = _hoi_setup(B=2, D=17, seed=60)
ln=loop.lanes[0]; o=loop._obj
rf=ln.render_faces.clone()
print("F1",o.F1,"V1",o.V1,"cap",o.cap_v,o.cap_f, "rf max", int(rf.max()))
s=torch.cuda.current_stream()
w=loop.phase_weights(2)
for phase in (1.5, 2):
    ww=loop.phase_weights(phase)
    try:
        loop._object_eval(ln, phase, False, ww, s)
        torch.cuda.synchronize()
    except Exception as e:
        print("phase",phase,"EXC",e)
    print("phase",phase,"faces1 intact", torch.equal(o.joint_faces[:o.F1], rf), "vo", o.ex.vert_offsets.tolist(), "fo", o.ex.face_offsets.tolist(), "flags", int(o.ex.flags))
    jf=o.joint_faces[o.F1:o.F1+int(o.ex.face_offsets[-1])]
    print("  set2 faces min/max", int(jf.min()) if jf.numel() else None, int(jf.max()) if jf.numel() else None, "allowed", o.V1, o.V1+int(o.ex.vert_offsets[-1]))
print("---- schedule with decoder volume")
from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder, lattice_points
from test_gpu_decoder import _vae
from followmyhold_b200.guidance.loop import set_timesteps_sigmas
cfg=loop.cfg
cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 2, 2, 2
cfg.with_steps(6); loop.sigmas=set_timesteps_sigmas(6); loop.nan_steps=torch.zeros(6,2,dtype=torch.int32,device="cuda")
vae=_vae(1,seed=21)
with torch.no_grad(): vae.geo_decoder.output_proj.weight.mul_(3.0)
dec=LatentDecoder(DecoderWeights(vae.state_dict(),"cuda:0"),2,query_chunk=2048,active_chunk=512); dec.set_queries(lattice_points(17))
g=torch.Generator().manual_seed(2)
loop.x_t.copy_(torch.randn(2,loop.L,generator=g)); vel=(0.5*torch.randn(2,loop.L,generator=g)).cuda()
orig=loop._object_eval
def wrapped(ln_, phase, late, w_, s_):
    o.ex.extract(loop.sdf, stream=s_); torch.cuda.synchronize()
    print("phase",phase,"sdf min/max",float(loop.sdf.min()),float(loop.sdf.max()),"neg frac",float((loop.sdf<0).float().mean()),"vo",o.ex.vert_offsets.tolist(),"fo",o.ex.face_offsets.tolist(),"flags",int(o.ex.flags), flush=True)
    jf=o.joint_faces[o.F1:]
    print("   tail faces min/max", int(jf.min()), int(jf.max()), "cap", o.V1+o.cap_v, flush=True)
    return orig(ln_, phase, late, w_, s_)
loop._object_eval=wrapped
try:
    loop.run_schedule_tc_decoder(lambda i,x: vel/(1.0+i), dec, last_step=4)
    torch.cuda.synchronize(); print("OK")
except Exception as e:
    print("EXC", str(e)[:200])
