import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from helpers import *
from simpledepthestimation_b200.functional import MonoLossPlan
from simpledepthestimation_b200.synthetic import mono_inputs, euler_pose
dev='cuda:0'
for (B,H,W,kw) in [(2,32,64,{}),(1,192,640,{}),(2,48,80,dict(automask=False)),(2,48,80,dict(reduce='mean')),(3,50,70,{}),(1,24,80,dict(ssim_w=0.0))]:
    inp = mono_inputs(B,H,W)
    tgt,src = build_pyramid(inp)
    ref = oracle_mono(inp, torch.float64, tgt, src, **kw)
    sizes=[d.shape[-2:] for d in inp['depth']]
    plan = MonoLossPlan(B, sizes, 2, (H,W), dev, automask=kw.get('automask',True), reduce=kw.get('reduce','min'), ssim_weight=kw.get('ssim_w',0.85))
    g = lambda t: t.to(dev).contiguous()
    pose=[g(euler_pose(v)) for v in inp['pose_vec']]
    losses, argmin = plan.forward([g(t) for t in tgt], [[g(s) for s in ss] for ss in src], [g(d) for d in inp['depth']], g(inp['K']), pose)
    torch.cuda.synchronize()
    l = losses.cpu().double()
    print(B,H,W,kw,'rec', l[0].item(), ref['rec_loss'].item(), 'rel', abs(l[0].item()-ref['rec_loss'].item())/ref['rec_loss'].item(), 'smooth', l[1].item(), float(ref['smooth_loss']), 'rel', abs(l[1].item()-float(ref['smooth_loss']))/float(ref['smooth_loss']))
    if 'argmin' in ref and kw.get('reduce','min')=='min':
        for i,a in enumerate(argmin):
            mism = (a.cpu().long()!=ref['argmin'][i])
            c = ref['cand'][i]; top2 = c.topk(2,dim=1,largest=False)[0]; gap=(top2[:,1]-top2[:,0])
            print('   scale',i,'argmin mismatches',int(mism.sum()),'max gap at mismatches', float(gap[mism].max()) if mism.any() else 0.0)
    # determinism
    l2,_ = plan.forward([g(t) for t in tgt], [[g(s) for s in ss] for ss in src], [g(d) for d in inp['depth']], g(inp['K']), pose)
    print('   deterministic', torch.equal(l2, losses))
