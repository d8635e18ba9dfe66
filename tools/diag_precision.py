import os, sys, numpy as np, torch
sys.path[:0]=['/root/repo','/root/repo/rc-mvsnet_b200']
from rcmvs import synth, weights
import models
G='/root/repo/tests/golden'
g=dict(np.load(G+'/mvs_full_192_64_8.npz'))
dev=torch.device('cuda:0')
imgs, proj, dv = synth.mvs_inputs(5, 512, 640, kind="textured", seed=4)
cu=lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
for mode in ('fp32','tf32x3','tf32'):
    os.environ['RCMVS_CONV']=mode
    m=models.CascadeMVSNet_eval(ndepths=[192,64,8]); weights.load_model(m, G+'/cas_model_000014.npz'); m.to(dev).eval()
    try:
        with torch.no_grad(): out=m(cu(imgs), {k:cu(v) for k,v in proj.items()}, cu(dv))
        torch.cuda.synchronize()
    except Exception as e:
        print(mode,'FAILED',repr(e)[:300]); continue
    for k in ('stage1','stage2','stage3'):
        d=out[k]['depth'][0].cpu().numpy(); r=np.abs(d-g['depth_'+k])/np.abs(g['depth_'+k])
        print(mode,k,'relLinf %.3e  frac>1e-3 %.2e  frac>1e-4 %.2e  p99.9 %.2e median %.2e'%(r.max(),(r>1e-3).mean(),(r>1e-4).mean(),np.quantile(r,0.999),np.median(r)))
