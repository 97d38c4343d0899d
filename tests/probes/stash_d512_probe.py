import sys, math, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import closed_form
from synergy_clip_b200 import ops
def rel(a,b): return float(np.sqrt(((a-b)**2).sum()/(b**2).sum()))
for (b,d,t,dt) in [(300,512,math.log(100.0),'f32'),(300,512,math.log(100.0),'bf16'),(300,512,math.log(50.0),'f32'),(300,512,3.5,'f32'),(300,640,math.log(100.0),'f32'),(300,768,math.log(100.0),'f32'),(512,512,math.log(100.0),'f32'),(300,1024,math.log(100.0),'f32')]:
    embs = closed_form.synthetic_embeddings(b, d, 31, 0.15)
    if dt=='bf16': embs=[closed_form.round_to_bf16(e) for e in embs]
    t3, g3 = (t,t,t), (0.25,0.5,0.125)
    want = closed_form.tri_contrastive(*embs, t3, g3)
    ten = [torch.from_numpy(e).cuda().to(torch.bfloat16 if dt=='bf16' else torch.float32) for e in embs]
    for st in (True, False):
        cfg = ops.TriContrastiveConfig(math="f16", grads_fp32=True, stash=st, check_status=False)
        loss3, dimg, dtxt, daud, dt3 = ops.forward_backward_raw(*ten, torch.tensor(t3,device='cuda'), torch.tensor(g3,device='cuda'), cfg)
        torch.cuda.synchronize()
        print(b,d,round(t,3),dt,'stash' if st else 'recompute', 'loss', np.max(np.abs(loss3.double().cpu().numpy()-want['loss'])/want['loss']),
              'dimg', rel(dimg.double().cpu().numpy(), want['dimg']), 'dtxt', rel(dtxt.double().cpu().numpy(), want['dtxt']), 'daud', rel(daud.double().cpu().numpy(), want['daud']),
              'dscale', np.max(np.abs(dt3.double().cpu().numpy()-want['dscale']))/np.max(np.abs(want['dscale'])), flush=True)
