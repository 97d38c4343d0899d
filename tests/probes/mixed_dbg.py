import math, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import closed_form
from synergy_clip_b200 import ops
def rel(a,b): return float(np.sqrt(((a-b)**2).sum()/ (b**2).sum()))
for (t3,g3,stash) in [((2.6592, math.log(43.5), math.log(100.0)), (0.5,1.0,0.25), True), ((2.6592, math.log(43.5), math.log(100.0)), (0.5,1.0,0.25), False), ((2.6592,2.6592,2.6592),(0.5,1.0,0.25), True), ((math.log(43.5),)*3,(0.5,1.0,0.25), True), ((math.log(100.0),)*3,(0.5,1.0,0.25), True), ((2.6592, 2.6592, math.log(100.0)), (1.0,1.0,1.0), True), ((math.log(100.0), 2.6592, 2.6592), (1.0,1.0,1.0), True)]:
    b,d=300,768
    embs=[closed_form.round_to_bf16(e) for e in closed_form.synthetic_embeddings(b,d,77,0.25)]
    want=closed_form.tri_contrastive(*embs,t3,g3)
    ten=[torch.from_numpy(e).cuda().bfloat16() for e in embs]
    cfg=ops.TriContrastiveConfig(math="f16",grads_fp32=True,stash=stash)
    loss3,dimg,dtxt,daud,dt3=ops.forward_backward_raw(*ten,torch.tensor(t3,dtype=torch.float32,device="cuda"),torch.tensor(g3,dtype=torch.float32,device="cuda"),cfg)
    print([round(x,3) for x in t3], stash, "loss", np.abs(loss3.double().cpu().numpy()-want["loss"])/want["loss"], "grads", [rel(g.double().cpu().numpy(), want[k]) for g,k in ((dimg,"dimg"),(dtxt,"dtxt"),(daud,"daud"))], "dt", np.abs(dt3.double().cpu().numpy()-want["dscale"])/np.max(np.abs(want["dscale"])))
