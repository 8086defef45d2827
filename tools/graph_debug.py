import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import textreid_b200 as trb
from textreid_b200.synthetic import loss_inputs
KEYS = ("instance_loss", "infonce_loss", "global_align_loss")
DEV = "cuda"
inp = loss_inputs(32, 64, 128, 257, seed=7)
def run(graph, precision, steps=3):
    a = {k: v.clone().to(DEV) for k, v in inp.items()}
    ptr = torch.zeros(1, dtype=torch.int64, device=DEV)
    ve, te, pr = a["v_embed"].requires_grad_(True), a["t_embed"].requires_grad_(True), a["projection"].requires_grad_(True)
    hist = []
    for step in range(steps):
        d = trb.moco_loss_dict(ve, te, a["v_key"], a["t_key"], a["labels"], a["v_queue"], a["t_queue"], a["id_queue"],
                               ptr, pr, epsilon=0.1, enqueue=True, precision=precision, cuda_graph=graph)
        ve.grad = te.grad = pr.grad = None
        sum(d.values()).backward()
        torch.cuda.synchronize()
        hist.append(dict(loss=[float(d[k]) for k in KEYS], gv=ve.grad.clone(), gt=te.grad.clone(), gp=pr.grad.clone(), ptr=int(ptr),
                         vq=a["v_queue"].clone()))
    return hist
for prec in ("fp32", "bf16"):
    e1, e2, g = run(False, prec), run(False, prec), run(True, prec)
    for s in range(3):
        for k in ("gv", "gt", "gp", "vq"):
            print(prec, "step", s, k, "eager-vs-eager %.3e  eager-vs-graph %.3e" % (float((e1[s][k] - e2[s][k]).abs().max()), float((e1[s][k] - g[s][k]).abs().max())),
                  "loss", e1[s]["loss"] == g[s]["loss"])
