"""Loss step probe (run on the GPU box): times the fp32 loss dict fwd+bwd(+enqueue) and the EMA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import textreid_b200 as trb
from textreid_b200.synthetic import loss_inputs

def main(N=128, D=256, K=2048, C=11003, iters=20):
    dev = "cuda"
    inp = {k: v.to(dev) for k, v in loss_inputs(N, D, K, C, seed=0).items()}
    ve, te, pr = inp["v_embed"].requires_grad_(True), inp["t_embed"].requires_grad_(True), inp["projection"].requires_grad_(True)
    ptr = torch.zeros(1, dtype=torch.int64, device=dev)
    def step():
        d = trb.moco_loss_dict(ve, te, inp["v_key"], inp["t_key"], inp["labels"], inp["v_queue"], inp["t_queue"], inp["id_queue"],
                               ptr, pr, epsilon=0.1, enqueue=True, precision=os.environ.get("TRB_LOSS_PRECISION", "fp32"),
                               cuda_graph=bool(int(os.environ.get("TRB_LOSS_GRAPH", "0"))))
        ve.grad = te.grad = pr.grad = None
        (d["instance_loss"] + d["infonce_loss"] + d["global_align_loss"]).backward()
        return d
    for _ in range(5): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): d = step()
    b.record(); torch.cuda.synchronize()
    print("loss step N=%d K=%d: %.1f us/step  losses=%s" % (N, K, a.elapsed_time(b) * 1e3 / iters, {k: round(float(v), 4) for k, v in d.items()}))

if __name__ == "__main__":
    if os.environ.get("TRB_PROBE_SHAPE") == "n256":
        main(N=256, K=4096)
    else:
        main()

def graph_only(iters=200):
    """GPU time of the captured step alone (loss + gradients + enqueue), replayed back to back."""
    from textreid_b200.losses import _GRAPHS
    if not _GRAPHS:
        return
    g = next(iter(_GRAPHS.values()))[0]
    for _ in range(10): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): g.replay()
    b.record(); torch.cuda.synchronize()
    print("graph replay alone: %.1f us/step" % (a.elapsed_time(b) * 1e3 / iters))

if __name__ == "__main__":
    graph_only()
