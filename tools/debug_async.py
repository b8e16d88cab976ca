"""debug aid: the bench's UNSYNCHRONISED step loop with per-step finiteness flags recorded on the device"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import numpy as np
import torch
import bench
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
from pnb200 import meanshift as pms
# make nms survive non-finite embeddings so that the flags can be read afterwards
_orig_nc = pms.nearest_center_batched
pms.nearest_center_batched = lambda X, Y: _orig_nc(X, Y).clamp_(0, X.shape[1] - 1)
hp = bench.HotPath(dev, world)
host = [bench.make_host_batch(bench.BATCH_PER_GPU, bench.N_POINTS, seed=100 * rank + i) for i in range(2)]
devb = [tuple(t.to(dev) for t in hb) for hb in host]
hnp = [(hb[1].numpy(), hb[2].numpy()) for hb in host]
flags = []
def fin(ts):
    return torch.stack([torch.isfinite(t).all() for t in ts]).all()
def step(i, e2e):
    if e2e:
        x, lab, prim = (t.to(dev, non_blocking=True) for t in host[i % 2])
    else:
        x, lab, prim = devb[i % 2]
    np.random.seed(i)
    hp.opt.zero_grad(set_to_none=True)
    emb, lp, el = hp.model(x, lab, True)
    loss = el.mean() + hp.primitive_loss(lp, prim)
    pts = x[:, 0:3].permute(0, 2, 1).contiguous(); nrm = x[:, 3:6].permute(0, 2, 1).contiguous()
    res, extra = hp.evaluation.fitting_loss(emb.permute(0, 2, 1), pts, nrm, hnp[i % 2][0], hnp[i % 2][1].copy(), lp,
                                            quantile=0.025, iterations=bench.MS_ITERS, lamb=0.1)
    loss = loss + torch.stack([r.reshape(()) for r in res[0::5]]).mean()
    loss.backward()
    g0 = fin([p.grad for p in hp.params if p.grad is not None])
    if world > 1:
        from pnb200.parallel import allreduce_mean_grads
        allreduce_mean_grads(hp.params, world)
    g1 = fin([p.grad for p in hp.params if p.grad is not None])
    hp.opt.step()
    w = fin(hp.params)
    gmax = torch.stack([p.grad.abs().max() for p in hp.params if p.grad is not None]).max()
    wmax = torch.stack([p.detach().abs().max() for p in hp.params]).max()
    flags.append(torch.stack([fin([emb]).float(), torch.isfinite(loss).all().float(), g0.float(), g1.float(), w.float(),
                              loss.detach().float().reshape(()), gmax, wmax]))
    return loss
seq = [(0, False), (0, True), (1, False), (1, True), (2, False), (2, True)] + [(i, False) for i in range(3)] + \
    [(i, True) for i in range(3)]          # exactly bench.py --steps 3 --warmup 3
flush = torch.empty(64 * 1024 * 1024, device=dev)
for n, (i, e2e) in enumerate(seq):
    flush.zero_()
    l = step(i, e2e)
    if e2e:
        l.item()
torch.cuda.synchronize()
F = torch.stack(flags).cpu().numpy()
np.set_printoptions(precision=4, suppress=True, linewidth=200)
print(f"[r{rank}] per step [emb ok, loss ok, grads ok before allreduce, after, weights ok, loss, max|grad|, max|w|]:\n{F}", flush=True)
