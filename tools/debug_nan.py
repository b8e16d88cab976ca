"""debug aid: run bench steps on a given rank's synthetic batches and report the first non-finite tensor"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import numpy as np
import torch
import bench
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", sys.argv[1] if len(sys.argv) > 1 else "1"))
local = int(os.environ.get("LOCAL_RANK", os.environ.get("DEV", "0")))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
_print = print
def print(*a, **k):
    _print(f"[r{rank}]", *a, **k)
hp = bench.HotPath(dev, world)
host = [bench.make_host_batch(bench.BATCH_PER_GPU, bench.N_POINTS, seed=100 * rank + i) for i in range(2)]
devb = [tuple(t.to(dev) for t in hb) for hb in host]
hnp = [(hb[1].numpy(), hb[2].numpy()) for hb in host]
for i in range(steps):
    x, lab, prim = devb[i % 2]
    np.random.seed(i)
    hp.opt.zero_grad(set_to_none=True)
    emb, lp, el = hp.model(x, lab, True)
    if not torch.isfinite(emb).all():
        badrows = (~torch.isfinite(emb)).any(1).nonzero()
        print(f"step {i}: NON-FINITE embedding: {badrows.shape[0]} points, first {badrows[:6].tolist()}; weights finite:",
              all(bool(torch.isfinite(p).all()) for p in hp.model.parameters()), flush=True)
        break
    print(f"step {i}: emb finite {bool(torch.isfinite(emb).all())} lp {bool(torch.isfinite(lp).all())} el {el.tolist()}", flush=True)
    loss = el.mean() + hp.primitive_loss(lp, prim)
    pts = x[:, 0:3].permute(0, 2, 1).contiguous(); nrm = x[:, 3:6].permute(0, 2, 1).contiguous()
    ctx = torch.autograd.set_detect_anomaly(True) if os.environ.get("ANOMALY") else None
    res, extra = hp.evaluation.fitting_loss(emb.permute(0, 2, 1), pts, nrm, hnp[i % 2][0], hnp[i % 2][1].copy(), lp,
                                            quantile=0.025, iterations=bench.MS_ITERS, lamb=0.1)
    fl = torch.stack([r.reshape(()) for r in res[0::5]])
    print("   fit losses", [round(float(v), 5) for v in fl], "kinds", sorted({v[0] for v in extra[0].values() if v is not None}), flush=True)
    loss = loss + fl.mean()
    loss.backward()
    bad = [n for n, p in hp.model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print("   loss", float(loss), "non-finite grads:", bad[:6], flush=True)
    if bad or not np.isfinite(float(loss)):
        # which shape?  re-run the fitting loss per shape
        for b in range(x.shape[0]):
            e = emb[b:b + 1].detach().permute(0, 2, 1).requires_grad_()
            r, ex = hp.evaluation.fitting_loss(e, pts[b:b + 1], nrm[b:b + 1], hnp[i % 2][0][b:b + 1], hnp[i % 2][1][b:b + 1].copy(),
                                               lp[b:b + 1].detach(), quantile=0.025, iterations=bench.MS_ITERS, lamb=0.1)
            r[0].reshape(()).backward()
            ok = bool(torch.isfinite(e.grad).all())
            print(f"   shape {b}: loss {float(r[0]):.5f} grad finite {ok} params",
                  {k: (v[0] if v is not None else None) for k, v in ex[0].items()}, flush=True)
        break
    if world > 1:
        from pnb200.parallel import allreduce_mean_grads
        allreduce_mean_grads(hp.params, world)
    hp.opt.step()
