"""profiling driver: where does one bench step go?  (a) wall time per phase with a sync after each phase,
(b) torch.profiler kernel table (device time per kernel name, all launches incl. torch's own),
(c) host-synchronisation points per source line (torch.cuda.set_sync_debug_mode).
usage: python tools/prof_step.py [out_prefix]      (writes <prefix>_phases.txt / _kernels.txt / _syncs.txt)"""
import collections
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import numpy as np
import torch

import bench

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "step")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
hp = bench.HotPath(dev, 1)
B = bench.BATCH_PER_GPU
host = bench.make_host_batch(B, bench.N_POINTS, seed=0)
from pnb200.input_pipeline import host_rotations
x = bench.device_input(host[0].to(dev), host[1].to(dev), torch.from_numpy(host_rotations(host[0].numpy())).to(dev))
lab, prim = host[2].to(dev), host[3].to(dev)
lab_np, prim_np = host[2].numpy(), host[3].numpy()


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def phased_step(i, rec):
    np.random.seed(i)
    t0 = sync()
    hp.opt.zero_grad(set_to_none=True)
    emb, lp, el = hp.model(x, lab, True)
    t1 = sync()
    loss = el.mean() + hp.primitive_loss(lp, prim)
    pts = x[:, 0:3].permute(0, 2, 1).contiguous()
    nrm = x[:, 3:6].permute(0, 2, 1).contiguous()
    t2 = sync()
    emb = bench.pin_clusters(emb, lab, hp.codes)
    res, extra = hp.evaluation.fitting_loss(emb.permute(0, 2, 1), pts, nrm, lab_np, prim_np.copy(), lp,
                                            quantile=0.025, iterations=bench.MS_ITERS, lamb=0.1)
    loss = loss + torch.stack([r.reshape(()) for r in res[0::5]]).mean()
    t3 = sync()
    loss.backward()
    t4 = sync()
    hp.opt.step()
    t5 = sync()
    rec.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))


for i in range(2):
    phased_step(i, [])
rec = []
for i in range(int(os.environ.get("PROF_STEPS", "3"))):
    phased_step(10 + i, rec)
names = ["seg-net fwd + triplet", "nll + input slices", "fitting_loss (bandwidth, mean-shift, nms, match, fit, residual)",
         "backward", "adam"]
with open(out + "_phases.txt", "w") as f:
    a = np.array(rec) * 1e3
    f.write(f"# wall ms per phase (sync after each phase), mean of {len(rec)} steps, B=16 x N=10000\n")
    for n, m in zip(names, a.mean(0)):
        f.write(f"{m:9.2f} ms  {n}\n")
    f.write(f"{a.sum(1).mean():9.2f} ms  total\n")
    f.write("# per step: " + " | ".join(" ".join(f"{v:.1f}" for v in row) for row in a) + "\n")
print(open(out + "_phases.txt").read())

# ---- kernel table
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    np.random.seed(20)
    hp.step(x, lab_np, prim_np, lab, prim)
    torch.cuda.synchronize()
evs = [e for e in p.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict()
for e in evs:
    n = e.name[:90]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(a[1] for a in agg.values())
with open(out + "_kernels.txt", "w") as f:
    f.write(f"# torch.profiler (CUPTI) device time per kernel, one step: {len(evs)} device activities, "
            f"{tot / 1e3:.1f} ms busy\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        f.write(f"| `{n}` | {c} | {t / 1e3:.3f} | {100 * t / tot:.1f}% |\n")
print(open(out + "_kernels.txt").read()[:6000])
with open(out + "_cpu_ops.txt", "w") as f:
    f.write(p.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=70))

# ---- host synchronisation points
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    np.random.seed(21)
    hp.step(x, lab_np, prim_np, lab, prim)
torch.cuda.set_sync_debug_mode("default")
cnt = collections.Counter()
for m in w:
    if "synchroniz" in str(m.message):
        cnt[f"{os.path.relpath(m.filename, ROOT)}:{m.lineno}"] += 1
with open(out + "_syncs.txt", "w") as f:
    f.write(f"# host-blocking synchronisations in one step: {sum(cnt.values())}\n")
    for k, v in cnt.most_common(60):
        f.write(f"{v:5d}  {k}\n")
print(open(out + "_syncs.txt").read())
