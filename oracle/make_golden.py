"""TEST INFRASTRUCTURE ONLY — dumps golden vectors from the UNMODIFIED reference (/root/reference, CPU, torch 2.11)
into tests/golden/*.npz.  Run in the build container:  python -m oracle.make_golden [names...]

The fixtures pin the CPU restatement in oracle/port (tests/test_oracle_golden.py); the GPU parity tests then
compare the CUDA path with the port on fresh seeded inputs and with these fixtures directly.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_loader as rl  # noqa: E402
from oracle.port import common  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def grad_summary(t):
    t = t.detach().reshape(-1).double()
    return np.array([t.sum().item(), t.norm().item()] + t[:14].tolist(), dtype=np.float64)


def fix_aliases(sd):
    for i in (1, 2, 3):
        for s in ("weight", "bias"):
            a, b = f"encoder.bn{i}.{s}", f"encoder.conv{i}.1.{s}"
            if a in sd and b in sd:
                sd[b] = sd[a]
    return sd


def gen_knn():
    PN = rl.ref("src.PointNet")
    out = {}
    for name, (N, C, k, metric, seed) in {"c3": (600, 3, 10, 0, 1), "c64": (700, 64, 80, 0, 2), "pn": (650, 6, 80, 1, 3)}.items():
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(2, C, N, generator=g) * 0.3
        if metric == 1:
            x[:, 3:] = torch.nn.functional.normalize(x[:, 3:], dim=1)
        idx = (PN.knn(x, k, k) if metric == 0 else PN.knn_points_normals(x, k, k)).numpy()
        out[name + "_x"] = x.numpy(); out[name + "_idx"] = idx.astype(np.int32)
        out[name + "_meta"] = np.array([N, C, k, metric])
    np.savez_compressed(os.path.join(OUT, "knn.npz"), **out)


def gen_segnet():
    PN = rl.ref("src.PointNet"); SL = rl.ref("src.segment_loss")
    B, N, k = 2, 320, 20
    pts, nrm, lab, prim = common.synth_cloud(B, N, seed=5)
    x = torch.from_numpy(np.concatenate([pts, nrm], 2)).permute(0, 2, 1).contiguous()
    loss = SL.EmbeddingLoss(margin=1.0)
    m = PN.PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=128, primitives=True, num_primitives=10,
                                     loss_function=loss.triplet_loss, mode=5, num_channels=6, nn_nb=k)
    shapes = {n: tuple(v.shape) for n, v in m.state_dict().items()}
    sd = fix_aliases(common.seeded_state_dict(shapes, seed=11))
    m.load_state_dict(sd)
    rec = []
    ok, okn = PN.knn, PN.knn_points_normals
    PN.knn = lambda *a, **kw: rec.append(ok(*a, **kw)) or rec[-1]
    PN.knn_points_normals = lambda *a, **kw: rec.append(okn(*a, **kw)) or rec[-1]
    np.random.seed(7)
    emb, lp, el = m(x, torch.from_numpy(lab), True)
    PN.knn, PN.knn_points_normals = ok, okn
    nll = SL.primitive_loss(lp, torch.from_numpy(prim))
    total = el.sum() + nll
    total.backward()
    out = dict(points=x.numpy(), labels=lab, prims=prim, meta=np.array([B, N, k, 11, 7]),
               embedding=emb.detach().numpy(), logprob=lp.detach().numpy(), embed_loss=el.detach().numpy(),
               nll=nll.detach().numpy(), idx1=rec[0].numpy().astype(np.int32), idx2=rec[1].numpy().astype(np.int32),
               idx3=rec[2].numpy().astype(np.int32))
    out["state_keys"] = np.array(sorted(shapes.keys()))
    out["state_shapes"] = np.array([str(shapes[k_]) for k_ in sorted(shapes.keys())])
    for n, p in m.named_parameters():
        if p.grad is not None:
            out["grad:" + n] = grad_summary(p.grad)
    np.savez_compressed(os.path.join(OUT, "segnet.npz"), **out)


from oracle.make_golden_helpers import clustered_embedding  # noqa: E402


def gen_meanshift():
    MS = rl.ref("src.mean_shift")
    ms = MS.MeanShift()
    out = {}
    for name, (N, ncl, seed, q, it) in {"a": (700, 6, 3, 0.05, 10), "b": (520, 3, 4, 0.1, 5)}.items():
        X, _ = clustered_embedding(N, 128, ncl, seed)
        Xr = X.clone().requires_grad_()
        np.random.seed(seed)
        newX, center, bw, labels = ms.mean_shift(Xr, N, q, it)
        g = torch.Generator().manual_seed(seed + 100)
        w = torch.randn(center.shape, generator=g)
        w2 = torch.randn(newX.shape, generator=g) * 0.01
        ((center * w).sum() + (newX * w2).sum()).backward()
        out[name + "_X"] = X.numpy(); out[name + "_newX"] = newX.detach().numpy()
        out[name + "_center"] = center.detach().numpy(); out[name + "_bw"] = bw.numpy()
        out[name + "_labels"] = labels.numpy(); out[name + "_gradX"] = Xr.grad.numpy()
        out[name + "_meta"] = np.array([N, ncl, seed, it], dtype=np.int64); out[name + "_q"] = np.array(q)
    np.savez_compressed(os.path.join(OUT, "meanshift.npz"), **out)


GENS = {"knn": gen_knn, "segnet": gen_segnet, "meanshift": gen_meanshift}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(GENS)
    for n in names:
        GENS[n]()
        print("wrote", n)
