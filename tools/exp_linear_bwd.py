"""A/B of the register cap of the FP32-pipe linear backward kernels (PN_LIN_BWD_OCC = 1: no cap / no spills, 2: 128
registers, 2 CTAs per SM) on the segmentation network backward at the BASELINE shape (16 x 10k points)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import numpy as np
import torch
from pnb200 import cabi
from src.PointNet import PrimitivesEmbeddingDGCNGn
from tools.synth import synth_cloud

B, N = 16, 10000
pts, nrm, lab, prim = synth_cloud(B, N, seed=0, n_patches=8)
x = torch.from_numpy(np.concatenate([pts, nrm], 2)).permute(0, 2, 1).contiguous().cuda()
torch.manual_seed(0)
m = PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=128, primitives=True, num_primitives=10, loss_function=None, mode=5,
                              num_channels=6, nn_nb=80).cuda()
labd = torch.from_numpy(lab).cuda()
for occ in ("2", "1", "2", "1"):
    os.environ["PN_LIN_BWD_OCC"] = occ
    cabi.TIMED["pn_linear_bwd_weight"] = []; cabi.TIMED["pn_linear_bwd_data"] = []
    emb, lp, _ = m(x, labd, False)
    (emb.square().mean() + lp.mean()).backward()
    torch.cuda.synchronize()
    w = sum(a.elapsed_time(b) for a, b in cabi.TIMED.pop("pn_linear_bwd_weight"))
    d = sum(a.elapsed_time(b) for a, b in cabi.TIMED.pop("pn_linear_bwd_data"))
    print(f"PN_LIN_BWD_OCC={occ}: bwd_weight {w:.2f} ms, bwd_data {d:.2f} ms per seg-net backward", flush=True)
