"""Diagnostic (GPU box): for every spline segment of the 3-shape stage test, compare the batched spline stage with the
per-entry legacy functions GIVEN THE SAME normalised weights: standardisation, SplineNet output, reconstruction, Chamfer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import src.residual_utils as RU
from oracle.make_golden_helpers import e2e_inputs
from pnb200 import fitstage as FS
from pnb200.staging import arena
from src.fitting_utils import rotation_matrix_a_to_b, standardize_point_torch
from src.primitive_forward import forward_closed_splines, forward_pass_open_spline
from src.utils import chamfer_distance_single_shape
from test_gpu_fitstage import _nets

nets = _nets()
shapes = [e2e_inputs(1400, 70 + i, False) for i in range(3)]
pts = torch.from_numpy(np.concatenate([s[0] for s in shapes])).cuda()
nrm = torch.from_numpy(np.concatenate([s[1] for s in shapes])).cuda()
lab = np.concatenate([s[2] for s in shapes]); prim = np.concatenate([s[3] for s in shapes])
emb = torch.cat([s[4] for s in shapes]); logp = torch.cat([s[5] for s in shapes]).cuda()
ev = RU.Evaluation(open_decoder=nets["open"], closed_decoder=nets["closed"])
np.random.seed(5)
with torch.no_grad():
    res, extra = ev.fitting_loss(emb.cuda(), pts, nrm, lab, prim.copy(), logp, quantile=0.015, iterations=10, lamb=0.1)
out = ev.last_fit
plan, Wn = out["plan"], out["Wn"]
stage = arena("diag", torch.device("cuda", 0))
rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
fitter = ev.fitter
fitter._basis_on(pts.device)
for e, (b, col, key, closed, gidx) in enumerate(plan.splines):
    P = pts[b, 0::2]
    w = Wn[b, 0::2, col:col + 1] + FS.EPS
    with torch.no_grad():
        p1, std1, mean1, R1 = standardize_point_torch(P, w)
        Ps, std2, mean2, R2, Rinv2 = FS.standardize_batched(P.unsqueeze(0), w.reshape(1, -1), rotation_matrix_a_to_b, stage)
        net = fitter.closed_control_decoder if closed else fitter.open_control_decoder
        o1 = net(p1.unsqueeze(0).permute(0, 2, 1), w.t())
        o2 = net(Ps.permute(0, 2, 1), w.reshape(1, -1))
        if closed:
            rec1 = forward_closed_splines(P.unsqueeze(0), net, fitter.nu, fitter.nv, weights=w, if_optimize=False)[2]
        else:
            rec1 = forward_pass_open_spline(P.unsqueeze(0), net, fitter.nu, fitter.nv, weights=w, if_optimize=False)[1]
        rec2 = out["recs"][e]
        gt = pts[b][torch.from_numpy(gidx).cuda()]
        d1 = chamfer_distance_single_shape(rec1[0], gt)
        d2 = FS.segment_distances(out, b)[key][1]
    print(f"spline {e}: shape {b} slot {col} key {key} {'closed' if closed else 'open'} gt {len(gidx)} pts | mask>0.8: "
          f"{int((w[:, 0] > 0.8).sum())} | std pts {rel(Ps[0], p1):.1e} extents {rel(std2[0], std1.reshape(3)):.1e} R {rel(R2[0], R1):.1e} | "
          f"net out (same input batch of 1) {rel(o2, o1):.1e} | recon {rel(rec2, rec1):.1e} | chamfer legacy {float(d1):.6e} "
          f"stage {float(d2):.6e}")
