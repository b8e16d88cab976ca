"""Host / torch-level pieces of the batched fit stage (pnb200/fitstage.py) on the CPU against the oracle port: the padded
(shape, slot) weight normalisation incl. single-cluster shapes and its gradient, the batched SplineNet input frame
(standardisation) and the segment plan.  The kernels of the stage are covered by tests/test_gpu_fitstage.py."""
import numpy as np
import pytest
import torch


class FakeStage:
    """stand-in for the pinned upload arena (pinned memory needs a CUDA runtime)"""

    def upload(self, arr, device):
        return torch.from_numpy(np.ascontiguousarray(arr)).to(device)

    def reset(self):
        pass


def test_normalized_weights_batched_equals_per_shape_port():
    from oracle.port import fitting as OP
    from pnb200 import fitstage as FS
    g = torch.Generator().manual_seed(0)
    B, N, S = 3, 500, FS.SLOTS
    K = [1, 7, 49]
    bws = torch.tensor([0.31, 0.05, 0.8])
    raw = (torch.rand(B, N, S, generator=g) * 2 - 1).requires_grad_()
    Wn = FS.normalized_weights(raw, bws, K, FakeStage())
    coef = torch.randn(B, N, S, generator=g)
    (Wn * coef).sum().backward()
    for b in range(B):
        r = raw[b, :, :K[b]].detach().t().clone().requires_grad_()            # (K,N) like the reference
        want = OP.weights_normalize(r, float(bws[b]))
        torch.testing.assert_close(Wn[b, :, :K[b]].t(), want, rtol=1e-6, atol=1e-7)
        assert not Wn[b, :, K[b]:].any(), "padded slots must be exact zeros"
        (want * coef[b, :, :K[b]].t()).sum().backward()
        scale = r.grad.abs().max().item()
        if K[b] == 1:       # exact gradient is 0 (prob == 1 everywhere); both sides hold rounding residue
            assert raw.grad[b, :, :1].abs().max().item() <= 1e-5 * coef.abs().max().item() * 100
        else:
            assert (raw.grad[b, :, :K[b]].t() - r.grad).abs().max().item() <= 1e-4 * scale
        assert not raw.grad[b, :, K[b]:].any(), "no gradient may reach the padded slots"


def test_standardize_batched_equals_port_per_entry():
    from oracle.port import e2e as PE
    from pnb200 import fitstage as FS
    g = torch.Generator().manual_seed(1)
    E, n = 4, 1200
    P = torch.randn(E, n, 3, generator=g) * torch.tensor([1.0, 0.4, 0.1]) + torch.randn(E, 1, 3, generator=g)
    w = torch.rand(E, n, generator=g)
    w[0] = w[0] * 0.5                   # nothing above 0.8 -> top-k fallback branch
    w[1, :900] = 0.95                   # plenty above 0.8 -> threshold branch
    Ps, std, mean, R, Rinv = FS.standardize_batched(P, w, PE.rotation_a_to_b, FakeStage())
    for e in range(E):
        pts, std_r, mean_r, R_r = PE.standardize_point(P[e], w[e].reshape(n, 1))
        torch.testing.assert_close(mean[e], mean_r.reshape(3), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(R[e], torch.as_tensor(R_r, dtype=torch.float32).reshape(3, 3), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(std[e], std_r.reshape(3), rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(Ps[e], pts, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(Rinv[e] @ R[e], torch.eye(3), rtol=0, atol=1e-5)


def test_plan_follows_the_segment_rules_of_fit_one_shape():
    """spline cap of 4 per shape, analytic kinds into the slot of the weight column, gt points into the matched slot"""
    from pnb200 import fitstage as FS
    N = 2000
    rs = np.random.RandomState(0)
    labels = np.repeat(np.arange(8), N // 8)[None].copy()
    prim_of = np.array([1, 5, 4, 3, 2, 9, 8, 0])                  # 4 analytic, 4 splines ... the fifth spline is dropped below
    primitives = prim_of[labels]
    primitives[0, labels[0] == 3] = 6                             # segment 3 becomes a 5th spline (closed)
    cluster = labels.copy()                                       # perfect clustering, cluster id == gt label

    def match_fn(t, p):
        ids = np.arange(50)
        return ids, ids, np.unique(t), np.unique(p)

    plan = FS.make_plan(labels, cluster, primitives, N, match_fn)
    assert [plan.kind[0, c] for c in range(8)] == [0, 1, 2, -1, -1, -1, -1, -1]
    assert (plan.kind[0, 8:] == -1).all()
    kinds = {k: v[0] for k, v in plan.keys[0].items()}
    # splines in order of appearance: seg 3 (closed), 4 (open), 5 (closed), 6 (open) are fitted, seg 7 is the 5th -> dropped
    assert kinds == {0: "plane", 1: "sphere", 2: "cylinder", 3: "closed", 4: "open", 5: "closed", 6: "open", 7: None}
    assert [(s[1], s[3]) for s in plan.splines] == [(3, True), (4, False), (5, True), (6, False)]
    for c in range(3):
        np.testing.assert_array_equal(np.nonzero(plan.seg[0] == c)[0], np.nonzero(labels[0] == c)[0])
    assert (plan.seg[0][labels[0] >= 3] == -1).all()
    for s in plan.splines:
        np.testing.assert_array_equal(s[4], np.nonzero(labels[0] == s[2])[0])
