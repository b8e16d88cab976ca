"""Segment matching on the device (csrc/assign.cu) against the host formulation the training path uses (itself bit-identical
to the reference's one-hot / relaxed_iou_fast arithmetic, tests/test_cpu_host_logic.py) and scipy's optimal assignment."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,kp,kg", [(1, 100, 3, 3), (4, 10000, 8, 8), (3, 7001, 49, 31), (2, 513, 1, 50)])
def test_iou_cost_kernel_is_bit_identical_to_the_host_cost(B, N, kp, kg):
    from pnb200.assign import iou_cost
    from src.segment_utils import iou_cost_host
    rs = np.random.RandomState(N)
    pred = rs.randint(0, kp, (B, N)); gt = rs.randint(0, kg, (B, N))
    gt[:, : N // 3] = pred[:, : N // 3] % kg                      # correlated labels: IoUs away from chance
    cost, bad = iou_cost(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), 50)
    assert int(bad) == 0
    for b in range(B):
        assert np.array_equal(cost[b].cpu().numpy(), iou_cost_host(pred[b], gt[b], 50))
    _, bad = iou_cost(torch.from_numpy(pred + 60).cuda(), torch.from_numpy(gt).cuda(), 50)
    assert int(bad) == 1


@pytest.mark.parametrize("n", [1, 2, 7, 50, 64])
def test_hungarian_kernel_finds_the_optimal_assignment(n):
    from pnb200.assign import hungarian
    rs = np.random.RandomState(n)
    costs = [rs.rand(n, n).astype(np.float32) for _ in range(6)]
    costs += [rs.randint(0, 3, (n, n)).astype(np.float32) for _ in range(3)]                 # heavy ties
    iou_like = np.ones((n, n), np.float32); k = min(n, 8); iou_like[:k, :k] = 1 - rs.rand(k, k).astype(np.float32)
    costs.append(iou_like)
    got = hungarian(torch.from_numpy(np.stack(costs)).cuda()).cpu().numpy()
    for c, cols in zip(costs, got):
        assert sorted(cols.tolist()) == list(range(n))
        r, cc = linear_sum_assignment(c)
        want = c[r, cc].astype(np.float64).sum()
        assert abs(c[np.arange(n), cols].astype(np.float64).sum() - want) <= 1e-9 * max(1.0, abs(want))
    for c, cols in zip(costs[:6], got[:6]):                        # continuous costs: the optimum is unique
        assert np.array_equal(cols, linear_sum_assignment(c)[1])


def test_match_batched_equals_host_match_on_clustered_labels():
    """the reference's `match` (cost of one-hot memberships + optimal assignment) for a batch of shapes in two launches: the
    matched pairs with a non-empty predicted cluster equal the host path's (iou_cost_host + scipy) pairs"""
    from pnb200.assign import match_batched
    from src.segment_utils import iou_cost_host
    rs = np.random.RandomState(0)
    B, N = 5, 10000
    gt = rs.randint(0, 8, (B, N))
    perm = np.stack([rs.permutation(8) for _ in range(B)])
    pred = np.take_along_axis(perm, gt, 1)
    noise = rs.rand(B, N) < 0.15
    pred[noise] = rs.randint(0, 11, noise.sum())
    cols, cost, bad = match_batched(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), 50)
    assert int(bad) == 0
    cols = cols.cpu().numpy()
    for b in range(B):
        r, c = linear_sum_assignment(iou_cost_host(pred[b], gt[b]))
        for p in np.unique(pred[b]):
            if p < 8:
                assert cols[b, p] == c[p] == np.argsort(perm[b])[p]
