"""kNN parity: CUDA path (through the C-ABI) vs the oracle (oracle/c/knn_oracle.c) — bit-exact indices."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(B, N, C, seed, metric):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, C, generator=g) * 0.3
    if metric == 1:
        x[..., 3:] = torch.nn.functional.normalize(x[..., 3:], dim=-1)
    return x


@pytest.mark.parametrize("B,N,C,k,metric", [
    (2, 1000, 3, 10, 0), (2, 777, 6, 80, 1), (1, 2048, 64, 80, 0), (2, 500, 64, 10, 0),
    (1, 1500, 128, 10, 0), (1, 600, 256, 10, 0), (1, 80, 6, 80, 1), (1, 4000, 64, 80, 0), (1, 300, 64, 200, 0),
])
def test_knn_bit_exact_vs_oracle(B, N, C, k, metric):
    from oracle import knn as oknn
    from pnb200 import ops
    x = _cloud(B, N, C, 1234 + N, metric)
    want, wdist = oknn.knn(x.numpy(), k, metric, return_dist=True)
    got, gdist = ops.knn_graph(x.cuda(), k, metric, out_dtype=torch.int64, return_dist=True)
    got = got.cpu().numpy()
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(gdist.cpu().numpy(), wdist)


def test_knn_strided_slice_and_int32():
    from oracle import knn as oknn
    from pnb200 import ops
    full = _cloud(2, 900, 256, 7, 0).cuda()
    sl = full[:, :, 64:128]
    want = oknn.knn(sl.cpu().contiguous().numpy(), 80, 0)
    got = ops.knn_graph(sl, 80, 0)
    assert got.dtype == torch.int32
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.int64), want)


def test_knn_duplicates_tie_break_lower_index():
    from oracle import knn as oknn
    from pnb200 import ops
    x = _cloud(1, 400, 3, 3, 0)
    x[0, 200:] = x[0, :200]  # exact duplicates -> exact ties
    want = oknn.knn(x.numpy(), 20, 0)
    got = ops.knn_graph(x.cuda(), 20, 0, out_dtype=torch.int64).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_knn_full_size_self_first_and_sorted():
    """BASELINE size (N=10^4,k=80): properties that do not need the O(N^2) CPU oracle on every row."""
    from oracle import knn as oknn
    from pnb200 import ops
    x = _cloud(2, 10000, 64, 11, 0)
    idx, dist = ops.knn_graph(x.cuda(), 80, 0, out_dtype=torch.int64, return_dist=True)
    idx = idx.cpu().numpy(); dist = dist.cpu().numpy()
    assert (np.diff(dist, axis=-1) <= 0).all()
    assert (idx[:, :, 0] == np.arange(10000)[None]).mean() > 0.999
    rows = [0, 1, 4999, 9999]
    for b in range(2):
        for r in rows:
            full = oknn.knn_row(x[b].numpy(), r, 0)
            order = np.lexsort((np.arange(10000), -full))[:80]
            np.testing.assert_array_equal(idx[b, r], order)


def test_cfg2_encoder_graphs_at_full_size_bit_exact():
    """BASELINE config 2 as written: DGCNN encoder forward on B = 8 synthetic 10k-point clouds (points + normals, k = 80);
    the kNN graph of EVERY layer (positions+normals metric, then the two 64-channel feature spaces) is compared bit-exactly
    with the C oracle on 512 sampled rows per shape and layer, each layer given the input the network itself fed it
    (the strided channel slices of the concat buffer, exactly as EncoderFn passes them)."""
    from oracle import knn as oknn
    from oracle.port import common
    from pnb200 import ops
    from src.PointNet import PrimitivesEmbeddingDGCNGn
    from tools.synth import ALL_KINDS, synth_cloud
    B, N, k = 8, 10000, 80
    pts, nrm, lab, prim = synth_cloud(B, N, seed=2024, n_patches=8, kinds=ALL_KINDS)
    x = torch.from_numpy(np.concatenate([pts, nrm], 2)).cuda()                       # (B,N,6) point-major
    m = PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=128, primitives=True, num_primitives=10, loss_function=None,
                                  mode=5, num_channels=6, nn_nb=k)
    sd = common.seeded_state_dict({n_: tuple(v.shape) for n_, v in m.state_dict().items()}, seed=5)
    for i in (1, 2, 3):
        for s_ in ("weight", "bias"):
            sd[f"encoder.conv{i}.1.{s_}"] = sd[f"encoder.bn{i}.{s_}"]
    m.load_state_dict(sd)
    m.cuda()
    with torch.no_grad():
        x4, xf = m.encoder(x.permute(0, 2, 1).contiguous())                          # xf (B,256,N) view of (B,N,256)
    xf_pm = xf.permute(0, 2, 1)                                                       # point-major concat buffer
    layer_inputs = [(x, 1), (xf_pm[:, :, 0:64], 0), (xf_pm[:, :, 64:128], 0)]
    rs = np.random.RandomState(0)
    for li, (inp, metric) in enumerate(layer_inputs):
        idx = ops.knn_graph(inp, k, metric, out_dtype=torch.int64).cpu().numpy()
        host = inp.contiguous().cpu().numpy()
        assert (idx[:, :, 0] == np.arange(N)[None]).mean() > 0.999
        for b in range(B):
            for r in rs.choice(N, 512, replace=False):
                full = oknn.knn_row(host[b], int(r), metric)
                order = np.lexsort((np.arange(N), -full))[:k]
                np.testing.assert_array_equal(idx[b, r], order, err_msg=f"layer {li + 1} shape {b} row {r}")


@pytest.mark.parametrize("B,N,C,k", [
    (2, 1000, 64, 80),        # below the sampling size: single pass
    (2, 5000, 64, 80),        # sampled admission threshold + main pass
    (1, 5000, 128, 10),       # SplineNet layers (k = 10, 64-entry buffers, 3-stage ring)
    (2, 4999, 256, 10),       # ragged tile, 8 channel chunks per tile
    (1, 130, 32, 16),         # two candidate tiles, one chunk
    (3, 10000, 64, 80),       # BASELINE size
])
def test_knn_tma_kernel_equals_loader_kernel(B, N, C, k):
    """csrc/knn_tma.cu (TMA-staged tiles, admission from registers, 16-bit buffer indices) against csrc/knn.cu (itself
    bit-exact vs the C oracle at every tested size): identical indices AND identical ranked values"""
    from pnb200.cabi import call
    x = _cloud(B, N, C, 99 + N + C, 0).cuda()
    x[0, N // 2:N // 2 + 7] = x[0, 3:10]            # a few exact duplicates: ties must break to the lower index
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for name in ("pn_knn", "pn_knn_tma"):
        idx = torch.full((B, N, k), -1, dtype=torch.int32, device="cuda")
        dist = torch.full((B, N, k), float("nan"), device="cuda")
        ws = torch.empty(B * N, device="cuda")
        call(name, x.data_ptr(), B, N, C, C, k, 0, idx.data_ptr(), 0, dist.data_ptr(), ws.data_ptr(), st)
        outs.append((idx, dist))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


def test_knn_tma_strided_slice_and_int64():
    """channel slice of a wider buffer (row pitch 256 floats, offset 64) as EncoderFn passes layer inputs; int64 output"""
    from oracle import knn as oknn
    from pnb200 import ops
    full = _cloud(2, 1500, 256, 17, 0).cuda()
    sl = full[:, :, 64:128]
    want = oknn.knn(sl.cpu().contiguous().numpy(), 80, 0)
    assert ops.lib.pn_knn_tma_supported(sl.data_ptr(), 1500, 64, 256, 80, 0) == 1
    got = ops.knn_graph(sl, 80, 0, out_dtype=torch.int64)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("N", [5, 333, 4000])
def test_knn_squared_difference_metric_bit_exact(N):
    """metric 2 (up-sampling helpers): 5 smallest sum((p_i - p_j)**2) with every op its own fp32 rounding, vs the C oracle"""
    from oracle import knn as oknn
    from pnb200 import ops
    g = torch.Generator().manual_seed(N)
    x = torch.rand(2, N, 3, generator=g)
    want, wd = oknn.knn(x.numpy(), 5, 2, return_dist=True)
    got, gd = ops.knn_graph(x.cuda(), 5, 2, out_dtype=torch.int64, return_dist=True)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(gd.cpu().numpy(), wd)


def test_up_sampling_helpers_vs_reference(golden_dir):
    """up_sample_points_torch (2 rounds), the memory-efficient variant and up_sample_points_in_range against the
    unmodified reference's outputs (tests/golden/upsample.npz): same neighbours -> centroids equal to fp32 rounding"""
    import os
    from src.fitting_utils import (up_sample_points_in_range, up_sample_points_torch, up_sample_points_torch_in_range,
                                   up_sample_points_torch_memory_efficient)
    g = np.load(os.path.join(golden_dir, "upsample.npz"))
    for name in ("a", "b"):
        p = torch.from_numpy(g[name + "_p"]).cuda()
        up = up_sample_points_torch(p, 2)
        assert up.shape == g[name + "_up2"].shape
        assert np.abs(up.cpu().numpy() - g[name + "_up2"]).max() <= 2e-7
        me = up_sample_points_torch_memory_efficient(p, 1)
        assert me.shape == g[name + "_me1"].shape
        assert np.abs(me.cpu().numpy() - g[name + "_me1"]).max() <= 2e-7
    np.random.seed(3)
    op, ow = up_sample_points_in_range(torch.from_numpy(g["r_p"]).cuda(), torch.from_numpy(g["r_w"]).cuda(), 1400, 1800)
    assert np.abs(op.cpu().numpy() - g["r_out_p"]).max() <= 2e-7
    np.testing.assert_array_equal(ow.cpu().numpy(), g["r_out_w"])
    np.random.seed(4)
    assert up_sample_points_torch_in_range(torch.from_numpy(g["r_p"]).cuda(), 1000, 1500).shape == (1500, 3)
    assert up_sample_points_torch_in_range(torch.from_numpy(g["b_p"]).cuda(), 100, 600).shape == (600, 3)


# ---------------------------------------------------------------------------------------------- tensor-core filter (knn_tc.cu)
def _tc_inputs(kind, B, N, C, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "randn":
        return (torch.randn(B, N, C, generator=g) * 0.3).cuda()
    if kind == "offset":                 # large common offset: norms >> distances, the error interval is wide relative to the gaps
        return (torch.randn(B, N, C, generator=g) * 0.05 + 3.0).cuda()
    if kind == "dups":                   # many exactly repeated points: ties resolved by index, windows full of equal costs
        base = torch.randn(B, N // 8, C, generator=g)
        return base.repeat(1, 8, 1)[:, torch.randperm(N // 8 * 8, generator=g)].contiguous().cuda()
    u = torch.rand(B, N, 3, generator=g)
    W1 = torch.randn(3, 32, generator=g); W2 = torch.randn(32, C, generator=g) * 0.3
    h = torch.nn.functional.leaky_relu(torch.sin(3 * u @ W1), 0.2)
    return torch.nn.functional.leaky_relu(h @ W2 + 0.5, 0.2).cuda().contiguous()


@pytest.mark.parametrize("kind,B,N,C,k", [("randn", 2, 10000, 64, 80), ("feat", 2, 10000, 64, 80), ("feat", 3, 5000, 128, 10),
                                          ("offset", 2, 4099, 64, 20), ("dups", 2, 4096, 128, 10), ("feat", 1, 2048, 64, 96),
                                          ("feat", 3, 5000, 256, 10), ("randn", 2, 4099, 256, 80), ("dups", 1, 4096, 256, 10),
                                          ("offset", 1, 3000, 256, 20)])
def test_tensor_core_filtered_knn_equals_the_exact_kernel(kind, B, N, C, k):
    """csrc/knn_tc.cu (tcgen05 filter with a proven error interval + exact fp32 refinement of the survivors + flagged fall-back)
    must give the bit-identical graph AND ranked values of the exact kernels, on ordinary features, on features with a large
    common offset (wide intervals), and on clouds full of exact duplicates (index tie-breaks, overflowing windows)"""
    from pnb200 import ops
    x = _tc_inputs(kind, B, N, C, 11)
    old = ops.KNN_IMPL
    try:
        ops.KNN_IMPL = "tma"
        i0, d0 = ops.knn_graph(x, k, 0, return_dist=True)
        ops.KNN_IMPL = "tc"
        assert ops.knn_tc_plan(N, k) is not None
        i1, d1 = ops.knn_graph(x, k, 0, return_dist=True)
    finally:
        ops.KNN_IMPL = old
    assert torch.equal(i0, i1)
    assert torch.equal(d0, d1)
    flags = [v for kk, v in ops._KNN_WS.items() if kk[2] == ops.knn_tc_plan(N, k)[2]][0]["flags"][: B * N]
    if kind in ("randn", "feat"):
        assert flags.float().mean().item() < 0.01, flags.float().mean().item()


def test_tensor_core_filtered_knn_on_a_channel_slice_vs_oracle():
    """a 64-channel slice of a 256-wide buffer (the encoder's concat buffer: row pitch 256), int64 output, against the C oracle"""
    from pnb200 import ops
    B, N, k = 1, 3000, 80
    g = torch.Generator().manual_seed(5)
    wide = torch.randn(B, N, 256, generator=g).cuda()
    x = wide[:, :, 64:128]
    old = ops.KNN_IMPL
    try:
        ops.KNN_IMPL = "tc"
        idx = ops.knn_graph(x, k, 0, out_dtype=torch.int64)
    finally:
        ops.KNN_IMPL = old
    from oracle import knn as oknn
    want = oknn.knn(x.contiguous().cpu().numpy(), k, 0)
    assert np.array_equal(idx.cpu().numpy(), want)


# ---------------------------------------------------------------------------------------------- low-dimensional metrics (knn_lowdim.cu)
def _lowdim_inputs(kind, B, N, metric, seed):
    g = torch.Generator().manual_seed(seed)
    C = 6 if metric == 1 else 3
    if kind == "dups":
        base = torch.randn(B, N // 4, C, generator=g) * 0.3
        x = base.repeat(1, 4, 1)[:, torch.randperm(N // 4 * 4, generator=g)].contiguous()
    elif kind == "grid":                 # points on a regular lattice: masses of exactly tied distances
        side = int(round(N ** (1 / 3))) + 1
        ax = torch.arange(side, dtype=torch.float32) / side
        pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)[:N]
        x = pts.unsqueeze(0).repeat(B, 1, 1)
        if C == 6:
            x = torch.cat([x, torch.randn(B, N, 3, generator=g)], 2)
    else:
        x = torch.randn(B, N, C, generator=g) * 0.3
    if metric == 1:
        x[..., 3:] = torch.nn.functional.normalize(x[..., 3:], dim=-1)
    return x.cuda().contiguous()


@pytest.mark.parametrize("kind,B,N,k,metric", [("randn", 2, 10000, 80, 1), ("randn", 3, 5000, 10, 0), ("randn", 2, 4099, 5, 2),
                                               ("dups", 2, 4096, 80, 1), ("grid", 1, 4000, 10, 0), ("randn", 1, 2048, 96, 1)])
def test_lowdim_bracketed_knn_equals_the_loader_thread_kernel(kind, B, N, k, metric):
    """csrc/knn_lowdim.cu (exact costs, one-pass bracketed selection, flagged fall-back) against the loader-thread kernel of
    knn.cu (itself bit-exact vs the C oracle): identical indices and ranked values for positions + normals, raw positions and
    the squared-difference metric; duplicates and lattice points exercise the index tie-breaks and the overflow flags"""
    from pnb200 import ops
    x = _lowdim_inputs(kind, B, N, metric, 21)
    old = ops.KNN_IMPL
    try:
        ops.KNN_IMPL = "tma"
        i0, d0 = ops.knn_graph(x, k, metric, return_dist=True)
        ops.KNN_IMPL = "tc"
        assert ops.knn_tc_plan(N, k, single_list=True) is not None
        i1, d1 = ops.knn_graph(x, k, metric, return_dist=True)
    finally:
        ops.KNN_IMPL = old
    assert torch.equal(i0, i1)
    assert torch.equal(d0, d1)


def test_lowdim_bracketed_knn_vs_oracle_int64():
    from oracle import knn as oknn
    from pnb200 import ops
    x = _lowdim_inputs("randn", 1, 3000, 1, 8)
    old = ops.KNN_IMPL
    try:
        ops.KNN_IMPL = "tc"
        idx, dist = ops.knn_graph(x, 80, 1, out_dtype=torch.int64, return_dist=True)
    finally:
        ops.KNN_IMPL = old
    want, wdist = oknn.knn(x.cpu().numpy(), 80, 1, return_dist=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
    np.testing.assert_array_equal(dist.cpu().numpy(), wdist)
