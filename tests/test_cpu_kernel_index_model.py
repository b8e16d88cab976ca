"""CPU model of the thread / shared-memory index arithmetic of csrc/meanshift.cu::ms_bwd_sparse_kernel (the experimental
sparse-row mean-shift backward, written without GPU access): the 256 threads of one CTA are replayed in Python with the
SAME index expressions as the CUDA source (tile loaders, register-tile orientation, the three P tiles, gemm_pr, the
partial-sum layout), and the result is compared with the closed form of oracle/port/meanshift.py.  A transposed tile or a
swapped (i, j) in the kernel's indexing shows up here as a wrong gradient."""
import os
import pytest
import numpy as np
import torch

D, T, NT, PK, PR = 128, 64, 256, 68, 132


def load_tile(g, r0, nrows, want_R, want_Kt):
    """csrc/meanshift.cu::load_tile: thread t loads row r0 + (t & 63), float4 column groups (t >> 6) + 4 i"""
    R = np.zeros((T, PR), np.float32) if want_R else None
    Kt = np.zeros((D, PK), np.float32) if want_Kt else None
    for t in range(NT):
        r = t & 63
        ok = (r0 + r) < nrows
        for i in range(8):
            c4 = (t >> 6) + 4 * i
            v = g[r0 + r, 4 * c4:4 * c4 + 4] if ok else np.zeros(4, np.float32)
            if R is not None:
                R[r, 4 * c4:4 * c4 + 4] = v
            if Kt is not None:
                for u in range(4):
                    Kt[4 * c4 + u, r] = v[u]
    return R, Kt


def gemm_pr(P, R, ty, tx, acc):
    """acc[4][8] += P[kk][4ty + i] * R[kk][4tx + j | 64 + 4tx + j - 4]"""
    for kk in range(T):
        a = P[kk, 4 * ty:4 * ty + 4]
        b = np.concatenate([R[kk, 4 * tx:4 * tx + 4], R[kk, 64 + 4 * tx:64 + 4 * tx + 4]])
        acc += np.outer(a, b)


def sparse_cta(Yp, Gn, gd, X, N, c, j0):
    """one CTA of ms_bwd_sparse_kernel: owns rows j0 .. j0+63 of X; returns (gX rows of the block, partial gY [64][128])"""
    _, Xt = load_tile(X, j0, N, False, True)
    _, Yt = load_tile(Yp, 0, T, False, True)
    _, Gt = load_tile(Gn, 0, T, False, True)
    P1 = np.zeros((T, PK), np.float32); P2 = np.zeros((T, PK), np.float32); Ps = np.zeros((T, PK), np.float32)
    for tid in range(NT):
        ty, tx = tid >> 4, tid & 15
        s = np.zeros((4, 4), np.float32); g = np.zeros((4, 4), np.float32)          # [jj][ii]
        for kk in range(D):
            av = Xt[kk, 4 * ty:4 * ty + 4]; yv = Yt[kk, 4 * tx:4 * tx + 4]; gv = Gt[kk, 4 * tx:4 * tx + 4]
            s += np.outer(av, yv); g += np.outer(av, gv)
        p1 = np.zeros((4, 4), np.float32); p2 = np.zeros((4, 4), np.float32)
        for jj in range(4):
            jv = (j0 + 4 * ty + jj) < N
            for ii in range(4):
                e = (s[jj, ii] - 1.0) * c
                cl = (e > 75.0) or (e < -75.0)
                k = np.exp(np.float32(min(max(e, -75.0), 75.0))) if jv else 0.0
                p2[jj, ii] = k
                p1[jj, ii] = (g[jj, ii] + gd[4 * tx + ii]) * k * c if (jv and not cl) else 0.0
        for ii in range(4):
            P1[4 * tx + ii, 4 * ty:4 * ty + 4] = p1[:, ii]
            P2[4 * tx + ii, 4 * ty:4 * ty + 4] = p2[:, ii]
        for jj in range(4):
            Ps[4 * ty + jj, 4 * tx:4 * tx + 4] = p1[jj, :]
    Xr, _ = load_tile(X, j0, N, True, False)
    Yr, _ = load_tile(Yp, 0, T, True, False)
    Gr, _ = load_tile(Gn, 0, T, True, False)
    gX = np.zeros((T, D), np.float32); part = np.zeros((T, D), np.float32)
    for tid in range(NT):
        ty, tx = tid >> 4, tid & 15
        o = np.zeros((4, 8), np.float32); q = np.zeros((4, 8), np.float32)
        gemm_pr(P1, Yr, ty, tx, o); gemm_pr(P2, Gr, ty, tx, o); gemm_pr(Ps, Xr, ty, tx, q)
        for i in range(4):
            gX[4 * ty + i, 4 * tx:4 * tx + 4] = o[i, :4]; gX[4 * ty + i, 64 + 4 * tx:64 + 4 * tx + 4] = o[i, 4:]
            part[4 * ty + i, 4 * tx:4 * tx + 4] = q[i, :4]; part[4 * ty + i, 64 + 4 * tx:64 + 4 * tx + 4] = q[i, 4:]
    return gX, part


def test_sparse_backward_kernel_index_model_matches_closed_form():
    from oracle.port import meanshift as oms
    g = torch.Generator().manual_seed(0)
    N, K, bw = 150, 7, 0.45                      # 3 column blocks, the last one ragged (150 = 2 * 64 + 22)
    X = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    rows = torch.tensor([2, 5, 63, 64, 100, 128, 149])
    Yprev = torch.nn.functional.normalize(X[rows] + 0.05 * torch.randn(K, D, generator=g), dim=1)
    # one forward iteration for those rows (values the product keeps from the forward kernel)
    Kmat = torch.exp(torch.clamp((Yprev @ X.t() - 1.0) / bw ** 2, -75.0, 75.0))
    den = Kmat.sum(1)
    u = (Kmat @ X) / den[:, None]
    un = u.norm(dim=1)
    Ynew = u / un[:, None]
    gout = torch.randn(K, D, generator=g)
    want_gY, want_gX = oms.sparse_rows_backward(gout, Ynew, Yprev, den, un, X, bw)
    # compact 64-row arrays exactly as the host layer builds them: slots beyond K repeat row 0 with a zero gradient
    pad = lambda t: torch.cat([t, t[:1].expand(T - K, *t.shape[1:])], 0)
    gpad = torch.cat([gout, torch.zeros(T - K, D)], 0)
    Yn_p, Yp_p, den_p, un_p = pad(Ynew), pad(Yprev), pad(den), pad(un)
    # prep (ms_bwd_prep_kernel)
    dot = (gpad * Yn_p).sum(1, keepdim=True)
    gu = (gpad - Yn_p * dot) / un_p[:, None]
    Gn = (gu / den_p[:, None]).numpy().astype(np.float32)
    gd = (-((gu * Yn_p).sum(1) * un_p) / den_p).numpy().astype(np.float32)
    c = np.float32(1.0 / bw ** 2)
    gX = np.zeros((N, D), np.float32); gY = np.zeros((T, D), np.float32)
    for blk in range((N + T - 1) // T):
        gx_blk, part = sparse_cta(Yp_p.numpy(), Gn, gd, X.numpy(), N, c, blk * T)
        n = min(T, N - blk * T)
        gX[blk * T:blk * T + n] += gx_blk[:n]
        assert np.abs(gx_blk[n:]).max(initial=0.0) == 0.0            # rows beyond N contribute nothing
        gY += part
    assert np.abs(gY[K:]).max() == 0.0                               # padded slots: exactly zero
    rel = lambda a, b: np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
    assert rel(gY[:K], want_gY.numpy()) < 1e-5
    assert rel(gX, want_gX.numpy()) < 1e-5


# ------------------------------------------------------------------------------------------------ TMA-fed operand addressing
# Model of what the tensor core reads in csrc/meanshift_tma.cu, built from the two layout facts the B200 probes established
# (profiles/r01_tc_probe.md): (1) a TMA box with CU_TENSOR_MAP_SWIZZLE_128B lands as rows of 128 B whose 16-byte chunk index
# is XOR-ed with bits 7..9 of the shared-memory address; (2) a K-major kind::tf32 descriptor with layout type 2 / SBO 1024
# reads row n of a K = 8 step at start + (n >> 3) * 1024 + (n & 7) * 128 + k * 4 with the same XOR applied to the address.
# The producer's TMA coordinates / destinations and the issuer's descriptor offsets are transcribed from the kernel source;
# the check is that every MMA of both products sees exactly the rows / columns of the streamed tile it is meant to see,
# for one CTA and for CTA pairs (each CTA staging half of both operands).
BN = 32
PART_BYTES = BN * D * 4


def _swz(addr):
    return addr ^ (((addr >> 7) & 7) << 4)


def tma_box(smem, dst, tensor, c0, c1, box0, box1):
    """tensor [rows][cols] (cols contiguous), box {box0 cols, box1 rows} at (c0, c1) -> smem bytes at dst (zero fill OOB)"""
    assert dst % 1024 == 0 or (dst % 128 == 0)
    for r in range(box1):
        for cc in range(box0):
            rr, col = c1 + r, c0 + cc
            v = tensor[rr, col] if (rr < tensor.shape[0] and col < tensor.shape[1]) else 0.0
            smem[_swz(dst + r * box0 * 4 + cc * 4) // 4] = v


def umma_b(smem, start, nrows):
    """the [nrows][8] B operand of one K = 8 MMA step whose descriptor start address is `start`"""
    out = np.zeros((nrows, 8), np.float32)
    for n in range(nrows):
        for k in range(8):
            out[n, k] = smem[_swz(start + (n >> 3) * 1024 + (n & 7) * 128 + k * 4) // 4]
    return out


def _check_stream(cg, N, t, X, Xs):
    rows, slab, part, drows = BN // cg, (BN // cg) * 128, PART_BYTES // cg, D // cg
    stage = 4 * part
    Xt, Xst = np.ascontiguousarray(X.T), np.ascontiguousarray(Xs.T)
    smems = []
    for rank in range(cg):                                   # tma_producer<CG>
        smem = np.full(stage // 4, np.nan, np.float32)
        row = t * BN + rank * rows
        for sl in range(4):
            tma_box(smem, sl * slab, X, 32 * sl, row, 32, rows)
            tma_box(smem, part + sl * slab, Xs, 32 * sl, row, 32, rows)
        tma_box(smem, 2 * part, Xt, t * BN, rank * drows, 32, drows)
        tma_box(smem, 3 * part, Xst, t * BN, rank * drows, 32, drows)
        assert not np.isnan(smem).any()                      # the boxes tile the stage exactly
        smems.append(smem)
    tile = np.zeros((BN, D), np.float32); tile_s = np.zeros((BN, D), np.float32)
    n_valid = max(0, min(BN, N - t * BN))
    tile[:n_valid] = X[t * BN:t * BN + n_valid]; tile_s[:n_valid] = Xs[t * BN:t * BN + n_valid]
    for ks in range(D // 8):                                 # first product: N = 32 tile rows, K = d
        off = (ks >> 2) * slab + (ks & 3) * 32
        big = np.concatenate([umma_b(s, off, rows) for s in smems], 0)
        small = np.concatenate([umma_b(s, part + off, rows) for s in smems], 0)
        assert np.array_equal(big, tile[:, ks * 8:ks * 8 + 8]) and np.array_equal(small, tile_s[:, ks * 8:ks * 8 + 8]), ks
    for ks in range(BN // 8):                                # second product: N = 128 d rows, K = tile row
        off = ks * 32
        big = np.concatenate([umma_b(s, 2 * part + off, drows) for s in smems], 0)
        small = np.concatenate([umma_b(s, 3 * part + off, drows) for s in smems], 0)
        assert np.array_equal(big, tile[ks * 8:ks * 8 + 8].T) and np.array_equal(small, tile_s[ks * 8:ks * 8 + 8].T), ks


def test_tma_fed_operand_addressing_one_cta_and_cta_pairs():
    rs = np.random.RandomState(0)
    N = 150                                                  # 4.7 tiles: the last one is ragged (TMA zero fill)
    X = rs.randn(N, D).astype(np.float32)
    Xs = (X - (X.view(np.int32) & -8192).view(np.float32)).astype(np.float32)
    Np = (N + 31) // 32 * 32
    Xp = np.zeros((Np, D), np.float32); Xp[:N] = X           # transposed forms are written for Np columns (zeros beyond N)
    Xsp = np.zeros((Np, D), np.float32); Xsp[:N] = Xs
    for cg in (1, 2):
        for t in (0, 3, 4):
            # (the zero rows of Xp beyond N stand for TMA's out-of-bounds zero fill of the N-row row-major forms)
            _check_stream(cg, N, t, Xp, Xsp)


# ------------------------------------------------------------------------------------------------ kNN selection algorithm
def _stream_select(d, k, cap, tau0, lo, hi):
    """the per-row selection of csrc/knn.cu over candidates lo..hi-1 in tiles of 128 / warps of 32: a candidate is admitted
    when d > tau; when a ballot would overflow the buffer it is compacted to the best k (value desc, index asc) and tau
    becomes the k-th best.  Returns (buffer of (value, index), tau)."""
    buf, tau = [], tau0
    for j0 in range(lo, hi, 32):
        js = [j for j in range(j0, min(j0 + 32, hi)) if d[j] > tau]
        if not js:
            continue
        if len(buf) + len(js) > cap:
            buf.sort(key=lambda e: (-e[0], e[1]))
            buf = buf[:k]
            if len(buf) == k:
                tau = buf[-1][0]
        buf += [(d[j], j) for j in js]          # (admission was decided with the tau of before the compaction, like the kernel)
    return buf, tau


def _knn_row(d, k, cap, sampled_r=0, sample_m=1024):
    n = len(d)
    tau0 = -np.inf
    if sampled_r:
        buf, _ = _stream_select(d, sampled_r, cap, -np.inf, 0, min(n, sample_m))
        buf.sort(key=lambda e: (-e[0], e[1]))
        tau0 = buf[sampled_r - 1][0] if len(buf) >= sampled_r else -np.inf
    buf, _ = _stream_select(d, k, cap, tau0, 0, n)
    if len(buf) < k:                                          # third phase of the kernel: exact re-run from -inf
        buf, _ = _stream_select(d, k, cap, -np.inf, 0, n)
    buf.sort(key=lambda e: (-e[0], e[1]))
    return [j for _, j in buf[:k]]


def test_knn_selection_is_exact_for_every_capacity_and_with_the_sampled_threshold():
    """streaming admission + compaction gives exactly the k best (ties to the lower index) whatever the buffer capacity,
    and also when the main pass starts from the sampled threshold (PN_KNN_SAMPLE) -- including rows full of duplicates,
    where the sampled threshold admits fewer than k candidates and the exact re-run has to kick in"""
    rs = np.random.RandomState(1)
    k = 80
    rows = [(-rs.rand(6000)).astype(np.float32),                                   # generic
            (-np.round(rs.rand(6000) * 40) / 40).astype(np.float32),               # heavy ties
            np.concatenate([np.zeros(3000, np.float32), -np.ones(3000, np.float32)]),   # two values only
            (-np.abs(rs.randn(6000)) ** 3).astype(np.float32)]                     # many near-zero distances
    admitted_fewer = 0
    for d in rows:
        want = sorted(range(len(d)), key=lambda j: (-d[j], j))[:k]
        for cap in (128, 256):
            assert _knn_row(d, k, cap) == want
            assert _knn_row(d, k, cap, sampled_r=25) == want
        buf, _ = _stream_select(d, k, 256, sorted(d[:1024])[-25], 0, len(d))
        admitted_fewer += len(buf) < k
    assert admitted_fewer >= 1          # the duplicate-heavy rows do exercise the fallback


# ------------------------------------------------------------------------------------------------ operand preparation
def _tf32_small(v):
    return (v - (v.view(np.int32) & -8192).view(np.float32)).astype(np.float32)


def test_operand_preparation_kernels_index_model():
    """csrc/meanshift_tma.cu::ms_prep_operands_kernel / ms_prep_concat_kernel replayed thread by thread: small split parts,
    transposes with zero padding, and the interleaved [16 rows of Y | 16 rows of Gn] tiles of the cols-backward kernel"""
    rs = np.random.RandomState(0)
    N = 75
    Np, Nq = (N + 31) // 32 * 32, (N + 15) // 16 * 16
    X = rs.randn(N, D).astype(np.float32); Y = rs.randn(N, D).astype(np.float32); G = rs.randn(N, D).astype(np.float32)
    # ---- ms_prep_operands_kernel: grid (Np / 32), 256 threads
    Xs = np.full((N, D), np.nan, np.float32); Xt = np.full((D, Np), np.nan, np.float32); Xst = np.full((D, Np), np.nan, np.float32)
    for blk in range(Np // 32):
        j0 = blk * 32
        tb = np.zeros((32, D + 1), np.float32); ts = np.zeros((32, D + 1), np.float32)
        for tid in range(256):
            for e in range(tid, 32 * D, 256):
                r, c = e >> 7, e & 127
                j = j0 + r
                v = X[j, c] if j < N else np.float32(0)
                sm = _tf32_small(np.array([v], np.float32))[0]
                if j < N:
                    Xs[j, c] = sm
                tb[r, c] = v; ts[r, c] = sm
        for tid in range(256):
            for e in range(tid, D * 32, 256):
                dd, r = e >> 5, e & 31
                Xt[dd, j0 + r] = tb[r, dd]; Xst[dd, j0 + r] = ts[r, dd]
    want_s = _tf32_small(X)
    assert np.array_equal(Xs, want_s)
    assert np.array_equal(Xt[:, :N], X.T) and np.array_equal(Xst[:, :N], want_s.T)
    assert (Xt[:, N:] == 0).all() and (Xst[:, N:] == 0).all()
    # ---- ms_prep_concat_kernel: grid (Nq / 16), 256 threads
    rows2 = 2 * Nq
    C = np.full((rows2, D), np.nan, np.float32); Ct = np.full((D, rows2), np.nan, np.float32)
    for t in range(Nq // 16):
        tb = np.zeros((32, D + 1), np.float32)
        for tid in range(256):
            for e in range(tid, 32 * D, 256):
                r, c = e >> 7, e & 127
                i = 16 * t + (r & 15)
                src = Y if r < 16 else G
                v = src[i, c] if i < N else np.float32(0)
                C[32 * t + r, c] = v
                tb[r, c] = v
        for tid in range(256):
            for e in range(tid, D * 32, 256):
                dd, r = e >> 5, e & 31
                Ct[dd, 32 * t + r] = tb[r, dd]
    assert not np.isnan(C).any() and np.array_equal(Ct, C.T)
    for t in range(Nq // 16):
        n = max(0, min(16, N - 16 * t))
        assert np.array_equal(C[32 * t:32 * t + n], Y[16 * t:16 * t + n])                 # tile rows 0..15: Y   -> S^T columns
        assert np.array_equal(C[32 * t + 16:32 * t + 16 + n], G[16 * t:16 * t + n])       # tile rows 16..31: Gn -> G^T columns
        assert (C[32 * t + n:32 * t + 16] == 0).all() and (C[32 * t + 16 + n:32 * t + 32] == 0).all()


# ------------------------------------------------------------------------------------- bracketed K-th distance (round 2)
def test_kth_select_host_twin_and_bracket_plan(tmp_path):
    """csrc/kth_select.cuh: the lane-local code of the warp K-th select (bit-transposed registers, one AND / POPC per key bit)
    run for 32 emulated lanes on the host (tests/c/kth_select_host.cpp) against a sort, for every list length the kernels
    use; and pnb200.meanshift.kth_bracket_plan: the sample order statistic sits 7 sigma above the hypergeometric mean and
    the expected list length fits the chosen capacity"""
    import ctypes, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "kthsel.so")
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-I", os.path.join(root, "parsenet-codebase_b200", "csrc"),
                           os.path.join(root, "tests", "c", "kth_select_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.kth_select_host.restype = ctypes.c_uint
    rng = np.random.default_rng(0)
    for n, K in [(1, 1), (5, 3), (33, 33), (1000, 43), (1024, 150), (430, 150), (430, 1), (777, 777), (64, 100)]:
        for _ in range(10):
            hi = int(rng.choice([8, 1 << 16, 1 << 32]))
            k = rng.integers(0, hi, size=n, dtype=np.uint64).astype(np.uint32)
            got = lib.kth_select_host(k.ctypes.data_as(ctypes.c_void_p), n, K)
            assert got == np.sort(k)[min(K, n) - 1], (n, K)
    # the plan: read the function out of the module source (importing pnb200 needs the CUDA library)
    src = open(os.path.join(root, "parsenet-codebase_b200", "pnb200", "meanshift.py")).read()
    ns = {}
    exec(src[src.index("KTH_CAPS ="):src.index("# one-pass bracketed K-th distance")] +
         src[src.index("def kth_bracket_plan"):src.index("def _kth_all_rows")], ns)
    plan = ns["kth_bracket_plan"]
    assert plan(10000, 150) == (10, 43, 1024) and plan(10000, 250)[2] == 2048
    assert plan(1000, 15) is None and plan(70000, 1000) is None and plan(10000, 3000) is None
    for N, K in [(10000, 150), (10000, 250), (10000, 562), (5000, 75), (2048, 30), (20000, 100)]:
        p = plan(N, K)
        assert p is not None, (N, K)
        stride, b, cap = p
        m = -(-N // stride)
        assert m <= 1024 and b <= m
        # Monte Carlo: how often does a random column sample put fewer than b ... i.e. the b-th sample point below the K-th?
        fails = 0
        for _ in range(2000):
            below = rng.hypergeometric(K, N - K, m)          # sample points among the K smallest
            fails += below >= b
        assert fails == 0
        assert b * N / m < cap / 2 * 1.6


# ------------------------------------------------------------------------------------- error model of the kNN filter (round 2)
def _tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)


def _f32_toward_zero(v64):
    f = v64.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(v64)
    f[over] = np.nextafter(f[over], np.float32(0))
    return f


@pytest.mark.parametrize("C,kind", [(64, "randn"), (128, "offset"), (256, "randn"), (256, "offset")])
def test_knn_filter_error_interval_covers_a_pessimistic_tensor_core_model(C, kind):
    """csrc/knn_tc.cu ranks pairs by a split-TF32 tensor-core cost and relies on |cost~ - cost| <= c0 (xx_i + xx_j) with
    c0 = knn_tc_c0(C).  Numpy model of the worst the hardware may do -- operands truncated to tf32 (big part) and the small
    part truncated again, the small x small product dropped (C = 64 / 128) , every K = 8 group summed exactly but the fp32
    accumulator rounded TOWARD ZERO after each of the 3 C / 8 MMAs -- against the reference-order fp32 fmaf chain the exact
    kernels use: the observed error must stay below HALF of the interval the kernel uses."""
    rs = np.random.RandomState(C)
    n = 300
    x = (rs.randn(n, C) * (0.05 if kind == "offset" else 0.3) + (3.0 if kind == "offset" else 0.0)).astype(np.float32)
    hi = _tf32_trunc(x)
    lo = _tf32_trunc((x - hi).astype(np.float32))
    # exact kernels: fmaf chain in channel order (float64 product is exact for fp32 operands; one rounding per step)
    chain = np.zeros((n, n), np.float32)
    for c in range(C):
        chain = (chain.astype(np.float64) + x[:, c:c + 1].astype(np.float64) * x[None, :, c].astype(np.float64)).astype(np.float32)
    xx = np.zeros(n, np.float32)
    for c in range(C):
        xx = (xx.astype(np.float64) + x[:, c].astype(np.float64) ** 2).astype(np.float32)
    inner = (np.float32(-2.0) * chain).astype(np.float32)
    cost = -(((-xx[None, :]) - inner).astype(np.float32) - xx[:, None]).astype(np.float32)
    # tensor-core model
    H, L = hi.astype(np.float64), lo.astype(np.float64)
    if C == 256:
        # collect256_kernel: the two split parts of the query rows are stacked along the TMEM lanes, two MMA chains (against the
        # small and the big part of the column tile) accumulate hi.(lo + hi) in one half of the lanes and lo.(lo + hi) in the
        # other; the halves are added in fp32 afterwards
        acc_h = np.zeros((n, n), np.float32); acc_l = np.zeros((n, n), np.float32)
        for k0 in range(0, C, 8):
            for B in (L, H):
                acc_h = _f32_toward_zero(acc_h.astype(np.float64) + H[:, k0:k0 + 8] @ B[:, k0:k0 + 8].T)
                acc_l = _f32_toward_zero(acc_l.astype(np.float64) + L[:, k0:k0 + 8] @ B[:, k0:k0 + 8].T)
        acc = (acc_h + acc_l).astype(np.float32)
    else:
        acc = np.zeros((n, n), np.float32)
        for k0 in range(0, C, 8):
            for A, B in ((L, H), (H, L), (H, H)):
                acc = _f32_toward_zero(acc.astype(np.float64) + A[:, k0:k0 + 8] @ B[:, k0:k0 + 8].T)
    cost_tc = ((np.float32(-2.0) * acc + xx[None, :]).astype(np.float32) + xx[:, None]).astype(np.float32)
    e = 3.0 / 1048576.0 + (3.0 * C / 8.0 + 8.0) / 4194304.0 + C / 16777216.0 + 8.0 / 16777216.0
    c0 = np.float32(1.25 * e)                                    # knn_tc_c0 (csrc/knn_tc.cu)
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "parsenet-codebase_b200", "csrc",
                            "knn_tc.cu")).read()
    assert "3.0 / 1048576.0 + (3.0 * C / 8.0 + 8.0) / 4194304.0 + C / 16777216.0 + 8.0 / 16777216.0" in src and "1.25 * e" in src
    ratio = np.abs(cost_tc.astype(np.float64) - cost.astype(np.float64)) / (c0 * (xx[:, None] + xx[None, :]).astype(np.float64))
    assert ratio.max() < 0.5, ratio.max()
