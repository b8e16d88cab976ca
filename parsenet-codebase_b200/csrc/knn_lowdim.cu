// kNN graphs of the low-dimensional inputs (reference `knn_points_normals` src/PointNet.py:29-69 on positions + normals,
// `knn` src/PointNet.py:9-26 / src/model.py:9-22 on raw positions, and the squared-difference metric of
// fitting_utils.up_sample_points_torch :150-163) with the one-pass bracketed selection of knn_tc.cu.  With 3 - 6 channels the
// distance of a pair costs a dozen FP32 operations, so the exact value (the same correctly rounded expression as knn.cu) is
// computed for every pair; what the loader-thread kernel of knn.cu spends its time on is the streaming top-k selection.  Here:
//   1. collect<ALL> : exact costs of every row to a strided column sample; T_i := b-th smallest (kth_select.cuh)
//   2. collect      : one pass over all columns, thread = query row; (cost, column) appended to the row's list when cost <= T_i
//   3. final        : k-th smallest cost L_i of the list, the entries <= L_i (k plus exact ties) sorted by (cost, index), first k
//                     written.  A short list (bracket too low), an overflow or more than 128 tied survivors FLAG the row;
//   4. knn_kernel re-does the tiles that contain a flagged row (pn_knn_flagged).
// cost = -D of knn.cu: metric 0  (xx_j - 2 <x_i, x_j>) + xx_i   [written as -((-xx_j - inner) - xx_i), the same roundings]
//                      metric 1  ((xx_j - 2 <p_i, p_j>) + xx_i) * (1 + (2 - 2 <n_i, n_j>))
//                      metric 2  ((dx dx) + (dy dy)) + dz dz
#include "common.cuh"
#include "kth_select.cuh"
#include "knn_select.cuh"

namespace pn {
namespace knnld {

constexpr int NT = 256, CB = 256, MAXSURV = 128;

template <int METRIC>
__device__ __forceinline__ float pair_cost(const float* q, float xxq, const float* v, float xxv) {
    if (METRIC == 2) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = __fsub_rn(q[c], v[c]);
            acc = __fadd_rn(acc, __fmul_rn(d, d));
        }
        return acc;
    }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) acc = fmaf(q[c], v[c], acc);
    if (METRIC == 0) {
        const float inner = __fmul_rn(-2.0f, acc);
        return -__fsub_rn(__fsub_rn(-xxv, inner), xxq);
    }
    float accn = 0.f;
#pragma unroll
    for (int c = 3; c < 6; ++c) accn = fmaf(q[c], v[c], accn);
    const float pd = __fadd_rn(__fsub_rn(xxv, __fmul_rn(2.0f, acc)), xxq);
    const float nd = __fsub_rn(2.0f, __fmul_rn(2.0f, accn));
    return __fmul_rn(pd, __fadd_rn(1.0f, nd));
}

// xx[r] = fmaf chain of the squares of the first cx channels (same as knn.cu::norms_kernel)
__global__ void norms_kernel(const float* __restrict__ x, long long rows, int ld, int cx, float* __restrict__ xx) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* p = x + r * ld;
    float acc = 0.f;
    for (int c = 0; c < cx; ++c) acc = fmaf(p[c], p[c], acc);
    xx[r] = acc;
}

// grid (ceil(N / 256), B), 256 threads, thread = query row.  Columns j = r * cstride, r < ncols, staged 256 at a time as
// channel-major rows in shared memory (all lanes of a warp read the same column: broadcasts).
template <int METRIC, bool ALL>
__global__ void __launch_bounds__(NT, 2)
collect_kernel(const float* __restrict__ x, const float* __restrict__ xx, int ld, int N, int ncols, int cstride,
               const float* __restrict__ T, int cap, unsigned* __restrict__ cval, unsigned short* __restrict__ ccol,
               int* __restrict__ cnt_out) {
    constexpr int C = METRIC == 1 ? 6 : 3;
    __shared__ __align__(16) float xs[C][CB];
    __shared__ __align__(16) float xxs[CB];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int i = blockIdx.x * NT + tid;
    const bool ok = i < N;
    const float* xb = x + (long long)b * N * ld;
    const float* xxb = xx + (long long)b * N;
    float q[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float xxq = 0.f;
    if (ok) {
#pragma unroll
        for (int c = 0; c < C; ++c) q[c] = xb[(long long)i * ld + c];
        xxq = xxb[i];
    }
    const long long grow = (long long)b * N + i;
    const float thr = (!ALL && ok) ? T[grow] : -INFINITY;
    const long long lbase = grow * cap;
    const int half_cap = cap >> 1;                 // (the final kernel reads two lists of cap / 2: the second one stays empty)
    int mine = 0;
    bool dead = !ok;
    for (int j0 = 0; j0 < ncols; j0 += CB) {
        __syncthreads();
        {
            const int r = j0 + tid;
            const bool in = r < ncols;
            const long long j = (long long)r * cstride;
#pragma unroll
            for (int c = 0; c < C; ++c) xs[c][tid] = in ? xb[j * ld + c] : 0.f;
            xxs[tid] = in ? xxb[j] : 0.f;
        }
        __syncthreads();
        const int nb = min(CB, ncols - j0);
#pragma unroll 1
        for (int g0 = 0; g0 < nb; g0 += 32) {
            float cost[32];
#pragma unroll
            for (int g = 0; g < 32; g += 4) {
                float4 ch[C];
#pragma unroll
                for (int c = 0; c < C; ++c) ch[c] = *reinterpret_cast<const float4*>(&xs[c][g0 + g]);
                const float4 n4 = *reinterpret_cast<const float4*>(&xxs[g0 + g]);
                const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float v[6];
#pragma unroll
                    for (int c = 0; c < C; ++c) v[c] = u == 0 ? ch[c].x : (u == 1 ? ch[c].y : (u == 2 ? ch[c].z : ch[c].w));
                    cost[g + u] = pair_cost<METRIC>(q, xxq, v, nn[u]);
                }
            }
            if (ALL) {
                if (ok) {
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        const uint4 kk = make_uint4(__float_as_uint(cost[u]), __float_as_uint(cost[u + 1]),
                                                    __float_as_uint(cost[u + 2]), __float_as_uint(cost[u + 3]));
                        // (slots beyond ncols of the last group hold costs of zero-filled columns: never read)
                        if (j0 + g0 + u + 3 < 1024) *reinterpret_cast<uint4*>(cval + lbase + j0 + g0 + u) = kk;
                    }
                }
            } else if (!dead) {
                uint32_t mask = 0u;
#pragma unroll
                for (int u = 0; u < 32; ++u) mask |= (cost[u] <= thr) ? (1u << u) : 0u;
                const int left = nb - g0;
                if (left < 32) mask &= (1u << left) - 1u;
                if (mask) {
                    const int add = __popc(mask);
                    if (mine + add > half_cap) {
                        mine = half_cap + 1;            // overflow: flagged by the final kernel
                        dead = true;
                    } else {
                        int slot = mine;
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            const uint32_t bit = (mask >> u) & 1u;
                            asm volatile(
                                "{\n"
                                ".reg .pred p;\n"
                                "setp.ne.u32 p, %0, 0;\n"
                                "@p st.global.u32 [%1], %2;\n"
                                "@p st.global.u16 [%3], %4;\n"
                                "}\n" ::"r"(bit), "l"(cval + lbase + slot), "r"(__float_as_uint(cost[u])), "l"(ccol + lbase + slot),
                                "h"((unsigned short)(j0 + g0 + u))
                                : "memory");
                            slot += (int)bit;
                        }
                        mine += add;
                    }
                }
            }
        }
    }
    if (!ALL && ok) { cnt_out[2 * grow] = mine; cnt_out[2 * grow + 1] = 0; }
}

// one warp per row: k-th smallest cost of the list, entries up to it sorted by (cost, index), first k written
template <int NW, typename IdxT>
__global__ void __launch_bounds__(256)
final_kernel(const unsigned* __restrict__ cval, const unsigned short* __restrict__ ccol, const int* __restrict__ cnt, int k,
             long long rows_total, IdxT* __restrict__ idx_out, float* __restrict__ dist_out, int* __restrict__ flags) {
    constexpr int CAP = 1024 * NW, HALF = CAP / 2;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows_total) return;
    const int n = cnt[2 * row];
    if (n > HALF || n < k) {
        if (lane == 0) flags[row] = 1;
        return;
    }
    unsigned Bt[NW][32];
    unsigned act[NW], valid[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        act[w] = 0u;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int p = w * 1024 + r * 32 + lane;
            const bool v = p < n;
            Bt[w][r] = v ? f2ord(__uint_as_float(cval[row * CAP + p])) : 0xffffffffu;
            act[w] |= v ? kthsel::reg_bit(r) : 0u;
        }
        valid[w] = act[w];
    }
    unsigned keep[NW][32];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int r = 0; r < 32; ++r) keep[w][r] = Bt[w][r];
        kthsel::bit_transpose32(Bt[w]);
    }
    int need = k;
    unsigned prefix = 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int c = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) c += __popc(kthsel::step_zeros(act[w], Bt[w][i]));
        c = __reduce_add_sync(FULL, c);
        const bool zero = c >= need;
        if (!zero) { need -= c; prefix |= 1u << (31 - i); }
#pragma unroll
        for (int w = 0; w < NW; ++w) act[w] = kthsel::step_next(act[w], Bt[w][i], zero);
    }
    // survivors: ordered key <= prefix (the k-th smallest); exact ties of the k-th cost are all kept and resolved by index
    int mycount = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int r = 0; r < 32; ++r) mycount += ((valid[w] & kthsel::reg_bit(r)) && keep[w][r] <= prefix) ? 1 : 0;
    const int total = __reduce_add_sync(FULL, mycount);
    if (total > MAXSURV) {
        if (lane == 0) flags[row] = 1;
        return;
    }
    if (lane == 0) flags[row] = 0;
    // keys of the survivors, gathered by a warp-wide compaction through shuffles: every lane contributes its survivors in turn
    unsigned long long key[MAXSURV / 32];
#pragma unroll
    for (int e = 0; e < MAXSURV / 32; ++e) key[e] = ~0ull;
    int incl = mycount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    int pos = incl - mycount;
    __shared__ unsigned long long stage[8][MAXSURV];
    unsigned long long* sk = stage[threadIdx.x >> 5];
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int r = 0; r < 32; ++r)
            if ((valid[w] & kthsel::reg_bit(r)) && keep[w][r] <= prefix) {
                const int p = w * 1024 + r * 32 + lane;
                const float cost = ord2f(keep[w][r]);
                sk[pos++] = knn::make_key(-cost, (int)ccol[row * CAP + p]);
            }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < MAXSURV / 32; ++e) {
        const int p = e * 32 + lane;
        if (p < total) key[e] = sk[p];
    }
    knn::warp_bitonic_sort<MAXSURV / 32>(key, lane);
#pragma unroll
    for (int e = 0; e < MAXSURV / 32; ++e) {
        const int p = e * 32 + lane;
        if (p < k) {
            idx_out[row * k + p] = (IdxT)(uint32_t)(key[e] & 0xffffffffu);
            if (dist_out) dist_out[row * k + p] = ord2f(~(uint32_t)(key[e] >> 32));
        }
    }
}

template <int METRIC>
static int run(const float* x, int B, int N, int ld, int k, void* idx_out, int idx_is_i64, float* dist_out, int stride,
               int b_sample, float* xx, float* T, unsigned* ws_val, unsigned short* ws_col, int* ws_cnt, int cap, int* flags,
               cudaStream_t st) {
    const long long rows = (long long)B * N;
    const int m = (N + stride - 1) / stride;
    norms_kernel<<<cdiv(rows, 256), 256, 0, st>>>(x, rows, ld, 3, xx);
    PN_COUNT_LAUNCH();
    const dim3 grid(cdiv(N, NT), B);
    const float* noT = nullptr;
    collect_kernel<METRIC, true><<<grid, NT, 0, st>>>(x, xx, ld, N, m, stride, noT, cap, ws_val, ws_col, ws_cnt);
    PN_COUNT_LAUNCH();
    if (cap == 1024) kthsel::kth_smallest_rows_kernel<1><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_val, m, b_sample, rows, T);
    else kthsel::kth_smallest_rows_kernel<2><<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(ws_val, m, b_sample, rows, T);
    PN_COUNT_LAUNCH();
    collect_kernel<METRIC, false><<<grid, NT, 0, st>>>(x, xx, ld, N, N, 1, (const float*)T, cap, ws_val, ws_col, ws_cnt);
    PN_COUNT_LAUNCH();
    const unsigned fg = (unsigned)cdiv(rows, 8);
    if (cap == 1024) {
        if (idx_is_i64) final_kernel<1, long long><<<fg, 256, 0, st>>>(ws_val, ws_col, ws_cnt, k, rows, (long long*)idx_out, dist_out, flags);
        else final_kernel<1, int><<<fg, 256, 0, st>>>(ws_val, ws_col, ws_cnt, k, rows, (int*)idx_out, dist_out, flags);
    } else {
        if (idx_is_i64) final_kernel<2, long long><<<fg, 256, 0, st>>>(ws_val, ws_col, ws_cnt, k, rows, (long long*)idx_out, dist_out, flags);
        else final_kernel<2, int><<<fg, 256, 0, st>>>(ws_val, ws_col, ws_cnt, k, rows, (int*)idx_out, dist_out, flags);
    }
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn_lowdim kernels");
    return PN_OK;
}

}  // namespace knnld
}  // namespace pn

using namespace pn;

// 1 if pn_knn_lowdim takes this problem: positions (metric 0 or 2, C == 3) or positions + normals (metric 1, C == 6), 16-bit columns
extern "C" int pn_knn_lowdim_supported(int N, int C, int k, int metric) {
    if (!((metric == 1 && C == 6) || ((metric == 0 || metric == 2) && C == 3))) return 0;
    if (N >= 65536 || N < 2048 || k < 1 || k > 96) return 0;
    return 1;
}

// kNN graph of pn_knn (same arguments, same bit-exact result) for the low-dimensional metrics through the one-pass bracketed
// selection described at the top.  stride / b_sample: column sample {0, stride, ...} (<= 1024 columns) and the order statistic
// used as bracket.  Workspaces: ws_norms [B*N], ws_T [B*N], ws_val [B*N][cap] u32, ws_col [B*N][cap] u16, ws_cnt [B*N][2];
// cap = 1024 or 2048 (a row's list may use cap / 2 entries).  flags [B*N] is written: rows with flag 1 are NOT written -- run
// pn_knn_flagged with the same flags next.
extern "C" int pn_knn_lowdim(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64,
                             float* dist_out, int stride, int b_sample, float* ws_norms, float* ws_T, unsigned* ws_val,
                             unsigned short* ws_col, int* ws_cnt, int cap, int* flags, void* stream) {
    PN_REQUIRE(x && idx_out && ws_norms && ws_T && ws_val && ws_col && ws_cnt && flags, "pn_knn_lowdim: null pointer");
    PN_REQUIRE(B > 0 && ld >= C && pn_knn_lowdim_supported(N, C, k, metric), "pn_knn_lowdim: unsupported problem (N=%d C=%d k=%d "
               "metric=%d)", N, C, k, metric);
    PN_REQUIRE(cap == 1024 || cap == 2048, "pn_knn_lowdim: lists are built for cap = 1024 or 2048 (got %d)", cap);
    PN_REQUIRE(stride >= 1, "pn_knn_lowdim: bad sample stride %d", stride);
    const int m = (N + stride - 1) / stride;
    PN_REQUIRE(m <= 1024 && b_sample >= 1 && b_sample <= m, "pn_knn_lowdim: need at most 1024 sample columns and 1 <= b_sample <= m "
               "(N=%d stride=%d m=%d b_sample=%d)", N, stride, m, b_sample);
    PN_REQUIRE(reinterpret_cast<uintptr_t>(ws_val) % 16 == 0, "pn_knn_lowdim: ws_val must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (metric == 0) return knnld::run<0>(x, B, N, ld, k, idx_out, idx_is_i64, dist_out, stride, b_sample, ws_norms, ws_T, ws_val,
                                          ws_col, ws_cnt, cap, flags, st);
    if (metric == 1) return knnld::run<1>(x, B, N, ld, k, idx_out, idx_is_i64, dist_out, stride, b_sample, ws_norms, ws_T, ws_val,
                                          ws_col, ws_cnt, cap, flags, st);
    return knnld::run<2>(x, B, N, ld, k, idx_out, idx_is_i64, dist_out, stride, b_sample, ws_norms, ws_T, ws_val, ws_col, ws_cnt,
                         cap, flags, st);
}
