// kNN graph build with TMA-staged point blocks (feature-space metric, C % 32 == 0): same arithmetic contract as knn.cu
// (bit-exact with oracle/c/knn_oracle.c: D = (-xx_j - (-2 dot(x_i, x_j))) - xx_i, dot = fmaf chain over the channels in
// increasing order, ranking D descending with ties to the lower index) and the same streaming top-k with per-row buffers.
// Replaces knn (reference src/PointNet.py:9-26, src/model.py:9-22) for the 64 / 128 / 256-channel feature spaces.
//
// What changed against knn.cu, and why (profiles/r01_ncu_knn_after.md: FMA pipe 29 %, issue-active 51 %, CTA barriers
// behind loading / sorting warps):
//   * candidate and query blocks arrive by TMA (cp.async.bulk.tensor, 128B swizzle) in a 2- or 3-stage mbarrier ring:
//     [128 candidates][32 channels] + [64 queries][32 channels] per stage.  No loader instructions, no transposing STS, the
//     next chunk is in flight while the current one is multiplied.
//   * the tiles stay ROW-major in shared memory; a thread reads the 4 channels of one point per LDS.128 (the swizzle makes
//     the 16 candidate rows of a half-warp hit 8 distinct bank groups) and runs the fmaf chain over them in order.
//   * the finished 4 x 8 distances of a thread never go to shared memory: a half-warp holds one query row's 128 candidates,
//     so admission (ballot per half-warp) and the append into the row buffer happen straight from registers.  The main
//     loop has NO CTA-wide barrier; warps only meet through the ring's empty / full mbarriers.
//   * candidate indices are stored as 16-bit in the row buffers (N < 65536), which keeps two CTAs per SM.
#include "common.cuh"
#define PN_KNN_SORT_ATTR __noinline__
#include "knn_select.cuh"
#include "tc05.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace pn {
namespace knn2 {
using namespace tc05;
using knn::compact_row;
using knn::compact_select;

constexpr int TQ = 64, TC = 128, KC = 32, NT = 256;
constexpr int X_BYTES = TC * KC * 4, Q_BYTES = TQ * KC * 4, STAGE_BYTES = X_BYTES + Q_BYTES;      // 16 KB + 8 KB
constexpr int SAMPLE_M = 1024;
typedef unsigned short BI;

template <int CAP> struct Cfg { static constexpr int NST = (CAP <= 64) ? 3 : 2; };

template <int CAP>
static constexpr size_t smem_bytes() {
    return (size_t)Cfg<CAP>::NST * STAGE_BYTES + (size_t)TQ * CAP * (sizeof(float) + sizeof(BI)) + 3 * TQ * sizeof(float) +
           2 * 4 * sizeof(uint64_t) + 1024;
}

// ld.shared.v4.f32 at a 32-bit shared-window address + compile-time offset (explicit state space: a pointer re-derived
// through integer arithmetic would compile to generic LD.E loads with 64-bit address math)
template <int IMM>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(IMM));
    return v;
}
// a tile written by TMA with the 128-byte swizzle is [rows][128 B] with the 16-byte chunk g (4 channels) of row r stored at
// chunk position g ^ (r & 7)

// full-barrier wait: the common case (data already landed) is one try_wait; the watchdog loop only runs while waiting
__device__ __forceinline__ void wait_ready(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done) mbar_wait_guarded(bar, parity);
}

// out of line on purpose: the compaction's register arrays must not merge into the register allocation of the FMA loop
// (they would spill the accumulators); the call sits in a rarely taken branch
template <int CAP>
__device__ __noinline__ int compact_rare(float* bv, BI* bi, int n, int k, int lane, float* tau_out) {
    return compact_select<CAP, BI>(bv, bi, n, k, lane, tau_out);
}
template <int CAP>
__device__ __noinline__ int sort_row(float* bv, BI* bi, int n, int k, int lane, float* tau_out) {
    return compact_row<CAP, BI>(bv, bi, n, k, lane, tau_out);
}

// state of the two query rows a warp-step touches (lower / upper half-warp), passed and returned BY VALUE so that it stays in
// registers across the out-of-line call
struct RowState { int n_mine, n_other; float tau; };

// SLOW path of an admission step (one of the two rows of the step would overflow its buffer): the whole warp compacts that row,
// then the lanes with `pass` append.  ONE out-of-line copy: inlined into the 32 unrolled steps of a tile the compactions made the
// kernel ~300 KB of SASS and the instruction fetch its top stall (ncu: no_instruction 3.6 per issue).  The common case (both rows
// have room) is a dozen inline instructions in the kernel.
template <int CAP>
__device__ __noinline__ RowState admit(float* bv_lo, BI* bi_lo, int half, int lane, unsigned lt_mask, unsigned m, bool pass,
                                       float d, int j, RowState st, int ksel) {
    __syncwarp();                               // the appends of earlier (inline, unsynchronised) steps are visible to the warp
    const int c_lo = __popc(m & 0xffffu), c_hi = __popc(m >> 16);
    const int n_lo = half ? st.n_other : st.n_mine, n_hi = half ? st.n_mine : st.n_other;
    float* bv_hi = bv_lo + 4 * CAP;             // rows 8 warp + a and 8 warp + 4 + a
    BI* bi_hi = bi_lo + 4 * CAP;
    if (n_lo + c_lo > CAP) {
        float t_new;
        const int n_new = compact_select<CAP, BI>(bv_lo, bi_lo, n_lo, ksel, lane, &t_new);
        if (!half) { st.n_mine = n_new; st.tau = t_new; } else st.n_other = n_new;
    }
    if (n_hi + c_hi > CAP) {
        float t_new;
        const int n_new = compact_select<CAP, BI>(bv_hi, bi_hi, n_hi, ksel, lane, &t_new);
        if (half) { st.n_mine = n_new; st.tau = t_new; } else st.n_other = n_new;
    }
    if (pass) {
        const int pos = st.n_mine + __popc(m & lt_mask);
        (half ? bv_hi : bv_lo)[pos] = d;
        (half ? bi_hi : bi_lo)[pos] = (BI)j;
    }
    st.n_mine += half ? c_hi : c_lo;
    st.n_other += half ? c_lo : c_hi;
    __syncwarp();
    return st;
}

template <int CAP, typename IdxT, bool SAMPLED>
__global__ void __launch_bounds__(NT, 2)
knn_tma_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mQ, const float* __restrict__ xx,
               int N, int C, int k, int r_sample, IdxT* __restrict__ idx_out, float* __restrict__ dist_out) {
    constexpr int NST = Cfg<CAP>::NST;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);          // (offset from the array: stays a shared pointer)
    unsigned char* stages = smem;                                                    // NST x (x tile | q tile)
    const uint32_t stages_u32 = smem_u32(stages);
    float* bufv = reinterpret_cast<float*>(smem + NST * STAGE_BYTES);                // [TQ][CAP]
    BI* bufi = reinterpret_cast<BI*>(bufv + TQ * CAP);                               // [TQ][CAP]
    float* tau_s = reinterpret_cast<float*>(bufi + TQ * CAP);                        // [TQ]
    int* cnt_s = reinterpret_cast<int*>(tau_s + TQ);                                 // [TQ]
    float* xxq = reinterpret_cast<float*>(cnt_s + TQ);                               // [TQ]
    uint64_t* full = reinterpret_cast<uint64_t*>(xxq + TQ);                          // [NST]
    uint64_t* empty = full + 4;                                                      // [NST]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, q0 = blockIdx.x * TQ;
    const float* xxb = xx + (size_t)b * N;
    const int tq = tid >> 4;            // queries 4 tq .. 4 tq + 3   (a warp: tq = 2 warp, 2 warp + 1 -> rows 8 warp .. 8 warp + 7)
    const int tc = tid & 15;            // candidates tc + 16 cc, cc = 0..7
    const int half = lane >> 4;         // which of the warp's two query groups this lane belongs to
    const int kx = tc & 7, kq = 4 * (tq & 1);
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const unsigned lt_mask = ((1u << lane) - 1u) & hmask;         // lanes of my half below me

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NT / 32); }
        mbar_fence_init();
        tma_prefetch_desc(&mX); tma_prefetch_desc(&mQ);
    }
    if (tid < TQ) {
        const int q = q0 + tid;
        xxq[tid] = (q < N) ? xxb[q] : 0.f;
    }
    __syncthreads();

    // per-thread copies of the state of its 4 query rows (identical in the 16 lanes of a half-warp)
    float tau[4];
    int cnt[4], cnt_o[4];               // cnt: my half's rows; cnt_o: the other half's rows (every lane tracks both from the ballots)
#pragma unroll
    for (int a = 0; a < 4; ++a) { tau[a] = -INFINITY; cnt[a] = 0; cnt_o[a] = 0; }
    float xq[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xq[a] = xxq[4 * tq + a];

    const int nchunk = C / KC;
    uint32_t issued = 0, consumed = 0;              // chunk counters over the whole kernel (ring position / parity)
    int jend = N, ksel = k;
    [[maybe_unused]] int phase = 1;
    if constexpr (SAMPLED) { phase = 0; ksel = r_sample; jend = min(N, SAMPLE_M); }

    auto issue = [&](int t, int ci) {               // one elected thread: chunk ci of candidate tile t into the next ring slot
        const int s = issued % NST;
        mbar_wait_guarded(&empty[s], ((issued / NST) & 1) ^ 1);
        unsigned char* st = stages + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        tma_load_3d(st, &mX, &full[s], ci * KC, t * TC, b);
        tma_load_3d(st + X_BYTES, &mQ, &full[s], ci * KC, q0, b);
    };

phase_begin:
    {
        const int ntiles = (jend + TC - 1) / TC;
        const int total = ntiles * nchunk;
        // prologue: fill the ring (issued is only advanced by the producer thread's own bookkeeping, mirrored in all threads)
        int next = 0;
        for (; next < min(NST - 1, total); ++next) {
            if (tid == 0) issue(next / nchunk, next % nchunk);
            ++issued;
        }
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const int j0 = t * TC;
            float acc[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) acc[a][cc] = 0.f;
            float xc8[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const int j = j0 + tc + 16 * cc;
                xc8[cc] = (j < N) ? __ldg(xxb + j) : 0.f;
            }
#pragma unroll 1
            for (int ci = 0; ci < nchunk; ++ci) {
                if (next < total) {                                  // keep the ring full: one chunk ahead per chunk consumed
                    if (tid == 0) issue(next / nchunk, next % nchunk);
                    ++issued; ++next;
                }
                const int s = consumed % NST;
                wait_ready(&full[s], (consumed / NST) & 1);
                // candidate rows tc + 16 cc (swizzle key tc & 7 for all of them), query rows 4 tq + a (key 4 (tq & 1) + a)
                const uint32_t xrow = stages_u32 + s * STAGE_BYTES + tc * 128;
                const uint32_t qrow = stages_u32 + s * STAGE_BYTES + X_BYTES + (4 * tq) * 128;
#pragma unroll 2
                for (int g = 0; g < KC / 4; ++g) {
                    float4 qv[4], xv[8];
                    const uint32_t xa = xrow + (uint32_t)((g ^ kx) << 4);
                    const uint32_t qa = qrow + (uint32_t)((g ^ kq) << 4);          // row a: chunk (g ^ kq) ^ a, kq has low bits 0
                    qv[0] = lds128<0>(qa);
                    qv[1] = lds128<128>(qa ^ 16u);
                    qv[2] = lds128<256>(qa ^ 32u);
                    qv[3] = lds128<384>(qa ^ 48u);
                    xv[0] = lds128<0>(xa);         xv[1] = lds128<2048>(xa);   xv[2] = lds128<4096>(xa);   xv[3] = lds128<6144>(xa);
                    xv[4] = lds128<8192>(xa);      xv[5] = lds128<10240>(xa);  xv[6] = lds128<12288>(xa);  xv[7] = lds128<14336>(xa);
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) {
                            float v = acc[a][cc];
                            v = fmaf(qv[a].x, xv[cc].x, v);
                            v = fmaf(qv[a].y, xv[cc].y, v);
                            v = fmaf(qv[a].z, xv[cc].z, v);
                            v = fmaf(qv[a].w, xv[cc].w, v);
                            acc[a][cc] = v;
                        }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                ++consumed;
            }
            // ---- distances + admission straight from registers: half-warp `half` owns rows 8 warp + 4 half + a
#pragma unroll
            for (int a = 0; a < 4; ++a) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int j = j0 + tc + 16 * cc;
                    const float inner = __fmul_rn(-2.0f, acc[a][cc]);
                    float d = __fsub_rn(__fsub_rn(-xc8[cc], inner), xq[a]);
                    d = (j < N) ? d : -INFINITY;
                    const bool pass = d > tau[a];
                    const unsigned m = __ballot_sync(FULL, pass);
                    if (m) {                                          // warp-uniform
                        const int c_lo = __popc(m & 0xffffu), c_hi = __popc(m >> 16);
                        const int c_mine = half ? c_hi : c_lo, c_other = half ? c_lo : c_hi;
                        if (max(cnt[a] + c_mine, cnt_o[a] + c_other) > CAP) {       // same value in every lane
                            RowState st;
                            st.n_mine = cnt[a]; st.n_other = cnt_o[a]; st.tau = tau[a];
                            st = admit<CAP>(bufv + (8 * warp + a) * CAP, bufi + (8 * warp + a) * CAP, half, lane, lt_mask, m, pass,
                                            d, j, st, ksel);
                            cnt[a] = st.n_mine; cnt_o[a] = st.n_other; tau[a] = st.tau;
                        } else {
                            if (pass) {
                                const int pos = cnt[a] + __popc(m & lt_mask);
                                bufv[(4 * tq + a) * CAP + pos] = d;
                                bufi[(4 * tq + a) * CAP + pos] = (BI)j;
                            }
                            cnt[a] += c_mine; cnt_o[a] += c_other;
                        }
                    }
                }
            }
        }
    }
    // ---- end of a pass: publish the per-row state (the phase logic and the final sort work row by row, warp-cooperatively)
    __syncwarp();
    if ((lane & 15) == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) { tau_s[4 * tq + a] = tau[a]; cnt_s[4 * tq + a] = cnt[a]; }
    }
    __syncwarp();
    if constexpr (SAMPLED) {
        if (phase == 0) {
            // admission threshold of the main pass = r-th best of the sample (-inf when the sample holds fewer than r)
#pragma unroll 1
            for (int r = 0; r < 8; ++r) {
                const int row = warp * 8 + r;
                float t;
                sort_row<CAP>(bufv + row * CAP, bufi + row * CAP, cnt_s[row], ksel, lane, &t);
                if (lane == 0) { tau_s[row] = t; cnt_s[row] = 0; }
            }
            __syncwarp();
#pragma unroll
            for (int a = 0; a < 4; ++a) { tau[a] = tau_s[4 * tq + a]; cnt[a] = 0; cnt_o[a] = 0; }
            phase = 1; ksel = k; jend = N;
            goto phase_begin;
        }
        if (phase == 1) {
            int short_rows = 0;
#pragma unroll 1
            for (int r = 0; r < 8; ++r) short_rows |= (cnt_s[warp * 8 + r] < k) ? 1 : 0;
            if (__syncthreads_or(short_rows)) {             // rare: some row admitted fewer than k -> exact re-run from -inf
#pragma unroll
                for (int a = 0; a < 4; ++a) { tau[a] = -INFINITY; cnt[a] = 0; cnt_o[a] = 0; }
                phase = 2;
                goto phase_begin;
            }
        }
    }
    // ---- final sort + write-out
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        const int row = warp * 8 + r;
        const int q = q0 + row;
        float t;
        float* bv = bufv + row * CAP;
        BI* bi = bufi + row * CAP;
        const int n = sort_row<CAP>(bv, bi, cnt_s[row], k, lane, &t);
        if (q < N) {
            const size_t o = ((size_t)b * N + q) * k;
            for (int p = lane; p < k; p += 32) {
                idx_out[o + p] = (IdxT)((p < n) ? (int)bi[p] : 0);
                if (dist_out) dist_out[o + p] = (p < n) ? bv[p] : -INFINITY;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {       // resolved through the runtime: the library keeps no link dependency on libcuda
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// x [B][N][ld] fp32 (the first C of ld floats of a row are the point): box {32 channels, rows, 1}, 128B swizzle, zero fill
static bool make_map(CUtensorMap* m, const float* x, int B, int N, int C, int ld, uint32_t rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)N * ld * 4};
    cuuint32_t box[3] = {(cuuint32_t)KC, rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sample_rank(int N, int k) {          // same rule as knn.cu
    const char* e = getenv("PN_KNN_SAMPLE");
    if ((e && e[0] == '0') || N < 4 * SAMPLE_M || k < 32) return 0;
    const int r = (int)((3.0 * k * SAMPLE_M + N - 1) / N);
    return (r >= 8 && r < k) ? r : 0;
}

template <int CAP, typename IdxT>
static int launch(const CUtensorMap& mx, const CUtensorMap& mq, const float* xx, int B, int N, int C, int k, void* idx,
                  float* dist, cudaStream_t st) {
    const int r = sample_rank(N, k);
    auto kern = r ? knn_tma_kernel<CAP, IdxT, true> : knn_tma_kernel<CAP, IdxT, false>;
    const size_t sm = smem_bytes<CAP>();
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid(cdiv(N, TQ), B);
    kern<<<grid, NT, sm, st>>>(mx, mq, xx, N, C, k, r, (IdxT*)idx, dist);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn_tma_kernel");
    return PN_OK;
}

// xx[r] = fmaf chain of the squares over the channels in increasing order (same as knn.cu::norms_kernel)
__global__ void norms_kernel(const float* __restrict__ x, long long rows, int ld, int C, float* __restrict__ xx) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* p = x + r * ld;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(p[c], p[c], acc);
    xx[r] = acc;
}

}  // namespace knn2
}  // namespace pn

using namespace pn;

// 1 if pn_knn_tma takes this problem (the caller uses pn_knn otherwise): feature-space metric, whole 32-channel chunks,
// TMA-compatible addressing, 16-bit candidate indices
extern "C" int pn_knn_tma_supported(const float* x, int N, int C, int ld, int k, int metric) {
    if (metric != 0 || C % 32 || C < 32 || ld % 4 || (reinterpret_cast<uintptr_t>(x) & 15u)) return 0;
    if (N >= 65536 || N < 128 || k > 96 || k < 1 || k > N) return 0;
    return 1;
}

extern "C" int pn_knn_tma(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64,
                          float* dist_out, float* ws_norms, void* stream) {
    PN_REQUIRE(x && idx_out && ws_norms, "pn_knn_tma: null pointer");
    PN_REQUIRE(B > 0 && pn_knn_tma_supported(x, N, C, ld, k, metric), "pn_knn_tma: unsupported problem (N=%d C=%d ld=%d k=%d)",
               N, C, ld, k);
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * N;
    knn2::norms_kernel<<<cdiv(rows, 256), 256, 0, st>>>(x, rows, ld, C, ws_norms);
    PN_COUNT_LAUNCH();
    PN_LAUNCH_CHECK("knn norms_kernel");
    CUtensorMap mx, mq;
    if (!knn2::make_map(&mx, x, B, N, C, ld, knn2::TC) || !knn2::make_map(&mq, x, B, N, C, ld, knn2::TQ)) {
        set_error("pn_knn_tma: cuTensorMapEncodeTiled failed or is unavailable");
        return PN_ERR_CUDA;
    }
    if (k <= 32) {
        return idx_is_i64 ? knn2::launch<64, long long>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st)
                          : knn2::launch<64, int>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st);
    }
    return idx_is_i64 ? knn2::launch<128, long long>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st)
                      : knn2::launch<128, int>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st);
}
