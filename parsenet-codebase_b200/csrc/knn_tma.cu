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

// shared-memory plan of one CTA (host and device agree through these two functions)
//   QRES: the CTA's 64 query rows stay resident for the whole kernel ([C/32][64][128 B], loaded once) and a ring stage holds the
//         candidate tile only (16 KB); otherwise a stage holds candidate + query chunk (24 KB).
template <int CAP>
static inline size_t fixed_bytes() {
    return (size_t)TQ * CAP * (sizeof(float) + sizeof(BI)) + 3 * TQ * sizeof(float) + 12 * sizeof(uint64_t);
}
static inline int stage_bytes(bool qres) { return qres ? X_BYTES : STAGE_BYTES; }

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

// out of line on purpose: the sort's register arrays must not merge into the register allocation of the FMA loop
template <int CAP>
__device__ __noinline__ int sort_row(float* bv, BI* bi, int n, int k, int lane, float* tau_out) {
    return compact_row<CAP, BI>(bv, bi, n, k, lane, tau_out);
}
// k-th best value of a row buffer without ordering it (quickselect; ~4x fewer instructions than the bitonic sort)
template <int CAP>
__device__ __noinline__ float kth_best(float* bv, BI* bi, int n, int k, int lane) {
    float t;
    compact_select<CAP, BI>(bv, bi, n, k, lane, &t);
    return t;
}

// state of the two query rows a warp-step touches (lower / upper half-warp), passed and returned BY VALUE so that it stays in
// registers across the out-of-line call
struct RowState { int n_mine, n_other; float tau; };

// SLOW path of the admission (one of the two rows of the step would overflow its buffer), one 16-candidate group at a time: the
// whole warp compacts the row that is full, then the lanes with `pass` append.  ONE out-of-line copy: inlined into the unrolled
// steps of a tile the compactions made the kernel ~300 KB of SASS and instruction fetch its top stall (ncu: no_instruction 3.6
// per issue).  The common case (both rows have room for everything the tile admits) is the inline scan in the kernel.
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

template <int CAP, typename IdxT, bool SAMPLED, bool QRES>
__global__ void __launch_bounds__(NT, 2)
knn_tma_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mQ, const float* __restrict__ xx,
               int N, int C, int k, int r_sample, int nst, IdxT* __restrict__ idx_out, float* __restrict__ dist_out,
               const int* __restrict__ flags) {
    // flags (optional, [B][N]): a CTA none of whose 64 query rows is flagged exits at once (exact fall-back of knn_tc.cu)
    if (flags) {
        const int q = blockIdx.x * TQ + threadIdx.x;
        const int f = (threadIdx.x < TQ && q < N) ? flags[(size_t)blockIdx.y * N + q] : 0;
        if (!__syncthreads_or(f)) return;
    }
    // The kernel has no static shared memory, so the dynamic window starts at offset 0 of the CTA's allocation, which is 1024-byte
    // aligned (what the 128B-swizzled TMA tiles need); the plan has no slack for padding (two CTAs per SM), hence the trap.
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    constexpr int SB = QRES ? X_BYTES : STAGE_BYTES;
    const int nchunk = C / KC;
    unsigned char* stages = smem;                                                    // nst x stage
    unsigned char* qres = smem + nst * SB;                                           // QRES: [nchunk][64][128 B]
    float* bufv = reinterpret_cast<float*>(qres + (QRES ? nchunk * Q_BYTES : 0));    // [TQ][CAP]
    BI* bufi = reinterpret_cast<BI*>(bufv + TQ * CAP);                               // [TQ][CAP]
    float* tau_s = reinterpret_cast<float*>(bufi + TQ * CAP);                        // [TQ]
    int* cnt_s = reinterpret_cast<int*>(tau_s + TQ);                                 // [TQ]
    float* xxq = reinterpret_cast<float*>(cnt_s + TQ);                               // [TQ]
    uint64_t* full = reinterpret_cast<uint64_t*>(xxq + TQ);                          // [4]
    uint64_t* empty = full + 4;                                                      // [4]
    uint64_t* qbar = full + 8;                                                       // [1]
    const uint32_t stages_u32 = smem_u32(stages);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, q0 = blockIdx.x * TQ;
    const float* xxb = xx + (size_t)b * N;
    const int tq = tid >> 4;            // queries 4 tq .. 4 tq + 3   (a warp: tq = 2 warp, 2 warp + 1 -> rows 8 warp .. 8 warp + 7)
    const int tc = tid & 15;            // candidates tc + 16 cc, cc = 0..7
    const int half = lane >> 4;         // which of the warp's two query groups this lane belongs to
    const int hl = lane & 15;           // lane inside the half-warp
    const int kx = tc & 7, kq = 4 * (tq & 1);
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const unsigned lt_mask = ((1u << lane) - 1u) & hmask;         // lanes of my half below me

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NT / 32); }
        mbar_init(qbar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&mX); tma_prefetch_desc(&mQ);
    }
    if (tid < TQ) {
        const int q = q0 + tid;
        xxq[tid] = (q < N) ? xxb[q] : 0.f;
    }
    __syncthreads();
    if constexpr (QRES) {
        if (tid == 0) {
            mbar_arrive_expect_tx(qbar, nchunk * Q_BYTES);
            for (int ci = 0; ci < nchunk; ++ci) tma_load_3d(qres + ci * Q_BYTES, &mQ, qbar, ci * KC, q0, b);
        }
    }

    // per-thread copies of the state of its 4 query rows (identical in the 16 lanes of a half-warp)
    float tau[4];
    int cnt[4], cnt_o[4];               // cnt: my half's rows; cnt_o: the other half's rows (every lane tracks both)
#pragma unroll
    for (int a = 0; a < 4; ++a) { tau[a] = -INFINITY; cnt[a] = 0; cnt_o[a] = 0; }
    float xq[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xq[a] = xxq[4 * tq + a];

    uint32_t issued = 0, consumed = 0;              // chunk counters over the whole kernel
    uint32_t is = 0, ip = 0, cs = 0, cp = 0;         // ring slot / phase bit of the next chunk to issue resp. consume
    int jend = N, ksel = k;
    [[maybe_unused]] int phase = 1;
    if constexpr (SAMPLED) { phase = 0; ksel = r_sample; jend = min(N, SAMPLE_M); }

    auto issue = [&](int t, int ci) {               // one elected thread: chunk ci of candidate tile t into the next ring slot
        mbar_wait_guarded(&empty[is], ip ^ 1);
        unsigned char* st = stages + is * SB;
        mbar_arrive_expect_tx(&full[is], SB);
        tma_load_3d(st, &mX, &full[is], ci * KC, t * TC, b);
        if constexpr (!QRES) tma_load_3d(st + X_BYTES, &mQ, &full[is], ci * KC, q0, b);
    };
    auto issued_step = [&]() { ++issued; if (++is == (uint32_t)nst) { is = 0; ip ^= 1; } };
    if constexpr (QRES) wait_ready(qbar, 0);

phase_begin:
    {
        const int ntiles = (jend + TC - 1) / TC;
        const int total = ntiles * nchunk;
        int next = 0;                                // chunks of this pass issued so far
        for (; next < min(nst - 1, total); ++next) {
            if (tid == 0) issue(next / nchunk, next % nchunk);
            issued_step();
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
                    issued_step(); ++next;
                }
                wait_ready(&full[cs], cp);
                // candidate rows tc + 16 cc (swizzle key tc & 7 for all of them), query rows 4 tq + a (key 4 (tq & 1) + a)
                const uint32_t xrow = stages_u32 + cs * SB + tc * 128;
                const uint32_t qrow = (QRES ? stages_u32 + nst * SB + ci * Q_BYTES : stages_u32 + cs * SB + X_BYTES) + (4 * tq) * 128;
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
                if (lane == 0) mbar_arrive(&empty[cs]);
                ++consumed;
                if (++cs == (uint32_t)nst) { cs = 0; cp ^= 1; }
            }
            // ---- distances + admission straight from registers: half-warp `half` owns rows 8 warp + 4 half + a.
            // Per row: every lane turns its 8 distances into a pass mask, a 16-lane prefix sum gives each lane its slots in the
            // row buffer, and the (d, j) pairs are stored -- one ballot and one scan per row instead of one ballot per 16
            // candidates.  A row that would overflow takes the out-of-line group-by-group path (compaction).
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float d8[8];
                unsigned pm = 0;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int j = j0 + tc + 16 * cc;
                    const float inner = __fmul_rn(-2.0f, acc[a][cc]);
                    float d = __fsub_rn(__fsub_rn(-xc8[cc], inner), xq[a]);
                    d = (j < N) ? d : -INFINITY;
                    d8[cc] = d;
                    pm |= (d > tau[a]) ? (1u << cc) : 0u;
                }
                if (__ballot_sync(FULL, pm != 0) == 0) continue;                    // nothing admitted in either row
                const int c = __popc(pm);
                int incl = c;                                                        // inclusive prefix sum inside the half-warp
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (hl >= o) incl += v;
                }
                const int tot_mine = __shfl_sync(FULL, incl, 15 + 16 * half);
                const int tot_other = __shfl_sync(FULL, incl, 15 + 16 * (half ^ 1));
                if (max(cnt[a] + tot_mine, cnt_o[a] + tot_other) <= CAP) {           // same value in every lane
                    float* bv = bufv + (4 * tq + a) * CAP + cnt[a] + (incl - c);
                    BI* bi = bufi + (4 * tq + a) * CAP + cnt[a] + (incl - c);
                    int w = 0;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        if (pm & (1u << cc)) { bv[w] = d8[cc]; bi[w] = (BI)(j0 + tc + 16 * cc); ++w; }
                    }
                    cnt[a] += tot_mine; cnt_o[a] += tot_other;
                } else {
#pragma unroll 1
                    for (int cc = 0; cc < 8; ++cc) {
                        float d = d8[0];                                               // d8[cc] without dynamic register indexing
#pragma unroll
                        for (int u = 1; u < 8; ++u) d = (cc == u) ? d8[u] : d;
                        const bool pass = d > tau[a];                                  // (tau may have risen in this loop)
                        const unsigned m = __ballot_sync(FULL, pass);
                        if (m) {
                            RowState st;
                            st.n_mine = cnt[a]; st.n_other = cnt_o[a]; st.tau = tau[a];
                            st = admit<CAP>(bufv + (8 * warp + a) * CAP, bufi + (8 * warp + a) * CAP, half, lane, lt_mask, m, pass,
                                            d, j0 + tc + 16 * cc, st, ksel);
                            cnt[a] = st.n_mine; cnt_o[a] = st.n_other; tau[a] = st.tau;
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
                const float t = kth_best<CAP>(bufv + row * CAP, bufi + row * CAP, cnt_s[row], ksel, lane);
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
                  float* dist, cudaStream_t st, const int* flags = nullptr) {
    const int r = sample_rank(N, k);
    // shared-memory plan for two CTAs per SM (<= 113 KB each): resident queries + 16 KB candidate stages when they fit with
    // at least 3 stages, else 24 KB (candidate + query) stages; as many stages as fit, at most 4
    const size_t budget = 113 * 1024, fixed = fixed_bytes<CAP>();          // (228 KB per SM - 1 KB reserved per CTA) / 2
    const size_t qres_bytes = (size_t)(C / KC) * Q_BYTES;
    bool qres = fixed + qres_bytes + 3 * (size_t)X_BYTES <= budget;
    int nst = (int)((budget - fixed - (qres ? qres_bytes : 0)) / stage_bytes(qres));
    nst = nst > 4 ? 4 : nst;
    PN_REQUIRE(nst >= 2, "pn_knn_tma: shared-memory plan does not fit (C=%d)", C);
    const size_t sm = fixed + (qres ? qres_bytes : 0) + (size_t)nst * stage_bytes(qres);
    void (*kern)(CUtensorMap, CUtensorMap, const float*, int, int, int, int, int, IdxT*, float*, const int*) =
        r ? (qres ? knn_tma_kernel<CAP, IdxT, true, true> : knn_tma_kernel<CAP, IdxT, true, false>)
          : (qres ? knn_tma_kernel<CAP, IdxT, false, true> : knn_tma_kernel<CAP, IdxT, false, false>);
    PN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid(cdiv(N, TQ), B);
    kern<<<grid, NT, sm, st>>>(mx, mq, xx, N, C, k, r, nst, (IdxT*)idx, dist, flags);
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

static int knn_tma_run(const char* who, const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out,
                       int idx_is_i64, float* dist_out, float* ws_norms, const int* flags, void* stream) {
    PN_REQUIRE(x && idx_out && ws_norms, "%s: null pointer", who);
    PN_REQUIRE(B > 0 && pn_knn_tma_supported(x, N, C, ld, k, metric), "%s: unsupported problem (N=%d C=%d ld=%d k=%d)", who,
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
        return idx_is_i64 ? knn2::launch<64, long long>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st, flags)
                          : knn2::launch<64, int>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st, flags);
    }
    return idx_is_i64 ? knn2::launch<128, long long>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st, flags)
                      : knn2::launch<128, int>(mx, mq, ws_norms, B, N, C, k, idx_out, dist_out, st, flags);
}

extern "C" int pn_knn_tma(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64,
                          float* dist_out, float* ws_norms, void* stream) {
    return knn_tma_run("pn_knn_tma", x, B, N, C, ld, k, metric, idx_out, idx_is_i64, dist_out, ws_norms, nullptr, stream);
}

// the same graph for the 64-row tiles that contain a row with flags[b][row] != 0 only (other tiles exit at once, their rows of
// idx_out / dist_out stay untouched): exact fall-back of pn_knn_tc
extern "C" int pn_knn_tma_flagged(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64,
                                  float* dist_out, float* ws_norms, const int* flags, void* stream) {
    PN_REQUIRE(flags, "pn_knn_tma_flagged: null flags");
    return knn_tma_run("pn_knn_tma_flagged", x, B, N, C, ld, k, metric, idx_out, idx_is_i64, dist_out, ws_norms, flags, stream);
}
