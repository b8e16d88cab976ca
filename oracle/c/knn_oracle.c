/*
 * TEST INFRASTRUCTURE ONLY (oracle) — CPU restatement of the reference kNN graph build.
 * Never linked into the product library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call it.
 *
 * Restates (reference paths relative to /root/reference):
 *   knn                 src/PointNet.py:9-26   (dup src/model.py:9-22)
 *   knn_points_normals  src/PointNet.py:29-69
 *
 * The reference ranks, per query row i, the fp32 value
 *   feature metric :  D[i][j] = (-xx[j] - inner[i][j]) - xx[i],  inner = -2 * <x_i, x_j>     (PointNet.py:15-17)
 *   pos+normal     :  D[i][j] = -( ((xx[j] - 2<p_i,p_j>) + xx[i]) * (1 + (2 - 2<n_i,n_j>)) )  (PointNet.py:44-52)
 *   squared diff   :  D[i][j] = -(((dx*dx) + (dy*dy)) + (dz*dz)),  d = x_i - x_j  (metric 2: up_sample_points_torch,
 *                     src/fitting_utils.py:150-163: torch.sum((p_i - p_j) ** 2, 2), topk(5, largest=False)); every square and
 *                     every sum is its own fp32 rounding (no fma), summed in channel order
 * and takes torch.topk(k) along j (largest first) (PointNet.py:22,65).
 *
 * The reference leaves the accumulation order of the dot product to the BLAS it runs on (cuBLAS / MKL), so
 * its near-ties are platform-defined.  This oracle PINS the order so that a CUDA kernel can be bit-exact:
 *   dot  = fmaf chain over channels c = 0..C-1 starting from +0.0f        (same as CUDA __fmaf_rn)
 *   xx   = fmaf chain of squares over c = 0..C-1 starting from +0.0f
 *   every other operation is a single correctly rounded fp32 op in the order written above
 *   ranking: D descending, ties broken by the lower candidate index j.
 * tests/test_oracle_golden.py audits this against the real reference output (golden vectors): every position
 * where the two disagree must be a near-tie of the reference's own distances.
 *
 * Layout: x is point-major [B][N][ld] (ld >= C floats per point row), idx out is [B][N][k] int32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float d; int32_t j; } cand_t;

/* strict ordering: larger distance value first, then lower index */
static inline int better(float da, int32_t ja, float db, int32_t jb) {
    return (da > db) || (da == db && ja < jb);
}

static void sift_down(cand_t *h, int n, int i) {
    /* min-heap on "better": root = worst of the kept k */
    for (;;) {
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && better(h[w].d, h[w].j, h[l].d, h[l].j)) w = l;
        if (r < n && better(h[w].d, h[w].j, h[r].d, h[r].j)) w = r;
        if (w == i) return;
        cand_t t = h[i]; h[i] = h[w]; h[w] = t; i = w;
    }
}

static int cmp_best_first(const void *a, const void *b) {
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (better(x->d, x->j, y->d, y->j)) return -1;
    if (better(y->d, y->j, x->d, x->j)) return 1;
    return 0;
}

static inline float sqdiff3(const float *a, const float *b) {
    float acc = 0.0f;
    for (int c = 0; c < 3; ++c) { float d = a[c] - b[c]; float q = d * d; acc = acc + q; }
    return acc;
}

static inline float dotf(const float *a, const float *b, int c0, int c1) {
    float acc = 0.0f;
    for (int c = c0; c < c1; ++c) acc = fmaf(a[c], b[c], acc);
    return acc;
}

/* metric 0: feature space over channels [0,C); metric 1: positions [0,3) + normals [3,6) */
int pn_oracle_knn(const float *x, int B, int N, int C, int ld, int k, int metric,
                  int32_t *idx_out, float *dist_out /* may be NULL, [B][N][k] */) {
    if (k > N || k <= 0) return 1;
    if (metric == 1 && C != 6) return 2;
    if (metric == 2 && C != 3) return 2;
    int cx = (metric == 1) ? 3 : C;
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (size_t)b * N * ld;
        float *xx = (float *)malloc(sizeof(float) * (size_t)N);
        for (int j = 0; j < N; ++j) xx[j] = dotf(xb + (size_t)j * ld, xb + (size_t)j * ld, 0, cx);
#pragma omp parallel
        {
            cand_t *heap = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
#pragma omp for schedule(dynamic, 16)
            for (int i = 0; i < N; ++i) {
                const float *xi = xb + (size_t)i * ld;
                int n = 0;
                for (int j = 0; j < N; ++j) {
                    const float *xj = xb + (size_t)j * ld;
                    float D;
                    if (metric == 2) {
                        D = -sqdiff3(xi, xj);
                    } else if (metric == 0) {
                        float inner = -2.0f * dotf(xi, xj, 0, C);
                        D = (-xx[j] - inner) - xx[i];
                    } else {
                        float pd = (xx[j] - 2.0f * dotf(xi, xj, 0, 3)) + xx[i];
                        float nd = 2.0f - 2.0f * dotf(xi, xj, 3, 6);
                        D = -(pd * (1.0f + nd));
                    }
                    if (n < k) {
                        heap[n].d = D; heap[n].j = j; ++n;
                        if (n == k) for (int t = k / 2 - 1; t >= 0; --t) sift_down(heap, k, t);
                    } else if (better(D, j, heap[0].d, heap[0].j)) {
                        heap[0].d = D; heap[0].j = j; sift_down(heap, k, 0);
                    }
                }
                qsort(heap, (size_t)k, sizeof(cand_t), cmp_best_first);
                size_t o = ((size_t)b * N + i) * k;
                for (int t = 0; t < k; ++t) {
                    idx_out[o + t] = heap[t].j;
                    if (dist_out) dist_out[o + t] = heap[t].d;
                }
            }
            free(heap);
        }
        free(xx);
    }
    return 0;
}

/* Full fp32 distance row for one query (used by the ambiguity audit in tests). */
int pn_oracle_knn_row(const float *x, int N, int C, int ld, int metric, int i, float *row_out) {
    int cx = (metric == 1) ? 3 : C;
    const float *xi = x + (size_t)i * ld;
    float xxi = dotf(xi, xi, 0, cx);
    for (int j = 0; j < N; ++j) {
        const float *xj = x + (size_t)j * ld;
        float xxj = dotf(xj, xj, 0, cx);
        if (metric == 2) {
            row_out[j] = -sqdiff3(xi, xj);
        } else if (metric == 0) {
            float inner = -2.0f * dotf(xi, xj, 0, C);
            row_out[j] = (-xxj - inner) - xxi;
        } else {
            float pd = (xxj - 2.0f * dotf(xi, xj, 0, 3)) + xxi;
            float nd = 2.0f - 2.0f * dotf(xi, xj, 3, 6);
            row_out[j] = -(pd * (1.0f + nd));
        }
    }
    return 0;
}
