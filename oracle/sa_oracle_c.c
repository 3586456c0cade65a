/* TEST INFRASTRUCTURE - plain-C oracle for the index-producing half of the
 * Ev2Hands set-abstraction encoder, plus a double-precision shared-MLP used as
 * an accuracy yardstick.  Never linked into the product library.
 *
 * Parity status: PINNED by tests/test_oracle_cpu.py against the golden vectors
 * that tests/golden/make_golden.py produced from the real reference.
 *
 * The reference (src/Ev2Hands/model/pointnet2_utils.py) is PyTorch; what makes
 * its index outputs reproducible is the exact order of fp32 roundings inside
 * aten/MKL.  Those orders, established by experiment (SURVEY.md section 7 hard
 * parts 1-2 and re-verified by tests/golden/make_golden.py), are restated here
 * with every rounding explicit.  Build with -ffp-contract=off (see Makefile).
 *
 *   orc_fps          <- farthest_point_sample   pointnet2_utils.py:63-84
 *   orc_sqdist       <- square_distance         pointnet2_utils.py:19-40
 *   orc_ball_query   <- query_ball_point        pointnet2_utils.py:87-107
 *   orc_mlp_max_f64  <- conv1x1+BN(eval)+ReLU stack and max over K
 *                                               pointnet2_utils.py:253-257
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* (dx*dx + dy*dy) + dz*dz, products and sums rounded separately:
 * torch.sum((xyz - centroid) ** 2, -1), pointnet2_utils.py:80. */
static float fps_dist(const float *p, const float *c) {
    volatile float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
    volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    volatile float s = xx + yy;
    return s + zz;
}

/* One window.  xyz [N,3]; out [S] int64.  Start index is an input because the
 * reference draws it from torch's CPU generator (pointnet2_utils.py:75). */
void orc_fps(const float *xyz, int64_t N, int64_t S, int64_t start, int64_t *out) {
    float *best = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) best[i] = 1e10f;
    int64_t cur = start;
    for (int64_t s = 0; s < S; ++s) {
        out[s] = cur;
        const float *c = xyz + 3 * cur;
        float top = -1.0f;            /* distances are >= 0, so the first point always wins */
        int64_t arg = 0;
        for (int64_t i = 0; i < N; ++i) {
            float d = fps_dist(xyz + 3 * i, c);
            if (d < best[i]) best[i] = d;          /* mask = dist < distance (:81-82) */
            if (best[i] > top) { top = best[i]; arg = i; }  /* first index of the max (:83) */
        }
        cur = arg;
    }
    free(best);
}

/* |v|^2 as torch.sum(v ** 2, -1): (x*x + y*y) + z*z. */
static float sqnorm(const float *v) {
    volatile float xx = v[0] * v[0], yy = v[1] * v[1], zz = v[2] * v[2];
    volatile float s = xx + yy;
    return s + zz;
}

/* -2 * (q . p) + |q|^2 + |p|^2 with the K=3 dot product as an FMA chain
 * x -> y -> z (what MKL's sgemm does for this shape) and the two adds rounded
 * one after the other (the two in-place += of pointnet2_utils.py:38-39). */
static float expanded_sqdist(const float *q, float qn, const float *p, float pn) {
    volatile float dot = q[0] * p[0];
    dot = fmaf(q[1], p[1], dot);
    dot = fmaf(q[2], p[2], dot);
    volatile float t = -2.0f * dot;
    t = t + qn;
    t = t + pn;
    return t;
}

void orc_sqdist(const float *centres, int64_t S, const float *xyz, int64_t N, float *out) {
    float *pn = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) pn[i] = sqnorm(xyz + 3 * i);
    for (int64_t s = 0; s < S; ++s) {
        float qn = sqnorm(centres + 3 * s);
        for (int64_t i = 0; i < N; ++i)
            out[s * N + i] = expanded_sqdist(centres + 3 * s, qn, xyz + 3 * i, pn[i]);
    }
    free(pn);
}

/* One window.  Keeps a point when NOT (d > r2) - the reference overwrites the
 * index with N where sqrdists > radius**2 (:102), sorts and keeps the first K
 * (:103), then pads with the first hit (:104-106).  aten compares an fp32
 * tensor with a Python scalar in fp32, i.e. against float(radius**2) rounded
 * to nearest - the caller passes exactly that value.  A centre with no hit
 * yields N in every slot, as the reference would. */
void orc_ball_query(const float *xyz, int64_t N, const float *centres, int64_t S,
                    float r2, int64_t K, int64_t *out) {
    float *pn = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) pn[i] = sqnorm(xyz + 3 * i);
    for (int64_t s = 0; s < S; ++s) {
        float qn = sqnorm(centres + 3 * s);
        int64_t n = 0;
        for (int64_t i = 0; i < N && n < K; ++i) {
            float d = expanded_sqdist(centres + 3 * s, qn, xyz + 3 * i, pn[i]);
            if (!(d > r2)) out[s * K + n++] = i;
        }
        int64_t first = n ? out[s * K] : N;
        for (; n < K; ++n) out[s * K + n] = first;
    }
    free(pn);
}

/* Shared MLP + max over the K rows of each group, all arithmetic in double.
 *   x        [G, K, C0]  grouped input rows (already gathered / centred)
 *   n_layers layers; layer l maps dims[l] -> dims[l+1]
 *   w[l] [dims[l+1], dims[l]], b/gamma/beta/mean/var[l] [dims[l+1]]
 *   out      [G, dims[n_layers]]
 * y = relu(gamma * (W x + b - mean) / sqrt(var + eps) + beta) per layer. */
void orc_mlp_max_f64(const float *x, int64_t G, int64_t K, int n_layers, const int64_t *dims,
                     const float *const *w, const float *const *b, const float *const *gamma,
                     const float *const *beta, const float *const *mean, const float *const *var,
                     double eps, double *out) {
    int64_t widest = 0;
    for (int l = 0; l <= n_layers; ++l) if (dims[l] > widest) widest = dims[l];
    double *cur = (double *)malloc(sizeof(double) * (size_t)widest);
    double *nxt = (double *)malloc(sizeof(double) * (size_t)widest);
    int64_t cout = dims[n_layers];
    for (int64_t g = 0; g < G; ++g) {
        double *o = out + g * cout;
        for (int64_t c = 0; c < cout; ++c) o[c] = -INFINITY;
        for (int64_t k = 0; k < K; ++k) {
            const float *row = x + (g * K + k) * dims[0];
            for (int64_t c = 0; c < dims[0]; ++c) cur[c] = row[c];
            for (int l = 0; l < n_layers; ++l) {
                int64_t ci = dims[l], co = dims[l + 1];
                for (int64_t j = 0; j < co; ++j) {
                    double acc = b[l][j];
                    const float *wr = w[l] + j * ci;
                    for (int64_t c = 0; c < ci; ++c) acc += (double)wr[c] * cur[c];
                    acc = (acc - mean[l][j]) / sqrt((double)var[l][j] + eps) * gamma[l][j] + beta[l][j];
                    nxt[j] = acc > 0.0 ? acc : 0.0;
                }
                double *t = cur; cur = nxt; nxt = t;
            }
            for (int64_t c = 0; c < cout; ++c) if (cur[c] > o[c]) o[c] = cur[c];
        }
    }
    free(cur);
    free(nxt);
}
