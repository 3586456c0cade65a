// Layout and grouping kernels around the shared MLP: neighbourhood gather and its
// backward, batched transpose, BatchNorm folding, max-pool over the neighbour
// axis with argmax and its backward.  All are streaming, HBM/L2-bound kernels:
// one element per thread, consecutive threads on consecutive addresses of the
// side that is written (reads of gathered rows are contiguous per row).
#include "common.cuh"

namespace ev2h {

// ---- gather: rows (b,s,j) = [feats(p) | xyz(p) - centre(s) | 0] -----------------
// reference: index_points + subtract + cat, pointnet2_utils.py:244-248
__global__ void __launch_bounds__(256)
group_gather_kernel(const float *__restrict__ xyz, int64_t sb, int64_t sc, int64_t sn,
                    const float *__restrict__ feats, int D, const float *__restrict__ centres,
                    const int32_t *__restrict__ idx, int idx_ld, int k_off, int N, int S, int K,
                    float *__restrict__ out, int ld_out, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % ld_out);
        const int64_t r = e / ld_out;            // row = (b*S + s)*K + j
        const int64_t bs = r / K;
        const int j = (int)(r - bs * K);
        const int64_t b = bs / S;
        float v = 0.f;
        if (c < D + 3) {
            const int p = idx[bs * idx_ld + k_off + j];
            // p == N is the ball query's "no neighbour" sentinel; the reference would raise
            // an IndexError there, we keep the row at zero instead of faulting.
            if (p < 0 || p >= N) {
                v = 0.f;
            } else if (c < D) {
                v = feats[(b * N + p) * (int64_t)D + c];
            } else {
                const int a = c - D;
                // grouped_xyz -= new_xyz  (:245), a plain fp32 subtraction
                v = __fsub_rn(xyz[b * sb + a * sc + (int64_t)p * sn], centres[bs * 3 + a]);
            }
        }
        out[e] = v;
    }
}

// reference: autograd of index_points (scatter-add of the gathered rows' gradient)
__global__ void __launch_bounds__(256)
group_gather_bwd_kernel(const float *__restrict__ grad_rows, int ld_grad, const int32_t *__restrict__ idx,
                        int idx_ld, int k_off, int N, int S, int K, int D,
                        float *__restrict__ grad_feats, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % D);
        const int64_t r = e / D;
        const int64_t bs = r / K;
        const int j = (int)(r - bs * K);
        const int64_t b = bs / S;
        const int p = idx[bs * idx_ld + k_off + j];
        if (p < 0 || p >= N) continue;
        atomicAdd(grad_feats + (b * N + p) * (int64_t)D + c, grad_rows[r * ld_grad + c]);
    }
}

// ---- batched transpose through a 32x33 shared tile -------------------------------
__global__ void __launch_bounds__(256)
transpose_kernel(const float *__restrict__ src, int64_t ssb, int64_t ssr, int64_t ssc, int R, int C,
                 float *__restrict__ dst, int64_t dsb, int64_t dld, int64_t dcol) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const float *s = src + (int64_t)b * ssb;
    float *d = dst + (int64_t)b * dsb;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        if (r < R && c < C) tile[ty + i][tx] = s[(int64_t)r * ssr + (int64_t)c * ssc];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (r < R && c < C) d[(int64_t)c * dld + dcol + r] = tile[tx][ty + i];
    }
}

// ---- conv bias + eval BatchNorm -> one affine map, written input-channel major ----
__global__ void __launch_bounds__(256)
fold_kernel(const float *__restrict__ w, const float *__restrict__ cb, const float *__restrict__ gamma,
            const float *__restrict__ beta, const float *__restrict__ mean, const float *__restrict__ var,
            double eps, int Cin, int Cout, int Cin_pad, int Cout_pad, float *__restrict__ wt,
            float *__restrict__ bias_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Cin_pad * Cout_pad) return;
    const int k = e / Cout_pad, n = e % Cout_pad;
    float v = 0.f;
    double scale = 0.0;
    if (n < Cout) scale = (double)gamma[n] / sqrt((double)var[n] + eps);
    if (n < Cout && k < Cin) v = (float)(scale * (double)w[(int64_t)n * Cin + k]);
    wt[e] = v;
    if (k == 0) bias_out[n] = n < Cout ? (float)(scale * ((double)cb[n] - (double)mean[n]) + (double)beta[n]) : 0.f;
}

// ---- max over K of channel-first [B,C,K,S] with first-index argmax -----------------
// reference: torch.max(new_points, 2)[0], pointnet2_utils.py:199, :257
__global__ void __launch_bounds__(256)
group_max_kernel(const float *__restrict__ x, int K, int S, float *__restrict__ out,
                 int32_t *__restrict__ arg, int64_t total) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (b*C + c)*S + s
    if (e >= total) return;
    const int64_t bc = e / S;
    const int s = (int)(e - bc * S);
    const float *p = x + bc * K * (int64_t)S + s;
    float m = p[0];
    int a = 0;
    for (int k = 1; k < K; ++k) {
        const float v = p[(int64_t)k * S];
        if (v > m || (v != v && m == m)) { m = v; a = k; }   // NaN propagates like torch.max
    }
    out[e] = m;
    if (arg) arg[e] = a;
}

__global__ void __launch_bounds__(256)
group_max_bwd_kernel(const float *__restrict__ go, const int32_t *__restrict__ arg, int K, int S,
                     float *__restrict__ gx, int64_t total) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over [B,C,K,S]
    if (e >= total) return;
    const int s = (int)(e % S);
    const int64_t t = e / S;
    const int k = (int)(t % K);
    const int64_t bc = t / K;
    const int64_t o = bc * S + s;
    gx[e] = arg[o] == k ? go[o] : 0.f;
}

static int grid_for(int64_t total, int block) {
    int64_t g = (total + block - 1) / block;
    const int64_t cap = 148 * 32;          // grid-stride beyond 32 CTAs per SM
    return (int)(g < cap ? g : cap);
}

// 32-byte point records [features (D <= 5) | x y z | 0 ...] for the gather mode of the fused kernel: one thread per point,
// the channel planes are read coalesced and a record leaves as two 16-byte stores (a full sector).  Two batched
// transposes into a zero-filled buffer did the same with three passes of partial-sector writes (0.34 ms for 256 windows
// of 16384 points against 0.07 ms).
__global__ void __launch_bounds__(256)
point_records_kernel(const float *__restrict__ feats, int64_t fb, int64_t fc, int64_t fn, int D,
                     const float *__restrict__ xyz, int64_t xb, int64_t xc, int64_t xn, int N, float4 *__restrict__ out, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {   // grid_for caps the grid
        const int64_t b = e / N, n = e % N;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = 0.f;
#pragma unroll
        for (int c = 0; c < 5; ++c)
            if (c < D) v[c] = feats[b * fb + c * fc + n * fn];
        const float x = xyz[b * xb + n * xn], y = xyz[b * xb + xc + n * xn], z = xyz[b * xb + 2 * xc + n * xn];
#pragma unroll
        for (int c = 0; c < 6; ++c) {            // the coordinates follow the D features
            if (c == D) { v[c] = x; v[c + 1] = y; v[c + 2] = z; }
        }
        out[2 * e] = make_float4(v[0], v[1], v[2], v[3]);
        out[2 * e + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

}  // namespace ev2h

extern "C" int ev2h_group_gather_f32(const float *xyz, int64_t stride_b, int64_t stride_c, int64_t stride_n,
                                     const float *feats_rows, int D, const float *centres_rows,
                                     const int32_t *idx, int idx_ld, int k_off, int B, int N, int S, int K,
                                     float *out, int ld_out, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(xyz && centres_rows && idx && out, "ev2h_group_gather_f32: null argument");
    EV2H_REQUIRE(D >= 0 && (D == 0 || feats_rows), "ev2h_group_gather_f32: D=%d needs feats_rows", D);
    EV2H_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0, "ev2h_group_gather_f32: sizes must be positive");
    EV2H_REQUIRE(ld_out >= D + 3, "ev2h_group_gather_f32: ld_out=%d < D+3=%d", ld_out, D + 3);
    EV2H_REQUIRE(k_off >= 0 && k_off + K <= idx_ld, "ev2h_group_gather_f32: k_off+K exceeds idx_ld");
    const int64_t total = (int64_t)B * S * K * ld_out;
    group_gather_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
        xyz, stride_b, stride_c, stride_n, feats_rows, D, centres_rows, idx, idx_ld, k_off, N, S, K, out, ld_out, total);
    return check_launch("ev2h_group_gather_f32");
}

extern "C" int ev2h_group_gather_bwd_f32(const float *grad_rows, int ld_grad, const int32_t *idx, int idx_ld,
                                         int k_off, int B, int N, int S, int K, int D, float *grad_feats_rows,
                                         ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(grad_rows && idx && grad_feats_rows, "ev2h_group_gather_bwd_f32: null argument");
    EV2H_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0 && D > 0 && ld_grad >= D, "ev2h_group_gather_bwd_f32: bad sizes");
    const int64_t total = (int64_t)B * S * K * D;
    group_gather_bwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
        grad_rows, ld_grad, idx, idx_ld, k_off, N, S, K, D, grad_feats_rows, total);
    return check_launch("ev2h_group_gather_bwd_f32");
}

extern "C" int ev2h_transpose_f32(const float *src, int64_t src_stride_b, int64_t src_stride_r, int64_t src_stride_c,
                                  int B, int R, int C, float *dst, int64_t dst_stride_b, int64_t dst_ld,
                                  int64_t dst_col_off, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(src && dst, "ev2h_transpose_f32: null argument");
    EV2H_REQUIRE(B > 0 && R > 0 && C > 0 && B <= 65535, "ev2h_transpose_f32: bad sizes");
    EV2H_REQUIRE(dst_ld >= dst_col_off + R, "ev2h_transpose_f32: dst_ld too small");
    dim3 grid((C + 31) / 32, (R + 31) / 32, B);
    EV2H_REQUIRE(grid.y <= 65535, "ev2h_transpose_f32: R too large");
    transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, src_stride_b, src_stride_r, src_stride_c, R, C, dst,
                                                           dst_stride_b, dst_ld, dst_col_off);
    return check_launch("ev2h_transpose_f32");
}

extern "C" int ev2h_fold_conv_bn_f32(const float *conv_w, const float *conv_b, const float *bn_gamma,
                                     const float *bn_beta, const float *bn_mean, const float *bn_var, double eps,
                                     int Cin, int Cout, float *wt, float *bias_out, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(conv_w && conv_b && bn_gamma && bn_beta && bn_mean && bn_var && wt && bias_out,
                 "ev2h_fold_conv_bn_f32: null argument");
    EV2H_REQUIRE(Cin > 0 && Cout > 0, "ev2h_fold_conv_bn_f32: bad sizes");
    const int cin_pad = round_up(Cin, 16), cout_pad = round_up(Cout, 128);
    const int total = cin_pad * cout_pad;
    fold_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(conv_w, conv_b, bn_gamma, bn_beta, bn_mean, bn_var,
                                                                     eps, Cin, Cout, cin_pad, cout_pad, wt, bias_out);
    return check_launch("ev2h_fold_conv_bn_f32");
}

extern "C" int ev2h_group_max_f32(const float *x, int B, int C, int K, int S, float *out, int32_t *arg,
                                  ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(x && out, "ev2h_group_max_f32: null argument");
    EV2H_REQUIRE(B > 0 && C > 0 && K > 0 && S > 0, "ev2h_group_max_f32: bad sizes");
    const int64_t total = (int64_t)B * C * S;
    group_max_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, K, S, out, arg, total);
    return check_launch("ev2h_group_max_f32");
}

extern "C" int ev2h_group_max_bwd_f32(const float *grad_out, const int32_t *arg, int B, int C, int K, int S,
                                      float *grad_x, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(grad_out && arg && grad_x, "ev2h_group_max_bwd_f32: null argument");
    EV2H_REQUIRE(B > 0 && C > 0 && K > 0 && S > 0, "ev2h_group_max_bwd_f32: bad sizes");
    const int64_t total = (int64_t)B * C * K * S;
    group_max_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(grad_out, arg, K, S, grad_x, total);
    return check_launch("ev2h_group_max_bwd_f32");
}

/* pts8[b, n, :] = [feats[b, 0..D, n] | xyz[b, 0..3, n] | 0 ...]: the 32-byte point records the gather mode of
 * ev2h_sa_msg_fused_tc reads (D + 3 <= 8; feats may be NULL with D = 0).  Both inputs are channel-first with element strides. */
extern "C" int ev2h_point_records_f32(const float *feats, int64_t feats_stride_b, int64_t feats_stride_c, int64_t feats_stride_n, int D,
                                      const float *xyz, int64_t stride_b, int64_t stride_c, int64_t stride_n,
                                      int B, int N, float *pts8, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(xyz && pts8 && ((uintptr_t)pts8 & 15) == 0, "ev2h_point_records_f32: null or misaligned argument");
    EV2H_REQUIRE(B > 0 && N > 0 && D >= 0 && D <= 5 && (D == 0 || feats), "ev2h_point_records_f32: bad sizes (D + 3 must fit 8 floats)");
    const int64_t total = (int64_t)B * N;
    point_records_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(feats, feats_stride_b, feats_stride_c, feats_stride_n, D,
                                                                              xyz, stride_b, stride_c, stride_n, N,
                                                                              reinterpret_cast<float4 *>(pts8), total);
    return check_launch("ev2h_point_records_f32");
}
