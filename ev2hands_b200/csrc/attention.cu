// Class-wise attention pooling of TEHNet (reference src/Ev2Hands/model/TEHNet.py:9-27, called at :191-192):
//
//     sim[b,c,d]     = scale * sum_n key[b,c,n] * query[b,d,n]          (bmm(key, query^T), scale = D^-1/2)
//     w[b,c,d]       = softmax over the C classes (dim=1)
//     context[b,c,n] = sum_d w[b,c,d] * value[b,d,n]                     (bmm(w, value))
//
// key = the segmentation logits (C = 4 classes), value = the decoder's point features, query = the per-hand
// query convolution's output (D = 256).  Everything is point-major here ([B*N, ld] rows, the layout the tensor-core
// layers leave their outputs in); the result is channel-first [B, C, N], which is what the hand regressor's set
// abstraction consumes (:194-195).  Two small kernels, exact fp32, fixed summation order (deterministic):
//   attn_sim_partial   grid (B, 8): every CTA sums one eighth of the window's points into partial[b, split, c, d]
//   attn_context       grid (B, N / 64): scale + softmax of the summed partials into shared memory, then one warp per
//                      point row: lanes stride over d, four dot products, warp shuffle reduction
// HBM traffic: query and value rows are read once each (2 * N * D * 4 B per window), nothing else of that size.
#include "common.cuh"

namespace ev2h {

constexpr int kAttnSplit = 8;
constexpr int kAttnMaxC = 8;

__global__ void __launch_bounds__(256)
attn_sim_partial_kernel(const float *__restrict__ key, int ld_k, const float *__restrict__ query, int ld_q,
                        int N, int C, int D, float *__restrict__ partial) {
    const int b = blockIdx.x, split = blockIdx.y;
    const int per = (N + kAttnSplit - 1) / kAttnSplit;
    const int n0 = split * per, n1 = min(n0 + per, N);
    __shared__ float ks[64][kAttnMaxC];
    for (int d0 = 0; d0 < D; d0 += blockDim.x) {
        const int d = d0 + threadIdx.x;
        float acc[kAttnMaxC];
#pragma unroll
        for (int c = 0; c < kAttnMaxC; ++c) acc[c] = 0.f;
        for (int t0 = n0; t0 < n1; t0 += 64) {
            const int nt = min(64, n1 - t0);
            __syncthreads();
            for (int i = threadIdx.x; i < nt * C; i += blockDim.x)
                ks[i / C][i % C] = key[((int64_t)b * N + t0 + i / C) * ld_k + i % C];
            __syncthreads();
            if (d < D) {
                for (int i = 0; i < nt; ++i) {
                    const float q = __ldg(query + ((int64_t)b * N + t0 + i) * ld_q + d);
#pragma unroll
                    for (int c = 0; c < kAttnMaxC; ++c)
                        if (c < C) acc[c] = fmaf(ks[i][c], q, acc[c]);
                }
            }
        }
        if (d < D) {
#pragma unroll
            for (int c = 0; c < kAttnMaxC; ++c)
                if (c < C) partial[(((int64_t)b * kAttnSplit + split) * C + c) * D + d] = acc[c];
        }
    }
}

__global__ void __launch_bounds__(256)
attn_context_kernel(const float *__restrict__ partial, const float *__restrict__ value, int ld_v, int N, int C, int D,
                    float scale, float *__restrict__ out_cf) {
    extern __shared__ float w_s[];                 // [C][D] softmax weights of this window
    const int b = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s[kAttnMaxC];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kAttnMaxC; ++c) {
            s[c] = 0.f;
            if (c < C) {
                for (int sp = 0; sp < kAttnSplit; ++sp) s[c] += partial[(((int64_t)b * kAttnSplit + sp) * C + c) * D + d];
                s[c] *= scale;
                mx = fmaxf(mx, s[c]);
            }
        }
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < kAttnMaxC; ++c)
            if (c < C) { s[c] = expf(s[c] - mx); den += s[c]; }
#pragma unroll
        for (int c = 0; c < kAttnMaxC; ++c)
            if (c < C) w_s[c * D + d] = s[c] / den;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows_per_cta = 64;
    for (int r = warp; r < rows_per_cta; r += blockDim.x / 32) {
        const int n = blockIdx.y * rows_per_cta + r;
        if (n >= N) break;
        const float *v = value + ((int64_t)b * N + n) * ld_v;
        float acc[kAttnMaxC];
#pragma unroll
        for (int c = 0; c < kAttnMaxC; ++c) acc[c] = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float x = __ldg(v + d);
#pragma unroll
            for (int c = 0; c < kAttnMaxC; ++c)
                if (c < C) acc[c] = fmaf(w_s[c * D + d], x, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < kAttnMaxC; ++c) {
            if (c < C) {
                float a = acc[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == c) out_cf[((int64_t)b * C + c) * N + n] = a;
            }
        }
    }
}

}  // namespace ev2h

extern "C" int ev2h_class_attention_f32(const float *key_rows, int ld_k, const float *query_rows, int ld_q,
                                        const float *value_rows, int ld_v, int B, int N, int C, int D, float scale,
                                        float *partial, float *out_cf, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(key_rows && query_rows && value_rows && partial && out_cf, "ev2h_class_attention_f32: null argument");
    EV2H_REQUIRE(B > 0 && N > 0 && C > 0 && D > 0, "ev2h_class_attention_f32: bad sizes");
    if (C > kAttnMaxC || D % 32 != 0 || D > 1024)
        return fail(EV2H_ERR_UNSUPPORTED, "ev2h_class_attention_f32: C=%d (<= %d), D=%d (multiple of 32, <= 1024)", C, kAttnMaxC, D);
    EV2H_REQUIRE(ld_k >= C && ld_q >= D && ld_v >= D, "ev2h_class_attention_f32: leading dimensions too small");
    cudaStream_t st = as_stream(stream);
    attn_sim_partial_kernel<<<dim3((unsigned)B, kAttnSplit), 256, 0, st>>>(key_rows, ld_k, query_rows, ld_q, N, C, D, partial);
    int rc = check_launch("ev2h_class_attention_f32 (similarity)");
    if (rc) return rc;
    attn_context_kernel<<<dim3((unsigned)B, (unsigned)((N + 63) / 64)), 256, (size_t)C * D * sizeof(float), st>>>(
        partial, value_rows, ld_v, N, C, D, scale, out_cf);
    return check_launch("ev2h_class_attention_f32 (context)");
}
