// Farthest point sampling: one CTA per event window.
//
// Replaces farthest_point_sample (reference src/Ev2Hands/model/pointnet2_utils.py:63-84).
// The whole window stays on chip: coordinates and the running min-distance live
// in registers (P points per thread), a copy of the coordinates in shared memory
// serves the broadcast read of the newly chosen centre.  One iteration is
//   distance update (non-fused mul/add, the reference's rounding order)
//   -> per-thread best -> warp REDUX max / min -> one __syncthreads -> REDUX again,
// so the S dependent iterations cost one block barrier each.
//
// Tie-break: the reference's torch.max returns the FIRST index of the maximum.
// Distances are >= +0, so their bit patterns order like unsigned integers; the
// winner is found as max over value bits, then min over the indices that hold it.
#include "common.cuh"

namespace ev2h {

constexpr unsigned kFull = 0xffffffffu;

template <int T, int P, bool XYZ_IN_REGS>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float *__restrict__ xyz, int64_t sb, int64_t sc, int64_t sn,
           const int64_t *__restrict__ start, int N, int S,
           int32_t *__restrict__ out_idx, float *__restrict__ out_rows, float *__restrict__ out_cf) {
    constexpr int NP = T * P;
    constexpr int W = T / 32;
    extern __shared__ float fps_smem[];
    float *sx = fps_smem, *sy = fps_smem + NP, *sz = fps_smem + 2 * NP;
    __shared__ unsigned long long slot[2][32];

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *base = xyz + (int64_t)b * sb;

    float x[XYZ_IN_REGS ? P : 1], y[XYZ_IN_REGS ? P : 1], z[XYZ_IN_REGS ? P : 1];
    float best[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = tid + j * T;
        float px = 0.f, py = 0.f, pz = 0.f;
        // Padding slots (i >= N) sit at the origin with best = 0: they can never
        // beat a real point (real values are >= 0 and real indices are smaller).
        best[j] = 0.f;
        if (i < N) {
            px = base[(int64_t)i * sn];
            py = base[sc + (int64_t)i * sn];
            pz = base[2 * sc + (int64_t)i * sn];
            best[j] = 1e10f;   // torch.ones(B, N) * 1e10, pointnet2_utils.py:74
        }
        sx[i] = px; sy[i] = py; sz[i] = pz;
        if (XYZ_IN_REGS) { x[j] = px; y[j] = py; z[j] = pz; }
    }
    // device-resident start indices cannot be validated on the host without a sync: clamp (the host binding
    // range-checks host tensors and raises like the reference's indexing would, pointnet2_utils.py:77)
    const int64_t st0 = start[b];
    int cur = st0 < 0 ? 0 : (st0 >= N ? N - 1 : (int)st0);
    __syncthreads();

    for (int s = 0; s < S; ++s) {
        const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
        if (tid == 0) {
            if (out_idx) out_idx[(int64_t)b * S + s] = cur;
            if (out_rows) {
                float *o = out_rows + ((int64_t)b * S + s) * 3;
                o[0] = cx; o[1] = cy; o[2] = cz;
            }
            if (out_cf) {
                float *o = out_cf + (int64_t)b * 3 * S + s;
                o[0] = cx; o[S] = cy; o[2 * (int64_t)S] = cz;
            }
        }
        unsigned bv = 0u, bi = 0u;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            float px, py, pz;
            if (XYZ_IN_REGS) { px = x[j]; py = y[j]; pz = z[j]; }
            else { const int i = tid + j * T; px = sx[i]; py = sy[i]; pz = sz[i]; }
            const float dx = __fsub_rn(px, cx), dy = __fsub_rn(py, cy), dz = __fsub_rn(pz, cz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float m = d < best[j] ? d : best[j];   // distance[mask] = dist[mask], :81-82
            best[j] = m;
            const unsigned vb = __float_as_uint(m);
            if (j == 0 || vb > bv) { bv = vb; bi = (unsigned)(tid + j * T); }   // ascending j == ascending index
        }
        unsigned m = __reduce_max_sync(kFull, bv);
        unsigned wi = __reduce_min_sync(kFull, bv == m ? bi : 0xffffffffu);
        if (W > 1) {
            if (lane == 0) slot[s & 1][warp] = ((unsigned long long)m << 32) | wi;
            __syncthreads();
            unsigned em = 0u, ei = 0xffffffffu;
            if (lane < W) {
                const unsigned long long e = slot[s & 1][lane];
                em = (unsigned)(e >> 32); ei = (unsigned)e;
            }
            m = __reduce_max_sync(kFull, em);
            wi = __reduce_min_sync(kFull, em == m ? ei : 0xffffffffu);
        }
        cur = (int)wi;
    }
}

template <int T, int P, bool R>
static int launch_fps(const float *xyz, int64_t sb, int64_t sc, int64_t sn, const int64_t *start,
                      int B, int N, int S, int32_t *oi, float *orows, float *ocf, cudaStream_t st) {
    const size_t smem = (size_t)3 * T * P * sizeof(float);
    auto k = fps_kernel<T, P, R>;
    if (smem + 1024 > 48 * 1024) {   // static slots count against the 48 KB default limit
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(EV2H_ERR_CUDA, "fps: smem attribute: %s", cudaGetErrorString(e));
    }
    k<<<B, T, smem, st>>>(xyz, sb, sc, sn, start, N, S, oi, orows, ocf);
    return check_launch("ev2h_fps_f32");
}

}  // namespace ev2h

extern "C" int ev2h_fps_f32(const float *xyz, int64_t stride_b, int64_t stride_c, int64_t stride_n,
                            const int64_t *start_idx, int B, int N, int S, int32_t *out_idx,
                            float *out_centres_rows, float *out_centres_cf, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(xyz && start_idx, "ev2h_fps_f32: null input");
    EV2H_REQUIRE(B > 0 && N > 0 && S > 0, "ev2h_fps_f32: B, N, S must be positive (got %d, %d, %d)", B, N, S);
    if (N > 16384) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_fps_f32: N=%d exceeds 16384 points per window", N);
    cudaStream_t st = as_stream(stream);
#define EV2H_FPS(T, P, R) \
    return launch_fps<T, P, R>(xyz, stride_b, stride_c, stride_n, start_idx, B, N, S, out_idx, out_centres_rows, out_centres_cf, st)
    if (N <= 128) EV2H_FPS(32, 4, true);
    if (N <= 256) EV2H_FPS(64, 4, true);
    if (N <= 512) EV2H_FPS(128, 4, true);
    if (N <= 1024) EV2H_FPS(256, 4, true);
    if (N <= 2048) EV2H_FPS(256, 8, true);      // 128 / 512 / 1024 threads per window measured 0.173 / 0.197 / 0.235 ms vs 0.179 ms (B = 64)
    if (N <= 4096) EV2H_FPS(512, 8, true);
    if (N <= 8192) EV2H_FPS(1024, 8, true);
    EV2H_FPS(1024, 16, false);
#undef EV2H_FPS
}
