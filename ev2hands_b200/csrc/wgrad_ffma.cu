// Weight gradient of a 1x1 convolution over point rows: dW[i, j] = sum_r dY[r, i] * X[r, j].
//
// The training path keeps the grouped tensor as rows [M, C] (M = B*S*K up to millions, C = 7 .. 1536), so the weight
// gradient of a layer (the autograd of Conv2d(1x1) in reference src/Ev2Hands/model/pointnet2_utils.py:253-256) is a GEMM
// whose CONTRACTION runs over the M rows and whose output is tiny ([Cout, Cin] <= 1024 x 1536).  Library SGEMM kernels
// reach ~11 TFLOP/s on that shape.  Here both operands are already in the layout the classic outer-product SGEMM wants -
// for one row r, dY[r, :] and X[r, :] are contiguous - so a CTA streams a slab of rows through shared memory with
// straight coalesced copies and every thread accumulates an 8 x 8 block of dW in registers; the slabs' partial results
// are summed in a fixed order by a second kernel (deterministic, exact fp32 FMAs).
#include "common.cuh"

namespace ev2h {
namespace wg {

constexpr int BK = 16, THREADS = 256;

// TM x TN outputs per thread (4 or 8 each): tiles of 16 TM x 16 TN, so narrow layers (32-96 channels) do not pay for a
// 128-wide tile
template <bool VEC, int TM, int TN>
__global__ void __launch_bounds__(THREADS, 2)       // two CTAs per SM: the 8 x 8 variant fits 128 registers
wgrad_partial_kernel(const float *__restrict__ dy, int ld_dy, const float *__restrict__ x, int ld_x,
                     int64_t M, int Cout, int Cin, int tiles_j, int64_t rows_per_split, float *__restrict__ partial,
                     float *__restrict__ bias_partial) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    __shared__ __align__(16) float As[2][BK][BM];      // dY rows: [k][i]
    __shared__ __align__(16) float Bs[2][BK][BN];      // X rows:  [k][j]
    const int tile = blockIdx.x, ti = tile / tiles_j, tj = tile % tiles_j;
    const int i0 = ti * BM, j0 = tj * BN;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r1 = r0 + rows_per_split < M ? r0 + rows_per_split : M;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[TM][TN], bsum[TM];
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        bsum[a] = 0.f;
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
    }
    // the bias gradient db[i] = sum_r dY[r, i] rides along in the tiles of the first column block: dY is streamed here anyway
    const bool with_bias = bias_partial != nullptr && tj == 0;

    // a stage = BK rows of each operand; BK * BM / 4 float4 of dY = TM / 4 per thread, likewise TN / 4 of X
    auto load4 = [&](const float *base, int ld, int64_t row, int c, int C) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < r1 && c < C) {
            const float *ptr = base + row * ld + c;
            if (VEC && c + 3 < C) v = __ldg(reinterpret_cast<const float4 *>(ptr));
            else { v.x = ptr[0]; if (c + 1 < C) v.y = ptr[1]; if (c + 2 < C) v.z = ptr[2]; if (c + 3 < C) v.w = ptr[3]; }
        }
        return v;
    };
    auto load_stage = [&](int64_t r, float4 (&va)[TM / 4], float4 (&vb)[TN / 4]) {
#pragma unroll
        for (int h = 0; h < TM / 4; ++h) {
            const int e = tid + h * THREADS, k = e / (BM / 4), c4 = (e % (BM / 4)) * 4;
            va[h] = load4(dy, ld_dy, r + k, i0 + c4, Cout);
        }
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            const int e = tid + h * THREADS, k = e / (BN / 4), c4 = (e % (BN / 4)) * 4;
            vb[h] = load4(x, ld_x, r + k, j0 + c4, Cin);
        }
    };
    auto store_stage = [&](int buf, const float4 (&va)[TM / 4], const float4 (&vb)[TN / 4]) {
#pragma unroll
        for (int h = 0; h < TM / 4; ++h) {
            const int e = tid + h * THREADS, k = e / (BM / 4), c4 = (e % (BM / 4)) * 4;
            *reinterpret_cast<float4 *>(&As[buf][k][c4]) = va[h];
        }
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            const int e = tid + h * THREADS, k = e / (BN / 4), c4 = (e % (BN / 4)) * 4;
            *reinterpret_cast<float4 *>(&Bs[buf][k][c4]) = vb[h];
        }
    };
    float4 va[TM / 4], vb[TN / 4];
    if (r0 < r1) {
        load_stage(r0, va, vb);
        store_stage(0, va, vb);
    }
    __syncthreads();
    int buf = 0;
    for (int64_t r = r0; r < r1; r += BK) {
        const bool more = r + BK < r1;
        if (more) load_stage(r + BK, va, vb);                  // global loads in flight during the FMAs
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            // thread (ty, tx): rows i = 64 h + 4 ty + {0..3}, columns j = 64 h + 4 tx + {0..3}
            float av[TM], bv[TN];
#pragma unroll
            for (int h = 0; h < TM / 4; ++h) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[buf][k][64 * h + ty * 4]);
                av[4 * h] = a.x; av[4 * h + 1] = a.y; av[4 * h + 2] = a.z; av[4 * h + 3] = a.w;
            }
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 * h + tx * 4]);
                bv[4 * h] = b.x; bv[4 * h + 1] = b.y; bv[4 * h + 2] = b.z; bv[4 * h + 3] = b.w;
            }
#pragma unroll
            for (int a = 0; a < TM; ++a)
#pragma unroll
                for (int b = 0; b < TN; ++b) acc[a][b] = __fmaf_rn(av[a], bv[b], acc[a][b]);
            if (with_bias) {
#pragma unroll
                for (int a = 0; a < TM; ++a) bsum[a] += av[a];
            }
        }
        if (more) store_stage(buf ^ 1, va, vb);
        __syncthreads();
        buf ^= 1;
    }
    float *out = partial + (int64_t)blockIdx.y * Cout * Cin;
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        const int i = i0 + 64 * (a / 4) + ty * 4 + (a & 3);
        if (i >= Cout) continue;
#pragma unroll
        for (int b = 0; b < TN; ++b) {
            const int j = j0 + 64 * (b / 4) + tx * 4 + (b & 3);
            if (j < Cin) out[(int64_t)i * Cin + j] = acc[a][b];
        }
        if (with_bias && tx == 0) bias_partial[(int64_t)blockIdx.y * Cout + i] = bsum[a];
    }
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ partial, int splits, int64_t n, float *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * n + e];     // fixed order: deterministic
    out[e] = s;
}

}  // namespace wg
}  // namespace ev2h

namespace ev2h {
namespace wg {
static int tile_of(int C) { return C <= 64 ? 64 : 128; }
}  // namespace wg
}  // namespace ev2h

extern "C" int ev2h_wgrad_splits(int64_t M, int Cout, int Cin) {
    using namespace ev2h::wg;
    if (M <= 0 || Cout <= 0 || Cin <= 0) return -1;
    const int bm = tile_of(Cout), bn = tile_of(Cin);
    const int64_t tiles = (int64_t)((Cout + bm - 1) / bm) * ((Cin + bn - 1) / bn);
    int64_t splits = (148 * 4 + tiles - 1) / tiles;            // four CTAs per SM's worth of slabs
    const int64_t max_by_rows = (M + 8 * BK - 1) / (8 * BK);   // at least 128 rows per slab
    if (splits > max_by_rows) splits = max_by_rows;
    if (splits < 1) splits = 1;
    if (splits > 2048) splits = 2048;
    return (int)splits;
}

/* dW [Cout, Cin] (row-major, contiguous) = dY^T X with dY [M, ld_dy >= Cout] and X [M, ld_x >= Cin] fp32 rows: the
 * weight gradient of a 1x1 convolution over M point rows (autograd of pointnet2_utils.py:253-256 in the row layout);
 * db [Cout] = column sums of dY (the bias gradient; NULL: not wanted).
 * workspace: ev2h_wgrad_splits(M, Cout, Cin) * Cout * (Cin + 1) floats.  Exact fp32 FMAs, deterministic. */
extern "C" int ev2h_wgrad_f32(const float *dy, int ld_dy, const float *x, int ld_x, int64_t M, int Cout, int Cin,
                              float *workspace, float *dw, float *db, ev2h_stream_t stream) {
    using namespace ev2h;
    using namespace ev2h::wg;
    EV2H_REQUIRE(dy && x && workspace && dw, "ev2h_wgrad_f32: null argument");
    EV2H_REQUIRE(M > 0 && Cout > 0 && Cin > 0 && ld_dy >= Cout && ld_x >= Cin, "ev2h_wgrad_f32: bad sizes");
    const int bm = tile_of(Cout), bn = tile_of(Cin);
    const int tiles_i = (Cout + bm - 1) / bm, tiles_j = (Cin + bn - 1) / bn;
    const int splits = ev2h_wgrad_splits(M, Cout, Cin);
    int64_t rows_per_split = (M + splits - 1) / splits;
    rows_per_split = (rows_per_split + BK - 1) / BK * BK;
    dim3 grid((unsigned)(tiles_i * tiles_j), (unsigned)splits);
    const bool vec = ld_dy % 4 == 0 && ld_x % 4 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0;
    cudaStream_t st = as_stream(stream);
    float *bias_partial = db != nullptr ? workspace + (int64_t)splits * Cout * Cin : nullptr;
#define EV2H_WG(V, TM, TN) \
    wgrad_partial_kernel<V, TM, TN><<<grid, THREADS, 0, st>>>(dy, ld_dy, x, ld_x, M, Cout, Cin, tiles_j, rows_per_split, workspace, bias_partial)
    if (vec) {
        if (bm == 64 && bn == 64) EV2H_WG(true, 4, 4);
        else if (bm == 64) EV2H_WG(true, 4, 8);
        else if (bn == 64) EV2H_WG(true, 8, 4);
        else EV2H_WG(true, 8, 8);
    } else {
        if (bm == 64 && bn == 64) EV2H_WG(false, 4, 4);
        else if (bm == 64) EV2H_WG(false, 4, 8);
        else if (bn == 64) EV2H_WG(false, 8, 4);
        else EV2H_WG(false, 8, 8);
    }
#undef EV2H_WG
    int rc = check_launch("ev2h_wgrad_f32");
    if (rc != EV2H_OK) return rc;
    const int64_t n = (int64_t)Cout * Cin;
    wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, splits, n, dw);
    if (db != nullptr) wgrad_reduce_kernel<<<(unsigned)((Cout + 255) / 256), 256, 0, st>>>(bias_partial, splits, Cout, db);
    return check_launch("ev2h_wgrad_f32 (reduce)");
}
