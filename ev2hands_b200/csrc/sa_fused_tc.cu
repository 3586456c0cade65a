// Fused grouping + shared MLP + max-pool for one radius scale of a multi-scale set abstraction
// layer, on the tcgen05 tensor cores.  Activations never leave the SM:
//
//   gather rows (loaders) -> smem operand ring -> UMMA -> TMEM accumulator
//        -> epilogue (bias, ReLU, split/convert) -> smem operand ring -> UMMA -> ... -> max over K
//
// Replaces, for one (radius, K) scale, the body of PointNetSetAbstractionMsg.forward's loop
// (reference src/Ev2Hands/model/pointnet2_utils.py:243-257): index_points / subtract / cat,
// the Conv2d(1x1)+BatchNorm2d(eval)+ReLU stack and torch.max over the K neighbours.
//
// A tile is 128 consecutive (centre, neighbour) rows, i.e. 128/K whole groups.  Its work is a
// fixed sequence of "K chunks" (32 input channels x 128 rows of one layer's input, plus that
// layer's weights for those channels).  Chunks flow through two shared-memory rings:
//   A ring  operand rows, written by the loader warps (first layer: gathered points) or by the
//           epilogue warps (later layers: the previous layer's activations straight from TMEM)
//   B ring  pre-packed weight images, one bulk async copy per chunk (weight-streamer thread)
// and are consumed in order by ONE issuing thread (tcgen05.mma), which signals ring slots free
// and accumulators complete through tcgen05.commit -> mbarrier.
//
// Two first-layer modes:
//   mode_b = 0  rows are [features(D) | xyz - centre] (<= 8 channels) gathered from a packed
//               [B,N,8] point table; all three layers run here.                (sa1, regressor)
//   mode_b = 1  layer 1 is linear before its ReLU, so it is evaluated once per POINT instead of
//               once per (centre, neighbour) row:  relu(W1 [f(p); xyz(p) - c(s)] + b1)
//               = relu(P[p] - C[s]) with P = W1 [f; xyz] + b1 per point and C = W1_xyz c per
//               centre, both precomputed; the loaders gather P, subtract C, apply ReLU, and the
//               kernel runs layers 2 and 3.  (sa2: 323 input channels, 32x fewer layer-1 MACs)
//
// Arithmetic modes as in linear_tc.cu: TF32X3 (fp32-level accuracy) or BF16.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <string.h>

namespace ev2h {

constexpr int FZ_BLOCK_M = 128;
constexpr int FZ_MAX_GEMMS = 3;
constexpr int FZ_MAX_RING = 8;
// Template parameters of the kernel:
//   KC  channels per K chunk (32, or 16 to halve the ring footprint so two CTAs share an SM)
//   LG  loader groups of 4 warps;  threads = 32 * (6 + 4 LG): 4 epilogue, issuer, weight streamer, loaders
//   OCC CTAs per SM the instance is built for (launch bound and TMEM share: 512 / OCC columns)
__host__ __device__ constexpr int fz_threads(int lg) { return 32 * (6 + 4 * lg); }

enum { FZ_MODE_BF16 = 0, FZ_MODE_TF32X3 = 1 };

struct FusedParams {
    // geometry of the grouping
    int B, N, S, K;
    const int32_t *idx; int idx_ld, k_off;       // ball-query result [B,S,idx_ld], this scale at k_off
    const float *centres;                        // [B,S,3]
    // first-layer source
    int mode_b;                                  // 0 = gather rows, 1 = per-point layer 1 (P - C)
    int ffma_first;                              // gather mode: loaders evaluate layer 1 (<= 8 -> c1 channels) in fp32 FFMA
    const float *first_wt; int first_ld, c1;     // folded layer-1 weights, input-channel major [16, first_ld], and width
    const float *first_bias;
    const float *pts8; int D;                    // gather modes: [B,N,8] rows = [features(D) | xyz | 0]
    const float *P; int ld_p, p_col;             // mode B: per-point layer-1 pre-activation [B*N, ld_p]
    const float *C; int ld_c, c_col;             // mode B: per-centre offset [B*S, ld_c]
    // the GEMM chain
    int G;
    int n[FZ_MAX_GEMMS];                         // accumulator width (multiple of 16, <= 256)
    int n_chunks[FZ_MAX_GEMMS];                  // K chunks of 32 input channels
    int k_steps_last[FZ_MAX_GEMMS];              // UMMA K steps in the last chunk
    int tmem_col[FZ_MAX_GEMMS];
    int bias_off[FZ_MAX_GEMMS];                  // offset into the bias array staged in shared memory
    const uint8_t *w[FZ_MAX_GEMMS];              // packed weight images (ev2h_tc_pack_weights)
    const float *bias[FZ_MAX_GEMMS];
    // output: pooled features of this scale, rows = centres
    float *out; int ld_out, out_col, c_out;
    int alias02;                                 // GEMM 2's accumulator reuses GEMM 0's TMEM columns
    int tmem_cols;                               // TMEM columns to allocate (power of two)
    // rings
    int sa, sb, a_slot_bytes, b_slot_bytes;
    long long *dbg;     // optional [gridDim.x][8] issuer wait-cycle counters (debug/profiling only)
};

struct Ring {
    int slot; uint32_t phase; int size;
    __device__ __forceinline__ void advance() { if (++slot == size) { slot = 0; phase ^= 1; } }
    __device__ __forceinline__ void advance(int k) { for (int i = 0; i < k; ++i) advance(); }
};

template <int MODE, int KC, int LG, int OCC>
__global__ void __launch_bounds__(fz_threads(LG), OCC)
sa_fused_tc_kernel(const FusedParams p) {
    extern __shared__ __align__(128) uint8_t fz_smem[];
    constexpr int FZ_KC = KC;
    constexpr int FZ_THREADS = fz_threads(LG);
    constexpr int FZ_LOADER_GROUPS = LG;
    constexpr int EB = MODE == FZ_MODE_BF16 ? 2 : 4;
    constexpr int PARTS = MODE == FZ_MODE_BF16 ? 1 : 2;
    constexpr int A_PART = FZ_BLOCK_M * FZ_KC * EB;      // per precision part: 16 KB (tf32, KC 32) / 8 KB
    constexpr int NCH = FZ_KC * EB / 16;                 // 16-byte operand chunks per row per K chunk
    constexpr int UMMA_K = 32 / EB;
    constexpr int K_STEPS = FZ_KC / UMMA_K;
    static_assert(NCH >= 2 && (MODE == FZ_MODE_TF32X3 || KC == 32), "unsupported chunk geometry");
    constexpr int CHUNK_ROWS_BYTES = FZ_BLOCK_M * 16;    // one 16-byte operand chunk for all 128 rows

    uint8_t *a_ring = fz_smem;
    uint8_t *b_ring = a_ring + (size_t)p.sa * p.a_slot_bytes;
    uint8_t *tail = b_ring + (size_t)p.sb * p.b_slot_bytes;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *a_empty = a_full + FZ_MAX_RING;
    uint64_t *b_full = a_empty + FZ_MAX_RING;
    uint64_t *b_empty = b_full + FZ_MAX_RING;
    uint64_t *acc_full = b_empty + FZ_MAX_RING;          // [3]
    uint64_t *acc_empty = acc_full + FZ_MAX_GEMMS;       // [3]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + FZ_MAX_GEMMS + 1);
    float *bias_s = reinterpret_cast<float *>(tmem_slot + 6);          // [sum n[g]], 16-byte aligned (tail offset 336)
    int bias_total = 0;
    for (int g = 0; g < p.G; ++g) bias_total += p.n[g];
    float *red = bias_s + bias_total;                                  // [2][4][n[G-1]]
    float *w1s = red + 2 * 4 * p.n[p.G - 1];                           // [c1_pad][8] layer-1 weights (ffma_first)
    float *b1s = w1s + (p.ffma_first ? p.n_chunks[0] * KC * 8 : 0);    // [c1_pad]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t M = (int64_t)p.B * p.S * p.K;
    const int64_t n_tiles = (M + FZ_BLOCK_M - 1) / FZ_BLOCK_M;
    int chunks_per_tile = 0;
    for (int g = 0; g < p.G; ++g) chunks_per_tile += p.n_chunks[g];

    if (tid == 0) {
        for (int s = 0; s < p.sa; ++s) { tc::mbar_init(a_full + s, 128); tc::mbar_init(a_empty + s, 1); }
        for (int s = 0; s < p.sb; ++s) { tc::mbar_init(b_full + s, 1); tc::mbar_init(b_empty + s, 1); }
        for (int g = 0; g < FZ_MAX_GEMMS; ++g) { tc::mbar_init(acc_full + g, 1); tc::mbar_init(acc_empty + g, 128); }
        tc::fence_mbar_init();
    }
    for (int g = 0; g < p.G; ++g)
        for (int i = tid; i < p.n[g]; i += FZ_THREADS) bias_s[p.bias_off[g] + i] = p.bias[g][i];
    if (p.ffma_first) {
        const int c1_pad = p.n_chunks[0] * KC;
        for (int i = tid; i < c1_pad * 8; i += FZ_THREADS) {
            const int ch = i >> 3, k = i & 7;
            w1s[i] = ch < p.c1 ? p.first_wt[(size_t)k * p.first_ld + ch] : 0.f;
        }
        for (int i = tid; i < c1_pad; i += FZ_THREADS) b1s[i] = i < p.c1 ? p.first_bias[i] : 0.f;
    }
    if (warp == 4) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 6) {
        // =============================== loaders: first layer's operand rows ===============================
        const int lw = warp - 6, grp = lw >> 2, wq = lw & 3;
        Ring ra{0, 0, p.sa};
        uint32_t ln = 0;                          // counts first-layer chunks; groups alternate on it
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t m0 = tile * FZ_BLOCK_M;
            if (p.ffma_first) {
                // ---- gather + layer 1 in exact fp32 on the CUDA cores; thread = row ------------------------
                // x = [features | xyz - centre] (8 floats, one 32-byte sector), then for every K chunk of the
                // SECOND layer's input: relu(W1' x + b1') for KC channels, split, store as operand rows.
                const int r = wq * 32 + lane;
                const int64_t R = m0 + r;
                float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                bool valid = false;
                if (R < M) {
                    const int64_t bs = R / p.K;
                    const int j = (int)(R - bs * p.K);
                    const int64_t b = bs / p.S;
                    const int pt = p.idx[bs * p.idx_ld + p.k_off + j];
                    if (pt >= 0 && pt < p.N) {
                        valid = true;
                        const float4 *src = reinterpret_cast<const float4 *>(p.pts8 + (b * p.N + pt) * 8);
                        const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
                        x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                        const float *c = p.centres + bs * 3;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float ca = __ldg(c + a);
#pragma unroll
                            for (int ch = 0; ch < 8; ++ch)
                                if (ch == p.D + a) x[ch] = __fsub_rn(x[ch], ca);   // grouped_xyz -= new_xyz (:245)
                        }
                    }
                }
                for (int kc = 0; kc < p.n_chunks[0]; ++kc, ++ln) {
                    if ((int)(ln % FZ_LOADER_GROUPS) == grp) {
                        float v[FZ_KC];
#pragma unroll
                        for (int j = 0; j < FZ_KC; ++j) {
                            const int ch = kc * FZ_KC + j;
                            const float4 wa = *reinterpret_cast<const float4 *>(w1s + ch * 8);
                            const float4 wb = *reinterpret_cast<const float4 *>(w1s + ch * 8 + 4);
                            float acc = b1s[ch];
                            acc = fmaf(wa.x, x[0], acc); acc = fmaf(wa.y, x[1], acc); acc = fmaf(wa.z, x[2], acc); acc = fmaf(wa.w, x[3], acc);
                            acc = fmaf(wb.x, x[4], acc); acc = fmaf(wb.y, x[5], acc); acc = fmaf(wb.z, x[6], acc); acc = fmaf(wb.w, x[7], acc);
                            v[j] = valid ? fmaxf(acc, 0.f) : 0.f;
                        }
                        tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 16);
                        uint8_t *st = a_ring + (size_t)ra.slot * p.a_slot_bytes;
                        if (MODE == FZ_MODE_TF32X3) {
#pragma unroll
                            for (int cc = 0; cc < NCH; ++cc) {
                                float4 hi, lo;
                                tc::split_tf32(v[4 * cc], hi.x, lo.x); tc::split_tf32(v[4 * cc + 1], hi.y, lo.y);
                                tc::split_tf32(v[4 * cc + 2], hi.z, lo.z); tc::split_tf32(v[4 * cc + 3], hi.w, lo.w);
                                *reinterpret_cast<float4 *>(st + cc * CHUNK_ROWS_BYTES + r * 16) = hi;
                                *reinterpret_cast<float4 *>(st + A_PART + cc * CHUNK_ROWS_BYTES + r * 16) = lo;
                            }
                        } else {
#pragma unroll
                            for (int cc = 0; cc < NCH; ++cc) {
                                __nv_bfloat162 q0 = __floats2bfloat162_rn(v[8 * cc], v[8 * cc + 1]);
                                __nv_bfloat162 q1 = __floats2bfloat162_rn(v[8 * cc + 2], v[8 * cc + 3]);
                                __nv_bfloat162 q2 = __floats2bfloat162_rn(v[8 * cc + 4], v[8 * cc + 5]);
                                __nv_bfloat162 q3 = __floats2bfloat162_rn(v[8 * cc + 6], v[8 * cc + 7]);
                                uint4 pk;
                                pk.x = *reinterpret_cast<uint32_t *>(&q0); pk.y = *reinterpret_cast<uint32_t *>(&q1);
                                pk.z = *reinterpret_cast<uint32_t *>(&q2); pk.w = *reinterpret_cast<uint32_t *>(&q3);
                                *reinterpret_cast<uint4 *>(st + cc * CHUNK_ROWS_BYTES + r * 16) = pk;
                            }
                        }
                        tc::fence_proxy_async();
                        tc::mbar_arrive(a_full + ra.slot);
                    } else {
                        tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 17);      // in-order walk
                    }
                    ra.advance();
                }
                for (int i = p.n_chunks[0]; i < chunks_per_tile; ++i) {
                    tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 18);
                    ra.advance();
                }
            } else if (!p.mode_b) {
                // ---- mode A: one chunk per tile; thread = row -------------------------------------------
                // Every loader walks EVERY chunk's "slot free" barrier in order, also for chunks other
                // warps fill: a parity wait is only meaningful when the waiter is at most one phase
                // ahead of the barrier, and skipping waits would let a group race a whole tile ahead.
                bool mine_waited = false;
                if ((int)(ln % FZ_LOADER_GROUPS) == grp) {
                    const int r = wq * 32 + lane;
                    const int64_t R = m0 + r;
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (R < M) {
                        const int64_t bs = R / p.K;
                        const int j = (int)(R - bs * p.K);
                        const int64_t b = bs / p.S;
                        const int pt = p.idx[bs * p.idx_ld + p.k_off + j];
                        if (pt >= 0 && pt < p.N) {
                            const float4 *src = reinterpret_cast<const float4 *>(p.pts8 + (b * p.N + pt) * 8);
                            v0 = __ldg(src); v1 = __ldg(src + 1);
                            float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                            const float *c = p.centres + bs * 3;
#pragma unroll
                            for (int a = 0; a < 3; ++a) {
                                const float ca = __ldg(c + a);
#pragma unroll
                                for (int ch = 0; ch < 8; ++ch)
                                    if (ch == p.D + a) f[ch] = __fsub_rn(f[ch], ca);   // grouped_xyz -= new_xyz (:245)
                            }
                            v0 = make_float4(f[0], f[1], f[2], f[3]); v1 = make_float4(f[4], f[5], f[6], f[7]);
                        }
                    }
                    tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 10);
                    mine_waited = true;
                    uint8_t *st = a_ring + (size_t)ra.slot * p.a_slot_bytes;
                    if (MODE == FZ_MODE_TF32X3) {
                        float4 h0, l0, h1, l1;
                        tc::split_tf32(v0.x, h0.x, l0.x); tc::split_tf32(v0.y, h0.y, l0.y);
                        tc::split_tf32(v0.z, h0.z, l0.z); tc::split_tf32(v0.w, h0.w, l0.w);
                        tc::split_tf32(v1.x, h1.x, l1.x); tc::split_tf32(v1.y, h1.y, l1.y);
                        tc::split_tf32(v1.z, h1.z, l1.z); tc::split_tf32(v1.w, h1.w, l1.w);
                        *reinterpret_cast<float4 *>(st + r * 16) = h0;
                        *reinterpret_cast<float4 *>(st + CHUNK_ROWS_BYTES + r * 16) = h1;
                        *reinterpret_cast<float4 *>(st + A_PART + r * 16) = l0;
                        *reinterpret_cast<float4 *>(st + A_PART + CHUNK_ROWS_BYTES + r * 16) = l1;
                    } else {
                        __nv_bfloat162 q0 = __floats2bfloat162_rn(v0.x, v0.y), q1 = __floats2bfloat162_rn(v0.z, v0.w);
                        __nv_bfloat162 q2 = __floats2bfloat162_rn(v1.x, v1.y), q3 = __floats2bfloat162_rn(v1.z, v1.w);
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t *>(&q0); pk.y = *reinterpret_cast<uint32_t *>(&q1);
                        pk.z = *reinterpret_cast<uint32_t *>(&q2); pk.w = *reinterpret_cast<uint32_t *>(&q3);
                        *reinterpret_cast<uint4 *>(st + r * 16) = pk;                             // channels 0-7
                        *reinterpret_cast<uint4 *>(st + CHUNK_ROWS_BYTES + r * 16) = make_uint4(0, 0, 0, 0);   // 8-15: zero
                    }
                    tc::fence_proxy_async();
                    tc::mbar_arrive(a_full + ra.slot);
                }
                ++ln;
                if (!mine_waited) tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 12);
                ra.advance();
                for (int i = 1; i < chunks_per_tile; ++i) {
                    tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 13);
                    ra.advance();
                }
            } else {
                // ---- mode B: n_chunks[0] chunks of relu(P[p] - C[s]); octet lane mapping ------------------
                const int l8 = lane & 7, oct = lane >> 3;
                // the 4 rows this lane touches in every chunk
                int64_t p_row[4], c_row[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int64_t R = m0 + 32 * wq + h * 8 + l8;
                    p_row[h] = -1; c_row[h] = 0;
                    if (R < M) {
                        const int64_t bs = R / p.K;
                        const int j = (int)(R - bs * p.K);
                        const int64_t b = bs / p.S;
                        const int pt = p.idx[bs * p.idx_ld + p.k_off + j];
                        if (pt >= 0 && pt < p.N) { p_row[h] = b * p.N + pt; c_row[h] = bs; }
                    }
                }
                for (int kc = 0; kc < p.n_chunks[0]; ++kc, ++ln) {
                    if ((int)(ln % FZ_LOADER_GROUPS) == grp) {
                        constexpr int QP = FZ_KC / 16;          // passes of 4 channel quads per 8-row group
                        float4 v[4 * QP];
#pragma unroll
                        for (int i = 0; i < 4 * QP; ++i) {
                            const int h = i / QP;
                            const int k = kc * FZ_KC + 4 * (oct + 4 * (i % QP));
                            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p_row[h] >= 0) {
                                const float4 a = __ldg(reinterpret_cast<const float4 *>(p.P + p_row[h] * p.ld_p + p.p_col + k));
                                const float4 c = __ldg(reinterpret_cast<const float4 *>(p.C + c_row[h] * p.ld_c + p.c_col + k));
                                v[i] = make_float4(fmaxf(a.x - c.x, 0.f), fmaxf(a.y - c.y, 0.f), fmaxf(a.z - c.z, 0.f), fmaxf(a.w - c.w, 0.f));
                            }
                        }
                        tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 11);
                        uint8_t *st = a_ring + (size_t)ra.slot * p.a_slot_bytes;
                        if (MODE == FZ_MODE_TF32X3) {
#pragma unroll
                            for (int i = 0; i < 4 * QP; ++i) {
                                const int row = 32 * wq + (i / QP) * 8 + l8;
                                const int c = oct + 4 * (i % QP);
                                float4 hi, lo;
                                tc::split_tf32(v[i].x, hi.x, lo.x); tc::split_tf32(v[i].y, hi.y, lo.y);
                                tc::split_tf32(v[i].z, hi.z, lo.z); tc::split_tf32(v[i].w, hi.w, lo.w);
                                *reinterpret_cast<float4 *>(st + c * CHUNK_ROWS_BYTES + row * 16) = hi;
                                *reinterpret_cast<float4 *>(st + A_PART + c * CHUNK_ROWS_BYTES + row * 16) = lo;
                            }
                        } else {
                            // bf16: a 16-byte operand chunk holds 8 channels = two of the fp32 float4s.
                            // lanes oct and oct+... own channel quads (oct + 4*(i&1)); pair them through shuffles:
                            // quad q (0..7) belongs to chunk q/2; lane octet `oct` holds quads oct and oct+4.
#pragma unroll
                            for (int i = 0; i < 4 * QP; ++i) {
                                const int row = 32 * wq + (i / QP) * 8 + l8;
                                const int quad = oct + 4 * (i % QP);
                                __nv_bfloat162 q0 = __floats2bfloat162_rn(v[i].x, v[i].y), q1 = __floats2bfloat162_rn(v[i].z, v[i].w);
                                uint2 pk;
                                pk.x = *reinterpret_cast<uint32_t *>(&q0); pk.y = *reinterpret_cast<uint32_t *>(&q1);
                                *reinterpret_cast<uint2 *>(st + (quad >> 1) * CHUNK_ROWS_BYTES + row * 16 + (quad & 1) * 8) = pk;
                            }
                        }
                        tc::fence_proxy_async();
                        tc::mbar_arrive(a_full + ra.slot);
                    } else {
                        tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 14);      // in-order walk, see mode A
                    }
                    ra.advance();
                }
                for (int i = p.n_chunks[0]; i < chunks_per_tile; ++i) {
                    tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 15);
                    ra.advance();
                }
            }
        }
    } else if (warp == 5) {
        // =============================== weight streamer ===============================
        if (lane == 0) {
            Ring rb{0, 0, p.sb};
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int g = 0; g < p.G; ++g) {
                    const uint32_t bytes = (uint32_t)(PARTS * p.n[g] * FZ_KC * EB);
                    for (int c = 0; c < p.n_chunks[g]; ++c) {
                        tc::mbar_wait(b_empty + rb.slot, rb.phase ^ 1, 20);
                        tc::mbar_arrive_expect_tx(b_full + rb.slot, bytes);
                        tc::bulk_g2s(b_ring + (size_t)rb.slot * p.b_slot_bytes, p.w[g] + (size_t)c * bytes, bytes, b_full + rb.slot);
                        rb.advance();
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        // =============================== UMMA issuer ===============================
        // The whole warp walks the chunk sequence (so every value below is warp-uniform and the
        // descriptor arithmetic stays on the uniform datapath); one elected lane issues.
        Ring ra{0, 0, p.sa}, rb{0, 0, p.sb};
        uint32_t it = 0;
        const bool prof = p.dbg != nullptr;
        long long w_a[3] = {0, 0, 0}, w_b = 0, w_acc = 0, w_commit = 0, t0 = 0, t1 = 0;
        const long long t_begin = prof ? clock64() : 0;
        const uint32_t a_lbo = CHUNK_ROWS_BYTES, sbo = 128;
        const uint32_t desc_hi = tc::smem_desc_hi(sbo);
        const uint32_t a_ring_addr = tc::smem_u32(a_ring), b_ring_addr = tc::smem_u32(b_ring);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            for (int g = 0; g < p.G; ++g) {
                const uint32_t idesc = tc::instr_desc(MODE == FZ_MODE_BF16 ? tc::FMT_BF16 : tc::FMT_TF32, FZ_BLOCK_M, (uint32_t)p.n[g]);
                const uint32_t b_lbo = (uint32_t)p.n[g] * 16;
                const uint32_t b_part = (uint32_t)(p.n[g] * FZ_KC * EB);
                const uint32_t d_tmem = tmem_base + (uint32_t)p.tmem_col[g];
                const int n_chunks = p.n_chunks[g];
                if (prof) t0 = clock64();
                tc::mbar_wait(acc_empty + g, (it & 1) ^ 1, 30 + g);       // previous tile's epilogue drained this accumulator
                if (p.alias02) {
                    // GEMM 0 and GEMM 2 share TMEM columns: GEMM 0 also needs the previous tile's pooled
                    // epilogue (GEMM 2) done, GEMM 2 needs THIS tile's GEMM 0 activations converted
                    if (g == 0) tc::mbar_wait(acc_empty + 2, (it & 1) ^ 1, 36);
                    if (g == 2) tc::mbar_wait(acc_empty + 0, it & 1, 37);
                }
                if (prof) w_acc += clock64() - t0;
                tc::tc_fence_after();
                for (int c = 0; c < n_chunks; ++c) {
                    if (prof) t0 = clock64();
                    tc::mbar_wait(a_full + ra.slot, ra.phase, 40 + g);
                    if (prof) { t1 = clock64(); w_a[g] += t1 - t0; }
                    tc::mbar_wait(b_full + rb.slot, rb.phase, 50 + g);
                    if (prof) w_b += clock64() - t1;
                    tc::tc_fence_after();
                    const uint32_t a0 = a_ring_addr + (uint32_t)ra.slot * (uint32_t)p.a_slot_bytes;
                    const uint32_t b0 = b_ring_addr + (uint32_t)rb.slot * (uint32_t)p.b_slot_bytes;
                    const int ks = (c == n_chunks - 1) ? p.k_steps_last[g] : K_STEPS;
                    if (tc::elect_one()) {
                        // descriptor low words; one K step = two 16-byte chunks further along K
                        uint32_t a_hi = tc::smem_desc_lo(a0, a_lbo), b_hi = tc::smem_desc_lo(b0, b_lbo);
                        const uint32_t a_step = (2 * a_lbo) >> 4, b_step = (2 * b_lbo) >> 4;
                        const uint32_t a_lo_off = A_PART >> 4, b_lo_off = b_part >> 4;
#pragma unroll 1
                        for (int j = 0; j < ks; ++j) {
                            const uint32_t acc = (c > 0 || j > 0) ? 1u : 0u;
                            if (MODE == FZ_MODE_TF32X3) {
                                tc::umma_tf32(d_tmem, tc::make_desc(a_hi + a_lo_off, desc_hi), tc::make_desc(b_hi, desc_hi), idesc, acc);
                                tc::umma_tf32(d_tmem, tc::make_desc(a_hi, desc_hi), tc::make_desc(b_hi + b_lo_off, desc_hi), idesc, 1u);
                                tc::umma_tf32(d_tmem, tc::make_desc(a_hi, desc_hi), tc::make_desc(b_hi, desc_hi), idesc, 1u);
                            } else {
                                tc::umma_f16(d_tmem, tc::make_desc(a_hi, desc_hi), tc::make_desc(b_hi, desc_hi), idesc, acc);
                            }
                            a_hi += a_step; b_hi += b_step;
                        }
                        if (prof) t0 = clock64();
                        tc::umma_commit(a_empty + ra.slot);
                        tc::umma_commit(b_empty + rb.slot);
                        if (c == n_chunks - 1) tc::umma_commit(acc_full + g);
                        if (prof) w_commit += clock64() - t0;
                    }
                    __syncwarp();
                    ra.advance(); rb.advance();
                }
            }
        }
        if (prof && lane == 0) {
            long long *d = p.dbg + (size_t)blockIdx.x * 8;
            d[0] = w_a[0]; d[1] = w_a[1]; d[2] = w_a[2]; d[3] = w_b; d[4] = w_acc; d[5] = clock64() - t_begin; d[6] = it; d[7] = w_commit;
        }
    } else {
        // =============================== epilogue warps ===============================
        const int q = warp, r = q * 32 + lane;
        const int K = p.K;
        Ring ra{0, 0, p.sa};
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int64_t m0 = tile * FZ_BLOCK_M;
            ra.advance(p.n_chunks[0]);                                 // first layer's chunks belong to the loaders
            for (int g = 0; g < p.G; ++g) {
                const uint32_t t_addr = tmem_base + (uint32_t)p.tmem_col[g] + ((uint32_t)(q * 32) << 16);
                const float *bias_g = bias_s + p.bias_off[g];
                tc::mbar_wait(acc_full + g, it & 1, 60 + g);
                tc::tc_fence_after();
                if (g < p.G - 1) {
                    // ---- activations of layer g -> operand chunks of layer g+1 ----
                    for (int c = 0; c < p.n_chunks[g + 1]; ++c) {
                        uint32_t raw[FZ_KC];
                        if constexpr (FZ_KC == 32) tc::tmem_ld32(t_addr + c * 32, raw);
                        else tc::tmem_ld16(t_addr + c * 16, raw);
                        tc::tmem_ld_wait();
                        float v[FZ_KC];
#pragma unroll
                        for (int j = 0; j < FZ_KC; ++j) {
                            const int col = c * FZ_KC + j;
                            v[j] = col < p.n[g] ? fmaxf(__uint_as_float(raw[j]) + bias_g[col], 0.f) : 0.f;
                        }
                        tc::mbar_wait(a_empty + ra.slot, ra.phase ^ 1, 70 + g);
                        uint8_t *st = a_ring + (size_t)ra.slot * p.a_slot_bytes;
                        if (MODE == FZ_MODE_TF32X3) {
#pragma unroll
                            for (int cc = 0; cc < NCH; ++cc) {
                                float4 hi, lo;
                                tc::split_tf32(v[4 * cc], hi.x, lo.x); tc::split_tf32(v[4 * cc + 1], hi.y, lo.y);
                                tc::split_tf32(v[4 * cc + 2], hi.z, lo.z); tc::split_tf32(v[4 * cc + 3], hi.w, lo.w);
                                *reinterpret_cast<float4 *>(st + cc * CHUNK_ROWS_BYTES + r * 16) = hi;
                                *reinterpret_cast<float4 *>(st + A_PART + cc * CHUNK_ROWS_BYTES + r * 16) = lo;
                            }
                        } else {
#pragma unroll
                            for (int cc = 0; cc < NCH; ++cc) {
                                __nv_bfloat162 q0 = __floats2bfloat162_rn(v[8 * cc], v[8 * cc + 1]);
                                __nv_bfloat162 q1 = __floats2bfloat162_rn(v[8 * cc + 2], v[8 * cc + 3]);
                                __nv_bfloat162 q2 = __floats2bfloat162_rn(v[8 * cc + 4], v[8 * cc + 5]);
                                __nv_bfloat162 q3 = __floats2bfloat162_rn(v[8 * cc + 6], v[8 * cc + 7]);
                                uint4 pk;
                                pk.x = *reinterpret_cast<uint32_t *>(&q0); pk.y = *reinterpret_cast<uint32_t *>(&q1);
                                pk.z = *reinterpret_cast<uint32_t *>(&q2); pk.w = *reinterpret_cast<uint32_t *>(&q3);
                                *reinterpret_cast<uint4 *>(st + cc * CHUNK_ROWS_BYTES + r * 16) = pk;
                            }
                        }
                        tc::fence_proxy_async();
                        tc::mbar_arrive(a_full + ra.slot);
                        ra.advance();
                    }
                    tc::tc_fence_before();
                    tc::mbar_arrive(acc_empty + g);
                } else {
                    // ---- last layer: bias + ReLU + max over the K rows of each group ----
                    const int n_last = p.n[g];
                    const bool row_ok = (m0 + r) < M;
                    float *red_w = red + ((size_t)(it & 1) * 4 + q) * n_last;
                    for (int c0 = 0; c0 < n_last; c0 += 32) {
                        uint32_t raw[32];
                        tc::tmem_ld32(t_addr + c0, raw);
                        tc::tmem_ld_wait();
                        float mine = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c0 + j;
                            const float v = col < n_last ? fmaxf(__uint_as_float(raw[j]) + bias_g[col], 0.f) : 0.f;
                            const unsigned m = __reduce_max_sync(0xffffffffu, row_ok ? __float_as_uint(v) : 0u);
                            if (lane == j) mine = __uint_as_float(m);
                        }
                        if (K == 32) {
                            const int64_t row0 = m0 + q * 32;
                            if (row0 < M && c0 + lane < p.c_out)
                                p.out[(row0 / 32) * (int64_t)p.ld_out + p.out_col + c0 + lane] = mine;
                        } else if (c0 + lane < n_last) {
                            red_w[c0 + lane] = mine;
                        }
                    }
                    tc::tc_fence_before();
                    tc::mbar_arrive(acc_empty + g);
                    if (K > 32) {
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        const float *rr = red + (size_t)(it & 1) * 4 * n_last;
                        if (K == 64) {
                            for (int i = tid; i < 2 * n_last; i += 128) {
                                const int gg = i / n_last, c = i % n_last;
                                const int64_t row0 = m0 + gg * 64;
                                if (row0 < M && c < p.c_out)
                                    p.out[(row0 / 64) * (int64_t)p.ld_out + p.out_col + c] =
                                        fmaxf(rr[(2 * gg) * n_last + c], rr[(2 * gg + 1) * n_last + c]);
                            }
                        } else {   // K == 128
                            for (int c = tid; c < n_last; c += 128)
                                if (c < p.c_out && m0 < M)
                                    p.out[(m0 / 128) * (int64_t)p.ld_out + p.out_col + c] =
                                        fmaxf(fmaxf(rr[c], rr[n_last + c]), fmaxf(rr[2 * n_last + c], rr[3 * n_last + c]));
                        }
                    }
                }
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace ev2h

namespace ev2h {
static long long *g_fused_dbg = nullptr;

// Which kernel instance serves a layer stack: two CTAs per SM (KC 16 for tf32) whenever the
// accumulators of one tile fit 256 TMEM columns, so one CTA's epilogues overlap the other's UMMAs.
struct FusedPlan { int kc, occ, lg, alias02, tmem_cols, col[FZ_MAX_GEMMS], n[FZ_MAX_GEMMS]; bool ok; };

static FusedPlan fused_plan(int mode, bool mode_b, int n_layers, const int32_t *cout) {
    FusedPlan pl;
    memset(&pl, 0, sizeof(pl));
    int sum = 0;
    for (int g = 0; g < n_layers; ++g) { pl.n[g] = round_up(cout[g], 16); sum += pl.n[g]; }
    auto extent = [&](bool alias) {
        int c = 0, ext = 0;
        for (int g = 0; g < n_layers; ++g) {
            if (alias && g == 2) { pl.col[2] = 0; }
            else if (alias && g == 0) { pl.col[0] = 0; c = pl.n[0] > pl.n[2] ? pl.n[0] : pl.n[2]; }
            else { pl.col[g] = c; c += pl.n[g]; }
            const int e = pl.col[g] + round_up(pl.n[g], 32);
            if (e > ext) ext = e;
        }
        return ext;
    };
    int ext = extent(false);
    pl.alias02 = 0;
    if (!mode_b && n_layers == 3 && ext > 256 && extent(true) <= 256) { pl.alias02 = 1; ext = extent(true); }
    else ext = extent(false);
    pl.ok = ext <= 512;
    pl.occ = ext <= 256 ? 2 : 1;
    pl.kc = (pl.occ == 2 && mode == FZ_MODE_TF32X3) ? 16 : 32;
    pl.lg = pl.occ == 2 ? 1 : 2;
    pl.tmem_cols = 32;
    while (pl.tmem_cols < ext) pl.tmem_cols *= 2;
    if (pl.occ == 2 && pl.tmem_cols > 256) pl.occ = 1;
    return pl;
}
}  // namespace ev2h

extern "C" int ev2h_fused_set_debug_buffer(void *buf) { ev2h::g_fused_dbg = (long long *)buf; return 0; }

extern "C" int ev2h_sa_msg_fused_kc(int mode, int per_point, int n_layers, const int32_t *cout_host) {
    using namespace ev2h;
    if (!cout_host || n_layers < 2 || n_layers > 3) return -1;
    const FusedPlan pl = fused_plan(mode, per_point != 0, n_layers, cout_host);
    return pl.ok ? pl.kc : -1;
}

template <int MODE, int KC, int LG, int OCC>
static int launch_fused(const ev2h::FusedParams &p, size_t smem, unsigned grid, cudaStream_t st) {
    using namespace ev2h;
    auto k = sa_fused_tc_kernel<MODE, KC, LG, OCC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(EV2H_ERR_CUDA, "ev2h_sa_msg_fused_tc: smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
    k<<<grid, fz_threads(LG), smem, st>>>(p);
    return check_launch("ev2h_sa_msg_fused_tc");
}

extern "C" int ev2h_sa_msg_fused_tc(
    const int32_t *idx, int idx_ld, int k_off, const float *centres_rows, int B, int N, int S, int K,
    const float *pts8, int D, const float *first_wt, int first_ld, const float *first_bias,
    const float *P, int ld_p, int p_col, const float *C, int ld_c, int c_col,
    int n_layers, const int32_t *cin_host, const int32_t *cout_host, const void *const *w_packed_host,
    const float *const *bias_host, float *out_rows, int ld_out, int out_col, int mode, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(idx && centres_rows && out_rows && cin_host && cout_host && w_packed_host && bias_host,
                 "ev2h_sa_msg_fused_tc: null argument");
    EV2H_REQUIRE(B > 0 && N > 0 && S > 0 && k_off >= 0 && k_off + K <= idx_ld, "ev2h_sa_msg_fused_tc: bad sizes");
    EV2H_REQUIRE(mode == FZ_MODE_BF16 || mode == FZ_MODE_TF32X3, "ev2h_sa_msg_fused_tc: unknown mode %d", mode);
    if (K != 32 && K != 64 && K != 128)
        return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: K=%d (supported: 32, 64, 128)", K);
    const bool mode_b = P != nullptr;
    const bool ffma_first = !mode_b && first_wt != nullptr;
    if (n_layers < 2 || n_layers > 3 || ((mode_b || ffma_first) && n_layers != 2) || (!mode_b && !ffma_first && n_layers != 3))
        return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: %d tensor-core layers in %s mode", n_layers,
                    mode_b ? "per-point" : (ffma_first ? "gather (layer 1 on CUDA cores)" : "gather"));
    if (!mode_b) {
        EV2H_REQUIRE(pts8 != nullptr, "ev2h_sa_msg_fused_tc: pts8 is null");
        if (D + 3 > 8 || (!ffma_first && cin_host[0] != D + 3))
            return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: gather mode needs D+3 <= 8 input channels (D=%d)", D);
        if (ffma_first) EV2H_REQUIRE(first_bias != nullptr && first_ld >= cin_host[0], "ev2h_sa_msg_fused_tc: first layer weights incomplete");
    } else {
        EV2H_REQUIRE(C != nullptr && ld_p % 4 == 0 && ld_c % 4 == 0 && p_col % 4 == 0 && c_col % 4 == 0,
                     "ev2h_sa_msg_fused_tc: per-point tables must be float4 addressable");
        if (cin_host[0] % 32 != 0)
            return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: per-point mode needs a multiple of 32 channels, got %d", cin_host[0]);
    }
    for (int g = 0; g < n_layers; ++g) {
        if (cout_host[g] > 256) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: layer width %d > 256", cout_host[g]);
        if (g > 0 && cin_host[g] != cout_host[g - 1]) return fail(EV2H_ERR_BAD_ARGUMENT, "ev2h_sa_msg_fused_tc: layer %d input width mismatch", g);
    }
    const FusedPlan pl = fused_plan(mode, mode_b, n_layers, cout_host);
    if (!pl.ok) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: accumulators exceed the 512 TMEM columns");
    const int KC = pl.kc;
    const int EB = mode == FZ_MODE_BF16 ? 2 : 4, PARTS = mode == FZ_MODE_BF16 ? 1 : 2, UMMA_K = 32 / EB;

    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.N = N; p.S = S; p.K = K; p.idx = idx; p.idx_ld = idx_ld; p.k_off = k_off; p.centres = centres_rows;
    p.mode_b = mode_b ? 1 : 0; p.pts8 = pts8; p.D = D; p.P = P; p.ld_p = ld_p; p.p_col = p_col; p.C = C; p.ld_c = ld_c; p.c_col = c_col;
    p.G = n_layers; p.alias02 = pl.alias02; p.tmem_cols = pl.tmem_cols;
    p.ffma_first = ffma_first ? 1 : 0; p.first_wt = first_wt; p.first_ld = first_ld; p.first_bias = first_bias; p.c1 = cin_host[0];
    int boff = 0, max_n = 0;
    for (int g = 0; g < n_layers; ++g) {
        const int cin = cin_host[g];
        p.n[g] = pl.n[g];
        p.n_chunks[g] = (cin + KC - 1) / KC;
        const int rem = cin - (p.n_chunks[g] - 1) * KC;
        p.k_steps_last[g] = (rem + UMMA_K - 1) / UMMA_K;
        p.tmem_col[g] = pl.col[g];
        p.bias_off[g] = boff; boff += p.n[g];
        p.w[g] = (const uint8_t *)w_packed_host[g]; p.bias[g] = bias_host[g];
        EV2H_REQUIRE(p.w[g] && p.bias[g] && ((uintptr_t)p.w[g] & 15) == 0, "ev2h_sa_msg_fused_tc: layer %d weights null or misaligned", g);
        if (p.n[g] > max_n) max_n = p.n[g];
    }
    p.out = out_rows; p.ld_out = ld_out; p.out_col = out_col; p.c_out = cout_host[n_layers - 1];
    EV2H_REQUIRE(ld_out >= out_col + p.c_out, "ev2h_sa_msg_fused_tc: ld_out too small");
    p.dbg = g_fused_dbg;

    p.a_slot_bytes = PARTS * FZ_BLOCK_M * KC * EB;
    p.b_slot_bytes = PARTS * max_n * KC * EB;
    const int tail = (4 * FZ_MAX_RING + 2 * FZ_MAX_GEMMS + 1) * 8 + 24 + (boff + 2 * 4 * p.n[n_layers - 1]) * 4 +
                     (ffma_first ? p.n_chunks[0] * KC * 9 * 4 : 0);
    int occ = pl.occ;
    int budget = (occ == 2 ? 113 : 227) * 1024 - tail - 512;
    if (occ == 2 && budget < 2 * p.a_slot_bytes + 2 * p.b_slot_bytes) { occ = 1; budget = 227 * 1024 - tail - 512; }
    // at least 2 slots each; a third operand slot when affordable; weights get the rest (prefetched furthest ahead)
    p.sa = 2;
    if (budget - 3 * p.a_slot_bytes >= 3 * p.b_slot_bytes) p.sa = 3;
    p.sb = (budget - p.sa * p.a_slot_bytes) / p.b_slot_bytes;
    if (p.sb > FZ_MAX_RING) p.sb = FZ_MAX_RING;
    if (p.sb < 2) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: rings do not fit in shared memory");
    const size_t smem = (size_t)p.sa * p.a_slot_bytes + (size_t)p.sb * p.b_slot_bytes + tail;

    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = ((int64_t)B * S * K + FZ_BLOCK_M - 1) / FZ_BLOCK_M;
    const int64_t slots = (int64_t)sms * occ;
    const unsigned grid = (unsigned)(n_tiles < slots ? n_tiles : slots);
    cudaStream_t st = as_stream(stream);
    if (mode == FZ_MODE_TF32X3) {
        if (KC == 16) return occ == 2 ? launch_fused<FZ_MODE_TF32X3, 16, 1, 2>(p, smem, grid, st)
                                      : launch_fused<FZ_MODE_TF32X3, 16, 1, 1>(p, smem, grid, st);
        return launch_fused<FZ_MODE_TF32X3, 32, 2, 1>(p, smem, grid, st);
    }
    if (pl.occ == 2) return occ == 2 ? launch_fused<FZ_MODE_BF16, 32, 1, 2>(p, smem, grid, st)
                                     : launch_fused<FZ_MODE_BF16, 32, 1, 1>(p, smem, grid, st);
    return launch_fused<FZ_MODE_BF16, 32, 2, 1>(p, smem, grid, st);
}
