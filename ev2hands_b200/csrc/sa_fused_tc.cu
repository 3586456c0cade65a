// Fused grouping + shared MLP + max-pool for one radius scale of a multi-scale set abstraction
// layer, on the tcgen05 tensor cores.  Activations never leave the SM:
//
//   loaders: gather + layer 1 -> smem operand ring -> UMMA (layer 2) -> TMEM accumulator
//        -> epilogue (bias, ReLU, split/convert) -> smem operand ring -> UMMA (layer 3) -> max over K
//
// Replaces, for one (radius, K) scale, the body of PointNetSetAbstractionMsg.forward's loop
// (reference src/Ev2Hands/model/pointnet2_utils.py:243-257): index_points / subtract / cat,
// the three Conv2d(1x1)+BatchNorm2d(eval)+ReLU layers and torch.max over the K neighbours.
//
// A tile is 128 consecutive (centre, neighbour) rows, i.e. 128/K whole groups.  Its work is a fixed
// sequence of "K chunks" (KC input channels x 128 rows of one layer's input, plus that layer's
// weights for those channels).  Chunks flow through two shared-memory rings:
//   A ring  operand rows: layer 2's input is written by the loader warps, layer 3's input by the
//           epilogue warps straight from layer 2's TMEM accumulator
//   B ring  pre-packed weight images, one bulk async copy per chunk (weight-streamer thread)
// and are consumed in order by ONE elected thread issuing tcgen05.mma, which signals completion
// through tcgen05.commit -> mbarrier.
//
// Layer 2 is computed as rows x channels (operand A = activation rows, B = weights): its epilogue
// thread owns a ROW and writes that row's 16-byte operand pieces for layer 3, conflict free.
// Layer 3 is computed TRANSPOSED, channels x rows (A = weights, one 128-channel block per UMMA,
// B = the same activation chunk): its accumulator has an output channel per TMEM lane and the
// tile's rows along the columns, so the max over the K neighbours of a group is a per-thread max
// over columns - no shuffles, no shared memory - and the pooled features leave coalesced.
//
// Layer 1 never runs as a per-row GEMM:
//   gather mode     (<= 8 input channels: sa1, regressor): the loader gathers the 32-byte point
//                   record [features | xyz], subtracts the centre and evaluates layer 1 in exact
//                   fp32 on the CUDA cores (8 FMAs per output channel) while producing the rows.
//   per-point mode  (wide inputs: sa2, 323 channels): layer 1 is linear before its ReLU, so
//                   relu(W1 [f(p); xyz(p) - c(s)] + b1) = relu(P[p] - C[s]) with P = W1 [f; xyz] + b1
//                   per POINT and C = W1_xyz c per centre, both precomputed (32x fewer MACs);
//                   the loader gathers P, subtracts C and applies the ReLU.
//
// Arithmetic modes: BF16 (one kind::f16 UMMA per product), TF32X3 (x = hi + lo, w = hi + lo in tf32, three
// kind::tf32 UMMAs lo*hi + hi*lo + hi*hi: fp32-level accuracy) and MIXED, the default fp32-level mode:
// hi*hi as kind::tf32 and the two correction products x_lo*w_hi + x_hi*w_lo as kind::f16 UMMAs on bf16
// copies.  The corrections are 2^-11 of the result, so bf16's 2^-9 relative error on them is 2^-20 of the
// result (measured end to end: same error against the reference as TF32X3), while tensor time and the
// shared-memory operand reads per product drop by a third (the kernel is shared-memory-bandwidth bound).
// F16X3 (round 2, the default fp32-level mode): x = hi + lo and w = hi + lo as fp16 pairs (tc::split_f16x2, ~22 bits),
// three kind::f16 UMMAs lo*hi + hi*lo + hi*hi.  Against MIXED: three quarters of the tensor time, HALF the operand
// bytes written to and fetched from shared memory (4 instead of 8 bytes per element), and 32-channel chunks fit the
// two-CTAs-per-SM instances, which halves the hand-shakes per tile.  Domain: activations below 65520 (fp16 range) -
// an overflow raises the caller's range flag instead of passing silently.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>

namespace ev2h {

constexpr int FZ_BLOCK_M = 128;
constexpr int FZ_GEMMS = 2;               // layers 2 and 3
constexpr int FZ_MAX_RING = 8;
constexpr int FZ_MAX_PRODUCERS = 4;       // up to two loader groups + up to two sets of epilogue warps
// Template parameters of the kernel:
//   KC  channels per K chunk (32, or 16 to halve the ring footprint so two CTAs share an SM)
//   LG  loader groups of 4 warps;  threads = 32 * (6 + 4 LG): loaders, weight streamer, issuer, 4 epilogue warps
//   OCC CTAs per SM the instance is built for (launch bound; TMEM share is 512 / OCC columns)
// With one CTA per SM nothing else fills the tensor pipe while the issuing thread waits, commits and sets up the
// next chunk (about half of its time per chunk), so those instances run TWO issuer warps that take alternate K
// chunks of a layer: one thread's hand-shakes overlap the other's UMMAs.  (Two CTAs per SM overlap each other.)
__host__ __device__ constexpr int fz_issuers(int occ) { return occ == 1 ? 2 : 1; }
__host__ __device__ constexpr int fz_threads_e(int lg, int ni, int es) { return 32 * (1 + 4 * lg + ni + 4 * es); }

enum { FZ_MODE_BF16 = 0, FZ_MODE_TF32X3 = 1, FZ_MODE_MIXED = 2, FZ_MODE_F16X3 = 3 };
__host__ __device__ constexpr bool fz_two_byte(int mode) { return mode == FZ_MODE_BF16 || mode == FZ_MODE_F16X3; }
constexpr int FZ_W1_MAX = 128;            // widest layer 1 evaluated in the loader warps (weights travel as kernel parameters)

struct FusedParams {
    // geometry of the grouping
    int B, N, S, K;
    const int32_t *idx; int idx_ld, k_off;       // ball-query result [B,S,idx_ld], this scale at k_off
    const float *centres;                        // [B,S,3]
    // optional compacted row list (ev2h_group_compact_i32): padded duplicate neighbours skipped
    const int32_t *rowmap;                       // [n_rows] global point row per compact row, -1 = no point
    const int32_t *blockgroup;                   // [n_rows/8] global group of every 8-row block
    const int32_t *n_rows_dev;                   // device scalar: compact rows of this scale
    // layer 1
    int per_point;                               // 0 = gather + FFMA, 1 = relu(P - C)
    const float *pts8; int D;                    // gather: [B,N,8] rows = [features(D) | xyz | 0]
    // gather: folded layer-1 weights and bias as KERNEL PARAMETERS (constant bank, read through the uniform datapath:
    // LDCU + FFMA2 with a uniform-register operand), so that the loaders' weight reads stay off the shared-memory
    // pipe the tensor cores' operand fetches and the operand stores saturate (round 1: 46-50 % of the LSU wavefronts
    // of the sa1 launches were these warp-uniform LDS.128).  Channel pairs interleaved: w1c[pair][k][2], b1c[ch].
    float w1c[FZ_W1_MAX * 8];
    float b1c[FZ_W1_MAX];
    const float *P; int ld_p, p_col;             // per-point: layer-1 pre-activation per point [B*N, ld_p]
    const float *C; int ld_c, c_col;             // per-point: per-centre offset [B*S, ld_c]
    int c1;                                      // layer-1 width = layer-2 input channels
    // the two tensor-core layers
    int n[FZ_GEMMS];                             // accumulator width (multiple of 16, <= 256)
    int n_chunks[FZ_GEMMS];                      // K chunks of KC input channels
    int k_steps_last[FZ_GEMMS];                  // UMMA K steps in the last chunk
    int tmem_col[FZ_GEMMS];
    int bias_off[FZ_GEMMS];
    const uint8_t *w[FZ_GEMMS];                  // packed weight images (ev2h_tc_pack_weights_kc)
    const float *bias[FZ_GEMMS];
    int mb3;                                     // 128-channel blocks of layer 3 (1 or 2)
    int tmem_cols;                               // TMEM columns to allocate (power of two)
    // output: pooled features of this scale, rows = centres
    float *out; int ld_out, out_col, c_out;
    // rings
    int sa, sb, a_slot_bytes, b_slot_bytes;
    int32_t *range_flag;                         // optional: bit 0 set when an F16X3 operand left the fp16 range
    long long *dbg;     // optional trace buffer (EV2H_FUSED_TRACE builds only, tools/fused_trace.py)
};

struct Ring {
    int slot; uint32_t phase; int size;
    __device__ __forceinline__ void advance() { if (++slot == size) { slot = 0; phase ^= 1; } }
};

#ifdef EV2H_FUSED_TRACE
// timeline trace of CTA 0 (steady-state tiles): (tag << 40 | clock) stored per role from row 300 of the debug
// buffer, 800 entries per role, private counters (plain stores only: no round trip, little perturbation)
#define FZ_TRACE(role, ev, it_, c_) do { if (p.dbg && blockIdx.x == 0 && (it_) >= 6 && (it_) < 9 && trace_n < 799) { \
    p.dbg[4800 + ((role) - 1) * 800 + 1 + trace_n] = ((long long)((role) * 1000000 + (ev) * 10000 + ((it_) % 100) * 100 + (c_)) << 40) | (clock64() & 0xFFFFFFFFFFll); \
    ++trace_n; p.dbg[4800 + ((role) - 1) * 800] = trace_n; } } while (0)
#else
#define FZ_TRACE(role, ev, it_, c_) do { } while (0)
#endif

// A producer's view of the operand ring: the slot of chunk n (absolute index over the CTA's
// lifetime) and, for n >= ring size, the wait for the issuer's grant of that slot.  `bits` holds one
// phase bit per slot, toggled every time this producer consumes a grant.
__device__ __forceinline__ int acquire_slot(uint64_t *my_grants, uint32_t n_abs, int ring, uint32_t &bits, int tag) {
    const int slot = (int)(ring == 2 ? (n_abs & 1u) : ring == 4 ? (n_abs & 3u) : ring == 6 ? (n_abs % 6u) : (n_abs % 3u));   // the operand ring has 2, 3, 4 or 6 slots
    if (n_abs >= (uint32_t)ring) {
        tc::mbar_wait(my_grants + slot, (bits >> slot) & 1u, tag);
        bits ^= 1u << slot;
    }
    return slot;
}

//   ES  sets of 4 epilogue warps (1 or 2).  The epilogue warps are the serial resource of a tile (hand-off of layer 2's
//       accumulator chunk by chunk, then the pool of layer 3): with two sets, set e takes the hand-off chunks c with
//       c % 2 == e and pools 128-channel block e (or, with one block, every other tile).
template <int MODE, int KC, int LG, int OCC, int NI, int ES>
__global__ void __launch_bounds__(fz_threads_e(LG, NI, ES), OCC)
sa_fused_tc_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(128) uint8_t fz_smem[];
    constexpr int THREADS = fz_threads_e(LG, NI, ES);
    constexpr int EB = fz_two_byte(MODE) ? 2 : 4;
    constexpr int PARTS = MODE == FZ_MODE_BF16 ? 1 : 2;
    constexpr int A_PART = FZ_BLOCK_M * KC * EB;         // per precision part: 16 KB (tf32, KC 32) ... 8 KB
    constexpr int A_B16 = FZ_BLOCK_M * KC * 2;           // MIXED: bf16 copy of hi at A_PART, bf16 lo at A_PART + A_B16
    constexpr int NCH = KC * EB / 16;                    // 16-byte operand chunks per row per K chunk
    constexpr int UMMA_K = 32 / EB;
    constexpr int K_STEPS = KC / UMMA_K;
    constexpr int CHUNK_ROWS_BYTES = FZ_BLOCK_M * 16;    // one 16-byte operand chunk for all 128 rows
    static_assert(NCH >= 2 && (!fz_two_byte(MODE) || KC == 32), "unsupported chunk geometry");

    uint8_t *a_ring = fz_smem;
    uint8_t *b_ring = a_ring + (size_t)p.sa * p.a_slot_bytes;
    uint8_t *tail = b_ring + (size_t)p.sb * p.b_slot_bytes;
    // Operand-ring protocol.  a_full[slot]: 128 producer arrivals.  A freed slot is GRANTED by the
    // issuer (tcgen05.commit) directly to the producer that fills it next - loader group 0..LG-1 or
    // the epilogue warps (producer id LG) - on a_grant[producer][slot].  Every producer therefore
    // waits only on barriers whose phase it alone consumes, so a parity wait can never be a phase
    // early or late, however far producers run ahead of or lag behind each other.
    uint64_t *a_full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *a_grant = a_full + FZ_MAX_RING;            // [FZ_MAX_PRODUCERS][FZ_MAX_RING]
    uint64_t *b_full = a_grant + FZ_MAX_PRODUCERS * FZ_MAX_RING;
    uint64_t *b_empty = b_full + FZ_MAX_RING;
    uint64_t *acc_full = b_empty + FZ_MAX_RING;          // [2]
    uint64_t *acc_empty = acc_full + FZ_GEMMS;           // [2]
    uint64_t *init_done = acc_empty + FZ_GEMMS;          // [2] two issuers: chunk 0 of a layer (the accumulator's overwrite) has completed
    uint64_t *turn = init_done + FZ_GEMMS;               // [2] two issuers: turn[i] = issuer i has ISSUED another of its chunks
    // layer 3's "accumulator full" when two loader groups pool ALTERNATE tiles (LG == 2, one 128-channel block): one
    // barrier per tile parity, so that each group consumes every phase of the barrier it waits on (a parity wait that
    // skips a phase returns at once on the phase in between)
    uint64_t *acc_full1_alt = turn + FZ_GEMMS;           // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full1_alt + FZ_GEMMS);
    float *bias_s = reinterpret_cast<float *>(tmem_slot + 4);          // [n0 + n1]; tail offset 480, 16-byte aligned

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef EV2H_FUSED_TRACE
    int trace_n = 0;
#endif
    const bool compact = p.rowmap != nullptr;
    const int64_t M = compact ? (int64_t)__ldg(p.n_rows_dev) : (int64_t)p.B * p.S * p.K;
    const int64_t n_tiles = (M + FZ_BLOCK_M - 1) / FZ_BLOCK_M;
    const int nc0 = p.n_chunks[0], nc1 = p.n_chunks[1];
    const uint32_t Q = (uint32_t)(nc0 + nc1);            // chunks per tile

    if (tid == 0) {
        for (int s = 0; s < p.sa; ++s) {
            tc::mbar_init(a_full + s, 128);
            for (int pr = 0; pr < FZ_MAX_PRODUCERS; ++pr) tc::mbar_init(a_grant + pr * FZ_MAX_RING + s, 1);
        }
        for (int s = 0; s < p.sb; ++s) { tc::mbar_init(b_full + s, 1); tc::mbar_init(b_empty + s, 1); }
        for (int g = 0; g < FZ_GEMMS; ++g) {
            tc::mbar_init(acc_full + g, NI); tc::mbar_init(acc_empty + g, g == 0 ? 128 * ES : (ES == 2 && p.mb3 == 2 ? 256 : 128)); tc::mbar_init(init_done + g, 1);
            tc::mbar_init(turn + g, 1); tc::mbar_init(acc_full1_alt + g, NI);
        }
        tc::fence_mbar_init();
    }
    for (int g = 0; g < FZ_GEMMS; ++g)
        for (int i = tid; i < p.n[g]; i += THREADS) bias_s[p.bias_off[g] + i] = p.bias[g][i];
    uint32_t hmax = 0;              // F16X3: running maximum of the fp16 hi words this thread wrote (range check)
    // Warp roles, lowest to highest warp id = lowest to highest scheduler priority: loaders (work with
    // slack), weight streamer, UMMA issuer, and the epilogue warps, which are the serial bottleneck of a tile.
    constexpr int STREAMER_WARP = 4 * LG, ISSUER_WARP = 4 * LG + 1, EPI_WARP0 = 4 * LG + 1 + NI;
    const int loader_idx = warp;
    const bool is_loader = warp < STREAMER_WARP;
    if (warp == ISSUER_WARP) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // 16 fp32 values of one row (channels j0 .. j0 + 15 of the K chunk, j0 a multiple of 16) -> operand pieces in the
    // K-major, no-swizzle UMMA layout.  Thread = row: consecutive threads write consecutive 16-byte pieces, conflict
    // free.  Producers work in 16-channel pieces so that at most 16 values (not KC) are live per thread.
    auto store_row_16 = [&](uint8_t *st, int r, int j0, const float (&v)[16]) {
        if (MODE == FZ_MODE_MIXED) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int c8 = j0 / 8 + c;
                float4 h0, l0, h1, l1;
                tc::split_tf32x2(v[8 * c], v[8 * c + 1], h0.x, h0.y, l0.x, l0.y);
                tc::split_tf32x2(v[8 * c + 2], v[8 * c + 3], h0.z, h0.w, l0.z, l0.w);
                tc::split_tf32x2(v[8 * c + 4], v[8 * c + 5], h1.x, h1.y, l1.x, l1.y);
                tc::split_tf32x2(v[8 * c + 6], v[8 * c + 7], h1.z, h1.w, l1.z, l1.w);
                *reinterpret_cast<float4 *>(st + (2 * c8) * CHUNK_ROWS_BYTES + r * 16) = h0;
                *reinterpret_cast<float4 *>(st + (2 * c8 + 1) * CHUNK_ROWS_BYTES + r * 16) = h1;
                uint4 xb, lb;
                xb.x = tc::bf16x2(v[8 * c], v[8 * c + 1]); xb.y = tc::bf16x2(v[8 * c + 2], v[8 * c + 3]);
                xb.z = tc::bf16x2(v[8 * c + 4], v[8 * c + 5]); xb.w = tc::bf16x2(v[8 * c + 6], v[8 * c + 7]);
                lb.x = tc::bf16x2(l0.x, l0.y); lb.y = tc::bf16x2(l0.z, l0.w);
                lb.z = tc::bf16x2(l1.x, l1.y); lb.w = tc::bf16x2(l1.z, l1.w);
                *reinterpret_cast<uint4 *>(st + A_PART + c8 * CHUNK_ROWS_BYTES + r * 16) = xb;
                *reinterpret_cast<uint4 *>(st + A_PART + A_B16 + c8 * CHUNK_ROWS_BYTES + r * 16) = lb;
            }
        } else if (MODE == FZ_MODE_F16X3) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {                    // 8 channels = one 16-byte piece of each part
                const int cc = j0 / 8 + c;
                uint4 hb, lb;
                tc::split_f16x2(v[8 * c], v[8 * c + 1], hb.x, lb.x); tc::split_f16x2(v[8 * c + 2], v[8 * c + 3], hb.y, lb.y);
                tc::split_f16x2(v[8 * c + 4], v[8 * c + 5], hb.z, lb.z); tc::split_f16x2(v[8 * c + 6], v[8 * c + 7], hb.w, lb.w);
                hmax = tc::max_u16x2(tc::max_u16x2(hmax, tc::max_u16x2(hb.x, hb.y)), tc::max_u16x2(hb.z, hb.w));
                *reinterpret_cast<uint4 *>(st + cc * CHUNK_ROWS_BYTES + r * 16) = hb;
                *reinterpret_cast<uint4 *>(st + A_PART + cc * CHUNK_ROWS_BYTES + r * 16) = lb;
            }
        } else if (MODE == FZ_MODE_TF32X3) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cc = j0 / 4 + c;
                float4 hi, lo;
                tc::split_tf32x2(v[4 * c], v[4 * c + 1], hi.x, hi.y, lo.x, lo.y);
                tc::split_tf32x2(v[4 * c + 2], v[4 * c + 3], hi.z, hi.w, lo.z, lo.w);
                *reinterpret_cast<float4 *>(st + cc * CHUNK_ROWS_BYTES + r * 16) = hi;
                *reinterpret_cast<float4 *>(st + A_PART + cc * CHUNK_ROWS_BYTES + r * 16) = lo;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int cc = j0 / 8 + c;
                uint4 pk;
                pk.x = tc::bf16x2(v[8 * c], v[8 * c + 1]); pk.y = tc::bf16x2(v[8 * c + 2], v[8 * c + 3]);
                pk.z = tc::bf16x2(v[8 * c + 4], v[8 * c + 5]); pk.w = tc::bf16x2(v[8 * c + 6], v[8 * c + 7]);
                *reinterpret_cast<uint4 *>(st + cc * CHUNK_ROWS_BYTES + r * 16) = pk;
            }
        }
    };

    // ---- max-pool of a tile's layer-3 accumulator (epilogue warps, after the tile's hand-off).  With two epilogue sets and
    // two 128-channel blocks set e pools block e; with one block the sets alternate tiles.  (Round 2 also tried the pool
    // in the LOADER warps, to overlap it with the next tile's hand-off: slower in every variant - DESIGN.md section 4 -
    // and removed.)
    const int pool_sets = ES;                                         // epilogue warp sets that share the pool
    const int pool_groups = (pool_sets == 2 && p.mb3 == 2) ? 2 : 1;   // 2: set g pools 128-channel block g of every tile
    const bool pool_alt = pool_sets == 2 && p.mb3 == 1;               // one block: the two sets pool alternate tiles
    auto pools_tile = [&](uint32_t it_, int set_) { return pool_sets == 1 || pool_groups == 2 || (int)(it_ & 1u) == set_; };
    auto pool_tile = [&](uint32_t it, int64_t tile, int q, int mb_lo, int mb_hi, bool ptrace) {
        // layer 3 (transposed accumulator): lane = output channel, columns = the tile's rows.  The max over the K rows
        // of a group = per-thread max over K columns; bias and ReLU commute with the max (both monotone) and are
        // applied once per pooled value.
        const int64_t m0 = tile * FZ_BLOCK_M;
        const int K = p.K;
        // group ids of the tile's sixteen 8-row blocks (compacted rows), fetched before the accumulator is waited for
        const int my_gid = (compact && lane < 16 && (tile * 16 + lane) * 8 < M) ? __ldg(p.blockgroup + tile * 16 + lane) : -1;   // -1: past the end
        if (pool_alt) tc::mbar_wait(acc_full1_alt + (it & 1), (it >> 1) & 1, 62);
        else tc::mbar_wait(acc_full + 1, it & 1, 61);
        if (ptrace) FZ_TRACE(4, 5, it, 0);
        tc::tc_fence_after();
        const int64_t rows_left = M - m0;                       // rows >= M do not exist (last tile)
        for (int mb = mb_lo; mb < mb_hi; ++mb) {
            const int ch = mb * 128 + q * 32 + lane;
            const uint32_t t_addr = tmem_base + (uint32_t)p.tmem_col[1] + (uint32_t)(mb * FZ_BLOCK_M) + ((uint32_t)(q * 32) << 16);
            const float b = ch < p.n[1] ? bias_s[p.bias_off[1] + ch] : 0.f;
            // compacted rows: a group is a run of 8-row blocks with the same group id (warp uniform).  Groups may straddle
            // tiles, so the first and the last group of a tile are merged into the (zero-initialised) output with an
            // integer atomic max - pooled values are >= 0 after the ReLU, where float order equals integer order - the
            // others are stored.  Dense rows: a group is K consecutive columns.
            int cur = -1;
            bool first_group = true;
            float acc = -INFINITY;
            auto flush = [&](bool edge) {
                if (cur >= 0 && ch < p.c_out) {
                    float *dst = p.out + (int64_t)cur * p.ld_out + p.out_col + ch;
                    const float v = fmaxf(acc + b, 0.f);
                    if (edge) atomicMax(reinterpret_cast<int *>(dst), __float_as_int(v));
                    else *dst = v;
                }
            };
            auto consume16 = [&](const uint32_t (&raw)[16], int col0) {
                if (compact) {
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        const int gid = __shfl_sync(0xffffffffu, my_gid, (col0 >> 3) + blk);
                        const float m8 = fmaxf(fmaxf(fmaxf(__uint_as_float(raw[8 * blk]), __uint_as_float(raw[8 * blk + 1])),
                                                     fmaxf(__uint_as_float(raw[8 * blk + 2]), __uint_as_float(raw[8 * blk + 3]))),
                                               fmaxf(fmaxf(__uint_as_float(raw[8 * blk + 4]), __uint_as_float(raw[8 * blk + 5])),
                                                     fmaxf(__uint_as_float(raw[8 * blk + 6]), __uint_as_float(raw[8 * blk + 7]))));
                        if (gid != cur) {
                            flush(first_group);
                            if (cur >= 0) first_group = false;
                            cur = gid; acc = m8;
                        } else {
                            acc = fmaxf(acc, m8);
                        }
                    }
                } else {
                    float m;
                    if (rows_left >= FZ_BLOCK_M) {           // every tile but possibly the last: no row mask
                        float m0_ = -INFINITY, m1_ = -INFINITY, m2_ = -INFINITY, m3_ = -INFINITY;   // 4 chains for ILP
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            m0_ = fmaxf(m0_, __uint_as_float(raw[j])); m1_ = fmaxf(m1_, __uint_as_float(raw[j + 1]));
                            m2_ = fmaxf(m2_, __uint_as_float(raw[j + 2])); m3_ = fmaxf(m3_, __uint_as_float(raw[j + 3]));
                        }
                        m = fmaxf(fmaxf(m0_, m1_), fmaxf(m2_, m3_));
                    } else {
                        m = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 16; ++j) m = fmaxf(m, (col0 + j) < rows_left ? __uint_as_float(raw[j]) : -INFINITY);
                    }
                    acc = fmaxf(acc, m);
                    if (((col0 + 16) % K) == 0) {            // a group of K rows is complete (K in {32, 64, 128})
                        const int64_t row0 = m0 + col0 + 16 - K;
                        if (row0 < M && ch < p.c_out)
                            p.out[(row0 / K) * (int64_t)p.ld_out + p.out_col + ch] = fmaxf(acc + b, 0.f);
                        acc = -INFINITY;
                    }
                }
            };
            // 16 columns per TMEM load, two buffers: the next load is in flight while this one is reduced
            uint32_t ra[16], rb[16];
            tc::tmem_ld16(t_addr, ra);
#pragma unroll 1
            for (int c0 = 0; c0 < FZ_BLOCK_M; c0 += 32) {
                tc::tmem_ld_wait();
                tc::tmem_ld16(t_addr + c0 + 16, rb);
                consume16(ra, c0);
                tc::tmem_ld_wait();
                if (c0 + 32 < FZ_BLOCK_M) tc::tmem_ld16(t_addr + c0 + 32, ra);
                else if (mb + 1 == mb_hi) {
                    // everything this warp reads of the accumulator is in registers: hand it back before the last
                    // reductions and the global stores
                    tc::tc_fence_before();
                    tc::mbar_arrive(acc_empty + 1);
                }
                consume16(rb, c0 + 16);
            }
            if (compact) flush(true);
        }
        if (ptrace) FZ_TRACE(4, 6, it, 0);
    };

    if (is_loader) {
        // =============================== loaders: layer 2's operand rows ===============================
        const int grp = loader_idx >> 2, wq = loader_idx & 3;
        uint64_t *my_grants = a_grant + grp * FZ_MAX_RING;
        uint32_t bits = 0;
        auto mine = [&](uint32_t it, int kc) { return (int)((it * (uint32_t)nc0 + (uint32_t)kc) % LG) == grp; };

        if (!p.per_point) {
            // ---- gather + layer 1 in exact fp32 on the CUDA cores; thread = row ------------------------
            const int r = wq * 32 + lane;
            // Three-stage software pipeline over the tiles of this CTA, so that no load is waited for where it is issued
            // (the chain row list -> point record -> centre is three dependent L2 round trips):
            //   A  tile t + 2: row-list entries (global point row, group id)
            //   B  tile t + 1: the 32-byte point record and the group's centre, addressed with A's result of a tile ago
            //   C  tile t    : subtract the centre, layer 1
            int a_pt = -1, a_bs = 0;                 // stage A result (compact: point row / group; dense: point index / group)
            float4 b_r0 = make_float4(0.f, 0.f, 0.f, 0.f), b_r1 = b_r0;      // stage B result
            float b_c0 = 0.f, b_c1 = 0.f, b_c2 = 0.f;
            bool b_ok = false;
            auto stage_a = [&](int64_t tile) {
                a_pt = -1; a_bs = 0;
                const int64_t R = tile * FZ_BLOCK_M + r;
                if (tile >= n_tiles || R >= M) return;
                if (compact) {
                    a_pt = __ldg(p.rowmap + R);
                    a_bs = __ldg(p.blockgroup + (R >> 3));
                } else {
                    const int64_t bs = R / p.K;
                    a_bs = (int)bs;
                    a_pt = p.idx[bs * p.idx_ld + p.k_off + (int)(R - bs * p.K)];
                }
            };
            auto stage_b = [&]() {
                b_ok = false;
                int64_t pt_row = a_pt;
                if (!compact) {
                    if (a_pt >= p.N) return;
                    pt_row = (int64_t)(a_bs / p.S) * p.N + a_pt;
                }
                if (a_pt < 0) return;
                b_ok = true;
                const float4 *src = reinterpret_cast<const float4 *>(p.pts8 + pt_row * 8);
                b_r0 = __ldg(src); b_r1 = __ldg(src + 1);
                const float *c = p.centres + (int64_t)a_bs * 3;
                b_c0 = __ldg(c); b_c1 = __ldg(c + 1); b_c2 = __ldg(c + 2);
            };
            float x[8];
            bool valid = false;
            stage_a(blockIdx.x);
            stage_b();
            stage_a((int64_t)blockIdx.x + gridDim.x);
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                // stage C of this tile
                valid = b_ok;
                x[0] = b_r0.x; x[1] = b_r0.y; x[2] = b_r0.z; x[3] = b_r0.w; x[4] = b_r1.x; x[5] = b_r1.y; x[6] = b_r1.z; x[7] = b_r1.w;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {                               // grouped_xyz -= new_xyz (:245)
                    const float ca = ch == p.D ? b_c0 : ch == p.D + 1 ? b_c1 : ch == p.D + 2 ? b_c2 : 0.f;
                    x[ch] = valid ? (ch >= p.D && ch < p.D + 3 ? __fsub_rn(x[ch], ca) : x[ch]) : 0.f;
                }
                uint64_t xx[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) xx[i] = tc::pack2(x[i], x[i]);
                const ulonglong2 *w1p = reinterpret_cast<const ulonglong2 *>(p.w1c);      // constant bank (kernel parameters)
                const uint64_t *b1p = reinterpret_cast<const uint64_t *>(p.b1c);
                stage_b();                                        // tile t + 1: records in flight during this tile
                stage_a(tile + 2 * (int64_t)gridDim.x);           // tile t + 2: row-list entries in flight
                // kc is a compile-time constant in every copy of the body, so every weight address is an immediate offset
                // into the parameter block: LDCU.128 into uniform registers feeding FFMA2 directly (with a run-time kc
                // the compiler falls back to one per-thread LDC.64 per FMA)
#pragma unroll
                for (int kc = 0; kc < FZ_W1_MAX / KC; ++kc) {
                    if (kc >= nc0) break;
                    if (!mine(it, kc)) continue;
                    // layer 1 for 16 channels of the row: exact fp32, packed FMAs, weights from the constant bank
                    auto layer1_16 = [&](int j0, float (&v)[16]) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const int ch = kc * KC + j0 + j;
                            const ulonglong2 *wp = w1p + ch * 2;              // [pair][k][2]: 16 floats per channel pair
                            const ulonglong2 w01 = wp[0], w23 = wp[1], w45 = wp[2], w67 = wp[3];
                            uint64_t acc = b1p[ch >> 1];
                            acc = tc::fma2(w01.x, xx[0], acc); acc = tc::fma2(w01.y, xx[1], acc);
                            acc = tc::fma2(w23.x, xx[2], acc); acc = tc::fma2(w23.y, xx[3], acc);
                            acc = tc::fma2(w45.x, xx[4], acc); acc = tc::fma2(w45.y, xx[5], acc);
                            acc = tc::fma2(w67.x, xx[6], acc); acc = tc::fma2(w67.y, xx[7], acc);
                            float a0, a1;
                            tc::unpack2(acc, a0, a1);
                            v[j] = valid ? fmaxf(a0, 0.f) : 0.f; v[j + 1] = valid ? fmaxf(a1, 0.f) : 0.f;
                        }
                    };
                    float v[16];
                    layer1_16(0, v);                                          // the first piece is ready before the slot is
                    if (tid == 0) FZ_TRACE(1, 1, it, kc);
                    const int slot = acquire_slot(my_grants, it * Q + (uint32_t)kc, p.sa, bits, 10);
                    if (tid == 0) FZ_TRACE(1, 2, it, kc);
                    uint8_t *st = a_ring + (size_t)slot * p.a_slot_bytes;
                    store_row_16(st, r, 0, v);
                    if constexpr (KC == 32) {
                        layer1_16(16, v);
                        store_row_16(st, r, 16, v);
                    }
                    tc::fence_proxy_async();
                    tc::mbar_arrive(a_full + slot);
                    if (tid == 0) FZ_TRACE(1, 3, it, kc);
                }
            }
        } else {
            // ---- per-point mode: relu(P[p] - C[s]); octet lane mapping, P rows prefetched a tile ahead ----
            // 8 consecutive lanes = 8 consecutive rows of ONE 16-byte operand chunk (conflict-free store),
            // the 4 lane octets = 4 adjacent channel quads, so a warp-wide load covers 8 rows x 64 bytes.
            constexpr int QP = KC / 16;             // passes of 4 channel quads per 8-row group
            constexpr int NV = 4 * QP;              // float4 per lane per chunk
            const int l8 = lane & 7, oct = lane >> 3;
            auto rows_of = [&](int64_t tile, int32_t (&p_row)[4], int32_t (&c_row)[4]) {
                const int64_t m0 = tile * FZ_BLOCK_M;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int64_t R = m0 + 32 * wq + h * 8 + l8;
                    p_row[h] = -1; c_row[h] = 0;
                    if (R < M) {
                        if (compact) {
                            p_row[h] = __ldg(p.rowmap + R);
                            c_row[h] = __ldg(p.blockgroup + (R >> 3));
                        } else {
                            const int64_t bs = R / p.K;
                            const int j = (int)(R - bs * p.K);
                            const int64_t b = bs / p.S;
                            const int pt = p.idx[bs * p.idx_ld + p.k_off + j];
                            if (pt >= 0 && pt < p.N) { p_row[h] = (int32_t)(b * p.N + pt); c_row[h] = (int32_t)bs; }
                        }
                    }
                }
            };
            auto load_p = [&](const int32_t (&p_row)[4], int kc, float4 (&v)[NV]) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int k = kc * KC + 4 * (oct + 4 * (i % QP));
                    v[i] = p_row[i / QP] >= 0 ? __ldg(reinterpret_cast<const float4 *>(p.P + (int64_t)p_row[i / QP] * p.ld_p + p.p_col + k))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            auto emit = [&](const int32_t (&p_row)[4], const int32_t (&c_row)[4], uint32_t it, int kc, float4 (&v)[NV]) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (p_row[i / QP] >= 0) {        // C comes from L1: a few rows per tile, shared by all K neighbours
                        const int k = kc * KC + 4 * (oct + 4 * (i % QP));
                        const float4 c = __ldg(reinterpret_cast<const float4 *>(p.C + (int64_t)c_row[i / QP] * p.ld_c + p.c_col + k));
                        v[i] = make_float4(fmaxf(v[i].x - c.x, 0.f), fmaxf(v[i].y - c.y, 0.f), fmaxf(v[i].z - c.z, 0.f), fmaxf(v[i].w - c.w, 0.f));
                    }
                }
                if (tid == 0) FZ_TRACE(1, 1, it, kc);
                const int slot = acquire_slot(my_grants, it * Q + (uint32_t)kc, p.sa, bits, 11);
                if (tid == 0) FZ_TRACE(1, 2, it, kc);
                uint8_t *st = a_ring + (size_t)slot * p.a_slot_bytes;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int row = 32 * wq + (i / QP) * 8 + l8;
                    const int quad = oct + 4 * (i % QP);
                    if (MODE == FZ_MODE_MIXED) {     // a 16-byte bf16 chunk holds two fp32 quads
                        float4 hi, lo;
                        tc::split_tf32x2(v[i].x, v[i].y, hi.x, hi.y, lo.x, lo.y);
                        tc::split_tf32x2(v[i].z, v[i].w, hi.z, hi.w, lo.z, lo.w);
                        *reinterpret_cast<float4 *>(st + quad * CHUNK_ROWS_BYTES + row * 16) = hi;
                        const uint32_t o16 = (quad >> 1) * CHUNK_ROWS_BYTES + row * 16 + (quad & 1) * 8;
                        *reinterpret_cast<uint2 *>(st + A_PART + o16) = make_uint2(tc::bf16x2(v[i].x, v[i].y), tc::bf16x2(v[i].z, v[i].w));
                        *reinterpret_cast<uint2 *>(st + A_PART + A_B16 + o16) = make_uint2(tc::bf16x2(lo.x, lo.y), tc::bf16x2(lo.z, lo.w));
                    } else if (MODE == FZ_MODE_F16X3) {     // a 16-byte fp16 chunk holds two fp32 quads
                        uint2 hb, lb;
                        tc::split_f16x2(v[i].x, v[i].y, hb.x, lb.x);
                        tc::split_f16x2(v[i].z, v[i].w, hb.y, lb.y);
                        hmax = tc::max_u16x2(hmax, tc::max_u16x2(hb.x, hb.y));
                        const uint32_t o16 = (quad >> 1) * CHUNK_ROWS_BYTES + row * 16 + (quad & 1) * 8;
                        *reinterpret_cast<uint2 *>(st + o16) = hb;
                        *reinterpret_cast<uint2 *>(st + A_PART + o16) = lb;
                    } else if (MODE == FZ_MODE_TF32X3) {
                        float4 hi, lo;
                        tc::split_tf32x2(v[i].x, v[i].y, hi.x, hi.y, lo.x, lo.y);
                        tc::split_tf32x2(v[i].z, v[i].w, hi.z, hi.w, lo.z, lo.w);
                        *reinterpret_cast<float4 *>(st + quad * CHUNK_ROWS_BYTES + row * 16) = hi;
                        *reinterpret_cast<float4 *>(st + A_PART + quad * CHUNK_ROWS_BYTES + row * 16) = lo;
                    } else {                         // bf16: a 16-byte operand chunk holds two fp32 quads
                        __nv_bfloat162 q0 = __floats2bfloat162_rn(v[i].x, v[i].y), q1 = __floats2bfloat162_rn(v[i].z, v[i].w);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t *>(&q0); pk.y = *reinterpret_cast<uint32_t *>(&q1);
                        *reinterpret_cast<uint2 *>(st + (quad >> 1) * CHUNK_ROWS_BYTES + row * 16 + (quad & 1) * 8) = pk;
                    }
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(a_full + slot);
                if (tid == 0) FZ_TRACE(1, 3, it, kc);
            };

            int32_t p_cur[4], c_cur[4], p_nxt[4], c_nxt[4];      // row indices fit 32 bits (B * N, B * S < 2^31 is checked on the host)
            // P rows of my first two chunks of the coming tile (only the first one with two epilogue sets: the second
            // buffer costs 32 registers that the 608-thread instance does not have)
            constexpr bool PF2 = ES == 1;
            float4 v0[NV], v1[PF2 ? NV : 1];
            int kc0 = -1, kc1 = -1;
            auto prefetch = [&](int64_t tile, uint32_t it) {
                kc0 = kc1 = -1;
                if (tile >= n_tiles) return;
                rows_of(tile, p_nxt, c_nxt);
                for (int kc = 0; kc < nc0; ++kc) {
                    if (!mine(it, kc)) continue;
                    if (kc0 < 0) { kc0 = kc; load_p(p_nxt, kc, v0); }
                    else if (PF2 && kc1 < 0) {
                        kc1 = kc;
                        if constexpr (PF2) load_p(p_nxt, kc, v1);
                    }
                }
            };
            prefetch(blockIdx.x, 0);
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
#pragma unroll
                for (int h = 0; h < 4; ++h) { p_cur[h] = p_nxt[h]; c_cur[h] = c_nxt[h]; }
                const int u0 = kc0, u1 = kc1;
                for (int kc = 0; kc < nc0; ++kc) {
                    if (!mine(it, kc)) continue;
                    if (kc == u0) emit(p_cur, c_cur, it, kc, v0);
                    else if (PF2 && kc == u1) {
                        if constexpr (PF2) emit(p_cur, c_cur, it, kc, v1);
                    }
                    else { float4 vj[NV]; load_p(p_cur, kc, vj); emit(p_cur, c_cur, it, kc, vj); }
                }
                prefetch(tile + gridDim.x, it + 1);     // my rows are out: fetch the next tile's P rows now
            }
        }
    } else if (warp == STREAMER_WARP) {
        // =============================== weight streamer ===============================
        if (lane == 0) {
            Ring rb{0, 0, p.sb};
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int g = 0; g < FZ_GEMMS; ++g) {
                    const uint32_t w_rows = g == 0 ? (uint32_t)p.n[0] : 128u * (uint32_t)p.mb3;
                    const uint32_t bytes = (uint32_t)PARTS * w_rows * KC * EB;
                    for (int c = 0; c < p.n_chunks[g]; ++c) {
                        tc::mbar_wait(b_empty + rb.slot, rb.phase ^ 1, 20);     // sole producer of this ring: always in step
                        FZ_TRACE(2, 1, (uint32_t)((tile - blockIdx.x) / gridDim.x), g * 50 + c);
                        tc::mbar_arrive_expect_tx(b_full + rb.slot, bytes);
                        tc::bulk_g2s(b_ring + (size_t)rb.slot * p.b_slot_bytes, p.w[g] + (size_t)c * bytes, bytes, b_full + rb.slot);
                        rb.advance();
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= ISSUER_WARP && warp < ISSUER_WARP + NI) {
        // =============================== UMMA issuer(s) ===============================
        // One thread walks the chunk sequence; its per-chunk latency bounds the kernel (every chunk of every tile
        // passes through it), so everything loop invariant is hoisted into registers, barriers are addressed in
        // the shared window directly and the K steps of a full chunk are straight-line code.
        // NI == 2: issuer `me` takes the chunks whose absolute number (over the CTA's lifetime) is me mod 2.  Both
        // rings have an even number of slots there (host side), so a slot - and every phase of its barriers -
        // belongs to one issuer: the two are independent pipelines sharing the tensor pipe, and no parity wait can
        // be a phase off.  The chunks are still ISSUED in sequence - an issuer passes the turn on (turn[me], a plain
        // arrive behind tcgen05.fence::before_thread_sync) once its chunk's UMMAs are issued, and takes it before
        // its next chunk - so the accumulation order, and with it every output bit, is that of one issuer; what
        // overlaps the other's UMMAs is everything else (operand waits, commits, descriptor set-up).  Chunk 0 of a
        // layer overwrites the accumulator: the owner of chunk 1 additionally waits until chunk 0 has COMPLETED
        // (tcgen05.commit -> init_done), and the accumulator goes to the epilogue after both issuers' commits
        // (acc_full counts NI).  Both wait for acc_empty at the start of every layer.
        // The WHOLE warp walks the chunk sequence (waits included) and one elected lane issues: every value that
        // reaches a tcgen05 instruction is then provably warp-uniform (kernel parameters, loop counters, and the two
        // run-time scalars - the TMEM base and the tile count - broadcast with a shuffle), so descriptors live in
        // uniform registers.  Under `if (lane == 0)` the compiler cannot know the branch holds one thread and wraps
        // each UTCHMMA in an ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY loop (~13 dependent instructions per UMMA: the
        // issuing thread, not the tensor pipe, then sets the pace - ~170 cycles per 64-cycle UMMA in the round-2 trace).
        const uint32_t me = (uint32_t)(warp - ISSUER_WARP);
        {
            const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
            const int64_t n_tiles = ((int64_t)__shfl_sync(0xffffffffu, (uint32_t)((M + FZ_BLOCK_M - 1) / FZ_BLOCK_M), 0));
            const uint32_t sa = (uint32_t)p.sa, sb = (uint32_t)p.sb, mb3 = (uint32_t)p.mb3;
            const uint32_t a_full_u = tc::smem_u32(a_full), a_grant_u = tc::smem_u32(a_grant), b_full_u = tc::smem_u32(b_full),
                           b_empty_u = tc::smem_u32(b_empty), acc_full_u = tc::smem_u32(acc_full), acc_empty_u = tc::smem_u32(acc_empty),
                           init_done_u = tc::smem_u32(init_done), turn_u = tc::smem_u32(turn);
            uint32_t own = 0;                                    // chunks this issuer has issued (NI == 2)
            const uint32_t desc_hi = tc::smem_desc_hi(128);
            const uint32_t a_slot_d = (uint32_t)p.a_slot_bytes >> 4, b_slot_d = (uint32_t)p.b_slot_bytes >> 4;
            const uint32_t a_base = tc::smem_desc_lo(tc::smem_u32(a_ring), CHUNK_ROWS_BYTES);
            const uint32_t b_ring_d = (tc::smem_u32(b_ring) >> 4) & 0x3fff;
            constexpr uint32_t a_step = (2 * CHUNK_ROWS_BYTES) >> 4;      // one K step = two 16-byte chunks further along K
            constexpr uint32_t A_PART_D = A_PART >> 4, A_B16_D = A_B16 >> 4;
            const uint32_t nch[2] = {(uint32_t)nc0, (uint32_t)nc1}, ksl[2] = {(uint32_t)p.k_steps_last[0], (uint32_t)p.k_steps_last[1]};
            const uint32_t wr[2] = {(uint32_t)p.n[0], 128u * mb3};
            const uint32_t dt[2] = {tmem_base + (uint32_t)p.tmem_col[0], tmem_base + (uint32_t)p.tmem_col[1]};
            uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0, par = 0;      // par: parity of the chunk's absolute number
            // (tile iteration, position in tile) of the chunk that will reuse the slot being freed: sa chunks ahead
            uint32_t g_it = sa / Q, g_q = sa % Q;
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
#pragma unroll
                for (int g = 0; g < FZ_GEMMS; ++g) {
                    // g == 0: D[rows x n0]      = X[rows x K] * W2[n0 x K]^T      (A = activations, B = weights)
                    // g == 1: D[chan x rows]^T: per 128-channel block  D = W3[128 x K] * X[rows x K]^T  (A = weights, B = activations)
                    const uint32_t n_umma = g == 0 ? wr[0] : (uint32_t)FZ_BLOCK_M;
                    const uint32_t idesc = tc::instr_desc(MODE == FZ_MODE_BF16 ? tc::FMT_BF16 : MODE == FZ_MODE_F16X3 ? tc::FMT_F16 : tc::FMT_TF32,
                                                          FZ_BLOCK_M, n_umma);
                    const uint32_t idesc16 = tc::instr_desc(tc::FMT_BF16, FZ_BLOCK_M, n_umma);     // MIXED: the correction products
                    const uint32_t b_lbo_d = wr[g];                      // (w_rows * 16) >> 4
                    const uint32_t b_step = 2 * b_lbo_d;
                    const uint32_t b_part_d = (wr[g] * KC * EB) >> 4;
                    const uint32_t b_base = b_ring_d | (b_lbo_d << 16);
                    const uint32_t n_mb = g == 0 ? 1u : mb3;
                    tc::mbar_wait_u32(acc_empty_u + 8 * g, (it & 1) ^ 1, 30 + g);      // previous tile's epilogue drained this accumulator
                    if (me == 0 && lane == 0) FZ_TRACE(3, 1, it, g * 50);
                    tc::tc_fence_after();
                    // one product: x = activation operand, w = weight operand (descriptor low words)
                    auto mma = [&](bool k16, uint32_t d, uint32_t x, uint32_t w, uint32_t acc) {
                        const uint64_t dx = tc::make_desc(x, desc_hi), dw = tc::make_desc(w, desc_hi);
                        if (k16) { if (g == 0) tc::umma_f16(d, dx, dw, idesc16, acc); else tc::umma_f16(d, dw, dx, idesc16, acc); }
                        else if (fz_two_byte(MODE)) { if (g == 0) tc::umma_f16(d, dx, dw, idesc, acc); else tc::umma_f16(d, dw, dx, idesc, acc); }
                        else { if (g == 0) tc::umma_tf32(d, dx, dw, idesc, acc); else tc::umma_tf32(d, dw, dx, idesc, acc); }
                    };
                    // the products of K step j (8 tf32 / 16 bf16 channels) of one chunk into accumulator block mb
                    auto k_step = [&](uint32_t x0, uint32_t w0, uint32_t mb, uint32_t j, uint32_t acc) {
                        const uint32_t d = dt[g] + mb * FZ_BLOCK_M, x = x0 + j * a_step, w = w0 + mb * 128u + j * b_step;
                        if (MODE == FZ_MODE_TF32X3 || MODE == FZ_MODE_F16X3) {      // small terms first: x_lo*w_hi + x_hi*w_lo + x_hi*w_hi
                            mma(false, d, x + A_PART_D, w, acc);
                            mma(false, d, x, w + b_part_d, 1u);
                            mma(false, d, x, w, 1u);
                        } else {
                            mma(false, d, x, w, acc);
                        }
                    };
                    // MIXED: correction products of 16-channel step j on the bf16 copies
                    auto k_step16 = [&](uint32_t x0, uint32_t w0, uint32_t mb, uint32_t j) {
                        const uint32_t d = dt[g] + mb * FZ_BLOCK_M;
                        const uint32_t x = x0 + A_PART_D + j * a_step, w = w0 + b_part_d + mb * 128u + j * b_step;
                        mma(true, d, x + A_B16_D, w, 1u);                    // x_lo * w_hi
                        mma(true, d, x, w + (b_part_d >> 1), 1u);            // x_hi * w_lo
                    };
                    for (uint32_t c = 0; c < nch[g]; ++c) {
                        if (NI == 1 || par == me) {
                        tc::mbar_wait2_u32(a_full_u + 8 * a_slot, a_phase, b_full_u + 8 * b_slot, b_phase, 40 + g, 50 + g);
                        if (me == 0 && lane == 0) FZ_TRACE(3, 3, it, g * 50 + c);
                        if (NI == 2) {
                            // my chunk is number 2 * own + me: the other issuer must have issued number 2 * own + me - 1
                            if (me == 1) tc::mbar_wait_u32(turn_u, own & 1u, 37);
                            else if (own > 0) tc::mbar_wait_u32(turn_u + 8, (own - 1u) & 1u, 38);
                            if (c == 1) tc::mbar_wait_u32(init_done_u + 8 * g, it & 1, 35 + g);
                        }
                        tc::tc_fence_after();
                        const uint32_t x0 = a_base + a_slot * a_slot_d, w0 = b_base + b_slot * b_slot_d;
                        const uint32_t acc0 = c > 0 ? 1u : 0u;
                        // the producer that fills this operand slot next: a loader group or the epilogue warps
                        const uint32_t next_prod = g_q < (uint32_t)nc0 ? (LG == 1 ? 0u : (g_it * (uint32_t)nc0 + g_q) % LG)
                                                                       : (uint32_t)LG + (ES == 1 ? 0u : ((g_q - (uint32_t)nc0) & 1u));
                        if (tc::elect_one()) {
                        if (c + 1 < nch[g] || ksl[g] == (uint32_t)K_STEPS) {
#pragma unroll
                            for (uint32_t mb = 0; mb < 2; ++mb) {
                                if (mb < n_mb) {
#pragma unroll
                                    for (uint32_t j = 0; j < (uint32_t)K_STEPS; ++j) k_step(x0, w0, mb, j, j > 0 ? 1u : acc0);
                                    if (MODE == FZ_MODE_MIXED) {
#pragma unroll
                                        for (uint32_t j = 0; j < (uint32_t)K_STEPS / 2; ++j) k_step16(x0, w0, mb, j);
                                    }
                                }
                            }
                        } else {                                             // last, partial chunk
                            const uint32_t ks = ksl[g];
                            for (uint32_t mb = 0; mb < n_mb; ++mb) {
#pragma unroll 1
                                for (uint32_t j = 0; j < ks; ++j) k_step(x0, w0, mb, j, j > 0 ? 1u : acc0);
                                if (MODE == FZ_MODE_MIXED) {
#pragma unroll 1
                                    for (uint32_t j = 0; j < (ks + 1) / 2; ++j) k_step16(x0, w0, mb, j);
                                }
                            }
                        }
                        if (NI == 2) {
                            tc::tc_fence_before();
                            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(turn_u + 8 * me) : "memory");
                        }
                        if (me == 0) FZ_TRACE(3, 6, it, g * 50 + c);
                        tc::umma_commit_u32(a_grant_u + 8 * (next_prod * FZ_MAX_RING + a_slot));
                        tc::umma_commit_u32(b_empty_u + 8 * b_slot);
                        if (NI == 2 && c == 0) tc::umma_commit_u32(init_done_u + 8 * g);
                        if (me == 0) FZ_TRACE(3, 7, it, g * 50 + c);
                        }      // elected lane
                        if (NI == 2) ++own;
                        __syncwarp();
                        }
                        if (++a_slot == sa) { a_slot = 0; a_phase ^= 1; }
                        if (++b_slot == sb) { b_slot = 0; b_phase ^= 1; }
                        if (++g_q == Q) { g_q = 0; ++g_it; }
                        par ^= 1u;
                    }
                    // this issuer's UMMAs into accumulator g are done (all of them with NI == 1)
                    if (tc::elect_one())
                        tc::umma_commit_u32(g == 1 && ES == 2 && mb3 == 1 ? tc::smem_u32(acc_full1_alt) + 8 * (it & 1) : acc_full_u + 8 * g);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue warps ===============================
        const int q = warp & 3, r = q * 32 + lane;       // TMEM lane quadrant of a warp is warp_id % 4
        const int eset = (warp - EPI_WARP0) >> 2;        // epilogue set: hand-off chunks c with c % ES == eset
        uint64_t *my_grants = a_grant + (LG + eset) * FZ_MAX_RING;
        uint32_t bits = 0;
        uint32_t it = 0;
        const bool eprof = warp == EPI_WARP0 && lane == 0;    // the thread that writes this role's trace events
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            {
                // ---- layer 2 activations -> operand chunks of layer 3 (the pool of layer 3 runs in the loader warps) ----
                const uint32_t t_addr = tmem_base + (uint32_t)p.tmem_col[0] + ((uint32_t)(q * 32) << 16);
                const float *bias_g = bias_s + p.bias_off[0];
                tc::mbar_wait(acc_full + 0, it & 1, 60);
                if (eprof) FZ_TRACE(4, 1, it, 0);
                tc::tc_fence_after();
                // bias + ReLU of 16 accumulator columns (columns past n[0] are padding: zero)
                auto act16 = [&](const uint32_t (&raw)[16], int col0, float (&v)[16]) {
#pragma unroll
                    for (int j4 = 0; j4 < 16; j4 += 4) {
                        const ulonglong2 b4 = *reinterpret_cast<const ulonglong2 *>(bias_g + col0 + j4);   // past n[0]: finite, masked below
                        float t0, t1, t2, t3;
                        tc::unpack2(tc::add2(tc::pack2(__uint_as_float(raw[j4 + 0]), __uint_as_float(raw[j4 + 1])), b4.x), t0, t1);
                        tc::unpack2(tc::add2(tc::pack2(__uint_as_float(raw[j4 + 2]), __uint_as_float(raw[j4 + 3])), b4.y), t2, t3);
                        v[j4 + 0] = fmaxf(t0, 0.f); v[j4 + 1] = fmaxf(t1, 0.f);
                        v[j4 + 2] = fmaxf(t2, 0.f); v[j4 + 3] = fmaxf(t3, 0.f);
                    }
                    if (col0 + 16 > p.n[0]) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (col0 + j >= p.n[0]) v[j] = 0.f;
                    }
                };
                // 16 columns per TMEM load, two buffers: the next load is in flight while this one is converted and
                // stored (tcgen05.wait::ld waits for every outstanding load, so exactly one is outstanding at a wait)
                uint32_t ra[16], rb[16];
                float v[16];
                if (eset < nc1) tc::tmem_ld16(t_addr + eset * KC, ra);
                for (int c = eset; c < nc1; c += ES) {
                    tc::tmem_ld_wait();                                       // ra: columns c * KC .. + 15
                    if constexpr (KC == 32) tc::tmem_ld16(t_addr + c * KC + 16, rb);
                    else if (c + ES < nc1) tc::tmem_ld16(t_addr + (c + ES) * KC, rb);
                    act16(ra, c * KC, v);
                    if (eprof) FZ_TRACE(4, 2, it, c);
                    const int slot = acquire_slot(my_grants, it * Q + (uint32_t)(nc0 + c), p.sa, bits, 70);
                    if (eprof) FZ_TRACE(4, 3, it, c);
                    uint8_t *st = a_ring + (size_t)slot * p.a_slot_bytes;
                    store_row_16(st, r, 0, v);
                    if constexpr (KC == 32) {
                        tc::tmem_ld_wait();                                   // rb: columns c * KC + 16 .. + 31
                        if (c + ES < nc1) tc::tmem_ld16(t_addr + (c + ES) * KC, ra);
                        act16(rb, c * KC + 16, v);
                        store_row_16(st, r, 16, v);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) ra[j] = rb[j];
                    }
                    tc::fence_proxy_async();
                    tc::mbar_arrive(a_full + slot);
                    if (eprof) FZ_TRACE(4, 4, it, c);
                }
                tc::tc_fence_before();
                tc::mbar_arrive(acc_empty + 0);
            }
            if (pools_tile(it, eset))
                pool_tile(it, tile, q, pool_groups == 2 ? eset : 0, pool_groups == 2 ? eset + 1 : p.mb3, eprof);
        }
    }

    if (MODE == FZ_MODE_F16X3 && p.range_flag != nullptr && tc::f16x2_overflowed(hmax)) atomicOr(p.range_flag, 1);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ISSUER_WARP) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

static long long *g_fused_dbg = nullptr;

// Which kernel instance serves a layer pair: two CTAs per SM (KC 16 for tf32, one loader group) whenever
// one tile's accumulators fit 256 TMEM columns, so one CTA's epilogues overlap the other's UMMAs.
struct FusedPlan { int kc, occ, tmem_cols, mb3, col[FZ_GEMMS], n[FZ_GEMMS]; bool ok; };

static FusedPlan fused_plan(int mode, const int32_t *cout) {
    FusedPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.n[0] = round_up(cout[0], 16);
    pl.n[1] = round_up(cout[1], 16);
    pl.mb3 = (cout[1] + 127) / 128;
    pl.col[0] = 0;
    pl.col[1] = round_up(pl.n[0], 32);                        // layer-2 accumulator is read in 16/32-column chunks
    const int ext = pl.col[1] + 128 * pl.mb3;                 // layer 3: one 128-column (= 128 rows) block per 128 channels
    pl.ok = cout[0] <= 256 && cout[1] <= 256 && ext <= 512;
    pl.occ = ext <= 256 ? 2 : 1;
    if (const char *e = getenv("EV2H_FUSED_OCC")) { if (e[0] == '1') pl.occ = 1; }     // experiment switch: one CTA per SM, 32-channel chunks
    pl.kc = (pl.occ == 2 && !fz_two_byte(mode)) ? 16 : 32;
    pl.tmem_cols = 32;
    while (pl.tmem_cols < ext) pl.tmem_cols *= 2;
    return pl;
}

template <int MODE, int KC, int LG, int OCC, int NI = fz_issuers(OCC), int ES = 1>
static int launch_fused(const FusedParams &p, size_t smem, unsigned grid, cudaStream_t st) {
    auto k = sa_fused_tc_kernel<MODE, KC, LG, OCC, NI, ES>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(EV2H_ERR_CUDA, "ev2h_sa_msg_fused_tc: smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
    k<<<grid, fz_threads_e(LG, NI, ES), smem, st>>>(p);
    return check_launch("ev2h_sa_msg_fused_tc");
}

}  // namespace ev2h

extern "C" int ev2h_fused_set_debug_buffer(void *buf) { ev2h::g_fused_dbg = (long long *)buf; return 0; }

extern "C" int ev2h_sa_msg_fused_kc(int mode, const int32_t *cout_host) {
    using namespace ev2h;
    if (!cout_host || mode < FZ_MODE_BF16 || mode > FZ_MODE_F16X3) return -1;
    const FusedPlan pl = fused_plan(mode, cout_host);
    return pl.ok ? pl.kc : -1;
}

static int sa_msg_fused_impl(
    const int32_t *idx, int idx_ld, int k_off, const float *centres_rows, int B, int N, int S, int K,
    const float *pts8, int D, const float *first_wt_host, int first_ld, const float *first_bias_host,
    const float *P, int ld_p, int p_col, const float *C, int ld_c, int c_col,
    int c1, const int32_t *cout_host, const void *const *w_packed_host, const float *const *bias_host,
    float *out_rows, int ld_out, int out_col, int mode, int32_t *range_flag,
    const int32_t *rowmap, const int32_t *blockgroup, const int32_t *n_rows_dev, ev2h_stream_t stream);

extern "C" int ev2h_sa_msg_fused_tc(
    const int32_t *idx, int idx_ld, int k_off, const float *centres_rows, int B, int N, int S, int K,
    const float *pts8, int D, const float *first_wt_host, int first_ld, const float *first_bias_host,
    const float *P, int ld_p, int p_col, const float *C, int ld_c, int c_col,
    int c1, const int32_t *cout_host, const void *const *w_packed_host, const float *const *bias_host,
    float *out_rows, int ld_out, int out_col, int mode, int32_t *range_flag, ev2h_stream_t stream) {
    return sa_msg_fused_impl(idx, idx_ld, k_off, centres_rows, B, N, S, K, pts8, D, first_wt_host, first_ld, first_bias_host, P, ld_p, p_col,
                             C, ld_c, c_col, c1, cout_host, w_packed_host, bias_host, out_rows, ld_out, out_col, mode, range_flag,
                             nullptr, nullptr, nullptr, stream);
}

extern "C" int ev2h_sa_msg_fused_compact_tc(
    const int32_t *rowmap, const int32_t *blockgroup, const int32_t *n_rows_dev,
    const float *centres_rows, int B, int N, int S, int K,
    const float *pts8, int D, const float *first_wt_host, int first_ld, const float *first_bias_host,
    const float *P, int ld_p, int p_col, const float *C, int ld_c, int c_col,
    int c1, const int32_t *cout_host, const void *const *w_packed_host, const float *const *bias_host,
    float *out_rows, int ld_out, int out_col, int mode, int32_t *range_flag, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(rowmap && blockgroup && n_rows_dev, "ev2h_sa_msg_fused_compact_tc: null row list");
    return sa_msg_fused_impl(rowmap, K, 0, centres_rows, B, N, S, K, pts8, D, first_wt_host, first_ld, first_bias_host, P, ld_p, p_col,
                             C, ld_c, c_col, c1, cout_host, w_packed_host, bias_host, out_rows, ld_out, out_col, mode, range_flag,
                             rowmap, blockgroup, n_rows_dev, stream);
}

static int sa_msg_fused_impl(
    const int32_t *idx, int idx_ld, int k_off, const float *centres_rows, int B, int N, int S, int K,
    const float *pts8, int D, const float *first_wt_host, int first_ld, const float *first_bias_host,
    const float *P, int ld_p, int p_col, const float *C, int ld_c, int c_col,
    int c1, const int32_t *cout_host, const void *const *w_packed_host, const float *const *bias_host,
    float *out_rows, int ld_out, int out_col, int mode, int32_t *range_flag,
    const int32_t *rowmap, const int32_t *blockgroup, const int32_t *n_rows_dev, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(idx && centres_rows && out_rows && cout_host && w_packed_host && bias_host, "ev2h_sa_msg_fused_tc: null argument");
    EV2H_REQUIRE(B > 0 && N > 0 && S > 0 && k_off >= 0 && k_off + K <= idx_ld && c1 > 0, "ev2h_sa_msg_fused_tc: bad sizes");
    EV2H_REQUIRE((int64_t)B * N < 2147483647LL && (int64_t)B * S < 2147483647LL, "ev2h_sa_msg_fused_tc: B * N and B * S must fit 32 bits");
    EV2H_REQUIRE(mode >= FZ_MODE_BF16 && mode <= FZ_MODE_F16X3, "ev2h_sa_msg_fused_tc: unknown mode %d", mode);
    if (K != 32 && K != 64 && K != 128)
        return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: K=%d (supported: 32, 64, 128)", K);
    const bool per_point = P != nullptr;
    if (!per_point) {
        EV2H_REQUIRE(pts8 && first_wt_host && first_bias_host && first_ld >= c1, "ev2h_sa_msg_fused_tc: gather mode needs pts8 and the folded layer-1 weights (host copies)");
        if (D < 0 || D + 3 > 8) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: gather mode needs D+3 <= 8 input channels (D=%d)", D);
        if (c1 > FZ_W1_MAX) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: gather mode evaluates at most %d layer-1 channels in the loaders (got %d)", FZ_W1_MAX, c1);
    } else {
        EV2H_REQUIRE(C != nullptr && ld_p % 4 == 0 && ld_c % 4 == 0 && p_col % 4 == 0 && c_col % 4 == 0,
                     "ev2h_sa_msg_fused_tc: per-point tables must be float4 addressable");
        if (c1 % 32 != 0) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: per-point mode needs a multiple of 32 layer-1 channels, got %d", c1);
    }
    if (c1 > 256) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: layer-1 width %d > 256", c1);
    const FusedPlan pl = fused_plan(mode, cout_host);
    if (!pl.ok) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: layer widths (%d, %d) exceed 256 or the 512 TMEM columns", cout_host[0], cout_host[1]);
    const int KC = pl.kc;
    const int EB = fz_two_byte(mode) ? 2 : 4, PARTS = mode == FZ_MODE_BF16 ? 1 : 2, UMMA_K = 32 / EB;

    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.N = N; p.S = S; p.K = K; p.idx = idx; p.idx_ld = idx_ld; p.k_off = k_off; p.centres = centres_rows;
    p.rowmap = rowmap; p.blockgroup = blockgroup; p.n_rows_dev = n_rows_dev;
    p.per_point = per_point ? 1 : 0; p.pts8 = pts8; p.D = D; p.range_flag = range_flag;
    if (!per_point) {
        // layer-1 weights into the parameter block: channel pairs interleaved, [pair][k][2], so one 64-bit word feeds one packed FMA
        for (int ch = 0; ch < c1; ++ch) {
            for (int k = 0; k < 8; ++k) p.w1c[(ch >> 1) * 16 + k * 2 + (ch & 1)] = k < D + 3 ? first_wt_host[(size_t)k * first_ld + ch] : 0.f;
            p.b1c[ch] = first_bias_host[ch];
        }
    }
    p.P = P; p.ld_p = ld_p; p.p_col = p_col; p.C = C; p.ld_c = ld_c; p.c_col = c_col; p.c1 = c1;
    p.tmem_cols = pl.tmem_cols; p.mb3 = pl.mb3;
    int boff = 0, max_n = 0;
    for (int g = 0; g < FZ_GEMMS; ++g) {
        const int cin = g == 0 ? c1 : cout_host[0];
        p.n[g] = pl.n[g];
        p.n_chunks[g] = (cin + KC - 1) / KC;
        const int rem = cin - (p.n_chunks[g] - 1) * KC;
        p.k_steps_last[g] = (rem + UMMA_K - 1) / UMMA_K;
        p.tmem_col[g] = pl.col[g];
        p.bias_off[g] = boff; boff += p.n[g];
        p.w[g] = (const uint8_t *)w_packed_host[g]; p.bias[g] = bias_host[g];
        EV2H_REQUIRE(p.w[g] && p.bias[g] && ((uintptr_t)p.w[g] & 15) == 0, "ev2h_sa_msg_fused_tc: layer %d weights null or misaligned", g + 2);
        const int w_rows = g == 0 ? p.n[0] : 128 * pl.mb3;
        if (w_rows > max_n) max_n = w_rows;
    }
    p.out = out_rows; p.ld_out = ld_out; p.out_col = out_col; p.c_out = cout_host[1];
    EV2H_REQUIRE(ld_out >= out_col + p.c_out, "ev2h_sa_msg_fused_tc: ld_out too small");
    p.dbg = g_fused_dbg;

    p.a_slot_bytes = PARTS * FZ_BLOCK_M * KC * EB;
    p.b_slot_bytes = PARTS * max_n * KC * EB;
    const int tail = ((3 + FZ_MAX_PRODUCERS) * FZ_MAX_RING + 5 * FZ_GEMMS) * 8 + 16 + boff * 4;      // barriers, TMEM slot, biases
    int occ = pl.occ;
    int budget = (occ == 2 ? 113 : 227) * 1024 - tail - 512;
    if (occ == 2 && budget < 2 * p.a_slot_bytes + 2 * p.b_slot_bytes) { occ = 1; budget = 227 * 1024 - tail - 512; }
    // at least 2 slots each; a third operand slot when affordable; weights get the rest (prefetched furthest ahead)
    p.sa = 2;
    if (budget - 3 * p.a_slot_bytes >= 3 * p.b_slot_bytes) p.sa = 3;
    // one-CTA-per-SM instances: two issuer warps (round 1) or one (EV2H_FUSED_NI=1; with the elected-lane issue of round 2
    // a single issuer no longer needs ~1.5 k cycles per chunk, and the turn / init_done hand-shakes disappear)
    static const int ni_env = [] { const char *e = getenv("EV2H_FUSED_NI"); return e ? atoi(e) : 0; }();
    const int ni = (occ == 1 && ni_env == 1 && mode == FZ_MODE_F16X3) ? 1 : fz_issuers(occ);
    if (ni == 2) p.sa = budget - 4 * p.a_slot_bytes >= 4 * p.b_slot_bytes ? 4 : 2;      // a slot belongs to one issuer
    p.sb = (budget - p.sa * p.a_slot_bytes) / p.b_slot_bytes;
    if (p.sb > FZ_MAX_RING) p.sb = FZ_MAX_RING;
    if (ni == 2) p.sb &= ~1;
    {   // experiment switches: ring depths (operand ring for two-CTA / one-CTA instances), checked against the budget below
        static const int sa2 = [] { const char *e = getenv("EV2H_FUSED_SA_OCC2"); return e ? atoi(e) : 0; }();
        static const int sa1 = [] { const char *e = getenv("EV2H_FUSED_SA_OCC1"); return e ? atoi(e) : 0; }();
        const int want = occ == 2 ? sa2 : sa1;
        if (want >= 2 && want <= FZ_MAX_RING && (want == 2 || want == 3 || want == 4 || want == 6) && (ni == 1 || want % 2 == 0)) {
            const int sb_new = (budget - want * p.a_slot_bytes) / p.b_slot_bytes;
            if (sb_new >= 2) { p.sa = want; p.sb = sb_new > FZ_MAX_RING ? FZ_MAX_RING : sb_new; if (ni == 2) p.sb &= ~1; }
        }
    }
    if (p.sb < 2) return fail(EV2H_ERR_UNSUPPORTED, "ev2h_sa_msg_fused_tc: rings do not fit in shared memory");
    const size_t smem = (size_t)p.sa * p.a_slot_bytes + (size_t)p.sb * p.b_slot_bytes + tail;

    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = ((int64_t)B * S * K + FZ_BLOCK_M - 1) / FZ_BLOCK_M;
    const int64_t slots = (int64_t)sms * occ;
    const unsigned grid = (unsigned)(n_tiles < slots ? n_tiles : slots);
    cudaStream_t st = as_stream(stream);
    if (mode == FZ_MODE_MIXED) {
        if (KC == 16) return occ == 2 ? launch_fused<FZ_MODE_MIXED, 16, 1, 2>(p, smem, grid, st)
                                      : launch_fused<FZ_MODE_MIXED, 16, 1, 1>(p, smem, grid, st);
        return launch_fused<FZ_MODE_MIXED, 32, 2, 1>(p, smem, grid, st);
    }
    if (mode == FZ_MODE_TF32X3) {
        if (KC == 16) return occ == 2 ? launch_fused<FZ_MODE_TF32X3, 16, 1, 2>(p, smem, grid, st)
                                      : launch_fused<FZ_MODE_TF32X3, 16, 1, 1>(p, smem, grid, st);
        return launch_fused<FZ_MODE_TF32X3, 32, 2, 1>(p, smem, grid, st);
    }
    if (mode == FZ_MODE_F16X3) {
        if (pl.occ == 2) return occ == 2 ? launch_fused<FZ_MODE_F16X3, 32, 1, 2>(p, smem, grid, st)
                                         : launch_fused<FZ_MODE_F16X3, 32, 1, 1>(p, smem, grid, st);
        if (ni == 1) return launch_fused<FZ_MODE_F16X3, 32, 2, 1, 1>(p, smem, grid, st);
        return launch_fused<FZ_MODE_F16X3, 32, 2, 1>(p, smem, grid, st);
    }
    if (pl.occ == 2) return occ == 2 ? launch_fused<FZ_MODE_BF16, 32, 1, 2>(p, smem, grid, st)
                                     : launch_fused<FZ_MODE_BF16, 32, 1, 1>(p, smem, grid, st);
    return launch_fused<FZ_MODE_BF16, 32, 2, 1>(p, smem, grid, st);
}
