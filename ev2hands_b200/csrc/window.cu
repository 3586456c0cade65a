// Event-window construction: raw camera events -> the [5, N] point set the encoder consumes.
// (the step right before the set-abstraction path.)
//
// Replaces the per-window numpy code of the reference's dataset classes
//   "stream"  src/Ev2Hands/dataset/evaluation_stream.py:188-215  (ERPCParser.__getitem__)
//   "erpc"    src/Ev2Hands/dataset/erpc.py:178-218, :249          (Ev2HandSDataset.__getitem__)
// which scatter the window's events into 346x260 float32 grids with np.add.at (time sum, positive
// count, negative count, count), list the occupied pixels in np.nonzero (row-major) order with their
// mean time, ["erpc": scale the mean by 1e-6, sort the pixels by it and rebase to the earliest,] draw N
// of them with replacement and normalise x, y, t to [-1, 1].
//
// A window is at most a few thousand events on a 90k-pixel sensor, so nothing here touches a dense
// grid: one CTA per window keeps everything in shared memory.
//   ev2h_window_aggregate_f64   key = pixel << 14 | event number, bitonic sort in shared memory ->
//                               events of a pixel are adjacent AND in stream order; the first thread
//                               of every run adds its events up the way np.add.at does (each sum
//                               is taken in double and rounded to the float32 grid cell, in stream
//                               order), a block scan numbers the occupied pixels.  "erpc" then sorts
//                               (mean time, pixel number) pairs - stable, where the reference's
//                               argsort leaves the order of equal means open.
//   ev2h_window_sample_f32      gathers the drawn records, normalises (IEEE divisions, the reference's
//                               operation order) and writes the channel-first window.
// The draw itself stays with the caller, like the FPS start indices: the reference takes it from
// numpy's global generator (np.random.choice(M, N)) and M is only known after the aggregation.
// HBM traffic per window: 32 B per raw event read once (+ one L2-resident re-read of t, p by the run
// heads), 20 B per occupied pixel written, 20 B per drawn point read, 20 B per point written.
#include "common.cuh"

namespace ev2h {

constexpr int WIN_THREADS = 1024;
constexpr int WIN_IDX_BITS = 14;                    // events per window <= 16384
constexpr int WIN_MAX_EVENTS = 1 << WIN_IDX_BITS;
constexpr int WIN_MAX_SORTED = 4096;                // "erpc" keeps the records in shared memory for the second sort
constexpr uint32_t WIN_PAD_KEY = 0xffffffffu;

// ascending bitonic sort of n = 2^k keys in shared memory by the whole CTA
template <typename T>
__device__ void bitonic_sort(T *keys, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));      // index with the stride bit clear
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const T a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// float -> unsigned with the same order (handles negatives; -0 < +0 is harmless for a sort of means)
__device__ __forceinline__ uint32_t orderable(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <bool ERPC>
__global__ void __launch_bounds__(WIN_THREADS, 2)      // 32 registers: two windows per SM (ncu: the erpc instance held one with 48)
window_aggregate_kernel(const double *__restrict__ events, int64_t row_stride, const int64_t *__restrict__ win_start,
                        const int32_t *__restrict__ win_count, int max_count, int n_sort, int width, int height,
                        float *__restrict__ records, int32_t *__restrict__ n_pixels, int32_t *__restrict__ n_bad) {
    extern __shared__ __align__(16) uint8_t win_smem[];
    uint32_t *keys = reinterpret_cast<uint32_t *>(win_smem);                        // [n_sort]
    float *rec_s = reinterpret_cast<float *>(win_smem + (size_t)n_sort * 4);          // ERPC: [5][n_sort]
    unsigned long long *keys2 = reinterpret_cast<unsigned long long *>(win_smem + (size_t)n_sort * 24);   // ERPC: [n_sort]
    __shared__ int warp_sums[32];
    __shared__ int bad_s, total_s;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = win_count[b];
    const double *ev = events + win_start[b] * row_stride;
    if (tid == 0) bad_s = 0;
    __syncthreads();

    // ---- keys: (pixel, event number); events outside the sensor are dropped and counted ----
    int bad = 0;
    for (int i = tid; i < n_sort; i += WIN_THREADS) {
        uint32_t key = WIN_PAD_KEY;
        if (i < n) {
            const double xd = ev[(int64_t)i * row_stride], yd = ev[(int64_t)i * row_stride + 1];
            const int xi = (int)xd, yi = (int)yd;                 // astype(np.int32): truncation
            if (xd == xd && yd == yd && xi >= 0 && yi >= 0 && xi < width && yi < height) key = ((uint32_t)(yi * width + xi) << WIN_IDX_BITS) | (uint32_t)i;
            else ++bad;
        }
        keys[i] = key;
    }
    if (bad) atomicAdd(&bad_s, bad);
    bitonic_sort(keys, n_sort);

    // ---- runs of equal pixels: heads are numbered by a block scan over contiguous chunks ----
    const int per = n_sort / WIN_THREADS > 0 ? n_sort / WIN_THREADS : 1;      // n_sort is a power of two >= 32
    const int s0 = tid * per, s1 = (s0 + per < n_sort) ? s0 + per : n_sort;
    int heads = 0;
    for (int s = s0; s < s1 && s < n_sort; ++s) {
        const uint32_t k = keys[s];
        if (k != WIN_PAD_KEY && (s == 0 || (keys[s - 1] >> WIN_IDX_BITS) != (k >> WIN_IDX_BITS))) ++heads;
    }
    if (s0 >= n_sort) heads = 0;
    int incl = heads;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        warp_sums[lane] = wi - w;                                  // exclusive
        if (lane == 31) total_s = wi;
    }
    __syncthreads();
    int u = warp_sums[warp] + incl - heads;                        // number of the first head of this thread
    const int M = total_s;
    const double t_first = n > 0 ? ev[2] : 0.0;                    // "stream": events[:, 2] -= events[0, 2]

    if (s0 < n_sort) {
        for (int s = s0; s < s1; ++s) {
            const uint32_t k = keys[s];
            const uint32_t pix = k >> WIN_IDX_BITS;
            if (k == WIN_PAD_KEY || (s > 0 && (keys[s - 1] >> WIN_IDX_BITS) == pix)) continue;
            // np.add.at on float32 cells with float64 values: cell = float32(double(cell) + value), event by event
            float sum_t = 0.f, pos = 0.f, neg = 0.f, cnt = 0.f;
            for (int r = s; r < n_sort; ++r) {
                const uint32_t kr = keys[r];
                if (kr == WIN_PAD_KEY || (kr >> WIN_IDX_BITS) != pix) break;
                const double *row = ev + (int64_t)(kr & (WIN_MAX_EVENTS - 1)) * row_stride;
                double t = row[2];
                if (!ERPC) t = __dsub_rn(t, t_first);
                sum_t = __double2float_rn(__dadd_rn((double)sum_t, t));
                const bool is_pos = row[3] == 1.0;
                pos += is_pos ? 1.f : 0.f;
                neg += is_pos ? 0.f : 1.f;
                cnt += 1.f;
            }
            float t_mean = __fdiv_rn(sum_t, cnt);
            if (ERPC) t_mean = __fmul_rn(t_mean, 1e-6f);           // erpc.py:192, float32 array times a weak scalar
            const float x = (float)(pix % (uint32_t)width), y = (float)(pix / (uint32_t)width);
            if (ERPC) {
                rec_s[u] = x; rec_s[n_sort + u] = y; rec_s[2 * n_sort + u] = t_mean;
                rec_s[3 * n_sort + u] = pos; rec_s[4 * n_sort + u] = neg;
            } else {
                float *o = records + ((int64_t)b * max_count + u) * 5;
                o[0] = x; o[1] = y; o[2] = t_mean; o[3] = pos; o[4] = neg;
            }
            ++u;
        }
    }
    if (tid == 0) { n_pixels[b] = M; n_bad[b] = bad_s; }

    if (ERPC) {
        // ---- pixels by mean time (erpc.py:210-211), pixel order among equals; time rebased to the earliest (:214) ----
        __syncthreads();
        for (int i = tid; i < n_sort; i += WIN_THREADS)
            keys2[i] = i < M ? ((unsigned long long)orderable(rec_s[2 * n_sort + i]) << 32) | (unsigned)i : ~0ull;
        bitonic_sort(keys2, n_sort);
        const float t0 = M > 0 ? rec_s[2 * n_sort + (int)(keys2[0] & 0xffffffffu)] : 0.f;
        for (int j = tid; j < M; j += WIN_THREADS) {
            const int src = (int)(keys2[j] & 0xffffffffu);
            float *o = records + ((int64_t)b * max_count + j) * 5;
            o[0] = rec_s[src]; o[1] = rec_s[n_sort + src]; o[2] = __fsub_rn(rec_s[2 * n_sort + src], t0);
            o[3] = rec_s[3 * n_sort + src]; o[4] = rec_s[4 * n_sort + src];
        }
    }
}

// gather + normalise: pc_normalize (evaluation_stream.py:12-29 / erpc.py:23-39) after the draw
__global__ void __launch_bounds__(256)
window_sample_kernel(const float *__restrict__ records, int max_count, const int32_t *__restrict__ n_pixels,
                     const int64_t *__restrict__ sample_idx, int N, float width, float height,
                     float *__restrict__ out, int32_t *__restrict__ n_bad) {
    __shared__ float red_min[8], red_max[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = n_pixels[b];
    const float *rec = records + (int64_t)b * max_count * 5;
    float *o = out + (int64_t)b * 5 * N;
    float t_min = INFINITY, t_max = -INFINITY;
    int bad = 0;
    for (int j = tid; j < N; j += 256) {
        int64_t src = sample_idx[(int64_t)b * N + j];
        if (src < 0 || src >= M) { ++bad; src = 0; }
        const float *r = rec + src * 5;
        const float x = r[0], y = r[1], t = M > 0 ? r[2] : 0.f;
        // pc[:, 0] /= W; pc[:, 1] /= H; pc[:, :2] = 2 * pc[:, :2] - 1
        o[j] = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(x, width)), 1.f);
        o[N + j] = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(y, height)), 1.f);
        o[2 * N + j] = t;                                       // normalised below, by the same thread
        o[3 * N + j] = r[3];
        o[4 * N + j] = r[4];
        t_min = fminf(t_min, t); t_max = fmaxf(t_max, t);
    }
    if (bad) atomicAdd(n_bad + b, bad);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        t_min = fminf(t_min, __shfl_xor_sync(0xffffffffu, t_min, d));
        t_max = fmaxf(t_max, __shfl_xor_sync(0xffffffffu, t_max, d));
    }
    if (lane == 0) { red_min[warp] = t_min; red_max[warp] = t_max; }
    __syncthreads();
    t_min = red_min[0]; t_max = red_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t_min = fminf(t_min, red_min[w]); t_max = fmaxf(t_max, red_max[w]); }
    const float span = __fsub_rn(t_max, t_min);
    // ts = (2 * ((ts - t_min) / (t_max - t_min))) - 1; a window whose points share one time gives 0/0 = NaN there too
    for (int j = tid; j < N; j += 256)
        o[2 * N + j] = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(__fsub_rn(o[2 * N + j], t_min), span)), 1.f);
}

}  // namespace ev2h

extern "C" int ev2h_window_aggregate_f64(const double *events, int64_t row_stride, const int64_t *win_start,
                                         const int32_t *win_count, int B, int max_count, int width, int height, int mode,
                                         float *records, int32_t *n_pixels, int32_t *n_bad, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(events && win_start && win_count && records && n_pixels && n_bad, "ev2h_window_aggregate_f64: null argument");
    EV2H_REQUIRE(B > 0 && max_count > 0 && row_stride >= 4, "ev2h_window_aggregate_f64: B, max_count must be positive and rows hold x, y, t, p");
    EV2H_REQUIRE(width > 0 && height > 0 && (int64_t)width * height <= (1 << (32 - WIN_IDX_BITS)) - 1,
                 "ev2h_window_aggregate_f64: sensor of %d x %d pixels does not fit the sort key", width, height);
    EV2H_REQUIRE(mode == EV2H_WINDOW_STREAM || mode == EV2H_WINDOW_ERPC, "ev2h_window_aggregate_f64: unknown mode %d", mode);
    const int limit = mode == EV2H_WINDOW_ERPC ? WIN_MAX_SORTED : WIN_MAX_EVENTS;
    if (max_count > limit)
        return fail(EV2H_ERR_UNSUPPORTED, "ev2h_window_aggregate_f64: %d events per window exceed %d (mode %d)", max_count, limit, mode);
    int n_sort = 32;
    while (n_sort < max_count) n_sort <<= 1;
    cudaStream_t st = as_stream(stream);
    cudaError_t e;
    if (mode == EV2H_WINDOW_ERPC) {
        const size_t smem = (size_t)n_sort * 32;          // keys 4 + records 20 + (time, pixel) keys 8 bytes per slot
        e = cudaFuncSetAttribute(window_aggregate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            window_aggregate_kernel<true><<<B, WIN_THREADS, smem, st>>>(events, row_stride, win_start, win_count, max_count, n_sort,
                                                                      width, height, records, n_pixels, n_bad);
    } else {
        const size_t smem = (size_t)n_sort * 4;
        e = cudaFuncSetAttribute(window_aggregate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            window_aggregate_kernel<false><<<B, WIN_THREADS, smem, st>>>(events, row_stride, win_start, win_count, max_count, n_sort,
                                                                       width, height, records, n_pixels, n_bad);
    }
    if (e != cudaSuccess) return fail(EV2H_ERR_CUDA, "ev2h_window_aggregate_f64: smem attribute: %s", cudaGetErrorString(e));
    return check_launch("ev2h_window_aggregate_f64");
}

extern "C" int ev2h_window_sample_f32(const float *records, int max_count, const int32_t *n_pixels, const int64_t *sample_idx,
                                      int B, int N, int width, int height, float *out, int32_t *n_bad, ev2h_stream_t stream) {
    using namespace ev2h;
    EV2H_REQUIRE(records && n_pixels && sample_idx && out && n_bad, "ev2h_window_sample_f32: null argument");
    EV2H_REQUIRE(B > 0 && N > 0 && max_count > 0 && width > 0 && height > 0, "ev2h_window_sample_f32: bad sizes");
    window_sample_kernel<<<B, 256, 0, as_stream(stream)>>>(records, max_count, n_pixels, sample_idx, N, (float)width, (float)height,
                                                          out, n_bad);
    return check_launch("ev2h_window_sample_f32");
}
