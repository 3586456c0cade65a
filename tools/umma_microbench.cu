// Microbenchmark of the hand-shake primitives the fused kernel is built from (sm_100a): tcgen05.mma issue and
// completion time, tcgen05.commit -> mbarrier latency, mbarrier try_wait cost, cross-warp arrive -> wait latency.
// One CTA per launch, clock64 on one SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   -I ev2hands_b200/csrc tools/umma_microbench.cu -o exp/umma_microbench ; run on the GPU box.
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace ev2h;

struct Out { long long v[64]; };

template <int KIND16>
__global__ void __launch_bounds__(64, 1) bench_kernel(Out *out, int n_mma, int N, int whole_warp) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);            // [8]
    uint32_t *tslot = reinterpret_cast<uint32_t *>(smem + 64);
    volatile long long *stamp = reinterpret_cast<volatile long long *>(smem + 128);
    uint8_t *ops = smem + 1024;                                    // operand area, 128 KB: garbage is fine
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) tc::mbar_init(bar + i, 1); tc::fence_mbar_init(); }
    for (int i = threadIdx.x; i < 32 * 1024; i += 64) reinterpret_cast<float *>(ops)[i] = 0.f;
    if (warp == 0) tc::tmem_alloc(tslot, 512);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = *tslot;
    tc::fence_proxy_async();
    __syncthreads();
    const uint32_t idesc = tc::instr_desc(KIND16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, (uint32_t)N);
    const uint32_t desc_hi = tc::smem_desc_hi(128);
    const uint32_t a_lo = tc::smem_desc_lo(tc::smem_u32(ops), 128 * 16), b_lo = tc::smem_desc_lo(tc::smem_u32(ops + 32768), (uint32_t)N * 16);
    long long r[16] = {0};
    if (warp == 0) {
        uint32_t ph = 0;
        // ---- T1: n back-to-back UMMAs, then commit, then wait
        for (int rep = 0; rep < 3; ++rep) {       // last repetition is reported
            __syncwarp();
            const bool me = whole_warp ? tc::elect_one() : lane == 0;
            long long t0 = clock64(), t1 = 0, t2 = 0;
            if (me) {
                for (int i = 0; i < n_mma; ++i) {
                    const uint64_t da = tc::make_desc(a_lo + (i & 3) * 256, desc_hi), db = tc::make_desc(b_lo + (i & 3) * 2 * N, desc_hi);
                    if (KIND16) tc::umma_f16(tmem, da, db, idesc, i > 0); else tc::umma_tf32(tmem, da, db, idesc, i > 0);
                }
                t1 = clock64();
                tc::umma_commit(bar + 0);
                t2 = clock64();
            }
            __syncwarp();
            tc::mbar_wait(bar + 0, ph, 1); ph ^= 1;
            long long t3 = clock64();
            if (me) { r[0] = t1 - t0; r[1] = t2 - t1; r[2] = t3 - t0; }
        }
        // ---- T2: commit with an empty pipe -> barrier completes
        if (lane == 0) {
            long long t0 = clock64();
            tc::umma_commit(bar + 1);
            tc::mbar_wait(bar + 1, 0, 2);
            r[3] = clock64() - t0;
            // ---- T3: try_wait on a completed phase, 16 times
            t0 = clock64();
            for (int i = 0; i < 16; ++i) tc::mbar_wait(bar + 1, 0, 3);
            r[4] = (clock64() - t0) / 16;
            // ---- T5: fences
            t0 = clock64();
            for (int i = 0; i < 16; ++i) tc::fence_proxy_async();
            r[5] = (clock64() - t0) / 16;
            t0 = clock64();
            for (int i = 0; i < 16; ++i) tc::tc_fence_after();
            r[6] = (clock64() - t0) / 16;
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- T4: ping-pong between lane 0 of the two warps: arrive -> partner's wait returns
    if (lane == 0) {
        uint64_t *mine = bar + 2 + warp, *other = bar + 2 + (1 - warp);
        uint32_t ph = 0;
        long long acc = 0;
        for (int i = 0; i < 64; ++i) {
            if (warp == 0) {
                stamp[0] = clock64();
                tc::mbar_arrive(other);
                tc::mbar_wait(mine, ph, 4);
                acc += clock64() - stamp[1];
            } else {
                tc::mbar_wait(mine, ph, 5);
                acc += clock64() - stamp[0];
                stamp[1] = clock64();
                tc::mbar_arrive(other);
            }
            ph ^= 1;
        }
        r[7 + warp] = acc / 64;
    }
    __syncthreads();
    if (lane == 0) for (int i = 0; i < 16; ++i) if (r[i]) out->v[i] = r[i];
    tc::tc_fence_before(); __syncthreads();
    if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

int main() {
    Out *d; cudaMalloc(&d, sizeof(Out));
    const size_t smem = 1024 + 160 * 1024;
    cudaFuncSetAttribute(bench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("# kind N n_mma issuer | issue_cycles commit_issue_cycles total_until_barrier | per-MMA issue | per-MMA total\n");
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {96, 128, 256})
            for (int ww = 0; ww < 2; ++ww)
                for (int n : {1, 2, 4, 8, 16, 32}) {
                    cudaMemset(d, 0, sizeof(Out));
                    if (kind) bench_kernel<1><<<1, 64, smem>>>(d, n, N, ww); else bench_kernel<0><<<1, 64, smem>>>(d, n, N, ww);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    Out h; cudaMemcpy(&h, d, sizeof(Out), cudaMemcpyDeviceToHost);
                    printf("%s N=%3d n=%2d %s | %5lld %4lld %6lld | %6.1f | %6.1f", kind ? "bf16(K16)" : "tf32(K8) ", N, n, ww ? "warp+elect " : "single lane",
                           h.v[0], h.v[1], h.v[2], (double)h.v[0] / n, (double)h.v[2] / n);
                    if (n == 1 && ww == 0 && N == 96)
                        printf("   || empty commit->barrier %lld, try_wait(done) %lld, fence.proxy.async %lld, tcgen05.fence %lld, arrive->wait one way %lld / %lld",
                               h.v[3], h.v[4], h.v[5], h.v[6], h.v[7], h.v[8]);
                    printf("\n");
                }
    return 0;
}
