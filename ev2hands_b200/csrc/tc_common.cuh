// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, bulk async copy,
// tcgen05 (TMEM allocation, UMMA issue / commit, TMEM loads) and the shared-memory / instruction
// descriptors.  Thin inline-PTX wrappers; bit layouts follow the PTX ISA tcgen05 descriptor
// tables (also documented in CUTLASS's cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ev2h {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel with a message, not as a hung GPU.
// `tag` identifies the wait site in the message.  (try_wait suspends for a hardware time slice per
// call, so the bound below is several seconds - far beyond any legitimate wait.)
static __device__ __noinline__ void mbar_timeout(int tag, uint32_t parity) {
    printf("ev2h: mbarrier wait timed out: tag=%d parity=%u block=%d thread=%d\n", tag, parity, (int)blockIdx.x,
           (int)threadIdx.x);
    const long long t0 = clock64();
    while (clock64() - t0 < 1000000000LL) {   // let the other stuck waiters report before the trap kills the grid
    }
    __trap();
}
// Probe with a suspend-time hint: the warp sleeps inside the instruction (no issue slots used) and is
// woken by the barrier's phase change, so a long hint costs no latency.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
#ifndef EV2H_WAIT_HINT_NS
#define EV2H_WAIT_HINT_NS 20000u      // suspend-time hint of the waits (experiment knob: -DEV2H_WAIT_HINT_NS=...)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int tag = 0) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(bar, parity, EV2H_WAIT_HINT_NS)) {
        if ((++spins & 15u) == 0 && clock64() - t0 > 4000000000LL) mbar_timeout(tag, parity);   // ~2 s
    }
}

// The same wait / commit on a shared-window address (no generic -> shared conversion per call)
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity, int tag = 0) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(EV2H_WAIT_HINT_NS) : "memory");
        if (ok) return;
        if ((++spins & 15u) == 0 && clock64() - t0 > 4000000000LL) mbar_timeout(tag, parity);   // ~2 s
    }
}
// Two barriers at once: both probes are in flight together (a probe costs ~80 cycles even when the phase is already
// complete, profiles/r01_umma_handshake_microbench.txt), then whichever is still pending is waited for.
__device__ __forceinline__ void mbar_wait2_u32(uint32_t bar_a, uint32_t parity_a, uint32_t bar_b, uint32_t parity_b, int tag_a, int tag_b) {
    uint32_t ok_a, ok_b;
    asm volatile("{\n.reg .pred p, q;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5;\n"
                 "selp.u32 %0, 1, 0, p;\nselp.u32 %1, 1, 0, q;\n}\n"
                 : "=r"(ok_a), "=r"(ok_b) : "r"(bar_a), "r"(parity_a), "r"(bar_b), "r"(parity_b) : "memory");
    if (!ok_a) mbar_wait_u32(bar_a, parity_a, tag_a);
    if (!ok_b) mbar_wait_u32(bar_b, parity_b, tag_b);
}
__device__ __forceinline__ void umma_commit_u32(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// generic-proxy writes (st.shared) -> visible to the async proxy (UMMA operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t n_cols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(n_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t n_cols) {     // same warp as alloc
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// all previously issued UMMAs of this thread done -> one arrival on the mbarrier
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (quadrant*32 + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------
// K-major operand without swizzle ("interleaved" canonical layout): core matrix = 8 rows x 16 bytes
// stored as 128 contiguous bytes; SBO = byte distance between consecutive 8-row groups,
// LBO = byte distance between the two 16-byte K chunks one UMMA consumes.
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version for sm_100
    return d;                      // base offset 0, layout type 0 = no swizzle
}

// Same descriptor from pre-shifted parts, so the per-K-step update is one 32-bit add on warp-uniform
// values: lo = (addr >> 4) | ((lbo >> 4) << 16), hi = (sbo >> 4) | version.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3fff) | (((lbo_bytes >> 4) & 0x3fff) << 16);
}
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3fff) | (1u << 14); }
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// One lane of a converged warp; lets the compiler keep the surrounding code warp-uniform
// (descriptor arithmetic on the uniform datapath, no per-instruction R2UR moves).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

enum : uint32_t { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };

// instruction descriptor: fp32 accumulate, both operands K-major, dense
__host__ __device__ constexpr uint32_t instr_desc(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread.  kind::tf32 reads fp32 words and uses
// their top 19 bits; kind::f16 reads bf16/f16.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// fp32 -> (hi, lo) with hi exactly representable in tf32 (round to nearest) and lo = x - hi
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

// Packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 issue two IEEE operations per lane per instruction; each
// half rounds exactly like the scalar instruction, so results are bit-identical to fmaf / __fadd_rn).
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// two fp32 -> packed bf16x2 (round to nearest even), first value in the low half
__device__ __forceinline__ uint32_t bf16x2(float lo_half, float hi_half) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_half), "f"(lo_half));
    return d;
}
// two fp32 -> (hi, lo) pairs; the subtraction is one packed instruction
__device__ __forceinline__ void split_tf32x2(float x0, float x1, float &h0, float &h1, float &l0, float &l1) {
    uint32_t a, b;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(x0));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(x1));
    h0 = __uint_as_float(a); h1 = __uint_as_float(b);
    unpack2(sub2(pack2(x0, x1), pack2(h0, h1)), l0, l1);
}

// fp16 split: x = hi + lo with hi = fp16(x) (11 significant bits) and lo = fp16(x - hi) (the next 11): the pair
// carries ~22 bits, and hi*hi + hi*lo + lo*hi (three kind::f16 UMMAs, fp32 accumulate) is an fp32-level product at
// three quarters of the tensor time and half the operand bytes of the tf32 + 2 bf16 split.  x - hi is exact in fp32.
// Domain: |x| < 65520 (fp16 range); values below 2^-14 lose relative, not absolute, precision (lo goes subnormal).
// Two values per call; the first one lands in the low half of each packed word.
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t &h, uint32_t &l) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    float f0, f1;
    asm("{\n.reg .f16 a, b;\nmov.b32 {a, b}, %2;\ncvt.f32.f16 %0, a;\ncvt.f32.f16 %1, b;\n}\n" : "=f"(f0), "=f"(f1) : "r"(h));
    float r0, r1;
    unpack2(sub2(pack2(x0, x1), pack2(f0, f1)), r0, r1);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
}
// running maximum of packed non-negative fp16 pairs (bit patterns of values >= +0 order like unsigned integers)
__device__ __forceinline__ uint32_t max_u16x2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
// true if either half is +inf / NaN (exponent all ones)
__device__ __forceinline__ bool f16x2_overflowed(uint32_t m) { return ((m & 0x7c00u) == 0x7c00u) || ((m & 0x7c000000u) == 0x7c000000u); }

}  // namespace tc
}  // namespace ev2h
