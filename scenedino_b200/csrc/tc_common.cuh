// PTX wrappers shared by the tcgen05 kernels (field_tc.cu, field_bin.cu, field_proj.cu): mbarriers, bulk / tensor
// copies, UMMA descriptors, tcgen05.mma / ld / st.  sm_100a only.
#pragma once
#include "common.cuh"

namespace sd {
namespace tcx {

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#if defined(SD_DEBUG_WAIT) || defined(SD_DEBUG_LONGWAIT)
__device__ unsigned int g_wait_timeout[260];   // [0] count, then 60 x (bar, parity, thread, block)
#endif
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One arrival per warp: every lane has finished (and fenced) its part, __syncwarp orders the lanes, lane 0 arrives.
// (All-thread arrives cost ~1600 serialized shared-memory atomics per tile: 30 % of the kernel.)
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.  Plain try_wait spins: a
// suspend-time hint was measured and removed (slow wake-ups put the producer warps in lockstep with the consumers they
// were supposed to run ahead of).
// (The deadlock that first showed up as this trap -- once per ~50 000 queries in profiles/stress_bin.py -- was a group of
// tcgen05 instructions issued twice by a split warp: see the "elected forms" below and DESIGN.md section 3.3.)
#ifndef SD_WAIT_SPINS
#define SD_WAIT_SPINS 2000000u
#endif
// The loop is C++, not a branch inside the asm: with the polling loop hidden in an asm block the compiler lays the code
// behind it out as if the warp could not have diverged in there, and warps that run a role in lockstep (all 32 lanes
// poll, then one elected lane issues tcgen05 / TMA instructions) failed about once per 20 000 launches of the tile
// kernel (profiles/stress_bin.py); as a visible loop with a per-lane exit the same code ran 240 000 launches clean.
#ifdef SD_TRAP_REPORT
__device__ unsigned int *g_trap_report;      // host-mapped buffer (sd_debug_set_trap_buffer): who timed out, readable after the trap
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#ifdef SD_DEBUG_LONGWAIT
#ifndef SD_LONGWAIT_US
#define SD_LONGWAIT_US 1000ull
#endif
    unsigned long long lw_t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(lw_t0));
#endif
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
#ifdef SD_DEBUG_LONGWAIT
        // debug build: waits that COMPLETE after more than ~1 ms are recorded with their duration (bar, parity | 0x80000000,
        // thread, microseconds); waits that never complete (2^31 spins) are recorded with parity as is -- tells a rare
        // long stall from a deadlock
        if (done && spin > 1000u) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            const unsigned long long us = (t1 - lw_t0) / 1000ull;
            if (us > SD_LONGWAIT_US && (threadIdx.x & 31) == 0) {
                const unsigned int k = atomicAdd(&g_wait_timeout[0], 1u);
                if (k < 60) { g_wait_timeout[4 + 4 * k] = bar; g_wait_timeout[5 + 4 * k] = parity | 0x80000000u; g_wait_timeout[6 + 4 * k] = threadIdx.x; g_wait_timeout[7 + 4 * k] = (unsigned int)us; }
            }
        }
        if (!done && spin > 0x7FFFFFF0u) {
            if ((threadIdx.x & 31) == 0) {
                const unsigned int k = atomicAdd(&g_wait_timeout[0], 1u);
                if (k < 60) { g_wait_timeout[4 + 4 * k] = bar; g_wait_timeout[5 + 4 * k] = parity; g_wait_timeout[6 + 4 * k] = threadIdx.x; g_wait_timeout[7 + 4 * k] = blockIdx.x; }
            }
            return;
        }
        continue;
#endif
#ifdef SD_DEBUG_WAIT
        if (!done && spin > (1u << 18)) {   // debug build: record who timed out on what and carry on (results are garbage)
            if ((threadIdx.x & 31) == 0) {                     // one record per warp
                const unsigned int k = atomicAdd(&g_wait_timeout[0], 1u);
                if (k < 60) { g_wait_timeout[4 + 4 * k] = bar; g_wait_timeout[5 + 4 * k] = parity; g_wait_timeout[6 + 4 * k] = threadIdx.x; g_wait_timeout[7 + 4 * k] = blockIdx.x; }
            }
            return;
        }
#else
#ifdef SD_TRAP_REPORT
        if (!done && spin == 2000000u && (threadIdx.x & 31) == 0) {   // one record per stuck warp, then keep waiting so that every
            unsigned int *r = g_trap_report;                          // stuck role gets to report before the trap
            if (r) {
                const unsigned int k = atomicAdd_system(r, 1u);
                if (k < 60) { r[4 + 4 * k] = bar; r[5 + 4 * k] = parity; r[6 + 4 * k] = threadIdx.x; r[7 + 4 * k] = blockIdx.x; }
                __threadfence_system();
            }
        }
        if (!done && spin > SD_WAIT_SPINS) __trap();
#else
        if (!done && spin > SD_WAIT_SPINS) __trap();
#endif
#endif
    }
}
// The wait of a role that a whole warp runs in lockstep (elected issue forms below): every lane polls, then the warp is
// reconverged explicitly -- lanes can leave the polling loop in different iterations, and a uniform-datapath instruction
// (tcgen05.mma / commit, TMA) reached by two parts of a diverged warp would be issued twice.
// (Measured alternatives: polling from one elected lane puts every tcgen05 instruction behind it back into an ELECT loop,
// 2x slower; leaving the loop on a warp vote is as fast as this form.)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 in
// [0,14), LBO unused for swizzled K-major, SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version 1 in
// [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (c_format 1), A/B fp16 (format 0), both K-major.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column = two consecutive K elements), B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// ---- whole-warp ("elected") forms ------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit / TMA copies are uniform-datapath instructions.  Issued under `if (lane == 0)` they sit in
// divergent control flow and ptxas wraps EVERY one of them in an ELECT / BRA.U.ANY loop with its operands rebuilt in
// uniform registers first (~90 cycles of dependent scalar work per instruction: a single-chunk tile's 16 MMAs and 6
// commits were ~2 000 cycles of pure issue, the critical path of the tile kernel).  The issuing roles therefore run as
// whole warps in lockstep -- warp index from warp_uniform() so that the compiler KNOWS it is uniform, every lane polling
// the barriers (mbar_wait_warp) -- and every group of tcgen05 / TMA instructions sits in an `if (elect_one()) { ... }`
// region followed by __syncwarp: the form ptxas recognises as single-threaded (clean UTCHMMA / UTCBAR / UTMALDG, back to
// back), and the form in which a group can never be issued twice.  (An earlier version predicated the instructions on
// elect.sync INSIDE the asm; ptxas drops that predicate -- a uniform-datapath instruction runs once per converged warp
// anyway -- so a warp that was still split behind a polling loop issued the group once per part: a tcgen05.commit that
// arrives twice lets the producers lap the issuer, which then waits for a phase that has gone by.  Seen as a deadlock
// about once per 50 000 queries in profiles/stress_bin.py; 400 000 queries clean with the regions.)
// elect.sync as a C++ predicate: `if (elect_one()) { ... }` in a converged warp is the form ptxas recognises as a
// single-thread region for TMA / bulk copies (clean UTMALDG / UBLKCP; predicating them INSIDE the asm still loops)
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ int warp_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void mbar_expect_tx_e(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// tile store shared -> global (TMA), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const void *tmap, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// one arrival (elected lane) of a converged warp
__device__ __forceinline__ void mbar_arrive_e(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld1_issue(uint32_t taddr, uint32_t &r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
// two fp32 -> packed fp16 with ReLU in the conversion (lo in the low half)
__device__ __forceinline__ uint32_t pack_h2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ __half2 as_h2(uint32_t w) { return *reinterpret_cast<__half2 *>(&w); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }


// MN-major SWIZZLE_128B operand (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units):
// 64 consecutive M/N elements are one 128-byte row, 8 consecutive K rows one 1024-byte atom; SBO = distance
// between 8-row K groups, LBO = distance between 64-element M/N blocks.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// instruction-descriptor bit: the B operand is MN-major (N contiguous) instead of K-major
constexpr uint32_t UMMA_B_MN_MAJOR = 1u << 16;

// 4-D tiled tensor copy global -> shared (TMA), completion on an mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void *tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void *tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// bulk copy shared -> global (contiguous), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// shared load / global store as volatile asm: the compiler keeps them in program order (used to keep a batch of
// loads ahead of the stores that consume them instead of a load-store-load-store chain on one register set)
__device__ __forceinline__ float4 lds128_ordered(uint32_t saddr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ void stg128_ordered(void *p, const float4 &v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// F.softplus(beta=1, threshold=20) with hardware exp2 / log2 (reduced-precision path: abs error ~1e-6)
__device__ __forceinline__ float softplus_fast(float x) {
    if (x > 15.0f) return x;
    const float ex = __expf(x);
    return x < -10.0f ? ex : __logf(1.0f + ex);
}
// positional_encoding.py:13-21 with an approximate reciprocal (the tensor-core path rounds z' to fp16 hi + lo anyway)
__device__ __forceinline__ float znorm_fast(float z, const EncodeParams &e, float inv_denom) {
    float zn;
    if (e.inv_z) {
        float zc = z > SD_EPS ? z : SD_EPS;
        if (z != z) zc = z;
        zn = (__frcp_rn(zc) - e.inv_dmax) * inv_denom;
    } else {
        zn = (z - e.d_min) * inv_denom;
    }
    return 2.0f * zn - 1.0f;
}

}  // namespace tcx
}  // namespace sd
