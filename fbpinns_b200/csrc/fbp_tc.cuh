// tcgen05 ("tensor") kernel family: the hidden-layer GEMMs of the jet propagation on the 5th-generation tensor cores
// of sm_100a, at FP32-equivalent accuracy through the 3xTF32 split (x = hi + lo, x*w ~ lo*w_hi + hi*w_lo + hi*w_hi,
// every part rounded to TF32, FP32 accumulation in tensor memory).
//
// Why: ncu on the FFMA2 kernels (profiles/r1e_*) shows the FP32 FMA pipe as the binding resource of the GEMM phases;
// the north star allows tensor-core MMA on the hidden layers in exactly that case, provided the tolerance holds.
//
// Shape of the mapping (plans with H = 32, two hidden layers, C <= 5 jet components):
//   * tile = 128 pairs of one subdomain = the 128 lanes of tensor memory: ROW p of every operand is point p;
//   * 256 threads: thread (g, p) owns point p and the 16 hidden units [16 g, 16 g + 16)  (g = warpgroup), so that
//     warp w touches only the TMEM lanes 32 (w % 4) .. +31 it is allowed to;
//   * A operands (activations, 128 x 32 per jet component) never touch shared memory: the thread that computed a row
//     writes its hi and lo parts straight into tensor memory (tcgen05.st), and the MMAs read A from TMEM (".ts" form);
//   * B operands (the subdomain's 32 x 32 matrix, hi and lo) are staged once per work item in shared memory in the
//     canonical no-swizzle K-major core-matrix layout and addressed through UMMA shared-memory descriptors;
//   * accumulators D[c] (128 x 32 per component) live in TMEM; one elected thread issues the MMAs
//     (M = 128, N = 16, K = 8 per instruction; 12 per component and unit half) and commits them to an mbarrier per
//     unit half, so warpgroup 0 starts its tanh-jet epilogue while the tensor core still works on the second half;
//   * the epilogue reads the accumulators back row-wise (tcgen05.ld 32x32b: thread = point, registers = units), so
//     the tanh jets, the output layer (a 32-long dot product per component) and the window/Leibniz product are all
//     thread-local; the two unit halves meet once per tile through 5 floats per point in shared memory.
//
// TMEM column map (one allocation of 512 columns per CTA, hence one CTA per SM):
//     [0, 32 C)  A hi   column c*32 + k          [32 C, 64 C)  A lo          [64 C, 96 C)  D   column c*32 + j
//
// Status: validated on a B200 (profiles/r1f_tc_bringup.md: 3e-7 from the FFMA2 kernel, 1.995 vs 2.368 ms on cfg 5) and
// used by mode 0 (auto) for that instance; the other instances on request (fbp_plan_set_kernel(plan, 3)).
// tests/test_gpu_tc.py (FBP_TC_TESTS=1) holds the oracle-level tests of the whole family.
#pragma once
#include "fbp_fast.cuh"

namespace fbptc {

// Per-CTA timeline of the tensor kernels (builds with -DFBP_BLOCK_TRACE only, tests/tools/item_cost_trace.py): thread 0 of every
// CTA records [8 * block + k] = 0 SM id, 1 entry, 2 prologue done, 3 its tile loop done, 4 every warp done, 5 exit (low 32 bits of
// the SM's cycle counter), 6 tiles, 7 exit on the global nanosecond timer.  Nothing of it exists in the product build.
#ifdef FBP_BLOCK_TRACE
__device__ uint32_t* fbp_blk_trace = nullptr;
__device__ __forceinline__ void blk_stamp(int k) {
#ifdef FBP_BLK_MASK
    if (!((FBP_BLK_MASK >> k) & 1)) return;
#endif
    if (threadIdx.x == 0 && fbp_blk_trace) fbp_blk_trace[8 * blockIdx.x + k] = (uint32_t)clock64();
}
__device__ __forceinline__ void blk_stamp_entry(int ntiles) {
    if (threadIdx.x == 0 && fbp_blk_trace) {
        uint32_t smid;
        uint64_t ns;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        fbp_blk_trace[8 * blockIdx.x + 0] = smid;
        fbp_blk_trace[8 * blockIdx.x + 6] = (uint32_t)ntiles;
        fbp_blk_trace[8 * blockIdx.x + 7] = (uint32_t)ns;
    }
}
#ifdef FBP_BLOCK_TRACE3
#define BLK_STAMP3 blk_stamp(3)
#else
#define BLK_STAMP3
#endif
#else
#define BLK_STAMP3
__device__ __forceinline__ void blk_stamp(int) {}
__device__ __forceinline__ void blk_stamp_entry(int) {}
#endif

constexpr int TP = 128;           // pairs per tile = TMEM lanes
constexpr int NT = 256;           // threads: 2 warpgroups x 128 point rows
constexpr int H = 32;
constexpr uint32_t TMEM_COLS = 512;

// ---- UMMA descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp of CUTLASS, SmemDescriptor / InstrDescriptor) ------
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: core matrices of 8 rows x 16 bytes (= 4 tf32) stored as 128
// contiguous bytes; LBO = byte step between the two core matrices an MMA reads along K (K = 8 tf32), SBO = byte step
// between groups of 8 rows.
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);            // [0,14)  start address >> 4
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;      // [16,30) leading byte offset >> 4
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;      // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                 // [46,48) descriptor version 1 (Blackwell)
    // base offset [49,52) = 0, lbo mode [52] = 0, layout type [61,64) = 0 (SWIZZLE_NONE)
    return d;
}
// descriptor of the same layout `bytes` further on (bytes % 16 == 0, same 256 KB window): ONE 64-bit add of a constant on the
// uniform datapath, where make_smem_desc of the new address costs the issuing warp ~6 dependent instructions per MMA
__host__ __device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// instruction descriptor of tcgen05.mma.kind::tf32, FP32 accumulate; a_mn / b_mn = 1: the operand is MN-major (the
// MN-major form failed its hardware self-test in round 2 and is not used; every operand of the family is K-major)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4)                    // [4,6)   D format F32
         | (2u << 7)                    // [7,10)  A format TF32
         | (2u << 10)                   // [10,13) B format TF32
         | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16)      // [15] A major, [16] B major (0 = K-major)
         | ((uint32_t)(N >> 3) << 17)   // [17,23) N >> 3
         | ((uint32_t)(M >> 4) << 24);  // [24,29) M >> 4
}

// Operand images of the tensor-core weight gradient (fbp_tc_bwd2.cuh): [32 rows][32 points of a quarter tile], K-major
// (K = point), no swizzle: 8 x 4 core matrices of 128 contiguous bytes; the 8 cores of a row group are GK_LBO = 144
// bytes apart along K (16 bytes of padding: the 32 lanes of a warp, which hold the 32 points of one row, then store to
// 32 different banks), row groups GK_SBO = 8 * 144 bytes apart.  One MMA (K = 8) reads two cores: next K step = +288 B.
constexpr uint32_t GK_LBO = 144, GK_SBO = 8 * GK_LBO;
constexpr int GK_IMG = 4 * GK_SBO / 4;                       // floats per 32-row image (1152)
__host__ __device__ constexpr int gk_index(int row, int k, uint32_t lbo = GK_LBO) {
    return (row >> 3) * (int)(8 * lbo / 4) + (k >> 2) * (int)(lbo / 4) + (row & 7) * 4 + (k & 3);
}

// canonical K-major no-swizzle placement of element (row n, column k) of a [rows][32] tf32 matrix, in floats:
// 8 x 4 core matrices of 32 floats, the 8 cores of a row group contiguous along K (LBO = 128 B), row groups 1024 B apart
__host__ __device__ constexpr int bcore_index(int n, int k) { return ((n >> 3) * 8 + (k >> 2)) * 32 + (n & 7) * 4 + (k & 3); }
constexpr uint32_t B_LBO = 128, B_SBO = 1024;

// ---- TF32 split ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rn(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - __uint_as_float(hi));
}

// 2-instruction split for the activation operands: hi = x with the low 13 mantissa bits cleared (exactly what the tensor
// core would read anyway), lo = x - hi (exact) passed as raw FP32 bits (the tensor core ignores its low 13 bits).
// Measured with fbp_tc_selftest variant 24: 2.8e-7 of |A||W| (5-instruction rounded split: 9.5e-8).
__device__ __forceinline__ void tf32_split_fast(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {       // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// D[tmem] (+)= A[smem descriptor] * B[smem descriptor]
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]; one thread issues for the CTA
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier receives one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// warp index as a value the compiler knows to be warp-uniform, and the single-lane election used for MMA issue:
// with both, ptxas keeps the MMA operands in uniform registers instead of wrapping every tcgen05.mma in a
// waterfall loop (ELECT/PLOP3/BRA per instruction when the guard is `tid == 0`)
__device__ __forceinline__ int warp_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// mbarrier wait that turns a lost arrival into a trap instead of a hung GPU
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long t_start = 0;
    for (int spin = 0;; ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (spin == 64) t_start = clock64();
        if (spin > 64 && clock64() - t_start > 4000000000ll) __trap();      // ~2 s
    }
}

// Stage a [32][32] matrix B[n][k] = src[n * sn + k * sk] (hi and lo TF32 parts) in the canonical layout.
__device__ __forceinline__ void stage_b(float* bhi, float* blo, const float* __restrict__ src, int sn, int sk, int tid,
                                        int nthreads = NT) {
    for (int i = tid; i < H * H; i += nthreads) {
        const int n = i >> 5, k = i & 31;
        uint32_t hi, lo;
        tf32_split(src[n * sn + k * sk], hi, lo);
        bhi[bcore_index(n, k)] = __uint_as_float(hi);
        blo[bcore_index(n, k)] = __uint_as_float(lo);
    }
}

// The 3xTF32 product of one 128 x 32 A block (columns a_hi.. / a_lo.. of TMEM) with the N = 16 rows [16 nh, 16 nh + 16)
// of B, into the 16 accumulator columns at d_tmem: 12 MMAs (small terms first).
__device__ __forceinline__ void issue_gemm_half(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                int nh) {
    constexpr uint32_t idesc = make_idesc_tf32(128, 16);
    const uint64_t row_off = (uint64_t)((nh * 2 * B_SBO) >> 4);
#pragma unroll
    for (int pr = 0; pr < 3; ++pr) {
        const uint32_t acol = (pr == 0) ? a_lo : a_hi;                 // lo*hi, hi*lo, hi*hi
        const uint64_t bd = ((pr == 1) ? b_lo : b_hi) + row_off;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            mma_tf32_ts(d_tmem, acol + ks * 8, bd + (uint64_t)((ks * 2 * B_LBO) >> 4), idesc, (pr | ks) != 0 ? 1u : 0u);
    }
}

// hi = the value itself (the tensor core reads only the upper 19 bits), lo = what that read drops
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// 3xTF32 product of one 128 x 32 A block in tensor memory with N rows of B (canonical layout; N = 32: one matrix, N = 64:
// two matrices stored back to back), small terms first; `fresh`: the first MMA overwrites the accumulator.  One
// instruction per K step and product (12 in all).
template <int N>
__device__ __forceinline__ void issue_gemm_acc(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, bool fresh) {
    constexpr uint32_t idesc = make_idesc_tf32(128, N);
#pragma unroll
    for (int pr = 0; pr < 3; ++pr) {
        const uint32_t acol = (pr == 0) ? a_lo : a_hi;
        const uint64_t bd = (pr == 1) ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            mma_tf32_ts(d_tmem, acol + ks * 8, bd + (uint64_t)((ks * 2 * B_LBO) >> 4), idesc, (fresh && pr == 0 && ks == 0) ? 0u : 1u);
    }
}

// The first layer is linear in the normalised point, so its jets factor per unit k: h1_0 = t, h1_s = g kappa_s[k],
// h1_ss = -2 kappa_s[k]^2 (t g) with kappa_s[k] = W0[k][axis s] / sd.  The hidden GEMM a2 = h1 W1^T therefore needs only the
// A operands (t, g kappa_s, t g) against W1 and, for the second-order components, W1 diag(-2 kappa_s^2):
//   CA1 = A operands,  NB1 = B variants;  component c of a2 uses A operand fac_ia(c) and B variant fac_vb(c).
template <class CF>
struct Factored {
    static constexpr int CA1 = 1 + CF::NS + (CF::NA2 > 0 ? 1 : 0);
    static constexpr int NB1 = 1 + CF::NA2;
    __host__ __device__ static constexpr int ia(int c) {
        return c == 0 ? 0 : (c <= 2 * CF::NA2 ? (((c - 1) & 1) ? CA1 - 1 : 1 + ((c - 1) >> 1)) : 1 + CF::NA2 + (c - 1 - 2 * CF::NA2));
    }
    __host__ __device__ static constexpr int vb(int c) { return (c > 0 && c <= 2 * CF::NA2 && ((c - 1) & 1)) ? 1 + ((c - 1) >> 1) : 0; }
};
// Stage W1 and its scaled variants (hi, lo) as B operands B1_v[n = j][k] = W1[j][k] sc_v[k]: [NB1][hi, lo][H*H] floats at b1.
template <class CF>
__device__ __forceinline__ void stage_b1_variants(float* b1, const float* __restrict__ w1, const float* sm, int tid, int nthreads) {
    for (int i = tid; i < H * H; i += nthreads) {
        const int j = i >> 5, k = i & 31;
        const float w = w1[i];
#pragma unroll
        for (int v = 0; v < Factored<CF>::NB1; ++v) {
            float sc = 1.0f;
            if (v > 0) {
                const float kap = sm[CF::SM_W0D + (v - 1) * H + k];
                sc = -2.0f * kap * kap;
            }
            uint32_t hi, lo;
            tf32_split(w * sc, hi, lo);
            b1[(2 * v) * H * H + bcore_index(j, k)] = __uint_as_float(hi);
            b1[(2 * v + 1) * H * H + bcore_index(j, k)] = __uint_as_float(lo);
        }
    }
}

// shared-memory layout of the forward kernel (floats, after CF::SM_PARAMS rounded up to 32)
template <class CF, int NWG>
struct FwdSmem {
    static constexpr int OFF_BHI = (CF::SM_PARAMS + 31) & ~31;
    static constexpr int OFF_BLO = OFF_BHI + H * H;
    static constexpr int OFF_EXCH = OFF_BLO + H * H;                    // [NWG-1][C][TP] partial output dots of warpgroups 1..
    static constexpr int OFF_OUT = OFF_EXCH + (NWG - 1) * CF::C * TP;    // [TP][C] tile output in external component order
    static constexpr int FLOATS = OFF_OUT + CF::C * TP;
};

// =====================================================================================================
// forward
// =====================================================================================================
// NWG warpgroups of 128 point rows; warpgroup g owns the hidden units [UPT g, UPT (g + 1)), UPT = 32 / NWG.
// a.dbg (timing experiments only, results are then wrong): 2 = no MMA issue / waits, 4 = no layer 0 / A stores,
// 8 = no tanh jets in the epilogue.
template <class CF, int NWG>
__global__ void __launch_bounds__(128 * NWG, 1) tc_forward_kernel(FastArgs a) {
    static_assert(CF::H == 32 && CF::NHID == 2, "tensor family: H = 32, two hidden layers");
    static_assert(3 * CF::C * 32 <= (int)TMEM_COLS, "A hi, A lo and D must fit the 512 TMEM columns");
    static_assert(NWG == 2 || NWG == 4, "2 or 4 warpgroups");
    constexpr int NT = 128 * NWG, UPT = 32 / NWG, NCH = UPT / 8;
    constexpr int C = CF::C, NS = CF::NS, NA2 = CF::NA2, NA1 = CF::NA1;
    constexpr uint32_t COL_AHI = 0, COL_ALO = C * 32, COL_D = 2 * C * 32;
    using L = FwdSmem<CF, NWG>;
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t mma_bar[2];
    __shared__ uint32_t tmem_slot;
    float* bhi = sm + L::OFF_BHI;
    float* blo = sm + L::OFF_BLO;
    float* exch = sm + L::OFF_EXCH;
    float* outN = sm + L::OFF_OUT;

    const int tid = threadIdx.x, warp = warp_uniform();
    const int g = tid >> 7;                 // warpgroup
    const int r = tid & 127;                // point row of the tile = TMEM lane
    const int j0 = UPT * g;
    const int nhalf = j0 >> 4;              // which of the two MMA commits covers this thread's units
    const int dbg = a.dbg;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

    const int item = a.order ? a.order[blockIdx.x] : (int)blockIdx.x;
    const int sp = a.items[item * 4 + 0], first = a.items[item * 4 + 1], count = a.items[item * 4 + 2];
    const int im = a.sub_ids[sp];
    const int xd = a.xd;
    const float* ss = a.sub_static + (int64_t)im * (2 * xd + 3);
    float mu[3], isd[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (d < xd) {
            const float lo = ss[d], hi = ss[xd + d];
            mu[d] = (hi + lo) * 0.5f;
            isd[d] = 1.0f / ((hi - lo) * 0.5f);
        } else { mu[d] = 0.0f; isd[d] = 0.0f; }
    }
    const float flag = ss[2 * xd], un_mu = ss[2 * xd + 1], un_sd = ss[2 * xd + 2];
    const float* prow = a.params + (int64_t)im * a.P;
    fast_load_params<CF, NT>(sm, prow, xd, isd, a.axis, false);
    stage_b(bhi, blo, prow + H * xd + H, H, 1, tid, NT);       // B[n = j][k] = W1[j][k]
    if (tid == 0) {
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();                                       // B is read by the tensor core (async proxy)
    __syncthreads();
    // the allocation comes after the staging so that a CTA waiting for the previous one's columns has its prologue done
    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint64_t bdesc_hi = make_smem_desc(smem_u32(bhi), B_LBO, B_SBO);
    const uint64_t bdesc_lo = make_smem_desc(smem_u32(blo), B_LBO, B_SBO);

    // software prefetch of the next tile's point (pair -> point index -> coordinates)
    int pf_pt = 0;
    float pf_x[3] = {0.0f, 0.0f, 0.0f};
    auto load_idx = [&](int t0n) {
        if (t0n < count) pf_pt = a.spair_point[first + t0n + (r < min(TP, count - t0n) ? r : 0)];
    };
    auto load_val = [&](int t0n) {
        if (t0n < count) {
#pragma unroll
            for (int d = 0; d < 3; ++d) pf_x[d] = d < xd ? a.x[(int64_t)pf_pt * xd + d] : 0.0f;
        }
    };
    load_idx(0);
    load_val(0);

    uint32_t parity = 0;
    for (int t0 = 0; t0 < count; t0 += TP, parity ^= 1) {
        const int cnt = min(TP, count - t0);
        float z[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) z[d] = d < xd ? (pf_x[d] - mu[d]) * isd[d] : 0.0f;
        load_idx(t0 + TP);

        // ---- layer 0 for this thread's 16 units: tanh jets -> A (hi, lo) in tensor memory -------------------
        if (!(dbg & 4))
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int jb = j0 + 8 * ch;
            float hv[8][C];
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
                const float4 w0 = *reinterpret_cast<const float4*>(sm + CF::SM_W0 + jb + 4 * q4);
                const float4 w1 = *reinterpret_cast<const float4*>(sm + CF::SM_W0 + H + jb + 4 * q4);
                const float4 w2 = *reinterpret_cast<const float4*>(sm + CF::SM_W0 + 2 * H + jb + 4 * q4);
                const float4 b0 = *reinterpret_cast<const float4*>(sm + CF::SM_B0 + jb + 4 * q4);
                float4 wd[NS > 0 ? NS : 1];
#pragma unroll
                for (int s = 0; s < NS; ++s) wd[s] = *reinterpret_cast<const float4*>(sm + CF::SM_W0D + s * H + jb + 4 * q4);
                const float w0a[4] = {w0.x, w0.y, w0.z, w0.w}, w1a[4] = {w1.x, w1.y, w1.z, w1.w};
                const float w2a[4] = {w2.x, w2.y, w2.z, w2.w}, b0a[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float (&av)[C] = hv[4 * q4 + e];
                    av[0] = fmaf(w2a[e], z[2], fmaf(w1a[e], z[1], fmaf(w0a[e], z[0], b0a[e])));
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const float wv = e == 0 ? wd[s].x : (e == 1 ? wd[s].y : (e == 2 ? wd[s].z : wd[s].w));
                        if (s < NA2) { av[1 + 2 * s] = wv; av[2 + 2 * s] = 0.0f; }
                        else av[1 + 2 * NA2 + (s - NA2)] = wv;
                    }
                    fast_tanh_jets<CF>(av);
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) tf32_split(hv[e][c], hi[e], lo[e]);
                tmem_st8(tbase + lane_base + COL_AHI + c * 32 + jb, hi);
                tmem_st8(tbase + lane_base + COL_ALO + c * 32 + jb, lo);
            }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();                                         // A complete; previous tile's D fully read

        // ---- hidden GEMM on the tensor core: D[c] = A[c] * W1^T, unit half 0 first ---------------------------
        if (warp == 0 && !(dbg & 2)) {
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int nh = 0; nh < 2; ++nh) {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        issue_gemm_half(tbase + COL_D + c * 32 + nh * 16, tbase + COL_AHI + c * 32, tbase + COL_ALO + c * 32,
                                        bdesc_hi, bdesc_lo, nh);
                    mma_commit(&mma_bar[nh]);
                }
            }
            __syncwarp();
        }
        load_val(t0 + TP);

        // ---- epilogue of this thread's unit half: bias, tanh jets, cache, partial output dot ------------------
        if (!(dbg & 2)) mbar_wait_or_trap(&mma_bar[nhalf], parity);
        tc_fence_after();
        float up[C];
#pragma unroll
        for (int c = 0; c < C; ++c) up[c] = 0.0f;
        float* cb = nullptr;
        int cntb = 0;
        if (a.cache != nullptr && r < cnt) {
            constexpr int TPB = CF::TPB;                          // the reverse kernel's tile (its cache layout)
            const int off = t0 + r;
            const int t0b = (off / TPB) * TPB;
            cntb = min(TPB, count - t0b);
            cb = a.cache + (int64_t)(first + t0b) * (H * C) + (off - t0b);
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int jb = j0 + 8 * ch;
            uint32_t v[C][8];
#pragma unroll
            for (int c = 0; c < C; ++c) tmem_ld8(tbase + lane_base + COL_D + c * 32 + jb, v[c]);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float acc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = __uint_as_float(v[c][e]);
                acc[0] += sm[CF::SM_B1 + jb + e];
                if (!(dbg & 8)) fast_tanh_jets<CF>(acc);
                if (cb != nullptr) {
#pragma unroll
                    for (int c = 0; c < C; ++c) cb[((jb + e) * C + c) * cntb] = acc[c];
                }
                const float wl = sm[CF::SM_WL + jb + e];
#pragma unroll
                for (int c = 0; c < C; ++c) up[c] = fmaf(wl, acc[c], up[c]);
            }
        }
        if (g > 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) exch[((g - 1) * C + c) * TP + r] = up[c];
        }
        if (!(dbg & 2)) mbar_wait_or_trap(&mma_bar[1], parity);  // every MMA of the tile is done: A may be rewritten
        tc_fence_before();
        __syncthreads();

        // ---- output layer, window jets and Leibniz product per point (warpgroup 0) ---------------------------
        if (g == 0) {
            float u[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float sum = up[c];
#pragma unroll
                for (int gg = 1; gg < NWG; ++gg) sum += exch[((gg - 1) * C + c) * TP + r];
                u[c] = un_sd * (sum + (c == 0 ? sm[CF::SM_BL] : 0.0f));
            }
            u[0] += un_mu;
            float w, w1[NS > 0 ? NS : 1], w2[NA2 > 0 ? NA2 : 1];
            fast_window<CF>(z, isd, xd, flag, a.axis, w, w1, w2);
            float* o = outN + r * C;
            o[a.ext[0]] = u[0] * w;
#pragma unroll
            for (int s = 0; s < NA2; ++s) {
                const float u1 = u[1 + 2 * s], u2 = u[2 + 2 * s];
                o[a.ext[1 + 2 * s]] = u1 * w + u[0] * w1[s];
                o[a.ext[2 + 2 * s]] = u2 * w + 2.0f * u1 * w1[s] + u[0] * w2[s];
            }
#pragma unroll
            for (int s = 0; s < NA1; ++s) {
                const int c = 1 + 2 * NA2 + s;
                o[a.ext[c]] = u[c] * w + u[0] * w1[NA2 + s];
            }
        }
        __syncthreads();
        float* dst = a.pair_out + (int64_t)(first + t0) * C;
        for (int i = tid; i < cnt * C; i += NT) dst[i] = outN[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TMEM_COLS);
}


// =====================================================================================================
// prologue of the second-generation kernels
// =====================================================================================================
// A work item costs a fixed ~10 000 cycles before its first tile when the launch order, the work list, the subdomain index
// and the parameters are four dependent global loads and the staging passes are separated by CTA barriers
// (profiles/r2h_item_cost.md).  Here: one 16-byte launch record (fbp_takes_view.d_launch_*), tensor-memory allocation and
// barrier initialisation issued before the loads, every parameter read straight from global memory (kappa_s recomputed per
// element instead of read back from shared memory), bank-conflict-free placement of the B operands, ONE barrier.
struct ItemRec { int first, count, im, item; };
__device__ __forceinline__ ItemRec tc_item(const FastArgs& a) {
    ItemRec r;
    if (a.launch != nullptr) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(a.launch) + blockIdx.x);
        r.first = v.x; r.count = v.y; r.im = v.z; r.item = v.w;
    } else {
        r.item = a.order ? a.order[blockIdx.x] : (int)blockIdx.x;
        r.first = a.items[r.item * 4 + 1];
        r.count = a.items[r.item * 4 + 2];
        r.im = a.sub_ids[a.items[r.item * 4 + 0]];
    }
    return r;
}

// the small parameter vectors (first layer, biases, output weights) into the CF::SM_* block; no barrier inside
template <class CF>
__device__ __forceinline__ void tc_stage_small(float* sm, const float* __restrict__ prow, int xd, const float (&isd)[3],
                                               const int (&axis)[FBP_MAX_XD], int tid, int nt) {
    for (int i = tid; i < 3 * H; i += nt) {
        const int d = i >> 5, j = i & 31;
        sm[CF::SM_W0 + i] = d < xd ? prow[j * xd + d] : 0.0f;
    }
    for (int i = tid; i < H; i += nt) sm[CF::SM_B0 + i] = prow[H * xd + i];
    for (int i = tid; i < CF::NS * H; i += nt) {
        const int s_ = i >> 5, j = i & 31, ax = axis[s_];
        sm[CF::SM_W0D + i] = prow[j * xd + ax] * sel3(ax, isd[0], isd[1], isd[2]);
    }
    const int off = H * xd + H + H * H;
    for (int i = tid; i < H; i += nt) {
        sm[CF::SM_B1 + i] = prow[off + i];
        sm[CF::SM_WL + i] = prow[off + H + i];
    }
    if (tid == 0) sm[CF::SM_BL] = prow[off + 2 * H];
}

// element i of the staging pass (lane = i % 32) -> W1[j][k]: lane bits = (j & 3, k & 3, bit 2 of j), the rest of i = (j >> 3, k >> 2)
__host__ __device__ constexpr int stage_b_row(int i) { return (((i >> 5) & 3) << 3) | ((((i & 31) >> 4) & 1) << 2) | (i & 3); }
__host__ __device__ constexpr int stage_b_col(int i) { return (((i >> 5) >> 2) << 2) | (((i & 31) >> 2) & 3); }

// B operands, hi / lo, canonical K-major layout:  b1[v][n = j][k] = W1[j][k] sc1_v[k]  (sc1 = 1, -2 kappa_s^2),  and, for the
// reverse kernel,  b2[v][n = k][j] = W1[j][k] sc2_v[k]  (sc2 = 1, kappa_s, -2 kappa_s^2).  Lane -> (j, k) is chosen so that the
// 32 stores of a warp hit 32 (b1) / 16 (b2) different banks.
template <class CF, bool WITH_B2>
__device__ __forceinline__ void tc_stage_b(float* b1, float* b2, const float* __restrict__ prow, int xd, const float (&isd)[3],
                                           const int (&axis)[FBP_MAX_XD], int tid, int nt) {
    constexpr int NS = CF::NS, NA2 = CF::NA2, NB1 = 1 + NA2, NB2 = 1 + NS + NA2, HH = H * H;
    const float* w1 = prow + H * xd + H;
    for (int i = tid; i < HH; i += nt) {
        const int j = stage_b_row(i), k = stage_b_col(i);
        const float w = w1[j * H + k];
        float kap[NS > 0 ? NS : 1];
#pragma unroll
        for (int s_ = 0; s_ < NS; ++s_) kap[s_] = prow[k * xd + axis[s_]] * sel3(axis[s_], isd[0], isd[1], isd[2]);
#pragma unroll
        for (int v = 0; v < NB1; ++v) {
            const float sc = v == 0 ? 1.0f : -2.0f * kap[v > 0 ? v - 1 : 0] * kap[v > 0 ? v - 1 : 0];
            uint32_t hi, lo;
            tf32_split(w * sc, hi, lo);
            b1[(2 * v) * HH + bcore_index(j, k)] = __uint_as_float(hi);
            b1[(2 * v + 1) * HH + bcore_index(j, k)] = __uint_as_float(lo);
        }
        if (WITH_B2) {
#pragma unroll
            for (int v = 0; v < NB2; ++v) {
                float sc = 1.0f;
                if (v >= 1 && v <= NS) sc = kap[v >= 1 && v <= NS ? v - 1 : 0];
                else if (v > NS) sc = -2.0f * kap[v > NS ? v - 1 - NS : 0] * kap[v > NS ? v - 1 - NS : 0];
                uint32_t hi, lo;
                tf32_split(w * sc, hi, lo);
                b2[(2 * v) * HH + bcore_index(k, j)] = __uint_as_float(hi);
                b2[(2 * v + 1) * HH + bcore_index(k, j)] = __uint_as_float(lo);
            }
        }
    }
}

// =====================================================================================================
// forward, software-pipelined kernel (the default: FBP_TC_FWD=2)
// =====================================================================================================
// The first kernel pays the tensor-core time and ~0.85 ms of latency serially because only one tile fits tensor memory
// (profiles/r1f_tc_bringup.md).  Here the CTA still owns one tile's worth of TMEM, but the phases of consecutive tiles are
// interleaved:
//     wait MMA(t)  ->  layer 0 of tile t+1 -> A   ->  read D(t) into registers  ->  barrier  ->  issue MMA(t+1)
//                  ->  epilogue math of tile t (tanh jets, cache, output dot, window) while MMA(t+1) runs
// A is free once MMA(t) is complete and D is free once every thread holds its accumulators, so MMA(t+1) overlaps the
// whole epilogue of tile t.  Round 2: the A operands are the factored ones (t, g kappa_s, t g: Factored<CF>, 4 instead of 5
// hi/lo pairs to split and store at C = 5), every product is one N = 32 (12 instructions per component instead of 24), and
// a 17th warp issues them, so that no epilogue warp is held back by the issue loop (the MMA warpgroup hands its registers
// to the 16 epilogue warps with setmaxnreg).
constexpr int F2_NPT = 512, F2_NT = F2_NPT + 128;
template <class CF>
struct Fwd2Smem {
    static constexpr int OFF_B1 = (CF::SM_PARAMS + 31) & ~31;                       // [NB1][hi, lo][H*H]
    static constexpr int OFF_EXCH = OFF_B1 + Factored<CF>::NB1 * 2 * H * H;          // [2][3][C][TP] partial output dots of warpgroups 1..3, by tile parity
    static constexpr int OFF_OUT = OFF_EXCH + 2 * 3 * CF::C * TP;                    // [TP][C] tile output in external component order
    static constexpr int FLOATS = OFF_OUT + CF::C * TP;
};

template <class CF>
__global__ void __launch_bounds__(F2_NT, 1) tc_forward_kernel2(FastArgs a) {
    static_assert(CF::H == 32 && CF::NHID == 2, "tensor family: H = 32, two hidden layers");
    using FA = Factored<CF>;
    constexpr int NWG = 4;
    constexpr int C = CF::C, NS = CF::NS, NA2 = CF::NA2, NA1 = CF::NA1, CA1 = FA::CA1;
    constexpr uint32_t COL_ALO = CA1 * 32, COL_D = 2 * CA1 * 32;
    static_assert(COL_D + C * 32 <= (int)TMEM_COLS, "A hi, A lo and D must fit the 512 TMEM columns");
    using L = Fwd2Smem<CF>;
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar_a, bar_m;
    __shared__ uint32_t tmem_slot;
    __shared__ int hdr_i[2];                  // first pair, pair count
    __shared__ float hdr_f[9];                // mu[3], 1/sd[3], window flag, output shift and scale
    float* exch = sm + L::OFF_EXCH;
    float* outN = sm + L::OFF_OUT;

    const int tid = threadIdx.x, warp = warp_uniform();
    blk_stamp(1);
    const int g = (tid >> 7) & 3, r = tid & 127;
    const int jb = 8 * g;                   // this thread's 8 hidden units
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

    const int xd = a.xd;
    {   // prologue: nothing computed here stays in registers (see tc_backward_kernel2)
        const ItemRec it = tc_item(a);
        if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
        const float* ss = a.sub_static + (int64_t)it.im * (2 * xd + 3);
        float isd_[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) isd_[d] = d < xd ? 1.0f / ((ss[xd + d] - ss[d]) * 0.5f) : 0.0f;
        if (tid == 32) {
            hdr_i[0] = it.first; hdr_i[1] = it.count;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                hdr_f[d] = d < xd ? (ss[xd + d] + ss[d]) * 0.5f : 0.0f;
                hdr_f[3 + d] = isd_[d];
            }
            hdr_f[6] = ss[2 * xd];
            hdr_f[7] = ss[2 * xd + 1];
            hdr_f[8] = ss[2 * xd + 2];
            mbar_init(&bar_a, F2_NPT);
            mbar_init(&bar_m, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        const float* prow = a.params + (int64_t)it.im * a.P;
        tc_stage_small<CF>(sm, prow, xd, isd_, a.axis, tid, F2_NT);
        tc_stage_b<CF, false>(sm + L::OFF_B1, nullptr, prow, xd, isd_, a.axis, tid, F2_NT);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const int first = hdr_i[0], count = hdr_i[1];
    float mu[3], isd[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { mu[d] = hdr_f[d]; isd[d] = hdr_f[3 + d]; }
    const float flag = hdr_f[6], un_mu = hdr_f[7], un_sd = hdr_f[8];
    const int ntiles = (count + TP - 1) / TP;
    if (warp < 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");

    if (warp == 16) {
        // ---- MMA warp: tile t's products as soon as its A operands are complete and the previous D has been read ----
        const uint64_t b1base = make_smem_desc(smem_u32(sm + L::OFF_B1), B_LBO, B_SBO);
        for (int t = 0; t < ntiles; ++t) {
            mbar_wait_or_trap(&bar_a, (uint32_t)(t & 1));
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int ia = FA::ia(c), vb = FA::vb(c);
                    issue_gemm_acc<32>(tbase + COL_D + c * 32, tbase + ia * 32, tbase + COL_ALO + ia * 32,
                                       desc_advance(b1base, (uint32_t)((2 * vb) * H * H * 4)),
                                       desc_advance(b1base, (uint32_t)((2 * vb + 1) * H * H * 4)), true);
                }
                mma_commit(&bar_m);
            }
            __syncwarp();
        }
    } else if (warp < 16) {
        blk_stamp(2);
        int pf_pt = 0;
        float pf_x[3] = {0.0f, 0.0f, 0.0f};
        auto load_idx = [&](int tt) {
            const int t0n = tt * TP;
            if (t0n < count) pf_pt = a.spair_point[first + t0n + (r < min(TP, count - t0n) ? r : 0)];
        };
        auto load_val = [&](int tt) {
            if (tt * TP < count) {
#pragma unroll
                for (int d = 0; d < 3; ++d) pf_x[d] = d < xd ? a.x[(int64_t)pf_pt * xd + d] : 0.0f;
            }
        };
        auto normalise = [&](float (&z)[3]) {
#pragma unroll
            for (int d = 0; d < 3; ++d) z[d] = d < xd ? (pf_x[d] - mu[d]) * isd[d] : 0.0f;
        };
        // layer 0 of this thread's 8 units at z -> factored A operands (t, g kappa_s, t g), hi / lo, in tensor memory
        auto layer0_to_tmem = [&](const float (&z)[3]) {
            float av[CA1][8];
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
                const float4 w0 = *reinterpret_cast<const float4*>(sm + CF::SM_W0 + jb + 4 * q4);
                const float4 w1 = *reinterpret_cast<const float4*>(sm + CF::SM_W0 + H + jb + 4 * q4);
                const float4 w2 = *reinterpret_cast<const float4*>(sm + CF::SM_W0 + 2 * H + jb + 4 * q4);
                const float4 b0 = *reinterpret_cast<const float4*>(sm + CF::SM_B0 + jb + 4 * q4);
                float4 wd[NS > 0 ? NS : 1];
#pragma unroll
                for (int s = 0; s < NS; ++s) wd[s] = *reinterpret_cast<const float4*>(sm + CF::SM_W0D + s * H + jb + 4 * q4);
                const float w0a[4] = {w0.x, w0.y, w0.z, w0.w}, w1a[4] = {w1.x, w1.y, w1.z, w1.w};
                const float w2a[4] = {w2.x, w2.y, w2.z, w2.w}, b0a[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const int e = 4 * q4 + e4;
                    const float a0 = fmaf(w2a[e4], z[2], fmaf(w1a[e4], z[1], fmaf(w0a[e4], z[0], b0a[e4])));
                    const float tv = fbp_tanh(a0);
                    const float gv = 1.0f - tv * tv;
                    av[0][e] = tv;
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        av[1 + s][e] = gv * (e4 == 0 ? wd[s].x : (e4 == 1 ? wd[s].y : (e4 == 2 ? wd[s].z : wd[s].w)));
                    if (NA2 > 0) av[CA1 - 1][e] = tv * gv;
                }
            }
#pragma unroll
            for (int i = 0; i < CA1; ++i) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    hi[e] = __float_as_uint(av[i][e]);
                    lo[e] = __float_as_uint(tf32_lo(av[i][e]));
                }
                tmem_st8(tbase + lane_base + i * 32 + jb, hi);
                tmem_st8(tbase + lane_base + COL_ALO + i * 32 + jb, lo);
            }
        };

        // ---- prologue: tile 0 into A; coordinates of tile 1 and the index of tile 2 on their way -----------------------
        float z_cur[3];
        load_idx(0);
        load_val(0);
        normalise(z_cur);
        load_idx(1);
        load_val(1);
        load_idx(2);
        if (ntiles > 0) {
            layer0_to_tmem(z_cur);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&bar_a);
        }

        for (int t = 0; t < ntiles; ++t) {
            const int t0 = t * TP;
            const int cnt = min(TP, count - t0);
            const bool more = t + 1 < ntiles;
            float z_next[3];
            normalise(z_next);                                        // pf_x holds the coordinates of tile t+1
            load_val(t + 2);
            load_idx(t + 3);

            mbar_wait_or_trap(&bar_m, (uint32_t)(t & 1));            // every MMA of tile t is complete: A is free, D is final
            tc_fence_after();
            if (more) layer0_to_tmem(z_next);
            uint32_t v[C][8];
#pragma unroll
            for (int c = 0; c < C; ++c) tmem_ld8(tbase + lane_base + COL_D + c * 32 + jb, v[c]);
            tmem_wait_ld();
            if (more) {
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&bar_a);                                  // A(t+1) complete, D(t) held in registers
            }

            // ---- epilogue math of tile t (the tensor core works on tile t+1 meanwhile) -----------------------------
            float up[C];
#pragma unroll
            for (int c = 0; c < C; ++c) up[c] = 0.0f;
            float* cb = nullptr;
            int cntb = 0;
            if (a.cache != nullptr && r < cnt) {
                constexpr int TPB = CF::TPB;
                const int off = t0 + r;
                const int t0b = (off / TPB) * TPB;
                cntb = min(TPB, count - t0b);
                cb = a.cache + (int64_t)(first + t0b) * (H * C) + (off - t0b);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float acc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = __uint_as_float(v[c][e]);
                acc[0] += sm[CF::SM_B1 + jb + e];
                fast_tanh_jets<CF>(acc);
                if (cb != nullptr) {
#pragma unroll
                    for (int c = 0; c < C; ++c) cb[((jb + e) * C + c) * cntb] = acc[c];
                }
                const float wl = sm[CF::SM_WL + jb + e];
#pragma unroll
                for (int c = 0; c < C; ++c) up[c] = fmaf(wl, acc[c], up[c]);
            }
            float* ex = exch + (t & 1) * 3 * C * TP;                 // double buffered: one barrier per tile is enough
            if (g > 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) ex[((g - 1) * C + c) * TP + r] = up[c];
            }
            // partial dots exchanged: only the four warps that share this quarter's 32 points meet (one named barrier per quarter)
            asm volatile("bar.sync %0, 128;" ::"r"(2 + (warp & 3)) : "memory");
            if (g == 0) {
                float u[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    float sum = up[c];
#pragma unroll
                    for (int gg = 1; gg < NWG; ++gg) sum += ex[((gg - 1) * C + c) * TP + r];
                    u[c] = un_sd * (sum + (c == 0 ? sm[CF::SM_BL] : 0.0f));
                }
                u[0] += un_mu;
                float w, w1[NS > 0 ? NS : 1], w2[NA2 > 0 ? NA2 : 1];
                fast_window<CF>(z_cur, isd, xd, flag, a.axis, w, w1, w2);
                float* o = outN + r * C;
                o[a.ext[0]] = u[0] * w;
#pragma unroll
                for (int s = 0; s < NA2; ++s) {
                    const float u1 = u[1 + 2 * s], u2 = u[2 + 2 * s];
                    o[a.ext[1 + 2 * s]] = u1 * w + u[0] * w1[s];
                    o[a.ext[2 + 2 * s]] = u2 * w + 2.0f * u1 * w1[s] + u[0] * w2[s];
                }
#pragma unroll
                for (int s = 0; s < NA1; ++s) {
                    const int c = 1 + 2 * NA2 + s;
                    o[a.ext[c]] = u[c] * w + u[0] * w1[NA2 + s];
                }
                // each warp of unit group 0 copies the 32 rows it staged (contiguous in pair_out): warp-level ordering is enough
                __syncwarp();
                const int lane = tid & 31, row0 = r - lane;
                const int nrow = min(32, cnt - row0);
                float* dst = a.pair_out + (int64_t)(first + t0 + row0) * C;
                const float* src = outN + row0 * C;
                for (int i = lane; i < nrow * C; i += 32) dst[i] = src[i];
                __syncwarp();
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) z_cur[d] = z_next[d];
        }
        BLK_STAMP3;
    }
    tc_fence_before();
    __syncthreads();
    blk_stamp(4);
    if (warp == 0) tmem_dealloc(tbase, TMEM_COLS);
    blk_stamp(5);
    blk_stamp_entry(ntiles);
}

}  // namespace fbptc

// launchers of the tensor family (fbp_tc.cu)
int fbp_tc_forward_launch(const FastSpec& f, const FastArgs& a, int grid, cudaStream_t st);
