// tcgen05 ("tensor") kernel family: instantiations, launch glue and the MMA self-test.
#include "fbp_tc_bwd2.cuh"

using namespace fbptc;

// -----------------------------------------------------------------------------------------------------------------
// Self-test of the one MMA form the family uses (A from tensor memory written row-wise by its owner threads, B from a
// no-swizzle K-major shared-memory descriptor, N = 16 halves, 3xTF32): out[128][32] = A[128][32] * W[32][32]^T.
// `variant` lets a hardware session probe descriptor conventions without rebuilding:
//   bit 0: swap LBO and SBO in the descriptor          bit 1: descriptor version field 0 instead of 1
//   bit 2: single TF32 pass (hi*hi only; expected error ~1e-3, distinguishes "wrong layout" from "wrong split")
//   bit 3: the lo parts (A and B) are passed as raw FP32 bit patterns (does the tensor core ignore the low 13 bits?)
//   bit 4: the hi parts are truncated (one LOP3) instead of rounded to nearest
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                            float* __restrict__ out, int variant) {
    __shared__ __align__(128) float bhi[H * H];
    __shared__ __align__(128) float blo[H * H];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = warp_uniform(), g = tid >> 7, r = tid & 127, j0 = 16 * g;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    constexpr uint32_t COL_AHI = 0, COL_ALO = 32, COL_D = 64;

    stage_b(bhi, blo, W, H, 1, tid);
    if (variant & 24) {
        for (int i = tid; i < H * H; i += NT) {
            const int n = i >> 5, k = i & 31;
            const float x = W[i];
            const uint32_t hi = (variant & 16) ? (__float_as_uint(x) & 0xffffe000u) : tf32_rn(x);
            const float lo = x - __uint_as_float(hi);
            bhi[bcore_index(n, k)] = __uint_as_float(hi);
            blo[bcore_index(n, k)] = (variant & 8) ? lo : __uint_as_float(tf32_rn(lo));
        }
    }
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;

#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float x = A[r * H + j0 + 8 * ch + e];
            tf32_split(x, hi[e], lo[e]);
            if (variant & 24) {
                hi[e] = (variant & 16) ? (__float_as_uint(x) & 0xffffe000u) : tf32_rn(x);
                const float l = x - __uint_as_float(hi[e]);
                lo[e] = (variant & 8) ? __float_as_uint(l) : tf32_rn(l);
            }
        }
        tmem_st8(tbase + lane_base + COL_AHI + j0 + 8 * ch, hi);
        tmem_st8(tbase + lane_base + COL_ALO + j0 + 8 * ch, lo);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
        tc_fence_after();
        const uint32_t lbo = (variant & 1) ? B_SBO : B_LBO, sbo = (variant & 1) ? B_LBO : B_SBO;
        uint64_t dh = make_smem_desc(smem_u32(bhi), lbo, sbo), dl = make_smem_desc(smem_u32(blo), lbo, sbo);
        if (variant & 2) { dh &= ~((uint64_t)3 << 46); dl &= ~((uint64_t)3 << 46); }
        for (int nh = 0; nh < 2; ++nh) {
            if (variant & 4) {
                constexpr uint32_t idesc = make_idesc_tf32(128, 16);
                for (int ks = 0; ks < 4; ++ks)
                    mma_tf32_ts(tbase + COL_D + nh * 16, tbase + COL_AHI + ks * 8,
                                dh + (uint64_t)((nh * 2 * B_SBO) >> 4) + (uint64_t)((ks * 2 * B_LBO) >> 4), idesc, ks != 0);
            } else {
                issue_gemm_half(tbase + COL_D + nh * 16, tbase + COL_AHI, tbase + COL_ALO, dh, dl, nh);
            }
            mma_commit(&bar[nh]);
        }
    }
    mbar_wait_or_trap(&bar[g], 0);
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[8];
        tmem_ld8(tbase + lane_base + COL_D + j0 + 8 * ch, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) out[r * H + j0 + 8 * ch + e] = __uint_as_float(v[e]);
    }
    mbar_wait_or_trap(&bar[1], 0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TMEM_COLS);
}

extern "C" int fbp_tc_selftest(const float* d_a, const float* d_w, float* d_out, int32_t variant, void* stream) {
    FBP_REQUIRE(d_a && d_w && d_out, "fbp_tc_selftest: null buffer");
    tc_selftest_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(d_a, d_w, d_out, variant);
    FBP_LAUNCH_CHECK();
    return 0;
}

// -----------------------------------------------------------------------------------------------------------------
// Self-test of the operand form of the tensor-core weight gradient (fbp_tc_bwd2.cuh): BOTH operands from shared memory
// (".ss"), contraction over the 128 "points" p of a tile:  out[m][n] = sum_p A[p][m] * B[p][n],  m < 128 (a stack of four
// 32-row images), n < 32.  The operands are staged exactly as the point warps do it: thread = point, one image row per
// store, K-major no-swizzle core matrices with the K step padded to GK_LBO = 144 bytes (conflict-free 32-lane stores),
// one slot per quarter tile (32 points = 4 MMAs of K = 8), accumulation over the four quarters in tensor memory.
// Inputs are rounded to TF32 so that one pass is exact up to FP32 accumulation.
//   variant bit 0: swap LBO and SBO     bit 1: unpadded K step (LBO 128, SBO 1024)
//   bit 2: M = 64 (only the first 64 rows of A take part; the caller maps accumulator rows to tensor-memory lanes)
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tc_selftest_g_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ out, int variant) {
    extern __shared__ __align__(128) float sg[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = warp_uniform();
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t lbo = (variant & 2) ? 128u : GK_LBO, sbo = 8u * lbo;
    const int img = 4 * (int)sbo / 4;                      // floats per 32-row image of one quarter
    const int slot = 5 * img;                              // 4 A images + 1 B image per quarter
    const int q = tid >> 5, pq = tid & 31;                 // quarter and point inside it
    for (int m = 0; m < 128; ++m)
        sg[q * slot + (m >> 5) * img + gk_index(m & 31, pq, lbo)] = __uint_as_float(tf32_rn(A[tid * 128 + m]));
    for (int n = 0; n < 32; ++n) sg[q * slot + 4 * img + gk_index(n, pq, lbo)] = __uint_as_float(tf32_rn(B[tid * 32 + n]));
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = (variant & 4) ? make_idesc_tf32(64, 32) : make_idesc_tf32(128, 32);     // bit 2: M = 64
        const uint32_t dl = (variant & 1) ? sbo : lbo, ds = (variant & 1) ? lbo : sbo;
        for (int qq = 0; qq < 4; ++qq) {
            const uint64_t ad = make_smem_desc(smem_u32(sg + qq * slot), dl, ds);
            const uint64_t bd = make_smem_desc(smem_u32(sg + qq * slot + 4 * img), dl, ds);
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t adv = (uint64_t)((ks * 2 * lbo) >> 4);
                mma_tf32_ss(tbase, ad + adv, bd + adv, idesc, (qq | ks) != 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait_or_trap(&bar, 0);
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[8];
        tmem_ld8(tbase + lane_base + 8 * ch, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) out[tid * 32 + 8 * ch + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 32);
}

extern "C" int fbp_tc_selftest_g(const float* d_a, const float* d_b, float* d_out, int32_t variant, void* stream) {
    FBP_REQUIRE(d_a && d_b && d_out, "fbp_tc_selftest_g: null buffer");
    constexpr int bytes = 4 * 5 * 8 * (int)GK_LBO * 4;     // 4 quarters x 5 images x 4 row groups x 8 K cores
    FBP_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest_g_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    tc_selftest_g_kernel<<<1, 128, bytes, (cudaStream_t)stream>>>(d_a, d_b, d_out, variant);
    FBP_LAUNCH_CHECK();
    return 0;
}

// -----------------------------------------------------------------------------------------------------------------
// plan matching and launchers
// -----------------------------------------------------------------------------------------------------------------
int fbp_tc_supported(const FastSpec& f, int C) {
    if (f.H != 32 || f.nhid != 2) return 0;
    if (3 * C * 32 > (int)TMEM_COLS) return 0;
    const int key = f.na2 * 4 + f.na1;
    return key == 0 || key == 1 || key == 4 || key == 5 || key == 8;
}

#include <cstdlib>
// FBP_TC_NWG = 2 | 4: warpgroups per CTA (A/B knob);  FBP_TC_DEBUG: timing-experiment bits (see tc_forward_kernel)
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <class CF, int NWG>
static int tc_forward_nwg(FastArgs a, int grid, cudaStream_t st) {
    constexpr size_t bytes = sizeof(float) * FwdSmem<CF, NWG>::FLOATS;
    // the attribute belongs to the (function, device) pair and the call is cheap: set it on every launch, so that a
    // process driving several GPUs configures each of them
    FBP_CHECK_CUDA(cudaFuncSetAttribute(tc_forward_kernel<CF, NWG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    a.dbg = env_int("FBP_TC_DEBUG", 0);
    tc_forward_kernel<CF, NWG><<<grid, 128 * NWG, bytes, st>>>(a);
    FBP_LAUNCH_CHECK();
    return 0;
}

template <class CF>
static int tc_forward_v2(FastArgs a, int grid, cudaStream_t st) {
    constexpr size_t bytes = sizeof(float) * Fwd2Smem<CF>::FLOATS;
    // the attribute belongs to the (function, device) pair and the call is cheap: set it on every launch, so that a
    // process driving several GPUs configures each of them
    FBP_CHECK_CUDA(cudaFuncSetAttribute(tc_forward_kernel2<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    tc_forward_kernel2<CF><<<grid, F2_NT, bytes, st>>>(a);
    FBP_LAUNCH_CHECK();
    return 0;
}

template <class CF>
static int tc_forward_one(const FastArgs& a, int grid, cudaStream_t st) {
    // the software-pipelined kernel is the default since it was timed (1.383 vs 1.716 ms without cache stores, 1.721 vs
    // 1.996 ms with them, cfg 5, profiles/r2a_tc_bringup.md); FBP_TC_FWD=1 selects the first kernel
    if (env_int("FBP_TC_FWD", 2) == 2) return tc_forward_v2<CF>(a, grid, st);
    // 4 warpgroups measured faster than 2 (1.995 vs 2.238 ms on cfg 5, profiles/r1f_tc_bringup.md)
    if (env_int("FBP_TC_NWG", 4) == 2) return tc_forward_nwg<CF, 2>(a, grid, st);
    return tc_forward_nwg<CF, 4>(a, grid, st);
}

int fbp_tc_forward_launch(const FastSpec& f, const FastArgs& a, int grid, cudaStream_t st) {
    switch (f.na2 * 4 + f.na1) {
        case 0: return tc_forward_one<FastCfg<32, 2, 0, 0>>(a, grid, st);
        case 1: return tc_forward_one<FastCfg<32, 2, 0, 1>>(a, grid, st);
        case 4: return tc_forward_one<FastCfg<32, 2, 1, 0>>(a, grid, st);
        case 5: return tc_forward_one<FastCfg<32, 2, 1, 1>>(a, grid, st);
        case 8: return tc_forward_one<FastCfg<32, 2, 2, 0>>(a, grid, st);
        default: break;
    }
    fbp_set_error("fbp_tc: no tensor-family instance for jets=(%d,%d)", f.na2, f.na1);
    return 3;
}

template <class CF>
static int tc_backward_one(FastArgs a, int grid, cudaStream_t st) {
    constexpr size_t bytes = sizeof(float) * BwdSmem<CF>::FLOATS;
    static_assert(bytes <= 227 * 1024, "reverse kernel: shared memory budget");
    // the attribute belongs to the (function, device) pair and the call is cheap: set it on every launch, so that a
    // process driving several GPUs configures each of them
    FBP_CHECK_CUDA(cudaFuncSetAttribute(tc_backward_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    a.dbg = env_int("FBP_TC_DEBUG", 0);
    tc_backward_kernel<CF><<<grid, BWD_NT, bytes, st>>>(a);
    FBP_LAUNCH_CHECK();
    return 0;
}

// second generation: weight gradient on the tensor core (fbp_tc_bwd2.cuh); FBP_TC_NG = 2 | 4 unit groups (8 | 16 point warps)
template <class CF, int NG>
static int tc_backward_two_ng(FastArgs a, int grid, cudaStream_t st) {
    constexpr size_t bytes = sizeof(float) * Bwd2Cfg<CF, NG>::FLOATS;
    static_assert(bytes <= 226 * 1024, "reverse kernel 2: shared memory budget");
    FBP_CHECK_CUDA(cudaFuncSetAttribute(tc_backward_kernel2<CF, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    a.dbg = env_int("FBP_TC_DEBUG", 0);
    tc_backward_kernel2<CF, NG><<<grid, B2Dim<NG>::NT, bytes, st>>>(a);
    FBP_LAUNCH_CHECK();
    return 0;
}
template <class CF>
static int tc_backward_two(const FastArgs& a, int grid, cudaStream_t st) {
    if (env_int("FBP_TC_NG", 4) == 2) return tc_backward_two_ng<CF, 2>(a, grid, st);
    return tc_backward_two_ng<CF, 4>(a, grid, st);
}

int fbp_tc_backward_launch(const FastSpec& f, const FastArgs& a, int grid, cudaStream_t st) {
    if (env_int("FBP_TC_BWD", 2) == 2) {       // default since it was timed: 4.43 vs 5.21 ms (profiles/r2c_tc_bwd2.md)
        switch (f.na2 * 4 + f.na1) {
            case 0: return tc_backward_two<FastCfg<32, 2, 0, 0>>(a, grid, st);
            case 1: return tc_backward_two<FastCfg<32, 2, 0, 1>>(a, grid, st);
            case 4: return tc_backward_two<FastCfg<32, 2, 1, 0>>(a, grid, st);
            case 5: return tc_backward_two<FastCfg<32, 2, 1, 1>>(a, grid, st);
            case 8: return tc_backward_two<FastCfg<32, 2, 2, 0>>(a, grid, st);
            default: break;
        }
    }
    switch (f.na2 * 4 + f.na1) {
        case 0: return tc_backward_one<FastCfg<32, 2, 0, 0>>(a, grid, st);
        case 1: return tc_backward_one<FastCfg<32, 2, 0, 1>>(a, grid, st);
        case 4: return tc_backward_one<FastCfg<32, 2, 1, 0>>(a, grid, st);
        case 5: return tc_backward_one<FastCfg<32, 2, 1, 1>>(a, grid, st);
        case 8: return tc_backward_one<FastCfg<32, 2, 2, 0>>(a, grid, st);
        default: break;
    }
    fbp_set_error("fbp_tc: no tensor-family reverse instance for jets=(%d,%d)", f.na2, f.na1);
    return 3;
}

#ifdef FBP_BLOCK_TRACE
// trace builds only (tests/tools/item_cost_trace.py): the buffer the CTAs of the tensor kernels record their timeline into
extern "C" int fbp_debug_set_block_trace(void* d_buf) {
    return cudaMemcpyToSymbol(fbptc::fbp_blk_trace, &d_buf, sizeof(void*)) == cudaSuccess ? 0 : 1;
}
#endif
