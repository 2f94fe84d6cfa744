// Tiled ("fast") kernel family: the roofline kernels of the FBPINN subdomain evaluation on sm_100a.
//
// One CTA per work item = (subdomain, contiguous range of its pairs).  The subdomain's parameters are staged
// once in shared memory (hidden matrix both transposed and raw), its points are streamed in tiles of TP (64) with
// coalesced, software-prefetched loads, and every linear layer of the jet propagation is a register-tiled FP32 GEMM
//      A[j][(c,p)] = sum_k W[j][k] * H[k][(c,p)]        j: output unit, c: jet component, p: point of the tile
// where H lives in shared memory as [k][c][p] (p contiguous: lanes read consecutive point pairs, conflict free) and
// W^T as [k][j] (a warp reads ONE float4 -> broadcast).  Each thread owns TM=8 output units x PPT=2 points x C
// components; the two points form the lanes of Blackwell's packed FP32 FMA (fma.rn.f32x2 -> FFMA2) with the weight as
// scalar-broadcast operand, so one k-step costs 2 LDS.128 + C LDS.64 for 8*C FFMA2.  tanh jets are applied to the
// accumulators in registers and written back IN PLACE (one activation buffer per hidden layer).
//
// Forward: layer 0 (K = xd) -> [hidden GEMM -> tanh jets] -> output layer fused as a partial dot -> window jets and
// Leibniz product per point -> coalesced store of the tile's [TP][C] block.  When an activation cache is given, the
// last hidden layer's jets (the tile's [H][C][TP] shared-memory image) are saved with TMA bulk stores.
//
// Reverse (same CTA shape): h0 is recomputed (cheap), h1 is TMA-loaded from the activation cache into shared memory
// (cp.async.bulk + mbarrier, issued one tile ahead) — or recomputed when no cache is given — then
//   tanh-jet transpose of the last hidden layer in registers (+ the output-layer weight gradient as per-thread
//     register partials, reduced once per work item),
//   G: Wbar[j][k] += sum_{c,p} abar[j][c][p] h[k][c][p]   lanes own an 8x4 (j,k) tile of float2 accumulators (FFMA2
//      over point pairs), warps split the points, operands read as float4 along p (row stride padded by 4 floats ->
//      the 4/8 distinct rows of a request fall in distinct banks), accumulators persist across the tiles of the item,
//   D: hbar[k][(c,p)] = sum_j W[j][k] abar[j][(c,p)]      same FFMA2 GEMM as the forward with the raw matrix,
//   tanh-jet transpose of layer 0 (+ the first-layer gradients as per-thread partial sums written into the thread's
//     own activation slots and summed by one thread per output, fixed order).
// Results per work item go to gpart[item][P]; a second kernel sums the items of a subdomain in fixed order
// (deterministic, no float atomics anywhere).
//
// Bound: FP32 FMA pipe (CUDA cores).  Algorithmic FLOPs per pair: SURVEY §8d (F_fwd = 2 MAC_0 + 2 C sum MAC_l).
// Measured evidence and the history of these choices: DESIGN.md §4.1, profiles/r1*_*.
#pragma once
#include "fbp_common.cuh"

#ifndef FBP_GEMM_UNROLL
#define FBP_GEMM_UNROLL 2      // k-steps unrolled in the GEMM loops (2 measured best: 2.58 vs 2.66 ms forward at 4)
#endif

struct FastArgs {
    const float* x;
    const float* params;
    const float* sub_static;
    const int32_t* sub_ids;
    const int32_t* spair_point;
    const int32_t* spair_row;
    const int32_t* items;
    const int32_t* order;   // optional launch order: block b works on item order[b] (longest first)
    const int32_t* launch;  // optional [grid][4] launch records (first pair, pair count, subdomain index, item) in launch order
    float* pair_out;     // forward output  [s][C]
    const float* grow;   // reverse input   [q][C]
    float* gpart;        // reverse output  [n_items_active][P]
    float* cache;        // optional [s][H*C]: jets of the last hidden layer, written by the forward kernel per reverse
                         // tile as [H][C][cnt] at float offset (first pair of the tile)*H*C, read back by the reverse
    int xd, P;
    int dbg;                // timing-experiment switches of the tensor family (0 in production)
    int axis[FBP_MAX_XD];   // slot -> axis (NA2 slots first)
    int ext[FBP_MAX_COMP];  // internal component -> external component index
};

template <int H_, int NHID_, int NA2_, int NA1_>
struct FastCfg {
    static constexpr int H = H_, NHID = NHID_, NA2 = NA2_, NA1 = NA1_;
    static constexpr int NS = NA2 + NA1;
    static constexpr int C = 1 + 2 * NA2 + NA1;
    static constexpr int TM = 8, PPT = 2, JG = H / TM;
    static constexpr int HQ = H < 32 ? H : 32;     // quadrant edge of the weight-gradient tiling
    static constexpr int JJ = HQ / 4, KK = HQ / 8; // per-lane (j,k) register tile of the G phase
    static constexpr int QW = H / HQ, NQ = QW * QW;

    // small-parameter block shared by both kernels (floats)
    static constexpr int SM_W0 = 0;                          // [3][H]  first-layer weights, axis major
    static constexpr int SM_B0 = SM_W0 + 3 * H;              // [H]
    static constexpr int SM_W0D = SM_B0 + H;                 // [max(NS,1)][H]  W0[:,axis]*inv_sd
    static constexpr int SM_WT1 = SM_W0D + (NS > 0 ? NS : 1) * H;   // [H][H] hidden matrix transposed (k major)
    static constexpr int SM_WR1 = SM_WT1 + (NHID == 2 ? H * H : 0); // [H][H] hidden matrix raw (j major)
    static constexpr int SM_B1 = SM_WR1 + (NHID == 2 ? H * H : 0);  // [H]
    static constexpr int SM_WL = SM_B1 + (NHID == 2 ? H : 0);       // [H] output weights
    static constexpr int SM_BL = SM_WL + H;                         // [4] output bias (+pad)
    static constexpr int SM_PARAMS = SM_BL + 4;

    // forward tile: largest TP in {128, 64, 32} whose footprint allows 2 CTAs per SM when possible
    static constexpr int fwd_floats(int tp) {
        return SM_PARAMS + 3 * tp + (NHID == 2 ? H * C * tp : 0) + JG * C * tp + C * tp;
    }
#ifndef FBP_TPF_CAP
#define FBP_TPF_CAP 64       // measured: 64-point tiles (more, smaller CTAs per SM) beat 128 (127.0 vs 116.4 steps/s)
#endif
    static constexpr bool fwd_ok(int tp) { return tp <= FBP_TPF_CAP && fwd_floats(tp) * 4 <= 110 * 1024 && JG * (tp / PPT) <= 256; }
    static constexpr int TPF = fwd_ok(128) ? 128 : (fwd_ok(64) ? 64 : 32);
    static constexpr int NTF = JG * (TPF / PPT);
    static constexpr int FWD_MINB = !fwd_ok(TPF) ? 1 : (fwd_floats(TPF) * 4 <= 56 * 1024 && NTF <= 128 ? 4 : 2);

    // backward tile
    static constexpr int SM_GRAD = 3 * H + H + (NS > 0 ? NS : 1) * H + (NHID == 2 ? H : 0) + H + 4;   // T, B0, S, B1, WL, BL
    static constexpr int bwd_floats(int tp) {
        return SM_PARAMS + SM_GRAD + 3 * tp + C * tp + NHID * H * (C * tp + 4);
    }
#ifndef FBP_TPB_CAP
#define FBP_TPB_CAP 64
#endif
    static constexpr bool bwd_ok(int tp) { return tp <= FBP_TPB_CAP && bwd_floats(tp) * 4 <= 200 * 1024 && JG * (tp / PPT) <= 256; }
    static constexpr int TPB = bwd_ok(128) ? 128 : (bwd_ok(64) ? 64 : 32);
    static constexpr int NTB = JG * (TPB / PPT);
};


// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a): global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy accesses to shared memory before this fence are ordered before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float sel3(int i, float a, float b, float c) { return i == 0 ? a : (i == 1 ? b : c); }

// window value / first / second derivative per slot for one point
template <class CF>
__device__ __forceinline__ void fast_window(const float z[3], const float isd[3], int xd, float flag, const int axis[FBP_MAX_XD],
                                            float& w, float w1[CF::NS > 0 ? CF::NS : 1], float w2[CF::NA2 > 0 ? CF::NA2 : 1]) {
    float f[3], f1[3], f2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (d < xd) fbp_window_dim(z[d], isd[d], f[d], f1[d], f2[d]);
        else { f[d] = 1.0f; f1[d] = 0.0f; f2[d] = 0.0f; }
    }
    const float fall = f[0] * f[1] * f[2];
    w = flag * fall + (1.0f - flag);
#pragma unroll
    for (int s = 0; s < CF::NS; ++s) {
        const int ax = axis[s];
        const float others = flag * sel3(ax, f[1] * f[2], f[0] * f[2], f[0] * f[1]);
        w1[s] = others * sel3(ax, f1[0], f1[1], f1[2]);
        if (s < CF::NA2) w2[s] = others * sel3(ax, f2[0], f2[1], f2[2]);
    }
}

// Stage the subdomain's parameters in shared memory.
template <class CF, int NT>
__device__ __forceinline__ void fast_load_params(float* sm, const float* __restrict__ prow, int xd, const float isd[3],
                                                 const int axis[FBP_MAX_XD], bool want_raw) {
    constexpr int H = CF::H;
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * H; i += NT) sm[CF::SM_W0 + i] = 0.0f;
    __syncthreads();
    for (int i = tid; i < H * xd; i += NT) {
        int j = i / xd, d = i - j * xd;
        sm[CF::SM_W0 + d * H + j] = prow[i];
    }
    for (int i = tid; i < H; i += NT) sm[CF::SM_B0 + i] = prow[H * xd + i];
    for (int i = tid; i < CF::NS * H; i += NT) {
        int s = i / H, j = i - s * H;
        int ax = axis[s];
        sm[CF::SM_W0D + i] = prow[j * xd + ax] * sel3(ax, isd[0], isd[1], isd[2]);
    }
    int off = H * xd + H;
    if (CF::NHID == 2) {
        for (int i = tid; i < H * H; i += NT) {
            int j = i / H, k = i - j * H;
            float v = prow[off + i];
            sm[CF::SM_WT1 + k * H + j] = v;
            if (want_raw) sm[CF::SM_WR1 + i] = v;
        }
        off += H * H;
        for (int i = tid; i < H; i += NT) sm[CF::SM_B1 + i] = prow[off + i];
        off += H;
    }
    for (int i = tid; i < H; i += NT) sm[CF::SM_WL + i] = prow[off + i];
    if (tid == 0) sm[CF::SM_BL] = prow[off + H];
}

// tanh jets applied in place to the accumulators of one (unit, point): a -> h
template <class CF>
__device__ __forceinline__ void fast_tanh_jets(float (&a)[CF::C]) {
    const float t = fbp_tanh(a[0]);
    const float g = 1.0f - t * t;
    a[0] = t;
#pragma unroll
    for (int s = 0; s < CF::NA2; ++s) {
        const float a1 = a[1 + 2 * s], a2 = a[2 + 2 * s];
        a[1 + 2 * s] = g * a1;
        a[2 + 2 * s] = g * (a2 - 2.0f * t * a1 * a1);
    }
#pragma unroll
    for (int s = 0; s < CF::NA1; ++s) a[1 + 2 * CF::NA2 + s] *= g;
}

// transpose of fast_tanh_jets: h (forward outputs), hb (their cotangents) -> ab (cotangents of the pre-activations)
template <class CF>
__device__ __forceinline__ void fast_tanh_jets_bwd(const float (&h)[CF::C], const float (&hb)[CF::C], float (&ab)[CF::C]) {
    const float t = h[0];
    const float g = 1.0f - t * t;
    float ab0 = g * hb[0];
#pragma unroll
    for (int s = 0; s < CF::NA2; ++s) {
        const float h1 = h[1 + 2 * s], h2 = h[2 + 2 * s];
        const float b1 = hb[1 + 2 * s], b2 = hb[2 + 2 * s];
        ab[2 + 2 * s] = g * b2;
        ab[1 + 2 * s] = g * b1 - 4.0f * t * b2 * h1;
        ab0 -= 2.0f * t * b1 * h1 + 2.0f * b2 * (t * h2 + h1 * h1);
    }
#pragma unroll
    for (int s = 0; s < CF::NA1; ++s) {
        const int c = 1 + 2 * CF::NA2 + s;
        ab[c] = g * hb[c];
        ab0 -= 2.0f * t * hb[c] * h[c];
    }
    ab[0] = ab0;
}

// First layer + tanh jets for TM units x PPT points into acc (registers).
template <class CF, int TP>
__device__ __forceinline__ void fast_layer0(const float* sm, const float* zs, int j0, int p0,
                                            float (&acc)[CF::TM][CF::PPT][CF::C]) {
    constexpr int H = CF::H;
    float z[3][CF::PPT];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float2 v = *reinterpret_cast<const float2*>(zs + d * TP + p0);
        z[d][0] = v.x;
        z[d][1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < CF::TM; ++j) {
        const float w0 = sm[CF::SM_W0 + j0 + j], w1 = sm[CF::SM_W0 + H + j0 + j], w2 = sm[CF::SM_W0 + 2 * H + j0 + j];
        const float b = sm[CF::SM_B0 + j0 + j];
        float wd[CF::NS > 0 ? CF::NS : 1];
#pragma unroll
        for (int s = 0; s < CF::NS; ++s) wd[s] = sm[CF::SM_W0D + s * H + j0 + j];
#pragma unroll
        for (int p = 0; p < CF::PPT; ++p) {
            float a[CF::C];
            a[0] = fmaf(w2, z[2][p], fmaf(w1, z[1][p], fmaf(w0, z[0][p], b)));
#pragma unroll
            for (int s = 0; s < CF::NA2; ++s) { a[1 + 2 * s] = wd[s]; a[2 + 2 * s] = 0.0f; }
#pragma unroll
            for (int s = 0; s < CF::NA1; ++s) a[1 + 2 * CF::NA2 + s] = wd[CF::NA2 + s];
            fast_tanh_jets<CF>(a);
#pragma unroll
            for (int c = 0; c < CF::C; ++c) acc[j][p][c] = a[c];
        }
    }
}

// Packed FP32 FMA of Blackwell (PTX fma.rn.f32x2 -> SASS FFMA2): two FMAs per issue slot.  The GEMM loops are
// issue-bound (ncu: FFMA ~70 % of the issued warp instructions), and the activations already arrive as float2 point
// pairs from LDS.64, so one FFMA2 with the weight as scalar-broadcast operand replaces two FFMA.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

// acc[j][p][c] = sum_k Wm[k*H + j0 + j] * act[k*RS + c*TP + p0 + p]   (Wm: k-major matrix in shared memory)
template <class CF, int TP, int RS>
__device__ __forceinline__ void fast_gemm(const float* __restrict__ Wm, const float* __restrict__ act, int j0, int p0,
                                          float (&acc)[CF::TM][CF::PPT][CF::C]) {
    constexpr int H = CF::H;
    static_assert(CF::PPT == 2, "the packed accumulators hold the two points of a thread");
    float2 accp[CF::TM][CF::C];
#pragma unroll
    for (int j = 0; j < CF::TM; ++j)
#pragma unroll
        for (int c = 0; c < CF::C; ++c) accp[j][c] = make_float2(0.0f, 0.0f);
    constexpr int kUnroll = FBP_GEMM_UNROLL;
#pragma unroll kUnroll
    for (int k = 0; k < H; ++k) {
        const float4 wa = *reinterpret_cast<const float4*>(Wm + k * H + j0);
        const float4 wb = *reinterpret_cast<const float4*>(Wm + k * H + j0 + 4);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        float2 av[CF::C];
#pragma unroll
        for (int c = 0; c < CF::C; ++c) av[c] = *reinterpret_cast<const float2*>(act + k * RS + c * TP + p0);
#pragma unroll
        for (int j = 0; j < CF::TM; ++j)
#pragma unroll
            for (int c = 0; c < CF::C; ++c) accp[j][c] = ffma2(make_float2(w[j], w[j]), av[c], accp[j][c]);
    }
#pragma unroll
    for (int j = 0; j < CF::TM; ++j)
#pragma unroll
        for (int c = 0; c < CF::C; ++c) {
            acc[j][0][c] = accp[j][c].x;
            acc[j][1][c] = accp[j][c].y;
        }
}

template <class CF, int TP, int RS>
__device__ __forceinline__ void fast_store_act(float* act, int j0, int p0, const float (&acc)[CF::TM][CF::PPT][CF::C]) {
#pragma unroll
    for (int j = 0; j < CF::TM; ++j)
#pragma unroll
        for (int c = 0; c < CF::C; ++c)
            *reinterpret_cast<float2*>(act + (j0 + j) * RS + c * TP + p0) = make_float2(acc[j][0][c], acc[j][1][c]);
}

// =====================================================================================================
// forward
// =====================================================================================================
template <class CF>
__global__ void __launch_bounds__(CF::NTF, CF::FWD_MINB) fast_forward_kernel(FastArgs a) {
    constexpr int H = CF::H, C = CF::C, TM = CF::TM, PPT = CF::PPT, JG = CF::JG, TP = CF::TPF, NT = CF::NTF;
    constexpr int RS = C * TP;
    extern __shared__ __align__(16) float sm[];
    float* zs = sm + CF::SM_PARAMS;                       // [3][TP]
    float* act = zs + 3 * TP;                             // [H][RS]      (NHID == 2)
    float* part = act + (CF::NHID == 2 ? H * RS : 0);     // [JG][C][TP]
    float* outN = part + JG * C * TP;                     // [TP][C] in external component order

    const int tid = threadIdx.x;
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
    fast_load_params<CF, NT>(sm, a.params + (int64_t)im * a.P, xd, isd, a.axis, false);
    __syncthreads();

    const int jg = tid / (TP / PPT), pp = tid % (TP / PPT);
    const int j0 = jg * TM, p0 = pp * PPT;

    constexpr bool TMA_SAVE = (CF::NHID == 2) && (CF::TPF == CF::TPB);   // forward tile == reverse tile: bulk stores
    // software prefetch of the next tile's points (pair -> point index -> coordinates: two dependent global loads)
    int pf_pt = 0;
    float pf_x[3] = {0.0f, 0.0f, 0.0f};
    auto load_idx = [&](int t0n) {
        if (tid < TP && t0n < count) pf_pt = a.spair_point[first + t0n + (tid < min(TP, count - t0n) ? tid : 0)];
    };
    auto load_val = [&](int t0n) {
        if (tid < TP && t0n < count) {
#pragma unroll
            for (int d = 0; d < 3; ++d) pf_x[d] = d < xd ? a.x[(int64_t)pf_pt * xd + d] : 0.0f;
        }
    };
    load_idx(0);
    load_val(0);
    for (int t0 = 0; t0 < count; t0 += TP) {
        const int cnt = min(TP, count - t0);
        if (TMA_SAVE && tid == 0) tma_store_wait_read();     // previous tile's bulk stores have read `act`
        if (tid < TP) {
#pragma unroll
            for (int d = 0; d < 3; ++d) zs[d * TP + tid] = d < xd ? (pf_x[d] - mu[d]) * isd[d] : 0.0f;
        }
        __syncthreads();
        load_idx(t0 + TP);

        float acc[TM][PPT][C];
        fast_layer0<CF, TP>(sm, zs, j0, p0, acc);
        if (CF::NHID == 2) {
            fast_store_act<CF, TP, RS>(act, j0, p0, acc);
            __syncthreads();
            fast_gemm<CF, TP, RS>(sm + CF::SM_WT1, act, j0, p0, acc);
#pragma unroll
            for (int j = 0; j < TM; ++j) {
                const float b = sm[CF::SM_B1 + j0 + j];
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    acc[j][p][0] += b;
                    fast_tanh_jets<CF>(acc[j][p]);
                }
            }
            if (a.cache != nullptr && !(TMA_SAVE && cnt == TP)) {
                // save the hidden jets for the reverse kernel in ITS tile layout: [H][C][cntb] at (first+t0b)*H*C
                constexpr int TPB = CF::TPB;
                const int off = t0 + p0;                         // pair offset inside the work item (even)
                if (off < count) {
                    const int t0b = (off / TPB) * TPB;
                    const int cntb = min(TPB, count - t0b);
                    float* cb = a.cache + (int64_t)(first + t0b) * (H * C) + (off - t0b);
                    const bool two = (off + 1 < count);
                    if (cntb == TPB) {                           // full reverse tile: rows are 8-byte aligned
#pragma unroll
                        for (int j = 0; j < TM; ++j)
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                *reinterpret_cast<float2*>(cb + ((j0 + j) * C + c) * TPB) = make_float2(acc[j][0][c], acc[j][1][c]);
                    } else {
#pragma unroll
                        for (int j = 0; j < TM; ++j)
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                cb[((j0 + j) * C + c) * cntb] = acc[j][0][c];
                                if (two) cb[((j0 + j) * C + c) * cntb + 1] = acc[j][1][c];
                            }
                    }
                }
            }
        }
        load_val(t0 + TP);
        // output layer (ud = 1): partial dot over this thread's TM units, reduced over the JG groups below
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int j = 0; j < TM; ++j) {
                const float wl = sm[CF::SM_WL + j0 + j];
                s0 = fmaf(wl, acc[j][0][c], s0);
                s1 = fmaf(wl, acc[j][1][c], s1);
            }
            *reinterpret_cast<float2*>(part + (jg * C + c) * TP + p0) = make_float2(s0, s1);
        }
        __syncthreads();
        const bool bulk_save = TMA_SAVE && a.cache != nullptr && cnt == TP;
        if (bulk_save) fast_store_act<CF, TP, RS>(act, j0, p0, acc);   // every GEMM read of act is done: h1 over h0

        if (tid < TP) {
            float u[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float r = (c == 0) ? sm[CF::SM_BL] : 0.0f;
#pragma unroll
                for (int g = 0; g < JG; ++g) r += part[(g * C + c) * TP + tid];
                u[c] = un_sd * r;
            }
            u[0] += un_mu;
            float z[3] = {zs[tid], zs[TP + tid], zs[2 * TP + tid]};
            float w, w1[CF::NS > 0 ? CF::NS : 1], w2[CF::NA2 > 0 ? CF::NA2 : 1];
            fast_window<CF>(z, isd, xd, flag, a.axis, w, w1, w2);
            float* o = outN + tid * C;
            o[a.ext[0]] = u[0] * w;
#pragma unroll
            for (int s = 0; s < CF::NA2; ++s) {
                const float u1 = u[1 + 2 * s], u2 = u[2 + 2 * s];
                o[a.ext[1 + 2 * s]] = u1 * w + u[0] * w1[s];
                o[a.ext[2 + 2 * s]] = u2 * w + 2.0f * u1 * w1[s] + u[0] * w2[s];
            }
#pragma unroll
            for (int s = 0; s < CF::NA1; ++s) {
                const int c = 1 + 2 * CF::NA2 + s;
                o[a.ext[c]] = u[c] * w + u[0] * w1[CF::NA2 + s];
            }
        }
        __syncthreads();
        if (bulk_save && tid == 0) {
            // the tile's [H][C][TP] block is contiguous both in shared memory and in the cache: H bulk stores
            fence_proxy_async();
            float* cdst = a.cache + (int64_t)(first + t0) * (H * C);
            for (int k = 0; k < H; ++k) tma_bulk_s2g(cdst + k * RS, act + k * RS, (uint32_t)(RS * sizeof(float)));
            tma_store_commit();
        }
        float* dst = a.pair_out + (int64_t)(first + t0) * C;
        for (int i = tid; i < cnt * C; i += NT) dst[i] = outN[i];
    }
    if (TMA_SAVE && tid == 0) tma_store_wait_read();
}

// =====================================================================================================
// reverse
// =====================================================================================================
template <class CF, bool USE_CACHE>
__global__ void __launch_bounds__(CF::NTB, (CF::bwd_floats(CF::TPB) * 4 <= 110 * 1024 && CF::NTB <= 128) ? 2 : 1)
fast_backward_kernel(FastArgs a) {
    constexpr int H = CF::H, C = CF::C, TM = CF::TM, PPT = CF::PPT, JG = CF::JG, TP = CF::TPB, NT = CF::NTB;
    constexpr int NS = CF::NS, NA2 = CF::NA2, NA1 = CF::NA1;
    constexpr int RS = C * TP + 4;
    constexpr int NW = NT / 32;
    constexpr int JJ = CF::JJ, KK = CF::KK, HQ = CF::HQ, QW = CF::QW, NQ = CF::NQ;
    constexpr int PG = NW / NQ;          // point groups of the G phase
    constexpr int PPG = TP / PG;         // points per group
    static_assert(NW % NQ == 0 && PG >= 1 && PPG % 4 == 0, "bad G-phase tiling");
    static_assert(NT >= TP, "point stage needs NT >= TP");

    extern __shared__ __align__(16) float sm[];
    float* gs = sm + CF::SM_PARAMS;                  // small-layer gradient accumulators
    float* gT = gs;                                  // [3][H]  sum abar0 * z_d
    float* gB0 = gT + 3 * H;                         // [H]
    float* gS = gB0 + H;                             // [NS][H] sum abar_first per slot
    float* gB1 = gS + (NS > 0 ? NS : 1) * H;         // [H] (NHID == 2)
    float* gWL = gB1 + (CF::NHID == 2 ? H : 0);      // [H]
    float* gBL = gWL + H;                            // [4]
    float* zs = gs + CF::SM_GRAD;                    // [3][TP]
    float* rb = zs + 3 * TP;                         // [C][TP]   cotangent of the output-layer jets
    float* act0 = rb + C * TP;                       // [H][RS]
    float* act1 = act0 + H * RS;                     // [H][RS]   (NHID == 2)
    __shared__ __align__(8) uint64_t cache_bar;      // mbarrier of the TMA loads into act1
    constexpr bool CACHED = USE_CACHE && CF::NHID == 2;
    constexpr uint32_t ROW_BYTES = C * TP * sizeof(float);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const float flag = ss[2 * xd], un_sd = ss[2 * xd + 2];
    fast_load_params<CF, NT>(sm, a.params + (int64_t)im * a.P, xd, isd, a.axis, true);
    for (int i = tid; i < CF::SM_GRAD; i += NT) gs[i] = 0.0f;
    if (CACHED && tid == 0) {
        mbar_init(&cache_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t cache_phase = 0;
    // TMA prefetch of a FULL tile's saved hidden jets into act1: H bulk copies of one [C][TP] row block each
    auto prefetch_tile = [&](int t0n) {
        if (CACHED && tid == 0 && t0n < count && count - t0n >= TP) {
            fence_proxy_async();
            mbar_expect_tx(&cache_bar, (uint32_t)H * ROW_BYTES);
            const float* src = a.cache + (int64_t)(first + t0n) * (H * C);
            for (int k = 0; k < H; ++k) tma_bulk_g2s(act1 + k * RS, src + k * (C * TP), ROW_BYTES, &cache_bar);
        }
    };
    prefetch_tile(0);

    const int jg = tid / (TP / PPT), pp = tid % (TP / PPT);
    const int j0 = jg * TM, p0 = pp * PPT;
    float* actL = CF::NHID == 2 ? act1 : act0;       // activations of the last hidden layer

    // G-phase ownership
    const int qd = warp % NQ, pg = warp / NQ;
    const int qj = qd / QW, qk = qd % QW;
    const int jt = lane >> 3, kt = lane & 7;
#ifndef FBP_G_FFMA2
#define FBP_G_FFMA2 1
#endif
    // weight-gradient accumulators: with FFMA2 each (j,k) keeps an (even point, odd point) pair, summed at the end
    float2 gacc[JJ][KK];
    float bacc[JJ];
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) {
        bacc[jj] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) gacc[jj][kk] = make_float2(0.0f, 0.0f);
    }

    // output-layer gradient partials live in registers for the whole work item (reduced once at the end):
    // wlacc[j] = sum over this thread's points of sum_c rbar_c h_c[j] ; blacc = sum of rbar_0 (threads < TP)
    float wlacc[TM];
#pragma unroll
    for (int j = 0; j < TM; ++j) wlacc[j] = 0.0f;
    float blacc = 0.0f;
    // first-layer gradients folded into the layer-0 tanh transpose when a thread's own activation slots can hold its
    // partial sums (bias, 3 coordinates, NS slots) — otherwise the warp-row reduction below is used
    constexpr int NQ0 = 4 + NS;
    constexpr bool FOLD0 = (CF::NHID == 2) && (NQ0 <= 2 * C);

    // software prefetch of the next tile's per-point inputs (two dependent global loads: pair -> point/row index ->
    // coordinates / row cotangent): indices are requested at the start of the long GEMM phases, values after them,
    // and both are consumed at the next tile's S0, so their latency never sits on the critical path
    int pf_pt = 0, pf_row = 0;
    float pf_x[3] = {0.0f, 0.0f, 0.0f}, pf_g[C];
#pragma unroll
    for (int c = 0; c < C; ++c) pf_g[c] = 0.0f;
    auto load_idx = [&](int t0n) {
        if (tid < TP && t0n < count) {
            const int cn = min(TP, count - t0n);
            const int pi = first + t0n + (tid < cn ? tid : 0);
            pf_pt = a.spair_point[pi];
            pf_row = a.spair_row[pi];
        }
    };
    auto load_val = [&](int t0n) {
        if (tid < TP && t0n < count) {
#pragma unroll
            for (int d = 0; d < 3; ++d) pf_x[d] = d < xd ? a.x[(int64_t)pf_pt * xd + d] : 0.0f;
            const float* gr = a.grow + (int64_t)pf_row * C;
#pragma unroll
            for (int c = 0; c < C; ++c) pf_g[c] = gr[a.ext[c]];
        }
    };
    load_idx(0);
    load_val(0);

    for (int t0 = 0; t0 < count; t0 += TP) {
        const int cnt = min(TP, count - t0);
        // ---- S0: points, window jets, cotangent of the output jets ------------------------------------
        if (tid < TP) {
            const bool valid = tid < cnt;
            float z[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                z[d] = d < xd ? (pf_x[d] - mu[d]) * isd[d] : 0.0f;
                zs[d * TP + tid] = z[d];
            }
            float w, w1[NS > 0 ? NS : 1], w2[NA2 > 0 ? NA2 : 1];
            fast_window<CF>(z, isd, xd, flag, a.axis, w, w1, w2);
            float G[C];
#pragma unroll
            for (int c = 0; c < C; ++c) G[c] = valid ? pf_g[c] : 0.0f;
            float ub0 = G[0] * w;
#pragma unroll
            for (int s = 0; s < NA2; ++s) {
                const float G1 = G[1 + 2 * s], G2 = G[2 + 2 * s];
                ub0 += G1 * w1[s] + G2 * w2[s];
                rb[(1 + 2 * s) * TP + tid] = un_sd * (G1 * w + 2.0f * G2 * w1[s]);
                rb[(2 + 2 * s) * TP + tid] = un_sd * (G2 * w);
            }
#pragma unroll
            for (int s = 0; s < NA1; ++s) {
                const int c = 1 + 2 * NA2 + s;
                ub0 += G[c] * w1[NA2 + s];
                rb[c * TP + tid] = un_sd * (G[c] * w);
            }
            rb[tid] = un_sd * ub0;
            blacc += un_sd * ub0;
        }
        __syncthreads();
        load_idx(t0 + TP);

        // ---- forward recompute -------------------------------------------------------------------------
        float acc[TM][PPT][C];
        fast_layer0<CF, TP>(sm, zs, j0, p0, acc);
        fast_store_act<CF, TP, RS>(act0, j0, p0, acc);
        if (CF::NHID == 2 && !CACHED) {
            __syncthreads();
            fast_gemm<CF, TP, RS>(sm + CF::SM_WT1, act0, j0, p0, acc);
#pragma unroll
            for (int j = 0; j < TM; ++j) {
                const float b = sm[CF::SM_B1 + j0 + j];
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    acc[j][p][0] += b;
                    fast_tanh_jets<CF>(acc[j][p]);
                }
            }
            fast_store_act<CF, TP, RS>(act1, j0, p0, acc);
        }
        if (CACHED) {
            if (cnt == TP) {                               // full tile: wait for the TMA bytes (issued one tile ahead)
                mbar_wait(&cache_bar, cache_phase);
                cache_phase ^= 1;
            } else {                                       // partial tile: plain loads of [H][C][cnt], zero padding
                const float* src = a.cache + (int64_t)(first + t0) * (H * C);
                for (int i = tid; i < H * C * TP; i += NT) {
                    const int row = i / TP, p = i - row * TP;          // row = k*C + c
                    const int k = row / C, c = row - k * C;
                    act1[k * RS + c * TP + p] = p < cnt ? src[(int64_t)row * cnt + p] : 0.0f;
                }
            }
        }
        __syncthreads();

        // ---- tanh transpose of the last hidden layer, in place: actL <- abar ----------------------------
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            const float wl = sm[CF::SM_WL + j0 + j];
            float2 hv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) hv[c] = *reinterpret_cast<const float2*>(actL + (j0 + j) * RS + c * TP + p0);
            float2 rv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) rv[c] = *reinterpret_cast<const float2*>(rb + c * TP + p0);
            float h0[C], h1[C], b0[C], b1[C], o0[C], o1[C];
#pragma unroll
            for (int c = 0; c < C; ++c) { h0[c] = hv[c].x; h1[c] = hv[c].y; b0[c] = wl * rv[c].x; b1[c] = wl * rv[c].y; }
            {   // output-layer weight gradient: sum_c rbar_c * h_c for this thread's two points
                float2 d = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int c = 0; c < C; ++c) d = ffma2(rv[c], hv[c], d);
                wlacc[j] += d.x + d.y;
            }
            fast_tanh_jets_bwd<CF>(h0, b0, o0);
            fast_tanh_jets_bwd<CF>(h1, b1, o1);
#pragma unroll
            for (int c = 0; c < C; ++c)
                *reinterpret_cast<float2*>(actL + (j0 + j) * RS + c * TP + p0) = make_float2(o0[c], o1[c]);
        }
        __syncthreads();
        if (CF::NHID == 1) load_val(t0 + TP);

        if (CF::NHID == 2) {
            // ---- G: Wbar1[j][k] += sum_{c,p} abar1[j][c][p] * h0[k][c][p] ; bbar1[j] += sum_p abar1[j][0][p]
            {
                const float* ab = act1 + (qj * HQ + jt) * RS;
                const float* hp = act0 + (qk * HQ + kt) * RS;
                for (int pq = pg * PPG; pq < (pg + 1) * PPG; pq += 4) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float4 av[JJ], hv[KK];
#pragma unroll
                        for (int jj = 0; jj < JJ; ++jj) av[jj] = *reinterpret_cast<const float4*>(ab + (4 * jj) * RS + c * TP + pq);
#pragma unroll
                        for (int kk = 0; kk < KK; ++kk) hv[kk] = *reinterpret_cast<const float4*>(hp + (8 * kk) * RS + c * TP + pq);
#if FBP_G_FFMA2
                        // two passes (points 0-1, then 2-3) so that the two updates of one accumulator are JJ*KK
                        // instructions apart instead of back to back
#pragma unroll
                        for (int jj = 0; jj < JJ; ++jj)
#pragma unroll
                            for (int kk = 0; kk < KK; ++kk)
                                gacc[jj][kk] = ffma2(make_float2(av[jj].x, av[jj].y), make_float2(hv[kk].x, hv[kk].y), gacc[jj][kk]);
#pragma unroll
                        for (int jj = 0; jj < JJ; ++jj)
#pragma unroll
                            for (int kk = 0; kk < KK; ++kk)
                                gacc[jj][kk] = ffma2(make_float2(av[jj].z, av[jj].w), make_float2(hv[kk].z, hv[kk].w), gacc[jj][kk]);
#endif
#pragma unroll
                        for (int jj = 0; jj < JJ; ++jj) {
#pragma unroll
                            for (int kk = 0; kk < KK; ++kk) {
#if FBP_G_FFMA2
#else
                                float g = gacc[jj][kk].x;
                                g = fmaf(av[jj].x, hv[kk].x, g);
                                g = fmaf(av[jj].y, hv[kk].y, g);
                                g = fmaf(av[jj].z, hv[kk].z, g);
                                g = fmaf(av[jj].w, hv[kk].w, g);
                                gacc[jj][kk].x = g;
#endif
                            }
                            if (c == 0) bacc[jj] += (av[jj].x + av[jj].y) + (av[jj].z + av[jj].w);
                        }
                    }
                }
            }
            // ---- D: hbar0[k][(c,p)] = sum_j W1[j][k] abar1[j][(c,p)]  (raw matrix is "j-major" = k contiguous)
            fast_gemm<CF, TP, RS>(sm + CF::SM_WR1, act1, j0, p0, acc);
            __syncthreads();      // every read of act0 (G) and of act1 (G, D) is done
            prefetch_tile(t0 + TP);
            load_val(t0 + TP);
            // ---- tanh transpose of layer 0, in place: act0 <- abar0 (or, folded, this thread's partial sums of the
            //      first-layer gradients over its two points, stored in its own activation slots)
            const float2 zv0 = *reinterpret_cast<const float2*>(zs + p0);
            const float2 zv1 = *reinterpret_cast<const float2*>(zs + TP + p0);
            const float2 zv2 = *reinterpret_cast<const float2*>(zs + 2 * TP + p0);
#pragma unroll
            for (int j = 0; j < TM; ++j) {
                float2 hv[C];
#pragma unroll
                for (int c = 0; c < C; ++c) hv[c] = *reinterpret_cast<const float2*>(act0 + (j0 + j) * RS + c * TP + p0);
                float h0[C], h1[C], o0[C], o1[C];
#pragma unroll
                for (int c = 0; c < C; ++c) { h0[c] = hv[c].x; h1[c] = hv[c].y; }
                fast_tanh_jets_bwd<CF>(h0, acc[j][0], o0);
                fast_tanh_jets_bwd<CF>(h1, acc[j][1], o1);
                if (FOLD0) {
                    float q[NQ0 + 1];
                    q[0] = o0[0] + o1[0];                                    // bias
                    q[1] = fmaf(o0[0], zv0.x, o1[0] * zv0.y);                // coordinates
                    q[2] = fmaf(o0[0], zv1.x, o1[0] * zv1.y);
                    q[3] = fmaf(o0[0], zv2.x, o1[0] * zv2.y);
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const int c = s < NA2 ? 1 + 2 * s : 1 + 2 * NA2 + (s - NA2);
                        q[4 + s] = o0[c] + o1[c];
                    }
                    // slot q of row j = element (component q/2, point p0 + q%2) of this thread's own block
#pragma unroll
                    for (int t = 0; t < NQ0; ++t) act0[(j0 + j) * RS + (t >> 1) * TP + p0 + (t & 1)] = q[t];
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        *reinterpret_cast<float2*>(act0 + (j0 + j) * RS + c * TP + p0) = make_float2(o0[c], o1[c]);
                }
            }
            __syncthreads();
            if (FOLD0) {
                // one thread per (unit j, quantity t): sum the partials of the TP/2 point pairs in fixed order
                for (int o = tid; o < H * NQ0; o += NT) {
                    const int j = o / NQ0, t = o - j * NQ0;
                    const float* src = act0 + j * RS + (t >> 1) * TP + (t & 1);
                    float v = 0.0f;
#pragma unroll 8
                    for (int pq = 0; pq < TP / 2; ++pq) v += src[2 * pq];
                    if (t == 0) gB0[j] += v;
                    else if (t < 4) gT[(t - 1) * H + j] += v;
                    else gS[(t - 4) * H + j] += v;
                }
                __syncthreads();
            }
        }

        // ---- first-layer gradients: rows j over warps, points over lanes (when not folded above) ---------
        if (!FOLD0)
        for (int j = warp; j < H; j += NW) {
            float t0s = 0.0f, t1s = 0.0f, t2s = 0.0f, bs = 0.0f;
            float ssl[NS > 0 ? NS : 1];
#pragma unroll
            for (int s = 0; s < NS; ++s) ssl[s] = 0.0f;
            for (int p = lane; p < TP; p += 32) {
                const float a0 = act0[j * RS + p];
                t0s = fmaf(a0, zs[p], t0s);
                t1s = fmaf(a0, zs[TP + p], t1s);
                t2s = fmaf(a0, zs[2 * TP + p], t2s);
                bs += a0;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int c = s < NA2 ? 1 + 2 * s : 1 + 2 * NA2 + (s - NA2);
                    ssl[s] += act0[j * RS + c * TP + p];
                }
            }
            t0s = fbp_warp_sum(t0s); t1s = fbp_warp_sum(t1s); t2s = fbp_warp_sum(t2s); bs = fbp_warp_sum(bs);
#pragma unroll
            for (int s = 0; s < NS; ++s) ssl[s] = fbp_warp_sum(ssl[s]);
            if (lane == 0) {
                gT[j] += t0s; gT[H + j] += t1s; gT[2 * H + j] += t2s; gB0[j] += bs;
#pragma unroll
                for (int s = 0; s < NS; ++s) gS[s * H + j] += ssl[s];
            }
        }
        if (!FOLD0) __syncthreads();
    }

    // ---- output-layer gradients: reduce the per-thread partials over the threads of each unit group, fixed order
    {
        constexpr int GW = (TP / PPT) < 32 ? (TP / PPT) : 32;      // lanes of a warp that share a unit group
        constexpr int NG = (TP / PPT) / GW;                        // such lane groups per unit group
        float* red = zs;                                           // [JG*NG][TM] + [TP/32] scratch (zs/rb are free now)
        static_assert(JG * NG * TM + 8 <= 3 * TP + C * TP, "reduction scratch does not fit");
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            float v = wlacc[j];
#pragma unroll
            for (int o = GW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid % GW) == 0) red[(tid / GW) * TM + j] = v;
        }
        float bv = fbp_warp_sum(blacc);                            // threads >= TP hold 0
        if (lane == 0) red[JG * NG * TM + warp] = bv;
        __syncthreads();
        for (int k = tid; k < H; k += NT) {
            const int g = k / TM, j = k - g * TM;
            float v = 0.0f;
#pragma unroll
            for (int i = 0; i < NG; ++i) v += red[(g * NG + i) * TM + j];
            gWL[k] += v;
        }
        if (tid == 0) {
            float v = 0.0f;
            for (int w = 0; w < NW; ++w) v += red[JG * NG * TM + w];
            gBL[0] += v;
        }
        __syncthreads();
    }

    // ---- write this work item's partial gradients -----------------------------------------------------------
    float* gp = a.gpart + (int64_t)item * a.P;
    int off = 0;
    // W0 (H x xd) and b0
    for (int i = tid; i < H * xd; i += NT) {
        const int j = i / xd, d = i - j * xd;
        float v = gT[d * H + j];
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (a.axis[s] == d) v = fmaf(sel3(d, isd[0], isd[1], isd[2]), gS[s * H + j], v);
        gp[i] = v;
    }
    for (int i = tid; i < H; i += NT) gp[H * xd + i] = gB0[i];
    off = H * xd + H;
    if constexpr (CF::NHID == 2) {
        // reduce the register accumulators over the PG point groups through shared memory (reusing the activation
        // buffers), one group per round in fixed order
        static_assert(NQ * (JJ * KK + JJ) * 32 <= 2 * H * RS, "stage does not fit in the activation buffers");
        float* stage = act0;                                   // [NQ][JJ*KK][32]
        float* stageB = act0 + NQ * JJ * KK * 32;              // [NQ][JJ][32]
        for (int g = 0; g < PG; ++g) {
            if (pg == g) {
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
#pragma unroll
                    for (int kk = 0; kk < KK; ++kk) {
                        const int idx = (qd * JJ * KK + jj * KK + kk) * 32 + lane;
                        stage[idx] = (g == 0 ? 0.0f : stage[idx]) + (gacc[jj][kk].x + gacc[jj][kk].y);
                    }
                    const int ib = (qd * JJ + jj) * 32 + lane;
                    stageB[ib] = (g == 0 ? 0.0f : stageB[ib]) + bacc[jj];
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < H * H; i += NT) {
            const int j = i / H, k = i - j * H;
            const int qj_ = j / HQ, jr = j % HQ, qk_ = k / HQ, kr = k % HQ;
            const int ln = (jr & 3) * 8 + (kr & 7);
            const int jj = jr >> 2, kk = kr >> 3;
            const int q = qj_ * QW + qk_;
            gp[off + i] = stage[(q * JJ * KK + jj * KK + kk) * 32 + ln];
        }
        off += H * H;
        for (int j = tid; j < H; j += NT) {
            const int qj_ = j / HQ, jr = j % HQ;
            const int ln = (jr & 3) * 8;           // lane with kt == 0
            const int q = qj_ * QW;                // quadrant with qk == 0
            gp[off + j] = stageB[(q * JJ + (jr >> 2)) * 32 + ln];
        }
        off += H;
    }
    for (int i = tid; i < H; i += NT) gp[off + i] = gWL[i];
    if (tid == 0) gp[off + H] = gBL[0];
}

// sums the partial gradients of every subdomain's work items in fixed order
__global__ void fast_grad_reduce_kernel(const float* __restrict__ gpart, const int32_t* __restrict__ sub_item_off,
                                        int m_active, int P, float* __restrict__ grads, int accumulate);

// per-translation-unit launchers (one TU per hidden width, see fbp_fast_inst_*.cu)
int fbp_fast_launch_h16(int nhid, int na2, int na1, bool backward, const FastArgs& a, int grid, cudaStream_t st);
int fbp_fast_launch_h32(int nhid, int na2, int na1, bool backward, const FastArgs& a, int grid, cudaStream_t st);
int fbp_fast_launch_h64(int nhid, int na2, int na1, bool backward, const FastArgs& a, int grid, cudaStream_t st);

template <class CF>
int fast_launch_one(bool backward, const FastArgs& a, int grid, cudaStream_t st) {
    if (!backward) {
        constexpr int TP = CF::TPF;
        constexpr size_t bytes = sizeof(float) * CF::fwd_floats(TP);
        static bool configured = false;
        if (!configured) {
            FBP_CHECK_CUDA(cudaFuncSetAttribute(fast_forward_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            configured = true;
        }
        fast_forward_kernel<CF><<<grid, CF::NTF, bytes, st>>>(a);
    } else {
        constexpr int TP = CF::TPB;
        constexpr size_t bytes = sizeof(float) * CF::bwd_floats(TP);
        static bool configured = false;
        if (!configured) {
            FBP_CHECK_CUDA(cudaFuncSetAttribute(fast_backward_kernel<CF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            FBP_CHECK_CUDA(cudaFuncSetAttribute(fast_backward_kernel<CF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            configured = true;
        }
        if (a.cache != nullptr && CF::NHID == 2) fast_backward_kernel<CF, true><<<grid, CF::NTB, bytes, st>>>(a);
        else fast_backward_kernel<CF, false><<<grid, CF::NTB, bytes, st>>>(a);
    }
    FBP_LAUNCH_CHECK();
    return 0;
}

#define FBP_FAST_JET_SWITCH(H, NHID)                                                                        \
    switch (na2 * 4 + na1) {                                                                                \
        case 0: return fast_launch_one<FastCfg<H, NHID, 0, 0>>(backward, a, grid, st);                      \
        case 1: return fast_launch_one<FastCfg<H, NHID, 0, 1>>(backward, a, grid, st);                      \
        case 4: return fast_launch_one<FastCfg<H, NHID, 1, 0>>(backward, a, grid, st);                      \
        case 5: return fast_launch_one<FastCfg<H, NHID, 1, 1>>(backward, a, grid, st);                      \
        case 8: return fast_launch_one<FastCfg<H, NHID, 2, 0>>(backward, a, grid, st);                      \
        case 12: return fast_launch_one<FastCfg<H, NHID, 3, 0>>(backward, a, grid, st);                     \
        default: break;                                                                                     \
    }
