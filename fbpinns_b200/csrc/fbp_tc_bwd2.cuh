// tcgen05 ("tensor") family, reverse kernel, second generation: EVERY GEMM of the reverse pass on the tensor core,
// including the weight gradient of the hidden matrix (the FFMA2 gradient warps of fbp_tc_bwd.cuh were the critical path:
// 5.21 ms with them, 3.34 ms without, profiles/r2a_tc_bringup.md).
//
//   warps 0 .. 4 NG - 1  "point warps": thread (g, p) = point row p of the 128-pair tile (= TMEM lane), units [UPT g, UPT (g + 1)), UPT = 32 / NG
//   warp  4 NG           "MMA warp": one elected lane issues every tcgen05.mma of the CTA, driven by mbarriers
//   (three more warps complete its warpgroup: setmaxnreg moves their registers to the point warps: 32 vs 112 per thread at NG = 4, the default)
//
// The first layer is linear in the normalised point, so its jets factor per unit k:  h1_0 = t,  h1_s = g kappa_s[k],
// h1_ss = -2 kappa_s[k]^2 (t g)   with t = tanh(a0), g = 1 - t^2, kappa_s[k] = W0[k][axis s] / sd.  That is used three times:
//   MMA 1   a2 = h1 W1^T       A operands (tensor memory, hi/lo): t, g kappa_s, t g      B: W1 and W1 diag(-2 kappa_s^2)
//   MMA 3   hbar1 = abar2 W1   A operands: abar2 (C components, hi/lo)   B: W1^T, diag(kappa_s) W1^T, diag(-2 kappa_s^2) W1^T
//           accumulated into THREE blocks  D_t, D_g, D_tg  so that  abar0 = g D_t - 2 t g D_g + g (1 - 3 t^2) D_tg
//   G       the weight gradient contracts over POINTS (= TMEM lanes), so both operands come from shared memory:
//           G_t = abar2_0^T t,  G_g[s] = abar2_s^T g,  G_tg[s] = abar2_ss^T (t g)     (raw products, 32 x 32 each), then
//           W1bar[j][k] = G_t + sum_s kappa_s[k] G_g[s] - 2 sum_s kappa_s[k]^2 G_tg[s]                    (end of item)
//           kappa_bar_s[k] = sum_j W1[j][k] (G_g[s][j][k] - 4 kappa_s[k] G_tg[s][j][k])   -> the derivative path of W0bar
//           The point threads write hi/lo images of abar2 and (t, g, t g) for their quarter tile (32 points) into a ring
//           slot in shared memory (K-major, K = point, padded K step: conflict-free 32-lane stores, fbp_tc.cuh GK_*);
//           the hi/lo images of two components are stacked on M = 128 rows, so that ONE accumulation
//           [x_hi; y_hi; x_lo; y_lo] (g_hi + g_lo) carries the whole 3xTF32 product (the fourth term lo*lo is harmless);
//           The tensor core accumulates with truncation, which shows after a few hundred chained MMAs (measured: 9e-6 of the
//           largest gradient entry against 2.5e-6 for the FP32 kernels when G stayed in tensor memory for a whole work
//           item), so every tile starts fresh accumulators (32 chained MMAs) and the point threads add the previous
//           tile's G into registers (round to nearest) while the tensor core works on the next one.
//
// TMEM columns (C = 5):  [0,320) A operands (MMA 1 uses 256 of them, MMA 3 all), D1 = a2 at [256,416) (its first 64 columns
// alias the lo parts of abar2's last two components: only ever touched by the thread that owns the same lane and units,
// read before written), D3 at [320,416), G at [416,512).
// Shared memory: parameters 9 KB, B operands (3 + 5 variants, hi/lo) 64 KB, two ring slots of 72 KB.
#pragma once
#include "fbp_tc_bwd.cuh"

namespace fbptc {

// NG = unit groups per point row: thread (g, p) owns point row p and the hidden units [UPT g, UPT (g + 1)).  NG = 2: 8 point
// warps of 16 units per thread (224 registers); NG = 4: 16 point warps of 8 units (112 registers, 4 warps per scheduler
// to hide the MMA / TMEM / MUFU latencies the 8-warp version is bound by: 29 % issue utilisation, profiles/r2bwd2_kernels.md).
template <int NG>
struct B2Dim {
    static_assert(NG == 2 || NG == 4, "2 or 4 unit groups");
    static constexpr int NPW = 4 * NG;              // point warps
    static constexpr int NPT = 128 * NG;            // point threads
    static constexpr int NT = NPT + 128;            // + the MMA warpgroup (its first warp issues, the others only help at the end)
    static constexpr int UPT = 32 / NG;             // units per thread
    static constexpr int NCH = UPT / 8;             // chunks of 8 units
    static constexpr int GCOLS = 96 / NG;           // G columns a thread accumulates for its lane
};
constexpr uint32_t COL_G = 416;           // G_g, G_tg, G_t: 32 columns each

template <class CF, int NG = 2>
struct Bwd2Cfg {
    using D = B2Dim<NG>;
    static constexpr int C = CF::C, NS = CF::NS, NA2 = CF::NA2, NA1 = CF::NA1;
    static constexpr int CA1 = 1 + NS + (NA2 > 0 ? 1 : 0);      // A operands of MMA 1: t, g kappa_s, t g
    static constexpr int NB1 = 1 + NA2;                         // B variants of MMA 1
    static constexpr int NB2 = 1 + NS + NA2;                    // B variants of MMA 3
    static constexpr uint32_t COL_A1LO = CA1 * 32, COL_A3LO = C * 32;
    static constexpr uint32_t COL_D1 = 2 * CA1 * 32, COL_D3 = 2 * C * 32;
    static_assert(COL_D1 + C * 32 <= COL_G && COL_D3 + 96 <= COL_G, "tensor memory budget");
    static constexpr int OFF_B1 = (CF::SM_PARAMS + 31) & ~31;   // [NB1][hi, lo][H*H]   B1_v[n = j][k] = W1[j][k] sc_v[k]
    static constexpr int OFF_B2 = OFF_B1 + NB1 * 2 * H * H;     // [NB2][hi, lo][H*H]   B2_v[n = k][j] = W1[j][k] sc_v[k]
    static constexpr int OFF_RING = OFF_B2 + NB2 * 2 * H * H;
    static constexpr int NIMG = 16;
    static constexpr int SLOT = NIMG * GK_IMG;                  // floats per ring slot (one quarter tile)
    static constexpr int FLOATS = OFF_RING + 2 * SLOT;
    // ring images (32 rows x 32 points each); the stacks of one MMA are contiguous
    static constexpr int img_ag(int part, int s) { return part * 2 + s; }        // first-order abar2 of slot s
    static constexpr int img_atg(int part, int s) { return 4 + part * 2 + s; }   // second-order abar2 of slot s
    static constexpr int img_at(int part) { return 8 + part; }                   // value abar2 (the M = 128 MMA also reads 10, 11)
    static constexpr int img_phi(int f, int part) { return 10 + 2 * f + part; }  // f: 0 t, 1 g, 2 t g
    // end-of-item scratch (floats, over the ring)
#ifndef FBP_B2_GDS
#define FBP_B2_GDS 100
#endif
    static constexpr int GDS = FBP_B2_GDS;                      // row stride of RED_GD: 100 floats = conflict-free 16-byte stores
    static constexpr int RED_GD = 0;                            // [128 lanes][GDS], 96 used
    static constexpr int RED_L0 = SLOT;                         // [point warps][chunks][32]
    static constexpr int RED_WL = RED_L0 + D::NPW * D::NCH * 32;    // [UPT][point threads]
    static constexpr int RED_B1 = RED_WL + D::UPT * D::NPT;     // [UPT][point threads]
    static constexpr int RED_BL = RED_B1 + D::UPT * D::NPT;     // [point warps]
    static constexpr int RED_KB = RED_BL + D::NPW;              // [max(NS,1)][8 partial sums over j][32]
    static_assert(128 * GDS <= SLOT && RED_KB + (NS > 0 ? NS : 1) * 8 * 32 <= 2 * SLOT, "reduction scratch must fit the ring");
};

// End of a work item: the partial sums the point threads left in shared memory (RED_*) -> this item's row of gpart.  A
// separate function on purpose: the instruction schedule of the tile loop of tc_backward_kernel2 is sensitive to any code
// compiled with it (the same loop ran at 14.1 k to 15.7 k cycles per tile across builds that only differed here,
// profiles/r2h_item_cost.md); w1g = W1[j][k] of the subdomain in global memory (L2-resident since the prologue).
__device__ __forceinline__ int sel3i(int i, int a, int b, int c) { return i == 0 ? a : (i == 1 ? b : c); }
#ifdef FBP_B2_FINISH_CALL
#define B2_FINISH_INLINE __noinline__
#else
#define B2_FINISH_INLINE __forceinline__
#endif
template <class CF, int NG>
__device__ B2_FINISH_INLINE void b2_finish(float* sm, const float* __restrict__ w1g, float* __restrict__ gp, int xd, int3 axis, float3 isd) {
    using L = Bwd2Cfg<CF, NG>;
    using DM = B2Dim<NG>;
    constexpr int B2_NPT = DM::NPT, B2_NT = DM::NT, UPT = DM::UPT, NCH = DM::NCH;
    constexpr int NS = CF::NS, NA2 = CF::NA2, HH = H * H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* red = sm + L::OFF_RING;
    const float* gd = red + L::RED_GD;
    constexpr int GDS = L::GDS;
    // G_g[s][j][k], G_tg[s][j][k], G_t[j][k]: the hi and lo rows of the stacks added
    auto Gg = [&](int s, int j, int k) { return gd[(32 * s + j) * GDS + k] + gd[(64 + 32 * s + j) * GDS + k]; };
    auto Gtg = [&](int s, int j, int k) { return gd[(32 * s + j) * GDS + 32 + k] + gd[(64 + 32 * s + j) * GDS + 32 + k]; };
    auto Gt = [&](int j, int k) { return gd[j * GDS + 64 + k] + gd[(32 + j) * GDS + 64 + k]; };
    // derivative path of the first layer: kappa_bar_s[k] = sum_j W1[j][k] (G_g[s][j][k] - 4 kappa_s[k] G_tg[s][j][k]);
    // a warp = (slot s, 4 rows j), lane = k; the 8 partial sums are added by the reader
    for (int i = tid; i < NS * 8 * H; i += B2_NT) {
        const int s = i >> 8, part = (i >> 5) & 7, k = i & 31;
        const float kap = sm[CF::SM_W0D + s * H + k];
        float v = 0.0f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * part + jj;
            float gsum = Gg(s, j, k);
            if (s < NA2) gsum = fmaf(-4.0f * kap, Gtg(s, j, k), gsum);
            v = fmaf(w1g[j * H + k], gsum, v);
        }
        red[L::RED_KB + i] = v;
    }
    __syncthreads();

    // first layer, value path: (unit k, quantity t) sits in lane (k % 8) * 4 + t of chunk (k % UPT) / 8 of the four point warps
    // of unit group k / UPT
    auto l0 = [&](int k, int t) {
        const int gg_ = k / UPT, ch = (k % UPT) >> 3, ln = (k & 7) * 4 + t;
        float v = 0.0f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) v += red[L::RED_L0 + ((gg_ * 4 + qq) * NCH + ch) * 32 + ln];
        return v;
    };
    for (int i = tid; i < H * xd; i += B2_NT) {
        const int j = i / xd, d = i - j * xd;
        float v = l0(j, 1 + d);
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (sel3i(s, axis.x, axis.y, axis.z) == d) {
                float kb = 0.0f;
#pragma unroll
                for (int part = 0; part < 8; ++part) kb += red[L::RED_KB + (s * 8 + part) * H + j];
                v = fmaf(sel3(d, isd.x, isd.y, isd.z), kb, v);
            }
        gp[i] = v;
    }
    for (int i = tid; i < H; i += B2_NT) gp[H * xd + i] = l0(i, 0);
    int off = H * xd + H;
    for (int i = tid; i < HH; i += B2_NT) {
        const int j = i >> 5, k = i & 31;
        float v = Gt(j, k);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float kap = sm[CF::SM_W0D + s * H + k];
            v = fmaf(kap, Gg(s, j, k), v);
            if (s < NA2) v = fmaf(-2.0f * kap * kap, Gtg(s, j, k), v);
        }
        gp[off + i] = v;
    }
    off += HH;
    for (int o = warp; o < 2 * H; o += B2_NT / 32) {   // hidden bias and output weights: sums over the 128 point threads of the unit's group
        const int j = o & 31;
        const float* src = red + (o < H ? L::RED_B1 : L::RED_WL) + (j % UPT) * B2_NPT + (j / UPT) * 128;
        const float v = fbp_warp_sum((src[lane] + src[lane + 32]) + (src[lane + 64] + src[lane + 96]));
        if (lane == 0) gp[off + o] = v;
    }
    off += 2 * H;
    if (tid == 0) {
        float v = 0.0f;
        for (int w = 0; w < 4; ++w) v += red[L::RED_BL + w];    // warps 0-3 are unit half 0 (the only ones that count it)
        gp[off] = v;
    }
}

template <class CF, int NG>
__global__ void __launch_bounds__(B2Dim<NG>::NT, 1) tc_backward_kernel2(FastArgs a) {
    static_assert(CF::H == 32 && CF::NHID == 2, "tensor family: H = 32, two hidden layers");
    using L = Bwd2Cfg<CF, NG>;
    using DM = B2Dim<NG>;
    constexpr int B2_NPW = DM::NPW, B2_NPT = DM::NPT, B2_NT = DM::NT, UPT = DM::UPT, NCH = DM::NCH, GCOLS = DM::GCOLS;
    constexpr int C = CF::C, NS = CF::NS, NA2 = CF::NA2, NA1 = CF::NA1, CA1 = L::CA1;
    constexpr int HH = H * H;

    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar_a1, bar_a3, bar_m1, bar_m3, bar_full[2], bar_free[2], bar_gtile;
    __shared__ uint32_t tmem_slot;
    __shared__ int hdr_i[4];                  // first pair, pair count, subdomain index, item
    __shared__ float hdr_f[8];                // mu[3], 1/sd[3], window flag, output scale

    const int tid = threadIdx.x, warp = warp_uniform(), lane = tid & 31;
    blk_stamp(1);
    // a.dbg (timing experiments only, results are then wrong): 1 = no weight-gradient staging / G MMAs, 4 = no ring stores
    // (the G MMAs run on stale images), 8 = no G MMA issue (the stores happen), 32 = no butterfly
    const int dbg = a.dbg;

    const int xd = a.xd;
    {   // prologue: nothing computed here stays in registers (the header in shared memory is re-read per role below, so that
        // the register allocation of the tile loop does not depend on this block)
        const ItemRec it = tc_item(a);
        if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);          // its latency hides under the parameter loads
        const float* ss = a.sub_static + (int64_t)it.im * (2 * xd + 3);
        float isd_[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) isd_[d] = d < xd ? 1.0f / ((ss[xd + d] - ss[d]) * 0.5f) : 0.0f;
        if (tid == 32) {
            hdr_i[0] = it.first; hdr_i[1] = it.count; hdr_i[2] = it.im; hdr_i[3] = it.item;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                hdr_f[d] = d < xd ? (ss[xd + d] + ss[d]) * 0.5f : 0.0f;
                hdr_f[3 + d] = isd_[d];
            }
            hdr_f[6] = ss[2 * xd];
            hdr_f[7] = ss[2 * xd + 2];
            mbar_init(&bar_a1, B2_NPT);
            mbar_init(&bar_a3, B2_NPT);
            mbar_init(&bar_gtile, 1);
            mbar_init(&bar_m1, 1);
            mbar_init(&bar_m3, 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bar_full[i], 32 * NG);   // the NG warps of a quarter tile
                mbar_init(&bar_free[i], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        const float* prow = a.params + (int64_t)it.im * a.P;
        tc_stage_small<CF>(sm, prow, xd, isd_, a.axis, tid, B2_NT);
        tc_stage_b<CF, true>(sm + L::OFF_B1, sm + L::OFF_B2, prow, xd, isd_, a.axis, tid, B2_NT);
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
    const float flag = hdr_f[6], un_sd = hdr_f[7];
    const int ntiles = (count + TP - 1) / TP;
    // registers move from the MMA warpgroup to the point warpgroups (what the former frees is what the latter take)
    if (NG == 2) {
        if (warp < B2_NPW) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else {
        if (warp < B2_NPW) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    }

    float* red = sm + L::OFF_RING;            // end-of-item scratch (the ring is free once the last G MMA has completed)

    if (warp < B2_NPW) {
        // =============================================================================================
        // point warps
        // =============================================================================================
        blk_stamp(2);
        const int g = warp >> 2, q = warp & 3;
        const int r = tid & 127;                    // point row of the tile = TMEM lane
        const int j0 = UPT * g;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int slot = q & 1;
        float* ring = sm + L::OFF_RING + slot * L::SLOT + (lane >> 2) * (int)(GK_LBO / 4) + (lane & 3) + NCH * g * (int)(GK_SBO / 4);
        float wlacc[UPT], b1acc[UPT], l0acc[NCH], blacc = 0.0f;      // per-thread partial sums over the item
#pragma unroll
        for (int u = 0; u < UPT; ++u) { wlacc[u] = 0.0f; b1acc[u] = 0.0f; }
#pragma unroll
        for (int u = 0; u < NCH; ++u) l0acc[u] = 0.0f;
        float gacc[GCOLS];                          // this thread's columns of its G lane, summed over the tiles
#pragma unroll
        for (int u = 0; u < GCOLS; ++u) gacc[u] = 0.0f;
        auto gather_g = [&](uint32_t parity) {      // add the G block of a finished tile (the MMA warp has committed it)
            mbar_wait_or_trap(&bar_gtile, parity);
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < GCOLS / 8; ++ch) {
                uint32_t v[8];
                tmem_ld8(tbase + lane_base + COL_G + GCOLS * g + 8 * ch, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 8; ++e) gacc[8 * ch + e] += __uint_as_float(v[e]);
            }
        };

        int pf_pt = 0, pf_row = 0;
        float pf_x[3] = {0.0f, 0.0f, 0.0f}, pf_g[C];
#pragma unroll
        for (int c = 0; c < C; ++c) pf_g[c] = 0.0f;
        auto load_idx = [&](int t0n) {
            if (t0n < count) {
                const int cn = min(TP, count - t0n);
                const int pi = first + t0n + (r < cn ? r : 0);
                pf_pt = a.spair_point[pi];
                pf_row = a.spair_row[pi];
            }
        };
        auto load_val = [&](int t0n) {
            if (t0n < count) {
#pragma unroll
                for (int d = 0; d < 3; ++d) pf_x[d] = d < xd ? a.x[(int64_t)pf_pt * xd + d] : 0.0f;
                const float* gr = a.grow + (int64_t)pf_row * C;
#pragma unroll
                for (int c = 0; c < C; ++c) pf_g[c] = gr[a.ext[c]];
            }
        };
        load_idx(0);
        load_val(0);

        // Phase trace (builds with -DFBP_B2_TRACE only: the three extra live registers cost the product kernel 14 %; run with
        // FBP_TC_DEBUG=64, tests/tools/bwd_phase_trace.py): thread 0 of block 0 records the cycle counter at the phase boundaries
        // of every tile into the buffer passed as activation cache: [tile][8] = start, a1 arrive, m1 done, a3 arrive, ring
        // done, m3 done, E2 done, end; [512 + tile * 4 ...] = inside the ring store
#ifdef FBP_B2_TRACE
        const bool rec = (dbg & 64) && a.cache != nullptr && blockIdx.x == 0 && tid == ((dbg >> 8) & 1023);   // FBP_TC_DEBUG = 64 + 256 * thread
        const long long rec0 = clock64();
        auto stamp = [&](int t, int k) { if (rec) a.cache[t * 8 + k] = (float)(clock64() - rec0); };
        auto stamp_ring = [&](int t, int k) { if (rec) a.cache[512 + t * 4 + k] = (float)(clock64() - rec0); };
#else
        auto stamp = [&](int, int) {};
        auto stamp_ring = [&](int, int) {};
#endif
        // S0 of tile t + 1 runs at the end of tile t, BEFORE the late ring store of the quarters 2, 3: those warps otherwise idle
        // until the weight-gradient MMAs of the quarters 0, 1 have freed the slot, and S0 + L of the next tile is what the tensor
        // core waits for after that store (profiles/r2g_bwd_phase_trace.md)
        float z[3], rb[C];
        auto do_s0 = [&](int tn) {
                // ---- S0: coordinates, window jets, cotangent of the output-layer jets ----------------------
#pragma unroll
                for (int d = 0; d < 3; ++d) z[d] = d < xd ? (pf_x[d] - mu[d]) * isd[d] : 0.0f;
                {
                    float w, w1[NS > 0 ? NS : 1], w2[NA2 > 0 ? NA2 : 1];
                    fast_window<CF>(z, isd, xd, flag, a.axis, w, w1, w2);
                    const bool valid = r < min(TP, count - tn * TP);
                    float G[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) G[c] = valid ? pf_g[c] : 0.0f;
                    float ub0 = G[0] * w;
#pragma unroll
                    for (int s = 0; s < NA2; ++s) {
                        const float G1 = G[1 + 2 * s], G2 = G[2 + 2 * s];
                        ub0 += G1 * w1[s] + G2 * w2[s];
                        rb[1 + 2 * s] = un_sd * (G1 * w + 2.0f * G2 * w1[s]);
                        rb[2 + 2 * s] = un_sd * (G2 * w);
                    }
#pragma unroll
                    for (int s = 0; s < NA1; ++s) {
                        const int c = 1 + 2 * NA2 + s;
                        ub0 += G[c] * w1[NA2 + s];
                        rb[c] = un_sd * (G[c] * w);
                    }
                    rb[0] = un_sd * ub0;
                    if (g == 0) blacc += rb[0];
                }
            load_idx((tn + 1) * TP);
        };
        if (ntiles > 0) do_s0(0);
        for (int t = 0; t < ntiles; ++t) {
            const int t0 = t * TP;
            const int cnt = min(TP, count - t0);
            const uint32_t par = (uint32_t)(t & 1);
            stamp(t, 0);


            // ---- L: t, g of this thread's 16 units; A operands of MMA 1 (t, g kappa_s, t g) -> tensor memory ----
            float tt[UPT], gg[UPT];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int jb = j0 + 8 * ch;
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
                        tt[8 * ch + e] = tv;
                        gg[8 * ch + e] = gv;
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
                    tmem_st8(tbase + lane_base + L::COL_A1LO + i * 32 + jb, lo);
                }
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&bar_a1);                              // A of MMA 1 written; D3 of the previous tile read
            stamp(t, 1);
            load_val(t0 + TP);
            // the previous tile's G block (complete before MMA 1 of this tile even starts: the tensor pipe runs in issue
            // order) is added into registers while MMA 1 runs; this tile's G is issued after MMA 3, i.e. after every
            // thread has passed this point
            if (t > 0 && !(dbg & 1)) gather_g((uint32_t)((t - 1) & 1));

            // ---- E1: a2 -> h2, output-layer gradient partials, tanh transpose -> abar2 -> A operands of MMA 3 ---------
            mbar_wait_or_trap(&bar_m1, par);                   // all of MMA 1: a2 is final and its A operands may be overwritten
            tc_fence_after();
            stamp(t, 2);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int jb = j0 + 8 * ch;
                uint32_t v[C][8];
#pragma unroll
                for (int c = 0; c < C; ++c) tmem_ld8(tbase + lane_base + L::COL_D1 + c * 32 + jb, v[c]);
                tmem_wait_ld();
                float ab[8][C];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float h2[C], hb[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) h2[c] = __uint_as_float(v[c][e]);
                    h2[0] += sm[CF::SM_B1 + jb + e];
                    fast_tanh_jets<CF>(h2);
                    const float wl = sm[CF::SM_WL + jb + e];
                    float dsum = 0.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        dsum = fmaf(rb[c], h2[c], dsum);
                        hb[c] = wl * rb[c];
                    }
                    wlacc[8 * ch + e] += dsum;
                    fast_tanh_jets_bwd<CF>(h2, hb, ab[e]);
                    b1acc[8 * ch + e] += ab[e][0];
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        hi[e] = __float_as_uint(ab[e][c]);
                        lo[e] = __float_as_uint(tf32_lo(ab[e][c]));
                    }
                    tmem_st8(tbase + lane_base + c * 32 + jb, hi);
                    tmem_st8(tbase + lane_base + L::COL_A3LO + c * 32 + jb, lo);
                }
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&bar_a3);
            stamp(t, 3);

            // hi/lo images of abar2 and (t, g, t g) of this thread's point and units -> ring slot of the quarter tile.
            // abar2 is read back from the hi columns of MMA 3's A operands (they hold the unsplit values) instead of
            // living in 80 registers across E2.
            auto ring_store = [&]() {
                const uint32_t use = 2u * (uint32_t)t + (uint32_t)(q >> 1);
                if (use > 0) mbar_wait_or_trap(&bar_free[slot], (use - 1) & 1);
                stamp_ring(t, 0);
                auto put = [&](int img, int e, float x) {
                    ring[img * GK_IMG + (e >> 3) * (int)(GK_SBO / 4) + (e & 7) * 4] = x;
                };
                if (!(dbg & 4))
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    uint32_t v[C][8];
#pragma unroll
                    for (int c = 0; c < C; ++c) tmem_ld8(tbase + lane_base + c * 32 + j0 + 8 * ch, v[c]);
                    tmem_wait_ld();
                    stamp_ring(t, 1);
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int e = 8 * ch + e8;
                        const float x0 = __uint_as_float(v[0][e8]);
                        put(L::img_at(0), e, x0);
                        put(L::img_at(1), e, tf32_lo(x0));
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            const float x1 = __uint_as_float(v[s < NA2 ? 1 + 2 * s : 1 + 2 * NA2 + (s - NA2)][e8]);
                            put(L::img_ag(0, s), e, x1);
                            put(L::img_ag(1, s), e, tf32_lo(x1));
                            if (s < NA2) {
                                const float x2 = __uint_as_float(v[2 + 2 * s][e8]);
                                put(L::img_atg(0, s), e, x2);
                                put(L::img_atg(1, s), e, tf32_lo(x2));
                            }
                        }
                        const float tv = tt[e], gv = gg[e], tg = tv * gv;
                        put(L::img_phi(0, 0), e, tv);
                        put(L::img_phi(0, 1), e, tf32_lo(tv));
                        if (NS > 0) {
                            put(L::img_phi(1, 0), e, gv);
                            put(L::img_phi(1, 1), e, tf32_lo(gv));
                        }
                        if (NA2 > 0) {
                            put(L::img_phi(2, 0), e, tg);
                            put(L::img_phi(2, 1), e, tf32_lo(tg));
                        }
                    }
                }
                stamp_ring(t, 2);
                fence_proxy_async();                           // generic-proxy stores -> visible to the tensor core's reads
                mbar_arrive(&bar_full[slot]);
                stamp_ring(t, 3);
            };
            if (q < 2 && !(dbg & 1)) ring_store();             // quarters 0, 1 own the slots first; 2, 3 after their E2
            stamp(t, 4);

            // ---- E2: abar0 = g D_t - 2 t g D_g + g (1 - 3 t^2) D_tg  -> first-layer gradient partials ------------------
            mbar_wait_or_trap(&bar_m3, par);                   // all of MMA 3: its A operands may be rewritten by the next tile
            tc_fence_after();
            stamp(t, 5);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int kb = j0 + 8 * ch;
                uint32_t vt[8], vg[8], vtg[8];
                tmem_ld8(tbase + lane_base + L::COL_D3 + kb, vt);
                if (NS > 0) tmem_ld8(tbase + lane_base + L::COL_D3 + 32 + kb, vg);
                if (NA2 > 0) tmem_ld8(tbase + lane_base + L::COL_D3 + 64 + kb, vtg);
                tmem_wait_ld();
                float qv[32];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float tv = tt[8 * ch + e], gv = gg[8 * ch + e];
                    float acc = __uint_as_float(vt[e]);
                    if (NS > 0) acc = fmaf(-2.0f * tv, __uint_as_float(vg[e]), acc);
                    if (NA2 > 0) acc = fmaf(fmaf(-3.0f * tv, tv, 1.0f), __uint_as_float(vtg[e]), acc);
                    const float ab0 = gv * acc;
                    qv[4 * e + 0] = ab0;
                    qv[4 * e + 1] = ab0 * z[0];
                    qv[4 * e + 2] = ab0 * z[1];
                    qv[4 * e + 3] = ab0 * z[2];
                }
                l0acc[ch] += (dbg & 32) ? qv[0] + qv[9] + qv[18] + qv[27] : warp_transpose_reduce32(qv, lane);
            }
            stamp(t, 6);
            if (t + 1 < ntiles) do_s0(t + 1);
            if (q >= 2 && !(dbg & 1)) ring_store();
            stamp(t, 7);
        }
        BLK_STAMP3;
        // ---- end of item: every partial to shared memory.  The last G MMA has completed (so has every ring store). ------
        if (ntiles > 0 && !(dbg & 1)) gather_g((uint32_t)((ntiles - 1) & 1));     // also: every ring read has completed
#pragma unroll
        for (int u = 0; u < GCOLS; ++u) red[L::RED_GD + (q * 32 + lane) * L::GDS + GCOLS * g + u] = gacc[u];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) red[L::RED_L0 + (warp * NCH + ch) * 32 + lane] = l0acc[ch];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            red[L::RED_WL + u * B2_NPT + tid] = wlacc[u];
            red[L::RED_B1 + u * B2_NPT + tid] = b1acc[u];
        }
        const float bv = fbp_warp_sum(blacc);
        if (lane == 0) red[L::RED_BL + warp] = bv;
    } else if (warp == B2_NPW) {
        // =============================================================================================
        // MMA warp
        // =============================================================================================
        // three base descriptors; every operand descriptor of the tile loop is one of them plus a compile-time constant
        const uint64_t b1base = make_smem_desc(smem_u32(sm + L::OFF_B1), B_LBO, B_SBO);
        const uint64_t b2base = make_smem_desc(smem_u32(sm + L::OFF_B2), B_LBO, B_SBO);
        const uint64_t gbase = make_smem_desc(smem_u32(sm + L::OFF_RING), GK_LBO, GK_SBO);
        auto b1d = [&](int v, int part) { return desc_advance(b1base, (uint32_t)((2 * v + part) * HH * 4)); };
        auto b2d = [&](int v, int part) { return desc_advance(b2base, (uint32_t)((2 * v + part) * HH * 4)); };
        constexpr uint32_t idesc_g = make_idesc_tf32(128, 32);
        for (int t = 0; t < ntiles; ++t) {
            const uint32_t par = (uint32_t)(t & 1);
            // ---- MMA 1: a2[c] = A1[ia(c)] * B1[vb(c)]^T, unit half 0 first ------------------------------------
            mbar_wait_or_trap(&bar_a1, par);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    int ia = 0, vb = 0;
                    if (c > 0 && c <= 2 * NA2) {
                        const int s = (c - 1) >> 1;
                        if ((c - 1) & 1) { ia = CA1 - 1; vb = 1 + s; }      // second order: (t g) against W1 diag(-2 kappa_s^2)
                        else ia = 1 + s;
                    } else if (c > 0) ia = 1 + NA2 + (c - 1 - 2 * NA2);
                    issue_gemm_acc<32>(tbase + L::COL_D1 + c * 32, tbase + ia * 32, tbase + L::COL_A1LO + ia * 32, b1d(vb, 0), b1d(vb, 1), true);
                }
                mma_commit(&bar_m1);
            }
            __syncwarp();
            // ---- MMA 3: D_t, D_g, D_tg ---------------------------------------------------------------------------
            mbar_wait_or_trap(&bar_a3, par);
            tc_fence_after();
            if (elect_one()) {
                issue_gemm_acc<32>(tbase + L::COL_D3, tbase, tbase + L::COL_A3LO, b2d(0, 0), b2d(0, 1), true);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int c1 = s < NA2 ? 1 + 2 * s : 1 + 2 * NA2 + (s - NA2);
                    issue_gemm_acc<32>(tbase + L::COL_D3 + 32, tbase + c1 * 32, tbase + L::COL_A3LO + c1 * 32, b2d(1 + s, 0), b2d(1 + s, 1), s == 0);
                }
#pragma unroll
                for (int s = 0; s < NA2; ++s) {
                    const int c2 = 2 + 2 * s;
                    issue_gemm_acc<32>(tbase + L::COL_D3 + 64, tbase + c2 * 32, tbase + L::COL_A3LO + c2 * 32, b2d(1 + NS + s, 0),
                                       b2d(1 + NS + s, 1), s == 0);
                }
                mma_commit(&bar_m3);
            }
            __syncwarp();
            // ---- G: weight-gradient products of the four quarter tiles, accumulated over the whole work item --------
            if (!(dbg & 1))
#pragma unroll 1
            for (int qi = 0; qi < 4; ++qi) {
                const int slot = qi & 1;
                mbar_wait_or_trap(&bar_full[slot], (uint32_t)(qi >> 1));
                if (elect_one()) {
                    const uint64_t gslot = slot ? desc_advance(gbase, (uint32_t)(L::SLOT * 4)) : gbase;
                    const bool fresh = (qi == 0);                 // every tile starts fresh accumulators
                    auto block = [&](int a_img, int b_img, uint32_t dcol) {
#pragma unroll
                        for (int part = 0; part < 2; ++part) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t adv = (uint32_t)(ks * 2 * GK_LBO);
                                mma_tf32_ss(tbase + dcol, desc_advance(gslot, (uint32_t)(a_img * GK_IMG * 4) + adv),
                                            desc_advance(gslot, (uint32_t)((b_img + part) * GK_IMG * 4) + adv), idesc_g,
                                            (fresh && part == 0 && ks == 0) ? 0u : 1u);
                            }
                        }
                    };
                    if (!(dbg & 8)) {
                        if (NS > 0) block(L::img_ag(0, 0), L::img_phi(1, 0), COL_G);
                        if (NA2 > 0) block(L::img_atg(0, 0), L::img_phi(2, 0), COL_G + 32);
                        block(L::img_at(0), L::img_phi(0, 0), COL_G + 64);
                    }
                    mma_commit(&bar_free[slot]);
                }
                __syncwarp();
            }
            if (elect_one()) mma_commit(&bar_gtile);           // this tile's G is complete when this arrives
            __syncwarp();
        }

    }

    __syncthreads();
    blk_stamp(4);

    b2_finish<CF, NG>(sm, a.params + (int64_t)hdr_i[2] * a.P + H * xd + H, a.gpart + (int64_t)hdr_i[3] * a.P, xd,
                      make_int3(a.axis[0], a.axis[1], a.axis[2]), make_float3(hdr_f[3], hdr_f[4], hdr_f[5]));
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TMEM_COLS);
    blk_stamp(5);
    blk_stamp_entry(ntiles);
}

}  // namespace fbptc
