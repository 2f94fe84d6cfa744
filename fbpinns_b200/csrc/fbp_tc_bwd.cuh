// tcgen05 ("tensor") family, reverse kernel: warp-specialised, one CTA (384 threads) per work item.
//
//   warps 0-7  "point warps"  thread (g, p): point row p of the 128-pair tile (= TMEM lane), hidden units [16 g, 16 g+16)
//   warps 8-11 "gradient warps": the weight gradient of the hidden matrix on the FFMA2 pipe
//
// Per tile the point warps run the row-per-point part of the reverse pass with both of its GEMMs on the tensor core
// (3xTF32, A operands written straight into tensor memory by the thread that computed the row, fbp_tc.cuh):
//   S0  window jets, cotangent of the output-layer jets (as fast_backward_kernel)
//   L   layer 0 + tanh jets h1 -> A (hi, lo)                                   MMA 1:  a2 = h1 W1^T
//   E1  a2 -> h2 (tanh jets), output-layer weight gradient partials, tanh transpose -> abar2
//       abar2 -> A (hi, lo) and -> shared memory image As[buf][j][c][p]        MMA 3:  hbar1 = abar2 W1
//   E2  hbar1, recomputed h1 -> abar1 (tanh transpose) -> first-layer gradient partials, reduced over the warp's 32
//       points with a transposing butterfly (31 shuffles per 32 values) into lane-distributed accumulators
// The weight gradient  Wbar1[j][k] += sum_{c,p} abar2_c[p][j] h1_c[p][k]  contracts over POINTS, i.e. over TMEM lanes:
// as a tensor-core GEMM both operands would have to come from shared memory and their hi/lo images are 320 KB per tile,
// more than an SM has.  It therefore stays on the FP32 pipe, but on its own warps: the gradient warps wait for As[buf]
// (double buffered, mbarrier full/empty pairs), recompute h1 for 8 points at a time into a warp-private 5.6 KB buffer
// (layer 0 has K = xd: cheaper than staging 80 KB) and accumulate 8x4 (j,k) register tiles of float2 (FFMA2 over point
// pairs) exactly like fast_backward_kernel's G phase.  They run one tile behind the point warps, so the FFMA2 stream
// overlaps the latency-bound TMEM / MMA / epilogue chain of the next tile instead of alternating with it.
//
// No activation cache: the tensor core recomputes a2 faster than HBM delivers it.
// Status: gradients validated on a B200 against the tiled kernel (3.6e-7, profiles/r1f_tc_bringup.md) on a small case
// through the C ABI; not yet timed, hence opt-in (fbp_plan_set_kernel mode 4 / kernel="tensor-full").
#pragma once
#include "fbp_tc.cuh"

namespace fbptc {

constexpr int BWD_NT = 384;
constexpr int NPW = 8, NGW = 4;           // point warps, gradient warps
constexpr int NPT = NPW * 32;             // 256 point threads
constexpr int HS_ROW = 44;                // floats per unit row of the warp-private h1 chunk [32][C][8] (C*8 = 40, +4: the 8
                                          // rows a float4 request touches fall in 8 disjoint bank quads)

template <class CF>
struct BwdSmem {
    static constexpr int C = CF::C;
    static constexpr int RS = C * TP + 4;                       // row stride of As (as in fast_backward_kernel)
    static constexpr int OFF_B1HI = (CF::SM_PARAMS + 31) & ~31; // B1[n = j][k] = W1[j][k]   (a2 = h1 W1^T)
    static constexpr int OFF_B1LO = OFF_B1HI + H * H;
    static constexpr int OFF_B2HI = OFF_B1LO + H * H;           // B2[n = k][j] = W1[j][k]   (hbar1 = abar2 W1)
    static constexpr int OFF_B2LO = OFF_B2HI + H * H;
    static constexpr int OFF_ZS = OFF_B2LO + H * H;             // [2][3][TP]  normalised coordinates, by tile parity
    static constexpr int OFF_AS = OFF_ZS + 2 * 3 * TP;          // [2][H][RS]  abar2 images, by tile parity
    static constexpr int OFF_HS = OFF_AS + 2 * H * RS;          // [NGW][H][HS_ROW]
    static constexpr int FLOATS = OFF_HS + NGW * H * HS_ROW;
    // end-of-item reduction scratch over the As and Hs regions (both are dead by then)
    static constexpr int RED_G = 0;                             // [NGW][JJ*KK + JJ][32]
    static constexpr int RED_L0 = RED_G + NGW * (8 * 4 + 8) * 32;   // [NPW][2][2][32]
    static constexpr int RED_WL = RED_L0 + NPW * 4 * 32;        // [16][NPT]
    static constexpr int RED_BL = RED_WL + 16 * NPT;            // [NPW]
    static_assert(OFF_AS + RED_BL + NPW <= FLOATS, "reduction scratch must fit the As + Hs regions");
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One stage of the transposing butterfly: N values per lane -> N/2, exchanging with lane ^ S.
template <int N, int S>
__device__ __forceinline__ void tr_stage(float (&v)[32], int lane) {
    const bool up = (lane & S) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float mine = up ? v[i + N / 2] : v[i];
        const float send = up ? v[i] : v[i + N / 2];
        v[i] = mine + __shfl_xor_sync(0xffffffffu, send, S);
    }
}
// 32 values per lane in, lane l returns the sum over the warp of value l (31 shuffles instead of 32 * 5).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
    tr_stage<32, 16>(v, lane);
    tr_stage<16, 8>(v, lane);
    tr_stage<8, 4>(v, lane);
    tr_stage<4, 2>(v, lane);
    tr_stage<2, 1>(v, lane);
    return v[0];
}

template <class CF>
__global__ void __launch_bounds__(BWD_NT, 1) tc_backward_kernel(FastArgs a) {
    static_assert(CF::H == 32 && CF::NHID == 2, "tensor family: H = 32, two hidden layers");
    static_assert(3 * CF::C * 32 <= (int)TMEM_COLS, "A hi, A lo and D must fit the 512 TMEM columns");
    constexpr int C = CF::C, NS = CF::NS, NA2 = CF::NA2, NA1 = CF::NA1;
    constexpr uint32_t COL_AHI = 0, COL_ALO = C * 32, COL_D = 2 * C * 32;
    constexpr int NQ0 = 4 + NS;                 // per unit: bias, 3 coordinate sums, NS slot sums
    static_assert(NQ0 <= 8, "layer-0 partials are padded to 8 per unit");
    using L = BwdSmem<CF>;
    constexpr int RS = L::RS;
    constexpr int JJ = 8, KK = 4;               // per-lane (j,k) register tile of the weight gradient: j = jt + 4 jj, k = kt + 8 kk

    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar_m1[2], bar_m3[2], bar_full[2], bar_empty[2];
    __shared__ uint32_t tmem_slot;
    float* b1hi = sm + L::OFF_B1HI;
    float* b1lo = sm + L::OFF_B1LO;
    float* b2hi = sm + L::OFF_B2HI;
    float* b2lo = sm + L::OFF_B2LO;
    float* zs = sm + L::OFF_ZS;
    float* As = sm + L::OFF_AS;

    const int tid = threadIdx.x, warp = warp_uniform(), lane = tid & 31;
    const bool is_p = warp < NPW;
    // a.dbg (timing experiments only, results are then wrong): 2 = no MMA issue / waits, 16 = gradient warps skip their
    // arithmetic, 32 = no butterfly reduction of the first-layer partials
    const int dbg = a.dbg;

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
    const float* prow = a.params + (int64_t)im * a.P;
    fast_load_params<CF, BWD_NT>(sm, prow, xd, isd, a.axis, false);
    stage_b(b1hi, b1lo, prow + H * xd + H, H, 1, tid, BWD_NT);
    stage_b(b2hi, b2lo, prow + H * xd + H, 1, H, tid, BWD_NT);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_m1[i], 1);
            mbar_init(&bar_m3[i], 1);
            mbar_init(&bar_full[i], NPT);           // every point thread arrives after its As / zs stores
            mbar_init(&bar_empty[i], NGW * 32);     // every gradient thread arrives after its last read
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const int ntiles = (count + TP - 1) / TP;

    if (is_p) {
        // =============================================================================================
        // point warps
        // =============================================================================================
        const int g = warp >> 2;                    // unit half
        const int r = tid & 127;                    // point row = TMEM lane
        const int j0 = 16 * g;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint64_t b1d_hi = make_smem_desc(smem_u32(b1hi), B_LBO, B_SBO), b1d_lo = make_smem_desc(smem_u32(b1lo), B_LBO, B_SBO);
        const uint64_t b2d_hi = make_smem_desc(smem_u32(b2hi), B_LBO, B_SBO), b2d_lo = make_smem_desc(smem_u32(b2lo), B_LBO, B_SBO);

        float wlacc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) wlacc[u] = 0.0f;
        float l0acc[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
        float blacc = 0.0f;

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
        // layer 0 + tanh jets of unit j at normalised point z
        auto layer0 = [&](int j, const float (&z)[3], float (&h)[C]) {
            h[0] = fmaf(sm[CF::SM_W0 + 2 * H + j], z[2], fmaf(sm[CF::SM_W0 + H + j], z[1], fmaf(sm[CF::SM_W0 + j], z[0], sm[CF::SM_B0 + j])));
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const float wv = sm[CF::SM_W0D + s * H + j];
                if (s < NA2) { h[1 + 2 * s] = wv; h[2 + 2 * s] = 0.0f; }
                else h[1 + 2 * NA2 + (s - NA2)] = wv;
            }
            fast_tanh_jets<CF>(h);
        };
        load_idx(0);
        load_val(0);

        for (int t = 0; t < ntiles; ++t) {
            const int t0 = t * TP;
            const int cnt = min(TP, count - t0);
            const int buf = t & 1;
            const uint32_t par = (uint32_t)(t & 1);            // phase parity of the per-tile MMA barriers
            const uint32_t use = (uint32_t)(t >> 1);           // how often this buffer has been used before
            float* Asb = As + buf * H * RS;
            float* zsb = zs + buf * 3 * TP;

            // ---- S0: coordinates, window jets, cotangent of the output-layer jets ----------------------
            float z[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) z[d] = d < xd ? (pf_x[d] - mu[d]) * isd[d] : 0.0f;
            float rb[C];
            {
                float w, w1[NS > 0 ? NS : 1], w2[NA2 > 0 ? NA2 : 1];
                fast_window<CF>(z, isd, xd, flag, a.axis, w, w1, w2);
                const bool valid = r < cnt;
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
            load_idx(t0 + TP);

            // ---- L: h1 of this thread's 16 units -> A (hi, lo) --------------------------------------------
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int jb = j0 + 8 * ch;
                float hv[8][C];
#pragma unroll
                for (int e = 0; e < 8; ++e) layer0(jb + e, z, hv[e]);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) tf32_split_fast(hv[e][c], hi[e], lo[e]);
                    tmem_st8(tbase + lane_base + COL_AHI + c * 32 + jb, hi);
                    tmem_st8(tbase + lane_base + COL_ALO + c * 32 + jb, lo);
                }
            }
            tmem_wait_st();
            tc_fence_before();
            named_bar_sync(1, NPT);                            // A complete; D of the previous tile fully read
            if (warp == 0 && !(dbg & 2)) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int nh = 0; nh < 2; ++nh) {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            issue_gemm_half(tbase + COL_D + c * 32 + nh * 16, tbase + COL_AHI + c * 32, tbase + COL_ALO + c * 32,
                                            b1d_hi, b1d_lo, nh);
                        mma_commit(&bar_m1[nh]);
                    }
                }
                __syncwarp();
            }
            load_val(t0 + TP);
            // the gradient warps must have finished the tile that used this buffer two tiles ago
            if (use > 0) mbar_wait_or_trap(&bar_empty[buf], (use - 1) & 1);
            if (g == 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) zsb[d * TP + r] = z[d];
            }

            // ---- E1: a2 -> h2, output-layer gradient partials, tanh transpose -> abar2 -> A and As ------------
            if (!(dbg & 2)) mbar_wait_or_trap(&bar_m1[g], par);
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int jb = j0 + 8 * ch;
                uint32_t v[C][8];
#pragma unroll
                for (int c = 0; c < C; ++c) tmem_ld8(tbase + lane_base + COL_D + c * 32 + jb, v[c]);
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
#pragma unroll
                    for (int c = 0; c < C; ++c) Asb[(jb + e) * RS + c * TP + r] = ab[e][c];
                }
                if (ch == 0 && !(dbg & 2)) {
                    mbar_wait_or_trap(&bar_m1[1], par);        // MMA 1 has read all of A: it may be overwritten
                    tc_fence_after();
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) tf32_split_fast(ab[e][c], hi[e], lo[e]);
                    tmem_st8(tbase + lane_base + COL_AHI + c * 32 + jb, hi);
                    tmem_st8(tbase + lane_base + COL_ALO + c * 32 + jb, lo);
                }
            }
            mbar_arrive(&bar_full[buf]);                       // releases this thread's As / zs stores to the gradient warps
            tmem_wait_st();
            tc_fence_before();
            named_bar_sync(1, NPT);                            // A complete; every a2 has been read from D
            if (warp == 0 && !(dbg & 2)) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int nh = 0; nh < 2; ++nh) {
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            issue_gemm_half(tbase + COL_D + c * 32 + nh * 16, tbase + COL_AHI + c * 32, tbase + COL_ALO + c * 32,
                                            b2d_hi, b2d_lo, nh);
                        mma_commit(&bar_m3[nh]);
                    }
                }
                __syncwarp();
            }

            // ---- E2: hbar1, recomputed h1 -> abar1 -> first-layer gradient partials (sum over the warp's 32 points) --
            if (!(dbg & 2)) mbar_wait_or_trap(&bar_m3[g], par);
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int kb = j0 + 8 * ch;
                uint32_t v[C][8];
#pragma unroll
                for (int c = 0; c < C; ++c) tmem_ld8(tbase + lane_base + COL_D + c * 32 + kb, v[c]);
                tmem_wait_ld();
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    float q[32];
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const int e = 4 * blk + e4;
                        float h1[C], hb[C], ab[C];
                        layer0(kb + e, z, h1);
#pragma unroll
                        for (int c = 0; c < C; ++c) hb[c] = __uint_as_float(v[c][e]);
                        fast_tanh_jets_bwd<CF>(h1, hb, ab);
                        q[e4 * 8 + 0] = ab[0];
                        q[e4 * 8 + 1] = ab[0] * z[0];
                        q[e4 * 8 + 2] = ab[0] * z[1];
                        q[e4 * 8 + 3] = ab[0] * z[2];
#pragma unroll
                        for (int s = 0; s < NS; ++s) q[e4 * 8 + 4 + s] = ab[s < NA2 ? 1 + 2 * s : 1 + 2 * NA2 + (s - NA2)];
#pragma unroll
                        for (int s = NQ0; s < 8; ++s) q[e4 * 8 + s] = 0.0f;
                    }
                    l0acc[ch][blk] += (dbg & 32) ? q[0] + q[9] + q[18] + q[27] : warp_transpose_reduce32(q, lane);
                }
            }
            if (!(dbg & 2)) mbar_wait_or_trap(&bar_m3[1], par);  // MMA 3 has read all of A before the next tile rewrites it
            tc_fence_after();
        }

        // ---- end of item: park the per-thread / per-lane partials in shared memory (As is free after the barrier) ----
        tc_fence_before();
        __syncthreads();
        float* red = As;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) red[L::RED_L0 + ((warp * 2 + ch) * 2 + blk) * 32 + lane] = l0acc[ch][blk];
#pragma unroll
        for (int u = 0; u < 16; ++u) red[L::RED_WL + u * NPT + tid] = wlacc[u];
        const float bv = fbp_warp_sum(blacc);
        if (lane == 0) red[L::RED_BL + warp] = bv;
    } else {
        // =============================================================================================
        // gradient warps: Wbar1[j][k] += sum_{c,p} abar2_c[p][j] h1_c[p][k],  bbar1[j] += sum_p abar2_0[p][j]
        // =============================================================================================
        const int gw = warp - NPW;                  // points [32 gw, 32 gw + 32) of every tile
        const int jt = lane >> 3, kt = lane & 7;    // accumulator ownership: j = jt + 4 jj, k = kt + 8 kk
        const int pt = lane & 7, kg = lane >> 3;    // h1 recompute ownership: point pt of the chunk, units kg + 4 i
        float* hs = sm + L::OFF_HS + gw * H * HS_ROW;
        float2 gacc[JJ][KK];
        float bacc[JJ];
#pragma unroll
        for (int jj = 0; jj < JJ; ++jj) {
            bacc[jj] = 0.0f;
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) gacc[jj][kk] = make_float2(0.0f, 0.0f);
        }
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            const uint32_t use = (uint32_t)(t >> 1);
            const float* Asb = As + buf * H * RS;
            const float* zsb = zs + buf * 3 * TP;
            mbar_wait_or_trap(&bar_full[buf], use & 1);
#pragma unroll 1
            for (int chk = 0; chk < ((dbg & 16) ? 0 : 4); ++chk) {
                const int p0 = 32 * gw + 8 * chk;
                // h1 of 8 points x 32 units into the warp-private buffer hs[k][c][pt]
                {
                    float z[3] = {zsb[p0 + pt], zsb[TP + p0 + pt], zsb[2 * TP + p0 + pt]};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = kg + 4 * i;
                        float h[C];
                        h[0] = fmaf(sm[CF::SM_W0 + 2 * H + k], z[2], fmaf(sm[CF::SM_W0 + H + k], z[1], fmaf(sm[CF::SM_W0 + k], z[0], sm[CF::SM_B0 + k])));
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            const float wv = sm[CF::SM_W0D + s * H + k];
                            if (s < NA2) { h[1 + 2 * s] = wv; h[2 + 2 * s] = 0.0f; }
                            else h[1 + 2 * NA2 + (s - NA2)] = wv;
                        }
                        fast_tanh_jets<CF>(h);
#pragma unroll
                        for (int c = 0; c < C; ++c) hs[k * HS_ROW + c * 8 + pt] = h[c];
                    }
                }
                __syncwarp();
                const float* ab = Asb + jt * RS + p0;
                const float* hp = hs + kt * HS_ROW;
#pragma unroll
                for (int pq = 0; pq < 8; pq += 4) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float4 av[JJ], hv[KK];
#pragma unroll
                        for (int jj = 0; jj < JJ; ++jj) av[jj] = *reinterpret_cast<const float4*>(ab + (4 * jj) * RS + c * TP + pq);
#pragma unroll
                        for (int kk = 0; kk < KK; ++kk) hv[kk] = *reinterpret_cast<const float4*>(hp + (8 * kk) * HS_ROW + c * 8 + pq);
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
                        if (c == 0) {
#pragma unroll
                            for (int jj = 0; jj < JJ; ++jj) bacc[jj] += (av[jj].x + av[jj].y) + (av[jj].z + av[jj].w);
                        }
                    }
                }
                __syncwarp();                                   // the next chunk overwrites hs
            }
            mbar_arrive(&bar_empty[buf]);
        }
        __syncthreads();                                        // pairs with the point warps' barrier: As is free
        float* red = As;
#pragma unroll
        for (int jj = 0; jj < JJ; ++jj) {
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) red[L::RED_G + (gw * (JJ * KK + JJ) + jj * KK + kk) * 32 + lane] = gacc[jj][kk].x + gacc[jj][kk].y;
            red[L::RED_G + (gw * (JJ * KK + JJ) + JJ * KK + jj) * 32 + lane] = bacc[jj];
        }
    }
    __syncthreads();

    // ---- this work item's partial gradients, every sum in fixed order ------------------------------------------
    const float* red = As;
    float* gp = a.gpart + (int64_t)item * a.P;
    // first layer: value (unit k, quantity t) sits in lane (k % 4) * 8 + t of block (k % 8) / 4, chunk (k % 16) / 8 of the
    // four point warps of unit half k / 16
    auto l0 = [&](int k, int t) {
        const int gg = k >> 4, ch = (k >> 3) & 1, blk = (k >> 2) & 1, ln = (k & 3) * 8 + t;
        float v = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) v += red[L::RED_L0 + (((gg * 4 + q) * 2 + ch) * 2 + blk) * 32 + ln];
        return v;
    };
    for (int i = tid; i < H * xd; i += BWD_NT) {
        const int j = i / xd, d = i - j * xd;
        float v = l0(j, 1 + d);
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (a.axis[s] == d) v = fmaf(sel3(d, isd[0], isd[1], isd[2]), l0(j, 4 + s), v);
        gp[i] = v;
    }
    for (int i = tid; i < H; i += BWD_NT) gp[H * xd + i] = l0(i, 0);
    int off = H * xd + H;
    for (int i = tid; i < H * H; i += BWD_NT) {
        const int j = i >> 5, k = i & 31;
        const int ln = (j & 3) * 8 + (k & 7), jj = j >> 2, kk = k >> 3;
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < NGW; ++w) v += red[L::RED_G + (w * (JJ * KK + JJ) + jj * KK + kk) * 32 + ln];
        gp[off + i] = v;
    }
    off += H * H;
    for (int j = tid; j < H; j += BWD_NT) {
        const int ln = (j & 3) * 8;                 // a lane with kt == 0
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < NGW; ++w) v += red[L::RED_G + (w * (JJ * KK + JJ) + JJ * KK + (j >> 2)) * 32 + ln];
        gp[off + j] = v;
    }
    off += H;
    for (int j = tid; j < H; j += BWD_NT) {         // output weights: sum over the 128 point threads of the unit's half
        const float* src = red + L::RED_WL + (j & 15) * NPT + (j >> 4) * 128;
        float v = 0.0f;
#pragma unroll 8
        for (int p = 0; p < 128; ++p) v += src[p];
        gp[off + j] = v;
    }
    if (tid == 0) {
        float v = 0.0f;
        for (int w = 0; w < 4; ++w) v += red[L::RED_BL + w];    // warps 0-3 are unit half 0 (the only ones that count it)
        gp[off + H] = v;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TMEM_COLS);
}

}  // namespace fbptc

int fbp_tc_backward_launch(const FastSpec& f, const FastArgs& a, int grid, cudaStream_t st);
