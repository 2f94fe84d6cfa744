// Generic per-pair kernels: any FCN layer sizes, ud <= FBP_MAX_UD, any closed jet set up to order 2
// (including mixed second derivatives).  One thread per (point, subdomain) pair, weights read through
// L1/L2, hidden-layer jets kept in a caller-provided global scratch laid out [unit][component][thread] so
// that a warp's accesses are coalesced.  This family is the correctness baseline and the fallback for
// shapes the tiled family (fbp_fast_*.cu) does not instantiate; it is NOT the roofline kernel.
//
// Maths per pair (reference: FBPINN_model_inner under nested jvp, fbpinns/trainers.py:113-118, 213-247):
//   z = (x - mu)/sd ; jets of z: dz/dx_k = 1/sd_k, higher = 0
//   a_c = W h_c (+ b for the value component)
//   tanh jets: t = tanh(a_0), g = 1 - t^2, h_0 = t, h_k = g a_k, h_kl = g (a_kl - 2 t a_k a_l)
//   u_c = sd_u r_c (+ mu_u for c = 0),  N_c = Leibniz(u, w)
// Reverse (hand-derived, checked against torch autograd of the oracle in tests):
//   abar_kl = g hbar_kl
//   abar_k  = g hbar_k - 2 t sum_{(k,l)} (1+delta_kl) hbar_kl h_l
//   abar_0  = g hbar_0 - 2 t sum_k hbar_k h_k - 2 sum_{kl} hbar_kl (t h_kl + h_k h_l)
#include "fbp_common.cuh"

namespace {

constexpr int GEN_THREADS = 128;

struct PairCtx {
    int pt, sp, im;
    float z[FBP_MAX_XD], isd[FBP_MAX_XD];
    float flag, un_mu, un_sd;
};

__device__ __forceinline__ void load_pair(const PlanDev& pd, const fbp_takes_view& tv, const float* __restrict__ x,
                                          const float* __restrict__ sub_static, int64_t i, PairCtx& pc) {
    pc.pt = tv.d_spair_point[i];
    pc.sp = tv.d_spair_sub[i];
    pc.im = tv.d_sub_ids[pc.sp];
    const float* ss = sub_static + (int64_t)pc.im * pd.ss;
#pragma unroll
    for (int d = 0; d < FBP_MAX_XD; ++d) {
        if (d < pd.xd) {
            float lo = ss[d], hi = ss[pd.xd + d];
            float mu = (hi + lo) * 0.5f, sd = (hi - lo) * 0.5f;
            pc.isd[d] = 1.0f / sd;
            pc.z[d] = (x[(int64_t)pc.pt * pd.xd + d] - mu) * pc.isd[d];
        } else {
            pc.isd[d] = 0.0f;
            pc.z[d] = 0.0f;
        }
    }
    pc.flag = ss[2 * pd.xd];
    pc.un_mu = ss[2 * pd.xd + 1];
    pc.un_sd = ss[2 * pd.xd + 2];
}

// Forward through all layers for one pair. Hidden jets go to scratch column `sc` (stride `stride`);
// output-layer jets to r[o][c].
__device__ __forceinline__ void forward_pair(const PlanDev& pd, const PairCtx& pc, const float* __restrict__ w,
                                             float* sc, int64_t stride, float r[FBP_MAX_UD][FBP_MAX_COMP]) {
    const int C = pd.C;
    for (int l = 0; l < pd.nl; ++l) {
        const int nin = pd.size[l], nout = pd.size[l + 1];
        const float* W = w + pd.woff[l];
        const float* B = w + pd.boff[l];
        const bool last = (l == pd.nl - 1);
        for (int j = 0; j < nout; ++j) {
            float a[FBP_MAX_COMP];
#pragma unroll
            for (int c = 0; c < FBP_MAX_COMP; ++c) a[c] = 0.0f;
            a[0] = B[j];
            if (l == 0) {
                for (int k = 0; k < nin; ++k) {
                    float wjk = W[j * nin + k];
                    a[0] = fmaf(wjk, pc.z[k], a[0]);
                    for (int c = 1; c < C; ++c)
                        if (pd.ord[c] == 1 && pd.ck[c] == k) a[c] = fmaf(wjk, pc.isd[k], a[c]);
                }
            } else {
                const int base = pd.hid_off[l - 1];
                for (int k = 0; k < nin; ++k) {
                    float wjk = W[j * nin + k];
                    const float* hk = sc + (int64_t)(base + k) * C * stride;
                    for (int c = 0; c < C; ++c) a[c] = fmaf(wjk, hk[(int64_t)c * stride], a[c]);
                }
            }
            if (last) {
                for (int c = 0; c < C; ++c) r[j][c] = a[c];
            } else {
                float t = fbp_tanh(a[0]);
                float g = 1.0f - t * t;
                float* hj = sc + (int64_t)(pd.hid_off[l] + j) * C * stride;
                hj[0] = t;
                for (int c = 1; c < C; ++c) {
                    float v;
                    if (pd.ord[c] == 1) v = g * a[c];
                    else v = g * (a[c] - 2.0f * t * a[pd.i1[c]] * a[pd.i2[c]]);
                    hj[(int64_t)c * stride] = v;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(GEN_THREADS)
generic_forward_kernel(PlanDev pd, fbp_takes_view tv, const float* __restrict__ x, const float* __restrict__ params,
                       const float* __restrict__ sub_static, float* __restrict__ pair_out, float* scratch,
                       int64_t pair0, int64_t npairs, int64_t stride) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npairs) return;
    int64_t i = pair0 + t;
    PairCtx pc;
    load_pair(pd, tv, x, sub_static, i, pc);
    const float* w = params + (int64_t)pc.im * pd.P;
    float r[FBP_MAX_UD][FBP_MAX_COMP];
    forward_pair(pd, pc, w, scratch + t, stride, r);

    float win[FBP_MAX_COMP];
    fbp_window_jets(pd, pc.z, pc.isd, pc.flag, win);
    const int C = pd.C, ud = pd.ud;
    float* out = pair_out + i * (int64_t)(C * ud);
    for (int o = 0; o < ud; ++o) {
        float u[FBP_MAX_COMP];
        for (int c = 0; c < C; ++c) u[c] = pc.un_sd * r[o][c];
        u[0] += pc.un_mu;
        for (int c = 0; c < C; ++c) {
            float v;
            if (pd.ord[c] == 0) v = u[0] * win[0];
            else if (pd.ord[c] == 1) v = u[c] * win[0] + u[0] * win[c];
            else v = u[c] * win[0] + u[pd.i1[c]] * win[pd.i2[c]] + u[pd.i2[c]] * win[pd.i1[c]] + u[0] * win[c];
            out[c * ud + o] = v;
        }
    }
}

// Add `v` of every lane into *addr. When the whole warp targets the same subdomain the sum is formed with
// shuffles and one atomic is issued; otherwise each lane issues its own.
__device__ __forceinline__ void grad_add(float* addr, float v, bool uniform, int lane) {
    if (uniform) {
        v = fbp_warp_sum(v);
        if (lane == 0) atomicAdd(addr, v);
    } else if (v != 0.0f) {
        atomicAdd(addr, v);
    }
}

__global__ void __launch_bounds__(GEN_THREADS)
generic_backward_kernel(PlanDev pd, fbp_takes_view tv, const float* __restrict__ x, const float* __restrict__ params,
                        const float* __restrict__ sub_static, const float* __restrict__ grow, float* grads,
                        float* scratch, int64_t pair0, int64_t npairs, int64_t stride) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = t < npairs;
    // invalid tail threads shadow the last valid pair with zero cotangent so that warps stay converged
    int64_t i = pair0 + (valid ? t : npairs - 1);
    int64_t tcol = valid ? t : npairs - 1;   // tail threads never touch scratch (all accesses are guarded)
    PairCtx pc;
    load_pair(pd, tv, x, sub_static, i, pc);
    const int C = pd.C, ud = pd.ud;
    const float* w = params + (int64_t)pc.im * pd.P;
    float* sc = scratch + tcol;
    float r[FBP_MAX_UD][FBP_MAX_COMP];
    if (valid) forward_pair(pd, pc, w, sc, stride, r);
    __syncwarp();

    // subdomains past m_active are fixed: their pairs are not in [pair0, pair0+npairs) (host restricts the range)
    const int sp0 = __shfl_sync(0xffffffffu, pc.sp, 0);
    const bool uniform = __all_sync(0xffffffffu, pc.sp == sp0);
    float* g = grads + (int64_t)pc.sp * pd.P;

    float win[FBP_MAX_COMP];
    fbp_window_jets(pd, pc.z, pc.isd, pc.flag, win);

    // ---- output layer: rbar[o][c]
    float rbar[FBP_MAX_UD][FBP_MAX_COMP];
    const int row = tv.d_spair_row[i];
    for (int o = 0; o < ud; ++o) {
        float G[FBP_MAX_COMP], ub[FBP_MAX_COMP];
        for (int c = 0; c < C; ++c) G[c] = valid ? grow[(int64_t)row * (C * ud) + c * ud + o] : 0.0f;
        for (int c = 0; c < C; ++c) ub[c] = 0.0f;
        for (int c = 0; c < C; ++c) {
            ub[0] += G[c] * win[c];
            if (pd.ord[c] == 1) ub[c] += G[c] * win[0];
            else if (pd.ord[c] == 2) {
                ub[c] += G[c] * win[0];
                ub[pd.i1[c]] += G[c] * win[pd.i2[c]];
                ub[pd.i2[c]] += G[c] * win[pd.i1[c]];
            }
        }
        for (int c = 0; c < C; ++c) rbar[o][c] = pc.un_sd * ub[c];
    }

    // ---- layers, last to first. For layer l: abar (its pre-activation cotangent) lives in
    //      scratch[hid_off[l]] for hidden layers (overwriting h) and in rbar for the output layer.
    for (int l = pd.nl - 1; l >= 0; --l) {
        const int nin = pd.size[l], nout = pd.size[l + 1];
        const bool last = (l == pd.nl - 1);
        const float* W = w + pd.woff[l];
        float* gW = g + pd.woff[l];
        float* gB = g + pd.boff[l];
        // bias gradients
        for (int j = 0; j < nout; ++j) {
            float ab0 = last ? rbar[j][0] : (valid ? sc[(int64_t)(pd.hid_off[l] + j) * C * stride] : 0.0f);
            grad_add(gB + j, ab0, uniform, lane);
        }
        for (int k = 0; k < nin; ++k) {
            // input jets of this layer for unit k
            float hin[FBP_MAX_COMP];
            if (l == 0) {
                hin[0] = pc.z[k];
                for (int c = 1; c < C; ++c) hin[c] = (pd.ord[c] == 1 && pd.ck[c] == k) ? pc.isd[k] : 0.0f;
            } else {
                const float* hk = sc + (int64_t)(pd.hid_off[l - 1] + k) * C * stride;
                for (int c = 0; c < C; ++c) hin[c] = valid ? hk[(int64_t)c * stride] : 0.0f;
            }
            float hbar[FBP_MAX_COMP];
            for (int c = 0; c < C; ++c) hbar[c] = 0.0f;
            for (int j = 0; j < nout; ++j) {
                float wjk = W[j * nin + k];
                float gw = 0.0f;
                for (int c = 0; c < C; ++c) {
                    float ab = last ? rbar[j][c]
                                    : (valid ? sc[((int64_t)(pd.hid_off[l] + j) * C + c) * stride] : 0.0f);
                    gw = fmaf(ab, hin[c], gw);
                    hbar[c] = fmaf(wjk, ab, hbar[c]);
                }
                grad_add(gW + j * nin + k, gw, uniform, lane);
            }
            if (l > 0) {
                // tanh reverse for unit k of hidden layer l-1, in place
                float tt = hin[0];
                float gg = 1.0f - tt * tt;
                float ab[FBP_MAX_COMP];
                float ab0 = gg * hbar[0];
                for (int c = 1; c < C; ++c) ab[c] = gg * hbar[c];
                for (int c = 1; c < C; ++c) {
                    if (pd.ord[c] == 1) {
                        ab0 -= 2.0f * tt * hbar[c] * hin[c];
                    } else {
                        int a1 = pd.i1[c], a2 = pd.i2[c];
                        ab0 -= 2.0f * hbar[c] * (tt * hin[c] + hin[a1] * hin[a2]);
                        ab[a1] -= 2.0f * tt * hbar[c] * hin[a2];
                        ab[a2] -= 2.0f * tt * hbar[c] * hin[a1];
                    }
                }
                ab[0] = ab0;
                if (valid) {
                    float* hk = sc + (int64_t)(pd.hid_off[l - 1] + k) * C * stride;
                    for (int c = 0; c < C; ++c) hk[(int64_t)c * stride] = ab[c];
                }
            }
        }
    }
}

__global__ void zero_kernel(float* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) p[i] = 0.0f;
}

}  // namespace

#ifndef FBP_HOST_EMU      // launch glue (the CPU emulation test calls the kernels above directly)
int fbp_generic_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                        const float* d_sub_static, float* d_pair_out, float* d_scratch, int64_t scratch_floats,
                        cudaStream_t stream) {
    const PlanDev& pd = plan->dev;
    if (tv->s == 0) return 0;
    int64_t per_pair = (int64_t)pd.hid_total * pd.C;
    int64_t chunk = tv->s;
    if (per_pair > 0) {
        FBP_REQUIRE(d_scratch != nullptr && scratch_floats >= per_pair * GEN_THREADS,
                    "fbp_forward(generic): scratch too small: need >= %lld floats (have %lld)",
                    (long long)(per_pair * GEN_THREADS), (long long)scratch_floats);
        chunk = scratch_floats / per_pair;
        chunk = (chunk / GEN_THREADS) * GEN_THREADS;
        if (chunk > tv->s) chunk = tv->s;
    }
    for (int64_t p0 = 0; p0 < tv->s; p0 += chunk) {
        int64_t np = tv->s - p0 < chunk ? tv->s - p0 : chunk;
        int blocks = (int)((np + GEN_THREADS - 1) / GEN_THREADS);
        generic_forward_kernel<<<blocks, GEN_THREADS, 0, stream>>>(pd, *tv, d_x, d_params, d_sub_static, d_pair_out,
                                                                   d_scratch, p0, np, chunk);
        FBP_LAUNCH_CHECK();
    }
    return 0;
}

int fbp_generic_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                         const float* d_sub_static, const float* d_grow, float* d_grads, int accumulate,
                         float* d_scratch, int64_t scratch_floats, cudaStream_t stream) {
    const PlanDev& pd = plan->dev;
    int64_t ng = (int64_t)tv->m_active * pd.P;
    if (!accumulate && ng > 0) {
        int blocks = (int)((ng + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        zero_kernel<<<blocks, 256, 0, stream>>>(d_grads, ng);
        FBP_LAUNCH_CHECK();
    }
    const int64_t s_active = tv->s_active;   // pairs of the trained (leading) subdomains
    if (s_active == 0) return 0;
    int64_t per_pair = (int64_t)pd.hid_total * pd.C;
    int64_t chunk = s_active;
    if (per_pair > 0) {
        FBP_REQUIRE(d_scratch != nullptr && scratch_floats >= per_pair * GEN_THREADS,
                    "fbp_backward(generic): scratch too small: need >= %lld floats (have %lld)",
                    (long long)(per_pair * GEN_THREADS), (long long)scratch_floats);
        chunk = scratch_floats / per_pair;
        chunk = (chunk / GEN_THREADS) * GEN_THREADS;
        if (chunk > s_active) chunk = s_active;
    }
    for (int64_t p0 = 0; p0 < s_active; p0 += chunk) {
        int64_t np = s_active - p0 < chunk ? s_active - p0 : chunk;
        int blocks = (int)((np + GEN_THREADS - 1) / GEN_THREADS);
        generic_backward_kernel<<<blocks, GEN_THREADS, 0, stream>>>(pd, *tv, d_x, d_params, d_sub_static, d_grow,
                                                                    d_grads, d_scratch, p0, np, chunk);
        FBP_LAUNCH_CHECK();
    }
    return 0;
}
#endif  // FBP_HOST_EMU
