// Generic per-pair kernels for the reference's other Network plug-ins (fbpinns/networks.py:70-194): same structure as
// fbp_generic.cu (one thread per pair, any layer sizes, any closed jet set up to order 2), but the hidden layers apply
// an activation chosen per layer, possibly with trainable per-unit parameters:
//     FBP_ACT_TANH           tanh(a)                     (the hidden layers of FourierFCN behind its feature layer)
//     FBP_ACT_ADAPTIVE_TANH  alpha tanh(a / alpha)       AdaptiveFCN
//     FBP_ACT_SIN            sin(a)                      SIREN; FourierFCN's static feature layer [sin, cos](omega z)
//     FBP_ACT_ADAPTIVE_SIN   c sin(o a)                  AdaptiveSIREN
// With f0..f3 = f, f', f'', f''' at the value component a_0:
//     forward   h_0 = f0,  h_k = f1 a_k,  h_kl = f1 a_kl + f2 a_k a_l
//     reverse   abar_kl = f1 hbar_kl ;  abar_k = f1 hbar_k + f2 sum_(k,l) (1+delta_kl) hbar_kl a_l
//               abar_0  = f1 hbar_0 + f2 sum_k hbar_k a_k + sum_kl hbar_kl (f2 a_kl + f3 a_k a_l)
//               pbar    = sum_c hbar_c dh_c/dp  for the unit's activation parameters p
// (tests/proto_activation_math.py is the float64 transcription, checked against torch autograd on the CPU.)
// sin cannot be inverted for its derivative, so the scratch keeps the pre-activation jets next to the activations:
// [unit][a | h][component][thread].  Correctness family, not a roofline kernel; the plain-FCN kernels in fbp_generic.cu
// are left untouched.
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

// f[0..3] = f, f', f'', f''' at a0; dp[q][0..2] = d f / dq, d f' / dq, d f'' / dq for the activation parameters q.
// Accurate library functions on purpose (tanhf, sincosf): this family is the correctness baseline.
__device__ __forceinline__ void act_eval(int kind, float a0, float p0, float p1, float f[4], float dp[2][3]) {
    if (kind == FBP_ACT_TANH) {
        const float t = tanhf(a0), g = 1.0f - t * t;
        f[0] = t; f[1] = g; f[2] = -2.0f * t * g; f[3] = -2.0f * g * (g - 2.0f * t * t);
    } else if (kind == FBP_ACT_ADAPTIVE_TANH) {
        const float al = p0, y = a0 / al;
        const float t = tanhf(y), g = 1.0f - t * t;
        const float gp = -2.0f * t * g, gpp = -2.0f * g * (g - 2.0f * t * t);
        f[0] = al * t; f[1] = g; f[2] = gp / al; f[3] = gpp / (al * al);
        const float dy = -y / al;
        dp[0][0] = t + al * g * dy;
        dp[0][1] = gp * dy;
        dp[0][2] = gpp * dy / al - gp / (al * al);
    } else if (kind == FBP_ACT_SIN) {
        float s, c;
        sincosf(a0, &s, &c);
        f[0] = s; f[1] = c; f[2] = -s; f[3] = -c;
    } else {   // FBP_ACT_ADAPTIVE_SIN
        const float cc = p0, o = p1;
        float s, c;
        sincosf(o * a0, &s, &c);
        f[0] = cc * s; f[1] = cc * o * c; f[2] = -cc * o * o * s; f[3] = -cc * o * o * o * c;
        dp[0][0] = s; dp[0][1] = o * c; dp[0][2] = -o * o * s;
        dp[1][0] = cc * a0 * c; dp[1][1] = cc * (c - o * a0 * s); dp[1][2] = -cc * (2.0f * o * s + o * o * a0 * c);
    }
}

// scratch addressing: unit u (global hidden-unit index), which = 0 pre-activation / 1 activation, component c
__device__ __forceinline__ int64_t sidx(int u, int which, int c, int C, int64_t stride) {
    return ((int64_t)(u * 2 + which) * C + c) * stride;
}

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
                    for (int c = 0; c < C; ++c) a[c] = fmaf(wjk, sc[sidx(base + k, 1, c, C, stride)], a[c]);
                }
            }
            if (last) {
                for (int c = 0; c < C; ++c) r[j][c] = a[c];
            } else {
                const int kind = pd.lkind[l];
                const float p0 = pd.n_extra > 0 ? w[pd.eoff[l][0] + j] : 1.0f;
                const float p1 = pd.n_extra > 1 ? w[pd.eoff[l][1] + j] : 1.0f;
                float f[4], dp[2][3];
                act_eval(kind, a[0], p0, p1, f, dp);
                const int u = pd.hid_off[l] + j;
                for (int c = 0; c < C; ++c) sc[sidx(u, 0, c, C, stride)] = a[c];
                sc[sidx(u, 1, 0, C, stride)] = f[0];
                for (int c = 1; c < C; ++c) {
                    float v;
                    if (pd.ord[c] == 1) v = f[1] * a[c];
                    else v = fmaf(f[1], a[c], f[2] * a[pd.i1[c]] * a[pd.i2[c]]);
                    sc[sidx(u, 1, c, C, stride)] = v;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(GEN_THREADS)
act_forward_kernel(PlanDev pd, fbp_takes_view tv, const float* __restrict__ x, const float* __restrict__ params,
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

__device__ __forceinline__ void grad_add(float* addr, float v, bool uniform, int lane) {
    if (uniform) {
        v = fbp_warp_sum(v);
        if (lane == 0) atomicAdd(addr, v);
    } else if (v != 0.0f) {
        atomicAdd(addr, v);
    }
}

__global__ void __launch_bounds__(GEN_THREADS)
act_backward_kernel(PlanDev pd, fbp_takes_view tv, const float* __restrict__ x, const float* __restrict__ params,
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

    // ---- layers, last to first.  The pre-activation cotangent of hidden layer l overwrites its ACTIVATION slot
    //      (slot 1) once the layer above has consumed the activations; the pre-activation slot (slot 0) keeps a.
    for (int l = pd.nl - 1; l >= 0; --l) {
        const int nin = pd.size[l], nout = pd.size[l + 1];
        const bool last = (l == pd.nl - 1);
        const bool frozen = pd.lfrozen[l] != 0;
        const float* W = w + pd.woff[l];
        float* gW = g + pd.woff[l];
        float* gB = g + pd.boff[l];
        if (!frozen)
            for (int j = 0; j < nout; ++j) {
                float ab0 = last ? rbar[j][0] : (valid ? sc[sidx(pd.hid_off[l] + j, 1, 0, C, stride)] : 0.0f);
                grad_add(gB + j, ab0, uniform, lane);
            }
        for (int k = 0; k < nin; ++k) {
            float hin[FBP_MAX_COMP];
            if (l == 0) {
                hin[0] = pc.z[k];
                for (int c = 1; c < C; ++c) hin[c] = (pd.ord[c] == 1 && pd.ck[c] == k) ? pc.isd[k] : 0.0f;
            } else {
                for (int c = 0; c < C; ++c) hin[c] = valid ? sc[sidx(pd.hid_off[l - 1] + k, 1, c, C, stride)] : 0.0f;
            }
            float hbar[FBP_MAX_COMP];
            for (int c = 0; c < C; ++c) hbar[c] = 0.0f;
            for (int j = 0; j < nout; ++j) {
                float wjk = W[j * nin + k];
                float gw = 0.0f;
                for (int c = 0; c < C; ++c) {
                    float ab = last ? rbar[j][c] : (valid ? sc[sidx(pd.hid_off[l] + j, 1, c, C, stride)] : 0.0f);
                    gw = fmaf(ab, hin[c], gw);
                    hbar[c] = fmaf(wjk, ab, hbar[c]);
                }
                if (!frozen) grad_add(gW + j * nin + k, gw, uniform, lane);
            }
            if (l > 0) {
                // activation reverse for unit k of hidden layer l-1
                const int kind = pd.lkind[l - 1];
                const int u = pd.hid_off[l - 1] + k;
                float a[FBP_MAX_COMP];
                for (int c = 0; c < C; ++c) a[c] = valid ? sc[sidx(u, 0, c, C, stride)] : 0.0f;
                const float p0 = pd.n_extra > 0 ? w[pd.eoff[l - 1][0] + k] : 1.0f;
                const float p1 = pd.n_extra > 1 ? w[pd.eoff[l - 1][1] + k] : 1.0f;
                float f[4], dp[2][3];
                act_eval(kind, a[0], p0, p1, f, dp);
                float ab[FBP_MAX_COMP];
                float ab0 = f[1] * hbar[0];
                for (int c = 1; c < C; ++c) ab[c] = f[1] * hbar[c];
                for (int c = 1; c < C; ++c) {
                    if (pd.ord[c] == 1) {
                        ab0 = fmaf(f[2] * hbar[c], a[c], ab0);
                    } else {
                        const int a1 = pd.i1[c], a2 = pd.i2[c];
                        ab0 += hbar[c] * (f[2] * a[c] + f[3] * a[a1] * a[a2]);
                        ab[a1] = fmaf(f[2] * hbar[c], a[a2], ab[a1]);
                        ab[a2] = fmaf(f[2] * hbar[c], a[a1], ab[a2]);
                    }
                }
                ab[0] = ab0;
                // gradients of the unit's own activation parameters
                for (int q = 0; q < pd.n_extra; ++q) {
                    float v = hbar[0] * dp[q][0];
                    for (int c = 1; c < C; ++c) {
                        if (pd.ord[c] == 1) v = fmaf(hbar[c] * dp[q][1], a[c], v);
                        else v += hbar[c] * (dp[q][1] * a[c] + dp[q][2] * a[pd.i1[c]] * a[pd.i2[c]]);
                    }
                    grad_add(g + pd.eoff[l - 1][q] + k, valid ? v : 0.0f, uniform, lane);
                }
                if (valid)
                    for (int c = 0; c < C; ++c) sc[sidx(u, 1, c, C, stride)] = ab[c];
            }
        }
    }
}

__global__ void zero_kernel(float* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) p[i] = 0.0f;
}

__global__ void pack_extra_kernel(float* vec, float* packed, int64_t m, int n, int off, int P, int to_packed) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * n) return;
    const int64_t i = e / n;
    const int j = (int)(e - i * n);
    if (to_packed) packed[i * P + off + j] = vec[e];
    else vec[e] = packed[i * P + off + j];
}

int64_t chunk_pairs(const PlanDev& pd, int64_t pairs, int64_t scratch_floats, int64_t* per_pair_out) {
    const int64_t per_pair = (int64_t)pd.hid_total * pd.C * 2;
    *per_pair_out = per_pair;
    if (per_pair == 0) return pairs;
    int64_t chunk = scratch_floats / per_pair;
    chunk = (chunk / GEN_THREADS) * GEN_THREADS;
    return chunk > pairs ? pairs : chunk;
}

}  // namespace

#ifndef FBP_HOST_EMU      // launch glue (the CPU emulation test calls the kernels above directly)
int fbp_generic_act_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                            const float* d_sub_static, float* d_pair_out, float* d_scratch, int64_t scratch_floats,
                            cudaStream_t stream) {
    const PlanDev& pd = plan->dev;
    if (tv->s == 0) return 0;
    int64_t per_pair;
    FBP_REQUIRE(pd.hid_total == 0 || (d_scratch != nullptr && scratch_floats >= (int64_t)pd.hid_total * pd.C * 2 * GEN_THREADS),
                "fbp_forward(generic, activation variant): scratch too small: need >= %lld floats (have %lld)",
                (long long)((int64_t)pd.hid_total * pd.C * 2 * GEN_THREADS), (long long)scratch_floats);
    const int64_t chunk = chunk_pairs(pd, tv->s, scratch_floats, &per_pair);
    for (int64_t p0 = 0; p0 < tv->s; p0 += chunk) {
        int64_t np = tv->s - p0 < chunk ? tv->s - p0 : chunk;
        int blocks = (int)((np + GEN_THREADS - 1) / GEN_THREADS);
        act_forward_kernel<<<blocks, GEN_THREADS, 0, stream>>>(pd, *tv, d_x, d_params, d_sub_static, d_pair_out, d_scratch,
                                                               p0, np, chunk);
        FBP_LAUNCH_CHECK();
    }
    return 0;
}

int fbp_generic_act_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
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
    const int64_t s_active = tv->s_active;
    if (s_active == 0) return 0;
    int64_t per_pair;
    FBP_REQUIRE(pd.hid_total == 0 || (d_scratch != nullptr && scratch_floats >= (int64_t)pd.hid_total * pd.C * 2 * GEN_THREADS),
                "fbp_backward(generic, activation variant): scratch too small: need >= %lld floats (have %lld)",
                (long long)((int64_t)pd.hid_total * pd.C * 2 * GEN_THREADS), (long long)scratch_floats);
    const int64_t chunk = chunk_pairs(pd, s_active, scratch_floats, &per_pair);
    for (int64_t p0 = 0; p0 < s_active; p0 += chunk) {
        int64_t np = s_active - p0 < chunk ? s_active - p0 : chunk;
        int blocks = (int)((np + GEN_THREADS - 1) / GEN_THREADS);
        act_backward_kernel<<<blocks, GEN_THREADS, 0, stream>>>(pd, *tv, d_x, d_params, d_sub_static, d_grow, d_grads,
                                                                d_scratch, p0, np, chunk);
        FBP_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int fbp_pack_extra(const fbp_plan* plan, int64_t m, int32_t layer, int32_t which, float* d_vec, float* d_params,
                              int32_t to_packed, void* stream) {
    FBP_REQUIRE(plan && d_vec && d_params, "fbp_pack_extra: null argument");
    const PlanDev& pd = plan->dev;
    FBP_REQUIRE(layer >= 0 && layer < pd.nl && which >= 0 && which < pd.n_extra,
                "fbp_pack_extra: layer %d / parameter %d out of range for this plan (%d layers, %d activation parameters)",
                layer, which, pd.nl, pd.n_extra);
    const int n = pd.size[layer + 1];
    if (m * n == 0) return 0;
    const int64_t total = m * n;
    pack_extra_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_vec, d_params, m, n, pd.eoff[layer][which],
                                                                                   pd.P, to_packed);
    FBP_LAUNCH_CHECK();
    return 0;
}
#endif  // FBP_HOST_EMU
