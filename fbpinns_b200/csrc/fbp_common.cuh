// Internal definitions shared by the kernels of libfbpinn_b200 (sm_100a only).
#pragma once

#ifdef FBP_HOST_EMU
#include "fbp_host_emu.h"      // tests/tools: the per-pair kernels compiled as plain C++ for the CPU emulation test
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fbpinn_b200.h"

#define FBP_PI_F 3.14159265358979323846f

// ---------------------------------------------------------------------------------------------------
// error handling (thread-local message, never throws across the C ABI)
// ---------------------------------------------------------------------------------------------------
void fbp_set_error(const char* fmt, ...);
void fbp_count_launch();   // every kernel launch of this library is counted (fbp_launch_count)

#define FBP_CHECK_CUDA(expr)                                                                        \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            fbp_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

#define FBP_REQUIRE(cond, ...)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            fbp_set_error(__VA_ARGS__);                                                             \
            return 2;                                                                               \
        }                                                                                           \
    } while (0)

#define FBP_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        fbp_count_launch();                                                                         \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            fbp_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------
// Device-side (by-value kernel parameter) description of the network + jet set.
struct PlanDev {
    int xd, ud, nl;                     // nl = linear layers
    int size[FBP_MAX_LAYERS + 1];
    int woff[FBP_MAX_LAYERS];           // offset of W_l inside a packed parameter row
    int boff[FBP_MAX_LAYERS];           // offset of b_l
    int hid_off[FBP_MAX_LAYERS];        // unit offset of hidden layer l in the generic scratch
    int hid_total;                      // sum of hidden widths
    int P;                              // parameters per subdomain
    int C;                              // jet components
    int ord[FBP_MAX_COMP];              // 0, 1, 2
    int ck[FBP_MAX_COMP];               // first axis (or -1)
    int cl[FBP_MAX_COMP];               // second axis (or -1)
    int i1[FBP_MAX_COMP];               // order 2: component index of d/dx_k
    int i2[FBP_MAX_COMP];               // order 2: component index of d/dx_l
    int ss;                             // floats per static subdomain record = 2*xd+3
    // activation variants (generic family only; all zero for plain FCN plans)
    int act;                            // FBP_ACT_*
    int n_extra;                        // activation parameters per unit (0, 1, 2)
    int lkind[FBP_MAX_LAYERS];          // per hidden layer: FBP_ACT_TANH / ADAPTIVE_TANH / SIN / ADAPTIVE_SIN
    int lfrozen[FBP_MAX_LAYERS];        // 1: the layer's W, b get no gradient (Fourier feature layer)
    int eoff[FBP_MAX_LAYERS][2];        // offsets of the layer's activation parameters inside a packed row
};

// Tiled ("fast") kernel family parameters, valid when fast_id >= 0.
struct FastSpec {
    int H;        // hidden width (all hidden layers equal)
    int nhid;     // number of hidden layers (>= 1)
    int na2;      // axis slots carrying (d, dd)
    int na1;      // axis slots carrying (d) only
    int axis[FBP_MAX_XD];   // slot -> axis (na2 slots first, then na1 slots)
    int ext[FBP_MAX_COMP];  // kernel-internal component -> external component index
    int tile_points;
};

struct fbp_plan {
    fbp_plan_desc desc;
    PlanDev dev;
    int fast_id;          // -1: no tiled instance
    FastSpec fast;
    bool tc_ok;           // the tcgen05 ("tensor") family has an instance for this plan
    bool tc_auto;         // ... and that instance has been validated on hardware, so mode 0 picks it (FBP_TC_AUTO=0: never)
    bool tc_auto_bwd;     // FBP_TC_AUTO=full: mode 0 also picks the tensor reverse kernel (switch for the next hardware session)
    int mode;             // 0 auto, 1 generic, 2 tiled, 3 tensor forward + tiled reverse, 4 tensor forward and reverse
    bool use_fast() const { return mode == 1 ? false : fast_id >= 0; }
    bool use_tc() const { return ((mode == 3 || mode == 4) && tc_ok) || (mode == 0 && tc_auto); }
    bool use_tc_bwd() const { return (mode == 4 && tc_ok) || (mode == 0 && tc_auto && tc_auto_bwd); }
};

// ---------------------------------------------------------------------------------------------------
// device math
// ---------------------------------------------------------------------------------------------------
#if defined(__CUDACC__) || defined(FBP_HOST_EMU)

// tanh with ~1.5e-7 absolute error in 5 instructions: 1 - 2/(exp(2x)+1) with exp through MUFU.EX2
// (ex2.approx.ftz) and the reciprocal through MUFU.RCP (rcp.approx.ftz), folded into one FFMA.
// Saturates correctly: x -> +inf gives e = inf -> rcp = 0 -> 1 ; x -> -inf gives e = 0 -> 1 - 2 = -1.
__device__ __forceinline__ float fbp_tanh(float x) {
#ifdef FBP_HOST_EMU
    return tanhf(x);
#else
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
#endif
}

// Per-dimension cosine window f(z) = ((1+cos(pi z))/2)^2 with z = (x-mu)/sd and its x-derivatives:
//   f  = A^2,  f' = -kappa sin(pi z) A,  f'' = -kappa^2 (2 cos(pi z) - 1) A,   A = (1+cos(pi z))/2, kappa = pi/sd.
// (windows.cosine, fbpinns/windows.py:25-35; the heaviside factors are 1 for every pair of the takes, which
//  are exactly the pairs passing the same float32 inside test, and have zero derivative in JAX.)
__device__ __forceinline__ void fbp_window_dim(float z, float inv_sd, float& f, float& f1, float& f2) {
    float s, c;
    sincospif(z, &s, &c);
    float A = 0.5f * (1.0f + c);
    float kap = FBP_PI_F * inv_sd;
    f = A * A;
    f1 = -kap * s * A;
    f2 = -kap * kap * (2.0f * c - 1.0f) * A;
}

// Window jets for an arbitrary closed component set. w[c] for c < pd.C.
__device__ __forceinline__ void fbp_window_jets(const PlanDev& pd, const float* z, const float* inv_sd, float flag,
                                                float* w) {
    float f[FBP_MAX_XD], f1[FBP_MAX_XD], f2[FBP_MAX_XD];
#pragma unroll
    for (int d = 0; d < FBP_MAX_XD; ++d) {
        if (d < pd.xd) fbp_window_dim(z[d], inv_sd[d], f[d], f1[d], f2[d]);
        else { f[d] = 1.0f; f1[d] = 0.0f; f2[d] = 0.0f; }
    }
    for (int c = 0; c < pd.C; ++c) {
        float prod = flag;
        int k = pd.ck[c], l = pd.cl[c];
#pragma unroll
        for (int d = 0; d < FBP_MAX_XD; ++d) {
            if (d < pd.xd) {
                int cnt = (d == k) + (d == l);
                prod *= (cnt == 0) ? f[d] : (cnt == 1 ? f1[d] : f2[d]);
            }
        }
        w[c] = (c == 0) ? prod + (1.0f - flag) : prod;
    }
}

__device__ __forceinline__ float fbp_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------------
// internal launchers (one per translation unit); all return 0 / non-zero with fbp_set_error
// ---------------------------------------------------------------------------------------------------
int fbp_generic_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                        const float* d_sub_static, float* d_pair_out, float* d_scratch, int64_t scratch_floats,
                        cudaStream_t stream);
int fbp_generic_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                         const float* d_sub_static, const float* d_grow, float* d_grads, int accumulate,
                         float* d_scratch, int64_t scratch_floats, cudaStream_t stream);

// generic family, activation variants (fbp_generic_act.cu): plans with dev.act != FBP_ACT_TANH
int fbp_generic_act_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                            const float* d_sub_static, float* d_pair_out, float* d_scratch, int64_t scratch_floats,
                            cudaStream_t stream);
int fbp_generic_act_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                             const float* d_sub_static, const float* d_grow, float* d_grads, int accumulate,
                             float* d_scratch, int64_t scratch_floats, cudaStream_t stream);

// tiled family: returns -1 if no instance matches
int fbp_fast_lookup(const fbp_plan_desc* desc, FastSpec* spec);
// tensor (tcgen05) family: 1 if it has an instance for this tiled spec with C jet components
int fbp_tc_supported(const FastSpec& f, int C);
int fbp_fast_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                     const float* d_sub_static, float* d_pair_out, float* d_cache, cudaStream_t stream);
int64_t fbp_fast_backward_workspace(const fbp_plan* plan, const fbp_takes_view* tv);
int fbp_fast_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                      const float* d_sub_static, const float* d_grow, float* d_grads, int accumulate, float* d_gpart,
                      const float* d_cache, cudaStream_t stream);
