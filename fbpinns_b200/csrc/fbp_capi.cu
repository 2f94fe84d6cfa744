// C ABI of libfbpinn_b200 (see include/fbpinn_b200.h): plan management, validation and dispatch between the
// tiled (fbp_fast_*.cu) and generic (fbp_generic.cu) kernel families.
#include "fbp_common.cuh"
#include <stdlib.h>

#include <stdarg.h>
#include <string.h>

#include <atomic>

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void fbp_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void fbp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

const char* fbp_last_error(void) { return g_err; }

int fbp_version(void) { return 100; }

int64_t fbp_launch_count(void) { return (int64_t)g_launches.load(); }

int fbp_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* smem_per_block_optin) {
    int dev = 0;
    FBP_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    FBP_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_per_block_optin) *smem_per_block_optin = (int64_t)prop.sharedMemPerBlockOptin;
    FBP_REQUIRE(prop.major == 10, "libfbpinn_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    return 0;
}

int fbp_plan_create(fbp_plan** out, const fbp_plan_desc* d) {
    FBP_REQUIRE(out && d, "fbp_plan_create: null argument");
    *out = nullptr;
    FBP_REQUIRE(d->xd >= 1 && d->xd <= FBP_MAX_XD, "fbp_plan_create: xd=%d unsupported (1..%d)", d->xd, FBP_MAX_XD);
    FBP_REQUIRE(d->ud >= 1 && d->ud <= FBP_MAX_UD, "fbp_plan_create: ud=%d unsupported (1..%d)", d->ud, FBP_MAX_UD);
    FBP_REQUIRE(d->n_layers >= 1 && d->n_layers <= FBP_MAX_LAYERS, "fbp_plan_create: n_layers=%d unsupported (1..%d)",
                d->n_layers, FBP_MAX_LAYERS);
    FBP_REQUIRE(d->layer_sizes[0] == d->xd, "fbp_plan_create: layer_sizes[0]=%d must equal xd=%d", d->layer_sizes[0], d->xd);
    FBP_REQUIRE(d->layer_sizes[d->n_layers] == d->ud, "fbp_plan_create: last layer size %d must equal ud=%d",
                d->layer_sizes[d->n_layers], d->ud);
    FBP_REQUIRE(d->activation >= FBP_ACT_TANH && d->activation <= FBP_ACT_FOURIER_TANH,
                "fbp_plan_create: activation %d not implemented (FBP_ACT_* 0..4)", d->activation);
    FBP_REQUIRE(d->activation != FBP_ACT_FOURIER_TANH || d->n_layers >= 2,
                "fbp_plan_create: FBP_ACT_FOURIER_TANH needs the feature layer plus at least one more layer");
    FBP_REQUIRE(d->window == FBP_WINDOW_COSINE, "fbp_plan_create: window %d not implemented (cosine only)", d->window);
    FBP_REQUIRE(d->n_comp >= 1 && d->n_comp <= FBP_MAX_COMP, "fbp_plan_create: n_comp=%d unsupported (1..%d)", d->n_comp, FBP_MAX_COMP);
    FBP_REQUIRE(d->comp_k[0] < 0 && d->comp_l[0] < 0, "fbp_plan_create: component 0 must be the value");

    fbp_plan* p = new fbp_plan();
    memset(p, 0, sizeof(*p));
    p->desc = *d;
    PlanDev& pd = p->dev;
    pd.xd = d->xd; pd.ud = d->ud; pd.nl = d->n_layers; pd.ss = 2 * d->xd + 3;
    int off = 0, hoff = 0;
    for (int l = 0; l <= d->n_layers; ++l) {
        if (d->layer_sizes[l] < 1) { delete p; FBP_REQUIRE(false, "fbp_plan_create: layer size must be >= 1"); }
        pd.size[l] = d->layer_sizes[l];
    }
    pd.act = d->activation;
    pd.n_extra = d->activation == FBP_ACT_ADAPTIVE_TANH ? 1 : (d->activation == FBP_ACT_ADAPTIVE_SIN ? 2 : 0);
    for (int l = 0; l < d->n_layers; ++l) {
        pd.woff[l] = off; off += pd.size[l] * pd.size[l + 1];
        pd.boff[l] = off; off += pd.size[l + 1];
        for (int e = 0; e < pd.n_extra; ++e) { pd.eoff[l][e] = off; off += pd.size[l + 1]; }
        pd.lkind[l] = d->activation == FBP_ACT_FOURIER_TANH ? (l == 0 ? FBP_ACT_SIN : FBP_ACT_TANH) : d->activation;
        pd.lfrozen[l] = (d->activation == FBP_ACT_FOURIER_TANH && l == 0) ? 1 : 0;
        if (l < d->n_layers - 1) { pd.hid_off[l] = hoff; hoff += pd.size[l + 1]; }
    }
    pd.P = off;
    pd.hid_total = hoff;
    pd.C = d->n_comp;
    // component tables + closure check
    for (int c = 0; c < d->n_comp; ++c) {
        int k = d->comp_k[c], l = d->comp_l[c];
        pd.ck[c] = k; pd.cl[c] = l; pd.i1[c] = 0; pd.i2[c] = 0;
        pd.ord[c] = (k < 0) ? 0 : (l < 0 ? 1 : 2);
        bool ok = (c == 0) ? (k < 0 && l < 0) : (k >= 0 && k < d->xd && l < d->xd);
        if (c > 0 && k < 0) ok = false;
        if (!ok) { delete p; FBP_REQUIRE(false, "fbp_plan_create: bad jet component %d (k=%d, l=%d)", c, k, l); }
    }
    for (int c = 0; c < d->n_comp; ++c) {
        for (int c2 = 0; c2 < c; ++c2)
            if (pd.ck[c2] == pd.ck[c] && pd.cl[c2] == pd.cl[c]) { delete p; FBP_REQUIRE(false, "fbp_plan_create: duplicate jet component %d", c); }
        if (pd.ord[c] != 2) continue;
        int f1 = -1, f2 = -1;
        for (int c2 = 0; c2 < d->n_comp; ++c2) {
            if (pd.ord[c2] == 1 && pd.ck[c2] == pd.ck[c]) f1 = c2;
            if (pd.ord[c2] == 1 && pd.ck[c2] == pd.cl[c]) f2 = c2;
        }
        if (f1 < 0 || f2 < 0) { delete p; FBP_REQUIRE(false, "fbp_plan_create: jet set not closed: order-2 component %d lacks its order-1 components", c); }
        pd.i1[c] = f1; pd.i2[c] = f2;
    }
    p->fast_id = d->activation == FBP_ACT_TANH ? fbp_fast_lookup(d, &p->fast) : -1;   // activation variants: generic family
    p->tc_ok = p->fast_id >= 0 && fbp_tc_supported(p->fast, p->dev.C) != 0;
    {   // instance validated on B200 (profiles/r1f_tc_bringup.md): two second-order axes, C = 5 (cfg 5)
        const char* e = getenv("FBP_TC_AUTO");
        p->tc_auto = p->tc_ok && p->fast.na2 == 2 && p->fast.na1 == 0 && !(e && e[0] == '0');
        // the tensor reverse kernel is part of auto mode since the round-2 timing run (profiles/r2a_tc_bringup.md:
        // tensor forward v2 without cache stores 1.38 ms + tensor reverse 5.21 ms beat 1.72 + 5.26 ms);
        // FBP_TC_AUTO=fwd keeps the tiled reverse kernel with the activation cache
        p->tc_auto_bwd = !(e && strcmp(e, "fwd") == 0);
    }
    p->mode = 0;
    *out = p;
    return 0;
}

int fbp_plan_destroy(fbp_plan* plan) {
    delete plan;
    return 0;
}

int64_t fbp_plan_param_count(const fbp_plan* plan) { return plan ? plan->dev.P : -1; }
int32_t fbp_plan_is_fast(const fbp_plan* plan) { return plan && plan->fast_id >= 0 ? 1 : 0; }
int32_t fbp_plan_tile_points(const fbp_plan* plan) { return (plan && plan->use_fast()) ? plan->fast.tile_points : 128; }

int32_t fbp_plan_n_extra(const fbp_plan* plan) { return plan ? plan->dev.n_extra : -1; }
int32_t fbp_plan_has_tensor(const fbp_plan* plan) { return plan && plan->fast_id >= 0 && plan->tc_ok ? 1 : 0; }
int32_t fbp_plan_forward_family(const fbp_plan* plan) {
    if (!plan) return -1;
    return !plan->use_fast() ? 0 : (plan->use_tc() ? 2 : 1);
}

int32_t fbp_plan_reverse_family(const fbp_plan* plan) {
    if (!plan) return -1;
    return !plan->use_fast() ? 0 : (plan->use_tc_bwd() ? 2 : 1);
}

int fbp_plan_set_kernel(fbp_plan* plan, int32_t mode) {
    FBP_REQUIRE(plan, "fbp_plan_set_kernel: null plan");
    FBP_REQUIRE(mode >= 0 && mode <= 4, "fbp_plan_set_kernel: mode must be 0 .. 4");
    FBP_REQUIRE(mode != 2 || plan->fast_id >= 0, "fbp_plan_set_kernel: no tiled kernel instance for this plan");
    FBP_REQUIRE((mode != 3 && mode != 4) || (plan->fast_id >= 0 && plan->tc_ok),
                "fbp_plan_set_kernel: no tensor (tcgen05) kernel instance for this plan (needs H = 32, two hidden layers, "
                "at most 5 jet components)");
    plan->mode = mode;
    return 0;
}

int64_t fbp_plan_cache_per_pair(const fbp_plan* plan) {
    if (!plan) return -1;
    if (plan->use_tc_bwd()) return 0;      // the tensor reverse kernel recomputes the hidden layer on the tensor core
    return (plan->use_fast() && plan->fast.nhid == 2) ? (int64_t)plan->fast.H * plan->dev.C : 0;
}

int64_t fbp_plan_scratch_per_pair(const fbp_plan* plan) {
    if (!plan) return -1;
    // the activation variants keep the pre-activation jets next to the activations
    return plan->use_fast() ? 0 : (int64_t)plan->dev.hid_total * plan->dev.C * (plan->dev.act != FBP_ACT_TANH ? 2 : 1);
}

static int check_view(const fbp_takes_view* tv, const char* who) {
    FBP_REQUIRE(tv, "%s: null takes view", who);
    FBP_REQUIRE(tv->s < (1ll << 31) && tv->n < (1ll << 31), "%s: sizes exceed int32 indexing", who);
    FBP_REQUIRE(tv->m_active >= 0 && tv->m_active <= tv->m_all, "%s: m_active out of range", who);
    return 0;
}

int fbp_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                const float* d_sub_static, float* d_pair_out, float* d_scratch, int64_t scratch_floats, float* d_act_cache,
                void* stream) {
    FBP_REQUIRE(plan, "fbp_forward: null plan");
    if (int rc = check_view(tv, "fbp_forward")) return rc;
    if (plan->use_fast())
        return fbp_fast_forward(plan, tv, d_x, d_params, d_sub_static, d_pair_out, d_act_cache, (cudaStream_t)stream);
    if (plan->dev.act != FBP_ACT_TANH)
        return fbp_generic_act_forward(plan, tv, d_x, d_params, d_sub_static, d_pair_out, d_scratch, scratch_floats,
                                       (cudaStream_t)stream);
    return fbp_generic_forward(plan, tv, d_x, d_params, d_sub_static, d_pair_out, d_scratch, scratch_floats, (cudaStream_t)stream);
}

int64_t fbp_backward_workspace_floats(const fbp_plan* plan, const fbp_takes_view* tv) {
    if (!plan || !tv) return -1;
    return plan->use_fast() ? fbp_fast_backward_workspace(plan, tv) : 0;
}

int fbp_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                 const float* d_sub_static, const float* d_grow, float* d_grads, int32_t accumulate, float* d_gpart,
                 float* d_scratch, int64_t scratch_floats, const float* d_act_cache, void* stream) {
    FBP_REQUIRE(plan, "fbp_backward: null plan");
    if (int rc = check_view(tv, "fbp_backward")) return rc;
    if (plan->use_fast())
        return fbp_fast_backward(plan, tv, d_x, d_params, d_sub_static, d_grow, d_grads, accumulate, d_gpart, d_act_cache,
                                 (cudaStream_t)stream);
    accumulate &= ~FBP_BWD_DIRECT;        // the generic kernels have no partial buffer: they always write d_grads themselves
    FBP_REQUIRE((accumulate & ~FBP_BWD_ACCUMULATE) == 0,
                "fbp_backward(generic): FBP_BWD_NO_REDUCE / FBP_BWD_REDUCE_ONLY need a tiled plan");
    if (plan->dev.act != FBP_ACT_TANH)
        return fbp_generic_act_backward(plan, tv, d_x, d_params, d_sub_static, d_grow, d_grads, accumulate, d_scratch,
                                        scratch_floats, (cudaStream_t)stream);
    return fbp_generic_backward(plan, tv, d_x, d_params, d_sub_static, d_grow, d_grads, accumulate, d_scratch,
                                scratch_floats, (cudaStream_t)stream);
}

}  // extern "C"
