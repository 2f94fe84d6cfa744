#include <cstdlib>
// Tiled kernel family: plan matching, launch glue and the deterministic partial-gradient reduction.
#include "fbp_tc_bwd.cuh"

__global__ void fast_grad_reduce_kernel(const float* __restrict__ gpart, const int32_t* __restrict__ sub_item_off,
                                        int m_active, int P, float* __restrict__ grads, int accumulate) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)m_active * P) return;
    const int sp = (int)(e / P);
    const int k = (int)(e - (int64_t)sp * P);
    float v = accumulate ? grads[e] : 0.0f;
    for (int it = sub_item_off[sp]; it < sub_item_off[sp + 1]; ++it) v += gpart[(int64_t)it * P + k];
    grads[e] = v;
}

// Which plans the tiled family covers: FCN [xd, H, (H,) 1] with H in {16, 32, 64}, ud = 1, pure per-axis jets of
// order <= 2 in one of the instantiated (second-order slots, first-order-only slots) shapes.
int fbp_fast_lookup(const fbp_plan_desc* d, FastSpec* spec) {
    const int nhid = d->n_layers - 1;
    if (d->ud != 1 || nhid < 1 || nhid > 2) return -1;
    const int H = d->layer_sizes[1];
    if (H != 16 && H != 32 && H != 64) return -1;
    for (int l = 1; l <= nhid; ++l)
        if (d->layer_sizes[l] != H) return -1;
    int has1[FBP_MAX_XD] = {0, 0, 0}, has2[FBP_MAX_XD] = {0, 0, 0};
    int c1[FBP_MAX_XD] = {-1, -1, -1}, c2[FBP_MAX_XD] = {-1, -1, -1};
    for (int c = 1; c < d->n_comp; ++c) {
        int k = d->comp_k[c], l = d->comp_l[c];
        if (l < 0) { has1[k] = 1; c1[k] = c; }
        else {
            if (k != l) return -1;      // mixed derivative: generic family
            has2[k] = 1; c2[k] = c;
        }
    }
    int na2 = 0, na1 = 0;
    spec->ext[0] = 0;
    for (int k = 0; k < d->xd; ++k)
        if (has2[k]) {
            if (!has1[k]) return -1;
            spec->axis[na2] = k;
            spec->ext[1 + 2 * na2] = c1[k];
            spec->ext[2 + 2 * na2] = c2[k];
            ++na2;
        }
    for (int k = 0; k < d->xd; ++k)
        if (has1[k] && !has2[k]) {
            spec->axis[na2 + na1] = k;
            spec->ext[1 + 2 * na2 + na1] = c1[k];
            ++na1;
        }
    const int key = na2 * 4 + na1;
    if (!(key == 0 || key == 1 || key == 4 || key == 5 || key == 8 || key == 12)) return -1;
    for (int s = na2 + na1; s < FBP_MAX_XD; ++s) spec->axis[s] = 0;
    spec->H = H; spec->nhid = nhid; spec->na2 = na2; spec->na1 = na1;
    spec->tile_points = 128;
    return H * 1000 + nhid * 100 + key;
}

static void fill_args(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                      const float* d_sub_static, FastArgs& a) {
    a.x = d_x; a.params = d_params; a.sub_static = d_sub_static;
    a.sub_ids = tv->d_sub_ids; a.spair_point = tv->d_spair_point; a.spair_row = tv->d_spair_row; a.items = tv->d_items;
    a.pair_out = nullptr; a.grow = nullptr; a.gpart = nullptr; a.cache = nullptr; a.order = nullptr; a.launch = nullptr;
    a.xd = plan->dev.xd; a.P = plan->dev.P; a.dbg = 0;
    for (int i = 0; i < FBP_MAX_XD; ++i) a.axis[i] = plan->fast.axis[i];
    for (int i = 0; i < FBP_MAX_COMP; ++i) a.ext[i] = plan->fast.ext[i];
}

static int dispatch(const fbp_plan* plan, bool backward, const FastArgs& a, int grid, cudaStream_t st) {
    const FastSpec& f = plan->fast;
    switch (f.H) {
        case 16: return fbp_fast_launch_h16(f.nhid, f.na2, f.na1, backward, a, grid, st);
        case 32: return fbp_fast_launch_h32(f.nhid, f.na2, f.na1, backward, a, grid, st);
        case 64: return fbp_fast_launch_h64(f.nhid, f.na2, f.na1, backward, a, grid, st);
    }
    fbp_set_error("fbp_fast: no instance for H=%d", f.H);
    return 3;
}

int fbp_fast_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                     const float* d_sub_static, float* d_pair_out, float* d_cache, cudaStream_t stream) {
    if (tv->n_items == 0) return 0;
    FBP_REQUIRE(tv->d_items != nullptr, "fbp_forward(tiled): takes view has no work list");
    FastArgs a;
    fill_args(plan, tv, d_x, d_params, d_sub_static, a);
    a.pair_out = d_pair_out;
    a.cache = d_cache;
    a.order = tv->d_item_order_fwd;
    a.launch = tv->d_launch_fwd;
    if (plan->use_tc()) return fbp_tc_forward_launch(plan->fast, a, tv->n_items, stream);
    return dispatch(plan, false, a, tv->n_items, stream);
}

int64_t fbp_fast_backward_workspace(const fbp_plan* plan, const fbp_takes_view* tv) {
    return (int64_t)tv->n_items_active * plan->dev.P;
}

int fbp_fast_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                      const float* d_sub_static, const float* d_grow, float* d_grads, int accumulate, float* d_gpart,
                      const float* d_cache, cudaStream_t stream) {
    if (tv->m_active == 0) return 0;
    // FBP_BWD_DIRECT: one work item per active subdomain (the caller's assertion, checked as far as the host can see):
    // item i IS subdomain position i, the kernels write the rows of d_grads themselves and no reduction pass runs
    const bool direct = accumulate == FBP_BWD_DIRECT;
    FBP_REQUIRE(!(accumulate & FBP_BWD_DIRECT) || direct, "fbp_backward: FBP_BWD_DIRECT cannot be combined with other flags");
    FBP_REQUIRE(!direct || tv->n_items_active == tv->m_active, "fbp_backward: FBP_BWD_DIRECT needs one work item per active subdomain");
    if (direct) d_gpart = d_grads;
    FBP_REQUIRE(d_gpart != nullptr || tv->n_items_active == 0, "fbp_backward(tiled): null workspace");
    const bool run_kernels = !(accumulate & FBP_BWD_REDUCE_ONLY);
    const bool run_reduce = !(accumulate & FBP_BWD_NO_REDUCE) && !direct;
    accumulate &= FBP_BWD_ACCUMULATE;
    if (run_kernels && tv->n_items_active > 0) {
        FastArgs a;
        fill_args(plan, tv, d_x, d_params, d_sub_static, a);
        a.grow = d_grow;
        a.gpart = d_gpart;
        a.cache = const_cast<float*>(d_cache);
        a.order = tv->d_item_order_bwd;
        a.launch = tv->d_launch_bwd;
        if (plan->use_tc_bwd()) {
            // no activation cache on this path; the pointer doubles as the trace buffer of the phase-timing experiment
            // (FBP_TC_DEBUG & 64, tests/tools/bwd_phase_trace.py)
            const char* dbg_env = getenv("FBP_TC_DEBUG");
            if (!(dbg_env && (atoi(dbg_env) & 64))) a.cache = nullptr;
            if (int rc = fbp_tc_backward_launch(plan->fast, a, tv->n_items_active, stream)) return rc;
        } else if (int rc = dispatch(plan, true, a, tv->n_items_active, stream)) return rc;
    }
    if (!run_reduce) return 0;
    const int64_t total = (int64_t)tv->m_active * plan->dev.P;
    fast_grad_reduce_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(d_gpart, tv->d_sub_item_off, tv->m_active,
                                                                            plan->dev.P, d_grads, accumulate);
    FBP_LAUNCH_CHECK();
    return 0;
}
