// CPU emulation of the streaming kernels of fbpinns_b200/csrc/fbp_reduce.cu (window sums, row sums, segment-sum + quotient
// rule forward and transpose, Adam): their own source compiled as C++ (fbp_host_emu.h), one row / point / element per
// "thread".  `plan` comes from the real library; every other pointer is a host array.
#define FBP_HOST_EMU 1
#include "../../fbpinns_b200/csrc/fbp_reduce.cu"

template <class F>
static void for_threads(int64_t n, F f) {
    blockDim.x = 1; gridDim.x = (unsigned)n;
    for (int64_t t = 0; t < n; ++t) { blockIdx.x = (unsigned)t; threadIdx.x = 0; f(); }
}

extern "C" int emu_window_sums(const fbp_plan* plan, const fbp_takes_view* tv, const float* x, const float* sub_static, float* dsum) {
    for_threads(tv->q, [&] { window_sums_kernel(plan->dev, *tv, x, sub_static, dsum); });
    return 0;
}
extern "C" int emu_row_sums(const fbp_plan* plan, const fbp_takes_view* tv, const float* pair_out, float* nsum) {
    for_threads(tv->q, [&] { row_sums_kernel(plan->dev, *tv, pair_out, nsum); });
    return 0;
}
extern "C" int emu_reduce_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* pair_out_or_rows, int from_rows,
                                  const float* dsum, const float* aff, float* ujets) {
    if (from_rows) for_threads(tv->n, [&] { reduce_forward_kernel<true>(plan->dev, *tv, pair_out_or_rows, dsum, aff, ujets, nullptr); });
    else for_threads(tv->n, [&] { reduce_forward_kernel<false>(plan->dev, *tv, pair_out_or_rows, dsum, aff, ujets, nullptr); });
    return 0;
}
extern "C" int emu_reduce_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* ubar, const float* dsum,
                                   const float* aff, float* grow) {
    for_threads(tv->q, [&] { reduce_backward_kernel<0>(plan->dev, *tv, ubar, dsum, aff, grow); });
    return 0;
}
extern "C" int emu_adam(float* params, float* mu, float* nu, const float* grads, const int32_t* row_ids, int64_t n_rows,
                        int64_t row_len, const int32_t* count, float lr, float b1, float b2, float eps, float eps_root) {
    for_threads(n_rows * row_len, [&] { adam_kernel(params, mu, nu, grads, row_ids, n_rows, row_len, count, lr, b1, b2, eps, eps_root); });
    return 0;
}

// The component-count instances (CT = C, ud = 1: everything in registers on the GPU) against the general kernels above.
#define EMU_CT_SWITCH(ct, CALL)                                                                          \
    switch (ct) {                                                                                        \
        case 1: { constexpr int CT = 1; CALL; } break;                                                   \
        case 2: { constexpr int CT = 2; CALL; } break;                                                   \
        case 3: { constexpr int CT = 3; CALL; } break;                                                   \
        case 4: { constexpr int CT = 4; CALL; } break;                                                   \
        case 5: { constexpr int CT = 5; CALL; } break;                                                   \
        case 6: { constexpr int CT = 6; CALL; } break;                                                   \
        case 7: { constexpr int CT = 7; CALL; } break;                                                   \
        default: return 1;                                                                               \
    }
extern "C" int emu_reduce_forward_ct(const fbp_plan* plan, const fbp_takes_view* tv, const float* pair_out_or_rows, int from_rows,
                                     const float* dsum, const float* aff, float* ujets, int ct) {
    if (ct != plan->dev.C || plan->dev.ud != 1) return 2;
    if (from_rows) {
        EMU_CT_SWITCH(ct, for_threads(tv->n, [&] { reduce_forward_kernel<true, CT>(plan->dev, *tv, pair_out_or_rows, dsum, aff, ujets, nullptr); }))
    } else {
        EMU_CT_SWITCH(ct, for_threads(tv->n, [&] { reduce_forward_kernel<false, CT>(plan->dev, *tv, pair_out_or_rows, dsum, aff, ujets, nullptr); }))
    }
    return 0;
}
extern "C" int emu_reduce_backward_ct(const fbp_plan* plan, const fbp_takes_view* tv, const float* ubar, const float* dsum,
                                      const float* aff, float* grow, int ct) {
    if (ct != plan->dev.C || plan->dev.ud != 1) return 2;
    EMU_CT_SWITCH(ct, for_threads(tv->q, [&] { reduce_backward_kernel<CT>(plan->dev, *tv, ubar, dsum, aff, grow); }))
    return 0;
}

