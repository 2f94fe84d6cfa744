// CPU emulation of fbpinns_b200/csrc/fbp_generic.cu (plain-FCN generic kernels): the kernels' own source compiled as C++
// (fbp_host_emu.h), one pair per "thread".  `plan` comes from the real library; every other pointer is a host array.
#define FBP_HOST_EMU 1
#include "../../fbpinns_b200/csrc/fbp_generic.cu"

extern "C" int emu_generic_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* x, const float* params,
                                   const float* sub_static, float* pair_out, float* scratch) {
    blockDim.x = 1;
    for (int64_t t = 0; t < tv->s; ++t) {
        blockIdx.x = (unsigned)t; threadIdx.x = 0;
        generic_forward_kernel(plan->dev, *tv, x, params, sub_static, pair_out, scratch, 0, tv->s, tv->s);
    }
    return 0;
}

extern "C" int emu_generic_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* x, const float* params,
                                    const float* sub_static, const float* grow, float* grads, float* scratch) {
    blockDim.x = 1;
    for (int64_t t = 0; t < tv->s_active; ++t) {
        blockIdx.x = (unsigned)t; threadIdx.x = 0;
        generic_backward_kernel(plan->dev, *tv, x, params, sub_static, grow, grads, scratch, 0, tv->s_active, tv->s_active);
    }
    return 0;
}
