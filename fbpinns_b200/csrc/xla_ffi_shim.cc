// XLA FFI shim over the C ABI of libfbpinn_b200 (include/fbpinn_b200.h).
//
// NOT BUILT INTO THE LIBRARY IN THIS IMAGE: jax / jaxlib are not installed, so the real "xla/ffi/api/ffi.h" does not
// exist here and csrc/Makefile (which only globs fbp_*.cu) never compiles this file.  The CPU test-suite type-checks it
// against a minimal stand-in of that header (tests/tools/xla_ffi_stub, test_xla_ffi_shim_compiles_against_stub), so
// that it at least parses and its calls match include/fbpinn_b200.h.  It is the binding a maintainer of the reference
// builds where JAX exists:
//     g++ -O2 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../../include \
//         xla_ffi_shim.cc -L. -lfbpinn_b200 -o libfbpinn_b200_xla.so
// and registers with jax.ffi.register_ffi_target.  The handlers add no arithmetic: they
// unpack XLA buffers into the plain pointers of the C ABI and forward the stream XLA's GPU executor provides.
//
// Plans and takes views are created on the host side once per active-set change and passed as opaque 64-bit
// attributes (pointer values), which is how the reference's static arguments (model_fns / jmaps,
// fbpinns/trainers.py:285) cross the jit boundary.
#include <cstdint>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"
#include "fbpinn_b200.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(ffi::ErrorCode::kInternal, fbp_last_error());
}

// optional operands travel as zero-element buffers when absent
template <class B>
static auto* opt(B& b) { return b.element_count() ? b.typed_data() : nullptr; }

// ujets = reduce(forward(params))      (FBPINN_forward; with `affine` = jets of A and B of an affine constraining
// operator the constrained jets, otherwise the host applies constraining_fn).  `act_cache` is a RESULT of the forward
// call (s * fbp_plan_cache_per_pair floats; zero elements when the plan's reverse kernel recomputes, e.g. the tensor
// family) that the backward call receives as an operand.
static ffi::Error ForwardImpl(cudaStream_t stream, int64_t plan, int64_t takes, ffi::Buffer<ffi::F32> x,
                              ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> sub_static,
                              ffi::Buffer<ffi::F32> dsum, ffi::Buffer<ffi::F32> affine,
                              ffi::ResultBuffer<ffi::F32> pair_out, ffi::ResultBuffer<ffi::F32> act_cache,
                              ffi::ResultBuffer<ffi::F32> ujets) {
    auto* p = reinterpret_cast<const fbp_plan*>(plan);
    auto* tv = reinterpret_cast<const fbp_takes_view*>(takes);
    int rc = fbp_forward(p, tv, x.typed_data(), params.typed_data(), sub_static.typed_data(), pair_out->typed_data(),
                         nullptr, 0, opt(*act_cache), stream);
    if (rc == 0) rc = fbp_reduce_forward(p, tv, pair_out->typed_data(), dsum.typed_data(), opt(affine), ujets->typed_data(), stream);
    return status(rc);
}

// grads (m_active, P) from the cotangent of ujets      (the custom_vjp backward rule)
static ffi::Error BackwardImpl(cudaStream_t stream, int64_t plan, int64_t takes, ffi::Buffer<ffi::F32> x,
                               ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> sub_static,
                               ffi::Buffer<ffi::F32> dsum, ffi::Buffer<ffi::F32> affine, ffi::Buffer<ffi::F32> act_cache,
                               ffi::Buffer<ffi::F32> ujets_bar, ffi::ResultBuffer<ffi::F32> grow,
                               ffi::ResultBuffer<ffi::F32> gpart, ffi::ResultBuffer<ffi::F32> grads) {
    auto* p = reinterpret_cast<const fbp_plan*>(plan);
    auto* tv = reinterpret_cast<const fbp_takes_view*>(takes);
    int rc = fbp_reduce_backward(p, tv, ujets_bar.typed_data(), dsum.typed_data(), opt(affine), grow->typed_data(), stream);
    if (rc == 0)
        rc = fbp_backward(p, tv, x.typed_data(), params.typed_data(), sub_static.typed_data(), grow->typed_data(),
                          grads->typed_data(), /*accumulate=*/0, gpart->typed_data(), nullptr, 0, opt(act_cache), stream);
    return status(rc);
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(FbpForward, ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("takes")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params
                                  .Arg<ffi::Buffer<ffi::F32>>()   // sub_static
                                  .Arg<ffi::Buffer<ffi::F32>>()   // dsum
                                  .Arg<ffi::Buffer<ffi::F32>>()   // affine (n, 2C) or zero elements
                                  .Ret<ffi::Buffer<ffi::F32>>()   // pair_out
                                  .Ret<ffi::Buffer<ffi::F32>>()   // act_cache (s * cache_per_pair, possibly zero elements)
                                  .Ret<ffi::Buffer<ffi::F32>>()); // ujets

XLA_FFI_DEFINE_HANDLER_SYMBOL(FbpBackward, BackwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("takes")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params
                                  .Arg<ffi::Buffer<ffi::F32>>()   // sub_static
                                  .Arg<ffi::Buffer<ffi::F32>>()   // dsum
                                  .Arg<ffi::Buffer<ffi::F32>>()   // affine
                                  .Arg<ffi::Buffer<ffi::F32>>()   // act_cache
                                  .Arg<ffi::Buffer<ffi::F32>>()   // ujets_bar
                                  .Ret<ffi::Buffer<ffi::F32>>()   // grow
                                  .Ret<ffi::Buffer<ffi::F32>>()   // gpart
                                  .Ret<ffi::Buffer<ffi::F32>>()); // grads
