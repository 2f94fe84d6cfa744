// MINIMAL STAND-IN for jaxlib's "xla/ffi/api/ffi.h" (jax / jaxlib are not installed in this image): just enough of the
// public binding API for fbpinns_b200/csrc/xla_ffi_shim.cc to be parsed and type-checked by a C++ compiler in the CPU
// test-suite (tests/test_host_logic.py::test_xla_ffi_shim_compiles_against_stub).  It is NOT the real header: it
// registers nothing and calls nothing.  Names and shapes follow the documented API (xla::ffi::Ffi::Bind().Ctx<>().Attr<>()
// .Arg<>().Ret<>(), Buffer<DataType>::typed_data(), Result<Buffer>::operator->, Error, XLA_FFI_DEFINE_HANDLER_SYMBOL).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace xla {
namespace ffi {

enum DataType { F32, S32 };
enum class ErrorCode { kOk, kInternal, kInvalidArgument };

class Error {
  public:
    Error() = default;
    Error(ErrorCode, std::string) {}
    static Error Success() { return Error(); }
};

template <DataType dtype>
struct NativeOf { using type = float; };
template <>
struct NativeOf<S32> { using type = int32_t; };

template <DataType dtype>
class Buffer {
  public:
    using T = typename NativeOf<dtype>::type;
    T* typed_data() const { return nullptr; }
    size_t element_count() const { return 0; }
};

template <typename T>
class Result {
  public:
    T* operator->() { return &v_; }
    T& operator*() { return v_; }

  private:
    T v_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

template <typename T>
struct PlatformStream {};

template <typename... Ts>
class Binding {
  public:
    template <typename T>
    Binding<Ts..., T> Ctx() const { return {}; }
    template <typename T>
    Binding<Ts..., T> Attr(const char*) const { return {}; }
    template <typename T>
    Binding<Ts..., T> Arg() const { return {}; }
    template <typename T>
    Binding<Ts..., T> Ret() const { return {}; }
};

struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an XLA_FFI_Handler symbol that decodes a call frame; here: take the address of the
// implementation so that its signature is instantiated and checked
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
    extern "C" const void* name##_stub_symbol() {           \
        (void)(binding);                                    \
        return reinterpret_cast<const void*>(&impl);        \
    }
