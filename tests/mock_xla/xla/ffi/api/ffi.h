// Minimal stand-in for jaxlib's xla/ffi/api/ffi.h: just enough of the typed-FFI surface (Buffer, Result, Error,
// RemainingArgs / RemainingRets, the Bind() builder and XLA_FFI_DEFINE_HANDLER_SYMBOL) for
// `g++ -fsyntax-only` to type-check tensorf-jax_b200/jax_ffi/xla_ffi_shim.cc in an image without JAX:
// every handler must be invocable with exactly the (context, argument, result, attribute) types its binding
// declares, in order, and must return ffi::Error.  TEST INFRASTRUCTURE ONLY - never shipped or linked.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};
template <typename T>
class ErrorOr {
 public:
  bool has_value() const { return ok_; }
  T& value() { return v_; }
  T& operator*() { return v_; }
  T* operator->() { return &v_; }
  Error error() const { return Error(); }

 private:
  bool ok_ = true;
  T v_{};
};

enum DataType { F32, U32, U8, S32, S64 };
template <DataType>
struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<S64> { using type = int64_t; };

template <typename T>
class Span {
 public:
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }

 private:
  const T* p_ = nullptr;
  size_t n_ = 0;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeType<dtype>::type;
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }

 private:
  T* data_ = nullptr;
};
class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  template <typename T> T* typed_data() const { return nullptr; }
  Span<int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_{};
};

class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<T> get(size_t) const { return {}; }
};
class RemainingRets {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<Result<T>> get(size_t) const { return {}; }
};

template <typename T>
struct PlatformStream {};

template <typename... Ts>
class Binding {
 public:
  template <typename C> auto Ctx() const { return CtxImpl(static_cast<C*>(nullptr)); }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingArgs> RemainingArgs() const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingRets> RemainingRets() const { return {}; }
  template <typename Fn>
  int To(Fn) const {
    static_assert(std::is_invocable_r_v<Error, Fn, Ts...>, "handler signature does not match its binding");
    return 0;
  }

 private:
  template <typename S> Binding<Ts..., S> CtxImpl(PlatformStream<S>*) const { return {}; }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding) \
  extern "C" int symbol() { return (binding).To(fn); }
