// TEST STAND-IN for jaxlib's `xla/ffi/api/ffi.h` (not the real header: jaxlib is not installable here).
// It declares the subset of the typed XLA FFI API that neural-tangents_b200/csrc/ntk_b200_ffi.cc uses, with the same
// names and call shapes -- Error / ErrorCode, DataType, AnyBuffer, Result<T>, PlatformStream<T>, the
// Ffi::Bind().Ctx<>().Attr<>().Arg<>().Ret<>() chain and XLA_FFI_DEFINE_HANDLER_SYMBOL -- so that the shim is
// type-checked and its handler body can be executed by tests/ffi_mock/harness.cc.  The binding chain is checked
// against the handler's signature at compile time (arguments arrive in binding order, as in the real API).
#pragma once
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

enum DataType { F32 = 11, F64 = 12 };

template <typename T>
class Span {
 public:
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T& back() const { return p_[n_ - 1]; }

 private:
  const T* p_;
  size_t n_;
};

class AnyBuffer {
 public:
  AnyBuffer(DataType t, void* data, std::vector<int64_t> dims) : t_(t), data_(data), dims_(std::move(dims)) {}
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
  DataType element_type() const { return t_; }
  void* untyped_data() const { return data_; }

 private:
  DataType t_;
  void* data_;
  std::vector<int64_t> dims_;
};

template <typename T>
class Result {
 public:
  explicit Result(T v) : v_(std::move(v)) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename T>
struct Decoded {
  using type = T;
};
template <typename T>
struct Decoded<PlatformStream<T>> {
  using type = T;
};
template <typename... Ts>
struct Binding {
  template <typename T>
  Binding<Ts..., typename Decoded<T>::type> Ctx() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename Fn>
  static constexpr bool matches() { return std::is_invocable_r<Error, Fn, Ts...>::value; }
};
}  // namespace internal

struct Ffi {
  static internal::Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                              \
  static_assert(decltype(binding)::template matches<decltype(&impl)>(),                                 \
                "handler signature does not match the Ffi::Bind() chain");                              \
  extern "C" const char* name##_mock_symbol() { return #name; }
