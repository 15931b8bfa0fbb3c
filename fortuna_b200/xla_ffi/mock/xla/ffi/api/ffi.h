// MOCK of jaxlib's xla/ffi/api/ffi.h - just enough of its public surface (names, member functions, the Bind() builder)
// for `g++ -fsyntax-only` to type-check fb200_xla_ffi.cc where jaxlib cannot be installed (scripts/check_xla_ffi_syntax.sh).
// It is NOT the real header and implements nothing: Binding::To() static_asserts that the handler is callable with the
// bound argument list, which is the check that matters (a handler whose signature drifts from its binding fails to
// compile here exactly as it would against jaxlib).  Written from the documented API of XLA's FFI (jax 0.4.30+).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla::ffi {

enum class DataType : uint8_t { INVALID, PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
inline constexpr DataType PRED = DataType::PRED, S8 = DataType::S8, S32 = DataType::S32, S64 = DataType::S64,
                          U8 = DataType::U8, U32 = DataType::U32, U64 = DataType::U64, F32 = DataType::F32,
                          F64 = DataType::F64, BF16 = DataType::BF16;

enum class ErrorCode : uint8_t { kOk, kCancelled, kUnknown, kInvalidArgument, kInternal, kUnimplemented };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class Span {
 public:
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T* begin() const { return d_; }
  const T* end() const { return d_ + n_; }

 private:
  const T* d_;
  size_t n_;
};

template <DataType dt> struct NativeTypeOf { using type = void; };
template <> struct NativeTypeOf<DataType::PRED> { using type = bool; };
template <> struct NativeTypeOf<DataType::S8> { using type = int8_t; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };
template <> struct NativeTypeOf<DataType::U8> { using type = uint8_t; };
template <> struct NativeTypeOf<DataType::U32> { using type = uint32_t; };
template <> struct NativeTypeOf<DataType::U64> { using type = uint64_t; };
template <> struct NativeTypeOf<DataType::F32> { using type = float; };
template <> struct NativeTypeOf<DataType::F64> { using type = double; };
template <> struct NativeTypeOf<DataType::BF16> { using type = uint16_t; };
template <DataType dt> using NativeType = typename NativeTypeOf<dt>::type;

class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  DataType element_type() const { return DataType::INVALID; }
  Span<int64_t> dimensions() const { return Span<int64_t>(nullptr, 0); }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <DataType dt>
class Buffer {
 public:
  NativeType<dt>* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  Span<int64_t> dimensions() const { return Span<int64_t>(nullptr, 0); }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

template <typename Stream> struct PlatformStream {};

namespace internal {
template <typename T> struct CtxType;
template <typename S> struct CtxType<PlatformStream<S>> { using type = S; };
}  // namespace internal

struct Handler {};

template <typename... Ts>
class Binding {
 public:
  template <typename C> Binding<Ts..., typename internal::CtxType<C>::type> Ctx() { return {}; }
  template <typename T> Binding<Ts..., T> Arg() { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) { return {}; }
  template <typename F>
  Handler To(F&&) {
    static_assert(std::is_invocable_r_v<Error, F, Ts...>, "handler signature does not match its XLA FFI binding");
    return {};
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)        \
  extern "C" void* name(void*) {                                  \
    static ::xla::ffi::Handler handler = (binding).To(impl);      \
    (void)handler;                                                \
    return nullptr;                                               \
  }
