// NOT the XLA header.  A stand-in with the names csrc/ffi/xla_ffi.cc uses from `xla/ffi/api/ffi.h`
// (jax.ffi.include_dir(), absent from this image), so that the shim can at least be TYPE-CHECKED here:
// XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that the implementation is callable with exactly the context,
// attribute, operand and result types its binding declares, in that order -- the class of bug (an operand
// missing from the binding) that XLA would only report at call time.  Test infrastructure only
// (tests/test_ffi_shim_typecheck.py); nothing built from it ships.
#ifndef FDM_TESTS_XLA_FFI_MOCK_H
#define FDM_TESTS_XLA_FFI_MOCK_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { F64, U8, S32 };
template <DataType T> struct NativeOf;
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };

struct Dims {
  const int64_t* p = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  int64_t operator[](size_t i) const { return p[i]; }
};

template <DataType T>
class Buffer {
 public:
  using Native = typename NativeOf<T>::type;
  Native* typed_data() const { return data_; }
  Dims dimensions() const { return dims_; }
  size_t element_count() const {
    size_t c = 1;
    for (size_t i = 0; i < dims_.n; ++i) c *= (size_t)dims_.p[i];
    return c;
  }
  size_t size_bytes() const { return element_count() * sizeof(Native); }
  Native* data_ = nullptr;
  Dims dims_;
};

template <class T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }
  T v_;
};

enum class ErrorCode { kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string m) : ok_(false), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return ok_; }
 private:
  bool ok_ = true;
  std::string msg_;
};

template <class T> struct PlatformStream {};

template <class... Ts>
struct Binding {
  template <class C> auto Ctx() const { return CtxOf(static_cast<C*>(nullptr)); }
  template <class T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <class T> Binding<Ts..., T> Arg() const { return {}; }
  template <class T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <class F> static constexpr bool Matches() { return std::is_invocable_r<Error, F, Ts...>::value; }
 private:
  template <class S> Binding<Ts..., S> CtxOf(PlatformStream<S>*) const { return {}; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                              \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                                \
                #sym ": the implementation's parameters do not match the binding (context, attributes, " \
                     "operands, results -- in that order)");                                             \
  extern "C" XLA_FFI_Error* sym(XLA_FFI_CallFrame*) { return nullptr; }
#endif
