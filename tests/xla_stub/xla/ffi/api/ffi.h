// TEST INFRASTRUCTURE -- NOT XLA.  A stand-in for the small part of the public XLA typed-FFI C++ API
// (xla/ffi/api/ffi.h) that folax_b200/ffi/xla_ffi_shim.cc uses, so that the shim can at least be COMPILED in an image
// without jaxlib: it checks that the file is valid C++, that it matches the current C ABI of include/folax_b200.h,
// and -- through the type list the binding accumulates -- that every handler's parameter list agrees in number,
// order and type with its Ffi::Bind() chain (the usual way such shims are wrong).  It says nothing about run-time
// behaviour under XLA.  Names follow the reference's use of the real header
// (fol/loss_functions/ffi_functions/kr_small_displacement_element.cc:1-60, 293-337).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { F32, F64, S32, U8 };

class Error {
 public:
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string) { return Error(); }
  static Error Internal(std::string) { return Error(); }
};

template <class T>
class Span {
 public:
  std::size_t size() const { return 0; }
  const T* begin() const { return nullptr; }
  const T& operator[](std::size_t) const { return *begin(); }
};

class AnyBuffer {
 public:
  DataType element_type() const { return F64; }
  Span<const int64_t> dimensions() const { return {}; }
  void* untyped_data() const { return nullptr; }
};

template <DataType D>
struct NativeOf;
template <>
struct NativeOf<S32> { using type = int32_t; };
template <>
struct NativeOf<U8> { using type = uint8_t; };

template <DataType D>
class Buffer {
 public:
  Span<const int64_t> dimensions() const { return {}; }
  typename NativeOf<D>::type* typed_data() const { return nullptr; }
};

template <class T>
class Result {
 public:
  T* operator->() { return &value_; }
 private:
  T value_;
};

template <class T>
struct PlatformStream { using type = T; };

template <class... Ts>
struct Binding {
  template <class C>
  Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <class A>
  Binding<Ts..., A> Arg() const { return {}; }
  template <class R>
  Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <class A>
  Binding<Ts..., A> Attr(const char*) const { return {}; }
  template <class F>
  static constexpr bool Matches() { return std::is_invocable_r_v<Error, F, Ts...>; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an exported XLA_FFI_Handler; here: the signature check + an exported marker of that name
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                                   \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                                     \
                #impl " does not accept the argument list declared by its Ffi::Bind() chain");              \
  extern "C" int sym##_stub_marker = 1
