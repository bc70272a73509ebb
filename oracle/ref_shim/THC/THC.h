// ORACLE build glue (test infrastructure, not product code).
// The reference's csrc/*.cu include <THC/THC.h>, which PyTorch removed.  This header is put on the
// include path under that name and force-included so the UNMODIFIED reference sources build against
// torch 2.11: it supplies the handful of macros / one overload they still expect.
#pragma once
#include <ATen/ATen.h>
#include <c10/cuda/CUDAException.h>
#include <c10/util/Exception.h>
#define THCudaCheck(x) C10_CUDA_CHECK(x)
#define THArgCheck(cond, argn, ...) TORCH_CHECK(cond, __VA_ARGS__)
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b), #a " != " #b)
#define CHECK_GT(a, b) TORCH_CHECK((a) > (b), #a " <= " #b)
#define CHECK_GE(a, b) TORCH_CHECK((a) >= (b), #a " < " #b)
#endif
namespace detail {  // AT_DISPATCH_* resolves ::detail::scalar_type(x.type())
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
