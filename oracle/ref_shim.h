// Force-included (-include) when compiling the UNMODIFIED reference CUDA extension from
// /root/reference against torch >= 2.x.  TEST INFRASTRUCTURE ONLY.
//
// The reference passes `value.type()` (an at::DeprecatedTypeProperties) to AT_DISPATCH_ALL_TYPES
// (box_attn.cu:54,117; instance_attn.cu:62,136).  Current ATen's dispatch macro resolves the dtype
// through ::detail::scalar_type(<arg>), whose DeprecatedTypeProperties overload was removed.
// Re-adding that one overload here lets the reference sources compile byte-for-byte unchanged.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
