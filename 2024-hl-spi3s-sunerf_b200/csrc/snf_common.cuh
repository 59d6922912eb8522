// Shared device/host helpers for the sunerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "../../include/sunerf_b200.h"

namespace snf {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// launch accounting for bench.py's gpu_launches
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

// per-device one-time state (snf_mlp_bf16.cu): the SM count and the opt-in shared-memory attributes of the kernels
// belong to a device, and one process may drive several (nn.DataParallel-style rendering from a thread pool)
constexpr int kMaxDevices = 64;
struct DeviceState {
  std::once_flag once;
  int num_sms = 0;
  int status = 0;
  int reserve = 0;   // SMs the persistent field-network grids leave free (snf_config_reserve_sms), atomic access
};
extern DeviceState g_devices[kMaxDevices];

inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

#define SNF_CHECK_PTR(p) \
  if ((p) == nullptr) return SNF_E_ARG
#define SNF_CHECK_ALIGN(p, a) \
  if ((reinterpret_cast<uintptr_t>(p) & ((a)-1)) != 0) return SNF_E_ALIGN

// ---- exactly-rounded fp32 primitives: the sampling path must reproduce torch's op-by-op rounding
// (SURVEY.md H1); these intrinsics are never contracted into FMAs.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sum3(float a, float b, float c) { return fadd(fadd(a, b), c); }  // torch sum(-1) over 3

// ---- warp scans
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(kFull, v, d); }
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(kFull, v, d); }

__device__ __forceinline__ double warp_incl_sum(double v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double n = shfl_up_d(v, d);
    if (lane >= d) v += n;
  }
  return v;
}
__device__ __forceinline__ double warp_incl_prod(double v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double n = shfl_up_d(v, d);
    if (lane >= d) v *= n;
  }
  return v;
}
// suffix (reverse inclusive) sum: result[lane] = sum_{l >= lane} v[l]
__device__ __forceinline__ double warp_suffix_sum(double v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double n = shfl_down_d(v, d);
    if (lane + d < 32) v += n;
  }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}

}  // namespace snf
// one-time setup of the CURRENT device (thread-safe); returns 0 or an error code, optionally the SM count
int snf_device_setup(int *num_sms_out);
int snf_set_kernel_attributes();
namespace snf {

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace snf
