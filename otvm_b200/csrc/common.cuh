// Shared device/host helpers for the otvm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/otvm_b200.h"

namespace otvm {

typedef __nv_bfloat16 bf16;

void set_cuda_error(cudaError_t e);          // abi.cu: remembers the string for otvm_last_cuda_error()

#define OTVM_CUDA_CHECK(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) { ::otvm::set_cuda_error(_e); return OTVM_ERR_CUDA; } \
  } while (0)

void count_launch();                         // abi.cu: kernels launched by this library (otvm_launch_count)
#define OTVM_LAUNCH_CHECK() do { ::otvm::count_launch(); OTVM_CUDA_CHECK(cudaGetLastError()); } while (0)

template <typename T> struct DT;
template <> struct DT<float> { static constexpr int code = OTVM_F32; };
template <> struct DT<bf16>  { static constexpr int code = OTVM_BF16; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float[4]
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
template <typename T> __host__ __device__ __forceinline__ bool aligned4(const T* p) {
  return (reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == OTVM_ACT_RELU) return fmaxf(v, 0.f);
  if (act == OTVM_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// Every kernel of the frame is launched with programmatic stream serialisation: the NEXT kernel's CTAs may become
// resident (and run their prologue: barrier init, TMEM allocation, tensor-map prefetch, constant loads) while this
// kernel drains.  Contract for every kernel in this library:
//   * pdl_wait() before the first access to global memory that another kernel may have written / may still read;
//   * pdl_trigger() only once the CTA holds every resource it will ever need (tcgen05 kernels: after TMEM
//     allocation, otherwise a dependent CTA could grab the columns a not-yet-allocated primary CTA is waiting for);
//   * no kernel exits without having executed pdl_wait() (completion order stays transitive along the stream).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }

bool pdl_enabled();                          // abi.cu: OTVM_PDL env (default on), otvm_set_pdl()

template <typename... Params, typename... Args>
inline cudaError_t launch_k(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

// number of SMs of the current device (cached)
int sm_count();

}  // namespace otvm
