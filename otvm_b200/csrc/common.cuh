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

// ---- split-bf16 elements ("bf16 x P") --------------------------------------------------------------------
// A value is the SUM of P bf16 planes: plane 0 = bf16(v), plane 1 = bf16(v - plane 0), plane 2 = bf16(remainder):
// 8 / 16 / 24 significant bits for P = 1 / 2 / 3.  The planes of one tensor are identical NHWC views `ps` elements
// apart (all split tensors of a call live in one arena, include/otvm_b200.h "dtype").  The tensor cores multiply
// planes pairwise (conv_tc.cu), every other kernel reads the sum and writes the split through the accessors below.
template <int P> struct bx {};                 // element tag: bx<2>, bx<3>
template <int P> struct BxPtr {
  bf16* p; int64_t ps;
  __host__ __device__ BxPtr() : p(nullptr), ps(0) {}
  __host__ __device__ BxPtr(decltype(nullptr)) : p(nullptr), ps(0) {}
  __host__ __device__ BxPtr(bf16* p_, int64_t ps_) : p(p_), ps(ps_) {}
  __host__ __device__ BxPtr operator+(int64_t o) const { return BxPtr(p + o, ps); }
  __host__ __device__ explicit operator bool() const { return p != nullptr; }
};
// pointer type of an element type: raw pointers for float / bf16, BxPtr for split elements
template <typename T> struct El { typedef T* ptr; typedef const T* cptr; };
template <int P> struct El<bx<P>> { typedef BxPtr<P> ptr; typedef BxPtr<P> cptr; };
template <typename T> using ptr_t = typename El<T>::ptr;
template <typename T> using cptr_t = typename El<T>::cptr;
template <typename T> struct MkPtr {
  static __host__ __device__ T* make(void* p, int64_t) { return static_cast<T*>(p); }
  static __host__ __device__ const T* cmake(const void* p, int64_t) { return static_cast<const T*>(p); }
};
template <int P> struct MkPtr<bx<P>> {
  static __host__ __device__ BxPtr<P> make(void* p, int64_t ps) { return BxPtr<P>(static_cast<bf16*>(p), ps); }
  static __host__ __device__ BxPtr<P> cmake(const void* p, int64_t ps) { return BxPtr<P>(static_cast<bf16*>(const_cast<void*>(p)), ps); }
};
template <typename T> __host__ __device__ __forceinline__ ptr_t<T> mkptr(void* p, int64_t ps) { return MkPtr<T>::make(p, ps); }
template <typename T> __host__ __device__ __forceinline__ cptr_t<T> mkcptr(const void* p, int64_t ps) { return MkPtr<T>::cmake(p, ps); }

template <int P> __device__ __forceinline__ void load4(BxPtr<P> q, float (&v)[4]) {
  load4(q.p, v);
#pragma unroll
  for (int k = 1; k < P; ++k) {
    float t[4];
    load4(q.p + k * q.ps, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += t[j];
  }
}
template <int P> __device__ __forceinline__ void store4(BxPtr<P> q, const float (&v)[4]) {
  float r[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
  for (int k = 0; k < P; ++k) {
    __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]), b = __floats2bfloat162_rn(r[2], r[3]);
    uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(q.p + k * q.ps) = t;
    r[0] -= __low2float(a); r[1] -= __high2float(a); r[2] -= __low2float(b); r[3] -= __high2float(b);
  }
}
template <int P> __host__ __device__ __forceinline__ bool aligned4(BxPtr<P> q) {
  return (reinterpret_cast<uintptr_t>(q.p) & 7) == 0 && (q.ps & 3) == 0;
}
// one element: ld1(ptr, index) / st1(ptr, index, value)
__device__ __forceinline__ float ld1(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ld1(const bf16* p, int64_t i) { return __bfloat162float(p[i]); }
template <int P> __device__ __forceinline__ float ld1(BxPtr<P> q, int64_t i) {
  float v = __bfloat162float(q.p[i]);
#pragma unroll
  for (int k = 1; k < P; ++k) v += __bfloat162float(q.p[i + k * q.ps]);
  return v;
}
__device__ __forceinline__ void st1(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void st1(bf16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }
template <int P> __device__ __forceinline__ void st1(BxPtr<P> q, int64_t i, float v) {
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const bf16 h = __float2bfloat16_rn(v);
    q.p[i + k * q.ps] = h;
    v -= __bfloat162float(h);
  }
}
// 8 consecutive elements <-> float[8] (one 16-byte bf16 access or two 16-byte fp32 accesses)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 t; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

template <int P> __device__ __forceinline__ void load8(BxPtr<P> q, float (&v)[8]) {
  load8(q.p, v);
#pragma unroll
  for (int k = 1; k < P; ++k) {
    float t[8];
    load8(q.p + k * q.ps, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += t[j];
  }
}
template <int P> __device__ __forceinline__ void store8(BxPtr<P> q, const float (&v)[8]) {
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = v[j];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    uint4 t; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
      r[2 * i] -= __low2float(h[i]); r[2 * i + 1] -= __high2float(h[i]);
    }
    *reinterpret_cast<uint4*>(q.p + k * q.ps) = t;
  }
}

// the value a consumer reads back after a store of `v` (GroupNorm statistics are those of the STORED tensor):
// exact for fp32, bf16 rounding for one plane, identity (to ~2^-17) for split elements
template <typename T> __device__ __forceinline__ float stored(float v) { return v; }
template <> __device__ __forceinline__ float stored<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// dtype word of the C ABI: low byte = element format, bits 8.. = plane stride of split formats in 4 KB units
__host__ __device__ __forceinline__ int dtype_fmt(int dtype) { return dtype & 0xff; }
__host__ __device__ __forceinline__ int dtype_planes(int dtype) {
  const int f = dtype & 0xff;
  return f == OTVM_BF16X2 ? 2 : f == OTVM_BF16X3 ? 3 : 1;
}
// plane stride in ELEMENTS (bf16)
__host__ __device__ __forceinline__ int64_t dtype_plane_stride(int dtype) { return ((int64_t)((uint32_t)dtype >> 8) * 4096) / 2; }

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

// same, as thread-block clusters of (1, 1, cluster_z) CTAs (grid.z must be a multiple of cluster_z)
template <typename... Params, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                    unsigned cluster_z, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = cluster_z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

// number of SMs of the CURRENT device (cached per device: several devices may be used from one process)
int sm_count();
int current_device();                        // cudaGetDevice, -1 on failure

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember per (kernel, device) that it was done
template <auto Kernel>
inline cudaError_t ensure_dynamic_smem(int bytes) {
  static bool done[64] = {};
  const int dev = current_device();
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

}  // namespace otvm
