// Hardware probe (dev / test only): does tcgen05.mma accept a SWIZZLE_128B K-major operand whose start address is
// offset by a whole number of 128-byte rows (not a multiple of the 1024-byte swizzle atom)?  A [136 x 64] bf16 tile
// is loaded once by TMA; for every row shift s = 0..8 the MMA computes D_s = A[s : s+128, :] . W^T from the SAME
// shared-memory copy, with the descriptor's base_offset field either 0 or (start >> 7) & 7.  The halo-reuse
// convolution (one input row box feeding the three horizontal filter taps) depends on the answer.
#include "tc_common.cuh"   // -I otvm_b200/csrc (scripts/umma_probe.py)

namespace otvm {
using namespace tc;

__global__ void __launch_bounds__(128) umma_shift_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmW,
                                                               float* __restrict__ out, int policy) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = smem;                    // 136 rows x 128 B = 17408 B (17 KB)
  uint8_t* sW = smem + 18 * 1024;        // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 28 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 136 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sW, &tmW, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  tcgen05_after_sync();
  constexpr uint32_t idesc = make_idesc_bf16(128, 64);
  for (int s = 0; s <= 8; ++s) {
    if (threadIdx.x == 0) {
      const uint32_t a_addr = base + (uint32_t)s * 128u;
      uint64_t ad = make_smem_desc(a_addr, 1024, 2);
      if (policy == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;
      const uint64_t wd = make_smem_desc(base + 18 * 1024, 1024, 2);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc, k != 0);
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], s & 1);
    tcgen05_after_sync();
    uint32_t r[32];
    float* o = out + ((size_t)s * 128 + warp * 32 + lane) * 64;
    for (int c = 0; c < 64; c += 32) {
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
      tmem_wait_ld();
      for (int i = 0; i < 32; ++i) o[c + i] = __uint_as_float(r[i]);
    }
    tcgen05_before_sync();
    __syncthreads();
    tcgen05_after_sync();
  }
  if (warp == 0) tmem_dealloc<64>(tmem_base);
}

}  // namespace otvm

using namespace otvm;

// A: [136][64] bf16 row-major, W: [64][64] bf16 row-major (both device), out: [9][128][64] fp32
extern "C" __attribute__((visibility("default"))) int otvm_debug_umma_shift_probe(const void* A, const void* W, float* out,
                                                                                  int policy, void* stream) {
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[2] = {64, 136}; uint64_t str[1] = {128}; uint32_t box[2] = {64, 136};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc;
  }
  {
    uint64_t dims[2] = {64, 64}; uint64_t str[1] = {128}; uint32_t box[2] = {64, 64};
    int rc = make_tmap_bf16(&tmW, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OTVM_CUDA_CHECK(cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  umma_shift_probe_kernel<<<1, 128, 30 * 1024, s>>>(tmA, tmW, out, policy);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}
