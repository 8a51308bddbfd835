// C-ABI entry points that dispatch between the FFMA (strict fp32 / fallback) and tcgen05 kernels.
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace otvm {

static thread_local char g_last_error[256] = "";

void set_cuda_error(cudaError_t e) {
  strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
  cudaGetLastError();                          // clear the sticky-less error state
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) { const char* e = getenv("OTVM_PDL"); g_pdl = (e && e[0] == '0') ? 0 : 1; }
  return g_pdl == 1;
}

int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {
  static int n[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return 148;
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
      cudaGetLastError();
      v = 148;
    }
    n[dev] = v;
  }
  return n[dev];
}

int conv2d_simt(const otvm_conv_params* p, cudaStream_t s);
int conv2d_tc(const otvm_conv_params* p, cudaStream_t s, bool dry_run = false);
bool conv2d_tc_supported(const otvm_conv_params* p);
int memory_read_simt(const otvm_read_params* p, cudaStream_t s);
int memory_read_tc(const otvm_read_params* p, cudaStream_t s);
bool memory_read_tc_supported(const otvm_read_params* p);
int64_t memory_read_tc_workspace(int M, int HW, int De, int Do);
int read_max_splits(int M, int HW, int Do, int rows_per_cta, int cols_per_cta, int keys_per_block);

}  // namespace otvm

using namespace otvm;

extern "C" int otvm_version(void) { return OTVM_ABI_VERSION; }

extern "C" const char* otvm_strerror(int code) {
  switch (code) {
    case OTVM_OK: return "ok";
    case OTVM_ERR_ARG: return "invalid argument";
    case OTVM_ERR_CUDA: return "CUDA error (see otvm_last_cuda_error)";
    case OTVM_ERR_UNSUPPORTED: return "unsupported shape or dtype for this kernel";
    case OTVM_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

extern "C" const char* otvm_last_cuda_error(void) { return g_last_error; }

extern "C" int64_t otvm_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" void otvm_set_pdl(int enabled) { otvm::g_pdl = enabled ? 1 : 0; }

extern "C" int otvm_zero_async(void* ptr, int64_t bytes, void* stream) {
  if (!ptr || bytes < 0) return OTVM_ERR_ARG;
  OTVM_CUDA_CHECK(cudaMemsetAsync(ptr, 0, (size_t)bytes, static_cast<cudaStream_t>(stream)));
  return OTVM_OK;
}

extern "C" int otvm_device_is_sm100(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return 0;
  return major == 10;
}

static int conv_check(const otvm_conv_params* p) {
  if (!p || !p->in || !p->weight || !p->out) return OTVM_ERR_ARG;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Cin <= 0 || p->Cout <= 0 || p->KH <= 0 || p->KW <= 0 ||
      p->stride <= 0 || p->dil <= 0 || p->pad < 0 || p->in_ld < p->Cin)
    return OTVM_ERR_ARG;
  { const int f = dtype_fmt(p->dtype); if (f < OTVM_F32 || f > OTVM_BF16X3) return OTVM_ERR_ARG; if (f >= OTVM_BF16X2 && (dtype_plane_stride(p->dtype) <= 0 || p->w_plane_stride <= 0)) return OTVM_ERR_ARG; }
  return OTVM_OK;
}

extern "C" int otvm_conv2d_uses_tensor_cores(const otvm_conv_params* p) {
  return conv_check(p) == OTVM_OK && conv2d_tc_supported(p) ? 1 : 0;
}

extern "C" int otvm_conv2d_can_fuse_gn(const otvm_conv_params* p) {
  if (conv_check(p) != OTVM_OK || !p->gn_gamma || !conv2d_tc_supported(p)) return 0;
  return conv2d_tc(p, nullptr, /*dry_run=*/true) == OTVM_OK ? 1 : 0;
}

extern "C" int otvm_conv2d(const otvm_conv_params* p, void* stream) {
  int rc = conv_check(p);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (conv2d_tc_supported(p)) return conv2d_tc(p, s);
  if (p->gn_gamma) return OTVM_ERR_UNSUPPORTED;            // only the tcgen05 kernel normalises in its epilogue
  return conv2d_simt(p, s);
}

extern "C" int64_t otvm_memory_read_workspace(int32_t M, int32_t HW, int32_t De, int32_t Do, int32_t dtype) {
  (void)dtype; (void)De;
  // partial O (fp32) + (m, l) per split; both kernels use at most 64 splits and never more than M/64 blocks
  int ns_simt = read_max_splits(M, HW, Do, 64, 128, 64);
  int64_t ns = ns_simt;
  int64_t tc = memory_read_tc_workspace(M, HW, De, Do);
  int64_t simt = ns * HW * ((int64_t)Do + 2) * (int64_t)sizeof(float);
  return simt > tc ? simt : tc;
}

extern "C" int otvm_memory_read(const otvm_read_params* p, void* stream) {
  if (!p || !p->keys || !p->vals || !p->query || !p->out || !p->workspace) return OTVM_ERR_ARG;
  if (p->M <= 0 || p->HW <= 0 || p->De <= 0 || p->Do <= 0) return OTVM_ERR_ARG;
  if (p->workspace_bytes < otvm_memory_read_workspace(p->M, p->HW, p->De, p->Do, p->dtype)) return OTVM_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!p->force_simt && memory_read_tc_supported(p)) return memory_read_tc(p, s);
  return memory_read_simt(p, s);
}
