// HBM-bound kernels of the OTVM frame: GroupNorm(32,C) statistics/apply (+activation, +residual), bilinear
// resampling, max/avg pooling, layout conversion.  All NHWC, 4 channels (16 B fp32 / 8 B bf16) per thread so
// a warp touches 128..512 contiguous bytes; grids are sized from the SM count (grid-stride loops).
#include "common.cuh"

namespace otvm {

static inline int launched() { OTVM_LAUNCH_CHECK(); return OTVM_OK; }

static inline int grid_for(int64_t work_items, int block) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// ------------------------------------------------------------------------------------------------------
// GroupNorm statistics: per (n, group) sum and sum of squares in fp64 (fp32 partials per thread/block).
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(cptr_t<T> x, int64_t ld, int64_t HW, int C,
                                                       double* __restrict__ stats) {
  pdl_sync();                                  // PDL contract (common.cuh)
  __shared__ float part[32][2];
  const int n = blockIdx.y;
  const int c4n = C >> 2;                        // channel quads per pixel
  const int64_t total = HW * c4n;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;       // multiple of c4n (host guarantees)
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cq = (int)(idx % c4n);
  const int cg = C >> 5;                         // channels per group
  if (threadIdx.x < 32) { part[threadIdx.x][0] = 0.f; part[threadIdx.x][1] = 0.f; }
  __syncthreads();
  float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
  cptr_t<T> base = x + ((int64_t)n * HW * ld + cq * 4);
  for (; idx < total; idx += nthreads) {
    int64_t pix = idx / c4n;
    float v[4];
    load4(base + pix * ld, v);
    s1a += v[0] + v[1]; s2a += v[0] * v[0] + v[1] * v[1];
    s1b += v[2] + v[3]; s2b += v[2] * v[2] + v[3] * v[3];
  }
  const int ga = (cq * 4) / cg, gb = (cq * 4 + 2) / cg;
  atomicAdd(&part[ga][0], s1a); atomicAdd(&part[ga][1], s2a);
  atomicAdd(&part[gb][0], s1b); atomicAdd(&part[gb][1], s2b);
  __syncthreads();
  if (threadIdx.x < 32) {
    atomicAdd(&stats[(n * 32 + threadIdx.x) * 2 + 0], (double)part[threadIdx.x][0]);
    atomicAdd(&stats[(n * 32 + threadIdx.x) * 2 + 1], (double)part[threadIdx.x][1]);
  }
}

// GroupNorm apply: every thread owns 8 fixed channels (scale/shift live in registers) and walks the pixels with
// 16-byte accesses, 4 pixels in flight per thread; a warp touches 512 contiguous bytes of one or more pixels.
template <typename T, bool RES>
__global__ void __launch_bounds__(256) gn_apply_kernel(cptr_t<T> x, int64_t ld, int64_t HW, int C,
                                                       const double* __restrict__ stats,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       cptr_t<T> res, int64_t res_ld, int act,
                                                       ptr_t<T> out, int64_t out_ld, double inv_cnt) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int n = blockIdx.y;
  const int tpp = C >> 3;                         // threads per pixel (divides 256)
  const int cv = (threadIdx.x % tpp) * 8;
  const int ppb = 256 / tpp;                      // pixels per block pass
  const int cg = C >> 5;
  __shared__ float g_mean[32], g_rstd[32];
  if (threadIdx.x < 32) {
    // fp64 only where E[x^2] - E[x]^2 cancels (two multiplies and one fma: fp64 division / square root run at 1/64
    // rate here and sat in every block's prologue); the reciprocal square root of the fp32 variance is exact enough
    const double mean = stats[(n * 32 + threadIdx.x) * 2] * inv_cnt;
    const double var = fma(-mean, mean, stats[(n * 32 + threadIdx.x) * 2 + 1] * inv_cnt);
    g_mean[threadIdx.x] = (float)mean;
    g_rstd[threadIdx.x] = 1.0f / sqrtf(fmaxf((float)var, 0.f) + eps);
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (cv + j) / cg;
    sc[j] = g_rstd[g] * gamma[cv + j];
    sh[j] = beta[cv + j] - g_mean[g] * sc[j];
  }
  const float slope = act == OTVM_ACT_NONE ? 1.f : act == OTVM_ACT_RELU ? 0.f : 0.01f;
  cptr_t<T> xb = x + ((int64_t)n * HW * ld + cv);
  cptr_t<T> rb = res + (RES ? (int64_t)n * HW * res_ld + cv : 0);
  ptr_t<T> ob = out + ((int64_t)n * HW * out_ld + cv);
  const int64_t stride = (int64_t)gridDim.x * ppb;
  for (int64_t p0 = (int64_t)blockIdx.x * ppb + threadIdx.x / tpp; p0 < HW; p0 += 4 * stride) {
    float v[4][8], r[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t p = p0 + u * stride;
      if (p < HW) { load8(xb + p * ld, v[u]); if (RES) load8(rb + p * res_ld, r[u]); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t p = p0 + u * stride;
      if (p < HW) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = fmaf(v[u][j], sc[j], sh[j]);
          if (RES) y += r[u][j];
          v[u][j] = fmaxf(y, 0.f) + slope * fminf(y, 0.f);
        }
        store8(ob + p * out_ld, v[u]);
      }
    }
  }
}

template <typename T>
static int gn_stats_t(const void* x, int64_t ld, int N, int HW, int C, double* stats, int64_t ps, cudaStream_t s) {
  OTVM_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 64 * N, s));
  int c4n = C / 4;
  int64_t total = (int64_t)HW * c4n;
  int64_t lcm = c4n > 256 ? c4n : 256;           // c4n is a power of two times {1}: 16..512
  if (lcm % 256 != 0 || lcm % c4n != 0) return OTVM_ERR_UNSUPPORTED;
  int64_t want = (int64_t)sm_count() * 4 * 256;
  int64_t threads = ((total < want ? total : want) + lcm - 1) / lcm * lcm;
  dim3 grid((unsigned)(threads / 256), N);
  launch_k(gn_stats_kernel<T>, grid, 256, 0, s, mkcptr<T>(x, ps), ld, HW, C, stats);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

template <typename T>
static int gn_apply_t(const void* x, int64_t ld, int N, int HW, int C, const double* stats, const float* gamma,
                      const float* beta, float eps, const void* res, int64_t res_ld, int act, void* out,
                      int64_t out_ld, int64_t ps, cudaStream_t s) {
  const int tpp = C / 8, ppb = 256 / tpp;
  const double inv_cnt = 1.0 / ((double)HW * (double)(C / 32));      // elements per (sample, group)
  int64_t blocks = ((int64_t)HW + (int64_t)ppb * 4 - 1) / ((int64_t)ppb * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, N);
  if (res)
    launch_k(gn_apply_kernel<T, true>, grid, 256, 0, s, mkcptr<T>(x, ps), ld, HW, C, stats, gamma, beta, eps,
                                                  mkcptr<T>(res, ps), res_ld, act, mkptr<T>(out, ps), out_ld, inv_cnt);
  else
    launch_k(gn_apply_kernel<T, false>, grid, 256, 0, s, mkcptr<T>(x, ps), ld, HW, C, stats, gamma, beta, eps,
                                                   nullptr, 0, act, mkptr<T>(out, ps), out_ld, inv_cnt);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

// ------------------------------------------------------------------------------------------------------
// bilinear resize, align_corners=False (ATen upsample_bilinear2d index rule), optional fused add
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256) upsample_kernel(cptr_t<T> in, int64_t in_ld, int Hi, int Wi,
                                                       int C, int Ho, int Wo, float sy, float sx,
                                                       cptr_t<T> add, int64_t add_ld,
                                                       ptr_t<T> out, int64_t out_ld,
                                                       ptr_t<T> out_relu, int64_t out_relu_ld, int N) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int c4n = C >> 2;
  const int64_t total = (int64_t)N * Ho * Wo * c4n;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
    int64_t pix = idx / c4n; int c = (int)(idx - pix * c4n) * 4;
    int n = (int)(pix / ((int64_t)Ho * Wo));
    int r = (int)(pix - (int64_t)n * Ho * Wo);
    int oy = r / Wo, ox = r - oy * Wo;
    int y0, y1, x0, x1; float ly, lx;
    src_index(sy, oy, Hi, y0, y1, ly);
    src_index(sx, ox, Wi, x0, x1, lx);
    cptr_t<T> b = in + ((int64_t)n * Hi * Wi * in_ld + c);
    float v00[4], v01[4], v10[4], v11[4], o[4];
    load4(b + ((int64_t)y0 * Wi + x0) * in_ld, v00);
    load4(b + ((int64_t)y0 * Wi + x1) * in_ld, v01);
    load4(b + ((int64_t)y1 * Wi + x0) * in_ld, v10);
    load4(b + ((int64_t)y1 * Wi + x1) * in_ld, v11);
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = hy * (hx * v00[j] + lx * v01[j]) + ly * (hx * v10[j] + lx * v11[j]);
    if (add) {
      float a[4];
      load4(add + (pix * add_ld + c), a);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = a[j] + o[j];
    }
    store4(out + (pix * out_ld + c), o);
    if (out_relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
      store4(out_relu + (pix * out_relu_ld + c), o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample8_kernel(cptr_t<T> in, int64_t in_ld, int Hi, int Wi,
                                                        int C, int Ho, int Wo, float sy, float sx,
                                                        cptr_t<T> add, int64_t add_ld,
                                                        ptr_t<T> out, int64_t out_ld,
                                                        ptr_t<T> out_relu, int64_t out_relu_ld, int N) {
  pdl_sync();                                  // PDL contract (common.cuh)
  // (32-bit index arithmetic: the host takes this kernel only for < 2^31 channel octets; three 64-bit divisions per
  // element made it issue-bound at 2.3 TB/s)
  const uint32_t c8n = (uint32_t)C >> 3;
  const uint32_t total = (uint32_t)N * (uint32_t)Ho * (uint32_t)Wo * c8n;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t hw = (uint32_t)Ho * (uint32_t)Wo;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
    const uint32_t pix32 = idx / c8n; const int c = (int)(idx - pix32 * c8n) * 8;
    const int64_t pix = pix32;
    const int n = (int)(pix32 / hw);
    const int r = (int)(pix32 - (uint32_t)n * hw);
    const int oy = r / Wo, ox = r - oy * Wo;
    int y0, y1, x0, x1; float ly, lx;
    src_index(sy, oy, Hi, y0, y1, ly);
    src_index(sx, ox, Wi, x0, x1, lx);
    cptr_t<T> b = in + ((int64_t)n * Hi * Wi * in_ld + c);
    float v00[8], v01[8], v10[8], v11[8], o[8], a8[8];
    load8(b + ((int64_t)y0 * Wi + x0) * in_ld, v00);
    load8(b + ((int64_t)y0 * Wi + x1) * in_ld, v01);
    load8(b + ((int64_t)y1 * Wi + x0) * in_ld, v10);
    load8(b + ((int64_t)y1 * Wi + x1) * in_ld, v11);
    if (add) load8(add + (pix * add_ld + c), a8);
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = hy * (hx * v00[j] + lx * v01[j]) + ly * (hx * v10[j] + lx * v11[j]);
      if (add) o[j] = a8[j] + o[j];
    }
    store8(out + (pix * out_ld + c), o);
    if (out_relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
      store8(out_relu + (pix * out_relu_ld + c), o);
    }
  }
}

// Exact x2 variant (Ho = 2 Hi, Wo = 2 Wi: every big resize of the frame): a thread owns one INPUT pixel and 8 channels,
// loads its 3 x 3 neighbourhood once and writes the 2 x 2 output block that lies inside it -- 9 loads per 4 outputs
// instead of 16.  Output (2m + a, 2k + b) interpolates rows {m - 1 + a, m + a} and columns {k - 1 + b, k + b} of the
// neighbourhood with the weights of src_index; where src_index clamps (image border) its weight on the clamped
// neighbour is exactly 0, so the statically chosen (clamped) neighbour contributes 0 as well: same arithmetic.
template <typename T>
__global__ void __launch_bounds__(256) upsample2x8_kernel(cptr_t<T> in, int64_t in_ld, int Hi, int Wi, int C,
                                                          cptr_t<T> add, int64_t add_ld, ptr_t<T> out, int64_t out_ld,
                                                          ptr_t<T> out_relu, int64_t out_relu_ld, int N) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const uint32_t c8n = (uint32_t)C >> 3, hw = (uint32_t)Hi * (uint32_t)Wi;
  const uint32_t total = (uint32_t)N * hw * c8n, nthreads = gridDim.x * blockDim.x;
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
    const uint32_t pix = idx / c8n; const int c = (int)(idx - pix * c8n) * 8;
    const int n = (int)(pix / hw);
    const int r = (int)(pix - (uint32_t)n * hw);
    const int m = r / Wi, k = r - m * Wi;
    cptr_t<T> b = in + ((int64_t)n * Hi * Wi * in_ld + c);
    float v[3][3][8];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = min(max(m - 1 + dy, 0), Hi - 1);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int x = min(max(k - 1 + dx, 0), Wi - 1);
        load8(b + ((int64_t)y * Wi + x) * in_ld, v[dy][dx]);
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      int y0, y1; float ly;
      src_index(0.5f, 2 * m + a, Hi, y0, y1, ly);
      const float hy = 1.f - ly;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        int x0, x1; float lx;
        src_index(0.5f, 2 * k + bb, Wi, x0, x1, lx);
        const float hx = 1.f - lx;
        const int64_t opix = ((int64_t)n * Ho + (2 * m + a)) * Wo + (2 * k + bb);
        float o[8], a8[8];
        if (add) load8(add + (opix * add_ld + c), a8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = hy * (hx * v[a][bb][j] + lx * v[a][bb + 1][j]) + ly * (hx * v[a + 1][bb][j] + lx * v[a + 1][bb + 1][j]);
          if (add) o[j] = a8[j] + o[j];
        }
        store8(out + (opix * out_ld + c), o);
        if (out_relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
          store8(out_relu + (opix * out_relu_ld + c), o);
        }
      }
    }
  }
}

// scalar-channel variant writing fp32 (NHWC with out_ld, or NCHW planes): the 3-channel STM logits (STM.py:136)
template <typename T>
__global__ void __launch_bounds__(256) upsample_scalar_kernel(cptr_t<T> in, int64_t in_ld, int Hi,
                                                              int Wi, int C, int Ho, int Wo, float sy, float sx,
                                                              float* __restrict__ out, int64_t out_ld, int nchw) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int64_t total = (int64_t)Ho * Wo;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += nthreads) {
    int oy = (int)(pix / Wo), ox = (int)(pix - (int64_t)oy * Wo);
    int y0, y1, x0, x1; float ly, lx;
    src_index(sy, oy, Hi, y0, y1, ly);
    src_index(sx, ox, Wi, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    for (int c = 0; c < C; ++c) {
      float v00 = ld1(in, ((int64_t)y0 * Wi + x0) * in_ld + c), v01 = ld1(in, ((int64_t)y0 * Wi + x1) * in_ld + c);
      float v10 = ld1(in, ((int64_t)y1 * Wi + x0) * in_ld + c), v11 = ld1(in, ((int64_t)y1 * Wi + x1) * in_ld + c);
      float o = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
      if (nchw) out[(int64_t)c * total + pix] = o; else out[pix * out_ld + c] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// MaxPool2d(3, 2, 1)
// ------------------------------------------------------------------------------------------------------
template <typename T, int V>                     // V channels per thread: 8 (16-byte accesses) when C % 8 == 0, else 4
__global__ void __launch_bounds__(256) maxpool_kernel(cptr_t<T> in, int64_t in_ld, int N, int H, int W,
                                                      int C, int Ho, int Wo, ptr_t<T> out, int64_t out_ld) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const uint32_t cvn = (uint32_t)C / V;
  const uint32_t hw = (uint32_t)Ho * (uint32_t)Wo;
  const uint32_t total = (uint32_t)N * hw * cvn;                    // (host: < 2^31)
  const uint32_t nthreads = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
    const uint32_t pix = idx / cvn; const int c = (int)(idx - pix * cvn) * V;
    const int n = (int)(pix / hw);
    const int r = (int)(pix - (uint32_t)n * hw);
    const int oy = r / Wo, ox = r - oy * Wo;
    float m[V];
#pragma unroll
    for (int j = 0; j < V; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = oy * 2 - 1 + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = ox * 2 - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        float v[V];
        if constexpr (V == 8) load8(in + (((int64_t)(n * H + iy) * W + ix) * in_ld + c), v);
        else load4(in + (((int64_t)(n * H + iy) * W + ix) * in_ld + c), v);
#pragma unroll
        for (int j = 0; j < V; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    }
    if constexpr (V == 8) store8(out + ((int64_t)pix * out_ld + c), m);
    else store4(out + ((int64_t)pix * out_ld + c), m);
  }
}

// ------------------------------------------------------------------------------------------------------
// Pyramid pooling: AdaptiveAvgPool2d(s) for s in {1,2,3,6} from ONE read of the feature map.
// pass 1: per row y, the 12 column-bin sums (1+2+3+6) for every channel; pass 2: the 50 cells from row sums.
// bin b of scale s covers [floor(b*L/s), ceil((b+1)*L/s))  (ATen adaptive pooling).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int bin_start(int b, int s, int L) { return (b * L) / s; }
__device__ __forceinline__ int bin_end(int b, int s, int L) { return ((b + 1) * L + s - 1) / s; }

template <typename T>
__global__ void __launch_bounds__(128) ppm_rows_kernel(cptr_t<T> in, int64_t in_ld, int H, int W, int C,
                                                       float* __restrict__ rows) {
  // One block = one row y x 128 channels: lane = channel quad, warp w = quarter w of the row.  A warp first parks its
  // (<= 32) pixels in shared memory -- every load of the quarter in flight at once -- then walks the 12 bins with
  // warp-uniform pixel ranges (static accumulators, no predicated adds); the four quarters' bin sums meet in shared
  // memory and are added in quarter order (deterministic).  The first version walked a whole row per thread with one
  // load in flight: 32 K threads, 42-47 us for the 33 MB layer-4 map.
  pdl_sync();                                  // PDL contract (common.cuh)
  constexpr int NB = 16;                         // pixels per batch
  __shared__ float4 px[4][NB][32];               // 32 KB; reused for the quarters' bin sums afterwards
  float (*part)[12][32][4] = reinterpret_cast<float (*)[12][32][4]>(&px[0][0][0]);      // [4][12][32][4] floats = 24 KB
  const int y = blockIdx.y, n = blockIdx.z;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * 4;
  const bool live = c < C;
  const int seg = (W + 3) >> 2, xa = wq * seg, xb = min(W, xa + seg);
  float4 acc[12];
#pragma unroll
  for (int bi = 0; bi < 12; ++bi) acc[bi] = make_float4(0.f, 0.f, 0.f, 0.f);
  cptr_t<T> row = in + (((int64_t)(n * H + y) * W) * in_ld + (live ? c : 0));
  for (int x0 = xa; x0 < xb; x0 += NB) {
    const int x1 = min(xb, x0 + NB);
    if (live) {
#pragma unroll 8
      for (int x = x0; x < x1; ++x) {
        float v[4];
        load4(row + (int64_t)x * in_ld, v);
        px[wq][x - x0][lane] = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    __syncwarp();
    int bi = 0;
#pragma unroll
    for (int si = 0; si < 4; ++si) {
      const int s = si == 0 ? 1 : si == 1 ? 2 : si == 2 ? 3 : 6;
#pragma unroll
      for (int b = 0; b < s; ++b, ++bi) {
        const int lo = max(bin_start(b, s, W), x0), hi = min(bin_end(b, s, W), x1);     // warp-uniform
        for (int x = lo; x < hi; ++x) {
          const float4 t = px[wq][x - x0][lane];
          acc[bi].x += t.x; acc[bi].y += t.y; acc[bi].z += t.z; acc[bi].w += t.w;
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();                               // every warp is done with its pixels: the buffer becomes `part`
#pragma unroll
  for (int bi = 0; bi < 12; ++bi) *reinterpret_cast<float4*>(part[wq][bi][lane]) = acc[bi];
  __syncthreads();
  if (!live) return;
  float* o = rows + (((int64_t)n * H + y) * 12) * C + c;
#pragma unroll
  for (int k = 0; k < 3; ++k) {                   // warp wq finishes bins 3 wq .. 3 wq + 2
    const int b = wq * 3 + k;
    float4 t = *reinterpret_cast<const float4*>(part[0][b][lane]);
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      const float4 u = *reinterpret_cast<const float4*>(part[q][b][lane]);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    *reinterpret_cast<float4*>(o + (int64_t)b * C) = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(128) ppm_cells_kernel(const float* __restrict__ rows, int H, int W, int C,
                                                        ptr_t<T> out) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int cell = blockIdx.x, n = blockIdx.z;          // 0..49: scale-major, row-major inside a scale
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= C) return;
  int s, local, xoff;
  if (cell < 1) { s = 1; local = cell; xoff = 0; }
  else if (cell < 5) { s = 2; local = cell - 1; xoff = 1; }
  else if (cell < 14) { s = 3; local = cell - 5; xoff = 3; }
  else { s = 6; local = cell - 14; xoff = 6; }
  const int by = local / s, bx = local - by * s;
  const int y0 = bin_start(by, s, H), y1 = bin_end(by, s, H);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int yb = y0; yb < y1; yb += 8) {            // 8 row sums in flight, added in row order
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (yb + u < y1) v[u] = __ldg(reinterpret_cast<const float4*>(rows + (((int64_t)n * H + yb + u) * 12 + xoff + bx) * C + c));
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (yb + u < y1) { acc[0] += v[u].x; acc[1] += v[u].y; acc[2] += v[u].z; acc[3] += v[u].w; }
  }
  const float inv = 1.f / (float)((y1 - y0) * (bin_end(bx, s, W) - bin_start(bx, s, W)));
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] *= inv;
  store4(out + (((int64_t)n * 50 + cell) * C + c), acc);
}

// ------------------------------------------------------------------------------------------------------
// layout conversion at the boundary
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int64_t HW, ptr_t<T> out,
                                    int64_t out_ld, int N) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int64_t total = (int64_t)N * HW * C;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
    int c = (int)(idx % C); int64_t pix = idx / C;             // pix = n*HW + p
    int64_t n = pix / HW, p = pix - n * HW;
    st1(out, pix * out_ld + c, in[(n * C + c) * HW + p]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(cptr_t<T> in, int64_t in_ld, int C, int64_t HW,
                                    float* __restrict__ out, int N) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int64_t total = (int64_t)N * HW * C;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
    int64_t p = idx % HW; int64_t r = idx / HW; int c = (int)(r % C); int64_t n = r / C;
    out[idx] = ld1(in, (n * HW + p) * in_ld + c);
  }
}

}  // namespace otvm

using namespace otvm;

// CALL is an expression template in the element type T; `ps` (plane stride in elements) is in scope for mkptr
#define DISPATCH_DTYPE(dtype, CALL)                                            \
  do {                                                                         \
    const int64_t ps = dtype_plane_stride(dtype); (void)ps;                    \
    switch (dtype_fmt(dtype)) {                                                \
      case OTVM_F32: { typedef float T; return CALL; }                         \
      case OTVM_BF16: { typedef bf16 T; return CALL; }                         \
      case OTVM_BF16X2: { typedef bx<2> T; return CALL; }                      \
      case OTVM_BF16X3: { typedef bx<3> T; return CALL; }                      \
      default: return OTVM_ERR_ARG;                                            \
    }                                                                          \
  } while (0)

extern "C" int otvm_gn_stats(const void* x, int64_t ld, int32_t N, int32_t HW, int32_t C, int32_t dtype,
                             double* stats, void* stream) {
  if (!x || !stats || C % 32 != 0 || C % 4 != 0 || ld % 4 != 0) return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DISPATCH_DTYPE(dtype, gn_stats_t<T>(x, ld, N, HW, C, stats, ps, s));
}

extern "C" int otvm_gn_apply(const void* x, int64_t ld, int32_t N, int32_t HW, int32_t C, int32_t dtype,
                             const double* stats, const float* gamma, const float* beta, float eps,
                             const void* res, int64_t res_ld, int32_t act, void* out, int64_t out_ld,
                             void* stream) {
  if (!x || !stats || !out || C % 64 != 0 || C > 2048 || ld % 8 != 0 || out_ld % 8 != 0 || (res && res_ld % 8 != 0))
    return OTVM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res)) & 15) return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DISPATCH_DTYPE(dtype, gn_apply_t<T>(x, ld, N, HW, C, stats, gamma, beta, eps, res, res_ld, act, out, out_ld, ps, s));
}

template <typename T>
static int upsample_t(const void* in, int64_t in_ld, int N, int Hi, int Wi, int C, int Ho, int Wo,
                      const void* add, int64_t add_ld, void* out, int64_t out_ld, void* out_relu,
                      int64_t out_relu_ld, int out_nchw_f32, int64_t ps, cudaStream_t s) {
  // ATen: scale = in/out computed in float (F.interpolate with an integer scale_factor gives the same value)
  float sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
  if (out_nchw_f32 || C % 4 != 0) {
    if (add || out_relu || N != 1) return OTVM_ERR_UNSUPPORTED;
    launch_k(upsample_scalar_kernel<T>, grid_for((int64_t)Ho * Wo, 256), 256, 0, s, 
        mkcptr<T>(in, ps), in_ld, Hi, Wi, C, Ho, Wo, sy, sx, static_cast<float*>(out), out_ld,
        out_nchw_f32 == 1);
  } else {
    if (in_ld % 4 || out_ld % 4 || (add && add_ld % 4) || (out_relu && out_relu_ld % 4)) return OTVM_ERR_ARG;
    const bool wide = C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 && (!add || add_ld % 8 == 0) &&
                      (!out_relu || out_relu_ld % 8 == 0) &&
                      !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(add) |
                         reinterpret_cast<uintptr_t>(out_relu)) & 15);
    if (wide && Ho == 2 * Hi && Wo == 2 * Wi && (int64_t)N * Ho * Wo * (C / 8) < (int64_t)0x7fffffff) {
      launch_k(upsample2x8_kernel<T>, grid_for((int64_t)N * Hi * Wi * (C / 8), 256), 256, 0, s, mkcptr<T>(in, ps), in_ld, Hi, Wi, C,
               mkcptr<T>(add, ps), add_ld, mkptr<T>(out, ps), out_ld, mkptr<T>(out_relu, ps), out_relu_ld, N);
      OTVM_LAUNCH_CHECK();
      return OTVM_OK;
    }
    if (wide && (int64_t)N * Ho * Wo * (C / 8) < (int64_t)0x7fffffff) {
      int64_t total8 = (int64_t)N * Ho * Wo * (C / 8);
      launch_k(upsample8_kernel<T>, grid_for(total8, 256), 256, 0, s, mkcptr<T>(in, ps), in_ld, Hi, Wi, C, Ho, Wo,
                                                                sy, sx, mkcptr<T>(add, ps), add_ld,
                                                                mkptr<T>(out, ps), out_ld,
                                                                mkptr<T>(out_relu, ps), out_relu_ld, N);
      OTVM_LAUNCH_CHECK();
      return OTVM_OK;
    }
    int64_t total = (int64_t)N * Ho * Wo * (C / 4);
    launch_k(upsample_kernel<T>, grid_for(total, 256), 256, 0, s, mkcptr<T>(in, ps), in_ld, Hi, Wi, C, Ho, Wo,
                                                           sy, sx, mkcptr<T>(add, ps), add_ld,
                                                           mkptr<T>(out, ps), out_ld,
                                                           mkptr<T>(out_relu, ps), out_relu_ld, N);
  }
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_upsample_bilinear(const void* in, int64_t in_ld, int32_t N, int32_t Hi, int32_t Wi, int32_t C,
                                      int32_t Ho, int32_t Wo, const void* add, int64_t add_ld, void* out,
                                      int64_t out_ld, void* out_relu, int64_t out_relu_ld, int32_t dtype,
                                      int32_t out_nchw_f32, void* stream) {
  if (!in || !out) return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DISPATCH_DTYPE(dtype, upsample_t<T>(in, in_ld, N, Hi, Wi, C, Ho, Wo, add, add_ld, out, out_ld, out_relu, out_relu_ld,
                                      out_nchw_f32, ps, s));
}

template <typename T>
static int maxpool_t(const void* in, int64_t in_ld, int N, int H, int W, int C, void* out, int64_t out_ld,
                     int64_t ps, cudaStream_t s) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  int64_t total = (int64_t)N * Ho * Wo * (C / 4);
  if (total >= (int64_t)0x7fffffff) return OTVM_ERR_UNSUPPORTED;
  const bool wide = C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 &&
                    !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15);
  if (wide)
    launch_k(maxpool_kernel<T, 8>, grid_for(total / 2, 256), 256, 0, s, mkcptr<T>(in, ps), in_ld, N, H, W, C, Ho, Wo,
                                                             mkptr<T>(out, ps), out_ld);
  else
    launch_k(maxpool_kernel<T, 4>, grid_for(total, 256), 256, 0, s, mkcptr<T>(in, ps), in_ld, N, H, W, C, Ho, Wo,
                                                          mkptr<T>(out, ps), out_ld);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_maxpool3x3s2(const void* in, int64_t in_ld, int32_t N, int32_t H, int32_t W, int32_t C,
                                 void* out, int64_t out_ld, int32_t dtype, void* stream) {
  if (!in || !out || C % 4 || in_ld % 4 || out_ld % 4) return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DISPATCH_DTYPE(dtype, maxpool_t<T>(in, in_ld, N, H, W, C, out, out_ld, ps, s));
}

template <typename T>
static int ppm_t(const void* in, int64_t in_ld, int N, int H, int W, int C, void* out, float* scratch,
                 int64_t ps, cudaStream_t s) {
  dim3 g1(ceil_div(C / 4, 32), H, N), g2(50, ceil_div(C / 4, 128), N);
  launch_k(ppm_rows_kernel<T>, g1, 128, 0, s, mkcptr<T>(in, ps), in_ld, H, W, C, scratch);
  OTVM_LAUNCH_CHECK();
  launch_k(ppm_cells_kernel<T>, g2, 128, 0, s, scratch, H, W, C, mkptr<T>(out, ps));
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_ppm_pool(const void* in, int64_t in_ld, int32_t N, int32_t H, int32_t W, int32_t C, void* out,
                             float* scratch, int32_t dtype, void* stream) {
  if (!in || !out || !scratch || C % 4 || in_ld % 4) return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DISPATCH_DTYPE(dtype, ppm_t<T>(in, in_ld, N, H, W, C, out, scratch, ps, s));
}

extern "C" int otvm_nchw_to_nhwc(const float* in, int32_t N, int32_t C, int32_t HW, void* out, int64_t out_ld,
                                 int32_t dtype, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int g = grid_for((int64_t)N * HW * C, 256);
  DISPATCH_DTYPE(dtype, (launch_k(nchw_to_nhwc_kernel<T>, g, 256, 0, s, in, C, HW, mkptr<T>(out, ps), out_ld, N), launched()));
}

extern "C" int otvm_nhwc_to_nchw(const void* in, int64_t in_ld, int32_t N, int32_t C, int32_t HW, float* out,
                                 int32_t dtype, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int g = grid_for((int64_t)N * HW * C, 256);
  DISPATCH_DTYPE(dtype, (launch_k(nhwc_to_nchw_kernel<T>, g, 256, 0, s, mkcptr<T>(in, ps), in_ld, C, HW, out, N), launched()));
}
