// Per-frame glue of EvalModel.forward (reference models/alpha/model.py:391-512) as fused HBM kernels:
// preprocessing + trimap dilation, the 8-channel trimap encoding with an EXACT on-device Euclidean distance
// transform (replaces the three host round trips through cv2.distanceTransform, utils/utils.py:12-23),
// the FBA fusion heads, and the output / memorize-input packing.
#include "common.cuh"

namespace otvm {

constexpr int EDT_INF = 0x3fffffff;

struct MeanStd { float mean[3], std[3]; };

static inline int grid1d(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ---------------------------------------------------------------------------------------------------
// preprocess_gt: composite + unknown mask (models/alpha/model.py:380-389, 342-349)
// ---------------------------------------------------------------------------------------------------
__global__ void composite_kernel(const float* __restrict__ a, const float* __restrict__ fg,
                                 const float* __restrict__ bg, int H, int W, int Wp, int pad_top, int pad_left,
                                 float* __restrict__ img, float* __restrict__ scaled, uint8_t* __restrict__ unk) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int64_t P = (int64_t)H * W;
  const float kScale = 1.f / 255;                     // IMG_SCALE, models/alpha/model.py:27
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const float al = a[p];
    int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    float* o = img + ((int64_t)(y + pad_top) * Wp + x + pad_left) * 4;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                     // flip([2]): BGR -> RGB
      float f = fg[(2 - c) * P + p] * kScale, b = bg[(2 - c) * P + p] * kScale;
      float v = f * al + b * (1.f - al);
      o[c] = v; scaled[c * P + p] = v;
    }
    o[3] = 0.f;
    unk[p] = (al > 0.f && al < 1.f) ? 1 : 0;
  }
}

// separable (2r+1)^2 max filter == F.max_pool2d(k=2r+1, s=1, p=r) on a {0,1} mask (:353)
__global__ void dilate_rows_kernel(const uint8_t* __restrict__ in, int H, int W, int r, uint8_t* __restrict__ out) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int64_t P = (int64_t)H * W;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    int lo = max(0, x - r), hi = min(W - 1, x + r);
    uint8_t m = 0;
    for (int i = lo; i <= hi; ++i) m |= in[(int64_t)y * W + i];
    out[p] = m;
  }
}

template <typename T>
__global__ void dilate_cols_onehot_kernel(const uint8_t* __restrict__ rows, const float* __restrict__ a, int H, int W,
                                          int Hp, int Wp, int pad_top, int pad_left, int r,
                                          float* __restrict__ img, float* __restrict__ tri3, MeanStd ms,
                                          ptr_t<T> imgn, int64_t imgn_ld) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int64_t P = (int64_t)Hp * Wp;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    int yp = (int)(p / Wp), xp = (int)(p - (int64_t)yp * Wp);
    int y = yp - pad_top, x = xp - pad_left;
    float4 t = make_float4(1.f, 0.f, 0.f, 0.f);       // padding is background (:409-410)
    if (y >= 0 && y < H && x >= 0 && x < W) {
      int lo = max(0, y - r), hi = min(H - 1, y + r);
      uint8_t m = 0;
      for (int i = lo; i <= hi; ++i) m |= rows[(int64_t)i * W + x];
      float al = a[(int64_t)y * W + x];
      al = al < 0.f ? 0.f : (al > 1.f ? 1.f : al);
      int cls = m ? 1 : (int)(2.f * al);              // torch.where(trimap>0.5, 1, 2*alpha).long()  (:360)
      t = make_float4(cls == 0 ? 1.f : 0.f, cls == 1 ? 1.f : 0.f, cls == 2 ? 1.f : 0.f, 0.f);
    } else {
      *reinterpret_cast<float4*>(img + p * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4*>(tri3 + p * 4) = t;
    // (f - mean) / std of the padded frame: the STM query encoder input (STM.py:93)
    float4 im = *reinterpret_cast<const float4*>(img + p * 4);
    float q[4] = {(im.x - ms.mean[0]) / ms.std[0], (im.y - ms.mean[1]) / ms.std[1], (im.z - ms.mean[2]) / ms.std[2], 0.f};
    store4(imgn + p * imgn_ld, q);
  }
}

// ---------------------------------------------------------------------------------------------------
// exact squared Euclidean distance transform (two passes, integer arithmetic)
//   pass 1 (rows)   : g2[y][x] = squared distance to the nearest seed in row y -- one warp per row, the nearest seed
//                     to the left / right comes from a warp max / min scan carried across 32-pixel chunks
//   pass 2 (columns): d2[y][x] = min_y' (y-y')^2 + g2[y'][x] -- a 32-column strip of g2 sits in shared memory and
//                     every pixel searches its column outward from its own row until no farther row can win
//                     (integer, exact; the exit test is warp-uniform)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edt_rowscan_kernel(const uint8_t* __restrict__ seed, int H, int W, int nmask,
                                                          int* __restrict__ g2) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= H * nmask) return;
  const uint8_t* s = seed + (int64_t)warp * W;        // rows of all masks are contiguous
  int* g = g2 + (int64_t)warp * W;
  // left-to-right: position of the nearest seed at or before x
  int carry = -EDT_INF;
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    int pos = (x < W && s[x]) ? x : -EDT_INF;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, pos, o); if (lane >= o) pos = max(pos, t); }
    pos = max(pos, carry);
    if (x < W) g[x] = pos;                            // stash the left position; finished below
    carry = __shfl_sync(0xffffffffu, pos, 31);
  }
  carry = EDT_INF;
  for (int x0 = ((W - 1) / 32) * 32; x0 >= 0; x0 -= 32) {
    const int x = x0 + lane;
    int pos = (x < W && s[x]) ? x : EDT_INF;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_down_sync(0xffffffffu, pos, o); if (lane + o < 32) pos = min(pos, t); }
    pos = min(pos, carry);
    if (x < W) {
      const int left = g[x];
      const int dl = left <= -EDT_INF ? EDT_INF : x - left, dr = pos >= EDT_INF ? EDT_INF : pos - x;
      const int d = min(dl, dr);
      g[x] = d >= 32768 ? EDT_INF : d * d;
    }
    carry = __shfl_sync(0xffffffffu, pos, 0);
  }
}

__global__ void __launch_bounds__(256) edt_cols_kernel(const int* __restrict__ g2, int H, int W, int* __restrict__ d2) {
  pdl_sync();                                  // PDL contract (common.cuh)
  extern __shared__ int strip[];                      // [H][32]
  const int x0 = blockIdx.x * 32, m = blockIdx.y;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5, nwy = blockDim.x >> 5;
  const int x = x0 + lane;
  const int* g = g2 + (int64_t)m * H * W;
  // columns beyond the image hold 0 so that they never keep the warp-uniform search below alive
  int any = 0;
  for (int y = wy; y < H; y += nwy) {
    const int v = x < W ? g[(int64_t)y * W + x] : 0;
    strip[y * 32 + lane] = v;
    any |= (x < W && v < EDT_INF);
  }
  const int has_seed = __syncthreads_or(any);
  // blockIdx.z splits the rows so that the grid covers the GPU (the whole strip is still needed for the search)
  const int rows_per = (H + gridDim.z - 1) / gridDim.z;
  const int y_lo = blockIdx.z * rows_per, y_hi = min(H, y_lo + rows_per);
  for (int y = y_lo + wy; y < y_hi; y += nwy) {
    // Outward search from the pixel's own row, 8 row pairs (y-d, y+d) per step, branch-free inside a step.  A row at
    // distance d can only contribute values >= d*d, so the warp stops once d*d >= the largest running minimum of its
    // 32 columns (warp-uniform exit: no divergence; a mask with few seeds degenerates into the full column scan at
    // the cost of the old exhaustive loop).  Row indices are clamped instead of predicated: a clamped row r' is
    // closer than d, so strip[r'] + d*d over-estimates a candidate that is (or was) also taken at its true distance.
    int best = strip[y * 32 + lane];
    if (has_seed) {                                    // an empty mask has no finite distance anywhere
      const int rmax = max(y, H - 1 - y);
      for (int r = 1; r <= rmax; r += 8) {
        if (r * r >= __reduce_max_sync(0xffffffffu, best)) break;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int d = r + u;
          const int up = strip[max(y - d, 0) * 32 + lane], dn = strip[min(y + d, H - 1) * 32 + lane];
          best = min(best, min(up, dn) + d * d);
        }
      }
    }
    if (x < W) d2[((int64_t)m * H + y) * W + x] = best >= EDT_INF ? EDT_INF : best;
  }
}

static int edt_launch(const uint8_t* seed, int H, int W, int nmask, int* d2, int* scratch, cudaStream_t s) {
  launch_k(edt_rowscan_kernel, ceil_div((int64_t)H * nmask * 32, 256), 256, 0, s, seed, H, W, nmask, scratch);
  OTVM_LAUNCH_CHECK();
  const size_t smem = (size_t)H * 32 * sizeof(int);
  if (smem > 200 * 1024) return OTVM_ERR_UNSUPPORTED;
  OTVM_CUDA_CHECK((ensure_dynamic_smem<edt_cols_kernel>(200 * 1024)));
  int ysplit = ceil_div(2 * sm_count(), ceil_div(W, 32) * nmask);
  if (ysplit > 32) ysplit = 32;
  if (ysplit < 1) ysplit = 1;
  launch_k(edt_cols_kernel, dim3(ceil_div(W, 32), nmask, ysplit), 256, smem, s, scratch, H, W, d2);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

// ---------------------------------------------------------------------------------------------------
// make_trimap (models/alpha/model.py:40-53): classes + seeds, then the 11-channel FBA input
// ---------------------------------------------------------------------------------------------------

__global__ void trimap_classes_kernel(const float* __restrict__ tri, int64_t tri_ld, int is_logit,
                                      const float* __restrict__ img, int64_t P, float* __restrict__ extras,
                                      uint8_t* __restrict__ seeds) {
  pdl_sync();                                  // PDL contract (common.cuh)
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    float t0 = tri[p * tri_ld], t1 = tri[p * tri_ld + 1], t2 = tri[p * tri_ld + 2];
    if (is_logit) {                                   // F.softmax(_logit_trimap, dim=1)  (:440)
      float mx = fmaxf(t0, fmaxf(t1, t2));
      float e0 = expf(t0 - mx), e1 = expf(t1 - mx), e2 = expf(t2 - mx);
      float inv = 1.f / (e0 + e1 + e2);
      t0 = e0 * inv; t1 = e1 * inv; t2 = e2 * inv;
    }
    int cls = 0; float best = t0;                     // tri.max(dim=2)[1]: first maximum wins
    if (t1 > best) { best = t1; cls = 1; }
    if (t2 > best) { cls = 2; }
    seeds[p] = cls == 0;                              // trimap2b = (scaled == 0)
    seeds[P + p] = cls == 2;                          // trimap2f = (scaled == 1)
    float4 im = *reinterpret_cast<const float4*>(img + p * 4);
    float* e = extras + p * 8;
    *reinterpret_cast<float4*>(e) = make_float4(im.x, im.y, im.z, t0);
    *reinterpret_cast<float4*>(e + 4) = make_float4(t2, t0, t1, t2);
  }
}

template <typename T>
__global__ void trimap_pack_kernel(const float* __restrict__ extras, const int* __restrict__ d2, int64_t P,
                                   MeanStd ms, ptr_t<T> x11, int64_t x11_ld, ptr_t<T> cat_dst,
                                   int64_t cat_ld, bool wide) {
  pdl_sync();                                  // PDL contract (common.cuh)
  // trimap_transform, utils/utils.py:25-39: exp(-d^2 / (2 (sigma L)^2)), sigma in {.02,.08,.16}, L = 320
  const float den0 = (float)(2.0 * (0.02 * 320) * (0.02 * 320));
  const float den1 = (float)(2.0 * (0.08 * 320) * (0.08 * 320));
  const float den2 = (float)(2.0 * (0.16 * 320) * (0.16 * 320));
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    float e[8];                                       // (two 16-byte loads; the scalar form issued five)
    {
      const float4 e0 = *reinterpret_cast<const float4*>(extras + p * 8), e1 = *reinterpret_cast<const float4*>(extras + p * 8 + 4);
      e[0] = e0.x; e[1] = e0.y; e[2] = e0.z; e[3] = e0.w; e[4] = e1.x; e[5] = e1.y; e[6] = e1.z; e[7] = e1.w;
    }
    float v[16];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (e[c] - ms.mean[c]) / ms.std[c];       // (:414)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int q = d2[(int64_t)k * P + p];
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (q < EDT_INF) {                              // no seed at all -> clicks stay 0 (utils/utils.py:31)
        float d = sqrtf((float)q);                    // cv2 returns the float distance; the reference squares it
        float dm = -(d * d);
        g0 = expf(dm / den0); g1 = expf(dm / den1); g2 = expf(dm / den2);
      }
      v[3 + 3 * k] = g0; v[4 + 3 * k] = g1; v[5 + 3 * k] = g2;
    }
    v[9] = e[3]; v[10] = e[4];                        // soft bg, soft fg (:51)
#pragma unroll
    for (int c = 11; c < 16; ++c) v[c] = 0.f;
    // 16-byte stores per plane (8-byte pieces at a 32 / 192-byte pixel pitch wrote a quarter of every sector they touched:
    // 51 MB of DRAM traffic and 30 us for 24 MB of output)
    ptr_t<T> o = x11 + p * x11_ld;
    if (wide) {
      float q8[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]}, r8[8] = {v[8], v[9], v[10], 0.f, 0.f, 0.f, 0.f, 0.f};
      store8(o, q8); store8(o + 8, r8);
      if (cat_dst) {                                  // cat(.., conv_out[-6][:, :3], img, two_chan_trimap) (:377-378)
        float c8[8] = {v[0], v[1], v[2], e[0], e[1], e[2], e[3], e[4]};
        store8(cat_dst + p * cat_ld, c8);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; c += 4) { float q4[4] = {v[c], v[c + 1], v[c + 2], v[c + 3]}; store4(o + c, q4); }
      if (cat_dst) {
        float q0[4] = {v[0], v[1], v[2], e[0]}, q1[4] = {e[1], e[2], e[3], e[4]};
        store4(cat_dst + p * cat_ld, q0); store4(cat_dst + (p * cat_ld + 4), q1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// clamp / sigmoid / fba_fusion (FBA/models.py:279-288)
// ---------------------------------------------------------------------------------------------------
// clamp / sigmoid / fba_fusion of one pixel (FBA/models.py:279-288, 383-390): raw = (alpha, F rgb, B rgb), img rgb
__device__ __forceinline__ void fba_fuse(const float (&raw)[7], const float (&img)[3], float& a2, float (&Fo)[3], float (&Bo)[3]) {
  const float al = fminf(fmaxf(raw[0], 0.f), 1.f);
  float F[3], B[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    F[c] = 1.f / (1.f + expf(-raw[1 + c]));
    B[c] = 1.f / (1.f + expf(-raw[4 + c]));
  }
  float num = 0.f, den = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float f = al * img[c] + (1.f - al * al) * F[c] - al * (1.f - al) * B[c];
    float b = (1.f - al) * img[c] + (2.f * al - al * al) * B[c] - al * (1.f - al) * f;   // uses updated F (:281)
    f = fminf(fmaxf(f, 0.f), 1.f);
    b = fminf(fmaxf(b, 0.f), 1.f);
    Fo[c] = f; Bo[c] = b;
    num += (img[c] - b) * (f - b);
    den += (f - b) * (f - b);
  }
  a2 = (al * 0.1f + num) / (den + 0.1f);
  a2 = fminf(fmaxf(a2, 0.f), 1.f);
}

template <typename TR, typename TA>
__global__ void fba_head_kernel(cptr_t<TR> raw, int64_t raw_ld, const float* __restrict__ extras,
                                int64_t P, float* __restrict__ out7, ptr_t<TA> alpha_dst, int64_t alpha_ld) {
  pdl_sync();                                  // PDL contract (common.cuh)
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    cptr_t<TR> r = raw + p * raw_ld;
    float rv[7], img[3], a2, Fo[3], Bo[3];
#pragma unroll
    for (int c = 0; c < 7; ++c) rv[c] = ld1(r, c);
#pragma unroll
    for (int c = 0; c < 3; ++c) img[c] = extras[p * 8 + c];
    fba_fuse(rv, img, a2, Fo, Bo);
    float* o = out7 + p * 8;
    *reinterpret_cast<float4*>(o) = make_float4(a2, Fo[0], Fo[1], Fo[2]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(Bo[0], Bo[1], Bo[2], 0.f);
    if (alpha_dst) st1(alpha_dst, p * alpha_ld, a2);
  }
}

// The 1x1 head convolution (16 -> 7 / 10 channels, FBA/models.py:347 `conv_up4.4`, :415 `pred.4`) + the fusion above in ONE
// pointwise pass: 160 multiply-adds per pixel are no tensor-core work -- as a tcgen05 layer it was 2048 one-tile CTAs and
// 19-22 us, followed by a second kernel that re-read its fp32 output.  fp32 weights and accumulation on the stored
// (split) activations; raw (the head's output, kept: the refined-trimap logits 7..9 are read by otvm_frame_outputs) is
// written as whole 16-byte quads, columns >= Cout as zeros.
template <typename T>
__global__ void __launch_bounds__(256) head_conv_fba_kernel(cptr_t<T> x, int64_t x_ld, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int Cout, float* __restrict__ raw,
                                                            int64_t raw_ld, const float* __restrict__ extras, int64_t P,
                                                            float* __restrict__ out7, ptr_t<T> alpha_dst, int64_t alpha_ld) {
  __shared__ float sw[12][16];
  __shared__ float sb[12];
  for (int i = threadIdx.x; i < 12 * 16; i += blockDim.x) sw[i >> 4][i & 15] = (i >> 4) < Cout ? w[i] : 0.f;   // constants
  if (threadIdx.x < 12) sb[threadIdx.x] = (threadIdx.x < Cout && bias) ? bias[threadIdx.x] : 0.f;
  pdl_sync();                                  // PDL contract (common.cuh)
  __syncthreads();
  const int nq = (int)(raw_ld >> 2);             // 16-byte quads per raw row (2 or 3)
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    float v[16];
    {
      float v0[8], v1[8];
      load8(x + p * x_ld, v0); load8(x + (p * x_ld + 8), v1);
#pragma unroll
      for (int c = 0; c < 8; ++c) { v[c] = v0[c]; v[8 + c] = v1[c]; }
    }
    float r[12];
#pragma unroll
    for (int co = 0; co < 12; ++co) {
      float acc = sb[co];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc = fmaf(sw[co][c], v[c], acc);
      r[co] = acc;
    }
    float* rp = raw + p * raw_ld;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (q < nq) *reinterpret_cast<float4*>(rp + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    float rv[7] = {r[0], r[1], r[2], r[3], r[4], r[5], r[6]}, img[3], a2, Fo[3], Bo[3];
    const float4 e0 = *reinterpret_cast<const float4*>(extras + p * 8);
    img[0] = e0.x; img[1] = e0.y; img[2] = e0.z;
    fba_fuse(rv, img, a2, Fo, Bo);
    float* o = out7 + p * 8;
    *reinterpret_cast<float4*>(o) = make_float4(a2, Fo[0], Fo[1], Fo[2]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(Bo[0], Bo[1], Bo[2], 0.f);
    if (alpha_dst) st1(alpha_dst, p * alpha_ld, a2);
  }
}

// ---------------------------------------------------------------------------------------------------
// refined-trimap softmax, memorize input (22 channels), cropped planar outputs
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void frame_outputs_kernel(const float* __restrict__ raw10, int64_t raw_ld, const float* __restrict__ fused,
                                     cptr_t<T> hid, int64_t hid_ld, const float* __restrict__ extras,
                                     int Hp, int Wp, int H, int W, int pad_top, int pad_left, MeanStd ms,
                                     ptr_t<T> mem_in, int64_t mem_ld, float* __restrict__ alpha_out,
                                     float* __restrict__ trimap_out, bool wide) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int64_t P = (int64_t)Hp * Wp, Pc = (int64_t)H * W;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const float* r = raw10 + p * raw_ld;
    float t0 = r[7], t1 = r[8], t2 = r[9];
    float mx = fmaxf(t0, fmaxf(t1, t2));
    float e0 = expf(t0 - mx), e1 = expf(t1 - mx), e2 = expf(t2 - mx);
    float inv = 1.f / (e0 + e1 + e2);
    t0 = e0 * inv; t1 = e1 * inv; t2 = e2 * inv;       // F.softmax(_logit_trimap_refine) (:460)
    const float al = fused[p * 8];
    if (mem_in) {
      // Encoder_M input order (STM.py:56-67): frame(3, normalised), unknown, fg, alpha, hid(16)
      ptr_t<T> o = mem_in + p * mem_ld;
      float v[8];
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = (extras[p * 8 + c] - ms.mean[c]) / ms.std[c];
      v[3] = t1; v[4] = t2; v[5] = al;
      cptr_t<T> h = hid + p * hid_ld;
      if (wide) {                                     // 16-byte accesses per plane (scalar loads / 8-byte stores before)
        float h0[8], h1[8];
        load8(h, h0); load8(h + 8, h1);
        float q0[8] = {v[0], v[1], v[2], v[3], v[4], v[5], h0[0], h0[1]};
        float q1[8] = {h0[2], h0[3], h0[4], h0[5], h0[6], h0[7], h1[0], h1[1]};
        float q2[8] = {h1[2], h1[3], h1[4], h1[5], h1[6], h1[7], 0.f, 0.f};
        store8(o, q0); store8(o + 8, q1); store8(o + 16, q2);
      } else {
        v[6] = ld1(h, 0); v[7] = ld1(h, 1);
        { float q[4] = {v[0], v[1], v[2], v[3]}; store4(o, q); }
        { float q[4] = {v[4], v[5], v[6], v[7]}; store4(o + 4, q); }
        for (int c = 2; c < 14; c += 4) {
          float q[4] = {ld1(h, c), ld1(h, c + 1), ld1(h, c + 2), ld1(h, c + 3)};
          store4(o + 6 + c, q);
        }
        { float q[4] = {ld1(h, 14), ld1(h, 15), 0.f, 0.f}; store4(o + 20, q); }
      }
    }
    int yp = (int)(p / Wp), xp = (int)(p - (int64_t)yp * Wp);
    int y = yp - pad_top, x = xp - pad_left;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      int64_t q = (int64_t)y * W + x;
      alpha_out[q] = al;
      trimap_out[q] = t0; trimap_out[Pc + q] = t1; trimap_out[2 * Pc + q] = t2;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// frame I/O around the loop (reference dataset.py:857-920 decodes on the host into fp32 tensors, eval.py:209 reads
// the alpha back as fp32 and converts it on the host): here the decoded 8-bit images are uploaded as they are
// (7 B / pixel instead of 28) and unpacked on the device; the matte goes back as 1 B / pixel.
// ---------------------------------------------------------------------------------------------------
__global__ void unpack_frame_u8_kernel(const uint8_t* __restrict__ fg, int fg_c, const uint8_t* __restrict__ bg,
                                       int64_t P, float* __restrict__ a, float* __restrict__ fg_out,
                                       float* __restrict__ bg_out) {
  pdl_sync();                                  // PDL contract (common.cuh)
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t* f = fg + p * fg_c;
    const uint8_t* b = bg + p * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) { fg_out[c * P + p] = (float)f[c]; bg_out[c * P + p] = (float)b[c]; }
    // np.float32(png[..., -1:]) / 255.  (dataset.py:864): IEEE fp32 division; an image without alpha is opaque (:868)
    a[p] = fg_c == 4 ? __fdiv_rn((float)f[3], 255.f) : 1.f;
  }
}

__global__ void alpha_to_u8_kernel(const float* __restrict__ alpha, int64_t P, uint8_t* __restrict__ out) {
  pdl_sync();                                  // PDL contract (common.cuh)
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    // (alphas * 255).byte()  (eval.py:209): fp32 product, conversion truncates toward zero; alpha is clamped to [0,1]
    const float v = __fmul_rn(alpha[p], 255.f);
    out[p] = (uint8_t)(int)fminf(fmaxf(v, 0.f), 255.f);
  }
}

}  // namespace otvm

using namespace otvm;

extern "C" int otvm_unpack_frame_u8(const uint8_t* fg, int32_t fg_channels, const uint8_t* bg, int32_t H, int32_t W,
                                    float* a, float* fg_out, float* bg_out, void* stream) {
  if (!fg || !bg || !a || !fg_out || !bg_out || (fg_channels != 3 && fg_channels != 4) || H <= 0 || W <= 0) return OTVM_ERR_ARG;
  const int64_t P = (int64_t)H * W;
  launch_k(unpack_frame_u8_kernel, grid1d(P, 256), 256, 0, static_cast<cudaStream_t>(stream), fg, (int)fg_channels, bg, P, a,
           fg_out, bg_out);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_alpha_to_u8(const float* alpha, int64_t P, uint8_t* out, void* stream) {
  if (!alpha || !out || P <= 0) return OTVM_ERR_ARG;
  launch_k(alpha_to_u8_kernel, grid1d(P, 256), 256, 0, static_cast<cudaStream_t>(stream), alpha, P, out);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

#define DISPATCH_DTYPE(dtype, STMT)                                            \
  do {                                                                         \
    const int64_t ps = dtype_plane_stride(dtype); (void)ps;                    \
    switch (dtype_fmt(dtype)) {                                                \
      case OTVM_F32: { typedef float T; STMT; break; }                         \
      case OTVM_BF16: { typedef bf16 T; STMT; break; }                         \
      case OTVM_BF16X2: { typedef bx<2> T; STMT; break; }                      \
      case OTVM_BF16X3: { typedef bx<3> T; STMT; break; }                      \
      default: return OTVM_ERR_ARG;                                            \
    }                                                                          \
  } while (0)

static MeanStd make_ms(const float* mean_std) {
  MeanStd ms;
  for (int c = 0; c < 3; ++c) { ms.mean[c] = mean_std[c]; ms.std[c] = mean_std[3 + c]; }
  return ms;
}

extern "C" int otvm_preprocess(const float* a, const float* fg, const float* bg, int32_t H, int32_t W, int32_t Hp,
                               int32_t Wp, int32_t pad_top, int32_t pad_left, int32_t radius, const float* mean_std,
                               float* img, float* scaled_img, float* tri3, void* imgn, int64_t imgn_ld, int32_t dtype,
                               uint8_t* scratch, void* stream) {
  if (!a || !fg || !bg || !img || !scaled_img || !tri3 || !scratch || !imgn || !mean_std || imgn_ld % 4)
    return OTVM_ERR_ARG;
  MeanStd ms = make_ms(mean_std);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t P = (int64_t)H * W;
  uint8_t* unk = scratch; uint8_t* rows = scratch + P;
  launch_k(composite_kernel, grid1d(P, 256), 256, 0, s, a, fg, bg, H, W, Wp, pad_top, pad_left, img, scaled_img, unk);
  OTVM_LAUNCH_CHECK();
  launch_k(dilate_rows_kernel, grid1d(P, 256), 256, 0, s, unk, H, W, radius, rows);
  OTVM_LAUNCH_CHECK();
  DISPATCH_DTYPE(dtype, launch_k(dilate_cols_onehot_kernel<T>, grid1d((int64_t)Hp * Wp, 256), 256, 0, s, rows, a, H, W, Hp, Wp,
                                 pad_top, pad_left, radius, img, tri3, ms, mkptr<T>(imgn, ps), imgn_ld));
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_edt_sq(const uint8_t* seed, int32_t H, int32_t W, int32_t* d2, int32_t* scratch, void* stream) {
  if (!seed || !d2 || !scratch || W > 8192) return OTVM_ERR_ARG;
  return edt_launch(seed, H, W, 1, d2, scratch, static_cast<cudaStream_t>(stream));
}


extern "C" int otvm_trimap_encode(const float* tri_in, int64_t tri_ld, int32_t is_logit, const float* img, int32_t Hp,
                                  int32_t Wp, const float* mean_std, void* x11, int64_t x11_ld, void* cat_dst,
                                  int64_t cat_ld, int32_t dtype, float* extras, int32_t* d2, int32_t* scratch,
                                  uint8_t* seeds, void* stream) {
  if (!tri_in || !img || !x11 || !extras || !d2 || !scratch || !seeds || !mean_std || x11_ld < 16 || x11_ld % 4 ||
      Wp > 8192 || (cat_dst && cat_ld % 4))
    return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t P = (int64_t)Hp * Wp;
  launch_k(trimap_classes_kernel, grid1d(P, 256), 256, 0, s, tri_in, tri_ld, is_logit, img, P, extras, seeds);
  OTVM_LAUNCH_CHECK();
  int rc = edt_launch(seeds, Hp, Wp, 2, d2, scratch, s);
  if (rc) return rc;
  MeanStd ms = make_ms(mean_std);
  // 16-byte stores: split / bf16 outputs whose pixel rows start on 16-byte boundaries
  const bool wide = dtype_fmt(dtype) != OTVM_F32 && x11_ld % 8 == 0 && !(reinterpret_cast<uintptr_t>(x11) & 15) &&
                    (!cat_dst || (cat_ld % 8 == 0 && !(reinterpret_cast<uintptr_t>(cat_dst) & 15)));
  DISPATCH_DTYPE(dtype, launch_k(trimap_pack_kernel<T>, grid1d(P, 256), 256, 0, s, extras, d2, P, ms, mkptr<T>(x11, ps), x11_ld,
                                 mkptr<T>(cat_dst, ps), cat_ld, wide));
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_fba_head(const void* raw, int64_t raw_ld, int32_t dtype, int32_t raw_f32, const float* extras,
                             int64_t P, float* out7, void* alpha_dst, int64_t alpha_ld, void* stream) {
  if (!raw || !extras || !out7) return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int g = grid1d(P, 256);
  if (raw_f32 || dtype_fmt(dtype) == OTVM_F32)
    DISPATCH_DTYPE(dtype, (launch_k(fba_head_kernel<float, T>, g, 256, 0, s, (const float*)raw, raw_ld, extras, P, out7,
                                    mkptr<T>(alpha_dst, ps), alpha_ld)));
  else
    DISPATCH_DTYPE(dtype, (launch_k(fba_head_kernel<T, T>, g, 256, 0, s, mkcptr<T>(raw, ps), raw_ld, extras, P, out7,
                                    mkptr<T>(alpha_dst, ps), alpha_ld)));
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_head_conv_fba(const void* x, int64_t x_ld, int32_t dtype, const float* w, const float* bias, int32_t Cout,
                                  float* raw, int64_t raw_ld, const float* extras, int64_t P, float* out7, void* alpha_dst,
                                  int64_t alpha_ld, void* stream) {
  if (!x || !w || !raw || !extras || !out7 || Cout < 7 || Cout > 12 || raw_ld < Cout || raw_ld % 4 || raw_ld > 12 || x_ld % 8)
    return OTVM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(extras) |
       reinterpret_cast<uintptr_t>(out7)) & 15)
    return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int g = grid1d(P, 256);
  DISPATCH_DTYPE(dtype, launch_k(head_conv_fba_kernel<T>, g, 256, 0, s, mkcptr<T>(x, ps), x_ld, w, bias, Cout, raw, raw_ld, extras, P,
                                 out7, mkptr<T>(alpha_dst, ps), alpha_ld));
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

extern "C" int otvm_frame_outputs(const float* raw10, int64_t raw_ld, const float* fused, const void* hid,
                                  int64_t hid_ld, const float* extras, int32_t Hp, int32_t Wp, int32_t H, int32_t W,
                                  int32_t pad_top, int32_t pad_left, const float* mean_std, void* mem_in,
                                  int64_t mem_ld, int32_t dtype, float* alpha_out, float* trimap_out, void* stream) {
  if (!raw10 || !fused || !extras || !alpha_out || !trimap_out || !mean_std) return OTVM_ERR_ARG;
  if (mem_in && (!hid || mem_ld < 24 || mem_ld % 4 || hid_ld % 2)) return OTVM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MeanStd ms = make_ms(mean_std);
  int g = grid1d((int64_t)Hp * Wp, 256);
  const bool wide = mem_in && dtype_fmt(dtype) != OTVM_F32 && mem_ld % 8 == 0 && hid_ld % 8 == 0 &&
                    !((reinterpret_cast<uintptr_t>(mem_in) | reinterpret_cast<uintptr_t>(hid)) & 15);
  DISPATCH_DTYPE(dtype, launch_k(frame_outputs_kernel<T>, g, 256, 0, s, raw10, raw_ld, fused, mkcptr<T>(hid, ps), hid_ld, extras,
                                 Hp, Wp, H, W, pad_top, pad_left, ms, mkptr<T>(mem_in, ps), mem_ld, alpha_out, trimap_out, wide));
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}
