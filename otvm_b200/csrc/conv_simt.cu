// FFMA implicit-GEMM convolution (NHWC).  This is the strict-fp32 arithmetic path (1e-3 parity gate) and the
// fallback for shapes the tcgen05 kernel does not take (7x7 stride-2 stems with 3..22 input channels, PPM
// 1x1 convs on 1..36 pixels, stride-2 convs).  GEMM view: M = N*Ho*Wo output pixels, N = Cout,
// K = KH*KW*Cin ordered (ky, kx, ci) so that a K tile of 16 is one filter tap and 16 contiguous channels.
//
// Tiling: 64x64 output tile per 256-thread CTA, BK = 16, 4x4 register micro-tile per thread, double-buffered
// shared memory fed through registers (global loads of tile k+1 are in flight while tile k is multiplied).
#include "common.cuh"

namespace otvm {

struct ConvArgs {
  const void* in; int64_t in_ld;
  int N, H, W, Cin, Ho, Wo;
  const void* w; const float* bias;
  int Cout, KH, KW, stride, pad, dil, K;
  void* out; int64_t out_ps, out_cs;
  const void* res; int64_t res_ld;
  void* out_relu; int64_t out_relu_ld;
  int act, relu_in, out_f32;
  double* gn_stats;
  int64_t M;
  int64_t ps, wps;             // split formats: plane strides (elements) of activations / weights
};

constexpr int BM = 64, BN = 64, BK = 16, LDS_PAD = 4;

template <typename T, bool FAST>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvArgs a) {
  pdl_sync();                                  // PDL contract (common.cuh)
  __shared__ __align__(16) float As[2][BK][BM + LDS_PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + LDS_PAD];
  __shared__ double gsum[BN][2];   // fp64: sums of fp32 partials are exact -> order-independent

  const int t = threadIdx.x;
  const cptr_t<T> in = mkcptr<T>(a.in, a.ps);
  const cptr_t<T> w = mkcptr<T>(a.w, a.wps);
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- loader coordinates: each thread owns 4 consecutive k of one A row and one B row
  const int lrow = t >> 2, lk = (t & 3) * 4;
  const int64_t pm = m0 + lrow;
  const bool prow_ok = pm < a.M;
  int pn = 0, poy = 0, pox = 0;
  if (prow_ok) {
    int hw = a.Ho * a.Wo;
    pn = (int)(pm / hw);
    int r = (int)(pm - (int64_t)pn * hw);
    poy = r / a.Wo; pox = r - poy * a.Wo;
  }
  const int iy0 = poy * a.stride - a.pad, ix0 = pox * a.stride - a.pad;
  const int co_l = n0 + lrow;
  const bool co_ok = co_l < a.Cout;
  const cptr_t<T> wrow = w + (int64_t)(co_ok ? co_l : 0) * a.K;

  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { ra[i] = 0.f; rb[i] = 0.f; }
    if (FAST) {
      // one tap per K tile, 4 contiguous channels per thread
      int tap = k0 / a.Cin, c = k0 - tap * a.Cin + lk;
      int ky = tap / a.KW, kx = tap - ky * a.KW;
      int iy = iy0 + ky * a.dil, ix = ix0 + kx * a.dil;
      if (prow_ok && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
        load4(in + (((int64_t)(pn * a.H + iy) * a.W + ix) * a.in_ld + c), ra);
      if (co_ok) load4(wrow + (k0 + lk), rb);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int kk = k0 + lk + i;
        if (kk < a.K) {
          int tap = kk / a.Cin, c = kk - tap * a.Cin;
          int ky = tap / a.KW, kx = tap - ky * a.KW;
          int iy = iy0 + ky * a.dil, ix = ix0 + kx * a.dil;
          if (prow_ok && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
            ra[i] = ld1(in, ((int64_t)(pn * a.H + iy) * a.W + ix) * a.in_ld + c);
          if (co_ok) rb[i] = ld1(wrow, kk);
        }
      }
    }
    if (a.relu_in) {
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = fmaxf(ra[i], 0.f);
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[buf][lk + i][lrow] = ra[i]; Bs[buf][lk + i][lrow] = rb[i]; }
  };

  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (a.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: bias, GN statistics of the pre-activation output, residual, activation, stores
  const int c0 = n0 + tx * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (c0 + j < a.Cout) bias[j] = a.bias[c0 + j];
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec_ok = (c0 + 3 < a.Cout);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t p = m0 + ty * 4 + i;
    if (p >= a.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (a.gn_stats) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // statistics of what GroupNorm will read back: the value as stored (rounded to T)
        float q = stored<T>(v[j]);
        s1[j] += q; s2[j] += q * q;
      }
    }
    if (a.res) {
      const cptr_t<T> r = mkcptr<T>(a.res, a.ps) + (p * a.res_ld + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (c0 + j < a.Cout) v[j] += ld1(r, j);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], a.act);
    if (a.out_f32) {
      float* o = static_cast<float*>(a.out) + p * a.out_ps + (int64_t)c0 * a.out_cs;
#pragma unroll
      for (int j = 0; j < 4; ++j) if (c0 + j < a.Cout) o[(int64_t)j * a.out_cs] = v[j];
    } else {
      const ptr_t<T> o = mkptr<T>(a.out, a.ps) + (p * a.out_ps + (int64_t)c0 * a.out_cs);
      if (a.out_cs == 1 && vec_ok && aligned4(o)) store4(o, v);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c0 + j < a.Cout) st1(o, (int64_t)j * a.out_cs, v[j]);
      }
    }
    if (a.out_relu) {
      const ptr_t<T> o = mkptr<T>(a.out_relu, a.ps) + (p * a.out_relu_ld + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (c0 + j < a.Cout) st1(o, j, fmaxf(v[j], 0.f));
    }
  }
  if (a.gn_stats) {
    // per-channel partial sums of this 64-pixel x 64-channel tile -> one fp64 atomic per channel
    if (t < BN) { gsum[t][0] = 0.0; gsum[t][1] = 0.0; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&gsum[tx * 4 + j][0], (double)s1[j]); atomicAdd(&gsum[tx * 4 + j][1], (double)s2[j]); }
    __syncthreads();
    if (t < BN && n0 + t < a.Cout) {
      int g = (n0 + t) / (a.Cout / 32);
      atomicAdd(&a.gn_stats[g * 2 + 0], gsum[t][0]);
      atomicAdd(&a.gn_stats[g * 2 + 1], gsum[t][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// 1x1 convolution on a handful of pixels (the pyramid-pooling branches: 1..36 pixels x 2048 -> 256 channels,
// FBA/models.py:300-305).  A 64x64 GEMM tile would leave all but a few SMs idle and walk K serially; here one warp
// owns one output channel, the 32 lanes split K, and the weights are read exactly once.
template <typename T, int MP>
__global__ void __launch_bounds__(128) conv1x1_smallm_kernel(const ConvArgs a) {
  pdl_sync();                                  // PDL contract (common.cuh)
  // one CTA per output channel; its 4 warps split K, every lane keeps one accumulator per pixel (M <= MP), so the
  // weights are read once and all activation loads of a K step are independent (latency-bound otherwise)
  __shared__ float part[4][MP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.x;
  const cptr_t<T> in = mkcptr<T>(a.in, a.ps);
  const cptr_t<T> wr = mkcptr<T>(a.w, a.wps) + (int64_t)co * a.K;
  const int M = (int)a.M;
  float acc[MP];
#pragma unroll
  for (int i = 0; i < MP; ++i) acc[i] = 0.f;
  const int kq = (a.K + 3) / 4;                       // K range of this warp (multiple of 4 by the FAST contract)
  const int kend = min(a.K, (warp + 1) * kq);
  for (int k = warp * kq + lane * 4; k < kend; k += 128) {
    float wv[4];
    load4(wr + k, wv);
#pragma unroll
    for (int i = 0; i < MP; ++i) {
      if (i < M) {
        float xv[4];
        load4(in + ((int64_t)i * a.in_ld + k), xv);
        acc[i] = fmaf(wv[0], xv[0], acc[i]); acc[i] = fmaf(wv[1], xv[1], acc[i]);
        acc[i] = fmaf(wv[2], xv[2], acc[i]); acc[i] = fmaf(wv[3], xv[3], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MP; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) part[warp][i] = v;
  }
  __syncthreads();
  if (warp == 0) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < M; i += 32) {
      float v = part[0][i] + part[1][i] + part[2][i] + part[3][i] + (a.bias ? a.bias[co] : 0.f);
      const float qv = stored<T>(v);
      s1 += qv; s2 += qv * qv;
      v = apply_act(v, a.act);
      if (a.out_f32) static_cast<float*>(a.out)[(int64_t)i * a.out_ps + (int64_t)co * a.out_cs] = v;
      else st1(mkptr<T>(a.out, a.ps), (int64_t)i * a.out_ps + (int64_t)co * a.out_cs, v);
    }
    if (a.gn_stats) {
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) {
        const int g = co / (a.Cout / 32);
        atomicAdd(&a.gn_stats[g * 2 + 0], (double)s1);
        atomicAdd(&a.gn_stats[g * 2 + 1], (double)s2);
      }
    }
  }
}

template <typename T>
static int launch_conv_simt(const ConvArgs& a, cudaStream_t s) {
  dim3 grid(ceil_div(a.M, BM), ceil_div(a.Cout, BN));
  const cptr_t<T> in = mkcptr<T>(a.in, a.ps);
  const cptr_t<T> w = mkcptr<T>(a.w, a.wps);
  bool fast = (a.Cin % BK == 0) && (a.in_ld % 4 == 0) && aligned4(in) && aligned4(w);
  if (fast && a.KH == 1 && a.KW == 1 && a.stride == 1 && a.pad == 0 && a.M <= 40 && a.K % 16 == 0 && !a.res && !a.out_relu &&
      !a.relu_in) {
    if (a.M <= 4) launch_k(conv1x1_smallm_kernel<T, 4>, a.Cout, 128, 0, s, a);
    else if (a.M <= 12) launch_k(conv1x1_smallm_kernel<T, 12>, a.Cout, 128, 0, s, a);
    else launch_k(conv1x1_smallm_kernel<T, 40>, a.Cout, 128, 0, s, a);
    OTVM_LAUNCH_CHECK();
    return OTVM_OK;
  }
  if (fast) launch_k(conv_simt_kernel<T, true>, grid, 256, 0, s, a);
  else launch_k(conv_simt_kernel<T, false>, grid, 256, 0, s, a);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

int conv2d_simt(const otvm_conv_params* p, cudaStream_t s) {
  if (p->groups > 1) return OTVM_ERR_UNSUPPORTED;               // grouped convolutions: tcgen05 path only
  ConvArgs a;
  a.in = p->in; a.in_ld = p->in_ld; a.N = p->N; a.H = p->H; a.W = p->W; a.Cin = p->Cin;
  a.Ho = (p->H + 2 * p->pad - p->dil * (p->KH - 1) - 1) / p->stride + 1;
  a.Wo = (p->W + 2 * p->pad - p->dil * (p->KW - 1) - 1) / p->stride + 1;
  a.w = p->weight; a.bias = p->bias; a.Cout = p->Cout; a.KH = p->KH; a.KW = p->KW;
  a.stride = p->stride; a.pad = p->pad; a.dil = p->dil; a.K = p->KH * p->KW * p->Cin;
  a.out = p->out; a.out_ps = p->out_ps; a.out_cs = p->out_cs;
  a.res = p->res; a.res_ld = p->res_ld; a.out_relu = p->out_relu; a.out_relu_ld = p->out_relu_ld;
  a.act = p->act; a.relu_in = p->relu_in; a.out_f32 = p->out_f32 || dtype_fmt(p->dtype) == OTVM_F32;
  a.ps = dtype_plane_stride(p->dtype); a.wps = p->w_plane_stride;
  a.gn_stats = p->gn_stats;
  a.M = (int64_t)p->N * a.Ho * a.Wo;
  if (a.M <= 0 || a.Cout <= 0) return OTVM_OK;
  if (p->gn_stats) {
    if (p->N != 1 || p->Cout % 32 != 0) return OTVM_ERR_UNSUPPORTED;
    if (p->gn_group_ch > 0 && p->gn_group_ch != p->Cout / 32) return OTVM_ERR_UNSUPPORTED;   // channel slices: tcgen05 path only
    if (!p->gn_stats_zeroed) OTVM_CUDA_CHECK(cudaMemsetAsync(p->gn_stats, 0, sizeof(double) * 64, s));
  }
  // (out_f32 with bf16 / split inputs: the kernel's T covers in / weight / res, `out` is written as float)
  switch (dtype_fmt(p->dtype)) {
    case OTVM_F32: return launch_conv_simt<float>(a, s);
    case OTVM_BF16: return launch_conv_simt<bf16>(a, s);
    case OTVM_BF16X2: return launch_conv_simt<bx<2>>(a, s);
    case OTVM_BF16X3: return launch_conv_simt<bx<3>>(a, s);
    default: return OTVM_ERR_ARG;
  }
}

}  // namespace otvm
