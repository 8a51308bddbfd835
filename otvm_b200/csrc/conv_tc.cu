// tcgen05 implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate in TMEM).
//
// GEMM view: M = output pixels, N = Cout, K = (ky, kx, ci).  One CTA computes a 128-pixel x BN-channel tile; the
// 128 pixels are a TH x TW rectangle of the output map, so for every filter tap the A operand is ONE 4-D TMA
// box {KC channels, TW, TH, 1} of the NHWC input at offset (kx*dil - pad, ky*dil - pad): TMA's out-of-bounds
// zero fill implements both the convolution padding and the ragged tile edge ("TMA im2col" without ever
// materialising im2col).  The box lands in shared memory as 128 rows of KC*2 bytes with the 128/64/32-byte
// swizzle, which is exactly the K-major canonical UMMA layout, so the MMA warp issues tcgen05.mma straight
// from it.  Weights [Cout][KH*KW*Cin] are the K-major B operand through a 2-D tensor map.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias / GroupNorm statistics / residual / activation -> global).
// Ring of NSTAGE {A,B} stages with full/empty mbarriers; tcgen05.commit releases stages and publishes the
// accumulator.  Up to two CTAs per SM (<= 113 KB smem, <= 128 TMEM columns each) so one CTA's epilogue overlaps
// another's main loop.
#include "tc_common.cuh"

namespace otvm {

using namespace tc;

struct ConvTcArgs {
  int N, H, W, Cin, Ho, Wo, Cout, KH, KW, pad, dil;
  int TW, TH, tiles_x, tiles_y;
  int KC, nchunk, nstage;
  uint32_t a_bytes, b_bytes, sbo, layout_type;
  const float* bias;
  void* out; int64_t out_ps, out_cs;
  const bf16* res; int64_t res_ld;
  bf16* out_relu; int64_t out_relu_ld;
  int act, out_f32;
  double* gn_stats;
};

constexpr int kConvThreads = 192;

template <int BN>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [stages x (A|B)] at 1024-byte alignment, then barriers
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage_bytes = a.a_bytes + a.b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstage * stage_bytes);
  uint64_t* empty_bar = full_bar + a.nstage;
  uint64_t* accum_bar = empty_bar + a.nstage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* sstat = reinterpret_cast<float*>(tmem_slot + 2);            // [BN][2] GroupNorm partials (per channel group)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  // tile coordinates
  int bx = blockIdx.x;
  const int tx_i = bx % a.tiles_x; bx /= a.tiles_x;
  const int ty_i = bx % a.tiles_y; const int n_img = bx / a.tiles_y;
  const int x0 = tx_i * a.TW, y0 = ty_i * a.TH;
  const int n0 = blockIdx.y * BN;
  const int num_k = a.KH * a.KW * a.nchunk;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < a.nstage; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < 2 * 128; i += kConvThreads) sstat[i] = 0.f;
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < num_k; ++it) {
        const int s = it % a.nstage;
        const uint32_t ph = (it / a.nstage) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = it / a.nchunk, chunk = it - tap * a.nchunk;
        const int ky = tap / a.KW, kx = tap - ky * a.KW;
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], a.a_bytes + (uint32_t)(BN * a.KC * 2));
        tma_load_4d(sa, &tmA, &full_bar[s], chunk * a.KC, x0 - a.pad + kx * a.dil, y0 - a.pad + ky * a.dil, n_img);
        tma_load_2d(sa + a.a_bytes, &tmB, &full_bar[s], tap * a.Cin + chunk * a.KC, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN);
      const int ksteps = a.KC / 16;
      for (int it = 0; it < num_k; ++it) {
        const int s = it % a.nstage;
        const uint32_t ph = (it / a.nstage) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_after_sync();
        const uint32_t sa = base + (uint32_t)s * stage_bytes;
        const uint64_t adesc = make_smem_desc(sa, a.sbo, a.layout_type);
        const uint64_t bdesc = make_smem_desc(sa + a.a_bytes, a.sbo, a.layout_type);
        for (int k = 0; k < ksteps; ++k)                     // +32 bytes (>>4 = 2) per UMMA_K=16 inside the swizzle span
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | k) != 0);
        umma_commit(&empty_bar[s]);                          // frees the stage when these MMAs have read it
      }
      umma_commit(accum_bar);                                // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                             // tile row = pixel
    const int ty = r / a.TW, tx = r - ty * a.TW;
    const int oy = y0 + ty, ox = x0 + tx;
    const bool valid = oy < a.Ho && ox < a.Wo;
    const int64_t pix = ((int64_t)n_img * a.Ho + oy) * a.Wo + ox;
    mbar_wait(accum_bar, 0);
    tcgen05_after_sync();
    const int cg = a.gn_stats ? a.Cout / 32 : 0;             // channels per GroupNorm group
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, raw);
      tmem_wait_ld();
      const int cbase = n0 + c;
      if (cbase >= a.Cout) break;                            // warp-uniform
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = __uint_as_float(raw[j]);
        if (a.bias && cbase + j < a.Cout) v[j] += __ldg(a.bias + cbase + j);
      }
      if (cg) {
        // statistics of the values GroupNorm will read back (rounded to bf16); rows outside the image count 0
        const int cgc = cg < 16 ? cg : 16, ngc = 16 / cgc;
        for (int gsub = 0; gsub < ngc; ++gsub) {
          float s1 = 0.f, s2 = 0.f;
          if (valid) {
            for (int j = gsub * cgc; j < (gsub + 1) * cgc; ++j) {
              float qv = __bfloat162float(__float2bfloat16_rn(v[j]));
              s1 += qv; s2 += qv * qv;
            }
          }
          s1 = warp_sum(s1); s2 = warp_sum(s2);
          if (lane == 0) {
            const int gl = (c + gsub * cgc) / cgc;            // local slot: one per cgc channels of this tile
            atomicAdd(&sstat[gl * 2 + 0], s1); atomicAdd(&sstat[gl * 2 + 1], s2);
          }
        }
      }
      if (valid) {
        if (a.res) {
          const bf16* rp = a.res + pix * a.res_ld + cbase;
          if (cbase + 15 < a.Cout && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
            uint4 r0 = *reinterpret_cast<const uint4*>(rp), r1 = *reinterpret_cast<const uint4*>(rp + 8);
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[2 * j] += __low2float(h0[j]); v[2 * j + 1] += __high2float(h0[j]);
              v[8 + 2 * j] += __low2float(h1[j]); v[9 + 2 * j] += __high2float(h1[j]);
            }
          } else {
            for (int j = 0; j < 16; ++j) if (cbase + j < a.Cout) v[j] += __bfloat162float(rp[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], a.act);
        if (a.out_f32) {
          float* op = static_cast<float*>(a.out) + pix * a.out_ps + (int64_t)cbase * a.out_cs;
          for (int j = 0; j < 16; ++j) if (cbase + j < a.Cout) op[(int64_t)j * a.out_cs] = v[j];
        } else {
          bf16* op = static_cast<bf16*>(a.out) + pix * a.out_ps + (int64_t)cbase * a.out_cs;
          if (a.out_cs == 1 && cbase + 15 < a.Cout && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
            for (int j = 0; j < 16; ++j) if (cbase + j < a.Cout) op[(int64_t)j * a.out_cs] = __float2bfloat16_rn(v[j]);
          }
        }
        if (a.out_relu) {
          bf16* op = a.out_relu + pix * a.out_relu_ld + cbase;
          if (cbase + 15 < a.Cout && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(v[2 * j], 0.f), fmaxf(v[2 * j + 1], 0.f));
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
            for (int j = 0; j < 16; ++j) if (cbase + j < a.Cout) op[j] = __float2bfloat16_rn(fmaxf(v[j], 0.f));
          }
        }
      }
    }
    if (cg) {
      asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps only
      const int cgc = cg < 16 ? cg : 16;
      const int e = threadIdx.x - 64;                         // 0..127
      if (e < BN / cgc && n0 + e * cgc < a.Cout) {
        const int g = (n0 + e * cgc) / cg;
        atomicAdd(&a.gn_stats[g * 2 + 0], (double)sstat[e * 2 + 0]);
        atomicAdd(&a.gn_stats[g * 2 + 1], (double)sstat[e * 2 + 1]);
      }
    }
    tcgen05_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
static int pick_bn(int Cout) { return Cout >= 128 ? 128 : Cout > 32 ? 64 : Cout > 16 ? 32 : 16; }

bool conv2d_tc_supported(const otvm_conv_params* p) {
  if (p->dtype != OTVM_BF16 || p->stride != 1 || p->relu_in) return false;
  if (p->Cin % 16 != 0 || p->in_ld % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p->in) & 15) || (reinterpret_cast<uintptr_t>(p->weight) & 15)) return false;
  const int Ho = p->H + 2 * p->pad - p->dil * (p->KH - 1), Wo = p->W + 2 * p->pad - p->dil * (p->KW - 1);
  if (Wo < 8 || Ho < 1 || (int64_t)Ho * Wo < 64) return false;
  if (((int64_t)p->KH * p->KW * p->Cin * 2) % 16 != 0) return false;
  if (p->gn_stats && (p->N != 1 || p->Cout % 32 != 0)) return false;
  const int bn = pick_bn(p->Cout);
  if (p->gn_stats) {            // a GroupNorm group must not straddle tiles / 16-column chunks irregularly
    const int cg = p->Cout / 32;
    if (cg > 16 && (cg % 16 != 0 || bn % cg != 0)) return false;
    if (cg <= 16 && 16 % cg != 0) return false;
  }
  static int sm100 = -1;
  if (sm100 < 0) { int dev = 0; cudaGetDevice(&dev); sm100 = otvm_device_is_sm100(dev); }
  return sm100 == 1;
}

template <int BN>
static int launch_conv_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvTcArgs& a, dim3 grid, size_t smem,
                          cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    OTVM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    attr = true;
  }
  conv_tc_kernel<BN><<<grid, kConvThreads, smem, s>>>(tmA, tmB, a);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

int conv2d_tc(const otvm_conv_params* p, cudaStream_t s) {
  ConvTcArgs a;
  a.N = p->N; a.H = p->H; a.W = p->W; a.Cin = p->Cin; a.Cout = p->Cout; a.KH = p->KH; a.KW = p->KW;
  a.pad = p->pad; a.dil = p->dil;
  a.Ho = p->H + 2 * p->pad - p->dil * (p->KH - 1); a.Wo = p->W + 2 * p->pad - p->dil * (p->KW - 1);
  int tw = 8; while (tw * 2 <= a.Wo && tw < 128) tw *= 2;
  a.TW = tw; a.TH = 128 / tw;
  a.tiles_x = ceil_div(a.Wo, a.TW); a.tiles_y = ceil_div(a.Ho, a.TH);
  a.KC = p->Cin % 64 == 0 ? 64 : p->Cin % 32 == 0 ? 32 : 16;
  a.nchunk = p->Cin / a.KC;
  const int bn = pick_bn(p->Cout);
  a.a_bytes = 128u * a.KC * 2;
  a.b_bytes = ((uint32_t)bn * a.KC * 2 + 1023u) & ~1023u;
  a.sbo = 8u * a.KC * 2;
  a.layout_type = a.KC == 64 ? 2u : a.KC == 32 ? 4u : 6u;
  const uint32_t stage = a.a_bytes + a.b_bytes;
  int nstage = (int)((96u * 1024u) / stage);
  if (nstage > 8) nstage = 8;
  const int num_k = a.KH * a.KW * a.nchunk;
  if (nstage > num_k) nstage = num_k < 2 ? 2 : num_k;
  a.nstage = nstage;
  a.bias = p->bias; a.out = p->out; a.out_ps = p->out_ps; a.out_cs = p->out_cs;
  a.res = static_cast<const bf16*>(p->res); a.res_ld = p->res_ld;
  a.out_relu = static_cast<bf16*>(p->out_relu); a.out_relu_ld = p->out_relu_ld;
  a.act = p->act; a.out_f32 = p->out_f32; a.gn_stats = p->gn_stats;
  if (p->gn_stats) OTVM_CUDA_CHECK(cudaMemsetAsync(p->gn_stats, 0, sizeof(double) * 64, s));

  const CUtensorMapSwizzle swz = a.KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : a.KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
    uint64_t str[3] = {(uint64_t)p->in_ld * 2, (uint64_t)p->W * p->in_ld * 2, (uint64_t)p->H * p->W * p->in_ld * 2};
    uint32_t box[4] = {(uint32_t)a.KC, (uint32_t)a.TW, (uint32_t)a.TH, 1};
    int rc = make_tmap_bf16(&tmA, p->in, 4, dims, str, box, swz);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)p->KH * p->KW * p->Cin;
    uint64_t dims[2] = {K, (uint64_t)p->Cout};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {(uint32_t)a.KC, (uint32_t)bn};
    int rc = make_tmap_bf16(&tmB, p->weight, 2, dims, str, box, swz);
    if (rc) return rc;
  }
  dim3 grid(a.tiles_x * a.tiles_y * p->N, ceil_div(p->Cout, bn));
  const size_t smem = (size_t)nstage * stage + 1024 + (2 * nstage + 1) * 8 + 16 + 2 * 128 * sizeof(float);
  switch (bn) {
    case 128: return launch_conv_tc<128>(tmA, tmB, a, grid, smem, s);
    case 64: return launch_conv_tc<64>(tmA, tmB, a, grid, smem, s);
    case 32: return launch_conv_tc<32>(tmA, tmB, a, grid, smem, s);
    default: return launch_conv_tc<16>(tmA, tmB, a, grid, smem, s);
  }
}

// ---------------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return OTVM_ERR_UNSUPPORTED;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OTVM_OK : OTVM_ERR_ARG;
}

}  // namespace otvm
