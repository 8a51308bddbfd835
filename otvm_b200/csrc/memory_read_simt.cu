// STM space-time memory read (reference models/trimap/STM.py:144-163), fp32 FFMA version, plus the
// log-sum-exp combine shared with the tcgen05 version.
//
//   O[q, :] = sum_m softmax_m( K[m,:].Q[q,:] / sqrt(De) ) * V[:, m]
//
// One CTA owns 64 queries x 128 value channels x one split of the memory axis and streams 64-key blocks:
// S = Q K^T (64x64x128) in registers, online softmax (running max / sum per query row, exp2 with the
// 1/sqrt(De)*log2(e) scale folded in), P through shared memory, O += P V (64x128x64).  The [THW x HW] affinity
// never exists in memory.  Partial (unnormalised O, m, l) per split go to the workspace; a small kernel merges.
#include <math_constants.h>
#include "common.cuh"

namespace otvm {

constexpr int RQ = 64;        // queries per CTA
constexpr int RK = 64;        // keys per block
constexpr int RDV = 128;      // value channels per CTA
constexpr int RDE = 128;      // key dimension (fixed by the architecture, STM.py:184-185)
constexpr int QP = RDE + 4;   // smem pitches (floats): +4 keeps 16 B alignment and spreads banks
constexpr int PP = RK + 4;

struct ReadArgs {
  const void* keys; const void* vals; int64_t ldv;
  const void* query; int64_t q_ld;
  int M, HW, Do;
  int nsplit, blocks_per_split;
  float scale_log2;
  float* o_part; float* ml_part;
  int64_t ps;                  // split formats: plane stride (elements)
};

template <typename T>
__global__ void __launch_bounds__(256) memory_read_simt_kernel(const ReadArgs a) {
  pdl_sync();                                  // PDL contract (common.cuh)
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                      // [RQ][QP]
  float* Ks = Qs + RQ * QP;              // [RK][QP]
  float* Ps = Ks + RK * QP;              // [RQ][PP]
  float* Vs = Ps + RQ * PP;              // [RDV][PP]   (channel-major like the bank)

  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int q0 = blockIdx.x * RQ, c0 = blockIdx.y * RDV, split = blockIdx.z;
  const cptr_t<T> keys = mkcptr<T>(a.keys, a.ps);
  const cptr_t<T> vals = mkcptr<T>(a.vals, a.ps);
  const cptr_t<T> query = mkcptr<T>(a.query, a.ps);

  // Q tile (zero rows beyond HW)
  for (int v = t; v < RQ * (RDE / 4); v += 256) {
    int r = v / (RDE / 4), k = (v % (RDE / 4)) * 4;
    float q[4] = {0.f, 0.f, 0.f, 0.f};
    if (q0 + r < a.HW) load4(query + ((int64_t)(q0 + r) * a.q_ld + k), q);
    *reinterpret_cast<float4*>(&Qs[r * QP + k]) = make_float4(q[0], q[1], q[2], q[3]);
  }

  float m_run[4], l_run[4], o[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -CUDART_INF_F; l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[i][c] = 0.f;
  }

  const int kb0 = split * a.blocks_per_split;
  const int nkb_total = (a.M + RK - 1) / RK;
  const int kb1 = min(kb0 + a.blocks_per_split, nkb_total);
  const bool vec_v = (a.ldv % 4 == 0) && aligned4(vals);

  for (int kb = kb0; kb < kb1; ++kb) {
    const int key0 = kb * RK;
    __syncthreads();                                   // previous iteration done with Ks / Vs / Ps
    for (int v = t; v < RK * (RDE / 4); v += 256) {
      int r = v / (RDE / 4), k = (v % (RDE / 4)) * 4;
      float q[4] = {0.f, 0.f, 0.f, 0.f};
      if (key0 + r < a.M) load4(keys + ((int64_t)(key0 + r) * RDE + k), q);
      *reinterpret_cast<float4*>(&Ks[r * QP + k]) = make_float4(q[0], q[1], q[2], q[3]);
    }
    for (int v = t; v < RDV * (RK / 4); v += 256) {
      int c = v / (RK / 4), j = (v % (RK / 4)) * 4;
      float q[4] = {0.f, 0.f, 0.f, 0.f};
      const cptr_t<T> src = vals + ((int64_t)(c0 + c) * a.ldv + key0 + j);
      if (vec_v && key0 + j + 3 < a.M) load4(src, q);
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (key0 + j + e < a.M) q[e] = ld1(src, e);
      }
      *reinterpret_cast<float4*>(&Vs[c * PP + j]) = make_float4(q[0], q[1], q[2], q[3]);
    }
    __syncthreads();

    // ---- S = Q K^T : rows ty+16*i, keys tx+16*j
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < RDE; k += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[(ty + 16 * i) * QP + k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(&Ks[(tx + 16 * j) * QP + k]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]); s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]); s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }
    // ---- online softmax over the memory axis
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (key0 + tx + 16 * j < a.M) ? s[i][j] * a.scale_log2 : -CUDART_INF_F;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);          // finite: every block holds at least one valid key
      alpha[i] = exp2f(m_run[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = exp2f(s[i][j] - m_new);
        rs += p;
        Ps[(ty + 16 * i) * PP + tx + 16 * j] = p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[i] = l_run[i] * alpha[i] + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < 8; ++c) o[i][c] *= alpha[i];
    }
    __syncthreads();
    // ---- O += P V : rows ty+16*i, channels tx+16*c
#pragma unroll 2
    for (int j = 0; j < RK; j += 4) {
      float4 pv[4], vv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(&Ps[(ty + 16 * i) * PP + j]);
#pragma unroll
      for (int c = 0; c < 8; ++c) vv[c] = *reinterpret_cast<const float4*>(&Vs[(tx + 16 * c) * PP + j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          o[i][c] = fmaf(pv[i].x, vv[c].x, o[i][c]); o[i][c] = fmaf(pv[i].y, vv[c].y, o[i][c]);
          o[i][c] = fmaf(pv[i].z, vv[c].z, o[i][c]); o[i][c] = fmaf(pv[i].w, vv[c].w, o[i][c]);
        }
    }
  }

  // ---- partial results
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty + 16 * i;
    if (q >= a.HW) continue;
    float* op = a.o_part + ((int64_t)split * a.HW + q) * a.Do + c0;
#pragma unroll
    for (int c = 0; c < 8; ++c) op[tx + 16 * c] = o[i][c];
    if (tx == 0 && blockIdx.y == 0) {
      a.ml_part[((int64_t)split * a.HW + q) * 2 + 0] = m_run[i];
      a.ml_part[((int64_t)split * a.HW + q) * 2 + 1] = l_run[i];
    }
  }
}

// merge the per-split partials:  out[q,c] = sum_s 2^(m_s - m*) O_s[q,c] / sum_s 2^(m_s - m*) l_s
template <typename T>
__global__ void __launch_bounds__(256) memory_read_combine_kernel(const float* __restrict__ o_part,
                                                                  const float* __restrict__ ml_part, int nsplit,
                                                                  int HW, int Do, ptr_t<T> out, int64_t out_ld,
                                                                  float* __restrict__ lse) {
  pdl_sync();                                  // PDL contract (common.cuh)
  const int c4n = Do >> 2;
  const int64_t total = (int64_t)HW * c4n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int q = (int)(idx / c4n), c = (int)(idx - (int64_t)q * c4n) * 4;
    float mstar = -CUDART_INF_F;
    for (int s = 0; s < nsplit; ++s) mstar = fmaxf(mstar, ml_part[((int64_t)s * HW + q) * 2]);
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, l = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      float w = exp2f(ml_part[((int64_t)s * HW + q) * 2] - mstar);
      l += w * ml_part[((int64_t)s * HW + q) * 2 + 1];
      float4 v = *reinterpret_cast<const float4*>(o_part + ((int64_t)s * HW + q) * Do + c);
      acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
    }
    const float inv = 1.f / l;
    if (lse && c == 0) lse[q] = mstar + log2f(l);        // log2-sum-exp of the scaled logits (for the backward)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] *= inv;
    store4(out + ((int64_t)q * out_ld + c), acc);
  }
}

// upper bound of read_pick_splits that is monotonic in M (workspace sizing)
int read_max_splits(int M, int HW, int Do, int rows_per_cta, int cols_per_cta, int keys_per_block) {
  int tiles = ceil_div(HW, rows_per_cta) * ceil_div(Do, cols_per_cta);
  int nkb = ceil_div(M, keys_per_block);
  int ns = ceil_div(2 * sm_count(), tiles);
  if (ns < 1) ns = 1;
  if (ns > nkb) ns = nkb;
  if (ns > 64) ns = 64;
  return ns;
}

int read_pick_splits(int M, int HW, int Do, int rows_per_cta, int cols_per_cta, int keys_per_block) {
  int tiles = ceil_div(HW, rows_per_cta) * ceil_div(Do, cols_per_cta);
  int nkb = ceil_div(M, keys_per_block);
  int want = ceil_div(2 * sm_count(), tiles);
  int ns = want < 1 ? 1 : want;
  if (ns > nkb) ns = nkb;
  if (ns > 64) ns = 64;
  int bps = ceil_div(nkb, ns);
  return ceil_div(nkb, bps);                   // no empty split
}

int read_combine(const otvm_read_params* p, int nsplit, cudaStream_t s) {
  float* o_part = static_cast<float*>(p->workspace);
  float* ml_part = o_part + (int64_t)nsplit * p->HW * p->Do;
  int64_t total = (int64_t)p->HW * (p->Do / 4);
  int g = ceil_div(total, 256);
  const int64_t ps = dtype_plane_stride(p->dtype);
  switch (dtype_fmt(p->dtype)) {
#define OTVM_COMBINE(T) launch_k(memory_read_combine_kernel<T>, g, 256, 0, s, o_part, ml_part, nsplit, p->HW, p->Do, mkptr<T>(p->out, ps), p->out_ld, p->lse)
    case OTVM_F32: OTVM_COMBINE(float); break;
    case OTVM_BF16: OTVM_COMBINE(bf16); break;
    case OTVM_BF16X2: OTVM_COMBINE(bx<2>); break;
    case OTVM_BF16X3: OTVM_COMBINE(bx<3>); break;
#undef OTVM_COMBINE
    default: return OTVM_ERR_ARG;
  }
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

template <typename T>
static int read_simt_t(const otvm_read_params* p, cudaStream_t s) {
  ReadArgs a;
  a.keys = p->keys; a.vals = p->vals; a.ldv = p->ldv; a.query = p->query; a.q_ld = p->q_ld;
  a.M = p->M; a.HW = p->HW; a.Do = p->Do;
  a.nsplit = read_pick_splits(p->M, p->HW, p->Do, RQ, RDV, RK);
  a.blocks_per_split = ceil_div(ceil_div(p->M, RK), a.nsplit);
  a.scale_log2 = (float)(1.4426950408889634 / sqrt((double)p->De));
  a.o_part = static_cast<float*>(p->workspace);
  a.ml_part = a.o_part + (int64_t)a.nsplit * p->HW * p->Do;
  a.ps = dtype_plane_stride(p->dtype);
  size_t smem = sizeof(float) * (RQ * QP + RK * QP + RQ * PP + RDV * PP);
  OTVM_CUDA_CHECK((ensure_dynamic_smem<memory_read_simt_kernel<T>>((int)smem)));
  dim3 grid(ceil_div(p->HW, RQ), p->Do / RDV, a.nsplit);
  launch_k(memory_read_simt_kernel<T>, grid, 256, smem, s, a);
  OTVM_LAUNCH_CHECK();
  return read_combine(p, a.nsplit, s);
}

int memory_read_simt(const otvm_read_params* p, cudaStream_t s) {
  if (p->De != RDE || p->Do % RDV != 0 || p->q_ld % 4 != 0 || p->out_ld % 4 != 0) return OTVM_ERR_UNSUPPORTED;
  switch (dtype_fmt(p->dtype)) {
    case OTVM_F32: return read_simt_t<float>(p, s);
    case OTVM_BF16: return read_simt_t<bf16>(p, s);
    case OTVM_BF16X2: return read_simt_t<bx<2>>(p, s);
    case OTVM_BF16X3: return read_simt_t<bx<3>>(p, s);
    default: return OTVM_ERR_ARG;
  }
}

}  // namespace otvm
