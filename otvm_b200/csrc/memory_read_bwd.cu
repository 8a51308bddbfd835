// Backward of the STM space-time memory read (reference models/trimap/STM.py:144-163 under autograd; SURVEY.md section
// 8(f) rank 2: the stage-4 training step differentiates through Memory.forward, train.py:349-375).
//
//   forward   S[m,q] = K[m,:].Q[q,:] / sqrt(De)     P[:,q] = softmax_m S[:,q]     O[q,:] = sum_m P[m,q] V[:,m]
//   backward  dP[m,q] = V[:,m].dO[q,:]       delta[q] = O[q,:].dO[q,:]      dS = P o (dP - delta)
//             dQ[q,:] = sum_m dS[m,q] K[m,:] / sqrt(De)     dK[m,:] = sum_q dS[m,q] Q[q,:] / sqrt(De)
//             dV[:,m] = sum_q P[m,q] dO[q,:]
//
// Like the forward, the [THW x HW] affinity is never materialised: both kernels RECOMPUTE 32 x 32 tiles of S and dP from
// the operands and the per-query log-sum-exp the forward saved (otvm_read_params.lse), flash-attention style.
//   read_bwd_dq_kernel : one CTA = 32 queries, loops over the memory axis, keeps dQ in registers
//   read_bwd_dkv_kernel: one CTA = 32 memory locations, loops over the queries, keeps dK and dV in registers
// fp32 FFMA arithmetic (the training problem is small: 320x320 crops, HW = 400, T <= 3: ~1.5 GFLOP per step); operands
// may be fp32 or bf16, gradients are fp32.
#include <math_constants.h>
#include "common.cuh"

namespace otvm {

constexpr int BQ = 32, BKY = 32;          // queries / memory locations per tile
constexpr int DE = 128, DOC = 512;        // key and value channels (STM.py:184-185)
constexpr int QP2 = DE + 4;               // smem pitches (floats)
constexpr int OP2 = DOC + 4;
constexpr int VP2 = BKY + 1;
constexpr int TP2 = BKY + 1;

struct ReadBwdArgs {
  const void* keys; const void* vals; int64_t ldv;
  const void* query; int64_t q_ld;
  const float* out; int64_t out_ld;
  const float* dout; int64_t dout_ld;
  const float* lse;
  float* dkeys; float* dvals; int64_t dldv; float* dquery;
  int M, HW;
  float scale, scale_log2;
};

// shared-memory tiles (floats): Qs [BQ][QP2], Ks [BKY][QP2], dOs [BQ][OP2], Vs [DOC][VP2], Pt / dSt [BQ][TP2], dl / ls [BQ]
constexpr size_t kBwdSmem = sizeof(float) * (BQ * QP2 + BKY * QP2 + BQ * OP2 + DOC * VP2 + 2 * BQ * TP2 + 2 * BQ);

template <typename T>
__device__ __forceinline__ void load_q_tile(const ReadBwdArgs& a, int q0, float* Qs, float* dOs, float* dl, float* ls) {
  const cptr_t<T> query = mkcptr<T>(a.query, 0);
  const int t = threadIdx.x;
  for (int v = t; v < BQ * (DE / 4); v += 256) {
    const int r = v / (DE / 4), k = (v % (DE / 4)) * 4;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (q0 + r < a.HW) load4(query + ((int64_t)(q0 + r) * a.q_ld + k), x);
    *reinterpret_cast<float4*>(&Qs[r * QP2 + k]) = make_float4(x[0], x[1], x[2], x[3]);
  }
  for (int v = t; v < BQ * (DOC / 4); v += 256) {
    const int r = v / (DOC / 4), c = (v % (DOC / 4)) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < a.HW) x = *reinterpret_cast<const float4*>(a.dout + (int64_t)(q0 + r) * a.dout_ld + c);
    *reinterpret_cast<float4*>(&dOs[r * OP2 + c]) = x;
  }
  __syncthreads();
  // delta[q] = O[q,:] . dO[q,:]: 8 threads per query
  {
    const int r = t >> 3, g = t & 7;
    float acc = 0.f;
    if (q0 + r < a.HW) {
      const float* o = a.out + (int64_t)(q0 + r) * a.out_ld;
      for (int c = g * 4; c < DOC; c += 32) {
        const float4 ov = *reinterpret_cast<const float4*>(o + c);
        const float4 dv = *reinterpret_cast<const float4*>(&dOs[r * OP2 + c]);
        acc += ov.x * dv.x + ov.y * dv.y + ov.z * dv.z + ov.w * dv.w;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2); acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (g == 0) { dl[r] = acc; ls[r] = q0 + r < a.HW ? a.lse[q0 + r] : 0.f; }
  }
}

template <typename T>
__device__ __forceinline__ void load_k_tile(const ReadBwdArgs& a, int m0, float* Ks, float* Vs) {
  const cptr_t<T> keys = mkcptr<T>(a.keys, 0), vals = mkcptr<T>(a.vals, 0);
  const int t = threadIdx.x;
  for (int v = t; v < BKY * (DE / 4); v += 256) {
    const int r = v / (DE / 4), k = (v % (DE / 4)) * 4;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + r < a.M) load4(keys + ((int64_t)(m0 + r) * DE + k), x);
    *reinterpret_cast<float4*>(&Ks[r * QP2 + k]) = make_float4(x[0], x[1], x[2], x[3]);
  }
  for (int v = t; v < DOC * BKY; v += 256) {
    const int c = v / BKY, j = v % BKY;
    Vs[c * VP2 + j] = m0 + j < a.M ? ld1(vals, (int64_t)c * a.ldv + m0 + j) : 0.f;
  }
}

// P and dS of one 32 x 32 tile into shared memory: thread (qi = t / 8, mg = t % 8) owns keys mg + 8 j
__device__ __forceinline__ void tile_p_ds(const ReadBwdArgs& a, int q0, int m0, const float* Qs, const float* Ks,
                                          const float* dOs, const float* Vs, const float* dl, const float* ls, float* Pt,
                                          float* dSt) {
  const int t = threadIdx.x, qi = t >> 3, mg = t & 7;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int d = 0; d < DE; d += 4) {
    const float4 q = *reinterpret_cast<const float4*>(&Qs[qi * QP2 + d]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 k = *reinterpret_cast<const float4*>(&Ks[(mg + 8 * j) * QP2 + d]);
      s[j] = fmaf(q.x, k.x, s[j]); s[j] = fmaf(q.y, k.y, s[j]); s[j] = fmaf(q.z, k.z, s[j]); s[j] = fmaf(q.w, k.w, s[j]);
    }
  }
#pragma unroll 4
  for (int c = 0; c < DOC; ++c) {
    const float g = dOs[qi * OP2 + c];
#pragma unroll
    for (int j = 0; j < 4; ++j) dp[j] = fmaf(g, Vs[c * VP2 + mg + 8 * j], dp[j]);
  }
  const bool qok = q0 + qi < a.HW;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = mg + 8 * j;
    const bool ok = qok && m0 + m < a.M;
    const float p = ok ? exp2f(fmaf(s[j], a.scale_log2, -ls[qi])) : 0.f;
    Pt[qi * TP2 + m] = p;
    dSt[qi * TP2 + m] = p * (dp[j] - dl[qi]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) read_bwd_dq_kernel(const ReadBwdArgs a) {
  pdl_sync();                                  // PDL contract (common.cuh)
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm; float* Ks = Qs + BQ * QP2; float* dOs = Ks + BKY * QP2; float* Vs = dOs + BQ * OP2;
  float* Pt = Vs + DOC * VP2; float* dSt = Pt + BQ * TP2; float* dl = dSt + BQ * TP2; float* ls = dl + BQ;
  const int q0 = blockIdx.x * BQ, t = threadIdx.x, qi = t >> 3, mg = t & 7;
  load_q_tile<T>(a, q0, Qs, dOs, dl, ls);
  float dq[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) dq[k] = 0.f;
  for (int m0 = 0; m0 < a.M; m0 += BKY) {
    __syncthreads();                                         // previous tile fully consumed (also orders dl / ls)
    load_k_tile<T>(a, m0, Ks, Vs);
    __syncthreads();
    tile_p_ds(a, q0, m0, Qs, Ks, dOs, Vs, dl, ls, Pt, dSt);
    __syncthreads();
    // dQ[qi][d] += sum_m dS[qi][m] K[m][d], d = mg + 8 k
    for (int m = 0; m < BKY; ++m) {
      const float ds = dSt[qi * TP2 + m];
#pragma unroll
      for (int k = 0; k < 16; ++k) dq[k] = fmaf(ds, Ks[m * QP2 + mg + 8 * k], dq[k]);
    }
  }
  if (q0 + qi < a.HW) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a.dquery[(int64_t)(q0 + qi) * DE + mg + 8 * k] = dq[k] * a.scale;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) read_bwd_dkv_kernel(const ReadBwdArgs a) {
  pdl_sync();                                  // PDL contract (common.cuh)
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm; float* Ks = Qs + BQ * QP2; float* dOs = Ks + BKY * QP2; float* Vs = dOs + BQ * OP2;
  float* Pt = Vs + DOC * VP2; float* dSt = Pt + BQ * TP2; float* dl = dSt + BQ * TP2; float* ls = dl + BQ;
  const int m0 = blockIdx.x * BKY, t = threadIdx.x;
  load_k_tile<T>(a, m0, Ks, Vs);
  // dK[m][d]: thread owns m = t / 8, d = (t % 8) + 8 k (16 values); dV[c][m]: thread owns m = t % 32, c = t / 32 + 8 k (64)
  const int km = t >> 3, kd = t & 7, vm = t & 31, vc = t >> 5;
  float dk[16], dv[64];
#pragma unroll
  for (int k = 0; k < 16; ++k) dk[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 64; ++k) dv[k] = 0.f;
  for (int q0 = 0; q0 < a.HW; q0 += BQ) {
    __syncthreads();
    load_q_tile<T>(a, q0, Qs, dOs, dl, ls);
    __syncthreads();
    tile_p_ds(a, q0, m0, Qs, Ks, dOs, Vs, dl, ls, Pt, dSt);
    __syncthreads();
    for (int q = 0; q < BQ; ++q) {
      const float ds = dSt[q * TP2 + km];
#pragma unroll
      for (int k = 0; k < 16; ++k) dk[k] = fmaf(ds, Qs[q * QP2 + kd + 8 * k], dk[k]);
      const float p = Pt[q * TP2 + vm];
#pragma unroll
      for (int k = 0; k < 64; ++k) dv[k] = fmaf(p, dOs[q * OP2 + vc + 8 * k], dv[k]);
    }
  }
  if (m0 + km < a.M) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a.dkeys[(int64_t)(m0 + km) * DE + kd + 8 * k] = dk[k] * a.scale;
  }
  if (m0 + vm < a.M) {
#pragma unroll
    for (int k = 0; k < 64; ++k) a.dvals[(int64_t)(vc + 8 * k) * a.dldv + m0 + vm] = dv[k];
  }
}

template <typename T>
static int read_bwd_t(const ReadBwdArgs& a, cudaStream_t s) {
  OTVM_CUDA_CHECK((ensure_dynamic_smem<read_bwd_dq_kernel<T>>((int)kBwdSmem)));
  OTVM_CUDA_CHECK((ensure_dynamic_smem<read_bwd_dkv_kernel<T>>((int)kBwdSmem)));
  launch_k(read_bwd_dq_kernel<T>, ceil_div(a.HW, BQ), 256, kBwdSmem, s, a);
  OTVM_LAUNCH_CHECK();
  launch_k(read_bwd_dkv_kernel<T>, ceil_div(a.M, BKY), 256, kBwdSmem, s, a);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

}  // namespace otvm

using namespace otvm;

extern "C" int otvm_memory_read_backward(const otvm_read_bwd_params* p, void* stream) {
  if (!p || !p->keys || !p->vals || !p->query || !p->out || !p->dout || !p->lse || !p->dkeys || !p->dvals || !p->dquery)
    return OTVM_ERR_ARG;
  if (p->M <= 0 || p->HW <= 0 || p->De != DE || p->Do != DOC) return OTVM_ERR_UNSUPPORTED;
  if (p->q_ld % 4 || p->out_ld % 4 || p->dout_ld % 4) return OTVM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p->out) | reinterpret_cast<uintptr_t>(p->dout) | reinterpret_cast<uintptr_t>(p->query) |
       reinterpret_cast<uintptr_t>(p->keys)) & 15)
    return OTVM_ERR_ARG;
  ReadBwdArgs a;
  a.keys = p->keys; a.vals = p->vals; a.ldv = p->ldv; a.query = p->query; a.q_ld = p->q_ld;
  a.out = p->out; a.out_ld = p->out_ld; a.dout = p->dout; a.dout_ld = p->dout_ld; a.lse = p->lse;
  a.dkeys = p->dkeys; a.dvals = p->dvals; a.dldv = p->dldv; a.dquery = p->dquery;
  a.M = p->M; a.HW = p->HW;
  a.scale = (float)(1.0 / sqrt((double)p->De));
  a.scale_log2 = (float)(1.4426950408889634 / sqrt((double)p->De));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype_fmt(p->dtype)) {
    case OTVM_F32: return read_bwd_t<float>(a, s);
    case OTVM_BF16: return read_bwd_t<bf16>(a, s);
    default: return OTVM_ERR_UNSUPPORTED;
  }
}
