// Fused STM space-time memory read on tcgen05 tensor cores (reference models/trimap/STM.py:144-163).
//
//   O[q, :] = sum_m softmax_m( K[m,:].Q[q,:] / sqrt(128) ) * V[:, m]          (bf16 operands, fp32 accumulation)
//
// One CTA = 128 queries x 256 of the 512 value channels x one split of the memory axis, streaming 64-key blocks:
//   TMA      : Q tile once; per block a K tile [64 keys x 128] and a V tile [256 ch x 64 keys] (channel-major
//              bank => both are K-major UMMA operands), 3-stage mbarrier ring
//   MMA warp : S_j = Q K_j^T (M128 N64 K128) into one of two TMEM S buffers, then O += P_{j-1} V_{j-1}
//              (M128 N256 K64) -- S_j is issued BEFORE PV_{j-1} so the softmax of block j overlaps the PV MMA
//   softmax  : 4 warps, thread = query row: tcgen05.ld S, online softmax in the exp2 domain with LAZY rescaling
//              (the O accumulator in TMEM is only rescaled when the running max grows by more than 2^8),
//              P written as bf16 into a swizzled K-major smem tile that the PV MMA reads directly
// TMEM: O = 256 columns, S = 2 x 64 columns.  The [THW x HW] affinity never leaves the SM.  Per-split partial
// (unnormalised O, m, l) go to the fp32 workspace and are merged by memory_read_combine_kernel.
//
// Split-bf16 operands (NP = 2 planes, common.cuh): Q, K, V arrive as hi + lo planes and the probabilities are split
// the same way before they go to shared memory, so both contractions are three plane products each
// (Qh Kh + Qh Kl + Ql Kh;  Ph Vh + Ph Vl + Pl Vh) and the read is good to ~2^-16 instead of 2^-8: with attention
// logits of several hundred (random-init weights) bf16 keys alone move the probabilities by e^0.8.  The tiles are half
// as long along the memory axis (32-key blocks) so that two planes of everything still fit the same 224 KB.  A third
// plane of the inputs (strict mode) is ignored here: one operator at 2^-16 is far below the 1e-3 budget.
#include <math_constants.h>
#include "tc_common.cuh"

namespace otvm {

using namespace tc;

int read_pick_splits(int M, int HW, int Do, int rows_per_cta, int cols_per_cta, int keys_per_block);
int read_max_splits(int M, int HW, int Do, int rows_per_cta, int cols_per_cta, int keys_per_block);
int read_combine(const otvm_read_params* p, int nsplit, cudaStream_t s);

constexpr int TQ = 128, TDV = 256, TDE = 128;
constexpr int TNK = 2;                                // K ring: super-block slots
constexpr uint32_t kAlignSlack = 256;                 // the dynamic window is 1024-aligned in practice; checked at run time
// tile geometry for NP operand planes: blocks of KB keys (one softmax half), super-blocks of 2 KB keys (one S product)
template <int NP> struct ReadCfg {
  static constexpr int KB = NP == 1 ? 64 : 32;
  static constexpr int TNV = NP == 1 ? 3 : 2;                        // V ring: KB-key slots
  static constexpr uint32_t kQPlane = TQ * TDE * 2;                  // 32 KB (two 64-wide swizzle atoms)
  static constexpr uint32_t kKPlane = 2 * KB * TDE * 2;              // one super-block, two 64-dim atoms
  static constexpr uint32_t kVPlane = TDV * KB * 2;                  // 256 channels x KB keys
  static constexpr uint32_t kPPlane = TQ * KB * 2;
  static constexpr uint32_t kQBytes = NP * kQPlane, kKBytes = NP * kKPlane, kVBytes = NP * kVPlane, kPBytes = NP * kPPlane;
  static constexpr uint32_t kOffK = kQBytes, kOffV = kOffK + TNK * kKBytes, kOffP = kOffV + TNV * kVBytes;
  static constexpr uint32_t kOffAux = kOffP + 2 * kPBytes;           // 224 KB: barriers (256 B) + row-max exchange (2 KB)
  static constexpr uint32_t kSmem = kOffAux + 256 + 8 * TQ * 2 + kAlignSlack;   // barriers + bf16 row-max exchange
  static constexpr uint32_t kRowB = KB * 2;                          // bytes per V / P tile row (128: SWIZZLE_128B, 64: 64B)
  static constexpr uint32_t kSwzMask = kRowB == 128 ? 7u : 3u, kLayout = kRowB == 128 ? 2u : 4u, kSbo = 8 * kRowB;
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
  static_assert(kOffV - kOffK + TNV * kVBytes >= 128 * 1024, "the partial-O staging tiles reuse the K / V rings");
};
constexpr int TKB = ReadCfg<1>::KB;                   // (block size the split heuristics are quoted in)
constexpr int kReadThreads = 608;        // TMA warp, S-MMA warp, 16 softmax warps, PV-MMA warp
constexpr float kLazyLog2 = 8.f;

// 2^x on the SFU (ex2.approx.ftz): relative error ~2^-22, far below the bf16 rounding of P
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct ReadTcArgs {
  int M, HW, Do, nsplit, blocks_per_split;
  float scale_log2;
  float* o_part; float* ml_part;
  long long* dbg;             // dev: per-CTA clock64 timestamps (NULL in production)
};

// Pipeline (per CTA: 128 queries x 256 value channels x one slice of the memory axis):
//   super-block J = keys [128J, 128J+128): S_J = Q K_J^T is ONE N=128 product (an N=64 tcgen05.mma costs the same
//   64 cycles as N=128, measured) into TMEM S buffer J&1;  blocks j = 2J, 2J+1 (64 keys each): O += P_j V_j.
//   softmax: warps 2-5 (half 0) own block 2J, warps 6-9 (half 1) own block 2J+1 of the SAME 128 rows (warps w and
//   w+4 share a TMEM lane quarter); the pair exchanges the row maximum once per super-block, so the barrier /
//   TMEM / fence latencies of the softmax are paid per 128 keys while its MUFU work (64 ex2 per thread) overlaps
//   the partner's.  P_j goes to shared-memory buffer j&1 (= half), written by that half's 4 warps only.
template <int NP>
__global__ void __launch_bounds__(kReadThreads, 1) memory_read_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                         const __grid_constant__ CUtensorMap tmK,
                                                                         const __grid_constant__ CUtensorMap tmV,
                                                                         const __grid_constant__ CUtensorMap tmO,
                                                                         const ReadTcArgs a) {
  typedef ReadCfg<NP> Cfg;
  constexpr int KB = Cfg::KB, TNV = Cfg::TNV, KQ = KB / 2;           // KQ: keys per softmax thread and super-block
  constexpr uint32_t kQBytes = Cfg::kQBytes, kKBytes = Cfg::kKBytes, kVBytes = Cfg::kVBytes, kPBytes = Cfg::kPBytes;
  constexpr uint32_t kOffK = Cfg::kOffK, kOffV = Cfg::kOffV, kOffP = Cfg::kOffP, kOffAux = Cfg::kOffAux;
  constexpr int NPAIR = NP == 1 ? 1 : 3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (base - smem_u32(smem_raw) > kAlignSlack) __trap();
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffAux);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // TNK
  uint64_t* k_empty = k_full + TNK;        // TNK
  uint64_t* v_full = k_empty + TNK;        // TNV
  uint64_t* v_empty = v_full + TNV;        // TNV
  uint64_t* s_full = v_empty + TNV;        // 2
  uint64_t* s_empty = s_full + 2;          // 2
  uint64_t* p_full = s_empty + 2;          // 2
  uint64_t* p_empty = p_full + 2;          // 2
  uint64_t* o_done = p_empty + 2;          // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);
  float* xmax = reinterpret_cast<float*>(smem + kOffAux + 256);        // [4 subs][128 rows] fp32 row-max exchange, 2 KB
  float* xsum = reinterpret_cast<float*>(smem + kOffAux + 256);        // the same bytes after the loop: [3][128] row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = a.dbg ? a.dbg + (size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  const int q0 = blockIdx.x * TQ, c0 = blockIdx.y * TDV, split = blockIdx.z;
  const int nb_total = (a.M + KB - 1) / KB;
  const int kb0 = split * a.blocks_per_split;
  const int nb = min(a.blocks_per_split, nb_total - kb0);          // KB-key blocks of this CTA
  const int nsb = (nb + 1) >> 1;                                   // super-blocks (2 KB keys)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < TNK; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < TNV; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 16);
      mbar_init(&p_full[b], 8); mbar_init(&p_empty[b], 1);
    }
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base, tmem_s = tmem_base + TDV;     // S buffers: 2 x (2 KB) columns
  pdl_trigger();                                       // PDL contract (common.cuh): resources held, then wait
  pdl_wait();
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      for (int pl = 0; pl < NP; ++pl) {
        tma_load_3d(sQ + pl * Cfg::kQPlane, &tmQ, q_full, 0, q0, pl);
        tma_load_3d(sQ + pl * Cfg::kQPlane + Cfg::kQPlane / 2, &tmQ, q_full, 64, q0, pl);
      }
      auto load_k = [&](int J) {                           // 2 KB keys x 128 dims (rows beyond M are zero-filled)
        const int s = J % TNK;
        mbar_wait(&k_empty[s], ((J / TNK) & 1) ^ 1);
        uint8_t* st = sK + (size_t)s * kKBytes;
        const int key0 = (kb0 + 2 * J) * KB;
        mbar_arrive_expect_tx(&k_full[s], kKBytes);
        for (int pl = 0; pl < NP; ++pl) {
          tma_load_3d(st + pl * Cfg::kKPlane, &tmK, &k_full[s], 0, key0, pl);
          tma_load_3d(st + pl * Cfg::kKPlane + Cfg::kKPlane / 2, &tmK, &k_full[s], 64, key0, pl);
        }
      };
      load_k(0);
      for (int j = 0; j < nb; ++j) {                       // issue order K_0, K_1, V_0, V_1, K_2, V_2, V_3, ...
        if ((j & 1) == 0 && (j >> 1) + 1 < nsb) load_k((j >> 1) + 1);
        const int s = j % TNV;
        mbar_wait(&v_empty[s], ((j / TNV) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[s], kVBytes);
        for (int pl = 0; pl < NP; ++pl)
          tma_load_3d(sV + (size_t)s * kVBytes + pl * Cfg::kVPlane, &tmV, &v_full[s], (kb0 + j) * KB, c0, pl);
      }
    }
  } else if (warp == 1) {
    // ===== S issuer: S_J = Q K_J^T (M128 N128 K128) into TMEM S buffer J & 1 =====
    // Two issuing threads (this one and the PV issuer, warp 18): a single thread walking wait -> issue -> commit for
    // both products was the pipeline's critical path (measured ~100 cycles per mbarrier operation).
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 2 * KB);
      mbar_wait(q_full, 0);
      for (int J = 0; J < nsb; ++J) {
        const int s = J % TNK, sb = J & 1;
        if (J >= 2) mbar_wait(&s_empty[sb], ((J >> 1) - 1) & 1);     // softmax finished reading S of super-block J-2
        mbar_wait(&k_full[s], (J / TNK) & 1);
        tcgen05_after_sync();
        if (dbg && J < 12) dbg[32 + J] = clock64();
        const uint32_t k_addr = base + kOffK + (uint32_t)s * kKBytes;
#pragma unroll
        for (int pr = 0; pr < NPAIR; ++pr) {               // plane products (hi hi | + hi lo + lo hi)
          const uint32_t qp = base + (pr == 2 ? Cfg::kQPlane : 0u), kp = k_addr + (pr == 1 ? Cfg::kKPlane : 0u);
#pragma unroll
          for (int k = 0; k < TDE / 16; ++k) {             // two 64-dim atoms, 4 K-steps each
            const uint32_t atom = (k >> 2), kk = (k & 3);
            const uint64_t qd = make_smem_desc(qp + atom * (Cfg::kQPlane / 2), 1024, 2) + (uint64_t)(2 * kk);
            const uint64_t kd = make_smem_desc(kp + atom * (Cfg::kKPlane / 2), 1024, 2) + (uint64_t)(2 * kk);
            umma_bf16(tmem_s + (uint32_t)sb * (2 * KB), qd, kd, idesc_s, (pr | k) != 0);
          }
        }
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[s]);                          // K slot free once S_J has been computed
      }
    }
  } else if (warp == 18) {
    // ===== PV issuer: O += P_j V_j (M128 N256 K64) =====
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(128, TDV);
      for (int j = 0; j < nb; ++j) {
        const int s = j % TNV, b = j & 1;
        mbar_wait(&v_full[s], (j / TNV) & 1);
        mbar_wait(&p_full[b], (j >> 1) & 1);
        tcgen05_after_sync();
        if (dbg && j < 12) dbg[44 + j] = clock64();
#pragma unroll
        for (int pr = 0; pr < NPAIR; ++pr) {
          const uint64_t pdesc = make_smem_desc(base + kOffP + (uint32_t)b * kPBytes + (pr == 2 ? Cfg::kPPlane : 0u), Cfg::kSbo, Cfg::kLayout);
          const uint64_t vdesc = make_smem_desc(base + kOffV + (uint32_t)s * kVBytes + (pr == 1 ? Cfg::kVPlane : 0u), Cfg::kSbo, Cfg::kLayout);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)
            umma_bf16(tmem_o, pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_o, (j | pr | k) != 0);
        }
        umma_commit(&p_empty[b]);                          // P buffer free, O holds blocks 0..j
        umma_commit(&v_empty[s]);                          // V slot free
      }
      umma_commit(o_done);
    }
  } else {
    // ===== softmax / correction / epilogue: 16 warps = FOUR threads per query row =====
    // warps w, w+4, w+8, w+12 (w = 2..5) share TMEM lane quarter w % 4.  sub = 0..3: half = sub >> 1 selects the
    // 64-key block of the super-block (2J or 2J+1), quarter = sub & 1 the 32 keys of that block the thread owns.
    // Four warps per scheduler hide each other's TMEM / barrier / fence latencies (the MUFU ex2 work, 1024 cycles per
    // super-block and scheduler, is the floor); the four threads of a row agree on the row maximum once per 128 keys.
    const int qd = warp & 3, sub = (warp - 2) >> 2, half = sub >> 1, quarter = sub & 1;
    const int r = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float m_used = -CUDART_INF_F, l_sum = 0.f;
    const float scale = a.scale_log2;
    for (int J = 0; J < nsb; ++J) {
      const int sb = J & 1, j = 2 * J + half;
      const bool has = j < nb;                             // (only the last super-block can lack its second block)
      mbar_wait(&s_full[sb], (J >> 1) & 1);
      tcgen05_after_sync();
      if (dbg && threadIdx.x == 64 && J < 12) dbg[8 + J] = clock64();
      uint32_t raw[KQ];
      if constexpr (KQ == 32) tmem_ld32(tmem_s + lane_base + (uint32_t)(sb * 2 * KB + sub * KQ), raw);
      else tmem_ld16(tmem_s + lane_base + (uint32_t)(sb * 2 * KB + sub * KQ), raw);
      tmem_wait_ld();
      if (dbg && threadIdx.x == 64 && J == 4) dbg[56] = clock64();
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);            // 16 arrivals: S buffer may be overwritten by super-block J+2
      const int key0 = (kb0 + j) * KB + quarter * KQ;
      if (!has || key0 + KQ > a.M) {                       // absent block / ragged last block (warp-uniform branch)
#pragma unroll
        for (int i = 0; i < KQ; ++i) if (!has || key0 + i >= a.M) raw[i] = 0xff800000u;    // -inf
      }
      float mx;
      {   // 4 independent chains instead of one 32-deep dependent FMNMX chain
        float m4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) m4[i] = __uint_as_float(raw[i]);
#pragma unroll
        for (int i = 4; i < KQ; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(raw[i]));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      // The four threads of a row exchange their maxima in fp32 through ONE buffer (2 KB is all that is left beside
      // 224 KB of tiles), so a second barrier keeps a fast warp from overwriting it for the next super-block.  (Round 1
      // exchanged bf16 values rounded up in two buffers: with logits of 1e5 and more -- undamped random weights -- the
      // 2^-8 slack pushed every probability of a row below 2^-126 and the row came out 0/0.)
      float* xrow = xmax + r;                              // [sub][row]
      xrow[sub * TQ] = mx;
      asm volatile("bar.sync %0, 128;" ::"r"(2 + qd) : "memory");         // the four warps of this lane quarter
      mx = fmaxf(fmaxf(xrow[0], xrow[TQ]), fmaxf(xrow[2 * TQ], xrow[3 * TQ])) * scale;            // scale > 0
      asm volatile("bar.sync %0, 128;" ::"r"(2 + qd) : "memory");
      if (dbg && threadIdx.x == 64 && J == 4) dbg[57] = clock64();
      // lazy rescale: keep the stale max unless it is exceeded by more than 2^8 (p stays <= 256, exact in fp32 sums)
      const bool grow = mx > m_used + kLazyLog2;
      if (J == 0) {
        m_used = mx;                                       // O not written yet: nothing to rescale
      } else if (__any_sync(0xffffffffu, grow)) {          // the four warps of a quarter see the same rows -> same decision
        const float m_new = grow ? mx : m_used;
        const float alpha = exp2f(m_used - m_new);
        mbar_wait(&p_empty[1], (J - 1) & 1);               // PV of block 2J-1 (and, in issue order, all earlier) complete
        tcgen05_after_sync();
#pragma unroll 1
        for (int c = sub * (TDV / 4); c < (sub + 1) * (TDV / 4); c += 32) {       // each warp rescales a quarter of the columns
          uint32_t o[32];
          tmem_ld32(tmem_o + lane_base + (uint32_t)c, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_o + lane_base + (uint32_t)c, o);
        }
        tmem_wait_st();
        tcgen05_before_sync();
        // P_j is published by the two warps of its block alone: none may release its PV product before all four
        // warps of this lane quarter have finished rescaling these rows
        asm volatile("bar.sync %0, 128;" ::"r"(2 + qd) : "memory");
        l_sum *= alpha;
        m_used = m_new;
      }
      if (has) {
        float l4[4] = {0.f, 0.f, 0.f, 0.f};                // independent partial sums
        uint32_t pk[KQ / 2], pk2[NP == 2 ? KQ / 2 : 1];
#pragma unroll
        for (int e = 0; e < KQ / 2; ++e) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(raw[2 * e]), scale, -m_used));
          const float p1 = fast_exp2(fmaf(__uint_as_float(raw[2 * e + 1]), scale, -m_used));
          // bf16 by truncation with integer ops: the F2FP conversion shares the 16-lane XU pipe with MUFU.EX2, which
          // paces this loop; the row sum uses the SAME truncated weights the MMA sees, so the normalisation is exact
          const uint32_t u0 = __float_as_uint(p0) & 0xffff0000u, u1 = __float_as_uint(p1) & 0xffff0000u;
          pk[e] = __byte_perm(u0, u1, 0x7632);             // low half = bf16(p0), high half = bf16(p1)
          if constexpr (NP == 2) {
            // second plane: the (exact) remainder, truncated the same way; the row sum takes hi + lo
            const uint32_t w0 = __float_as_uint(p0 - __uint_as_float(u0)) & 0xffff0000u;
            const uint32_t w1 = __float_as_uint(p1 - __uint_as_float(u1)) & 0xffff0000u;
            pk2[e] = __byte_perm(w0, w1, 0x7632);
            l4[e & 3] += (__uint_as_float(u0) + __uint_as_float(w0)) + (__uint_as_float(u1) + __uint_as_float(w1));
          } else {
            l4[e & 3] += __uint_as_float(u0) + __uint_as_float(u1);
          }
        }
        l_sum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        if (dbg && threadIdx.x == 64 && J == 4) dbg[58] = clock64();
        if (J >= 1) mbar_wait(&p_empty[half], (J - 1) & 1);     // PV of block j-2 finished reading P[half]
        if (dbg && threadIdx.x == 64 && J == 4) dbg[59] = clock64();
        uint8_t* prow = sP + (size_t)half * kPBytes;
        constexpr int NCH = KQ / 8;                         // 16-byte pieces of this thread's half row
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t off = (uint32_t)r * Cfg::kRowB + (uint32_t)(quarter * NCH + ch) * 16u;
          off ^= ((off >> 7) & Cfg::kSwzMask) << 4;
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
          if constexpr (NP == 2)
            *reinterpret_cast<uint4*>(prow + Cfg::kPPlane + off) = make_uint4(pk2[4 * ch], pk2[4 * ch + 1], pk2[4 * ch + 2], pk2[4 * ch + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[half]);         // 8 arrivals (this block's warps)
      }
      if (dbg && threadIdx.x == 64 && J < 12) dbg[20 + J] = clock64();
    }
    // ---- partial results of this split
    mbar_wait(o_done, 0);
    tcgen05_after_sync();
    if (dbg && threadIdx.x == 64) dbg[2] = clock64();
    const int q = q0 + r;
    asm volatile("bar.sync 1, 512;" ::: "memory");         // every warp has read the last row maxima: reuse the buffer
    if (sub != 0) xsum[(sub - 1) * TQ + r] = l_sum;
    // partial O: TMEM -> swizzled fp32 tiles [8 chunks][128 rows][32 floats] in the drained K/V rings -> TMA store
    // (per-thread row stores would scatter 16-byte pieces over 32 rows per instruction); each of the four warps of
    // a lane quarter drains 64 of the 256 columns
    uint8_t* stage = sK;                                   // 128 KB behind Q: K ring + first two V slots
#pragma unroll 1
    for (int c = sub * (TDV / 4); c < (sub + 1) * (TDV / 4); c += 32) {
      uint32_t o[32];
      tmem_ld32(tmem_o + lane_base + (uint32_t)c, o);
      tmem_wait_ld();
      uint8_t* tile = stage + (size_t)(c >> 5) * (TQ * 128);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t off = (uint32_t)r * 128u + (uint32_t)i * 16u;
        off ^= ((off >> 7) & 7u) << 4;
        *reinterpret_cast<uint4*>(tile + off) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x == 64) {
#pragma unroll 1
      for (int c = 0; c < TDV; c += 32) tma_store_3d(&tmO, stage + (size_t)(c >> 5) * (TQ * 128), c0 + c, q0, split);
      tma_store_commit_and_wait();
    }
    if (sub == 0 && q < a.HW && blockIdx.y == 0) {
      a.ml_part[((int64_t)split * a.HW + q) * 2 + 0] = m_used;
      a.ml_part[((int64_t)split * a.HW + q) * 2 + 1] = l_sum + xsum[r] + xsum[TQ + r] + xsum[2 * TQ + r];
    }
    if (dbg && threadIdx.x == 64) dbg[3] = clock64();
    tcgen05_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// one CTA per SM (224 KB of shared memory): split the memory axis so that the grid is one wave; kb = keys per block
static int read_tc_splits(int M, int HW, int Do, int kb) {
  const int tiles = ceil_div(HW, TQ) * (Do / TDV);
  const int nkb = ceil_div(M, kb);
  int ns = sm_count() / tiles;
  if (ns < 1) ns = 1;
  if (ns > nkb) ns = nkb;
  if (ns > 64) ns = 64;
  int bps = ceil_div(nkb, ns);
  if (bps > 1 && (bps & 1)) ++bps;              // whole super-blocks per slice
  return ceil_div(nkb, bps);
}

long long* g_read_dbg = nullptr;   // dev hook (otvm_debug_set_read_timestamps)

bool memory_read_tc_supported(const otvm_read_params* p) {
  const int fmt = dtype_fmt(p->dtype);
  if ((fmt != OTVM_BF16 && fmt != OTVM_BF16X2 && fmt != OTVM_BF16X3) || p->De != TDE || p->Do % TDV != 0) return false;
  if (p->q_ld % 8 != 0 || p->ldv % 8 != 0 || p->out_ld % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p->keys) & 15) || (reinterpret_cast<uintptr_t>(p->vals) & 15) ||
      (reinterpret_cast<uintptr_t>(p->query) & 15))
    return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  return otvm_device_is_sm100(dev) == 1;
}

int64_t memory_read_tc_workspace(int M, int HW, int De, int Do) {
  (void)De;
  // upper bound of read_tc_splits over both block sizes that is monotonic in M (callers size the workspace for the
  // bank CAPACITY and read fewer frames): the split count before it is rounded to whole super-blocks
  const int tiles = ceil_div(HW, TQ) * (Do / TDV);
  int64_t ns = sm_count() / tiles;
  if (ns < 1) ns = 1;
  const int64_t nkb = ceil_div(M, ReadCfg<2>::KB);
  if (ns > nkb) ns = nkb;
  if (ns > 64) ns = 64;
  { int64_t nsm = read_max_splits(M, HW, Do, TQ, TDV, TKB); if (nsm > ns) ns = nsm; }
  return ns * HW * ((int64_t)Do + 2) * (int64_t)sizeof(float);
}

template <int NP>
static int memory_read_tc_np(const otvm_read_params* p, cudaStream_t s) {
  typedef ReadCfg<NP> Cfg;
  constexpr int KB = Cfg::KB;
  ReadTcArgs a;
  a.M = p->M; a.HW = p->HW; a.Do = p->Do;
  a.nsplit = read_tc_splits(p->M, p->HW, p->Do, KB);
  a.blocks_per_split = ceil_div(ceil_div(p->M, KB), a.nsplit);
  if (a.blocks_per_split > 1 && (a.blocks_per_split & 1)) ++a.blocks_per_split;
  a.scale_log2 = (float)(1.4426950408889634 / sqrt((double)p->De));
  a.o_part = static_cast<float*>(p->workspace);
  a.ml_part = a.o_part + (int64_t)a.nsplit * p->HW * p->Do;
  a.dbg = g_read_dbg;
  // (the operand planes are the outermost tensor-map dimension; a third plane of the inputs is not read)
  const uint64_t pb = NP > 1 ? (uint64_t)dtype_plane_stride(p->dtype) * 2 : 0;
  const CUtensorMapSwizzle vswz = Cfg::kRowB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)TDE, (uint64_t)p->HW, (uint64_t)NP};
    uint64_t str[2] = {(uint64_t)p->q_ld * 2, NP > 1 ? pb : (uint64_t)p->HW * p->q_ld * 2};
    uint32_t box[3] = {64, TQ, 1};
    int rc = make_tmap_bf16(&tmQ, p->query, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)TDE, (uint64_t)p->M, (uint64_t)NP};
    uint64_t str[2] = {(uint64_t)TDE * 2, NP > 1 ? pb : (uint64_t)p->M * TDE * 2};
    uint32_t box[3] = {64, 2 * KB, 1};
    int rc = make_tmap_bf16(&tmK, p->keys, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p->M, (uint64_t)p->Do, (uint64_t)NP};
    uint64_t str[2] = {(uint64_t)p->ldv * 2, NP > 1 ? pb : (uint64_t)p->Do * p->ldv * 2};
    uint32_t box[3] = {KB, TDV, 1};
    int rc = make_tmap_bf16(&tmV, p->vals, 3, dims, str, box, vswz); if (rc) return rc;
  }
  OTVM_CUDA_CHECK((ensure_dynamic_smem<memory_read_tc_kernel<NP>>((int)Cfg::kSmem)));
  dim3 grid(ceil_div(p->HW, TQ), p->Do / TDV, a.nsplit);
  CUtensorMap tmO;
  {   // fp32 partial outputs [nsplit][HW][Do]: 32-float (128-byte) rows, 128-row boxes, clipped at HW
    uint64_t dims[3] = {(uint64_t)p->Do, (uint64_t)p->HW, (uint64_t)a.nsplit};
    uint64_t str[2] = {(uint64_t)p->Do * 4, (uint64_t)p->HW * p->Do * 4};
    uint32_t box[3] = {32, TQ, 1};
    int rc = make_tmap(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.o_part, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  launch_k(memory_read_tc_kernel<NP>, grid, kReadThreads, Cfg::kSmem, s, tmQ, tmK, tmV, tmO, a);
  OTVM_LAUNCH_CHECK();
  return read_combine(p, a.nsplit, s);
}

int memory_read_tc(const otvm_read_params* p, cudaStream_t s) {
  return dtype_planes(p->dtype) == 1 ? memory_read_tc_np<1>(p, s) : memory_read_tc_np<2>(p, s);
}

}  // namespace otvm

extern "C" __attribute__((visibility("default"))) void otvm_debug_set_read_timestamps(long long* buf) {
  otvm::g_read_dbg = buf;
}
