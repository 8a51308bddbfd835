// tcgen05 implicit-GEMM convolution for sm_100a (bf16 planes in, fp32 accumulate in TMEM).
//
// Split-bf16 operands: an activation / weight is the sum of P bf16 planes (P = 1, 2, 3: 8 / 16 / 24 significant bits,
// common.cuh).  A pipeline stage holds the P planes of the A tile and the P planes of the B tile, and the MMA warp
// issues the plane products (a0 b0 | + a0 b1 + a1 b0 | + a1 b1 + a0 b2 + a2 b0) into ONE fp32 accumulator: fp32-grade
// convolutions at tensor-core rate, 3x / 6x the MMA work of plain bf16 but the same TMEM / epilogue / tile traffic
// per output.  The cross-plane products go to a SECOND accumulator (columns BN..2BN-1) that the epilogue adds in fp32:
// the tensor core truncates (rounds toward zero) every time it adds an instruction's products to the accumulator, a
// bias of ~half an ulp per MMA that grows linearly with the number of instructions (measured: three planes in one
// accumulator were LESS exact than two, 8.5e-5 vs 5.9e-5 on the 9216-deep key projection); the cross terms are 2^-8 of
// the result, so their truncation is harmless in an accumulator of their own and the main one sees one product per K
// step again.  The epilogue splits the fp32 result back into planes.  The reference is fp32 end to end
// (models/trimap/STM.py, models/alpha/FBA/models.py) and the random-weight networks amplify storage rounding: plain
// bf16 misses its outputs by 0.3-0.7 in the max norm, two planes reach ~5e-4.
//
// GEMM view: M = output pixels, N = Cout, K = (ky, kx, ci).  One CTA computes a 128-pixel x BN-channel tile; the
// 128 pixels are a TH x TW rectangle of the output map, so for every filter tap the A operand is ONE 4-D TMA
// box {KC channels, TW, TH, 1} of the NHWC input at offset (kx*dil - pad, ky*dil - pad): TMA's out-of-bounds
// zero fill implements both the convolution padding and the ragged tile edge ("TMA im2col" without ever
// materialising im2col).  The box lands in shared memory as 128 rows of KC*2 bytes with the 128/64/32-byte
// swizzle, which is exactly the K-major canonical UMMA layout, so the MMA warp issues tcgen05.mma straight
// from it.  Weights [Cout][KH*KW*Cin] are the K-major B operand through a 2-D tensor map.
//
// Warp roles (192 or 320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 (+ 6..9 for tiles of >= 64 channels: two groups split the tile's columns) = epilogue (TMEM -> registers ->
// bias / GroupNorm statistics / residual, which arrives as a TMA tile / activation -> swizzled shared-memory tile that
// reuses the drained pipeline stages -> TMA tensor store per 64-column sub-tile, which also clips ragged tiles).
// Variants: GroupNorm statistics or the whole GroupNorm in the epilogue (grid barrier), split-K through a workspace or
// through a thread-block cluster (CLUSTER), grouped filter banks (one launch for the two STM encoders).
// Ring of NSTAGE {A,B} stages with full/empty mbarriers; tcgen05.commit releases stages and publishes the
// accumulator.  Up to two CTAs per SM (<= 113 KB smem, <= 128 TMEM columns each) so one CTA's epilogue overlaps
// another's main loop.
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>
#include "tc_common.cuh"

namespace otvm {

using namespace tc;

struct ConvTcArgs {
  int N, H, W, Cin, Ho, Wo, Cout, KH, KW, pad, dil, stride;
  int TW, TH, tiles_x, tiles_y;
  int tw_shift;                // TW = 1 << tw_shift
  uint32_t tiles_x_magic, tiles_y_magic;   // ceil(2^32 / d) for the tile-coordinate divisions (0: d == 1)
  int KC, nchunk, nstage, ksub; // nstage ring groups of ksub K-chunks each
  int k_per_split;             // K iterations per blockIdx.z slice (split-K); == num_k without split
  int cluster_k;               // > 1 (CLUSTER kernels): the blockIdx.z slices of a tile form a thread-block cluster; CTA 0 adds
                               // the other slices' fp32 partial tiles through distributed shared memory and runs the epilogue
  int64_t split_stride;        // elements between the fp32 partial outputs of consecutive K slices
  uint32_t aux_off;            // barriers / tmem slot / stats / bias live after max(pipeline, staging) bytes
  uint32_t a_bytes, b_bytes, sbo, layout_type;
  // halo mode (3x3 stride-1): ONE input patch {KC, TW+2d, TH+2d} per K-chunk feeds all 9 taps; the tap (ky,kx) operand is
  // the same shared-memory copy read through a descriptor whose start is shifted by whole pixel rows
  int halo, na;                // na = patch ring slots
  uint32_t patch_bytes, b_off; // patch slot size (1024-aligned); byte offset of the B ring behind the patch ring
  uint32_t row_bytes;          // KC * 2
  // split-bf16: planes per operand, plane products per K step (1 / 3 / 6), bytes of ONE operand plane inside a stage
  // (a_bytes = planes * a_plane, b_bytes = planes * b_plane, a patch-ring slot = planes * patch_bytes)
  int planes, npair;
  int wide;                    // split operands: products (a0 b0) and (a0 b1) are ONE MMA of N = 2 BN (see the MMA issuers)
  uint32_t a_plane, b_plane;
  int64_t act_plane;           // elements between the planes of out / res / out_relu in global memory
  const float* bias;
  void* out; int64_t out_ps, out_cs;
  const bf16* res; int64_t res_ld;
  bf16* out_relu; int64_t out_relu_ld;
  int act, out_f32;
  double* gn_stats;
  const float* gn_gamma; const float* gn_beta; float gn_eps;   // GN == 2: normalise in this kernel (grid barrier)
  double gn_inv_cnt;           // 1 / elements per GroupNorm group
  int gn_cg;                   // channels per GroupNorm group (Cout / 32 unless the caller runs a channel slice)
  int w_group_rows;            // grouped convolution (otvm_conv_params.groups): image n reads filter rows n * Cout + ..; else 0
  long long* dbg;              // dev: per-CTA clock64 timestamps [grid][8] (NULL in production)
  // persistent patch-mode kernel (conv_tc_persist_kernel): tiles walked per CTA, smem carve-up
  int ntiles;                  // tiles_x * tiles_y * N
  uint32_t p_off, stage_off;   // byte offsets of the patch ring and of the two epilogue staging tiles (weights sit at 0)
};

constexpr int kConvThreads = 192;       // TMA + MMA warps + ONE group of four epilogue warps
constexpr int kConvThreadsMax = 320;    // ... + a second epilogue group (BN >= 64): the groups split the tile's columns

// per-chunk GroupNorm partial sums: CGC consecutive channels form one slot.  Every thread (= tile row) parks its
// partial (sum, sum of squares) per slot in shared memory, sred[which][slot][129]; after the chunk loop one thread
// per (slot, which) adds up the 128 rows.  (Warp-shuffle trees cost 10 shuffles per slot per thread, which made
// the epilogue of the 64-channel full-resolution convolutions, 32 groups of 2, as long as their main loop.)
constexpr int kSredPitch = 129;
template <int CH, int CGC>
__device__ __forceinline__ void gn_chunk(const float (&qv)[CH], int r, float* sred, int slot0) {
#pragma unroll
  for (int g = 0; g < CH / CGC; ++g) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < CGC; ++j) { const float x = qv[g * CGC + j]; s1 += x; s2 += x * x; }
    sred[(slot0 + g) * kSredPitch + r] = s1;
    sred[(32 + slot0 + g) * kSredPitch + r] = s2;
  }
}

enum { EPI_RES = 1, EPI_RELU2 = 2, EPI_DIRECT = 4 };
enum { GN_NONE = 0, GN_STATS = 1, GN_FUSED = 2 };   // GN_FUSED: statistics -> grid barrier -> normalise + affine in the epilogue

__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// sticky device-side error flags, read (and cleared) by otvm_device_error_flags():
//   bit 0: the grid barrier of a GroupNorm-fused convolution timed out (a CTA of the grid never became resident)
__device__ unsigned int g_device_error_flags = 0;

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// plane products of a K step, in issue order: (A plane, B plane) of product i = (kPairA >> 4i & 15, kPairB >> 4i & 15);
// the first 1 / 3 / 6 entries are the products of 1 / 2 / 3 planes (terms below 2^-16 / 2^-24 of the result dropped)
constexpr uint32_t kPairA = 0x201100u, kPairB = 0x021010u;

// 8 fp32 values -> `planes` bf16 planes, 16 bytes each at base + plane * stride (x is consumed)
__device__ __forceinline__ uint4 pack8_and_subtract(float (&x)[8], bool subtract) {
  uint32_t pk[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
    pk[e] = *reinterpret_cast<const uint32_t*>(&h);
    if (subtract) { x[2 * e] -= __low2float(h); x[2 * e + 1] -= __high2float(h); }
  }
  return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}
__device__ __forceinline__ void split_store8(uint8_t* base, uint32_t stride, int planes, float (&x)[8]) {
  // (straight-line per plane count: the remainder is not computed behind the last plane)
  if (planes == 1) {
    *reinterpret_cast<uint4*>(base) = pack8_and_subtract(x, false);
  } else if (planes == 2) {
    *reinterpret_cast<uint4*>(base) = pack8_and_subtract(x, true);
    *reinterpret_cast<uint4*>(base + stride) = pack8_and_subtract(x, false);
  } else {
    *reinterpret_cast<uint4*>(base) = pack8_and_subtract(x, true);
    *reinterpret_cast<uint4*>(base + stride) = pack8_and_subtract(x, true);
    *reinterpret_cast<uint4*>(base + 2 * (size_t)stride) = pack8_and_subtract(x, false);
  }
}

__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Sum of the 128 row partials of every (statistic, slot) column of `sred` -> one fp64 atomic per column.  `nthr` epilogue
// threads (e = 0 .. nthr-1, whole warps) share the 2 * nslot columns: `per` = 1, 2 or 4 threads per column, each adds
// 128 / per rows in an order that keeps the 32 lanes of a warp on 32 different banks (column pitch 129), then a shuffle
// tree.  (One thread per column walking 128 rows was ~600 cycles of the GroupNorm-fused epilogue's critical path.)
__device__ __forceinline__ void gn_reduce_columns(const float* sred, int e, int nthr, int nslot, int cgc, int cg, int n0, int Cout,
                                                  double* gn_stats) {
  const int ncol = 2 * nslot;
  const int per = nthr >= 4 * ncol ? 4 : nthr >= 2 * ncol ? 2 : 1;
  const int col = e / per, sub = e - col * per;
  const int rows = 128 / per, rot = (32 / per) * sub;
  const int which = col / nslot, slot = col - which * nslot;
  const bool live = col < ncol && n0 + slot * cgc < Cout;
  float acc = 0.f;
  if (live) {
    const float* colp = sred + (which * 32 + slot) * kSredPitch + rows * sub;
#pragma unroll 8
    for (int j = 0; j < rows; ++j) acc += colp[(j + rot) & (rows - 1)];
  }
  if (per >= 2) acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (per >= 4) acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (live && sub == 0) atomicAdd(&gn_stats[((n0 + slot * cgc) / cg) * 2 + which], (double)acc);
}

template <int BN, int GN, int EPI, bool HALO, bool CLUSTER = false>
__global__ void __launch_bounds__(kConvThreadsMax, ((EPI & 4) != 0 ? 1 : 2)) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmO,
                                                               const __grid_constant__ CUtensorMap tmR,
                                                               const ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [stages x (A|B)] at 1024-byte alignment, then barriers
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage_bytes = a.a_bytes + a.b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + a.aux_off);
  uint64_t* empty_bar = full_bar + a.nstage;
  uint64_t* accum_bar = empty_bar + a.nstage;
  uint64_t* a_empty = accum_bar + 1;                                  // [4] patch ring (halo mode)
  uint64_t* a_full = a_empty + 4;                                     // [4]
  uint64_t* res_bar = a_full + 3;                                     // residual tile landed (GroupNorm-fused + residual); the
                                                                      // patch ring never has more than two slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 4);
  float* sstat = reinterpret_cast<float*>(tmem_slot + 2);            // [<=128] GroupNorm partials (sum, sumsq per slot)
  float* sbias = sstat + 256;                                        // [BN] bias of this tile's channels

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = a.dbg ? a.dbg + (size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64 : nullptr;
  if (dbg && threadIdx.x == 0) { dbg[0] = clock64(); dbg[10] = gtime_ns(); }
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;            // columns of one accumulator
  const uint32_t tmem_cols = a.planes > 1 ? 2 * ACC_COLS : ACC_COLS;   // split operands: main + cross-term accumulator

  // tile coordinates
  int bx = blockIdx.x;
  // (runtime integer division costs ~100 cycles a piece in every CTA's prologue: magic-number reciprocals instead)
  const int bq = a.tiles_x_magic ? (int)__umulhi((uint32_t)bx, a.tiles_x_magic) : bx;
  const int tx_i = bx - bq * a.tiles_x;
  const int n_img = a.tiles_y_magic ? (int)__umulhi((uint32_t)bq, a.tiles_y_magic) : bq;
  const int ty_i = bq - n_img * a.tiles_y;
  const int x0 = tx_i * a.TW, y0 = ty_i * a.TH;
  const int n0 = blockIdx.y * BN;
  const int wrow = n_img * a.w_group_rows + n0;               // first filter row (grouped: the bank of this image)
  const int num_k_all = a.KH * a.KW * a.nchunk;
  const int it0 = blockIdx.z * a.k_per_split;
  const int num_k = min(a.k_per_split, num_k_all - it0);      // this CTA's K range is [it0, it0 + num_k)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < a.nstage; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&a_empty[i], 1); mbar_init(&a_full[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_n(tmem_slot, tmem_cols);
  // (the bias is fetched by the epilogue warps behind this barrier: a global load in front of it put ~700 cycles of
  // memory latency into every CTA's prologue)
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barriers, TMEM, tensor-map prefetch, bias = constant weights) overlapped the previous
  // kernel's tail; this CTA now owns all its resources, so dependents may be scheduled, and from here on it
  // touches activations produced by earlier kernels
  // The weight-producer warp does NOT wait here: filters are never written inside a frame, so it issues the weight boxes
  // of the whole first ring pass (every slot is free at kernel start) while the previous kernel is still running, and
  // only then joins the wait.  Weights stream from HBM (150 MB per frame do not stay in L2), so they were the slowest
  // part of a CTA's first stage.
  constexpr int kWeightWarp = HALO ? 0 : 2;
  pdl_trigger();
  if (warp != kWeightWarp) pdl_wait();
  if (dbg && threadIdx.x == 32) { dbg[1] = clock64(); dbg[11] = gtime_ns(); }

  // NOTE on the single-thread loops below: one thread issuing a dependent scalar chain is the pipeline's critical
  // path (measured ~830 cycles per K iteration with runtime div/mod for the stage / tap / chunk indices, and ~430
  // when all 32 lanes polled the mbarrier, vs 256 cycles of MMA work), so indices are carried incrementally and
  // exactly one thread per role runs its loop.
  if constexpr (HALO) {
    // iteration order: K-chunk major, the 9 taps inner; `it0`/`num_k` are multiples of 9 (whole chunks per K slice)
    const int d = a.dil, pw = a.TW + 2 * a.dil;                 // patch width in pixels
    if (warp == 2) {
      // ===== patch producer (joins the epilogue afterwards): one input patch per K-chunk into the patch ring; it runs
      // up to `na` chunks ahead of the MMAs, independent of the depth of the weight ring =====
      int as = 0; uint32_t aph = 0;
      const int cx = x0 - a.pad, cy = y0 - a.pad;
      const uint32_t a_tx = (uint32_t)((a.TW + 2 * d) * (a.TH + 2 * d)) * a.row_bytes * (uint32_t)a.planes;
      const uint32_t slot_bytes = (uint32_t)a.planes * a.patch_bytes;
      const int c0 = it0 / 9, c1 = c0 + num_k / 9;
      for (int chunk = c0; chunk < c1; ++chunk) {
        mbar_wait(&a_empty[as], aph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[as], a_tx);
          for (int pl = 0; pl < a.planes; ++pl)
            tma_load_5d(smem + (size_t)as * slot_bytes + (size_t)pl * a.patch_bytes, &tmA, &a_full[as], chunk * a.KC, cx, cy,
                        n_img, pl);
        }
        __syncwarp();
        if (++as == a.na) { as = 0; aph ^= 1; }
      }
    } else if (warp == 0) {
      // ===== weight producer: one [BN x KC] box per (chunk, tap) =====
      int s = 0; uint32_t ph = 0;
      int tap = 0, chunk = it0 / 9;
      const uint32_t b_tx = (uint32_t)(BN * a.KC * 2 * a.planes);
      bool waited = false;
      for (int it = 0; it < num_k; ++it) {
        if (it >= a.nstage) {                                  // first ring pass: slots are free, no PDL wait yet
          if (!waited) { pdl_wait(); waited = true; }
          mbar_wait(&empty_bar[s], ph ^ 1);
        }
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], b_tx);
          for (int pl = 0; pl < a.planes; ++pl)
            tma_load_3d(smem + a.b_off + (size_t)s * a.b_bytes + (size_t)pl * a.b_plane, &tmB, &full_bar[s],
                        tap * a.Cin + chunk * a.KC, wrow, pl);
        }
        __syncwarp();
        if (++s == a.nstage) { s = 0; ph ^= 1; }
        if (++tap == 9) { tap = 0; ++chunk; }
      }
      if (!waited) pdl_wait();
    } else if (warp == 1) {
      // ===== MMA issuer: A descriptor = patch slot + (ky*d*pw + kx*d) pixel rows; 8-row groups are tile rows (TW = 8),
      // SBO = pw pixel rows.  Row-shifted SWIZZLE_* operands are valid because the hardware swizzles on absolute
      // shared-memory address bits (probed on B200: csrc/umma_probe.cu) =====
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN);
      constexpr uint32_t idesc2 = make_idesc_bf16(128, BN < 16 ? 32 : 2 * BN);
      const int ksteps = a.KC / 16;
      const uint64_t adesc0 = make_smem_desc(base, (uint32_t)pw * a.row_bytes, a.layout_type);
      const uint64_t bdesc0 = make_smem_desc(base + a.b_off, a.sbo, a.layout_type);
      const uint32_t bstage16 = a.b_bytes >> 4, patch16 = ((uint32_t)a.planes * a.patch_bytes) >> 4;
      const uint32_t aplane16 = a.patch_bytes >> 4, bplane16 = a.b_plane >> 4;
      const uint32_t kx16 = ((uint32_t)d * a.row_bytes) >> 4, ky16 = ((uint32_t)(d * pw) * a.row_bytes) >> 4;
      int s = 0; uint32_t ph = 0, soff = 0, aoff = 0, toff = 0, aph = 0; int as = 0, kx = 0, ky = 0;
      for (int it = 0; it < num_k; ++it) {
        if ((kx | ky) == 0) mbar_wait(&a_full[as], aph);
        mbar_wait(&full_bar[s], ph);
        tcgen05_after_sync();
        const bool last_tap = kx == 2 && ky == 2;
        if (elect_one()) {
          const uint64_t ad0 = adesc0 + aoff + toff, bd0 = bdesc0 + soff;
          // product 0 -> main accumulator, cross-plane products -> the second one (columns ACC_COLS..).  `wide`: the B
          // planes are contiguous [hi | lo] rows and the two accumulators adjacent columns, so a0 . [b0 | b1] is one
          // instruction of N = 2 BN writing both (product 1 is then skipped): a third fewer issues per K step, which is
          // what paces the N <= 64 layers (one thread issues ~55 cycles per instruction, 32 of tensor work)
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_base, ad0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), a.wide ? idesc2 : idesc, (it | k) != 0);
          for (int pr = a.wide ? 2 : 1; pr < a.npair; ++pr) {
            const uint64_t ad = ad0 + ((kPairA >> (4 * pr)) & 15u) * aplane16, bd = bd0 + ((kPairB >> (4 * pr)) & 15u) * bplane16;
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(tmem_base + ACC_COLS, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc,
                        a.wide ? 1u : (uint32_t)((it | (pr - 1) | k) != 0));
          }
          umma_commit(&empty_bar[s]);
          if (last_tap) umma_commit(&a_empty[as]);             // all 9 taps of this chunk have read the patch
          if (it == num_k - 1) umma_commit(accum_bar);
        }
        __syncwarp();
        soff += bstage16;
        if (++s == a.nstage) { s = 0; ph ^= 1; soff = 0; }
        if (last_tap) { kx = 0; ky = 0; toff = 0; aoff += patch16; if (++as == a.na) { as = 0; aoff = 0; aph ^= 1; } }
        else if (++kx == 3) { kx = 0; ++ky; toff += ky16 - 2 * kx16; }
        else toff += kx16;
      }
    }
  } else {
  // ===== one-box-per-tap mode.  Each of the three loops below is ONE serial instruction stream (an elected lane) that
  // paces the whole pipeline, so they are written to the instruction: raw 32-bit shared-memory addresses carried
  // incrementally, 32-bit descriptor low words, K steps unrolled at compile time, no debug stamps inside.  Measured
  // before (scripts/conv_ts3.py): the MMA warp passed one stage per 495 cycles and the producers issued one per 425,
  // against 256 cycles of tensor work per 128x128x64 stage. =====
  if (warp == 0) {
    // ----- activation producer: one 4-D box {KC, TW*s, TH*s, 1} per (tap, chunk); posts the stage's byte count -----
    int chunk = 0, kx = 0, ky = 0;
    if (it0 != 0) { const int tap = it0 / a.nchunk; chunk = it0 - tap * a.nchunk; ky = tap / a.KW; kx = tap - ky * a.KW; }
    const int cx = x0 * a.stride - a.pad, cy = y0 * a.stride - a.pad;
    int c0 = chunk * a.KC, xx = cx + kx * a.dil, yy = cy + ky * a.dil;
    const uint32_t tx_bytes = a.a_bytes + (uint32_t)(BN * a.KC * 2 * a.planes);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t gbytes = (uint32_t)a.ksub * stage_bytes;            // one ring slot = ksub consecutive {A,B} stages
    uint32_t fa = full0, ea = empty0, sa = base, ph = 1;
    int s = 0;
    auto advance = [&]() {
      c0 += a.KC;
      if (++chunk == a.nchunk) { chunk = 0; c0 = 0; xx += a.dil; if (++kx == a.KW) { kx = 0; xx = cx; yy += a.dil; } }
    };
    for (int it = 0; it < num_k;) {
      const bool two = a.ksub == 2 && it + 1 < num_k;
      mbar_wait_a(ea, ph);
      const int c0a = c0, xa = xx, ya = yy;
      advance();
      const int c0b = c0, xb = xx, yb = yy;
      if (two) advance();
      if (elect_one()) {
        mbar_arrive_expect_tx_a(fa, two ? 2 * tx_bytes : tx_bytes);
        for (int pl = 0; pl < a.planes; ++pl) tma_load_5d_a(sa + (uint32_t)pl * a.a_plane, &tmA, fa, c0a, xa, ya, n_img, pl);
        if (two) tma_load_5d_a(sa + stage_bytes, &tmA, fa, c0b, xb, yb, n_img, 0);      // (ksub = 2 only with one plane)
        if (dbg && it == 0) dbg[2] = clock64();
      }
      __syncwarp();
      it += two ? 2 : 1;
      fa += 8; ea += 8; sa += gbytes;
      if (++s == a.nstage) { s = 0; ph ^= 1; fa = full0; ea = empty0; sa = base; }
    }
  } else if (warp == 2) {
    // ----- weight producer (joins the epilogue afterwards): box [BN x KC] at K offset (it0 + it) * KC (the packed K
    // axis is tap-major, chunk-minor, so consecutive iterations are consecutive K columns).  A weight box landing
    // before the activation thread's expect_tx only makes the transaction count transiently negative. -----
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t gbytes = (uint32_t)a.ksub * stage_bytes;
    uint32_t fa = full0, ea = empty0, sb = base + a.a_bytes, ph = 1;
    int s = 0, k0 = it0 * a.KC;
    bool waited = false;
    for (int it = 0, gi = 0; it < num_k; ++gi) {
      const bool two = a.ksub == 2 && it + 1 < num_k;
      if (gi >= a.nstage) {                                    // first ring pass: slots are free, no PDL wait yet
        if (!waited) { pdl_wait(); waited = true; }
        mbar_wait_a(ea, ph);
      }
      if (elect_one()) {
        for (int pl = 0; pl < a.planes; ++pl) tma_load_3d_a(sb + (uint32_t)pl * a.b_plane, &tmB, fa, k0, wrow, pl);
        if (two) tma_load_3d_a(sb + stage_bytes, &tmB, fa, k0 + a.KC, wrow, 0);
      }
      __syncwarp();
      it += two ? 2 : 1;
      k0 += two ? 2 * a.KC : a.KC;
      fa += 8; ea += 8; sb += gbytes;
      if (++s == a.nstage) { s = 0; ph ^= 1; fa = full0; ea = empty0; sb = base + a.a_bytes; }
    }
    if (!waited) pdl_wait();
  } else if (warp == 1) {
    // ----- MMA issuer -----
    constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN);
    const uint32_t idesc0 = a.wide ? make_idesc_bf16(128, BN < 16 ? 32 : 2 * BN) : idesc;   // see the halo issuer
    const uint64_t adesc0 = make_smem_desc(base, a.sbo, a.layout_type);
    const uint64_t bdesc0 = make_smem_desc(base + a.a_bytes, a.sbo, a.layout_type);
    const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0, hi = (uint32_t)(adesc0 >> 32);   // same SBO / layout
    const uint32_t stage16 = stage_bytes >> 4, aplane16 = a.a_plane >> 4, bplane16 = a.b_plane >> 4;
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), accum_a = smem_u32(accum_bar);
    auto mma_loop = [&](auto ks_tag) {
      constexpr int KS = decltype(ks_tag)::value;
      const uint32_t g16 = (uint32_t)a.ksub * stage16;
      uint32_t fa = full0, ea = empty0, a_lo = a_lo0, b_lo = b_lo0, ph = 0;
      int s = 0;
      for (int it = 0; it < num_k;) {
        const bool two = a.ksub == 2 && it + 1 < num_k;
        const bool last = it + (two ? 2 : 1) >= num_k;
        mbar_wait_a(fa, ph);
        tcgen05_after_sync();
        if (elect_one()) {
          if (dbg && it == 0) dbg[3] = clock64();
          umma_bf16_lh(tmem_base, a_lo, hi, b_lo, hi, idesc0, it != 0);
#pragma unroll
          for (int k = 1; k < KS; ++k) umma_bf16_lh(tmem_base, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc0, 1u);
#pragma unroll
          for (int pr = 1; pr < 6; ++pr) {                     // cross-plane products of split operands
            if (pr < a.npair && !(pr == 1 && a.wide)) {
              const uint32_t ap = a_lo + ((kPairA >> (4 * pr)) & 15u) * aplane16, bp = b_lo + ((kPairB >> (4 * pr)) & 15u) * bplane16;
#pragma unroll
              for (int k = 0; k < KS; ++k)
                umma_bf16_lh(tmem_base + ACC_COLS, ap + 2 * k, hi, bp + 2 * k, hi, idesc,
                             a.wide ? 1u : (uint32_t)((it | (pr - 1) | k) != 0));
            }
          }
          if (two) {
#pragma unroll
            for (int k = 0; k < KS; ++k)
              umma_bf16_lh(tmem_base, a_lo + stage16 + 2 * k, hi, b_lo + stage16 + 2 * k, hi, idesc, 1u);
          }
          umma_commit_a(ea);                                   // frees the slot when these MMAs have read it
          if (last) {
            umma_commit_a(accum_a);                            // accumulator complete
            if (dbg) dbg[4] = clock64();
          }
        }
        __syncwarp();
        it += two ? 2 : 1;
        fa += 8; ea += 8; a_lo += g16; b_lo += g16;
        if (++s == a.nstage) { s = 0; ph ^= 1; fa = full0; ea = empty0; a_lo = a_lo0; b_lo = b_lo0; }
      }
    };
    if (a.KC == 64) mma_loop(std::integral_constant<int, 4>{});
    else if (a.KC == 32) mma_loop(std::integral_constant<int, 2>{});
    else mma_loop(std::integral_constant<int, 1>{});
  }
}
  // Cluster split-K (CLUSTER kernels, a.cluster_k CTAs along blockIdx.z): every thread of every CTA of the cluster passes two
  // cluster barriers: #1 once the partial accumulators of ranks 1.. are parked in their shared memory, #2 once rank 0 has
  // read them.  A template parameter, not a run-time switch: the extra branches cost every epilogue 1.3 % when they were.
  const uint32_t crank = CLUSTER ? blockIdx.z : 0u;
  if constexpr (CLUSTER) { if (warp < 2) { __syncwarp(); cluster_sync_all(); } }
  if (warp >= 2) {
    // ===== epilogue: warps 2..5 (and 6..9 when the block has a second group) own TMEM lanes 32*(warp%4) .. +31 =====
    // Compile-time variants (EPI) keep the per-element instruction count low: the three store paths and the
    // residual paths would otherwise all be issued as predicated-off instructions (measured: 1800 SASS
    // instructions per 32-column chunk, which made the epilogue issue-bound at ~2000 cycles per chunk).
    //
    // Short-K layers are bound by THIS code, not by the MMAs (scripts/conv_ts4.py, 64 -> 256 1x1 at 128^2, two planes:
    // accumulator ready after 2.1 us, epilogue 4.3 us plain / 7.6 us with a residual / 15 us GroupNorm-fused).  Hence:
    //   * two groups of four warps split the tile's columns (the chain tcgen05.ld -> bias -> residual -> activation ->
    //     plane split -> st.shared is latency-bound at one warp per scheduler),
    //   * the residual of a chunk is requested before the accumulator wait / while the previous chunk is processed,
    //   * each 64-column sub-tile is handed to the TMA store as soon as it is staged, so the store overlaps the rest.
    constexpr bool kRes = (EPI & EPI_RES) != 0, kRelu2 = (EPI & EPI_RELU2) != 0, kDirect = (EPI & EPI_DIRECT) != 0;
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;                         // epilogue group 0 / 1
    const int r = q * 32 + lane;                             // tile row = pixel
    const int ty = r >> a.tw_shift, tx = r - (ty << a.tw_shift);
    const int oy = y0 + ty, ox = x0 + tx;
    const bool valid = oy < a.Ho && ox < a.Wo;
    const int64_t pix = ((int64_t)n_img * a.Ho + oy) * a.Wo + ox;
    // columns per TMEM load: 32; 16 with a residual (two prefetched residual planes + main / cross accumulator chunks must
    // fit the 96 registers that keep two 320-thread CTAs on an SM)
    // Residual: the residual TILE is fetched by TMA into the staging tiles as soon as the pipeline stages are dead (all
    // MMAs complete), read back from shared memory and the result written in its place: no global loads and no residual
    // registers in the epilogue.  GroupNorm-fused layers hide the fetch behind their statistics pass + grid barrier; the
    // others wait ~1 us once instead of two exposed global-load latencies per 16-column chunk (64 -> 256 1x1 at 128^2:
    // 3.2 us of a 5 us epilogue).  tmR is the residual's tensor map; the variants with a second (ReLU) output use tmR for
    // that output and keep the register path (kResTma false).
    constexpr bool kResTma = kRes && !kRelu2;
    constexpr int CH = (BN >= 32 && (!kRes || kResTma)) ? 32 : 16;
    constexpr int NCHUNK = BN / CH;
    // column range of this group: BN = 128 -> one 64-column sub-tile per group, BN = 64 -> one 32-column chunk per group
    const int ngrp = (NCHUNK >= 2 && blockDim.x > kConvThreads) ? 2 : 1;
    const int nthr = 128 * ngrp;                             // epilogue threads (GroupNorm barriers / reductions)
    const int c_lo = grp * (BN / ngrp), c_hi = c_lo + BN / ngrp;
    const bool idle = grp >= ngrp;                           // (a second group on a tile with a single chunk)
    const int cg = GN != GN_NONE ? a.gn_cg : 0;              // channels per GroupNorm group
    const int cgc = cg < CH ? cg : CH;
    constexpr uint32_t TILE = 128u * BN * 2u;                  // one bf16 staging tile (one plane of the output tile)
    const uint32_t nplane = (uint32_t)a.planes;
    float* sred = reinterpret_cast<float*>(smem + nplane * TILE);     // behind the staging tiles, inside the drained stages
    // act(v) = max(v, slope * v): none -> 1, ReLU -> 0, LeakyReLU -> 0.01
    const float slope = a.act == OTVM_ACT_NONE ? 1.f : a.act == OTVM_ACT_RELU ? 0.f : 0.01f;
    const int e = threadIdx.x - 64;                          // 0 .. nthr-1
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);

    // residual planes 0 / 1 of one chunk (a third plane is fetched where it is added)
    uint4 rr[2][CH / 8];
    auto load_res = [&](int cbase) {
      if constexpr (kRes && !kResTma) {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          if ((uint32_t)pl < nplane) {
            const uint4* rp = reinterpret_cast<const uint4*>(a.res + (int64_t)pl * a.act_plane + pix * a.res_ld + cbase);
#pragma unroll
            for (int j = 0; j < CH / 8; ++j) rr[pl][j] = valid ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
          }
        }
      }
    };
    // accumulator chunk (main + cross-plane columns) -> raw[]
    auto load_acc = [&](int c, uint32_t (&raw)[CH]) {
      uint32_t raw2[CH];
      if constexpr (CH == 32) tmem_ld32(trow + (uint32_t)c, raw); else tmem_ld16(trow + (uint32_t)c, raw);
      if (nplane > 1) {
        if constexpr (CH == 32) tmem_ld32(trow + ACC_COLS + (uint32_t)c, raw2); else tmem_ld16(trow + ACC_COLS + (uint32_t)c, raw2);
      }
      tmem_wait_ld();
      if (nplane > 1) {
#pragma unroll
        for (int j = 0; j < CH; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
      }
      if constexpr (CLUSTER) {
        if (crank == 0) {                                      // + the K slices of the other CTAs of the cluster
          const uint32_t poff = (((uint32_t)c >> 2) * 128u + (uint32_t)r) * 16u;
#pragma unroll 1
          for (uint32_t pr = 1; pr < (uint32_t)a.cluster_k; ++pr) {
            const uint32_t rem = dsmem_addr(base, pr) + poff;
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 t = ld_dsmem_f4(rem + (uint32_t)j * (128u * 16u));
              raw[4 * j] = __float_as_uint(__uint_as_float(raw[4 * j]) + t.x);
              raw[4 * j + 1] = __float_as_uint(__uint_as_float(raw[4 * j + 1]) + t.y);
              raw[4 * j + 2] = __float_as_uint(__uint_as_float(raw[4 * j + 2]) + t.z);
              raw[4 * j + 3] = __float_as_uint(__uint_as_float(raw[4 * j + 3]) + t.w);
            }
          }
        }
      }
    };
    auto stats_chunk = [&](const float (&qv)[CH], int c) {
      const int slot0 = c / cgc;
      switch (cgc) {
        case 1: gn_chunk<CH, 1>(qv, r, sred, slot0); break;
        case 2: gn_chunk<CH, 2>(qv, r, sred, slot0); break;
        case 4: gn_chunk<CH, 4>(qv, r, sred, slot0); break;
        case 8: gn_chunk<CH, 8>(qv, r, sred, slot0); break;
        case 16: gn_chunk<CH, 16>(qv, r, sred, slot0); break;
        default: gn_chunk<CH, CH>(qv, r, sred, slot0); break;
      }
    };

    if constexpr (CLUSTER) {
      // partial tile of a rank >= 1 CTA: fp32 [BN/4 column quads][128 rows][4] at the start of its (drained) ring, so that
      // a warp's 16-byte accesses are contiguous on both sides
      if (crank != 0 && !idle) {
        mbar_wait(accum_bar, 0);
        tcgen05_after_sync();
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += CH) {
          uint32_t raw[CH];
          load_acc(c, raw);
          uint8_t* pp = smem + (((uint32_t)c >> 2) * 128u + (uint32_t)r) * 16u;
#pragma unroll
          for (int j = 0; j < CH / 4; ++j)
            *reinterpret_cast<uint4*>(pp + (size_t)j * (128 * 16)) = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
        }
      }
      __syncwarp();
      cluster_sync_all();                                    // #1
    }
    if (!idle && crank == 0) {                               // (whole warps: a second group on a single-chunk tile does nothing)
    // bias (a constant: no PDL dependency, and warp 2 -- weight / patch producer first -- has waited anyway)
    for (int i = e; i < BN; i += nthr) sbias[i] = (a.bias && n0 + i < a.Cout) ? a.bias[wrow + i] : 0.f;
    if (n0 + c_lo < a.Cout) load_res(n0 + c_lo);             // in flight while the last MMAs finish
    bar_sync_n(1, nthr);                                     // bias visible to every epilogue thread
    mbar_wait(accum_bar, 0);
    tcgen05_after_sync();
    if (dbg && threadIdx.x == 64) dbg[5] = clock64();
    if constexpr (kResTma) {
      // every MMA has completed: the pipeline stages are dead, the staging tiles may be filled
      if (threadIdx.x == 64) {
        constexpr uint32_t ROWB = (BN < 64 ? BN : 64) * 2;
        constexpr int NSUB = BN > 64 ? BN / 64 : 1;
        int nsub = 0;
        for (int sub = 0; sub < NSUB; ++sub) nsub += (n0 + sub * 64 < a.Cout) ? 1 : 0;
        mbar_arrive_expect_tx(res_bar, (uint32_t)nsub * nplane * 128u * ROWB);
        for (uint32_t pl = 0; pl < nplane; ++pl)
          for (int sub = 0; sub < nsub; ++sub)
            tma_load_5d(smem + pl * TILE + (size_t)sub * 128 * ROWB, &tmR, res_bar, n0 + sub * 64, x0, y0, n_img, (int)pl);
      }
    }
    if constexpr (GN == GN_FUSED) {
      // ---- fused GroupNorm: pass 1 accumulates the statistics of the fp32 convolution output (the accumulator
      // stays in TMEM), a grid-wide barrier makes every CTA's contribution visible, pass 2 (the loop below)
      // normalises.  The host only selects this variant when the whole grid is co-resident (one wave).
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += CH) {
        if (n0 + c >= a.Cout) break;
        uint32_t raw[CH];
        load_acc(c, raw);
        float qv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) qv[j] = valid ? __uint_as_float(raw[j]) + sbias[c + j] : 0.f;
        stats_chunk(qv, c);
      }
      if (dbg && e == 0) dbg[13] = clock64();                 // statistics pass done (this thread)
      bar_sync_n(1, nthr);
      gn_reduce_columns(sred, e, nthr, BN / cgc, cgc, cg, n0, a.Cout, a.gn_stats);
      bar_sync_n(1, nthr);                                    // this CTA's atomics are issued
      if (dbg && e == 0) dbg[14] = clock64();                 // column sums + atomics issued
      if (e == 0) {
        unsigned int* ctr = reinterpret_cast<unsigned int*>(a.gn_stats + 64);     // zeroed with the statistics arena
        const unsigned int total = gridDim.x * gridDim.y;
        // release at gpu scope: the CTA's statistics atomics (ordered before this by the barrier) are visible to whoever
        // acquires the incremented counter
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        if (dbg) dbg[15] = clock64();                         // release-increment issued
        // Bounded wait.  The host only selects this variant for grids it computed to be co-resident, but it cannot see
        // what else shares the device (other streams / processes, MPS, a debugger): if a CTA is still missing after
        // ~1 s the kernel gives up WITHOUT trapping (a trap would poison the whole CUDA context, graph replays included):
        // it raises a sticky error flag the host reads with otvm_device_error_flags() and finishes with whatever
        // statistics have arrived, so this launch's output is wrong but the process and the stream stay usable.
        uint32_t spins = 0;
        while (ld_acquire_gpu(ctr) < total) {
          __nanosleep(20);
          if (++spins > (1u << 24)) { atomicOr(&g_device_error_flags, 1u); break; }
        }
        if (dbg) { dbg[16] = clock64(); dbg[18] = spins; }    // barrier passed; polls that saw it closed
      }
      bar_sync_n(1, nthr);
      if (e < BN && n0 + e < a.Cout) {
        const int ch = n0 + e, g = ch / cg;
        // (fp64 only for the cancelling subtraction; fp64 division / square root run at 1/64 rate on this part)
        const double mean = __ldcg(&a.gn_stats[g * 2]) * a.gn_inv_cnt;
        const double var = fma(-mean, mean, __ldcg(&a.gn_stats[g * 2 + 1]) * a.gn_inv_cnt);
        const float rstd = 1.0f / sqrtf(fmaxf((float)var, 0.f) + a.gn_eps);
        const float sc = rstd * a.gn_gamma[ch];
        sstat[e] = sc;
        sstat[128 + e] = a.gn_beta[ch] - (float)mean * sc;
      }
      bar_sync_n(1, nthr);
      if (dbg && e == 0) dbg[17] = clock64();                 // scale / shift of this tile's channels ready
    }
    // sub-tile hand-over to the TMA store: the threads that staged a 64-column sub-tile meet on a named barrier and one
    // of them issues its stores; BN = 128 with two groups: each group owns one sub-tile (barriers 2 / 3, 128 threads)
    const bool own_sub = ngrp == 2 && BN == 128;
    const int store_bar = own_sub ? 2 + grp : 2, store_thr = own_sub ? 128 : nthr;
    const bool storer = own_sub ? (threadIdx.x == 64 + 128 * grp) : threadIdx.x == 64;
    auto store_sub = [&](int sub) {
      if constexpr (!kDirect) {
        fence_proxy_async_smem();                             // generic-proxy smem writes -> visible to the TMA engine
        bar_sync_n(store_bar, store_thr);
        if (storer && n0 + sub * 64 < a.Cout) {
          constexpr uint32_t ROWB = (BN < 64 ? BN : 64) * 2;
#pragma unroll 1
          for (uint32_t pl = 0; pl < nplane; ++pl) {
            tma_store_5d(&tmO, smem + pl * TILE + (size_t)sub * 128 * ROWB, n0 + sub * 64, x0, y0, n_img, (int)pl);
            if constexpr (kRelu2)
              tma_store_5d(&tmR, smem + (nplane + pl) * TILE + (size_t)sub * 128 * ROWB, n0 + sub * 64, x0, y0, n_img, (int)pl);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    };
    if constexpr (kResTma) mbar_wait(res_bar, 0);
    int pending_sub = -1;                                    // staged sub-tile not yet handed to the TMA store
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += CH) {
      const int cbase = n0 + c;
      if (cbase >= a.Cout) break;                            // uniform over the threads that share a sub-tile
      uint32_t raw[CH];
      load_acc(c, raw);
      if (dbg && threadIdx.x == 64 && c == 0) dbg[8] = clock64();
      float v[CH];
#pragma unroll
      for (int j = 0; j < CH; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sbias + c + j);
        v[j] = __uint_as_float(raw[j]) + b4.x; v[j + 1] = __uint_as_float(raw[j + 1]) + b4.y;
        v[j + 2] = __uint_as_float(raw[j + 2]) + b4.z; v[j + 3] = __uint_as_float(raw[j + 3]) + b4.w;
      }
      if constexpr (GN == GN_FUSED) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = fmaf(v[j], sstat[c + j], sstat[128 + c + j]);
      }
      if constexpr (GN == GN_STATS) {
        // statistics of the values GroupNorm will read back (one plane: rounded to bf16; split: the fp32 value to
        // 2^-17); rows outside the image count 0
        float qv[CH];
        if (nplane == 1) {
#pragma unroll
          for (int j = 0; j < CH; ++j) qv[j] = valid ? __bfloat162float(__float2bfloat16_rn(v[j])) : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j) qv[j] = valid ? v[j] : 0.f;
        }
        stats_chunk(qv, c);
      }
      if constexpr (kResTma) {
        constexpr uint32_t ROWB = (BN < 64 ? BN : 64) * 2, MASK = ROWB == 128 ? 7u : ROWB == 64 ? 3u : 1u;
#pragma unroll
        for (int j = 0; j < CH / 8; ++j) {
          const int cc = c + 8 * j;
          uint32_t off = (uint32_t)(cc >> 6) * (128u * ROWB) + (uint32_t)r * ROWB + (uint32_t)(cc & 63) * 2u;
          off ^= ((off >> 7) & MASK) << 4;
#pragma unroll 1
          for (uint32_t pl = 0; pl < nplane; ++pl) {
            const uint4 t = *reinterpret_cast<const uint4*>(smem + pl * TILE + off);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int k = 0; k < 4; ++k) { v[8 * j + 2 * k] += __low2float(h[k]); v[8 * j + 2 * k + 1] += __high2float(h[k]); }
          }
        }
      } else if constexpr (kRes) {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          if ((uint32_t)pl < nplane) {
#pragma unroll
            for (int j = 0; j < CH / 8; ++j) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr[pl][j]);
#pragma unroll
              for (int k = 0; k < 4; ++k) { v[8 * j + 2 * k] += __low2float(h[k]); v[8 * j + 2 * k + 1] += __high2float(h[k]); }
            }
          }
        }
        if (nplane > 2) {                                      // third plane of a split residual
          const uint4* rp = reinterpret_cast<const uint4*>(a.res + 2 * a.act_plane + pix * a.res_ld + cbase);
#pragma unroll
          for (int j = 0; j < CH / 8; ++j) {
            const uint4 t = valid ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int k = 0; k < 4; ++k) { v[8 * j + 2 * k] += __low2float(h[k]); v[8 * j + 2 * k + 1] += __high2float(h[k]); }
          }
        }
        if (c + CH < c_hi && cbase + CH < a.Cout) load_res(cbase + CH);     // next chunk's residual: in flight during the split
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], slope * v[j]);
      if constexpr (!kDirect) {
        // swizzled staging tile(s): [BN/64][128 rows][min(BN,64) ch]; 16-byte chunk j of row r lands at the address the
        // TMA swizzle expects, so 8 consecutive rows cover all 32 banks (conflict-free 16 B stores).  Rows outside
        // the image are staged too and clipped by the tensor store.
        constexpr uint32_t ROWB = (BN < 64 ? BN : 64) * 2, MASK = ROWB == 128 ? 7u : ROWB == 64 ? 3u : 1u;
#pragma unroll
        for (int j = 0; j < CH / 8; ++j) {
          const int cc = c + 8 * j;
          uint32_t off = (uint32_t)(cc >> 6) * (128u * ROWB) + (uint32_t)r * ROWB + (uint32_t)(cc & 63) * 2u;
          off ^= ((off >> 7) & MASK) << 4;
          float x8[8];
          if constexpr (kRelu2) {                              // second output = ReLU(out): tiles nplane .. 2 nplane - 1
#pragma unroll
            for (int k = 0; k < 8; ++k) x8[k] = fmaxf(v[8 * j + k], 0.f);
            split_store8(smem + nplane * TILE + off, TILE, a.planes, x8);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) x8[k] = v[8 * j + k];
          split_store8(smem + off, TILE, a.planes, x8);
        }
        // a 64-column sub-tile is complete after its last chunk (BN = 128); with one group the remaining chunks overlap
        // the store of the first sub-tile
        pending_sub = c >> 6;
        if constexpr (BN == 128) { if ((c & 63) == 64 - CH) { store_sub(pending_sub); pending_sub = -1; } }
      } else if (valid) {
        // direct stores: fp32 heads ([P][8] / [P][12]) and the channel-major value bank (lanes = consecutive pixels)
        if (a.out_f32) {
          float* op = static_cast<float*>(a.out) + (int64_t)blockIdx.z * a.split_stride + pix * a.out_ps + (int64_t)cbase * a.out_cs;
          if (a.out_cs == 1 && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
            // channel-contiguous rows (the fp32 heads, split-K partial tiles): 16-byte stores for whole quads inside Cout
            // (scalar stores at a 32 / 48-byte row pitch touched every sector four times)
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              if (cbase + j + 3 < a.Cout) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              else {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (cbase + j + k < a.Cout) op[j + k] = v[j + k];
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (cbase + j < a.Cout) op[(int64_t)j * a.out_cs] = v[j];
          }
        } else {
          bf16* op = static_cast<bf16*>(a.out) + pix * a.out_ps + (int64_t)cbase * a.out_cs;
#pragma unroll 1
          for (uint32_t pl = 0; pl < nplane; ++pl, op += a.act_plane) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              const bf16 h = __float2bfloat16_rn(v[j]);
              if (cbase + j < a.Cout) op[(int64_t)j * a.out_cs] = h;
              v[j] -= __bfloat162float(h);
            }
          }
        }
      }
    }
    if (dbg && threadIdx.x == 64) dbg[9] = clock64();
    if constexpr (!kDirect) {
      // BN = 128: a sub-tile whose second chunk lies beyond Cout is still pending; BN <= 64: ONE sub-tile, staged by every
      // epilogue thread (a group whose columns lie beyond Cout staged nothing but takes part in the hand-over barrier)
      if constexpr (BN == 128) { if (pending_sub >= 0) store_sub(pending_sub); } else store_sub(0);
      if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tiles read before the CTA exits
    }
    if constexpr (GN == GN_STATS) {
      bar_sync_n(1, nthr);                                    // the epilogue warps only
      gn_reduce_columns(sred, e, nthr, BN / cgc, cgc, cg, n0, a.Cout, a.gn_stats);
    }
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();
    }
    tcgen05_before_sync();
  }
  if constexpr (CLUSTER) { __syncwarp(); cluster_sync_all(); }   // #2: rank 0 has read every partial tile
  __syncthreads();
  if (dbg && threadIdx.x == 0) { dbg[7] = clock64(); dbg[12] = gtime_ns(); }
  if (warp == 1) {
    tcgen05_after_sync();
    tmem_dealloc_n(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent patch-mode kernel for the full-resolution 3x3 stride-1 layers (Cout = BN <= 64, Cin <= 128: the decoder
// tail and the refinement module, thousands of 128-pixel tiles per layer).  The one-tile-per-CTA kernel above pays
// its fixed costs (barrier init, TMEM allocation, descriptor fetch, first-load latency, drain) once per 16 KB of
// output and re-reads the whole filter bank (up to 108 KB) from L2 for every tile.  Here ONE CTA per SM
//   * loads the complete filter bank into shared memory once (before the PDL wait: weights are constants),
//   * walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... : one TMA patch {KC, 8+2d, 16+2d} per K-chunk through a
//     ring that runs ahead across tile boundaries,
//   * runs TWO independent tile pipelines (even / odd tiles of the CTA): each has its own MMA issuer warp, its own
//     pair of accumulators in tensor memory, its own 4 epilogue warps and its own bf16 staging tile.  Measured on B200
//     (scripts/persist_ts.py): ONE thread cannot issue N <= 64 MMAs faster than ~55 cycles apiece (uniform-datapath
//     descriptor arithmetic), which left the tensor pipe half idle; two issuers interleave on the pipe, and the
//     epilogue of one tile overlaps the MMAs of the next two,
//   * keeps the GroupNorm partial sums in registers across all its tiles and issues its 64 fp64 atomics once.
// Warp roles (352 threads): 0 = patch producer, 1..2 = MMA issuers (warp 1 allocates TMEM), 3..6 / 7..10 = epilogue
// of pipeline 0 / 1.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPersistThreads = 352;
constexpr int kSredPitch2 = 257;             // 256 rows (two epilogue groups) + 1

template <int BN, int GN, int KSTEPS>
__global__ void __launch_bounds__(kPersistThreads, 1) conv_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                             const __grid_constant__ CUtensorMap tmB,
                                                                             const __grid_constant__ CUtensorMap tmO,
                                                                             const ConvTcArgs a) {
  // statistics variant: Cout = 64 = 32 GroupNorm groups of 2 channels (host-checked); a CTA owns BN of them
  static_assert(GN == GN_NONE || (GN == GN_STATS && BN >= 32), "statistics variant");
  // Split operands (a.planes > 1): the filter bank holds [chunk][tap][plane] tiles, a patch-ring slot the planes of one
  // patch, and every tile has a main + a cross-plane accumulator (2 BN columns, see conv_tc_kernel).  Two planes of a
  // 64-channel layer do not fit one CTA (147 KB of weights alone), so such a layer runs as TWO 32-channel halves:
  // blockIdx.y picks the half, each half's CTAs walk all tiles.
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + a.aux_off);
  uint64_t* p_full = w_full + 1;                                      // [2][8] patch ring: one "full" barrier per pipeline and slot
  uint64_t* p_empty = p_full + 16;                                    // [8]
  uint64_t* acc_full = p_empty + 8;                                   // [4] accumulator 2g+b ready for epilogue group g
  uint64_t* acc_empty = acc_full + 4;                                 // [4] accumulator 2g+b drained (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);             // [BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;                                     // first output channel of this CTA
  const uint32_t nplane = (uint32_t)a.planes;
  const uint32_t ACCW = nplane > 1 ? 2u * BN : (uint32_t)BN;          // columns per tile: main (+ cross-plane) accumulator
  const uint32_t tmem_cols = 4u * ACCW < 32u ? 32u : 4u * ACCW;       // two tiles in flight per pipeline (power of two)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmO);
    mbar_init(w_full, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&p_full[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&p_empty[i], 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_n(tmem_slot, tmem_cols);
  for (int i = threadIdx.x; i < BN; i += kPersistThreads) sbias[i] = (a.bias && n0 + i < a.Cout) ? a.bias[n0 + i] : 0.f;
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    // the filter bank: 9 taps x nchunk boxes [BN x KC] behind one barrier.  Weights are never written inside a frame,
    // so this is issued BEFORE the PDL wait and overlaps the previous kernel's tail.
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, (uint32_t)(9 * a.nchunk) * (uint32_t)(BN * a.KC * 2) * nplane);
      for (int chunk = 0; chunk < a.nchunk; ++chunk)
        for (int tap = 0; tap < 9; ++tap)
          for (int pl = 0; pl < a.planes; ++pl)
            tma_load_3d(smem + (size_t)(chunk * 9 + tap) * a.b_bytes + (size_t)pl * a.b_plane, &tmB, w_full,
                        tap * a.Cin + chunk * a.KC, n0, pl);
    }
    __syncwarp();
  }
  pdl_trigger();
  pdl_wait();

  const int d = a.dil, pw = a.TW + 2 * a.dil;
  // dev: per-tile clock64 stamps [grid][32 tiles][8] (NULL in production)
  long long* dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 256 : nullptr;
#define OTVM_PSTAMP(i, k) do { if (dbg && (i) < 32 && lane == 0) dbg[(i) * 8 + (k)] = clock64(); } while (0)
  if (warp == 0) {
    // ===== patch producer =====
    int slot = 0; uint32_t ph = 0; int i = 0;
    const uint32_t p_tx = (uint32_t)((a.TW + 2 * d) * (a.TH + 2 * d)) * a.row_bytes * nplane;
    const uint32_t slot_bytes = nplane * a.patch_bytes;
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++i) {
      const int bq = a.tiles_x_magic ? (int)__umulhi((uint32_t)t, a.tiles_x_magic) : t;
      const int tx_i = t - bq * a.tiles_x;
      const int n_img = a.tiles_y_magic ? (int)__umulhi((uint32_t)bq, a.tiles_y_magic) : bq;
      const int ty_i = bq - n_img * a.tiles_y;
      const int cx = tx_i * a.TW - a.pad, cy = ty_i * a.TH - a.pad;
      for (int chunk = 0; chunk < a.nchunk; ++chunk) {
        mbar_wait(&p_empty[slot], ph ^ 1);
        if (elect_one()) {
          // the patch completes on the barrier of the pipeline that consumes this tile (i even / odd): each pipeline's
          // barrier of a slot then counts only its OWN fills, and the parity it waits for is exact at any ring depth
          uint64_t* fb = &p_full[(i & 1) * 8 + slot];
          mbar_arrive_expect_tx(fb, p_tx);
          for (int pl = 0; pl < a.planes; ++pl)
            tma_load_5d(smem + a.p_off + (size_t)slot * slot_bytes + (size_t)pl * a.patch_bytes, &tmA, fb,
                        chunk * a.KC, cx, cy, n_img, pl);
        }
        __syncwarp();
        if (++slot == a.na) { slot = 0; ph ^= 1; }
      }
      OTVM_PSTAMP(i, 0);
    }
  } else if (warp <= 2) {
    // ===== MMA issuers: warp 1 takes the CTA's even tiles (accumulator 0), warp 2 the odd ones (accumulator 1) =====
    const int g = warp - 1;
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    constexpr uint32_t idesc2 = make_idesc_bf16(128, 2 * BN);       // a0 . [b0 | b1] -> [main | cross] in one instruction
    const uint32_t idesc0 = nplane > 1 ? idesc2 : idesc;
    const uint64_t adesc0 = make_smem_desc(base + a.p_off, (uint32_t)pw * a.row_bytes, a.layout_type);
    const uint64_t bdesc0 = make_smem_desc(base, a.sbo, a.layout_type);
    const uint32_t b16 = a.b_bytes >> 4, patch16 = (nplane * a.patch_bytes) >> 4;
    const uint32_t aplane16 = a.patch_bytes >> 4, bplane16 = a.b_plane >> 4;
    const uint32_t kx16 = ((uint32_t)d * a.row_bytes) >> 4, ky16 = ((uint32_t)(d * pw) * a.row_bytes) >> 4;
    mbar_wait(w_full, 0);
    // patch-ring position of this pipeline's first tile; every tile consumes nchunk consecutive slots
    int slot = 0; uint32_t fills = 0;
    auto skip = [&](int n) { slot += n; while (slot >= a.na) slot -= a.na; };
    if (g) skip(a.nchunk);
    int j = 0;                                                   // tiles done by this pipeline
    for (int t = blockIdx.x + g * gridDim.x; t < a.ntiles; t += 2 * gridDim.x, ++j) {
      const int ab = 2 * g + (j & 1);                              // the pipeline alternates between its two accumulators
      mbar_wait(&acc_empty[ab], (((uint32_t)j >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
      tcgen05_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)ab * ACCW;
      OTVM_PSTAMP(2 * j + g, 1);
      for (int chunk = 0; chunk < a.nchunk; ++chunk) {
        // own "full" barrier of this slot, parity of the number of own fills so far (bit `slot` of fills)
        mbar_wait(&p_full[g * 8 + slot], (fills >> slot) & 1u);
        fills ^= 1u << slot;
        tcgen05_after_sync();
        if (chunk == a.nchunk - 1) OTVM_PSTAMP(2 * j + g, 2);
        if (elect_one()) {
          const uint64_t ad0 = adesc0 + (uint64_t)((uint32_t)slot * patch16);
          const uint64_t bd0 = bdesc0 + (uint64_t)((uint32_t)(chunk * 9) * b16);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t ad = ad0 + (uint64_t)((uint32_t)ky * ky16 + (uint32_t)kx * kx16);
              const uint64_t bd = bd0 + (uint64_t)((uint32_t)(ky * 3 + kx) * b16);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc0, (chunk | ky | kx | k) != 0);
              // remaining cross-plane products (a1 b0 | a1 b1, a0 b2, a2 b0) into the cross accumulator
              for (int pr = 2; pr < a.npair; ++pr) {
                const uint64_t ap = ad + ((kPairA >> (4 * pr)) & 15u) * aplane16, bp = bd + ((kPairB >> (4 * pr)) & 15u) * bplane16;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_bf16(d_tmem + BN, ap + (uint64_t)(2 * k), bp + (uint64_t)(2 * k), idesc, 1u);
              }
            }
          }
          umma_commit(&p_empty[slot]);                         // all 9 taps have read this patch
          if (chunk == a.nchunk - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
        if (++slot == a.na) slot = 0;
      }
      OTVM_PSTAMP(2 * j + g, 3);
      skip(a.nchunk);                                            // the other pipeline's tile
    }
  } else {
    // ===== epilogue: group g = warps 3+4g .. 6+4g; a warp owns TMEM lanes 32*(warp%4) .. +31 =====
    const int g = (warp - 3) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ty = r >> a.tw_shift, tx = r - (ty << a.tw_shift);
    const bool leader = (warp - 3 - 4 * g) == 0 && lane == 0;   // issues this group's TMA stores
    constexpr int CH = BN >= 32 ? 32 : 16;
    constexpr uint32_t ROWB = BN * 2, MASK = ROWB == 128 ? 7u : ROWB == 64 ? 3u : 1u;
    constexpr uint32_t STAGE_BYTES = 128u * ROWB;             // one plane of one output tile
    constexpr int NG = GN == GN_STATS ? BN / 2 : 1;            // GroupNorm groups (2 channels each) of this CTA
    const float slope = a.act == OTVM_ACT_NONE ? 1.f : a.act == OTVM_ACT_RELU ? 0.f : 0.01f;
    uint8_t* stg = smem + a.stage_off + (size_t)g * nplane * STAGE_BYTES;
    float s1[NG], s2[NG];
    if constexpr (GN == GN_STATS) {
#pragma unroll
      for (int u = 0; u < NG; ++u) { s1[u] = 0.f; s2[u] = 0.f; }
    }
    int j = 0;
    for (int t = blockIdx.x + g * gridDim.x; t < a.ntiles; t += 2 * gridDim.x, ++j) {
      const int bq = a.tiles_x_magic ? (int)__umulhi((uint32_t)t, a.tiles_x_magic) : t;
      const int tx_i = t - bq * a.tiles_x;
      const int n_img = a.tiles_y_magic ? (int)__umulhi((uint32_t)bq, a.tiles_y_magic) : bq;
      const int ty_i = bq - n_img * a.tiles_y;
      const int x0 = tx_i * a.TW, y0 = ty_i * a.TH;
      const bool valid = (y0 + ty) < a.Ho && (x0 + tx) < a.Wo;
      const int ab = 2 * g + (j & 1);
      mbar_wait(&acc_full[ab], ((uint32_t)j >> 1) & 1u);
      tcgen05_after_sync();
      if (q == 3) OTVM_PSTAMP(2 * j + g, 4);
      // the group's staging tile was handed to the TMA store of its previous tile: wait until that store has read it
      if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
#pragma unroll
      for (int c = 0; c < BN; c += CH) {
        uint32_t raw[CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ab * ACCW + (uint32_t)c;
        if constexpr (CH == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
        tmem_wait_ld();
        if (nplane > 1) {                                      // + cross-plane accumulator
          uint32_t raw2[CH];
          if constexpr (CH == 32) tmem_ld32(taddr + BN, raw2); else tmem_ld16(taddr + BN, raw2);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < CH; ++e) raw[e] = __float_as_uint(__uint_as_float(raw[e]) + __uint_as_float(raw2[e]));
        }
        float v[CH];
#pragma unroll
        for (int e = 0; e < CH; e += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + c + e);
          v[e] = __uint_as_float(raw[e]) + b4.x; v[e + 1] = __uint_as_float(raw[e + 1]) + b4.y;
          v[e + 2] = __uint_as_float(raw[e + 2]) + b4.z; v[e + 3] = __uint_as_float(raw[e + 3]) + b4.w;
        }
        if constexpr (GN == GN_STATS) {
          // statistics of the values GroupNorm will read back (one plane: rounded to bf16); rows outside the image count 0
#pragma unroll
          for (int e = 0; e < CH; e += 2) {
            const float q0 = !valid ? 0.f : nplane > 1 ? v[e] : __bfloat162float(__float2bfloat16_rn(v[e]));
            const float q1 = !valid ? 0.f : nplane > 1 ? v[e + 1] : __bfloat162float(__float2bfloat16_rn(v[e + 1]));
            s1[(c + e) >> 1] += q0 + q1;
            s2[(c + e) >> 1] += q0 * q0 + q1 * q1;
          }
        }
#pragma unroll
        for (int e = 0; e < CH; ++e) v[e] = fmaxf(v[e], 0.f) + slope * fminf(v[e], 0.f);
#pragma unroll
        for (int e8 = 0; e8 < CH / 8; ++e8) {
          const int cc = c + 8 * e8;
          uint32_t off = (uint32_t)r * ROWB + (uint32_t)cc * 2u;
          off ^= ((off >> 7) & MASK) << 4;
          float x8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x8[e] = v[8 * e8 + e];
          split_store8(stg + off, STAGE_BYTES, a.planes, x8);
        }
      }
      // this warp's share of the accumulator is staged: hand the TMEM buffer back to the pipeline's MMA warp
      tcgen05_before_sync();
      __syncwarp();
      if (q == 3) OTVM_PSTAMP(2 * j + g, 5);
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      fence_proxy_async_smem();                               // generic-proxy smem writes -> visible to the TMA engine
      asm volatile("bar.sync %0, 128;" ::"r"(3 + g) : "memory");
      if (leader) {
        for (uint32_t pl = 0; pl < nplane; ++pl)
          tma_store_5d(&tmO, stg + pl * STAGE_BYTES, n0, x0, y0, n_img, (int)pl);      // clips ragged tiles
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (q == 3) OTVM_PSTAMP(2 * j + g, 6);
    }
    if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if constexpr (GN == GN_STATS) {
      // Both groups have seen their last accumulator, so every MMA of this CTA has completed and (after the wait above)
      // every store has read its staging tile: the patch ring + staging region is dead.  Park the per-row partials
      // there, add up the 256 rows per (group, moment) in fp64 and issue this CTA's 64 atomics.
      asm volatile("bar.sync 5, 256;" ::: "memory");
      float* sred = reinterpret_cast<float*>(smem + a.p_off);
      const int row = g * 128 + r;
#pragma unroll
      for (int u = 0; u < NG; ++u) { sred[u * kSredPitch2 + row] = s1[u]; sred[(NG + u) * kSredPitch2 + row] = s2[u]; }
      asm volatile("bar.sync 5, 256;" ::: "memory");
      const int e = threadIdx.x - 96;
      if (e < 2 * NG) {
        // (fp64 adds are ~64x slower than fp32 on this part: 8 independent fp32 chains, fixed order, combined in fp64)
        const float* rowp = sred + e * kSredPitch2;          // e = which * NG + group
        float p8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int k = 0; k < 256; k += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) p8[u] += rowp[k + u];
        }
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += (double)p8[u];
        atomicAdd(&a.gn_stats[(n0 / 2 + e % NG) * 2 + e / NG], acc);
      }
    }
    tcgen05_before_sync();
  }
#undef OTVM_PSTAMP
  __syncthreads();
  if (warp == 1) {
    tcgen05_after_sync();
    tmem_dealloc_n(tmem_base, tmem_cols);
  }
}

// split-K second pass: sum the fp32 partial tiles and run the fused epilogue (bias, GroupNorm statistics of the
// stored values, residual, activation, optional ReLU'd second output) on the 4 channels each thread owns
template <typename T>
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ ws, int S, int64_t M, int Cout,
                                                            const float* __restrict__ bias, cptr_t<T> res,
                                                            int64_t res_ld, int act, void* __restrict__ out, int64_t out_ps,
                                                            int64_t out_cs, int out_f32, ptr_t<T> out_relu,
                                                            int64_t out_relu_ld, double* __restrict__ gn_stats, int64_t ps,
                                                            int gn_cg) {
  pdl_sync();                                  // PDL contract (common.cuh)
  // fp64 partials: (near-)exact sums, so the atomic order does not change the result.  One copy per warp: shared-memory
  // fp64 atomics are compare-and-swap loops, and eight warps retrying on the same 64 words made this kernel 3x slower
  // than its memory traffic (21.7 us for 12 MB behind the 3072 -> 256 convolution)
  __shared__ double sstat_w[8][32][2];
  double (*sstat)[2] = sstat_w[threadIdx.x >> 5];
  const int c4n = Cout >> 2;
  const int64_t total = M * c4n;
  const int cg = gn_stats ? gn_cg : 1;
  if (gn_stats) {
    sstat[threadIdx.x & 31][0] = 0.0; sstat[threadIdx.x & 31][1] = 0.0;
    __syncthreads();
  }
  // channel-major destinations (the value bank, out_cs = row length): consecutive threads take consecutive PIXELS of one
  // channel quad, so their 2-byte stores are contiguous (pixel-fastest threads scattered them a row apart: 16 vs 5 us)
  const bool pix_fast = out_cs != 1;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t p; int c;
    if (pix_fast) { const int64_t q = idx / M; p = idx - q * M; c = (int)q * 4; }
    else { p = idx / c4n; c = (int)(idx - p * c4n) * 4; }
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int z = 0; z < S; ++z) {
      const float4 t = *reinterpret_cast<const float4*>(ws + ((int64_t)z * M + p) * Cout + c);
      v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
    }
    if (bias) { v[0] += bias[c]; v[1] += bias[c + 1]; v[2] += bias[c + 2]; v[3] += bias[c + 3]; }
    if (gn_stats) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float q0 = stored<T>(v[2 * h]), q1 = stored<T>(v[2 * h + 1]);
        if (cg >= 2) {
          atomicAdd(&sstat[(c + 2 * h) / cg][0], (double)q0 + (double)q1);
          atomicAdd(&sstat[(c + 2 * h) / cg][1], (double)q0 * q0 + (double)q1 * q1);
        } else {
          atomicAdd(&sstat[c + 2 * h][0], (double)q0); atomicAdd(&sstat[c + 2 * h][1], (double)q0 * q0);
          atomicAdd(&sstat[c + 2 * h + 1][0], (double)q1); atomicAdd(&sstat[c + 2 * h + 1][1], (double)q1 * q1);
        }
      }
    }
    if (res) {
      float r[4];
      load4(res + (p * res_ld + c), r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], act);
    if (out_f32) {
      float* o = static_cast<float*>(out) + p * out_ps + (int64_t)c * out_cs;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[(int64_t)j * out_cs] = v[j];
    } else {
      const ptr_t<T> o = mkptr<T>(out, ps) + (p * out_ps + (int64_t)c * out_cs);
      if (out_cs == 1 && aligned4(o)) store4(o, v);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) st1(o, (int64_t)j * out_cs, v[j]);
      }
    }
    if (out_relu) {
      float rl[4] = {fmaxf(v[0], 0.f), fmaxf(v[1], 0.f), fmaxf(v[2], 0.f), fmaxf(v[3], 0.f)};
      store4(out_relu + (p * out_relu_ld + c), rl);
    }
  }
  if (gn_stats) {
    __syncthreads();
    if (threadIdx.x < 64) {
      const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sstat_w[w][g][which];
      atomicAdd(&gn_stats[g * 2 + which], t);
    }
  }
}

long long* g_conv_dbg = nullptr;   // dev hook (otvm_debug_set_conv_timestamps)

// ---------------------------------------------------------------------------------------------------------
static int g_conv_budget_kb = 0;
static int g_conv_halo = -2;              // -2 unset, -1 auto (default), 0 off, 1 forced on
static int conv_halo_mode() {
  if (g_conv_halo == -2) { const char* e = getenv("OTVM_CONV_HALO"); g_conv_halo = e ? atoi(e) : -1; }
  return g_conv_halo;
}

static int g_conv_ksub = -2;              // K-chunks per barrier pair on one-wave grids: 2 (default) or 1
static int conv_ksub_mode() {
  if (g_conv_ksub == -2) { const char* e = getenv("OTVM_CONV_KSUB"); g_conv_ksub = e ? atoi(e) : 2; }
  return g_conv_ksub;
}
static long long g_conv_persist_launches = 0;
static long long g_conv_cluster_launches = 0;
static int g_conv_persist = -2;           // -2 unset, -1 auto (default), 0 off, 1 whenever the shape allows
static int conv_persist_mode() {
  if (g_conv_persist == -2) { const char* e = getenv("OTVM_CONV_PERSIST"); g_conv_persist = e ? atoi(e) : -1; }
  return g_conv_persist;
}

static int g_conv_small_bn = -2;          // narrower tiles on SM-starved grids (default 1; env OTVM_CONV_SMALL_BN)
static int conv_small_bn_mode() {
  if (g_conv_small_bn == -2) { const char* e = getenv("OTVM_CONV_SMALL_BN"); g_conv_small_bn = e ? atoi(e) : 1; }
  return g_conv_small_bn;
}

// Tile width (output channels per CTA).  128 wherever the grid can fill the SMs; a grid that would leave more than half
// of them idle takes 64-channel tiles instead: twice the CTAs, and each CTA's K loop is paced by its single MMA-issuing
// thread at ~77 cycles per N=128 instruction against ~50 per N=64 one (scripts/conv_ts3.py), with half the epilogue.
static int pick_bn(const otvm_conv_params* p) {
  const int Cout = p->Cout;
  int bn = Cout >= 128 ? 128 : Cout > 32 ? 64 : Cout > 16 ? 32 : 16;
  if (bn == 128 && conv_small_bn_mode() > 0) {
    const int Ho = (p->H + 2 * p->pad - p->dil * (p->KH - 1) - 1) / p->stride + 1;
    const int Wo = (p->W + 2 * p->pad - p->dil * (p->KW - 1) - 1) / p->stride + 1;
    int tw = 8; while (tw * 2 <= Wo && tw < 128) tw *= 2;
    const int64_t tiles = (int64_t)ceil_div(Wo, tw) * ceil_div(Ho, 128 / tw) * p->N;
    const int kc = p->Cin % 64 == 0 ? 64 : p->Cin % 32 == 0 ? 32 : 16;
    const int num_k = p->KH * p->KW * (p->Cin / kc);            // long-K layers keep 128 and slice K instead (split-K)
    if (tiles * ceil_div(Cout, 128) * 2 <= sm_count() && num_k < 48) bn = 64;
    // (32-channel tiles on the even smaller grids measured slower: 429.6 vs 433.2 frames/s)
  }
  return bn;
}

static bool aligned_view(const void* ptr, int64_t ld) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0;
}

// epilogue variant for this problem, or -1 when no compiled variant covers it
static int conv_tc_epi(const otvm_conv_params* p, int bn) {
  const int ch = bn >= 32 ? 32 : 16;
  int epi = 0;
  if (p->res) {
    if (!aligned_view(p->res, p->res_ld) || p->Cout % ch != 0) return -1;
    epi |= EPI_RES;
  }
  const bool tma_ok = !p->out_f32 && p->out_cs == 1 && aligned_view(p->out, p->out_ps) && p->Cout % 8 == 0;
  if (!tma_ok) {
    if (p->out_relu || p->gn_stats || p->res) return -1;
    return EPI_DIRECT;
  }
  if (p->out_relu) {
    if (!aligned_view(p->out_relu, p->out_relu_ld) || p->gn_stats) return -1;
    epi |= EPI_RELU2;
  }
  if (p->gn_stats && (epi & EPI_RELU2)) return -1;
  return epi;
}

bool conv2d_tc_supported(const otvm_conv_params* p) {
  const int fmt = dtype_fmt(p->dtype);
  if ((fmt != OTVM_BF16 && fmt != OTVM_BF16X2 && fmt != OTVM_BF16X3) || p->stride > 2 || p->relu_in) return false;
  if (p->Cin % 16 != 0 || p->in_ld % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p->in) & 15) || (reinterpret_cast<uintptr_t>(p->weight) & 15)) return false;
  const int Ho = (p->H + 2 * p->pad - p->dil * (p->KH - 1) - 1) / p->stride + 1;
  const int Wo = (p->W + 2 * p->pad - p->dil * (p->KW - 1) - 1) / p->stride + 1;
  if (Wo < 8 || Ho < 1) return false;
  if (((int64_t)p->KH * p->KW * p->Cin * 2) % 16 != 0) return false;
  if (p->groups > 1 && (p->N != p->groups || p->gn_stats)) return false;    // grouped: one image per filter bank, no GroupNorm
  const int gcg = p->gn_group_ch > 0 ? p->gn_group_ch : p->Cout / 32;      // channels per GroupNorm group
  if (p->gn_stats && (p->N != 1 || gcg < 1 || p->Cout % gcg != 0 || p->Cout / gcg > 32)) return false;
  const int bn = pick_bn(p);
  if (conv_tc_epi(p, bn) < 0) return false;
  if (p->gn_stats) {            // a GroupNorm group must not straddle tiles / 16-column chunks irregularly
    const int cg = gcg;
    if (bn < 32) return false;
    if (cg > 32 && (cg % 32 != 0 || bn % cg != 0)) return false;
    if (cg <= 32 && 32 % cg != 0) return false;
    if (bn / (cg < 32 ? cg : 32) > 32) return false;          // at most 32 (slot, row) partial columns in smem
  }
  const int dev = current_device();
  return dev >= 0 && otvm_device_is_sm100(dev) == 1;
}

// block size: tiles of >= 64 channels (two or four 32-column chunks) get a second group of four epilogue warps
// (OTVM_CONV_EPI_GROUPS=1 keeps one)
template <int BN>
static int conv_threads() {
  static const int groups = getenv("OTVM_CONV_EPI_GROUPS") ? atoi(getenv("OTVM_CONV_EPI_GROUPS")) : 2;
  return (BN >= 64 && groups >= 2) ? kConvThreadsMax : kConvThreads;
}

template <int BN, int GN, int EPI, bool HALO>
static int launch_conv_tc_cluster(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                                  const CUtensorMap& tmR, const ConvTcArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
  OTVM_CUDA_CHECK((ensure_dynamic_smem<conv_tc_kernel<BN, GN, EPI, HALO, true>>(220 * 1024)));
  launch_k_cluster(conv_tc_kernel<BN, GN, EPI, HALO, true>, grid, conv_threads<BN>(), smem, s, (unsigned)a.cluster_k, tmA, tmB,
                   tmO, tmR, a);
  OTVM_LAUNCH_CHECK();
  ++g_conv_cluster_launches;
  return OTVM_OK;
}

template <int BN, int GN, int EPI, bool HALO>
static int launch_conv_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                          const CUtensorMap& tmR, const ConvTcArgs& a, dim3 grid, size_t smem, cudaStream_t s, bool dry_run) {
  if constexpr (GN == GN_NONE && (EPI & EPI_DIRECT) == 0 && BN >= 64) {
    if (a.cluster_k > 1) return dry_run ? OTVM_OK : launch_conv_tc_cluster<BN, GN, EPI, HALO>(tmA, tmB, tmO, tmR, a, grid, smem, s);
  }
  OTVM_CUDA_CHECK((ensure_dynamic_smem<conv_tc_kernel<BN, GN, EPI, HALO>>(220 * 1024)));
  if constexpr (GN == GN_FUSED) {
    // the grid barrier needs every CTA resident at once.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for
    // every tcgen05 kernel on this toolkit, whatever its footprint -- ncu shows 2-4 resident CTAs -- so residency is
    // derived from the variant's own registers / shared memory / TMEM columns, capped at 2 for margin.)
    static int regs = 0;
    if (regs == 0) {
      cudaFuncAttributes fa;
      OTVM_CUDA_CHECK(cudaFuncGetAttributes(&fa, conv_tc_kernel<BN, GN, EPI, HALO>));
      regs = fa.numRegs > 0 ? fa.numRegs : 255;
    }
    const int regs_per_warp = ((regs * 32 + 255) / 256) * 256;
    const int by_regs = 65536 / (regs_per_warp * (conv_threads<BN>() / 32));
    const int by_smem = (int)((227u * 1024u) / (smem + 1024));
    const int tmem_cols = (BN < 32 ? 32 : BN) * (a.planes > 1 ? 2 : 1);
    int occ = by_regs < by_smem ? by_regs : by_smem;
    if (occ > 512 / tmem_cols) occ = 512 / tmem_cols;
    if (occ > 2) occ = 2;
    if (getenv("OTVM_DEBUG_OCC")) fprintf(stderr, "conv_tc<%d,%d,%d,%d> smem=%zu regs=%d grid=(%u,%u,%u) occ=%d\n", BN, GN, EPI, (int)HALO, smem, regs, grid.x, grid.y, grid.z, occ);
    if ((int64_t)grid.x * grid.y * grid.z > (int64_t)occ * sm_count()) return OTVM_ERR_UNSUPPORTED;
  }
  if (dry_run) return OTVM_OK;
  launch_k(conv_tc_kernel<BN, GN, EPI, HALO>, grid, conv_threads<BN>(), smem, s, tmA, tmB, tmO, tmR, a);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}

template <int BN, int GN, int KSTEPS>
static int launch_conv_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvTcArgs& a,
                               dim3 grid, size_t smem, cudaStream_t s) {
  OTVM_CUDA_CHECK((ensure_dynamic_smem<conv_tc_persist_kernel<BN, GN, KSTEPS>>(226 * 1024)));
  launch_k(conv_tc_persist_kernel<BN, GN, KSTEPS>, grid, kPersistThreads, smem, s, tmA, tmB, tmO, a);
  OTVM_LAUNCH_CHECK();
  return OTVM_OK;
}
template <int BN, int GN>
static int launch_conv_persist_k(int ksteps, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                                 const ConvTcArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
  return ksteps == 4 ? launch_conv_persist<BN, GN, 4>(tmA, tmB, tmO, a, grid, smem, s)
                     : launch_conv_persist<BN, GN, 2>(tmA, tmB, tmO, a, grid, smem, s);
}

// Persistent patch-mode variant (conv_tc_persist_kernel) when the layer qualifies; returns 1 when it launched, 0 when the
// caller should take the one-tile-per-CTA kernel, < 0 on error.
static int try_conv_persist(const otvm_conv_params* p, const ConvTcArgs& a0, int bn, int epi, int gn,
                            const CUtensorMap& tmA, const CUtensorMap& tmB0, const CUtensorMap& tmO0, cudaStream_t s) {
  const int mode = conv_persist_mode();
  if (mode == 0 || !a0.halo || p->Cout != bn || bn > 64 || epi != 0 || gn == GN_FUSED || p->groups > 1) return 0;
  if (gn == GN_STATS && (p->Cout != 64 || (p->gn_group_ch > 0 && p->gn_group_ch != 2))) return 0;
  if (a0.KC != 64 && a0.KC != 32) return 0;
  ConvTcArgs a = a0;
  a.ntiles = a.tiles_x * a.tiles_y * p->N;
  // auto: only where a CTA gets several tiles (measured: 11.2 vs 16.1 us at 3.5 tiles per SM, 64->64 at 256^2);
  // smaller grids are latency-bound by their single wave and keep the deeper per-tile rings
  if (mode < 0 && a.ntiles < 2 * sm_count()) return 0;
  // channels per CTA: split operands double the filter bank, so a 64-channel layer runs as two 32-channel halves
  // (blockIdx.y); each half's CTAs walk every tile and re-read the patches (from L2)
  const int pbn = (a.planes > 1 && bn == 64) ? 32 : bn;
  const int ny = bn / pbn;
  a.b_plane = (uint32_t)pbn * a.KC * 2;
  if (a.b_plane % 1024u != 0) return 0;
  a.b_bytes = (uint32_t)a.planes * a.b_plane;
  const uint32_t slot_bytes = (uint32_t)a.planes * a.patch_bytes;
  const uint32_t w_bytes = (uint32_t)(9 * a.nchunk) * a.b_bytes;
  const uint32_t stage_bytes = 2u * (uint32_t)a.planes * 128u * (uint32_t)pbn * 2u;
  // 200 KB; up to 220 KB when that buys the ring a slot per K-chunk of a tile (96 -> 64 / 32 with two planes: 108 KB of
  // filters + 32 KB of staging leave 2 slots of 24 KB in 200 KB, 3 in 220 KB)
  uint32_t budget = 200u * 1024u;
  if (w_bytes + stage_bytes + 2 * slot_bytes > budget) return 0;
  if (4u * (uint32_t)pbn * (a.planes > 1 ? 2u : 1u) > 512u) return 0;           // tensor-memory columns
  if ((int)((budget - w_bytes - stage_bytes) / slot_bytes) < a.nchunk && w_bytes + stage_bytes + (uint32_t)a.nchunk * slot_bytes <= 220u * 1024u)
    budget = 220u * 1024u;
  int na = (int)((budget - w_bytes - stage_bytes) / slot_bytes);
  if (na > 8) na = 8;
  // Ring depth: each pipeline has its own "full" barrier per slot (the producer completes a patch on the barrier of the
  // pipeline that consumes the tile), so the parity a pipeline waits for counts only its own fills and is exact at any
  // depth >= 2.  (Round 2 first shared one barrier per slot, which needed na >= nchunk + 1 and sent the 3-chunk 96 -> 64 /
  // 96 -> 32 layers to the one-tile kernel: 155 + 96 us per frame.)
  if (na < 2) return 0;
  // the final GroupNorm reduction parks [pbn][257] floats in the (then dead) patch ring + staging region
  if (gn == GN_STATS && (uint32_t)na * slot_bytes + stage_bytes < (uint32_t)pbn * kSredPitch2 * sizeof(float)) return 0;
  a.na = na;
  a.p_off = w_bytes;                                   // b_bytes and patch_bytes are multiples of 1024
  a.stage_off = a.p_off + (uint32_t)na * slot_bytes;
  a.aux_off = a.stage_off + stage_bytes;
  const size_t smem = (size_t)a.aux_off + 1024 + 40 * 8 + 16 + 64 * sizeof(float) + 64;
  int gx = sm_count() / ny;
  if (gx > a.ntiles) gx = a.ntiles;
  CUtensorMap tmB = tmB0, tmO = tmO0;
  if (pbn != bn) {                                     // boxes of pbn channels
    const CUtensorMapSwizzle swz = a.KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const uint64_t K = (uint64_t)p->KH * p->KW * p->Cin;
    uint64_t dims[3] = {K, (uint64_t)p->Cout, (uint64_t)a.planes};
    uint64_t str[2] = {K * 2, a.planes > 1 ? (uint64_t)p->w_plane_stride * 2 : K * 2 * (uint64_t)p->Cout};
    uint32_t box[3] = {(uint32_t)a.KC, (uint32_t)pbn, 1};
    int rc = make_tmap_bf16(&tmB, p->weight, 3, dims, str, box, swz);
    if (rc) return rc;
    const CUtensorMapSwizzle oswz = pbn == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : pbn == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    uint64_t odims[5] = {(uint64_t)p->Cout, (uint64_t)a.Wo, (uint64_t)a.Ho, (uint64_t)p->N, (uint64_t)a.planes};
    uint32_t obox[5] = {(uint32_t)pbn, (uint32_t)a.TW, (uint32_t)a.TH, 1, 1};
    uint64_t ostr[4] = {(uint64_t)p->out_ps * 2, (uint64_t)a.Wo * p->out_ps * 2, (uint64_t)a.Ho * a.Wo * p->out_ps * 2,
                        a.planes > 1 ? (uint64_t)a.act_plane * 2 : (uint64_t)p->N * a.Ho * a.Wo * p->out_ps * 2};
    rc = make_tmap_bf16(&tmO, p->out, 5, odims, ostr, obox, oswz);
    if (rc) return rc;
  }
  const dim3 grid(gx, ny);
  int rc;
  const int ks = a.KC / 16;
  if (pbn == 64) rc = gn == GN_STATS ? launch_conv_persist_k<64, GN_STATS>(ks, tmA, tmB, tmO, a, grid, smem, s)
                                     : launch_conv_persist_k<64, GN_NONE>(ks, tmA, tmB, tmO, a, grid, smem, s);
  else if (pbn == 32) rc = gn == GN_STATS ? launch_conv_persist_k<32, GN_STATS>(ks, tmA, tmB, tmO, a, grid, smem, s)
                                          : launch_conv_persist_k<32, GN_NONE>(ks, tmA, tmB, tmO, a, grid, smem, s);
  else rc = launch_conv_persist_k<16, GN_NONE>(ks, tmA, tmB, tmO, a, grid, smem, s);
  if (rc == OTVM_OK) ++g_conv_persist_launches;
  return rc == OTVM_OK ? 1 : rc;
}

template <int BN>
static int dispatch_conv_tc(int gn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                            const CUtensorMap& tmR, const ConvTcArgs& a, dim3 grid, size_t smem, cudaStream_t s,
                            bool dry_run = false) {
#define OTVM_CASE(G, E)                                                                                     \
  if (gn == G && epi == (E))                                                                                \
    return a.halo ? launch_conv_tc<BN, G, (E), true>(tmA, tmB, tmO, tmR, a, grid, smem, s, dry_run)         \
                  : launch_conv_tc<BN, G, (E), false>(tmA, tmB, tmO, tmR, a, grid, smem, s, dry_run);
  OTVM_CASE(GN_NONE, 0) OTVM_CASE(GN_NONE, EPI_RES) OTVM_CASE(GN_NONE, EPI_RELU2) OTVM_CASE(GN_NONE, EPI_RES | EPI_RELU2)
  OTVM_CASE(GN_NONE, EPI_DIRECT)
  if constexpr (BN >= 32) {
    OTVM_CASE(GN_STATS, 0) OTVM_CASE(GN_STATS, EPI_RES) OTVM_CASE(GN_FUSED, 0) OTVM_CASE(GN_FUSED, EPI_RES)
  }
#undef OTVM_CASE
  return OTVM_ERR_UNSUPPORTED;
}

int conv2d_tc(const otvm_conv_params* p, cudaStream_t s, bool dry_run) {
  ConvTcArgs a;
  a.ntiles = 0; a.p_off = 0; a.stage_off = 0;
  const int fmt = dtype_fmt(p->dtype);
  a.planes = dtype_planes(p->dtype);
  a.npair = a.planes == 1 ? 1 : a.planes == 2 ? 3 : 6;
  a.act_plane = dtype_plane_stride(p->dtype);
  if (a.planes > 1 && (a.act_plane <= 0 || p->w_plane_stride <= 0)) return OTVM_ERR_ARG;
  (void)fmt;
  a.N = p->N; a.H = p->H; a.W = p->W; a.Cin = p->Cin; a.Cout = p->Cout; a.KH = p->KH; a.KW = p->KW;
  a.pad = p->pad; a.dil = p->dil; a.stride = p->stride;
  a.Ho = (p->H + 2 * p->pad - p->dil * (p->KH - 1) - 1) / p->stride + 1;
  a.Wo = (p->W + 2 * p->pad - p->dil * (p->KW - 1) - 1) / p->stride + 1;
  a.KC = p->Cin % 64 == 0 ? 64 : p->Cin % 32 == 0 ? 32 : 16;
  a.nchunk = p->Cin / a.KC;
  a.row_bytes = (uint32_t)a.KC * 2;
  // halo mode: 3x3 stride-1 convolutions read ONE input patch per K-chunk for all 9 taps (16 x 8 output tile, so that
  // every 8-row core-matrix group of the UMMA operand is one tile row and SBO = patch row pitch)
  // Measured on B200 (scripts/bench_conv2.py, profiles/r01c_conv_sweep.txt): the patch mode wins where the per-tap
  // boxes are small and numerous (Cout <= 64, Cin <= 128: the full-resolution decoder / refinement layers, 1.4-1.8x);
  // with BN = 128 or many K-chunks the 9-tap granularity of the patch ring costs more than the saved L2 traffic.
  const int halo_mode = conv_halo_mode();              // -1 auto, 0 never, 1 whenever the shape allows
  a.halo = halo_mode != 0 && p->KH == 3 && p->KW == 3 && p->stride == 1 && p->dil <= 4 &&
           (halo_mode == 1 || (p->Cout <= 64 && p->Cin <= 128));
  const int bn = pick_bn(p);
  if (a.halo) {
    a.TW = 8; a.TH = 16;
    a.patch_bytes = (((uint32_t)(a.TW + 2 * p->dil) * (a.TH + 2 * p->dil) * a.row_bytes) + 1023u) & ~1023u;
    a.na = a.nchunk > 1 ? 2 : 1;
    // the patch ring (planes x na patches) + at least two weight stages must fit one CTA; else one box per tap
    const uint32_t wstage = (uint32_t)a.planes * (((uint32_t)bn * a.KC * 2 + 1023u) & ~1023u);
    if ((uint32_t)a.na * a.planes * a.patch_bytes + 2 * wstage > 200u * 1024u) a.na = 1;
    if ((uint32_t)a.na * a.planes * a.patch_bytes + 2 * wstage > 200u * 1024u) a.halo = 0;
  }
  if (!a.halo) {
    int tw = 8; while (tw * 2 <= a.Wo && tw < 128) tw *= 2;
    a.TW = tw; a.TH = 128 / tw;
    a.patch_bytes = 0; a.na = 0;
  }
  a.tiles_x = ceil_div(a.Wo, a.TW); a.tiles_y = ceil_div(a.Ho, a.TH);
  a.tw_shift = 0; while ((1 << a.tw_shift) < a.TW) ++a.tw_shift;
  auto magic = [](int d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
  a.tiles_x_magic = magic(a.tiles_x); a.tiles_y_magic = magic(a.tiles_y);     // exact for n * d < 2^32 (n < 2^16 tiles)
  // Multi-wave grids with split operands: the usual 200 KB ring leaves ONE CTA per SM, so nothing overlaps a CTA's
  // (long: two accumulators, plane split, two staging tiles) epilogue, and a GroupNorm-fused request with 149..296 CTAs
  // has no co-resident wave at all (-> a separate gn_apply pass over the whole output).  Half-depth K chunks (32) and a
  // ~100 KB footprint keep two CTAs per SM.  Measured: 231.6 -> 235.7 frames/s from the fused-GroupNorm layers alone (8
  // gn_apply passes fewer), -> 239.7 for every multi-wave layer (512 -> 2048 at 64^2: 58 -> 44 us).
  // OTVM_CONV_SMALL_RING=0 restores the deep ring.
  const int64_t ctas0 = (int64_t)a.tiles_x * a.tiles_y * p->N * ceil_div(p->Cout, bn);
  static const int small_ring_on = getenv("OTVM_CONV_SMALL_RING") ? atoi(getenv("OTVM_CONV_SMALL_RING")) : 1;
  const bool small_ring = small_ring_on && a.planes > 1 && !a.halo && a.KC == 64 && ctas0 > sm_count();
  if (small_ring) { a.KC = 32; a.nchunk = p->Cin / a.KC; a.row_bytes = (uint32_t)a.KC * 2; }
  a.a_plane = a.halo ? 0u : 128u * a.KC * 2;
  a.b_plane = ((uint32_t)bn * a.KC * 2 + 1023u) & ~1023u;
  a.a_bytes = (uint32_t)a.planes * a.a_plane;
  a.b_bytes = (uint32_t)a.planes * a.b_plane;
  a.b_off = (uint32_t)a.na * (uint32_t)a.planes * a.patch_bytes;
  // [hi | lo] weight planes are one contiguous 2 BN-row tile when a plane is not padded; BN >= 32: accumulators BN apart
  a.wide = a.planes > 1 && bn >= 32 && a.b_plane == (uint32_t)bn * a.KC * 2;
  a.sbo = 8u * a.KC * 2;
  a.layout_type = a.KC == 64 ? 2u : a.KC == 32 ? 4u : 6u;
  const uint32_t stage = a.a_bytes + a.b_bytes;
  // two CTAs per SM (<= 96 KB of ring each) when the grid is larger than one wave; a grid that fits in one
  // wave is latency-bound instead, so it gets a deeper ring (up to ~190 KB, one CTA per SM)
  const int64_t ctas = (int64_t)a.tiles_x * a.tiles_y * p->N * ceil_div(p->Cout, bn);
  const int budget_kb = g_conv_budget_kb;   // dev override of the multi-wave ring budget (KB)
  // multi-wave grids: narrow tiles (BN <= 64) have short K loops and are bound by per-CTA latency, so they trade ring
  // depth for residency (4+ CTAs per SM); patch mode keeps 3 weight stages behind the patch ring
  // split operands: a stage is `planes` times larger, so the ring always takes one CTA's worth of shared memory (ring()
  // clamps it to the K iterations: short-K layers stay small and still co-reside)
  uint32_t budget = ctas <= sm_count() ? (a.planes > 1 ? 200u * 1024u : 190u * 1024u)
                  : budget_kb > 0 ? (uint32_t)budget_kb * 1024u
                  : a.halo ? a.b_off + 3 * a.b_bytes
                  : small_ring ? 98u * 1024u
                  : a.planes > 1 ? 200u * 1024u
                  : bn <= 64 ? 48u * 1024u : 96u * 1024u;
  if (a.halo && a.b_off + 3 * a.b_bytes > budget) budget = 200u * 1024u;
  const int num_k = a.KH * a.KW * a.nchunk;
  // K-chunk groups: a one-wave grid is paced by the per-stage round trip of its single-thread loops (wait, elect,
  // issue, commit: ~420 cycles per 128x128x64 stage against 256 cycles of tensor work, scripts/conv_ts3.py), so there
  // one barrier pair covers TWO consecutive K-chunks (8 MMAs / 2+2 TMA boxes per round trip).  Multi-wave grids keep
  // single chunks: with two CTAs per SM the tensor pipe is already shared at ~2 x 256 cycles per stage pair.
  auto ring = [&](int ksub_, int iters) {
    uint32_t bud = budget;
    if (ksub_ == 2 && bud < 200u * 1024u) bud = 200u * 1024u;
    int n = (int)((bud - a.b_off) / (stage * ksub_));
    if (n > 8) n = 8;
    const int groups = ceil_div(iters, ksub_);
    if (n > groups) n = groups;
    if (n < 2) n = groups < 2 ? 1 : 2;      // (a one-iteration K loop needs one stage: the CTA stays small and co-resides)
    return n;
  };
  int ksub = (!a.halo && a.planes == 1 && ctas <= sm_count() && num_k >= 8 && conv_ksub_mode() >= 2) ? 2 : 1;
  // split-K: a grid that fills less than half of the SMs walks K serially at TMA/L2 latency; slice K across
  // blockIdx.z, write fp32 partial tiles to the caller's workspace and finish with a small fused-epilogue kernel
  int nsplit = 1;
  const int64_t Mtot = (int64_t)p->N * a.Ho * a.Wo;
  // (measured with split operands, 3 plane products per K iteration: lowering this threshold to num_k * 3 >= 48 made the
  // 16-36 iteration layers SLOWER, 206 vs 224 frames/s -- they are bound by their fixed costs, not by the K walk, and the
  // workspace round trip + finish kernel add to those)
  const int num_kp = num_k;
  if (p->workspace && p->groups <= 1 && ctas * 2 <= sm_count() && num_kp >= 48 && p->Cout % 4 == 0 &&
      (!p->res || p->res_ld % 4 == 0) && (!p->out_relu || p->out_relu_ld % 4 == 0)) {
    nsplit = (int)(sm_count() / ctas);
    if (nsplit > num_kp / 12) nsplit = num_kp / 12;        // >= 12 K iterations' worth per slice: the extra pass must pay off
    if (nsplit > num_k) nsplit = num_k;
    if (nsplit > 16) nsplit = 16;
    while (nsplit > 1 && (int64_t)nsplit * Mtot * p->Cout * 4 > p->workspace_bytes) --nsplit;
  }
  // Cluster split-K for the layers too short for the workspace route (< 48 K iterations) whose grid leaves at least half
  // of the SMs idle: the K slices (blockIdx.z) of a tile form a thread-block cluster and meet in rank 0 through distributed
  // shared memory -- no workspace, no finish kernel, the epilogue stays in the kernel.  Plain / residual / second-output
  // epilogues of >= 64-channel tiles only (the STM encoders' and decoder's 1/16-resolution layers: 256 -> 256 3x3 at 32^2,
  // 36 iterations on 32 or 64 CTAs).  OTVM_CONV_CLUSTER_K=0 disables.
  static const int cluster_on = getenv("OTVM_CONV_CLUSTER_K") ? atoi(getenv("OTVM_CONV_CLUSTER_K")) : 1;
  a.cluster_k = 1;
  if (cluster_on && nsplit == 1 && !a.halo && bn >= 64 && !p->gn_stats && !p->out_f32 && p->out_cs == 1 &&
      ctas * 2 <= sm_count() && num_k >= 16 && conv_tc_epi(p, bn) >= 0 && !(conv_tc_epi(p, bn) & EPI_DIRECT)) {
    int cs = (int)(sm_count() / ctas);
    if (cs > 8) cs = 8;                                    // portable cluster size
    if (cs > num_k / 8) cs = num_k / 8;                    // >= 8 K iterations per slice
    if (cs > 1) { nsplit = cs; a.cluster_k = cs; }
  }
  if (a.halo) {                            // halo mode slices whole K-chunks (9 taps each)
    const int cps = ceil_div(a.nchunk, nsplit);
    a.k_per_split = cps * 9;
  } else {
    a.k_per_split = ceil_div(num_k, nsplit);
  }
  nsplit = ceil_div(num_k, a.k_per_split);
  if (a.cluster_k > 1) a.cluster_k = nsplit;               // (no empty slice)
  const bool clustered = a.cluster_k > 1;
  a.split_stride = Mtot * p->Cout;
  if (a.k_per_split < 4) ksub = 1;
  a.ksub = ksub;
  a.nstage = ring(ksub, a.k_per_split);
  const int nstage = a.nstage;
  a.bias = p->bias; a.out = p->out; a.out_ps = p->out_ps; a.out_cs = p->out_cs;
  a.res = static_cast<const bf16*>(p->res); a.res_ld = p->res_ld;
  a.out_relu = static_cast<bf16*>(p->out_relu); a.out_relu_ld = p->out_relu_ld;
  a.act = p->act; a.out_f32 = p->out_f32; a.gn_stats = p->gn_stats;
  a.gn_gamma = p->gn_gamma; a.gn_beta = p->gn_beta; a.gn_eps = p->gn_eps;
  a.w_group_rows = p->groups > 1 ? p->Cout : 0;
  a.gn_cg = p->gn_group_ch > 0 ? p->gn_group_ch : (p->Cout >= 32 ? p->Cout / 32 : 1);
  a.gn_inv_cnt = 1.0 / ((double)a.Ho * a.Wo * (double)a.gn_cg);
  a.dbg = g_conv_dbg;
  const bool fuse_gn = p->gn_gamma != nullptr;
  if (fuse_gn) {
    // the in-kernel GroupNorm needs: statistics slot (+ barrier counter) already zeroed, one co-resident wave, no split-K
    if (!p->gn_stats || !p->gn_stats_zeroed || !p->gn_beta || nsplit > 1 || p->out_relu) return OTVM_ERR_UNSUPPORTED;
  }
  if (p->gn_stats && !p->gn_stats_zeroed && !dry_run) OTVM_CUDA_CHECK(cudaMemsetAsync(p->gn_stats, 0, sizeof(double) * 64, s));

  const CUtensorMapSwizzle swz = a.KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : a.KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tmA, tmB;
  {
    // (the plane of a split tensor is the outermost dimension; one plane: a dimension of extent 1)
    const uint64_t img_bytes = (uint64_t)p->N * p->H * p->W * p->in_ld * 2;
    uint64_t dims[5] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N, (uint64_t)a.planes};
    uint64_t str[4] = {(uint64_t)p->in_ld * 2, (uint64_t)p->W * p->in_ld * 2, (uint64_t)p->H * p->W * p->in_ld * 2,
                       a.planes > 1 ? (uint64_t)a.act_plane * 2 : img_bytes};
    // stride-2 convolutions: TMA traversal stride 2 along W and H (a box of 2*TW x 2*TH input pixels yields TW x TH)
    uint32_t box[5] = {(uint32_t)a.KC, (uint32_t)(a.TW * p->stride), (uint32_t)(a.TH * p->stride), 1, 1};
    if (a.halo) { box[1] = (uint32_t)(a.TW + 2 * p->dil); box[2] = (uint32_t)(a.TH + 2 * p->dil); }
    uint32_t es[5] = {1, (uint32_t)p->stride, (uint32_t)p->stride, 1, 1};
    int rc = make_tmap_bf16(&tmA, p->in, 5, dims, str, box, swz, es);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)p->KH * p->KW * p->Cin;
    const uint64_t wrows = (uint64_t)p->Cout * (uint64_t)(p->groups > 1 ? p->groups : 1);
    uint64_t dims[3] = {K, wrows, (uint64_t)a.planes};
    uint64_t str[2] = {K * 2, a.planes > 1 ? (uint64_t)p->w_plane_stride * 2 : K * 2 * wrows};
    uint32_t box[3] = {(uint32_t)a.KC, (uint32_t)bn, 1};
    int rc = make_tmap_bf16(&tmB, p->weight, 3, dims, str, box, swz);
    if (rc) return rc;
  }
  // epilogue through shared memory + TMA tensor store when the destination is a 16-byte aligned bf16 NHWC view
  CUtensorMap tmO = tmA, tmR = tmA;
  const int boxc = bn < 64 ? bn : 64;
  const bool tma_store = !(conv_tc_epi(p, bn) & EPI_DIRECT);
  if (tma_store) {
    const CUtensorMapSwizzle oswz = boxc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : boxc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    uint64_t dims[5] = {(uint64_t)p->Cout, (uint64_t)a.Wo, (uint64_t)a.Ho, (uint64_t)p->N, (uint64_t)a.planes};
    uint32_t box[5] = {(uint32_t)boxc, (uint32_t)a.TW, (uint32_t)a.TH, 1, 1};
    uint64_t str[4] = {(uint64_t)p->out_ps * 2, (uint64_t)a.Wo * p->out_ps * 2, (uint64_t)a.Ho * a.Wo * p->out_ps * 2,
                       a.planes > 1 ? (uint64_t)a.act_plane * 2 : (uint64_t)p->N * a.Ho * a.Wo * p->out_ps * 2};
    int rc = make_tmap_bf16(&tmO, p->out, 5, dims, str, box, oswz);
    if (rc) return rc;
    if (p->res && !p->out_relu) {            // the residual tile arrives by TMA (kResTma)
      uint64_t str2[4] = {(uint64_t)p->res_ld * 2, (uint64_t)a.Wo * p->res_ld * 2, (uint64_t)a.Ho * a.Wo * p->res_ld * 2,
                          a.planes > 1 ? (uint64_t)a.act_plane * 2 : (uint64_t)p->N * a.Ho * a.Wo * p->res_ld * 2};
      rc = make_tmap_bf16(&tmR, p->res, 5, dims, str2, box, oswz);
      if (rc) return rc;
    }
    if (p->out_relu) {
      uint64_t str2[4] = {(uint64_t)p->out_relu_ld * 2, (uint64_t)a.Wo * p->out_relu_ld * 2,
                          (uint64_t)a.Ho * a.Wo * p->out_relu_ld * 2,
                          a.planes > 1 ? (uint64_t)a.act_plane * 2 : (uint64_t)p->N * a.Ho * a.Wo * p->out_relu_ld * 2};
      rc = make_tmap_bf16(&tmR, p->out_relu, 5, dims, str2, box, oswz);
      if (rc) return rc;
    }
  }
  dim3 grid(a.tiles_x * a.tiles_y * p->N, ceil_div(p->Cout, bn), nsplit);
  if (nsplit > 1 && !clustered) {
    ConvTcArgs b = a;                       // raw fp32 partial tiles into the workspace, epilogue deferred
    b.bias = nullptr; b.out = p->workspace; b.out_ps = p->Cout; b.out_cs = 1; b.res = nullptr; b.out_relu = nullptr;
    b.act = OTVM_ACT_NONE; b.out_f32 = 1; b.gn_stats = nullptr;
    const size_t pipe_s = (size_t)a.b_off + (size_t)a.nstage * a.ksub * stage;
    const size_t smem_s = pipe_s + 1024 + 16 * 8 + 128 + 16 + 16 + (2 * 128 + 128) * sizeof(float) + (size_t)a.k_per_split * 16;
    b.aux_off = (uint32_t)pipe_s;
    int rc;
    switch (bn) {
      case 128: rc = dispatch_conv_tc<128>(GN_NONE, EPI_DIRECT, tmA, tmB, tmA, tmA, b, grid, smem_s, s); break;
      case 64: rc = dispatch_conv_tc<64>(GN_NONE, EPI_DIRECT, tmA, tmB, tmA, tmA, b, grid, smem_s, s); break;
      case 32: rc = dispatch_conv_tc<32>(GN_NONE, EPI_DIRECT, tmA, tmB, tmA, tmA, b, grid, smem_s, s); break;
      default: rc = dispatch_conv_tc<16>(GN_NONE, EPI_DIRECT, tmA, tmB, tmA, tmA, b, grid, smem_s, s); break;
    }
    if (rc) return rc;
    const int64_t total = Mtot * (p->Cout / 4);
    int g = (int)((total + 255) / 256); if (g > sm_count() * 8) g = sm_count() * 8;
    const int64_t ps = a.act_plane;
#define OTVM_FINISH(T)                                                                                              \
    launch_k(splitk_finish_kernel<T>, g, 256, 0, s, static_cast<const float*>(p->workspace), nsplit, Mtot, p->Cout,  \
             p->bias, mkcptr<T>(p->res, ps), p->res_ld, p->act, p->out, p->out_ps, p->out_cs, p->out_f32,            \
             mkptr<T>(p->out_relu, ps), p->out_relu_ld, p->gn_stats, ps, a.gn_cg)
    if (a.planes == 1) OTVM_FINISH(bf16); else if (a.planes == 2) OTVM_FINISH(bx<2>); else OTVM_FINISH(bx<3>);
#undef OTVM_FINISH
    OTVM_LAUNCH_CHECK();
    return OTVM_OK;
  }
  size_t pipe = (size_t)a.b_off + (size_t)a.nstage * ksub * stage;
  const size_t staging = (size_t)128 * bn * 2 * (p->out_relu ? 2 : 1) * a.planes;
  size_t need = tma_store ? staging : 0;
  if (p->gn_stats) need += (size_t)64 * 129 * sizeof(float);  // GroupNorm row partials (sred)
  if (clustered && need < (size_t)128 * bn * 4) need = (size_t)128 * bn * 4;   // fp32 partial tile of a rank >= 1 CTA
  if (need > pipe) pipe = need;           // the epilogue tile reuses the drained stages
  const size_t smem = pipe + 1024 + 16 * 8 + 128 + 16 + 16 + (2 * 128 + 128) * sizeof(float) + (size_t)num_k * 16;
  a.aux_off = (uint32_t)pipe;
  const int gn = fuse_gn ? GN_FUSED : p->gn_stats != nullptr ? GN_STATS : GN_NONE;
  const int epi = conv_tc_epi(p, bn);
  if (fuse_gn) {
    const int64_t per_sm = (int64_t)(227 * 1024) / (int64_t)(smem + 1024);
    const int64_t resident = (int64_t)sm_count() * (per_sm > 2 ? 2 : per_sm);
    if (ctas > resident || (epi & ~EPI_RES) != 0) return OTVM_ERR_UNSUPPORTED;
  } else if (dry_run) {
    return OTVM_ERR_UNSUPPORTED;
  }
  if (!fuse_gn && tma_store) {
    const int pr = try_conv_persist(p, a, bn, epi, gn, tmA, tmB, tmO, s);
    if (pr != 0) return pr > 0 ? OTVM_OK : pr;
  }
  switch (bn) {
    case 128: return dispatch_conv_tc<128>(gn, epi, tmA, tmB, tmO, tmR, a, grid, smem, s, dry_run);
    case 64: return dispatch_conv_tc<64>(gn, epi, tmA, tmB, tmO, tmR, a, grid, smem, s, dry_run);
    case 32: return dispatch_conv_tc<32>(gn, epi, tmA, tmB, tmO, tmR, a, grid, smem, s, dry_run);
    default: return dispatch_conv_tc<16>(gn, epi, tmA, tmB, tmO, tmR, a, grid, smem, s, dry_run);
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

int make_tmap(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz, const uint32_t* elem_strides) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return OTVM_ERR_UNSUPPORTED;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides) for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OTVM_OK : OTVM_ERR_ARG;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swz, const uint32_t* elem_strides) {
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swz, elem_strides);
}

}  // namespace otvm

extern "C" int otvm_device_error_flags(int clear) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, otvm::g_device_error_flags, sizeof(v)) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (clear && v) {
    const unsigned int z = 0;
    if (cudaMemcpyToSymbol(otvm::g_device_error_flags, &z, sizeof(z)) != cudaSuccess) cudaGetLastError();
  }
  return (int)v;
}

// dev hook (not part of the documented ABI): per-CTA clock64 timestamps of the next tcgen05 conv launches
extern "C" __attribute__((visibility("default"))) void otvm_debug_set_conv_timestamps(long long* buf) {
  otvm::g_conv_dbg = buf;
}
extern "C" __attribute__((visibility("default"))) void otvm_debug_set_conv_budget_kb(int kb) { otvm::g_conv_budget_kb = kb; }
// dev hook: K-chunks per barrier pair on one-wave grids (1 or 2; env OTVM_CONV_KSUB)
extern "C" __attribute__((visibility("default"))) void otvm_debug_set_conv_ksub(int n) { otvm::g_conv_ksub = n; }
// dev hook: persistent patch-mode kernel (-1 auto, 0 off, 1 whenever the shape allows; env OTVM_CONV_PERSIST)
extern "C" __attribute__((visibility("default"))) void otvm_debug_set_conv_persist(int mode) { otvm::g_conv_persist = mode; }
extern "C" __attribute__((visibility("default"))) long long otvm_debug_conv_persist_launches() { return otvm::g_conv_persist_launches; }
extern "C" __attribute__((visibility("default"))) long long otvm_debug_conv_cluster_launches() { return otvm::g_conv_cluster_launches; }
// dev hook: 3x3 halo-patch mode on/off (default on; env OTVM_CONV_HALO=0)
extern "C" __attribute__((visibility("default"))) void otvm_debug_set_conv_halo(int enabled) {
  otvm::g_conv_halo = enabled;              // -1 auto, 0 off, 1 forced on
}
