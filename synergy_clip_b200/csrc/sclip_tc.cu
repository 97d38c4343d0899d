// Tensor-core kernels of the tri-modal contrastive objective (tcgen05 / TMEM / TMA, sm_100a only).
//
//   forward_tiles   S = Xhat_rows . Xhat_cols^T per 128x256 tile, epilogue: E = exp(s*S - ref), row sums,
//                   column sums, diagonal -- the logits never leave TMEM          (model.py:254-265, 52-58)
//   backward_tiles  recompute S, epilogue: G' = kappa c_p ((softmax_rows + softmax_cols)/2 - I) as fp16 tiles
//                   + partial sums of dL/dlogit_scale                               (autograd of model.py:52-58)
//   gemm            dXhat = G' . Xhat_cols,  dXhat += G'^T . Xhat_rows              (autograd of model.py:255,260,265)
//
// One persistent CTA per SM (or one CTA pair per two SMs with cta_group::2), looping over tiles:
//   warps 0..EW-1  epilogue: EW = 8 or 16 warps, EW / 4 per TMEM lane quarter, each on its slice of the 256 columns
//   warp EW        TMA producer: ring of `stages` k blocks (A: 128 x 64, B: 256/CG x 64 fp16, 128-byte swizzle)
//   warp EW+1      single-thread tcgen05.mma issuer (leader CTA only when CG == 2) and TMEM owner
// (the two single-thread roles sit on the highest warp ids: the warp scheduler favours higher ids, and a delayed
//  TMA / MMA issue stalls the whole SM while a delayed epilogue warp does not)
// The 512 TMEM columns hold two 128x256 fp32 accumulators, so the epilogue of tile n runs under the MMAs of tile n+1.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace sclip {
using namespace ptx;

namespace {

constexpr int kMaxStages = 6;
constexpr int kAccStages = 2;
constexpr int kMaxEpiWarps = 16;

struct __align__(8) PipeBarriers {
  uint64_t full[kMaxStages];    // TMA -> MMA: operand bytes have landed (leader CTA's barrier collects both CTAs' bytes)
  uint64_t empty[kMaxStages];   // MMA -> TMA: the MMAs that read this slot have completed
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint32_t tmem_base;
  uint32_t pad;
};

template <int CG>
struct Geo {
  static constexpr int kBRows = BN / CG;                               // B rows loaded by one CTA per stage
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBBytes;               // 48 KiB (CG 1) / 32 KiB (CG 2)
  static constexpr int kLoadBytes = A_STAGE_BYTES + kBBytes;                // bytes the tensor loads deliver
  static constexpr int kTileM = BM * CG;                               // rows of one cluster tile
};

__device__ __forceinline__ uint8_t* aligned_dyn_smem() {
  extern __shared__ uint8_t dyn_smem_raw[];
  uint32_t a = smem_u32(dyn_smem_raw);
  uint32_t pad = (1024u - (a & 1023u)) & 1023u;
  return dyn_smem_raw + pad;
}

struct Tile {
  int job;    // job / pair index
  int m0;     // first row of the cluster tile
  int n0;     // first column
  int split;  // k split index
  int ti;     // cluster-row-tile index
  int tj;
};

struct RingState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int stages) {
    if (++stage == stages) {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// k blocks [lo, hi) of a segment handled by split `split` of `ksplits`
__device__ __forceinline__ void split_range(int num_kb, int split, int ksplits, int& lo, int& hi) {
  lo = static_cast<int>((static_cast<long long>(num_kb) * split) / ksplits);
  hi = static_cast<int>((static_cast<long long>(num_kb) * (split + 1)) / ksplits);
}

// ---------------------------------------------------------------------------------------------- mainloop roles
// Both roles are convergent warp code (see ptx.cuh, "warp-uniform role helpers"): every lane runs the loops and the
// barrier waits, `elected` is true in exactly one lane, and only the TMA / MMA / commit instructions sit under it.
template <int CG, bool A_MN, bool B_MN>
__device__ __forceinline__ void produce_kblocks(const CUtensorMap* ma, const CUtensorMap* mb, int m0, int n0, int kb_lo,
                                                int kb_hi, uint32_t smem0, uint32_t full_sig0, uint32_t full_loc0,
                                                PipeBarriers* bars, int stages, RingState& rs, bool elected, bool arm) {
  using G = Geo<CG>;
  for (int kb = kb_lo; kb < kb_hi; ++kb) {
    mbar_wait_bounded<false>(&bars->empty[rs.stage], rs.phase ^ 1u, 1);
    if (elected) {
      const uint32_t sa = smem0 + rs.stage * G::kStageBytes;
      const uint32_t sb = sa + A_STAGE_BYTES;
      const uint32_t bar = full_sig0 + rs.stage * 8;
      const int k = kb * BK;
      // CG 2: both CTAs of the pair signal the leader's barrier; the leader arms it for the bytes of both
      if (arm) mbar_expect_tx_u32(full_loc0 + rs.stage * 8, CG * G::kLoadBytes);
      if constexpr (CG == 1) {
        if constexpr (!A_MN) {
          tma_load_2d_u32(sa, ma, bar, k, m0);
        } else {
#pragma unroll
          for (int g = 0; g < BM / 64; ++g) tma_load_2d_u32(sa + g * MN_BOX_BYTES, ma, bar, m0 + g * 64, k);
        }
        if constexpr (!B_MN) {
          tma_load_2d_u32(sb, mb, bar, k, n0);
        } else {
#pragma unroll
          for (int g = 0; g < G::kBRows / 64; ++g) tma_load_2d_u32(sb + g * MN_BOX_BYTES, mb, bar, n0 + g * 64, k);
        }
      } else {
        if constexpr (!A_MN) {
          tma_load_2d_2sm_u32(sa, ma, bar, k, m0);
        } else {
#pragma unroll
          for (int g = 0; g < BM / 64; ++g) tma_load_2d_2sm_u32(sa + g * MN_BOX_BYTES, ma, bar, m0 + g * 64, k);
        }
        if constexpr (!B_MN) {
          tma_load_2d_2sm_u32(sb, mb, bar, k, n0);
        } else {
#pragma unroll
          for (int g = 0; g < G::kBRows / 64; ++g) tma_load_2d_2sm_u32(sb + g * MN_BOX_BYTES, mb, bar, n0 + g * 64, k);
        }
      }
    }
    rs.advance(stages);
  }
}

// Called by all lanes of the producer warp.
template <int CG>
__device__ __forceinline__ void producer_tile(const CUtensorMap* maps, const Job& job, const Tile& t, uint32_t rank,
                                              uint8_t* smem, PipeBarriers* bars, int stages, RingState& rs,
                                              bool elected) {
  using G = Geo<CG>;
  const int m0 = t.m0 + static_cast<int>(rank) * BM;         // this CTA's accumulator rows
  const int n0 = t.n0 + static_cast<int>(rank) * G::kBRows;  // this CTA's share of the B rows
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full_loc0 = smem_u32(&bars->full[0]);
  const uint32_t full_sig0 = CG == 1 ? full_loc0 : mapa(full_loc0, 0);
  const bool arm = rank == 0;
  for (int s = 0; s < job.nseg; ++s) {
    const CUtensorMap* ma = maps + job.seg[s].map_a;
    const CUtensorMap* mb = maps + job.seg[s].map_b;
    const int a_mn = job.seg[s].a_mn, b_mn = job.seg[s].b_mn;
    int kb_lo, kb_hi;
    split_range(job.seg[s].num_kb, t.split, job.ksplits, kb_lo, kb_hi);
    if (!a_mn && !b_mn)
      produce_kblocks<CG, false, false>(ma, mb, m0, n0, kb_lo, kb_hi, smem0, full_sig0, full_loc0, bars, stages, rs, elected, arm);
    else if (!a_mn)
      produce_kblocks<CG, false, true>(ma, mb, m0, n0, kb_lo, kb_hi, smem0, full_sig0, full_loc0, bars, stages, rs, elected, arm);
    else if (!b_mn)
      produce_kblocks<CG, true, false>(ma, mb, m0, n0, kb_lo, kb_hi, smem0, full_sig0, full_loc0, bars, stages, rs, elected, arm);
    else
      produce_kblocks<CG, true, true>(ma, mb, m0, n0, kb_lo, kb_hi, smem0, full_sig0, full_loc0, bars, stages, rs, elected, arm);
  }
}

// descriptor steps: K-major operand: 8-row groups 1024 B apart, k step of 16 elements = 32 B inside the 128-byte swizzle
// row (low word + 2).  MN-major operand: 64-element groups one box (8 KiB) apart, 8-k groups 1024 B apart, k step =
// 16 rows of 128 B = 2 KiB (low word + 128).
template <bool MN>
struct OperandDesc {
  static constexpr uint32_t kLbo = MN ? MN_BOX_BYTES : 16;
  static constexpr uint32_t kStep = MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
};

template <int CG, bool A_MN, bool B_MN>
__device__ __forceinline__ void mma_kblocks(int nkb, uint32_t smem0, uint32_t empty0, PipeBarriers* bars, int stages,
                                            RingState& rs, uint32_t tmem_acc, uint32_t& accumulate, bool elected) {
  using G = Geo<CG>;
  constexpr uint32_t idesc = make_idesc_f16(BM * CG, BN, /*fp16*/ 0, A_MN, B_MN);
  constexpr uint32_t hi = smem_desc_hi_sw128(1024);
  for (int kb = 0; kb < nkb; ++kb) {
    mbar_wait_bounded<false>(&bars->full[rs.stage], rs.phase, 2);
    tc_fence_after();
    if (elected) {
      const uint32_t a_lo = smem_desc_lo(smem0 + rs.stage * G::kStageBytes, OperandDesc<A_MN>::kLbo);
      const uint32_t b_lo = smem_desc_lo(smem0 + rs.stage * G::kStageBytes + A_STAGE_BYTES, OperandDesc<B_MN>::kLbo);
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        umma_f16<CG>(tmem_acc, smem_desc_join(a_lo + k * OperandDesc<A_MN>::kStep, hi),
                     smem_desc_join(b_lo + k * OperandDesc<B_MN>::kStep, hi), idesc, accumulate);
        accumulate = 1;
      }
      // frees the smem slot (in both CTAs) once these MMAs have read it
      if constexpr (CG == 1) umma_commit_1sm_u32(empty0 + rs.stage * 8);
      else umma_commit_2sm_u32(empty0 + rs.stage * 8, 0b11);
    }
    rs.advance(stages);
  }
}

// Called by all lanes of the MMA warp of the leader CTA.
template <int CG>
__device__ __forceinline__ void mma_tile(const Job& job, const Tile& t, uint8_t* smem, PipeBarriers* bars, int stages,
                                         RingState& rs, uint32_t tmem_acc, int acc, uint32_t acc_phase, bool elected) {
  // the epilogue must have drained this accumulator (two tiles ago)
  mbar_wait_bounded<false>(&bars->tmem_empty[acc], acc_phase ^ 1u, 4);
  tc_fence_after();
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t empty0 = smem_u32(&bars->empty[0]);
  uint32_t accumulate = 0;
  for (int s = 0; s < job.nseg; ++s) {
    const int a_mn = job.seg[s].a_mn, b_mn = job.seg[s].b_mn;
    int kb_lo, kb_hi;
    split_range(job.seg[s].num_kb, t.split, job.ksplits, kb_lo, kb_hi);
    const int nkb = kb_hi - kb_lo;
    if (!a_mn && !b_mn) mma_kblocks<CG, false, false>(nkb, smem0, empty0, bars, stages, rs, tmem_acc, accumulate, elected);
    else if (!a_mn) mma_kblocks<CG, false, true>(nkb, smem0, empty0, bars, stages, rs, tmem_acc, accumulate, elected);
    else if (!b_mn) mma_kblocks<CG, true, false>(nkb, smem0, empty0, bars, stages, rs, tmem_acc, accumulate, elected);
    else mma_kblocks<CG, true, true>(nkb, smem0, empty0, bars, stages, rs, tmem_acc, accumulate, elected);
  }
  if (elected) {
    if constexpr (CG == 1) umma_commit_1sm_u32(smem_u32(&bars->tmem_full[acc]));
    else umma_commit_2sm_u32(smem_u32(&bars->tmem_full[acc]), 0b11);
  }
}

// Common prologue: barrier init, TMEM allocation (all 512 columns = two accumulators).
template <int CG, int EW>
__device__ __forceinline__ uint32_t kernel_setup(PipeBarriers* bars, int warp, int lane) {
  if (warp == EW && lane == 0) {
#pragma unroll
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
#pragma unroll
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], CG * EW);
    }
    fence_mbar_init();
  }
  if (warp == EW + 1) {
    tmem_alloc<CG>(&bars->tmem_base, kAccStages * BN);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if constexpr (CG == 1) __syncthreads();
  else cluster_sync();  // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);
}

template <int CG, int EW>
__device__ __forceinline__ void kernel_teardown(uint32_t tmem_base, int warp) {
  tc_fence_before();
  if constexpr (CG == 1) __syncthreads();
  else cluster_sync();  // no CTA of the pair may exit while the other can still signal its barriers
  if (warp == EW + 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, kAccStages * BN);
  }
}

// the epilogue warp has finished reading accumulator `acc`: hand it back to the MMA issuer (leader CTA)
template <int CG>
__device__ __forceinline__ void release_accumulator(PipeBarriers* bars, int acc, uint32_t rank, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (CG == 1 || rank == 0) mbar_arrive(&bars->tmem_empty[acc]);
    else mbar_arrive_cluster(&bars->tmem_empty[acc], 0);
  }
}

// grouped rasterisation: consecutive tiles walk down 8 row tiles before moving to the next column tile, so the
// CTAs running at the same time share a compact set of operand rows in L2
__device__ __forceinline__ void decode_grouped(int id, int nti, int ntj, int& ti, int& tj) {
  constexpr int GM = 8;
  const int per_group = GM * ntj;
  const int group = id / per_group;
  const int first = group * GM;
  const int gm = min(nti - first, GM);
  const int r = id - group * per_group;
  ti = first + r % gm;
  tj = r / gm;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, w));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) v += __shfl_xor_sync(0xffffffffu, v, w);
  return v;
}

// ---- fragment algebra of tmem_ld_block32 (ptx.cuh): element i of thread `lane` is
//        row 16 (i >> 4) + 8 ((i >> 1) & 1) + lane / 4,   column 8 ((i >> 2) & 3) + 2 (lane % 4) + (i & 1)
//      i.e. every thread owns 4 rows x 8 columns of the 32 x 32 block.
__device__ __forceinline__ int frag_row(int i, int lane) { return 16 * (i >> 4) + 8 * ((i >> 1) & 1) + (lane >> 2); }
__device__ __forceinline__ int frag_col(int i, int lane) { return 8 * ((i >> 2) & 3) + 2 * (lane & 3) + (i & 1); }

// Column sums of a 32 x 32 block: 24 in-thread adds (4 rows per column), then a reduce-scatter of the 8 column partials
// over the 8 threads that share lane % 4 (7 shuffles).  Returns the sum over the 32 rows of block column `col`.
__device__ __forceinline__ float frag_column_sum(const float (&e)[32], int lane, int& col) {
  float cp[8];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int c = 0; c < 2; ++c)
      cp[2 * n + c] = (e[4 * n + c] + e[4 * n + 2 + c]) + (e[16 + 4 * n + c] + e[16 + 4 * n + 2 + c]);
#pragma unroll
  for (int w = 16, half = 4; half >= 1; w >>= 1, half >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const float send = up ? cp[k] : cp[k + half];
      const float keep = up ? cp[k + half] : cp[k];
      cp[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  col = 8 * (idx >> 1) + 2 * (lane & 3) + (idx & 1);
  return cp[0];
}

// In-thread row partials of a block: rp[2 g + h] += sum over the thread's 8 columns of row 16 g + 8 h + lane / 4.
__device__ __forceinline__ void frag_row_partials(const float (&e)[32], float (&rp)[4]) {
#pragma unroll
  for (int g = 0; g < 2; ++g)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = 16 * g + 2 * h;
      rp[2 * g + h] += ((e[o] + e[o + 1]) + (e[o + 4] + e[o + 5])) + ((e[o + 8] + e[o + 9]) + (e[o + 12] + e[o + 13]));
    }
}

// Reduce-scatter of the 4 row partials over the 4 threads of a quad (3 shuffles): returns the complete sum of
// block row `row`.
__device__ __forceinline__ float frag_row_sum(float (&rp)[4], int lane, int& row) {
  {
    const bool up = (lane & 1) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float send = up ? rp[k] : rp[k + 2];
      const float keep = up ? rp[k + 2] : rp[k];
      rp[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
  }
  {
    const bool up = (lane & 2) != 0;
    const float send = up ? rp[0] : rp[1];
    const float keep = up ? rp[1] : rp[0];
    rp[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  row = 16 * (lane & 1) + 8 * ((lane >> 1) & 1) + (lane >> 2);
  return rp[0];
}

constexpr uint32_t kBarAll = 1;     // named barrier of all epilogue threads
constexpr uint32_t kBarSlice = 2;   // + slice: the 128 threads (4 warps, one per lane quarter) of one column slice

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// saturating variant: values above 65504 (a stash element more than ~14 nats above both positive pairs) clamp to 65504;
// backward_scale_kernel recognises that value and hands the step over to the recompute kernel
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

constexpr int kSlabBytes = BM * 64 * 2;  // one 128-row x 64-column fp16 slab (16 KiB) staged for a TMA store

// ============================================================================================== forward tiles
// Pairs whose scale allows the folded-exponent epilogue of forward_fast_kernel: |L| <= s < 44 keeps
// 2^(c acc - h_i), h_i = (L_ii / 2) log2(e) + 2, and its row / column sums inside the normal fp32 range.
constexpr float kFoldMaxScale = 44.0f;

// pairs handled by a launch: those of P.pair_list with (fast ? s_p < kFoldMaxScale : s_p >= kFoldMaxScale)
__device__ __forceinline__ int select_pairs(const FwdParams& P, bool fast, int (&list)[3]) {
  int n = 0;
  for (int u = 0; u < P.npairs; ++u) {
    const int p = P.pair_list[u];
    const bool is_fast = expf(P.t3[p]) < kFoldMaxScale;
    if (is_fast == fast) list[n++] = p;
  }
  return n;
}

// ntj_wrap > 0: the column tiles [tj_begin, tj_begin + ntj) are taken modulo ntj_wrap (a range of ranks' columns that
// wraps around the end of the global batch)
template <int CG>
__device__ __forceinline__ Tile decode_similarity(int t, int nti_c, int ntj, int tj_begin = 0, int ntj_wrap = 0) {
  Tile r;
  const int per_pair = nti_c * ntj;
  r.job = t / per_pair;
  decode_grouped(t - r.job * per_pair, nti_c, ntj, r.ti, r.tj);
  r.tj += tj_begin;
  if (ntj_wrap > 0 && r.tj >= ntj_wrap) r.tj -= ntj_wrap;
  r.m0 = r.ti * Geo<CG>::kTileM;
  r.n0 = r.tj * BN;
  r.split = 0;
  return r;
}

// SCLIP_FWD_WAIT_PEERS: the tiles are taken wave by wave -- every (pair, row tile) on this rank's own columns first,
// then on the columns of rank - 1, rank - 2, ...: every rank pushes to rank + 1 first (sclip_push_shards), so the first
// shard to land here is that of rank - 1, the second that of rank - 2, and so on.
template <int CG>
__device__ __forceinline__ Tile decode_forward(const FwdParams& P, int t, int npairs, int nti_c, int& wave) {
  if (!P.wait_peers) {
    wave = 0;
    return decode_similarity<CG>(t, nti_c, P.tj_count, P.tj_begin, P.ntj);
  }
  const int per_wave = npairs * nti_c * P.tiles_per_rank;
  wave = t / per_wave;
  const int src = (P.world - wave) % P.world;  // ranks behind this one, modulo world
  return decode_similarity<CG>(t - wave * per_wave, nti_c, P.tiles_per_rank, P.tj_begin + src * P.tiles_per_rank, P.ntj);
}

// Block until the shard of wave `wave` (source rank (rank - wave) mod world) is complete in this workspace.
__device__ __forceinline__ void wait_landed(const FwdParams& P, int wave) {
  if (wave <= 0) return;
  const int* flag = P.landed + (P.rank + P.world - wave) % P.world;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");  // written by the peer GPU
    if (v - P.epoch >= 0) return;
    __nanosleep(200);
    if ((++spins & 0xFFF) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {
        printf("sclip: peer shard of wave %d never landed (block %d)\n", wave, static_cast<int>(blockIdx.x));
        __trap();
      }
    }
  }
}

template <int CG, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1) forward_tiles_kernel(const __grid_constant__ FwdParams P) {
  constexpr int S = EW / 4;        // column slices
  constexpr int CS = BN / S;       // columns per slice
  constexpr int NCH = CS / 32;     // 32-column chunks per warp
  constexpr uint32_t kEpiThreads = 32 * EW;
  __shared__ PipeBarriers bars;
  constexpr int NSLAB = CS / 64;   // 64-column slabs per slice (stash stores)
  __shared__ float colacc[kAccStages][4][BN];
  __shared__ float rowacc[kAccStages][S][BM];
  __shared__ float redw[kAccStages][kMaxEpiWarps];
  __shared__ __align__(8) float colh[kAccStages][BN];
  uint8_t* smem = aligned_dyn_smem();
  uint8_t* staging = smem + P.stages * Geo<CG>::kStageBytes;  // stash mode: S * NSLAB = 4 slabs
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  const int nti_c = P.nti / CG;
  // pair_filter: the pairs with s < kFoldMaxScale were taken by forward_fast_kernel
  int pairs[3] = {P.pair_list[0], P.pair_list[1], P.pair_list[2]};
  const int npairs = P.pair_filter ? select_pairs(P, false, pairs) : P.npairs;
  const int total = npairs * nti_c * P.tj_count;
  if (total == 0) return;  // every pair was taken by forward_fast_kernel

  const uint32_t tmem_base = kernel_setup<CG, EW>(&bars, warp, lane);

  if (warp == EW) {
    const bool elected = elect_one();
    RingState rs;
    for (int t = cluster_id; t < total; t += num_clusters) {
      Tile tile = decode_similarity<CG>(t, nti_c, P.tj_count, P.tj_begin, P.ntj);
      tile.job = pairs[tile.job];
      producer_tile<CG>(P.maps, P.jobs[tile.job], tile, rank, smem, &bars, P.stages, rs, elected);
    }
  } else if (warp == EW + 1) {
    if (rank == 0) {  // the leader CTA issues the MMAs of the pair
      const bool elected = elect_one();
      RingState rs;
      int it = 0;
      for (int t = cluster_id; t < total; t += num_clusters, ++it) {
        Tile tile = decode_similarity<CG>(t, nti_c, P.tj_count, P.tj_begin, P.ntj);
        tile.job = pairs[tile.job];
        const int acc = it & 1;
        mma_tile<CG>(P.jobs[tile.job], tile, smem, &bars, P.stages, rs, tmem_base + acc * BN, acc, (it >> 1) & 1, elected);
      }
    }
  } else {
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int slice = warp >> 2;   // which CS accumulator columns
    const int epi_tid = warp * 32 + lane;
    const int col0 = slice * CS;
    // positive-pair logits of the rows / columns a tile touches (stash scaling), loaded one tile ahead
    constexpr int kDiagCols = (BN + static_cast<int>(kEpiThreads) - 1) / static_cast<int>(kEpiThreads);
    float pre_r[4] = {0.f, 0.f, 0.f, 0.f};
    float pre_c[kDiagCols] = {};
    bool pre_valid = false;
    auto load_diag = [&](int pp, int wr0, int nn0, float (&r4)[4], float (&c1)[kDiagCols]) {
      const float* dg = P.diag_all + static_cast<size_t>(pp) * P.rows_global;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = wr0 + 16 * (j >> 1) + 8 * (j & 1) + (lane >> 2);
        r4[j] = row < P.rows_local ? dg[P.row_offset + row] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kDiagCols; ++u) {
        const int cc = epi_tid + u * kEpiThreads;
        c1[u] = (cc < BN && nn0 + cc < P.rows_global) ? dg[nn0 + cc] : 0.f;
      }
    };
    int it = 0;
    for (int t = cluster_id; t < total; t += num_clusters, ++it) {
      Tile tile = decode_similarity<CG>(t, nti_c, P.tj_count, P.tj_begin, P.ntj);
      tile.job = pairs[tile.job];
      const int acc = it & 1;
      const int p = tile.job;
      const int ti = tile.ti * CG + static_cast<int>(rank);  // 128-row tile index of this CTA
      const int tj = tile.tj;
      const int m0 = ti * BM, n0 = tile.n0;
      const float s = expf(P.t3[p]);
      const float c = s * kLog2e * P.acc_scale;  // accumulator -> logit in log2 units
      const int wrow0 = m0 + q * 32;             // first local row of this warp's block
      const bool edge = (m0 + BM > P.rows_local) || (n0 + BN > P.rows_global);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col0;

      // stash scaling: E~_ij = 2^(acc c - h_i - h_j), h = (L_ii / 2) log2(e) + 2.  While s < 64 the exponential E of
      // the statistics is reused: E~ = E sigma_i tau_j with sigma = 2^-h_i, tau = 2^-h_j (two multiplies instead of a
      // second exponential); above, E is relative to the tile maximum and E~ gets its own exponential.
      // The positive-pair logits were prefetched during the previous tile (pre_r / pre_c).
      const bool fast_tile = s < kFastPathMaxScale;
      float hr[4] = {0.f, 0.f, 0.f, 0.f};
      if (P.stash) {
        if (!pre_valid) load_diag(p, wrow0, n0, pre_r, pre_c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float h = fmaf(0.5f * kLog2e, pre_r[j], 2.0f);
          hr[j] = fast_tile ? ex2_approx(-h) : h;
        }
#pragma unroll
        for (int u = 0; u < kDiagCols; ++u) {
          const float h = fmaf(0.5f * kLog2e, pre_c[u], 2.0f);
          if (epi_tid + u * kEpiThreads < BN) colh[acc][epi_tid + u * kEpiThreads] = fast_tile ? ex2_approx(-h) : h;
        }
        // every staging slab was handed to a TMA store one tile ago: make sure those stores have read it
        if ((epi_tid & 127) == 0) tma_store_wait_read<0>();
        named_bar_sync(kBarAll, kEpiThreads);
      }

      mbar_wait_bounded<false>(&bars.tmem_full[acc], (it >> 1) & 1, 3);
      tc_fence_after();

      // exponent reference of this tile, in log2 units.  s < 64: |logit| <= s (cosines), so exp(logit) and its sums
      // are normal fp32 numbers without any shift; otherwise the true maximum of the tile is taken in a first pass.
      float ref2 = 0.f;
      if (!(s < kFastPathMaxScale)) {
        float mx = -INFINITY;
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t v[32];
          tmem_ld_block32(taddr + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = !edge || ((wrow0 + frag_row(i, lane)) < P.rows_local &&
                                      (n0 + col0 + ch * 32 + frag_col(i, lane)) < P.rows_global);
            mx = fmaxf(mx, ok ? __uint_as_float(v[i]) : -INFINITY);
          }
        }
        mx = warp_max(mx);
        if (lane == 0) redw[acc][warp] = mx;
        named_bar_sync(kBarAll, kEpiThreads);
#pragma unroll
        for (int w = 0; w < EW; ++w) mx = fmaxf(mx, redw[acc][w]);
        ref2 = mx * c;
      }

      float rp[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t va[32];
      [[maybe_unused]] uint32_t vb[32];
      auto process = [&](uint32_t (&v)[32], int ch, uint8_t* stash_slab = nullptr, int hc = 0) {
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(v[i]), c, -ref2));
        const int gcol0 = n0 + col0 + ch * 32;          // first global column of the block
        if (edge) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!((wrow0 + frag_row(i, lane)) < P.rows_local && (gcol0 + frag_col(i, lane)) < P.rows_global)) e[i] = 0.f;
        }
        if (stash_slab != nullptr) {
          float cf[8];
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float2 f = *reinterpret_cast<const float2*>(&colh[acc][col0 + ch * 32 + 8 * n + 2 * (lane & 3)]);
            cf[2 * n] = f.x;
            cf[2 * n + 1] = f.y;
          }
          // the four threads of a quad fill one 16-byte chunk of the swizzled staging slab
#pragma unroll
          for (int gh = 0; gh < 4; ++gh) {
            const int r_in_tile = q * 32 + 16 * (gh >> 1) + 8 * (gh & 1) + (lane >> 2);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              const int i0 = 16 * (gh >> 1) + 4 * n + 2 * (gh & 1);
              float s0, s1;
              if (fast_tile) {
                s0 = e[i0] * (hr[gh] * cf[2 * n]);
                s1 = e[i0 + 1] * (hr[gh] * cf[2 * n + 1]);
              } else {
                s0 = ex2_approx(fmaf(__uint_as_float(v[i0]), c, -(hr[gh] + cf[2 * n])));
                s1 = ex2_approx(fmaf(__uint_as_float(v[i0 + 1]), c, -(hr[gh] + cf[2 * n + 1])));
              }
              const int chunk16 = hc * 4 + n;
              const uint32_t off = r_in_tile * 128 + ((chunk16 ^ (r_in_tile & 7)) << 4) + 4 * (lane & 3);
              *reinterpret_cast<uint32_t*>(stash_slab + off) = pack_half2_sat(s0, s1);
            }
          }
        }
        const int grow0 = P.row_offset + wrow0;         // global index of the block's first row
        if (grow0 < gcol0 + 32 && gcol0 < grow0 + 32) {  // the block touches the diagonal: positive-pair logits
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (grow0 + frag_row(i, lane) == gcol0 + frag_col(i, lane) && (wrow0 + frag_row(i, lane)) < P.rows_local)
              P.diag[static_cast<size_t>(p) * P.rows_local + wrow0 + frag_row(i, lane)] =
                  __uint_as_float(v[i]) * s * P.acc_scale;
        }
        frag_row_partials(e, rp);
        int col;
        const float csum = frag_column_sum(e, lane, col);
        colacc[acc][q][col0 + ch * 32 + col] = csum;
      };
      if (P.stash) {
        const int slice_tid = epi_tid & 127;
        const CUtensorMap* map = &P.maps[P.store_map[p]];
#pragma unroll
        for (int sl = 0; sl < NSLAB; ++sl) {
          uint8_t* slab = staging + (slice * NSLAB + sl) * kSlabBytes;  // one buffer per slab, reused one tile later
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            const int ch = sl * 2 + hc;
            tmem_ld_block32(taddr + ch * 32, va);
            tmem_ld_wait();
            process(va, ch, slab, hc);
          }
          fence_proxy_async_smem();
          named_bar_sync(kBarSlice + slice, 128);
          const int gcol = n0 + col0 + sl * 64;
          if (slice_tid == 0) {
            if (gcol < P.rows_global) tma_store_2d(map, slab, gcol, m0);
            tma_store_commit();
          }
        }
      } else if constexpr (EW == 8) {
        // software pipeline: the TMEM load of chunk ch + 1 is in flight while chunk ch is processed
        tmem_ld_block32(taddr, va);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_wait();
          if (ch & 1) {
            if (ch + 1 < NCH) tmem_ld_block32(taddr + (ch + 1) * 32, va);
            process(vb, ch);
          } else {
            if (ch + 1 < NCH) tmem_ld_block32(taddr + (ch + 1) * 32, vb);
            process(va, ch);
          }
        }
      } else {  // 16 warps: four warps per scheduler hide the load latency, registers are the scarcer resource
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_block32(taddr + ch * 32, va);
          tmem_ld_wait();
          process(va, ch);
        }
      }
      release_accumulator<CG>(&bars, acc, rank, lane);  // all TMEM reads of this warp are complete
      pre_valid = false;
      if (P.stash && t + num_clusters < total) {  // start the next tile's diagonal loads under this tile's tail
        Tile nx = decode_similarity<CG>(t + num_clusters, nti_c, P.tj_count, P.tj_begin, P.ntj);
        const int np = pairs[nx.job];
        load_diag(np, (nx.ti * CG + static_cast<int>(rank)) * BM + q * 32, nx.n0, pre_r, pre_c);
        pre_valid = true;
      }
      {
        int r;
        const float rsum = frag_row_sum(rp, lane, r);
        rowacc[acc][slice][q * 32 + r] = rsum;
      }
      named_bar_sync(kBarAll, kEpiThreads);
      for (int cc = epi_tid; cc < BN; cc += kEpiThreads) {
        if (n0 + cc < P.rows_global)
          P.col_part[(static_cast<size_t>(p) * P.nti + ti) * P.rows_global + n0 + cc] =
              (colacc[acc][0][cc] + colacc[acc][1][cc]) + (colacc[acc][2][cc] + colacc[acc][3][cc]);
      }
      if (epi_tid < BM && m0 + epi_tid < P.rows_local) {
        float rs = 0.f;
#pragma unroll
        for (int u = 0; u < S; ++u) rs += rowacc[acc][u][epi_tid];
        // two slots per column tile (forward_fast_kernel writes one per column slice)
        P.row_part[(static_cast<size_t>(p) * P.ntj * 2 + tj * 2) * P.rows_local + m0 + epi_tid] = rs;
        P.row_part[(static_cast<size_t>(p) * P.ntj * 2 + tj * 2 + 1) * P.rows_local + m0 + epi_tid] = 0.f;
      }
      if (epi_tid == 0) P.tile_ref[(static_cast<size_t>(p) * P.nti + ti) * P.ntj + tj] = ref2 / kLog2e;
    }
    if (P.stash && (epi_tid & 127) == 0) tma_store_wait_all<0>();
  }
  kernel_teardown<CG, EW>(tmem_base, warp);
}

// ============================================================================================== fast forward tiles
// The forward tile kernel for the common case s = exp(logit_scale) < 64 (no exponent reference), CTA pairs, 8 epilogue
// warps.  Same mainloop; the epilogue is rebuilt around what bounded the generic one (17 instructions per logit,
// two all-warp barriers per tile):
//   * packed fp32 pairs (FFMA2 / FADD2 / FMUL2) on the two adjacent columns every thread owns,
//   * the row factor of the stash folded into the exponent: e' = 2^(c acc - h_i) = E sigma_i feeds the row sums
//     (scaled back once per row), the column sums (FFMA2 with weight 1 / sigma_i) and the stash E~ = e' tau_j,
//   * stmatrix for the fp16 staging (one instruction per 8 x 32 block instead of four stores),
//   * no all-warp barrier: column statistics are combined by the four warps that share a column slice, the two
//     column slices of a row write separate row partials, the column factors of tile n + 1 are staged during tile n,
//     and a staging slab is reused two slabs later (the storing thread drains its bulk reads before the barrier).
// Pairs with s >= kFoldMaxScale are left to forward_tiles_kernel (pair_filter).
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void stmatrix_x4(uint32_t smem_addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(r0), "r"(r1),
               "r"(r2), "r"(r3)
               : "memory");
}

template <bool STASH>
__global__ void __launch_bounds__(64 + 32 * 8, 1) forward_fast_kernel(const __grid_constant__ FwdParams P) {
  constexpr int CG = 2, EW = 8;
  constexpr int CS = BN / 2;       // columns per slice (two slices of four warps)
  constexpr int NCH = CS / 32;     // 32-column chunks per warp and tile
  __shared__ PipeBarriers bars;
  __shared__ float colacc[kAccStages][4][BN];
  __shared__ __align__(8) float coltau[kAccStages][BN];
  uint8_t* smem = aligned_dyn_smem();
  uint8_t* staging = smem + P.stages * Geo<CG>::kStageBytes;  // STASH: 2 slices x 2 slabs of 16 KiB
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  const int nti_c = P.nti / CG;
  int pairs[3];
  const int npairs = select_pairs(P, true, pairs);
  const int total = npairs * nti_c * P.tj_count;
  if (total == 0 && !P.wait_peers) return;  // nothing for this kernel (every pair has s >= kFoldMaxScale)

  const uint32_t tmem_base = kernel_setup<CG, EW>(&bars, warp, lane);

  if (warp == EW) {
    const bool elected = elect_one();
    RingState rs;
    int landed_wave = 0;
    for (int t = cluster_id; t < total; t += num_clusters) {
      int wave;
      Tile tile = decode_forward<CG>(P, t, npairs, nti_c, wave);
      tile.job = pairs[tile.job];
      for (; landed_wave < wave; ++landed_wave) wait_landed(P, landed_wave + 1);
      producer_tile<CG>(P.maps, P.jobs[tile.job], tile, rank, smem, &bars, P.stages, rs, elected);
    }
    // a kernel launched behind this one (the s >= 44 pairs, the backward) may read every shard without checking
    if (P.wait_peers)
      for (; landed_wave < P.world - 1; ++landed_wave) wait_landed(P, landed_wave + 1);
  } else if (warp == EW + 1) {
    if (rank == 0) {  // the leader CTA issues the MMAs of the pair
      const bool elected = elect_one();
      RingState rs;
      int it = 0;
      for (int t = cluster_id; t < total; t += num_clusters, ++it) {
        int wave;
        Tile tile = decode_forward<CG>(P, t, npairs, nti_c, wave);
        tile.job = pairs[tile.job];
        const int acc = it & 1;
        mma_tile<CG>(P.jobs[tile.job], tile, smem, &bars, P.stages, rs, tmem_base + acc * BN, acc, (it >> 1) & 1, elected);
      }
    }
  } else {
    const int q = warp & 3;        // TMEM lane quarter
    const int slice = warp >> 2;   // column half of the tile
    const int slice_tid = (warp & 3) * 32 + lane;
    const int col0 = slice * CS;
    const int my_col = col0 + slice_tid;  // the tile column whose factor / statistics this thread stages
    // stmatrix row addresses: lane L supplies row (L & 7) of 8 x 8 matrix (L >> 3) = 16-byte chunk (L >> 3) of the
    // 64-byte block; rows 128 bytes apart, chunk index XOR row & 7 (the TMA 128-byte swizzle)
    const uint32_t st_row = static_cast<uint32_t>(q * 32 + (lane & 7)) * 128u;
    const uint32_t st_chunk0 = static_cast<uint32_t>(((lane >> 3) ^ (lane & 7)) << 4);
    const uint32_t st_chunk1 = static_cast<uint32_t>((((lane >> 3) + 4) ^ (lane & 7)) << 4);

    // positive-pair logits of a tile's rows / columns -> stash factors (loaded one tile ahead)
    int landed_wave = 0;  // waves whose positive-pair logits may be read (WAIT_PEERS)
    auto tile_coords = [&](int t, int& p, int& ti, int& tj, int& m0, int& n0) {
      int wave;
      Tile tile = decode_forward<CG>(P, t, npairs, nti_c, wave);
      for (; landed_wave < wave; ++landed_wave) wait_landed(P, landed_wave + 1);
      p = pairs[tile.job];
      ti = tile.ti * CG + static_cast<int>(rank);
      tj = tile.tj;
      m0 = ti * BM;
      n0 = tile.n0;
    };
    auto load_diag = [&](int t, float (&r4)[4], float& c1) {
      int p, ti, tj, m0, n0;
      tile_coords(t, p, ti, tj, m0, n0);
      const float* dg = P.diag_all + static_cast<size_t>(p) * P.rows_global;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = m0 + q * 32 + 16 * (j >> 1) + 8 * (j & 1) + (lane >> 2);
        r4[j] = row < P.rows_local ? __ldcg(dg + P.row_offset + row) : 0.f;
      }
      c1 = (n0 + my_col < P.rows_global) ? __ldcg(dg + n0 + my_col) : 0.f;
    };
    float nxt_r[4] = {0.f, 0.f, 0.f, 0.f}, nxt_c = 0.f;  // raw positive-pair logits of the next tile
    float cur_r[4] = {0.f, 0.f, 0.f, 0.f};
    if (STASH && cluster_id < total) {
      float c1;
      load_diag(cluster_id, cur_r, c1);
      coltau[0][my_col] = ex2_approx(-fmaf(0.5f * kLog2e, c1, 2.0f));
    }
    named_bar_sync(kBarSlice + slice, 128);

    int it = 0;
    for (int t = cluster_id; t < total; t += num_clusters, ++it) {
      int p, ti, tj, m0, n0;
      tile_coords(t, p, ti, tj, m0, n0);
      const int acc = it & 1;
      const float s = expf(P.t3[p]);
      const float c = s * kLog2e * P.acc_scale;  // accumulator -> logit in log2 units
      const int wrow0 = m0 + q * 32;
      const bool edge = (m0 + BM > P.rows_local) || (n0 + BN > P.rows_global);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col0;
      const bool has_next = t + num_clusters < total;
      if (STASH && has_next) load_diag(t + num_clusters, nxt_r, nxt_c);

      // h_i = (L_ii / 2) log2(e) + 2: e' = 2^(c acc - h_i) = E sigma_i;  isig = 1 / sigma_i restores E in the sums
      uint64_t nh2[4], isig2[4];
      float isig[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float h = STASH ? fmaf(0.5f * kLog2e, cur_r[j], 2.0f) : 0.f;
        nh2[j] = pack2(-h, -h);
        isig[j] = STASH ? ex2_approx(h) : 1.0f;
        isig2[j] = pack2(isig[j], isig[j]);
      }
      const uint64_t c2 = pack2(c, c);

      mbar_wait_bounded<false>(&bars.tmem_full[acc], (it >> 1) & 1, 3);
      tc_fence_after();
      uint64_t rp2[4] = {0ull, 0ull, 0ull, 0ull};  // row partials (pairs of adjacent columns), in e' units
      uint32_t va[32], vb[32];
      auto process = [&](uint32_t (&v)[32], int ch) {
        const int gcol0 = n0 + col0 + ch * 32;  // first global column of the block
        const int grow0 = P.row_offset + wrow0;
        if (grow0 < gcol0 + 32 && gcol0 < grow0 + 32) {  // the block touches the diagonal: positive-pair logits
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (grow0 + frag_row(i, lane) == gcol0 + frag_col(i, lane) && (wrow0 + frag_row(i, lane)) < P.rows_local)
              P.diag[static_cast<size_t>(p) * P.rows_local + wrow0 + frag_row(i, lane)] =
                  __uint_as_float(v[i]) * s * P.acc_scale;
        }
        uint64_t e2[16];  // e2[8 g + 2 n + h] = columns (8 n + 2 (lane % 4), + 1) of row 16 g + 8 h + lane / 4
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int i = 16 * g + 4 * n + 2 * h;
              const uint64_t x = ffma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nh2[2 * g + h]);
              float x0, x1;
              unpack2(x, x0, x1);
              e2[8 * g + 2 * n + h] = pack2(ex2_approx(x0), ex2_approx(x1));
            }
        if (edge) {  // ragged last row / column tile: entries outside the matrix contribute nothing
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = 16 * g + 4 * n + 2 * h;
                float e0, e1;
                unpack2(e2[8 * g + 2 * n + h], e0, e1);
                const bool rok = (wrow0 + frag_row(i, lane)) < P.rows_local;
                if (!(rok && (gcol0 + frag_col(i, lane)) < P.rows_global)) e0 = 0.f;
                if (!(rok && (gcol0 + frag_col(i + 1, lane)) < P.rows_global)) e1 = 0.f;
                e2[8 * g + 2 * n + h] = pack2(e0, e1);
              }
        }
        // row partials and column partials (weights 1 / sigma_i turn e' back into E)
        uint64_t cp2[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          cp2[n] = fmul2(e2[2 * n], isig2[0]);
          cp2[n] = ffma2(e2[2 * n + 1], isig2[1], cp2[n]);
          cp2[n] = ffma2(e2[8 + 2 * n], isig2[2], cp2[n]);
          cp2[n] = ffma2(e2[8 + 2 * n + 1], isig2[3], cp2[n]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t a = fadd2(e2[8 * g + h], e2[8 * g + 2 + h]);
            const uint64_t b = fadd2(e2[8 * g + 4 + h], e2[8 * g + 6 + h]);
            rp2[2 * g + h] = fadd2(rp2[2 * g + h], fadd2(a, b));
          }
        if constexpr (STASH) {
          // E~ = e' tau_j, packed to fp16 and staged through stmatrix in the TMA store's swizzled layout
          uint8_t* slab = staging + (slice * 2 + (ch >> 1)) * kSlabBytes;
          const uint32_t slab_addr = smem_u32(slab) + st_row + ((ch & 1) ? st_chunk1 : st_chunk0);
          uint64_t tau2[4];
#pragma unroll
          for (int n = 0; n < 4; ++n)
            tau2[n] = *reinterpret_cast<const uint64_t*>(&coltau[acc][col0 + ch * 32 + 8 * n + 2 * (lane & 3)]);
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t r[4];
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                float s0, s1;
                unpack2(fmul2(e2[8 * g + 2 * n + h], tau2[n]), s0, s1);
                r[n] = pack_half2_sat(s0, s1);
              }
              stmatrix_x4(slab_addr + static_cast<uint32_t>(16 * g + 8 * h) * 128u, r[0], r[1], r[2], r[3]);
            }
        }
        // column sums of the 32 rows: reduce-scatter of the 8 partials over the 8 lanes that share lane % 4
        float cp[8];
#pragma unroll
        for (int n = 0; n < 4; ++n) unpack2(cp2[n], cp[2 * n], cp[2 * n + 1]);
#pragma unroll
        for (int w = 16, half = 4; half >= 1; w >>= 1, half >>= 1) {
          const bool up = (lane & w) != 0;
#pragma unroll
          for (int k = 0; k < half; ++k) {
            const float send = up ? cp[k] : cp[k + half];
            const float keep = up ? cp[k + half] : cp[k];
            cp[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
          }
        }
        const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        colacc[acc][q][col0 + ch * 32 + 8 * (idx >> 1) + 2 * (lane & 3) + (idx & 1)] = cp[0];
      };

      // software pipeline over the four chunks: the TMEM load of chunk ch + 1 is in flight while chunk ch is processed
      tmem_ld_block32(taddr, va);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        tmem_ld_wait();
        if (ch & 1) {
          if (ch + 1 < NCH) tmem_ld_block32(taddr + (ch + 1) * 32, va);
          else release_accumulator<CG>(&bars, acc, rank, lane);  // all TMEM reads of this warp are complete
          process(vb, ch);
        } else {
          tmem_ld_block32(taddr + (ch + 1) * 32, vb);
          process(va, ch);
        }
        if (ch == 1 && STASH && has_next)  // column factors of the next tile (ordered by the slab barriers below)
          coltau[acc ^ 1][my_col] = ex2_approx(-fmaf(0.5f * kLog2e, nxt_c, 2.0f));
        if (ch & 1) {  // a 64-column slab of this slice is staged (or, without stash, just keep the slices in step)
          if constexpr (STASH) {
            fence_proxy_async_smem();
            if (slice_tid == 0) tma_store_wait_read<0>();  // every earlier slab store has left shared memory
          }
          named_bar_sync(kBarSlice + slice, 128);
          if constexpr (STASH) {
            if (slice_tid == 0) {
              const int gcol = n0 + col0 + (ch >> 1) * 64;
              if (gcol < P.rows_global)
                tma_store_2d(&P.maps[P.store_map[p]], staging + (slice * 2 + (ch >> 1)) * kSlabBytes, gcol, m0);
              tma_store_commit();
            }
          }
        }
      }
      // column statistics of this slice: the four lane quarters are complete after the last slab barrier
      if (n0 + my_col < P.rows_global)
        P.col_part[(static_cast<size_t>(p) * P.nti + ti) * P.rows_global + n0 + my_col] =
            (colacc[acc][0][my_col] + colacc[acc][1][my_col]) + (colacc[acc][2][my_col] + colacc[acc][3][my_col]);
      {
        float rp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a, b;
          unpack2(rp2[j], a, b);
          rp[j] = (a + b) * isig[j];
        }
        int r;
        const float rsum = frag_row_sum(rp, lane, r);
        if (wrow0 + r < P.rows_local)
          P.row_part[(static_cast<size_t>(p) * P.ntj * 2 + tj * 2 + slice) * P.rows_local + wrow0 + r] = rsum;
      }
      if (slice_tid == 0 && slice == 0) P.tile_ref[(static_cast<size_t>(p) * P.nti + ti) * P.ntj + tj] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) cur_r[j] = nxt_r[j];
    }
    if (STASH && slice_tid == 0) tma_store_wait_all<0>();
  }
  kernel_teardown<CG, EW>(tmem_base, warp);
}

// ============================================================================================== backward tiles
template <int CG, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1) backward_tiles_kernel(const __grid_constant__ BwdParams P) {
  constexpr int S = EW / 4;
  constexpr int CS = BN / S;
  constexpr int NSLAB = CS / 64;   // 64-column slabs per slice
  constexpr uint32_t kEpiThreads = 32 * EW;
  __shared__ PipeBarriers bars;
  __shared__ __align__(8) float colfac[kAccStages][BN];
  __shared__ float redw[kAccStages][kMaxEpiWarps];
  uint8_t* smem = aligned_dyn_smem();
  uint8_t* staging = smem + P.stages * Geo<CG>::kStageBytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  const int nti_c = P.nti / CG;
  const int total = 3 * nti_c * P.ntj;
  // fallback launch behind backward_scale_kernel: only when the forward reported a stash overflow (same word for every CTA)
  if (P.only_if != nullptr && *reinterpret_cast<const volatile int*>(P.only_if) == 0) return;

  const uint32_t tmem_base = kernel_setup<CG, EW>(&bars, warp, lane);

  if (warp == EW) {
    const bool elected = elect_one();
    RingState rs;
    for (int t = cluster_id; t < total; t += num_clusters) {
      const Tile tile = decode_similarity<CG>(t, nti_c, P.ntj);
      producer_tile<CG>(P.maps, P.jobs[tile.job], tile, rank, smem, &bars, P.stages, rs, elected);
    }
  } else if (warp == EW + 1) {
    if (rank == 0) {  // the leader CTA issues the MMAs of the pair
      const bool elected = elect_one();
      RingState rs;
      int it = 0;
      for (int t = cluster_id; t < total; t += num_clusters, ++it) {
        const Tile tile = decode_similarity<CG>(t, nti_c, P.ntj);
        const int acc = it & 1;
        mma_tile<CG>(P.jobs[tile.job], tile, smem, &bars, P.stages, rs, tmem_base + acc * BN, acc, (it >> 1) & 1, elected);
      }
    }
  } else {
    const int q = warp & 3;
    const int slice = warp >> 2;
    const int epi_tid = warp * 32 + lane;
    const int slice_tid = epi_tid & 127;
    const int col0 = slice * CS;
    // c_p = s_p g_p / max_q |s_q g_q|
    float mxsg = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) mxsg = fmaxf(mxsg, fabsf(expf(P.t3[r]) * P.g3[r]));
    int it = 0;
    for (int t = cluster_id; t < total; t += num_clusters, ++it) {
      const Tile tile = decode_similarity<CG>(t, nti_c, P.ntj);
      const int acc = it & 1;
      const int p = tile.job;
      const int ti = tile.ti * CG + static_cast<int>(rank);
      const int tj = tile.tj;
      const int m0 = ti * BM, n0 = tile.n0;
      const float s = expf(P.t3[p]);
      const float c = s * kLog2e * P.acc_scale;
      const float cp = mxsg > 0.f ? (s * P.g3[p]) / mxsg : 0.f;
      const float half_kc = 0.5f * kKappa * cp;
      const bool fast = s < kFastPathMaxScale;
      const int wrow0 = m0 + q * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col0;
      const bool has_lo = P.store_map_lo[p] >= 0;
      const CUtensorMap* map_hi = &P.maps[P.store_map[p]];
      const CUtensorMap* map_lo = has_lo ? &P.maps[P.store_map_lo[p]] : nullptr;

      // per-row / per-column softmax normalisers, prepared while the MMAs of this tile run.
      // fast path: (kappa c_p / 2) / sum of exp(logit); safe path: log-sum-exps in log2 units
      const float* rown = fast ? P.row_inv : P.lse_row;
      const float* coln = fast ? P.col_inv : P.lse_col;
      const float nscale = fast ? half_kc : kLog2e;
      float rf[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = wrow0 + 16 * (j >> 1) + 8 * (j & 1) + (lane >> 2);
        rf[j] = (row < P.rows_local ? rown[static_cast<size_t>(p) * P.rows_local + row] : 0.f) * nscale;
      }
      for (int cc = epi_tid; cc < BN; cc += kEpiThreads)
        colfac[acc][cc] =
            ((n0 + cc < P.rows_global) ? coln[static_cast<size_t>(p) * P.rows_global + n0 + cc] : 0.f) * nscale;
      named_bar_sync(kBarAll, kEpiThreads);

      mbar_wait_bounded<false>(&bars.tmem_full[acc], (it >> 1) & 1, 3);
      tc_fence_after();
      float dtacc = 0.f;
#pragma unroll
      for (int sl = 0; sl < NSLAB; ++sl) {
        // f16: every slab of the tile has its own staging buffer (reused one tile later); f16x3: one hi and one lo
        // buffer per slice, reused by each of its slabs
        uint8_t* stage_hi = staging + (has_lo ? slice * 2 : slice * NSLAB + sl) * kSlabBytes;
        uint8_t* stage_lo = staging + (slice * 2 + 1) * kSlabBytes;
        if (slice_tid == 0) {
          if (has_lo || NSLAB == 1) tma_store_wait_read<0>();
          else tma_store_wait_read<NSLAB - 1>();
        }
        named_bar_sync(kBarSlice + slice, 128);
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          const int ch = sl * 2 + hc;  // 32-column chunk inside this slice
          uint32_t v[32];
          tmem_ld_block32(taddr + ch * 32, v);
          float cf[8];
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float2 f = *reinterpret_cast<const float2*>(&colfac[acc][col0 + ch * 32 + 8 * n + 2 * (lane & 3)]);
            cf[2 * n] = f.x;
            cf[2 * n + 1] = f.y;
          }
          tmem_ld_wait();
          if (sl == NSLAB - 1 && hc == 1) release_accumulator<CG>(&bars, acc, rank, lane);  // last TMEM read
          float g[32];
          if (fast) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float e = ex2_approx(__uint_as_float(v[i]) * c);
              g[i] = e * (rf[2 * (i >> 4) + ((i >> 1) & 1)] + cf[2 * ((i >> 2) & 3) + (i & 1)]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float l2 = __uint_as_float(v[i]) * c;
              g[i] = (ex2_approx(l2 - rf[2 * (i >> 4) + ((i >> 1) & 1)]) +
                      ex2_approx(l2 - cf[2 * ((i >> 2) & 3) + (i & 1)])) * half_kc;
            }
          }
          const int gcol0 = n0 + col0 + ch * 32;
          const int grow0 = P.row_offset + wrow0;
          if (grow0 < gcol0 + 32 && gcol0 < grow0 + 32) {  // the block touches the diagonal: - kappa c_p I
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (grow0 + frag_row(i, lane) == gcol0 + frag_col(i, lane)) g[i] -= kKappa * cp;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) dtacc = fmaf(g[i], __uint_as_float(v[i]), dtacc);
          // fp16 pack + swizzled staging (128-byte rows, 16-byte chunk index XOR row & 7 == TMA SWIZZLE_128B);
          // the four threads of a quad fill one 16-byte chunk
#pragma unroll
          for (int gh = 0; gh < 4; ++gh) {
            const int r_in_tile = q * 32 + 16 * (gh >> 1) + 8 * (gh & 1) + (lane >> 2);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              const int i0 = 16 * (gh >> 1) + 4 * n + 2 * (gh & 1);
              const int chunk16 = hc * 4 + n;
              const uint32_t off = r_in_tile * 128 + ((chunk16 ^ (r_in_tile & 7)) << 4) + 4 * (lane & 3);
              *reinterpret_cast<uint32_t*>(stage_hi + off) = pack_half2(g[i0], g[i0 + 1]);
              if (has_lo) {
                const float l0 = g[i0] - __half2float(__float2half_rn(g[i0]));
                const float l1 = g[i0 + 1] - __half2float(__float2half_rn(g[i0 + 1]));
                *reinterpret_cast<uint32_t*>(stage_lo + off) = pack_half2(l0, l1);
              }
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(kBarSlice + slice, 128);
        const int gcol = n0 + col0 + sl * 64;
        if (slice_tid == 0) {
          if (gcol < P.rows_global) {
            tma_store_2d(map_hi, stage_hi, gcol, m0);
            if (has_lo) tma_store_2d(map_lo, stage_lo, gcol, m0);
          }
          tma_store_commit();
        }
      }
      dtacc = warp_sum(dtacc);
      if (lane == 0) redw[acc][warp] = dtacc;
      named_bar_sync(kBarAll, kEpiThreads);
      if (epi_tid == 0) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < EW; ++w) sum += redw[acc][w];
        P.dt_part[(static_cast<size_t>(p) * P.nti + ti) * P.ntj + tj] = sum * P.acc_scale;
      }
    }
    if (slice_tid == 0) tma_store_wait_all<0>();
  }
  kernel_teardown<CG, EW>(tmem_base, warp);
}

// ============================================================================================== plain GEMM tiles
template <int CG>
__device__ __forceinline__ Tile decode_gemm(const GemmParams& P, int t) {
  Tile r;
  int j = 0;
#pragma unroll
  for (int u = 1; u < kMaxJobs; ++u)
    if (u < P.njobs && t >= P.jobs[u].tile_base) j = u;
  const Job& job = P.jobs[j];
  const int local = t - job.tile_base;
  const int tn = local % job.n_tiles;  // n fastest: the tiles sharing an A row panel run together
  const int rest = local / job.n_tiles;
  r.job = j;
  r.split = rest % job.ksplits;
  r.ti = rest / job.ksplits;
  r.tj = tn;
  r.m0 = r.ti * Geo<CG>::kTileM;
  r.n0 = tn * BN;
  return r;
}

template <int CG, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1) gemm_tiles_kernel(const __grid_constant__ GemmParams P) {
  constexpr int S = EW / 4;
  constexpr int CS = BN / S;
  constexpr int NCH = CS / 32;
  __shared__ PipeBarriers bars;
  uint8_t* smem = aligned_dyn_smem();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  const int total = P.total_tiles;

  const uint32_t tmem_base = kernel_setup<CG, EW>(&bars, warp, lane);

  if (warp == EW) {
    const bool elected = elect_one();
    RingState rs;
    for (int t = cluster_id; t < total; t += num_clusters) {
      const Tile tile = decode_gemm<CG>(P, t);
      producer_tile<CG>(P.maps, P.jobs[tile.job], tile, rank, smem, &bars, P.stages, rs, elected);
    }
  } else if (warp == EW + 1) {
    if (rank == 0) {  // the leader CTA issues the MMAs of the pair
      const bool elected = elect_one();
      RingState rs;
      int it = 0;
      for (int t = cluster_id; t < total; t += num_clusters, ++it) {
        const Tile tile = decode_gemm<CG>(P, t);
        const int acc = it & 1;
        mma_tile<CG>(P.jobs[tile.job], tile, smem, &bars, P.stages, rs, tmem_base + acc * BN, acc, (it >> 1) & 1, elected);
      }
    }
  } else {
    const int q = warp & 3;
    const int slice = warp >> 2;
    float alpha = P.alpha0;
    if (P.t3 != nullptr) {
      float mx = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) mx = fmaxf(mx, fabsf(expf(P.t3[r]) * P.g3[r]));
      alpha *= mx;
    }
    if (P.log_alpha != nullptr) alpha *= expf(*P.log_alpha);  // a learnable log-temperature read on the device
    int it = 0;
    for (int t = cluster_id; t < total; t += num_clusters, ++it) {
      const Tile tile = decode_gemm<CG>(P, t);
      const int acc = it & 1;
      const int j = tile.job;
      const bool accumulate_out = P.jobs[j].ksplits > 1;
      const int row = tile.m0 + static_cast<int>(rank) * BM + q * 32 + lane;
      const bool row_ok = row < P.m[j];
      const int ncols = P.n[j];
      const int c0 = tile.n0 + slice * CS;
      float* out = P.out[j] + static_cast<size_t>(row) * P.ldc[j] + c0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + slice * CS;
      mbar_wait_bounded<false>(&bars.tmem_full[acc], (it >> 1) & 1, 3);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (ch == NCH - 1) release_accumulator<CG>(&bars, acc, rank, lane);
        if (row_ok) {
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const int col = c0 + ch * 32 + k4 * 4;
            if (col + 3 < ncols) {
              float4 o;
              o.x = __uint_as_float(v[k4 * 4 + 0]) * alpha;
              o.y = __uint_as_float(v[k4 * 4 + 1]) * alpha;
              o.z = __uint_as_float(v[k4 * 4 + 2]) * alpha;
              o.w = __uint_as_float(v[k4 * 4 + 3]) * alpha;
              float* dst = out + ch * 32 + k4 * 4;
              if (!accumulate_out) {
                *reinterpret_cast<float4*>(dst) = o;
              } else {  // k-split chunks are combined with round-to-nearest fp32 adds in L2
                red_add_f32(dst + 0, o.x);
                red_add_f32(dst + 1, o.y);
                red_add_f32(dst + 2, o.z);
                red_add_f32(dst + 3, o.w);
              }
            }
          }
        }
      }
    }
  }
  kernel_teardown<CG, EW>(tmem_base, warp);
}

// ============================================================================================== wide GEMM tiles
// Gradient GEMMs (long k, narrow n): one 256 x wn tile per CTA pair, wn = 256 | 384 | 512 accumulator columns in a
// single TMEM accumulator.  Per k block a CTA receives 16 KiB of A and wn * 64 bytes of B for 128 x wn x 64 MACs:
// 157 (wn = 384) / 171 (wn = 512) flop per byte through the L2 -> SM fabric instead of 128 with the 256-column tiles,
// which is what bounds the 256-column mainloop (measured: 43 B/clk/SM delivered, the chip-wide TMA ceiling).  With a
// tile taking hundreds of k blocks the epilogue does not need a second accumulator to hide behind.
//   B operands are MN-major only (xhat stored [sample][feature], k = sample).
//   MMA 1 covers accumulator columns [0, 256): CTA r of the pair supplies columns 128 r .. 128 r + 127 (two boxes),
//   MMA 2 covers columns [256, wn):            CTA r supplies (wn - 256) / 2 columns (one or two boxes).
struct WideBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ Tile decode_wide(const GemmParams& P, int t) {
  Tile r;
  int j = 0;
#pragma unroll
  for (int u = 1; u < kMaxJobs; ++u)
    if (u < P.njobs && t >= P.jobs[u].tile_base) j = u;
  const Job& job = P.jobs[j];
  const int local = t - job.tile_base;
  const int tn = local % job.n_tiles;  // n fastest: the tiles sharing an A row panel run together
  const int rest = local / job.n_tiles;
  r.job = j;
  r.split = rest % job.ksplits;
  r.ti = rest / job.ksplits;
  r.tj = tn;
  r.m0 = r.ti * (2 * BM);
  r.n0 = tn * P.wn;
  return r;
}

template <int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1) gemm_wide_kernel(const __grid_constant__ GemmParams P) {
  constexpr int S = EW / 4;
  __shared__ WideBarriers bars;
  uint8_t* smem = aligned_dyn_smem();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t half = cluster_ctarank();  // which 128 rows of the pair's 256; CTA 0 issues the MMAs
  const int cluster_id = blockIdx.x / 2, num_clusters = gridDim.x / 2;
  const int total = P.total_tiles;
  const int wn = P.wn;
  const int n2 = wn - 256;                          // columns of the second MMA (0, 128 or 256)
  const int stage_bytes = A_STAGE_BYTES + wn * 64;  // A 128 x 64 + B (wn / 2) x 64 fp16
  const int stages = P.stages;

  if (warp == EW && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bars.full[i], 1);
      mbar_init(&bars.empty[i], 1);
    }
    mbar_init(&bars.tmem_full, 1);
    mbar_init(&bars.tmem_empty, 2 * EW);
    fence_mbar_init();
  }
  if (warp == EW + 1) {
    tmem_alloc<2>(&bars.tmem_base, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars.tmem_base);
  const uint32_t smem0 = smem_u32(smem);

  if (warp == EW) {  // TMA producer (convergent warp code, one elected lane issues)
    const bool elected = elect_one();
    RingState rs;
    const uint32_t full_loc0 = smem_u32(&bars.full[0]);
    const uint32_t full_sig0 = mapa(full_loc0, 0);
    for (int u = cluster_id; u < total; u += num_clusters) {
      const Tile tile = decode_wide(P, u);
      const Job& job = P.jobs[tile.job];
      const int m0 = tile.m0 + static_cast<int>(half) * BM;
      const int nb1 = tile.n0 + static_cast<int>(half) * 128;             // this CTA's columns of MMA 1
      const int nb2 = tile.n0 + 256 + static_cast<int>(half) * (n2 / 2);  // ... of MMA 2
      for (int s = 0; s < job.nseg; ++s) {
        const CUtensorMap* ma = P.maps + job.seg[s].map_a;
        const CUtensorMap* mb = P.maps + job.seg[s].map_b;
        const int a_mn = job.seg[s].a_mn;
        int kb_lo, kb_hi;
        split_range(job.seg[s].num_kb, tile.split, job.ksplits, kb_lo, kb_hi);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait_bounded<false>(&bars.empty[rs.stage], rs.phase ^ 1u, 1);
          if (elected) {
            const uint32_t sa = smem0 + rs.stage * stage_bytes;
            const uint32_t sb = sa + A_STAGE_BYTES;
            const uint32_t bar = full_sig0 + rs.stage * 8;
            const int k = kb * BK;
            if (half == 0) mbar_expect_tx_u32(full_loc0 + rs.stage * 8, 2 * stage_bytes);
            if (!a_mn) {
              tma_load_2d_2sm_u32(sa, ma, bar, k, m0);
            } else {
              tma_load_2d_2sm_u32(sa, ma, bar, m0, k);
              tma_load_2d_2sm_u32(sa + MN_BOX_BYTES, ma, bar, m0 + 64, k);
            }
            tma_load_2d_2sm_u32(sb, mb, bar, nb1, k);
            tma_load_2d_2sm_u32(sb + MN_BOX_BYTES, mb, bar, nb1 + 64, k);
            if (n2 >= 128) tma_load_2d_2sm_u32(sb + 2 * MN_BOX_BYTES, mb, bar, nb2, k);
            if (n2 >= 256) tma_load_2d_2sm_u32(sb + 3 * MN_BOX_BYTES, mb, bar, nb2 + 64, k);
          }
          rs.advance(stages);
        }
      }
    }
  } else if (warp == EW + 1) {
    if (half == 0) {  // MMA issuer (leader CTA; convergent warp code, one elected lane issues)
      const bool elected = elect_one();
      RingState rs;
      int it = 0;
      const uint32_t empty0 = smem_u32(&bars.empty[0]);
      const uint32_t tfull = smem_u32(&bars.tmem_full);
      constexpr uint32_t hi = smem_desc_hi_sw128(1024);
      for (int u = cluster_id; u < total; u += num_clusters, ++it) {
        const Tile tile = decode_wide(P, u);
        const Job& job = P.jobs[tile.job];
        mbar_wait_bounded<false>(&bars.tmem_empty, (it & 1) ^ 1u, 4);  // the epilogue has drained the accumulator
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int s = 0; s < job.nseg; ++s) {
          const int a_mn = job.seg[s].a_mn;
          const uint32_t idesc1 = make_idesc_f16(2 * BM, 256, 0, a_mn, 1);
          const uint32_t idesc2 = make_idesc_f16(2 * BM, n2 > 0 ? n2 : 256, 0, a_mn, 1);
          const uint32_t a_lbo = a_mn ? OperandDesc<true>::kLbo : OperandDesc<false>::kLbo;
          const uint32_t a_step = a_mn ? OperandDesc<true>::kStep : OperandDesc<false>::kStep;
          int kb_lo, kb_hi;
          split_range(job.seg[s].num_kb, tile.split, job.ksplits, kb_lo, kb_hi);
          for (int kb = kb_lo; kb < kb_hi; ++kb) {
            mbar_wait_bounded<false>(&bars.full[rs.stage], rs.phase, 2);
            tc_fence_after();
            if (elected) {
              const uint32_t a_lo = smem_desc_lo(smem0 + rs.stage * stage_bytes, a_lbo);
              const uint32_t b_lo = smem_desc_lo(smem0 + rs.stage * stage_bytes + A_STAGE_BYTES, MN_BOX_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t adesc = smem_desc_join(a_lo + k * a_step, hi);
                umma_f16<2>(tmem_base, adesc, smem_desc_join(b_lo + k * OperandDesc<true>::kStep, hi), idesc1, accumulate);
                if (n2 > 0)
                  umma_f16<2>(tmem_base + 256, adesc,
                              smem_desc_join(b_lo + ((2 * MN_BOX_BYTES) >> 4) + k * OperandDesc<true>::kStep, hi), idesc2,
                              accumulate);
                accumulate = 1;
              }
              umma_commit_2sm_u32(empty0 + rs.stage * 8, 0b11);  // frees the slot in both CTAs
            }
            rs.advance(stages);
          }
        }
        if (elected) umma_commit_2sm_u32(tfull, 0b11);
      }
    }
  } else {
    const int q = warp & 3;
    const int slice = warp >> 2;
    const int cs = wn / S;  // columns per slice (a multiple of 32)
    float alpha = P.alpha0;
    if (P.t3 != nullptr) {
      float mx = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) mx = fmaxf(mx, fabsf(expf(P.t3[r]) * P.g3[r]));
      alpha *= mx;
    }
    int it = 0;
    for (int u = cluster_id; u < total; u += num_clusters, ++it) {
      const Tile tile = decode_wide(P, u);
      const int j = tile.job;
      const bool accumulate_out = P.jobs[j].ksplits > 1;
      const int row = tile.m0 + static_cast<int>(half) * BM + q * 32 + lane;
      const bool row_ok = row < P.m[j];
      const int ncols = P.n[j];
      const int c0 = tile.n0 + slice * cs;
      float* out = P.out[j] + static_cast<size_t>(row) * P.ldc[j] + c0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slice * cs;
      mbar_wait_bounded<false>(&bars.tmem_full, it & 1, 3);
      tc_fence_after();
      const int nch = cs / 32;
      for (int ch = 0; ch < nch; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (ch == nch - 1) {  // last TMEM read of this warp: hand the accumulator back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (half == 0) mbar_arrive(&bars.tmem_empty);
            else mbar_arrive_cluster(&bars.tmem_empty, 0);
          }
        }
        if (row_ok) {
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const int col = c0 + ch * 32 + k4 * 4;
            if (col + 3 < ncols) {
              float4 o;
              o.x = __uint_as_float(v[k4 * 4 + 0]) * alpha;
              o.y = __uint_as_float(v[k4 * 4 + 1]) * alpha;
              o.z = __uint_as_float(v[k4 * 4 + 2]) * alpha;
              o.w = __uint_as_float(v[k4 * 4 + 3]) * alpha;
              float* dst = out + ch * 32 + k4 * 4;
              if (!accumulate_out) *reinterpret_cast<float4*>(dst) = o;
              else red_add_v4(dst, o);  // k-split partial sums are combined by fp32 adds in L2
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync();  // no CTA of the pair may exit while the other can still signal its barriers
  if (warp == EW + 1) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

// ============================================================================================== converting GEMM
// The gradient GEMMs of the stash path with the stash -> G' conversion inside the A-operand path, so that the B x B
// strips make no separate trip through HBM (the in-place pass is 17 % of the north-star step and executes no flop).
// The A operand never goes back to shared memory: TMA delivers the raw stash tile, eight converter warps (thread =
// accumulator row = TMEM lane) multiply it by the rank-2 factor R1_i C1_j + R2_i C2_j, subtract the identity term in
// fp32, round to fp16 and write it with tcgen05.st into one of four 32-column TMEM slots beside the 384-column
// accumulator (384 + 4 * 32 = 512 columns); the MMAs take A from TMEM.  Shared-memory traffic per k block is that of
// the plain kernel (the converter's read replaces the tensor core's read of A), which is what bounded the in-smem
// conversion of round 1.  Values are bit-identical to backward_scale_kernel + gemm_wide_kernel: same fp32 expression,
// same rounding, same MMA order.
//   warps 0-7   epilogue (as gemm_wide_kernel)
//   warps 8-15  converters: lane quarter q = warp & 3; warps 8-11 take the even k blocks, 12-15 the odd ones (a warp
//               needs ~2 k-block times per block it converts: two barrier waits, the loads, ~280 instructions, the
//               TMEM store and its completion; splitting one block over two warps left that latency on every block)
//   warp 16     TMA producer of the raw stash tiles + k-side factors   (ring of kRawStages)
//   warp 17     TMA producer of the B operand                          (ring of kBStages)
//   warp 18     MMA issuer (leader CTA)
// Two rings because the two operands have different latencies and lifetimes: the stash streams from HBM (2-3 k-block
// times away) but is only needed until it is converted; the B operand comes from L2 and has to stay until its MMAs
// have completed.  With one 5-stage ring the conversion sat in the middle of the load -> MMA -> free -> load cycle and
// the kernel ran at (HBM latency + conversion + MMA) / 5 per k block.
constexpr int kConvWarps = 8;
constexpr int kASlots = 4;
constexpr int kConvWn = 384;
constexpr int kRawStages = 6;
constexpr int kBStages = 4;
constexpr int kConvFacBytes = 512;                             // 64 pair-interleaved fp32 k-side factor pairs per stage
constexpr int kRawBytes = kRawStages * A_STAGE_BYTES;          // raw tiles (1 KiB aligned for the 128-byte swizzle),
                                                               // then the factor blocks, then the B ring
constexpr int kConvBBytes = kConvWn * 64;                      // 24 KiB of B per CTA and k block

struct ConvBarriers {
  uint64_t bfull[kBStages];     // leader: the B bytes of both CTAs have landed
  uint64_t bempty[kBStages];    // each CTA: the MMAs that read this B slot have completed (1 commit)
  uint64_t araw[kRawStages];    // each CTA: its raw A tile and the k-side factors have landed
  uint64_t aempty[kRawStages];  // each CTA: the four converter warps of the block's parity are done with the raw tile
  uint64_t aconv[kASlots];      // leader: the slot is converted in both CTAs (2 x 4 warp arrivals)
  uint64_t afree[kASlots];      // each CTA: the MMAs that read the slot have completed (1 commit)
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(32 * (8 + kConvWarps + 3), 1) gemm_conv_kernel(const __grid_constant__ GemmParams P) {
  constexpr int EW = 8, S = 2;
  constexpr int kProducer = EW + kConvWarps, kProducerB = kProducer + 1, kIssuer = kProducer + 2;
  constexpr int n2 = kConvWn - 256;
  __shared__ ConvBarriers bars;
  uint8_t* smem = aligned_dyn_smem();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t half = cluster_ctarank();
  const int cluster_id = blockIdx.x / 2, num_clusters = gridDim.x / 2;
  const int total = P.total_tiles;

  if (warp == kProducer && lane == 0) {
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bars.bfull[i], 1);
      mbar_init(&bars.bempty[i], 1);
    }
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&bars.araw[i], 1);
      mbar_init(&bars.aempty[i], kConvWarps / 2);
    }
    for (int i = 0; i < kASlots; ++i) {
      mbar_init(&bars.aconv[i], kConvWarps);  // 4 warps (one per lane quarter) of each CTA
      mbar_init(&bars.afree[i], 1);
    }
    mbar_init(&bars.tmem_full, 1);
    mbar_init(&bars.tmem_empty, 2 * EW);
    fence_mbar_init();
  }
  if (warp == kIssuer) {
    tmem_alloc<2>(&bars.tmem_base, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars.tmem_base);
  const uint32_t smem0 = smem_u32(smem);                     // raw ring: kRawStages tiles, then kRawStages factor blocks
  uint8_t* smem_f = smem + kRawBytes;
  uint8_t* smem_b = smem_f + kRawStages * kConvFacBytes;     // B ring: kBStages x 24 KiB (99 KiB in: 1 KiB aligned)
  const uint32_t smem_b0 = smem_u32(smem_b);
  const uint32_t a_slot0 = tmem_base + kConvWn;              // four 32-column A slots behind the accumulator

  if (warp == kProducer) {  // raw stash tiles + k-side factors -> this CTA's araw barriers
    const bool elected = elect_one();
    RingState rs;
    const uint32_t araw0 = smem_u32(&bars.araw[0]);
    for (int u = cluster_id; u < total; u += num_clusters) {
      const Tile tile = decode_wide(P, u);
      const Job& job = P.jobs[tile.job];
      const int m0 = tile.m0 + static_cast<int>(half) * BM;
      for (int s = 0; s < job.nseg; ++s) {
        const CUtensorMap* ma = P.maps + job.seg[s].map_a;
        const int a_mn = job.seg[s].a_mn, p = job.seg[s].pair;
        // k-side factors, pair-interleaved (b1_k, b1_k+1, b2_k, b2_k+1), 512 bytes per k block: row role (K-major A)
        // k = global column -> the column factors; column role k = local row -> the row factors
        const int kld = a_mn ? P.ld_row : P.ld_col;
        const float* kpk = (a_mn ? P.fac_row : P.fac_col) + static_cast<size_t>(6 + 2 * p) * kld;
        int kb_lo, kb_hi;
        split_range(job.seg[s].num_kb, tile.split, job.ksplits, kb_lo, kb_hi);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait_bounded<false>(&bars.aempty[rs.stage], rs.phase ^ 1u, 1);
          if (elected) {
            const uint32_t sa = smem0 + rs.stage * A_STAGE_BYTES;
            const uint32_t abar = araw0 + rs.stage * 8;
            const int k = kb * BK;
            mbar_expect_tx_u32(abar, A_STAGE_BYTES + kConvFacBytes);
            if (!a_mn) {
              tma_load_2d_u32(sa, ma, abar, k, m0);
            } else {
              tma_load_2d_u32(sa, ma, abar, m0, k);
              tma_load_2d_u32(sa + MN_BOX_BYTES, ma, abar, m0 + 64, k);
            }
            bulk_load_1d_u32(smem0 + kRawBytes + rs.stage * kConvFacBytes, kpk + 2 * k, kConvFacBytes, abar);
          }
          rs.advance(kRawStages);
        }
      }
    }
  } else if (warp == kProducerB) {  // B operand of both CTAs -> the leader's bfull barriers
    const bool elected = elect_one();
    RingState rs;
    const uint32_t bfull_loc0 = smem_u32(&bars.bfull[0]);
    const uint32_t bfull_sig0 = mapa(bfull_loc0, 0);
    for (int u = cluster_id; u < total; u += num_clusters) {
      const Tile tile = decode_wide(P, u);
      const Job& job = P.jobs[tile.job];
      const int nb1 = tile.n0 + static_cast<int>(half) * 128;
      const int nb2 = tile.n0 + 256 + static_cast<int>(half) * (n2 / 2);
      for (int s = 0; s < job.nseg; ++s) {
        const CUtensorMap* mb = P.maps + job.seg[s].map_b;
        int kb_lo, kb_hi;
        split_range(job.seg[s].num_kb, tile.split, job.ksplits, kb_lo, kb_hi);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait_bounded<false>(&bars.bempty[rs.stage], rs.phase ^ 1u, 1);
          if (elected) {
            const uint32_t sb = smem_b0 + rs.stage * kConvBBytes;
            const uint32_t bbar = bfull_sig0 + rs.stage * 8;
            const int k = kb * BK;
            if (half == 0) mbar_expect_tx_u32(bfull_loc0 + rs.stage * 8, 2 * kConvBBytes);
            tma_load_2d_2sm_u32(sb, mb, bbar, nb1, k);
            tma_load_2d_2sm_u32(sb + MN_BOX_BYTES, mb, bbar, nb1 + 64, k);
            tma_load_2d_2sm_u32(sb + 2 * MN_BOX_BYTES, mb, bbar, nb2, k);
          }
          rs.advance(kBStages);
        }
      }
    }
  } else if (warp == kIssuer) {
    if (half == 0) {
      const bool elected = elect_one();
      RingState rs, as;
      int it = 0;
      const uint32_t bempty0 = smem_u32(&bars.bempty[0]);
      const uint32_t afree0 = smem_u32(&bars.afree[0]);
      const uint32_t tfull = smem_u32(&bars.tmem_full);
      constexpr uint32_t hi = smem_desc_hi_sw128(1024);
      // A comes from TMEM (always "K-major"), B is MN-major in shared memory
      constexpr uint32_t idesc1 = make_idesc_f16(2 * BM, 256, 0, 0, 1);
      constexpr uint32_t idesc2 = make_idesc_f16(2 * BM, n2, 0, 0, 1);
      for (int u = cluster_id; u < total; u += num_clusters, ++it) {
        const Tile tile = decode_wide(P, u);
        const Job& job = P.jobs[tile.job];
        mbar_wait_bounded<false>(&bars.tmem_empty, (it & 1) ^ 1u, 4);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int s = 0; s < job.nseg; ++s) {
          int kb_lo, kb_hi;
          split_range(job.seg[s].num_kb, tile.split, job.ksplits, kb_lo, kb_hi);
          for (int kb = kb_lo; kb < kb_hi; ++kb) {
            mbar_wait_bounded<false>(&bars.bfull[rs.stage], rs.phase, 2);
            mbar_wait_bounded<true>(&bars.aconv[as.stage], as.phase, 5);  // arrivals come from both CTAs
            tc_fence_after();
            if (elected) {
              const uint32_t b_lo = smem_desc_lo(smem_b0 + rs.stage * kConvBBytes, MN_BOX_BYTES);
              const uint32_t a_t = a_slot0 + as.stage * 32;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                umma_f16_ts_2sm(tmem_base, a_t + k * 8, smem_desc_join(b_lo + k * OperandDesc<true>::kStep, hi), idesc1,
                                accumulate);
                umma_f16_ts_2sm(tmem_base + 256, a_t + k * 8,
                                smem_desc_join(b_lo + ((2 * MN_BOX_BYTES) >> 4) + k * OperandDesc<true>::kStep, hi), idesc2,
                                accumulate);
                accumulate = 1;
              }
              umma_commit_2sm_u32(bempty0 + rs.stage * 8, 0b11);  // the B slot (both CTAs)
              umma_commit_2sm_u32(afree0 + as.stage * 8, 0b11);   // the TMEM slot
            }
            rs.advance(kBStages);
            as.advance(kASlots);
          }
        }
        if (elected) umma_commit_2sm_u32(tfull, 0b11);
      }
    }
  } else if (warp >= EW) {
    // ---- converters.  A warp converts 32 accumulator rows (its TMEM lane quarter) x 64 k of every other k block.
    //   row role (K-major raw tile): thread = row; 4 x LDS.128 of its swizzled 128-byte row, 16 packed factor entries,
    //     one tcgen05.st.32x32b.x16.
    //   column role (MN-major raw tile, the accumulator row is a COLUMN of the stash): ldmatrix.trans with row addresses
    //     chosen so that every thread receives, for its lanes (t/4, t/4 + 8), the k pairs 4 (t%4) + 2 c + {0, 1} + 16 n --
    //     exactly the register layout of tcgen05.st.16x256b; 4 m-side factor pairs in registers, 4 k-side entries.
    const int cw = warp - EW;
    const int q = cw & 3, parity = cw >> 2;
    RingState rs, as;
    int kcount = 0;  // k blocks seen so far (all tiles): this warp converts those with kcount % 2 == parity
    const uint32_t aconv_sig0 = mapa(smem_u32(&bars.aconv[0]), 0);
    const uint32_t aempty0 = smem_u32(&bars.aempty[0]);
    float mxsg = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) mxsg = fmaxf(mxsg, fabsf(expf(P.t3[r]) * P.g3[r]));
    auto slot_ready = [&]() {
      mbar_wait_bounded<false>(&bars.araw[rs.stage], rs.phase, 6);
      mbar_wait_bounded<true>(&bars.afree[as.stage], as.phase ^ 1u, 7);
      tc_fence_after();
    };
    auto slot_done = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive_u32(aempty0 + rs.stage * 8);  // every lane has read its part of the raw tile
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_u32(aconv_sig0 + as.stage * 8);  // counted in the leader CTA
      rs.advance(kRawStages);
      as.advance(kASlots);
    };
    // G' pair = E~ pair * (a1 b1 + a2 b2) in fp32, rounded to fp16; `entry` = (b1_k, b1_k+1, b2_k, b2_k+1)
    auto convert_pair = [&](uint32_t e2, unsigned long long a1p, unsigned long long a2p, const ulonglong2& entry, float& lo,
                            float& hi) {
      const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&e2));
      unpack2(fmul2(pack2(e.x, e.y), ffma2(a1p, entry.x, fmul2(a2p, entry.y))), lo, hi);
    };
    for (int u = cluster_id; u < total; u += num_clusters) {
      const Tile tile = decode_wide(P, u);
      const Job& job = P.jobs[tile.job];
      const int tile_row0 = tile.m0 + static_cast<int>(half) * BM + q * 32;  // first output row of this warp
      for (int s = 0; s < job.nseg; ++s) {
        const int a_mn = job.seg[s].a_mn, p = job.seg[s].pair;
        const float kcp = mxsg > 0.f ? kKappa * expf(P.t3[p]) * P.g3[p] / mxsg : 0.f;
        int kb_lo, kb_hi;
        split_range(job.seg[s].num_kb, tile.split, job.ksplits, kb_lo, kb_hi);
        if (!a_mn) {
          // ---------------- row role: thread = row `mrow`, k = global column
          const int row = q * 32 + lane;
          const int mrow = tile_row0 + lane;
          const float* mf = P.fac_row + static_cast<size_t>(p) * 2 * P.ld_row;
          const float a1 = mrow < P.ld_row ? __ldg(mf + mrow) : 0.f;
          const float a2 = mrow < P.ld_row ? __ldg(mf + P.ld_row + mrow) : 0.f;
          const unsigned long long a1p = pack2(a1, a1), a2p = pack2(a2, a2);
          const int kdiag = P.row_offset + mrow;  // the positive pair of this row sits at this global column
          const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
          const uint32_t sw = static_cast<uint32_t>(row & 7);
          for (int kb = kb_lo; kb < kb_hi; ++kb, ++kcount) {
            if ((kcount & 1) != parity) {
              rs.advance(kRawStages);
              as.advance(kASlots);
              continue;
            }
            slot_ready();
            const uint8_t* sa_g = smem + rs.stage * A_STAGE_BYTES + row_off;
#pragma unroll 1
            for (int kh = 0; kh < 2; ++kh) {  // the two 32-k halves of the block
            const uint8_t* sf_g = smem_f + rs.stage * kConvFacBytes + kh * 256;
            uint32_t out[16];
            const int dk = kdiag - (kb * BK + kh * 32);  // in [0, 32): element dk of this thread is the positive pair
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // 8 k: 16-byte chunk (4 kh + c) of the row, at position chunk ^ (row & 7).  Plain loads (not volatile asm):
              // the compiler may batch them; the barrier waits above carry memory clobbers, nothing moves across them
              const uint4 raw = *reinterpret_cast<const uint4*>(sa_g + (((static_cast<uint32_t>(kh * 4 + c)) ^ sw) << 4));
              const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const ulonglong2 entry = *reinterpret_cast<const ulonglong2*>(sf_g + (c * 4 + t) * 16);
                float lo, hi;
                convert_pair(w4[t], a1p, a2p, entry, lo, hi);
                out[c * 4 + t] = pack_half2(lo, hi);
              }
            }
            if (static_cast<unsigned>(dk) < 32u) {  // rare: redo the pair that holds the identity term, in fp32
              const int j = dk >> 1;
              const uint32_t e2 = *reinterpret_cast<const uint32_t*>(
                  sa_g + (((static_cast<uint32_t>(kh * 4 + (j >> 2))) ^ sw) << 4) + (j & 3) * 4);
              const ulonglong2 entry = *reinterpret_cast<const ulonglong2*>(sf_g + j * 16);
              float lo, hi;
              convert_pair(e2, a1p, a2p, entry, lo, hi);
              if (dk & 1) hi -= kcp;
              else lo -= kcp;
              const uint32_t fixed = pack_half2(lo, hi);
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i == j) out[i] = fixed;
            }
            tmem_st_32x32b_x16(a_slot0 + as.stage * 32 + kh * 16 + (static_cast<uint32_t>(q * 32) << 16), out);
            }
            slot_done();
          }
        } else {
          // ---------------- column role: accumulator row = global column of the stash, k = local row
          const int t4 = lane >> 2, tq = lane & 3;
          const float* mf = P.fac_col + static_cast<size_t>(p) * 2 * P.ld_col;
          unsigned long long a1p[2][2], a2p[2][2];  // [g][h]: lanes 16 g + 8 h + t / 4 of this warp's quarter
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int mrow = tile_row0 + 16 * g + 8 * h + t4;
              const float a1 = mrow < P.ld_col ? __ldg(mf + mrow) : 0.f;
              const float a2 = mrow < P.ld_col ? __ldg(mf + P.ld_col + mrow) : 0.f;
              a1p[g][h] = pack2(a1, a1);
              a2p[g][h] = pack2(a2, a2);
            }
          // ldmatrix row addresses: lane L supplies row r = L & 7 of matrix j = L >> 3 = 2 h + c:
          //   k row (within the block)  32 kh + 16 n + 4 (r >> 1) + 2 c + (r & 1),  m chunk of (32 q + 16 g + 8 h)
          const int lj = lane >> 3, lr = lane & 7;
          const int lh = lj >> 1, lc = lj & 1;
          const int kk0 = 4 * (lr >> 1) + 2 * lc + (lr & 1);  // + 16 n + 32 kh
          uint32_t ld_off[2][2];  // [g][n]: byte offset inside a stage, for kh = 0 (kh = 1: + 32 rows = 4096 bytes; the
                                  // swizzle term only depends on kk & 7, which 32 more rows do not change)
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              const int m_local = q * 32 + 16 * g + 8 * lh;  // first of the 8 m of this matrix (CTA-local accumulator row)
              const int kk = kk0 + 16 * n;
              ld_off[g][n] = static_cast<uint32_t>((m_local >> 6) * MN_BOX_BYTES + kk * 128 +
                                                   ((((m_local & 63) >> 3) ^ (kk & 7)) << 4));
            }
          for (int kb = kb_lo; kb < kb_hi; ++kb, ++kcount) {
            if ((kcount & 1) != parity) {
              rs.advance(kRawStages);
              as.advance(kASlots);
              continue;
            }
            slot_ready();
#pragma unroll 1
            for (int kh = 0; kh < 2; ++kh) {  // the two 32-k halves of the block
            const uint32_t sa = smem0 + rs.stage * A_STAGE_BYTES + kh * 4096;
            const uint8_t* sf_g = smem_f + rs.stage * kConvFacBytes + kh * 256;
            ulonglong2 entry[2][2];  // [n][c]: k pair 32 kh + 16 n + 4 (t % 4) + 2 c of the block
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
              for (int c = 0; c < 2; ++c)
                entry[n][c] = *reinterpret_cast<const ulonglong2*>(sf_g + (8 * n + 2 * tq + c) * 16);
            const int krow0 = P.row_offset + kb * BK + kh * 32;  // global row of this warp's first k
            // identity term: element (lane 16 g + 8 h + t / 4, k 16 n + 4 (t % 4) + 2 c + {0, 1}) is a positive pair when
            // d0 + 16 g + 8 h - 16 n - 2 c is 0 or 1; only threads with d0 in [-24, 19] can hold one (rare: one warp
            // per 32 columns and k block), the others take the loop without the per-element test
            const int d0 = (tile_row0 + t4) - (krow0 + 4 * tq);
            auto convert_half = [&](int g, auto diag_tag) {
              constexpr bool kDiag = decltype(diag_tag)::value;
              uint32_t out[8];
#pragma unroll
              for (int n = 0; n < 2; ++n) {
                uint32_t r4[4];
                ldmatrix_x4_trans(sa + ld_off[g][n], r4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int h = j >> 1, c = j & 1;
                  float lo, hi;
                  convert_pair(r4[j], a1p[g][h], a2p[g][h], entry[n][c], lo, hi);
                  if constexpr (kDiag) {  // in fp32, before the rounding to fp16
                    const int dj = d0 + 16 * g + 8 * h - 16 * n - 2 * c;
                    if (dj == 0) lo -= kcp;
                    if (dj == 1) hi -= kcp;
                  }
                  out[4 * n + j] = pack_half2(lo, hi);
                }
              }
              tmem_st_16x256b_x2(a_slot0 + as.stage * 32 + kh * 16 + (static_cast<uint32_t>(q * 32 + 16 * g) << 16), out);
            };
            if (__any_sync(0xffffffffu, d0 >= -24 && d0 <= 19)) {
              convert_half(0, std::true_type{});
              convert_half(1, std::true_type{});
            } else {
              convert_half(0, std::false_type{});
              convert_half(1, std::false_type{});
            }
            }
            slot_done();
          }
        }
      }
    }
  } else {
    // ---- epilogue (as gemm_wide_kernel)
    const int q = warp & 3;
    const int slice = warp >> 2;
    constexpr int cs = kConvWn / S;
    float alpha = P.alpha0;
    {
      float mx = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) mx = fmaxf(mx, fabsf(expf(P.t3[r]) * P.g3[r]));
      alpha *= mx;
    }
    int it = 0;
    for (int u = cluster_id; u < total; u += num_clusters, ++it) {
      const Tile tile = decode_wide(P, u);
      const int j = tile.job;
      const bool accumulate_out = P.jobs[j].ksplits > 1;
      const int orow = tile.m0 + static_cast<int>(half) * BM + q * 32 + lane;
      const bool row_ok = orow < P.m[j];
      const int ncols = P.n[j];
      const int c0 = tile.n0 + slice * cs;
      float* outp = P.out[j] + static_cast<size_t>(orow) * P.ldc[j] + c0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slice * cs;
      mbar_wait_bounded<false>(&bars.tmem_full, it & 1, 3);
      tc_fence_after();
      constexpr int nch = cs / 32;
      for (int ch = 0; ch < nch; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (ch == nch - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (half == 0) mbar_arrive(&bars.tmem_empty);
            else mbar_arrive_cluster(&bars.tmem_empty, 0);
          }
        }
        if (row_ok) {
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const int col = c0 + ch * 32 + k4 * 4;
            if (col + 3 < ncols) {
              float4 o;
              o.x = __uint_as_float(v[k4 * 4 + 0]) * alpha;
              o.y = __uint_as_float(v[k4 * 4 + 1]) * alpha;
              o.z = __uint_as_float(v[k4 * 4 + 2]) * alpha;
              o.w = __uint_as_float(v[k4 * 4 + 3]) * alpha;
              float* dst = outp + ch * 32 + k4 * 4;
              if (!accumulate_out) *reinterpret_cast<float4*>(dst) = o;
              else red_add_v4(dst, o);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync();
  if (warp == kIssuer) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host launchers
}  // namespace

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
  }
  return cached;
}

namespace {

// cudaFuncSetAttribute once per (kernel, device) instead of on every launch
int ensure_smem_attribute(const void* kernel, int smem_bytes) {
  struct Entry {
    const void* fn;
    int dev;
    int bytes;
  };
  static Entry table[64];
  static int used = 0;
  static std::mutex mu;
  int dev = 0;
  SCLIP_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (table[i].fn == kernel && table[i].dev == dev) {
      if (table[i].bytes >= smem_bytes) return SCLIP_OK;
      SCLIP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      table[i].bytes = smem_bytes;
      return SCLIP_OK;
    }
  SCLIP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  if (used < 64) table[used++] = Entry{kernel, dev, smem_bytes};
  return SCLIP_OK;
}

template <class Params>
int launch_persistent(void (*kernel)(Params), const Params& p, int cg, int ew, int smem_bytes, int total_cluster_tiles,
                      int max_sms, cudaStream_t stream) {
  const int rc = ensure_smem_attribute(reinterpret_cast<const void*>(kernel), smem_bytes);
  if (rc) return rc;
  int sms = sm_count();
  if (max_sms > 0 && max_sms < sms) sms = max_sms;
  int clusters = sms / cg;
  if (total_cluster_tiles < clusters) clusters = total_cluster_tiles;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * cg, 1, 1);
  cfg.blockDim = dim3(64 + 32 * ew, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  count_launch();
  SCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p));
  return SCLIP_OK;
}

// dispatch on the two compile-time knobs (CTA group 1 | 2, epilogue warps 8 | 16)
#define SCLIP_DISPATCH(KERNEL, ...)                                                    \
  do {                                                                                 \
    if (cg == 2 && ew == 16) return launch_persistent(KERNEL<2, 16>, p, 2, 16, __VA_ARGS__); \
    if (cg == 2) return launch_persistent(KERNEL<2, 8>, p, 2, 8, __VA_ARGS__);           \
    if (ew == 16) return launch_persistent(KERNEL<1, 16>, p, 1, 16, __VA_ARGS__);        \
    return launch_persistent(KERNEL<1, 8>, p, 1, 8, __VA_ARGS__);                        \
  } while (0)

}  // namespace

int staging_slabs(int ew, bool split) { return split ? (ew / 4) * 2 : 4; }

int tile_smem_bytes(int cg, int stages, int slabs) {
  const int stage_bytes = cg == 2 ? Geo<2>::kStageBytes : Geo<1>::kStageBytes;
  return stages * stage_bytes + slabs * kSlabBytes + 1024;
}

int launch_forward_tiles(const FwdParams& p0, int cg, int ew, int max_sms, cudaStream_t stream) {
  FwdParams p = p0;
  const int smem = tile_smem_bytes(cg, p.stages, p.stash ? 4 : 0);
  const int total = p.npairs * (p.nti / cg) * p.tj_count;
  if (total <= 0) return SCLIP_OK;
  static const bool fast_on = [] {
    const char* e = getenv("SCLIP_FAST");
    return !(e != nullptr && e[0] == '0');
  }();
  if (cg == 2 && fast_on) {
    // pairs with s < kFoldMaxScale (decided on the device: the scales are never read by the host), then the rest
    const int rc = p.stash ? launch_persistent(forward_fast_kernel<true>, p, 2, 8, smem, total, max_sms, stream)
                           : launch_persistent(forward_fast_kernel<false>, p, 2, 8, smem, total, max_sms, stream);
    if (rc) return rc;
    p.pair_filter = 1;
    if (p.wait_peers) {  // forward_fast_kernel has waited for every shard: the rest runs in the plain column order
      p.wait_peers = 0;
      p.tj_begin = 0;
      p.tj_count = p.ntj;
    }
  } else if (p.wait_peers) {
    set_error("SCLIP_FWD_WAIT_PEERS needs the CTA-pair forward kernel (SCLIP_CTA_GROUP=2, SCLIP_FAST!=0)");
    return SCLIP_ERR_UNSUPPORTED;
  }
  SCLIP_DISPATCH(forward_tiles_kernel, smem, total, max_sms, stream);
}

int launch_backward_tiles(const BwdParams& p, int cg, int ew, cudaStream_t stream) {
  const int smem = tile_smem_bytes(cg, p.stages, staging_slabs(ew, p.store_map_lo[0] >= 0));
  const int total = 3 * (p.nti / cg) * p.ntj;
  SCLIP_DISPATCH(backward_tiles_kernel, smem, total, 0, stream);
}

int launch_gemm(const GemmParams& p, int cg, int ew, int max_sms, cudaStream_t stream) {
  const int smem = tile_smem_bytes(cg, p.stages, 0);
  const int total = p.total_tiles;
  if (total <= 0) return SCLIP_OK;
  SCLIP_DISPATCH(gemm_tiles_kernel, smem, total, max_sms, stream);
}

int wide_stages(int wn) {
  const int st = (227 * 1024 - 2048) / (A_STAGE_BYTES + wn * 64);
  return st > kMaxStages ? kMaxStages : st;
}

int conv_stages() { return kBStages; }

int launch_gemm_conv(const GemmParams& p, int max_sms, cudaStream_t stream) {
  if (p.total_tiles <= 0) return SCLIP_OK;
  if (p.wn != kConvWn) {
    set_error("internal: the converting GEMM takes 384-column tiles (got %d)", p.wn);
    return SCLIP_ERR_ARGUMENT;
  }
  const int smem = kRawStages * (A_STAGE_BYTES + kConvFacBytes) + kBStages * kConvBBytes + 1024;
  return launch_persistent(gemm_conv_kernel, p, 2, 8 + kConvWarps + 1, smem, p.total_tiles, max_sms, stream);
}

int launch_gemm_wide(const GemmParams& p, int ew, int max_sms, cudaStream_t stream) {
  if (p.total_tiles <= 0) return SCLIP_OK;
  if (p.wn != 256 && p.wn != 384 && p.wn != 512) {
    set_error("internal: wide GEMM tile width %d", p.wn);
    return SCLIP_ERR_ARGUMENT;
  }
  const int smem = p.stages * (A_STAGE_BYTES + p.wn * 64) + 1024;
  if (ew == 16) return launch_persistent(gemm_wide_kernel<16>, p, 2, 16, smem, p.total_tiles, max_sms, stream);
  return launch_persistent(gemm_wide_kernel<8>, p, 2, 8, smem, p.total_tiles, max_sms, stream);
}

}  // namespace sclip
