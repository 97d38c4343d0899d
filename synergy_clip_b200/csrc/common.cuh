// Shared host/device definitions of the sclip library (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/sclip.h"

namespace sclip {

// ------------------------------------------------------------------ tile geometry of the tensor-core kernels
constexpr int BM = 128;      // rows of the accumulator tile (TMEM lanes)
constexpr int BN = 256;      // columns of the accumulator tile (TMEM columns, fp32)
constexpr int BK = 64;       // fp16 elements per k block = one 128-byte swizzle row
constexpr int UMMA_K = 16;   // k extent of one tcgen05.mma kind::f16
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int MN_BOX_BYTES = 64 * BK * 2;    // one 64(mn) x 64(k) box of an MN-major operand

constexpr float kKappa = 32768.0f;           // scale of the fp16 softmax-gradient tiles (|G'| <= kappa)
// While s = exp(logit_scale) < 64 every exp(+-s) and every row / column sum of exp(logit) is a normal fp32 number
// (|logit| <= s because the operands are cosines), so the exponentials need no reference shift at all.  Above that
// (e.g. CLIP's clamp at s = 100) each tile is shifted by its own maximum and the backward uses log-sum-exps.
constexpr float kFastPathMaxScale = 64.0f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kOperandScaleX3 = 256.0f;    // xhat is stored as 256*xhat in the F16X3 mode (keeps lo parts normal)

constexpr int kMaxSegments = 6;
constexpr int kMaxJobs = 6;
constexpr int kFwdMaps = 12;    // tensor maps carried in the kernel parameters (128 B each)
constexpr int kBwdMaps = 18;
constexpr int kGemmMaps = 24;

// One k range of a tile contraction: acc += A_seg(m0.., :) . B_seg(n0.., :)^T
struct Segment {
  int map_a;   // index into the tensor-map table of the kernel parameters
  int map_b;
  int a_mn;    // 0: operand stored [rows][k] (K-major)   1: stored [k][rows] (MN-major)
  int b_mn;
  int num_kb;  // number of BK-wide k blocks
  int pair;    // converting GEMM: the pair whose stash this segment reads (factor lookup); else unused
};

struct Job {
  Segment seg[kMaxSegments];
  int nseg;
  int ksplits;   // >1: the k blocks of every segment are divided over this many CTAs which add into the output
  int m_tiles;   // BM tiles
  int n_tiles;   // BN tiles
  int tile_base; // first linear tile id of this job inside a batched launch
};

// thread-local error text (defined in sclip_api.cu)
void set_error(const char* fmt, ...);

#define SCLIP_CUDA_OK(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::sclip::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SCLIP_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

// every kernel launch of the library is counted (sclip_kernel_launches: the bench reports it as gpu_launches)
void count_launch();
#define SCLIP_LAUNCHED()                     \
  do {                                       \
    ::sclip::count_launch();                 \
    SCLIP_CUDA_OK(cudaGetLastError());       \
  } while (0)

// pair p: rows = modality p, cols = modality (p + 1) % 3  (model.py:255,260,265)
__host__ __device__ inline int pair_row_modality(int p) { return p; }
__host__ __device__ inline int pair_col_modality(int p) { return (p + 1) % 3; }
// modality m takes the row role in pair m and the column role in pair (m + 2) % 3
__host__ __device__ inline int modality_row_pair(int m) { return m; }
__host__ __device__ inline int modality_col_pair(int m) { return (m + 2) % 3; }

// ------------------------------------------------------------------ kernel launchers (defined in the .cu files)
struct Workspace {  // resolved device pointers of one workspace blob
  sclip_problem pb;
  sclip_layout lay;
  uint8_t* base;
  __half* xhat[3];
  __half* xhat_lo[3];
  float* inv_norm;
  float* row_part;
  float* col_part;
  float* tile_ref;
  float* diag;
  float* lse_row;
  float* lse_col_local;
  float* lse_col;
  float* row_inv;
  float* col_sum_local;
  float* col_inv;
  float* loss_part;
  __half* g[3];
  __half* g_lo[3];
  float* dt_part;
  float* dxhat_row;
  float* dxhat_col;
  float* diag_all;
  float* fac_row;
  float* fac_col;
  float* dot_part;
  double* rowterm_part;
  int* status;
  int* sync;  // flags and counters that live across launches (zeroed once by the owner of the workspace)
};

// slots of the `sync` area (32-bit words)
constexpr int kSyncLanded = 0;       // [SCLIP_MAX_PEERS] epoch of the last complete shard per source rank (peers write)
constexpr int kSyncArrived = 16;     // [SCLIP_MAX_PEERS] block counters of sclip_push_shards (per destination)
constexpr int kSyncFinishDone = 32;  // block counter of sclip_backward_finish
constexpr int kSyncLossDone = 33;    // [3] block counters of the loss kernel
constexpr int kSyncWords = 64;

// status words of a workspace (sclip_read_status)
constexpr int kStatusNonFinite = 0;      // a row / column log-sum-exp of the last forward was not finite
constexpr int kStatusStashOverflow = 1;  // non-zero: the backward recomputes G' instead of converting the stash.
                                         //   bit 0 (set by the conversion pass): a stash element sits at fp16's
                                         //          saturation value (a negative pair more than 13.86 nats above the
                                         //          mean of its two positive pairs);
                                         //   bit 1 (forward_loss_kernel): a loss is below kStashMinLoss -- the softmax
                                         //          is so peaked that what is left of the gradient sits in elements the
                                         //          stash keeps with few or no bits (its fp16 window ends 6.9 nats below
                                         //          the positive pairs);
                                         //   bit 2 (backward_factors_kernel): a scale exp(t_p) >= kStashMaxScale.  There a
                                         //          handful of elements carry each row and their fp16 rounding no longer
                                         //          averages out of dlogit_scale, which the stash route derives from the
                                         //          rounded G' (measured 1.4e-3 .. 2.1e-3 at s = 100; 3.9e-4 at 43.5)
constexpr float kStashMinLoss = 0.03f;
constexpr float kStashMaxScale = 44.0f;  // = the limit of the folded-exponent forward (kFoldMaxScale)

struct FwdParams {
  CUtensorMap maps[kFwdMaps];
  Job jobs[3];
  const float* t3;
  float* row_part;
  float* col_part;
  float* tile_ref;
  float* diag;
  int rows_local, rows_global, row_offset;
  int nti, ntj;     // layout strides: 128-row tiles (padded to even) and 256-column tiles
  int tj_begin, tj_count;  // column tiles this launch covers
  int pair_list[3], npairs;  // pairs this launch covers
  int stash;                 // also store E~ = exp(L_ij - (L_ii + L_jj)/2)/16 as fp16 tiles (backward without recompute)
  int store_map[3];          // tensor maps (box 64 x 128) of the stash strips
  const float* diag_all;     // [3][rows_global] positive-pair logits (stash scaling)
  int stages;       // depth of the TMA ring
  int pair_filter;  // forward_tiles_kernel: skip the pairs forward_fast_kernel has taken (s < 44)
  // SCLIP_FWD_WAIT_PEERS: tiles are taken rank by rank (this rank's own columns first, then rank + 1, ...) and the
  // producer acquires landed[source rank] >= epoch before the first tile on a rank's columns
  int wait_peers;
  int tiles_per_rank;     // 256-column tiles per rank
  int world, rank;
  const int* landed;
  int epoch;
  float acc_scale;  // accumulator -> cosine (1 in F16 mode, 2^-16 in F16X3 mode)
};

struct BwdParams {
  CUtensorMap maps[kBwdMaps];
  Job jobs[3];
  int store_map[3];      // tensor map (box 64 x 128) for the G' tile stores, per pair
  int store_map_lo[3];   // low halves (F16X3) or -1
  const float* t3;
  const float* g3;
  const float* lse_row;
  const float* lse_col;
  const float* row_inv;
  const float* col_inv;
  float* dt_part;
  int rows_local, rows_global, row_offset;
  int nti, ntj;
  int stages;
  float acc_scale;
  const int* only_if;  // when non-null the kernel does nothing unless *only_if != 0 (stash overflow fallback)
};

struct GemmParams {
  CUtensorMap maps[kGemmMaps];
  Job jobs[kMaxJobs];
  float* out[kMaxJobs];
  long long ldc[kMaxJobs];
  int m[kMaxJobs];
  int n[kMaxJobs];
  int njobs;
  int total_tiles;
  int stages;
  int wn;            // wide kernel: accumulator columns of one CTA-pair tile (256 | 384 | 512)
  const float* t3;   // when non-null alpha = alpha0 * max_q |exp(t_q) g_q| (backward); else alpha = alpha0
  const float* g3;
  const float* log_alpha;  // when non-null alpha is further multiplied by exp(*log_alpha) (zero-shot scorers)
  float alpha0;
  // converting GEMM (gemm_conv_kernel): the A operand is the forward's stash E~; G' = E~ (R1 C1 + R2 C2) - kappa c_p I is
  // formed on the fly from the factor arrays of sclip_backward_factors
  const float* fac_row;  // [3][2][ld_row]
  const float* fac_col;  // [3][2][ld_col]
  int ld_row, ld_col, row_offset;
};

// cg = 1: one CTA per tile of 128 rows; cg = 2: CTA pairs (cta_group::2) on tiles of 256 rows
// ew = 8 | 16 epilogue warps per CTA
// max_sms > 0: the persistent grid takes at most that many SMs (the rest is left to communication kernels)
int launch_forward_tiles(const FwdParams& p, int cg, int ew, int max_sms, cudaStream_t stream);
int launch_backward_tiles(const BwdParams& p, int cg, int ew, cudaStream_t stream);
int launch_gemm(const GemmParams& p, int cg, int ew, int max_sms, cudaStream_t stream);
// CTA pairs on 256 x wn tiles with one accumulator (MN-major B operands only); see gemm_wide_kernel
int launch_gemm_wide(const GemmParams& p, int ew, int max_sms, cudaStream_t stream);
// the same tiles (wn = 384 only) with the stash -> G' conversion in the A-operand path (TMEM)
int launch_gemm_conv(const GemmParams& p, int max_sms, cudaStream_t stream);
int conv_stages();
int launch_backward_factors(const Workspace& w, const float* t3, const float* g3, cudaStream_t stream);
int wide_stages(int wn);  // depth of the TMA ring that fits beside nothing else in shared memory
int cta_group();   // SCLIP_CTA_GROUP environment override (1 or 2), default 2
int epi_warps();   // SCLIP_EPI_WARPS environment override (8 or 16), default 16
int sm_count();    // SMs of the current device
int staging_slabs(int ew, bool split);  // 16 KiB G' staging slabs the backward tile kernel needs

// t3_for_diag != null: also write this rank's positive-pair logits into diag_all (stash forward)
int launch_prologue(const Workspace& w, const void* const x3[3], const float* t3_for_diag, cudaStream_t stream);
int reduce_row_blocks(const sclip_problem& pb);  // entries per pair of rowterm_part
int loss_col_chunks(const sclip_problem& pb);    // entries per pair of the column-term partial sums behind it
int launch_forward_reduce(const Workspace& w, int row_tiles_done, cudaStream_t stream);
// col_lse_all: the ranks' statistics gathered by a collective; peer_ws: read them from the peers' workspaces (and
// compute the complete losses); both null: world == 1
int launch_forward_loss(const Workspace& w, const float* col_lse_all, const void* const* peer_ws, float* loss3,
                        cudaStream_t stream);
int launch_backward_finish(const Workspace& w, const void* const x3[3], const float* t3, const float* g3,
                           const float* col_contrib, float grad_mult, void* const dx3[3], int out_f32, int stash,
                           float* dt3, cudaStream_t stream);
int launch_diag(const Workspace& w, const float* t3, cudaStream_t stream);
// rows x dim matrix -> unit rows as fp16 operands (hi, and lo when split), times opscale
int launch_normalise(const void* x, int dtype, int rows, int dim, __half* hi, __half* lo, float* inv_norm, bool split,
                     cudaStream_t stream);
// peer-memory exchanges (world > 1, workspaces in symmetric memory); peer_ws[r] = base of rank r's workspace
int launch_push_shards(const Workspace& w, void* const* peer_ws, int max_blocks, int block_threads, int epoch,
                       cudaStream_t stream);
int launch_wait_shards(const Workspace& w, int epoch, cudaStream_t stream);
int launch_pull_reduce(const Workspace& w, const void* const* peer_ws, int max_blocks, int block_threads,
                       cudaStream_t stream);
int launch_backward_scale(const Workspace& w, const float* t3, const float* g3, cudaStream_t stream);

}  // namespace sclip
