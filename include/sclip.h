/*
 * sclip.h -- C ABI of the B200 (sm_100a) tri-modal contrastive objective.
 *
 * Drop-in boundary for the loss tail of Synergy-CLIP's Tri_CLIP.forward:
 *   reference model.py:247-272  (normalise -> 3 scaled similarity matmuls -> 3 x clip_loss)
 *   reference model.py:52-58    (contrastive_loss / clip_loss)
 *   and the autograd backward of those lines (main_pretraining.py:172-173 calls .backward()).
 * The reference has no native interface of its own (it is pure PyTorch); this header is the
 * interface a maintainer binds with ctypes.  Everything is plain pointers,
 * sizes and a cudaStream_t passed as void*: no torch / ATen types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a negative sclip_status on failure; nothing aborts or throws.
 *     sclip_last_error() returns a thread-local human-readable description of the last failure.
 *   - all device buffers are owned by the caller.  The library allocates no device memory; it only
 *     builds TMA descriptors on the host and passes them as kernel parameters.  `ws` is one caller-allocated blob of sclip_layout.total_bytes
 *     (256-byte aligned) whose sub-buffers are at the byte offsets sclip_plan() reports.
 *   - all work is enqueued on the caller's stream; no call synchronises the host.
 *   - t3 / g3 / loss3 / dt3 are DEVICE pointers to 3 floats ordered (IT, TA, AI)
 *     (logit_scale_for_IT / _TA / _AI, model.py:80-82).
 *   - pair / role table (model.py:255,260,265): IT rows=image cols=text, TA rows=text cols=audio,
 *     AI rows=audio cols=image.  Modalities are ordered (image, text, audio).
 *   - sharding: a rank owns `rows_local` consecutive rows [row_offset, row_offset+rows_local) of the
 *     global batch of `rows_global` samples, for all three modalities (what DistributedSampler gives,
 *     main_pretraining.py:124).  Single GPU: rows_local == rows_global, row_offset == 0.
 */
#ifndef SCLIP_H_
#define SCLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCLIP_ABI_VERSION 2

typedef enum sclip_status {
  SCLIP_OK = 0,
  SCLIP_ERR_ARGUMENT = -1,   /* bad shape / dtype / alignment / null pointer */
  SCLIP_ERR_CUDA = -2,       /* a CUDA runtime or driver call failed */
  SCLIP_ERR_UNSUPPORTED = -3 /* device is not sm_100 or the driver lacks cuTensorMapEncodeTiled */
} sclip_status;

/* dtype of the embeddings handed in and of the embedding gradients handed back */
typedef enum sclip_dtype { SCLIP_F32 = 0, SCLIP_BF16 = 1 } sclip_dtype;

/* arithmetic mode of the tensor-core contractions
 *   SCLIP_MATH_F16   : fp16 operands, fp32 accumulate in TMEM (throughput mode; the bf16 I/O contract, 1e-3)
 *   SCLIP_MATH_F16X3 : every operand split into an fp16 (hi, lo) pair, three tensor-core products per
 *                      contraction (hi*hi + lo*hi + hi*lo), fp32 accumulate (fp32 parity mode, 1e-5)   */
typedef enum sclip_math { SCLIP_MATH_F16 = 0, SCLIP_MATH_F16X3 = 1 } sclip_math;

typedef struct sclip_problem {
  int32_t rows_local;  /* rows of this rank                                   */
  int32_t rows_global; /* global batch B                                      */
  int32_t row_offset;  /* first global row owned by this rank                 */
  int32_t dim;         /* embedding dim D, multiple of 8                      */
  int32_t dtype;       /* sclip_dtype of img/txt/aud and of dimg/dtxt/daud    */
  int32_t math;        /* sclip_math                                          */
  int32_t world;       /* number of ranks sharing the global batch (>= 1)     */
  int32_t parity;      /* 0 | 1: which copy of the exchange buffers (xhat, diag_all) this step uses.  With
                          sclip_push_shards a rank writes its shard straight into the peers' workspaces without
                          waiting for them, so consecutive steps must alternate; otherwise 0 (world == 1: 0)  */
} sclip_problem;

/* Byte offsets into the workspace blob.  Buffers marked [exchange] are the ones the host moves
 * between ranks (NCCL) when world > 1; everything else is private to the library. */
typedef struct sclip_layout {
  uint64_t total_bytes;
  uint64_t xhat;          /* [3][rows_global][dim] fp16 normalised embeddings; this rank's rows are written by
                             sclip_prologue at row_offset [exchange: all-gather of the row shards].  world > 1:
                             two copies; the offset reported here is that of problem->parity               */
  uint64_t xhat_lo;       /* same shape, low halves (SCLIP_MATH_F16X3 only, else == xhat)                    */
  uint64_t inv_norm;      /* [3][rows_local] fp32                                                            */
  uint64_t row_part;      /* [3][col_tiles][2][rows_local] fp32 partial row sums (per 128-column slice)      */
  uint64_t col_part;      /* [3][row_tiles][rows_global] fp32 partial column sums                            */
  uint64_t tile_ref;      /* [3][row_tiles][col_tiles] fp32 exponent reference of each tile                  */
  uint64_t diag;          /* [3][rows_local] fp32 positive-pair logits L_ii                                  */
  uint64_t lse_row;       /* [3][rows_local] fp32                                                            */
  uint64_t lse_col_local; /* [3][rows_global] fp32 column log-sum-exp over this rank's rows [exchange: all-gather] */
  uint64_t lse_col;       /* [3][rows_global] fp32 column log-sum-exp over all rows                          */
  uint64_t row_inv;       /* [3][rows_local] fp32 1 / sum_j exp(L_ij)   (only used while exp(logit_scale) < 64)   */
  uint64_t col_sum_local; /* [3][rows_global] fp32 sum over this rank's rows of exp(L_ij) (same condition)        */
  uint64_t col_inv;       /* [3][rows_global] fp32 1 / sum_i exp(L_ij) over all rows (same condition)             */
  uint64_t loss_part;     /* [3] fp32 this rank's share of the three losses [exchange: all-reduce sum]       */
  uint64_t grad_tiles;    /* [3][rows_local][ld_g] fp16 scaled softmax-gradient strip G'                     */
  uint64_t grad_tiles_lo; /* low halves (SCLIP_MATH_F16X3 only)                                              */
  uint64_t dt_part;       /* [3][row_tiles*col_tiles] fp32                                                   */
  uint64_t dxhat_row;     /* [3][rows_local][dim] fp32 d/dxhat from the row role (+ column role if world==1) */
  uint64_t dxhat_col;     /* [3][rows_global][dim] fp32 column-role partial sums (world > 1 only)
                             [exchange: reduce-scatter sum over ranks]                                        */
  uint64_t col_contrib;   /* [3][rows_local][dim] fp32 landing buffer of that reduce-scatter (world > 1 only)            */
  uint64_t diag_all;      /* [3][rows_global] fp32 positive-pair logits of ALL rows (stash scaling); sclip_forward_diag
                             writes this rank's rows [exchange: all-gather when world > 1]                     */
  uint64_t fac_row;       /* [3][2][rows_local rounded up to 64] fp32 row factors of the stash -> G' conversion,
                             followed by the same values pair-interleaved [3][ld/2][4] (converting GEMM)         */
  uint64_t fac_col;       /* [3][2][rows_global rounded up to 64] fp32 column factors, likewise                  */
  uint64_t dot_part;      /* [3][ceil(rows_local/8)] fp32 partial sums of <xhat, dxhat> (stash mode dlogit_scale) */
  uint64_t status;        /* [4] int32 device-side status words (bit 0: a row/column sum under- or overflowed);
                             cleared by sclip_prologue, read back by sclip_read_status                            */
  uint64_t rowterm_part;  /* [3][ceil(rows_local/64)] fp64 partial sums of the row term of the loss
                             [read by the peers in sclip_forward_loss_peers], followed by
                             [3][ceil(rows_global/1024)] fp64 partial sums of its column term                     */
  uint64_t sync;          /* [64] int32 flags and counters that live across calls (per-source-rank "shard landed"
                             epochs -- written by the peers --, block counters).  THE OWNER ZEROES THIS AREA ONCE when the workspace is
                             allocated; the library never needs it cleared again.                                 */
  int32_t row_tiles;      /* ceil(rows_local / 128)   */
  int32_t col_tiles;      /* ceil(rows_global / 256)  */
  int32_t ld_g;           /* leading dimension (elements) of grad_tiles */
  int32_t reserved;
} sclip_layout;

int sclip_abi_version(void);
const char* sclip_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process so far (all threads). */
long long sclip_kernel_launches(void);

/* Workspace sizing.  Pure host arithmetic; callable without a GPU. */
int sclip_plan(const sclip_problem* problem, sclip_layout* layout);

/* ---- forward (model.py:248-272) -------------------------------------------------------------- */

/* x / x.norm(p=2, dim=-1, keepdim=True) (model.py:248-250, no epsilon) for this rank's rows of the
 * three modalities, rounded to the tensor-core operand format and written at row_offset of `xhat`.
 * flags & SCLIP_PRO_DIAG (SCLIP_MATH_F16, needs t3): the same kernel also writes the positive-pair logits L_ii of
 * this rank's rows into diag_all -- what a forward with SCLIP_FWD_STASH needs; t3 may be NULL otherwise. */
#define SCLIP_PRO_DIAG 1
int sclip_prologue(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                   const float* t3, int flags, void* stream);

/* The three similarity strips (rows_local x rows_global each) with the exp / row-sum / column-sum /
 * diagonal epilogue (model.py:254-265 and the softmax statistics of model.py:52-58); logits are never
 * written to memory.  Needs the complete `xhat` (after the all-gather when world > 1). */
int sclip_forward_tiles(const sclip_problem* problem, void* ws, const float* t3, void* stream);

/* Same, restricted to the pairs in pair_mask (bit 0 IT, bit 1 TA, bit 2 AI) and to the 256-column tiles
 * [col_tile_begin, col_tile_end): lets the host run the tiles whose column operand has already arrived (this rank's
 * own rows first, then one modality after the other) while the all-gather of the rest is still in flight.  Every
 * (pair, column tile) must be covered exactly once before sclip_forward_reduce.
 * flags & SCLIP_FWD_STASH (SCLIP_MATH_F16 only): additionally store E~_ij = exp(L_ij - (L_ii + L_jj)/2) / 16 as fp16
 * tiles into grad_tiles, so that the backward needs no recomputation of the similarities (sclip_backward_scale
 * instead of sclip_backward_tiles).  Needs diag_all (sclip_forward_diag, all-gathered when world > 1). */
#define SCLIP_FWD_STASH 1
/* flags & SCLIP_FWD_WRAP: the column tiles [col_tile_begin, col_tile_end) are taken modulo col_tiles (a range of
 * ranks' columns that wraps around the end of the global batch); col_tile_end - col_tile_begin <= col_tiles. */
#define SCLIP_FWD_WRAP 2
/* flags & SCLIP_FWD_WAIT_PEERS (world > 1, rows_local a multiple of 256, the full column range): ONE launch covers
 * every column; the tiles are taken rank by rank -- this rank's own columns first, then those of rank - 1, rank - 2,
 * ... (the order in which the pushes of the peers land here) -- and the kernel itself waits (acquire loads of the per-rank "landed" flags in `sync`) until the rank that owns
 * a shard has pushed it (sclip_push_shards(..., epoch) on that rank) before touching its columns.  The pushes run
 * concurrently on another stream; give them SMs with max_sms.  `epoch` must be the value every rank passes to
 * sclip_push_shards in this step (a counter that grows by one per forward). */
#define SCLIP_FWD_WAIT_PEERS 4
/* max_sms > 0: the persistent grid takes at most that many SMs, leaving the rest to concurrently running
 * communication kernels (0 = all). */
int sclip_forward_tiles_cols(const sclip_problem* problem, void* ws, const float* t3, int pair_mask,
                             int col_tile_begin, int col_tile_end, int flags, int max_sms, int epoch, void* stream);

/* Positive-pair logits L_ii of this rank's rows for the three pairs, written into diag_all at row_offset. */
int sclip_forward_diag(const sclip_problem* problem, void* ws, const float* t3, void* stream);

/* Merge the per-tile statistics into lse_row, lse_col_local and the partial sums of the loss' row term. */
int sclip_forward_reduce(const sclip_problem* problem, void* ws, void* stream);

/* lse_col = logsumexp over ranks of `col_lse_all` ([world][3][rows_global], NULL => this rank's own
 * lse_col_local), then this rank's share of the three losses:
 *   loss_part[p] = ( sum_i (lse_row_i - L_ii) + sum_i (lse_col_{off+i} - L_ii) ) / (2 * rows_global)
 * written to the workspace and to loss3 (device, 3 floats).  Summed over ranks it is clip_loss. */
int sclip_forward_loss(const sclip_problem* problem, void* ws, const float* col_lse_all, float* loss3,
                       void* stream);

/* Peer-memory variant (see "peer-memory exchanges" below): one kernel reads every rank's lse_col_local and row-term
 * partial sums straight from the peers' workspaces (after a barrier behind sclip_forward_reduce), merges lse_col and
 * computes the COMPLETE three losses -- identical on every rank -- into loss3.  Replaces the all-gather of the column
 * statistics, sclip_forward_loss and the all-reduce of the loss shares. */
int sclip_forward_loss_peers(const sclip_problem* problem, void* ws, const void* const* peer_ws, float* loss3,
                             void* stream);

/* ---- backward (autograd of model.py:248-272) --------------------------------------------------- */

/* Recompute the similarity strips and emit the scaled softmax gradient
 *   G'_p = kappa * c_p * ( (softmax_rows + softmax_cols) / 2 - I ),  c_p = s_p g_p / max_q |s_q g_q|
 * as fp16 tiles, plus the per-tile partial sums of dL/dlogit_scale.  g3: upstream gradients of the three
 * losses (device). */
int sclip_backward_tiles(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream);

/* The same tiles without recomputation, after a forward with SCLIP_FWD_STASH: in place on grad_tiles,
 *   G'_ij = E~_ij (R1_i C1_j + R2_i C2_j) - kappa c_p [i == j]   (= kappa c_p ((softmax_rows + softmax_cols)/2 - I));
 * the identity term is subtracted in fp32 before the rounding to fp16, so the positive-pair entry keeps its
 * precision when the softmax is sharply peaked.  HBM-bound elementwise pass.  When the stash cannot carry the
 * gradient (status word 1, see sclip_read_status: a saturated element, a vanishing loss, a scale >= 44) a recompute
 * launch behind the pass (the kernel of sclip_backward_tiles; it returns at once otherwise) writes G' instead -- decided
 * on the device, no host synchronisation, same results as sclip_backward_tiles.  (The converting GEMMs of
 * SCLIP_GEMM_CONVERT_STASH have no such fallback.) */
int sclip_backward_scale(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream);

/* dxhat_row[m] = G'_{rowpair(m)} . xhat_{col modality}   (rows_local x dim, complete)
 * dxhat_col[m] = G'_{colpair(m)}^T . xhat_{row modality} (rows_global x dim partial sums; world == 1: added
 *                into dxhat_row by the same accumulator instead). */
int sclip_backward_gemms(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream);

/* Same, one role at a time (world > 1): SCLIP_ROLE_COLUMN writes only dxhat_col (so its exchange can start),
 * SCLIP_ROLE_ROW only dxhat_row; SCLIP_ROLE_BOTH == sclip_backward_gemms.  flags: reserved, pass 0. */
#define SCLIP_ROLE_BOTH 0
#define SCLIP_ROLE_COLUMN 1
#define SCLIP_ROLE_ROW 2
/* flags & SCLIP_GEMM_CONVERT_STASH: grad_tiles still holds the forward's stash E~ (no sclip_backward_scale ran):
 * G' = E~ (R1 C1 + R2 C2) - kappa c_p I is formed inside the A-operand path of the GEMM kernel (converted tiles go
 * through tensor memory, never back to HBM or shared memory) from the factor arrays of sclip_backward_factors.  The
 * values are those of sclip_backward_scale + the plain GEMMs, bit for bit.  Only where sclip_gemm_converts_stash()
 * says so (fp16 operands and gradient tiles of 384 columns, e.g. dim 768).  The stash survives the call.
 * EXPERIMENTAL: measured slower than sclip_backward_scale + the plain GEMMs on B200 (12.7-13.3 ms against 10.6 ms at
 * 32768 x 768); the repo's host op does not use it unless SCLIP_CONVERT_IN_GEMM=1.
 * max_sms > 0: at most that many SMs (see sclip_forward_tiles_cols). */
#define SCLIP_GEMM_CONVERT_STASH 1
int sclip_backward_gemms_role(const sclip_problem* problem, void* ws, const float* t3, const float* g3, int role,
                              int flags, int max_sms, void* stream);
int sclip_gemm_converts_stash(const sclip_problem* problem); /* 1 | 0; pure host arithmetic */
/* The per-row / per-column factors of the stash -> G' conversion (fac_row, fac_col) alone: what a converting GEMM
 * needs instead of sclip_backward_scale. */
int sclip_backward_factors(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream);

/* Backward of the normalisation: d x = (d - xhat <xhat, d>) / ||x|| with d = dxhat_row (+ col_contrib, the
 * reduce-scattered column-role gradients [3][rows_local][dim] fp32, NULL when world == 1), times grad_mult
 * (1, or world when the caller's DDP wrapper will average over ranks).  Also finishes dlogit_scale:
 * dt3[p] = grad_mult * (this rank's share), device, 3 floats.
 * dimg/dtxt/daud have the dtype of the problem unless out_f32 != 0.
 * flags & SCLIP_BWD_STASHED: the tiles came from sclip_backward_scale: dlogit_scale is then derived from the row dots
 * <xhat, dxhat> this kernel computes anyway (no tile kernel has summed G' cos). */
#define SCLIP_BWD_STASHED 1
int sclip_backward_finish(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                          const float* t3, const float* g3, const float* col_contrib, float grad_mult, void* dimg,
                          void* dtxt, void* daud, int out_f32, int flags, float* dt3, void* stream);

/* ---- zero-shot scorers (model.py:126-203 get_img_txt_sim_score / get_aud_txt_sim_score, and the return_logits
 * branch model.py:275-277): logits = exp(*log_scale) * unit(a) . unit(b)^T, MATERIALISED (m x n, n = number of
 * prompts / classes).  Same normalise kernel and tile kernel as the training path with a store epilogue.
 * a: (m, dim), b: (n, dim) row-major, dtype SCLIP_F32 | SCLIP_BF16; logits: (m, ldc) fp32 with ldc >= n rounded up to
 * a multiple of 4 (columns [n, ldc) receive zeros); scratch: sclip_cosine_logits_scratch bytes, 256-byte aligned. */
int sclip_cosine_logits_scratch(int m, int n, int dim, int math, uint64_t* bytes);
int sclip_cosine_logits(const void* a, const void* b, const float* log_scale, int m, int n, int dim, int dtype, int math,
                        void* scratch, float* logits, int64_t ldc, void* stream);

/* ---- peer-memory exchanges (world > 1) ------------------------------------------------------------
 * When every rank's workspace lives in symmetric memory (mapped into all processes of the node over NVLink /
 * NVSwitch), the exchanges the reference would do with torch.distributed collectives are kernels of this library that
 * load straight from the peers' workspaces.  peer_ws[r] is the base of rank r's workspace in this process' address
 * space (peer_ws[rank] == ws), world <= SCLIP_MAX_PEERS.  The caller orders the ranks (a signal-pad / flag barrier on
 * the same stream before each call: the sources must be complete and must not be rewritten while peers read). */
#define SCLIP_MAX_PEERS 16

/* Push all-gather: write this rank's normalised operand shard (and its positive-pair logits) into the xhat / diag_all
 * of every other rank, rank + 1 first, then rank + 2, ... -- every rank does the same, so at any moment every rank
 * receives from exactly one peer -- and, as each destination is complete, publish landed[this rank] = epoch in the
 * DESTINATION's `sync` area (system-scope release), which that rank's forward tiles launched with
 * SCLIP_FWD_WAIT_PEERS acquire.  No barrier is needed before the call: the exchange buffers exist twice
 * (problem->parity, alternate it every step), and a peer that is still one step behind reads the other copy.
 * At most max_blocks thread blocks of block_threads (<= 1024) threads, on the SMs the tile kernel leaves free (max_sms).
 * max_blocks == 0: the bytes move on the copy engines instead (strided peer copies enqueued on `stream`, a one-thread
 * kernel behind the copies to each destination publishes its flag): no SMs, the tile kernel may keep all of them. */
int sclip_push_shards(const sclip_problem* problem, void* ws, void* const* peer_ws, int max_blocks, int block_threads,
                      int epoch, void* stream);

/* A one-block kernel that completes once every other rank's shard of this epoch has landed in this workspace: for
 * forward launches that do not wait themselves (SCLIP_FWD_WAIT_PEERS needs shards that are multiples of 256 rows). */
int sclip_wait_shards(const sclip_problem* problem, void* ws, int epoch, void* stream);

/* Pull reduce-scatter: col_contrib[m][i][:] = sum over ranks r (rank order) of rank r's dxhat_col[m][row_offset + i][:]
 * -- the column-role gradients of this rank's rows, ready for sclip_backward_finish. */
int sclip_pull_reduce_cols(const sclip_problem* problem, void* ws, const void* const* peer_ws, int max_blocks,
                           int block_threads /* <= 512 */, void* stream);

/* Copy the four device status words of the workspace to the host (synchronises `stream`).  status_host[0] != 0: a row
 * or column log-sum-exp of the last forward was not finite (non-finite embeddings, or exp(logit_scale) beyond what
 * fp32 can hold); the losses are then non-finite too.  status_host[1] != 0 (complete after sclip_backward_scale): the
 * fp16 stash of this step was not used, because it could not carry the gradient to 1e-3 --
 *   bit 0: an element sits at fp16's saturation value (a negative pair whose logit exceeds the mean of its two
 *          positive-pair logits by more than ln(16 * 65504) = 13.86), found by the conversion pass;
 *   bit 1: one of the losses is below 0.03: the softmax is so peaked that what is left of the gradient sits in elements
 *          the stash holds with few or no bits (its window ends 6.9 nats below the positive pairs).  world == 1 and the
 *          peer-memory loss only: a rank that sees only its share of the loss does not apply this test;
 *   bit 2: a scale exp(t_p) >= 44: a handful of elements carry each row and the fp16 rounding of G' no longer averages
 *          out of dlogit_scale, which the stash route derives from the rounded G' (1.4e-3 .. 2.1e-3 measured at 100).
 * Nothing is wrong in any of these cases: sclip_backward_scale reads the same word on the device and a recompute launch
 * behind the pass (the kernel of sclip_backward_tiles) writes G' instead, at that kernel's cost for the step. */
int sclip_read_status(const sclip_problem* problem, void* ws, int32_t* status_host, void* stream);

/* ---- single-GPU convenience (world == 1): the whole tail in two calls ---------------------------
 * keep_for_backward != 0: sclip_backward will follow on the same workspace.  With SCLIP_MATH_F16 and dim >= 512 the
 * forward then stashes its tiles and the backward converts them (4 bytes of HBM traffic per logit instead of 2 dim
 * flop: the measured crossover); otherwise the backward recomputes the similarities.  0: forward only (evaluation
 * loops).  A stash serves ONE backward: a second sclip_backward on the same forward returns SCLIP_ERR_ARGUMENT. */
int sclip_forward(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                  const float* t3, int keep_for_backward, float* loss3, void* stream);
int sclip_backward(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                   const float* t3, const float* g3, void* dimg, void* dtxt, void* daud, int out_f32, float* dt3,
                   void* stream);

/* ---- building block exposed for tests and for the zero-shot scorers (model.py:126-203) ---------- */
/* C[M x N] (fp32, ld ldc) = alpha * A . B with fp16 operands.
 * a_mn == 0: A is stored [M][K] (K contiguous);  a_mn != 0: A is stored [K][M] (M contiguous).
 * b_mn == 0: B is stored [N][K] (K contiguous);  b_mn != 0: B is stored [K][N] (N contiguous).  */
int sclip_gemm_f16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, float* c, int64_t ldc,
                   int m, int n, int k, float alpha, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCLIP_H_ */
