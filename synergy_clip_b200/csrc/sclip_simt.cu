// HBM-bound SIMT kernels around the tensor-core tiles: the fused normalise + cast prologue
// (model.py:248-250), the merge of per-tile softmax statistics into log-sum-exps and the three
// losses (model.py:52-58), and the backward of the normalisation plus dL/dlogit_scale.
#include <cstring>

#include "common.cuh"

namespace sclip {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) v += __shfl_xor_sync(0xffffffffu, v, w);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) v += __shfl_xor_sync(0xffffffffu, v, w);
  return v;
}

// 8 consecutive elements of a row as fp32 (16-byte / 32-byte coalesced vector loads)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 raw;
  uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ uint4 pack8_half(const float (&v)[8]) {
  uint4 raw;
  uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return raw;
}

constexpr int kRowsPerBlock = 8;  // one warp per embedding row

// ------------------------------------------------------------------------------------------------ prologue
// x / ||x||_2 (no epsilon: a zero row gives NaN exactly like model.py:248), rounded to fp16 operands.
// F16X3: operand = 256 * xhat split into hi + lo fp16 halves.
struct PrologueArgs {
  const void* x[3];
  __half* hi[3];
  __half* lo[3];
  float* inv_norm;
  int rows, dim;
  int row_offset;
  float opscale;
  int split;
};

template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock * 32) prologue_kernel(const PrologueArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRowsPerBlock + warp;
  const int m = blockIdx.y;
  if (row >= a.rows) return;
  const T* xr = static_cast<const T*>(a.x[m]) + static_cast<size_t>(row) * a.dim;
  float ss = 0.f;
  for (int i = lane * 8; i < a.dim; i += 256) {
    float v[8];
    load8(xr + i, v);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) part = fmaf(v[k], v[k], part);
    ss += part;
  }
  ss = warp_sum_f(ss);
  const float nrm = sqrtf(ss);
  if (lane == 0) a.inv_norm[static_cast<size_t>(m) * a.rows + row] = 1.0f / nrm;
  __half* hr = a.hi[m] + static_cast<size_t>(a.row_offset + row) * a.dim;
  __half* lr = a.lo[m] + static_cast<size_t>(a.row_offset + row) * a.dim;
  for (int i = lane * 8; i < a.dim; i += 256) {
    float v[8];
    load8(xr + i, v);  // second read hits L1/L2
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (v[k] / nrm) * a.opscale;
    *reinterpret_cast<uint4*>(hr + i) = pack8_half(v);
    if (a.split) {
      float r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = v[k] - __half2float(__float2half_rn(v[k]));
      *reinterpret_cast<uint4*>(lr + i) = pack8_half(r);
    }
  }
}

// All three modalities of one row per warp, and -- when a stash forward follows -- the positive-pair logits
// diag_all[p][row_offset + i] = s_p <xr_i, xc_i> of the fp16 operands as the tensor cores will see them.
struct Prologue3Args {
  const void* x[3];
  __half* hi[3];
  __half* lo[3];
  float* inv_norm;
  const float* t3;    // null: no positive-pair logits
  float* diag_all;    // [3][rows_global]
  int rows, dim, rows_global;
  int row_offset;
  float opscale;
  int split;
};

template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock * 32) prologue3_kernel(const Prologue3Args a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRowsPerBlock + warp;
  if (row >= a.rows) return;
  const T* xr[3];
  float ss[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int m = 0; m < 3; ++m) xr[m] = static_cast<const T*>(a.x[m]) + static_cast<size_t>(row) * a.dim;
  for (int i = lane * 8; i < a.dim; i += 256) {
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float v[8];
      load8(xr[m] + i, v);
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) part = fmaf(v[k], v[k], part);
      ss[m] += part;
    }
  }
  float nrm[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    nrm[m] = sqrtf(warp_sum_f(ss[m]));
    if (lane == 0) a.inv_norm[static_cast<size_t>(m) * a.rows + row] = 1.0f / nrm[m];
  }
  float dots[3] = {0.f, 0.f, 0.f};
  const size_t out_off = static_cast<size_t>(a.row_offset + row) * a.dim;
  for (int i = lane * 8; i < a.dim; i += 256) {
    float r[3][8];  // operands rounded to fp16, as the tensor cores see them
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float v[8];
      load8(xr[m] + i, v);  // second read hits L1/L2
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (v[k] / nrm[m]) * a.opscale;
      const uint4 packed = pack8_half(v);
      *reinterpret_cast<uint4*>(a.hi[m] + out_off + i) = packed;
      const __half2* h = reinterpret_cast<const __half2*>(&packed);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        r[m][2 * k] = f.x;
        r[m][2 * k + 1] = f.y;
      }
      if (a.split) {
        float l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) l[k] = v[k] - r[m][k];
        *reinterpret_cast<uint4*>(a.lo[m] + out_off + i) = pack8_half(l);
      }
    }
    if (a.t3 != nullptr) {
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) part = fmaf(r[p][k], r[(p + 1) % 3][k], part);
        dots[p] += part;
      }
    }
  }
  if (a.t3 != nullptr) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const float d = warp_sum_f(dots[p]);
      if (lane == 0) a.diag_all[static_cast<size_t>(p) * a.rows_global + a.row_offset + row] = expf(a.t3[p]) * d;
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward reduce
// Merge of the per-tile statistics: lse_row / row_inv of this rank's rows, lse_col_local / col_sum_local over this
// rank's rows, and per-block partial sums of the row term of the loss.  A block handles 64 rows (or columns); its
// 256 threads split the tiles to merge four ways so that enough loads are in flight (the round-1 kernel, one thread
// per row walking all tiles, ran at 1.7 TB/s), and the four partial sums are combined in a fixed order.
struct ReduceArgs {
  const float* row_part;
  const float* col_part;
  const float* tile_ref;
  const float* diag;
  float* lse_row;
  float* lse_col_local;
  float* row_inv;
  float* col_sum_local;
  double* rowterm_part;  // [3][row_blocks]: sum over the block's rows of (lse_row_i - 2 L_ii)
  int* status;
  int rows_local, rows_global;
  int nti;       // layout stride (128-row tiles, padded)
  int nti_done;  // 128-row tiles the forward kernel actually produced
  int ntj;
  int row_blocks, col_blocks;
};

constexpr int kReduceRows = 64;

__global__ void __launch_bounds__(256) forward_reduce_kernel(const ReduceArgs a) {
  __shared__ float part_sum[4][kReduceRows];
  __shared__ double term[kReduceRows];
  const int r = threadIdx.x & (kReduceRows - 1);
  const int part = threadIdx.x / kReduceRows;  // 0..3
  const int p = blockIdx.y;
  const bool row_side = static_cast<int>(blockIdx.x) < a.row_blocks;
  bool bad = false;
  if (row_side) {
    const int i = blockIdx.x * kReduceRows + r;
    const bool ok = i < a.rows_local;
    const int ii = ok ? i : a.rows_local - 1;
    const int ti = ii / BM;
    const float* ref = a.tile_ref + (static_cast<size_t>(p) * a.nti + ti) * a.ntj;
    float R = -INFINITY;
    for (int tj = 0; tj < a.ntj; ++tj) R = fmaxf(R, __ldg(ref + tj));
    const float* rp0 = a.row_part + static_cast<size_t>(p) * a.ntj * 2 * a.rows_local + ii;
    const size_t pitch = static_cast<size_t>(2) * a.rows_local;  // two slots per column tile (one per 128-column slice)
    auto term_of = [&](int tj) {
      const float* rp = rp0 + tj * pitch;
      return (__ldg(rp) + __ldg(rp + a.rows_local)) * __expf(__ldg(ref + tj) - R);
    };
    const int per = (a.ntj + 3) / 4, lo = part * per, hi = min(a.ntj, lo + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int tj = lo;
    for (; tj + 3 < hi; tj += 4) {
      const float t0 = term_of(tj), t1 = term_of(tj + 1), t2 = term_of(tj + 2), t3 = term_of(tj + 3);
      s0 += t0; s1 += t1; s2 += t2; s3 += t3;
    }
    for (; tj < hi; ++tj) s0 += term_of(tj);
    part_sum[part][r] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (part == 0) {
      const float sum = (part_sum[0][r] + part_sum[1][r]) + (part_sum[2][r] + part_sum[3][r]);
      const float lse = R + logf(sum);
      double t = 0.0;
      if (ok) {
        a.lse_row[static_cast<size_t>(p) * a.rows_local + i] = lse;
        a.row_inv[static_cast<size_t>(p) * a.rows_local + i] = 1.0f / sum;  // meaningful when every tile reference is 0
        bad |= !isfinite(lse);
        t = static_cast<double>(lse) - 2.0 * static_cast<double>(a.diag[static_cast<size_t>(p) * a.rows_local + i]);
      }
      term[r] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int k = 0; k < kReduceRows; ++k) acc += term[k];
      a.rowterm_part[static_cast<size_t>(p) * a.row_blocks + blockIdx.x] = acc;
    }
  } else {
    const int j = (blockIdx.x - a.row_blocks) * kReduceRows + r;
    const bool ok = j < a.rows_global;
    const int jj = ok ? j : a.rows_global - 1;
    const int tj = jj / BN;
    float R = -INFINITY;
    for (int ti = 0; ti < a.nti_done; ++ti) R = fmaxf(R, __ldg(a.tile_ref + (static_cast<size_t>(p) * a.nti + ti) * a.ntj + tj));
    auto term_of = [&](int ti) {
      return __ldg(a.col_part + (static_cast<size_t>(p) * a.nti + ti) * a.rows_global + jj) *
             __expf(__ldg(a.tile_ref + (static_cast<size_t>(p) * a.nti + ti) * a.ntj + tj) - R);
    };
    const int per = (a.nti_done + 3) / 4, lo = part * per, hi = min(a.nti_done, lo + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int ti = lo;
    for (; ti + 3 < hi; ti += 4) {
      const float t0 = term_of(ti), t1 = term_of(ti + 1), t2 = term_of(ti + 2), t3 = term_of(ti + 3);
      s0 += t0; s1 += t1; s2 += t2; s3 += t3;
    }
    for (; ti < hi; ++ti) s0 += term_of(ti);
    part_sum[part][r] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (part == 0 && ok) {
      const float sum = (part_sum[0][r] + part_sum[1][r]) + (part_sum[2][r] + part_sum[3][r]);
      const float lse = R + logf(sum);
      a.lse_col_local[static_cast<size_t>(p) * a.rows_global + j] = lse;
      a.col_sum_local[static_cast<size_t>(p) * a.rows_global + j] = sum;
      bad |= !isfinite(lse);
    }
  }
  if (bad) atomicOr(a.status, 1);
}

// ------------------------------------------------------------------------------------------------ forward loss
// lse_col = log-sum-exp over the ranks' column statistics, the normalisers of the backward, and the losses.
//   share mode (world == 1, or the statistics of the other ranks came through a collective into `col_lse_all`):
//     loss3[p] = this rank's share ( sum_i (lse_row_i - L_ii) + sum_i (lse_col_{off+i} - L_ii) ) / (2 B);
//     summed over ranks it is clip_loss.
//   peer mode (`peer` = every rank's workspace, mapped): the column statistics and the row terms of all ranks are read
//     straight from their workspaces, and every rank computes the complete losses
//     loss3[p] = ( sum_r rowterm_r + sum_j lse_col_j ) / (2 B)   -- same data, same order, same result on every rank.
struct LossArgs {
  const float* lse_col_local;
  const float* col_lse_all;  // [world][3][rows_global] or null
  const uint8_t* peer[SCLIP_MAX_PEERS];  // peer mode: workspace bases in rank order (peer[0] == null: share mode)
  unsigned long long lse_col_local_off, rowterm_off;
  const double* rowterm_part;
  double* colterm_part;  // [3][column chunks] per-block sums of the column term
  unsigned int* done;    // [3] block counters (zero between launches)
  const float* col_sum_local;
  float* lse_col;
  float* col_inv;
  float* loss_part;
  float* loss3;
  int rows_local, rows_global, row_offset, world, row_blocks;
  int* status;  // non-null when `loss` below is a complete loss (world == 1 or peer mode): the peaked-softmax guard
};

__device__ __forceinline__ float ld_relaxed_sys(const float* p) {  // peer data: L2 only, never a stale L1 line
  return __ldcg(p);
}

// grid (column chunks of kLossCols, 3 pairs); one column per thread, so a rank's NVLink loads of all peers' statistics are
// in flight at once (the first version walked 32 columns per thread one after the other: 0.24 ms at 4 ranks).  The
// per-block column terms go to colterm_part and the last block to finish a pair adds everything up in a fixed order.
constexpr int kLossCols = 1024;

__global__ void __launch_bounds__(kLossCols) forward_loss_kernel(const LossArgs a) {
  __shared__ double red[32];
  __shared__ bool last;
  const int p = blockIdx.y;
  const bool peers = a.peer[0] != nullptr;
  const int j = blockIdx.x * kLossCols + threadIdx.x;
  double acc = 0.0;
  if (j < a.rows_global) {
    float v;
    if (!peers && a.col_lse_all == nullptr) {
      v = a.lse_col_local[static_cast<size_t>(p) * a.rows_global + j];
      a.col_inv[static_cast<size_t>(p) * a.rows_global + j] = 1.0f / a.col_sum_local[static_cast<size_t>(p) * a.rows_global + j];
    } else {
      float x[SCLIP_MAX_PEERS];
      float mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < SCLIP_MAX_PEERS; ++w) {
        if (w < a.world) {
          x[w] = peers ? ld_relaxed_sys(reinterpret_cast<const float*>(a.peer[w] + a.lse_col_local_off) +
                                        static_cast<size_t>(p) * a.rows_global + j)
                       : a.col_lse_all[(static_cast<size_t>(w) * 3 + p) * a.rows_global + j];
          mx = fmaxf(mx, x[w]);
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < SCLIP_MAX_PEERS; ++w)
        if (w < a.world) sum += expf(x[w] - mx);
      v = mx + logf(sum);
      a.col_inv[static_cast<size_t>(p) * a.rows_global + j] = expf(-v);
    }
    a.lse_col[static_cast<size_t>(p) * a.rows_global + j] = v;
    // column term: every column in peer mode, the columns of this rank's own rows otherwise
    if (peers || (j >= a.row_offset && j < a.row_offset + a.rows_local)) acc = static_cast<double>(v);
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
    v = warp_sum_d(v);
    if (threadIdx.x == 0) {
      a.colterm_part[static_cast<size_t>(p) * gridDim.x + blockIdx.x] = v;
      __threadfence();
      last = atomicAdd(a.done + p, 1u) == gridDim.x - 1;
      if (last) a.done[p] = 0u;
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block of this pair: column terms of all blocks + row terms sum_i (lse_row_i - 2 L_ii) (per-block partial
  // sums of forward_reduce_kernel; in peer mode those of every rank, read from the peers)
  double tot = 0.0;
  for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += blockDim.x)
    tot += __ldcg(a.colterm_part + static_cast<size_t>(p) * gridDim.x + i);
  if (peers) {
    for (int w = 0; w < a.world; ++w) {
      const double* rt = reinterpret_cast<const double*>(a.peer[w] + a.rowterm_off) + static_cast<size_t>(p) * a.row_blocks;
      for (int i = threadIdx.x; i < a.row_blocks; i += blockDim.x) tot += __ldcg(rt + i);
    }
  } else {
    for (int i = threadIdx.x; i < a.row_blocks; i += blockDim.x)
      tot += a.rowterm_part[static_cast<size_t>(p) * a.row_blocks + i];
  }
  tot = warp_sum_d(tot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
    v = warp_sum_d(v);
    if (threadIdx.x == 0) {
      const float loss = static_cast<float>(v / (2.0 * a.rows_global));
      a.loss_part[p] = loss;
      if (a.loss3 != nullptr) a.loss3[p] = loss;
      if (a.status != nullptr && loss < kStashMinLoss) atomicOr(a.status + kStatusStashOverflow, 2);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward finish
struct FinishArgs {
  const void* x[3];
  void* dx[3];
  const float* dxhat_row;    // [3][rows][dim]
  const float* col_contrib;  // [3][rows][dim] or null
  const float* inv_norm;
  const __half* xhat[3];     // normalised operands at this rank's rows (stash mode: the -I term of G' is applied here)
  const float* t3;
  const float* g3;
  float* dot_part;           // [3][gridDim.x] per-block sums of <xhat, dxhat_total> (stash mode: dlogit_scale)
  const float* dt_part;      // recompute mode: [3][ntiles] tile sums of G' cos
  float* dt3;                // dlogit_scale out (null: not wanted)
  unsigned int* done;        // block counter (zero between launches): the last block to finish reduces dlogit_scale
  int rows, dim, rows_global;
  int ntiles;
  int stash;
  float grad_mult;
};

// d x = (d - xhat <xhat, d>) / ||x||, with xhat recomputed in fp32 from the caller's embeddings.
// Stash mode: no tile kernel has summed G' cos for dlogit_scale, so <xhat_m, dxhat_total_m> = dt_{rowpair(m)} +
// dt_{colpair(m)} is accumulated here (three equations for the three dt, solved by dt_finish_kernel).
template <typename T, typename TO>
__global__ void __launch_bounds__(kRowsPerBlock * 32) backward_finish_kernel(const FinishArgs a) {
  __shared__ float blockdot[kRowsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRowsPerBlock + warp;
  const int m = blockIdx.y;
  float dot = 0.f;
  if (row < a.rows) {
    const size_t base = (static_cast<size_t>(m) * a.rows + row) * a.dim;
    const T* xr = static_cast<const T*>(a.x[m]) + static_cast<size_t>(row) * a.dim;
    TO* outr = static_cast<TO*>(a.dx[m]) + static_cast<size_t>(row) * a.dim;
    const float inv = a.inv_norm[static_cast<size_t>(m) * a.rows + row];
    auto total = [&](int i, float (&d)[8]) {
      load8(a.dxhat_row + base + i, d);
      if (a.col_contrib != nullptr) {
        float e[8];
        load8(a.col_contrib + base + i, e);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] += e[k];
      }
    };
    for (int i = lane * 8; i < a.dim; i += 256) {
      float x[8], d[8];
      load8(xr + i, x);
      total(i, d);
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) part = fmaf(x[k] * inv, d[k], part);
      dot += part;
    }
    dot = warp_sum_f(dot);
    const float scale = inv * a.grad_mult;
    for (int i = lane * 8; i < a.dim; i += 256) {
      float x[8], d[8], o[8];
      load8(xr + i, x);
      total(i, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (d[k] - (x[k] * inv) * dot) * scale;
      store8(outr + i, o);
    }
  }
  if (a.dot_part != nullptr) {
    if (lane == 0) blockdot[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kRowsPerBlock; ++w) sum += blockdot[w];
      a.dot_part[static_cast<size_t>(m) * gridDim.x + blockIdx.x] = sum;
    }
  }
  if (a.dt3 == nullptr) return;
  // dlogit_scale: the last block to arrive sums the partials (fixed order: deterministic) -- no second launch
  __shared__ bool last;
  __shared__ double red[kRowsPerBlock];
  __shared__ double total[3];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int n = gridDim.x * gridDim.y;
    last = atomicAdd(a.done, 1u) == n - 1;
    if (last) *a.done = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int count = a.stash ? static_cast<int>(gridDim.x) : a.ntiles;
  const float* src = a.stash ? a.dot_part : a.dt_part;
  for (int p = 0; p < 3; ++p) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) acc += __ldcg(src + static_cast<size_t>(p) * count + i);
    acc = warp_sum_d(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kRowsPerBlock; ++w) v += red[w];
      total[p] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) {
    const int p = threadIdx.x;
    double dt;
    if (a.stash) {
      // S_m = dt_m + dt_{(m+2)%3}  =>  dt_p = (S_p + S_{(p+1)%3} - S_{(p+2)%3}) / 2
      dt = 0.5 * (total[p] + total[(p + 1) % 3] - total[(p + 2) % 3]);
    } else {
      float mx = 0.f;
      for (int r = 0; r < 3; ++r) mx = fmaxf(mx, fabsf(expf(a.t3[r]) * a.g3[r]));
      dt = total[p] * (static_cast<double>(mx) / (static_cast<double>(kKappa) * a.rows_global));
    }
    a.dt3[p] = static_cast<float>(dt * a.grad_mult);
  }
}

// ------------------------------------------------------------------------------------------------ stash path
// Positive-pair logits of this rank's rows from the fp16 operands: diag_all[p][row_offset + i] = s_p <xr_i, xc_i>
struct DiagArgs {
  const __half* xhat[3];  // at this rank's first row
  const float* t3;
  float* diag_all;        // [3][rows_global]
  int rows, dim, rows_global, row_offset;
};

__global__ void __launch_bounds__(kRowsPerBlock * 32) diag_kernel(const DiagArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRowsPerBlock + warp;
  if (row >= a.rows) return;
  float dots[3] = {0.f, 0.f, 0.f};
  for (int i = lane * 8; i < a.dim; i += 256) {
    float x[3][8];
#pragma unroll
    for (int m = 0; m < 3; ++m) load8(a.xhat[m] + static_cast<size_t>(row) * a.dim + i, x[m]);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) part = fmaf(x[p][k], x[(p + 1) % 3][k], part);
      dots[p] += part;
    }
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const float d = warp_sum_f(dots[p]);
    if (lane == 0) a.diag_all[static_cast<size_t>(p) * a.rows_global + a.row_offset + row] = expf(a.t3[p]) * d;
  }
}

// Per-row and per-column factors of the stash -> G' conversion (see backward_scale_kernel)
struct FactorArgs {
  const float* t3;
  const float* g3;
  const float* diag_all;  // [3][rows_global]
  const float* lse_row;   // [3][rows_local]
  const float* lse_col;   // [3][rows_global]
  float* fac_row;         // [3][2][ld_row]   (ld = rows rounded up to 64, tail zeroed)
  float* fac_col;         // [3][2][ld_col]
  int rows_local, rows_global, row_offset;
  int ld_row, ld_col;
  int* fallback;          // status word kStatusStashOverflow (bit 2: a scale beyond kStashMaxScale)
};

__global__ void __launch_bounds__(256) backward_factors_kernel(const FactorArgs a) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int p = blockIdx.y;
  float mx = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) mx = fmaxf(mx, fabsf(expf(a.t3[r]) * a.g3[r]));
  const float cp = mx > 0.f ? expf(a.t3[p]) * a.g3[p] / mx : 0.f;
  if (i == 0 && a.fallback != nullptr && expf(a.t3[p]) >= kStashMaxScale) atomicOr(a.fallback, 4);
  const float k8 = 8.0f * kKappa * cp;  // (kappa c_p / 2) * 16: the stash carries a 2^-4 headroom factor
  if (i < a.ld_row) {
    float r1 = 0.f, r2 = 0.f;
    if (i < a.rows_local) {
      const float hd = 0.5f * a.diag_all[static_cast<size_t>(p) * a.rows_global + a.row_offset + i];
      r1 = k8 * expf(hd - a.lse_row[static_cast<size_t>(p) * a.rows_local + i]);
      r2 = expf(hd);
    }
    a.fac_row[(static_cast<size_t>(p) * 2 + 0) * a.ld_row + i] = r1;
    a.fac_row[(static_cast<size_t>(p) * 2 + 1) * a.ld_row + i] = r2;
    // the same factors interleaved by pairs of rows -- (R1_i, R1_i+1, R2_i, R2_i+1) -- behind the planar arrays: what
    // the converting GEMM bulk-loads per k block and reads as packed fp32 pairs
    float* pk = a.fac_row + static_cast<size_t>(3) * 2 * a.ld_row + (static_cast<size_t>(p) * (a.ld_row / 2) + i / 2) * 4;
    pk[i & 1] = r1;
    pk[2 + (i & 1)] = r2;
  }
  if (i < a.ld_col) {
    float c1 = 0.f, c2 = 0.f;
    if (i < a.rows_global) {
      const float hd = 0.5f * a.diag_all[static_cast<size_t>(p) * a.rows_global + i];
      c1 = expf(hd);
      c2 = k8 * expf(hd - a.lse_col[static_cast<size_t>(p) * a.rows_global + i]);
    }
    a.fac_col[(static_cast<size_t>(p) * 2 + 0) * a.ld_col + i] = c1;
    a.fac_col[(static_cast<size_t>(p) * 2 + 1) * a.ld_col + i] = c2;
    float* pk = a.fac_col + static_cast<size_t>(3) * 2 * a.ld_col + (static_cast<size_t>(p) * (a.ld_col / 2) + i / 2) * 4;
    pk[i & 1] = c1;
    pk[2 + (i & 1)] = c2;
  }
}

// In place on the fp16 strip: stash E~_ij = exp(L_ij - (L_ii + L_jj)/2) / 16  ->  G'_ij (the -kappa c_p I term included)
//   G'_ij = E~_ij (R1_i C1_j + R2_i C2_j),  R1 = 8 kappa c_p exp(L_ii/2 - lse_row_i), C1 = exp(L_jj/2),
//                                            R2 = exp(L_ii/2),                       C2 = 8 kappa c_p exp(L_jj/2 - lse_col_j)
// HBM-bound: 2 bytes in + 2 bytes out per element; each thread keeps the factors of its 8 columns in registers and
// walks down kScaleRows rows.
constexpr int kScaleRows = 16;
struct ScaleArgs {
  __half* g;             // [3][rows_local][ld]
  const float* fac_row;  // [3][2][ld_row]
  const float* fac_col;  // [3][2][ld_col]
  const float* t3;
  const float* g3;
  int rows_local, rows_global, ld;
  int ld_row, ld_col;
  int row_offset;
  int* fallback;  // status word kStatusStashOverflow: non-zero on entry -> the pass does nothing (the recompute kernel
                  // launched behind it produces G'); the pass itself sets bit 0 when it meets a saturated element
};

// packed fp32 pairs (FFMA2 / FMUL2): half the floating-point instructions of the scalar form -- at the clock the power
// cap leaves after the tile kernels (about 1.2 GHz) the scalar version of this pass was bound by instruction issue, not
// by HBM (5.8 TB/s alone at 1.95 GHz, 4.9 TB/s inside the step)
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long x, unsigned long long y, unsigned long long z) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long x, unsigned long long y) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y));
  return d;
}

// DIAG: the block's 16 rows x 2048 columns contain positive pairs (one block in 2048 at the north-star shape); the
// other blocks run the loop without the per-row test, which keeps four rows of loads in flight.
template <bool DIAG>
__device__ __forceinline__ void scale_rows(__half* ptr, size_t ld, const float* fr1, const float* fr2, int nrows,
                                           const unsigned long long (&c1)[4], const unsigned long long (&c2)[4], int d0,
                                           float kcp, uint32_t& seen_max) {
  constexpr int kBatch = 4;  // rows loaded before the first is stored (the pass is in place: the compiler must not be
                             // left to order a row's load behind the previous row's store)
  for (int r0 = 0; r0 < nrows; r0 += kBatch) {
    uint4 raw[kBatch];
    float r1[kBatch], r2[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int r = min(r0 + j, nrows - 1);
      raw[j] = *reinterpret_cast<const uint4*>(ptr + static_cast<size_t>(r) * ld);
      r1[j] = __ldg(fr1 + r);
      r2[j] = __ldg(fr2 + r);
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int r = r0 + j;
      if (r >= nrows) break;
      const unsigned long long r1p = pk2(r1[j], r1[j]), r2p = pk2(r2[j], r2[j]);
      const uint32_t w[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        {  // largest raw stash value seen by this thread (the stash is non-negative; 0x7BFF = saturated in the forward)
          const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&seen_max), *reinterpret_cast<const __half2*>(&w[k]));
          seen_max = *reinterpret_cast<const uint32_t*>(&m);
        }
        const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        const unsigned long long v = mul2(pk2(e.x, e.y), fma2(r1p, c1[k], mul2(r2p, c2[k])));
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
        if constexpr (DIAG) {  // the identity term, in fp32 before the rounding to fp16
          if (d0 + r == 2 * k) lo -= kcp;
          if (d0 + r == 2 * k + 1) hi -= kcp;
        }
        const __half2 hh = __floats2half2_rn(lo, hi);
        o[k] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(ptr + static_cast<size_t>(r) * ld) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

__global__ void __launch_bounds__(256) backward_scale_kernel(const ScaleArgs a) {
  if (a.fallback != nullptr && (*reinterpret_cast<const volatile int*>(a.fallback) & ~1) != 0) return;
  const int p = blockIdx.z;
  const int col = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (col >= a.rows_global) return;
  unsigned long long c1[4], c2[4];  // column factors of this thread's 8 columns, as pairs
  const float* fc = a.fac_col + static_cast<size_t>(p) * 2 * a.ld_col;  // (padded to 64 and zero-filled: no bound checks)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c1[k] = pk2(fc[col + 2 * k], fc[col + 2 * k + 1]);
    c2[k] = pk2(fc[a.ld_col + col + 2 * k], fc[a.ld_col + col + 2 * k + 1]);
  }
  const float* fr = a.fac_row + static_cast<size_t>(p) * 2 * a.ld_row;
  const int row0 = blockIdx.y * kScaleRows;
  // the "- kappa c_p I" term of G' is applied to the positive-pair entry here, in fp32 before the rounding to fp16:
  // when the softmax is sharply peaked (trained model, large scale) that entry is kappa c_p ((P_row + P_col) / 2 - 1),
  // a small difference of two numbers close to kappa c_p, and subtracting after the rounding (as a separate fp32
  // term behind the fp16 GEMM) loses it: measured 1.7e-3 instead of 3e-4 on the gradients at s = 43.5, cos = 0.25.
  float mxsg = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) mxsg = fmaxf(mxsg, fabsf(expf(a.t3[r]) * a.g3[r]));
  const float kcp = mxsg > 0.f ? kKappa * expf(a.t3[p]) * a.g3[p] / mxsg : 0.f;
  const int nrows = min(kScaleRows, a.rows_local - row0);
  __half* ptr = a.g + (static_cast<size_t>(p) * a.rows_local + row0) * a.ld + col;
  const int d0 = a.row_offset + row0 - col;  // row r holds the positive pair at column offset d0 + r of this thread
  const int g0 = a.row_offset + row0, bc0 = blockIdx.x * 2048;  // block-uniform: does the diagonal cross this block?
  uint32_t seen_max = 0;
  if (g0 + kScaleRows > bc0 && g0 < bc0 + 2048)
    scale_rows<true>(ptr, a.ld, fr + row0, fr + a.ld_row + row0, nrows, c1, c2, d0, kcp, seen_max);
  else
    scale_rows<false>(ptr, a.ld, fr + row0, fr + a.ld_row + row0, nrows, c1, c2, d0, kcp, seen_max);
  if (a.fallback != nullptr && ((seen_max & 0xFFFFu) >= 0x7BFFu || (seen_max >> 16) >= 0x7BFFu)) atomicOr(a.fallback, 1);
}

// ------------------------------------------------------------------------------------------------ peer memory
// world > 1 with the workspaces in symmetric memory (every rank's blob mapped into every process over NVLink /
// NVSwitch): the exchanges are plain kernels that load straight from the peers' workspaces -- a pull all-gather of
// the operand shards, a pull reduce(-scatter) of the column-role gradient partial sums, and gathers of the small
// per-column / per-rank statistics.  One NVSwitch hop per byte; ordering between ranks comes from the host's
// signal-pad barriers on the same stream.
__device__ __forceinline__ uint4 ld_peer(const uint4* p) { return __ldcg(p); }  // L2 only: never a stale L1 line

struct PushShardArgs {
  uint8_t* peer[SCLIP_MAX_PEERS];   // workspace bases of the destination ranks, in push order
  int peer_rank[SCLIP_MAX_PEERS];
  int count;
  const uint8_t* local;
  unsigned long long xhat_off, xhat_lo_off, diag_off, sync_off;
  int nseg;  // 3, or 6 with the low halves
  int rows_local, rows_global, dim, rank;
  unsigned int* arrived;  // [SCLIP_MAX_PEERS] block counters of this launch, per destination (zero between launches)
  int epoch;
};

// This rank's shard (its rows of the three normalised operand matrices, and their positive-pair logits) written into
// every peer's workspace at the same offsets.  Every block walks the destinations in the same order and copies its
// slice to each, so the destinations complete one after the other; the last block to finish a destination publishes
// landed[this rank] = epoch in the DESTINATION's sync area with a system-scope release (the peers' forward tiles,
// launched with SCLIP_FWD_WAIT_PEERS, acquire it).  Local loads, posted NVLink stores.
__global__ void __launch_bounds__(1024) push_shards_kernel(const PushShardArgs a) {
  const size_t seg_units = static_cast<size_t>(a.rows_local) * a.dim * 2 / 16;  // 16-byte units per modality shard
  const size_t total = a.nseg * seg_units;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t row0 = static_cast<size_t>(a.rank) * a.rows_local;
  auto locate = [&](size_t u) {
    const int sg = static_cast<int>(u / seg_units);
    const size_t w = u - sg * seg_units;
    return (sg < 3 ? a.xhat_off : a.xhat_lo_off) + ((static_cast<size_t>(sg % 3) * a.rows_global + row0) * a.dim) * 2 + w * 16;
  };
  for (int pi = 0; pi < a.count; ++pi) {
    uint8_t* dst = a.peer[pi];
    size_t u = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; u + 3 * stride < total; u += 4 * stride) {
      size_t o[4];
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = locate(u + k * stride);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const uint4*>(a.local + o[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + o[k]) = v[k];
    }
    for (; u < total; u += stride) {
      const size_t o = locate(u);
      *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(a.local + o);
    }
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3u * a.rows_local; i += stride) {
      const size_t p = i / a.rows_local, r = i - p * a.rows_local;
      const size_t o = a.diag_off + (p * a.rows_global + row0 + r) * 4;
      *reinterpret_cast<float*>(dst + o) = *reinterpret_cast<const float*>(a.local + o);
    }
    __threadfence_system();  // this thread's stores to the peer are ordered before whatever follows the block barrier
    __syncthreads();
    if (threadIdx.x == 0) {
      const int r = a.peer_rank[pi];
      if (atomicAdd(&a.arrived[r], 1u) == gridDim.x - 1) {
        a.arrived[r] = 0u;
        __threadfence_system();
        int* flag = reinterpret_cast<int*>(dst + a.sync_off) + kSyncLanded + a.rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(a.epoch) : "memory");
      }
    }
  }
}

// Copy-engine variant of the push (sclip_push_shards with max_blocks == 0): the bulk bytes go through strided
// cudaMemcpy2DAsync calls (no SMs, so the similarity tiles keep all of them), and behind the copies to one destination
// this one-thread kernel publishes landed[this rank] = epoch there.  Stream order puts it behind the completed copies.
__global__ void publish_landed_kernel(int* flag, int epoch) {
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

// One small block that returns once every other rank's shard of this epoch has landed (for launches that do not wait
// themselves: ragged shards, the collective-free fallback order).
__global__ void wait_shards_kernel(const int* landed, int world, int rank, int epoch) {
  const int r = threadIdx.x;
  if (r >= world || r == rank) return;
  unsigned long long spins = 0;
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(landed + r) : "memory");
    if (v - epoch >= 0) return;
    __nanosleep(500);
    if (++spins > 8000000ull) {  // ~4 s: a peer never pushed -- fail loudly instead of hanging the stream
      printf("sclip: shard of rank %d never landed (epoch %d)\n", r, epoch);
      __trap();
    }
  }
}

struct PullReduceArgs {
  const uint8_t* peer[SCLIP_MAX_PEERS];  // all ranks in rank order (this rank included)
  int world;
  unsigned long long src_off;   // dxhat_col: [3][rows_global][dim] fp32 in every workspace
  float* out;                   // col_contrib: [3][rows_local][dim] fp32
  int rows_local, rows_global, row_offset, dim;
};

// out[m][i][:] = sum over ranks r (in rank order: deterministic) of dxhat_col_r[m][row_offset + i][:]
template <int W>
__global__ void __launch_bounds__(512) pull_reduce_kernel(const PullReduceArgs a) {
  const size_t per_m = static_cast<size_t>(a.rows_local) * a.dim / 4;  // float4 units per modality
  const size_t total = 3 * per_m;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int world = W > 0 ? W : a.world;
  for (size_t u = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < total; u += stride) {
    const size_t m = u / per_m, w = u - m * per_m;
    const size_t o = a.src_off + ((m * a.rows_global + a.row_offset) * a.dim) * 4 + w * 16;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (W > 0) {
      uint4 v[W];
#pragma unroll
      for (int r = 0; r < W; ++r) v[r] = ld_peer(reinterpret_cast<const uint4*>(a.peer[r] + o));
#pragma unroll
      for (int r = 0; r < W; ++r) {
        acc.x += __uint_as_float(v[r].x);
        acc.y += __uint_as_float(v[r].y);
        acc.z += __uint_as_float(v[r].z);
        acc.w += __uint_as_float(v[r].w);
      }
    } else {
      for (int r = 0; r < world; ++r) {
        const uint4 v = ld_peer(reinterpret_cast<const uint4*>(a.peer[r] + o));
        acc.x += __uint_as_float(v.x);
        acc.y += __uint_as_float(v.y);
        acc.z += __uint_as_float(v.z);
        acc.w += __uint_as_float(v.w);
      }
    }
    reinterpret_cast<float4*>(a.out)[u] = acc;
  }
}

}  // namespace

int launch_prologue(const Workspace& w, const void* const x3[3], const float* t3_for_diag, cudaStream_t stream) {
  Prologue3Args a;
  for (int m = 0; m < 3; ++m) {
    a.x[m] = x3[m];
    a.hi[m] = w.xhat[m];
    a.lo[m] = w.xhat_lo[m];
  }
  a.inv_norm = w.inv_norm;
  a.t3 = t3_for_diag;
  a.diag_all = w.diag_all;
  a.rows = w.pb.rows_local;
  a.dim = w.pb.dim;
  a.rows_global = w.pb.rows_global;
  a.row_offset = w.pb.row_offset;
  a.split = w.pb.math == SCLIP_MATH_F16X3;
  a.opscale = a.split ? kOperandScaleX3 : 1.0f;
  const int grid = (a.rows + kRowsPerBlock - 1) / kRowsPerBlock;
  if (w.pb.dtype == SCLIP_F32)
    prologue3_kernel<float><<<grid, kRowsPerBlock * 32, 0, stream>>>(a);
  else
    prologue3_kernel<__nv_bfloat16><<<grid, kRowsPerBlock * 32, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_normalise(const void* x, int dtype, int rows, int dim, __half* hi, __half* lo, float* inv_norm, bool split,
                     cudaStream_t stream) {
  PrologueArgs a;
  memset(&a, 0, sizeof(a));
  a.x[0] = x;
  a.hi[0] = hi;
  a.lo[0] = lo;
  a.inv_norm = inv_norm;
  a.rows = rows;
  a.dim = dim;
  a.row_offset = 0;
  a.split = split ? 1 : 0;
  a.opscale = split ? kOperandScaleX3 : 1.0f;
  dim3 grid((rows + kRowsPerBlock - 1) / kRowsPerBlock, 1);
  if (dtype == SCLIP_F32)
    prologue_kernel<float><<<grid, kRowsPerBlock * 32, 0, stream>>>(a);
  else
    prologue_kernel<__nv_bfloat16><<<grid, kRowsPerBlock * 32, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int reduce_row_blocks(const sclip_problem& pb) { return (pb.rows_local + kReduceRows - 1) / kReduceRows; }
int loss_col_chunks(const sclip_problem& pb) { return (pb.rows_global + kLossCols - 1) / kLossCols; }

int launch_forward_reduce(const Workspace& w, int row_tiles_done, cudaStream_t stream) {
  const int rb = reduce_row_blocks(w.pb), cb = (w.pb.rows_global + kReduceRows - 1) / kReduceRows;
  ReduceArgs a{w.row_part, w.col_part, w.tile_ref, w.diag, w.lse_row, w.lse_col_local, w.row_inv, w.col_sum_local,
               w.rowterm_part, w.status, w.pb.rows_local, w.pb.rows_global, w.lay.row_tiles, row_tiles_done,
               w.lay.col_tiles, rb, cb};
  forward_reduce_kernel<<<dim3(rb + cb, 3), 256, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_forward_loss(const Workspace& w, const float* col_lse_all, const void* const* peer_ws, float* loss3,
                        cudaStream_t stream) {
  LossArgs a;
  memset(&a, 0, sizeof(a));
  a.lse_col_local = w.lse_col_local;
  a.col_lse_all = col_lse_all;
  if (peer_ws != nullptr)
    for (int r = 0; r < w.pb.world; ++r) a.peer[r] = static_cast<const uint8_t*>(peer_ws[r]);
  a.lse_col_local_off = w.lay.lse_col_local;
  a.rowterm_off = w.lay.rowterm_part;
  a.rowterm_part = w.rowterm_part;
  a.colterm_part = w.rowterm_part + 3 * static_cast<size_t>(reduce_row_blocks(w.pb));
  a.done = reinterpret_cast<unsigned int*>(w.sync) + kSyncLossDone;
  a.col_sum_local = w.col_sum_local;
  a.lse_col = w.lse_col;
  a.col_inv = w.col_inv;
  a.loss_part = w.loss_part;
  a.loss3 = loss3;
  a.rows_local = w.pb.rows_local;
  a.rows_global = w.pb.rows_global;
  a.row_offset = w.pb.row_offset;
  a.world = w.pb.world;
  a.row_blocks = reduce_row_blocks(w.pb);
  a.status = (w.pb.world == 1 || peer_ws != nullptr) ? w.status : nullptr;
  forward_loss_kernel<<<dim3(loss_col_chunks(w.pb), 3), kLossCols, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_backward_finish(const Workspace& w, const void* const x3[3], const float* t3, const float* g3,
                           const float* col_contrib, float grad_mult, void* const dx3[3], int out_f32, int stash,
                           float* dt3, cudaStream_t stream) {
  FinishArgs a;
  const size_t d = w.pb.dim;
  for (int m = 0; m < 3; ++m) {
    a.x[m] = x3[m];
    a.dx[m] = dx3[m];
    a.xhat[m] = w.xhat[m] + static_cast<size_t>(w.pb.row_offset) * d;
  }
  a.dxhat_row = w.dxhat_row;
  a.col_contrib = col_contrib;
  a.inv_norm = w.inv_norm;
  a.t3 = t3;
  a.g3 = g3;
  a.rows = w.pb.rows_local;
  a.dim = w.pb.dim;
  a.rows_global = w.pb.rows_global;
  a.stash = stash;
  a.grad_mult = grad_mult;
  dim3 grid((a.rows + kRowsPerBlock - 1) / kRowsPerBlock, 3);
  a.dot_part = stash ? w.dot_part : nullptr;
  a.dt_part = w.dt_part;
  a.dt3 = dt3;
  a.done = reinterpret_cast<unsigned int*>(w.sync) + kSyncFinishDone;
  a.ntiles = w.lay.row_tiles * w.lay.col_tiles;
  const int threads = kRowsPerBlock * 32;
  if (w.pb.dtype == SCLIP_F32)
    backward_finish_kernel<float, float><<<grid, threads, 0, stream>>>(a);
  else if (out_f32)
    backward_finish_kernel<__nv_bfloat16, float><<<grid, threads, 0, stream>>>(a);
  else
    backward_finish_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, threads, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_diag(const Workspace& w, const float* t3, cudaStream_t stream) {
  DiagArgs a;
  for (int m = 0; m < 3; ++m) a.xhat[m] = w.xhat[m] + static_cast<size_t>(w.pb.row_offset) * w.pb.dim;
  a.t3 = t3;
  a.diag_all = w.diag_all;
  a.rows = w.pb.rows_local;
  a.dim = w.pb.dim;
  a.rows_global = w.pb.rows_global;
  a.row_offset = w.pb.row_offset;
  diag_kernel<<<(a.rows + kRowsPerBlock - 1) / kRowsPerBlock, kRowsPerBlock * 32, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_backward_factors(const Workspace& w, const float* t3, const float* g3, cudaStream_t stream) {
  const int ld_row = (w.pb.rows_local + 63) / 64 * 64, ld_col = (w.pb.rows_global + 63) / 64 * 64;
  FactorArgs f{t3, g3, w.diag_all, w.lse_row, w.lse_col, w.fac_row, w.fac_col,
               w.pb.rows_local, w.pb.rows_global, w.pb.row_offset, ld_row, ld_col, w.status + kStatusStashOverflow};
  const int n = ld_row > ld_col ? ld_row : ld_col;
  backward_factors_kernel<<<dim3((n + 255) / 256, 3), 256, 0, stream>>>(f);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_backward_scale(const Workspace& w, const float* t3, const float* g3, cudaStream_t stream) {
  const int rc = launch_backward_factors(w, t3, g3, stream);
  if (rc) return rc;
  const int ld_row = (w.pb.rows_local + 63) / 64 * 64, ld_col = (w.pb.rows_global + 63) / 64 * 64;
  ScaleArgs a{w.g[0], w.fac_row, w.fac_col, t3, g3, w.pb.rows_local, w.pb.rows_global, w.lay.ld_g, ld_row, ld_col,
              w.pb.row_offset, w.status + kStatusStashOverflow};  // (ScaleArgs::fallback)
  dim3 grid((w.pb.rows_global + 2047) / 2048, (w.pb.rows_local + kScaleRows - 1) / kScaleRows, 3);
  backward_scale_kernel<<<grid, 256, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

}  // namespace sclip

namespace sclip {

int launch_push_shards(const Workspace& w, void* const* peer_ws, int max_blocks, int block_threads, int epoch,
                       cudaStream_t stream) {
  PushShardArgs a;
  memset(&a, 0, sizeof(a));
  const int world = w.pb.world, rank = w.pb.row_offset / w.pb.rows_local;
  for (int i = 0; i + 1 < world; ++i) {
    const int r = (rank + 1 + i) % world;
    a.peer[i] = static_cast<uint8_t*>(peer_ws[r]);
    a.peer_rank[i] = r;
  }
  a.count = world - 1;
  a.local = w.base;
  a.xhat_off = w.lay.xhat;
  a.xhat_lo_off = w.lay.xhat_lo;
  a.diag_off = w.lay.diag_all;
  a.sync_off = w.lay.sync;
  a.nseg = w.pb.math == SCLIP_MATH_F16X3 ? 6 : 3;
  a.rows_local = w.pb.rows_local;
  a.rows_global = w.pb.rows_global;
  a.dim = w.pb.dim;
  a.rank = rank;
  a.arrived = reinterpret_cast<unsigned int*>(w.sync) + kSyncArrived;
  a.epoch = epoch;
  if (max_blocks == 0) {  // copy engines
    const size_t row0 = static_cast<size_t>(rank) * a.rows_local;
    const size_t seg_bytes = static_cast<size_t>(a.rows_local) * a.dim * 2;
    const size_t seg_pitch = static_cast<size_t>(a.rows_global) * a.dim * 2;
    const size_t first = row0 * a.dim * 2;
    for (int pi = 0; pi < a.count; ++pi) {
      uint8_t* dst = a.peer[pi];
      cudaError_t e = cudaMemcpy2DAsync(dst + a.xhat_off + first, seg_pitch, a.local + a.xhat_off + first, seg_pitch,
                                        seg_bytes, 3, cudaMemcpyDeviceToDevice, stream);
      if (e == cudaSuccess && a.nseg == 6)
        e = cudaMemcpy2DAsync(dst + a.xhat_lo_off + first, seg_pitch, a.local + a.xhat_lo_off + first, seg_pitch,
                              seg_bytes, 3, cudaMemcpyDeviceToDevice, stream);
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(dst + a.diag_off + row0 * 4, static_cast<size_t>(a.rows_global) * 4,
                              a.local + a.diag_off + row0 * 4, static_cast<size_t>(a.rows_global) * 4,
                              static_cast<size_t>(a.rows_local) * 4, 3, cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) {
        set_error("sclip_push_shards: peer copy to rank %d failed: %s", a.peer_rank[pi], cudaGetErrorString(e));
        return SCLIP_ERR_CUDA;
      }
      publish_landed_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<int*>(dst + a.sync_off) + kSyncLanded + rank, epoch);
      SCLIP_LAUNCHED();
    }
    return SCLIP_OK;
  }
  const int bx = max_blocks < 1 ? 1 : max_blocks;
  push_shards_kernel<<<bx, block_threads, 0, stream>>>(a);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_wait_shards(const Workspace& w, int epoch, cudaStream_t stream) {
  wait_shards_kernel<<<1, 32, 0, stream>>>(w.sync + kSyncLanded, w.pb.world, w.pb.row_offset / w.pb.rows_local, epoch);
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

int launch_pull_reduce(const Workspace& w, const void* const* peer_ws, int max_blocks, int block_threads,
                       cudaStream_t stream) {
  PullReduceArgs a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < w.pb.world; ++r) a.peer[r] = static_cast<const uint8_t*>(peer_ws[r]);
  a.world = w.pb.world;
  a.src_off = w.lay.dxhat_col;
  a.out = reinterpret_cast<float*>(w.base + w.lay.col_contrib);
  a.rows_local = w.pb.rows_local;
  a.rows_global = w.pb.rows_global;
  a.row_offset = w.pb.row_offset;
  a.dim = w.pb.dim;
  const int blocks = max_blocks < 1 ? 1 : max_blocks;
  switch (w.pb.world) {
    case 2: pull_reduce_kernel<2><<<blocks, block_threads, 0, stream>>>(a); break;
    case 4: pull_reduce_kernel<4><<<blocks, block_threads, 0, stream>>>(a); break;
    case 8: pull_reduce_kernel<8><<<blocks, block_threads, 0, stream>>>(a); break;
    default: pull_reduce_kernel<0><<<blocks, block_threads, 0, stream>>>(a); break;
  }
  SCLIP_LAUNCHED();
  return SCLIP_OK;
}

}  // namespace sclip
