// C ABI of the sclip library (include/sclip.h): argument checking, workspace layout, TMA descriptor
// construction and the stage launchers.  Host code only; the kernels live in sclip_tc.cu / sclip_simt.cu.
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace sclip {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cta_group() {
  static int cached = 0;
  if (cached == 0) {
    const char* e = getenv("SCLIP_CTA_GROUP");
    cached = (e != nullptr && e[0] == '1') ? 1 : 2;
  }
  return cached;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int epi_warps() {
  static int cached = 0;
  if (cached == 0) {
    const char* e = getenv("SCLIP_EPI_WARPS");
    cached = (e != nullptr && e[0] == '1') ? 16 : 8;
  }
  return cached;
}

namespace {

// deepest TMA ring that fits next to `slabs` 16 KiB staging slabs in the 227 KiB of one CTA
int ring_stages(int slabs) {
  const int stage_bytes = cta_group() == 2 ? (A_STAGE_BYTES + B_STAGE_BYTES / 2) : STAGE_BYTES;
  const int budget = 227 * 1024 - 14 * 1024 - 1024 - slabs * 16384;  // static smem + alignment slack
  int st = budget / stage_bytes;
  st = st > 6 ? 6 : (st < 2 ? 2 : st);
  static const int forced = [] {
    const char* e = getenv("SCLIP_STAGES");  // profiling experiments only
    return e != nullptr ? atoi(e) : 0;
  }();
  if (forced >= 1 && forced < st) st = forced;
  return st;
}

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// The gradient GEMMs run on 256 x wn CTA-pair tiles (gemm_wide_kernel) unless SCLIP_WIDE=0 or SCLIP_CTA_GROUP=1.
bool wide_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("SCLIP_WIDE");
    cached = (cta_group() == 2 && !(e != nullptr && e[0] == '0')) ? 1 : 0;
  }
  return cached != 0;
}

// accumulator columns per tile: the fewest tiles of at most 512 columns, rounded to whole 64-column boxes per CTA
int wide_width(int n) {
  const int nt = ceil_div(n, 512);
  const int w = ceil_div(ceil_div(n, nt), 128) * 128;
  return w < 256 ? 256 : w;
}

// k splits per tile so that the persistent clusters finish together: minimise rounds x (k blocks per unit + fixed
// cost of a unit: ring fill and epilogue, about 8 k-block times), keeping at least 32 k blocks per unit.  At most two:
// the partial sums meet in a zeroed output through fp32 red.add, and 0 + a + b does not depend on the arrival order
// (IEEE addition is commutative) whereas three or more addends would make the result run-to-run non-deterministic.
int wide_ksplits(int tiles, int kb, int max_sms) {
  int sms = sm_count();
  if (max_sms > 0 && max_sms < sms) sms = max_sms;
  const int clusters = sms / 2 > 0 ? sms / 2 : 1;
  static const int forced = [] {
    const char* e = getenv("SCLIP_KSPLITS");  // profiling experiments only
    return e != nullptr ? atoi(e) : 0;
  }();
  if (forced >= 1) return forced > kb ? kb : forced;
  int best = 1;
  double best_cost = 0;
  for (int ks = 1; ks <= 2; ++ks) {
    if (ks > 1 && kb / ks < 32) break;
    const double cost = static_cast<double>(ceil_div(tiles * ks, clusters)) * (ceil_div(kb, ks) + 8);
    if (ks == 1 || cost < best_cost * 0.995) {
      best = ks;
      best_cost = cost;
    }
  }
  return best;
}

int check_problem(const sclip_problem* pb) {
  if (pb == nullptr) {
    set_error("problem is null");
    return SCLIP_ERR_ARGUMENT;
  }
  if (pb->rows_local < 1 || pb->rows_global < pb->rows_local || pb->row_offset < 0 ||
      pb->row_offset + pb->rows_local > pb->rows_global) {
    set_error("bad row partition: rows_local=%d rows_global=%d row_offset=%d", pb->rows_local, pb->rows_global,
              pb->row_offset);
    return SCLIP_ERR_ARGUMENT;
  }
  if (pb->dim < 8 || pb->dim % 8 != 0) {
    set_error("dim=%d must be a positive multiple of 8 (16-byte rows for TMA)", pb->dim);
    return SCLIP_ERR_ARGUMENT;
  }
  if (pb->dtype != SCLIP_F32 && pb->dtype != SCLIP_BF16) {
    set_error("dtype=%d is neither SCLIP_F32 nor SCLIP_BF16", pb->dtype);
    return SCLIP_ERR_ARGUMENT;
  }
  if (pb->math != SCLIP_MATH_F16 && pb->math != SCLIP_MATH_F16X3) {
    set_error("math=%d is neither SCLIP_MATH_F16 nor SCLIP_MATH_F16X3", pb->math);
    return SCLIP_ERR_ARGUMENT;
  }
  if (pb->parity != 0 && (pb->parity != 1 || pb->world == 1)) {
    set_error("parity=%d must be 0 or 1 (and 0 when world == 1)", pb->parity);
    return SCLIP_ERR_ARGUMENT;
  }
  if (pb->world < 1 || (pb->world == 1 && pb->rows_local != pb->rows_global)) {
    set_error("world=%d inconsistent with rows_local=%d rows_global=%d", pb->world, pb->rows_local, pb->rows_global);
    return SCLIP_ERR_ARGUMENT;
  }
  return SCLIP_OK;
}

int plan(const sclip_problem* pb, sclip_layout* lay) {
  int rc = check_problem(pb);
  if (rc) return rc;
  if (lay == nullptr) {
    set_error("layout is null");
    return SCLIP_ERR_ARGUMENT;
  }
  memset(lay, 0, sizeof(*lay));
  const uint64_t bl = pb->rows_local, bg = pb->rows_global, d = pb->dim;
  const bool x3 = pb->math == SCLIP_MATH_F16X3;
  lay->row_tiles = 2 * ceil_div(pb->rows_local, 2 * BM);  // even: a CTA pair works on 256 rows
  lay->col_tiles = ceil_div(pb->rows_global, BN);
  lay->ld_g = static_cast<int32_t>(align_up(bg, 64));
  const uint64_t nti = lay->row_tiles, ntj = lay->col_tiles, ldg = lay->ld_g;
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) {
    const uint64_t at = off;
    off = align_up(off + bytes, 256);
    return at;
  };
  // world > 1: the exchange buffers exist twice (problem->parity picks the copy, see sclip_push_shards)
  const uint64_t copies = pb->world > 1 ? 2 : 1, par = pb->world > 1 ? static_cast<uint64_t>(pb->parity) : 0;
  lay->xhat = take(copies * 3 * bg * d * 2) + par * 3 * bg * d * 2;
  lay->xhat_lo = x3 ? take(copies * 3 * bg * d * 2) + par * 3 * bg * d * 2 : lay->xhat;
  lay->inv_norm = take(3 * bl * 4);
  lay->row_part = take(3 * 2 * ntj * bl * 4);
  lay->col_part = take(3 * nti * bg * 4);
  lay->tile_ref = take(3 * nti * ntj * 4);
  lay->diag = take(3 * bl * 4);
  lay->lse_row = take(3 * bl * 4);
  lay->lse_col_local = take(3 * bg * 4);
  lay->lse_col = take(3 * bg * 4);
  lay->row_inv = take(3 * bl * 4);
  lay->col_sum_local = take(3 * bg * 4);
  lay->col_inv = take(3 * bg * 4);
  lay->loss_part = take(3 * 4);
  lay->grad_tiles = take(3 * bl * ldg * 2);
  lay->grad_tiles_lo = x3 ? take(3 * bl * ldg * 2) : lay->grad_tiles;
  lay->dt_part = take(3 * nti * ntj * 4);
  lay->dxhat_row = take(3 * bl * d * 4);
  lay->dxhat_col = pb->world > 1 ? take(3 * bg * d * 4) : lay->dxhat_row;
  lay->col_contrib = pb->world > 1 ? take(3 * bl * d * 4) : lay->dxhat_row;
  lay->diag_all = take(copies * 3 * bg * 4) + par * 3 * bg * 4;
  lay->fac_row = take(2 * 3 * 2 * align_up(bl, 64) * 4);  // planar [3][2][ld] + pair-interleaved [3][ld/2][4]
  lay->fac_col = take(2 * 3 * 2 * align_up(bg, 64) * 4);
  lay->dot_part = take(3 * ((bl + 7) / 8) * 4);
  lay->status = take(4 * 4);
  lay->rowterm_part = take(3 * static_cast<uint64_t>(reduce_row_blocks(*pb) + loss_col_chunks(*pb)) * 8);
  lay->sync = take(kSyncWords * 4);
  lay->total_bytes = off;
  return SCLIP_OK;
}

int resolve(const sclip_problem* pb, void* ws, Workspace* w) {
  int rc = plan(pb, &w->lay);
  if (rc) return rc;
  if (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 255u) != 0) {
    set_error("workspace pointer must be non-null and 256-byte aligned");
    return SCLIP_ERR_ARGUMENT;
  }
  w->pb = *pb;
  uint8_t* b = static_cast<uint8_t*>(ws);
  const sclip_layout& l = w->lay;
  const size_t bl = pb->rows_local, bg = pb->rows_global, d = pb->dim, ldg = l.ld_g;
  w->base = b;
  for (int m = 0; m < 3; ++m) {
    w->xhat[m] = reinterpret_cast<__half*>(b + l.xhat) + m * bg * d;
    w->xhat_lo[m] = reinterpret_cast<__half*>(b + l.xhat_lo) + m * bg * d;
    w->g[m] = reinterpret_cast<__half*>(b + l.grad_tiles) + m * bl * ldg;
    w->g_lo[m] = reinterpret_cast<__half*>(b + l.grad_tiles_lo) + m * bl * ldg;
  }
  w->inv_norm = reinterpret_cast<float*>(b + l.inv_norm);
  w->row_part = reinterpret_cast<float*>(b + l.row_part);
  w->col_part = reinterpret_cast<float*>(b + l.col_part);
  w->tile_ref = reinterpret_cast<float*>(b + l.tile_ref);
  w->diag = reinterpret_cast<float*>(b + l.diag);
  w->lse_row = reinterpret_cast<float*>(b + l.lse_row);
  w->lse_col_local = reinterpret_cast<float*>(b + l.lse_col_local);
  w->lse_col = reinterpret_cast<float*>(b + l.lse_col);
  w->row_inv = reinterpret_cast<float*>(b + l.row_inv);
  w->col_sum_local = reinterpret_cast<float*>(b + l.col_sum_local);
  w->col_inv = reinterpret_cast<float*>(b + l.col_inv);
  w->loss_part = reinterpret_cast<float*>(b + l.loss_part);
  w->dt_part = reinterpret_cast<float*>(b + l.dt_part);
  w->dxhat_row = reinterpret_cast<float*>(b + l.dxhat_row);
  w->dxhat_col = reinterpret_cast<float*>(b + l.dxhat_col);
  w->diag_all = reinterpret_cast<float*>(b + l.diag_all);
  w->fac_row = reinterpret_cast<float*>(b + l.fac_row);
  w->fac_col = reinterpret_cast<float*>(b + l.fac_col);
  w->dot_part = reinterpret_cast<float*>(b + l.dot_part);
  w->status = reinterpret_cast<int*>(b + l.status);
  w->rowterm_part = reinterpret_cast<double*>(b + l.rowterm_part);
  w->sync = reinterpret_cast<int*>(b + l.sync);
  return SCLIP_OK;
}

// ------------------------------------------------------------------------------------------------ TMA descriptors
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// fp16 matrix [outer][inner] (inner contiguous, row pitch ld elements); box = box_inner x box_outer, 128-byte swizzle
int make_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
             uint32_t box_outer) {
  auto fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SCLIP_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld * 2) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ptr=%p ld=%llu)", ptr,
              static_cast<unsigned long long>(ld));
    return SCLIP_ERR_ARGUMENT;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    return SCLIP_ERR_CUDA;
  }
  return SCLIP_OK;
}

// tensor-map slots of a workspace
enum MapSlot {
  kXhatRowsK = 0,     // + m : xhat[m] local rows, K-major A box 64 x 128
  kXhatColsK = 3,     // + m : xhat[m] all rows,   K-major B box 64 x 256
  kXloRowsK = 6,
  kXloColsK = 9,
  kGK = 12,           // + p : G'[p] [rows_local][rows_global], box 64 x 128 (store target and K-major A operand)
  kGloK = 15,
  kGMN = 18,          // + p : same memory, box 64 x 64 (MN-major A operand = G'^T)
  kGloMN = 21,
  kXhatAllMN = 24,    // + m : xhat[m] all rows, box 64(d) x 64(rows): MN-major B operand, k = global row
  kXhatLocMN = 27,    // + m : xhat[m] local rows, MN-major B operand, k = local row
  kXloAllMN = 30,
  kXloLocMN = 33,
  kNumSlots = 36
};

// Encode the descriptor of one slot.
int encode_slot(const Workspace& w, int slot, CUtensorMap* out) {
  const sclip_problem& pb = w.pb;
  const uint64_t bl = pb.rows_local, bg = pb.rows_global, d = pb.dim, ldg = w.lay.ld_g;
  const int m = slot % 3;
  const __half* loc = w.xhat[m] + static_cast<size_t>(pb.row_offset) * d;
  const __half* loc_lo = w.xhat_lo[m] + static_cast<size_t>(pb.row_offset) * d;
  switch (slot - m) {
    case kXhatRowsK: return make_map(out, loc, d, bl, d, BK, BM);
    case kXhatColsK: return make_map(out, w.xhat[m], d, bg, d, BK, BN / cta_group());
    case kXloRowsK: return make_map(out, loc_lo, d, bl, d, BK, BM);
    case kXloColsK: return make_map(out, w.xhat_lo[m], d, bg, d, BK, BN / cta_group());
    case kGK: return make_map(out, w.g[m], bg, bl, ldg, 64, BM);
    case kGloK: return make_map(out, w.g_lo[m], bg, bl, ldg, 64, BM);
    case kGMN: return make_map(out, w.g[m], bg, bl, ldg, 64, BK);
    case kGloMN: return make_map(out, w.g_lo[m], bg, bl, ldg, 64, BK);
    case kXhatAllMN: return make_map(out, w.xhat[m], d, bg, d, 64, BK);
    case kXhatLocMN: return make_map(out, loc, d, bl, d, 64, BK);
    case kXloAllMN: return make_map(out, w.xhat_lo[m], d, bg, d, 64, BK);
    case kXloLocMN: return make_map(out, loc_lo, d, bl, d, 64, BK);
  }
  set_error("internal: unknown tensor-map slot %d", slot);
  return SCLIP_ERR_ARGUMENT;
}

// Descriptors are cached per (workspace, problem): encoding one costs a driver call (~1-2 us) and a stage launch needs
// 6 to 12 of them, which showed in the host time of the 8-GPU step (about 30 launches in 2.7 ms).  A small
// most-recently-used table under a mutex; an entry is revalidated by comparing the whole key.
struct MapCacheEntry {
  const void* ws = nullptr;
  int dev = -1;
  sclip_problem pb{};
  int cg = 0;
  unsigned long long valid = 0;  // bit per slot
  CUtensorMap maps[kNumSlots];
  unsigned long long stamp = 0;
};
constexpr int kMapCacheEntries = 16;
MapCacheEntry g_map_cache[kMapCacheEntries];
unsigned long long g_map_clock = 0;
std::mutex g_map_mutex;

int cached_slot(const Workspace& w, int slot, CUtensorMap* out) {
  int dev = 0;
  SCLIP_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_map_mutex);
  MapCacheEntry* hit = nullptr;
  MapCacheEntry* victim = &g_map_cache[0];
  for (MapCacheEntry& e : g_map_cache) {
    if (e.ws == w.base && e.dev == dev && e.cg == cta_group() && memcmp(&e.pb, &w.pb, sizeof(sclip_problem)) == 0) {
      hit = &e;
      break;
    }
    if (e.stamp < victim->stamp) victim = &e;
  }
  if (hit == nullptr) {
    hit = victim;
    hit->ws = w.base;
    hit->dev = dev;
    hit->pb = w.pb;
    hit->cg = cta_group();
    hit->valid = 0;
  }
  hit->stamp = ++g_map_clock;
  if (!(hit->valid >> slot & 1ull)) {
    const int rc = encode_slot(w, slot, &hit->maps[slot]);
    if (rc) return rc;
    hit->valid |= 1ull << slot;
  }
  *out = hit->maps[slot];
  return SCLIP_OK;
}

// Collects the descriptors one kernel launch needs into the table carried by its parameters.
struct MapTable {
  const Workspace& w;
  CUtensorMap* dst;
  int cap;
  int count = 0;
  int rc = 0;
  int local[kNumSlots];
  MapTable(const Workspace& ws, CUtensorMap* d, int c) : w(ws), dst(d), cap(c) {
    for (int i = 0; i < kNumSlots; ++i) local[i] = -1;
  }
  int use(int slot) {
    if (local[slot] >= 0) return local[slot];
    if (count >= cap) {
      set_error("internal: tensor-map table overflow");
      rc = SCLIP_ERR_ARGUMENT;
      return 0;
    }
    const int r = cached_slot(w, slot, &dst[count]);
    if (r) rc = r;
    local[slot] = count;
    return count++;
  }
};

Segment seg(int a, int b, int a_mn, int b_mn, int num_kb) { return Segment{a, b, a_mn, b_mn, num_kb}; }

// similarity job of pair p (forward and the backward recompute)
void similarity_job(const Workspace& w, MapTable& t, int p, Job* job) {
  const int rm = pair_row_modality(p), cm = pair_col_modality(p);
  const int nkb = ceil_div(w.pb.dim, BK);
  memset(job, 0, sizeof(*job));
  // F16X3: the two small cross products go first so that the tensor core's truncating fp32 accumulation works on
  // small partial sums for two thirds of the k range
  if (w.pb.math == SCLIP_MATH_F16X3) {
    job->seg[job->nseg++] = seg(t.use(kXloRowsK + rm), t.use(kXhatColsK + cm), 0, 0, nkb);
    job->seg[job->nseg++] = seg(t.use(kXhatRowsK + rm), t.use(kXloColsK + cm), 0, 0, nkb);
  }
  job->seg[job->nseg++] = seg(t.use(kXhatRowsK + rm), t.use(kXhatColsK + cm), 0, 0, nkb);
  job->ksplits = 1;
  job->m_tiles = w.lay.row_tiles;
  job->n_tiles = w.lay.col_tiles;
}

}  // namespace
}  // namespace sclip

using namespace sclip;

extern "C" {

int sclip_abi_version(void) { return SCLIP_ABI_VERSION; }
long long sclip_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
const char* sclip_last_error(void) { return g_error; }

int sclip_plan(const sclip_problem* problem, sclip_layout* layout) { return plan(problem, layout); }

int sclip_prologue(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                   const float* t3, int flags, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  const void* x3[3] = {img, txt, aud};
  for (int m = 0; m < 3; ++m)
    if (x3[m] == nullptr || (reinterpret_cast<uintptr_t>(x3[m]) & 15u) != 0) {
      set_error("embedding pointer %d must be non-null and 16-byte aligned", m);
      return SCLIP_ERR_ARGUMENT;
    }
  const bool diag = (flags & SCLIP_PRO_DIAG) != 0;
  if (diag && (t3 == nullptr || w.pb.math != SCLIP_MATH_F16)) {
    set_error("SCLIP_PRO_DIAG needs t3 and a SCLIP_MATH_F16 problem");
    return SCLIP_ERR_ARGUMENT;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SCLIP_CUDA_OK(cudaMemsetAsync(w.status, 0, 16, st));
  return launch_prologue(w, x3, diag ? t3 : nullptr, st);
}

int sclip_forward_tiles(const sclip_problem* problem, void* ws, const float* t3, void* stream) {
  return sclip_forward_tiles_cols(problem, ws, t3, 7, 0, 1 << 30, 0, 0, 0, stream);
}

int sclip_forward_diag(const sclip_problem* problem, void* ws, const float* t3, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (t3 == nullptr) {
    set_error("t3 is null");
    return SCLIP_ERR_ARGUMENT;
  }
  return launch_diag(w, t3, static_cast<cudaStream_t>(stream));
}

static int backward_tiles_impl(const Workspace& w, const float* t3, const float* g3, const int* only_if,
                               cudaStream_t stream);

int sclip_backward_scale(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (t3 == nullptr || g3 == nullptr || w.pb.math != SCLIP_MATH_F16) {
    set_error("sclip_backward_scale needs t3, g3 and a SCLIP_MATH_F16 problem");
    return SCLIP_ERR_ARGUMENT;
  }
  rc = launch_backward_scale(w, t3, g3, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  // a stash element outside fp16's range (status word kStatusStashOverflow, set by the forward): the pass above has
  // skipped itself and this launch recomputes G' from the similarities; otherwise it returns before touching anything
  return backward_tiles_impl(w, t3, g3, w.status + kStatusStashOverflow, static_cast<cudaStream_t>(stream));
}

int sclip_forward_tiles_cols(const sclip_problem* problem, void* ws, const float* t3, int pair_mask,
                             int col_tile_begin, int col_tile_end, int flags, int max_sms, int epoch, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (t3 == nullptr) {
    set_error("t3 is null");
    return SCLIP_ERR_ARGUMENT;
  }
  const bool wait_peers = (flags & SCLIP_FWD_WAIT_PEERS) != 0;
  if (wait_peers) {
    const sclip_problem& q = w.pb;
    if (q.world < 2 || q.world > SCLIP_MAX_PEERS || q.rows_local % 256 != 0 || q.rows_local * q.world != q.rows_global ||
        q.row_offset % q.rows_local != 0 || max_sms < 0) {
      set_error("SCLIP_FWD_WAIT_PEERS needs 2 <= world <= %d equal row shards that are multiples of 256", SCLIP_MAX_PEERS);
      return SCLIP_ERR_ARGUMENT;
    }
    // one launch over every column, starting at this rank's own (the kernel wraps around)
    col_tile_begin = q.row_offset / 256;
    col_tile_end = col_tile_begin + w.lay.col_tiles;
    flags |= SCLIP_FWD_WRAP;
  }
  const bool wrap = (flags & SCLIP_FWD_WRAP) != 0;
  if (!wrap && col_tile_end > w.lay.col_tiles) col_tile_end = w.lay.col_tiles;
  if (col_tile_begin < 0 || col_tile_begin > col_tile_end || col_tile_begin >= w.lay.col_tiles + (wrap ? 0 : 1) ||
      col_tile_end - col_tile_begin > w.lay.col_tiles) {
    set_error("bad column tile range [%d, %d)", col_tile_begin, col_tile_end);
    return SCLIP_ERR_ARGUMENT;
  }
  FwdParams p;
  memset(&p, 0, sizeof(p));
  MapTable tab(w, p.maps, kFwdMaps);
  for (int q = 0; q < 3; ++q) similarity_job(w, tab, q, &p.jobs[q]);
  if (tab.rc) return tab.rc;
  p.t3 = t3;
  p.row_part = w.row_part;
  p.col_part = w.col_part;
  p.tile_ref = w.tile_ref;
  p.diag = w.diag;
  p.rows_local = w.pb.rows_local;
  p.rows_global = w.pb.rows_global;
  p.row_offset = w.pb.row_offset;
  p.nti = w.lay.row_tiles;
  p.ntj = w.lay.col_tiles;
  p.tj_begin = col_tile_begin;
  p.tj_count = col_tile_end - col_tile_begin;
  for (int q = 0; q < 3; ++q)
    if (pair_mask & (1 << q)) p.pair_list[p.npairs++] = q;
  if (flags & SCLIP_FWD_STASH) {
    if (w.pb.math != SCLIP_MATH_F16) {
      set_error("SCLIP_FWD_STASH is only defined for SCLIP_MATH_F16");
      return SCLIP_ERR_ARGUMENT;
    }
    p.stash = 1;
    p.diag_all = w.diag_all;
    for (int q = 0; q < 3; ++q) p.store_map[q] = tab.use(kGK + q);
    if (tab.rc) return tab.rc;
  }
  p.stages = ring_stages(p.stash ? 4 : 0);
  p.acc_scale = w.pb.math == SCLIP_MATH_F16X3 ? 1.0f / (kOperandScaleX3 * kOperandScaleX3) : 1.0f;
  if (wait_peers) {
    p.wait_peers = 1;
    p.tiles_per_rank = w.pb.rows_local / 256;
    p.world = w.pb.world;
    p.rank = w.pb.row_offset / w.pb.rows_local;
    p.landed = w.sync + kSyncLanded;
    p.epoch = epoch;
  }
  return launch_forward_tiles(p, cta_group(), epi_warps(), max_sms, static_cast<cudaStream_t>(stream));
}

int sclip_forward_reduce(const sclip_problem* problem, void* ws, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  return launch_forward_reduce(w, w.lay.row_tiles, static_cast<cudaStream_t>(stream));
}

int sclip_forward_loss(const sclip_problem* problem, void* ws, const float* col_lse_all, float* loss3, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  return launch_forward_loss(w, col_lse_all, nullptr, loss3, static_cast<cudaStream_t>(stream));
}

int sclip_backward_tiles(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (t3 == nullptr || g3 == nullptr) {
    set_error("t3 / g3 is null");
    return SCLIP_ERR_ARGUMENT;
  }
  return backward_tiles_impl(w, t3, g3, nullptr, static_cast<cudaStream_t>(stream));
}

static int backward_tiles_impl(const Workspace& w, const float* t3, const float* g3, const int* only_if,
                               cudaStream_t stream) {
  BwdParams p;
  memset(&p, 0, sizeof(p));
  MapTable tab(w, p.maps, kBwdMaps);
  const bool x3 = w.pb.math == SCLIP_MATH_F16X3;
  for (int q = 0; q < 3; ++q) {
    similarity_job(w, tab, q, &p.jobs[q]);
    p.store_map[q] = tab.use(kGK + q);
    p.store_map_lo[q] = x3 ? tab.use(kGloK + q) : -1;
  }
  if (tab.rc) return tab.rc;
  p.t3 = t3;
  p.g3 = g3;
  p.lse_row = w.lse_row;
  p.lse_col = w.lse_col;
  p.row_inv = w.row_inv;
  p.col_inv = w.col_inv;
  p.dt_part = w.dt_part;
  p.rows_local = w.pb.rows_local;
  p.rows_global = w.pb.rows_global;
  p.row_offset = w.pb.row_offset;
  p.nti = w.lay.row_tiles;
  p.ntj = w.lay.col_tiles;
  p.stages = ring_stages(staging_slabs(epi_warps(), x3));
  p.acc_scale = x3 ? 1.0f / (kOperandScaleX3 * kOperandScaleX3) : 1.0f;
  p.only_if = only_if;
  return launch_backward_tiles(p, cta_group(), epi_warps(), stream);
}

int sclip_backward_gemms(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream) {
  return sclip_backward_gemms_role(problem, ws, t3, g3, SCLIP_ROLE_BOTH, 0, 0, stream);
}

int sclip_gemm_converts_stash(const sclip_problem* problem) {
  if (check_problem(problem)) return 0;
  return (problem->math == SCLIP_MATH_F16 && wide_enabled() && wide_width(problem->dim) == 384) ? 1 : 0;
}

int sclip_backward_factors(const sclip_problem* problem, void* ws, const float* t3, const float* g3, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (t3 == nullptr || g3 == nullptr || w.pb.math != SCLIP_MATH_F16) {
    set_error("sclip_backward_factors needs t3, g3 and a SCLIP_MATH_F16 problem");
    return SCLIP_ERR_ARGUMENT;
  }
  return launch_backward_factors(w, t3, g3, static_cast<cudaStream_t>(stream));
}

int sclip_backward_gemms_role(const sclip_problem* problem, void* ws, const float* t3, const float* g3, int role,
                              int flags, int max_sms, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (t3 == nullptr || g3 == nullptr) {
    set_error("t3 / g3 is null");
    return SCLIP_ERR_ARGUMENT;
  }
  const sclip_problem& pb = w.pb;
  const bool x3 = pb.math == SCLIP_MATH_F16X3;
  const int kb_g = ceil_div(pb.rows_global, BK), kb_l = ceil_div(pb.rows_local, BK);
  const size_t bl = pb.rows_local, bg = pb.rows_global, d = pb.dim;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  MapTable tab(w, p.maps, kGemmMaps);
  p.t3 = t3;
  p.g3 = g3;
  // dXhat = (max_q |s_q g_q| / (kappa B)) G' Xhat ; operands of the F16X3 mode carry an extra factor 256
  p.alpha0 = 1.0f / (kKappa * static_cast<float>(pb.rows_global)) / (x3 ? kOperandScaleX3 : 1.0f);
  // F16X3 (fp32 parity mode): bound the length of one tensor-core accumulation chain to 1024 k elements; the
  // chunks are combined with round-to-nearest fp32 adds (red.global.add.f32 into the zeroed output)
  auto splits_for = [&](int kb) { return x3 ? ceil_div(kb, 16) : 1; };
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (role < SCLIP_ROLE_BOTH || role > SCLIP_ROLE_ROW || (pb.world == 1 && role != SCLIP_ROLE_BOTH)) {
    set_error("role=%d is invalid (single-rank problems only take SCLIP_ROLE_BOTH)", role);
    return SCLIP_ERR_ARGUMENT;
  }
  const bool do_row = role != SCLIP_ROLE_COLUMN, do_col = pb.world > 1 && role != SCLIP_ROLE_ROW;
  if (x3) {
    if (do_row) SCLIP_CUDA_OK(cudaMemsetAsync(w.dxhat_row, 0, 3 * bl * d * 4, st));
    if (do_col) SCLIP_CUDA_OK(cudaMemsetAsync(w.dxhat_col, 0, 3 * bg * d * 4, st));
  }
  int nj = 0, tiles = 0;
  if (max_sms < 0) max_sms = 0;
  const bool convert = (flags & SCLIP_GEMM_CONVERT_STASH) != 0;
  if (convert && !sclip_gemm_converts_stash(problem)) {
    set_error("SCLIP_GEMM_CONVERT_STASH needs SCLIP_MATH_F16 and a dim whose gradient tiles are 384 columns wide "
              "(sclip_gemm_converts_stash)");
    return SCLIP_ERR_ARGUMENT;
  }
  auto add_role = [&](Job& job, int m, bool row_role) {
    if (row_role) {  // G'_{pair m} (rows_local x rows_global, K-major) . xhat_{col modality} (k = global row)
      const int pr = modality_row_pair(m), cm = pair_col_modality(pr);
      if (x3) {
        job.seg[job.nseg++] = seg(tab.use(kGloK + pr), tab.use(kXhatAllMN + cm), 0, 1, kb_g);
        job.seg[job.nseg++] = seg(tab.use(kGK + pr), tab.use(kXloAllMN + cm), 0, 1, kb_g);
      }
      job.seg[job.nseg] = seg(tab.use(kGK + pr), tab.use(kXhatAllMN + cm), 0, 1, kb_g);
      job.seg[job.nseg++].pair = pr;
    } else {  // G'^T_{pair (m+2)%3} (rows_global x rows_local, MN-major view of G') . xhat_{row modality} (k = local row)
      const int pc = modality_col_pair(m), rm = pair_row_modality(pc);
      if (x3) {
        job.seg[job.nseg++] = seg(tab.use(kGloMN + pc), tab.use(kXhatLocMN + rm), 1, 1, kb_l);
        job.seg[job.nseg++] = seg(tab.use(kGMN + pc), tab.use(kXloLocMN + rm), 1, 1, kb_l);
      }
      job.seg[job.nseg] = seg(tab.use(kGMN + pc), tab.use(kXhatLocMN + rm), 1, 1, kb_l);
      job.seg[job.nseg++].pair = pc;
    }
  };
  for (int m = 0; m < 3 && do_row; ++m) {
    Job& job = p.jobs[nj];
    add_role(job, m, true);
    if (pb.world == 1) add_role(job, m, false);
    job.m_tiles = ceil_div(pb.rows_local, BM * cta_group());
    job.n_tiles = ceil_div(pb.dim, BN);
    job.ksplits = splits_for(pb.world == 1 ? (kb_g > kb_l ? kb_g : kb_l) : kb_g);
    job.tile_base = tiles;
    tiles += job.m_tiles * job.n_tiles * job.ksplits;
    p.out[nj] = w.dxhat_row + m * bl * d;
    p.ldc[nj] = pb.dim;
    p.m[nj] = pb.rows_local;
    p.n[nj] = pb.dim;
    ++nj;
  }
  if (do_col) {
    for (int m = 0; m < 3; ++m) {
      Job& job = p.jobs[nj];
      add_role(job, m, false);
      job.m_tiles = ceil_div(pb.rows_global, BM * cta_group());
      job.n_tiles = ceil_div(pb.dim, BN);
      job.ksplits = splits_for(kb_l);
      job.tile_base = tiles;
      tiles += job.m_tiles * job.n_tiles * job.ksplits;
      p.out[nj] = w.dxhat_col + m * bg * d;
      p.ldc[nj] = pb.dim;
      p.m[nj] = pb.rows_global;
      p.n[nj] = pb.dim;
      ++nj;
    }
  }
  if (tab.rc) return tab.rc;
  p.njobs = nj;
  if (wide_enabled() && !x3) {
    // 256 x wn tiles, k split so that the clusters finish together; split partial sums are added into zeroed outputs
    p.wn = wide_width(pb.dim);
    int kb_max = 0, base_tiles = 0;
    for (int j = 0; j < nj; ++j) {
      Job& job = p.jobs[j];
      job.n_tiles = ceil_div(pb.dim, p.wn);
      int kb = 0;
      for (int s = 0; s < job.nseg; ++s) kb += job.seg[s].num_kb;
      kb_max = kb > kb_max ? kb : kb_max;
      base_tiles += job.m_tiles * job.n_tiles;
    }
    const int ks = wide_ksplits(base_tiles, kb_max, max_sms);
    tiles = 0;
    for (int j = 0; j < nj; ++j) {
      Job& job = p.jobs[j];
      job.ksplits = ks;
      job.tile_base = tiles;
      tiles += job.m_tiles * job.n_tiles * ks;
    }
    if (ks > 1) {
      if (do_row) SCLIP_CUDA_OK(cudaMemsetAsync(w.dxhat_row, 0, 3 * bl * d * 4, st));
      if (do_col) SCLIP_CUDA_OK(cudaMemsetAsync(w.dxhat_col, 0, 3 * bg * d * 4, st));
    }
    p.total_tiles = tiles;
    if (convert) {  // the A operand is the forward's stash: G' is formed in the A-operand path (gemm_conv_kernel)
      p.fac_row = w.fac_row;
      p.fac_col = w.fac_col;
      p.ld_row = static_cast<int>(align_up(pb.rows_local, 64));
      p.ld_col = static_cast<int>(align_up(pb.rows_global, 64));
      p.row_offset = pb.row_offset;
      p.stages = conv_stages();
      return launch_gemm_conv(p, max_sms, st);
    }
    p.stages = wide_stages(p.wn);
    return launch_gemm_wide(p, epi_warps(), max_sms, st);
  }
  p.total_tiles = tiles;
  p.stages = ring_stages(0);
  return launch_gemm(p, cta_group(), epi_warps(), max_sms, st);
}

int sclip_backward_finish(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                          const float* t3, const float* g3, const float* col_contrib, float grad_mult, void* dimg,
                          void* dtxt, void* daud, int out_f32, int flags, float* dt3, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  const void* x3[3] = {img, txt, aud};
  void* dx3[3] = {dimg, dtxt, daud};
  for (int m = 0; m < 3; ++m)
    if (x3[m] == nullptr || dx3[m] == nullptr || (reinterpret_cast<uintptr_t>(dx3[m]) & 15u) != 0) {
      set_error("embedding / gradient pointer %d must be non-null and 16-byte aligned", m);
      return SCLIP_ERR_ARGUMENT;
    }
  if (w.pb.world > 1 && col_contrib == nullptr) {
    set_error("world > 1 needs the reduce-scattered column-role gradients (col_contrib)");
    return SCLIP_ERR_ARGUMENT;
  }
  if (t3 == nullptr || g3 == nullptr) {
    set_error("t3 / g3 is null");
    return SCLIP_ERR_ARGUMENT;
  }
  return launch_backward_finish(w, x3, t3, g3, col_contrib, grad_mult, dx3, out_f32, (flags & SCLIP_BWD_STASHED) ? 1 : 0,
                                dt3, static_cast<cudaStream_t>(stream));
}

int sclip_cosine_logits_scratch(int m, int n, int dim, int math, uint64_t* bytes) {
  if (m < 1 || n < 1 || dim < 8 || dim % 8 != 0 || bytes == nullptr ||
      (math != SCLIP_MATH_F16 && math != SCLIP_MATH_F16X3)) {
    set_error("sclip_cosine_logits_scratch: bad argument (m=%d n=%d dim=%d math=%d)", m, n, dim, math);
    return SCLIP_ERR_ARGUMENT;
  }
  const uint64_t parts = math == SCLIP_MATH_F16X3 ? 2 : 1;
  *bytes = parts * (align_up(static_cast<uint64_t>(m) * dim * 2, 256) + align_up(static_cast<uint64_t>(n) * dim * 2, 256)) +
           align_up(static_cast<uint64_t>(m + n) * 4, 256);
  return SCLIP_OK;
}

int sclip_cosine_logits(const void* a, const void* b, const float* log_scale, int m, int n, int dim, int dtype, int math,
                        void* scratch, float* logits, int64_t ldc, void* stream) {
  uint64_t need = 0;
  int rc = sclip_cosine_logits_scratch(m, n, dim, math, &need);
  if (rc) return rc;
  const int n4 = (n + 3) / 4 * 4;
  if (a == nullptr || b == nullptr || log_scale == nullptr || scratch == nullptr || logits == nullptr ||
      (dtype != SCLIP_F32 && dtype != SCLIP_BF16) || ldc < n4 || ldc % 4 != 0 ||
      (reinterpret_cast<uintptr_t>(scratch) & 255u) != 0 || (reinterpret_cast<uintptr_t>(logits) & 15u) != 0 ||
      (reinterpret_cast<uintptr_t>(a) & 15u) != 0 || (reinterpret_cast<uintptr_t>(b) & 15u) != 0) {
    set_error("sclip_cosine_logits: bad argument (null / misaligned pointer, dtype, or ldc < n rounded up to 4)");
    return SCLIP_ERR_ARGUMENT;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool x3 = math == SCLIP_MATH_F16X3;
  uint8_t* base = static_cast<uint8_t*>(scratch);
  const uint64_t abytes = align_up(static_cast<uint64_t>(m) * dim * 2, 256);
  const uint64_t bbytes = align_up(static_cast<uint64_t>(n) * dim * 2, 256);
  __half* a_hi = reinterpret_cast<__half*>(base);
  __half* b_hi = reinterpret_cast<__half*>(base + abytes);
  __half* a_lo = x3 ? reinterpret_cast<__half*>(base + abytes + bbytes) : a_hi;
  __half* b_lo = x3 ? reinterpret_cast<__half*>(base + 2 * abytes + bbytes) : b_hi;
  float* inv = reinterpret_cast<float*>(base + (x3 ? 2 : 1) * (abytes + bbytes));
  rc = launch_normalise(a, dtype, m, dim, a_hi, a_lo, inv, x3, st);
  if (!rc) rc = launch_normalise(b, dtype, n, dim, b_hi, b_lo, inv + m, x3, st);
  if (rc) return rc;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int cg = cta_group();
  rc = make_map(&p.maps[0], a_hi, dim, m, dim, BK, BM);
  if (!rc) rc = make_map(&p.maps[1], b_hi, dim, n, dim, BK, BN / cg);
  if (!rc && x3) rc = make_map(&p.maps[2], a_lo, dim, m, dim, BK, BM);
  if (!rc && x3) rc = make_map(&p.maps[3], b_lo, dim, n, dim, BK, BN / cg);
  if (rc) return rc;
  Job& job = p.jobs[0];
  const int nkb = ceil_div(dim, BK);
  if (x3) {  // small cross terms first (see similarity_job)
    job.seg[job.nseg++] = seg(2, 1, 0, 0, nkb);
    job.seg[job.nseg++] = seg(0, 3, 0, 0, nkb);
  }
  job.seg[job.nseg++] = seg(0, 1, 0, 0, nkb);
  job.ksplits = 1;
  job.m_tiles = ceil_div(m, BM * cg);
  job.n_tiles = ceil_div(n4, BN);
  job.tile_base = 0;
  p.out[0] = logits;
  p.ldc[0] = ldc;
  p.m[0] = m;
  p.n[0] = n4;
  p.njobs = 1;
  p.total_tiles = job.m_tiles * job.n_tiles;
  p.alpha0 = x3 ? 1.0f / (kOperandScaleX3 * kOperandScaleX3) : 1.0f;
  p.log_alpha = log_scale;
  p.stages = ring_stages(0);
  return launch_gemm(p, cg, epi_warps(), 0, st);
}

static int check_peers(const Workspace& w, void* ws, const void* const* peer_ws) {
  if (w.pb.world < 2 || w.pb.world > SCLIP_MAX_PEERS || peer_ws == nullptr) {
    set_error("peer-memory calls need 2 <= world <= %d and a peer workspace table", SCLIP_MAX_PEERS);
    return SCLIP_ERR_ARGUMENT;
  }
  if (w.pb.rows_local * w.pb.world != w.pb.rows_global || w.pb.row_offset % w.pb.rows_local != 0) {
    set_error("peer-memory calls need equal row shards (rows_global = world * rows_local)");
    return SCLIP_ERR_ARGUMENT;
  }
  const int rank = w.pb.row_offset / w.pb.rows_local;
  for (int r = 0; r < w.pb.world; ++r)
    if (peer_ws[r] == nullptr || (reinterpret_cast<uintptr_t>(peer_ws[r]) & 255u) != 0) {
      set_error("peer workspace %d is null or not 256-byte aligned", r);
      return SCLIP_ERR_ARGUMENT;
    }
  if (peer_ws[rank] != ws) {
    set_error("peer_ws[rank] must be this rank's own workspace");
    return SCLIP_ERR_ARGUMENT;
  }
  return SCLIP_OK;
}

static bool bad_block(int block_threads, int limit) {
  if (block_threads >= 32 && block_threads <= limit && block_threads % 32 == 0) return false;
  set_error("block_threads=%d must be a multiple of 32 in [32, %d]", block_threads, limit);
  return true;
}

int sclip_push_shards(const sclip_problem* problem, void* ws, void* const* peer_ws, int max_blocks, int block_threads,
                      int epoch, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (!rc) rc = check_peers(w, ws, peer_ws);
  if (rc) return rc;
  if (bad_block(block_threads, 1024)) return SCLIP_ERR_ARGUMENT;
  return launch_push_shards(w, peer_ws, max_blocks, block_threads, epoch, static_cast<cudaStream_t>(stream));
}

int sclip_wait_shards(const sclip_problem* problem, void* ws, int epoch, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (w.pb.world < 2 || w.pb.world > SCLIP_MAX_PEERS || w.pb.rows_local * w.pb.world != w.pb.rows_global) {
    set_error("sclip_wait_shards needs 2 <= world <= %d equal row shards", SCLIP_MAX_PEERS);
    return SCLIP_ERR_ARGUMENT;
  }
  return launch_wait_shards(w, epoch, static_cast<cudaStream_t>(stream));
}

int sclip_forward_loss_peers(const sclip_problem* problem, void* ws, const void* const* peer_ws, float* loss3,
                             void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (!rc) rc = check_peers(w, ws, peer_ws);
  if (rc) return rc;
  if (loss3 == nullptr) {
    set_error("loss3 is null");
    return SCLIP_ERR_ARGUMENT;
  }
  return launch_forward_loss(w, nullptr, peer_ws, loss3, static_cast<cudaStream_t>(stream));
}

int sclip_read_status(const sclip_problem* problem, void* ws, int32_t* status_host, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (rc) return rc;
  if (status_host == nullptr) {
    set_error("status_host is null");
    return SCLIP_ERR_ARGUMENT;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SCLIP_CUDA_OK(cudaMemcpyAsync(status_host, w.status, 16, cudaMemcpyDeviceToHost, st));
  SCLIP_CUDA_OK(cudaStreamSynchronize(st));
  return SCLIP_OK;
}

int sclip_pull_reduce_cols(const sclip_problem* problem, void* ws, const void* const* peer_ws, int max_blocks,
                           int block_threads, void* stream) {
  Workspace w;
  int rc = resolve(problem, ws, &w);
  if (!rc) rc = check_peers(w, ws, peer_ws);
  if (rc) return rc;
  if (w.pb.dim % 4 != 0) {
    set_error("dim must be a multiple of 4");
    return SCLIP_ERR_ARGUMENT;
  }
  if (bad_block(block_threads, 512)) return SCLIP_ERR_ARGUMENT;
  return launch_pull_reduce(w, peer_ws, max_blocks, block_threads, static_cast<cudaStream_t>(stream));
}

// The single-GPU convenience calls remember, per workspace, what the last forward left for the backward:
// 0 nothing (forward only, or consumed), 1 statistics only (the backward recomputes the similarities), 2 a stash.
namespace {
struct KeptState {
  const void* ws;
  int state;
};
constexpr int kKeptEntries = 32;
KeptState g_kept[kKeptEntries];
std::mutex g_kept_mutex;

void set_kept(const void* ws, int state) {
  std::lock_guard<std::mutex> lock(g_kept_mutex);
  KeptState* slot = nullptr;
  for (KeptState& k : g_kept) {
    if (k.ws == ws) {
      slot = &k;
      break;
    }
    if (slot == nullptr && (k.ws == nullptr || k.state == 0)) slot = &k;
  }
  if (slot == nullptr) slot = &g_kept[0];
  slot->ws = ws;
  slot->state = state;
}
int get_kept(const void* ws) {
  std::lock_guard<std::mutex> lock(g_kept_mutex);
  for (const KeptState& k : g_kept)
    if (k.ws == ws) return k.state;
  return 0;
}
}  // namespace

int sclip_forward(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                  const float* t3, int keep_for_backward, float* loss3, void* stream) {
  if (problem != nullptr && problem->world != 1) {
    set_error("sclip_forward is the single-GPU entry point (world must be 1); use the stage calls when sharded");
    return SCLIP_ERR_ARGUMENT;
  }
  if (t3 == nullptr) {
    set_error("t3 is null");
    return SCLIP_ERR_ARGUMENT;
  }
  // stash when a backward follows, the operands are fp16 and the row is long enough that 4 bytes of HBM traffic per
  // logit beat 2 dim flop of recomputation (measured crossover on B200 between dim 512 and 768)
  const bool stash = keep_for_backward && problem != nullptr && problem->math == SCLIP_MATH_F16 && problem->dim >= 512;
  int rc = sclip_prologue(problem, ws, img, txt, aud, t3, stash ? SCLIP_PRO_DIAG : 0, stream);
  if (!rc) rc = sclip_forward_tiles_cols(problem, ws, t3, 7, 0, 1 << 30, stash ? SCLIP_FWD_STASH : 0, 0, 0, stream);
  if (!rc) rc = sclip_forward_reduce(problem, ws, stream);
  if (!rc) rc = sclip_forward_loss(problem, ws, nullptr, loss3, stream);
  if (ws != nullptr) set_kept(ws, rc ? 0 : (stash ? 2 : (keep_for_backward ? 1 : 0)));
  return rc;
}

int sclip_backward(const sclip_problem* problem, void* ws, const void* img, const void* txt, const void* aud,
                   const float* t3, const float* g3, void* dimg, void* dtxt, void* daud, int out_f32, float* dt3,
                   void* stream) {
  if (problem != nullptr && problem->world != 1) {
    set_error("sclip_backward is the single-GPU entry point (world must be 1); use the stage calls when sharded");
    return SCLIP_ERR_ARGUMENT;
  }
  const int kept = ws != nullptr ? get_kept(ws) : 0;
  if (kept == 0) {
    set_error("sclip_backward needs a preceding sclip_forward(keep_for_backward = 1) on this workspace; a forward "
              "serves one backward (its stashed tiles are converted in place)");
    return SCLIP_ERR_ARGUMENT;
  }
  const bool stashed = kept == 2;
  set_kept(ws, 0);
  int rc = stashed ? sclip_backward_scale(problem, ws, t3, g3, stream) : sclip_backward_tiles(problem, ws, t3, g3, stream);
  if (!rc) rc = sclip_backward_gemms_role(problem, ws, t3, g3, SCLIP_ROLE_BOTH, 0, 0, stream);
  if (!rc)
    rc = sclip_backward_finish(problem, ws, img, txt, aud, t3, g3, nullptr, 1.0f, dimg, dtxt, daud, out_f32,
                               stashed ? SCLIP_BWD_STASHED : 0, dt3, stream);
  return rc;
}

int sclip_gemm_f16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, float* c, int64_t ldc,
                   int m, int n, int k, float alpha, void* stream) {
  if (a == nullptr || b == nullptr || c == nullptr || m < 1 || n < 1 || k < 1 || n % 4 != 0 || ldc % 4 != 0 ||
      (reinterpret_cast<uintptr_t>(c) & 15u) != 0) {
    set_error("sclip_gemm_f16: bad argument (null pointer, non-positive size, n or ldc not a multiple of 4, or alignment)");
    return SCLIP_ERR_ARGUMENT;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = a_mn ? make_map(&p.maps[0], a, m, k, lda, 64, BK) : make_map(&p.maps[0], a, k, m, lda, BK, BM);
  if (!rc)
    rc = b_mn ? make_map(&p.maps[1], b, n, k, ldb, 64, BK) : make_map(&p.maps[1], b, k, n, ldb, BK, BN / cta_group());
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p.jobs[0].seg[0] = Segment{0, 1, a_mn ? 1 : 0, b_mn ? 1 : 0, ceil_div(k, BK)};
  p.jobs[0].nseg = 1;
  p.jobs[0].ksplits = 1;
  p.jobs[0].m_tiles = ceil_div(m, BM * cta_group());
  p.jobs[0].n_tiles = ceil_div(n, BN);
  p.jobs[0].tile_base = 0;
  p.out[0] = c;
  p.ldc[0] = ldc;
  p.m[0] = m;
  p.n[0] = n;
  p.njobs = 1;
  p.alpha0 = alpha;
  if (wide_enabled() && b_mn && n > 256) {
    p.wn = wide_width(n);
    p.jobs[0].n_tiles = ceil_div(n, p.wn);
    const int ks = wide_ksplits(p.jobs[0].m_tiles * p.jobs[0].n_tiles, ceil_div(k, BK), 0);
    p.jobs[0].ksplits = ks;
    if (ks > 1) {
      if (ldc != n) {
        set_error("sclip_gemm_f16: a k-split launch needs a dense output (ldc == n)");
        return SCLIP_ERR_ARGUMENT;
      }
      SCLIP_CUDA_OK(cudaMemsetAsync(c, 0, static_cast<size_t>(m) * n * 4, st));
    }
    p.total_tiles = p.jobs[0].m_tiles * p.jobs[0].n_tiles * ks;
    p.stages = wide_stages(p.wn);
    return launch_gemm_wide(p, epi_warps(), 0, st);
  }
  p.total_tiles = p.jobs[0].m_tiles * p.jobs[0].n_tiles;
  p.stages = ring_stages(0);
  return launch_gemm(p, cta_group(), epi_warps(), 0, st);
}

}  // extern "C"
