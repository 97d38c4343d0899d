// Probe (not part of the product): how many bytes per second can TMA deliver from an L2-resident matrix into the
// shared memory of all SMs, and does multicast inside a cluster raise that ceiling?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2probe tests/probes/l2_fabric_probe.cu -lcuda
//   ./l2probe
// Every CTA runs a producer thread (TMA loads into a ring of stages) and a consumer thread that frees each stage as
// soon as it has landed: no tensor-core work, the kernel measures the L2 -> SM path alone.
//   mode 0  unicast: every CTA loads its own A box (16 KiB) and the B box (16 KiB) of its cluster's column tile
//   mode 1  multicast: the B box is split over the CTAs of the cluster, each part multicast to all of them
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kStages = 6;  // ring slots allocated; P.stages of them are used
constexpr int kBoxBytes = 128 * 64 * 2;  // 128 rows x 64 fp16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

struct Params {
  CUtensorMap map_a;      // box 64 x 128
  CUtensorMap map_part;   // box 64 x (128 / cluster)
  int row_tiles;          // 128-row tiles of the matrix
  int kblocks;            // k blocks per tile (dim / 64)
  int tiles_per_cluster;  // tiles every cluster walks through
  int cl;
  int mode;
  int stages;
};

__global__ void __launch_bounds__(64, 1) probe_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[kStages], empty[kStages];
  const uint32_t rank = P.cl > 1 ? ctarank() : 0;
  const int cluster_id = blockIdx.x / P.cl, num_clusters = gridDim.x / P.cl;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], P.mode ? P.cl : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (P.cl > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  else __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < P.tiles_per_cluster; ++t) {
      // forward-like pattern: consecutive clusters share column tiles in groups of 8 row tiles
      const int id = cluster_id + t * num_clusters;
      const int ti = (id % 8 + (id / (8 * 64)) * 8) * P.cl + rank;   // A: this CTA's row tile
      const int tj = (id / 8) % 64;                                  // B: the cluster's column tile
      for (int kb = 0; kb < P.kblocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * 2 * kBoxBytes;
        uint8_t* sb = sa + kBoxBytes;
        mbar_expect(&full[stage], 2 * kBoxBytes);
        tma_load(sa, &P.map_a, &full[stage], kb * 64, (ti % P.row_tiles) * 128);
        if (P.mode == 0) {
          tma_load(sb, &P.map_a, &full[stage], kb * 64, ((tj + 128) % P.row_tiles) * 128);
        } else {
          const int part = 128 / P.cl;
          tma_load_mc(sb + rank * part * 128, &P.map_part, &full[stage], kb * 64,
                      ((tj + 128) % P.row_tiles) * 128 + rank * part, (uint16_t)((1u << P.cl) - 1));
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    const int total = P.tiles_per_cluster * P.kblocks;
    for (int i = 0; i < total; ++i) {
      mbar_wait(&full[stage], phase);
      if (P.mode == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      } else {
        for (int c = 0; c < P.cl; ++c) mbar_arrive_remote(&empty[stage], c);
      }
      if (++stage == P.stages) { stage = 0; phase ^= 1u; }
    }
  }
  if (P.cl > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

static PFN_cuTensorMapEncodeTiled_v12000 encoder() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return (PFN_cuTensorMapEncodeTiled_v12000)p;
}

static void make_map(CUtensorMap* m, void* ptr, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
  const int rows = 32768, dim = 768;
  void* buf;
  CK(cudaMalloc(&buf, (size_t)rows * dim * 2));
  CK(cudaMemset(buf, 0, (size_t)rows * dim * 2));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int smem = kStages * 2 * kBoxBytes + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  struct Cfg { int cl, mode; };
  const Cfg cfgs[] = {{1, 0}, {2, 0}, {2, 1}, {4, 0}, {4, 1}, {8, 0}, {8, 1}};
  for (int stages = 4; stages <= 6; stages += 2)
  for (const Cfg& c : cfgs) {
    Params p;
    memset(&p, 0, sizeof(p));
    make_map(&p.map_a, buf, dim, rows, 128);
    make_map(&p.map_part, buf, dim, rows, 128 / c.cl);
    p.row_tiles = rows / 128;
    p.kblocks = dim / 64;
    p.tiles_per_cluster = 64;
    p.cl = c.cl;
    p.mode = c.mode;
    p.stages = stages;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c.cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    cfg.gridDim = dim3(sms / c.cl * c.cl);
    CK(cudaOccupancyMaxActiveClusters(&max_clusters, probe_kernel, &cfg));
    int clusters = max_clusters < sms / c.cl ? max_clusters : sms / c.cl;
    cfg.gridDim = dim3(clusters * c.cl);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, probe_kernel, p));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) CK(cudaLaunchKernelEx(&cfg, probe_kernel, p));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double delivered = (double)clusters * c.cl * p.tiles_per_cluster * p.kblocks * 2.0 * kBoxBytes;
    const double read = c.mode ? (double)clusters * c.cl * p.tiles_per_cluster * p.kblocks * (kBoxBytes + kBoxBytes / (double)c.cl) : delivered;
    printf("L2PROBE stages=%d cluster=%d mode=%s active_clusters=%d (max %d) ctas=%d: %.3f ms  delivered %.2f TB/s  L2 reads %.2f TB/s\n",
           p.stages, c.cl, c.mode ? "multicast" : "unicast", clusters, max_clusters, clusters * c.cl, ms, delivered / ms / 1e9,
           read / ms / 1e9);
  }
  return 0;
}
