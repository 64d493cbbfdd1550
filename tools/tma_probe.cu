// L2 -> shared-memory delivery probe (tuning aid, not part of the product library).
//
// tools/mma_probe shows that the tensor pipe itself runs a 256x256x16 shared-memory-operand MMA in 128 cycles, on the
// whole chip.  The fusion kernel needs 32 KB of operands per CTA and k-block (A: its 128 rows, B: its half of the
// weight tile) = 64 B/clk per SM at that rate.  This probe streams exactly that traffic pattern through TMA with no
// MMAs at all and reports the bytes per clock and SM the L2 / fabric delivers, for
//   - unicast loads (cluster of 2, what the kernel does),
//   - a cluster of 4 (two CTA pairs) where every weight box is multicast to the two CTAs that need it.
//
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../pointnet_refine_b200/csrc/ptx.cuh"

using namespace lrn;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct TmaProbeParams {
  int kblocks;     // k-blocks per CTA
  int stages;
  int load_a, load_b;
  int multicast;   // cluster of 4 only: weight boxes are multicast to the two CTAs with the same rank parity
  int a_tiles;     // row tiles of the A matrix (the CTA walks tiles cta, cta + grid, ...)
  long long* out;  // per CTA: cycles, ns
};

constexpr int kBox = 16384;       // [128 rows x 64 bf16]
constexpr int kStage = 2 * kBox;  // A box + B box
constexpr int kMaxStages = 6;

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// tmA: tiled operand matrix [tile][32 column blocks][128 rows][64]; tmB: (1024, 2048) row-major weights, box 64 x 128 (or 64 x 64)
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(64, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBh, const TmaProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool mc = CL == 4 && p.multicast;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], mc ? 2 : 1);   // multicast: released by this CTA's consumer and by its partner's
    }
    ptx::fence_mbar_init();
  }
  ptx::cluster_sync();
  const uint32_t partner = rank ^ 2u;
  const uint32_t bytes = (p.load_a ? kBox : 0) + (p.load_b ? kBox : 0);
  if (warp == 0 && lane == 0) {
    const uint64_t g0 = globaltimer_ns();
    const long long t0 = clock64();
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; it < p.kblocks; ++it) {
      const int kb = it & 31, pass = it >> 5;           // 32 k-blocks per (tile, n-chunk) like the fusion kernel
      const int n_blk = pass & 3;
      const int tile = (blockIdx.x + (pass >> 2) * gridDim.x) % p.a_tiles;
      ptx::mbar_wait(&empty[st], ph ^ 1);
      ptx::mbar_arrive_expect_tx(&full[st], bytes);
      uint8_t* dst = smem + st * kStage;
      if (p.load_a) tma_load_4d(dst, &tmA, &full[st], 0, 0, kb, tile);
      if (p.load_b) {
        const int half = rank & 1;
        if (mc) {   // this CTA loads 64 of the 128 rows of its half and delivers them to itself and to its partner
          const int q = rank >> 1;
          tma_load_2d_mc(dst + kBox + q * (kBox / 2), &tmBh, &full[st], kb * 64, n_blk * 256 + half * 128 + q * 64,
                         static_cast<uint16_t>((1u << rank) | (1u << partner)));
        } else {
          ptx::tma_load_2d(dst + kBox, &tmB, &full[st], kb * 64, n_blk * 256 + half * 128);
        }
      }
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
    // drain
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_wait(&empty[st], ph ^ 1);
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
    const long long t1 = clock64();
    const uint64_t g1 = globaltimer_ns();
    p.out[blockIdx.x * 2 + 0] = t1 - t0;
    p.out[blockIdx.x * 2 + 1] = static_cast<long long>(g1 - g0);
  } else if (warp == 1 && lane == 0) {
    int st = 0;
    uint32_t ph = 0;
    const uint32_t e_partner0 = ptx::mapa(ptx::smem_u32(&empty[0]), partner);
    for (int it = 0; it < p.kblocks; ++it) {
      ptx::mbar_wait(&full[st], ph);
      ptx::mbar_arrive(&empty[st]);
      if (mc) ptx::mbar_arrive_cluster(e_partner0 + 8u * st);
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
  }
  __syncwarp();
  ptx::cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CL>
void run(const char* name, int ctas, TmaProbeParams p, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbh,
         long long* d_out) {
  auto kern = tma_probe_kernel<CL>;
  const int smem = kMaxStages * kStage + 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int max_clusters = 0;
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  }
  p.out = d_out;
  std::vector<long long> h(ctas * 2);
  double best = 0, best_cyc = 0, mhz = 0;
  for (int rep = 0; rep < 3; ++rep) {
    kern<<<ctas, 64, smem>>>(ta, tb, tbh, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, ctas * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<double> cyc;
    double max_ns = 0;
    for (int c = 0; c < ctas; ++c) { cyc.push_back(double(h[2 * c])); max_ns = std::max(max_ns, double(h[2 * c + 1])); }
    std::sort(cyc.begin(), cyc.end());
    if (rep == 0) continue;
    const double bytes = double(p.kblocks) * ((p.load_a ? kBox : 0) + (p.load_b ? kBox : 0));
    const double bpc = bytes / cyc[cyc.size() / 2];
    if (bpc > best) { best = bpc; best_cyc = cyc[cyc.size() / 2] / p.kblocks; mhz = double(h[0]) / double(h[1]) * 1e3; }
  }
  printf("%-40s ctas=%3d (max active clusters %3d x %d) stages=%d A=%d B=%d mc=%d : %6.1f B/clk/SM delivered  %7.1f cycles/k-block  %5.0f MHz  chip %6.0f B/clk\n",
         name, ctas, max_clusters, CL, p.stages, p.load_a, p.load_b, p.multicast, best, best_cyc, mhz, best * ctas);
  fflush(stdout);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  if (prop.major != 10) return 0;
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  auto enc = reinterpret_cast<EncodeTiledFn>(fnp);
  const int a_tiles = 2368;   // 303,104 points = one wave of the encoder: 1.24 GB, read 4 x per wave (3 x from L2)
  void *dA, *dB;
  CK(cudaMalloc(&dA, size_t(a_tiles) * 32 * kBox));
  CK(cudaMalloc(&dB, size_t(1024) * 2048 * 2));
  CK(cudaMemset(dA, 0x11, size_t(a_tiles) * 32 * kBox));
  CK(cudaMemset(dB, 0x22, size_t(1024) * 2048 * 2));
  CUtensorMap ta, tb, tbh;
  {
    cuuint64_t dims[4] = {64, 128, 32, cuuint64_t(a_tiles)};
    cuuint64_t strides[3] = {128, 16384, 32 * 16384};
    cuuint32_t box[4] = {64, 128, 1, 1}, es[4] = {1, 1, 1, 1};
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
  }
  for (int hrows : {128, 64}) {
    cuuint64_t dims[2] = {2048, 1024};
    cuuint64_t strides[1] = {4096};
    cuuint32_t box[2] = {64, cuuint32_t(hrows)}, es[2] = {1, 1};
    if (enc(hrows == 128 ? &tb : &tbh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
  }
  long long* d_out;
  CK(cudaMalloc(&d_out, 148 * 2 * sizeof(long long)));
  const int sms = prop.multiProcessorCount;
  TmaProbeParams p{};
  p.a_tiles = a_tiles;
  p.kblocks = 32 * 4 * 16;   // 16 tiles x 4 n-chunks x 32 k-blocks: one fusion launch
  for (int stages : {5, 6}) {
    p.stages = stages;
    p.load_a = 1; p.load_b = 1; p.multicast = 0;
    run<2>("pair, A + B unicast (fusion kernel)", sms, p, ta, tb, tbh, d_out);
    p.load_a = 0; run<2>("pair, B only (weights, L2 resident)", sms, p, ta, tb, tbh, d_out);
    p.load_a = 1; p.load_b = 0; run<2>("pair, A only (operand rows)", sms, p, ta, tb, tbh, d_out);
    p.load_b = 1;
    run<4>("cluster 4, A + B unicast", sms / 4 * 4, p, ta, tb, tbh, d_out);
    p.multicast = 1; run<4>("cluster 4, A + B multicast x2", sms / 4 * 4, p, ta, tb, tbh, d_out);
    p.load_a = 0; run<4>("cluster 4, B multicast x2 only", sms / 4 * 4, p, ta, tb, tbh, d_out);
    p.multicast = 0;
  }
  // one cluster alone: latency-bound rate of a 5-stage ring (no contention)
  p.stages = 5; p.load_a = 1; p.load_b = 1;
  run<2>("pair alone, A + B", 2, p, ta, tb, tbh, d_out);
  return 0;
}
