// tcgen05.mma cadence probe (tuning aid, not part of the product library).
//
// Question it answers ("what's weak" 3): is the ~194 cycles per 256x256x16 shared-memory-operand MMA that
// the fusion kernel's timeline shows a property of the kernel (barriers, TMA, epilogue interference, power capping)
// or of the tensor pipe's operand delivery?  The probe issues a long back-to-back stream of MMAs on operands that are
// already resident (no TMA, no epilogue, optional side traffic) and reports cycles per MMA from clock64() together
// with the SM clock the stream ran at (cycles / %globaltimer ns), on one cluster and on the whole chip.
//
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe tools/mma_probe.cu
//   run  :  tools/mma_probe            (prints one line per configuration)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../pointnet_refine_b200/csrc/ptx.cuh"

using namespace lrn;

struct ProbeParams {
  int n;          // MMA N (64 / 128 / 256)
  int kblocks;    // k-blocks of 4 MMAs (K = 16 each) issued back to back
  int stages;     // operand stages walked round-robin (1 = the same 16 KB blocks every time)
  int commit;     // 1 = tcgen05.commit to a scratch barrier after every k-block (as a pipelined GEMM does)
  int alt;        // 1 = alternate between two accumulators per k-block (N <= 256 each)
  int side;       // side warps: 0 idle, 1 = tcgen05.ld loop, 2 = tcgen05.st loop (into the TMEM A region / spare columns)
  int zero;       // 1 = all-zero operands (power comparison)
  int fence;      // 1 = tcgen05.fence::after_thread_sync before every k-block (as after an mbarrier wait)
  int ring;       // 1 = full/empty mbarrier ring with a producer warp that only signals (a GEMM main loop without the loads)
  int ring_stages;
  int order;      // ring only: 0 = wait, MMAs, commit;  1 = peek at the NEXT stage's barrier before issuing this stage's MMAs;
                  // 2 = 1 + the commit of stage i is issued after the first MMA of stage i + 1
  int mma_per_kb; // MMAs per barrier round trip (4 = BK 64, 8 = BK 128)
  long long* out; // per cluster: cycles, ns, side iterations
};

constexpr int kProbeThreads = 192;
constexpr int kStageBytes = 16384 + 32768;  // A [128 x 64] bf16 + B up to [256 x 64] bf16
constexpr int kMaxStages = 4;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 2) ptx::tc_mma_ss_pair<false>(d, da, db, idesc, acc);
  else ptx::tc_mma_ss<false>(d, da, db, idesc, acc);
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 2) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (CG == 2) ptx::tc_commit_pair(bar, 3);
  else ptx::tc_commit(bar);
}

template <int CG, bool TS>
__global__ void __cluster_dims__(CG, 1, 1) __launch_bounds__(kProbeThreads, 1) probe_kernel(const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_done, bar_scratch;
  __shared__ uint64_t bar_full[8], bar_empty[8];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / CG;

  // operands: pseudo-random bf16 in [-0.5, 0.5) (or zeros)
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    const int words = p.stages * kStageBytes / 4;
    for (int i = threadIdx.x; i < words; i += kProbeThreads) {
      const uint32_t h = hash32(i * 2654435761u + blockIdx.x * 97u);
      // two bf16: exponent 0x3E (0.125..0.25) / 0x3F with random sign and mantissa -> |x| < 1
      const uint32_t lo = (h & 0x80FFu) | 0x3E00u, hi = ((h >> 16) & 0x80FFu) | 0x3E00u;
      w[i] = p.zero ? 0u : (lo | (hi << 16));
    }
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_done, 1);
    ptx::mbar_init(&bar_scratch, 1);
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(&bar_full[i], 1); ptx::mbar_init(&bar_empty[i], 1); }
    ptx::fence_mbar_init();
    stop_flag = 0;
  }
  ptx::fence_proxy_async_smem();
  if (warp == 1) {
    if constexpr (CG == 2) ptx::tmem_alloc_pair<512>(&tmem_ptr);
    else ptx::tmem_alloc<512>(&tmem_ptr);
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;

  if (warp >= 2) {  // fill TMEM columns [256, 512) (the A operand of the TS form) with packed bf16
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = 0; c < 8; ++c) {
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t h = hash32((threadIdx.x * 256 + c * 32 + j) * 40503u + blockIdx.x);
        r[j] = p.zero ? 0u : (((h & 0x80FFu) | 0x3E00u) | ((((h >> 16) & 0x80FFu) | 0x3E00u) << 16));
      }
      ptx::tmem_st_32x32b_x32(t_lane + 256 + 32 * c, r);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();

  if (warp == 0 && p.order >= 3 && leader) {
    // Warp-converged issue loop (CUTLASS style): every lane runs the loop, so descriptors and barrier addresses are
    // provably warp-uniform (uniform registers, no R2UR / per-lane serialisation); only the tcgen05 instructions are
    // predicated on elect.sync.  order 3 = wait, MMAs, commit;  order 4 = + peek at the next stage before the MMAs.
    const uint32_t idesc = ptx::make_idesc(false, 128u * CG, static_cast<uint32_t>(p.n));
    const uint64_t g0 = globaltimer_ns();
    const long long t0 = clock64();
    int st = 0, rs = 0;
    uint32_t rph = 0;
    bool peeked = false;
    for (int kb = 0; kb < p.kblocks; ++kb) {
      if (p.ring && !peeked) {
        uint32_t spins = 0;
        while (!ptx::mbar_try_wait(&bar_full[rs], rph)) { if (++spins > (1u << 22)) { printf("probe: ring timeout\n"); __trap(); } }
      }
      if (p.fence) ptx::tc_fence_after();
      const uint32_t a_addr = ptx::smem_u32(smem + st * kStageBytes);
      const uint64_t da = ptx::make_smem_desc_sw128(a_addr);
      const uint64_t db = ptx::make_smem_desc_sw128(a_addr + 16384);
      const int rs_cur = rs;
      if (p.ring) {
        if (++rs == p.ring_stages) { rs = 0; rph ^= 1; }
        peeked = p.order == 4 && kb + 1 < p.kblocks && ptx::mbar_try_wait(&bar_full[rs], rph);
      }
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (TS) mma_ts<CG>(tmem_base, tmem_base + 256 + ((kb * 32 + k * 8) & 255), db + 2 * k, idesc, 1u);
          else mma_ss<CG>(tmem_base, da + 2 * k, db + 2 * k, idesc, 1u);
        }
        if (p.ring) commit<CG>(&bar_empty[rs_cur]);
        else if (p.commit) commit<CG>(&bar_scratch);
      }
      __syncwarp();
      if (++st == p.stages) st = 0;
    }
    if (ptx::elect_one()) commit<CG>(&bar_done);
    __syncwarp();
    uint32_t spins = 0;
    while (!ptx::mbar_try_wait(&bar_done, 0)) { if (++spins > (1u << 22)) { printf("probe: timeout\n"); __trap(); } }
    const long long t1 = clock64();
    const uint64_t g1 = globaltimer_ns();
    if (lane == 0) {
      stop_flag = 1;
      p.out[cluster_id * 4 + 0] = t1 - t0;
      p.out[cluster_id * 4 + 1] = static_cast<long long>(g1 - g0);
    }
  } else if (warp == 0) {
    if (leader && lane == 0) {
      const uint32_t m = 128u * CG;
      const uint32_t idesc = ptx::make_idesc(false, m, static_cast<uint32_t>(p.n));
      const uint64_t g0 = globaltimer_ns();
      const long long t0 = clock64();
      int st = 0, rs = 0;
      uint32_t rph = 0;
      bool peeked = false;
      int pending_commit = -1;
      const int mpk = p.mma_per_kb > 0 ? p.mma_per_kb : 4;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        if (p.ring && !peeked) {
          uint32_t spins = 0;
          while (!ptx::mbar_try_wait(&bar_full[rs], rph)) { if (++spins > (1u << 22)) { printf("probe: ring timeout\n"); __trap(); } }
        }
        if (p.fence) ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + st * kStageBytes);
        const uint64_t da = ptx::make_smem_desc_sw128(a_addr);
        const uint64_t db = ptx::make_smem_desc_sw128(a_addr + 16384);
        const uint32_t d = tmem_base + ((p.alt && (kb & 1)) ? 128u : 0u) * (p.n > 128 ? 0u : 1u);
        const int rs_cur = rs;
        if (p.ring) {
          if (++rs == p.ring_stages) { rs = 0; rph ^= 1; }
          peeked = p.order >= 1 && kb + 1 < p.kblocks && ptx::mbar_try_wait(&bar_full[rs], rph);   // one non-blocking look
        }
        for (int k = 0; k < mpk; ++k) {
          if constexpr (TS) mma_ts<CG>(d, tmem_base + 256 + ((kb * 32 + k * 8) & 255), db + 2 * (k & 3), idesc, 1u);
          else mma_ss<CG>(d, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1u);
          if (k == 0 && pending_commit >= 0) { commit<CG>(&bar_empty[pending_commit]); pending_commit = -1; }
        }
        if (p.ring) {
          if (p.order == 2) pending_commit = rs_cur; else commit<CG>(&bar_empty[rs_cur]);
        } else if (p.commit) commit<CG>(&bar_scratch);
        if (++st == p.stages) st = 0;
      }
      if (pending_commit >= 0) commit<CG>(&bar_empty[pending_commit]);
      commit<CG>(&bar_done);
      uint32_t spins = 0;
      while (!ptx::mbar_try_wait(&bar_done, 0)) {
        if (++spins > (1u << 22)) { printf("probe: timeout\n"); __trap(); }
      }
      const long long t1 = clock64();
      const uint64_t g1 = globaltimer_ns();
      stop_flag = 1;
      p.out[cluster_id * 4 + 0] = t1 - t0;
      p.out[cluster_id * 4 + 1] = static_cast<long long>(g1 - g0);
    } else if (!leader && lane == 0) {
      uint32_t spins = 0;
      while (!ptx::mbar_try_wait(&bar_done, 0)) {
        if (++spins > (1u << 22)) { printf("probe: peer timeout\n"); __trap(); }
      }
      stop_flag = 1;
    }
    __syncwarp();
  } else if (warp == 1 && p.ring) {
    if (lane == 0) {   // producer of a GEMM main loop without its loads: waits for the stage to be free, signals it full (leader only)
      int rs = 0;
      uint32_t rph = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        uint32_t spins = 0;
        while (!ptx::mbar_try_wait(&bar_empty[rs], rph ^ 1)) { if (++spins > (1u << 22)) { printf("probe: producer timeout\n"); __trap(); } }
        if (leader) ptx::mbar_arrive(&bar_full[rs]);
        if (++rs == p.ring_stages) { rs = 0; rph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 2 && p.side) {
    // side traffic on tensor memory while the MMA stream runs
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    long long n = 0;
    uint32_t acc = 0;
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = 0x3E003E00u + j;
    while (!stop_flag) {
      const uint32_t col = 256u + 32u * (static_cast<uint32_t>(n) & 7u);  // SS: spare columns; TS: the A operand region
      if (p.side == 1) {
        ptx::tmem_ld_32x32b_x32(t_lane + col, r);
        ptx::tmem_ld_wait();
        acc += r[n & 31];
      } else {
        ptx::tmem_st_32x32b_x32(t_lane + col, r);
        ptx::tmem_st_wait();
      }
      ++n;
    }
    if (lane == 0 && warp == 2 && leader) p.out[cluster_id * 4 + 2] = n + (acc == 0x12345678u);
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CG == 2) ptx::tmem_dealloc_pair<512>(tmem_base);
    else ptx::tmem_dealloc<512>(tmem_base);
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int CG, bool TS>
void run(const char* name, int clusters, ProbeParams p, long long* d_out) {
  auto kern = probe_kernel<CG, TS>;
  const int smem = kMaxStages * kStageBytes + 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  p.out = d_out;
  std::vector<long long> h(clusters * 4);
  double best_cyc = 0, best_mhz = 0, best_tf = 0, side = 0;
  for (int rep = 0; rep < 4; ++rep) {   // rep 0 = warm-up
    CK(cudaMemset(d_out, 0, clusters * 4 * sizeof(long long)));
    kern<<<clusters * CG, kProbeThreads, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, clusters * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<double> cyc, ns;
    for (int c = 0; c < clusters; ++c) { cyc.push_back(double(h[c * 4])); ns.push_back(double(h[c * 4 + 1])); }
    std::sort(cyc.begin(), cyc.end());
    const double med = cyc[cyc.size() / 2], mx_ns = *std::max_element(ns.begin(), ns.end());
    const double mmas = double(p.kblocks) * (p.mma_per_kb > 0 ? p.mma_per_kb : 4);
    const double flop = 2.0 * 128 * CG * p.n * 16 * mmas * clusters;
    if (rep == 0) continue;
    const double tf = flop / (mx_ns * 1e-9) / 1e12;
    if (tf > best_tf) { best_tf = tf; best_cyc = med / mmas; best_mhz = double(h[0]) / double(h[1]) * 1e3; side = double(h[2]); }
  }
  printf("%-34s clusters=%3d N=%3d stages=%d commit=%d alt=%d side=%d zero=%d fence=%d ring=%d/%d order=%d mpk=%d : %7.1f cycles/MMA (median cluster)  %6.0f MHz  %7.1f TFLOP/s  floor=%d  side_iters=%.0f\n",
         name, clusters, p.n, p.stages, p.commit, p.alt, p.side, p.zero, p.fence, p.ring, p.ring_stages, p.order, p.mma_per_kb, best_cyc, best_mhz, best_tf, 128 * p.n / 256, side);
  fflush(stdout);
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  if (prop.major != 10) { printf("not sm_100: nothing to measure\n"); return 0; }
  long long* d_out;
  CK(cudaMalloc(&d_out, 148 * 4 * sizeof(long long)));
  const int kb_short = 4096, kb_long = 32768;   // 16k MMAs (~3M cycles) / 131k MMAs (~25M cycles = ~15 ms: long enough for the power cap)
  const int full = prop.multiProcessorCount / 2;
  ProbeParams p{};
  p.stages = 4; p.commit = 1; p.ring_stages = 5;
  if (argc > 1 && atoi(argv[1]) == 3) {   // third experiment: issue-loop orderings that hide the barrier round trip
    p.n = 256; p.kblocks = kb_short; p.ring = 1; p.ring_stages = 5;
    for (int order : {0, 1, 2}) for (int mpk : {4, 8}) {
      p.order = order; p.mma_per_kb = mpk; p.kblocks = kb_short * 4 / mpk;
      run<2, false>("SS pair ring", 1, p, d_out);
    }
    p.mma_per_kb = 4; p.kblocks = kb_short;
    p.order = 3; run<2, false>("SS pair ring converged", 1, p, d_out);
    p.order = 4; run<2, false>("SS pair ring converged + peek", 1, p, d_out);
    p.fence = 1; run<2, false>("SS pair ring converged + peek + fence", 1, p, d_out); p.fence = 0;
    p.ring = 0; p.order = 3; run<2, false>("SS pair bare converged", 1, p, d_out); p.ring = 1;
    p.order = 4; run<2, true>("TS pair ring converged + peek", 1, p, d_out);
    p.order = 4; p.kblocks = kb_long;
    run<2, false>("SS pair ring converged + peek chip", full, p, d_out);
    p.order = 2; p.mma_per_kb = 4; p.kblocks = kb_long;
    run<2, false>("SS pair ring order 2 chip", full, p, d_out);
    p.kblocks = kb_short;
    run<2, true>("TS pair ring order 2", 1, p, d_out);
    p.order = 1; run<2, true>("TS pair ring order 1", 1, p, d_out);
    CK(cudaFree(d_out));
    return 0;
  }
  if (argc > 1 && atoi(argv[1]) == 2) {   // second experiment: what a GEMM main loop's synchronisation adds to the bare stream
    p.n = 256; p.kblocks = kb_short;
    run<2, false>("SS pair bare", 1, p, d_out);
    p.fence = 1; run<2, false>("SS pair + fence per k-block", 1, p, d_out); p.fence = 0;
    for (int rs : {2, 3, 4, 5, 6, 8}) { p.ring = 1; p.ring_stages = rs; run<2, false>("SS pair + full/empty ring", 1, p, d_out); }
    p.ring_stages = 5; p.fence = 1; run<2, false>("SS pair + ring + fence", 1, p, d_out);
    p.kblocks = kb_long; run<2, false>("SS pair + ring + fence chip", full, p, d_out);
    p.kblocks = kb_short; p.fence = 0;
    run<1, false>("SS single + ring", 1, p, d_out);
    run<2, true>("TS pair + ring", 1, p, d_out);
    CK(cudaFree(d_out));
    return 0;
  }

  // 1. one cluster vs whole chip, SS, pair, N = 256 (the fusion kernel's MMA)
  p.n = 256; p.kblocks = kb_short;
  run<2, false>("SS pair 256x256x16  1 cluster", 1, p, d_out);
  p.kblocks = kb_long;
  run<2, false>("SS pair 256x256x16  chip", full, p, d_out);
  p.zero = 1; run<2, false>("SS pair 256x256x16  chip zeros", full, p, d_out); p.zero = 0;
  p.commit = 0; run<2, false>("SS pair 256x256x16  chip nocommit", full, p, d_out); p.commit = 1;
  p.stages = 1; run<2, false>("SS pair 256x256x16  chip 1 stage", full, p, d_out); p.stages = 4;
  // 2. N sweep, pair and single CTA (operand-delivery model: cycles = (A bytes + full-B bytes) / X)
  p.kblocks = kb_short;
  for (int n : {64, 128, 192, 256}) { p.n = n; run<2, false>("SS pair  N sweep  1 cluster", 1, p, d_out); }
  for (int n : {64, 128, 192, 256}) { p.n = n; run<1, false>("SS single N sweep  1 CTA", 1, p, d_out); }
  p.n = 128; p.alt = 1; run<2, false>("SS pair N=128 alternating acc", 1, p, d_out); p.alt = 0;
  // 3. A operand in tensor memory
  for (int n : {64, 128, 256}) { p.n = n; run<2, true>("TS pair  N sweep  1 cluster", 1, p, d_out); }
  for (int n : {128, 256}) { p.n = n; run<1, true>("TS single N sweep  1 CTA", 1, p, d_out); }
  p.n = 256; p.kblocks = kb_long;
  run<2, true>("TS pair 256x256x16  chip", full, p, d_out);
  run<1, true>("TS single 128x256x16 chip", 2 * full, p, d_out);
  run<1, false>("SS single 128x256x16 chip", 2 * full, p, d_out);
  // 4. side traffic on tensor memory (what an epilogue / an A-operand producer does concurrently)
  p.kblocks = kb_short;
  p.side = 1; run<2, false>("SS pair + tcgen05.ld loop", 1, p, d_out);
  p.side = 2; run<2, false>("SS pair + tcgen05.st loop", 1, p, d_out);
  p.side = 1; run<2, true>("TS pair + tcgen05.ld loop", 1, p, d_out);
  p.side = 2; run<2, true>("TS pair + tcgen05.st loop", 1, p, d_out);
  p.kblocks = kb_long;
  p.side = 2; run<2, true>("TS pair + tcgen05.st loop chip", full, p, d_out);
  p.side = 1; run<2, false>("SS pair + tcgen05.ld loop chip", full, p, d_out);
  p.side = 0;
  CK(cudaFree(d_out));
  return 0;
}
