// Persistent, warp-specialised tcgen05 GEMM for the shared-MLP layers:
//     D[M x N] = A[M x K] * W[N x K]^T      (A = per-point activations, W = folded layer weights)
// rows (M) are context points, columns (N) are output channels.  Both operands are K-major and are
// staged global -> shared by TMA with the 128-byte swizzle; accumulators live in TMEM
// (128 lanes = 128 points of the tile, one 32-bit column per output channel).
//
//   warp 0      : TMA producer   (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (lane 0)
//   warps 2..5  : epilogue, one TMEM lane quarter each (tcgen05.ld -> registers -> global)
//
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, 2 accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1), and a static persistent tile loop.
//
// Epilogues
//   EPI_ACT    : out = [relu](acc + bias) stored bf16 / fp32 row-major (chain layers, context_proj)
//   EPI_FUSION : two accumulators per tile -- F (fusion conv, K = 1984) and G (gate layer 2, K = 64,
//                the trailing k-blocks of the same A/W rows) --
//                v = relu(F + bf) * (0.5 + 0.5 * sigmoid(G + bg))                 src/model.py:51,54-55
//                optional stores of v ((B,1024,N) fp32 and/or point-major operand type) and the
//                dual pooling (max / mean over the points of a segment, src/model.py:58-59) as a
//                warp transpose-reduce followed by one atomic per (warp, channel).
#pragma once
#include "ptx.cuh"

namespace lrn {

constexpr int BM = 128;             // points per tile (= TMEM lanes)
constexpr int kGemmThreads = 192;   // 6 warps
constexpr int kTileRowBytes = 128;  // one swizzle row: 64 bf16 or 32 tf32 along K

enum { EPI_ACT = 0, EPI_FUSION = 1 };
enum { FUSE_POOL = 1, FUSE_ARGMAX = 2, FUSE_STORE_CN = 4, FUSE_STORE_PM = 8 };

struct GemmParams {
  int M;        // rows (points) covered by this launch
  int m_tiles;  // ceil(M / 128)
  int n_tiles;  // N / BN
  int kb_main;  // k-blocks accumulated into F
  int kb_gate;  // k-blocks accumulated into G (EPI_FUSION only; they follow kb_main along K)
  int a_col0;   // first A column (elements) inside the A tensor map
  // EPI_ACT
  const float* bias;
  void* out;
  long long ldo;  // elements
  int out_f32;
  int a_tiled;  // A operand in the tiled layout ([row tile][column block][128 rows][128 B]); tmA is then a 4-D map
  int relu;
  int round_tf32;  // fp32 output is a TF32 operand of the next layer: round to nearest instead of truncating later
  int out_col0;    // pair kernels with staged TMA stores: first output column inside the output tensor map
  int k_splits;    // pair EPI_ACT kernels, fp32 output: > 1 = split-K, every split atomically adds its partial
                   // product into the (zeroed) output; 0 / 1 = one pass over K
  int kb_per_split;
  // EPI_FUSION
  const float* bias_f;
  const float* bias_g;
  long long row0;  // global point index (b*N + n) of row 0 of this launch
  int npts;        // points per segment (N)
  float inv_npts;
  int flags;
  float* global_feat;            // (B, 2048): max at [b*2048 + c], running mean at [b*2048 + 1024 + c]
  unsigned long long* pool_key;  // (B, 1024) packed (value bits << 32 | ~index) for argmax
  float* fused_cn;               // (B, 1024, N) fp32
  void* fused_pm;                // (M, 1024) operand type, rows local to this launch
  long long* dbg;                // optional: clock64() stamps of cluster 0 (tools/timeline.py), normally null
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kA = BM * kTileRowBytes;
  static constexpr int kB = BN * kTileRowBytes;
  static constexpr int kStage = kA + kB;
  static constexpr int kTiles = STAGES * kStage;
  static constexpr int kBarOff = kTiles;                 // full[STAGES], empty[STAGES], tfull[2], tempty[2]
  static constexpr int kTmemPtrOff = kBarOff + (2 * STAGES + 4) * 8;
  static constexpr int kBiasOff = kTmemPtrOff + 16;      // float sbias[2][2 * BN]
  static constexpr int kTotal = kBiasOff + 2 * 2 * BN * 4;
  static constexpr int kDynamic = kTotal + 1024;         // slack for manual 1024-byte alignment
};

// In-warp transpose-reduce: every lane holds 32 values (one per channel of the chunk, for its own
// point); afterwards v[0] on lane l is the reduction over the warp's 32 points of channel l.
// 31 shuffles instead of 32 x 5.
template <typename T, typename Op>
__device__ __forceinline__ void warp_transpose_reduce(T (&v)[32], int lane, Op op) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const T send = upper ? v[j] : v[j + o];
      const T keep = upper ? v[j + o] : v[j];
      const T recv = __shfl_xor_sync(0xffffffffu, send, o);
      v[j] = op(keep, recv);
    }
  }
}

struct OpMaxF { __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpAddF { __device__ __forceinline__ float operator()(float a, float b) const { return a + b; } };
struct OpMaxU64 {
  __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const {
    return a > b ? a : b;
  }
};

// Pool one 32-channel chunk of one warp (32 points) into segment `seg`; `mask` selects the lanes
// (points) that belong to it.  v >= 0 always (ReLU x positive gate), so 0 is the identity of both
// reductions and the signed-int ordering of the float bit patterns is the float ordering.
__device__ __forceinline__ void pool_chunk(const float (&v)[32], bool mask, long long seg, int n_in_seg, int ch0,
                                           int lane, const GemmParams& p) {
  float s[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = mask ? v[j] : 0.f;
  if (p.flags & FUSE_ARGMAX) {
    unsigned long long k[32];
    const unsigned int inv_idx = 0xFFFFFFFFu - static_cast<unsigned int>(n_in_seg);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      k[j] = mask ? ((static_cast<unsigned long long>(__float_as_uint(v[j])) << 32) | inv_idx) : 0ull;
    warp_transpose_reduce(k, lane, OpMaxU64());
    atomicMax(p.pool_key + seg * 1024 + ch0 + lane, k[0]);
  } else {
    float m[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) m[j] = s[j];
    warp_transpose_reduce(m, lane, OpMaxF());
    if (m[0] > 0.f) atomicMax(reinterpret_cast<int*>(p.global_feat + seg * 2048 + ch0 + lane), __float_as_int(m[0]));
  }
  warp_transpose_reduce(s, lane, OpAddF());
  if (s[0] > 0.f) atomicAdd(p.global_feat + seg * 2048 + 1024 + ch0 + lane, s[0] * p.inv_npts);
}

template <bool TF32>
__device__ __forceinline__ void store_row_chunk(void* out, long long elem_off, const float (&v)[32], bool as_f32,
                                                bool rna_tf32) {
  if (as_f32) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + elem_off);
    if (rna_tf32) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        dst[q] = make_float4(ptx::round_tf32(v[4 * q]), ptx::round_tf32(v[4 * q + 1]), ptx::round_tf32(v[4 * q + 2]),
                             ptx::round_tf32(v[4 * q + 3]));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  } else {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + elem_off);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_uint4(ptx::pack_bf16x2(v[8 * q], v[8 * q + 1]), ptx::pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                          ptx::pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), ptx::pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
  }
}

template <int BN, bool TF32, int EPI, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using L = GemmSmem<BN, STAGES>;
  constexpr int BK = TF32 ? 32 : 64;          // elements per 128-byte swizzle row
  constexpr int kMmaPerKb = 4;                // 128 B / 32 B per tcgen05.mma K step
  constexpr uint32_t kAccCols = (EPI == EPI_FUSION) ? 2 * BN : BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols;
  static_assert(kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns must be a power of two");
  constexpr uint32_t kIdesc = ptx::make_idesc(TF32, BM, BN);

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the shared address space: LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tfull = bar_empty + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  float* sbias = reinterpret_cast<float*>(smem + L::kBiasOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int kb_total = p.kb_main + p.kb_gate;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&bar_full[s], 1);
      ptx::mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bar_tfull[s], 1);
      ptx::mbar_init(&bar_tempty[s], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles;
        const int n_blk = tile - m_blk * p.n_tiles;
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(&bar_empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStage;
          ptx::mbar_arrive_expect_tx(&bar_full[stage], L::kStage);
          ptx::tma_load_2d(sa, &tmA, &bar_full[stage], p.a_col0 + kb * BK, m_blk * BM);
          ptx::tma_load_2d(sa + L::kA, &tmB, &bar_full[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&bar_tempty[as], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t acc_f = tmem_base + as * kAccCols;
      const uint32_t acc_g = acc_f + BN;
      for (int kb = 0; kb < kb_total; ++kb) {
        ptx::mbar_wait(&bar_full[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = ptx::smem_u32(smem + stage * L::kStage);
          const uint64_t da = ptx::make_smem_desc_sw128(a_addr);
          const uint64_t db = ptx::make_smem_desc_sw128(a_addr + L::kA);
          const bool gate = kb >= p.kb_main;
          const uint32_t d = gate ? acc_g : acc_f;
          const int kb_rel = gate ? kb - p.kb_main : kb;
#pragma unroll
          for (int k = 0; k < kMmaPerKb; ++k)
            ptx::tc_mma_ss<TF32>(d, da + 2 * k, db + 2 * k, kIdesc, (kb_rel > 0 || k > 0) ? 1u : 0u);
          ptx::tc_commit(&bar_empty[stage]);
          if (kb == kb_total - 1) ptx::tc_commit(&bar_tfull[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 64;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile - m_blk * p.n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      float* sb = sbias + as * 2 * BN;
      for (int c = et; c < BN; c += 128) {
        if (EPI == EPI_FUSION) {
          sb[c] = p.bias_f[n_blk * BN + c];
          sb[BN + c] = p.bias_g[n_blk * BN + c];
        } else {
          sb[c] = p.bias ? p.bias[n_blk * BN + c] : 0.f;
        }
      }
      ptx::named_bar_sync(1, 128);

      const int row = m_blk * BM + q * 32 + lane;  // local row of this thread
      const bool valid = row < p.M;
      ptx::mbar_wait(&bar_tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccCols;

      if (EPI == EPI_ACT) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_acc + c0, r);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(r[j]) + sb[c0 + j];
            v[j] = p.relu ? fmaxf(x, 0.f) : x;
          }
          if (valid)
            store_row_chunk<TF32>(p.out, static_cast<long long>(row) * p.ldo + n_blk * BN + c0, v, p.out_f32 != 0,
                                  p.round_tf32 != 0);
        }
      } else {
        const long long grow = p.row0 + row;
        const long long seg = valid ? grow / p.npts : 0;
        const int n_in_seg = valid ? static_cast<int>(grow - seg * p.npts) : 0;
        // warp-uniform fast path: all 32 points valid and in one segment
        const long long seg_lo = __shfl_sync(0xffffffffu, seg, 0);
        const bool uniform = __all_sync(0xffffffffu, valid && seg == seg_lo);
        long long seg_hi = seg_lo;
        if (!uniform) {
          // last valid lane's segment (rows are monotone in lane)
          const unsigned vm = __ballot_sync(0xffffffffu, valid);
          seg_hi = vm ? __shfl_sync(0xffffffffu, seg, 31 - __clz(vm)) : seg_lo - 1;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t rf[32], rg[32];
          ptx::tmem_ld_32x32b_x32(t_acc + c0, rf);
          ptx::tmem_ld_32x32b_x32(t_acc + BN + c0, rg);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f = fmaxf(__uint_as_float(rf[j]) + sb[c0 + j], 0.f);
            const float z = __uint_as_float(rg[j]) + sb[BN + c0 + j];
            const float g = 0.5f + __fdividef(0.5f, 1.f + __expf(-z));
            v[j] = valid ? f * g : 0.f;
          }
          const int ch0 = n_blk * BN + c0;
          if ((p.flags & FUSE_STORE_CN) && valid) {
            float* dst = p.fused_cn + (seg * 1024 + ch0) * p.npts + n_in_seg;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[static_cast<long long>(j) * p.npts] = v[j];
          }
          if ((p.flags & FUSE_STORE_PM) && valid)
            store_row_chunk<TF32>(p.fused_pm, static_cast<long long>(row) * 1024 + ch0, v, TF32, TF32);
          if (p.flags & (FUSE_POOL | FUSE_ARGMAX)) {
            if (uniform) {
              pool_chunk(v, true, seg_lo, n_in_seg, ch0, lane, p);
            } else {
              for (long long s = seg_lo; s <= seg_hi; ++s) pool_chunk(v, valid && seg == s, s, n_in_seg, ch0, lane, p);
            }
          }
        }
      }
      // accumulator stage drained -> hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_tempty[as]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace lrn
