// Shared definitions of the tcgen05 GEMM kernels (gemm_pair_sm100.cuh, chain_pair_sm100.cuh): tile constants,
// kernel parameters, the pooling reductions and row stores used by the epilogues.
//
//     D[M x N] = A[M x K] * W[N x K]^T      (A = per-point activations, W = folded layer weights)
// rows (M) are context points, columns (N) are output channels.  Both operands are K-major and are
// staged global -> shared by TMA with the 128-byte swizzle; accumulators live in TMEM
// (128 lanes = 128 points of a CTA's tile, one 32-bit column per output channel).
//
// Epilogues
//   EPI_ACT    : out = [relu](acc + bias) stored bf16 / fp32 row-major (tf32-tier layers, context_proj, decoder linears)
//   EPI_FUSION : two accumulators per tile -- F (fusion conv, K = 1984) and G (gate layer 2, K = 64,
//                the trailing k-blocks of the same A/W rows) --
//                v = relu(F + bf) * (0.5 + 0.5 * sigmoid(G + bg))                 src/model.py:51,54-55
//                optional stores of v ((B,1024,N) fp32 and/or point-major operand type) and the
//                dual pooling (max / mean over the points of a segment, src/model.py:58-59) as a
//                warp transpose-reduce followed by one atomic per (warp, channel).
#pragma once
#include "ptx.cuh"

// clock64() stamps for tools/timeline.py exist only in the tuning build (-DLRN_TIMELINE, liblrn_b200_timeline.so): the
// product library carries no timing code in its TMA / MMA / epilogue loops.
#ifdef LRN_TIMELINE
#define LRN_STAMP(cond, buf, idx) do { if (cond) (buf)[idx] = clock64(); } while (0)
#else
#define LRN_STAMP(cond, buf, idx) do { } while (0)
#endif

namespace lrn {

constexpr int BM = 128;             // points per tile (= TMEM lanes)
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
  // fp32x3 tier (3 x TF32 split, "exact" argmax tier): every fp32 operand is stored as hi = rna_tf32(v) and
  // lo = rna_tf32(v - hi), `*_lo_off` elements further along the row, and the k-loop runs three passes
  // A_hi B_hi + A_lo B_hi + A_hi B_lo into the same accumulator (kb_main / kb_gate count all three passes).
  int split3;
  int a_lo_off, b_lo_off;  // elements
  long long out_lo_off;    // EPI_ACT with round_tf32 / FUSE_STORE_PM: where the lo half of an output row goes
  int exact_gate;          // EPI_FUSION: keep G in tensor memory until F is complete and evaluate the gate in full fp32
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
  int a_ring;                    // -DLRN_TIMELINE build only (tools/l2_ring_probe.py): A row tiles are read modulo this count
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

// fp32 values as pairs of TF32 operands (fp32x3 tier): hi = rna_tf32(v) at `elem_off`, lo = rna_tf32(v - hi) `lo_off` further
__device__ __forceinline__ void store_row_chunk_split(float* out, long long elem_off, long long lo_off, const float (&v)[32]) {
  float4* dh = reinterpret_cast<float4*>(out + elem_off);
  float4* dl = reinterpret_cast<float4*>(out + elem_off + lo_off);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = ptx::round_tf32(v[4 * q + j]);
      l[j] = ptx::round_tf32(v[4 * q + j] - h[j]);
    }
    dh[q] = make_float4(h[0], h[1], h[2], h[3]);
    dl[q] = make_float4(l[0], l[1], l[2], l[3]);
  }
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

}  // namespace lrn
