// Train-mode cross attention of the decoder (reference: nn.MultiheadAttention cross_attn inside
// DetrTransformerDecoderLayer.forward under model.train(), src/model.py:84,123-128: 8 heads x 32, dropout 0.1 on the attention
// weights) for the shape this model has: 32 queries per segment, N context points, K / V of all six layers living in one
// (B, N, 6, 8, 32) bf16 buffer (the output of the K / V projection GEMM).  A generic flash kernel pads the 32 queries to its
// 128-row tile and wants (B, H, N, 32) tensors; here one CTA owns a (segment, head): its four warps take the key blocks of 32
// round-robin, the 32 x 32 query block lives in registers as mma.sync A fragments, K / V blocks are read ONCE, straight from
// the projection's layout, and the backward writes dK / dV of a key block in place (all 32 queries of the segment are in the
// CTA, so a key's gradient is complete there): no head transposes, no gradient gathers.
//
// Tensor cores through mma.sync.m16n8k16 (bf16, fp32 accumulate): per key block the work is 32 x 32 x 32, far below a
// tcgen05 tile, and the kernel is bound by reading K / V (64 + 64 bytes per key and head), not by the MMAs.
// Scores are kept in log2 units: q' = q * log2(e) / sqrt(32).
//
// Dropout is a counter-based hash of (seed, segment, head, query, key): forward and backward regenerate the same mask, no
// mask tensor exists.  One 32-bit hash serves two adjacent keys (16 bits each against p * 2^16); kept weights are scaled by
// 1 / (1 - p).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace lrn {

struct TrainAttnParams {
  const float* q;            // (B, 32, 256) fp32: in-projected queries (bias included), head h = columns [32h, 32h + 32)
  const __nv_bfloat16* k;    // key n of segment b: k + (b * N + n) * ld + 32 h   (a layer's column block of the K buffer)
  const __nv_bfloat16* v;
  long long ld, ldv;         // elements between consecutive keys of k / of v (6 * 256 for the six-layer buffers)
  int B, N;
  float* out;                // (B, 32, 256) fp32, heads concatenated (input of out_proj)
  float* lse;                // (B, 8, 32): log2 of the softmax denominator (+ running maximum), for the backward
  float p_drop;              // 0 = no dropout
  unsigned long long seed;
  const unsigned long long* seed_state;   // optional device word added to `seed` (a CUDA-graph replay draws a new mask without a new launch argument)
  // backward only
  const float* dout;         // (B, 32, 256)
  float* dq;                 // (B, 32, 256)
  __nv_bfloat16* dk;         // same addressing as k / v (its own row pitch)
  __nv_bfloat16* dv;
  long long ldg, ldgv;       // row pitch of dk / of dv
};

constexpr int kTaPitch = 40;                 // bf16 elements per shared-memory row (32 + 8: conflict-free ldmatrix)
constexpr int kTaTile = 32 * kTaPitch * 2;   // bytes of one 32 x 32 bf16 tile

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// Fragments of a 32 x 32 bf16 tile `t` (shared-memory byte address, row pitch kTaPitch):
//  A operand, rows [r0, r0 + 16), k columns [c0, c0 + 16)  (tile row-major [m][k])
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], uint32_t t, int r0, int c0, int lane) {
  const int mi = lane >> 3, row = lane & 7;
  ldsm_x4(a, t + ((r0 + (mi & 1) * 8 + row) * kTaPitch + c0 + (mi >> 1) * 8) * 2);
}
//  B operand from a tile stored [n][k]: n rows [n0, n0 + 8), all 32 k columns -> {b0, b1} of k-step 0 in r[0], r[1], of k-step 1 in r[2], r[3]
__device__ __forceinline__ void frag_b_nk(uint32_t (&r)[4], uint32_t t, int n0, int lane) {
  const int mi = lane >> 3, row = lane & 7;
  ldsm_x4(r, t + ((n0 + row) * kTaPitch + mi * 8) * 2);
}
//  B operand from a tile stored [k][n]: k rows [k0, k0 + 16), n columns [n0, n0 + 16) -> n-tile n0 in r[0], r[1], n-tile n0 + 8 in r[2], r[3]
__device__ __forceinline__ void frag_b_kn(uint32_t (&r)[4], uint32_t t, int k0, int n0, int lane) {
  const int mi = lane >> 3, row = lane & 7;
  ldsm_x4_t(r, t + ((k0 + (mi & 1) * 8 + row) * kTaPitch + n0 + (mi >> 1) * 8) * 2);
}

__device__ __forceinline__ float ta_ex2(float x) {   // 2^x, one MUFU op (exp2f adds range fix-ups the softmax does not need)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One 32-bit hash decides TWO adjacent keys (its low / high 16 bits against p * 2^16): `idx` is the element index of the
// even key of the pair.  Returns the keep / (1 - p) factors of (even key, odd key).
__device__ __forceinline__ void ta_keep_pair(unsigned long long seed, unsigned long long idx, uint32_t thresh16, float inv_keep,
                                             float& k0, float& k1) {
  // 32-bit mix of the 64-bit element index and the seed, then two multiply-xorshift rounds ("lowbias32")
  uint32_t x = static_cast<uint32_t>(idx) ^ (static_cast<uint32_t>(idx >> 32) * 0x85EBCA77u) ^ static_cast<uint32_t>(seed) ^
               (static_cast<uint32_t>(seed >> 32) * 0xC2B2AE3Du);
  x ^= x >> 16; x *= 0x21F0AAADu;
  x ^= x >> 15; x *= 0x735A2D97u;
  x ^= x >> 15;
  k0 = (x & 0xFFFFu) >= thresh16 ? inv_keep : 0.f;
  k1 = (x >> 16) >= thresh16 ? inv_keep : 0.f;
}

// 32 keys x 32 channels of K and V (one head): global -> registers (rows past N are zero), registers -> this warp's
// shared-memory tiles.  Split so that the next block's loads are in flight while the current block is multiplied.
struct TaKvRegs {
  uint4 k[4], v[4];
};
__device__ __forceinline__ void ta_fetch_kv(TaKvRegs& r, const __nv_bfloat16* k, const __nv_bfloat16* v, long long ld, long long ldv,
                                            int key0, int N, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i, row = id >> 2, c = id & 3;
    const bool in = key0 + row < N;
    r.k[i] = in ? *reinterpret_cast<const uint4*>(k + static_cast<long long>(key0 + row) * ld + 8 * c) : make_uint4(0, 0, 0, 0);
    r.v[i] = in ? *reinterpret_cast<const uint4*>(v + static_cast<long long>(key0 + row) * ldv + 8 * c) : make_uint4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void ta_park_kv(const TaKvRegs& r, uint8_t* sk, uint8_t* sv, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i, row = id >> 2, c = id & 3;
    *reinterpret_cast<uint4*>(sk + (row * kTaPitch + 8 * c) * 2) = r.k[i];
    *reinterpret_cast<uint4*>(sv + (row * kTaPitch + 8 * c) * 2) = r.v[i];
  }
}

// (32, 32) fp32 rows of one head, scaled -> bf16 shared-memory tile (CTA-wide, 128 threads)
__device__ __forceinline__ void ta_stage_rows(const float* src, float scale, uint8_t* tile, int tid) {
  for (int i = tid; i < 32 * 8; i += 128) {
    const int row = i >> 3, c4 = i & 7;
    const float4 t = *reinterpret_cast<const float4*>(src + row * 256 + 4 * c4);
    *reinterpret_cast<uint2*>(tile + (row * kTaPitch + 4 * c4) * 2) =
        make_uint2(ptx::pack_bf16x2(t.x * scale, t.y * scale), ptx::pack_bf16x2(t.z * scale, t.w * scale));
  }
}

constexpr float kTaQScale = 0.17677669529663688f * 1.4426950408889634f;  // log2(e) / sqrt(32)

// ---------------------------------------------------------------------------------------------------------------------
// forward: out = softmax(q k^T / sqrt(32)) [dropout] v per (segment, head); lse for the backward
// shared memory: Q tile | per warp K, V tiles | merge scratch (aliases the K / V tiles)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTaFwdSmem = kTaTile + 4 * 2 * kTaTile + 4 * 34 * 32 * 4;

__global__ void __launch_bounds__(128, 3) train_attn_fwd_kernel(const TrainAttnParams p) {
  extern __shared__ __align__(16) uint8_t ta_smem[];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x >> 3, h = blockIdx.x & 7;
  uint8_t* sq = ta_smem;
  uint8_t* sk = ta_smem + kTaTile + w * 2 * kTaTile;
  uint8_t* sv = sk + kTaTile;
  float* red = reinterpret_cast<float*>(ta_smem + kTaTile + 4 * 2 * kTaTile);  // [warp][34][32]: m, l, O[32 channels] per query
  ta_stage_rows(p.q + static_cast<long long>(b) * 32 * 256 + h * 32, kTaQScale, sq, tid);
  __syncthreads();
  uint32_t qa[2][2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) frag_a(qa[mt][ks], ptx::smem_u32(sq), 16 * mt, 16 * ks, lane);

  const __nv_bfloat16* kb = p.k + static_cast<long long>(b) * p.N * p.ld + h * 32;
  const __nv_bfloat16* vb = p.v + static_cast<long long>(b) * p.N * p.ldv + h * 32;
  const uint32_t thresh = static_cast<uint32_t>(p.p_drop * 65536.f);
  const unsigned long long seed = p.seed + (p.seed_state ? __ldg(p.seed_state) : 0ull);
  const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
  float o[2][4][4], mrun[2][2], lrun[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
    mrun[mt][0] = mrun[mt][1] = -INFINITY;
    lrun[mt][0] = lrun[mt][1] = 0.f;
  }
  const int nblk = (p.N + 31) >> 5;
  TaKvRegs kv;
  if (w < nblk) ta_fetch_kv(kv, kb, vb, p.ld, p.ldv, w << 5, p.N, lane);
  for (int blk = w; blk < nblk; blk += 4) {
    const int key0 = blk << 5;
    __syncwarp();
    ta_park_kv(kv, sk, sv, lane);
    __syncwarp();
    if (blk + 4 < nblk) ta_fetch_kv(kv, kb, vb, p.ld, p.ldv, (blk + 4) << 5, p.N, lane);
    float s[2][4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t kf[4];
      frag_b_nk(kf, ptx::smem_u32(sk), 8 * nt, lane);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
        mma16816(s[mt][nt], qa[mt][0], kf[0], kf[1]);
        mma16816(s[mt][nt], qa[mt][1], kf[2], kf[3]);
      }
    }
    // element e of s[mt][nt]: query 16 mt + g + 8 (e >> 1), key key0 + 8 nt + 2 t + (e & 1)
    uint32_t pa[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float bm[2] = {mrun[mt][0], mrun[mt][1]};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (key0 + 8 * nt + 2 * t + (e & 1) >= p.N) s[mt][nt][e] = -INFINITY;
          bm[e >> 1] = fmaxf(bm[e >> 1], s[mt][nt][e]);
        }
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        bm[hi] = fmaxf(bm[hi], __shfl_xor_sync(0xffffffffu, bm[hi], 1));
        bm[hi] = fmaxf(bm[hi], __shfl_xor_sync(0xffffffffu, bm[hi], 2));
        const float sc = ta_ex2(mrun[mt][hi] - bm[hi]);  // 0 on the first block
        lrun[mt][hi] *= sc;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          o[mt][nt][2 * hi] *= sc;
          o[mt][nt][2 * hi + 1] *= sc;
        }
        mrun[mt][hi] = bm[hi];
      }
      float ls[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const float p0 = ta_ex2(s[mt][nt][2 * hi] - bm[hi]), p1 = ta_ex2(s[mt][nt][2 * hi + 1] - bm[hi]);
          ls[hi] += p0 + p1;
          float k0 = 1.f, k1 = 1.f;
          if (thresh) {
            const int qi = 16 * mt + g + 8 * hi, key = key0 + 8 * nt + 2 * t;
            ta_keep_pair(seed, (static_cast<unsigned long long>(blockIdx.x) * 32 + qi) * p.N + key, thresh, inv_keep, k0, k1);
          }
          s[mt][nt][2 * hi] = p0 * k0;
          s[mt][nt][2 * hi + 1] = p1 * k1;
        }
      lrun[mt][0] += ls[0];
      lrun[mt][1] += ls[1];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        pa[mt][u][0] = ptx::pack_bf16x2(s[mt][2 * u][0], s[mt][2 * u][1]);
        pa[mt][u][1] = ptx::pack_bf16x2(s[mt][2 * u][2], s[mt][2 * u][3]);
        pa[mt][u][2] = ptx::pack_bf16x2(s[mt][2 * u + 1][0], s[mt][2 * u + 1][1]);
        pa[mt][u][3] = ptx::pack_bf16x2(s[mt][2 * u + 1][2], s[mt][2 * u + 1][3]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t vf[4];
        frag_b_kn(vf, ptx::smem_u32(sv), 16 * u, 16 * np, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(o[mt][2 * np], pa[mt][u], vf[0], vf[1]);
          mma16816(o[mt][2 * np + 1], pa[mt][u], vf[2], vf[3]);
        }
      }
  }
  // the row sums are still spread over the four lanes that share a row
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      lrun[mt][hi] += __shfl_xor_sync(0xffffffffu, lrun[mt][hi], 1);
      lrun[mt][hi] += __shfl_xor_sync(0xffffffffu, lrun[mt][hi], 2);
    }
  __syncthreads();  // every warp is done with its K / V tiles (the merge scratch lies behind them, but keep it simple)
  float* my = red + w * 34 * 32;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int qi = 16 * mt + g + 8 * hi;
      if (t == 0) {
        my[0 * 32 + qi] = mrun[mt][hi];
        my[1 * 32 + qi] = lrun[mt][hi];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        my[(2 + 8 * nt + 2 * t) * 32 + qi] = o[mt][nt][2 * hi];
        my[(2 + 8 * nt + 2 * t + 1) * 32 + qi] = o[mt][nt][2 * hi + 1];
      }
    }
  __syncthreads();
  for (int i = tid; i < 32 * 32; i += 128) {
    const int qi = i >> 5, c = i & 31;
    float M = -INFINITY;
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) M = fmaxf(M, red[(ww * 34) * 32 + qi]);
    float L = 0.f, acc = 0.f;
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) {
      const float mw = red[(ww * 34) * 32 + qi];
      const float f = mw == -INFINITY ? 0.f : ta_ex2(mw - M);
      L = fmaf(red[(ww * 34 + 1) * 32 + qi], f, L);
      acc = fmaf(red[(ww * 34 + 2 + c) * 32 + qi], f, acc);
    }
    p.out[(static_cast<long long>(b) * 32 + qi) * 256 + h * 32 + c] = acc / L;
    if (c == 0) p.lse[static_cast<long long>(blockIdx.x) * 32 + qi] = M + log2f(L);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward.  With P = softmax weights, D the dropout keep / (1 - p) mask, dPd = dO V^T, delta = rowsum(dO * O):
//   G = P * (D * dPd - delta);  dq = G K / sqrt(32);  dK = G^T q / sqrt(32);  dV = (P * D)^T dO
// Per key block a warp forms S and dPd as [query x key] fragments (they feed dq += G K directly), parks G and P * D as
// bf16 tiles in shared memory and reads them back TRANSPOSED with ldmatrix.trans as the A operands of dK = G^T q' and
// dV = Pd^T dO - no second pass of exponentials and mask hashes.
// shared memory: Q' tile | dO tile | lse[32] delta[32] | per warp K, V, G, Pd tiles; the dq merge scratch aliases them
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTaBwdSmem = 2 * kTaTile + 64 * 4 + 4 * 4 * kTaTile;
static_assert(4 * 4 * kTaTile >= 4 * 32 * 33 * 4, "dq merge scratch must fit behind the per-warp tiles");

//  A operand = TRANSPOSE of a tile stored [k][m]: m rows [m0, m0 + 16), k columns [k0, k0 + 16)
__device__ __forceinline__ void frag_a_t(uint32_t (&a)[4], uint32_t t, int m0, int k0, int lane) {
  const int mi = lane >> 3, row = lane & 7;
  // matrix mi covers (m half = mi & 1, k half = mi >> 1): stored rows = k, stored columns = m
  ldsm_x4_t(a, t + ((k0 + (mi >> 1) * 8 + row) * kTaPitch + m0 + (mi & 1) * 8) * 2);
}

__global__ void __launch_bounds__(128) train_attn_bwd_kernel(const TrainAttnParams p) {
  extern __shared__ __align__(16) uint8_t ta_smem[];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x >> 3, h = blockIdx.x & 7;
  uint8_t* sq = ta_smem;
  uint8_t* sdo = ta_smem + kTaTile;
  float* slse = reinterpret_cast<float*>(ta_smem + 2 * kTaTile);
  float* sdelta = slse + 32;
  uint8_t* wbase = ta_smem + 2 * kTaTile + 256;
  uint8_t* sk = wbase + w * 4 * kTaTile;
  uint8_t* sv = sk + kTaTile;
  uint8_t* sg = sv + kTaTile;
  uint8_t* sp = sg + kTaTile;
  float* red = reinterpret_cast<float*>(wbase);  // [warp][32 q][33], after the main loop
  const long long qoff = static_cast<long long>(b) * 32 * 256 + h * 32;
  ta_stage_rows(p.q + qoff, kTaQScale, sq, tid);
  ta_stage_rows(p.dout + qoff, 1.f, sdo, tid);
  if (tid < 32) {
    const float* orow = p.out + qoff + tid * 256;
    const float* drow = p.dout + qoff + tid * 256;
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      const float4 a = *reinterpret_cast<const float4*>(orow + c), bb = *reinterpret_cast<const float4*>(drow + c);
      d += a.x * bb.x + a.y * bb.y + a.z * bb.z + a.w * bb.w;
    }
    sdelta[tid] = d;
    slse[tid] = p.lse[static_cast<long long>(blockIdx.x) * 32 + tid];
  }
  __syncthreads();
  const uint32_t aq = ptx::smem_u32(sq), ado = ptx::smem_u32(sdo), ak = ptx::smem_u32(sk), av = ptx::smem_u32(sv);
  const uint32_t ag = ptx::smem_u32(sg), ap = ptx::smem_u32(sp);
  uint32_t qa[2][2][4], doa[2][2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      frag_a(qa[mt][ks], aq, 16 * mt, 16 * ks, lane);
      frag_a(doa[mt][ks], ado, 16 * mt, 16 * ks, lane);
    }
  float lse_r[2][2], del_r[2][2];  // this lane's rows
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      lse_r[mt][hi] = slse[16 * mt + g + 8 * hi];
      del_r[mt][hi] = sdelta[16 * mt + g + 8 * hi];
    }
  const __nv_bfloat16* kb = p.k + static_cast<long long>(b) * p.N * p.ld + h * 32;
  const __nv_bfloat16* vb = p.v + static_cast<long long>(b) * p.N * p.ldv + h * 32;
  __nv_bfloat16* dkb = p.dk + static_cast<long long>(b) * p.N * p.ldg + h * 32;
  __nv_bfloat16* dvb = p.dv + static_cast<long long>(b) * p.N * p.ldgv + h * 32;
  const uint32_t thresh = static_cast<uint32_t>(p.p_drop * 65536.f);
  const unsigned long long seed = p.seed + (p.seed_state ? __ldg(p.seed_state) : 0ull);
  const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
  const unsigned long long idx0 = static_cast<unsigned long long>(blockIdx.x) * 32;
  float dq[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[mt][nt][e] = 0.f;

  const int nblk = (p.N + 31) >> 5;
  TaKvRegs kv;
  if (w < nblk) ta_fetch_kv(kv, kb, vb, p.ld, p.ldv, w << 5, p.N, lane);
  for (int blk = w; blk < nblk; blk += 4) {
    const int key0 = blk << 5;
    __syncwarp();
    ta_park_kv(kv, sk, sv, lane);
    __syncwarp();
    if (blk + 4 < nblk) ta_fetch_kv(kv, kb, vb, p.ld, p.ldv, (blk + 4) << 5, p.N, lane);
    // ---- S = q' K^T, dPd = dO V^T as [query x key]; G and Pd
    uint32_t ga[2][2][4];
    {
      float s[2][4][4], dp[2][4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t kf[4], vf[4];
        frag_b_nk(kf, ak, 8 * nt, lane);
        frag_b_nk(vf, av, 8 * nt, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) s[mt][nt][e] = dp[mt][nt][e] = 0.f;
          mma16816(s[mt][nt], qa[mt][0], kf[0], kf[1]);
          mma16816(s[mt][nt], qa[mt][1], kf[2], kf[3]);
          mma16816(dp[mt][nt], doa[mt][0], vf[0], vf[1]);
          mma16816(dp[mt][nt], doa[mt][1], vf[2], vf[3]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int hi = 0; hi < 2; ++hi) {
            const int qi = 16 * mt + g + 8 * hi, key = key0 + 8 * nt + 2 * t;
            float keep[2] = {1.f, 1.f};
            if (thresh) ta_keep_pair(seed, (idx0 + qi) * p.N + key, thresh, inv_keep, keep[0], keep[1]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int e = 2 * hi + j;
              const float pe = key + j < p.N ? ta_ex2(s[mt][nt][e] - lse_r[mt][hi]) : 0.f;
              s[mt][nt][e] = pe * (keep[j] * dp[mt][nt][e] - del_r[mt][hi]);  // G
              dp[mt][nt][e] = pe * keep[j];                                    // Pd
            }
          }
          // park both as bf16 tiles [query][key]: element pairs (e = 0, 1) and (e = 2, 3) are adjacent keys of rows g, g + 8
          const int col = 8 * nt + 2 * t;
          *reinterpret_cast<uint32_t*>(sg + ((16 * mt + g) * kTaPitch + col) * 2) = ptx::pack_bf16x2(s[mt][nt][0], s[mt][nt][1]);
          *reinterpret_cast<uint32_t*>(sg + ((16 * mt + g + 8) * kTaPitch + col) * 2) = ptx::pack_bf16x2(s[mt][nt][2], s[mt][nt][3]);
          *reinterpret_cast<uint32_t*>(sp + ((16 * mt + g) * kTaPitch + col) * 2) = ptx::pack_bf16x2(dp[mt][nt][0], dp[mt][nt][1]);
          *reinterpret_cast<uint32_t*>(sp + ((16 * mt + g + 8) * kTaPitch + col) * 2) = ptx::pack_bf16x2(dp[mt][nt][2], dp[mt][nt][3]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          ga[mt][u][0] = ptx::pack_bf16x2(s[mt][2 * u][0], s[mt][2 * u][1]);
          ga[mt][u][1] = ptx::pack_bf16x2(s[mt][2 * u][2], s[mt][2 * u][3]);
          ga[mt][u][2] = ptx::pack_bf16x2(s[mt][2 * u + 1][0], s[mt][2 * u + 1][1]);
          ga[mt][u][3] = ptx::pack_bf16x2(s[mt][2 * u + 1][2], s[mt][2 * u + 1][3]);
        }
      }
    }
    // ---- dq += G K   (k = keys of this block)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t kf[4];
        frag_b_kn(kf, ak, 16 * u, 16 * np, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(dq[mt][2 * np], ga[mt][u], kf[0], kf[1]);
          mma16816(dq[mt][2 * np + 1], ga[mt][u], kf[2], kf[3]);
        }
      }
    __syncwarp();  // G / Pd tiles written by every lane
    // ---- dK = G^T q', dV = Pd^T dO   (m = keys, k = queries, n = channels)
    {
      float dkc[2][4][4], dvc[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) dkc[mt][nt][e] = dvc[mt][nt][e] = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {  // queries 16 u ..
        uint32_t gta[2][4], pta[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          frag_a_t(gta[mt], ag, 16 * mt, 16 * u, lane);
          frag_a_t(pta[mt], ap, 16 * mt, 16 * u, lane);
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t qf[4], df[4];
          frag_b_kn(qf, aq, 16 * u, 16 * np, lane);
          frag_b_kn(df, ado, 16 * u, 16 * np, lane);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma16816(dkc[mt][2 * np], gta[mt], qf[0], qf[1]);
            mma16816(dkc[mt][2 * np + 1], gta[mt], qf[2], qf[3]);
            mma16816(dvc[mt][2 * np], pta[mt], df[0], df[1]);
            mma16816(dvc[mt][2 * np + 1], pta[mt], df[2], df[3]);
          }
        }
      }
      // dK = ln2 * G^T q'  (q' = q log2e / sqrt(32));  rows = keys, columns = channels 8 nt + 2 t, + 1
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const int key = key0 + 16 * mt + g + 8 * hi;
          if (key < p.N) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              *reinterpret_cast<uint32_t*>(dkb + static_cast<long long>(key) * p.ldg + 8 * nt + 2 * t) =
                  ptx::pack_bf16x2(dkc[mt][nt][2 * hi] * 0.6931471805599453f, dkc[mt][nt][2 * hi + 1] * 0.6931471805599453f);
              *reinterpret_cast<uint32_t*>(dvb + static_cast<long long>(key) * p.ldgv + 8 * nt + 2 * t) =
                  ptx::pack_bf16x2(dvc[mt][nt][2 * hi], dvc[mt][nt][2 * hi + 1]);
            }
          }
        }
    }
  }
  // dq = G K / sqrt(32): sum the four warps' partial results
  __syncthreads();
  float* my = red + w * 32 * 33;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) my[(16 * mt + g + 8 * (e >> 1)) * 33 + 8 * nt + 2 * t + (e & 1)] = dq[mt][nt][e];
  __syncthreads();
  for (int i = tid; i < 32 * 32; i += 128) {
    const int qi = i >> 5, c = i & 31;
    const float sum = red[qi * 33 + c] + red[(32 + qi) * 33 + c] + red[(64 + qi) * 33 + c] + red[(96 + qi) * 33 + c];
    p.dq[qoff + qi * 256 + c] = sum * 0.17677669529663688f;
  }
}

}  // namespace lrn
