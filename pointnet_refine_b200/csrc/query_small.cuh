// Query side of the decoder for FEW polyline rows: the whole-scene loop of the reference calls
// the model with B = 1 (inference_whole_scene.py:130-139), i.e. 32 query rows, where a 256-row tensor-core tile would be
// 7/8 padding and the time is set by weight bandwidth and launch count, not by FLOPs.  Here every nn.Linear of
// DetrTransformerDecoderLayer.forward / PositionalEncoding / point_mlp (src/model.py:104-135,64-75,150-159) on such a batch is
// one launch of rows_linear_kernel: fp32 FMA, the weights of a layer read exactly once and spread over N/8 blocks, with
// the small producers fused into the operand load (the residual add `tgt + query_pos`, or the K = 3 first layer of an MLP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace lrn {

enum RowsLinearInput {
  ROWS_IN_PLAIN = 0,  // x
  ROWS_IN_SUM = 1,    // x + x2                      (with_pos_embed, src/model.py:101-102,113,124)
  ROWS_IN_MLP3 = 2,   // relu(W1 c + b1), c = 3 coordinates per row  (PositionalEncoding.mlp[0:2] :68-72, point_mlp[0:3] :151-153)
};

struct RowsLinearArgs {
  const float* x;     // PLAIN / SUM: (M, K) fp32, row pitch ldx.  MLP3: (M, 3) coordinates, row pitch ldx
  const float* x2;    // SUM: second addend (M, K), row pitch ldx2
  const float* w1;    // MLP3: (K, 3) first-layer weight and (K) bias
  const float* b1;
  const float* w;     // (N, K) fp32 row-major (nn.Linear.weight)
  const float* bias;  // (N) or null
  void* out;          // (M, N) fp32 or bf16, row pitch ldo
  long long ldx, ldx2, ldo;
  int M, N, K;        // N % 8 == 0, K % 32 == 0
  int relu, out_bf16;
};

// out[m][n] = act(sum_k x'[m][k] w[n][k] + bias[n]).  Block = 32 rows (blockIdx.y) x NB output columns (blockIdx.x); its
// eight warps split K; lane = row.  Per 32-wide piece of its k-slice a warp fetches the 32 rows (coalesced) and the NB
// weight rows (coalesced) into registers, parks them in shared memory (rows transposed: conflict-free column reads; weights
// read back as broadcast float4) and multiplies - while the loads of the next piece are already in flight, so a piece
// costs one L2 round trip and not one per dependent load.  The eight partial sums meet in shared memory in a fixed
// order: results do not depend on scheduling.  NB = 8, or 4 when that is needed to give the launch enough blocks.
template <int MODE, int NB>
__global__ void __launch_bounds__(256) rows_linear_kernel(const RowsLinearArgs a) {
  __shared__ float xs[8][32 * 33];
  __shared__ __align__(16) float wsm[8][NB * 32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * NB, m0 = blockIdx.y * 32;
  const int kslice = a.K >> 3;  // multiple of 4
  float* my = xs[w];
  float* myw = wsm[w];
  float acc[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) acc[c] = 0.f;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (MODE == ROWS_IN_MLP3 && m0 + lane < a.M) {
    const float* c = a.x + static_cast<long long>(m0 + lane) * a.ldx;
    cx = c[0];
    cy = c[1];
    cz = c[2];
  }
  const int kbeg = w * kslice, kend = kbeg + kslice;
  float xr[MODE == ROWS_IN_MLP3 ? 4 : 32], wr[NB];
  auto fetch = [&](int k0) {  // registers <- global: piece [k0, k0 + 32) of this warp's slice (columns past kend read as 0)
    const bool kin = k0 + lane < kend;
    if (MODE == ROWS_IN_MLP3) {  // lane = hidden unit k0 + lane: its three weights and bias
      xr[0] = kin ? a.w1[3 * (k0 + lane)] : 0.f;
      xr[1] = kin ? a.w1[3 * (k0 + lane) + 1] : 0.f;
      xr[2] = kin ? a.w1[3 * (k0 + lane) + 2] : 0.f;
      xr[3] = kin ? a.b1[k0 + lane] : 0.f;
    } else {
#pragma unroll
      for (int m = 0; m < 32; ++m) {
        float v = 0.f;
        if (m0 + m < a.M && kin) {
          v = a.x[static_cast<long long>(m0 + m) * a.ldx + k0 + lane];
          if (MODE == ROWS_IN_SUM) v += a.x2[static_cast<long long>(m0 + m) * a.ldx2 + k0 + lane];
        }
        xr[m] = v;
      }
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) wr[c] = kin ? __ldg(a.w + static_cast<long long>(n0 + c) * a.K + k0 + lane) : 0.f;
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
    __syncwarp();  // the previous piece has been consumed
    if (MODE == ROWS_IN_MLP3) {
#pragma unroll 8
      for (int m = 0; m < 32; ++m) {
        const float px = __shfl_sync(0xffffffffu, cx, m), py = __shfl_sync(0xffffffffu, cy, m), pz = __shfl_sync(0xffffffffu, cz, m);
        my[lane * 33 + m] = fmaxf(fmaf(xr[0], px, fmaf(xr[1], py, fmaf(xr[2], pz, xr[3]))), 0.f);
      }
    } else {
#pragma unroll
      for (int m = 0; m < 32; ++m) my[lane * 33 + m] = xr[m];
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) myw[c * 32 + lane] = wr[c];
    __syncwarp();
    if (k0 + 32 < kend) fetch(k0 + 32);
    const int kc = min(32, kend - k0);
#pragma unroll 2
    for (int kk = 0; kk < kc; kk += 4) {
      const float x0 = my[kk * 33 + lane], x1 = my[(kk + 1) * 33 + lane], x2 = my[(kk + 2) * 33 + lane], x3 = my[(kk + 3) * 33 + lane];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const float4 wv = *reinterpret_cast<const float4*>(myw + c * 32 + kk);
        acc[c] = fmaf(wv.x, x0, fmaf(wv.y, x1, fmaf(wv.z, x2, fmaf(wv.w, x3, acc[c]))));
      }
    }
  }
  __syncthreads();
  float* red = &xs[0][0];  // [warp][column][row]
#pragma unroll
  for (int c = 0; c < NB; ++c) red[(w * NB + c) * 32 + lane] = acc[c];
  __syncthreads();
  if (threadIdx.x < 32 * NB) {
    const int m = threadIdx.x / NB, c = threadIdx.x % NB;
    float s = a.bias ? a.bias[n0 + c] : 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += red[(ww * NB + c) * 32 + m];
    if (a.relu) s = fmaxf(s, 0.f);
    if (m0 + m < a.M) {
      const long long o = static_cast<long long>(m0 + m) * a.ldo + n0 + c;
      if (a.out_bf16) reinterpret_cast<__nv_bfloat16*>(a.out)[o] = __float2bfloat16_rn(s);
      else reinterpret_cast<float*>(a.out)[o] = s;
    }
  }
}

// out[r][0:256) = [rna_tf32](relu(W1 c_r + b1)) in fp32: first layer of PositionalEncoding (src/model.py:68-72) on the current
// polyline points (:212; rows of 3 floats) or on the context points (:197; rows of 4 floats, tf32 tier) when its second
// layer runs on the tensor cores.  One warp per row, lane = 8 consecutive channels; `ld` = floats per coordinate row.
__global__ void __launch_bounds__(256) query_pos_hidden_kernel(const float* __restrict__ cur, long long ld, long long rows,
                                                               const float* __restrict__ w1 /* (256,3) */,
                                                               const float* __restrict__ b1, float* __restrict__ out, int rna) {
  const int lane = threadIdx.x & 31;
  float wx[8], wy[8], wz[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = 8 * lane + j;
    wx[j] = w1[3 * c];
    wy[j] = w1[3 * c + 1];
    wz[j] = w1[3 * c + 2];
    bb[j] = b1[c];
  }
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float px = cur[ld * r], py = cur[ld * r + 1], pz = cur[ld * r + 2];
    float h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = fmaxf(fmaf(wx[j], px, fmaf(wy[j], py, fmaf(wz[j], pz, bb[j]))), 0.f);
      if (rna) h[j] = ptx::round_tf32(h[j]);
    }
    float4* o = reinterpret_cast<float4*>(out + r * 256 + 8 * lane);
    o[0] = make_float4(h[0], h[1], h[2], h[3]);
    o[1] = make_float4(h[4], h[5], h[6], h[7]);
  }
}

// out = [rna_tf32](a + b) over n4 float4 elements (with_pos_embed ahead of a tensor-core linear); b may be null
// (then: a copy rounded to the nearest TF32 value, the operand form of the tf32 tier's GEMMs).
__global__ void __launch_bounds__(256) add_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                  float4* __restrict__ out, long long n4, int rna) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 u = a[i];
    if (b) {
      const float4 v = b[i];
      u = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
    if (rna) u = make_float4(ptx::round_tf32(u.x), ptx::round_tf32(u.y), ptx::round_tf32(u.z), ptx::round_tf32(u.w));
    out[i] = u;
  }
}

// Cross attention of the 32 polyline queries of a segment over its N context points in fp32 (tf32 tier: K / V come from
// the hoisted TF32 GEMMs; nn.MultiheadAttention cross_attn in eval mode, src/model.py:123-128).  Block = (segment, head);
// lane = query; the eight warps take the key blocks of 32 round-robin, each with its own running (max, sum, 32 outputs),
// merged through shared memory at the end.  K / V rows of a block are parked in shared memory and read as broadcasts.
//   q (B*32, 256) fp32 in-projected queries; k, v: row (b*N + n) at k + (b*N + n) * ld, head h at + 32 h; out (B*32, 256).
constexpr int kCrossAttnSmem = 8 * 2 * 32 * 33 * 4;
__global__ void __launch_bounds__(256) cross_attn32_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                           const float* __restrict__ v, long long ld, int N,
                                                           float* __restrict__ out) {
  extern __shared__ float sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x >> 3, h = blockIdx.x & 7;
  float* ks = sm + w * 2 * 32 * 33;
  float* vs = ks + 32 * 33;
  float qr[32], acc[32];
  {
    const float* qp = q + (static_cast<long long>(b) * 32 + lane) * 256 + h * 32;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qp + c);
      qr[c] = t.x * 0.17677669529663688f; qr[c + 1] = t.y * 0.17677669529663688f;   // 1 / sqrt(32)
      qr[c + 2] = t.z * 0.17677669529663688f; qr[c + 3] = t.w * 0.17677669529663688f;
    }
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  const float* kb = k + static_cast<long long>(b) * N * ld + h * 32 + lane;
  const float* vb = v + static_cast<long long>(b) * N * ld + h * 32 + lane;
  for (int n0 = w * 32; n0 < N; n0 += 8 * 32) {
    const int cnt = min(32, N - n0);
    __syncwarp();
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 8) {
      float tk[8], tv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool in = j0 + j < cnt;
        tk[j] = in ? kb[static_cast<long long>(n0 + j0 + j) * ld] : 0.f;
        tv[j] = in ? vb[static_cast<long long>(n0 + j0 + j) * ld] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ks[(j0 + j) * 33 + lane] = tk[j];
        vs[(j0 + j) * 33 + lane] = tv[j];
      }
    }
    __syncwarp();
    float s[32], bm = m;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        a0 = fmaf(qr[c], ks[j * 33 + c], a0);
        a1 = fmaf(qr[c + 1], ks[j * 33 + c + 1], a1);
      }
      s[j] = j < cnt ? a0 + a1 : -INFINITY;
      bm = fmaxf(bm, s[j]);
    }
    const float scale = __expf(m - bm);  // 0 for the first block (m = -inf)
    l *= scale;
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] *= scale;
    m = bm;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float pj = __expf(s[j] - m);  // exp(-inf) = 0 for the padded keys
      l += pj;
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = fmaf(pj, vs[j * 33 + c], acc[c]);
    }
  }
  __syncthreads();
  // merge the eight warps: red[w][34][32 lanes] = (m, l, acc[32])
  float* red = sm;
  red[(w * 34 + 0) * 32 + lane] = m;
  red[(w * 34 + 1) * 32 + lane] = l;
#pragma unroll
  for (int c = 0; c < 32; ++c) red[(w * 34 + 2 + c) * 32 + lane] = acc[c];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    const int qi = i >> 5, c = i & 31;
    float M = -INFINITY;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) M = fmaxf(M, red[(ww * 34) * 32 + qi]);
    float L = 0.f, o = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      const float mw = red[(ww * 34) * 32 + qi];
      const float f = mw == -INFINITY ? 0.f : __expf(mw - M);
      L = fmaf(red[(ww * 34 + 1) * 32 + qi], f, L);
      o = fmaf(red[(ww * 34 + 2 + c) * 32 + qi], f, o);
    }
    out[(static_cast<long long>(b) * 32 + qi) * 256 + h * 32 + c] = o / L;
  }
}

// Merge the point splits of ctx_attn_kernel: part (B, S, 256, 256) fp32 holds softmax(q K_s^T) M_s normalised inside split s,
// lse (B, S, 256) the log2 of that split's sum of 2^score.  out[b][q] = sum_s w_s part[b][s][q], w_s = 2^(lse_s - max) / sum.
// One warp per (segment, folded query); lane = 8 columns.  The split weights are computed lane-parallel (lane = split) and
// handed round by shuffles; the part rows of four splits are in flight together.
__global__ void __launch_bounds__(256) ctx_merge_kernel(const float* __restrict__ part, const float* __restrict__ lse, int rows /* B*256 */,
                                                        int S, void* __restrict__ out, int out_bf16) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const int b = r >> 8, q = r & 255;
    const float* l = lse + static_cast<long long>(b) * S * 256 + q;
    float mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmaxf(mx, l[s * 256]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float den = 0.f;
    for (int s = lane; s < S; s += 32) den += exp2f(l[s * 256] - mx);
#pragma unroll
    for (int o = 16; o; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    const float inv = 1.f / den;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float* prow = part + (static_cast<long long>(b) * S * 256 + q) * 256 + 8 * lane;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const float wl = s0 + lane < S ? exp2f(l[(s0 + lane) * 256] - mx) * inv : 0.f;  // lane = split s0 + lane
      const int cnt = min(32, S - s0);
      for (int t = 0; t < cnt; t += 4) {
        float4 u[4], v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int s = min(s0 + t + j, S - 1);
          const float4* p = reinterpret_cast<const float4*>(prow + static_cast<long long>(s) * 65536);
          u[j] = p[0];
          v[j] = p[1];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float ws = t + j < cnt ? __shfl_sync(0xffffffffu, wl, (t + j) & 31) : 0.f;
          acc[0] = fmaf(ws, u[j].x, acc[0]); acc[1] = fmaf(ws, u[j].y, acc[1]); acc[2] = fmaf(ws, u[j].z, acc[2]); acc[3] = fmaf(ws, u[j].w, acc[3]);
          acc[4] = fmaf(ws, v[j].x, acc[4]); acc[5] = fmaf(ws, v[j].y, acc[5]); acc[6] = fmaf(ws, v[j].z, acc[6]); acc[7] = fmaf(ws, v[j].w, acc[7]);
        }
      }
    }
    const long long o = static_cast<long long>(r) * 256 + 8 * lane;
    if (out_bf16) {
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + o) =
          make_uint4(ptx::pack_bf16x2(acc[0], acc[1]), ptx::pack_bf16x2(acc[2], acc[3]), ptx::pack_bf16x2(acc[4], acc[5]),
                     ptx::pack_bf16x2(acc[6], acc[7]));
    } else {
      float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o);
      d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

}  // namespace lrn
