// CUDA-core kernels around the tensor-core GEMMs: BatchNorm folding / packing, the fp32 first
// layer (conv1 + gate layer 1) that turns raw points into the first operand columns, the argmax
// finalisation, and the regression head.
#pragma once
#include "ptx.cuh"

namespace lrn {

// ---------------------------------------------------------------------------------------------
// Fold eval-mode BatchNorm into a 1x1 conv and write the K-major operand matrix.
//   W'[co,ci] = W[co,ci] * s[co],  b'[co] = (b[co] - mean[co]) * s[co] + beta[co],  s = g / sqrt(var + eps)
// (src/model.py:43-47,51).  g == nullptr means "no BatchNorm" (gate layer 2, context_proj).
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void fold_linear_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                   const float* __restrict__ g, const float* __restrict__ beta,
                                   const float* __restrict__ mean, const float* __restrict__ var, float eps, int cout,
                                   int cin, OutT* __restrict__ out_w, long long ld, int col0,
                                   float* __restrict__ out_b, int rna_tf32, long long lo_off = 0) {
  const long long total = static_cast<long long>(cout) * cin;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i / cin);
    const int ci = static_cast<int>(i - static_cast<long long>(co) * cin);
    const float s = g ? g[co] / sqrtf(var[co] + eps) : 1.f;
    const float v = w[i] * s;
    if constexpr (sizeof(OutT) == 2) {
      out_w[co * ld + col0 + ci] = __float2bfloat16_rn(v);
    } else if (lo_off) {  // fp32x3 tier: hi / lo TF32 halves of the folded weight
      const float h = ptx::round_tf32(v);
      out_w[co * ld + col0 + ci] = h;
      out_w[co * ld + lo_off + col0 + ci] = ptx::round_tf32(v - h);
    } else {
      out_w[co * ld + col0 + ci] = rna_tf32 ? ptx::round_tf32(v) : v;
    }
    if (ci == 0 && out_b) {
      out_b[co] = g ? (b[co] - mean[co]) * s + beta[co] : b[co];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// First layer, fp32 FMA (raw xyz in metres and raw intensity counts must not be rounded to bf16):
//   feat1 = relu(W1' [x y z I] + b1')   -> operand columns [0, 64)        src/model.py:43
//   hgate = relu(wg1 * I + bg1)         -> operand columns [1984, 2048)   src/model.py:33-34
// 16 threads per point, each produces 4 + 4 channels, so every 16-thread group writes two
// contiguous runs (128 B bf16 / 256 B fp32) of the point's operand row.  HBM-bound: 16 B in.
// ---------------------------------------------------------------------------------------------
struct EmbedWeights {
  const float* w1;   // (64, 4) folded
  const float* b1;   // (64)
  const float* wg1;  // (64)
  const float* bg1;  // (64)
};

// ld > 0: row-major operand rows of pitch ld.  ld == 0 (bf16 only): the tiled layout of the default path,
// [row tile of 128][column block of 64][128 rows][64 columns] - feat1 is column block 0, the gate hidden block is block 31.
template <bool TF32>
__global__ void __launch_bounds__(256) point_embed_kernel(const float4* __restrict__ ctx, long long rows,
                                                          EmbedWeights w, void* __restrict__ cat, int ld, int lo_off = 0) {
  const int sub = threadIdx.x & 15;  // channel group: channels [4*sub, 4*sub + 4)
  float4 wr[4];
  float br[4], gw[4], gb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    wr[j] = reinterpret_cast<const float4*>(w.w1)[4 * sub + j];
    br[j] = w.b1[4 * sub + j];
    gw[j] = w.wg1[4 * sub + j];
    gb[j] = w.bg1[4 * sub + j];
  }
  const long long pts_per_block = blockDim.x / 16;
  const long long stride = gridDim.x * pts_per_block;
  auto emit = [&](long long pt, const float4& x) {
    float f[4], h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[j] = fmaxf(fmaf(wr[j].x, x.x, fmaf(wr[j].y, x.y, fmaf(wr[j].z, x.z, fmaf(wr[j].w, x.w, br[j])))), 0.f);
      h[j] = fmaxf(fmaf(gw[j], x.w, gb[j]), 0.f);
    }
    if constexpr (TF32) {
      float* row = reinterpret_cast<float*>(cat) + pt * ld;
      if (lo_off) {  // fp32x3 tier: lo = rna_tf32(v - hi) in the second half of the operand row
        float fl[4], hl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          fl[j] = ptx::round_tf32(f[j] - ptx::round_tf32(f[j]));
          hl[j] = ptx::round_tf32(h[j] - ptx::round_tf32(h[j]));
        }
        *reinterpret_cast<float4*>(row + lo_off + 4 * sub) = make_float4(fl[0], fl[1], fl[2], fl[3]);
        *reinterpret_cast<float4*>(row + lo_off + 1984 + 4 * sub) = make_float4(hl[0], hl[1], hl[2], hl[3]);
      }
      *reinterpret_cast<float4*>(row + 4 * sub) =
          make_float4(ptx::round_tf32(f[0]), ptx::round_tf32(f[1]), ptx::round_tf32(f[2]), ptx::round_tf32(f[3]));
      *reinterpret_cast<float4*>(row + 1984 + 4 * sub) =
          make_float4(ptx::round_tf32(h[0]), ptx::round_tf32(h[1]), ptx::round_tf32(h[2]), ptx::round_tf32(h[3]));
    } else {
      uint16_t* row = reinterpret_cast<uint16_t*>(cat) + (ld ? pt * ld : (pt >> 7) * (32ll * 8192) + (pt & 127) * 64);
      const long long gate_off = ld ? 1984 : 31ll * 8192;
      *reinterpret_cast<uint2*>(row + 4 * sub) = make_uint2(ptx::pack_bf16x2(f[0], f[1]), ptx::pack_bf16x2(f[2], f[3]));
      *reinterpret_cast<uint2*>(row + gate_off + 4 * sub) =
          make_uint2(ptx::pack_bf16x2(h[0], h[1]), ptx::pack_bf16x2(h[2], h[3]));
    }
  };
  // four points (loads) in flight per thread: the kernel is HBM-latency-bound otherwise
  long long pt = blockIdx.x * pts_per_block + (threadIdx.x >> 4);
  for (; pt + 3 * stride < rows; pt += 4 * stride) {
    float4 x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = __ldg(ctx + pt + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(pt + k * stride, x[k]);
  }
  for (; pt < rows; pt += stride) emit(pt, __ldg(ctx + pt));
}

// ---------------------------------------------------------------------------------------------
// First layer of the positional embedding on the context points (reference PositionalEncoding.mlp[0] + ReLU,
// src/model.py:66-75 applied to context[:, :, :3] at :197): out[pt][0:256) = bf16(relu(W1 xyz + b1)).
// One warp writes one point's 512-byte row; lane = 8 consecutive channels, their weights live in registers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pos_hidden_kernel(const float4* __restrict__ ctx, long long rows,
                                                         const float* __restrict__ w1 /* (256,3) */,
                                                         const float* __restrict__ b1, uint16_t* __restrict__ out, long long ld) {
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
  long long pt = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  auto emit = [&](long long q, const float4& x) {
    float h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = fmaxf(fmaf(wx[j], x.x, fmaf(wy[j], x.y, fmaf(wz[j], x.z, bb[j]))), 0.f);
    *reinterpret_cast<uint4*>(out + q * ld + 8 * lane) = make_uint4(ptx::pack_bf16x2(h[0], h[1]), ptx::pack_bf16x2(h[2], h[3]),
                                                                    ptx::pack_bf16x2(h[4], h[5]), ptx::pack_bf16x2(h[6], h[7]));
  };
  for (; pt + 3 * warps < rows; pt += 4 * warps) {
    float4 x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = __ldg(ctx + pt + k * warps);
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(pt + k * warps, x[k]);
  }
  for (; pt < rows; pt += warps) emit(pt, __ldg(ctx + pt));
}

// Backward of pos_hidden_kernel w.r.t. its parameters (train mode): with m = [h > 0],
//   dW1[c][k] += sum_p dh[p][c] * m * xyz[p][k],   db1[c] += sum_p dh[p][c] * m.
// Same thread layout (warp per point, lane = 8 channels); every warp keeps its partial sums in registers over all its
// points and issues one atomicAdd per (channel, term) at the end.  dW1 / db1 are zeroed by the caller.
__global__ void __launch_bounds__(256) pos_hidden_bwd_kernel(const float4* __restrict__ ctx, long long rows,
                                                             const uint16_t* __restrict__ h, long long ldh,
                                                             const uint16_t* __restrict__ dh, long long ldd,
                                                             float* __restrict__ dw1 /* (256,3) */, float* __restrict__ db1) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  float ax[8], ay[8], az[8], ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ax[j] = ay[j] = az[j] = ab[j] = 0.f;
  auto bf = [](uint32_t w, int hi) { return __uint_as_float(hi ? (w & 0xFFFF0000u) : (w << 16)); };
  for (long long pt = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; pt < rows; pt += warps) {
    const float4 x = __ldg(ctx + pt);
    const uint4 hv = *reinterpret_cast<const uint4*>(h + pt * ldh + 8 * lane);
    const uint4 dv = *reinterpret_cast<const uint4*>(dh + pt * ldd + 8 * lane);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = bf(hw[j >> 1], j & 1) > 0.f ? bf(dw[j >> 1], j & 1) : 0.f;
      ax[j] = fmaf(g, x.x, ax[j]);
      ay[j] = fmaf(g, x.y, ay[j]);
      az[j] = fmaf(g, x.z, az[j]);
      ab[j] += g;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = 8 * lane + j;
    atomicAdd(dw1 + 3 * c, ax[j]);
    atomicAdd(dw1 + 3 * c + 1, ay[j]);
    atomicAdd(dw1 + 3 * c + 2, az[j]);
    atomicAdd(db1 + c, ab[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// Unpack argmax keys: key = (float bits << 32) | (0xFFFFFFFF - n)  ->  max value and int64 index.
// ---------------------------------------------------------------------------------------------
__global__ void argmax_finalize_kernel(const unsigned long long* __restrict__ keys, long long B,
                                       float* __restrict__ global_feat, long long* __restrict__ argmax) {
  const long long total = B * 1024;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long k = keys[i];
    const long long b = i >> 10;
    const int c = static_cast<int>(i & 1023);
    global_feat[b * 2048 + c] = __uint_as_float(static_cast<unsigned int>(k >> 32));
    argmax[i] = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(k & 0xFFFFFFFFu));
  }
}

// ---------------------------------------------------------------------------------------------
// Regression head + cumulative-offset bookkeeping (src/model.py:172-179,220,227-231), fp32:
//   delta = W2 relu(W1 t + b1) + b2 ; current += delta ; cum = current - noisy
// One block handles kHeadRows rows; thread j owns hidden unit j for all of them.
// ---------------------------------------------------------------------------------------------
constexpr int kHeadRows = 32;

__global__ void __launch_bounds__(128) head_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                                   const float* __restrict__ tgt, long long rows,
                                                   float* __restrict__ current, const float* __restrict__ noisy,
                                                   float* __restrict__ cum_out) {
  __shared__ __align__(16) float st[kHeadRows][256];
  float(*sh)[129] = reinterpret_cast<float(*)[129]>(&st[0][0]);  // hidden activations reuse the input tile
  const long long r0 = static_cast<long long>(blockIdx.x) * kHeadRows;
  const int nrows = static_cast<int>(min(static_cast<long long>(kHeadRows), rows - r0));
  for (int i = threadIdx.x; i < kHeadRows * 64; i += 128) {
    const int r = i >> 6, c4 = i & 63;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows) v = __ldg(reinterpret_cast<const float4*>(tgt + (r0 + r) * 256) + c4);
    reinterpret_cast<float4*>(&st[r][0])[c4] = v;
  }
  __syncthreads();
  const int j = threadIdx.x;
  float acc[kHeadRows];
  const float bj = b1[j];
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) acc[r] = bj;
  const float4* wrow = reinterpret_cast<const float4*>(w1 + j * 256);
#pragma unroll 2
  for (int k4 = 0; k4 < 64; ++k4) {
    const float4 w = __ldg(wrow + k4);
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const float4 t = reinterpret_cast<const float4*>(&st[r][0])[k4];
      acc[r] = fmaf(w.x, t.x, fmaf(w.y, t.y, fmaf(w.z, t.z, fmaf(w.w, t.w, acc[r]))));
    }
  }
  __syncthreads();  // every thread is done reading st
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) sh[r][j] = fmaxf(acc[r], 0.f);
  __syncthreads();
  if (threadIdx.x < kHeadRows * 3) {
    const int r = threadIdx.x / 3, o = threadIdx.x - 3 * r;
    if (r < nrows) {
      float d = b2[o];
      for (int k = 0; k < 128; ++k) d = fmaf(w2[o * 128 + k], sh[r][k], d);
      const long long idx = (r0 + r) * 3 + o;
      const float c = current[idx] + d;
      current[idx] = c;
      cum_out[idx] = c - noisy[idx];
    }
  }
}

}  // namespace lrn
