// Query side of the decoder: the small per-polyline-point operations between the tensor-core
// linears of DetrTransformerDecoderLayer.forward (reference src/model.py:104-135), each as one HBM-bound kernel:
// residual add + LayerNorm, the 32 x 32 self attention of one segment's polyline points (8 heads x 32), and the second
// layer of a regression head with the cumulative-offset update (src/model.py:220,227-231).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lrn {

// out[r, :] = LayerNorm(x[r, :] + y[r, :]) * gamma + beta over 256 columns (norm1/2/3, eps 1e-5, biased variance).
// One warp per row; lane l holds columns [4l, 4l+4) and [128+4l, 128+4l+4).  y may be null.
// `stats` (optional, (rows, 2) = mean, rstd) is what the backward kernel needs besides x + y.
__global__ void __launch_bounds__(256)
add_layernorm256_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, float* __restrict__ out, float* __restrict__ stats,
                        long long rows) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[lane], g1 = reinterpret_cast<const float4*>(gamma)[32 + lane];
  const float4 b0 = reinterpret_cast<const float4*>(beta)[lane], b1 = reinterpret_cast<const float4*>(beta)[32 + lane];
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * 256);
    float4 a = xr[lane], b = xr[32 + lane];
    if (y) {
      const float4* yr = reinterpret_cast<const float4*>(y + r * 256);
      const float4 c = yr[lane], d = yr[32 + lane];
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    float s = (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 256.f);
    a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
    b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
    float q = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w) + (b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w);
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / 256.f) + eps);
    if (stats && lane == 0) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
    float4* orow = reinterpret_cast<float4*>(out + r * 256);
    orow[lane] = make_float4(a.x * rstd * g0.x + b0.x, a.y * rstd * g0.y + b0.y, a.z * rstd * g0.z + b0.z, a.w * rstd * g0.w + b0.w);
    orow[32 + lane] = make_float4(b.x * rstd * g1.x + b1.x, b.y * rstd * g1.y + b1.y, b.z * rstd * g1.z + b1.z, b.w * rstd * g1.w + b1.w);
  }
}

// Backward of out = LayerNorm(x + y) * gamma + beta over 256 columns.  With z = x + y, xh = (z - mean) * rstd, g = gamma * dy:
//   dz = rstd * (g - mean_c(g) - xh * mean_c(g * xh))    (the gradient of both x and y)
//   dgamma += sum_rows dy * xh,  dbeta += sum_rows dy    (register partial sums per warp, reduced per block)
__global__ void __launch_bounds__(256)
add_layernorm256_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                            const float* __restrict__ stats, const float* __restrict__ gamma, float* __restrict__ dz,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[lane], g1 = reinterpret_cast<const float4*>(gamma)[32 + lane];
  float ag[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ab[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    float4 a = reinterpret_cast<const float4*>(x + r * 256)[lane], b = reinterpret_cast<const float4*>(x + r * 256)[32 + lane];
    if (y) {
      const float4 c = reinterpret_cast<const float4*>(y + r * 256)[lane], d = reinterpret_cast<const float4*>(y + r * 256)[32 + lane];
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    const float4 d0 = reinterpret_cast<const float4*>(dy + r * 256)[lane], d1 = reinterpret_cast<const float4*>(dy + r * 256)[32 + lane];
    const float xh[8] = {(a.x - mean) * rstd, (a.y - mean) * rstd, (a.z - mean) * rstd, (a.w - mean) * rstd,
                         (b.x - mean) * rstd, (b.y - mean) * rstd, (b.z - mean) * rstd, (b.w - mean) * rstd};
    const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float g[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] = gm[i] * dv[i];
      s1 += g[i];
      s2 += g[i] * xh[i];
      ag[i] += dv[i] * xh[i];
      ab[i] += dv[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.f / 256.f);
    s2 *= (1.f / 256.f);
    float o8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o8[i] = rstd * (g[i] - s1 - xh[i] * s2);
    reinterpret_cast<float4*>(dz + r * 256)[lane] = make_float4(o8[0], o8[1], o8[2], o8[3]);
    reinterpret_cast<float4*>(dz + r * 256)[32 + lane] = make_float4(o8[4], o8[5], o8[6], o8[7]);
  }
  // block-level reduction of the 8 warps' partial sums, then one atomicAdd per block and column
  __shared__ float red[8][512];
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[w][4 * lane + i] = ag[i];
    red[w][128 + 4 * lane + i] = ag[4 + i];
    red[w][256 + 4 * lane + i] = ab[i];
    red[w][384 + 4 * lane + i] = ab[4 + i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 512; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][c];
    atomicAdd((c < 256 ? dgamma : dbeta - 256) + c, t);
  }
}

// Self attention over the 32 polyline points of one segment, 8 heads of 32 channels (nn.MultiheadAttention self_attn in
// eval mode, src/model.py:113-117): qk (B*32, 512) = [q | k] projections, v (B*32, 256) -> out (B*32, 256), heads
// concatenated.  One block per segment, warp = head, lane = query; K / V of the head in shared memory (broadcast reads).
constexpr int kSelfAttnSmem = 8 * 3 * 32 * 33 * 4;
__global__ void __launch_bounds__(256)
self_attn32_kernel(const float* __restrict__ qk, const float* __restrict__ v, float* __restrict__ out, int B) {
  extern __shared__ float sm[];
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = sm + h * 3 * 32 * 33;
  float* ks = qs + 32 * 33;
  float* vs = ks + 32 * 33;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const long long row0 = static_cast<long long>(b) * 32;
    __syncwarp();
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 8) {  // coalesced 128-byte rows, 24 loads in flight per lane (one segment per block: latency-bound)
      float tq[8], tk[8], tv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        tq[j] = qk[(row0 + j0 + j) * 512 + h * 32 + lane];
        tk[j] = qk[(row0 + j0 + j) * 512 + 256 + h * 32 + lane];
        tv[j] = v[(row0 + j0 + j) * 256 + h * 32 + lane];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        qs[(j0 + j) * 33 + lane] = tq[j];
        ks[(j0 + j) * 33 + lane] = tk[j];
        vs[(j0 + j) * 33 + lane] = tv[j];
      }
    }
    __syncwarp();
    float q[32], s[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) q[c] = qs[lane * 33 + c] * 0.17677669529663688f;  // 1 / sqrt(32)
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four independent chains: one segment per block is latency-bound
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        a0 = fmaf(q[c], ks[j * 33 + c], a0);
        a1 = fmaf(q[c + 1], ks[j * 33 + c + 1], a1);
        a2 = fmaf(q[c + 2], ks[j * 33 + c + 2], a2);
        a3 = fmaf(q[c + 3], ks[j * 33 + c + 3], a3);
      }
      const float acc = (a0 + a1) + (a2 + a3);
      s[j] = acc;
      m = fmaxf(m, acc);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = expf(s[j] - m);
      l += s[j];
    }
    const float inv = 1.f / l;
    float o[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float pj = s[j] * inv;
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] = fmaf(pj, vs[j * 33 + c], o[c]);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 32; ++c) qs[lane * 33 + c] = o[c];  // stage through shared memory for coalesced rows
    __syncwarp();
#pragma unroll 4
    for (int j = 0; j < 32; ++j) out[(row0 + j) * 256 + h * 32 + lane] = qs[j * 33 + lane];
  }
}

// Second layer of a regression head + cumulative-offset bookkeeping (src/model.py:220,227-231):
//   delta = hidden[r, :] W2^T + b2 (128 -> 3);  current[r] += delta;  cum[r] = current[r] - noisy[r].  One warp per row.
__global__ void __launch_bounds__(256)
head_update_kernel(const float* __restrict__ hidden, const float* __restrict__ w2 /* (3,128) */, const float* __restrict__ b2,
                   long long rows, float* __restrict__ current, const float* __restrict__ noisy, float* __restrict__ cum) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float4 wx = reinterpret_cast<const float4*>(w2)[lane], wy = reinterpret_cast<const float4*>(w2 + 128)[lane],
               wz = reinterpret_cast<const float4*>(w2 + 256)[lane];
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float4 hv = reinterpret_cast<const float4*>(hidden + r * 128)[lane];
    float dx = hv.x * wx.x + hv.y * wx.y + hv.z * wx.z + hv.w * wx.w;
    float dy = hv.x * wy.x + hv.y * wy.y + hv.z * wy.z + hv.w * wy.w;
    float dz = hv.x * wz.x + hv.y * wz.y + hv.z * wz.z + hv.w * wz.w;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      dx += __shfl_xor_sync(0xffffffffu, dx, o);
      dy += __shfl_xor_sync(0xffffffffu, dy, o);
      dz += __shfl_xor_sync(0xffffffffu, dz, o);
    }
    if (lane < 3) {
      const float d = (lane == 0 ? dx : lane == 1 ? dy : dz) + b2[lane];
      const float c = current[r * 3 + lane] + d;
      current[r * 3 + lane] = c;
      cum[r * 3 + lane] = c - noisy[r * 3 + lane];
    }
  }
}

}  // namespace lrn
