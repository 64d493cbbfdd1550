// Probe kernel (tuning / validation aid, not on the product path): D[128 x N] = A[128 x K] * W[N x K]^T with the
// A operand read from TENSOR MEMORY (tcgen05.mma "TS" form), to pin down the TMEM layout of a bf16 A operand:
// row m = lane m, K elements packed two per 32-bit column, element 2j in the low half.
#pragma once
#include "ptx.cuh"

namespace lrn {

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// one CTA, 128 threads; K = 64, N = 64; A (128, 64) bf16 row-major, W via tensor map (box 64 x 64, 128B swizzle)
__global__ void __launch_bounds__(128, 1)
ts_probe_kernel(const __grid_constant__ CUtensorMap tmW, const uint32_t* __restrict__ a_packed /* (128, 32) u32 */,
                float* __restrict__ out /* (128, 64) */) {
  __shared__ __align__(1024) uint8_t sW[64 * 128];
  __shared__ __align__(8) uint64_t bar_w, bar_d;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_w, 1);
    ptx::mbar_init(&bar_d, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc<128>(&tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_w, 64 * 128);
    ptx::tma_load_2d(sW, &tmW, &bar_w, 0, 0);
  }
  // A: thread = row; 32 packed registers (64 bf16) -> TMEM columns [64, 96) of its lane
  uint32_t r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = a_packed[threadIdx.x * 32 + j];
  ptx::tmem_st_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 64, r);
  ptx::tmem_st_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    ptx::mbar_wait(&bar_w, 0);
    ptx::tc_fence_after();
    const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(sW));
    const uint32_t idesc = ptx::make_idesc(false, 128, 64);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_ts(tmem, tmem + 64 + 8 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
    ptx::tc_commit(&bar_d);
  }
  ptx::mbar_wait(&bar_d, 0);
  ptx::tc_fence_after();
  uint32_t d[32];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    ptx::tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 32 * h, d);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 64 + 32 * h + j] = __uint_as_float(d[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<128>(tmem);
  }
}

}  // namespace lrn
