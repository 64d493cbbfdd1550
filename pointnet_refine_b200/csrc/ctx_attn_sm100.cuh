// Cross attention of the 32 polyline queries of a segment over its N context points (SURVEY.md §8f row 1;
// reference: nn.MultiheadAttention cross_attn in DetrTransformerDecoderLayer.forward, src/model.py:123-128),
// without ever materialising K or V:
//
//   scores_h = (q_h / sqrt(32)) K_h^T,  K_h = (memory + pos) Wk_h^T + bk_h
//            = [(q_h Wk_h) / sqrt(32)] (memory + pos)^T + const(row)          -- the bias term is softmax-invariant
//   out_h    = softmax(scores_h) V_h,   V_h = memory Wv_h^T + bv_h
//            = [softmax(scores_h) memory] Wv_h^T + bv_h                       -- rows of the softmax sum to 1
//
// so the 8 heads x 32 queries of one segment become 256 "folded" queries of width 256 (the host folds Wk and the
// 1/sqrt(32) * log2(e) scale into them), and the kernel is one flash-style pass over the segment's points:
//
//   S = Qf Kp^T        tcgen05.mma cta_group::2, 256 folded queries x 128 points x K = 256 (Qf resident in smem)
//   P = exp2(S - m)    one thread per folded query (TMEM lane), lazily rescaled running maximum
//   O += P Mem         tcgen05.mma with P (bf16) as the A operand in TENSOR MEMORY, Mem as an MN-major smem operand
//
// Per layer the only HBM traffic is Kp = memory + pos and Mem = memory (bf16, 1 KB per point) plus 256 KB of output per
// segment.  Work item = (segment, split): a segment can be split over several items (few segments), the host merges the
// partial results with their log-sum-exp.  The kernel is persistent - one cluster of two CTAs per SM pair loops over the
// items with one running step counter for all pipeline barriers, so the next item's folded queries and first point tiles
// load, and its first S products run, while the softmax warps still write the previous item's output.
#pragma once
#include "ptx.cuh"
#include "chain_pair_sm100.cuh"  // tc_mma_ts_pair

namespace lrn {

constexpr int kAttnThreads = 320;  // warp 0 TMA, warp 1 MMA / TMEM alloc, warps 2..9 softmax: two warps per TMEM lane quarter,
                                   // each thread = one folded query x one half (64 points) of the step
constexpr int kAttnStep = 128;     // points per step

struct AttnSmem {
  static constexpr uint32_t kQ = 0;                   // 4 k-blocks x [128 folded queries x 128 B]
  static constexpr uint32_t kKp = kQ + 65536;         // 2 stages x 4 k-blocks x [64 points x 128 B]
  static constexpr uint32_t kV = kKp + 2 * 32768;     // 2 stages x 2 point-blocks x 2 dim-blocks x [64 points x 128 B]
  static constexpr uint32_t kScratch = kV + 2 * 32768;  // 4 warps x [32][33] floats (output transpose)
  static constexpr uint32_t kXchg = kScratch + 4 * 32 * 33 * 4;   // [2 halves][128 queries] floats: row maximum / row sum exchange
  static constexpr uint32_t kBar = kXchg + 2 * 128 * 4;
  static constexpr uint32_t kTmemPtr = kBar + 16 * 8;
  static constexpr uint32_t kDynamic = kTmemPtr + 16 + 1024;
};

struct AttnParams {
  int items;      // work items = segments * splits
  int N;          // points per segment
  int splits;     // clusters per segment
  int steps_per_split;
  void* out;      // (B * splits * 256, 256) fp32 (or bf16 if out_bf16): normalised partial attention over memory
  int out_bf16;
  float* lse;     // (B * splits * 256) fp32: log2-sum-exp2 of the scaled scores of the split
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAttnThreads, 1)
ctx_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKp,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using L = AttnSmem;
  constexpr uint32_t kIdescS = ptx::make_idesc(false, 256, 128);
  constexpr uint32_t kIdescO = ptx::make_idesc(false, 256, 256) | (1u << 16);  // B operand (memory) MN-major

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* q_full = bars;         // leader: folded queries of both CTAs landed
  uint64_t* kp_full = bars + 1;    // [2] leader: Kp stage landed (both CTAs' halves)
  uint64_t* v_full = bars + 3;     // [2] leader: Mem stage landed
  uint64_t* s_full = bars + 5;     // [2] both: S buffer written by the tensor pipe (also: Kp stage consumed)
  uint64_t* p_ready = bars + 7;    // [2] leader: P written to tensor memory by the softmax warps of both CTAs
  uint64_t* o_done = bars + 9;     // [2] both: O += P Mem of that step finished (also: Mem stage and S buffer free)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int total_steps = (p.N + kAttnStep - 1) / kAttnStep;
  // item -> (segment, first step, number of steps); every role walks the same item sequence
  auto item_seg = [&](int item) { return item / p.splits; };
  auto item_step0 = [&](int item) { return (item % p.splits) * p.steps_per_split; };
  auto item_steps = [&](int item) { return min(p.steps_per_split, total_steps - item_step0(item)); };  // >= 1 by construction

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmKp);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&kp_full[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_ready[i], 16);
      ptx::mbar_init(&o_done[i], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair<512>(tmem_ptr);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_o = tmem_base + 256;

  // TMA and tcgen05 instructions are issued from warp-converged loops under elect.sync (uniform-register operands,
  // no per-lane serialisation loops): see gemm_pair_sm100.cuh / tools/mma_probe.cu.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t q_leader = ptx::mapa(ptx::smem_u32(q_full), 0);
    const uint32_t kp_leader0 = ptx::mapa(ptx::smem_u32(&kp_full[0]), 0);
    const uint32_t v_leader0 = ptx::mapa(ptx::smem_u32(&v_full[0]), 0);
    int g = 0;  // running step counter over all items of this cluster: stage = g & 1, phase = (g >> 1) & 1
    for (int item = cluster_id; item < p.items; item += num_clusters) {
      const int seg = item_seg(item), step0 = item_step0(item), steps = item_steps(item);
      const int64_t seg_row0 = static_cast<int64_t>(seg) * p.N;
      // the previous item's last S product has finished reading the folded queries
      if (g >= 1) ptx::mbar_wait(&s_full[(g - 1) & 1], ((g - 1) >> 1) & 1);
      if (ptx::elect_one()) {
        if (leader) ptx::mbar_arrive_expect_tx(q_full, 2 * 65536);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
          ptx::tma_load_2d_pair(smem + L::kQ + kb * 16384, &tmQ, q_leader, kb * 64, seg * 256 + static_cast<int>(rank) * 128);
      }
      __syncwarp();
      for (int j = 0; j < steps; ++j, ++g) {
        const int st = g & 1;
        const uint32_t prev = ((g - 2) >> 1) & 1;  // parity of the phase that step g - 2 completed
        const int row = static_cast<int>(seg_row0) + (step0 + j) * kAttnStep;
        // Kp: this CTA's 64 points of the step (the B operand of S is split by points across the pair)
        if (g >= 2) ptx::mbar_wait(&s_full[st], prev);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(&kp_full[st], 2 * 32768);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
            ptx::tma_load_2d_pair(smem + L::kKp + st * 32768 + kb * 8192, &tmKp, kp_leader0 + 8u * st, kb * 64,
                                  row + static_cast<int>(rank) * 64);
        }
        __syncwarp();
        // Mem: all 128 points, this CTA's 128 of the 256 output dims (the B operand of O is split by dims)
        if (g >= 2) ptx::mbar_wait(&o_done[st], prev);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(&v_full[st], 2 * 32768);
#pragma unroll
          for (int pb = 0; pb < 2; ++pb)
#pragma unroll
            for (int db = 0; db < 2; ++db)
              ptx::tma_load_2d_pair(smem + L::kV + st * 32768 + pb * 16384 + db * 8192, &tmV, v_leader0 + 8u * st,
                                    static_cast<int>(rank) * 128 + db * 64, row + pb * 64);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (leader) {
      auto issue_s = [&](int g) {  // S = Qf Kp^T of running step g -> S buffer g & 1
        const int st = g & 1;
        ptx::mbar_wait(&kp_full[st], (g >> 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d = tmem_base + st * 128;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t da = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + L::kQ + kb * 16384));
            const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + L::kKp + st * 32768 + kb * 8192));
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::tc_mma_ss_pair<false>(d, da + 2 * k, db + 2 * k, kIdescS, (kb | k) ? 1u : 0u);
          }
          ptx::tc_commit_pair(&s_full[st], 3);
        }
        __syncwarp();
      };
      int g0 = 0, it = 0;
      for (int item = cluster_id; item < p.items; item += num_clusters, ++it) {
        const int steps = item_steps(item);
        ptx::mbar_wait(q_full, it & 1);
        issue_s(g0);
        if (steps > 1) issue_s(g0 + 1);
        for (int j = 0; j < steps; ++j) {
          const int g = g0 + j, st = g & 1;
          const uint32_t par = (g >> 1) & 1;
          // P_j written by the softmax warps of both CTAs (for j = 0 this also says: the previous item's O was read out)
          ptx::mbar_wait(&p_ready[st], par);
          ptx::mbar_wait(&v_full[st], par);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {  // O += P_j Mem_j : A = P_j in the first 64 columns of S buffer st (2 points per column)
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) {
              const uint64_t db = ptx::make_smem_desc_sw128_mn(ptx::smem_u32(smem + L::kV + st * 32768 + pb * 16384), 8192);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_ts_pair(tmem_o, tmem_base + st * 128 + pb * 32 + k * 8, db + 128 * k, kIdescO, (j | pb | k) ? 1u : 0u);
            }
            ptx::tc_commit_pair(&o_done[st], 3);
          }
          __syncwarp();
          if (j + 2 < steps) issue_s(g + 2);  // in order behind O_j on the tensor pipe, which is still reading P_j
        }
        g0 += steps;
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps (both CTAs)
    // Two warps share a TMEM lane quarter (32 folded queries): warp `half` owns points [64 half, 64 half + 64) of every
    // step, loads its 64 scores with two tensor-memory loads in flight and ONE wait, and works from registers.  (The first
    // version - one warp per quarter, eight serialised load / wait round trips per step - made the softmax the limiter:
    // 4.6k cycles per step against 2k of tensor work, 3.6 TB/s.  Now the kernel reads Kp / Mem at 5.2 TB/s, 0.79 of the
    // measured copy bandwidth: HBM-bound for its 1 KB per point and layer.)
    // The row maximum is exchanged through shared memory (one 64-thread named barrier per step); each warp keeps its own
    // partial row sum, added up once per item.
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which 64 of the step's 128 points
    const int row = q * 32 + lane;     // folded query of this thread within the CTA's 128
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t ready_leader[2] = {ptx::mapa(ptx::smem_u32(&p_ready[0]), 0), ptx::mapa(ptx::smem_u32(&p_ready[1]), 0)};
    float* xchg = reinterpret_cast<float*>(smem + L::kXchg);
    int g0 = 0;
    for (int item = cluster_id; item < p.items; item += num_clusters) {
      const int step0 = item_step0(item), steps = item_steps(item);
      float m_ref = -INFINITY, l_sum = 0.f;
      for (int j = 0; j < steps; ++j) {
        const int g = g0 + j;
        const int st = g & 1;
        const uint32_t t_s = t_lane + st * 128;
        const int valid = min(kAttnStep, p.N - (step0 + j) * kAttnStep) - 64 * half;  // valid points of this half (may be <= 0)
        ptx::mbar_wait(&s_full[st], (g >> 1) & 1);
        ptx::tc_fence_after();
        // this thread's 64 scores: two loads in flight, one wait; everything below works from registers
        uint32_t sr[64];
        ptx::tmem_ld_32x32b_x32(t_s + 64 * half, reinterpret_cast<uint32_t(&)[32]>(sr[0]));
        ptx::tmem_ld_32x32b_x32(t_s + 64 * half + 32, reinterpret_cast<uint32_t(&)[32]>(sr[32]));
        ptx::tmem_ld_wait();
        if (valid < 64) {  // last step of a ragged segment: points past N never win the maximum and get p = 0
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) sr[i] = 0xff800000u;
        }
        float m_t;
        {
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(sr[i]));
            m1 = fmaxf(m1, __uint_as_float(sr[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(sr[i + 2]));
            m3 = fmaxf(m3, __uint_as_float(sr[i + 3]));
          }
          m_t = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        }
        xchg[half * 128 + row] = m_t;
        ptx::named_bar_sync(1 + q, 64);
        m_t = fmaxf(m_t, xchg[(half ^ 1) * 128 + row]);
        ptx::named_bar_sync(1 + q, 64);   // both warps have read before the next step overwrites
        // Lazy rescale: the reference maximum only moves when the row maximum grew by more than 2^8 (p <= 256 stays
        // exact enough in bf16 / fp32); then this row of O and its running sum are scaled down once.  Both warps of a
        // quarter see the same m_t / m_ref and take the same decision; each rescales half of the 256 output columns.
        const bool grow = m_t > m_ref + 8.f;
        if (j == 0) {
          m_ref = m_t;
        } else if (__any_sync(0xffffffffu, grow)) {
          const float f = grow ? ex2_approx(m_ref - m_t) : 1.f;
          ptx::mbar_wait(&o_done[st ^ 1], ((g - 1) >> 1) & 1);  // O_{j-1} has been accumulated
          ptx::tc_fence_after();
#pragma unroll 1
          for (int c = 4 * half; c < 4 * half + 4; ++c) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tmem_o + (static_cast<uint32_t>(q * 32) << 16) + 32 * c, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            ptx::tmem_st_32x32b_x32(tmem_o + (static_cast<uint32_t>(q * 32) << 16) + 32 * c, r);
          }
          l_sum *= f;
          if (grow) m_ref = m_t;
        }
        // P = exp2(S - m_ref) as bf16 pairs: packed column c of the S buffer = points 2c, 2c + 1; this warp writes packed
        // columns [32 half, 32 half + 32).  (The other warp's scores in those columns are already in ITS registers: the
        // second named barrier above comes after both warps' loads.)
        {
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
          uint32_t o[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = ex2_approx(__uint_as_float(sr[2 * i]) - m_ref), e1 = ex2_approx(__uint_as_float(sr[2 * i + 1]) - m_ref);
            const float e2 = ex2_approx(__uint_as_float(sr[2 * i + 2]) - m_ref), e3 = ex2_approx(__uint_as_float(sr[2 * i + 3]) - m_ref);
            l0 += e0; l1 += e1; l2 += e2; l3 += e3;
            o[i] = ptx::pack_bf16x2(e0, e1);
            o[i + 1] = ptx::pack_bf16x2(e2, e3);
          }
          ptx::tmem_st_32x32b_x32(t_s + 32 * half, o);
          l_sum += (l0 + l1) + (l2 + l3);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ready_leader[st]);
      }
      // ---- epilogue: O / l -> global (through a per-warp transpose so that rows are written 128 B at a time); the two
      //      warps of a quarter add their partial row sums and take four of the eight 32-column chunks each
      {
        const int last = g0 + steps - 1;
        xchg[half * 128 + row] = l_sum;
        ptx::named_bar_sync(1 + q, 64);
        l_sum += xchg[(half ^ 1) * 128 + row];
        ptx::named_bar_sync(1 + q, 64);
        ptx::mbar_wait(&o_done[last & 1], (last >> 1) & 1);
        ptx::tc_fence_after();
        const float inv = 1.f / l_sum;
        float* scratch = reinterpret_cast<float*>(smem + L::kScratch) + q * 32 * 33;   // shared by the two warps of a quarter, in turn
        const int64_t out_row0 = (static_cast<int64_t>(item) * 256 + rank * 128 + q * 32);
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          if ((c >> 2) == half) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tmem_o + (static_cast<uint32_t>(q * 32) << 16) + 32 * c, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) scratch[lane * 33 + i] = __uint_as_float(r[i]) * inv;
            __syncwarp();
            if (p.out_bf16) {  // two rows per instruction, 16 lanes x bf16 pair = 64 contiguous bytes per row
              uint32_t* o16 = reinterpret_cast<uint32_t*>(p.out);
              const int hh = lane >> 4, l = lane & 15;
#pragma unroll 4
              for (int rr = 0; rr < 32; rr += 2) {
                const float* src = scratch + (rr + hh) * 33 + 2 * l;
                o16[((out_row0 + rr + hh) * 256 + 32 * c) / 2 + l] = ptx::pack_bf16x2(src[0], src[1]);
              }
            } else {
              float* o32 = reinterpret_cast<float*>(p.out);
#pragma unroll 4
              for (int rr = 0; rr < 32; ++rr) o32[(out_row0 + rr) * 256 + 32 * c + lane] = scratch[rr * 33 + lane];
            }
            __syncwarp();
          }
          if (c == 3) ptx::named_bar_sync(1 + q, 64);   // hand the quarter's scratch from warp `half 0` to warp `half 1`
        }
        if (half == 0) p.lse[out_row0 + lane] = m_ref + log2f(l_sum);
      }
      g0 += steps;
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<512>(tmem_base);
  }
}

}  // namespace lrn
