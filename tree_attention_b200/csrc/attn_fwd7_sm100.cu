// attn_fwd7 -- EXPERIMENTAL (compile-checked only; not yet validated on hardware, see docs/NEXT.md): tcgen05 flash forward
// with 2-CTA MMAs (cta_group::2).
//
// Same contract as attn_fwd_sm100.cu's local mode (shard-local partial (o, lse); replaces /root/reference/model.py:74-80).
// A cluster of two CTAs owns 256 query rows of one (batch, q-head): CTA r keeps rows [m0 + 128 r, +128) -- its own Q tile in
// shared memory, its own S / P / O in TMEM, its own four softmax warps.  Every MMA is ONE tcgen05.mma.cta_group::2 with
// M = 256 issued by the leader CTA (cluster rank 0):
//     S = Q K^T   A = Q (each CTA's rows, same smem offset)     B = K tile, each CTA holding HALF of its 128 keys
//     O += P V    A = P from each CTA's TMEM                    B = V tile (MN-major), each CTA holding HALF of the channels
// so a CTA TMA-loads 32 KB per KV step instead of 64 KB and each SM reads half of the B operand from its shared memory:
// the forward is power-bound at 1 kW and operand traffic is the part of the energy a schedule can cut.
// Barriers: k_full / v_full / q_full / p_full live in the leader (both CTAs' TMA loads post their bytes there with the
// .cta_group::2 form; the peer's softmax warps arrive remotely); s_full / pv_done / k_empty / v_empty are signalled in BOTH
// CTAs by multicast commits.  head_dim 128 only, no fused multi-GPU mode.
#include "attn_fwd_common.cuh"

namespace ta {
namespace {
using namespace fwd_detail;

constexpr int kF7Threads = 192;
constexpr int kF7Stages = 3;
constexpr int kF7D = 128;
constexpr int kF7QBytes = kBlockM * kF7D * 2;           // 32 KB: this CTA's query tile
constexpr int kF7KHalf = (kBlockN / 2) * kF7D * 2;      // 16 KB: 64 keys x 128 channels
constexpr int kF7VHalf = kBlockN * (kF7D / 2) * 2;      // 16 KB: 128 keys x 64 channels
constexpr size_t kF7Smem = 1024 + kF7QBytes + size_t(kF7Stages) * (kF7KHalf + kF7VHalf) + 512;

__device__ __forceinline__ uint32_t f7_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void f7_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the LEADER's (even CTA's) copy of a barrier / buffer at the same offset
__device__ __forceinline__ uint32_t f7_leader_addr(const void* p) { return smem_u32(p) & 0xFEFFFFFFu; }
// both CTAs issue their half; the transaction bytes are posted to the leader's barrier
__device__ __forceinline__ void f7_tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(f7_leader_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void f7_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  const uint32_t z = 0;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z) : "memory");
}
__device__ __forceinline__ void f7_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  const uint32_t z = 0;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void f7_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the LEADER's barrier (local arrive in the leader, remote arrive from the peer)
__device__ __forceinline__ void f7_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(f7_leader_addr(bar)) : "memory");
}

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kF7Threads, 1)
attn_fwd7_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                 const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap omap,
                 const FwdParams p, const int num_pairs) {
  constexpr int D = kF7D;
  constexpr int NS = kF7Stages;
  constexpr int kAtomBytes = kBlockM * 128;        // one 64-channel atom column of a 128-row tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;                              // [2 atoms][128 rows][128 B]
  uint8_t* k_s = q_s + kF7QBytes;                   // [NS][2 atoms][64 keys][128 B]
  uint8_t* v_s = k_s + NS * kF7KHalf;               // [NS][128 keys][128 B] (64 channels)
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + NS * kF7VHalf);
  uint64_t* q_full = bars;            // leader
  uint64_t* k_full = bars + 1;        // leader, NS
  uint64_t* k_empty = k_full + NS;    // both
  uint64_t* v_full = k_empty + NS;    // leader
  uint64_t* v_empty = v_full + NS;    // both
  uint64_t* s_full = v_empty + NS;    // both, 2
  uint64_t* p_full = s_full + 2;      // leader, 2 (count 8)
  uint64_t* pv_done = p_full + 2;     // both, 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = f7_cluster_rank();
  const bool leader_cta = crank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int pair = num_pairs - 1 - (cluster_id % num_pairs);   // heaviest (causal) pairs first
  const int bh = cluster_id / num_pairs;
  const int hq = bh % p.Hq, b = bh / p.Hq;
  const int hkv = hq / p.G;
  const int m0 = (pair * 2 + (int)crank) * kBlockM;             // this CTA's 128 rows

  int n_end = p.S;
  if (p.causal) {
    const long long last_q = p.q_pos0 + min(pair * 2 * kBlockM + 2 * kBlockM - 1, p.Sq - 1);
    n_end = (int)max(0LL, min((long long)p.S, last_q - p.kv_pos0 + 1));
  }
  const int n_tiles = (n_end + kBlockN - 1) / kBlockN;   // identical in both CTAs of the cluster
  if (n_tiles == 0) {
    if (warp < 4) {
      const int row = m0 + tid;
      if (row < p.Sq) {
        uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)hq * p.o_sh + (long long)row * p.o_ss;
        for (int d = 0; d < D; d += 8) *reinterpret_cast<uint4*>(op + d) = make_uint4(0, 0, 0, 0);
        p.lse[((long long)b * p.Hq + hq) * p.Sq + row] = neg_inf_f();
      }
    }
    return;
  }

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&qmap); tma_prefetch_desc(&kmap); tma_prefetch_desc(&vmap); tma_prefetch_desc(&omap); }
  if (warp == 5) {   // the same warp in both CTAs allocates the same columns in both TMEMs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  f7_cluster_sync();   // barriers of both CTAs initialised and both TMEM allocations visible before anyone signals
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;       // S0 [0,128) | S1 [128,256) | O [256,384)
  const uint32_t tmem_o = tmem + 256;

  if (warp == 4) {
    // =============================== TMA producer (both CTAs) ===================================
    if (lane == 0) {
      if (leader_cta) mbar_arrive_expect_tx(q_full, 2 * kF7QBytes);        // both CTAs' query tiles
#pragma unroll
      for (int a = 0; a < 2; ++a) f7_tma_load_4d(q_s + a * kAtomBytes, &qmap, q_full, a * 64, m0, hq, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        if (leader_cta) mbar_arrive_expect_tx(&k_full[st], 2 * kF7KHalf);
#pragma unroll
        for (int a = 0; a < 2; ++a)    // this CTA's 64 keys of the tile, two 64-channel atoms
          f7_tma_load_4d(k_s + st * kF7KHalf + a * (kBlockN / 2) * 128, &kmap, &k_full[st], a * 64, j * kBlockN + (int)crank * (kBlockN / 2), hkv, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        if (leader_cta) mbar_arrive_expect_tx(&v_full[st], 2 * kF7VHalf);
        // all 128 keys, this CTA's 64 channels
        f7_tma_load_4d(v_s + st * kF7VHalf, &vmap, &v_full[st], (int)crank * 64, j * kBlockN, hkv, b);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (leader_cta) {
      const bool leader = elect_one();
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc_qk = umma_idesc(fmt, fmt, 2 * kBlockM, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc(fmt, fmt, 2 * kBlockM, D, 0, 1);
      const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(q_s), 0, 1024);
      auto issue_qk = [&](int j) {
        const int st = j % NS;
        mbar_wait(&k_full[st], (j / NS) & 1);
        tc_fence_after();
        const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(k_s + st * kF7KHalf), 0, 1024);
        const uint32_t d_tmem = tmem + (j & 1) * 128;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            f7_mma_ss(d_tmem, q_desc0 + (((kk / 4) * kAtomBytes + (kk % 4) * 32) >> 4),
                      k_desc0 + (((kk / 4) * ((kBlockN / 2) * 128) + (kk % 4) * 32) >> 4), idesc_qk, kk > 0 ? 1u : 0u);
          f7_commit_both(&k_empty[st]);
          f7_commit_both(&s_full[j & 1]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        const int st = j % NS;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        mbar_wait(&v_full[st], (j / NS) & 1);
        tc_fence_after();
        const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(v_s + st * kF7VHalf), kBlockN * 128, 1024);
        const uint32_t p_tmem = tmem + (j & 1) * 128;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk)
            f7_mma_ts(tmem_o, p_tmem + kk * 8, v_desc0 + ((kk * 2048) >> 4), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          f7_commit_both(&v_empty[st]);
          f7_commit_both(&pv_done[j & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== softmax warps (both CTAs) ==================================
    const int row = tid;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const long long q_pos = p.q_pos0 + m0 + row;
    float m_used = neg_inf_f();
    float l_sum = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int n0 = j * kBlockN;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_tmem = tmem + (j & 1) * 128 + lane_addr;
      const bool need_mask = (n0 + kBlockN > p.S) || (p.causal && (p.kv_pos0 + n0 + kBlockN - 1 > p.q_pos0 + m0));
      uint32_t pk[64];
      bool done = false;
      // ---- fast path (interior tiles): exponentiate against the CURRENT reference maximum while the tile maximum is
      // computed in the same loop (MUFU and FMNMX overlap instead of running as two serial phases).  If some row's
      // maximum turns out to have grown past the lazy-rescale threshold, nothing has been committed yet: the scores are
      // still in TMEM and the tile is redone on the exact path below.
      if (!need_mask && !__any_sync(0xffffffffu, m_used == neg_inf_f())) {
        // TMEM reads run at 64 B/clk per SM: the 64 KB score tile alone costs 1024 cycles, as much as the tile's MMAs
        // or its 16K ex2.  The row is therefore consumed in four 32-column chunks, the tcgen05.ld of chunk c + 1 in
        // flight while chunk c is exponentiated.
        const float neg_m = -m_used;
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(neg_m, neg_m);
        uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
        float mx4[4] = {neg_inf_f(), neg_inf_f(), neg_inf_f(), neg_inf_f()};
        uint32_t cb[2][32];
        tmem_ld_32x32b_x32(s_tmem, cb[0]);
        tmem_ld_wait_on(cb[0]);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc < 3) tmem_ld_32x32b_x32(s_tmem + 32 * (cc + 1), cb[(cc + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = __uint_as_float(cb[cc & 1][i]), s1 = __uint_as_float(cb[cc & 1][i + 1]);
            float x0, x1;
            unpack_f32x2(fma2_f32x2(pack_f32x2(s0, s1), sc2, nm2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            ls2[(i >> 1) & 3] = add2_f32x2(ls2[(i >> 1) & 3], pack_f32x2(p0, p1));
            pk[cc * 16 + (i >> 1)] = pack2<BF16>(p0, p1);
            mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(s0, s1));
          }
          if (cc < 3) tmem_ld_wait_on(cb[(cc + 1) & 1]);
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const bool grow = fmaxf(m_used, mx * p.scale_log2) - m_used > kRescaleThreshold;
        if (!__any_sync(0xffffffffu, grow)) {
          float a0, a1, b0, b1, c0, c1, d0, d1;
          unpack_f32x2(ls2[0], a0, a1); unpack_f32x2(ls2[1], b0, b1); unpack_f32x2(ls2[2], c0, c1); unpack_f32x2(ls2[3], d0, d1);
          l_sum += ((a0 + a1) + (b0 + b1)) + ((c0 + c1) + (d0 + d1));
          done = true;
        }
      }
      if (!done) {
        // ---- exact path: all 128 scores, mask (diagonal tiles and the ragged last tile only), row maximum, refresh,
        // exponentiate
        uint32_t sr[128];
        tmem_ld_32x32b_x32(s_tmem + 0, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
        tmem_ld_32x32b_x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
        tmem_ld_32x32b_x32(s_tmem + 64, *reinterpret_cast<uint32_t(*)[32]>(&sr[64]));
        tmem_ld_32x32b_x32(s_tmem + 96, *reinterpret_cast<uint32_t(*)[32]>(&sr[96]));
        tmem_ld_wait();
        if (need_mask) {
          long long lim = (long long)p.S - n0 - 1;
          if (p.causal) lim = min(lim, q_pos - p.kv_pos0 - n0);
          const int limc = (int)max(-1LL, min(lim, 127LL));
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c > limc) sr[c] = 0xff800000u;
        }
        float mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(sr[i]);
#pragma unroll
        for (int c = 8; c < 128; c += 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) mx8[i] = fmaxf(mx8[i], __uint_as_float(sr[c + i]));
        }
        const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        const float m_new = fmaxf(m_used, mx * p.scale_log2);
        // lazy rescale: refresh the reference max only when it moved by more than the threshold
        const bool refresh = (m_new - m_used > kRescaleThreshold) || (m_used == neg_inf_f() && m_new != neg_inf_f());
        if (__any_sync(0xffffffffu, refresh) ) {
          const float alpha = refresh ? fast_exp2(m_used - m_new) : 1.f;  // m_used=-inf -> 0
          if (refresh) { l_sum *= alpha; m_used = m_new; }
          if (j > 0) {
            mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
              uint32_t orow[32];
              tmem_ld_32x32b_x32(tmem_o + lane_addr + c0, orow);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * alpha);
              tmem_st_32x32b_x32(tmem_o + lane_addr + c0, orow);
            }
          }
        }
        const float m_sub = (m_used == neg_inf_f()) ? 0.f : m_used;
        const float neg_m = -m_sub;
        // exp2(s * c - m) with packed f32x2 scale-subtract and row-sum (FFMA2 / FADD2: half the issue slots)
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(neg_m, neg_m);
        uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float x0, x1;
            unpack_f32x2(fma2_f32x2(pack_f32x2(__uint_as_float(sr[c + i]), __uint_as_float(sr[c + i + 1])), sc2, nm2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            ls2[i >> 1] = add2_f32x2(ls2[i >> 1], pack_f32x2(p0, p1));
            pk[(c + i) >> 1] = pack2<BF16>(p0, p1);
          }
        }
        {
          float a0, a1, b0, b1, c0, c1, d0, d1;
          unpack_f32x2(ls2[0], a0, a1); unpack_f32x2(ls2[1], b0, b1); unpack_f32x2(ls2[2], c0, c1); unpack_f32x2(ls2[3], d0, d1);
          l_sum += ((a0 + a1) + (b0 + b1)) + ((c0 + c1) + (d0 + d1));
        }
      }
      tmem_st_32x32b_x32(s_tmem + 0, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st_32x32b_x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) f7_arrive_leader(&p_full[j & 1]);
    }
    // ------------------------------- epilogue --------------------------------------------------
    const int jl = n_tiles - 1;
    mbar_wait(&pv_done[jl & 1], (jl >> 1) & 1);
    tc_fence_after();
    const float inv_l = l_sum > 0.f ? 1.f / l_sum : 0.f;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t orow[32];
      tmem_ld_32x32b_x32(tmem_o + lane_addr + c0, orow);
      tmem_ld_wait();
      uint8_t* base = q_s + (c0 >> 6) * kAtomBytes + row * 128;   // all QK^T MMAs are complete: the Q tile buffer is free
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 w;
        w.x = pack2<BF16>(__uint_as_float(orow[g * 8 + 0]) * inv_l, __uint_as_float(orow[g * 8 + 1]) * inv_l);
        w.y = pack2<BF16>(__uint_as_float(orow[g * 8 + 2]) * inv_l, __uint_as_float(orow[g * 8 + 3]) * inv_l);
        w.z = pack2<BF16>(__uint_as_float(orow[g * 8 + 4]) * inv_l, __uint_as_float(orow[g * 8 + 5]) * inv_l);
        w.w = pack2<BF16>(__uint_as_float(orow[g * 8 + 6]) * inv_l, __uint_as_float(orow[g * 8 + 7]) * inv_l);
        const int chunk = ((c0 & 63) >> 3) + g;
        *reinterpret_cast<uint4*>(base + ((chunk ^ (row & 7)) << 4)) = w;
      }
    }
    const float lse_row = l_sum > 0.f ? (m_used + fast_log2(l_sum)) * 0.6931471805599453f : neg_inf_f();
    if (m0 + row < p.Sq) p.lse[((long long)b * p.Hq + hq) * p.Sq + m0 + row] = lse_row;
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(1, 128);
    if (tid == 0) {
#pragma unroll
      for (int a = 0; a < 2; ++a) tma_store_4d(&omap, q_s + a * kAtomBytes, a * 64, m0, hq, b);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  f7_cluster_sync();   // nobody frees TMEM or exits while the peer's MMAs / remote arrives may still target this CTA
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <bool BF16>
void launch_fwd7(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t stream) {
  constexpr int D = kF7D;
  CUtensorMap qmap = make_tmap_bhsd(q, 2, s.B, s.Hq, s.Sq, D, s.q_sb, s.q_sh, s.q_ss, 64, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap kmap = make_tmap_bhsd(k, 2, s.B, s.Hkv, s.S, D, s.k_sb, s.k_sh, s.k_ss, 64, kBlockN / 2, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap vmap = make_tmap_bhsd(v, 2, s.B, s.Hkv, s.S, D, s.v_sb, s.v_sh, s.v_ss, 64, kBlockN, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap omap = make_tmap_bhsd(out, 2, s.B, s.Hq, s.Sq, D, s.o_sb, s.o_sh, s.o_ss, 64, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B);
  FwdParams p;
  p.lse = lse; p.out = out; p.o_sb = s.o_sb; p.o_sh = s.o_sh; p.o_ss = s.o_ss;
  p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.G = s.Hq / s.Hkv; p.Sq = s.Sq; p.S = s.S;
  p.scale_log2 = s.softmax_scale * 1.4426950408889634f;
  p.causal = s.causal; p.q_pos0 = s.q_pos0; p.kv_pos0 = s.kv_pos0;
  p.num_m_tiles = (s.Sq + kBlockM - 1) / kBlockM;
  p.n_items = p.num_m_tiles * s.Hq * s.B;
  p.lag = 0;
  p.q_in_tmem = 0;
  p.comm = to_device_ctx(CommCtxHost{});
  const int num_pairs = (p.num_m_tiles + 1) / 2;
  auto kern = attn_fwd7_kernel<BF16>;
  static bool configured = false;
  if (!configured) {
    TA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF7Smem));
    configured = true;
  }
  const int grid = 2 * num_pairs * s.Hq * s.B;   // clusters of 2 (static __cluster_dims__)
  kern<<<grid, kF7Threads, kF7Smem, stream>>>(qmap, kmap, vmap, omap, p, num_pairs);
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void attn_fwd7_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                      const CommCtxHost& comm, cudaStream_t stream) {
  if (s.D != 128) throw std::runtime_error("attn_fwd7 (experimental 2-CTA forward): head_dim must be 128");
  if (comm.world > 1) throw std::runtime_error("attn_fwd7 (experimental 2-CTA forward): no fused multi-GPU mode");
  if (s.Hq % s.Hkv != 0) throw std::runtime_error("attn_fwd7: Hq must be a multiple of Hkv");
  if (s.S <= 0 || s.Sq <= 0) throw std::runtime_error("attn_fwd7: empty problem");
  if (s.is_bf16) launch_fwd7<true>(s, q, k, v, out, lse, stream);
  else launch_fwd7<false>(s, q, k, v, out, lse, stream);
}

}  // namespace ta
