// attn_partial_fwd -- flash-attention forward for sm_100a on the 5th-generation tensor cores.
//
// Replaces the reference's local attention (/root/reference/model.py:74-80: matmul -> softmax -> matmul
// with the full score row materialised in HBM) for Sq >= 1 with GQA, causal masks
// over GLOBAL positions, and the (o, lse) contract the tree combine needs.
//
// One CTA = one 128-row query tile of one (batch, q-head); 6 warps, warp-specialised:
//   warps 0-3  softmax: tcgen05.ld S (one row per thread) -> online softmax with lazy rescale ->
//              P (bf16/fp16) written back to TMEM over S -> epilogue O/l, lse, swizzled smem, TMA store
//   warp  4    TMA producer: Q once, then K and V tiles through two 3-stage mbarrier rings
//   warp  5    MMA issuer (one elected lane): S = Q K^T (SS form, K-major operands),
//              O += P V (TS form: A = P from TMEM, B = V MN-major from smem); owns the TMEM allocation
// TMEM (512 columns): S0 [0,128) | S1 [128,256) | O [256,256+D).  S is double-buffered so that
// QK^T of tile j+1 and PV of tile j-1 run on the tensor pipe while the softmax of tile j runs on the
// SIMT pipes; P_j aliases the first 64 columns of S_j.
// All layout conventions used here are verified on hardware by csrc/umma_probe.cu (tests/test_gpu_probe.py).
#include "attn_fwd_common.cuh"

namespace ta {
namespace {
using namespace fwd_detail;

// profiling aid: cycles spent by softmax thread 0 of CTA 0 in {waiting for S, fast path, exact path, P store + signal}, tiles
__device__ unsigned long long g_fwd_phase_cycles[5];

template <int D, bool BF16, bool kComm>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap omap,
                const FwdParams p) {
  using SM = FwdSmem<D>;
  constexpr int NS = SM::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + SM::kQBytes;
  uint8_t* v_s = k_s + NS * SM::kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + NS * SM::kKVBytes);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // NS
  uint64_t* k_empty = k_full + NS;    // NS
  uint64_t* v_full = k_empty + NS;    // NS
  uint64_t* v_empty = v_full + NS;    // NS
  uint64_t* s_full = v_empty + NS;    // 2
  uint64_t* p_full = s_full + 2;      // 2
  uint64_t* pv_done = p_full + 2;     // 2
  uint64_t* qt_ready = pv_done + 2;   // 1 (count 4): the query tile has been copied to TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qt_ready + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  // ---- work decode.  Single GPU: block x computes item x.  Fused multi-GPU: compute CTAs and merge CTAs
  // share the launch, merges trail their item by `lag` compute CTAs:
  //   x < L: compute x | L <= x < 2N-L: even -> compute, odd -> merge | x >= 2N-L: merge
  int item = blockIdx.x;
  uint32_t epoch = 0;
  if constexpr (kComm) {
    epoch = ld_relaxed_sys_u32(p.comm.epoch) + 1;
    const int N = p.n_items, L = p.lag, x = blockIdx.x;
    bool is_merge = false;
    if (x < L) item = x;
    else if (x < 2 * N - L) { const int k = x - L; if (k & 1) { is_merge = true; item = k >> 1; } else item = L + (k >> 1); }
    else { is_merge = true; item = (N - L) + (x - (2 * N - L)); }
    if (is_merge) {
      merge_item<D, BF16>(p, item, epoch, smem);
      comm_kernel_exit(p, epoch);
      return;
    }
  }
  const int m_tile = p.num_m_tiles - 1 - (item % p.num_m_tiles);  // heaviest (causal) tiles first
  const int bh = item / p.num_m_tiles;
  const int hq = bh % p.Hq, b = bh / p.Hq;
  const int hkv = hq / p.G;
  const int m0 = m_tile * kBlockM;
  // fused combine: this tile's partial goes to the OWNER of the tile only (reduce-scatter over the query tiles)
  using SL = FwdSlots<D>;
  int owner = 0;
  size_t slot_off = 0, flag_off = 0;
  if constexpr (kComm) {
    owner = min(m_tile / p.tiles_per_rank, p.comm.world - 1);
    const int litem = bh * p.tiles_per_rank + (m_tile - owner * p.tiles_per_rank);
    slot_off = SL::part_off(p, epoch, p.comm.rank, litem);
    flag_off = SL::part_flag(p, epoch, p.comm.rank, litem);
  }

  // number of KV tiles this query tile can see
  // (two-segment shards -- zigzag sharding of a causal sequence: rows [0, seg_len) sit at kv_pos0 + row, rows [seg_len, S) at
  // kv_pos0 + seg_gap + row with seg_gap >= 0, so whatever is visible is still a PREFIX of the local rows)
  int n_end = p.S;
  if (p.causal) {
    const long long last_q = p.q_pos0 + min(m0 + kBlockM - 1, p.Sq - 1);
    const int len_a = min(p.seg_len, p.S);
    const long long vis_a = max(0LL, min((long long)len_a, last_q - p.kv_pos0 + 1));
    const long long vis_b = max(0LL, min((long long)(p.S - len_a), last_q - (p.kv_pos0 + p.seg_gap + len_a) + 1));
    n_end = (int)(vis_b > 0 ? len_a + vis_b : vis_a);
  }
  const int n_tiles = (n_end + kBlockN - 1) / kBlockN;

  if (n_tiles == 0) {
    // nothing visible: the monoid identity (0, -inf)
    if constexpr (kComm) {
      if (!p.comm.skip_publish) {
        uint8_t* slot = reinterpret_cast<uint8_t*>(p.comm.data[owner]) + slot_off;
        for (int c = tid; c < kBlockM * (D / 8); c += kFwdThreads) reinterpret_cast<uint4*>(slot)[c] = make_uint4(0, 0, 0, 0);
        if (tid < kBlockM) reinterpret_cast<float*>(slot + kBlockM * D * 2)[tid] = neg_inf_f();
        __syncthreads();
        if (tid == 0) { fence_acq_rel_sys(); st_release_sys_u32(p.comm.flags[owner] + flag_off, epoch); }
      }
      comm_kernel_exit(p, epoch);
    } else if (warp < 4) {
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
    mbar_init(qt_ready, 4);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&qmap); tma_prefetch_desc(&kmap); tma_prefetch_desc(&vmap); tma_prefetch_desc(&omap);
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + 256;
  const uint32_t tmem_q = tmem + 384;   // q_in_tmem: Q as the TMEM A operand, D/2 columns

  if (warp == 4) {
    // =============================== TMA producer ===============================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, SM::kQBytes);
#pragma unroll
      for (int a = 0; a < SM::kAtoms; ++a) tma_load_4d(q_s + a * SM::kAtomBytes, &qmap, q_full, a * 64, m0, hq, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], SM::kKVBytes);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(k_s + st * SM::kKVBytes + a * SM::kAtomBytes, &kmap, &k_full[st], a * 64, j * kBlockN, hkv, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], SM::kKVBytes);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(v_s + st * SM::kKVBytes + a * SM::kAtomBytes, &vmap, &v_full[st], a * 64, j * kBlockN, hkv, b);
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer =================================================
    // The whole warp runs this loop convergently so that every descriptor / address stays in uniform registers
    // (issuing from inside `if (lane == 0)` makes ptxas wrap each tcgen05.mma in an R2UR waterfall loop: ~17
    // instructions per MMA, enough to make the single issuing thread the bottleneck); one elected lane issues.
    {
      const bool leader = elect_one();
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc_qk = umma_idesc(fmt, fmt, kBlockM, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc(fmt, fmt, kBlockM, D, 0, 1);
      const uint32_t q_addr = smem_u32(q_s);
      const uint64_t q_desc0 = umma_smem_desc_sw128(q_addr, 0, 1024);
      auto issue_qk = [&](int j) {
        const int st = j % NS;
        mbar_wait(&k_full[st], (j / NS) & 1);
        tc_fence_after();
        const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(k_s + st * SM::kKVBytes), 0, 1024);
        const uint32_t d_tmem = tmem + (j & 1) * 128;
        if (leader) {
#pragma unroll
          if (p.q_in_tmem) {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
              umma_ts_f16(d_tmem, tmem_q + kk * 8, k_desc0 + (((kk / 4) * SM::kAtomBytes + (kk % 4) * 32) >> 4), idesc_qk, kk > 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
              const uint32_t off = ((kk / 4) * SM::kAtomBytes + (kk % 4) * 32) >> 4;   // descriptor address field: bytes / 16
              umma_ss_f16(d_tmem, q_desc0 + off, k_desc0 + off, idesc_qk, kk > 0 ? 1u : 0u);
            }
          }
          if (j + NS < n_tiles) umma_commit(&k_empty[st]);   // only when the producer will refill this stage: no arrival is left un-waited
          umma_commit(&s_full[j & 1]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      if (p.q_in_tmem) { mbar_wait(qt_ready, 0); tc_fence_after(); }
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        const int st = j % NS;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        mbar_wait(&v_full[st], (j / NS) & 1);
        tc_fence_after();
        const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(v_s + st * SM::kKVBytes), kBlockN * 128, 1024);
        const uint32_t p_tmem = tmem + (j & 1) * 128;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk)
            umma_ts_f16(tmem_o, p_tmem + kk * 8, v_desc0 + ((kk * 2048) >> 4), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          if (j + NS < n_tiles) umma_commit(&v_empty[st]);
          umma_commit(&pv_done[j & 1]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // =============================== softmax warps ==============================================
    const int row = tid;                                   // row of the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;  // this warp's lane quarter
    const long long q_pos = p.q_pos0 + m0 + row;
    float m_used = neg_inf_f();  // running (stale) max, scaled log2 domain
    float l_sum = 0.f;
    if (p.q_in_tmem) {
      // one-time: this thread's query row, un-swizzled from the TMA tile, becomes lane `row` of the TMEM A operand
      mbar_wait(q_full, 0);
#pragma unroll
      for (int a = 0; a < SM::kAtoms; ++a) {
        uint32_t w[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(q_s + a * SM::kAtomBytes + row * 128 + ((c ^ (row & 7)) << 4));
          w[c * 4 + 0] = u.x; w[c * 4 + 1] = u.y; w[c * 4 + 2] = u.z; w[c * 4 + 3] = u.w;
        }
        tmem_st_32x32b_x32(tmem_q + lane_addr + a * 32, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qt_ready);
    }
    const bool prof = (blockIdx.x == 0 && tid == 0);
    long long c_wait = 0, c_fast = 0, c_exact = 0, c_store = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int n0 = j * kBlockN;
      const long long tp0 = clock64();
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const long long tp1 = clock64();
      const uint32_t s_tmem = tmem + (j & 1) * 128 + lane_addr;
      const long long tile_pos = p.kv_pos0 + n0 + (n0 >= p.seg_len ? p.seg_gap : 0LL);   // global position of the tile's key 0
      const bool need_mask = (n0 + kBlockN > p.S) || (p.causal && (tile_pos + kBlockN - 1 > p.q_pos0 + m0));
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
      const long long tp2 = clock64();
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
          if (p.causal) lim = min(lim, q_pos - tile_pos);
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
      const long long tp3 = clock64();
      tmem_st_32x32b_x32(s_tmem + 0, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st_32x32b_x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
      c_wait += tp1 - tp0; c_fast += tp2 - tp1; c_exact += tp3 - tp2; c_store += clock64() - tp3;
    }
    if (prof) {
      g_fwd_phase_cycles[0] = c_wait; g_fwd_phase_cycles[1] = c_fast; g_fwd_phase_cycles[2] = c_exact; g_fwd_phase_cycles[3] = c_store;
      g_fwd_phase_cycles[4] = n_tiles;
    }
    // ------------------------------- epilogue --------------------------------------------------
    const int jl = n_tiles - 1;
    // both PV barriers are waited to their final phase: the MMAs retire in order, so the second wait is free, and no
    // barrier phase is left completed-but-never-waited at CTA exit (compute-sanitizer synccheck "Missing wait", round 1)
    if (jl >= 1) mbar_wait(&pv_done[(jl - 1) & 1], ((jl - 1) >> 1) & 1);
    mbar_wait(&pv_done[jl & 1], (jl >> 1) & 1);
    tc_fence_after();
    const float inv_l = l_sum > 0.f ? 1.f / l_sum : 0.f;
    // all QK^T MMAs are complete -> the Q tile buffer is free: stage O there in the TMA/UMMA 128B swizzle
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t orow[32];
      tmem_ld_32x32b_x32(tmem_o + lane_addr + c0, orow);
      tmem_ld_wait();
      const int atom = c0 >> 6;
      uint8_t* base = q_s + atom * SM::kAtomBytes + row * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {  // 4 chunks of 8 elements (16 B)
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
    if constexpr (!kComm) {
      if (m0 + row < p.Sq) p.lse[((long long)b * p.Hq + hq) * p.Sq + m0 + row] = lse_row;
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, kSoftmaxThreads);
      if (tid == 0) {
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a) tma_store_4d(&omap, q_s + a * SM::kAtomBytes, a * 64, m0, hq, b);
        tma_store_commit();
        tma_store_wait<0>();
      }
    } else {
      // fused tree combine, step 1 (reduce-scatter): push this rank's partial tile (o in the I/O dtype, lse fp32) into
      // slot [my rank][local item] of the tile's OWNER with coalesced 16-byte P2P stores, then release its epoch flag.
      tc_fence_before();
      named_bar_sync(1, kSoftmaxThreads);
      if (!p.comm.skip_publish) {
        constexpr int CPR = D / 8;
        uint8_t* slot = reinterpret_cast<uint8_t*>(p.comm.data[owner]) + slot_off;
#pragma unroll 4
        for (int c = tid; c < kBlockM * CPR; c += kSoftmaxThreads) {
          const int r = c / CPR, ch = c - r * CPR;
          const uint4 w = *reinterpret_cast<const uint4*>(q_s + (ch >> 3) * SM::kAtomBytes + r * 128 + (((ch & 7) ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(slot + (size_t)r * D * 2 + ch * 16) = w;
        }
        reinterpret_cast<float*>(slot + kBlockM * D * 2)[row] = lse_row;
        named_bar_sync(1, kSoftmaxThreads);
        if (tid == 0) { fence_acq_rel_sys(); st_release_sys_u32(p.comm.flags[owner] + flag_off, epoch); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
  if constexpr (kComm) comm_kernel_exit(p, epoch);
}

template <int D, bool BF16, bool kComm>
void launch_fwd(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                const CommCtxHost& comm, cudaStream_t stream, int q_in_tmem, int comm_mode, int sq_out) {
  using SM = FwdSmem<D>;
  CUtensorMap qmap = make_tmap_bhsd(q, 2, s.B, s.Hq, s.Sq, D, s.q_sb, s.q_sh, s.q_ss, 64, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap kmap = make_tmap_bhsd(k, 2, s.B, s.Hkv, s.S, D, s.k_sb, s.k_sh, s.k_ss, 64, kBlockN, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap vmap = make_tmap_bhsd(v, 2, s.B, s.Hkv, s.S, D, s.v_sb, s.v_sh, s.v_ss, 64, kBlockN, CU_TENSOR_MAP_SWIZZLE_128B);
  // (fused mode writes `out` with plain stores from the merge CTAs; the TMA store map is only used single-GPU)
  const int out_rows = (kComm && comm_mode == 2) ? sq_out : s.Sq;
  CUtensorMap omap = make_tmap_bhsd(out, 2, s.B, s.Hq, out_rows, D, s.o_sb, s.o_sh, s.o_ss, 64, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B);
  FwdParams p;
  p.lse = lse; p.out = out; p.o_sb = s.o_sb; p.o_sh = s.o_sh; p.o_ss = s.o_ss;
  p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.G = s.Hq / s.Hkv; p.Sq = s.Sq; p.S = s.S;
  p.scale_log2 = s.softmax_scale * 1.4426950408889634f;
  p.causal = s.causal; p.q_pos0 = s.q_pos0; p.kv_pos0 = s.kv_pos0;
  p.seg_len = s.kv_seg_len > 0 ? s.kv_seg_len : 0x7fffffff;
  p.seg_gap = s.kv_seg_len > 0 ? s.kv_seg_gap : 0;
  p.num_m_tiles = (s.Sq + kBlockM - 1) / kBlockM;
  p.n_items = p.num_m_tiles * s.Hq * s.B;
  p.lag = std::min(p.n_items, 2 * num_sms());
  p.q_in_tmem = q_in_tmem;
  p.comm = to_device_ctx(comm);
  p.mode = 0; p.tiles_per_rank = p.num_m_tiles; p.n_local = p.n_items; p.final_off = 0; p.final_flag_off = 0;
  p.sq_out = s.Sq;
  if (kComm) {
    const int W = comm.world;
    const size_t slot = (size_t)kBlockM * D * 2 + kBlockM * 4;
    p.mode = comm_mode == 2 ? 2 : 1;
    p.tiles_per_rank = (p.num_m_tiles + W - 1) / W;
    p.n_local = s.B * s.Hq * p.tiles_per_rank;
    const size_t part_bytes = (size_t)(p.mode == 2 ? 2 : 1) * W * p.n_local * slot;
    const size_t part_flags = (size_t)(p.mode == 2 ? 2 : 1) * W * p.n_local;
    p.final_off = (long long)part_bytes;
    p.final_flag_off = (int)part_flags;
    const size_t need_data = part_bytes + (p.mode == 1 ? (size_t)p.n_items * slot : 0);
    const size_t need_flags = (part_flags + (p.mode == 1 ? (size_t)p.n_items : 0)) * 4;
    if (need_data > comm.data_bytes || need_flags > comm.flag_bytes)
      throw std::runtime_error("attn_fwd(fused): symmetric buffer too small");
    if (p.mode == 2) p.sq_out = sq_out;
  }
  auto kern = attn_fwd_kernel<D, BF16, kComm>;
  static bool configured = false;
  if (!configured) {
    TA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::kTotal));
    configured = true;
  }
  dim3 grid(kComm ? 2 * p.n_items : p.n_items);
  kern<<<grid, kFwdThreads, SM::kTotal, stream>>>(qmap, kmap, vmap, omap, p);
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void attn_fwd_phase_cycles(unsigned long long* out5) {
  TA_CUDA_CHECK(cudaMemcpyFromSymbol(out5, g_fwd_phase_cycles, 5 * sizeof(unsigned long long)));
}

size_t attn_fwd_comm_bytes(const AttnShape& s, int world, size_t* flag_bytes, int comm_mode) {
  const size_t num_m = (size_t)((s.Sq + kBlockM - 1) / kBlockM);
  const size_t n_items = num_m * s.Hq * s.B;
  const size_t tpr = (num_m + world - 1) / world;
  const size_t n_local = (size_t)s.B * s.Hq * tpr;
  const size_t slot = (size_t)kBlockM * s.D * 2 + kBlockM * 4;
  const size_t par = comm_mode == 2 ? 2 : 1;
  *flag_bytes = (par * world * n_local + (comm_mode == 2 ? 0 : n_items)) * 4;
  return par * world * n_local * slot + (comm_mode == 2 ? 0 : n_items * slot);
}

void attn_fwd_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                     const CommCtxHost& comm, cudaStream_t stream, int q_in_tmem, int comm_mode, int sq_out) {
  if (s.D != 64 && s.D != 128) throw std::runtime_error("attn_fwd: head_dim must be 64 or 128");
  if (s.kv_seg_len < 0 || s.kv_seg_len % kBlockN != 0 || s.kv_seg_gap < 0)
    throw std::runtime_error("attn_fwd: a two-segment shard needs kv_seg_len % 128 == 0 and kv_seg_gap >= 0");
  if (s.Hq % s.Hkv != 0) throw std::runtime_error("attn_fwd: Hq must be a multiple of Hkv");
  if (s.S <= 0 || s.Sq <= 0) throw std::runtime_error("attn_fwd: empty problem");
  const bool fused = comm.world > 1;
#define TA_FWD(DD, BB)                                                               \
  if (fused) launch_fwd<DD, BB, true>(s, q, k, v, out, lse, comm, stream, q_in_tmem, comm_mode, sq_out);           \
  else launch_fwd<DD, BB, false>(s, q, k, v, out, lse, comm, stream, q_in_tmem, comm_mode, sq_out);
  if (s.D == 128) {
    if (s.is_bf16) { TA_FWD(128, true) } else { TA_FWD(128, false) }
  } else {
    if (s.is_bf16) { TA_FWD(64, true) } else { TA_FWD(64, false) }
  }
#undef TA_FWD
}

}  // namespace ta
