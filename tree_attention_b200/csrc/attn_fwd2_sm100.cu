// attn_fwd2 -- tcgen05 flash forward with TWO ping-ponged 128-row query tiles per CTA (M = 256).
//
// Same contract as attn_fwd_sm100.cu (local partial (o, lse) or, in fused mode, the whole cross-GPU tree
// combine in the same launch).  Differences that buy throughput:
//   * two softmax warpgroups (A: warps 0-3, B: warps 4-7) each own one query tile; the single MMA thread
//     interleaves  PV_A(j) | S_A(j+1) | PV_B(j) | S_B(j+1)  so that while one warpgroup runs its softmax the
//     tensor pipe is busy with the other tile's GEMMs (two resident softmax warps per SM sub-partition hide
//     each other's MUFU / TMEM latencies);
//   * every K/V tile fetched from L2 feeds 256 query rows (half the L2->SMEM traffic per FLOP);
//   * the softmax is two-pass over TMEM in 32-column chunks (max, then exp/pack), so a row needs ~40
//     registers instead of 128 and the CTA runs 12 warps without spills.
// TMEM: S_A [0,128) | S_B [128,256) | O_A [256,384) | O_B [384,512);  P_t aliases S_t[0,64).
// Reference: replaces /root/reference/model.py:74-80 (matmul -> softmax -> matmul -> logsumexp), like attn_fwd_sm100.cu.
#include "attn_fwd_common.cuh"

namespace ta {
namespace {
using namespace fwd_detail;

constexpr int kFwd2Threads = 384;
constexpr int kWG = 128;

template <int D>
struct Fwd2Smem {
  static constexpr int kStages = 2;
  static constexpr int kAtoms = D / 64;
  static constexpr int kTileBytes = 128 * D * 2;
  static constexpr int kAtomBytes = 128 * 128;
  static constexpr size_t kTotal = 1024 + size_t(2 + 2 * kStages) * kTileBytes + 256;
};

template <int D, bool BF16, bool kComm>
__global__ void __launch_bounds__(kFwd2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                 const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap omap,
                 const FwdParams p, const int num_pairs, const int n_compute) {
  using SM = Fwd2Smem<D>;
  constexpr int NS = SM::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;                          // [2 tiles]
  uint8_t* k_s = q_s + 2 * SM::kTileBytes;      // [NS]
  uint8_t* v_s = k_s + NS * SM::kTileBytes;     // [NS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + NS * SM::kTileBytes);
  uint64_t* q_full = bars;             // 1
  uint64_t* k_full = bars + 1;         // NS
  uint64_t* k_empty = k_full + NS;
  uint64_t* v_full = k_empty + NS;
  uint64_t* v_empty = v_full + NS;
  uint64_t* s_full = v_empty + NS;     // 2 (per tile)
  uint64_t* p_full = s_full + 2;       // 2
  uint64_t* pv_done = p_full + 2;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t epoch = 0;
  if constexpr (kComm) {
    epoch = ld_relaxed_sys_u32(p.comm.epoch) + 1;
    if ((int)blockIdx.x >= n_compute) {  // merge CTAs trail all compute CTAs of the launch
      merge_item<D, BF16>(p, (int)blockIdx.x - n_compute, epoch, smem);
      comm_kernel_exit(p, epoch);
      return;
    }
  }
  const int pair = num_pairs - 1 - (int)(blockIdx.x % num_pairs);  // heaviest (causal) pairs first
  const int bh = blockIdx.x / num_pairs;
  const int hq = bh % p.Hq, b = bh / p.Hq;
  const int hkv = hq / p.G;
  const int m0 = pair * 2 * kBlockM;
  const bool has_b = (pair * 2 + 1) < p.num_m_tiles;
  constexpr int kSlotBytes = kBlockM * D * 2 + kBlockM * 4;

  int n_end = p.S;
  if (p.causal) {
    const long long last_q = p.q_pos0 + min(m0 + 2 * kBlockM - 1, p.Sq - 1);
    n_end = (int)max(0LL, min((long long)p.S, last_q - p.kv_pos0 + 1));
  }
  const int n_tiles = (n_end + kBlockN - 1) / kBlockN;

  // slot / item bookkeeping for the fused mode
  auto item_of = [&](int t) { return bh * p.num_m_tiles + (p.num_m_tiles - 1 - (pair * 2 + t)); };
  auto slot_off = [&](int t) { return ((size_t)((epoch & 1) * p.comm.world + p.comm.rank) * p.n_items + item_of(t)) * kSlotBytes; };
  auto flag_off = [&](int t) { return (size_t)((epoch & 1) * p.comm.world + p.comm.rank) * p.n_items + item_of(t); };

  if (n_tiles == 0) {
    // nothing visible for either tile: the monoid identity (0, -inf)
    for (int t = 0; t < (has_b ? 2 : 1); ++t) {
      if constexpr (kComm) {
        if (!p.comm.skip_publish) {
          for (int dst = 0; dst < p.comm.world; ++dst) {
            uint8_t* slot = reinterpret_cast<uint8_t*>(p.comm.data[dst]) + slot_off(t);
            for (int c = tid; c < kBlockM * (D / 8); c += kFwd2Threads) reinterpret_cast<uint4*>(slot)[c] = make_uint4(0, 0, 0, 0);
            if (tid < kBlockM) reinterpret_cast<float*>(slot + kBlockM * D * 2)[tid] = neg_inf_f();
          }
          __syncthreads();
          if (tid < p.comm.world) { fence_acq_rel_sys(); st_release_sys_u32(p.comm.flags[tid] + flag_off(t), epoch); }
        }
      } else if (tid < kBlockM) {
        const int row = m0 + t * kBlockM + tid;
        if (row < p.Sq) {
          uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)hq * p.o_sh + (long long)row * p.o_ss;
          for (int d = 0; d < D; d += 8) *reinterpret_cast<uint4*>(op + d) = make_uint4(0, 0, 0, 0);
          p.lse[((long long)b * p.Hq + hq) * p.Sq + row] = neg_inf_f();
        }
      }
    }
    if constexpr (kComm) comm_kernel_exit(p, epoch);
    return;
  }

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) { tma_prefetch_desc(&qmap); tma_prefetch_desc(&kmap); tma_prefetch_desc(&vmap); tma_prefetch_desc(&omap); }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // =============================== TMA producer ===============================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * SM::kTileBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(q_s + t * SM::kTileBytes + a * SM::kAtomBytes, &qmap, q_full, a * 64, m0 + t * kBlockM, hq, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], SM::kTileBytes);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(k_s + st * SM::kTileBytes + a * SM::kAtomBytes, &kmap, &k_full[st], a * 64, j * kBlockN, hkv, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], SM::kTileBytes);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(v_s + st * SM::kTileBytes + a * SM::kAtomBytes, &vmap, &v_full[st], a * 64, j * kBlockN, hkv, b);
      }
    }
  } else if (warp == 9) {
    // =============================== MMA issuer =================================================
    if (lane == 0) {
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc_qk = umma_idesc(fmt, fmt, kBlockM, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc(fmt, fmt, kBlockM, D, 0, 1);
      auto issue_qk = [&](int t, int j) {
        const uint32_t q_addr = smem_u32(q_s + t * SM::kTileBytes);
        const uint32_t k_addr = smem_u32(k_s + (j % NS) * SM::kTileBytes);
        const uint32_t d_tmem = tmem + t * 128;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk / 4) * SM::kAtomBytes + (kk % 4) * 32;
          umma_ss_f16(d_tmem, umma_smem_desc_sw128(q_addr + off, 0, 1024), umma_smem_desc_sw128(k_addr + off, 0, 1024),
                      idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {
        const uint32_t v_addr = smem_u32(v_s + (j % NS) * SM::kTileBytes);
        const uint32_t p_tmem = tmem + t * 128;
        const uint32_t o_tmem = tmem + 256 + t * 128;
#pragma unroll
        for (int kk = 0; kk < kBlockN / 16; ++kk)
          umma_ts_f16(o_tmem, p_tmem + kk * 8, umma_smem_desc_sw128(v_addr + kk * 2048, kBlockN * 128, 1024), idesc_pv,
                      (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&pv_done[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      if (has_b) issue_qk(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(&p_full[0], j & 1);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        issue_pv(0, j);
        if (j + 1 < n_tiles) {
          mbar_wait(&k_full[(j + 1) % NS], ((j + 1) / NS) & 1);
          tc_fence_after();
          issue_qk(0, j + 1);
        }
        if (has_b) {
          mbar_wait(&p_full[1], j & 1);
          tc_fence_after();
          issue_pv(1, j);
        }
        umma_commit(&v_empty[st]);
        if (j + 1 < n_tiles) {
          if (has_b) issue_qk(1, j + 1);
          umma_commit(&k_empty[(j + 1) % NS]);
        }
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // =============================== softmax warpgroups =========================================
    const int t = warp >> 2;  // 0 = tile A, 1 = tile B
    if (t == 0 || has_b) {
      const int row = tid & (kWG - 1);
      const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
      const int m0t = m0 + t * kBlockM;
      const long long q_pos = p.q_pos0 + m0t + row;
      const uint32_t s_tmem = tmem + t * 128 + lane_addr;
      const uint32_t o_tmem = tmem + 256 + t * 128 + lane_addr;
      uint8_t* stage = q_s + t * SM::kTileBytes;
      float m_used = neg_inf_f();
      float l_sum = 0.f;
      for (int j = 0; j < n_tiles; ++j) {
        const int n0 = j * kBlockN;
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        int limc = 127;
        if ((n0 + kBlockN > p.S) || (p.causal && (p.kv_pos0 + n0 + kBlockN - 1 > p.q_pos0 + m0t))) {
          long long lim = (long long)p.S - n0 - 1;
          if (p.causal) lim = min(lim, q_pos - p.kv_pos0 - n0);
          limc = (int)max(-1LL, min(lim, 127LL));
        }
        // ---- pass 1: row max
        float mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = neg_inf_f();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t sr[32];
          tmem_ld_32x32b_x32(s_tmem + c0, sr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = (c0 + i <= limc) ? __uint_as_float(sr[i]) : neg_inf_f();
            mx8[i & 7] = fmaxf(mx8[i & 7], v);
          }
        }
        const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        const float m_new = fmaxf(m_used, mx * p.scale_log2);
        const bool refresh = (m_new - m_used > kRescaleThreshold) || (m_used == neg_inf_f() && m_new != neg_inf_f());
        if (__any_sync(0xffffffffu, refresh)) {
          const float alpha = refresh ? fast_exp2(m_used - m_new) : 1.f;
          if (refresh) { l_sum *= alpha; m_used = m_new; }
          if (j > 0) {
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
              uint32_t orow[32];
              tmem_ld_32x32b_x32(o_tmem + c0, orow);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * alpha);
              tmem_st_32x32b_x32(o_tmem + c0, orow);
            }
          }
        }
        const float neg_m = (m_used == neg_inf_f()) ? 0.f : -m_used;
        // ---- pass 2: P = exp2(S * c - m), row sum, pack, store over S
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t sr[32];
          tmem_ld_32x32b_x32(s_tmem + c0, sr);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2, neg_m));
            float p1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2, neg_m));
            p0 = (c0 + i <= limc) ? p0 : 0.f;
            p1 = (c0 + i + 1 <= limc) ? p1 : 0.f;
            ls[(i >> 1) & 3] += p0 + p1;
            pk[i >> 1] = pack2<BF16>(p0, p1);
          }
          tmem_st_32x32b_x16(s_tmem + (c0 >> 1), pk);
        }
        l_sum += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ------------------------------- epilogue --------------------------------------------------
      mbar_wait(&pv_done[t], (n_tiles - 1) & 1);
      tc_fence_after();
      const float inv_l = l_sum > 0.f ? 1.f / l_sum : 0.f;
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t orow[32];
        tmem_ld_32x32b_x32(o_tmem + c0, orow);
        tmem_ld_wait();
        uint8_t* base = stage + (c0 >> 6) * SM::kAtomBytes + row * 128;
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
      if constexpr (!kComm) {
        if (m0t + row < p.Sq) p.lse[((long long)b * p.Hq + hq) * p.Sq + m0t + row] = lse_row;
        fence_proxy_async_smem();
        tc_fence_before();
        named_bar_sync(1 + t, kWG);
        if (row == 0) {
#pragma unroll
          for (int a = 0; a < SM::kAtoms; ++a) tma_store_4d(&omap, stage + a * SM::kAtomBytes, a * 64, m0t, hq, b);
          tma_store_commit();
          tma_store_wait<0>();
        }
      } else {
        tc_fence_before();
        named_bar_sync(1 + t, kWG);
        if (!p.comm.skip_publish) {
          constexpr int CPR = D / 8;
          const size_t so = slot_off(t);
          for (int dst = 0; dst < p.comm.world; ++dst) {
            uint8_t* slot = reinterpret_cast<uint8_t*>(p.comm.data[dst]) + so;
#pragma unroll 4
            for (int c = row; c < kBlockM * CPR; c += kWG) {
              const int r = c / CPR, ch = c - r * CPR;
              const uint4 w = *reinterpret_cast<const uint4*>(stage + (ch >> 3) * SM::kAtomBytes + r * 128 + (((ch & 7) ^ (r & 7)) << 4));
              *reinterpret_cast<uint4*>(slot + (size_t)r * D * 2 + ch * 16) = w;
            }
            reinterpret_cast<float*>(slot + kBlockM * D * 2)[row] = lse_row;
          }
          named_bar_sync(1 + t, kWG);
          if (row < p.comm.world) { fence_acq_rel_sys(); st_release_sys_u32(p.comm.flags[row] + flag_off(t), epoch); }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) { tc_fence_after(); tmem_dealloc<512>(tmem); }
  if constexpr (kComm) comm_kernel_exit(p, epoch);
}

template <int D, bool BF16, bool kComm>
void launch_fwd2(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                 const CommCtxHost& comm, cudaStream_t stream) {
  using SM = Fwd2Smem<D>;
  CUtensorMap qmap = make_tmap_bhsd(q, 2, s.B, s.Hq, s.Sq, D, s.q_sb, s.q_sh, s.q_ss, 64, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap kmap = make_tmap_bhsd(k, 2, s.B, s.Hkv, s.S, D, s.k_sb, s.k_sh, s.k_ss, 64, kBlockN, CU_TENSOR_MAP_SWIZZLE_128B);
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
  p.comm = to_device_ctx(comm);
  const int num_pairs = (p.num_m_tiles + 1) / 2;
  const int n_compute = num_pairs * s.Hq * s.B;
  if (kComm) {
    const size_t slot = (size_t)kBlockM * D * 2 + kBlockM * 4;
    if ((size_t)2 * comm.world * p.n_items * slot > comm.data_bytes || (size_t)2 * comm.world * p.n_items * 4 > comm.flag_bytes)
      throw std::runtime_error("attn_fwd2(fused): symmetric buffer too small");
  }
  auto kern = attn_fwd2_kernel<D, BF16, kComm>;
  static bool configured = false;
  if (!configured) {
    TA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::kTotal));
    configured = true;
  }
  const int grid = n_compute + (kComm ? p.n_items : 0);
  kern<<<grid, kFwd2Threads, SM::kTotal, stream>>>(qmap, kmap, vmap, omap, p, num_pairs, n_compute);
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void attn_fwd2_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                      const CommCtxHost& comm, cudaStream_t stream) {
  if (s.D != 64 && s.D != 128) throw std::runtime_error("attn_fwd2: head_dim must be 64 or 128");
  if (s.Hq % s.Hkv != 0) throw std::runtime_error("attn_fwd2: Hq must be a multiple of Hkv");
  if (s.S <= 0 || s.Sq <= 0) throw std::runtime_error("attn_fwd2: empty problem");
  const bool fused = comm.world > 1;
#define TA_FWD2(DD, BB)                                                               \
  if (fused) launch_fwd2<DD, BB, true>(s, q, k, v, out, lse, comm, stream);           \
  else launch_fwd2<DD, BB, false>(s, q, k, v, out, lse, comm, stream);
  if (s.D == 128) {
    if (s.is_bf16) { TA_FWD2(128, true) } else { TA_FWD2(128, false) }
  } else {
    if (s.is_bf16) { TA_FWD2(64, true) } else { TA_FWD2(64, false) }
  }
#undef TA_FWD2
}

}  // namespace ta
