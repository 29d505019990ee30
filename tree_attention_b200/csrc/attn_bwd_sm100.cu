// attn_bwd -- flash-attention backward for sm_100a on tcgen05/TMEM/TMA.  The reference has no backward
// (SURVEY.md 5.7/7.4); BASELINE.json's "1M GQA fwd+bwd" config needs one.  With Q replicated and KV
// sharded, rank r runs this over its LOCAL shard with the GLOBAL lse: dK_r/dV_r come out complete with
// no communication and dQ_r is a partial that is summed over ranks.
//
// Three kernels (no atomics anywhere => bitwise deterministic):
//   bwd_prep   delta = rowsum(dO * O), lse2 = lse * log2(e), both padded to 64-row multiples
//              (pad rows carry lse2 = +inf so that their probabilities are exactly 0)
//   bwd_dq     query-tile-outer:  S = Q K^T, dP = dO V^T  ->  dS = P o (dP - delta)  ->  dQ += dS K
//   bwd_dkv    kv-tile-outer over (GQA group x query tiles):  S^T = K Q^T, dP^T = V dO^T  ->
//              dV += P^T dO,  dK += dS^T Q      (transposed products put the kv rows on the TMEM lanes)
// Both main kernels are warp-specialised exactly like the forward (4 softmax warps, TMA warp, MMA warp),
// with S/dP double-buffered in TMEM (64-column tiles) so that the tensor pipe works on tile j+1 while the
// SIMT pipes work on tile j.  The same swizzled smem tile is consumed K-major by one MMA and MN-major by
// another (Q and dO in bwd_dkv, K in bwd_dq) -- only the descriptor differs.
// Reference: none -- /root/reference/model.py has no backward; this differentiates the op it defines at model.py:74-80 (local
// attention) and model.py:103-124 (combine).
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace {

constexpr int kBM = 128;       // rows owned by the CTA (q rows in bwd_dq, kv rows in bwd_dkv)
constexpr int kBN = 64;        // streamed tile (kv rows in bwd_dq, q rows in bwd_dkv)
constexpr int kBwdThreads = 192;
constexpr int kBwdStages = 3;

__device__ __forceinline__ float ninf() { return __int_as_float(0xff800000); }
__device__ __forceinline__ float pinf() { return __int_as_float(0x7f800000); }

template <bool BF16>
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  if constexpr (BF16) return pack_bf16x2(lo, hi);
  else return pack_f16x2(lo, hi);
}
template <bool BF16>
__device__ __forceinline__ float ld16(const uint16_t* p) {
  if constexpr (BF16) return __uint_as_float(uint32_t(*p) << 16);
  else return __half2float(__ushort_as_half(*p));
}

// ------------------------------------------------------------------------------------------------
// prep: one warp per (b, h, row)
// ------------------------------------------------------------------------------------------------
struct PrepParams {
  const void* o; const void* dout; const float* lse;
  float* delta; float* lse2;
  int B, Hq, Sq, Sq_pad, D;
  long long o_sb, o_sh, o_ss, d_sb, d_sh, d_ss;
};

template <bool BF16>
__global__ void bwd_prep_kernel(const PrepParams p) {
  const int warps_per_block = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = (long long)p.B * p.Hq * p.Sq_pad;
  if (gw >= total) return;
  const int row = (int)(gw % p.Sq_pad);
  const long long bh = gw / p.Sq_pad;
  const int h = (int)(bh % p.Hq), b = (int)(bh / p.Hq);
  if (row >= p.Sq) {
    if (lane == 0) { p.delta[gw] = 0.f; p.lse2[gw] = pinf(); }
    return;
  }
  const uint16_t* op = reinterpret_cast<const uint16_t*>(p.o) + b * p.o_sb + h * p.o_sh + (long long)row * p.o_ss;
  const uint16_t* dp = reinterpret_cast<const uint16_t*>(p.dout) + b * p.d_sb + h * p.d_sh + (long long)row * p.d_ss;
  float acc = 0.f;
  for (int d = lane; d < p.D; d += 32) acc = fmaf(ld16<BF16>(op + d), ld16<BF16>(dp + d), acc);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    p.delta[gw] = acc;
    const float l = p.lse[(bh)*p.Sq + row];
    // a fully masked row has lse = -inf and o = 0: make its probabilities 0 instead of NaN
    p.lse2[gw] = (l == ninf()) ? pinf() : l * 1.4426950408889634f;
  }
}

// ------------------------------------------------------------------------------------------------
// shared parameter block
// ------------------------------------------------------------------------------------------------
struct BwdParams {
  const float* delta;  // (B, Hq, Sq_pad)
  const float* lse2;   // (B, Hq, Sq_pad)
  float* dq;           // (B, Hq, Sq, D) fp32 contiguous
  void* dk; void* dv;  // (B, Hkv, S, D) contiguous, I/O dtype
  int B, Hq, Hkv, G, Sq, Sq_pad, S, D;
  float scale, scale_log2;
  int causal;
  long long q_pos0, kv_pos0;
  int num_q_tiles128, num_kv_tiles128;
};

template <int D>
struct BwdSmem {
  static constexpr int kAtoms = D / 64;
  static constexpr int kBigBytes = kBM * D * 2;    // 128-row tile
  static constexpr int kSmallBytes = kBN * D * 2;  // 64-row tile
  static constexpr int kBigAtom = kBM * 128;
  static constexpr int kSmallAtom = kBN * 128;
  static constexpr size_t kTotal = 1024 + 2 * size_t(kBigBytes) + 2 * size_t(kBwdStages) * kSmallBytes +
                                   size_t(kBwdStages) * 2 * kBN * 4 + 512;
};

// ------------------------------------------------------------------------------------------------
// bwd_dq : CTA = one 128-row query tile of one (b, hq); streams 64-row K/V tiles
// TMEM: S0 [0,64) dP0 [64,128) | S1 [128,192) dP1 [192,256) | dQ [256, 256+D);  dS_j aliases S_j[0,32)
// ------------------------------------------------------------------------------------------------
template <int D, bool BF16>
__global__ void __launch_bounds__(kBwdThreads, 1)
bwd_dq_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap domap,
              const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
              const BwdParams p) {
  using SM = BwdSmem<D>;
  constexpr int NS = kBwdStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* do_s = q_s + SM::kBigBytes;
  uint8_t* k_s = do_s + SM::kBigBytes;
  uint8_t* v_s = k_s + NS * SM::kSmallBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + NS * SM::kSmallBytes + NS * 2 * kBN * 4);
  uint64_t* qdo_full = bars;           // 1
  uint64_t* k_full = bars + 1;         // NS
  uint64_t* k_empty = k_full + NS;
  uint64_t* v_full = k_empty + NS;
  uint64_t* v_empty = v_full + NS;
  uint64_t* sdp_full = v_empty + NS;   // 2
  uint64_t* ds_full = sdp_full + 2;    // 2
  uint64_t* dq_done = ds_full + 2;     // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tile = p.num_q_tiles128 - 1 - (int)(blockIdx.x % p.num_q_tiles128);
  const int bh = blockIdx.x / p.num_q_tiles128;
  const int hq = bh % p.Hq, b = bh / p.Hq;
  const int hkv = hq / p.G;
  const int m0 = m_tile * kBM;

  int n_end = p.S;
  if (p.causal) {
    const long long last_q = p.q_pos0 + min(m0 + kBM - 1, p.Sq - 1);
    n_end = (int)max(0LL, min((long long)p.S, last_q - p.kv_pos0 + 1));
  }
  const int n_tiles = (n_end + kBN - 1) / kBN;
  float* dq_base = p.dq + (((long long)b * p.Hq + hq) * p.Sq + m0) * D;
  if (n_tiles == 0) {
    if (warp < 4 && m0 + tid < p.Sq)
      for (int d = 0; d < D; d += 4) *reinterpret_cast<float4*>(dq_base + (long long)tid * D + d) = make_float4(0, 0, 0, 0);
    return;
  }

  if (tid == 0) {
    mbar_init(qdo_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&ds_full[i], 4); mbar_init(&dq_done[i], 1); }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_dq = tmem + 256;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(qdo_full, 2 * SM::kBigBytes);
#pragma unroll
      for (int a = 0; a < SM::kAtoms; ++a) {
        tma_load_4d(q_s + a * SM::kBigAtom, &qmap, qdo_full, a * 64, m0, hq, b);
        tma_load_4d(do_s + a * SM::kBigAtom, &domap, qdo_full, a * 64, m0, hq, b);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], SM::kSmallBytes);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(k_s + st * SM::kSmallBytes + a * SM::kSmallAtom, &kmap, &k_full[st], a * 64, j * kBN, hkv, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], SM::kSmallBytes);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(v_s + st * SM::kSmallBytes + a * SM::kSmallAtom, &vmap, &v_full[st], a * 64, j * kBN, hkv, b);
      }
    }
  } else if (warp == 5) {
    // The whole warp runs the issue loop convergently (descriptors stay in uniform registers, no per-MMA R2UR
    // waterfall); one elected lane issues the tcgen05 instructions.
    {
      const bool leader = elect_one();
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc_s = umma_idesc(fmt, fmt, kBM, kBN, 0, 0);   // [128 q] x [64 kv], K = d
      constexpr uint32_t idesc_dq = umma_idesc(fmt, fmt, kBM, D, 0, 1);    // [128 q] x [D], K = 64 kv, B MN-major
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(q_s), 0, 1024), do_desc = umma_smem_desc_sw128(smem_u32(do_s), 0, 1024);
      auto issue_sdp = [&](int j) {
        const int st = j % NS;
        const uint32_t ph = (j / NS) & 1;
        const uint32_t buf = tmem + (j & 1) * 128;
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(k_s + st * SM::kSmallBytes), 0, 1024);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t offb = ((kk / 4) * SM::kBigAtom + (kk % 4) * 32) >> 4, offs = ((kk / 4) * SM::kSmallAtom + (kk % 4) * 32) >> 4;
            umma_ss_f16(buf, q_desc + offb, k_desc + offs, idesc_s, kk > 0 ? 1u : 0u);
          }
        }
        __syncwarp();
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(v_s + st * SM::kSmallBytes), 0, 1024);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t offb = ((kk / 4) * SM::kBigAtom + (kk % 4) * 32) >> 4, offs = ((kk / 4) * SM::kSmallAtom + (kk % 4) * 32) >> 4;
            umma_ss_f16(buf + 64, do_desc + offb, v_desc + offs, idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[st]);
          umma_commit(&sdp_full[j & 1]);
        }
        __syncwarp();
      };
      mbar_wait(qdo_full, 0);
      issue_sdp(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_sdp(j + 1);
        const int st = j % NS;
        mbar_wait(&ds_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(k_s + st * SM::kSmallBytes), SM::kSmallAtom, 1024);
        const uint32_t ds_tmem = tmem + (j & 1) * 128;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kBN / 16; ++kk)
            umma_ts_f16(tmem_dq, ds_tmem + kk * 8, k_desc + ((kk * 2048) >> 4), idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&k_empty[st]);
          umma_commit(&dq_done[j & 1]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    const int row = tid;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const long long q_pos = p.q_pos0 + m0 + row;
    const long long stat_idx = ((long long)b * p.Hq + hq) * p.Sq_pad + m0 + row;
    const bool row_ok = (m0 + row) < p.Sq_pad;
    const float lse2 = row_ok ? p.lse2[stat_idx] : pinf();
    const float delta = row_ok ? p.delta[stat_idx] : 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int n0 = j * kBN;
      mbar_wait(&sdp_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t buf = tmem + (j & 1) * 128 + lane_addr;
      uint32_t sr[64], dpr[64];
      tmem_ld_32x32b_x32(buf + 0, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld_32x32b_x32(buf + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tmem_ld_32x32b_x32(buf + 64, *reinterpret_cast<uint32_t(*)[32]>(&dpr[0]));
      tmem_ld_32x32b_x32(buf + 96, *reinterpret_cast<uint32_t(*)[32]>(&dpr[32]));
      tmem_ld_wait();
      int limc = 63;
      if ((n0 + kBN > p.S) || (p.causal && (p.kv_pos0 + n0 + kBN - 1 > p.q_pos0 + m0))) {
        long long lim = (long long)p.S - n0 - 1;
        if (p.causal) lim = min(lim, q_pos - p.kv_pos0 - n0);
        limc = (int)max(-1LL, min(lim, 63LL));
      }
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        float p0 = fast_exp2(fmaf(__uint_as_float(sr[c]), p.scale_log2, -lse2));
        float p1 = fast_exp2(fmaf(__uint_as_float(sr[c + 1]), p.scale_log2, -lse2));
        p0 = (c <= limc) ? p0 : 0.f;
        p1 = (c + 1 <= limc) ? p1 : 0.f;
        pk[c >> 1] = pk2<BF16>(p0 * (__uint_as_float(dpr[c]) - delta), p1 * (__uint_as_float(dpr[c + 1]) - delta));
      }
      tmem_st_32x32b_x32(buf, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[j & 1]);
    }
    const int jl = n_tiles - 1;
    mbar_wait(&dq_done[jl & 1], (jl >> 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_dq + lane_addr + c0, r);
      tmem_ld_wait();
      if (m0 + row < p.Sq) {
        float* dst = dq_base + (long long)row * D + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(r[i]) * p.scale, __uint_as_float(r[i + 1]) * p.scale,
                                                            __uint_as_float(r[i + 2]) * p.scale, __uint_as_float(r[i + 3]) * p.scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// ------------------------------------------------------------------------------------------------
// bwd_dkv : CTA = one 128-row kv tile of one (b, hkv); streams 64-row Q/dO tiles over the GQA group
// TMEM: S^T0 [0,64) dP^T0 [64,128) | S^T1 [128,192) dP^T1 [192,256) | dV [256,256+D) | dK [384,384+D)
//       P^T_i aliases S^T_i[0,32), dS^T_i aliases dP^T_i[0,32)
// ------------------------------------------------------------------------------------------------
template <int D, bool BF16>
__global__ void __launch_bounds__(kBwdThreads, 1)
bwd_dkv_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap domap,
               const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
               const BwdParams p) {
  using SM = BwdSmem<D>;
  constexpr int NS = kBwdStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* k_s = smem;
  uint8_t* v_s = k_s + SM::kBigBytes;
  uint8_t* q_s = v_s + SM::kBigBytes;
  uint8_t* do_s = q_s + NS * SM::kSmallBytes;
  float* stat_s = reinterpret_cast<float*>(do_s + NS * SM::kSmallBytes);  // [NS][2][64]: lse2, delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(stat_s + NS * 2 * kBN);
  uint64_t* kv_full = bars;            // 1
  uint64_t* q_full = bars + 1;         // NS  (Q tile + lse2 + delta)
  uint64_t* q_empty = q_full + NS;
  uint64_t* do_full = q_empty + NS;
  uint64_t* do_empty = do_full + NS;
  uint64_t* sdp_full = do_empty + NS;  // 2
  uint64_t* ds_full = sdp_full + 2;    // 2
  uint64_t* acc_done = ds_full + 2;    // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x % p.num_kv_tiles128;  // causal: low kv tiles are the heaviest -> first
  const int bh = blockIdx.x / p.num_kv_tiles128;
  const int hkv = bh % p.Hkv, b = bh / p.Hkv;
  const int n0 = n_tile * kBM;

  // query tiles (64 rows) that can see this kv tile
  const int q_tiles_total = p.Sq_pad / kBN;
  int i_begin = 0;
  if (p.causal) {
    const long long first_q = p.kv_pos0 + n0 - p.q_pos0;  // first query row that sees key n0
    i_begin = (int)min((long long)q_tiles_total, max(0LL, first_q) / kBN);
  }
  const int per_head = q_tiles_total - i_begin;
  const int n_iter = per_head * p.G;
  uint16_t* dk_base = reinterpret_cast<uint16_t*>(p.dk) + (((long long)b * p.Hkv + hkv) * p.S + n0) * D;
  uint16_t* dv_base = reinterpret_cast<uint16_t*>(p.dv) + (((long long)b * p.Hkv + hkv) * p.S + n0) * D;
  if (n_iter <= 0) {
    if (warp < 4 && n0 + tid < p.S)
      for (int d = 0; d < D; d += 8) {
        *reinterpret_cast<uint4*>(dk_base + (long long)tid * D + d) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dv_base + (long long)tid * D + d) = make_uint4(0, 0, 0, 0);
      }
    return;
  }

  if (tid == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&do_full[i], 1); mbar_init(&do_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&ds_full[i], 4); mbar_init(&acc_done[i], 1); }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_dv = tmem + 256, tmem_dk = tmem + 384;

  auto iter_to = [&](int it, int& g, int& i) { g = it / per_head; i = i_begin + (it - g * per_head); };

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * SM::kBigBytes);
#pragma unroll
      for (int a = 0; a < SM::kAtoms; ++a) {
        tma_load_4d(k_s + a * SM::kBigAtom, &kmap, kv_full, a * 64, n0, hkv, b);
        tma_load_4d(v_s + a * SM::kBigAtom, &vmap, kv_full, a * 64, n0, hkv, b);
      }
      for (int it = 0; it < n_iter; ++it) {
        int g, i;
        iter_to(it, g, i);
        const int hq = hkv * p.G + g;
        const int st = it % NS;
        const uint32_t ph = (it / NS) & 1;
        mbar_wait(&q_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&q_full[st], SM::kSmallBytes + 2 * kBN * 4);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(q_s + st * SM::kSmallBytes + a * SM::kSmallAtom, &qmap, &q_full[st], a * 64, i * kBN, hq, b);
        const long long sidx = ((long long)b * p.Hq + hq) * p.Sq_pad + (long long)i * kBN;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(stat_s + st * 2 * kBN)), "l"(p.lse2 + sidx), "r"(kBN * 4), "r"(smem_u32(&q_full[st])) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(stat_s + st * 2 * kBN + kBN)), "l"(p.delta + sidx), "r"(kBN * 4), "r"(smem_u32(&q_full[st])) : "memory");
        mbar_wait(&do_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&do_full[st], SM::kSmallBytes);
#pragma unroll
        for (int a = 0; a < SM::kAtoms; ++a)
          tma_load_4d(do_s + st * SM::kSmallBytes + a * SM::kSmallAtom, &domap, &do_full[st], a * 64, i * kBN, hq, b);
      }
    }
  } else if (warp == 5) {
    {
      const bool leader = elect_one();   // whole warp convergent, one elected lane issues (see bwd_dq_kernel)
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc_s = umma_idesc(fmt, fmt, kBM, kBN, 0, 0);    // [128 kv] x [64 q], K = d
      constexpr uint32_t idesc_acc = umma_idesc(fmt, fmt, kBM, D, 0, 1);    // [128 kv] x [D], K = 64 q, B MN-major
      const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(k_s), 0, 1024), v_desc = umma_smem_desc_sw128(smem_u32(v_s), 0, 1024);
      auto issue_sdp = [&](int it) {
        const int st = it % NS;
        const uint32_t ph = (it / NS) & 1;
        const uint32_t buf = tmem + (it & 1) * 128;
        mbar_wait(&q_full[st], ph);
        tc_fence_after();
        const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(q_s + st * SM::kSmallBytes), 0, 1024);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t offb = ((kk / 4) * SM::kBigAtom + (kk % 4) * 32) >> 4, offs = ((kk / 4) * SM::kSmallAtom + (kk % 4) * 32) >> 4;
            umma_ss_f16(buf, k_desc + offb, q_desc + offs, idesc_s, kk > 0 ? 1u : 0u);
          }
        }
        __syncwarp();
        mbar_wait(&do_full[st], ph);
        tc_fence_after();
        const uint64_t do_desc = umma_smem_desc_sw128(smem_u32(do_s + st * SM::kSmallBytes), 0, 1024);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t offb = ((kk / 4) * SM::kBigAtom + (kk % 4) * 32) >> 4, offs = ((kk / 4) * SM::kSmallAtom + (kk % 4) * 32) >> 4;
            umma_ss_f16(buf + 64, v_desc + offb, do_desc + offs, idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&sdp_full[it & 1]);
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_iter; ++it) {
        if (it + 1 < n_iter) issue_sdp(it + 1);
        const int st = it % NS;
        mbar_wait(&ds_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(q_s + st * SM::kSmallBytes), SM::kSmallAtom, 1024);
        const uint64_t do_desc = umma_smem_desc_sw128(smem_u32(do_s + st * SM::kSmallBytes), SM::kSmallAtom, 1024);
        const uint32_t buf = tmem + (it & 1) * 128;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kBN / 16; ++kk)   // dV += P^T dO
            umma_ts_f16(tmem_dv, buf + kk * 8, do_desc + ((kk * 2048) >> 4), idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&do_empty[st]);
#pragma unroll
          for (int kk = 0; kk < kBN / 16; ++kk)   // dK += dS^T Q
            umma_ts_f16(tmem_dk, buf + 64 + kk * 8, q_desc + ((kk * 2048) >> 4), idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&q_empty[st]);
          umma_commit(&acc_done[it & 1]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    const int row = tid;  // kv row == TMEM lane
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const bool row_in = (n0 + row) < p.S;
    const long long kv_pos = p.kv_pos0 + n0 + row;
    for (int it = 0; it < n_iter; ++it) {
      int g, i;
      iter_to(it, g, i);
      const int st = it % NS;
      // lse2 / delta of this query tile arrive with the Q tile (same barrier); the MMA warp waited on it
      // before issuing, and sdp_full is signalled after those MMAs complete.
      mbar_wait(&sdp_full[it & 1], (it >> 1) & 1);
      tc_fence_after();
      const float* lse2_s = stat_s + st * 2 * kBN;
      const float* delta_s = lse2_s + kBN;
      const uint32_t buf = tmem + (it & 1) * 128 + lane_addr;
      uint32_t sr[64], dpr[64];
      tmem_ld_32x32b_x32(buf + 0, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld_32x32b_x32(buf + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tmem_ld_32x32b_x32(buf + 64, *reinterpret_cast<uint32_t(*)[32]>(&dpr[0]));
      tmem_ld_32x32b_x32(buf + 96, *reinterpret_cast<uint32_t(*)[32]>(&dpr[32]));
      tmem_ld_wait();
      // column c is query i*64 + c; visible iff kv_pos <= q_pos
      int cmin = 0;
      if (p.causal) {
        const long long first = kv_pos - p.q_pos0 - (long long)i * kBN;  // first visible column
        cmin = (int)max(0LL, min(first, 64LL));
      }
      if (!row_in) cmin = 64;
      uint32_t ppk[32], dspk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(lse2_s + c);
        const float4 d4 = *reinterpret_cast<const float4*>(delta_s + c);
        float p0 = fast_exp2(fmaf(__uint_as_float(sr[c + 0]), p.scale_log2, -l4.x));
        float p1 = fast_exp2(fmaf(__uint_as_float(sr[c + 1]), p.scale_log2, -l4.y));
        float p2 = fast_exp2(fmaf(__uint_as_float(sr[c + 2]), p.scale_log2, -l4.z));
        float p3 = fast_exp2(fmaf(__uint_as_float(sr[c + 3]), p.scale_log2, -l4.w));
        p0 = (c + 0 >= cmin) ? p0 : 0.f;
        p1 = (c + 1 >= cmin) ? p1 : 0.f;
        p2 = (c + 2 >= cmin) ? p2 : 0.f;
        p3 = (c + 3 >= cmin) ? p3 : 0.f;
        ppk[(c >> 1) + 0] = pk2<BF16>(p0, p1);
        ppk[(c >> 1) + 1] = pk2<BF16>(p2, p3);
        dspk[(c >> 1) + 0] = pk2<BF16>(p0 * (__uint_as_float(dpr[c + 0]) - d4.x), p1 * (__uint_as_float(dpr[c + 1]) - d4.y));
        dspk[(c >> 1) + 1] = pk2<BF16>(p2 * (__uint_as_float(dpr[c + 2]) - d4.z), p3 * (__uint_as_float(dpr[c + 3]) - d4.w));
      }
      tmem_st_32x32b_x32(buf, ppk);
      tmem_st_32x32b_x32(buf + 64, dspk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[it & 1]);
    }
    const int il = n_iter - 1;
    mbar_wait(&acc_done[il & 1], (il >> 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t rv[32], rk[32];
      tmem_ld_32x32b_x32(tmem_dv + lane_addr + c0, rv);
      tmem_ld_32x32b_x32(tmem_dk + lane_addr + c0, rk);
      tmem_ld_wait();
      if (row_in) {
#pragma unroll
        for (int i8 = 0; i8 < 32; i8 += 8) {
          uint4 wv, wk;
          wv.x = pk2<BF16>(__uint_as_float(rv[i8 + 0]), __uint_as_float(rv[i8 + 1]));
          wv.y = pk2<BF16>(__uint_as_float(rv[i8 + 2]), __uint_as_float(rv[i8 + 3]));
          wv.z = pk2<BF16>(__uint_as_float(rv[i8 + 4]), __uint_as_float(rv[i8 + 5]));
          wv.w = pk2<BF16>(__uint_as_float(rv[i8 + 6]), __uint_as_float(rv[i8 + 7]));
          wk.x = pk2<BF16>(__uint_as_float(rk[i8 + 0]) * p.scale, __uint_as_float(rk[i8 + 1]) * p.scale);
          wk.y = pk2<BF16>(__uint_as_float(rk[i8 + 2]) * p.scale, __uint_as_float(rk[i8 + 3]) * p.scale);
          wk.z = pk2<BF16>(__uint_as_float(rk[i8 + 4]) * p.scale, __uint_as_float(rk[i8 + 5]) * p.scale);
          wk.w = pk2<BF16>(__uint_as_float(rk[i8 + 6]) * p.scale, __uint_as_float(rk[i8 + 7]) * p.scale);
          *reinterpret_cast<uint4*>(dv_base + (long long)row * D + c0 + i8) = wv;
          *reinterpret_cast<uint4*>(dk_base + (long long)row * D + c0 + i8) = wk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int D, bool BF16>
void launch_bwd(const AttnShape& s, const void* q, const void* k, const void* v, const void* o, const void* dout,
                const float* lse, float* dq, void* dk, void* dv, float* delta, float* lse2, int64_t do_sb, int64_t do_sh,
                int64_t do_ss, cudaStream_t stream) {
  using SM = BwdSmem<D>;
  const int Sq_pad = (s.Sq + kBN - 1) / kBN * kBN;
  PrepParams pp;
  pp.o = o; pp.dout = dout; pp.lse = lse; pp.delta = delta; pp.lse2 = lse2;
  pp.B = s.B; pp.Hq = s.Hq; pp.Sq = s.Sq; pp.Sq_pad = Sq_pad; pp.D = D;
  pp.o_sb = s.o_sb; pp.o_sh = s.o_sh; pp.o_ss = s.o_ss; pp.d_sb = do_sb; pp.d_sh = do_sh; pp.d_ss = do_ss;
  const long long rows = (long long)s.B * s.Hq * Sq_pad;
  bwd_prep_kernel<BF16><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(pp);
  TA_CUDA_CHECK(cudaGetLastError());

  // Q / dO are streamed in 64-row tiles by bwd_dkv and loaded as 128-row tiles by bwd_dq: two maps each.
  CUtensorMap q128 = make_tmap_bhsd(q, 2, s.B, s.Hq, s.Sq, D, s.q_sb, s.q_sh, s.q_ss, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap do128 = make_tmap_bhsd(dout, 2, s.B, s.Hq, s.Sq, D, do_sb, do_sh, do_ss, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap q64 = make_tmap_bhsd(q, 2, s.B, s.Hq, s.Sq, D, s.q_sb, s.q_sh, s.q_ss, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap do64 = make_tmap_bhsd(dout, 2, s.B, s.Hq, s.Sq, D, do_sb, do_sh, do_ss, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap k128 = make_tmap_bhsd(k, 2, s.B, s.Hkv, s.S, D, s.k_sb, s.k_sh, s.k_ss, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap v128 = make_tmap_bhsd(v, 2, s.B, s.Hkv, s.S, D, s.v_sb, s.v_sh, s.v_ss, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap k64 = make_tmap_bhsd(k, 2, s.B, s.Hkv, s.S, D, s.k_sb, s.k_sh, s.k_ss, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap v64 = make_tmap_bhsd(v, 2, s.B, s.Hkv, s.S, D, s.v_sb, s.v_sh, s.v_ss, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B);

  BwdParams p;
  p.delta = delta; p.lse2 = lse2; p.dq = dq; p.dk = dk; p.dv = dv;
  p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.G = s.Hq / s.Hkv; p.Sq = s.Sq; p.Sq_pad = Sq_pad; p.S = s.S; p.D = D;
  p.scale = s.softmax_scale; p.scale_log2 = s.softmax_scale * 1.4426950408889634f;
  p.causal = s.causal; p.q_pos0 = s.q_pos0; p.kv_pos0 = s.kv_pos0;
  p.num_q_tiles128 = (s.Sq + kBM - 1) / kBM;
  p.num_kv_tiles128 = (s.S + kBM - 1) / kBM;
  static bool configured = false;
  if (!configured) {
    TA_CUDA_CHECK(cudaFuncSetAttribute(bwd_dq_kernel<D, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::kTotal));
    TA_CUDA_CHECK(cudaFuncSetAttribute(bwd_dkv_kernel<D, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::kTotal));
    configured = true;
  }
  bwd_dkv_kernel<D, BF16><<<p.num_kv_tiles128 * s.Hkv * s.B, kBwdThreads, SM::kTotal, stream>>>(q64, do64, k128, v128, p);
  TA_CUDA_CHECK(cudaGetLastError());
  bwd_dq_kernel<D, BF16><<<p.num_q_tiles128 * s.Hq * s.B, kBwdThreads, SM::kTotal, stream>>>(q128, do128, k64, v64, p);
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void attn_bwd_launch(const AttnShape& s, const void* q, const void* k, const void* v, const void* o, const void* dout,
                     const float* lse, float* dq, void* dk, void* dv, float* delta, float* lse2, int64_t do_sb,
                     int64_t do_sh, int64_t do_ss, cudaStream_t stream) {
  if (s.D != 64 && s.D != 128) throw std::runtime_error("attn_bwd: head_dim must be 64 or 128");
  if (s.Hq % s.Hkv != 0) throw std::runtime_error("attn_bwd: Hq must be a multiple of Hkv");
#define TA_BWD(DD, BB) launch_bwd<DD, BB>(s, q, k, v, o, dout, lse, dq, dk, dv, delta, lse2, do_sb, do_sh, do_ss, stream)
  if (s.D == 128) { if (s.is_bf16) TA_BWD(128, true); else TA_BWD(128, false); }
  else { if (s.is_bf16) TA_BWD(64, true); else TA_BWD(64, false); }
#undef TA_BWD
}

}  // namespace ta
