// attn_tree_fused_decode (tensor-core variant) -- split-KV decode on tcgen05 with the tree combine in the epilogue.
//
// The CUDA-core decode kernel (decode_simt.cu) is HBM-bound for MHA (1 query row per KV head) but issue-bound once
// a KV head serves several query rows (GQA 32q/8kv: 51-60 % of HBM peak, profiles/r1_decode/).  Here the R = G x Sq
// query rows of one KV head are PACKED into one 128-row MMA tile (rows >= R are zero padding), so the tensor pipe
// does all the math at a cost that does not depend on R and the kernel stays HBM-bound:
//
//   * persistent CTAs, stream-K over the flattened (batch, kv-head, 128-row KV tile) space -- same split, workspace
//     and ticket scheme as decode_simt.cu; the TMA warp streams K/V tiles continuously across head boundaries;
//   * per head segment: the softmax warps stage the packed Q tile into 128B-swizzled smem (generic stores +
//     fence.proxy.async), the MMA thread runs S = Q K^T (SS) / O += P V (TS, P from TMEM) exactly like
//     attn_fwd_sm100.cu (double-buffered S, lazy rescale), only warps that own valid rows do softmax work;
//   * epilogue: (O, m, l) of the valid rows -> workspace -> atomic ticket -> the last CTA of the head merges the
//     splits in part order -> output, or LL-tagged 8-byte words to every peer + deferred cross-GPU merge.
// Reference: replaces flash_res_lse (/root/reference/model.py:60-83) and the combine of tree_decode (model.py:85-124) for up to
// 128 packed query rows per KV head.
#include "common.cuh"
#include "decode_comm.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace {

constexpr int kTM = 128;  // MMA M (packed query rows, zero padded)
constexpr int kTN = 128;  // KV rows per tile
constexpr int kTcThreads = 192;
constexpr int kSmx = 128;
constexpr int kTcStages = 3;
constexpr int kTcMaxPending = 32;
constexpr float kTcRescale = 8.0f;

struct DecodeTcParams {
  const void* q;
  void* out;
  float* lse;
  uint64_t* part;             // workspace: tagged partial words (decode_comm.cuh)
  unsigned long long* wctr;   // workspace arrival counter (launch tag of the partial words), zero-initialised once
  const float* kscale;  // KV8: per-channel scales (B, Hkv, D) of the e4m3 K / V shards
  const float* vscale;
  const int* kv_len;      // optional device scalar: valid rows of this shard (<= S); rows past it inside the last tile
                          // must hold FINITE data (the tensor pipe multiplies them by exact zeros)
  int B, Hq, Hkv, G, Sq, S, R;   // S = capacity of the shard (rows covered by the tensor maps)
  float scale_log2;
  int causal;
  long long q_pos0, kv_pos0;
  long long q_sb, q_sh, q_ss, o_sb, o_sh, o_ss;
  int max_parts;
  CommCtx comm;
};

// KV8: K, V and the staged Q are e4m3 (1 byte): a 128-row tile of D = 128 is ONE 128B-swizzle atom
template <int D, bool KV8>
struct TcSmem {
  static constexpr int kElem = KV8 ? 1 : 2;
  static constexpr int kAtoms = D * kElem / 128;
  static constexpr int kTileBytes = 128 * D * kElem;
  static constexpr int kAtomBytes = 128 * 128;
  static constexpr int kStages = KV8 ? 6 : kTcStages;
  static constexpr size_t kTotal = 1024 + size_t(1 + 2 * kStages) * kTileBytes + 512 + (KV8 ? 2 * D * 4 : 0);
};

__device__ __forceinline__ float tc_ninf() { return __int_as_float(0xff800000); }
template <bool BF16>
__device__ __forceinline__ uint32_t tc_pack2(float lo, float hi) {
  if constexpr (BF16) return pack_bf16x2(lo, hi);
  else return pack_f16x2(lo, hi);
}
template <bool BF16>
__device__ __forceinline__ uint16_t tc_to16(float f) {
  if constexpr (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(f));
  else return __half_as_ushort(__float2half_rn(f));
}

template <int D, bool BF16, bool KV8>
__global__ void __launch_bounds__(kTcThreads, 1)
decode_tc_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                 const __grid_constant__ DecodeTcParams p) {
  using SM = TcSmem<D, KV8>;
  constexpr int NS = SM::kStages;
  constexpr int EPA = 128 / SM::kElem;  // elements per 128-byte atom row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + SM::kTileBytes;
  uint8_t* v_s = k_s + NS * SM::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + NS * SM::kTileBytes);
  uint64_t* q_ready = bars;            // 1 (count 128)
  uint64_t* k_full = bars + 1;         // NS
  uint64_t* k_empty = k_full + NS;
  uint64_t* v_full = k_empty + NS;
  uint64_t* v_empty = v_full + NS;
  uint64_t* s_full = v_empty + NS;     // 2
  uint64_t* p_full = s_full + 2;       // 2
  uint64_t* pv_done = p_full + 2;      // 2
  uint64_t* stamps = pv_done + 2;      // 2: globaltimer at CTA start / last publish (thread 0)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stamps + 2);
  int* s_misc = reinterpret_cast<int*>(tmem_slot + 2);  // [0] ticket, [1] n_pending, [2] inline-combine head+1
  int* s_tags = s_misc + 4;      // [0] intra-GPU tag, [1] cross-GPU tag, [2] ready (written by the TMA thread)
  int* pending = s_tags + 4;                            // [kTcMaxPending]
  [[maybe_unused]] float* ch_scale = reinterpret_cast<float*>(pending + kTcMaxPending);  // KV8: [2][D] K / V channel scales

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int world = p.comm.world;
  const int BH = p.B * p.Hkv;
  const int R = p.R;
  const dcomm::Geom geo = dcomm::make_geom(p.S, p.kv_len, BH, gridDim.x);
  const int t_lo = dcomm::cta_lo(geo, cta), t_hi = dcomm::cta_lo(geo, cta + 1);
  int jvis = geo.tph;   // tiles of a head that hold visible keys (causal prefix)
  if (p.causal) {
    const long long last = p.q_pos0 + p.Sq - 1 - p.kv_pos0;   // last visible local key index
    jvis = (int)max(0LL, min((long long)geo.tph, last < 0 ? 0LL : last / kTN + 1));
  }
  const int n_active_warps = (R + 31) / 32;

  if (tid == 0) {
    mbar_init(q_ready, kSmx);
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], n_active_warps); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
    s_misc[0] = 0; s_misc[1] = 0; s_tags[2] = 0;
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&kmap); tma_prefetch_desc(&vmap); }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  // zero the whole Q tile once: rows >= R stay zero padding, rows < R are rewritten per head segment
  if (warp < 4) {
    for (int c = tid; c < SM::kTileBytes / 16; c += kSmx) reinterpret_cast<uint4*>(q_s)[c] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + 256;

  if (tid == 0) { stamps[0] = globaltimer_ns(); stamps[1] = 0; }

  // segment iteration shared by all roles: consecutive tiles of one head, clipped to the visible (causal) prefix
  auto next_segment = [&](int t, int& x, int& j0, int& n, int& t_next) {
    x = t / geo.tph;
    const int seg_end = min(t_hi, (x + 1) * geo.tph);
    j0 = t - x * geo.tph;
    const int j1 = seg_end - x * geo.tph;
    n = max(0, min(j1, jvis) - j0);
    t_next = seg_end;
  };

  if (warp == 4) {
    // =============================== TMA producer ===============================================
    if (lane == 0) {
      int it = 0;
      bool tags_done = false;
      // launch tags (decode_comm.cuh): arrival atomics issued once the ring is full (the producer would block on the
      // first `empty` barrier anyway), so their round trip is never exposed
      auto fetch_tags = [&]() {
        const uint32_t wtag = dcomm::launch_tag(p.wctr);
        const uint32_t ctag = world > 1 ? dcomm::launch_tag(reinterpret_cast<unsigned long long*>(p.comm.epoch)) : 0u;
        volatile int* sm = s_tags;
        sm[0] = (int)wtag; sm[1] = (int)ctag;
        __threadfence_block();
        sm[2] = 1;
        tags_done = true;
      };
      for (int t = t_lo; t < t_hi;) {
        int x, j0, n, tn;
        next_segment(t, x, j0, n, tn);
        const int b = x / p.Hkv, h = x - b * p.Hkv;
        for (int jj = 0; jj < n; ++jj, ++it) {
          const int st = it % NS;
          const uint32_t ph = (it / NS) & 1;
          if (it == NS && !tags_done) fetch_tags();
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], SM::kTileBytes);
#pragma unroll
          for (int a = 0; a < SM::kAtoms; ++a)
            tma_load_4d(k_s + st * SM::kTileBytes + a * SM::kAtomBytes, &kmap, &k_full[st], a * EPA, (j0 + jj) * kTN, h, b);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], SM::kTileBytes);
#pragma unroll
          for (int a = 0; a < SM::kAtoms; ++a)
            tma_load_4d(v_s + st * SM::kTileBytes + a * SM::kAtomBytes, &vmap, &v_full[st], a * EPA, (j0 + jj) * kTN, h, b);
        }
        t = tn;
      }
      if (!tags_done) fetch_tags();
    }
  } else if (warp == 5) {
    // =============================== MMA issuer =================================================
    {
      // whole warp convergent (descriptors in uniform registers, no per-MMA R2UR waterfall); one elected lane issues
      const bool leader = elect_one();
      constexpr uint32_t fmt = KV8 ? 0u : (BF16 ? 1u : 0u);  // kind::f8f6f4: 0 = e4m3 ; kind::f16: 0 = f16, 1 = bf16
      constexpr uint32_t idesc_qk = umma_idesc(fmt, fmt, kTM, kTN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc(fmt, fmt, kTM, D, 0, 1);
      constexpr int KSTEP = 32 / SM::kElem;     // elements per MMA along K (32 bytes)
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(q_s), 0, 1024);
      auto issue_qk = [&](int i) {
        const int st = i % NS;
        mbar_wait(&k_full[st], (i / NS) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(k_s + st * SM::kTileBytes), 0, 1024);
        const uint32_t d_tmem = tmem + (i & 1) * 128;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / KSTEP; ++kk) {
            const uint32_t off = ((kk / 4) * SM::kAtomBytes + (kk % 4) * 32) >> 4;
            if constexpr (KV8) umma_ss_f8(d_tmem, q_desc + off, k_desc + off, idesc_qk, kk > 0 ? 1u : 0u);
            else umma_ss_f16(d_tmem, q_desc + off, k_desc + off, idesc_qk, kk > 0 ? 1u : 0u);
          }
          umma_commit(&k_empty[st]);
          umma_commit(&s_full[i & 1]);
        }
        __syncwarp();
      };
      int it = 0, seg = 0;
      for (int t = t_lo; t < t_hi;) {
        int x, j0, n, tn;
        next_segment(t, x, j0, n, tn);
        mbar_wait(q_ready, seg & 1);   // the packed Q tile of this head is in smem (and visible to the async proxy)
        tc_fence_after();
        if (n > 0) {
          issue_qk(it);
          for (int jj = 0; jj < n; ++jj) {
            const int i = it + jj;
            if (jj + 1 < n) issue_qk(i + 1);
            const int st = i % NS;
            mbar_wait(&p_full[i & 1], (i >> 1) & 1);
            mbar_wait(&v_full[st], (i / NS) & 1);
            tc_fence_after();
            const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(v_s + st * SM::kTileBytes), kTN * 128, 1024);
            const uint32_t p_tmem = tmem + (i & 1) * 128;
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < kTN / KSTEP; ++kk) {
                // one MMA consumes 32 bytes of P per row (8 TMEM columns) and KSTEP rows of V (KSTEP x 128 B)
                if constexpr (KV8) umma_ts_f8(tmem_o, p_tmem + kk * 8, v_desc + ((kk * (KSTEP * 128)) >> 4), idesc_pv, (jj > 0 || kk > 0) ? 1u : 0u);
                else umma_ts_f16(tmem_o, p_tmem + kk * 8, v_desc + ((kk * (KSTEP * 128)) >> 4), idesc_pv, (jj > 0 || kk > 0) ? 1u : 0u);
              }
              umma_commit(&v_empty[st]);
              umma_commit(&pv_done[i & 1]);
            }
            __syncwarp();
          }
          it += n;
        }
        ++seg;
        t = tn;
      }
    }
    __syncwarp();
  } else {
    // =============================== softmax / epilogue warps ===================================
    const int row = tid;  // packed query row == TMEM lane;  row r -> (g = r / Sq, i = r % Sq)
    const bool row_valid = row < R;
    const bool warp_active = (warp * 32) < R;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const int g_row = row / p.Sq, i_row = row - g_row * p.Sq;
    const long long q_pos = p.q_pos0 + i_row;

    auto store_out = [&](int x, int r, int d, float o_norm, float lse2) {
      const int b = x / p.Hkv, h = x - b * p.Hkv;
      const int g = r / p.Sq, i = r - g * p.Sq;
      uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)(h * p.G + g) * p.o_sh +
                     (long long)i * p.o_ss + d;
      *op = tc_to16<BF16>(o_norm);
      if (d == 0 && p.lse != nullptr) p.lse[((long long)b * p.Hq + (h * p.G + g)) * p.Sq + i] = lse2 * 0.6931471805599453f;
    };
    // split merge + LL-word cross-GPU combine: decode_comm.cuh (shared with decode_simt.cu / decode_swap_sm100.cu)
    int n_pend = 0;   // heads queued for the merge in drain (uniform across the softmax threads)
    auto make_tail = [&]() {   // built on demand: keeps the tail's bookkeeping out of the tile loop's live registers
      dcomm::Tail tl;
      volatile int* sm = s_tags;
      while (sm[2] == 0) { }    // launch tags fetched by the TMA thread (long done by the first segment end)
      tl.comm = &p.comm; tl.part = p.part; tl.max_parts = p.max_parts; tl.BH = BH;
      tl.R = R; tl.rows_valid = R; tl.wtag = (uint32_t)sm[0]; tl.ctag = (uint32_t)sm[1]; tl.geo = geo;
      tl.pending = pending; tl.max_pending = kTcMaxPending; tl.stamps = stamps;
      return tl;
    };

    int it = 0;
    for (int t = t_lo; t < t_hi;) {
      int x, j0, n, tn;
      next_segment(t, x, j0, n, tn);
      const int b = x / p.Hkv, h = x - b * p.Hkv;
      // ---- stage the packed Q tile of head x: thread r writes row r (global 16-byte loads -> swizzled smem)
      [[maybe_unused]] float q_scale = 1.f;  // KV8: per-row scale of the e4m3 query
      if constexpr (KV8) {
        // per-channel scales of this head's K and V shards; K's are folded into q before it is quantised
        if (tid < D) {
          ch_scale[tid] = __ldg(p.kscale + (long long)x * D + tid);
          ch_scale[D + tid] = __ldg(p.vscale + (long long)x * D + tid);
        }
        named_bar_sync(1, kSmx);
      }
      if (row_valid) {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.q) + (long long)b * p.q_sb +
                                                          (long long)(h * p.G + g_row) * p.q_sh + (long long)i_row * p.q_ss);
        if constexpr (!KV8) {
#pragma unroll
          for (int ch = 0; ch < D / 8; ++ch)
            *reinterpret_cast<uint4*>(q_s + (ch >> 3) * SM::kAtomBytes + row * 128 + (((ch & 7) ^ (row & 7)) << 4)) = __ldg(src + ch);
        } else {
          float qf[D];
          float amax = 0.f;
#pragma unroll
          for (int ch = 0; ch < D / 8; ++ch) {
            const uint4 w = __ldg(src + ch);
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float lo, hi;
              if constexpr (BF16) { lo = bf16lo(ws[u]); hi = bf16hi(ws[u]); } else { lo = f16lo(ws[u]); hi = f16hi(ws[u]); }
              lo *= ch_scale[ch * 8 + 2 * u]; hi *= ch_scale[ch * 8 + 2 * u + 1];
              qf[ch * 8 + 2 * u] = lo; qf[ch * 8 + 2 * u + 1] = hi;
              amax = fmaxf(amax, fmaxf(fabsf(lo), fabsf(hi)));
            }
          }
          q_scale = amax > 0.f ? amax * (1.f / 448.f) : 1.f;
          const float inv = 1.f / q_scale;
#pragma unroll
          for (int ch = 0; ch < D / 16; ++ch) {  // 16 e4m3 per 16-byte chunk
            uint4 w;
            w.x = pack_e4m3x4(qf[ch * 16 + 0] * inv, qf[ch * 16 + 1] * inv, qf[ch * 16 + 2] * inv, qf[ch * 16 + 3] * inv);
            w.y = pack_e4m3x4(qf[ch * 16 + 4] * inv, qf[ch * 16 + 5] * inv, qf[ch * 16 + 6] * inv, qf[ch * 16 + 7] * inv);
            w.z = pack_e4m3x4(qf[ch * 16 + 8] * inv, qf[ch * 16 + 9] * inv, qf[ch * 16 + 10] * inv, qf[ch * 16 + 11] * inv);
            w.w = pack_e4m3x4(qf[ch * 16 + 12] * inv, qf[ch * 16 + 13] * inv, qf[ch * 16 + 14] * inv, qf[ch * 16 + 15] * inv);
            *reinterpret_cast<uint4*>(q_s + row * 128 + (((ch & 7) ^ (row & 7)) << 4)) = w;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(q_ready);
      const float sc_row = p.scale_log2 * q_scale;  // log2-domain softmax scale of this row

      float m_used = tc_ninf(), l_sum = 0.f;
      if (warp_active && n > 0) {
        for (int jj = 0; jj < n; ++jj) {
          const int i = it + jj;
          const int n0 = (j0 + jj) * kTN;
          mbar_wait(&s_full[i & 1], (i >> 1) & 1);
          tc_fence_after();
          const uint32_t s_tmem = tmem + (i & 1) * 128 + lane_addr;
          uint32_t sr[128];
          tmem_ld_32x32b_x32(s_tmem + 0, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
          tmem_ld_32x32b_x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
          tmem_ld_32x32b_x32(s_tmem + 64, *reinterpret_cast<uint32_t(*)[32]>(&sr[64]));
          tmem_ld_32x32b_x32(s_tmem + 96, *reinterpret_cast<uint32_t(*)[32]>(&sr[96]));
          tmem_ld_wait();
          const bool need_mask = (n0 + kTN > geo.S) || (p.causal && (p.kv_pos0 + n0 + kTN - 1 > p.q_pos0));
          if (need_mask) {
            long long lim = (long long)geo.S - n0 - 1;
            if (p.causal) lim = min(lim, q_pos - p.kv_pos0 - n0);
            const int limc = (int)max(-1LL, min(lim, 127LL));
#pragma unroll
            for (int c = 0; c < 128; ++c)
              if (c > limc) sr[c] = 0xff800000u;
          }
          float mx8[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) mx8[u] = __uint_as_float(sr[u]);
#pragma unroll
          for (int c = 8; c < 128; c += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) mx8[u] = fmaxf(mx8[u], __uint_as_float(sr[c + u]));
          }
          const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
          const float m_new = fmaxf(m_used, mx * sc_row);
          const bool refresh = (m_new - m_used > kTcRescale) || (m_used == tc_ninf() && m_new != tc_ninf());
          if (__any_sync(0xffffffffu, refresh)) {
            const float alpha = refresh ? fast_exp2(m_used - m_new) : 1.f;
            if (refresh) { l_sum *= alpha; m_used = m_new; }
            if (jj > 0) {
              mbar_wait(&pv_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
              tc_fence_after();
#pragma unroll
              for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t orow[32];
                tmem_ld_32x32b_x32(tmem_o + lane_addr + c0, orow);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 32; ++u) orow[u] = __float_as_uint(__uint_as_float(orow[u]) * alpha);
                tmem_st_32x32b_x32(tmem_o + lane_addr + c0, orow);
              }
            }
          }
          const float neg_m = (m_used == tc_ninf()) ? 0.f : -m_used;
          float ls[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if constexpr (!KV8) {
            uint32_t pk[64];
#pragma unroll
            for (int c = 0; c < 128; c += 8) {
#pragma unroll
              for (int u = 0; u < 8; u += 2) {
                const float p0 = fast_exp2(fmaf(__uint_as_float(sr[c + u]), sc_row, neg_m));
                const float p1 = fast_exp2(fmaf(__uint_as_float(sr[c + u + 1]), sc_row, neg_m));
                ls[u] += p0; ls[u + 1] += p1;
                pk[(c + u) >> 1] = tc_pack2<BF16>(p0, p1);
              }
            }
            tmem_st_32x32b_x32(s_tmem + 0, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
            tmem_st_32x32b_x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
          } else {
            // P <= 2^8 (lazy rescale) fits e4m3 (max 448): four probabilities per 32-bit TMEM column
            uint32_t pk[32];
#pragma unroll
            for (int c = 0; c < 128; c += 8) {
              float pv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) { pv[u] = fast_exp2(fmaf(__uint_as_float(sr[c + u]), sc_row, neg_m)); ls[u] += pv[u]; }
              pk[c >> 2] = pack_e4m3x4(pv[0], pv[1], pv[2], pv[3]);
              pk[(c >> 2) + 1] = pack_e4m3x4(pv[4], pv[5], pv[6], pv[7]);
            }
            tmem_st_32x32b_x32(s_tmem, pk);
          }
          l_sum += ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[i & 1]);
        }
      }
      // ---- segment epilogue: (O, m, l) of the valid rows -> CTA partial in the workspace
      const dcomm::Tail tail = make_tail();
      int nparts, pidx;
      dcomm::head_parts(geo, x, cta, nparts, pidx);
      uint64_t* my_part = dcomm::part_ptr<D>(tail, x, pidx);   // tagged words: no fence, no ticket (decode_comm.cuh)
      if (warp_active) {
        if (n > 0) {
          const int il = it + n - 1;
          mbar_wait(&pv_done[il & 1], (il >> 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
          uint32_t orow[32];
          if (n > 0) {
            tmem_ld_32x32b_x32(tmem_o + lane_addr + c0, orow);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int u = 0; u < 32; ++u) orow[u] = 0u;
          }
          if (row_valid) {
            uint64_t* dst = my_part + row * (D + 2) + c0;
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              float o1 = __uint_as_float(orow[u]);
              if constexpr (KV8) o1 *= ch_scale[D + c0 + u];   // de-quantise the V channels
              dcomm::ll_store_gpu(dst + u, o1, tail.wtag);
            }
          }
        }
        if (row_valid) {
          dcomm::ll_store_gpu(my_part + row * (D + 2) + D, m_used, tail.wtag);
          dcomm::ll_store_gpu(my_part + row * (D + 2) + D + 1, l_sum, tail.wtag);
        }
        tc_fence_before();
      }
      it += n;
      dcomm::segment_done<D, kSmx, 4>(tail, x, tn, n_pend, tid, 1, store_out);   // the owner of the head's last tile queues the merge
      named_bar_sync(1, kSmx);  // Q smem / s_misc reuse by the next segment
      t = tn;
    }
    const dcomm::Tail tail = make_tail();
    dcomm::drain<D, kSmx, 4>(tail, n_pend, tid, 1, store_out);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int D, bool BF16, bool KV8>
std::function<void(cudaStream_t)> make_tc_pass(const CUtensorMap& kmap, const CUtensorMap& vmap, const DecodeTcParams& p, int grid) {
  auto kern = decode_tc_kernel<D, BF16, KV8>;
  static bool configured = false;
  if (!configured) {
    TA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcSmem<D, KV8>::kTotal));
    configured = true;
  }
  return [kern, kmap, vmap, p, grid](cudaStream_t stream) {
    kern<<<grid, kTcThreads, TcSmem<D, KV8>::kTotal, stream>>>(kmap, vmap, p);
    TA_CUDA_CHECK(cudaGetLastError());
  };
}

}  // namespace

// kept for callers of the round-1 name: same split as decode_simt (decode_split, decode_simt.cu)
void decode_tc_split(const AttnShape& s, int ncta, int* grid, int* max_parts) {
  decode_split(s.B * s.Hkv, s.S, ncta, grid, max_parts);
}

void decode_tc_plan(const AttnShape& s, int nsm, int* grid, int* max_parts, int* rows, size_t* part_floats,
                    size_t* comm_bytes) {
  const int BH = s.B * s.Hkv;
  decode_split(BH, s.S, nsm, grid, max_parts);
  const int R = (s.Hq / s.Hkv) * s.Sq;
  *rows = R;
  *part_floats = (size_t)BH * *max_parts * R * (s.D + 2) * 2;   // tagged 8-byte words, counted in floats
  *comm_bytes = (size_t)2 * kMaxWorldHost * BH * R * (s.D + 2) * 8;
}

PreparedLaunch decode_tc_prepare(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                                 float* part, uint32_t* tickets, const CommCtxHost& comm, int nsm, const float* kscale,
                                 const float* vscale, const int* kv_len) {
  const bool kv8 = kscale != nullptr;
  if (s.D != 64 && s.D != 128) throw std::runtime_error("decode_tc: head_dim must be 64 or 128");
  if (kv8 && s.D != 128) throw std::runtime_error("decode_tc(fp8): head_dim must be 128");
  if (s.Hq % s.Hkv != 0) throw std::runtime_error("decode_tc: Hq must be a multiple of Hkv");
  const int G = s.Hq / s.Hkv;
  const int R = G * s.Sq;
  if (R > kTM) throw std::runtime_error("decode_tc: (Hq / Hkv) * Sq must be <= 128");
  if (s.S <= 0) throw std::runtime_error("decode_tc: the KV shard must have capacity for at least one row");
  int grid, max_parts, rows;
  size_t pf, cb;
  decode_tc_plan(s, nsm, &grid, &max_parts, &rows, &pf, &cb);
  if ((long long)s.B * s.Hkv / grid + 2 > kTcMaxPending)
    throw std::runtime_error("decode_tc: batch x kv-heads too large for one launch (split the batch)");
  if (comm.world > 1) {
    const size_t need = (size_t)2 * comm.world * s.B * s.Hkv * R * (s.D + 2) * 8;
    if (need > comm.data_bytes) throw std::runtime_error("decode_tc: symmetric buffer too small for this problem");
  }
  const int eb = kv8 ? 1 : 2;
  CUtensorMap kmap = make_tmap_bhsd(k, eb, s.B, s.Hkv, s.S, s.D, s.k_sb, s.k_sh, s.k_ss, 128 / eb, kTN, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap vmap = make_tmap_bhsd(v, eb, s.B, s.Hkv, s.S, s.D, s.v_sb, s.v_sh, s.v_ss, 128 / eb, kTN, CU_TENSOR_MAP_SWIZZLE_128B);
  DecodeTcParams p;
  p.kscale = kscale; p.vscale = vscale; p.kv_len = kv_len;
  p.q = q; p.out = out; p.lse = lse;
  p.part = reinterpret_cast<uint64_t*>(part); p.wctr = reinterpret_cast<unsigned long long*>(tickets);
  p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.G = G; p.Sq = s.Sq; p.S = s.S; p.R = R;
  p.scale_log2 = s.softmax_scale * 1.4426950408889634f;
  p.causal = s.causal; p.q_pos0 = s.q_pos0; p.kv_pos0 = s.kv_pos0;
  p.q_sb = s.q_sb; p.q_sh = s.q_sh; p.q_ss = s.q_ss; p.o_sb = s.o_sb; p.o_sh = s.o_sh; p.o_ss = s.o_ss;
  p.max_parts = max_parts;
  p.comm = dcomm::to_device_ctx(comm);
  PreparedLaunch pl;
  if (kv8) pl.passes.push_back(s.is_bf16 ? make_tc_pass<128, true, true>(kmap, vmap, p, grid) : make_tc_pass<128, false, true>(kmap, vmap, p, grid));
  else if (s.D == 128) pl.passes.push_back(s.is_bf16 ? make_tc_pass<128, true, false>(kmap, vmap, p, grid) : make_tc_pass<128, false, false>(kmap, vmap, p, grid));
  else pl.passes.push_back(s.is_bf16 ? make_tc_pass<64, true, false>(kmap, vmap, p, grid) : make_tc_pass<64, false, false>(kmap, vmap, p, grid));
  return pl;
}

void decode_tc_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse, float* part,
                      uint32_t* tickets, const CommCtxHost& comm, int nsm, cudaStream_t stream, const float* kscale,
                      const float* vscale, const int* kv_len) {
  decode_tc_prepare(s, q, k, v, out, lse, part, tickets, comm, nsm, kscale, vscale, kv_len).run(stream);
}

}  // namespace ta
