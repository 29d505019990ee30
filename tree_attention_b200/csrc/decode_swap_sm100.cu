// attn_tree_fused_decode, swap-AB formulation -- S^T = K Q^T on tcgen05 with the KEYS on the TMEM lanes.
//
// decode_tc_sm100.cu packs the R query rows into the M dimension: every softmax thread then owns a full 128-key row
// of scores (~2700 cycles per tile, independent of dtype), which caps fp8 KV at ~1.3x over bf16.  Here the tile is
// transposed (SURVEY.md 7.1 "swap-AB"):
//     S^T[128 keys, N = 16 queries] = K_tile (A, K-major)  x  Q^T (B = packed query rows, K-major)
//     O^T[D = 128,  N = 16 queries] += V_tile^T (A, MN-major: the [key][d] tile exactly as TMA delivers it) x P^T (B)
// so a softmax thread (= one key) handles 16 scores instead of 128, the two GEMMs cost 128 tensor cycles per tile
// instead of 1024, and the kernel is HBM-bound for fp8 as well.  Column (per-query) maxima across the 128 keys are
// taken with redux.sync + one smem exchange per tile; row sums stay per-thread until the end of a head segment.
// P^T is written to shared memory by the softmax threads (B operands cannot come from TMEM) in the 128B-swizzled
// K-major layout and made visible with fence.proxy.async.  Split-KV scheduling, workspace, tickets and the LL-tagged
// cross-GPU combine are those of decode_simt.cu / decode_tc_sm100.cu.  R = (Hq/Hkv) x Sq <= 16, head_dim = 128.
// Reference: replaces flash_res_lse (/root/reference/model.py:60-83) and the combine of tree_decode (model.py:85-124) for <= 16
// packed query rows per KV head.
#include "common.cuh"
#include "decode_comm.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace {

constexpr int kSwN = 16;       // query columns of the transposed tile
constexpr int kSwKV = 128;     // keys per tile == MMA M
constexpr int kSwD = 128;
constexpr int kSwThreads = 192;
constexpr int kSmx = 128;
constexpr int kSwMaxPending = 32;
constexpr float kSwRescale = 8.0f;
// MX mode: scale-factor columns in TMEM (4 columns per 128-row operand, see umma_probe.cu)
constexpr uint32_t kSfK = 96;    // + 4 * (tile & 1): K tile scales (A of S^T = K Q^T)
constexpr uint32_t kSfV = 104;   // + 4 * (tile & 1): V tile scales (A of O^T += V^T P^T)
constexpr uint32_t kSfQ = 112;   // packed query rows (B of S^T)
constexpr uint32_t kSfP = 116;   // P^T (B of O^T): constant 2^0
constexpr uint32_t kSfOne = 0x7f7f7f7fu;

struct DecodeSwParams {
  const void* q;
  void* out;
  float* lse;
  uint64_t* part;             // workspace: tagged partial words (decode_comm.cuh)
  unsigned long long* wctr;   // workspace arrival counter (launch tag of the partial words), zero-initialised once
  const float* kscale;  // KV8: per-channel fp32 scales (B, Hkv, D) of the e4m3 K / V shards
  const float* vscale;
  const uint32_t* k_sf;   // MX: one word per key = the 4 UE8M0 scales of its 4 blocks of 32 channels, (B, Hkv, S)
  const uint32_t* v_sf;   // MX: one word per (128-key tile, channel) = the scales of its 4 blocks of 32 keys, (B, Hkv, tiles, D)
  const int* kv_len;      // optional device scalar: valid rows of this shard (<= S); rows past it inside the last tile
                          // must hold FINITE data (the tensor pipe multiplies them by exact zeros)
  int B, Hq, Hkv, G, Sq, S, R;   // S = capacity of the shard (rows covered by the tensor maps)
  float scale_log2;
  int causal;
  long long q_pos0, kv_pos0;
  long long q_sb, q_sh, q_ss, o_sb, o_sh, o_ss;
  int tph_cap;            // tiles per head at capacity (stride of the MX V-scale tensor)
  int max_parts;
  CommCtx comm;
};

template <bool KV8>
struct SwSmem {
  static constexpr int kElem = KV8 ? 1 : 2;
  static constexpr int kAtoms = kSwD * kElem / 128;            // 128-byte atoms per row of K / V / Q
  static constexpr int kTileBytes = kSwKV * kSwD * kElem;      // one K or V tile
  static constexpr int kAtomBytes = kSwKV * 128;
  static constexpr int kQBytes = kSwN * kSwD * kElem;          // packed query rows
  static constexpr int kQAtomBytes = kSwN * 128;
  static constexpr int kPBytes = kSwN * kSwKV * kElem;         // P^T [16 queries][128 keys]
  static constexpr int kPAtoms = kSwKV * kElem / 128;
  static constexpr int kStages = KV8 ? 6 : 3;
  static constexpr size_t kTotal = 1024 + size_t(2 * kStages) * kTileBytes + kQBytes + 2 * kPBytes + 2048 + 1024;
};

__device__ __forceinline__ float sw_ninf() { return __int_as_float(0xff800000); }
template <bool BF16>
__device__ __forceinline__ uint16_t sw_to16(float f) {
  if constexpr (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(f));
  else return __half_as_ushort(__float2half_rn(f));
}
// order-preserving float <-> int so that redux.sync.max.s32 can reduce floats
__device__ __forceinline__ int sw_f2ord(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float sw_ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

template <bool BF16, bool KV8, bool MX>
__global__ void __launch_bounds__(kSwThreads, 1)
decode_swap_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                   const __grid_constant__ DecodeSwParams p) {
  using SM = SwSmem<KV8>;
  constexpr int D = kSwD;
  constexpr int NS = SM::kStages;
  constexpr int EPA = 128 / SM::kElem;    // elements per 128-byte atom row
  constexpr int KSTEP = 32 / SM::kElem;   // elements per MMA along the contraction (32 bytes)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* k_s = smem;
  uint8_t* v_s = k_s + NS * SM::kTileBytes;
  uint8_t* q_s = v_s + NS * SM::kTileBytes;          // [atoms][16 rows][128 B]
  uint8_t* p_s = q_s + SM::kQBytes;                  // [2][atoms][16 rows][128 B]
  float* xch = reinterpret_cast<float*>(p_s + 2 * SM::kPBytes);   // [2 parity][4 warps][16] column maxima
  float* red_s = xch + 2 * 4 * kSwN; // [4 warps][16] row-sum reduction
  uint64_t* bars = reinterpret_cast<uint64_t*>(red_s + 4 * kSwN);
  uint64_t* q_ready = bars;            // count 128
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + NS;
  uint64_t* v_full = k_empty + NS;
  uint64_t* v_empty = v_full + NS;
  uint64_t* s_full = v_empty + NS;     // 2
  uint64_t* p_full = s_full + 2;       // 2 (count 4)
  uint64_t* pv_done = p_full + 2;      // 2
  uint64_t* stamps = pv_done + 2;      // 2: globaltimer at CTA start / last publish (thread 0)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stamps + 2);
  int* s_misc = reinterpret_cast<int*>(tmem_slot + 2);
  int* s_tags = s_misc + 4;      // [0] intra-GPU tag, [1] cross-GPU tag, [2] ready (written by the TMA thread)
  int* pending = s_tags + 4;
  [[maybe_unused]] float* ch_scale = reinterpret_cast<float*>(pending + kSwMaxPending);  // KV8: [2][D]

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
    jvis = (int)max(0LL, min((long long)geo.tph, last < 0 ? 0LL : last / kSwKV + 1));
  }

  if (tid == 0) {
    mbar_init(q_ready, kSmx);
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
    s_misc[0] = 0; s_misc[1] = 0; s_tags[2] = 0;
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&kmap); tma_prefetch_desc(&vmap); }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  if (warp < 4) {  // zero Q (padding rows) and both P^T buffers once
    for (int c = tid; c < (SM::kQBytes + 2 * SM::kPBytes) / 16; c += kSmx) reinterpret_cast<uint4*>(q_s)[c] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;       // S^T0 [0,16) | S^T1 [32,48) | O^T [64,80)
  const uint32_t tmem_o = tmem + 64;

  if (tid == 0) { stamps[0] = globaltimer_ns(); stamps[1] = 0; }

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
            tma_load_4d(k_s + st * SM::kTileBytes + a * SM::kAtomBytes, &kmap, &k_full[st], a * EPA, (j0 + jj) * kSwKV, h, b);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], SM::kTileBytes);
#pragma unroll
          for (int a = 0; a < SM::kAtoms; ++a)
            tma_load_4d(v_s + st * SM::kTileBytes + a * SM::kAtomBytes, &vmap, &v_full[st], a * EPA, (j0 + jj) * kSwKV, h, b);
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
      constexpr uint32_t fmt = KV8 ? 0u : (BF16 ? 1u : 0u);
      constexpr uint32_t idesc_qk = umma_idesc(fmt, fmt, kSwKV, kSwN, 0, 0);   // A = K (K-major), B = Q (K-major)
      constexpr uint32_t idesc_pv = umma_idesc(fmt, fmt, D, kSwN, 1, 0);       // A = V^T (MN-major), B = P^T (K-major)
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(q_s), 0, 1024);
      const uint64_t pt_desc0 = umma_smem_desc_sw128(smem_u32(p_s), 0, 1024);
      auto issue_qk = [&](int i) {
        const int st = i % NS;
        mbar_wait(&k_full[st], (i / NS) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(k_s + st * SM::kTileBytes), 0, 1024);
        const uint32_t d_tmem = tmem + (i & 1) * 32;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / KSTEP; ++kk) {
            const uint64_t ad = k_desc + (((kk / 4) * SM::kAtomBytes + (kk % 4) * 32) >> 4);
            const uint64_t bd = q_desc + (((kk / 4) * SM::kQAtomBytes + (kk % 4) * 32) >> 4);
            if constexpr (MX)
              umma_ss_mxf8_block_scale(d_tmem, ad, bd, umma_idesc_block_scaled(0, 0, kSwKV, kSwN, 0, 0, kk, kk),
                                       tmem + kSfK + 4 * (i & 1), tmem + kSfQ, kk > 0 ? 1u : 0u);
            else if constexpr (KV8) umma_ss_f8(d_tmem, ad, bd, idesc_qk, kk > 0 ? 1u : 0u);
            else umma_ss_f16(d_tmem, ad, bd, idesc_qk, kk > 0 ? 1u : 0u);
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
        mbar_wait(q_ready, seg & 1);
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
            // A: KSTEP keys x 128 d of V^T: rows of the [key][d] tile, MN-major (LBO = next 64-wide d atom)
            const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(v_s + st * SM::kTileBytes), SM::kAtomBytes, 1024);
            const uint64_t pt_desc = pt_desc0 + (((i & 1) * SM::kPBytes) >> 4);
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < kSwKV / KSTEP; ++kk) {
                const uint64_t ad = v_desc + ((kk * (KSTEP * 128)) >> 4);
                const uint64_t bd = pt_desc + (((kk / 4) * SM::kQAtomBytes + (kk % 4) * 32) >> 4);
                if constexpr (MX)
                  umma_ss_mxf8_block_scale(tmem_o, ad, bd, umma_idesc_block_scaled(0, 0, D, kSwN, 1, 0, kk, kk),
                                           tmem + kSfV + 4 * (i & 1), tmem + kSfP, (jj > 0 || kk > 0) ? 1u : 0u);
                else if constexpr (KV8) umma_ss_f8(tmem_o, ad, bd, idesc_pv, (jj > 0 || kk > 0) ? 1u : 0u);
                else umma_ss_f16(tmem_o, ad, bd, idesc_pv, (jj > 0 || kk > 0) ? 1u : 0u);
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
    // =============================== softmax / epilogue warps: thread = key row / output channel ==
    const int row = tid;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    int qi_of[kSwN];  // token index (within Sq) of packed query column qn
#pragma unroll
    for (int qn = 0; qn < kSwN; ++qn) qi_of[qn] = qn % p.Sq;
    auto store_out = [&](int x, int r, int d, float o_norm, float lse2) {
      const int b = x / p.Hkv, h = x - b * p.Hkv;
      const int g = r / p.Sq, i = r - g * p.Sq;
      uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)(h * p.G + g) * p.o_sh +
                     (long long)i * p.o_ss + d;
      *op = sw_to16<BF16>(o_norm);
      if (d == 0 && p.lse != nullptr) p.lse[((long long)b * p.Hq + (h * p.G + g)) * p.Sq + i] = lse2 * 0.6931471805599453f;
    };
    // split merge + LL-word cross-GPU combine: decode_comm.cuh (shared with decode_simt.cu / decode_tc_sm100.cu)
    int n_pend = 0;   // heads queued for the merge in drain (uniform across the softmax threads)
    auto make_tail = [&]() {   // built on demand: keeps the tail's bookkeeping out of the tile loop's live registers
      dcomm::Tail tl;
      volatile int* sm = s_tags;
      while (sm[2] == 0) { }    // launch tags fetched by the TMA thread (long done by the first segment end)
      tl.comm = &p.comm; tl.part = p.part; tl.max_parts = p.max_parts; tl.BH = BH;
      tl.R = R; tl.rows_valid = R; tl.wtag = (uint32_t)sm[0]; tl.ctag = (uint32_t)sm[1]; tl.geo = geo;
      tl.pending = pending; tl.max_pending = kSwMaxPending; tl.stamps = stamps;
      return tl;
    };

    const uint32_t xch_a = smem_u32(xch), pt_a = smem_u32(p_s);
    [[maybe_unused]] uint32_t* sfq_s = reinterpret_cast<uint32_t*>(ch_scale);   // MX: the 16 query-row scale words
    // MX: scale words of K tile j (one per key) / V tile j (one per channel) for the 4 rows 32 c + lane this lane
    // stages into its lane quarter of the scale-factor columns
    [[maybe_unused]] auto load_ksf = [&](int x, int j, uint32_t (&w)[4]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int key = j * kSwKV + 32 * c + lane;
        w[c] = key < p.S ? __ldg(p.k_sf + (long long)x * p.S + key) : kSfOne;   // p.S: capacity (index bound), not the valid length
      }
    };
    [[maybe_unused]] auto load_vsf = [&](int x, int j, uint32_t (&w)[4]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = __ldg(p.v_sf + ((long long)x * p.tph_cap + j) * D + 32 * c + lane);
    };
    if constexpr (MX) {
      tmem_st_32x32b_x4(tmem + lane_addr + kSfP, kSfOne, kSfOne, kSfOne, kSfOne);   // P <= 2^8 fits e4m3 unscaled
      tmem_st_wait();
    }
    int it = 0;
    for (int t = t_lo; t < t_hi;) {
      int x, j0, n, tn;
      next_segment(t, x, j0, n, tn);
      const int b = x / p.Hkv, h = x - b * p.Hkv;
      // ---- stage the packed query rows (B operand, K-major): thread -> (row = tid / 8, 2 chunks of 16 bytes)
      [[maybe_unused]] float q_scale_mine = 1.f;
      if constexpr (KV8 && !MX) {
        if (tid < D) { ch_scale[tid] = __ldg(p.kscale + (long long)x * D + tid); ch_scale[D + tid] = __ldg(p.vscale + (long long)x * D + tid); }
        named_bar_sync(1, kSmx);
      }
      {
        const int qr = tid >> 3, part8 = tid & 7;  // 16 rows x 8 threads; every thread runs the same shuffles
        const bool qv = qr < R;
        const int g = qv ? qr / p.Sq : 0, qi = qv ? qr - g * p.Sq : 0;
        const uint16_t* src = reinterpret_cast<const uint16_t*>(p.q) + (long long)b * p.q_sb + (long long)(h * p.G + g) * p.q_sh +
                              (long long)qi * p.q_ss + part8 * 16;
        if constexpr (!KV8) {
          if (qv) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ch = part8 * 2 + u;  // 16-byte chunk (8 elements) within the 256-byte row
              const uint4 w = __ldg(reinterpret_cast<const uint4*>(src) + u);
              *reinterpret_cast<uint4*>(q_s + (ch >> 3) * SM::kQAtomBytes + qr * 128 + (((ch & 7) ^ (qr & 7)) << 4)) = w;
            }
          }
        } else if constexpr (MX) {
          // MX-quantise the row: one power-of-two scale per 32 channels (two neighbouring threads share a block)
          float qf[16];
          float amax = 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float v = 0.f;
            if (qv) { if constexpr (BF16) v = __uint_as_float(uint32_t(src[u]) << 16); else v = __half2float(__ushort_as_half(src[u])); }
            qf[u] = v;
            amax = fmaxf(amax, fabsf(v));
          }
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
          int e = 0;
          if (amax > 0.f) e = max(-126, min(126, (int)ceilf(log2f(amax * (1.f / 448.f)))));
          const float inv = __int_as_float((127 - e) << 23);   // 2^-e
          uint4 w;
          w.x = pack_e4m3x4(qf[0] * inv, qf[1] * inv, qf[2] * inv, qf[3] * inv);
          w.y = pack_e4m3x4(qf[4] * inv, qf[5] * inv, qf[6] * inv, qf[7] * inv);
          w.z = pack_e4m3x4(qf[8] * inv, qf[9] * inv, qf[10] * inv, qf[11] * inv);
          w.w = pack_e4m3x4(qf[12] * inv, qf[13] * inv, qf[14] * inv, qf[15] * inv);
          if (qv) *reinterpret_cast<uint4*>(q_s + qr * 128 + ((part8 ^ (qr & 7)) << 4)) = w;
          const uint32_t eb = uint32_t(e + 127);
          const int l0 = lane & ~7;
          const uint32_t word = __shfl_sync(0xffffffffu, eb, l0) | (__shfl_sync(0xffffffffu, eb, l0 + 2) << 8) |
                                (__shfl_sync(0xffffffffu, eb, l0 + 4) << 16) | (__shfl_sync(0xffffffffu, eb, l0 + 6) << 24);
          if (part8 == 0) sfq_s[qr] = word;
        } else {
          // fold K's channel scales into q, quantise the row to e4m3 with one scale per row (8 threads cooperate)
          float qf[16];
          float amax = 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float v = 0.f;
            if (qv) {
              if constexpr (BF16) v = __uint_as_float(uint32_t(src[u]) << 16); else v = __half2float(__ushort_as_half(src[u]));
              v *= ch_scale[part8 * 16 + u];
            }
            qf[u] = v;
            amax = fmaxf(amax, fabsf(v));
          }
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
          const float qs = amax > 0.f ? amax * (1.f / 448.f) : 1.f;
          const float inv = 1.f / qs;
          uint4 w;
          w.x = pack_e4m3x4(qf[0] * inv, qf[1] * inv, qf[2] * inv, qf[3] * inv);
          w.y = pack_e4m3x4(qf[4] * inv, qf[5] * inv, qf[6] * inv, qf[7] * inv);
          w.z = pack_e4m3x4(qf[8] * inv, qf[9] * inv, qf[10] * inv, qf[11] * inv);
          w.w = pack_e4m3x4(qf[12] * inv, qf[13] * inv, qf[14] * inv, qf[15] * inv);
          if (qv) *reinterpret_cast<uint4*>(q_s + qr * 128 + ((part8 ^ (qr & 7)) << 4)) = w;
          if (part8 == 0) red_s[qr] = qs;  // per-query scale, read by every softmax thread below
        }
      }
      if constexpr (MX) {
        // scale factors of the query rows and of the first two K tiles of this head -> TMEM, before the MMA warp starts
        named_bar_sync(1, kSmx);
        tmem_st_32x32b_x4(tmem + lane_addr + kSfQ, lane < kSwN ? sfq_s[lane] : kSfOne, kSfOne, kSfOne, kSfOne);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u < n) {
            uint32_t w[4];
            load_ksf(x, j0 + u, w);
            tmem_st_32x32b_x4(tmem + lane_addr + kSfK + 4 * ((it + u) & 1), w[0], w[1], w[2], w[3]);
          }
        }
        tmem_st_wait();
        tc_fence_before();
      }
      fence_proxy_async_smem();
      mbar_arrive(q_ready);
      float sc_q[kSwN];
      if constexpr (KV8 && !MX) {
        named_bar_sync(1, kSmx);
#pragma unroll
        for (int qn = 0; qn < kSwN; ++qn) sc_q[qn] = p.scale_log2 * red_s[qn];
        named_bar_sync(1, kSmx);  // red_s is reused for the row-sum reduction at the end of the segment
      } else {
#pragma unroll
        for (int qn = 0; qn < kSwN; ++qn) sc_q[qn] = p.scale_log2;
      }

      // Running state per query column.  m_used is the (stale) maximum the exponent uses; it is refreshed for all
      // 16 queries at once, and only when some column maximum grew by more than kSwRescale (exact lazy rescale:
      // the same m_used is applied to P and to the row sums, so the final normalisation is exact).
      float m_used[kSwN], neg_ms[kSwN], l_thr[kSwN];
#pragma unroll
      for (int qn = 0; qn < kSwN; ++qn) { m_used[qn] = sw_ninf(); neg_ms[qn] = 0.f; l_thr[qn] = 0.f; }
      // P^T store addresses: element (query qn, key row) of the 128B-swizzled K-major B tile
      uint32_t pa[8];
      {
        const int ch = KV8 ? (row >> 4) : ((row & 63) >> 3);       // 16-byte chunk of this key inside a 128-byte row
        const uint32_t pbase = KV8 ? uint32_t(row & 15) : uint32_t((row >> 6) * SM::kQAtomBytes + (row & 7) * 2);
#pragma unroll
        for (int c = 0; c < 8; ++c) pa[c] = pbase + uint32_t((ch ^ c) << 4);
      }
      for (int jj = 0; jj < n; ++jj) {
        const int i = it + jj;
        const int n0 = (j0 + jj) * kSwKV;
        mbar_wait(&s_full[i & 1], (i >> 1) & 1);
        tc_fence_after();
        uint32_t sr[16];
        tmem_ld_32x32b_x16(tmem + (i & 1) * 32 + lane_addr, sr);
        [[maybe_unused]] uint32_t ksf_n[4], vsf_c[4];
        if constexpr (MX) {   // in flight during the softmax: K scales of tile i + 2 (same buffer as tile i), V scales of tile i
          if (jj + 2 < n) load_ksf(x, j0 + jj + 2, ksf_n);
          load_vsf(x, j0 + jj, vsf_c);
        }
        tmem_ld_wait();
        float sv[kSwN];
#pragma unroll
        for (int qn = 0; qn < kSwN; ++qn) sv[qn] = __uint_as_float(sr[qn]) * sc_q[qn];
        if (p.causal || n0 + kSwKV > geo.S) {  // edge / causal tiles only (padding columns qn >= R hold q = 0: harmless)
          const bool row_in = (n0 + row) < geo.S;
          const long long kvpos = p.kv_pos0 + n0 + row;
#pragma unroll
          for (int qn = 0; qn < kSwN; ++qn) {
            const bool vis = row_in && (!p.causal || kvpos <= p.q_pos0 + qi_of[qn]);
            sv[qn] = vis ? sv[qn] : sw_ninf();
          }
        }
        // column maxima over the 128 keys: redux within the warp, one smem exchange across the 4 warps
        const uint32_t xc = xch_a + uint32_t(i & 1) * (4 * kSwN * 4);
        {
          float wm[kSwN];
#pragma unroll
          for (int qn = 0; qn < kSwN; ++qn) wm[qn] = sw_ord2f(__reduce_max_sync(0xffffffffu, sw_f2ord(sv[qn])));
          if (lane == 0) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4)
              st_shared_v4f(xc + warp * (kSwN * 4) + g4 * 16, wm[4 * g4], wm[4 * g4 + 1], wm[4 * g4 + 2], wm[4 * g4 + 3]);
          }
        }
        named_bar_sync(2, kSmx);
        float m4[kSwN];
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          float4 a0 = ld_shared_v4f(xc + g4 * 16), a1 = ld_shared_v4f(xc + kSwN * 4 + g4 * 16);
          float4 a2 = ld_shared_v4f(xc + 2 * kSwN * 4 + g4 * 16), a3 = ld_shared_v4f(xc + 3 * kSwN * 4 + g4 * 16);
          m4[4 * g4 + 0] = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x));
          m4[4 * g4 + 1] = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
          m4[4 * g4 + 2] = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z));
          m4[4 * g4 + 3] = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
        }
        float grow = sw_ninf();   // max over queries of (new maximum - used maximum); NaN (-inf - -inf) drops out of fmaxf
#pragma unroll
        for (int qn = 0; qn < kSwN; ++qn) grow = fmaxf(grow, m4[qn] - m_used[qn]);
        if (grow > kSwRescale) {  // uniform across the CTA: every thread sees the same maxima
          float alpha[kSwN];
#pragma unroll
          for (int qn = 0; qn < kSwN; ++qn) {
            const float m_new = fmaxf(m_used[qn], m4[qn]);
            alpha[qn] = (m_new == sw_ninf()) ? 1.f : fast_exp2(m_used[qn] - m_new);
            l_thr[qn] *= alpha[qn];
            m_used[qn] = m_new;
            neg_ms[qn] = (m_new == sw_ninf()) ? 0.f : -m_new;
          }
          if (jj > 0) {
            mbar_wait(&pv_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
            tc_fence_after();
            uint32_t orow[16];
            tmem_ld_32x32b_x16(tmem_o + lane_addr, orow);
            tmem_ld_wait();
#pragma unroll
            for (int qn = 0; qn < kSwN; ++qn) orow[qn] = __float_as_uint(__uint_as_float(orow[qn]) * alpha[qn]);
            tmem_st_32x32b_x16(tmem_o + lane_addr, orow);
            tmem_st_wait();
          }
        }
        // the P^T buffer of tile i was last read by PV(i-2)
        if (jj >= 2) { mbar_wait(&pv_done[i & 1], ((i - 2) >> 1) & 1); }
        const uint32_t pt = pt_a + uint32_t(i & 1) * SM::kPBytes;
#pragma unroll
        for (int qn = 0; qn < kSwN; ++qn) {
          const float pv = fast_exp2(sv[qn] + neg_ms[qn]);
          l_thr[qn] += pv;
          if constexpr (!KV8) {
            st_shared_u16(pt + pa[qn & 7] + qn * 128, sw_to16<BF16>(pv));
          } else {
            uint16_t e2;
            asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(e2) : "f"(0.f), "f"(pv));
            st_shared_u8(pt + pa[qn & 7] + qn * 128, e2);
          }
        }
        if constexpr (MX) {
          // QK(i) has completed (s_full) and PV(i - 2) too (pv_done above): both scale buffers of parity i & 1 are free
          if (jj + 2 < n) tmem_st_32x32b_x4(tmem + lane_addr + kSfK + 4 * (i & 1), ksf_n[0], ksf_n[1], ksf_n[2], ksf_n[3]);
          tmem_st_32x32b_x4(tmem + lane_addr + kSfV + 4 * (i & 1), vsf_c[0], vsf_c[1], vsf_c[2], vsf_c[3]);
          tmem_st_wait();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[i & 1]);
      }
      // ---- segment epilogue: O^T (thread = channel d, 16 query columns), row sums reduced over the 128 keys
      const dcomm::Tail tail = make_tail();
      int nparts, pidx;
      dcomm::head_parts(geo, x, cta, nparts, pidx);
      uint64_t* my_part = dcomm::part_ptr<D>(tail, x, pidx);   // tagged words: no fence, no ticket (decode_comm.cuh)
      {
        float lw[kSwN];
#pragma unroll
        for (int qn = 0; qn < kSwN; ++qn) {
          float v = l_thr[qn];
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
          lw[qn] = v;
        }
        if (lane == 0) {
#pragma unroll
          for (int qn = 0; qn < kSwN; ++qn) red_s[warp * kSwN + qn] = lw[qn];
        }
        named_bar_sync(2, kSmx);
        uint32_t orow[16];
        if (n > 0) {
          const int il = it + n - 1;
          mbar_wait(&pv_done[il & 1], (il >> 1) & 1);
          tc_fence_after();
          tmem_ld_32x32b_x16(tmem_o + lane_addr, orow);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int qn = 0; qn < kSwN; ++qn) orow[qn] = 0u;
        }
        float vs_d = 1.f;
        if constexpr (KV8 && !MX) vs_d = ch_scale[D + row];
#pragma unroll
        for (int qn = 0; qn < kSwN; ++qn)
          if (qn < R) dcomm::ll_store_gpu(my_part + qn * (D + 2) + row, __uint_as_float(orow[qn]) * vs_d, tail.wtag);   // thread = channel d
        if (tid < R) {
          const float lsum = red_s[tid] + red_s[kSwN + tid] + red_s[2 * kSwN + tid] + red_s[3 * kSwN + tid];
          float mq = sw_ninf();
#pragma unroll
          for (int qn = 0; qn < kSwN; ++qn) mq = (qn == tid) ? m_used[qn] : mq;
          dcomm::ll_store_gpu(my_part + tid * (D + 2) + D, mq, tail.wtag);
          dcomm::ll_store_gpu(my_part + tid * (D + 2) + D + 1, lsum, tail.wtag);
        }
        tc_fence_before();
      }
      it += n;
      dcomm::segment_done<D, kSmx, 4>(tail, x, tn, n_pend, tid, 1, store_out);   // the owner of the head's last tile queues the merge
      named_bar_sync(1, kSmx);  // Q / P^T smem and s_misc reuse by the next segment
      t = tn;
    }
    const dcomm::Tail tail = make_tail();
    dcomm::drain<D, kSmx, 4>(tail, n_pend, tid, 1, store_out);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc<128>(tmem); }
}

template <bool BF16, bool KV8, bool MX = false>
std::function<void(cudaStream_t)> make_sw_pass(const CUtensorMap& kmap, const CUtensorMap& vmap, const DecodeSwParams& p, int grid) {
  auto kern = decode_swap_kernel<BF16, KV8, MX>;
  static bool configured = false;
  if (!configured) {
    TA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SwSmem<KV8>::kTotal));
    configured = true;
  }
  return [kern, kmap, vmap, p, grid](cudaStream_t stream) {
    kern<<<grid, kSwThreads, SwSmem<KV8>::kTotal, stream>>>(kmap, vmap, p);
    TA_CUDA_CHECK(cudaGetLastError());
  };
}

}  // namespace

PreparedLaunch decode_swap_prepare(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                                   float* part, uint32_t* tickets, const CommCtxHost& comm, int nsm, const float* kscale,
                                   const float* vscale, const uint32_t* k_sf, const uint32_t* v_sf, const int* kv_len) {
  const bool mx = k_sf != nullptr;
  const bool kv8 = kscale != nullptr || mx;
  if (s.D != 128) throw std::runtime_error("decode_swap: head_dim must be 128");
  if (s.Hq % s.Hkv != 0) throw std::runtime_error("decode_swap: Hq must be a multiple of Hkv");
  const int G = s.Hq / s.Hkv;
  const int R = G * s.Sq;
  if (R > kSwN) throw std::runtime_error("decode_swap: (Hq / Hkv) * Sq must be <= 16");
  if (s.S <= 0) throw std::runtime_error("decode_swap: the KV shard must have capacity for at least one row");
  int grid, max_parts;
  // one persistent CTA per SM (the driver keeps tcgen05 kernels at one resident CTA per SM: a second CTA only queues)
  decode_split(s.B * s.Hkv, s.S, nsm, &grid, &max_parts);
  if ((long long)s.B * s.Hkv / grid + 2 > kSwMaxPending)
    throw std::runtime_error("decode_swap: batch x kv-heads too large for one launch (split the batch)");
  if (comm.world > 1) {
    const size_t need = (size_t)2 * comm.world * s.B * s.Hkv * R * (s.D + 2) * 8;
    if (need > comm.data_bytes) throw std::runtime_error("decode_swap: symmetric buffer too small for this problem");
  }
  const int eb = kv8 ? 1 : 2;
  CUtensorMap kmap = make_tmap_bhsd(k, eb, s.B, s.Hkv, s.S, s.D, s.k_sb, s.k_sh, s.k_ss, 128 / eb, kSwKV, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap vmap = make_tmap_bhsd(v, eb, s.B, s.Hkv, s.S, s.D, s.v_sb, s.v_sh, s.v_ss, 128 / eb, kSwKV, CU_TENSOR_MAP_SWIZZLE_128B);
  DecodeSwParams p;
  p.kscale = kscale; p.vscale = vscale; p.k_sf = k_sf; p.v_sf = v_sf; p.kv_len = kv_len;
  p.q = q; p.out = out; p.lse = lse;
  p.part = reinterpret_cast<uint64_t*>(part); p.wctr = reinterpret_cast<unsigned long long*>(tickets);
  p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.G = G; p.Sq = s.Sq; p.S = s.S; p.R = R;
  p.scale_log2 = s.softmax_scale * 1.4426950408889634f;
  p.causal = s.causal; p.q_pos0 = s.q_pos0; p.kv_pos0 = s.kv_pos0;
  p.q_sb = s.q_sb; p.q_sh = s.q_sh; p.q_ss = s.q_ss; p.o_sb = s.o_sb; p.o_sh = s.o_sh; p.o_ss = s.o_ss;
  p.tph_cap = (s.S + kSwKV - 1) / kSwKV;
  p.max_parts = max_parts;
  p.comm = dcomm::to_device_ctx(comm);
  PreparedLaunch pl;
  if (mx) pl.passes.push_back(s.is_bf16 ? make_sw_pass<true, true, true>(kmap, vmap, p, grid) : make_sw_pass<false, true, true>(kmap, vmap, p, grid));
  else if (kv8) pl.passes.push_back(s.is_bf16 ? make_sw_pass<true, true>(kmap, vmap, p, grid) : make_sw_pass<false, true>(kmap, vmap, p, grid));
  else pl.passes.push_back(s.is_bf16 ? make_sw_pass<true, false>(kmap, vmap, p, grid) : make_sw_pass<false, false>(kmap, vmap, p, grid));
  return pl;
}

void decode_swap_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse, float* part,
                        uint32_t* tickets, const CommCtxHost& comm, int nsm, cudaStream_t stream, const float* kscale,
                        const float* vscale, const uint32_t* k_sf, const uint32_t* v_sf, const int* kv_len) {
  decode_swap_prepare(s, q, k, v, out, lse, part, tickets, comm, nsm, kscale, vscale, k_sf, v_sf, kv_len).run(stream);
}

}  // namespace ta
