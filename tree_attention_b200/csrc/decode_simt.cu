// attn_tree_fused_decode -- split-KV streaming decode attention for sm_100a with the whole
// tree combine in its epilogue.
//
// Replaces, in ONE launch per rank, everything the reference runs for a decode step
// (/root/reference/model.py:74-80 local attention, :103-124 combine; SURVEY.md 2.3 rows K1-K13 and
// the three NCCL all-reduces N1-N3):
//
//   1. persistent CTAs stream this rank's K/V shard with TMA (cp.async.bulk.tensor, 128B swizzle)
//      through a 3-6 stage mbarrier ring; work is stream-K balanced over (batch, kv-head, tile);
//   2. online-softmax partials per warp -> per CTA -> (atomic ticket) per (batch, kv-head);
//   3. the last-arriving CTA of a head publishes the normalised partial (o, lse) with plain P2P
//      stores into every peer's symmetric buffer, then releases a per-(head, source) epoch flag with
//      st.release.sys;  after finishing its remaining tiles it acquires the peers' flags and merges
//      the W partials in fixed rank order (bitwise identical output on every rank).
//
// The math is CUDA-core (decode is a GEMV per head: ~1 FMA per byte of KV, HBM bound).  Query rows
// R = (GQA group) x Sq <= 4 per pass; larger Sq goes to the tcgen05 kernel (attn_fwd_sm100.cu).
#include "common.cuh"
#include "decode_comm.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {

namespace {

constexpr int kTileRows = 128;
constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int kRowsPerWarp = kTileRows / kConsumerWarps;  // 16
constexpr int kMaxPending = 64;

struct DecodeParams {
  const void* q;
  void* out;
  float* lse;
  const uint32_t* kscale;  // KV8 only: (B, Hkv, S) words of 4 UE8M0 scales (one per 32 elements of D = 128)
  const uint32_t* vscale;
  uint64_t* part;             // workspace: tagged partial words (decode_comm.cuh)
  unsigned long long* wctr;   // workspace arrival counter (launch tag of the partial words), zero-initialised once
  const int* kv_len;  // optional device scalar: valid rows of this shard (<= S); read by the kernel => graph-replayable
  int B, Hq, Hkv, G, Sq, S;   // S = capacity of the shard (rows covered by the tensor maps)
  int rows_valid;     // G * Sq - r_base, clipped to R
  int r_base;
  float scale_log2;
  int causal;
  long long q_pos0, kv_pos0;
  long long q_sb, q_sh, q_ss, o_sb, o_sh, o_ss;
  int max_parts;      // bound on the CTAs sharing one head, for ANY number of valid rows (decode_simt_plan)
  unsigned long long* trace;  // optional [grid][16] timeline stamps (bench_tools/trace_decode.py); null in production
  int pdl;  // 0: plain launch; 1: programmatic dependent launch; 2: + K/V prefetch before the dependency wait (static KV)
  CommCtx comm;
};

// KV8: block-scaled fp8 (e4m3 + UE8M0 per 32) K/V: rows are D bytes = one 128B-swizzle atom for D = 128
template <int D, bool KV8>
struct SmemLayout {
  static constexpr int kElem = KV8 ? 1 : 2;
  static constexpr int kAtoms = D * kElem / 128;
  static constexpr int kTensorBytes = kTileRows * D * kElem;  // one of K or V
  static constexpr int kStageBytes = 2 * kTensorBytes;
  static constexpr int kStages = (D * kElem == 256) ? 3 : 6;
  static constexpr int kPsPerWarp = KV8 ? 4 : 1;  // P x V-scale is kept per 32-wide block of D
};

template <int D, int R, bool KV8>
constexpr size_t smem_bytes() {
  using L = SmemLayout<D, KV8>;
  return 1024 /*align slack*/ + size_t(L::kStages) * L::kStageBytes +
         sizeof(float) * (2 * R * D + kConsumerWarps * L::kPsPerWarp * R * kRowsPerWarp +
                          (R <= 2 ? 2 : 1) * kConsumerWarps * R * (D + 4)) +
         sizeof(int) * (kMaxPending + 8) + sizeof(uint64_t) * (2 * L::kStages + 2);
}

template <bool BF16>
__device__ __forceinline__ void cvt8(const uint4& w, float (&f)[8]) {
  if constexpr (BF16) {
    f[0] = bf16lo(w.x); f[1] = bf16hi(w.x); f[2] = bf16lo(w.y); f[3] = bf16hi(w.y);
    f[4] = bf16lo(w.z); f[5] = bf16hi(w.z); f[6] = bf16lo(w.w); f[7] = bf16hi(w.w);
  } else {
    f[0] = f16lo(w.x); f[1] = f16hi(w.x); f[2] = f16lo(w.y); f[3] = f16hi(w.y);
    f[4] = f16lo(w.z); f[5] = f16hi(w.z); f[6] = f16lo(w.w); f[7] = f16hi(w.w);
  }
}
template <bool BF16>
__device__ __forceinline__ float cvt1(uint16_t h) {
  if constexpr (BF16) return __uint_as_float(uint32_t(h) << 16);
  else return __half2float(__ushort_as_half(h));
}
template <bool BF16>
__device__ __forceinline__ uint16_t to16(float f) {
  if constexpr (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(f));
  else return __half_as_ushort(__float2half_rn(f));
}

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

template <int D, int R, bool BF16, bool KV8>
__global__ void __launch_bounds__(kThreads, 1)
decode_simt_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                   const __grid_constant__ DecodeParams p) {
  using L = SmemLayout<D, KV8>;
  constexpr int NS = L::kStages;
  constexpr int EPL = D / 32;  // output elements per lane in the PV phase
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* q_s = reinterpret_cast<float*>(smem + size_t(NS) * L::kStageBytes);  // [2][R][D], pre-scaled: this CTA's first two heads
  float* p_s = q_s + 2 * R * D;                                                // [warps][(blk)][R][16]
  float* merge_s = p_s + kConsumerWarps * L::kPsPerWarp * R * kRowsPerWarp;    // [2][warps][R][D+4] (double-buffered by segment)
  constexpr int kMergeBufs = R <= 2 ? 2 : 1;   // R = 4: no room for a second buffer, a trailing barrier instead
  int* pending = reinterpret_cast<int*>(merge_s + kMergeBufs * kConsumerWarps * R * (D + 4));
  int* s_misc = pending + kMaxPending;  // [0]=n_pending, [1]=scratch, [4]=intra-GPU tag, [5]=cross-GPU tag, [6]=tags ready
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_misc + 8);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* stamps = empty_bar + NS;   // [0] CTA start, [1] last publish (globaltimer, thread 0)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int world = p.comm.world;
  // timeline stamps: [0] entry (globaltimer) [1] entry (clock64) [2] prologue done [3] dependency resolved [4] first tile
  // landed [5] last tile consumed [6] last partial stored [7] merges published [8] end (clock64) [9] end (globaltimer)
  // [10] producer: first TMA issued [11] producer: last TMA issued [12] tiles of this CTA
  unsigned long long* trc = p.trace ? p.trace + (size_t)cta * 16 : nullptr;
  if (trc && tid == 0) {
    trc[0] = globaltimer_ns(); trc[1] = clock64();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trc[13] = smid; trc[14] = 0; trc[15] = 0;   // [13] SM id [14] cycles inside mid-stream finalize_segment calls [15] their count
  }

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kConsumerWarps);
    }
    fence_mbar_init();
    s_misc[0] = 0; s_misc[1] = 0; s_misc[6] = 0;
  }
  if (warp == kConsumerWarps && lane == 0) {
    tma_prefetch_desc(&kmap);
    tma_prefetch_desc(&vmap);
  }
  __syncthreads();
  if (trc && tid == 0) trc[2] = clock64();

  // Programmatic dependent launch: let the NEXT launch of the stream start its prologue (and, for a static KV
  // cache, its first K/V tile loads) on SMs this grid has already vacated, while our last CTAs still wait for
  // their peers.  Everything that depends on earlier kernels (q, workspace, tickets, epoch) is touched only after
  // griddepcontrol.wait.
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int BH = p.B * p.Hkv;
  const long long q_pos_max = p.q_pos0 + p.Sq - 1;

  auto tile_visible = [&](int j) -> bool {
    return !p.causal || (p.kv_pos0 + (long long)j * kTileRows <= q_pos_max);
  };

  if (warp == kConsumerWarps) {
    // ------------------------------- TMA producer --------------------------------------------
    if (lane == 0) {
      // pdl == 2 streams K/V (and reads kv_len) BEFORE the dependency wait: only valid when neither was written by
      // the kernel this launch depends on (the session downgrades to pdl = 1 for the step after an append)
      if (p.pdl == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
      const dcomm::Geom geo = dcomm::make_geom(p.S, p.kv_len, BH, gridDim.x);
      const int t_lo = dcomm::cta_lo(geo, cta), t_hi = dcomm::cta_lo(geo, cta + 1);
      int stage = 0, issued = 0;
      uint32_t phase = 0;
      bool tags_done = false;
      // launch tags (decode_comm.cuh): the arrival atomics are issued once the ring is full -- the producer would block
      // on the first `empty` barrier anyway, and no consumer can free a stage before the dependency of a programmatic
      // launch has resolved -- so their round trip is never exposed; consumers pick the tags up at their first segment end
      auto fetch_tags = [&]() {
        if (p.pdl == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
        const uint32_t wtag = dcomm::launch_tag(p.wctr);
        const uint32_t ctag = world > 1 ? dcomm::launch_tag(reinterpret_cast<unsigned long long*>(p.comm.epoch)) : 0u;
        volatile int* sm = s_misc;
        sm[4] = (int)wtag; sm[5] = (int)ctag;
        __threadfence_block();
        sm[6] = 1;
        tags_done = true;
      };
      for (int t = t_lo; t < t_hi; ++t) {
        const int x = t / geo.tph, j = t - x * geo.tph;
        if (!tile_visible(j)) continue;
        const int b = x / p.Hkv, h = x - b * p.Hkv;
        if (issued == NS && !tags_done) fetch_tags();
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (trc && issued == 0) trc[10] = clock64();
        mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        uint8_t* ks = stage_base + size_t(stage) * L::kStageBytes;
        uint8_t* vs = ks + L::kTensorBytes;
#pragma unroll
        for (int a = 0; a < L::kAtoms; ++a) {
          tma_load_4d(ks + a * (kTileRows * 128), &kmap, &full_bar[stage], a * (128 / L::kElem), j * kTileRows, h, b);
          tma_load_4d(vs + a * (kTileRows * 128), &vmap, &full_bar[stage], a * (128 / L::kElem), j * kTileRows, h, b);
        }
        ++issued;
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
      if (!tags_done) fetch_tags();
      if (trc) { trc[11] = clock64(); trc[12] = (unsigned long long)issued; }
    }
    return;
  }

  // --------------------------------- consumers ----------------------------------------------
  if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (trc && tid == 0) trc[3] = clock64();
  const dcomm::Geom geo = dcomm::make_geom(p.S, p.kv_len, BH, gridDim.x);
  const int t_lo = dcomm::cta_lo(geo, cta), t_hi = dcomm::cta_lo(geo, cta + 1);
  constexpr int RB = R;  // rows of one output channel gathered per batch of the cross-GPU merge (R <= 4)
  if (tid == 0) { stamps[0] = globaltimer_ns(); stamps[1] = 0; }
  // built on demand (kept out of the streaming loop's live registers)
  auto make_tail = [&]() {
    dcomm::Tail tl;
    volatile int* sm = s_misc;
    while (sm[6] == 0) { }   // launch tags fetched by the producer warp (long done by the first segment end)
    tl.comm = &p.comm; tl.part = p.part; tl.max_parts = p.max_parts; tl.BH = BH;
    tl.R = R; tl.rows_valid = p.rows_valid; tl.wtag = (uint32_t)sm[4]; tl.ctag = (uint32_t)sm[5]; tl.geo = geo;
    tl.pending = pending; tl.max_pending = kMaxPending; tl.stamps = stamps;
    return tl;
  };
  const int r16 = lane & 15;
  const int half = lane >> 4;
  float m_run[R], l_run[R], o_acc[R][EPL];
  int cur_x = -1;
  int stage = 0;
  uint32_t phase = 0;

  auto reset_state = [&]() {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      m_run[r] = neg_inf();
      l_run[r] = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) o_acc[r][e] = 0.f;
    }
  };

  // The query rows of a head, pre-scaled, into buffer `buf` of q_s.  A head switch in the middle of the stream must
  // not stall the consumers: the stage ring holds exactly the bytes that keep this SM's share of HBM busy (Little's
  // law: 192 KB in flight ~ 4 us of loaded latency), so a 2.5 us stall (finalize + a global load of q + 3 barriers)
  // drains it and costs the CTA 5-10 us -- and every CTA with a head boundary in its range became the kernel's
  // straggler (bench_tools/trace_decode.py, profiles/r2_decode/).  Hence: the first TWO heads of the CTA's range are
  // loaded up front (one latency, overlapped with the first TMA loads); only a third head (short sequences) loads late.
  auto load_q = [&](int x, int buf) {
    const int b = x / p.Hkv, h = x - b * p.Hkv;
    for (int idx = tid; idx < R * D; idx += kConsumerThreads) {
      const int r = idx / D, d = idx - r * D;
      float val = 0.f;
      if (r < p.rows_valid) {
        const int rr = p.r_base + r;
        const int g = rr / p.Sq, i = rr - g * p.Sq;
        const uint16_t* qp = reinterpret_cast<const uint16_t*>(p.q) + (long long)b * p.q_sb +
                             (long long)(h * p.G + g) * p.q_sh + (long long)i * p.q_ss + d;
        val = cvt1<BF16>(*qp) * p.scale_log2;
      }
      if constexpr (KV8) reinterpret_cast<__half*>(q_s + buf * R * D)[idx] = __float2half_rn(val);
      else q_s[buf * R * D + idx] = val;
    }
  };
  const int x_first = t_lo / geo.tph;
  if (t_lo < t_hi) {
    load_q(x_first, 0);
    if ((x_first + 1) * geo.tph < t_hi) load_q(x_first + 1, 1);
  }
  named_bar_sync(1, kConsumerThreads);
  const float* q_cur = q_s;
  uint32_t wtag = 0;   // launch tag of the partial words (picked up from the producer thread at the first segment end)
  int n_fin = 0;       // segments finalised so far (selects the merge_s buffer)
  int n_pend = 0;      // heads queued for the merge in drain (uniform across the threads)

  // write a finished (b, kv-head) result (normalised o, log2-domain lse) to the user tensors
  auto store_out = [&](int x, int r, int d, float o_norm, float lse2) {
    const int b = x / p.Hkv, h = x - b * p.Hkv;
    const int rr = p.r_base + r;
    const int g = rr / p.Sq, i = rr - g * p.Sq;
    uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)(h * p.G + g) * p.o_sh +
                   (long long)i * p.o_ss + d;
    *op = to16<BF16>(o_norm);
    if (d == 0 && p.lse != nullptr)
      p.lse[((long long)b * p.Hq + (h * p.G + g)) * p.Sq + i] = lse2 * 0.6931471805599453f;
  };

  // per-CTA partial of head x is complete: merge warps, write the partial, then the shared tail (decode_comm.cuh):
  // ticket -> last CTA merges the parts -> output (world == 1) or tagged 8-byte words to every rank + deferred merge
  auto finalize_segment = [&](int x) {
    // (a) warps -> smem (double-buffered by segment: a fast warp may start the NEXT segment's (a) while slow threads of
    // this one are still in (b); two finalizes are at least one tile apart, (b) lasts a fraction of a tile)
    float* mbuf = merge_s + (kMergeBufs == 2 ? (n_fin & 1) : 0) * (kConsumerWarps * R * (D + 4));
    ++n_fin;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float lt = l_run[r];
      lt += __shfl_xor_sync(0xffffffffu, lt, 8);
      lt += __shfl_xor_sync(0xffffffffu, lt, 4);
      lt += __shfl_xor_sync(0xffffffffu, lt, 2);
      lt += __shfl_xor_sync(0xffffffffu, lt, 1);
      float* ms = mbuf + (warp * R + r) * (D + 4);
#pragma unroll
      for (int e = 0; e < EPL; ++e) ms[lane * EPL + e] = o_acc[r][e];
      if (lane == 0) { ms[D] = m_run[r]; ms[D + 1] = lt; }
    }
    named_bar_sync(1, kConsumerThreads);   // the ONLY barrier of a head switch
    // (b) CTA partial -> workspace as tagged words (no fence, no ticket: every word validates itself).  Nothing here may
    // call out of line or touch local memory: with 227 KB of shared memory there is no L1 left for spills, and every
    // microsecond the consumers spend here drains the stage ring (see load_q above)
    if (wtag == 0) {
      volatile int* sm = s_misc;
      while (sm[6] == 0) { }   // launch tags fetched by the producer thread (long done by the first segment end)
      wtag = (uint32_t)sm[4];
    }
    int nparts, pidx;
    dcomm::head_parts(geo, x, cta, nparts, pidx);
    uint64_t* my_part = p.part + ((size_t)x * p.max_parts + pidx) * (size_t)(R * (D + 2));
    for (int idx = tid; idx < R * D; idx += kConsumerThreads) {
      const int r = idx / D, d = idx - r * D;
      float M = neg_inf();
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) M = fmaxf(M, mbuf[(w * R + r) * (D + 4) + D]);
      const float Ms = (M == neg_inf()) ? 0.f : M;
      float acc = 0.f, Lsum = 0.f;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        const float* ms = mbuf + (w * R + r) * (D + 4);
        const float sc = fast_exp2(ms[D] - Ms);
        acc = fmaf(ms[d], sc, acc);
        Lsum = fmaf(ms[D + 1], sc, Lsum);
      }
      dcomm::ll_store_gpu(my_part + r * (D + 2) + d, acc, wtag);
      if (d == 0) {
        dcomm::ll_store_gpu(my_part + r * (D + 2) + D, M, wtag);
        dcomm::ll_store_gpu(my_part + r * (D + 2) + D + 1, Lsum, wtag);
      }
    }
    // (c) the owner of the head's last tile merges it -- after this CTA's own last tile (drain below).  The list cannot
    // overflow: decode_simt_prepare bounds the heads per CTA by kMaxPending
    if (min(t_hi, (x + 1) * geo.tph) == (x + 1) * geo.tph) {
      if (tid == 0) pending[n_pend] = x;
      ++n_pend;
    }
    if constexpr (kMergeBufs == 1) named_bar_sync(1, kConsumerThreads);   // single merge buffer: (b) must finish before the next (a)
  };

  // KV8: the block scales of the NEXT tile are prefetched into registers one iteration ahead, so that their
  // global-load latency hides behind the current tile's math and the mbarrier wait.
  [[maybe_unused]] uint32_t nxt_ksc = 0, nxt_vsc = 0;
  [[maybe_unused]] auto fetch_scales = [&](int tt) {
    if constexpr (KV8) {
      if (tt < t_hi) {
        const int xx = tt / geo.tph, jj = tt - xx * geo.tph;
        const long long srow = (long long)xx * p.S + min((long long)jj * kTileRows + warp * kRowsPerWarp + (lane & 15), (long long)p.S - 1);
        nxt_ksc = __ldg(p.kscale + srow);
        nxt_vsc = __ldg(p.vscale + srow);
      }
    }
  };
  fetch_scales(t_lo);

  // bf16 / fp16 hot loop: explicit shared-state-space accesses with 32-bit addresses and per-lane offsets computed ONCE.
  // (Generic pointers carved out of the dynamic smem block compile to LD.E with 64-bit address arithmetic per access: the
  // round-2 SASS of this loop spent as many IADD3/IADD3.X/LOP3 on addresses as FFMAs on the math -- and the kernel is
  // issue / power sensitive: 22 % fewer instructions bought 7 % more HBM bandwidth.)
  const uint32_t stage_a = smem_u32(stage_base);
  const uint32_t ps_a = smem_u32(p_s);
  [[maybe_unused]] uint32_t k_off[D / 16];     // QK: this lane's D/16 16-byte chunks of its key row (128B swizzle applied)
  [[maybe_unused]] uint32_t v_sw[8];           // PV: swizzled chunk offset of this lane's columns for row & 7 = 0..7
  [[maybe_unused]] uint32_t q_off[D / 16];     // QK: byte offset of the matching 8 query values
  if constexpr (!KV8) {
    const int row = warp * kRowsPerWarp + r16;
#pragma unroll
    for (int c = 0; c < D / 16; ++c) {
      const int cg = half * (D / 16) + c;
      const int atom = cg >> 3, cia = cg & 7;
      k_off[c] = uint32_t(atom * (kTileRows * 128) + row * 128 + ((cia ^ (row & 7)) << 4));
      q_off[c] = uint32_t(cg * 8 * 4);
    }
    const int c0 = lane * EPL;
    const int atom = (c0 * 2) >> 7, inner = (c0 * 2) & 127;
    const int chunk = inner >> 4, off = inner & 15;
#pragma unroll
    for (int k7 = 0; k7 < 8; ++k7)
      v_sw[k7] = uint32_t(L::kTensorBytes + atom * (kTileRows * 128) + off + ((chunk ^ k7) << 4) + (warp * kRowsPerWarp) * 128);
  }

  for (int t = t_lo; t < t_hi; ++t) {
    const int x = t / geo.tph, j = t - x * geo.tph;
    [[maybe_unused]] const uint32_t ksc_word = nxt_ksc;
    [[maybe_unused]] const uint32_t vsc_cur = nxt_vsc;
    fetch_scales(t + 1);
    bool switched = false;
    long long ts0 = 0, ts1 = 0, ts2 = 0;
    if (x != cur_x) {
      switched = cur_x >= 0;
      ts0 = clock64();
      if (cur_x >= 0) finalize_segment(cur_x);
      ts1 = clock64();
      const int rel = x - x_first;
      if (rel >= 2) {   // third+ head of this CTA's range (short sequences): its buffer was last used two heads ago
        load_q(x, rel & 1);
        named_bar_sync(1, kConsumerThreads);
      }
      q_cur = q_s + (rel & 1) * R * D;
      reset_state();
      cur_x = x;
      ts2 = clock64();
    }
    if (!tile_visible(j)) continue;
    mbar_wait(&full_bar[stage], phase);
    if (trc && tid == 0 && t == t_lo) trc[4] = clock64();
    if (trc && tid == 0 && switched) {   // mid-stream head switch: [14] finalize cycles | load_q cycles << 32, [15] wait for the next tile | t - t_lo << 32
      trc[14] = (unsigned long long)(ts1 - ts0) | ((unsigned long long)(ts2 - ts1) << 32);
      trc[15] = (unsigned long long)(clock64() - ts2) | ((unsigned long long)(t - t_lo) << 32);
      trc[7] = (unsigned long long)ts0;
    }
    const uint8_t* ks = stage_base + size_t(stage) * L::kStageBytes;
    const uint8_t* vs = ks + L::kTensorBytes;

    // ---------------- S = q . K^T : lane = (row r16 of this warp's 16 rows, half of D) -------------
    const int row = warp * kRowsPerWarp + r16;
    float s_sum[R];
    [[maybe_unused]] uint32_t vsc_word = 0;
    if constexpr (!KV8) {
      const uint32_t ks_a = stage_a + uint32_t(stage) * L::kStageBytes;
      const uint32_t q_a = smem_u32(q_cur);
      uint64_t acc2[R][2];   // packed f32x2 accumulators: fma.rn.f32x2 issues ONE instruction for two FMAs on sm_100
#pragma unroll
      for (int r = 0; r < R; ++r) { acc2[r][0] = 0ull; acc2[r][1] = 0ull; }
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        const uint4 kw = ld_shared_v4u(ks_a + k_off[c]);
        uint64_t k2[4];
        if constexpr (BF16) {
          k2[0] = pack_f32x2(bf16lo(kw.x), bf16hi(kw.x)); k2[1] = pack_f32x2(bf16lo(kw.y), bf16hi(kw.y));
          k2[2] = pack_f32x2(bf16lo(kw.z), bf16hi(kw.z)); k2[3] = pack_f32x2(bf16lo(kw.w), bf16hi(kw.w));
        } else {
          k2[0] = pack_f32x2(f16lo(kw.x), f16hi(kw.x)); k2[1] = pack_f32x2(f16lo(kw.y), f16hi(kw.y));
          k2[2] = pack_f32x2(f16lo(kw.z), f16hi(kw.z)); k2[3] = pack_f32x2(f16lo(kw.w), f16hi(kw.w));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          uint64_t q01, q23, q45, q67;
          ld_shared_v2u64(q_a + uint32_t(r * D * 4) + q_off[c], q01, q23);
          ld_shared_v2u64(q_a + uint32_t(r * D * 4) + q_off[c] + 16, q45, q67);
          acc2[r][0] = fma2_f32x2(k2[0], q01, acc2[r][0]);
          acc2[r][1] = fma2_f32x2(k2[1], q23, acc2[r][1]);
          acc2[r][0] = fma2_f32x2(k2[2], q45, acc2[r][0]);
          acc2[r][1] = fma2_f32x2(k2[3], q67, acc2[r][1]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float a0, a1, a2, a3;
        unpack_f32x2(acc2[r][0], a0, a1);
        unpack_f32x2(acc2[r][1], a2, a3);
        s_sum[r] = (a0 + a1) + (a2 + a3);
      }
    } else {
      // block-scaled fp8: this lane covers 64 elements = 4 chunks of 16 = blocks {2*half, 2*half+1}; products are
      // accumulated per chunk in fp16x2 (8 HFMA2), then scaled by the block's UE8M0 scale in fp32.
      vsc_word = vsc_cur;
#pragma unroll
      for (int r = 0; r < R; ++r) s_sum[r] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cg = half * 4 + c;
        const uint4 kw = *reinterpret_cast<const uint4*>(ks + row * 128 + ((cg ^ (row & 7)) << 4));
        uint32_t kh[8];
        const uint32_t kw4[4] = {kw.x, kw.y, kw.z, kw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(kh[2 * i]) : "h"((uint16_t)(kw4[i] & 0xffffu)));
          asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(kh[2 * i + 1]) : "h"((uint16_t)(kw4[i] >> 16)));
        }
        const float ksc = __uint_as_float(((ksc_word >> (8 * (cg >> 1))) & 0xffu) << 23);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint4 qa = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(q_cur) + r * D + cg * 16);
          const uint4 qb = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(q_cur) + r * D + cg * 16 + 8);
          const uint32_t qh[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
          __half2 a0 = __float2half2_rn(0.f), a1 = a0;
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            a0 = __hfma2(*reinterpret_cast<const __half2*>(&kh[i]), *reinterpret_cast<const __half2*>(&qh[i]), a0);
            a1 = __hfma2(*reinterpret_cast<const __half2*>(&kh[i + 1]), *reinterpret_cast<const __half2*>(&qh[i + 1]), a1);
          }
          const float2 f0 = __half22float2(a0), f1 = __half22float2(a1);
          s_sum[r] = fmaf((f0.x + f0.y) + (f1.x + f1.y), ksc, s_sum[r]);
        }
      }
    }
    const long long grow = (long long)j * kTileRows + row;
    const bool inb = grow < geo.S;
    const long long kvpos = p.kv_pos0 + grow;
    float alpha[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = s_sum[r];
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      const int i = (p.r_base + r) % p.Sq;
      const bool vis = inb && (!p.causal || kvpos <= p.q_pos0 + i);
      s = vis ? s : neg_inf();
      float mx = s;
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      const float m_new = fmaxf(m_run[r], mx);
      const float m_safe = (m_new == neg_inf()) ? 0.f : m_new;
      alpha[r] = fast_exp2(m_run[r] - m_safe);
      const float pv = fast_exp2(s - m_safe);
      l_run[r] = fmaf(l_run[r], alpha[r], pv);
      m_run[r] = m_new;
      if constexpr (!KV8) {
        if (half == 0) st_shared_f32(ps_a + uint32_t(((warp * R + r) * kRowsPerWarp + r16) * 4), pv);
      } else {
        // P x (V block scale), one copy per 32-wide block of D; this lane owns blocks 2*half and 2*half+1
        const int b0 = 2 * half;
        p_s[((warp * 4 + b0) * R + r) * kRowsPerWarp + r16] = pv * __uint_as_float(((vsc_word >> (8 * b0)) & 0xffu) << 23);
        p_s[((warp * 4 + b0 + 1) * R + r) * kRowsPerWarp + r16] = pv * __uint_as_float(((vsc_word >> (8 * b0 + 8)) & 0xffu) << 23);
      }
#pragma unroll
      for (int e = 0; e < EPL; ++e) o_acc[r][e] *= alpha[r];
    }
    __syncwarp();

    // ---------------- O += P . V : lane = EPL consecutive output columns ---------------------------
    // Rows past the valid length may hold anything, and 0 x NaN is NaN: on the one tile per head that contains the end of
    // the valid range, every warp zeroes ITS invalid V rows in the stage buffer first (the hot loop stays branch-free; a
    // per-row guard inside it cost 18 % more issued instructions and 5 % of the single-GPU time in round 2).
    {
      const int nv = geo.S - (j * kTileRows + warp * kRowsPerWarp);   // valid rows of this warp's 16 (>= 16: all)
      if (nv < kRowsPerWarp) {
        constexpr int kRowBytes = D * L::kElem;        // 256 (two 128-byte atoms) or 128
        constexpr int kLanes = kRowBytes / 8;
        uint8_t* vw = const_cast<uint8_t*>(vs);
        for (int rr = max(nv, 0); rr < kRowsPerWarp; ++rr) {
          const int vrow = warp * kRowsPerWarp + rr;
          if (lane < kLanes)
            *reinterpret_cast<uint2*>(vw + (lane >> 4) * (kTileRows * 128) + vrow * 128 + (lane & 15) * 8) = make_uint2(0u, 0u);
        }
        fence_proxy_async_smem();   // generic-proxy writes to a buffer the TMA (async proxy) will overwrite after the release below
        __syncwarp();
      }
    }
    if constexpr (KV8) {
      // fp8: 4 output columns = 4 bytes; the P x scale coefficients of this lane's block come from smem
      const int blk = lane >> 3;
      const int chunk = lane >> 2, off = (lane & 3) * 4;
#pragma unroll
      for (int jj = 0; jj < kRowsPerWarp; jj += 4) {
        float4 pr[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
          pr[r] = *reinterpret_cast<const float4*>(p_s + ((warp * 4 + blk) * R + r) * kRowsPerWarp + jj);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int vrow = warp * kRowsPerWarp + jj + u;
          const uint32_t w = *reinterpret_cast<const uint32_t*>(vs + vrow * 128 + ((chunk ^ (vrow & 7)) << 4) + off);
          uint32_t h0, h1;
          asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h0) : "h"((uint16_t)(w & 0xffffu)));
          asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h1) : "h"((uint16_t)(w >> 16)));
          const float2 v01 = __half22float2(*reinterpret_cast<const __half2*>(&h0));
          const float2 v23 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float pp = (u == 0) ? pr[r].x : (u == 1) ? pr[r].y : (u == 2) ? pr[r].z : pr[r].w;
            o_acc[r][0] = fmaf(pp, v01.x, o_acc[r][0]);
            o_acc[r][1] = fmaf(pp, v01.y, o_acc[r][1]);
            o_acc[r][2] = fmaf(pp, v23.x, o_acc[r][2]);
            o_acc[r][3] = fmaf(pp, v23.y, o_acc[r][3]);
          }
        }
      }
    } else
    {
      const uint32_t vs_a = stage_a + uint32_t(stage) * L::kStageBytes;   // v_sw already includes the V tensor offset
#pragma unroll
      for (int jj = 0; jj < kRowsPerWarp; jj += 4) {
        float4 pr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) pr[r] = ld_shared_v4f(ps_a + uint32_t(((warp * R + r) * kRowsPerWarp + jj) * 4));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t va = vs_a + v_sw[(jj + u) & 7] + uint32_t((jj + u) * 128);
          float vf[EPL];
          if constexpr (EPL == 4) {
            const uint2 w = ld_shared_v2u(va);
            if constexpr (BF16) { vf[0] = bf16lo(w.x); vf[1] = bf16hi(w.x); vf[2] = bf16lo(w.y); vf[3] = bf16hi(w.y); }
            else { vf[0] = f16lo(w.x); vf[1] = f16hi(w.x); vf[2] = f16lo(w.y); vf[3] = f16hi(w.y); }
          } else {
            const uint32_t w = ld_shared_u32(va);
            if constexpr (BF16) { vf[0] = bf16lo(w); vf[1] = bf16hi(w); }
            else { vf[0] = f16lo(w); vf[1] = f16hi(w); }
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float pp = (u == 0) ? pr[r].x : (u == 1) ? pr[r].y : (u == 2) ? pr[r].z : pr[r].w;
#pragma unroll
            for (int e = 0; e < EPL; ++e) o_acc[r][e] = fmaf(pp, vf[e], o_acc[r][e]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == NS) { stage = 0; phase ^= 1; }
  }
  if (trc && tid == 0) trc[5] = clock64();
  if (cur_x >= 0) finalize_segment(cur_x);
  if (trc && tid == 0) trc[6] = clock64();

  // ------------- deferred cross-GPU merges for the heads this CTA finished, end-of-kernel arrival -----
  const dcomm::Tail tail = make_tail();
  dcomm::drain<D, kConsumerThreads, RB>(tail, n_pend, tid, 1, store_out);
  if (trc && tid == 0) { trc[8] = clock64(); trc[9] = globaltimer_ns(); }
}

template <int D, int R, bool BF16, bool KV8>
std::function<void(cudaStream_t)> make_pass(const CUtensorMap& kmap, const CUtensorMap& vmap, const DecodeParams& p, int grid) {
  auto kern = decode_simt_kernel<D, R, BF16, KV8>;
  constexpr size_t smem = smem_bytes<D, R, KV8>();
  static bool configured = false;
  if (!configured) {
    TA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // the tensor maps and the parameter block are captured BY VALUE: a prepared pass re-launches with one runtime call
  return [kern, kmap, vmap, p, grid](cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes<D, R, KV8>();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    TA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, kmap, vmap, p));
  };
}

int pick_rows(int total_rows) { return total_rows >= 4 ? 4 : (total_rows >= 2 ? 2 : 1); }

unsigned long long* g_decode_trace = nullptr;   // set by decode_simt_set_trace (profiling aid)

}  // namespace

void decode_simt_set_trace(unsigned long long* buf) { g_decode_trace = buf; }

void decode_split(int BH, int cap, int ncta, int* grid, int* max_parts) {
  const int tph = std::max(1, (cap + kTileRows - 1) / kTileRows);
  const long long total = (long long)BH * tph;
  const int g = (int)std::min<long long>(ncta, total);
  // The kernels split BH * max(1, ceil(s / 128)) tiles over the SAME g CTAs for whatever number s <= cap of rows is
  // valid at run time (device-resident kv_len).  Every CTA then owns >= floor(BH * tph(s) / g) tiles, which bounds the
  // CTAs that share one head by ceil(2 g / BH) + 1 for every s (tests/test_split_cpu.py checks this exhaustively).
  const int mp = std::min<long long>(g, (2LL * g + BH - 1) / BH + 1);
  *grid = g;
  *max_parts = std::max(mp, 1);
}

void decode_simt_plan(const AttnShape& s, int nsm, int* grid, int* max_parts, int* rows_per_pass,
                      size_t* part_floats, size_t* comm_floats, size_t* comm_flags) {
  const int BH = s.B * s.Hkv;
  decode_split(BH, s.S, nsm, grid, max_parts);
  const int R = pick_rows((s.Hq / s.Hkv) * s.Sq);
  *rows_per_pass = R;
  *part_floats = (size_t)BH * *max_parts * R * (s.D + 2) * 2;   // tagged 8-byte words, counted in floats
  *comm_floats = (size_t)2 * kMaxWorldHost * BH * R * (s.D + 2) * 2;
  *comm_flags = (size_t)2 * kMaxWorldHost * BH;
}

PreparedLaunch decode_simt_prepare(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                                   float* part, uint32_t* tickets, const CommCtxHost& comm, int nsm,
                                   const uint32_t* kscale, const uint32_t* vscale, int pdl, const int* kv_len) {
  if ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(tickets)) & 7)
    throw std::runtime_error("decode_simt: workspace must be 8-byte aligned");
  const bool kv8 = kscale != nullptr;
  if (s.D != 64 && s.D != 128) throw std::runtime_error("decode_simt: head_dim must be 64 or 128");
  if (kv8 && s.D != 128) throw std::runtime_error("decode_simt(mxfp8): head_dim must be 128");
  if (s.Hq % s.Hkv != 0) throw std::runtime_error("decode_simt: Hq must be a multiple of Hkv");
  if (s.S <= 0) throw std::runtime_error("decode_simt: the KV shard must have capacity for at least one row");
  int grid, max_parts, R;
  size_t pf, cf, cfl;
  decode_simt_plan(s, nsm, &grid, &max_parts, &R, &pf, &cf, &cfl);
  const int G = s.Hq / s.Hkv;
  const int total_rows = G * s.Sq;
  // a CTA queues the heads whose last tile it owns (at most one per head in its range, ranges shrink to one tile per head
  // for an almost empty cache): (B * Hkv) / grid + 2 of them at worst
  if ((long long)s.B * s.Hkv / grid + 2 > kMaxPending)
    throw std::runtime_error("decode_simt: batch x kv-heads too large for one launch (split the batch)");
  if (comm.world > 1) {
    const size_t need_data = (size_t)2 * comm.world * s.B * s.Hkv * R * (s.D + 2) * 8;
    if (need_data > comm.data_bytes)
      throw std::runtime_error("decode_simt: symmetric buffer too small for this problem");
  }
  const int eb = kv8 ? 1 : 2;
  CUtensorMap kmap = make_tmap_bhsd(k, eb, s.B, s.Hkv, s.S, s.D, s.k_sb, s.k_sh, s.k_ss, 128 / eb, kTileRows,
                                    CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap vmap = make_tmap_bhsd(v, eb, s.B, s.Hkv, s.S, s.D, s.v_sb, s.v_sh, s.v_ss, 128 / eb, kTileRows,
                                    CU_TENSOR_MAP_SWIZZLE_128B);
  DecodeParams p;
  p.kscale = kscale; p.vscale = vscale; p.pdl = pdl; p.kv_len = kv_len; p.trace = g_decode_trace;
  p.q = q; p.out = out; p.lse = lse;
  p.part = reinterpret_cast<uint64_t*>(part); p.wctr = reinterpret_cast<unsigned long long*>(tickets);
  p.B = s.B; p.Hq = s.Hq; p.Hkv = s.Hkv; p.G = G; p.Sq = s.Sq; p.S = s.S;
  p.scale_log2 = s.softmax_scale * 1.4426950408889634f;
  p.causal = s.causal; p.q_pos0 = s.q_pos0; p.kv_pos0 = s.kv_pos0;
  p.q_sb = s.q_sb; p.q_sh = s.q_sh; p.q_ss = s.q_ss; p.o_sb = s.o_sb; p.o_sh = s.o_sh; p.o_ss = s.o_ss;
  p.max_parts = max_parts;
  p.comm = dcomm::to_device_ctx(comm);
  PreparedLaunch pl;
  for (int r_base = 0; r_base < total_rows; r_base += R) {
    p.r_base = r_base;
    p.rows_valid = std::min(R, total_rows - r_base);
#define TA_PASS(DD, RR, KV)                                                                       \
  pl.passes.push_back(s.is_bf16 ? make_pass<DD, RR, true, KV>(kmap, vmap, p, grid)               \
                                : make_pass<DD, RR, false, KV>(kmap, vmap, p, grid));
    if (kv8) {
      if (R == 4) { TA_PASS(128, 4, true) } else if (R == 2) { TA_PASS(128, 2, true) } else { TA_PASS(128, 1, true) }
    } else if (s.D == 128) {
      if (R == 4) { TA_PASS(128, 4, false) } else if (R == 2) { TA_PASS(128, 2, false) } else { TA_PASS(128, 1, false) }
    } else {
      if (R == 4) { TA_PASS(64, 4, false) } else if (R == 2) { TA_PASS(64, 2, false) } else { TA_PASS(64, 1, false) }
    }
#undef TA_PASS
  }
  return pl;
}

void decode_simt_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                        float* part, uint32_t* tickets, const CommCtxHost& comm, int nsm, cudaStream_t stream,
                        const uint32_t* kscale, const uint32_t* vscale, int pdl, const int* kv_len) {
  decode_simt_prepare(s, q, k, v, out, lse, part, tickets, comm, nsm, kscale, vscale, pdl, kv_len).run(stream);
}

}  // namespace ta
