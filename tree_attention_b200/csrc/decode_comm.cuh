// decode_comm.cuh -- the tail every fused decode kernel shares (decode_simt.cu, decode_tc_sm100.cu, decode_swap_sm100.cu):
//
//   1. stream-K geometry of the flattened (batch, kv-head, 128-key tile) space, computed IN the kernel from the
//      (possibly device-resident, i.e. CUDA-graph-replayable) number of valid keys of this rank's shard;
//   2. split merge WITHOUT fences, tickets or atomics on the critical path: every CTA writes its partial of a head as
//      NCCL-LL style 8-byte words {fp32 value, launch tag} (a naturally aligned 64-bit store is single-copy atomic, so
//      every word validates itself); the CTA that owns the head's LAST tile is the head's merger: after its own last
//      tile it polls the nparts x (o[d], m, l) words with all loads in flight and merges them in part order.  The merger
//      has the highest block index of the head, i.e. it only ever waits for CTAs the hardware dispatched before it
//      (the forward-progress assumption of every decoupled look-back scan), and the wait is bounded;
//   3. cross-GPU combine: the merged, normalised (o, lse) goes to slot [parity][my rank][head] of EVERY rank as the same
//      kind of tagged word over NVLink; the consumer polls the W x (o, lse) words of a head with ALL loads in flight
//      (one NVLink-write-visible round trip instead of 2 W serial ones -- the round-1 kernels polled word after word,
//      which is what their 13-15 us "wait_peers" was) and merges in rank order => bitwise identical on every rank;
//   4. launch tags from arrival counters: each CTA adds to a 64-bit device-resident counter when it starts (CTA 0 adds
//      the complement to 4096), tag = counter / 4096 + 1.  No CTA has to stay behind to bump an epoch when the kernel
//      drains, the value is needed only at the first segment end (the atomic's latency is hidden behind the first
//      tiles), and a captured CUDA graph replays correctly.  One counter in the workspace tags the intra-GPU partials,
//      one in the symmetric region tags the cross-GPU words (parity of the latter double-buffers the slots).
//
// Replaces the combine of /root/reference/model.py:103-124 (all_reduce MAX, SUM, SUM + rescale + divide).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ta {
namespace dcomm {

constexpr int kTileKeys = 128;
constexpr int kPartChunk = 8;   // CTA partials loaded per batch by the last-arriver merge

__device__ __forceinline__ float ninf() { return __int_as_float(0xff800000); }

// ---------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------
struct Geom {
  int S;          // valid keys of this shard (<= capacity)
  int tph;        // 128-key tiles per (batch, kv-head): max(1, ceil(S / 128)) -- an EMPTY shard still runs one fully
                  // masked tile per head so that every head publishes the monoid identity (peers wait for it)
  int total;      // BH * tph
  int tiles_q, tiles_rem;
};

// kv_len: optional device scalar (number of valid rows, clamped to [0, cap]); null = the whole shard is valid
__device__ __forceinline__ Geom make_geom(int cap, const int* kv_len, int BH, int ncta) {
  Geom g;
  int s = cap;
  if (kv_len != nullptr) s = max(0, min(cap, *reinterpret_cast<const volatile int*>(kv_len)));
  g.S = s;
  g.tph = max(1, (s + kTileKeys - 1) / kTileKeys);
  g.total = BH * g.tph;
  g.tiles_q = g.total / ncta;
  g.tiles_rem = g.total - g.tiles_q * ncta;
  return g;
}
__device__ __forceinline__ int cta_lo(const Geom& g, int c) { return c * g.tiles_q + min(c, g.tiles_rem); }
__device__ __forceinline__ int cta_of_tile(const Geom& g, int t) {
  const int big = g.tiles_rem * (g.tiles_q + 1);
  return t < big ? t / (g.tiles_q + 1) : g.tiles_rem + (t - big) / max(g.tiles_q, 1);
}

// ---------------------------------------------------------------------------------------------------------
// tagged words: {fp32 bits (low), launch tag (high)} in one naturally aligned 64-bit access
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store_sys(uint64_t* w, float v, uint32_t tag) {
  const uint64_t x = ((uint64_t)tag << 32) | (uint64_t)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(w), "l"(x) : "memory");
}
__device__ __forceinline__ uint64_t ll_load_sys(const uint64_t* w) {
  uint64_t r;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(r) : "l"(w) : "memory");
  return r;
}
__device__ __forceinline__ void ll_store_gpu(uint64_t* w, float v, uint32_t tag) {
  const uint64_t x = ((uint64_t)tag << 32) | (uint64_t)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(w), "l"(x) : "memory");
}
__device__ __forceinline__ uint64_t ll_load_gpu(const uint64_t* w) {
  uint64_t r;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(w) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t ll_tag(uint64_t w) { return (uint32_t)(w >> 32); }
__device__ __forceinline__ float ll_val(uint64_t w) { return __uint_as_float((uint32_t)w); }

// slow paths: spin on one word until its tag matches, bounded (the caller detects a timeout by the returned tag).  By
// value on purpose: a reference parameter of a non-inlined function would force the caller's whole batch of in-flight
// words into local memory.
static __device__ __noinline__ uint64_t ll_spin_sys(const uint64_t* w, uint32_t tag, unsigned long long timeout_ns) {
  const uint64_t t0 = globaltimer_ns();
  uint32_t it = 0;
  uint64_t out;
  while (true) {
    out = ll_load_sys(w);
    if (ll_tag(out) == tag) break;
    if ((++it & 0x3fu) == 0 && globaltimer_ns() - t0 > timeout_ns) break;
  }
  return out;
}
static __device__ __noinline__ uint64_t ll_spin_gpu(const uint64_t* w, uint32_t tag) {
  const uint64_t t0 = globaltimer_ns();
  uint32_t it = 0;
  uint64_t out;
  while (true) {
    out = ll_load_gpu(w);
    if (ll_tag(out) == tag) break;
    if ((++it & 0x3fu) == 0 && globaltimer_ns() - t0 > TA_SPIN_TIMEOUT_NS) {
      // a CTA of THIS launch never delivered its partial: a bug (or a device without forward progress of earlier
      // blocks), never a peer problem -- fail the launch loudly instead of hanging (same policy as mbar_wait)
      printf("[tree_attention] split-merge wait timeout block=%d thread=%d tag=%u\n", blockIdx.x, threadIdx.x, tag);
      __trap();
    }
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------
// launch tags from arrival counters
// ---------------------------------------------------------------------------------------------------------
constexpr int kTagShift = 12;   // every launch advances a counter by exactly 4096 (grids are <= #SMs <= 4096 CTAs)

// One thread per CTA, once per launch and counter, AFTER everything the launch depends on has completed (stream order, or
// griddepcontrol.wait under programmatic dependent launch): then all arrivals of launch n lie in [4096 n, 4096 (n + 1)).
__device__ __forceinline__ uint32_t launch_tag(unsigned long long* ctr) {
  const unsigned long long inc = blockIdx.x == 0 ? (unsigned long long)((1u << kTagShift) - (gridDim.x - 1)) : 1ull;
  const unsigned long long old = atomicAdd(ctr, inc);
  return (uint32_t)(old >> kTagShift) + 1u;
}

// everything the tail needs; built on demand by the kernels (kept out of the streaming loop's live registers)
struct Tail {
  const CommCtx* comm;   // points into the kernel's __grid_constant__ parameter block
  uint64_t* part;        // workspace: [BH][max_parts][R][D + 2] tagged words: o (unnormalised) | m | l   (log2 domain)
  int max_parts;
  int BH;
  int R;                 // rows per head in the part / word layouts
  int rows_valid;        // rows that are real (<= R)
  uint32_t wtag;         // tag of this launch's intra-GPU partial words (workspace counter)
  uint32_t ctag;         // tag of this launch's cross-GPU words (region counter); its parity selects the slot set
  Geom geo;
  int* pending;          // smem: heads this CTA merges (it owns their last tile), merged after the CTA's last tile
  int max_pending;
  uint64_t* stamps;      // smem (thread 0 only): [0] globaltimer at CTA start, [1] at the last publish (0 = none yet)
};

template <int D>
__device__ __forceinline__ uint64_t* part_ptr(const Tail& t, int x, int pidx) {
  return t.part + ((size_t)x * t.max_parts + pidx) * (size_t)(t.R * (D + 2));
}
template <int D>
__device__ __forceinline__ uint64_t* word_ptr(const Tail& t, int dst, int src, int x) {
  return reinterpret_cast<uint64_t*>(t.comm->data[dst]) +
         ((size_t)((t.ctag & 1) * t.comm->world + src) * t.BH + x) * (size_t)(t.R * (D + 2));
}
// CTAs that share head x, and this CTA's index among them
__device__ __forceinline__ void head_parts(const Geom& g, int x, int cta, int& nparts, int& pidx) {
  const int first_cta = cta_of_tile(g, x * g.tph);
  nparts = cta_of_tile(g, (x + 1) * g.tph - 1) - first_cta + 1;
  pidx = cta - first_cta;
}

// ---------------------------------------------------------------------------------------------------------
// merger: gather the nparts partials of row r, channel d of head x (all loads of a batch in flight), merge in part order
// ---------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void merge_parts(const Tail& t, int x, int nparts, int r, int d, float& o_norm, float& lse2) {
  const size_t pstride = (size_t)t.R * (D + 2);
  const uint64_t* base = part_ptr<D>(t, x, 0) + (size_t)r * (D + 2);
  float M = ninf(), acc = 0.f, L = 0.f;
  for (int q0 = 0; q0 < nparts; q0 += kPartChunk) {
    uint64_t wa[kPartChunk], wm[kPartChunk], wl[kPartChunk];
#pragma unroll
    for (int i = 0; i < kPartChunk; ++i) {
      if (q0 + i < nparts) {
        const uint64_t* pp = base + (size_t)(q0 + i) * pstride;
        wa[i] = ll_load_gpu(pp + d); wm[i] = ll_load_gpu(pp + D); wl[i] = ll_load_gpu(pp + D + 1);
      }
    }
    float m[kPartChunk], a[kPartChunk], l[kPartChunk];
    float Mn = M;
#pragma unroll
    for (int i = 0; i < kPartChunk; ++i) {
      m[i] = ninf(); a[i] = 0.f; l[i] = 0.f;
      if (q0 + i < nparts) {
        const uint64_t* pp = base + (size_t)(q0 + i) * pstride;
        if (ll_tag(wa[i]) != t.wtag) wa[i] = ll_spin_gpu(pp + d, t.wtag);
        if (ll_tag(wm[i]) != t.wtag) wm[i] = ll_spin_gpu(pp + D, t.wtag);
        if (ll_tag(wl[i]) != t.wtag) wl[i] = ll_spin_gpu(pp + D + 1, t.wtag);
        a[i] = ll_val(wa[i]); m[i] = ll_val(wm[i]); l[i] = ll_val(wl[i]);
        Mn = fmaxf(Mn, m[i]);
      }
    }
    const float Ms = (Mn == ninf()) ? 0.f : Mn;
    const float sc0 = fast_exp2(M - Ms);   // M = -inf -> 0 (acc and L are 0 then)
    acc *= sc0; L *= sc0;
#pragma unroll
    for (int i = 0; i < kPartChunk; ++i) {
      const float sc = fast_exp2(m[i] - Ms);
      acc = fmaf(a[i], sc, acc);
      L = fmaf(l[i], sc, L);
    }
    M = Mn;
  }
  const float Ms = (M == ninf()) ? 0.f : M;
  o_norm = L > 0.f ? acc / L : 0.f;
  lse2 = L > 0.f ? Ms + fast_log2(L) : ninf();
}

// publish one element (and the row's lse, by the d == 0 thread) to every rank, own slot included
template <int D>
__device__ __forceinline__ void publish(const Tail& t, int x, int r, int d, float o_norm, float lse2) {
  const int world = t.comm->world;
#pragma unroll 4
  for (int dst = 0; dst < world; ++dst) {
    uint64_t* wp = word_ptr<D>(t, dst, t.comm->rank, x) + r * (D + 2);
    ll_store_sys(wp + d, o_norm, t.ctag);
    if (d == 0) ll_store_sys(wp + D, lse2, t.ctag);
  }
}

// Gather the W published (o[d], lse) words of rows r0 .. r0 + nr - 1 of head x and merge them in rank order.
// All 2 * nr * min(W, kSrcChunk) loads of a batch are in flight before the first tag is examined.
// RB = query rows of one output channel handled per batch.
template <int D, int RB>
__device__ __forceinline__ bool gather_rows(const Tail& t, int x, int r0, int nr, int d, float (&o_norm)[RB],
                                            float (&lse2)[RB], int& bad_src) {
  constexpr int kSrcChunk = RB >= 4 ? 4 : 8;   // peers polled per batch: <= 32 rows x sources x 2 words in flight per thread
  const int world = t.comm->world;
  const size_t src_stride = (size_t)t.BH * t.R * (D + 2);
  const uint64_t* w0 = word_ptr<D>(t, t.comm->rank, 0, x);
  float m_run[RB], num[RB], den[RB];
#pragma unroll
  for (int i = 0; i < RB; ++i) { m_run[i] = ninf(); num[i] = 0.f; den[i] = 0.f; }
  bool ok = true;
  for (int s0 = 0; s0 < world; s0 += kSrcChunk) {
    uint64_t lw[RB][kSrcChunk], vw[RB][kSrcChunk];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
#pragma unroll
      for (int s = 0; s < kSrcChunk; ++s) {
        if (i < nr && s0 + s < world) {
          const uint64_t* wr = w0 + (size_t)(s0 + s) * src_stride + (size_t)(r0 + i) * (D + 2);
          lw[i][s] = ll_load_sys(wr + D);
          vw[i][s] = ll_load_sys(wr + d);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      if (i < nr) {
        float ls[kSrcChunk], vs[kSrcChunk];
        float mc = m_run[i];
#pragma unroll
        for (int s = 0; s < kSrcChunk; ++s) {
          ls[s] = ninf(); vs[s] = 0.f;
          if (s0 + s < world) {
            const uint64_t* wr = w0 + (size_t)(s0 + s) * src_stride + (size_t)(r0 + i) * (D + 2);
            if (ll_tag(lw[i][s]) != t.ctag) lw[i][s] = ll_spin_sys(wr + D, t.ctag, t.comm->timeout_ns);
            if (ll_tag(vw[i][s]) != t.ctag) vw[i][s] = ll_spin_sys(wr + d, t.ctag, t.comm->timeout_ns);
            if (ll_tag(lw[i][s]) != t.ctag || ll_tag(vw[i][s]) != t.ctag) { ok = false; bad_src = s0 + s; }
            ls[s] = ll_val(lw[i][s]);
            vs[s] = ll_val(vw[i][s]);
            mc = fmaxf(mc, ls[s]);
          }
        }
        const float ms = (mc == ninf()) ? 0.f : mc;
        const float sc0 = fast_exp2(m_run[i] - ms);
        num[i] *= sc0; den[i] *= sc0;
#pragma unroll
        for (int s = 0; s < kSrcChunk; ++s) {
          const float w = fast_exp2(ls[s] - ms);   // absent source: exp2(-inf) = 0
          num[i] = fmaf(w, vs[s], num[i]);
          den[i] += w;
        }
        m_run[i] = mc;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RB; ++i) {
    const float ms = (m_run[i] == ninf()) ? 0.f : m_run[i];
    o_norm[i] = den[i] > 0.f ? num[i] / den[i] : 0.f;
    lse2[i] = den[i] > 0.f ? ms + fast_log2(den[i]) : ninf();
  }
  return ok;
}

// merge the W published partials of head x and hand every element to store_out(x, r, d, o_norm, lse2 /*log2 domain*/).
// NT threads (tid in [0, NT), NT a multiple of D) take part; bar_id is a named barrier private to those threads.
// (noinline on purpose: the tail runs once per head; keeping it out of line keeps its registers out of the streaming loop)
template <int D, int NT, int RB, typename StoreOut>
__device__ __noinline__ void combine_ranks(const Tail& t, int x, int tid, int bar_id, StoreOut& store_out) {
  static_assert(NT % D == 0, "thread count must be a multiple of head_dim");
  constexpr int NG = NT / D;
  const int d = tid % D, g = tid / D;
  uint64_t t_got = 0;
  for (int r0 = g * RB; r0 < t.rows_valid; r0 += NG * RB) {
    const int nr = min(RB, t.rows_valid - r0);
    float o_norm[RB], lse2[RB];
    int bad_src = -1;
    const bool ok = gather_rows<D, RB>(t, x, r0, nr, d, o_norm, lse2, bad_src);
    if (tid == 0 && r0 == 0) t_got = globaltimer_ns();
    if (!ok) {
      t.comm->status[0] = kCommTimeout; t.comm->status[1] = x; t.comm->status[2] = bad_src; t.comm->status[3] = t.ctag;
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      if (i < nr) {
        const float nan = __int_as_float(0x7fc00000);
        store_out(x, r0 + i, d, ok ? o_norm[i] : nan, ok ? lse2[i] : nan);
      }
    }
  }
  if (tid == 0 && t.stamps[1] != 0) {  // in-kernel stamps of the combine step
    const uint64_t t_done = globaltimer_ns(), t_publish = t.stamps[1], t_cta0 = t.stamps[0];
    atomicMax(t.comm->status + 10, (uint32_t)min((unsigned long long)(t_got - t_publish), 0xffffffffull));
    atomicMax(t.comm->status + 11, (uint32_t)min((unsigned long long)(t_done - t_publish), 0xffffffffull));
    atomicMax(t.comm->status + 12, (uint32_t)min((unsigned long long)(t_publish - t_cta0), 0xffffffffull));
  }
}

// merger side of the split merge for head x: poll + merge the parts, then write the result (world == 1) or publish it
template <int D, int NT, typename StoreOut>
__device__ __noinline__ void merge_head(const Tail& t, int x, int tid, StoreOut& store_out) {
  static_assert(NT % D == 0, "thread count must be a multiple of head_dim");
  constexpr int NG = NT / D;
  int nparts, pidx;
  head_parts(t.geo, x, blockIdx.x, nparts, pidx);
  const int world = t.comm->world;
  const int d = tid % D;
  for (int r = tid / D; r < t.rows_valid; r += NG) {
    float o_norm, lse2;
    merge_parts<D>(t, x, nparts, r, d, o_norm, lse2);
    if (world == 1) store_out(x, r, d, o_norm, lse2);
    else if (!t.comm->skip_publish) publish<D>(t, x, r, d, o_norm, lse2);
  }
}

// The CTA partial of head x has been written to the workspace as tagged words by the calling threads (part_ptr /
// ll_store_gpu; nothing to fence).  If this CTA owns the head's last tile it is the head's merger: queue the head; the
// merge itself happens after the CTA's last tile (drain), when the other CTAs' words are on their way.
// No barrier on the common path (a head switch must not stall the stream): `n_pend` is the number of queued heads, kept
// in a register by EVERY calling thread (all threads see the same segments), thread 0 mirrors the heads into smem.
template <int D, int NT, int RB, typename StoreOut>
__device__ __forceinline__ void segment_done(const Tail& t, int x, int seg_end_tile, int& n_pend, int tid, int bar_id,
                                             StoreOut& store_out) {
  if (seg_end_tile != (x + 1) * t.geo.tph) return;   // not the owner of the head's last tile: nothing else to do
  // The list cannot overflow: the launchers bound the heads per CTA by max_pending (decode_*_prepare).  Nothing here may
  // call out of line -- an out-of-line call inside the streaming loop makes ptxas spill around it, and with 227 KB of
  // shared memory carved out there is no L1 left to catch the spills (round 2: 10 us per head switch).
  if (n_pend >= t.max_pending) {
    printf("[tree_attention] pending-merge list overflow block=%d head=%d\n", blockIdx.x, x);
    __trap();
  }
  if (tid == 0) t.pending[n_pend] = x;
  ++n_pend;
}

// after the CTA's last tile: merge (and publish) every queued head, THEN wait for the peers' words of those heads --
// all publishes are in flight over NVLink before the first wait
template <int D, int NT, int RB, typename StoreOut>
__device__ __forceinline__ void drain(const Tail& t, int n_pend, int tid, int bar_id, StoreOut& store_out) {
  named_bar_sync(bar_id, NT);   // thread 0's pending[] writes
  for (int i = 0; i < n_pend; ++i) merge_head<D, NT>(t, t.pending[i], tid, store_out);
  if (t.comm->world <= 1 || n_pend == 0) return;
  if (tid == 0) t.stamps[1] = globaltimer_ns();
  named_bar_sync(bar_id, NT);
  for (int i = 0; i < n_pend; ++i) combine_ranks<D, NT, RB>(t, t.pending[i], tid, bar_id, store_out);
}

inline CommCtx to_device_ctx(const CommCtxHost& h) {
  CommCtx c;
  c.rank = h.rank;
  c.world = h.world;
  for (int i = 0; i < kMaxWorld; ++i) {
    c.data[i] = reinterpret_cast<float*>(h.data[i]);
    c.flags[i] = reinterpret_cast<uint32_t*>(h.flags[i]);
  }
  c.epoch = reinterpret_cast<uint32_t*>(h.epoch);
  c.status = reinterpret_cast<uint32_t*>(h.status);
  c.timeout_ns = h.timeout_ns;
  c.skip_publish = h.skip_publish;
  return c;
}

}  // namespace dcomm
}  // namespace ta
