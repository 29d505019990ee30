// decode_comm.cuh -- the tail every fused decode kernel shares (decode_simt.cu, decode_tc_sm100.cu, decode_swap_sm100.cu):
//
//   1. stream-K geometry of the flattened (batch, kv-head, 128-key tile) space, computed IN the kernel from the
//      (possibly device-resident, i.e. CUDA-graph-replayable) number of valid keys of this rank's shard;
//   2. split merge: CTA partial -> workspace -> atomic ticket -> the last CTA of a (batch, kv-head) merges the
//      parts in part order;
//   3. cross-GPU combine: the merged, normalised (o, lse) goes to slot [parity][my rank][head] of EVERY rank as
//      NCCL-LL style 8-byte words {fp32 value, epoch tag}; the consumer polls the W x (o, lse) words of a head
//      with ALL loads in flight at once (one NVLink-write-visible round trip instead of 2 W serial ones -- the
//      round-1 kernels polled word after word, which is what their 13-15 us "wait_peers" was) and merges in rank
//      order, so the result is bitwise identical on every rank.
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
// LL words
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint2* w, float v, uint32_t epoch) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(w), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 ll_load(const uint2* w) {
  uint2 r;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(w) : "memory");
  return r;
}
// slow path: spin on one word until its tag is `epoch`; bounded by the communicator's timeout (the caller detects a
// timeout by the tag of the returned word).  By value on purpose: a reference parameter of a non-inlined function would
// force the caller's whole batch of in-flight words into local memory.
static __device__ __noinline__ uint2 ll_spin(const uint2* w, uint32_t epoch, unsigned long long timeout_ns) {
  const uint64_t t0 = globaltimer_ns();
  uint32_t it = 0;
  uint2 out;
  while (true) {
    out = ll_load(w);
    if (out.y == epoch) break;
    if ((++it & 0x3fu) == 0 && globaltimer_ns() - t0 > timeout_ns) break;
  }
  return out;
}

// everything the tail needs; one per kernel instance, lives in registers
struct Tail {
  const CommCtx* comm;  // points into the kernel's __grid_constant__ parameter block
  float* part;         // [BH][max_parts][R][D + 4]
  uint32_t* tickets;   // [BH] head tickets, [BH] = exit counter
  int max_parts;
  int BH;
  int R;               // rows per head in the part / word layouts
  int rows_valid;      // rows that are real (<= R)
  uint32_t epoch;
  int parity;
  int* s_misc;         // smem: [0] ticket, [1] n_pending, [2] head + 1 to combine inline (pending list full)
  int* pending;        // smem: heads this CTA finished and still has to merge across ranks
  int max_pending;
  uint64_t* stamps;    // smem (thread 0 only): [0] globaltimer at CTA start, [1] at the last publish (0 = none yet)
};

template <int D>
__device__ __forceinline__ uint2* word_ptr(const Tail& t, int dst, int src, int x) {
  return reinterpret_cast<uint2*>(t.comm->data[dst]) + ((size_t)(t.parity * t.comm->world + src) * t.BH + x) * (size_t)(t.R * (D + 2));
}

// ---------------------------------------------------------------------------------------------------------
// last-arriver merge of the CTA partials of one head (part order => deterministic); loads are issued kPartChunk
// parts at a time so a head split over n CTAs costs ceil(n / 8) L2 round trips instead of 2 n
// ---------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void merge_parts(const float* parts, int nparts, int R, int r, int d, float& o_norm, float& lse2) {
  const size_t pstride = (size_t)R * (D + 4);
  const float* base = parts + (size_t)r * (D + 4);
  float M = ninf(), acc = 0.f, L = 0.f;
  for (int q0 = 0; q0 < nparts; q0 += kPartChunk) {
    float m[kPartChunk], a[kPartChunk], l[kPartChunk];
#pragma unroll
    for (int i = 0; i < kPartChunk; ++i) {
      m[i] = ninf(); a[i] = 0.f; l[i] = 0.f;
      if (q0 + i < nparts) {
        const float* pp = base + (size_t)(q0 + i) * pstride;
        m[i] = __ldcg(pp + D); l[i] = __ldcg(pp + D + 1); a[i] = __ldcg(pp + d);
      }
    }
    float Mn = M;
#pragma unroll
    for (int i = 0; i < kPartChunk; ++i) Mn = fmaxf(Mn, m[i]);
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
    uint2* wp = word_ptr<D>(t, dst, t.comm->rank, x) + r * (D + 2);
    ll_store(wp + d, o_norm, t.epoch);
    if (d == 0) ll_store(wp + D, lse2, t.epoch);
  }
}

// Gather the W published (o[d], lse) words of rows r0 .. r0 + nr - 1 of head x and merge them in rank order.
// All 2 * nr * min(W, kSrcChunk) loads of a batch are in flight before the first tag is examined.
// RB = query rows of one output channel handled per batch.
template <int D, int RB>
__device__ __forceinline__ bool gather_rows(const Tail& t, int x, int r0, int nr, int d, float (&o_norm)[RB],
                                            float (&lse2)[RB], int& bad_src) {
  constexpr int kRowBatch = RB;
  constexpr int kSrcChunk = RB >= 4 ? 4 : 8;   // peers polled per batch: <= 32 rows x sources x 2 words in flight per thread
  const int world = t.comm->world;
  const size_t src_stride = (size_t)t.BH * t.R * (D + 2);
  const uint2* w0 = word_ptr<D>(t, t.comm->rank, 0, x);
  float m_run[kRowBatch], num[kRowBatch], den[kRowBatch];
#pragma unroll
  for (int i = 0; i < kRowBatch; ++i) { m_run[i] = ninf(); num[i] = 0.f; den[i] = 0.f; }
  bool ok = true;
  for (int s0 = 0; s0 < world; s0 += kSrcChunk) {
    uint2 lw[kRowBatch][kSrcChunk], vw[kRowBatch][kSrcChunk];
#pragma unroll
    for (int i = 0; i < kRowBatch; ++i) {
#pragma unroll
      for (int s = 0; s < kSrcChunk; ++s) {
        if (i < nr && s0 + s < world) {
          const uint2* wr = w0 + (size_t)(s0 + s) * src_stride + (size_t)(r0 + i) * (D + 2);
          lw[i][s] = ll_load(wr + D);
          vw[i][s] = ll_load(wr + d);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kRowBatch; ++i) {
      if (i < nr) {
        float ls[kSrcChunk], vs[kSrcChunk];
        float mc = m_run[i];
#pragma unroll
        for (int s = 0; s < kSrcChunk; ++s) {
          ls[s] = ninf(); vs[s] = 0.f;
          if (s0 + s < world) {
            const uint2* wr = w0 + (size_t)(s0 + s) * src_stride + (size_t)(r0 + i) * (D + 2);
            if (lw[i][s].y != t.epoch) lw[i][s] = ll_spin(wr + D, t.epoch, t.comm->timeout_ns);
            if (vw[i][s].y != t.epoch) vw[i][s] = ll_spin(wr + d, t.epoch, t.comm->timeout_ns);
            if (lw[i][s].y != t.epoch || vw[i][s].y != t.epoch) { ok = false; bad_src = s0 + s; }
            ls[s] = __uint_as_float(lw[i][s].x);
            vs[s] = __uint_as_float(vw[i][s].x);
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
  for (int i = 0; i < kRowBatch; ++i) {
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
  constexpr int kRowBatch = RB;
  const int d = tid % D, g = tid / D;
  uint64_t t_got = 0;
  for (int r0 = g * kRowBatch; r0 < t.rows_valid; r0 += NG * kRowBatch) {
    const int nr = min(kRowBatch, t.rows_valid - r0);
    float o_norm[kRowBatch], lse2[kRowBatch];
    int bad_src = -1;
    const bool ok = gather_rows<D, RB>(t, x, r0, nr, d, o_norm, lse2, bad_src);
    if (tid == 0 && r0 == 0) t_got = globaltimer_ns();
    if (!ok) {
      t.comm->status[0] = kCommTimeout; t.comm->status[1] = x; t.comm->status[2] = bad_src; t.comm->status[3] = t.epoch;
    }
#pragma unroll
    for (int i = 0; i < kRowBatch; ++i) {
      if (i < nr) {
        const float nan = __int_as_float(0x7fc00000);
        store_out(x, r0 + i, d, ok ? o_norm[i] : nan, ok ? lse2[i] : nan);
      }
    }
  }
  named_bar_sync(bar_id, NT);
  if (tid == 0 && t.stamps[1] != 0) {  // in-kernel stamps of the combine step
    const uint64_t t_done = globaltimer_ns(), t_publish = t.stamps[1], t_cta0 = t.stamps[0];
    atomicMax(t.comm->status + 10, (uint32_t)min((unsigned long long)(t_got - t_publish), 0xffffffffull));
    atomicMax(t.comm->status + 11, (uint32_t)min((unsigned long long)(t_done - t_publish), 0xffffffffull));
    atomicMax(t.comm->status + 12, (uint32_t)min((unsigned long long)(t_publish - t_cta0), 0xffffffffull));
  }
}

// The CTA partial of head x has been written to the workspace by the NT calling threads (plain / .cg stores, not yet
// fenced).  Take a ticket; the last CTA of the head merges all parts, then either writes the result (world == 1) or
// publishes it to every rank and queues the head for the deferred cross-GPU merge.
template <int D, int NT, int RB, typename StoreOut>
__device__ __noinline__ void finish_head(const Tail& t, int x, int nparts, int tid, int bar_id, StoreOut& store_out) {
  static_assert(NT % D == 0, "thread count must be a multiple of head_dim");
  constexpr int NG = NT / D;
  __threadfence();
  named_bar_sync(bar_id, NT);
  if (tid == 0) t.s_misc[0] = (int)atomicAdd(&t.tickets[x], 1u);
  named_bar_sync(bar_id, NT);
  if (t.s_misc[0] != nparts - 1) return;
  __threadfence();
  if (tid == 0) t.tickets[x] = 0;
  const float* parts = t.part + (size_t)x * t.max_parts * (size_t)(t.R * (D + 4));
  const int world = t.comm->world;
  const int d = tid % D;
  for (int r = tid / D; r < t.rows_valid; r += NG) {
    float o_norm, lse2;
    merge_parts<D>(parts, nparts, t.R, r, d, o_norm, lse2);
    if (world == 1) store_out(x, r, d, o_norm, lse2);
    else if (!t.comm->skip_publish) publish<D>(t, x, r, d, o_norm, lse2);
  }
  if (world > 1) {
    if (tid == 0) {
      t.stamps[1] = globaltimer_ns();
      const int n = t.s_misc[1];
      if (n < t.max_pending) { t.pending[n] = x; t.s_misc[1] = n + 1; }
      else t.s_misc[2] = x + 1;  // list full: combine inline below
    }
    named_bar_sync(bar_id, NT);
    if (t.s_misc[2] != 0) {
      named_bar_sync(bar_id, NT);
      if (tid == 0) t.s_misc[2] = 0;
      combine_ranks<D, NT, RB>(t, x, tid, bar_id, store_out);
    }
  }
}

// after the last tile: deferred cross-GPU merges of the heads this CTA finished, then the end-of-kernel arrival (the
// last CTA to leave bumps the device-resident epoch => the kernel is CUDA-graph replayable)
template <int D, int NT, int RB, typename StoreOut>
__device__ __forceinline__ void drain_and_exit(const Tail& t, int tid, int bar_id, StoreOut& store_out) {
  if (t.comm->world <= 1) return;
  named_bar_sync(bar_id, NT);
  const int n = t.s_misc[1];
  for (int i = 0; i < n; ++i) combine_ranks<D, NT, RB>(t, t.pending[i], tid, bar_id, store_out);
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(&t.tickets[t.BH], 1u);
    if (done == gridDim.x - 1) {
      t.tickets[t.BH] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(t.comm->epoch) = t.epoch;
    }
  }
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
