// Shared pieces of the tcgen05 forward kernel (attn_fwd_sm100.cu: M = 128 per CTA): parameters, the fused-mode merge CTA
// and the epoch bookkeeping.
// Reference: merge_item() is the combine of /root/reference/model.py:103-124 (global max, exp-rescaled numerator / denominator
// sums, divide) applied to one 128-row tile across ranks, in fp32 and in fixed rank order.
#pragma once
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace fwd_detail {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kSoftmaxThreads = 128;
constexpr int kFwdThreads = 192;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays <= 2^8 before the max is refreshed

struct FwdParams {
  float* lse;        // (B, Hq, Sq) natural log
  void* out;         // used only by the "no visible keys" path (plain stores)
  long long o_sb, o_sh, o_ss;
  int B, Hq, Hkv, G, Sq, S;
  float scale_log2;
  int causal;
  long long q_pos0, kv_pos0;
  int seg_len;         // two-segment shard (zigzag): local rows >= seg_len sit seg_gap positions further on; INT_MAX = one segment
  long long seg_gap;
  int num_m_tiles;
  int n_items;   // B * Hq * num_m_tiles
  int lag;       // merge CTA of item i is dispatched about `lag` compute CTAs after item i
  int q_in_tmem; // attn_fwd_kernel: keep the query tile in TMEM (QK^T as a TS MMA: half the smem operand traffic)
  // fused multi-GPU combine = reduce-scatter (+ all-gather) over the query tiles:
  //   every 128-row query tile has an OWNER rank (contiguous blocks of `tiles_per_rank` tiles along Sq); a compute CTA pushes
  //   its partial tile (o in the I/O dtype + fp32 lse) to the owner only; the owner's merge CTA combines the W partials;
  //   mode 1 (replicated output): the owner then pushes the FINAL tile to every peer, whose merge CTA copies it out;
  //   mode 2 (sharded output):    the owner writes its rows of a (B, Hq, Sq / W, D) output, nothing else moves.
  // NVLink bytes per rank: (W-1)/W |O| (mode 2) or 2 (W-1)/W |O| (mode 1) instead of the (W-1) |O| of an all-to-all push.
  int mode;
  int tiles_per_rank;   // query tiles owned per rank and (b, head)
  int n_local;          // B * Hq * tiles_per_rank: partial slots per source rank on the owner
  long long final_off;  // byte offset of the final-tile area in the data region (mode 1)
  int final_flag_off;   // u32 offset of the final flags in the flag region (mode 1)
  int sq_out;           // rows of the output tensor per (b, head): Sq (mode 1) or the owner's shard (mode 2)
  CommCtx comm;  // world == 1: unused
};

template <int D>
struct FwdSmem {
  static constexpr int kStages = 3;
  static constexpr int kAtoms = D / 64;
  static constexpr int kQBytes = kBlockM * D * 2;
  static constexpr int kKVBytes = kBlockN * D * 2;
  static constexpr int kAtomBytes = 128 * 128;  // 128 rows x 128 B
  static constexpr size_t kTotal = 1024 + size_t(kQBytes) + size_t(2 * kStages) * kKVBytes + 256;
};

__device__ __forceinline__ float neg_inf_f() { return __int_as_float(0xff800000); }

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (BF16) return pack_bf16x2(lo, hi);
  else return pack_f16x2(lo, hi);
}

// slot addressing of the reduce-scatter combine (see FwdParams)
template <int D>
struct FwdSlots {
  static constexpr int kSlotBytes = kBlockM * D * 2 + kBlockM * 4;
  // partial slot of (source rank, local item) in the OWNER's data region; mode 2 double-buffers by launch parity (the
  // owner may still be merging launch e while a fast rank already pushes launch e + 1), mode 1 needs none: a rank leaves
  // launch e only after it has received EVERY final tile, i.e. after every owner has finished reading its partial slots
  __device__ static size_t part_off(const FwdParams& p, uint32_t epoch, int src, int litem) {
    const int par = p.mode == 2 ? (int)(epoch & 1) : 0;
    return ((size_t)(par * p.comm.world + src) * p.n_local + litem) * kSlotBytes;
  }
  __device__ static size_t part_flag(const FwdParams& p, uint32_t epoch, int src, int litem) {
    const int par = p.mode == 2 ? (int)(epoch & 1) : 0;
    return (size_t)(par * p.comm.world + src) * p.n_local + litem;
  }
};

// Merge CTA (fused multi-GPU mode).  Owner of the tile: all W partial tiles of `item` have been pushed into THIS rank's
// symmetric buffer by the compute CTAs of every rank; merge them in rank order (deterministic), write the final tile and,
// in replicated mode, push it to every peer.  Non-owner in replicated mode: wait for the owner's final tile, copy it out.
template <int D, bool BF16>
__device__ __forceinline__ void merge_item(const FwdParams& p, int item, uint32_t epoch, uint8_t* smem) {
  using SL = FwdSlots<D>;
  constexpr int kSlotBytes = SL::kSlotBytes;
  constexpr int CPR = D / 8;  // 16-byte chunks per row
  const int tid = threadIdx.x;
  const int world = p.comm.world, rank = p.comm.rank;
  const int mi = item % p.num_m_tiles;
  const int bh = item / p.num_m_tiles;
  const int hq = bh % p.Hq, b = bh / p.Hq;
  const int m_tile = p.num_m_tiles - 1 - mi;
  const int m0 = m_tile * kBlockM;
  const int owner = min(m_tile / p.tiles_per_rank, world - 1);
  const int litem = bh * p.tiles_per_rank + (m_tile - owner * p.tiles_per_rank);
  const int row0_out = p.mode == 2 ? m0 - owner * p.tiles_per_rank * kBlockM : m0;   // first row of the tile in `out`
  float* w_s = reinterpret_cast<float*>(smem);           // [128][world]
  int* ok_s = reinterpret_cast<int*>(w_s + kBlockM * kMaxWorld);
  if (tid == 0) *ok_s = 1;
  __syncthreads();
  uint8_t* my_data = reinterpret_cast<uint8_t*>(p.comm.data[rank]);
  if (rank != owner) {
    if (p.mode != 1) return;
    // ---- replicated output, non-owner: the final tile arrives in my final slot `item`
    if (tid == 0) {
      const uint32_t* f = p.comm.flags[rank] + p.final_flag_off + item;
      if (!spin_flag_acquire(f, epoch, p.comm.timeout_ns)) {
        p.comm.status[0] = kCommTimeout; p.comm.status[1] = item; p.comm.status[2] = owner; p.comm.status[3] = epoch;
        *ok_s = 0;
      }
    }
    __syncthreads();
    const bool ok = *ok_s != 0;
    const uint8_t* fs = my_data + p.final_off + (size_t)item * kSlotBytes;
    for (int c = tid; c < kBlockM * CPR; c += (int)blockDim.x) {
      const int row = c / CPR, ch = c - row * CPR;
      if (m0 + row >= p.Sq) continue;
      float4 raw = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(fs + (size_t)row * D * 2 + ch * 16));
      if (!ok) raw = make_float4(__int_as_float(0x7fc07fc0), __int_as_float(0x7fc07fc0), __int_as_float(0x7fc07fc0), __int_as_float(0x7fc07fc0));
      uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)hq * p.o_sh +
                     (long long)(m0 + row) * p.o_ss + ch * 8;
      *reinterpret_cast<float4*>(op) = raw;
    }
    if (tid < kBlockM && m0 + tid < p.Sq) {
      float l = ld_relaxed_sys_f(reinterpret_cast<const float*>(fs + kBlockM * D * 2) + tid);
      if (!ok) l = __int_as_float(0x7fc00000);
      p.lse[((long long)b * p.Hq + hq) * p.sq_out + m0 + tid] = l;
    }
    return;
  }
  // ---- owner: acquire the W partial flags, merge in rank order
  if (tid < world) {
    const uint32_t* f = p.comm.flags[rank] + SL::part_flag(p, epoch, tid, litem);
    if (!spin_flag_acquire(f, epoch, p.comm.timeout_ns)) {
      p.comm.status[0] = kCommTimeout; p.comm.status[1] = item; p.comm.status[2] = tid; p.comm.status[3] = epoch;
      *ok_s = 0;
    }
  }
  __syncthreads();
  const bool ok = *ok_s != 0;
  float lse_row = neg_inf_f();
  if (tid < kBlockM) {
    const int row = tid;
    float ls[kMaxWorld];
    float mx = neg_inf_f();
#pragma unroll
    for (int s = 0; s < kMaxWorld; ++s) {
      ls[s] = neg_inf_f();
      if (s < world) {
        ls[s] = ld_relaxed_sys_f(reinterpret_cast<const float*>(my_data + SL::part_off(p, epoch, s, litem) + kBlockM * D * 2) + row);
        mx = fmaxf(mx, ls[s]);
      }
    }
    const float ms = mx == neg_inf_f() ? 0.f : mx;
    float den = 0.f;
#pragma unroll
    for (int s = 0; s < kMaxWorld; ++s) {
      if (s < world) {
        const float w = fast_exp2((ls[s] - ms) * 1.4426950408889634f);
        w_s[row * kMaxWorld + s] = w;
        den += w;
      }
    }
    const float inv = den > 0.f ? 1.f / den : 0.f;
    for (int s = 0; s < world; ++s) w_s[row * kMaxWorld + s] *= inv;
    lse_row = den > 0.f ? ms + fast_log2(den) * 0.6931471805599453f : neg_inf_f();
    if (!ok) lse_row = __int_as_float(0x7fc00000);
    if (m0 + row < p.Sq) p.lse[((long long)b * p.Hq + hq) * p.sq_out + row0_out + row] = lse_row;
  }
  __syncthreads();
  for (int c = tid; c < kBlockM * CPR; c += (int)blockDim.x) {
    const int row = c / CPR, ch = c - row * CPR;
    if (m0 + row >= p.Sq) continue;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float4 raw[8];
    for (int s0 = 0; s0 < world; s0 += 8) {   // all loads of a batch in flight before the first use
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s0 + s < world)
          raw[s] = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(my_data + SL::part_off(p, epoch, s0 + s, litem) + (size_t)row * D * 2 + ch * 16));
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (s0 + s < world) {
          const uint32_t w4[4] = {__float_as_uint(raw[s].x), __float_as_uint(raw[s].y), __float_as_uint(raw[s].z), __float_as_uint(raw[s].w)};
          const float w = w_s[row * kMaxWorld + s0 + s];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float lo, hi;
            if constexpr (BF16) { lo = bf16lo(w4[i]); hi = bf16hi(w4[i]); } else { lo = f16lo(w4[i]); hi = f16hi(w4[i]); }
            acc[2 * i] = fmaf(w, lo, acc[2 * i]);
            acc[2 * i + 1] = fmaf(w, hi, acc[2 * i + 1]);
          }
        }
      }
    }
    if (!ok) { for (int i = 0; i < 8; ++i) acc[i] = __int_as_float(0x7fc00000); }
    uint4 o;
    o.x = pack2<BF16>(acc[0], acc[1]); o.y = pack2<BF16>(acc[2], acc[3]);
    o.z = pack2<BF16>(acc[4], acc[5]); o.w = pack2<BF16>(acc[6], acc[7]);
    uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)hq * p.o_sh +
                   (long long)(row0_out + row) * p.o_ss + ch * 8;
    *reinterpret_cast<uint4*>(op) = o;
    if (p.mode == 1 && !p.comm.skip_publish) {   // all-gather: the final tile to every peer's final slot
      for (int d = 0; d < world; ++d) {
        if (d == rank) continue;
        uint8_t* fs = reinterpret_cast<uint8_t*>(p.comm.data[d]) + p.final_off + (size_t)item * kSlotBytes;
        *reinterpret_cast<uint4*>(fs + (size_t)row * D * 2 + ch * 16) = o;
      }
    }
  }
  if (p.mode == 1 && !p.comm.skip_publish) {
    if (tid < kBlockM) {
      for (int d = 0; d < world; ++d) {
        if (d == rank) continue;
        uint8_t* fs = reinterpret_cast<uint8_t*>(p.comm.data[d]) + p.final_off + (size_t)item * kSlotBytes;
        reinterpret_cast<float*>(fs + kBlockM * D * 2)[tid] = lse_row;
      }
    }
    __syncthreads();
    if (tid < world && tid != rank) { fence_acq_rel_sys(); st_release_sys_u32(p.comm.flags[tid] + p.final_flag_off + item, epoch); }
  }
}

// end-of-kernel arrival; the last CTA bumps the device-resident epoch for the next launch
__device__ __forceinline__ void comm_kernel_exit(const FwdParams& p, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(p.comm.status + 8, 1u);
    if (done == gridDim.x - 1) {
      p.comm.status[8] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(p.comm.epoch) = epoch;
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


}  // namespace fwd_detail
}  // namespace ta
