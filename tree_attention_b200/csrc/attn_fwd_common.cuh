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
  int num_m_tiles;
  int n_items;   // B * Hq * num_m_tiles
  int lag;       // merge CTA of item i is dispatched about `lag` compute CTAs after item i
  int q_in_tmem; // attn_fwd_kernel: keep the query tile in TMEM (QK^T as a TS MMA: half the smem operand traffic)
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

// Merge CTA (fused multi-GPU mode): all W partial tiles of `item` have been pushed into THIS rank's
// symmetric buffer by the compute CTAs of every rank; merge them in rank order and write the final tile.
template <int D, bool BF16>
__device__ __forceinline__ void merge_item(const FwdParams& p, int item, uint32_t epoch, uint8_t* smem) {
  constexpr int kSlotBytes = kBlockM * D * 2 + kBlockM * 4;
  constexpr int CPR = D / 8;  // 16-byte chunks per row
  const int tid = threadIdx.x;
  const int world = p.comm.world;
  const int parity = epoch & 1;
  const int mi = item % p.num_m_tiles;
  const int bh = item / p.num_m_tiles;
  const int hq = bh % p.Hq, b = bh / p.Hq;
  const int m0 = (p.num_m_tiles - 1 - mi) * kBlockM;
  float* w_s = reinterpret_cast<float*>(smem);           // [128][world]
  int* ok_s = reinterpret_cast<int*>(w_s + kBlockM * kMaxWorld);
  if (tid == 0) *ok_s = 1;
  __syncthreads();
  const uint8_t* base = reinterpret_cast<const uint8_t*>(p.comm.data[p.comm.rank]);
  if (tid < world) {
    const uint32_t* f = p.comm.flags[p.comm.rank] + ((size_t)(parity * world + tid) * p.n_items + item);
    if (!spin_flag_acquire(f, epoch, p.comm.timeout_ns)) {
      p.comm.status[0] = kCommTimeout; p.comm.status[1] = item; p.comm.status[2] = tid; p.comm.status[3] = epoch;
      *ok_s = 0;
    }
  }
  __syncthreads();
  const bool ok = *ok_s != 0;
  if (tid < kBlockM) {
    const int row = tid;
    float mx = neg_inf_f();
    for (int s = 0; s < world; ++s) {
      const float* lp = reinterpret_cast<const float*>(base + ((size_t)(parity * world + s) * p.n_items + item) * kSlotBytes + kBlockM * D * 2);
      mx = fmaxf(mx, ld_relaxed_sys_f(lp + row));
    }
    const float ms = mx == neg_inf_f() ? 0.f : mx;
    float den = 0.f;
    for (int s = 0; s < world; ++s) {
      const float* lp = reinterpret_cast<const float*>(base + ((size_t)(parity * world + s) * p.n_items + item) * kSlotBytes + kBlockM * D * 2);
      const float w = fast_exp2((ld_relaxed_sys_f(lp + row) - ms) * 1.4426950408889634f);
      w_s[row * kMaxWorld + s] = w;
      den += w;
    }
    const float inv = den > 0.f ? 1.f / den : 0.f;
    for (int s = 0; s < world; ++s) w_s[row * kMaxWorld + s] *= inv;
    if (m0 + row < p.Sq) {
      float l = den > 0.f ? ms + fast_log2(den) * 0.6931471805599453f : neg_inf_f();
      if (!ok) l = __int_as_float(0x7fc00000);
      p.lse[((long long)b * p.Hq + hq) * p.Sq + m0 + row] = l;
    }
  }
  __syncthreads();
  for (int c = tid; c < kBlockM * CPR; c += (int)blockDim.x) {
    const int row = c / CPR, ch = c - row * CPR;
    if (m0 + row >= p.Sq) continue;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < world; ++s) {
      const uint8_t* sp = base + ((size_t)(parity * world + s) * p.n_items + item) * kSlotBytes + (size_t)row * D * 2 + ch * 16;
      const float4 raw = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(sp));
      const uint32_t w4[4] = {__float_as_uint(raw.x), __float_as_uint(raw.y), __float_as_uint(raw.z), __float_as_uint(raw.w)};
      const float w = w_s[row * kMaxWorld + s];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        if constexpr (BF16) { lo = bf16lo(w4[i]); hi = bf16hi(w4[i]); } else { lo = f16lo(w4[i]); hi = f16hi(w4[i]); }
        acc[2 * i] = fmaf(w, lo, acc[2 * i]);
        acc[2 * i + 1] = fmaf(w, hi, acc[2 * i + 1]);
      }
    }
    if (!ok) { for (int i = 0; i < 8; ++i) acc[i] = __int_as_float(0x7fc00000); }
    uint4 o;
    o.x = pack2<BF16>(acc[0], acc[1]); o.y = pack2<BF16>(acc[2], acc[3]);
    o.z = pack2<BF16>(acc[4], acc[5]); o.w = pack2<BF16>(acc[6], acc[7]);
    uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)b * p.o_sb + (long long)hq * p.o_sh +
                   (long long)(m0 + row) * p.o_ss + ch * 8;
    *reinterpret_cast<uint4*>(op) = o;
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
