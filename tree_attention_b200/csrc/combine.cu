// combine_partials -- stand-alone cross-GPU merge of per-rank attention partials (o, lse) over
// symmetric memory.  Replaces model.py:103-124 (three NCCL all-reduces + ~10 elementwise kernels,
// SURVEY.md 2.3 K7-K13/N1-N3) with one launch and no NCCL.  It is the unfused stepping stone and
// the fallback when the partial was not produced by one of the fused attention kernels.
//
//   mode 0  one-shot: every rank pushes its rows into slot [src] of every peer, releases one flag per
//           (row-chunk, src), acquires the W flags of the chunk and merges in rank order.
//           Depth 1 -- optimal on a uniform NVSwitch fabric for KB..MB payloads.
//   mode 1  butterfly ("tree"): log2(W) rounds; in round k a rank exchanges its running partial with
//           rank ^ (1 << k) and both apply merge(lower-rank operand, higher-rank operand), so all
//           ranks finish with bitwise-identical results after depth log2(W).
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace {

constexpr int kCombThreads = 256;
constexpr int kRowsPerCta = 8;

struct CombineParams {
  const float* o_part;
  const float* lse_part;
  void* out;
  float* lse_out;
  int out_dtype;  // 0 fp32, 1 bf16, 2 fp16
  long long rows;
  int D;
  int nchunks;
  int rounds;
  CommCtx comm;
};

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

__device__ __forceinline__ void store_out(const CombineParams& p, long long row, int d, float v) {
  if (p.out_dtype == 0) reinterpret_cast<float*>(p.out)[row * p.D + d] = v;
  else if (p.out_dtype == 1) reinterpret_cast<__nv_bfloat16*>(p.out)[row * p.D + d] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p.out)[row * p.D + d] = __float2half_rn(v);
}

__device__ __forceinline__ void report_timeout(const CommCtx& c, int item, int src, uint32_t epoch) {
  c.status[0] = kCommTimeout; c.status[1] = item; c.status[2] = src; c.status[3] = epoch;
}

__global__ void __launch_bounds__(kCombThreads) combine_oneshot_kernel(const CombineParams p) {
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  const int world = p.comm.world, rank = p.comm.rank;
  const uint32_t epoch = ld_relaxed_sys_u32(p.comm.epoch) + 1;
  const int parity = epoch & 1;
  const int D = p.D, RS = D + 4;
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
    const long long row0 = (long long)chunk * kRowsPerCta;
    const int nrows = (int)min((long long)kRowsPerCta, p.rows - row0);
    if (tid == 0) s_ok = 1;
    // publish my rows to every rank (own slot included)
    if (!p.comm.skip_publish) {
      for (int idx = tid; idx < nrows * RS; idx += kCombThreads) {
        const int r = idx / RS, d = idx - r * RS;
        float v = 0.f;
        if (d < D) v = p.o_part[(row0 + r) * D + d];
        else if (d == D) v = p.lse_part[row0 + r] * LOG2E;
        for (int dst = 0; dst < world; ++dst) {
          float* sp = p.comm.data[dst] + ((size_t)(parity * world + rank) * p.rows + row0 + r) * RS;
          sp[d] = v;
        }
      }
    }
    __syncthreads();
    if (tid < world) {
      if (!p.comm.skip_publish) {
        fence_acq_rel_sys();
        st_release_sys_u32(p.comm.flags[tid] + (size_t)(parity * world + rank) * p.nchunks + chunk, epoch);
      }
      const uint32_t* f = p.comm.flags[rank] + (size_t)(parity * world + tid) * p.nchunks + chunk;
      if (!spin_flag_acquire(f, epoch, p.comm.timeout_ns)) { report_timeout(p.comm, chunk, tid, epoch); s_ok = 0; }
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    for (int idx = tid; idx < nrows * D; idx += kCombThreads) {
      const int r = idx / D, d = idx - r * D;
      float mx = neg_inf();
      for (int s = 0; s < world; ++s)
        mx = fmaxf(mx, ld_relaxed_sys_f(p.comm.data[rank] + ((size_t)(parity * world + s) * p.rows + row0 + r) * RS + D));
      const float ms = mx == neg_inf() ? 0.f : mx;
      float num = 0.f, den = 0.f;
      for (int s = 0; s < world; ++s) {
        const float* sp = p.comm.data[rank] + ((size_t)(parity * world + s) * p.rows + row0 + r) * RS;
        const float w = fast_exp2(ld_relaxed_sys_f(sp + D) - ms);
        num = fmaf(w, ld_relaxed_sys_f(sp + d), num);
        den += w;
      }
      float o = den > 0.f ? num / den : 0.f;
      float l = den > 0.f ? (ms + fast_log2(den)) * LN2 : neg_inf();
      if (!ok) { o = __int_as_float(0x7fc00000); l = o; }
      store_out(p, row0 + r, d, o);
      if (d == 0 && p.lse_out) p.lse_out[row0 + r] = l;
    }
    __syncthreads();
  }
  // epoch bump by the last CTA
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(p.comm.status + 8, 1u);
    if (done == gridDim.x - 1) {
      p.comm.status[8] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(p.comm.epoch) = epoch;
    }
  }
}

// Butterfly: the running partial lives in shared memory; slot [parity][round] of the PARTNER receives it.
__global__ void __launch_bounds__(kCombThreads) combine_butterfly_kernel(const CombineParams p) {
  extern __shared__ float cur[];  // [kRowsPerCta][D+4]
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  const int rank = p.comm.rank;
  const uint32_t epoch = ld_relaxed_sys_u32(p.comm.epoch) + 1;
  const int parity = epoch & 1;
  const int D = p.D, RS = D + 4;
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
    const long long row0 = (long long)chunk * kRowsPerCta;
    const int nrows = (int)min((long long)kRowsPerCta, p.rows - row0);
    if (tid == 0) s_ok = 1;
    for (int idx = tid; idx < nrows * RS; idx += kCombThreads) {
      const int r = idx / RS, d = idx - r * RS;
      float v = 0.f;
      if (d < D) v = p.o_part[(row0 + r) * D + d];
      else if (d == D) v = p.lse_part[row0 + r] * LOG2E;
      cur[idx] = v;
    }
    __syncthreads();
    for (int k = 0; k < p.rounds; ++k) {
      const int partner = rank ^ (1 << k);
      float* dst = p.comm.data[partner] + ((size_t)(parity * p.rounds + k) * p.rows + row0) * RS;
      if (!p.comm.skip_publish)
        for (int idx = tid; idx < nrows * RS; idx += kCombThreads) dst[idx] = cur[idx];
      __syncthreads();
      if (tid == 0) {
        if (!p.comm.skip_publish) {
          fence_acq_rel_sys();
          st_release_sys_u32(p.comm.flags[partner] + (size_t)(parity * p.rounds + k) * p.nchunks + chunk, epoch);
        }
        const uint32_t* f = p.comm.flags[rank] + (size_t)(parity * p.rounds + k) * p.nchunks + chunk;
        if (!spin_flag_acquire(f, epoch, p.comm.timeout_ns)) { report_timeout(p.comm, chunk, partner, epoch); s_ok = 0; }
      }
      __syncthreads();
      const float* in = p.comm.data[rank] + ((size_t)(parity * p.rounds + k) * p.rows + row0) * RS;
      const bool i_am_low = rank < partner;
      // merge(low, high): identical arithmetic on both partners => identical bits
      float newv[(kRowsPerCta * 260 + kCombThreads - 1) / kCombThreads];
      int cnt = 0;
      for (int idx = tid; idx < nrows * RS; idx += kCombThreads, ++cnt) {
        const int r = idx / RS, d = idx - r * RS;
        const float l_mine = cur[r * RS + D], l_theirs = ld_relaxed_sys_f(in + r * RS + D);
        const float l_lo = i_am_low ? l_mine : l_theirs, l_hi = i_am_low ? l_theirs : l_mine;
        const float mx = fmaxf(l_lo, l_hi);
        const float ms = mx == neg_inf() ? 0.f : mx;
        const float w_lo = fast_exp2(l_lo - ms), w_hi = fast_exp2(l_hi - ms);
        const float den = w_lo + w_hi;
        float v;
        if (d < D) {
          const float o_mine = cur[idx], o_theirs = ld_relaxed_sys_f(in + idx);
          const float o_lo = i_am_low ? o_mine : o_theirs, o_hi = i_am_low ? o_theirs : o_mine;
          v = den > 0.f ? fmaf(w_lo, o_lo, w_hi * o_hi) / den : 0.f;
        } else if (d == D) {
          v = den > 0.f ? ms + fast_log2(den) : neg_inf();
        } else {
          v = 0.f;
        }
        newv[cnt] = v;
      }
      __syncthreads();
      cnt = 0;
      for (int idx = tid; idx < nrows * RS; idx += kCombThreads, ++cnt) cur[idx] = newv[cnt];
      __syncthreads();
    }
    const bool ok = s_ok != 0;
    for (int idx = tid; idx < nrows * D; idx += kCombThreads) {
      const int r = idx / D, d = idx - r * D;
      float o = cur[r * RS + d];
      if (!ok) o = __int_as_float(0x7fc00000);
      store_out(p, row0 + r, d, o);
      if (d == 0 && p.lse_out) p.lse_out[row0 + r] = ok ? cur[r * RS + D] * LN2 : o;
    }
    __syncthreads();
  }
  if (tid == 0) {
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

}  // namespace

void combine_launch(const float* o_part, const float* lse_part, void* out, int out_dtype, float* lse_out,
                    int64_t rows, int D, const CommCtxHost& comm, int mode, cudaStream_t stream) {
  if (comm.world < 2) throw std::runtime_error("combine: world size must be >= 2");
  if (D > 256 || D % 4 != 0) throw std::runtime_error("combine: head_dim must be a multiple of 4, <= 256");
  CombineParams p;
  p.o_part = o_part; p.lse_part = lse_part; p.out = out; p.lse_out = lse_out; p.out_dtype = out_dtype;
  p.rows = rows; p.D = D;
  p.nchunks = (int)((rows + kRowsPerCta - 1) / kRowsPerCta);
  p.comm = to_device_ctx(comm);
  int rounds = 0;
  while ((1 << rounds) < comm.world) ++rounds;
  p.rounds = rounds;
  const size_t RS = D + 4;
  // co-residency: every CTA may spin on peers, so the grid never exceeds what is resident at once
  const int grid = std::min(p.nchunks, 2 * num_sms());
  if (mode == 0) {
    const size_t need = (size_t)2 * comm.world * rows * RS * sizeof(float);
    const size_t needf = (size_t)2 * comm.world * p.nchunks * sizeof(uint32_t);
    if (need > comm.data_bytes || needf > comm.flag_bytes)
      throw std::runtime_error("combine(one-shot): symmetric buffer too small");
    combine_oneshot_kernel<<<grid, kCombThreads, 0, stream>>>(p);
  } else if (mode == 1) {
    if ((1 << rounds) != comm.world) throw std::runtime_error("combine(butterfly): world size must be a power of two");
    const size_t need = (size_t)2 * rounds * rows * RS * sizeof(float);
    const size_t needf = (size_t)2 * rounds * p.nchunks * sizeof(uint32_t);
    if (need > comm.data_bytes || needf > comm.flag_bytes)
      throw std::runtime_error("combine(butterfly): symmetric buffer too small");
    combine_butterfly_kernel<<<grid, kCombThreads, kRowsPerCta * RS * sizeof(float), stream>>>(p);
  } else {
    throw std::runtime_error("combine: unknown mode");
  }
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ta
