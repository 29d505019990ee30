// symm_allreduce_sum -- fp32 sum over ranks on symmetric memory, one launch, no NCCL.  This is the backward
// pass's partial-gradient tree reduce (dQ = sum_r dQ_r): the same publish / acquire pattern
// as the forward combine, with the two-shot schedule that suits GB-scale payloads:
//
//   phase A  copy my partial into my symmetric stage (skipped when the producer kernel already wrote there),
//            release one "chunk ready" flag per 128 KB chunk on every peer;
//   phase B  reduce-scatter by PULL: rank r owns slice r; for each of its chunks it acquires the W ready flags,
//            reads the W copies with 16-byte P2P loads (NVLink), sums them in rank order (deterministic) and
//            PUSHES the result chunk into every rank's result area, then releases a "result ready" flag;
//   phase C  every rank acquires the result flags of all chunks, so that when the kernel retires the
//            replicated sum is complete in local memory.
// Per rank NVLink traffic: (W-1)/W * n in + (W-1)/W * n out.  Slots are double-buffered by epoch parity.
// Reference: the reference sums across ranks with blocking ncclAllReduce (/root/reference/model.py:108-115); it has no backward.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace {

constexpr int kRedThreads = 256;
constexpr int kChunkFloats = 32768;  // 128 KB per flag: a system-scope fence + release per chunk must be amortised over enough bytes
                                      // (16 KB chunks: 7.1 ms for 1 GiB at W = 2 against NCCL's 1.9 ms -- the fences, not the wires)

struct ReduceParams {
  const float* x;      // local input (nullptr: already staged)
  float* y;            // local output: the replicated sum
  long long n;         // floats (multiple of 4)
  long long n_pad;     // per-parity area size in floats
  int nchunks;         // total chunks over the whole vector
  CommCtx comm;
};

// data layout per rank: [parity 2][stage n_pad | result n_pad]; flags: [parity 2][kind 2][src W][nchunks]
__device__ __forceinline__ float* stage_ptr(const ReduceParams& p, int rank, int parity) {
  return p.comm.data[rank] + (size_t)parity * 2 * p.n_pad;
}
__device__ __forceinline__ float* result_ptr(const ReduceParams& p, int rank, int parity) {
  return p.comm.data[rank] + (size_t)parity * 2 * p.n_pad + p.n_pad;
}
__device__ __forceinline__ uint32_t* flag_ptr(const ReduceParams& p, int rank, int parity, int kind, int src, int chunk) {
  return p.comm.flags[rank] + (((size_t)(parity * 2 + kind) * p.comm.world + src) * p.nchunks + chunk);
}

__global__ void __launch_bounds__(kRedThreads, 2) symm_allreduce_kernel(const ReduceParams p) {
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  const int world = p.comm.world, rank = p.comm.rank;
  const uint32_t epoch = ld_relaxed_sys_u32(p.comm.epoch) + 1;
  const int parity = epoch & 1;
  if (tid == 0) s_ok = 1;
  __syncthreads();

  // ---- phase A: stage + publish readiness
  for (int c = blockIdx.x; c < p.nchunks; c += gridDim.x) {
    const long long lo = (long long)c * kChunkFloats, hi = min(lo + kChunkFloats, p.n);
    if (p.x != nullptr) {
      float4* dst = reinterpret_cast<float4*>(stage_ptr(p, rank, parity) + lo);
      const float4* src = reinterpret_cast<const float4*>(p.x + lo);
      for (int i = tid; i < (hi - lo) / 4; i += kRedThreads) dst[i] = src[i];
    }
    __syncthreads();
    if (tid < world && !p.comm.skip_publish) {
      fence_acq_rel_sys();
      st_release_sys_u32(flag_ptr(p, tid, parity, 0, rank, c), epoch);
    }
  }
  // ---- phase B: reduce my slice (chunks c with c % world == rank), push the result everywhere
  for (int c = rank + blockIdx.x * world; c < p.nchunks; c += gridDim.x * world) {
    const long long lo = (long long)c * kChunkFloats, hi = min(lo + kChunkFloats, p.n);
    if (tid < world) {
      if (!spin_flag_acquire(flag_ptr(p, rank, parity, 0, tid, c), epoch, p.comm.timeout_ns)) {
        p.comm.status[0] = kCommTimeout; p.comm.status[1] = c; p.comm.status[2] = tid; p.comm.status[3] = epoch;
        s_ok = 0;
      }
    }
    __syncthreads();
    // Every thread owns kPerThread float4 of the chunk; ALL their W remote loads are issued before the first add
    // (round 1 issued load, add, load, add: W serialized NVLink round trips per element -- 214 GB/s bus bandwidth against
    // NCCL's 718 at 1 GiB).  The sum runs in rank order (deterministic, identical on every rank).
    constexpr int kPerThread = kChunkFloats / 4 / kRedThreads;   // 32 float4 per thread and chunk, two at a time
    const int n4 = (int)((hi - lo) / 4);
#pragma unroll 1
    for (int u0 = 0; u0 < kPerThread; u0 += 2) {
      float4 acc[2];
      acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s0 = 0; s0 < world; s0 += 8) {
        float4 v[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = tid + (u0 + u) * kRedThreads;
#pragma unroll
          for (int s = 0; s < 8; ++s)
            if (i < n4 && s0 + s < world)
              v[u][s] = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(stage_ptr(p, s0 + s, parity) + lo) + i);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = tid + (u0 + u) * kRedThreads;
#pragma unroll
          for (int s = 0; s < 8; ++s)
            if (i < n4 && s0 + s < world) { acc[u].x += v[u][s].x; acc[u].y += v[u][s].y; acc[u].z += v[u][s].z; acc[u].w += v[u][s].w; }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = tid + (u0 + u) * kRedThreads;
        if (i < n4)
          for (int d = 0; d < world; ++d) reinterpret_cast<float4*>(result_ptr(p, d, parity) + lo)[i] = acc[u];
      }
    }
    __syncthreads();
    if (tid < world && !p.comm.skip_publish) {
      fence_acq_rel_sys();
      st_release_sys_u32(flag_ptr(p, tid, parity, 1, rank, c), epoch);
    }
  }
  // ---- phase C: acquire every result chunk and copy it to the caller's output, so that the replicated sum
  // is complete in ordinary local memory when the kernel retires
  for (int c = blockIdx.x; c < p.nchunks; c += gridDim.x) {
    const long long lo = (long long)c * kChunkFloats, hi = min(lo + kChunkFloats, p.n);
    const int owner = c % world;
    if (tid == 0) {
      if (!spin_flag_acquire(flag_ptr(p, rank, parity, 1, owner, c), epoch, p.comm.timeout_ns)) {
        p.comm.status[0] = kCommTimeout; p.comm.status[1] = c; p.comm.status[2] = owner; p.comm.status[3] = epoch;
        s_ok = 0;
      }
    }
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(result_ptr(p, rank, parity) + lo);
    float4* dst = reinterpret_cast<float4*>(p.y + lo);
    const bool ok = s_ok != 0;
    {
      constexpr int kPerThread = kChunkFloats / 4 / kRedThreads;
      const int n4 = (int)((hi - lo) / 4);
#pragma unroll 1
      for (int u0 = 0; u0 < kPerThread; u0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (tid + (u0 + u) * kRedThreads < n4) v[u] = ld_relaxed_sys_f4(src + tid + (u0 + u) * kRedThreads);   // all loads in flight
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (tid + (u0 + u) * kRedThreads < n4) {
            if (!ok) v[u] = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
            dst[tid + (u0 + u) * kRedThreads] = v[u];
          }
        }
      }
    }
  }
  __syncthreads();
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

void symm_allreduce_sizes(int64_t n, int world, size_t* data_bytes, size_t* flag_bytes) {
  const int64_t n_pad = (n + kChunkFloats - 1) / kChunkFloats * kChunkFloats;
  const int64_t nchunks = n_pad / kChunkFloats;
  *data_bytes = (size_t)4 * n_pad * sizeof(float);
  *flag_bytes = (size_t)4 * world * nchunks * sizeof(uint32_t);
}

// y = sum over ranks of x (fp32, n % 4 == 0); y may alias x.
void symm_allreduce_launch(const float* x, float* y, int64_t n, const CommCtxHost& comm, cudaStream_t stream) {
  if (comm.world < 2) throw std::runtime_error("symm_allreduce: world size must be >= 2");
  if (n % 4 != 0) throw std::runtime_error("symm_allreduce: element count must be a multiple of 4");
  size_t db, fb;
  symm_allreduce_sizes(n, comm.world, &db, &fb);
  if (db > comm.data_bytes || fb > comm.flag_bytes) throw std::runtime_error("symm_allreduce: symmetric buffer too small");
  ReduceParams p;
  p.x = x; p.y = y; p.n = n;
  p.n_pad = (n + kChunkFloats - 1) / kChunkFloats * kChunkFloats;
  p.nchunks = (int)(p.n_pad / kChunkFloats);
  p.comm = to_device_ctx(comm);
  const int grid = std::min(p.nchunks, 2 * num_sms());
  symm_allreduce_kernel<<<grid, kRedThreads, 0, stream>>>(p);
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ta
