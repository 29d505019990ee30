// quant_mxfp8 -- OCP-MX style block-scaled fp8: e4m3 elements with one UE8M0 (power-of-two) scale per 32
// consecutive elements of the innermost (head) dimension.  Used for the block-scaled-fp8 KV cache
// (BASELINE.json config "256K, block-scaled fp8 forward").  The scale is the smallest power of two that maps
// the block's amax into the e4m3 range (|x| <= 448), so quantisation never saturates.
// Reference: none -- /root/reference/model.py is fp16 only (model.py:51-53); fp8 KV caches are BASELINE.json's fp8 config.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace {

template <int IN>  // 0 = bf16, 1 = fp16, 2 = fp32
__device__ __forceinline__ float load_in(const void* p, long long i) {
  if constexpr (IN == 0) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  else if constexpr (IN == 1) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  else return reinterpret_cast<const float*>(p)[i];
}

template <int IN>
__global__ void quant_mxfp8_kernel(const void* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sc,
                                   long long nblocks) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= nblocks) return;
  float v[32];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { v[i] = load_in<IN>(x, blk * 32 + i); amax = fmaxf(amax, fabsf(v[i])); }
  int e = 0;
  if (amax > 0.f && isfinite(amax)) {
    int ex;
    const float m = frexpf(amax / 448.f, &ex);
    e = (m == 0.5f) ? ex - 1 : ex;  // ceil(log2(amax / 448))
    e = max(-127, min(127, e));
  }
  const float inv = exp2f((float)-e);
  sc[blk] = (uint8_t)(e + 127);
  uint32_t out[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(v[4 * i + 1] * inv), "f"(v[4 * i + 0] * inv));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(v[4 * i + 3] * inv), "f"(v[4 * i + 2] * inv));
    out[i] = uint32_t(lo) | (uint32_t(hi) << 16);
  }
  uint4* qp = reinterpret_cast<uint4*>(q + blk * 32);
  qp[0] = make_uint4(out[0], out[1], out[2], out[3]);
  qp[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

// Sequence-blocked MX (the V-cache layout of the block-scaled tensor-core decode): one scale per 32 KEYS per channel;
// the four scale bytes of a (128-key tile, channel) pair are stored as one word.  x / q: (BH, S, D), sc: (BH, T, D, 4).
// One CTA per (bh, 32-key block), one thread per channel: loads and stores are coalesced across the channels.
template <int IN>
__global__ void quant_mxfp8_seq_kernel(const void* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sc,
                                       int S, int D, int T) {
  const int g = blockIdx.x % (T * 4);          // 32-key block within the (padded) sequence
  const long long bh = blockIdx.x / (T * 4);
  const int d = threadIdx.x;
  float v[32];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int key = g * 32 + i;
    v[i] = key < S ? load_in<IN>(x, (bh * S + key) * D + d) : 0.f;
    amax = fmaxf(amax, fabsf(v[i]));
  }
  int e = 0;
  if (amax > 0.f && isfinite(amax)) {
    int ex;
    const float m = frexpf(amax / 448.f, &ex);
    e = (m == 0.5f) ? ex - 1 : ex;  // ceil(log2(amax / 448))
    e = max(-127, min(127, e));
  }
  const float inv = exp2f((float)-e);
  sc[((bh * T + g / 4) * D + d) * 4 + (g & 3)] = (uint8_t)(e + 127);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int key = g * 32 + i;
    if (key < S) {
      uint16_t b2;
      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(b2) : "f"(0.f), "f"(v[i] * inv));
      q[(bh * S + key) * D + d] = (uint8_t)(b2 & 0xff);
    }
  }
}

// KV append into a sequence-blocked MX cache: rows [pos, pos + n) of every (b, h) are replaced by x (BH, n, D).  A 32-key
// block shares one scale per channel, so every block the new rows touch is re-quantised as a whole: old rows are
// de-quantised with the old scale, the new rows come from x, a new power-of-two scale is chosen and all 32 rows are
// written back.  One CTA per (bh, touched block), one thread per channel (coalesced across channels).
template <int IN>
__global__ void mxfp8_seq_append_kernel(const void* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sc,
                                        int S, int D, int T, int pos, int n, int g0, int nblk) {
  const int g = g0 + blockIdx.x % nblk;        // 32-key block of the cache
  const long long bh = blockIdx.x / nblk;
  const int d = threadIdx.x;
  uint8_t* scp = sc + ((bh * T + g / 4) * D + d) * 4 + (g & 3);
  const float s_old = exp2f((float)((int)*scp - 127));
  float v[32];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int key = g * 32 + i;
    float val = 0.f;
    if (key < S) {
      if (key >= pos && key < pos + n) {
        val = load_in<IN>(x, (bh * n + (key - pos)) * D + d);
      } else {
        uint32_t h2;
        asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"((uint16_t)q[(bh * S + key) * D + d]));
        val = f16lo(h2) * s_old;
      }
    }
    v[i] = val;
    amax = fmaxf(amax, fabsf(val));
  }
  int e = 0;
  if (amax > 0.f && isfinite(amax)) {
    int ex;
    const float m = frexpf(amax / 448.f, &ex);
    e = (m == 0.5f) ? ex - 1 : ex;  // ceil(log2(amax / 448))
    e = max(-127, min(127, e));
  }
  const float inv = exp2f((float)-e);
  *scp = (uint8_t)(e + 127);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int key = g * 32 + i;
    if (key < S) {
      uint16_t b2;
      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(b2) : "f"(0.f), "f"(v[i] * inv));
      q[(bh * S + key) * D + d] = (uint8_t)(b2 & 0xff);
    }
  }
}

__global__ void dequant_mxfp8_kernel(const uint8_t* __restrict__ q, const uint8_t* __restrict__ sc, float* __restrict__ y,
                                     long long nblocks) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= nblocks) return;
  const float s = exp2f((float)((int)sc[blk] - 127));
  const uint16_t* qp = reinterpret_cast<const uint16_t*>(q + blk * 32);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint32_t h2;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"(qp[i]));
    y[blk * 32 + 2 * i] = f16lo(h2) * s;
    y[blk * 32 + 2 * i + 1] = f16hi(h2) * s;
  }
}

}  // namespace

void quant_mxfp8_launch(const void* x, int in_dtype, uint8_t* q, uint8_t* scales, int64_t nblocks, cudaStream_t stream) {
  const int threads = 128;
  const unsigned grid = (unsigned)((nblocks + threads - 1) / threads);
  if (in_dtype == 0) quant_mxfp8_kernel<0><<<grid, threads, 0, stream>>>(x, q, scales, nblocks);
  else if (in_dtype == 1) quant_mxfp8_kernel<1><<<grid, threads, 0, stream>>>(x, q, scales, nblocks);
  else quant_mxfp8_kernel<2><<<grid, threads, 0, stream>>>(x, q, scales, nblocks);
  TA_CUDA_CHECK(cudaGetLastError());
}

void quant_mxfp8_seq_launch(const void* x, int in_dtype, uint8_t* q, uint8_t* scales, int64_t bh, int S, int D,
                            cudaStream_t stream) {
  if (D % 32 != 0 || D > 1024) throw std::runtime_error("quant_mxfp8_seq: head_dim must be a multiple of 32, <= 1024");
  const int T = (S + 127) / 128;
  const unsigned grid = (unsigned)(bh * T * 4);
  if (in_dtype == 0) quant_mxfp8_seq_kernel<0><<<grid, D, 0, stream>>>(x, q, scales, S, D, T);
  else if (in_dtype == 1) quant_mxfp8_seq_kernel<1><<<grid, D, 0, stream>>>(x, q, scales, S, D, T);
  else quant_mxfp8_seq_kernel<2><<<grid, D, 0, stream>>>(x, q, scales, S, D, T);
  TA_CUDA_CHECK(cudaGetLastError());
}

void mxfp8_seq_append_launch(const void* x, int in_dtype, uint8_t* q, uint8_t* scales, int64_t bh, int S, int D, int pos, int n,
                             cudaStream_t stream) {
  if (D % 32 != 0 || D > 1024) throw std::runtime_error("mxfp8_seq_append: head_dim must be a multiple of 32, <= 1024");
  if (pos < 0 || n <= 0 || pos + n > S) throw std::runtime_error("mxfp8_seq_append: rows out of range");
  const int T = (S + 127) / 128;
  const int g0 = pos / 32, g1 = (pos + n - 1) / 32;
  const int nblk = g1 - g0 + 1;
  const unsigned grid = (unsigned)(bh * nblk);
  if (in_dtype == 0) mxfp8_seq_append_kernel<0><<<grid, D, 0, stream>>>(x, q, scales, S, D, T, pos, n, g0, nblk);
  else if (in_dtype == 1) mxfp8_seq_append_kernel<1><<<grid, D, 0, stream>>>(x, q, scales, S, D, T, pos, n, g0, nblk);
  else mxfp8_seq_append_kernel<2><<<grid, D, 0, stream>>>(x, q, scales, S, D, T, pos, n, g0, nblk);
  TA_CUDA_CHECK(cudaGetLastError());
}

void dequant_mxfp8_launch(const uint8_t* q, const uint8_t* scales, float* y, int64_t nblocks, cudaStream_t stream) {
  const int threads = 128;
  dequant_mxfp8_kernel<<<(unsigned)((nblocks + threads - 1) / threads), threads, 0, stream>>>(q, scales, y, nblocks);
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ta
