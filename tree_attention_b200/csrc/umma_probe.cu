// umma_probe -- a one-CTA tcgen05 GEMM used by tests to pin down, on real hardware, every layout
// convention the attention kernels rely on: TMA 128B-swizzled tiles, K-major and MN-major shared
// memory descriptors, the instruction descriptor, A-operand-from-TMEM, and the tcgen05.ld lane map.
//
//   C[128, N] (fp32) = A[128, K] (bf16, row-major) x B
//     b_mn_major = 0 : B is (N, K) row-major  ("K-major", like K in Q K^T)
//     b_mn_major = 1 : B is (K, N) row-major  ("MN-major", like V in P V)
//     a_from_tmem = 1: A is written to TMEM with tcgen05.st and consumed by the .ts MMA form (like P)
// Reference: none (hardware layout probes for the tcgen05 kernels).
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace ta {
namespace {

struct ProbeParams {
  const uint16_t* a;
  float* c;
  int N, K, b_mn_major, a_from_tmem;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                  const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K;
  uint8_t* a_s = smem;                       // [K/64][128][128B]
  uint8_t* b_s = smem + (K / 64) * 16384;    // K-major: [K/64][N][128B] ; MN-major: [N/64][K][128B]

  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t bytes = 128 * K * 2 + N * K * 2;
    mbar_arrive_expect_tx(&bar_load, bytes);
    for (int ka = 0; ka < K / 64; ++ka) tma_load_2d(a_s + ka * 16384, &amap, &bar_load, ka * 64, 0);
    if (!p.b_mn_major) {
      for (int ka = 0; ka < K / 64; ++ka) tma_load_2d(b_s + ka * (N * 128), &bmap, &bar_load, ka * 64, 0);
    } else {
      for (int na = 0; na < N / 64; ++na) tma_load_2d(b_s + na * (K * 128), &bmap, &bar_load, na * 64, 0);
    }
  }
  if (p.a_from_tmem) {
    // thread t owns row t of A: pack bf16 pairs, 32 columns (64 elements) per tcgen05.st
    const uint16_t* arow = p.a + (size_t)tid * K;
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = uint32_t(arow[(c0 + i) * 2]) | (uint32_t(arow[(c0 + i) * 2 + 1]) << 16);
      tmem_st_32x32b_x32(tmem + (uint32_t(warp * 32) << 16) + 256 + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc(1, 1, 128, N, 0, p.b_mn_major ? 1 : 0);
      for (int kk = 0; kk < K / 16; ++kk) {
        uint64_t bdesc;
        if (!p.b_mn_major)
          bdesc = umma_smem_desc_sw128(smem_u32(b_s) + (kk / 4) * (N * 128) + (kk % 4) * 32, 0, 1024);
        else
          bdesc = umma_smem_desc_sw128(smem_u32(b_s) + kk * 2048, K * 128, 1024);
        if (p.a_from_tmem) {
          umma_ts_f16(tmem, tmem + 256 + kk * 8, bdesc, idesc, kk > 0 ? 1u : 0u);
        } else {
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(a_s) + (kk / 4) * 16384 + (kk % 4) * 32, 0, 1024);
          umma_ss_f16(tmem, adesc, bdesc, idesc, kk > 0 ? 1u : 0u);
        }
      }
      umma_commit(&bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) p.c[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ---- block-scaled fp8 probe: C[128, N] = (A8 o SFA)[128, 128] x (B8 o SFB)[N, 128]^T, e4m3 + UE8M0 per 32 ----
struct BsProbeParams {
  const uint32_t* sfa;  // [128] words: 4 UE8M0 bytes per row (K-blocks 0..3)
  const uint32_t* sfb;  // [N] words
  float* c;
  int N;
  int a_mn_major;   // A is stored transposed ([K][M], the layout of a [key][channel] tile consumed along the keys)
};

__global__ void __launch_bounds__(128, 1)
umma_bs_probe_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                     const BsProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N;
  uint8_t* a_s = smem;              // [128 rows][128 B]  (K = 128 e4m3 = one swizzle atom)
  uint8_t* b_s = smem + 128 * 128;  // [N rows][128 B]
  if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_load, 128 * 128 + N * 128);
    tma_load_2d(a_s, &amap, &bar_load, 0, 0);
    tma_load_2d(b_s, &bmap, &bar_load, 0, 0);
  }
  // scale factors -> TMEM: lane l of EVERY lane quarter holds, in column j, the scale word of row 32 j + l
  {
    const uint32_t la = uint32_t(warp * 32) << 16;
    tmem_st_32x32b_x4(tmem + la + 256, p.sfa[lane], p.sfa[32 + lane], p.sfa[64 + lane], p.sfa[96 + lane]);
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = (32 * j + lane < N) ? p.sfb[32 * j + lane] : 0x7f7f7f7fu;
    tmem_st_32x32b_x4(tmem + la + 264, w[0], w[1], w[2], w[3]);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (elect_one()) {
      for (int kk = 0; kk < 4; ++kk) {  // K = 32 per instruction; scale byte kk of the column words
        const uint32_t idesc = umma_idesc_block_scaled(0, 0, 128, N, p.a_mn_major ? 1 : 0, 0, kk, kk);
        // K-major A: 32 bytes further along the row; MN-major A: 32 rows (K) further down the [K][M] tile
        const uint64_t a_desc = p.a_mn_major ? umma_smem_desc_sw128(smem_u32(a_s) + kk * 32 * 128, 128 * 128, 1024)
                                             : umma_smem_desc_sw128(smem_u32(a_s) + kk * 32, 0, 1024);
        umma_ss_mxf8_block_scale(tmem, a_desc,
                                 umma_smem_desc_sw128(smem_u32(b_s) + kk * 32, 0, 1024), idesc, tmem + 256, tmem + 264,
                                 kk > 0 ? 1u : 0u);
      }
      umma_commit(&bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) p.c[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ---- 2-CTA probe (EXPERIMENTAL: compile-checked only, not yet run on hardware -- docs/NEXT.md) ---------------------------
// One tcgen05.mma.cta_group::2 GEMM, M = 256 over a CTA pair: CTA r holds rows [128 r, +128) of A and HALF of B's N rows
// (rows [N/2 r, +N/2)) in its own shared memory at identical offsets; each CTA's TMEM receives its 128 rows of C.
// C[256, N] = A[256, 64] * B[N, 64]^T, bf16 inputs, N = 128.  The leader CTA issues the MMAs and signals both CTAs with a
// multicast commit.  PTX forms follow cute/arch/mma_sm100_umma.hpp and cutlass/arch/barrier.h.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_2cta_probe_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, float* c, int N) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  uint8_t* a_s = smem;                 // [128 rows][64 bf16] = one 128B-swizzled atom column
  uint8_t* b_s = smem + 128 * 128;     // [N / 2 rows][64 bf16]
  if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) {   // the same warp id in both CTAs allocates the same columns in both TMEMs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {    // every CTA loads its own operand halves into its own shared memory
    mbar_arrive_expect_tx(&bar_load, 128 * 128 + (N / 2) * 128);
    tma_load_2d(a_s, &amap, &bar_load, 0, (int)rank * 128);
    tma_load_2d(b_s, &bmap, &bar_load, 0, (int)rank * (N / 2));
  }
  mbar_wait(&bar_load, 0);
  cluster_sync_all();   // both CTAs' operands are in place before the leader issues
  if (rank == 0 && warp == 1) {
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc(1, 1, 256, N, 0, 0);
      const uint32_t zero = 0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = umma_smem_desc_sw128(smem_u32(a_s) + kk * 32, 0, 1024);
        const uint64_t bd = umma_smem_desc_sw128(smem_u32(b_s) + kk * 32, 0, 1024);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
            ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(kk > 0 ? 1u : 0u), "r"(zero)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar_mma)), "h"((uint16_t)3) : "memory");
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) c[((size_t)rank * 128 + tid) * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

// TMEM read bandwidth probe: `warps` warps (4 or 8) of one CTA each read 128 fp32 columns of their 32-lane quarter
// (4 x tcgen05.ld.32x32b.x32 + wait) `iters` times; out[0] = cycles, out[1] = checksum.
__global__ void __launch_bounds__(256, 1) tmem_ld_bw_kernel(long long* out, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_s + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[128];
    tmem_ld_32x32b_x32(taddr + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
    tmem_ld_32x32b_x32(taddr + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
    tmem_ld_32x32b_x32(taddr + 96, *reinterpret_cast<uint32_t(*)[32]>(&r[96]));
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 128; i += 16) acc ^= r[i];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345u) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<128>(tmem_base_s); }
}

}  // namespace

long long tmem_ld_bw_probe(int warps, int iters, cudaStream_t stream) {
  long long* d;
  TA_CUDA_CHECK(cudaMalloc(&d, 2 * sizeof(long long)));
  tmem_ld_bw_kernel<<<1, warps * 32, 0, stream>>>(d, iters);
  TA_CUDA_CHECK(cudaGetLastError());
  long long h[2] = {0, 0};
  TA_CUDA_CHECK(cudaMemcpyAsync(h, d, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  TA_CUDA_CHECK(cudaStreamSynchronize(stream));
  TA_CUDA_CHECK(cudaFree(d));
  return h[0];
}

void umma_2cta_probe_launch(const void* a, const void* b, float* c, int N, cudaStream_t stream) {
  if (N != 128) throw std::runtime_error("umma_2cta_probe: N must be 128");
  auto enc = get_encode_tiled();
  auto mk = [&](const void* base, uint64_t rows, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {64 * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("umma_2cta_probe: cuTensorMapEncodeTiled failed " + std::to_string((int)r));
    return m;
  };
  CUtensorMap amap = mk(a, 256, 128), bmap = mk(b, N, N / 2);
  const size_t smem = 1024 + 128 * 128 + (size_t)(N / 2) * 128;
  TA_CUDA_CHECK(cudaFuncSetAttribute(umma_2cta_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_2cta_probe_kernel<<<2, 128, smem, stream>>>(amap, bmap, c, N);
  TA_CUDA_CHECK(cudaGetLastError());
}

void umma_bs_probe_launch(const void* a8, const void* b8, const void* sfa, const void* sfb, float* c, int N,
                          cudaStream_t stream, int a_mn_major) {
  if (N % 16 != 0 || N < 16 || N > 128) throw std::runtime_error("umma_bs_probe: bad N");
  auto enc = get_encode_tiled();
  auto mk = [&](const void* base, uint64_t rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {128, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {128, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("umma_bs_probe: cuTensorMapEncodeTiled failed " + std::to_string((int)r));
    return m;
  };
  CUtensorMap amap = mk(a8, 128), bmap = mk(b8, N);
  BsProbeParams p{reinterpret_cast<const uint32_t*>(sfa), reinterpret_cast<const uint32_t*>(sfb), c, N, a_mn_major};
  const size_t smem = 1024 + 128 * 128 + (size_t)N * 128;
  TA_CUDA_CHECK(cudaFuncSetAttribute(umma_bs_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_bs_probe_kernel<<<1, 128, smem, stream>>>(amap, bmap, p);
  TA_CUDA_CHECK(cudaGetLastError());
}

void umma_probe_launch(const void* a, const void* b, float* c, int N, int K, int b_mn_major, int a_from_tmem,
                       cudaStream_t stream) {
  if (K % 64 != 0 || K > 256 || N % 16 != 0 || N < 16 || N > 256) throw std::runtime_error("umma_probe: bad N/K");
  if (b_mn_major && N % 64 != 0) throw std::runtime_error("umma_probe: MN-major B needs N % 64 == 0");
  // the 2-D TMA instruction needs rank-2 maps
  CUtensorMap amap, bmap;
  {
    auto enc = get_encode_tiled();
    auto mk = [&](const void* base, uint64_t inner, uint64_t outer, uint32_t box_in, uint32_t box_out) {
      CUtensorMap m;
      cuuint64_t dims[2] = {inner, outer};
      cuuint64_t strides[1] = {inner * 2};
      cuuint32_t box[2] = {box_in, box_out};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) throw std::runtime_error("umma_probe: cuTensorMapEncodeTiled failed " + std::to_string((int)r));
      return m;
    };
    amap = mk(a, K, 128, 64, 128);
    bmap = b_mn_major ? mk(b, N, K, 64, K) : mk(b, K, N, 64, N);
  }
  ProbeParams p{reinterpret_cast<const uint16_t*>(a), c, N, K, b_mn_major, a_from_tmem};
  const size_t smem = 1024 + (size_t)128 * K * 2 + (size_t)N * K * 2;
  TA_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, stream>>>(amap, bmap, p);
  TA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ta
