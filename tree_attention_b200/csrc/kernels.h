// Plain-C++ launcher interface between the CUDA translation units and the torch bindings.
// Reference: none (launcher declarations of the native layer; the reference is a single Python file, /root/reference/model.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <vector>

namespace ta {

// A kernel launch (or a short sequence of launches) whose tensor maps and parameter block were encoded once.
// run() costs one cudaLaunchKernelEx per pass: the host path of a serving loop (models/decoder.py, bindings DecodeStep).
struct PreparedLaunch {
  std::vector<std::function<void(cudaStream_t)>> passes;
  void run(cudaStream_t stream) const {
    for (const auto& f : passes) f(stream);
  }
};

constexpr int kMaxWorldHost = 16;

struct CommCtxHost {
  int rank = 0;
  int world = 1;
  void* data[kMaxWorldHost] = {nullptr};
  void* flags[kMaxWorldHost] = {nullptr};
  void* epoch = nullptr;
  void* status = nullptr;
  unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
  int skip_publish = 0;
  size_t data_bytes = 0;
  size_t flag_bytes = 0;
};

struct AttnShape {
  int B = 1, Hq = 1, Hkv = 1, Sq = 1, S = 0, D = 128;
  int is_bf16 = 1;
  float softmax_scale = 1.f;
  int causal = 0;
  int64_t q_pos0 = 0;   // global position of query row 0
  int64_t kv_pos0 = 0;  // global position of local key row 0
  // two-segment shard (zigzag sharding of a causal sequence, forward kernel only): local rows [kv_seg_len, S) sit at
  // kv_pos0 + kv_seg_gap + row.  kv_seg_len = 0: one contiguous segment.  Needs kv_seg_len % 128 == 0, kv_seg_gap >= 0.
  int kv_seg_len = 0;
  int64_t kv_seg_gap = 0;
  // element strides, innermost (D) contiguous
  int64_t q_sb = 0, q_sh = 0, q_ss = 0;
  int64_t k_sb = 0, k_sh = 0, k_ss = 0;
  int64_t v_sb = 0, v_sh = 0, v_ss = 0;
  int64_t o_sb = 0, o_sh = 0, o_ss = 0;
};

// ---- stream-K split shared by the three decode kernels: BH x ceil(cap / 128) tiles over <= ncta persistent CTAs.
// max_parts bounds the CTAs sharing one (batch, kv-head) for EVERY run-time number of valid rows <= cap.
void decode_split(int BH, int cap, int ncta, int* grid, int* max_parts);

// ---- split-KV streaming decode (CUDA-core math, TMA-fed), fused split merge + cross-GPU combine ----
// workspace sizes for a given problem; `grid` is returned so that callers can cache it.
void decode_simt_plan(const AttnShape& s, int num_sms, int* grid, int* max_parts, int* rows_per_pass,
                      size_t* part_floats, size_t* comm_floats, size_t* comm_flags);
// out: same dtype as q.  lse: fp32 (B, Hq, Sq) natural log, may be null.
// part: float workspace, tickets: uint32 [B*Hkv + 2] zero-initialised once.
// kv_len: optional DEVICE scalar with the number of valid rows of the shard (<= S, the capacity the tensor maps
// cover); the kernel reads it at run time, so a captured CUDA graph follows a growing KV cache.
void decode_simt_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                        float* part, uint32_t* tickets, const CommCtxHost& comm, int num_sms,
                        cudaStream_t stream, const uint32_t* kscale = nullptr, const uint32_t* vscale = nullptr,
                        int pdl = 0, const int* kv_len = nullptr);
// profiling aid: every later decode_simt launch/prepare writes [grid][16] u64 timeline stamps into buf (nullptr = off)
void decode_simt_set_trace(unsigned long long* buf);
PreparedLaunch decode_simt_prepare(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                                   float* part, uint32_t* tickets, const CommCtxHost& comm, int num_sms,
                                   const uint32_t* kscale = nullptr, const uint32_t* vscale = nullptr, int pdl = 0,
                                   const int* kv_len = nullptr);
// kscale/vscale != null: K/V are block-scaled fp8 (e4m3 bytes, D = 128) and the scales are (B, Hkv, S) words of
// four UE8M0 exponents (one per 32 elements); strides in AttnShape are then in BYTES == elements.

// ---- tensor-core decode: the R = (Hq/Hkv) x Sq <= 128 query rows of a KV head packed into one tcgen05 tile ----
void decode_tc_split(const AttnShape& s, int ncta, int* grid, int* max_parts);
void decode_tc_plan(const AttnShape& s, int num_sms, int* grid, int* max_parts, int* rows, size_t* part_floats,
                    size_t* comm_bytes);
// kscale / vscale != null: K, V are e4m3 bytes with per-CHANNEL fp32 scales (B, Hkv, D); both GEMMs then run as
// tcgen05 kind::f8f6f4 (q is quantised per row in the kernel, P per element, K's scales are folded into q).
void decode_tc_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse, float* part,
                      uint32_t* tickets, const CommCtxHost& comm, int num_sms, cudaStream_t stream,
                      const float* kscale = nullptr, const float* vscale = nullptr, const int* kv_len = nullptr);
PreparedLaunch decode_tc_prepare(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                                 float* part, uint32_t* tickets, const CommCtxHost& comm, int num_sms,
                                 const float* kscale = nullptr, const float* vscale = nullptr,
                                 const int* kv_len = nullptr);

// ---- swap-AB tensor-core decode (keys on the TMEM lanes, <= 16 packed query columns); head_dim 128 ----
void decode_swap_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse, float* part,
                        uint32_t* tickets, const CommCtxHost& comm, int num_sms, cudaStream_t stream,
                        const float* kscale = nullptr, const float* vscale = nullptr, const uint32_t* k_sf = nullptr,
                        const uint32_t* v_sf = nullptr, const int* kv_len = nullptr);
PreparedLaunch decode_swap_prepare(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                                   float* part, uint32_t* tickets, const CommCtxHost& comm, int num_sms,
                                   const float* kscale = nullptr, const float* vscale = nullptr,
                                   const uint32_t* k_sf = nullptr, const uint32_t* v_sf = nullptr,
                                   const int* kv_len = nullptr);

// ---- stand-alone combine of W per-rank partials (o fp32 normalised, lse natural log) ----
// local: o_part (rows, D) fp32 + lse_part (rows); result written to out (dtype of `is_bf16`/fp16/fp32)
// mode 0: one-shot push (every rank publishes to every peer, merges all W in rank order)
// mode 1: butterfly (log2 W rounds of pairwise exchange)
void combine_launch(const float* o_part, const float* lse_part, void* out, int out_dtype, float* lse_out,
                    int64_t rows, int D, const CommCtxHost& comm, int mode, cudaStream_t stream);

// ---- tcgen05 probes / attention forward (attn_fwd_sm100.cu) ----
// C[M=128, N] (fp32) = A[128, K] * B^T ; a, b bf16.  b_mn_major: B given as (K, N) row-major (V-like).
// a_from_tmem: A is first staged through TMEM (tcgen05.st) and consumed with the .ts MMA form.
void umma_probe_launch(const void* a, const void* b, float* c, int N, int K, int b_mn_major, int a_from_tmem,
                       cudaStream_t stream);

// block-scaled fp8 probe: C[128, N] fp32 = (A8 o SFA)(128 x 128) x (B8 o SFB)(N x 128)^T, e4m3 + UE8M0 per 32 of K,
// through tcgen05.mma.kind::mxf8f6f4.block_scale with the scale factors staged in TMEM.
long long tmem_ld_bw_probe(int warps, int iters, cudaStream_t stream);   // cycles for iters x (warps x 16 KB) of tcgen05.ld
// EXPERIMENTAL (compile-checked only): C[256, 128] = A[256, 64] B[128, 64]^T with one cta_group::2 MMA over a CTA pair
void umma_2cta_probe_launch(const void* a, const void* b, float* c, int N, cudaStream_t stream);
void umma_bs_probe_launch(const void* a8, const void* b8, const void* sfa, const void* sfb, float* c, int N,
                          cudaStream_t stream, int a_mn_major = 0);

// ---- tcgen05 flash-attention forward: shard-local partial (o normalised, lse natural log) ----
// comm.world > 1: fused mode -- compute CTAs push their partial tiles to every peer, merge CTAs in the same
// launch produce the final (replicated) output; no NCCL.
// comm_mode (fused mode only): 1 = replicated output (reduce-scatter of the partial tiles to their owner rank + all-gather
// of the final tiles), 2 = output sharded over Sq: `out` / `lse` hold sq_out = tiles_per_rank * 128 rows per (b, head), the
// rows [rank * sq_out, (rank + 1) * sq_out) of the global result.
void attn_fwd_launch(const AttnShape& s, const void* q, const void* k, const void* v, void* out, float* lse,
                     const CommCtxHost& comm, cudaStream_t stream, int q_in_tmem = 0, int comm_mode = 1, int sq_out = 0);
void attn_fwd_phase_cycles(unsigned long long* out5);   // profiling aid, see attn_fwd_sm100.cu
size_t attn_fwd_comm_bytes(const AttnShape& s, int world, size_t* flag_bytes, int comm_mode = 1);
// ---- tcgen05 flash-attention backward over one KV shard with the GLOBAL o / lse ----
// dq: fp32 (B, Hq, Sq, D) contiguous (this shard's partial); dk, dv: (B, Hkv, S, D) contiguous, I/O dtype;
// delta, lse2: fp32 scratch (B, Hq, ceil64(Sq)).
void attn_bwd_launch(const AttnShape& s, const void* q, const void* k, const void* v, const void* o, const void* dout,
                     const float* lse, float* dq, void* dk, void* dv, float* delta, float* lse2, int64_t do_sb,
                     int64_t do_sh, int64_t do_ss, cudaStream_t stream);

// ---- fp32 all-reduce (sum) over symmetric memory: the backward's dQ reduce (reduce.cu) ----
void symm_allreduce_sizes(int64_t n, int world, size_t* data_bytes, size_t* flag_bytes);
void symm_allreduce_launch(const float* x, float* y, int64_t n, const CommCtxHost& comm, cudaStream_t stream);

// ---- block-scaled fp8 (MX): e4m3 + one UE8M0 scale per 32 elements of the innermost dimension ----
void quant_mxfp8_launch(const void* x, int in_dtype /*0 bf16, 1 fp16, 2 fp32*/, uint8_t* q, uint8_t* scales,
                        int64_t nblocks, cudaStream_t stream);
void quant_mxfp8_seq_launch(const void* x, int in_dtype, uint8_t* q, uint8_t* scales, int64_t bh, int S, int D,
                            cudaStream_t stream);
// KV append into a sequence-blocked MX cache: rows [pos, pos + n) <- x (BH, n, D); the touched 32-key blocks are re-quantised
void mxfp8_seq_append_launch(const void* x, int in_dtype, uint8_t* q, uint8_t* scales, int64_t bh, int S, int D, int pos, int n,
                             cudaStream_t stream);
void dequant_mxfp8_launch(const uint8_t* q, const uint8_t* scales, float* y, int64_t nblocks, cudaStream_t stream);

}  // namespace ta
