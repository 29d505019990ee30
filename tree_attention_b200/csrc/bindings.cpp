// Python bindings (torch extension) for the sm_100a kernels and the symmetric-memory runtime.
// Reference: the Python <-> native boundary; /root/reference/model.py calls stock torch ops instead (model.py:74-80, 108-115).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include <cstring>

#include "host_utils.h"
#include "kernels.h"

namespace py = pybind11;
using ta::AttnShape;
using ta::CommCtxHost;

namespace {

// ---------------------------------------------------------------------------------------------
// Symmetric memory: cudaMalloc + CUDA IPC handles.  One buffer per rank, opened by every peer on the
// node; NVLink P2P access is enabled lazily by cudaIpcOpenMemHandle.  Handle exchange happens in
// Python over the bootstrap process group (parallel/symm.py).
// ---------------------------------------------------------------------------------------------
py::tuple symm_alloc(int64_t nbytes) {
  void* p = nullptr;
  TA_CUDA_CHECK(cudaMalloc(&p, (size_t)nbytes));
  TA_CUDA_CHECK(cudaMemset(p, 0, (size_t)nbytes));
  TA_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  TA_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  return py::make_tuple((int64_t) reinterpret_cast<uintptr_t>(p),
                        py::bytes(reinterpret_cast<const char*>(&h), sizeof(h)));
}
int64_t symm_open(const std::string& handle) {
  if (handle.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("symm_open: bad handle size");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  TA_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return (int64_t) reinterpret_cast<uintptr_t>(p);
}
void symm_close(int64_t ptr) { TA_CUDA_CHECK(cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr))); }
void symm_free(int64_t ptr) { TA_CUDA_CHECK(cudaFree(reinterpret_cast<void*>(ptr))); }
void symm_memset(int64_t ptr, int value, int64_t nbytes) {
  TA_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<void*>(ptr), value, (size_t)nbytes,
                                at::cuda::getCurrentCUDAStream()));
}
std::vector<uint32_t> symm_read_u32(int64_t ptr, int n) {
  std::vector<uint32_t> v(n);
  TA_CUDA_CHECK(cudaMemcpy(v.data(), reinterpret_cast<void*>(ptr), n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return v;
}

struct Comm {
  CommCtxHost h;
  Comm(int rank, int world, std::vector<int64_t> data, std::vector<int64_t> flags, int64_t epoch, int64_t status,
       int64_t data_bytes, int64_t flag_bytes, double timeout_s) {
    if (world > ta::kMaxWorldHost) throw std::runtime_error("world size > 16 not supported");
    if ((int)data.size() != world || (int)flags.size() != world) throw std::runtime_error("Comm: pointer table size");
    h.rank = rank;
    h.world = world;
    for (int i = 0; i < world; ++i) {
      h.data[i] = reinterpret_cast<void*>(data[i]);
      h.flags[i] = reinterpret_cast<void*>(flags[i]);
    }
    h.epoch = reinterpret_cast<void*>(epoch);
    h.status = reinterpret_cast<void*>(status);
    h.data_bytes = (size_t)data_bytes;
    h.flag_bytes = (size_t)flag_bytes;
    h.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  }
};

// host_io: q and out may live in pinned (page-locked, device-mapped) HOST memory -- the decode kernels read the few KB of q
// and write the result with plain global accesses, so a latency-bound lone step can skip both copy-engine hops.
AttnShape make_shape(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& out,
                     double scale, bool causal, int64_t q_pos0, int64_t kv_pos0, bool host_io = false) {
  TORCH_CHECK(k.is_cuda() && v.is_cuda(), "k and v must be CUDA tensors");
  TORCH_CHECK((q.is_cuda() || (host_io && q.is_pinned())) && (out.is_cuda() || (host_io && out.is_pinned())),
              host_io ? "q / out must be CUDA or pinned host tensors" : "tensors must be CUDA");
  TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4 && out.dim() == 4, "expected (B, H, S, D) tensors");
  TORCH_CHECK(q.scalar_type() == at::kBFloat16 || q.scalar_type() == at::kHalf, "q must be bf16 or fp16");
  TORCH_CHECK(k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type() &&
                  out.scalar_type() == q.scalar_type(), "q, k, v, out must share a dtype");
  TORCH_CHECK(q.stride(3) == 1 && k.stride(3) == 1 && v.stride(3) == 1 && out.stride(3) == 1,
              "head_dim must be contiguous");
  TORCH_CHECK(k.sizes() == v.sizes(), "k and v shapes differ");
  TORCH_CHECK(q.size(0) == k.size(0) && q.size(3) == k.size(3), "batch / head_dim mismatch");
  TORCH_CHECK(out.sizes() == q.sizes(), "out must have q's shape");
  AttnShape s;
  s.B = (int)q.size(0); s.Hq = (int)q.size(1); s.Sq = (int)q.size(2); s.D = (int)q.size(3);
  s.Hkv = (int)k.size(1); s.S = (int)k.size(2);
  s.is_bf16 = q.scalar_type() == at::kBFloat16;
  s.softmax_scale = (float)scale;
  s.causal = causal;
  s.q_pos0 = q_pos0; s.kv_pos0 = kv_pos0;
  s.q_sb = q.stride(0); s.q_sh = q.stride(1); s.q_ss = q.stride(2);
  s.k_sb = k.stride(0); s.k_sh = k.stride(1); s.k_ss = k.stride(2);
  s.v_sb = v.stride(0); s.v_sh = v.stride(1); s.v_ss = v.stride(2);
  s.o_sb = out.stride(0); s.o_sh = out.stride(1); s.o_ss = out.stride(2);
  return s;
}

// optional device scalar with the number of valid KV rows (int32, 1 element); nullptr when absent
const int* kv_len_ptr(const c10::optional<at::Tensor>& kv_len, const at::Tensor& like) {
  if (!kv_len.has_value()) return nullptr;
  TORCH_CHECK(kv_len->is_cuda() && kv_len->device() == like.device() && kv_len->scalar_type() == at::kInt && kv_len->numel() == 1,
              "kv_len must be a 1-element int32 CUDA tensor on the KV cache's device");
  return kv_len->data_ptr<int>();
}

py::tuple decode_plan(int B, int Hq, int Hkv, int Sq, int S, int D) {
  AttnShape s;
  s.B = B; s.Hq = Hq; s.Hkv = Hkv; s.Sq = Sq; s.S = S; s.D = D;
  int grid, mp, R;
  size_t pf, cf, cfl;
  ta::decode_simt_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cf, &cfl);
  return py::make_tuple(grid, mp, R, (int64_t)pf, (int64_t)cf, (int64_t)cfl);
}

void decode_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, at::Tensor& out,
                c10::optional<at::Tensor> lse, at::Tensor& part, at::Tensor& tickets, py::object comm, double scale,
                bool causal, int64_t q_pos0, int64_t kv_pos0, int pdl, c10::optional<at::Tensor> kv_len) {
  c10::cuda::CUDAGuard guard(q.device());
  AttnShape s = make_shape(q, k, v, out, scale, causal, q_pos0, kv_pos0);
  TORCH_CHECK(part.scalar_type() == at::kFloat && part.is_contiguous(), "part must be contiguous fp32");
  TORCH_CHECK(tickets.scalar_type() == at::kInt && tickets.numel() >= s.B * s.Hkv + 2, "tickets too small");
  int grid, mp, R;
  size_t pf, cf, cfl;
  ta::decode_simt_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cf, &cfl);
  TORCH_CHECK((size_t)part.numel() >= pf, "part workspace too small: need ", pf, " floats");
  float* lse_p = nullptr;
  if (lse.has_value()) {
    TORCH_CHECK(lse->scalar_type() == at::kFloat && lse->is_contiguous() && lse->numel() == (int64_t)s.B * s.Hq * s.Sq,
                "lse must be contiguous fp32 (B, Hq, Sq)");
    lse_p = lse->data_ptr<float>();
  }
  CommCtxHost c;
  if (!comm.is_none()) c = comm.cast<Comm&>().h;
  ta::decode_simt_launch(s, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse_p, part.data_ptr<float>(),
                         reinterpret_cast<uint32_t*>(tickets.data_ptr<int>()), c, ta::num_sms(),
                         at::cuda::getCurrentCUDAStream(), nullptr, nullptr, pdl, kv_len_ptr(kv_len, k));
}

py::tuple attn_fwd_comm_bytes(int B, int Hq, int Sq, int D, int world, int comm_mode) {
  AttnShape s;
  s.B = B; s.Hq = Hq; s.Sq = Sq; s.D = D;
  size_t fb = 0;
  size_t db = ta::attn_fwd_comm_bytes(s, world, &fb, comm_mode);
  return py::make_tuple((int64_t)db, (int64_t)fb);
}

// comm_mode (with a Comm): 1 = replicated output, 2 = output sharded over Sq -- out is (B, Hq, sq_out, D), lse (B, Hq, sq_out)
// with sq_out = ceil(ceil(Sq / 128) / world) * 128, the rows [rank * sq_out, ...) of the global result.
void attn_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, at::Tensor& out, at::Tensor& lse,
              double scale, bool causal, int64_t q_pos0, int64_t kv_pos0, py::object comm, int variant, int comm_mode,
              int64_t kv_seg_len, int64_t kv_seg_gap) {
  c10::cuda::CUDAGuard guard(q.device());
  const bool sharded = !comm.is_none() && comm_mode == 2;
  AttnShape s = make_shape(q, k, v, sharded ? q : out, scale, causal, q_pos0, kv_pos0);
  s.kv_seg_len = (int)kv_seg_len; s.kv_seg_gap = kv_seg_gap;   // zigzag shards: see kernels.h
  int64_t sq_out = s.Sq;
  if (sharded) {
    TORCH_CHECK(out.dim() == 4 && out.is_cuda() && out.scalar_type() == q.scalar_type() && out.stride(3) == 1 &&
                    out.size(0) == q.size(0) && out.size(1) == q.size(1) && out.size(3) == q.size(3),
                "sharded out must be (B, Hq, sq_out, D) in q's dtype");
    sq_out = out.size(2);
    const int world = comm.cast<Comm&>().h.world;
    const int64_t num_m = (s.Sq + 127) / 128;
    TORCH_CHECK(sq_out == (num_m + world - 1) / world * 128, "sharded out must have ceil(ceil(Sq/128)/world)*128 rows");
    s.o_sb = out.stride(0); s.o_sh = out.stride(1); s.o_ss = out.stride(2);
  }
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.numel() == (int64_t)s.B * s.Hq * sq_out,
              "lse must be contiguous fp32 (B, Hq, Sq) [(B, Hq, sq_out) for a sharded output]");
  CommCtxHost c;
  if (!comm.is_none()) c = comm.cast<Comm&>().h;
  // variant 0 / 1: the M = 128 kernel with double-buffered S; 6: the same kernel with the query tile kept in TMEM.
  // The other pipelines of rounds 1-2 were measured and removed (DESIGN.md section 5b keeps the numbers): M = 256 /
  // split-softmax variants 2-5 landed within +-4 % of variant 1; the 2-CTA kernel (cta_group::2 MMAs + TMA multicast,
  // variant 7) passed its tests in round 2 but ran at 734-800 TFLOP/s against 1165 / 1006 for variant 1.
  TORCH_CHECK(variant == 0 || variant == 1 || variant == 6, "attn_fwd: unknown variant ", variant);
  ta::attn_fwd_launch(s, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), c,
                      at::cuda::getCurrentCUDAStream(), variant == 6 ? 1 : 0, comm_mode, (int)sq_out);
}

py::tuple decode_tc_plan(int B, int Hq, int Hkv, int Sq, int S, int D) {
  AttnShape s;
  s.B = B; s.Hq = Hq; s.Hkv = Hkv; s.Sq = Sq; s.S = S; s.D = D;
  int grid, mp, R;
  size_t pf, cb;
  ta::decode_tc_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cb);
  return py::make_tuple(grid, mp, R, (int64_t)pf, (int64_t)cb);
}

void decode_tc_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, at::Tensor& out,
                   c10::optional<at::Tensor> lse, at::Tensor& part, at::Tensor& tickets, py::object comm, double scale,
                   bool causal, int64_t q_pos0, int64_t kv_pos0, bool swap, c10::optional<at::Tensor> kv_len) {
  c10::cuda::CUDAGuard guard(q.device());
  AttnShape s = make_shape(q, k, v, out, scale, causal, q_pos0, kv_pos0);
  TORCH_CHECK(part.scalar_type() == at::kFloat && part.is_contiguous(), "part must be contiguous fp32");
  int grid, mp, R;
  size_t pf, cb;
  ta::decode_tc_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cb);
  TORCH_CHECK((size_t)part.numel() >= pf, "part workspace too small: need ", pf, " floats");
  TORCH_CHECK(tickets.scalar_type() == at::kInt && tickets.numel() >= s.B * s.Hkv + 2, "tickets too small");
  TORCH_CHECK(q.stride(2) % 8 == 0 || q.size(2) == 1, "q rows must be 16-byte aligned");
  float* lse_p = nullptr;
  if (lse.has_value()) {
    TORCH_CHECK(lse->scalar_type() == at::kFloat && lse->is_contiguous() && lse->numel() == (int64_t)s.B * s.Hq * s.Sq);
    lse_p = lse->data_ptr<float>();
  }
  CommCtxHost c;
  if (!comm.is_none()) c = comm.cast<Comm&>().h;
  if (swap)
    ta::decode_swap_launch(s, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse_p, part.data_ptr<float>(),
                           reinterpret_cast<uint32_t*>(tickets.data_ptr<int>()), c, ta::num_sms(), at::cuda::getCurrentCUDAStream(),
                           nullptr, nullptr, nullptr, nullptr, kv_len_ptr(kv_len, k));
  else
    ta::decode_tc_launch(s, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse_p, part.data_ptr<float>(),
                         reinterpret_cast<uint32_t*>(tickets.data_ptr<int>()), c, ta::num_sms(), at::cuda::getCurrentCUDAStream(),
                         nullptr, nullptr, kv_len_ptr(kv_len, k));
}

// block-scaled (MX) fp8 KV cache on the tensor cores (tcgen05.mma.kind::mxf8f6f4.block_scale, swap-AB decode kernel):
//   k8 / v8 uint8 (B, Hkv, S, 128) e4m3;  k_sf uint8 (B, Hkv, S, 4): K's UE8M0 scales per 32 channels (standard MX);
//   v_sf uint8 (B, Hkv, ceil(S / 128), 128, 4): V's UE8M0 scales per 32 KEYS, grouped per 128-key tile and channel.
void decode_mx_tc_fwd(const at::Tensor& q, const at::Tensor& k8, const at::Tensor& v8, const at::Tensor& k_sf,
                      const at::Tensor& v_sf, at::Tensor& out, c10::optional<at::Tensor> lse, at::Tensor& part,
                      at::Tensor& tickets, py::object comm, double scale, bool causal, int64_t q_pos0, int64_t kv_pos0,
                      c10::optional<at::Tensor> kv_len) {
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(q.dim() == 4 && k8.dim() == 4 && v8.dim() == 4 && out.sizes() == q.sizes(), "expected (B, H, S, D) tensors");
  TORCH_CHECK(q.scalar_type() == at::kBFloat16 || q.scalar_type() == at::kHalf, "q must be bf16 or fp16");
  TORCH_CHECK(out.scalar_type() == q.scalar_type() && q.stride(3) == 1 && out.stride(3) == 1);
  TORCH_CHECK(k8.scalar_type() == at::kByte && v8.scalar_type() == at::kByte && k8.stride(3) == 1 && v8.stride(3) == 1);
  TORCH_CHECK(k8.sizes() == v8.sizes() && k8.size(3) == 128 && q.size(3) == 128, "mx fp8 decode needs head_dim 128");
  const int64_t B = k8.size(0), H = k8.size(1), S = k8.size(2), T = (S + 127) / 128;
  TORCH_CHECK(k_sf.scalar_type() == at::kByte && k_sf.is_contiguous() && k_sf.numel() == B * H * S * 4, "k_sf must be uint8 (B, Hkv, S, 4)");
  TORCH_CHECK(v_sf.scalar_type() == at::kByte && v_sf.is_contiguous() && v_sf.numel() == B * H * T * 128 * 4,
              "v_sf must be uint8 (B, Hkv, ceil(S/128), 128, 4)");
  TORCH_CHECK(q.stride(2) % 8 == 0 || q.size(2) == 1, "q rows must be 16-byte aligned");
  AttnShape s;
  s.B = (int)q.size(0); s.Hq = (int)q.size(1); s.Sq = (int)q.size(2); s.D = 128;
  s.Hkv = (int)k8.size(1); s.S = (int)k8.size(2);
  s.is_bf16 = q.scalar_type() == at::kBFloat16;
  s.softmax_scale = (float)scale; s.causal = causal; s.q_pos0 = q_pos0; s.kv_pos0 = kv_pos0;
  s.q_sb = q.stride(0); s.q_sh = q.stride(1); s.q_ss = q.stride(2);
  s.k_sb = k8.stride(0); s.k_sh = k8.stride(1); s.k_ss = k8.stride(2);
  s.v_sb = v8.stride(0); s.v_sh = v8.stride(1); s.v_ss = v8.stride(2);
  s.o_sb = out.stride(0); s.o_sh = out.stride(1); s.o_ss = out.stride(2);
  int grid, mp, R;
  size_t pf, cb;
  ta::decode_tc_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cb);
  TORCH_CHECK(R <= 16, "mx fp8 tensor-core decode packs at most 16 query rows per KV head");
  TORCH_CHECK((size_t)part.numel() >= pf && tickets.numel() >= s.B * s.Hkv + 2, "workspace too small");
  float* lse_p = lse.has_value() ? lse->data_ptr<float>() : nullptr;
  CommCtxHost c;
  if (!comm.is_none()) c = comm.cast<Comm&>().h;
  ta::decode_swap_launch(s, q.data_ptr(), k8.data_ptr(), v8.data_ptr(), out.data_ptr(), lse_p, part.data_ptr<float>(),
                         reinterpret_cast<uint32_t*>(tickets.data_ptr<int>()), c, ta::num_sms(), at::cuda::getCurrentCUDAStream(),
                         nullptr, nullptr, reinterpret_cast<const uint32_t*>(k_sf.data_ptr()),
                         reinterpret_cast<const uint32_t*>(v_sf.data_ptr()), kv_len_ptr(kv_len, k8));
}

// per-channel-scaled fp8 KV cache on the tensor cores: k8/v8 uint8 (B, Hkv, S, 128) e4m3, ksc/vsc fp32 (B, Hkv, 128)
void decode_tc_fwd8(const at::Tensor& q, const at::Tensor& k8, const at::Tensor& v8, const at::Tensor& ksc,
                    const at::Tensor& vsc, at::Tensor& out, c10::optional<at::Tensor> lse, at::Tensor& part,
                    at::Tensor& tickets, py::object comm, double scale, bool causal, int64_t q_pos0, int64_t kv_pos0,
                    bool swap, c10::optional<at::Tensor> kv_len) {
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(q.dim() == 4 && k8.dim() == 4 && v8.dim() == 4 && out.sizes() == q.sizes(), "expected (B, H, S, D) tensors");
  TORCH_CHECK(q.scalar_type() == at::kBFloat16 || q.scalar_type() == at::kHalf, "q must be bf16 or fp16");
  TORCH_CHECK(out.scalar_type() == q.scalar_type() && q.stride(3) == 1 && out.stride(3) == 1);
  TORCH_CHECK(k8.scalar_type() == at::kByte && v8.scalar_type() == at::kByte && k8.stride(3) == 1 && v8.stride(3) == 1);
  TORCH_CHECK(k8.sizes() == v8.sizes() && k8.size(3) == 128 && q.size(3) == 128, "fp8 decode needs head_dim 128");
  TORCH_CHECK(ksc.scalar_type() == at::kFloat && vsc.scalar_type() == at::kFloat && ksc.is_contiguous() && vsc.is_contiguous() &&
                  ksc.numel() == k8.size(0) * k8.size(1) * 128 && vsc.numel() == ksc.numel(), "scales must be fp32 (B, Hkv, 128)");
  TORCH_CHECK(q.stride(2) % 8 == 0 || q.size(2) == 1, "q rows must be 16-byte aligned");
  AttnShape s;
  s.B = (int)q.size(0); s.Hq = (int)q.size(1); s.Sq = (int)q.size(2); s.D = 128;
  s.Hkv = (int)k8.size(1); s.S = (int)k8.size(2);
  s.is_bf16 = q.scalar_type() == at::kBFloat16;
  s.softmax_scale = (float)scale; s.causal = causal; s.q_pos0 = q_pos0; s.kv_pos0 = kv_pos0;
  s.q_sb = q.stride(0); s.q_sh = q.stride(1); s.q_ss = q.stride(2);
  s.k_sb = k8.stride(0); s.k_sh = k8.stride(1); s.k_ss = k8.stride(2);
  s.v_sb = v8.stride(0); s.v_sh = v8.stride(1); s.v_ss = v8.stride(2);
  s.o_sb = out.stride(0); s.o_sh = out.stride(1); s.o_ss = out.stride(2);
  int grid, mp, R;
  size_t pf, cb;
  ta::decode_tc_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cb);
  TORCH_CHECK((size_t)part.numel() >= pf && tickets.numel() >= s.B * s.Hkv + 2, "workspace too small");
  float* lse_p = lse.has_value() ? lse->data_ptr<float>() : nullptr;
  CommCtxHost c;
  if (!comm.is_none()) c = comm.cast<Comm&>().h;
  if (swap)
    ta::decode_swap_launch(s, q.data_ptr(), k8.data_ptr(), v8.data_ptr(), out.data_ptr(), lse_p, part.data_ptr<float>(),
                           reinterpret_cast<uint32_t*>(tickets.data_ptr<int>()), c, ta::num_sms(), at::cuda::getCurrentCUDAStream(),
                           ksc.data_ptr<float>(), vsc.data_ptr<float>(), nullptr, nullptr, kv_len_ptr(kv_len, k8));
  else
    ta::decode_tc_launch(s, q.data_ptr(), k8.data_ptr(), v8.data_ptr(), out.data_ptr(), lse_p, part.data_ptr<float>(),
                         reinterpret_cast<uint32_t*>(tickets.data_ptr<int>()), c, ta::num_sms(), at::cuda::getCurrentCUDAStream(),
                         ksc.data_ptr<float>(), vsc.data_ptr<float>(), kv_len_ptr(kv_len, k8));
}

// block-scaled fp8 KV cache decode: k8/v8 uint8 (B, Hkv, S, 128) e4m3, ks/vs uint8 (B, Hkv, S, 4) UE8M0
void decode_fwd_mx(const at::Tensor& q, const at::Tensor& k8, const at::Tensor& v8, const at::Tensor& ks,
                   const at::Tensor& vs, at::Tensor& out, c10::optional<at::Tensor> lse, at::Tensor& part,
                   at::Tensor& tickets, py::object comm, double scale, bool causal, int64_t q_pos0, int64_t kv_pos0,
                   c10::optional<at::Tensor> kv_len) {
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(q.dim() == 4 && k8.dim() == 4 && v8.dim() == 4 && out.sizes() == q.sizes(), "expected (B, H, S, D) tensors");
  TORCH_CHECK(q.scalar_type() == at::kBFloat16 || q.scalar_type() == at::kHalf, "q must be bf16 or fp16");
  TORCH_CHECK(out.scalar_type() == q.scalar_type() && q.stride(3) == 1 && out.stride(3) == 1);
  TORCH_CHECK(k8.scalar_type() == at::kByte && v8.scalar_type() == at::kByte && k8.stride(3) == 1 && v8.stride(3) == 1);
  TORCH_CHECK(k8.sizes() == v8.sizes() && k8.size(3) == 128 && q.size(3) == 128, "mxfp8 decode needs head_dim 128");
  TORCH_CHECK(ks.scalar_type() == at::kByte && vs.scalar_type() == at::kByte && ks.is_contiguous() && vs.is_contiguous() &&
                  ks.numel() == k8.numel() / 32 && vs.numel() == v8.numel() / 32, "scales must be contiguous uint8 (B, Hkv, S, 4)");
  AttnShape s;
  s.B = (int)q.size(0); s.Hq = (int)q.size(1); s.Sq = (int)q.size(2); s.D = 128;
  s.Hkv = (int)k8.size(1); s.S = (int)k8.size(2);
  s.is_bf16 = q.scalar_type() == at::kBFloat16;
  s.softmax_scale = (float)scale; s.causal = causal; s.q_pos0 = q_pos0; s.kv_pos0 = kv_pos0;
  s.q_sb = q.stride(0); s.q_sh = q.stride(1); s.q_ss = q.stride(2);
  s.k_sb = k8.stride(0); s.k_sh = k8.stride(1); s.k_ss = k8.stride(2);
  s.v_sb = v8.stride(0); s.v_sh = v8.stride(1); s.v_ss = v8.stride(2);
  s.o_sb = out.stride(0); s.o_sh = out.stride(1); s.o_ss = out.stride(2);
  int grid, mp, R;
  size_t pf, cf, cfl;
  ta::decode_simt_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cf, &cfl);
  TORCH_CHECK((size_t)part.numel() >= pf && tickets.numel() >= s.B * s.Hkv + 2, "workspace too small");
  float* lse_p = lse.has_value() ? lse->data_ptr<float>() : nullptr;
  CommCtxHost c;
  if (!comm.is_none()) c = comm.cast<Comm&>().h;
  ta::decode_simt_launch(s, q.data_ptr(), k8.data_ptr(), v8.data_ptr(), out.data_ptr(), lse_p, part.data_ptr<float>(),
                         reinterpret_cast<uint32_t*>(tickets.data_ptr<int>()), c, ta::num_sms(),
                         at::cuda::getCurrentCUDAStream(), reinterpret_cast<const uint32_t*>(ks.data_ptr()),
                         reinterpret_cast<const uint32_t*>(vs.data_ptr()), 0, kv_len_ptr(kv_len, k8));
}

// ---------------------------------------------------------------------------------------------
// DecodeStep: a decode-attention step prepared ONCE (tensor maps encoded, parameter block filled, workspace and
// symmetric region bound) and re-launched with a single runtime call -- the host path of a serving loop.  It holds
// references to every tensor the kernel touches, so the memory a captured CUDA graph or an in-flight launch points
// at cannot be freed or re-used while the step object is alive (workspace / region lifetime).
// ---------------------------------------------------------------------------------------------
struct DecodeStep {
  ta::PreparedLaunch launches[3];   // indexed by pdl level (0 plain, 1 dependent launch, 2 + K/V prefetch)
  bool has[3] = {false, false, false};
  std::vector<at::Tensor> keep;
  py::object comm_keep;
  c10::Device device;
  int kernels_per_step = 1;
  std::string impl;

  explicit DecodeStep(c10::Device d) : device(d) {}

  void launch(int pdl) {
    c10::cuda::CUDAGuard guard(device);
    if (pdl < 0 || pdl > 2 || !has[pdl]) pdl = 0;
    launches[pdl].run(at::cuda::getCurrentCUDAStream());
  }
};

// device-visible address of a CUDA tensor or of a pinned (mapped) host tensor
void* io_ptr(const at::Tensor& t) {
  if (t.is_cuda()) return t.data_ptr();
  void* d = nullptr;
  const cudaError_t e = cudaHostGetDevicePointer(&d, t.data_ptr(), 0);
  TORCH_CHECK(e == cudaSuccess && d != nullptr, "pinned host tensor is not mapped into the device address space: ", cudaGetErrorString(e));
  return d;
}

std::shared_ptr<DecodeStep> decode_step(const std::string& impl, const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
                                        at::Tensor& out, c10::optional<at::Tensor> lse, at::Tensor& part, at::Tensor& tickets,
                                        py::object comm, double scale, bool causal, int64_t q_pos0, int64_t kv_pos0,
                                        c10::optional<at::Tensor> kv_len) {
  c10::cuda::CUDAGuard guard(k.device());
  // q / out: device tensors, or pinned host tensors (zero-copy step: the kernel pulls q over PCIe and posts the result back)
  AttnShape s = make_shape(q, k, v, out, scale, causal, q_pos0, kv_pos0, /*host_io=*/true);
  TORCH_CHECK(part.scalar_type() == at::kFloat && part.is_contiguous(), "part must be contiguous fp32");
  TORCH_CHECK(tickets.scalar_type() == at::kInt && tickets.numel() >= s.B * s.Hkv + 2, "tickets too small");
  float* lse_p = nullptr;
  if (lse.has_value()) {
    TORCH_CHECK(lse->scalar_type() == at::kFloat && lse->is_contiguous() && lse->numel() == (int64_t)s.B * s.Hq * s.Sq,
                "lse must be contiguous fp32 (B, Hq, Sq)");
    lse_p = lse->data_ptr<float>();
  }
  CommCtxHost c;
  if (!comm.is_none()) c = comm.cast<Comm&>().h;
  auto st = std::make_shared<DecodeStep>(k.device());
  st->impl = impl;
  st->comm_keep = comm;
  st->keep = {q, k, v, out, part, tickets};
  if (lse.has_value()) st->keep.push_back(*lse);
  if (kv_len.has_value()) st->keep.push_back(*kv_len);
  const int* kvl = kv_len_ptr(kv_len, k);
  uint32_t* tk = reinterpret_cast<uint32_t*>(tickets.data_ptr<int>());
  if (impl == "simt") {
    int grid, mp, R;
    size_t pf, cf, cfl;
    ta::decode_simt_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cf, &cfl);
    TORCH_CHECK((size_t)part.numel() >= pf, "part workspace too small: need ", pf, " floats");
    for (int pdl = 0; pdl < 3; ++pdl) {
      st->launches[pdl] = ta::decode_simt_prepare(s, io_ptr(q), k.data_ptr(), v.data_ptr(), io_ptr(out), lse_p,
                                                  part.data_ptr<float>(), tk, c, ta::num_sms(), nullptr, nullptr, pdl, kvl);
      st->has[pdl] = true;
    }
    st->kernels_per_step = (int)st->launches[0].passes.size();
  } else if (impl == "swap" || impl == "tc") {
    int grid, mp, R;
    size_t pf, cb;
    ta::decode_tc_plan(s, ta::num_sms(), &grid, &mp, &R, &pf, &cb);
    TORCH_CHECK((size_t)part.numel() >= pf, "part workspace too small: need ", pf, " floats");
    TORCH_CHECK(q.stride(2) % 8 == 0 || q.size(2) == 1, "q rows must be 16-byte aligned");
    st->launches[0] = impl == "swap"
                          ? ta::decode_swap_prepare(s, io_ptr(q), k.data_ptr(), v.data_ptr(), io_ptr(out), lse_p,
                                                    part.data_ptr<float>(), tk, c, ta::num_sms(), nullptr, nullptr, nullptr, nullptr, kvl)
                          : ta::decode_tc_prepare(s, io_ptr(q), k.data_ptr(), v.data_ptr(), io_ptr(out), lse_p,
                                                  part.data_ptr<float>(), tk, c, ta::num_sms(), nullptr, nullptr, kvl);
    st->has[0] = true;
  } else {
    TORCH_CHECK(false, "decode_step: impl must be 'simt', 'swap' or 'tc'");
  }
  return st;
}

py::tuple quant_mxfp8(const at::Tensor& x) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && x.size(-1) % 32 == 0, "x must be contiguous with last dim % 32 == 0");
  int dt = x.scalar_type() == at::kBFloat16 ? 0 : x.scalar_type() == at::kHalf ? 1 : 2;
  TORCH_CHECK(dt != 2 || x.scalar_type() == at::kFloat, "x must be bf16, fp16 or fp32");
  auto q = at::empty(x.sizes(), x.options().dtype(at::kByte));
  auto sizes = x.sizes().vec();
  sizes.back() /= 32;
  auto sc = at::empty(sizes, x.options().dtype(at::kByte));
  ta::quant_mxfp8_launch(x.data_ptr(), dt, q.data_ptr<uint8_t>(), sc.data_ptr<uint8_t>(), x.numel() / 32,
                         at::cuda::getCurrentCUDAStream());
  return py::make_tuple(q, sc);
}

// sequence-blocked MX quantiser: x (B, H, S, D) -> (uint8 e4m3 (B, H, S, D), uint8 UE8M0 (B, H, ceil(S/128), D, 4))
py::tuple quant_mxfp8_seq(const at::Tensor& x) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && x.dim() == 4 && x.size(3) % 32 == 0, "x must be contiguous (B, H, S, D), D % 32 == 0");
  int dt = x.scalar_type() == at::kBFloat16 ? 0 : x.scalar_type() == at::kHalf ? 1 : 2;
  TORCH_CHECK(dt != 2 || x.scalar_type() == at::kFloat, "x must be bf16, fp16 or fp32");
  const int64_t B = x.size(0), H = x.size(1), S = x.size(2), D = x.size(3), T = (S + 127) / 128;
  auto q = at::empty(x.sizes(), x.options().dtype(at::kByte));
  auto sc = at::empty({B, H, T, D, 4}, x.options().dtype(at::kByte));
  ta::quant_mxfp8_seq_launch(x.data_ptr(), dt, q.data_ptr<uint8_t>(), sc.data_ptr<uint8_t>(), B * H, (int)S, (int)D,
                             at::cuda::getCurrentCUDAStream());
  return py::make_tuple(q, sc);
}

// in-place KV append into a sequence-blocked MX cache (data (B, H, S, D) uint8, scales (B, H, T, D, 4) uint8): x is (B, H, n, D)
void mxfp8_seq_append(at::Tensor& data, at::Tensor& sc, const at::Tensor& x, int64_t position) {
  c10::cuda::CUDAGuard guard(data.device());
  TORCH_CHECK(data.is_cuda() && data.is_contiguous() && data.dim() == 4 && data.scalar_type() == at::kByte, "data must be contiguous uint8 (B, H, S, D)");
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.size(0) == data.size(0) && x.size(1) == data.size(1) && x.size(3) == data.size(3),
              "x must be (B, H, n, D)");
  const int64_t B = data.size(0), H = data.size(1), S = data.size(2), D = data.size(3), T = (S + 127) / 128;
  TORCH_CHECK(sc.is_cuda() && sc.is_contiguous() && sc.scalar_type() == at::kByte && sc.numel() == B * H * T * D * 4, "scales must be uint8 (B, H, T, D, 4)");
  at::Tensor xc = x.contiguous();
  int dt = xc.scalar_type() == at::kBFloat16 ? 0 : xc.scalar_type() == at::kHalf ? 1 : 2;
  TORCH_CHECK(dt != 2 || xc.scalar_type() == at::kFloat, "x must be bf16, fp16 or fp32");
  ta::mxfp8_seq_append_launch(xc.data_ptr(), dt, data.data_ptr<uint8_t>(), sc.data_ptr<uint8_t>(), B * H, (int)S, (int)D,
                              (int)position, (int)xc.size(2), at::cuda::getCurrentCUDAStream());
}

at::Tensor dequant_mxfp8(const at::Tensor& q, const at::Tensor& sc) {
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(q.is_cuda() && q.is_contiguous() && sc.is_contiguous() && q.scalar_type() == at::kByte &&
              sc.scalar_type() == at::kByte && sc.numel() * 32 == q.numel());
  auto y = at::empty(q.sizes(), q.options().dtype(at::kFloat));
  ta::dequant_mxfp8_launch(q.data_ptr<uint8_t>(), sc.data_ptr<uint8_t>(), y.data_ptr<float>(), q.numel() / 32,
                           at::cuda::getCurrentCUDAStream());
  return y;
}

py::tuple symm_allreduce_sizes(int64_t n, int world) {
  size_t db, fb;
  ta::symm_allreduce_sizes(n, world, &db, &fb);
  return py::make_tuple((int64_t)db, (int64_t)fb);
}

void symm_allreduce(const at::Tensor& x, at::Tensor& y, Comm& comm) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.scalar_type() == at::kFloat && y.scalar_type() == at::kFloat && x.is_contiguous() && y.is_contiguous() &&
              x.numel() == y.numel() && x.numel() % 4 == 0, "symm_allreduce: contiguous fp32 tensors, numel % 4 == 0");
  ta::symm_allreduce_launch(x.data_ptr<float>(), y.data_ptr<float>(), x.numel(), comm.h, at::cuda::getCurrentCUDAStream());
}

void attn_bwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& o, const at::Tensor& dout,
              const at::Tensor& lse, at::Tensor& dq, at::Tensor& dk, at::Tensor& dv, at::Tensor& delta, at::Tensor& lse2,
              double scale, bool causal, int64_t q_pos0, int64_t kv_pos0) {
  c10::cuda::CUDAGuard guard(q.device());
  AttnShape s = make_shape(q, k, v, o, scale, causal, q_pos0, kv_pos0);
  TORCH_CHECK(dout.sizes() == q.sizes() && dout.scalar_type() == q.scalar_type() && dout.stride(3) == 1, "bad dout");
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.numel() == (int64_t)s.B * s.Hq * s.Sq, "bad lse");
  TORCH_CHECK(dq.scalar_type() == at::kFloat && dq.is_contiguous() && dq.numel() == q.numel(), "dq must be contiguous fp32");
  TORCH_CHECK(dk.is_contiguous() && dv.is_contiguous() && dk.numel() == k.numel() && dv.numel() == v.numel() &&
                  dk.scalar_type() == q.scalar_type() && dv.scalar_type() == q.scalar_type(), "bad dk/dv");
  const int64_t sq_pad = (s.Sq + 63) / 64 * 64;
  TORCH_CHECK(delta.scalar_type() == at::kFloat && lse2.scalar_type() == at::kFloat && delta.is_contiguous() &&
                  lse2.is_contiguous() && delta.numel() >= (int64_t)s.B * s.Hq * sq_pad && lse2.numel() >= (int64_t)s.B * s.Hq * sq_pad,
              "delta / lse2 scratch too small");
  ta::attn_bwd_launch(s, q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(), lse.data_ptr<float>(),
                      dq.data_ptr<float>(), dk.data_ptr(), dv.data_ptr(), delta.data_ptr<float>(), lse2.data_ptr<float>(),
                      dout.stride(0), dout.stride(1), dout.stride(2), at::cuda::getCurrentCUDAStream());
}

void combine(const at::Tensor& o_part, const at::Tensor& lse_part, at::Tensor& out, c10::optional<at::Tensor> lse_out,
             Comm& comm, int mode) {
  c10::cuda::CUDAGuard guard(o_part.device());
  TORCH_CHECK(o_part.scalar_type() == at::kFloat && o_part.is_contiguous(), "o_part must be contiguous fp32");
  TORCH_CHECK(lse_part.scalar_type() == at::kFloat && lse_part.is_contiguous(), "lse_part must be contiguous fp32");
  TORCH_CHECK(out.is_contiguous() && out.numel() == o_part.numel(), "out must be contiguous with o_part's size");
  const int D = (int)o_part.size(-1);
  const int64_t rows = o_part.numel() / D;
  TORCH_CHECK(lse_part.numel() == rows, "lse_part must have one entry per row");
  int dt = out.scalar_type() == at::kFloat ? 0 : out.scalar_type() == at::kBFloat16 ? 1 : 2;
  TORCH_CHECK(dt != 2 || out.scalar_type() == at::kHalf, "out must be fp32, bf16 or fp16");
  float* lo = nullptr;
  if (lse_out.has_value()) {
    TORCH_CHECK(lse_out->scalar_type() == at::kFloat && lse_out->is_contiguous() && lse_out->numel() == rows);
    lo = lse_out->data_ptr<float>();
  }
  ta::combine_launch(o_part.data_ptr<float>(), lse_part.data_ptr<float>(), out.data_ptr(), dt, lo, rows, D, comm.h,
                     mode, at::cuda::getCurrentCUDAStream());
}

void umma_probe(const at::Tensor& a, const at::Tensor& b, at::Tensor& c, bool b_mn_major, bool a_from_tmem) {
  c10::cuda::CUDAGuard guard(a.device());
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && c.scalar_type() == at::kFloat);
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && c.is_contiguous());
  TORCH_CHECK(a.size(0) == 128);
  const int K = (int)a.size(1);
  const int N = b_mn_major ? (int)b.size(1) : (int)b.size(0);
  TORCH_CHECK((b_mn_major ? b.size(0) : b.size(1)) == K);
  TORCH_CHECK(c.size(0) == 128 && c.size(1) == N);
  ta::umma_probe_launch(a.data_ptr(), b.data_ptr(), c.data_ptr<float>(), N, K, b_mn_major, a_from_tmem,
                        at::cuda::getCurrentCUDAStream());
}

void umma_bs_probe(const at::Tensor& a8, const at::Tensor& b8, const at::Tensor& sfa, const at::Tensor& sfb, at::Tensor& c,
                   bool a_mn_major) {
  c10::cuda::CUDAGuard guard(a8.device());
  TORCH_CHECK(a8.scalar_type() == at::kByte && b8.scalar_type() == at::kByte && sfa.scalar_type() == at::kByte &&
              sfb.scalar_type() == at::kByte && c.scalar_type() == at::kFloat);
  TORCH_CHECK(a8.is_contiguous() && b8.is_contiguous() && sfa.is_contiguous() && sfb.is_contiguous() && c.is_contiguous());
  TORCH_CHECK(a8.size(0) == 128 && a8.size(1) == 128 && b8.size(1) == 128 && sfa.numel() == 128 * 4 && sfb.numel() == b8.size(0) * 4);
  TORCH_CHECK(c.size(0) == 128 && c.size(1) == b8.size(0));
  ta::umma_bs_probe_launch(a8.data_ptr(), b8.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), c.data_ptr<float>(), (int)b8.size(0),
                           at::cuda::getCurrentCUDAStream(), a_mn_major ? 1 : 0);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "tree_attention_b200 native extension (sm_100a)";
  m.def("symm_alloc", &symm_alloc);
  m.def("symm_open", &symm_open);
  m.def("symm_close", &symm_close);
  m.def("symm_free", &symm_free);
  m.def("symm_memset", &symm_memset);
  m.def("symm_read_u32", &symm_read_u32);
  py::class_<Comm>(m, "Comm")
      .def(py::init<int, int, std::vector<int64_t>, std::vector<int64_t>, int64_t, int64_t, int64_t, int64_t, double>())
      .def_property("skip_publish", [](Comm& c) { return c.h.skip_publish; }, [](Comm& c, int v) { c.h.skip_publish = v; })
      .def_property("timeout_s", [](Comm& c) { return c.h.timeout_ns * 1e-9; },
                    [](Comm& c, double v) { c.h.timeout_ns = (unsigned long long)(v * 1e9); })
      .def_property_readonly("rank", [](Comm& c) { return c.h.rank; })
      .def_property_readonly("world", [](Comm& c) { return c.h.world; });
  m.def("decode_plan", &decode_plan);
  m.def("decode_fwd", &decode_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out"), py::arg("lse"), py::arg("part"),
        py::arg("tickets"), py::arg("comm"), py::arg("scale"), py::arg("causal"), py::arg("q_pos0"), py::arg("kv_pos0"),
        py::arg("pdl") = 0, py::arg("kv_len") = py::none());
  m.def("decode_fwd_mx", &decode_fwd_mx, py::arg("q"), py::arg("k8"), py::arg("v8"), py::arg("ks"), py::arg("vs"), py::arg("out"),
        py::arg("lse"), py::arg("part"), py::arg("tickets"), py::arg("comm"), py::arg("scale"), py::arg("causal"),
        py::arg("q_pos0"), py::arg("kv_pos0"), py::arg("kv_len") = py::none());
  m.def("decode_set_trace", [](c10::optional<at::Tensor> t) {   // profiling aid (bench_tools/trace_decode.py)
    if (!t.has_value()) { ta::decode_simt_set_trace(nullptr); return; }
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kLong && t->is_contiguous() && t->numel() >= 16 * ta::num_sms());
    ta::decode_simt_set_trace(reinterpret_cast<unsigned long long*>(t->data_ptr<int64_t>()));
  });
  m.def("decode_tc_plan", &decode_tc_plan);
  m.def("decode_tc_fwd", &decode_tc_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out"), py::arg("lse"), py::arg("part"),
        py::arg("tickets"), py::arg("comm"), py::arg("scale"), py::arg("causal"), py::arg("q_pos0"), py::arg("kv_pos0"),
        py::arg("swap"), py::arg("kv_len") = py::none());
  m.def("decode_tc_fwd8", &decode_tc_fwd8, py::arg("q"), py::arg("k8"), py::arg("v8"), py::arg("ksc"), py::arg("vsc"), py::arg("out"),
        py::arg("lse"), py::arg("part"), py::arg("tickets"), py::arg("comm"), py::arg("scale"), py::arg("causal"),
        py::arg("q_pos0"), py::arg("kv_pos0"), py::arg("swap"), py::arg("kv_len") = py::none());
  m.def("decode_mx_tc_fwd", &decode_mx_tc_fwd, py::arg("q"), py::arg("k8"), py::arg("v8"), py::arg("k_sf"), py::arg("v_sf"),
        py::arg("out"), py::arg("lse"), py::arg("part"), py::arg("tickets"), py::arg("comm"), py::arg("scale"), py::arg("causal"),
        py::arg("q_pos0"), py::arg("kv_pos0"), py::arg("kv_len") = py::none());
  py::class_<DecodeStep, std::shared_ptr<DecodeStep>>(m, "DecodeStep")
      .def("launch", &DecodeStep::launch, py::arg("pdl") = 0)
      .def_readonly("kernels_per_step", &DecodeStep::kernels_per_step)
      .def_readonly("impl", &DecodeStep::impl);
  m.def("decode_step", &decode_step, py::arg("impl"), py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out"), py::arg("lse"),
        py::arg("part"), py::arg("tickets"), py::arg("comm"), py::arg("scale"), py::arg("causal"), py::arg("q_pos0"),
        py::arg("kv_pos0"), py::arg("kv_len") = py::none());
  m.def("quant_mxfp8", &quant_mxfp8);
  m.def("quant_mxfp8_seq", &quant_mxfp8_seq);
  m.def("mxfp8_seq_append", &mxfp8_seq_append);
  m.def("dequant_mxfp8", &dequant_mxfp8);
  m.def("attn_fwd", &attn_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out"), py::arg("lse"), py::arg("scale"),
        py::arg("causal"), py::arg("q_pos0"), py::arg("kv_pos0"), py::arg("comm"), py::arg("variant") = 0, py::arg("comm_mode") = 1,
        py::arg("kv_seg_len") = 0, py::arg("kv_seg_gap") = 0);
  m.def("attn_fwd_comm_bytes", &attn_fwd_comm_bytes, py::arg("B"), py::arg("Hq"), py::arg("Sq"), py::arg("D"), py::arg("world"),
        py::arg("comm_mode") = 1);
  m.def("attn_fwd_phase_cycles", []() { unsigned long long c[5]; ta::attn_fwd_phase_cycles(c); return py::make_tuple(c[0], c[1], c[2], c[3], c[4]); });
  m.def("attn_bwd", &attn_bwd);
  m.def("symm_allreduce", &symm_allreduce);
  m.def("symm_allreduce_sizes", &symm_allreduce_sizes);
  m.def("combine", &combine);
  m.def("umma_probe", &umma_probe);
  m.def("umma_bs_probe", &umma_bs_probe, py::arg("a8"), py::arg("b8"), py::arg("sfa"), py::arg("sfb"), py::arg("c"),
        py::arg("a_mn_major") = false);
  m.def("umma_2cta_probe", [](const at::Tensor& a, const at::Tensor& b, at::Tensor& c) {   // EXPERIMENTAL, see umma_probe.cu
    c10::cuda::CUDAGuard guard(a.device());
    TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && c.scalar_type() == at::kFloat);
    TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && c.is_contiguous() && a.size(0) == 256 && a.size(1) == 64 &&
                b.size(0) == 128 && b.size(1) == 64 && c.size(0) == 256 && c.size(1) == 128);
    ta::umma_2cta_probe_launch(a.data_ptr(), b.data_ptr(), c.data_ptr<float>(), 128, at::cuda::getCurrentCUDAStream());
  });
  m.def("num_sms", []() { return ta::num_sms(); });
  // host-only: the stream-K split of (B x Hkv x ceil(S / 128)) tiles over `ncta` persistent CTAs -> (grid, max_parts).
  // Needs no device; tests/test_split_cpu.py checks that max_parts really bounds the CTAs sharing one KV head.
  m.def("decode_split_for", [](int B, int Hkv, int S, int ncta) {
    AttnShape s{};
    s.B = B; s.Hkv = Hkv; s.Hq = Hkv; s.Sq = 1; s.S = S; s.D = 128;
    int g = 0, mp = 0;
    ta::decode_split(B * Hkv, S, ncta, &g, &mp);
    return py::make_tuple(g, mp);
  });
  m.def("tmem_ld_bw_probe", [](int warps, int iters) { return ta::tmem_ld_bw_probe(warps, iters, at::cuda::getCurrentCUDAStream()); });
}
