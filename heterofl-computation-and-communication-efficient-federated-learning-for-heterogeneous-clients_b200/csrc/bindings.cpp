// Python bindings (pybind11 / ATen) for the hfl_b200 sm_100a kernels and the peer-memory runtime.
#include <torch/extension.h>
#include <cstdlib>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

extern "C" {
void hfl_aggregate(const void*, const void*, int, const void*, int, int, const void*, const void*, const void*, int,
                   int, const void*, int, cudaStream_t);
void hfl_distribute(const void*, const void*, int, const void*, const void*, int, const float*, const void*,
                    const void*, const void*, const void*, int, cudaStream_t);
void hfl_peer_sum(const void*, int, float*, int, cudaStream_t);
int hfl_fed_chunk();
int hfl_fed2_chunk();
void hfl_aggregate_v2(const void*, const void*, int, const int*, const void*, int, int, const void*, const int*, int,
                      const void*, int, const void*, int, const void*, int, float*, int, cudaStream_t);
int hfl_fed_max_clients();
void hfl_sgd_clip_step(float*, void*, void*, int, float*, long, const float*, float, float, float, float*,
                       cudaStream_t);
void hfl_masked_ce(const void*, int, const long long*, const unsigned char*, int, int, int, void*, float*, float*,
                   float, float*, double*, cudaStream_t);
void hfl_augment(const unsigned char*, const long long*, const int*, const float*, const float*, int, int, int, int,
                 int, void*, int, cudaStream_t);
int hfl_mx_sf_bytes(int, int);
int hfl_mx_quantize(const void*, int, int, int, void*, int, void*, cudaStream_t);
int hfl_gemm_mxfp8(const void*, int, const void*, const void*, int, const void*, void*, int, int, const float*, float, int,
                   int, int, cudaStream_t);
int hfl_maxpool2_fwd(const void*, void*, int, int, int, int, cudaStream_t);
int hfl_maxpool2_bwd(const void*, const void*, void*, int, int, int, int, cudaStream_t);
int hfl_gn_act_fwd(const void*, const void*, const void*, const void*, int, void*, void*, float*, float*, int, int, int,
                   int, float, int, cudaStream_t);
int hfl_gn_act_bwd(const void*, const void*, const void*, const void*, const float*, const float*, const void*,
                   const void*, int, void*, float*, void*, void*, int, int, int, int, int, int, float, cudaStream_t);
int hfl_bn_act_fwd(const void*, const void*, void*, void*, const void*, const void*, int, float, int, int, int, int,
                   float*, float*, float*, float*, int, const int*, int, cudaStream_t);
void hfl_affine_act(const void*, const void*, void*, void*, const float*, const float*, long, int, int, cudaStream_t);
int hfl_bn_act_bwd(const void*, const void*, const void*, const void*, const float*, const float*, const void*,
                   const void*, int, int, int, int, float, int, void*, void*, void*, float*, int, const int*, int,
                   cudaStream_t);
int hfl_ms_chunk();
void hfl_emb_sparse_norm(const int*, int, const float*, int, float*, cudaStream_t);
void hfl_emb_sparse_apply(const int*, int, float*, int, long, float*, void*, float*, const float*, float, const float*,
                          cudaStream_t);
void hfl_multi_sgd_clip_step(const void*, const void*, int, float*, void*, float*, const float*, float, float, float,
                             float*, cudaStream_t);
int hfl_gemm_bf16_tn(const void*, int, const void*, int, void*, int, int, const float*, const void*, float, int, int,
                     int, int, cudaStream_t);
void hfl_gemm_timing(long long*);
int hfl_gemm_mc_wanted(int, int, int);
int hfl_gemm_bf16_tn_mc(const void*, int, const void*, int, void*, int, const float*, const void*, float, int, int, int, int,
                        cudaStream_t);
int hfl_lm_max_e();
int hfl_lm_max_s();
int hfl_lm_max_dh();
void hfl_lm_embed_fwd(const long long*, const void*, const void*, const void*, const void*, int, int, int, int, float,
                      float, float, const long long*, const long long*, unsigned, void*, int*, void*, float*,
                      unsigned char*, cudaStream_t);
void hfl_lm_embed_bwd(const void* const*, int, const unsigned char*, float, const void*, const float*, const void*,
                      const int*, int, int, int, float*, float*, float*, float*, cudaStream_t);
void hfl_lm_add_ln_fwd(const void*, const void*, const void*, const void*, int, int, float, float, const long long*,
                       const long long*, unsigned, void*, void*, float*, unsigned char*, cudaStream_t);
void hfl_lm_add_ln_bwd(const void* const*, int, const void*, const float*, const void*, const unsigned char*, float, int,
                       int, void*, void*, float*, float*, float*, float, const float*, cudaStream_t);
int hfl_lm_vocab_supported(int, int, int);
int hfl_lm_vocab_tiles(int);
int hfl_lm_vocab_fwd(const void*, int, const void*, int, const void*, const unsigned char*, const long long*, int, int, int,
                     float*, float*, int*, float*, float*, float, float*, double*, float*, int, cudaStream_t);
int hfl_lm_vocab_bwd(const void*, int, const void*, int, const void*, const unsigned char*, const long long*, const float*,
                     int, int, int, void*, int, float*, float*, cudaStream_t);
void hfl_lm_gelu_drop_fwd(const void*, long, float, const long long*, const long long*, unsigned, void*, unsigned char*,
                          cudaStream_t);
void hfl_lm_gelu_drop_bwd(const void*, const void*, const unsigned char*, float, int, int, void*, float*, float,
                          cudaStream_t);
void hfl_lm_colsum(const void*, int, int, int, float, float*, cudaStream_t);
void hfl_lm_attn_fwd(const void*, const void*, const void*, int, int, int, int, float, void*, void*, cudaStream_t);
int hfl_lm_attn_bwd(const void*, const void*, const void*, const void*, const void*, const void*, int, int, int, int,
                    float, void*, void*, void*, float*, float*, float*, float, int, cudaStream_t);
int hfl_head_smem_bytes(int, int, int);
int hfl_head_fwd_bwd(const void*, int, int, int, const void*, const void*, int, int, const long long*,
                     const unsigned char*, float, float*, double*, void*, void*, void*, float*, const int*,
                     cudaStream_t);
int hfl_bn_apply_sums(const void*, void*, const void*, const void*, int, float, int, int, int, float*, float*, float*,
                      float*, const int*, int, cudaStream_t);
int hfl_conv_tc_supported(int, int, int, int, int, int, int, int);
void hfl_conv_set_ksplit(int);
int hfl_conv_fprop_tc(const void*, const void*, void*, const void*, float*, const int*, int, int, int, int, int, int, int,
                      int, cudaStream_t);
int hfl_conv_fprop_tc2(const void*, const void*, void*, const void*, float*, const int*, int, int, int, int, int, int, int,
                       int, const void*, const void*, int, int, cudaStream_t);
int hfl_gemm_bf16_general(const void*, int, int, const void*, int, int, void*, int, int, float, int, int, int, int,
                          cudaStream_t);
int hfl_gemm_general_splits(int, int, int);
int hfl_gemm_grouped(int, const void* const*, int, int, const void* const*, int, int, void* const*, int, const void* const*,
                     const float*, int, int, int, int, cudaStream_t);
int hfl_conv_dgrad_tc(const void*, const void*, void*, const void*, int, int, int, int, int, int, int, int, cudaStream_t);
int hfl_conv_dgrad_tc_supported(int, int, int, int, int, int, int, int);
int hfl_conv_wgrad_tc_supported(int, int, int, int, int, int, int, int);
int hfl_conv_wgrad_tc(const void*, const void*, float*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int hfl_mega_phase_bytes();
int hfl_mega_smem_bytes();
int hfl_mega_max_grid();
void hfl_mega_occupancy_probe(int*);
int hfl_mega_launch(const void*, int, int, int, void*, int, void*, cudaStream_t);
int hfl_mega_max_jobs();
int hfl_mega_encode_tmap(void*, const void*, int, const unsigned long long*, const unsigned long long*, const unsigned*,
                         const unsigned*);
long hfl_launch_count();
void hfl_launch_count_add(long);
void hfl_peer_barrier(const void*, int, int, unsigned, int*, cudaStream_t);
}

#define CHECK_CUDA_T(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIG(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CUDA_OK(e)                                                                        \
  do {                                                                                    \
    cudaError_t _e = (e);                                                                 \
    TORCH_CHECK(_e == cudaSuccess, "CUDA error: ", cudaGetErrorString(_e), " at ", #e);   \
  } while (0)

static cudaStream_t cur_stream() { return c10::cuda::getCurrentCUDAStream().stream(); }
static void check_launch() { CUDA_OK(cudaGetLastError()); }

// ---- federation ------------------------------------------------------------------------------------------
void aggregate(torch::Tensor tensor_tab, torch::Tensor level_tab, torch::Tensor block_tab, int64_t block_lo,
               int64_t n_blocks, torch::Tensor client_ptr, torch::Tensor client_level, torch::Tensor client_mask,
               torch::Tensor out_ptr) {
  CHECK_CUDA_T(tensor_tab); CHECK_CUDA_T(level_tab); CHECK_CUDA_T(block_tab); CHECK_CUDA_T(client_ptr);
  CHECK_CUDA_T(client_level); CHECK_CUDA_T(client_mask); CHECK_CUDA_T(out_ptr);
  const int T = tensor_tab.size(0), M = client_ptr.size(0);
  TORCH_CHECK(M <= hfl_fed_max_clients(), "too many clients for one aggregate launch");
  TORCH_CHECK(client_mask.dim() == 2 && client_mask.size(0) == M && client_mask.is_contiguous());
  c10::cuda::CUDAGuard guard(tensor_tab.device());
  hfl_aggregate(tensor_tab.data_ptr(), level_tab.data_ptr(), T, block_tab.data_ptr(), (int)block_lo, (int)n_blocks,
                client_ptr.data_ptr(), client_level.data_ptr(), client_mask.data_ptr(), (int)client_mask.size(1), M,
                out_ptr.data_ptr(), (int)out_ptr.size(0), cur_stream());
  check_launch();
}


// clients sorted by level: lvl_start[L+1]; optional fused small all-reduce (red_src pointers -> red_dst)
void aggregate_v2(torch::Tensor tensor_tab, torch::Tensor level_tab, torch::Tensor vec4, torch::Tensor block_tab,
                  int64_t block_lo, int64_t n_blocks, torch::Tensor client_ptr, torch::Tensor lvl_start,
                  torch::Tensor client_mask, torch::Tensor out_ptr, c10::optional<torch::Tensor> red_src,
                  c10::optional<torch::Tensor> red_dst) {
  CHECK_CUDA_T(tensor_tab); CHECK_CUDA_T(level_tab); CHECK_CUDA_T(block_tab); CHECK_CUDA_T(client_ptr);
  CHECK_CUDA_T(lvl_start); CHECK_CUDA_T(client_mask); CHECK_CUDA_T(out_ptr); CHECK_CUDA_T(vec4);
  const int T = tensor_tab.size(0), M = client_ptr.size(0), L = lvl_start.numel() - 1;
  TORCH_CHECK(client_mask.dim() == 2 && client_mask.size(0) == M && client_mask.is_contiguous());
  TORCH_CHECK(lvl_start.scalar_type() == torch::kInt32 && vec4.scalar_type() == torch::kInt32 && vec4.numel() == T);
  const void* rs = nullptr; float* rd = nullptr; int rn = 0, rns = 0;
  if (red_src.has_value() && red_dst.has_value()) {
    CHECK_CUDA_T(*red_src); CHECK_CUDA_T(*red_dst);
    TORCH_CHECK(red_dst->scalar_type() == torch::kFloat32);
    rs = red_src->data_ptr(); rd = red_dst->data_ptr<float>(); rn = (int)red_dst->numel(); rns = (int)red_src->numel();
  }
  c10::cuda::CUDAGuard guard(tensor_tab.device());
  hfl_aggregate_v2(tensor_tab.data_ptr(), level_tab.data_ptr(), T, vec4.data_ptr<int>(), block_tab.data_ptr(),
                   (int)block_lo, (int)n_blocks, client_ptr.data_ptr(), lvl_start.data_ptr<int>(), L,
                   client_mask.data_ptr(), (int)client_mask.size(1), out_ptr.data_ptr(), (int)out_ptr.size(0), rs, rns,
                   rd, rn, cur_stream());
  check_launch();
}

void distribute(torch::Tensor tensor_tab, torch::Tensor level_tab, torch::Tensor lvl_block_tab,
                torch::Tensor lvl_block_start, int64_t max_blocks, torch::Tensor global_w, torch::Tensor w_ptr,
                c10::optional<torch::Tensor> wc_ptr, c10::optional<torch::Tensor> mom_ptr,
                torch::Tensor client_level) {
  CHECK_CUDA_T(global_w); CHECK_CONTIG(global_w);
  TORCH_CHECK(global_w.scalar_type() == torch::kFloat32);
  const int T = tensor_tab.size(0), M = w_ptr.size(0);
  c10::cuda::CUDAGuard guard(global_w.device());
  hfl_distribute(tensor_tab.data_ptr(), level_tab.data_ptr(), T, lvl_block_tab.data_ptr(),
                 lvl_block_start.data_ptr(), (int)max_blocks, global_w.data_ptr<float>(), w_ptr.data_ptr(),
                 wc_ptr.has_value() ? wc_ptr->data_ptr() : nullptr,
                 mom_ptr.has_value() ? mom_ptr->data_ptr() : nullptr, client_level.data_ptr(), M, cur_stream());
  check_launch();
}

void peer_sum(torch::Tensor src_ptr, torch::Tensor dst) {
  CHECK_CUDA_T(dst); CHECK_CONTIG(dst);
  TORCH_CHECK(dst.scalar_type() == torch::kFloat32);
  c10::cuda::CUDAGuard guard(dst.device());
  hfl_peer_sum(src_ptr.data_ptr(), (int)src_ptr.size(0), dst.data_ptr<float>(), (int)dst.numel(), cur_stream());
  check_launch();
}

// ---- optimiser -------------------------------------------------------------------------------------------
void sgd_clip_step(torch::Tensor w, c10::optional<torch::Tensor> wc, torch::Tensor g, torch::Tensor mom,
                   torch::Tensor lr, double mu, double wd, double max_norm, torch::Tensor scratch) {
  CHECK_CUDA_T(w); CHECK_CONTIG(w); CHECK_CONTIG(g); CHECK_CONTIG(mom);
  TORCH_CHECK(w.scalar_type() == torch::kFloat32 && mom.scalar_type() == torch::kFloat32);
  TORCH_CHECK(w.numel() % 8 == 0 && g.numel() == w.numel() && mom.numel() == w.numel());
  TORCH_CHECK(lr.is_cuda() && lr.scalar_type() == torch::kFloat32 && scratch.numel() >= 4);
  const bool gb = g.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(gb || g.scalar_type() == torch::kFloat32);
  void* wcp = nullptr;
  if (wc.has_value()) { TORCH_CHECK(wc->scalar_type() == torch::kBFloat16 && wc->numel() == w.numel()); wcp = wc->data_ptr(); }
  c10::cuda::CUDAGuard guard(w.device());
  hfl_sgd_clip_step(w.data_ptr<float>(), wcp, g.data_ptr(), gb ? 1 : 0, mom.data_ptr<float>(), (long)w.numel(),
                    lr.data_ptr<float>(), (float)mu, (float)wd, (float)max_norm, scratch.data_ptr<float>(),
                    cur_stream());
  check_launch();
}

// ---- loss ------------------------------------------------------------------------------------------------
// returns (loss[1] fp32, dlogits like logits)
std::vector<torch::Tensor> masked_ce(torch::Tensor logits, torch::Tensor target, c10::optional<torch::Tensor> mask,
                                     double n_weight, c10::optional<torch::Tensor> meter, bool need_grad) {
  CHECK_CUDA_T(logits); CHECK_CONTIG(target);
  TORCH_CHECK(logits.dim() == 2 && target.scalar_type() == torch::kInt64 && target.numel() == logits.size(0));
  TORCH_CHECK(logits.stride(1) == 1 && logits.stride(0) >= logits.size(1), "masked_ce: rows must be dense");
  const int ld = (int)logits.stride(0);
  const bool bf = logits.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(bf || logits.scalar_type() == torch::kFloat32);
  const int rows = logits.size(0), V = logits.size(1);
  const unsigned char* mp = nullptr;
  if (mask.has_value()) {
    TORCH_CHECK(mask->is_cuda() && mask->scalar_type() == torch::kUInt8 && mask->numel() >= V);
    mp = mask->data_ptr<unsigned char>();
  }
  double* mt = nullptr;
  if (meter.has_value()) { TORCH_CHECK(meter->scalar_type() == torch::kFloat64 && meter->numel() >= 4); mt = meter->data_ptr<double>(); }
  c10::cuda::CUDAGuard guard(logits.device());
  auto fopt = torch::TensorOptions().dtype(torch::kFloat32).device(logits.device());
  auto loss = torch::empty({1}, fopt);
  auto rows_buf = torch::empty({2, rows}, fopt);
  torch::Tensor dl = need_grad ? torch::empty_strided(logits.sizes(), logits.strides(), logits.options()) : torch::Tensor();
  hfl_masked_ce(logits.data_ptr(), bf ? 1 : 0, (const long long*)target.data_ptr<int64_t>(), mp, rows, V, ld,
                need_grad ? dl.data_ptr() : nullptr, rows_buf.data_ptr<float>(), rows_buf.data_ptr<float>() + rows,
                (float)n_weight, loss.data_ptr<float>(), mt, cur_stream());
  check_launch();
  return {loss, dl};
}

// ---- data ------------------------------------------------------------------------------------------------
torch::Tensor augment(torch::Tensor src, c10::optional<torch::Tensor> idx, c10::optional<torch::Tensor> params,
                      torch::Tensor mean, torch::Tensor inv_std, int64_t pad, bool out_bf16) {
  CHECK_CUDA_T(src); CHECK_CONTIG(src);
  TORCH_CHECK(src.scalar_type() == torch::kUInt8 && src.dim() == 4);
  const int H = src.size(1), W = src.size(2), C = src.size(3);
  const int B = idx.has_value() ? idx->numel() : src.size(0);
  if (idx.has_value()) TORCH_CHECK(idx->is_cuda() && idx->scalar_type() == torch::kInt64 && idx->is_contiguous());
  if (params.has_value()) TORCH_CHECK(params->is_cuda() && params->scalar_type() == torch::kInt32 && params->is_contiguous() && params->numel() >= 3 * B);
  TORCH_CHECK(mean.numel() == C && inv_std.numel() == C && mean.scalar_type() == torch::kFloat32);
  c10::cuda::CUDAGuard guard(src.device());
  auto out = torch::empty({B, C, H, W}, torch::TensorOptions().dtype(out_bf16 ? torch::kBFloat16 : torch::kFloat32)
                                           .device(src.device()).memory_format(torch::MemoryFormat::ChannelsLast));
  hfl_augment(src.data_ptr<unsigned char>(), idx.has_value() ? (const long long*)idx->data_ptr<int64_t>() : nullptr,
              params.has_value() ? params->data_ptr<int>() : nullptr, mean.data_ptr<float>(),
              inv_std.data_ptr<float>(), B, H, W, C, (int)pad, out.data_ptr(), out_bf16 ? 1 : 0, cur_stream());
  check_launch();
  return out;
}

// ---- fused norm ------------------------------------------------------------------------------------------
static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
static const void* optp(const c10::optional<torch::Tensor>& t) { return t.has_value() ? t->data_ptr() : nullptr; }

// x, res: NHWC-dense bf16 activations (any logical shape with C the fastest physical dim); returns
// [s (or undefined), y, mean[G,C], invstd[G,C], var_unbiased[G,C]]
std::vector<torch::Tensor> bn_act_fwd(torch::Tensor x, c10::optional<torch::Tensor> res, torch::Tensor gamma,
                                      torch::Tensor beta, double eps, int64_t groups, int64_t C, bool relu,
                                      c10::optional<torch::Tensor> ws_persistent,
                                      c10::optional<torch::Tensor> nvalid, int64_t samples) {
  // nvalid (int32 [1], device): only the first *nvalid of `samples` samples are real (partial last batch of an epoch
  // inside a captured step); statistics skip the rest and s / y are zero there
  CHECK_CUDA_T(x);
  TORCH_CHECK(x.scalar_type() == torch::kBFloat16, "bn_act_fwd: bf16 activations only");
  TORCH_CHECK(is_pow2(C) && C >= 4 && C <= 2048, "bn_act_fwd: C must be a power of two in [4, 2048]");
  const int64_t M = x.numel() / C;
  TORCH_CHECK(M % groups == 0 && ((M / groups) * C) % 8 == 0);
  const bool pb = gamma.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(pb || gamma.scalar_type() == torch::kFloat32);
  TORCH_CHECK(gamma.is_contiguous() && beta.is_contiguous() && beta.scalar_type() == gamma.scalar_type());
  c10::cuda::CUDAGuard guard(x.device());
  auto y = torch::empty_like(x);
  torch::Tensor s;
  if (res.has_value()) { TORCH_CHECK(res->numel() == x.numel() && res->scalar_type() == torch::kBFloat16); s = torch::empty_like(x); }
  auto fopt = torch::TensorOptions().dtype(torch::kFloat32).device(x.device());
  auto stats = torch::empty({3, groups, C}, fopt);
  float* sp = stats.data_ptr<float>();
  // scratch of the split (wide) path: per call, or a caller-owned zero-initialised fp32 [2C+1] slab that the apply
  // kernel re-zeroes itself (no memset node in captured graphs)
  const bool wp = ws_persistent.has_value() && groups == 1;
  if (wp) TORCH_CHECK(ws_persistent->scalar_type() == torch::kFloat32 && ws_persistent->numel() >= 2 * C + 1);
  auto ws = wp ? *ws_persistent : torch::empty({groups, 2, C}, fopt);
  const int* nvp = nullptr;
  int rps = 0;
  if (nvalid.has_value()) {
    TORCH_CHECK(nvalid->scalar_type() == torch::kInt32 && nvalid->is_cuda() && groups == 1 && samples > 0 &&
                M % samples == 0 && ((M / samples) * C) % 8 == 0, "bn_act_fwd: bad nvalid / samples");
    nvp = nvalid->data_ptr<int>();
    rps = (int)(M / samples);
  }
  int e = hfl_bn_act_fwd(x.data_ptr(), optp(res), res.has_value() ? s.data_ptr() : nullptr, y.data_ptr(),
                         gamma.data_ptr(), beta.data_ptr(), pb ? 1 : 0, (float)eps, (int)groups, (int)(M / groups),
                         (int)C, relu ? 1 : 0, sp, sp + groups * C, sp + 2 * groups * C, ws.data_ptr<float>(), wp ? 1 : 0,
                         nvp, rps, cur_stream());
  TORCH_CHECK(e == 0, "bn_act_fwd launch failed: ", cudaGetErrorString((cudaError_t)e));
  check_launch();
  return {s, y, stats[0], stats[1], stats[2]};
}

// y = relu(bn(s)) where the per-channel (sum, sum of squares) of s were accumulated into `ws` (fp32 [2C+1], self-cleaning)
// by the producing convolution's epilogue.  Returns [y, mean[1,C], invstd[1,C], var_unbiased[1,C]].
std::vector<torch::Tensor> bn_apply_sums(torch::Tensor s, torch::Tensor gamma, torch::Tensor beta, double eps, int64_t C,
                                         bool relu, torch::Tensor ws, c10::optional<torch::Tensor> nvalid,
                                         int64_t samples) {
  CHECK_CUDA_T(s);
  TORCH_CHECK(s.scalar_type() == torch::kBFloat16 && is_pow2(C) && C >= 8 && C <= 2048);
  const int64_t M = s.numel() / C;
  TORCH_CHECK((M * C) % 8 == 0);
  const bool pb = gamma.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(pb || gamma.scalar_type() == torch::kFloat32);
  TORCH_CHECK(gamma.is_contiguous() && beta.is_contiguous() && beta.scalar_type() == gamma.scalar_type());
  TORCH_CHECK(ws.scalar_type() == torch::kFloat32 && ws.numel() >= 2 * C + 1 && ws.is_contiguous());
  c10::cuda::CUDAGuard guard(s.device());
  auto y = torch::empty_like(s);
  auto stats = torch::empty({3, 1, C}, torch::TensorOptions().dtype(torch::kFloat32).device(s.device()));
  float* sp = stats.data_ptr<float>();
  const int* nvp = nullptr;
  int rps = 0;
  if (nvalid.has_value()) {
    TORCH_CHECK(nvalid->scalar_type() == torch::kInt32 && nvalid->is_cuda() && samples > 0 && M % samples == 0 &&
                ((M / samples) * C) % 8 == 0, "bn_apply_sums: bad nvalid / samples");
    nvp = nvalid->data_ptr<int>();
    rps = (int)(M / samples);
  }
  int e = hfl_bn_apply_sums(s.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), pb ? 1 : 0, (float)eps, (int)M,
                            (int)C, relu ? 1 : 0, ws.data_ptr<float>(), sp, sp + C, sp + 2 * C, nvp, rps, cur_stream());
  TORCH_CHECK(e == 0, "bn_apply_sums launch failed: ", cudaGetErrorString((cudaError_t)e));
  check_launch();
  return {y, stats[0], stats[1], stats[2]};
}

std::vector<torch::Tensor> affine_act(torch::Tensor x, c10::optional<torch::Tensor> res, torch::Tensor a,
                                      torch::Tensor b, int64_t C, bool relu) {
  CHECK_CUDA_T(x);
  TORCH_CHECK(x.scalar_type() == torch::kBFloat16 && x.numel() % 8 == 0 && is_pow2(C));
  TORCH_CHECK(a.scalar_type() == torch::kFloat32 && b.scalar_type() == torch::kFloat32 && a.numel() == C);
  c10::cuda::CUDAGuard guard(x.device());
  auto y = torch::empty_like(x);
  torch::Tensor s;
  if (res.has_value()) s = torch::empty_like(x);
  hfl_affine_act(x.data_ptr(), optp(res), res.has_value() ? s.data_ptr() : nullptr, y.data_ptr(),
                 a.data_ptr<float>(), b.data_ptr<float>(), (long)x.numel(), (int)C, relu ? 1 : 0, cur_stream());
  check_launch();
  return {s, y};
}

// returns ds; writes dgamma / dbeta in place (views into the gradient arena)
torch::Tensor bn_act_bwd(torch::Tensor dy, c10::optional<torch::Tensor> dy2, c10::optional<torch::Tensor> dres,
                         torch::Tensor s, torch::Tensor mean, torch::Tensor invstd, torch::Tensor gamma,
                         torch::Tensor beta, torch::Tensor dgamma, torch::Tensor dbeta, int64_t C, int64_t hw,
                         double dy_scale, bool relu, c10::optional<torch::Tensor> ws_persistent,
                         c10::optional<torch::Tensor> nvalid, int64_t samples) {
  CHECK_CUDA_T(dy); CHECK_CUDA_T(s);
  TORCH_CHECK(s.scalar_type() == torch::kBFloat16 && dy.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(is_pow2(C) && C >= 4 && C <= 2048);
  const int64_t M = s.numel() / C;
  TORCH_CHECK((M * C) % 8 == 0);
  if (hw > 0) {
    TORCH_CHECK(dy.numel() * hw == s.numel() && (hw * C) % 8 == 0);
  } else {
    TORCH_CHECK(dy.numel() == s.numel());
  }
  const bool pb = gamma.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(dgamma.scalar_type() == gamma.scalar_type() && dbeta.scalar_type() == gamma.scalar_type());
  TORCH_CHECK(dgamma.is_contiguous() && dbeta.is_contiguous() && mean.scalar_type() == torch::kFloat32);
  c10::cuda::CUDAGuard guard(s.device());
  auto ds = torch::empty_like(s);
  const bool wp = ws_persistent.has_value();
  if (wp) TORCH_CHECK(ws_persistent->scalar_type() == torch::kFloat32 && ws_persistent->numel() >= 2 * C + 1);
  auto ws = wp ? *ws_persistent : torch::empty({2, C}, torch::TensorOptions().dtype(torch::kFloat32).device(s.device()));
  const int* nvp = nullptr;
  int rps = 0;
  if (nvalid.has_value()) {
    TORCH_CHECK(nvalid->scalar_type() == torch::kInt32 && nvalid->is_cuda() && samples > 0 && M % samples == 0 &&
                ((M / samples) * C) % 8 == 0, "bn_act_bwd: bad nvalid / samples");
    nvp = nvalid->data_ptr<int>();
    rps = (int)(M / samples);
  }
  int e = hfl_bn_act_bwd(dy.data_ptr(), optp(dy2), optp(dres), s.data_ptr(), mean.data_ptr<float>(),
                         invstd.data_ptr<float>(), gamma.data_ptr(), beta.data_ptr(), pb ? 1 : 0, (int)M, (int)C,
                         (int)hw, (float)dy_scale, relu ? 1 : 0, ds.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                         ws.data_ptr<float>(), wp ? 1 : 0, nvp, rps, cur_stream());
  TORCH_CHECK(e == 0, "bn_act_bwd launch failed: ", cudaGetErrorString((cudaError_t)e));
  check_launch();
  return ds;
}

// ---- MXFP8 block-scaled GEMM (csrc/gemm_mxfp8_tcgen05.cu) --------------------------------------------------
// x: bf16 [R, K] (K % 128 == 0, K contiguous) -> [q uint8 [R, K] (e4m3 bit patterns), sf uint8 [tiles * 512] (ue8m0)]
std::vector<torch::Tensor> mx_quantize(torch::Tensor x) {
  CHECK_CUDA_T(x);
  TORCH_CHECK(x.dim() == 2 && x.scalar_type() == torch::kBFloat16 && x.stride(1) == 1, "mx_quantize: bf16 [R, K]");
  const int64_t R = x.size(0), K = x.size(1);
  TORCH_CHECK(K % 128 == 0 && x.stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15) == 0,
              "mx_quantize: K % 128 == 0 and 16-byte aligned rows required");
  c10::cuda::CUDAGuard guard(x.device());
  auto bopt = torch::TensorOptions().dtype(torch::kUInt8).device(x.device());
  auto q = torch::empty({R, K}, bopt);
  auto sf = torch::zeros({(int64_t)hfl_mx_sf_bytes((int)R, (int)K)}, bopt);
  int e = hfl_mx_quantize(x.data_ptr(), (int)R, (int)K, (int)x.stride(0), q.data_ptr(), (int)K, sf.data_ptr(),
                          cur_stream());
  TORCH_CHECK(e == 0, "mx_quantize failed: ", e);
  return {q, sf};
}

// C[M, N] = alpha * dequant(aq, sfa) . dequant(bq, sfb)^T (+ bias fp32 [N]); bf16 or fp32 output
torch::Tensor gemm_mxfp8(torch::Tensor aq, torch::Tensor sfa, torch::Tensor bq, torch::Tensor sfb,
                         c10::optional<torch::Tensor> bias, double alpha, bool out_fp32) {
  CHECK_CUDA_T(aq); CHECK_CUDA_T(bq); CHECK_CUDA_T(sfa); CHECK_CUDA_T(sfb);
  TORCH_CHECK(aq.dim() == 2 && bq.dim() == 2 && aq.size(1) == bq.size(1) && aq.scalar_type() == torch::kUInt8 &&
              bq.scalar_type() == torch::kUInt8 && aq.is_contiguous() && bq.is_contiguous(), "gemm_mxfp8: operands");
  const int64_t M = aq.size(0), N = bq.size(0), K = aq.size(1);
  TORCH_CHECK(K % 128 == 0 && sfa.numel() == hfl_mx_sf_bytes((int)M, (int)K) &&
              sfb.numel() == hfl_mx_sf_bytes((int)N, (int)K), "gemm_mxfp8: scale-factor buffers do not match");
  const float* bp = nullptr;
  if (bias.has_value()) {
    TORCH_CHECK(bias->scalar_type() == torch::kFloat32 && bias->numel() == N && bias->is_contiguous());
    bp = bias->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(aq.device());
  auto C = torch::empty({M, N}, torch::TensorOptions().dtype(out_fp32 ? torch::kFloat32 : torch::kBFloat16).device(aq.device()));
  int e = hfl_gemm_mxfp8(aq.data_ptr(), (int)K, sfa.data_ptr(), bq.data_ptr(), (int)K, sfb.data_ptr(), C.data_ptr(),
                         (int)N, out_fp32 ? 0 : 1, bp, (float)alpha, (int)M, (int)N, (int)K, cur_stream());
  TORCH_CHECK(e == 0, "gemm_mxfp8 failed: ", e);
  return C;
}

// ---- 2x2 max pooling (csrc/pool_kernels.cu) ---------------------------------------------------------------
static void check_nhwc_bf16(const torch::Tensor& t, const char* what) {
  CHECK_CUDA_T(t);
  TORCH_CHECK(t.dim() == 4 && t.scalar_type() == torch::kBFloat16 && t.is_contiguous(torch::MemoryFormat::ChannelsLast),
              what, ": channels-last bf16 [N, C, H, W] expected");
}
torch::Tensor maxpool2_fwd(torch::Tensor x) {
  check_nhwc_bf16(x, "maxpool2_fwd");
  const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
  TORCH_CHECK(H >= 2 && W >= 2);
  c10::cuda::CUDAGuard guard(x.device());
  auto y = torch::empty({N, C, H / 2, W / 2}, x.options().memory_format(torch::MemoryFormat::ChannelsLast));
  int e = hfl_maxpool2_fwd(x.data_ptr(), y.data_ptr(), N, H, W, C, cur_stream());
  TORCH_CHECK(e == 0, "maxpool2_fwd launch failed: ", e);
  return y;
}
torch::Tensor maxpool2_bwd(torch::Tensor dy, torch::Tensor x) {
  check_nhwc_bf16(x, "maxpool2_bwd");
  check_nhwc_bf16(dy, "maxpool2_bwd");
  const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
  TORCH_CHECK(dy.size(0) == N && dy.size(1) == C && dy.size(2) == H / 2 && dy.size(3) == W / 2);
  c10::cuda::CUDAGuard guard(x.device());
  auto dx = torch::empty_like(x);
  int e = hfl_maxpool2_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), N, H, W, C, cur_stream());
  TORCH_CHECK(e == 0, "maxpool2_bwd launch failed: ", e);
  return dx;
}

// ---- fused GroupNorm / InstanceNorm / LayerNorm (csrc/gn_kernels.cu) ---------------------------------------
// x, res: NHWC-dense bf16 [N, HW, C]; returns [s (or undefined), y, mean[N,G], rstd[N,G]]
std::vector<torch::Tensor> gn_act_fwd(torch::Tensor x, c10::optional<torch::Tensor> res, torch::Tensor gamma,
                                      torch::Tensor beta, double eps, int64_t N, int64_t C, int64_t G, bool relu) {
  CHECK_CUDA_T(x);
  TORCH_CHECK(x.scalar_type() == torch::kBFloat16, "gn_act_fwd: bf16 activations only");
  TORCH_CHECK(G >= 1 && C % G == 0 && x.numel() % (N * C) == 0);
  const bool pb = gamma.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(pb || gamma.scalar_type() == torch::kFloat32);
  TORCH_CHECK(gamma.is_contiguous() && beta.is_contiguous() && beta.scalar_type() == gamma.scalar_type());
  c10::cuda::CUDAGuard guard(x.device());
  auto y = torch::empty_like(x);
  torch::Tensor s;
  if (res.has_value()) { TORCH_CHECK(res->numel() == x.numel() && res->scalar_type() == torch::kBFloat16); s = torch::empty_like(x); }
  auto stats = torch::empty({2, N, G}, torch::TensorOptions().dtype(torch::kFloat32).device(x.device()));
  float* sp = stats.data_ptr<float>();
  int e = hfl_gn_act_fwd(x.data_ptr(), optp(res), gamma.data_ptr(), beta.data_ptr(), pb ? 1 : 0,
                         res.has_value() ? s.data_ptr() : nullptr, y.data_ptr(), sp, sp + N * G, (int)N,
                         (int)(x.numel() / (N * C)), (int)C, (int)G, (float)eps, relu ? 1 : 0, cur_stream());
  TORCH_CHECK(e == 0, "gn_act_fwd launch failed: ", e);
  check_launch();
  return {s, y, stats[0], stats[1]};
}

// returns ds; writes dgamma / dbeta (dtype of gamma) in place
torch::Tensor gn_act_bwd(torch::Tensor dy, c10::optional<torch::Tensor> dy2, c10::optional<torch::Tensor> dres,
                         torch::Tensor s, torch::Tensor mean, torch::Tensor rstd, torch::Tensor gamma,
                         torch::Tensor beta, torch::Tensor dgamma, torch::Tensor dbeta, int64_t N, int64_t C, int64_t G,
                         int64_t hw, double dy_scale, bool relu) {
  CHECK_CUDA_T(dy); CHECK_CUDA_T(s);
  TORCH_CHECK(s.scalar_type() == torch::kBFloat16 && dy.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(G >= 1 && C % G == 0 && s.numel() % (N * C) == 0 && mean.numel() == N * G && rstd.numel() == N * G);
  if (hw > 0) {
    TORCH_CHECK(dy.numel() * hw == s.numel());
  } else {
    TORCH_CHECK(dy.numel() == s.numel());
  }
  const bool pb = gamma.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(dgamma.scalar_type() == gamma.scalar_type() && dbeta.scalar_type() == gamma.scalar_type());
  TORCH_CHECK(dgamma.is_contiguous() && dbeta.is_contiguous() && mean.scalar_type() == torch::kFloat32 &&
              mean.is_contiguous() && rstd.is_contiguous());
  c10::cuda::CUDAGuard guard(s.device());
  auto ds = torch::empty_like(s);
  auto part = torch::empty({N, 2, C}, torch::TensorOptions().dtype(torch::kFloat32).device(s.device()));
  int e = hfl_gn_act_bwd(dy.data_ptr(), optp(dy2), optp(dres), s.data_ptr(), mean.data_ptr<float>(),
                         rstd.data_ptr<float>(), gamma.data_ptr(), beta.data_ptr(), pb ? 1 : 0, ds.data_ptr(),
                         part.data_ptr<float>(), dgamma.data_ptr(), dbeta.data_ptr(), (int)N,
                         (int)(s.numel() / (N * C)), (int)C, (int)G, relu ? 1 : 0, (int)hw, (float)dy_scale,
                         cur_stream());
  TORCH_CHECK(e == 0, "gn_act_bwd launch failed: ", e);
  check_launch();
  return ds;
}

void multi_sgd_clip_step(torch::Tensor segs, torch::Tensor blocks, torch::Tensor w, c10::optional<torch::Tensor> wc,
                         torch::Tensor mom, torch::Tensor lr, double mu, double wd, double max_norm,
                         torch::Tensor scratch) {
  CHECK_CUDA_T(segs); CHECK_CUDA_T(blocks); CHECK_CUDA_T(w);
  TORCH_CHECK(segs.scalar_type() == torch::kInt64 && segs.dim() == 2 && segs.size(1) == 4 && segs.is_contiguous());
  TORCH_CHECK(blocks.scalar_type() == torch::kInt32 && blocks.dim() == 2 && blocks.size(1) == 2);
  c10::cuda::CUDAGuard guard(w.device());
  hfl_multi_sgd_clip_step(segs.data_ptr(), blocks.data_ptr(), (int)blocks.size(0), w.data_ptr<float>(),
                          wc.has_value() ? wc->data_ptr() : nullptr, mom.data_ptr<float>(), lr.data_ptr<float>(),
                          (float)mu, (float)wd, (float)max_norm, scratch.data_ptr<float>(), cur_stream());
  check_launch();
}

// row-sparse embedding gradient (csrc/optim_kernels.cu): src int32 [M] (table row per token), g fp32 [rows, E]
void emb_sparse_norm(torch::Tensor src, torch::Tensor g, torch::Tensor scratch) {
  CHECK_CUDA_T(src); CHECK_CUDA_T(g);
  TORCH_CHECK(src.scalar_type() == torch::kInt32 && src.is_contiguous() && g.scalar_type() == torch::kFloat32 &&
              g.dim() == 2 && g.is_contiguous() && scratch.scalar_type() == torch::kFloat32);
  c10::cuda::CUDAGuard guard(g.device());
  hfl_emb_sparse_norm(src.data_ptr<int>(), (int)src.numel(), g.data_ptr<float>(), (int)g.size(1), scratch.data_ptr<float>(),
                      cur_stream());
  check_launch();
}
void emb_sparse_apply(torch::Tensor src, torch::Tensor g, int64_t off, torch::Tensor w, c10::optional<torch::Tensor> wc,
                      torch::Tensor mom, torch::Tensor lr, double max_norm, torch::Tensor scratch) {
  CHECK_CUDA_T(src); CHECK_CUDA_T(g); CHECK_CUDA_T(w);
  TORCH_CHECK(src.scalar_type() == torch::kInt32 && src.is_contiguous() && g.scalar_type() == torch::kFloat32 &&
              g.dim() == 2 && g.is_contiguous() && w.scalar_type() == torch::kFloat32 && mom.scalar_type() == torch::kFloat32);
  TORCH_CHECK(off >= 0 && off + g.numel() <= w.numel());
  c10::cuda::CUDAGuard guard(g.device());
  hfl_emb_sparse_apply(src.data_ptr<int>(), (int)src.numel(), g.data_ptr<float>(), (int)g.size(1), (long)off,
                       w.data_ptr<float>(), wc.has_value() ? wc->data_ptr() : nullptr, mom.data_ptr<float>(),
                       lr.data_ptr<float>(), (float)max_norm, scratch.data_ptr<float>(), cur_stream());
  check_launch();
}

// ---- classifier head (pool + linear + masked CE, forward and backward) ----------------------------------
// y4: [N, C, H, W] channels-last bf16.  Returns [loss, dW, db, dpooled[N, C] (bf16)]
std::vector<torch::Tensor> head_fwd_bwd(torch::Tensor y4, torch::Tensor W, torch::Tensor bias, torch::Tensor target,
                                        c10::optional<torch::Tensor> mask, double n_weight,
                                        c10::optional<torch::Tensor> meter, c10::optional<torch::Tensor> nvalid) {
  CHECK_CUDA_T(y4);
  TORCH_CHECK(y4.dim() == 4 && y4.scalar_type() == torch::kBFloat16 && y4.is_contiguous(torch::MemoryFormat::ChannelsLast));
  const int N = y4.size(0), C = y4.size(1), HW = y4.size(2) * y4.size(3), Kc = W.size(0);
  TORCH_CHECK(W.is_contiguous() && W.size(1) == C && bias.numel() == Kc && target.numel() == N);
  const bool pb = W.scalar_type() == torch::kBFloat16;
  TORCH_CHECK((pb || W.scalar_type() == torch::kFloat32) && bias.scalar_type() == W.scalar_type());
  TORCH_CHECK(hfl_head_smem_bytes(N, C, Kc) <= 200 * 1024, "head_fwd_bwd: problem too large for one CTA");
  const unsigned char* mp = nullptr;
  if (mask.has_value()) { TORCH_CHECK(mask->scalar_type() == torch::kUInt8 && mask->numel() >= Kc); mp = mask->data_ptr<unsigned char>(); }
  double* mt = nullptr;
  if (meter.has_value()) { TORCH_CHECK(meter->scalar_type() == torch::kFloat64 && meter->numel() >= 4); mt = meter->data_ptr<double>(); }
  c10::cuda::CUDAGuard guard(y4.device());
  auto loss = torch::empty({1}, torch::TensorOptions().dtype(torch::kFloat32).device(y4.device()));
  auto dW = torch::empty_like(W);
  auto db = torch::empty_like(bias);
  auto dp = torch::empty({N, C}, torch::TensorOptions().dtype(torch::kBFloat16).device(y4.device()));
  auto pooled = torch::empty({N, C}, torch::TensorOptions().dtype(torch::kFloat32).device(y4.device()));
  int e = hfl_head_fwd_bwd(y4.data_ptr(), N, HW, C, W.data_ptr(), bias.data_ptr(), pb ? 1 : 0, Kc,
                           (const long long*)target.data_ptr<int64_t>(), mp, (float)n_weight, loss.data_ptr<float>(), mt,
                           dW.data_ptr(), db.data_ptr(), dp.data_ptr(), pooled.data_ptr<float>(),
                           nvalid.has_value() ? nvalid->data_ptr<int>() : nullptr, cur_stream());
  TORCH_CHECK(e == 0, "head_fwd_bwd failed: ", e);
  return {loss, dW, db, dp};
}

// ---- tcgen05 GEMM ----------------------------------------------------------------------------------------
// y[M,N] = act(alpha * x[M,K] @ w[N,K]^T + bias)   (x, w bf16 with K contiguous; bias fp32; y bf16 or fp32)
torch::Tensor gemm_bf16_tn(torch::Tensor x, torch::Tensor w, c10::optional<torch::Tensor> bias, double alpha,
                           bool gelu, bool out_fp32, bool bias_inside, c10::optional<torch::Tensor> out) {
  CHECK_CUDA_T(x); CHECK_CUDA_T(w);
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1), "gemm_bf16_tn: shape mismatch");
  TORCH_CHECK(x.scalar_type() == torch::kBFloat16 && w.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(x.stride(1) == 1 && w.stride(1) == 1, "gemm_bf16_tn: K must be contiguous");
  const int64_t M = x.size(0), K = x.size(1), N = w.size(0);
  TORCH_CHECK(x.stride(0) % 8 == 0 && w.stride(0) % 8 == 0, "gemm_bf16_tn: 16-byte row alignment");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(x.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(w.data_ptr()) & 15) == 0);
  const float* bp = nullptr;
  const void* bhp = nullptr;
  torch::Tensor bf;
  if (bias.has_value()) {
    TORCH_CHECK(bias->numel() == N && bias->is_contiguous());
    if (bias_inside) {               // bf16 bias read in place, y = alpha * (x w^T + b)
      TORCH_CHECK(bias->scalar_type() == torch::kBFloat16, "gemm_bf16_tn: bias_inside needs a bf16 bias");
      bhp = bias->data_ptr();
    } else {
      bf = bias->scalar_type() == torch::kFloat32 ? *bias : bias->to(torch::kFloat32);
      bp = bf.data_ptr<float>();
    }
  }
  c10::cuda::CUDAGuard guard(x.device());
  torch::Tensor y;
  if (out.has_value()) {             // caller-provided destination (may be a row-padded view: ldc = stride(0))
    y = *out;
    TORCH_CHECK(y.is_cuda() && y.dim() == 2 && y.size(0) == M && y.size(1) == N && y.stride(1) == 1 &&
                y.scalar_type() == (out_fp32 ? torch::kFloat32 : torch::kBFloat16));
  } else {
    y = torch::empty({M, N}, torch::TensorOptions().dtype(out_fp32 ? torch::kFloat32 : torch::kBFloat16).device(x.device()));
  }
  // big problems: 2-CTA clusters with TMA multicast of the shared operand + 128x256 tiles (csrc/gemm_mc_tcgen05.cu)
  static const bool mc_enabled = []() { const char* e = std::getenv("HFL_GEMM_MC"); return !(e && e[0] == '0'); }();
  if (mc_enabled && !out_fp32 && hfl_gemm_mc_wanted((int)M, (int)N, (int)K)) {
    int e2 = hfl_gemm_bf16_tn_mc(x.data_ptr(), (int)x.stride(0), w.data_ptr(), (int)w.stride(0), y.data_ptr(),
                                 (int)y.stride(0), bp, bhp, (float)alpha, gelu ? 1 : 0, (int)M, (int)N, (int)K, cur_stream());
    TORCH_CHECK(e2 == 0, "gemm_bf16_tn_mc failed with code ", e2);
    return y;
  }
  int e = hfl_gemm_bf16_tn(x.data_ptr(), (int)x.stride(0), w.data_ptr(), (int)w.stride(0), y.data_ptr(), (int)y.stride(0),
                           out_fp32 ? 0 : 1, bp, bhp, (float)alpha, gelu ? 1 : 0, (int)M, (int)N, (int)K, cur_stream());
  TORCH_CHECK(e == 0, "gemm_bf16_tn failed with code ", e);
  return y;
}

// C[M,N] = alpha * op(A) . op(B)^T with explicit layouts: a is [M,K] (a_mn=false) or [K,M] (a_mn=true); b is [N,K]
// (b_mn=false) or [K,N] (b_mn=true); all row-major with the last dim contiguous.
torch::Tensor gemm_bf16_general(torch::Tensor a, bool a_mn, torch::Tensor b, bool b_mn, double alpha, bool out_fp32) {
  CHECK_CUDA_T(a); CHECK_CUDA_T(b);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.scalar_type() == torch::kBFloat16 && b.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(a.stride(1) == 1 && b.stride(1) == 1);
  const int64_t M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0), Kb = b_mn ? b.size(0) : b.size(1);
  TORCH_CHECK(K == Kb, "gemm_bf16_general: reduction dims differ");
  TORCH_CHECK(a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0, "gemm_bf16_general: 16-byte row strides");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(a.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(b.data_ptr()) & 15) == 0);
  c10::cuda::CUDAGuard guard(a.device());
  // skinny output + long reduction: split-K into a zero-filled fp32 buffer (atomics), cast afterwards if bf16 is wanted
  const int splits = hfl_gemm_general_splits((int)M, (int)N, (int)K);
  const bool direct = splits == 1;
  auto y = direct ? torch::empty({M, N}, torch::TensorOptions().dtype(out_fp32 ? torch::kFloat32 : torch::kBFloat16).device(a.device()))
                  : torch::zeros({M, N}, torch::TensorOptions().dtype(torch::kFloat32).device(a.device()));
  int e = hfl_gemm_bf16_general(a.data_ptr(), (int)a.stride(0), a_mn ? 1 : 0, b.data_ptr(), (int)b.stride(0), b_mn ? 1 : 0,
                                y.data_ptr(), (int)N, (direct && !out_fp32) ? 1 : 0, (float)alpha, (int)M, (int)N, (int)K,
                                splits, cur_stream());
  TORCH_CHECK(e == 0, "gemm_bf16_general failed with code ", e);
  if (!direct && !out_fp32) return y.to(torch::kBFloat16);
  return y;
}


// ---- fused masked-LM transformer pieces (csrc/lm_kernels.cu) -----------------------------------------------
static const long long* i64p(const torch::Tensor& t) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kInt64 && t.numel() >= 1);
  return reinterpret_cast<const long long*>(t.data_ptr<int64_t>());
}
static void check_bf(const torch::Tensor& t, const char* what) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kBFloat16, what, ": contiguous CUDA bf16 expected");
}
static float* f32p(const torch::Tensor& t, int64_t n) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32 && t.is_contiguous() && t.numel() == n);
  return t.data_ptr<float>();
}
struct DyList {
  const void* p[4];
  int n;
  DyList(const std::vector<torch::Tensor>& dys, int64_t numel, bool allow_empty = false) {
    TORCH_CHECK((allow_empty || !dys.empty()) && dys.size() <= 4, "1..4 gradient branches");
    n = (int)dys.size();
    for (int i = 0; i < 4; ++i) p[i] = nullptr;
    for (int i = 0; i < n; ++i) { check_bf(dys[i], "dy"); TORCH_CHECK(dys[i].numel() == numel); p[i] = dys[i].data_ptr(); }
  }
};

std::vector<torch::Tensor> lm_embed_fwd(torch::Tensor tok, torch::Tensor emb, torch::Tensor pos, torch::Tensor gamma,
                                        torch::Tensor beta, int64_t S, int64_t mask_id, double eps, double p_mask,
                                        double p_drop, torch::Tensor seed, torch::Tensor step, int64_t site) {
  check_bf(emb, "emb"); check_bf(pos, "pos"); check_bf(gamma, "gamma"); check_bf(beta, "beta");
  TORCH_CHECK(tok.is_cuda() && tok.scalar_type() == torch::kInt64 && tok.is_contiguous());
  const int M = tok.numel(), E = emb.size(1);
  TORCH_CHECK(E <= hfl_lm_max_e() && pos.size(1) == E && pos.size(0) >= S && gamma.numel() == E && mask_id < emb.size(0));
  c10::cuda::CUDAGuard guard(tok.device());
  auto ob = emb.options();
  auto out = torch::empty({M, E}, ob), xhat = torch::empty({M, E}, ob);
  auto src = torch::empty({M}, ob.dtype(torch::kInt32));
  auto rstd = torch::empty({M}, ob.dtype(torch::kFloat32));
  auto dmask = torch::empty({p_drop > 0 ? M : 0, E}, ob.dtype(torch::kUInt8));
  hfl_lm_embed_fwd(i64p(tok), emb.data_ptr(), pos.data_ptr(), gamma.data_ptr(), beta.data_ptr(), M, (int)S, E,
                   (int)mask_id, (float)eps, (float)p_mask, (float)p_drop, i64p(seed), i64p(step), (unsigned)site,
                   out.data_ptr(), src.data_ptr<int>(), xhat.data_ptr(), rstd.data_ptr<float>(),
                   p_drop > 0 ? dmask.data_ptr<uint8_t>() : nullptr, cur_stream());
  check_launch();
  return {out, src, xhat, rstd, dmask};
}

void lm_embed_bwd(std::vector<torch::Tensor> dys, torch::Tensor dmask, double p_drop, torch::Tensor xhat,
                  torch::Tensor rstd, torch::Tensor gamma, torch::Tensor src, int64_t S, torch::Tensor demb,
                  torch::Tensor dpos, torch::Tensor dgamma, torch::Tensor dbeta) {
  check_bf(xhat, "xhat"); check_bf(gamma, "gamma");
  const int M = xhat.size(0), E = xhat.size(1);
  DyList dl(dys, (int64_t)M * E);
  TORCH_CHECK(E <= hfl_lm_max_e() && demb.size(1) == E && dpos.size(1) == E && dpos.size(0) >= S);
  TORCH_CHECK(p_drop <= 0 || dmask.numel() == (int64_t)M * E);
  c10::cuda::CUDAGuard guard(xhat.device());
  hfl_lm_embed_bwd(dl.p, dl.n, p_drop > 0 ? dmask.data_ptr<uint8_t>() : nullptr, (float)p_drop, xhat.data_ptr(),
                   rstd.data_ptr<float>(), gamma.data_ptr(), src.data_ptr<int>(), M, (int)S, E,
                   f32p(demb, demb.numel()), f32p(dpos, dpos.numel()), f32p(dgamma, E), f32p(dbeta, E), cur_stream());
  check_launch();
}

std::vector<torch::Tensor> lm_add_ln_fwd(c10::optional<torch::Tensor> x, torch::Tensor h, torch::Tensor gamma,
                                         torch::Tensor beta, double eps, double p_drop, torch::Tensor seed,
                                         torch::Tensor step, int64_t site) {
  check_bf(h, "h"); check_bf(gamma, "gamma"); check_bf(beta, "beta");
  TORCH_CHECK(h.dim() == 2);
  const int M = h.size(0), E = h.size(1);
  TORCH_CHECK(E <= hfl_lm_max_e() && gamma.numel() == E);
  if (x.has_value()) { check_bf(*x, "x"); TORCH_CHECK(x->sizes() == h.sizes()); }
  c10::cuda::CUDAGuard guard(h.device());
  auto y = torch::empty_like(h), xhat = torch::empty_like(h);
  auto rstd = torch::empty({M}, h.options().dtype(torch::kFloat32));
  auto dmask = torch::empty({p_drop > 0 ? M : 0, E}, h.options().dtype(torch::kUInt8));
  hfl_lm_add_ln_fwd(x.has_value() ? x->data_ptr() : nullptr, h.data_ptr(), gamma.data_ptr(), beta.data_ptr(), M, E,
                    (float)eps, (float)p_drop, i64p(seed), i64p(step), (unsigned)site, y.data_ptr(), xhat.data_ptr(),
                    rstd.data_ptr<float>(), p_drop > 0 ? dmask.data_ptr<uint8_t>() : nullptr, cur_stream());
  check_launch();
  return {y, xhat, rstd, dmask};
}

// returns (dz, dh); dh aliases dz when there is no dropout
std::vector<torch::Tensor> lm_add_ln_bwd(std::vector<torch::Tensor> dys, torch::Tensor xhat, torch::Tensor rstd,
                                         torch::Tensor gamma, torch::Tensor dmask, double p_drop, torch::Tensor dgamma,
                                         torch::Tensor dbeta, c10::optional<torch::Tensor> dbias, double bias_alpha,
                                         c10::optional<torch::Tensor> dy_f32) {
  check_bf(xhat, "xhat"); check_bf(gamma, "gamma");
  const int M = xhat.size(0), E = xhat.size(1);
  DyList dl(dys, (int64_t)M * E, dy_f32.has_value());
  const float* dyf = nullptr;
  if (dy_f32.has_value()) {   // one more gradient branch, fp32 (the fused vocabulary head's dX accumulator)
    TORCH_CHECK(dy_f32->is_cuda() && dy_f32->scalar_type() == torch::kFloat32 && dy_f32->is_contiguous() &&
                dy_f32->numel() == (int64_t)M * E);
    dyf = dy_f32->data_ptr<float>();
  }
  TORCH_CHECK(E <= hfl_lm_max_e());
  TORCH_CHECK(p_drop <= 0 || dmask.numel() == (int64_t)M * E);
  c10::cuda::CUDAGuard guard(xhat.device());
  auto dz = torch::empty_like(xhat);
  auto dh = p_drop > 0 ? torch::empty_like(xhat) : dz;
  hfl_lm_add_ln_bwd(dl.p, dl.n, xhat.data_ptr(), rstd.data_ptr<float>(), gamma.data_ptr(),
                    p_drop > 0 ? dmask.data_ptr<uint8_t>() : nullptr, (float)p_drop, M, E, dz.data_ptr(),
                    p_drop > 0 ? dh.data_ptr() : nullptr, f32p(dgamma, E), f32p(dbeta, E),
                    dbias.has_value() ? f32p(*dbias, E) : nullptr, (float)bias_alpha, dyf, cur_stream());
  check_launch();
  return {dz, dh};
}

// ---- fused vocabulary head (csrc/lm_vocab_tcgen05.cu) ----------------------------------------------------------
bool lm_vocab_supported(int64_t M, int64_t V, int64_t E) { return hfl_lm_vocab_supported((int)M, (int)V, (int)E) == 0; }

// d: [M, E] bf16 (decoder LayerNorm output), W: [V, E] bf16, b: [V] bf16, target: [M] int64, mask: uint8 [V] (optional).
// Returns [loss (fp32 [1]), dW (bf16 [V, E]), dX (fp32 [M, E])]; dbias (fp32 [V]) is accumulated in place.  train = false:
// loss / meter only.
std::vector<torch::Tensor> lm_vocab_head(torch::Tensor d, torch::Tensor W, torch::Tensor b, torch::Tensor target,
                                         c10::optional<torch::Tensor> mask, double n_weight,
                                         c10::optional<torch::Tensor> meter, c10::optional<torch::Tensor> dbias, bool train) {
  check_bf(d, "d"); check_bf(W, "W"); check_bf(b, "b");
  TORCH_CHECK(d.dim() == 2 && W.dim() == 2 && d.size(1) == W.size(1) && b.numel() == W.size(0) && b.is_contiguous());
  TORCH_CHECK(d.stride(1) == 1 && W.stride(1) == 1 && d.stride(0) % 8 == 0 && W.stride(0) % 8 == 0);
  TORCH_CHECK(((uintptr_t)d.data_ptr() & 15) == 0 && ((uintptr_t)W.data_ptr() & 15) == 0);
  const int M = d.size(0), E = d.size(1), V = W.size(0);
  TORCH_CHECK(hfl_lm_vocab_supported(M, V, E) == 0, "lm_vocab_head: unsupported shape");
  TORCH_CHECK(target.is_cuda() && target.scalar_type() == torch::kInt64 && target.is_contiguous() && target.numel() == M);
  const unsigned char* mp = nullptr;
  if (mask.has_value()) {
    TORCH_CHECK(mask->is_cuda() && mask->scalar_type() == torch::kUInt8 && mask->numel() >= V);
    mp = mask->data_ptr<unsigned char>();
  }
  double* mt = nullptr;
  if (meter.has_value()) { TORCH_CHECK(meter->scalar_type() == torch::kFloat64 && meter->numel() >= 4); mt = meter->data_ptr<double>(); }
  c10::cuda::CUDAGuard guard(d.device());
  const int T = hfl_lm_vocab_tiles(V);
  auto fopt = torch::TensorOptions().dtype(torch::kFloat32).device(d.device());
  auto part = torch::empty({2, M, T}, fopt);
  auto parg = torch::empty({M, T}, fopt.dtype(torch::kInt32));
  auto rowbuf = torch::empty({2, M}, fopt);          // target logit, log-sum-exp
  auto loss = torch::empty({1}, fopt);
  torch::Tensor dW, dX;
  if (train) {
    dW = torch::empty({V, E}, d.options());
    dX = torch::empty({M, E}, fopt);
  }
  const long long* tp = (const long long*)target.data_ptr<int64_t>();
  int e = hfl_lm_vocab_fwd(d.data_ptr(), (int)d.stride(0), W.data_ptr(), (int)W.stride(0), b.data_ptr(), mp, tp, M, V, E,
                           part.data_ptr<float>(), part.data_ptr<float>() + (int64_t)M * T, parg.data_ptr<int>(),
                           rowbuf.data_ptr<float>(), rowbuf.data_ptr<float>() + M, (float)n_weight, loss.data_ptr<float>(),
                           mt, train ? dX.data_ptr<float>() : nullptr, train ? M * E : 0, cur_stream());
  TORCH_CHECK(e == 0, "lm_vocab_fwd failed with code ", e);
  if (train) {
    TORCH_CHECK(dbias.has_value() && dbias->scalar_type() == torch::kFloat32 && dbias->numel() >= V && dbias->is_contiguous());
    e = hfl_lm_vocab_bwd(d.data_ptr(), (int)d.stride(0), W.data_ptr(), (int)W.stride(0), b.data_ptr(), mp, tp,
                         rowbuf.data_ptr<float>() + M, M, V, E, dW.data_ptr(), E, dbias->data_ptr<float>(),
                         dX.data_ptr<float>(), cur_stream());
    TORCH_CHECK(e == 0, "lm_vocab_bwd failed with code ", e);
  }
  check_launch();
  return {loss, dW, dX};
}

std::vector<torch::Tensor> lm_gelu_drop_fwd(torch::Tensor u, double p_drop, torch::Tensor seed, torch::Tensor step,
                                            int64_t site) {
  check_bf(u, "u");
  c10::cuda::CUDAGuard guard(u.device());
  auto h = torch::empty_like(u);
  auto dmask = torch::empty({p_drop > 0 ? u.numel() : 0}, u.options().dtype(torch::kUInt8));
  hfl_lm_gelu_drop_fwd(u.data_ptr(), (long)u.numel(), (float)p_drop, i64p(seed), i64p(step), (unsigned)site,
                       h.data_ptr(), p_drop > 0 ? dmask.data_ptr<uint8_t>() : nullptr, cur_stream());
  check_launch();
  return {h, dmask};
}

torch::Tensor lm_gelu_drop_bwd(torch::Tensor dh, torch::Tensor u, torch::Tensor dmask, double p_drop,
                               c10::optional<torch::Tensor> dbias, double bias_alpha) {
  check_bf(dh, "dh"); check_bf(u, "u");
  TORCH_CHECK(u.dim() == 2 && dh.numel() == u.numel());
  const int M = u.size(0), F = u.size(1);
  TORCH_CHECK(p_drop <= 0 || dmask.numel() == u.numel());
  c10::cuda::CUDAGuard guard(u.device());
  auto du = torch::empty_like(u);
  hfl_lm_gelu_drop_bwd(dh.data_ptr(), u.data_ptr(), p_drop > 0 ? dmask.data_ptr<uint8_t>() : nullptr, (float)p_drop, M, F,
                       du.data_ptr(), dbias.has_value() ? f32p(*dbias, F) : nullptr, (float)bias_alpha, cur_stream());
  check_launch();
  return du;
}

void lm_colsum(torch::Tensor x, double alpha, torch::Tensor out) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kBFloat16 && x.dim() == 2 && x.stride(1) == 1);
  c10::cuda::CUDAGuard guard(x.device());
  hfl_lm_colsum(x.data_ptr(), (int)x.size(0), (int)x.size(1), (int)x.stride(0), (float)alpha, f32p(out, x.size(1)),
                cur_stream());
  check_launch();
}

std::vector<torch::Tensor> lm_attn_fwd(torch::Tensor q, torch::Tensor k, torch::Tensor v, int64_t N, int64_t S,
                                       int64_t H, double scale) {
  check_bf(q, "q"); check_bf(k, "k"); check_bf(v, "v");
  TORCH_CHECK(q.dim() == 2 && q.size(0) == N * S && q.size(1) % H == 0 && k.sizes() == q.sizes() && v.sizes() == q.sizes());
  const int dh = q.size(1) / H;
  TORCH_CHECK(S <= hfl_lm_max_s() && dh <= hfl_lm_max_dh(), "lm_attn: S <= 64 and head dim <= 32 only");
  c10::cuda::CUDAGuard guard(q.device());
  auto o = torch::empty_like(q);
  auto P = torch::empty({N, H, S, S}, q.options());
  hfl_lm_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), (int)N, (int)S, (int)H, dh, (float)scale, o.data_ptr(),
                  P.data_ptr(), cur_stream());
  check_launch();
  return {o, P};
}

std::vector<torch::Tensor> lm_attn_bwd(torch::Tensor dO, torch::Tensor q, torch::Tensor k, torch::Tensor v,
                                       torch::Tensor O, torch::Tensor P, int64_t N, int64_t S, int64_t H, double scale,
                                       torch::Tensor dbq, torch::Tensor dbk, torch::Tensor dbv, double alpha_v,
                                       bool scale_dv) {
  check_bf(dO, "dO"); check_bf(q, "q"); check_bf(k, "k"); check_bf(v, "v"); check_bf(O, "O"); check_bf(P, "P");
  TORCH_CHECK(O.sizes() == q.sizes());
  const int E = q.size(1), dh = E / H;
  TORCH_CHECK(q.size(0) == N * S && dO.sizes() == q.sizes() && P.numel() == N * H * S * S);
  TORCH_CHECK(S <= hfl_lm_max_s() && dh <= hfl_lm_max_dh());
  c10::cuda::CUDAGuard guard(q.device());
  auto dq = torch::empty_like(q), dk = torch::empty_like(q), dv = torch::empty_like(q);
  int e = hfl_lm_attn_bwd(dO.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), O.data_ptr(), P.data_ptr(), (int)N,
                          (int)S, (int)H, dh, (float)scale, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), f32p(dbq, E), f32p(dbk, E),
                          f32p(dbv, E), (float)alpha_v, scale_dv ? 1 : 0, cur_stream());
  TORCH_CHECK(e == 0, "lm_attn_bwd failed with code ", e);
  return {dq, dk, dv};
}

// ---- tcgen05 implicit-GEMM convolution (forward) ---------------------------------------------------------
bool conv_tc_supported(int64_t N, int64_t H, int64_t W, int64_t C, int64_t Kout, int64_t R, int64_t S, int64_t stride) {
  return hfl_conv_tc_supported((int)N, (int)H, (int)W, (int)C, (int)Kout, (int)R, (int)S, (int)stride) == 0;
}

// x: [N, C, H, W] channels-last bf16; w: [K, C, R, S] with channels-last strides (physically K,R,S,C);
// residual: like the output (optional); stats: fp32 [2, K] accumulated in place (optional).  padding = R/2.
// x2 / w2 (optional): block input [N, C2, s2*OH, s2*OW] and 1x1 shortcut filter [K, C2, 1, 1]; their stride-s2 convolution is
// accumulated into the same output tile (one launch for  conv2(o2) + shortcut(o1)  of a ResNet transition block)
torch::Tensor conv_fprop_tc(torch::Tensor x, torch::Tensor w, c10::optional<torch::Tensor> residual,
                            c10::optional<torch::Tensor> stats, int64_t stride, c10::optional<torch::Tensor> nvalid,
                            c10::optional<torch::Tensor> x2, c10::optional<torch::Tensor> w2, int64_t stride2) {
  CHECK_CUDA_T(x); CHECK_CUDA_T(w);
  TORCH_CHECK(x.dim() == 4 && w.dim() == 4 && x.scalar_type() == torch::kBFloat16 && w.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(x.is_contiguous(torch::MemoryFormat::ChannelsLast), "conv_fprop_tc: x must be channels-last");
  TORCH_CHECK(w.permute({0, 2, 3, 1}).is_contiguous(), "conv_fprop_tc: w must be stored K,R,S,C");
  const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
  const int K = w.size(0), R = w.size(2), S = w.size(3);
  TORCH_CHECK(w.size(1) == C);
  TORCH_CHECK(hfl_conv_tc_supported(N, H, W, C, K, R, S, (int)stride) == 0, "conv_fprop_tc: unsupported problem");
  const int pad = R / 2;
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  c10::cuda::CUDAGuard guard(x.device());
  auto y = torch::empty({N, K, OH, OW}, x.options().memory_format(torch::MemoryFormat::ChannelsLast));
  const void* rp = nullptr;
  if (residual.has_value()) {
    TORCH_CHECK(residual->sizes() == y.sizes() && residual->is_contiguous(torch::MemoryFormat::ChannelsLast) &&
                residual->scalar_type() == torch::kBFloat16);
    rp = residual->data_ptr();
  }
  float* sp = nullptr;
  if (stats.has_value()) { TORCH_CHECK(stats->scalar_type() == torch::kFloat32 && stats->numel() >= 2 * K && stats->is_contiguous()); sp = stats->data_ptr<float>(); }
  const int* nvp = nullptr;   // int32 [1] on the device: samples >= *nvalid are padding (zero outputs, no statistics)
  if (nvalid.has_value()) { TORCH_CHECK(nvalid->is_cuda() && nvalid->scalar_type() == torch::kInt32); nvp = nvalid->data_ptr<int>(); }
  const void* x2p = nullptr;
  const void* w2p = nullptr;
  int C2 = 0;
  if (x2.has_value()) {
    TORCH_CHECK(w2.has_value() && x2->dim() == 4 && x2->scalar_type() == torch::kBFloat16 && w2->scalar_type() == torch::kBFloat16);
    TORCH_CHECK(x2->is_contiguous(torch::MemoryFormat::ChannelsLast), "conv_fprop_tc: x2 must be channels-last");
    C2 = x2->size(1);
    TORCH_CHECK(x2->size(0) == N && x2->size(2) == OH * stride2 && x2->size(3) == OW * stride2, "conv_fprop_tc: x2 extent");
    TORCH_CHECK(w2->dim() == 4 && w2->size(0) == K && w2->size(1) == C2 && w2->size(2) == 1 && w2->size(3) == 1 &&
                w2->permute({0, 2, 3, 1}).is_contiguous(), "conv_fprop_tc: w2 must be a [K, C2, 1, 1] filter");
    x2p = x2->data_ptr();
    w2p = w2->data_ptr();
  }
  int e = hfl_conv_fprop_tc2(x.data_ptr(), w.data_ptr(), y.data_ptr(), rp, sp, nvp, N, H, W, C, K, R, S, (int)stride, x2p, w2p, C2,
                             (int)stride2, cur_stream());
  TORCH_CHECK(e == 0, "conv_fprop_tc failed with code ", e);
  return y;
}

// dy: [N, K, OH, OW] channels-last bf16 (gradient of a stride-1 or stride-2 convolution's output, padding R/2, even input
// extent); w: [K, C, R, S] stored K,R,S,C; returns dx [N, C, stride*OH, stride*OW]
bool conv_dgrad_tc_supported(int64_t N, int64_t H, int64_t W, int64_t C, int64_t Kout, int64_t R, int64_t S, int64_t stride) {
  return hfl_conv_dgrad_tc_supported((int)N, (int)H, (int)W, (int)C, (int)Kout, (int)R, (int)S, (int)stride) == 0;
}

torch::Tensor conv_dgrad_tc(torch::Tensor dy, torch::Tensor w, c10::optional<torch::Tensor> residual, int64_t stride) {
  CHECK_CUDA_T(dy); CHECK_CUDA_T(w);
  TORCH_CHECK(dy.dim() == 4 && w.dim() == 4 && dy.scalar_type() == torch::kBFloat16 && w.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(dy.is_contiguous(torch::MemoryFormat::ChannelsLast) && w.permute({0, 2, 3, 1}).is_contiguous());
  const int N = dy.size(0), K = dy.size(1), H = dy.size(2) * stride, W = dy.size(3) * stride;
  const int C = w.size(1), R = w.size(2), S = w.size(3);
  TORCH_CHECK(w.size(0) == K);
  TORCH_CHECK(hfl_conv_dgrad_tc_supported(N, H, W, C, K, R, S, (int)stride) == 0, "conv_dgrad_tc: unsupported problem");
  c10::cuda::CUDAGuard guard(dy.device());
  auto dx = torch::empty({N, C, H, W}, dy.options().memory_format(torch::MemoryFormat::ChannelsLast));
  const void* rp = nullptr;
  if (residual.has_value()) {
    TORCH_CHECK(residual->sizes() == dx.sizes() && residual->is_contiguous(torch::MemoryFormat::ChannelsLast));
    rp = residual->data_ptr();
  }
  int e = hfl_conv_dgrad_tc(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), rp, N, H, W, C, K, R, S, (int)stride, cur_stream());
  TORCH_CHECK(e == 0, "conv_dgrad_tc failed with code ", e);
  return dx;
}

bool conv_wgrad_tc_supported(int64_t N, int64_t H, int64_t W, int64_t C, int64_t Kout, int64_t R, int64_t S,
                             int64_t stride) {
  return hfl_conv_wgrad_tc_supported((int)N, (int)H, (int)W, (int)C, (int)Kout, (int)R, (int)S, (int)stride) == 0;
}

// dy: [N, K, OH, OW], x: [N, C, H, W] (channels-last bf16) -> dw fp32, logical [K, C, R, S] stored K,R,S,C
torch::Tensor conv_wgrad_tc(torch::Tensor dy, torch::Tensor x, int64_t R, int64_t stride) {
  CHECK_CUDA_T(dy); CHECK_CUDA_T(x);
  TORCH_CHECK(dy.scalar_type() == torch::kBFloat16 && x.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(dy.is_contiguous(torch::MemoryFormat::ChannelsLast) && x.is_contiguous(torch::MemoryFormat::ChannelsLast));
  const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3), K = dy.size(1);
  TORCH_CHECK(hfl_conv_wgrad_tc_supported(N, H, W, C, K, (int)R, (int)R, (int)stride) == 0, "conv_wgrad_tc: unsupported");
  c10::cuda::CUDAGuard guard(x.device());
  auto dw = torch::empty({K, R, R, C}, torch::TensorOptions().dtype(torch::kFloat32).device(x.device()));
  int e = hfl_conv_wgrad_tc(dy.data_ptr(), x.data_ptr(), dw.data_ptr<float>(), N, H, W, C, K, (int)R, (int)R, (int)stride, 1,
                            cur_stream());
  TORCH_CHECK(e >= 1, "conv_wgrad_tc failed with code ", e);
  return dw.permute({0, 3, 1, 2});
}


// ---- peer memory runtime (cudaMalloc + CUDA IPC) ---------------------------------------------------------
int64_t peer_alloc(int64_t bytes) {
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
  return (int64_t)p;
}
void peer_free(int64_t ptr) { CUDA_OK(cudaFree((void*)ptr)); }
py::bytes ipc_get_handle(int64_t ptr) {
  cudaIpcMemHandle_t h;
  CUDA_OK(cudaIpcGetMemHandle(&h, (void*)ptr));
  return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
}
int64_t ipc_open_handle(const std::string& handle) {
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return (int64_t)p;
}
void ipc_close_handle(int64_t ptr) { CUDA_OK(cudaIpcCloseMemHandle((void*)ptr)); }

torch::Tensor tensor_from_ptr(int64_t ptr, int64_t numel, const std::string& dtype, int64_t device) {
  torch::ScalarType st = torch::kFloat32;
  if (dtype == "float32") st = torch::kFloat32;
  else if (dtype == "bfloat16") st = torch::kBFloat16;
  else if (dtype == "uint8") st = torch::kUInt8;
  else if (dtype == "int32") st = torch::kInt32;
  else if (dtype == "int64") st = torch::kInt64;
  else if (dtype == "float64") st = torch::kFloat64;
  else TORCH_CHECK(false, "unsupported dtype ", dtype);
  return torch::from_blob((void*)ptr, {numel}, torch::TensorOptions().dtype(st).device(torch::kCUDA, (int)device));
}

void peer_barrier(torch::Tensor flag_ptrs, int64_t rank, int64_t world, int64_t epoch, torch::Tensor err) {
  CHECK_CUDA_T(flag_ptrs);
  c10::cuda::CUDAGuard guard(flag_ptrs.device());
  hfl_peer_barrier(flag_ptrs.data_ptr(), (int)rank, (int)world, (unsigned)epoch, err.data_ptr<int>(), cur_stream());
  check_launch();
}



// Grouped GEMMs sharing one launch (csrc/gemm_grouped_tcgen05.cu).  `a` / `b`: 1 or G tensors each (a single tensor is
// shared by every group); layouts as in gemm_bf16_general.  accum = false: returns G outputs C_g = alpha_g (A_g B_g^T +
// bias_g); accum = true: returns ONE output alpha_0 * sum_g A_g B_g^T.
std::vector<torch::Tensor> gemm_grouped(std::vector<torch::Tensor> a, bool a_mn, std::vector<torch::Tensor> b, bool b_mn,
                                        c10::optional<std::vector<torch::Tensor>> bias, std::vector<double> alpha,
                                        bool accum) {
  const int G = (int)std::max(a.size(), b.size());
  TORCH_CHECK(G >= 1 && G <= 3 && (a.size() == 1 || (int)a.size() == G) && (b.size() == 1 || (int)b.size() == G));
  TORCH_CHECK((int)alpha.size() == G || alpha.size() == 1);
  auto A = [&](int g) -> torch::Tensor& { return a[a.size() == 1 ? 0 : g]; };
  auto B = [&](int g) -> torch::Tensor& { return b[b.size() == 1 ? 0 : g]; };
  const int64_t M = a_mn ? A(0).size(1) : A(0).size(0), K = a_mn ? A(0).size(0) : A(0).size(1);
  const int64_t N = b_mn ? B(0).size(1) : B(0).size(0);
  const void* ap[3]; const void* bp[3]; const void* biasp[3] = {nullptr, nullptr, nullptr}; void* cp[3]; float al[3];
  for (int g = 0; g < G; ++g) {
    auto& x = A(g); auto& w = B(g);
    CHECK_CUDA_T(x); CHECK_CUDA_T(w);
    TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.scalar_type() == torch::kBFloat16 && w.scalar_type() == torch::kBFloat16);
    TORCH_CHECK(x.stride(1) == 1 && w.stride(1) == 1 && x.stride(0) % 8 == 0 && w.stride(0) % 8 == 0);
    TORCH_CHECK(x.sizes() == A(0).sizes() && w.sizes() == B(0).sizes() && x.stride(0) == A(0).stride(0) &&
                w.stride(0) == B(0).stride(0), "gemm_grouped: groups must share shapes and strides");
    TORCH_CHECK((b_mn ? w.size(0) : w.size(1)) == K, "gemm_grouped: reduction dims differ");
    TORCH_CHECK((reinterpret_cast<uintptr_t>(x.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(w.data_ptr()) & 15) == 0);
    ap[g] = x.data_ptr(); bp[g] = w.data_ptr();
    al[g] = (float)alpha[alpha.size() == 1 ? 0 : g];
    if (bias.has_value()) {
      auto& bt = (*bias)[g];
      TORCH_CHECK(bt.scalar_type() == torch::kBFloat16 && bt.numel() == N && bt.is_contiguous());
      biasp[g] = bt.data_ptr();
    }
  }
  c10::cuda::CUDAGuard guard(A(0).device());
  std::vector<torch::Tensor> out;
  const int n_out = accum ? 1 : G;
  for (int g = 0; g < n_out; ++g) {
    out.push_back(torch::empty({M, N}, torch::TensorOptions().dtype(torch::kBFloat16).device(A(0).device())));
    cp[g] = out.back().data_ptr();
  }
  for (int g = n_out; g < 3; ++g) cp[g] = cp[0];
  int e = hfl_gemm_grouped(G, ap, (int)A(0).stride(0), a_mn ? 1 : 0, bp, (int)B(0).stride(0), b_mn ? 1 : 0, cp, (int)N,
                           bias.has_value() ? biasp : nullptr, al, accum ? 1 : 0, (int)M, (int)N, (int)K, cur_stream());
  TORCH_CHECK(e == 0, "gemm_grouped failed with code ", e);
  return out;
}

// ---- resident client-training megakernel (csrc/mega_kernel.cu) ------------------------------------------
// jobs: int64 [K, 4] device tensor of {program ptr, nvalid-table ptr (0: none), per-job ctl ptr, nsteps}
void mega_launch(torch::Tensor jobs, int64_t nphase, int64_t nsteps, torch::Tensor ctl, int64_t grid,
                 c10::optional<torch::Tensor> prof) {
  CHECK_CUDA_T(jobs); CHECK_CUDA_T(ctl); CHECK_CONTIG(jobs);
  TORCH_CHECK(jobs.scalar_type() == torch::kInt64 && jobs.dim() == 2 && jobs.size(1) == 4);
  TORCH_CHECK(jobs.size(0) >= 1 && jobs.size(0) <= hfl_mega_max_jobs(), "mega_launch: 1..", hfl_mega_max_jobs(), " jobs");
  TORCH_CHECK(ctl.scalar_type() == torch::kInt32 && ctl.numel() >= 16);
  void* pr = nullptr;
  if (prof.has_value()) {   // per-phase (work, barrier) cycle counters of CTA 0: int64 [2 * nphase]
    CHECK_CUDA_T(*prof);
    TORCH_CHECK(prof->scalar_type() == torch::kInt64 && prof->numel() >= 2 * nphase && prof->is_contiguous());
    pr = prof->data_ptr();
  }
  c10::cuda::CUDAGuard guard(jobs.device());
  int e = hfl_mega_launch(jobs.data_ptr(), (int)jobs.size(0), (int)nphase, (int)nsteps, ctl.data_ptr(), (int)grid, pr,
                          cur_stream());
  TORCH_CHECK(e == 0, "mega_launch failed with code ", e, " (", cudaGetErrorString((cudaError_t)e), ")");
}

// 128-byte CUtensorMap (bf16, 128B swizzle, zero OOB fill) for a device pointer; dims/box/estr innermost first
torch::Tensor mega_encode_tmap(int64_t base_ptr, std::vector<int64_t> dims, std::vector<int64_t> strides_bytes,
                               std::vector<int64_t> box, std::vector<int64_t> estr) {
  const int rank = (int)dims.size();
  TORCH_CHECK(rank >= 2 && rank <= 5 && (int)strides_bytes.size() == rank - 1 && (int)box.size() == rank &&
              (int)estr.size() == rank);
  unsigned long long d[5], s[4];
  unsigned b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = (unsigned long long)dims[i]; b[i] = (unsigned)box[i]; e[i] = (unsigned)estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = (unsigned long long)strides_bytes[i];
  auto out = torch::empty({128}, torch::dtype(torch::kUInt8));
  int r = hfl_mega_encode_tmap(out.data_ptr(), reinterpret_cast<const void*>(base_ptr), rank, d, s, b, e);
  TORCH_CHECK(r == 0, "cuTensorMapEncodeTiled failed with code ", r);
  return out;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("aggregate", &aggregate);
  m.def("aggregate_v2", &aggregate_v2);
  m.def("fed2_chunk", &hfl_fed2_chunk);
  m.def("distribute", &distribute);
  m.def("peer_sum", &peer_sum);
  m.def("fed_chunk", &hfl_fed_chunk);
  m.def("fed_max_clients", &hfl_fed_max_clients);
  m.def("sgd_clip_step", &sgd_clip_step);
  m.def("masked_ce", &masked_ce);
  m.def("augment", &augment);
  m.def("bn_act_fwd", &bn_act_fwd);
  m.def("bn_apply_sums", &bn_apply_sums);
  m.def("affine_act", &affine_act);
  m.def("bn_act_bwd", &bn_act_bwd);
  m.def("mx_quantize", &mx_quantize);
  m.def("gemm_mxfp8", &gemm_mxfp8);
  m.def("maxpool2_fwd", &maxpool2_fwd);
  m.def("maxpool2_bwd", &maxpool2_bwd);
  m.def("gn_act_fwd", &gn_act_fwd);
  m.def("gn_act_bwd", &gn_act_bwd);
  m.def("multi_sgd_clip_step", &multi_sgd_clip_step);
  m.def("ms_chunk", &hfl_ms_chunk);
  m.def("emb_sparse_norm", &emb_sparse_norm);
  m.def("emb_sparse_apply", &emb_sparse_apply);
  m.def("head_fwd_bwd", &head_fwd_bwd);
  m.def("head_smem_bytes", &hfl_head_smem_bytes);
  m.def("gemm_bf16_tn", &gemm_bf16_tn, py::arg("x"), py::arg("w"), py::arg("bias"), py::arg("alpha"), py::arg("gelu"),
        py::arg("out_fp32"), py::arg("bias_inside") = false, py::arg("out") = py::none());
  m.def("gemm_bf16_general", &gemm_bf16_general);
  m.def("gemm_grouped", &gemm_grouped);
  m.def("gemm_timing", []() { std::vector<long long> v(16); hfl_gemm_timing(v.data()); return v; });
  m.def("lm_embed_fwd", &lm_embed_fwd);
  m.def("lm_embed_bwd", &lm_embed_bwd);
  m.def("lm_add_ln_fwd", &lm_add_ln_fwd);
  m.def("lm_add_ln_bwd", &lm_add_ln_bwd, py::arg("dys"), py::arg("xhat"), py::arg("rstd"), py::arg("gamma"), py::arg("dmask"),
        py::arg("p_drop"), py::arg("dgamma"), py::arg("dbeta"), py::arg("dbias"), py::arg("bias_alpha"),
        py::arg("dy_f32") = py::none());
  m.def("lm_vocab_supported", &lm_vocab_supported);
  m.def("lm_vocab_head", &lm_vocab_head);
  m.def("lm_gelu_drop_fwd", &lm_gelu_drop_fwd);
  m.def("lm_gelu_drop_bwd", &lm_gelu_drop_bwd);
  m.def("lm_colsum", &lm_colsum);
  m.def("lm_attn_fwd", &lm_attn_fwd);
  m.def("lm_attn_bwd", &lm_attn_bwd, py::arg("dO"), py::arg("q"), py::arg("k"), py::arg("v"), py::arg("O"), py::arg("P"),
        py::arg("N"), py::arg("S"), py::arg("H"), py::arg("scale"), py::arg("dbq"), py::arg("dbk"), py::arg("dbv"),
        py::arg("alpha_v"), py::arg("scale_dv") = false);
  m.def("lm_max_e", &hfl_lm_max_e);
  m.def("conv_tc_supported", &conv_tc_supported);
  m.def("conv_fprop_tc", &conv_fprop_tc, py::arg("x"), py::arg("w"), py::arg("residual"), py::arg("stats"), py::arg("stride"),
        py::arg("nvalid") = py::none(), py::arg("x2") = py::none(), py::arg("w2") = py::none(), py::arg("stride2") = 1);
  m.def("conv_dgrad_tc", &conv_dgrad_tc, py::arg("dy"), py::arg("w"), py::arg("residual"), py::arg("stride") = 1);
  m.def("conv_dgrad_tc_supported", &conv_dgrad_tc_supported);
  m.def("conv_set_ksplit", [](int64_t cap) { hfl_conv_set_ksplit((int)cap); });
  m.def("conv_wgrad_tc", &conv_wgrad_tc);
  m.def("conv_wgrad_tc_supported", &conv_wgrad_tc_supported);
  m.def("mega_launch", &mega_launch, py::arg("jobs"), py::arg("nphase"), py::arg("nsteps"), py::arg("ctl"), py::arg("grid"),
        py::arg("prof") = py::none());
  m.def("mega_max_jobs", &hfl_mega_max_jobs);
  m.def("mega_encode_tmap", &mega_encode_tmap);
  m.def("mega_phase_bytes", &hfl_mega_phase_bytes);
  m.def("mega_smem_bytes", &hfl_mega_smem_bytes);
  m.def("mega_max_grid", &hfl_mega_max_grid);
  m.def("mega_occupancy_probe", []() { std::vector<int> v(16); hfl_mega_occupancy_probe(v.data()); return v; });
  m.def("launch_count", &hfl_launch_count);
  m.def("launch_count_add", &hfl_launch_count_add);
  m.def("peer_alloc", &peer_alloc);
  m.def("peer_free", &peer_free);
  m.def("ipc_get_handle", &ipc_get_handle);
  m.def("ipc_open_handle", &ipc_open_handle);
  m.def("ipc_close_handle", &ipc_close_handle);
  m.def("tensor_from_ptr", &tensor_from_ptr);
  m.def("peer_barrier", &peer_barrier);
}
