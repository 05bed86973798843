// Monte-Carlo aggregation on device: softmax of every sample's logits, running sum of p and p^2.
//
// Replaces   output_ = torch.stack(output_mc); softmax; mean(0)
// (/root/reference/bayesian_torch/examples/main_bayesian_cifar_dnn2bnn.py:545-557) and the
// per-sample `.cpu().numpy()` round trip of examples/main_bayesian_imagenet.py:617-624.
// The [2,B,C] fp32 result is the ONLY thing that crosses GPUs (one NCCL all-reduce).
#include "bt_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float ld_logit(const T* p);
template <>
__device__ __forceinline__ float ld_logit<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_logit<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

// one warp per batch row b; loops over the S samples in order (deterministic).
// CPL = classes per lane held in registers (C <= 32 * CPL)
// ent_sum (nullable, [B]): running sum over the samples of the per-sample entropy  -sum_c p log(p + 1e-15)
// (utils/util.py:41-42 applied to every MC member, the second term of mutual_information, util.py:54-60).
template <typename T, int CPL>
__global__ void mc_accumulate_kernel(const T* __restrict__ logits, int S, int B, int C,
                                     float* __restrict__ sums, float* __restrict__ ent_sum, int accumulate) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  float a1[CPL], a2[CPL];
  float ent = 0.f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) a1[j] = a2[j] = 0.f;
  // The samples are ACCUMULATED strictly in order (deterministic, bit-identical for any batching), but their logits are
  // LOADED SB samples at a time: one dependent global load per sample made this kernel a 64-long latency chain (46 us per
  // step for 8192 x 10 logits, profiles/r02_launches_bf16.md).
  constexpr int SB = CPL == 1 ? 8 : (CPL == 4 ? 4 : 1);
  for (int s0 = 0; s0 < S; s0 += SB) {
    float vv[SB][CPL];
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      const int su = s0 + u < S ? s0 + u : S - 1;
      const T* row = logits + ((long long)su * B + warp) * C;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        vv[u][j] = c < C ? ld_logit<T>(row + c) : -INFINITY;
      }
    }
#pragma unroll
    for (int u = 0; u < SB; ++u) {
    if (s0 + u >= S) break;
    float v[CPL];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      v[j] = vv[u][j];
      mx = fmaxf(mx, v[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      v[j] = (lane + 32 * j < C) ? __expf(v[j] - mx) : 0.f;
      den += v[j];
    }
    den = bt_warp_sum(den);
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const float p = v[j] * inv;
      a1[j] += p;
      a2[j] = fmaf(p, p, a2[j]);
      if (ent_sum != nullptr && lane + 32 * j < C) ent = fmaf(-p, __logf(p + 1e-15f), ent);
    }
    }
  }
  if (ent_sum != nullptr) {
    ent = bt_warp_sum(ent);
    if (lane == 0) ent_sum[warp] = accumulate ? ent_sum[warp] + ent : ent;
  }
  float* o1 = sums + (long long)warp * C;
  float* o2 = sums + ((long long)B + warp) * C;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    if (c < C) {
      o1[c] = accumulate ? o1[c] + a1[j] : a1[j];
      o2[c] = accumulate ? o2[c] + a2[j] : a2[j];
    }
  }
}

__global__ void mc_finalize_kernel(const float* __restrict__ sums, long long n, float inv_total,
                                   float* __restrict__ mean, float* __restrict__ var) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float m = sums[i] * inv_total;
    mean[i] = m;
    if (var) var[i] = fmaxf(sums[n + i] * inv_total - m * m, 0.f);
  }
}

// predictive entropy  H(mean_s p_s)  and mutual information  H(mean p) - mean_s H(p_s)  (utils/util.py:45-60)
// from the (all-reduced) moment buffer: one warp per batch row.
__global__ void mc_uncertainty_kernel(const float* __restrict__ sums, const float* __restrict__ ent_sum, int B, int C,
                                      float inv_total, float* __restrict__ pred_entropy, float* __restrict__ mutual_info) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  float h = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float m = sums[(long long)warp * C + c] * inv_total;
    h = fmaf(-m, __logf(m + 1e-15f), h);
  }
  h = bt_warp_sum(h);
  if (lane == 0) {
    pred_entropy[warp] = h;
    if (mutual_info != nullptr) mutual_info[warp] = h - ent_sum[warp] * inv_total;
  }
}

template <typename T>
int launch_acc(const void* logits, int S, int B, int C, float* sums, float* ent, int acc, cudaStream_t st) {
  const int threads = 128;  // 4 rows per block
  const unsigned blocks = (unsigned)((B + 3) / 4);
  const T* l = static_cast<const T*>(logits);
  if (C <= 32) mc_accumulate_kernel<T, 1><<<blocks, threads, 0, st>>>(l, S, B, C, sums, ent, acc);
  else if (C <= 128) mc_accumulate_kernel<T, 4><<<blocks, threads, 0, st>>>(l, S, B, C, sums, ent, acc);
  else if (C <= 1024) mc_accumulate_kernel<T, 32><<<blocks, threads, 0, st>>>(l, S, B, C, sums, ent, acc);
  else {
    bt_set_error("bt_mc_accumulate: n_classes %d > 1024 not supported", C);
    return BT_ERR_UNSUPPORTED;
  }
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

}  // namespace

extern "C" {

int bt_mc_accumulate_ex(const void* logits, int dtype, int32_t n_samples, int32_t batch,
                        int32_t n_classes, float* sums, float* entropy_sum, int accumulate, void* stream) {
  BT_REQUIRE(n_samples > 0 && batch > 0 && n_classes > 0, BT_ERR_BAD_SHAPE,
             "bt_mc_accumulate: bad shape S=%d B=%d C=%d", n_samples, batch, n_classes);
  BT_REQUIRE(dtype == BT_F32 || dtype == BT_BF16, BT_ERR_BAD_DTYPE, "bt_mc_accumulate: dtype %d", dtype);
  int rc;
  if ((rc = bt_check_device_ptr(logits, "logits")) != BT_OK) return rc;
  if ((rc = bt_check_device_ptr(sums, "sums")) != BT_OK) return rc;
  if (entropy_sum != nullptr && (rc = bt_check_device_ptr(entropy_sum, "entropy_sum")) != BT_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == BT_F32 ? launch_acc<float>(logits, n_samples, batch, n_classes, sums, entropy_sum, accumulate, st)
                         : launch_acc<__nv_bfloat16>(logits, n_samples, batch, n_classes, sums, entropy_sum,
                                                     accumulate, st);
}

int bt_mc_accumulate(const void* logits, int dtype, int32_t n_samples, int32_t batch,
                     int32_t n_classes, float* sums, int accumulate, void* stream) {
  return bt_mc_accumulate_ex(logits, dtype, n_samples, batch, n_classes, sums, nullptr, accumulate, stream);
}

int bt_mc_uncertainty(const float* sums, const float* entropy_sum, int32_t batch, int32_t n_classes, int32_t n_total,
                      float* pred_entropy, float* mutual_info, void* stream) {
  BT_REQUIRE(batch > 0 && n_classes > 0 && n_total > 0, BT_ERR_BAD_SHAPE, "bt_mc_uncertainty: bad shape");
  BT_REQUIRE(mutual_info == nullptr || entropy_sum != nullptr, BT_ERR_BAD_POINTER,
             "bt_mc_uncertainty: mutual information needs the per-sample entropy sums");
  int rc;
  if ((rc = bt_check_device_ptr(sums, "sums")) != BT_OK) return rc;
  if ((rc = bt_check_device_ptr(pred_entropy, "pred_entropy")) != BT_OK) return rc;
  const unsigned blocks = (unsigned)((batch + 3) / 4);
  mc_uncertainty_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, entropy_sum, batch, n_classes, 1.0f / (float)n_total, pred_entropy, mutual_info);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

int bt_mc_finalize(const float* sums, int32_t batch, int32_t n_classes, int32_t n_total,
                   float* mean, float* var, void* stream) {
  BT_REQUIRE(batch > 0 && n_classes > 0 && n_total > 0, BT_ERR_BAD_SHAPE, "bt_mc_finalize: bad shape");
  int rc;
  if ((rc = bt_check_device_ptr(sums, "sums")) != BT_OK) return rc;
  if ((rc = bt_check_device_ptr(mean, "mean")) != BT_OK) return rc;
  const long long n = (long long)batch * n_classes;
  const unsigned blocks = (unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
  mc_finalize_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, n, 1.0f / (float)n_total, mean, var);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

}  // extern "C"
