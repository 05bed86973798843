// LSTM cell pointwise stage: gates = ff_i + ff_h; i, f, o = sigmoid, g = tanh; c' = f*c + i*g; h' = o*tanh(c').
//
// Replaces the 12 ATen launches per time step of LSTMReparameterization / LSTMFlipout.forward
// (/root/reference/bayesian_torch/layers/variational_layers/rnn_variational.py:127-141,
//  flipout_layers/rnn_flipout.py:127-141); the two gate GEMMs of the step are bt_layer_forward launches on the
// layer's `ih` / `hh` Bayesian linears.  h' and c' are also written into the [B, T, H] output sequences (the
// reference's torch.cat + transpose(0, 1).contiguous()).
#include "bt_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p, long long i);
template <>
__device__ __forceinline__ float ldf<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
template <typename T>
__device__ __forceinline__ void stf(T* p, long long i, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, long long i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

template <typename T>
__global__ void lstm_cell_kernel(const T* __restrict__ gi, const T* __restrict__ gh, const T* __restrict__ c_prev,
                                 T* __restrict__ h_out, T* __restrict__ c_out, T* __restrict__ h_seq, T* __restrict__ c_seq,
                                 int B, int H, int T_len, int t) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx - (long long)b * H);
  const long long g0 = (long long)b * 4 * H + j;
  const float i_t = sigmoidf_(ldf(gi, g0) + ldf(gh, g0));
  const float f_t = sigmoidf_(ldf(gi, g0 + H) + ldf(gh, g0 + H));
  const float g_t = tanhf(ldf(gi, g0 + 2 * H) + ldf(gh, g0 + 2 * H));
  const float o_t = sigmoidf_(ldf(gi, g0 + 3 * H) + ldf(gh, g0 + 3 * H));
  const float c = f_t * ldf(c_prev, idx) + i_t * g_t;
  const float h = o_t * tanhf(c);
  stf(h_out, idx, h);
  stf(c_out, idx, c);
  const long long sq = ((long long)b * T_len + t) * H + j;
  stf(h_seq, sq, h);
  stf(c_seq, sq, c);
}

}  // namespace

extern "C" int bt_lstm_cell(const void* gates_i, const void* gates_h, const void* c_prev, void* h_out, void* c_out,
                            void* h_seq, void* c_seq, int dtype, int32_t batch, int32_t hidden, int32_t seq_len, int32_t t,
                            void* stream) {
  BT_REQUIRE(dtype == BT_F32 || dtype == BT_BF16, BT_ERR_BAD_DTYPE, "bt_lstm_cell: dtype %d", dtype);
  BT_REQUIRE(batch >= 1 && hidden >= 1 && seq_len >= 1 && t >= 0 && t < seq_len, BT_ERR_BAD_SHAPE, "bt_lstm_cell: bad sizes");
  int rc;
  if ((rc = bt_device_check()) != BT_OK) return rc;
  const void* ptrs[7] = {gates_i, gates_h, c_prev, h_out, c_out, h_seq, c_seq};
  const char* names[7] = {"gates_i", "gates_h", "c_prev", "h_out", "c_out", "h_seq", "c_seq"};
  for (int i = 0; i < 7; ++i)
    if ((rc = bt_check_device_ptr(ptrs[i], names[i])) != BT_OK) return rc;
  const long long n = (long long)batch * hidden;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == BT_F32)
    lstm_cell_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float*>(gates_i), static_cast<const float*>(gates_h),
                                                       static_cast<const float*>(c_prev), static_cast<float*>(h_out),
                                                       static_cast<float*>(c_out), static_cast<float*>(h_seq),
                                                       static_cast<float*>(c_seq), batch, hidden, seq_len, t);
  else
    lstm_cell_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(gates_i), static_cast<const __nv_bfloat16*>(gates_h),
        static_cast<const __nv_bfloat16*>(c_prev), static_cast<__nv_bfloat16*>(h_out), static_cast<__nv_bfloat16*>(c_out),
        static_cast<__nv_bfloat16*>(h_seq), static_cast<__nv_bfloat16*>(c_seq), batch, hidden, seq_len, t);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}
