// Regenerates into global memory the random draws of a fused forward launch (parity / debug and
// the on-demand `eps_weight`, `eps_kernel`, `eps_bias` buffers of the reference API,
// /root/reference/bayesian_torch/layers/variational_layers/linear_variational.py:90-92,161,173).
// Uses the SAME device functions and counter layout as bt_fused.cu (bt_philox.cuh).
#include "bt_common.cuh"
#include "bt_philox.cuh"

namespace {

// weight eps: physical K order is (tap, channel); the reference's logical order is (channel, tap).
__global__ void rng_weight_eps(float* out, long long rows, long long K, int taps, BtRngKey key,
                               uint32_t sample) {
  const long long kq_per_row = (K + 3) / 4;
  const long long total = rows * kq_per_row;
  const int cpt = (int)(K / taps);  // channels per tap
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / kq_per_row, kq = i % kq_per_row;
    const float4 z = bt_eps_quad(key, BT_STREAM_W_EPS, (uint32_t)kq, (uint32_t)n, sample);
    const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long k = kq * 4 + j;
      if (k < K) {
        const long long tap = k / cpt, c = k % cpt;
        out[n * K + c * taps + tap] = zz[j];
      }
    }
  }
}

__global__ void rng_bias_eps(float* out, long long n, BtRngKey key, uint32_t sample) {
  const long long nq = (n + 3) / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 z = bt_eps_quad(key, BT_STREAM_B_EPS, (uint32_t)i, 0u, sample);
    const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = zz[j];
  }
}

// signs: out [rows, cols] +-1; block index inside a row = (group << 20) | (local col / 128)
__global__ void rng_signs(float* out, long long rows, long long cols, int cols_per_group,
                          uint32_t stream, BtRngKey key, uint32_t sample) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    uint32_t blk, bit;
    if (stream == BT_STREAM_SIGN_OUT) {
      const int g = c / cols_per_group, cl = c % cols_per_group;
      blk = ((uint32_t)g << 20) | (uint32_t)(cl >> 7);
      bit = cl & 127;
    } else {
      blk = (uint32_t)(c >> 7);
      bit = c & 127;
    }
    const uint4 b = bt_sign_block(key, stream, blk, (uint32_t)r, sample);
    const uint32_t w = bt_sign_word(b, bit >> 5);
    out[i] = ((w >> (bit & 31)) & 1u) ? -1.0f : 1.0f;
  }
}

}  // namespace

extern "C" int bt_rng_export(int what, float* out, int64_t rows, int64_t cols, int32_t taps,
                             int32_t cols_per_group, uint64_t seed, uint32_t layer_key,
                             uint32_t sample_idx, void* stream) {
  BT_REQUIRE(what >= 0 && what <= 3, BT_ERR_UNSUPPORTED, "bt_rng_export: what=%d", what);
  BT_REQUIRE(rows > 0 && (what == 1 || cols > 0), BT_ERR_BAD_SHAPE, "bt_rng_export: empty shape");
  BT_REQUIRE(layer_key < (1u << 28), BT_ERR_BAD_SHAPE, "bt_rng_export: layer_key must be < 2^28");
  int rc;
  if ((rc = bt_check_device_ptr(out, "out")) != BT_OK) return rc;
  BtRngKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), layer_key << 4};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  auto nblocks = [&](long long work) {
    long long b = (work + threads - 1) / threads;
    return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
  };
  if (what == 0) {
    BT_REQUIRE(taps >= 1 && cols % taps == 0, BT_ERR_BAD_SHAPE,
               "bt_rng_export: cols (%lld) must be a multiple of taps (%d)", (long long)cols, taps);
    rng_weight_eps<<<nblocks(rows * ((cols + 3) / 4)), threads, 0, st>>>(out, rows, cols, taps, key,
                                                                        sample_idx);
  } else if (what == 1) {
    rng_bias_eps<<<nblocks((rows + 3) / 4), threads, 0, st>>>(out, rows, key, sample_idx);
  } else {
    BT_REQUIRE(cols_per_group >= 1 && cols % cols_per_group == 0, BT_ERR_BAD_SHAPE,
               "bt_rng_export: cols_per_group");
    rng_signs<<<nblocks(rows * cols), threads, 0, st>>>(
        out, rows, cols, cols_per_group, what == 2 ? BT_STREAM_SIGN_IN : BT_STREAM_SIGN_OUT, key,
        sample_idx);
  }
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}
