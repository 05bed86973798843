// Channels-last 2-D max pooling for the MC-batched activations ([S*B, H, W, C], bf16 or fp32).
// Not part of the Bayesian layers themselves: torchvision's ResNet stem runs nn.MaxPool2d right after the
// first Bayesian conv, and on the [S*B]-stacked activations ATen's max_pool_forward_nhwc reaches only ~0.6 TB/s
// (profiles/r01b); this streaming kernel (16-byte loads, one output vector per thread) is HBM-bound instead.
// Used by bayesian_torch_b200.fuse.fuse_inference (whole-forward MC batching).
#include "bt_common.cuh"

namespace {

struct PoolArgs {
  const void* x;
  void* out;
  long long n_img;
  int H, W, C, OH, OW, kh, kw, sh, sw, ph, pw;
};

// IDX = uint32_t whenever the output has < 2^31 vectors (always, in practice): the four index divisions per output
// vector are then 32-bit (the 64-bit ones made the kernel instruction-bound at ~3 TB/s, profiles/r01h).
template <bool BF16, typename IDX>
__global__ void maxpool_nhwc_kernel(const PoolArgs a) {
  constexpr int VE = BF16 ? 8 : 4;  // elements per 16-byte vector
  const IDX cv = (IDX)(a.C / VE);
  const IDX total = (IDX)(a.n_img * a.OH * a.OW * cv);
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    IDX r = i / cv;
    const int ow = (int)(r % (IDX)a.OW);
    r /= (IDX)a.OW;
    const int oh = (int)(r % (IDX)a.OH);
    const long long n = (long long)(r / (IDX)a.OH);
    const int h0 = oh * a.sh - a.ph, w0 = ow * a.sw - a.pw;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int dh = 0; dh < a.kh; ++dh) {
      const int h = h0 + dh;
      if ((unsigned)h >= (unsigned)a.H) continue;
      for (int dw = 0; dw < a.kw; ++dw) {
        const int w = w0 + dw;
        if ((unsigned)w >= (unsigned)a.W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.x) + ((n * a.H + h) * a.W + w) * cv + c);
        if (BF16) {
          const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            m[2 * j] = fmaxf(m[2 * j], bt_bf16_lo(ws[j]));
            m[2 * j + 1] = fmaxf(m[2 * j + 1], bt_bf16_hi(ws[j]));
          }
        } else {
          m[0] = fmaxf(m[0], __uint_as_float(v.x));
          m[1] = fmaxf(m[1], __uint_as_float(v.y));
          m[2] = fmaxf(m[2], __uint_as_float(v.z));
          m[3] = fmaxf(m[3], __uint_as_float(v.w));
        }
      }
    }
    uint4 o;
    if (BF16) {
      o.x = bt_pack_bf16x2(m[0], m[1]); o.y = bt_pack_bf16x2(m[2], m[3]);
      o.z = bt_pack_bf16x2(m[4], m[5]); o.w = bt_pack_bf16x2(m[6], m[7]);
    } else {
      o.x = __float_as_uint(m[0]); o.y = __float_as_uint(m[1]);
      o.z = __float_as_uint(m[2]); o.w = __float_as_uint(m[3]);
    }
    reinterpret_cast<uint4*>(a.out)[i] = o;
  }
}

}  // namespace

extern "C" int bt_maxpool2d_nhwc(const void* x, int dtype, int64_t n_img, int32_t H, int32_t W, int32_t C,
                                 int32_t kh, int32_t kw, int32_t sh, int32_t sw, int32_t ph, int32_t pw,
                                 void* out, void* stream) {
  BT_REQUIRE(dtype == BT_F32 || dtype == BT_BF16, BT_ERR_BAD_DTYPE, "bt_maxpool2d_nhwc: dtype %d", dtype);
  const int ve = dtype == BT_BF16 ? 8 : 4;
  BT_REQUIRE(n_img > 0 && H > 0 && W > 0 && C > 0 && C % ve == 0, BT_ERR_BAD_SHAPE,
             "bt_maxpool2d_nhwc: C (%d) must be a positive multiple of %d", C, ve);
  BT_REQUIRE(kh >= 1 && kw >= 1 && sh >= 1 && sw >= 1 && ph >= 0 && pw >= 0 && 2 * ph <= kh && 2 * pw <= kw,
             BT_ERR_BAD_SHAPE, "bt_maxpool2d_nhwc: bad window");
  int rc;
  if ((rc = bt_check_device_ptr(x, "x")) != BT_OK) return rc;
  if ((rc = bt_check_device_ptr(out, "out")) != BT_OK) return rc;
  BT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             BT_ERR_BAD_POINTER, "bt_maxpool2d_nhwc: pointers must be 16-byte aligned");
  PoolArgs a;
  a.x = x; a.out = out; a.n_img = n_img; a.H = H; a.W = W; a.C = C;
  a.kh = kh; a.kw = kw; a.sh = sh; a.sw = sw; a.ph = ph; a.pw = pw;
  a.OH = (H + 2 * ph - kh) / sh + 1;
  a.OW = (W + 2 * pw - kw) / sw + 1;
  BT_REQUIRE(a.OH >= 1 && a.OW >= 1, BT_ERR_BAD_SHAPE, "bt_maxpool2d_nhwc: empty output");
  const long long total = n_img * a.OH * a.OW * (C / ve);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool small = total < (1ll << 31) - (long long)blocks * 256;   // (the grid-stride increment must not wrap either)
  if (dtype == BT_BF16) {
    if (small) maxpool_nhwc_kernel<true, uint32_t><<<(unsigned)blocks, 256, 0, st>>>(a);
    else maxpool_nhwc_kernel<true, unsigned long long><<<(unsigned)blocks, 256, 0, st>>>(a);
  } else {
    if (small) maxpool_nhwc_kernel<false, uint32_t><<<(unsigned)blocks, 256, 0, st>>>(a);
    else maxpool_nhwc_kernel<false, unsigned long long><<<(unsigned)blocks, 256, 0, st>>>(a);
  }
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}
