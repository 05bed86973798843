// Materialised im2col for few-channel 2-D convolutions (the RGB stem of a ResNet): x [N, C, H, W] (any strides) ->
// rows [N * OH * OW, kpad], column k = (kh, kw, c) for k < KH*KW*C, zero beyond.
//
// The stem of dnn_to_bnn(torchvision ResNet) (conv_variational.py:357-402 with C_in = 3) has 3 input channels = 6 bytes
// per pixel: too narrow for a TMA box or a 16-byte gather.  Its im2col matrix is tiny (CIFAR: 12.6 MB for B = 128),
// shared by all MC samples of a step, and turns the layer into a linear layer whose rows TMA stages with full 128-byte
// lines (bt_tma_kernel, tiled map).  Replaces the F.pad + unfold + copy_ ATen sequence of round 1 with ONE launch.
#include "bt_common.cuh"

namespace {

struct Im2colArgs {
  const void* x;
  void* out;
  long long sn, sc, sh, sw;   // element strides of x
  long long rows;
  int C, H, W, KH, KW, SH, SW, PH, PW, DH, DW, OH, OW, ktrue, kpad;
};

// Table form (ktrue <= IM2COL_TAB): the (tap, channel) decode of column k -- four integer divisions per ELEMENT in the
// generic kernel below, which made the 12.6 MB CIFAR stem matrix cost 27 us per step -- is done once per block into
// shared memory: element offset of (c, kh, kw) inside an image and the (dh, dw) shift of the tap.
constexpr int IM2COL_TAB = 1024;
template <typename T, int VEC>
__global__ void __launch_bounds__(256) im2col2d_tab_kernel(const Im2colArgs a) {
  __shared__ int t_off[IM2COL_TAB];
  __shared__ short2 t_d[IM2COL_TAB];
  for (int k = threadIdx.x; k < a.kpad; k += blockDim.x) {
    int off = 0;
    short2 d = make_short2(-30000, -30000);            // columns >= ktrue: always out of range -> zero
    if (k < a.ktrue) {
      const int c = k % a.C, tap = k / a.C;
      const int kw = tap % a.KW, kh = tap / a.KW;
      d = make_short2((short)(kh * a.DH), (short)(kw * a.DW));
      off = (int)(c * a.sc + (long long)kh * a.DH * a.sh + (long long)kw * a.DW * a.sw);
    }
    t_off[k] = off;
    t_d[k] = d;
  }
  __syncthreads();
  const int chunks = a.kpad / VEC;
  const long long total = a.rows * chunks;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / chunks;
    const int k0 = (int)(idx - m * chunks) * VEC;
    const int ow = (int)(m % a.OW);
    const long long t = m / a.OW;
    const int oh = (int)(t % a.OH);
    const long long n = t / a.OH;
    const int ih0 = oh * a.SH - a.PH, iw0 = ow * a.SW - a.PW;
    const T* x = static_cast<const T*>(a.x) + n * a.sn + (long long)ih0 * a.sh + (long long)iw0 * a.sw;
    T v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const short2 d = t_d[k0 + j];
      const bool in = (unsigned)(ih0 + d.x) < (unsigned)a.H && (unsigned)(iw0 + d.y) < (unsigned)a.W;
      v[j] = in ? x[t_off[k0 + j]] : T(0);
    }
    *reinterpret_cast<uint4*>(static_cast<T*>(a.out) + m * a.kpad + k0) = *reinterpret_cast<const uint4*>(v);
  }
}

template <typename T, int VEC>
__global__ void im2col2d_kernel(const Im2colArgs a) {
  const int chunks = a.kpad / VEC;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.rows * chunks) return;
  const long long m = idx / chunks;
  const int k0 = (int)(idx - m * chunks) * VEC;
  const int ow = (int)(m % a.OW);
  const long long t = m / a.OW;
  const int oh = (int)(t % a.OH);
  const long long n = t / a.OH;
  const T* x = static_cast<const T*>(a.x) + n * a.sn;
  T v[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int k = k0 + j;
    T val = T(0);
    if (k < a.ktrue) {
      const int c = k % a.C, tap = k / a.C;
      const int kw = tap % a.KW, kh = tap / a.KW;
      const int ih = oh * a.SH - a.PH + kh * a.DH, iw = ow * a.SW - a.PW + kw * a.DW;
      if ((unsigned)ih < (unsigned)a.H && (unsigned)iw < (unsigned)a.W) val = x[c * a.sc + ih * a.sh + iw * a.sw];
    }
    v[j] = val;
  }
  T* dst = static_cast<T*>(a.out) + m * a.kpad + k0;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
}

}  // namespace

extern "C" int bt_im2col2d(const void* x, int dtype, int64_t n_img, int32_t C, int32_t H, int32_t W, const int64_t* strides_nchw,
                           int32_t kh, int32_t kw, int32_t sh, int32_t sw, int32_t ph, int32_t pw, int32_t dh, int32_t dw,
                           int32_t kpad, void* out, void* stream) {
  BT_REQUIRE(dtype == BT_F32 || dtype == BT_BF16, BT_ERR_BAD_DTYPE, "bt_im2col2d: dtype %d", dtype);
  BT_REQUIRE(strides_nchw != nullptr, BT_ERR_BAD_POINTER, "bt_im2col2d: strides is NULL");
  int rc;
  if ((rc = bt_device_check()) != BT_OK) return rc;
  if ((rc = bt_check_device_ptr(x, "x")) != BT_OK) return rc;
  if ((rc = bt_check_device_ptr(out, "out")) != BT_OK) return rc;
  const int vec = dtype == BT_BF16 ? 8 : 4;
  Im2colArgs a;
  a.x = x; a.out = out;
  a.sn = strides_nchw[0]; a.sc = strides_nchw[1]; a.sh = strides_nchw[2]; a.sw = strides_nchw[3];
  a.C = C; a.H = H; a.W = W; a.KH = kh; a.KW = kw; a.SH = sh; a.SW = sw; a.PH = ph; a.PW = pw; a.DH = dh; a.DW = dw;
  a.OH = (H + 2 * ph - dh * (kh - 1) - 1) / sh + 1;
  a.OW = (W + 2 * pw - dw * (kw - 1) - 1) / sw + 1;
  BT_REQUIRE(n_img >= 1 && a.OH >= 1 && a.OW >= 1 && kpad % vec == 0 && kpad >= kh * kw * C, BT_ERR_BAD_SHAPE,
             "bt_im2col2d: bad geometry (kpad %d must be a multiple of %d and >= %d)", kpad, vec, kh * kw * C);
  BT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, BT_ERR_BAD_POINTER, "bt_im2col2d: out must be 16-byte aligned");
  a.ktrue = kh * kw * C; a.kpad = kpad;
  a.rows = n_img * a.OH * a.OW;
  const long long work = a.rows * (kpad / vec);
  const int threads = 256;
  const unsigned blocks = (unsigned)((work + threads - 1) / threads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tab = kpad <= IM2COL_TAB && (long long)(C - 1) * a.sc + (long long)(kh - 1) * dh * a.sh + (long long)(kw - 1) * dw * a.sw < (1ll << 31) &&
                   kh * dh < 30000 && kw * dw < 30000;
  if (tab) {
    const unsigned tb = blocks < 148u * 8u ? blocks : 148u * 8u;      // grid-stride: the table is built once per block
    if (dtype == BT_BF16) im2col2d_tab_kernel<__nv_bfloat16, 8><<<tb, threads, 0, st>>>(a);
    else im2col2d_tab_kernel<float, 4><<<tb, threads, 0, st>>>(a);
  } else if (dtype == BT_BF16) im2col2d_kernel<__nv_bfloat16, 8><<<blocks, threads, 0, st>>>(a);
  else im2col2d_kernel<float, 4><<<blocks, threads, 0, st>>>(a);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}
