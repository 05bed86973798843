// On-chip random numbers for the fused Bayesian layers: Philox4x32-10 (Salmon et al. SC'11,
// Random123), counter-based so that a draw depends only on (seed, layer, sample, element) and
// never on tiling, launch geometry or GPU count.  CPU statement: oracle/philox_ref.py.
//
// Replaces the ATen streams behind  eps.data.normal_()  (linear_variational.py:161,173) and
// x.clone().uniform_(-1,1).sign()  (linear_flipout.py:169-170) of the reference.
#pragma once
#include <stdint.h>

#define BT_STREAM_W_EPS 0u
#define BT_STREAM_B_EPS 1u
#define BT_STREAM_SIGN_IN 2u
#define BT_STREAM_SIGN_OUT 3u

struct BtRngKey {
  uint32_t k0, k1;   // seed lo / hi
  uint32_t c3_base;  // layer_key << 4
};

__device__ __forceinline__ uint4 bt_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                  uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Box-Muller on ONE 32-bit word: two 16-bit uniforms built by bit insertion (no I2F on the MUFU pipe)
//   u = 1 - (w & 0xffff) * 2^-16  in (0,1],   v = (w >> 16) * 2^-16  in [0,1)
//   r = sqrt(-2 ln u);  z0 = r cos(2 pi v);  z1 = r sin(2 pi v)          (|z| <= 4.71)
// so one Philox4x32 call yields EIGHT normals.  4 MUFU per pair (lg2, sqrt, sin, cos).
// CPU statement: oracle/philox_ref.py::box_muller8.
__device__ __forceinline__ void bt_box_muller16(uint32_t w, float& z0, float& z1) {
  // 2^23 + halfword as a float (one PRMT each), then ONE fma per uniform -- both exact:
  //   u     = (2^23 + lo) * -2^-16 + 129         = 1 - lo * 2^-16
  //   theta = (2^23 + hi) * c - 2^23 * c         = fl(hi * c),  c = fl(2 pi) * 2^-16   (== fl(fl(2 pi) * (hi * 2^-16)))
  const float lo_m = __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7610));
  const float hi_m = __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7632));
  const float u = fmaf(lo_m, -1.52587890625e-05f, 129.0f);
  const float th = fmaf(hi_m, 6.283185307179586f * 1.52587890625e-05f, -(6.283185307179586f * 128.0f));
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
  float r;  // sqrt(-2 ln u) = sqrt(-2 ln2 * log2(u)); MUFU sqrt instead of the IEEE sequence
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(-1.3862943611198906f * lg, 0.0f)));
  float s, c;
  __sincosf(th, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// 8 standard normals of weight "oct" (row n, elements 8*ko .. 8*ko+7) for one MC sample.
__device__ __forceinline__ void bt_eps_oct(const BtRngKey& key, uint32_t stream, uint32_t ko, uint32_t n,
                                           uint32_t sample, float (&z)[8]) {
  const uint4 r = bt_philox4x32_10(ko, n, sample, key.c3_base | stream, key.k0, key.k1);
  bt_box_muller16(r.x, z[0], z[1]);
  bt_box_muller16(r.y, z[2], z[3]);
  bt_box_muller16(r.z, z[4], z[5]);
  bt_box_muller16(r.w, z[6], z[7]);
}

// the 4 normals of quad kq (elements 4*kq .. 4*kq+3) = one half of oct kq >> 1 (generic / export paths)
__device__ __forceinline__ float4 bt_eps_quad(const BtRngKey& key, uint32_t stream, uint32_t kq,
                                              uint32_t n, uint32_t sample) {
  const uint4 r = bt_philox4x32_10(kq >> 1, n, sample, key.c3_base | stream, key.k0, key.k1);
  float4 z;
  bt_box_muller16((kq & 1u) ? r.z : r.x, z.x, z.y);
  bt_box_muller16((kq & 1u) ? r.w : r.y, z.z, z.w);
  return z;
}

// 128 sign bits (1 -> negative) of `row` (pixel or output row), 128-column block `blk`.
__device__ __forceinline__ uint4 bt_sign_block(const BtRngKey& key, uint32_t stream, uint32_t blk,
                                               uint32_t row, uint32_t sample) {
  return bt_philox4x32_10(blk, row, sample, key.c3_base | stream, key.k0, key.k1);
}

__device__ __forceinline__ uint32_t bt_sign_word(const uint4& b, uint32_t w) {
  return w == 0 ? b.x : (w == 1 ? b.y : (w == 2 ? b.z : b.w));
}
