// K2t: TMA-fed weight-stationary kernel -- the activation operand of tcgen05.mma is staged by the Tensor Memory
// Accelerator (cp.async.bulk.tensor, tiled or im2col tensor maps, 128B swizzle, mbarrier complete_tx), so no warp
// spends an instruction on gathering activations.  (Included by bt_fused.cu inside its anonymous namespace; shares
// FusedParams and the PTX wrappers.)
//
// Same reference op sequences as bt_fused_kernel (linear_variational.py:157-201, conv_variational.py:183-227 /
// 357-402 / 530-574): Reparameterization layers whose sampled weight tile [BLOCK_N x K] fits shared memory next to
// an activation ring.  Operands: bf16 (kind::f16) for bf16 activations, tf32 (kind::tf32, fp32 words in shared
// memory) for fp32 parameters with fp32 activations.
//
//   warps 0-7  sample W_s = mu + softplus(rho) * eps for every k-block of this CTA's (n-tile, MC sample) ONCE into the
//              resident region (same Philox counters / arithmetic as the other kernel families => same draws), then
//              become the epilogue: TMEM lane quarter = warp & 3, column half = warp >> 2;
//   warp  8    TMA producer: per (row tile, k-block) one cp.async.bulk.tensor into a ring of 16 KB stages --
//              linear layers / materialised-im2col stems: a 2-D tiled map over x [rows, K];
//              convolutions: an im2col map over x [N, (D,) H, W, C] (stride / padding / dilation in the map, the
//              filter tap as the instruction's im2col offsets, out-of-image taps zero-filled by the hardware);
//   warp  9    MMA issuer (whole warp runs the loop, elect.sync issues), two accumulator buffers in TMEM.
#include <cuda.h>   // CUtensorMap, cuTensorMapEncode* prototypes (resolved at run time through cudaGetDriverEntryPoint)

constexpr int TM_SAMP_WARPS = 8;
constexpr int TM_TMA_WARP = 8;
constexpr int TM_MMA_WARP = 9;
constexpr int TM_CONV_WARP0 = 10;     // tf32 only: warps 10-13 round the TMA-staged fp32 activations to tf32 in place
constexpr int TM_CONV_WARPS = 4;
// warps are allocated in fours: 12 warps x 168 registers (bf16: warps 10-11 idle) or 16 warps x 128 registers (tf32:
// warps 14-15 idle) fill the register file
template <bool TF32>
constexpr int tm_threads() { return TF32 ? 16 * 32 : 12 * 32; }
constexpr int TM_THREADS = 12 * 32;   // (bf16 instantiations; reported by the plan)
constexpr int TM_AUX_BYTES = 8192;    // barriers + up to 4 x [3][128] fp32 epilogue constants
constexpr int TM_MAX_NSMP = 4;
constexpr int DT_AUX_BYTES = 2560;    // bt_dtma_kernel: barriers + one [3][128] constant table

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmaDriver {
  PFN_encodeTiled tiled = nullptr;
  PFN_encodeIm2col im2col = nullptr;
  int driver_version = 0;
  int state = 0;   // 0 unknown, 1 ready, -1 unavailable
};
TmaDriver g_tma;

// libcuda is not linked: the two encoders are looked up once through the runtime.  Returns false (and the callers
// fall back to the cp.async kernel families) if the driver does not export them.
inline bool tma_driver_ready() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_tma.state != 0) return g_tma.state > 0;
  g_tma.state = -1;
  void* f1 = nullptr;
  void* f2 = nullptr;
  cudaDriverEntryPointQueryResult q1, q2;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1) != cudaSuccess ||
      q1 != cudaDriverEntryPointSuccess || f1 == nullptr) {
    cudaGetLastError();
    return false;
  }
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2) != cudaSuccess ||
      q2 != cudaDriverEntryPointSuccess || f2 == nullptr) {
    cudaGetLastError();
    return false;
  }
  g_tma.tiled = reinterpret_cast<PFN_encodeTiled>(f1);
  g_tma.im2col = reinterpret_cast<PFN_encodeIm2col>(f2);
  cudaDriverGetVersion(&g_tma.driver_version);
  g_tma.state = 1;
  return true;
}

// What the TMA path needs to know about the activation operand of one layer (pure host arithmetic, shared by the
// launch, the plan and the probe).
struct TmaAPlan {
  int mode;        // 1 = tiled 2-D (rows x K), 2 = im2col
  int nd;          // im2col: spatial dims that are not degenerate from the left (1, 2 or 3)
  int kbe;         // elements per k-block (64 bf16 | 32 tf32) = one 128-byte swizzle row
  int es;          // element size
  long long n_img; // images in x ((x_shared ? 1 : S) * B)
};

// Is this layer's A operand expressible as one TMA box per (row tile, k-block)?
inline bool tma_a_plan(const FusedParams& p, bool tf32, TmaAPlan* out) {
  TmaAPlan a;
  a.es = p.x_is_bf16 ? 2 : 4;
  a.kbe = 128 / a.es;
  if (!p.x_is_bf16 && !tf32) return false;          // fp32 activations need the tf32 operand path
  a.n_img = (long long)(p.x_shared ? 1 : p.S) * p.B;
  if (((long long)p.C_in * a.es) % 16 != 0) return false;
  const int taps_all = p.KD * p.KH * p.KW;
  const bool linear_like = taps_all == 1 && p.sd == 1 && p.sh == 1 && p.sw == 1 && p.pd == 0 && p.ph == 0 && p.pw == 0;
  if (linear_like && (p.groups == 1 || p.Cin_g % a.kbe == 0)) {
    a.mode = 1;
    a.nd = 0;
    *out = a;
    return true;
  }
  if (p.Cin_g % a.kbe != 0) return false;           // a k-block must not straddle filter taps
  a.mode = 2;
  a.nd = p.ID > 1 || p.KD > 1 || p.OD > 1 ? 3 : (p.IH > 1 || p.KH > 1 || p.OH > 1 ? 2 : 1);
  // corner / offset field widths of the im2col descriptor and instruction (16 bits shared by the spatial dims)
  const int cbits = a.nd == 1 ? 16 : (a.nd == 2 ? 8 : 5);
  const int clim = 1 << (cbits - 1);
  const int sp_in[3] = {p.IW, p.IH, p.ID}, sp_k[3] = {p.KW, p.KH, p.KD}, sp_p[3] = {p.pw, p.ph, p.pd},
            sp_d[3] = {p.dw, p.dh, p.dd}, sp_s[3] = {p.sw, p.sh, p.sd};
  for (int i = 0; i < a.nd; ++i) {
    const int lower = -sp_p[i], upper = sp_p[i] - (sp_k[i] - 1) * sp_d[i];
    if (lower < -clim || lower > clim - 1 || upper < -clim || upper > clim - 1) return false;
    if ((sp_k[i] - 1) * sp_d[i] >= (1 << cbits)) return false;
    if (sp_s[i] > 8) return false;                  // traversal stride field: 3 bits (1..8)
    (void)sp_in;
  }
  for (int i = a.nd; i < 3; ++i)
    if (sp_k[i] != 1 || sp_p[i] != 0 || sp_in[i] != 1) return false;
  if (a.n_img >= (1ll << 31)) return false;
  *out = a;
  return true;
}

// Encode the tensor map of x for this layer.  x: channels-last activations [n_img, ID, IH, IW, C_in].
inline int tma_encode_a(const FusedParams& p, const TmaAPlan& a, const void* x, CUtensorMap* map) {
  BT_REQUIRE(tma_driver_ready(), BT_ERR_UNSUPPORTED, "TMA: cuTensorMapEncode* not available from this driver");
  const CUtensorMapDataType dt = p.x_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r;
  if (a.mode == 1) {
    const long long rows = a.n_img * p.ID * p.IH * p.IW;
    cuuint64_t dims[2] = {(cuuint64_t)p.C_in, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)p.C_in * a.es};
    cuuint32_t box[2] = {(cuuint32_t)a.kbe, (cuuint32_t)BLOCK_M};
    cuuint32_t estr[2] = {1, 1};
    r = g_tma.tiled(map, dt, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const int rank = a.nd + 2;
    const int sp_in[3] = {p.IW, p.IH, p.ID}, sp_k[3] = {p.KW, p.KH, p.KD}, sp_p[3] = {p.pw, p.ph, p.pd},
              sp_d[3] = {p.dw, p.dh, p.dd}, sp_s[3] = {p.sw, p.sh, p.sd};
    cuuint64_t dims[5], strides[4];
    cuuint32_t trav[5];
    int lower[3], upper[3];
    dims[0] = (cuuint64_t)p.C_in;
    trav[0] = 1;
    unsigned long long st = (unsigned long long)p.C_in * a.es;
    for (int i = 0; i < a.nd; ++i) {
      dims[1 + i] = (cuuint64_t)sp_in[i];
      strides[i] = st;
      st *= (unsigned long long)sp_in[i];
      trav[1 + i] = (cuuint32_t)sp_s[i];
      lower[i] = -sp_p[i];
      upper[i] = sp_p[i] - (sp_k[i] - 1) * sp_d[i];
    }
    dims[1 + a.nd] = (cuuint64_t)a.n_img;
    strides[a.nd] = st;
    trav[1 + a.nd] = 1;
    r = g_tma.im2col(map, dt, (cuuint32_t)rank, const_cast<void*>(x), dims, strides, lower, upper, (cuuint32_t)a.kbe,
                     (cuuint32_t)BLOCK_M, trav, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // Drivers up to CUDA 13.1 mis-set a descriptor bit for tensors smaller than 128 KiB (same fix-up as CUTLASS,
    // cute/atom/copy_traits_sm90_im2col.hpp)
    const unsigned long long total = st * (unsigned long long)a.n_img;
    if (r == CUDA_SUCCESS && g_tma.driver_version <= 13010 && total < 131072ull)
      reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  }
  BT_REQUIRE(r == CUDA_SUCCESS, BT_ERR_CUDA, "TMA: cuTensorMapEncode%s failed with CUresult %d", a.mode == 1 ? "Tiled" : "Im2col", (int)r);
  return BT_OK;
}

// Kernel-side view of the A operand (part of TmaParams).
struct TmaA {
  int mode, nd, kbe;
  int slabs;        // k-blocks per filter tap (Cin_g / kbe); tiled mode: unused
  int cln;          // bt_tms_kernel: thread-block cluster size along the n-tiles (1 | 2): the CTAs of a cluster read the SAME
                    // activation tiles, each loads 1/cln of them and multicasts (L2 -> SM traffic / cln)
  int probe;        // measurement switches (BT_TMA_PROBE, never set in production): 1 = the samplers skip their arithmetic
                    // (stale weights): what is left is the MMA / TMA / epilogue time of the launch
  int nsmp;         // bt_tma_kernel: MC samples per CTA (> 1 only when every sample reads the same x: each staged
                    // activation tile is multiplied with the resident W_s of nsmp samples -> 1/nsmp of the L2 traffic)
};

struct __align__(64) TmaParams {
  CUtensorMap map_a;
  FusedParams f;
  TmaA a;
};

struct DtGeom {          // host-computed window geometry of bt_dtma_kernel
  int hb;                // padded rows per BODY box inside one plane (divides Ph); == Ph when the box spans whole planes
  int nbp;               // whole padded planes per body box (2-D convolutions with tiny images), else 1
  int unit;              // padded rows per body box = hb (nbp == 1) or nbp * Ph
  int k;                 // padded rows per tile (multiple of unit), k * Pw <= 128
  int hr;                // halo rows on each side, each staged by a one-row box
  int nbox;              // TMA boxes per window and slab = k / unit + 2 hr (k / unit when halo == 0)
  int halo;              // 1: the halo rows are staged by one-row boxes; 0: tiles are whole padded planes (2-D, hb == Ph):
                         // the rows above a tile are the previous image's zero rows / a zeroed region of the slot, the rows
                         // below are only read by the discarded pad-row outputs -- no halo boxes at all
  int R;                 // rows (128 B each) of one slab plane of a window slot
  int Z;                 // permanently-zero rows in front of the data (>= pw)
  int Pw, Ph, Pd;        // padded extents W + pw, H + ph, D + pd
  long long NR;          // padded rows per sample = B * Pd * Ph
  int slots;             // window ring depth
  uint32_t mulw, shw, mulh, shh, muld, shd;   // reciprocals of Pw, Ph, Pd (n / d == (n * mul) >> sh for n < 2^31)
};

struct __align__(64) DtParams {
  CUtensorMap map_a;     // body boxes {kbe, Pw, hb, 1, nbp}
  CUtensorMap map_h;     // halo boxes {kbe, Pw, 1, 1, 1}
  FusedParams f;
  DtGeom g;
  int kbe, slabs;
  int clx;     // thread-block cluster size along x (the CTAs of one (sample, n-tile)): they split the k-blocks of the
               // W_s sampling prologue and write their tiles into every peer's shared memory (0 / 1 = off)
};

#ifdef BT_TMA_DEVICE
#ifndef BT_TM_PINGPONG
#define BT_TM_PINGPONG 0      // sampler k loop unrolled by two over two parameter register sets: measured SLOWER (3-10%,
                              // profiles/r02l: twice the loop code, instruction-fetch stalls) than one register copy per k-block
#endif
#ifndef BT_TM_PREFETCH
#define BT_TM_PREFETCH 1      // sampler: parameter words of k-block kb+1 are loaded while kb is sampled (second register set)
#endif
#ifndef BT_TM_PIN_ROWOFF
#define BT_TM_PIN_ROWOFF 1    // sampler: keep the per-row parameter offsets instead of re-deriving them per k-block
#endif
// ------------------------------------------------------------------ device: TMA wrappers (elect.sync inside)
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                  uint16_t mask = 0) {
  if (mask) {      // the box lands at the same offset of every CTA in `mask` and completes on each one's barrier
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;\n\t}\n" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
    return;
  }
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_3d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w,
                                                          int n, uint16_t ow, uint16_t mask = 0) {
  if (mask) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.3d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], {%6}, %7;\n\t}\n" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(n), "h"(ow), "h"(mask)
        : "memory");
    return;
  }
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.3d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2], {%6};\n\t}\n" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(n), "h"(ow)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w,
                                                          int h, int n, uint16_t ow, uint16_t oh, uint16_t mask = 0) {
  if (mask) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;\n\t}\n" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh), "h"(mask)
        : "memory");
    return;
  }
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n\t}\n" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_5d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w,
                                                          int h, int d, int n, uint16_t ow, uint16_t oh, uint16_t od,
                                                          uint16_t mask = 0) {
  if (mask) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10}, %11;\n\t}\n" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(ow),
            "h"(oh), "h"(od), "h"(mask)
        : "memory");
    return;
  }
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};\n\t}\n" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(ow),
          "h"(oh), "h"(od)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Issue the TMA load of the A tile (128 output rows starting at row m0 of MC sample s, k-block = (tap_i, slab)) into
// `dst`, completing on `bar`.  Executed by the whole TMA warp with warp-uniform arguments.
__device__ __forceinline__ void tma_issue_a(const TmaParams& tp, uint32_t dst, uint32_t bar, int img_base, int g,
                                            long long m0, int tap_i, int slab, int b, int od, int oh, int ow, uint16_t mask = 0) {
  const FusedParams& p = tp.f;
  if (tp.a.mode == 1) {
    const long long row = (long long)img_base * p.ID * p.IH * p.IW + m0;     // (linear-like: one row per pixel)
    tma_load_2d_elect(dst, &tp.map_a, bar, g * p.Cin_g + slab * tp.a.kbe, (int)row, mask);
    return;
  }
  const uint32_t t = p.taps[tap_i];
  const int kd = t & 0xff, kh = (t >> 8) & 0xff, kw = (t >> 16) & 0xff;
  const int c = g * p.Cin_g + slab * tp.a.kbe;
  const int n = img_base + b;
  if (tp.a.nd == 1)
    tma_load_im2col_3d_elect(dst, &tp.map_a, bar, c, ow * p.sw - p.pw, n, (uint16_t)(kw * p.dw), mask);
  else if (tp.a.nd == 2)
    tma_load_im2col_4d_elect(dst, &tp.map_a, bar, c, ow * p.sw - p.pw, oh * p.sh - p.ph, n, (uint16_t)(kw * p.dw),
                             (uint16_t)(kh * p.dh), mask);
  else
    tma_load_im2col_5d_elect(dst, &tp.map_a, bar, c, ow * p.sw - p.pw, oh * p.sh - p.ph, od * p.sd - p.pd, n,
                             (uint16_t)(kw * p.dw), (uint16_t)(kh * p.dh), (uint16_t)(kd * p.dd), mask);
}

// ------------------------------------------------------------------ probe: one A tile through TMA -> global (tests)
__global__ void __launch_bounds__(32, 1) bt_tma_probe_kernel(const __grid_constant__ TmaParams tp, long long m0, int s, int g,
                                                             int tap_i, int slab, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_TILE_BYTES);
  const FusedParams& p = tp.f;
  const int lane = threadIdx.x;
  if (lane == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  __syncwarp();
  const long long out_sp = (long long)p.OD * p.OH * p.OW;
  const int b = (int)(m0 / out_sp);
  long long rem = m0 - (long long)b * out_sp;
  const int od = (int)(rem / ((long long)p.OH * p.OW));
  rem -= (long long)od * p.OH * p.OW;
  const int oh = (int)(rem / p.OW), ow = (int)(rem - (long long)oh * p.OW);
  const int img_base = p.x_shared ? 0 : s * p.B;
  mbar_expect_tx_elect(smem_u32(bar), A_TILE_BYTES);
  tma_issue_a(tp, smem_u32(smem), smem_u32(bar), img_base, g, m0, tap_i, slab, b, od, oh, ow);
  mbar_wait(smem_u32(bar), 0);
  for (int i = lane; i < A_TILE_BYTES / 16; i += 32)
    reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(smem)[i];
}

// Generic probe of a tiled 4-D map (C, W, H, N): box {bc, bw, bh, 1} at coords {c, w, h, n} -> shared memory at
// byte offset dst_off (a multiple of 128, NOT necessarily of 1024) of a 1024-aligned 32 KB buffer; the whole buffer is
// copied out.  Pins two hardware facts the window loader of the direct kernel relies on: out-of-range coordinates are
// zero-filled, and the 128B-swizzle XOR is a function of the absolute shared-memory address.
__global__ void __launch_bounds__(32, 1) bt_tma_probe4d_kernel(const __grid_constant__ CUtensorMap map, int c, int w, int h, int n,
                                                               uint32_t dst_off, uint32_t bytes, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  const int lane = threadIdx.x;
  for (int i = lane; i < 32768 / 16; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0xA5A5A5A5u, 0xA5A5A5A5u, 0xA5A5A5A5u, 0xA5A5A5A5u);
  if (lane == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncwarp();
  mbar_expect_tx_elect(smem_u32(bar), bytes);
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}\n" ::
          "r"(smem_u32(smem) + dst_off), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
  mbar_wait(smem_u32(bar), 0);
  for (int i = lane; i < 32768 / 16; i += 32)
    reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(smem)[i];
}

// ------------------------------------------------------------------ shared pieces of the two TMA kernels
// Sampler of one [BLOCK_N x KBE] weight tile: 256 threads (warps 0-7); thread = oct `wo` (8 consecutive k = one Philox
// call) of rows wrb + RPP*i.  Same counters (kphys >> 3, row, sample, stream) and arithmetic as every other kernel
// family, so bt_rng_export re-materialises exactly these draws.
template <int BLOCK_N, bool P_BF16, bool TF32, bool FLIP = false>
struct TmSampler {
  static constexpr int KBE = TF32 ? 32 : 64;
  static constexpr int OPR = KBE / 8;
  static constexpr int RPP = (TM_SAMP_WARPS * 32) / OPR;    // rows per pass: 32 | 64
  static constexpr int WO = (BLOCK_N + RPP - 1) / RPP;      // octs per thread
  static constexpr int PW = P_BF16 ? 4 : 8;
  static constexpr int P_ES = P_BF16 ? 2 : 4;
  int wo, wrb;
  long long row_off[WO];
  bool nvalid[WO], rvalid[WO];
  uint32_t nrow[WO];

  __device__ __forceinline__ void init(const FusedParams& p, int tid, int g, int n0) {
    wo = tid % OPR;
    wrb = tid / OPR;
#pragma unroll
    for (int i = 0; i < WO; ++i) {
      const int nl = wrb + RPP * i;
      const int n = n0 + nl;
      rvalid[i] = nl < BLOCK_N;
      nvalid[i] = rvalid[i] && n < p.N;
      row_off[i] = ((long long)g * p.N + (nvalid[i] ? n : p.N - 1)) * p.K_phys * P_ES;   // bytes
      nrow[i] = (uint32_t)(g * p.N + n);
      // opaque to the optimiser: ptxas otherwise re-derives the row offsets (compares, selects, 64-bit multiplies)
      // in every k-block instead of keeping them (~100 of ~850 instructions per thread and k-block)
#if BT_TM_PIN_ROWOFF
      asm volatile("" : "+l"(row_off[i]), "+r"(nrow[i]));
#endif
    }
  }
  // k offset (inside the k-block) of this thread's oct
  __device__ __forceinline__ int koff() const { return wo * 8; }

  // Two register sets of parameter words (PH = 0 / 1): the loads of k-block kb+1 go into the set that is not being
  // turned into weights, so their L2 latency (~600+ clocks, exposed once per k-block with only two sampler warps per
  // scheduler) hides behind the Philox / Box-Muller arithmetic of k-block kb.  The k loops are unrolled by two
  // (tm_sample_loop) so that the set index is a compile-time constant and no register is copied.
  uint32_t mu_r[2][WO][PW], rho_r[2][WO][PW];
  long long k_of[2];
  bool v_of[2];
  uint32_t ncl = 1;       // > 1: every sampled chunk is written to the same offset of all `ncl` CTAs of the cluster (DSMEM)
  __device__ __forceinline__ void put16(uint32_t addr, const uint4& v) const {
    if (ncl <= 1) {
      sts16(addr, v);
    } else {
      for (uint32_t c = 0; c < ncl; ++c) bt_sts16_cluster(addr, c, v);
    }
  }

  template <int PH>
  __device__ __forceinline__ void prefetch(const FusedParams& p, long long kphys0, bool kvalid) {
    const uint8_t* mu_w = static_cast<const uint8_t*>(p.mu_w);
    const uint8_t* rho_w = static_cast<const uint8_t*>(p.rho_w);
    k_of[PH] = kvalid ? kphys0 : 0;
    v_of[PH] = kvalid;
#pragma unroll
    for (int i = 0; i < WO; ++i) {
      const long long off = row_off[i] + k_of[PH] * P_ES;
      const uint4 a = ldg16(mu_w + off);
      const uint4 b = ldg16(rho_w + off);
      mu_r[PH][i][0] = a.x; mu_r[PH][i][1] = a.y; mu_r[PH][i][2] = a.z; mu_r[PH][i][3] = a.w;
      rho_r[PH][i][0] = b.x; rho_r[PH][i][1] = b.y; rho_r[PH][i][2] = b.z; rho_r[PH][i][3] = b.w;
      if constexpr (!P_BF16) {
        const uint4 a2 = ldg16(mu_w + off + 16);
        const uint4 b2 = ldg16(rho_w + off + 16);
        mu_r[PH][i][4] = a2.x; mu_r[PH][i][5] = a2.y; mu_r[PH][i][6] = a2.z; mu_r[PH][i][7] = a2.w;
        rho_r[PH][i][4] = b2.x; rho_r[PH][i][5] = b2.y; rho_r[PH][i][6] = b2.z; rho_r[PH][i][7] = b2.w;
      }
    }
  }

  __device__ __forceinline__ void advance() {     // set 1 (prefetched) becomes set 0 (current)
    k_of[0] = k_of[1];
    v_of[0] = v_of[1];
#pragma unroll
    for (int i = 0; i < WO; ++i) {
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        mu_r[0][i][j] = mu_r[1][i][j];
        rho_r[0][i][j] = rho_r[1][i][j];
      }
    }
  }

  // turn the parameter words of set PH into the weight tile(s) at sb: shared-memory tile [BLOCK_N][128 B]
  template <int PH>
  __device__ __forceinline__ void compute(const FusedParams& p, uint32_t smp, uint32_t sb) const {
    const long long kl = k_of[PH];
    const bool kvalid = v_of[PH];
    uint32_t c[WO][4];
#pragma unroll
    for (int i = 0; i < WO; ++i) {
      c[i][0] = (uint32_t)(kl >> 3);
      c[i][1] = nrow[i];
      c[i][2] = smp;
      c[i][3] = p.key.c3_base | BT_STREAM_W_EPS;
    }
    philox_multi<WO>(c, p.key.k0, p.key.k1);
#pragma unroll
    for (int i = 0; i < WO; ++i) {
      float e[8], m8[8], r8[8];
      bt_box_muller16(c[i][0], e[0], e[1]);
      bt_box_muller16(c[i][1], e[2], e[3]);
      bt_box_muller16(c[i][2], e[4], e[5]);
      bt_box_muller16(c[i][3], e[6], e[7]);
      if constexpr (P_BF16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m8[2 * j] = bt_bf16_lo(mu_r[PH][i][j]);
          m8[2 * j + 1] = bt_bf16_hi(mu_r[PH][i][j]);
          r8[2 * j] = bt_bf16_lo(rho_r[PH][i][j]);
          r8[2 * j + 1] = bt_bf16_hi(rho_r[PH][i][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          m8[j] = __uint_as_float(mu_r[PH][i][j]);
          r8[j] = __uint_as_float(rho_r[PH][i][j]);
        }
      }
      const bool ok = kvalid && nvalid[i];
      float w0[8], w1[8];
      if (!p.rho_is_sigma) {   // (warp-uniform) tf32: the full-precision softplus (parity at 1e-4), bf16: the fast one
#pragma unroll
        for (int j = 0; j < 8; ++j) r8[j] = TF32 ? bt_softplus(r8[j]) : bt_softplus_fast(r8[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if constexpr (FLIP) {            // Flipout: the mean tile and the perturbation tile sigma * eps
          w0[j] = ok ? m8[j] : 0.f;
          w1[j] = ok ? r8[j] * e[j] : 0.f;
        } else {
          w0[j] = ok ? fmaf(r8[j], e[j], m8[j]) : 0.f;
        }
      }
      const int nl = wrb + RPP * i;
      if (rvalid[i]) {
#pragma unroll
        for (int t = 0; t < (FLIP ? 2 : 1); ++t) {
          const float* w = t == 0 ? w0 : w1;
          const uint32_t tb = sb + (uint32_t)(t * BLOCK_N * 128);
          if constexpr (TF32) {
            const uint32_t r0 = tb + (uint32_t)(nl * 128);
            put16(r0 + (uint32_t)(((2 * wo) ^ (nl & 7)) << 4), make_uint4(bt_tf32(w[0]), bt_tf32(w[1]), bt_tf32(w[2]), bt_tf32(w[3])));
            put16(r0 + (uint32_t)(((2 * wo + 1) ^ (nl & 7)) << 4),
                  make_uint4(bt_tf32(w[4]), bt_tf32(w[5]), bt_tf32(w[6]), bt_tf32(w[7])));
          } else {
            put16(tb + (uint32_t)(nl * 128 + ((wo ^ (nl & 7)) << 4)),
                  make_uint4(bt_pack_bf16x2(w[0], w[1]), bt_pack_bf16x2(w[2], w[3]), bt_pack_bf16x2(w[4], w[5]),
                             bt_pack_bf16x2(w[6], w[7])));
          }
        }
      }
    }
  }
};

// k loop of a sampler: `next(ph)` issues the parameter loads of the following k-block into register set ph (a
// std::integral_constant), `body(ph, kb)` turns set ph into the tile(s) of k-block kb.  Unrolled by two: set indices
// are compile-time constants.
template <int V> struct TmPh { static constexpr int value = V; };
// k0 / kstep: the k-blocks this CTA samples (a cluster splits them: k0 = rank, kstep = cluster size); `next` is called
// once per sampled k-block, in order
template <class Smp, class Next, class Body>
__device__ __forceinline__ void tm_sample_loop_strided(Smp& smp, int num_kb, int k0, int kstep, Next&& next, Body&& body) {
  if (k0 >= num_kb) return;
  next(TmPh<1>{});
#pragma unroll 1
  for (int kb = k0; kb < num_kb; kb += kstep) {
    smp.advance();
    if (kb + kstep < num_kb) next(TmPh<1>{});
    body(TmPh<0>{}, kb);
  }
}
template <class Smp, class Next, class Body>
__device__ __forceinline__ void tm_sample_loop(Smp& smp, int num_kb, Next&& next, Body&& body) {
#if BT_TM_PINGPONG
  next(TmPh<0>{});
#pragma unroll 1
  for (int kb = 0; kb < num_kb; kb += 2) {
    if (kb + 1 < num_kb) next(TmPh<1>{});              // in flight while k-block kb is sampled
    body(TmPh<0>{}, kb);
    if (kb + 1 < num_kb) {
      if (kb + 2 < num_kb) next(TmPh<0>{});
      body(TmPh<1>{}, kb + 1);
    }
  }
#elif BT_TM_PREFETCH
  next(TmPh<1>{});
#pragma unroll 1
  for (int kb = 0; kb < num_kb; ++kb) {
    smp.advance();                                     // set 1 -> set 0 (register copies; half the loop's code size)
    if (kb + 1 < num_kb) next(TmPh<1>{});              // in flight while k-block kb is sampled
    body(TmPh<0>{}, kb);
  }
#else
#pragma unroll 1
  for (int kb = 0; kb < num_kb; ++kb) {               // one register set: the loads sit in front of this k-block's Philox /
    next(TmPh<0>{});                                   // Box-Muller arithmetic, which needs them only for its last fma
    body(TmPh<0>{}, kb);
  }
#endif
}

// per-column constants of the epilogue in shared memory: [0,128) bias, [128,256) scale, [256,384) bias*scale + shift
// (Flipout: [0,128) mean bias mu_b, [128,256) scale, [256,384) shift, [384,512) perturbation bias sigma_b * eps_b)
template <bool P_BF16, bool FLIP = false>
__device__ __forceinline__ void tm_fill_bias(const FusedParams& p, float* bias_s, int tid, int g, int n0, uint32_t sample) {
  const int n = n0 + tid;
  float b0 = 0.f, b1 = 0.f, sc = 1.f, sh = 0.f;
  if (n < p.N) {
    const int ng = g * p.N + n;
    if (p.mu_b != nullptr) {
      float mu, rho;
      if (P_BF16) {
        mu = __bfloat162float(static_cast<const __nv_bfloat16*>(p.mu_b)[ng]);
        rho = __bfloat162float(static_cast<const __nv_bfloat16*>(p.rho_b)[ng]);
      } else {
        mu = static_cast<const float*>(p.mu_b)[ng];
        rho = static_cast<const float*>(p.rho_b)[ng];
      }
      const float4 z = bt_eps_quad(p.key, BT_STREAM_B_EPS, (uint32_t)(ng >> 2), 0u, sample);
      const int j = ng & 3;
      const float eps = j == 0 ? z.x : (j == 1 ? z.y : (j == 2 ? z.z : z.w));
      if constexpr (FLIP) {
        b0 = mu;
        b1 = bt_softplus(rho) * eps;
      } else {
        b0 = mu + bt_softplus(rho) * eps;
      }
    }
    if (p.ep_scale != nullptr) {
      sc = __ldg(p.ep_scale + ng);
      sh = __ldg(p.ep_shift + ng);
    }
  }
  // out = (acc + b) * sc + sh  ==  fma(acc, sc, b * sc + sh): one FMA and two constants per element
  bias_s[tid] = b0;
  bias_s[128 + tid] = sc;
  bias_s[256 + tid] = FLIP ? sh : fmaf(b0, sc, sh);
  if constexpr (FLIP) bias_s[384 + tid] = b1;
}

// Flipout: acc0 / acc1 of one 16-column group -> (acc0 + mu_b) + s_out * (acc1 + sigma_b eps_b), then the affine
// (reference: conv_flipout.py:430-433 `outputs + perturbed_outputs`); sblk = the 128 output sign bits of this row and
// 128-column block, nbit0 = (n0 & 127) + column of o[0].  v1 = the perturbation accumulator.
__device__ __forceinline__ void tm_flip_combine16(const float* bias_s, int col0, const uint32_t (&v0)[16], const uint32_t (&v1)[16],
                                                  const uint4& sblk, int nbit0, bool has_affine, float (&o)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int bit = nbit0 + j;
    const bool neg = (bt_sign_word(sblk, bit >> 5) >> (bit & 31)) & 1u;
    const float pert = __uint_as_float(v1[j]) + bias_s[384 + col0 + j];
    float val = __uint_as_float(v0[j]) + bias_s[col0 + j] + (neg ? -pert : pert);
    if (has_affine) val = fmaf(val, bias_s[128 + col0 + j], bias_s[256 + col0 + j]);
    o[j] = val;
  }
}

// 16 accumulator columns [col0, col0+16) of this lane's row: TMEM -> (+bias, affine, residual, ReLU) -> global.
// orow: output row (s * M + m); mvalid: the row exists.
// 16 accumulator columns of one row -> 16 outputs in the output dtype, as 16-byte chunks (bf16: 2, fp32: 4):
//   y = max(fma(acc, scale, shift') [+ residual], floor),  floor = 0 with a ReLU, -inf without.
// cst = shared-memory address of column 0's bias slot (scale at +512 bytes, shift' = bias * scale + shift at +1024;
// scale is 1 without a folded affine, so fma(acc, 1, bias) == acc + bias exactly).  Packed pairs throughout: 8 FFMA2,
// (8 FADD2,) and for bf16 outputs the ReLU runs on the packed words (max commutes with the monotonic rounding): per 16
// columns 8 + 8 + 8 issue slots instead of 16 + 16 + 8 -- the epilogue warps are issue / latency bound with two warps
// per scheduler (profiles/r02j: 620 instructions per warp and 128x64 tile before this form).
template <bool TF32>
__device__ __forceinline__ void tm_out16(const uint32_t (&v)[16], uint32_t cst, const uint4* res, uint32_t floor_bits,
                                         uint4 (&ch)[TF32 ? 4 : 2]) {
  uint64_t o2[8];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const uint4 sc = lds16(cst + 512 + 16 * jj), sh = lds16(cst + 1024 + 16 * jj);
    o2[2 * jj] = bt_ffma2(bt_pk2u(v[4 * jj], v[4 * jj + 1]), bt_pk2u(sc.x, sc.y), bt_pk2u(sh.x, sh.y));
    o2[2 * jj + 1] = bt_ffma2(bt_pk2u(v[4 * jj + 2], v[4 * jj + 3]), bt_pk2u(sc.z, sc.w), bt_pk2u(sh.z, sh.w));
  }
  if (res != nullptr) {
    if constexpr (TF32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o2[2 * k] = bt_fadd2(o2[2 * k], bt_pk2u(res[k].x, res[k].y));
        o2[2 * k + 1] = bt_fadd2(o2[2 * k + 1], bt_pk2u(res[k].z, res[k].w));
      }
    } else {
      const uint32_t w[8] = {res[0].x, res[0].y, res[0].z, res[0].w, res[1].x, res[1].y, res[1].z, res[1].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o2[j] = bt_fadd2(o2[j], bt_pk2u(w[j] << 16, w[j] & 0xffff0000u));
    }
  }
  uint32_t f[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) bt_upk2(o2[j], f[2 * j], f[2 * j + 1]);
  if constexpr (TF32) {
#pragma unroll
    for (int j = 0; j < 16; ++j) asm("max.NaN.f32 %0, %0, %1;" : "+r"(f[j]) : "r"(floor_bits));
#pragma unroll
    for (int k = 0; k < 4; ++k) ch[k] = make_uint4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
  } else {
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = bt_pack_bf16x2(__uint_as_float(f[2 * j]), __uint_as_float(f[2 * j + 1]));
      asm("max.NaN.bf16x2 %0, %0, %1;" : "+r"(h[j]) : "r"(floor_bits));
    }
    ch[0] = make_uint4(h[0], h[1], h[2], h[3]);
    ch[1] = make_uint4(h[4], h[5], h[6], h[7]);
  }
}
// floor operand of tm_out16: 0 (ReLU) or -inf, as fp32 / packed bf16 bits
template <bool TF32>
__device__ __forceinline__ uint32_t tm_floor_bits(bool relu) {
  return relu ? 0u : (TF32 ? 0xff800000u : 0xff80ff80u);
}

template <bool TF32, bool FLIP = false>
__device__ __forceinline__ void tm_epilogue16(const FusedParams& p, const float* bias_s, uint32_t taddr, int g, int n0,
                                              int col0, long long orow, bool mvalid, uint32_t flip_off = 0u,
                                              uint4 sblk = make_uint4(0u, 0u, 0u, 0u)) {
  constexpr int O_ES = TF32 ? 4 : 2;
  uint8_t* outb = static_cast<uint8_t*>(p.out);
  const uint8_t* resb = static_cast<const uint8_t*>(p.ep_residual);
  uint32_t v0[16], v1[16];
  tmem_ld16(taddr, v0);
  if constexpr (FLIP) tmem_ld16(taddr + flip_off, v1);
  const int nfirst = n0 + col0;
  const long long eoff = orow * p.C_out + g * p.N + nfirst;
  const bool vec_ok = p.out_vec && nfirst + 16 <= p.N;
  uint4 rr[4];
  const bool res_vec = p.ep_residual != nullptr && mvalid && vec_ok;
  if (res_vec) {                               // in flight during the TMEM round trip
    const uint8_t* rsd = resb + eoff * O_ES;
#pragma unroll
    for (int j = 0; j < (TF32 ? 4 : 2); ++j) rr[j] = ldg16(rsd + 16 * j);
  }
  tmem_ld_wait();
  if constexpr (!FLIP) {
    if (vec_ok) {                              // full 16-column group: packed arithmetic, 16-byte stores
      if (!mvalid) return;
      uint4 ch[TF32 ? 4 : 2];
      tm_out16<TF32>(v0, smem_u32(bias_s) + (uint32_t)(col0 * 4), res_vec ? rr : nullptr, tm_floor_bits<TF32>(p.ep_relu != 0), ch);
      uint4* dst4 = reinterpret_cast<uint4*>(outb + eoff * O_ES);
#pragma unroll
      for (int j = 0; j < (TF32 ? 4 : 2); ++j) dst4[j] = ch[j];
      return;
    }
  }
  float o[16];
  const bool has_affine = p.ep_scale != nullptr;
  if constexpr (FLIP) {
    tm_flip_combine16(bias_s, col0, v0, v1, sblk, (n0 & 127) + col0, has_affine, o);
  } else {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float4 sh = *reinterpret_cast<const float4*>(bias_s + 256 + col0 + 4 * jj);
      float v[4] = {__uint_as_float(v0[4 * jj]), __uint_as_float(v0[4 * jj + 1]), __uint_as_float(v0[4 * jj + 2]),
                    __uint_as_float(v0[4 * jj + 3])};
      if (has_affine) {
        const float4 sc = *reinterpret_cast<const float4*>(bias_s + 128 + col0 + 4 * jj);
        v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y);
        v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w);
      } else {
        v[0] += sh.x; v[1] += sh.y; v[2] += sh.z; v[3] += sh.w;     // shift' = bias
      }
      o[4 * jj] = v[0]; o[4 * jj + 1] = v[1]; o[4 * jj + 2] = v[2]; o[4 * jj + 3] = v[3];
    }
  }
  if (!mvalid) return;
  uint8_t* dst = outb + eoff * O_ES;
  if (p.ep_residual != nullptr) {
    if (res_vec) {
      if constexpr (TF32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[4 * j] += __uint_as_float(rr[j].x); o[4 * j + 1] += __uint_as_float(rr[j].y);
          o[4 * j + 2] += __uint_as_float(rr[j].z); o[4 * j + 3] += __uint_as_float(rr[j].w);
        }
      } else {
        const uint32_t w[8] = {rr[0].x, rr[0].y, rr[0].z, rr[0].w, rr[1].x, rr[1].y, rr[1].z, rr[1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[2 * j] += bt_bf16_lo(w[j]);
          o[2 * j + 1] += bt_bf16_hi(w[j]);
        }
      }
    } else {
      const uint8_t* rsd = resb + eoff * O_ES;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (nfirst + j < p.N)
          o[j] += TF32 ? reinterpret_cast<const float*>(rsd)[j]
                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rsd)[j]);
    }
  }
  if (p.ep_relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
  }
  if (vec_ok) {
    if constexpr (TF32) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    } else {
      reinterpret_cast<uint4*>(dst)[0] = make_uint4(bt_pack_bf16x2(o[0], o[1]), bt_pack_bf16x2(o[2], o[3]),
                                                    bt_pack_bf16x2(o[4], o[5]), bt_pack_bf16x2(o[6], o[7]));
      reinterpret_cast<uint4*>(dst)[1] = make_uint4(bt_pack_bf16x2(o[8], o[9]), bt_pack_bf16x2(o[10], o[11]),
                                                    bt_pack_bf16x2(o[12], o[13]), bt_pack_bf16x2(o[14], o[15]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (nfirst + j < p.N) {
        if constexpr (TF32) reinterpret_cast<float*>(dst)[j] = o[j];
        else reinterpret_cast<__nv_bfloat16*>(dst)[j] = __float2bfloat16_rn(o[j]);
      }
    }
  }
}

// Epilogue of one warp's share of a tile: 32 accumulator rows (lane = row) x EN columns.  With a staging buffer
// (`stg` != 0: 32 rows x EN outputs, 16-byte chunks XOR-swizzled) global memory is touched ROW-CONTIGUOUSLY -- CPR
// consecutive lanes cover one row's EN outputs -- instead of 32 lanes writing 32 different rows per instruction (32 LSU
// wavefronts per store; measured: the lane-per-row form made the stem and the layer1 / layer2 kernels store-bound,
// profiles/r02f).  The residual takes the same route in the other direction.  Ragged n-tiles / unaligned outputs and
// launches without room for the buffer use the lane-per-row form (tm_epilogue16).
// Residual prefetch for the staged epilogue: the coalesced residual chunks of a tile (CPR x 16 bytes per lane), to be
// issued one tile ahead so that their HBM latency never sits on the per-tile critical path (measured: fetching them at
// the start of the tile's own epilogue cost ~2k clocks per tile, profiles/r02g).
template <int EN, bool TF32>
struct TmResidual {
  static constexpr int O_ES = TF32 ? 4 : 2;
  static constexpr int CPR = EN * O_ES / 16;
  static constexpr int RPI = 32 / CPR;
  static constexpr bool PREFETCH = CPR <= 8;      // (16 chunks per lane would cost 64 registers)
  uint4 v[PREFETCH ? CPR : 1];
  bool have;
  __device__ __forceinline__ void fetch(const FusedParams& p, int g, int n0, int ncol0, long long orow, bool mvalid, uint32_t stg,
                                        int lane) {
    have = false;
    if constexpr (PREFETCH) {
      const bool tile_vec = p.out_vec && n0 + ncol0 + EN <= p.N;
      if (stg == 0u || !tile_vec || p.ep_residual == nullptr || p.probe >= 2) return;
      const uint8_t* resb = static_cast<const uint8_t*>(p.ep_residual);
      const int crow = lane / CPR, cch = lane % CPR;
      const long long orow_v = mvalid ? orow : -1ll;
      const long long col_b = ((long long)g * p.N + n0 + ncol0) * O_ES + cch * 16;
#pragma unroll
      for (int i = 0; i < CPR; ++i) {
        // UNCONDITIONAL load from a clamped row: a select between the loaded value and zero right behind the load made
        // every lane wait for the data here (profiles/r02j: 20% of this kernel's stall samples on that one move) and
        // turned the prefetch into a blocking read.  Rows outside the sample are never stored, so their value is free.
        const long long ro = __shfl_sync(0xffffffffu, orow_v, i * RPI + crow);
        v[i] = ldg16(resb + (ro >= 0 ? ro : 0ll) * p.C_out * O_ES + col_b);
      }
      have = true;
    }
  }
  // Branch-free form for the prefetch loops: NO control flow between the loads and the slot's registers (under an `if`
  // ptxas loads into temporaries and merges them into the slot with moves that wait for the data: profiles/r02q, r02r).
  // The caller has checked once that the staged residual path applies (ep_residual, staging buffer, full n-tile).
  __device__ __forceinline__ void fetch_always(const FusedParams& p, int g, int n0, int ncol0, long long orow, bool mvalid, int lane) {
    static_assert(PREFETCH, "fetch_always needs CPR <= 8");
    const uint8_t* resb = static_cast<const uint8_t*>(p.ep_residual);
    const int crow = lane / CPR, cch = lane % CPR;
    const long long orow_v = mvalid ? orow : 0ll;          // rows outside the sample: any valid row (never stored)
    const long long col_b = ((long long)g * p.N + n0 + ncol0) * O_ES + cch * 16;
#pragma unroll
    for (int i = 0; i < CPR; ++i) {
      const long long ro = __shfl_sync(0xffffffffu, orow_v, i * RPI + crow);
      v[i] = ldg16(resb + ro * p.C_out * O_ES + col_b);
    }
    have = true;
  }
};

template <int EN, bool TF32, bool FLIP = false>
__device__ __forceinline__ void tm_epilogue_tile(const FusedParams& p, const float* bias_s, uint32_t taddr, int g, int n0,
                                                 int ncol0, long long orow, bool mvalid, uint32_t stg, int lane,
                                                 uint32_t flip_off = 0u, uint4 sblk = make_uint4(0u, 0u, 0u, 0u),
                                                 const TmResidual<EN, TF32>* pre = nullptr) {
  constexpr int O_ES = TF32 ? 4 : 2;
  constexpr int ROWB = EN * O_ES;              // staged bytes per row
  constexpr int CPR = ROWB / 16;               // 16-byte chunks per row = lanes per row in the coalesced passes
  constexpr int RPI = 32 / CPR;                // rows per coalesced instruction
  constexpr int CP16 = 16 * O_ES / 16;         // chunks per 16 columns: 2 (bf16) | 4 (fp32)
  const bool tile_vec = p.out_vec && n0 + ncol0 + EN <= p.N;
  if (stg == 0u || !tile_vec) {
#pragma unroll 1
    for (int cb = 0; cb < EN; cb += 16)
      tm_epilogue16<TF32, FLIP>(p, bias_s, taddr + cb, g, n0, ncol0 + cb, orow, mvalid, flip_off, sblk);
    return;
  }
  auto swz = [](int c, int r) -> int {
    return CPR >= 8 ? (c ^ (r & (CPR - 1))) : (CPR == 4 ? (c ^ ((r >> 1) & 3)) : (CPR == 2 ? (c ^ ((r >> 2) & 1)) : c));
  };
  uint8_t* outb = static_cast<uint8_t*>(p.out);
  const uint8_t* resb = p.probe == 3 ? nullptr : static_cast<const uint8_t*>(p.ep_residual);
  const int crow = lane / CPR, cch = lane % CPR;
  const long long orow_v = mvalid ? orow : -1ll;
  const long long col_b = ((long long)g * p.N + n0 + ncol0) * O_ES + cch * 16;
  if (resb != nullptr) {                       // residual -> staging, coalesced (prefetched one tile ahead when possible)
    const bool pf = pre != nullptr && pre->have;
#pragma unroll
    for (int i = 0; i < CPR; ++i) {
      const int r = i * RPI + crow;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if constexpr (TmResidual<EN, TF32>::PREFETCH) {
        if (pf) v = pre->v[i];
      }
      if (!pf) {
        const long long ro = __shfl_sync(0xffffffffu, orow_v, r);
        if (ro >= 0) v = ldg16(resb + ro * p.C_out * O_ES + col_b);
      }
      sts16(stg + (uint32_t)(r * ROWB + (swz(cch, r) << 4)), v);
    }
    __syncwarp();
  }
  const bool has_affine = p.ep_scale != nullptr;
  if constexpr (!FLIP) {
    // two 16-column groups per TMEM round trip, packed arithmetic (tm_out16)
    const uint32_t cst = smem_u32(bias_s) + (uint32_t)(ncol0 * 4);
    const uint32_t floor_bits = tm_floor_bits<TF32>(p.ep_relu != 0);
    constexpr int GRP = EN >= 32 ? 2 : 1;
#pragma unroll
    for (int cb = 0; cb < EN; cb += 16 * GRP) {
      uint32_t v[GRP][16];
#pragma unroll
      for (int h = 0; h < GRP; ++h) tmem_ld16(taddr + cb + 16 * h, v[h]);
      tmem_ld_wait();
#pragma unroll
      for (int h = 0; h < GRP; ++h) {
        const int c0 = (cb + 16 * h) * O_ES / 16;
        uint32_t sa[CP16];
#pragma unroll
        for (int k = 0; k < CP16; ++k) sa[k] = stg + (uint32_t)(lane * ROWB + (swz(c0 + k, lane) << 4));
        uint4 rr[CP16], ch[CP16];
        if (resb != nullptr) {
#pragma unroll
          for (int k = 0; k < CP16; ++k) rr[k] = lds16(sa[k]);
        }
        tm_out16<TF32>(v[h], cst + (uint32_t)((cb + 16 * h) * 4), resb != nullptr ? rr : nullptr, floor_bits, ch);
#pragma unroll
        for (int k = 0; k < CP16; ++k) sts16(sa[k], ch[k]);
      }
    }
  }
#pragma unroll 1
  for (int cb = 0; FLIP && cb < EN; cb += 16) {
    uint32_t v0[16], v1[16];
    tmem_ld16(taddr + cb, v0);
    if constexpr (FLIP) tmem_ld16(taddr + flip_off + cb, v1);
    tmem_ld_wait();
    float o[16];
    if constexpr (FLIP) {
      tm_flip_combine16(bias_s, ncol0 + cb, v0, v1, sblk, (n0 & 127) + ncol0 + cb, has_affine, o);
    } else {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 sh = *reinterpret_cast<const float4*>(bias_s + 256 + ncol0 + cb + 4 * jj);
        float v[4] = {__uint_as_float(v0[4 * jj]), __uint_as_float(v0[4 * jj + 1]), __uint_as_float(v0[4 * jj + 2]),
                      __uint_as_float(v0[4 * jj + 3])};
        if (has_affine) {
          const float4 sc = *reinterpret_cast<const float4*>(bias_s + 128 + ncol0 + cb + 4 * jj);
          v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y);
          v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w);
        } else {
          v[0] += sh.x; v[1] += sh.y; v[2] += sh.z; v[3] += sh.w;
        }
        o[4 * jj] = v[0]; o[4 * jj + 1] = v[1]; o[4 * jj + 2] = v[2]; o[4 * jj + 3] = v[3];
      }
    }
    const int c0 = cb * O_ES / 16;             // first chunk of these 16 columns inside the staged row
    uint32_t sa[CP16];
#pragma unroll
    for (int k = 0; k < CP16; ++k) sa[k] = stg + (uint32_t)(lane * ROWB + (swz(c0 + k, lane) << 4));
    if (resb != nullptr) {
#pragma unroll
      for (int k = 0; k < CP16; ++k) {
        uint4 a;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sa[k]));
        if constexpr (TF32) {
          o[4 * k] += __uint_as_float(a.x); o[4 * k + 1] += __uint_as_float(a.y);
          o[4 * k + 2] += __uint_as_float(a.z); o[4 * k + 3] += __uint_as_float(a.w);
        } else {
          const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[8 * k + 2 * j] += bt_bf16_lo(w[j]);
            o[8 * k + 2 * j + 1] += bt_bf16_hi(w[j]);
          }
        }
      }
    }
    if (p.ep_relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    if constexpr (TF32) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sts16(sa[k], make_uint4(__float_as_uint(o[4 * k]), __float_as_uint(o[4 * k + 1]), __float_as_uint(o[4 * k + 2]),
                                __float_as_uint(o[4 * k + 3])));
    } else {
      sts16(sa[0], make_uint4(bt_pack_bf16x2(o[0], o[1]), bt_pack_bf16x2(o[2], o[3]), bt_pack_bf16x2(o[4], o[5]), bt_pack_bf16x2(o[6], o[7])));
      sts16(sa[1], make_uint4(bt_pack_bf16x2(o[8], o[9]), bt_pack_bf16x2(o[10], o[11]), bt_pack_bf16x2(o[12], o[13]),
                              bt_pack_bf16x2(o[14], o[15])));
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < CPR; ++i) {              // copy-out, coalesced: CPR lanes per row
    const int r = i * RPI + crow;
    const long long ro = __shfl_sync(0xffffffffu, orow_v, r);
    if (ro >= 0) {
      uint4 v;
      const uint32_t a = stg + (uint32_t)(r * ROWB + (swz(cch, r) << 4));
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
      *reinterpret_cast<uint4*>(outb + ro * p.C_out * O_ES + col_b) = v;
    }
  }
  __syncwarp();                                // the staging buffer is rewritten by the next tile
}

// Epilogue + 3x3 / stride-2 / pad-1 max-pool of one 128-row tile of one MC sample (torchvision's ResNet stem:
// conv -> bn -> relu -> maxpool; resnet.py's `self.maxpool`), run by the 256 threads of warps 0-7 together.  The tile
// holds R = 128 / OW whole output rows of ONE image (host guarantees OW | 128, R even, OH * OW a multiple of 128), the
// CTA walks the T = OH * OW / 128 tiles of an image in order, and
//   1. every warp turns its 32 rows x EN accumulator columns into outputs (affine, ReLU, rounding to the output dtype)
//      and parks them in the CTA-wide tile buffer tile_s [128][BLOCK_N] (16-byte chunks XOR-swizzled by row & 7); the
//      last conv row of the tile is ALSO written to carry_s[(t + 1) & 1] for the next tile of the image (pooled row
//      2r - 1 of the next tile);
//   2. named barrier; each thread reduces (pooled pixel, 16-byte channel chunk) items over their <= 9 taps -- rows
//      2pr-1 (the carry of tile t-1, absent for t = 0: padding), 2pr, 2pr+1; columns 2pw-1 (absent for pw = 0), 2pw,
//      2pw+1 -- and stores the pooled vector: 1/4 of the unpooled bytes reach HBM and the separate pooling kernel (a
//      full read + write of the [S*B, OH, OW, C] activation) disappears;
//   3. named barrier (the tile buffer is rewritten by the next tile-sample).
// max of values already rounded to the output dtype == rounding of the max (rounding is monotonic), so the result is
// bit-identical to epilogue -> store -> bt_maxpool2d_nhwc.
// per-thread constants of the pooling pass, computed once per CTA: the thread's ITEMS (pooled pixel, 16-byte channel
// chunk) items -- tap offsets inside the tile buffer / the carry rows and the output offset inside the tile's pooled rows
template <int BLOCK_N, bool TF32>
struct TmPoolPlan {
  static constexpr int O_ES = TF32 ? 4 : 2;
  static constexpr int ROWB = BLOCK_N * O_ES;
  static constexpr int CPR = ROWB / 16;
  static constexpr int ITEMS = (32 * CPR + 255) / 256;      // 1 (bf16, 64 columns) | 2 | 4
  uint32_t off[ITEMS][9];      // [dh + 1][dw + 1]: byte offset of the tap's chunk (tile buffer; carry rows for top_carry)
  uint32_t out_off[ITEMS];     // byte offset of the pooled vector inside this tile's pooled rows of `out`
  bool top_carry[ITEMS], has_left[ITEMS];
  __device__ __forceinline__ void init(const FusedParams& p, int g, int n0, int etid) {
    const int OW = p.pool_ow, PWd = OW >> 1;
    const int ckey = (BLOCK_M - OW) & 7;                     // swizzle key of carry pixel ow: (BLOCK_M - OW + ow) & 7
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      const int it = etid + 256 * q;
      const int pp = it / CPR, c = it % CPR;
      const int pr = pp / PWd, pw = pp - pr * PWd;
      top_carry[q] = pr == 0;
      has_left[q] = pw != 0;
#pragma unroll
      for (int dh = -1; dh <= 1; ++dh) {
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
          const int lr = 2 * pr + dh, ow = 2 * pw + dw;
          uint32_t o;
          if (lr < 0) o = (uint32_t)(ow * ROWB + ((c ^ ((ckey + ow) & 7)) << 4));
          else {
            const int ri = lr * OW + ow;
            o = (uint32_t)(ri * ROWB + ((c ^ (ri & 7)) << 4));
          }
          off[q][(dh + 1) * 3 + dw + 1] = o;
        }
      }
      out_off[q] = (uint32_t)((((pr * PWd + pw) * p.C_out) + g * p.N + n0) * O_ES + c * 16);
    }
  }
};

// Epilogue + 3x3 / stride-2 / pad-1 max-pool of one 128-row tile of one MC sample (torchvision's ResNet stem:
// conv -> bn -> relu -> maxpool; resnet.py's `self.maxpool`), run by the 256 threads of warps 0-7 together.  The tile
// holds R = 128 / OW whole output rows of ONE image (host guarantees OW | 128, R even, OH * OW a multiple of 128), the
// CTA walks the T = OH * OW / 128 tiles of an image in order, and
//   1. every warp turns its 32 rows x EN accumulator columns into outputs (tm_out16: affine, ReLU, rounding to the
//      output dtype) and parks them in the CTA-wide tile buffer tile_s [128][BLOCK_N] (16-byte chunks XOR-swizzled by
//      row & 7); the last conv row of the tile is ALSO written to carry_s[(t + 1) & 1] for the next tile of the image
//      (pooled row 2r - 1 of the next tile);
//   2. named barrier; each thread reduces its (pooled pixel, 16-byte channel chunk) items over their <= 9 taps -- rows
//      2pr-1 (the carry of tile t-1, absent for t = 0: padding), 2pr, 2pr+1; columns 2pw-1 (absent for pw = 0), 2pw,
//      2pw+1 -- and stores the pooled vector: 1/4 of the unpooled bytes reach HBM and the separate pooling kernel (a
//      full read + write of the [S*B, OH, OW, C] activation) disappears;
//   3. named barrier (the tile buffer is rewritten by the next tile-sample).
// max of values already rounded to the output dtype == rounding of the max (rounding is monotonic), so the result is
// bit-identical to epilogue -> store -> bt_maxpool2d_nhwc.
template <int BLOCK_N, bool TF32>
__device__ __forceinline__ void tm_epilogue_pool(const FusedParams& p, const float* bias_s, uint32_t taddr, int ncol0, int q4,
                                                 uint32_t tile_s, uint32_t carry_s, int t, int T, long long gimg, int lane,
                                                 const TmPoolPlan<BLOCK_N, TF32>& pl) {
  constexpr int O_ES = TF32 ? 4 : 2;
  constexpr int EN = BLOCK_N / 2;
  constexpr int ROWB = BLOCK_N * O_ES;         // bytes per pixel row of the tile buffer
  constexpr int CP16 = O_ES;                   // chunks per 16 columns: 2 (bf16) | 4 (fp32)
  constexpr int ITEMS = TmPoolPlan<BLOCK_N, TF32>::ITEMS;
  const int OW = p.pool_ow, R = BLOCK_M / OW;
  const int row = q4 * 32 + lane;
  const bool to_carry = t + 1 < T && row >= BLOCK_M - OW;
  const uint32_t carry_w = carry_s + (uint32_t)(((t + 1) & 1) * OW * ROWB + (row - (BLOCK_M - OW)) * ROWB);
  const uint32_t cst = smem_u32(bias_s) + (uint32_t)(ncol0 * 4);
  const uint32_t floor_bits = tm_floor_bits<TF32>(p.ep_relu != 0);
  constexpr int GRP = EN >= 32 ? 2 : 1;
#pragma unroll
  for (int cb = 0; cb < EN; cb += 16 * GRP) {
    uint32_t v[GRP][16];
#pragma unroll
    for (int h = 0; h < GRP; ++h) tmem_ld16(taddr + cb + 16 * h, v[h]);
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < GRP; ++h) {
      uint4 ch[CP16];
      tm_out16<TF32>(v[h], cst + (uint32_t)((cb + 16 * h) * 4), nullptr, floor_bits, ch);
      const int c0 = (ncol0 + cb + 16 * h) * O_ES / 16;   // first chunk of these 16 columns inside the pixel row
#pragma unroll
      for (int k = 0; k < CP16; ++k) {
        const uint32_t sw = (uint32_t)(((c0 + k) ^ (row & 7)) << 4);
        sts16(tile_s + (uint32_t)(row * ROWB) + sw, ch[k]);
        if (to_carry) sts16(carry_w + sw, ch[k]);           // (same chunk permutation as the tile row: key row & 7)
      }
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  {
    const int PWd = OW >> 1, PH = p.pool_oh >> 1;
    const uint32_t carry_r = carry_s + (uint32_t)((t & 1) * OW * ROWB);
    uint8_t* outb = static_cast<uint8_t*>(p.out) + ((gimg * PH + (long long)t * (R >> 1)) * PWd) * p.C_out * O_ES;
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      uint4 m;
      bool first = true;
      auto take = [&](uint32_t a) {
        const uint4 v = lds16(a);
        if (first) {
          m = v;
          first = false;
        } else if constexpr (TF32) {
          m.x = __float_as_uint(fmaxf(__uint_as_float(m.x), __uint_as_float(v.x)));
          m.y = __float_as_uint(fmaxf(__uint_as_float(m.y), __uint_as_float(v.y)));
          m.z = __float_as_uint(fmaxf(__uint_as_float(m.z), __uint_as_float(v.z)));
          m.w = __float_as_uint(fmaxf(__uint_as_float(m.w), __uint_as_float(v.w)));
        } else {
          asm("max.bf16x2 %0, %0, %1;" : "+r"(m.x) : "r"(v.x));
          asm("max.bf16x2 %0, %0, %1;" : "+r"(m.y) : "r"(v.y));
          asm("max.bf16x2 %0, %0, %1;" : "+r"(m.z) : "r"(v.z));
          asm("max.bf16x2 %0, %0, %1;" : "+r"(m.w) : "r"(v.w));
        }
      };
      // the centre tap always exists: start from it, then the (predicated) rest
      take(tile_s + pl.off[q][4]);
      take(tile_s + pl.off[q][5]);
      take(tile_s + pl.off[q][7]);
      take(tile_s + pl.off[q][8]);
      if (pl.has_left[q]) {
        take(tile_s + pl.off[q][3]);
        take(tile_s + pl.off[q][6]);
      }
      if (!(pl.top_carry[q] && t == 0)) {                    // row 2pr - 1: the tile itself, or the carry of tile t - 1
        const uint32_t tb = pl.top_carry[q] ? carry_r : tile_s;
        take(tb + pl.off[q][1]);
        take(tb + pl.off[q][2]);
        if (pl.has_left[q]) take(tb + pl.off[q][0]);
      }
      *reinterpret_cast<uint4*>(outb + pl.out_off[q]) = m;
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

// first output pixel (b, od, oh, ow) of row m0 (warp-uniform; once per tile)
__device__ __forceinline__ void tm_decode_row(const FusedParams& p, long long m0, int& b, int& od, int& oh, int& ow) {
  const long long out_sp = (long long)p.OD * p.OH * p.OW;
  b = (int)(m0 / out_sp);
  long long rem = m0 - (long long)b * out_sp;
  od = (int)(rem / ((long long)p.OH * p.OW));
  rem -= (long long)od * p.OH * p.OW;
  oh = (int)(rem / p.OW);
  ow = (int)(rem - (long long)oh * p.OW);
}

// tf32: round one TMA-staged activation tile (128 rows x 128 B of fp32 words) to tf32, in place, with the 128 threads
// of the converter warps.  tcgen05.mma kind::tf32 ignores the low 13 mantissa bits of its operands (truncation, a
// systematic -3.4e-4 relative shrink that compounds over a deep network); rounding to nearest first makes the product
// the same as on the generic path (both operands cvt.rna.tf32.f32) -- rel-RMS 2.9e-4 instead of 4.6e-4 per layer.
__device__ __forceinline__ void tm_round_tile_tf32(uint32_t tile, int ctid) {
#pragma unroll
  for (int i = 0; i < A_TILE_BYTES / 16 / (TM_CONV_WARPS * 32); ++i) {
    const uint32_t a = tile + (uint32_t)((i * TM_CONV_WARPS * 32 + ctid) * 16);
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    sts16(a, make_uint4(bt_tf32(__uint_as_float(v.x)), bt_tf32(__uint_as_float(v.y)), bt_tf32(__uint_as_float(v.z)),
                        bt_tf32(__uint_as_float(v.w))));
  }
}

// ------------------------------------------------------------------ kernel 1: weight-stationary (W_s resident)
template <int BLOCK_N, bool P_BF16, bool TF32>
__global__ void __launch_bounds__(tm_threads<TF32>(), 1) bt_tma_kernel(const __grid_constant__ TmaParams tp) {
  const FusedParams& p = tp.f;
  constexpr int B_TILE_BYTES = BLOCK_N * 128;
  constexpr int KBE = TF32 ? 32 : 64;               // k per k-block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int NSMP = tp.a.nsmp;                        // MC samples of this CTA (shared x only)
  const int res_bytes = NSMP * p.num_kb * B_TILE_BYTES;
  const int NSTG = p.stages;
  uint8_t* aux = smem + res_bytes + NSTG * A_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  float* bias_all = reinterpret_cast<float*>(aux + 512);         // [NSMP][3][128]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 5);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_base = smem_base + res_bytes;
  const uint32_t full_bar0 = smem_u32(bars);                        // stage ready for the tensor core
  const uint32_t empty_bar0 = smem_u32(bars + MAX_STAGES);
  const uint32_t bready_bar = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t acc_bar0 = smem_u32(bars + 2 * MAX_STAGES + 1);    // [2]
  const uint32_t tfree_bar0 = smem_u32(bars + 2 * MAX_STAGES + 3);  // [2]
  const uint32_t afull_bar0 = smem_u32(bars + 2 * MAX_STAGES + 5);  // tf32: TMA bytes landed (converter warps wait here)
  const uint32_t land_bar0 = TF32 ? afull_bar0 : full_bar0;         // where the TMA transaction completes

  const int s = blockIdx.z * NSMP;                   // first MC sample of this CTA
  const int ns_live = p.S - s < NSMP ? p.S - s : NSMP;
  const int g = blockIdx.y / p.n_tiles_per_group;
  const int n0 = (blockIdx.y % p.n_tiles_per_group) * BLOCK_N;
  const uint32_t sample = p.sample0 + (uint32_t)s + (p.sample_ptr != nullptr ? __ldg(p.sample_ptr) : 0u);
  const int img_base = p.x_shared ? 0 : s * p.B;
  const long long n_rt = p.n_groups;   // 128-row tiles per sample; this CTA takes blockIdx.x, +gridDim.x, ...
  const int slabs = tp.a.slabs;
  // Fused max-pool: a CTA takes whole images (T consecutive tiles each) so that the pooling windows that straddle two
  // tiles find the previous tile's last row in this CTA's shared memory.  T = 1 otherwise.
  const int T = p.pool_ow ? (p.pool_oh * p.pool_ow) / BLOCK_M : 1;
  auto tile_of = [&](long long it, int& t) -> long long {   // it-th tile of this CTA (>= n_rt: done)
    if (T == 1) {
      t = 0;
      return blockIdx.x + it * gridDim.x;
    }
    const long long q = it / T;
    t = (int)(it - q * T);
    return (blockIdx.x + q * gridDim.x) * T + t;
  };

  if (warp == TM_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < NSTG; ++i) {
        // bf16: one arrive.expect_tx by the TMA warp + the transaction bytes; tf32: the converter warps
        mbar_init(full_bar0 + 8 * i, TF32 ? TM_CONV_WARPS : 1);
        mbar_init(empty_bar0 + 8 * i, 1);
        mbar_init(afull_bar0 + 8 * i, 1);
      }
      mbar_init(bready_bar, TM_SAMP_WARPS);
      for (int i = 0; i < 2; ++i) {
        mbar_init(acc_bar0 + 8 * i, 1);
        mbar_init(tfree_bar0 + 8 * i, TM_SAMP_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  } else if (warp == TM_TMA_WARP) {
    if (lane == 0) tma_prefetch_desc(&tp.map_a);
  } else if (tid < BLOCK_N) {
    for (int j = 0; j < ns_live; ++j) tm_fill_bias<P_BF16>(p, bias_all + j * 384, tid, g, n0, sample + (uint32_t)j);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TM_MMA_WARP) {
    // ============================================================== MMA issuer (whole warp, elected issue)
    const uint32_t idesc = make_idesc(BLOCK_N, TF32);
    const uint64_t desc_hi = make_smem_desc(0u);
    int stage = 0;
    uint32_t phase = 0;
    int t_ = 0;
    if (bt_elect_one()) {                             // one thread issues every MMA (umma1_x4)
    mbar_wait_idle(bready_bar, 0, 256);
    tc_fence_after();
    for (long long it = 0; tile_of(it, t_) < n_rt; ++it) {
      const int buf = (int)(it & 1);
      if (it >= 2) {  // the epilogue has drained this accumulator buffer
        mbar_wait_idle(tfree_bar0 + 8 * buf, (uint32_t)(((it >> 1) - 1) & 1), 64);
        tc_fence_after();
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait_idle(full_bar0 + 8 * stage, phase, 32);
        tc_fence_after();
        const uint32_t sa16 = ((ring_base + stage * A_TILE_BYTES) & 0x3FFFFu) >> 4;
        for (int j = 0; j < ns_live; ++j) {           // the staged activation tile x the resident tile of every sample
          const uint32_t sb16 = ((smem_base + (j * p.num_kb + kb) * B_TILE_BYTES) & 0x3FFFFu) >> 4;
          umma1_x4<TF32>(tmem_base + (uint32_t)((buf * NSMP + j) * BLOCK_N), sa16 | (1u << 16), sb16 | (1u << 16),
                         (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
        }
        umma1_commit(empty_bar0 + 8 * stage);
        if (++stage == NSTG) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma1_commit(acc_bar0 + 8 * buf);
    }
    }
    __syncwarp();
  } else if (warp == TM_TMA_WARP) {
    // ============================================================== TMA producer (whole warp, elected issue)
    int stage = 0;
    uint32_t phase = 0;
    int t_ = 0;
    long long rt;
    for (long long it = 0; (rt = tile_of(it, t_)) < n_rt; ++it) {
      const long long m0 = rt * BLOCK_M;
      int b = 0, od = 0, oh = 0, ow = 0;
      if (tp.a.mode == 2) tm_decode_row(p, m0, b, od, oh, ow);
      int tap_i = 0, slab = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
        mbar_expect_tx_elect(land_bar0 + 8 * stage, A_TILE_BYTES);
        tma_issue_a(tp, ring_base + stage * A_TILE_BYTES, land_bar0 + 8 * stage, img_base, g, m0, tap_i, slab, b, od, oh, ow);
        if (++slab == slabs) {
          slab = 0;
          ++tap_i;
        }
        if (++stage == NSTG) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (TF32 && warp >= TM_CONV_WARP0 && warp < TM_CONV_WARP0 + TM_CONV_WARPS) {
    // ============================================================== tf32: round the staged activations to nearest
    const int ctid = tid - TM_CONV_WARP0 * 32;
    int stage = 0;
    uint32_t phase = 0;
    int t_ = 0;
    for (long long it = 0; tile_of(it, t_) < n_rt; ++it) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(afull_bar0 + 8 * stage, phase);
        tm_round_tile_tf32(ring_base + stage * A_TILE_BYTES, ctid);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar0 + 8 * stage);
        if (++stage == NSTG) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp < TM_SAMP_WARPS) {
    // ============================================================== warps 0-7: sample W_s, then epilogue
    {
      TmSampler<BLOCK_N, P_BF16, TF32> smp;
      smp.init(p, tid, g, n0);
      int tap_i = 0, slab = 0;
      // physical k of this thread's oct in the NEXT k-block: (tap, channel) -> tap.lin * Cin_g + channel (tiled mode: k itself)
      auto prefetch_next = [&](auto ph) {
        const int kc = slab * KBE + smp.koff();
        const bool kvalid = tp.a.mode == 1 ? (kc < p.K_used) : true;
        const long long kphys0 = tp.a.mode == 1 ? (long long)kc : (long long)decode_tap(p, tap_i).lin * p.Cin_g + kc;
        if (++slab == slabs) {
          slab = 0;
          ++tap_i;
        }
        smp.template prefetch<decltype(ph)::value>(p, kphys0, kvalid);
      };
      tm_sample_loop(smp, p.num_kb, prefetch_next, [&](auto ph, int kb) {
        for (int j = 0; j < ns_live; ++j)               // (shared x: the same parameter words serve every sample of the CTA)
          smp.template compute<decltype(ph)::value>(p, sample + (uint32_t)j, smem_base + (j * p.num_kb + kb) * B_TILE_BYTES);
      });
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bready_bar);
    }
    // ---- epilogue: lane quarter q4 = warp & 3, column half = warp >> 2 (EN columns each)
    constexpr int EN = BLOCK_N / 2;
    const int q4 = warp & 3, ncol0 = (warp >> 2) * EN;
    const uint32_t stg = p.dr_stage ? smem_u32(aux + TM_AUX_BYTES) + (uint32_t)(warp * 32 * EN * (TF32 ? 4 : 2)) : 0u;
    const uint32_t pool_tile = smem_u32(aux + TM_AUX_BYTES);                     // fused max-pool: [128][BLOCK_N] outputs
    const uint32_t pool_carry = pool_tile + (uint32_t)(BLOCK_M * BLOCK_N * (TF32 ? 4 : 2));   // [NSMP][2][OW][BLOCK_N]
    const int carry_b = 2 * p.pool_ow * BLOCK_N * (TF32 ? 4 : 2);
    int t = 0;
    long long rt;
    constexpr bool POOL_OK = BLOCK_N * (TF32 ? 4 : 2) >= 128;
    TmPoolPlan<POOL_OK ? BLOCK_N : 64, POOL_OK ? TF32 : true> pool_plan;
    if (POOL_OK && p.pool_ow) pool_plan.init(p, g, n0, tid);
    for (long long it = 0; (rt = tile_of(it, t)) < n_rt; ++it) {
      const int buf = (int)(it & 1);
      const long long m = rt * BLOCK_M + q4 * 32 + lane;
      mbar_wait_idle(acc_bar0 + 8 * buf, (uint32_t)((it >> 1) & 1), 128);
      tc_fence_after();
      if (p.pool_ow) {
        if constexpr (POOL_OK) {
          for (int j = 0; j < ns_live; ++j)
            tm_epilogue_pool<BLOCK_N, TF32>(
                p, bias_all + j * 384,
                tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((buf * NSMP + j) * BLOCK_N + ncol0), ncol0, q4,
                pool_tile, pool_carry + (uint32_t)(j * carry_b), t, T, (long long)(s + j) * (n_rt / T) + rt / T, lane,
                pool_plan);
        }
      } else
      for (int j = 0; j < ns_live; ++j)
        tm_epilogue_tile<EN, TF32>(p, bias_all + j * 384,
                                   tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((buf * NSMP + j) * BLOCK_N + ncol0), g, n0,
                                   ncol0, (long long)(s + j) * p.M + m, m < p.M, stg, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tfree_bar0 + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------ kernel 2: streaming (W_s tile per k-block)
// For layers whose sampled tile [BLOCK_N x K] does not fit shared memory, or whose M is a few tiles only: CTA =
// (group of MT 128-row tiles, n-tile, MC sample).  Per k-block the TMA warp loads the MT activation tiles while warps 0-7
// sample the [BLOCK_N x KBE] weight tile into the same stage -- every sampled tile is used by MT x 128 output rows --
// and the MMA warp issues MT x 4 MMAs into MT accumulators.  Warps 0-7 run the epilogue after the last k-block.
// FLIP (Flipout, linear_flipout.py:145-197 / conv_flipout.py:370-439): the samplers write the pair (mu tile, sigma*eps
// tile), the transform warps (10-13) build the second activation plane x * s_in from the TMA-staged tile -- one Philox
// call per row and k-block gives its input sign bits -- two accumulators per row tile, and the epilogue combines
// (acc0 + mu_b) + s_out * (acc1 + sigma_b eps_b) with on-chip output signs.
template <int BLOCK_N, bool P_BF16, bool TF32, bool FLIP>
__global__ void __launch_bounds__(tm_threads<TF32 || FLIP>(), 1) bt_tms_kernel(const __grid_constant__ TmaParams tp) {
  const FusedParams& p = tp.f;
  constexpr int NB = FLIP ? 2 : 1;
  constexpr bool XFORM = TF32 || FLIP;               // warps 10-13 post-process every landed activation tile
  constexpr int B_TILE_BYTES = BLOCK_N * 128;
  constexpr int KBE = TF32 ? 32 : 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MT = p.MT;
  const int NSTG = p.stages;
  // stage: [NB weight tiles][MT x NB activation tiles: (plane 0 = x, plane 1 = x * s_in) per row tile]
  const int stage_bytes = NB * (B_TILE_BYTES + MT * A_TILE_BYTES);
  constexpr int A_OFF = NB * B_TILE_BYTES;
  uint8_t* aux = smem + NSTG * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  float* bias_s = reinterpret_cast<float*>(aux + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 5);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full_bar0 = smem_u32(bars);
  const uint32_t empty_bar0 = smem_u32(bars + MAX_STAGES);
  const uint32_t acc_bar = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t afull_bar0 = smem_u32(bars + 2 * MAX_STAGES + 5);  // tf32: TMA bytes landed (converter warps wait here)
  const uint32_t land_bar0 = XFORM ? afull_bar0 : full_bar0;

  const int s = blockIdx.z;
  const int g = blockIdx.y / p.n_tiles_per_group;
  const int n0 = (blockIdx.y % p.n_tiles_per_group) * BLOCK_N;
  const uint32_t sample = p.sample0 + (uint32_t)s + (p.sample_ptr != nullptr ? __ldg(p.sample_ptr) : 0u);
  const int img_base = p.x_shared ? 0 : s * p.B;
  const long long m_base = (long long)blockIdx.x * MT * BLOCK_M;   // first row of this CTA's M-group
  const int slabs = tp.a.slabs;
  // A-operand multicast: the CLN CTAs of a cluster (consecutive n-tiles of one group, same M-group and sample) need the
  // same activation tiles; CTA `crank` loads the row tiles mt = crank (mod CLN) and multicasts them to all.  A stage may be
  // refilled only when EVERY CTA of the cluster has consumed it: the MMA commits are multicast to all empty barriers.
  const int CLN = tp.a.cln > 1 ? tp.a.cln : 1;
  const int crank = CLN > 1 ? (int)bt_cluster_ctarank() : 0;
  const uint16_t cmask = CLN > 1 ? (uint16_t)((1u << CLN) - 1u) : (uint16_t)0;

  if (warp == TM_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < NSTG; ++i) {
        // 8 sampler warps + (bf16) the TMA warp's arrive.expect_tx / (tf32) the converter warps
        mbar_init(full_bar0 + 8 * i, TM_SAMP_WARPS + (XFORM ? TM_CONV_WARPS : 1));
        mbar_init(empty_bar0 + 8 * i, CLN);              // the MMA commits of every CTA that reads this stage's A tiles
        mbar_init(afull_bar0 + 8 * i, 1);
      }
      mbar_init(acc_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  } else if (warp == TM_TMA_WARP) {
    if (lane == 0) tma_prefetch_desc(&tp.map_a);
  } else if (tid < BLOCK_N) {
    tm_fill_bias<P_BF16, FLIP>(p, bias_s, tid, g, n0, sample);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (CLN > 1) bt_cluster_sync();                        // every CTA's barriers exist before a peer's TMA completes on them

  if (warp == TM_MMA_WARP) {
    const uint32_t idesc = make_idesc(BLOCK_N, TF32);
    const uint64_t desc_hi = make_smem_desc(0u);
    int stage = 0;
    uint32_t phase = 0;
    const long long left = (p.M - m_base + BLOCK_M - 1) / BLOCK_M;
    const int mt_live = left < MT ? (int)left : MT;     // tiles of the group that start inside the sample
    if (bt_elect_one()) {                               // one thread issues every MMA (umma1_x4)
    for (int kb = 0; kb < p.num_kb; ++kb) {
      mbar_wait_idle(full_bar0 + 8 * stage, phase, 32);
      tc_fence_after();
      const uint32_t sst = smem_base + stage * stage_bytes;
      const uint32_t sb16 = (sst & 0x3FFFFu) >> 4;
      for (int mt = 0; mt < mt_live; ++mt) {
        const uint32_t sa16 = ((sst + A_OFF + mt * NB * A_TILE_BYTES) & 0x3FFFFu) >> 4;
        umma1_x4<TF32>(tmem_base + (uint32_t)(mt * NB * BLOCK_N), sa16 | (1u << 16), sb16 | (1u << 16),
                       (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
        if constexpr (FLIP)
          umma1_x4<TF32>(tmem_base + (uint32_t)((mt * NB + 1) * BLOCK_N), (sa16 + (A_TILE_BYTES >> 4)) | (1u << 16),
                         (sb16 + (B_TILE_BYTES >> 4)) | (1u << 16), (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
      }
      if (CLN > 1) umma1_commit_mc(empty_bar0 + 8 * stage, cmask);
      else umma1_commit(empty_bar0 + 8 * stage);
      if (++stage == NSTG) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma1_commit(acc_bar);
    }
    __syncwarp();
  } else if (warp == TM_TMA_WARP) {
    int stage = 0;
    uint32_t phase = 0;
    int b[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0}, oh[4] = {0, 0, 0, 0}, ow[4] = {0, 0, 0, 0};
    int mt_live = 0;                                   // tiles of the group that start inside the sample
    for (int mt = 0; mt < MT; ++mt) {
      const long long m0 = m_base + (long long)mt * BLOCK_M;
      if (m0 < p.M) {
        mt_live = mt + 1;
        if (tp.a.mode == 2) tm_decode_row(p, m0, b[mt], od[mt], oh[mt], ow[mt]);
      }
    }
    int tap_i = 0, slab = 0;
    for (int kb = 0; kb < p.num_kb; ++kb) {
      mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
      const uint32_t sst = smem_base + stage * stage_bytes;
      mbar_expect_tx_elect(land_bar0 + 8 * stage, (uint32_t)(mt_live * A_TILE_BYTES));
      for (int mt = 0; mt < mt_live; ++mt)
        if (CLN == 1 || mt % CLN == crank)
          tma_issue_a(tp, sst + A_OFF + mt * NB * A_TILE_BYTES, land_bar0 + 8 * stage, img_base, g,
                      m_base + (long long)mt * BLOCK_M, tap_i, slab, b[mt], od[mt], oh[mt], ow[mt], cmask);
      if (++slab == slabs) {
        slab = 0;
        ++tap_i;
      }
      if (++stage == NSTG) {
        stage = 0;
        phase ^= 1;
      }
    }
    __syncwarp();
  } else if (XFORM && warp >= TM_CONV_WARP0 && warp < TM_CONV_WARP0 + TM_CONV_WARPS) {
    // ============================================================== transform warps: tf32 rounding and / or the x * s_in plane
    const int ctid = tid - TM_CONV_WARP0 * 32;           // 0..127 = the row of a tile this thread owns (Flipout)
    const long long left = (p.M - m_base + BLOCK_M - 1) / BLOCK_M;
    const int mt_live = left < MT ? (int)left : MT;
    // Flipout: where this thread's row of every tile sits in the input (pixel of the window origin)
    int rb[4] = {0, 0, 0, 0}, rz[4] = {0, 0, 0, 0}, ry[4] = {0, 0, 0, 0}, rx[4] = {0, 0, 0, 0};
    bool rok[4] = {false, false, false, false};
    if constexpr (FLIP) {
      for (int mt = 0; mt < mt_live; ++mt) {
        const long long m = m_base + (long long)mt * BLOCK_M + ctid;
        rok[mt] = m < p.M;
        if (rok[mt]) {
          int b_, od_, oh_, ow_;
          tm_decode_row(p, m, b_, od_, oh_, ow_);
          rb[mt] = b_; rz[mt] = od_ * p.sd - p.pd; ry[mt] = oh_ * p.sh - p.ph; rx[mt] = ow_ * p.sw - p.pw;
        }
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    int tap_i = 0, slab = 0;
    uint32_t sb_row[4] = {0xfffffffeu, 0xfffffffeu, 0xfffffffeu, 0xfffffffeu}, sb_blk[4] = {0u, 0u, 0u, 0u};   // cached sign blocks
    uint4 sb_val[4];
    for (int kb = 0; kb < p.num_kb; ++kb) {
      int dz = 0, dy = 0, dx = 0;
      if (FLIP && tp.a.mode == 2) {
        const TapCoord tc = decode_tap(p, tap_i);
        dz = tc.dz; dy = tc.dy; dx = tc.dx;
      }
      const int cg = g * p.Cin_g + slab * KBE;           // global channel of chunk 0 of this k-block
      if (++slab == slabs) {
        slab = 0;
        ++tap_i;
      }
      mbar_wait(afull_bar0 + 8 * stage, phase);
      const uint32_t sst = smem_base + stage * stage_bytes + A_OFF;
      for (int mt = 0; mt < mt_live; ++mt) {
        const uint32_t t0 = sst + mt * NB * A_TILE_BYTES;
        if constexpr (!FLIP) {
          tm_round_tile_tf32(t0, ctid);
        } else {
          // input sign bits of this row: Philox block (128 channels) of the input pixel this row reads for this tap
          // (the sign block of a row covers 128 channels = 2 (bf16) / 4 (tf32) k-blocks of a linear layer or of one tap:
          //  it is recomputed only when the (pixel, channel block) pair changes)
          const int z = rz[mt] + dz, y = ry[mt] + dy, xw = rx[mt] + dx;
          const bool inb = rok[mt] && (unsigned)z < (unsigned)p.ID && (unsigned)y < (unsigned)p.IH && (unsigned)xw < (unsigned)p.IW;
          const uint32_t prow = inb ? (uint32_t)((((long long)rb[mt] * p.ID + z) * p.IH + y) * p.IW + xw) : 0xffffffffu;
          const uint32_t bkey = (uint32_t)(cg >> 7);
          if (prow != sb_row[mt] || bkey != sb_blk[mt]) {
            sb_row[mt] = prow;
            sb_blk[mt] = bkey;
            sb_val[mt] = inb ? bt_sign_block(p.key, BT_STREAM_SIGN_IN, bkey, prow, sample) : make_uint4(0u, 0u, 0u, 0u);
          }
          const uint4 blk = sb_val[mt];
          const int r = ctid;
          // all eight 16-byte chunks of the row in flight before the first is used (one warp per scheduler: latency-bound)
          uint4 vv[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) vv[c] = lds16(t0 + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)));
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = t0 + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
            uint4 v = vv[c];
            const int ch = (cg & 127) + c * (TF32 ? 4 : 8);          // channel (inside the 128-block) of the chunk's element 0
            const uint32_t bits = bt_sign_word(blk, (uint32_t)(ch >> 5)) >> (ch & 31);
            if constexpr (TF32) {
              v = make_uint4(bt_tf32(__uint_as_float(v.x)), bt_tf32(__uint_as_float(v.y)), bt_tf32(__uint_as_float(v.z)),
                             bt_tf32(__uint_as_float(v.w)));
              sts16(a, v);
              v.x ^= (bits & 1u) << 31; v.y ^= (bits & 2u) << 30; v.z ^= (bits & 4u) << 29; v.w ^= (bits & 8u) << 28;
            } else {
              const uint4 mk = sign_masks8(bits & 0xffu);
              v.x ^= mk.x; v.y ^= mk.y; v.z ^= mk.z; v.w ^= mk.w;
            }
            sts16(a + A_TILE_BYTES, v);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar0 + 8 * stage);
      if (++stage == NSTG) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp < TM_SAMP_WARPS) {
    TmSampler<BLOCK_N, P_BF16, TF32, FLIP> smp;
    smp.init(p, tid, g, n0);
    int stage = 0;
    uint32_t phase = 0;
    int tap_i = 0, slab = 0;
    auto prefetch_next = [&](auto ph) {
      const int kc = slab * KBE + smp.koff();
      const bool kvalid = tp.a.mode == 1 ? (kc < p.K_used) : true;
      const long long kphys0 = tp.a.mode == 1 ? (long long)kc : (long long)decode_tap(p, tap_i).lin * p.Cin_g + kc;
      if (++slab == slabs) {
        slab = 0;
        ++tap_i;
      }
      smp.template prefetch<decltype(ph)::value>(p, kphys0, kvalid);
    };
    tm_sample_loop(smp, p.num_kb, prefetch_next, [&](auto ph, int) {
      mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
      if (tp.a.probe != 1) smp.template compute<decltype(ph)::value>(p, sample, smem_base + stage * stage_bytes);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar0 + 8 * stage);
      if (++stage == NSTG) {
        stage = 0;
        phase ^= 1;
      }
    });
    // ---- epilogue of the MT accumulators
    constexpr int EN = BLOCK_N / 2;
    const int q4 = warp & 3, ncol0 = (warp >> 2) * EN;
    const uint32_t stg = p.dr_stage ? smem_u32(aux + TM_AUX_BYTES) + (uint32_t)(warp * 32 * EN * (TF32 ? 4 : 2)) : 0u;
    mbar_wait_idle(acc_bar, 0, 128);
    tc_fence_after();
    for (int mt = 0; mt < MT; ++mt) {
      const long long m = m_base + (long long)mt * BLOCK_M + q4 * 32 + lane;
      if (m_base + (long long)mt * BLOCK_M >= p.M) break;      // (warp-uniform) tile beyond the sample: never loaded
      uint4 sblk = make_uint4(0u, 0u, 0u, 0u);
      if constexpr (FLIP)
        sblk = bt_sign_block(p.key, BT_STREAM_SIGN_OUT, ((uint32_t)g << 20) | (uint32_t)(n0 >> 7), (uint32_t)m, sample);
      tm_epilogue_tile<EN, TF32, FLIP>(p, bias_s, tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(mt * NB * BLOCK_N + ncol0), g, n0,
                                       ncol0, (long long)s * p.M + m, m < p.M, stg, lane, (uint32_t)BLOCK_N, sblk);
    }
  }

  tc_fence_before();
  __syncthreads();
  // no CTA leaves while a peer may still multicast into its shared memory or arrive on its barriers
  if (CLN > 1) bt_cluster_sync();
  if (warp == TM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------ kernel 3: direct windows staged by TMA
// Stride-1 "same" convolutions (groups == 1, C_in a multiple of the k-block): the trick of bt_direct_kernel -- number
// the pixels of one MC sample in a PADDED flattening (every padded row = W real pixels followed by pw zero pixels, every
// padded plane = H real rows followed by ph zero rows, ...) so that filter tap t pairs output pixel q with input pixel
// q + delta_t for EVERY q; a window of consecutive padded pixels, stored row by row in the 128B-swizzled K-major
// layout, then IS the A operand of every tap (the tap's descriptor starts delta_t rows further) -- with the window
// staged by the TMA instead of 1216 cp.async per tile issued by 7 producer warps:
//   * a tile is dt_k whole padded rows (dt_k * Pw <= 128 pixels); its window = those rows plus dt_hr halo rows on each
//     side = a handful of TILED TMA boxes {kbe channels, Pw pixels, dt_hb rows}: pixels w >= W, rows h >= H, planes
//     d >= D and images outside the sample are OUT OF RANGE for the tensor map and arrive as zeros -- the zero padding is
//     produced by the copy engine (tests/test_gpu_tma.py pins the zero fill and the address-based swizzle of a box
//     that lands at a 128-byte-aligned row of the slot);
//   * L2 / HBM see every activation ONCE per n-tile instead of once per filter tap (an im2col map re-reads it per tap:
//     measured L2-bound, profiles/r02*), and no warp spends instructions on the gather;
//   * warps 0-7 sample the resident W_s then run the epilogue, warp 8 issues the TMA boxes, warp 9 the MMAs;
//     tf32: warps 10-13 round each landed window to tf32 in place.
__device__ __forceinline__ void tma_load_5d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int d, int n) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}\n" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n)
      : "memory");
}

// FLIP (Flipout): every window slot holds two operand copies -- plane 0 = x as staged by the TMA, plane 1 = x * s_in
// written by the transform warps (one Philox call per window pixel and 128-channel block) -- the resident tiles come in
// (mu, sigma*eps) pairs, two accumulators per tile, output signs in the epilogue.
template <int BLOCK_N, bool P_BF16, bool TF32, bool FLIP>
__global__ void __launch_bounds__(tm_threads<TF32 || FLIP>(), 1) bt_dtma_kernel(const __grid_constant__ DtParams dp) {
  const FusedParams& p = dp.f;
  const DtGeom& G = dp.g;
  constexpr int NB = FLIP ? 2 : 1;
  constexpr bool XFORM = TF32 || FLIP;
  constexpr int B_TILE_BYTES = BLOCK_N * 128;
  constexpr int KBE = TF32 ? 32 : 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int slabs = dp.slabs;
  const int res_bytes = p.num_kb * NB * B_TILE_BYTES;
  const uint32_t plane_bytes = (uint32_t)(G.R * 128);              // one slab of one operand copy
  const uint32_t copy_bytes = (uint32_t)slabs * plane_bytes;       // one operand copy of a window (all slabs)
  const uint32_t slot_bytes = NB * copy_bytes;
  const int NS = G.slots;
  uint8_t* aux = smem + res_bytes + NS * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  float* bias_s = reinterpret_cast<float*>(aux + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 5);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t win0 = smem_base + res_bytes;
  const uint32_t wfull_bar0 = smem_u32(bars);                       // window ready for the tensor core
  const uint32_t wempty_bar0 = smem_u32(bars + MAX_STAGES);
  const uint32_t bready_bar = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t acc_bar0 = smem_u32(bars + 2 * MAX_STAGES + 1);    // [2]
  const uint32_t tfree_bar0 = smem_u32(bars + 2 * MAX_STAGES + 3);  // [2]
  const uint32_t wland_bar0 = smem_u32(bars + 2 * MAX_STAGES + 5);  // tf32: TMA bytes landed (converter warps wait here)
  const uint32_t land_bar0 = XFORM ? wland_bar0 : wfull_bar0;

  const int s = blockIdx.z;
  const int n0 = blockIdx.y * BLOCK_N;               // groups == 1
  const uint32_t sample = p.sample0 + (uint32_t)s + (p.sample_ptr != nullptr ? __ldg(p.sample_ptr) : 0u);
  const int img_base = p.x_shared ? 0 : s * p.B;
  const long long n_rt = p.n_groups;                 // tiles of dt_k padded rows per sample
  const int CLX = dp.clx > 1 ? dp.clx : 1;            // cluster along x: the CTAs of one (sample, n-tile) share the sampling
  const int crank = CLX > 1 ? (int)bt_cluster_ctarank() : 0;
  const int data_rows = (G.k + 2 * G.hr) * G.Pw;     // window rows that carry data
  const int load_rows = G.halo ? data_rows : G.k * G.Pw;           // rows the TMA writes per slab (halo == 0: body only)
  const uint32_t win_bytes = (uint32_t)slabs * (uint32_t)load_rows * 128u;
  const int zero_rows = G.Z + (G.halo ? 0 : G.hr * G.Pw);         // never written by a box: zero for the whole kernel
  const int xf_lo = G.halo ? 0 : G.hr * G.Pw, xf_hi = xf_lo + load_rows;   // window rows the transform warps touch

  if (warp == TM_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) {
        mbar_init(wfull_bar0 + 8 * i, XFORM ? TM_CONV_WARPS : 1);
        mbar_init(wempty_bar0 + 8 * i, 1);
        mbar_init(wland_bar0 + 8 * i, 1);
      }
      mbar_init(bready_bar, TM_SAMP_WARPS);
      for (int i = 0; i < 2; ++i) {
        mbar_init(acc_bar0 + 8 * i, 1);
        mbar_init(tfree_bar0 + 8 * i, TM_SAMP_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  } else if (warp == TM_TMA_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&dp.map_a);
      tma_prefetch_desc(&dp.map_h);
    }
  } else if (tid < BLOCK_N) {
    tm_fill_bias<P_BF16, FLIP>(p, bias_s, tid, 0, n0, sample);
  }
  // the Z rows in front of every slab plane stay zero for the whole kernel (pad pixels just before the window)
  for (int i = tid; i < NS * NB * slabs * zero_rows * 8; i += blockDim.x) {
    const int pl = i / (zero_rows * 8), r = i - pl * (zero_rows * 8);
    sts16(win0 + (uint32_t)pl * plane_bytes + (uint32_t)r * 16u, make_uint4(0u, 0u, 0u, 0u));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // split cluster barrier of the shared sampling prologue: every thread arrives once -- the samplers after their last
  // remote store, everybody else now -- and waits once: the MMA warp before its first MMA, the others before the exit
  if (CLX > 1 && warp >= TM_SAMP_WARPS) bt_cluster_arrive();

  if (warp == TM_MMA_WARP) {
    // ============================================================== MMA issuer
    const uint32_t idesc = make_idesc(BLOCK_N, TF32);
    const uint64_t desc_hi = make_smem_desc(0u);
    long long it = 0;
    int slot = 0;
    uint32_t wpar = 0;
    if (CLX > 1) bt_cluster_wait();                   // every CTA of the cluster has written its k-blocks of W_s
    if (bt_elect_one()) {                             // one thread issues every MMA (umma1_x4)
    mbar_wait_idle(bready_bar, 0, 256);
    tc_fence_after();
    for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      mbar_wait_idle(wfull_bar0 + 8 * slot, wpar, 32);
      if (it >= 2) mbar_wait_idle(tfree_bar0 + 8 * buf, (uint32_t)(((it >> 1) - 1) & 1), 32);
      tc_fence_after();
      const uint32_t wslot16 = ((win0 + (uint32_t)slot * slot_bytes) & 0x3FFFFu) >> 4;
      const uint32_t acc = tmem_base + (uint32_t)(buf * NB * BLOCK_N);
      uint32_t b16 = (smem_base & 0x3FFFFu) >> 4;
#pragma unroll 2
      for (int kb = 0; kb < p.num_kb; ++kb, b16 += (uint32_t)(NB * B_TILE_BYTES) >> 4) {
        const uint32_t a16 = wslot16 + (uint32_t)p.dr_aoff[kb];   // (slab * R + Z + hr * Pw + delta_tap) rows, in 16-byte units
        umma1_x4<TF32>(acc, a16 | (1u << 16), b16 | (1u << 16), (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
        if constexpr (FLIP)
          umma1_x4<TF32>(acc + BLOCK_N, (a16 + (copy_bytes >> 4)) | (1u << 16), (b16 + (B_TILE_BYTES >> 4)) | (1u << 16),
                         (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
      }
      umma1_commit(wempty_bar0 + 8 * slot);
      umma1_commit(acc_bar0 + 8 * buf);
      if (++slot == NS) {
        slot = 0;
        wpar ^= 1u;
      }
    }
    }
    __syncwarp();
  } else if (warp == TM_TMA_WARP) {
    // ============================================================== TMA producer: the window boxes of every tile
    int slot = 0;
    uint32_t wpar = 0;
    for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x) {
      mbar_wait(wempty_bar0 + 8 * slot, wpar ^ 1);
      mbar_expect_tx_elect(land_bar0 + 8 * slot, win_bytes);
      const long long row0 = rt * G.k;                              // first padded row (within the sample) of the tile
      const uint32_t wslot = win0 + (uint32_t)slot * slot_bytes + (uint32_t)G.Z * 128u;
      // one box: padded row rr (decoded to (h, d, image)); rows outside the sample come from image -1 = all zeros
      auto issue = [&](const CUtensorMap* map, long long rr, uint32_t dst_row) {
        int h = 0, d = 0, n = -1;
        if (rr >= 0 && rr < G.NR) {
          const uint32_t r32 = (uint32_t)rr;
          const uint32_t t1 = (uint32_t)(((unsigned long long)r32 * G.mulh) >> G.shh);      // r / Ph
          h = (int)(r32 - t1 * (uint32_t)G.Ph);
          const uint32_t b = (uint32_t)(((unsigned long long)t1 * G.muld) >> G.shd);        // / Pd
          d = (int)(t1 - b * (uint32_t)G.Pd);
          n = img_base + (int)b;
        }
        const uint32_t dst = wslot + dst_row * 128u;
        for (int sl = 0; sl < slabs; ++sl)
          tma_load_5d_elect(dst + (uint32_t)sl * plane_bytes, map, land_bar0 + 8 * slot, sl * KBE, 0, h, d, n);
      };
      if (G.halo)
        for (int i = 0; i < G.hr; ++i) issue(&dp.map_h, row0 - G.hr + i, (uint32_t)(i * G.Pw));
      for (int i = 0; i * G.unit < G.k; ++i) issue(&dp.map_a, row0 + (long long)i * G.unit, (uint32_t)((G.hr + i * G.unit) * G.Pw));
      if (G.halo)
        for (int i = 0; i < G.hr; ++i) issue(&dp.map_h, row0 + G.k + i, (uint32_t)((G.hr + G.k + i) * G.Pw));
      if (++slot == NS) {
        slot = 0;
        wpar ^= 1u;
      }
    }
    __syncwarp();
  } else if (XFORM && warp >= TM_CONV_WARP0 && warp < TM_CONV_WARP0 + TM_CONV_WARPS) {
    // ============================================================== transform warps: tf32 rounding in place and / or the
    // x * s_in copy of every landed window
    const int ctid = tid - TM_CONV_WARP0 * 32;
    int slot = 0;
    uint32_t wpar = 0;
    for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x) {
      mbar_wait(wland_bar0 + 8 * slot, wpar);
      const uint32_t wslot = win0 + (uint32_t)slot * slot_bytes + (uint32_t)G.Z * 128u;
      if constexpr (!FLIP) {
        for (int sl = 0; sl < slabs; ++sl) {
          const uint32_t base = wslot + (uint32_t)sl * plane_bytes;
          for (int c = xf_lo * 8 + ctid; c < xf_hi * 8; c += TM_CONV_WARPS * 32) {
            const uint32_t a = base + (uint32_t)c * 16u;
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
            sts16(a, make_uint4(bt_tf32(__uint_as_float(v.x)), bt_tf32(__uint_as_float(v.y)), bt_tf32(__uint_as_float(v.z)),
                                bt_tf32(__uint_as_float(v.w))));
          }
        }
      } else {
        constexpr int SPB = 128 / KBE;                            // slabs per 128-channel sign block: 2 (bf16) | 4 (tf32)
        const long long pix_first = (rt * G.k - G.hr) * (long long)G.Pw;     // padded pixel of window row 0
        for (int j = xf_lo + ctid; j < xf_hi; j += TM_CONV_WARPS * 32) {
          // window row j = padded pixel pix_first + j -> (valid real pixel?, dense pixel index inside the sample)
          const long long q = pix_first + j;
          bool ok = q >= 0;
          uint32_t prow = 0;
          if (ok) {
            const uint32_t q32 = (uint32_t)q;
            const uint32_t rr = (uint32_t)(((unsigned long long)q32 * G.mulw) >> G.shw);   // padded row
            const uint32_t w = q32 - rr * (uint32_t)G.Pw;
            const uint32_t t1 = (uint32_t)(((unsigned long long)rr * G.mulh) >> G.shh);
            const uint32_t h = rr - t1 * (uint32_t)G.Ph;
            const uint32_t b = (uint32_t)(((unsigned long long)t1 * G.muld) >> G.shd);
            const uint32_t d = t1 - b * (uint32_t)G.Pd;
            ok = (long long)rr < G.NR && w < (uint32_t)p.IW && h < (uint32_t)p.IH && d < (uint32_t)p.ID;
            prow = ((b * (uint32_t)p.ID + d) * (uint32_t)p.IH + h) * (uint32_t)p.IW + w;
          }
          for (int sb_ = 0; sb_ * SPB < slabs; ++sb_) {
            uint4 blk = make_uint4(0u, 0u, 0u, 0u);
            if (ok) blk = bt_sign_block(p.key, BT_STREAM_SIGN_IN, (uint32_t)sb_, prow, sample);
#pragma unroll
            for (int hs = 0; hs < SPB; ++hs) {
              const int sl = sb_ * SPB + hs;
              if (sl < slabs) {
                const uint32_t ra = wslot + (uint32_t)sl * plane_bytes + (uint32_t)j * 128u;
                const int jj = G.Z + j;                            // absolute row inside the plane (swizzle phase)
                uint4 vv[8];                                       // the whole row in flight before the first use
#pragma unroll
                for (int c = 0; c < 8; ++c) vv[c] = lds16(ra + (uint32_t)((c ^ (jj & 7)) << 4));
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const uint32_t a = ra + (uint32_t)((c ^ (jj & 7)) << 4);
                  uint4 v = vv[c];
                  const int ch = hs * KBE + c * (TF32 ? 4 : 8);    // channel inside the 128-block
                  const uint32_t bits = bt_sign_word(blk, (uint32_t)(ch >> 5)) >> (ch & 31);
                  if constexpr (TF32) {
                    v = make_uint4(bt_tf32(__uint_as_float(v.x)), bt_tf32(__uint_as_float(v.y)), bt_tf32(__uint_as_float(v.z)),
                                   bt_tf32(__uint_as_float(v.w)));
                    sts16(a, v);
                    v.x ^= (bits & 1u) << 31; v.y ^= (bits & 2u) << 30; v.z ^= (bits & 4u) << 29; v.w ^= (bits & 8u) << 28;
                  } else {
                    const uint4 mk = sign_masks8(bits & 0xffu);
                    v.x ^= mk.x; v.y ^= mk.y; v.z ^= mk.z; v.w ^= mk.w;
                  }
                  sts16(a + copy_bytes, v);
                }
              }
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(wfull_bar0 + 8 * slot);
      if (++slot == NS) {
        slot = 0;
        wpar ^= 1u;
      }
    }
  } else if (warp < TM_SAMP_WARPS) {
    // ============================================================== warps 0-7: sample W_s, then epilogue
    {
      TmSampler<BLOCK_N, P_BF16, TF32, FLIP> smp;
      smp.init(p, tid, 0, n0);
      if (CLX > 1) {
        // Cluster-shared prologue: the CLX CTAs of this (sample, n-tile) would each sample the SAME W_s; instead CTA r
        // samples the k-blocks kb = r (mod CLX) and writes every tile into all CLX shared memories (DSMEM).
        smp.ncl = (uint32_t)CLX;
        int kb_next = crank;
        auto prefetch_next = [&](auto ph) {
          const int t_ = kb_next / slabs, sl_ = kb_next - t_ * slabs;
          kb_next += CLX;
          smp.template prefetch<decltype(ph)::value>(p, (long long)decode_tap(p, t_).lin * p.Cin_g + sl_ * KBE + smp.koff(), true);
        };
        tm_sample_loop_strided(smp, p.num_kb, crank, CLX, prefetch_next, [&](auto ph, int kb) {
          smp.template compute<decltype(ph)::value>(p, sample, smem_base + kb * NB * B_TILE_BYTES);
        });
        fence_proxy_async_all();                          // generic-proxy stores (local and remote) -> async proxy (UMMA)
        __syncwarp();
        bt_cluster_arrive();                               // (the non-sampler warps arrived right after the set-up)
      } else {
      int tap_i = 0, slab = 0;
      auto prefetch_next = [&](auto ph) {
        const long long kphys0 = (long long)decode_tap(p, tap_i).lin * p.Cin_g + slab * KBE + smp.koff();
        if (++slab == slabs) {
          slab = 0;
          ++tap_i;
        }
        smp.template prefetch<decltype(ph)::value>(p, kphys0, true);
      };
      tm_sample_loop(smp, p.num_kb, prefetch_next, [&](auto ph, int kb) {
        smp.template compute<decltype(ph)::value>(p, sample, smem_base + kb * NB * B_TILE_BYTES);
      });
      fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bready_bar);
    }
    constexpr int EN = BLOCK_N / 2;
    const int q4 = warp & 3, ncol0 = (warp >> 2) * EN;
    const uint32_t stg = p.dr_stage ? smem_u32(aux + DT_AUX_BYTES) + (uint32_t)(warp * 32 * EN * (TF32 ? 4 : 2)) : 0u;
    const int j = q4 * 32 + lane;                                   // this lane's pixel inside the tile
    const uint32_t jr = (uint32_t)(((unsigned long long)(uint32_t)j * G.mulw) >> G.shw);   // j / Pw
    const int jw = j - (int)jr * G.Pw;
    const bool jok = (int)jr < G.k && jw < p.IW;
    // output row of this lane in tile rt: padded row rr -> (b, d, h); valid iff a real pixel of a real image
    auto row_of = [&](long long rt, long long& m) -> bool {
      const long long rr = rt * G.k + jr;
      bool ok = jok && rr < G.NR;
      m = 0;
      if (ok) {
        const uint32_t r32 = (uint32_t)rr;
        const uint32_t t1 = (uint32_t)(((unsigned long long)r32 * G.mulh) >> G.shh);
        const uint32_t h = r32 - t1 * (uint32_t)G.Ph;
        const uint32_t b = (uint32_t)(((unsigned long long)t1 * G.muld) >> G.shd);
        const uint32_t d = t1 - b * (uint32_t)G.Pd;
        ok = h < (uint32_t)p.IH && d < (uint32_t)p.ID;
        m = (((long long)b * p.ID + d) * p.IH + h) * p.IW + jw;
      }
      return ok;
    };
    // Residual prefetch TWO tiles ahead, in two register slots that are never copied: one MMA-bound tile (~1.4 us) is
    // about the HBM latency under load, and a struct copy of a slot (`cur = pre`) is a use of the loaded registers -- it
    // waited for the data right there (profiles/r02q: 25% of the kernel's stall samples on those moves; tools/res_probe.py:
    // no prefetch 197 us, one tile ahead 141 us, residual ignored 92 us per layer1 launch).  The tile loop is unrolled by
    // two so that slot A serves the even and slot B the odd tiles of this CTA.
    const bool res_pf = TmResidual<EN, TF32>::PREFETCH && !FLIP && p.ep_residual != nullptr && stg != 0u && p.out_vec &&
                        n0 + ncol0 + EN <= p.N && p.probe < 2;
    auto tile_loop = [&](auto PF) {
      constexpr bool pf = decltype(PF)::value;
      TmResidual<EN, TF32> slot_a, slot_b;
      slot_a.have = slot_b.have = false;
      long long it = 0;
      auto prefetch = [&](TmResidual<EN, TF32>& slot, long long rt) {
        if constexpr (pf) {                                  // (past the CTA's last tile: re-read the last one, unused)
          long long m_;
          const bool ok = row_of(rt < n_rt ? rt : n_rt - 1, m_);
          slot.fetch_always(p, 0, n0, ncol0, (long long)s * p.M + (ok ? m_ : 0), ok, lane);
        }
      };
      auto do_tile = [&](TmResidual<EN, TF32>& slot, long long rt) {
        const int buf = (int)(it & 1);
        long long m = 0;
        const bool mvalid = row_of(rt, m);
        mbar_wait_idle(acc_bar0 + 8 * buf, (uint32_t)((it >> 1) & 1), 128);
        tc_fence_after();
        uint4 sblk = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (FLIP) sblk = bt_sign_block(p.key, BT_STREAM_SIGN_OUT, (uint32_t)(n0 >> 7), (uint32_t)m, sample);
        tm_epilogue_tile<EN, TF32, FLIP>(p, bias_s, tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * NB * BLOCK_N + ncol0), 0,
                                         n0, ncol0, (long long)s * p.M + m, mvalid, stg, lane, (uint32_t)BLOCK_N, sblk, &slot);
        prefetch(slot, rt + 2 * (long long)gridDim.x);   // the slot is free again: the residual of this CTA's tile after next
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tfree_bar0 + 8 * buf);
        ++it;
      };
      if ((long long)blockIdx.x < n_rt) {
        prefetch(slot_a, blockIdx.x);
        prefetch(slot_b, (long long)blockIdx.x + gridDim.x);
      }
      for (long long rt = blockIdx.x; rt < n_rt; rt += 2 * (long long)gridDim.x) {
        do_tile(slot_a, rt);
        if (rt + gridDim.x < n_rt) do_tile(slot_b, rt + gridDim.x);
      }
    };
    if constexpr (TmResidual<EN, TF32>::PREFETCH && !FLIP) {
      if (res_pf) tile_loop(TmPh<1>{});
      else tile_loop(TmPh<0>{});
    } else {
      tile_loop(TmPh<0>{});
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CLX > 1 && warp != TM_MMA_WARP) bt_cluster_wait();   // (the MMA warp waited before its first MMA)
  if (warp == TM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

#endif  // BT_TMA_DEVICE

// host: window geometry of the TMA direct kernel for this layer, or false
inline bool dt_plan(const FusedParams& p, bool tf32, bool flip, int bn, int nkb, DtGeom* out, int* smem_total, int* ep_stage) {
  const int kbe = p.x_is_bf16 ? 64 : 32;
  if (!p.x_is_bf16 && !tf32) return false;
  const bool same = p.sd == 1 && p.sh == 1 && p.sw == 1 && p.ID == p.OD && p.IH == p.OH && p.IW == p.OW && p.groups == 1 &&
                    p.Cin_g % kbe == 0 && !p.transposed;
  if (!same) return false;
  DtGeom g;
  memset(&g, 0, sizeof(g));
  g.Pw = p.IW + p.pw; g.Ph = p.IH + p.ph; g.Pd = p.ID + p.pd;
  if (g.Pw > 128 || g.Pw > 256) return false;
  g.NR = (long long)p.B * g.Pd * g.Ph;
  if (g.NR >= (1ll << 31)) return false;
  const int hrn = p.pd * g.Ph + p.ph;                  // halo rows a tile needs on each side
  const long long halo = (long long)hrn * g.Pw + p.pw;
  const int slabs = p.Cin_g / kbe;
  const int NBp = flip ? 2 : 1;                        // operand copies (Flipout: x and x * s_in; mu and sigma*eps tiles)
  const long long res = (long long)nkb * NBp * bn * 128;
  int best_hb = 0;
  double best_score = 1e300;
  // Body boxes: `hb` padded rows inside one plane (hb divides Ph, so a box never straddles two planes) or, for 2-D
  // convolutions on tiny images, `nbp` whole padded planes (= images); halo rows: one-row boxes.  A TMA box costs
  // ~350 clocks of the copy engine almost independently of its size (measured, profiles/r02e: 16 one-row boxes per
  // tile made the kernel TMA-issue bound), so pick the shape with the least time per useful output pixel.
  // clocks per tcgen05.mma with both operands in shared memory: 43 + N / 2 (measured, tools/probes/mma_probe.cu)
  static const bool no_align = getenv("BT_DTMA_NO_ALIGN") != nullptr;             // A/B switches
  static const int hb_only = getenv("BT_DTMA_HB") ? atoi(getenv("BT_DTMA_HB")) : 0;
  const double mma1 = 43.0 + 0.5 * bn;
  const double t_mma = NBp * nkb * 4.0 * mma1 + 100.0;
  if (hrn > 8) return false;
  for (int hb = 1; hb <= g.Ph && hb * g.Pw <= 128; ++hb) {
    if (g.Ph % hb != 0 || hb > 256) continue;
    if (hb_only && hb != hb_only) continue;
    const int nbp_max = (hb == g.Ph && g.Pd == 1) ? 128 / (g.Pw * g.Ph) : 1;
    for (int nbp = 1; nbp <= nbp_max && nbp <= 256; ++nbp) {
      const int unit = hb * nbp;
      const int k = (128 / g.Pw) / unit * unit;
      if (k < unit) continue;
      const int hr = hrn;
      const bool aligned = hb == g.Ph && g.Pd == 1 && !no_align;
      const int nbox = k / unit + (aligned ? 0 : 2 * hr);
      const int Z = p.pw;
      long long rows = (long long)(k + 2 * hr) * g.Pw;
      const long long reach = (long long)hr * g.Pw + halo + 128;
      if (reach > rows) rows = reach;
      const int R = (int)((Z + rows + 7) / 8 * 8);
      const long long slot = (long long)NBp * slabs * R * 128;
      // epilogue staging buffer (coalesced global access): taken when it leaves >= 2 window slots (>= 3 preferred)
      const long long stage_b = 128ll * bn * (p.x_is_bf16 ? 2 : 4);
      long long ns = (SMEM_BUDGET - DT_AUX_BYTES - 1024 - res - stage_b) / slot;
      int stg_on = 1;
      if (ns < 2) {
        ns = (SMEM_BUDGET - DT_AUX_BYTES - 1024 - res) / slot;
        stg_on = 0;
      }
      if (ns > MAX_STAGES) ns = MAX_STAGES;
      if (ns < 2) continue;
      const double t_tma = nbox * slabs * 350.0 + 300.0;
      const double score = (t_mma > t_tma ? t_mma : t_tma) * (ns < 3 ? 1.2 : 1.0) * (stg_on ? 1.0 : 1.3) / (double)(k * p.IW);
      if (score >= best_score) continue;
      g.hb = hb; g.nbp = nbp; g.unit = unit; g.k = k; g.hr = hr; g.nbox = nbox; g.R = R; g.Z = Z; g.slots = (int)ns;
      g.halo = aligned ? 0 : 1;
      *smem_total = (int)(res + ns * slot + DT_AUX_BYTES + (stg_on ? stage_b : 0) + 1024);
      *ep_stage = stg_on;
      best_hb = hb; best_score = score;
    }
  }
  if (!best_hb) return false;
  const long long divs[3] = {g.Pw, g.Ph, g.Pd};
  uint32_t* muls[3] = {&g.mulw, &g.mulh, &g.muld};
  uint32_t* shs[3] = {&g.shw, &g.shh, &g.shd};
  for (int i = 0; i < 3; ++i) {
    int l = 0;
    while ((1ll << l) < divs[i]) ++l;
    *shs[i] = (uint32_t)(31 + l);
    *muls[i] = (uint32_t)((((unsigned long long)1 << (31 + l)) + (unsigned long long)divs[i] - 1) / (unsigned long long)divs[i]);
  }
  *out = g;
  return true;
}

// tensor maps of x for the window boxes: 5-D (C, W, H, D, N); body box {kbe, Pw, hb, 1, nbp}, halo box {kbe, Pw, 1, 1, 1}
inline int dt_encode(const FusedParams& p, const DtGeom& g, const void* x, CUtensorMap* map, bool halo_map) {
  BT_REQUIRE(tma_driver_ready(), BT_ERR_UNSUPPORTED, "TMA: cuTensorMapEncodeTiled not available from this driver");
  const int es = p.x_is_bf16 ? 2 : 4;
  const long long n_img = (long long)(p.x_shared ? 1 : p.S) * p.B;
  cuuint64_t dims[5] = {(cuuint64_t)p.C_in, (cuuint64_t)p.IW, (cuuint64_t)p.IH, (cuuint64_t)p.ID, (cuuint64_t)n_img};
  cuuint64_t st[4];
  st[0] = (cuuint64_t)p.C_in * es;
  st[1] = st[0] * p.IW;
  st[2] = st[1] * p.IH;
  st[3] = st[2] * p.ID;
  cuuint32_t box[5] = {(cuuint32_t)(128 / es), (cuuint32_t)g.Pw, (cuuint32_t)(halo_map ? 1 : g.hb), 1,
                       (cuuint32_t)(halo_map ? 1 : g.nbp)};
  cuuint32_t es5[5] = {1, 1, 1, 1, 1};
  const CUresult r = g_tma.tiled(map, p.x_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                                 const_cast<void*>(x), dims, st, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BT_REQUIRE(r == CUDA_SUCCESS, BT_ERR_CUDA, "TMA: cuTensorMapEncodeTiled (window map) failed with CUresult %d", (int)r);
  return BT_OK;
}

#ifdef BT_TMA_DEVICE
template <int BN, bool PB, bool TF32, bool FLIP>
int launch_dtma(const DtParams& dp, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  static bool attr_done[64] = {};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!attr_done[dev]) {
      BT_CHECK_CUDA(cudaFuncSetAttribute(bt_dtma_kernel<BN, PB, TF32, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
      attr_done[dev] = true;
    }
  }
  if (dp.clx > 1) {        // thread-block clusters along x (shared sampling prologue)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(tm_threads<TF32 || FLIP>(), 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)dp.clx;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    BT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, bt_dtma_kernel<BN, PB, TF32, FLIP>, dp));
    return BT_OK;
  }
  bt_dtma_kernel<BN, PB, TF32, FLIP><<<grid, tm_threads<TF32 || FLIP>(), smem_bytes, st>>>(dp);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

template <int BN>
int dispatch_dtma(const DtParams& dp, bool tf32, bool flip, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  if (flip) {
    if (tf32) return launch_dtma<BN, false, true, true>(dp, grid, smem_bytes, dev, st);
    return dp.f.p_is_bf16 ? launch_dtma<BN, true, false, true>(dp, grid, smem_bytes, dev, st)
                          : launch_dtma<BN, false, false, true>(dp, grid, smem_bytes, dev, st);
  }
  if (tf32) return launch_dtma<BN, false, true, false>(dp, grid, smem_bytes, dev, st);
  return dp.f.p_is_bf16 ? launch_dtma<BN, true, false, false>(dp, grid, smem_bytes, dev, st)
                        : launch_dtma<BN, false, false, false>(dp, grid, smem_bytes, dev, st);
}

template <int BN, bool PB, bool TF32>
int launch_tma(const TmaParams& tp, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  static bool attr_done[64] = {};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!attr_done[dev]) {
      BT_CHECK_CUDA(cudaFuncSetAttribute(bt_tma_kernel<BN, PB, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
      attr_done[dev] = true;
    }
  }
  bt_tma_kernel<BN, PB, TF32><<<grid, tm_threads<TF32>(), smem_bytes, st>>>(tp);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

template <int BN, bool PB, bool TF32, bool FLIP>
int launch_tms(const TmaParams& tp, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  static bool attr_done[64] = {};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!attr_done[dev]) {
      BT_CHECK_CUDA(cudaFuncSetAttribute(bt_tms_kernel<BN, PB, TF32, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
      attr_done[dev] = true;
    }
  }
  if (tp.a.cln > 1) {      // thread-block clusters along the n-tiles (A-operand multicast)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(tm_threads<TF32 || FLIP>(), 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = (unsigned)tp.a.cln;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    BT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, bt_tms_kernel<BN, PB, TF32, FLIP>, tp));
    return BT_OK;
  }
  bt_tms_kernel<BN, PB, TF32, FLIP><<<grid, tm_threads<TF32 || FLIP>(), smem_bytes, st>>>(tp);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

template <int BN>
int dispatch_tma(const TmaParams& tp, bool tf32, bool stream_mode, bool flip, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  if (stream_mode && flip) {
    if constexpr (BN >= 64) {
      if (tf32) return launch_tms<BN, false, true, true>(tp, grid, smem_bytes, dev, st);
      return tp.f.p_is_bf16 ? launch_tms<BN, true, false, true>(tp, grid, smem_bytes, dev, st)
                            : launch_tms<BN, false, false, true>(tp, grid, smem_bytes, dev, st);
    } else {
      bt_set_error("bt_tms_kernel: Flipout needs BLOCK_N >= 64");
      return BT_ERR_UNSUPPORTED;
    }
  }
  if (stream_mode) {
    if (tf32) return launch_tms<BN, false, true, false>(tp, grid, smem_bytes, dev, st);
    return tp.f.p_is_bf16 ? launch_tms<BN, true, false, false>(tp, grid, smem_bytes, dev, st)
                          : launch_tms<BN, false, false, false>(tp, grid, smem_bytes, dev, st);
  }
  if (tf32) return launch_tma<BN, false, true>(tp, grid, smem_bytes, dev, st);
  return tp.f.p_is_bf16 ? launch_tma<BN, true, false>(tp, grid, smem_bytes, dev, st)
                        : launch_tma<BN, false, false>(tp, grid, smem_bytes, dev, st);
}
#endif  // BT_TMA_DEVICE
