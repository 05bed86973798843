// Device / host pieces shared by the kernel families of libbtb200 (bt_fused.cu: bt_fused_kernel, bt_ws_kernel,
// bt_direct_kernel; bt_tma.cu: bt_tma_kernel, bt_tms_kernel, bt_dtma_kernel): launch parameters, PTX wrappers
// (mbarrier, tcgen05 alloc / mma / commit / ld, cp.async), UMMA descriptors.  Everything lives in an anonymous
// namespace of the including translation unit.
#pragma once
#include "bt_common.cuh"
#include "bt_philox.cuh"
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace {


constexpr int GENERIC_WARPS = 8;   // producer warps of the generic path
constexpr int FAST_WARPS = 16;     // producer warps of the fast path (4 per SM sub-partition)
constexpr int BLOCK_M = 128;                        // rows of one accumulator (UMMA M)
constexpr int BLOCK_K = 64;                         // bf16 per 128-byte swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * 128;
constexpr int MAX_MT = 4;
constexpr int MAX_STAGES = 8;
constexpr int MAX_TAPS = 64;
constexpr int AUX_BYTES = 12288;
constexpr int SMEM_BUDGET = 227 * 1024;

struct FusedParams {
  const void* x;
  void* out;
  const void* mu_w;
  const void* rho_w;
  const void* mu_b;
  const void* rho_b;
  const float* eps_w_in;
  const float* eps_b_in;
  const float* sign_in;
  const float* sign_out;
  const float* ep_scale;   // fused epilogue: out = out * scale[n] + shift[n]  (eval-mode BatchNorm)
  const float* ep_shift;
  const void* ep_residual; // out += residual (same layout / dtype as out)
  int ep_relu;
  float* kl_partials;
  long long M;  // output rows per sample = B*OD*OH*OW
  int S, x_shared, B;
  int C_in, C_out, groups, Cin_g, N;
  int K_phys, K_used, num_kb;
  int ID, IH, IW, OD, OH, OW, KD, KH, KW;
  int sd, sh, sw, pd, ph, pw, dd, dh, dw;
  int taps_explicit;        // taps[] lists the taps to iterate (always set on the fast path: no div/mod per stage)
  int taps_natural;         // taps[] is simply every tap in (kd, kh, kw) order
  int q64, r64;             // 64 / Cin_g, 64 % Cin_g: k-block step of a (tap, channel) cursor
  uint32_t taps[MAX_TAPS];  // kd | kh << 8 | kw << 16 of the taps that touch real data
  int MT, stages;
  int ws;        // fast path only: weight-stationary CTA (all k-blocks of the sampled tile resident in smem,
                 // the CTA loops over several groups of MT M-subtiles)
  int n_groups;  // ceil(m_tiles / MT)
  int ws_async;  // ws: gather activations with the cp.async pipeline (bf16 activations, no Flipout)
  int tc_rows;   // bt_ws_kernel tap-copy mode (stride-1 'same' convs): rows of the input window buffer, 0 = off
  int tc_halo;   //   pixels in front of the row tile held in that buffer
  int tc_padoff; //   (pd*IH + ph)*IW + pw: window-origin offset of an output pixel
  // bt_direct_kernel (bt_direct.cuh): padded pixel numbering, window geometry
  int dr_R, dr_halo;          // window rows (multiple of 8); max |tap shift|
  int dr_Pw, dr_Ph, dr_Pd;    // padded extents W+pw, H+ph, D+pd
  long long dr_Mp;            // padded pixels per sample = B * Pd * Ph * Pw
  int dr_slots;               // window ring depth (2..8)
  int dr_stage;               // 1: the epilogue has its smem staging buffer (coalesced global access)
  uint32_t dr_mul[3], dr_sh[3];   // reciprocals of dr_Pw, dr_Ph, dr_Pd (dr_div)
  int dr_aoff[64];            // A-descriptor start of k-block kb inside a window slot, in 16-byte units
  long long* dr_times;        // phase probe buffer (BT_DIRECT_TIMES), normally NULL
  int x_is_bf16, p_is_bf16;
  int rho_is_sigma;   // rho_w holds sigma = softplus(rho) already (cached by the caller for frozen parameters)
  int a_vec, w_vec, out_vec;
  int n_tiles_per_group;
  float prior_mu, log_prior_sigma, inv_2ps2;
  BtRngKey key;
  uint32_t sample0;
  const uint32_t* sample_ptr;   // optional DEVICE word added to sample0 at run time (fresh draws per CUDA-graph replay)
  int transposed;               // generic path: fractionally-strided gather (ConvTranspose{1,2,3}d)
  int probe;                    // measurement switches (BT_TMA_PROBE with BT_DYNAMIC_ENV; 0 in production): 2 = no residual
                                // prefetch, 3 = the epilogue ignores the residual (timing only)
  int pool_oh, pool_ow;         // bt_tma_kernel: 3x3 / stride 2 / pad 1 max-pool fused behind the epilogue (0 = off); the
                                // output rows of one image are pool_oh x pool_ow pixels
  uint32_t tmem_cols;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  // NOTE: no suspend-time hint -- with a hint ptxas emits a NANOSLEEP back-off loop that quantises every
  // hand-off to the hint (profiles/r01d: 10% of the stall samples); the plain form blocks in hardware until the
  // phase flips or a short system time limit expires.
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait.  Every 1024 failed probes the wall clock (%globaltimer, ns) is consulted; a protocol bug traps after 3 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0u) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 3000000000ull) __trap();
    }
  }
}
// Wait with back-off for roles that idle for a long time (epilogue warps waiting for a whole row tile, the MMA
// thread waiting for producers): a tight try_wait loop steals issue slots from the producer warps that share the
// SM sub-partition (profiles/r01g: 21% of the samples of bt_ws_kernel sat in such a loop).
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    if ((++spins & 1023u) == 0u) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 3000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate.
// Warp-uniform issue: ALL 32 lanes of the MMA warp execute the surrounding loop (so every operand is provably
// warp-uniform and lives in uniform registers) and elect.sync picks the one lane that issues.  Issuing from inside
// an `if (lane == 0)` branch instead makes ptxas wrap every tcgen05.mma in a R2UR.BROADCAST "waterfall" loop:
// ~125 clocks per instruction measured (tools/direct_probe.py) against the 32-64 clocks the MMA itself takes.
// The four K=16 steps of one 64-wide k-block in ONE asm block: a single elect.sync, descriptors advanced by +32 bytes
// (+2 in the start-address field) between the steps.  Keeps the MMA warp's instruction stream short -- it shares an
// SM sub-partition scheduler with producer / epilogue warps (34 instructions per MMA before this, profiles/r01h).
__device__ __forceinline__ void umma_bf16_elect_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                   uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %5, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pa;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// Same issue pattern for fp32 parameters: kind::tf32 (fp32 words in shared memory, the tensor core reads the upper
// 19 bits; the producers round to nearest with cvt.rna.tf32.f32 first).  One MMA covers K = 8 (32 bytes), so the four
// steps below walk one 128-byte swizzle row = 32 k -- the same +32-byte descriptor advance as the bf16 form.
__device__ __forceinline__ void umma_tf32_elect_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                   uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %5, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pa;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// The same four steps issued by ONE thread: the caller runs its whole MMA loop under `if (bt_elect_one())`, so there is
// no elect.sync / vote / predicate shuffling per call (~45 uniform-datapath instructions per k-block with the elected
// form, profiles/r02j).  What bounds the loop then is the instruction itself: tools/probes/mma_probe.cu measures
// 43 + N/2 clocks per tcgen05.mma with both operands in shared memory (M = 128: N = 32: 59, 64: 75, 128: 107, 256: 171;
// the same for kind::f16 K = 16 and kind::tf32 K = 8, aligned or unaligned A start rows) -- the A-operand read is not
// hidden behind the math, so a 64-column tile runs the tensor pipe at 43% at best.
template <bool TF32>
__device__ __forceinline__ void umma1_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                         uint32_t accumulate_first) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred pa, pt;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 pa, %5, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pa;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pa, pt;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 pa, %5, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pa;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
        "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
        : "memory");
  }
}
// thread-block cluster helpers (A-operand multicast of the streaming kernel)
__device__ __forceinline__ uint32_t bt_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void bt_cluster_sync() {     // every thread of every CTA of the cluster, converged warps
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t bt_cluster_nctaid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(r));
  return r;
}
// split cluster barrier: every thread of the cluster arrives ONCE and waits ONCE (warps converged: .aligned)
__device__ __forceinline__ void bt_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void bt_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 16-byte store to the same shared-memory offset of CTA `rank` of the cluster (distributed shared memory)
__device__ __forceinline__ void bt_sts16_cluster(uint32_t addr, uint32_t rank, const uint4& v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(rank));
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// commit of the issuing thread's MMAs arriving on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma1_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// true in exactly one lane of a converged warp.  `if (bt_elect_one()) { ...MMA loop... }` is the form ptxas understands:
// inside it the UTCHMMAs are emitted back to back; under `if (lane == 0)` every single MMA is wrapped in an
// ELECT / BRA.U.ANY loop over the "active lanes" (measured with tools/probes/mma_probe.cu: 97 instead of 75 clocks per
// 128x64x16 MMA).
__device__ __forceinline__ bool bt_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma1_commit(uint32_t bar) {     // by the thread that issued the MMAs
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_elect_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate_first) {
  if constexpr (TF32) umma_tf32_elect_x4(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate_first);
  else umma_bf16_elect_x4(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate_first);
}
// round-to-nearest (ties away) fp32 -> tf32, result as an fp32 bit pattern with 13 zero low bits
__device__ __forceinline__ uint32_t bt_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void sts4(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//  [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64: 1024 B
//  between 8-row groups) | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (bit4), a=b=bf16 (1<<7, 1<<10),
// K-major both, N>>3 at [17,23), M>>4 at [24,29)
// (tf32: a = b = 2 in the same fields)
__device__ __forceinline__ uint32_t make_idesc(int n, bool tf32 = false) {
  const uint32_t fmt = tf32 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ uint4 ldg16(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
// packed fp32 pairs (FFMA2 / FADD2: one issue slot for two lanes of the epilogue's affine) and 16-byte shared loads
__device__ __forceinline__ uint64_t bt_pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t bt_pk2u(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void bt_upk2(uint64_t v, uint32_t& a, uint32_t& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t bt_ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t bt_fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts2(uint32_t addr, uint16_t a) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(a) : "memory");
}

// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills the destination (padding taps)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// 8 sign bits (bit j -> element j) -> xor masks for 4 packed bf16x2 words
__device__ __forceinline__ uint4 sign_masks8(uint32_t bits) {
  uint4 m;
  m.x = ((bits & 1u) << 15) | ((bits & 2u) << 30);
  m.y = ((bits & 4u) << 13) | ((bits & 8u) << 28);
  m.z = ((bits & 16u) << 11) | ((bits & 32u) << 26);
  m.w = ((bits & 64u) << 9) | ((bits & 128u) << 24);
  return m;
}

struct TapCoord {
  int dz, dy, dx, lin;
};
__device__ __forceinline__ TapCoord decode_tap(const FusedParams& p, int tap_i) {
  int kd, kh, kw;
  if (p.taps_explicit) {
    const uint32_t t = p.taps[tap_i];
    kd = t & 0xff;
    kh = (t >> 8) & 0xff;
    kw = (t >> 16) & 0xff;
  } else {
    kw = tap_i % p.KW;
    const int r = tap_i / p.KW;
    kh = r % p.KH;
    kd = r / p.KH;
  }
  TapCoord c;
  c.dz = kd * p.dd;
  c.dy = kh * p.dh;
  c.dx = kw * p.dw;
  c.lin = (kd * p.KH + kh) * p.KW + kw;
  return c;
}


template <int WQ>
__device__ __forceinline__ void philox_multi(uint32_t (&c)[WQ][4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#pragma unroll
    for (int u = 0; u < WQ; ++u) {  // WQ independent chains, interleaved by the scheduler
      const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[u][0];
      const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[u][2];
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[u][1] ^ k0;
      const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[u][3] ^ k1;
      c[u][1] = (uint32_t)p1;
      c[u][3] = (uint32_t)p0;
      c[u][0] = n0;
      c[u][2] = n2;
    }
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}


}  // namespace

// entry points of bt_tma.cu (the TMA kernel families are compiled in their own translation unit); `params` points to a
// TmaParams (kernel 0 = bt_tma_kernel, 1 = bt_tms_kernel) or a DtParams (kernel 2 = bt_dtma_kernel) of bt_tma.cuh
int bt_tma_family_launch(int kernel, const void* params, int bn, int tf32, int flip, unsigned gx, unsigned gy, unsigned gz,
                         int smem_bytes, int dev, void* stream);
int bt_tma_probe_launch(const void* params, long long m0, int sample, int group, int tap, int slab, void* out, void* stream);
int bt_tma_probe4d_launch(const void* map, int c, int w, int h, int n, uint32_t dst_off, uint32_t bytes, void* out, void* stream);
