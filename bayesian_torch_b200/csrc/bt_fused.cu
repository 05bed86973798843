// K2/K3: ONE fused sm_100a kernel per Bayesian layer forward (Linear and Conv1d/2d/3d,
// Reparameterization and Flipout).
//
// Reference op sequences it replaces (/root/reference/bayesian_torch/layers/...):
//   variational_layers/linear_variational.py:157-201   exp, log1p, normal_, mul, add, F.linear (+KL)
//   variational_layers/conv_variational.py:183-227 / 357-402 / 530-574     ... F.convNd
//   flipout_layers/linear_flipout.py:145-197           2x F.linear, uniform_().sign() x2, ...
//   flipout_layers/conv_flipout.py:175-244 / 370-439 / 568-637             2x F.convNd ...
//
// Design (B200-first):
//   * implicit GEMM  out[m, n] = sum_k A[m, k] * W[n, k],  m = (image, od, oh, ow) of a
//     channels-last activation, k = (tap, channel) of a channels-last weight, so both operands
//     are K-major and the weight row of the reference's [Cout, Cin, k...] parameter is contiguous.
//   * W is never materialised: 8 producer warps stream mu/rho with 16-byte loads, draw eps with
//     Philox4x32-10 + Box-Muller in registers, form  W = mu + softplus(rho) * eps  (Flipout: the
//     pair  mu , softplus(rho) * eps) and write bf16 straight into the 128B-swizzled K-major
//     shared-memory layout that tcgen05.mma consumes (fence.proxy.async + mbarrier hand-off).
//   * the same warps gather the activation tile (im2col on the fly, zero fill at the borders,
//     optional fp32->bf16 conversion, Flipout input signs applied in registers).
//   * one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) into TMEM; up to 4
//     M-subtiles (4 x 128 rows) share every sampled weight tile, i.e. the Philox/softplus work
//     is amortised over up to 512 output rows; an MC-sample index is a grid dimension, so one
//     launch evaluates S independent weight samples.
//   * epilogue: tcgen05.ld TMEM -> registers, sampled bias, Flipout combine with on-chip output
//     signs, store.  Optional KL side output (warp-reduced per CTA, deterministic finalize).
#include "bt_kernels.cuh"

namespace {

// ------------------------------------------------------------------ the kernel
// Template axes:
//   BLOCK_N  64 | 128           output columns per CTA (UMMA N)
//   FLIP     Reparameterization | Flipout (two operand pairs, two accumulators per M-subtile)
//   NPW      producer warps: 8 (generic path) or 16 (fast path)
//   FAST     branch-free, interleaved sampler; requires vectorisable weights and activations, no debug
//            import hooks and no KL side output (those launches take the generic instantiation)
//   P_BF16 / X_BF16   compile-time dtypes of the fast path (the generic path reads them from FusedParams)
//   TF32     generic path only: fp32 parameters AND fp32 activations -> fp32 words in shared memory (rounded to tf32),
//            tcgen05.mma kind::tf32, 32 k per k-block (one 128-byte swizzle row)
template <int BLOCK_N, bool FLIP, int NPW, bool FAST, bool P_BF16, bool X_BF16, int FMT, bool TF32 = false>
__global__ void __launch_bounds__(NPW * 32 + 32, 1) bt_fused_kernel(const __grid_constant__ FusedParams p) {
  static_assert(!(TF32 && FAST), "the tf32 instantiation exists for the generic path only");
  constexpr int NB = FLIP ? 2 : 1;
  constexpr int KB = TF32 ? 32 : BLOCK_K;   // k per k-block
  constexpr int B_TILE_BYTES = BLOCK_N * 128;
  constexpr int NPT = NPW * 32;  // producer threads

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int CMT = FMT > 0 ? FMT : MAX_MT;   // fast path: M-subtiles per group are a compile-time constant
  const int MT = FMT > 0 ? FMT : p.MT;
  const bool ws = FAST && p.ws != 0;
  // stage of the ring: [sampled B tile(s)][MT activation tile(s)]; weight-stationary: activations only, the
  // sampled tiles of ALL k-blocks live in front of the ring
  const int stage_bytes = ws ? NB * MT * A_TILE_BYTES : NB * (B_TILE_BYTES + MT * A_TILE_BYTES);
  const int a_off = ws ? 0 : NB * B_TILE_BYTES;
  const int res_bytes = ws ? p.num_kb * NB * B_TILE_BYTES : 0;
  const bool x_bf16 = FAST ? X_BF16 : (p.x_is_bf16 != 0);
  const bool p_bf16 = FAST ? P_BF16 : (p.p_is_bf16 != 0);

  uint8_t* aux = smem + res_bytes + p.stages * stage_bytes;
  int4* row_info = reinterpret_cast<int4*>(aux);                              // MAX_MT*128 * 16 B
  float* bias_s = reinterpret_cast<float*>(aux + MAX_MT * BLOCK_M * 16);       // [4][128]: bias0, bias1, scale, shift
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux + MAX_MT * BLOCK_M * 16 + 2048);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 3);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);                        // [16]

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full_bar0 = smem_u32(bars);
  const uint32_t empty_bar0 = smem_u32(bars + MAX_STAGES);
  const uint32_t acc_bar = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t bready_bar = smem_u32(bars + 2 * MAX_STAGES + 1);   // ws: sampled tiles are in smem
  const uint32_t tfree_bar = smem_u32(bars + 2 * MAX_STAGES + 2);    // ws: epilogue has drained the accumulators
  const uint32_t ring_base = smem_u32(smem) + res_bytes;

  const int s = blockIdx.z;
  const int g = blockIdx.y / p.n_tiles_per_group;
  const int n0 = (blockIdx.y % p.n_tiles_per_group) * BLOCK_N;  // first output column inside the group
  // M-groups (MT * 128 rows each) handled by this CTA: exactly one, or a strided set when weight-stationary
  const long long g_first = blockIdx.x;
  const long long g_step = ws ? (long long)gridDim.x : (1ll << 40);
  const long long g_end = ws ? (long long)p.n_groups : g_first + 1;
  const uint32_t sample = p.sample0 + (uint32_t)s + (p.sample_ptr != nullptr ? __ldg(p.sample_ptr) : 0u);
  const int img_base = p.x_shared ? 0 : s * p.B;
  const long long out_sp = (long long)p.OD * p.OH * p.OW;
  const long long in_sp = (long long)p.ID * p.IH * p.IW;
  const bool do_kl = !FAST && (p.kl_partials != nullptr) && blockIdx.x == 0 && blockIdx.z == 0;

  // per-row gather metadata of one group of MT*128 output rows (producer threads only)
  auto fill_rows = [&](long long m0) {
    for (int r = tid; r < MT * BLOCK_M; r += NPT) {
      const long long m = m0 + r;
      int4 info = FAST ? make_int4(0, 0, 0, 0) : make_int4(-1, 0, 0, 0);
      if (m < p.M) {
        int b, od, oh, ow;
        if constexpr (FAST) {  // the host guarantees M < 2^31 on the fast path: 32-bit divisions
          const uint32_t mm = (uint32_t)m, sp = (uint32_t)out_sp, hw = (uint32_t)(p.OH * p.OW);
          b = (int)(mm / sp);
          uint32_t rem = mm - (uint32_t)b * sp;
          od = (int)(rem / hw);
          rem -= (uint32_t)od * hw;
          oh = (int)(rem / (uint32_t)p.OW);
          ow = (int)(rem - (uint32_t)oh * (uint32_t)p.OW);
        } else {
          const long long bb = m / out_sp;
          long long rem = m - bb * out_sp;
          od = (int)(rem / ((long long)p.OH * p.OW));
          rem -= (long long)od * p.OH * p.OW;
          oh = (int)(rem / p.OW);
          ow = (int)(rem - (long long)oh * p.OW);
          b = (int)bb;
        }
        // (transposed: the row carries o + pad; the gather divides (o + pad - k*dil) by the stride)
        const int z0 = p.transposed ? od + p.pd : od * p.sd - p.pd, y0 = p.transposed ? oh + p.ph : oh * p.sh - p.ph,
                  x0 = p.transposed ? ow + p.pw : ow * p.sw - p.pw;
        if constexpr (FAST) {
          // pixel index of the window origin (mod 2^32; the host guarantees < 2^32 pixels) and a bit mask of the
          // filter taps (in iteration order, <= 64) that fall inside the image for this output position
          const long long pix0 = (((long long)(img_base + b) * p.ID + z0) * p.IH + y0) * p.IW + x0;
          unsigned long long mask = 0ull;
          if (!p.taps_natural) {
            const int n_taps = p.K_used / p.Cin_g;
            for (int t = 0; t < n_taps; ++t) {
              const uint32_t tp = p.taps[t];
              const int kd = tp & 0xff, kh = (tp >> 8) & 0xff, kw = (tp >> 16) & 0xff;
              const bool inb = (unsigned)(z0 + kd * p.dd) < (unsigned)p.ID &&
                               (unsigned)(y0 + kh * p.dh) < (unsigned)p.IH &&
                               (unsigned)(x0 + kw * p.dw) < (unsigned)p.IW;
              mask |= (unsigned long long)(inb ? 1 : 0) << t;
            }
          } else {  // all taps in (kd, kh, kw) order: the mask is separable per dimension
            unsigned long long xm = 0ull;
            for (int kw = 0; kw < p.KW; ++kw)
              xm |= (unsigned long long)((unsigned)(x0 + kw * p.dw) < (unsigned)p.IW ? 1 : 0) << kw;
            for (int kd = 0; kd < p.KD; ++kd) {
              if ((unsigned)(z0 + kd * p.dd) >= (unsigned)p.ID) continue;
              for (int kh = 0; kh < p.KH; ++kh)
                if ((unsigned)(y0 + kh * p.dh) < (unsigned)p.IH) mask |= xm << ((kd * p.KH + kh) * p.KW);
            }
          }
          info = make_int4((int)(uint32_t)pix0, (int)(uint32_t)mask, (int)(uint32_t)(mask >> 32), 1);
        } else {
          info = make_int4(img_base + b, z0, y0, x0);
        }
      }
      row_info[r] = info;
    }
  };
  // ---------------------------------------------------------------- setup
  if (warp == NPW) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(full_bar0 + 8 * i, NPW);
        mbar_init(empty_bar0 + 8 * i, 1);
      }
      mbar_init(acc_bar, 1);
      mbar_init(bready_bar, NPW);
      mbar_init(tfree_bar, NPW);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  } else {
    fill_rows((long long)g_first * (MT * BLOCK_M));
    if (tid < BLOCK_N) {
      const int n = n0 + tid;
      float b0 = 0.f, b1 = 0.f;
      if (p.mu_b != nullptr && n < p.N) {
        const int ng = g * p.N + n;
        float mu, rho;
        if (p_bf16) {
          mu = __bfloat162float(static_cast<const __nv_bfloat16*>(p.mu_b)[ng]);
          rho = __bfloat162float(static_cast<const __nv_bfloat16*>(p.rho_b)[ng]);
        } else {
          mu = static_cast<const float*>(p.mu_b)[ng];
          rho = static_cast<const float*>(p.rho_b)[ng];
        }
        float eps;
        if (!FAST && p.eps_b_in != nullptr) {
          eps = p.eps_b_in[ng];
        } else {
          const float4 z = bt_eps_quad(p.key, BT_STREAM_B_EPS, (uint32_t)(ng >> 2), 0u, sample);
          const int j = ng & 3;
          eps = j == 0 ? z.x : (j == 1 ? z.y : (j == 2 ? z.z : z.w));
        }
        const float d = bt_softplus(rho) * eps;
        if (FLIP) {
          b0 = mu;
          b1 = d;
        } else {
          b0 = mu + d;
        }
      }
      bias_s[tid] = b0;
      bias_s[128 + tid] = b1;
      float sc = 1.f, sh = 0.f;
      if (p.ep_scale != nullptr && n < p.N) {
        sc = __ldg(p.ep_scale + g * p.N + n);
        sh = __ldg(p.ep_shift + g * p.N + n);
      }
      bias_s[256 + tid] = sc;
      bias_s[384 + tid] = sh;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == NPW) {
    // ============================================================== MMA issuer: the whole warp runs the loop with
    // warp-uniform operands, one elected lane issues (see umma_bf16_elect_x4)
    {
      const uint32_t idesc = make_idesc(BLOCK_N, TF32);
      const uint64_t desc_hi = make_smem_desc(0u);
      int stage = 0;
      uint32_t phase = 0;
      if (ws) {
        mbar_wait(bready_bar, 0);      // every k-block of the sampled tile is resident
        tc_fence_after();
      }
      int it = 0;
      for (long long gi = g_first; gi < g_end; gi += g_step, ++it) {
        if (it > 0) {                  // previous group's accumulators have been read out
          mbar_wait(tfree_bar, (uint32_t)((it - 1) & 1));
          tc_fence_after();
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_idle(full_bar0 + 8 * stage, phase, 32);
          tc_fence_after();
          const uint32_t sst = ring_base + stage * stage_bytes;
          const uint32_t sb16 = ((ws ? smem_base + kb * NB * B_TILE_BYTES : sst) & 0x3FFFFu) >> 4;
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t sa16 = ((sst + a_off + mt * NB * A_TILE_BYTES) & 0x3FFFFu) >> 4;
            umma_elect_x4<TF32>(tmem_base + (uint32_t)(mt * NB * BLOCK_N), sa16 | (1u << 16), sb16 | (1u << 16),
                                (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
            if (FLIP)
              umma_elect_x4<TF32>(tmem_base + (uint32_t)((mt * NB + 1) * BLOCK_N), (sa16 + (A_TILE_BYTES >> 4)) | (1u << 16),
                                  (sb16 + (B_TILE_BYTES >> 4)) | (1u << 16), (uint32_t)(desc_hi >> 32), idesc,
                                  kb != 0 ? 1u : 0u);
          }
          umma_commit_elect(empty_bar0 + 8 * stage);  // frees this stage's smem once the MMAs have read it
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_elect(acc_bar);  // this group's accumulators are complete
      }
    }
    __syncwarp();
  } else {
    // ============================================================== producers
    float kl_acc = 0.f;
    const uint8_t* mu_w = static_cast<const uint8_t*>(p.mu_w);
    const uint8_t* rho_w = static_cast<const uint8_t*>(p.rho_w);
    const uint8_t* xb = static_cast<const uint8_t*>(p.x);

    // ---------------------------------------------------------------- epilogue (all producer warps)
    auto epilogue = [&](long long m0, uint32_t acc_parity) {
    mbar_wait_idle(acc_bar, acc_parity, 128);
    tc_fence_after();
    constexpr int PARTS = NPW / 4;                       // column slices (warps sharing a TMEM lane quarter)
    constexpr int COLS_PER_WARP = BLOCK_N / PARTS;
    const int q = warp & 3, part = warp >> 2;
    const int o_es = x_bf16 ? 2 : 4;
    uint8_t* outb = static_cast<uint8_t*>(p.out);
    for (int mt = 0; mt < MT; ++mt) {
      const int rl = q * 32 + lane;
      const long long m = m0 + (long long)mt * BLOCK_M + rl;
      const bool mvalid = m < p.M;
      const long long orow = (long long)s * p.M + m;
      uint4 sblk = make_uint4(0u, 0u, 0u, 0u);
      if (FLIP && (FAST || p.sign_out == nullptr))
        sblk = bt_sign_block(p.key, BT_STREAM_SIGN_OUT, ((uint32_t)g << 20) | (uint32_t)(n0 >> 7),
                             (uint32_t)m, sample);
#pragma unroll 1
      for (int cc = 0; cc < COLS_PER_WARP; cc += 16) {
        const int col0 = part * COLS_PER_WARP + cc;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NB * BLOCK_N + col0);
        uint32_t v0[16], v1[16];
        tmem_ld16(taddr, v0);
        if (FLIP) tmem_ld16(taddr + BLOCK_N, v1);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = col0 + j;
          float val = __uint_as_float(v0[j]) + bias_s[col];
          if (FLIP) {
            float pert = __uint_as_float(v1[j]) + bias_s[128 + col];
            bool neg;
            if (!FAST && p.sign_out != nullptr) {
              const int n = n0 + col;
              neg = (mvalid && n < p.N) ? (__ldg(p.sign_out + orow * p.C_out + g * p.N + n) < 0.f) : false;
            } else {
              const int bit = (n0 & 127) + col;
              neg = (bt_sign_word(sblk, bit >> 5) >> (bit & 31)) & 1u;
            }
            val += neg ? -pert : pert;
          }
          if (p.ep_scale != nullptr) val = fmaf(val, bias_s[256 + col], bias_s[384 + col]);
          o[j] = val;
        }
        if (mvalid) {
          const int nfirst = n0 + col0;
          const long long eoff = orow * p.C_out + g * p.N + nfirst;
          uint8_t* dst = outb + eoff * o_es;
          const bool vec_ok = p.out_vec && nfirst + 16 <= p.N;
          if (p.ep_residual != nullptr) {
            const uint8_t* rsd = static_cast<const uint8_t*>(p.ep_residual) + eoff * o_es;
            if (vec_ok) {
              if (x_bf16) {
                const uint4 a = ldg16(rsd), b = ldg16(rsd + 16);
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  o[2 * j] += bt_bf16_lo(w[j]);
                  o[2 * j + 1] += bt_bf16_hi(w[j]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 r = __ldg(reinterpret_cast<const float4*>(rsd) + j);
                  o[4 * j] += r.x; o[4 * j + 1] += r.y; o[4 * j + 2] += r.z; o[4 * j + 3] += r.w;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (nfirst + j < p.N)
                  o[j] += x_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rsd)[j])
                                 : reinterpret_cast<const float*>(rsd)[j];
              }
            }
          }
          if (p.ep_relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
          }
          if (vec_ok) {
            if (x_bf16) {
              uint4 a, b;
              a.x = bt_pack_bf16x2(o[0], o[1]);   a.y = bt_pack_bf16x2(o[2], o[3]);
              a.z = bt_pack_bf16x2(o[4], o[5]);   a.w = bt_pack_bf16x2(o[6], o[7]);
              b.x = bt_pack_bf16x2(o[8], o[9]);   b.y = bt_pack_bf16x2(o[10], o[11]);
              b.z = bt_pack_bf16x2(o[12], o[13]); b.w = bt_pack_bf16x2(o[14], o[15]);
              reinterpret_cast<uint4*>(dst)[0] = a;
              reinterpret_cast<uint4*>(dst)[1] = b;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                reinterpret_cast<float4*>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (nfirst + j < p.N) {
                if (x_bf16) reinterpret_cast<__nv_bfloat16*>(dst)[j] = __float2bfloat16_rn(o[j]);
                else reinterpret_cast<float*>(dst)[j] = o[j];
              }
            }
          }
        }
      }
    }
    };


    if constexpr (FAST) {
      // ------------------------------------------------------------ fast path (16 warps)
      constexpr int WO = BLOCK_N / 64;          // weight "octs" (8 consecutive k = one Philox call) per thread
      constexpr int AT = 2;                     // activation chunks per thread and subtile: rows arb + 64*i
      const int wo = tid & 7, wrb = tid >> 3;   // oct `wo` of rows wrb + 64*i
      const int ac = tid & 7, arb = tid >> 3;
      constexpr int P_ES = P_BF16 ? 2 : 4;
      constexpr int PW = P_BF16 ? 4 : 8;        // 32-bit words per oct of parameters

      // weight oct registers of the CURRENT k-block (raw words)
      uint32_t mu_r[WO][PW], rho_r[WO][PW];
      uint32_t ko_cur = 0;
      bool kvalid_cur = false;
      long long row_off[WO];
      bool nvalid[WO];
#pragma unroll
      for (int i = 0; i < WO; ++i) {
        const int n = n0 + wrb + 64 * i;
        nvalid[i] = n < p.N;
        row_off[i] = ((long long)g * p.N + (nvalid[i] ? n : p.N - 1)) * p.K_phys;
      }
      // (tap, channel) cursor of this thread's weight oct; advanced by one k-block per load_weights call
      int w_tap = (wo * 8) / p.Cin_g, w_c = (wo * 8) - ((wo * 8) / p.Cin_g) * p.Cin_g;
      auto load_weights = [&](int kb) {
        const int ku0 = kb * BLOCK_K + wo * 8;
        kvalid_cur = ku0 < p.K_used;
        long long kphys0 = 0;
        if (kvalid_cur) kphys0 = (long long)decode_tap(p, w_tap).lin * p.Cin_g + w_c;
        w_tap += p.q64;
        w_c += p.r64;
        if (w_c >= p.Cin_g) {
          w_c -= p.Cin_g;
          ++w_tap;
        }
        ko_cur = (uint32_t)(kphys0 >> 3);
#pragma unroll
        for (int i = 0; i < WO; ++i) {
          const long long off = (row_off[i] + kphys0) * P_ES;
          const uint4 a = ldg16(mu_w + off);
          const uint4 b = ldg16(rho_w + off);
          mu_r[i][0] = a.x; mu_r[i][1] = a.y; mu_r[i][2] = a.z; mu_r[i][3] = a.w;
          rho_r[i][0] = b.x; rho_r[i][1] = b.y; rho_r[i][2] = b.z; rho_r[i][3] = b.w;
          if constexpr (!P_BF16) {
            const uint4 a2 = ldg16(mu_w + off + 16);
            const uint4 b2 = ldg16(rho_w + off + 16);
            mu_r[i][4] = a2.x; mu_r[i][5] = a2.y; mu_r[i][6] = a2.z; mu_r[i][7] = a2.w;
            rho_r[i][4] = b2.x; rho_r[i][5] = b2.y; rho_r[i][6] = b2.z; rho_r[i][7] = b2.w;
          }
        }
      };
      load_weights(0);

      // ---- sample one [BLOCK_N x 64] weight tile (from the octs in mu_r / rho_r) into smem at `sb`:
      //      one Philox call -> 8 normals -> 8 weights -> one 16-byte swizzled store; no branches
      auto sample_tile = [&](uint32_t sb) {
        uint32_t c[WO][4];
#pragma unroll
        for (int i = 0; i < WO; ++i) {
          c[i][0] = ko_cur;
          c[i][1] = (uint32_t)(g * p.N + n0 + wrb + 64 * i);
          c[i][2] = sample;
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
              m8[2 * j] = bt_bf16_lo(mu_r[i][j]);
              m8[2 * j + 1] = bt_bf16_hi(mu_r[i][j]);
              r8[2 * j] = bt_bf16_lo(rho_r[i][j]);
              r8[2 * j + 1] = bt_bf16_hi(rho_r[i][j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              m8[j] = __uint_as_float(mu_r[i][j]);
              r8[j] = __uint_as_float(rho_r[i][j]);
            }
          }
          const bool ok = kvalid_cur && nvalid[i];
          float w0[8], w1[8];
          if (!p.rho_is_sigma) {   // (warp-uniform) sigma = softplus(rho); skipped when the caller cached sigma
#pragma unroll
            for (int j = 0; j < 8; ++j) r8[j] = bt_softplus_fast(r8[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float sg = r8[j];
            if (FLIP) {
              w0[j] = ok ? m8[j] : 0.f;
              w1[j] = ok ? sg * e[j] : 0.f;
            } else {
              w0[j] = ok ? fmaf(sg, e[j], m8[j]) : 0.f;
            }
          }
          const int nl = wrb + 64 * i;
          const uint32_t soff = (uint32_t)(nl * 128 + ((wo ^ (nl & 7)) << 4));
          sts16(sb + soff, make_uint4(bt_pack_bf16x2(w0[0], w0[1]), bt_pack_bf16x2(w0[2], w0[3]),
                                      bt_pack_bf16x2(w0[4], w0[5]), bt_pack_bf16x2(w0[6], w0[7])));
          if (FLIP)
            sts16(sb + B_TILE_BYTES + soff,
                  make_uint4(bt_pack_bf16x2(w1[0], w1[1]), bt_pack_bf16x2(w1[2], w1[3]),
                             bt_pack_bf16x2(w1[4], w1[5]), bt_pack_bf16x2(w1[6], w1[7])));
        }
      };

      const uint32_t pix_bytes = (uint32_t)p.C_in * (X_BF16 ? 2u : 4u);
      const uint32_t base_pix = (uint32_t)((long long)img_base * in_sp);
      int stage = 0;
      uint32_t phase = 0;
      // gather metadata of this thread's rows (window-origin pixel + tap mask), held in registers per M-group
      uint32_t rpix[CMT][AT], rmlo[CMT][AT], rmhi[CMT][AT];
      auto load_rows = [&]() {
#pragma unroll
        for (int mt = 0; mt < CMT; ++mt) {
#pragma unroll
          for (int i = 0; i < AT; ++i) {
            rpix[mt][i] = rmlo[mt][i] = rmhi[mt][i] = 0u;
            if (mt < MT) {
              const int4 info = row_info[mt * BLOCK_M + arb + 64 * i];
              rpix[mt][i] = (uint32_t)info.x;
              rmlo[mt][i] = (uint32_t)info.y;
              rmhi[mt][i] = (uint32_t)info.z;
            }
          }
        }
      };
      load_rows();

      // (tap, channel) cursor of this thread's activation chunk; reset at the start of every M-group
      const int a_tap0 = (ac * 8) / p.Cin_g, a_c0 = (ac * 8) - ((ac * 8) / p.Cin_g) * p.Cin_g;
      int a_tap = a_tap0, a_c = a_c0;
      // ---- one k-block of the ring: gather MT activation tiles (and, unless weight-stationary, sample the
      //      weight tile) into stage `stage`, then hand it to the tensor core
      auto produce_stage = [&](int kb) {
        // activation chunk geometry of this k-block: tap -> pixel delta + bit in the row's tap mask
        const int ku = kb * BLOCK_K + ac * 8;
        const bool kv = ku < p.K_used;
        int cg = 0, tap_i = 0;
        uint32_t dpix = 0;
        if (kv) {
          tap_i = a_tap;
          const TapCoord tc = decode_tap(p, tap_i);
          cg = g * p.Cin_g + a_c;
          dpix = (uint32_t)((tc.dz * p.IH + tc.dy) * p.IW + tc.dx);
        }
        a_tap += p.q64;
        a_c += p.r64;
        if (a_c >= p.Cin_g) {
          a_c -= p.Cin_g;
          ++a_tap;
        }
        const uint8_t* xcol = xb + (size_t)cg * (X_BF16 ? 2 : 4);
        // 1. issue the activation loads (bf16 activations: all subtiles in flight while we sample)
        uint4 va[CMT][AT];
        uint32_t prow[CMT][AT];
        bool oka[CMT][AT];
        if constexpr (X_BF16) {
#pragma unroll
          for (int mt = 0; mt < CMT; ++mt) {
            if (mt < MT) {
#pragma unroll
              for (int i = 0; i < AT; ++i) {
                const uint32_t mword = tap_i < 32 ? rmlo[mt][i] : rmhi[mt][i];
                oka[mt][i] = kv && ((mword >> (tap_i & 31)) & 1u);
                const uint32_t pix = rpix[mt][i] + dpix;
                prow[mt][i] = pix - base_pix;
                va[mt][i] = make_uint4(0u, 0u, 0u, 0u);
                if (oka[mt][i]) va[mt][i] = ldg16(xcol + (unsigned long long)pix * pix_bytes);
              }
            }
          }
        }
        // 2. wait until the tensor core has drained this stage's buffers
        mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
        const uint32_t sst = ring_base + stage * stage_bytes;
        // 3. sample the weight tile, then prefetch the next k-block's mu / rho
        if (!ws) {
          sample_tile(sst);
          if (kb + 1 < p.num_kb) load_weights(kb + 1);
        }
        // 4. activation tiles -> swizzled smem (+ Flipout sign-flipped copy)
#pragma unroll
        for (int mt = 0; mt < CMT; ++mt) {
          if (mt < MT) {
            const uint32_t sa = sst + a_off + mt * NB * A_TILE_BYTES;
#pragma unroll
            for (int i = 0; i < AT; ++i) {
              const int rl = arb + 64 * i;
              uint4 v;
              bool ok;
              uint32_t pr;
              if constexpr (X_BF16) {
                v = va[mt][i];
                ok = oka[mt][i];
                pr = prow[mt][i];
              } else {
                const uint32_t mword = tap_i < 32 ? rmlo[mt][i] : rmhi[mt][i];
                ok = kv && ((mword >> (tap_i & 31)) & 1u);
                const uint32_t pix = rpix[mt][i] + dpix;
                pr = pix - base_pix;
                v = make_uint4(0u, 0u, 0u, 0u);
                if (ok) {
                  const float4* src = reinterpret_cast<const float4*>(xcol + (unsigned long long)pix * pix_bytes);
                  const float4 a = __ldg(src), b = __ldg(src + 1);
                  v.x = bt_pack_bf16x2(a.x, a.y);
                  v.y = bt_pack_bf16x2(a.z, a.w);
                  v.z = bt_pack_bf16x2(b.x, b.y);
                  v.w = bt_pack_bf16x2(b.z, b.w);
                }
              }
              const uint32_t soff = (uint32_t)(rl * 128 + ((ac ^ (rl & 7)) << 4));
              sts16(sa + soff, v);
              if (FLIP) {
                const uint4 blk = bt_sign_block(p.key, BT_STREAM_SIGN_IN, (uint32_t)(cg >> 7), pr, sample);
                const uint32_t bits = ok ? ((bt_sign_word(blk, (cg & 127) >> 5) >> (cg & 31)) & 0xffu) : 0u;
                const uint4 mk = sign_masks8(bits);
                v.x ^= mk.x; v.y ^= mk.y; v.z ^= mk.z; v.w ^= mk.w;
                sts16(sa + A_TILE_BYTES + soff, v);
              }
            }
          }
        }
        // 5. publish the stage to the tensor core
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar0 + 8 * stage);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      };

      // ---- weight-stationary, bf16 activations, no Flipout: the gather is a multi-stage cp.async (LDGSTS)
      //      pipeline -- a warp keeps up to ASYNC_DEPTH k-blocks of 16-byte copies in flight and publishes a
      //      stage only when its copies have landed, so the L2/HBM latency of the im2col reads is overlapped
      //      instead of being paid once per k-block (profiles/r01d: 45% long-scoreboard stalls before this)
      constexpr bool ASYNC_OK = X_BF16 && !FLIP;
      constexpr int ASYNC_DEPTH = 4;
      int arr_stage = 0;  // next stage to publish (the issue cursor is `stage` / `phase`)
      auto issue_stage_async = [&](int kb) {
        const int ku = kb * BLOCK_K + ac * 8;
        const bool kv = ku < p.K_used;
        int cg = 0, tap_i = 0;
        uint32_t dpix = 0;
        if (kv) {
          tap_i = a_tap;
          const TapCoord tc = decode_tap(p, tap_i);
          cg = g * p.Cin_g + a_c;
          dpix = (uint32_t)((tc.dz * p.IH + tc.dy) * p.IW + tc.dx);
        }
        a_tap += p.q64;
        a_c += p.r64;
        if (a_c >= p.Cin_g) {
          a_c -= p.Cin_g;
          ++a_tap;
        }
        const uint8_t* xcol = xb + (size_t)cg * 2;
        mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
        const uint32_t sst = ring_base + stage * stage_bytes;
#pragma unroll
        for (int mt = 0; mt < CMT; ++mt) {
          if (mt < MT) {
#pragma unroll
            for (int i = 0; i < AT; ++i) {
              const int rl = arb + 64 * i;
              const uint32_t mword = tap_i < 32 ? rmlo[mt][i] : rmhi[mt][i];
              const bool ok = kv && ((mword >> (tap_i & 31)) & 1u);
              const uint32_t pix = ok ? rpix[mt][i] + dpix : 0u;
              cp_async16(sst + mt * A_TILE_BYTES + (uint32_t)(rl * 128 + ((ac ^ (rl & 7)) << 4)),
                         xcol + (unsigned long long)pix * pix_bytes, ok ? 16u : 0u);
            }
          }
        }
        cp_async_commit();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto publish_stage = [&]() {   // the oldest in-flight stage of this warp has landed
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar0 + 8 * arr_stage);
        if (++arr_stage == p.stages) arr_stage = 0;
      };

      if (!ws) {
        for (int kb = 0; kb < p.num_kb; ++kb) produce_stage(kb);
        epilogue((long long)g_first * (MT * BLOCK_M), 0u);
      } else {
        // ---- weight-stationary: sample every k-block of this (n-tile, sample) ONCE ...
        for (int kb = 0; kb < p.num_kb; ++kb) {
          sample_tile(smem_base + kb * NB * B_TILE_BYTES);
          if (kb + 1 < p.num_kb) load_weights(kb + 1);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bready_bar);
        // ---- ... then stream the activation tiles of all the CTA's M-groups past it
        int it = 0;
        for (long long gi = g_first; gi < g_end; gi += g_step, ++it) {
          if (it > 0) {
            fill_rows(gi * (MT * BLOCK_M));
            named_bar_sync(1, NPT);
            load_rows();
          }
          a_tap = a_tap0;
          a_c = a_c0;
          if (ASYNC_OK && p.ws_async && p.stages > ASYNC_DEPTH) {
            int in_flight = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
              issue_stage_async(kb);
              if (++in_flight == ASYNC_DEPTH) {
                cp_async_wait<ASYNC_DEPTH - 1>();
                publish_stage();
                --in_flight;
              }
            }
            cp_async_wait<0>();
            for (; in_flight > 0; --in_flight) publish_stage();
          } else {
            for (int kb = 0; kb < p.num_kb; ++kb) produce_stage(kb);
            arr_stage = stage;
          }
          epilogue(gi * (MT * BLOCK_M), (uint32_t)(it & 1));
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tfree_bar);
          // all epilogue reads of row-independent state are done; the next fill_rows may overwrite row_info only
          // after every warp has left its gather loop, which the accumulator barrier already guarantees
        }
      }
    } else {
      // ------------------------------------------------------------ generic path (8 warps): any shape /
      // alignment, debug import hooks, KL side output; bf16 operands (64 k per k-block) or, for fp32 parameters with
      // fp32 activations, tf32 operands (32 k per k-block)
      constexpr int QPR = KB / 4;            // weight quads (4 consecutive k) per k-block row: 16 | 8
      constexpr int RPP = NPT / QPR;         // weight rows per pass: 16 | 32
      constexpr int CE = TF32 ? 4 : 8;       // activation elements per 16-byte shared-memory chunk
      // weights: quad wq of rows wrb + RPP*i
      const int wq = tid % QPR, wrb = tid / QPR;
      // activations, vector path: 16-byte chunk `ac` (CE channels) of rows arb + 32*i
      const int ac = tid & 7, arb = tid >> 3;
      // activations, scalar path: k column aj of rows asr + (NPT/KB)*i
      const int aj = tid % KB, asr = tid / KB;
      const int x_es = x_bf16 ? 2 : 4;

      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        // ------------------------------------------------ 1. global loads of the weight quads
        constexpr int WQ = BLOCK_N / RPP;  // quads per thread
        float mu[WQ][4], rho[WQ][4];
        const int ku0 = kb * KB + wq * 4;  // index in the (tap-compacted) K
        bool kvalid = ku0 < p.K_used;
        long long kphys0 = ku0;
        if (p.taps_explicit && kvalid) {
          const int tap_i = ku0 / p.Cin_g;
          kphys0 = (long long)decode_tap(p, tap_i).lin * p.Cin_g + (ku0 - tap_i * p.Cin_g);
        }
#pragma unroll
        for (int i = 0; i < WQ; ++i) {
          const int nl = wrb + RPP * i;
          const int n = n0 + nl;
          const bool ok = kvalid && n < p.N;
          const long long off = ((long long)g * p.N + n) * p.K_phys + kphys0;
#pragma unroll
          for (int j = 0; j < 4; ++j) mu[i][j] = rho[i][j] = 0.f;
          if (ok) {
            if (p.w_vec) {
              if (p_bf16) {
                const uint2 a = __ldg(reinterpret_cast<const uint2*>(mu_w + off * 2));
                const uint2 b = __ldg(reinterpret_cast<const uint2*>(rho_w + off * 2));
                mu[i][0] = bt_bf16_lo(a.x); mu[i][1] = bt_bf16_hi(a.x);
                mu[i][2] = bt_bf16_lo(a.y); mu[i][3] = bt_bf16_hi(a.y);
                rho[i][0] = bt_bf16_lo(b.x); rho[i][1] = bt_bf16_hi(b.x);
                rho[i][2] = bt_bf16_lo(b.y); rho[i][3] = bt_bf16_hi(b.y);
              } else {
                const float4 a = __ldg(reinterpret_cast<const float4*>(mu_w + off * 4));
                const float4 b = __ldg(reinterpret_cast<const float4*>(rho_w + off * 4));
                mu[i][0] = a.x; mu[i][1] = a.y; mu[i][2] = a.z; mu[i][3] = a.w;
                rho[i][0] = b.x; rho[i][1] = b.y; rho[i][2] = b.z; rho[i][3] = b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (ku0 + j < p.K_used) {
                  if (p_bf16) {
                    mu[i][j] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(mu_w)[off + j]);
                    rho[i][j] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rho_w)[off + j]);
                  } else {
                    mu[i][j] = reinterpret_cast<const float*>(mu_w)[off + j];
                    rho[i][j] = reinterpret_cast<const float*>(rho_w)[off + j];
                  }
                }
              }
            }
          }
        }

        // ------------------------------------------------ 2. wait for the stage to be free
        mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
        const uint32_t sb = smem_base + stage * stage_bytes;

        // ------------------------------------------------ 3. sample the weight tile
#pragma unroll
        for (int i = 0; i < WQ; ++i) {
          const int nl = wrb + RPP * i;
          const int n = n0 + nl;
          const bool ok = kvalid && n < p.N;
          const uint32_t ng = (uint32_t)(g * p.N + n);
          float w0[4], w1[4];
          if (ok) {
            float e[4];
            if (p.eps_w_in != nullptr) {
              const long long off = (long long)ng * p.K_phys + kphys0;
#pragma unroll
              for (int j = 0; j < 4; ++j) e[j] = (ku0 + j < p.K_used) ? __ldg(p.eps_w_in + off + j) : 0.f;
            } else {
              const float4 z = bt_eps_quad(p.key, BT_STREAM_W_EPS, (uint32_t)(kphys0 >> 2), ng, sample);
              e[0] = z.x; e[1] = z.y; e[2] = z.z; e[3] = z.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool ev = p.w_vec || (ku0 + j < p.K_used);
              const float sg = p.rho_is_sigma ? rho[i][j] : bt_softplus(rho[i][j]);
              if (FLIP) {
                w0[j] = ev ? mu[i][j] : 0.f;
                w1[j] = ev ? sg * e[j] : 0.f;
              } else {
                w0[j] = ev ? fmaf(sg, e[j], mu[i][j]) : 0.f;
              }
              if (do_kl && ev)
                kl_acc += bt_kl_elem(mu[i][j], sg, p.prior_mu, p.log_prior_sigma, p.inv_2ps2);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) w0[j] = w1[j] = 0.f;
          }
          if constexpr (TF32) {
            const uint32_t soff = (uint32_t)(nl * 128 + ((wq ^ (nl & 7)) << 4));
            sts16(sb + soff, make_uint4(bt_tf32(w0[0]), bt_tf32(w0[1]), bt_tf32(w0[2]), bt_tf32(w0[3])));
            if (FLIP)
              sts16(sb + B_TILE_BYTES + soff, make_uint4(bt_tf32(w1[0]), bt_tf32(w1[1]), bt_tf32(w1[2]), bt_tf32(w1[3])));
          } else {
            const uint32_t soff = (uint32_t)(nl * 128 + (((wq >> 1) ^ (nl & 7)) << 4) + ((wq & 1) << 3));
            sts8(sb + soff, bt_pack_bf16x2(w0[0], w0[1]), bt_pack_bf16x2(w0[2], w0[3]));
            if (FLIP)
              sts8(sb + B_TILE_BYTES + soff, bt_pack_bf16x2(w1[0], w1[1]), bt_pack_bf16x2(w1[2], w1[3]));
          }
        }

        // ------------------------------------------------ 4. gather the activation tiles
        if (p.a_vec) {
          const int ku = kb * KB + ac * CE;
          const bool kv = ku < p.K_used;
          TapCoord tc = {0, 0, 0, 0};
          int cg = 0;
          if (kv) {
            const int tap_i = ku / p.Cin_g;
            tc = decode_tap(p, tap_i);
            cg = g * p.Cin_g + (ku - tap_i * p.Cin_g);
          }
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t sa = sb + NB * B_TILE_BYTES + mt * NB * A_TILE_BYTES;
            uint4 v[4];
            long long pix[4];
            bool okr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = arb + 32 * i;
              const int4 info = row_info[mt * BLOCK_M + rl];
              int z, y, xw;
              bool inb;
              if (p.transposed) {   // fractionally-strided gather (ConvTranspose): input = (o + pad - k*dil) / stride
                const int zn = info.y - tc.dz, yn = info.z - tc.dy, xn = info.w - tc.dx;
                z = zn / p.sd; y = yn / p.sh; xw = xn / p.sw;
                inb = zn >= 0 && yn >= 0 && xn >= 0 && z * p.sd == zn && y * p.sh == yn && xw * p.sw == xn &&
                      z < p.ID && y < p.IH && xw < p.IW;
              } else {
                z = info.y + tc.dz; y = info.z + tc.dy; xw = info.w + tc.dx;
                inb = (unsigned)z < (unsigned)p.ID && (unsigned)y < (unsigned)p.IH && (unsigned)xw < (unsigned)p.IW;
              }
              okr[i] = kv && info.x >= 0 && inb;
              pix[i] = (((long long)info.x * p.ID + z) * p.IH + y) * p.IW + xw;
              v[i] = make_uint4(0u, 0u, 0u, 0u);
              if (okr[i]) {
                const uint8_t* src = xb + (pix[i] * p.C_in + cg) * x_es;
                if constexpr (TF32) {
                  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
                  v[i] = make_uint4(bt_tf32(a.x), bt_tf32(a.y), bt_tf32(a.z), bt_tf32(a.w));
                } else if (x_bf16) {
                  v[i] = ldg16(src);
                } else {
                  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
                  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
                  v[i].x = bt_pack_bf16x2(a.x, a.y);
                  v[i].y = bt_pack_bf16x2(a.z, a.w);
                  v[i].z = bt_pack_bf16x2(b.x, b.y);
                  v[i].w = bt_pack_bf16x2(b.z, b.w);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = arb + 32 * i;
              const uint32_t soff = (uint32_t)(rl * 128 + ((ac ^ (rl & 7)) << 4));
              sts16(sa + soff, v[i]);
              if (FLIP) {
                uint4 f = v[i];
                if (okr[i]) {
                  uint32_t bits;
                  const long long spix = pix[i] + (long long)(s * p.B - img_base) * in_sp;  // pixel incl. sample
                  if (p.sign_in != nullptr) {
                    const float* sp_ = p.sign_in + spix * p.C_in + cg;
                    bits = 0;
#pragma unroll
                    for (int j = 0; j < CE; ++j) bits |= (__ldg(sp_ + j) < 0.f ? 1u : 0u) << j;
                  } else {
                    const uint32_t prow = (uint32_t)(pix[i] - (long long)img_base * in_sp);
                    const uint4 blk = bt_sign_block(p.key, BT_STREAM_SIGN_IN, (uint32_t)(cg >> 7), prow, sample);
                    bits = (bt_sign_word(blk, (cg & 127) >> 5) >> (cg & 31)) & (TF32 ? 0xfu : 0xffu);
                  }
                  if constexpr (TF32) {
                    f.x ^= (bits & 1u) << 31; f.y ^= (bits & 2u) << 30; f.z ^= (bits & 4u) << 29; f.w ^= (bits & 8u) << 28;
                  } else {
                    const uint4 mk = sign_masks8(bits);
                    f.x ^= mk.x; f.y ^= mk.y; f.z ^= mk.z; f.w ^= mk.w;
                  }
                }
                sts16(sa + A_TILE_BYTES + soff, f);
              }
            }
          }
        } else {
          // scalar gather: any channel count / alignment (e.g. a Cin=3 stem that was not channel-padded)
          const int k = kb * KB + aj;
          const bool kv = k < p.K_used;
          TapCoord tc = {0, 0, 0, 0};
          int cg = 0;
          if (kv) {
            const int tap_i = k / p.Cin_g;
            tc = decode_tap(p, tap_i);
            cg = g * p.Cin_g + (k - tap_i * p.Cin_g);
          }
          constexpr int RSTEP = NPT / KB;   // rows covered per pass: 4 | 8
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t sa = sb + NB * B_TILE_BYTES + mt * NB * A_TILE_BYTES;
#pragma unroll 4
            for (int i = 0; i < BLOCK_M / RSTEP; ++i) {
              const int rl = asr + RSTEP * i;
              const int4 info = row_info[mt * BLOCK_M + rl];
              int z, y, xw;
              bool inb;
              if (p.transposed) {
                const int zn = info.y - tc.dz, yn = info.z - tc.dy, xn = info.w - tc.dx;
                z = zn / p.sd; y = yn / p.sh; xw = xn / p.sw;
                inb = zn >= 0 && yn >= 0 && xn >= 0 && z * p.sd == zn && y * p.sh == yn && xw * p.sw == xn &&
                      z < p.ID && y < p.IH && xw < p.IW;
              } else {
                z = info.y + tc.dz; y = info.z + tc.dy; xw = info.w + tc.dx;
                inb = (unsigned)z < (unsigned)p.ID && (unsigned)y < (unsigned)p.IH && (unsigned)xw < (unsigned)p.IW;
              }
              const bool ok = kv && info.x >= 0 && inb;
              const long long pix = (((long long)info.x * p.ID + z) * p.IH + y) * p.IW + xw;
              uint32_t h = 0, hf = 0;   // bf16: low 16 bits; tf32: the fp32 word
              if (ok) {
                const long long e = pix * p.C_in + cg;
                if constexpr (TF32) {
                  h = bt_tf32(__ldg(reinterpret_cast<const float*>(xb) + e));
                } else if (x_bf16) {
                  h = __ldg(reinterpret_cast<const uint16_t*>(xb) + e);
                } else {
                  const __nv_bfloat16 t = __float2bfloat16_rn(__ldg(reinterpret_cast<const float*>(xb) + e));
                  h = *reinterpret_cast<const uint16_t*>(&t);
                }
                if (FLIP) {
                  uint32_t neg;
                  if (p.sign_in != nullptr) {
                    const long long spix = pix + (long long)(s * p.B - img_base) * in_sp;
                    neg = __ldg(p.sign_in + spix * p.C_in + cg) < 0.f ? 1u : 0u;
                  } else {
                    const uint32_t prow = (uint32_t)(pix - (long long)img_base * in_sp);
                    const uint4 blk = bt_sign_block(p.key, BT_STREAM_SIGN_IN, (uint32_t)(cg >> 7), prow, sample);
                    neg = (bt_sign_word(blk, (cg & 127) >> 5) >> (cg & 31)) & 1u;
                  }
                  hf = h ^ (neg << (TF32 ? 31 : 15));
                }
              }
              if constexpr (TF32) {
                const uint32_t soff = (uint32_t)(rl * 128 + (((aj >> 2) ^ (rl & 7)) << 4) + ((aj & 3) << 2));
                sts4(sa + soff, h);
                if (FLIP) sts4(sa + A_TILE_BYTES + soff, hf);
              } else {
                const uint32_t soff = (uint32_t)(rl * 128 + (((aj >> 3) ^ (rl & 7)) << 4) + ((aj & 7) << 1));
                sts2(sa + soff, (uint16_t)h);
                if (FLIP) sts2(sa + A_TILE_BYTES + soff, (uint16_t)hf);
              }
            }
          }
        }

        // ------------------------------------------------ 5. publish the stage to the tensor core
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar0 + 8 * stage);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }

    // ---------------------------------------------------------------- KL side output (generic path only)
    if (!FAST && p.kl_partials != nullptr && blockIdx.x == 0 && blockIdx.z == 0) {
      kl_acc = bt_warp_sum(kl_acc);
      if (lane == 0) red[warp] = kl_acc;
      named_bar_sync(1, NPT);
      if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NPW; ++w) t += red[w];
        p.kl_partials[blockIdx.y] = t;
      }
    }
    if constexpr (!FAST) epilogue((long long)g_first * (MT * BLOCK_M), 0u);
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == NPW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------ persistent weight-stationary kernel
// Reparameterization, bf16 activations.  For layers whose sampled weight tile [BLOCK_N x K] fits shared memory
// next to an activation ring (ResNet stem / layer1 / layer2 at CIFAR resolution; any conv with many more output
// rows than K): the CTA samples W_s for its (n-tile, MC sample) ONCE, keeps every k-block resident in the
// swizzled UMMA layout, and then streams 128-row activation tiles past it as a software pipeline that never
// drains between tiles:
//   warps 0-7   producers : sample the resident tiles, then gather im2col rows with cp.async (LDGSTS, zero-fill
//                           for padding taps) into a ring of 16 KB stages, publishing a stage WS_DEPTH k-blocks
//                           after issuing it -- the copies of several k-blocks (and of the next row tile) are
//                           always in flight, so L2/HBM latency is overlapped;
//   warps 8-11  epilogue  : TMEM -> registers -> (+bias, BatchNorm affine, residual, ReLU) -> global, one TMEM
//                           lane quarter each, on the accumulator buffer the MMA warp finished last;
//   warp  12    MMA       : one thread issues tcgen05.mma into one of TWO accumulator buffers, so the epilogue of
//                           row tile i overlaps the gather + MMA of row tile i+1.
constexpr int WS_PROD_WARPS = 8;
constexpr int WS_EPI_WARPS = 4;
constexpr int WS_THREADS = (WS_PROD_WARPS + WS_EPI_WARPS + 1) * 32;
constexpr int WS_DEPTH = 3;  // k-blocks of cp.async in flight per producer warp

template <int BLOCK_N, bool P_BF16>
__global__ void __launch_bounds__(WS_THREADS, 1) bt_ws_kernel(const __grid_constant__ FusedParams p) {
  constexpr int B_TILE_BYTES = BLOCK_N * 128;
  constexpr int NPT = WS_PROD_WARPS * 32;
  constexpr int P_ES = P_BF16 ? 2 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int res_bytes = p.num_kb * B_TILE_BYTES;
  const int tc_bytes = p.tc_rows > 0 ? 2 * (p.Cin_g / BLOCK_K) * p.tc_rows * 128 : 0;   // input windows (tap-copy mode)
  uint8_t* aux = smem + res_bytes + p.stages * A_TILE_BYTES + tc_bytes;
  int4* row_info = reinterpret_cast<int4*>(aux);                          // [2][128]
  float* bias_s = reinterpret_cast<float*>(aux + 2 * BLOCK_M * 16);       // [3][128]: bias, scale, shift
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux + 2 * BLOCK_M * 16 + 1536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 5);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_base = smem_base + res_bytes;
  const uint32_t full_bar0 = smem_u32(bars);
  const uint32_t empty_bar0 = smem_u32(bars + MAX_STAGES);
  const uint32_t bready_bar = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t acc_bar0 = smem_u32(bars + 2 * MAX_STAGES + 1);    // [2]
  const uint32_t tfree_bar0 = smem_u32(bars + 2 * MAX_STAGES + 3);  // [2]

  const int s = blockIdx.z;
  const int g = blockIdx.y / p.n_tiles_per_group;
  const int n0 = (blockIdx.y % p.n_tiles_per_group) * BLOCK_N;
  const uint32_t sample = p.sample0 + (uint32_t)s + (p.sample_ptr != nullptr ? __ldg(p.sample_ptr) : 0u);
  const int img_base = p.x_shared ? 0 : s * p.B;
  const uint32_t out_sp = (uint32_t)(p.OD * p.OH * p.OW);
  const long long in_sp = (long long)p.ID * p.IH * p.IW;
  const long long n_rt = p.n_groups;  // row tiles (128 rows) per sample; this CTA takes blockIdx.x, +gridDim.x, ...

  if (warp == WS_PROD_WARPS + WS_EPI_WARPS) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(full_bar0 + 8 * i, WS_PROD_WARPS);
        mbar_init(empty_bar0 + 8 * i, 1);
      }
      mbar_init(bready_bar, WS_PROD_WARPS);
      for (int i = 0; i < 2; ++i) {
        mbar_init(acc_bar0 + 8 * i, 1);
        mbar_init(tfree_bar0 + 8 * i, WS_EPI_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  } else if (tid < BLOCK_N) {
    const int n = n0 + tid;
    float b0 = 0.f, sc = 1.f, sh = 0.f;
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
        b0 = mu + bt_softplus(rho) * eps;
      }
      if (p.ep_scale != nullptr) {
        sc = __ldg(p.ep_scale + ng);
        sh = __ldg(p.ep_shift + ng);
      }
    }
    bias_s[tid] = b0;
    bias_s[128 + tid] = sc;
    bias_s[256 + tid] = sh;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WS_PROD_WARPS + WS_EPI_WARPS) {
    // ============================================================== MMA issuer (whole warp, elected issue)
    {
      const uint32_t idesc = make_idesc(BLOCK_N);
      const uint64_t desc_hi = make_smem_desc(0u);
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait_idle(bready_bar, 0, 256);
      tc_fence_after();
      long long it = 0;
      for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
        const int buf = (int)(it & 1);
        if (it >= 2) {  // the epilogue has drained this accumulator buffer
          mbar_wait_idle(tfree_bar0 + 8 * buf, (uint32_t)(((it >> 1) - 1) & 1), 64);
          tc_fence_after();
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_idle(full_bar0 + 8 * stage, phase, 32);
          tc_fence_after();
          const uint32_t sa16 = ((ring_base + stage * A_TILE_BYTES) & 0x3FFFFu) >> 4;
          const uint32_t sb16 = ((smem_base + kb * B_TILE_BYTES) & 0x3FFFFu) >> 4;
          umma_bf16_elect_x4(tmem_base + (uint32_t)(buf * BLOCK_N), sa16 | (1u << 16), sb16 | (1u << 16),
                             (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
          umma_commit_elect(empty_bar0 + 8 * stage);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_elect(acc_bar0 + 8 * buf);
      }
    }
    __syncwarp();
  } else if (warp >= WS_PROD_WARPS) {
    // ============================================================== epilogue warps (one TMEM lane quarter each)
    const int q = warp - WS_PROD_WARPS;
    uint8_t* outb = static_cast<uint8_t*>(p.out);
    long long it = 0;
    for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      mbar_wait_idle(acc_bar0 + 8 * buf, (uint32_t)((it >> 1) & 1), 256);
      tc_fence_after();
      const long long m = rt * BLOCK_M + q * 32 + lane;
      const bool mvalid = m < p.M;
      const long long orow = (long long)s * p.M + m;
#pragma unroll 1
      for (int col0 = 0; col0 < BLOCK_N; col0 += 16) {
        uint32_t v0[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BLOCK_N + col0), v0);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float val = __uint_as_float(v0[j]) + bias_s[col0 + j];
          o[j] = fmaf(val, bias_s[128 + col0 + j], bias_s[256 + col0 + j]);
        }
        if (mvalid) {
          const int nfirst = n0 + col0;
          const long long eoff = orow * p.C_out + g * p.N + nfirst;
          uint8_t* dst = outb + eoff * 2;
          const bool vec_ok = p.out_vec && nfirst + 16 <= p.N;
          if (p.ep_residual != nullptr) {
            const uint8_t* rsd = static_cast<const uint8_t*>(p.ep_residual) + eoff * 2;
            if (vec_ok) {
              const uint4 a = ldg16(rsd), b = ldg16(rsd + 16);
              const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                o[2 * j] += bt_bf16_lo(w[j]);
                o[2 * j + 1] += bt_bf16_hi(w[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (nfirst + j < p.N) o[j] += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rsd)[j]);
            }
          }
          if (p.ep_relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
          }
          if (vec_ok) {
            uint4 a, b;
            a.x = bt_pack_bf16x2(o[0], o[1]);   a.y = bt_pack_bf16x2(o[2], o[3]);
            a.z = bt_pack_bf16x2(o[4], o[5]);   a.w = bt_pack_bf16x2(o[6], o[7]);
            b.x = bt_pack_bf16x2(o[8], o[9]);   b.y = bt_pack_bf16x2(o[10], o[11]);
            b.z = bt_pack_bf16x2(o[12], o[13]); b.w = bt_pack_bf16x2(o[14], o[15]);
            reinterpret_cast<uint4*>(dst)[0] = a;
            reinterpret_cast<uint4*>(dst)[1] = b;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (nfirst + j < p.N) reinterpret_cast<__nv_bfloat16*>(dst)[j] = __float2bfloat16_rn(o[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tfree_bar0 + 8 * buf);
    }
  } else {
    // ============================================================== producers
    const uint8_t* mu_w = static_cast<const uint8_t*>(p.mu_w);
    const uint8_t* rho_w = static_cast<const uint8_t*>(p.rho_w);
    const uint8_t* xb = static_cast<const uint8_t*>(p.x);
    // ---- 1. sample every k-block of W_s for this (n-tile, sample) into the resident region
    {
      constexpr int WO = BLOCK_N / 32;  // octs (8 consecutive k = one Philox call) per thread: rows wrb + 32*i
      constexpr int PW = P_BF16 ? 4 : 8;
      const int wo = tid & 7, wrb = tid >> 3;
      long long row_off[WO];
      bool nvalid[WO];
#pragma unroll
      for (int i = 0; i < WO; ++i) {
        const int n = n0 + wrb + 32 * i;
        nvalid[i] = n < p.N;
        row_off[i] = ((long long)g * p.N + (nvalid[i] ? n : p.N - 1)) * p.K_phys;
      }
      int w_tap = (wo * 8) / p.Cin_g, w_c = (wo * 8) - ((wo * 8) / p.Cin_g) * p.Cin_g;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int ku0 = kb * BLOCK_K + wo * 8;
        const bool kvalid = ku0 < p.K_used;
        long long kphys0 = 0;
        if (kvalid) kphys0 = (long long)decode_tap(p, w_tap).lin * p.Cin_g + w_c;
        w_tap += p.q64;
        w_c += p.r64;
        if (w_c >= p.Cin_g) {
          w_c -= p.Cin_g;
          ++w_tap;
        }
        uint32_t mu_r[WO][PW], rho_r[WO][PW];
#pragma unroll
        for (int i = 0; i < WO; ++i) {
          const long long off = (row_off[i] + kphys0) * P_ES;
          const uint4 a = ldg16(mu_w + off);
          const uint4 b = ldg16(rho_w + off);
          mu_r[i][0] = a.x; mu_r[i][1] = a.y; mu_r[i][2] = a.z; mu_r[i][3] = a.w;
          rho_r[i][0] = b.x; rho_r[i][1] = b.y; rho_r[i][2] = b.z; rho_r[i][3] = b.w;
          if constexpr (!P_BF16) {
            const uint4 a2 = ldg16(mu_w + off + 16);
            const uint4 b2 = ldg16(rho_w + off + 16);
            mu_r[i][4] = a2.x; mu_r[i][5] = a2.y; mu_r[i][6] = a2.z; mu_r[i][7] = a2.w;
            rho_r[i][4] = b2.x; rho_r[i][5] = b2.y; rho_r[i][6] = b2.z; rho_r[i][7] = b2.w;
          }
        }
        uint32_t c[WO][4];
#pragma unroll
        for (int i = 0; i < WO; ++i) {
          c[i][0] = (uint32_t)(kphys0 >> 3);
          c[i][1] = (uint32_t)(g * p.N + n0 + wrb + 32 * i);
          c[i][2] = sample;
          c[i][3] = p.key.c3_base | BT_STREAM_W_EPS;
        }
        philox_multi<WO>(c, p.key.k0, p.key.k1);
        const uint32_t sb = smem_base + kb * B_TILE_BYTES;
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
              m8[2 * j] = bt_bf16_lo(mu_r[i][j]);
              m8[2 * j + 1] = bt_bf16_hi(mu_r[i][j]);
              r8[2 * j] = bt_bf16_lo(rho_r[i][j]);
              r8[2 * j + 1] = bt_bf16_hi(rho_r[i][j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              m8[j] = __uint_as_float(mu_r[i][j]);
              r8[j] = __uint_as_float(rho_r[i][j]);
            }
          }
          const bool ok = kvalid && nvalid[i];
          float w0[8];
#pragma unroll
          if (!p.rho_is_sigma) {   // (warp-uniform)
#pragma unroll
            for (int j = 0; j < 8; ++j) r8[j] = bt_softplus_fast(r8[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) w0[j] = ok ? fmaf(r8[j], e[j], m8[j]) : 0.f;
          const int nl = wrb + 32 * i;
          sts16(sb + (uint32_t)(nl * 128 + ((wo ^ (nl & 7)) << 4)),
                make_uint4(bt_pack_bf16x2(w0[0], w0[1]), bt_pack_bf16x2(w0[2], w0[3]),
                           bt_pack_bf16x2(w0[4], w0[5]), bt_pack_bf16x2(w0[6], w0[7])));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bready_bar);
    }

    // ---- 2. stream the row tiles: cp.async gather, publish WS_DEPTH k-blocks behind the issue cursor.
    // Mapping: 8 consecutive lanes copy the 8 16-byte chunks of one 128-byte row (fully coalesced; a row-per-thread
    // mapping halves the instruction count but makes the L2 reads strided and was 1.7x slower, profiles/r01f);
    // thread t handles chunk (t & 7) of rows (t >> 3) + 32 i.  Everything that depends only on the k-block (tap ->
    // pixel delta, channel byte offset, mask bit) comes from a small table built once per CTA.
    const int ac = tid & 7, arb = tid >> 3;
    const uint32_t pix_bytes = (uint32_t)p.C_in * 2u;
    int4* ktab = reinterpret_cast<int4*>(aux + 2 * BLOCK_M * 16 + 1536 + 256);   // [num_kb][8] {dpix, byte off, tap, valid}
    for (int e = tid; e < p.num_kb * 8; e += NPT) {
      const int ku = (e >> 3) * BLOCK_K + (e & 7) * 8;
      int4 ent = make_int4(0, 0, 0, 0);
      if (ku < p.K_used) {
        const int tap_i = ku / p.Cin_g;
        const TapCoord tc = decode_tap(p, tap_i);
        ent = make_int4((tc.dz * p.IH + tc.dy) * p.IW + tc.dx, (g * p.Cin_g + (ku - tap_i * p.Cin_g)) * 2, tap_i, 1);
      }
      ktab[e] = ent;
    }
    uint32_t soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) soff[i] = (uint32_t)((arb + 32 * i) * 128 + ((ac ^ ((arb + 32 * i) & 7)) << 4));
    auto fill_rows = [&](long long rt, int slot) {  // threads 0..127: metadata of one row each
      if (tid < BLOCK_M) {
        const long long m = rt * BLOCK_M + tid;
        int4 info = make_int4(0, 0, 0, 0);
        if (m < p.M) {
          const uint32_t mm = (uint32_t)m, hw = (uint32_t)(p.OH * p.OW);
          const int b = (int)(mm / out_sp);
          uint32_t rem = mm - (uint32_t)b * out_sp;
          const int od = (int)(rem / hw);
          rem -= (uint32_t)od * hw;
          const int oh = (int)(rem / (uint32_t)p.OW);
          const int ow = (int)(rem - (uint32_t)oh * (uint32_t)p.OW);
          const int z0 = od * p.sd - p.pd, y0 = oh * p.sh - p.ph, x0 = ow * p.sw - p.pw;
          const long long pix0 = (((long long)(img_base + b) * p.ID + z0) * p.IH + y0) * p.IW + x0;
          unsigned long long mask = 0ull;
          if (!p.taps_natural) {
            const int n_taps = p.K_used / p.Cin_g;
            for (int t = 0; t < n_taps; ++t) {
              const uint32_t tp = p.taps[t];
              const int kd = tp & 0xff, kh = (tp >> 8) & 0xff, kw = (tp >> 16) & 0xff;
              const bool inb = (unsigned)(z0 + kd * p.dd) < (unsigned)p.ID && (unsigned)(y0 + kh * p.dh) < (unsigned)p.IH &&
                               (unsigned)(x0 + kw * p.dw) < (unsigned)p.IW;
              mask |= (unsigned long long)(inb ? 1 : 0) << t;
            }
          } else {
            unsigned long long xm = 0ull;
            for (int kw = 0; kw < p.KW; ++kw)
              xm |= (unsigned long long)((unsigned)(x0 + kw * p.dw) < (unsigned)p.IW ? 1 : 0) << kw;
            for (int kd = 0; kd < p.KD; ++kd) {
              if ((unsigned)(z0 + kd * p.dd) >= (unsigned)p.ID) continue;
              for (int kh = 0; kh < p.KH; ++kh)
                if ((unsigned)(y0 + kh * p.dh) < (unsigned)p.IH) mask |= xm << ((kd * p.KH + kh) * p.KW);
            }
          }
          info = make_int4((int)(uint32_t)pix0, (int)(uint32_t)mask, (int)(uint32_t)(mask >> 32), 1);
        }
        row_info[slot * BLOCK_M + tid] = info;
      }
    };
    int stage = 0, arr_stage = 0, in_flight = 0;
    uint32_t phase = 0;
    const int depth = p.stages - 2 >= WS_DEPTH ? WS_DEPTH : (p.stages - 2 >= 2 ? 2 : 1);   // needs stages >= depth + 2
    auto publish = [&]() {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar0 + 8 * arr_stage);
      if (++arr_stage == p.stages) arr_stage = 0;
      --in_flight;
    };
    if ((long long)blockIdx.x < n_rt) fill_rows(blockIdx.x, 0);
    long long it = 0;
    if (p.tc_rows > 0) {
      // ---- tap-copy mode (stride-1 "same" convolutions, Cin multiple of 64): the pixels a row tile needs --
      // [m0 - halo, m0 + 128 + halo) of the sample's flattened pixel sequence -- are loaded ONCE per 64-channel slab
      // into a double-buffered input window (cp.async, prefetched one tile ahead); every filter tap's A tile is then
      // a row-shifted shared->shared copy (source row = r + delta_tap, zero where the tap mask says padding).
      // L2 sees each activation once instead of once per tap (9x for 3x3).
      const int slabs = p.Cin_g / BLOCK_K;
      const int R = p.tc_rows;
      const uint32_t inbuf0 = ring_base + p.stages * A_TILE_BYTES;            // [2][slabs][R][128 B]
      const uint32_t inbuf_bytes = (uint32_t)(slabs * R * 128);
      const long long sample_pix0 = (long long)img_base * in_sp;              // first pixel of this sample in x
      auto load_window = [&](long long rt, int slot) {
        const long long first = rt * BLOCK_M - p.tc_halo;                     // pixel (within the sample) of buffer row 0
        for (int e = tid; e < slabs * R * 8; e += NPT) {
          const int c = e & 7, row = (e >> 3) % R, sl = (e >> 3) / R;
          const long long pix = first + row;
          const bool ok = pix >= 0 && pix < p.M;
          cp_async16(inbuf0 + slot * inbuf_bytes + (uint32_t)((sl * R + row) * 128 + c * 16),
                     xb + ((sample_pix0 + (ok ? pix : 0)) * p.C_in + g * p.Cin_g + sl * BLOCK_K + c * 8) * 2,
                     ok ? 16u : 0u);
        }
        cp_async_commit();
      };
      if ((long long)blockIdx.x < n_rt) load_window(blockIdx.x, 0);
      for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
        cp_async_wait<0>();                 // this thread's part of window[it & 1] has landed ...
        named_bar_sync(1, NPT);             // ... and everybody else's; row_info[it & 1] / ktab complete
        unsigned long long rmask[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int4 info = row_info[(int)(it & 1) * BLOCK_M + arb + 32 * i];
          rmask[i] = (unsigned long long)(uint32_t)info.y | ((unsigned long long)(uint32_t)info.z << 32);
        }
        if (rt + gridDim.x < n_rt) {
          fill_rows(rt + gridDim.x, (int)((it + 1) & 1));
          load_window(rt + gridDim.x, (int)((it + 1) & 1));   // in flight during this tile's tap copies
        }
        const uint32_t win = inbuf0 + (uint32_t)(it & 1) * inbuf_bytes;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
          const uint32_t sst = ring_base + stage * A_TILE_BYTES;
          const int4 e = ktab[kb * 8];                          // all 8 chunks of the k-block share the tap
          const int sl = kb % slabs;
          const int shift = e.x - p.tc_padoff + p.tc_halo;     // buffer row of output row 0 for this tap
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = arb + 32 * i;
            const bool ok = e.w != 0 && ((rmask[i] >> e.z) & 1ull);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (ok) {
              const uint32_t a = win + (uint32_t)((sl * R + rl + shift) * 128 + ac * 16);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
            }
            sts16(sst + soff[i], v);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar0 + 8 * stage);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      cp_async_wait<0>();
    } else
    for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
      named_bar_sync(1, NPT);  // row_info[it & 1] (and, the first time, ktab) complete; row_info[(it+1)&1] is free
      uint32_t rpix[4];
      unsigned long long rmask[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int4 info = row_info[(int)(it & 1) * BLOCK_M + arb + 32 * i];
        rpix[i] = (uint32_t)info.x;
        rmask[i] = (unsigned long long)(uint32_t)info.y | ((unsigned long long)(uint32_t)info.z << 32);
      }
      if (rt + gridDim.x < n_rt) fill_rows(rt + gridDim.x, (int)((it + 1) & 1));
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
        const uint32_t sst = ring_base + stage * A_TILE_BYTES;
        const int4 e = ktab[kb * 8 + ac];
        const uint8_t* xcol = xb + (uint32_t)e.y;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = e.w != 0 && ((rmask[i] >> e.z) & 1ull);
          cp_async16(sst + soff[i], xcol + (unsigned long long)(ok ? rpix[i] + (uint32_t)e.x : 0u) * pix_bytes,
                     ok ? 16u : 0u);
        }
        cp_async_commit();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
        if (++in_flight > depth) {  // the oldest in-flight k-block has landed
          if (depth == 3) cp_async_wait<3>();
          else if (depth == 2) cp_async_wait<2>();
          else cp_async_wait<1>();
          publish();
        }
      }
    }
    cp_async_wait<0>();
    while (in_flight > 0) publish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WS_PROD_WARPS + WS_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// KL finalize: fixed-order sum of the per-tile partials + the (tiny) bias term.
__global__ void bt_fused_kl_finalize(const float* partials, int n_partials, long long n_w,
                                     const void* mu_b, const void* rho_b, int n_b, int p_is_bf16,
                                     float pmu, float log_ps, float inv2, float* out) {
  const int lane = threadIdx.x;
  float s = 0.f;
  for (int i = lane; i < n_partials; i += 32) s += partials[i];
  s = bt_warp_sum(s);
  float sb = 0.f;
  for (int i = lane; i < n_b; i += 32) {
    float mu, rho;
    if (p_is_bf16) {
      mu = __bfloat162float(static_cast<const __nv_bfloat16*>(mu_b)[i]);
      rho = __bfloat162float(static_cast<const __nv_bfloat16*>(rho_b)[i]);
    } else {
      mu = static_cast<const float*>(mu_b)[i];
      rho = static_cast<const float*>(rho_b)[i];
    }
    sb += bt_kl_elem(mu, bt_softplus(rho), pmu, log_ps, inv2);
  }
  sb = bt_warp_sum(sb);
  if (lane == 0) *out = s / (float)n_w + (n_b > 0 ? sb / (float)n_b : 0.f);
}

// ------------------------------------------------------------------ host side
struct DevInfo {
  int sm_count = 0;
};
std::mutex g_mu;
DevInfo g_dev[64];

template <int BN, bool FLIP, int NPW, bool FAST, bool PB, bool XB, int FMT, bool TF32 = false>
int launch_fused(const FusedParams& p, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  static bool attr_done[64] = {};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!attr_done[dev]) {
      BT_CHECK_CUDA(cudaFuncSetAttribute(bt_fused_kernel<BN, FLIP, NPW, FAST, PB, XB, FMT, TF32>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
      attr_done[dev] = true;
    }
  }
  bt_fused_kernel<BN, FLIP, NPW, FAST, PB, XB, FMT, TF32><<<grid, NPW * 32 + 32, smem_bytes, st>>>(p);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

template <int BN, bool FLIP, bool PB, bool XB>
int dispatch_fast(const FusedParams& p, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  switch (p.MT) {
    case 1: return launch_fused<BN, FLIP, FAST_WARPS, true, PB, XB, 1>(p, grid, smem_bytes, dev, st);
    case 2: return launch_fused<BN, FLIP, FAST_WARPS, true, PB, XB, 2>(p, grid, smem_bytes, dev, st);
    default: return launch_fused<BN, FLIP, FAST_WARPS, true, PB, XB, 4>(p, grid, smem_bytes, dev, st);
  }
}

template <int BN, bool FLIP>
int dispatch_fused(const FusedParams& p, bool fast, bool tf32, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  if (fast) {
    if (p.p_is_bf16 && p.x_is_bf16) return dispatch_fast<BN, FLIP, true, true>(p, grid, smem_bytes, dev, st);
    if (!p.p_is_bf16 && !p.x_is_bf16) return dispatch_fast<BN, FLIP, false, false>(p, grid, smem_bytes, dev, st);
    if (!p.p_is_bf16 && p.x_is_bf16) return dispatch_fast<BN, FLIP, false, true>(p, grid, smem_bytes, dev, st);
  }
  if (tf32) return launch_fused<BN, FLIP, GENERIC_WARPS, false, false, false, 0, true>(p, grid, smem_bytes, dev, st);
  return launch_fused<BN, FLIP, GENERIC_WARPS, false, false, false, 0>(p, grid, smem_bytes, dev, st);
}

template <int BN, bool PB>
int launch_ws(const FusedParams& p, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  static bool attr_done[64] = {};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!attr_done[dev]) {
      BT_CHECK_CUDA(cudaFuncSetAttribute(bt_ws_kernel<BN, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BUDGET));
      attr_done[dev] = true;
    }
  }
  bt_ws_kernel<BN, PB><<<grid, WS_THREADS, smem_bytes, st>>>(p);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

#include "bt_direct.cuh"
#include "bt_tma.cuh"

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

int64_t bt_forward_workspace_bytes(void) { return 65536 * 4; }

static thread_local int g_last_path = -1;
int bt_last_forward_path(void) { return g_last_path; }

// bt_layer_forward and bt_layer_forward_plan share this body: validation + tiling search + kernel selection are pure
// host arithmetic on the geometry (plan != NULL: stop there and report the decision -- no device is touched, `sm_plan`
// stands in for the SM count and NULL pointers count as 16-byte aligned); the launch follows for plan == NULL.
static int layer_forward_impl(BtForwardPlan* plan, int sm_plan, int mode, const BtLayerGeom* gm, const void* x, int x_dtype,
                              const void* mu_w, const void* rho_w, const void* mu_b, const void* rho_b, int p_dtype,
                              void* out, float* kl_out, float prior_mu_s, float prior_sigma_s, uint64_t seed,
                              uint32_t layer_key, uint32_t sample_idx0, const BtDebugIO* dbg,
                              const BtEpilogue* epi, void* workspace, void* stream) {
  const bool plan_only = plan != nullptr;
  BT_REQUIRE(gm != nullptr, BT_ERR_BAD_POINTER, "bt_layer_forward: geom is NULL");
  BT_REQUIRE(mode == BT_MODE_REPARAM || mode == BT_MODE_FLIPOUT, BT_ERR_UNSUPPORTED,
             "bt_layer_forward: mode %d", mode);
  BT_REQUIRE((x_dtype == BT_F32 || x_dtype == BT_BF16) && (p_dtype == BT_F32 || p_dtype == BT_BF16),
             BT_ERR_BAD_DTYPE, "bt_layer_forward: dtypes x=%d p=%d (supported: f32, bf16)", x_dtype, p_dtype);
  BT_REQUIRE(gm->n_samples >= 1 && gm->n_samples <= 65535, BT_ERR_BAD_SHAPE,
             "bt_layer_forward: n_samples %d out of [1,65535]", gm->n_samples);
  BT_REQUIRE(gm->batch >= 1 && gm->c_in >= 1 && gm->c_out >= 1 && gm->groups >= 1, BT_ERR_BAD_SHAPE,
             "bt_layer_forward: batch/c_in/c_out/groups must be >= 1");
  BT_REQUIRE(gm->c_in % gm->groups == 0 && gm->c_out % gm->groups == 0, BT_ERR_BAD_SHAPE,
             "bt_layer_forward: channels (%d,%d) not divisible by groups %d", gm->c_in, gm->c_out, gm->groups);
  BT_REQUIRE(layer_key < (1u << 28), BT_ERR_BAD_SHAPE, "bt_layer_forward: layer_key must be < 2^28");
  for (int i = 0; i < 3; ++i) {
    BT_REQUIRE(gm->in_dhw[i] >= 1 && gm->k_dhw[i] >= 1 && gm->k_dhw[i] <= 255 && gm->stride[i] >= 1 &&
                   gm->dil[i] >= 1 && gm->pad[i] >= 0,
               BT_ERR_BAD_SHAPE, "bt_layer_forward: bad conv geometry in dim %d", i);
    const long long eff = (long long)gm->dil[i] * (gm->k_dhw[i] - 1) + 1;
    if (gm->transposed) {   // out = (in - 1) * stride - 2 pad + dil (k - 1) + output_padding + 1, output_padding < max(stride, dil)
      const long long o0 = ((long long)gm->in_dhw[i] - 1) * gm->stride[i] - 2ll * gm->pad[i] + eff;
      const long long slack = gm->stride[i] > gm->dil[i] ? gm->stride[i] : gm->dil[i];
      BT_REQUIRE(o0 >= 1 && gm->out_dhw[i] >= o0 && gm->out_dhw[i] < o0 + slack, BT_ERR_BAD_SHAPE,
                 "bt_layer_forward: transposed out_dhw[%d]=%d inconsistent with input %d k %d s %d p %d d %d", i,
                 gm->out_dhw[i], gm->in_dhw[i], gm->k_dhw[i], gm->stride[i], gm->pad[i], gm->dil[i]);
      continue;
    }
    const long long o = ((long long)gm->in_dhw[i] + 2ll * gm->pad[i] - eff) / gm->stride[i] + 1;
    BT_REQUIRE((long long)gm->in_dhw[i] + 2ll * gm->pad[i] >= eff && o == gm->out_dhw[i], BT_ERR_BAD_SHAPE,
               "bt_layer_forward: out_dhw[%d]=%d inconsistent with input %d k %d s %d p %d d %d", i,
               gm->out_dhw[i], gm->in_dhw[i], gm->k_dhw[i], gm->stride[i], gm->pad[i], gm->dil[i]);
  }
  int rc = BT_OK;
  int dev = 0;
  int sm_count = sm_plan;
  if (!plan_only) {
    if ((rc = bt_device_check()) != BT_OK) return rc;
    if ((rc = bt_check_device_ptr(x, "x")) != BT_OK) return rc;
    if ((rc = bt_check_device_ptr(mu_w, "mu_w")) != BT_OK) return rc;
    if ((rc = bt_check_device_ptr(rho_w, "rho_w")) != BT_OK) return rc;
    if ((rc = bt_check_device_ptr(out, "out")) != BT_OK) return rc;
    BT_REQUIRE((mu_b == nullptr) == (rho_b == nullptr), BT_ERR_BAD_POINTER,
               "bt_layer_forward: mu_b and rho_b must both be given or both NULL");
    if (kl_out != nullptr) {
      if ((rc = bt_check_device_ptr(kl_out, "kl_out")) != BT_OK) return rc;
      if ((rc = bt_check_device_ptr(workspace, "workspace")) != BT_OK) return rc;
      BT_REQUIRE(prior_sigma_s > 0.f, BT_ERR_BAD_SHAPE, "bt_layer_forward: prior sigma must be > 0");
    }
    BT_CHECK_CUDA(cudaGetDevice(&dev));
    BT_REQUIRE(dev >= 0 && dev < 64, BT_ERR_UNSUPPORTED, "device index %d", dev);
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if (g_dev[dev].sm_count == 0)
        BT_CHECK_CUDA(cudaDeviceGetAttribute(&g_dev[dev].sm_count, cudaDevAttrMultiProcessorCount, dev));
      sm_count = g_dev[dev].sm_count;
    }
  } else {
    BT_REQUIRE(sm_plan >= 1, BT_ERR_BAD_SHAPE, "bt_layer_forward_plan: sm_count must be >= 1");
  }

  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.out = out; p.mu_w = mu_w; p.rho_w = rho_w; p.mu_b = mu_b; p.rho_b = rho_b;
  int dbg_any = 0;
  if (dbg) {
    p.eps_w_in = dbg->eps_w_in; p.eps_b_in = dbg->eps_b_in;
    p.sign_in = dbg->sign_in; p.sign_out = dbg->sign_out;
    dbg_any = (dbg->eps_w_in || dbg->eps_b_in || dbg->sign_in || dbg->sign_out) ? 1 : 0;
  }
  if (epi) {
    BT_REQUIRE((epi->scale == nullptr) == (epi->shift == nullptr), BT_ERR_BAD_POINTER,
               "bt_layer_forward: epilogue scale and shift must both be given or both NULL");
    p.ep_scale = epi->scale; p.ep_shift = epi->shift; p.ep_residual = epi->residual; p.ep_relu = epi->relu ? 1 : 0;
    if (!plan_only && epi->scale && (rc = bt_check_device_ptr(epi->scale, "epilogue scale")) != BT_OK) return rc;
    if (!plan_only && epi->residual && (rc = bt_check_device_ptr(epi->residual, "epilogue residual")) != BT_OK) return rc;
  }
  p.S = gm->n_samples; p.x_shared = gm->x_shared ? 1 : 0; p.B = gm->batch;
  p.C_in = gm->c_in; p.C_out = gm->c_out; p.groups = gm->groups;
  p.Cin_g = gm->c_in / gm->groups; p.N = gm->c_out / gm->groups;
  p.ID = gm->in_dhw[0]; p.IH = gm->in_dhw[1]; p.IW = gm->in_dhw[2];
  p.OD = gm->out_dhw[0]; p.OH = gm->out_dhw[1]; p.OW = gm->out_dhw[2];
  p.KD = gm->k_dhw[0]; p.KH = gm->k_dhw[1]; p.KW = gm->k_dhw[2];
  p.sd = gm->stride[0]; p.sh = gm->stride[1]; p.sw = gm->stride[2];
  p.pd = gm->pad[0]; p.ph = gm->pad[1]; p.pw = gm->pad[2];
  p.dd = gm->dil[0]; p.dh = gm->dil[1]; p.dw = gm->dil[2];
  const long long out_sp = (long long)p.OD * p.OH * p.OW;
  const long long in_sp = (long long)p.ID * p.IH * p.IW;
  p.M = (long long)p.B * out_sp;
  BT_REQUIRE((long long)p.B * in_sp < (1ll << 32) && p.M < (1ll << 32), BT_ERR_BAD_SHAPE,
             "bt_layer_forward: more than 2^32 pixels per sample");
  BT_REQUIRE((long long)p.S * p.B < (1ll << 31), BT_ERR_BAD_SHAPE, "bt_layer_forward: S*B too large");
  const int taps_all = p.KD * p.KH * p.KW;
  const long long kphys = (long long)taps_all * p.Cin_g;
  BT_REQUIRE(kphys < (1ll << 31), BT_ERR_BAD_SHAPE, "bt_layer_forward: K too large");
  p.K_phys = (int)kphys;
  p.x_is_bf16 = x_dtype == BT_BF16;
  p.p_is_bf16 = p_dtype == BT_BF16;
  p.rho_is_sigma = gm->rho_is_sigma ? 1 : 0;
  BT_REQUIRE(!(p.rho_is_sigma && kl_out != nullptr), BT_ERR_UNSUPPORTED,
             "bt_layer_forward: the KL side output needs rho, not a cached sigma (geom.rho_is_sigma)");
  const int p_es = p.p_is_bf16 ? 2 : 4, x_es = p.x_is_bf16 ? 2 : 4;

  p.transposed = gm->transposed ? 1 : 0;
  // fused 3x3 / stride-2 / pad-1 max-pool behind the epilogue (BtLayerGeom.pool_hw): whole output rows per 128-row tile
  p.probe = (getenv("BT_DYNAMIC_ENV") != nullptr && getenv("BT_TMA_PROBE") != nullptr) ? atoi(getenv("BT_TMA_PROBE")) : 0;
  p.pool_oh = gm->pool_hw[0];
  p.pool_ow = gm->pool_hw[1];
  const bool want_pool = p.pool_oh != 0 || p.pool_ow != 0;
  bool pool_geom_ok = false;
  if (want_pool) {
    BT_REQUIRE(p.pool_oh > 0 && p.pool_ow > 0, BT_ERR_BAD_SHAPE, "bt_layer_forward: pool_hw must be {OH, OW} > 0 or {0, 0}");
    const long long pix = (long long)p.pool_oh * p.pool_ow;
    BT_REQUIRE(p.M % pix == 0 && (out_sp == 1 || (p.OD == 1 && p.OH == p.pool_oh && p.OW == p.pool_ow)), BT_ERR_BAD_SHAPE,
               "bt_layer_forward: pool_hw %dx%d does not tile the layer's output rows", p.pool_oh, p.pool_ow);
    pool_geom_ok = p.pool_ow <= 64 && BLOCK_M % p.pool_ow == 0 && (BLOCK_M / p.pool_ow) % 2 == 0 && pix % BLOCK_M == 0;
  }
  p.sample_ptr = gm->sample_offset;
  if (!plan_only && p.sample_ptr != nullptr && (rc = bt_check_device_ptr(p.sample_ptr, "sample_offset")) != BT_OK) return rc;
  // fp32 parameters with fp32 activations (the reference's default dtype): tf32 operands, 32 k per k-block
  static const bool tf32_disabled = getenv("BT_DISABLE_TF32") != nullptr;   // A/B switch (bf16 operand rounding instead)
  const bool tf32 = !p.p_is_bf16 && !p.x_is_bf16 && !tf32_disabled;
  const int KB = tf32 ? 32 : BLOCK_K;
  const int a_ce = tf32 ? 4 : 8;   // activation elements per 16-byte shared-memory chunk
  p.w_vec = (p.Cin_g % 4 == 0) && ((reinterpret_cast<uintptr_t>(mu_w) % (4 * p_es)) == 0) &&
            ((reinterpret_cast<uintptr_t>(rho_w) % (4 * p_es)) == 0);
  p.a_vec = (p.Cin_g % a_ce == 0) && (p.C_in % a_ce == 0) && al16(x);
  p.out_vec = ((long long)p.C_out * x_es) % 16 == 0 && al16(out) && ((long long)p.N * x_es) % 16 == 0 &&
              (p.ep_residual == nullptr || al16(p.ep_residual));

  // taps that touch at least one real input element for at least one output position; the others
  // multiply zero padding only and are skipped exactly (their weights are neither read nor sampled).
  // Disabled when the KL side output is requested (KL needs every weight).
  int n_used = taps_all;
  p.taps_explicit = 0;
  p.taps_natural = 1;
  if (p.a_vec && p.w_vec && taps_all <= MAX_TAPS && !p.transposed) {
    auto dim_ok = [](int k, int dil, int pad, int stride, int in, int outn) {
      for (int o = 0; o < outn; ++o) {
        const int i = o * stride - pad + k * dil;
        if (i >= 0 && i < in) return true;
      }
      return false;
    };
    const bool may_skip = kl_out == nullptr;   // the KL side output needs every weight
    int cnt = 0;
    for (int kd = 0; kd < p.KD; ++kd)
      for (int kh = 0; kh < p.KH; ++kh)
        for (int kw = 0; kw < p.KW; ++kw)
          if (!may_skip || (dim_ok(kd, p.dd, p.pd, p.sd, p.ID, p.OD) && dim_ok(kh, p.dh, p.ph, p.sh, p.IH, p.OH) &&
                            dim_ok(kw, p.dw, p.pw, p.sw, p.IW, p.OW)))
            p.taps[cnt++] = (uint32_t)kd | ((uint32_t)kh << 8) | ((uint32_t)kw << 16);
    BT_REQUIRE(cnt >= 1, BT_ERR_BAD_SHAPE, "bt_layer_forward: no filter tap touches the input");
    n_used = cnt;
    p.taps_explicit = 1;          // the kernels read tap coordinates from the list (no div/mod per k-block)
    p.taps_natural = cnt == taps_all;
  }
  p.q64 = BLOCK_K / p.Cin_g;
  p.r64 = BLOCK_K % p.Cin_g;
  p.K_used = n_used * p.Cin_g;
  p.num_kb = (p.K_used + KB - 1) / KB;

  const bool flip = mode == BT_MODE_FLIPOUT;
  const int NB = flip ? 2 : 1;
  const int max_mt = flip ? 2 : 4;
  const long long m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const bool fast = p.w_vec && p.a_vec && dbg_any == 0 && kl_out == nullptr && n_used <= 64 && !tf32 && !p.transposed &&
                    (long long)(p.x_shared ? 1 : p.S) * p.B * in_sp < (1ll << 32);
  // Tiling: minimise  waves * per-CTA work  over the column tile BN, (a) the M-subtiles per CTA that share one
  // sampled weight tile and (b) -- fast path -- a weight-stationary schedule: the CTA samples all k-blocks of its
  // (n-tile, sample) once, keeps them in shared memory and streams `groups` of MT M-subtiles past them.
  // Sampling a weight element costs ~10x gathering an activation element, so re-sampling is what the search
  // avoids (a narrower BN can make the sampled tiles fit shared memory at the price of gathering A once more).
  static const bool ws_disabled = getenv("BT_DISABLE_WS") != nullptr;   // A/B switch for benchmarking
  const double c_s = 1.0, c_a = p.a_vec ? 0.08 : 0.6, c_e = 0.2;
  static const bool tc_disabled = getenv("BT_DISABLE_TAPCOPY") != nullptr;   // A/B switch
  int BN = 128, mt = 1, ws = 0, ws_x = 1;
  int tc_rows_sel = 0, tc_halo_sel = 0, tc_padoff_sel = 0, tc_cand_halo = 0, tc_cand_padoff = 0;
  long long tc_bytes_sel = 0;
  double best = 1e300;
  const int bn_cands[2] = {p.N <= 64 ? 64 : 128, 64};
  for (int bi = 0; bi < (p.N <= 64 ? 1 : 2); ++bi) {
    const int bn = bn_cands[bi];
    const long long nt = (long long)((p.N + bn - 1) / bn) * p.groups;
    for (int cand = 1; cand <= max_mt; cand <<= 1) {
      if (cand > 1 && cand / 2 >= m_tiles) break;
      const long long groups = (m_tiles + cand - 1) / cand;
      const long long ctas = groups * nt * p.S;
      const double waves = (double)((ctas + sm_count - 1) / sm_count);
      const double t_cta = p.num_kb * (bn * 64.0 * c_s + cand * 128.0 * 64.0 * c_a) + cand * 128.0 * bn * c_e;
      if (waves * t_cta < best) {
        best = waves * t_cta;
        BN = bn; mt = cand; ws = 0;
      }
    }
    if (!fast || ws_disabled) continue;
    const long long res_bytes = (long long)p.num_kb * NB * bn * 128;
    // (b1) persistent weight-stationary kernel (bt_ws_kernel): reparameterization + bf16 activations
    if (!flip && p.x_is_bf16 && m_tiles >= 2 && p.M < (1ll << 31) && p.num_kb <= 48) {
      // tap-copy variant: stride-1 "same" convolution whose input window fits next to the resident tiles
      long long tc_bytes = 0;
      int tc_rows_c = 0;
      {
        const bool same = p.sd == 1 && p.sh == 1 && p.sw == 1 && p.ID == p.OD && p.IH == p.OH && p.IW == p.OW &&
                          p.groups == 1 && p.Cin_g % BLOCK_K == 0 && taps_all > 1 && !tc_disabled;
        if (same) {
          const int padoff = (p.pd * p.IH + p.ph) * p.IW + p.pw;
          const int maxd = ((p.KD - 1) * p.dd * p.IH + (p.KH - 1) * p.dh) * p.IW + (p.KW - 1) * p.dw;
          const int halo = padoff > maxd - padoff ? padoff : maxd - padoff;
          const long long bytes = 2ll * (p.Cin_g / BLOCK_K) * (BLOCK_M + 2 * halo) * 128;
          if (halo <= 96 && res_bytes + bytes + 3 * A_TILE_BYTES + AUX_BYTES + 1024 <= SMEM_BUDGET) {
            tc_bytes = bytes;
            tc_rows_c = BLOCK_M + 2 * halo;
            tc_cand_halo = halo;
            tc_cand_padoff = padoff;
          }
        }
      }
      long long st = (SMEM_BUDGET - AUX_BYTES - 1024 - res_bytes - tc_bytes) / A_TILE_BYTES;
      if (st > MAX_STAGES) st = MAX_STAGES;
      if (st >= 3) {
        const long long xmax = m_tiles < 4 * sm_count ? m_tiles : 4 * sm_count;
        for (long long x = 1; x <= xmax; ++x) {
          const long long ctas = x * nt * p.S;
          const double waves = (double)((ctas + sm_count - 1) / sm_count);
          const double per = (double)((m_tiles + x - 1) / x);
          // measured (profiles/r01f): a gathered 16 KB stage costs ~800 clocks (L2-bound im2col reads), a tap-copied
          // one ~300; the 8 producer warps sample at ~1.5 clocks per element
          const double t_cta = p.num_kb * bn * 64.0 * 1.5 * c_s + per * (p.num_kb * (tc_rows_c ? 300.0 : 800.0) + 300.0);
          if (waves * t_cta < 0.95 * best) {
            best = waves * t_cta / 0.95;
            BN = bn; mt = 1; ws = 2; ws_x = (int)x;
            tc_rows_sel = tc_rows_c; tc_bytes_sel = tc_bytes; tc_halo_sel = tc_cand_halo; tc_padoff_sel = tc_cand_padoff;
          }
        }
      }
    }
    // (b2) weight-stationary mode of the general kernel (Flipout / fp32 activations); its M-groups are not
    // overlapped (same warps gather and run the epilogue), hence the 1.5x handicap
    // the sampled tiles are resident, so MT no longer buys weight reuse: prefer small M-groups = a deeper
    // activation ring (the cp.async gather needs > 4 stages); ties keep the first candidate
    const int ws_cands[3] = {1, 2, 4};
    for (int ci = 0; ci < 3; ++ci) {
      const int cand = ws_cands[ci];
      if (cand > max_mt) continue;
      if (cand > 1 && cand / 2 >= m_tiles) continue;
      const long long a_stage = (long long)NB * cand * A_TILE_BYTES;
      if (res_bytes + 2 * a_stage + AUX_BYTES + 1024 > SMEM_BUDGET) continue;
      const long long groups = (m_tiles + cand - 1) / cand;
      if (groups < 2) continue;  // nothing to amortise
      const bool deep = (SMEM_BUDGET - AUX_BYTES - 1024 - res_bytes) / a_stage >= 5 && p.x_is_bf16 && !flip;
      const double c_aw = deep ? 0.5 * c_a : c_a;
      const long long xmax = groups < 4 * sm_count ? groups : 4 * sm_count;
      for (long long x = 1; x <= xmax; ++x) {
        const long long ctas = x * nt * p.S;
        const double waves = (double)((ctas + sm_count - 1) / sm_count);
        const double per = (double)((groups + x - 1) / x);
        const double t_cta = p.num_kb * bn * 64.0 * c_s +
                             1.5 * per * (p.num_kb * cand * 128.0 * 64.0 * c_aw + cand * 128.0 * bn * c_e);
        if (waves * t_cta < 0.95 * best) {
          best = waves * t_cta / 0.95;
          BN = bn; mt = cand; ws = 1; ws_x = (int)x;
        }
      }
    }
  }
  // (c) direct mode (bt_direct_kernel, bt_direct.cuh): stride-1 "same" convolutions / linears with C_in % 64 == 0
  // and bf16 activations whose sampled tiles AND two input windows fit shared memory -- the A operand is read in
  // place from the window (descriptor row shift per tap), no im2col copies at all.
  // (read on every call so that the parity tests can A/B the paths inside one process)
  // The A/B switches are read from the environment on every call ONLY when BT_DYNAMIC_ENV was set at load time (the
  // parity tests flip them inside one process); production reads them once (9 getenv calls per launch before).
  static const bool dyn_env = getenv("BT_DYNAMIC_ENV") != nullptr;
  struct DrEnv { bool disabled, force; int bn_only, x_only, slots_max; bool times; };
  auto read_env = []() {
    DrEnv e;
    e.disabled = getenv("BT_DISABLE_DIRECT") != nullptr;
    e.force = getenv("BT_FORCE_DIRECT") != nullptr;
    e.bn_only = getenv("BT_DIRECT_BN") ? atoi(getenv("BT_DIRECT_BN")) : 0;
    e.x_only = getenv("BT_DIRECT_X") ? atoi(getenv("BT_DIRECT_X")) : 0;
    e.slots_max = getenv("BT_DIRECT_SLOTS") ? atoi(getenv("BT_DIRECT_SLOTS")) : 6;
    e.times = getenv("BT_DIRECT_TIMES") != nullptr;
    return e;
  };
  static const DrEnv env0 = read_env();
  const DrEnv env = dyn_env ? read_env() : env0;
  const bool dr_disabled = env.disabled;   // A/B switch
  const bool dr_force = env.force;         // tests: take it whenever it is legal
  int dr = 0, dr_x = 1, dr_smem = 0, dr_ns = 2, dr_stage = 0;
  {
    const bool same = p.sd == 1 && p.sh == 1 && p.sw == 1 && p.ID == p.OD && p.IH == p.OH && p.IW == p.OW &&
                      p.groups == 1 && p.Cin_g % BLOCK_K == 0;
    const long long Pw = p.IW + p.pw, Ph = p.IH + p.ph, Pd = p.ID + p.pd;
    const long long Mp = (long long)p.B * Pd * Ph * Pw;
    const long long halo = ((long long)p.pd * Ph + p.ph) * Pw + p.pw;
    if (fast && same && p.x_is_bf16 && !dr_disabled && Mp < (1ll << 31) && halo <= 512 && p.num_kb <= 64) {
      const int R = (int)((BLOCK_M + 2 * halo + 7) / 8 * 8);
      const int slabs = p.Cin_g / BLOCK_K;
      const long long n_rt = (Mp + BLOCK_M - 1) / BLOCK_M;
      const int bns[3] = {128, 64, 32};
      const int bn_only = env.bn_only, x_only = env.x_only, slots_max = env.slots_max;          // diagnostics
      double dbest = 1e300;
      for (int bi = 0; bi < 3; ++bi) {
        const int bn = bns[bi];
        if (bn_only && bn != bn_only) continue;
        if (bn > 32 && bn / 2 >= p.N) continue;
        const long long res = (long long)p.num_kb * NB * bn * 128;
        const long long slot = (long long)NB * slabs * R * 128;
        if (2 * NB * bn > 512) continue;
        const long long nt = (p.N + bn - 1) / bn;
        // one K=16 MMA: 128*bn/256 tensor clocks, but never less than ~48 (operand reads from smem: 6 KB at N = 64)
        const double t_mma = (double)p.num_kb * NB * 4.0 * (0.5 * bn + 8.0 > 48.0 ? 0.5 * bn + 8.0 : 48.0);
        const double t_prod = (double)R * slabs * (flip ? 20.0 : 14.0);        // 7 producer warps, latency-bound chains
        // with / without the epilogue staging buffer (8 warps x 32 rows x bn/2 bf16): staging buys coalesced global
        // access, dropping it buys window slots (prefetch depth D = slots - 2)
        for (int stg_on = 1; stg_on >= 0; --stg_on) {
          const long long stage_b = stg_on ? 256ll * bn : 0;
          long long ns = (SMEM_BUDGET - DR_AUX_BYTES - stage_b - 1024 - res) / slot;
          if (ns > slots_max) ns = slots_max;
          if (ns > 8) ns = 8;
          if (ns < 2) continue;
          const long long need = res + ns * slot + DR_AUX_BYTES + stage_b + 1024;
          const double t_epi = (bn * (flip ? 24.0 : 16.0) + 600.0) * (stg_on ? 1.0 : 1.25);
          // a window needs ~2500 clocks from "slot free" to "landed and published"; D of them overlap
          const double t_lat = 2500.0 / (double)(ns >= 3 ? ns - 2 : 1);
          double t_tile = t_mma > t_prod ? t_mma : t_prod;
          if (t_epi > t_tile) t_tile = t_epi;
          if (t_lat > t_tile) t_tile = t_lat;
          t_tile *= 1.3;                                   // the roles share the issue slots of one SM
          const long long xmax = n_rt < 4 * sm_count ? n_rt : 4 * sm_count;
          for (long long x = 1; x <= xmax; ++x) {
            if (x_only && x != x_only) continue;
            const long long ctas = x * nt * p.S;
            const double waves = (double)((ctas + sm_count - 1) / sm_count);
            const double per = (double)((n_rt + x - 1) / x);
            // sampling prologue (measured, tools/direct_probe.py): ~500 + 25 bn clocks per k-block
            const double t_cta = p.num_kb * (500.0 + 25.0 * bn) + per * (t_tile + 100.0) + 6000.0;
            if (waves * t_cta < dbest) {
              dbest = waves * t_cta;
              dr = bn; dr_x = (int)x; dr_smem = (int)need; dr_ns = (int)ns; dr_stage = stg_on;
            }
          }
        }
      }
      // (the im2col-path estimates in `best` are ~2x optimistic against measurements, profiles/r01h_direct_probe.log:
      //  only drop the direct kernel when it looks clearly worse)
      if (dr && !dr_force && dbest >= 1.5 * best) dr = 0;
      if (dr) {
        BN = dr;
        p.dr_R = R; p.dr_halo = (int)halo;
        p.dr_Pw = (int)Pw; p.dr_Ph = (int)Ph; p.dr_Pd = (int)Pd;
        p.dr_Mp = Mp;
        p.dr_slots = dr_ns;
        p.dr_stage = dr_stage;
        {
          const int slabs = p.Cin_g / BLOCK_K;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const int t = kb / slabs, sl = kb - t * slabs;
            const uint32_t tp = p.taps[t];
            const int kd = tp & 0xff, kh = (tp >> 8) & 0xff, kw = (tp >> 16) & 0xff;
            const long long delta = ((long long)(kd * p.dd - p.pd) * Ph + (kh * p.dh - p.ph)) * Pw + (kw * p.dw - p.pw);
            p.dr_aoff[kb] = (int)(((long long)sl * R + halo + delta) * 8);
          }
        }
        const long long divs[3] = {Pw, Ph, Pd};
        for (int i = 0; i < 3; ++i) {   // n / d == (n * mul) >> sh for every n < 2^31 (l = ceil(log2 d), mul = ceil(2^(31+l) / d))
          int l = 0;
          while ((1ll << l) < divs[i]) ++l;
          p.dr_sh[i] = (uint32_t)(31 + l);
          p.dr_mul[i] = (uint32_t)((((unsigned long long)1 << (31 + l)) + (unsigned long long)divs[i] - 1) / (unsigned long long)divs[i]);
        }
        p.dr_times = (env.times && workspace != nullptr) ? static_cast<long long*>(workspace) : nullptr;
      }
    }
  }
  // (d) TMA mode (bt_tma_kernel, bt_tma.cuh): Reparameterization layers whose activation operand is one TMA box per
  // (row tile, k-block) -- linear layers, materialised-im2col stems, and every convolution with C_in/groups a multiple
  // of the k-block (im2col tensor map: stride / padding / dilation are in the map) -- and whose sampled tile fits shared
  // memory next to >= 3 stages.  Takes over from the cp.async families (bt_ws_kernel, bt_fused_kernel fast path);
  // the direct kernel keeps the stride-1 "same" convolutions it was selected for (it reads every activation once
  // instead of once per tap), unless BT_TMA_PREFER is set.
  int tm = 0, tm_x = 1, tm_stages = 0, tm_smem = 0, tm_stream = 0, tm_mt = 1, tm_nsmp = 1, tm_epst = 0, tm_cln = 1;
  TmaAPlan tma_a;
  memset(&tma_a, 0, sizeof(tma_a));
  {
    struct TmEnv { bool disabled, prefer, cln_off, cln_force; int bn_only, mode_only; };
    auto read_tm = []() {
      TmEnv e;
      e.disabled = getenv("BT_DISABLE_TMA") != nullptr;
      e.prefer = getenv("BT_TMA_PREFER") != nullptr;
      e.cln_off = getenv("BT_DISABLE_CLUSTER") != nullptr;
      e.cln_force = getenv("BT_FORCE_CLUSTER") != nullptr;
      e.bn_only = getenv("BT_TMA_BN") ? atoi(getenv("BT_TMA_BN")) : 0;
      e.mode_only = getenv("BT_TMA_MODE") ? atoi(getenv("BT_TMA_MODE")) : 0;     // 1: resident only, 2: streaming only
      return e;
    };
    static const TmEnv tenv0 = read_tm();
    const TmEnv tenv = dyn_env ? read_tm() : tenv0;
    const bool base_ok = p.w_vec && dbg_any == 0 && kl_out == nullptr && n_used <= 64 && !p.transposed &&
                         (p.taps_explicit || taps_all == 1) && p.K_phys % 8 == 0 && p.Cin_g % 8 == 0 && !tenv.disabled &&
                         (long long)(p.x_shared ? 1 : p.S) * p.B * in_sp < (1ll << 31) && p.M < (1ll << 31) &&
                         (p.x_is_bf16 || tf32) && (!dr || tenv.prefer) && (plan_only || al16(x));
    if (base_ok && tma_a_plan(p, tf32, &tma_a) && (plan_only || tma_driver_ready())) {
      const int kbe = tma_a.kbe;
      const int nkb = tma_a.mode == 1 ? (p.K_used + kbe - 1) / kbe : p.num_kb;
      const long long n_rt = m_tiles;
      const int bns[3] = {128, 64, 32};
      // sampler clocks per weight element with 256 threads, everything included (measured on the layer3 / layer4
      // launches, profiles/r02f: 0.47-0.53 with the sigma cache; softplus adds 2 MUFU + ~10 ALU, the tf32 one more)
      const double c_el = p.rho_is_sigma ? 0.5 : (tf32 ? 0.8 : 0.65);
      const long long ep_bytes_of[3] = {128ll * 128 * x_es, 128ll * 64 * x_es, 128ll * 32 * x_es};   // epilogue staging, per bn
      const double l2_bpc = 16.0;                           // L2 -> SM bytes per clock per SM with every SM pulling (measured:
                                                            // ~5-6 TB/s chip-wide on the im2col re-reads, profiles/r02)
      double tbest = 1e300;
      for (int bi = 0; bi < 3; ++bi) {
        const int bn = bns[bi];
        if (tenv.bn_only && bn != tenv.bn_only) continue;
        if (bn > 32 && bn / 2 >= p.N) continue;
        const long long nt = (long long)((p.N + bn - 1) / bn) * p.groups;
        const double mma1 = 43.0 + 0.5 * bn;   // clocks per tcgen05.mma, both operands in smem (tools/probes/mma_probe.cu)
        const double t_epi = bn * 5.0 + 300.0;
        // (1) resident W_s, row tiles streamed past it.  When every MC sample reads the SAME x (first layer of an MC
        // pass) a CTA keeps the sampled tiles of `nsmp` samples and multiplies each staged activation tile with all of
        // them: 1/nsmp of the L2 -> SM traffic (the stem is L2-bound otherwise: 64 samples re-read one small matrix).
        for (int nsmp = p.x_shared ? TM_MAX_NSMP : 1; nsmp >= 1 && tenv.mode_only != 2 && !flip; nsmp >>= 1) {
          if (nsmp > p.S || 2 * nsmp * bn > 512) continue;
          const long long res = (long long)nsmp * nkb * bn * 128;
          int epst = 1;                                     // epilogue staging buffer when >= 3 stages remain
          // fused max-pool: the staging buffer is the CTA-wide tile buffer, plus two carry rows per sample
          const long long pool_b = want_pool ? (long long)nsmp * 2 * p.pool_ow * bn * x_es : 0;
          if (want_pool && !(pool_geom_ok && bn * x_es >= 128 && p.N % bn == 0 && p.ep_residual == nullptr)) continue;
          long long stg = (SMEM_BUDGET - TM_AUX_BYTES - 1024 - res - ep_bytes_of[bi] - pool_b) / A_TILE_BYTES;
          if (stg < 3 && want_pool) continue;
          if (stg < 3) {
            stg = (SMEM_BUDGET - TM_AUX_BYTES - 1024 - res) / A_TILE_BYTES;
            epst = 0;
          }
          if (stg > MAX_STAGES) stg = MAX_STAGES;
          if (stg < 3) continue;
          const double t_mma = nsmp * (nkb * 4.0 * mma1) + 100.0;
          const double t_l2 = nkb * (double)A_TILE_BYTES / l2_bpc;
          double t_tile = t_mma > t_l2 ? t_mma : t_l2;
          if (nsmp * t_epi * (epst ? 1.0 : 2.0) > t_tile) t_tile = nsmp * t_epi * (epst ? 1.0 : 2.0);
          const double t_samp = nsmp * nkb * (400.0 + bn * kbe * c_el);
          const long long zs = (p.S + nsmp - 1) / nsmp;
          const long long tpi = want_pool ? (long long)p.pool_oh * p.pool_ow / BLOCK_M : 1;   // tiles per unit of work
          const long long n_units = n_rt / tpi;
          const long long xmax = n_units < 4 * sm_count ? n_units : 4 * sm_count;
          for (long long x_ = 1; x_ <= xmax; ++x_) {
            const long long ctas = x_ * nt * zs;
            const double waves = (double)((ctas + sm_count - 1) / sm_count);
            const double per = (double)(((n_units + x_ - 1) / x_) * tpi);
            const double t_cta = t_samp + per * t_tile * 1.1 + 5000.0;
            if (waves * t_cta < tbest) {
              tbest = waves * t_cta;
              tm = bn; tm_x = (int)x_; tm_stages = (int)stg; tm_stream = 0; tm_mt = 1; tm_nsmp = nsmp; tm_epst = epst;
              tm_cln = 1;
              tm_smem = (int)(res + stg * A_TILE_BYTES + TM_AUX_BYTES + (epst ? ep_bytes_of[bi] : 0) + pool_b + 1024);
            }
          }
        }
        // (2) streaming: a sampled [bn x kbe] tile per k-block, shared by MT row tiles (Flipout: two weight tiles, two
        // activation planes and two accumulators per row tile; the transform warps build the x * s_in plane)
        if (tenv.mode_only == 1 || (flip && bn < 64) || want_pool) continue;
        for (int mt = 1; mt <= 4; mt <<= 1) {
          if (mt * NB * bn > 512) break;
          if (mt > 1 && mt / 2 >= n_rt) break;
          const long long stage_b = (long long)NB * ((long long)bn * 128 + (long long)mt * A_TILE_BYTES);
          int epst2 = 1;
          long long stg2 = (SMEM_BUDGET - TM_AUX_BYTES - 1024 - ep_bytes_of[bi]) / stage_b;
          if (stg2 < 3 && stg2 < nkb) {
            const long long alt = (SMEM_BUDGET - TM_AUX_BYTES - 1024) / stage_b;
            if (alt > stg2) { stg2 = alt; epst2 = 0; }
          }
          if (stg2 > MAX_STAGES) stg2 = MAX_STAGES;
          if (stg2 > nkb) stg2 = nkb;
          if (stg2 < 2 && nkb >= 2) continue;
          const long long groups_m0 = (n_rt + mt - 1) / mt;
          // the 16 B/clk/SM were measured with every SM pulling; a launch that leaves SMs idle (small S: one rank of an
          // N-GPU job) gives each busy SM a larger share of the L2 -> SM fabric (capped at 3x)
          const double busy = (double)(groups_m0 * nt * p.S) / sm_count;
          const double l2_eff = l2_bpc * (busy >= 1.0 ? 1.0 : (busy <= 1.0 / 3.0 ? 3.0 : 1.0 / busy));
          // A-operand multicast: two consecutive n-tile CTAs of a group form a cluster, each loads half of the row tiles
          // and multicasts them -- the L2 -> SM traffic per CTA halves (tenv.cln_off: A/B switch BT_DISABLE_CLUSTER)
          const int nt_g = (p.N + bn - 1) / bn;
          // Used when it was measured to pay (profiles/r02_log.md, call V): >= 2 row tiles per k-block (so both CTAs load)
          // and a single wave of CTAs -- with one row tile only rank 0 loads and the pair just runs in lock-step (layer4:
          // +5%), and with more CTAs than SMs the pairwise scheduling costs more than the traffic saves (4096^3: +25%).
          const long long ctas0 = groups_m0 * nt * p.S;
          const int cln = (!tenv.cln_off && nt_g % 2 == 0 && !flip && ((mt >= 2 && ctas0 <= sm_count) || tenv.cln_force)) ? 2 : 1;
          const double t_s = 400.0 + bn * kbe * c_el * (flip ? 1.1 : 1.0), t_m = NB * mt * 4.0 * mma1,
                       t_l = mt * (double)A_TILE_BYTES / l2_eff / cln + (cln > 1 ? 150.0 : 0.0),
                       // Flipout: the four transform warps build the x * s_in plane of every row tile -- one warp per
                       // scheduler, latency-bound: ~2000 clocks per row tile and k-block (measured with the sampler
                       // arithmetic switched off, BT_TMA_PROBE=1: C5 4096^3 bn 128 / mt 2 = 4245 clocks per k-block)
                       t_x = flip ? 300.0 + mt * 1900.0 : 0.0;
          double t_kb = t_s > t_m ? t_s : t_m;
          if (t_l > t_kb) t_kb = t_l;
          if (t_x > t_kb) t_kb = t_x;
          if (stg2 < 3) t_kb *= 1.1;
          const long long groups_m = (n_rt + mt - 1) / mt;
          const long long ctas = groups_m * nt * p.S;
          const double waves = (double)((ctas + sm_count - 1) / sm_count);
          const double t_cta = nkb * t_kb * 1.1 + mt * t_epi + 5000.0;
          if (waves * t_cta < tbest) {
            tbest = waves * t_cta;
            tm = bn; tm_x = (int)groups_m; tm_stages = (int)stg2; tm_stream = 1; tm_mt = mt; tm_nsmp = 1; tm_epst = epst2;
            tm_cln = cln;
            tm_smem = (int)(stg2 * stage_b + TM_AUX_BYTES + (epst2 ? ep_bytes_of[bi] : 0) + 1024);
          }
        }
      }
      if (tm) {
        dr = 0;
        BN = tm;
        p.num_kb = nkb;
      }
    }
  }
  // (e) TMA direct mode (bt_dtma_kernel, bt_tma.cuh): the direct kernel's in-place A operand with the input windows
  // staged by tiled TMA boxes (zero padding = out-of-range fill) -- bf16 and tf32.  Takes over the stride-1 "same"
  // convolutions with more than one filter tap from both the cp.async direct kernel and the im2col TMA kernels (which
  // re-read every activation once per tap from L2).
  int dtm = 0, dtm_x = 1, dtm_smem = 0, dtm_epst = 0, dtm_clx = 1;
  DtGeom dtg;
  memset(&dtg, 0, sizeof(dtg));
  {
    struct DtEnv { bool disabled; int bn_only; };
    auto read_dt = []() {
      DtEnv e;
      e.disabled = getenv("BT_DISABLE_DTMA") != nullptr || getenv("BT_DISABLE_TMA") != nullptr;
      e.bn_only = getenv("BT_DTMA_BN") ? atoi(getenv("BT_DTMA_BN")) : 0;
      return e;
    };
    static const DtEnv denv0 = read_dt();
    const DtEnv denv = dyn_env ? read_dt() : denv0;
    const bool ok = p.w_vec && dbg_any == 0 && kl_out == nullptr && n_used <= 64 && n_used > 1 && !p.transposed &&
                    p.taps_explicit && p.Cin_g % 8 == 0 && !denv.disabled && !dr_force && !want_pool && (p.x_is_bf16 || tf32) && (plan_only || al16(x)) &&
                    (long long)(p.x_shared ? 1 : p.S) * p.B < (1ll << 31) && p.M < (1ll << 31) && (plan_only || tma_driver_ready());
    if (ok) {
      const int kbe = p.x_is_bf16 ? 64 : 32;
      const int nkb = n_used * (p.Cin_g / kbe);
      double dbest = 1e300;
      const int bns[3] = {128, 64, 32};
      for (int bi = 0; bi < 3 && p.Cin_g % kbe == 0; ++bi) {
        const int bn = bns[bi];
        if (denv.bn_only && bn != denv.bn_only) continue;
        if (bn > 32 && bn / 2 >= p.N) continue;
        DtGeom g;
        int sm = 0, epst = 0;
        if (2 * NB * bn > 512) continue;
        if (!dt_plan(p, tf32, flip, bn, nkb, &g, &sm, &epst)) continue;
        const long long n_rt = (g.NR + g.k - 1) / g.k;
        const long long nt = (p.N + bn - 1) / bn;
        const double mma1 = 43.0 + 0.5 * bn;   // clocks per tcgen05.mma, both operands in smem (tools/probes/mma_probe.cu)
        const double t_mma = NB * nkb * 4.0 * mma1 + 100.0, t_epi = (bn * (flip ? 8.0 : 5.0) + 300.0) * (epst ? 1.0 : 2.0),
                     t_tma = g.nbox * (p.Cin_g / kbe) * 350.0 + 300.0,
                     t_xf = flip ? 400.0 + (g.k + 2.0 * g.hr) * g.Pw * (p.Cin_g / 64.0) * 4.0 : 0.0;
        double t_tile = t_mma > t_epi ? t_mma : t_epi;
        if (t_tma > t_tile) t_tile = t_tma;
        if (t_xf > t_tile) t_tile = t_xf;
        if (g.slots < 3) t_tile *= 1.25;
        const double t_samp = nkb * (400.0 + bn * kbe * (p.rho_is_sigma ? 0.5 : (tf32 ? 0.8 : 0.65)) * (flip ? 1.3 : 1.0));
        const long long xmax = n_rt < 4 * sm_count ? n_rt : 4 * sm_count;
        for (long long x_ = 1; x_ <= xmax; ++x_) {
          const long long ctas = x_ * nt * p.S;
          const double waves = (double)((ctas + sm_count - 1) / sm_count);
          const double per = (double)((n_rt + x_ - 1) / x_);
          // the x_ CTAs of a (sample, n-tile) in clusters of clx (largest power of two <= 8 dividing x_, <= nkb): each samples
          // 1/clx of the k-blocks and writes them to all (DSMEM) instead of every CTA sampling the whole W_s
          // MEASURED SLOWER than every CTA sampling the whole W_s (profiles/r02_log.md, call X: layer1 family +2% at S = 64
          // with clusters of 2, S = 8 per rank 0.446 -> 0.528 ms with clusters of 8): off unless BT_CLUSTER_PROLOGUE is set.
          int clx = 1;
          static const bool clx_on0 = getenv("BT_CLUSTER_PROLOGUE") != nullptr;
          const bool clx_on = dyn_env ? getenv("BT_CLUSTER_PROLOGUE") != nullptr : clx_on0;
          while (clx_on && clx < 8 && x_ % (2 * clx) == 0 && 2 * clx <= nkb) clx *= 2;
          const double t_cta = t_samp / clx + (clx > 1 ? 800.0 + 0.1 * t_samp : 0.0) + per * t_tile * 1.1 + 5000.0;
          if (waves * t_cta < dbest) {
            dbest = waves * t_cta;
            dtm = bn; dtm_x = (int)x_; dtm_smem = sm; dtg = g; dtm_epst = epst; dtm_clx = clx;
          }
        }
      }
      if (dtm) {
        dr = 0; tm = 0;
        BN = dtm;
        p.num_kb = nkb;
        const int slabs = p.Cin_g / kbe;
        for (int kb = 0; kb < nkb; ++kb) {
          const int t = kb / slabs, sl = kb - t * slabs;
          const uint32_t tp_ = p.taps[t];
          const int kd = tp_ & 0xff, kh = (tp_ >> 8) & 0xff, kw = (tp_ >> 16) & 0xff;
          const long long delta = ((long long)(kd * p.dd - p.pd) * dtg.Ph + (kh * p.dh - p.ph)) * dtg.Pw + (kw * p.dw - p.pw);
          p.dr_aoff[kb] = (int)(((long long)sl * dtg.R + dtg.Z + (long long)dtg.hr * dtg.Pw + delta) * 8);   // (copy 0)
        }
      }
    }
  }
  p.n_tiles_per_group = (p.N + BN - 1) / BN;
  const long long n_tiles = (long long)p.n_tiles_per_group * p.groups;
  BT_REQUIRE(n_tiles <= 65535, BT_ERR_BAD_SHAPE, "bt_layer_forward: too many N tiles");
  static const bool async_disabled = getenv("BT_DISABLE_ASYNC") != nullptr;   // A/B switch
  p.ws_async = async_disabled ? 0 : 1;
  p.MT = mt;
  p.ws = ws;
  p.n_groups = (int)((m_tiles + mt - 1) / mt);
  const int stage_bytes = ws ? NB * mt * A_TILE_BYTES : NB * (BN * 128 + mt * A_TILE_BYTES);   // (ws == 2: 16 KB)
  const int res_total = ws ? p.num_kb * NB * BN * 128 : 0;
  const int tc_total = ws == 2 ? (int)tc_bytes_sel : 0;
  p.tc_rows = ws == 2 ? tc_rows_sel : 0;
  p.tc_halo = tc_halo_sel;
  p.tc_padoff = tc_padoff_sel;
  int stages = (tm || dtm) ? 1 : (SMEM_BUDGET - AUX_BYTES - 1024 - res_total - tc_total) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (!ws && stages > p.num_kb) stages = p.num_kb < 1 ? 1 : p.num_kb;   // (the ws ring runs across M-groups)
  BT_REQUIRE(dr || tm || dtm || stages >= 1, BT_ERR_UNSUPPORTED, "bt_layer_forward: tile does not fit shared memory");
  p.stages = stages;
  const int smem_bytes = res_total + stages * stage_bytes + tc_total + AUX_BYTES + 1024;
  uint32_t cols = (uint32_t)(ws == 2 ? 2 * BN : NB * mt * BN), pc = 32;   // ws == 2: two accumulator buffers
  while (pc < cols) pc <<= 1;
  p.tmem_cols = pc;

  p.prior_mu = prior_mu_s;
  if (kl_out != nullptr) {
    BT_REQUIRE(n_tiles <= 65536, BT_ERR_UNSUPPORTED, "bt_layer_forward: KL partial buffer too small");
    p.kl_partials = static_cast<float*>(workspace);
    p.log_prior_sigma = logf(prior_sigma_s);
    p.inv_2ps2 = 0.5f / (prior_sigma_s * prior_sigma_s);
  }
  p.key.k0 = (uint32_t)(seed & 0xffffffffu);
  p.key.k1 = (uint32_t)(seed >> 32);
  p.key.c3_base = layer_key << 4;
  p.sample0 = sample_idx0;

  const long long gx = ws ? ws_x : (m_tiles + mt - 1) / mt;
  BT_REQUIRE(gx < (1ll << 31), BT_ERR_BAD_SHAPE, "bt_layer_forward: grid too large");
  dim3 grid((unsigned)gx, (unsigned)n_tiles, (unsigned)p.S);
  const bool pool_fused = want_pool && tm && !tm_stream && !dtm;
  if (plan_only) {   // report the decision (the launch below does exactly this)
    memset(plan, 0, sizeof(*plan));
    plan->pool_fused = pool_fused ? 1 : 0;
    plan->path = dtm ? BT_PATH_TMA_DIRECT : tm ? (tm_stream ? BT_PATH_TMA_STREAM : BT_PATH_TMA) : (dr ? BT_PATH_DIRECT : (ws == 2 ? BT_PATH_WS : (fast ? (ws ? BT_PATH_FAST_WS : BT_PATH_FAST) : BT_PATH_GENERIC)));
    plan->block_n = BN;
    plan->k_blocks = p.num_kb;
    if (dtm) {
      uint32_t tcols = (uint32_t)(2 * NB * dtm), tpc = 32;
      while (tpc < tcols) tpc <<= 1;
      plan->m_subtiles = 1;
      plan->grid[0] = dtm_x; plan->grid[1] = (int32_t)n_tiles; plan->grid[2] = p.S;
      plan->threads = (tf32 || flip) ? tm_threads<true>() : tm_threads<false>();
      plan->smem_bytes = dtm_smem;
      plan->tmem_cols = (int32_t)tpc;
      plan->window_slots = dtg.slots; plan->window_rows = dtg.R; plan->staged_epilogue = dtm_epst;
      plan->window_boxes = dtg.nbox;
      plan->cluster_n = dtm_clx;
    } else if (tm) {
      uint32_t tcols = (uint32_t)(tm_stream ? NB * tm_mt * tm : 2 * tm_nsmp * tm), tpc = 32;
      while (tpc < tcols) tpc <<= 1;
      plan->m_subtiles = tm_mt;
      plan->grid[0] = tm_x; plan->grid[1] = (int32_t)n_tiles; plan->grid[2] = (p.S + tm_nsmp - 1) / tm_nsmp;
      plan->samples_per_cta = tm_nsmp;
      plan->cluster_n = tm_stream ? tm_cln : 1;
      plan->staged_epilogue = tm_epst;
      plan->threads = (tf32 || flip) ? tm_threads<true>() : tm_threads<false>();
      plan->smem_bytes = tm_smem;
      plan->tmem_cols = (int32_t)tpc;
      plan->window_slots = tm_stages;
    } else if (dr) {
      uint32_t dcols = (uint32_t)(2 * NB * dr), dpc = 32;
      while (dpc < dcols) dpc <<= 1;
      plan->m_subtiles = 1;
      plan->grid[0] = dr_x; plan->grid[1] = (int32_t)n_tiles; plan->grid[2] = p.S;
      plan->threads = DR_THREADS;
      plan->smem_bytes = dr_smem;
      plan->tmem_cols = (int32_t)dpc;
      plan->window_slots = p.dr_slots; plan->window_rows = p.dr_R; plan->staged_epilogue = p.dr_stage;
    } else {
      plan->m_subtiles = mt;
      plan->grid[0] = (int32_t)gx; plan->grid[1] = (int32_t)n_tiles; plan->grid[2] = p.S;
      plan->threads = ws == 2 ? WS_THREADS : (fast ? FAST_WARPS : GENERIC_WARPS) * 32 + 32;
      plan->smem_bytes = smem_bytes;
      plan->tmem_cols = (int32_t)p.tmem_cols;
    }
    return BT_OK;
  }
  BT_REQUIRE(!want_pool || pool_fused, BT_ERR_UNSUPPORTED,
             "bt_layer_forward: no kernel with a fused max-pool for this layer (BtForwardPlan.pool_fused == 0): "
             "leave BtLayerGeom.pool_hw zero and pool separately");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtm) {
    DtParams dp;
    p.MT = 1; p.ws = 0; p.tc_rows = 0; p.stages = 0;
    p.dr_stage = dtm_epst;
    p.n_groups = (int)((dtg.NR + dtg.k - 1) / dtg.k);
    uint32_t tcols = (uint32_t)(2 * NB * dtm), tpc = 32;
    while (tpc < tcols) tpc <<= 1;
    p.tmem_cols = tpc;
    if ((rc = dt_encode(p, dtg, x, &dp.map_a, false)) != BT_OK) return rc;
    if ((rc = dt_encode(p, dtg, x, &dp.map_h, true)) != BT_OK) return rc;
    dp.f = p;
    dp.g = dtg;
    dp.kbe = p.x_is_bf16 ? 64 : 32;
    dp.slabs = p.Cin_g / dp.kbe;
    dp.clx = dtm_clx;
    dim3 dgrid((unsigned)dtm_x, (unsigned)n_tiles, (unsigned)p.S);
    rc = bt_tma_family_launch(2, &dp, dtm, tf32 ? 1 : 0, flip ? 1 : 0, dgrid.x, dgrid.y, dgrid.z, dtm_smem, dev, stream);
  } else if (tm) {
    TmaParams tp;
    p.MT = tm_mt; p.ws = 0; p.tc_rows = 0;
    p.dr_stage = tm_epst;
    p.stages = tm_stages;
    p.n_groups = (int)m_tiles;
    uint32_t tcols = (uint32_t)(tm_stream ? NB * tm_mt * tm : 2 * tm_nsmp * tm), tpc = 32;
    while (tpc < tcols) tpc <<= 1;
    p.tmem_cols = tpc;
    if ((rc = tma_encode_a(p, tma_a, x, &tp.map_a)) != BT_OK) return rc;
    tp.a.nsmp = tm_nsmp;
    tp.a.cln = tm_stream ? tm_cln : 1;
    tp.a.probe = dyn_env && getenv("BT_TMA_PROBE") ? atoi(getenv("BT_TMA_PROBE")) : 0;
    tp.f = p;
    tp.a.mode = tma_a.mode; tp.a.nd = tma_a.nd; tp.a.kbe = tma_a.kbe;
    tp.a.slabs = tma_a.mode == 2 ? p.Cin_g / tma_a.kbe : p.num_kb;
    dim3 tgrid((unsigned)tm_x, (unsigned)n_tiles, (unsigned)((p.S + tm_nsmp - 1) / tm_nsmp));
    rc = bt_tma_family_launch(tm_stream ? 1 : 0, &tp, tm, tf32 ? 1 : 0, flip ? 1 : 0, tgrid.x, tgrid.y, tgrid.z, tm_smem, dev, stream);
  } else if (dr) {
    p.MT = 1; p.ws = 0; p.tc_rows = 0; p.stages = 0;
    p.n_groups = (int)((p.dr_Mp + BLOCK_M - 1) / BLOCK_M);
    uint32_t dcols = (uint32_t)(2 * NB * dr), dpc = 32;
    while (dpc < dcols) dpc <<= 1;
    p.tmem_cols = dpc;
    dim3 dgrid((unsigned)dr_x, (unsigned)n_tiles, (unsigned)p.S);
    if (dr == 128) rc = dispatch_direct<128>(p, flip, dgrid, dr_smem, dev, st);
    else if (dr == 64) rc = dispatch_direct<64>(p, flip, dgrid, dr_smem, dev, st);
    else rc = dispatch_direct<32>(p, flip, dgrid, dr_smem, dev, st);
  } else if (ws == 2) {
    if (BN == 64) rc = p.p_is_bf16 ? launch_ws<64, true>(p, grid, smem_bytes, dev, st)
                                   : launch_ws<64, false>(p, grid, smem_bytes, dev, st);
    else rc = p.p_is_bf16 ? launch_ws<128, true>(p, grid, smem_bytes, dev, st)
                          : launch_ws<128, false>(p, grid, smem_bytes, dev, st);
  } else if (BN == 64) rc = flip ? dispatch_fused<64, true>(p, fast, tf32, grid, smem_bytes, dev, st)
                          : dispatch_fused<64, false>(p, fast, tf32, grid, smem_bytes, dev, st);
  else rc = flip ? dispatch_fused<128, true>(p, fast, tf32, grid, smem_bytes, dev, st)
                 : dispatch_fused<128, false>(p, fast, tf32, grid, smem_bytes, dev, st);
  if (rc != BT_OK) return rc;
  g_last_path = dtm ? BT_PATH_TMA_DIRECT : tm ? (tm_stream ? BT_PATH_TMA_STREAM : BT_PATH_TMA) : (dr ? BT_PATH_DIRECT : (ws == 2 ? BT_PATH_WS : (fast ? (ws ? BT_PATH_FAST_WS : BT_PATH_FAST) : BT_PATH_GENERIC)));
  if (kl_out != nullptr) {
    bt_fused_kl_finalize<<<1, 32, 0, st>>>(p.kl_partials, (int)n_tiles, (long long)p.C_out * p.K_phys, mu_b,
                                           rho_b, mu_b ? p.C_out : 0, p.p_is_bf16, prior_mu_s,
                                           p.log_prior_sigma, p.inv_2ps2, kl_out);
    BT_CHECK_CUDA(cudaGetLastError());
  }
  return BT_OK;
}

int bt_layer_forward(int mode, const BtLayerGeom* gm, const void* x, int x_dtype, const void* mu_w,
                     const void* rho_w, const void* mu_b, const void* rho_b, int p_dtype, void* out,
                     float* kl_out, float prior_mu_s, float prior_sigma_s, uint64_t seed,
                     uint32_t layer_key, uint32_t sample_idx0, const BtDebugIO* dbg,
                     const BtEpilogue* epi, void* workspace, void* stream) {
  return layer_forward_impl(nullptr, 0, mode, gm, x, x_dtype, mu_w, rho_w, mu_b, rho_b, p_dtype, out, kl_out, prior_mu_s,
                            prior_sigma_s, seed, layer_key, sample_idx0, dbg, epi, workspace, stream);
}

// Test hook: stage ONE activation tile (128 output rows from row m0 of MC sample `sample`, group `group`, filter tap
// number `tap` in (kd, kh, kw) order, 64/32-channel slab `slab`) through exactly the tensor map and TMA instruction
// bt_tma_kernel uses, and copy the 16 KB shared-memory image (128B-swizzled rows) to `out`.
int bt_tma_probe(const BtLayerGeom* gm, const void* x, int x_dtype, int64_t m0, int sample, int group, int tap, int slab,
                 void* out, void* stream) {
  BT_REQUIRE(gm != nullptr && x != nullptr && out != nullptr, BT_ERR_BAD_POINTER, "bt_tma_probe: NULL argument");
  int rc;
  if ((rc = bt_device_check()) != BT_OK) return rc;
  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.S = gm->n_samples; p.x_shared = gm->x_shared ? 1 : 0; p.B = gm->batch;
  p.C_in = gm->c_in; p.C_out = gm->c_out; p.groups = gm->groups;
  p.Cin_g = gm->c_in / gm->groups; p.N = gm->c_out / gm->groups;
  p.ID = gm->in_dhw[0]; p.IH = gm->in_dhw[1]; p.IW = gm->in_dhw[2];
  p.OD = gm->out_dhw[0]; p.OH = gm->out_dhw[1]; p.OW = gm->out_dhw[2];
  p.KD = gm->k_dhw[0]; p.KH = gm->k_dhw[1]; p.KW = gm->k_dhw[2];
  p.sd = gm->stride[0]; p.sh = gm->stride[1]; p.sw = gm->stride[2];
  p.pd = gm->pad[0]; p.ph = gm->pad[1]; p.pw = gm->pad[2];
  p.dd = gm->dil[0]; p.dh = gm->dil[1]; p.dw = gm->dil[2];
  p.x_is_bf16 = x_dtype == BT_BF16;
  p.M = (long long)p.B * p.OD * p.OH * p.OW;
  int cnt = 0;
  for (int kd = 0; kd < p.KD; ++kd)
    for (int kh = 0; kh < p.KH; ++kh)
      for (int kw = 0; kw < p.KW; ++kw)
        if (cnt < MAX_TAPS) p.taps[cnt++] = (uint32_t)kd | ((uint32_t)kh << 8) | ((uint32_t)kw << 16);
  p.taps_explicit = 1;
  TmaAPlan a;
  BT_REQUIRE(tma_a_plan(p, !p.x_is_bf16, &a), BT_ERR_UNSUPPORTED, "bt_tma_probe: geometry has no TMA form");
  BT_REQUIRE(tap >= 0 && tap < cnt && m0 >= 0 && m0 < p.M, BT_ERR_BAD_SHAPE, "bt_tma_probe: tap / row out of range");
  TmaParams tp;
  if ((rc = tma_encode_a(p, a, x, &tp.map_a)) != BT_OK) return rc;
  tp.f = p;
  tp.a.mode = a.mode; tp.a.nd = a.nd; tp.a.kbe = a.kbe; tp.a.slabs = a.mode == 2 ? p.Cin_g / a.kbe : 1;
  tp.a.nsmp = 1;
  return bt_tma_probe_launch(&tp, (long long)m0, sample, group, tap, slab, out, stream);
}

int bt_tma_probe4d(const void* x, int x_dtype, const int64_t* dims, const int32_t* box, const int32_t* coords,
                   uint32_t dst_off, void* out, void* stream) {
  BT_REQUIRE(x != nullptr && dims != nullptr && box != nullptr && coords != nullptr && out != nullptr, BT_ERR_BAD_POINTER,
             "bt_tma_probe4d: NULL argument");
  int rc;
  if ((rc = bt_device_check()) != BT_OK) return rc;
  BT_REQUIRE(tma_driver_ready(), BT_ERR_UNSUPPORTED, "TMA: cuTensorMapEncodeTiled not available from this driver");
  const int es = x_dtype == BT_BF16 ? 2 : 4;
  BT_REQUIRE(dst_off % 128 == 0 && box[0] * es == 128 && (long long)box[1] * box[2] * 128 + dst_off <= 32768, BT_ERR_BAD_SHAPE,
             "bt_tma_probe4d: box must be one 128-byte swizzle row wide and fit the 32 KB buffer");
  CUtensorMap map;
  cuuint64_t d[4] = {(cuuint64_t)dims[0], (cuuint64_t)dims[1], (cuuint64_t)dims[2], (cuuint64_t)dims[3]};
  cuuint64_t st[3] = {d[0] * es, d[0] * d[1] * es, d[0] * d[1] * d[2] * es};
  cuuint32_t bx[4] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2], 1};
  cuuint32_t es4[4] = {1, 1, 1, 1};
  const CUresult r = g_tma.tiled(&map, x_dtype == BT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                 const_cast<void*>(x), d, st, bx, es4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BT_REQUIRE(r == CUDA_SUCCESS, BT_ERR_CUDA, "bt_tma_probe4d: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return bt_tma_probe4d_launch(&map, coords[0], coords[1], coords[2], coords[3], dst_off, (uint32_t)(box[1] * box[2] * 128), out, stream);
}

int bt_layer_forward_plan(int mode, const BtLayerGeom* gm, int x_dtype, int p_dtype, int with_kl, int with_debug_hooks,
                          int with_residual, int sm_count, BtForwardPlan* plan) {
  BT_REQUIRE(plan != nullptr, BT_ERR_BAD_POINTER, "bt_layer_forward_plan: plan is NULL");
  // stand-ins that are only compared against NULL (never dereferenced in plan mode); 16-byte "aligned"
  float* const kl = with_kl ? reinterpret_cast<float*>(uintptr_t(16)) : nullptr;
  BtDebugIO dbg;
  memset(&dbg, 0, sizeof(dbg));
  if (with_debug_hooks) dbg.eps_w_in = reinterpret_cast<const float*>(uintptr_t(16));
  BtEpilogue epi;
  memset(&epi, 0, sizeof(epi));
  if (with_residual) epi.residual = reinterpret_cast<const void*>(uintptr_t(16));
  return layer_forward_impl(plan, sm_count, mode, gm, nullptr, x_dtype, nullptr, nullptr, nullptr, nullptr, p_dtype, nullptr,
                            kl, 0.f, 1.f, 0, 0, 0, with_debug_hooks ? &dbg : nullptr, with_residual ? &epi : nullptr,
                            kl ? reinterpret_cast<void*>(uintptr_t(16)) : nullptr, nullptr);
}

}  // extern "C"
