// K2d: "direct" implicit-GEMM convolution -- the activation operand of tcgen05.mma is read IN PLACE from a
// shared-memory input window; no im2col tile is ever built.  (Included by bt_fused.cu inside its anonymous
// namespace; shares FusedParams and the PTX wrappers.)
//
// Same reference op sequences as bt_fused_kernel (conv_variational.py:183-227 / 357-402 / 530-574,
// conv_flipout.py:175-244 / 370-439 / 568-637, linear_*.py forward), restricted to what makes the trick exact:
// stride 1, "same" output extent, groups == 1, C_in % 64 == 0, bf16 activations.
//
// Idea.  Number the pixels of one MC sample in a PADDED flattening
//     q = ((b * (D+pd) + d) * (H+ph) + h) * (W+pw) + w,      d < D+pd, h < H+ph, w < W+pw
// where every index with d >= D, h >= H or w >= W is a zero pixel (the trailing pad of one row / plane / image is
// also the leading pad of the next one).  In that numbering the input pixel that filter tap (kd,kh,kw) pairs with
// output pixel q is simply  q + delta_tap,  delta_tap = ((kd*dd-pd)*(H+ph) + (kh*dh-ph))*(W+pw) + (kw*dw-pw),
// the same shift for every q -- borders included, because out-of-image taps land on zero pixels.
// A 64-channel slab of one pixel is 128 bytes = one row of the 128B-swizzled K-major UMMA layout, so a window
// [q0 - halo, q0 + 128 + halo) of consecutive padded pixels, stored row by row with the hardware swizzle
// (16-byte chunk c of row j at j*128 + ((c ^ (j & 7)) << 4), window base 1024-aligned), already IS the A operand
// of every tap: the descriptor of tap t just starts (halo + delta_t) rows into the window.  Rows of the tile that
// are pad pixels produce garbage accumulator rows which the epilogue drops ((D+pd)(H+ph)(W+pw) / DHW extra MMA
// work: 1.27x at 8x8, 1.04x at 56x56) -- in exchange L2 and shared memory see every activation ONCE instead of
// once per tap, and the producer warps do nothing per k-block.
//
//   warps 0-7   sample the resident weight tiles W_s (all k-blocks of this CTA's (n-tile, sample)); afterwards
//   warps 0-6   stream the ring of input windows with cp.async (Flipout: + the x*s_in copy);
//   warp  7     MMA : the whole warp runs the issue loop, one elected lane issues tcgen05.mma; two accumulator
//               buffers in TMEM;
//   warps 8-15  epilogue (TMEM lane quarter = warp % 4, column half = (warp - 8) / 4): TMEM -> registers -> bias /
//               Flipout combine / BatchNorm affine / residual / ReLU -> HBM.
// Every role is a chain of dependent instructions per tile (a lone warp retires ~1 instruction per 5-7 clocks), so the
// per-tile work of the producers and of the epilogue is spread over 7 and 8 warps (profiles/r01h: with 4 + 4 warps
// the tile period was 3.3-4.5k clocks against the 1.9k the 36 MMAs of a 64-channel 3x3 tile need).
constexpr int DR_SAMP_WARPS = 8;   // warps 0-7 (the MMA warp samples too: it has nothing to issue before W_s exists)
constexpr int DR_PROD_WARPS = 7;   // warps 0-6
constexpr int DR_MMA_WARP = 7;
constexpr int DR_EPI_WARPS = 8;    // warps 8-15
constexpr int DR_THREADS = 16 * 32;   // 16 warps x 128 registers fill the register file (warps are allocated in fours)
constexpr int DR_AUX_BYTES = 4096;

// A descriptor may start at ANY 128-byte row of a 1024-aligned swizzled buffer: the hardware applies the swizzle XOR
// to absolute shared-memory address bits, so a row shift needs no correction and the descriptor's base-offset field
// stays 0 (verified on B200: tests/test_gpu_direct.py; setting it to (addr >> 7) & 7 gives wrong results).
//
// n / d for n < 2^31 by a host-computed reciprocal:  mul = ceil(2^(31+l) / d), l = ceil(log2 d), sh = 31 + l (exact;
// computed in bt_layer_forward) -- the window fill and the epilogue decode thousands of pixel indices per tile.
__device__ __forceinline__ uint32_t dr_div(uint32_t n, uint32_t mul, uint32_t sh) {
  return (uint32_t)(((unsigned long long)n * mul) >> sh);
}
// padded pixel index -> (is a real pixel, its dense index inside the sample)
__device__ __forceinline__ bool dr_decode(const FusedParams& p, long long q, uint32_t& m) {
  m = 0;
  if (q < 0 || q >= p.dr_Mp) return false;
  const uint32_t qq = (uint32_t)q;
  const uint32_t t1 = dr_div(qq, p.dr_mul[0], p.dr_sh[0]), w = qq - t1 * (uint32_t)p.dr_Pw;
  const uint32_t t2 = dr_div(t1, p.dr_mul[1], p.dr_sh[1]), h = t1 - t2 * (uint32_t)p.dr_Ph;
  const uint32_t b = dr_div(t2, p.dr_mul[2], p.dr_sh[2]), d = t2 - b * (uint32_t)p.dr_Pd;
  if (w >= (uint32_t)p.IW || h >= (uint32_t)p.IH || d >= (uint32_t)p.ID) return false;
  m = ((b * (uint32_t)p.ID + d) * (uint32_t)p.IH + h) * (uint32_t)p.IW + w;
  return true;
}

// optional phase probe (BT_DIRECT_TIMES=1 + a workspace): CTA (0,0,0) records SM-clock stamps
//   [0] kernel start, [1] resident tiles sampled, then per tile t < 60 and role r:  [8 + (r*60 + t)*2 + {0,1}]
//   r=0 producer thread 0 (window landed, next window issued)   r=1 MMA thread (window + accumulator available, MMAs issued)
//   r=2 epilogue warp 0 (accumulator complete, tile stored)
__device__ __forceinline__ void dr_stamp(const FusedParams& p, int idx) {
  if (p.dr_times != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dr_times[idx] = clock64();
}
__device__ __forceinline__ void dr_stamp_tile(const FusedParams& p, int role, long long it, int which) {
  if (it < 60) dr_stamp(p, 8 + (role * 60 + (int)it) * 2 + which);
}

template <int BLOCK_N, bool FLIP, bool P_BF16>
__global__ void __launch_bounds__(DR_THREADS, 1) bt_direct_kernel(const __grid_constant__ FusedParams p) {
  constexpr int NB = FLIP ? 2 : 1;
  constexpr int B_TILE_BYTES = BLOCK_N * 128;
  constexpr int P_ES = P_BF16 ? 2 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int slabs = p.Cin_g >> 6;
  const int R = p.dr_R;                                         // window rows (multiple of 8)
  const int res_bytes = p.num_kb * NB * B_TILE_BYTES;           // resident sampled tiles [kb][NB]
  const uint32_t plane_bytes = (uint32_t)(slabs * R * 128);     // one operand copy of one window slot [slab][R][128]
  const uint32_t slot_bytes = NB * plane_bytes;                 // Flipout: plane 0 = x, plane 1 = x * s_in
  const int NS = p.dr_slots;                                     // window ring depth (2..8)
  uint8_t* aux = smem + res_bytes + NS * slot_bytes;
  float* bias_s = reinterpret_cast<float*>(aux);                // [4][128]: bias (mean), bias (perturbation), scale, shift
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux + 2048);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t win0 = smem_base + res_bytes;
  const uint32_t bready_bar = smem_u32(bars);
  const uint32_t wfull_bar0 = smem_u32(bars + 1);    // [8] window slot filled (count = producer warps)
  const uint32_t wempty_bar0 = smem_u32(bars + 9);   // [8] window slot consumed by the tensor core
  const uint32_t acc_bar0 = smem_u32(bars + 17);     // [2] accumulator buffer complete
  const uint32_t tfree_bar0 = smem_u32(bars + 19);   // [2] accumulator buffer drained by the epilogue

  const int s = blockIdx.z;
  const int n0 = blockIdx.y * BLOCK_N;               // groups == 1
  const uint32_t sample = p.sample0 + (uint32_t)s + (p.sample_ptr != nullptr ? __ldg(p.sample_ptr) : 0u);
  const int img_base = p.x_shared ? 0 : s * p.B;
  const long long in_sp = (long long)p.ID * p.IH * p.IW;
  const long long n_rt = p.n_groups;                 // 128-row tiles of the PADDED pixel sequence
  const int n_taps = p.K_used / p.Cin_g;

  if (warp == DR_MMA_WARP) {
    if (lane == 0) {
      mbar_init(bready_bar, DR_SAMP_WARPS);
      for (int i = 0; i < NS; ++i) {
        mbar_init(wfull_bar0 + 8 * i, DR_PROD_WARPS);
        mbar_init(wempty_bar0 + 8 * i, 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(acc_bar0 + 8 * i, 1);
        mbar_init(tfree_bar0 + 8 * i, DR_EPI_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  } else {
    if (tid < BLOCK_N) {
      const int n = n0 + tid;
      float b0 = 0.f, b1 = 0.f, sc = 1.f, sh = 0.f;
      if (n < p.N) {
        if (p.mu_b != nullptr) {
          float mu, rho;
          if (P_BF16) {
            mu = __bfloat162float(static_cast<const __nv_bfloat16*>(p.mu_b)[n]);
            rho = __bfloat162float(static_cast<const __nv_bfloat16*>(p.rho_b)[n]);
          } else {
            mu = static_cast<const float*>(p.mu_b)[n];
            rho = static_cast<const float*>(p.rho_b)[n];
          }
          const float4 z = bt_eps_quad(p.key, BT_STREAM_B_EPS, (uint32_t)(n >> 2), 0u, sample);
          const int j = n & 3;
          const float eps = j == 0 ? z.x : (j == 1 ? z.y : (j == 2 ? z.z : z.w));
          const float d = bt_softplus(rho) * eps;
          if (FLIP) {
            b0 = mu;
            b1 = d;
          } else {
            b0 = mu + d;
          }
        }
        if (p.ep_scale != nullptr) {
          sc = __ldg(p.ep_scale + n);
          sh = __ldg(p.ep_shift + n);
        }
      }
      // Reparameterization: out = (acc + b) * sc + sh is applied as fma(acc, sc, b * sc + sh) -- one FMA and two
      // constants per element in the epilogue (exact when there is no affine: sc = 1, shift = b).
      bias_s[tid] = b0;
      bias_s[128 + tid] = b1;
      bias_s[256 + tid] = sc;
      bias_s[384 + tid] = FLIP ? sh : fmaf(b0, sc, sh);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) dr_stamp(p, 0);

  {
    // ============================================================== all warps
    const uint8_t* mu_w = static_cast<const uint8_t*>(p.mu_w);
    const uint8_t* rho_w = static_cast<const uint8_t*>(p.rho_w);
    const uint8_t* xb = static_cast<const uint8_t*>(p.x);
    // ---- 0. input-window machinery (the first NS - 1 windows are requested BEFORE the weights are sampled, so their
    //         HBM/L2 latency hides behind the sampling prologue).  8 consecutive lanes copy the 8 16-byte chunks of one
    //         128-byte pixel slab (coalesced); a lane owns chunk (lane & 7) of rows i*PROWS + warp*4 + (lane >> 3).
    const int ac = lane & 7, g4 = lane >> 3;      // this lane's 16-byte chunk and its row inside a 4-row warp pass
    constexpr int PROWS = DR_PROD_WARPS * 4;      // window rows per pass of the 4 producer warps
    const int wrow0 = warp * 4;                   // pass i covers window rows i*PROWS + wrow0 + g4
    const int n_pass = (R + PROWS - 1) / PROWS;
    const long long sample_pix0 = (long long)img_base * in_sp;
    const int D = NS >= 3 ? (NS - 2 > 6 ? 6 : NS - 2) : 1;   // windows in flight
    // Index decode shared through shuffles: in a round of 8 passes the warp touches 32 distinct window rows; lane e
    // decodes the row of (pass i0 + e/4, group e%4) ONCE and the 8 lanes that copy that row fetch the result with a
    // shuffle (the first version decoded every row in each of its 8 lanes: 475 instructions per warp per window and an
    // issue-bound kernel, profiles/r01h).  0xFFFFFFFF = zero pixel.
    auto decode_round = [&](long long first, int i0) -> uint32_t {
      const int j = (i0 + (lane >> 2)) * PROWS + wrow0 + (lane & 3);
      uint32_t m;
      const bool ok = dr_decode(p, first + j, m) && j < R;
      return ok ? m : 0xFFFFFFFFu;
    };
    auto load_window = [&](long long rt, int slot) {
      const long long first = rt * BLOCK_M - p.dr_halo;
      const uint32_t wbase = win0 + (uint32_t)slot * slot_bytes;
      for (int i0 = 0; i0 < n_pass; i0 += 8) {
        const uint32_t mdec = decode_round(first, i0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (i0 + u < n_pass) {                                  // (warp-uniform)
            const uint32_t m = __shfl_sync(0xffffffffu, mdec, u * 4 + g4);
            const int j = (i0 + u) * PROWS + wrow0 + g4;
            if (j < R) {
              const bool ok = m != 0xFFFFFFFFu;
              const uint8_t* src = xb + ((sample_pix0 + (ok ? m : 0u)) * p.C_in + ac * 8) * 2;
              const uint32_t dst = wbase + (uint32_t)(j * 128 + ((ac ^ (j & 7)) << 4));
              for (int sl = 0; sl < slabs; ++sl)
                cp_async16(dst + (uint32_t)(sl * R * 128), src + sl * 128, ok ? 16u : 0u);
            }
          }
        }
      }
      cp_async_commit();
    };
    // Flipout: plane 1 = plane 0 with the input signs applied (each thread re-reads exactly the chunks it copied).
    // One Philox call per window row and 128-channel block (by the lane that decoded the row), its words shuffled
    // to the 8 lanes of the row.
    auto sign_window = [&](long long rt, int slot) {
      const long long first = rt * BLOCK_M - p.dr_halo;
      const uint32_t wbase = win0 + (uint32_t)slot * slot_bytes;
      for (int i0 = 0; i0 < n_pass; i0 += 8) {
        const uint32_t mdec = decode_round(first, i0);
        for (int blk_i = 0; 2 * blk_i < slabs; ++blk_i) {
          uint4 blk = make_uint4(0u, 0u, 0u, 0u);
          if (mdec != 0xFFFFFFFFu) blk = bt_sign_block(p.key, BT_STREAM_SIGN_IN, (uint32_t)blk_i, mdec, sample);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (i0 + u < n_pass) {                                // (warp-uniform)
              const int srcl = u * 4 + g4;
              const int j = (i0 + u) * PROWS + wrow0 + g4;
#pragma unroll
              for (int hs = 0; hs < 2; ++hs) {                    // the two 64-channel slabs of this 128-channel block
                const int sl = 2 * blk_i + hs;
                if (sl < slabs) {                                 // (warp-uniform)
                  const uint32_t w_lo = __shfl_sync(0xffffffffu, hs ? blk.z : blk.x, srcl);
                  const uint32_t w_hi = __shfl_sync(0xffffffffu, hs ? blk.w : blk.y, srcl);
                  if (j < R) {
                    const uint32_t bits = (((ac >> 2) ? w_hi : w_lo) >> ((ac * 8) & 31)) & 0xffu;   // 0 for zero pixels
                    const uint4 mk = sign_masks8(bits);
                    const uint32_t a = wbase + (uint32_t)(j * 128 + ((ac ^ (j & 7)) << 4)) + (uint32_t)(sl * R * 128);
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
                    v.x ^= mk.x; v.y ^= mk.y; v.z ^= mk.z; v.w ^= mk.w;
                    sts16(a + plane_bytes, v);
                  }
                }
              }
            }
          }
        }
      }
    };
    if (warp < DR_PROD_WARPS) {
      // ring of NS window slots with D windows in flight (D = NS - 2, so that the slot a prefetch needs was consumed
      // two tiles ago and the producers never wait for the tensor core; D = 1 when only two slots fit).  Exactly one
      // cp.async group is committed per prologue step and per tile (empty once the tiles run out), so
      // "all but the newest D - 1 groups" == "window `it` has landed".
      for (int i = 0; i < D; ++i) {
        const long long rt = blockIdx.x + (long long)i * gridDim.x;
        if (rt < n_rt) load_window(rt, i);
        else cp_async_commit();
      }
    }
    if (warp < DR_SAMP_WARPS)
    // ---- 1. sample every k-block of this CTA's (n-tile, sample) into the resident region (same counters, same
    //         arithmetic and the same bf16 rounding as bt_fused_kernel's fast sampler => identical W_s)
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
        row_off[i] = (long long)(nvalid[i] ? n : p.N - 1) * p.K_phys;
      }
      int kb = 0;
      for (int t = 0; t < n_taps; ++t) {
        const long long tap_k = (long long)decode_tap(p, t).lin * p.Cin_g;
        for (int sl = 0; sl < slabs; ++sl, ++kb) {
          const long long kphys0 = tap_k + sl * BLOCK_K + wo * 8;
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
            c[i][1] = (uint32_t)(n0 + wrb + 32 * i);
            c[i][2] = sample;
            c[i][3] = p.key.c3_base | BT_STREAM_W_EPS;
          }
          philox_multi<WO>(c, p.key.k0, p.key.k1);
          const uint32_t sb = smem_base + (uint32_t)(kb * NB * B_TILE_BYTES);
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
            const bool ok = nvalid[i];
            float w0[8], w1[8];
            if (!p.rho_is_sigma) {   // (warp-uniform) skipped when the caller cached sigma = softplus(rho)
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
            const int nl = wrb + 32 * i;
            const uint32_t soff = (uint32_t)(nl * 128 + ((wo ^ (nl & 7)) << 4));
            sts16(sb + soff, make_uint4(bt_pack_bf16x2(w0[0], w0[1]), bt_pack_bf16x2(w0[2], w0[3]),
                                        bt_pack_bf16x2(w0[4], w0[5]), bt_pack_bf16x2(w0[6], w0[7])));
            if (FLIP)
              sts16(sb + B_TILE_BYTES + soff,
                    make_uint4(bt_pack_bf16x2(w1[0], w1[1]), bt_pack_bf16x2(w1[2], w1[3]),
                               bt_pack_bf16x2(w1[4], w1[5]), bt_pack_bf16x2(w1[6], w1[7])));
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bready_bar);
      if (tid == 0) dr_stamp(p, 1);
    }

    if (warp == DR_MMA_WARP) {
    // ============================================================== MMA issuer: the whole warp runs this loop
    // (warp-uniform operands -> uniform registers), one elected lane issues (umma_bf16_elect)
    {
      const uint32_t idesc = make_idesc(BLOCK_N);
      const uint64_t desc_hi = make_smem_desc(0u);                 // everything but the start-address field
      mbar_wait_idle(bready_bar, 0, 256);
      tc_fence_after();
      long long it = 0;
      int slot = 0;
      uint32_t wpar = 0;
      for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
        const int buf = (int)(it & 1);
        mbar_wait_idle(wfull_bar0 + 8 * slot, wpar, 32);
        if (it >= 2) mbar_wait_idle(tfree_bar0 + 8 * buf, (uint32_t)(((it >> 1) - 1) & 1), 32);
        tc_fence_after();
        if (lane == 0) dr_stamp_tile(p, 1, it, 0);
        const uint32_t wslot16 = ((win0 + (uint32_t)slot * slot_bytes) & 0x3FFFFu) >> 4;
        const uint32_t acc = tmem_base + (uint32_t)(buf * NB * BLOCK_N);
        uint32_t b16 = (smem_base & 0x3FFFFu) >> 4;               // start-address field of the resident tile of kb
        // p.dr_aoff[kb] (host-computed, constant bank -> uniform loads): where k-block kb = (tap, slab) starts inside a
        // window slot, in 16-byte units.  (Decoding the tap in this loop cost ~180 clocks per tap, profiles/r01h.)
#pragma unroll 2
        for (int kb = 0; kb < p.num_kb; ++kb, b16 += (uint32_t)(NB * B_TILE_BYTES) >> 4) {
          const uint32_t a16 = wslot16 + (uint32_t)p.dr_aoff[kb];
          // (low descriptor word = start-address field | LBO field (1 << 16); +32 bytes per K=16 step inside the asm)
          umma_bf16_elect_x4(acc, a16 | (1u << 16), b16 | (1u << 16), (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
          if (FLIP)
            umma_bf16_elect_x4(acc + BLOCK_N, (a16 + (plane_bytes >> 4)) | (1u << 16),
                               (b16 + (B_TILE_BYTES >> 4)) | (1u << 16), (uint32_t)(desc_hi >> 32), idesc, kb != 0 ? 1u : 0u);
        }
        umma_commit_elect(wempty_bar0 + 8 * slot);
        umma_commit_elect(acc_bar0 + 8 * buf);
        if (lane == 0) dr_stamp_tile(p, 1, it, 1);
        if (++slot == NS) {
          slot = 0;
          wpar ^= 1u;
        }
      }
    }
    __syncwarp();
    } else if (warp < DR_PROD_WARPS) {
      // ---- 2. stream the remaining windows
      long long it = 0;
      int slot = 0;                          // slot of tile `it`
      for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
        switch (D) {                         // this thread's part of window `it` has landed
          case 1: cp_async_wait<0>(); break;
          case 2: cp_async_wait<1>(); break;
          case 3: cp_async_wait<2>(); break;
          case 4: cp_async_wait<3>(); break;
          case 5: cp_async_wait<4>(); break;
          default: cp_async_wait<5>(); break;
        }
        if (tid == 0) dr_stamp_tile(p, 0, it, 0);
        if (FLIP) sign_window(rt, slot);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(wfull_bar0 + 8 * slot);
        const long long nxt = rt + (long long)D * gridDim.x;
        if (nxt < n_rt) {
          const long long prev = it + D - NS;            // tile that used the target slot last (if any)
          const int pslot = (int)((it + D) % NS);
          if (prev >= 0) mbar_wait_idle(wempty_bar0 + 8 * pslot, (uint32_t)((prev / NS) & 1), 64);
          load_window(nxt, pslot);
        } else {
          cp_async_commit();
        }
        if (tid == 0) dr_stamp_tile(p, 0, it, 1);
        if (++slot == NS) slot = 0;
      }
      cp_async_wait<0>();
    } else {
      // ============================================================== epilogue: 8 warps, (TMEM lane quarter, column half)
      // Full n-tiles go through a per-warp staging buffer [32 rows][EN bf16] (16-byte chunks XOR-swizzled): a lane owns
      // accumulator row `lane`, but global memory is touched row-contiguously -- CPR consecutive lanes cover one row's
      // EN * 2 bytes.  The residual of tile it+1 is fetched (coalesced, into registers) while tile it is processed, so
      // its HBM latency never sits on the per-tile critical path, and the accumulator buffer is handed back to the MMA
      // warp before the copy-out.
      constexpr int EN = BLOCK_N / 2;              // columns of this warp
      constexpr int ROWB = EN * 2;                 // staged bytes per row
      constexpr int CPR = EN / 8;                  // 16-byte chunks per row = lanes per row in the coalesced passes
      constexpr int RPI = 32 / CPR;                // rows per coalesced instruction
      constexpr int CSTEP = EN < 32 ? EN : 32;     // columns per TMEM round trip
      const int q4 = warp & 3, chalf = (warp - 8) >> 2;  // warps 8-11: half 0, warps 12-15: half 1
      const int ncol0 = chalf * EN;                // first column (inside the n-tile) of this warp
      uint8_t* outb = static_cast<uint8_t*>(p.out);
      const uint8_t* resb = static_cast<const uint8_t*>(p.ep_residual);
      const uint32_t stg = smem_u32(aux + DR_AUX_BYTES) + (uint32_t)((chalf * 4 + q4) * 32 * ROWB);
      auto swz = [](int c, int r) -> int {
        return CPR == 8 ? (c ^ (r & 7)) : (CPR == 4 ? (c ^ ((r >> 1) & 3)) : (c ^ ((r >> 2) & 1)));
      };
      const bool tile_vec = p.out_vec && n0 + BLOCK_N <= p.N;   // whole 16-byte chunks, no ragged columns
      const bool staged = tile_vec && p.dr_stage != 0;           // (the host drops the staging buffer when smem is short)
      const bool has_affine = p.ep_scale != nullptr;
      const bool pre_res = staged && p.ep_residual != nullptr;
      const int crow = lane / CPR, cch = lane % CPR;             // this lane's (row, chunk) in the coalesced passes
      auto row_of = [&](long long rt, uint32_t& m) -> long long {   // output row of this lane's accumulator row, -1 = pad
        const bool v = dr_decode(p, rt * BLOCK_M + q4 * 32 + lane, m);
        return v ? (long long)s * p.M + m : -1ll;
      };
      uint4 rv[CPR];                                              // residual chunks of the NEXT tile (pre_res)
      auto fetch_residual = [&](long long orow_t) {
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
          const long long ro = __shfl_sync(0xffffffffu, orow_t, i * RPI + crow);
          rv[i] = make_uint4(0u, 0u, 0u, 0u);
          if (ro >= 0) rv[i] = ldg16(resb + (ro * p.C_out + n0 + ncol0) * 2 + cch * 16);
        }
      };
      uint32_t m = 0, m_n = 0;
      long long orow = ((long long)blockIdx.x < n_rt) ? row_of(blockIdx.x, m) : -1ll;
      if (pre_res) fetch_residual(orow);
      long long it = 0;
      for (long long rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
        const int buf = (int)(it & 1);
        const bool mvalid = orow >= 0;
        if (pre_res) {                                            // this tile's residual -> staging
#pragma unroll
          for (int i = 0; i < CPR; ++i) {
            const int r = i * RPI + crow;
            sts16(stg + (uint32_t)(r * ROWB + (swz(cch, r) << 4)), rv[i]);
          }
          __syncwarp();
        }
        const long long rt_n = rt + gridDim.x;
        const long long orow_n = rt_n < n_rt ? row_of(rt_n, m_n) : -1ll;
        if (pre_res && rt_n < n_rt) fetch_residual(orow_n);       // in flight during this tile's wait + math + copy-out
        uint4 sblk = make_uint4(0u, 0u, 0u, 0u);
        if (FLIP) sblk = bt_sign_block(p.key, BT_STREAM_SIGN_OUT, (uint32_t)(n0 >> 7), m, sample);
        mbar_wait_idle(acc_bar0 + 8 * buf, (uint32_t)((it >> 1) & 1), 256);
        tc_fence_after();
        if (warp == 8 && lane == 0) dr_stamp_tile(p, 2, it, 0);
#pragma unroll 1
        for (int colb = 0; colb < EN; colb += CSTEP) {           // up to 32 columns per TMEM round trip
          const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * NB * BLOCK_N + ncol0 + colb);
          uint32_t va[CSTEP / 16][16], vb[CSTEP / 16][16];
#pragma unroll
          for (int h = 0; h < CSTEP / 16; ++h) {
            tmem_ld16(taddr + 16 * h, va[h]);
            if (FLIP) tmem_ld16(taddr + BLOCK_N + 16 * h, vb[h]);
          }
          uint4 rres[CSTEP / 16][2];
          if (tile_vec && !staged && p.ep_residual != nullptr && mvalid) {   // overlaps the TMEM round trip
            const uint8_t* rsd = resb + (orow * p.C_out + n0 + ncol0 + colb) * 2;
#pragma unroll
            for (int h = 0; h < CSTEP / 16; ++h) {
              rres[h][0] = ldg16(rsd + 32 * h);
              rres[h][1] = ldg16(rsd + 32 * h + 16);
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < CSTEP / 16; ++h) {
            const int col0 = ncol0 + colb + 16 * h;               // column inside the n-tile
            float o[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {                     // per-column constants: one LDS.128 per 4 columns
              const int col = col0 + 4 * jj;
              float v[4] = {__uint_as_float(va[h][4 * jj]), __uint_as_float(va[h][4 * jj + 1]),
                            __uint_as_float(va[h][4 * jj + 2]), __uint_as_float(va[h][4 * jj + 3])};
              const float4 sh = *reinterpret_cast<const float4*>(bias_s + 384 + col);
              if (FLIP) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias_s + col);
                const float4 b1 = *reinterpret_cast<const float4*>(bias_s + 128 + col);
                const float pb[4] = {b1.x, b1.y, b1.z, b1.w};
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float pert = __uint_as_float(vb[h][4 * jj + e]) + pb[e];
                  const int bit = (n0 & 127) + col + e;
                  const bool neg = (bt_sign_word(sblk, bit >> 5) >> (bit & 31)) & 1u;
                  v[e] += neg ? -pert : pert;
                }
                if (has_affine) {
                  const float4 sc = *reinterpret_cast<const float4*>(bias_s + 256 + col);
                  v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y);
                  v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w);
                }
              } else if (has_affine) {
                const float4 sc = *reinterpret_cast<const float4*>(bias_s + 256 + col);
                v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y);
                v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w);
              } else {
                v[0] += sh.x; v[1] += sh.y; v[2] += sh.z; v[3] += sh.w;     // shift = bias
              }
              o[4 * jj] = v[0]; o[4 * jj + 1] = v[1]; o[4 * jj + 2] = v[2]; o[4 * jj + 3] = v[3];
            }
            if (staged) {
              const int c0 = (colb + 16 * h) >> 3;                // chunk inside this warp's staged row
              const uint32_t sa0 = stg + (uint32_t)(lane * ROWB + (swz(c0, lane) << 4));
              const uint32_t sa1 = stg + (uint32_t)(lane * ROWB + (swz(c0 + 1, lane) << 4));
              if (p.ep_residual != nullptr) {
                uint4 a, b;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sa0));
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(sa1));
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  o[2 * j] += bt_bf16_lo(w[j]);
                  o[2 * j + 1] += bt_bf16_hi(w[j]);
                }
              }
              if (p.ep_relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
              }
              sts16(sa0, make_uint4(bt_pack_bf16x2(o[0], o[1]), bt_pack_bf16x2(o[2], o[3]),
                                    bt_pack_bf16x2(o[4], o[5]), bt_pack_bf16x2(o[6], o[7])));
              sts16(sa1, make_uint4(bt_pack_bf16x2(o[8], o[9]), bt_pack_bf16x2(o[10], o[11]),
                                    bt_pack_bf16x2(o[12], o[13]), bt_pack_bf16x2(o[14], o[15])));
            } else if (tile_vec) {   // no staging buffer: lane-per-row 16-byte stores (32 bytes of the row per pass)
              if (mvalid) {
                uint8_t* dst = outb + (orow * p.C_out + n0 + col0) * 2;
                if (p.ep_residual != nullptr) {
                  const uint4 a = rres[h][0], b = rres[h][1];
                  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    o[2 * j] += bt_bf16_lo(w[j]);
                    o[2 * j + 1] += bt_bf16_hi(w[j]);
                  }
                }
                if (p.ep_relu) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
                }
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(bt_pack_bf16x2(o[0], o[1]), bt_pack_bf16x2(o[2], o[3]),
                                                              bt_pack_bf16x2(o[4], o[5]), bt_pack_bf16x2(o[6], o[7]));
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(bt_pack_bf16x2(o[8], o[9]), bt_pack_bf16x2(o[10], o[11]),
                                                              bt_pack_bf16x2(o[12], o[13]), bt_pack_bf16x2(o[14], o[15]));
              }
            } else if (mvalid) {   // ragged n-tile / unaligned output: element-wise, lane per row
              const int nfirst = n0 + col0;
              const long long eoff = orow * p.C_out + nfirst;
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(outb + eoff * 2);
              const __nv_bfloat16* rsd = reinterpret_cast<const __nv_bfloat16*>(resb + eoff * 2);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (nfirst + j < p.N) {
                  float v = o[j];
                  if (p.ep_residual != nullptr) v += __bfloat162float(rsd[j]);
                  if (p.ep_relu) v = fmaxf(v, 0.f);
                  dst[j] = __float2bfloat16_rn(v);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tfree_bar0 + 8 * buf);     // accumulator drained: the MMA warp may reuse it
        if (staged) {
#pragma unroll
          for (int i = 0; i < CPR; ++i) {
            const int r = i * RPI + crow;
            const long long ro = __shfl_sync(0xffffffffu, orow, r);
            if (ro >= 0) {
              uint4 v;
              const uint32_t sa = stg + (uint32_t)(r * ROWB + (swz(cch, r) << 4));
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
              *reinterpret_cast<uint4*>(outb + (ro * p.C_out + n0 + ncol0) * 2 + cch * 16) = v;
            }
          }
          __syncwarp();                                        // staging is rewritten by the next tile
        }
        if (warp == 8 && lane == 0) dr_stamp_tile(p, 2, it, 1);
        orow = orow_n;
        m = m_n;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == DR_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

template <int BN, bool FLIP, bool PB>
int launch_direct(const FusedParams& p, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  static bool attr_done[64] = {};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!attr_done[dev]) {
      BT_CHECK_CUDA(cudaFuncSetAttribute(bt_direct_kernel<BN, FLIP, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BUDGET));
      attr_done[dev] = true;
    }
  }
  bt_direct_kernel<BN, FLIP, PB><<<grid, DR_THREADS, smem_bytes, st>>>(p);
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

template <int BN>
int dispatch_direct(const FusedParams& p, bool flip, dim3 grid, int smem_bytes, int dev, cudaStream_t st) {
  if (flip) return p.p_is_bf16 ? launch_direct<BN, true, true>(p, grid, smem_bytes, dev, st)
                               : launch_direct<BN, true, false>(p, grid, smem_bytes, dev, st);
  return p.p_is_bf16 ? launch_direct<BN, false, true>(p, grid, smem_bytes, dev, st)
                     : launch_direct<BN, false, false>(p, grid, smem_bytes, dev, st);
}
