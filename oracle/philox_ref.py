"""numpy restatement of the device RNG used by the fused kernels (TEST INFRASTRUCTURE).

The reference draws eps with ``Tensor.normal_()`` and signs with
``Tensor.uniform_(-1, 1).sign()``
(/root/reference/bayesian_torch/layers/variational_layers/linear_variational.py:161,173;
 /root/reference/bayesian_torch/layers/flipout_layers/linear_flipout.py:169-170).
Those are ATen's RNG streams; the B200 kernels instead generate eps / signs on
chip from a counter-based Philox4x32-10 (Salmon et al., "Parallel random
numbers: as easy as 1, 2, 3", SC'11 -- Random123).  This file is the CPU
statement of exactly that generator so tests can (a) check the device Philox
against the published Random123 known-answer vectors and (b) predict, element
by element, the eps / sign tensors a kernel launch will use.

Counter layout (must match bayesian_torch_b200/csrc/bt_philox.cuh):
    key     = (seed_lo, seed_hi)
    counter = (c0, c1, c2, c3)
      c3 = (layer_key << 4) | stream       stream: 0 weight eps, 1 bias eps,
                                                    2 input signs, 3 output signs
      c2 = global MC sample index
      weight eps : c1 = output row n,  c0 = k // 8   (k = physical K index), lane = k % 8
      bias eps   : c1 = 0,             c0 = n // 8,  lane = n % 8
      input sign : c1 = pixel index inside the sample, c0 = channel // 128, bit = channel % 128
      output sign: c1 = output row m inside the sample, c0 = n // 128,      bit = n % 128
Normals: every 32-bit output word w gives TWO normals (8 per Philox call) by Box-Muller on 16-bit uniforms:
      u = 1 - (w & 0xffff) * 2^-16  in (0,1],  v = (w >> 16) * 2^-16  in [0,1)
      r = sqrt(-2 ln u);  z0 = r cos(2 pi v);  z1 = r sin(2 pi v)      (lanes 2j, 2j+1 from word j)
Signs: bit b of the 128-bit block (word b // 32, bit b % 32): 1 -> -1.0, 0 -> +1.0.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

STREAM_W_EPS, STREAM_B_EPS, STREAM_SIGN_IN, STREAM_SIGN_OUT = 0, 1, 2, 3


def philox4x32_10(ctr, key):
    """ctr: uint32[..., 4], key: uint32[..., 2] (broadcastable) -> uint32[..., 4]."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    shape = np.broadcast(ctr[..., 0], key[..., 0]).shape
    k0 = np.broadcast_to(key[..., 0], shape).astype(np.uint32).copy()
    k1 = np.broadcast_to(key[..., 1], shape).astype(np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
        n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        with np.errstate(over="ignore"):
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _c3(layer_key, stream):
    return np.uint32(((int(layer_key) << 4) | int(stream)) & 0xFFFFFFFF)


def box_muller8(x):
    """x: uint32[..., 4] -> float32[..., 8] standard normals (float64 math, rounded)."""
    x = np.asarray(x, dtype=np.uint32)
    u = 1.0 - (x & np.uint32(0xFFFF)).astype(np.float64) * 2.0 ** -16
    v = (x >> np.uint32(16)).astype(np.float64) * 2.0 ** -16
    r = np.sqrt(-2.0 * np.log(u))
    z = np.empty(x.shape[:-1] + (8,), dtype=np.float64)
    z[..., 0::2] = r * np.cos(2.0 * np.pi * v)
    z[..., 1::2] = r * np.sin(2.0 * np.pi * v)
    return z.astype(np.float32)


def weight_eps(n_rows, k_cols, seed, layer_key, sample_idx):
    """eps for a [n_rows, k_cols] weight in PHYSICAL (row, k) order."""
    ko = (k_cols + 7) // 8
    ctr = np.zeros((n_rows, ko, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(ko, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.arange(n_rows, dtype=np.uint32)[:, None]
    ctr[..., 2] = np.uint32(sample_idx)
    ctr[..., 3] = _c3(layer_key, STREAM_W_EPS)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    z = box_muller8(philox4x32_10(ctr, key))
    return z.reshape(n_rows, ko * 8)[:, :k_cols]


def bias_eps(n, seed, layer_key, sample_idx):
    no = (n + 7) // 8
    ctr = np.zeros((no, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(no, dtype=np.uint32)
    ctr[:, 2] = np.uint32(sample_idx)
    ctr[:, 3] = _c3(layer_key, STREAM_B_EPS)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return box_muller8(philox4x32_10(ctr, key)).reshape(-1)[:n]


def sign_bits(n_rows, n_cols, seed, layer_key, sample_idx, stream):
    """+-1 float32 [n_rows, n_cols]; row = pixel (input signs) or output row (output signs)."""
    blocks = (n_cols + 127) // 128
    ctr = np.zeros((n_rows, blocks, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(blocks, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.arange(n_rows, dtype=np.uint32)[:, None]
    ctr[..., 2] = np.uint32(sample_idx)
    ctr[..., 3] = _c3(layer_key, stream)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    words = philox4x32_10(ctr, key)                      # [rows, blocks, 4]
    bits = ((words[..., None] >> np.arange(32, dtype=np.uint32)) & 1)  # [rows, blocks, 4, 32]
    bits = bits.reshape(n_rows, blocks * 128)[:, :n_cols]
    return np.where(bits == 1, -1.0, 1.0).astype(np.float32)


# Random123 known-answer vectors for philox4x32-10 (kat_vectors in the Random123 distribution).
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]
