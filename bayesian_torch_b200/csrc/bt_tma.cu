// Translation unit of the TMA-fed kernel families (bt_tma.cuh): bt_tma_kernel (W_s resident), bt_tms_kernel
// (streaming), bt_dtma_kernel (in-place windows) and the two TMA probes.  bt_fused.cu plans the launch
// (layer_forward_impl) and hands the filled parameter block to the entry points below.
#include "bt_kernels.cuh"

namespace {
std::mutex g_mu;
#define BT_TMA_DEVICE 1
#include "bt_tma.cuh"
}  // namespace

int bt_tma_family_launch(int kernel, const void* params, int bn, int tf32, int flip, unsigned gx, unsigned gy, unsigned gz,
                         int smem_bytes, int dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(gx, gy, gz);
  if (kernel == 2) {
    const DtParams& dp = *static_cast<const DtParams*>(params);
    if (bn == 128) return dispatch_dtma<128>(dp, tf32 != 0, flip != 0, grid, smem_bytes, dev, st);
    if (bn == 64) return dispatch_dtma<64>(dp, tf32 != 0, flip != 0, grid, smem_bytes, dev, st);
    return dispatch_dtma<32>(dp, tf32 != 0, flip != 0, grid, smem_bytes, dev, st);
  }
  const TmaParams& tp = *static_cast<const TmaParams*>(params);
  if (bn == 128) return dispatch_tma<128>(tp, tf32 != 0, kernel == 1, flip != 0, grid, smem_bytes, dev, st);
  if (bn == 64) return dispatch_tma<64>(tp, tf32 != 0, kernel == 1, flip != 0, grid, smem_bytes, dev, st);
  return dispatch_tma<32>(tp, tf32 != 0, kernel == 1, flip != 0, grid, smem_bytes, dev, st);
}

int bt_tma_probe_launch(const void* params, long long m0, int sample, int group, int tap, int slab, void* out, void* stream) {
  static bool attr_done = false;
  if (!attr_done) {
    BT_CHECK_CUDA(cudaFuncSetAttribute(bt_tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_TILE_BYTES + 2048));
    attr_done = true;
  }
  bt_tma_probe_kernel<<<1, 32, A_TILE_BYTES + 2048, static_cast<cudaStream_t>(stream)>>>(
      *static_cast<const TmaParams*>(params), m0, sample, group, tap, slab, static_cast<uint8_t*>(out));
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}

int bt_tma_probe4d_launch(const void* map, int c, int w, int h, int n, uint32_t dst_off, uint32_t bytes, void* out, void* stream) {
  static bool attr_done = false;
  if (!attr_done) {
    BT_CHECK_CUDA(cudaFuncSetAttribute(bt_tma_probe4d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 2048));
    attr_done = true;
  }
  bt_tma_probe4d_kernel<<<1, 32, 32768 + 2048, static_cast<cudaStream_t>(stream)>>>(*static_cast<const CUtensorMap*>(map), c, w, h, n,
                                                                                    dst_off, bytes, static_cast<uint8_t*>(out));
  BT_CHECK_CUDA(cudaGetLastError());
  return BT_OK;
}
