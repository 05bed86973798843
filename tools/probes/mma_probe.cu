// tcgen05.mma issue / throughput probe (SS mode, K-major SW128 operands in shared memory, one issuing thread).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mma_probe tools/probes/mma_probe.cu
//   run  : tools/probes/mma_probe            (prints clocks per MMA for N in {32, 64, 128, 256}, A start row offsets)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(int n, int reps, int a_row_off, int tf32, int same_acc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tslot;
  uint32_t elected = 0;
  if (tid < 32) {
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
  }
  if (elected) {
    // idesc: c fp32 (bit 4), a/b bf16 (1<<7, 1<<10) or tf32 (2<<7, 2<<10), N>>3 at bit 17, M>>4 at bit 24
    const uint32_t fmt = tf32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hi = ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t a0 = ((smem_u32(smem) + (uint32_t)a_row_off * 128u) & 0x3FFFFu) >> 4;
    const uint32_t b0 = ((smem_u32(smem) + 40960u) & 0x3FFFFu) >> 4;
    const uint64_t da = hi | (1ull << 16) | a0, db = hi | (1ull << 16) | b0;
    if (same_acc == 2) {   // best case: 16 MMAs per iteration, descriptors = loop-invariant base + compile-time offsets
      long long t0 = clock64();
      for (int r = 0; r < reps; r += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const uint64_t dak = da + (uint64_t)((u & 3) * 2), dbk = db + (uint64_t)((u & 3) * 2);
          if (tf32)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(dak), "l"(dbk), "r"(idesc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(dak), "l"(dbk), "r"(idesc) : "memory");
        }
      }
      long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    } else {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (same_acc ? 0u : (uint32_t)((r & 1) * n));
      const uint64_t dak = da + (uint64_t)((r & 3) * 2), dbk = db + (uint64_t)((r & 3) * 2);
      if (tf32)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(dak), "l"(dbk), "r"(idesc), "r"(r) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(dak), "l"(dbk), "r"(idesc), "r"(r) : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 2048;
  printf("clocks per MMA (M=128, K=16 bf16 / K=8 tf32), %d MMAs back to back from one thread: issue-loop / until-commit\n", reps);
  for (int tf32 = 0; tf32 < 2; ++tf32)
    for (int n : {32, 64, 128, 256})
      for (int off : {0, 9})
        for (int same : {1, 2}) {
          probe<<<1, 128, 100 * 1024>>>(n, reps, off, tf32, same, out);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2] = {0, 0};
          cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
          printf("%s N=%3d a_row_off=%d %s: %.1f / %.1f  (%s)\n", tf32 ? "tf32" : "bf16", n, off, same == 2 ? "unrolled16" : "loop      ",
                 (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(e));
        }
  // all 148 SMs at once (shared-memory bandwidth is per SM; checks there is no chip-level limit)
  probe<<<148, 128, 100 * 1024>>>(64, reps, 9, 0, 1, out);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("148 CTAs bf16 N=64 off=9: %.1f / %.1f\n", (double)h[0] / reps, (double)h[1] / reps);
  return 0;
}
