// umma_probe.cu — standalone micro-benchmark (NOT part of libvlmerge): issue rate of tcgen05.mma with both operands in
// shared memory, to find out what bounds the SYRK mainloop.
//
//   umma_probe [iters]
//
// Every CTA pair (cluster of 2, one CTA per SM, all 148 SMs) spins on MMAs over four resident 48 KB stage buffers —
// no loads, no epilogue — for the operand layouts in question:
//   group 1 : each CTA issues its own M = 128, N = 256 instructions (the round-1 multicast pair kernel)
//   group 2 : the leader issues ONE M = 256, N = 256 cta_group::2 instruction for the pair (each CTA holds its 128 A rows
//             and half of B)
//   MN-major (the Gram layout: X^T X with X row-major) vs K-major (the usual GEMM layout), tf32 vs bf16,
//   and optionally a second warp streaming 48 KB per iteration from global memory into the same shared memory with
//   cp.async.bulk (the TMA write traffic of the real mainloop).
// Prints TFLOP/s per variant (dense count 2*M*N*K per instruction) from CUDA-event time around the launch.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

namespace {

constexpr int kStages = 4;
constexpr int kStageBytes = 49152;
constexpr int kSmem = kStages * kStageBytes + 1024 + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

template <int GROUP>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  if constexpr (GROUP == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int GROUP>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (GROUP == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int GROUP, int TF32>
__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if constexpr (GROUP == 1 && TF32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc)
                 : "memory");
  else if constexpr (GROUP == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc)
                 : "memory");
  else if constexpr (TF32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc)
                 : "memory");
}
template <int GROUP>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (GROUP == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// GROUP 1|2, TF32 1 (kind::tf32, K = 8) | 0 (kind::f16 with bf16 operands, K = 16), MN 1 (both operands MN-major) | 0
// (K-major), WRITER: a second warp streams 48 KB per iteration into shared memory
template <int GROUP, int TF32, int MN, int WRITER>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_kernel(int iters, const uint8_t* __restrict__ src) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* done = bars;          // all MMAs complete
  uint64_t* wbar = bars + 1;      // [2] writer's bulk copies
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ctarank();
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    mbar_init(wbar, 1);
    mbar_init(wbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<GROUP>(slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;

  // descriptors.  MN-major: LBO = one column group of BK rows (BK*128 B), SBO = one swizzle atom of rows; layout
  // SWIZZLE_128B_BASE32B (1) for tf32, SWIZZLE_128B (2) for 16-bit — exactly the SYRK kernels'.  K-major: 128-byte
  // rows (32 tf32 / 64 bf16 K elements), 8-row atoms 1024 B apart, SWIZZLE_128B; one MMA advances K by 32 bytes.
  constexpr uint32_t kLayout = MN ? (TF32 ? 1u : 2u) : 2u;
  constexpr uint32_t kLbo = MN ? (uint32_t)((TF32 ? 32 : 64) * 128) : 16u;
  constexpr uint32_t kSbo = MN ? (TF32 ? 512u : 1024u) : 1024u;
  constexpr uint32_t kStep = MN ? (uint32_t)((TF32 ? 8 : 16) * 128) : 32u;
  constexpr uint32_t fmt = TF32 ? 2u : 1u;
  constexpr uint32_t kM = GROUP == 1 ? 128u : 256u;
  constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)MN << 15) | ((uint32_t)MN << 16) |
                             ((256u >> 3) << 17) | ((kM >> 4) << 24);
  if (warp == 1 && (GROUP == 1 || rank == 0)) {
    const uint32_t base = smem_u32(smem);
    for (int it = 0; it < iters; ++it) {
      const uint32_t sb = base + (uint32_t)(it & (kStages - 1)) * kStageBytes;
      // group 1: [B 32 KB][A 16 KB]; group 2: [B half 16 KB][A half 16 KB]
      const uint32_t sa = sb + (GROUP == 1 ? 32768u : 16384u);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma<GROUP, TF32>(tmem, smem_desc(sa + kk * kStep, kLbo, kSbo, kLayout), smem_desc(sb + kk * kStep, kLbo, kSbo, kLayout),
                           idesc);
      }
      __syncwarp();
    }
    if (elect_one()) commit<GROUP>(done);
    __syncwarp();
  }
  if (WRITER && warp == 2 && lane == 0) {
    // 48 KB per iteration in three 16 KB bulk copies, two iterations in flight
    const uint32_t base = smem_u32(smem);
    for (int it = 0; it < iters; ++it) {
      uint64_t* b = wbar + (it & 1);
      if (it >= 2) mbar_wait(b, ((it - 2) >> 1) & 1);
      mbar_expect_tx(b, 3 * 16384);
      const uint32_t dst = base + (uint32_t)((it + 2) & (kStages - 1)) * kStageBytes;
      const uint8_t* g = src + ((size_t)(blockIdx.x * 7 + it) % 512) * 49152;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst + j * 16384),
                     "l"(g + j * 16384), "r"(16384), "r"(smem_u32(b))
                     : "memory");
    }
    for (int it = (iters > 2 ? iters - 2 : 0); it < iters; ++it) mbar_wait(wbar + (it & 1), (it >> 1) & 1);
  }
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0) tmem_dealloc<GROUP>(tmem, 512);
}

template <int GROUP, int TF32, int MN, int WRITER>
void run(const char* name, int iters, int nsm, const uint8_t* src) {
  auto k = probe_kernel<GROUP, TF32, MN, WRITER>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  k<<<nsm, 128, kSmem>>>(iters / 10 + 1, src);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  k<<<nsm, 128, kSmem>>>(iters, src);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double kdim = TF32 ? 8.0 : 16.0;
  // per iteration and per SM: 4 instructions x 2 * 128 rows * 256 columns * K (a cta_group::2 instruction covers 2 SMs)
  const double flops = (double)nsm * iters * 4.0 * 2.0 * 128.0 * 256.0 * kdim;
  const double cyc = ms * 1e-3 * 1.965e9 / iters;
  printf("%-44s %8.3f ms  %8.1f TFLOP/s   ~%6.0f cycles per 4-MMA step at 1.965 GHz\n", name, ms, flops / ms * 1e-9, cyc);
  fflush(stdout);
}

}  // namespace

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount & ~1;
  printf("device: %s, %d SMs used, %d iterations of 4 MMAs per CTA (pair)\n", prop.name, nsm, iters);
  uint8_t* src = nullptr;
  CK(cudaMalloc(&src, (size_t)512 * 49152 + 65536));
  CK(cudaMemset(src, 0, (size_t)512 * 49152 + 65536));
  run<1, 1, 1, 0>("group1 tf32 MN-major", iters, nsm, src);
  run<1, 1, 0, 0>("group1 tf32 K-major", iters, nsm, src);
  run<2, 1, 1, 0>("group2 tf32 MN-major", iters, nsm, src);
  run<2, 1, 0, 0>("group2 tf32 K-major", iters, nsm, src);
  run<1, 0, 1, 0>("group1 bf16 MN-major", iters, nsm, src);
  run<1, 0, 0, 0>("group1 bf16 K-major", iters, nsm, src);
  run<2, 0, 1, 0>("group2 bf16 MN-major", iters, nsm, src);
  run<2, 0, 0, 0>("group2 bf16 K-major", iters, nsm, src);
  run<1, 1, 1, 1>("group1 tf32 MN-major + 48 KB/iter smem writes", iters, nsm, src);
  run<2, 1, 1, 1>("group2 tf32 MN-major + 48 KB/iter smem writes", iters, nsm, src);
  run<1, 0, 1, 1>("group1 bf16 MN-major + 48 KB/iter smem writes", iters, nsm, src);
  return 0;
}
