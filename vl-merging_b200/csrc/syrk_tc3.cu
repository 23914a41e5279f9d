// syrk_tc3.cu — kernel (a), third generation: ONE tcgen05.mma.cta_group::2 per K step for the CTA pair.
//
// Same super-tile decomposition, schedule and epilogue as syrk_tc2.cu.  What changes is the MMA itself:
// the pair issues a single M = 256, N = 256 instruction (cta_group::2) from the leader CTA.  CTA r supplies
// the A rows of its block 2a+r and HALF of the B operand (column block 2b+r) from its own shared memory,
// and the tensor cores of both SMs read both halves.  Consequences:
//   * a pipeline stage is 32 KB per CTA ([B half][A half]) instead of 48 KB, so six stages fit where the
//     multicast kernel had four: 50 % more load latency is covered with the same shared memory;
//   * nothing is multicast: each CTA writes only its own 32 KB per stage (16 KB on a diagonal super-tile,
//     where the A half of CTA r IS its B half);
//   * L2 -> SM traffic is the same 32 KB per CTA per stage.
// The price is the 2-CTA protocol: TMA loads of both CTAs complete on the LEADER's full barrier
// (.cta_group::2), tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs,
// and both CTAs' epilogue threads release the accumulator on the leader's barrier.
// On a diagonal super-tile the block below the diagonal is computed (one instruction cannot skip it) but
// not stored.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "syrk.h"
#include "umma.cuh"

namespace vlm {

namespace {

constexpr int kStages = 6;
constexpr int kStageBytes = 2 * kBlockBytes;  // [B half][A half]
constexpr int kStagingBytes = 16384;
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStagingBytes + 256 + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank3() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all3() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes complete on the mbarrier at cluster address `bar`
// (the leader's full barrier)
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* tm, uint32_t bar_cluster_addr, void* smem_dst,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
template <int FMT>
__device__ __forceinline__ void umma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  if constexpr (FMT == 2) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// instruction descriptor as make_idesc, but M = 256 (the pair's rows)
__host__ __device__ constexpr uint32_t make_idesc_m256(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 15) | (1u << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int ELEM_BYTES, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
syrk_tc3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                const PairSeg* __restrict__ segs, const int* __restrict__ seg_off, int d) {
  using G = Geo<ELEM_BYTES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* full = bars;                       // [kStages] used in the LEADER: both CTAs' loads land here
  uint64_t* empty = bars + kStages;            // [kStages] per CTA: the pair's MMAs have read this CTA's slot
  uint64_t* tfull = bars + 2 * kStages;        // [2] per CTA: accumulator complete
  uint64_t* tempty = bars + 2 * kStages + 2;   // [2] used in the LEADER: both CTAs' epilogues have drained it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank3();
  const int cluster_id = blockIdx.x >> 1;
  const int seg_begin = seg_off[cluster_id];
  const int seg_end = seg_off[cluster_id + 1];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 2);   // one arrive(+expect_tx) per CTA's producer
      mbar_init(&empty[i], 1);  // the leader's tcgen05.commit (multicast)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 256);  // 128 epilogue threads of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all3();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs): own B half and own A half, completing on the leader's barrier =====
    int stage = 0;
    uint32_t phase = 0;
    for (int s = seg_begin; s < seg_end; ++s) {
      const PairSeg seg = segs[s];
      const bool diag = seg.sa == seg.sb;
      const int a_group = (2 * seg.sa + (int)rank) * G::GB;
      const int b_group = (2 * seg.sb + (int)rank) * G::GB;
      const uint32_t bytes = (diag ? 1u : 2u) * kBlockBytes;
      for (int k = seg.k0; k < seg.k1; ++k) {
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t leader_full = mapa_u32(smem_u32(&full[stage]), 0);
        mbar_arrive_expect_tx_cluster(leader_full, bytes);
        uint8_t* sb = stage_base + stage * kStageBytes;
        const int row = k * G::BK;
        tma_load_3d_2sm(&tm_x, leader_full, sb, 0, row, b_group);
        if (!diag) tma_load_3d_2sm(&tm_x, leader_full, sb + kBlockBytes, 0, row, a_group);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===== MMA issuer: leader CTA only =====
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t idesc = make_idesc_m256(FMT, 256);
    for (int s = seg_begin; s < seg_end; ++s) {
      const PairSeg seg = segs[s];
      const bool diag = seg.sa == seg.sb;
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int k = seg.k0; k < seg.k1; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sb = smem_u32(stage_base + stage * kStageBytes);
        const uint32_t sa = diag ? sb : sb + kBlockBytes;
#pragma unroll
        for (int kk = 0; kk < G::NUM_MMA; ++kk) {
          const uint64_t adesc = make_smem_desc<G::LAYOUT_TYPE>(sa + kk * G::KSTEP_BYTES, G::BOX_BYTES, G::SBO_BYTES);
          const uint64_t bdesc = make_smem_desc<G::LAYOUT_TYPE>(sb + kk * G::KSTEP_BYTES, G::BOX_BYTES, G::SBO_BYTES);
          umma_2sm<FMT>(d_tmem, adesc, bdesc, idesc, (k > seg.k0 || kk > 0) ? 1u : 0u);
        }
        tc_commit_2sm_mcast(&empty[stage], (uint16_t)0x3);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit_2sm_mcast(&tfull[acc], (uint16_t)0x3);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): own 128 rows x 256 columns =====
    const int q = warp - 4;
    const int epi_tid = threadIdx.x - 128;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_counter = 0;
    for (int s = seg_begin; s < seg_end; ++s) {
      const PairSeg seg = segs[s];
      const bool diag = seg.sa == seg.sb;
      const int row0 = (2 * seg.sa + (int)rank) * 128;
      const int col0 = 2 * seg.sb * 128;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int first = (diag && rank == 1) ? 4 : 0;  // the block below the diagonal is not stored
      const int nslab = (row0 < d) ? min(8, (d - col0 + 31) / 32) : 0;
      for (int sl = first; sl < nslab; ++sl) {
        uint8_t* buf = staging + (slab_counter & 1) * kStagingBytes;
        if (epi_tid == 0) bulk_wait_group_read<1>();
        named_bar_sync(1, 128);
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols + sl * 32, v);
        tmem_ld_wait();
        const uint32_t rbase = smem_u32(buf) + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = rbase + ((uint32_t)(c ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                       "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (epi_tid == 0) {
          tma_reduce_add_2d(&tm_g, buf, col0 + sl * 32, row0);
          bulk_commit_group();
        }
        ++slab_counter;
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));  // release the accumulator on the leader's barrier
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (epi_tid == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  cluster_sync_all3();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
}

struct DeviceSchedule3 {
  int nclusters = 0;
  PairSeg* d_segs = nullptr;
  int* d_off = nullptr;
};
std::mutex g_mu3;
std::map<std::tuple<int, int64_t, int, int, int>, DeviceSchedule3> g_sched3;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode3 = nullptr;

template <int ELEM_BYTES, int FMT>
int launch_kernel3(int dev, const DeviceSchedule3& sched, const CUtensorMap& tm_x, const CUtensorMap& tm_g, int d,
                   cudaStream_t stream) {
  static std::atomic<bool> attr_done[64];
  auto kernel = syrk_tc3_kernel<ELEM_BYTES, FMT>;
  if (dev >= 64 || !attr_done[dev].load(std::memory_order_acquire)) {
    VLM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    if (dev < 64) attr_done[dev].store(true, std::memory_order_release);
  }
  kernel<<<2 * sched.nclusters, kThreads, kSmemBytes, stream>>>(tm_x, tm_g, sched.d_segs, sched.d_off, d);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int syrk_tc3_launch(const void* x, int dtype, int64_t rows, int d, int64_t ldx, float* g, int64_t ldg,
                    cudaStream_t stream) {
  const int elem = (dtype == VLM_F32) ? 4 : 2;
  const int bk = 128 / elem;
  const int gc = 128 / elem;
  VLM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((ldx * elem) & 15) == 0, VLM_ERR_ALIGNMENT,
              "vlm_syrk_accum: x must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
  VLM_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (ldg & 3) == 0, VLM_ERR_ALIGNMENT,
              "vlm_syrk_accum: g must be 16-byte aligned with ldg %% 4 == 0");
  VLM_REQUIRE(rows < (int64_t)1 << 31, VLM_ERR_INVALID_ARG, "vlm_syrk_accum: rows too large");
  int dev = 0, nsm = 0;
  VLM_CUDA(cudaGetDevice(&dev));
  if (int rc = device_sm_count(&nsm)) return rc;
  {
    std::lock_guard<std::mutex> lk(g_mu3);
    if (!g_encode3) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      VLM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      VLM_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, VLM_ERR_DRIVER,
                  "cuTensorMapEncodeTiled not available from the driver");
      g_encode3 = reinterpret_cast<EncodeTiledFn>(fn);
    }
  }
  const int64_t kc = (rows + bk - 1) / bk;
  DeviceSchedule3 sched;
  {
    std::lock_guard<std::mutex> lk(g_mu3);
    auto key = std::make_tuple(dev, kc, d, bk, nsm);
    auto it = g_sched3.find(key);
    if (it == g_sched3.end()) {
      std::vector<PairSeg> segs;
      std::vector<int> off;
      build_pair_schedule(kc, d, nsm / 2, &segs, &off);
      DeviceSchedule3 ds;
      ds.nclusters = (int)off.size() - 1;
      VLM_CUDA(cudaMalloc(&ds.d_segs, std::max<size_t>(1, segs.size()) * sizeof(PairSeg)));
      VLM_CUDA(cudaMalloc(&ds.d_off, off.size() * sizeof(int)));
      VLM_CUDA(cudaMemcpyAsync(ds.d_segs, segs.data(), segs.size() * sizeof(PairSeg), cudaMemcpyHostToDevice, stream));
      VLM_CUDA(cudaMemcpyAsync(ds.d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
      VLM_CUDA(cudaStreamSynchronize(stream));
      it = g_sched3.emplace(key, ds).first;
    }
    sched = it->second;
  }
  CUtensorMap tm_x, tm_g;
  {
    const CUtensorMapDataType dt = dtype == VLM_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : dtype == VLM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                       : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t gdim[3] = {(cuuint64_t)gc, (cuuint64_t)rows, (cuuint64_t)(d / gc)};
    cuuint64_t gstr[2] = {(cuuint64_t)ldx * elem, 128};
    cuuint32_t box[3] = {(cuuint32_t)gc, (cuuint32_t)bk, (cuuint32_t)elem};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle swz = elem == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = g_encode3(&tm_x, dt, 3, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLM_REQUIRE(r == CUDA_SUCCESS, VLM_ERR_DRIVER, "cuTensorMapEncodeTiled(X, 3-D) failed: CUresult %d", (int)r);
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)d};
    cuuint64_t gstr[1] = {(cuuint64_t)ldg * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode3(&tm_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLM_REQUIRE(r == CUDA_SUCCESS, VLM_ERR_DRIVER, "cuTensorMapEncodeTiled(G) failed: CUresult %d", (int)r);
  }
  if (dtype == VLM_F32) return launch_kernel3<4, 2>(dev, sched, tm_x, tm_g, d, stream);
  if (dtype == VLM_BF16) return launch_kernel3<2, 1>(dev, sched, tm_x, tm_g, d, stream);
  return launch_kernel3<2, 0>(dev, sched, tm_x, tm_g, d, stream);
}

}  // namespace vlm
