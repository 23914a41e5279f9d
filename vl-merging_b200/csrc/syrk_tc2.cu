// syrk_tc2.cu — kernel (a), second generation: CTA pairs with TMA multicast.
//
// Same arithmetic and epilogue as syrk_tc.cu (tcgen05.mma kind::tf32 / kind::f16 on MN-major TMA tiles,
// fp32 accumulators in TMEM, TMA reduce-add into G).  What changes is how bytes reach shared memory,
// because the first-generation kernel was bound by L2->SM traffic (48 KB per 512 tensor-pipe cycles per
// SM, issued as 12 small TMA boxes per stage):
//
//  * work unit = a 256x256 SUPER-TILE (a, b), b >= a, computed by a cluster of two CTAs: CTA r owns the
//    row block 2a+r and both column blocks 2b, 2b+1.  The two CTAs need the same B columns, so each
//    loads ONE B block and multicasts it into both CTAs' shared memory (.multicast::cluster), plus its
//    own A block: 32 KB from L2 per CTA per stage instead of 48 KB.  On a diagonal super-tile the A
//    block of CTA r IS B block r: 16 KB per CTA per stage, and CTA 1 only computes its diagonal block
//    (N = 128) — the block below the diagonal is never formed.
//  * X is described to TMA as a 3-D tensor {columns-in-group, rows, column groups} (strides 4 B,
//    pitch, 128 B), so ONE cp.async.bulk.tensor box {128 B, BK rows, groups-per-block} fetches a whole
//    128-column block in exactly the [group][row][128 B] order the UMMA descriptor wants: 2 TMA
//    instructions per stage instead of 12.
//
// Synchronisation (per CTA): full[s] gets 1 arrival + the bytes of its own A/B loads AND of the peer's
// multicast half; empty[s] needs 2 arrivals — this CTA's and the peer's tcgen05.commit (multicast
// commit) — because a slot is overwritten by the peer's multicast as well.  The two CTAs of a cluster
// walk the same segment list, so they stay in lock-step by construction.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "syrk.h"
#include "umma.cuh"

namespace vlm {

namespace {

constexpr int kStages = 4;
constexpr int kStageBytes = 3 * kBlockBytes;  // [B0][B1][A]
constexpr int kStagingBytes = 16384;
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStagingBytes + 256 + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// L2 eviction policies (createpolicy): the activation X is a stream that is only re-read within a short window
// (by the other clusters sweeping the same rows), the Gram tiles are re-read and re-written by every K segment's
// reduce-add for the whole launch.
__device__ __forceinline__ uint64_t l2_policy(int kind) {  // 0 evict_normal, 1 evict_first, 2 evict_last
  uint64_t p;
  if (kind == 1)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// one box of a 3-D tensor map; lands in this CTA only
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint64_t* bar, void* smem_dst, int c0, int c1,
                                            int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "l"(pol)
      : "memory");
}
// same, delivered to the same smem offset (and signalling the same mbarrier offset) in every CTA of `mask`
__device__ __forceinline__ void tma_load_3d_mcast(const CUtensorMap* tm, uint64_t* bar, void* smem_dst, int c0,
                                                  int c1, int c2, uint16_t mask, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%4, %5, %6}], [%2], %3, %7;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
// 4-D variants for row-SEGMENTED activations {column in group, row in segment, column group, segment}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint64_t* bar, void* smem_dst, int c0, int c1,
                                            int c2, int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_mcast(const CUtensorMap* tm, uint64_t* bar, void* smem_dst, int c0,
                                                  int c1, int c2, int c3, uint16_t mask, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%4, %5, %6, %7}], [%2], %3, %8;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_hint(const CUtensorMap* tm, const void* smem_src, int c0, int c1,
                                                       uint64_t pol) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
      ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// One unit of work of a cluster in a BATCHED launch (several independent Gram problems in one grid): as PairSeg,
// plus which problem it belongs to (index into the tensor-map array) and that problem's column count.
// cps = chunks per row segment of a segmented activation (0: contiguous rows, 3-D tensor map).
struct BatchSeg {
  int32_t sa, sb, k0, k1, pid, d, cps, pad;
};
__device__ __forceinline__ void tensormap_acquire(const CUtensorMap* tm) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// BATCH = false: one problem, tensor maps in kernel parameters.  BATCH = true: the segments of several problems
// (e.g. all Grams of the text tower of one forward) share the grid; tensor maps live in global memory (`maps`:
// [2*pid] = X, [2*pid+1] = G), written by the host before the launch.
template <int ELEM_BYTES, int FMT, bool BATCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
syrk_tc2_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_g1,
                const CUtensorMap* __restrict__ maps, const void* __restrict__ segs_raw,
                const int* __restrict__ seg_off, int d1, int cps1, int l2_hints) {
  using G = Geo<ELEM_BYTES>;
  using Seg = typename std::conditional<BATCH, BatchSeg, PairSeg>::type;
  const Seg* __restrict__ segs = static_cast<const Seg*>(segs_raw);
  auto map_x = [&](const Seg& sg) -> const CUtensorMap* {
    if constexpr (BATCH) return maps + 2 * sg.pid; else return &tm_x1;
  };
  auto map_g = [&](const Seg& sg) -> const CUtensorMap* {
    if constexpr (BATCH) return maps + 2 * sg.pid + 1; else return &tm_g1;
  };
  auto cols_of = [&](const Seg& sg) -> int {
    if constexpr (BATCH) return sg.d; else return d1;
  };
  auto cps_of = [&](const Seg& sg) -> int {
    if constexpr (BATCH) return sg.cps; else return cps1;
  };
  constexpr int kRows = G::BK;            // rows of X per pipeline stage
  constexpr int kBlk = kBlockBytes;       // bytes of one 128-column block per stage
  constexpr int kBox = G::BOX_BYTES;      // bytes of one column group per stage (= LBO of the UMMA descriptor)
  constexpr int kNumMma = G::NUM_MMA;
  constexpr int kNStages = kStages;
  constexpr int kStageB = kStageBytes;    // [B0][B1][A]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kNStages * kStageB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kNStages;
  uint64_t* tfull = bars + 2 * kNStages;
  uint64_t* tempty = bars + 2 * kNStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 / 1: which row block of the super-tile
  const int cluster_id = blockIdx.x >> 1;
  const int seg_begin = seg_off[cluster_id];
  const int seg_end = seg_off[cluster_id + 1];

  if (!BATCH && warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x1);
    tma_prefetch_desc(&tm_g1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kNStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 2);  // this CTA's MMA commit + the peer's
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits visible cluster-wide before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    int cur_pid = -1;
    const uint64_t pol_x = l2_policy(l2_hints & 3);
    for (int s = seg_begin; s < seg_end; ++s) {
      const Seg seg = segs[s];
      const CUtensorMap* tm_x = map_x(seg);
      if constexpr (BATCH) {
        if (seg.pid != cur_pid) {
          tensormap_acquire(tm_x);
          cur_pid = seg.pid;
        }
      }
      const bool diag = seg.sa == seg.sb;
      const int a_group = (2 * seg.sa + (int)rank) * G::GB;       // first column group of this CTA's A block
      const int b_group = (2 * seg.sb + (int)rank) * G::GB;       // ... of the B block this CTA fetches
      const uint32_t bytes = (diag ? 2u : 3u) * kBlk;             // both B blocks (+ own A block)
      // segmented activation: chunk k = (row segment k / cps, chunk k % cps inside it); rows past the end of a
      // segment are zero-filled by TMA (they lie outside the tensor map's row dimension)
      const int cps = cps_of(seg);
      int xseg = cps > 0 ? seg.k0 / cps : 0;
      int kin = seg.k0 - xseg * cps;
      for (int k = seg.k0; k < seg.k1; ++k) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], bytes);
        uint8_t* sb = stage_base + stage * kStageB;
        const int row = kin * kRows;
        if (cps > 0) {
          tma_load_4d_mcast(tm_x, &full[stage], sb + rank * kBlk, 0, row, b_group, xseg, (uint16_t)0x3, pol_x);
          if (!diag) tma_load_4d(tm_x, &full[stage], sb + 2 * kBlk, 0, row, a_group, xseg, pol_x);
        } else {
          tma_load_3d_mcast(tm_x, &full[stage], sb + rank * kBlk, 0, row, b_group, (uint16_t)0x3, pol_x);
          if (!diag) tma_load_3d(tm_x, &full[stage], sb + 2 * kBlk, 0, row, a_group, pol_x);
        }
        if (++kin == cps) {
          kin = 0;
          ++xseg;
        }
        if (++stage == kNStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The WHOLE warp walks the loop and one elected lane issues: with warp-uniform control flow the descriptors
    // live in uniform registers.  (Issued from inside an `if (lane == 0)` region every tcgen05.mma cost ~23 SASS
    // instructions of ELECT / R2UR.BROADCAST / descriptor arithmetic, and the issue loop took about as long as
    // the MMAs themselves: ncu showed the issuing warp busy, not waiting, while the tensor pipe idled 16 %.)
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t stage0 = smem_u32(stage_base);
    // descriptor words that never change: high word = SBO | version | layout type, low word = LBO (start address added)
    constexpr uint32_t kDescHi = (uint32_t)((G::SBO_BYTES >> 4) & 0x3FFF) | (1u << 14) | ((uint32_t)G::LAYOUT_TYPE << 29);
    constexpr uint32_t kDescLo = (uint32_t)((kBox >> 4) & 0x3FFF) << 16;
    for (int s = seg_begin; s < seg_end; ++s) {
      const Seg seg = segs[s];
      const int sa_t = __shfl_sync(0xffffffffu, seg.sa, 0), sb_t = __shfl_sync(0xffffffffu, seg.sb, 0);
      const int k0 = __shfl_sync(0xffffffffu, seg.k0, 0), k1 = __shfl_sync(0xffffffffu, seg.k1, 0);
      const bool diag = sa_t == sb_t;
      // diagonal super-tile: CTA 1 forms only its diagonal block (B block 1, N = 128)
      const int n_off = (diag && rank == 1) ? 1 : 0;
      const uint32_t idesc = make_idesc(FMT, 128 * (2 - n_off));
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      const uint32_t a_off = diag ? rank * kBlk : 2 * kBlk;
      const uint32_t b_off = n_off * kBlk;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int k = k0; k < k1; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sb = stage0 + stage * kStageB;
        const uint32_t alo = (((sb + a_off) & 0x3FFFFu) >> 4) | kDescLo;
        const uint32_t blo = (((sb + b_off) & 0x3FFFFu) >> 4) | kDescLo;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kNumMma; ++kk) {
            const uint64_t adesc = ((uint64_t)kDescHi << 32) | (alo + kk * (G::KSTEP_BYTES >> 4));
            const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (blo + kk * (G::KSTEP_BYTES >> 4));
            umma<FMT>(d_tmem, adesc, bdesc, idesc, (k > k0 || kk > 0) ? 1u : 0u);
          }
          tc_commit_mcast(&empty[stage], (uint16_t)0x3);  // slot free in BOTH CTAs once these MMAs have read it
        }
        __syncwarp();
        if (++stage == kNStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) tc_commit(&tfull[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp - 4;
    const int epi_tid = threadIdx.x - 128;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_counter = 0;
    int cur_pid = -1;
    const uint64_t pol_g = l2_policy((l2_hints >> 2) & 3);
    for (int s = seg_begin; s < seg_end; ++s) {
      const Seg seg = segs[s];
      const CUtensorMap* tm_g = map_g(seg);
      const int d = cols_of(seg);
      if constexpr (BATCH) {
        if (epi_tid == 0 && seg.pid != cur_pid) {
          tensormap_acquire(tm_g);
          cur_pid = seg.pid;
        }
      }
      const bool diag = seg.sa == seg.sb;
      const int n_off = (diag && rank == 1) ? 1 : 0;
      const int w = 2 - n_off;
      const int row0 = (2 * seg.sa + (int)rank) * 128;
      const int col0 = (2 * seg.sb + n_off) * 128;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int nslab = (row0 < d) ? min(4 * w, (d - col0 + 31) / 32) : 0;
      // After this CTA's LAST segment the pipeline stages are dead (every load was consumed by the MMAs that
      // tfull just reported complete, in this CTA and — for the multicast halves — in the peer), so each slab
      // gets its own 16 KB staging slot there and no slab waits for an earlier TMA store to drain.
      const bool last = (s == seg_end - 1);
      for (int sl = 0; sl < nslab; ++sl) {
        uint8_t* buf = last ? stage_base + sl * kStagingBytes : staging + (slab_counter & 1) * kStagingBytes;
        if (!last) {
          if (epi_tid == 0) bulk_wait_group_read<1>();
          named_bar_sync(1, 128);
        }
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
          tma_reduce_add_2d_hint(tm_g, buf, col0 + sl * 32, row0, pol_g);
          bulk_commit_group();
        }
        ++slab_counter;
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (epi_tid == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's shared memory until here
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace vlm

#include "syrk_2sm.cuh"

namespace vlm {
namespace {
struct DeviceSchedule2 {
  int nclusters = 0;
  PairSeg* d_segs = nullptr;
  int* d_off = nullptr;
};
std::mutex g_mu2;
std::map<std::tuple<int, int64_t, int, int, int>, DeviceSchedule2> g_sched2;
// schedules evicted from the cache: freed at the NEXT eviction, after a device synchronisation, never while a
// launch that was handed their pointers may still be pending (the lock is held from lookup to launch)
std::vector<DeviceSchedule2> g_retired2;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode2 = nullptr;

int ensure_encode() {
  std::lock_guard<std::mutex> lk(g_mu2);
  if (!g_encode2) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VLM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VLM_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, VLM_ERR_DRIVER,
                "cuTensorMapEncodeTiled not available from the driver");
    g_encode2 = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return 0;
}

inline int elem_bytes(int dtype) { return (dtype == VLM_F32 || dtype == VLM_TF32X2) ? 4 : 2; }
// rows of X per pipeline stage = schedule chunk
inline int chunk_rows(int dtype) { return dtype == VLM_TF32X2 ? 16 : 128 / elem_bytes(dtype); }

// X as a 3-D tensor {column in group, row, column group} (strides: row pitch, 128 bytes); G as a 2-D fp32 tensor.
// seg_rows > 0: 4-D, {column in group, row in segment, column group, segment}.
// VLM_TF32X2: 4-D, {column in group, row, column group, plane}, planes rows * ldx elements apart, and one box
// fetches both planes of a 128-column block for 16 rows.
int encode_maps(const void* x, int dtype, int64_t rows, int d, int64_t ldx, int64_t seg_rows, int64_t seg_stride,
                float* g, int64_t ldg, CUtensorMap* tm_x, CUtensorMap* tm_g) {
  const int elem = elem_bytes(dtype);
  const int bk = chunk_rows(dtype), gc = 128 / elem;
  {
    // fp32 activations are described to TMA as TFLOAT32: the copy engine then ROUNDS each value to TF32 on its
    // way into shared memory.  With the plain FLOAT32 element type the tensor core truncates the low 13
    // mantissa bits of both operands instead, a systematic -6.8e-4 relative bias on every Gram (measured on the
    // B200, 36928 x 3072: rel. Frobenius error 7.6e-4 truncated vs 2.5e-5 rounded, same speed).
    // (The planes of VLM_TF32X2 hold TF32 values already: the conversion is then the identity.)
    const CUtensorMapDataType dt = elem == 4             ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                   : dtype == VLM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                       : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const bool split = dtype == VLM_TF32X2;
    const bool segmented = !split && seg_rows > 0;
    const int rank = (split || segmented) ? 4 : 3;
    cuuint64_t gdim[4] = {(cuuint64_t)gc, (cuuint64_t)(segmented ? seg_rows : rows), (cuuint64_t)(d / gc),
                          (cuuint64_t)(split ? 2 : segmented ? rows / seg_rows : 1)};
    cuuint64_t gstr[3] = {(cuuint64_t)ldx * elem, 128, (cuuint64_t)(split ? rows * ldx : seg_stride) * elem};
    cuuint32_t box[4] = {(cuuint32_t)gc, (cuuint32_t)bk, (cuuint32_t)elem /* groups per 128-column block */,
                         (cuuint32_t)(split ? 2 : 1)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle swz = elem == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = g_encode2(tm_x, dt, rank, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLM_REQUIRE(r == CUDA_SUCCESS, VLM_ERR_DRIVER, "cuTensorMapEncodeTiled(X, %d-D) failed: CUresult %d", rank, (int)r);
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)d};
    cuuint64_t gstr[1] = {(cuuint64_t)ldg * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode2(tm_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLM_REQUIRE(r == CUDA_SUCCESS, VLM_ERR_DRIVER, "cuTensorMapEncodeTiled(G) failed: CUresult %d", (int)r);
  }
  return 0;
}

int check_alignment(const void* x, int elem, int64_t rows, int64_t ldx, const float* g, int64_t ldg) {
  VLM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((ldx * elem) & 15) == 0, VLM_ERR_ALIGNMENT,
              "vlm_syrk_accum: x must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
  VLM_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (ldg & 3) == 0, VLM_ERR_ALIGNMENT,
              "vlm_syrk_accum: g must be 16-byte aligned with ldg %% 4 == 0");
  VLM_REQUIRE(rows < (int64_t)1 << 31, VLM_ERR_INVALID_ARG, "vlm_syrk_accum: rows too large");
  return 0;
}

// device scratch for batched launches: a small ring per (device, stream); reuse is ordered by the stream itself
struct Scratch {
  void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[4] = {0, 0, 0, 0};
  int next = 0;
};
std::map<std::pair<int, cudaStream_t>, Scratch> g_scratch;

// bits 0-1: policy of the X loads, bits 2-3: policy of the G reduce-adds (0 normal, 1 evict_first, 2 evict_last)
int l2_hints() {
  static const int v = [] {
    const char* e = getenv("VLM_SYRK_L2_HINTS");
    return e ? atoi(e) : (2 << 2);
  }();
  return v;
}

// Which pair kernel: 3 = one cta_group::2 instruction stream per pair (syrk_2sm.cuh, the default), 2 = two
// cta_group::1 streams with TMA multicast (kept selectable with VLM_SYRK_VARIANT=2 for A/B measurements).
// VLM_TF32X2 exists only in the 2-SM kernel.
int pair_variant(int dtype) {
  static const int v = [] {
    const char* e = getenv("VLM_SYRK_VARIANT");
    return (e && atoi(e) == 2) ? 2 : 3;
  }();
  return dtype == VLM_TF32X2 ? 3 : v;
}

typedef void (*PairKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap*, const void*, const int*, int, int,
                           int);
template <bool BATCH>
PairKernel pick_kernel(int dtype, int variant, int* smem) {
  if (variant == 2) {
    *smem = kSmemBytes;
    if (dtype == VLM_F32) return syrk_tc2_kernel<4, 2, BATCH>;
    if (dtype == VLM_BF16) return syrk_tc2_kernel<2, 1, BATCH>;
    return syrk_tc2_kernel<2, 0, BATCH>;
  }
  *smem = k2SmemBytes;
  if (dtype == VLM_TF32X2) return syrk_2sm_kernel<4, 2, BATCH, true>;
  if (dtype == VLM_F32) return syrk_2sm_kernel<4, 2, BATCH, false>;
  if (dtype == VLM_BF16) return syrk_2sm_kernel<2, 1, BATCH, false>;
  return syrk_2sm_kernel<2, 0, BATCH, false>;
}

}  // namespace

// Work decomposition for the CTA-pair kernel: K-ALIGNED tile ownership.
//
// Every cluster sweeps the rows of X from the top for "its" super-tile, so at any moment all clusters read
// the same thin band of rows: X streams from HBM once, every re-read (each column block is used by ~nsb
// tiles) hits L2, and no row panel has to fit in L2.  T super-tiles over C clusters:
//   * T >= C: floor(T/C) rounds of whole tiles (cluster c owns tiles c, c+C, ...), then the T mod C
//     left-over tiles are cut along K into C equal shares (stream-K) so that nobody idles;
//   * T <  C: every tile is cut along K into m = floor(C/T) equal pieces with the SAME boundaries for all
//     tiles (clusters on different tiles then stay aligned in K); T*m clusters are used.
// With at least one round, a cluster's left-over pieces are merged into its last sweep next to the own-tile
// segment covering the same rows (see below), so they too read the band of X that is in L2.  The reduce-adds
// into G carry an L2 evict_last policy: each tile is re-read and re-written by every K segment while ~45 MB of X
// stream through L2 in between (ncu, 36928 x 3072 fp32: DRAM read+write 636 MB without either, 526 MB with both).
// A K range longer than seg_cap chunks is processed as consecutive segments of at most seg_cap chunks:
// each ends with its own reduce-add into G (hidden behind the next segment's mainloop by the double-
// buffered TMEM accumulator), which bounds the tensor core's truncating fp32 accumulation.
// Measured against the alternatives on the B200 (36928 x 3072 fp32): panel-major stream-K with chunk-granular
// shares 672 TFLOP/s (many short segments whose 128 KB epilogues cannot hide), this schedule 753-758; an L2
// look-ahead prefetch (cp.async.bulk.prefetch.tensor) cost 8 %, half-height stages x 8 cost 14 %.  A variant
// issuing ONE tcgen05.mma.cta_group::2 (M = 256) per K step for the pair (32 KB stages x 6, no multicast) was
// bit-for-bit as accurate but ran at exactly half the speed (380 TFLOP/s, tensor pipe 37 % active in ncu) with
// these MN-major operands, so the pair keeps two independent cta_group::1 instruction streams.
void build_pair_schedule(int64_t kc, int d, int nclusters_max, std::vector<PairSeg>* segs, std::vector<int>* off) {
  // chunks per accumulation: 4096 fp32 rows / 8192 16-bit rows.  Measured on the B200 with all-positive
  // activations 36928 x 3072: cap 128 / 256 / 512 / none -> rel. error 7.6e-4 / 7.8e-4 / 8.2e-4 / 1.0e-3 (fp32),
  // 2.2e-5 / 5.0e-5 / 8.8e-5 (bf16), at 742 / 753 / 766 / 779 TFLOP/s.
  int64_t seg_cap = 128;
  if (const char* e = getenv("VLM_SYRK_SEG_CHUNKS")) seg_cap = std::max(1, atoi(e));
  const int nsb = (d + 255) / 256;
  struct T {
    int a, b;
  };
  std::vector<T> tiles;
  for (int a = 0; a < nsb; ++a)
    for (int b = a; b < nsb; ++b) tiles.push_back({a, b});
  const int64_t ntile = (int64_t)tiles.size();
  const int64_t min_chunks = 8;  // do not cut a K range below this: every piece pays a 128 KB epilogue per CTA
  const int C = nclusters_max;
  std::vector<std::vector<PairSeg>> per;
  auto emit = [&](int c, const T& t, int64_t k0, int64_t k1) {
    if (k1 <= k0) return;
    const int64_t pieces = (k1 - k0 + seg_cap - 1) / seg_cap;
    for (int64_t i = 0; i < pieces; ++i) {
      const int64_t a = k0 + (k1 - k0) * i / pieces, b = k0 + (k1 - k0) * (i + 1) / pieces;
      per[c].push_back({t.a, t.b, (int)a, (int)b});
    }
  };
  per.resize(C);
  const int64_t rounds = ntile / C;  // whole tiles per cluster, all clusters sweeping K from 0 together
  for (int64_t r = 0; r < rounds; ++r)
    for (int c = 0; c < C; ++c) emit(c, tiles[r * C + c], 0, kc);
  // The R = T mod C left-over tiles (all T tiles when T < C) are cut along K into P equal pieces with the
  // same boundaries for every tile; the (piece, tile) items are dealt round-robin in piece-major order,
  // so concurrently processed items belong to the same one or two pieces (K-aligned).  P minimises the
  // makespan ceil(R*P/C)/P among piece lengths between min_chunks.. and seg_cap chunks.
  const int64_t rem = ntile - rounds * C;
  if (rem > 0) {
    const int64_t p_lo = std::max<int64_t>(1, (kc + seg_cap - 1) / seg_cap);
    const int64_t p_hi = std::max<int64_t>(p_lo, kc / (kc >= 8 * 64 ? 64 : min_chunks));
    int64_t best_p = p_lo;
    double best = 1e30;
    for (int64_t p = p_lo; p <= p_hi; ++p) {
      const double makespan = (double)((rem * p + C - 1) / C) / (double)p;
      if (makespan < best - 1e-9) best = makespan, best_p = p;
    }
    // With at least one round of whole tiles, a left-over piece is not appended after the sweep (by then its rows
    // have long left L2: X is streamed once per round) but INTERLEAVED into the cluster's last sweep, right after
    // the own-tile segment that covers the same rows — every cluster keeps reading the one band of X that is in L2.
    std::vector<std::vector<PairSeg>> sweep;
    static const bool interleave_on = [] {
      const char* e = getenv("VLM_SYRK_INTERLEAVE");
      return e ? atoi(e) != 0 : true;
    }();
    const bool interleave = interleave_on && rounds > 0;
    if (interleave) {
      sweep.resize(C);
      for (int c = 0; c < C; ++c) {
        const size_t n_last = (size_t)((kc + seg_cap - 1) / seg_cap);  // segments of the last whole tile
        sweep[c].assign(per[c].end() - n_last, per[c].end());
        per[c].resize(per[c].size() - n_last);
      }
    }
    int64_t q = 0;
    for (int64_t p = 0; p < best_p; ++p)
      for (int64_t t = 0; t < rem; ++t, ++q)
        emit((int)(q % C), tiles[rounds * C + t], kc * p / best_p, kc * (p + 1) / best_p);
    if (interleave) {
      for (int c = 0; c < C; ++c) {
        // per[c] = earlier rounds + this cluster's left-over pieces (ascending k0); merge them into the sweep
        const size_t n_before = (size_t)(rounds - 1) * (size_t)((kc + seg_cap - 1) / seg_cap);
        std::vector<PairSeg> pieces(per[c].begin() + n_before, per[c].end());
        per[c].resize(n_before);
        size_t i = 0;
        for (const PairSeg& own : sweep[c]) {
          per[c].push_back(own);
          while (i < pieces.size() && pieces[i].k0 < own.k1) per[c].push_back(pieces[i++]);
        }
        while (i < pieces.size()) per[c].push_back(pieces[i++]);
      }
    }
  }
  while (!per.empty() && per.back().empty()) per.pop_back();  // clusters without work are not launched
  segs->clear();
  off->assign(1, 0);
  for (auto& v : per) {
    segs->insert(segs->end(), v.begin(), v.end());
    off->push_back((int)segs->size());
  }
}

// Host view for the CPU tests: segments as {super_row, super_col, k0, k1} per cluster.
void build_syrk_pair_schedule_host(int64_t kc, int d, int nsm, std::vector<int32_t>* flat, std::vector<int>* off) {
  std::vector<PairSeg> segs;
  build_pair_schedule(kc, d, nsm / 2, &segs, off);
  flat->clear();
  for (const PairSeg& s : segs) {
    flat->push_back(s.sa);
    flat->push_back(s.sb);
    flat->push_back(s.k0);
    flat->push_back(s.k1);
  }
}

bool syrk_tc2_supported(int dtype, int d, int64_t ldx) {
  return d % (128 / elem_bytes(dtype)) == 0 && ldx >= d;  // whole 128-byte column groups (the 3-D tensor map needs them)
}

// chunks per row segment (0 = contiguous) and in total
static inline void seg_chunks(int64_t rows, int64_t seg_rows, int bk, int64_t* cps, int64_t* kc) {
  if (seg_rows > 0) {
    *cps = (seg_rows + bk - 1) / bk;
    *kc = (rows / seg_rows) * *cps;
  } else {
    *cps = 0;
    *kc = (rows + bk - 1) / bk;
  }
}

int syrk_tc2_launch(const void* x, int dtype, int64_t rows, int d, int64_t ldx, int64_t seg_rows, int64_t seg_stride,
                    float* g, int64_t ldg, cudaStream_t stream) {
  const int elem = elem_bytes(dtype);
  const int bk = chunk_rows(dtype);
  if (int rc = check_alignment(x, elem, rows, ldx, g, ldg)) return rc;
  if (seg_rows >= rows || dtype == VLM_TF32X2) seg_rows = 0;  // one segment: plain rows
  int dev = 0, nsm = 0;
  VLM_CUDA(cudaGetDevice(&dev));
  if (int rc = device_sm_count(&nsm)) return rc;
  if (int rc = ensure_encode()) return rc;
  int64_t cps, kc;
  seg_chunks(rows, seg_rows, bk, &cps, &kc);
  VLM_REQUIRE(kc < (int64_t)1 << 30, VLM_ERR_INVALID_ARG, "vlm_syrk_accum: too many row chunks");
  CUtensorMap tm_x, tm_g;
  if (int rc = encode_maps(x, dtype, rows, d, ldx, seg_rows, seg_stride, g, ldg, &tm_x, &tm_g)) return rc;
  int smem = 0;
  PairKernel kernel = pick_kernel<false>(dtype, pair_variant(dtype), &smem);

  // lookup and launch under one lock: an eviction by another host thread cannot free a schedule between the two
  std::lock_guard<std::mutex> lk(g_mu2);
  auto key = std::make_tuple(dev, kc, d, bk, nsm);
  auto it = g_sched2.find(key);
  if (it == g_sched2.end()) {
    if (g_sched2.size() >= 512) {  // ragged workloads (a new row count every call): start over, do not grow forever
      VLM_CUDA(cudaDeviceSynchronize());  // launches that read the schedules retired LAST time have finished
      for (auto& ds : g_retired2) {
        cudaFree(ds.d_segs);
        cudaFree(ds.d_off);
      }
      g_retired2.clear();
      for (auto& kv : g_sched2) g_retired2.push_back(kv.second);
      g_sched2.clear();
    }
    std::vector<PairSeg> segs;
    std::vector<int> off;
    build_pair_schedule(kc, d, nsm / 2, &segs, &off);
    DeviceSchedule2 ds;
    ds.nclusters = (int)off.size() - 1;
    VLM_CUDA(cudaMalloc(&ds.d_segs, std::max<size_t>(1, segs.size()) * sizeof(PairSeg)));
    VLM_CUDA(cudaMalloc(&ds.d_off, off.size() * sizeof(int)));
    VLM_CUDA(cudaMemcpyAsync(ds.d_segs, segs.data(), segs.size() * sizeof(PairSeg), cudaMemcpyHostToDevice, stream));
    VLM_CUDA(cudaMemcpyAsync(ds.d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    VLM_CUDA(cudaStreamSynchronize(stream));
    it = g_sched2.emplace(key, ds).first;
  }
  const DeviceSchedule2& sched = it->second;
  VLM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kernel<<<2 * sched.nclusters, kThreads, smem, stream>>>(tm_x, tm_g, nullptr, sched.d_segs, sched.d_off, d, (int)cps,
                                                          l2_hints());
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int tf32_split_launch(const float* x, int64_t rows, int d, int64_t ldx, int64_t seg_rows, int64_t seg_stride, float* out,
                      cudaStream_t stream) {
  int nsm = 0;
  if (int rc = device_sm_count(&nsm)) return rc;
  if (seg_rows >= rows) seg_rows = 0;
  const int64_t n4 = rows * (d / 4);
  const int64_t blocks = std::min<int64_t>((n4 + 255) / 256, (int64_t)nsm * 8);
  tf32_split_kernel<<<(unsigned)std::max<int64_t>(1, blocks), 256, 0, stream>>>(x, rows, d, ldx, seg_rows, seg_stride, out);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int syrk_tc2_batch_launch(const vlm_syrk_problem* probs, int n, int dtype, cudaStream_t stream) {
  const int elem = elem_bytes(dtype);
  const int bk = chunk_rows(dtype);
  int dev = 0, nsm = 0;
  VLM_CUDA(cudaGetDevice(&dev));
  if (int rc = device_sm_count(&nsm)) return rc;
  if (int rc = ensure_encode()) return rc;
  const int C = nsm / 2;
  std::vector<CUtensorMap> maps(2 * (size_t)n);
  std::vector<std::vector<BatchSeg>> per(C);
  std::vector<int64_t> load(C, 0);
  // per-shape schedule (segments, per-cluster offsets, shares sorted heaviest first): a batch usually repeats a
  // handful of shapes (36 x [2560, 768] + 12 x [2560, 3072] for the text tower), so it is built once per shape
  struct ShapeSched {
    std::vector<PairSeg> segs;
    std::vector<int> off;
    std::vector<std::pair<int64_t, int>> shares;
  };
  std::map<std::pair<int64_t, int>, ShapeSched> by_shape;
  for (int p = 0; p < n; ++p) {
    const vlm_syrk_problem& q = probs[p];
    if (int rc = check_alignment(q.x, elem, q.rows, q.ldx, q.g, q.ldg)) return rc;
    const int64_t seg_rows = (dtype != VLM_TF32X2 && q.seg_rows > 0 && q.seg_rows < q.rows) ? q.seg_rows : 0;
    if (int rc = encode_maps(q.x, dtype, q.rows, q.d, q.ldx, seg_rows, q.seg_stride, q.g, q.ldg, &maps[2 * p],
                             &maps[2 * p + 1]))
      return rc;
    int64_t cps, kc;
    seg_chunks(q.rows, seg_rows, bk, &cps, &kc);
    auto found = by_shape.find({kc, q.d});
    if (found == by_shape.end()) {
      ShapeSched ss;
      build_pair_schedule(kc, q.d, C, &ss.segs, &ss.off);
      for (int c = 0; c + 1 < (int)ss.off.size(); ++c) {
        int64_t cost = 0;
        for (int s = ss.off[c]; s < ss.off[c + 1]; ++s) cost += ss.segs[s].k1 - ss.segs[s].k0;
        ss.shares.push_back({cost, c});
      }
      std::sort(ss.shares.begin(), ss.shares.end(),
                [](auto& a, auto& b) { return a.first > b.first || (a.first == b.first && a.second < b.second); });
      found = by_shape.emplace(std::make_pair(kc, q.d), std::move(ss)).first;
    }
    const std::vector<PairSeg>& segs = found->second.segs;
    const std::vector<int>& off = found->second.off;
    const std::vector<std::pair<int64_t, int>>& shares = found->second.shares;
    // this problem's per-cluster shares go to the least loaded clusters, heaviest share first
    std::vector<int> order(C);
    for (int c = 0; c < C; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return load[a] < load[b]; });
    for (size_t i = 0; i < shares.size(); ++i) {
      const int dst = order[i % C], c = shares[i].second;
      for (int s = off[c]; s < off[c + 1]; ++s)
        per[dst].push_back({segs[s].sa, segs[s].sb, segs[s].k0, segs[s].k1, p, q.d, (int)cps, 0});
      load[dst] += shares[i].first;
    }
  }
  int ncl = C;
  while (ncl > 0 && per[ncl - 1].empty()) --ncl;  // (clusters are filled least-loaded first, so gaps are rare)
  std::vector<BatchSeg> flat;
  std::vector<int> foff(1, 0);
  for (int c = 0; c < ncl; ++c) {
    flat.insert(flat.end(), per[c].begin(), per[c].end());
    foff.push_back((int)flat.size());
  }
  if (flat.empty()) return 0;
  const size_t maps_bytes = maps.size() * sizeof(CUtensorMap);
  const size_t off_bytes = ((foff.size() * sizeof(int)) + 127) / 128 * 128;
  const size_t seg_bytes = flat.size() * sizeof(BatchSeg);
  const size_t total = maps_bytes + off_bytes + seg_bytes;
  std::vector<uint8_t> host(total);
  memcpy(host.data(), maps.data(), maps_bytes);
  memcpy(host.data() + maps_bytes, foff.data(), foff.size() * sizeof(int));
  memcpy(host.data() + maps_bytes + off_bytes, flat.data(), seg_bytes);
  uint8_t* dptr = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu2);
    Scratch& sc = g_scratch[{dev, stream}];
    const int i = sc.next;
    sc.next = (sc.next + 1) % 4;
    if (sc.cap[i] < total) {
      if (sc.ptr[i]) VLM_CUDA(cudaFree(sc.ptr[i]));  // synchronises: nothing in flight can still read it
      sc.cap[i] = std::max<size_t>(total * 2, 1 << 18);
      VLM_CUDA(cudaMalloc(&sc.ptr[i], sc.cap[i]));
    }
    dptr = static_cast<uint8_t*>(sc.ptr[i]);
  }
  // pageable source: staged before the call returns; ordered on `stream` before the kernel below
  VLM_CUDA(cudaMemcpyAsync(dptr, host.data(), total, cudaMemcpyHostToDevice, stream));
  int smem = 0;
  PairKernel kernel = pick_kernel<true>(dtype, pair_variant(dtype), &smem);
  VLM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kernel<<<2 * ncl, kThreads, smem, stream>>>(maps[0], maps[1], reinterpret_cast<const CUtensorMap*>(dptr),
                                              dptr + maps_bytes + off_bytes,
                                              reinterpret_cast<const int*>(dptr + maps_bytes), 0, 0, l2_hints());
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace vlm
