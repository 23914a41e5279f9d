// syrk_tc.cu — kernel (a): G += X^T X on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// activation tiles staged by TMA, upper-triangular tile schedule, fp32 accumulation into the persistent
// Gram buffer with TMA reduce-add.  Replaces src/cache_gram_matrices.py:246-254 of the reference
// (fp64 cast + full DGEMM + synchronous .cpu() per hook call).
//
// Layout.  X is the hooked activation viewed as [rows, d] row-major, so for G = X^T X the contraction
// dimension (rows) is the SLOW one: both MMA operands are MN-major.  One TMA box is {128 bytes of
// columns, BK rows} with a 128-byte swizzle, which is exactly one column-group of the canonical
// MN-major UMMA layout  ((T,8,m),(R,k)) : ((1,T,LBO),(8T,SBO)):
//   LBO = BK*128 bytes (next 128-byte column group = next TMA box), SBO = R*128 bytes (next R rows);
//   R = 8 rows (SWIZZLE_128B) for bf16/f16, R = 4 rows (SWIZZLE_128B_BASE32B, TMA ..._ATOM_32B) for
//   TF32 — the only swizzled layout the tensor core takes for MN-major 32-bit operands.
// A 128-column block of X for one stage is therefore always 16 KB (fp32: 4 boxes x 32 rows,
// bf16/f16: 2 boxes x 64 rows) and one stage holds [B block 0][B block 1][A block].
//
// Work decomposition.  Output tiles are 128 x (128*w), w in {1,2}, over the block upper triangle
// (j >= i).  On a diagonal tile the A block IS the first B block, so it is loaded once.  The host
// builds a per-CTA segment list (build_syrk_schedule below): panel-major stream-K, i.e. all CTAs sweep
// the rows of X panel by panel (X is read from HBM once and re-read from L2) and the linearised
// (panel, tile, chunk) space is cut into equal shares.  Every segment ends with an fp32 reduce-add
// into G, which is also what "G +=" needs across hook calls.
//
// Warp roles (256 threads, 1 CTA / SM): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-7 = epilogue (TMEM -> registers -> swizzled smem -> cp.reduce.async.bulk.tensor).
// TMEM holds two 128x256 fp32 accumulators so the epilogue of one segment overlaps the next mainloop.
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

constexpr int kStages = 4;
constexpr int kStageBytes = 3 * kBlockBytes;  // [B0][B1][A]
constexpr int kStagingBytes = 16384;          // 128 rows x 32 fp32 (128-byte rows, swizzled)
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStagingBytes + 256 /*barriers*/ + 1024 /*align*/;

template <int ELEM_BYTES, int FMT>
__global__ void __launch_bounds__(kThreads, 1)
syrk_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
               const SyrkSeg* __restrict__ segs, const int* __restrict__ seg_off, int d) {
  using G = Geo<ELEM_BYTES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* full = bars;             // [kStages] TMA -> MMA
  uint64_t* empty = bars + kStages;  // [kStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kStages;       // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seg_begin = seg_off[blockIdx.x];
  const int seg_end = seg_off[blockIdx.x + 1];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    for (int s = seg_begin; s < seg_end; ++s) {
      const SyrkSeg seg = segs[s];
      const bool diag = seg.col_a == seg.col_b;
      // column groups that start inside the matrix (groups entirely past column d are not loaded:
      // whatever smem holds there only reaches output elements that the store clips)
      const int nb_groups = min(seg.w * G::GB, (d - seg.col_b + G::GC - 1) / G::GC);
      const int na_groups = diag ? 0 : min(G::GB, (d - seg.col_a + G::GC - 1) / G::GC);
      const uint32_t bytes = (uint32_t)(nb_groups + na_groups) * G::BOX_BYTES;
      for (int k = seg.k0; k < seg.k1; ++k) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], bytes);
        uint8_t* sb = stage_base + stage * kStageBytes;
        const int row = k * G::BK;
        for (int g = 0; g < nb_groups; ++g)
          tma_load_2d(&tm_x, &full[stage], sb + g * G::BOX_BYTES, seg.col_b + g * G::GC, row);
        for (int g = 0; g < na_groups; ++g)
          tma_load_2d(&tm_x, &full[stage], sb + 2 * kBlockBytes + g * G::BOX_BYTES, seg.col_a + g * G::GC, row);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer (one thread) =====
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int s = seg_begin; s < seg_end; ++s) {
      const SyrkSeg seg = segs[s];
      const bool diag = seg.col_a == seg.col_b;
      const uint32_t idesc = make_idesc(FMT, 128 * seg.w);
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int k = seg.k0; k < seg.k1; ++k) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sb = smem_u32(stage_base + stage * kStageBytes);
        const uint32_t sa = diag ? sb : sb + 2 * kBlockBytes;
#pragma unroll
        for (int kk = 0; kk < G::NUM_MMA; ++kk) {
          const uint64_t adesc = make_smem_desc<G::LAYOUT_TYPE>(sa + kk * G::KSTEP_BYTES, G::BOX_BYTES, G::SBO_BYTES);
          const uint64_t bdesc = make_smem_desc<G::LAYOUT_TYPE>(sb + kk * G::KSTEP_BYTES, G::BOX_BYTES, G::SBO_BYTES);
          umma<FMT>(d_tmem, adesc, bdesc, idesc, (k > seg.k0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit(&tfull[acc]);  // accumulator complete
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> swizzled smem -> TMA reduce-add into G =====
    const int q = warp - 4;  // == warp % 4: the TMEM lane quadrant this warp may read
    const int epi_tid = threadIdx.x - 128;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_counter = 0;
    for (int s = seg_begin; s < seg_end; ++s) {
      const SyrkSeg seg = segs[s];
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int nslab = min(4 * seg.w, (d - seg.col_b + 31) / 32);
      for (int sl = 0; sl < nslab; ++sl) {
        uint8_t* buf = staging + (slab_counter & 1) * kStagingBytes;
        if (epi_tid == 0) bulk_wait_group_read<1>();  // the store that last read `buf` is done with it
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
          tma_reduce_add_2d(&tm_g, buf, seg.col_b + sl * 32, seg.col_a);
          bulk_commit_group();
        }
        ++slab_counter;
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);  // 128 arrivals: accumulator drained
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (epi_tid == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- host: schedule -----------------------------------------------------------------------------

struct DeviceSchedule {
  int nctas = 0;
  SyrkSeg* d_segs = nullptr;
  int* d_off = nullptr;
};

std::mutex g_mu;
std::map<std::tuple<int, int64_t, int, int, int>, DeviceSchedule> g_sched;  // (dev, kc, d, bk, nsm)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int get_encode(EncodeTiledFn* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VLM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VLM_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, VLM_ERR_DRIVER,
                "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = g_encode;
  return 0;
}

template <int ELEM_BYTES, int FMT>
int launch_kernel(int dev, const DeviceSchedule& sched, const CUtensorMap& tm_x, const CUtensorMap& tm_g, int d,
                  cudaStream_t stream) {
  static std::atomic<bool> attr_done[64];
  auto kernel = syrk_tc_kernel<ELEM_BYTES, FMT>;
  if (dev >= 64 || !attr_done[dev].load(std::memory_order_acquire)) {
    VLM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    if (dev < 64) attr_done[dev].store(true, std::memory_order_release);
  }
  kernel<<<sched.nctas, kThreads, kSmemBytes, stream>>>(tm_x, tm_g, sched.d_segs, sched.d_off, d);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

// Exposed for tests (tests/test_schedule.py drives it through vlm_syrk_schedule_host).
//
// Panel-major stream-K.  The rows of X are cut into panels of kPanelChunks chunks.  Inside ONE panel
// the work units are the tiles in row-major order, each costing w * chunks_in_panel, and that linear
// cost axis is cut into ncta equal contiguous shares; every CTA then moves on to its share of the
// next panel.  Consequences: (1) at any moment all CTAs read the same row panel of X, so X streams
// from HBM once and every re-read (each column block is used by ~nb/2 tiles) hits L2; (2) load
// balance is exact whatever the tile count; (3) no accumulator runs over more than one panel of rows,
// which bounds the tensor core's truncating fp32 accumulation (the cross-panel sum is the
// round-to-nearest fp32 reduce-add into G).
void build_syrk_schedule(int64_t kc, int d, int nsm, std::vector<SyrkSeg>* segs, std::vector<int>* off) {
  int64_t panel_chunks = syrk_panel_chunks(d);
  const int nb = (d + 127) / 128;
  struct Tile {
    int i, j, w;
  };
  std::vector<Tile> tiles;
  int64_t tile_w_sum = 0;
  for (int i = 0; i < nb; ++i)
    for (int j = i; j < nb;) {
      const int w = std::min(2, nb - j);
      tiles.push_back({i, j, w});
      tile_w_sum += w;
      j += w;
    }
  const int64_t npanels = (kc + panel_chunks - 1) / panel_chunks;
  const int64_t total_cost = tile_w_sum * kc;
  // do not split below ~16 chunks of a wide tile per CTA: a segment's epilogue (128 KB reduce-add)
  // must stay small next to its mainloop
  const int64_t min_cost = 32;
  const int ncta = (int)std::max<int64_t>(1, std::min<int64_t>(nsm, total_cost / min_cost));

  std::vector<std::vector<SyrkSeg>> per_cta(ncta);
  for (int64_t p = 0; p < npanels; ++p) {
    const int64_t k_lo = p * panel_chunks, k_hi = std::min(kc, k_lo + panel_chunks);
    const int64_t panel_cost = tile_w_sum * (k_hi - k_lo);
    // rotate the CTA that gets the first share so that leftovers do not pile up on CTA 0
    const int rot = (int)((p * 37) % ncta);
    int share = 0;
    int64_t next_cut = panel_cost / ncta;
    int64_t pos = 0;
    for (const Tile& t : tiles) {
      int64_t k = k_lo;
      while (k < k_hi) {
        while (share < ncta - 1 && pos >= next_cut) {
          ++share;
          next_cut = panel_cost * (share + 1) / ncta;
        }
        const int64_t room = (share == ncta - 1) ? (k_hi - k) : (next_cut - pos + t.w - 1) / t.w;
        const int64_t take = std::min(k_hi - k, std::max<int64_t>(room, 1));
        per_cta[(share + rot) % ncta].push_back({t.i * 128, t.j * 128, t.w, (int)k, (int)(k + take)});
        k += take;
        pos += take * t.w;
      }
    }
  }
  segs->clear();
  off->assign(1, 0);
  for (int c = 0; c < ncta; ++c) {
    segs->insert(segs->end(), per_cta[c].begin(), per_cta[c].end());
    off->push_back((int)segs->size());
  }
}

int syrk_tc_launch(const void* x, int dtype, int64_t rows, int d, int64_t ldx, float* g, int64_t ldg,
                   cudaStream_t stream) {
  const int elem = (dtype == VLM_F32) ? 4 : 2;
  const int bk = 128 / elem;
  VLM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((ldx * elem) & 15) == 0, VLM_ERR_ALIGNMENT,
              "vlm_syrk_accum: x must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
  VLM_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (ldg & 3) == 0, VLM_ERR_ALIGNMENT,
              "vlm_syrk_accum: g must be 16-byte aligned with ldg %% 4 == 0");
  VLM_REQUIRE(rows < (int64_t)1 << 31, VLM_ERR_INVALID_ARG, "vlm_syrk_accum: rows too large");
  int dev = 0, nsm = 0;
  VLM_CUDA(cudaGetDevice(&dev));
  if (int rc = device_sm_count(&nsm)) return rc;
  EncodeTiledFn encode = nullptr;
  if (int rc = get_encode(&encode)) return rc;

  const int64_t kc = (rows + bk - 1) / bk;
  DeviceSchedule sched;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, kc, d, bk, nsm);
    auto it = g_sched.find(key);
    if (it == g_sched.end()) {
      if (g_sched.size() >= 512) {  // ragged workloads: start over instead of growing forever
        VLM_CUDA(cudaDeviceSynchronize());
        for (auto& kv : g_sched) {
          cudaFree(kv.second.d_segs);
          cudaFree(kv.second.d_off);
        }
        g_sched.clear();
      }
      std::vector<SyrkSeg> segs;
      std::vector<int> off;
      build_syrk_schedule(kc, d, nsm, &segs, &off);
      DeviceSchedule ds;
      ds.nctas = (int)off.size() - 1;
      VLM_CUDA(cudaMalloc(&ds.d_segs, std::max<size_t>(1, segs.size()) * sizeof(SyrkSeg)));
      VLM_CUDA(cudaMalloc(&ds.d_off, off.size() * sizeof(int)));
      // stream-ordered w.r.t. the launch below; the vectors are pageable, so the copy is staged
      // before cudaMemcpyAsync returns
      VLM_CUDA(cudaMemcpyAsync(ds.d_segs, segs.data(), segs.size() * sizeof(SyrkSeg), cudaMemcpyHostToDevice, stream));
      VLM_CUDA(cudaMemcpyAsync(ds.d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
      VLM_CUDA(cudaStreamSynchronize(stream));
      it = g_sched.emplace(key, ds).first;
    }
    sched = it->second;
  }

  CUtensorMap tm_x, tm_g;
  {
    // fp32 activations are described to TMA as TFLOAT32: the copy engine then ROUNDS each value to TF32 on its
    // way into shared memory.  With the plain FLOAT32 element type the tensor core truncates the low 13
    // mantissa bits of both operands instead, a systematic -6.8e-4 relative bias on every Gram (measured on the
    // B200, 36928 x 3072: rel. Frobenius error 7.6e-4 truncated vs 2.5e-5 rounded, same speed).
    const CUtensorMapDataType dt = dtype == VLM_F32    ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                   : dtype == VLM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                       : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ldx * elem};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elem), (cuuint32_t)bk};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle swz = elem == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = encode(&tm_x, dt, 2, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLM_REQUIRE(r == CUDA_SUCCESS, VLM_ERR_DRIVER, "cuTensorMapEncodeTiled(X) failed: CUresult %d", (int)r);
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)d};
    cuuint64_t gstr[1] = {(cuuint64_t)ldg * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLM_REQUIRE(r == CUDA_SUCCESS, VLM_ERR_DRIVER, "cuTensorMapEncodeTiled(G) failed: CUresult %d", (int)r);
  }

  if (dtype == VLM_F32) return launch_kernel<4, 2>(dev, sched, tm_x, tm_g, d, stream);
  if (dtype == VLM_BF16) return launch_kernel<2, 1>(dev, sched, tm_x, tm_g, d, stream);
  return launch_kernel<2, 0>(dev, sched, tm_x, tm_g, d, stream);
}

}  // namespace vlm
