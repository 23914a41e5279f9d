// syrk_i8.cuh — kernel (a), exact variant on the INTEGER tensor cores ("Ozaki" splitting), included by syrk_pair.cu.
//
// The RegMean-grade Gram needs ~1e-8 (regmean amplifies Gram error a thousandfold on near-singular
// sums), which the fp32 accumulator of the floating-point tensor pipes cannot give at any segment length, and the
// fp64 DMMA path (syrk_f64.cu) gives at 30 TFLOP/s.  Integer MMAs accumulate EXACTLY:
//
//   x[r][c]  ~=  q[r][c] * 2^(E_c - 26),   q = rint(x * 2^(26 - E_c)),  |q| < 2^26,  2^E_c > max_r |x[r][c]|
//   q = D0 * 2^21 + D1 * 2^14 + D2 * 2^7 + D3,   balanced digits D_p in [-64, 64]   (four int8 planes)
//   G[i][j] = 2^(E_i + E_j - 52) * sum_{p,q} 2^(7 (6 - p - q)) * (D_p[:, i] . D_q[:, j])
//
// The dot products of digit planes are int8 x int8 -> int32 tensor-core products (tcgen05.mma kind::i8, exact: 2^12 per
// product); the thirteen pairs with p + q <= 4 are kept.  (Group 4 matters: its D_2 . D_2 term is a sum of SQUARES on
// the diagonal — with p + q <= 3 only, normally distributed activations came out 1.7e-7 off, on the B200 and in a numpy
// emulation alike.  What is dropped now is below 1e-9 of sqrt(G_ii G_jj).)  Pairs of one group s = p + q share a scale, so a segment accumulates ONE group in one
// int32 TMEM accumulator and its epilogue adds ldexp(acc, E_i + E_j - 10 - 7 s) to the fp64 Gram (red.global.add.f64,
// coalesced through a per-warp shared-memory transpose: the split-K reduction, the sum over groups and the `+=`
// across hook calls in one).  Quantisation error: 2^-27 of the
// column maximum per element, unbiased.
//
// Same CTA-pair structure as syrk_2sm_kernel (one tcgen05.mma.cta_group::2 stream, M = 256, N = 256; leader-owned
// full / tempty barriers; six 32 KB stages, MN-major SWIZZLE_128B planes).  Every MMA consumes one 4 KB plane of A
// and of B per CTA for 128 tensor cycles — 64 B/clk, the SM's ingest limit — unless loaded planes are reused, so a
// segment forms up to TWO groups at once, one per 256-column TMEM accumulator, in three phases per K range:
//   phase 0 (groups 4 and 3): a stage = 32 rows of all four planes ([A: 4 x 4 KB][B: 4 x 4 KB]), 7 MMAs per 32 KB;
//   phase 1 (groups 2 and 1): a stage = 32 rows of planes 0..2 ([A: 3 x 4 KB][B: 3 x 4 KB]), 5 MMAs per 24 KB;
//   phase 2 (group 0): a stage = 128 rows of plane 0 ([A: 16 KB][B: 16 KB]), 4 MMAs per 32 KB (ingest-bound, 1/13 of
//   the work);
// one TMA box per operand per stage (three tensor maps over the same planes).  The accumulators belong to the
// running segment, so its epilogue is not hidden (~10 % with 65536-row segments; int32 holds 2^17 rows of a
// four-pair group).  (First version, one group per segment with 4 KB boxes: tensor pipe 43 % of elapsed in ncu.)
#pragma once

namespace vlm {
namespace {

constexpr int kI8PlaneBytes = 32 * 128;   // 32 rows x 128 int8 columns
constexpr int kI8Planes = 4;
constexpr int kI8Phases = 3;              // {groups 4, 3}, {groups 2, 1}, {group 0}
constexpr int kI8SmemBytes = k2Stages * k2StageBytes + 1280 + 4 * 32 * 33 * 4 + 1024;   // stages, barriers, 4 transpose tiles

struct I8Args {
  double* g;
  int64_t ldg;
  const int* exps;   // E_c per column
};

// D s32, A/B signed 8-bit, both MN-major, M = 256 over the pair
__host__ __device__ constexpr uint32_t make_idesc_i8(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void umma2_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// segs: PairSeg with the phase (0: groups 4 and 3, 1: groups 2 and 1, 2: group 0) in bits 16.. of sb; k in 32-row chunks
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
syrk_i8x4_kernel(const __grid_constant__ CUtensorMap tm_p0, const __grid_constant__ CUtensorMap tm_p1,
                 const __grid_constant__ CUtensorMap tm_p2, const PairSeg* __restrict__ segs,
                 const int* __restrict__ seg_off, int d, const __grid_constant__ I8Args args) {
  constexpr int kBlk = kBlockBytes;       // one operand of a stage
  constexpr int kNS = k2Stages;
  constexpr int kStageB = k2StageBytes;   // [A][B]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kNS * kStageB);
  uint64_t* full = bars;                  // used in the leader only
  uint64_t* empty = bars + kNS;
  uint64_t* tfull = bars + 2 * kNS;
  uint64_t* tempty = bars + 2 * kNS + 1;  // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNS + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int seg_begin = seg_off[cluster_id];
  const int seg_end = seg_off[cluster_id + 1];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_p0);
    tma_prefetch_desc(&tm_p1);
    tma_prefetch_desc(&tm_p2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kNS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem2_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs): one box for its own A block, one for its own half of B =====
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_x = l2_policy(0);
    const uint32_t full0 = mapa_rank(smem_u32(full), 0);
    for (int s = seg_begin; s < seg_end; ++s) {
      const PairSeg seg = segs[s];
      const int sb_t = seg.sb & 0xFFFF, ph = seg.sb >> 16;
      const bool diag = seg.sa == sb_t;
      const CUtensorMap* tm = ph == 0 ? &tm_p0 : ph == 1 ? &tm_p1 : &tm_p2;
      const int a_group = 2 * seg.sa + (int)rank;
      const int b_group = 2 * sb_t + (int)rank;
      const uint32_t bytes_pair = (diag ? 2u : 4u) * (ph == 1 ? 3u * kI8PlaneBytes : (uint32_t)kBlk);
      for (int k = seg.k0; k < seg.k1; k += ph == 2 ? 4 : 1) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&full[stage], bytes_pair);
        uint8_t* sb = stage_base + stage * kStageB;
        const uint32_t bar = full0 + (uint32_t)stage * 8u;
        tma2_load_4d(tm, bar, sb + kBlk, 0, k * 32, b_group, 0, pol_x);
        if (!diag) tma2_load_4d(tm, bar, sb, 0, k * 32, a_group, 0, pol_x);
        if (++stage == kNS) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader only) =====
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t stage0 = smem_u32(stage_base);
    // MN-major SWIZZLE_128B: 8-row atoms 1024 bytes apart (SBO); one 128-byte column group per CTA, LBO unused
    constexpr uint32_t kDescHi = (uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
    constexpr uint32_t kDescLo = (uint32_t)((kI8PlaneBytes >> 4) & 0x3FFF) << 16;
    constexpr uint32_t idesc = make_idesc_i8(256);
    constexpr uint32_t kP = kI8PlaneBytes >> 4;     // 32 rows of one plane, in descriptor units
    for (int s = seg_begin; s < seg_end; ++s) {
      const PairSeg seg = segs[s];
      const int sa_t = __shfl_sync(0xffffffffu, seg.sa, 0), sbw = __shfl_sync(0xffffffffu, seg.sb, 0);
      const int k0 = __shfl_sync(0xffffffffu, seg.k0, 0), k1 = __shfl_sync(0xffffffffu, seg.k1, 0);
      const int ph = sbw >> 16;
      const uint32_t a_off = (sa_t == (sbw & 0xFFFF)) ? kBlk : 0u;
      const uint32_t d_hi = tmem_base, d_lo = tmem_base + kAccCols;   // the higher / lower group of this phase
      mbar_wait(tempty, acc_phase ^ 1);
      tc_fence_after();
      for (int k = k0; k < k1; k += ph == 2 ? 4 : 1) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sb = stage0 + stage * kStageB;
        const uint32_t alo = (((sb + a_off) & 0x3FFFFu) >> 4) | kDescLo;
        const uint32_t blo = (((sb + kBlk) & 0x3FFFFu) >> 4) | kDescLo;
        if (elect_one()) {
          auto mma = [&](uint32_t dst, uint32_t a_units, uint32_t b_units, uint32_t accumulate) {
            umma2_i8(dst, ((uint64_t)kDescHi << 32) | (alo + a_units), ((uint64_t)kDescHi << 32) | (blo + b_units), idesc,
                     accumulate);
          };
          const uint32_t first = (k == k0) ? 0u : 1u;
          if (ph == 0) {
            // planes p at p * 4 KB.  group 4: (1,3) (2,2) (3,1); group 3: (0,3) (1,2) (2,1) (3,0)
            mma(d_hi, 1 * kP, 3 * kP, first);
            mma(d_hi, 2 * kP, 2 * kP, 1u);
            mma(d_hi, 3 * kP, 1 * kP, 1u);
            mma(d_lo, 0 * kP, 3 * kP, first);
            mma(d_lo, 1 * kP, 2 * kP, 1u);
            mma(d_lo, 2 * kP, 1 * kP, 1u);
            mma(d_lo, 3 * kP, 0 * kP, 1u);
          } else if (ph == 1) {
            // planes 0..2.  group 2: (0,2) (1,1) (2,0); group 1: (0,1) (1,0)
            mma(d_hi, 0 * kP, 2 * kP, first);
            mma(d_hi, 1 * kP, 1 * kP, 1u);
            mma(d_hi, 2 * kP, 0 * kP, 1u);
            mma(d_lo, 0 * kP, 1 * kP, first);
            mma(d_lo, 1 * kP, 0 * kP, 1u);
          } else {
            // 128 rows of plane 0: up to four K steps of 32 rows, as many as still lie inside this segment's K range
            const int nks = min(4, k1 - k);
            for (int ks = 0; ks < nks; ++ks) mma(d_hi, ks * kP, ks * kP, ks == 0 ? first : 1u);
          }
          tc2_commit_mcast(&empty[stage]);
        }
        __syncwarp();
        if (++stage == kNS) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) tc2_commit_mcast(tfull);
      __syncwarp();
      acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): exact int32 sums of the two groups -> scaled fp64 adds into G =====
    // A thread owns one accumulator ROW, so adding straight from its registers would touch 32 rows of G per warp
    // instruction.  Each warp transposes its 32 x 32 chunk through shared memory and lane j adds column j of all 32
    // rows: one contiguous 256-byte red.global.add.f64 per instruction.
    const int q = warp - 4;
    int* tile = reinterpret_cast<int*>(smem + kNS * kStageB + 1280) + q * (32 * 33);
    uint32_t acc_phase = 0;
    const uint32_t tempty0 = mapa_rank(smem_u32(tempty), 0);
    for (int s = seg_begin; s < seg_end; ++s) {
      const PairSeg seg = segs[s];
      const int sb_t = seg.sb & 0xFFFF, ph = seg.sb >> 16;
      const bool diag = seg.sa == sb_t;
      const int n_off = (diag && rank == 1) ? 1 : 0;   // peer on a diagonal tile: only block (2a+1, 2a+1)
      const int row_base = (2 * seg.sa + (int)rank) * 128 + q * 32;
      const int col0 = (2 * sb_t + n_off) * 128;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
      const int e_lane = (row_base + lane < d) ? __ldg(args.exps + row_base + lane) - 10 : 0;
      const int nchunk = min(4 * (2 - n_off), (d - col0 + 31) / 32);
      for (int which = 0; which < (ph == 2 ? 1 : 2); ++which) {
        const int grp = 4 - 2 * ph - which;            // accumulator 0: the higher group of the phase
        for (int ch = 0; ch < nchunk; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + which * kAccCols + n_off * 128 + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = (int)v[j];
          __syncwarp();
          const int c = col0 + ch * 32 + lane;
          const int e_col = (c < d ? __ldg(args.exps + c) : 0) - 7 * grp;
#pragma unroll 4
          for (int r = 0; r < 32; ++r) {
            const int iv = tile[r * 33 + lane];
            const int e_row = __shfl_sync(0xffffffffu, e_lane, r);
            if (row_base + r < d && c < d && iv != 0)
              atomicAdd(args.g + (int64_t)(row_base + r) * args.ldg + c, ldexp((double)iv, e_row + e_col));
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem2_dealloc(tmem_base, kTmemCols);
}

// ---- pre-pass: column maxima -> exponents, then the four digit planes ---------------------------------------------
__device__ __forceinline__ const float* i8_row_ptr(const float* x, int64_t r, int64_t ldx, int64_t seg_rows,
                                                    int64_t seg_stride) {
  return seg_rows > 0 ? x + (r / seg_rows) * seg_stride + (r % seg_rows) * ldx : x + r * ldx;
}

// grid (d / 128, row slabs); thread t of 32 x 8: columns 4 (t & 31) .. +3 of the block, rows (t >> 5) + 8 i of the slab
__global__ void __launch_bounds__(256) i8_colmax_kernel(const float* __restrict__ x, int64_t rows, int d, int64_t ldx,
                                                        int64_t seg_rows, int64_t seg_stride, int64_t rows_per_slab,
                                                        unsigned* __restrict__ amax_bits) {
  const int c = blockIdx.x * 128 + (threadIdx.x & 31) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab, r1 = min(rows, r0 + rows_per_slab);
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  for (int64_t r = r0 + (threadIdx.x >> 5); r < r1; r += 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(i8_row_ptr(x, r, ldx, seg_rows, seg_stride) + c));
    m0 = fmaxf(m0, fabsf(v.x)), m1 = fmaxf(m1, fabsf(v.y)), m2 = fmaxf(m2, fabsf(v.z)), m3 = fmaxf(m3, fabsf(v.w));
  }
  // non-negative floats order like their bit patterns
  atomicMax(amax_bits + c, __float_as_uint(m0));
  atomicMax(amax_bits + c + 1, __float_as_uint(m1));
  atomicMax(amax_bits + c + 2, __float_as_uint(m2));
  atomicMax(amax_bits + c + 3, __float_as_uint(m3));
}

__global__ void i8_exps_kernel(const unsigned* __restrict__ amax_bits, int d, int* __restrict__ exps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float m = __uint_as_float(amax_bits[c]);
  exps[c] = (m > 0.f && isfinite(m)) ? ilogbf(m) + 1 : 0;     // 2^E > m
}

__device__ __forceinline__ void i8_digits(float x, int e, int8_t (&dg)[4]) {
  int t = __float2int_rn(ldexpf(x, 26 - e));   // |t| < 2^26 (+1 from rounding at the very top)
#pragma unroll
  for (int p = 3; p > 0; --p) {
    const int r = ((t + 64) & 127) - 64;       // balanced digit in [-64, 63]
    dg[p] = (int8_t)r;
    t = (t - r) >> 7;
  }
  dg[0] = (int8_t)t;                           // |t| <= 33
}

// one thread: 4 consecutive columns of one row -> a char4 in each of the four planes ([plane][row][d])
__global__ void __launch_bounds__(256) i8_slice_kernel(const float* __restrict__ x, int64_t rows, int d, int64_t ldx,
                                                       int64_t seg_rows, int64_t seg_stride,
                                                       const int* __restrict__ exps, int8_t* __restrict__ planes) {
  const int d4 = d >> 2;
  const int64_t n4 = rows * d4;
  const int64_t plane = rows * (int64_t)d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d4;
    const int c = (int)(i - r * d4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(i8_row_ptr(x, r, ldx, seg_rows, seg_stride) + c));
    const int4 e = __ldg(reinterpret_cast<const int4*>(exps + c));
    int8_t a[4], b[4], cc[4], dd[4];
    i8_digits(v.x, e.x, a);
    i8_digits(v.y, e.y, b);
    i8_digits(v.z, e.z, cc);
    i8_digits(v.w, e.w, dd);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      char4 o;
      o.x = a[p], o.y = b[p], o.z = cc[p], o.w = dd[p];
      *reinterpret_cast<char4*>(planes + p * plane + r * d + c) = o;
    }
  }
}

}  // namespace
}  // namespace vlm
