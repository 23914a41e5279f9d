// merge.cu — kernel (b): one-pass streaming merge of flat fp32 parameter segments.
// Replaces the per-tensor ATen CPU loops of merge_weights (src/vilt/modules/vilt_module.py:586-635),
// sum_task_vectors (:696-744) and regmean's simple averages (:436-457, :486-529): every source
// element is read from HBM exactly once and every merged element written once, in ONE launch.
//
// HBM-bound (0.25-0.4 flop/byte): the design points are 128-bit streaming loads/stores, all of a
// thread's loads issued before the first use (4 vectors x n_src sources in flight per thread), a
// grid of SM-count x 4 persistent CTAs walking a chunk table, and no per-tensor launches.
// Rounding order is the reference's (separately rounded products and sums, no FMA contraction), so
// results are bit-identical to the torch CPU path.
#include <vector>

#include "common.cuh"

struct vlm_merge_plan {
  int n_seg = 0;
  int n_chunks = 0;
  uint64_t bytes = 0;
  vlm_merge_seg* d_segs = nullptr;
  int2* d_chunks = nullptr;  // (segment index, chunk index inside the segment)
  int device = 0;
};

namespace vlm {
namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;
constexpr int kChunkElems = kThreads * kVecPerThread * 4;  // 4096 fp32 = 16 KB per source per chunk

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int MODE>
__device__ __forceinline__ float combine(const float (&x)[VLM_MERGE_MAX_SRC], const float (&c)[VLM_MERGE_MAX_SRC],
                                         int n_src) {
  float acc;
  if (MODE == VLM_MERGE_WSUM) {
    acc = __fmul_rn(c[0], x[0]);
#pragma unroll
    for (int m = 1; m < VLM_MERGE_MAX_SRC; ++m)
      if (m < n_src) acc = __fadd_rn(acc, __fmul_rn(c[m], x[m]));
  } else if (MODE == VLM_MERGE_SEQ_LERP) {
    acc = x[0];
#pragma unroll
    for (int m = 1; m < VLM_MERGE_MAX_SRC; ++m)
      if (m < n_src) acc = __fadd_rn(acc, __fmul_rn(c[m], __fsub_rn(x[m], acc)));
  } else {
    acc = x[0];
#pragma unroll
    for (int m = 1; m < VLM_MERGE_MAX_SRC; ++m)
      if (m < n_src) acc = __fadd_rn(acc, x[m]);
    acc = __fdiv_rn(acc, (float)n_src);
  }
  return acc;
}

template <int MODE, int NSRC>
__device__ __forceinline__ float4 combine4(const float4 (&v)[NSRC], const float (&c)[VLM_MERGE_MAX_SRC]) {
  float xs[4][VLM_MERGE_MAX_SRC];
#pragma unroll
  for (int m = 0; m < NSRC; ++m) {
    xs[0][m] = v[m].x;
    xs[1][m] = v[m].y;
    xs[2][m] = v[m].z;
    xs[3][m] = v[m].w;
  }
  float4 o;
  o.x = combine<MODE>(xs[0], c, NSRC);
  o.y = combine<MODE>(xs[1], c, NSRC);
  o.z = combine<MODE>(xs[2], c, NSRC);
  o.w = combine<MODE>(xs[3], c, NSRC);
  return o;
}

template <int MODE, int NSRC>
__device__ __forceinline__ void run_chunk(const vlm_merge_seg& seg, uint64_t base, uint32_t n_here, bool aligned) {
  float c[VLM_MERGE_MAX_SRC];
#pragma unroll
  for (int m = 0; m < VLM_MERGE_MAX_SRC; ++m) c[m] = seg.coef[m];
  const float4* __restrict__ src[NSRC];
#pragma unroll
  for (int m = 0; m < NSRC; ++m) src[m] = reinterpret_cast<const float4*>(seg.src[m] + base);
  float4* __restrict__ dst = reinterpret_cast<float4*>(seg.dst + base);
  uint32_t done = 0;
  if (aligned && n_here == kChunkElems) {
    // full chunk: every load of a pass is issued before the first use.  VPT vectors x NSRC sources
    // are in flight per thread; VPT shrinks for 3-4 sources so they stay in registers at 4 CTAs/SM.
    constexpr int VPT = NSRC <= 2 ? kVecPerThread : 2;
#pragma unroll
    for (int pass = 0; pass < kVecPerThread / VPT; ++pass) {
      float4 v[VPT][NSRC];
#pragma unroll
      for (int u = 0; u < VPT; ++u)
#pragma unroll
        for (int m = 0; m < NSRC; ++m) v[u][m] = ld_stream(src[m] + (pass * VPT + u) * kThreads + threadIdx.x);
#pragma unroll
      for (int u = 0; u < VPT; ++u)
        st_stream(dst + (pass * VPT + u) * kThreads + threadIdx.x, combine4<MODE, NSRC>(v[u], c));
    }
    return;
  }
  if (aligned) {  // ragged last chunk of a segment: one vector at a time
    const uint32_t nvec = n_here >> 2;
    for (uint32_t i = threadIdx.x; i < nvec; i += kThreads) {
      float4 v[NSRC];
#pragma unroll
      for (int m = 0; m < NSRC; ++m) v[m] = ld_stream(src[m] + i);
      st_stream(dst + i, combine4<MODE, NSRC>(v, c));
    }
    done = nvec << 2;
  }
  // scalar tail (n % 4 elements), or the whole chunk when a pointer is not 16-byte aligned
  for (uint32_t e = done + threadIdx.x; e < n_here; e += kThreads) {
    float xs[VLM_MERGE_MAX_SRC];
#pragma unroll
    for (int m = 0; m < NSRC; ++m) xs[m] = seg.src[m][base + e];
    seg.dst[base + e] = combine<MODE>(xs, c, NSRC);
  }
}

template <int MODE>
__device__ __forceinline__ void dispatch_nsrc(const vlm_merge_seg& seg, uint64_t base, uint32_t n_here, bool al) {
  switch (seg.n_src) {
    case 1: run_chunk<MODE, 1>(seg, base, n_here, al); break;
    case 2: run_chunk<MODE, 2>(seg, base, n_here, al); break;
    case 3: run_chunk<MODE, 3>(seg, base, n_here, al); break;
    default: run_chunk<MODE, 4>(seg, base, n_here, al); break;
  }
}

__global__ void __launch_bounds__(kThreads, 4)
merge_segments_kernel(const vlm_merge_seg* __restrict__ segs, const int2* __restrict__ chunks, int n_chunks) {
  __shared__ vlm_merge_seg seg;
  for (int ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int2 cs = chunks[ch];
    __syncthreads();
    if (threadIdx.x < sizeof(vlm_merge_seg) / 4)
      reinterpret_cast<uint32_t*>(&seg)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&segs[cs.x])[threadIdx.x];
    __syncthreads();
    const uint64_t base = (uint64_t)cs.y * kChunkElems;
    const uint32_t n_here = (uint32_t)min((uint64_t)kChunkElems, seg.n - base);
    uintptr_t bits = reinterpret_cast<uintptr_t>(seg.dst);
    for (int m = 0; m < seg.n_src; ++m) bits |= reinterpret_cast<uintptr_t>(seg.src[m]);
    const bool aligned = (bits & 15) == 0;
    switch (seg.mode) {
      case VLM_MERGE_WSUM: dispatch_nsrc<VLM_MERGE_WSUM>(seg, base, n_here, aligned); break;
      case VLM_MERGE_SEQ_LERP: dispatch_nsrc<VLM_MERGE_SEQ_LERP>(seg, base, n_here, aligned); break;
      default: dispatch_nsrc<VLM_MERGE_MEAN>(seg, base, n_here, aligned); break;
    }
  }
}

}  // namespace
}  // namespace vlm

using namespace vlm;

extern "C" int vlm_merge_plan_create(const vlm_merge_seg* segs_host, int n_seg, vlm_merge_plan** out) {
  VLM_REQUIRE(out != nullptr && n_seg >= 0 && (segs_host != nullptr || n_seg == 0), VLM_ERR_INVALID_ARG,
              "vlm_merge_plan_create: bad arguments");
  std::vector<int2> chunks;
  uint64_t bytes = 0;
  for (int s = 0; s < n_seg; ++s) {
    const vlm_merge_seg& g = segs_host[s];
    VLM_REQUIRE(g.n_src >= 1 && g.n_src <= VLM_MERGE_MAX_SRC, VLM_ERR_INVALID_ARG,
                "vlm_merge_plan_create: segment %d has n_src=%d", s, g.n_src);
    VLM_REQUIRE(g.mode >= VLM_MERGE_WSUM && g.mode <= VLM_MERGE_MEAN, VLM_ERR_INVALID_ARG,
                "vlm_merge_plan_create: segment %d has mode=%d", s, g.mode);
    VLM_REQUIRE(g.dst != nullptr || g.n == 0, VLM_ERR_INVALID_ARG, "vlm_merge_plan_create: segment %d dst is NULL", s);
    for (int m = 0; m < g.n_src; ++m)
      VLM_REQUIRE(g.src[m] != nullptr || g.n == 0, VLM_ERR_INVALID_ARG,
                  "vlm_merge_plan_create: segment %d src[%d] is NULL", s, m);
    VLM_REQUIRE(g.n < ((uint64_t)1 << 31) * kChunkElems, VLM_ERR_INVALID_ARG, "segment %d too large", s);
    const uint64_t nch = (g.n + kChunkElems - 1) / kChunkElems;
    for (uint64_t c = 0; c < nch; ++c) chunks.push_back(make_int2(s, (int)c));
    bytes += (uint64_t)(g.n_src + 1) * g.n * 4;
  }
  vlm_merge_plan* p = new vlm_merge_plan();
  p->n_seg = n_seg;
  p->n_chunks = (int)chunks.size();
  p->bytes = bytes;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess && n_seg > 0) e = cudaMalloc(&p->d_segs, sizeof(vlm_merge_seg) * n_seg);
  if (e == cudaSuccess && !chunks.empty()) e = cudaMalloc(&p->d_chunks, sizeof(int2) * chunks.size());
  if (e == cudaSuccess && n_seg > 0)
    e = cudaMemcpy(p->d_segs, segs_host, sizeof(vlm_merge_seg) * n_seg, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !chunks.empty())
    e = cudaMemcpy(p->d_chunks, chunks.data(), sizeof(int2) * chunks.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(p->d_segs);
    cudaFree(p->d_chunks);
    delete p;
    return fail((int)e, "vlm_merge_plan_create: %s", cudaGetErrorString(e));
  }
  *out = p;
  return 0;
}

extern "C" int vlm_merge_plan_run(const vlm_merge_plan* plan, void* stream) {
  VLM_REQUIRE(plan != nullptr, VLM_ERR_INVALID_ARG, "vlm_merge_plan_run: plan is NULL");
  if (plan->n_chunks == 0) return 0;
  int nsm = 0;
  if (int rc = device_sm_count(&nsm)) return rc;
  const int grid = std::min(plan->n_chunks, nsm * 4);
  merge_segments_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(plan->d_segs, plan->d_chunks,
                                                                                  plan->n_chunks);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int vlm_merge_plan_destroy(vlm_merge_plan* plan) {
  if (!plan) return 0;
  cudaFree(plan->d_segs);
  cudaFree(plan->d_chunks);
  delete plan;
  return 0;
}

extern "C" uint64_t vlm_merge_plan_bytes(const vlm_merge_plan* plan) { return plan ? plan->bytes : 0; }

/* Many small copies in one call (staging a host checkpoint into the input arena: ~400 tensors, most of them a few
 * KB; issued one by one from Python they cost ~10 us of host time each, 4 ms of an 19 ms end-to-end merge). */
extern "C" int vlm_copy_batch(void* dst_base, const uint64_t* dst_off_bytes, const void* const* src_host_or_dev,
                              const uint64_t* nbytes, int n, void* stream) {
  VLM_REQUIRE(n >= 0 && (n == 0 || (dst_base && dst_off_bytes && src_host_or_dev && nbytes)), VLM_ERR_INVALID_ARG,
              "vlm_copy_batch: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n; ++i) {
    if (nbytes[i] == 0) continue;
    VLM_REQUIRE(src_host_or_dev[i] != nullptr, VLM_ERR_INVALID_ARG, "vlm_copy_batch: source %d is NULL", i);
    VLM_CUDA(cudaMemcpyAsync(static_cast<char*>(dst_base) + dst_off_bytes[i], src_host_or_dev[i], nbytes[i],
                             cudaMemcpyDefault, st));
  }
  return 0;
}
