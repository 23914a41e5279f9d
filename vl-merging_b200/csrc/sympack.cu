// sympack.cu — packed-upper-triangle form of a Gram matrix (the on-disk container of gramfile.py).
//
// A Gram is symmetric, and the SYRK kernels only ever form its upper triangle; the reference nevertheless
// writes the full matrix in fp64 (src/cache_gram_matrices.py:251,349: 2.15 GB for VLMo-base, 7.65 GB for ViT-L).
// Packed row-major upper triangle in fp32 — row r holds columns r..d-1 at offset r*d - r(r-1)/2 — is a quarter
// of that.  Both kernels are HBM-bound copies; unpack goes through a 32x32 shared-memory tile so that the mirrored
// half is written with coalesced rows too.
#include "common.cuh"
#include "../../include/vlmerge.h"

namespace vlm {
namespace {

__device__ __forceinline__ int64_t packed_row_offset(int r, int d) {
  return (int64_t)r * d - ((int64_t)r * (r - 1)) / 2;
}

// one warp per row, rows interleaved over the grid (row r has d - r elements: interleaving balances the warps)
template <typename T>
__global__ void __launch_bounds__(256) sym_pack_kernel(const T* __restrict__ g, int d, int64_t ldg,
                                                       T* __restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < d; r += nwarp) {
    const T* src = g + (int64_t)r * ldg;
    T* dst = packed + packed_row_offset(r, d) - r;
    for (int c = r + lane; c < d; c += 32) dst[c] = src[c];
  }
}

// tile (bi, bj), bj >= bi: read the packed rows of the tile once, write out[r][c] and, transposed, out[c][r]
template <typename IN, typename OUT>
__global__ void __launch_bounds__(256) sym_unpack_kernel(const IN* __restrict__ packed, int d,
                                                         OUT* __restrict__ out, int64_t ldo) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  __shared__ IN t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const int r = bi * 32 + rr, c = bj * 32 + tx;
    IN v = 0;
    if (r < d && c < d) {
      v = (c >= r) ? packed[packed_row_offset(r, d) + (c - r)] : packed[packed_row_offset(c, d) + (r - c)];
      out[(int64_t)r * ldo + c] = (OUT)v;
    }
    t[rr][tx] = v;
  }
  if (bj == bi) return;
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int r = bj * 32 + rr, c = bi * 32 + tx;  // mirrored position, below the diagonal
    if (r < d && c < d) out[(int64_t)r * ldo + c] = (OUT)t[tx][rr];
  }
}

}  // namespace
}  // namespace vlm

using namespace vlm;

extern "C" int vlm_sym_pack_upper(const float* g, int d, int64_t ldg, float* packed, void* stream) {
  VLM_REQUIRE(g != nullptr && packed != nullptr && d > 0 && ldg >= d, VLM_ERR_INVALID_ARG,
              "vlm_sym_pack_upper: bad arguments");
  int nsm = 0;
  if (int rc = device_sm_count(&nsm)) return rc;
  const int blocks = std::min(nsm * 8, (d + 7) / 8);
  sym_pack_kernel<float><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, d, ldg, packed);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// fp64 Grams (the RegMean-grade cache modes): the exchange buffer of GramCache.all_reduce
extern "C" int vlm_sym_pack_upper_f64(const double* g, int d, int64_t ldg, double* packed, void* stream) {
  VLM_REQUIRE(g != nullptr && packed != nullptr && d > 0 && ldg >= d, VLM_ERR_INVALID_ARG,
              "vlm_sym_pack_upper_f64: bad arguments");
  int nsm = 0;
  if (int rc = device_sm_count(&nsm)) return rc;
  const int blocks = std::min(nsm * 8, (d + 7) / 8);
  sym_pack_kernel<double><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, d, ldg, packed);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int vlm_sym_unpack_f64(const double* packed, int d, double* out, int64_t ldo, void* stream) {
  VLM_REQUIRE(packed != nullptr && out != nullptr && d > 0 && ldo >= d, VLM_ERR_INVALID_ARG,
              "vlm_sym_unpack_f64: bad arguments");
  const int nt = (d + 31) / 32;
  sym_unpack_kernel<double, double><<<dim3(nt, nt), 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, d, out, ldo);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int vlm_sym_unpack(const float* packed, int d, void* out, int out_dtype, int64_t ldo, void* stream) {
  VLM_REQUIRE(packed != nullptr && out != nullptr && d > 0 && ldo >= d, VLM_ERR_INVALID_ARG,
              "vlm_sym_unpack: bad arguments");
  VLM_REQUIRE(out_dtype == VLM_F32 || out_dtype == VLM_F64, VLM_ERR_INVALID_ARG,
              "vlm_sym_unpack: out_dtype must be VLM_F32 or VLM_F64 (got %d)", out_dtype);
  const int nt = (d + 31) / 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_dtype == VLM_F32)
    sym_unpack_kernel<float, float><<<dim3(nt, nt), 256, 0, s>>>(packed, d, static_cast<float*>(out), ldo);
  else
    sym_unpack_kernel<float, double><<<dim3(nt, nt), 256, 0, s>>>(packed, d, static_cast<double*>(out), ldo);
  VLM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
