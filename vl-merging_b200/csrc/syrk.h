// syrk.h — declarations shared by the SYRK translation units.
#pragma once
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

namespace vlm {

// One unit of work of the tcgen05 SYRK kernel: output tile rows [col_a, col_a+128), columns
// [col_b, col_b+128*w) of G, accumulated over row chunks [k0, k1) of X (chunk = BK rows).
struct SyrkSeg {
  int32_t col_a, col_b, w, k0, k1;
};

void build_syrk_schedule(int64_t kc, int d, int nsm, std::vector<SyrkSeg>* segs, std::vector<int>* off);

int syrk_tc_launch(const void* x, int dtype, int64_t rows, int d, int64_t ldx, float* g, int64_t ldg,
                   cudaStream_t stream);
int syrk_simt_launch(const void* x, int dtype, int64_t rows, int d, int64_t ldx, float* g, int64_t ldg,
                     cudaStream_t stream);
int sym_finalize_launch(float* g, int d, int64_t ldg, double* out_f64, int64_t ld64, cudaStream_t stream);

}  // namespace vlm
