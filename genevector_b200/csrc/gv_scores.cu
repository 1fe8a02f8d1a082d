// gv_scores.cu -- matrix fast path of the score-building step.
//
// Replaces GeneVectorDataset._build_training_tensors (genevector/data.py:377-411): the reference
// walks the nested score dict in a double Python loop into a dense float32 matrix (:383-388),
// scales by c^2 (:390), clamps negatives when unsigned (:392-393) and flattens all off-diagonal
// ordered pairs in row-major order into (i_idx, j_idx, xij) (:396-411).  Here the dense device
// matrix from gv_mi_tiles / gv_moment_tiles goes straight to the three tensors: one thread per
// output element, coalesced writes.  HBM-bound: 4 B read + 20 B written per ordered pair.
#include <cuda_runtime.h>
#include <cstdint>
#include "gv_internal.h"

namespace gv {

__global__ void k_build_pairs(const float* __restrict__ scores, int64_t ld, int G, float c2,
                              int clamp_negative, int round5, int64_t* __restrict__ i_idx,
                              int64_t* __restrict__ j_idx, float* __restrict__ xij) {
  const int64_t total = int64_t(G) * (G - 1);
  const int gm1 = G - 1;
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += int64_t(gridDim.x) * blockDim.x) {
    const int i = int(p / gm1);
    const int r = int(p - int64_t(i) * gm1);
    const int j = r + (r >= i ? 1 : 0);
    float s = scores[int64_t(i) * ld + j];
    if (round5) s = float(rint(double(s) * 1e5) / 1e5);   // the dict stores round(score, 5): metrics.py:139
    s *= c2;                                              // float32 multiply like score_matrix *= c**2
    if (clamp_negative && s < 0.0f) s = 0.0f;
    i_idx[p] = i;
    j_idx[p] = j;
    xij[p] = s;
  }
}

int build_pairs_dispatch(const float* scores, int64_t ld, int G, float c2, int clamp_negative,
                         int round5, int64_t* i_idx, int64_t* j_idx, float* xij, cudaStream_t st) {
  if (G < 2) return GV_OK;
  const int64_t total = int64_t(G) * (G - 1);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = int64_t(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  k_build_pairs<<<unsigned(blocks), 256, 0, st>>>(scores, ld, G, c2, clamp_negative, round5, i_idx, j_idx, xij);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  return GV_OK;
}

}  // namespace gv
