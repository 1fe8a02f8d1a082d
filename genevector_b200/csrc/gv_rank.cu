// gv_rank.cu -- per-gene value sort of the gene-major (cell, value) segments.
//
// Used off the hot path only: the rank transform behind the Spearman target
// (genevector/metrics.py:423-431, scipy.stats.spearmanr = Pearson of average ranks) and the
// unique-value counts behind get_gene_entropy (genevector/data.py:186-203) for data that are not
// small integer counts.  The sort itself is CUB's segmented sort (library code, like cuBLAS); the
// rank / entropy arithmetic on the sorted segments is in gv_discretize.cu.
#include <cub/device/device_segmented_sort.cuh>
#include <cuda_runtime.h>
#include <cstdint>
#include "gv_internal.h"

namespace gv {

template <typename T>
static cudaError_t sort_impl(void* temp, size_t& temp_bytes, const void* keys_in, void* keys_out,
                             const int32_t* cells_in, int32_t* cells_out, int64_t nnz, int G,
                             const int64_t* gene_ptr, cudaStream_t st) {
  return cub::DeviceSegmentedSort::SortPairs(temp, temp_bytes, static_cast<const T*>(keys_in),
                                             static_cast<T*>(keys_out), cells_in, cells_out, int(nnz), G,
                                             gene_ptr, gene_ptr + 1, st);
}

size_t segsort_temp_bytes(int64_t nnz, int G, int value_dtype) {
  if (nnz <= 0 || nnz >= (int64_t(1) << 31) || G <= 0) return 0;
  size_t bytes = 0;
  cudaError_t e = value_dtype == GV_F64
      ? sort_impl<double>(nullptr, bytes, nullptr, nullptr, nullptr, nullptr, nnz, G, nullptr, nullptr)
      : sort_impl<float>(nullptr, bytes, nullptr, nullptr, nullptr, nullptr, nnz, G, nullptr, nullptr);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return bytes;
}

int segsort_pairs(void* temp, size_t temp_bytes, const void* keys_in, void* keys_out,
                  const int32_t* cells_in, int32_t* cells_out, int64_t nnz, int G,
                  const int64_t* gene_ptr, int value_dtype, cudaStream_t st) {
  if (nnz <= 0) return GV_OK;
  if (nnz >= (int64_t(1) << 31))
    return set_error(GV_ERR_UNSUPPORTED, "per-gene sort (ranks / entropy of non-count data) needs nnz < 2^31");
  cudaError_t e = value_dtype == GV_F64
      ? sort_impl<double>(temp, temp_bytes, keys_in, keys_out, cells_in, cells_out, nnz, G, gene_ptr, st)
      : sort_impl<float>(temp, temp_bytes, keys_in, keys_out, cells_in, cells_out, nnz, G, gene_ptr, st);
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  return GV_OK;
}

}  // namespace gv
