// gv_internal.h -- declarations shared by the translation units of libgvb200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "../../include/gvb200.h"

namespace gv {

constexpr int MARG_STRIDE_ = GV_MARG_STRIDE;   // int32 per gene in `marg`
constexpr int EDGE_STRIDE_ = GV_EDGE_STRIDE;   // float64 per gene in `edges`
constexpr int QT_DIM_ = GV_QTABLE_DIM;         // the quantile-fraction table is [QT_DIM_][QT_DIM_]

int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* file, int line);
int device_sm_count();

struct DiscretizeArgs {
  const void* values; int value_dtype;
  const int32_t* col_idx; const int64_t* row_ptr;
  int64_t n_cells; int n_genes; int64_t nnz; int n_bins;
  const double* q_table;            // host, [32][32]
  void* bins_packed; int64_t bins_pitch_bytes;
  int32_t* n_bins_per_gene; int32_t* marg; double* edges;
  int64_t* gsum; int64_t* gsq;
  void* val_rows; int64_t val_pitch; int64_t val_rows_alloc;
  int val_mode; int val_limbs; int val_limb_bits;
  uint64_t* data_flags_out;         // host, [2]
  void* workspace; size_t workspace_bytes;
  int path;                         // GV_PATH_AUTO / GV_PATH_COUNTS / GV_PATH_GENERAL
  int* path_used;                   // host, optional
};

int discretize_dispatch(const DiscretizeArgs& a, cudaStream_t st);
size_t discretize_workspace_bytes(int64_t nnz, int64_t n_cells, int G, int value_dtype, int sorted);
int gene_entropy_dispatch(const void* values, int value_dtype, const int32_t* col_idx,
                          const int64_t* row_ptr, int64_t nnz, int64_t n_cells, int G, double* ent,
                          void* workspace, size_t ws_bytes, cudaStream_t st);
int csr_check_dispatch(const void* col_idx, int col16, const int64_t* row_ptr, int64_t n_cells, int G,
                       int* scratch_dev, int* result_host, cudaStream_t st);
int shard_transpose_dispatch(const void* values, int value_dtype, const int32_t* col_idx, const int64_t* row_ptr,
                             int64_t n_cells_local, int G, int64_t blk_begin, int64_t blk_end,
                             void* const* raw_dests, int n_dests, int own_index, int64_t raw_pitch,
                             uint32_t* ghist_local, unsigned long long* flags_dev,
                             unsigned long long* flags_host, int phase, cudaStream_t st);
int shard_bins_dispatch(const DiscretizeArgs& a, const uint8_t* raw, const uint32_t* const* hists, int n_hists,
                        uint32_t* ghist_total, double* qtab_dev, uint8_t* lut_all, cudaStream_t st);
int expand_onehot_dispatch(const void* bins_packed, int64_t pitch_bytes, int G, int rows_per_gene,
                           void* onehot, int64_t kp, int n_blocks, int operand_format,
                           cudaStream_t st);

}  // namespace gv
