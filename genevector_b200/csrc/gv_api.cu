// gv_api.cu -- the extern "C" surface of libgvb200.so (include/gvb200.h): argument checks,
// error reporting, device gate, host-side tile schedule.  No algorithm lives here.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "gv_internal.h"

namespace gv {

static thread_local std::string g_last_error;

int set_error(int code, const char* msg) {
  g_last_error = msg ? msg : "";
  return code;
}
int set_cuda_error(cudaError_t e, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d", int(e), cudaGetErrorString(e), file, line);
  g_last_error = buf;
  return GV_ERR_CUDA;
}

int device_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  return sms > 0 ? sms : 1;
}

static int check_device() {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10 || minor != 0) {
    char buf[160];
    snprintf(buf, sizeof(buf),
             "libgvb200 is built for sm_100a (B200) only; current device is compute capability %d.%d",
             major, minor);
    return set_error(GV_ERR_DEVICE, buf);
  }
  return GV_OK;
}

// implemented in gv_tiles.cu
int mi_tiles_dispatch(const void*, int64_t, int64_t, int, const int32_t*, const int8_t*, int64_t, int,
                      const int32_t*, int64_t, float*, int64_t, int, int, void*, cudaStream_t);
int gram_dump_dispatch(const void*, int64_t, int64_t, const int32_t*, int64_t, int32_t*, int64_t, int,
                       int, int, cudaStream_t);
int moment_tiles_dispatch(int, const void*, int64_t, int64_t, const int64_t*, const int64_t*,
                          const int32_t*, const double*, int64_t, int, int, int, const int32_t*, int64_t, int8_t*,
                          int64_t, float*, int64_t, int, void*, int, cudaStream_t);

int mma_peak_dispatch(int, int, int, int, double*, cudaStream_t);
int xcorr_tiles_dispatch(const void*, int64_t, int, const void*, int64_t, int, int64_t, int, const int64_t*,
                         const int64_t*, const double*, int64_t, const int64_t*, const int64_t*,
                         const double*, int64_t, int, const int32_t*, int64_t, float*, int64_t, int, void*,
                         cudaStream_t);
int graph_aggregate_dispatch(const void*, int64_t, int, int, const double*, const double*, int64_t, const int64_t*,
                             const int32_t*, const double*, double*, int, int64_t, int, void*, int64_t, int,
                             int64_t*, int64_t*, double*, cudaStream_t);
int symmetrize_dispatch(const float*, int64_t, int, float*, int64_t, cudaStream_t);
int build_pairs_dispatch(const float*, int64_t, int, float, int, int, int64_t*, int64_t*, float*,
                         cudaStream_t);
int compute_mi_pairs_impl(const int32_t*, int64_t, int, const int32_t*, const float*, float*, int64_t*, int,
                          cudaStream_t);

}  // namespace gv

using namespace gv;

extern "C" {

int gv_abi_version(void) { return GV_ABI_VERSION; }
const char* gv_last_error(void) { return g_last_error.c_str(); }
int gv_device_check(void) { return check_device(); }

size_t gv_discretize_workspace_bytes(int64_t nnz, int64_t n_cells, int n_genes, int value_dtype, int sorted) {
  return discretize_workspace_bytes(nnz < 0 ? 0 : nnz, n_cells, n_genes < 0 ? 0 : n_genes, value_dtype, sorted);
}

int gv_discretize_csr(const void* values, int value_dtype, const int32_t* col_idx,
                      const int64_t* row_ptr, int64_t n_cells, int n_genes, int64_t nnz, int n_bins,
                      const double* q_table_host, void* bins_packed, int64_t bins_pitch_bytes,
                      int32_t* n_bins_per_gene, int32_t* marg, double* edges, int64_t* gsum,
                      int64_t* gsq, void* val_rows, int64_t val_pitch, int64_t val_rows_alloc,
                      int val_mode, int val_limbs, int val_limb_bits, uint64_t* data_flags_out_host,
                      void* workspace, size_t workspace_bytes, int path, int* path_used_host,
                      void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (n_cells < 0 || n_genes <= 0 || nnz < 0) return set_error(GV_ERR_ARG, "bad matrix shape");
  if (n_cells >= (int64_t(1) << 31)) return set_error(GV_ERR_ARG, "n_cells must be < 2^31");
  if (!row_ptr || !q_table_host || !bins_packed || !n_bins_per_gene || !marg || !edges || !gsum ||
      !gsq || !workspace)
    return set_error(GV_ERR_ARG, "null pointer argument");
  if (nnz > 0 && (!values || !col_idx)) return set_error(GV_ERR_ARG, "null CSR arrays");
  if (val_rows && (val_pitch < n_cells || val_pitch % 128 != 0))
    return set_error(GV_ERR_ARG, "val_pitch must be a multiple of 128 covering n_cells");
  if (val_rows && (val_mode < 0 || val_mode > 3)) return set_error(GV_ERR_ARG, "unknown val_mode");
  if (val_rows && val_mode != 1 && (val_limbs < 1 || val_limbs > 4))
    return set_error(GV_ERR_ARG, "val_limbs must be in [1,4]");
  if (val_rows && val_rows_alloc < gv_value_rows(n_genes, val_mode == 1 ? 1 : val_limbs))
    return set_error(GV_ERR_ARG, "val_rows_alloc < gv_value_rows(n_genes, val_limbs)");
  DiscretizeArgs a;
  a.values = values; a.value_dtype = value_dtype; a.col_idx = col_idx; a.row_ptr = row_ptr;
  a.n_cells = n_cells; a.n_genes = n_genes; a.nnz = nnz; a.n_bins = n_bins; a.q_table = q_table_host;
  a.bins_packed = bins_packed; a.bins_pitch_bytes = bins_pitch_bytes;
  a.n_bins_per_gene = n_bins_per_gene; a.marg = marg; a.edges = edges; a.gsum = gsum; a.gsq = gsq;
  a.val_rows = val_rows; a.val_pitch = val_pitch; a.val_rows_alloc = val_rows_alloc;
  a.val_mode = val_mode; a.val_limbs = val_limbs; a.val_limb_bits = val_limb_bits;
  a.data_flags_out = data_flags_out_host; a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  if (path != GV_PATH_AUTO && path != GV_PATH_COUNTS && path != GV_PATH_GENERAL)
    return set_error(GV_ERR_ARG, "unknown discretisation path");
  a.path = path; a.path_used = path_used_host;
  return discretize_dispatch(a, static_cast<cudaStream_t>(stream));
}

int gv_csr_check(const void* col_idx, int value_dtype, const int64_t* row_ptr, int64_t n_cells, int n_genes,
                 int* scratch_dev, int* result_host, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!row_ptr || !scratch_dev || !result_host || n_cells < 0 || n_genes <= 0)
    return set_error(GV_ERR_ARG, "gv_csr_check: bad argument");
  *result_host = 0;
  return csr_check_dispatch(col_idx, value_dtype == GV_COUNTS_U8_COL16 ? 1 : 0, row_ptr, n_cells, n_genes,
                            scratch_dev, result_host, static_cast<cudaStream_t>(stream));
}

int gv_shard_transpose(const void* values, int value_dtype, const int32_t* col_idx, const int64_t* row_ptr_local,
                       int64_t n_cells_local, int n_genes, int64_t blk_begin, int64_t blk_end,
                       void* const* raw_dests_host, int n_dests, int own_index, int64_t raw_pitch,
                       uint32_t* ghist_local, void* flags_dev, uint64_t* data_flags_out_host, int phase,
                       void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!row_ptr_local || !raw_dests_host || !ghist_local || !flags_dev || !data_flags_out_host || n_genes <= 0 ||
      n_cells_local < 0)
    return set_error(GV_ERR_ARG, "gv_shard_transpose: bad argument");
  for (int p = 0; p < n_dests && p < 8; ++p)
    if (!raw_dests_host[p]) return set_error(GV_ERR_ARG, "gv_shard_transpose: null destination");
  if (n_cells_local > (blk_end - blk_begin) * 128)
    return set_error(GV_ERR_ARG, "gv_shard_transpose: more local cells than the block range holds");
  if (phase < 0 || phase > 3) return set_error(GV_ERR_ARG, "gv_shard_transpose: phase is a 2-bit mask");
  unsigned long long hf[2] = {0, 0};
  rc = shard_transpose_dispatch(values, value_dtype, col_idx, row_ptr_local, n_cells_local, n_genes, blk_begin,
                                blk_end, raw_dests_host, n_dests, own_index, raw_pitch, ghist_local,
                                static_cast<unsigned long long*>(flags_dev), hf, phase,
                                static_cast<cudaStream_t>(stream));
  if (phase & 2) {
    data_flags_out_host[0] = hf[0];
    data_flags_out_host[1] = hf[1];
  }
  return rc;
}

int gv_shard_bins(int value_dtype, int64_t n_cells, int n_genes, int n_bins, const double* q_table_host,
                  const void* raw, const void* const* hists_host, int n_hists, uint32_t* ghist_total,
                  double* qtab_dev, void* lut_ws, void* bins_packed, int64_t bins_pitch_bytes, int32_t* n_bins_per_gene,
                  int32_t* marg, double* edges, int64_t* gsum, int64_t* gsq, int value_rows,
                  int64_t val_rows_alloc, int val_limb_bits, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!q_table_host || !raw || !hists_host || !ghist_total || !qtab_dev || !lut_ws || !bins_packed || !n_bins_per_gene ||
      !marg || !edges || !gsum || !gsq || n_genes <= 0 || n_cells < 0)
    return set_error(GV_ERR_ARG, "gv_shard_bins: bad argument");
  if (bins_pitch_bytes % 64 != 0 || bins_pitch_bytes * 2 < n_cells)
    return set_error(GV_ERR_ARG, "gv_shard_bins: bins pitch must be a multiple of 64 bytes covering n_cells/2");
  DiscretizeArgs a;
  memset(&a, 0, sizeof(a));
  a.value_dtype = value_dtype; a.n_cells = n_cells; a.n_genes = n_genes; a.n_bins = n_bins;
  a.q_table = q_table_host; a.bins_packed = bins_packed; a.bins_pitch_bytes = bins_pitch_bytes;
  a.n_bins_per_gene = n_bins_per_gene; a.marg = marg; a.edges = edges; a.gsum = gsum; a.gsq = gsq;
  a.val_rows = value_rows ? const_cast<void*>(raw) : nullptr;
  a.val_pitch = bins_pitch_bytes * 2; a.val_rows_alloc = val_rows_alloc; a.val_mode = 0; a.val_limbs = 1;
  a.val_limb_bits = val_limb_bits;
  if (value_rows && val_rows_alloc < gv_value_rows(n_genes, 1))
    return set_error(GV_ERR_ARG, "gv_shard_bins: val_rows_alloc < gv_value_rows(n_genes, 1)");
  return shard_bins_dispatch(a, static_cast<const uint8_t*>(raw), reinterpret_cast<const uint32_t* const*>(hists_host),
                             n_hists, ghist_total, qtab_dev, static_cast<uint8_t*>(lut_ws),
                             static_cast<cudaStream_t>(stream));
}

int gv_gene_entropy(const void* values, int value_dtype, const int32_t* col_idx,
                    const int64_t* row_ptr, int64_t nnz, int64_t n_cells, int n_genes,
                    double* entropy_out, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (n_genes <= 0 || n_cells < 0 || nnz < 0 || !entropy_out || !workspace || (nnz > 0 && (!values || !col_idx)))
    return set_error(GV_ERR_ARG, "gv_gene_entropy: bad argument");
  return gene_entropy_dispatch(values, value_dtype, col_idx, row_ptr, nnz, n_cells, n_genes, entropy_out,
                               workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int gv_rows_per_gene(int n_bins) {
  if (n_bins < 1 || n_bins > GV_MAX_BINS) return set_error(GV_ERR_ARG, "n_bins must be in [1,31]");
  if (n_bins <= 5) return 5;
  if (n_bins <= 8) return 8;
  if (n_bins <= 10) return 10;
  if (n_bins <= 15) return 16;
  return 32;
}

int64_t gv_onehot_rows(int n_genes, int rows_per_gene) {
  if (rows_per_gene != 5 && rows_per_gene != 8 && rows_per_gene != 10 && rows_per_gene != 16 &&
      rows_per_gene != 32)
    return set_error(GV_ERR_ARG, "rows_per_gene must be 5, 8, 10, 16 or 32");
  const int gpb = 32 / rows_per_gene;
  return int64_t((n_genes + gpb - 1) / gpb) * 32;
}

int64_t gv_value_rows(int n_genes, int n_limbs) {
  if (n_limbs < 1 || n_limbs > 4 || n_genes < 0) return set_error(GV_ERR_ARG, "n_limbs must be in [1,4]");
  const int gpb = 32 / n_limbs;
  return int64_t((n_genes + gpb - 1) / gpb) * 32;
}

int gv_expand_onehot(const void* bins_packed, int64_t bins_pitch_bytes, int n_genes,
                     int rows_per_gene, void* onehot, int64_t kp, int operand_format, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!bins_packed || !onehot || n_genes <= 0) return set_error(GV_ERR_ARG, "bad argument");
  const int64_t rows = gv_onehot_rows(n_genes, rows_per_gene);
  if (rows < 0) return int(rows);
  return expand_onehot_dispatch(bins_packed, bins_pitch_bytes, n_genes, rows_per_gene, onehot, kp,
                                int(rows / 32), operand_format, static_cast<cudaStream_t>(stream));
}

int gv_tile_m(int cta_group) {
  if (cta_group == 1) return 128;
  if (cta_group == 2) return 256;
  return set_error(GV_ERR_ARG, "cta_group must be 1 or 2");
}

int64_t gv_tile_schedule(int64_t op_rows, int tile_m, int band, int32_t* out_pairs_host,
                         int64_t capacity) {
  if (op_rows <= 0 || (tile_m != 128 && tile_m != 256) || band < 1)
    return set_error(GV_ERR_ARG, "bad tile schedule argument");
  const int64_t tr = (op_rows + tile_m - 1) / tile_m;
  const int64_t tc = (op_rows + 255) / 256;
  int64_t n = 0;
  for (int64_t b0 = 0; b0 < tr; b0 += band) {
    const int64_t b1 = b0 + band < tr ? b0 + band : tr;
    for (int64_t c = 0; c < tc; ++c) {
      for (int64_t r = b0; r < b1; ++r) {
        // the tile holds some pair (row <= col) iff its first row is not past its last column
        if (r * tile_m > c * 256 + 255) continue;
        if (out_pairs_host && n < capacity) {
          out_pairs_host[2 * n] = int32_t(r * tile_m);
          out_pairs_host[2 * n + 1] = int32_t(c * 256);
        }
        ++n;
      }
    }
  }
  return n;
}

int gv_mi_tiles(const void* onehot, int64_t op_rows, int64_t kp, int rows_per_gene,
                const int32_t* marg, const int8_t* sgn, int64_t ld_sgn, int n_genes,
                const int32_t* tiles, int64_t n_tiles, float* out, int64_t ld_out, int cta_group,
                int operand_format, void* sync_ws, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!onehot || !marg || !tiles || !out || n_genes <= 0 || ld_out < n_genes)
    return set_error(GV_ERR_ARG, "gv_mi_tiles: bad argument");
  if (sgn && ld_sgn < n_genes) return set_error(GV_ERR_ARG, "gv_mi_tiles: ld_sgn < n_genes");
  if (op_rows < gv_onehot_rows(n_genes, rows_per_gene))
    return set_error(GV_ERR_ARG, "gv_mi_tiles: operand has too few rows");
  return mi_tiles_dispatch(onehot, op_rows, kp, rows_per_gene, marg, sgn, ld_sgn, n_genes, tiles,
                           n_tiles, out, ld_out, cta_group, operand_format, sync_ws,
                           static_cast<cudaStream_t>(stream));
}

int gv_gram_dump(const void* operand, int64_t op_rows, int64_t kp, const int32_t* tiles,
                 int64_t n_tiles, int32_t* out, int64_t ld_out, int cta_group, int active_rows,
                 int operand_format, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!operand || !tiles || (out && ld_out < op_rows)) return set_error(GV_ERR_ARG, "gv_gram_dump: bad argument");
  return gram_dump_dispatch(operand, op_rows, kp, tiles, n_tiles, out, ld_out, cta_group, active_rows,
                            operand_format, static_cast<cudaStream_t>(stream));
}

int gv_moment_tiles(int kind, const void* val_rows, int64_t op_rows, int64_t kp,
                    const int64_t* gsum, const int64_t* gsq, const int32_t* marg, const double* edges,
                    int64_t n_cells,
                    int n_genes, int n_limbs, int limb_bits, const int32_t* tiles, int64_t n_tiles,
                    int8_t* sgn, int64_t ld_sgn, float* out, int64_t ld_out, int cta_group,
                    void* sync_ws, int sign_zero, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (sign_zero != 0 && sign_zero != 1) return set_error(GV_ERR_ARG, "gv_moment_tiles: sign_zero must be 0 or 1");
  if (!val_rows || !tiles || n_genes <= 0 || n_limbs < 1 || n_limbs > 4 ||
      op_rows < gv_value_rows(n_genes, n_limbs))
    return set_error(GV_ERR_ARG, "gv_moment_tiles: bad argument");
  if ((sgn && ld_sgn < n_genes) || (out && ld_out < n_genes))
    return set_error(GV_ERR_ARG, "gv_moment_tiles: leading dimension < n_genes");
  return moment_tiles_dispatch(kind, val_rows, op_rows, kp, gsum, gsq, marg, edges, n_cells, n_genes, n_limbs,
                               limb_bits, tiles, n_tiles, sgn, ld_sgn, out, ld_out, cta_group, sync_ws,
                               sign_zero, static_cast<cudaStream_t>(stream));
}

int64_t gv_tile_schedule_rect(int64_t rows_a, int64_t rows_b, int tile_m, int32_t* out_pairs_host,
                              int64_t capacity) {
  if (rows_a <= 0 || rows_b <= 0 || (tile_m != 128 && tile_m != 256))
    return set_error(GV_ERR_ARG, "bad tile schedule argument");
  const int64_t tr = (rows_a + tile_m - 1) / tile_m, tc = (rows_b + 255) / 256;
  int64_t n = 0;
  for (int64_t b0 = 0; b0 < tr; b0 += 8)
    for (int64_t c = 0; c < tc; ++c)
      for (int64_t r = b0; r < (b0 + 8 < tr ? b0 + 8 : tr); ++r, ++n)
        if (out_pairs_host && n < capacity) {
          out_pairs_host[2 * n] = int32_t(r * tile_m);
          out_pairs_host[2 * n + 1] = int32_t(c * 256);
        }
  return n;
}

int gv_graph_aggregate(const void* x_rows, int64_t op_rows_x, int64_t kp, int x_limbs, int limb_bits,
                       const double* x_scale, const double* x_offset, int64_t x_scale_stride,
                       const int64_t* w_row_ptr,
                       const int32_t* w_col, const double* w_val, double* w_inv_ws, int include_self,
                       int64_t n_cells, int n_genes, void* y_rows, int64_t op_rows_y, int y_limbs,
                       int64_t* ysum, int64_t* ysq, double* yscale, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!x_rows || !w_row_ptr || !w_inv_ws || !y_rows || !ysum || !ysq || !yscale || n_genes <= 0 ||
      n_cells < 0 || kp < n_cells || x_limbs < 1 || x_limbs > 4 || y_limbs < 1 || y_limbs > 4 ||
      limb_bits < 1 || limb_bits > 7)
    return set_error(GV_ERR_ARG, "gv_graph_aggregate: bad argument");
  if (op_rows_x < gv_value_rows(n_genes, x_limbs) || op_rows_y < gv_value_rows(n_genes, y_limbs))
    return set_error(GV_ERR_ARG, "gv_graph_aggregate: value rows too short");
  return graph_aggregate_dispatch(x_rows, kp, x_limbs, limb_bits, x_scale, x_offset, x_scale_stride, w_row_ptr, w_col,
                                  w_val, w_inv_ws, include_self, n_cells, n_genes, y_rows, op_rows_y,
                                  y_limbs, ysum, ysq, yscale, static_cast<cudaStream_t>(stream));
}

int gv_xcorr_tiles(const void* x_rows, int64_t op_rows_x, int x_limbs, const void* y_rows,
                   int64_t op_rows_y, int y_limbs, int64_t kp, int limb_bits, const int64_t* xsum,
                   const int64_t* xsq, const double* x_scale, int64_t x_scale_stride,
                   const int64_t* ysum, const int64_t* ysq, const double* yscale, int64_t n_cells,
                   int n_genes, const int32_t* tiles, int64_t n_tiles, float* out, int64_t ld_out,
                   int cta_group, void* sync_ws, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!x_rows || !y_rows || !xsum || !xsq || !ysum || !ysq || !yscale || !tiles || !out || n_genes <= 0 ||
      ld_out < n_genes || x_limbs < 1 || x_limbs > 4 || y_limbs < 1 || y_limbs > 4)
    return set_error(GV_ERR_ARG, "gv_xcorr_tiles: bad argument");
  if (op_rows_x < gv_value_rows(n_genes, x_limbs) || op_rows_y < gv_value_rows(n_genes, y_limbs))
    return set_error(GV_ERR_ARG, "gv_xcorr_tiles: value rows too short");
  return xcorr_tiles_dispatch(x_rows, op_rows_x, x_limbs, y_rows, op_rows_y, y_limbs, kp, limb_bits, xsum, xsq,
                              x_scale, x_scale_stride, ysum, ysq, yscale, n_cells, n_genes, tiles, n_tiles,
                              out, ld_out, cta_group, sync_ws, static_cast<cudaStream_t>(stream));
}

int gv_symmetrize(const float* t, int64_t ld_t, int n_genes, float* out, int64_t ld_out, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!t || !out || n_genes <= 0 || ld_t < n_genes || ld_out < n_genes)
    return set_error(GV_ERR_ARG, "gv_symmetrize: bad argument");
  return symmetrize_dispatch(t, ld_t, n_genes, out, ld_out, static_cast<cudaStream_t>(stream));
}

int gv_mma_peak(int operand_format, int cta_group, int iters, int mode, double* tops_out_host, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  return mma_peak_dispatch(operand_format, cta_group, iters, mode, tops_out_host, static_cast<cudaStream_t>(stream));
}

int gv_compute_mi_pairs(const int32_t* x_disc_host, int64_t n_cells, int n_genes,
                        const int32_t* n_bins_per_gene_host, const float* corr_signs_host,
                        float* scores_out_host, int64_t* n_pairs_out_host, int operand_format, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!x_disc_host || !n_bins_per_gene_host || !scores_out_host || n_genes <= 0 || n_cells < 0)
    return set_error(GV_ERR_ARG, "gv_compute_mi_pairs: bad argument");
  if (n_cells >= (int64_t(1) << 31)) return set_error(GV_ERR_ARG, "n_cells must be < 2^31");
  return compute_mi_pairs_impl(x_disc_host, n_cells, n_genes, n_bins_per_gene_host, corr_signs_host,
                               scores_out_host, n_pairs_out_host, operand_format,
                               static_cast<cudaStream_t>(stream));
}

int gv_build_training_tensors(const float* scores, int64_t ld_scores, int n_genes, float c,
                              int clamp_negative, int round5, int64_t* i_idx, int64_t* j_idx,
                              float* xij, void* stream) {
  int rc = check_device();
  if (rc != GV_OK) return rc;
  if (!scores || !i_idx || !j_idx || !xij || n_genes < 1 || ld_scores < n_genes)
    return set_error(GV_ERR_ARG, "gv_build_training_tensors: bad argument");
  return build_pairs_dispatch(scores, ld_scores, n_genes, c * c, clamp_negative, round5, i_idx, j_idx,
                              xij, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
