// gv_pairs.cu -- the reference's native boundary, shape for shape:
//   genevector._rust.compute_mi_pairs(x_disc, n_bins_per_gene, corr_signs)   src/lib.rs:9-99,
//   stub genevector/_rust.pyi:1-5, called from compute_mi_rust (genevector/metrics.py:312-339)
// A maintainer who swaps only the Rust module binds gv_compute_mi_pairs: host arrays in (the
// pre-discretised int32 matrix, the per-gene bin counts, the optional float32 correlation matrix whose
// sign signs the score: +1 unless negative, lib.rs:91-94), a dense score matrix out.  The library
// allocates its own device memory here (the caller of this entry has no device allocator) and frees it
// before returning.  Pipeline: x_disc [cells][genes] int32 -> packed 4-bit bins [genes][cells] +
// marginals (k_pack_xdisc: shared-memory transposition) -> one-hot operand -> tcgen05 contraction with
// the fused MI epilogue (the same kernels gv_mi_tiles launches).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include "gv_internal.h"

namespace gv {

int mi_tiles_dispatch(const void*, int64_t, int64_t, int, const int32_t*, const int8_t*, int64_t, int,
                      const int32_t*, int64_t, float*, int64_t, int, int, void*, cudaStream_t);

#define GV_PAIRS_TRY(expr)                                                  \
  do {                                                                      \
    cudaError_t e_ = (expr);                                                \
    if (e_ != cudaSuccess) { rc = set_cuda_error(e_, __FILE__, __LINE__); goto done; } \
  } while (0)

// x_disc [n_cells][n_genes] int32 (C order, as discretize_genes returns it) -> packed bins
// [n_genes][pitch_bytes] (two cells per byte, low nibble first) and marg[g][0] = #cells with bin > 0,
// marg[g][a] = #cells in bin a.  One CTA: 32 genes x 256 cells through a padded shared-memory tile;
// reads are 128-byte rows of x_disc, writes 128-byte runs of a gene's packed row.
// flags[0] |= 1 when a bin index is negative, above 15 or not below the gene's n_bins_per_gene.
constexpr int PX_CELLS = 256;
__global__ void __launch_bounds__(256)
k_pack_xdisc(const int32_t* __restrict__ x_disc, int64_t n_cells, int G, const int32_t* __restrict__ nbpg,
             uint8_t* __restrict__ bins, int64_t pitch_bytes, int32_t* __restrict__ marg,
             unsigned int* __restrict__ flags) {
  __shared__ __align__(8) uint8_t tile[32][PX_CELLS + 8];
  __shared__ unsigned int cnt[32][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.y * 32;
  const int64_t c0 = int64_t(blockIdx.x) * PX_CELLS;
  for (int i = tid; i < 32 * 16; i += 256) (&cnt[0][0])[i] = 0u;
  const int g = g0 + lane;
  const int nb = g < G ? nbpg[g] : 0;
  unsigned bad = 0;
  // each warp reads 32 rows (cells), lane = gene
  for (int r = warp; r < PX_CELLS; r += 8) {
    const int64_t c = c0 + r;
    int v = 0;
    if (c < n_cells && g < G) {
      v = x_disc[c * G + g];
      if (v < 0 || v > 15 || v >= nb) { bad = 1; v = 0; }
    }
    tile[lane][r] = uint8_t(v);
  }
  __syncthreads();
  // each warp packs 4 genes: lane handles 8 cells -> 4 bytes
  for (int gl = warp * 4; gl < warp * 4 + 4; ++gl) {
    const uint2 w = *reinterpret_cast<const uint2*>(&tile[gl][lane * 8]);
    const uint32_t lo = (w.x & 0x0F) | ((w.x >> 4) & 0xF0) | ((w.x >> 8) & 0xF00) | ((w.x >> 12) & 0xF000);
    const uint32_t hi = (w.y & 0x0F) | ((w.y >> 4) & 0xF0) | ((w.y >> 8) & 0xF00) | ((w.y >> 12) & 0xF000);
    const uint32_t packed = lo | (hi << 16);
    if (g0 + gl < G && c0 + lane * 8 < pitch_bytes * 2)
      *reinterpret_cast<uint32_t*>(bins + int64_t(g0 + gl) * pitch_bytes + (c0 >> 1) + lane * 4) = packed;
    const uint32_t ws[2] = {w.x, w.y};
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const unsigned v = (ws[k] >> (8 * b)) & 0xFFu;
        if (v) atomicAdd(&cnt[gl][v], 1u);
      }
  }
  __syncthreads();
  for (int i = tid; i < 32 * 16; i += 256) {
    const int gl = i >> 4, a = i & 15;
    const unsigned n = cnt[gl][a];
    if (a >= 1 && n && g0 + gl < G) {
      atomicAdd(&marg[int64_t(g0 + gl) * MARG_STRIDE_ + a], int(n));
      atomicAdd(&marg[int64_t(g0 + gl) * MARG_STRIDE_], int(n));
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flags, 1u);
}

// corr_signs float32 [G][G] -> int8 [G][ld]: +1 unless negative (lib.rs:91-94; NaN compares false: -1)
__global__ void k_corr_to_sign(const float* __restrict__ corr, int G, int8_t* __restrict__ sgn, int64_t ld) {
  const int64_t total = int64_t(G) * G;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / G, c = i - r * G;
    sgn[r * ld + c] = corr[i] >= 0.0f ? int8_t(1) : int8_t(-1);
  }
}

int compute_mi_pairs_impl(const int32_t* x_disc_host, int64_t n_cells, int G, const int32_t* nbpg_host,
                          const float* corr_host, float* out_host, int64_t* n_pairs_out, int operand_format,
                          cudaStream_t st) {
  int rc = GV_OK;
  int32_t *d_x = nullptr, *d_nbpg = nullptr, *d_marg = nullptr, *d_tiles = nullptr;
  uint8_t* d_bins = nullptr;
  void* d_onehot = nullptr;
  float *d_corr = nullptr, *d_out = nullptr;
  int8_t* d_sgn = nullptr;
  unsigned int* d_flags = nullptr;
  void* d_sync = nullptr;
  // layout parameters from the largest bin index any gene can hold
  int max_nb = 1, expressed = 0;
  for (int g = 0; g < G; ++g) {
    if (nbpg_host[g] > max_nb) max_nb = nbpg_host[g];
    if (nbpg_host[g] > 1) ++expressed;
  }
  if (n_pairs_out) *n_pairs_out = int64_t(expressed) * (expressed - 1) / 2;      // lib.rs:30-38
  const int n_bins = max_nb - 1 < 1 ? 1 : max_nb - 1;
  if (n_bins > 15) return set_error(GV_ERR_UNSUPPORTED, "gv_compute_mi_pairs: more than 15 non-zero bins per gene");
  const int B = gv_rows_per_gene(n_bins);
  const int64_t kp = (n_cells + 255) / 256 * 256 > 0 ? (n_cells + 255) / 256 * 256 : 256;
  const int64_t pitch = kp / 2;
  const int64_t op_rows = gv_onehot_rows(G, B);
  const bool fp4 = operand_format == GV_OPERAND_FP4 ||
                   (operand_format < 0 && (B == 10 || B == 5) && n_cells < (int64_t(1) << 24));
  if (fp4 && !(B == 10 || B == 5)) return set_error(GV_ERR_ARG, "gv_compute_mi_pairs: fp4 needs 5 or 10 rows per gene");
  const int64_t row_bytes = fp4 ? kp / 2 : kp;
  const int64_t ld_sgn = (int64_t(G) + 15) / 16 * 16;
  const int64_t n_tiles = gv_tile_schedule(op_rows, 256, 8, nullptr, 0);
  std::vector<int32_t> tiles(size_t(n_tiles > 0 ? n_tiles : 1) * 2);
  gv_tile_schedule(op_rows, 256, 8, tiles.data(), n_tiles);

  GV_PAIRS_TRY(cudaMalloc(&d_x, size_t(n_cells > 0 ? n_cells : 1) * G * 4));
  GV_PAIRS_TRY(cudaMalloc(&d_nbpg, size_t(G) * 4));
  GV_PAIRS_TRY(cudaMalloc(&d_marg, size_t(G) * MARG_STRIDE_ * 4));
  GV_PAIRS_TRY(cudaMalloc(&d_bins, size_t(G) * pitch));
  GV_PAIRS_TRY(cudaMalloc(&d_flags, 64));
  GV_PAIRS_TRY(cudaMalloc(&d_sync, 64));
  GV_PAIRS_TRY(cudaMalloc(&d_tiles, tiles.size() * 4));
  GV_PAIRS_TRY(cudaMalloc(&d_out, size_t(G) * G * 4));
  GV_PAIRS_TRY(cudaMemsetAsync(d_marg, 0, size_t(G) * MARG_STRIDE_ * 4, st));
  GV_PAIRS_TRY(cudaMemsetAsync(d_bins, 0, size_t(G) * pitch, st));
  GV_PAIRS_TRY(cudaMemsetAsync(d_flags, 0, 64, st));
  GV_PAIRS_TRY(cudaMemsetAsync(d_out, 0, size_t(G) * G * 4, st));
  rc = gv_upload_pageable(d_x, x_disc_host, size_t(n_cells) * G * 4, 0, st);
  if (rc != GV_OK) goto done;
  GV_PAIRS_TRY(cudaMemcpyAsync(d_nbpg, nbpg_host, size_t(G) * 4, cudaMemcpyHostToDevice, st));
  GV_PAIRS_TRY(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice, st));
  if (n_cells > 0) {
    dim3 grid(unsigned((n_cells + PX_CELLS - 1) / PX_CELLS), unsigned((G + 31) / 32));
    k_pack_xdisc<<<grid, 256, 0, st>>>(d_x, n_cells, G, d_nbpg, d_bins, pitch, d_marg, d_flags);
    GV_PAIRS_TRY(cudaGetLastError());
  }
  {
    unsigned int hflags = 0;
    GV_PAIRS_TRY(cudaMemcpyAsync(&hflags, d_flags, 4, cudaMemcpyDeviceToHost, st));
    GV_PAIRS_TRY(cudaStreamSynchronize(st));
    cudaFree(d_x);
    d_x = nullptr;
    if (hflags & 1u) {
      rc = set_error(GV_ERR_ARG, "gv_compute_mi_pairs: x_disc holds a bin index outside [0, n_bins_per_gene)");
      goto done;
    }
  }
  if (corr_host != nullptr) {
    GV_PAIRS_TRY(cudaMalloc(&d_corr, size_t(G) * G * 4));
    GV_PAIRS_TRY(cudaMalloc(&d_sgn, size_t(G) * ld_sgn));
    rc = gv_upload_pageable(d_corr, corr_host, size_t(G) * G * 4, 0, st);
    if (rc != GV_OK) goto done;
    k_corr_to_sign<<<device_sm_count() * 8, 256, 0, st>>>(d_corr, G, d_sgn, ld_sgn);
    GV_PAIRS_TRY(cudaGetLastError());
  }
  GV_PAIRS_TRY(cudaMalloc(&d_onehot, size_t(op_rows) * row_bytes));
  rc = expand_onehot_dispatch(d_bins, pitch, G, B, d_onehot, kp, int(op_rows / 32),
                              fp4 ? GV_OPERAND_FP4 : GV_OPERAND_INT8, st);
  if (rc != GV_OK) goto done;
  rc = mi_tiles_dispatch(d_onehot, op_rows, row_bytes, B, d_marg, d_sgn, ld_sgn, G, d_tiles, n_tiles, d_out, G, 2,
                         fp4 ? GV_OPERAND_FP4 : GV_OPERAND_INT8,
                         size_t(op_rows) * row_bytes > (size_t(96) << 20) ? d_sync : nullptr, st);
  if (rc != GV_OK) goto done;
  GV_PAIRS_TRY(cudaMemcpyAsync(out_host, d_out, size_t(G) * G * 4, cudaMemcpyDeviceToHost, st));
  GV_PAIRS_TRY(cudaStreamSynchronize(st));
done:
  if (rc != GV_OK) cudaStreamSynchronize(st);
  cudaFree(d_x); cudaFree(d_nbpg); cudaFree(d_marg); cudaFree(d_bins); cudaFree(d_flags); cudaFree(d_sync);
  cudaFree(d_tiles); cudaFree(d_out); cudaFree(d_corr); cudaFree(d_sgn); cudaFree(d_onehot);
  return rc;
}

}  // namespace gv
