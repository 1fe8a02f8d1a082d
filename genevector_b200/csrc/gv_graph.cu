// gv_graph.cu -- graph-neighbour aggregation for the graph_xcorr target.
//
// Replaces, for target_graph_xcorr (genevector/_graph_targets.py:10-55), the dense steps
//   X_agg = aggr_mean(X_dense, graph)        genevector/_aggregation.py:48-109
//           = (D^-1 W) @ X_dense  (rows of W normalised to sum 1, zero-degree rows left 0; with
//             include_self: 0.5 * X + 0.5 * that)
// on the gene-major value rows the co-moment contraction reads.  One CTA per gene: the gene's row
// (a few hundred KB, L1/L2-resident) is gathered through the graph's CSR, y_c = Sum_k wn_ck x_{col k}
// in float64 in CSR order like scipy's csr @ dense, and y is written back as fixed-point digit rows
// (the same representation gv_discretize_csr uses for non-count data), with Sum q, Sum q^2 and the
// per-gene scale the xcorr epilogue needs.  Off the headline path; HBM/L2-bound gather.
#include <cuda_runtime.h>
#include <cstdint>
#include "gv_internal.h"

namespace gv {

#define GV_LAUNCH_CHECK_G()                                   \
  do {                                                        \
    cudaError_t e_ = cudaGetLastError();                      \
    if (e_ != cudaSuccess) return set_cuda_error(e_, __FILE__, __LINE__); \
  } while (0)

// 1 / row sum of the adjacency (rows summing to 0 keep weight 1: _aggregation.py:61-62)
__global__ void k_graph_row_inv(const int64_t* __restrict__ w_ptr, const double* __restrict__ w_val,
                                int64_t n_cells, double* __restrict__ w_inv) {
  for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n_cells;
       c += int64_t(gridDim.x) * blockDim.x) {
    double s = 0.0;
    for (int64_t k = w_ptr[c]; k < w_ptr[c + 1]; ++k) s += w_val[k];
    if (s == 0.0) s = 1.0;
    w_inv[c] = 1.0 / s;
  }
}

__device__ __forceinline__ int64_t vrow0(int g, int n_limbs) {
  const int gpb = 32 / n_limbs;
  const int blk = g / gpb;
  return int64_t(blk) * 32 + (g - blk * gpb) * n_limbs;
}

constexpr int GA_THREADS = 256;

__global__ void __launch_bounds__(GA_THREADS)
k_graph_aggregate(const int8_t* __restrict__ xrows, int64_t pitch, int lx, int limb_bits,
                  const double* __restrict__ xscale, const double* __restrict__ xoffset,
                  int64_t xscale_stride, const int64_t* __restrict__ w_ptr, const int32_t* __restrict__ w_col,
                  const double* __restrict__ w_val, const double* __restrict__ w_inv, int include_self,
                  int64_t n_cells, int G, int8_t* __restrict__ yrows, int ly,
                  int64_t* __restrict__ ysum, int64_t* __restrict__ ysq, double* __restrict__ yscale) {
  __shared__ double red_d[GA_THREADS / 32][2];
  __shared__ double s_off;
  __shared__ long long red_i[GA_THREADS / 32][2];
  __shared__ int s_shift;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned dmask = (1u << limb_bits) - 1u;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int8_t* xr = xrows + vrow0(g, lx) * pitch;
    // genes with negative values are stored shifted by q0 (GV_EDGE_OFFSET_SLOT; DBL_MAX = none)
    double q0x = 0.0;
    if (xoffset != nullptr) { const double o = xoffset[int64_t(g) * xscale_stride]; q0x = o < 1e300 ? o : 0.0; }
    auto xval = [&](int64_t c) -> double {
      unsigned long long v = 0;
      for (int l = 0; l < lx; ++l) v |= (unsigned long long)(uint8_t(xr[int64_t(l) * pitch + c])) << (l * limb_bits);
      return double(v) - q0x;
    };
    auto yval = [&](int64_t c) -> double {
      double acc = 0.0;
      const double inv = w_inv[c];
      for (int64_t k = w_ptr[c]; k < w_ptr[c + 1]; ++k)
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(inv, w_val[k]), xval(w_col[k])));
      if (include_self) acc = __dadd_rn(__dmul_rn(0.5, xval(c)), __dmul_rn(0.5, acc));
      return acc;
    };
    // ---- sweep 1: range of the aggregated values of the gene
    double mx = 0.0, mn = 0.0;
    for (int64_t c = tid; c < n_cells; c += GA_THREADS) {
      const double y = yval(c);
      mx = y > mx ? y : mx; mn = y < mn ? y : mn;
    }
    for (int d = 16; d >= 1; d >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, mx, d); mx = o > mx ? o : mx;
      const double u = __shfl_xor_sync(0xffffffffu, mn, d); mn = u < mn ? u : mn;
    }
    __syncthreads();
    if (lane == 0) { red_d[warp][0] = mx; red_d[warp][1] = mn; }
    __syncthreads();
    if (tid == 0) {
      double m = 0.0, lo = 0.0;
      for (int w = 0; w < GA_THREADS / 32; ++w) {
        m = red_d[w][0] > m ? red_d[w][0] : m;
        lo = red_d[w][1] < lo ? red_d[w][1] : lo;
      }
      int ex = 0;
      if (m - lo > 0.0) frexp(m - lo, &ex);
      s_shift = (m - lo) > 0.0 ? ly * limb_bits - ex : 0;
      s_off = lo;                                   // <= 0; the rows hold rint((y - lo) * 2^shift)
    }
    __syncthreads();
    const int shift = s_shift;
    const double yoff = s_off;
    const unsigned long long cap = (1ull << (ly * limb_bits)) - 1ull;
    // ---- sweep 2: fixed-point digits, sums
    long long s1 = 0, s2 = 0;
    int8_t* yr = yrows + vrow0(g, ly) * pitch;
    for (int64_t c = tid; c < n_cells; c += GA_THREADS) {
      const double y = yval(c);
      const double ys = y - yoff;
      unsigned long long q = ys > 0.0 ? (unsigned long long)llrint(ldexp(ys, shift)) : 0ull;
      q = q > cap ? cap : q;
      s1 += (long long)q; s2 += (long long)(q * q);
      for (int l = 0; l < ly; ++l) yr[int64_t(l) * pitch + c] = int8_t((q >> (l * limb_bits)) & dmask);
    }
    for (int d = 16; d >= 1; d >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) { red_i[warp][0] = s1; red_i[warp][1] = s2; }
    __syncthreads();
    if (tid == 0) {
      long long a = 0, b = 0;
      for (int w = 0; w < GA_THREADS / 32; ++w) { a += red_i[w][0]; b += red_i[w][1]; }
      ysum[g] = a; ysq[g] = b;
      const double xs = xscale ? xscale[int64_t(g) * xscale_stride] : 1.0;
      yscale[g] = ldexp(xs, -shift);
    }
  }
}

int graph_aggregate_dispatch(const void* x_rows, int64_t kp, int x_limbs, int limb_bits,
                             const double* x_scale, const double* x_offset, int64_t x_scale_stride,
                             const int64_t* w_ptr,
                             const int32_t* w_col, const double* w_val, double* w_inv_ws,
                             int include_self, int64_t n_cells, int G, void* y_rows, int64_t y_rows_alloc,
                             int y_limbs, int64_t* ysum, int64_t* ysq, double* yscale, cudaStream_t st) {
  const int sms = device_sm_count();
  cudaError_t e = cudaMemsetAsync(y_rows, 0, size_t(y_rows_alloc) * size_t(kp), st);
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  if (n_cells > 0) {
    k_graph_row_inv<<<sms * 4, 256, 0, st>>>(w_ptr, w_val, n_cells, w_inv_ws);
    GV_LAUNCH_CHECK_G();
  }
  k_graph_aggregate<<<G < sms * 8 ? G : sms * 8, GA_THREADS, 0, st>>>(
      static_cast<const int8_t*>(x_rows), kp, x_limbs, limb_bits, x_scale, x_offset, x_scale_stride, w_ptr, w_col,
      w_val, w_inv_ws, include_self, n_cells, G, static_cast<int8_t*>(y_rows), y_limbs, ysum, ysq, yscale);
  GV_LAUNCH_CHECK_G();
  return GV_OK;
}

// out[a][b] = (t[a][b] + t[b][a]) / 2, zero diagonal (_graph_targets.py:51-52); 32 x 32 tiles
// through shared memory so both reads are coalesced.
__global__ void k_symmetrize(const float* __restrict__ t, int64_t ld_t, int G, float* __restrict__ out,
                             int64_t ld_out) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int a = bx + r, b = by + tx;               // t[b-block rows][a-block cols] transposed in
    tile[r][tx] = (by + r < G && bx + tx < G) ? t[int64_t(by + r) * ld_t + bx + tx] : 0.0f;
    (void)a; (void)b;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int a = bx + r, b = by + tx;
    if (a < G && b < G) {
      const float v = t[int64_t(a) * ld_t + b];
      const float w = tile[tx][r];                   // t[b][a]
      out[int64_t(a) * ld_out + b] = a == b ? 0.0f : 0.5f * (v + w);
    }
  }
}

int symmetrize_dispatch(const float* t, int64_t ld_t, int G, float* out, int64_t ld_out, cudaStream_t st) {
  dim3 grid((G + 31) / 32, (G + 31) / 32), block(32, 8);
  k_symmetrize<<<grid, block, 0, st>>>(t, ld_t, G, out, ld_out);
  GV_LAUNCH_CHECK_G();
  return GV_OK;
}

}  // namespace gv
