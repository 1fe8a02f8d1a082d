// gv_discretize.cu -- per-gene quantile discretisation of a cells x genes CSR matrix.
//
// Reproduces genevector/metrics.py:25-69 (discretize_genes) bit for bit:
//   bin 0 for x <= 0; k = min(n_bins, #unique positive values); k <= 1 -> every positive value
//   in bin 1; else k-1 edges = np.percentile(positive values, linspace(0,100,k+1)[1:-1])
//   (method 'linear': numpy/lib/_function_base_impl.py _quantile/_get_indexes/_lerp) and
//   bin = 1 + #{edges <= x}  (np.searchsorted(..., side='right') + 1).
// The arithmetic of the edges (float64, one subtraction in the input dtype, two-branch lerp,
// separately rounded multiply and add) is restated in SURVEY.md Appendix A.
//
// Pipeline (all HBM/L2-bound integer work, no tensor cores):
//   k_count_positive   CSR values -> per-gene count of x>0 (+ data flags for the co-moment path)
//   k_scan_counts      exclusive scan -> gene_ptr (gene-major segment offsets)
//   k_scatter          CSR (cell-major) -> gene-major (cell, value) segments, one warp per cell row
//   k_gene_bins        one CTA per gene: 11-bit histogram + distinct-value set, multi-target MSD
//                      radix select of the <= 2(k-1) order statistics, edges, then bin assignment:
//                      packed 4-bit bins [G][pitch], marginals [G][16], per-gene sums, optional
//                      INT8 value rows for the co-moment contraction
//   k_expand_onehot    packed bins -> K-major INT8 one-hot operand rows for the tcgen05 contraction
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>
#include <cstdlib>
#include <type_traits>
#include "gv_internal.h"

namespace gv {

// ---------------------------------------------------------------- value keys
template <typename T> struct ValKey;
template <> struct ValKey<float> {
  static constexpr int W = 31;   // significant key bits (sign bit of a positive float is 0)
  __device__ static uint64_t key(float v) { return uint64_t(__float_as_uint(v)); }
  __device__ static float val(uint64_t k) { return __uint_as_float(uint32_t(k)); }
  __device__ static float sub(float b, float a) { return __fsub_rn(b, a); }
};
template <> struct ValKey<double> {
  static constexpr int W = 63;
  __device__ static uint64_t key(double v) { return uint64_t(__double_as_longlong(v)); }
  __device__ static double val(uint64_t k) { return __longlong_as_double((long long)k); }
  __device__ static double sub(double b, double a) { return __dsub_rn(b, a); }
};

// ---------------------------------------------------------------- value rows
// Operand of the co-moment contraction (gv_moment_tiles): n_limbs INT8 rows per gene holding the
// base-2^limb_bits digits of a non-negative integer per cell, in 32-row blocks of 32/n_limbs genes
// (row = block*32 + slot*n_limbs + digit; see MomentEpi in gv_epilogues.cuh).  What the integer is:
//   val_mode 0  the count itself (exact co-moments)
//   val_mode 1  1 where the gene is detected (one row per gene; Jaccard)
//   val_mode 2  fixed point: rint(x * 2^e_g) with a per-gene power-of-two scale that puts the gene's
//               largest value just below 2^(n_limbs*limb_bits) (Pearson / cosine are invariant to it)
//   val_mode 3  doubled average rank minus (n_zero + 1): 0 for the gene's zero cells,
//               n_zero + #{values < x} + #{values <= x} for the others (Spearman = Pearson of ranks)
__device__ __forceinline__ int64_t val_row0(int g, int n_limbs) {
  const int gpb = 32 / n_limbs;
  const int blk = g / gpb;
  return int64_t(blk) * 32 + (g - blk * gpb) * n_limbs;
}
__device__ __forceinline__ void write_digits(int8_t* __restrict__ val_rows, int64_t val_pitch, int g,
                                             int64_t cell, unsigned long long xi, int n_limbs,
                                             int limb_bits) {
  int8_t* p = val_rows + val_row0(g, n_limbs) * val_pitch + cell;
  const unsigned long long mask = (1ull << limb_bits) - 1ull;
  for (int l = 0; l < n_limbs; ++l) p[int64_t(l) * val_pitch] = int8_t((xi >> (l * limb_bits)) & mask);
}

// ---------------------------------------------------------------- k_count_positive
// flags[0] |= 1 if a value is negative, |= 2 if a positive value is not an integer;
// flags[1] = max over ceil(values) (as uint64).
// ALL: count every stored entry (the signed rank / entropy regroup), not only x > 0.
template <typename T, bool ALL = false>
__global__ void k_count_positive(const T* __restrict__ values, const int32_t* __restrict__ col_idx,
                                 int64_t nnz, uint32_t* __restrict__ cnt,
                                 unsigned long long* __restrict__ flags) {
  unsigned f = 0;
  unsigned long long vmax = 0;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    const T v = values[e];
    if (ALL && !(v > T(0))) atomicAdd(&cnt[col_idx[e]], 1u);
    if (v > T(0)) {
      atomicAdd(&cnt[col_idx[e]], 1u);
      const T c = ceil(v);
      if (c != v) f |= 2u;
      const unsigned long long ci = c < T(1.8e19) ? (unsigned long long)c : ~0ull;
      vmax = ci > vmax ? ci : vmax;
    } else if (v < T(0)) {
      f |= 1u;
    }
    if (!(v - v == T(0))) f |= 8u;           // NaN or infinity
  }
  f = __reduce_or_sync(0xffffffffu, f);
  for (int d = 16; d >= 1; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, vmax, d);
    vmax = o > vmax ? o : vmax;
  }
  if ((threadIdx.x & 31) == 0) {
    if (f) atomicOr(&flags[0], (unsigned long long)f);
    if (vmax) atomicMax(&flags[1], vmax);
  }
}

// ---------------------------------------------------------------- k_scan_counts
// Single CTA exclusive scan of G counts into int64 offsets; also resets the cursors.
__global__ void k_scan_counts(const uint32_t* __restrict__ cnt, int64_t* __restrict__ gene_ptr,
                              unsigned long long* __restrict__ cursor, int G) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < G; base += blockDim.x) {
    const int g = base + threadIdx.x;
    const unsigned long long x = g < G ? cnt[g] : 0ull;
    unsigned long long inc = x;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0ull;
      unsigned long long winc = w;
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += y;
      }
      warp_tot[lane] = winc - w;   // exclusive
    }
    __syncthreads();
    const unsigned long long excl = carry + warp_tot[warp] + inc - x;
    if (g < G) { gene_ptr[g] = (int64_t)excl; cursor[g] = excl; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) gene_ptr[G] = (int64_t)carry;
}

// ---------------------------------------------------------------- k_scatter
// One warp per CSR row (cell): coalesced reads, scattered 4/8-byte writes into gene segments.
template <typename T, bool ALL = false>
__global__ void k_scatter(const T* __restrict__ values, const int32_t* __restrict__ col_idx,
                          const int64_t* __restrict__ row_ptr, int64_t n_cells,
                          unsigned long long* __restrict__ cursor, int32_t* __restrict__ g_cell,
                          T* __restrict__ g_val) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_cells; r += warps) {
    const int64_t e0 = row_ptr[r], e1 = row_ptr[r + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
      const T v = values[e];
      if (ALL || v > T(0)) {
        const unsigned long long pos = atomicAdd(&cursor[col_idx[e]], 1ull);
        g_cell[pos] = int32_t(r);
        g_val[pos] = v;
      }
    }
  }
}

// ---------------------------------------------------------------- k_gene_bins
constexpr int EDGE_SCALE_SLOT = GV_EDGE_SCALE_SLOT;
constexpr int GB_THREADS = 256;
constexpr int H0_BITS = 11;
constexpr int H0_SIZE = 1 << H0_BITS;
constexpr int SET_CAP = 64;          // distinct-value set capacity (n_bins <= 31)

// WIDE = n_bins 16..31: up to 30 edges (60 order statistics) and one byte per cell in the bin matrix;
// otherwise up to 14 edges and two cells per byte.  MAX_T >= 2 * edges.
template <int MAX_T>
struct GeneSmem {
  uint32_t h0[H0_SIZE];              // level-0 histogram, then its exclusive scan
  uint32_t hs[MAX_T][256];           // per-slot digit histograms of a refinement pass
  unsigned long long set[SET_CAP];   // distinct keys (0 = empty)
  unsigned long long slot_prefix[MAX_T];
  unsigned long long slot_min[MAX_T];
  unsigned long long slot_max[MAX_T];
  unsigned long long tgt_prefix[MAX_T];
  unsigned long long tgt_val[MAX_T];
  uint32_t tgt_rank[MAX_T];
  int tgt_slot[MAX_T];
  int tgt_done[MAX_T];
  double edges[EDGE_STRIDE_];
  uint32_t marg[MARG_STRIDE_];
  uint32_t scan_tmp[GB_THREADS];
  unsigned long long red[GB_THREADS / 32][2];
  unsigned long long kmax[GB_THREADS / 32];
  int n_distinct;
  int n_slots;
  int n_undone;
};

template <typename T, bool WIDE>
__global__ void __launch_bounds__(GB_THREADS)
k_gene_bins(const int64_t* __restrict__ gene_ptr, const int32_t* __restrict__ g_cell,
            const T* __restrict__ g_val, int G, int n_bins, const double* __restrict__ q_table,
            uint32_t* __restrict__ bins_packed, int64_t pitch_words,
            int32_t* __restrict__ n_bins_per_gene, int32_t* __restrict__ marg,
            double* __restrict__ edges_out, int64_t* __restrict__ gsum, int64_t* __restrict__ gsq,
            int8_t* __restrict__ val_rows, int64_t val_pitch, int val_mode, int limb_bits,
            int n_limbs, int64_t n_cells) {
  using K = ValKey<T>;
  constexpr int NE = WIDE ? 30 : 14;              // most edges a gene can have
  constexpr int MAX_T = WIDE ? 64 : 32;
  extern __shared__ __align__(16) uint8_t gb_smem[];
  GeneSmem<MAX_T>& sm = *reinterpret_cast<GeneSmem<MAX_T>*>(gb_smem);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int64_t e0 = gene_ptr[g], e1 = gene_ptr[g + 1];
    const int64_t m = e1 - e0;
    __syncthreads();   // previous gene fully done with smem
    if (m == 0) {
      if (tid == 0) {
        n_bins_per_gene[g] = 1;               // metrics.py:55-57
        gsum[g] = 0; gsq[g] = 0;
      }
      if (tid < MARG_STRIDE_) marg[int64_t(g) * MARG_STRIDE_ + tid] = 0;
      if (tid < EDGE_STRIDE_) edges_out[int64_t(g) * EDGE_STRIDE_ + tid] = tid == EDGE_SCALE_SLOT ? 1.0 : DBL_MAX;
      continue;
    }
    // ---- pass 0: level-0 histogram, distinct set, integer sums
    for (int i = tid; i < H0_SIZE; i += GB_THREADS) sm.h0[i] = 0;
    if (tid < SET_CAP) sm.set[tid] = 0;
    if (tid < MARG_STRIDE_) sm.marg[tid] = 0;
    if (tid == 0) sm.n_distinct = 0;
    __syncthreads();
    unsigned long long kmax = 0;
    for (int64_t e = e0 + tid; e < e1; e += GB_THREADS) {
      const T v = g_val[e];
      const uint64_t key = K::key(v);
      atomicAdd(&sm.h0[uint32_t(key >> (K::W - H0_BITS))], 1u);
      kmax = key > kmax ? key : kmax;
      if (*(volatile int*)&sm.n_distinct <= n_bins) {
        // open-addressing insert; stops mattering once more than n_bins values are known
        uint32_t h = uint32_t((key * 0x9E3779B97F4A7C15ull) >> 58);   // 6 bits
        for (int probe = 0; probe < SET_CAP; ++probe) {
          const unsigned long long cur = *(volatile unsigned long long*)&sm.set[h];
          if (cur == key) break;
          if (cur == 0ull) {
            const unsigned long long old = atomicCAS(&sm.set[h], 0ull, (unsigned long long)key);
            if (old == 0ull) { atomicAdd(&sm.n_distinct, 1); break; }
            if (old == key) break;
          }
          h = (h + 1) & (SET_CAP - 1);
        }
      }
    }
    // largest value of the gene (positive floats order like their bit patterns)
    for (int d = 16; d >= 1; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, kmax, d);
      kmax = o > kmax ? o : kmax;
    }
    if (lane == 0) sm.kmax[warp] = kmax;
    __syncthreads();
    const int n_dist = sm.n_distinct;
    const int k = n_dist < n_bins ? n_dist : n_bins;     // metrics.py:59
    const int n_edges = k > 1 ? k - 1 : 0;
    const int nt = 2 * n_edges;
    if (tid < EDGE_STRIDE_) sm.edges[tid] = DBL_MAX;

    if (n_edges > 0) {
      // ---- target ranks (numpy _get_indexes: floor(vi), floor(vi)+1, clamped at the top)
      if (tid < nt) {
        const int ei = tid >> 1;
        const double q = q_table[k * QT_DIM_ + ei];
        const double vi = __dmul_rn(double(m - 1), q);
        int64_t rank;
        if (vi >= double(m - 1)) rank = m - 1;
        else rank = int64_t(floor(vi)) + (tid & 1);
        sm.tgt_rank[tid] = uint32_t(rank);
        sm.tgt_done[tid] = 0;
      }
      // ---- exclusive scan of h0 (2048 = 256 threads x 8)
      uint32_t loc[8]; uint32_t tsum = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { loc[i] = sm.h0[tid * 8 + i]; tsum += loc[i]; }
      uint32_t inc = tsum;
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
      }
      if (lane == 31) sm.scan_tmp[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        const uint32_t w = lane < GB_THREADS / 32 ? sm.scan_tmp[lane] : 0u;
        uint32_t winc = w;
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, winc, d);
          if (lane >= d) winc += y;
        }
        sm.scan_tmp[lane] = winc - w;
      }
      __syncthreads();
      uint32_t run = sm.scan_tmp[warp] + inc - tsum;
#pragma unroll
      for (int i = 0; i < 8; ++i) { sm.h0[tid * 8 + i] = run; run += loc[i]; }
      __syncthreads();
      // ---- locate each target's level-0 bucket: largest b with excl[b] <= rank
      if (tid < nt) {
        const uint32_t rank = sm.tgt_rank[tid];
        int lo = 0, hi = H0_SIZE - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (sm.h0[mid] <= rank) lo = mid; else hi = mid - 1;
        }
        sm.tgt_prefix[tid] = (unsigned long long)lo;
        sm.tgt_rank[tid] = rank - sm.h0[lo];
      }
      __syncthreads();
      int shift = K::W - H0_BITS;   // key bits below the current prefixes
      while (true) {
        // ---- group undone targets into slots by prefix
        if (tid == 0) {
          int ns = 0, undone = 0;
          for (int t = 0; t < nt; ++t) {
            if (sm.tgt_done[t]) continue;
            ++undone;
            int s = -1;
            for (int u = 0; u < ns; ++u) if (sm.slot_prefix[u] == sm.tgt_prefix[t]) { s = u; break; }
            if (s < 0) { s = ns++; sm.slot_prefix[s] = sm.tgt_prefix[t]; }
            sm.tgt_slot[t] = s;
          }
          sm.n_slots = ns; sm.n_undone = undone;
        }
        __syncthreads();
        if (sm.n_undone == 0) break;
        const int ns = sm.n_slots;
        const int width = shift < 8 ? shift : 8;
        const int nshift = shift - width;
        for (int i = tid; i < ns * 256; i += GB_THREADS) sm.hs[i >> 8][i & 255] = 0;
        if (tid < ns) { sm.slot_min[tid] = ~0ull; sm.slot_max[tid] = 0ull; }
        __syncthreads();
        for (int64_t e = e0 + tid; e < e1; e += GB_THREADS) {
          const uint64_t key = K::key(g_val[e]);
          const uint64_t pre = key >> shift;
          for (int s = 0; s < ns; ++s) {
            if (sm.slot_prefix[s] == pre) {
              atomicAdd(&sm.hs[s][uint32_t(key >> nshift) & ((1u << width) - 1u)], 1u);
              atomicMin(&sm.slot_min[s], (unsigned long long)key);
              atomicMax(&sm.slot_max[s], (unsigned long long)key);
              break;
            }
          }
        }
        __syncthreads();
        if (tid < nt && !sm.tgt_done[tid]) {
          const int s = sm.tgt_slot[tid];
          if (sm.slot_min[s] == sm.slot_max[s]) {          // bucket holds one distinct value
            sm.tgt_val[tid] = sm.slot_min[s]; sm.tgt_done[tid] = 1;
          } else {
            uint32_t rank = sm.tgt_rank[tid], cum = 0; int d = 0;
            for (; d < 255; ++d) {
              const uint32_t c = sm.hs[s][d];
              if (rank < cum + c) break;
              cum += c;
            }
            sm.tgt_rank[tid] = rank - cum;
            const unsigned long long np = (sm.tgt_prefix[tid] << width) | (unsigned long long)d;
            sm.tgt_prefix[tid] = np;
            if (nshift == 0) { sm.tgt_val[tid] = np; sm.tgt_done[tid] = 1; }
          }
        }
        __syncthreads();
        shift = nshift;
      }
      // ---- edges (numpy _lerp)
      if (tid < n_edges) {
        const double q = q_table[k * QT_DIM_ + tid];
        const double vi = __dmul_rn(double(m - 1), q);
        const T a = K::val(sm.tgt_val[2 * tid]);
        const T b = K::val(sm.tgt_val[2 * tid + 1]);
        double gam;
        if (vi >= double(m - 1)) gam = __dadd_rn(vi, 1.0);   // previous index is -1 there
        else gam = __dsub_rn(vi, floor(vi));
        const T d = K::sub(b, a);                            // subtraction in the input dtype
        double e;
        if (gam >= 0.5) e = __dsub_rn(double(b), __dmul_rn(double(d), __dsub_rn(1.0, gam)));
        else e = __dadd_rn(double(a), __dmul_rn(double(d), gam));
        sm.edges[tid] = e;
      }
    }
    __syncthreads();
    // ---- bin assignment: bin = 1 + #{edges <= x}; packed nibbles, marginals, value rows
    int fx_shift = 0;                       // val_mode 2: x -> rint(x * 2^fx_shift)
    if (val_mode == 2) {
      unsigned long long km = 0;
      for (int w = 0; w < GB_THREADS / 32; ++w) km = sm.kmax[w] > km ? sm.kmax[w] : km;
      int ex;
      frexp(double(K::val(km)), &ex);       // max = f * 2^ex, f in [0.5, 1)
      fx_shift = n_limbs * limb_bits - ex;
    }
    const unsigned long long xi_cap = (1ull << (n_limbs * limb_bits)) - 1ull;
    long long s1 = 0, s2 = 0;
    uint32_t* brow = bins_packed + int64_t(g) * pitch_words;
    for (int64_t e = e0 + tid; e < e1; e += GB_THREADS) {
      const T v = g_val[e];
      const int cell = g_cell[e];
      const double vd = double(v);
      int bin = 1;
#pragma unroll
      for (int t = 0; t < NE; ++t) bin += (t < n_edges && sm.edges[t] <= vd) ? 1 : 0;
      if constexpr (WIDE) reinterpret_cast<uint8_t*>(brow)[cell] = uint8_t(bin);
      else atomicOr(&brow[cell >> 3], uint32_t(bin) << ((cell & 7) * 4));
      atomicAdd(&sm.marg[bin], 1u);
      // the integer behind the value rows and the per-gene sums
      unsigned long long xi;
      if (val_mode == 2) {
        xi = (unsigned long long)llrint(ldexp(vd, fx_shift));
        xi = xi > xi_cap ? xi_cap : xi;
      } else if (val_mode == 3) {
        // segments are sorted by value (k_rank path): #{values < x} + #{values <= x}
        int64_t lo = e0, hi = e;            // first index holding v
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (g_val[mid] < v) lo = mid + 1; else hi = mid; }
        const int64_t lb0 = lo - e0;
        lo = e; hi = e1 - 1;                // last index holding v
        while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if (g_val[mid] > v) hi = mid - 1; else lo = mid; }
        xi = (unsigned long long)((n_cells - m) + lb0 + (lo - e0 + 1));
      } else {
        xi = (unsigned long long)v;         // exact when the data are counts (flagged otherwise)
      }
      s1 += (long long)xi; s2 += (long long)(xi * xi);
      if (val_rows != nullptr) {
        if (val_mode == 1) val_rows[int64_t(g) * val_pitch + cell] = 1;   // detection indicator (jaccard)
        else write_digits(val_rows, val_pitch, g, cell, xi, n_limbs, limb_bits);
      }
    }
    // block-reduce the sums
    for (int d = 16; d >= 1; d >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) { sm.red[warp][0] = (unsigned long long)s1; sm.red[warp][1] = (unsigned long long)s2; }
    __syncthreads();
    if (tid == 32) {
      long long a = 0, b = 0;
      for (int w = 0; w < GB_THREADS / 32; ++w) { a += (long long)sm.red[w][0]; b += (long long)sm.red[w][1]; }
      gsum[g] = a; gsq[g] = b;
    }
    if (tid == 0) {
      n_bins_per_gene[g] = (k <= 1) ? 2 : k + 1;             // metrics.py:62,67
      marg[int64_t(g) * MARG_STRIDE_] = int32_t(m);
    } else if (tid < MARG_STRIDE_) {
      marg[int64_t(g) * MARG_STRIDE_ + tid] = int32_t(sm.marg[tid]);
    }
    // slots 30 / 31 are never edges (n_bins <= 31): 31 carries the real value of one value-row unit
    if (tid < EDGE_STRIDE_)
      edges_out[int64_t(g) * EDGE_STRIDE_ + tid] =
          tid == EDGE_SCALE_SLOT ? (val_mode == 2 ? ldexp(1.0, -fx_shift) : 1.0) : sm.edges[tid];
  }
}

// ---------------------------------------------------------------- k_expand_onehot
// bins_packed [G][pitch] (two cells per byte, low nibble first) -> operand rows
// [n_blocks*32][kp] INT8, block = 32 rows = GPB genes x B bins (+ zero pad rows).
// One warp covers 1024 cells of one 32-row block per iteration: every lane converts the 32
// cells of its 16 packed bytes for each gene of the block and writes 32 bytes per row.
template <int B>
__global__ void k_expand_onehot(const uint32_t* __restrict__ bins_packed, int64_t pitch_words,
                                int G, int8_t* __restrict__ onehot, int64_t kp, int n_blocks) {
  constexpr int GPB = 32 / B;
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t chunks = kp / 1024 + ((kp % 1024) ? 1 : 0);
  for (int blk = blockIdx.y; blk < n_blocks; blk += gridDim.y) {
    for (int64_t ch = int64_t(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5); ch < chunks;
         ch += int64_t(gridDim.x) * warps_per_cta) {
      const int64_t cell0 = ch * 1024 + lane * 32;     // first of this lane's 32 cells
      if (cell0 >= kp) continue;
      int8_t* obase = onehot + int64_t(blk) * 32 * kp + cell0;
#pragma unroll
      for (int sj = 0; sj < GPB; ++sj) {
        const int g = blk * GPB + sj;
        uint4 w = make_uint4(0, 0, 0, 0);
        if (g < G && (cell0 >> 3) + 3 < pitch_words)
          w = *reinterpret_cast<const uint4*>(bins_packed + int64_t(g) * pitch_words + (cell0 >> 3));
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        // bytes of 8 cells each, in cell order
        uint32_t cells[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t lo = ws[i] & 0x0F0F0F0Fu, hi = (ws[i] >> 4) & 0x0F0F0F0Fu;
          cells[2 * i] = __byte_perm(lo, hi, 0x5140);
          cells[2 * i + 1] = __byte_perm(lo, hi, 0x7362);
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const uint32_t pat = uint32_t(b + 1) * 0x01010101u;
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __vcmpeq4(cells[i], pat) & 0x01010101u;
          int8_t* orow = obase + int64_t(sj * B + b) * kp;
          *reinterpret_cast<uint4*>(orow) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(orow + 16) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      // zero pad rows of the block
#pragma unroll
      for (int r = GPB * B; r < 32; ++r) {
        int8_t* orow = obase + int64_t(r) * kp;
        *reinterpret_cast<uint4*>(orow) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(orow + 16) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}


// FP4 variant: the same rows as 4-bit E2M1 one-hot values (0x2 = 1.0), two cells per byte in the
// packing of bins_packed itself (low nibble = even cell), so an operand row is exactly as long
// as a packed-bin row.  Every lane turns 16 packed bytes (32 cells) into 16 bytes per bin row.
template <int B>
__global__ void k_expand_onehot_fp4(const uint32_t* __restrict__ bins_packed, int64_t pitch_words,
                                    int G, uint8_t* __restrict__ onehot, int64_t row_bytes,
                                    int n_blocks) {
  constexpr int GPB = 32 / B;
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t chunks = (row_bytes + 511) / 512;       // 512 bytes = 1024 cells per warp step
  for (int blk = blockIdx.y; blk < n_blocks; blk += gridDim.y) {
    for (int64_t ch = int64_t(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5); ch < chunks;
         ch += int64_t(gridDim.x) * warps_per_cta) {
      const int64_t byte0 = ch * 512 + lane * 16;
      if (byte0 >= row_bytes) continue;
      uint8_t* obase = onehot + int64_t(blk) * 32 * row_bytes + byte0;
#pragma unroll
      for (int sj = 0; sj < GPB; ++sj) {
        const int g = blk * GPB + sj;
        uint4 w = make_uint4(0, 0, 0, 0);
        if (g < G && (byte0 >> 2) + 3 < pitch_words)
          w = *reinterpret_cast<const uint4*>(bins_packed + int64_t(g) * pitch_words + (byte0 >> 2));
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const uint32_t pat = uint32_t(b + 1) * 0x11111111u;
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t x = ws[i] ^ pat;                       // nibble == 0 where bin == b+1
            x = (x | (x >> 2)) & 0x33333333u;
            x = (x | (x >> 1)) & 0x11111111u;               // 1 where the nibble was non-zero
            o[i] = (x ^ 0x11111111u) << 1;                  // 0x2 (E2M1 1.0) where bin == b+1
          }
          *reinterpret_cast<uint4*>(obase + int64_t(sj * B + b) * row_bytes) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
#pragma unroll
      for (int r = GPB * B; r < 32; ++r)
        *reinterpret_cast<uint4*>(obase + int64_t(r) * row_bytes) = make_uint4(0, 0, 0, 0);
    }
  }
}

// One byte per cell (n_bins 16..31): one gene per 32-row block, rows 0..30 = bins 1..31, row 31 zero.
// Every lane turns 16 cells (16 bytes) into 16 bytes per bin row.
__global__ void k_expand_onehot_wide(const uint8_t* __restrict__ bins, int64_t pitch_bytes, int G,
                                     int8_t* __restrict__ onehot, int64_t kp, int n_blocks) {
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t chunks = (kp + 511) / 512;
  for (int blk = blockIdx.y; blk < n_blocks; blk += gridDim.y) {
    for (int64_t ch = int64_t(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5); ch < chunks;
         ch += int64_t(gridDim.x) * warps_per_cta) {
      const int64_t cell0 = ch * 512 + lane * 16;
      if (cell0 >= kp) continue;
      uint4 w = make_uint4(0, 0, 0, 0);
      if (blk < G && cell0 + 16 <= pitch_bytes)
        w = *reinterpret_cast<const uint4*>(bins + int64_t(blk) * pitch_bytes + cell0);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
      int8_t* obase = onehot + int64_t(blk) * 32 * kp + cell0;
#pragma unroll
      for (int b = 0; b < 31; ++b) {
        const uint32_t pat = uint32_t(b + 1) * 0x01010101u;
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __vcmpeq4(ws[i], pat) & 0x01010101u;
        *reinterpret_cast<uint4*>(obase + int64_t(b) * kp) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      *reinterpret_cast<uint4*>(obase + int64_t(31) * kp) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ================================================================ count-data fast path
// For non-negative integer data with values <= 255 (scRNA-seq counts, the documented input) the
// order statistics follow from a per-gene VALUE histogram, and the cell-major -> gene-major regroup
// can be a dense byte transposition staged through shared memory, with no global atomics at all:
//   k_transpose_counts  CSR (cell-major, sorted columns) -> dense uint8 rows raw[gene][cell]: a CTA
//                       owns 128 consecutive cells and walks the genes in ranges of TP_GR; entries
//                       are scattered into a shared-memory tile [gene][cell] and the tile leaves as
//                       full 128-byte lines, so every byte of `raw` is written exactly once (zeros
//                       included: no memset) and the CSR is read once, coalesced
//                       the per-gene value histogram is accumulated on the way in packed byte
//                       counters in shared memory (one flush of a few atomics per gene and cell block)
//   k_rows_to_bins      one CTA per gene: value histogram -> n_unique, k, the 2(k-1) order
//                       statistics, numpy-exact edges, a value -> bin lookup table, marginals, Sum x,
//                       Sum x^2 (gene_lut_warp); then one sweep of the gene's dense row: packed 4-bit
//                       bins and, when they differ from `raw`, the value rows (count digits /
//                       detection / rank digits)
// For signed MI with counts <= 127 `raw` IS the INT8 value-row operand of the sign pass.
// Results are bit-identical to the general path (tests compare both with the oracle).
// k_hist2d (scatter histogram with global atomics) remains for the gene-entropy entry point.
constexpr int VH = 256;   // histogram slots per gene: integer values 1..255

template <typename T>
__global__ void k_hist2d(const T* __restrict__ values, const int32_t* __restrict__ col_idx,
                         int64_t nnz, uint32_t* __restrict__ hist,
                         unsigned long long* __restrict__ flags) {
  unsigned f = 0;
  unsigned long long vmax = 0;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    const T v = values[e];
    if (v > T(0)) {
      const T c = ceil(v);
      if (c != v) f |= 2u;
      const unsigned long long ci = c < T(1.8e19) ? (unsigned long long)c : ~0ull;
      vmax = ci > vmax ? ci : vmax;
      if (c == v && ci < VH) atomicAdd(&hist[int64_t(col_idx[e]) * VH + int(ci)], 1u);
    } else if (v < T(0)) {
      f |= 1u;
    }
    if (!(v - v == T(0))) f |= 8u;           // NaN or infinity
  }
  f = __reduce_or_sync(0xffffffffu, f);
  for (int d = 16; d >= 1; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, vmax, d);
    vmax = o > vmax ? o : vmax;
  }
  if ((threadIdx.x & 31) == 0) {
    if (f) atomicOr(&flags[0], (unsigned long long)f);
    if (vmax) atomicMax(&flags[1], vmax);
  }
}

// One warp per gene.  Lane l owns the 8 histogram slots v = 8l .. 8l+7 (c[i] = #cells with value
// 8l+i; c[0] of lane 0 is ignored).  Writes the value -> bin table (and value -> doubled rank when
// rank_lut != nullptr) and the per-gene outputs of discretize_genes.
template <typename T, int NE>
__device__ void gene_lut_warp(const uint32_t (&c)[8], int lane, int g, int n_bins,
                              const double* __restrict__ q_table, int64_t n_cells,
                              uint8_t* lut, uint32_t* rank_lut,
                              int32_t* __restrict__ n_bins_per_gene, int32_t* __restrict__ marg,
                              double* __restrict__ edges_out, int64_t* __restrict__ gsum,
                              int64_t* __restrict__ gsq) {
  using K = ValKey<T>;
  uint32_t tot = 0, nun = 0;
  long long s1 = 0, s2 = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    tot += c[i]; nun += c[i] ? 1u : 0u;
    const long long v = lane * 8 + i;
    s1 += v * (long long)c[i]; s2 += v * v * (long long)c[i];
  }
  uint32_t inc = tot;
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  const uint32_t lane_excl = inc - tot;
  const int64_t m = __shfl_sync(0xffffffffu, inc, 31);
  for (int d = 16; d >= 1; d >>= 1) nun += __shfl_xor_sync(0xffffffffu, nun, d);
  if (rank_lut != nullptr) {
    // val_mode 3: value -> n_zero + #{values < v} + #{values <= v}; the sums are those of the ranks
    uint32_t run = lane_excl;
    s1 = 0; s2 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long r = (n_cells - m) + 2ll * run + c[i];
      rank_lut[lane * 8 + i] = c[i] ? uint32_t(r) : 0u;
      s1 += c[i] ? r * (long long)c[i] : 0ll;
      s2 += c[i] ? r * r * (long long)c[i] : 0ll;
      run += c[i];
    }
  }
  for (int d = 16; d >= 1; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane == 0) { gsum[g] = s1; gsq[g] = s2; }
  double edge[NE];
#pragma unroll
  for (int t = 0; t < NE; ++t) edge[t] = DBL_MAX;
  int k = 0, n_edges = 0;
  if (m > 0) {
    k = int(nun) < n_bins ? int(nun) : n_bins;          // metrics.py:59
    n_edges = k > 1 ? k - 1 : 0;
    // value (as the integer it is) holding sorted rank r: the slot whose cumulative range has r
    auto value_at = [&](int64_t r) -> int {
      int found = 0;
      uint32_t run = lane_excl;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (r >= int64_t(run) && r < int64_t(run) + int64_t(c[i])) found = lane * 8 + i;
        run += c[i];
      }
      for (int d = 16; d >= 1; d >>= 1) {
        const int o = __shfl_xor_sync(0xffffffffu, found, d);
        found = o > found ? o : found;
      }
      return found;
    };
#pragma unroll
    for (int t = 0; t < NE; ++t) {
      if (t < n_edges) {                                 // warp-uniform
        const double q = q_table[k * QT_DIM_ + t];
        const double vi = __dmul_rn(double(m - 1), q);
        int64_t ra, rb; double gam;
        if (vi >= double(m - 1)) { ra = rb = m - 1; gam = __dadd_rn(vi, 1.0); }
        else { const double fl = floor(vi); ra = int64_t(fl); rb = ra + 1; gam = __dsub_rn(vi, fl); }
        const T a = T(value_at(ra));
        const T b = T(value_at(rb));
        const T d = K::sub(b, a);
        edge[t] = (gam >= 0.5) ? __dsub_rn(double(b), __dmul_rn(double(d), __dsub_rn(1.0, gam)))
                               : __dadd_rn(double(a), __dmul_rn(double(d), gam));
      }
    }
  }
  // lookup table + marginals
  uint32_t mg[NE + 2];
#pragma unroll
  for (int a = 0; a < NE + 2; ++a) mg[a] = 0;
  uint32_t packed[2] = {0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double vd = double(lane * 8 + i);
    int bin = 1;
#pragma unroll
    for (int t = 0; t < NE; ++t) bin += (t < n_edges && edge[t] <= vd) ? 1 : 0;
    if (lane * 8 + i == 0) bin = 0;
    packed[i >> 2] |= uint32_t(bin) << ((i & 3) * 8);
#pragma unroll
    for (int a = 1; a < NE + 2; ++a) mg[a] += (a == bin) ? c[i] : 0u;
  }
  *reinterpret_cast<uint2*>(lut + lane * 8) = make_uint2(packed[0], packed[1]);
#pragma unroll
  for (int a = 1; a < NE + 2; ++a)
    for (int d = 16; d >= 1; d >>= 1) mg[a] += __shfl_xor_sync(0xffffffffu, mg[a], d);
  if (lane == 0) {
    n_bins_per_gene[g] = m == 0 ? 1 : (k <= 1 ? 2 : k + 1);   // metrics.py:55-57,62,67
    marg[int64_t(g) * MARG_STRIDE_] = int32_t(m);
#pragma unroll
    for (int a = 1; a < MARG_STRIDE_; ++a) marg[int64_t(g) * MARG_STRIDE_ + a] = a < NE + 2 ? int32_t(mg[a < NE + 2 ? a : 0]) : 0;
#pragma unroll
    for (int t = 0; t < EDGE_STRIDE_; ++t)
      edges_out[int64_t(g) * EDGE_STRIDE_ + t] = t < NE ? edge[t < NE ? t : 0] : DBL_MAX;
    edges_out[int64_t(g) * EDGE_STRIDE_ + EDGE_SCALE_SLOT] = 1.0;   // one value-row unit = 1 (counts, ranks)
  }
}

// ---------------------------------------------------------------- k_transpose_counts
// flags[0] |= 1 negative value, |= 2 fractional value, |= 4 column indices of a row not ascending
// (the walk below needs canonical CSR; the caller then takes the general path), |= 8 NaN / infinity;
// flags[1] = max over ceil(values).
constexpr int TP_CB = 128;                 // cells per CTA block = bytes per output line
constexpr int TP_STRIDE = TP_CB + 4;       // tile row stride: 33 words, conflict-free both ways
constexpr int TP_GR = 704;                 // genes per range
constexpr int TP_THREADS = 512;
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int TP_ROWS_PER_WARP = TP_CB / TP_WARPS;   // 8, all in flight at once
constexpr int TP_HV = 16;                  // values below this are counted in shared memory
constexpr size_t TP_TILE_BYTES = size_t(TP_GR) * TP_STRIDE;
constexpr size_t TP_HIST_BYTES = size_t(TP_GR) * TP_HV;   // one byte per (gene, value): <= 128 cells
constexpr size_t TP_SMEM = TP_TILE_BYTES + TP_HIST_BYTES + 2 * TP_CB * sizeof(int64_t);

constexpr int TP_MAX_PEERS = 8;            // ranks of one box in the sharded front end

// row_ptr / n_cells describe the CSR rows this launch owns (all of them, or a rank's block of cells);
// blk_offset is the index of its first 128-cell block in the whole matrix (column offset in `raw`).
// VT / CT: storage types of the CSR values and column indices (T / int32_t as handed over, or
// uint8_t / uint16_t for count data packed on the way up: gv_upload_packed_counts); T stays the dtype the
// edges are computed in.
template <typename T, typename VT = T, typename CT = int32_t>
__global__ void __launch_bounds__(TP_THREADS, 2)
k_transpose_counts(const VT* __restrict__ values, const CT* __restrict__ col_idx,
                   const int64_t* __restrict__ row_ptr, int64_t n_cells, int G,
                   uint8_t* __restrict__ raw, int64_t raw_pitch, int64_t n_blocks, int64_t blk_offset,
                   uint32_t* __restrict__ ghist, unsigned long long* __restrict__ flags) {
  // ghist [G][256]: per-gene value histogram, accumulated on the way (small values in packed byte
  // counters in shared memory, flushed once per (cell block, gene range); the rest directly)
  extern __shared__ __align__(16) uint8_t tp_smem[];
  uint8_t* tile = tp_smem;
  uint32_t* hcnt = reinterpret_cast<uint32_t*>(tp_smem + TP_TILE_BYTES);
  // per-row cursors, relative to the block's first entry (a block holds at most 128 x G entries)
  uint32_t* cur_s = reinterpret_cast<uint32_t*>(tp_smem + TP_TILE_BYTES + TP_HIST_BYTES);
  uint32_t* end_s = cur_s + TP_CB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned f = 0;
  unsigned long long vmax = 0;
  for (int i = tid; i < int((TP_TILE_BYTES + TP_HIST_BYTES) / 4); i += TP_THREADS)
    reinterpret_cast<uint32_t*>(tile)[i] = 0u;

  for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const int64_t c0 = blk * TP_CB;
    const int64_t base = row_ptr[c0 < n_cells ? c0 : n_cells];
    const VT* __restrict__ vals_b = values + base;
    const CT* __restrict__ cols_b = col_idx + base;
    if (tid < TP_CB) {
      const int64_t r = c0 + tid;
      cur_s[tid] = r < n_cells ? uint32_t(row_ptr[r] - base) : 0u;
      end_s[tid] = r < n_cells ? uint32_t(row_ptr[r + 1] - base) : 0u;
    }
    __syncthreads();
    for (int g0 = 0; g0 < G; g0 += TP_GR) {
      const int g1 = g0 + TP_GR < G ? g0 + TP_GR : G;
      // ---- scatter this range's entries of my rows into the tile: all eight rows of the warp are in
      // flight at once (16 independent loads per lane and step; the walk is latency-bound)
      {
        constexpr int rg = 0;
        uint32_t cur[TP_ROWS_PER_WARP], end[TP_ROWS_PER_WARP];
#pragma unroll
        for (int k = 0; k < TP_ROWS_PER_WARP; ++k) {
          const int rl = warp + (rg + k) * TP_WARPS;
          cur[k] = cur_s[rl]; end[k] = end_s[rl];
        }
        bool more = true;
        while (more) {
          int32_t cc[TP_ROWS_PER_WARP]; T vv[TP_ROWS_PER_WARP];
#pragma unroll
          for (int k = 0; k < TP_ROWS_PER_WARP; ++k) {
            const uint32_t e = cur[k] + lane;
            const bool ok = e < end[k];
            cc[k] = ok ? int32_t(cols_b[e]) : 0x7fffffff;
            vv[k] = ok ? T(vals_b[e]) : T(0);
          }
          more = false;
#pragma unroll
          for (int k = 0; k < TP_ROWS_PER_WARP; ++k) {
            const bool in = cc[k] < g1;
            const int32_t prev = __shfl_up_sync(0xffffffffu, cc[k], 1);
            if (in) {
              if (cc[k] < g0 || (lane > 0 && cc[k] <= prev)) f |= 4u;
              const T v = vv[k];
              if (v > T(0)) {
                const T cl = ceil(v);
                if (cl != v) f |= 2u;
                const unsigned long long ci = cl < T(1.8e19) ? (unsigned long long)cl : ~0ull;
                vmax = ci > vmax ? ci : vmax;
                if (cc[k] >= g0) {
                  const int gl = cc[k] - g0;
                  const unsigned u = unsigned(ci < 255ull ? ci : 255ull);
                  tile[gl * TP_STRIDE + warp + (rg + k) * TP_WARPS] = uint8_t(u);
                  if (u < unsigned(TP_HV)) atomicAdd(&hcnt[(gl * TP_HV + int(u)) >> 2], 1u << (8 * (u & 3u)));
                  else atomicAdd(&ghist[int64_t(cc[k]) * VH + u], 1u);
                }
              } else if (v < T(0)) {
                f |= 1u;
              }
              if (!(v - v == T(0))) f |= 8u;   // NaN or infinity
            }
            const int cnt = __popc(__ballot_sync(0xffffffffu, in));
            cur[k] += cnt;
            more |= (cnt == 32);
          }
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < TP_ROWS_PER_WARP; ++k) cur_s[warp + (rg + k) * TP_WARPS] = cur[k];
        }
      }
      __syncthreads();
      // ---- the tile leaves as full 128-byte lines and is zeroed for the next range
      for (int gl = warp; gl < g1 - g0; gl += TP_WARPS) {
        uint32_t* w = reinterpret_cast<uint32_t*>(tile + gl * TP_STRIDE) + lane;
        const uint32_t x = *w;
        *w = 0u;
        *reinterpret_cast<uint32_t*>(raw + int64_t(g0 + gl) * raw_pitch + (blk_offset * TP_CB + c0) + 4 * lane) = x;
      }
      // ---- flush the packed byte counters of this (cell block, gene range)
      for (int gl = tid; gl < g1 - g0; gl += TP_THREADS) {
        uint4* hp = reinterpret_cast<uint4*>(hcnt + gl * (TP_HV / 4));
        const uint4 h = *hp;
        if ((h.x | h.y | h.z | h.w) != 0u) {
          const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (hw[q] == 0u) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint32_t cnt = (hw[q] >> (8 * b)) & 0xffu;
              if (cnt) atomicAdd(&ghist[int64_t(g0 + gl) * VH + q * 4 + b], cnt);
            }
          }
          *hp = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      __syncthreads();
    }
  }
  f = __reduce_or_sync(0xffffffffu, f);
  for (int d = 16; d >= 1; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, vmax, d);
    vmax = o > vmax ? o : vmax;
  }
  if (lane == 0) {
    if (f) atomicOr(&flags[0], (unsigned long long)f);
    if (vmax) atomicMax(&flags[1], vmax);
  }
}

// ---------------------------------------------------------------- k_gene_luts
// One warp per gene: value histogram -> the per-gene outputs of discretize_genes and the value -> bin
// table (value -> doubled rank for the rank rows), written to global memory for the row sweep.  Kept out
// of k_rows_to_bins: there one warp built the table while the other seven of the CTA waited, a serial
// third of every gene's time.
template <typename T, int NE>
__global__ void __launch_bounds__(256)
k_gene_luts(const uint32_t* __restrict__ ghist, int G, int n_bins, const double* __restrict__ q_table,
            int64_t n_cells, uint8_t* __restrict__ lut_all, uint32_t* __restrict__ rlut_all,
            int32_t* __restrict__ n_bins_per_gene, int32_t* __restrict__ marg, double* __restrict__ edges_out,
            int64_t* __restrict__ gsum, int64_t* __restrict__ gsq) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < G; g += warps) {
    uint32_t c[8];
    const uint4* hp = reinterpret_cast<const uint4*>(ghist + int64_t(g) * VH + lane * 8);
    const uint4 h0 = hp[0], h1 = hp[1];
    c[0] = lane ? h0.x : 0u; c[1] = h0.y; c[2] = h0.z; c[3] = h0.w;
    c[4] = h1.x; c[5] = h1.y; c[6] = h1.z; c[7] = h1.w;
    gene_lut_warp<T, NE>(c, lane, g, n_bins, q_table, n_cells, lut_all + int64_t(g) * VH,
                         rlut_all ? rlut_all + int64_t(g) * VH : nullptr, n_bins_per_gene, marg, edges_out,
                         gsum, gsq);
  }
}

// ---------------------------------------------------------------- k_rows_to_bins
constexpr int RB_THREADS = 256;

template <typename T, bool WIDE = false>
__global__ void __launch_bounds__(RB_THREADS, 3)
k_rows_to_bins(const uint8_t* __restrict__ raw, int64_t raw_pitch, int64_t kp, int G,
               uint32_t* __restrict__ bins_packed, int64_t pitch_words,
               int8_t* __restrict__ val_rows, int64_t val_pitch, int val_mode, int limb_bits,
               int n_limbs, const uint8_t* __restrict__ lut_all, const uint32_t* __restrict__ rlut_all) {
  // lut_all [G][256] / rlut_all [G][256]: the per-gene tables from k_gene_luts; the row is swept once
  __shared__ __align__(16) uint8_t lut[VH];
  __shared__ uint32_t rlut[VH];
  const int tid = threadIdx.x;
  const int64_t n16 = kp / 16;
  const bool ranks = val_rows != nullptr && val_mode == 3;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const uint4* row = reinterpret_cast<const uint4*>(raw + int64_t(g) * raw_pitch);
    if (tid < VH / 4) reinterpret_cast<uint32_t*>(lut)[tid] = reinterpret_cast<const uint32_t*>(lut_all + int64_t(g) * VH)[tid];
    if (ranks) rlut[tid] = rlut_all[int64_t(g) * VH + tid];
    __syncthreads();
    // ---- one sweep of the row: packed bins and, when they are not `raw` itself, value rows
    uint2* brow = reinterpret_cast<uint2*>(bins_packed + int64_t(g) * pitch_words);
    const bool digits = val_rows != nullptr && (val_mode == 0 || val_mode == 3) &&
                        reinterpret_cast<const uint8_t*>(val_rows) != raw;
    const bool detect = val_rows != nullptr && val_mode == 1;
    const int64_t vrow0 = val_rows != nullptr ? val_row0(g, val_mode == 1 ? 1 : n_limbs) : 0;
    const uint32_t dmask = (1u << limb_bits) - 1u;
    // 16 cells per thread and step; four independent 16-byte loads are in flight per thread before the
    // first is consumed (3 CTAs x 256 threads x 64 bytes = 48 KB per SM: what HBM latency x bandwidth asks)
    auto process = [&](const int64_t i, const uint4 w) {
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
      if constexpr (WIDE) {
        // one byte per cell: the bins of 16 cells are 16 bytes
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t acc = 0u;
          if (ws[k] != 0u) {
#pragma unroll
            for (int b = 0; b < 4; ++b) acc |= uint32_t(lut[(ws[k] >> (8 * b)) & 0xffu]) << (8 * b);
          }
          o[k] = acc;
        }
        reinterpret_cast<uint4*>(brow)[i] = make_uint4(o[0], o[1], o[2], o[3]);
      } else {
        uint32_t pk[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t x = ws[k];
          if (x == 0u) continue;
#pragma unroll
          for (int b = 0; b < 4; ++b)
            pk[k >> 1] |= uint32_t(lut[(x >> (8 * b)) & 0xffu]) << (((k & 1) * 4 + b) * 4);
        }
        brow[i] = make_uint2(pk[0], pk[1]);
      }
      if (detect) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = __vcmpne4(ws[k], 0u) & 0x01010101u;
        *reinterpret_cast<uint4*>(val_rows + vrow0 * val_pitch + 16 * i) = make_uint4(o[0], o[1], o[2], o[3]);
      } else if (digits) {
        for (int l = 0; l < n_limbs; ++l) {
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t acc = 0u;
            if (ws[k] != 0u) {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const uint32_t v = (ws[k] >> (8 * b)) & 0xffu;
                const uint32_t xi = val_mode == 3 ? rlut[v] : v;
                acc |= ((xi >> (l * limb_bits)) & dmask) << (8 * b);
              }
            }
            o[k] = acc;
          }
          *reinterpret_cast<uint4*>(val_rows + (vrow0 + l) * val_pitch + 16 * i) =
              make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    };
    int64_t i = tid;
    for (; i + 3 * RB_THREADS < n16; i += 4 * RB_THREADS) {
      const uint4 w0 = row[i], w1 = row[i + RB_THREADS], w2 = row[i + 2 * RB_THREADS], w3 = row[i + 3 * RB_THREADS];
      process(i, w0);
      process(i + RB_THREADS, w1);
      process(i + 2 * RB_THREADS, w2);
      process(i + 3 * RB_THREADS, w3);
    }
    for (; i < n16; i += RB_THREADS) process(i, row[i]);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- k_gene_entropy
// GeneVectorDataset.get_gene_entropy (data.py:186-203): per gene, np.unique(column,
// return_counts=True), drop the first count (the smallest value's: the zeros when the gene has any,
// data.py:201-202) and scipy.stats.entropy (natural log) of the rest.  One warp per gene over the
// value histogram of the count-data path; the zero count is n_cells minus the positive entries.
__global__ void k_gene_entropy(const uint32_t* __restrict__ hist, int G, int64_t n_cells,
                               double* __restrict__ ent) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < G; g += warps) {
    uint32_t c[8];
    const uint4* hp = reinterpret_cast<const uint4*>(hist + int64_t(g) * VH + lane * 8);
    const uint4 h0 = hp[0], h1 = hp[1];
    c[0] = h0.x; c[1] = h0.y; c[2] = h0.z; c[3] = h0.w; c[4] = h1.x; c[5] = h1.y; c[6] = h1.z; c[7] = h1.w;
    long long pos = 0;
    int first = VH;                       // smallest value present
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pos += c[i];
      if (c[i] && lane * 8 + i < first) first = lane * 8 + i;
    }
    for (int d = 16; d >= 1; d >>= 1) {
      pos += __shfl_xor_sync(0xffffffffu, pos, d);
      const int o = __shfl_xor_sync(0xffffffffu, first, d);
      first = o < first ? o : first;
    }
    const long long zeros = n_cells - pos;
    // counts kept = all positive-value counts, minus the smallest one when there are no zeros
    const int drop = zeros > 0 ? -1 : first;
    long long kept = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) kept += (lane * 8 + i == drop) ? 0 : c[i];
    for (int d = 16; d >= 1; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
    double h = 0.0;
    if (kept > 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (c[i] && lane * 8 + i != drop) {
          const double pk = double(c[i]) / double(kept);
          h -= pk * log(pk);
        }
      }
    }
    for (int d = 16; d >= 1; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
    if (lane == 0) ent[g] = h;
  }
}

// ================================================================ signed fixed-point value rows
// Data with negative values (centred / scaled matrices): the discretisation still treats x <= 0 as
// "not expressed" (metrics.py:53), but the correlation-type scores use the raw values.  Pearson, its
// sign and the cross-correlation do not change when a gene is shifted, so the value rows hold
// rint((x - o_g) * 2^e_g) >= 0 with o_g = min(0, min_g x): the cells without a CSR entry (x = 0) all
// hold q0_g = rint(-o_g * 2^e_g), written by a row fill before the stored entries are scattered.
// edges[g][14] carries q0_g (cosine, which is not shift-invariant, subtracts it again) and
// edges[g][15] the unit 2^-e_g.
constexpr int EDGE_OFFSET_SLOT = GV_EDGE_OFFSET_SLOT;

__device__ __forceinline__ unsigned long long ordered_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_val(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct SignedGene {            // per-gene scratch of the signed path
  unsigned long long min_key, max_key;
  double offset;               // o_g <= 0
  long long q0;
  unsigned long long acc1, acc2;   // Sum (q - q0), Sum (q^2 - q0^2) over stored entries (wrapping)
  int shift, pad;
};

__global__ void k_signed_init(SignedGene* __restrict__ sg, int G) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
    sg[g].min_key = ~0ull; sg[g].max_key = 0ull; sg[g].acc1 = 0ull; sg[g].acc2 = 0ull;
  }
}

template <typename T>
__global__ void k_gene_minmax(const T* __restrict__ values, const int32_t* __restrict__ col_idx,
                              int64_t nnz, SignedGene* __restrict__ sg) {
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    const unsigned long long k = ordered_key(double(values[e]));
    SignedGene* s = sg + col_idx[e];
    if (k < *(volatile unsigned long long*)&s->min_key) atomicMin(&s->min_key, k);
    if (k > *(volatile unsigned long long*)&s->max_key) atomicMax(&s->max_key, k);
  }
}

__global__ void k_signed_params(SignedGene* __restrict__ sg, int G, int total_bits,
                                double* __restrict__ edges_out) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
    SignedGene s = sg[g];
    double lo = 0.0, hi = 0.0;
    if (s.max_key != 0ull) { lo = ordered_val(s.min_key); hi = ordered_val(s.max_key); }
    const double o = lo < 0.0 ? lo : 0.0;
    const double top = (hi > 0.0 ? hi : 0.0) - o;      // largest shifted value (implicit zeros included)
    int shift = 0; long long q0 = 0;
    if (top > 0.0) {
      int ex; frexp(top, &ex);
      shift = total_bits - ex;
      q0 = llrint(ldexp(-o, shift));
      const long long cap = (1ll << total_bits) - 1ll;
      q0 = q0 > cap ? cap : q0;
    }
    sg[g].offset = o; sg[g].shift = shift; sg[g].q0 = q0;
    edges_out[int64_t(g) * EDGE_STRIDE_ + EDGE_OFFSET_SLOT] = q0 ? double(q0) : DBL_MAX;
    edges_out[int64_t(g) * EDGE_STRIDE_ + EDGE_SCALE_SLOT] = ldexp(1.0, -shift);
  }
}

// every cell of the gene's rows gets the digits of q0 (16 cells per thread); cells >= n_cells stay 0
__global__ void k_signed_fill(const SignedGene* __restrict__ sg, int G, int64_t n_cells,
                              int8_t* __restrict__ val_rows, int64_t val_pitch, int n_limbs,
                              int limb_bits) {
  const unsigned mask = (1u << limb_bits) - 1u;
  for (int g = blockIdx.y; g < G; g += gridDim.y) {
    const long long q0 = sg[g].q0;
    if (q0 == 0) continue;
    int8_t* base = val_rows + val_row0(g, n_limbs) * val_pitch;
    for (int64_t c = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 16; c < n_cells;
         c += int64_t(gridDim.x) * blockDim.x * 16) {
      for (int l = 0; l < n_limbs; ++l) {
        const unsigned d = unsigned(q0 >> (l * limb_bits)) & mask;
        int8_t* p = base + int64_t(l) * val_pitch + c;
        if (c + 16 <= n_cells) {
          const uint32_t w = d * 0x01010101u;
          *reinterpret_cast<uint4*>(p) = make_uint4(w, w, w, w);
        } else {
          for (int64_t k = c; k < n_cells; ++k) p[k - c] = int8_t(d);
        }
      }
    }
  }
}

// one warp per CSR row: every stored entry (any sign, explicit zeros too) overwrites its cell
template <typename T>
__global__ void k_signed_scatter(const T* __restrict__ values, const int32_t* __restrict__ col_idx,
                                 const int64_t* __restrict__ row_ptr, int64_t n_cells,
                                 SignedGene* __restrict__ sg, int8_t* __restrict__ val_rows,
                                 int64_t val_pitch, int n_limbs, int limb_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const long long cap = (1ll << (n_limbs * limb_bits)) - 1ll;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_cells; r += warps) {
    const int64_t e0 = row_ptr[r], e1 = row_ptr[r + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
      const int g = col_idx[e];
      SignedGene* s = sg + g;
      long long q = llrint(ldexp(__dsub_rn(double(values[e]), s->offset), s->shift));
      q = q < 0 ? 0 : (q > cap ? cap : q);
      write_digits(val_rows, val_pitch, g, r, (unsigned long long)q, n_limbs, limb_bits);
      const long long q0 = s->q0;
      atomicAdd(&s->acc1, (unsigned long long)(q - q0));
      atomicAdd(&s->acc2, (unsigned long long)(q * q - q0 * q0));
    }
  }
}

__global__ void k_signed_finish(const SignedGene* __restrict__ sg, int G, int64_t n_cells,
                                int64_t* __restrict__ gsum, int64_t* __restrict__ gsq) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
    const long long q0 = sg[g].q0;
    gsum[g] = n_cells * q0 + (long long)sg[g].acc1;
    gsq[g] = n_cells * q0 * q0 + (long long)sg[g].acc2;
  }
}

// ---------------------------------------------------------------- signed rank rows
// Rank rows (val_mode 3) for data with negative values, on per-gene segments that hold EVERY stored
// entry sorted by value (negatives, explicit zeros, positives).  With n_neg negatives and n_zero zero
// cells (implicit + explicit), the doubled average rank minus one of a cell is #{y < x} + #{y <= x}:
//   negatives: their tie range inside the sorted negatives;  zero cells: 2 n_neg + n_zero;
//   positives: 2 (n_neg + n_zero) + their tie range inside the sorted positives.
// Pearson of the ranks is shift-invariant: genes without negatives keep the usual shift by n_zero
// (zero cells hold 0, no fill); the others are stored unshifted and their rows are first filled with
// the zero cells' value (k_signed_fill), then the stored entries are written.
template <typename T>
__global__ void k_rank_params(const int64_t* __restrict__ gene_ptr, const T* __restrict__ s_val, int G,
                              int64_t n_cells, SignedGene* __restrict__ sg) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
    const int64_t e0 = gene_ptr[g], e1 = gene_ptr[g + 1];
    int64_t lo = e0, hi = e1;                       // first index with value >= 0
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (s_val[mid] < T(0)) lo = mid + 1; else hi = mid; }
    const int64_t lbz = lo;
    hi = e1;                                        // first index with value > 0
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (s_val[mid] <= T(0)) lo = mid + 1; else hi = mid; }
    const int64_t ubz = lo;
    const long long n_neg = lbz - e0;
    const long long n_zero = n_cells - (e1 - e0) + (ubz - lbz);
    SignedGene s;
    s.min_key = (unsigned long long)lbz;            // reused: index of the first stored zero
    s.max_key = (unsigned long long)ubz;            // reused: index of the first positive
    s.offset = double(n_neg);
    s.q0 = n_neg > 0 ? 2 * n_neg + n_zero : 0;      // value of the zero cells
    s.acc1 = 0ull; s.acc2 = 0ull;
    s.shift = int(n_neg > 0 ? 0 : 1);               // 1: shifted by n_zero (no negatives)
    s.pad = 0;
    sg[g] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_rank_scatter(const int64_t* __restrict__ gene_ptr, const int32_t* __restrict__ s_cell,
               const T* __restrict__ s_val, int G, int64_t n_cells, const SignedGene* __restrict__ sg,
               int8_t* __restrict__ val_rows, int64_t val_pitch, int n_limbs, int limb_bits,
               int64_t* __restrict__ gsum, int64_t* __restrict__ gsq) {
  __shared__ long long red[8][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int64_t e0 = gene_ptr[g], e1 = gene_ptr[g + 1];
    const SignedGene s = sg[g];
    const int64_t lbz = int64_t(s.min_key), ubz = int64_t(s.max_key);
    const long long n_neg = lbz - e0;
    const long long n_zero = n_cells - (e1 - e0) + (ubz - lbz);
    const long long shift = s.shift ? n_zero : 0;
    const long long q0 = s.q0;
    long long s1 = 0, s2 = 0;
    for (int64_t e = e0 + tid; e < e1; e += 256) {
      const T v = s_val[e];
      long long q;
      if (v == T(0)) {
        q = 2 * n_neg + n_zero - shift;
      } else {
        int64_t lo = v < T(0) ? e0 : ubz, hi = e;             // first index holding v
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (s_val[mid] < v) lo = mid + 1; else hi = mid; }
        const int64_t first = lo;
        lo = e; hi = (v < T(0) ? lbz : e1) - 1;               // last index holding v
        while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if (s_val[mid] > v) hi = mid - 1; else lo = mid; }
        const int64_t last = lo;
        if (v < T(0)) q = (first - e0) + (last - e0 + 1);
        else q = 2 * (n_neg + n_zero) + (first - ubz) + (last - ubz + 1) - shift;
      }
      write_digits(val_rows, val_pitch, g, s_cell[e], (unsigned long long)q, n_limbs, limb_bits);
      s1 += q - q0; s2 += q * q - q0 * q0;
    }
    for (int d = 16; d >= 1; d >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    __syncthreads();
    if (lane == 0) { red[warp][0] = s1; red[warp][1] = s2; }
    __syncthreads();
    if (tid == 0) {
      long long a = 0, b = 0;
      for (int w = 0; w < 8; ++w) { a += red[w][0]; b += red[w][1]; }
      gsum[g] = n_cells * q0 + a;
      gsq[g] = n_cells * q0 * q0 + b;
    }
  }
}

// ================================================================ host-side launchers
#define GV_LAUNCH_CHECK()                                     \
  do {                                                        \
    cudaError_t e_ = cudaGetLastError();                      \
    if (e_ != cudaSuccess) return set_cuda_error(e_, __FILE__, __LINE__); \
  } while (0)

// implemented in gv_rank.cu
size_t segsort_temp_bytes(int64_t nnz, int G, int value_dtype);
int segsort_pairs(void* temp, size_t temp_bytes, const void* keys_in, void* keys_out,
                  const int32_t* cells_in, int32_t* cells_out, int64_t nnz, int G,
                  const int64_t* gene_ptr, int value_dtype, cudaStream_t st);

// hflags: [0] bit0 negative value seen, bit1 fractional value seen; [1] max ceil(value)
static int check_val_config(const DiscretizeArgs& a, const unsigned long long* hflags) {
  if (a.val_rows == nullptr || a.val_mode == 1) return GV_OK;
  const int n_limbs = a.val_limbs, limb_bits = a.val_limb_bits;
  if (n_limbs < 1 || n_limbs > 4 || limb_bits < 1 || limb_bits > 7)
    return set_error(GV_ERR_ARG, "value rows: n_limbs must be in [1,4] and limb_bits in [1,7]");
  const int tb = n_limbs * limb_bits;
  if (hflags[0] & 8ull)
    return set_error(GV_ERR_UNSUPPORTED,
                     "NaN or infinite values: the correlation-type scores are undefined for such genes "
                     "(the discretisation alone treats NaN as not expressed, like the reference)");
  if ((hflags[0] & 1ull) && a.val_mode != 2 && a.val_mode != 3)
    return set_error(GV_ERR_UNSUPPORTED,
                     "count value rows need non-negative values (negative values found): "
                     "use the fixed-point value rows (val_mode 2)");
  if (a.val_mode == 0) {
    if (hflags[0] & 2ull)
      return set_error(GV_ERR_UNSUPPORTED,
                       "co-moment path (count digits) needs integer values (fractional values found): "
                       "use the fixed-point value rows (val_mode 2)");
    if ((hflags[1] >> tb) != 0)
      return set_error(GV_ERR_UNSUPPORTED, "co-moment path: values exceed the configured limb range");
  } else if (a.val_mode == 3) {
    if (((unsigned long long)(2 * a.n_cells) >> tb) != 0)
      return set_error(GV_ERR_UNSUPPORTED, "rank rows: 2*n_cells exceeds the configured limb range");
  }
  return GV_OK;
}

#define GV_CUDA_TRY(expr)                                             \
  do {                                                                \
    cudaError_t e_ = (expr);                                          \
    if (e_ != cudaSuccess) return set_cuda_error(e_, __FILE__, __LINE__); \
  } while (0)

static inline size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

// PACKED: the CSR arrays are uint8 values + uint16 column indices (count data packed on the way up,
// gv_upload_packed_counts); such input always qualifies for the count path.
template <typename T, bool PACKED = false>
static int discretize_impl(const DiscretizeArgs& a, cudaStream_t st) {
  using VT = typename std::conditional<PACKED, uint8_t, T>::type;
  using CT = typename std::conditional<PACKED, uint16_t, int32_t>::type;
  const int G = a.n_genes;
  // workspace carve-up (all 256-byte aligned); the two paths share the front part
  uint8_t* ws = static_cast<uint8_t*>(a.workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = ws + off; off += al256(bytes); return p; };
  unsigned long long* flags = static_cast<unsigned long long*>(take(64));
  double* qtab = static_cast<double*>(take(QT_DIM_ * QT_DIM_ * 8));
  const size_t front = off;
  const VT* values = static_cast<const VT*>(a.values);
  const CT* col_idx = reinterpret_cast<const CT*>(a.col_idx);
  const int sms = device_sm_count();
  int8_t* val_rows = static_cast<int8_t*>(a.val_rows);
  const int n_limbs = a.val_mode == 1 ? 1 : a.val_limbs;
  const bool ranks = val_rows != nullptr && a.val_mode == 3;

  GV_CUDA_TRY(cudaMemsetAsync(flags, 0, 64, st));
  GV_CUDA_TRY(cudaMemcpyAsync(qtab, a.q_table, QT_DIM_ * QT_DIM_ * 8, cudaMemcpyHostToDevice, st));
  unsigned long long hflags[2] = {0, 0};
  const bool wide = a.n_bins > 15;                   // one byte per cell in the bin matrix
  const int64_t kcells = wide ? a.bins_pitch_bytes : a.bins_pitch_bytes * 2;     // cells covered by a bin row

  // ---------------- count-data fast path (tried first unless the caller forces the general one;
  // fixed-point value rows are for data that are not counts, so they go straight to the general path)
  if (a.path != GV_PATH_GENERAL && !(val_rows != nullptr && a.val_mode == 2) && kcells % TP_CB == 0 &&
      (val_rows == nullptr || a.val_pitch == kcells)) {
    // dense uint8 rows: the value-row operand itself when that is what the caller wants
    const bool raw_is_val = val_rows != nullptr && a.val_mode == 0 && n_limbs == 1;
    uint8_t* raw = raw_is_val ? reinterpret_cast<uint8_t*>(val_rows)
                              : static_cast<uint8_t*>(take(size_t(G) * size_t(kcells)));
    uint32_t* ghist = static_cast<uint32_t*>(take(size_t(G) * VH * 4));
    uint8_t* lut_all = static_cast<uint8_t*>(take(size_t(G) * VH));
    uint32_t* rlut_all = ranks ? static_cast<uint32_t*>(take(size_t(G) * VH * 4)) : nullptr;
    if (off > a.workspace_bytes) return set_error(GV_ERR_ARG, "discretize: workspace too small");
    GV_CUDA_TRY(cudaMemsetAsync(ghist, 0, size_t(G) * VH * 4, st));
    auto kt = k_transpose_counts<T, VT, CT>;
    GV_CUDA_TRY(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TP_SMEM)));
    const int64_t n_blocks = kcells / TP_CB;
    const int64_t tgrid = n_blocks < int64_t(sms) * 2 ? n_blocks : int64_t(sms) * 2;
    kt<<<unsigned(tgrid), TP_THREADS, TP_SMEM, st>>>(values, col_idx, a.row_ptr, a.n_cells, G, raw,
                                                     kcells, n_blocks, 0, ghist, flags);
    GV_LAUNCH_CHECK();
    GV_CUDA_TRY(cudaMemcpyAsync(hflags, flags, 16, cudaMemcpyDeviceToHost, st));
    GV_CUDA_TRY(cudaStreamSynchronize(st));
    // (rank rows of data with negative values need every stored entry: general path)
    const bool eligible = (hflags[0] & 14ull) == 0 && hflags[1] < (unsigned long long)VH &&
                          !(ranks && (hflags[0] & 1ull));
    if (eligible) {
      if (a.data_flags_out) { a.data_flags_out[0] = hflags[0]; a.data_flags_out[1] = hflags[1]; }
      int rc = check_val_config(a, hflags);
      if (rc != GV_OK) return rc;
      if (val_rows != nullptr) {
        // rows no gene writes: the two spare rows of every 32-row block when a gene has three
        // digits, and everything past the last gene
        if (n_limbs == 3)
          GV_CUDA_TRY(cudaMemset2DAsync(val_rows + 30 * a.val_pitch, size_t(32) * a.val_pitch, 0,
                                        size_t(2) * a.val_pitch, size_t(a.val_rows_alloc / 32), st));
        const int gpb = 32 / n_limbs;
        const int64_t used = int64_t(G / gpb) * 32 + int64_t(G % gpb) * n_limbs;
        if (a.val_rows_alloc > used)
          GV_CUDA_TRY(cudaMemsetAsync(val_rows + used * a.val_pitch, 0,
                                      size_t(a.val_rows_alloc - used) * a.val_pitch, st));
      }
      auto kl = wide ? k_gene_luts<T, 30> : k_gene_luts<T, 14>;
      kl<<<(G + 7) / 8, 256, 0, st>>>(ghist, G, a.n_bins, qtab, a.n_cells, lut_all, rlut_all, a.n_bins_per_gene,
                                      a.marg, a.edges, a.gsum, a.gsq);
      GV_LAUNCH_CHECK();
      auto kr = wide ? k_rows_to_bins<T, true> : k_rows_to_bins<T, false>;
      kr<<<G < sms * 8 ? G : sms * 8, RB_THREADS, 0, st>>>(
          raw, kcells, kcells, G, static_cast<uint32_t*>(a.bins_packed), int64_t(a.bins_pitch_bytes / 4),
          val_rows, a.val_pitch, a.val_mode, a.val_limb_bits, n_limbs, lut_all, rlut_all);
      GV_LAUNCH_CHECK();
      if (a.path_used) *a.path_used = GV_PATH_COUNTS;
      return GV_OK;
    }
    if (a.path == GV_PATH_COUNTS)
      return set_error(GV_ERR_UNSUPPORTED,
                       "count-data path requested but the values are not integers <= 255 in canonical CSR");
    off = front;   // fall through to the general path, reusing the workspace
  } else if (a.path == GV_PATH_COUNTS) {
    return set_error(GV_ERR_UNSUPPORTED, "count-data path needs a bins pitch of a multiple of 64 bytes "
                                         "matching val_pitch, and count value rows");
  }

  if constexpr (PACKED) {
    return set_error(GV_ERR_UNSUPPORTED, "packed count input (uint8 values, uint16 columns) takes the count path only");
  } else {
  GV_CUDA_TRY(cudaMemsetAsync(a.bins_packed, 0, size_t(G) * a.bins_pitch_bytes, st));
  if (val_rows != nullptr)
    GV_CUDA_TRY(cudaMemsetAsync(val_rows, 0, size_t(a.val_rows_alloc) * a.val_pitch, st));
  // ---------------- general path: gene-major regroup + radix select (any float data)
  uint32_t* cnt = static_cast<uint32_t*>(take(size_t(G) * 4));
  unsigned long long* cursor = static_cast<unsigned long long*>(take(size_t(G) * 8));
  int64_t* gene_ptr = static_cast<int64_t*>(take(size_t(G + 1) * 8));
  int32_t* g_cell = static_cast<int32_t*>(take(size_t(a.nnz) * 4));
  T* g_val = static_cast<T*>(take(size_t(a.nnz) * sizeof(T)));
  int32_t* s_cell = nullptr; T* s_val = nullptr; void* sort_tmp = nullptr; size_t sort_bytes = 0;
  if (ranks) {
    s_cell = static_cast<int32_t*>(take(size_t(a.nnz) * 4));
    s_val = static_cast<T*>(take(size_t(a.nnz) * sizeof(T)));
    sort_bytes = segsort_temp_bytes(a.nnz, G, a.value_dtype);
    sort_tmp = take(sort_bytes);
  }
  if (off > a.workspace_bytes) return set_error(GV_ERR_ARG, "discretize: workspace too small");
  GV_CUDA_TRY(cudaMemsetAsync(cnt, 0, size_t(G) * 4, st));
  GV_CUDA_TRY(cudaMemsetAsync(flags, 0, 64, st));
  if (a.nnz > 0) {
    k_count_positive<T><<<sms * 8, 256, 0, st>>>(values, a.col_idx, a.nnz, cnt, flags);
    GV_LAUNCH_CHECK();
  }
  // data flags decide the limb layout of the value rows; read them back (one small sync)
  GV_CUDA_TRY(cudaMemcpyAsync(hflags, flags, 16, cudaMemcpyDeviceToHost, st));
  GV_CUDA_TRY(cudaStreamSynchronize(st));
  if (a.data_flags_out) { a.data_flags_out[0] = hflags[0]; a.data_flags_out[1] = hflags[1]; }
  int rc = check_val_config(a, hflags);
  if (rc != GV_OK) return rc;
  k_scan_counts<<<1, 1024, 0, st>>>(cnt, gene_ptr, cursor, G);
  GV_LAUNCH_CHECK();
  if (a.nnz > 0) {
    k_scatter<T><<<sms * 8, 256, 0, st>>>(values, a.col_idx, a.row_ptr, a.n_cells, cursor, g_cell, g_val);
    GV_LAUNCH_CHECK();
  }
  // negative values: the value rows come from the signed kernels below instead
  const bool signed_rows = val_rows != nullptr && a.val_mode == 2 && (hflags[0] & 1ull);
  const bool signed_ranks = ranks && (hflags[0] & 1ull);
  int32_t* u_cell = g_cell; T* u_val = g_val;        // the unsorted regroup buffers
  if (ranks && !signed_ranks) {
    // explicit zeros / negatives are dropped by k_scatter, so the segments hold positives only
    rc = segsort_pairs(sort_tmp, sort_bytes, g_val, s_val, g_cell, s_cell, a.nnz, G, gene_ptr,
                       a.value_dtype, st);
    if (rc != GV_OK) return rc;
    g_cell = s_cell; g_val = s_val;
  }
  {
    auto kg = wide ? k_gene_bins<T, true> : k_gene_bins<T, false>;
    const size_t gsm = wide ? sizeof(GeneSmem<64>) : sizeof(GeneSmem<32>);
    GV_CUDA_TRY(cudaFuncSetAttribute(kg, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gsm)));
    kg<<<G < sms * 16 ? (G > 0 ? G : 1) : sms * 16, GB_THREADS, gsm, st>>>(
        gene_ptr, g_cell, g_val, G, a.n_bins, qtab, static_cast<uint32_t*>(a.bins_packed),
        int64_t(a.bins_pitch_bytes / 4), a.n_bins_per_gene, a.marg, a.edges, a.gsum, a.gsq,
        (signed_rows || signed_ranks) ? nullptr : val_rows, a.val_pitch, a.val_mode, a.val_limb_bits, n_limbs,
        a.n_cells);
    GV_LAUNCH_CHECK();
  }
  if (signed_ranks) {
    // second regroup with every stored entry (the buffers of the first one are free again), sorted
    SignedGene* sg = static_cast<SignedGene*>(take(size_t(G) * sizeof(SignedGene)));
    if (off > a.workspace_bytes) return set_error(GV_ERR_ARG, "discretize: workspace too small");
    GV_CUDA_TRY(cudaMemsetAsync(cnt, 0, size_t(G) * 4, st));
    k_count_positive<T, true><<<sms * 8, 256, 0, st>>>(values, a.col_idx, a.nnz, cnt, flags);
    GV_LAUNCH_CHECK();
    k_scan_counts<<<1, 1024, 0, st>>>(cnt, gene_ptr, cursor, G);
    GV_LAUNCH_CHECK();
    k_scatter<T, true><<<sms * 8, 256, 0, st>>>(values, a.col_idx, a.row_ptr, a.n_cells, cursor, u_cell, u_val);
    GV_LAUNCH_CHECK();
    rc = segsort_pairs(sort_tmp, sort_bytes, u_val, s_val, u_cell, s_cell, a.nnz, G, gene_ptr, a.value_dtype, st);
    if (rc != GV_OK) return rc;
    k_rank_params<T><<<(G + 255) / 256, 256, 0, st>>>(gene_ptr, s_val, G, a.n_cells, sg);
    GV_LAUNCH_CHECK();
    dim3 fgrid(unsigned((a.n_cells + 4095) / 4096 < 64 ? (a.n_cells + 4095) / 4096 : 64),
               unsigned(G < 65535 ? G : 65535));
    k_signed_fill<<<fgrid, 256, 0, st>>>(sg, G, a.n_cells, val_rows, a.val_pitch, n_limbs, a.val_limb_bits);
    GV_LAUNCH_CHECK();
    k_rank_scatter<T><<<G < sms * 8 ? G : sms * 8, 256, 0, st>>>(gene_ptr, s_cell, s_val, G, a.n_cells, sg, val_rows,
                                                             a.val_pitch, n_limbs, a.val_limb_bits, a.gsum, a.gsq);
    GV_LAUNCH_CHECK();
  }
  if (signed_rows) {
    SignedGene* sg = static_cast<SignedGene*>(take(size_t(G) * sizeof(SignedGene)));
    if (off > a.workspace_bytes) return set_error(GV_ERR_ARG, "discretize: workspace too small");
    const int gb = (G + 255) / 256;
    k_signed_init<<<gb, 256, 0, st>>>(sg, G);
    GV_LAUNCH_CHECK();
    if (a.nnz > 0) {
      k_gene_minmax<T><<<sms * 8, 256, 0, st>>>(values, a.col_idx, a.nnz, sg);
      GV_LAUNCH_CHECK();
    }
    k_signed_params<<<gb, 256, 0, st>>>(sg, G, n_limbs * a.val_limb_bits, a.edges);
    GV_LAUNCH_CHECK();
    if (a.n_cells > 0) {
      dim3 fgrid(unsigned((a.n_cells + 4095) / 4096 < 64 ? (a.n_cells + 4095) / 4096 : 64),
                 unsigned(G < 65535 ? G : 65535));
      k_signed_fill<<<fgrid, 256, 0, st>>>(sg, G, a.n_cells, val_rows, a.val_pitch, n_limbs, a.val_limb_bits);
      GV_LAUNCH_CHECK();
      k_signed_scatter<T><<<sms * 8, 256, 0, st>>>(values, a.col_idx, a.row_ptr, a.n_cells, sg, val_rows,
                                                    a.val_pitch, n_limbs, a.val_limb_bits);
      GV_LAUNCH_CHECK();
    }
    k_signed_finish<<<gb, 256, 0, st>>>(sg, G, a.n_cells, a.gsum, a.gsq);
    GV_LAUNCH_CHECK();
  }
  if (a.path_used) *a.path_used = GV_PATH_GENERAL;
  return GV_OK;
  }  // !PACKED
}

// ---------------------------------------------------------------- gene entropy, general data
// Per-gene segments holding every stored entry sorted by value: runs of equal values are the unique
// values of np.unique (data.py:198), the zero run also takes the cells without an entry; the count of
// the smallest value is dropped (data.py:201-202: the zeros, unless the gene has negative values or
// no zero at all) and scipy.stats.entropy (natural log) is taken of the rest.
template <typename T>
__global__ void __launch_bounds__(256)
k_gene_entropy_sorted(const int64_t* __restrict__ gene_ptr, const T* __restrict__ s_val, int G,
                      int64_t n_cells, double* __restrict__ ent) {
  __shared__ double red[8];
  __shared__ int64_t s_lbz, s_ubz;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int64_t e0 = gene_ptr[g], e1 = gene_ptr[g + 1];
    const int64_t m = e1 - e0;
    __syncthreads();
    if (tid == 0) {
      int64_t lo = e0, hi = e1;                     // first index with value >= 0
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (s_val[mid] < T(0)) lo = mid + 1; else hi = mid; }
      s_lbz = lo;
      hi = e1;                                      // first index with value > 0
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (s_val[mid] <= T(0)) lo = mid + 1; else hi = mid; }
      s_ubz = lo;
    }
    __syncthreads();
    const int64_t lbz = s_lbz, ubz = s_ubz;
    const int64_t n_zero = n_cells - m + (ubz - lbz);
    auto run_end = [&](int64_t e) {          // last index holding s_val[e]
      const T v = s_val[e];
      int64_t lo = e, hi = e1 - 1;
      while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if (s_val[mid] > v) hi = mid - 1; else lo = mid; }
      return lo;
    };
    // the dropped first value: the smallest negative run, else the zeros, else the smallest positive run
    int64_t first_stored = e0;               // stored runs start here after the drop
    bool zeros_kept = n_zero > 0;
    int64_t dropped;
    if (lbz > e0) { const int64_t r = run_end(e0); dropped = r - e0 + 1; first_stored = r + 1; }
    else if (n_zero > 0) { dropped = n_zero; zeros_kept = false; }
    else if (m > 0) { const int64_t r = run_end(e0); dropped = r - e0 + 1; first_stored = r + 1; }
    else dropped = 0;
    const double kept = double(n_cells - dropped);
    double h = 0.0;
    for (int64_t e = first_stored + tid; e < e1; e += 256) {
      if (e >= lbz && e < ubz) continue;     // stored zeros belong to the zero run below
      if (e == first_stored || e == ubz || s_val[e - 1] != s_val[e]) {
        const double pk = double(run_end(e) - e + 1) / kept;
        h -= pk * log(pk);
      }
    }
    if (tid == 0 && zeros_kept && kept > 0.0) {
      const double pk = double(n_zero) / kept;
      h -= pk * log(pk);
    }
    for (int d = 16; d >= 1; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
    if (lane == 0) red[warp] = h;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      ent[g] = kept > 0.0 ? t : 0.0;
    }
  }
}

template <typename T>
static int gene_entropy_impl(const void* values_v, const int32_t* col_idx, const int64_t* row_ptr,
                             int64_t nnz, int64_t n_cells, int G, double* ent, void* workspace,
                             size_t ws_bytes, cudaStream_t st) {
  const T* values = static_cast<const T*>(values_v);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = ws + off; off += al256(bytes); return p; };
  unsigned long long* flags = static_cast<unsigned long long*>(take(64));
  const size_t front = off;
  uint32_t* hist = static_cast<uint32_t*>(take(size_t(G) * VH * 4));
  if (off > ws_bytes) return set_error(GV_ERR_ARG, "gene entropy: workspace too small");
  GV_CUDA_TRY(cudaMemsetAsync(ws, 0, off, st));
  const int sms = device_sm_count();
  if (nnz > 0) {
    k_hist2d<T><<<sms * 8, 256, 0, st>>>(values, col_idx, nnz, hist, flags);
    GV_LAUNCH_CHECK();
  }
  unsigned long long hflags[2] = {0, 0};
  GV_CUDA_TRY(cudaMemcpyAsync(hflags, flags, 16, cudaMemcpyDeviceToHost, st));
  GV_CUDA_TRY(cudaStreamSynchronize(st));
  if (hflags[0] & 8ull) return set_error(GV_ERR_UNSUPPORTED, "gene entropy: NaN or infinite values");
  if ((hflags[0] & 3ull) == 0 && hflags[1] < (unsigned long long)VH) {
    k_gene_entropy<<<(G + 7) / 8, 256, 0, st>>>(hist, G, n_cells, ent);
    GV_LAUNCH_CHECK();
    return GV_OK;
  }
  // general data: gene-major regroup, per-gene sort, run lengths
  if (!row_ptr) return set_error(GV_ERR_ARG, "gene entropy of non-count data needs row_ptr");
  off = front;
  uint32_t* cnt = static_cast<uint32_t*>(take(size_t(G) * 4));
  unsigned long long* cursor = static_cast<unsigned long long*>(take(size_t(G) * 8));
  int64_t* gene_ptr = static_cast<int64_t*>(take(size_t(G + 1) * 8));
  int32_t* g_cell = static_cast<int32_t*>(take(size_t(nnz) * 4));
  T* g_val = static_cast<T*>(take(size_t(nnz) * sizeof(T)));
  int32_t* s_cell = static_cast<int32_t*>(take(size_t(nnz) * 4));
  T* s_val = static_cast<T*>(take(size_t(nnz) * sizeof(T)));
  const int dt = sizeof(T) == 8 ? GV_F64 : GV_F32;
  const size_t sort_bytes = segsort_temp_bytes(nnz, G, dt);
  void* sort_tmp = take(sort_bytes);
  if (off > ws_bytes) return set_error(GV_ERR_ARG, "gene entropy: workspace too small");
  GV_CUDA_TRY(cudaMemsetAsync(cnt, 0, size_t(G) * 4, st));
  // every stored entry (negatives and explicit zeros too) goes into the per-gene segments
  k_count_positive<T, true><<<sms * 8, 256, 0, st>>>(values, col_idx, nnz, cnt, flags);
  GV_LAUNCH_CHECK();
  k_scan_counts<<<1, 1024, 0, st>>>(cnt, gene_ptr, cursor, G);
  GV_LAUNCH_CHECK();
  k_scatter<T, true><<<sms * 8, 256, 0, st>>>(values, col_idx, row_ptr, n_cells, cursor, g_cell, g_val);
  GV_LAUNCH_CHECK();
  int rc = segsort_pairs(sort_tmp, sort_bytes, g_val, s_val, g_cell, s_cell, nnz, G, gene_ptr, dt, st);
  if (rc != GV_OK) return rc;
  k_gene_entropy_sorted<T><<<G < sms * 8 ? G : sms * 8, 256, 0, st>>>(gene_ptr, s_val, G, n_cells, ent);
  GV_LAUNCH_CHECK();
  return GV_OK;
}

int gene_entropy_dispatch(const void* values, int value_dtype, const int32_t* col_idx,
                          const int64_t* row_ptr, int64_t nnz, int64_t n_cells, int G, double* ent,
                          void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (value_dtype == GV_F32)
    return gene_entropy_impl<float>(values, col_idx, row_ptr, nnz, n_cells, G, ent, workspace, ws_bytes, st);
  if (value_dtype == GV_F64)
    return gene_entropy_impl<double>(values, col_idx, row_ptr, nnz, n_cells, G, ent, workspace, ws_bytes, st);
  return set_error(GV_ERR_ARG, "value dtype must be GV_F32 or GV_F64");
}

// ---------------------------------------------------------------- canonical-CSR check
// One warp per row: every column index must lie in [0, G) and exceed its predecessor in the row
// (sorted, no duplicate entries) -- what scipy calls canonical format and what the discretisation
// kernels assume.  result: 0 canonical, 1 unsorted or duplicate entries, 2 an index out of range.
template <typename CT>
__global__ void k_csr_check(const CT* __restrict__ col_idx, const int64_t* __restrict__ row_ptr,
                            int64_t n_cells, int G, int* __restrict__ result) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  int bad = 0;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_cells; r += warps) {
    const int64_t e0 = row_ptr[r], e1 = row_ptr[r + 1];
    if (e1 < e0) bad = 2;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
      const int32_t c = int32_t(col_idx[e]);
      if (c < 0 || c >= G) bad = 2;
      else if (e > e0 && int32_t(col_idx[e - 1]) >= c && bad < 1) bad = 1;
    }
  }
  bad = __reduce_max_sync(0xffffffffu, bad);
  if (lane == 0 && bad) atomicMax(result, bad);
}

int csr_check_dispatch(const void* col_idx, int col16, const int64_t* row_ptr, int64_t n_cells, int G,
                       int* scratch_dev, int* result_host, cudaStream_t st) {
  GV_CUDA_TRY(cudaMemsetAsync(scratch_dev, 0, sizeof(int), st));
  if (n_cells > 0) {
    const int sms = device_sm_count();
    if (col16)
      k_csr_check<uint16_t><<<sms * 8, 256, 0, st>>>(static_cast<const uint16_t*>(col_idx), row_ptr, n_cells, G,
                                                      scratch_dev);
    else
      k_csr_check<int32_t><<<sms * 8, 256, 0, st>>>(static_cast<const int32_t*>(col_idx), row_ptr, n_cells, G,
                                                     scratch_dev);
    GV_LAUNCH_CHECK();
  }
  GV_CUDA_TRY(cudaMemcpyAsync(result_host, scratch_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  GV_CUDA_TRY(cudaStreamSynchronize(st));
  return GV_OK;
}

// ================================================================ sharded front end
// One process per GPU: every rank owns a contiguous run of 128-cell blocks of the matrix, uploads
// only those CSR rows over its own PCIe link and transposes them into the dense rows of EVERY rank
// (shard_transpose_dispatch: k_transpose_counts with peer destinations).  After a host hand-shake all
// ranks hold the same dense rows; each sums the ranks' partial value histograms (k_sum_hists, peer
// loads) and bins all genes locally (shard_bins_dispatch: k_rows_to_bins), so the bin block is
// replicated without a broadcast.
struct HistSrcs {
  const uint32_t* h[TP_MAX_PEERS];
  int n;
};
__global__ void k_sum_hists(const HistSrcs srcs, int64_t n_words, uint32_t* __restrict__ total) {
  for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n_words;
       i += int64_t(gridDim.x) * blockDim.x * 4) {
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    for (int p = 0; p < srcs.n; ++p) {
      const uint4 v = *reinterpret_cast<const uint4*>(srcs.h[p] + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<uint4*>(total + i) = acc;
  }
}

// Side streams for the peer copies of the sharded front end (one per destination, per device).
static cudaStream_t* peer_streams(int device) {
  static cudaStream_t streams[16][TP_MAX_PEERS] = {};
  static bool made[16] = {};
  if (device < 0 || device >= 16) return nullptr;
  if (!made[device]) {
    for (int p = 0; p < TP_MAX_PEERS; ++p)
      if (cudaStreamCreateWithFlags(&streams[device][p], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    made[device] = true;
  }
  return streams[device];
}

// own_index: which of raw_dests is this rank's own buffer.  The rank's block of cells is transposed into
// its own dense rows; the slab [G][its cells] then goes to every peer as one 2-D peer copy per
// destination on its own stream (copy engines over NVLink: full-rate 25 KB rows, and the SMs are free
// again after the 0.5 ms transposition).  Storing every 128-byte line to all destinations from inside the
// transposition kernel was the first version: 3.9 ms at 2 GPUs but 53 ms at 8 (4-byte peer stores from a
// latency-bound kernel do not fill NVLink) against 9 ms for the copies -- measured, removed.
template <typename T, bool PACKED = false>
static int shard_transpose_impl(const void* values_v, const int32_t* col_idx_v, const int64_t* row_ptr,
                                int64_t n_cells_local, int G, int64_t blk_begin, int64_t blk_end,
                                void* const* raw_dests, int n_dests, int own_index, int64_t raw_pitch,
                                uint32_t* ghist_local, unsigned long long* flags_dev,
                                unsigned long long* flags_host, int phase, cudaStream_t st) {
  using VT = typename std::conditional<PACKED, uint8_t, T>::type;
  using CT = typename std::conditional<PACKED, uint16_t, int32_t>::type;
  const VT* values = static_cast<const VT*>(values_v);
  const CT* col_idx = reinterpret_cast<const CT*>(col_idx_v);
  uint8_t* own_raw = static_cast<uint8_t*>(raw_dests[own_index]);
  // phase bit 0: first chunk of this rank's cells (histogram and flags start from zero); bit 1: last
  // chunk (flags are read back and the stream is synchronised).  A rank may feed its block of cells in
  // several chunks, each one's transposition and peer copies running under the upload of the next.
  if (phase & 1) {
    GV_CUDA_TRY(cudaMemsetAsync(flags_dev, 0, 64, st));
    GV_CUDA_TRY(cudaMemsetAsync(ghist_local, 0, size_t(G) * VH * 4, st));
  }
  const int64_t n_blocks = blk_end - blk_begin;
  if (n_blocks > 0) {
    auto kt = k_transpose_counts<T, VT, CT>;
    GV_CUDA_TRY(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TP_SMEM)));
    const int sms = device_sm_count();
    const int64_t tgrid = n_blocks < int64_t(sms) * 2 ? n_blocks : int64_t(sms) * 2;
    kt<<<unsigned(tgrid), TP_THREADS, TP_SMEM, st>>>(values, col_idx, row_ptr, n_cells_local, G, own_raw, raw_pitch,
                                                     n_blocks, blk_begin, ghist_local, flags_dev);
    GV_LAUNCH_CHECK();
    if (n_dests > 1) {
      int device = 0;
      GV_CUDA_TRY(cudaGetDevice(&device));
      cudaStream_t* ps = peer_streams(device);
      if (!ps) return set_error(GV_ERR_CUDA, "shard transpose: could not create the peer copy streams");
      cudaEvent_t ready, done;
      GV_CUDA_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
      GV_CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
      cudaError_t e = cudaEventRecord(ready, st);
      const uint8_t* src = static_cast<const uint8_t*>(raw_dests[own_index]) + blk_begin * TP_CB;
      const size_t width = size_t(n_blocks) * TP_CB;
      for (int k = 1; k < n_dests && e == cudaSuccess; ++k) {
        const int p = (own_index + k) % n_dests;            // every rank starts with a different peer
        e = cudaStreamWaitEvent(ps[p], ready, 0);
        if (e == cudaSuccess)
          e = cudaMemcpy2DAsync(static_cast<uint8_t*>(raw_dests[p]) + blk_begin * TP_CB, size_t(raw_pitch), src,
                                size_t(raw_pitch), width, size_t(G), cudaMemcpyDeviceToDevice, ps[p]);
        if (e == cudaSuccess) e = cudaEventRecord(done, ps[p]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, done, 0);
      }
      cudaEventDestroy(ready);
      cudaEventDestroy(done);
      if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    }
  }
  if (phase & 2) {
    GV_CUDA_TRY(cudaMemcpyAsync(flags_host, flags_dev, 16, cudaMemcpyDeviceToHost, st));
    GV_CUDA_TRY(cudaStreamSynchronize(st));      // the peer stores / copies are complete
  }
  return GV_OK;
}

int shard_transpose_dispatch(const void* values, int value_dtype, const int32_t* col_idx, const int64_t* row_ptr,
                             int64_t n_cells_local, int G, int64_t blk_begin, int64_t blk_end,
                             void* const* raw_dests, int n_dests, int own_index, int64_t raw_pitch,
                             uint32_t* ghist_local, unsigned long long* flags_dev,
                             unsigned long long* flags_host, int phase, cudaStream_t st) {
  if (n_dests < 1 || n_dests > TP_MAX_PEERS) return set_error(GV_ERR_ARG, "shard transpose: 1..8 destinations");
  if (own_index < 0 || own_index >= n_dests) return set_error(GV_ERR_ARG, "shard transpose: own_index out of range");
  if (raw_pitch % TP_CB != 0 || blk_begin < 0 || blk_end < blk_begin || blk_end * TP_CB > raw_pitch)
    return set_error(GV_ERR_ARG, "shard transpose: cell blocks outside the dense rows");
  if (value_dtype == GV_F32)
    return shard_transpose_impl<float>(values, col_idx, row_ptr, n_cells_local, G, blk_begin, blk_end, raw_dests,
                                       n_dests, own_index, raw_pitch, ghist_local, flags_dev, flags_host, phase, st);
  if (value_dtype == GV_F64)
    return shard_transpose_impl<double>(values, col_idx, row_ptr, n_cells_local, G, blk_begin, blk_end, raw_dests,
                                        n_dests, own_index, raw_pitch, ghist_local, flags_dev, flags_host, phase, st);
  if (value_dtype == GV_COUNTS_U8_COL16)
    return shard_transpose_impl<float, true>(values, col_idx, row_ptr, n_cells_local, G, blk_begin, blk_end,
                                             raw_dests, n_dests, own_index, raw_pitch, ghist_local, flags_dev,
                                             flags_host, phase, st);
  return set_error(GV_ERR_ARG, "value dtype must be GV_F32 or GV_F64");
}

template <typename T>
static int shard_bins_impl(const DiscretizeArgs& a, const uint8_t* raw, const uint32_t* const* hists, int n_hists,
                           uint32_t* ghist_total, double* qtab_dev, uint8_t* lut_all, cudaStream_t st) {
  const int G = a.n_genes;
  const int sms = device_sm_count();
  const int64_t kcells = a.bins_pitch_bytes * 2;
  int8_t* val_rows = static_cast<int8_t*>(a.val_rows);
  HistSrcs srcs;
  srcs.n = n_hists;
  for (int p = 0; p < n_hists; ++p) srcs.h[p] = hists[p];
  GV_CUDA_TRY(cudaMemcpyAsync(qtab_dev, a.q_table, QT_DIM_ * QT_DIM_ * 8, cudaMemcpyHostToDevice, st));
  k_sum_hists<<<sms * 4, 256, 0, st>>>(srcs, int64_t(G) * VH, ghist_total);
  GV_LAUNCH_CHECK();
  if (val_rows != nullptr && a.val_rows_alloc > G)
    GV_CUDA_TRY(cudaMemsetAsync(val_rows + int64_t(G) * a.val_pitch, 0, size_t(a.val_rows_alloc - G) * a.val_pitch, st));
  k_gene_luts<T, 14><<<(G + 7) / 8, 256, 0, st>>>(ghist_total, G, a.n_bins, qtab_dev, a.n_cells, lut_all, nullptr,
                                                  a.n_bins_per_gene, a.marg, a.edges, a.gsum, a.gsq);
  GV_LAUNCH_CHECK();
  k_rows_to_bins<T><<<G < sms * 8 ? G : sms * 8, RB_THREADS, 0, st>>>(
      raw, kcells, kcells, G, static_cast<uint32_t*>(a.bins_packed), int64_t(a.bins_pitch_bytes / 4), val_rows,
      a.val_pitch, a.val_mode, a.val_limb_bits, 1, lut_all, nullptr);
  GV_LAUNCH_CHECK();
  return GV_OK;
}

// raw: this rank's complete dense rows [G][kcells] (== val_rows when value rows are wanted: count
// digits, one per gene); hists: every rank's partial histogram (peer pointers).
int shard_bins_dispatch(const DiscretizeArgs& a, const uint8_t* raw, const uint32_t* const* hists, int n_hists,
                        uint32_t* ghist_total, double* qtab_dev, uint8_t* lut_all, cudaStream_t st) {
  if (n_hists < 1 || n_hists > TP_MAX_PEERS) return set_error(GV_ERR_ARG, "shard bins: 1..8 histograms");
  if (a.n_bins < 1 || a.n_bins > 15)
    return set_error(GV_ERR_UNSUPPORTED, "the sharded front end covers n_bins in [1,15] (4-bit bins)");
  if (a.val_rows != nullptr && (a.val_mode != 0 || a.val_limbs != 1 || a.val_rows != raw ||
                                a.val_pitch != a.bins_pitch_bytes * 2))
    return set_error(GV_ERR_ARG, "shard bins: value rows must be the dense count rows themselves (one digit)");
  if (a.value_dtype == GV_F32 || a.value_dtype == GV_COUNTS_U8_COL16)
    return shard_bins_impl<float>(a, raw, hists, n_hists, ghist_total, qtab_dev, lut_all, st);
  if (a.value_dtype == GV_F64)
    return shard_bins_impl<double>(a, raw, hists, n_hists, ghist_total, qtab_dev, lut_all, st);
  return set_error(GV_ERR_ARG, "value dtype must be GV_F32 or GV_F64");
}

int discretize_dispatch(const DiscretizeArgs& a, cudaStream_t st) {
  if (a.n_bins < 1 || a.n_bins > GV_MAX_BINS) return set_error(GV_ERR_ARG, "n_bins must be in [1,31]");
  if (a.bins_pitch_bytes % 16 != 0 || int64_t(a.bins_pitch_bytes) * (a.n_bins > 15 ? 1 : 2) < a.n_cells)
    return set_error(GV_ERR_ARG, "bins pitch must be a multiple of 16 bytes covering n_cells (n_cells/2 when "
                                 "n_bins <= 15)");
  if (a.value_dtype == GV_F32) return discretize_impl<float>(a, st);
  if (a.value_dtype == GV_F64) return discretize_impl<double>(a, st);
  if (a.value_dtype == GV_COUNTS_U8_COL16) return discretize_impl<float, true>(a, st);
  return set_error(GV_ERR_ARG, "value dtype must be GV_F32 or GV_F64");
}

// sorted = the caller will need the per-gene sort (rank value rows, entropy of non-count data).
// Sized for packed-bin rows of the engine's pitch (n_cells rounded up to 256 cells).
size_t discretize_workspace_bytes(int64_t nnz, int64_t n_cells, int G, int value_dtype, int sorted) {
  const size_t vs = value_dtype == GV_F64 ? 8 : 4;
  const size_t front = al256(64) + al256(QT_DIM_ * QT_DIM_ * 8);
  const size_t kcells = (size_t(n_cells > 0 ? n_cells : 1) + 255) / 256 * 256;
  const size_t counts = al256(size_t(G) * kcells) + al256(size_t(G) * VH * 4) + al256(size_t(G) * VH) +
                        al256(size_t(G) * VH * 4);
  const size_t entropy = al256(size_t(G) * VH * 4);
  size_t general = al256(size_t(G) * 4) + al256(size_t(G) * 8) + al256(size_t(G + 1) * 8) +
                   al256(size_t(nnz) * 4) + al256(size_t(nnz) * vs) + al256(size_t(G) * sizeof(SignedGene));
  if (sorted)
    general += al256(size_t(nnz) * 4) + al256(size_t(nnz) * vs) + al256(segsort_temp_bytes(nnz, G, value_dtype));
  size_t m = counts > general ? counts : general;
  m = m > entropy ? m : entropy;
  return front + m + 256;
}

int expand_onehot_dispatch(const void* bins_packed, int64_t pitch_bytes, int G, int rows_per_gene,
                           void* onehot, int64_t kp, int n_blocks, int operand_format,
                           cudaStream_t st) {
  const uint32_t* bp = static_cast<const uint32_t*>(bins_packed);
  if (operand_format == GV_OPERAND_FP4) {
    if (kp % 256 != 0) return set_error(GV_ERR_ARG, "expand (fp4): K extent must be a multiple of 256 cells");
    if (pitch_bytes % 16 != 0 || pitch_bytes * 2 < kp)
      return set_error(GV_ERR_ARG, "expand: bins pitch must cover the padded K extent");
    const int64_t row_bytes = kp / 2;
    const int64_t chunks = (row_bytes + 511) / 512;
    dim3 grid((unsigned)((chunks + 7) / 8), (unsigned)(n_blocks < 65535 ? n_blocks : 65535));
    uint8_t* oh = static_cast<uint8_t*>(onehot);
    switch (rows_per_gene) {
      case 5:  k_expand_onehot_fp4<5><<<grid, 256, 0, st>>>(bp, pitch_bytes / 4, G, oh, row_bytes, n_blocks); break;
      case 10: k_expand_onehot_fp4<10><<<grid, 256, 0, st>>>(bp, pitch_bytes / 4, G, oh, row_bytes, n_blocks); break;
      default: return set_error(GV_ERR_ARG, "expand (fp4): rows_per_gene must be 5 or 10");
    }
    GV_LAUNCH_CHECK();
    return GV_OK;
  }
  if (operand_format != GV_OPERAND_INT8) return set_error(GV_ERR_ARG, "unknown operand format");
  if (kp % 128 != 0) return set_error(GV_ERR_ARG, "expand: K extent must be a multiple of 128");
  if (rows_per_gene == 32) {
    if (pitch_bytes % 16 != 0 || pitch_bytes < kp)
      return set_error(GV_ERR_ARG, "expand: bins pitch must cover the padded K extent (one byte per cell)");
    const int64_t wchunks = (kp + 511) / 512;
    dim3 wgrid((unsigned)((wchunks + 7) / 8), (unsigned)(n_blocks < 65535 ? n_blocks : 65535));
    k_expand_onehot_wide<<<wgrid, 256, 0, st>>>(static_cast<const uint8_t*>(bins_packed), pitch_bytes, G,
                                                static_cast<int8_t*>(onehot), kp, n_blocks);
    GV_LAUNCH_CHECK();
    return GV_OK;
  }
  if (pitch_bytes % 16 != 0 || pitch_bytes * 2 < kp)
    return set_error(GV_ERR_ARG, "expand: bins pitch must cover the padded K extent");
  const int64_t chunks = (kp + 1023) / 1024;
  dim3 grid((unsigned)((chunks + 7) / 8), (unsigned)(n_blocks < 65535 ? n_blocks : 65535));
  int8_t* oh = static_cast<int8_t*>(onehot);
  switch (rows_per_gene) {
    case 5:  k_expand_onehot<5><<<grid, 256, 0, st>>>(bp, pitch_bytes / 4, G, oh, kp, n_blocks); break;
    case 8:  k_expand_onehot<8><<<grid, 256, 0, st>>>(bp, pitch_bytes / 4, G, oh, kp, n_blocks); break;
    case 10: k_expand_onehot<10><<<grid, 256, 0, st>>>(bp, pitch_bytes / 4, G, oh, kp, n_blocks); break;
    case 16: k_expand_onehot<16><<<grid, 256, 0, st>>>(bp, pitch_bytes / 4, G, oh, kp, n_blocks); break;
    default: return set_error(GV_ERR_ARG, "expand: rows_per_gene must be 5, 8, 10, 16 or 32");
  }
  GV_LAUNCH_CHECK();
  return GV_OK;
}

}  // namespace gv
