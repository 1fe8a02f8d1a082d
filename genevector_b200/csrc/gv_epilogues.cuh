// gv_epilogues.cuh -- epilogue functors for gram_tiles_kernel (gv_gram.cuh).
//
//   DumpEpi      raw INT32 accumulators -> HBM (tests only: bit-exact joint-count check)
//   MiEpi<B>     joint counts of (gene,bin) x (gene,bin) -> mutual information per gene pair,
//                fused, counts never leave the SM.  Restates the per-pair body of the reference
//                (metrics.py:74-81 _mi_from_joint, metrics.py:179-196, src/lib.rs:62-88) on the
//                recast of SURVEY.md section 7.2: the zero-bin row/column of the joint table and the
//                normaliser come from the per-gene marginals, only non-zero bins are contracted.
//   MomentEpi<K> co-moment Sum_c x_ic*x_jc (or co-detection counts) -> Pearson sign / Pearson /
//                cosine / Jaccard (metrics.py:106-111, 414-420, 452-461, 434-449)
#pragma once
#include "gv_gram.cuh"

namespace gv {

constexpr int MARG_STRIDE = 32;  // int32 per gene: [0] = #cells with x>0, [a] = #cells in bin a (GV_MARG_STRIDE)
constexpr int EDGE_STRIDE = 32;  // float64 per gene (GV_EDGE_STRIDE); slot 30 = value-row integer of a zero cell

__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(GRAM_EPI_BAR), "n"(GRAM_EPI_WARPS * 32) : "memory");
}

// ------------------------------------------------------------------ DumpEpi
struct DumpParams {
  int32_t* out;     // [rows][ld] accumulators, row/col = operand row indices
  int64_t ld;
  int rows;         // operand rows (bounds)
};
// ACT = data rows per 32-row block on the B side (see GramCfg): accumulator column
// cb*ACT + c holds operand row col0 + cb*32 + c.
// FP32 accumulators of the FP4 operand path hold exact integers: back to INT32.
template <bool FP4>
__device__ __forceinline__ void acc_to_int(uint32_t (&v)[32]) {
  if constexpr (FP4) {
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = uint32_t(__float2int_rn(__uint_as_float(v[c])));
  }
}

template <int ACT, bool FP4 = false>
struct DumpEpi {
  using Params = DumpParams;
  static constexpr size_t SMEM_BYTES = 16;
  static constexpr int ACTIVE_ROWS = ACT;
  static constexpr bool OPERAND_FP4 = FP4;
  const Params& p;
  __device__ DumpEpi(const Params& p_, uint8_t*) : p(p_) {}
  __device__ void prepare(const EpiTile&) {}
  __device__ void run(const EpiTile& et) {
    if (p.out == nullptr) return;      // discard mode: the main loop alone (tools/probe/l2_probe.py)
    const int row = et.row0 + et.quarter * 32 + et.lane;
#pragma unroll 1
    for (int cbi = 0; cbi < 4; ++cbi) {
      const int cb = et.half * 4 + cbi;
      uint32_t v[32];
      tmem_ld_32x32(et.tmem_acc + (uint32_t(et.quarter * 32) << 16) + uint32_t(cb * ACT), v);
      tmem_ld_wait();
      acc_to_int<FP4>(v);
      if (row < p.rows) {
#pragma unroll
        for (int c = 0; c < ACT; ++c) {
          const int col = et.col0 + cb * 32 + c;
          if (col < p.rows) p.out[int64_t(row) * p.ld + col] = int32_t(v[c]);
        }
      }
    }
  }
};

// ------------------------------------------------------------------ MiEpi
struct MiParams {
  const int32_t* marg;   // [G][MARG_STRIDE]
  const int8_t* sgn;     // [G][ld_sgn] Pearson sign in {-1,0,+1}, or nullptr (unsigned MI)
  int64_t ld_sgn;
  float* out;            // [G][ld_out] scores; (i,j) and (j,i) both written, diagonal 0
  int64_t ld_out;
  int G;
};

// Sum over the lanes [seg0, seg0+B) of a gene segment, result valid in lane seg0.
template <int B, typename T>
__device__ __forceinline__ T seg_reduce(T x, int pos) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    if (d < B) {   // compile-time prunable: B <= 16 never needs d == 16
      const T y = __shfl_down_sync(0xffffffffu, x, d);
      if (pos + d < B) x += y;
    }
  }
  return x;
}

template <int B, bool FP4 = false>
struct MiEpi {
  using Params = MiParams;
  static constexpr bool OPERAND_FP4 = FP4;
  static constexpr int GPB = 32 / B;          // genes per 32-row block of the operand
  static constexpr int ACTIVE = GPB * B;      // rows of a block that carry data
  // B side staged densely when a block has exactly 30 data rows (B = 10, 5): N = 240
  static constexpr int ACTIVE_ROWS = (ACTIVE == 30) ? 30 : 32;
  static constexpr int CG_PER_TILE = 8 * GPB; // column genes per tile
  struct Meta {
    int cm[GRAM_TILE_N];        // marginal count of the (gene,bin) behind each column
    float cinv[GRAM_TILE_N];    // 1 / cm (0 where cm == 0)
    int cg_nnz[CG_PER_TILE];    // #cells with x>0 per column gene (0 => skip pair)
  };
  // warp-private transpose scratch: 32 rows x ACTIVE columns, odd row stride (no bank conflicts)
  static constexpr int SCR_STRIDE = (ACTIVE % 2 == 0) ? ACTIVE + 1 : ACTIVE;
  static constexpr size_t SCR_INTS = 32 * SCR_STRIDE;
  static constexpr size_t SMEM_BYTES = sizeof(Meta) + GRAM_EPI_WARPS * SCR_INTS * 4 + 16;

  const Params& p;
  Meta* meta;
  int* scr_all;
  __device__ MiEpi(const Params& p_, uint8_t* smem) : p(p_) {
    meta = reinterpret_cast<Meta*>(smem);
    scr_all = reinterpret_cast<int*>(smem + sizeof(Meta));
  }

  // Column metadata of the tile: one epilogue thread per accumulator column.
  __device__ void prepare(const EpiTile& et) {
    epi_bar_sync();   // everyone is done with the previous tile's metadata
    const int c = et.epi_tid;                 // accumulator column
    const int blk = c / ACTIVE_ROWS;          // 32-row operand block behind it
    const int within = c - blk * ACTIVE_ROWS;
    int m = 0;
    if (blk < 8 && within < ACTIVE) {
      const int sj = within / B, b = within - sj * B;
      const int gj = ((et.col0 >> 5) + blk) * GPB + sj;
      if (gj < p.G) {
        if (1 + b < MARG_STRIDE) m = p.marg[int64_t(gj) * MARG_STRIDE + 1 + b];  // B=32: bin 32 never occurs
        if (b == 0) meta->cg_nnz[blk * GPB + sj] = p.marg[int64_t(gj) * MARG_STRIDE];
      } else if (b == 0) {
        meta->cg_nnz[blk * GPB + sj] = 0;
      }
    }
    meta->cm[c] = m;
    meta->cinv[c] = m > 0 ? 1.0f / float(m) : 0.0f;
    epi_bar_sync();
  }

  __device__ void run(const EpiTile& et) {
    const int lane = et.lane;
    const bool active = lane < ACTIVE;
    const int si = lane / B;
    const int pos = lane - si * B;      // bin index a-1 of this row
    const int seg0 = lane - pos;
    const int gi = ((et.row0 >> 5) + et.quarter) * GPB + si;
    const bool row_ok = active && gi < p.G;
    int nnz_i = 0, m_ia = 0;
    if (row_ok) {
      nnz_i = p.marg[int64_t(gi) * MARG_STRIDE];
      if (1 + pos < MARG_STRIDE) m_ia = p.marg[int64_t(gi) * MARG_STRIDE + 1 + pos];
    }
    const float inv_rx = m_ia > 0 ? 1.0f / float(m_ia) : 0.0f;
    int* scr = scr_all + ((et.quarter + 4 * et.half) * SCR_INTS);

#pragma unroll 1
    for (int cbi = 0; cbi < 4; ++cbi) {
      const int cb = et.half * 4 + cbi;
      uint32_t v[32];
      tmem_ld_32x32(et.tmem_acc + (uint32_t(et.quarter * 32) << 16) + uint32_t(cb * ACTIVE_ROWS), v);
      tmem_ld_wait();
      acc_to_int<FP4>(v);
      // stage this warp's 32x32 count block for the column sums (bank = lane + c: no conflicts)
#pragma unroll
      for (int c = 0; c < ACTIVE; ++c) scr[lane * SCR_STRIDE + c] = int(v[c]);
      __syncwarp();

#pragma unroll
      for (int sj = 0; sj < GPB; ++sj) {
        const int cgi = cb * GPB + sj;
        const int gj = ((et.col0 >> 5) + cb) * GPB + sj;
        const int nnz_j = meta->cg_nnz[cgi];
        const int colb = cb * ACTIVE_ROWS + sj * B;    // first accumulator column of gene j

        // row sum over b (this row's bin a against all non-zero bins of gene j)
        int r = 0;
#pragma unroll
        for (int b = 0; b < B; ++b) r += int(v[sj * B + b]);
        // column sum over a for b = pos+1 (all non-zero bins of gene i against bin b of gene j)
        int csum = 0;
        if (active) {
#pragma unroll
          for (int a2 = 0; a2 < B; ++a2) csum += scr[(seg0 + a2) * SCR_STRIDE + sj * B + pos];
        }
        // S = #cells where both genes are non-zero; T = #cells where either is (metrics.py:124)
        int S = seg_reduce<B>(r, pos);
        S = __shfl_sync(0xffffffffu, S, seg0);
        const int T = nnz_i + nnz_j - S;
        const float Tf = float(T);

        float acc = 0.0f;
        // inner cells J[a,b], a,b >= 1: px = m_i[a]/T, py = m_j[b]/T
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const int J = int(v[sj * B + b]);
          const float Jf = float(J);
          const float x = (Jf * inv_rx) * (Tf * meta->cinv[colb + b]);
          const float term = Jf * __log2f(x);
          acc += J > 0 ? term : 0.0f;
        }
        // J[a,0] = m_i[a] - r : gene i in bin a, gene j zero.  py[0] = (nnz_i - S)/T
        {
          const int J = m_ia - r;
          const float Jf = float(J);
          const float x = (Jf * inv_rx) * __fdividef(Tf, float(nnz_i - S));
          const float term = Jf * __log2f(x);
          acc += J > 0 ? term : 0.0f;
        }
        // J[0,b] = m_j[b] - csum, b = pos+1 : gene i zero, gene j in bin b.  px[0] = (nnz_j - S)/T
        {
          const int J = meta->cm[colb + pos] - csum;
          const float Jf = float(J);
          const float x = __fdividef(Jf, float(nnz_j - S)) * (Tf * meta->cinv[colb + pos]);
          const float term = Jf * __log2f(x);
          acc += (active && J > 0) ? term : 0.0f;
        }
        const float total = seg_reduce<B>(acc, pos);

        if (pos == 0 && row_ok && gj < p.G) {
          if (gi < gj) {
            float mi = (nnz_i > 0 && nnz_j > 0 && T > 0) ? total / Tf : 0.0f;
            if (p.sgn != nullptr) mi *= float(p.sgn[int64_t(gi) * p.ld_sgn + gj]);
            p.out[int64_t(gi) * p.ld_out + gj] = mi;
            p.out[int64_t(gj) * p.ld_out + gi] = mi;
          } else if (gi == gj) {
            p.out[int64_t(gi) * p.ld_out + gi] = 0.0f;
          }
        }
      }
      __syncwarp();   // scratch is rewritten by the next column block
    }
  }
};

// ------------------------------------------------------------------ MomentEpi
enum MomentKind : int { MOM_SIGN = 0, MOM_PEARSON = 1, MOM_COSINE = 2, MOM_JACCARD = 3 };

struct MomentParams {
  const int64_t* gsum;   // [G] Sum_c x            (sign / pearson / cosine)
  const int64_t* gsq;    // [G] Sum_c x^2          (sign / pearson / cosine)
  const int32_t* marg;   // [G][MARG_STRIDE], [0] = #cells with x>0 (jaccard)
  const double* edges;   // [G][32] or nullptr: slot 30 = value-row integer of a zero cell (cosine)
  int8_t* sgn;           // [G][ld_sgn]  (MOM_SIGN)
  int64_t ld_sgn;
  float* out;            // [G][ld_out]  (other kinds)
  int64_t ld_out;
  int G;
  int64_t n_cells;
  int limb_bits;         // digit width of the value rows (two-digit operands)
  int sign_zero;         // MOM_SIGN: value written where the correlation is 0 or undefined: 0 follows
                         // np.sign (metrics.py:111,134), +1 the Rust tier (src/lib.rs:91-94)
};

// 128-bit integers keep n*Sum(xy) - Sum(x)*Sum(y) exact for 21-bit fixed-point values over 2^18 cells.
typedef __int128 i128;
__device__ __forceinline__ double i128_to_double(i128 v) {
  const bool neg = v < 0;
  const unsigned __int128 m = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
  const double d = double((unsigned long long)(m >> 64)) * 18446744073709551616.0 +
                   double((unsigned long long)m);
  return neg ? -d : d;
}

// L = operand rows per gene: base-2^limb_bits digits of the (count / fixed-point / doubled rank)
// value, least significant first.  The value rows use the same 32-row block layout as the one-hot
// operand: a block holds GPB = 32/L genes (row = block*32 + slot*L + digit; L = 3 leaves two zero
// rows per block), so a gene's digits never straddle a 32-lane TMEM quarter.
// Sum_c x_i x_j = Sum_{p,q} 2^{limb_bits (p+q)} acc[(i,p),(j,q)]: the q sum is taken in registers
// (adjacent columns), the p sum across the L adjacent lanes of gene i.
template <int KIND, int L>
struct MomentEpi {
  static_assert(L >= 1 && L <= 4, "one to four digits per gene");
  using Params = MomentParams;
  static constexpr int GPB = 32 / L;              // genes per 32-row block
  static constexpr int ACTIVE = GPB * L;          // data rows per block
  static constexpr int GENES_PER_TILE = 8 * GPB;  // column genes per tile
  struct Meta {
    int64_t csum[GENES_PER_TILE];
    int64_t csq[GENES_PER_TILE];
    int64_t coff[GENES_PER_TILE];
  };
  static constexpr size_t SMEM_BYTES = sizeof(Meta) + 16;
  static constexpr int ACTIVE_ROWS = 32;
  static constexpr bool OPERAND_FP4 = false;
  const Params& p;
  Meta* meta;
  __device__ MomentEpi(const Params& p_, uint8_t* smem) : p(p_) {
    meta = reinterpret_cast<Meta*>(smem);
  }
  // value-row integer of the gene's zero cells (genes with negative values are stored shifted)
  __device__ int64_t zero_offset(int g) const {
    if (p.edges == nullptr) return 0;
    const double o = p.edges[int64_t(g) * EDGE_STRIDE + 30];
    return o < 1e300 ? int64_t(o) : 0;
  }
  __device__ void prepare(const EpiTile& et) {
    epi_bar_sync();
    const int c = et.epi_tid;
    if (c < GENES_PER_TILE) {
      const int gj = (et.col0 >> 5) * GPB + c;
      if constexpr (KIND == MOM_JACCARD) {
        meta->csum[c] = gj < p.G ? int64_t(p.marg[int64_t(gj) * MARG_STRIDE]) : 0;
        meta->csq[c] = 0;
      } else {
        meta->csum[c] = gj < p.G ? p.gsum[gj] : 0;
        meta->csq[c] = gj < p.G ? p.gsq[gj] : 0;
        if constexpr (KIND == MOM_COSINE) meta->coff[c] = gj < p.G ? zero_offset(gj) : 0;
      }
    }
    epi_bar_sync();
  }
  __device__ void run(const EpiTile& et) {
    const int si = et.lane / L;
    const int digit = et.lane - si * L;
    const int gi = ((et.row0 >> 5) + et.quarter) * GPB + si;
    const bool row_ok = et.lane < ACTIVE && gi < p.G;
    int64_t sx = 0, sxx = 0;
    if (row_ok) {
      if constexpr (KIND == MOM_JACCARD) {
        sx = int64_t(p.marg[int64_t(gi) * MARG_STRIDE]);
      } else {
        sx = p.gsum[gi];
        sxx = p.gsq[gi];
      }
    }
    const i128 n = i128(p.n_cells);
    const i128 var_i = n * sxx - i128(sx) * sx;
    int64_t ox = 0;
    i128 txx = 0;
    if constexpr (KIND == MOM_COSINE) {
      ox = row_ok ? zero_offset(gi) : 0;
      txx = i128(sxx) - 2 * i128(ox) * sx + n * ox * ox;
    }
    const int lb = p.limb_bits;
#pragma unroll 1
    for (int cbi = 0; cbi < 4; ++cbi) {
      const int cb = et.half * 4 + cbi;
      uint32_t v[32];
      tmem_ld_32x32(et.tmem_acc + (uint32_t(et.quarter * 32) << 16) + uint32_t(cb * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < GPB; ++c) {
        const int cg = cb * GPB + c;            // column gene within the tile
        const int gj = ((et.col0 >> 5) + cb) * GPB + c;
        // accumulators are read as unsigned 32-bit (sums of digit products < 2^32)
        int64_t sxy = int64_t(v[c * L]);
        if constexpr (L > 1) {
#pragma unroll
          for (int q = 1; q < L; ++q) sxy += int64_t(v[c * L + q]) << (q * lb);
          sxy <<= (digit * lb);
          sxy = seg_reduce<L>(sxy, digit);      // add the other digit rows of gene i
        }
        if (digit != 0 || !row_ok || gj >= p.G || gi > gj) continue;
        const int64_t sy = meta->csum[cg];
        if constexpr (KIND == MOM_SIGN) {
          int8_t s = 0;
          if (gi != gj) {
            const int64_t syy = meta->csq[cg];
            const i128 var_j = n * syy - i128(sy) * sy;
            const i128 cov = n * sxy - i128(sx) * sy;
            const int8_t z = int8_t(p.sign_zero);
            s = (var_i == 0 || var_j == 0) ? z : (cov > 0 ? int8_t(1) : (cov < 0 ? int8_t(-1) : z));
          }
          p.sgn[int64_t(gi) * p.ld_sgn + gj] = s;
          p.sgn[int64_t(gj) * p.ld_sgn + gi] = s;
        } else {
          float r = 0.0f;
          if (gi != gj) {
            if constexpr (KIND == MOM_PEARSON) {
              const int64_t syy = meta->csq[cg];
              const i128 var_j = n * syy - i128(sy) * sy;
              const i128 cov = n * sxy - i128(sx) * sy;
              // 0/0 -> NaN as np.corrcoef
              double q = i128_to_double(cov) / sqrt(i128_to_double(var_i) * i128_to_double(var_j));
              q = q > 1.0 ? 1.0 : (q < -1.0 ? -1.0 : q);
              r = float(q);
            } else if constexpr (KIND == MOM_COSINE) {
              // cosine is not shift-invariant: undo the per-gene offsets (zero for non-negative data)
              const int64_t oy = meta->coff[cg];
              const i128 txy = i128(sxy) - i128(oy) * sx - i128(ox) * sy + n * ox * oy;
              const i128 tyy = i128(meta->csq[cg]) - 2 * i128(oy) * sy + n * oy * oy;
              r = (txx == 0 || tyy == 0)
                      ? 0.0f
                      : float(i128_to_double(txy) / sqrt(i128_to_double(txx) * i128_to_double(tyy)));
            } else {  // JACCARD: float32 division of exact integers, as metrics.py:443-447
              int64_t uni = sx + sy - sxy;
              if (uni == 0) uni = 1;
              r = float(sxy) / float(uni);
            }
          }
          p.out[int64_t(gi) * p.ld_out + gj] = r;
          p.out[int64_t(gj) * p.ld_out + gi] = r;
        }
      }
    }
  }
};

// ------------------------------------------------------------------ XcorrEpi
// Cross-correlation between expression (A side: LA digit rows per gene) and graph-aggregated
// expression (B side: LB digit rows per gene), genevector/_graph_targets.py:45-50:
//   X_std = (X - mean) / (std + 1e-8), likewise X_agg;  xcorr = X_std^T @ X_agg_std / n_cells
// from exact integer co-moments: Sum (x - mx)(y - my) = (n Sxy - Sx Sy) / n in value-row units, times
// the per-gene unit scales.  Writes the non-symmetric xcorr[a][b]; gv_symmetrize finishes (:51-52).
struct XcorrParams {
  const int64_t* asum; const int64_t* asq; const double* ascale; int64_t ascale_stride;   // A side per gene
  const int64_t* bsum; const int64_t* bsq; const double* bscale;                          // B side per gene
  float* out; int64_t ld_out;
  int G; int64_t n_cells; int limb_bits;
};

template <int LA, int LB>
struct XcorrEpi {
  using Params = XcorrParams;
  static constexpr int GPB_A = 32 / LA, ACTIVE_A = GPB_A * LA;
  static constexpr int GPB_B = 32 / LB;
  static constexpr int GENES_PER_TILE_B = 8 * GPB_B;
  struct Meta {
    int64_t csum[GENES_PER_TILE_B];
    double cfac[GENES_PER_TILE_B];      // scale_b / (std_b + 1e-8)
  };
  static constexpr size_t SMEM_BYTES = sizeof(Meta) + 16;
  static constexpr int ACTIVE_ROWS = 32;
  static constexpr bool OPERAND_FP4 = false;
  const Params& p;
  Meta* meta;
  __device__ XcorrEpi(const Params& p_, uint8_t* smem) : p(p_) { meta = reinterpret_cast<Meta*>(smem); }

  // scale / (population std + 1e-8) of a gene from its integer sums
  __device__ static double std_factor(int64_t s, int64_t sq, double scale, int64_t n) {
    const i128 var = i128(n) * sq - i128(s) * s;
    const double sd = sqrt(i128_to_double(var)) / double(n) * scale;
    return scale / (sd + 1e-8);
  }
  __device__ void prepare(const EpiTile& et) {
    epi_bar_sync();
    const int c = et.epi_tid;
    if (c < GENES_PER_TILE_B) {
      const int gj = (et.col0 >> 5) * GPB_B + c;
      const bool ok = gj < p.G;
      meta->csum[c] = ok ? p.bsum[gj] : 0;
      meta->cfac[c] = ok ? std_factor(p.bsum[gj], p.bsq[gj], p.bscale[gj], p.n_cells) : 0.0;
    }
    epi_bar_sync();
  }
  __device__ void run(const EpiTile& et) {
    const int si = et.lane / LA;
    const int digit = et.lane - si * LA;
    const int gi = ((et.row0 >> 5) + et.quarter) * GPB_A + si;
    const bool row_ok = et.lane < ACTIVE_A && gi < p.G;
    int64_t sx = 0;
    double fx = 0.0;
    if (row_ok) {
      sx = p.asum[gi];
      fx = std_factor(sx, p.asq[gi], p.ascale ? p.ascale[int64_t(gi) * p.ascale_stride] : 1.0, p.n_cells);
    }
    const double inv_n2 = 1.0 / (double(p.n_cells) * double(p.n_cells));
    const int lb = p.limb_bits;
#pragma unroll 1
    for (int cbi = 0; cbi < 4; ++cbi) {
      const int cb = et.half * 4 + cbi;
      uint32_t v[32];
      tmem_ld_32x32(et.tmem_acc + (uint32_t(et.quarter * 32) << 16) + uint32_t(cb * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < GPB_B; ++c) {
        const int cg = cb * GPB_B + c;
        const int gj = ((et.col0 >> 5) + cb) * GPB_B + c;
        int64_t sxy = int64_t(v[c * LB]);
#pragma unroll
        for (int q = 1; q < LB; ++q) sxy += int64_t(v[c * LB + q]) << (q * lb);
        if constexpr (LA > 1) {
          sxy <<= (digit * lb);
          sxy = seg_reduce<LA>(sxy, digit);
        }
        if (digit != 0 || !row_ok || gj >= p.G) continue;
        const i128 cov = i128(p.n_cells) * sxy - i128(sx) * meta->csum[cg];
        p.out[int64_t(gi) * p.ld_out + gj] = float(i128_to_double(cov) * inv_n2 * fx * meta->cfac[cg]);
      }
    }
  }
};

}  // namespace gv
