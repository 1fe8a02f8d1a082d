/* gvb200.h -- C ABI of libgvb200.so: the B200 (sm_100a) all-pairs gene co-expression engine.
 *
 * Drop-in boundary for the hot path of nceglia/genevector (citations are into the reference):
 *   - the FFI it replaces is genevector._rust.compute_mi_pairs(x_disc, n_bins_per_gene, corr_signs)
 *     (src/lib.rs:9-15, stub genevector/_rust.pyi:1-5), called from compute_mi_rust
 *     (genevector/metrics.py:312-339), plus the Python steps around it that the reference runs on
 *     the CPU for every tier: discretize_genes (metrics.py:25-69) and the Pearson sign
 *     (metrics.py:106-111);
 *   - the matrix targets target_pearson / target_cosine / target_jaccard (metrics.py:414-420,
 *     452-461, 434-449) share the same contraction.
 *
 * Conventions
 *   - every function returns 0 (GV_OK) or a negative gv_status; gv_last_error() gives the text of
 *     the last failure on the calling thread.  Nothing throws, nothing calls exit().
 *   - all data pointers are DEVICE pointers on the current CUDA device unless the name ends in
 *     _host.  The caller owns every buffer (PyTorch tensors are used as plain allocations); the
 *     library allocates nothing that outlives a call.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it.  gv_discretize_csr
 *     synchronises the stream once (it reads two data flags back); the tile functions do not.
 *   - there is no CPU fallback: on a device that is not compute capability 10.0 every compute
 *     entry point fails with GV_ERR_DEVICE.
 */
#ifndef GVB200_H_
#define GVB200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GV_ABI_VERSION 6
#define GV_MAX_BINS 31    /* n_bins of discretize_genes: 1..31 */
#define GV_MARG_STRIDE 32 /* int32 per gene in `marg` */
#define GV_EDGE_STRIDE 32 /* float64 per gene in `edges` (at most 30 edges) */
#define GV_EDGE_SCALE_SLOT 31 /* edges[g][31]: real value of one unit of the gene's value-row integer */
#define GV_EDGE_OFFSET_SLOT 30 /* edges[g][30]: value-row integer of a zero cell (DBL_MAX = 0): only
                                  genes with negative values are stored shifted (val_mode 2) */
#define GV_QTABLE_DIM 32  /* q_table_host is [32][32] float64 */

typedef enum gv_status {
  GV_OK = 0,
  GV_ERR_ARG = -1,         /* bad argument (shape, alignment, range) */
  GV_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed */
  GV_ERR_DEVICE = -3,      /* not an sm_100 device */
  GV_ERR_UNSUPPORTED = -4, /* input outside what the exact integer path covers */
  GV_ERR_TIMEOUT = -5      /* another rank of the box did not arrive in time */
} gv_status;

/* Storage of the CSR arrays.  GV_COUNTS_U8_COL16: `values` are uint8 and `col_idx` uint16 -- count data
 * (integers <= 255, fewer than 65,537 genes) packed on the way up by gv_upload_packed_counts; it takes
 * the count path of the discretisation, with the edge arithmetic of float32 input (what the data were). */
typedef enum gv_dtype { GV_F32 = 0, GV_F64 = 1, GV_COUNTS_U8_COL16 = 2 } gv_dtype;

/* Discretisation strategies (same results, bit for bit):
 *   GV_PATH_COUNTS  dense uint8 transposition through shared memory + per-gene value histogram and
 *                   lookup table; needs integer values <= 255 in canonical CSR (ascending columns)
 *   GV_PATH_GENERAL gene-major regroup + multi-target radix select; any float32/float64 data
 *   GV_PATH_AUTO    try COUNTS, fall back to GENERAL when the data do not qualify */
typedef enum gv_disc_path { GV_PATH_AUTO = 0, GV_PATH_COUNTS = 1, GV_PATH_GENERAL = 2 } gv_disc_path;

/* What gv_discretize_csr writes into the value rows (see there). */
typedef enum gv_value_mode {
  GV_VAL_COUNTS = 0, GV_VAL_DETECTED = 1, GV_VAL_FIXED_POINT = 2, GV_VAL_RANKS = 3
} gv_value_mode;

/* Storage of the one-hot operand of the MI contraction (both give bit-exact joint counts):
 *   GV_OPERAND_INT8  one byte per (row, cell); tcgen05 kind::i8, INT32 accumulators; every n_bins
 *   GV_OPERAND_FP4   4-bit E2M1 (two cells per byte); tcgen05 kind::mxf4.block_scale with unit scales,
 *                    FP32 accumulators (exact below 2^24 cells), half the operand bytes and twice the
 *                    cells per MMA; needs 10 or 5 rows per gene.  The Python engine picks FP4 where it
 *                    applies (the reference's default n_bins = 10 does) and INT8 otherwise. */
typedef enum gv_operand_format { GV_OPERAND_INT8 = 0, GV_OPERAND_FP4 = 1 } gv_operand_format;

/* Score kinds of gv_moment_tiles. */
typedef enum gv_moment_kind {
  GV_MOM_SIGN = 0,    /* int8 sign of the Pearson correlation, np.sign(nan_to_num(corrcoef)):
                         metrics.py:111,134 */
  GV_MOM_PEARSON = 1, /* metrics.py:414-420 */
  GV_MOM_COSINE = 2,  /* metrics.py:452-461 */
  GV_MOM_JACCARD = 3  /* metrics.py:434-449 */
} gv_moment_kind;

int gv_abi_version(void);
const char* gv_last_error(void);
/* GV_OK iff the current device is compute capability 10.0 (B200). */
int gv_device_check(void);

/* ---------------------------------------------------------------- discretisation
 * Replaces discretize_genes(X, n_bins) (metrics.py:25-69) for a CSR matrix [n_cells, n_genes].
 *
 *   values/col_idx/row_ptr : CSR arrays (canonical: no duplicate entries); value_dtype GV_F32/GV_F64
 *   q_table_host[k][t-1]   : np.linspace(0,100,k+1)[1:-1][t-1] / 100 for k = 2..n_bins (host memory;
 *                            numpy's own table so that the quantile fractions match bit for bit)
 *   bins_packed            : out, uint8 [n_genes][bins_pitch_bytes]; value 0..n_bins.  n_bins <= 15: two
 *                            cells per byte, cell c of gene g is nibble (c & 1) of byte c >> 1 (low
 *                            nibble first), 2*bins_pitch_bytes >= padded K extent.  n_bins 16..31: one
 *                            byte per cell, bins_pitch_bytes >= padded K extent.  bins_pitch_bytes % 16 == 0.
 *   n_bins_per_gene        : out, int32 [n_genes]  (second return value of discretize_genes)
 *   marg                   : out, int32 [n_genes][32]; [0] = #cells with x > 0, [a] = #cells in bin a
 *   edges                  : out, float64 [n_genes][32] bin edges (unused slots = DBL_MAX); slot
 *                            GV_EDGE_SCALE_SLOT holds the value-row scale (1.0, or 2^-e_g for val_mode 2)
 *   gsum, gsq              : out, int64 [n_genes]  Sum x and Sum x^2 over cells of the value-row integer
 *   val_rows (optional)    : out, int8 [val_rows_alloc][val_pitch] K-major operand of gv_moment_tiles:
 *                            val_limbs rows per gene holding the base-2^val_limb_bits digits (least
 *                            significant first) of one non-negative integer per cell, in 32-row blocks of
 *                            32/val_limbs genes: row = (g / gpb)*32 + (g % gpb)*val_limbs + digit,
 *                            gpb = 32/val_limbs; val_rows_alloc >= gv_value_rows(n_genes, val_limbs).
 *                            The integer is, by val_mode (gv_value_mode):
 *                              0 the count itself (integer data; exact co-moments);
 *                              1 1 where the gene is detected (one row per gene; Jaccard);
 *                              2 fixed point rint((x - o_g) * 2^e_g), o_g = min(0, min x) (0 unless the
 *                                gene has negative values), e_g per gene such that the largest shifted
 *                                value lands just below 2^(val_limbs*val_limb_bits) (any float data;
 *                                Pearson, its sign and xcorr depend on neither, cosine undoes o_g);
 *                              3 doubled average rank minus one, #{y < x} + #{y <= x} over all cells
 *                                (minus n_zero for genes without negative values, whose zero cells
 *                                then hold 0) (Spearman; needs 2*n_cells < 2^(val_limbs*val_limb_bits)).
 *                            gsum / gsq are then the sums of that integer and of its square.
 *   data_flags_out_host[2] : [0] bit0 = negative value seen, bit1 = fractional value seen, bit3 = NaN or
 *                            infinity seen (value rows are then refused); [1] = max ceil(value)
 *   workspace              : >= gv_discretize_workspace_bytes(nnz, n_cells, n_genes, value_dtype, sorted) bytes,
 *                            sorted = 1 when val_mode is 3 (the rank rows of non-count data need a
 *                            per-gene sort: twice the regroup scratch)
 *   path / path_used_host  : gv_disc_path to use; the one taken is written to *path_used_host (optional)
 * n_bins must be in [1, 31] (GV_MAX_BINS). */
size_t gv_discretize_workspace_bytes(int64_t nnz, int64_t n_cells, int n_genes, int value_dtype, int sorted);
int gv_discretize_csr(const void* values, int value_dtype, const int32_t* col_idx,
                      const int64_t* row_ptr, int64_t n_cells, int n_genes, int64_t nnz, int n_bins,
                      const double* q_table_host, void* bins_packed, int64_t bins_pitch_bytes,
                      int32_t* n_bins_per_gene, int32_t* marg, double* edges, int64_t* gsum,
                      int64_t* gsq, void* val_rows, int64_t val_pitch, int64_t val_rows_alloc,
                      int val_mode, int val_limbs, int val_limb_bits, uint64_t* data_flags_out_host,
                      void* workspace, size_t workspace_bytes, int path, int* path_used_host,
                      void* stream);

/* Is the CSR matrix canonical (scipy's has_canonical_format: per row ascending, unique column indices)
 * with every index in [0, n_genes)?  The discretisation assumes it; the reference gets it from
 * scipy.  Synchronous; *result_host = 0 canonical, 1 unsorted or duplicate entries (canonicalise on the
 * host: sum_duplicates), 2 an index out of range.  scratch_dev: 4 bytes of device memory. */
int gv_csr_check(const void* col_idx, int value_dtype, const int64_t* row_ptr, int64_t n_cells, int n_genes,
                 int* scratch_dev, int* result_host, void* stream);

/* ---------------------------------------------------------------- sharded discretisation front end
 * One process per GPU of one box; count data (integers <= 255, canonical CSR).  Instead of rank 0
 * discretising everything and broadcasting the result, every rank uploads only its own run of 128-cell
 * blocks [blk_begin, blk_end) of the matrix (row_ptr_local: the n_cells_local + 1 row pointers of those
 * cells, starting at 0) and gv_shard_transpose writes their dense uint8 rows raw[gene][cell] into the
 * `raw` buffer of EVERY rank: raw_dests_host holds n_dests device pointers (own buffer, at index
 * own_index, and the gv_ipc_open mappings of the peers'), each [n_genes][raw_pitch], raw_pitch = padded
 * cell count (% 256 == 0).  The rank's slab [n_genes][its cells] is transposed into its own buffer and
 * pushed to every peer by one 2-D peer copy per destination (copy engines over NVLink / NVSwitch).  The block ranges of the ranks must tile [0, raw_pitch / 128).
 * ghist_local (uint32 [n_genes][256]) receives this rank's part of the per-gene value histogram;
 * flags_dev: 64 bytes of device scratch; data_flags_out_host as gv_discretize_csr (the caller combines
 * them over the ranks: OR of [0], max of [1]).  A rank may feed its cells in several chunks (each a run
 * of whole blocks with its own CSR arrays), so that a chunk's transposition and peer copies run under the
 * upload of the next: phase bit 0 marks the first chunk (histogram and flags are zeroed), bit 1 the last
 * (the flags are read back and the stream is synchronised: when that call returns, all of this rank's
 * rows have landed in every destination); phase 3 = everything in one call.
 * After every rank has returned from it (host hand-shake, gv_flag_*), gv_shard_bins sums the ranks'
 * histograms (hists_host: n_hists device pointers, peers' via gv_ipc_open) into ghist_total and bins all
 * genes from this rank's complete `raw`: the same outputs as gv_discretize_csr on the count path.
 * value_rows = 1: `raw` itself is the one-digit INT8 value-row operand (needs every value <= 127 and
 * val_rows_alloc = gv_value_rows(n_genes, 1) rows allocated; rows >= n_genes are zeroed).
 * qtab_dev: 8 KB of device scratch; lut_ws: n_genes * 256 bytes of device scratch (value -> bin tables). */
int gv_shard_transpose(const void* values, int value_dtype, const int32_t* col_idx, const int64_t* row_ptr_local,
                       int64_t n_cells_local, int n_genes, int64_t blk_begin, int64_t blk_end,
                       void* const* raw_dests_host, int n_dests, int own_index, int64_t raw_pitch,
                       uint32_t* ghist_local, void* flags_dev, uint64_t* data_flags_out_host, int phase,
                       void* stream);
int gv_shard_bins(int value_dtype, int64_t n_cells, int n_genes, int n_bins, const double* q_table_host,
                  const void* raw, const void* const* hists_host, int n_hists, uint32_t* ghist_total,
                  double* qtab_dev, void* lut_ws, void* bins_packed, int64_t bins_pitch_bytes, int32_t* n_bins_per_gene,
                  int32_t* marg, double* edges, int64_t* gsum, int64_t* gsq, int value_rows,
                  int64_t val_rows_alloc, int val_limb_bits, void* stream);

/* ---------------------------------------------------------------- gene entropy
 * Replaces GeneVectorDataset.get_gene_entropy (data.py:186-203), which runs before every score
 * computation (data.py:349): per gene the entropy (nats) of the counts of its unique values after
 * dropping the first.  Non-negative integer counts <= 255 take a value-histogram path, anything else
 * (floats, negative values) a per-gene sort of every stored entry (needs row_ptr and nnz < 2^31).  entropy_out float64 [n_genes];
 * workspace >= gv_discretize_workspace_bytes(nnz, n_cells, n_genes, value_dtype, 1) bytes. */
int gv_gene_entropy(const void* values, int value_dtype, const int32_t* col_idx,
                    const int64_t* row_ptr, int64_t nnz, int64_t n_cells, int n_genes,
                    double* entropy_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- operand layout
 * rows_per_gene B for a given n_bins: 5 (n_bins<=5), 8 (<=8), 10 (<=10), 16 (<=15), 32 (<=31).
 * The one-hot operand has 32-row blocks of floor(32/B) genes x B non-zero bins (+ zero rows). */
int gv_rows_per_gene(int n_bins);
/* number of 32-row blocks / operand rows for n_genes */
int64_t gv_onehot_rows(int n_genes, int rows_per_gene);
/* operand rows of the value matrix for n_limbs digits per gene (32-row blocks of 32/n_limbs genes) */
int64_t gv_value_rows(int n_genes, int n_limbs);
/* packed bins (4-bit, or bytes when rows_per_gene is 32) -> one-hot operand rows.  kp = K extent in cells.
 *   GV_OPERAND_INT8: int8 [gv_onehot_rows][kp], kp % 128 == 0
 *   GV_OPERAND_FP4 : uint8 [gv_onehot_rows][kp / 2], kp % 256 == 0 */
int gv_expand_onehot(const void* bins_packed, int64_t bins_pitch_bytes, int n_genes,
                     int rows_per_gene, void* onehot, int64_t kp, int operand_format, void* stream);

/* ---------------------------------------------------------------- tile schedule (host)
 * Upper-triangular tile list over an operand of `op_rows` rows: tiles are tile_m x 256 operand
 * rows, ordered in bands of `band` tile rows walked column by column so that tiles running at the
 * same time share operand panels in L2.  Writes (row0, col0) int32 pairs into out_pairs_host
 * (capacity in pairs) and returns the number of tiles (also when capacity is too small), or a
 * negative status.  A rank takes a contiguous slice of this list. */
int64_t gv_tile_schedule(int64_t op_rows, int tile_m, int band, int32_t* out_pairs_host,
                         int64_t capacity);
/* rows per tile on the A side for a cta_group (1 -> 128, 2 -> 256) */
int gv_tile_m(int cta_group);

/* ---------------------------------------------------------------- contraction + MI epilogue
 * Replaces the per-pair body of compute_mi_pairs (src/lib.rs:40-98) / _compute_all_mi_numba
 * (metrics.py:147-200) for every gene pair covered by `tiles`: joint counts by tcgen05 INT8 MMA,
 * MI in a fused fp32 epilogue.  Writes out[i][j] and out[j][i] (i<j), 0 on the diagonal and for
 * skipped pairs (a gene without positive values: metrics.py:119-120; lib.rs:33).
 *   sgn: optional int8 [n_genes][ld_sgn] from gv_moment_tiles(GV_MOM_SIGN); score = sign * MI
 *        (metrics.py:132-135).
 *   row_bytes: bytes per operand row (= padded cells for INT8, half of them for FP4), % 128 == 0.
 *   sync_ws: optional >= 64 bytes of device scratch.  When given, the persistent CTA pairs run the
 *        tile schedule in lockstep waves (one arrive per pair per wave) so that the tiles in flight
 *        share operand panels out of L2; pass it when the operand is larger than L2, NULL otherwise.
 *        The same buffer must not be used by two launches that can overlap. */
int gv_mi_tiles(const void* onehot, int64_t op_rows, int64_t row_bytes, int rows_per_gene,
                const int32_t* marg, const int8_t* sgn, int64_t ld_sgn, int n_genes,
                const int32_t* tiles, int64_t n_tiles, float* out, int64_t ld_out, int cta_group,
                int operand_format, void* sync_ws, void* stream);

/* ---------------------------------------------------------------- the reference's FFI, shape for shape
 * genevector._rust.compute_mi_pairs(x_disc, n_bins_per_gene, corr_signs) (src/lib.rs:9-15, stub
 * genevector/_rust.pyi:1-5; called at genevector/metrics.py:330): what a maintainer binds who swaps only
 * the Rust module.  HOST arrays in, as the PyO3 function borrows them:
 *   x_disc_host           int32 [n_cells][n_genes], C order: discretize_genes' first return value
 *   n_bins_per_gene_host  int32 [n_genes]: its second (a gene with <= 1 is skipped, lib.rs:33)
 *   corr_signs_host       float32 [n_genes][n_genes] or NULL: the score is multiplied by +1 unless the
 *                         entry is negative (lib.rs:91-94)
 * and a dense float32 [n_genes][n_genes] matrix out: scores_out_host[i][j] = [j][i] = MI of the pair, 0
 * on the diagonal and for the pairs the Rust function does not return; *n_pairs_out_host (optional) =
 * the number of (i, j, mi) triples it would return.  Bin indices must be < 16 (n_bins <= 15).
 * operand_format: GV_OPERAND_INT8, GV_OPERAND_FP4 or -1 (FP4 where it applies).  Synchronous; the device
 * memory it needs (n_cells * n_genes * 4 bytes for x_disc among others) is allocated and freed inside:
 * the caller of this entry has no device allocator. */
int gv_compute_mi_pairs(const int32_t* x_disc_host, int64_t n_cells, int n_genes,
                        const int32_t* n_bins_per_gene_host, const float* corr_signs_host,
                        float* scores_out_host, int64_t* n_pairs_out_host, int operand_format, void* stream);

/* Test-only: raw INT32 accumulators (= joint counts over non-zero bins) of the tiles.
 * active_rows = 32: every operand row is a column of the tile (N = 256); 30: the B side stages only
 * the first 30 rows of every 32-row block (N = 240, what gv_mi_tiles does for 10 or 5 rows per
 * gene) and columns of the other two rows are left untouched.  out == NULL discards the accumulators
 * (the main loop alone, for probing). */
int gv_gram_dump(const void* operand, int64_t op_rows, int64_t row_bytes, const int32_t* tiles,
                 int64_t n_tiles, int32_t* out, int64_t ld_out, int cta_group, int active_rows,
                 int operand_format, void* stream);

/* Diagnostic: the rate (TOP/s, whole device) at which the tensor pipe executes the MI contraction's
 * own tcgen05.mma instruction (M = 256 per CTA pair, N = 240, operands resident in shared memory)
 * when issued back to back with no loads: the live roofline denominator of bench.py.  Synchronous
 * (times `iters` x 4 MMAs per CTA pair with CUDA events on `stream`, best of 3).  mode 0: MMAs only
 * (the peak); 1: a tcgen05.commit after every 4 MMAs, as the main loop issues; 2: plus its fence;
 * 3: plus a wait on a completed mbarrier phase; 4: MMAs only, all into one accumulator; 5: MMAs only,
 * operands taken from six stage buffers in turn as the main loop does; 6: like 5 with pseudo-random
 * operand bytes instead of zeros (switching power); 7: like 5 with one-hot-like operands (~10 % ones). */
int gv_mma_peak(int operand_format, int cta_group, int iters, int mode, double* tops_out_host,
                void* stream);

/* ---------------------------------------------------------------- co-moment targets
 * val_rows from gv_discretize_csr: n_limbs INT8 rows per gene (layout there; one 0/1 row for
 * GV_MOM_JACCARD), op_rows >= gv_value_rows(n_genes, n_limbs); tiles are in operand rows.  kind
 * GV_MOM_SIGN writes int8 `sgn` (both triangles, 0 on the diagonal); the other kinds write float32
 * `out` like gv_mi_tiles.  GV_MOM_PEARSON on rank rows (val_mode 3) is the Spearman correlation
 * (metrics.py:423-431).  The co-moments are exact integers (128-bit covariance arithmetic): requires
 * n_cells * (2^limb_bits - 1)^2 < 2^32, n_limbs in 1..4 and n_cells * 4^(n_limbs*limb_bits) < 2^63.
 * edges: the `edges` array of gv_discretize_csr or NULL; GV_MOM_COSINE reads GV_EDGE_OFFSET_SLOT
 * from it (genes with negative values are stored shifted; cosine is not shift-invariant).
 * sign_zero (GV_MOM_SIGN only): what is written where the correlation is exactly 0 or undefined (a
 * constant gene): 0 = np.sign(nan_to_num(corrcoef)) as the NumPy / Numba / torch tiers use it
 * (metrics.py:111,134,227,302); 1 = the Rust tier's rule, +1 unless negative (src/lib.rs:91-94). */
int gv_moment_tiles(int kind, const void* val_rows, int64_t op_rows, int64_t kp,
                    const int64_t* gsum, const int64_t* gsq, const int32_t* marg, const double* edges,
                    int64_t n_cells,
                    int n_genes, int n_limbs, int limb_bits, const int32_t* tiles, int64_t n_tiles,
                    int8_t* sgn, int64_t ld_sgn, float* out, int64_t ld_out, int cta_group,
                    void* sync_ws, int sign_zero, void* stream);

/* ---------------------------------------------------------------- graph cross-correlation target
 * Replaces target_graph_xcorr (genevector/_graph_targets.py:10-55) with the built-in "mean"
 * aggregation (genevector/_aggregation.py:48-109); graph weights must be non-negative.
 *
 * gv_graph_aggregate: X_agg = (D^-1 W) X (0.5 X + 0.5 that with include_self) on the gene-major value
 *   rows: x_rows / x_limbs / limb_bits / x_scale as produced by gv_discretize_csr (x_scale = the
 *   `edges` array + GV_EDGE_SCALE_SLOT, x_offset = `edges` + GV_EDGE_OFFSET_SLOT, both with
 *   x_scale_stride 16, or NULL for unit scale / no offset); W is the
 *   adjacency in CSR (w_row_ptr int64 [n_cells+1], w_col int32, w_val float64, any non-negative
 *   weights); w_inv_ws: n_cells doubles of scratch.  Out: y_rows (fixed point, y_limbs digits per
 *   gene, op_rows_y >= gv_value_rows(n_genes, y_limbs)), ysum / ysq int64 [n_genes], yscale float64
 *   [n_genes] (real value of one y unit).
 * gv_xcorr_tiles: xcorr[a][b] = X_std[:,a] . X_agg_std[:,b] / n_cells over the full tile rectangle
 *   (gv_tile_schedule_rect), exact integer co-moments, non-symmetric float32 [n_genes][ld_out];
 *   y_limbs must be 3.
 * gv_symmetrize: out = (t + t^T) / 2 with a zero diagonal (_graph_targets.py:51-52). */
int64_t gv_tile_schedule_rect(int64_t rows_a, int64_t rows_b, int tile_m, int32_t* out_pairs_host,
                              int64_t capacity);
int gv_graph_aggregate(const void* x_rows, int64_t op_rows_x, int64_t kp, int x_limbs, int limb_bits,
                       const double* x_scale, const double* x_offset, int64_t x_scale_stride,
                       const int64_t* w_row_ptr,
                       const int32_t* w_col, const double* w_val, double* w_inv_ws, int include_self,
                       int64_t n_cells, int n_genes, void* y_rows, int64_t op_rows_y, int y_limbs,
                       int64_t* ysum, int64_t* ysq, double* yscale, void* stream);
int gv_xcorr_tiles(const void* x_rows, int64_t op_rows_x, int x_limbs, const void* y_rows,
                   int64_t op_rows_y, int y_limbs, int64_t kp, int limb_bits, const int64_t* xsum,
                   const int64_t* xsq, const double* x_scale, int64_t x_scale_stride,
                   const int64_t* ysum, const int64_t* ysq, const double* yscale, int64_t n_cells,
                   int n_genes, const int32_t* tiles, int64_t n_tiles, float* out, int64_t ld_out,
                   int cta_group, void* sync_ws, void* stream);
int gv_symmetrize(const float* t, int64_t ld_t, int n_genes, float* out, int64_t ld_out, void* stream);

/* ---------------------------------------------------------------- score-building fast path
 * Replaces GeneVectorDataset._build_training_tensors (data.py:377-411) for a dense device score
 * matrix: all n_genes*(n_genes-1) off-diagonal ordered pairs in row-major order,
 * xij = float32(round5 ? round(score,5) : score) * c^2, negatives clamped to 0 when
 * clamp_negative (unsigned training, data.py:392-393).  i_idx/j_idx int64, xij float32. */
int gv_build_training_tensors(const float* scores, int64_t ld_scores, int n_genes, float c,
                              int clamp_negative, int round5, int64_t* i_idx, int64_t* j_idx,
                              float* xij, void* stream);

/* ---------------------------------------------------------------- host runtime (no kernels)
 * The reference hands its backends host arrays (scipy CSR in: data.py:171, 319-326; a score structure
 * out: metrics.py:137-140) and runs in one process; these entries move the same data for one process
 * per GPU without a collective.
 *
 * gv_upload_pageable: pageable host memory -> device on `stream` through a ring of pinned staging
 *   buffers filled by n_threads worker threads (0 = pick from the core count).  Returns when the
 *   source has been read (it may be reused); the copies complete in stream order.  The ring (a few
 *   pinned 8 MB slots) outlives the call; gv_runtime_shutdown() frees it.
 * gv_shm_open / gv_shm_close: a POSIX shared-memory segment (`name` starts with '/'), optionally
 *   page-locked for CUDA (cudaHostRegister) in the calling process.  create = 1 replaces a stale
 *   segment of that name.  Every rank of a box maps the same G x G result matrix and DMAs the
 *   rectangles of its own tiles into it (gv_copy_rect_d2h): result assembly needs no reduction.
 * gv_flag_store / gv_flag_load / gv_flag_wait_ge: release / acquire accesses to a uint64 in such a
 *   segment (completion hand-shake between the ranks); the wait returns GV_ERR_TIMEOUT after
 *   timeout_ms (negative: wait forever).
 * gv_copy_rect_d2h: rows x width_bytes rectangle, device (pitch src_pitch_bytes) -> host. */
int gv_upload_pageable(void* dst_dev, const void* src_host, size_t bytes, int n_threads, void* stream);
/* The same copy without a CPU pass over the data: the caller's pages are page-locked in place
 * (cudaHostRegister, in 64 MB chunks, each chunk's DMA running under the next chunk's registration) and
 * read by the copy engine directly.  The source must stay untouched until gv_upload_release(*token_out),
 * which waits for the copies and unlocks the pages (call it once the rest of the work is enqueued). */
/* Count data (float32 values that are integers in [0, 255], column indices < 65,536): the staging pass
 * packs while it copies -- uint8 values and uint16 column indices reach the device, 3 bytes per stored
 * entry instead of 8 (storage GV_COUNTS_U8_COL16).  *ok_host = 0 when an entry does not fit (device
 * contents are then undefined: upload the arrays as they are).  gv_pack_counts_host is the packing pass
 * alone, host to host (destinations 32-byte aligned); it exists for the tests. */
int gv_upload_packed_counts(void* dst_vals_u8, void* dst_cols_u16, const float* src_vals_host,
                            const int32_t* src_cols_host, size_t nnz, int n_threads, int* ok_host, void* stream);
int gv_pack_counts_host(const float* src_vals, const int32_t* src_cols, size_t nnz, uint8_t* dst_vals,
                        uint16_t* dst_cols, int* ok_host);
int gv_upload_registered(void* dst_dev, const void* src_host, size_t bytes, void* stream, void** token_out);
int gv_upload_release(void* token);
int gv_runtime_shutdown(void);
int gv_shm_open(const char* name, size_t bytes, int create, int cuda_register, void** ptr_out_host);
int gv_shm_close(void* ptr_host, size_t bytes, int cuda_registered, const char* unlink_name);
int gv_flag_store(uint64_t* flag_host, uint64_t value);
uint64_t gv_flag_load(const uint64_t* flag_host);
int gv_flag_wait_ge(const uint64_t* flag_host, uint64_t value, int timeout_ms);
int gv_copy_rect_d2h(void* dst_host, size_t dst_pitch_bytes, const void* src_dev, size_t src_pitch_bytes,
                     size_t width_bytes, size_t rows, void* stream);
/* Device allocations other processes of the box can map (CUDA IPC, peer access over NVLink):
 * gv_ipc_alloc cudaMallocs `bytes` and writes the 64-byte handle a peer passes to gv_ipc_open. */
int gv_ipc_alloc(size_t bytes, void** dev_ptr_out, void* handle_out_64_host);
int gv_ipc_open(const void* handle_64_host, void** dev_ptr_out);
int gv_ipc_close(void* dev_ptr);
int gv_ipc_free(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* GVB200_H_ */
