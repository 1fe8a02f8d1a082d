// gv_gram.cuh -- persistent tcgen05 kernel skeleton for C = P * P^T over tiles of the
// upper triangle, where P is a K-major INT8 operand (rows = (gene, bin) one-hot rows for the
// MI path, or one value row per gene for the co-moment path; K = cells).
//
// Replaces the per-pair cell loops of the reference (genevector/metrics.py:167-174 Numba,
// src/lib.rs:47-56 Rust, metrics.py:122-129 NumPy): one tile of INT32 accumulators in TMEM
// holds the joint counts of (TILE_M/B x TILE_N/B) gene pairs at once.
//
// Structure (one CTA per SM, CG CTAs per cluster; CG=2 pairs two SMs on one 256x256 tile):
//   warp 0      TMA producer  : operand tiles global -> shared, 128-byte K slices, SWIZZLE_128B
//   warp 1      MMA issuer    : tcgen05.mma.kind::i8 (leader CTA only), accumulators in TMEM
//   warp 2      TMEM allocator
//   warps 4..11 epilogue      : tcgen05.ld the accumulators and run the Epi functor; TMEM is
//                               double buffered so the epilogue of tile t overlaps the MMA of t+1
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty pair (MMA <-> epilogue),
// static tile schedule (tile list in global memory, cluster c takes tiles c, c+nclusters, ...).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "gv_ptx.cuh"

namespace gv {

constexpr int GRAM_BK = 128;           // bytes (= int8 elements) of K per pipeline stage
constexpr int GRAM_TILE_N = 256;       // accumulator columns per tile
constexpr int GRAM_EPI_WARPS = 8;      // warps 4..11
constexpr int GRAM_THREADS = (4 + GRAM_EPI_WARPS) * 32;
constexpr int GRAM_EPI_BAR = 1;        // named barrier id used by the epilogue warps

constexpr int GRAM_SMEM_MAX = 227 * 1024;   // dynamic shared memory a CTA may opt in to on sm_100

// ACTIVE = operand rows that carry data in every 32-row block (32, or 30 for 3 genes x 10 bins).
// On the B side the data rows of the 8 blocks of a tile are staged back to back (one TMA box of
// ACTIVE rows per block), so the MMA runs with N = 8*ACTIVE columns instead of 256 and no
// accumulator column is spent on padding rows.  The A side keeps whole 32-row blocks: a gene's
// bins must stay inside one 32-lane TMEM quarter for the epilogue.
//
// FP4 = true: the operand holds 4-bit E2M1 one-hot values (two cells per byte) and the MMA is
// kind::mxf4.block_scale with unit UE8M0 scales: a 128-byte K slice then covers 256 cells and every
// MMA 64 of them, at the same bytes per instruction as INT8, i.e. twice the cells per second.
// The scale factors live in the 16 + 16 TMEM columns the N = 240 layout leaves free
// (columns 240..255 and 496..511), so FP4 needs ACTIVE == 30.
template <int CG, int EPI_SMEM_BYTES, int ACTIVE = 32, bool FP4 = false>
struct GramCfg {
  static_assert(ACTIVE == 32 || ACTIVE == 30, "unsupported block occupancy");
  static_assert(!FP4 || ACTIVE == 30, "the FP4 operand needs the 240-column accumulator layout");
  static constexpr uint32_t SFA_COL = 240, SFB_COL = 496;
  static constexpr int TILE_M = 128 * CG;                 // operand rows on the A side per tile
  static constexpr int N_COLS = 8 * ACTIVE;               // accumulator columns = MMA N
  static constexpr int A_ROWS = 128;                      // rows of A each CTA stages
  static constexpr int B_ROWS = N_COLS / CG;              // data rows of B each CTA stages
  static constexpr int B_BLOCKS = 8 / CG;                 // 32-row operand blocks behind them
  static constexpr int A_BYTES = A_ROWS * GRAM_BK;
  static constexpr int B_BYTES = B_ROWS * GRAM_BK;
  static constexpr int B_SLOT_BYTES = (GRAM_TILE_N / CG) * GRAM_BK;   // smem reserved for B
  static constexpr int STAGE_BYTES = A_BYTES + B_SLOT_BYTES;
  static constexpr int TX_BYTES = A_BYTES + B_BYTES;      // bytes TMA delivers per CTA per stage
  static constexpr int BAR_BYTES = 256;                   // barriers + tmem pointer
  // as many pipeline stages as fit next to the epilogue's scratch (1024 = alignment slack)
  static constexpr int STAGES_FIT = (GRAM_SMEM_MAX - 1024 - BAR_BYTES - EPI_SMEM_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static_assert(STAGES >= 3, "not enough shared memory for a 3-stage operand pipeline");
  static_assert(8 * (2 * STAGES + 5) <= BAR_BYTES, "barrier block too small");
  static constexpr uint32_t IDESC = FP4 ? idesc_mxf4(TILE_M, N_COLS) : idesc_i8(TILE_M, N_COLS, true, true);
};

// What the epilogue functor gets for one tile.
struct EpiTile {
  uint32_t tmem_acc;   // TMEM address of this tile's accumulator (lane 0, first column)
  int row0;            // operand row held by TMEM lane 0 of this CTA
  int col0;            // operand row (of the B side) held by accumulator column 0
  int quarter;         // lane quarter this warp may read (warp_idx % 4): lanes 32q..32q+31
  int half;            // which half of the 8 column blocks (of 32 columns) this warp handles
  int lane;            // lane id
  int epi_tid;         // 0..255 within the epilogue warps
};

struct GramArgs {
  const int2* tiles;   // (row0, col0) per tile, in operand rows; row0 % TILE_M == 0, col0 % 256 == 0
  int n_tiles;
  int k_blocks;        // operand K extent / 128
  // Wave lockstep (optional, zero-initialised device counter): every cluster arrives once per
  // wave of n_clusters tiles before loading the wave's tile, so tiles in flight walk K together
  // and share operand panels out of L2 (without it the clusters drift apart and at 20k x 200k
  // the kernel re-reads the operand from DRAM 380 times: profiles/r01_ncu_mi_c4_v1_nolockstep.txt).
  unsigned int* wave_sync;
};

template <int CG, class Epi>
constexpr size_t gram_smem_bytes() {
  using Cfg = GramCfg<CG, int(Epi::SMEM_BYTES), Epi::ACTIVE_ROWS, Epi::OPERAND_FP4>;
  return 1024 + size_t(Cfg::STAGES) * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + Epi::SMEM_BYTES;
}

template <int CG, class Epi>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_tiles_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_b,
                  const GramArgs ga, const typename Epi::Params ep) {
  // tmap: 2-D, 128-row boxes over the A-side operand;
  // tmap_b: the B-side operand (the same matrix for the symmetric products, another one for the
  // cross-correlation target): 2-D 128-row boxes when ACTIVE == 32, else a 3-D view
  // [block][32 rows][K] with a box of B_BLOCKS blocks x ACTIVE rows x 128 bytes
  using Cfg = GramCfg<CG, int(Epi::SMEM_BYTES), Epi::ACTIVE_ROWS, Epi::OPERAND_FP4>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t bars = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));
  uint8_t* epi_smem = smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / CG;
  const int n_clusters = gridDim.x / CG;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap); tma_prefetch_desc(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), CG);   // one producer arrive per CTA of the pair (on the leader)
      mbar_init(empty_bar(s), 1);   // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);                       // one tcgen05.commit
      mbar_init(tempty_bar(a), CG * GRAM_EPI_WARPS);    // every epilogue warp of the pair
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  if constexpr (Epi::OPERAND_FP4) {
    // unit scale factors (UE8M0 0x7F = 2^0) for every row of A and B, in every lane quarter
    if (warp >= 4 && warp < 8) {
      const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
      tmem_fill_32x16(tmem_base + lane_base + Cfg::SFA_COL, 0x7F7F7F7Fu);
      tmem_fill_32x16(tmem_base + lane_base + Cfg::SFB_COL, 0x7F7F7F7Fu);
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
  }

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The warp runs converged; one elected lane issues the loads and the barrier arrives (see the MMA
    // warp below for why not a single-lane branch).
    {
      const bool elected = elect_one();
      const uint32_t full0 = (CG == 2) ? mapa_u32(full_bar(0), 0) : full_bar(0);
      int s = 0; uint32_t ph = 0;
      unsigned int wave = 0;
      const int n_waves = (ga.n_tiles + n_clusters - 1) / n_clusters;
      for (int t = cluster_id; int(wave) < n_waves; t += n_clusters, ++wave) {
        if (ga.wave_sync != nullptr && leader) {
          // the peer's loads only fill its own stages; MMA cannot start before the leader's arrive.
          // wave_sync[0] counts arrivals; wave_sync[1] != 0 means the lockstep was given up: a cluster
          // that waited a full second for the others (they are not co-resident: something else holds
          // SMs) sets it and every cluster runs on unsynchronised -- slower, never a hang or a trap
          volatile unsigned int* ws = reinterpret_cast<volatile unsigned int*>(ga.wave_sync);
          if (elected) atomicAdd(ga.wave_sync, 1u);
          __syncwarp();
          // every lane polls (one broadcast load per poll): the warp stays converged
          const unsigned int target = unsigned(n_clusters) * (wave + 1u);
          if (ws[1] == 0u && ws[0] < target) {
            const uint64_t t0 = global_timer_ns();
            while (ws[0] < target && ws[1] == 0u) {
              __nanosleep(128);
              if (global_timer_ns() - t0 > 1000000000ull) ws[1] = 1u;
            }
          }
          __syncwarp();
        }
        if (t >= ga.n_tiles) break;
        const int2 tc = ga.tiles[t];
        const int arow = tc.x + int(rank) * Cfg::A_ROWS;
        const int brow = tc.y + int(rank) * (GRAM_TILE_N / CG);   // first operand row of my B blocks
        // odd waves walk K backwards: the K slices a wave touched last are still in L2 when the next
        // wave (same band: same A panels) starts with them; the exact sums do not depend on the order
        const bool backwards = (wave & 1u) != 0u;
        for (int kq = 0; kq < ga.k_blocks; ++kq) {
          const int kb = backwards ? ga.k_blocks - 1 - kq : kq;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint32_t fb = full0 + 8u * s;   // leader's barrier (cluster address if CG==2)
          if (elected) {
            if constexpr (CG == 2) {
              tma_load_2d_pair(&tmap, fb, sa, kb * GRAM_BK, arow);
              if constexpr (Epi::ACTIVE_ROWS == 32) {
                tma_load_2d_pair(&tmap_b, fb, sb, kb * GRAM_BK, brow);
              } else {
                // one 3-D box: 128 bytes of K x ACTIVE rows x B_BLOCKS blocks, landing back to back
                tma_load_3d_pair(&tmap_b, fb, sb, kb * GRAM_BK, 0, brow >> 5);
              }
              if (leader) mbar_arrive_expect_tx(full_bar(s), Cfg::TX_BYTES * CG);
              else        mbar_arrive_cluster(fb);
            } else {
              tma_load_2d(&tmap, fb, sa, kb * GRAM_BK, arow);
              if constexpr (Epi::ACTIVE_ROWS == 32) {
                // TMA boxes are at most 256 rows; 256 rows here and the box is 128 rows
                tma_load_2d(&tmap_b, fb, sb, kb * GRAM_BK, brow);
                tma_load_2d(&tmap_b, fb, sb + Cfg::A_BYTES, kb * GRAM_BK, brow + 128);
              } else {
                tma_load_3d(&tmap_b, fb, sb, kb * GRAM_BK, 0, brow >> 5);
              }
              mbar_arrive_expect_tx(full_bar(s), Cfg::TX_BYTES);
            }
          }
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // One thread issues every tcgen05.mma of the CTA pair, and ncu's source page shows that this thread,
    // not the operand supply, bounds the streaming loop (profiles/r02_ncu_mi_c4_fp4_stalls.txt: 115 SASS
    // instructions per stage, 2.5 % of the warp's samples on each of them, 12 % waiting for data).  So the
    // loop is unrolled over the pipeline stages: with the stage a compile-time constant, barrier
    // addresses and shared-memory descriptors are immediates off uniform registers, nothing is
    // recomputed per stage, and the first MMA of a tile is the only one that carries a predicate.
    // The stage counter keeps running across tiles (k_blocks need not be a multiple of STAGES): a
    // switch enters the unrolled body at the current stage (Duff's device).  The whole warp runs the
    // loop converged and one elected lane issues: inside a single-lane branch the compiler keeps the
    // operands in vector registers and moves them to uniform ones around every instruction.
    if (leader) {
      const bool elected = elect_one();
      int s = 0; uint32_t ph = 0; int it = 0;
      const uint64_t desc0 = smem_desc_k128(smem_base);                 // A tile of stage 0
      constexpr uint64_t STAGE_D = uint64_t(Cfg::STAGE_BYTES >> 4);     // in 16-byte address units
      constexpr uint64_t A_D = uint64_t(Cfg::A_BYTES >> 4);
      const uint32_t sfa = tmem_base + Cfg::SFA_COL, sfb = tmem_base + Cfg::SFB_COL;
      (void)sfa; (void)sfb;
      const int K = ga.k_blocks;
      for (int t = cluster_id; t < ga.n_tiles; t += n_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc) * GRAM_TILE_N;
        int kb = 0;
#define GV_MMA_ONE(AD, BD, ACCUM)                                                                    \
        do {                                                                                         \
          if constexpr (Epi::OPERAND_FP4) mma_mxf4<CG>(d_tmem, (AD), (BD), Cfg::IDESC, sfa, sfb, (ACCUM)); \
          else mma_i8<CG>(d_tmem, (AD), (BD), Cfg::IDESC, (ACCUM));                                  \
        } while (0)
#define GV_MMA_STAGE(S)                                                                              \
        case (S):                                                                                    \
          if constexpr ((S) < Cfg::STAGES) {                                                         \
            mbar_wait(bars + 8u * (S), ph);                                                          \
            tc_fence_after();                                                                        \
            const uint64_t adesc = desc0 + uint64_t(S) * STAGE_D;                                    \
            const uint64_t bdesc = adesc + A_D;                                                      \
            /* +2 in the 16-byte-granular start-address field = 32 bytes = one K step of the swizzle row */ \
            if (elected) {                                                                           \
              GV_MMA_ONE(adesc, bdesc, kb != 0 ? 1u : 0u);                                           \
              GV_MMA_ONE(adesc + 2, bdesc + 2, 1u);                                                  \
              GV_MMA_ONE(adesc + 4, bdesc + 4, 1u);                                                  \
              GV_MMA_ONE(adesc + 6, bdesc + 6, 1u);                                                  \
              mma_commit<CG>(bars + 8u * (Cfg::STAGES + (S)));    /* frees the smem slot (both CTAs) */ \
            }                                                                                        \
            if constexpr ((S) + 1 == Cfg::STAGES) ph ^= 1u;                                          \
            if (++kb == K) { s = ((S) + 1) % Cfg::STAGES; goto tile_done; }                          \
          }
        static_assert(GRAM_BK / 32 == 4, "four K steps per stage");
        switch (s) {
          for (;;) {
            GV_MMA_STAGE(0) GV_MMA_STAGE(1) GV_MMA_STAGE(2) GV_MMA_STAGE(3)
            GV_MMA_STAGE(4) GV_MMA_STAGE(5) GV_MMA_STAGE(6) GV_MMA_STAGE(7)
          }
        }
      tile_done:
        if (elected) mma_commit<CG>(tfull_bar(acc));             // accumulator ready
#undef GV_MMA_STAGE
#undef GV_MMA_ONE
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    EpiTile et;
    et.quarter = warp & 3;
    et.half = ew >> 2;
    et.lane = lane;
    et.epi_tid = ew * 32 + lane;
    const uint32_t tempty0 = (CG == 2) ? mapa_u32(tempty_bar(0), 0) : tempty_bar(0);
    Epi epi(ep, epi_smem);
    int it = 0;
    for (int t = cluster_id; t < ga.n_tiles; t += n_clusters, ++it) {
      const int2 tc = ga.tiles[t];
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1u;
      et.row0 = tc.x + int(rank) * 128;
      et.col0 = tc.y;
      et.tmem_acc = tmem_base + uint32_t(acc) * GRAM_TILE_N;
      epi.prepare(et);                       // tile metadata -> smem (overlaps the MMA)
      // one warp waits for the accumulator; the other seven sleep on the named barrier instead of
      // polling the mbarrier (they woke on every barrier event of the SM: 2.5e10 polls per C4 launch)
      if (ew == 0) mbar_wait(tfull_bar(acc), acc_ph);
      asm volatile("bar.sync %0, %1;" ::"n"(GRAM_EPI_BAR), "n"(GRAM_EPI_WARPS * 32) : "memory");
      tc_fence_after();
      epi.run(et);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + 8u * acc);
    }
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

}  // namespace gv
