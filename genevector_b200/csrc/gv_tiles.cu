// gv_tiles.cu -- host launchers for the tcgen05 tile kernels (gv_gram.cuh + gv_epilogues.cuh):
// TMA tensor map of the K-major INT8 operand, cluster launch, template dispatch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "gv_internal.h"
#include "gv_epilogues.cuh"

namespace gv {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D map over operand[op_rows][kp] (uint8), box = 128 bytes of K x box_rows rows, SWIZZLE_128B.
static int make_operand_map(CUtensorMap* map, const void* operand, int64_t op_rows, int64_t kp,
                            int box_rows = 128) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(GV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (kp % 128 != 0 || kp <= 0) return set_error(GV_ERR_ARG, "operand K extent must be a positive multiple of 128");
  if ((reinterpret_cast<uintptr_t>(operand) & 127) != 0)
    return set_error(GV_ERR_ARG, "operand must be 128-byte aligned");
  cuuint64_t dims[2] = {cuuint64_t(kp), cuuint64_t(op_rows)};
  cuuint64_t strides[1] = {cuuint64_t(kp)};
  cuuint32_t box[2] = {128, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(operand), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(GV_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return GV_OK;
}

// 3-D map over the same operand seen as [op_rows/32 blocks][32 rows][kp bytes]; one box fetches
// the first `active` rows of `blocks` consecutive 32-row blocks (they land contiguously in smem).
static int make_block_map(CUtensorMap* map, const void* operand, int64_t op_rows, int64_t kp,
                          int active, int blocks) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(GV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (op_rows % 32 != 0) return set_error(GV_ERR_ARG, "operand rows must be a multiple of 32");
  cuuint64_t dims[3] = {cuuint64_t(kp), 32, cuuint64_t(op_rows / 32)};
  cuuint64_t strides[2] = {cuuint64_t(kp), cuuint64_t(kp) * 32};
  cuuint32_t box[3] = {128, cuuint32_t(active), cuuint32_t(blocks)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(operand), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(GV_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed");
  return GV_OK;
}

// operand_b / op_rows_b: the B-side matrix when it is not the A-side one (cross-correlation target)
template <int CG, class Epi>
static int launch_gram(const void* operand, int64_t op_rows, int64_t kp, const int32_t* tiles,
                       int64_t n_tiles, const typename Epi::Params& ep, void* sync_ws,
                       cudaStream_t st, const void* operand_b = nullptr, int64_t op_rows_b = 0) {
  if (n_tiles <= 0) return GV_OK;
  if (operand_b == nullptr) { operand_b = operand; op_rows_b = op_rows; }
  CUtensorMap map, map_b;
  int rc = make_operand_map(&map, operand, op_rows, kp);
  if (rc != GV_OK) return rc;
  if (Epi::ACTIVE_ROWS == 32) rc = make_operand_map(&map_b, operand_b, op_rows_b, kp, 128);
  else rc = make_block_map(&map_b, operand_b, op_rows_b, kp, Epi::ACTIVE_ROWS, 8 / CG);
  if (rc != GV_OK) return rc;
  GramArgs ga;
  ga.tiles = reinterpret_cast<const int2*>(tiles);
  ga.n_tiles = int(n_tiles);
  ga.k_blocks = int(kp / GRAM_BK);
  ga.wave_sync = static_cast<unsigned int*>(sync_ws);
  if (sync_ws != nullptr) {
    cudaError_t e0 = cudaMemsetAsync(sync_ws, 0, 64, st);
    if (e0 != cudaSuccess) return set_cuda_error(e0, __FILE__, __LINE__);
  }
  auto kern = gram_tiles_kernel<CG, Epi>;
  const size_t smem = gram_smem_bytes<CG, Epi>();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  const int sms = device_sm_count();
  int64_t clusters = sms / CG;
  if (clusters > n_tiles) clusters = n_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(unsigned(clusters * CG));
  cfg.blockDim = dim3(GRAM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // The persistent grid is one CTA per SM; never more clusters than the device can hold at once
  // (GPC floorsweeping, a smaller part): the wave barrier counts on every cluster being resident.
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 &&
      clusters > max_clusters) {
    clusters = max_clusters;
    cfg.gridDim = dim3(unsigned(clusters * CG));
  } else {
    cudaGetLastError();
  }
  if (sync_ws != nullptr) {
    // Wave lockstep: a cooperative launch, so that SMs held by anything else (another stream, an MPS
    // client, NCCL) give a launch error here instead of clusters spinning on a barrier that cannot
    // complete; the launch then falls back to the plain one, whose barrier gives itself up after a
    // second (gv_gram.cuh).  GV_COOPERATIVE=0 skips the attempt (Nsight Compute cannot replay
    // cooperative cluster launches).
    static const bool coop = !(getenv("GV_COOPERATIVE") != nullptr && getenv("GV_COOPERATIVE")[0] == '0');
    if (coop) {
      cfg.numAttrs = 2;
      e = cudaLaunchKernelEx(&cfg, kern, map, map_b, ga, ep);
      if (e == cudaSuccess) return GV_OK;
      cudaGetLastError();                    // not launched: clear and retry as a plain launch
      cfg.numAttrs = 1;
    }
  }
  e = cudaLaunchKernelEx(&cfg, kern, map, map_b, ga, ep);
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  return GV_OK;
}

template <class Epi>
static int launch_cg(int cta_group, const void* operand, int64_t op_rows, int64_t kp,
                     const int32_t* tiles, int64_t n_tiles, const typename Epi::Params& ep,
                     void* sync_ws, cudaStream_t st, const void* operand_b = nullptr,
                     int64_t op_rows_b = 0) {
  if (cta_group == 2)
    return launch_gram<2, Epi>(operand, op_rows, kp, tiles, n_tiles, ep, sync_ws, st, operand_b, op_rows_b);
  if (cta_group == 1)
    return launch_gram<1, Epi>(operand, op_rows, kp, tiles, n_tiles, ep, sync_ws, st, operand_b, op_rows_b);
  return set_error(GV_ERR_ARG, "cta_group must be 1 or 2");
}

int mi_tiles_dispatch(const void* onehot, int64_t op_rows, int64_t kp, int rows_per_gene,
                      const int32_t* marg, const int8_t* sgn, int64_t ld_sgn, int n_genes,
                      const int32_t* tiles, int64_t n_tiles, float* out, int64_t ld_out,
                      int cta_group, int operand_format, void* sync_ws, cudaStream_t st) {
  MiParams p;
  p.marg = marg; p.sgn = sgn; p.ld_sgn = ld_sgn; p.out = out; p.ld_out = ld_out; p.G = n_genes;
  if (operand_format == GV_OPERAND_FP4) {
    switch (rows_per_gene) {
      case 5:  return launch_cg<MiEpi<5, true>>(cta_group, onehot, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
      case 10: return launch_cg<MiEpi<10, true>>(cta_group, onehot, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    }
    return set_error(GV_ERR_ARG, "the FP4 operand needs rows_per_gene 5 or 10 (n_bins <= 5 or 9..10)");
  }
  if (operand_format != GV_OPERAND_INT8) return set_error(GV_ERR_ARG, "unknown operand format");
  switch (rows_per_gene) {
    case 5:  return launch_cg<MiEpi<5>>(cta_group, onehot, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case 8:  return launch_cg<MiEpi<8>>(cta_group, onehot, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case 10: return launch_cg<MiEpi<10>>(cta_group, onehot, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case 16: return launch_cg<MiEpi<16>>(cta_group, onehot, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case 32: return launch_cg<MiEpi<32>>(cta_group, onehot, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
  }
  return set_error(GV_ERR_ARG, "rows_per_gene must be 5, 8, 10, 16 or 32");
}

int gram_dump_dispatch(const void* operand, int64_t op_rows, int64_t kp, const int32_t* tiles,
                       int64_t n_tiles, int32_t* out, int64_t ld_out, int cta_group, int active_rows,
                       int operand_format, cudaStream_t st) {
  DumpParams p;
  p.out = out; p.ld = ld_out; p.rows = int(op_rows);
  if (operand_format == GV_OPERAND_FP4) {
    if (active_rows != 30) return set_error(GV_ERR_ARG, "gram dump (fp4): active_rows must be 30");
    return launch_cg<DumpEpi<30, true>>(cta_group, operand, op_rows, kp, tiles, n_tiles, p, nullptr, st);
  }
  if (active_rows == 32)
    return launch_cg<DumpEpi<32>>(cta_group, operand, op_rows, kp, tiles, n_tiles, p, nullptr, st);
  if (active_rows == 30)
    return launch_cg<DumpEpi<30>>(cta_group, operand, op_rows, kp, tiles, n_tiles, p, nullptr, st);
  return set_error(GV_ERR_ARG, "gram dump: active_rows must be 32 or 30");
}

template <int L>
static int moment_dispatch_l(int kind, int cta_group, const void* val_rows, int64_t op_rows, int64_t kp,
                             const int32_t* tiles, int64_t n_tiles, const MomentParams& p,
                             void* sync_ws, cudaStream_t st) {
  switch (kind) {
    case GV_MOM_SIGN:
      return launch_cg<MomentEpi<MOM_SIGN, L>>(cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case GV_MOM_PEARSON:
      return launch_cg<MomentEpi<MOM_PEARSON, L>>(cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case GV_MOM_COSINE:
      return launch_cg<MomentEpi<MOM_COSINE, L>>(cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case GV_MOM_JACCARD:
      return launch_cg<MomentEpi<MOM_JACCARD, L>>(cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
  }
  return set_error(GV_ERR_ARG, "unknown moment kind");
}

int moment_tiles_dispatch(int kind, const void* val_rows, int64_t op_rows, int64_t kp,
                          const int64_t* gsum, const int64_t* gsq, const int32_t* marg,
                          const double* edges, int64_t n_cells, int n_genes, int n_limbs, int limb_bits,
                          const int32_t* tiles, int64_t n_tiles, int8_t* sgn, int64_t ld_sgn,
                          float* out, int64_t ld_out, int cta_group, void* sync_ws, int sign_zero,
                          cudaStream_t st) {
  MomentParams p;
  p.sign_zero = sign_zero;
  p.gsum = gsum; p.gsq = gsq; p.marg = marg; p.edges = edges; p.sgn = sgn; p.ld_sgn = ld_sgn; p.out = out;
  p.ld_out = ld_out; p.G = n_genes; p.n_cells = n_cells; p.limb_bits = limb_bits;
  if (kind == GV_MOM_SIGN && (!sgn || !gsum || !gsq)) return set_error(GV_ERR_ARG, "sign needs sgn, gsum, gsq");
  if ((kind == GV_MOM_PEARSON || kind == GV_MOM_COSINE) && (!out || !gsum || !gsq))
    return set_error(GV_ERR_ARG, "pearson / cosine need out, gsum, gsq");
  if (kind == GV_MOM_JACCARD && (!out || !marg || n_limbs != 1))
    return set_error(GV_ERR_ARG, "jaccard needs out, marg and one detection row per gene");
  if (limb_bits < 1 || limb_bits > 7) return set_error(GV_ERR_ARG, "limb_bits must be in [1,7]");
  if (n_limbs < 1 || n_limbs > 4)
    return set_error(GV_ERR_UNSUPPORTED, "co-moment path supports one to four digits per gene");
  // exactness: every digit product summed over all cells must fit the (unsigned) 32-bit accumulator,
  // and the recombined Sum x_i x_j (values < 2^(n_limbs*limb_bits)) the signed 64-bit sums
  const uint64_t dmax = (1ull << limb_bits) - 1ull;
  if (uint64_t(n_cells) * dmax * dmax >= (1ull << 32))
    return set_error(GV_ERR_UNSUPPORTED, "co-moment accumulators would overflow: lower limb_bits");
  const int tb = n_limbs * limb_bits;
  if (tb > 31 || (long double)n_cells * (long double)((1ull << tb) - 1ull) * (long double)((1ull << tb) - 1ull) >=
                     9.2e18L)
    return set_error(GV_ERR_UNSUPPORTED, "co-moment sums would overflow 64 bits: fewer value bits needed");
  switch (n_limbs) {
    case 1: return moment_dispatch_l<1>(kind, cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case 2: return moment_dispatch_l<2>(kind, cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    case 3: return moment_dispatch_l<3>(kind, cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
    default: return moment_dispatch_l<4>(kind, cta_group, val_rows, op_rows, kp, tiles, n_tiles, p, sync_ws, st);
  }
}

int xcorr_tiles_dispatch(const void* x_rows, int64_t op_rows_x, int x_limbs, const void* y_rows,
                         int64_t op_rows_y, int y_limbs, int64_t kp, int limb_bits, const int64_t* xsum,
                         const int64_t* xsq, const double* xscale, int64_t xscale_stride,
                         const int64_t* ysum, const int64_t* ysq, const double* yscale, int64_t n_cells,
                         int n_genes, const int32_t* tiles, int64_t n_tiles, float* out, int64_t ld_out,
                         int cta_group, void* sync_ws, cudaStream_t st) {
  XcorrParams p;
  p.asum = xsum; p.asq = xsq; p.ascale = xscale; p.ascale_stride = xscale_stride;
  p.bsum = ysum; p.bsq = ysq; p.bscale = yscale; p.out = out; p.ld_out = ld_out; p.G = n_genes;
  p.n_cells = n_cells; p.limb_bits = limb_bits;
  if (limb_bits < 1 || limb_bits > 7) return set_error(GV_ERR_ARG, "limb_bits must be in [1,7]");
  const uint64_t dmax = (1ull << limb_bits) - 1ull;
  if (uint64_t(n_cells) * dmax * dmax >= (1ull << 32))
    return set_error(GV_ERR_UNSUPPORTED, "co-moment accumulators would overflow: lower limb_bits");
  const int tb = (x_limbs > y_limbs ? x_limbs : y_limbs) * limb_bits;
  if (tb > 31 || (long double)n_cells * (long double)((1ull << tb) - 1ull) * (long double)((1ull << tb) - 1ull) >= 9.2e18L)
    return set_error(GV_ERR_UNSUPPORTED, "co-moment sums would overflow 64 bits: fewer value bits needed");
  if (y_limbs != 3) return set_error(GV_ERR_UNSUPPORTED, "xcorr: the aggregated rows must have three digits");
  switch (x_limbs) {
    case 1: return launch_cg<XcorrEpi<1, 3>>(cta_group, x_rows, op_rows_x, kp, tiles, n_tiles, p, sync_ws, st, y_rows, op_rows_y);
    case 2: return launch_cg<XcorrEpi<2, 3>>(cta_group, x_rows, op_rows_x, kp, tiles, n_tiles, p, sync_ws, st, y_rows, op_rows_y);
    case 3: return launch_cg<XcorrEpi<3, 3>>(cta_group, x_rows, op_rows_x, kp, tiles, n_tiles, p, sync_ws, st, y_rows, op_rows_y);
  }
  return set_error(GV_ERR_UNSUPPORTED, "xcorr: the expression rows must have one to three digits");
}

// ------------------------------------------------------------------ instruction-rate probe
// Issues the MI contraction's own tcgen05.mma (same instruction descriptor, M = 256, N = 240,
// one-hot-like zero operands already in shared memory) back to back with no loads in between: the
// rate the tensor pipe sustains when nothing else is in the way, i.e. the live roofline denominator
// bench.py divides by.  Diagnostic only; not on the product path.
template <int CG, bool FP4>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, int mode) {
  using Cfg = GramCfg<CG, 0, 30, FP4>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int PK_STAGES = 6;            // mode 5 walks through as many stage buffers as the main loop
  const uint32_t bar = smem_base + PK_STAGES * Cfg::STAGE_BYTES;
  const uint32_t tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + PK_STAGES * Cfg::STAGE_BYTES + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = (CG == 2) ? cluster_ctarank() == 0 : true;
  // operand bytes: zeros (modes 0-5), pseudo-random (mode 6: every bit toggles, worst-case switching
  // power) or one-hot-like (mode 7: one E2M1 1.0 / INT8 1 in ~10 % of the elements, like the MI operand)
  for (int i = threadIdx.x; i < PK_STAGES * Cfg::STAGE_BYTES / 16; i += blockDim.x) {
    uint4 w = make_uint4(0, 0, 0, 0);
    if (mode >= 6) {
      uint32_t h[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t x = (uint32_t(i) * 4u + q) * 2654435761u + blockIdx.x * 40503u + 12345u;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
        if (mode == 6) {
          h[q] = FP4 ? (x & 0x77777777u) : (x & 0x7F7F7F7Fu);          // finite, non-negative values
        } else {
          // keep an element with probability ~1/10: value 1.0 (E2M1 0x2) or 1 (INT8)
          uint32_t o = 0;
          if (FP4) { for (int n = 0; n < 8; ++n) if (((x >> (4 * n)) & 0xF) == 0 || (((x >> (4 * n)) & 0xF) == 1 && (n & 1))) o |= 0x2u << (4 * n); }
          else     { for (int n = 0; n < 4; ++n) if (((x >> (8 * n)) & 0xFF) < 26) o |= 0x1u << (8 * n); }
          h[q] = o;
        }
      }
      w = make_uint4(h[0], h[1], h[2], h[3]);
    }
    reinterpret_cast<uint4*>(smem_gen)[i] = w;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(bar + 16, 1); mbar_init(bar + 32, 1);
    fence_mbar_init();
    mbar_arrive_local(bar + 32);          // phase 0 of this one is complete from the start
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 2) tmem_alloc<CG>(tmem_slot, 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  if constexpr (FP4) {
    const uint32_t lane_base = uint32_t(warp * 32) << 16;
    tmem_fill_32x16(tmem_base + lane_base + Cfg::SFA_COL, 0x7F7F7F7Fu);
    tmem_fill_32x16(tmem_base + lane_base + Cfg::SFB_COL, 0x7F7F7F7Fu);
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
  }
  if (warp == 1 && lane == 0 && leader) {
    int stage = 0;
    for (int it = 0; it < iters; ++it) {
      // mode 5: a different stage buffer every 4 MMAs (operands streamed, never re-read back to back)
      const uint32_t sa = smem_base + (mode >= 5 ? stage : 0) * Cfg::STAGE_BYTES;
      if (++stage == PK_STAGES) stage = 0;
      const uint64_t adesc = smem_desc_k128(sa);
      const uint64_t bdesc = smem_desc_k128(sa + Cfg::A_BYTES);
      // mode 4: one accumulator throughout (every MMA depends on the previous one, as inside a tile)
      const uint32_t d_tmem = tmem_base + (mode == 4 ? 0u : uint32_t(it & 1) * GRAM_TILE_N);
#pragma unroll
      for (int k = 0; k < GRAM_BK / 32; ++k) {
        if constexpr (FP4)
          mma_mxf4<CG>(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), Cfg::IDESC,
                       tmem_base + Cfg::SFA_COL, tmem_base + Cfg::SFB_COL, 1u);
        else
          mma_i8<CG>(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), Cfg::IDESC, 1u);
      }
      // mode 1: a commit after every 4 MMAs like the real main loop (to a barrier nobody waits on);
      // mode 2: plus the fence the real loop executes after its full-barrier wait
      if (mode >= 1 && mode <= 3) mma_commit<CG>(bar + 16);
      if (mode >= 2 && mode <= 3) tc_fence_after();
      // mode 3: plus a wait on an already completed barrier phase (the shared-memory round trip of
      // the main loop's full-barrier wait)
      if (mode == 3) mbar_wait(bar + 32, 0);
    }
    mma_commit<CG>(bar);
    mbar_wait(bar, 0);
  }
  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

template <int CG, bool FP4>
static int mma_peak_launch(int iters, int mode, double* tops_out, cudaStream_t st) {
  using Cfg = GramCfg<CG, 0, 30, FP4>;
  auto kern = mma_peak_kernel<CG, FP4>;
  const size_t smem = 1024 + 6 * Cfg::STAGE_BYTES + 64;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  const int sms = device_sm_count();
  const int clusters = sms / CG;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(unsigned(clusters * CG));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  if ((e = cudaEventCreate(&e0)) != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return set_cuda_error(e, __FILE__, __LINE__); }
  float best = 1e30f;
  for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {       // first launch is the warm-up
    cudaEventRecord(e0, st);
    e = cudaLaunchKernelEx(&cfg, kern, iters, mode);
    cudaEventRecord(e1, st);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
  // cells per MMA: 32 bytes of K = 32 (INT8) or 64 (FP4)
  const double ops = double(clusters) * double(iters) * (GRAM_BK / 32) * 2.0 * Cfg::TILE_M * Cfg::N_COLS *
                     (FP4 ? 64.0 : 32.0);
  *tops_out = ops / (double(best) * 1e-3) / 1e12;
  return GV_OK;
}

int mma_peak_dispatch(int operand_format, int cta_group, int iters, int mode, double* tops_out,
                      cudaStream_t st) {
  if (!tops_out || iters < 1) return set_error(GV_ERR_ARG, "gv_mma_peak: bad argument");
  const bool fp4 = operand_format == GV_OPERAND_FP4;
  if (!fp4 && operand_format != GV_OPERAND_INT8) return set_error(GV_ERR_ARG, "unknown operand format");
  if (cta_group == 2)
    return fp4 ? mma_peak_launch<2, true>(iters, mode, tops_out, st) : mma_peak_launch<2, false>(iters, mode, tops_out, st);
  if (cta_group == 1)
    return fp4 ? mma_peak_launch<1, true>(iters, mode, tops_out, st) : mma_peak_launch<1, false>(iters, mode, tops_out, st);
  return set_error(GV_ERR_ARG, "cta_group must be 1 or 2");
}

}  // namespace gv
