// gv_host.cu -- host-side runtime of libgvb200.so: the pieces around the kernels that move data
// between the caller's host buffers and the device without a collective.
//
//   gv_upload_pageable   pageable host array -> device through a ring of pinned staging buffers that
//                        worker threads fill in parallel (the reference hands over scipy CSR arrays,
//                        data.py:171; a plain cudaMemcpy of pageable memory runs at a few GB/s)
//   gv_shm_*             one POSIX shared-memory segment, page-locked for CUDA in every process of
//                        the box: the G x G result matrix every rank DMAs its own tiles into
//                        (north_star: "each rank writes its own result tiles", no reduction)
//   gv_flag_*            release / acquire flags in such a segment (completion hand-shake, no NCCL)
//   gv_copy_rect_d2h     one rectangle of the device result matrix -> the host matrix
//   gv_ipc_*             device allocations mapped into the peer processes of the box (NVLink /
//                        NVSwitch peer stores of the sharded discretisation front)
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <immintrin.h>
#include <cstdlib>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "gv_internal.h"

namespace gv {

#define GV_HOST_TRY(expr)                                                 \
  do {                                                                    \
    cudaError_t e_ = (expr);                                              \
    if (e_ != cudaSuccess) return set_cuda_error(e_, __FILE__, __LINE__); \
  } while (0)

// ---------------------------------------------------------------- staged upload
constexpr size_t UP_CHUNK = size_t(8) << 20;   // bytes per staging slot
constexpr int UP_SLOTS_PER_THREAD = 2;
constexpr int UP_MAX_THREADS = 16;

// Copy into a staging slot with non-temporal stores: the slot is read next by the copy engine, not by
// this core, so its lines need neither be fetched for ownership nor kept in the cache.  With every rank
// of a box staging at once the copies are bound by host memory traffic; this removes a quarter of it.
__attribute__((target("avx2"))) static void copy_stream_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
  size_t i = 0;
  // dst is 4 KB aligned (pinned slot); src has any alignment
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}
static void copy_to_slot(uint8_t* dst, const uint8_t* src, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2") && !(getenv("GV_STAGE_MEMCPY") && getenv("GV_STAGE_MEMCPY")[0] == '1');
  if (avx2) copy_stream_avx2(dst, src, n);
  else memcpy(dst, src, n);
}

struct UploadRing {
  std::mutex mu;                  // one upload at a time per process
  int device = -1;
  int n_slots = 0;
  uint8_t* slot[UP_MAX_THREADS * UP_SLOTS_PER_THREAD] = {};
  cudaEvent_t ev[UP_MAX_THREADS * UP_SLOTS_PER_THREAD] = {};
};
static UploadRing g_ring;

static int ring_release_locked() {
  for (int i = 0; i < g_ring.n_slots; ++i) {
    if (g_ring.ev[i]) { cudaEventSynchronize(g_ring.ev[i]); cudaEventDestroy(g_ring.ev[i]); g_ring.ev[i] = nullptr; }
    if (g_ring.slot[i]) { cudaFreeHost(g_ring.slot[i]); g_ring.slot[i] = nullptr; }
  }
  g_ring.n_slots = 0;
  g_ring.device = -1;
  return GV_OK;
}

static int ring_prepare_locked(int device, int n_slots) {
  if (g_ring.device == device && g_ring.n_slots >= n_slots) return GV_OK;
  if (g_ring.device != device) ring_release_locked();
  g_ring.device = device;
  for (int i = g_ring.n_slots; i < n_slots; ++i) {
    GV_HOST_TRY(cudaHostAlloc(reinterpret_cast<void**>(&g_ring.slot[i]), UP_CHUNK, cudaHostAllocPortable));
    GV_HOST_TRY(cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming));
    g_ring.n_slots = i + 1;
  }
  return GV_OK;
}

}  // namespace gv

using namespace gv;

extern "C" {

int gv_upload_pageable(void* dst_dev, const void* src_host, size_t bytes, int n_threads, void* stream) {
  if (bytes == 0) return GV_OK;
  if (!dst_dev || !src_host) return set_error(GV_ERR_ARG, "gv_upload_pageable: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int device = 0;
  GV_HOST_TRY(cudaGetDevice(&device));
  if (bytes < 2 * UP_CHUNK) {
    // small arrays: the runtime's own staging is as good
    GV_HOST_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
    return GV_OK;
  }
  if (n_threads <= 0) {
    const unsigned hw = std::thread::hardware_concurrency();
    n_threads = hw >= 16 ? 8 : (hw >= 8 ? 4 : 2);
  }
  if (n_threads > UP_MAX_THREADS) n_threads = UP_MAX_THREADS;
  const size_t n_chunks = (bytes + UP_CHUNK - 1) / UP_CHUNK;
  if (size_t(n_threads) > n_chunks) n_threads = int(n_chunks);
  std::lock_guard<std::mutex> lock(g_ring.mu);
  int rc = ring_prepare_locked(device, n_threads * UP_SLOTS_PER_THREAD);
  if (rc != GV_OK) return rc;
  std::atomic<int> err{0};
  std::atomic<size_t> next{0};
  auto worker = [&](int w) {
    if (cudaSetDevice(device) != cudaSuccess) { err.store(1); return; }
    int use = 0;
    for (;;) {
      const size_t c = next.fetch_add(1);
      if (c >= n_chunks || err.load()) break;
      const size_t off = c * UP_CHUNK;
      const size_t n = bytes - off < UP_CHUNK ? bytes - off : UP_CHUNK;
      const int s = w * UP_SLOTS_PER_THREAD + (use++ % UP_SLOTS_PER_THREAD);
      // the DMA that last read this slot (this call or an earlier one) must be done
      if (cudaEventSynchronize(g_ring.ev[s]) != cudaSuccess) { err.store(2); break; }
      copy_to_slot(g_ring.slot[s], static_cast<const uint8_t*>(src_host) + off, n);
      if (cudaMemcpyAsync(static_cast<uint8_t*>(dst_dev) + off, g_ring.slot[s], n, cudaMemcpyHostToDevice, st) !=
              cudaSuccess ||
          cudaEventRecord(g_ring.ev[s], st) != cudaSuccess) {
        err.store(3);
        break;
      }
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(n_threads - 1);
  for (int w = 1; w < n_threads; ++w) pool.emplace_back(worker, w);
  worker(0);
  for (auto& t : pool) t.join();
  if (err.load()) {
    cudaError_t e = cudaGetLastError();
    return e != cudaSuccess ? set_cuda_error(e, __FILE__, __LINE__)
                            : set_error(GV_ERR_CUDA, "gv_upload_pageable: a staging copy failed");
  }
  return GV_OK;
}

// ---------------------------------------------------------------- packed count upload
// scRNA-seq count matrices hand over 8 bytes per stored entry (float32 value + int32 column) of which
// 3 carry information: the value is an integer <= 255 and the column index fits 16 bits.  The staging
// pass already touches every byte, so it packs while it copies (uint8 value, uint16 column): 2.7x fewer
// bytes cross the PCIe link, which is what bounds the upload when every rank of a box uploads at once.
// ok = 0 as soon as an entry does not fit (the caller then uploads the arrays as they are).
static bool pack_chunk_scalar(const float* v, const int32_t* c, size_t n, uint8_t* dv, uint16_t* dc);
__attribute__((target("avx2"))) static bool pack_chunk_avx2(const float* v, const int32_t* c, size_t n,
                                                            uint8_t* dv, uint16_t* dc) {
  const __m256i fix = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  __m256i bad = _mm256_setzero_si256();     // OR of everything that must be zero
  __m256 inexact = _mm256_setzero_ps();     // OR of the lanes whose value is not an integer
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    __m256i iv[4];
#pragma GCC unroll 4
    for (int k = 0; k < 4; ++k) {
      const __m256 f = _mm256_loadu_ps(v + i + 8 * k);
      iv[k] = _mm256_cvttps_epi32(f);
      inexact = _mm256_or_ps(inexact, _mm256_cmp_ps(_mm256_cvtepi32_ps(iv[k]), f, _CMP_NEQ_UQ));
      bad = _mm256_or_si256(bad, _mm256_andnot_si256(_mm256_set1_epi32(0xFF), iv[k]));
    }
    const __m256i p01 = _mm256_packus_epi32(iv[0], iv[1]);
    const __m256i p23 = _mm256_packus_epi32(iv[2], iv[3]);
    const __m256i b = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(p01, p23), fix);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dv + i), b);
#pragma GCC unroll 2
    for (int k = 0; k < 2; ++k) {
      const __m256i a0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(c + i + 16 * k));
      const __m256i a1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(c + i + 16 * k + 8));
      bad = _mm256_or_si256(bad, _mm256_andnot_si256(_mm256_set1_epi32(0xFFFF), _mm256_or_si256(a0, a1)));
      const __m256i w = _mm256_permute4x64_epi64(_mm256_packus_epi32(a0, a1), 0xD8);
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dc + i + 16 * k), w);
    }
  }
  _mm_sfence();
  bool ok = _mm256_testz_si256(bad, bad) && _mm256_testz_ps(inexact, inexact);
  if (i < n && !pack_chunk_scalar(v + i, c + i, n - i, dv + i, dc + i)) ok = false;
  return ok;
}
static bool pack_chunk_scalar(const float* v, const int32_t* c, size_t n, uint8_t* dv, uint16_t* dc) {
  bool ok = true;
  for (size_t i = 0; i < n; ++i) {
    const float f = v[i];
    const bool in = f >= 0.0f && f <= 255.0f;          // also false for NaN
    const int32_t iv = in ? int32_t(f) : 0;
    if (!in || float(iv) != f || (c[i] & ~0xFFFF)) ok = false;
    dv[i] = uint8_t(iv);
    dc[i] = uint16_t(c[i]);
  }
  return ok;
}
static bool pack_chunk(const float* v, const int32_t* c, size_t n, uint8_t* dv, uint16_t* dc) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  // cvttps yields 0x80000000 for NaN / out-of-range values: caught by the range mask
  return avx2 ? pack_chunk_avx2(v, c, n, dv, dc) : pack_chunk_scalar(v, c, n, dv, dc);
}

constexpr size_t PK_ELEMS = size_t(2) << 20;      // entries per staging slot: 2 MB of values + 4 MB of columns

// ---------------------------------------------------------------- registered upload
// The other way to read pageable memory at PCIe speed: page-lock the caller's pages in place
// (cudaHostRegister), DMA straight out of them, unlock afterwards.  No CPU copy touches the data, so
// it does not compete for host memory bandwidth -- what matters when the eight ranks of a box upload
// their blocks of the matrix at the same time.  The range is registered in page-aligned chunks, each
// chunk's DMA running under the registration of the next.
struct RegToken {
  std::vector<std::pair<void*, size_t>> ranges;
  cudaEvent_t done = nullptr;
};
// chunk size of the registration (bytes); GV_REG_CHUNK_MB overrides (0 = the whole range in one call)
static size_t reg_chunk_bytes() {
  static const size_t v = [] {
    const char* e = getenv("GV_REG_CHUNK_MB");
    if (!e) return size_t(0);
    return size_t(atol(e)) << 20;
  }();
  return v;
}

int gv_upload_registered(void* dst_dev, const void* src_host, size_t bytes, void* stream, void** token_out) {
  if (!token_out) return set_error(GV_ERR_ARG, "gv_upload_registered: null token");
  *token_out = nullptr;
  if (bytes == 0) return GV_OK;
  if (!dst_dev || !src_host) return set_error(GV_ERR_ARG, "gv_upload_registered: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t page = size_t(sysconf(_SC_PAGESIZE));
  const uintptr_t begin = reinterpret_cast<uintptr_t>(src_host), end = begin + bytes;
  RegToken* tok = new RegToken();
  uintptr_t cur = begin;
  int rc = GV_OK;
  while (cur < end) {
    const size_t chunk = reg_chunk_bytes();
    uintptr_t nxt = chunk ? ((cur + chunk) / page) * page : end;      // chunk boundaries on page boundaries
    if (nxt > end || nxt <= cur) nxt = end;
    const uintptr_t p0 = (cur / page) * page, p1 = ((nxt + page - 1) / page) * page;
    cudaError_t e = cudaHostRegister(reinterpret_cast<void*>(p0), p1 - p0, cudaHostRegisterPortable);
    if (e == cudaSuccess) {
      tok->ranges.emplace_back(reinterpret_cast<void*>(p0), p1 - p0);
    } else {
      cudaGetLastError();      // e.g. a page shared with a range that is already registered: plain copy
    }
    e = cudaMemcpyAsync(static_cast<uint8_t*>(dst_dev) + (cur - begin), reinterpret_cast<const void*>(cur),
                        nxt - cur, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = set_cuda_error(e, __FILE__, __LINE__); break; }
    cur = nxt;
  }
  if (rc == GV_OK) {
    cudaError_t e = cudaEventCreateWithFlags(&tok->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(tok->done, st);
    if (e != cudaSuccess) rc = set_cuda_error(e, __FILE__, __LINE__);
  }
  if (rc != GV_OK) {
    cudaStreamSynchronize(st);
    for (auto& r : tok->ranges) cudaHostUnregister(r.first);
    if (tok->done) cudaEventDestroy(tok->done);
    delete tok;
    return rc;
  }
  *token_out = tok;
  return GV_OK;
}

int gv_upload_release(void* token) {
  if (!token) return GV_OK;
  RegToken* tok = static_cast<RegToken*>(token);
  int rc = GV_OK;
  if (tok->done) {
    cudaError_t e = cudaEventSynchronize(tok->done);      // the DMAs have read the pages
    if (e != cudaSuccess) rc = set_cuda_error(e, __FILE__, __LINE__);
    cudaEventDestroy(tok->done);
  }
  for (auto& r : tok->ranges) cudaHostUnregister(r.first);
  delete tok;
  return rc;
}

// host-only entry for tests: the packing pass alone
int gv_pack_counts_host(const float* src_vals, const int32_t* src_cols, size_t nnz, uint8_t* dst_vals,
                        uint16_t* dst_cols, int* ok_host) {
  if (!ok_host || (nnz && (!src_vals || !src_cols || !dst_vals || !dst_cols)))
    return set_error(GV_ERR_ARG, "gv_pack_counts_host: null pointer");
  if ((reinterpret_cast<uintptr_t>(dst_vals) | reinterpret_cast<uintptr_t>(dst_cols)) & 31)
    return set_error(GV_ERR_ARG, "gv_pack_counts_host: destinations must be 32-byte aligned");
  *ok_host = pack_chunk(src_vals, src_cols, nnz, dst_vals, dst_cols) ? 1 : 0;
  return GV_OK;
}

int gv_upload_packed_counts(void* dst_vals_u8, void* dst_cols_u16, const float* src_vals, const int32_t* src_cols,
                            size_t nnz, int n_threads, int* ok_host, void* stream) {
  if (!ok_host) return set_error(GV_ERR_ARG, "gv_upload_packed_counts: null ok_host");
  *ok_host = 1;
  if (nnz == 0) return GV_OK;
  if (!dst_vals_u8 || !dst_cols_u16 || !src_vals || !src_cols)
    return set_error(GV_ERR_ARG, "gv_upload_packed_counts: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int device = 0;
  GV_HOST_TRY(cudaGetDevice(&device));
  if (n_threads <= 0) {
    const unsigned hw = std::thread::hardware_concurrency();
    n_threads = hw >= 16 ? 8 : (hw >= 8 ? 4 : 2);
  }
  if (n_threads > UP_MAX_THREADS) n_threads = UP_MAX_THREADS;
  const size_t n_chunks = (nnz + PK_ELEMS - 1) / PK_ELEMS;
  if (size_t(n_threads) > n_chunks) n_threads = int(n_chunks);
  std::lock_guard<std::mutex> lock(g_ring.mu);
  int rc = ring_prepare_locked(device, n_threads * UP_SLOTS_PER_THREAD);
  if (rc != GV_OK) return rc;
  std::atomic<int> err{0};
  std::atomic<int> fits{1};
  std::atomic<size_t> next{0};
  auto worker = [&](int w) {
    if (cudaSetDevice(device) != cudaSuccess) { err.store(1); return; }
    int use = 0;
    for (;;) {
      const size_t c = next.fetch_add(1);
      if (c >= n_chunks || err.load() || !fits.load()) break;
      const size_t off = c * PK_ELEMS;
      const size_t n = nnz - off < PK_ELEMS ? nnz - off : PK_ELEMS;
      const int s = w * UP_SLOTS_PER_THREAD + (use++ % UP_SLOTS_PER_THREAD);
      if (cudaEventSynchronize(g_ring.ev[s]) != cudaSuccess) { err.store(2); break; }
      uint8_t* sv = g_ring.slot[s];
      uint16_t* sc = reinterpret_cast<uint16_t*>(g_ring.slot[s] + PK_ELEMS);
      if (!pack_chunk(src_vals + off, src_cols + off, n, sv, sc)) { fits.store(0); break; }
      if (cudaMemcpyAsync(static_cast<uint8_t*>(dst_vals_u8) + off, sv, n, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaMemcpyAsync(static_cast<uint16_t*>(dst_cols_u16) + off, sc, 2 * n, cudaMemcpyHostToDevice, st) !=
              cudaSuccess ||
          cudaEventRecord(g_ring.ev[s], st) != cudaSuccess) {
        err.store(3);
        break;
      }
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(n_threads - 1);
  for (int w = 1; w < n_threads; ++w) pool.emplace_back(worker, w);
  worker(0);
  for (auto& t : pool) t.join();
  if (err.load()) {
    cudaError_t e = cudaGetLastError();
    return e != cudaSuccess ? set_cuda_error(e, __FILE__, __LINE__)
                            : set_error(GV_ERR_CUDA, "gv_upload_packed_counts: a staging copy failed");
  }
  *ok_host = fits.load();
  return GV_OK;
}

int gv_runtime_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_ring.mu);
  return ring_release_locked();
}

// ---------------------------------------------------------------- shared host segments
int gv_shm_open(const char* name, size_t bytes, int create, int cuda_register, void** ptr_out_host) {
  if (!name || name[0] != '/' || bytes == 0 || !ptr_out_host)
    return set_error(GV_ERR_ARG, "gv_shm_open: name must start with '/', bytes > 0");
  int fd = -1;
  if (create) {
    shm_unlink(name);                                   // a stale segment of a crashed run
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  } else {
    fd = shm_open(name, O_RDWR, 0600);
  }
  if (fd < 0) {
    char buf[256];
    snprintf(buf, sizeof(buf), "gv_shm_open: shm_open(%s) failed: %s", name, strerror(errno));
    return set_error(GV_ERR_ARG, buf);
  }
  if (create && ftruncate(fd, off_t(bytes)) != 0) {
    char buf[256];
    snprintf(buf, sizeof(buf), "gv_shm_open: ftruncate(%s, %zu) failed: %s", name, bytes, strerror(errno));
    close(fd);
    shm_unlink(name);
    return set_error(GV_ERR_ARG, buf);
  }
  if (!create) {
    struct stat sb;
    if (fstat(fd, &sb) != 0 || size_t(sb.st_size) < bytes) {
      close(fd);
      return set_error(GV_ERR_ARG, "gv_shm_open: segment is smaller than requested");
    }
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    if (create) shm_unlink(name);
    return set_error(GV_ERR_ARG, "gv_shm_open: mmap failed");
  }
  if (cuda_register) {
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
      munmap(p, bytes);
      if (create) shm_unlink(name);
      return set_cuda_error(e, __FILE__, __LINE__);
    }
  }
  *ptr_out_host = p;
  return GV_OK;
}

int gv_shm_close(void* ptr, size_t bytes, int cuda_registered, const char* unlink_name) {
  int rc = GV_OK;
  if (ptr) {
    if (cuda_registered) {
      cudaError_t e = cudaHostUnregister(ptr);
      if (e != cudaSuccess) rc = set_cuda_error(e, __FILE__, __LINE__);
    }
    munmap(ptr, bytes);
  }
  if (unlink_name && unlink_name[0] == '/') shm_unlink(unlink_name);
  return rc;
}

int gv_flag_store(uint64_t* flag, uint64_t value) {
  if (!flag) return set_error(GV_ERR_ARG, "gv_flag_store: null");
  __atomic_store_n(flag, value, __ATOMIC_RELEASE);
  return GV_OK;
}

uint64_t gv_flag_load(const uint64_t* flag) { return flag ? __atomic_load_n(flag, __ATOMIC_ACQUIRE) : 0; }

int gv_flag_wait_ge(const uint64_t* flag, uint64_t value, int timeout_ms) {
  if (!flag) return set_error(GV_ERR_ARG, "gv_flag_wait_ge: null");
  const auto t0 = std::chrono::steady_clock::now();
  int spins = 0;
  while (__atomic_load_n(flag, __ATOMIC_ACQUIRE) < value) {
    if (++spins < 2000) {
      __builtin_ia32_pause();
    } else {
      std::this_thread::sleep_for(std::chrono::microseconds(20));
      const auto dt = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0);
      if (timeout_ms >= 0 && dt.count() > timeout_ms)
        return set_error(GV_ERR_TIMEOUT, "gv_flag_wait_ge: timed out waiting for another rank");
    }
  }
  return GV_OK;
}

// ---------------------------------------------------------------- result rectangles
int gv_copy_rect_d2h(void* dst_host, size_t dst_pitch_bytes, const void* src_dev, size_t src_pitch_bytes,
                     size_t width_bytes, size_t rows, void* stream) {
  if (width_bytes == 0 || rows == 0) return GV_OK;
  if (!dst_host || !src_dev || dst_pitch_bytes < width_bytes || src_pitch_bytes < width_bytes)
    return set_error(GV_ERR_ARG, "gv_copy_rect_d2h: bad argument");
  GV_HOST_TRY(cudaMemcpy2DAsync(dst_host, dst_pitch_bytes, src_dev, src_pitch_bytes, width_bytes, rows,
                                cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return GV_OK;
}

// ---------------------------------------------------------------- peer-mapped device buffers
int gv_ipc_alloc(size_t bytes, void** dev_ptr_out, void* handle_out_64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
  if (bytes == 0 || !dev_ptr_out || !handle_out_64) return set_error(GV_ERR_ARG, "gv_ipc_alloc: bad argument");
  void* p = nullptr;
  GV_HOST_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return set_cuda_error(e, __FILE__, __LINE__); }
  memcpy(handle_out_64, &h, 64);
  *dev_ptr_out = p;
  return GV_OK;
}

int gv_ipc_open(const void* handle_64, void** dev_ptr_out) {
  if (!handle_64 || !dev_ptr_out) return set_error(GV_ERR_ARG, "gv_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, 64);
  GV_HOST_TRY(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return GV_OK;
}

int gv_ipc_close(void* dev_ptr) {
  if (dev_ptr) GV_HOST_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return GV_OK;
}

int gv_ipc_free(void* dev_ptr) {
  if (dev_ptr) GV_HOST_TRY(cudaFree(dev_ptr));
  return GV_OK;
}

}  // extern "C"
