// extras.cu -- handle teardown, the host-buffer (end-to-end) entry point and the optional NCCL all-gather.
//
// scimlop_kron_mul_host streams the columns of V through the device in chunks (H2D copy, the kron kernels
// and the D2H copy of the previous chunks overlap on three streams): columns are independent units of the
// operator (src/SciMLOperators.jl:177-198), so chunking is exact.
#include <chrono>
#include <cstdio>
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "context.h"
#include "krylov.cuh"

namespace {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int ensure_host_streams(scimlop_context* c) {
  if (c->host_streams_ready) return SCIMLOP_OK;
  SCIMLOP_CUDA(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  SCIMLOP_CUDA(c, cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
  SCIMLOP_CUDA(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < SCIMLOP_HOST_SLOTS; i++) {
    SCIMLOP_CUDA(c, cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
    SCIMLOP_CUDA(c, cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
    SCIMLOP_CUDA(c, cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
  }
  c->host_streams_ready = true;
  return SCIMLOP_OK;
}

}  // namespace

extern "C" int scimlop_kron_mul_host(scimlop_handle_t h, int dtype, int precision, int nfactors,
                                     const void* const* factors, const int64_t* dims, const int64_t* lds,
                                     const int32_t* kinds, const void* V, int64_t ldv, void* W, int64_t ldw,
                                     int64_t k, const void* alpha, const void* beta, int64_t chunk_cols) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
  if (nfactors < 1 || nfactors > 8 || !dims || !kinds || !lds || !factors)
    return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_host: bad factor description");
  if (k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_mul_host: negative k");
  const size_t es = dtype == SCIMLOP_F64 ? 8 : 4;
  long long rows = 1, cols = 1;
  for (int f = 0; f < nfactors; f++) {
    if (dims[2 * f] < 0 || dims[2 * f + 1] < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_mul_host: negative size");
    rows *= dims[2 * f]; cols *= dims[2 * f + 1];
  }
  if (k == 0 || rows == 0) return SCIMLOP_OK;
  if (!V || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_host: V/W null");
  if (ldv != cols || ldw != rows) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_mul_host: V and W must be contiguous");
  scimlop_device_guard guard(h->device);
  if (int rc = ensure_host_streams(h)) return rc;

  double beta_val = 0.0;
  if (beta) beta_val = dtype == SCIMLOP_F64 ? *static_cast<const double*>(beta) : (double)*static_cast<const float*>(beta);
  const bool readw = beta_val != 0.0;

  // chunk size: ~32 MiB of V per chunk by default, at least 1 column
  long long cc = chunk_cols > 0 ? chunk_cols : std::max<long long>(1, (32ll << 20) / std::max<long long>(1, cols * (long long)es));
  cc = std::min<long long>(cc, k);

  // ---- carve the staging allocation: factors | V slots | W slots | workspace ----
  size_t off = 0;
  size_t fac_off[8];
  for (int f = 0; f < nfactors; f++) {
    fac_off[f] = off;
    if ((kinds[f] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE) {
      const int64_t srows = (kinds[f] & SCIMLOP_TRANS) ? dims[2 * f + 1] : dims[2 * f];
      if (!factors[f] || lds[f] < std::max<int64_t>(1, srows)) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_mul_host: bad dense factor %d", f);
      off += align_up((size_t)lds[f] * (size_t)(dims[2 * f] + dims[2 * f + 1] - srows) * es, 256);
    } else if (kinds[f] == SCIMLOP_DIAG) {
      if (!factors[f]) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_host: null diagonal factor %d", f);
      off += align_up((size_t)dims[2 * f] * es, 256);
    } else if (kinds[f] == SCIMLOP_BANDED) {
      if (!factors[f]) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_host: null banded factor %d", f);
      if (lds[f] < 0 || lds[f] > 4) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_mul_host: bad half-bandwidth");
      off += align_up((size_t)dims[2 * f] * (size_t)(2 * lds[f] + 1) * es, 256);
    } else if (kinds[f] == SCIMLOP_CSR) {
      return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "kron_mul_host: sparse factor %d: upload the CSR arrays and call scimlop_kron_mul", f);
    }
  }
  const size_t vslot = align_up((size_t)cols * cc * es, 256), wslot = align_up((size_t)rows * cc * es, 256);
  const size_t v_off = off; off += vslot * SCIMLOP_HOST_SLOTS;
  const size_t w_off = off; off += wslot * SCIMLOP_HOST_SLOTS;
  size_t ws_bytes = 0;
  if (int rc = scimlop_kron_workspace_bytes(dtype, nfactors, dims, kinds, cc, &ws_bytes)) return scimlop_fail(h, rc, "kron_mul_host: bad factors");
  const size_t ws_off = off; off += ws_bytes;
  if (off > h->stage_bytes) {
    if (h->stage) { SCIMLOP_CUDA(h, cudaDeviceSynchronize()); SCIMLOP_CUDA(h, cudaFree(h->stage)); h->stage = nullptr; h->stage_bytes = 0; }
    SCIMLOP_CUDA(h, cudaMalloc(&h->stage, off));
    h->stage_bytes = off;
  }
  char* base = static_cast<char*>(h->stage);

  // factors: uploaded once per call on the H2D stream
  const void* dfac[8];
  for (int f = 0; f < nfactors; f++) {
    dfac[f] = nullptr;
    size_t bytes = 0;
    if ((kinds[f] & ~SCIMLOP_TRANS) == SCIMLOP_DENSE) {
      // the logical extent only: ld * (stored cols - 1) + stored rows (a view with ld > rows may end at the end of
      // its allocation, so the padding behind the last column must not be read)
      const bool tr = (kinds[f] & SCIMLOP_TRANS) != 0;
      const size_t scols = (size_t)(tr ? dims[2 * f] : dims[2 * f + 1]), srows = (size_t)(tr ? dims[2 * f + 1] : dims[2 * f]);
      bytes = scols ? ((size_t)lds[f] * (scols - 1) + srows) * es : 0;
    }
    else if (kinds[f] == SCIMLOP_DIAG) bytes = (size_t)dims[2 * f] * es;
    else if (kinds[f] == SCIMLOP_BANDED) bytes = (size_t)dims[2 * f] * (size_t)(2 * lds[f] + 1) * es;
    if (bytes) {
      dfac[f] = base + fac_off[f];
      SCIMLOP_CUDA(h, cudaMemcpyAsync(base + fac_off[f], factors[f], bytes, cudaMemcpyHostToDevice, h->s_h2d));
    }
  }

  // chunk schedule: ramp up (cc/8, cc/4, cc/2), full chunks, ramp down -- the first upload and the last download are
  // the only transfers nothing overlaps with, so they should be small (auto mode only; an explicit chunk_cols is kept)
  std::vector<long long> sizes;
  {
    std::vector<long long> ramp;
    long long used = 0;
    if (chunk_cols <= 0)
      for (long long sz = std::max<long long>(1, cc / 8); sz < cc && 2 * (used + sz) <= k / 2; sz *= 2) { ramp.push_back(sz); used += sz; }
    long long mid = k - 2 * used;
    sizes = ramp;
    while (mid > 0) { const long long sz = std::min(cc, mid); sizes.push_back(sz); mid -= sz; }
    for (auto it = ramp.rbegin(); it != ramp.rend(); ++it) sizes.push_back(*it);
  }
  const long long nchunks = (long long)sizes.size();
#ifdef SCIMLOP_HOST_TRACE
  // timing build (tools/e2e_trace.py): timed events around every transfer and compute of a chunk + host enqueue times
  static cudaEvent_t tev[64][6];
  static bool tev_ready = false;
  if (!tev_ready) { for (auto& r : tev) for (auto& e : r) cudaEventCreate(&e); tev_ready = true; }
  static double host_us[64];
  const auto host_t0 = std::chrono::steady_clock::now();
  cudaEventRecord(tev[63][0], h->s_h2d);
#define HOST_TEV(ci, i, stream) do { if ((ci) < 63) cudaEventRecord(tev[ci][i], stream); } while (0)
#else
#define HOST_TEV(ci, i, stream) do { } while (0)
#endif
  long long c0 = 0;
  for (long long ci = 0; ci < nchunks; c0 += sizes[ci], ci++) {
    const int slot = (int)(ci % SCIMLOP_HOST_SLOTS);
    const long long nc = sizes[ci];
    char* dV = base + v_off + vslot * slot;
    char* dW = base + w_off + wslot * slot;
    if (ci >= SCIMLOP_HOST_SLOTS) {
      // the slot is free once its previous compute (V) / download (W) finished
      SCIMLOP_CUDA(h, cudaStreamWaitEvent(h->s_h2d, h->ev_comp[slot], 0));
      if (readw) SCIMLOP_CUDA(h, cudaStreamWaitEvent(h->s_h2d, h->ev_d2h[slot], 0));
    }
    HOST_TEV(ci, 0, h->s_h2d);
    SCIMLOP_CUDA(h, cudaMemcpyAsync(dV, static_cast<const char*>(V) + (size_t)c0 * cols * es, (size_t)nc * cols * es,
                                    cudaMemcpyHostToDevice, h->s_h2d));
    HOST_TEV(ci, 1, h->s_h2d);
    if (readw)
      SCIMLOP_CUDA(h, cudaMemcpyAsync(dW, static_cast<const char*>(W) + (size_t)c0 * rows * es, (size_t)nc * rows * es,
                                      cudaMemcpyHostToDevice, h->s_h2d));
    SCIMLOP_CUDA(h, cudaEventRecord(h->ev_h2d[slot], h->s_h2d));
    SCIMLOP_CUDA(h, cudaStreamWaitEvent(h->s_comp, h->ev_h2d[slot], 0));
    if (ci >= SCIMLOP_HOST_SLOTS) SCIMLOP_CUDA(h, cudaStreamWaitEvent(h->s_comp, h->ev_d2h[slot], 0));
    HOST_TEV(ci, 2, h->s_comp);
#ifdef SCIMLOP_HOST_TRACE_NOCOMPUTE
    int rc = SCIMLOP_OK;                            // timing experiment: the chunked copies alone (W is garbage)
#else
    int rc = scimlop_kron_mul_impl(h, dtype, precision, nfactors, dfac, dims, lds, kinds, dV, cols, dW, rows, nc, alpha,
                                   beta, base + ws_off, ws_bytes, h->s_comp);
#endif
    if (rc) return rc;
    HOST_TEV(ci, 3, h->s_comp);
    SCIMLOP_CUDA(h, cudaEventRecord(h->ev_comp[slot], h->s_comp));
    SCIMLOP_CUDA(h, cudaStreamWaitEvent(h->s_d2h, h->ev_comp[slot], 0));
    HOST_TEV(ci, 4, h->s_d2h);
    SCIMLOP_CUDA(h, cudaMemcpyAsync(static_cast<char*>(W) + (size_t)c0 * rows * es, dW, (size_t)nc * rows * es,
                                    cudaMemcpyDeviceToHost, h->s_d2h));
    SCIMLOP_CUDA(h, cudaEventRecord(h->ev_d2h[slot], h->s_d2h));
    HOST_TEV(ci, 5, h->s_d2h);
#ifdef SCIMLOP_HOST_TRACE
    if (ci < 63) host_us[ci] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
#endif
  }
  SCIMLOP_CUDA(h, cudaStreamSynchronize(h->s_d2h));
  SCIMLOP_CUDA(h, cudaStreamSynchronize(h->s_comp));
  SCIMLOP_CUDA(h, cudaStreamSynchronize(h->s_h2d));
#ifdef SCIMLOP_HOST_TRACE
  fprintf(stderr, "e2e_trace: %lld chunks; times in us from the first enqueue on the H2D stream\n", nchunks);
  for (long long ci = 0; ci < nchunks && ci < 63; ci++) {
    float t[6];
    for (int i = 0; i < 6; i++) cudaEventElapsedTime(&t[i], tev[63][0], tev[ci][i]);
    fprintf(stderr, "e2e_trace chunk %2lld (%3lld cols): H2D %8.1f..%8.1f | compute %8.1f..%8.1f | D2H %8.1f..%8.1f | host enqueued at %8.1f\n", ci,
            sizes[ci], t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3, host_us[ci]);
  }
#endif
  return SCIMLOP_OK;
}

// ================================================================================================
// NCCL (dlopen'ed: the library must load, and the single-GPU path must work, without libnccl)
// ================================================================================================
namespace {

typedef struct { char internal[128]; } nccl_unique_id;
typedef int (*nccl_get_unique_id_fn)(nccl_unique_id*);
typedef int (*nccl_comm_init_rank_fn)(void**, int, nccl_unique_id, int);
typedef int (*nccl_comm_destroy_fn)(void*);
typedef int (*nccl_all_gather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_broadcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_group_fn)(void);
typedef int (*nccl_all_reduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_get_error_string_fn)(int);
constexpr int NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8;  // ncclDataType_t (nccl.h)

void* g_nccl = nullptr;

void* nccl_sym(const char* name) {
  if (!g_nccl) {
    g_nccl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl) g_nccl = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  }
  return g_nccl ? dlsym(g_nccl, name) : nullptr;
}

}  // namespace

extern "C" int scimlop_comm_unique_id(void* id) {
  if (!id) return SCIMLOP_ERR_NULL_POINTER;
  auto fn = reinterpret_cast<nccl_get_unique_id_fn>(nccl_sym("ncclGetUniqueId"));
  if (!fn) return SCIMLOP_ERR_NCCL;
  return fn(static_cast<nccl_unique_id*>(id)) == 0 ? SCIMLOP_OK : SCIMLOP_ERR_NCCL;
}

extern "C" int scimlop_comm_init(scimlop_handle_t h, int nranks, int rank, const void* id) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!id) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "comm_init: unique id is null");
  if (nranks < 1 || rank < 0 || rank >= nranks) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "comm_init: bad rank %d of %d", rank, nranks);
  auto fn = reinterpret_cast<nccl_comm_init_rank_fn>(nccl_sym("ncclCommInitRank"));
  if (!fn) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "libnccl.so.2 not found: %s", dlerror());
  scimlop_device_guard guard(h->device);
  nccl_unique_id uid;
  memcpy(&uid, id, sizeof(uid));
  void* comm = nullptr;
  int r = fn(&comm, nranks, uid, rank);
  if (r != 0) {
    auto es = reinterpret_cast<nccl_get_error_string_fn>(nccl_sym("ncclGetErrorString"));
    return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclCommInitRank failed: %s", es ? es(r) : "?");
  }
  h->nccl_comm = comm; h->nranks = nranks; h->rank = rank;
  return SCIMLOP_OK;
}

extern "C" int scimlop_comm_destroy(scimlop_handle_t h) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (h->nccl_comm) {
    auto fn = reinterpret_cast<nccl_comm_destroy_fn>(nccl_sym("ncclCommDestroy"));
    if (fn) fn(h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  return SCIMLOP_OK;
}

extern "C" int scimlop_allgather_cols(scimlop_handle_t h, int dtype, void* W_full, int64_t rows, int64_t cols_per_rank,
                                      void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!h->nccl_comm) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "allgather_cols: call scimlop_comm_init first");
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
  if (rows < 0 || cols_per_rank < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "allgather_cols: negative size");
  if (rows == 0 || cols_per_rank == 0) return SCIMLOP_OK;
  if (!W_full) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "allgather_cols: W_full null");
  auto fn = reinterpret_cast<nccl_all_gather_fn>(nccl_sym("ncclAllGather"));
  if (!fn) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclAllGather not found");
  scimlop_device_guard guard(h->device);
  const size_t es = dtype == SCIMLOP_F64 ? 8 : 4;
  const size_t count = (size_t)rows * (size_t)cols_per_rank;
  const char* send = static_cast<const char*>(W_full) + (size_t)h->rank * count * es;  // in-place form
  int r = fn(send, W_full, count, dtype == SCIMLOP_F64 ? NCCL_FLOAT64 : NCCL_FLOAT32, h->nccl_comm,
             static_cast<cudaStream_t>(stream));
  if (r != 0) {
    auto es2 = reinterpret_cast<nccl_get_error_string_fn>(nccl_sym("ncclGetErrorString"));
    return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclAllGather failed: %s", es2 ? es2(r) : "?");
  }
  return SCIMLOP_OK;
}

extern "C" int scimlop_kron_mul_allgather(scimlop_handle_t h, int dtype, int precision, int nfactors,
                                          const void* const* factors, const int64_t* dims, const int64_t* lds,
                                          const int32_t* kinds, const void* V, void* W_full, int64_t cols_per_rank,
                                          const void* alpha, int64_t chunk_cols, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!h->nccl_comm) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "kron_mul_allgather: call scimlop_comm_init first");
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
  if (nfactors < 1 || nfactors > 8 || !dims || !kinds) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_allgather: bad factor description");
  if (cols_per_rank < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "kron_mul_allgather: negative column count");
  if (cols_per_rank == 0) return SCIMLOP_OK;
  if (!V || !W_full) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "kron_mul_allgather: V/W null");
  auto bcast = reinterpret_cast<nccl_broadcast_fn>(nccl_sym("ncclBroadcast"));
  auto gstart = reinterpret_cast<nccl_group_fn>(nccl_sym("ncclGroupStart"));
  auto gend = reinterpret_cast<nccl_group_fn>(nccl_sym("ncclGroupEnd"));
  if (!bcast || !gstart || !gend) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "NCCL symbols not found");
  scimlop_device_guard guard(h->device);
  if (int rc = ensure_host_streams(h)) return rc;   // s_comp = compute, s_d2h = exchange
  const size_t es = dtype == SCIMLOP_F64 ? 8 : 4;
  long long rows = 1, cols = 1;
  for (int f = 0; f < nfactors; f++) { rows *= dims[2 * f]; cols *= dims[2 * f + 1]; }
  long long cc = chunk_cols > 0 ? chunk_cols : std::max<long long>(1, (cols_per_rank + 3) / 4);
  cc = std::min<long long>(cc, cols_per_rank);
  cudaStream_t user = static_cast<cudaStream_t>(stream), comp = h->s_comp, comm = h->s_d2h;
  // both internal streams start after whatever the caller has enqueued
  SCIMLOP_CUDA(h, cudaEventRecord(h->ev_h2d[0], user));
  SCIMLOP_CUDA(h, cudaStreamWaitEvent(comp, h->ev_h2d[0], 0));
  SCIMLOP_CUDA(h, cudaStreamWaitEvent(comm, h->ev_h2d[0], 0));
  const int ncclType = dtype == SCIMLOP_F64 ? NCCL_FLOAT64 : NCCL_FLOAT32;
  char* Wb = static_cast<char*>(W_full);
  // (Reserving SMs for the NCCL kernels next to the persistent GEMM CTAs was measured on 2 GPUs and does not help:
  // DESIGN.md section 6 -- the GEMMs use every SM.)
  for (long long c0 = 0; c0 < cols_per_rank; c0 += cc) {
    const long long nc = std::min<long long>(cc, cols_per_rank - c0);
    char* mine = Wb + ((size_t)h->rank * cols_per_rank + c0) * rows * es;
    int rc = scimlop_kron_mul_impl(h, dtype, precision, nfactors, factors, dims, lds, kinds,
                                   static_cast<const char*>(V) + (size_t)c0 * cols * es, cols, mine, rows, nc, alpha,
                                   nullptr, workspace, workspace_bytes, comp);
    if (rc) return rc;
    SCIMLOP_CUDA(h, cudaEventRecord(h->ev_comp[0], comp));
    SCIMLOP_CUDA(h, cudaStreamWaitEvent(comm, h->ev_comp[0], 0));
    if (gstart() != 0) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclGroupStart failed");
    for (int r = 0; r < h->nranks; r++) {
      char* slab = Wb + ((size_t)r * cols_per_rank + c0) * rows * es;
      int nr = bcast(slab, slab, (size_t)nc * rows, ncclType, r, h->nccl_comm, comm);
      if (nr != 0) { gend(); return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclBroadcast failed (%d)", nr); }
    }
    if (gend() != 0) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclGroupEnd failed");
  }
  // the caller's stream continues after the last exchange (which is ordered after the last compute)
  SCIMLOP_CUDA(h, cudaEventRecord(h->ev_d2h[0], comm));
  SCIMLOP_CUDA(h, cudaStreamWaitEvent(user, h->ev_d2h[0], 0));
  return SCIMLOP_OK;
}

// ---- peer memory (CUDA IPC) + stream barrier ---------------------------------------------------------------------
extern "C" int scimlop_peer_export(scimlop_handle_t h, const void* dev_ptr, void* handle64, int64_t* offset) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!dev_ptr || !handle64 || !offset) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "peer_export: null argument");
  scimlop_device_guard guard(h->device);
  // IPC handles name whole allocations: find the base of the allocation this pointer lives in
  typedef CUresult (*get_range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
    return scimlop_fail(h, SCIMLOP_ERR_CUDA, "cuMemGetAddressRange not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (reinterpret_cast<get_range_fn>(fn)(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS)
    return scimlop_fail(h, SCIMLOP_ERR_CUDA, "cuMemGetAddressRange failed");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  SCIMLOP_CUDA(h, cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), (void*)base));
  *offset = (int64_t)((CUdeviceptr)dev_ptr - base);
  return SCIMLOP_OK;
}

extern "C" int scimlop_peer_open(scimlop_handle_t h, const void* handle64, int64_t offset, void** peer_ptr) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!handle64 || !peer_ptr) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "peer_open: null argument");
  scimlop_device_guard guard(h->device);
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, sizeof(hd));
  void* base = nullptr;
  SCIMLOP_CUDA(h, cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = static_cast<char*>(base) + offset;
  return SCIMLOP_OK;
}

extern "C" int scimlop_peer_close(scimlop_handle_t h, void* peer_ptr, int64_t offset) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!peer_ptr) return SCIMLOP_OK;
  scimlop_device_guard guard(h->device);
  SCIMLOP_CUDA(h, cudaIpcCloseMemHandle(static_cast<char*>(peer_ptr) - offset));
  return SCIMLOP_OK;
}

extern "C" int scimlop_comm_barrier(scimlop_handle_t h, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!h->nccl_comm) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "comm_barrier: call scimlop_comm_init first");
  auto ar = reinterpret_cast<nccl_all_reduce_fn>(nccl_sym("ncclAllReduce"));
  if (!ar) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclAllReduce not found");
  scimlop_device_guard guard(h->device);
  if (!h->comm_scratch) SCIMLOP_CUDA(h, cudaMalloc(&h->comm_scratch, 256));
  int r = ar(h->comm_scratch, h->comm_scratch, 1, NCCL_FLOAT32, /*ncclSum*/ 0, h->nccl_comm, static_cast<cudaStream_t>(stream));
  if (r != 0) return scimlop_fail(h, SCIMLOP_ERR_NCCL, "ncclAllReduce (barrier) failed (%d)", r);
  return SCIMLOP_OK;
}

extern "C" int scimlop_destroy(scimlop_handle_t h) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  scimlop_device_guard guard(h->device);
  scimlop_comm_destroy(h);
  if (h->host_streams_ready) {
    cudaStreamSynchronize(h->s_h2d); cudaStreamSynchronize(h->s_comp); cudaStreamSynchronize(h->s_d2h);
    for (int i = 0; i < SCIMLOP_HOST_SLOTS; i++) {
      cudaEventDestroy(h->ev_h2d[i]); cudaEventDestroy(h->ev_comp[i]); cudaEventDestroy(h->ev_d2h[i]);
    }
    cudaStreamDestroy(h->s_h2d); cudaStreamDestroy(h->s_comp); cudaStreamDestroy(h->s_d2h);
  }
  if (h->pipe_ready) {
    cudaStreamSynchronize(h->s_side);
    for (int i = 0; i < 10; i++) cudaEventDestroy(h->ev_pipe[i]);
    cudaStreamDestroy(h->s_side);
  }
  if (h->stage) cudaFree(h->stage);
  if (h->comm_scratch) cudaFree(h->comm_scratch);
  if (h->split_scratch) { cudaDeviceSynchronize(); cudaFree(h->split_scratch); }
  h->magic = 0;
  delete h;
  return SCIMLOP_OK;
}

// ================================================================================================
// solver-loop vector kernels
// ================================================================================================
namespace {
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
int col_dot_t(scimlop_context* h, int64_t n, int64_t k, const T* X, int64_t ldx, const T* Y, int64_t ldy, T* out,
              cudaStream_t st) {
  constexpr int NV = scimlop::kry::V16<T>::N;
  const bool vec = (n % NV == 0) && (ldx % NV == 0) && (ldy % NV == 0) && al16(X) && al16(Y);
  const long long chunk = (long long)scimlop::kry::THREADS * (vec ? NV : 1) * scimlop::kry::U;
  long long bx = std::max<long long>(1, std::min<long long>((n + chunk - 1) / chunk, (long long)h->sm_count * 8 / std::max<int64_t>(1, std::min<int64_t>(k, 8))));
  SCIMLOP_CUDA(h, cudaMemsetAsync(out, 0, (size_t)k * sizeof(T), st));
  for (int64_t j0 = 0; j0 < k; j0 += 65535) {
    const unsigned ky = (unsigned)std::min<int64_t>(65535, k - j0);
    dim3 grid((unsigned)bx, ky);
    if (vec) scimlop::kry::col_dot_kernel<T, true><<<grid, scimlop::kry::THREADS, 0, st>>>(n, X + j0 * ldx, ldx, Y + j0 * ldy, ldy, out + j0);
    else scimlop::kry::col_dot_kernel<T, false><<<grid, scimlop::kry::THREADS, 0, st>>>(n, X + j0 * ldx, ldx, Y + j0 * ldy, ldy, out + j0);
    h->launches++;
  }
  SCIMLOP_CUDA(h, cudaGetLastError());
  return SCIMLOP_OK;
}

template <typename T>
int col_axpby_t(scimlop_context* h, int64_t n, int64_t k, const T* X, int64_t ldx, double sa, const T* an, const T* ad,
                T* W, int64_t ldw, double sb, const T* bn, const T* bd, T* dot, cudaStream_t st) {
  constexpr int NV = scimlop::kry::V16<T>::N;
  const bool vec = (n % NV == 0) && (ldx % NV == 0) && (ldw % NV == 0) && al16(X) && al16(W);
  const long long chunk = (long long)scimlop::kry::THREADS * (vec ? NV : 1) * scimlop::kry::U;
  long long bx = std::max<long long>(1, std::min<long long>((n + chunk - 1) / chunk, (long long)h->sm_count * 8 / std::max<int64_t>(1, std::min<int64_t>(k, 8))));
  if (dot) SCIMLOP_CUDA(h, cudaMemsetAsync(dot, 0, (size_t)k * sizeof(T), st));
  for (int64_t j0 = 0; j0 < k; j0 += 65535) {
    const unsigned ky = (unsigned)std::min<int64_t>(65535, k - j0);
    scimlop::kry::AxpbyParams<T> p;
    p.n = n; p.X = X + j0 * ldx; p.ldx = ldx; p.W = W + j0 * ldw; p.ldw = ldw;
    p.sa = (T)sa; p.sb = (T)sb;
    p.an = an ? an + j0 : nullptr; p.ad = ad ? ad + j0 : nullptr; p.bn = bn ? bn + j0 : nullptr; p.bd = bd ? bd + j0 : nullptr;
    p.dot = dot ? dot + j0 : nullptr;
    p.read_w = !(sb == 0.0 && !bn && !bd);
    dim3 grid((unsigned)bx, ky);
    if (vec) scimlop::kry::col_axpby_dot_kernel<T, true><<<grid, scimlop::kry::THREADS, 0, st>>>(p);
    else scimlop::kry::col_axpby_dot_kernel<T, false><<<grid, scimlop::kry::THREADS, 0, st>>>(p);
    h->launches++;
  }
  SCIMLOP_CUDA(h, cudaGetLastError());
  return SCIMLOP_OK;
}
}  // namespace

extern "C" int scimlop_col_dot(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* X, int64_t ldx,
                               const void* Y, int64_t ldy, void* out_dev, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
  if (n < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "col_dot: negative size");
  if (k == 0) return SCIMLOP_OK;
  if (!out_dev || (n > 0 && (!X || !Y))) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "col_dot: null operand");
  if (n > 0 && (ldx < n || ldy < n)) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "col_dot: leading dimension smaller than n");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64) return col_dot_t<double>(h, n, k, (const double*)X, ldx, (const double*)Y, ldy, (double*)out_dev, st);
  return col_dot_t<float>(h, n, k, (const float*)X, ldx, (const float*)Y, ldy, (float*)out_dev, st);
}

extern "C" int scimlop_col_axpby_dot(scimlop_handle_t h, int dtype, int64_t n, int64_t k, const void* X, int64_t ldx,
                                     double sa, const void* an_dev, const void* ad_dev, void* W, int64_t ldw, double sb,
                                     const void* bn_dev, const void* bd_dev, void* dot_dev, void* stream) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (dtype != SCIMLOP_F64 && dtype != SCIMLOP_F32) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
  if (n < 0 || k < 0) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "col_axpby: negative size");
  if (n == 0 || k == 0) return SCIMLOP_OK;
  if (!X || !W) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "col_axpby: null operand");
  if (ldx < n || ldw < n) return scimlop_fail(h, SCIMLOP_ERR_BAD_SHAPE, "col_axpby: leading dimension smaller than n");
  scimlop_device_guard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SCIMLOP_F64)
    return col_axpby_t<double>(h, n, k, (const double*)X, ldx, sa, (const double*)an_dev, (const double*)ad_dev, (double*)W,
                               ldw, sb, (const double*)bn_dev, (const double*)bd_dev, (double*)dot_dev, st);
  return col_axpby_t<float>(h, n, k, (const float*)X, ldx, sa, (const float*)an_dev, (const float*)ad_dev, (float*)W, ldw,
                            sb, (const float*)bn_dev, (const float*)bd_dev, (float*)dot_dev, st);
}
