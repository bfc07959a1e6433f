// context.h -- library handle and error plumbing shared by the translation units of libscimlop_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/scimlop_b200.h"

typedef CUresult (*scimlop_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                            const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int SCIMLOP_HOST_SLOTS = 3;
constexpr int SCIMLOP_MAX_SPLITS = 128;

// A Float32 matrix whose hi / lo halves (3xTF32 split) the caller prepared once (scimlop_split_attach): the streamed
// tcgen05 GEMM uses them instead of splitting the operand on every call.
struct scimlop_split_entry {
  const void* src;        // the matrix this is the split of (device pointer, the lookup key)
  long long ld, rows, cols;   // its stored layout
  float* hi;              // rows x cols, leading dimension ld_split; lo follows at hi + ld_split * cols
  float* lo;
  long long ld_split;
};

struct scimlop_context {
  uint32_t magic;
  int device;
  int sm_count;
  int sm_reserve;  // SMs left free by the persistent GEMM kernels (for concurrently running NCCL kernels)
  int peer_pipeline;  // scimlop_set_option(SCIMLOP_OPT_PEER_PIPELINE): column-chunk pipeline of the fused all-gather entries (default 1)
  int grid_limit;     // != 0: persistent GEMM kernels launch at most this many CTAs (set per stage by that pipeline)
  int mode_pipeline;   // SCIMLOP_OPT_MODE_PIPELINE: column chunks of the two-pass Float32 three-factor pipeline (0 = off)
  cudaStream_t split_stream;  // != nullptr: kron_mul_t continues on this stream after its first (fused two-mode) pass
  int peer_chunks, peer_stage2_sms;  // SCIMLOP_OPT_PEER_CHUNKS / _STAGE2_SMS: explicit pipeline shape (0 = chosen by the library)
  int no_small_splitk; // != 0: a GEMM with fewer tiles than SMs still runs the plain persistent kernel (chunks of that pipeline keep the summation order of the whole problem)
  int stage_limit[2]; // grid limits of the first / the last stage of a kron_mul call (0 = none)
  cudaEvent_t stage_event;   // != nullptr: recorded after the first stage of a kron_mul call
  cudaStream_t s_side;       // second stream + events of the pipeline (created on first use)
  cudaEvent_t ev_pipe[10];
  bool pipe_ready;
  int tail_split;  // scimlop_set_option(SCIMLOP_OPT_TAIL_SPLIT): run the last partial wave of a many-tile FP64 GEMM split-K (default 1)
  scimlop_encode_tiled_fn encode_tiled;
  // Float32 streamed GEMM: hi / lo images attached by the caller (cached operators: nothing is split or allocated per
  // call), and a grow-only scratch for the UNCACHED convenience entries (scimlop_matmul & co. on unattached operands)
  scimlop_split_entry splits[SCIMLOP_MAX_SPLITS];
  int nsplits;
  void* split_scratch;
  size_t split_scratch_bytes;
  int dmma_max_clusters[9];  // co-resident clusters of the split-K DMMA kernel per cluster size (occupancy query)
  char last_error[512];
  int64_t launches;

  // ---- host-streaming state (scimlop_kron_mul_host) ----
  cudaStream_t s_h2d, s_comp, s_d2h;
  cudaEvent_t ev_h2d[SCIMLOP_HOST_SLOTS], ev_comp[SCIMLOP_HOST_SLOTS], ev_d2h[SCIMLOP_HOST_SLOTS];
  bool host_streams_ready;
  void* stage;  // one device allocation carved into V/W slots, factors and workspace
  size_t stage_bytes;

  // ---- NCCL (loaded with dlopen on first use) ----
  void* nccl_lib;
  void* nccl_comm;
  int nranks, rank;
  void* comm_scratch;  // one word on the device for scimlop_comm_barrier
};

constexpr uint32_t SCIMLOP_MAGIC = 0x5C1B200u;

inline int scimlop_fail(scimlop_context* c, int status, const char* fmt, ...) {
  if (c) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->last_error, sizeof(c->last_error), fmt, ap);
    va_end(ap);
  }
  return status;
}

#define SCIMLOP_CUDA(c, call)                                                                      \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return scimlop_fail((c), SCIMLOP_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                   \
                          cudaGetErrorString(e__), __FILE__, __LINE__);                            \
  } while (0)

// RAII: make the handle's device current for the duration of a call.
struct scimlop_device_guard {
  int prev = -1;
  bool switched = false;
  explicit scimlop_device_guard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~scimlop_device_guard() {
    if (switched) cudaSetDevice(prev);
  }
};

inline bool scimlop_valid(scimlop_context* c) { return c && c->magic == SCIMLOP_MAGIC; }

// the handle's second stream + events (column-chunk pipeline of the fused all-gather, look-ahead of the LU): created on
// first use, destroyed with the handle
inline int scimlop_ensure_side_stream(scimlop_context* c) {
  if (c->pipe_ready) return SCIMLOP_OK;
  SCIMLOP_CUDA(c, cudaStreamCreateWithFlags(&c->s_side, cudaStreamNonBlocking));
  for (int i = 0; i < 10; i++) SCIMLOP_CUDA(c, cudaEventCreateWithFlags(&c->ev_pipe[i], cudaEventDisableTiming));
  c->pipe_ready = true;
  return SCIMLOP_OK;
}

// internal entry shared between scimlop.cu and extras.cu
int scimlop_kron_mul_impl(scimlop_context* c, int dtype, int precision, int nfactors,
                          const void* const* factors, const int64_t* dims, const int64_t* lds,
                          const int32_t* kinds, const void* V, int64_t ldv, void* W, int64_t ldw, int64_t k,
                          const void* alpha, const void* beta, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream);

// internal entry of sparse.cu used by the Kronecker planner (scimlop.cu): one mode product with a CSR factor,
// Y(l, i, r) = alpha * sum_j A(i, j) X(l, j, r) + beta * Y(l, i, r) over the (L x n x R) -> (L x m x R) views.
// `stage`: optional buffer of n * R elements for the innermost mode (L == 1): the row-contiguous copy of X.
int scimlop_csr_mode_impl(scimlop_context* c, int dtype, const void* csr_factor_desc, long long L, long long m, long long n,
                          long long R, const void* X, void* Y, const void* alpha, const void* beta, void* stage,
                          size_t stage_bytes, cudaStream_t stream);
