// bgc_capi.cu — the C ABI of include/bgc_b200.h: context, parameter upload,
// the two memory spaces (host Fortran layout / device SoA) and the launches.
//
// There is NO CPU fallback anywhere in this file: without a CUDA device every
// compute entry point fails with BGC_ERR_NO_DEVICE / BGC_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bgc_b200.h"
#include "bgc_kernels.cuh"

namespace bgc {
cudaError_t upload_bgc_tables_eco(const BgcTables &t, cudaStream_t s);
cudaError_t upload_bgc_tables_co3(const BgcTables &t, cudaStream_t s);
}

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(BGC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define RC(call) do { int rc_ = (call); if (rc_ != BGC_OK) return rc_; } while (0)

extern "C" const char *bgc_last_error(void) { return g_err; }
extern "C" const char *bgc_kernel_name(int kernel_id);
extern "C" const char *bgc_version(void) { return "ocean-bgc_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------ minimal NCCL binding (dlopen)
// Only ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy are
// needed, for one 64-element FP64 sum per step; binding them at run time keeps
// the library loadable on machines without NCCL.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
static int nccl_load() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.h) return BGC_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return fail(BGC_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(ncclUniqueId_ *))dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId_, int))dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    dlclose(g_nccl.h);
    g_nccl = NcclApi();   // a later call tries again instead of calling through null pointers
    return fail(BGC_ERR_NCCL, "libnccl is missing required symbols");
  }
  return BGC_OK;
}
#define NC(call)                                                                         \
  do {                                                                                   \
    int r_ = (call);                                                                     \
    if (r_ != 0)                                                                         \
      return fail(BGC_ERR_NCCL, "%s failed: %s", #call,                                  \
                  g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error");     \
  } while (0)
enum { kNcclFloat64 = 8, kNcclSum = 0 };

// ------------------------------------------------------------------ context
struct DevBuf { void *p = nullptr; size_t bytes = 0; };
struct HostStage { void *p = nullptr; size_t bytes = 0; };   // page-locked host staging

struct bgc_ctx {
  int device = 0;
  int nL = 0, nC = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t pipe_stream = nullptr;       // second slot of the host-layout pipeline
  cudaEvent_t pipe_event = nullptr;
  // bgc_source_sink runs the carbonate kernel (+ saturation-depth scan) beside the column
  // sweep: one side stream per pipeline slot, forked from / joined to the slot's stream
  cudaStream_t side_stream[2] = {nullptr, nullptr};
  cudaEvent_t fork_event[2] = {nullptr, nullptr}, join_event[2] = {nullptr, nullptr};
  int concurrent_co3 = 1;                   // bgc_ctx_set_concurrency / BGC_CONCURRENT_CO3 (0, 1, 2, 3: see source_sink_device)
  int co3_confined = 0;                     // BGC_CO3_CONFINED=1: confine the carbonate kernel to a sub-wave sweep's idle SMs (measured: does not pay)
  int co3_share_pct = 0;                    // BGC_CO3_SHARE: percent of the modelled share (0 = default)
  unsigned long long *d_block_trace = nullptr;   // BGC_BLOCK_TRACE_FILE (debugging aid, bgc_kernels.cuh)
  int co3_pblocks = 0, co3_after = 0;       // experiments: BGC_CO3_PBLOCKS (confined blocks), BGC_CO3_AFTER (launched after the sweep)
  int sm_count = 148;
  int zero_shortcut = 1;                    // bgc_ctx_set_zero_shortcut / BGC_ZERO_SHORTCUT
  bool diag_accumulate = false;             // bgc_diag_accumulate_enable
  bool defer_join = false;                  // bgc_ctx_set_deferred_join
  bool pending_join = false;                // a carbonate side stream has not been joined to the ctx stream yet
  int host_chunk_columns = 0;               // columns per pipeline chunk (0 = automatic; BGC_HOST_CHUNK_COLUMNS)
  bgc::BgcTables bgc_tab;
  bgc::DmsTables dms_tab;
  bgc::MacrosTables macros_tab;
  bool have_bgc = false, have_dms = false, have_macros = false;
  unsigned long long *d_status = nullptr;   // 4 counters
  double *d_inventory = nullptr;            // BGC_INVENTORY_LEN
  bool inventory_on = false;
  double *h_inventory = nullptr;            // page-locked landing buffer of bgc_inventory_allreduce_begin
  bool inventory_pending = false;
  bool capturing = false;                   // between bgc_graph_capture_begin and _end
  bool cap_uses_bgc = false, cap_uses_dms = false, cap_uses_macros = false;   // tables the captured calls read
  unsigned long long capture_base[BGC_KERNEL_ID_COUNT] = {0};
  int eco_variant = 0;                      // launch shape of the column sweep (BGC_ECO_VARIANT, tuning only)
  int dms_variant = 0;                      // launch shape of the DMS tile kernel (BGC_DMS_VARIANT, tuning only)
  std::map<std::string, DevBuf> arena;      // persistent device buffers (host-layout mode, scratch)
  std::map<std::string, HostStage> host_stage;   // page-locked staging buffers of the host-layout calls
  struct ZeroJob { char *p; size_t bytes; };
  unsigned long long h2d_bytes = 0, d2h_bytes = 0;   // what the host-layout calls have copied (bgc_transfer_bytes)
  std::vector<ZeroJob> zero_jobs;                // host ranges a host-layout call fills with zeros itself (see down_k)
  // bottom-cell values of the sediment diagnostics, scattered into the zero-filled host arrays once the call's
  // transfers are complete: host(kmax(col) - 1, col) = vec[col] for col in [c0, c0 + cc)
  struct ScatterJob { double *host; const double *vec; const int *kmax; int nL, c0, cc, nColumns; };
  std::vector<ScatterJob> scatter_jobs;
  ncclComm_t comm = nullptr;
  int nranks = 1;
  // launch accounting (always on) and optional per-kernel CUDA-event timing
  unsigned long long launches[BGC_KERNEL_ID_COUNT] = {0};
  bool timing_on = false;
  struct Span { int kid; cudaEvent_t a, b; };
  std::vector<Span> spans;          // recorded, not yet resolved
  std::vector<cudaEvent_t> ev_pool;
  double timed_ms[BGC_KERNEL_ID_COUNT] = {0};
  unsigned long long timed_launches[BGC_KERNEL_ID_COUNT] = {0};
};

// Every kernel launch of the library goes through LAUNCH: it counts `n` launches
// under kernel id `kid` and, when timing is enabled, brackets them with CUDA
// events on the ctx stream (resolved lazily in bgc_timing_get).
static int span_begin(bgc_ctx *c, int kid, cudaEvent_t *a, cudaEvent_t *b) {
  auto get = [&](cudaEvent_t *e) -> int {
    if (!c->ev_pool.empty()) { *e = c->ev_pool.back(); c->ev_pool.pop_back(); return BGC_OK; }
    cudaError_t r = cudaEventCreate(e);
    if (r != cudaSuccess) return BGC_ERR_CUDA;
    return BGC_OK;
  };
  if (get(a) != BGC_OK || get(b) != BGC_OK) return BGC_ERR_CUDA;
  (void)kid;
  return cudaEventRecord(*a, c->stream) == cudaSuccess ? BGC_OK : BGC_ERR_CUDA;
}
#define LAUNCH(kid, n, call)                                                                  \
  do {                                                                                        \
    cudaEvent_t ea_ = nullptr, eb_ = nullptr;                                                 \
    if (c->timing_on && span_begin(c, (kid), &ea_, &eb_) != BGC_OK)                           \
      return fail(BGC_ERR_CUDA, "cannot create timing events");                               \
    CU(call);                                                                                 \
    c->launches[(kid)] += (n);                                                                \
    if (c->timing_on) {                                                                       \
      CU(cudaEventRecord(eb_, c->stream));                                                    \
      c->spans.push_back({(kid), ea_, eb_});                                                  \
      c->timed_launches[(kid)] += (n);                                                        \
    }                                                                                         \
  } while (0)

// Which ctx's tables currently sit in each device's __constant__ memory.
static std::map<int, std::pair<const bgc_ctx *, unsigned long long>> g_const_owner_bgc, g_const_owner_dms,
    g_const_owner_macros;
static unsigned long long g_version_counter = 1;
static std::mutex g_mu;   // guards the maps below and g_version_counter
struct CtxVersions { unsigned long long bgc = 0, dms = 0, macros = 0; };
static std::map<const bgc_ctx *, CtxVersions> g_versions;

static int use_device(bgc_ctx *c) {
  if (!c) return fail(BGC_ERR_ARG, "null ctx");
  CU(cudaSetDevice(c->device));
  return BGC_OK;
}

// Join point of the deferred carbonate join: the ctx stream waits for the side stream.
static int join_pending(bgc_ctx *c) {
  if (c->pending_join) {
    CU(cudaStreamWaitEvent(c->stream, c->join_event[0], 0));
    c->pending_join = false;
  }
  return BGC_OK;
}

static int arena_get(bgc_ctx *c, const std::string &key, size_t bytes, void **out) {
  DevBuf &b = c->arena[key];
  if (b.bytes < bytes && c->capturing)
    return fail(BGC_ERR_ARG, "device arena would grow during graph capture (%s): run the same calls once before capturing", key.c_str());
  if (b.bytes < bytes) {
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.bytes = 0;
    CU(cudaMalloc(&b.p, bytes ? bytes : 8));
    b.bytes = bytes;
  }
  *out = b.p;
  return BGC_OK;
}
static int host_stage_get(bgc_ctx *c, const std::string &key, size_t bytes, void **out) {
  HostStage &hs = c->host_stage[key];
  if (hs.bytes < bytes) {
    if (hs.p) cudaFreeHost(hs.p);
    hs.p = nullptr; hs.bytes = 0;
    CU(cudaHostAlloc(&hs.p, bytes, cudaHostAllocDefault));
    hs.bytes = bytes;
  }
  *out = hs.p;
  return BGC_OK;
}

static int arena_d(bgc_ctx *c, const std::string &key, size_t n, double **out) {
  void *p = nullptr;
  RC(arena_get(c, key, n * sizeof(double), &p));
  *out = (double *)p;
  return BGC_OK;
}

static int ctx_init(bgc_ctx *c, int device, int nLevelsMax, int nColumnsMax);
extern "C" int bgc_ctx_destroy(bgc_ctx *c);

extern "C" int bgc_ctx_create(int device, int nLevelsMax, int nColumnsMax, bgc_ctx **out) {
  if (!out || nLevelsMax < 1 || nColumnsMax < 1) return fail(BGC_ERR_ARG, "bgc_ctx_create: bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return fail(BGC_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev) return fail(BGC_ERR_ARG, "device %d out of range [0,%d)", device, ndev);
  CU(cudaSetDevice(device));
  bgc_ctx *c = new bgc_ctx();
  { std::lock_guard<std::mutex> lock(g_mu); g_versions[c] = CtxVersions(); }
  const int rc = ctx_init(c, device, nLevelsMax, nColumnsMax);
  if (rc != BGC_OK) { bgc_ctx_destroy(c); return rc; }   // releases whatever was created before the failure
  *out = c;
  return BGC_OK;
}

static int ctx_init(bgc_ctx *c, int device, int nLevelsMax, int nColumnsMax) {
  c->device = device;
  c->nL = nLevelsMax;
  c->nC = nColumnsMax;
  if (const char *v = getenv("BGC_ECO_VARIANT")) c->eco_variant = atoi(v);
  if (const char *v = getenv("BGC_DMS_VARIANT")) c->dms_variant = atoi(v);
  if (const char *v = getenv("BGC_HOST_CHUNK_COLUMNS")) c->host_chunk_columns = atoi(v);
  CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->pipe_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->pipe_event, cudaEventDisableTiming));
  if (const char *v = getenv("BGC_CONCURRENT_CO3")) c->concurrent_co3 = atoi(v);
  if (const char *v = getenv("BGC_CO3_CONFINED")) c->co3_confined = atoi(v);
  if (const char *v = getenv("BGC_CO3_SHARE")) c->co3_share_pct = atoi(v);
  if (const char *v = getenv("BGC_CO3_PBLOCKS")) c->co3_pblocks = atoi(v);
  if (getenv("BGC_BLOCK_TRACE_FILE")) {
    const size_t bytes = (4 + 4 * (size_t)bgc::kBlockTraceCap) * sizeof(unsigned long long);
    CU(cudaMalloc(&c->d_block_trace, bytes));
    CU(cudaMemset(c->d_block_trace, 0, bytes));
  }
  if (const char *v = getenv("BGC_CO3_AFTER")) c->co3_after = atoi(v);
  CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  if (const char *v = getenv("BGC_ZERO_SHORTCUT")) c->zero_shortcut = atoi(v);
  {
    int lo = 0, hi = 0;   // the side stream gets the LOWER priority: the sweep's blocks are placed first
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    if (const char *v = getenv("BGC_SIDE_PRIORITY")) lo = atoi(v);   // tuning only
    for (int i = 0; i < 2; ++i) {
      CU(cudaStreamCreateWithPriority(&c->side_stream[i], cudaStreamNonBlocking, lo));
      CU(cudaEventCreateWithFlags(&c->fork_event[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->join_event[i], cudaEventDisableTiming));
    }
  }
  c->stream = c->own_stream;
  CU(cudaMalloc(&c->d_status, 4 * sizeof(unsigned long long)));
  CU(cudaMemset(c->d_status, 0, 4 * sizeof(unsigned long long)));
  CU(cudaMalloc(&c->d_inventory, BGC_INVENTORY_LEN * sizeof(double)));
  CU(cudaMemset(c->d_inventory, 0, BGC_INVENTORY_LEN * sizeof(double)));
  return BGC_OK;
}

extern "C" int bgc_ctx_destroy(bgc_ctx *c) {
  if (!c) return BGC_OK;
  cudaSetDevice(c->device);
  // the streams may still carry the all-reduce: drain them before the communicator goes
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->own_stream && c->own_stream != c->stream) cudaStreamSynchronize(c->own_stream);
  for (int i = 0; i < 2; ++i) if (c->side_stream[i]) cudaStreamSynchronize(c->side_stream[i]);
  if (c->pipe_stream) cudaStreamSynchronize(c->pipe_stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  c->comm = nullptr;
  if (c->d_block_trace) {   // debugging aid: one line per traced thread block
    std::vector<unsigned long long> h(4 + 4 * (size_t)bgc::kBlockTraceCap);
    const char *path = getenv("BGC_BLOCK_TRACE_FILE");
    if (path && cudaMemcpy(h.data(), c->d_block_trace, h.size() * sizeof(h[0]), cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (FILE *f = fopen(path, "a")) {
        const size_t n = h[0] < bgc::kBlockTraceCap ? (size_t)h[0] : (size_t)bgc::kBlockTraceCap;
        for (size_t r = 0; r < n; ++r)
          fprintf(f, "%u %u %llu %llu %llu\n", (unsigned)(h[4 + 4 * r] >> 32), (unsigned)(h[4 + 4 * r] & 0xffffffffu),
                  h[5 + 4 * r], h[6 + 4 * r], h[7 + 4 * r]);
        fclose(f);
      }
    }
    cudaFree(c->d_block_trace);
  }
  for (auto &kv : c->arena) if (kv.second.p) cudaFree(kv.second.p);
  for (auto &kv : c->host_stage) if (kv.second.p) cudaFreeHost(kv.second.p);
  cudaFree(c->d_status); cudaFree(c->d_inventory);
  if (c->h_inventory) cudaFreeHost(c->h_inventory);
  for (auto &sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->pipe_stream) cudaStreamDestroy(c->pipe_stream);
  if (c->pipe_event) cudaEventDestroy(c->pipe_event);
  for (int i = 0; i < 2; ++i) {
    if (c->side_stream[i]) cudaStreamDestroy(c->side_stream[i]);
    if (c->fork_event[i]) cudaEventDestroy(c->fork_event[i]);
    if (c->join_event[i]) cudaEventDestroy(c->join_event[i]);
  }
  std::lock_guard<std::mutex> lock(g_mu);
  g_versions.erase(c);
  for (auto *m : {&g_const_owner_bgc, &g_const_owner_dms, &g_const_owner_macros}) {
    auto it = m->find(c->device);
    if (it != m->end() && it->second.first == c) m->erase(it);
  }
  delete c;
  return BGC_OK;
}

extern "C" int bgc_ctx_set_stream(bgc_ctx *c, void *cuda_stream) {
  if (!c) return fail(BGC_ERR_ARG, "null ctx");
  RC(join_pending(c));
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return BGC_OK;
}

extern "C" int bgc_ctx_set_deferred_join(bgc_ctx *c, int enable) {
  RC(use_device(c));
  if (!enable) RC(join_pending(c));
  c->defer_join = enable != 0;
  return BGC_OK;
}

extern "C" int bgc_ctx_set_concurrency(bgc_ctx *c, int enable) {
  RC(use_device(c));
  RC(join_pending(c));
  // 0 = same stream, 1 = side stream, placement chosen by the sweep's size (the default),
  // 2 = side stream after the sweep, 3 = confined to the sweep's idle SMs where it has any, else 2
  c->concurrent_co3 = enable;
  return BGC_OK;
}

extern "C" int bgc_transfer_bytes(bgc_ctx *c, unsigned long long bytes[2], int reset) {
  if (!c || !bytes) return fail(BGC_ERR_ARG, "bgc_transfer_bytes: null argument");
  bytes[0] = c->h2d_bytes; bytes[1] = c->d2h_bytes;
  if (reset) c->h2d_bytes = c->d2h_bytes = 0;
  return BGC_OK;
}

extern "C" int bgc_ctx_set_zero_shortcut(bgc_ctx *c, int enable) {
  if (!c) return fail(BGC_ERR_ARG, "null ctx");
  c->zero_shortcut = enable != 0;
  return BGC_OK;
}

extern "C" int bgc_carbonate_join(bgc_ctx *c) {
  RC(use_device(c));
  return join_pending(c);
}

extern "C" int bgc_ctx_synchronize(bgc_ctx *c) {
  RC(use_device(c));
  RC(join_pending(c));
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

extern "C" int bgc_get_status(bgc_ctx *c, BgcStatus *out, int reset) {
  RC(use_device(c));
  if (!out) return fail(BGC_ERR_ARG, "null out");
  unsigned long long h[4];
  RC(join_pending(c));
  CU(cudaMemcpyAsync(h, c->d_status, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  if (reset) CU(cudaMemsetAsync(c->d_status, 0, sizeof h, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  out->no_bracket = h[0]; out->no_convergence = h[1]; out->poc_error = h[2]; out->nonfinite = h[3];
  return BGC_OK;
}

// ------------------------------------------------------------------ launch accounting / timing
static int resolve_spans(bgc_ctx *c) {
  if (c->spans.empty()) return BGC_OK;
  RC(join_pending(c));
  CU(cudaStreamSynchronize(c->stream));
  // BGC_TRACE_FILE=<path> (debugging / tuning): one line per timed launch, start and end in ms
  // relative to the first launch of the batch - a poor man's timeline of the streams.
  FILE *trace = nullptr;
  if (const char *tf = getenv("BGC_TRACE_FILE")) trace = fopen(tf, "a");
  for (auto &sp : c->spans) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, sp.a, sp.b));
    c->timed_ms[sp.kid] += (double)ms;
    if (trace) {
      float t0 = 0.f;
      cudaEventElapsedTime(&t0, c->spans.front().a, sp.a);
      fprintf(trace, "%-24s %10.4f %10.4f\n", bgc_kernel_name(sp.kid), (double)t0, (double)(t0 + ms));
    }
  }
  if (trace) { fprintf(trace, "--\n"); fclose(trace); }
  for (auto &sp : c->spans) {
    c->ev_pool.push_back(sp.a);
    c->ev_pool.push_back(sp.b);
  }
  c->spans.clear();
  return BGC_OK;
}

extern "C" int bgc_timing_enable(bgc_ctx *c, int enable) {
  RC(use_device(c));
  RC(resolve_spans(c));
  c->timing_on = enable != 0;
  return BGC_OK;
}
extern "C" int bgc_timing_reset(bgc_ctx *c) {
  RC(use_device(c));
  RC(resolve_spans(c));
  for (int i = 0; i < BGC_KERNEL_ID_COUNT; ++i) { c->timed_ms[i] = 0.0; c->timed_launches[i] = 0; c->launches[i] = 0; }
  return BGC_OK;
}
extern "C" int bgc_timing_get(bgc_ctx *c, int kernel_id, double *total_ms, unsigned long long *timed_launches,
                              unsigned long long *launches) {
  RC(use_device(c));
  if (kernel_id < 0 || kernel_id >= BGC_KERNEL_ID_COUNT) return fail(BGC_ERR_ARG, "bad kernel id %d", kernel_id);
  RC(resolve_spans(c));
  if (total_ms) *total_ms = c->timed_ms[kernel_id];
  if (timed_launches) *timed_launches = c->timed_launches[kernel_id];
  if (launches) *launches = c->launches[kernel_id];
  return BGC_OK;
}
extern "C" const char *bgc_kernel_name(int kernel_id) {
  static const char *names[BGC_KERNEL_ID_COUNT] = {
      "co3_cells_kernel", "eco_columns_kernel", "dms_columns_kernel", "macros_cells_kernel",
      "surface_fluxes_kernel", "dms_surface_kernel", "co2calc_points_kernel", "inventory kernels",
      "transpose_kernel", "zsat_columns_kernel", "accumulate_kernel"};
  return (kernel_id >= 0 && kernel_id < BGC_KERNEL_ID_COUNT) ? names[kernel_id] : "";
}

// ------------------------------------------------------------------ parameters
static bool index_ok(int v, int n) { return v >= 1 && v <= n; }

extern "C" int bgc_set_params(bgc_ctx *c, const BgcParams *p, const BgcAutotroph a[BGC_AUTOTROPH_CNT],
                              const BgcIndices *ind) {
  if (!c || !p || !a || !ind) return fail(BGC_ERR_ARG, "bgc_set_params: null argument");
  // every tracer slot must be written exactly once: the 16 plain tracers plus
  // the autotroph slots have to form a permutation of 1..30
  int seen[BGC_TRACER_CNT + 1] = {0};
  const int plain[16] = {ind->po4_ind, ind->no3_ind, ind->sio3_ind, ind->nh4_ind, ind->fe_ind, ind->o2_ind,
                         ind->dic_ind, ind->dic_alt_co2_ind, ind->alk_ind, ind->doc_ind, ind->don_ind,
                         ind->dofe_ind, ind->dop_ind, ind->dopr_ind, ind->donr_ind, ind->zooC_ind};
  int count = 0;
  for (int v : plain) {
    if (!index_ok(v, BGC_TRACER_CNT) || seen[v]++) return fail(BGC_ERR_ARG, "bgc_set_params: bad/duplicate tracer index %d", v);
    ++count;
  }
  for (int g = 0; g < BGC_AUTOTROPH_CNT; ++g) {
    const int req[3] = {a[g].Chl_ind, a[g].C_ind, a[g].Fe_ind};
    for (int v : req) {
      if (!index_ok(v, BGC_TRACER_CNT) || seen[v]++)
        return fail(BGC_ERR_ARG, "bgc_set_params: autotroph %d has bad/duplicate tracer index %d (call bgc_init first)", g + 1, v);
      ++count;
    }
    const int opt[2] = {a[g].Si_ind, a[g].CaCO3_ind};
    for (int v : opt) {
      if (v == 0) continue;
      if (!index_ok(v, BGC_TRACER_CNT) || seen[v]++)
        return fail(BGC_ERR_ARG, "bgc_set_params: autotroph %d has bad/duplicate optional tracer index %d", g + 1, v);
      ++count;
    }
    if (!index_ok(a[g].grazee_ind, BGC_AUTOTROPH_CNT)) return fail(BGC_ERR_ARG, "bgc_set_params: bad grazee_ind");
  }
  if (count != BGC_TRACER_CNT)
    return fail(BGC_ERR_ARG, "bgc_set_params: tracer indices cover %d of %d slots", count, BGC_TRACER_CNT);
  if (!index_ok(ind->diat_ind, BGC_AUTOTROPH_CNT)) return fail(BGC_ERR_ARG, "bgc_set_params: bad diat_ind");

  // The Fortran shim calls this before every BGC_SourceSink: identical tables keep their
  // version, so the __constant__ copy is not uploaded again.
  if (c->have_bgc && memcmp(&c->bgc_tab.p, p, sizeof *p) == 0 && memcmp(c->bgc_tab.a, a, sizeof(BgcAutotroph) * BGC_AUTOTROPH_CNT) == 0 &&
      memcmp(&c->bgc_tab.ind, ind, sizeof *ind) == 0)
    return BGC_OK;
  c->bgc_tab.p = *p;
  for (int g = 0; g < BGC_AUTOTROPH_CNT; ++g) c->bgc_tab.a[g] = a[g];
  c->bgc_tab.ind = *ind;
  for (int i = 0; i < BGC_AUTOTROPH_CNT; ++i)
    for (int j = 0; j < BGC_AUTOTROPH_CNT; ++j)
      c->bgc_tab.same_grazee[i][j] = (a[j].grazee_ind == a[i].grazee_ind) ? 1 : 0;
  c->have_bgc = true;
  { std::lock_guard<std::mutex> lock(g_mu); g_versions[c].bgc = g_version_counter++; }
  return BGC_OK;
}

extern "C" int dms_set_params(bgc_ctx *c, const DmsParams *p, const DmsIndices *ind) {
  if (!c || !p || !ind) return fail(BGC_ERR_ARG, "dms_set_params: null argument");
  const int *v = &ind->dms_ind;
  int seen[DMS_TRACER_CNT + 1] = {0};
  for (int i = 0; i < DMS_TRACER_CNT; ++i)
    if (!index_ok(v[i], DMS_TRACER_CNT) || seen[v[i]]++) return fail(BGC_ERR_ARG, "dms_set_params: bad/duplicate tracer index %d", v[i]);
  if (c->have_dms && memcmp(&c->dms_tab.p, p, sizeof *p) == 0 && memcmp(&c->dms_tab.ind, ind, sizeof *ind) == 0) return BGC_OK;
  c->dms_tab.p = *p;
  c->dms_tab.ind = *ind;
  c->have_dms = true;
  { std::lock_guard<std::mutex> lock(g_mu); g_versions[c].dms = g_version_counter++; }
  return BGC_OK;
}

extern "C" int macros_set_params(bgc_ctx *c, const MacrosParams *p, const MacrosIndices *ind) {
  if (!c || !p || !ind) return fail(BGC_ERR_ARG, "macros_set_params: null argument");
  const int *v = &ind->prot_ind;
  int seen[MACROS_TRACER_CNT + 1] = {0};
  for (int i = 0; i < MACROS_TRACER_CNT; ++i)
    if (!index_ok(v[i], MACROS_TRACER_CNT) || seen[v[i]]++) return fail(BGC_ERR_ARG, "macros_set_params: bad/duplicate tracer index %d", v[i]);
  if (c->have_macros && memcmp(&c->macros_tab.p, p, sizeof *p) == 0 && memcmp(&c->macros_tab.ind, ind, sizeof *ind) == 0)
    return BGC_OK;
  c->macros_tab.p = *p;
  c->macros_tab.ind = *ind;
  c->have_macros = true;
  { std::lock_guard<std::mutex> lock(g_mu); g_versions[c].macros = g_version_counter++; }
  return BGC_OK;
}

// __constant__ memory is one copy per device: re-upload only when another ctx
// (or a newer *_set_params) owns what is there now.
// Kernels of the previous owner (this ctx's side streams, or another ctx of the device) may still
// be reading the table: the device is drained before it is overwritten.  This happens only when
// *_set_params changed something or two ctxs share a device - never in a steady time loop.
static int drain_before_table_change(bgc_ctx *c, const char *what) {
  if (c->capturing)
    return fail(BGC_ERR_PARAMS, "%s parameter tables changed during graph capture: run the same calls once before capturing", what);
  CU(cudaDeviceSynchronize());
  return BGC_OK;
}

static int ensure_bgc_tables(bgc_ctx *c) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!c->have_bgc) return fail(BGC_ERR_PARAMS, "bgc_set_params has not been called on this ctx");
  if (c->capturing) c->cap_uses_bgc = true;
  auto &own = g_const_owner_bgc[c->device];
  const unsigned long long v = g_versions[c].bgc;
  if (own.first != c || own.second != v) {
    RC(drain_before_table_change(c, "BGC"));
    CU(bgc::upload_bgc_tables_eco(c->bgc_tab, c->stream));
    CU(bgc::upload_bgc_tables_co3(c->bgc_tab, c->stream));
    own = {c, v};
  }
  return BGC_OK;
}
static int ensure_dms_tables(bgc_ctx *c) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!c->have_dms) return fail(BGC_ERR_PARAMS, "dms_set_params has not been called on this ctx");
  if (c->capturing) c->cap_uses_dms = true;
  auto &own = g_const_owner_dms[c->device];
  const unsigned long long v = g_versions[c].dms;
  if (own.first != c || own.second != v) {
    RC(drain_before_table_change(c, "DMS"));
    CU(bgc::upload_dms_tables(c->dms_tab, c->stream));
    own = {c, v};
  }
  return BGC_OK;
}
static int ensure_macros_tables(bgc_ctx *c) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!c->have_macros) return fail(BGC_ERR_PARAMS, "macros_set_params has not been called on this ctx");
  if (c->capturing) c->cap_uses_macros = true;
  auto &own = g_const_owner_macros[c->device];
  const unsigned long long v = g_versions[c].macros;
  if (own.first != c || own.second != v) {
    RC(drain_before_table_change(c, "MACROS"));
    CU(bgc::upload_macros_tables(c->macros_tab, c->stream));
    own = {c, v};
  }
  return BGC_OK;
}

// ------------------------------------------------------------------ host-layout transport
// BGC_MEM_HOST_FORTRAN calls are PCIe-bound (EC60to30: 11 GB up, 25 GB down per step against
// 9 ms of kernels), so the transport is a two-slot pipeline over COLUMN CHUNKS: chunk i is
// uploaded, transposed, computed, transposed back and downloaded on stream (i & 1) with its
// own set of arena buffers, so the download of one chunk overlaps the upload and the kernels
// of the next (PCIe is full duplex).  In the reference layout A(k,col,n) the columns
// [c0, c0+cc) of slab n are one contiguous run of nL*cc doubles: a chunk of an array is one
// 2-D copy (one row per slab), and on the device the chunk is simply a mesh of cc columns.
struct HostChunk {
  int nL, nC;      // the caller's extents (host pitch)
  int c0, cc;      // first column and number of columns of this chunk
  int slot;        // pipeline slot: selects the stream and the arena buffer set
  // Batched transfers.  Copies and transposes of a chunk are issued in two passes so that the
  // copy engine never waits for a transpose between two arrays: uploads are copied into
  // consecutive slabs of the slot's staging area first and transposed afterwards
  // (flush_up); downloads are transposed into the staging area first and copied afterwards
  // (flush_down).
  struct Xfer { double *host; double *dev; int nSlabs; size_t stage_slab; };
  std::vector<Xfer> ups, downs;
  size_t up_slabs = 0, down_slabs = 0;   // staging slabs used so far
  size_t stage_slabs = 0;                // capacity of the slot's staging area, in (k,col) slabs
  double *stage = nullptr;
};

static std::string slot_key(const HostChunk &h, const char *key) { return std::string(key) + (h.slot ? "#1" : "#0"); }

// Reserve the slot's staging area: `slabs` (k,col) slabs of the chunk (the larger of what the
// call uploads and what it downloads).
static int stage_reserve(bgc_ctx *c, HostChunk &h, size_t slabs) {
  h.stage_slabs = slabs;
  return arena_d(c, slot_key(h, "stage"), (size_t)h.nL * h.cc * slabs, &h.stage);
}

// Upload the chunk of a (k,col,n) array (copy now, transpose in flush_up).  Slabs whose bit is
// set in skip_mask are not transferred (inputs no kernel reads); their device slab keeps stale data.
static int up_k(bgc_ctx *c, HostChunk &h, const char *key, const double *host, int nSlabs, double **dev_out,
                unsigned skip_mask = 0u) {
  const size_t n2c = (size_t)h.nL * h.cc, n2 = (size_t)h.nL * h.nC;
  double *dev = nullptr;
  RC(arena_d(c, slot_key(h, key), n2c * nSlabs, &dev));
  *dev_out = dev;
  if (!host) return fail(BGC_ERR_ARG, "null host array for %s", key);
  for (int s0 = 0; s0 < nSlabs;) {
    if (s0 < 32 && ((skip_mask >> s0) & 1u)) { ++s0; continue; }
    int ns = 1;
    while (s0 + ns < nSlabs && !((s0 + ns) < 32 && ((skip_mask >> (s0 + ns)) & 1u))) ++ns;
    if (h.up_slabs + ns > h.stage_slabs) return fail(BGC_ERR_ARG, "internal: staging area too small (%s)", key);
    double *st = h.stage + h.up_slabs * n2c;
    CU(cudaMemcpy2DAsync(st, n2c * sizeof(double), host + (size_t)s0 * n2 + (size_t)h.c0 * h.nL, n2 * sizeof(double),
                         n2c * sizeof(double), (size_t)ns, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += n2c * sizeof(double) * (size_t)ns;
    h.ups.push_back({nullptr, dev + (size_t)s0 * n2c, ns, h.up_slabs});
    h.up_slabs += ns;
    s0 += ns;
  }
  return BGC_OK;
}

static int flush_up(bgc_ctx *c, HostChunk &h) {
  const size_t n2c = (size_t)h.nL * h.cc;
  for (const auto &x : h.ups)
    LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_transpose(h.stage + x.stage_slab * n2c, x.dev, h.nL, h.cc, x.nSlabs, c->stream));
  h.ups.clear();
  return BGC_OK;
}

// Download the chunk of a (k,col,n) array (transpose now, copy in flush_down).  Slabs whose bit is
// set in zero_mask are STRUCTURALLY zero - the reference assigns them the constant zero whatever the
// inputs are (e.g. twelve of the fourteen DMS tendencies, DMS_mod.F90:413, :718-719) - and do not
// cross PCIe: the host range is zero-filled by host threads while the other slabs are in flight
// (run_zero_jobs).
static int down_k(bgc_ctx *c, HostChunk &h, const double *dev, double *host, int nSlabs, unsigned long long zero_mask = 0ull) {
  const size_t n2c = (size_t)h.nL * h.cc, n2 = (size_t)h.nL * h.nC;
  for (int s0 = 0; s0 < nSlabs;) {
    if (s0 < 64 && ((zero_mask >> s0) & 1ull)) {
      c->zero_jobs.push_back({(char *)(host + (size_t)s0 * n2 + (size_t)h.c0 * h.nL), n2c * sizeof(double)});
      ++s0;
      continue;
    }
    int ns = 1;
    while (s0 + ns < nSlabs && !((s0 + ns) < 64 && ((zero_mask >> (s0 + ns)) & 1ull))) ++ns;
    if (h.down_slabs + ns > h.stage_slabs) return fail(BGC_ERR_ARG, "internal: staging area too small (download)");
    LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_transpose(dev + (size_t)s0 * n2c, h.stage + h.down_slabs * n2c, h.cc, h.nL, ns, c->stream));
    h.downs.push_back({host + (size_t)s0 * n2, nullptr, ns, h.down_slabs});
    h.down_slabs += ns;
    s0 += ns;
  }
  return BGC_OK;
}

// The zero fills collected by down_k, spread over a few host threads (the calling thread is one of
// them); runs while the GPU and the copy engines work on what the call enqueued.
static void run_zero_jobs(bgc_ctx *c) {
  if (c->zero_jobs.empty()) return;
  std::vector<bgc_ctx::ZeroJob> jobs;
  jobs.swap(c->zero_jobs);
  size_t total = 0;
  for (const auto &j : jobs) total += j.bytes;
  unsigned nthreads = std::thread::hardware_concurrency() / 2;
  if (nthreads > 4) nthreads = 4;
  if (total < ((size_t)8 << 20) || nthreads < 2) nthreads = 1;
  std::atomic<size_t> next(0);
  auto work = [&]() {
    for (size_t i = next.fetch_add(1); i < jobs.size(); i = next.fetch_add(1)) memset(jobs[i].p, 0, jobs[i].bytes);
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(work);
  work();
  for (auto &th : pool) th.join();
}

static int flush_down(bgc_ctx *c, HostChunk &h) {
  const size_t n2c = (size_t)h.nL * h.cc, n2 = (size_t)h.nL * h.nC;
  for (const auto &x : h.downs)
    CU(cudaMemcpy2DAsync(x.host + (size_t)h.c0 * h.nL, n2 * sizeof(double), h.stage + x.stage_slab * n2c,
                         n2c * sizeof(double), n2c * sizeof(double), (size_t)x.nSlabs, cudaMemcpyDeviceToHost, c->stream));
  for (const auto &x : h.downs) c->d2h_bytes += n2c * sizeof(double) * (size_t)x.nSlabs;
  h.downs.clear();
  return BGC_OK;
}

// Diagnostics accumulation: full-size, zero-initialised accumulator of a diagnostic array
// (SoA layout of the caller's whole block) and the add of one chunk into it.
static int acc_buffer(bgc_ctx *c, const char *key, size_t n, double **out) {
  const std::string k = std::string("acc.") + key;
  const bool fresh = c->arena.find(k) == c->arena.end() || c->arena[k].bytes < n * sizeof(double);
  RC(arena_d(c, k, n, out));
  if (fresh) {   // both pipeline slots add into this buffer: the zero fill must be complete before either does
    CU(cudaMemsetAsync(*out, 0, n * sizeof(double), c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return BGC_OK;
}
static int acc_add(bgc_ctx *c, const HostChunk &h, const char *key, const double *dev_chunk, int rows, int nSlabs,
                   const int *dev_kmax, int cols) {
  // rows = nL for (k,col[,n]) arrays, 1 for per-column arrays
  double *acc = nullptr;
  RC(acc_buffer(c, key, (size_t)rows * h.nC * nSlabs, &acc));
  LAUNCH(BGC_K_ACCUMULATE, 1, bgc::launch_accumulate(dev_chunk, acc, rows, h.cc, h.nC, h.c0, nSlabs, dev_kmax, cols, 1.0, c->stream));
  return BGC_OK;
}

// (col[,n]) arrays have the same layout in both spaces; elem = bytes per element
static int up_c(bgc_ctx *c, const HostChunk &h, const char *key, const void *host, size_t elem, int nSlabs, void **dev_out) {
  void *dev = nullptr;
  RC(arena_get(c, slot_key(h, key), (size_t)h.cc * elem * nSlabs, &dev));
  *dev_out = dev;
  if (!host) return fail(BGC_ERR_ARG, "null host array for %s", key);
  CU(cudaMemcpy2DAsync(dev, (size_t)h.cc * elem, (const char *)host + (size_t)h.c0 * elem, (size_t)h.nC * elem,
                       (size_t)h.cc * elem, (size_t)nSlabs, cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes += (size_t)h.cc * elem * (size_t)nSlabs;
  return BGC_OK;
}
static int down_c(bgc_ctx *c, const HostChunk &h, const void *dev, void *host, size_t elem, int nSlabs) {
  CU(cudaMemcpy2DAsync((char *)host + (size_t)h.c0 * elem, (size_t)h.nC * elem, dev, (size_t)h.cc * elem,
                       (size_t)h.cc * elem, (size_t)nSlabs, cudaMemcpyDeviceToHost, c->stream));
  c->d2h_bytes += (size_t)h.cc * elem * (size_t)nSlabs;
  return BGC_OK;
}

// Chunk plan: up to 32 k columns per chunk (large enough for full-speed DMA and kernels, small
// enough that several chunks overlap), a multiple of 32 columns so every device slab stays
// 256-byte aligned; BGC_HOST_CHUNK_COLUMNS overrides it (tests use tiny chunks).
static int chunk_columns(const bgc_ctx *c, int nC) {
  if (c->host_chunk_columns > 0) {
    const int cc = (c->host_chunk_columns + 31) / 32 * 32;
    return cc < nC ? cc : nC;
  }
  if (nC <= 8192) return nC;                 // small blocks: one chunk
  // at least four chunks, so that the download of one overlaps the upload and the kernels of the
  // next also when a GPU owns only a slab of the mesh; at most 32 k columns
  int cc = ((nC + 3) / 4 + 31) / 32 * 32;
  if (cc > 32768) cc = 32768;
  return cc < nC ? cc : nC;
}

// Run `body(chunk)` over all column chunks, alternating between the two pipeline slots.  The
// second slot's stream first waits for whatever was queued on the ctx stream before the call.
// The inventory fold accumulates into one device vector, so with the inventory on every chunk
// goes through slot 0 (sequential, still chunked).
template <class Body>
static int host_pipeline(bgc_ctx *c, int nL, int nC, Body body) {
  RC(join_pending(c));
  const int cc = chunk_columns(c, nC);
  const int nchunks = (nC + cc - 1) / cc;
  cudaStream_t user = c->stream;
  const bool two = nchunks > 1 && !c->inventory_on;
  if (two) {
    CU(cudaEventRecord(c->pipe_event, user));
    CU(cudaStreamWaitEvent(c->pipe_stream, c->pipe_event, 0));
  }
  int rc = BGC_OK;
  for (int i = 0; i < nchunks && rc == BGC_OK; ++i) {
    HostChunk h;
    h.nL = nL; h.nC = nC; h.c0 = i * cc; h.cc = (nC - h.c0 < cc) ? nC - h.c0 : cc;
    h.slot = two ? (i & 1) : 0;
    c->stream = h.slot ? c->pipe_stream : user;
    rc = body(h);
  }
  c->stream = user;
  if (rc == BGC_OK) run_zero_jobs(c); else c->zero_jobs.clear();
  // Fortran semantics: results are in host memory on return
  cudaError_t e1 = cudaStreamSynchronize(user), e2 = two ? cudaStreamSynchronize(c->pipe_stream) : cudaSuccess;
  if (rc == BGC_OK && e1 == cudaSuccess && e2 == cudaSuccess) {
    for (const auto &j : c->scatter_jobs)
      for (int col = 0; col < j.cc; ++col) {
        int km = (j.c0 + col) < j.nColumns ? j.kmax[j.c0 + col] : 0;
        if (km > j.nL) km = j.nL;
        if (km > 0) j.host[(size_t)(j.c0 + col) * j.nL + (km - 1)] = j.vec[col];
      }
  }
  c->scatter_jobs.clear();
  if (rc != BGC_OK) return rc;
  if (e1 != cudaSuccess) return fail(BGC_ERR_CUDA, "%s", cudaGetErrorString(e1));
  if (e2 != cudaSuccess) return fail(BGC_ERR_CUDA, "%s", cudaGetErrorString(e2));
  return BGC_OK;
}

// true when every column of the chunk is active over all levels: then every element of a
// (k,col) output is overwritten and the caller's previous contents need not be uploaded
static bool chunk_fully_active(const HostChunk &h, const int *kmax, int nCols) {
  if (h.c0 + h.cc > nCols) return false;
  for (int i = 0; i < h.cc; ++i) if (kmax[h.c0 + i] < h.nL) return false;
  return true;
}

static int check_dims(bgc_ctx *c, int nL, int nC, int nCols) {
  if (nL < 1 || nC < 1 || nCols < 0 || nCols > nC) return fail(BGC_ERR_ARG, "bad dimensions (%d,%d,%d)", nL, nC, nCols);
  // the kernels index elements with 32 bits (k_eco.cu); 30 * nL * nC must stay below 2^32
  if ((unsigned long long)nL * (unsigned long long)nC * BGC_TRACER_CNT >= (1ull << 32))
    return fail(BGC_ERR_ARG, "block of %d x %d cells is too large for one call (30*nL*nC >= 2^32): split the columns", nL, nC);
  (void)c;
  return BGC_OK;
}

// ------------------------------------------------------------------ inventory
// Stage 2 of the inventory reduction (stage 1 is fused into the source-sink kernels, which
// write one partial per block): add the partials into the ctx inventory vector.
static int inventory_fold(bgc_ctx *c, const double *partials, int nParts, int nGroups,
                          const int out_index[][bgc::kInvGroup]) {
  bgc::InventoryFoldArgs fa;
  memset(&fa, 0, sizeof fa);
  fa.nGroups = nGroups;
  for (int g = 0; g < bgc::kInvMaxGroups; ++g)
    for (int j = 0; j < bgc::kInvGroup; ++j) fa.out_index[g][j] = (g < nGroups) ? out_index[g][j] : -1;
  fa.partials = partials;
  fa.inventory = c->d_inventory;
  LAUNCH(BGC_K_INVENTORY, 1, bgc::launch_inventory_fold(fa, nParts, c->stream));
  return BGC_OK;
}

extern "C" int bgc_inventory_enable(bgc_ctx *c, int enable) {
  if (!c) return fail(BGC_ERR_ARG, "null ctx");
  c->inventory_on = enable != 0;
  return BGC_OK;
}
extern "C" int bgc_inventory_reset(bgc_ctx *c) {
  RC(use_device(c));
  CU(cudaMemsetAsync(c->d_inventory, 0, BGC_INVENTORY_LEN * sizeof(double), c->stream));
  return BGC_OK;
}
extern "C" int bgc_inventory_get(bgc_ctx *c, double out[BGC_INVENTORY_LEN]) {
  RC(use_device(c));
  if (!out) return fail(BGC_ERR_ARG, "null out");
  RC(join_pending(c));
  CU(cudaMemcpyAsync(out, c->d_inventory, BGC_INVENTORY_LEN * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}
extern "C" int bgc_inventory_device_ptr(bgc_ctx *c, double **dev_ptr) {
  if (!c || !dev_ptr) return fail(BGC_ERR_ARG, "null argument");
  *dev_ptr = c->d_inventory;
  return BGC_OK;
}

// ------------------------------------------------------------------ BGC_SourceSink
static int source_sink_device(bgc_ctx *c, const BgcInput *in, const BgcForcing *fo, BgcOutput *out,
                              const BgcDiagnostics *diag, int nL, int nC, int nCols, int alt_co2_use_eco,
                              bool host_call = false) {
  RC(ensure_bgc_tables(c));
  const size_t n2 = (size_t)nL * nC;
  BgcDiagnostics d;
  if (diag) d = *diag; else memset(&d, 0, sizeof d);
  bool any_diag = false;
  {
    double *const *pp = (double *const *)&d;
    for (size_t i = 0; i < sizeof(BgcDiagnostics) / sizeof(double *); ++i) any_diag = any_diag || pp[i] != nullptr;
  }
  if (!in->BGC_tracers || !in->PotentialTemperature || !in->Salinity || !in->cell_center_depth ||
      !in->cell_thickness || !in->cell_bottom_depth || !in->cell_latitude || !in->number_of_active_levels ||
      !fo->FESEDFLUX || !fo->dust_FLUX_IN || !fo->ShortWaveFlux_surface || !out->BGC_tendencies ||
      !out->PH_PREV_3D || !out->PH_PREV_ALT_CO2_3D)
    return fail(BGC_ERR_ARG, "bgc_source_sink: a required array pointer is NULL");
  const BgcParams &P = c->bgc_tab.p;
  if ((P.lrest_no3 || P.lrest_po4 || P.lrest_sio3) && !fo->NUTR_RESTORE_RTAU)
    return fail(BGC_ERR_ARG, "bgc_source_sink: lrest_* set but NUTR_RESTORE_RTAU is NULL");
  if ((P.lrest_no3 && !fo->NO3_CLIM) || (P.lrest_po4 && !fo->PO4_CLIM) || (P.lrest_sio3 && !fo->SiO3_CLIM))
    return fail(BGC_ERR_ARG, "bgc_source_sink: lrest_* set but the climatology array is NULL");

  // The column sweep does not depend on the carbonate kernel: it is launched FIRST on the ctx
  // stream, the carbonate kernel and the saturation-depth scan follow on a lower-priority side
  // stream forked here and joined before returning.  The sweep's last, partial wave leaves
  // most SMs idle (6.2 waves of 148 blocks on the EC60to30 mesh) and the FP64-bound carbonate
  // blocks fill them.
  RC(join_pending(c));   // a deferred join of the previous call ends here at the latest
  const int slot = (c->stream == c->pipe_stream) ? 1 : 0;
  cudaStream_t main_stream = c->stream;
  // Placement of the carbonate kernel (measured, profiles/concurrency_sweep_r02.txt).  Beside the
  // sweep (mode 1) it fills the SMs that the sweep's last, partial wave leaves idle: -0.35 ms on the
  // full EC60to30 mesh, -0.4 ms at half the mesh.  But a sweep block needs a WHOLE SM (every register,
  // 219 KB of shared memory) while a carbonate block fits into any SM that is still draining the
  // previous kernel, and once carbonate blocks sit there the SM never empties: when the sweep is less
  // than one wave (fewer blocks than SMs: a GPU's slab in an 8-way split) its blocks starve behind
  // them and the sweep takes 1.06 instead of 0.58 ms.  There the fork moves behind the sweep
  // (mode 2: the carbonate kernel runs beside the DMS / MACROS / surface kernels) - or, better, the
  // carbonate work is CONFINED to the SMs the sub-wave sweep cannot use (mode 3): as many 512-thread
  // carbonate blocks as there are idle SMs are launched before the sweep; each fills an SM's register
  // file, so they land on that many different SMs and the sweep's blocks find exactly the others free,
  // whichever kernel the block scheduler serves first.  They walk a share of the cells sized to last as
  // long as the sweep (its duration is one column's latency however few columns there are); the rest
  // of the cells follow behind the sweep as in mode 2.
  int co3_mode = c->concurrent_co3;
  const int sweep_blocks = bgc::eco_sweep_blocks(nC, c->eco_variant);
  const int idle_sms = c->sm_count - sweep_blocks;
  // (host-layout chunks are PCIe-bound and run two at a time on the pipeline slots: not confined)
  if (co3_mode == 1 && idle_sms > 0) co3_mode = (idle_sms >= 8 && c->co3_confined && !host_call) ? 3 : 2;
  else if (co3_mode == 3 && idle_sms <= 0) co3_mode = 1;
  size_t co3_split = 0;   // mode 3: cells [0, co3_split) run beside the sweep, the rest behind it
  if (co3_mode == 3) {
    // measured on B200 (profiles/concurrency_sweep_r02.txt): the sweep takes ~9.7 us per level, a
    // carbonate-filled SM retires ~80 cells per us; the share is cut to 85 % of that so that the
    // confined blocks finish with the sweep rather than after it (BGC_CO3_SHARE: percent, experiments)
    const double share = c->co3_share_pct > 0 ? c->co3_share_pct * 0.01 : 0.85;
    const double cells = share * (double)idle_sms * 80.0 * (9.7 * (double)nL);
    co3_split = (size_t)cells & ~(size_t)31;
    if (co3_split >= n2) co3_split = n2;
    if (co3_split == 0) co3_mode = 2;
  }
  cudaStream_t co3_stream = co3_mode ? c->side_stream[slot] : main_stream;
  if (co3_mode == 1 || co3_mode == 3) {   // fork BEFORE the sweep: the carbonate kernel runs beside it
    CU(cudaEventRecord(c->fork_event[slot], main_stream));
    CU(cudaStreamWaitEvent(co3_stream, c->fork_event[slot], 0));
  }

  // ecosystem + particle sweep, column-parallel
  bgc::EcoArgs ea;
  ea.nL = nL; ea.nC = nC; ea.nColumns = nCols; ea.alt_co2_use_eco = alt_co2_use_eco;
  ea.zero_shortcut = c->zero_shortcut;
  ea.any_restore = (P.lrest_no3 || P.lrest_po4 || P.lrest_sio3) ? 1 : 0;
  ea.tracers = in->BGC_tracers; ea.T = in->PotentialTemperature; ea.S = in->Salinity;
  ea.zmid = in->cell_center_depth; ea.dz = in->cell_thickness; ea.zbot = in->cell_bottom_depth;
  ea.lat = in->cell_latitude; ea.kmax = in->number_of_active_levels;
  ea.fesedflux = fo->FESEDFLUX; ea.rtau = fo->NUTR_RESTORE_RTAU; ea.no3_clim = fo->NO3_CLIM;
  ea.po4_clim = fo->PO4_CLIM; ea.sio3_clim = fo->SiO3_CLIM;
  ea.dust_flux_in = fo->dust_FLUX_IN; ea.sw_flux = fo->ShortWaveFlux_surface;
  ea.tend = out->BGC_tendencies;
  ea.d = d;
  // written by the carbonate kernel / the saturation-depth scan
  ea.d.diag_CO3 = ea.d.diag_HCO3 = ea.d.diag_H2CO3 = ea.d.diag_pH_3D = nullptr;
  ea.d.diag_CO3_ALT_CO2 = ea.d.diag_HCO3_ALT_CO2 = ea.d.diag_H2CO3_ALT_CO2 = ea.d.diag_pH_3D_ALT_CO2 = nullptr;
  ea.d.diag_co3_sat_calc = ea.d.diag_co3_sat_arag = nullptr;
  ea.d.diag_zsatcalc = ea.d.diag_zsatarag = nullptr;
  // declared in BGC_diagnostics_type but never zeroed nor written by the reference
  ea.d.diag_POC_ACCUM = ea.d.diag_DONr_remin = ea.d.diag_DOPr_remin = nullptr;
  ea.status = c->d_status;
  ea.block_trace = c->d_block_trace;
  if (!bgc::eco_rows_from_tables(c->bgc_tab, ea))
    return fail(BGC_ERR_PARAMS, "bgc_source_sink: the tracer index tables do not cover the %d tracer slots", BGC_TRACER_CNT);
  // diag_mode 2 = every array the sweep owns is present -> unchecked stores.  The arrays of
  // the other two kernels and the three never-touched members of the reference type may be
  // NULL without leaving that mode.
  int diag_mode = 0;
  if (any_diag) {
    diag_mode = 2;
    double *const *pp = (double *const *)&ea.d;
    const size_t first_carb = offsetof(BgcDiagnostics, diag_CO3) / sizeof(double *);
    const size_t last_carb = offsetof(BgcDiagnostics, diag_co3_sat_arag) / sizeof(double *);
    const size_t other[5] = {offsetof(BgcDiagnostics, diag_POC_ACCUM) / sizeof(double *),
                             offsetof(BgcDiagnostics, diag_DONr_remin) / sizeof(double *),
                             offsetof(BgcDiagnostics, diag_DOPr_remin) / sizeof(double *),
                             offsetof(BgcDiagnostics, diag_zsatcalc) / sizeof(double *),
                             offsetof(BgcDiagnostics, diag_zsatarag) / sizeof(double *)};
    for (size_t i = 0; i < sizeof(BgcDiagnostics) / sizeof(double *); ++i) {
      if (pp[i] || (i >= first_carb && i <= last_carb)) continue;
      if (i == other[0] || i == other[1] || i == other[2] || i == other[3] || i == other[4]) continue;
      diag_mode = 1;
      break;
    }
  }
  ea.inv_partials = nullptr;
  int inv_parts = 0;
  if (c->inventory_on) {
    inv_parts = bgc::eco_inventory_parts(ea, diag_mode, c->eco_variant);
    RC(arena_d(c, "inv_partials_bgc", (size_t)inv_parts * bgc::kEcoInvGroups * bgc::kInvGroup, &ea.inv_partials));
  }
  // carbonate chemistry, cell-parallel, then the saturation-depth scan (side stream)
  bgc::Co3Args ca;
  ca.nL = nL; ca.nC = nC; ca.nColumns = nCols;
  ca.tracers = in->BGC_tracers; ca.T = in->PotentialTemperature; ca.S = in->Salinity;
  ca.zmid = in->cell_center_depth; ca.kmax = in->number_of_active_levels;
  ca.ph_prev = out->PH_PREV_3D; ca.ph_prev_alt = out->PH_PREV_ALT_CO2_3D;
  ca.co3 = d.diag_CO3; ca.hco3 = d.diag_HCO3; ca.h2co3 = d.diag_H2CO3; ca.ph = d.diag_pH_3D;
  ca.co3_alt = d.diag_CO3_ALT_CO2; ca.hco3_alt = d.diag_HCO3_ALT_CO2; ca.h2co3_alt = d.diag_H2CO3_ALT_CO2;
  ca.ph_alt = d.diag_pH_3D_ALT_CO2; ca.sat_calc = d.diag_co3_sat_calc; ca.sat_arag = d.diag_co3_sat_arag;
  ca.status = c->d_status;
  ca.block_trace = c->d_block_trace;
  const bool want_zsat = d.diag_zsatcalc || d.diag_zsatarag;
  if (want_zsat) {   // the scan consumes these three: ctx scratch where the caller has no array
    const std::string sfx = slot ? "#1" : "#0";
    if (!ca.co3) RC(arena_d(c, "scratch_co3" + sfx, n2, &ca.co3));
    if (!ca.sat_calc) RC(arena_d(c, "scratch_satc" + sfx, n2, &ca.sat_calc));
    if (!ca.sat_arag) RC(arena_d(c, "scratch_sata" + sfx, n2, &ca.sat_arag));
  }
  auto launch_confined = [&]() -> int {
    c->stream = co3_stream;   // LAUNCH brackets its timing events on c->stream
    bgc::Co3Args part = ca;
    part.cell_begin = 0; part.cell_end = co3_split;
    const int pblocks = c->co3_pblocks > 0 ? c->co3_pblocks : idle_sms;
    int rc_part = [&]() -> int { LAUNCH(BGC_K_CO3_CELLS, 1, bgc::launch_co3_cells(part, pblocks, c->stream)); return BGC_OK; }();
    c->stream = main_stream;
    ca.cell_begin = co3_split; ca.cell_end = 0;
    return rc_part;
  };
  if (co3_mode == 3 && !c->co3_after) RC(launch_confined());
  LAUNCH(BGC_K_ECO_COLUMNS, 1, bgc::launch_eco_columns(ea, diag_mode, c->eco_variant, c->stream));
  if (co3_mode == 3 && c->co3_after) RC(launch_confined());
  if (co3_mode >= 2) {   // fork AFTER the sweep: the carbonate kernel runs beside whatever follows on the ctx stream
    CU(cudaEventRecord(c->fork_event[slot], main_stream));
    CU(cudaStreamWaitEvent(co3_stream, c->fork_event[slot], 0));
  }
  c->stream = co3_stream;   // LAUNCH brackets its timing events on c->stream
  int rc_side = [&]() -> int {
    if ((size_t)ca.cell_begin < n2) LAUNCH(BGC_K_CO3_CELLS, 1, bgc::launch_co3_cells(ca, 0, c->stream));
    if (want_zsat) {
      bgc::ZsatArgs za;
      za.nL = nL; za.nC = nC; za.nColumns = nCols; za.kmax = in->number_of_active_levels;
      za.co3 = ca.co3; za.sat_calc = ca.sat_calc; za.sat_arag = ca.sat_arag;
      za.zmid = in->cell_center_depth; za.zbot = in->cell_bottom_depth;
      za.zsatcalc = d.diag_zsatcalc; za.zsatarag = d.diag_zsatarag;
      LAUNCH(BGC_K_ZSAT_COLUMNS, 1, bgc::launch_zsat_columns(za, c->stream));
    }
    return BGC_OK;
  }();
  c->stream = main_stream;
  if (rc_side != BGC_OK) return rc_side;
  if (co3_mode) {
    CU(cudaEventRecord(c->join_event[slot], co3_stream));
    if (c->defer_join && !host_call && slot == 0) c->pending_join = true;   // joined at the next join point
    else CU(cudaStreamWaitEvent(main_stream, c->join_event[slot], 0));
  }
  if (c->inventory_on) {
    // destination of every value the sweep produced (layout: bgc_kernels.cuh, kEcoInvGroups)
    int oi[bgc::kEcoInvGroups][bgc::kInvGroup];
    for (int v = 0; v < bgc::kEcoInvGroups * bgc::kInvGroup; ++v) {
      int dst;
      if (v < BGC_TRACER_CNT) dst = v;                    // tracer slot
      else if (v < 32) dst = 60 + (v - 30);               // active cells, active columns
      else dst = any_diag ? 52 + (v - 32) : -1;           // the Jint_* sums exist only with diagnostics
      oi[v / bgc::kInvGroup][v % bgc::kInvGroup] = dst;
    }
    RC(inventory_fold(c, ea.inv_partials, inv_parts, bgc::kEcoInvGroups, oi));
  }
  return BGC_OK;
}

extern "C" int bgc_source_sink(bgc_ctx *c, const BgcInput *in, const BgcForcing *fo, BgcOutput *out,
                               BgcDiagnostics *diag, int nL, int nC, int nCols, int alt_co2_use_eco,
                               int mem_space) {
  RC(use_device(c));
  if (!in || !fo || !out) return fail(BGC_ERR_ARG, "bgc_source_sink: null argument block");
  RC(check_dims(c, nL, nC, nCols));
  if (mem_space == BGC_MEM_DEVICE_SOA)
    return source_sink_device(c, in, fo, out, diag, nL, nC, nCols, alt_co2_use_eco);
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  if (!c->have_bgc) return fail(BGC_ERR_PARAMS, "bgc_set_params has not been called on this ctx");

  RC(ensure_bgc_tables(c));
  const BgcParams &P = c->bgc_tab.p;
  // DIC_ALT_CO2 is clamped and then never read by BGC_SourceSink (BGC_mod.F90:748): not uploaded
  const unsigned dead_slabs = 1u << (c->bgc_tab.ind.dic_alt_co2_ind - 1);
  return host_pipeline(c, nL, nC, [&](HostChunk &h) -> int {
    const size_t n2 = (size_t)h.nL * h.cc;
    int cols = nCols - h.c0;
    if (cols < 0) cols = 0;
    if (cols > h.cc) cols = h.cc;
    BgcInput din; BgcForcing dfo; BgcOutput dout; BgcDiagnostics dd;
    memset(&din, 0, sizeof din); memset(&dfo, 0, sizeof dfo); memset(&dout, 0, sizeof dout); memset(&dd, 0, sizeof dd);
    double *t = nullptr; void *v = nullptr;
    // staging: 30 + 5 + 1 + 4 + 2 slabs up at most; 30 + 2 + 61 + 18*4 down at most
    RC(stage_reserve(c, h, BGC_TRACER_CNT + 2 + 61 + 18 * BGC_AUTOTROPH_CNT));
    RC(up_k(c, h, "bgc.tracers", in->BGC_tracers, BGC_TRACER_CNT, &t, dead_slabs)); din.BGC_tracers = t;
    RC(up_k(c, h, "bgc.T", in->PotentialTemperature, 1, &t)); din.PotentialTemperature = t;
    RC(up_k(c, h, "bgc.S", in->Salinity, 1, &t)); din.Salinity = t;
    RC(up_k(c, h, "bgc.zmid", in->cell_center_depth, 1, &t)); din.cell_center_depth = t;
    RC(up_k(c, h, "bgc.dz", in->cell_thickness, 1, &t)); din.cell_thickness = t;
    RC(up_k(c, h, "bgc.zbot", in->cell_bottom_depth, 1, &t)); din.cell_bottom_depth = t;
    RC(up_c(c, h, "bgc.lat", in->cell_latitude, sizeof(double), 1, &v)); din.cell_latitude = (double *)v;
    RC(up_c(c, h, "bgc.kmax", in->number_of_active_levels, sizeof(int), 1, &v)); din.number_of_active_levels = (int *)v;
    RC(up_k(c, h, "bgc.fesed", fo->FESEDFLUX, 1, &t)); dfo.FESEDFLUX = t;
    if (P.lrest_no3 || P.lrest_po4 || P.lrest_sio3) { RC(up_k(c, h, "bgc.rtau", fo->NUTR_RESTORE_RTAU, 1, &t)); dfo.NUTR_RESTORE_RTAU = t; }
    if (P.lrest_no3) { RC(up_k(c, h, "bgc.no3clim", fo->NO3_CLIM, 1, &t)); dfo.NO3_CLIM = t; }
    if (P.lrest_po4) { RC(up_k(c, h, "bgc.po4clim", fo->PO4_CLIM, 1, &t)); dfo.PO4_CLIM = t; }
    if (P.lrest_sio3) { RC(up_k(c, h, "bgc.sio3clim", fo->SiO3_CLIM, 1, &t)); dfo.SiO3_CLIM = t; }
    RC(up_c(c, h, "bgc.dust", fo->dust_FLUX_IN, sizeof(double), 1, &v)); dfo.dust_FLUX_IN = (double *)v;
    RC(up_c(c, h, "bgc.sw", fo->ShortWaveFlux_surface, sizeof(double), 1, &v)); dfo.ShortWaveFlux_surface = (double *)v;
    RC(up_k(c, h, "bgc.phprev", out->PH_PREV_3D, 1, &t)); dout.PH_PREV_3D = t;
    RC(up_k(c, h, "bgc.phprevalt", out->PH_PREV_ALT_CO2_3D, 1, &t)); dout.PH_PREV_ALT_CO2_3D = t;
    RC(arena_d(c, slot_key(h, "bgc.tend"), n2 * BGC_TRACER_CNT, &dout.BGC_tendencies));

    if (diag) {
#define DEV_K2(name) if (diag->name) RC(arena_d(c, slot_key(h, "bgc.d." #name), n2, &dd.name));
#define DEV_KA(name) if (diag->name) RC(arena_d(c, slot_key(h, "bgc.d." #name), n2 * BGC_AUTOTROPH_CNT, &dd.name));
#define DEV_CA(name) if (diag->name) RC(arena_d(c, slot_key(h, "bgc.d." #name), (size_t)h.cc * BGC_AUTOTROPH_CNT, &dd.name));
#define DEV_C1(name) if (diag->name) RC(arena_d(c, slot_key(h, "bgc.d." #name), (size_t)h.cc, &dd.name));
      BGC_DIAG_K2_LIST(DEV_K2) BGC_DIAG_KA_LIST(DEV_KA) BGC_DIAG_CA_LIST(DEV_CA) BGC_DIAG_C1_LIST(DEV_C1)
#undef DEV_K2
#undef DEV_KA
#undef DEV_CA
#undef DEV_C1
    }

    RC(flush_up(c, h));
    RC(source_sink_device(c, &din, &dfo, &dout, diag ? &dd : nullptr, h.nL, h.cc, cols, alt_co2_use_eco, true));

    // structurally zero: the DIC_ALT_CO2 tendency when alt_co2_use_eco is off (BGC_mod.F90:1741-1745)
    const unsigned long long zero_tend = alt_co2_use_eco ? 0ull : (1ull << (c->bgc_tab.ind.dic_alt_co2_ind - 1));
    RC(down_k(c, h, dout.BGC_tendencies, out->BGC_tendencies, BGC_TRACER_CNT, zero_tend));
    RC(down_k(c, h, dout.PH_PREV_3D, out->PH_PREV_3D, 1));
    RC(down_k(c, h, dout.PH_PREV_ALT_CO2_3D, out->PH_PREV_ALT_CO2_3D, 1));
    if (diag && c->diag_accumulate) {
      dd.diag_POC_ACCUM = dd.diag_DONr_remin = dd.diag_DOPr_remin = nullptr;
#define AC_K2(name) if (dd.name) RC(acc_add(c, h, "bgc." #name, dd.name, h.nL, 1, nullptr, cols));
#define AC_KA(name) if (dd.name) RC(acc_add(c, h, "bgc." #name, dd.name, h.nL, BGC_AUTOTROPH_CNT, nullptr, cols));
#define AC_CA(name) if (dd.name) RC(acc_add(c, h, "bgc." #name, dd.name, 1, BGC_AUTOTROPH_CNT, nullptr, cols));
#define AC_C1(name) if (dd.name) RC(acc_add(c, h, "bgc." #name, dd.name, 1, 1, nullptr, cols));
      BGC_DIAG_K2_LIST(AC_K2) BGC_DIAG_KA_LIST(AC_KA) BGC_DIAG_CA_LIST(AC_CA) BGC_DIAG_C1_LIST(AC_C1)
#undef AC_K2
#undef AC_KA
#undef AC_CA
#undef AC_C1
    } else if (diag) {
      // the three never-touched arrays stay exactly as the caller left them
      dd.diag_POC_ACCUM = dd.diag_DONr_remin = dd.diag_DOPr_remin = nullptr;
      // structurally zero: the three restoring terms when restoring is switched off (BGC_mod.F90:1545-1552
      // and the like assign RESTORE = c0 then)
      const double *const off_no3 = P.lrest_no3 ? nullptr : dd.diag_NO3_RESTORE, *const off_po4 = P.lrest_po4 ? nullptr : dd.diag_PO4_RESTORE,
                   *const off_sio3 = P.lrest_sio3 ? nullptr : dd.diag_SiO3_RESTORE;
      // structurally zero for a functional group, whatever the inputs are: N fixation of a group that
      // is no N fixer (:1331-1338), bSi formation / the SiO3 limitation term of a group without Si quota or
      // kSiO3 (:1140-1150, :1230), CaCO3 formation of a group that is no implicit calcifier (:1255-1278)
      unsigned long long z_nfix = 0, z_bsi = 0, z_caco3 = 0, z_sio3lim = 0;
      for (int g = 0; g < BGC_AUTOTROPH_CNT; ++g) {
        const BgcAutotroph &at = c->bgc_tab.a[g];
        if (!at.Nfixer) z_nfix |= 1ull << g;
        if (at.Si_ind <= 0) z_bsi |= 1ull << g;
        if (!at.imp_calcifier) z_caco3 |= 1ull << g;
        if (!(at.kSiO3 > 0.0)) z_sio3lim |= 1ull << g;
      }
      auto ka_zero = [&](const double *p) -> unsigned long long {
        return p == dd.diag_Nfix ? z_nfix : p == dd.diag_bSi_form ? z_bsi : p == dd.diag_CaCO3_form ? z_caco3 :
               p == dd.diag_SiO3_lim ? z_sio3lim : 0ull;
      };
      // The nine sediment diagnostics are set in a column's BOTTOM cell alone (:2522-2631; zero fill elsewhere,
      // :625-727): the bottom values come down as one vector per array, the host range is zero-filled by the
      // host threads and the values are put in place when the call's transfers are complete.
      double *const sed[9] = {dd.diag_calcToSed, dd.diag_pocToSed, dd.diag_ponToSed, dd.diag_popToSed, dd.diag_bsiToSed,
                              dd.diag_dustToSed, dd.diag_pfeToSed, dd.diag_SedDenitrif, dd.diag_OtherRemin};
      double *const sed_host[9] = {diag->diag_calcToSed, diag->diag_pocToSed, diag->diag_ponToSed, diag->diag_popToSed,
                                   diag->diag_bsiToSed, diag->diag_dustToSed, diag->diag_pfeToSed, diag->diag_SedDenitrif,
                                   diag->diag_OtherRemin};
      bgc::BottomGatherArgs ga;
      memset(&ga, 0, sizeof ga);
      int sed_of[9];
      for (int j = 0; j < 9; ++j) if (sed[j]) { ga.src[ga.n] = sed[j]; sed_of[ga.n++] = j; }
      if (ga.n) {
        void *hv = nullptr;
        RC(host_stage_get(c, "bgc.bottom", (size_t)9 * h.nC * sizeof(double), &hv));
        RC(arena_d(c, slot_key(h, "bgc.bottom"), (size_t)ga.n * h.cc, &ga.out));
        ga.kmax = din.number_of_active_levels; ga.nL = h.nL; ga.cc = h.cc; ga.nColumns = cols;
        LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_bottom_gather(ga, c->stream));
        for (int q = 0; q < ga.n; ++q) {
          double *vec = (double *)hv + (size_t)sed_of[q] * h.nC + h.c0;
          CU(cudaMemcpyAsync(vec, ga.out + (size_t)q * h.cc, (size_t)h.cc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
          c->d2h_bytes += (size_t)h.cc * sizeof(double);
          c->scatter_jobs.push_back({sed_host[sed_of[q]], vec, in->number_of_active_levels, h.nL, h.c0, h.cc, nCols});
        }
      }
      auto is_sed = [&](const double *p) { for (int j = 0; j < 9; ++j) if (p == sed[j]) return true; return false; };
#define DN_K2(name) if (dd.name) RC(down_k(c, h, dd.name, diag->name, 1, (dd.name == off_no3 || dd.name == off_po4 || dd.name == off_sio3 || is_sed(dd.name)) ? 1ull : 0ull));
#define DN_KA(name) if (dd.name) RC(down_k(c, h, dd.name, diag->name, BGC_AUTOTROPH_CNT, ka_zero(dd.name)));
#define DN_CA(name) if (dd.name) RC(down_c(c, h, dd.name, diag->name, sizeof(double), BGC_AUTOTROPH_CNT));
#define DN_C1(name) if (dd.name) RC(down_c(c, h, dd.name, diag->name, sizeof(double), 1));
      BGC_DIAG_K2_LIST(DN_K2) BGC_DIAG_KA_LIST(DN_KA) BGC_DIAG_CA_LIST(DN_CA) BGC_DIAG_C1_LIST(DN_C1)
#undef DN_K2
#undef DN_KA
#undef DN_CA
#undef DN_C1
    }
    return flush_down(c, h);
  });
}


// Level 1 of a host array A(k,col,n) (level fastest): nC*nSlabs doubles at a stride of nL.  The
// copy engine handles an 8-byte-wide 2-D copy one row at a time (millions of rows on a full
// mesh), so the surface-flux calls gather level 1 on the host instead - a few threads, each
// striding through its share into a page-locked staging buffer of the ctx - and upload that
// with ONE contiguous copy.
static int gather_level1(bgc_ctx *c, const char *key, const double *host, int nL, size_t count, double *dev) {
  HostStage &hs = c->host_stage[key];
  if (hs.bytes < count * sizeof(double)) {
    if (hs.p) cudaFreeHost(hs.p);
    hs.p = nullptr; hs.bytes = 0;
    CU(cudaHostAlloc(&hs.p, count * sizeof(double), cudaHostAllocDefault));
    hs.bytes = count * sizeof(double);
  }
  double *st = (double *)hs.p;
  // the previous call's upload from this buffer has completed: every host-layout call synchronises before returning
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads > 8) nthreads = 8;
  if (count < (size_t)1 << 16 || nthreads < 2) nthreads = 1;
  auto work = [=](size_t a, size_t b) { for (size_t i = a; i < b; ++i) st[i] = host[i * (size_t)nL]; };
  if (nthreads == 1) {
    work(0, count);
  } else {
    std::vector<std::thread> pool;
    const size_t per = (count + nthreads - 1) / nthreads;
    for (unsigned t = 0; t < nthreads; ++t) {
      const size_t a = (size_t)t * per, b = a + per < count ? a + per : count;
      if (a < b) pool.emplace_back(work, a, b);
    }
    for (auto &th : pool) th.join();
  }
  CU(cudaMemcpyAsync(dev, st, count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes += count * sizeof(double);
  return BGC_OK;
}

// ------------------------------------------------------------------ BGC_SurfaceFluxes
static int surface_fluxes_device(bgc_ctx *c, const BgcInput *in, BgcForcing *fo, BgcFluxDiagnostics *diag,
                                 int nL, int nC, int nCols) {
  RC(ensure_bgc_tables(c));
  if (!in->BGC_tracers || !fo->depositionFlux || !fo->riverFlux || !fo->gasFlux || !fo->seaIceFlux ||
      !fo->netFlux || !fo->iceFraction || !fo->windSpeedSquared10m || !fo->SST || !fo->SSS || !fo->surfacePressure)
    return fail(BGC_ERR_ARG, "bgc_surface_fluxes: a required array pointer is NULL");
  if (fo->lcalc_CO2_gas_flux && (!fo->surface_pH || !fo->surface_pH_alt_co2 || !fo->atmCO2 || !fo->atmCO2_ALT_CO2))
    return fail(BGC_ERR_ARG, "bgc_surface_fluxes: CO2 flux requested but pH / atmCO2 arrays are NULL");
  bgc::SurfArgs sa;
  sa.nL = nL; sa.nC = nC; sa.nColumns = nCols;
  sa.tracers = in->BGC_tracers;
  sa.f = *fo;
  if (diag) sa.d = *diag; else memset(&sa.d, 0, sizeof sa.d);
  sa.status = c->d_status;
  LAUNCH(BGC_K_SURFACE_FLUXES, 1, bgc::launch_surface_fluxes(sa, c->stream));
  return BGC_OK;
}

extern "C" int bgc_surface_fluxes(bgc_ctx *c, const BgcInput *in, BgcForcing *fo, BgcFluxDiagnostics *diag,
                                  int nL, int nC, int nCols, int mem_space) {
  RC(use_device(c));
  if (!in || !fo) return fail(BGC_ERR_ARG, "bgc_surface_fluxes: null argument block");
  RC(check_dims(c, nL, nC, nCols));
  if (mem_space == BGC_MEM_DEVICE_SOA) return surface_fluxes_device(c, in, fo, diag, nL, nC, nCols);
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  if (!c->have_bgc) return fail(BGC_ERR_PARAMS, "bgc_set_params has not been called on this ctx");

  // Only level 1 of the tracer array is read: upload that single level per tracer.  Per-column
  // data only (a few hundred MB at most): one chunk.
  HostChunk h; h.nL = nL; h.nC = nC; h.c0 = 0; h.cc = nC; h.slot = 0;
  BgcInput din; BgcForcing dfo; BgcFluxDiagnostics dd;
  memset(&din, 0, sizeof din); memset(&dd, 0, sizeof dd);
  dfo = *fo;
  double *surf = nullptr;
  RC(arena_d(c, "surf.tracers", (size_t)nC * BGC_TRACER_CNT, &surf));
  if (!in->BGC_tracers) return fail(BGC_ERR_ARG, "bgc_surface_fluxes: BGC_tracers is NULL");
  din.BGC_tracers = surf;
  void *v = nullptr;
  // (the forcing uploads are queued first: they fly while the host threads gather level 1 of the tracers)
#define UPC(member, n) do { if (fo->member) { RC(up_c(c, h, "surf." #member, fo->member, sizeof(double), (n), &v)); dfo.member = (double *)v; } } while (0)
  UPC(surfacePressure, 1); UPC(iceFraction, 1); UPC(windSpeedSquared10m, 1); UPC(atmCO2, 1); UPC(atmCO2_ALT_CO2, 1);
  UPC(surface_pH, 1); UPC(surface_pH_alt_co2, 1); UPC(surfaceDepth, 1); UPC(SST, 1); UPC(SSS, 1);
  UPC(depositionFlux, BGC_TRACER_CNT); UPC(riverFlux, BGC_TRACER_CNT); UPC(gasFlux, BGC_TRACER_CNT);
  UPC(seaIceFlux, BGC_TRACER_CNT);
  // netFlux is written for every column below numColumns (BGC_mod.F90:2929-2942): the caller's values are needed
  // only where a block has columns beyond numColumns
  if (nCols < nC) { UPC(netFlux, BGC_TRACER_CNT); }
  else if (fo->netFlux) { double *nf = nullptr; RC(arena_d(c, "surf.netFlux#0", (size_t)nC * BGC_TRACER_CNT, &nf)); dfo.netFlux = nf; }
#undef UPC
  RC(gather_level1(c, "surf.tracers", in->BGC_tracers, nL, (size_t)nC * BGC_TRACER_CNT, surf));
  if (diag) {
#define DEV_F(name) if (diag->name) RC(arena_d(c, "surf.d." #name, (size_t)nC, &dd.name));
    BGC_FLUX_DIAG_LIST(DEV_F)
#undef DEV_F
  }
  // device tracer "array" holds level 1 only: nL = 1 on the device side
  RC(surface_fluxes_device(c, &din, &dfo, diag ? &dd : nullptr, 1, nC, nCols));
#define DNC(member, n) do { if (fo->member) RC(down_c(c, h, dfo.member, fo->member, sizeof(double), (n))); } while (0)
  DNC(iceFraction, 1); DNC(surface_pH, 1); DNC(surface_pH_alt_co2, 1);
  DNC(netFlux, BGC_TRACER_CNT);
#undef DNC
  {   // of the four input fluxes the routine touches the iron slot (bioavailable fraction, :2828-2838) and the gas
      // fluxes it computes (O2, DIC, DIC_ALT_CO2, :2860-2925): only those (col) vectors come back
    const BgcIndices &I = c->bgc_tab.ind;
    auto slot_down = [&](double *dev, double *hostp, int ind) -> int {
      if (!dev || !hostp) return BGC_OK;
      return down_c(c, h, dev + (size_t)(ind - 1) * nC, hostp + (size_t)(ind - 1) * nC, sizeof(double), 1);
    };
    RC(slot_down(dfo.depositionFlux, fo->depositionFlux, I.fe_ind));
    RC(slot_down(dfo.riverFlux, fo->riverFlux, I.fe_ind));
    RC(slot_down(dfo.seaIceFlux, fo->seaIceFlux, I.fe_ind));
    const int gas[4] = {I.fe_ind, I.o2_ind, I.dic_ind, I.dic_alt_co2_ind};
    for (int q = 0; q < 4; ++q) RC(slot_down(dfo.gasFlux, fo->gasFlux, gas[q]));
  }
  if (diag) {
#define DN_F(name) if (dd.name) RC(down_c(c, h, dd.name, diag->name, sizeof(double), 1));
    BGC_FLUX_DIAG_LIST(DN_F)
#undef DN_F
  }
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

// ------------------------------------------------------------------ co2calc_1point, batched
extern "C" int bgc_co2calc_points(bgc_ctx *c, int n, const double *depth, const double *temp, const double *salt,
                                  const double *dic, const double *ta, const double *pt, const double *sit,
                                  const double *phlo, const double *phhi, const double *xco2,
                                  const double *atmpres, double *ph, double *co2star, double *dco2star,
                                  double *pco2surf, double *dpco2, int mem_space) {
  RC(use_device(c));
  if (n < 0) return fail(BGC_ERR_ARG, "negative n");
  if (n == 0) return BGC_OK;
  if (!temp || !salt || !dic || !ta || !pt || !sit || !phlo || !phhi || !xco2 || !atmpres || !ph || !co2star ||
      !dco2star || !pco2surf || !dpco2)
    return fail(BGC_ERR_ARG, "bgc_co2calc_points: null array");
  bgc::Co2PointsArgs a;
  a.n = n; a.status = c->d_status;
  if (mem_space == BGC_MEM_DEVICE_SOA) {
    a.depth = depth; a.temp = temp; a.salt = salt; a.dic = dic; a.ta = ta; a.pt = pt; a.sit = sit;
    a.phlo = phlo; a.phhi = phhi; a.xco2 = xco2; a.atmpres = atmpres;
    a.ph = ph; a.co2star = co2star; a.dco2star = dco2star; a.pco2surf = pco2surf; a.dpco2 = dpco2;
    LAUNCH(BGC_K_CO2CALC_POINTS, 1, bgc::launch_co2calc_points(a, c->stream));
    return BGC_OK;
  }
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  const size_t b = (size_t)n * sizeof(double);
  double *in_dev = nullptr, *out_dev = nullptr;
  RC(arena_d(c, "pts.in", (size_t)n * 10, &in_dev));
  RC(arena_d(c, "pts.out", (size_t)n * 5, &out_dev));
  const double *src[10] = {temp, salt, dic, ta, pt, sit, phlo, phhi, xco2, atmpres};
  for (int i = 0; i < 10; ++i) CU(cudaMemcpyAsync(in_dev + (size_t)i * n, src[i], b, cudaMemcpyHostToDevice, c->stream));
  a.depth = nullptr;   // level 1: depth does not enter the result (co2calc.F90:149-160)
  a.temp = in_dev; a.salt = in_dev + (size_t)n; a.dic = in_dev + 2 * (size_t)n; a.ta = in_dev + 3 * (size_t)n;
  a.pt = in_dev + 4 * (size_t)n; a.sit = in_dev + 5 * (size_t)n; a.phlo = in_dev + 6 * (size_t)n;
  a.phhi = in_dev + 7 * (size_t)n; a.xco2 = in_dev + 8 * (size_t)n; a.atmpres = in_dev + 9 * (size_t)n;
  a.ph = out_dev; a.co2star = out_dev + (size_t)n; a.dco2star = out_dev + 2 * (size_t)n;
  a.pco2surf = out_dev + 3 * (size_t)n; a.dpco2 = out_dev + 4 * (size_t)n;
  LAUNCH(BGC_K_CO2CALC_POINTS, 1, bgc::launch_co2calc_points(a, c->stream));
  double *dst[5] = {ph, co2star, dco2star, pco2surf, dpco2};
  for (int i = 0; i < 5; ++i) CU(cudaMemcpyAsync(dst[i], out_dev + (size_t)i * n, b, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

// ------------------------------------------------------------------ comp_CO3terms / comp_co3_sat_vals, batched
// Host-layout helper of the point entries: `nin` input arrays and `nout` output arrays of n doubles
// (plus an optional int array) staged through the arena.
static int points_stage(bgc_ctx *c, const char *key, int n, const double *const *in, int nin, double **din,
                        const int *k_level, const int **dk, int nout, double **dout) {
  double *buf = nullptr;
  RC(arena_d(c, std::string(key) + ".buf", (size_t)n * (nin + nout), &buf));
  for (int i = 0; i < nin; ++i) {
    din[i] = buf + (size_t)i * n;
    CU(cudaMemcpyAsync(din[i], in[i], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  for (int i = 0; i < nout; ++i) dout[i] = buf + (size_t)(nin + i) * n;
  *dk = nullptr;
  if (k_level) {
    void *kb = nullptr;
    RC(arena_get(c, std::string(key) + ".k", (size_t)n * sizeof(int), &kb));
    CU(cudaMemcpyAsync(kb, k_level, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    *dk = (const int *)kb;
  }
  return BGC_OK;
}

extern "C" int bgc_comp_co3terms(bgc_ctx *c, int n, const int *k_level, int k_all, const double *depth,
                                 const double *temp, const double *salt, const double *dic, const double *ta,
                                 const double *pt, const double *sit, const double *phlo, const double *phhi,
                                 double *ph, double *h2co3, double *hco3, double *co3, int mem_space) {
  RC(use_device(c));
  if (n < 0) return fail(BGC_ERR_ARG, "negative n");
  if (n == 0) return BGC_OK;
  if (!depth || !temp || !salt || !dic || !ta || !pt || !sit || !phlo || !phhi || !ph || !h2co3 || !hco3 || !co3)
    return fail(BGC_ERR_ARG, "bgc_comp_co3terms: null array");
  if (!k_level && k_all < 1) return fail(BGC_ERR_ARG, "bgc_comp_co3terms: k_level is NULL and k_all < 1");
  RC(join_pending(c));
  bgc::Co3TermsPointsArgs a;
  a.n = n; a.k_all = k_all; a.status = c->d_status;
  if (mem_space == BGC_MEM_DEVICE_SOA) {
    a.k = k_level; a.depth = depth; a.temp = temp; a.salt = salt; a.dic = dic; a.ta = ta; a.pt = pt; a.sit = sit;
    a.phlo = phlo; a.phhi = phhi; a.ph = ph; a.h2co3 = h2co3; a.hco3 = hco3; a.co3 = co3;
    LAUNCH(BGC_K_CO2CALC_POINTS, 1, bgc::launch_co3terms_points(a, c->stream));
    return BGC_OK;
  }
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  const double *in[9] = {depth, temp, salt, dic, ta, pt, sit, phlo, phhi};
  double *din[9], *dout[4];
  RC(points_stage(c, "co3terms", n, in, 9, din, k_level, &a.k, 4, dout));
  a.depth = din[0]; a.temp = din[1]; a.salt = din[2]; a.dic = din[3]; a.ta = din[4]; a.pt = din[5]; a.sit = din[6];
  a.phlo = din[7]; a.phhi = din[8];
  a.ph = dout[0]; a.h2co3 = dout[1]; a.hco3 = dout[2]; a.co3 = dout[3];
  LAUNCH(BGC_K_CO2CALC_POINTS, 1, bgc::launch_co3terms_points(a, c->stream));
  double *dst[4] = {ph, h2co3, hco3, co3};
  for (int i = 0; i < 4; ++i) CU(cudaMemcpyAsync(dst[i], dout[i], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

extern "C" int bgc_comp_co3_sat_vals(bgc_ctx *c, int n, const int *k_level, int k_all, const double *depth,
                                     const double *temp, const double *salt, double *co3_sat_calc,
                                     double *co3_sat_arag, int mem_space) {
  RC(use_device(c));
  if (n < 0) return fail(BGC_ERR_ARG, "negative n");
  if (n == 0) return BGC_OK;
  if (!depth || !temp || !salt || !co3_sat_calc || !co3_sat_arag) return fail(BGC_ERR_ARG, "bgc_comp_co3_sat_vals: null array");
  if (!k_level && k_all < 1) return fail(BGC_ERR_ARG, "bgc_comp_co3_sat_vals: k_level is NULL and k_all < 1");
  RC(join_pending(c));
  bgc::Co3SatPointsArgs a;
  a.n = n; a.k_all = k_all;
  if (mem_space == BGC_MEM_DEVICE_SOA) {
    a.k = k_level; a.depth = depth; a.temp = temp; a.salt = salt; a.sat_calc = co3_sat_calc; a.sat_arag = co3_sat_arag;
    LAUNCH(BGC_K_CO2CALC_POINTS, 1, bgc::launch_co3_sat_points(a, c->stream));
    return BGC_OK;
  }
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  const double *in[3] = {depth, temp, salt};
  double *din[3], *dout[2];
  RC(points_stage(c, "co3sat", n, in, 3, din, k_level, &a.k, 2, dout));
  a.depth = din[0]; a.temp = din[1]; a.salt = din[2]; a.sat_calc = dout[0]; a.sat_arag = dout[1];
  LAUNCH(BGC_K_CO2CALC_POINTS, 1, bgc::launch_co3_sat_points(a, c->stream));
  CU(cudaMemcpyAsync(co3_sat_calc, dout[0], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(co3_sat_arag, dout[1], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

// ------------------------------------------------------------------ DMS
static int dms_source_sink_device(bgc_ctx *c, const DmsInput *in, const DmsForcing *fo, DmsOutput *out,
                                  const DmsDiagnostics *diag, int nL, int nC, int nCols) {
  RC(ensure_dms_tables(c));
  if (!in->DMS_tracers || !in->cell_thickness || !in->number_of_active_levels || !fo->SST ||
      !fo->ShortWaveFlux_surface || !out->DMS_tendencies)
    return fail(BGC_ERR_ARG, "dms_source_sink: a required array pointer is NULL");
  bgc::DmsArgs a;
  a.nL = nL; a.nC = nC; a.nColumns = nCols;
  a.tracers = in->DMS_tracers; a.dz = in->cell_thickness; a.kmax = in->number_of_active_levels;
  a.sst = fo->SST; a.sw_flux = fo->ShortWaveFlux_surface; a.tend = out->DMS_tendencies;
  if (diag) a.d = *diag; else memset(&a.d, 0, sizeof a.d);
  const bool inv = c->inventory_on;
  a.inv_partials = nullptr;
  if (inv) RC(arena_d(c, "inv_partials_dms", (size_t)bgc::dms_inventory_parts(nL, nC, c->dms_variant) * bgc::kInvGroup, &a.inv_partials));
  LAUNCH(BGC_K_DMS_COLUMNS, 1, bgc::launch_dms_columns(a, c->dms_variant, c->stream));
  if (inv) {   // only DMS and DMSP have non-zero tendencies (DMS_mod.F90:413, :741-742)
    int oi[1][bgc::kInvGroup];
    for (int j = 0; j < bgc::kInvGroup; ++j) oi[0][j] = -1;
    oi[0][0] = 30 + c->dms_tab.ind.dms_ind - 1;
    oi[0][1] = 30 + c->dms_tab.ind.dmsp_ind - 1;
    RC(inventory_fold(c, a.inv_partials, bgc::dms_inventory_parts(nL, nC, c->dms_variant), 1, oi));
  }
  return BGC_OK;
}

extern "C" int dms_source_sink(bgc_ctx *c, const DmsInput *in, const DmsForcing *fo, DmsOutput *out,
                               DmsDiagnostics *diag, int nL, int nC, int nCols, int mem_space) {
  RC(use_device(c));
  if (!in || !fo || !out) return fail(BGC_ERR_ARG, "dms_source_sink: null argument block");
  RC(check_dims(c, nL, nC, nCols));
  if (mem_space == BGC_MEM_DEVICE_SOA) return dms_source_sink_device(c, in, fo, out, diag, nL, nC, nCols);
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  if (!in->number_of_active_levels) return fail(BGC_ERR_ARG, "dms_source_sink: number_of_active_levels is NULL");
  RC(ensure_dms_tables(c));
  // NO3 and DOC are copied by the reference and never reach an output (DMS_mod.F90:471-472): not uploaded
  const unsigned dead_slabs = (1u << (c->dms_tab.ind.no3_ind - 1)) | (1u << (c->dms_tab.ind.doc_ind - 1));
  return host_pipeline(c, nL, nC, [&](HostChunk &h) -> int {
    const size_t n2 = (size_t)h.nL * h.cc;
    int cols = nCols - h.c0;
    if (cols < 0) cols = 0;
    if (cols > h.cc) cols = h.cc;
    DmsInput din; DmsForcing dfo; DmsOutput dout; DmsDiagnostics dd;
    memset(&din, 0, sizeof din); memset(&dfo, 0, sizeof dfo); memset(&dout, 0, sizeof dout); memset(&dd, 0, sizeof dd);
    double *t = nullptr; void *v = nullptr;
    RC(stage_reserve(c, h, DMS_TRACER_CNT + 1 + sizeof(DmsDiagnostics) / sizeof(double *)));
    RC(up_k(c, h, "dms.tracers", in->DMS_tracers, DMS_TRACER_CNT, &t, dead_slabs)); din.DMS_tracers = t;
    RC(up_k(c, h, "dms.dz", in->cell_thickness, 1, &t)); din.cell_thickness = t;
    RC(up_c(c, h, "dms.kmax", in->number_of_active_levels, sizeof(int), 1, &v)); din.number_of_active_levels = (int *)v;
    RC(up_c(c, h, "dms.sst", fo->SST, sizeof(double), 1, &v)); dfo.SST = (double *)v;
    RC(up_c(c, h, "dms.sw", fo->ShortWaveFlux_surface, sizeof(double), 1, &v)); dfo.ShortWaveFlux_surface = (double *)v;
    RC(arena_d(c, slot_key(h, "dms.tend"), n2 * DMS_TRACER_CNT, &dout.DMS_tendencies));
    if (diag) {
      // DMS diagnostics are NOT zeroed by the reference: inactive cells keep the caller's
      // values, so the caller's arrays are uploaded first - unless every cell of the chunk is
      // active, in which case every element is overwritten anyway.
      const bool keep = !c->diag_accumulate && !chunk_fully_active(h, in->number_of_active_levels, nCols);
#define UP_D(name) if (diag->name) { if (keep) { RC(up_k(c, h, "dms.d." #name, diag->name, 1, &t)); } \
                                     else { RC(arena_d(c, slot_key(h, "dms.d." #name), n2, &t)); } dd.name = t; }
      DMS_DIAG_LIST(UP_D)
#undef UP_D
    }
    RC(flush_up(c, h));
    RC(dms_source_sink_device(c, &din, &dfo, &dout, diag ? &dd : nullptr, h.nL, h.cc, cols));
    // only DMS and DMSP have tendencies; the other twelve slabs are the constant zero (DMS_mod.F90:413, :718-719)
    const unsigned long long dms_live = (1ull << (c->dms_tab.ind.dms_ind - 1)) | (1ull << (c->dms_tab.ind.dmsp_ind - 1));
    RC(down_k(c, h, dout.DMS_tendencies, out->DMS_tendencies, DMS_TRACER_CNT, ((1ull << DMS_TRACER_CNT) - 1ull) & ~dms_live));
    if (diag && c->diag_accumulate) {
#define AC_D(name) if (dd.name) RC(acc_add(c, h, "dms." #name, dd.name, h.nL, 1, din.number_of_active_levels, cols));
      DMS_DIAG_LIST(AC_D)
#undef AC_D
    } else if (diag) {
#define DN_D(name) if (dd.name) RC(down_k(c, h, dd.name, diag->name, 1));
      DMS_DIAG_LIST(DN_D)
#undef DN_D
    }
    return flush_down(c, h);
  });
}

static int dms_surface_device(bgc_ctx *c, const DmsInput *in, DmsForcing *fo, DmsFluxDiagnostics *diag, int nL,
                              int nC, int nCols) {
  RC(ensure_dms_tables(c));
  if (!fo->lcalc_DMS_gas_flux) return BGC_OK;   // DMS_mod.F90:846: nothing at all happens
  if (!in->DMS_tracers || !fo->SST || !fo->iceFraction || !fo->windSpeedSquared10m || !fo->surfacePressure || !fo->netFlux)
    return fail(BGC_ERR_ARG, "dms_surface_fluxes: a required array pointer is NULL");
  bgc::DmsSurfArgs a;
  a.nL = nL; a.nC = nC; a.nColumns = nCols;
  a.tracers = in->DMS_tracers; a.f = *fo;
  if (diag) a.d = *diag; else memset(&a.d, 0, sizeof a.d);
  LAUNCH(BGC_K_DMS_SURFACE, 1, bgc::launch_dms_surface(a, c->stream));
  return BGC_OK;
}

extern "C" int dms_surface_fluxes(bgc_ctx *c, const DmsInput *in, DmsForcing *fo, DmsFluxDiagnostics *diag,
                                  int nL, int nC, int nCols, int mem_space) {
  RC(use_device(c));
  if (!in || !fo) return fail(BGC_ERR_ARG, "dms_surface_fluxes: null argument block");
  RC(check_dims(c, nL, nC, nCols));
  if (mem_space == BGC_MEM_DEVICE_SOA) return dms_surface_device(c, in, fo, diag, nL, nC, nCols);
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  if (!fo->lcalc_DMS_gas_flux) return BGC_OK;
  HostChunk h; h.nL = nL; h.nC = nC; h.c0 = 0; h.cc = nC; h.slot = 0;   // per-column data only: one chunk
  DmsInput din; DmsForcing dfo = *fo; DmsFluxDiagnostics dd;
  memset(&din, 0, sizeof din); memset(&dd, 0, sizeof dd);
  if (!in->DMS_tracers) return fail(BGC_ERR_ARG, "dms_surface_fluxes: DMS_tracers is NULL");
  double *surf = nullptr;
  RC(arena_d(c, "dmssurf.tracers", (size_t)nC * DMS_TRACER_CNT, &surf));
  RC(gather_level1(c, "dmssurf.tracers", in->DMS_tracers, nL, (size_t)nC * DMS_TRACER_CNT, surf));
  din.DMS_tracers = surf;
  void *v = nullptr;
#define UPC(member, n) do { if (fo->member) { RC(up_c(c, h, "dmssurf." #member, fo->member, sizeof(double), (n), &v)); dfo.member = (double *)v; } } while (0)
  UPC(surfacePressure, 1); UPC(iceFraction, 1); UPC(windSpeedSquared10m, 1); UPC(SST, 1); UPC(SSS, 1);
  UPC(netFlux, DMS_TRACER_CNT);
#undef UPC
  if (diag) {
    // not zeroed by the reference: columns beyond numColumns keep the caller's values
#define UP_F(name) if (diag->name) { RC(up_c(c, h, "dmssurf.d." #name, diag->name, sizeof(double), 1, &v)); dd.name = (double *)v; }
    DMS_FLUX_DIAG_LIST(UP_F)
#undef UP_F
  }
  RC(dms_surface_device(c, &din, &dfo, diag ? &dd : nullptr, 1, nC, nCols));
  RC(down_c(c, h, dfo.iceFraction, fo->iceFraction, sizeof(double), 1));
  RC(down_c(c, h, dfo.netFlux, fo->netFlux, sizeof(double), DMS_TRACER_CNT));
  if (diag) {
#define DN_F(name) if (dd.name) RC(down_c(c, h, dd.name, diag->name, sizeof(double), 1));
    DMS_FLUX_DIAG_LIST(DN_F)
#undef DN_F
  }
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

// ------------------------------------------------------------------ MACROS
static int macros_device(bgc_ctx *c, const MacrosInput *in, MacrosOutput *out, const MacrosDiagnostics *diag,
                         int nL, int nC, int nCols) {
  RC(ensure_macros_tables(c));
  if (!in->MACROS_tracers || !in->number_of_active_levels || !out->MACROS_tendencies)
    return fail(BGC_ERR_ARG, "macros_source_sink: a required array pointer is NULL");
  bgc::MacrosArgs a;
  a.nL = nL; a.nC = nC; a.nColumns = nCols;
  a.tracers = in->MACROS_tracers; a.kmax = in->number_of_active_levels; a.tend = out->MACROS_tendencies;
  if (diag) a.d = *diag; else memset(&a.d, 0, sizeof a.d);
  const bool inv = c->inventory_on && in->cell_thickness;
  a.dz = in->cell_thickness;
  a.inv_partials = nullptr;
  if (inv) RC(arena_d(c, "inv_partials_macros", (size_t)bgc::macros_inventory_parts(nL, nC) * bgc::kInvGroup, &a.inv_partials));
  LAUNCH(BGC_K_MACROS_CELLS, 1, bgc::launch_macros_cells(a, c->stream));
  if (inv) {   // only PROT, POLY and LIP have non-zero tendencies (MACROS_mod.F90:267, :389-391)
    int oi[1][bgc::kInvGroup];
    for (int j = 0; j < bgc::kInvGroup; ++j) oi[0][j] = -1;
    oi[0][0] = 44 + c->macros_tab.ind.prot_ind - 1;
    oi[0][1] = 44 + c->macros_tab.ind.poly_ind - 1;
    oi[0][2] = 44 + c->macros_tab.ind.lip_ind - 1;
    RC(inventory_fold(c, a.inv_partials, bgc::macros_inventory_parts(nL, nC), 1, oi));
  }
  return BGC_OK;
}

extern "C" int macros_source_sink(bgc_ctx *c, const MacrosInput *in, MacrosOutput *out, MacrosDiagnostics *diag,
                                  int nL, int nC, int nCols, int mem_space) {
  RC(use_device(c));
  if (!in || !out) return fail(BGC_ERR_ARG, "macros_source_sink: null argument block");
  RC(check_dims(c, nL, nC, nCols));
  if (mem_space == BGC_MEM_DEVICE_SOA) return macros_device(c, in, out, diag, nL, nC, nCols);
  if (mem_space != BGC_MEM_HOST_FORTRAN) return fail(BGC_ERR_ARG, "unknown mem_space %d", mem_space);
  if (!in->number_of_active_levels) return fail(BGC_ERR_ARG, "macros_source_sink: number_of_active_levels is NULL");
  RC(ensure_macros_tables(c));
  return host_pipeline(c, nL, nC, [&](HostChunk &h) -> int {
    const size_t n2 = (size_t)h.nL * h.cc;
    int cols = nCols - h.c0;
    if (cols < 0) cols = 0;
    if (cols > h.cc) cols = h.cc;
    MacrosInput din; MacrosOutput dout; MacrosDiagnostics dd;
    memset(&din, 0, sizeof din); memset(&dout, 0, sizeof dout); memset(&dd, 0, sizeof dd);
    double *t = nullptr; void *v = nullptr;
    RC(stage_reserve(c, h, MACROS_TRACER_CNT + 1 + sizeof(MacrosDiagnostics) / sizeof(double *)));
    RC(up_k(c, h, "mac.tracers", in->MACROS_tracers, MACROS_TRACER_CNT, &t)); din.MACROS_tracers = t;
    if (c->inventory_on && in->cell_thickness) { RC(up_k(c, h, "mac.dz", in->cell_thickness, 1, &t)); din.cell_thickness = t; }
    RC(up_c(c, h, "mac.kmax", in->number_of_active_levels, sizeof(int), 1, &v)); din.number_of_active_levels = (int *)v;
    RC(arena_d(c, slot_key(h, "mac.tend"), n2 * MACROS_TRACER_CNT, &dout.MACROS_tendencies));
    if (diag) {   // not zeroed by the reference either: see dms_source_sink
      const bool keep = !c->diag_accumulate && !chunk_fully_active(h, in->number_of_active_levels, nCols);
#define UP_D(name) if (diag->name) { if (keep) { RC(up_k(c, h, "mac.d." #name, diag->name, 1, &t)); } \
                                     else { RC(arena_d(c, slot_key(h, "mac.d." #name), n2, &t)); } dd.name = t; }
      MACROS_DIAG_LIST(UP_D)
#undef UP_D
    }
    RC(flush_up(c, h));
    RC(macros_device(c, &din, &dout, diag ? &dd : nullptr, h.nL, h.cc, cols));
    // only PROT, POLY and LIP have tendencies (MACROS_mod.F90:267, :389-391)
    const unsigned long long mac_live = (1ull << (c->macros_tab.ind.prot_ind - 1)) | (1ull << (c->macros_tab.ind.poly_ind - 1)) |
                                        (1ull << (c->macros_tab.ind.lip_ind - 1));
    RC(down_k(c, h, dout.MACROS_tendencies, out->MACROS_tendencies, MACROS_TRACER_CNT, ((1ull << MACROS_TRACER_CNT) - 1ull) & ~mac_live));
    if (diag && c->diag_accumulate) {
#define AC_D(name) if (dd.name) RC(acc_add(c, h, "mac." #name, dd.name, h.nL, 1, din.number_of_active_levels, cols));
      MACROS_DIAG_LIST(AC_D)
#undef AC_D
    } else if (diag) {
#define DN_D(name) if (dd.name) RC(down_k(c, h, dd.name, diag->name, 1));
      MACROS_DIAG_LIST(DN_D)
#undef DN_D
    }
    return flush_down(c, h);
  });
}

// ------------------------------------------------------------------ CUDA graphs
// A model step is a fixed sequence of ~15 dependent launches on two streams; replaying it
// as one CUDA graph removes the launch gaps between them.  Everything the BGC_MEM_DEVICE_SOA
// entry points enqueue is capturable once the ctx is warm (no arena growth, tables uploaded).
struct bgc_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  unsigned long long launches[BGC_KERNEL_ID_COUNT] = {0};   // kernel launches one replay stands for
  bool uses_bgc = false, uses_dms = false, uses_macros = false;   // __constant__ tables its kernels read
};

extern "C" int bgc_graph_capture_begin(bgc_ctx *c) {
  RC(use_device(c));
  if (c->capturing) return fail(BGC_ERR_ARG, "bgc_graph_capture_begin: already capturing");
  if (c->timing_on) return fail(BGC_ERR_ARG, "bgc_graph_capture_begin: disable per-kernel timing first");
  RC(join_pending(c));
  CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
  c->capturing = true;
  c->cap_uses_bgc = c->cap_uses_dms = c->cap_uses_macros = false;
  for (int i = 0; i < BGC_KERNEL_ID_COUNT; ++i) c->capture_base[i] = c->launches[i];
  return BGC_OK;
}

extern "C" int bgc_graph_capture_end(bgc_ctx *c, bgc_graph **out) {
  RC(use_device(c));
  if (!c->capturing || !out) return fail(BGC_ERR_ARG, "bgc_graph_capture_end: not capturing / null out");
  int rc = join_pending(c);   // a deferred carbonate join must close inside the graph
  c->capturing = false;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(c->stream, &g);
  if (rc != BGC_OK) { if (g) cudaGraphDestroy(g); return rc; }
  if (e != cudaSuccess) return fail(BGC_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  bgc_graph *bg = new bgc_graph();
  bg->graph = g;
  bg->uses_bgc = c->cap_uses_bgc; bg->uses_dms = c->cap_uses_dms; bg->uses_macros = c->cap_uses_macros;
  e = cudaGraphInstantiate(&bg->exec, g, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(g); delete bg; return fail(BGC_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
  for (int i = 0; i < BGC_KERNEL_ID_COUNT; ++i) {
    bg->launches[i] = c->launches[i] - c->capture_base[i];
    c->launches[i] = c->capture_base[i];   // captured, not executed: counted per replay instead
  }
  *out = bg;
  return BGC_OK;
}

extern "C" int bgc_graph_launch(bgc_ctx *c, bgc_graph *g) {
  RC(use_device(c));
  if (!g || !g->exec) return fail(BGC_ERR_ARG, "bgc_graph_launch: null graph");
  RC(join_pending(c));
  // the captured kernels read the per-device __constant__ tables: make sure they are still this
  // ctx's current ones (another ctx of the device, or a *_set_params, may have replaced them)
  if (g->uses_bgc) RC(ensure_bgc_tables(c));
  if (g->uses_dms) RC(ensure_dms_tables(c));
  if (g->uses_macros) RC(ensure_macros_tables(c));
  CU(cudaGraphLaunch(g->exec, c->stream));
  for (int i = 0; i < BGC_KERNEL_ID_COUNT; ++i) c->launches[i] += g->launches[i];
  return BGC_OK;
}

extern "C" int bgc_graph_destroy(bgc_graph *g) {
  if (!g) return BGC_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return BGC_OK;
}

// ------------------------------------------------------------------ MPAS tracer layout
static int mpas_map(int nT, const int *slot, bgc::MpasMap *m) {
  if (nT < 1 || nT > bgc::kMpasMaxTracers || !slot) return fail(BGC_ERR_ARG, "MPAS layout: 1..%d tracers and a slot map are required", bgc::kMpasMaxTracers);
  m->nT = nT;
  m->used = 0ull;
  for (int n = 0; n < bgc::kMpasMaxTracers; ++n) {
    m->slot[n] = (n < nT) ? slot[n] : 0;
    if (n < nT && slot[n] > 0) m->used |= 1ull << n;
  }
  for (int n = 0; n < nT; ++n)
    if (slot[n] < 0 || slot[n] > bgc::kMpasMaxTracers) return fail(BGC_ERR_ARG, "MPAS layout: slot %d of tracer %d out of range", slot[n], n);
  return BGC_OK;
}
extern "C" int bgc_layout_mpas_to_soa(bgc_ctx *c, const double *mpas, double *soa, int nT, const int *slot, int nL,
                                      int nC) {
  RC(use_device(c));
  if (!mpas || !soa) return fail(BGC_ERR_ARG, "bgc_layout_mpas_to_soa: null array");
  RC(check_dims(c, nL, nC, nC));
  bgc::MpasMap m;
  RC(mpas_map(nT, slot, &m));
  LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_mpas_to_soa(mpas, soa, m, nL, nC, c->stream));
  return BGC_OK;
}
extern "C" int bgc_layout_soa_to_mpas_weighted(bgc_ctx *c, const double *soa, double *mpas, int nT, const int *slot,
                                               int nL, int nC, double alpha, double beta, const double *weight) {
  RC(use_device(c));
  if (!mpas || !soa) return fail(BGC_ERR_ARG, "bgc_layout_soa_to_mpas: null array");
  RC(check_dims(c, nL, nC, nC));
  bgc::MpasMap m;
  RC(mpas_map(nT, slot, &m));
  LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_soa_to_mpas(soa, mpas, m, nL, nC, alpha, beta, weight, c->stream));
  return BGC_OK;
}
extern "C" int bgc_layout_soa_to_mpas(bgc_ctx *c, const double *soa, double *mpas, int nT, const int *slot, int nL,
                                      int nC, double alpha, double beta) {
  return bgc_layout_soa_to_mpas_weighted(c, soa, mpas, nT, slot, nL, nC, alpha, beta, nullptr);
}

// ------------------------------------------------------------------ device-resident model state
static int state_buffer(bgc_ctx *c, int which, int nL, int nC, double **out, size_t *count) {
  if (which < 0 || which >= BGC_STATE_COUNT) return fail(BGC_ERR_ARG, "bgc_state: unknown field %d", which);
  RC(check_dims(c, nL, nC, nC));
  const bool is3d = which == BGC_STATE_PH_PREV_3D || which == BGC_STATE_PH_PREV_ALT_CO2_3D;
  const size_t n = is3d ? (size_t)nL * nC : (size_t)nC;
  char key[64];
  snprintf(key, sizeof key, "state.%d.%dx%d", which, is3d ? nL : 1, nC);
  const bool fresh = c->arena.find(key) == c->arena.end();
  RC(arena_d(c, key, n, out));
  if (fresh) CU(cudaMemsetAsync(*out, 0, n * sizeof(double), c->stream));   // 0 = "no previous pH"
  if (count) *count = n;
  return BGC_OK;
}

extern "C" int bgc_state_device_ptr(bgc_ctx *c, int which, int nL, int nC, double **dev_ptr) {
  RC(use_device(c));
  if (!dev_ptr) return fail(BGC_ERR_ARG, "bgc_state_device_ptr: null out");
  return state_buffer(c, which, nL, nC, dev_ptr, nullptr);
}

extern "C" int bgc_state_set(bgc_ctx *c, int which, const double *host, int nL, int nC) {
  RC(use_device(c));
  if (!host) return fail(BGC_ERR_ARG, "bgc_state_set: null host array");
  RC(join_pending(c));
  double *dev = nullptr; size_t n = 0;
  RC(state_buffer(c, which, nL, nC, &dev, &n));
  if (n == (size_t)nC) {   // per-column: same layout in both spaces
    CU(cudaMemcpyAsync(dev, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  } else {
    double *st = nullptr;
    RC(arena_d(c, "state.stage", n, &st));
    CU(cudaMemcpyAsync(st, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_transpose(st, dev, nL, nC, 1, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

extern "C" int bgc_state_get(bgc_ctx *c, int which, double *host, int nL, int nC) {
  RC(use_device(c));
  if (!host) return fail(BGC_ERR_ARG, "bgc_state_get: null host array");
  RC(join_pending(c));
  double *dev = nullptr; size_t n = 0;
  RC(state_buffer(c, which, nL, nC, &dev, &n));
  if (n == (size_t)nC) {
    CU(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  } else {
    double *st = nullptr;
    RC(arena_d(c, "state.stage", n, &st));
    LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_transpose(dev, st, nC, nL, 1, c->stream));
    CU(cudaMemcpyAsync(host, st, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

// ------------------------------------------------------------------ diagnostics accumulation
extern "C" int bgc_diag_accumulate_enable(bgc_ctx *c, int enable) {
  RC(use_device(c));
  c->diag_accumulate = enable != 0;
  return BGC_OK;
}

extern "C" int bgc_diag_flush(bgc_ctx *c, BgcDiagnostics *bgc, DmsDiagnostics *dms, MacrosDiagnostics *macros,
                              int nL, int nC, double scale, int reset) {
  RC(use_device(c));
  RC(check_dims(c, nL, nC, nC));
  RC(join_pending(c));
  // one array at a time through a single-slab staging area (flushes are rare)
  auto flush_one = [&](const char *key, double *host, int rows, int nSlabs) -> int {
    if (!host) return BGC_OK;
    const std::string k = std::string("acc.") + key;
    auto it = c->arena.find(k);
    const size_t n = (size_t)rows * nC * nSlabs;
    if (it == c->arena.end() || it->second.bytes < n * sizeof(double))
      return fail(BGC_ERR_ARG, "bgc_diag_flush: nothing accumulated for %s at these extents", key);
    double *acc = (double *)it->second.p;
    // transpose into the staging area (a per-column array is a 1-row transpose), scale the
    // staged copy, download; the accumulator itself keeps the unscaled sum
    HostChunk h; h.nL = rows; h.nC = nC; h.c0 = 0; h.cc = nC; h.slot = 0;
    RC(stage_reserve(c, h, (size_t)nSlabs));
    RC(down_k(c, h, acc, host, nSlabs));
    if (scale != 1.0) LAUNCH(BGC_K_ACCUMULATE, 1, bgc::launch_scale(h.stage, n, scale, c->stream));
    RC(flush_down(c, h));
    if (reset) CU(cudaMemsetAsync(acc, 0, n * sizeof(double), c->stream));
    return BGC_OK;
  };
  if (bgc) {
    BgcDiagnostics d = *bgc;
    d.diag_POC_ACCUM = d.diag_DONr_remin = d.diag_DOPr_remin = nullptr;   // never touched by the reference
#define FL_K2(name) RC(flush_one("bgc." #name, d.name, nL, 1));
#define FL_KA(name) RC(flush_one("bgc." #name, d.name, nL, BGC_AUTOTROPH_CNT));
#define FL_CA(name) RC(flush_one("bgc." #name, d.name, 1, BGC_AUTOTROPH_CNT));
#define FL_C1(name) RC(flush_one("bgc." #name, d.name, 1, 1));
    BGC_DIAG_K2_LIST(FL_K2) BGC_DIAG_KA_LIST(FL_KA) BGC_DIAG_CA_LIST(FL_CA) BGC_DIAG_C1_LIST(FL_C1)
#undef FL_K2
#undef FL_KA
#undef FL_CA
#undef FL_C1
  }
  if (dms) {
#define FL_D(name) RC(flush_one("dms." #name, dms->name, nL, 1));
    DMS_DIAG_LIST(FL_D)
#undef FL_D
  }
  if (macros) {
#define FL_D(name) RC(flush_one("mac." #name, macros->name, nL, 1));
    MACROS_DIAG_LIST(FL_D)
#undef FL_D
  }
  CU(cudaStreamSynchronize(c->stream));
  return BGC_OK;
}

// ------------------------------------------------------------------ multi-GPU
extern "C" int bgc_comm_unique_id(unsigned char id[128]) {
  if (!id) return fail(BGC_ERR_ARG, "null id");
  RC(nccl_load());
  ncclUniqueId_ u;
  NC(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return BGC_OK;
}

extern "C" int bgc_comm_init_rank(bgc_ctx *c, int nranks, int rank, const unsigned char id[128]) {
  RC(use_device(c));
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) return fail(BGC_ERR_ARG, "bgc_comm_init_rank: bad arguments");
  RC(nccl_load());
  ncclUniqueId_ u;
  memcpy(u.internal, id, 128);
  NC(g_nccl.CommInitRank(&c->comm, nranks, u, rank));
  c->nranks = nranks;
  return BGC_OK;
}

// The all-reduce is stream-ordered: bgc_inventory_allreduce_begin enqueues the NCCL all-reduce
// (a device copy for a single rank) and the copy of the 512-byte result into a page-locked
// buffer of the ctx, and returns without waiting; bgc_inventory_allreduce_end waits for that
// copy only.  A model that consumes the inventory once per output interval keeps issuing
// steps without a host synchronisation in between.
extern "C" int bgc_inventory_allreduce_begin(bgc_ctx *c) {
  RC(use_device(c));
  double *buf = nullptr;
  RC(join_pending(c));
  RC(arena_d(c, "inv_reduced", BGC_INVENTORY_LEN, &buf));
  if (!c->h_inventory)
    CU(cudaHostAlloc((void **)&c->h_inventory, BGC_INVENTORY_LEN * sizeof(double), cudaHostAllocDefault));
  if (c->comm) {
    NC(g_nccl.AllReduce(c->d_inventory, buf, BGC_INVENTORY_LEN, kNcclFloat64, kNcclSum, c->comm, c->stream));
  } else {
    CU(cudaMemcpyAsync(buf, c->d_inventory, BGC_INVENTORY_LEN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  }
  CU(cudaMemcpyAsync(c->h_inventory, buf, BGC_INVENTORY_LEN * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  c->inventory_pending = true;
  return BGC_OK;
}

extern "C" int bgc_inventory_allreduce_end(bgc_ctx *c, double out[BGC_INVENTORY_LEN]) {
  RC(use_device(c));
  if (!out) return fail(BGC_ERR_ARG, "null out");
  if (!c->inventory_pending) return fail(BGC_ERR_ARG, "bgc_inventory_allreduce_end without a begin");
  CU(cudaStreamSynchronize(c->stream));   // (also correct when the _begin was replayed from a captured graph)
  memcpy(out, c->h_inventory, BGC_INVENTORY_LEN * sizeof(double));
  c->inventory_pending = false;
  return BGC_OK;
}

extern "C" int bgc_inventory_allreduce(bgc_ctx *c, double out[BGC_INVENTORY_LEN]) {
  if (!out) return fail(BGC_ERR_ARG, "null out");
  RC(bgc_inventory_allreduce_begin(c));
  return bgc_inventory_allreduce_end(c, out);
}

// ------------------------------------------------------------------ host memory, layout helpers
extern "C" int bgc_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return fail(BGC_ERR_ARG, "null ptr");
  CU(cudaHostAlloc(ptr, bytes ? bytes : 8, cudaHostAllocDefault));
  return BGC_OK;
}
extern "C" int bgc_host_free(void *ptr) {
  if (ptr) CU(cudaFreeHost(ptr));
  return BGC_OK;
}
extern "C" int bgc_host_register(void *ptr, size_t bytes) {
  if (!ptr) return fail(BGC_ERR_ARG, "null ptr");
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return BGC_OK;
}
extern "C" int bgc_host_unregister(void *ptr) {
  if (ptr) CU(cudaHostUnregister(ptr));
  return BGC_OK;
}

extern "C" int bgc_layout_to_soa(bgc_ctx *c, const double *dev_fortran, double *dev_soa, int nL, int nC, int nSlabs) {
  RC(use_device(c));
  if (!dev_fortran || !dev_soa) return fail(BGC_ERR_ARG, "null array");
  LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_transpose(dev_fortran, dev_soa, nL, nC, nSlabs, c->stream));
  return BGC_OK;
}
extern "C" int bgc_layout_to_fortran(bgc_ctx *c, const double *dev_soa, double *dev_fortran, int nL, int nC, int nSlabs) {
  RC(use_device(c));
  if (!dev_fortran || !dev_soa) return fail(BGC_ERR_ARG, "null array");
  LAUNCH(BGC_K_TRANSPOSE, 1, bgc::launch_transpose(dev_soa, dev_fortran, nC, nL, nSlabs, c->stream));
  return BGC_OK;
}
