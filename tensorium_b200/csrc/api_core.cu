// api_core.cu — lifetime, memory and stream entry points of the C ABI.
// Replaces TNNCuda<T>.Create/Destroy/createDeviceBuffer/readBuffer/writeBuffer/finish
// (reference: source/nncuda.pas:443-522,1575-1584).
#include <stdarg.h>
#include <map>
#include <set>
#include <tuple>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";
void tnn_wcache_destroy(tnn_context *ctx);

void tnn_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

const char *tnn_last_error(void) { return g_err; }

const char *tnn_version(void) { return "tensorium_b200 0.1 (sm_100a; tcgen05 3xTF32 implicit-GEMM conv)"; }

int32_t tnn_device_count(int32_t *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    tnn_set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    if (count) *count = 0;
    return 2;
  }
  if (count) *count = n;
  return 0;
}

int32_t tnn_create(int32_t device_index, tnn_context **out) {
  if (!out) {
    tnn_set_error("tnn_create: null out pointer");
    return 1;
  }
  *out = nullptr;
  int n = 0;
  TNN_CUDA(cudaGetDeviceCount(&n));
  if (device_index < 0 || device_index >= n) {
    tnn_set_error("tnn_create: device %d out of range (%d devices)", device_index, n);
    return 1;
  }
  tnn_context *ctx = new tnn_context();
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device_index;
  cudaError_t e = cudaGetDeviceProperties(&ctx->prop, device_index);
  if (e != cudaSuccess) {
    delete ctx;
    tnn_set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    return 2;
  }
  if (ctx->prop.major != 10) {
    // no fallback path: this library only carries sm_100a code
    tnn_set_error("tnn_create: device %d (%s, sm_%d%d) is not a Blackwell sm_100 GPU", device_index,
                  ctx->prop.name, ctx->prop.major, ctx->prop.minor);
    delete ctx;
    return 5;
  }
  e = cudaSetDevice(device_index);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    tnn_set_error("tnn_create: %s", cudaGetErrorString(e));
    return 2;
  }
  ctx->own_stream = true;
  ctx->math_mode = TNN_MATH_3XTF32;
  {
    // kernel-selection switch for A/B measurements: TNN_TMA_ACT=0 keeps the gathering producer everywhere
    const char *e = getenv("TNN_TMA_ACT");
    ctx->tma_activations = (e && e[0] == '0') ? 0 : 1;
    const char *g = getenv("TNN_TMA_GEMM");
    ctx->tma_gemm = (g && g[0] >= '0' && g[0] <= '3') ? g[0] - '0' : 1;  // 2 / 3: force 32 / 16 k per stage
    const char *k16 = getenv("TNN_GATHER_KB16");
    ctx->gather_kb16 = (k16 && k16[0] == '1') ? 1 : 0;  // measured slower (173 vs 195 TFLOP/s): off unless asked for
    const char *pk = getenv("TNN_PATCH_KB16");
    ctx->patch_kb16 = (pk && pk[0] == '1') ? 1 : 0;  // measured: no difference (97.3 vs 96.6 TFLOP/s), so the 32-k path stays
    const char *ts = getenv("TNN_TAIL_SPLIT");
    ctx->tail_split = (ts && ts[0] == '0') ? 0 : 1;
    const char *np_ = getenv("TNN_NHWC_PAIR");
    ctx->nhwc_pair = (np_ && np_[0] == '0') ? 0 : 1;
    const char *gn = getenv("TNN_GEMM_NHWC");
    ctx->gemm_nhwc = (gn && gn[0] == '0') ? 0 : 1;
    const char *ba = getenv("TNN_NHWC_BN_AUTO");
    ctx->nhwc_bn_auto = (ba && ba[0] == '0') ? 0 : 1;
    const char *ni = getenv("TNN_NHWC_IMAGE");
    ctx->nhwc_image = (ni && ni[0] == '0') ? 0 : 1;
    const char *es = getenv("TNN_NHWC_EPI_SETS");
    ctx->nhwc_epi_sets = es ? atoi(es) : 0;
    if (ctx->nhwc_epi_sets < 0 || ctx->nhwc_epi_sets > 2) ctx->nhwc_epi_sets = 0;
    const char *pd = getenv("TNN_PDL");
    ctx->pdl = (pd && pd[0] == '0') ? 0 : 1;
    const char *nk = getenv("TNN_NHWC_KB");
    ctx->nhwc_kb = (nk && atoi(nk) == 32) ? 32 : (nk && atoi(nk) == 16) ? 16 : 0;
  }
  *out = ctx;
  return 0;
}

int32_t tnn_destroy(tnn_context *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->red_partials) cudaFree(ctx->red_partials);
  if (ctx->red_counters) cudaFree(ctx->red_counters);
  if (ctx->gemm_scratch) cudaFree(ctx->gemm_scratch);
  if (ctx->work) cudaFree(ctx->work);
  for (int i = 0; i < ctx->n_retired; ++i) cudaFree(ctx->retired[i]);
  tnn_wcache_destroy(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

const char *tnn_device_name(tnn_context *ctx) { return ctx ? ctx->prop.name : ""; }

int32_t tnn_device_info(tnn_context *ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor) {
  TNN_CHECK_CTX(ctx);
  if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
  if (cc_major) *cc_major = ctx->prop.major;
  if (cc_minor) *cc_minor = ctx->prop.minor;
  return 0;
}

int32_t tnn_set_stream(tnn_context *ctx, void *cuda_stream) {
  TNN_CHECK_CTX(ctx);
  if (ctx->own_stream) {
    TNN_CUDA(cudaStreamSynchronize(ctx->stream));
    TNN_CUDA(cudaStreamDestroy(ctx->stream));
    ctx->own_stream = false;
  }
  ctx->stream = (cudaStream_t)cuda_stream;
  return 0;
}

void *tnn_get_stream(tnn_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int32_t tnn_set_math_mode(tnn_context *ctx, int32_t mode) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(mode >= TNN_MATH_FP32 && mode <= TNN_MATH_BF16X3, "unknown math mode");
  ctx->math_mode = mode;
  return 0;
}

int32_t tnn_get_math_mode(tnn_context *ctx) { return ctx ? ctx->math_mode : -1; }

int64_t tnn_launch_count(tnn_context *ctx) { return ctx ? ctx->launches : 0; }

int32_t tnn_create_device_buffer(tnn_context *ctx, int64_t bytes, float **mem) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(mem != nullptr && bytes >= 0, "bad arguments");
  *mem = nullptr;
  if (bytes == 0) return 0;
  TNN_CUDA(cudaSetDevice(ctx->device));
  TNN_CUDA(cudaMalloc((void **)mem, (size_t)bytes));
  return 0;
}

void tnn_wcache_forget(tnn_context *ctx, const float *mem);

int32_t tnn_free_device_buffer(tnn_context *ctx, float *mem) {
  TNN_CHECK_CTX(ctx);
  if (mem) {
    tnn_wcache_forget(ctx, mem);  // a later allocation may reuse the address
    TNN_CUDA(cudaFree(mem));
  }
  return 0;
}

int32_t tnn_read_buffer(tnn_context *ctx, const float *mem, size_t bytes, void *host) {
  TNN_CHECK_CTX(ctx);
  TNN_CUDA(cudaMemcpyAsync(host, mem, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TNN_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t tnn_write_buffer(tnn_context *ctx, float *mem, size_t bytes, const void *host) {
  TNN_CHECK_CTX(ctx);
  TNN_CUDA(cudaMemcpyAsync(mem, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TNN_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t tnn_read_buffer_async(tnn_context *ctx, const float *mem, size_t bytes, void *host) {
  TNN_CHECK_CTX(ctx);
  TNN_CUDA(cudaMemcpyAsync(host, mem, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

int32_t tnn_write_buffer_async(tnn_context *ctx, float *mem, size_t bytes, const void *host) {
  TNN_CHECK_CTX(ctx);
  TNN_CUDA(cudaMemcpyAsync(mem, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

struct tnn_event {
  cudaEvent_t ev;
};

int32_t tnn_event_create(tnn_context *ctx, tnn_event **ev) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(ev != nullptr, "null out pointer");
  tnn_event *e = new tnn_event();
  cudaError_t r = cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming);
  if (r != cudaSuccess) {
    delete e;
    tnn_set_error("cudaEventCreate failed: %s", cudaGetErrorString(r));
    return 2;
  }
  *ev = e;
  return 0;
}

int32_t tnn_event_record(tnn_context *ctx, tnn_event *ev) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(ev != nullptr, "null event");
  TNN_CUDA(cudaEventRecord(ev->ev, ctx->stream));
  return 0;
}

int32_t tnn_stream_wait_event(tnn_context *ctx, tnn_event *ev) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(ev != nullptr, "null event");
  TNN_CUDA(cudaStreamWaitEvent(ctx->stream, ev->ev, 0));
  return 0;
}

int32_t tnn_event_destroy(tnn_event *ev) {
  if (ev) {
    cudaEventDestroy(ev->ev);
    delete ev;
  }
  return 0;
}

struct tnn_graph {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  int64_t launches;  // kernel launches captured (tnn_launch_count advances by this much per replay)
  tnn_context *ctx;
};

int32_t tnn_graph_begin_capture(tnn_context *ctx) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(!ctx->capturing, "already capturing");
  TNN_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = 1;
  ctx->capture_launches0 = ctx->launches;
  return 0;
}

int32_t tnn_graph_end_capture(tnn_context *ctx, tnn_graph **graph) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(ctx->capturing && graph != nullptr, "not capturing");
  ctx->capturing = 0;
  tnn_graph *g = new tnn_graph();
  g->launches = ctx->launches - ctx->capture_launches0;
  ctx->launches = ctx->capture_launches0;  // nothing ran yet
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete g;
    tnn_set_error("graph capture failed: %s", cudaGetErrorString(e));
    return 2;
  }
  g->ctx = ctx;
  ctx->live_graphs++;
  *graph = g;
  return 0;
}

int32_t tnn_graph_launch(tnn_context *ctx, tnn_graph *graph) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(graph != nullptr, "null graph");
  TNN_CUDA(cudaGraphLaunch(graph->exec, ctx->stream));
  ctx->launches += graph->launches;
  return 0;
}

int32_t tnn_graph_destroy(tnn_graph *graph) {
  if (graph) {
    if (graph->ctx && graph->ctx->live_graphs > 0) graph->ctx->live_graphs--;
    cudaGraphExecDestroy(graph->exec);
    cudaGraphDestroy(graph->graph);
    delete graph;
  }
  return 0;
}

int32_t tnn_event_synchronize(tnn_event *ev) {
  if (!ev) {
    tnn_set_error("tnn_event_synchronize: null event");
    return 1;
  }
  TNN_CUDA(cudaEventSynchronize(ev->ev));
  return 0;
}

int32_t tnn_host_alloc(tnn_context *ctx, int64_t bytes, void **host) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(host != nullptr && bytes > 0, "bad arguments");
  *host = nullptr;
  TNN_CUDA(cudaSetDevice(ctx->device));
  TNN_CUDA(cudaHostAlloc(host, (size_t)bytes, cudaHostAllocDefault));
  return 0;
}

int32_t tnn_host_free(tnn_context *ctx, void *host) {
  TNN_CHECK_CTX(ctx);
  if (host) TNN_CUDA(cudaFreeHost(host));
  return 0;
}

int32_t tnn_finish(tnn_context *ctx) {
  TNN_CHECK_CTX(ctx);
  TNN_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"

int tnn_release_scratch(tnn_context *ctx, void *p) {
  if (ctx->live_graphs > 0) {
    // an instantiated graph may replay kernels that read or write this buffer: keep it alive
    TNN_REQUIRE(ctx->n_retired < (int)(sizeof(ctx->retired) / sizeof(ctx->retired[0])),
                "too many scratch buffers retired while CUDA graphs are alive: destroy the graphs first");
    ctx->retired[ctx->n_retired++] = p;
    return 0;
  }
  TNN_CUDA(cudaFree(p));
  return 0;
}

int tnn_reserve_reduction(tnn_context *ctx, size_t partial_floats, size_t counters) {
  TNN_REQUIRE(!ctx->capturing || (ctx->red_partials_bytes >= partial_floats * sizeof(float) && ctx->red_counters_n >= counters),
              "scratch buffer would grow during graph capture: run the step once eagerly first");
  if (ctx->red_partials_bytes < partial_floats * sizeof(float)) {
    // the old buffer may still be in use by kernels queued on the stream
    TNN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->red_partials) {
      int rc = tnn_release_scratch(ctx, ctx->red_partials);
      if (rc) return rc;
    }
    ctx->red_partials = nullptr;
    size_t bytes = partial_floats * sizeof(float) * 2;
    TNN_CUDA(cudaMalloc((void **)&ctx->red_partials, bytes));
    ctx->red_partials_bytes = bytes;
  }
  if (ctx->red_counters_n < counters) {
    TNN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->red_counters) {
      int rc = tnn_release_scratch(ctx, ctx->red_counters);
      if (rc) return rc;
    }
    ctx->red_counters = nullptr;
    size_t n = counters * 2;
    TNN_CUDA(cudaMalloc((void **)&ctx->red_counters, n * sizeof(unsigned int)));
    TNN_CUDA(cudaMemsetAsync(ctx->red_counters, 0, n * sizeof(unsigned int), ctx->stream));
    ctx->red_counters_n = n;
  }
  return 0;
}

struct tnn_wcache {
  std::set<const float *> statics;
  std::map<std::tuple<const float *, int, int, size_t>, float *> copies;  // (weights, kind, math mode, count) -> buffer
};

static void wcache_drop(tnn_context *ctx, const float *weights) {
  if (!ctx->wcache) return;
  for (auto it = ctx->wcache->copies.begin(); it != ctx->wcache->copies.end();) {
    if (std::get<0>(it->first) == weights) {
      cudaFree(it->second);
      it = ctx->wcache->copies.erase(it);
    } else {
      ++it;
    }
  }
}

extern "C" int32_t tnn_set_weights_static(tnn_context *ctx, const float *weights, int32_t is_static) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(weights != nullptr, "null weights");
  if (!ctx->wcache) ctx->wcache = new tnn_wcache();
  TNN_CUDA(cudaStreamSynchronize(ctx->stream));  // kernels still reading an old prepared copy
  wcache_drop(ctx, weights);                     // (re)declaring invalidates what was prepared from the old contents
  if (is_static) ctx->wcache->statics.insert(weights);
  else ctx->wcache->statics.erase(weights);
  return 0;
}

void tnn_wcache_forget(tnn_context *ctx, const float *mem) {
  if (!ctx->wcache || !ctx->wcache->statics.count(mem)) return;
  cudaStreamSynchronize(ctx->stream);
  wcache_drop(ctx, mem);
  ctx->wcache->statics.erase(mem);
}

void tnn_wcache_destroy(tnn_context *ctx) {
  if (!ctx->wcache) return;
  for (auto &kv : ctx->wcache->copies) cudaFree(kv.second);
  delete ctx->wcache;
  ctx->wcache = nullptr;
}

int tnn_weight_buffer(tnn_context *ctx, const float *weights, int kind, size_t count, float **hi, float **lo, bool *fresh) {
  if (ctx->wcache && ctx->wcache->statics.count(weights)) {
    const auto key = std::make_tuple(weights, kind, ctx->math_mode, count);
    auto it = ctx->wcache->copies.find(key);
    if (it != ctx->wcache->copies.end()) {
      *hi = it->second;
      *lo = it->second + count;
      *fresh = false;
      return 0;
    }
    TNN_REQUIRE(!ctx->capturing, "a static weight copy would be allocated during graph capture: run the step once first");
    float *buf = nullptr;
    TNN_CUDA(cudaMalloc((void **)&buf, 2 * count * sizeof(float)));
    ctx->wcache->copies[key] = buf;
    *hi = buf;
    *lo = buf + count;
    *fresh = true;
    return 0;
  }
  int rc = tnn_reserve_work(ctx, 2 * count * sizeof(float));
  if (rc) return rc;
  *hi = ctx->work;
  *lo = ctx->work + count;
  *fresh = true;
  return 0;
}

int tnn_reserve_work(tnn_context *ctx, size_t bytes) {
  if (ctx->work_bytes < bytes) {
    TNN_REQUIRE(!ctx->capturing, "scratch buffer would grow during graph capture: run the step once eagerly first");
    TNN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->work) {
      int rc = tnn_release_scratch(ctx, ctx->work);
      if (rc) return rc;
    }
    ctx->work = nullptr;
    TNN_CUDA(cudaMalloc((void **)&ctx->work, bytes));
    ctx->work_bytes = bytes;
  }
  return 0;
}
