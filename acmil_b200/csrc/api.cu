// C-ABI entry points (include/acmil_b200.h).  No torch types, no exceptions, caller-owned memory.
#include <stdarg.h>

#include <mutex>
#include <utility>
#include <vector>

#include "gp_common.cuh"

// gp_umma.cu
int gp_launch_main_umma(const GpMainParams& p, const acmil_gp_consts* consts, const unsigned char* d_umma, cudaStream_t st);
int gp_umma_supported(const acmil_gp_shape& s);
int gp_umma_pack(const acmil_gp_shape& s, const acmil_gp_weights& w, unsigned char* d_umma, acmil_gp_consts* consts,
                 const float* d_f32, const GpPackLayout& lay, cudaStream_t st);
int gp_umma_build_plan(const acmil_gp_batch& b, int sm_count, GpSegTable* t);

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_acmil_launches{0};

void acmil_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#define RT_FINISH_THREADS 256

namespace {

// row-pass timing for bench.py (acmil_prof_enable / acmil_prof_collect): a diagnostic, guarded by a mutex so that host
// threads driving different streams cannot corrupt the event pool
struct ProfState {
  std::mutex mu;
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  size_t used = 0;
} g_prof;

// SM count of the CURRENT device (cached per device ordinal: a process may drive several GPUs)
int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

int check_shape(const acmil_gp_shape* s) {
  ACMIL_REQUIRE(s != nullptr, ACMIL_E_INVALID, "shape is NULL");
  ACMIL_REQUIRE(s->n_branch >= 1 && s->n_branch <= KMAX, ACMIL_E_INVALID, "n_branch must be in [1, %d] (got %d)", KMAX,
                s->n_branch);
  ACMIL_REQUIRE(s->d_attn == GP_DATTN, ACMIL_E_UNSUPPORTED, "d_attn must be %d (got %d)", GP_DATTN, s->d_attn);
  ACMIL_REQUIRE(s->d_in > 0 && s->d_inner > 0, ACMIL_E_INVALID, "bad widths d_in=%d d_inner=%d", s->d_in, s->d_inner);
  ACMIL_REQUIRE(s->front == 0 || s->front == 1, ACMIL_E_INVALID, "front must be 0 or 1");
  ACMIL_REQUIRE(s->front_act == ACMIL_ACT_RELU || s->front_act == ACMIL_ACT_GELU, ACMIL_E_INVALID,
                "front_act must be relu or gelu");
  ACMIL_REQUIRE(s->act_a >= 0 && s->act_a <= 2, ACMIL_E_INVALID, "act_a out of range");
  return ACMIL_OK;
}

int check_batch(const acmil_gp_batch* b) {
  ACMIL_REQUIRE(b != nullptr && b->row_offsets != nullptr, ACMIL_E_INVALID, "batch / row_offsets is NULL");
  ACMIL_REQUIRE(b->n_slides >= 0 && b->n_slides <= SMAX, ACMIL_E_INVALID, "n_slides must be in [0, %d] (got %d)", SMAX,
                b->n_slides);
  ACMIL_REQUIRE(b->n_masked >= 0 && b->n_masked <= NMAX, ACMIL_E_INVALID, "n_masked must be in [0, %d] (got %d)", NMAX,
                b->n_masked);
  ACMIL_REQUIRE(b->row_offsets[0] == 0, ACMIL_E_INVALID, "row_offsets[0] must be 0");
  for (int s = 0; s < b->n_slides; ++s)
    ACMIL_REQUIRE(b->row_offsets[s + 1] >= b->row_offsets[s], ACMIL_E_INVALID, "row_offsets must be non-decreasing");
  ACMIL_REQUIRE(b->row_offsets[b->n_slides] < (int64_t)0x7fffffff, ACMIL_E_INVALID, "too many rows in one batch");
  return ACMIL_OK;
}

int pick_impl(const acmil_gp_shape& s, int impl) {
  if (impl == ACMIL_IMPL_FFMA) return ACMIL_IMPL_FFMA;
  if (impl == ACMIL_IMPL_UMMA) return ACMIL_IMPL_UMMA;
  return gp_umma_supported(s) ? ACMIL_IMPL_UMMA : ACMIL_IMPL_FFMA;
}

size_t rescue_offset(size_t umma_ws_bytes) { return (umma_ws_bytes + 1023) / 1024 * 1024; }

int plan(const acmil_gp_shape& s, const acmil_gp_batch& b, int impl, GpSegTable* seg, GpWorkspace* wl) {
  if (impl == ACMIL_IMPL_UMMA)
    ACMIL_REQUIRE(gp_umma_build_plan(b, sm_count(), seg) == 0, ACMIL_E_INVALID, "bad row_offsets");
  else
    ACMIL_REQUIRE(gp_build_segments(b, 64, 4 * sm_count(), s.n_branch, seg) == 0, ACMIL_E_INVALID, "bad row_offsets");
  *wl = gp_workspace_layout(s, *seg);
  return ACMIL_OK;
}

__global__ void gp_pack_f32_kernel(acmil_gp_shape s, acmil_gp_weights w, GpPackLayout l, float* __restrict__ out) {
  const size_t total = l.f32_floats;
  const size_t L = s.d_inner, D = GP_DATTN, DIN = s.d_in;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < l.b1) {  // w1t[k][n] = W1[n][k]
      const size_t k = i / L, n = i % L;
      v = w.d_w1[n * DIN + k];
    } else if (i < l.wvt) {
      v = (s.front && s.front_bias && w.d_b1) ? w.d_b1[i - l.b1] : 0.f;
    } else if (i < l.wut) {
      const size_t j = i - l.wvt, k = j / D, n = j % D;
      v = w.d_wv[n * L + k];
    } else if (i < l.bv) {
      const size_t j = i - l.wut, k = j / D, n = j % D;
      v = (s.gated && w.d_wu) ? w.d_wu[n * L + k] : 0.f;
    } else if (i < l.bu) {
      v = (s.gate_bias && w.d_bv) ? w.d_bv[i - l.bv] : 0.f;
    } else if (i < l.ww) {
      v = (s.gated && s.gate_bias && w.d_bu) ? w.d_bu[i - l.bu] : 0.f;
    } else if (i < l.bw) {
      const size_t j = i - l.ww, k = j / D, n = j % D;
      v = k < (size_t)s.n_branch ? w.d_ww[k * D + n] : 0.f;
    } else {
      const size_t k = i - l.bw;
      v = (s.score_bias && w.d_bw && k < (size_t)s.n_branch) ? w.d_bw[k] : 0.f;
    }
    out[i] = v;
  }
}

}  // namespace

extern "C" {

const char* acmil_last_error(void) { return g_err; }
int acmil_abi_version(void) { return ACMIL_ABI_VERSION; }
int acmil_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int64_t acmil_launch_count(void) { return g_acmil_launches.load(); }

int acmil_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.on = on != 0;
  if (!on) g_prof.used = 0;
  return ACMIL_OK;
}

int acmil_prof_collect(double* main_ms_sum, int64_t* n_launches) {
  double sum = 0.0;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (size_t i = 0; i < g_prof.used; ++i) {
    ACMIL_CHECK_CUDA(cudaEventSynchronize(g_prof.pool[i].second));
    float ms = 0.f;
    ACMIL_CHECK_CUDA(cudaEventElapsedTime(&ms, g_prof.pool[i].first, g_prof.pool[i].second));
    sum += ms;
  }
  if (main_ms_sum) *main_ms_sum = sum;
  if (n_launches) *n_launches = (int64_t)g_prof.used;
  g_prof.used = 0;
  return ACMIL_OK;
}

int acmil_gp_packed_bytes(const acmil_gp_shape* shape, size_t* bytes) {
  if (int rc = check_shape(shape)) return rc;
  ACMIL_REQUIRE(bytes != nullptr, ACMIL_E_INVALID, "bytes is NULL");
  *bytes = gp_pack_layout(*shape).total_bytes;
  return ACMIL_OK;
}

int acmil_gp_umma_supported(const acmil_gp_shape* shape) { return shape != nullptr && gp_umma_supported(*shape); }

int acmil_gp_pack(const acmil_gp_shape* shape, const acmil_gp_weights* w, void* d_packed, size_t packed_bytes,
                  acmil_gp_consts* consts, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  ACMIL_REQUIRE(w != nullptr && d_packed != nullptr, ACMIL_E_INVALID, "weights / d_packed is NULL");
  ACMIL_REQUIRE(w->d_wv != nullptr && w->d_ww != nullptr, ACMIL_E_INVALID, "d_wv and d_ww are required");
  ACMIL_REQUIRE(!shape->front || w->d_w1 != nullptr, ACMIL_E_INVALID, "front == 1 needs d_w1");
  ACMIL_REQUIRE(!shape->gated || w->d_wu != nullptr, ACMIL_E_INVALID, "gated == 1 needs d_wu");
  const GpPackLayout l = gp_pack_layout(*shape);
  ACMIL_REQUIRE(packed_bytes >= l.total_bytes, ACMIL_E_WORKSPACE, "packed buffer too small: %zu < %zu", packed_bytes,
                l.total_bytes);
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  cudaStream_t st = (cudaStream_t)stream;
  gp_pack_f32_kernel<<<296, 256, 0, st>>>(*shape, *w, l, reinterpret_cast<float*>(d_packed));
  ++g_acmil_launches;
  ACMIL_CHECK_CUDA(cudaGetLastError());
  if (consts) consts->valid = 0;
  if (gp_umma_supported(*shape)) {
    if (int rc = gp_umma_pack(*shape, *w, reinterpret_cast<unsigned char*>(d_packed) + l.umma_off, consts,
                              reinterpret_cast<const float*>(d_packed), l, st))
      return rc;
  }
  return ACMIL_OK;
}

int acmil_gp_sizes(const acmil_gp_shape* shape, const acmil_gp_batch* batch, int impl, size_t* workspace_bytes,
                   size_t* partial_bytes) {
  if (int rc = check_shape(shape)) return rc;
  if (int rc = check_batch(batch)) return rc;
  GpSegTable seg;
  GpWorkspace wl;
  if (int rc = plan(*shape, *batch, pick_impl(*shape, impl), &seg, &wl)) return rc;
  size_t ws = wl.total_bytes;
  if (pick_impl(*shape, impl) == ACMIL_IMPL_UMMA) {
    // the tcgen05 kernel's workspace is followed by the FFMA kernel's: the rescue pass of bags that ran out of parking
    // slots works there (and AUTO may run the FFMA kernel alone when no host constants were supplied)
    if (int rc = plan(*shape, *batch, ACMIL_IMPL_FFMA, &seg, &wl)) return rc;
    ws = rescue_offset(ws) + wl.total_bytes;
  }
  if (workspace_bytes) *workspace_bytes = ws;
  if (partial_bytes) *partial_bytes = gp_record(*shape, batch->n_masked).stride() * 4 * (size_t)batch->n_slides;
  return ACMIL_OK;
}

static int to_device_exchange(const acmil_gp_exchange* x, const GpRecord& rec, int n_slides, GpExchange* gx) {
  ACMIL_REQUIRE(x->n_ranks >= 1 && x->n_ranks <= ACMIL_MAX_PEERS, ACMIL_E_INVALID, "exchange: n_ranks must be in [1, %d]",
                ACMIL_MAX_PEERS);
  ACMIL_REQUIRE(x->rank >= 0 && x->rank < x->n_ranks, ACMIL_E_INVALID, "exchange: bad rank");
  ACMIL_REQUIRE(x->d_epoch && x->d_ticket, ACMIL_E_INVALID, "exchange: d_epoch / d_ticket is NULL");
  ACMIL_REQUIRE(x->gather_bytes >= 2 * (size_t)x->n_ranks * rec.stride() * 4 * (size_t)n_slides, ACMIL_E_WORKSPACE,
                "exchange: gather buffer too small");
  memset(gx, 0, sizeof(*gx));
  gx->n_ranks = x->n_ranks;
  gx->rank = x->rank;
  for (int r = 0; r < x->n_ranks; ++r) {
    ACMIL_REQUIRE(x->d_gather[r] && x->d_flags[r], ACMIL_E_INVALID, "exchange: peer pointer %d is NULL", r);
    gx->gather[r] = reinterpret_cast<float*>(x->d_gather[r]);
    gx->flags[r] = x->d_flags[r];
  }
  gx->epoch = x->d_epoch;
  gx->ticket = x->d_ticket;
  return ACMIL_OK;
}

static int gp_partial_common(const acmil_gp_shape* shape, const void* d_packed, const acmil_gp_consts* consts,
                             const acmil_gp_batch* batch, int impl, void* d_workspace, size_t workspace_bytes,
                             void* d_partial, size_t partial_bytes, const acmil_gp_exchange* x, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  if (int rc = check_batch(batch)) return rc;
  ACMIL_REQUIRE(d_packed && d_workspace && (d_partial || x), ACMIL_E_INVALID, "NULL device buffer");
  ACMIL_REQUIRE(batch->d_x != nullptr || batch->row_offsets[batch->n_slides] == 0, ACMIL_E_INVALID, "d_x is NULL");
  ACMIL_REQUIRE(batch->d_a_out == nullptr || batch->a_ld >= batch->row_offsets[batch->n_slides], ACMIL_E_INVALID,
                "a_ld smaller than the number of rows");
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  int use = pick_impl(*shape, impl);
  if (impl == ACMIL_IMPL_AUTO && use == ACMIL_IMPL_UMMA && shape->n_branch > 6 && batch->n_masked > 0)
    use = ACMIL_IMPL_FFMA;   // masking with > 6 branches: stay on the general kernel
  ACMIL_REQUIRE(use != ACMIL_IMPL_UMMA || gp_umma_supported(*shape), ACMIL_E_UNSUPPORTED,
                "tcgen05 kernel does not support this shape");
  ACMIL_REQUIRE(batch->d_z == nullptr || (shape->front == 0 && batch->x_f16 == 0 && use == ACMIL_IMPL_FFMA &&
                                          ((uintptr_t)batch->d_z & 15) == 0),
                ACMIL_E_INVALID, "d_z needs front == 0, fp32 rows, the FFMA kernel and a 16-byte aligned pointer");
  GpMainParams p;
  memset(&p, 0, sizeof(p));
  p.sh = *shape;
  if (int rc = plan(*shape, *batch, use, &p.seg, &p.wl)) return rc;
  const GpRecord rec = gp_record(*shape, batch->n_masked);
  ACMIL_REQUIRE(workspace_bytes >= p.wl.total_bytes, ACMIL_E_WORKSPACE, "workspace too small: %zu < %zu",
                workspace_bytes, p.wl.total_bytes);
  ACMIL_REQUIRE(x != nullptr || partial_bytes >= rec.stride() * 4 * (size_t)batch->n_slides, ACMIL_E_WORKSPACE,
                "partial buffer too small");
  GpExchange gx;
  const bool rescue = use == ACMIL_IMPL_UMMA && batch->n_masked > 0;
  if (x) {
    if (int rc = to_device_exchange(x, rec, batch->n_slides, &gx)) return rc;
    gx.reduce_ctas = batch->n_slides * shape->n_branch * (rescue ? 2 : 1);
  }
  const GpExchange* gxp = x ? &gx : nullptr;
  p.x = batch->d_x;
  p.x_f16 = batch->x_f16 != 0;
  p.z = batch->d_z;
  p.a_out = batch->d_a_out;
  p.a_ld = batch->a_ld;
  p.pack = reinterpret_cast<const float*>(d_packed);
  p.lay = gp_pack_layout(*shape);
  p.ws = reinterpret_cast<unsigned char*>(d_workspace);
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.on) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (g_prof.used == g_prof.pool.size()) {
      ACMIL_CHECK_CUDA(cudaEventCreate(&e0));
      ACMIL_CHECK_CUDA(cudaEventCreate(&e1));
      g_prof.pool.emplace_back(e0, e1);
    }
    e0 = g_prof.pool[g_prof.used].first;
    e1 = g_prof.pool[g_prof.used].second;
    ++g_prof.used;
    ACMIL_CHECK_CUDA(cudaEventRecord(e0, st));
  }
  int rc = use == ACMIL_IMPL_UMMA
               ? gp_launch_main_umma(p, consts, reinterpret_cast<const unsigned char*>(d_packed) + p.lay.umma_off, st)
               : gp_launch_main_ffma(p, st);
  if (rc) return rc;
  if (e1) ACMIL_CHECK_CUDA(cudaEventRecord(e1, st));
  if (!rescue) return gp_launch_reduce(p, rec, reinterpret_cast<float*>(d_partial), nullptr, GP_REDUCE_ALL, gxp, st);
  // Training-mode masking on the tcgen05 kernel: a bag whose score order defeats the bounded parking scratch (e.g. rows
  // sorted by ascending score) is flagged per bag; the exact FFMA kernel redoes exactly those bags (its CTAs return at
  // once otherwise) and each reduce launch takes the bags of its side.  No host round trip, graph-capturable.
  const int* d_flags = reinterpret_cast<const int*>(p.ws + p.wl.flags);
  static thread_local GpMainParams q;      // large: keep off the stack
  q = p;
  if ((rc = plan(*shape, *batch, ACMIL_IMPL_FFMA, &q.seg, &q.wl))) return rc;
  q.ws = p.ws + rescue_offset(p.wl.total_bytes);
  ACMIL_REQUIRE(workspace_bytes >= rescue_offset(p.wl.total_bytes) + q.wl.total_bytes, ACMIL_E_WORKSPACE,
                "workspace too small for the rescue pass: %zu < %zu", workspace_bytes,
                rescue_offset(p.wl.total_bytes) + q.wl.total_bytes);
  q.rescue_flags = d_flags;
  if ((rc = gp_launch_main_ffma(q, st))) return rc;
  if ((rc = gp_launch_reduce(p, rec, reinterpret_cast<float*>(d_partial), d_flags, GP_REDUCE_UNFLAGGED, gxp, st))) return rc;
  return gp_launch_reduce(q, rec, reinterpret_cast<float*>(d_partial), d_flags, GP_REDUCE_FLAGGED, gxp, st);
}

int acmil_gp_partial(const acmil_gp_shape* shape, const void* d_packed, const acmil_gp_consts* consts,
                     const acmil_gp_batch* batch, int impl, void* d_workspace, size_t workspace_bytes, void* d_partial,
                     size_t partial_bytes, void* stream) {
  ACMIL_REQUIRE(d_partial != nullptr, ACMIL_E_INVALID, "d_partial is NULL");
  return gp_partial_common(shape, d_packed, consts, batch, impl, d_workspace, workspace_bytes, d_partial, partial_bytes,
                           nullptr, stream);
}

int acmil_gp_partial_x(const acmil_gp_shape* shape, const void* d_packed, const acmil_gp_consts* consts,
                       const acmil_gp_batch* batch, int impl, void* d_workspace, size_t workspace_bytes,
                       const acmil_gp_exchange* x, void* stream) {
  ACMIL_REQUIRE(x != nullptr, ACMIL_E_INVALID, "exchange is NULL");
  return gp_partial_common(shape, d_packed, consts, batch, impl, d_workspace, workspace_bytes, nullptr, 0, x, stream);
}

int acmil_gp_overflow_flags(const acmil_gp_shape* shape, const acmil_gp_batch* batch, int impl, const void* d_workspace,
                            int32_t* host_flags, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  if (int rc = check_batch(batch)) return rc;
  ACMIL_REQUIRE(d_workspace && host_flags, ACMIL_E_INVALID, "NULL argument");
  for (int s = 0; s < batch->n_slides; ++s) host_flags[s] = 0;
  if (pick_impl(*shape, impl) != ACMIL_IMPL_UMMA || batch->n_masked == 0) return ACMIL_OK;
  GpSegTable seg;
  GpWorkspace wl;
  if (int rc = plan(*shape, *batch, ACMIL_IMPL_UMMA, &seg, &wl)) return rc;
  int32_t tmp[SMAX];
  cudaStream_t st = (cudaStream_t)stream;
  ACMIL_CHECK_CUDA(cudaMemcpyAsync(tmp, reinterpret_cast<const unsigned char*>(d_workspace) + wl.flags,
                                   sizeof(int32_t) * batch->n_slides, cudaMemcpyDeviceToHost, st));
  ACMIL_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int s = 0; s < batch->n_slides; ++s) host_flags[s] = tmp[s] == 1 ? 1 : 0;
  return ACMIL_OK;
}

static int gp_finish_common(const acmil_gp_shape* shape, const acmil_gp_batch* batch, const void* d_partials,
                            size_t partial_bytes, int n_ranks, const int32_t* keep, const int64_t* d_rsel, int32_t keep_ld,
                            const float* d_rand, int32_t rand_ld, const acmil_gp_heads* heads, const acmil_gp_outputs* out,
                            void* stream, const acmil_gp_exchange* x = nullptr) {
  if (int rc = check_shape(shape)) return rc;
  if (int rc = check_batch(batch)) return rc;
  ACMIL_REQUIRE((d_partials || x) && heads && out, ACMIL_E_INVALID, "NULL argument");
  ACMIL_REQUIRE(n_ranks >= 1 && n_ranks <= 64, ACMIL_E_INVALID, "n_ranks must be in [1, 64]");
  ACMIL_REQUIRE(heads->n_class >= 0 && heads->n_class <= ACMIL_MAX_CLASS, ACMIL_E_INVALID, "n_class must be <= %d",
                ACMIL_MAX_CLASS);
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  GpFinishParams p;
  memset(&p, 0, sizeof(p));
  p.sh = *shape;
  p.rec = gp_record(*shape, batch->n_masked);
  ACMIL_REQUIRE(x != nullptr || partial_bytes >= p.rec.stride() * 4 * (size_t)batch->n_slides * n_ranks, ACMIL_E_WORKSPACE,
                "partials buffer too small for %d ranks", n_ranks);
  if (x) {
    if (int rc = to_device_exchange(x, p.rec, batch->n_slides, &p.x)) return rc;
    ACMIL_REQUIRE(n_ranks <= RT_FINISH_THREADS, ACMIL_E_INVALID, "exchange: too many ranks");
  }
  p.records = reinterpret_cast<const float*>(d_partials);
  p.n_ranks = n_ranks;
  p.n_slides = batch->n_slides;
  p.n_masked = batch->n_masked;
  p.keep_ld = keep_ld > 0 ? keep_ld : 1;
  for (int s = 0; s < batch->n_slides; ++s) {
    p.keep[s] = (keep && batch->n_masked > 0) ? keep[s] : 0;
    ACMIL_REQUIRE(p.keep[s] >= 0 && p.keep[s] <= batch->n_masked && p.keep[s] <= p.keep_ld, ACMIL_E_INVALID,
                  "keep[%d]=%d out of range", s, p.keep[s]);
    ACMIL_REQUIRE(p.keep[s] == 0 || d_rsel != nullptr || d_rand != nullptr, ACMIL_E_INVALID, "masking needs d_rsel or d_rand");
    p.row_off[s] = batch->row_offsets[s];
    p.shard_begin[s] = batch->shard_row_begin ? batch->shard_row_begin[s] : 0;
  }
  p.row_off[batch->n_slides] = batch->row_offsets[batch->n_slides];
  p.rsel = d_rsel;
  p.rand = d_rand;
  p.rand_ld = rand_ld;
  p.a_out = batch->d_a_out;
  p.a_ld = batch->a_ld;
  p.heads = *heads;
  p.out = *out;
  ACMIL_REQUIRE(!(heads->n_branch_heads > 0) || (heads->d_wc && heads->d_bc), ACMIL_E_INVALID, "branch heads need d_wc/d_bc");
  ACMIL_REQUIRE(!(heads->slide_head || heads->shared_head) || (heads->d_ws && heads->d_bs), ACMIL_E_INVALID,
                "slide/shared head needs d_ws/d_bs");
  return gp_launch_finish(p, (cudaStream_t)stream);
}

int acmil_gp_finish(const acmil_gp_shape* shape, const acmil_gp_batch* batch, const void* d_partials,
                    size_t partial_bytes, int n_ranks, const int32_t* keep, const int64_t* d_rsel, int32_t keep_ld,
                    const acmil_gp_heads* heads, const acmil_gp_outputs* out, void* stream) {
  return gp_finish_common(shape, batch, d_partials, partial_bytes, n_ranks, keep, d_rsel, keep_ld, nullptr, 0, heads, out, stream);
}

int acmil_gp_finish_x(const acmil_gp_shape* shape, const acmil_gp_batch* batch, const acmil_gp_exchange* x,
                      const int32_t* keep, const int64_t* d_rsel, const float* d_rand, int32_t rand_ld, int32_t keep_ld,
                      const acmil_gp_heads* heads, const acmil_gp_outputs* out, void* stream) {
  ACMIL_REQUIRE(x != nullptr, ACMIL_E_INVALID, "exchange is NULL");
  ACMIL_REQUIRE(!(d_rsel && d_rand), ACMIL_E_INVALID, "give d_rsel or d_rand, not both");
  return gp_finish_common(shape, batch, nullptr, 0, x->n_ranks, keep, d_rsel, keep_ld, d_rand, d_rand ? rand_ld : 0, heads, out,
                          stream, x);
}

int acmil_gp_finish_rand(const acmil_gp_shape* shape, const acmil_gp_batch* batch, const void* d_partials,
                         size_t partial_bytes, int n_ranks, const int32_t* keep, const float* d_rand, int32_t rand_ld,
                         int32_t keep_ld, const acmil_gp_heads* heads, const acmil_gp_outputs* out, void* stream) {
  ACMIL_REQUIRE(d_rand != nullptr && rand_ld >= 1, ACMIL_E_INVALID, "finish_rand needs the draws");
  return gp_finish_common(shape, batch, d_partials, partial_bytes, n_ranks, keep, nullptr, keep_ld, d_rand, rand_ld, heads, out, stream);
}

int acmil_gp_attn_stats(const float* d_a, int64_t a_ld, int32_t n_branch, const int64_t* row_offsets, int32_t n_slides,
                        const float* d_lse_m, const float* d_lse_l, float* d_gram, float* d_ent, float* d_div,
                        void* stream) {
  ACMIL_REQUIRE(d_a && row_offsets && d_lse_m && d_lse_l, ACMIL_E_INVALID, "NULL argument");
  ACMIL_REQUIRE(n_branch >= 1 && n_branch <= KMAX, ACMIL_E_INVALID, "n_branch must be in [1, %d]", KMAX);
  ACMIL_REQUIRE(n_slides >= 0 && n_slides <= SMAX, ACMIL_E_INVALID, "n_slides must be in [0, %d]", SMAX);
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  return gp_launch_stats(d_a, a_ld, n_branch, row_offsets, n_slides, d_lse_m, d_lse_l, d_gram, d_ent, d_div,
                         (cudaStream_t)stream);
}

int acmil_softmax_rows(const float* d_a, int64_t a_ld, int32_t n_rows, int64_t n, float* d_out, int64_t out_ld,
                       void* stream) {
  ACMIL_REQUIRE(d_a && d_out, ACMIL_E_INVALID, "NULL argument");
  ACMIL_REQUIRE(n_rows >= 0 && n >= 0, ACMIL_E_INVALID, "negative size");
  ACMIL_REQUIRE(acmil_device_count() > 0, ACMIL_E_CUDA, "no CUDA device: acmil_b200 has no CPU path");
  return gp_launch_softmax_rows(d_a, a_ld, n_rows, n, d_out, out_ld, (cudaStream_t)stream);
}

}  // extern "C"
