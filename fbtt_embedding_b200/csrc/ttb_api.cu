// extern "C" entry points of libttb for the TT forward/backward ops (include/ttb.h),
// shape validation and path dispatch.  Cache / hash / preprocess entry points live in
// ttb_cache.cu.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "ttb_common.cuh"

namespace ttb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_path{TTB_PATH_AUTO};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool tuning_flag(const char* name) {
  // a handful of names, each looked up in the environment the first time it is asked for
  static std::mutex mu;
  static std::vector<std::pair<std::string, bool>> seen;
  std::lock_guard<std::mutex> lk(mu);
  for (auto& kv : seen)
    if (kv.first == name) return kv.second;
  const char* v = getenv(name);
  const bool on = v && v[0] == '1' && v[1] == 0;
  seen.emplace_back(name, on);
  return on;
}
int current_path() { return g_path.load(std::memory_order_relaxed); }

namespace {
struct TimingRec {
  int kind;
  cudaEvent_t a, b;
};
std::atomic<int> g_timing{0};
std::mutex g_timing_mu;
std::vector<TimingRec> g_timing_recs;
}  // namespace

KernelTimer::KernelTimer(int kind, cudaStream_t stream) : slot_(-1), stream_(stream) {
  if (!g_timing.load(std::memory_order_relaxed)) return;
  TimingRec r;
  r.kind = kind;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, stream);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timing_recs.push_back(r);
  slot_ = (int)g_timing_recs.size() - 1;
}
KernelTimer::~KernelTimer() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (slot_ < (int)g_timing_recs.size()) cudaEventRecord(g_timing_recs[slot_].b, stream_);
}

int make_chain_dims(const ttb_shape_t* s, ChainDims* d) {
  TTB_CHECK(s != nullptr, "shape is NULL");
  TTB_CHECK(s->T >= 2 && s->T <= TTB_MAX_CORES, "T=%d not in [2,4] (tt_embeddings_ops.py:475-476)",
            s->T);
  TTB_CHECK(s->num_tables > 0, "num_tables must be > 0");
  TTB_CHECK(s->B > 0, "B must be > 0");
  TTB_CHECK(s->D > 0, "D must be > 0");                      // tt_embeddings_cuda.cu:988
  TTB_CHECK(s->D % 4 == 0, "D=%d must be divisible by 4", s->D);  // :989
  TTB_CHECK(s->R[0] == 1 && s->R[s->T] == 1, "ranks must start and end with 1");
  memset(d, 0, sizeof(*d));
  d->T = s->T;
  d->num_tables = s->num_tables;
  d->B = s->B;
  d->D = s->D;
  long long prodq = 1, Lv = 1;
  for (int t = 0; t <= s->T; ++t) {
    TTB_CHECK(s->R[t] > 0, "rank %d must be > 0", t);
    d->R[t] = s->R[t];
  }
  for (int t = s->T - 1; t >= 0; --t) {
    TTB_CHECK(s->p[t] > 0 && s->q[t] > 0, "p/q must be > 0");
    d->L[t] = Lv;
    TTB_CHECK(s->L[t] == Lv, "L[%d]=%lld is not prod(p[%d:])=%lld (tt_embeddings_ops.py:506-512)", t,
              (long long)s->L[t], t + 1, Lv);
    Lv *= s->p[t];
  }
  int mm = 1;
  d->vmax = 0;
  d->vsum = 0;
  for (int t = 0; t < s->T; ++t) {
    d->p[t] = s->p[t];
    d->q[t] = s->q[t];
    prodq *= s->q[t];
    d->S[t] = s->R[t] * s->q[t] * s->R[t + 1];
    mm *= s->q[t];
    d->m[t] = mm;
    d->n[t] = s->q[t] * s->R[t + 1];
    d->vsize[t] = mm * s->R[t + 1];
    if (d->vsize[t] > d->vmax) d->vmax = d->vsize[t];
    if (t < s->T - 1) {
      d->voff[t] = d->vsum;
      d->vsum += (d->vsize[t] + 3) & ~3;
    }
  }
  if (d->D > d->vmax) d->vmax = d->D;
  d->vmax = (d->vmax + 3) & ~3;
  TTB_CHECK(prodq == s->D, "prod(q)=%lld != D=%d", prodq, s->D);
  d->total_rows = Lv;  // prod(p) after the L loop
  d->small32 = Lv < (1LL << 31) ? 1 : 0;
  return 0;
}

// implemented in ttb_tt_generic.cu
int launch_fwd_generic(const ChainDims&, int64_t, const int64_t*, const int64_t*, const int64_t*,
                       const CorePtrs&, float*, const int32_t* mask, cudaStream_t);
int launch_bwd_generic(const ChainDims&, int64_t, const int64_t*, const int64_t*, const int64_t*,
                       const float*, const CorePtrs&, const CorePtrsRW&, const int32_t* mask, cudaStream_t);
int launch_optimizer_sweep(const ChainDims&, int, float, float, const CorePtrsRW&,
                           const CorePtrsRW&, const CorePtrsRW&, cudaStream_t, int core_mask = 0xF);
// implemented in ttb_tt_fast.cu
void set_trace(long long* fwd, long long* bwd);
bool fast_supported(const ChainDims&);
bool bf16_supported(const ChainDims&);
size_t fast_workspace_bytes(const ChainDims&, int64_t nnz);
size_t fast_workspace_header_bytes(const ChainDims&, int64_t nnz);
int launch_fwd_fast(const ChainDims&, const LookupBatch&, const CorePtrs&, float*, void*, size_t, int, cudaStream_t);
int launch_bwd_fast(const ChainDims&, const LookupBatch&, int optim, float lr, float eps, const float*,
                    const CorePtrs&, const CorePtrsRW& grads, const CorePtrsRW& state, void*, size_t, int,
                    int* sweep_mask, cudaStream_t);

static bool use_fast(const ChainDims& d, int* err) {
  *err = 0;
  const int path = current_path();
  if (path == TTB_PATH_GENERIC) return false;
  const bool ok = fast_supported(d);
  if (path == TTB_PATH_FAST && !ok) {
    set_error("TTB_PATH_FAST requested but this shape is not supported by the bucketed kernels");
    *err = 1;
  }
  return ok;
}

static LookupBatch coo_batch(const ChainDims& d, int64_t nnz, const int64_t* indices, const int64_t* rowidx,
                             const int64_t* tableidx, const int32_t* mask) {
  LookupBatch b;
  b.nnz = nnz;
  b.indices = indices;
  b.rowidx = rowidx;
  b.tableidx = tableidx;
  b.offsets = nullptr;
  b.num_bags = 0;
  b.B = d.B;
  b.mask = mask;
  b.zero_output = 0;
  b.bf16_cores = 0;
  return b;
}

// body of every forward entry point once the chain and the batch are described
static int forward_impl(const ChainDims& d, const LookupBatch& b, const float* const* cores, float* output,
                        void* workspace, size_t workspace_bytes, int plan_ready, cudaStream_t stream) {
  TTB_CHECK(b.nnz >= 0, "nnz must be >= 0");
  if (b.nnz == 0) return 0;  // tt_embeddings_cuda.cu:983-985
  const bool csr = b.offsets != nullptr && !b.rowidx && !b.tableidx;
  TTB_CHECK(b.indices && cores && output && (csr || (b.rowidx && b.tableidx)), "NULL pointer argument");
  CorePtrs c;
  for (int t = 0; t < TTB_MAX_CORES; ++t) c.c[t] = t < d.T ? cores[t] : nullptr;
  for (int t = 0; t < d.T; ++t) TTB_CHECK(c.c[t] != nullptr, "core %d is NULL", t);
  int err;
  if (use_fast(d, &err)) return launch_fwd_fast(d, b, c, output, workspace, workspace_bytes, plan_ready, stream);
  if (err) return 1;
  TTB_CHECK(!csr, "a CSR batch needs the bucketed path; run ttb_preprocess_rowidx first for this shape / path");
  if (b.zero_output)
    TTB_CUDA(cudaMemsetAsync(output, 0, (size_t)(d.het ? d.het_tables : d.num_tables) * d.B * d.D * sizeof(float), stream));
  return launch_fwd_generic(d, b.nnz, b.indices, b.rowidx, b.tableidx, c, output, b.mask, stream);
}

// body of every backward entry point
static int backward_impl(const ChainDims& d, int optim, float lr, float eps, const LookupBatch& b,
                         const float* d_output, float* const* cores, float* const* grads, float* const* opt_state,
                         void* workspace, size_t workspace_bytes, int plan_ready, cudaStream_t stream) {
  TTB_CHECK(optim == TTB_OPTIM_SGD || optim == TTB_OPTIM_ADAGRAD || optim == TTB_OPTIM_DENSE,
            "unknown optimizer %d", optim);
  TTB_CHECK(b.nnz >= 0, "nnz must be >= 0");
  if (b.nnz == 0) return 0;  // tt_embeddings_cuda.cu:448-450
  const bool csr = b.offsets != nullptr && !b.rowidx && !b.tableidx;
  TTB_CHECK(b.indices && d_output && cores && grads && (csr || (b.rowidx && b.tableidx)), "NULL pointer argument");
  CorePtrs c;
  CorePtrsRW cw, g, s;
  for (int t = 0; t < TTB_MAX_CORES; ++t) {
    c.c[t] = t < d.T ? cores[t] : nullptr;
    cw.c[t] = t < d.T ? cores[t] : nullptr;
    g.c[t] = t < d.T ? grads[t] : nullptr;
    s.c[t] = (t < d.T && optim == TTB_OPTIM_ADAGRAD && opt_state) ? opt_state[t] : nullptr;
  }
  for (int t = 0; t < d.T; ++t) {
    TTB_CHECK(c.c[t] && g.c[t], "core/grad %d is NULL", t);
    if (optim == TTB_OPTIM_ADAGRAD) TTB_CHECK(s.c[t] != nullptr, "optimizer_state %d is NULL", t);
  }
  int err;
  int sweep_mask = 0xF;  // cores the dense sweep still has to visit (the tcgen05 backward applies the optimizer itself)
  if (use_fast(d, &err)) {
    if (launch_bwd_fast(d, b, optim, lr, eps, d_output, c, g, s, workspace, workspace_bytes, plan_ready, &sweep_mask,
                        stream))
      return 1;
  } else {
    if (err) return 1;
    TTB_CHECK(!csr, "a CSR batch needs the bucketed path; run ttb_preprocess_rowidx first for this shape / path");
    if (launch_bwd_generic(d, b.nnz, b.indices, b.rowidx, b.tableidx, d_output, c, g, b.mask, stream)) return 1;
  }
  if (optim == TTB_OPTIM_DENSE || sweep_mask == 0) return 0;
  return launch_optimizer_sweep(d, optim, lr, eps, cw, g, s, stream, sweep_mask);
}

// chain of a fused heterogeneous batch: the concatenated shape (one table, P_t slices per core) plus the
// per-table radices on the device
static int make_chain_dims_het(const ttb_shape_t* cat_shape, int32_t n_tables, const ttb_het_table_t* tables_dev,
                               const ttb_row_map_t* map, ChainDims* d) {
  if (make_chain_dims(cat_shape, d)) return 1;
  TTB_CHECK(cat_shape->num_tables == 1,
            "het: cat_shape.num_tables must be 1 (the tables are concatenated along the slice dimension), got %d",
            cat_shape->num_tables);
  TTB_CHECK(n_tables > 0, "het: n_tables must be > 0");
  TTB_CHECK(tables_dev != nullptr, "het: table descriptors are NULL");
  for (int t = 0; t < d->T; ++t)
    TTB_CHECK(d->p[t] >= n_tables, "het: core %d has %d slices for %d tables", t, d->p[t], n_tables);
  d->het = tables_dev;
  d->het_tables = n_tables;
  if (map) {
    TTB_CHECK(map->world > 0 && map->rows_per_rank > 0, "row map: world and rows_per_rank must be > 0");
    TTB_CHECK((long long)map->world * map->rows_per_rank == d->B, "row map: world (%d) x rows_per_rank (%d) != B (%d)",
              map->world, map->rows_per_rank, d->B);
    TTB_CHECK(map->tables_total >= n_tables, "row map: tables_total (%d) < local tables (%d)", map->tables_total,
              n_tables);
    TTB_CHECK(map->peer_offset && map->table_gid, "row map: peer_offset / table_gid is NULL");
    d->map_peer_off = (const long long*)map->peer_offset;
    d->map_gid = map->table_gid;
    d->map_bw = map->rows_per_rank;
    d->map_tt = map->tables_total;
  }
  return 0;
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_abi_version(void) { return TTB_ABI_VERSION; }
const char* ttb_last_error(void) { return g_err; }
int ttb_set_path(int path) {
  TTB_CHECK(path >= TTB_PATH_AUTO && path <= TTB_PATH_FAST, "unknown path %d", path);
  g_path.store(path);
  return 0;
}
int ttb_get_path(void) { return current_path(); }
int64_t ttb_launch_count(void) { return g_launches.load(); }

int ttb_trace_set(int64_t* fwd, int64_t* bwd) {
  set_trace((long long*)fwd, (long long*)bwd);
  return 0;
}

int ttb_timing_enable(int on) {
  g_timing.store(on ? 1 : 0);
  return 0;
}
int ttb_timing_collect(double* ms, int64_t* counts, int n) {
  TTB_CHECK(ms && counts && n > 0 && n <= TTB_KIND_COUNT, "bad timing buffers");
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int i = 0; i < n; ++i) {
    ms[i] = 0.0;
    counts[i] = 0;
  }
  for (auto& r : g_timing_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess &&
        r.kind < n) {
      ms[r.kind] += t;
      counts[r.kind] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_timing_recs.clear();
  return 0;
}

size_t ttb_tt_workspace_bytes(const ttb_shape_t* shape, int64_t nnz) {
  ChainDims d;
  if (make_chain_dims(shape, &d)) return 0;
  if (current_path() != TTB_PATH_GENERIC && fast_supported(d)) return fast_workspace_bytes(d, nnz);
  return 0;
}

size_t ttb_tt_workspace_header_bytes(const ttb_shape_t* shape, int64_t nnz) {
  ChainDims d;
  if (make_chain_dims(shape, &d)) return 0;
  if (current_path() != TTB_PATH_GENERIC && fast_supported(d)) return fast_workspace_header_bytes(d, nnz);
  return 0;
}

int ttb_tt_forward(const ttb_shape_t* shape, int64_t nnz, const int64_t* indices,
                   const int64_t* rowidx, const int64_t* tableidx, const float* const* cores,
                   float* output, void* workspace, size_t workspace_bytes, int plan_ready,
                   cudaStream_t stream) {
  return ttb_tt_forward_masked(shape, nnz, indices, rowidx, tableidx, nullptr, cores, output, workspace,
                               workspace_bytes, plan_ready, stream);
}

int ttb_tt_forward_masked(const ttb_shape_t* shape, int64_t nnz, const int64_t* indices,
                          const int64_t* rowidx, const int64_t* tableidx,
                          const int32_t* cache_locations, const float* const* cores, float* output,
                          void* workspace, size_t workspace_bytes, int plan_ready,
                          cudaStream_t stream) {
  ChainDims d;
  if (make_chain_dims(shape, &d)) return 1;
  return forward_impl(d, coo_batch(d, nnz, indices, rowidx, tableidx, cache_locations), cores, output, workspace,
                      workspace_bytes, plan_ready, stream);
}

int ttb_tt_backward(const ttb_shape_t* shape, int optim, float lr, float eps, int64_t nnz,
                    const int64_t* indices, const int64_t* rowidx, const int64_t* tableidx,
                    const float* d_output, float* const* cores, float* const* grads,
                    float* const* opt_state, void* workspace, size_t workspace_bytes,
                    int plan_ready, cudaStream_t stream) {
  return ttb_tt_backward_masked(shape, optim, lr, eps, nnz, indices, rowidx, tableidx, nullptr, d_output,
                                cores, grads, opt_state, workspace, workspace_bytes, plan_ready, stream);
}

int ttb_tt_backward_masked(const ttb_shape_t* shape, int optim, float lr, float eps, int64_t nnz,
                           const int64_t* indices, const int64_t* rowidx, const int64_t* tableidx,
                           const int32_t* cache_locations, const float* d_output,
                           float* const* cores, float* const* grads, float* const* opt_state,
                           void* workspace, size_t workspace_bytes, int plan_ready,
                           cudaStream_t stream) {
  ChainDims d;
  if (make_chain_dims(shape, &d)) return 1;
  return backward_impl(d, optim, lr, eps, coo_batch(d, nnz, indices, rowidx, tableidx, cache_locations), d_output,
                       cores, grads, opt_state, workspace, workspace_bytes, plan_ready, stream);
}

int ttb_het_describe(int32_t T, int32_t n_tables, const int32_t* p_shapes, ttb_het_table_t* tables,
                     int32_t* P) {
  TTB_CHECK(T >= 2 && T <= TTB_MAX_CORES, "T=%d not in [2,4]", T);
  TTB_CHECK(n_tables > 0, "n_tables must be > 0");
  TTB_CHECK(p_shapes && tables && P, "NULL pointer argument");
  long long tot[TTB_MAX_CORES] = {0, 0, 0, 0};
  for (int k = 0; k < n_tables; ++k) {
    ttb_het_table_t& h = tables[k];
    memset(&h, 0, sizeof(h));
    long long Lv = 1;
    for (int t = T - 1; t >= 0; --t) {
      const int32_t pt = p_shapes[(size_t)k * T + t];
      TTB_CHECK(pt > 0, "table %d: p[%d]=%d must be > 0", k, t, pt);
      h.p[t] = pt;
      h.L[t] = Lv;
      TTB_CHECK(Lv <= (1LL << 62) / pt, "table %d: prod(p) overflows int64", k);
      Lv *= pt;
    }
    h.rows = Lv;
    for (int t = 0; t < T; ++t) {
      h.off[t] = (int32_t)tot[t];
      tot[t] += h.p[t];
      TTB_CHECK(tot[t] < (1LL << 31), "core %d: more than 2^31 concatenated slices", t);
    }
  }
  for (int t = 0; t < TTB_MAX_CORES; ++t) P[t] = t < T ? (int32_t)tot[t] : 0;
  return 0;
}

int ttb_het_digits(int32_t T, int32_t n_tables, const ttb_het_table_t* tables, int64_t table,
                   int64_t index, int32_t* digits, int32_t* valid) {
  TTB_CHECK(T >= 2 && T <= TTB_MAX_CORES, "T=%d not in [2,4]", T);
  TTB_CHECK(n_tables > 0 && tables && digits && valid, "bad arguments");
  ChainDims d;
  memset(&d, 0, sizeof(d));
  d.T = T;
  d.het = tables;
  d.het_tables = n_tables;
  int i[TTB_MAX_CORES] = {0, 0, 0, 0};
  *valid = het_digits(d, table, index, i) ? 1 : 0;
  for (int t = 0; t < T; ++t) digits[t] = *valid ? i[t] : 0;
  return 0;
}

int ttb_row_map_offset(const ttb_row_map_t* map, int32_t B, int32_t D, int64_t table, int64_t row,
                       int64_t* offset) {
  TTB_CHECK(map && offset, "NULL pointer argument");
  TTB_CHECK(map->world > 0 && map->rows_per_rank > 0 && (long long)map->world * map->rows_per_rank == B,
            "row map: world x rows_per_rank != B");
  TTB_CHECK(map->peer_offset && map->table_gid, "row map: peer_offset / table_gid is NULL");
  TTB_CHECK(row >= 0 && row < B && table >= 0, "row / table out of range");
  ChainDims d;
  memset(&d, 0, sizeof(d));
  d.B = B;
  d.D = D;
  d.map_peer_off = (const long long*)map->peer_offset;
  d.map_gid = map->table_gid;
  d.map_bw = map->rows_per_rank;
  d.map_tt = map->tables_total;
  *offset = out_row_offset(d, table, row);
  return 0;
}

int ttb_tt_forward_het(const ttb_shape_t* cat_shape, int32_t n_tables, const ttb_het_table_t* tables_dev,
                       const ttb_row_map_t* row_map, int64_t nnz, const int64_t* indices,
                       const int64_t* rowidx, const int64_t* tableidx, const float* const* cores,
                       float* output, void* workspace, size_t workspace_bytes, int plan_ready,
                       cudaStream_t stream) {
  ChainDims d;
  if (make_chain_dims_het(cat_shape, n_tables, tables_dev, row_map, &d)) return 1;
  return forward_impl(d, coo_batch(d, nnz, indices, rowidx, tableidx, nullptr), cores, output, workspace,
                      workspace_bytes, plan_ready, stream);
}

int ttb_tt_backward_het(const ttb_shape_t* cat_shape, int32_t n_tables,
                        const ttb_het_table_t* tables_dev, const ttb_row_map_t* row_map, int optim,
                        float lr, float eps, int64_t nnz, const int64_t* indices, const int64_t* rowidx,
                        const int64_t* tableidx, const float* d_output, float* const* cores,
                        float* const* grads, float* const* opt_state, void* workspace,
                        size_t workspace_bytes, int plan_ready, cudaStream_t stream) {
  ChainDims d;
  if (make_chain_dims_het(cat_shape, n_tables, tables_dev, row_map, &d)) return 1;
  return backward_impl(d, optim, lr, eps, coo_batch(d, nnz, indices, rowidx, tableidx, nullptr), d_output, cores,
                       grads, opt_state, workspace, workspace_bytes, plan_ready, stream);
}

static int batch_chain(const ttb_shape_t* shape, const ttb_batch_t* batch, ChainDims* d, LookupBatch* b) {
  TTB_CHECK(batch != nullptr, "batch is NULL");
  if (batch->n_het_tables > 0) {
    if (make_chain_dims_het(shape, batch->n_het_tables, batch->het_tables, batch->row_map, d)) return 1;
  } else {
    TTB_CHECK(batch->row_map == nullptr, "a row map needs a heterogeneous batch (n_het_tables > 0)");
    if (make_chain_dims(shape, d)) return 1;
  }
  b->nnz = batch->nnz;
  b->indices = batch->indices;
  b->rowidx = batch->rowidx;
  b->tableidx = batch->tableidx;
  b->offsets = batch->offsets;
  b->num_bags = batch->num_bags_total;
  b->B = d->B;
  b->mask = batch->cache_locations;
  b->zero_output = (batch->flags & TTB_BATCH_ZERO_OUTPUT) ? 1 : 0;
  b->bf16_cores = (batch->flags & TTB_BATCH_BF16_CORES) ? 1 : 0;
  if (b->bf16_cores)
    TTB_CHECK(current_path() != TTB_PATH_GENERIC && bf16_supported(*d),
              "bf16 cores need the tcgen05 kernel family (T = 3, q0 = 4, equal ranks 16 / 32 / 64 / 128) and a non-generic path");
  TTB_CHECK(!(b->zero_output && batch->row_map), "TTB_BATCH_ZERO_OUTPUT does not apply to a row-mapped (peer) output");
  if (batch->offsets && !batch->rowidx) {
    const long long tables = batch->n_het_tables > 0 ? batch->n_het_tables : d->num_tables;
    TTB_CHECK(batch->tableidx == nullptr, "CSR batch: pass offsets with rowidx == tableidx == NULL");
    TTB_CHECK(batch->num_bags_total == tables * d->B, "CSR batch: offsets must cover tables x B = %lld bags, got %lld",
              tables * d->B, (long long)batch->num_bags_total);
  }
  return 0;
}

int ttb_tt_forward_batch(const ttb_shape_t* shape, const ttb_batch_t* batch, const float* const* cores,
                         float* output, void* workspace, size_t workspace_bytes, int plan_ready,
                         cudaStream_t stream) {
  ChainDims d;
  LookupBatch b;
  if (batch_chain(shape, batch, &d, &b)) return 1;
  return forward_impl(d, b, cores, output, workspace, workspace_bytes, plan_ready, stream);
}

int ttb_tt_backward_batch(const ttb_shape_t* shape, const ttb_batch_t* batch, int optim, float lr, float eps,
                          const float* d_output, float* const* cores, float* const* grads,
                          float* const* opt_state, void* workspace, size_t workspace_bytes, int plan_ready,
                          cudaStream_t stream) {
  ChainDims d;
  LookupBatch b;
  if (batch_chain(shape, batch, &d, &b)) return 1;
  return backward_impl(d, optim, lr, eps, b, d_output, cores, grads, opt_state, workspace, workspace_bytes,
                       plan_ready, stream);
}

int ttb_optimizer_step(const ttb_shape_t* shape, int optim, float lr, float eps,
                       float* const* cores, float* const* grads, float* const* opt_state,
                       cudaStream_t stream) {
  ChainDims d;
  if (make_chain_dims(shape, &d)) return 1;
  TTB_CHECK(optim == TTB_OPTIM_SGD || optim == TTB_OPTIM_ADAGRAD,
            "ttb_optimizer_step: optimizer must be SGD or ADAGRAD, got %d", optim);
  TTB_CHECK(cores && grads, "NULL pointer argument");
  CorePtrsRW cw, g, s;
  for (int t = 0; t < TTB_MAX_CORES; ++t) {
    cw.c[t] = t < d.T ? cores[t] : nullptr;
    g.c[t] = t < d.T ? grads[t] : nullptr;
    s.c[t] = (t < d.T && optim == TTB_OPTIM_ADAGRAD && opt_state) ? opt_state[t] : nullptr;
  }
  for (int t = 0; t < d.T; ++t) {
    TTB_CHECK(cw.c[t] && g.c[t], "core/grad %d is NULL", t);
    if (optim == TTB_OPTIM_ADAGRAD) TTB_CHECK(s.c[t] != nullptr, "optimizer_state %d is NULL", t);
  }
  return launch_optimizer_sweep(d, optim, lr, eps, cw, g, s, stream);
}

}  // extern "C"
