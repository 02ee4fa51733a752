/*
 * ttb.h -- C ABI of libttb.so, the B200-native (sm_100a) TT-EmbeddingBag hot path.
 *
 * This is the drop-in boundary for the reference's `tt_embeddings` extension
 * (facebookresearch/FBTT-Embedding, tt_embeddings.cpp:131-161: eleven pybind
 * functions).  Every entry point below replaces exactly one of them and cites it.
 * No ATen / torch types cross this ABI: raw device pointers, sizes, scalars and a
 * cudaStream_t.  All outputs and scratch are allocated by the caller (the Python
 * shim uses torch's caching allocator, preserving the reference's stream semantics,
 * tt_embeddings_cuda.cu:54-55).  All work is enqueued on `stream`; nothing blocks
 * the host except ttb_preprocess_cached (the reference blocks there too,
 * tt_embeddings_cuda.cu:1481-1488).
 *
 * Return value: 0 on success, non-zero on error; ttb_last_error() returns a
 * thread-local message (the shim raises RuntimeError, matching TORCH_CHECK).
 *
 * Data layout (identical to the reference, tt_embeddings_ops.py:506-530):
 *   core t   : float [num_tables][p_t][r_t * q_t * r_{t+1}], one slice is a row-major
 *              r_t x (q_t*r_{t+1}) matrix
 *   L[t]     : prod_{s>t} p_s ; digits i_t = idx / L[t], idx %= L[t]
 *   output   : float [num_tables][B][D]        d_output: same, contiguous
 *   indices / rowidx / tableidx : int64 [nnz]  (COO of the CSR offsets)
 */
#ifndef TTB_H_
#define TTB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define TTB_MAX_CORES 4
#define TTB_ABI_VERSION 8

/* POD shape descriptor.  R has T+1 entries, R[0] == R[T] == 1. */
typedef struct ttb_shape {
  int32_t T;          /* number of TT cores, 2..4 (tt_embeddings_ops.py:475-476) */
  int32_t num_tables; /* leading dim of every core */
  int32_t B;          /* bags per table */
  int32_t D;          /* embedding dim == prod(q), must be % 4 == 0 (tt_embeddings_cuda.cu:989) */
  int32_t p[TTB_MAX_CORES];
  int32_t q[TTB_MAX_CORES];
  int32_t R[TTB_MAX_CORES + 1];
  int64_t L[TTB_MAX_CORES];
} ttb_shape_t;

/* optimizer selector of ttb_tt_backward (tt_embeddings_cuda.cu:31-35) */
enum { TTB_OPTIM_SGD = 0, TTB_OPTIM_ADAGRAD = 1, TTB_OPTIM_DENSE = 2 };

/* compute path selector (process-wide, see ttb_set_path) */
enum {
  TTB_PATH_AUTO = 0,    /* fast kernels when the shape qualifies, else generic */
  TTB_PATH_GENERIC = 1, /* fp32 FFMA kernels, any T/p/q/r  (reference-test exact, rtol 1.3e-6) */
  TTB_PATH_FAST = 2     /* force the bucketed tensor-core kernels; error if shape unsupported */
};

int ttb_abi_version(void);
const char* ttb_last_error(void);
int ttb_set_path(int path);
int ttb_get_path(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t ttb_launch_count(void);

/* Per-kernel-class device timing for bench.py's roofline pass: while enabled, every launch is
 * bracketed by CUDA events on its stream; ttb_timing_collect synchronises them, adds the
 * elapsed milliseconds and launch counts per class into ms[] / counts[] (n entries, n <=
 * TTB_KIND_COUNT), and clears the record.  Off by default (no events, no overhead). */
enum { TTB_KIND_FWD = 0, TTB_KIND_BWD = 1, TTB_KIND_SWEEP = 2, TTB_KIND_PLAN = 3, TTB_KIND_CACHE = 4,
       TTB_KIND_COUNT = 5 };
int ttb_timing_enable(int on);
/* Debug phase trace of the tcgen05 kernels: device buffers of 16 int64 per CTA (>= 1024 CTAs) that thread 0 of every
 * CTA stamps with %globaltimer at its phase boundaries (slot 0 = CTA start, 15 = CTA end); NULL switches it off.
 * scripts/trace_phases.py prints the per-phase medians. */
int ttb_trace_set(int64_t* fwd, int64_t* bwd);
int ttb_timing_collect(double* ms, int64_t* counts, int n);

/* ---- tt_forward  (replaces tt_embeddings_forward_cuda, tt_embeddings.cpp:13-26,
 *      tt_embeddings_cuda.cu:964-1075).  `output` must be zero-filled by the caller
 *      (the reference does at::zeros, :981-982).  nnz == 0 is a no-op.
 *      `workspace` (ttb_tt_workspace_bytes() bytes; may be NULL when that is 0) holds the
 *      bucketing plan of the tensor-core path: a permutation of the lookups grouped by
 *      (table, middle-core index) plus the tile list.  With plan_ready == 0 the call builds the
 *      plan there; with plan_ready != 0 it trusts a plan built by an earlier call for the SAME
 *      (shape, nnz, indices, tableidx) -- the backward of a step reuses the forward's plan. */
int ttb_tt_forward(const ttb_shape_t* shape, int64_t nnz, const int64_t* indices,
                   const int64_t* rowidx, const int64_t* tableidx,
                   const float* const* cores, float* output, void* workspace,
                   size_t workspace_bytes, int plan_ready, cudaStream_t stream);

/* ---- tt_dense_backward / tt_sgd_backward / tt_adagrad_backward (replace
 *      tt_embeddings_backward_{dense,sgd,adagrad}_cuda, tt_embeddings.cpp:28-72,
 *      tt_embeddings_cuda.cu:419-752).
 *      grads[t]: core-shaped fp32, zero on entry.  TTB_OPTIM_DENSE leaves the batch
 *      gradient there (the op's return value).  SGD / ADAGRAD use them as scratch,
 *      apply the update to cores[t] (and opt_state[t]) over EVERY row (the reference
 *      launch skips rows when p_t > S_t, SURVEY Q1 -- not replicated) and re-zero
 *      them before returning so the caller can reuse the buffers without a memset. */
int ttb_tt_backward(const ttb_shape_t* shape, int optim, float lr, float eps, int64_t nnz,
                    const int64_t* indices, const int64_t* rowidx, const int64_t* tableidx,
                    const float* d_output, float* const* cores, float* const* grads,
                    float* const* opt_state, void* workspace, size_t workspace_bytes,
                    int plan_ready, cudaStream_t stream);

/* ---- the optimizer half of the fused backward on its own: applies SGD / Adagrad
 *      (tt_embeddings_cuda.cu:392, 412-414) to cores[t] (and opt_state[t]) from the dense,
 *      core-shaped gradients grads[t] and re-zeroes them.  For data-parallel replicas
 *      (SURVEY 8f-3): ttb_tt_backward(TTB_OPTIM_DENSE) per rank, all-reduce of grads over
 *      the ranks, then this call -- every replica applies the same summed gradient.
 *      Elements whose gradient is exactly 0 are left untouched (like the fused path). */
int ttb_optimizer_step(const ttb_shape_t* shape, int optim, float lr, float eps,
                       float* const* cores, float* const* grads, float* const* opt_state,
                       cudaStream_t stream);

/* scratch needed by ttb_tt_forward / ttb_tt_backward for this shape and nnz.  Its first
 * ttb_tt_workspace_header_bytes() bytes (bucket counters + sync words of the plan kernels) must
 * be ZERO whenever a call builds a plan (plan_ready == 0); the kernels leave them zero, so a
 * buffer that was zero-filled once can be reused for any number of plans without a memset. */
size_t ttb_tt_workspace_bytes(const ttb_shape_t* shape, int64_t nnz);
size_t ttb_tt_workspace_header_bytes(const ttb_shape_t* shape, int64_t nnz);

/* ---- table groups (a rank's heterogeneous tables in ONE call).  The reference's
 *      TableBatchedTTEmbeddingBag needs identical shapes (tt_embeddings_ops.py:424, README.md:136-141),
 *      so a DLRM with 26 differently-sized tables pays 26 x (3 pybind calls + autograd node) of host
 *      time per step (BASELINE config 4 was host-launch bound: DESIGN.md section 8).  A group is an
 *      array of independent items -- each one exactly the argument set of ttb_preprocess_rowidx /
 *      ttb_tt_forward / ttb_tt_backward for one table family -- walked by the library: one host
 *      call per phase for the whole group, same kernels, same numerics as the per-table calls.
 *      Items must not alias each other's outputs, gradients or workspaces.  With
 *      ttb_group_set_streams(k > 1) the items are spread round-robin over k internal streams that
 *      fork from and join back into `stream` (event record / wait, capturable in a CUDA graph), so
 *      small tables overlap on the GPU; k == 1 (default) enqueues everything on `stream`. */
typedef struct ttb_group_item {
  ttb_shape_t shape;            /* this table family; every item of a group may differ */
  int64_t nnz;                  /* lookups of this item (0: the item is skipped) */
  const int64_t* indices;       /* int64[nnz] */
  const int64_t* offsets;       /* int64[num_tables*B + 1], CSR; read by ttb_group_preprocess only */
  int64_t* rowidx;              /* int64[nnz]: written by preprocess, read by forward / backward */
  int64_t* tableidx;            /* int64[nnz]: same */
  float* cores[TTB_MAX_CORES];
  float* grads[TTB_MAX_CORES];      /* backward: core-shaped, zero on entry (ttb_tt_backward contract) */
  float* opt_state[TTB_MAX_CORES];  /* backward, TTB_OPTIM_ADAGRAD only */
  float* output;                /* forward: [num_tables][B][D], zero-filled by the caller */
  const float* d_output;        /* backward: [num_tables][B][D] */
  void* workspace;              /* plan buffer of this item (ttb_tt_workspace_bytes), header zero */
  size_t workspace_bytes;
  int32_t plan_ready;           /* as in ttb_tt_forward / ttb_tt_backward */
  int32_t reserved;
} ttb_group_item_t;

int ttb_group_set_streams(int k);  /* 1..16 */
int ttb_group_get_streams(void);
int ttb_group_preprocess(int n_items, const ttb_group_item_t* items, cudaStream_t stream);
int ttb_group_forward(int n_items, const ttb_group_item_t* items, cudaStream_t stream);
int ttb_group_backward(int n_items, const ttb_group_item_t* items, int optim, float lr, float eps,
                       cudaStream_t stream);

/* ---- fused heterogeneous table batch (second step: ONE plan / forward / backward launch
 *      for tables of DIFFERENT sizes; the optimizer is applied inside the backward launch).  The reference
 *      batches tables only when their TT shapes are identical (tt_embeddings_ops.py:424: one
 *      [num_tables, p_t, S_t] tensor per core).  Tables that share the
 *      q-shapes and ranks (every table of a DLRM does: same D, same rank setting) differ only in their
 *      p-shapes, i.e. in HOW MANY slices each core has -- so their cores can be concatenated along the slice
 *      dimension: core t = float [1][P_t][S_t] with P_t = sum over tables of p_t(table), table k owning slices
 *      [off_t(k), off_t(k) + p_t(k)).  To every kernel that walks core slices this is one table with P_t
 *      slices; only the index decomposition differs: lookup n of table k = tableidx[n] has digits
 *      i_t = off_t(k) + (idx / L_t(k)) % p_t(k).  ttb_het_describe fills the per-table descriptors (host) and
 *      the concatenated slice counts from the tables' p-shapes; the caller copies the descriptors to device
 *      memory once and passes that pointer.  `cat_shape` is an ordinary ttb_shape_t with num_tables == 1,
 *      p[t] = P_t, L = prod(P[t+1:]) (what ttb_het_describe returns in P), B = bags per table; it is also
 *      the shape for ttb_tt_workspace_bytes / _header_bytes / ttb_optimizer_step.  output / d_output are
 *      [n_tables][B][D] and rowidx / tableidx come from ttb_preprocess_rowidx with num_bags_total =
 *      n_tables * B, exactly as for identical tables.  A lookup whose tableidx is outside [0, n_tables) or
 *      whose index is outside its table's [0, rows) contributes nothing. */
typedef struct ttb_het_table {
  int64_t rows;                /* prod(p): valid indices of this table are [0, rows) */
  int64_t L[TTB_MAX_CORES];    /* prod(p[t+1:]) of THIS table */
  int32_t p[TTB_MAX_CORES];
  int32_t off[TTB_MAX_CORES];  /* first slice of this table in concatenated core t */
} ttb_het_table_t;

/* p_shapes: int32 [n_tables][T] (row-major).  Writes tables[n_tables] (host) and P[T]. */
int ttb_het_describe(int32_t T, int32_t n_tables, const int32_t* p_shapes, ttb_het_table_t* tables,
                     int32_t* P);
/* Host evaluation of the index decomposition the het kernels perform (the same inline function, compiled for
 * the host): digits[t] = concatenated slice number of `index` of table `table` in core t, *valid = 0 when the
 * table or the index is out of range (such a lookup contributes nothing).  `tables` is a HOST array here. */
int ttb_het_digits(int32_t T, int32_t n_tables, const ttb_het_table_t* tables, int64_t table,
                   int64_t index, int32_t* digits, int32_t* valid);

/* ---- fused exchange (the table-parallel all-to-all of pooled rows, folded into the kernels).
 *      Table-parallel sharding gives rank s the WHOLE batch for its tables and wants, on rank w, the batch
 *      slice [w*bw, (w+1)*bw) of ALL tables: X_w = float [bw][tables_total][D].  With a row map the forward
 *      does not write output[table][row][:] but adds the pooled row straight into
 *          X_w[row % bw][table_gid[table]][:],   w = row / bw,
 *      where X_w lives in the memory of rank w (peer-mapped over NVLink: one unified address space), and the
 *      backward reads d_output from the same place of rank w's gradient buffer -- so there is no all-to-all
 *      kernel and no staging copy; the transfer overlaps the math tile by tile.  Addresses are formed as
 *      `output + peer_offset[w] + ...` with peer_offset[w] = (X_w - X_self) in ELEMENTS (both 16-byte aligned):
 *      plain 64-bit pointer arithmetic, which is why only the plan / index stage knows about the map and the
 *      tile kernels are the ordinary ones.  For the backward to reuse the forward's plan, d_output must sit at
 *      the same distance from the peers' gradient buffers: allocate X and dX as two regions of ONE symmetric
 *      buffer per rank.  The caller zero-fills every X_w and synchronises the ranks (stream-ordered barrier)
 *      before the forward, and synchronises again before anyone reads X_w / after everyone wrote dX_w.
 *      row_map == NULL: ordinary [n_tables][B][D] output. */
typedef struct ttb_row_map {
  int32_t world;               /* ranks the batch is split over; cat_shape.B must equal world * rows_per_rank */
  int32_t rows_per_rank;       /* bw */
  int32_t tables_total;        /* tables of ALL ranks: the middle dimension of X_w */
  int32_t reserved;
  const int64_t* peer_offset;  /* DEVICE int64[world]: (X_w - X_self) in floats; peer_offset[self] == 0 */
  const int32_t* table_gid;    /* DEVICE int32[n_tables]: global number of each local table */
} ttb_row_map_t;
/* host evaluation of the same address function (host arrays in `map`): element offset of (table, row)'s
 * pooled row relative to `output` / `d_output` */
int ttb_row_map_offset(const ttb_row_map_t* map, int32_t B, int32_t D, int64_t table, int64_t row,
                       int64_t* offset);

int ttb_tt_forward_het(const ttb_shape_t* cat_shape, int32_t n_tables, const ttb_het_table_t* tables_dev,
                       const ttb_row_map_t* row_map, int64_t nnz, const int64_t* indices,
                       const int64_t* rowidx, const int64_t* tableidx, const float* const* cores,
                       float* output, void* workspace, size_t workspace_bytes, int plan_ready,
                       cudaStream_t stream);
int ttb_tt_backward_het(const ttb_shape_t* cat_shape, int32_t n_tables,
                        const ttb_het_table_t* tables_dev, const ttb_row_map_t* row_map, int optim,
                        float lr, float eps, int64_t nnz, const int64_t* indices, const int64_t* rowidx,
                        const int64_t* tableidx, const float* d_output, float* const* cores,
                        float* const* grads, float* const* opt_state, void* workspace,
                        size_t workspace_bytes, int plan_ready, cudaStream_t stream);

/* ---- one batch descriptor for every lookup flavour (ABI v8).  COO (rowidx / tableidx, the reference's op
 *      interface, tt_embeddings.cpp:13-72) or CSR: rowidx == tableidx == NULL and `offsets` over
 *      tables x B bags (+ the end offset) -- the CSR -> COO step of compute_rowidx_kernel
 *      (tt_embeddings_cuda.cu:1338-1354) then happens inside the plan kernel of the bucketed path, so a training
 *      step through the modules is plan + forward + backward(with optimizer) = 3 launches.  A CSR batch on a shape
 *      / path the bucketed kernels do not cover is an error (run ttb_preprocess_rowidx and pass COO).
 *      cache_locations: optional mask of the async cache front-end.  n_het_tables > 0: fused heterogeneous batch
 *      (`shape` is the concatenated shape, het_tables the DEVICE descriptors, row_map optional). */
typedef struct ttb_batch {
  int64_t nnz;
  const int64_t* indices;
  const int64_t* rowidx;
  const int64_t* tableidx;
  const int64_t* offsets;
  int64_t num_bags_total;
  const int32_t* cache_locations;
  int32_t n_het_tables;
  int32_t flags; /* TTB_BATCH_* */
  const ttb_het_table_t* het_tables;
  const ttb_row_map_t* row_map;
} ttb_batch_t;
/* forward: `output` is uninitialised memory; the library zero-fills it -- inside the plan kernel when the call builds
 * a plan (one launch less than a memset in front of the call) */
#define TTB_BATCH_ZERO_OUTPUT 1
/* cores[t] (passed through the `float*` parameters) hold bf16 values, same shapes: BASELINE configs[2], "bf16 cores /
 * fp32 accumulate".  The tcgen05 kernels stage them as they are (one-term products, fp32 accumulation in TMEM; the
 * last link and all gradients in fp32); the fused optimizers update them as w = bf16_rn(float(w) - step), the
 * Adagrad state stays fp32; TTB_OPTIM_DENSE returns fp32 gradients.  Shapes outside the tcgen05 family: error. */
#define TTB_BATCH_BF16_CORES 2

int ttb_tt_forward_batch(const ttb_shape_t* shape, const ttb_batch_t* batch, const float* const* cores,
                         float* output, void* workspace, size_t workspace_bytes, int plan_ready,
                         cudaStream_t stream);
int ttb_tt_backward_batch(const ttb_shape_t* shape, const ttb_batch_t* batch, int optim, float lr, float eps,
                          const float* d_output, float* const* cores, float* const* grads,
                          float* const* opt_state, void* workspace, size_t workspace_bytes, int plan_ready,
                          cudaStream_t stream);

/* ---- update_cache_state (replaces update_cache_state_cuda, tt_embeddings.cpp:74,
 *      tt_embeddings_cuda.cu:1077-1113; hashtbl_insert hashtbl_cuda_utils.cuh:102-133) */
int ttb_update_cache_state(int64_t nnz, const int64_t* indices, int64_t hashtbl_size,
                           int64_t* hashtbl, int64_t* cache_freq, cudaStream_t stream);

/* ---- cache_populate (replaces cache_populate_cuda, tt_embeddings.cpp:76-86,
 *      tt_embeddings_cuda.cu:1260-1336).  sorted_keys / sorted_freq: int64[hashtbl_size]
 *      scratch; temp: ttb_cache_populate_temp_bytes(hashtbl_size) bytes. */
size_t ttb_cache_populate_temp_bytes(int64_t hashtbl_size);
int ttb_cache_populate(const ttb_shape_t* shape, const float* const* cores,
                       int64_t hashtbl_size, int64_t* hashtbl, int64_t* cache_freq,
                       int32_t* cache_state, int64_t cache_size, float* cache_weight,
                       int64_t* sorted_keys, int64_t* sorted_freq, void* temp,
                       size_t temp_bytes, cudaStream_t stream);

/* ---- preprocess_indices_sync (replaces preprocess_indices_sync_cuda,
 *      tt_embeddings.cpp:88-95, tt_embeddings_cuda.cu:1377-1496), split in its two
 *      branches.  ttb_preprocess_rowidx is the warm-up / multi-table branch (CSR->COO).
 *      ttb_preprocess_cached additionally looks every index up in the LFU table and
 *      stably partitions (colidx,rowidx,cache_locations): TT lookups first in order,
 *      cached lookups at the tail in REVERSE order (cub::DevicePartition::Flagged
 *      semantics, :1436-1479); writes the TT count to *h_num_tt (pinned or pageable
 *      host memory) and synchronises `stream` before returning, like the reference.
 *      tile_scratch: int32[ttb_preprocess_tile_count(nnz)+1]. */
int ttb_preprocess_rowidx(int64_t nnz, int64_t num_bags_total, int32_t B,
                          const int64_t* offsets, int64_t* rowidx, int64_t* tableidx,
                          cudaStream_t stream);
int64_t ttb_preprocess_tile_count(int64_t nnz);
int ttb_preprocess_cached(int64_t nnz, const int64_t* colidx, const int64_t* rowidx,
                          int64_t hashtbl_size, const int64_t* hashtbl,
                          const int32_t* cache_state, int64_t* out_colidx,
                          int64_t* out_rowidx, int32_t* out_cache_locations,
                          int32_t* tile_scratch, int32_t* h_num_tt, cudaStream_t stream);

/* ---- async cache front-end (no counterpart among the reference's ops).  In steady state
 *      the reference spends, per step, update_cache_state + preprocess_indices_sync = 7 launches, a
 *      D2H copy of the TT count and a stream synchronisation (tt_embeddings_cuda.cu:1481-1488) before it
 *      can launch the lookup.  ttb_cache_frontend does the same bookkeeping in ONE launch and leaves the
 *      batch in order: hashtbl / cache_freq are updated exactly as by ttb_update_cache_state, rowidx /
 *      tableidx as by ttb_preprocess_rowidx, and cache_locations[n] is the cache row of lookup n or -1
 *      (what the reference's cache_lookup_kernel returns, :1356-1375) -- no partition, no count, no
 *      host round trip, so a whole cached training step can sit in a CUDA graph.
 *      ttb_tt_forward_masked / ttb_tt_backward_masked are ttb_tt_forward / ttb_tt_backward over the
 *      lookups with cache_locations[n] == -1 (NULL: all of them; a plan built with a mask must be
 *      reused with the same mask); ttb_cache_forward and the three ttb_cache_backward_* skip entries
 *      with a negative location, so both halves take the same full-length arrays.  An entry the
 *      offsets do not cover gets -2 and is skipped by both. */
int ttb_cache_frontend(int64_t nnz, const int64_t* colidx, int64_t num_bags_total, int32_t B,
                       const int64_t* offsets, int64_t hashtbl_size, int64_t* hashtbl,
                       int64_t* cache_freq, const int32_t* cache_state, int64_t* rowidx,
                       int64_t* tableidx, int32_t* cache_locations, cudaStream_t stream);
int ttb_tt_forward_masked(const ttb_shape_t* shape, int64_t nnz, const int64_t* indices,
                          const int64_t* rowidx, const int64_t* tableidx,
                          const int32_t* cache_locations, const float* const* cores, float* output,
                          void* workspace, size_t workspace_bytes, int plan_ready,
                          cudaStream_t stream);
int ttb_tt_backward_masked(const ttb_shape_t* shape, int optim, float lr, float eps, int64_t nnz,
                           const int64_t* indices, const int64_t* rowidx, const int64_t* tableidx,
                           const int32_t* cache_locations, const float* d_output,
                           float* const* cores, float* const* grads, float* const* opt_state,
                           void* workspace, size_t workspace_bytes, int plan_ready,
                           cudaStream_t stream);

/* ---- cache_forward (replaces cache_forward_cuda, tt_embeddings.cpp:97-103,
 *      tt_embeddings_cuda.cu:1498-1572): output[row] += cache_weight[loc] */
int ttb_cache_forward(int32_t B, int64_t nnz, int32_t D, const int32_t* cache_locations,
                      const int64_t* rowidx, const float* cache_weight, float* output,
                      cudaStream_t stream);

/* ---- cache_backward_sgd / _dense / _rowwise_adagrad_approx (replace
 *      tt_embeddings.cpp:105-129, tt_embeddings_cuda.cu:1574-1835) */
int ttb_cache_backward_sgd(int64_t nnz, int32_t D, const float* grad_output,
                           const int32_t* cache_locations, const int64_t* rowidx, float lr,
                           float* cache_weight, cudaStream_t stream);
int ttb_cache_backward_dense(int64_t nnz, int32_t D, const float* grad_output,
                             const int32_t* cache_locations, const int64_t* rowidx,
                             float* grad_cache_weight /* zero on entry */, cudaStream_t stream);
int ttb_cache_backward_rowwise_adagrad_approx(int64_t nnz, int32_t D, const float* grad_output,
                                              const int32_t* cache_locations,
                                              const int64_t* rowidx, float lr, float eps,
                                              float* cache_optimizer_state,
                                              float* cache_weight, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TTB_H_ */
