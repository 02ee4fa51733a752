// Bucketed tensor-core TT-EmbeddingBag path for sm_100a (T == 3): plan kernels, launchers, and the legacy tf32 kernels.
//
// Idea: one lookup's GEMMs are tiny (M = q0 = 4), but the lookups of a
// batch that share the middle-core index i1 also share the whole B operand core1[i1]
// (r1 x q1*r2).  So the batch is bucketed by (table, i1); inside a bucket the q0-row A tiles
// core0[i0] of 32 lookups are stacked to M = 128 and ONE tcgen05.mma (fp32 accumulate in TMEM)
// computes tr0 for all 32 lookups.  The tiny last link (K = r2, N = q2, per-lookup operand
// core2[i2]) and the bag pooling run in the epilogue straight out of TMEM (tcgen05.ld) on the
// FFMA pipe.  The backward does the same for the gradient GEMMs: the per-lookup 16 KB atomic
// scatter of dCore1 (reference K6/K7) becomes one tensor-core GEMM per tile whose K dimension
// runs over the lookups of the bucket.
//
//   plan     : (CSR -> COO) + histogram by (table,i1) -> exclusive scan + tile / run lists ->
//              scatter of packed per-lookup records {i0, i2, output-row offset} in bucket order;
//              ONE launch for small batches (last-arriving CTA scans), three for large ones
//   forward  : ttb_tt_x.cuh x_fwd_kernel (kind::f16 on bf16 hi/lo split operands, fp32-grade);
//              the round-1 kind::tf32 kernels below stay selectable with TTB_LEGACY_TC=1
//   backward : ttb_tt_x.cuh x_bwd_kernel: three MMAs per tile (recompute, dCore0 rows, dCore1),
//              SIMT stage for G and dCore2, optimizer applied in the same launch
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ttb_common.cuh"
#include "ttb_sm100.cuh"

namespace ttb {

using namespace sm100;

// The single-launch plan kernel spins (bounded, then traps) until the last of its CTAs has arrived, so ALL of its
// CTAs must be co-resident: it is launched cooperatively with at most one CTA per SM.  A table group running on k
// lanes (ttb_group.cu) may have k of them in flight at once; each then gets 1/k of the CTA budget (larger plans take
// the three-launch path, which never spins).
static thread_local int g_onepass_share = 1;
// debug phase trace (ttb_trace_set): device buffers of 16 int64 per CTA, or nullptr
static long long* g_trace_fwd = nullptr;
static long long* g_trace_bwd = nullptr;
static long long* g_trace_plan = nullptr;  // the forward buffer's second half (CTAs 1024..)
void set_trace(long long* fwd, long long* bwd) {
  g_trace_fwd = fwd;
  g_trace_bwd = bwd;
  g_trace_plan = fwd ? fwd + 1024 * 16 : nullptr;
}
void set_onepass_share(int k) { g_onepass_share = k < 1 ? 1 : k; }

namespace {

constexpr int kTileLookups = 32;   // 32 lookups x q0(=4) rows = one M=128 MMA tile == one work item
constexpr int kFastThreads = 256;  // 8 warps: warp w and w+4 share TMEM lane quarter w%4

struct __align__(16) LookupRec {
  int i0, i2;
  long long orow;  // element offset of the bag's row in output / d_output: (table*B + row)*D
};

struct PlanView {
  int* counts;        // [nb]   lookups per bucket        } header: must be ZERO when a plan is
  int* bucket_done;   // [nb]   landed items of split buckets (backward)  } built; the kernels
  int* sync_words;    // [8]    tickets / flags           } leave it zero
  size_t header_bytes;
  int* bucket_start;  // [nb+1]
  int* cursor;        // [nb]
  int* num_tiles;     // [4]: {tiles, runs, max_run_tiles, -}
  int* tile_bucket;   // [max_tiles]
  int* tile_begin;    // [max_tiles]  offset into recs
  int* tile_count;    // [max_tiles]
  int* run_bucket;    // [max_tiles]  runs: up to max_run_tiles consecutive tiles of ONE bucket
  int* run_begin;     // [max_tiles]
  int* run_count;     // [max_tiles]  lookups in the run
  LookupRec* recs;    // [nnz]  lookups grouped by bucket
  int nb, max_tiles;
  size_t bytes;
};

// sync_words: [0] plan arrival ticket, [1] plan "scan published" flag, [2] plan departure ticket,
//             [3] backward CTAs that have finished their items, [4] tail sweepers that have finished
constexpr int kSyncWords = 8;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

PlanView carve_plan(const ChainDims& d, int64_t nnz, void* ws) {
  PlanView p;
  p.nb = d.num_tables * d.p[1];
  p.max_tiles = p.nb + (int)(nnz / kTileLookups) + 1;
  char* base = (char*)ws;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = base + off;
    off += align_up(bytes, 256);
    return r;
  };
  p.counts = (int*)take((size_t)p.nb * 4);
  p.bucket_done = (int*)take((size_t)p.nb * 4);
  p.sync_words = (int*)take(kSyncWords * 4);
  p.header_bytes = off;
  p.cursor = (int*)take((size_t)p.nb * 4);
  p.bucket_start = (int*)take((size_t)(p.nb + 1) * 4);
  p.num_tiles = (int*)take(16);
  p.tile_bucket = (int*)take((size_t)p.max_tiles * 4);
  p.tile_begin = (int*)take((size_t)p.max_tiles * 4);
  p.tile_count = (int*)take((size_t)p.max_tiles * 4);
  p.run_bucket = (int*)take((size_t)p.max_tiles * 4);
  p.run_begin = (int*)take((size_t)p.max_tiles * 4);
  p.run_count = (int*)take((size_t)p.max_tiles * 4);
  p.recs = (LookupRec*)take((size_t)nnz * sizeof(LookupRec));
  p.bytes = off;
  return p;
}

// Runs: a bucket with up to max_run tiles is ONE run (one CTA sees every lookup of the bucket, so the complete
// dCore1 slice sits in its TMEM accumulator and the optimizer is applied from there); larger buckets are cut into
// runs of max_run tiles for load balance (their partial slices meet in the gradient scratch).  max_run grows with
// the batch so that the number of runs stays near 1.5x the CTAs a kernel can keep resident.
inline int plan_max_run(const ChainDims& d, int64_t nnz, int nb) {
  const long long est_tiles = nnz / kTileLookups + nb / 2 + 1;
  const long long slots = (long long)sm_count() * 2;
  long long m = (3 * est_tiles + 2 * slots - 1) / (2 * slots);
  if (m < 4) m = 4;
  if (m > 64) m = 64;
  return (int)m;
}

// ---- plan kernels ---------------------------------------------------------------------------
// digits of one index; 32-bit arithmetic when the table has < 2^31 rows (a 64-bit division costs
// ~10x a 32-bit one and every lookup needs two)
__device__ __forceinline__ bool digits3(const ChainDims& d, long long tb, long long idx, int& i0, int& i1, int& i2) {
  if (d.het) {  // heterogeneous batch: table tb's own radices, slice numbers in the concatenated cores
    int i[TTB_MAX_CORES];
    if (!het_digits(d, tb, idx, i)) return false;
    i0 = i[0];
    i1 = i[1];
    i2 = i[2];
    return true;
  }
  if (idx < 0 || idx >= d.total_rows) return false;
  if (d.small32) {
    const unsigned u = (unsigned)idx, L0 = (unsigned)d.L[0], L1 = (unsigned)d.L[1];
    const unsigned a = u / L0;
    const unsigned rem = u - a * L0;
    const unsigned b = rem / L1;
    i0 = (int)a;
    i1 = (int)b;
    i2 = (int)(rem - b * L1);
  } else {
    const long long a = idx / d.L[0];
    const long long rem = idx - a * d.L[0];
    const long long b = rem / d.L[1];
    i0 = (int)a;
    i1 = (int)b;
    i2 = (int)(rem - b * d.L[1]);
  }
  return true;
}

__device__ __forceinline__ int bucket_of(const ChainDims& d, long long idx, long long tb) {
  int i0, i1, i2;
  if (!digits3(d, tb, idx, i0, i1, i2)) return -1;
  return d.het ? i1 : (int)(tb * d.p[1] + i1);  // het: i1 already counts slices across all tables
}

__device__ __forceinline__ void write_rec(const ChainDims& d, LookupRec* recs, int pos, long long idx,
                                          long long tb, long long row) {
  int i0 = 0, i1 = 0, i2 = 0;
  digits3(d, tb, idx, i0, i1, i2);
  LookupRec r;
  r.i0 = i0;
  r.i2 = i2;
  r.orow = out_row_offset(d, tb, row);  // local [table][row][:] or a peer's batch-slice buffer (fused exchange)
  recs[pos] = r;
}

// what a plan is built from: COO triples (reference op interface) or CSR offsets (rowidx == nullptr: the
// CSR -> COO step of compute_rowidx_kernel, tt_embeddings_cuda.cu:1338-1354, happens here, per lookup)
struct PlanIn {
  const long long* indices;
  const long long* rowidx;    // COO: bag row of lookup n (nullptr with offsets: derived; nullptr without: row n)
  const long long* tableidx;  // COO: table of lookup n (nullptr: table 0 / derived from offsets)
  const long long* offsets;   // CSR: [num_bags + 1], bag b = table b / B, row b % B
  long long num_bags;
  int B;
  const int* mask;            // optional: only mask[n] == -1 is a TT lookup (async cache front-end)
  long long nnz;
  float4* zero_ptr;           // optional: the forward's output, zero-filled by the plan kernel (one launch less
  long long zero_n4;          // than a memset in front of it)
  float bags_per_lookup;      // num_bags / nnz: the proportional guess of the bag search
};

// the forward's output is accumulated into with red.add: it has to start from zero
__device__ __forceinline__ void plan_zero_fill(const PlanIn& in, long long first, long long stride) {
  if (!in.zero_ptr) return;
  for (long long i = first; i < in.zero_n4; i += stride) in.zero_ptr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// optional: the plan kernel pulls the core slices its lookups will touch towards L2 (fire-and-forget prefetches), so
// that the forward's dependent gathers hit L2 instead of HBM when the step starts cold
struct PlanPrefetch {
  const char* core[3];
  int slice_bytes[3];
  int on;
};

struct PlanOut {
  int* counts;
  int* cursor;
  int* sync_words;
  int* bucket_start;
  LookupRec* recs;
  int* tile_bucket;
  int* tile_begin;
  int* tile_count;
  int* run_bucket;
  int* run_begin;
  int* run_count;
  int* num_tiles;
  int nb, max_run;
  long long* trace;  // debug phase trace (ttb_trace_set) or nullptr
};

__device__ __forceinline__ void plan_stamp(const PlanOut& o, int slot) {
  if (o.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    o.trace[(size_t)blockIdx.x * 16 + slot] = (long long)t;
  }
}

// lookup n of the batch -> (index, table, bag row); false when it is not a TT lookup of this batch
__device__ __forceinline__ bool plan_resolve(const PlanIn& in, long long n, long long& idx, long long& tb,
                                             long long& row) {
  if (in.mask && __ldg(in.mask + n) != -1) return false;  // served by the LFU cache
  idx = __ldg(in.indices + n);
  if (in.rowidx || !in.offsets) {
    tb = in.tableidx ? __ldg(in.tableidx + n) : 0;
    row = in.rowidx ? __ldg(in.rowidx + n) : n;
    return true;
  }
  // every load of the common case is issued before the first one is looked at: the range bounds, the proportional
  // guess and its successor are ONE round trip to memory instead of two
  // (single precision: FP64 is a slow pipe on this GPU, and a guess that is off by one costs one more probe)
  long long g = (long long)((float)n * in.bags_per_lookup);
  g = g < 0 ? 0 : (g > in.num_bags - 1 ? in.num_bags - 1 : g);
  const long long first = __ldg(in.offsets), last = __ldg(in.offsets + in.num_bags);
  const long long og = __ldg(in.offsets + g), og1 = __ldg(in.offsets + g + 1);
  if (n < first || n >= last) return false;  // not covered by any bag
  const long long bag = bag_of_probe(in.offsets, in.num_bags, n, g, og, og1);
  tb = bag / in.B;
  row = bag - tb * in.B;
  return true;
}

// per-bucket part of the scan: publishes the bucket's offsets, its 32-lookup tiles and its runs
__device__ __forceinline__ void plan_emit_bucket(const PlanOut& o, int b, int v, int cbase, int& sbase, int& rbase) {
  o.bucket_start[b] = cbase;
  o.cursor[b] = cbase;
  for (int off = 0; off < v; off += kTileLookups) {
    o.tile_bucket[sbase] = b;
    o.tile_begin[sbase] = cbase + off;
    o.tile_count[sbase] = min(kTileLookups, v - off);
    ++sbase;
  }
  const int run_lookups = o.max_run * kTileLookups;
  for (int off = 0; off < v; off += run_lookups) {
    o.run_bucket[rbase] = b;
    o.run_begin[rbase] = cbase + off;
    o.run_count[rbase] = min(run_lookups, v - off);
    ++rbase;
  }
}
__device__ __forceinline__ int tiles_of(int v) { return (v + kTileLookups - 1) / kTileLookups; }
__device__ __forceinline__ int runs_of(int v, int max_run) {
  const int rl = max_run * kTileLookups;
  return (v + rl - 1) / rl;
}

__global__ void __launch_bounds__(256)
    plan_hist_kernel(const ChainDims d, const PlanIn in, int* __restrict__ counts) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  plan_zero_fill(in, n, (long long)gridDim.x * blockDim.x);
  if (n >= in.nnz) return;
  long long idx, tb, row;
  if (!plan_resolve(in, n, idx, tb, row)) return;
  const int b = bucket_of(d, idx, tb);
  if (b >= 0) atomicAdd(counts + b, 1);
}

// one CTA: exclusive scan of bucket counts, then the tile and run lists
__global__ void __launch_bounds__(1024) plan_scan_kernel(const PlanOut o) {
  __shared__ int s_cnt[1024], s_seg[1024], s_run[1024];
  const int tid = threadIdx.x, nb = o.nb;
  const int per = (nb + 1023) / 1024;
  const int lo = min(nb, tid * per), hi = min(nb, lo + per);
  int c = 0, s = 0, r = 0;
  for (int b = lo; b < hi; ++b) {
    const int v = o.counts[b];
    c += v;
    s += tiles_of(v);
    r += runs_of(v, o.max_run);
  }
  s_cnt[tid] = c;
  s_seg[tid] = s;
  s_run[tid] = r;
  __syncthreads();
  for (int st = 1; st < 1024; st <<= 1) {  // Hillis-Steele inclusive scan over 1024 partials
    int a = 0, b2 = 0, c2 = 0;
    if (tid >= st) {
      a = s_cnt[tid - st];
      b2 = s_seg[tid - st];
      c2 = s_run[tid - st];
    }
    __syncthreads();
    s_cnt[tid] += a;
    s_seg[tid] += b2;
    s_run[tid] += c2;
    __syncthreads();
  }
  int cbase = s_cnt[tid] - c, sbase = s_seg[tid] - s, rbase = s_run[tid] - r;
  for (int b = lo; b < hi; ++b) {
    const int v = o.counts[b];
    o.counts[b] = 0;  // header contract: zero on entry, zero on exit
    plan_emit_bucket(o, b, v, cbase, sbase, rbase);
    cbase += v;
  }
  if (tid == 1023) {
    o.bucket_start[nb] = s_cnt[1023];
    o.num_tiles[0] = s_seg[1023];
    o.num_tiles[1] = s_run[1023];
    o.num_tiles[2] = o.max_run;
  }
}

__global__ void __launch_bounds__(256)
    plan_scatter_kernel(const ChainDims d, const PlanIn in, int* __restrict__ cursor, LookupRec* __restrict__ recs) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= in.nnz) return;
  long long idx, tb, row;
  if (!plan_resolve(in, n, idx, tb, row)) return;
  const int b = bucket_of(d, idx, tb);
  if (b >= 0) write_rec(d, recs, atomicAdd(cursor + b, 1), idx, tb, row);
}

// Small batches: the whole plan in ONE launch.  Shared-memory atomics are too slow for a
// redundant per-CTA histogram (2 cycles per lane: 40 us for 10^4 lookups), global (L2) atomics
// are not, so: every CTA histograms its 256 lookups with L2 atomics, takes a ticket; the LAST
// CTA to arrive scans the counts, publishes bucket cursors + the tile / run lists and raises a flag;
// the others wait for the flag, then scatter their lookups.  The last CTA to finish re-zeroes the sync
// words; `counts` is re-zeroed by the scanner -- the plan buffer's header is zero on entry and on exit.
// The wait needs every CTA of the grid resident at once: the launch is a COOPERATIVE launch (the driver
// refuses it instead of deadlocking when the grid does not fit next to whatever else holds SM slots), its
// size is checked against the occupancy calculator, and the wait itself is bounded (traps, like mbar_wait).
constexpr int kOnePassMaxNnz = 131072;
constexpr int kOnePassMaxBuckets = 8192;
constexpr int kOnePassThreads = 256;

__device__ __forceinline__ void plan_prefetch(const ChainDims& d, const PlanPrefetch& pf, long long n, long long tb,
                                              int i0, int i1, int i2) {
  if (!pf.on) return;
  const long long t = d.het ? 0 : tb;
  const char* c0 = pf.core[0] + ((size_t)t * d.p[0] + i0) * pf.slice_bytes[0];
  const char* c2 = pf.core[2] + ((size_t)t * d.p[2] + i2) * pf.slice_bytes[2];
  for (int b = 0; b < pf.slice_bytes[0]; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c0 + b));
  for (int b = 0; b < pf.slice_bytes[2]; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c2 + b));
  // the (large) core-1 slice is shared by the bucket: every lookup takes a different 1/32 of it
  const int lines = pf.slice_bytes[1] / 128, per = (lines + 31) / 32;
  const char* c1 = pf.core[1] + ((size_t)t * d.p[1] + i1) * pf.slice_bytes[1];
  for (int k = 0; k < per; ++k) {
    const int line = (int)(n & 31) * per + k;
    if (line < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(c1 + (size_t)line * 128));
  }
}

__global__ void __launch_bounds__(kOnePassThreads)
    plan_onepass_kernel(const ChainDims d, const PlanIn in, const PlanOut o, const PlanPrefetch pf) {
  __shared__ int s_wc[8], s_ws[8], s_wr[8];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n = (long long)blockIdx.x * kOnePassThreads + tid;
  const int nb = o.nb;
  plan_stamp(o, 0);
  plan_zero_fill(in, n, (long long)gridDim.x * kOnePassThreads);
  long long idx = 0, tb = 0, my_row = 0;
  int my_bucket = -1, my_rank = 0, i0 = 0, i1 = 0, i2 = 0;
  if (n < in.nnz && plan_resolve(in, n, idx, tb, my_row) && digits3(d, tb, idx, i0, i1, i2)) {  // digits: ONCE
    my_bucket = d.het ? i1 : (int)(tb * d.p[1] + i1);
    my_rank = atomicAdd(o.counts + my_bucket, 1);  // the histogram atomic already hands out the rank in the bucket
    plan_prefetch(d, pf, n, tb, i0, i1, i2);
  }
  plan_stamp(o, 1);  // indices resolved, histogram atomics issued
  fence_gpu();  // acq_rel at gpu scope is all the ticket protocol needs (__threadfence() is the dearer fence.sc:
  __syncthreads();  // 21 % of this kernel's stall samples in profiles/r2/readme_step_ncu_summary.txt)
  plan_stamp(o, 2);  // fence done
  if (tid == 0) s_last = (atomicAdd(o.sync_words + 0, 1) == (int)gridDim.x - 1);
  __syncthreads();
  plan_stamp(o, 3);  // ticket taken
  if (s_last) {
    fence_gpu();
    const int per = (nb + kOnePassThreads - 1) / kOnePassThreads;
    const int lo = min(nb, tid * per), hi = min(nb, lo + per);
    int c = 0, s = 0, r = 0;
    for (int b = lo; b < hi; ++b) {
      const int v = __ldcg(o.counts + b);
      c += v;
      s += tiles_of(v);
      r += runs_of(v, o.max_run);
    }
    int ic = c, is = s, ir = r;
    for (int st = 1; st < 32; st <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, ic, st), b2 = __shfl_up_sync(0xffffffffu, is, st),
                c2 = __shfl_up_sync(0xffffffffu, ir, st);
      if (lane >= st) {
        ic += a;
        is += b2;
        ir += c2;
      }
    }
    if (lane == 31) {
      s_wc[warp] = ic;
      s_ws[warp] = is;
      s_wr[warp] = ir;
    }
    __syncthreads();
    int wc_off = 0, ws_off = 0, wr_off = 0, wc_tot = 0, ws_tot = 0, wr_tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) {
        wc_off += s_wc[w];
        ws_off += s_ws[w];
        wr_off += s_wr[w];
      }
      wc_tot += s_wc[w];
      ws_tot += s_ws[w];
      wr_tot += s_wr[w];
    }
    int cbase = ic - c + wc_off, sbase = is - s + ws_off, rbase = ir - r + wr_off;
    for (int b = lo; b < hi; ++b) {
      const int v = __ldcg(o.counts + b);
      o.counts[b] = 0;  // leave the header clean for the next plan built in this buffer
      plan_emit_bucket(o, b, v, cbase, sbase, rbase);
      cbase += v;
    }
    if (tid == 0) {
      o.bucket_start[nb] = wc_tot;
      o.num_tiles[0] = ws_tot;
      o.num_tiles[1] = wr_tot;
      o.num_tiles[2] = o.max_run;
    }
    fence_gpu();
    __syncthreads();
    plan_stamp(o, 4);  // scan published (scanner CTA only)
    if (tid == 0) atomicExch(o.sync_words + 1, 1);
  }
  if (tid == 0) {
    unsigned spins = 0;
    while (atomicAdd(o.sync_words + 1, 0) == 0) {
      __nanosleep(64);
      if (++spins > (1u << 24)) __trap();  // > 1 s: the scanner CTA never ran -- fail loudly instead of hanging
    }
  }
  __syncthreads();
  plan_stamp(o, 5);  // flag seen
  fence_gpu();
  if (my_bucket >= 0) {
    LookupRec r;
    r.i0 = i0;
    r.i2 = i2;
    r.orow = out_row_offset(d, tb, my_row);
    o.recs[__ldcg(o.bucket_start + my_bucket) + my_rank] = r;
  }
  __syncthreads();
  plan_stamp(o, 6);  // records written
  if (tid == 0) {
    if (atomicAdd(o.sync_words + 2, 1) == (int)gridDim.x - 1) {
      o.sync_words[0] = 0;
      o.sync_words[1] = 0;
      o.sync_words[2] = 0;
    }
  }
}

// Small batches, an experiment (opt-in, see build_plan): the whole plan inside ONE thread-block cluster (8 CTAs x 1024 threads, nnz <= 32768,
// <= 16384 buckets).  The histogram lives in DISTRIBUTED SHARED MEMORY -- CTA c owns the counters of buckets
// [c*nbc, (c+1)*nbc) -- so a lookup is ONE remote shared-memory atomicAdd whose return value is already its rank
// inside the bucket; the scan is a block scan per CTA plus 8 published totals; positions are rank + bucket start read
// back from the owner's shared memory.  No global atomics, no memory fences, no flag to wait on: four hardware
// cluster barriers and one round trip to HBM for the indices (the single-launch kernel above pays ~8 dependent L2
// round trips: 13 us at the README shape against the forward's 10 us).
constexpr int kClusterCtas = 8;
constexpr int kClusterThreads = 1024;
constexpr int kClusterPerThread = 4;
constexpr int kClusterMaxNnz = kClusterCtas * kClusterThreads * kClusterPerThread;
constexpr int kClusterBucketsPerCta = 2048;
constexpr int kClusterMaxBuckets = kClusterCtas * kClusterBucketsPerCta;

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads)
    plan_cluster_kernel(const ChainDims d, const PlanIn in, const PlanOut o) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int s_cnt[kClusterBucketsPerCta];  // my buckets: lookup counts, later their first record
  __shared__ int s_tot[3][kClusterCtas];        // per CTA: lookups, tiles, runs (every CTA holds a full copy)
  __shared__ int s_part[3][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = (int)cluster.block_rank();
  const int nb = o.nb, nbc = (nb + kClusterCtas - 1) / kClusterCtas;
  for (int i = tid; i < nbc; i += kClusterThreads) s_cnt[i] = 0;
  plan_zero_fill(in, (long long)c * kClusterThreads + tid, (long long)kClusterCtas * kClusterThreads);
  cluster.sync();
  // ---- phase 1: one remote shared-memory atomic per lookup; its return value is the rank inside the bucket
  long long idx[kClusterPerThread], tb[kClusterPerThread], row[kClusterPerThread];
  int bucket[kClusterPerThread], rank[kClusterPerThread];
#pragma unroll
  for (int k = 0; k < kClusterPerThread; ++k) {
    const long long n = (long long)k * (kClusterCtas * kClusterThreads) + c * kClusterThreads + tid;
    bucket[k] = -1;
    rank[k] = 0;
    idx[k] = tb[k] = row[k] = 0;
    if (n < in.nnz && plan_resolve(in, n, idx[k], tb[k], row[k])) bucket[k] = bucket_of(d, idx[k], tb[k]);
  }
#pragma unroll
  for (int k = 0; k < kClusterPerThread; ++k) {
    if (bucket[k] >= 0) {
      const int owner = bucket[k] / nbc;
      int* remote = cluster.map_shared_rank(s_cnt, owner);
      rank[k] = atomicAdd(remote + (bucket[k] - owner * nbc), 1);
    }
  }
  cluster.sync();
  // ---- phase 2: exclusive scan over my buckets (lookups, tiles, runs), totals published to every CTA
  const int per = (nbc + kClusterThreads - 1) / kClusterThreads;
  const int lo = min(nbc, tid * per), hi = min(nbc, lo + per);
  int cs = 0, ts = 0, rs = 0;
  for (int i = lo; i < hi; ++i) {
    const int v = (c * nbc + i < nb) ? s_cnt[i] : 0;
    cs += v;
    ts += tiles_of(v);
    rs += runs_of(v, o.max_run);
  }
  int ic = cs, it = ts, ir = rs;
#pragma unroll
  for (int st = 1; st < 32; st <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, ic, st), b2 = __shfl_up_sync(0xffffffffu, it, st),
              c2 = __shfl_up_sync(0xffffffffu, ir, st);
    if (lane >= st) {
      ic += a;
      it += b2;
      ir += c2;
    }
  }
  if (lane == 31) {
    s_part[0][warp] = ic;
    s_part[1][warp] = it;
    s_part[2][warp] = ir;
  }
  __syncthreads();
  int off_c = 0, off_t = 0, off_r = 0, tot_c = 0, tot_t = 0, tot_r = 0;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    if (w < warp) {
      off_c += s_part[0][w];
      off_t += s_part[1][w];
      off_r += s_part[2][w];
    }
    tot_c += s_part[0][w];
    tot_t += s_part[1][w];
    tot_r += s_part[2][w];
  }
  if (tid < kClusterCtas) {  // thread j tells CTA j about my totals
    int* remote = cluster.map_shared_rank(&s_tot[0][0], tid);
    remote[0 * kClusterCtas + c] = tot_c;
    remote[1 * kClusterCtas + c] = tot_t;
    remote[2 * kClusterCtas + c] = tot_r;
  }
  cluster.sync();
  int cbase = ic - cs + off_c, sbase = it - ts + off_t, rbase = ir - rs + off_r;
  int all_c = 0, all_t = 0, all_r = 0;
#pragma unroll
  for (int j = 0; j < kClusterCtas; ++j) {
    if (j < c) {
      cbase += s_tot[0][j];
      sbase += s_tot[1][j];
      rbase += s_tot[2][j];
    }
    all_c += s_tot[0][j];
    all_t += s_tot[1][j];
    all_r += s_tot[2][j];
  }
  for (int i = lo; i < hi; ++i) {
    const int b = c * nbc + i;
    if (b >= nb) break;
    const int v = s_cnt[i];
    s_cnt[i] = cbase;  // phase 3 reads the bucket's first record from here
    plan_emit_bucket(o, b, v, cbase, sbase, rbase);
    cbase += v;
  }
  if (c == 0 && tid == 0) {
    o.bucket_start[nb] = all_c;
    o.num_tiles[0] = all_t;
    o.num_tiles[1] = all_r;
    o.num_tiles[2] = o.max_run;
  }
  cluster.sync();
  // ---- phase 3: record position = first record of the bucket (owner's shared memory) + rank
#pragma unroll
  for (int k = 0; k < kClusterPerThread; ++k) {
    if (bucket[k] >= 0) {
      const int owner = bucket[k] / nbc;
      const int* remote = cluster.map_shared_rank(s_cnt, owner);
      write_rec(d, o.recs, remote[bucket[k] - owner * nbc] + rank[k], idx[k], tb[k], row[k]);
    }
  }
  cluster.sync();  // nobody leaves while a peer may still read its shared memory
}

// <<<grid, block>>> as a cooperative launch: all CTAs co-resident or cudaErrorCooperativeLaunchTooLarge
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cooperative(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                      cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// CTAs of `kernel` that can be resident on the current device at once.  The kernel is first told to prefer the
// largest shared-memory carveout (the occupancy calculator -- and the cooperative-launch check that uses it --
// otherwise assumes the default L1 / shared split, under which a 98 KB CTA is alone on its SM); `tmem_cols` bounds
// the CTAs by their TMEM allocation (512 columns per SM), which the calculator does not know about.
template <typename K>
inline int resident_ctas(K kernel, int threads, size_t smem, int tmem_cols = 0) {
  (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int per_sm = 0;
  const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
  if (e != cudaSuccess) per_sm = 0;
  (void)cudaGetLastError();
  {
    // The calculator answers 1 for every tcgen05 kernel of this library although ncu shows several of their CTAs
    // resident per SM (round 1: 2.3 on average for the forward).  None of the callers NEEDS co-residency (no kernel
    // sized by this function waits on another CTA), so take the resource arithmetic instead: registers, shared
    // memory (+1 KB per CTA reserved by the driver), threads.
    cudaFuncAttributes fa;
    memset(&fa, 0, sizeof(fa));
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess && fa.numRegs > 0) {
      const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * threads;
      int manual = std::min(65536 / std::max(1, regs_per_cta), 2048 / std::max(1, threads));
      manual = std::min(manual, (int)(233472 / (smem + fa.sharedSizeBytes + 1024)));
      manual = std::min(manual, 32);
      if (manual > per_sm) per_sm = manual;
    }
    (void)cudaGetLastError();
  }
  if (tuning_flag("TTB_DEBUG")) {
    cudaFuncAttributes fa;
    memset(&fa, 0, sizeof(fa));
    const cudaError_t e2 = cudaFuncGetAttributes(&fa, kernel);
    fprintf(stderr, "[ttb] occupancy: threads=%d smem=%zu -> per_sm=%d (%s); regs=%d static_smem=%zu max_dyn=%d carveout=%d (%s)\n",
            threads, smem, per_sm, cudaGetErrorString(e), fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes,
            fa.preferredShmemCarveout, cudaGetErrorString(e2));
  }
  if (tmem_cols > 0 && per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
  return per_sm * sm_count();
}

int build_plan(const ChainDims& d, const PlanIn& in, const PlanView& p, cudaStream_t stream,
               const PlanPrefetch* prefetch = nullptr) {
  KernelTimer timer(TTB_KIND_PLAN, stream);
  PlanOut o;
  o.counts = p.counts;
  o.cursor = p.cursor;
  o.sync_words = p.sync_words;
  o.bucket_start = p.bucket_start;
  o.recs = p.recs;
  o.tile_bucket = p.tile_bucket;
  o.tile_begin = p.tile_begin;
  o.tile_count = p.tile_count;
  o.run_bucket = p.run_bucket;
  o.run_begin = p.run_begin;
  o.run_count = p.run_count;
  o.num_tiles = p.num_tiles;
  o.nb = p.nb;
  o.max_run = plan_max_run(d, in.nnz, p.nb);
  o.trace = g_trace_plan;
  const long long nnz = in.nnz;
  // opt-in (TTB_CLUSTER_PLAN=1): measured SLOWER than the ticket-and-flag kernel below at the README shape -- 23.1 us
  // against 16.2 us (CUDA events, eager): ~46 same-address remote shared-memory atomics per bucket serialise at the
  // SM-to-SM round trip (~215 cycles each).  Kept as the documented experiment; parity-tested (tests pass with it on).
  static const bool use_cluster = tuning_flag("TTB_CLUSTER_PLAN");
  if (use_cluster && nnz <= kClusterMaxNnz && p.nb <= kClusterMaxBuckets) {
    plan_cluster_kernel<<<kClusterCtas, kClusterThreads, 0, stream>>>(d, in, o);
    TTB_LAUNCH_CHECK();
    return 0;
  }
  if (nnz <= kOnePassMaxNnz / g_onepass_share && p.nb <= kOnePassMaxBuckets) {
    static int cap[16] = {0};
    int& c = cap[current_device() & 15];
    if (c == 0) c = resident_ctas(plan_onepass_kernel, kOnePassThreads, 0);
    const unsigned ctas = (unsigned)((nnz + kOnePassThreads - 1) / kOnePassThreads);
    // one CTA per SM at most: resident under ANY occupancy assumption (and inside stream capture, where a refused
    // cooperative launch would invalidate the capture instead of returning an error we could fall back from)
    if ((long long)ctas * g_onepass_share <= std::min(c, sm_count())) {
      PlanPrefetch pf;
      memset(&pf, 0, sizeof(pf));
      if (prefetch) pf = *prefetch;
      const cudaError_t e =
          launch_cooperative(plan_onepass_kernel, dim3(ctas), dim3(kOnePassThreads), 0, stream, d, in, o, pf);
      if (e == cudaSuccess) {
        TTB_LAUNCH_CHECK();
        return 0;
      }
      (void)cudaGetLastError();  // refused (SM slots held elsewhere): the three-launch path never waits
    }
  }
  const unsigned blocks = (unsigned)((nnz + 255) / 256);
  plan_hist_kernel<<<blocks, 256, 0, stream>>>(d, in, p.counts);
  TTB_LAUNCH_CHECK();
  plan_scan_kernel<<<1, 1024, 0, stream>>>(o);
  TTB_LAUNCH_CHECK();
  plan_scatter_kernel<<<blocks, 256, 0, stream>>>(d, in, p.cursor, p.recs);
  TTB_LAUNCH_CHECK();
  return 0;
}

// ---- shape family handled by the tensor-core kernels ----------------------------------------
// T == 3, q0 == 4 (32 lookups per 128-row tile), r1 <= 32 (K padded to 32 with zeros),
// N1 = q1*r2 == 128 (four 32-wide column blocks), r2 == 32, q1 == 4, q2 in {4, 8}.
//
// Every tensor-core operand is K-major with the 128-byte swizzle (validated layout, see
// tests/cuda/mma_probe.cu: MN-major tf32 needs a different swizzle atom, so wherever a GEMM
// needs the transpose of a staged tile, the transpose is staged too).
//   sA   [128 rows (l,j0)][32 r]     A of MMA-1
//   sAT  [32 r][128 row-index]       B of MMA-2                      (backward only)
//   sB1T [128 n=(j1,k)][32 r]        B of MMA-1
//   sB1  [32 r][128 n]               B of MMA-3                      (backward only)
//   sG   [128 rows][128 n]           A of MMA-3  (G = dTr0)          (backward only)
//   sGT  [128 n][128 row-index]      A of MMA-2                      (backward only)
constexpr int N1 = 128, R2 = 32, Q1 = 4;
constexpr int kC2StrideBase = 4;  // pad floats per lookup (bank spread)

struct TileMeta {
  LookupRec rec[kTileLookups];
};

// core1[tb][i1] (r1 x 128) -> sB1T [n][r] (always) and sB1 [r][n] (backward), tf32-rounded.
// One 4x4 block per thread: four coalesced 128-bit loads, 128-bit shared stores in both layouts
// (the transposed rows are written in a lane-rotated order so a quarter warp hits 8 distinct
// swizzle chunks -> no bank conflicts).
template <bool WITH_NATURAL>
__device__ __forceinline__ void stage_core1(const float* __restrict__ c1, int r1, uint8_t* sB1T,
                                            uint8_t* sB1, int tid) {
  const int rg = tid >> 5, lane = tid & 31;  // 8 row groups x 32 column groups == 256 threads
  const int r0 = rg * 4, n0 = lane * 4;
  float x[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + i < r1) v = to_tf32(__ldg(reinterpret_cast<const float4*>(c1 + (size_t)(r0 + i) * N1 + n0)));
    x[i][0] = v.x;
    x[i][1] = v.y;
    x[i][2] = v.z;
    x[i][3] = v.w;
    if (WITH_NATURAL) *reinterpret_cast<float4*>(sB1 + sw128_offset(32, r0 + i, n0)) = v;
  }
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = (jj + (lane >> 1)) & 3;
    float4 t;
    t.x = (j == 0) ? x[0][0] : (j == 1) ? x[0][1] : (j == 2) ? x[0][2] : x[0][3];
    t.y = (j == 0) ? x[1][0] : (j == 1) ? x[1][1] : (j == 2) ? x[1][2] : x[1][3];
    t.z = (j == 0) ? x[2][0] : (j == 1) ? x[2][1] : (j == 2) ? x[2][2] : x[2][3];
    t.w = (j == 0) ? x[3][0] : (j == 1) ? x[3][1] : (j == 2) ? x[3][2] : x[3][3];
    *reinterpret_cast<float4*>(sB1T + sw128_offset(128, n0 + j, r0)) = t;
  }
}

// A rows of the tile: row = l*4 + j0 <- core0[tb][i0_l][j0][0..r1)
template <bool WITH_TRANSPOSE, int THREADS = kFastThreads>
__device__ __forceinline__ void gather_core0(const ChainDims& d, const float* __restrict__ core0, int tb,
                                             const TileMeta* m, int nl, uint8_t* sA, uint8_t* sAT, int tid) {
  const int r1 = d.R[1];
  for (int it = tid; it < 128 * 8; it += THREADS) {
    const int row = it >> 3, ch = it & 7;
    const int l = row >> 2, j0 = row & 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < nl && ch * 4 < r1) {
      const float* src = core0 + ((size_t)tb * d.p[0] + m->rec[l].i0) * d.S[0] + j0 * r1 + ch * 4;
      v = to_tf32(__ldg(reinterpret_cast<const float4*>(src)));
    }
    *reinterpret_cast<float4*>(sA + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
    if (WITH_TRANSPOSE) {
      *reinterpret_cast<float*>(sAT + sw128_offset(32, ch * 4 + 0, row)) = v.x;
      *reinterpret_cast<float*>(sAT + sw128_offset(32, ch * 4 + 1, row)) = v.y;
      *reinterpret_cast<float*>(sAT + sw128_offset(32, ch * 4 + 2, row)) = v.z;
      *reinterpret_cast<float*>(sAT + sw128_offset(32, ch * 4 + 3, row)) = v.w;
    }
  }
}

// per-lookup last-core slices core2[tb][i2_l] (R2 x Q2 fp32 = 512 B / 1 KB, contiguous in HBM) ->
// sC2[l][..] (padded stride) with the TMA bulk-copy engine: warp 0 arms the mbarrier with the tile's
// byte count and lane l issues the copy of lookup l; consumers wait on the mbarrier only when they
// first need core2, so the copies overlap the core0 gather and the MMA.
template <int Q2>
__device__ __forceinline__ void tma_core2(const ChainDims& d, const float* __restrict__ core2, int tb,
                                          const TileMeta* m, int nl, float* sC2, uint64_t* mbar, int tid) {
  constexpr int kStride = R2 * Q2 + kC2StrideBase;
  constexpr uint32_t kBytes = R2 * Q2 * 4;
  if (tid < 32) {
    if (tid == 0) mbar_arrive_expect_tx(mbar, (uint32_t)nl * kBytes);
    __syncwarp();
    if (tid < nl)
      tma_bulk_g2s(sC2 + tid * kStride, core2 + ((size_t)tb * d.p[2] + m->rec[tid].i2) * d.S[2], kBytes, mbar);
  }
}

__device__ __forceinline__ void load_tile_meta(TileMeta* m, int tid, int nl, const LookupRec* __restrict__ recs) {
  if (tid < kTileLookups) {
    LookupRec r;
    r.i0 = 0;
    r.i2 = 0;
    r.orow = 0;
    if (tid < nl) r = recs[tid];
    m->rec[tid] = r;
  }
}

// tr0[128 x 128] = A[128 x 32] * B1[32 x 128]  (4 K-steps of 8), issued by one thread
__device__ __forceinline__ void issue_mma1(uint32_t d_tmem, const uint8_t* sA, const uint8_t* sB1T) {
  constexpr uint32_t kIdesc = make_idesc_tf32(128, N1, 0, 0);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t adesc = make_desc_sw128(smem_u32(sA) + ks * 32, 16, 1024);
    const uint64_t bdesc = make_desc_sw128(smem_u32(sB1T) + ks * 32, 16, 1024);
    mma_tf32(d_tmem, adesc, bdesc, kIdesc, ks > 0);
  }
}

template <int Q2>
struct FwdSmem {
  static constexpr int kAStage = 128 * 128;     // 16 KB
  static constexpr int kBStage = 128 * 128;     // sB1T: 128 rows x 32 tf32 = 16 KB
  static constexpr int kC2Stride = R2 * Q2 + kC2StrideBase;
  static constexpr int kC2Stage = kTileLookups * kC2Stride * 4;
  static constexpr int kMeta = 1024;
  static constexpr int kBytes = 1024 /*align slack*/ + kAStage + kBStage + kC2Stage + kMeta;
};

template <int Q2>
__global__ void __launch_bounds__(kFastThreads)
    tt_fwd_tc_kernel(const ChainDims d, const LookupRec* __restrict__ recs,
                     const int* __restrict__ tile_bucket, const int* __restrict__ tile_begin,
                     const int* __restrict__ tile_count, const int* __restrict__ num_tiles,
                     const CorePtrs cores, float* __restrict__ out, const int pdl) {
  using SM = FwdSmem<Q2>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB1T = sA + SM::kAStage;
  float* sC2 = (float*)(sB1T + SM::kBStage);
  uint8_t* metab = (uint8_t*)sC2 + SM::kC2Stage;
  uint64_t* mbar = (uint64_t*)metab;
  uint64_t* mbarC = (uint64_t*)(metab + 8);  // core2 slices landed (TMA)
  uint32_t* tmem_slot = (uint32_t*)(metab + 16);
  TileMeta* meta = (TileMeta*)(metab + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger(pdl);  // the backward (or whatever follows) may start its own prologue
  pdl_wait(pdl);     // the plan (previous kernel in the stream) is complete and visible from here on
  const int ntiles = *num_tiles;
  if ((int)blockIdx.x >= ntiles) return;  // whole CTA exits before touching TMEM
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbarC, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < SM::kC2Stage / 4; i += kFastThreads) sC2[i] = 0.f;  // padding rows stay finite
  fence_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int bucket = tile_bucket[tile];
    const int tb = bucket / d.p[1];
    const int i1 = bucket - tb * d.p[1];
    const int nl = tile_count[tile];
    load_tile_meta(meta, tid, nl, recs + tile_begin[tile]);
    stage_core1<false>(cores.c[1] + ((size_t)tb * d.p[1] + i1) * d.S[1], d.R[1], sB1T, nullptr, tid);
    __syncthreads();
    tma_core2<Q2>(d, cores.c[2], tb, meta, nl, sC2, mbarC, tid);
    gather_core0<false>(d, cores.c[0], tb, meta, nl, sA, nullptr, tid);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
      issue_mma1(tmem_base, sA, sB1T);
      mma_commit(mbar);
    }
    mbar_wait(mbarC, phase);
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after_sync();
    // ---- epilogue: row (l, j0) x column half -> last link (K = r2) + pooling
    {
      const int row = (warp & 3) * 32 + lane;
      const int half = warp >> 2;
      const int l = row >> 2, j0 = row & 3;
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const float* c2 = sC2 + l * SM::kC2Stride;
#pragma unroll
      for (int jj = 0; jj < Q1 / 2; ++jj) {
        const int j1 = half * (Q1 / 2) + jj;
        float acc[Q2];
#pragma unroll
        for (int j2 = 0; j2 < Q2; ++j2) acc[j2] = 0.f;
#pragma unroll
        for (int kc = 0; kc < R2; kc += 16) {
          float v[16];
          tmem_ld16(taddr + j1 * R2 + kc, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) {
#pragma unroll
            for (int j4 = 0; j4 < Q2; j4 += 4) {
              const float4 w = *reinterpret_cast<const float4*>(c2 + (kc + k) * Q2 + j4);
              acc[j4 + 0] = fmaf(v[k], w.x, acc[j4 + 0]);
              acc[j4 + 1] = fmaf(v[k], w.y, acc[j4 + 1]);
              acc[j4 + 2] = fmaf(v[k], w.z, acc[j4 + 2]);
              acc[j4 + 3] = fmaf(v[k], w.w, acc[j4 + 3]);
            }
          }
        }
        if (l < nl) {
          float* dst = out + meta->rec[l].orow + (j0 * Q1 + j1) * Q2;
#pragma unroll
          for (int j4 = 0; j4 < Q2; j4 += 4)
            red_add_f32x4(dst + j4, make_float4(acc[j4], acc[j4 + 1], acc[j4 + 2], acc[j4 + 3]));
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();  // smem tiles, metadata and the TMEM accumulator are reused by the next tile
  }
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// backward: per 32-lookup tile
//   MMA-1  tr0 = A0 * B1                      (recompute, reference K5)
//   SIMT   G = dOut * C2^T  (per lookup, K = q2)      -> sG, sGT          (reference K8, t = 1)
//          dCore2[i2] += tr0^T * dOut (per lookup)    -> red.add          (reference K6/K7, t = 1)
//   MMA-3  dCore0 rows = G * B1^T                     -> red.add          (reference K8/K7, t = 0)
//   MMA-2  dCore1[i1]^T = G^T * A0: ONE GEMM whose K dimension runs over the tile's lookups
//          replaces the per-lookup 16 KB atomic scatter                   (reference K6/K7, t = 0)
// G needs no tensor-core result, so it is computed while MMA-1 runs; the dCore2 stage (needs tr0)
// runs while MMA-2/3 run.  16 warps: thread = (tile row, k-quarter): warp w owns TMEM lane quarter
// w % 4 and the rank indices k in [8*(w/4), 8*(w/4)+8) of every j1 block, so each (lookup, k) of
// dCore2 is reduced by exactly one 4-lane group (one red.add per 16 bytes).
// ---------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 512;

template <int Q2>
struct BwdSmem {
  static constexpr int kA = 128 * 128;         // sA   16 KB
  static constexpr int kAT = 32 * 128 * 4;     // sAT  16 KB
  static constexpr int kB1T = 128 * 128;       // sB1T 16 KB
  static constexpr int kB1 = 32 * 128 * 4;     // sB1  16 KB
  static constexpr int kG = 128 * 128 * 4;     // sG   64 KB
  static constexpr int kGT = 128 * 128 * 4;    // sGT  64 KB
  static constexpr int kC2Stride = R2 * Q2 + kC2StrideBase;
  static constexpr int kC2 = kTileLookups * kC2Stride * 4;
  static constexpr int kMeta = 1024;
  static constexpr int kBytes = 1024 + kA + kAT + kB1T + kB1 + kG + kGT + kC2 + kMeta;
};

// VEC_FLUSH (opt-in, TTB_BWD_VEC_FLUSH=1): the dCore1 flush goes through shared memory so that every
// reduction carries 16 bytes (see flush_d2).
template <int Q2, bool VEC_FLUSH>
__global__ void __launch_bounds__(kBwdThreads, 1)
    tt_bwd_tc_kernel(const ChainDims d, const LookupRec* __restrict__ recs,
                     const int* __restrict__ tile_bucket, const int* __restrict__ tile_begin,
                     const int* __restrict__ tile_count, const int* __restrict__ num_tiles,
                     const int chunk_tiles, const float* __restrict__ d_output, const CorePtrs cores,
                     const CorePtrsRW grads, const int pdl) {
  using SM = BwdSmem<Q2>;
  static_assert(Q2 == 4 || Q2 == 8, "backward epilogue is written for q2 in {4, 8}");
  constexpr int H = Q2 / 4;  // float4s per (row, j1) of dOut and per k of core2
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sAT = sA + SM::kA;
  uint8_t* sB1T = sAT + SM::kAT;
  uint8_t* sB1 = sB1T + SM::kB1T;
  uint8_t* sG = sB1 + SM::kB1;
  uint8_t* sGT = sG + SM::kG;
  float* sC2 = (float*)(sGT + SM::kGT);
  uint8_t* metab = (uint8_t*)sC2 + SM::kC2;
  uint64_t* mbar1 = (uint64_t*)metab;         // MMA-1 done
  uint64_t* mbar2 = (uint64_t*)(metab + 8);   // MMA-2 + MMA-3 done
  uint64_t* mbarC = (uint64_t*)(metab + 16);  // core2 slices landed (TMA)
  uint32_t* tmem_slot = (uint32_t*)(metab + 24);
  TileMeta* meta = (TileMeta*)(metab + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r1 = d.R[1];
  pdl_trigger(pdl);  // the optimizer sweep may be scheduled as SMs drain (it waits before reading gradients)
  pdl_wait(pdl);     // forward / plan / producer of d_output complete and visible
  const int ntiles = *num_tiles;
  if ((int)blockIdx.x * chunk_tiles >= ntiles) return;
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  if (tid == 0) {
    mbar_init(mbar1, 1);
    mbar_init(mbar2, 1);
    mbar_init(mbarC, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < SM::kC2 / 4; i += kBwdThreads) sC2[i] = 0.f;  // padding rows stay finite
  fence_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tD1 = tmem_base, tD3 = tmem_base + 128, tD2 = tmem_base + 160;
  uint32_t phase = 0;
  constexpr uint32_t kIdesc32 = make_idesc_tf32(128, 32, 0, 0);

  const int row = (warp & 3) * 32 + lane;  // TMEM lane == tile row (l, j0) == n index in D2
  const int kq = warp >> 2;                // k-quarter: k in [8*kq, 8*kq + 8)
  const int l = row >> 2, j0 = row & 3;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;

  // dCore1[tb][i1][r][n] += D2[n][r]  (TMEM lane = n): one red.add pass per run of same-bucket tiles.
  // A thread holds D2[n = row][8 kq .. 8 kq + 8): consecutive n sit in consecutive LANES, so the direct flush
  // is one 4-byte reduction per element (4096 lane-operations per flush; REDG issues ~1.3 cycles per lane).
  // VEC_FLUSH transposes through the idle sG tile (every MMA reading it has completed: mbar2 was waited on)
  // into [r][n] rows and issues one 16-byte reduction per four elements instead.
  auto flush_d2 = [&](int bucket) {
    const int tb = bucket / d.p[1];
    const int i1 = bucket - tb * d.p[1];
    float w[8];
    tmem_ld8(tD2 + lane_addr + kq * 8, w);
    tmem_ld_wait();
    if (VEC_FLUSH) {
      float* sT = reinterpret_cast<float*>(sG);  // [32 r][128 n] fp32 = 16 KB of the 64 KB tile
#pragma unroll
      for (int c = 0; c < 8; ++c) sT[(kq * 8 + c) * N1 + row] = w[c];  // lanes -> consecutive n: conflict-free
      __syncthreads();
      float* g1 = grads.c[1] + ((size_t)tb * d.p[1] + i1) * d.S[1];
#pragma unroll
      for (int it = 0; it < (32 * N1 / 4) / kBwdThreads; ++it) {  // 1024 float4 over 512 threads
        const int e = it * kBwdThreads + tid;
        const int r = e >> 5, n4 = (e & 31) << 2;
        if (r < r1) red_add_f32x4(g1 + (size_t)r * N1 + n4, *reinterpret_cast<const float4*>(sT + r * N1 + n4));
      }
    } else {
      float* g1 = grads.c[1] + ((size_t)tb * d.p[1] + i1) * d.S[1] + row;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int r = kq * 8 + c;
        if (r < r1) red_add_f32(g1 + (size_t)r * N1, w[c]);
      }
    }
    tc_fence_before_sync();
    __syncthreads();
  };

  // A CTA takes runs of `chunk_tiles` consecutive tiles.  Tiles of one bucket are consecutive in the
  // plan, so inside a run the core1 slice is staged once per bucket and dCore1 accumulates in TMEM
  // across the bucket's tiles (chunk_tiles == 1 for small batches: pure tile-level load balance).
  for (int chunk = blockIdx.x; chunk * chunk_tiles < ntiles; chunk += gridDim.x) {
    int prev_bucket = -1;
    const int tile_end = min(ntiles, (chunk + 1) * chunk_tiles);
    for (int tile = chunk * chunk_tiles; tile < tile_end; ++tile) {
      const int bucket = tile_bucket[tile];
      const int tb = bucket / d.p[1];
      const int i1 = bucket - tb * d.p[1];
      const int nl = tile_count[tile];
      const bool new_bucket = bucket != prev_bucket;
      if (new_bucket && prev_bucket >= 0) flush_d2(prev_bucket);
      load_tile_meta(meta, tid, nl, recs + tile_begin[tile]);
      if (new_bucket && tid < kFastThreads)
        stage_core1<true>(cores.c[1] + ((size_t)tb * d.p[1] + i1) * d.S[1], r1, sB1T, sB1, tid);
      prev_bucket = bucket;
      __syncthreads();
      tma_core2<Q2>(d, cores.c[2], tb, meta, nl, sC2, mbarC, tid);
      gather_core0<true, kBwdThreads>(d, cores.c[0], tb, meta, nl, sA, sAT, tid);
      const bool valid = l < nl;
      const bool warp_has_rows = (row & ~31) < nl * 4;
      float4 go[Q1][H];  // dOut[l][j0][j1][0..Q2) for all four j1
#pragma unroll
      for (int j1 = 0; j1 < Q1; ++j1)
#pragma unroll
        for (int h = 0; h < H; ++h)
          go[j1][h] = valid ? __ldg(reinterpret_cast<const float4*>(d_output + meta->rec[l].orow +
                                                                   (j0 * Q1 + j1) * Q2 + h * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      fence_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after_sync();
        issue_mma1(tD1, sA, sB1T);
        mma_commit(mbar1);
      }
      // ---- G = dOut * C2^T while MMA-1 runs: row-major into sG, transposed into sGT
      mbar_wait(mbarC, phase);
      if (!warp_has_rows) {
#pragma unroll
        for (int j1 = 0; j1 < Q1; ++j1) {
          const int n0 = j1 * R2 + kq * 8;
          *reinterpret_cast<float4*>(sG + sw128_offset(128, row, n0)) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(sG + sw128_offset(128, row, n0 + 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 8; ++k) *reinterpret_cast<float*>(sGT + sw128_offset(128, n0 + k, row)) = 0.f;
        }
      } else {
        const float* c2 = sC2 + l * SM::kC2Stride + kq * 8 * Q2;
        float g[Q1][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float acc[Q1];
#pragma unroll
          for (int j1 = 0; j1 < Q1; ++j1) acc[j1] = 0.f;
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const float4 w = *reinterpret_cast<const float4*>(c2 + k * Q2 + h * 4);
#pragma unroll
            for (int j1 = 0; j1 < Q1; ++j1)
              acc[j1] = fmaf(go[j1][h].x, w.x,
                             fmaf(go[j1][h].y, w.y, fmaf(go[j1][h].z, w.z, fmaf(go[j1][h].w, w.w, acc[j1]))));
          }
#pragma unroll
          for (int j1 = 0; j1 < Q1; ++j1) g[j1][k] = to_tf32(acc[j1]);
        }
#pragma unroll
        for (int j1 = 0; j1 < Q1; ++j1) {
          const int n0 = j1 * R2 + kq * 8;  // G[row][n0 .. n0+7]
          *reinterpret_cast<float4*>(sG + sw128_offset(128, row, n0)) =
              make_float4(g[j1][0], g[j1][1], g[j1][2], g[j1][3]);
          *reinterpret_cast<float4*>(sG + sw128_offset(128, row, n0 + 4)) =
              make_float4(g[j1][4], g[j1][5], g[j1][6], g[j1][7]);
#pragma unroll
          for (int k = 0; k < 8; ++k) *reinterpret_cast<float*>(sGT + sw128_offset(128, n0 + k, row)) = g[j1][k];
        }
      }
      fence_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {  // D3[128 x 32] = G[128 x 128] * B1^T
          const uint64_t adesc = make_desc_sw128(smem_u32(sG) + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024);
          const uint64_t bdesc = make_desc_sw128(smem_u32(sB1) + (ks >> 2) * (32 * 128) + (ks & 3) * 32, 16, 1024);
          mma_tf32(tD3, adesc, bdesc, kIdesc32, ks > 0);
        }
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {  // D2[128 n x 32 r] (+)= G^T[128 n x 128 rows] * A0[128 rows x 32 r]
          const uint64_t adesc = make_desc_sw128(smem_u32(sGT) + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024);
          const uint64_t bdesc = make_desc_sw128(smem_u32(sAT) + (ks >> 2) * (32 * 128) + (ks & 3) * 32, 16, 1024);
          mma_tf32(tD2, adesc, bdesc, kIdesc32, (!new_bucket) || (ks > 0));
        }
        mma_commit(mbar2);
      }
      // ---- dCore2[i2_l][k][j2] += sum_{j0,j1} tr0[j0][j1][k] * dOut[j0][j1][j2]   (while MMA-2/3 run)
      mbar_wait(mbar1, phase);
      tc_fence_after_sync();
      if (warp_has_rows) {
#pragma unroll
       for (int h = 0; h < H; ++h) {  // one float4 of j2 per pass keeps the accumulators at 32 registers
        float4 part[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) part[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j1 = 0; j1 < Q1; ++j1) {
          float v[8];
          tmem_ld8(tD1 + lane_addr + j1 * R2 + kq * 8, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            part[k].x = fmaf(v[k], go[j1][h].x, part[k].x);
            part[k].y = fmaf(v[k], go[j1][h].y, part[k].y);
            part[k].z = fmaf(v[k], go[j1][h].z, part[k].z);
            part[k].w = fmaf(v[k], go[j1][h].w, part[k].w);
          }
        }
        // reduce over the 4 rows (j0) of this lookup, then lane j0 issues the k with k%4 == j0
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          part[k].x += __shfl_xor_sync(0xffffffffu, part[k].x, 1);
          part[k].y += __shfl_xor_sync(0xffffffffu, part[k].y, 1);
          part[k].z += __shfl_xor_sync(0xffffffffu, part[k].z, 1);
          part[k].w += __shfl_xor_sync(0xffffffffu, part[k].w, 1);
          part[k].x += __shfl_xor_sync(0xffffffffu, part[k].x, 2);
          part[k].y += __shfl_xor_sync(0xffffffffu, part[k].y, 2);
          part[k].z += __shfl_xor_sync(0xffffffffu, part[k].z, 2);
          part[k].w += __shfl_xor_sync(0xffffffffu, part[k].w, 2);
        }
        if (valid) {
          float* g2 = grads.c[2] + ((size_t)tb * d.p[2] + meta->rec[l].i2) * d.S[2] + kq * 8 * Q2 + h * 4;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if ((k & 3) == j0) red_add_f32x4(g2 + k * Q2, part[k]);
        }
       }
      }
      mbar_wait(mbar2, phase);
      phase ^= 1;
      tc_fence_after_sync();
      // ---- dCore0[i0_l][j0][r] += D3[row][r]
      {
        float v[8];
        tmem_ld8(tD3 + lane_addr + kq * 8, v);
        tmem_ld_wait();
        if (valid) {
          float* g0 = grads.c[0] + ((size_t)tb * d.p[0] + meta->rec[l].i0) * d.S[0] + j0 * r1 + kq * 8;
          if (kq * 8 < r1) red_add_f32x4(g0, make_float4(v[0], v[1], v[2], v[3]));
          if (kq * 8 + 4 < r1) red_add_f32x4(g0 + 4, make_float4(v[4], v[5], v[6], v[7]));
        }
      }
      tc_fence_before_sync();
      __syncthreads();
    }
    if (prev_bucket >= 0) flush_d2(prev_bucket);
  }
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// tcgen05 family
bool shape_ok(const ChainDims& d) {
  return d.T == 3 && d.q[0] == 4 && d.q[1] == 4 && d.R[2] == 32 && d.R[1] <= 32 && (d.R[1] % 4) == 0 &&
         (d.q[2] == 4 || d.q[2] == 8) && d.D % 4 == 0 && (long long)d.num_tables * d.p[1] < (1 << 24);
}

#include "ttb_tt_bk.cuh"

// warp-MMA (mma.sync) family: equal ranks 16 / 64 / 128, any q1, q2 in {4, 8}
bool bk_ok(const ChainDims& d) {
  return d.T == 3 && d.q[0] == 4 && d.R[1] == d.R[2] && (d.R[1] == 16 || d.R[1] == 64 || d.R[1] == 128) &&
         (d.q[2] == 4 || d.q[2] == 8) && d.D % 4 == 0 && (long long)d.num_tables * d.p[1] < (1 << 24);
}

template <int R, int Q2>
int launch_fwd_bk_t(const ChainDims& d, const PlanView& p, const CorePtrs& cores, float* output,
                    cudaStream_t stream) {
  using C = bk::Cfg<R, R, Q2, 128>;
  static SmemAttr attr;
  TTB_CUDA(attr.ensure(bk::tt_fwd_bk_kernel<R, R, Q2, 128>, C::kFwdBytes));
  const int per_sm = std::max(1, std::min(4, (227 * 1024) / (C::kFwdBytes + 1024)));
  const long long items = (long long)p.max_tiles * d.q[1];
  const int grid = (int)std::min<long long>(items, (long long)sm_count() * per_sm);
  bk::tt_fwd_bk_kernel<R, R, Q2, 128><<<grid, C::kThreads, C::kFwdBytes, stream>>>(
      d, p.recs, p.tile_bucket, p.tile_begin, p.tile_count, p.num_tiles, cores, output);
  return 0;
}

template <int R, int Q2>
int launch_bwd_bk_t(const ChainDims& d, const PlanView& p, int chunk_tiles, const float* d_output,
                    const CorePtrs& cores, const CorePtrsRW& grads, cudaStream_t stream) {
  using C = bk::Cfg<R, R, Q2, 128>;
  static SmemAttr attr;
  TTB_CUDA(attr.ensure(bk::tt_bwd_bk_kernel<R, R, Q2, 128>, C::kBwdBytes));
  const int per_sm = std::max(1, std::min(2, (227 * 1024) / (C::kBwdBytes + 1024)));
  const long long items = (long long)((p.max_tiles + chunk_tiles - 1) / chunk_tiles) * d.q[1];
  const int grid = (int)std::min<long long>(items, (long long)sm_count() * per_sm);
  bk::tt_bwd_bk_kernel<R, R, Q2, 128><<<grid, C::kThreads, C::kBwdBytes, stream>>>(
      d, p.recs, p.tile_bucket, p.tile_begin, p.tile_count, p.num_tiles, chunk_tiles, d_output, cores, grads);
  return 0;
}

#define TTB_BK_DISPATCH(FN, ...)                                  \
  do {                                                            \
    const int r_ = d.R[1], q2_ = d.q[2];                          \
    if (r_ == 16 && q2_ == 4) return FN<16, 4>(__VA_ARGS__);      \
    if (r_ == 16 && q2_ == 8) return FN<16, 8>(__VA_ARGS__);      \
    if (r_ == 64 && q2_ == 4) return FN<64, 4>(__VA_ARGS__);      \
    if (r_ == 64 && q2_ == 8) return FN<64, 8>(__VA_ARGS__);      \
    if (r_ == 128 && q2_ == 4) return FN<128, 4>(__VA_ARGS__);    \
    if (r_ == 128 && q2_ == 8) return FN<128, 8>(__VA_ARGS__);    \
  } while (0)

int launch_fwd_bk(const ChainDims& d, const PlanView& p, const CorePtrs& cores, float* output, cudaStream_t stream) {
  TTB_BK_DISPATCH(launch_fwd_bk_t, d, p, cores, output, stream);
  set_error("bucketed warp-MMA forward: unsupported shape");
  return 1;
}
int launch_bwd_bk(const ChainDims& d, const PlanView& p, int chunk_tiles, const float* d_output,
                  const CorePtrs& cores, const CorePtrsRW& grads, cudaStream_t stream) {
  TTB_BK_DISPATCH(launch_bwd_bk_t, d, p, chunk_tiles, d_output, cores, grads, stream);
  set_error("bucketed warp-MMA backward: unsupported shape");
  return 1;
}

// core-0 + core-2 gradients up to this size are swept inside the backward kernel by its last 32 CTAs to finish
// (larger ones -- beyond any BASELINE config -- by x_sweep02_kernel launched behind it: a 4th launch).  A single
// sweeping CTA was measured at +29 us for 200 KB against +7 us for the extra launch; 32 sharing it are below either.
constexpr long long kTailSweepMaxFloats = 16LL << 20;

#include "ttb_tt_x.cuh"

// tcgen05 / bf16-operand family (ttb_tt_x.cuh): equal ranks 16 / 32 / 64 / 128
bool x_ok(const ChainDims& d) {
  static const bool legacy = tuning_flag("TTB_LEGACY_TC");  // round-1 tf32 tcgen05 / mma.sync kernels, for A/B runs
  if (legacy) return false;
  static const bool legacy16 = tuning_flag("TTB_LEGACY_R16");  // rank 16 on the warp-level mma.sync kernels (A/B runs)
  const int R = d.R[1];
  if (R == 16 && legacy16) return false;
  const int nb = R == 16 ? 64 : 128;  // xk::BwdBlock<R>::kNB
  return d.T == 3 && d.q[0] == 4 && d.R[2] == R && (R == 16 || R == 32 || R == 64 || R == 128) && (d.q[1] * R) % nb == 0 &&
         (d.q[2] == 4 || d.q[2] == 8) && d.D % 4 == 0 && (long long)d.num_tables * d.p[1] < (1 << 24);
}

template <int R, int Q2, typename CoreT>
int launch_fwd_x_t(const ChainDims& d, const PlanView& p, const CorePtrs& cores, float* output, cudaStream_t stream) {
  constexpr int NB = xk::BwdBlock<R>::kNB;
  using C = xk::XCfg<R, Q2, NB>;
  auto kernel = xk::x_fwd_kernel<R, Q2, CoreT>;
  static SmemAttr attr;
  TTB_CUDA(attr.ensure(kernel, C::kFwdBytes));
  static int cap[16] = {0};
  int& c = cap[current_device() & 15];
  if (c == 0) c = std::max(sm_count(), resident_ctas(kernel, C::kFwdThreads, C::kFwdBytes, 128));
  const long long items = (long long)p.max_tiles * (d.q[1] * R / NB);
  const int grid = (int)std::min<long long>(items, c);
  // the forward has nothing to accumulate across the tiles of a bucket: its work items are single tiles
  kernel<<<grid, C::kFwdThreads, C::kFwdBytes, stream>>>(d, p.recs, p.tile_bucket, p.tile_begin, p.tile_count,
                                                          p.num_tiles, (const CoreT*)cores.c[0],
                                                          (const CoreT*)cores.c[1], (const CoreT*)cores.c[2], output,
                                                          g_trace_fwd);
  return 0;
}

template <int R, int Q2, typename CoreT>
int launch_bwd_x_t(const ChainDims& d, const PlanView& p, int optim, float lr, float eps, const float* d_output,
                   const CorePtrs& cores, const CorePtrsRW& grads, const CorePtrsRW& state, int* sweep_mask,
                   cudaStream_t stream) {
  constexpr int NB = xk::BwdBlock<R>::kNB;
  using C = xk::XCfg<R, Q2, NB>;
  auto kernel = xk::x_bwd_kernel<R, Q2, CoreT>;
  static SmemAttr attr;
  TTB_CUDA(attr.ensure(kernel, C::kBwdBytes));
  static int cap[16] = {0};
  int& c = cap[current_device() & 15];
  if (c == 0) c = std::max(sm_count(), resident_ctas(kernel, C::kBwdThreads, C::kBwdBytes, C::kBwdTmem));
  xk::XBwdArgs a;
  a.recs = p.recs;
  a.run_bucket = p.run_bucket;
  a.run_begin = p.run_begin;
  a.run_count = p.run_count;
  a.num_tiles = p.num_tiles;
  a.bucket_start = p.bucket_start;
  a.sync_words = p.sync_words;
  a.bucket_done = p.bucket_done;
  a.nb = p.nb;
  // cores 0 and 2: swept by the last CTA of this launch when small, by the dense sweep kernel otherwise
  const long long small = (long long)d.num_tables * ((long long)d.p[0] * d.S[0] + (long long)d.p[2] * d.S[2]);
  static const long long tail_max = [] {
    const char* v = getenv("TTB_TAIL_SWEEP_FLOATS");  // tuning override
    return v ? atoll(v) : kTailSweepMaxFloats;
  }();
  a.tail_sweep = small <= tail_max ? 1 : 0;
  *sweep_mask = 0;  // this family applies the optimizer itself (cores 0 / 2: tail of the kernel or the launch below)
  a.d_output = d_output;
  for (int t = 0; t < 3; ++t) {
    a.core[t] = (void*)cores.c[t];
    a.grad[t] = grads.c[t];
    a.state[t] = optim == TTB_OPTIM_ADAGRAD ? state.c[t] : nullptr;
  }
  a.optim = optim;
  a.lr = lr;
  a.eps = eps;
  a.trace = g_trace_bwd;
  const long long items = (long long)p.max_tiles * (d.q[1] * R / NB);
  const int grid = (int)std::min<long long>(items, c);
  kernel<<<grid, C::kBwdThreads, C::kBwdBytes, stream>>>(d, a);
  if (optim != TTB_OPTIM_DENSE && !a.tail_sweep) *sweep_mask = 0x100;  // caller: launch_sweep02_x after this kernel
  return 0;
}

// cores 0 and 2 of the tcgen05 family when they are too large for the backward's last CTA (see kTailSweepMaxFloats)
template <typename CoreT>
int launch_sweep02_x(const ChainDims& d, int optim, float lr, float eps, const CorePtrs& cores, const CorePtrsRW& grads,
                     const CorePtrsRW& state, cudaStream_t stream) {
  const long long small = (long long)d.num_tables * ((long long)d.p[0] * d.S[0] + (long long)d.p[2] * d.S[2]);
  const long long blocks = std::min<long long>((small / 4 + 255) / 256, (long long)sm_count() * 4);
  const bool ada = optim == TTB_OPTIM_ADAGRAD;
  xk::x_sweep02_kernel<CoreT><<<(unsigned)std::max<long long>(1, blocks), 256, 0, stream>>>(
      d, (void*)cores.c[0], (void*)cores.c[2], grads.c[0], grads.c[2], ada ? state.c[0] : nullptr,
      ada ? state.c[2] : nullptr, optim, lr, eps);
  return 0;
}

#define TTB_X_DISPATCH(FN, T, ...)                                 \
  do {                                                             \
    const int r_ = d.R[1], q2_ = d.q[2];                           \
    if (r_ == 16 && q2_ == 4) return FN<16, 4, T>(__VA_ARGS__);    \
    if (r_ == 16 && q2_ == 8) return FN<16, 8, T>(__VA_ARGS__);    \
    if (r_ == 32 && q2_ == 4) return FN<32, 4, T>(__VA_ARGS__);    \
    if (r_ == 32 && q2_ == 8) return FN<32, 8, T>(__VA_ARGS__);    \
    if (r_ == 64 && q2_ == 4) return FN<64, 4, T>(__VA_ARGS__);    \
    if (r_ == 64 && q2_ == 8) return FN<64, 8, T>(__VA_ARGS__);    \
    if (r_ == 128 && q2_ == 4) return FN<128, 4, T>(__VA_ARGS__);  \
    if (r_ == 128 && q2_ == 8) return FN<128, 8, T>(__VA_ARGS__);  \
  } while (0)

int launch_fwd_x(const ChainDims& d, const PlanView& p, const CorePtrs& cores, float* output, bool bf16,
                 cudaStream_t stream) {
  if (bf16)
    TTB_X_DISPATCH(launch_fwd_x_t, __nv_bfloat16, d, p, cores, output, stream);
  else
    TTB_X_DISPATCH(launch_fwd_x_t, float, d, p, cores, output, stream);
  set_error("tcgen05 forward: unsupported shape");
  return 1;
}
int launch_bwd_x(const ChainDims& d, const PlanView& p, int optim, float lr, float eps, const float* d_output,
                 const CorePtrs& cores, const CorePtrsRW& grads, const CorePtrsRW& state, int* sweep_mask, bool bf16,
                 cudaStream_t stream) {
  if (bf16)
    TTB_X_DISPATCH(launch_bwd_x_t, __nv_bfloat16, d, p, optim, lr, eps, d_output, cores, grads, state, sweep_mask, stream);
  else
    TTB_X_DISPATCH(launch_bwd_x_t, float, d, p, optim, lr, eps, d_output, cores, grads, state, sweep_mask, stream);
  set_error("tcgen05 backward: unsupported shape");
  return 1;
}

}  // namespace

bool fast_supported(const ChainDims& d) { return x_ok(d) || shape_ok(d) || bk_ok(d); }
bool bf16_supported(const ChainDims& d) { return x_ok(d); }

size_t fast_workspace_bytes(const ChainDims& d, int64_t nnz) {
  return carve_plan(d, nnz, nullptr).bytes + 256;
}
size_t fast_workspace_header_bytes(const ChainDims& d, int64_t nnz) {
  return carve_plan(d, nnz, nullptr).header_bytes + 256;
}

static PlanIn make_plan_in(const LookupBatch& b) {
  PlanIn in;
  in.indices = (const long long*)b.indices;
  in.rowidx = (const long long*)b.rowidx;
  in.tableidx = (const long long*)b.tableidx;
  in.offsets = (const long long*)b.offsets;
  in.num_bags = b.num_bags;
  in.B = b.B;
  in.mask = b.mask;
  in.nnz = b.nnz;
  in.zero_ptr = nullptr;
  in.zero_n4 = 0;
  in.bags_per_lookup = b.nnz > 0 ? (float)((double)b.num_bags / (double)b.nnz) : 0.f;
  return in;
}

int launch_fwd_fast(const ChainDims& d, const LookupBatch& batch, const CorePtrs& cores, float* output,
                    void* workspace, size_t workspace_bytes, int plan_ready, cudaStream_t stream) {
  const int64_t nnz = batch.nnz;
  TTB_CHECK(nnz < 2147483647LL, "nnz too large for the bucketed path");
  void* ws = (void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  PlanView p = carve_plan(d, nnz, ws);
  TTB_CHECK(workspace && workspace_bytes >= p.bytes + 256, "workspace too small (%zu < %zu)",
            workspace_bytes, p.bytes + 256);
  {
    PlanIn in = make_plan_in(batch);
    const size_t out_floats = (size_t)(d.het ? d.het_tables : d.num_tables) * d.B * d.D;
    if (batch.zero_output && !plan_ready) {  // the plan kernel zero-fills the output on its way
      in.zero_ptr = reinterpret_cast<float4*>(output);
      in.zero_n4 = (long long)(out_floats / 4);
    } else if (batch.zero_output) {
      TTB_CUDA(cudaMemsetAsync(output, 0, out_floats * sizeof(float), stream));
    }
    PlanPrefetch pf;
    memset(&pf, 0, sizeof(pf));
    static const bool no_prefetch = tuning_flag("TTB_NO_PLAN_PREFETCH");
    if (d.T == 3 && !no_prefetch) {
      const int esz = batch.bf16_cores ? 2 : 4;
      for (int t = 0; t < 3; ++t) {
        pf.core[t] = (const char*)cores.c[t];
        pf.slice_bytes[t] = d.S[t] * esz;
      }
      pf.on = 1;
    }
    if (!plan_ready && build_plan(d, in, p, stream, &pf)) return 1;
  }
  const int grid = std::min(p.max_tiles, sm_count() * 4);  // 4 CTAs/SM: 4 x 128 TMEM columns, 4 x 52 KB smem
  KernelTimer timer(TTB_KIND_FWD, stream);
  if (x_ok(d)) {
    if (launch_fwd_x(d, p, cores, output, batch.bf16_cores != 0, stream)) return 1;
    TTB_LAUNCH_CHECK();
    return 0;
  }
  TTB_CHECK(!batch.bf16_cores, "bf16 cores need the tcgen05 kernel family (equal ranks 16 / 32 / 64 / 128)");
  if (!shape_ok(d)) {
    if (launch_fwd_bk(d, p, cores, output, stream)) return 1;
    TTB_LAUNCH_CHECK();
    return 0;
  }
#define TTB_LAUNCH_FWD(Q2)                                                                          \
  do {                                                                                              \
    static SmemAttr attr;                                                                           \
    TTB_CUDA(attr.ensure(tt_fwd_tc_kernel<Q2>, FwdSmem<Q2>::kBytes));                               \
    TTB_CUDA(launch_kernel(pdl, tt_fwd_tc_kernel<Q2>, dim3(grid), dim3(kFastThreads),               \
                           FwdSmem<Q2>::kBytes, stream, d, (const LookupRec*)p.recs,                \
                           (const int*)p.tile_bucket, (const int*)p.tile_begin,                     \
                           (const int*)p.tile_count, (const int*)p.num_tiles, cores, output,        \
                           pdl ? 1 : 0));                                                           \
  } while (0)
  const bool pdl = tuning_flag("TTB_PDL");
  if (d.q[2] == 4)
    TTB_LAUNCH_FWD(4);
  else
    TTB_LAUNCH_FWD(8);
#undef TTB_LAUNCH_FWD
  TTB_LAUNCH_CHECK();
  return 0;
}

int launch_bwd_fast(const ChainDims& d, const LookupBatch& batch, int optim, float lr, float eps,
                    const float* d_output, const CorePtrs& cores, const CorePtrsRW& grads, const CorePtrsRW& state,
                    void* workspace, size_t workspace_bytes, int plan_ready, int* sweep_mask, cudaStream_t stream) {
  *sweep_mask = optim == TTB_OPTIM_DENSE ? 0 : 0x7;  // cores the caller still has to run the dense sweep over
  const int64_t nnz = batch.nnz;
  TTB_CHECK(nnz < 2147483647LL, "nnz too large for the bucketed path");
  void* ws = (void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  PlanView p = carve_plan(d, nnz, ws);
  TTB_CHECK(workspace && workspace_bytes >= p.bytes + 256, "workspace too small (%zu < %zu)",
            workspace_bytes, p.bytes + 256);
  if (!plan_ready && build_plan(d, make_plan_in(batch), p, stream)) return 1;
  // runs of consecutive tiles per CTA visit: ~4 runs per SM for balance, 1 tile per run for small batches
  const long long est_tiles = nnz / kTileLookups + p.nb / 2 + 1;
  const int chunk_tiles = (int)std::max(1LL, std::min(16LL, est_tiles / ((long long)sm_count() * 4)));
  const int grid = std::min((p.max_tiles + chunk_tiles - 1) / chunk_tiles, sm_count());
  if (x_ok(d)) {
    {
      KernelTimer timer(TTB_KIND_BWD, stream);
      if (launch_bwd_x(d, p, optim, lr, eps, d_output, cores, grads, state, sweep_mask, batch.bf16_cores != 0, stream))
        return 1;
      TTB_LAUNCH_CHECK();
    }
    if (*sweep_mask == 0x100) {
      *sweep_mask = 0;
      KernelTimer timer(TTB_KIND_SWEEP, stream);
      if (batch.bf16_cores ? launch_sweep02_x<__nv_bfloat16>(d, optim, lr, eps, cores, grads, state, stream)
                           : launch_sweep02_x<float>(d, optim, lr, eps, cores, grads, state, stream))
        return 1;
      TTB_LAUNCH_CHECK();
    }
    return 0;
  }
  KernelTimer timer(TTB_KIND_BWD, stream);
  TTB_CHECK(!batch.bf16_cores, "bf16 cores need the tcgen05 kernel family (equal ranks 16 / 32 / 64 / 128)");
  if (!shape_ok(d)) {
    if (launch_bwd_bk(d, p, chunk_tiles, d_output, cores, grads, stream)) return 1;
    TTB_LAUNCH_CHECK();
    return 0;
  }
#define TTB_LAUNCH_BWD(Q2, VEC)                                                                     \
  do {                                                                                              \
    static SmemAttr attr;                                                                           \
    TTB_CUDA(attr.ensure(tt_bwd_tc_kernel<Q2, VEC>, BwdSmem<Q2>::kBytes));                          \
    TTB_CUDA(launch_kernel(pdl, tt_bwd_tc_kernel<Q2, VEC>, dim3(grid), dim3(kBwdThreads),           \
                           BwdSmem<Q2>::kBytes, stream, d, (const LookupRec*)p.recs,                \
                           (const int*)p.tile_bucket, (const int*)p.tile_begin,                     \
                           (const int*)p.tile_count, (const int*)p.num_tiles, chunk_tiles,          \
                           d_output, cores, grads, pdl ? 1 : 0));                                   \
  } while (0)
  const bool pdl = tuning_flag("TTB_PDL");
  const bool vec_flush = tuning_flag("TTB_BWD_VEC_FLUSH");
  if (d.q[2] == 4) {
    if (vec_flush)
      TTB_LAUNCH_BWD(4, true);
    else
      TTB_LAUNCH_BWD(4, false);
  } else {
    if (vec_flush)
      TTB_LAUNCH_BWD(8, true);
    else
      TTB_LAUNCH_BWD(8, false);
  }
#undef TTB_LAUNCH_BWD
  TTB_LAUNCH_CHECK();
  return 0;
}

}  // namespace ttb
