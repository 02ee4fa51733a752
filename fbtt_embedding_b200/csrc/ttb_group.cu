// Table groups (include/ttb.h): one host call per phase for a rank's heterogeneous
// tables.  No device code of its own: every item goes through the same entry points as a per-table
// call (ttb_preprocess_rowidx / ttb_tt_forward / ttb_tt_backward), so kernels and numerics are
// identical; what disappears is the per-table host cost above the ABI (Python, ctypes, autograd).
//
// Optional lanes: with ttb_group_set_streams(k > 1) item i is enqueued on lane i % k.  Lane 0 is the
// caller's stream; lanes 1..k-1 are library-owned non-blocking streams that fork from the caller's
// stream (event record + wait) before the first item and join back into it after the last, so the
// call is still "everything is ordered on `stream`" for the caller (and for torch's caching
// allocator: all buffers were allocated on `stream` and every lane is joined before the call
// returns).  Event record / wait is the capturable cross-stream pattern, so a group call can sit
// inside a CUDA graph.
#include <atomic>
#include <mutex>
#include <vector>

#include "ttb_common.cuh"

namespace ttb {

void set_onepass_share(int k);  // ttb_tt_fast.cu: lanes that may run single-launch plan kernels at once

namespace {

constexpr int kMaxLanes = 16;
std::atomic<int> g_lanes{1};
std::mutex g_mu;

struct Lanes {
  std::vector<cudaStream_t> streams;  // side lanes 1..k-1
  std::vector<cudaEvent_t> done;
  cudaEvent_t fork = nullptr;
};
Lanes g_dev_lanes[16];

int ensure_lanes(Lanes& L, int side) {
  if (!L.fork) TTB_CUDA(cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming));
  while ((int)L.streams.size() < side) {
    cudaStream_t s;
    cudaEvent_t e;
    TTB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    TTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    L.streams.push_back(s);
    L.done.push_back(e);
  }
  return 0;
}

// runs per_item(i, lane_stream) for every item with nnz > 0
template <typename F>
int run_group(int n_items, const ttb_group_item_t* items, cudaStream_t stream, F per_item) {
  TTB_CHECK(n_items >= 0, "n_items must be >= 0");
  if (n_items == 0) return 0;
  TTB_CHECK(items != nullptr, "items is NULL");
  int active = 0;
  for (int i = 0; i < n_items; ++i) {
    TTB_CHECK(items[i].nnz >= 0, "item %d: nnz must be >= 0", i);
    active += items[i].nnz > 0;
  }
  const int k = std::min(std::min(g_lanes.load(std::memory_order_relaxed), active), kMaxLanes);
  if (k <= 1) {
    for (int i = 0; i < n_items; ++i)
      if (items[i].nnz > 0 && per_item(i, stream)) return 1;
    return 0;
  }
  std::lock_guard<std::mutex> lk(g_mu);  // the fork / done events are shared per device
  Lanes& L = g_dev_lanes[current_device() & 15];
  if (ensure_lanes(L, k - 1)) return 1;
  TTB_CUDA(cudaEventRecord(L.fork, stream));
  for (int s = 0; s < k - 1; ++s) TTB_CUDA(cudaStreamWaitEvent(L.streams[s], L.fork, 0));
  // k plan kernels may now be co-resident: each gets 1/k of the single-launch plan's CTA budget
  set_onepass_share(k);
  int rc = 0, lane = 0;
  for (int i = 0; i < n_items && !rc; ++i) {
    if (items[i].nnz == 0) continue;
    rc = per_item(i, lane == 0 ? stream : L.streams[lane - 1]);
    lane = (lane + 1) % k;
  }
  set_onepass_share(1);
  // join even after a failure: a stream capture must not be left forked
  for (int s = 0; s < k - 1; ++s) {
    const cudaError_t e1 = cudaEventRecord(L.done[s], L.streams[s]);
    const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(stream, L.done[s], 0) : e1;
    if (e2 != cudaSuccess && !rc) {
      set_error("table group: joining lane %d failed: %s", s + 1, cudaGetErrorString(e2));
      rc = 1;
    }
  }
  return rc;
}

// prefixes the item number to whatever the per-table entry point reported
int item_failed(int i) {
  char msg[400];
  snprintf(msg, sizeof(msg), "%s", ttb_last_error());
  set_error("table group item %d: %s", i, msg);
  return 1;
}

}  // namespace
}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_group_set_streams(int k) {
  TTB_CHECK(k >= 1 && k <= kMaxLanes, "ttb_group_set_streams: k=%d not in [1, %d]", k, kMaxLanes);
  g_lanes.store(k);
  return 0;
}
int ttb_group_get_streams(void) { return g_lanes.load(); }

int ttb_group_preprocess(int n_items, const ttb_group_item_t* items, cudaStream_t stream) {
  return run_group(n_items, items, stream, [&](int i, cudaStream_t s) {
    const ttb_group_item_t& it = items[i];
    const int64_t bags = (int64_t)it.shape.num_tables * it.shape.B;
    return ttb_preprocess_rowidx(it.nnz, bags, it.shape.B, it.offsets, it.rowidx, it.tableidx, s)
               ? item_failed(i) : 0;
  });
}

int ttb_group_forward(int n_items, const ttb_group_item_t* items, cudaStream_t stream) {
  return run_group(n_items, items, stream, [&](int i, cudaStream_t s) {
    const ttb_group_item_t& it = items[i];
    return ttb_tt_forward(&it.shape, it.nnz, it.indices, it.rowidx, it.tableidx, it.cores, it.output,
                          it.workspace, it.workspace_bytes, it.plan_ready, s)
               ? item_failed(i) : 0;
  });
}

int ttb_group_backward(int n_items, const ttb_group_item_t* items, int optim, float lr, float eps,
                       cudaStream_t stream) {
  return run_group(n_items, items, stream, [&](int i, cudaStream_t s) {
    const ttb_group_item_t& it = items[i];
    return ttb_tt_backward(&it.shape, optim, lr, eps, it.nnz, it.indices, it.rowidx, it.tableidx,
                           it.d_output, it.cores, it.grads,
                           optim == TTB_OPTIM_ADAGRAD ? it.opt_state : nullptr, it.workspace,
                           it.workspace_bytes, it.plan_ready, s)
               ? item_failed(i) : 0;
  });
}

}  // extern "C"
