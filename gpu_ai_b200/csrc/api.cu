// gpu_ai_b200/csrc/api.cu -- the C ABI of include/b2p.h: contexts, buffers, sharding, launches.
//
// Host-side replacement for the four Device*PlayoutDriver::runPlayouts launchers of the
// reference (src/singlePlayout.cu:71-120, src/multiplePlayout.cu:53-98,
// src/coarsePlayout.cu:91-163, src/heuristicPlayout.cu:102-147) and for genMovesTest
// (src/genMovesTest.cu:26-100).  Differences by design: grow-only buffers and pinned staging
// owned by a context instead of cudaMalloc/cudaFree per call; asynchronous copies on
// per-device streams; no process-global state (the reference sets cudaLimitStackSize on
// every call); error codes instead of exit(1); leaves are packed to 16 B before they cross
// PCIe (the reference ships 776 B per state).
#include "../../include/b2p.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "hostpool.h"
#include "pack776.h"
#include "kernels.cuh"

using namespace b2p;

namespace {

thread_local std::string g_create_error;

struct Buffer {
  void *ptr = nullptr;
  size_t cap = 0;
  bool pinned = false;
};

// one pipeline slot of the asynchronous count call (b2p_run_counts_async): its own device buffers, runs on aux[slot]
constexpr int kSlots = 4;
struct Slot {
  Buffer d_states, d_counts, d_misc, h_misc;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // around the kernel: b2p_wait_slot reports the kernel time
};

struct Device {
  int id = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux[4] = {nullptr, nullptr, nullptr, nullptr};  // chunk pipeline of b2p_run_states776
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Buffer d_states, d_winners, d_plies, d_final, d_moves, d_counts, d_misc;
  Buffer h_states, h_winners, h_misc;  // pinned staging
  unsigned int *d_next_ring = nullptr;  // kRing work-queue heads, rotated per launch
  unsigned ring_pos = 0;  // advanced under the caller's serialisation (one caller per context)
  // one event per ring slot, recorded behind the launch that uses the slot: a slot is only handed out again
  // after that launch has finished, whatever stream it ran on (launches may sit on caller-supplied streams)
  std::vector<cudaEvent_t> ring_ev;
  std::vector<char> ring_busy;
  unsigned next_aux = 0;  // round-robin over `aux` per DEVICE (segments of b2p_run_states776)
  Slot slots[kSlots];
};


}  // namespace

struct b2p_ctx {
  std::vector<Device> devs;
  uint64_t seed = 0;
  uint64_t calls = 0;
  std::atomic<uint64_t> launches{0};
  std::string err;
  b2p::Pool pool;
  int slot_span[kSlots] = {0, 0, 0, 0};  // devices used by the call in flight on each slot (0 = idle)
  Buffer slot_leaves[kSlots], slot_wins[kSlots];  // page-locked staging lent to pipelined callers (b2p_slot_staging)
};

namespace {

int fail(b2p_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  else g_create_error = msg;
  return code;
}

#define B2P_CUDA(ctx, call)                                                                          \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(ctx, B2P_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
  } while (0)

int ensure(b2p_ctx *ctx, Buffer &b, size_t bytes, bool pinned) {
  if (bytes <= b.cap) return B2P_OK;
  size_t cap = std::max(bytes, b.cap + b.cap / 2);
  if (b.ptr) {
    if (b.pinned) cudaFreeHost(b.ptr);
    else cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
  }
  cudaError_t e = pinned ? cudaMallocHost(&b.ptr, cap) : cudaMalloc(&b.ptr, cap);
  if (e != cudaSuccess) return fail(ctx, B2P_ENOMEM, std::string("allocation of ") + std::to_string(cap) + " bytes failed: " + cudaGetErrorString(e));
  b.cap = cap;
  b.pinned = pinned;
  return B2P_OK;
}

void release(Buffer &b) {
  if (!b.ptr) return;
  if (b.pinned) cudaFreeHost(b.ptr);
  else cudaFree(b.ptr);
  b.ptr = nullptr;
  b.cap = 0;
}

constexpr unsigned kRing = 256;

// B2P_SCHED_AUTO: warp-per-playout only pays below the batch size at which thread-per-playout
// fills the machine's latency budget (measured on B200, profiles/: see DESIGN.md 4.5)
// (profiles/r01_final_sched_compare.jsonl: both modes cross over between 4096 and 8192 playouts per launch)
constexpr size_t kAutoWarpMaxRandom = 4096, kAutoWarpMaxHeuristic = 4096;
bool use_warp_kernel(int sched, size_t total, KernelMode km) {
  return sched == B2P_SCHED_WARP ||
         (sched == B2P_SCHED_AUTO && total <= (km == kHeuristic ? kAutoWarpMaxHeuristic : kAutoWarpMaxRandom));
}

// One playout launch on device `d`: takes the next work-queue head of the ring (waiting, if the ring has
// wrapped, for the launch that last used it -- so any number of launches may be in flight on any streams),
// launches, and records the slot's event behind the kernel.
cudaError_t launch_playouts(b2p_ctx *ctx, Device &d, PlayoutParams &prm, KernelMode km, int sched, cudaStream_t st) {
  const unsigned slot = d.ring_pos++ % kRing;
  cudaError_t e;
  if (d.ring_busy[slot] && (e = cudaEventSynchronize(d.ring_ev[slot])) != cudaSuccess) return e;
  prm.next = d.d_next_ring + slot;
  e = use_warp_kernel(sched, prm.total, km) ? launch_playout_warp(prm, km, d.sm_count, st, nullptr)
                                            : launch_playout_lanes(prm, km, d.sm_count, st, nullptr);
  if (e != cudaSuccess) return e;
  ctx->launches++;
  d.ring_busy[slot] = 1;
  return cudaEventRecord(d.ring_ev[slot], st);
}

bool mode_to_kernel(int mode, int order, KernelMode *out) {
  if (mode == B2P_MODE_RANDOM && order == B2P_ORDER_CANONICAL) { *out = kRandomCanonical; return true; }
  if (mode == B2P_MODE_RANDOM && order == B2P_ORDER_FAST) { *out = kRandomFast; return true; }
  if (mode == B2P_MODE_HEURISTIC) { *out = kHeuristic; return true; }
  return false;
}

// ---- reference layout (probed: SURVEY.md 8a) ---------------------------------------------------
constexpr size_t kStateBytes = 776, kItemBytes = 12, kTurnOff = 768, kMscOff = 772;
constexpr size_t kMoveBytes = 38;

inline void unpack_one(const b2p_state16 *p, unsigned char *s) {
  std::memset(s, 0, kStateBytes);
  for (int i = 0; i < 32; i++) {
    const int r = i >> 2, c = 2 * (i & 3) + ((r & 1) ^ 1);
    unsigned char *q = s + kItemBytes * (size_t)(r * 8 + c);
    const uint32_t bit = 1u << i;
    if (!((p->p1 | p->p2) & bit)) continue;
    q[0] = 1;
    const int32_t type = (p->kings & bit) ? 1 : 0, owner = (p->p1 & bit) ? 0 : 1;
    std::memcpy(q + 4, &type, 4);
    std::memcpy(q + 8, &owner, 4);
  }
  const int32_t turn = (int32_t)(p->meta & 1u);
  const uint32_t msc = p->meta >> 8;
  std::memcpy(s + kTurnOff, &turn, 4);
  std::memcpy(s + kMscOff, &msc, 4);
}

template <class F>
void parallel_for(size_t n, size_t grain, F &&f) {
  unsigned hw = std::thread::hardware_concurrency();
  size_t workers = std::min<size_t>(hw ? hw : 1, 32);
  workers = std::min(workers, (n + grain - 1) / grain);
  if (workers <= 1) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  const size_t chunk = (n + workers - 1) / workers;
  for (size_t t = 0; t < workers; t++) {
    const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=, &f] { f(lo, hi); });
  }
  for (auto &t : th) t.join();
}

// sharding policy of the host-buffer calls: a second device is only used once every shard keeps at least
// kMinShard playouts (a 50-leaf MCTS batch spread over 8 GPUs costs 8 copies and 8 launches for nothing:
// measured 0.46 ms vs 0.10 ms, profiles/r01x_sweep_tests_sh_protocol_8gpu.jsonl)
constexpr size_t kMinShard = 8192;
constexpr size_t kPackGrain = 512;    // reference States per host pack task (400 KB of `State`s, ~35 us)
constexpr size_t kSegmentMin = 8192, kSegmentMax = 32768;  // leaves per launch of the 776-byte pipeline
int device_span(const b2p_ctx *ctx, size_t n, size_t work) {
  const size_t by_work = std::max<size_t>(1, work / kMinShard);
  return (int)std::min(std::min(ctx->devs.size(), n), by_work);
}

struct Shard {
  size_t lo, hi;
};
Shard shard_of(size_t n, int g, int G) { return {n * (size_t)g / G, n * (size_t)(g + 1) / G}; }

}  // namespace

// ==================================================================================================
extern "C" {

const char *b2p_version(void) { return "b2p 0.1 (sm_100a)"; }

const char *b2p_last_error(const b2p_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b2p_create(b2p_ctx **out, const int *device_ids, int n_dev, uint64_t seed) {
  if (!out) return fail(nullptr, B2P_EINVAL, "b2p_create: out is NULL");
  *out = nullptr;
  int visible = 0;
  cudaError_t e = cudaGetDeviceCount(&visible);
  if (e != cudaSuccess || visible == 0)
    return fail(nullptr, B2P_ENODEV, std::string("no CUDA device: ") + (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)) +
                                         " (this library has no CPU execution path)");
  if (n_dev <= 0) n_dev = visible;
  b2p_ctx *ctx = new b2p_ctx();
  ctx->seed = seed;
  for (int i = 0; i < n_dev; i++) {
    Device d;
    d.id = device_ids ? device_ids[i] : i;
    if (d.id < 0 || d.id >= visible) {
      b2p_destroy(ctx);
      return fail(nullptr, B2P_EINVAL, "b2p_create: device id out of range");
    }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(d.id)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, d.id)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&d.ev0)) != cudaSuccess || (e = cudaEventCreate(&d.ev1)) != cudaSuccess ||
        (e = cudaMalloc(&d.d_next_ring, kRing * sizeof(unsigned int))) != cudaSuccess) {
      ctx->devs.push_back(d);
      b2p_destroy(ctx);
      return fail(nullptr, B2P_ECUDA, std::string("b2p_create: ") + cudaGetErrorString(e));
    }
    if (prop.major < 10) {
      ctx->devs.push_back(d);
      b2p_destroy(ctx);
      return fail(nullptr, B2P_ENODEV, std::string("b2p_create: device ") + prop.name + " is not sm_100 class; the kernels are built for sm_100a only");
    }
    d.sm_count = prop.multiProcessorCount;
    d.ring_ev.assign(kRing, nullptr);
    d.ring_busy.assign(kRing, 0);
    for (cudaEvent_t &ev : d.ring_ev)
      if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) break;
    if (e != cudaSuccess) {
      ctx->devs.push_back(d);
      b2p_destroy(ctx);
      return fail(nullptr, B2P_ECUDA, std::string("b2p_create: ") + cudaGetErrorString(e));
    }
    for (cudaStream_t &a : d.aux)
      if (cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking) != cudaSuccess) a = nullptr;
    ctx->devs.push_back(d);
  }
  *out = ctx;
  return B2P_OK;
}

void b2p_destroy(b2p_ctx *ctx) {
  if (!ctx) return;
  for (int sl = 0; sl < kSlots; sl++) {
    release(ctx->slot_leaves[sl]);
    release(ctx->slot_wins[sl]);
  }
  for (Device &d : ctx->devs) {
    if (d.id < 0) continue;
    cudaSetDevice(d.id);
    if (d.stream) cudaStreamSynchronize(d.stream);
    for (cudaStream_t a : d.aux)
      if (a) cudaStreamSynchronize(a);
    for (cudaEvent_t ev : d.ring_ev)
      if (ev) cudaEventDestroy(ev);
    for (Slot &sl : d.slots) {
      for (Buffer *b : {&sl.d_states, &sl.d_counts, &sl.d_misc, &sl.h_misc}) release(*b);
      if (sl.ev0) cudaEventDestroy(sl.ev0);
      if (sl.ev1) cudaEventDestroy(sl.ev1);
    }
    for (Buffer *b : {&d.d_states, &d.d_winners, &d.d_plies, &d.d_final, &d.d_moves, &d.d_counts, &d.d_misc, &d.h_states, &d.h_winners, &d.h_misc}) release(*b);
    if (d.d_next_ring) cudaFree(d.d_next_ring);
    if (d.ev0) cudaEventDestroy(d.ev0);
    if (d.ev1) cudaEventDestroy(d.ev1);
    for (cudaStream_t a : d.aux)
      if (a) cudaStreamDestroy(a);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  delete ctx;
}

int b2p_device_count(const b2p_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int b2p_device_info(const b2p_ctx *ctx, int dev_index, b2p_devinfo *out) {
  if (!ctx || !out || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return B2P_EINVAL;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, ctx->devs[dev_index].id) != cudaSuccess) return B2P_ECUDA;
  std::memset(out, 0, sizeof *out);
  out->device_id = ctx->devs[dev_index].id;
  out->sm_count = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->devs[dev_index].id);
  out->clock_khz = khz;
  out->cc_major = prop.major;
  out->cc_minor = prop.minor;
  out->total_mem = prop.totalGlobalMem;
  std::snprintf(out->name, sizeof out->name, "%s", prop.name);
  return B2P_OK;
}

uint64_t b2p_launch_count(const b2p_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int b2p_sync(b2p_ctx *ctx) {
  if (!ctx) return B2P_EINVAL;
  for (Device &d : ctx->devs) {
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    B2P_CUDA(ctx, cudaStreamSynchronize(d.stream));
  }
  return B2P_OK;
}

// ---- caller-visible pinned host memory -----------------------------------------------------------------
int b2p_alloc_host(void **out, size_t bytes) {
  if (!out) return B2P_EINVAL;
  *out = nullptr;
  if (bytes == 0) return B2P_OK;
  const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    *out = nullptr;
    return fail(nullptr, e == cudaErrorMemoryAllocation ? B2P_ENOMEM : B2P_ECUDA, std::string("b2p_alloc_host: ") + cudaGetErrorString(e));
  }
  return B2P_OK;
}

int b2p_free_host(void *ptr) {
  if (!ptr) return B2P_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? B2P_OK : B2P_ECUDA;
}

// ---- converters ------------------------------------------------------------------------------------
int b2p_pack776(const void *states, size_t n, b2p_state16 *out) {
  if (n && (!states || !out)) return B2P_EINVAL;
  const unsigned char *s = (const unsigned char *)states;
  parallel_for(n, 8192, [&](size_t lo, size_t hi) {
    pack776_range(s + kStateBytes * lo, hi - lo, out + lo);
  });
  return B2P_OK;
}

const char *b2p_pack776_impl(void) { return pack776_impl(); }

int b2p_unpack776(const b2p_state16 *states, size_t n, void *states_out) {
  if (n && (!states || !states_out)) return B2P_EINVAL;
  unsigned char *s = (unsigned char *)states_out;
  parallel_for(n, 8192, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) unpack_one(states + i, s + kStateBytes * i);
  });
  return B2P_OK;
}

int b2p_expand_move(b2p_move_t m, void *move38_out) {
  if (!move38_out) return B2P_EINVAL;
  unsigned char *o = (unsigned char *)move38_out;
  std::memset(o, 0, kMoveBytes);
  auto row = [](int i) { return i >> 2; };
  auto col = [](int i) { return 2 * (i & 3) + (((i >> 2) & 1) ^ 1); };
  const int from = (int)(m & 31), to = (int)((m >> 5) & 31), hops = (int)((m >> 10) & 7);
  o[0] = (unsigned char)row(from); o[1] = (unsigned char)col(from);
  o[2] = (unsigned char)row(to);   o[3] = (unsigned char)col(to);
  int prev = from;
  for (int k = 0; k < hops; k++) {
    const int land = (int)((m >> (16 + 5 * k)) & 31);
    o[4 + 2 * k] = (unsigned char)((row(prev) + row(land)) / 2);      // Move::removed[k]
    o[5 + 2 * k] = (unsigned char)((col(prev) + col(land)) / 2);
    o[20 + 2 * k] = (unsigned char)row(land);                          // Move::intermediate[k]
    o[21 + 2 * k] = (unsigned char)col(land);
    prev = land;
  }
  o[36] = (unsigned char)hops;
  o[37] = (unsigned char)((m >> 13) & 1);
  return B2P_OK;
}

// ---- device-resident calls ----------------------------------------------------------------------------
int b2p_run_packed_device(b2p_ctx *ctx, int dev_index, const b2p_state16 *d_states, size_t n, uint32_t reps,
                          uint64_t key, uint64_t pid_base, int mode, int sched, int order, int max_plies,
                          int8_t *d_winners, uint32_t *d_plies, b2p_state16 *d_final, uint64_t *d_counters,
                          void *cuda_stream) {
  if (!ctx) return B2P_EINVAL;
  if (dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(ctx, B2P_EINVAL, "bad device index");
  if (n == 0 || reps == 0) return B2P_OK;
  if (!d_states) return fail(ctx, B2P_EINVAL, "d_states is NULL");
  KernelMode km;
  if (!mode_to_kernel(mode, order, &km)) return fail(ctx, B2P_EINVAL, "unknown mode/order");
  if ((unsigned long long)n * reps >= (1ull << 31)) return fail(ctx, B2P_EINVAL, "n*reps must be < 2^31 per launch");
  Device &d = ctx->devs[dev_index];
  B2P_CUDA(ctx, cudaSetDevice(d.id));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PlayoutParams prm;
  std::memset(&prm, 0, sizeof prm);
  prm.states = reinterpret_cast<const uint4 *>(d_states);
  prm.n = (uint32_t)n;
  prm.total = (uint32_t)(n * reps);
  prm.rep_stride = n;
  prm.key = key;
  prm.pid_base = pid_base;
  prm.max_plies = max_plies;
  prm.winners = d_winners;
  prm.plies = d_plies;
  prm.final_states = reinterpret_cast<uint4 *>(d_final);
  prm.counters = reinterpret_cast<unsigned long long *>(d_counters);
  const cudaError_t e = launch_playouts(ctx, d, prm, km, sched, st);
  if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("playout launch: ") + cudaGetErrorString(e));
  return B2P_OK;
}

int b2p_genmoves_device(b2p_ctx *ctx, int dev_index, const b2p_state16 *d_states, size_t n, int max_moves,
                        b2p_move_t *d_moves, uint8_t *d_counts, void *cuda_stream) {
  if (!ctx) return B2P_EINVAL;
  if (dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(ctx, B2P_EINVAL, "bad device index");
  if (n == 0) return B2P_OK;
  if (!d_states || !d_moves || !d_counts || max_moves <= 0 || n >= (1ull << 31)) return fail(ctx, B2P_EINVAL, "bad genmoves arguments");
  Device &d = ctx->devs[dev_index];
  B2P_CUDA(ctx, cudaSetDevice(d.id));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaError_t e = launch_genmoves(reinterpret_cast<const uint4 *>(d_states), (uint32_t)n, max_moves,
                                  reinterpret_cast<unsigned long long *>(d_moves), d_counts, st);
  if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("genmoves launch: ") + cudaGetErrorString(e));
  ctx->launches++;
  return B2P_OK;
}

int b2p_gen_leaves_device(b2p_ctx *ctx, int dev_index, size_t n, uint64_t key, uint64_t first_index, b2p_state16 *d_out,
                          void *cuda_stream) {
  if (!ctx) return B2P_EINVAL;
  if (dev_index < 0 || dev_index >= (int)ctx->devs.size()) return fail(ctx, B2P_EINVAL, "bad device index");
  if (n == 0) return B2P_OK;
  if (!d_out || n >= (1ull << 31)) return fail(ctx, B2P_EINVAL, "bad gen_leaves arguments");
  Device &d = ctx->devs[dev_index];
  B2P_CUDA(ctx, cudaSetDevice(d.id));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PlayoutParams prm;
  std::memset(&prm, 0, sizeof prm);
  prm.n = (uint32_t)n;
  prm.total = (uint32_t)n;
  prm.rep_stride = n;
  prm.key = key;
  prm.pid_base = first_index;
  prm.max_plies = 0;
  prm.final_states = reinterpret_cast<uint4 *>(d_out);
  const cudaError_t e = launch_playouts(ctx, d, prm, kLeafGen, B2P_SCHED_THREAD, st);
  if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("leafgen launch: ") + cudaGetErrorString(e));
  return B2P_OK;
}

// ---- host-buffer calls ---------------------------------------------------------------------------------
int b2p_run_packed(b2p_ctx *ctx, const b2p_state16 *states, size_t n, uint32_t reps, uint64_t key, uint64_t pid_base,
                   int mode, int sched, int order, int max_plies, int8_t *winners_out, uint32_t *plies_out,
                   b2p_state16 *final_out, uint64_t counters_out[4]) {
  if (!ctx) return B2P_EINVAL;
  if (counters_out) counters_out[0] = counters_out[1] = counters_out[2] = counters_out[3] = 0;
  if (n == 0 || reps == 0) return B2P_OK;
  if (!states) return fail(ctx, B2P_EINVAL, "states is NULL");
  KernelMode km;
  if (!mode_to_kernel(mode, order, &km)) return fail(ctx, B2P_EINVAL, "unknown mode/order");
  const int G = device_span(ctx, n, n * (size_t)reps);
  // phase 1: enqueue everything on every device (copies and kernels are asynchronous)
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    const Shard sh = shard_of(n, g, G);
    const size_t nl = sh.hi - sh.lo, tl = nl * reps;
    if ((unsigned long long)tl >= (1ull << 31)) return fail(ctx, B2P_EINVAL, "per-device n*reps must be < 2^31");
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    int rc;
    if ((rc = ensure(ctx, d.d_states, nl * sizeof(b2p_state16), false))) return rc;
    if ((rc = ensure(ctx, d.d_misc, 4 * sizeof(uint64_t), false))) return rc;
    if ((rc = ensure(ctx, d.h_misc, 4 * sizeof(uint64_t), true))) return rc;
    if (winners_out && (rc = ensure(ctx, d.d_winners, tl, false))) return rc;
    if (plies_out && (rc = ensure(ctx, d.d_plies, tl * sizeof(uint32_t), false))) return rc;
    if (final_out && (rc = ensure(ctx, d.d_final, tl * sizeof(b2p_state16), false))) return rc;
    B2P_CUDA(ctx, cudaMemcpyAsync(d.d_states.ptr, states + sh.lo, nl * sizeof(b2p_state16), cudaMemcpyHostToDevice, d.stream));
    B2P_CUDA(ctx, cudaMemsetAsync(d.d_misc.ptr, 0, 4 * sizeof(uint64_t), d.stream));
    PlayoutParams prm;
    std::memset(&prm, 0, sizeof prm);
    prm.states = reinterpret_cast<const uint4 *>(d.d_states.ptr);
    prm.n = (uint32_t)nl;
    prm.total = (uint32_t)tl;
    prm.rep_stride = n;
    prm.key = key;
    prm.pid_base = pid_base + sh.lo;
    prm.max_plies = max_plies;
    prm.winners = winners_out ? (int8_t *)d.d_winners.ptr : nullptr;
    prm.plies = plies_out ? (uint32_t *)d.d_plies.ptr : nullptr;
    prm.final_states = final_out ? (uint4 *)d.d_final.ptr : nullptr;
    prm.counters = (unsigned long long *)d.d_misc.ptr;
    const cudaError_t e = launch_playouts(ctx, d, prm, km, sched, d.stream);
    if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("playout launch: ") + cudaGetErrorString(e));
    // gather: device layout [rep][local leaf] -> host layout [rep][global leaf]
    if (winners_out)
      B2P_CUDA(ctx, cudaMemcpy2DAsync(winners_out + sh.lo, n, d.d_winners.ptr, nl, nl, reps, cudaMemcpyDeviceToHost, d.stream));
    if (plies_out)
      B2P_CUDA(ctx, cudaMemcpy2DAsync(plies_out + sh.lo, n * 4, d.d_plies.ptr, nl * 4, nl * 4, reps, cudaMemcpyDeviceToHost, d.stream));
    if (final_out)
      B2P_CUDA(ctx, cudaMemcpy2DAsync(final_out + sh.lo, n * 16, d.d_final.ptr, nl * 16, nl * 16, reps, cudaMemcpyDeviceToHost, d.stream));
    B2P_CUDA(ctx, cudaMemcpyAsync(d.h_misc.ptr, d.d_misc.ptr, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream));
  }
  // phase 2: wait and combine the per-device counters on the host ("host gather" of north_star)
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    B2P_CUDA(ctx, cudaStreamSynchronize(d.stream));
    if (counters_out)
      for (int k = 0; k < 4; k++) counters_out[k] += ((const uint64_t *)d.h_misc.ptr)[k];
  }
  return B2P_OK;
}

// K playouts per leaf, win COUNTS per leaf back: what a tree search needs from a batch.
// 16 B per leaf up, 8 B per leaf down, whatever `reps` is; the per-playout winners never leave the device.
int b2p_run_counts(b2p_ctx *ctx, const b2p_state16 *states, size_t n, uint32_t reps, uint64_t key, uint64_t pid_base,
                   int mode, int sched, int order, uint32_t *wins_out, uint64_t counters_out[4]) {
  if (!ctx) return B2P_EINVAL;
  if (counters_out) counters_out[0] = counters_out[1] = counters_out[2] = counters_out[3] = 0;
  if (n == 0 || reps == 0) return B2P_OK;
  if (!states || !wins_out) return fail(ctx, B2P_EINVAL, "NULL buffer");
  KernelMode km;
  if (!mode_to_kernel(mode, order, &km)) return fail(ctx, B2P_EINVAL, "unknown mode/order");
  const int G = device_span(ctx, n, n * (size_t)reps);
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    const Shard sh = shard_of(n, g, G);
    const size_t nl = sh.hi - sh.lo, tl = nl * reps;
    if ((unsigned long long)tl >= (1ull << 31)) return fail(ctx, B2P_EINVAL, "per-device n*reps must be < 2^31");
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    int rc;
    if ((rc = ensure(ctx, d.d_states, nl * sizeof(b2p_state16), false))) return rc;
    if ((rc = ensure(ctx, d.d_plies, nl * 2 * sizeof(uint32_t), false))) return rc;  // reused as the per-leaf count buffer
    if ((rc = ensure(ctx, d.d_misc, 4 * sizeof(uint64_t), false))) return rc;
    if ((rc = ensure(ctx, d.h_misc, 4 * sizeof(uint64_t), true))) return rc;
    B2P_CUDA(ctx, cudaMemcpyAsync(d.d_states.ptr, states + sh.lo, nl * sizeof(b2p_state16), cudaMemcpyHostToDevice, d.stream));
    B2P_CUDA(ctx, cudaMemsetAsync(d.d_misc.ptr, 0, 4 * sizeof(uint64_t), d.stream));
    B2P_CUDA(ctx, cudaMemsetAsync(d.d_plies.ptr, 0, nl * 2 * sizeof(uint32_t), d.stream));
    PlayoutParams prm;
    std::memset(&prm, 0, sizeof prm);
    prm.states = reinterpret_cast<const uint4 *>(d.d_states.ptr);
    prm.n = (uint32_t)nl;
    prm.total = (uint32_t)tl;
    prm.rep_stride = n;
    prm.key = key;
    prm.pid_base = pid_base + sh.lo;
    prm.max_plies = -1;
    prm.leaf_wins = (unsigned int *)d.d_plies.ptr;
    prm.counters = (unsigned long long *)d.d_misc.ptr;
    const cudaError_t e = launch_playouts(ctx, d, prm, km, sched, d.stream);
    if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("playout launch: ") + cudaGetErrorString(e));
    B2P_CUDA(ctx, cudaMemcpyAsync(wins_out + 2 * sh.lo, d.d_plies.ptr, nl * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, d.stream));
    B2P_CUDA(ctx, cudaMemcpyAsync(d.h_misc.ptr, d.d_misc.ptr, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream));
  }
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    B2P_CUDA(ctx, cudaStreamSynchronize(d.stream));
    if (counters_out)
      for (int k = 0; k < 4; k++) counters_out[k] += ((const uint64_t *)d.h_misc.ptr)[k];
  }
  return B2P_OK;
}

// Asynchronous form of b2p_run_counts for a pipelined caller (b2p_tree_search_ex): queues H2D copy, kernel and D2H
// copy of the per-leaf counts on pipeline slot `slot` of every device the batch is sharded over and returns at once.
static int run_counts_async_impl(b2p_ctx *ctx, int slot, const b2p_state16 *states, size_t n, uint32_t reps, uint64_t key,
                         uint64_t pid_base, int mode, int sched, int order, uint32_t *wins_out) {
  if (!ctx) return B2P_EINVAL;
  if (slot < 0 || slot >= kSlots) return fail(ctx, B2P_EINVAL, "bad pipeline slot");
  if (ctx->slot_span[slot] != 0) return fail(ctx, B2P_EINVAL, "pipeline slot still in flight: call b2p_wait_slot first");
  if (n == 0 || reps == 0) return B2P_OK;
  if (!states || !wins_out) return fail(ctx, B2P_EINVAL, "NULL buffer");
  KernelMode km;
  if (!mode_to_kernel(mode, order, &km)) return fail(ctx, B2P_EINVAL, "unknown mode/order");
  const int G = device_span(ctx, n, n * (size_t)reps);
  for (int g = 0; g < G; g++)
    if ((unsigned long long)(shard_of(n, g, G).hi - shard_of(n, g, G).lo) * reps >= (1ull << 31))
      return fail(ctx, B2P_EINVAL, "per-device n*reps must be < 2^31");
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    Slot &sl = d.slots[slot];
    cudaStream_t st = d.aux[slot] ? d.aux[slot] : d.stream;
    const Shard sh = shard_of(n, g, G);
    const size_t nl = sh.hi - sh.lo;
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    int rc;
    if ((rc = ensure(ctx, sl.d_states, nl * sizeof(b2p_state16), false))) return rc;
    if ((rc = ensure(ctx, sl.d_counts, nl * 2 * sizeof(uint32_t), false))) return rc;
    if ((rc = ensure(ctx, sl.d_misc, 4 * sizeof(uint64_t), false))) return rc;
    if ((rc = ensure(ctx, sl.h_misc, 4 * sizeof(uint64_t), true))) return rc;
    if (!sl.ev0) B2P_CUDA(ctx, cudaEventCreate(&sl.ev0));
    if (!sl.ev1) B2P_CUDA(ctx, cudaEventCreate(&sl.ev1));
    B2P_CUDA(ctx, cudaMemcpyAsync(sl.d_states.ptr, states + sh.lo, nl * sizeof(b2p_state16), cudaMemcpyHostToDevice, st));
    B2P_CUDA(ctx, cudaMemsetAsync(sl.d_misc.ptr, 0, 4 * sizeof(uint64_t), st));
    B2P_CUDA(ctx, cudaMemsetAsync(sl.d_counts.ptr, 0, nl * 2 * sizeof(uint32_t), st));
    PlayoutParams prm;
    std::memset(&prm, 0, sizeof prm);
    prm.states = reinterpret_cast<const uint4 *>(sl.d_states.ptr);
    prm.n = (uint32_t)nl;
    prm.total = (uint32_t)(nl * reps);
    prm.rep_stride = n;
    prm.key = key;
    prm.pid_base = pid_base + sh.lo;
    prm.max_plies = -1;
    prm.leaf_wins = (unsigned int *)sl.d_counts.ptr;
    prm.counters = (unsigned long long *)sl.d_misc.ptr;
    B2P_CUDA(ctx, cudaEventRecord(sl.ev0, st));
    const cudaError_t e = launch_playouts(ctx, d, prm, km, sched, st);
    if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("playout launch: ") + cudaGetErrorString(e));
    B2P_CUDA(ctx, cudaEventRecord(sl.ev1, st));
    B2P_CUDA(ctx, cudaMemcpyAsync(wins_out + 2 * sh.lo, sl.d_counts.ptr, nl * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    B2P_CUDA(ctx, cudaMemcpyAsync(sl.h_misc.ptr, sl.d_misc.ptr, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    ctx->slot_span[slot] = g + 1;
  }
  return B2P_OK;
}

int b2p_run_counts_async(b2p_ctx *ctx, int slot, const b2p_state16 *states, size_t n, uint32_t reps, uint64_t key,
                         uint64_t pid_base, int mode, int sched, int order, uint32_t *wins_out) {
  if (ctx && slot >= 0 && slot < kSlots && ctx->slot_span[slot] != 0)
    return fail(ctx, B2P_EINVAL, "pipeline slot still in flight: call b2p_wait_slot first");
  const int rc = run_counts_async_impl(ctx, slot, states, n, reps, key, pid_base, mode, sched, order, wins_out);
  if (rc != B2P_OK && ctx && slot >= 0 && slot < kSlots && ctx->slot_span[slot] != 0) {
    // failed half-way through the devices: nothing of this batch may stay in flight, and the slot must be reusable
    const std::string why = ctx->err;
    for (int g = 0; g < ctx->slot_span[slot]; g++) {
      Device &d = ctx->devs[g];
      if (cudaSetDevice(d.id) == cudaSuccess) cudaStreamSynchronize(d.aux[slot] ? d.aux[slot] : d.stream);
    }
    ctx->slot_span[slot] = 0;
    ctx->err = why;
  }
  return rc;
}

// Context-owned page-locked staging for a pipelined caller: room for `leaves` packed states and 2*`leaves` win
// counts on `slot`.  Grow-only and kept for the life of the context, so a caller that creates many short-lived
// trees (a tournament, a benchmark) pays for page-locking once.  The slot must be idle.
int b2p_slot_staging(b2p_ctx *ctx, int slot, size_t leaves, b2p_state16 **leaves_out, uint32_t **wins_out) {
  if (!ctx) return B2P_EINVAL;
  if (slot < 0 || slot >= kSlots || !leaves_out || !wins_out) return fail(ctx, B2P_EINVAL, "bad staging arguments");
  if (ctx->slot_span[slot] != 0) return fail(ctx, B2P_EINVAL, "pipeline slot still in flight: call b2p_wait_slot first");
  int rc;
  if ((rc = ensure(ctx, ctx->slot_leaves[slot], leaves * sizeof(b2p_state16), true))) return rc;
  if ((rc = ensure(ctx, ctx->slot_wins[slot], leaves * 2 * sizeof(uint32_t), true))) return rc;
  // the slot's device buffers too: growing them later means cudaFree, which synchronises the whole device
  for (Device &d : ctx->devs) {
    Slot &sl = d.slots[slot];
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    if ((rc = ensure(ctx, sl.d_states, leaves * sizeof(b2p_state16), false))) return rc;
    if ((rc = ensure(ctx, sl.d_counts, leaves * 2 * sizeof(uint32_t), false))) return rc;
  }
  *leaves_out = (b2p_state16 *)ctx->slot_leaves[slot].ptr;
  *wins_out = (uint32_t *)ctx->slot_wins[slot].ptr;
  return B2P_OK;
}

int b2p_wait_slot(b2p_ctx *ctx, int slot, uint64_t counters_out[4], float *kernel_ms_out) {
  if (!ctx) return B2P_EINVAL;
  if (slot < 0 || slot >= kSlots) return fail(ctx, B2P_EINVAL, "bad pipeline slot");
  if (counters_out) counters_out[0] = counters_out[1] = counters_out[2] = counters_out[3] = 0;
  if (kernel_ms_out) *kernel_ms_out = 0.f;
  const int G = ctx->slot_span[slot];
  ctx->slot_span[slot] = 0;
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    Slot &sl = d.slots[slot];
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    B2P_CUDA(ctx, cudaStreamSynchronize(d.aux[slot] ? d.aux[slot] : d.stream));
    if (counters_out)
      for (int k = 0; k < 4; k++) counters_out[k] += ((const uint64_t *)sl.h_misc.ptr)[k];
    if (kernel_ms_out) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sl.ev0, sl.ev1) == cudaSuccess && ms > *kernel_ms_out) *kernel_ms_out = ms;
    }
  }
  return B2P_OK;
}

int b2p_run_states776(b2p_ctx *ctx, const void *states, size_t n, int mode, int sched, int32_t *winners_out) {
  if (!ctx) return B2P_EINVAL;
  if (n == 0) return B2P_OK;  // reference: empty vector in, empty vector out (src/singlePlayout.cu:73-75)
  if (!states || !winners_out) return fail(ctx, B2P_EINVAL, "NULL buffer");
  if (mode != B2P_MODE_RANDOM && mode != B2P_MODE_HEURISTIC) return fail(ctx, B2P_EINVAL, "unknown mode");
  const uint64_t key = ctx->seed + 0x9E3779B97F4A7C15ull * ctx->calls;
  ctx->calls++;
  const int G = device_span(ctx, n, n);
  const unsigned char *src = (const unsigned char *)states;
  const KernelMode km = mode == B2P_MODE_HEURISTIC ? kHeuristic : kRandomFast;

  // Two granularities.  PACK tasks (kPackGrain states) are what the host
  // workers pull: even a 2 k-leaf MCTS batch is packed by several cores.  LAUNCH segments (kSegment leaves)
  // are what the GPU sees: the worker that packs the last task of a segment queues the segment's H2D copy,
  // kernel and D2H copy on one of the device's streams, so the GPU work of segment s hides behind the
  // packing of segment s+1 and no launch is smaller than it has to be.  Packing goes 776 B -> 16 B straight
  // into pinned staging (48x less PCIe traffic than the reference's raw State copy).  Playout ids are
  // global leaf indices, so neither granularity changes any result.
  // a quarter of the batch per launch, within [8192, 32768]: mid-size batches (the MCTS range 2 k - 65 k) still
  // overlap packing with GPU work, large ones do not pay for more launches than they need
  const size_t kSegment = std::min(kSegmentMax, std::max(kSegmentMin, (n / 4 + kPackGrain - 1) / kPackGrain * kPackGrain));
  struct Segment { int g; size_t lo, len; std::atomic<int> todo; };
  struct Task { int seg; size_t lo, len; };
  std::vector<std::unique_ptr<Segment>> segs;
  std::vector<Task> tasks;
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    const Shard sh = shard_of(n, g, G);
    const size_t nl = sh.hi - sh.lo;
    if (nl >= (1ull << 31)) return fail(ctx, B2P_EINVAL, "per-device n must be < 2^31");
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    int rc;
    if ((rc = ensure(ctx, d.h_states, nl * sizeof(b2p_state16), true))) return rc;
    if ((rc = ensure(ctx, d.d_states, nl * sizeof(b2p_state16), false))) return rc;
    if ((rc = ensure(ctx, d.d_winners, nl, false))) return rc;
    if ((rc = ensure(ctx, d.h_winners, nl, true))) return rc;
    // the tail is merged into the last segment when it is short (no 100-leaf launches)
    for (size_t lo = 0; lo < nl;) {
      size_t len = std::min(kSegment, nl - lo);
      if (nl - lo - len < kSegment / 4) len = nl - lo;
      auto sg = std::make_unique<Segment>();
      sg->g = g; sg->lo = lo; sg->len = len;
      sg->todo.store((int)((len + kPackGrain - 1) / kPackGrain));
      segs.push_back(std::move(sg));
      lo += len;
    }
  }
  // interleave devices so that every GPU gets work early
  std::stable_sort(segs.begin(), segs.end(), [](const std::unique_ptr<Segment> &a, const std::unique_ptr<Segment> &b) { return a->lo < b->lo; });
  for (size_t s = 0; s < segs.size(); s++)
    for (size_t lo = 0; lo < segs[s]->len; lo += kPackGrain) tasks.push_back({(int)s, lo, std::min(kPackGrain, segs[s]->len - lo)});

  std::atomic<size_t> next_task{0};
  std::mutex enqueue_mu;
  std::string err;
  auto worker = [&]() {
    for (;;) {
      const size_t t = next_task.fetch_add(1);
      if (t >= tasks.size()) return;
      const Task tk = tasks[t];
      Segment &sg = *segs[tk.seg];
      Device &d = ctx->devs[sg.g];
      const Shard sh = shard_of(n, sg.g, G);
      b2p_state16 *stage = (b2p_state16 *)d.h_states.ptr + sg.lo + tk.lo;
      const unsigned char *from = src + kStateBytes * (sh.lo + sg.lo + tk.lo);
      pack776_range(from, tk.len, stage);
      if (sg.todo.fetch_sub(1) != 1) continue;
      // last task of the segment: queue its GPU work
      std::lock_guard<std::mutex> lock(enqueue_mu);
      if (!err.empty()) continue;
      const unsigned a = d.next_aux++ % 4;
      cudaStream_t st = d.aux[a] ? d.aux[a] : d.stream;
      cudaError_t e = cudaSetDevice(d.id);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync((b2p_state16 *)d.d_states.ptr + sg.lo, (b2p_state16 *)d.h_states.ptr + sg.lo, sg.len * sizeof(b2p_state16),
                            cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) {
        PlayoutParams prm;
        std::memset(&prm, 0, sizeof prm);
        prm.states = reinterpret_cast<const uint4 *>((b2p_state16 *)d.d_states.ptr + sg.lo);
        prm.n = (uint32_t)sg.len;
        prm.total = (uint32_t)sg.len;
        prm.rep_stride = n;
        prm.key = key;
        prm.pid_base = sh.lo + sg.lo;
        prm.max_plies = -1;
        prm.winners = (int8_t *)d.d_winners.ptr + sg.lo;
        e = launch_playouts(ctx, d, prm, km, sched, st);  // AUTO decides on the size of THIS launch
      }
      if (e == cudaSuccess)
        e = cudaMemcpyAsync((int8_t *)d.h_winners.ptr + sg.lo, (int8_t *)d.d_winners.ptr + sg.lo, sg.len, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) err = std::string("b2p_run_states776 pipeline: ") + cudaGetErrorString(e);
    }
  };
  // at most 32 packers per context: the loop is bound by host memory bandwidth, and one process per GPU may run it
  ctx->pool.run(std::min<size_t>(tasks.size(), 32), worker);
  // wait for every stream that may carry work of this call -- also on the error path: the next call may
  // grow (free + reallocate) buffers that kernels still in flight are writing
  cudaError_t sync_err = cudaSuccess;
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    cudaError_t e = cudaSetDevice(d.id);
    for (cudaStream_t a : d.aux)
      if (a && e == cudaSuccess) e = cudaStreamSynchronize(a);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    if (e != cudaSuccess && sync_err == cudaSuccess) sync_err = e;
  }
  if (!err.empty()) return fail(ctx, B2P_ECUDA, err);
  if (sync_err != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("b2p_run_states776: ") + cudaGetErrorString(sync_err));
  // PlayerId is a 4-byte enum: widen the int8 winners into the caller's array
  std::atomic<size_t> next_block{0};
  const size_t kWiden = 1 << 16, blocks = (n + kWiden - 1) / kWiden;
  auto widen = [&]() {
    for (;;) {
      const size_t b = next_block.fetch_add(1);
      if (b >= blocks) return;
      const size_t lo = b * kWiden, hi = std::min(n, lo + kWiden);
      int g = 0;
      for (size_t i = lo; i < hi;) {
        while (shard_of(n, g, G).hi <= i) g++;
        const Shard sh = shard_of(n, g, G);
        const int8_t *w = (const int8_t *)ctx->devs[g].h_winners.ptr - sh.lo;
        const size_t stop = std::min(hi, sh.hi);
        for (; i < stop; i++) winners_out[i] = (int32_t)w[i];
      }
    }
  };
  ctx->pool.run(std::min<size_t>(blocks, 32), widen);
  return B2P_OK;
}

int b2p_genmoves(b2p_ctx *ctx, const b2p_state16 *states, size_t n, int max_moves, b2p_move_t *moves_out,
                 uint8_t *counts_out) {
  if (!ctx) return B2P_EINVAL;
  if (n == 0) return B2P_OK;
  if (!states || !moves_out || !counts_out || max_moves <= 0) return fail(ctx, B2P_EINVAL, "bad genmoves arguments");
  const int G = device_span(ctx, n, n);
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    const Shard sh = shard_of(n, g, G);
    const size_t nl = sh.hi - sh.lo;
    if (nl >= (1ull << 31)) return fail(ctx, B2P_EINVAL, "per-device n must be < 2^31");
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    int rc;
    if ((rc = ensure(ctx, d.d_states, nl * sizeof(b2p_state16), false))) return rc;
    if ((rc = ensure(ctx, d.d_moves, nl * (size_t)max_moves * sizeof(b2p_move_t), false))) return rc;
    if ((rc = ensure(ctx, d.d_counts, nl, false))) return rc;
    B2P_CUDA(ctx, cudaMemcpyAsync(d.d_states.ptr, states + sh.lo, nl * sizeof(b2p_state16), cudaMemcpyHostToDevice, d.stream));
    B2P_CUDA(ctx, cudaMemsetAsync(d.d_moves.ptr, 0, nl * (size_t)max_moves * sizeof(b2p_move_t), d.stream));
    cudaError_t e = launch_genmoves((const uint4 *)d.d_states.ptr, (uint32_t)nl, max_moves, (unsigned long long *)d.d_moves.ptr,
                                    (uint8_t *)d.d_counts.ptr, d.stream);
    if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("genmoves launch: ") + cudaGetErrorString(e));
    ctx->launches++;
    B2P_CUDA(ctx, cudaMemcpyAsync(moves_out + sh.lo * (size_t)max_moves, d.d_moves.ptr, nl * (size_t)max_moves * sizeof(b2p_move_t), cudaMemcpyDeviceToHost, d.stream));
    B2P_CUDA(ctx, cudaMemcpyAsync(counts_out + sh.lo, d.d_counts.ptr, nl, cudaMemcpyDeviceToHost, d.stream));
  }
  return b2p_sync(ctx);
}

int b2p_gen_leaves(b2p_ctx *ctx, size_t n, uint64_t key, uint64_t first_index, b2p_state16 *out) {
  if (!ctx) return B2P_EINVAL;
  if (n == 0) return B2P_OK;
  if (!out) return fail(ctx, B2P_EINVAL, "out is NULL");
  const int G = device_span(ctx, n, n);
  for (int g = 0; g < G; g++) {
    Device &d = ctx->devs[g];
    const Shard sh = shard_of(n, g, G);
    const size_t nl = sh.hi - sh.lo;
    B2P_CUDA(ctx, cudaSetDevice(d.id));
    int rc;
    if ((rc = ensure(ctx, d.d_final, nl * sizeof(b2p_state16), false))) return rc;
    if ((rc = b2p_gen_leaves_device(ctx, g, nl, key, first_index + sh.lo, (b2p_state16 *)d.d_final.ptr, (void *)d.stream))) return rc;
    B2P_CUDA(ctx, cudaMemcpyAsync(out + sh.lo, d.d_final.ptr, nl * sizeof(b2p_state16), cudaMemcpyDeviceToHost, d.stream));
  }
  return b2p_sync(ctx);
}

int b2p_microbench(b2p_ctx *ctx, int dev_index, int which, int iters, double *thread_ops_per_s, double *ms_out) {
  if (!ctx) return B2P_EINVAL;
  if (dev_index < 0 || dev_index >= (int)ctx->devs.size() || iters <= 0) return fail(ctx, B2P_EINVAL, "bad microbench arguments");
  Device &d = ctx->devs[dev_index];
  B2P_CUDA(ctx, cudaSetDevice(d.id));
  int rc;
  if ((rc = ensure(ctx, d.d_misc, 4 * sizeof(uint64_t), false))) return rc;
  double ops = 0;
  cudaError_t e = launch_microbench(which, std::max(1, iters / 8), d.sm_count, (uint32_t *)d.d_misc.ptr, d.stream, &ops);  // warm-up
  if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("microbench launch: ") + cudaGetErrorString(e));
  B2P_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
  e = launch_microbench(which, iters, d.sm_count, (uint32_t *)d.d_misc.ptr, d.stream, &ops);
  if (e != cudaSuccess) return fail(ctx, B2P_ECUDA, std::string("microbench launch: ") + cudaGetErrorString(e));
  B2P_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
  B2P_CUDA(ctx, cudaEventSynchronize(d.ev1));
  float ms = 0;
  B2P_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
  ctx->launches += 2;
  if (thread_ops_per_s) *thread_ops_per_s = ops / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  return B2P_OK;
}

}  // extern "C"
