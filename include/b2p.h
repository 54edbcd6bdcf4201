/*
 * include/b2p.h -- C ABI of the B200-native checkers playout engine (libb2p.so).
 *
 * This is the drop-in boundary for ONE hot path of krame505/gpu_ai: batched checkers playouts
 * from MCTS leaf states behind `PlayoutDriver::runPlayouts(std::vector<State>)`
 * (reference: src/playout.hpp:27-33).  Every entry point names the reference interface it
 * replaces.  Plain pointers and sizes only; no C++ or torch types cross this line.  The
 * reference-side binding (a C++ shim TU that defines the reference's five device symbols on
 * top of these calls) is shim/playout_shim.cpp; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative B2P_E* code; nothing throws or exits
 *     across the ABI (the reference prints and exit(1)s on CUDA errors,
 *     src/singlePlayout.cu:91-108).  b2p_last_error() returns the message.
 *   - a context owns its devices' streams, staging and device buffers (grow-only, reused
 *     across calls; the reference mallocs and frees on every call, src/singlePlayout.cu:80-116).
 *   - one context = one caller at a time; different contexts may be used concurrently from
 *     different host threads (MCTSPlayer worker threads, src/player.cpp:119-150).
 *   - the device-resident calls may be issued on any caller stream and any number of them may be in
 *     flight: each launch takes a work-queue head from a 256-slot ring whose slots are guarded by events
 *     (a wrapped slot waits for the launch that used it last).
 *   - host-buffer calls use a second device only when every shard keeps >= 8192 playouts.
 *   - there is NO CPU execution path: without a usable CUDA device b2p_create fails.
 *
 * Packed state (16 bytes): square i = row*4 + col/2 over the 32 dark squares (the numbering of
 * the reference's genTypeMovesParallel, src/state.cu:185-188).
 */
#ifndef B2P_H
#define B2P_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2p_ctx b2p_ctx;

/* replaces struct State (776 B AoS, src/state.hpp:119-122) on the device and on the wire */
typedef struct b2p_state16 {
  uint32_t p1;    /* PLAYER_1 pieces */
  uint32_t p2;    /* PLAYER_2 pieces */
  uint32_t kings; /* kings of both players */
  uint32_t meta;  /* bit 0: turn (0 = PLAYER_1, 1 = PLAYER_2); bits 8..31: movesSinceLastCapture */
} b2p_state16;

/* replaces struct Move (38 B, src/state.hpp:253-296) in b2p_genmoves output:
 * [0:5) from, [5:10) to, [10:13) jumps, [13] promoted, [16+5k : 21+5k) landing square of hop k.
 * Move::removed[k] is the midpoint of consecutive landings (Move::addJump, src/state.cu:456-462). */
typedef uint64_t b2p_move_t;

/* winners use the reference's PlayerId values (src/state.hpp:21-26) */
#define B2P_PLAYER_1 0
#define B2P_PLAYER_2 1
#define B2P_PLAYER_NONE (-1)
#define B2P_UNFINISHED 2 /* only with max_plies >= 0 */

enum { B2P_MODE_RANDOM = 0,     /* HostPlayoutDriver semantics, src/playout.cpp:17-32 */
       B2P_MODE_HEURISTIC = 1   /* HostHeuristicPlayoutDriver semantics, src/heuristicPlayout.cpp:12-48 */ };
enum { B2P_SCHED_THREAD = 0,    /* one lane per playout, persistent lanes with warp-aggregated refill
                                   (replaces singlePlayoutKernel / coarsePlayoutKernel scheduling) */
       B2P_SCHED_WARP = 1,      /* one warp per playout (replaces playoutKernel / heuristicPlayoutKernel
                                   scheduling); lowest latency for small batches */
       B2P_SCHED_AUTO = 2 };
enum { B2P_ORDER_CANONICAL = 0, /* rank j -> j-th move of State::getMoves() */
       B2P_ORDER_FAST = 1       /* rank j -> j-th move in direction-major order (same uniform law) */ };

#define B2P_OK 0
#define B2P_EINVAL (-1)
#define B2P_ECUDA (-2)
#define B2P_ENOMEM (-3)
#define B2P_ENODEV (-4)

typedef struct b2p_devinfo {
  int device_id;
  int sm_count;
  int clock_khz;
  int cc_major, cc_minor;
  size_t total_mem;
  char name[128];
} b2p_devinfo;

/* ---- lifetime -------------------------------------------------------------------------- */
/* device_ids == NULL: use devices 0..n_dev-1; n_dev <= 0: all visible devices. */
int b2p_create(b2p_ctx **out, const int *device_ids, int n_dev, uint64_t seed);
void b2p_destroy(b2p_ctx *ctx);
const char *b2p_last_error(const b2p_ctx *ctx); /* ctx may be NULL: error of the last failed b2p_create */
int b2p_device_count(const b2p_ctx *ctx);
int b2p_device_info(const b2p_ctx *ctx, int dev_index, b2p_devinfo *out);
const char *b2p_version(void);

/* ---- the reference-facing call ---------------------------------------------------------------
 * Replaces Device{Single,Multiple,Coarse,Heuristic}PlayoutDriver::runPlayouts
 * (src/singlePlayout.cu:71-120, multiplePlayout.cu:53-98, coarsePlayout.cu:91-163,
 * heuristicPlayout.cu:102-147).  `states` = n reference `State` objects (776 B each) in host
 * memory; winners_out[i] = PlayerId of a playout from states[i].  Packs on the host, shards
 * contiguously over the context's devices, plays, gathers.  n == 0 is a no-op.  The RNG stream
 * is f(context seed, per-context call counter, global leaf index): repeated calls differ (the
 * reference reuses SEED 12345 every call) but a context replays identically. */
int b2p_run_states776(b2p_ctx *ctx, const void *states, size_t n, int mode, int sched, int32_t *winners_out);

/* ---- packed host-buffer call ---------------------------------------------------------------
 * Playout id of (rep, leaf) = pid_base + rep*n + leaf; it is the Philox counter, so results do
 * not depend on the number of devices.  Outputs (each may be NULL) have n*reps entries in that
 * order; counters_out = {draws, PLAYER_1 wins, PLAYER_2 wins, plies played} summed over all
 * devices.  max_plies < 0: play to the end. */
int b2p_run_packed(b2p_ctx *ctx, const b2p_state16 *states, size_t n, uint32_t reps, uint64_t key, uint64_t pid_base,
                   int mode, int sched, int order, int max_plies, int8_t *winners_out, uint32_t *plies_out,
                   b2p_state16 *final_out, uint64_t counters_out[4]);

/* K playouts per leaf with win COUNTS per leaf returned: wins_out[2*i + p] = number of the
 * `reps` playouts from states[i] won by PLAYER_(p+1); draws = reps - both.  Same playout ids and rules as
 * b2p_run_packed; only 8 bytes per leaf come back. */
int b2p_run_counts(b2p_ctx *ctx, const b2p_state16 *states, size_t n, uint32_t reps, uint64_t key, uint64_t pid_base,
                   int mode, int sched, int order, uint32_t *wins_out, uint64_t counters_out[4]);

/* Asynchronous, pipelined form of b2p_run_counts (the caller side of SURVEY.md 8f-1/8f-2: select of batch k+1 overlaps
 * the playouts of batch k).  `slot` in [0, 4): each slot has its own stream and device buffers on every device, so up
 * to four batches may be in flight.  b2p_run_counts_async queues H2D copy + kernel + D2H copy and returns at once;
 * `states` and `wins_out` must stay valid (and should be page-locked: b2p_alloc_host) until b2p_wait_slot(slot)
 * returns.  counters_out / kernel_ms_out may be NULL; kernel_ms = device time of the slot's kernel (max over devices). */
int b2p_run_counts_async(b2p_ctx *ctx, int slot, const b2p_state16 *states, size_t n, uint32_t reps, uint64_t key,
                         uint64_t pid_base, int mode, int sched, int order, uint32_t *wins_out);
int b2p_wait_slot(b2p_ctx *ctx, int slot, uint64_t counters_out[4], float *kernel_ms_out);
/* context-owned page-locked staging for `slot`: room for `leaves` states and 2*leaves counts; grow-only, kept for the
 * life of the context (callers that build many short-lived trees pay for page-locking once); the slot must be idle */
int b2p_slot_staging(b2p_ctx *ctx, int slot, size_t leaves, b2p_state16 **leaves_out, uint32_t **wins_out);

/* ---- move generation (replaces genMovesKernel / genMovesTest, src/genMovesTest.cu:10-100) ----
 * moves_out[i*max_moves + k] = k-th move of State::getMoves() for states[i] (k < max_moves);
 * counts_out[i] = number of legal moves (may exceed max_moves). */
int b2p_genmoves(b2p_ctx *ctx, const b2p_state16 *states, size_t n, int max_moves, b2p_move_t *moves_out,
                 uint8_t *counts_out);

/* ---- device-resident calls (single device `dev_index` of the context, asynchronous on
 * `cuda_stream` (a cudaStream_t, passed through as is: NULL = the CUDA default stream).  All
 * pointers are device pointers on that device.  counters (4 x u64) are ACCUMULATED with atomics: zero them first. */
int b2p_run_packed_device(b2p_ctx *ctx, int dev_index, const b2p_state16 *d_states, size_t n, uint32_t reps,
                          uint64_t key, uint64_t pid_base, int mode, int sched, int order, int max_plies,
                          int8_t *d_winners, uint32_t *d_plies, b2p_state16 *d_final, uint64_t *d_counters,
                          void *cuda_stream);
int b2p_genmoves_device(b2p_ctx *ctx, int dev_index, const b2p_state16 *d_states, size_t n, int max_moves,
                        b2p_move_t *d_moves, uint8_t *d_counts, void *cuda_stream);
/* D_ref leaf set: the reference's genRandomStates recipe (src/driver.cpp:76-104) with a
 * reproducible per-leaf Philox stream; leaf j = first_index + j. */
int b2p_gen_leaves_device(b2p_ctx *ctx, int dev_index, size_t n, uint64_t key, uint64_t first_index,
                          b2p_state16 *d_out, void *cuda_stream);
int b2p_gen_leaves(b2p_ctx *ctx, size_t n, uint64_t key, uint64_t first_index, b2p_state16 *out);
int b2p_sync(b2p_ctx *ctx);

/* ---- caller-visible pinned host memory (the zero-copy side of the boundary) ----------------
 * The reference hands `std::vector<State>` storage (pageable) to cudaMemcpy on every call
 * (src/singlePlayout.cu:83-91).  Buffers obtained here are page-locked and portable across the context's
 * devices: b2p_run_packed / b2p_run_counts / b2p_genmoves then copy to and from them with true asynchronous
 * DMA instead of the driver's pageable staging path.  Any host pointer is still accepted everywhere. */
int b2p_alloc_host(void **out, size_t bytes);
int b2p_free_host(void *ptr);

/* ---- layout converters (host, multi-threaded; no device involved) ------------------------------
 * struct State <-> b2p_state16.  type/owner are read only where `occupied` is set: State::move
 * leaves stale fields in vacated squares (src/state.cu:78-84). */
int b2p_pack776(const void *states, size_t n, b2p_state16 *out);
/* which packer this process runs: "avx512bw+bmi2" (one VPTESTMB + three PEXT per cache line of the State) or
 * "scalar"; B2P_PACK_SCALAR=1 in the environment forces the portable one.  Same bits either way. */
const char *b2p_pack776_impl(void);
int b2p_unpack776(const b2p_state16 *states, size_t n, void *states_out);
/* b2p_move_t -> struct Move (38 B: from@0 to@2 removed@4 intermediate@20 jumps@36 promoted@37) */
int b2p_expand_move(b2p_move_t move, void *move38_out);

/* ---- measurement helpers ------------------------------------------------------------------------
 * Dependency-light integer-pipe microbenchmarks that freeze the INT32 roofline denominator
 * (SURVEY.md 8d).  which: 0 LOP3, 1 IADD3, 2 SHF, 3 POPC, 4 IMAD, 5 LOP3+IMAD 1:1, 6 BREV, 7 FLO, 8 LOP3+IMAD 3:1,
 * 9 IMAD.HI, 10 LOP3 + mul.hi-as-shift, 11 LOP3 + SHF.R.
 * Returns thread-level ops per second over the whole chip. */
int b2p_microbench(b2p_ctx *ctx, int dev_index, int which, int iters, double *thread_ops_per_s, double *ms);
/* number of kernels this context has launched so far (bench.py reports it as gpu_launches) */
uint64_t b2p_launch_count(const b2p_ctx *ctx);

/* ---- search tree: the caller side of the path ----------------------------------------
 * Replaces GameTree (src/mcts.hpp:12-64, src/mcts.cpp:11-191) decision for decision -- fed the same playout
 * results it selects the same leaves in the same order -- on 16-byte packed states, writing leaves straight
 * into the caller's buffer (no per-level vector<State> concatenation, src/mcts.cpp:144-156).  Host-side, like
 * the reference's tree; no CUDA context needed except for b2p_tree_search. */
typedef struct b2p_tree b2p_tree;
typedef struct b2p_tree_stats {
  uint64_t nodes;
  uint64_t total_trials; /* GameTree::getTotalTrials of the root */
  uint64_t wins_p1, wins_p2;
  uint32_t root_children, root_moves;
  b2p_state16 root_state;
} b2p_tree_stats;

int b2p_tree_create(b2p_tree **out, const b2p_state16 *root);            /* GameTree::GameTree(State) */
void b2p_tree_destroy(b2p_tree *tree);
/* GameTree::select (src/mcts.cpp:63-157): leaves_out needs room for `trials` states; *n_out = leaves written */
int b2p_tree_select(b2p_tree *tree, uint32_t trials, b2p_state16 *leaves_out, uint32_t *n_out);
/* GameTree::update (src/mcts.cpp:159-180).  winners: PlayerId per trial of the last select, n = its leaf count;
 * reps > 1: `reps` playouts per selected leaf laid out [rep][leaf] exactly as b2p_run_packed returns them */
int b2p_tree_update(b2p_tree *tree, const int8_t *winners, uint32_t n, uint32_t reps);
/* the same with per-leaf win counts (b2p_run_counts layout): every selected leaf was played `reps` times */
int b2p_tree_update_counts(b2p_tree *tree, const uint32_t *wins, uint32_t n, uint32_t reps);
int b2p_tree_best_move(const b2p_tree *tree, int player, b2p_move_t *move_out);   /* GameTree::getOptMove */
/* the root move with the most trials (the "robust child"): the move rule that goes with B2P_POLICY_UCT */
int b2p_tree_robust_move(const b2p_tree *tree, int player, b2p_move_t *move_out);
int b2p_tree_move(b2p_tree *tree, b2p_move_t move);                      /* GameTree::move (subtree reuse) */
int b2p_tree_info(const b2p_tree *tree, b2p_tree_stats *out);
/* root move list with per-child statistics; returns the number of legal root moves */
int b2p_tree_root_moves(const b2p_tree *tree, b2p_move_t *moves_out, uint64_t *trials_out, uint64_t *wins_p1_out,
                        uint64_t *wins_p2_out, uint32_t capacity);
const char *b2p_tree_last_error(const b2p_tree *tree);
/* MCTSPlayer::worker (src/player.cpp:134-150) fused with the playout engine: select -> playouts -> update,
 * batch = max(initial_batch, scale * leaf selections so far), `reps` playouts per selected leaf, until `iterations`
 * rounds or `seconds` of wall clock (0 = unlimited on that axis; both 0 = nothing). */
int b2p_tree_search(b2p_ctx *ctx, b2p_tree *tree, uint32_t iterations, double seconds, uint32_t initial_batch,
                    float scale, uint32_t reps, int mode, uint64_t key, uint64_t *playouts_out);

/* The same loop with its knobs and its accounting exposed.  The loop is a pipeline: leaf selection runs on
 * `threads` host threads (disjoint subtrees), batches travel through page-locked staging to b2p_run_counts_async,
 * and with depth >= 2 the selection + statistics update of one batch overlap the playouts of the previous one
 * (in-flight trials count as visits without wins).  depth == 1 is the strictly serial loop: it takes exactly the
 * decisions of b2p_tree_select -> b2p_run_counts -> b2p_tree_update_counts, for any `threads` and any number of
 * devices.  Results are deterministic for given options (the time limit only decides how many rounds run). */
enum { B2P_POLICY_REFERENCE = 0, /* GameTree::select's rule: a node's trials are shared out in proportion to the children's
                                    UCB1 values (src/mcts.cpp:93-139) */
       B2P_POLICY_UCT = 1        /* a node's trials go down one by one to the child with the highest UCB1 value, in-flight
                                    trials counting as visits without wins: what a sequential UCT search does */ };
typedef struct b2p_search_opts {
  uint32_t iterations;    /* rounds; 0 = until `seconds` */
  double seconds;         /* wall-clock budget; 0 = until `iterations` */
  uint32_t initial_batch; /* leaves per round, lower bound */
  float scale;            /* leaves per round = max(initial_batch, scale * leaf selections so far) ... */
  uint32_t max_batch;     /* ... capped here (0 = 2^20) and by 2^31 / reps */
  uint32_t reps;          /* playouts per selected leaf */
  int mode;               /* B2P_MODE_* */
  uint64_t key;           /* Philox key of round r = key + r; playout ids count up over the whole search */
  int threads;            /* host threads for select/update; 0 = min(hardware threads, 64) */
  int depth;              /* batches in flight: 1 serial, 2..4 pipelined; 0 = 2 */
  int policy;             /* B2P_POLICY_*: how a node hands its trials to its children */
} b2p_search_opts;
typedef struct b2p_search_stats {
  uint64_t playouts, leaves, batches, nodes;
  double seconds;                     /* wall clock of the whole call */
  double select_s, update_s, wait_s;  /* host time: selecting, updating, blocked on the GPU */
  double kernel_s;                    /* device time of the playout kernels (max over devices per batch) */
  uint32_t threads, depth;
} b2p_search_stats;
/* The two host halves of one pipelined round, for a caller that runs the playouts itself.  select_batch picks
 * `trials` leaves on `threads` host threads (same decisions as b2p_tree_select), counts them into the visited nodes
 * as `reps` trials each and remembers the batch in `slot` (0..3); policy 0 = the reference's rule with the reference's
 * expression types (bit-identical decisions, what depth 1 uses), 1 = the same rule in single precision (what
 * depth >= 2 uses), 2 = B2P_POLICY_UCT; update_batch folds the batch's per-leaf win counts
 * (b2p_run_counts layout) into the nodes it visited.  Several slots may be selected before the first is updated;
 * b2p_tree_move refuses to re-root while a selected batch has not been folded in. */
int b2p_tree_select_batch(b2p_tree *tree, int slot, uint32_t trials, uint32_t reps, int threads, int policy,
                          b2p_state16 *leaves_out);
int b2p_tree_update_batch(b2p_tree *tree, int slot, const uint32_t *wins, int threads);
int b2p_tree_search_ex(b2p_ctx *ctx, b2p_tree *tree, const b2p_search_opts *opts, b2p_search_stats *stats_out);

#ifdef __cplusplus
}
#endif
#endif /* B2P_H */
