/*
 * oracle/checkers_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU-only restatement of the checkers rules and playout loops of
 * krame505/gpu_ai (the "reference", /root/reference).  It exists so that the
 * CUDA path in gpu_ai_b200/ can be checked bit-for-bit on machines where the
 * reference sources are absent (the GPU box).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * nothing under gpu_ai_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *  (1) oracle/_ref/libref_harness.so = the unmodified reference rules
 *      (src/state.cu, src/state.cpp, src/heuristic.cu) compiled in the build
 *      container, move-list for move-list and playout for playout, and
 *  (2) the golden vectors under tests/golden/ that were produced by that
 *      reference build (tools/make_golden.py), plus the perft table and the
 *      known-answer positions listed in SURVEY.md section 8c.
 *
 * It deliberately works on an 8x8 (row, col) board like the reference -- not on
 * the 32-bit bitboards the CUDA kernels use -- so that a bit-twiddling mistake
 * in the product cannot be mirrored here.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#include "chooser.h"

#define OR_P1 0
#define OR_P2 1
#define OR_NONE (-1)
#define OR_DRAW_PLIES 50u   /* NUM_DRAW_MOVES, src/state.hpp:14 */
#define OR_MAX_MOVES 128    /* reference buffers hold 100 (src/state.hpp:19) */
#define OR_MAX_HOPS 8       /* MAX_MOVE_JUMPS, src/state.hpp:17 */

typedef struct {
  uint8_t occ;   /* BoardItem::occupied */
  uint8_t king;  /* BoardItem::type == CHECKER_KING */
  uint8_t owner; /* BoardItem::owner (0/1) */
} or_square;

typedef struct {
  or_square at[8][8]; /* State::board, src/state.hpp:120 */
  int turn;           /* State::turn */
  unsigned msc;       /* State::movesSinceLastCapture */
} or_state;

typedef struct {
  int8_t fr, fc, tr, tc;                        /* Move::from / Move::to */
  uint8_t hops;                                 /* Move::jumps */
  uint8_t crowned;                              /* Move::promoted */
  int8_t cap_r[OR_MAX_HOPS], cap_c[OR_MAX_HOPS]; /* Move::removed */
  int8_t via_r[OR_MAX_HOPS], via_c[OR_MAX_HOPS]; /* Move::intermediate (landing squares) */
} or_move;

static int on_board(int r, int c) { return r >= 0 && r < 8 && c >= 0 && c < 8; } /* Loc::isValid, src/state.cu:9-11 */

/* dark squares: row even -> cols 1,3,5,7 ; row odd -> cols 0,2,4,6 (src/state.cu:174) */
static int first_dark_col(int r) { return 1 - (r & 1); }

/* ------------------------------------------------------------------ */
/* packed <-> board conversion (square index i = row*4 + col/2, the     */
/* numbering of the reference's own parallel generator, state.cu:185-188) */
/* ------------------------------------------------------------------ */

void or_from_packed(const uint32_t w[4], or_state *s) {
  memset(s, 0, sizeof *s);
  for (int i = 0; i < 32; i++) {
    int r = i >> 2, c = 2 * (i & 3) + first_dark_col(r);
    uint32_t bit = 1u << i;
    if (w[0] & bit) { s->at[r][c].occ = 1; s->at[r][c].owner = OR_P1; }
    if (w[1] & bit) { s->at[r][c].occ = 1; s->at[r][c].owner = OR_P2; }
    if ((w[0] | w[1]) & w[2] & bit) s->at[r][c].king = 1;
  }
  s->turn = (int)(w[3] & 1u);
  s->msc = w[3] >> 8;
}

void or_to_packed(const or_state *s, uint32_t w[4]) {
  w[0] = w[1] = w[2] = 0;
  for (int i = 0; i < 32; i++) {
    int r = i >> 2, c = 2 * (i & 3) + first_dark_col(r);
    const or_square *q = &s->at[r][c];
    if (!q->occ) continue;
    w[q->owner == OR_P1 ? 0 : 1] |= 1u << i;
    if (q->king) w[2] |= 1u << i;
  }
  unsigned m = s->msc > 0xFFFFFFu ? 0xFFFFFFu : s->msc;
  w[3] = (uint32_t)(s->turn & 1) | (m << 8);
}

/* the reference's 776-byte AoS State: BoardItem = {bool occupied @0; int type @4; int owner @8} (12 B),
 * board[8][8] @0, turn @768, movesSinceLastCapture @772 (src/state.hpp:71-122; sizes probed) */
void or_from_ref776(const unsigned char *p, or_state *s) {
  memset(s, 0, sizeof *s);
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) {
      const unsigned char *q = p + 12 * (r * 8 + c);
      int32_t type, owner;
      memcpy(&type, q + 4, 4);
      memcpy(&owner, q + 8, 4);
      if (q[0]) {
        s->at[r][c].occ = 1;
        s->at[r][c].king = (type == 1);
        s->at[r][c].owner = (uint8_t)(owner & 1);
      }
    }
  int32_t turn;
  uint32_t msc;
  memcpy(&turn, p + 768, 4);
  memcpy(&msc, p + 772, 4);
  s->turn = turn;
  s->msc = msc;
}

void or_to_ref776(const or_state *s, unsigned char *p) {
  memset(p, 0, 776);
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) {
      unsigned char *q = p + 12 * (r * 8 + c);
      int32_t type = s->at[r][c].king, owner = s->at[r][c].owner;
      q[0] = s->at[r][c].occ;
      if (s->at[r][c].occ) {
        memcpy(q + 4, &type, 4);
        memcpy(q + 8, &owner, 4);
      }
    }
  int32_t turn = s->turn;
  uint32_t msc = s->msc;
  memcpy(p + 768, &turn, 4);
  memcpy(p + 772, &msc, 4);
}

/* getStartingState, src/state.cpp:25-40: P1 on rows 0-2, P2 on rows 5-7, P1 to move */
void or_start(or_state *s) {
  memset(s, 0, sizeof *s);
  for (int r = 0; r < 8; r++) {
    if (r == 3 || r == 4) continue;
    for (int c = first_dark_col(r); c < 8; c += 2) {
      s->at[r][c].occ = 1;
      s->at[r][c].owner = r < 3 ? OR_P1 : OR_P2;
    }
  }
  s->turn = OR_P1;
  s->msc = 0;
}

/* ------------------------------------------------------------------ */
/* move generation                                                     */
/* ------------------------------------------------------------------ */

/* State::isValidJump, src/state.cu:122-142: the jumped square holds an enemy of the side to
 * move, the landing square is empty ON THE UNMODIFIED BOARD, and (kings only) the landing
 * square was not already landed on earlier in this sequence. */
static int hop_ok(const or_state *s, const or_move *sofar, int jr, int jc, int lr, int lc, int check_cycles) {
  if (!on_board(lr, lc) || !on_board(jr, jc)) return 0;
  if (!s->at[jr][jc].occ || s->at[jr][jc].owner == s->turn || s->at[lr][lc].occ) return 0;
  if (check_cycles)
    for (int i = (int)sofar->hops - 1; i >= 0; i--)
      if (sofar->via_r[i] == lr && sofar->via_c[i] == lc) return 0;
  return 1;
}

/* Move::addJump, src/state.cu:456-462 */
static void push_hop(or_move *m, int lr, int lc) {
  m->via_r[m->hops] = (int8_t)lr;
  m->via_c[m->hops] = (int8_t)lc;
  m->cap_r[m->hops] = (int8_t)((lr - m->tr) / 2 + m->tr);
  m->cap_c[m->hops] = (int8_t)((lc - m->tc) / 2 + m->tc);
  m->hops++;
  m->tr = (int8_t)lr;
  m->tc = (int8_t)lc;
}

/* genLocCaptureReg, src/state.cu:283-340: forward-only DFS, LEFT (dc = -1) before RIGHT (dc = +1);
 * a sequence is emitted when no further hop exists; promoted iff it ends on the far row. */
static int man_captures(const or_state *s, const or_move *sofar, or_move *out, int n) {
  int fwd = s->turn == OR_P1 ? 1 : -1; /* P1 moves toward row 7, src/state.cu:292-302 */
  int r = sofar->tr, c = sofar->tc;
  int any = 0;
  for (int side = -1; side <= 1; side += 2) {
    if (!hop_ok(s, sofar, r + fwd, c + side, r + 2 * fwd, c + 2 * side, 0)) continue;
    any = 1;
    or_move next = *sofar;
    push_hop(&next, r + 2 * fwd, c + 2 * side);
    n = man_captures(s, &next, out, n);
  }
  if (!any && sofar->hops > 0) {
    out[n] = *sofar;
    out[n].crowned = (s->turn == OR_P1 && sofar->tr == 7) || (s->turn == OR_P2 && sofar->tr == 0); /* :315-320 */
    n++;
  }
  return n;
}

/* genLocCaptureKing, src/state.cu:388-420: directions (+1,+1),(+1,-1),(-1,+1),(-1,-1), cycle check on */
static int king_captures(const or_state *s, const or_move *sofar, or_move *out, int n) {
  static const int dr[4] = {1, 1, -1, -1}, dc[4] = {1, -1, 1, -1};
  int r = sofar->tr, c = sofar->tc;
  int any = 0;
  for (int d = 0; d < 4; d++) {
    if (!hop_ok(s, sofar, r + dr[d], c + dc[d], r + 2 * dr[d], c + 2 * dc[d], 1)) continue;
    any = 1;
    or_move next = *sofar;
    push_hop(&next, r + 2 * dr[d], c + 2 * dc[d]);
    n = king_captures(s, &next, out, n);
  }
  if (!any && sofar->hops > 0) out[n++] = *sofar; /* kings are never "promoted" */
  return n;
}

/* genLocDirectMoves, src/state.cu:255-281: king (+1,+1),(+1,-1),(-1,+1),(-1,-1); P1 man the first
 * two, P2 man the last two; promoted iff a man reaches row 7 (P1) / row 0 (P2). */
static int step_moves(const or_state *s, int r, int c, or_move *out, int n) {
  static const int dr[4] = {1, 1, -1, -1}, dc[4] = {1, -1, 1, -1};
  const or_square *q = &s->at[r][c];
  int lo = 0, hi = 4;
  if (!q->king) {
    lo = q->owner == OR_P1 ? 0 : 2;
    hi = lo + 2;
  }
  for (int d = lo; d < hi; d++) {
    int tr = r + dr[d], tc = c + dc[d];
    if (!on_board(tr, tc) || s->at[tr][tc].occ) continue;
    or_move m;
    memset(&m, 0, sizeof m);
    m.fr = (int8_t)r; m.fc = (int8_t)c; m.tr = (int8_t)tr; m.tc = (int8_t)tc;
    m.crowned = !q->king && tr == (q->owner == OR_P1 ? 7 : 0);
    out[n++] = m;
  }
  return n;
}

/* State::genMoves / genTypeMoves / genLocMoves, src/state.cu:239-245,171-180,144-169:
 * every complete capture sequence (squares row-major), else every direct move. */
int or_gen_moves(const or_state *s, or_move out[OR_MAX_MOVES]) {
  int n = 0;
  for (int pass = 0; pass < 2 && n == 0; pass++)
    for (int r = 0; r < 8; r++)
      for (int c = first_dark_col(r); c < 8; c += 2) {
        const or_square *q = &s->at[r][c];
        if (!q->occ || q->owner != s->turn) continue;
        if (pass == 0) {
          or_move seed;
          memset(&seed, 0, sizeof seed);
          seed.fr = seed.tr = (int8_t)r;
          seed.fc = seed.tc = (int8_t)c;
          n = q->king ? king_captures(s, &seed, out, n) : man_captures(s, &seed, out, n);
        } else {
          n = step_moves(s, r, c, out, n);
        }
      }
  return n;
}

/* State::move, src/state.cu:57-92 */
void or_apply(or_state *s, const or_move *m) {
  for (int i = 0; i < m->hops; i++) s->at[m->cap_r[i]][m->cap_c[i]].occ = 0;
  or_square piece = s->at[m->fr][m->fc];
  s->at[m->fr][m->fc].occ = 0;
  s->at[m->tr][m->tc].occ = 1;
  s->at[m->tr][m->tc].king = m->crowned ? 1 : piece.king;
  s->at[m->tr][m->tc].owner = piece.owner;
  s->turn = s->turn == OR_P1 ? OR_P2 : OR_P1;
  s->msc = m->hops ? 0 : s->msc + 1;
}

/* State::isGameOver / getWinner, src/state.cpp:16-23: no move or 50 plies without capture;
 * the draw test wins over "no moves". */
static int outcome(const or_state *s, int n_moves) {
  if (s->msc >= OR_DRAW_PLIES) return OR_NONE;
  if (n_moves == 0) return s->turn == OR_P1 ? OR_P2 : OR_P1;
  return 2; /* still running */
}

uint64_t or_perft(const or_state *s, int depth) {
  or_move mv[OR_MAX_MOVES];
  int n = or_gen_moves(s, mv);
  if (depth == 1) return (uint64_t)n;
  uint64_t total = 0;
  for (int i = 0; i < n; i++) {
    or_state t = *s;
    or_apply(&t, &mv[i]);
    total += or_perft(&t, depth - 1);
  }
  return total;
}

/* ------------------------------------------------------------------ */
/* compact 64-bit move record shared with the product's genmoves output  */
/* (include/b2p.h, b2p_move_t): [0:5) from, [5:10) to, [10:13) hops,     */
/* [13] promoted, [16+5k : 21+5k) landing square of hop k (k < 7)         */
/* ------------------------------------------------------------------ */
static uint32_t sq_index(int r, int c) { return (uint32_t)(r * 4 + c / 2); }

uint64_t or_encode_move(const or_move *m) {
  uint64_t e = sq_index(m->fr, m->fc) | ((uint64_t)sq_index(m->tr, m->tc) << 5) |
               ((uint64_t)(m->hops & 7u) << 10) | ((uint64_t)(m->crowned != 0) << 13);
  for (int k = 0; k < m->hops && k < 7; k++) e |= (uint64_t)sq_index(m->via_r[k], m->via_c[k]) << (16 + 5 * k);
  return e;
}

/* ------------------------------------------------------------------ */
/* fast-order rank: the order in which the throughput kernel enumerates  */
/* the legal moves (DESIGN.md "move order").  Positions whose capture     */
/* list needs a full enumeration (see below) keep the canonical order.    */
/* Otherwise moves are sorted by (direction of the first hop or step as   */
/* seen by the mover, origin in the mover's frame = square index for P1,  */
/* 31 - index for P2).                                                    */
/* Returns the canonical index of the move with fast-order rank j.        */
/* ------------------------------------------------------------------ */
static int fast_order_pick(const or_state *s, const or_move *mv, int n, int j) {
  /* full enumeration shape: a king sequence with >= 3 hops, or two sequences that share origin and first
   * landing square (a choice on a later square) -> canonical order (reversed for PLAYER_2) */
  int full = 0;
  for (int i = 0; i < n && !full; i++) {
    if (mv[i].hops >= 3 && s->at[mv[i].fr][mv[i].fc].king) full = 1;
    for (int k = i + 1; k < n && !full; k++)
      if (mv[i].hops >= 1 && mv[k].hops >= 1 && mv[i].fr == mv[k].fr && mv[i].fc == mv[k].fc &&
          mv[i].via_r[0] == mv[k].via_r[0] && mv[i].via_c[0] == mv[k].via_c[0]) full = 1;
  }
  if (full) return s->turn == OR_P1 ? j : n - 1 - j;
  int key[OR_MAX_MOVES];
  for (int i = 0; i < n; i++) {
    /* direction of the first hop (or of the step) */
    int t_r = mv[i].hops ? mv[i].via_r[0] : mv[i].tr, t_c = mv[i].hops ? mv[i].via_c[0] : mv[i].tc;
    int dr = (t_r > mv[i].fr) ? 1 : -1, dc = (t_c > mv[i].fc) ? 1 : -1;
    /* direction as seen by the mover (board rotated by 180 degrees for P2) */
    if (s->turn == OR_P2) { dr = -dr; dc = -dc; }
    int dir = dr > 0 ? (dc > 0 ? 0 : 1) : (dc > 0 ? 2 : 3); /* UR, UL, DR, DL */
    int origin = (int)sq_index(mv[i].fr, mv[i].fc);
    if (s->turn == OR_P2) origin = 31 - origin;
    key[i] = dir * 32 + origin;
  }
  /* rank j in ascending key order (keys are distinct) */
  for (int i = 0; i < n; i++) {
    int below = 0;
    for (int k = 0; k < n; k++) below += key[k] < key[i];
    if (below == j) return i;
  }
  return -1;
}

#define OR_ORDER_CANONICAL 0
#define OR_ORDER_FAST 1

/* ------------------------------------------------------------------ */
/* playouts                                                            */
/* ------------------------------------------------------------------ */

/* HostPlayoutDriver::runPlayouts body, src/playout.cpp:22-29, with the reference's
 * `moves[rand() % n]` (src/player.cpp:13-16) replaced by the Philox chooser of chooser.h.
 * max_plies < 0 = play to the end.  Returns the winner (-1/0/1) or 2 if stopped early. */
int or_random_playout(or_state *s, uint64_t key, uint64_t pid, uint32_t domain, uint32_t first_draw,
                      int order, int max_plies, uint32_t *plies_out) {
  or_move mv[OR_MAX_MOVES];
  uint32_t ply = 0;
  int result;
  for (;;) {
    int n = or_gen_moves(s, mv);
    result = outcome(s, n);
    if (result != 2) break;
    if (max_plies >= 0 && (int)ply >= max_plies) break;
    uint32_t j = ch_mulhi32(ch_draw(key, pid, domain, first_draw + ply), (uint32_t)n);
    int pick = order == OR_ORDER_FAST ? fast_order_pick(s, mv, n, (int)j) : (int)j;
    or_apply(s, &mv[pick]);
    ply++;
  }
  if (plies_out) *plies_out = ply;
  return result;
}

/* pieceValue / scoreState, src/heuristic.cu:7-31 (man 1, king 4) */
static void material(const or_state *s, unsigned score[2]) {
  score[0] = score[1] = 0;
  for (int r = 0; r < 8; r++)
    for (int c = first_dark_col(r); c < 8; c += 2)
      if (s->at[r][c].occ) score[s->at[r][c].owner] += s->at[r][c].king ? 4u : 1u;
}

/* HostHeuristicPlayoutDriver::runPlayouts body, src/heuristicPlayout.cpp:20-45:
 * scoreMove (src/heuristic.cu:33-42): promotion +3 to the mover, minus the value of every
 * captured piece for the opponent; getWeight (src/heuristic.cu:44-49): unsigned sums, then
 * float(my) / opp; argmax of weight + noise with strict '>' (first maximum wins); the running
 * material score is updated incrementally. */
int or_heuristic_playout(or_state *s, uint64_t key, uint64_t pid, int max_plies, uint32_t *plies_out) {
  or_move mv[OR_MAX_MOVES];
  unsigned score[2];
  material(s, score);
  uint32_t ply = 0;
  int result;
  for (;;) {
    int n = or_gen_moves(s, mv);
    result = outcome(s, n);
    if (result != 2) break;
    if (max_plies >= 0 && (int)ply >= max_plies) break;
    int me = s->turn, you = 1 - s->turn;
    int best = -1, best_d_me = 0, best_d_you = 0;
    float best_w = -INFINITY;
    for (int i = 0; i < n; i++) {
      int d_me = mv[i].crowned ? 3 : 0, d_you = 0;
      for (int k = 0; k < mv[i].hops; k++) d_you -= s->at[mv[i].cap_r[k]][mv[i].cap_c[k]].king ? 4 : 1;
      unsigned a = score[me] + (unsigned)d_me, b = score[you] + (unsigned)d_you;
      float w = (float)a / (float)b + ch_gauss_sigma(ch_noise_draw(key, pid, ply, (uint32_t)i));
      if (w > best_w) { best_w = w; best = i; best_d_me = d_me; best_d_you = d_you; }
    }
    or_apply(s, &mv[best]);
    score[me] += (unsigned)best_d_me;
    score[you] += (unsigned)best_d_you;
    ply++;
  }
  if (plies_out) *plies_out = ply;
  return result;
}

/* ------------------------------------------------------------------ */
/* batch entry points (ctypes)                                          */
/* ------------------------------------------------------------------ */

/* canonical move lists of n packed states; moves_out has n*max_moves records */
void or_genmoves_batch(const uint32_t *packed, size_t n, int max_moves, uint64_t *moves_out, uint8_t *counts_out) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    or_state s;
    or_move mv[OR_MAX_MOVES];
    or_from_packed(packed + 4 * i, &s);
    int cnt = or_gen_moves(&s, mv);
    counts_out[i] = (uint8_t)cnt;
    for (int k = 0; k < cnt && k < max_moves; k++) moves_out[i * (size_t)max_moves + k] = or_encode_move(&mv[k]);
  }
}

/* mode 0 = random, 1 = heuristic.  Playout id of (rep, leaf) = pid_base + rep*n + leaf.
 * winners_out/plies_out/final_out (4 words per playout) may be NULL. */
void or_playouts_batch(const uint32_t *packed, size_t n, uint32_t reps, uint64_t key, uint64_t pid_base, int mode,
                       int order, int max_plies, int8_t *winners_out, uint32_t *plies_out, uint32_t *final_out,
                       uint64_t counters_out[4]) {
  uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  size_t total = n * (size_t)reps;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c0, c1, c2, c3)
  for (size_t w = 0; w < total; w++) {
    or_state s;
    or_from_packed(packed + 4 * (w % n), &s);
    uint32_t plies = 0;
    int res = mode == 1 ? or_heuristic_playout(&s, key, pid_base + w, max_plies, &plies)
                        : or_random_playout(&s, key, pid_base + w, CH_DOMAIN_RANDOM, 0, order, max_plies, &plies);
    if (winners_out) winners_out[w] = (int8_t)res;
    if (plies_out) plies_out[w] = plies;
    if (final_out) or_to_packed(&s, final_out + 4 * w);
    if (res == OR_NONE) c0++; else if (res == OR_P1) c1++; else if (res == OR_P2) c2++;
    c3 += plies;
  }
  if (counters_out) { counters_out[0] = c0; counters_out[1] = c1; counters_out[2] = c2; counters_out[3] = c3; }
}

/* D_ref leaf set: the reference's genRandomStates recipe (src/driver.cpp:76-104: from the
 * start position play U{1..100} uniformly random plies, stop early when the game ends), with its
 * shared default_random_engine / rand() replaced by the Philox chooser so that leaf j is reproducible. */
void or_gen_leaves(size_t n, uint64_t key, uint64_t first_index, uint32_t *packed_out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t j = 0; j < n; j++) {
    or_state s;
    or_start(&s);
    uint64_t pid = first_index + j;
    int prefix = 1 + (int)ch_mulhi32(ch_draw(key, pid, CH_DOMAIN_LEAF, 0), 100u);
    or_random_playout(&s, key, pid, CH_DOMAIN_LEAF, 1, OR_ORDER_CANONICAL, prefix, NULL);
    or_to_packed(&s, packed_out + 4 * j);
  }
}

void or_pack776_batch(const unsigned char *states776, size_t n, uint32_t *packed_out) {
  for (size_t i = 0; i < n; i++) {
    or_state s;
    or_from_ref776(states776 + 776 * i, &s);
    or_to_packed(&s, packed_out + 4 * i);
  }
}

void or_unpack776_batch(const uint32_t *packed, size_t n, unsigned char *states776_out) {
  for (size_t i = 0; i < n; i++) {
    or_state s;
    or_from_packed(packed + 4 * i, &s);
    or_to_ref776(&s, states776_out + 776 * i);
  }
}

uint64_t or_perft_packed(const uint32_t packed[4], int depth) {
  or_state s;
  or_from_packed(packed, &s);
  return depth <= 0 ? 1 : or_perft(&s, depth);
}

void or_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { ch_philox4x32_10(ctr, key, out); }
uint32_t or_draw(uint64_t key, uint64_t pid, uint32_t domain, uint32_t t) { return ch_draw(key, pid, domain, t); }
float or_gauss(uint32_t r) { return ch_gauss_sigma(r); }
