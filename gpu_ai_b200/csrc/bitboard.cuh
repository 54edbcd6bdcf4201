// gpu_ai_b200/csrc/bitboard.cuh
//
// Checkers rules of krame505/gpu_ai on a packed 32-square bitboard, written for sm_100a
// registers: every function is branch-light shift/mask arithmetic on 32-bit words, no local
// arrays, no recursion.  Replaces the reference's 776-byte AoS State and its recursive
// generators (reference: src/state.hpp:119-250, src/state.cu:57-92,122-180,239-340,388-420).
//
// Geometry (same numbering as the reference's own parallel generator, src/state.cu:185-188):
//   square i = row*4 + col/2 ; dark squares only: row even -> cols 1,3,5,7, row odd -> 0,2,4,6.
//   PLAYER_1 starts on rows 0-2 (bits 0-11) and moves toward row 7; PLAYER_2 the other way.
//
// Mover-normalised frame: all rule code below assumes "the side to move moves UP (toward
// row 7)".  When PLAYER_2 is to move the three words are bit-reversed (180-degree board
// rotation, i -> 31-i, one BREV each).  Under that rotation the reference's canonical move
// list (squares row-major, then its per-square generator order, then DFS order) is exactly
// REVERSED, because every direction order in the reference -- direct (+1,+1),(+1,-1),(-1,+1),
// (-1,-1) (src/state.cu:257-258), man capture left-then-right (src/state.cu:326-337), king
// capture (+1,+1),(+1,-1),(-1,+1),(-1,-1) (src/state.cu:391-392) -- is mapped onto its own
// reverse.  So canonical index k of PLAYER_2 = index n-1-k in the normalised frame.
//
// The B2P_HD macro lets tests/host_build compile this header with g++ to unit-test the bit
// logic on CPU against the oracle; the shipped library only uses the __device__ build.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define B2P_HD __host__ __device__ __forceinline__
#else
#define B2P_HD inline
#endif

// keep a float in a register: stops the compiler from sinking a division into a loop that
// only uses its result conditionally
#if defined(__CUDA_ARCH__)
#define B2P_PIN_FLOAT(x) asm volatile("" : "+f"(x))
#define B2P_PIN_INT(x) asm volatile("" : "+r"(x))
// explicit reconvergence of a known group of lanes: without it ptxas may keep sub-groups that took
// different rare paths apart and run the common code after them once per sub-group
#define B2P_REJOIN(mask) __syncwarp(mask)
#else
#define B2P_REJOIN(mask) ((void)(mask))
#define B2P_PIN_FLOAT(x) ((void)(x))
#define B2P_PIN_INT(x) ((void)(x))
#endif

namespace b2p {

// ---- portable bit primitives ------------------------------------------------------------
B2P_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
B2P_HD uint32_t brev(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  return __builtin_bswap32(x);
#endif
}
// index of the lowest set bit (x != 0)
B2P_HD int lowbit(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
B2P_HD uint32_t mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ---- one diagonal step, as a set image ------------------------------------------------------
// direction codes: 0 = UR (+1,+1)  1 = UL (+1,-1)  2 = DR (-1,+1)  3 = DL (-1,-1)
// From an even row: UL=i+4 UR=i+5 DL=i-4 DR=i-3 ; from an odd row: UL=i+3 UR=i+4 DL=i-5 DR=i-4.
// Masks drop the squares that would leave the board sideways (col 7 going right = even-row
// bits with i%4==3; col 0 going left = odd-row bits with i%4==0); rows fall off the word.
constexpr uint32_t kEvenRows = 0x0F0F0F0Fu;
constexpr uint32_t kOddRows = 0xF0F0F0F0u;
constexpr uint32_t kEvenNotCol7 = 0x07070707u;
constexpr uint32_t kOddNotCol0 = 0xE0E0E0E0u;

// x >> S for a compile-time S.  Experiment B2P_SHR_ON_FMA issues it as IMAD.HI (x * 2^(32-S) >> 32)
// to move work from the saturated ALU pipe (ncu: 85 %) to the FMA pipe (20 %).  Measured on B200
// (profiles/r01i_ab.txt): IMAD.HI is half rate (32/clk/SM) and the kernel gets 2.4 % SLOWER -- off.
template <int S>
B2P_HD uint32_t shr(uint32_t x) {
#if defined(__CUDA_ARCH__) && defined(B2P_SHR_ON_FMA)
  return __umulhi(x, 1u << (32 - S));
#else
  return x >> S;
#endif
}

B2P_HD uint32_t stepUR(uint32_t x) { return ((x & kEvenNotCol7) << 5) | ((x & kOddRows) << 4); }
B2P_HD uint32_t stepUL(uint32_t x) { return ((x & kEvenRows) << 4) | ((x & kOddNotCol0) << 3); }
B2P_HD uint32_t stepDR(uint32_t x) { return shr<3>(x & kEvenNotCol7) | shr<4>(x & kOddRows); }
B2P_HD uint32_t stepDL(uint32_t x) { return shr<4>(x & kEvenRows) | shr<5>(x & kOddNotCol0); }

// two steps in one direction (a jump landing): +9 / +7 / -7 / -9, legal while col+-2 stays on
// the board (i%4 <= 2 going right, i%4 >= 1 going left)
constexpr uint32_t kNotRight2 = 0x77777777u;
constexpr uint32_t kNotLeft2 = 0xEEEEEEEEu;
B2P_HD uint32_t jumpUR(uint32_t x) { return (x & kNotRight2) << 9; }
B2P_HD uint32_t jumpUL(uint32_t x) { return (x & kNotLeft2) << 7; }
B2P_HD uint32_t jumpDR(uint32_t x) { return shr<7>(x & kNotRight2); }
B2P_HD uint32_t jumpDL(uint32_t x) { return shr<9>(x & kNotLeft2); }

// per-square scalars: target index of one step / one jump from square s in direction d
B2P_HD int step_target(int s, int d) {
  // {+5,+4,-3,-4} on even rows, one less on odd rows
  const int base = (int)((0xFCFD0405u >> (8 * d)) & 0xFF);  // bytes: 5, 4, -3, -4 (two's complement)
  return s + (int)(int8_t)base - ((s >> 2) & 1);
}
B2P_HD int jump_target(int s, int d) {
  const int base = (int)((0xF7F90709u >> (8 * d)) & 0xFF);  // bytes: 9, 7, -7, -9
  return s + (int)(int8_t)base;
}

// the same as single-bit MASK arithmetic: the square one step / one jump away from the single-bit mask `from` in
// direction d is a rotation of the word by the table amount (minus one on odd rows for a step); the caller
// guarantees that the target is on the board, so nothing wraps.  One funnel shift instead of
// bit index -> arithmetic -> 1 << index.
B2P_HD uint32_t rotl32(uint32_t x, uint32_t amt) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(x, x, amt);
#else
  amt &= 31u;
  return amt ? (x << amt) | (x >> (32u - amt)) : x;
#endif
}
B2P_HD uint32_t step_mask(uint32_t from, int d) {
  const uint32_t odd = (from & kOddRows) ? 1u : 0u;
  return rotl32(from, ((0x1C1D0405u >> (8 * d)) & 0xFFu) - odd);  // +5, +4, -3, -4 (mod 32), one less on odd rows
}
B2P_HD uint32_t jump_mask(uint32_t from, int d) {
  return rotl32(from, (0x17190709u >> (8 * d)) & 0xFFu);          // +9, +7, -7, -9 (mod 32)
}

// ---- position in the mover-normalised frame -----------------------------------------------
struct Pos {
  uint32_t own;    // pieces of the side to move
  uint32_t opp;    // pieces of the other side
  uint32_t kings;  // kings of both sides
};

// jump availability per direction, independent of what stands on the origin square:
// J[d] bit s set <=> the square one step from s in direction d holds an enemy and the square
// two steps away is empty.  The reference never modifies the board while it extends a
// capture sequence (src/state.cu:129-131 test the original board), so these four masks are
// valid for EVERY hop of every sequence of the ply.
struct JumpMasks {
  uint32_t j[4];
};

B2P_HD JumpMasks jump_masks(const Pos &p) {
  const uint32_t empty = ~(p.own | p.opp);
  JumpMasks m;
  m.j[0] = stepDL(p.opp & stepDL(empty));  // s --UR--> enemy --UR--> empty
  m.j[1] = stepDR(p.opp & stepDR(empty));
  m.j[2] = stepUL(p.opp & stepUL(empty));
  m.j[3] = stepUR(p.opp & stepUR(empty));
  return m;
}

// squares of the mover that have a first hop, per direction (men: up only)
B2P_HD void capture_origins(const Pos &p, const JumpMasks &m, uint32_t out[4]) {
  const uint32_t ownK = p.own & p.kings;
  out[0] = p.own & m.j[0];
  out[1] = p.own & m.j[1];
  out[2] = ownK & m.j[2];
  out[3] = ownK & m.j[3];
}

// squares of the mover that have a direct move, per direction (reference: genLocDirectMoves,
// src/state.cu:255-281)
B2P_HD void step_origins(const Pos &p, uint32_t out[4]) {
  const uint32_t empty = ~(p.own | p.opp);
  const uint32_t ownK = p.own & p.kings;
  out[0] = p.own & stepDL(empty);
  out[1] = p.own & stepDR(empty);
  out[2] = ownK & stepUL(empty);
  out[3] = ownK & stepUR(empty);
}

// ---- applying a move ----------------------------------------------------------------------
// reference: State::move, src/state.cu:57-92.  `from`/`to` are single-bit masks, `captured`
// the set of jumped squares (empty for a direct move).  A man that ends on row 7 (bits 28-31
// of the normalised frame) is crowned: direct moves src/state.cu:274-276, captures :315-320.
B2P_HD void apply_move(Pos &p, uint32_t from, uint32_t to, uint32_t captured) {
  const uint32_t was_king = p.kings & from;
  p.own ^= from | to;
  p.opp &= ~captured;
  p.kings &= ~(captured | from);
  if (was_king | (to & 0xF0000000u)) p.kings |= to;
}

// hand the move to the other side: rotate the board by 180 degrees
B2P_HD Pos flip(const Pos &p) {
  Pos q;
  q.own = brev(p.opp);
  q.opp = brev(p.own);
  q.kings = brev(p.kings);
  return q;
}

// ---- selecting the k-th set bit ---------------------------------------------------------------
B2P_HD int select_bit(uint32_t m, int k) {
#if defined(__CUDA_ARCH__) && !defined(B2P_NO_PEEL_SELECT)
  // the masks are sparse (a handful of origins per direction): clearing k low bits beats the
  // 5-step popcount search even at the warp's worst lane (+2 % playouts/s, profiles/r01i_ab.txt)
  for (; k > 0; k--) m &= m - 1;
  return lowbit(m);
#endif
  int r = 0, c;
  c = popc(m & 0xFFFFu); if (k >= c) { k -= c; m >>= 16; r = 16; }
  c = popc(m & 0xFFu);   if (k >= c) { k -= c; m >>= 8;  r += 8; }
  c = popc(m & 0xFu);    if (k >= c) { k -= c; m >>= 4;  r += 4; }
  c = popc(m & 0x3u);    if (k >= c) { k -= c; m >>= 2;  r += 2; }
  c = (int)(m & 1u);     if (k >= c) { r += 1; }
  return r;
}

// Order of the (origin, slot) pairs of four origin masks a[0..3]:
//   kOrderFast      direction-major: all of a[0] by ascending origin, then a[1], ...
//   kOrderCanonical origin-major   : ascending origin, then slot 0..3 -- the reference's list
//                                    order in the normalised frame
// Returns origin | slot << 5 for rank k (0 <= k < total).
enum { kOrderCanonical = 0, kOrderFast = 1 };

// direction-major rank k as (single-bit origin mask, slot): no bit index is ever formed
B2P_HD uint32_t select_dir_major_mask(const uint32_t a[4], int n0, int n1, int n2, int k, int &slot) {
  slot = 0;
  uint32_t m = a[0];
  if (k >= n0) { k -= n0; m = a[1]; slot = 1;
    if (k >= n1) { k -= n1; m = a[2]; slot = 2;
      if (k >= n2) { k -= n2; m = a[3]; slot = 3; } } }
  for (; k > 0; k--) m &= m - 1;
  return m & (0u - m);
}

B2P_HD int select_dir_major(const uint32_t a[4], int n0, int n1, int n2, int k) {
  int slot = 0;
  uint32_t m = a[0];
  if (k >= n0) { k -= n0; m = a[1]; slot = 1;
    if (k >= n1) { k -= n1; m = a[2]; slot = 2;
      if (k >= n2) { k -= n2; m = a[3]; slot = 3; } } }
  return select_bit(m, k) | (slot << 5);
}

// Experiment B2P_SELECT2 (not yet measured on a GPU; same result, tests/host_build checks it against the search
// below on random masks): two levels -- the byte (two board rows) from three independent prefix counts, then a
// peel over the few origins inside that byte -- instead of five dependent search steps.
B2P_HD int select_origin_major_two_level(const uint32_t a[4], int k) {
  const int c8 = popc(a[0] & 0xFFu) + popc(a[1] & 0xFFu) + popc(a[2] & 0xFFu) + popc(a[3] & 0xFFu);
  const int c16 = popc(a[0] & 0xFFFFu) + popc(a[1] & 0xFFFFu) + popc(a[2] & 0xFFFFu) + popc(a[3] & 0xFFFFu);
  const int c24 = popc(a[0] & 0xFFFFFFu) + popc(a[1] & 0xFFFFFFu) + popc(a[2] & 0xFFFFFFu) + popc(a[3] & 0xFFFFFFu);
  const int byte = (k >= c8) + (k >= c16) + (k >= c24);
  k -= byte == 0 ? 0 : byte == 1 ? c8 : byte == 2 ? c16 : c24;
  const int shift = 8 * byte;
  const uint32_t x0 = (a[0] >> shift) & 0xFFu, x1 = (a[1] >> shift) & 0xFFu, x2 = (a[2] >> shift) & 0xFFu, x3 = (a[3] >> shift) & 0xFFu;
  uint32_t u = x0 | x1 | x2 | x3, bit = 0;
  for (;;) {
    bit = u & (0u - u);
    const int cnt = ((x0 & bit) != 0) + ((x1 & bit) != 0) + ((x2 & bit) != 0) + ((x3 & bit) != 0);
    if (k < cnt || u == 0) break;
    k -= cnt;
    u ^= bit;
  }
  uint32_t t = ((x0 & bit) ? 1u : 0u) | ((x1 & bit) ? 2u : 0u) | ((x2 & bit) ? 4u : 0u) | ((x3 & bit) ? 8u : 0u);
  if (k >= 1) t &= t - 1;
  if (k >= 2) t &= t - 1;
  if (k >= 3) t &= t - 1;
  return (lowbit(bit | 0x100u) + shift) | (lowbit(t | 16u) << 5);
}

B2P_HD int select_origin_major(const uint32_t a[4], int k) {
#if defined(B2P_SELECT2)
  return select_origin_major_two_level(a, k);
#endif
  // binary search for the origin o with  count(origins < o) <= k < count(origins <= o)
  int lo = 0, below = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int w = 16; w >= 1; w >>= 1) {
    const uint32_t lm = (1u << (lo + w)) - 1u;
    const int c = popc(a[0] & lm) + popc(a[1] & lm) + popc(a[2] & lm) + popc(a[3] & lm);
    if (c <= k) { lo += w; below = c; }
  }
  int r = k - below;  // rank among the slots present at origin lo
  const uint32_t nib = ((a[0] >> lo) & 1u) | (((a[1] >> lo) & 1u) << 1) | (((a[2] >> lo) & 1u) << 2) | (((a[3] >> lo) & 1u) << 3);
  // r-th set bit of a 4-bit value
  uint32_t t = nib;
  if (r >= 1) t &= t - 1;
  if (r >= 2) t &= t - 1;
  if (r >= 3) t &= t - 1;
  return lo | (lowbit(t | 16u) << 5);  // | 16: defined result (slot 4) when the masks are empty
}

// ---- capture sequences: iterative DFS over the static jump graph -----------------------------
// Enumerates every complete capture sequence in the normalised canonical order: origins
// ascending; men try UL then UR (reference: left before right, src/state.cu:326-337), kings
// UR, UL, DR, DL with "landing square not landed on before" (src/state.cu:134-139); a
// sequence is complete when no hop is possible (src/state.cu:314-322, :403-409).
// The whole DFS state lives in six registers: per-depth "next slot" counters (3 bits each),
// the landing path (5 bits each), visited and captured masks.
struct CaptureMove {
  int from, to, hops;
  uint32_t captured;  // jumped squares
  uint64_t path;      // landing square of hop k in bits [5k, 5k+5)
};

// single-bit mask of square s, 0 when s is off the board (s may be negative or >= 32)
B2P_HD uint32_t bit_or_zero(int s) { return (uint32_t)(1ull << (s & 63)); }

// Sequences that start on square `o` (which must hold a piece of the mover with a first hop).
// Calls visit(const CaptureMove&) for each complete sequence in reference order; visit returns
// true to stop (then `stopped` is set).  Returns the number of sequences visited.
//
// The king's "never land twice on the same square" rule (src/state.cu:134-139) is kept in the
// jump masks themselves: entering a landing square removes, in a working copy of the four masks,
// the (at most four) hops that end on it; leaving it restores them.  Expanding a node is then
// four bit extractions instead of four visited-set tests.
template <class Visit>
B2P_HD int for_each_capture_from(const Pos &p, const JumpMasks &m, int o, Visit &&visit, bool &stopped) {
  int count = 0;
  const bool king = (p.kings >> o) & 1u;
  int cur = o, depth = 0;
  uint32_t next = 0;       // next slot to try at each depth (3 bits per depth)
  uint32_t captured = 0;
  uint64_t path = 0;
  uint32_t w0 = m.j[0], w1 = m.j[1], w2 = m.j[2], w3 = m.j[3];  // working masks (kings)
  for (;;) {
    // slots that can be hopped from `cur` (slot order = reference order for this piece type)
    uint32_t v;
    if (king) v = ((w0 >> cur) & 1u) | (((w1 >> cur) & 1u) << 1) | (((w2 >> cur) & 1u) << 2) | (((w3 >> cur) & 1u) << 3);
    else v = ((m.j[1] >> cur) & 1u) | (((m.j[0] >> cur) & 1u) << 1);
    const int ns = (int)((next >> (3 * depth)) & 7u);
    const uint32_t todo = v & (0xFu << ns);
    if (todo == 0) {
      if (ns == 0 && depth > 0) {  // nothing was ever hoppable from here: a complete sequence
        CaptureMove cm;
        cm.from = o; cm.to = cur; cm.hops = depth; cm.captured = captured; cm.path = path;
        count++;
        if (visit(cm)) { stopped = true; return count; }
      }
      if (depth == 0) break;
      // pop: undo the hop that led to `cur`
      depth--;
      const int slot = (int)((next >> (3 * depth)) & 7u) - 1;
      const int d = king ? slot : (slot ^ 1);
      const int parent = cur - (jump_target(cur, d) - cur);
      if (king) {
        w0 |= m.j[0] & bit_or_zero(cur - 9);
        w1 |= m.j[1] & bit_or_zero(cur - 7);
        w2 |= m.j[2] & bit_or_zero(cur + 7);
        w3 |= m.j[3] & bit_or_zero(cur + 9);
      }
      captured &= ~(1u << step_target(parent, d));
      path &= ~((uint64_t)31 << (5 * depth));
      cur = parent;
    } else {
      const int slot = lowbit(todo);
      const int d = king ? slot : (slot ^ 1);
      next = (next & ~(7u << (3 * depth))) | ((uint32_t)(slot + 1) << (3 * depth));
      const int land = jump_target(cur, d);
      captured |= 1u << step_target(cur, d);
      if (king) {
        w0 &= ~bit_or_zero(land - 9);
        w1 &= ~bit_or_zero(land - 7);
        w2 &= ~bit_or_zero(land + 7);
        w3 &= ~bit_or_zero(land + 9);
      }
      path |= (uint64_t)land << (5 * depth);
      depth++;
      next &= ~(7u << (3 * depth));
      cur = land;
    }
  }
  return count;
}

// All capture sequences of the position, origins ascending.  Returns the number visited.
template <class Visit>
B2P_HD int for_each_capture(const Pos &p, const JumpMasks &m, Visit &&visit) {
  uint32_t cap[4];
  capture_origins(p, m, cap);
  uint32_t origins = cap[0] | cap[1] | cap[2] | cap[3];
  int count = 0;
  bool stopped = false;
  while (origins && !stopped) {
    const int o = lowbit(origins);
    origins &= origins - 1;
    count += for_each_capture_from(p, m, o, visit, stopped);
  }
  return count;
}

// ---- material (heuristic) -----------------------------------------------------------------
// reference: pieceValue / scoreState, src/heuristic.cu:7-31 (man 1, king 4)
B2P_HD uint32_t material(uint32_t pieces, uint32_t kings) { return (uint32_t)(popc(pieces) + 3 * popc(pieces & kings)); }

// ---- the 64-bit move record of b2p_genmoves (include/b2p.h) -----------------------------------
B2P_HD uint64_t encode_move(int from, int to, int hops, bool promoted, uint64_t path) {
  return (uint64_t)from | ((uint64_t)to << 5) | ((uint64_t)hops << 10) | ((uint64_t)(promoted ? 1 : 0) << 13) | (path << 16);
}

// Writes the reference-canonical move list (absolute frame, reference order) of a packed
// state; returns the number of legal moves (may exceed max_out; only max_out are stored).
// reference: State::genMoves, src/state.cu:239-245.
B2P_HD int gen_moves_canonical(uint32_t p1, uint32_t p2, uint32_t kings, uint32_t turn, uint64_t *out, int max_out) {
  Pos p;
  if (turn == 0) { p.own = p1; p.opp = p2; p.kings = kings; }
  else { p.own = brev(p2); p.opp = brev(p1); p.kings = brev(kings); }
  const JumpMasks jm = jump_masks(p);
  const uint32_t ownMen = p.own & ~p.kings;
  // pass 1: count (needed because PLAYER_2's list is the normalised list reversed)
  uint32_t cap[4];
  capture_origins(p, jm, cap);
  int n;
  const bool capture = (cap[0] | cap[1] | cap[2] | cap[3]) != 0;
  uint32_t st[4];
  if (capture) {
    n = for_each_capture(p, jm, [](const CaptureMove &) { return false; });
  } else {
    step_origins(p, st);
    n = popc(st[0]) + popc(st[1]) + popc(st[2]) + popc(st[3]);
  }
  // pass 2: emit
  auto put = [&](int idx_norm, int from, int to, int hops, bool promoted, uint64_t path) {
    int idx = idx_norm;
    if (turn != 0) {
      idx = n - 1 - idx_norm;
      from = 31 - from; to = 31 - to;
      uint64_t q = 0;
      for (int k = 0; k < hops; k++) q |= (uint64_t)(31 - (int)((path >> (5 * k)) & 31)) << (5 * k);
      path = q;
    }
    if (idx < max_out) out[idx] = encode_move(from, to, hops, promoted, path);
  };
  if (capture) {
    int i = 0;
    for_each_capture(p, jm, [&](const CaptureMove &cm) {
      const bool promoted = ((ownMen >> cm.from) & 1u) && cm.to >= 28;
      put(i++, cm.from, cm.to, cm.hops, promoted, cm.path);
      return false;
    });
  } else {
    for (int k = 0; k < n; k++) {
      const int sel = select_origin_major(st, k);
      const int o = sel & 31, d = sel >> 5;
      const int to = step_target(o, d);
      const bool promoted = ((ownMen >> o) & 1u) && to >= 28;
      put(k, o, to, 0, promoted, 0);
    }
  }
  return n;
}

}  // namespace b2p
