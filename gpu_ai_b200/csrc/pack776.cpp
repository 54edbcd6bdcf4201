// gpu_ai_b200/csrc/pack776.cpp -- reference `State` (776-byte AoS, src/state.hpp:119-122) -> b2p_state16, the
// host-side hot loop of the reference-facing call: b2p_run_states776 streams 813 MB of `State`s per 2^20 leaves
// through this function, and the call's throughput IS this function's throughput.
//
// Layout (probed): board[8][8] of BoardItem {bool occupied @0, int32 type @4, int32 owner @8} = 12 B,
// turn @768, movesSinceLastCapture @772.  Only the 32 dark squares matter; type/owner count only where `occupied`
// is set (State::move leaves stale fields in vacated squares, src/state.cu:78-84).
//
// Two implementations behind one entry point, chosen once at start-up:
//   * AVX-512BW + BMI2 (every x86 server CPU since 2017): the 768 board bytes are 12 cache lines; one VPTESTMB per line
//     turns "byte != 0" into a 64-bit mask, and three PEXTs per line pull out the occupied / type / owner bits of
//     the dark squares in that line, already in square order.  ~120 instructions per state.
//   * portable scalar: one 8-byte and one 4-byte load per dark square.  ~320 instructions per state.
#include "pack776.h"

#include <cstdlib>
#include <cstring>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define B2P_HAVE_X86_DISPATCH 1
#endif

namespace b2p {

namespace {

constexpr size_t kStateBytes = 776, kItemBytes = 12, kTurnOff = 768, kMscOff = 772;

inline uint32_t min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

void pack_scalar(const unsigned char *s, size_t n, b2p_state16 *out) {
  for (size_t i = 0; i < n; i++, s += kStateBytes) {
    uint32_t occ = 0, p2 = 0, k = 0;
    for (int r = 0; r < 8; r++) {
      const unsigned char *row = s + 8 * kItemBytes * (size_t)r + kItemBytes * (size_t)((r & 1) ^ 1);
      for (int j = 0; j < 4; j++) {
        const unsigned char *q = row + 2 * kItemBytes * (size_t)j;
        uint64_t w;
        uint32_t owner;
        std::memcpy(&w, q, 8);
        std::memcpy(&owner, q + 8, 4);
        // the LOW BYTE of each field decides, in both implementations (the enums only take the values 0 and 1
        // where a piece stands): any byte pattern packs to the same bits on either path
        const uint32_t oc = (uint32_t)(w & 0xFFu) != 0u;          // BoardItem::occupied
        const uint32_t ty = (uint32_t)((w >> 32) & 0xFFu) != 0u;  // CHECKER_KING = 1
        const uint32_t ow = (owner & 0xFFu) != 0u;                // PLAYER_2 = 1
        const int sq = r * 4 + j;
        occ |= oc << sq;
        p2 |= (oc & ow) << sq;
        k |= (oc & ty) << sq;
      }
    }
    int32_t turn;
    uint32_t msc;
    std::memcpy(&turn, s + kTurnOff, 4);
    std::memcpy(&msc, s + kMscOff, 4);
    out[i].p1 = occ & ~p2;
    out[i].p2 = p2;
    out[i].kings = k;
    out[i].meta = (turn == 1 ? 1u : 0u) | (min_u32(msc, 0xFFFFFFu) << 8);
  }
}

#if defined(B2P_HAVE_X86_DISPATCH)
// per cache line (12 of them) and field (occupied, type, owner): which bytes of the line are that field's low byte
// of a dark square, and how many such bytes the lines before it hold (= where this line's bits go)
struct LineTables {
  uint64_t mask[12][3];
  uint32_t base[12][3];
  LineTables() {
    std::memset(mask, 0, sizeof mask);
    std::memset(base, 0, sizeof base);
    static const int field_off[3] = {0, 4, 8};
    for (int f = 0; f < 3; f++) {
      int seen = 0, line = 0;
      for (int sq = 0; sq < 32; sq++) {
        const int r = sq >> 2, c = 2 * (sq & 3) + ((r & 1) ^ 1);
        const int byte = (int)kItemBytes * (8 * r + c) + field_off[f];
        while (line < byte / 64) base[++line][f] = (uint32_t)seen;
        mask[byte / 64][f] |= 1ull << (byte % 64);
        seen++;
      }
      while (line < 11) base[++line][f] = (uint32_t)seen;
    }
  }
};

__attribute__((target("avx512f,avx512bw,bmi2"))) void pack_avx512(const unsigned char *s, size_t n, b2p_state16 *out) {
  static const LineTables T;
  for (size_t i = 0; i < n; i++, s += kStateBytes) {
    uint64_t occ = 0, typ = 0, own = 0;  // 64-bit accumulators: a line's bits may be shifted by up to 31
#pragma GCC unroll 12
    for (int l = 0; l < 12; l++) {
      const __m512i v = _mm512_loadu_si512((const void *)(s + 64 * l));
      const uint64_t nz = (uint64_t)_mm512_test_epi8_mask(v, v);  // byte != 0
      occ |= _pext_u64(nz, T.mask[l][0]) << T.base[l][0];
      typ |= _pext_u64(nz, T.mask[l][1]) << T.base[l][1];
      own |= _pext_u64(nz, T.mask[l][2]) << T.base[l][2];
    }
    // the type / owner enums only take the values 0 and 1 where a piece stands, so "low byte != 0" is "== 1" there;
    // stale values in vacated squares are masked by `occupied`
    const uint32_t o = (uint32_t)occ, p2 = o & (uint32_t)own, k = o & (uint32_t)typ;
    int32_t turn;
    uint32_t msc;
    std::memcpy(&turn, s + kTurnOff, 4);
    std::memcpy(&msc, s + kMscOff, 4);
    out[i].p1 = o & ~p2;
    out[i].p2 = p2;
    out[i].kings = k;
    out[i].meta = (turn == 1 ? 1u : 0u) | (min_u32(msc, 0xFFFFFFu) << 8);
  }
}
#endif

using PackFn = void (*)(const unsigned char *, size_t, b2p_state16 *);

PackFn choose() {
  const char *force = std::getenv("B2P_PACK_SCALAR");
  if (force && force[0] == '1') return pack_scalar;
#if defined(B2P_HAVE_X86_DISPATCH)
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("bmi2")) return pack_avx512;
#endif
  return pack_scalar;
}

const PackFn g_pack = choose();

}  // namespace

void pack776_range(const unsigned char *states, size_t n, b2p_state16 *out) { g_pack(states, n, out); }
void pack776_range_scalar(const unsigned char *states, size_t n, b2p_state16 *out) { pack_scalar(states, n, out); }
const char *pack776_impl() {
#if defined(B2P_HAVE_X86_DISPATCH)
  return g_pack == pack_avx512 ? "avx512bw+bmi2" : "scalar";
#else
  return "scalar";
#endif
}

}  // namespace b2p
