// obs_swar.cuh -- the observation half of the state phase (window gather, process_vis, Grid.encode, tile kinds) on
// whole ROWS of the 7x7 window at a time: seven bytes per 64-bit register instead of one cell per instruction.
//
// Same results as gather_view + visibility + sym_of_code in env_logic.cuh (the per-cell form, which stays the
// reference restatement and serves grids narrower than the window); ~2.2x fewer instructions per environment.
// Host+device like env_logic.cuh, so tests/csrc/host_model.cpp checks it against the per-cell form and the oracle on
// the CPU tier.
//
//   window rows     the 7x7 window is an axis-aligned square of the row-major grid: 7 unaligned 7-byte row loads
//                   (two aligned 64-bit loads + funnel shift), cells outside the grid = wall  (Grid.slice, upstream)
//   orientation     rotate_left x (dir + 1) = byte-matrix transpose (odd dir) + byte / row reversal, giving seven
//                   groups g[vi] whose byte vj is the cell the agent sees at view (vi, vj)
//   transparency    byte-parallel type test (wall, closed / locked door are opaque) -> 8x8 bit matrix -> transposed
//                   to the row masks process_vis sweeps (bits over vi, one row per vj)
//   visibility      the same row recurrence as env_logic.cuh::visibility, rows 8 bits apart
//   encode          (type, colour, state) byte vectors per group, masked by visibility, interleaved to the 21-byte
//                   piece of the [vi][vj][3] image; tile kinds = visible codes (0 = unseen), agent slot patched in
#pragma once
#include <stdint.h>
#include <string.h>

#include "env_logic.cuh"

namespace merlin {

constexpr uint64_t kB01 = 0x0101010101010101ull;
constexpr uint64_t kLow7 = 0x00ffffffffffffffull;                 // bytes 0..6
constexpr uint64_t kWall7 = (kB01 * CODE_WALL) & kLow7;

// Grid loads carry an L2 cache policy on the device (`pol`, from grid_policy() in env_kernels_common.cuh: evict-last for the
// shared layout pool, which every step re-reads while gigabytes of observations stream through the same L2).
MERLIN_HD uint64_t ld64(const uint8_t* p, uint64_t pol = 0) {
#if defined(__CUDA_ARCH__)
  uint64_t v;  // callers pass 8-byte aligned addresses
  asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
#else
  (void)pol;
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
#endif
}

// 16 bytes from a 16-byte aligned address as two little-endian 64-bit halves.
MERLIN_HD void ld128(const uint8_t* p, uint64_t& lo, uint64_t& hi, uint64_t pol = 0) {
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(lo), "=l"(hi) : "l"(p), "l"(pol));
#else
  (void)pol;
  memcpy(&lo, p, 8);
  memcpy(&hi, p + 8, 8);
#endif
}

MERLIN_HD uint64_t bswap64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
#else
  return __builtin_bswap64(x);
#endif
}

// The seven window rows: r[v] = codes of cells (x0 .. x0+6, y0 + v), byte u at bits 8u, CODE_WALL outside the grid.
// Needs W >= 7 and an 8-byte aligned `grid`.  Straight-line on purpose: all fourteen loads (two aligned 64-bit words
// per row; the second repeats the first when the seven cells do not straddle an 8-byte boundary) are independent of any
// branch, so they are in flight together -- with the loads behind per-row branches a lone warp pays one L2 round trip
// per row.  Every address lies inside [0, cell_stride) of this grid: a row outside the grid reads row 0 instead and is
// replaced by walls afterwards, and the second word ends at most at the next multiple of 8 after the row's last cell.
MERLIN_HD void window_rows(const uint8_t* grid, int W, int H, int x0, int y0, uint64_t (&r)[8], uint64_t pol = 0) {
  const int x0c = x0 < 0 ? 0 : (x0 > W - kView ? W - kView : x0);
  const int d = x0c - x0;                        // > 0: the window starts left of the grid, < 0: it ends right of it
  const int shl = d > 0 ? 8 * d : 0, shr = d < 0 ? -8 * d : 0;
  uint64_t col_valid = 0;
#pragma unroll
  for (int u = 0; u < kView; ++u)
    if ((unsigned)(x0 + u) < (unsigned)W) col_valid |= 0xffull << (8 * u);
  uint64_t lo[kView], hi[kView];
  int sh[kView];
  if (W == 16) {
    // the BASELINE grid: a row is ONE aligned 16-byte word -- seven 128-bit loads instead of fourteen 64-bit ones (the
    // symbolic-only kernel keeps the load/store unit ~70 % busy: ncu, profiles/r02_sym_kernel_ncu.md); x0c <= 9, so the
    // seven cells start in the low half (shift < 64, may spill into the high half) or lie in the high half entirely
    const int s16 = x0c * 8;
#pragma unroll
    for (int v = 0; v < kView; ++v) {
      const int wy = y0 + v;
      ld128(grid + ((unsigned)wy < (unsigned)H ? wy : 0) * 16, lo[v], hi[v], pol);
    }
#pragma unroll
    for (int v = 0; v < kView; ++v) {
      uint64_t x = s16 < 64 ? ((lo[v] >> s16) | ((hi[v] << 1) << (63 - s16))) : (hi[v] >> (s16 - 64));
      x = (x << shl) >> shr;
      x = (x & col_valid) | (kWall7 & ~col_valid);
      r[v] = (unsigned)(y0 + v) < (unsigned)H ? x : kWall7;
    }
    r[7] = 0;
    return;
  }
#pragma unroll
  for (int v = 0; v < kView; ++v) {
    const int wy = y0 + v;
    const int base = ((unsigned)wy < (unsigned)H ? wy : 0) * W + x0c;
    const int a0 = base & ~7;
    sh[v] = (base & 7) * 8;
    lo[v] = ld64(grid + a0, pol);
    hi[v] = ld64(grid + (sh[v] > 8 ? a0 + 8 : a0), pol);
  }
#pragma unroll
  for (int v = 0; v < kView; ++v) {
    uint64_t x = lo[v] >> sh[v];
    if (sh[v] > 8) x |= (hi[v] << 1) << (63 - sh[v]);
    x = (x << shl) >> shr;
    x = (x & col_valid) | (kWall7 & ~col_valid);
    r[v] = (unsigned)(y0 + v) < (unsigned)H ? x : kWall7;
  }
  r[7] = 0;
}

// 8x8 byte-matrix transpose: r[i] byte j <-> r[j] byte i.
MERLIN_HD void transpose_bytes8(uint64_t (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const uint64_t t = ((r[i] >> 8) ^ r[i + 1]) & 0x00ff00ff00ff00ffull;
    r[i + 1] ^= t; r[i] ^= t << 8;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i & 2) continue;
    const uint64_t t = ((r[i] >> 16) ^ r[i + 2]) & 0x0000ffff0000ffffull;
    r[i + 2] ^= t; r[i] ^= t << 16;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t t = ((r[i] >> 32) ^ r[i + 4]) & 0x00000000ffffffffull;
    r[i + 4] ^= t; r[i] ^= t << 32;
  }
}

// 8x8 bit-matrix transpose (row i at bits 8i..8i+7): bit (i, j) <-> bit (j, i).
MERLIN_HD uint64_t transpose_bits8(uint64_t x) {
  uint64_t t;
  t = (x ^ (x >> 7)) & 0x00aa00aa00aa00aaull; x ^= t ^ (t << 7);
  t = (x ^ (x >> 14)) & 0x0000cccc0000ccccull; x ^= t ^ (t << 14);
  t = (x ^ (x >> 28)) & 0x00000000f0f0f0f0ull; x ^= t ^ (t << 28);
  return x;
}

// 0x01 in every byte whose low nibble differs from c (bytes hold values <= 0x0f).
MERLIN_HD uint64_t nibble_ne(uint64_t t4, uint32_t c) {
  return (((t4 ^ (kB01 * c)) + kB01 * 0x0f) >> 4) & kB01;
}

// Grid.process_vis(agent_pos = (3, 6)) on an 8-bit-strided matrix: row vj at bits 8vj, bit vi set = transparent;
// returns the visibility matrix in the same layout (the recurrence of env_logic.cuh::visibility).
MERLIN_HD uint64_t visibility8(uint64_t transp) {
  uint64_t vis = 0;
  uint32_t seed = 1u << (kView / 2);
#pragma unroll
  for (int vj = kView - 1; vj >= 0; --vj) {
    uint32_t v;
    seed = vis_row(seed, (uint32_t)(transp >> (8 * vj)) & 0x7f, v);   // env_logic.cuh: carry-chain form of a row
    vis |= (uint64_t)v << (8 * vj);
  }
  return vis;
}

// What the agent sees.  g[vi]: byte vj = packed code at view (vi, vj) AFTER the invisible cells were erased (0) and the
// agent's own cell (3, 6) was replaced by what it carries (or empty); seen[vi]: 0xff in byte vj where visible.
// `doors` = false promises that no cell of the grid is a closed / locked door (walls are then the only opaque cells):
// one byte-parallel type test per group instead of three.
MERLIN_HD void observe_swar(const EnvState& s, const uint8_t* grid, int W, int H, uint64_t (&g)[kView],
                            uint64_t (&seen)[kView], bool doors = true, uint64_t pol = 0) {
  // window origin: get_view_exts
  const int x0 = s.dir == 0 ? s.x : (s.dir == 2 ? s.x - (kView - 1) : s.x - kView / 2);
  const int y0 = s.dir == 1 ? s.y : (s.dir == 3 ? s.y - (kView - 1) : s.y - kView / 2);
  uint64_t r[8];
  window_rows(grid, W, H, x0, y0, r, pol);
  // orientation: view (vi, vj) = window (u, v) with   dir 0: (6 - vj, vi)   dir 1: (6 - vi, 6 - vj)
  //                                                   dir 2: (vj, 6 - vi)   dir 3: (vi, vj)
  // i.e. g[vi] = rows (even dir) or columns (odd dir), bytes reversed for dir 0 / 1, index reversed for dir 1 / 2.
  uint64_t c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = r[i];
  transpose_bytes8(c);
  const bool odd = s.dir & 1, rev_bytes = s.dir < 2, rev_index = s.dir == 1 || s.dir == 2;
#pragma unroll
  for (int vi = 0; vi < kView; ++vi) {
    const uint64_t a = odd ? c[vi] : r[vi], b = odd ? c[kView - 1 - vi] : r[kView - 1 - vi];
    const uint64_t x = rev_index ? b : a;
    g[vi] = rev_bytes ? (bswap64(x) >> 8) : x;
  }
  // transparency (before the agent's cell is overwritten, like gen_obs_grid): opaque = wall, closed / locked door
  uint64_t tm = 0;  // row vi at bits 8vi, bit vj = transparent
#pragma unroll
  for (int vi = 0; vi < kView; ++vi) {
    const uint64_t t4 = g[vi] & (kB01 * 0x0f);
    uint64_t tb = nibble_ne(t4, T_WALL) & kLow7;
    if (doors) tb &= nibble_ne(t4, T_DOOR_CLOSED) & nibble_ne(t4, T_DOOR_LOCKED);
    // gather the seven byte LSBs into bits 0..6
    const uint64_t bits = (tb * 0x0102040810204080ull) >> 56;
    tm |= bits << (8 * vi);
  }
  const uint64_t vis_by_vj = visibility8(transpose_bits8(tm));
  const uint64_t vis_by_vi = transpose_bits8(vis_by_vj);
  const uint64_t own = s.carry ? s.carry : CODE_EMPTY;
  g[kView / 2] = (g[kView / 2] & ~(0xffull << (8 * (kView - 1)))) | (own << (8 * (kView - 1)));
#pragma unroll
  for (int vi = 0; vi < kView; ++vi) {
    const uint64_t bits = (vis_by_vi >> (8 * vi)) & 0x7f;
    // spread bit k to byte k, then widen to 0x00 / 0xff
    const uint64_t lsb = (bits * 0x0002040810204081ull) & kB01;
    seen[vi] = (lsb << 8) - lsb;
    g[vi] &= seen[vi];
  }
}

MERLIN_HD uint32_t perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  const uint64_t pool = ((uint64_t)b << 32) | a;
  uint32_t out = 0;
  for (int i = 0; i < 4; ++i) out |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return out;
#endif
}

// The 21-byte piece of the symbolic image for group vi (cells vj = 0..6, three bytes each) as six little-endian words
// (the last one holds a single byte): Grid.encode's (type, colour, state) of the visible codes, zeros where unseen.
// `doors` = false promises that no code is a closed / locked door: type = the low nibble, state = 0.
MERLIN_HD void encode_group(uint64_t codes, uint32_t (&w)[6], bool doors = true) {
  const uint64_t t4 = codes & (kB01 * 0x0f);
  const uint64_t col = (codes >> 4) & (kB01 * 0x07);
  uint64_t typ = t4, st = 0;
  if (doors) {
    const uint64_t door_lsb = ((t4 + kB01 * 0x05) >> 4) & kB01;          // low nibble >= 11: closed / locked door
    const uint64_t door = (door_lsb << 8) - door_lsb;
    typ = (t4 & ~door) | (kB01 * T_DOOR_OPEN & door);                    // doors encode as type 4 ...
    st = (((t4 | kB01 * 0x10) - kB01 * 0x0a) & (kB01 * 0x03)) & door;    // ... with state 1 / 2
  }
  const uint32_t t0 = (uint32_t)typ, t1 = (uint32_t)(typ >> 32), c0 = (uint32_t)col, c1 = (uint32_t)(col >> 32);
  const uint32_t s0 = (uint32_t)st, s1 = (uint32_t)(st >> 32);
  // byte p = 3 vj + k of the piece; words: T0 C0 S0 T1 | C1 S1 T2 C2 | S2 T3 C3 S3 | T4 C4 S4 T5 | C5 S5 T6 C6 | S6
  w[0] = perm(perm(t0, c0, 0x1040), s0, 0x3410);
  w[1] = perm(perm(c0, s0, 0x2051), t0, 0x3610);
  w[2] = perm(perm(s0, t0, 0x3072), c0, 0x3710);
  w[3] = perm(perm(t1, c1, 0x1040), s1, 0x3410);
  w[4] = perm(perm(c1, s1, 0x2051), t1, 0x3610);
  w[5] = (s1 >> 16) & 0xff;
}

// The 49 tile kinds of one env ([vi*7 + vj], what the frame phase indexes the atlas with) as 13 little-endian words
// (49 bytes + 3 of padding): the visible codes (0 = KIND_UNSEEN), the agent's cell = the agent tile over what it carries.
MERLIN_HD void kind_words(const uint64_t (&g)[kView], uint32_t carry, uint32_t (&w)[13]) {
  uint64_t k[kView];
#pragma unroll
  for (int vi = 0; vi < kView; ++vi) k[vi] = g[vi];
  k[kView / 2] = (k[kView / 2] & ~(0xffull << 48)) | ((uint64_t)agent_kind(carry) << 48);
  // byte stream: group vi occupies bytes 7vi .. 7vi+6, so its 64-bit word m is the tail of group m and the head of m+1
#pragma unroll
  for (int m = 0; m < kView; ++m) {
    const uint64_t q = (k[m] >> (8 * m)) | (m + 1 < kView ? k[m + 1] << (56 - 8 * m) : 0);
    w[2 * m] = (uint32_t)q;
    if (2 * m + 1 < 13) w[2 * m + 1] = (uint32_t)(q >> 32);
  }
}

}  // namespace merlin
