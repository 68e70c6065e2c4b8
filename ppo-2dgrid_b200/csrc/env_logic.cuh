// env_logic.cuh -- per-environment arithmetic of the MERLIN rollout hot path, written once as
// host+device inline functions.  The CUDA kernels (env_kernels*.cu, render_kernels.cu) call these per lane; the
// host-only model in tests/csrc/host_model.cpp compiles the SAME functions with g++ so the logic can
// be checked against the oracle in the CPU test tier (it is a test vehicle, not a CPU fallback: the
// product library exports no host compute path).
//
// What each function restates (reference anchors; upstream = minigrid 3.0.0, un-vendored):
//   step_logic      MiniGridEnv.step  [upstream]           reached from src/ppo.py:76, src/fomaml.py:71
//                   ThreeActionWrapper                      src/wrappers/three_action_wrapper.py:10-17
//   shape_reward    StuckPenaltyWrapper.step                src/wrappers/stuck_penalty_wrapper.py:29-58
//                   + first-visit exploration bonus (builder-specified; absent from the reference)
//   gather_view     get_view_exts + Grid.slice + rotate_left x (dir+1) [upstream], as the closed form
//                   world = agent + (6 - vj) * f + (vi - 3) * r        
//   visibility      Grid.process_vis [upstream] as 7-bit row masks     
//   cell_kind / sym_of_code   Grid.encode / Grid.render tile choice [upstream]
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MERLIN_HD __host__ __device__ __forceinline__
#else
#define MERLIN_HD inline
#endif

namespace merlin {

constexpr int kView = 7;                              // agent_view_size (upstream default; base_env.py passes none)
constexpr int kTile = 8;                              // RGBImgPartialObsWrapper tile_size default
constexpr int kCells = kView * kView;                 // 49
constexpr int kRowBytes = kView * kTile * 3;          // 168
constexpr int kImgBytes = kRowBytes * kView * kTile;  // 9408
constexpr int kChunks = kImgBytes / 16;               // 588 16-byte chunks per frame
constexpr int kUnitsPerRow = kRowBytes / 8;           // 21 8-byte units per pixel row
constexpr int kSymBytes = kCells * 3;                 // 147
constexpr int kTileBytes = kTile * kTile * 3;         // 192
constexpr int kAtlasTiles = 128;

// packed cell codes (see include/merlin_b200.h)
constexpr uint32_t T_EMPTY = 1, T_WALL = 2, T_FLOOR = 3, T_DOOR_OPEN = 4, T_KEY = 5, T_BALL = 6, T_BOX = 7,
                   T_GOAL = 8, T_LAVA = 9, T_AGENT = 10, T_DOOR_CLOSED = 11, T_DOOR_LOCKED = 12;
constexpr uint32_t CODE_EMPTY = 1;
constexpr uint32_t CODE_WALL = T_WALL | (5u << 4);  // grey wall: what Grid.slice puts outside the grid
constexpr uint32_t KIND_UNSEEN = 0;                 // atlas slot of an invisible (erased, un-highlighted) cell
constexpr uint32_t KIND_AGENT = T_AGENT;            // agent triangle over an empty highlighted cell

// type-indexed predicate bitmasks
constexpr uint32_t M_OVERLAP = (1u << T_EMPTY) | (1u << T_FLOOR) | (1u << T_DOOR_OPEN) | (1u << T_GOAL) | (1u << T_LAVA);
constexpr uint32_t M_PICKUP = (1u << T_KEY) | (1u << T_BALL) | (1u << T_BOX);
constexpr uint32_t M_OPAQUE = (1u << T_WALL) | (1u << T_DOOR_CLOSED) | (1u << T_DOOR_LOCKED);

// actions (MiniGridEnv.Actions)
enum : int { A_LEFT = 0, A_RIGHT = 1, A_FORWARD = 2, A_PICKUP = 3, A_DROP = 4, A_TOGGLE = 5, A_DONE = 6 };

struct EnvState {
  int x, y, dir;        // agent pose
  uint32_t carry;       // packed code of the carried object, 0 = nothing
  int step_count;
  int layout;           // pool index of the loaded layout; negative = ~index pending for the first reset
  int stay;             // StuckPenalty counter
  int last_x, last_y;   // StuckPenalty last position
};

MERLIN_HD void unpack_state(int sx, int sy, int sz, int sw, EnvState& s) {
  s.x = sx & 0xff; s.y = (sx >> 8) & 0xff; s.dir = (sx >> 16) & 3; s.carry = ((uint32_t)sx >> 24) & 0x7f;
  s.step_count = sy; s.layout = sz;
  s.stay = sw & 0xffff; s.last_x = (sw >> 16) & 0xff; s.last_y = ((uint32_t)sw >> 24) & 0xff;
}
MERLIN_HD void pack_state(const EnvState& s, int& sx, int& sy, int& sz, int& sw) {
  sx = s.x | (s.y << 8) | (s.dir << 16) | (int)(s.carry << 24);
  sy = s.step_count; sz = s.layout;
  int stay = s.stay > 0xffff ? 0xffff : s.stay;
  sw = stay | (s.last_x << 16) | (int)((uint32_t)s.last_y << 24);
}

MERLIN_HD int dir_dx(int d) { return d == 0 ? 1 : (d == 2 ? -1 : 0); }
MERLIN_HD int dir_dy(int d) { return d == 1 ? 1 : (d == 3 ? -1 : 0); }

struct StepResult {
  double reward;
  bool terminated, truncated, bad_action;
  int write_idx;        // >= 0: grid cell index to overwrite with write_code (pickup/drop/toggle)
  uint32_t write_code;
};

// MiniGridEnv.step on one env. `fwd` is the packed code in front of the agent (CODE_WALL when outside the grid).
MERLIN_HD StepResult step_logic(EnvState& s, long long action, int n_actions, uint32_t fwd, bool fwd_in_grid,
                                int fwd_idx, int max_steps) {
  StepResult r;
  r.reward = 0.0; r.terminated = false; r.truncated = false; r.write_idx = -1; r.write_code = 0;
  r.bad_action = (action < 0 || action >= n_actions);
  const int a = r.bad_action ? A_DONE : (int)action;  // 3-action mode: {0,1,2} are already left/right/forward
  s.step_count += 1;
  const uint32_t ft = fwd & 0xf;
  const int fx = s.x + dir_dx(s.dir), fy = s.y + dir_dy(s.dir);
  if (a == A_LEFT) {
    s.dir = (s.dir + 3) & 3;
  } else if (a == A_RIGHT) {
    s.dir = (s.dir + 1) & 3;
  } else if (a == A_FORWARD) {
    if ((M_OVERLAP >> ft) & 1u) { s.x = fx; s.y = fy; }
    if (ft == T_GOAL) {
      r.terminated = true;
      r.reward = 1 - 0.9 * ((double)s.step_count / (double)max_steps);  // MiniGridEnv._reward, float64
    }
    if (ft == T_LAVA) r.terminated = true;
  } else if (a == A_PICKUP) {
    if (((M_PICKUP >> ft) & 1u) && s.carry == 0 && fwd_in_grid) {
      s.carry = fwd; r.write_idx = fwd_idx; r.write_code = CODE_EMPTY;
    }
  } else if (a == A_DROP) {
    if (ft == T_EMPTY && s.carry != 0 && fwd_in_grid) {
      r.write_idx = fwd_idx; r.write_code = s.carry; s.carry = 0;
    }
  } else if (a == A_TOGGLE) {
    if (fwd_in_grid) {
      const uint32_t col = fwd & 0x70;
      if (ft == T_DOOR_LOCKED) {
        if ((s.carry & 0xf) == T_KEY && (s.carry & 0x70) == col) { r.write_idx = fwd_idx; r.write_code = col | T_DOOR_OPEN; }
      } else if (ft == T_DOOR_CLOSED) {
        r.write_idx = fwd_idx; r.write_code = col | T_DOOR_OPEN;
      } else if (ft == T_DOOR_OPEN) {
        r.write_idx = fwd_idx; r.write_code = col | T_DOOR_CLOSED;
      } else if (ft == T_BOX) {
        r.write_idx = fwd_idx; r.write_code = CODE_EMPTY;  // boxes are empty here: replaced by their (None) contents
      }
    }
  }
  if (s.step_count >= max_steps) r.truncated = true;
  return r;
}

// Wrapper-stack reward shaping, in float64 like the python floats of the reference, cast to f32 by the caller.
// visited_word: the 32-bit word of the per-episode visited bitmap holding the agent's cell (in/out).
MERLIN_HD double shape_reward(EnvState& s, double reward, bool stuck_on, int max_stay, double penalty, bool explore_on,
                              double bonus, uint32_t& visited_word, int cell_bit, bool& stuck) {
  stuck = false;
  if (stuck_on) {
    s.stay = (s.x == s.last_x && s.y == s.last_y) ? s.stay + 1 : 0;
    if (s.stay >= max_stay) { reward += penalty; stuck = true; }
    s.last_x = s.x; s.last_y = s.y;
  }
  if (explore_on) {
    const uint32_t bit = 1u << cell_bit;
    if (!(visited_word & bit)) { visited_word |= bit; reward += bonus; }
  }
  return reward;
}

// ---------------------------------------------------------------------------------------------------------------
// Action sampling inside the step kernel ("act -> sample -> step -> store").  The reference samples
// with torch.distributions.Categorical(logits).sample() and stores .log_prob(a) (src/actor_critic.py act();
// src/ppo.py:70-86, src/fomaml.py:65-84).  torch's own generator stream cannot be reproduced, so the draw is
// specified here: Philox4x32-10 keyed by the sampler seed, counter = (env index, that env's draw number, 0, 0); the
// first output word gives u = (x >> 8) * 2^-24 in [0, 1); the action is the first a with cumsum(exp(l - max))[a] > u * sum
// (inverse CDF in float32, left to right); log-probability = (l_a - max) - log(sum).
MERLIN_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): first output word.
MERLIN_HD uint32_t philox4x32_10_x0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                    uint32_t* out4 = nullptr) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  if (out4) { out4[0] = c0; out4[1] = c1; out4[2] = c2; out4[3] = c3; }
  return c0;
}

MERLIN_HD float sampler_uniform(uint32_t seed_lo, uint32_t seed_hi, uint32_t env, uint32_t draw) {
  return (float)(philox4x32_10_x0(env, draw, 0u, 0u, seed_lo, seed_hi) >> 8) * (1.0f / 16777216.0f);
}

constexpr int kMaxActions = 7;

struct PolicySample {
  int action;
  float logp;
};

// Categorical(logits) over `n` <= 7 actions: inverse-CDF sample with uniform `u` (or argmax, first maximum, when
// `greedy`) and its log-probability.  Non-finite logits never index out of range: the fallback is the last action.
MERLIN_HD PolicySample sample_policy(const float (&logit)[kMaxActions], int n, float u, bool greedy) {
  float m = logit[0];
  int arg = 0;
#pragma unroll
  for (int a = 1; a < kMaxActions; ++a)
    if (a < n && logit[a] > m) { m = logit[a]; arg = a; }
  float ex[kMaxActions];
  float sum = 0.f;
#pragma unroll
  for (int a = 0; a < kMaxActions; ++a) {
    ex[a] = a < n ? expf(logit[a] - m) : 0.f;
    sum += ex[a];
  }
  int pick = n - 1;
  if (greedy) {
    pick = arg;
  } else {
    const float target = u * sum;
    float acc = 0.f;
    bool found = false;
#pragma unroll
    for (int a = 0; a < kMaxActions; ++a) {
      acc += ex[a];
      if (a < n && !found && acc > target) { pick = a; found = true; }
    }
  }
  float lsel = logit[0];
#pragma unroll
  for (int a = 1; a < kMaxActions; ++a)
    if (a == pick) lsel = logit[a];
  PolicySample s;
  s.action = pick;
  s.logp = (lsel - m) - logf(sum);
  return s;
}

// Egocentric 7x7 window. kind[vi*7+vj] receives the packed code (CODE_WALL outside the grid); returns the
// 49-bit transparency mask, bit (vj*7 + vi).  `cells` is the env's row-major grid.
template <typename LoadCell>
MERLIN_HD uint64_t gather_view(const EnvState& s, int W, int H, LoadCell load, uint8_t* kind) {
  const int fx = dir_dx(s.dir), fy = dir_dy(s.dir);
  const int rx = -fy, ry = fx;
  uint64_t transp = 0;
#pragma unroll
  for (int vj = 0; vj < kView; ++vj) {
#pragma unroll
    for (int vi = 0; vi < kView; ++vi) {
      const int a = (kView - 1) - vj, b = vi - kView / 2;
      const int wx = s.x + a * fx + b * rx;
      const int wy = s.y + a * fy + b * ry;
      uint32_t code = CODE_WALL;
      if ((unsigned)wx < (unsigned)W && (unsigned)wy < (unsigned)H) code = load(wy * W + wx);
      kind[vi * kView + vj] = (uint8_t)code;
      if (!((M_OPAQUE >> (code & 0xf)) & 1u)) transp |= (uint64_t)1 << (vj * kView + vi);
    }
  }
  return transp;
}

// Grid.process_vis(agent_pos=(3,6)) on row bitmasks; returns the 49-bit visibility mask, bit (vj*7 + vi).
// LITERAL form: the upstream sweeps restated bit by bit (six propagation steps per direction).  It is the reference
// the fast form below is tested against (exhaustively per row); the kernels use the fast form.
MERLIN_HD uint32_t vis_row_literal(uint32_t seed, uint32_t T, uint32_t& lit) {
  uint32_t v = seed;
  const uint32_t TL = T & 0x3f;  // cells 0..5 may push right
#pragma unroll
  for (int k = 0; k < kView - 1; ++k) v |= (v & TL) << 1;
  const uint32_t A = v & TL;
  const uint32_t TR = T & 0x7e;  // cells 6..1 may push left
#pragma unroll
  for (int k = 0; k < kView - 1; ++k) v |= (v & TR) >> 1;
  const uint32_t B = v & TR;
  lit = v;
  return (A | (A << 1) | B | (B >> 1)) & 0x7f;
}

MERLIN_HD uint64_t visibility_literal(uint64_t transp) {
  uint64_t vis = 0;
  uint32_t seed = 1u << (kView / 2);
#pragma unroll
  for (int vj = kView - 1; vj >= 0; --vj) {
    uint32_t v;
    seed = vis_row_literal(seed, (uint32_t)(transp >> (vj * kView)) & 0x7f, v);
    vis |= (uint64_t)v << (vj * kView);
  }
  return vis;
}

// FAST form of one row.  A sweep lights, from every lit transparent cell, the rest of its run of transparent cells in
// the sweep direction plus the cell after the run: an occluded fill, which the carry chain of ONE addition computes
// for all runs at once -- adding the pushers q (a subset of p) to p clears each run from its lowest pusher to its top
// and leaves later pushers set, so fill = (((p + q) ^ p) & p) | q.  The right-to-left sweep is the same fill on the
// bit-reversed row.  7 + 10 operations per row instead of 36; same results for all 128 x 128 (seed, T) pairs (tested).
MERLIN_HD uint32_t fill_up(uint32_t q, uint32_t p) { return (((p + q) ^ p) & p) | q; }
MERLIN_HD uint32_t rev7(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __brev(x) >> 25;
#else
  uint32_t r = 0;
  for (int i = 0; i < 7; ++i) r |= ((x >> i) & 1u) << (6 - i);
  return r;
#endif
}
MERLIN_HD uint32_t vis_row(uint32_t seed, uint32_t T, uint32_t& lit) {
  const uint32_t TL = T & 0x3f;
  uint32_t v = seed | (fill_up(seed & TL, TL) << 1);
  const uint32_t A = v & TL;
  const uint32_t pr = rev7(T) & 0x3f;   // cells 6..1 as bits 0..5
  uint32_t vr = rev7(v);
  vr |= fill_up(vr & pr, pr) << 1;
  v = rev7(vr);
  const uint32_t B = v & T & 0x7e;
  lit = v;
  return (A | (A << 1) | B | (B >> 1)) & 0x7f;
}

MERLIN_HD uint64_t visibility(uint64_t transp) {
  uint64_t vis = 0;
  uint32_t seed = 1u << (kView / 2);
#pragma unroll
  for (int vj = kView - 1; vj >= 0; --vj) {
    uint32_t v;
    seed = vis_row(seed, (uint32_t)(transp >> (vj * kView)) & 0x7f, v);
    vis |= (uint64_t)v << (vj * kView);
  }
  return vis;
}

// The same with one row per BYTE (row vj in bits 8*vj .. 8*vj+6), in and out: what env_kernel_quad's lanes exchange.
MERLIN_HD uint64_t visibility_rows(uint64_t transp_rows) {
  uint64_t vis = 0;
  uint32_t seed = 1u << (kView / 2);
#pragma unroll
  for (int vj = kView - 1; vj >= 0; --vj) {
    uint32_t v;
    seed = vis_row(seed, (uint32_t)(transp_rows >> (8 * vj)) & 0x7f, v);
    vis |= (uint64_t)v << (8 * vj);
  }
  return vis;
}

// Atlas slot shown for the agent cell (view (3,6)): the carried object under the agent triangle.
MERLIN_HD uint32_t agent_kind(uint32_t carry) {
  return carry == 0 ? KIND_AGENT : ((carry & 0x70) | ((carry & 0xf) + 8));  // key/ball/box -> 13/14/15
}

// (type, colour, state) bytes of Grid.encode for a packed code.
MERLIN_HD void sym_of_code(uint32_t code, uint8_t& t, uint8_t& c, uint8_t& st) {
  const uint32_t ty = code & 0xf;
  c = (uint8_t)((code >> 4) & 7);
  if (ty == T_DOOR_CLOSED) { t = 4; st = 1; }
  else if (ty == T_DOOR_LOCKED) { t = 4; st = 2; }
  else { t = (uint8_t)ty; st = 0; }
}

// Chunk lookup for the RGB blit: 16-byte chunk c of the 56x56x3 frame is two 8-byte units; each unit lies in
// one tile row.  Returns cell0 | off0<<8 | cell1<<16 | off1<<24 with cell = vi*7+vj and off = py*3 + part
// (the 8-byte unit index inside the 192-byte tile).
MERLIN_HD uint32_t chunk_lut(int c) {
  uint32_t packed = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int u = 2 * c + h;
    const int row = u / kUnitsPerRow, w = u - row * kUnitsPerRow;
    const int vi = w / 3, part = w - vi * 3;
    const int vj = row / kTile, py = row - vj * kTile;
    packed |= (uint32_t)((vi * kView + vj) | ((py * 3 + part) << 8)) << (16 * h);
  }
  return packed;
}

// Chunk lookup for the BLOCKED frame layout u8[14][14][48] (4x4 pixel blocks, channel-major inside a block): 16-byte
// chunk c = part (c % 3) of block b = c / 3 at (by, bx) = (b / 14, b % 14); the block lies inside cell
// (vi, vj) = (bx / 2, by / 2) as its sub-block (by % 2, bx % 2).  Returns cell | unit16 << 8 with unit16 the 16-byte
// unit inside the blocked 192-byte tile (sub * 3 + part).
MERLIN_HD uint32_t chunk_lut_blocked(int c) {
  const int b = c / 3, part = c - b * 3;
  const int by = b / (kView * 2), bx = b - by * (kView * 2);
  const int vi = bx >> 1, vj = by >> 1, sub = ((by & 1) << 1) | (bx & 1);
  return (uint32_t)(vi * kView + vj) | ((uint32_t)(sub * 3 + part) << 8);
}

// Tile kind shown for a cell of a stored symbolic observation (Grid.encode triple); `agent_cell` = view cell (3, 6).
MERLIN_HD uint32_t kind_of_sym(uint32_t t, uint32_t c, uint32_t st, bool agent_cell) {
  uint32_t code = 0;  // (0,0,0) = not visible
  if (t != 0) {
    uint32_t tt = t & 0xf;
    if (tt == 4) tt = st == 1 ? T_DOOR_CLOSED : (st == 2 ? T_DOOR_LOCKED : T_DOOR_OPEN);
    code = tt | ((c & 7) << 4);
  }
  if (agent_cell) return agent_kind((code & 0xf) == T_EMPTY ? 0u : code);  // the cell under the agent shows what it carries
  return code;
}

}  // namespace merlin
