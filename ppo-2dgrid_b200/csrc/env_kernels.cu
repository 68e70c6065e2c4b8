// env_kernels.cu -- fused MERLIN env kernels for sm_100a.
//
// Every kernel does the same work per environment -- step + wrappers + auto-reset (STEP = 1: actions given; STEP = 2:
// actions drawn in the kernel from the policy's logits, merlin_env_policy_step) or a masked reset (STEP = 0), then gen_obs (+process_vis), the symbolic encode and the RGB frame -- and differs only in how
// environments are mapped onto the machine.  Two phases:
//
//   state phase   step logic, reward shaping, restart, 49-cell window gather, bitmask visibility -> the 49 tile
//                 kinds (and 147 symbolic bytes) of the env, in shared memory.  ~20 warp-instructions per env when
//                 run one env per LANE (state_phase<G>), ~250 when one WARP serves one env cooperatively.
//   frame phase   the 9408-byte frame as 588 coalesced 16-byte streaming stores per env, each assembled from two
//                 8-byte reads of the tile atlas in shared memory (blit_frame); always one warp per env.
//
//   env_kernel<G>        a warp owns G consecutive envs for both phases (G = 32 at scale).  Cheapest in instructions;
//                        the unit of work is G frames (300 KB at G = 32), so it wants many groups per warp.
//   env_kernel_tile<T>   a CTA owns a tile of T envs: warp 0 runs the state phase one env per lane, then ALL warps
//                        of the CTA share the T frames.  Same cheap state phase, 8x finer frame-phase granularity:
//                        fills the SMs from a few thousand envs up and has no tail at a few tiles per CTA.
//   env_kernel_warp      one warp per env, cooperative state phase (two window cells per lane, warp ballots for the
//                        transparency mask).  Lowest latency for batches too small to give each SM a tile.
//
// HBM-bound streaming writers; tensor cores are not involved (there is no contraction on this path).
// Algorithmic HBM bytes per env-step (16x16, RGB): 9408 obs + 256 grid + 32 state + 8 action + 6 = 9710.
#include <cuda_runtime.h>
#include <stdint.h>

#include "env_kernels.cuh"
#include "env_logic.cuh"
#include "obs_swar.cuh"

namespace merlin {

// Streaming (evict-first) 16-byte store.  Measured on B200 at 1M envs: .cs 0.99 of the HBM copy peak, plain / .cg
// stores 0.93, 256-bit st.global.v8.b32 (with or without L2::evict_first) 0.92-0.94.
__device__ __forceinline__ void st_stream_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Grid loads carry an L2 cache policy: evict-LAST when the grid is an entry of the shared layout pool.  The pool
// (16 MB at the benchmark's 65 536 layouts) is re-read by every step while 10 GB of frames stream through the same L2
// between two uses of a line; tagged evict-last it stays resident.  Measured on B200, 1M envs, RGB, 65 536 layouts:
// 1.484 -> 1.418 ms per step (7.07e8 -> 7.39e8 env-steps/s, +4.6 %); with the 8192-layout pool of round 1 +0.9 %
// (profiles/r02_pool_evict_last_ab.txt).  Private (mutable) grids -- 256 B per env, no reuse across envs -- keep the
// normal policy.
__device__ __forceinline__ uint64_t grid_policy(bool shared_pool) {
  uint64_t last, normal;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(last));
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(normal));
  return shared_pool ? last : normal;
}
__device__ __forceinline__ uint32_t ld_cell(const uint8_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

constexpr int kWarpKindStride = 64;   // per-warp slot for the 49 tile kinds of one env

// bit `tile` of the 128-bit "present" mask (four device words, read through the read-only path; they live in device
// memory rather than in the kernel parameters so that a captured CUDA graph sees a re-uploaded layout pool's mask)
__device__ __forceinline__ bool tile_bit(const uint32_t* m, int tile) {
  return (__ldg(m + (tile >> 5)) >> (tile & 31)) & 1u;
}

// Can a closed / locked door occur in a grid of this handle?  Read from the device-resident "present" mask (codes
// type | colour << 4 with type 11 / 12; mutable grids set every bit), so a CUDA graph captured before a re-upload
// still sees the current pool.  Without doors, walls are the only opaque cells and Grid.encode needs no state byte:
// the row-parallel observation takes its short path (warp-uniform branch).
__device__ __forceinline__ bool pool_has_doors(const uint32_t* present) {
  constexpr uint32_t kDoorBits = (1u << T_DOOR_CLOSED) | (1u << T_DOOR_LOCKED);
  constexpr uint32_t kMask = kDoorBits | (kDoorBits << 16);   // colours 2w and 2w + 1 share word w
  return ((__ldg(present) | __ldg(present + 1) | __ldg(present + 2) | __ldg(present + 3)) & kMask) != 0;
}

struct Flags {
  int n_actions;
  bool mutable_grid, stuck_on, explore_on, auto_reset, advance, want_rgb, want_sym;
  bool doors = true;   // set by the kernels that run the row-parallel observation
  uint64_t pol;        // L2 cache policy of the grid loads (grid_policy)
  __device__ __forceinline__ explicit Flags(const EnvParams& p)
      : n_actions((p.flags & MERLIN_F_SEVEN_ACTIONS) ? 7 : 3), mutable_grid(p.cells != nullptr),
        stuck_on(p.flags & MERLIN_F_STUCK_PENALTY), explore_on(p.flags & MERLIN_F_EXPLORE_BONUS),
        auto_reset(p.flags & MERLIN_F_AUTO_RESET), advance(!(p.flags & MERLIN_F_RESET_SAME)),
        want_rgb(p.obs_rgb != nullptr), want_sym(p.obs_sym != nullptr), pol(grid_policy(p.cells == nullptr)) {}
};

// The action env e takes this step: read from `actions`, or -- policy I/O -- drawn here from the policy's logits
// (sample_policy, env_logic.cuh).  `commit`: this thread stores the action / log-probability / value rows and advances
// the env's draw counter (one lane per env in the lane-per-env kernels, lane 0 in the warp-per-env kernel, where every
// lane computes the same sample from the same warp-uniform loads).
struct ActionDraw {
  long long action;
  float logp;
  uint32_t draw;
};
// The step kernels are instantiated three times: STEP = 0 (masked reset), 1 (step, actions given), 2 (step, actions
// drawn here from the policy's logits).  The policy code exists only in the STEP = 2 instances: compiled into the
// others -- even behind a uniform branch, even out of line -- it degrades the register allocation of the frame kernels'
// state phase, which runs under a 128-register cap (B200, 1M envs: 1461 us per step vs 1434 us without it;
// profiles/r02_tile_kernel_ab.txt).
template <bool POLICY>
__device__ __forceinline__ ActionDraw draw_action(const EnvParams& p, int n_actions, int e) {
  ActionDraw d;
  d.logp = 0.f;
  d.draw = 0;
  if (!POLICY) {
    d.action = p.actions[e];
    return d;
  }
  float lg[kMaxActions];
  const float* row = p.logits + (size_t)e * p.logits_stride;
#pragma unroll
  for (int a = 0; a < kMaxActions; ++a) lg[a] = a < n_actions ? row[a] : 0.f;
  float u = 0.f;
  if (!p.greedy) {
    d.draw = p.draws[e];
    u = sampler_uniform(p.seed_lo, p.seed_hi, (uint32_t)e, d.draw);
  }
  const PolicySample smp = sample_policy(lg, n_actions, u, p.greedy != 0);
  d.action = smp.action;
  d.logp = smp.logp;
  return d;
}
template <bool POLICY>
__device__ __forceinline__ void commit_action(const EnvParams& p, int e, const ActionDraw& d) {
  if (!POLICY) return;
  if (!p.greedy) p.draws[e] = d.draw + 1u;
  p.out_action[e] = d.action;
  p.out_logp[e] = d.logp;
  if (p.out_value) p.out_value[e] = p.value_in[(size_t)e * p.value_stride];
}
// First-episode record of deterministic evaluation (see EnvParams::rec_finished).
template <bool POLICY>
__device__ __forceinline__ void record_first_episode(const EnvParams& p, int e, bool done, bool goal, float ep_ret, int len) {
  if (!POLICY || p.rec_finished == nullptr || !done || p.rec_finished[e]) return;
  p.rec_finished[e] = 1;
  p.rec_return[e] = ep_ret;
  p.rec_length[e] = len;
  p.rec_goal[e] = goal ? 1 : 0;
}

// Stage the tile atlas: only the slots a frame of this handle's layout pool can show (5 of 128 for the MERLIN
// scenarios: 960 B instead of 24 KB), at their usual offsets.  All threads of the CTA take part.
__device__ __forceinline__ void stage_atlas(const EnvParams& p, uint8_t* atlas_s) {
  const int4* src = reinterpret_cast<const int4*>(p.atlas);
  int4* dst = reinterpret_cast<int4*>(atlas_s);
  for (int i = threadIdx.x; i < kAtlasBytes / 16; i += blockDim.x) {
    const int tile = i / (kTileBytes / 16);
    if (tile_bit(p.tile_present, tile)) dst[i] = __ldg(src + i);
  }
}

// Per-lane blit map: chunk c = lane + 32*k -> (cell0, off0, cell1, off1), from the table built at handle creation.
__device__ __forceinline__ void load_lut(const EnvParams& p, int lane, uint32_t (&lut)[kChunksPerLane]) {
#pragma unroll
  for (int k = 0; k < kChunksPerLane; ++k) lut[k] = __ldg(p.blit_lut + k * 32 + lane);
}

// Frame phase for one env: `kp` = its 49 tile kinds in shared memory.
__device__ __forceinline__ void blit_frame(const uint8_t* atlas_s, const uint8_t* kp, const uint32_t (&lut)[kChunksPerLane],
                                           uint8_t* frame, int lane) {
  const uint2* atlas64 = reinterpret_cast<const uint2*>(atlas_s);
#pragma unroll
  for (int k = 0; k < kChunksPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < kChunks) {
      const uint32_t q = lut[k];
      const uint32_t k0 = kp[q & 0xff], k1 = kp[(q >> 16) & 0xff];
      const uint2 a = atlas64[k0 * (kTileBytes / 8) + ((q >> 8) & 0xff)];
      const uint2 b = atlas64[k1 * (kTileBytes / 8) + (q >> 24)];
      st_stream_v4(frame + c * 16, a.x, a.y, b.x, b.y);
    }
  }
}

// Symbolic rows of `n_here` consecutive envs from shared memory (same layout as the output), by `nthreads` threads.
__device__ __forceinline__ void emit_sym_rows(uint8_t* out, const uint8_t* sym_s, int n_here, unsigned render_mask,
                                              int tid, int nthreads) {
  const unsigned full = n_here >= 32 ? 0xffffffffu : ((1u << n_here) - 1u);
  if (render_mask == full && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (n_here * kSymBytes) % 16 == 0) {
    const int4* src = reinterpret_cast<const int4*>(sym_s);
    int4* dst = reinterpret_cast<int4*>(out);
    for (int i = tid; i < n_here * kSymBytes / 16; i += nthreads) dst[i] = src[i];
  } else {
    for (int b = tid; b < n_here * kSymBytes; b += nthreads) {
      const int i = b / kSymBytes;
      if ((render_mask >> i) & 1) out[b] = sym_s[b];
    }
  }
}

// One env's 147-byte symbolic image from its seven visible-code groups (obs_swar.cuh) into shared memory at `row`
// (any alignment: rows of consecutive lanes are 147 bytes apart).  The image is assembled as 37 little-endian words in
// registers, funnel-shifted by this lane's misalignment and stored as 35 aligned words; the words that straddle the
// row's two ends are shared with the neighbouring lanes' rows and go out as single bytes.
__device__ __forceinline__ void store_sym_row(uint8_t* row, const uint64_t (&g)[kView], bool doors) {
  uint32_t r[38];
  {
    uint64_t acc = 0;
    int have = 0, n = 0;  // compile-time after unrolling: every shift below is an immediate
#pragma unroll
    for (int vi = 0; vi < kView; ++vi) {
      uint32_t w[6];
      encode_group(g[vi], w, doors);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        acc |= (uint64_t)w[i] << (8 * have);
        have += i < 5 ? 4 : 1;
        if (have >= 4) { r[n++] = (uint32_t)acc; acc >>= 32; have -= 4; }
      }
    }
    r[n++] = (uint32_t)acc;  // bytes 144..146 (+ one zero)
    r[n] = 0;                // n == 37
  }
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(row);
  const uint32_t s = addr & 3u, sh = 8u * s;
  uint8_t* base = row - s;  // word-aligned
  uint32_t prev = 0;
#pragma unroll
  for (int k = 0; k < 38; ++k) {
    const uint32_t word = __funnelshift_l(prev, r[k], sh);  // bytes 4k - s .. 4k - s + 3 of the image
    prev = r[k];
    if (k >= 1 && k <= 35) {
      *reinterpret_cast<uint32_t*>(base + 4 * k) = word;
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = 4 * k + b - (int)s;  // image byte held by byte b of this word
        if (i >= 0 && i < kSymBytes) base[4 * k + b] = (uint8_t)(word >> (8 * b));
      }
    }
  }
}

// State phase, one env per lane, for the G envs e0 .. e0+G-1 (lanes >= G idle).  Must be called by a full warp.
// Leaves kinds_s[lane][kKindStride] / sym_s[lane][147] filled for the envs whose bit is set in the returned mask.
// SWAR = true: the observation is computed on window rows (obs_swar.cuh; needs W >= 7), else cell by cell.
template <int G, int STEP, bool SWAR = false>
__device__ __forceinline__ unsigned state_phase(const EnvParams& p, const Flags& f, int e0, int lane, uint8_t* kinds_s,
                                                uint8_t* sym_s) {
  const int e = e0 + lane;
  const bool active = lane < G && e < p.N;
  EnvState s{};
  bool restart = false;   // this lane's env (re)loads a layout now
  bool render = active;   // this lane's env gets its observation written
  if (active) {
    const int4 st = p.state[e];
    unpack_state(st.x, st.y, st.z, st.w, s);
  }
  float ep_ret = active ? p.ep_return[e] : 0.f;

  if (STEP) {
    if (active) {
      const uint8_t* grid = f.mutable_grid ? p.cells + (size_t)e * p.cell_stride
                                           : p.pool_cells + (size_t)s.layout * p.cell_stride;
      const int fx = s.x + dir_dx(s.dir), fy = s.y + dir_dy(s.dir);
      const bool inb = (unsigned)fx < (unsigned)p.W && (unsigned)fy < (unsigned)p.H;
      const int fidx = fy * p.W + fx;
      const uint32_t fwd = inb ? ld_cell(grid + fidx, f.pol) : CODE_WALL;
      const ActionDraw act = draw_action<STEP == 2>(p, f.n_actions, e);
      commit_action<STEP == 2>(p, e, act);
      StepResult r = step_logic(s, act.action, f.n_actions, fwd, inb, fidx, p.max_steps);
      if (r.bad_action) atomicAdd(p.bad_actions, 1ull);
      if (r.write_idx >= 0 && f.mutable_grid) p.cells[(size_t)e * p.cell_stride + r.write_idx] = (uint8_t)r.write_code;

      uint32_t vword = 0;
      const int cell = s.y * p.W + s.x;
      uint32_t* vptr = nullptr;
      if (f.explore_on) { vptr = p.visited + (size_t)e * p.vis_words + (cell >> 5); vword = *vptr; }
      bool stuck = false;
      const uint32_t vword_in = vword;
      const double rew_d = shape_reward(s, r.reward, f.stuck_on, p.stuck_max_stay, p.stuck_penalty, f.explore_on,
                                        p.explore_bonus, vword, cell & 31, stuck);
      if (f.explore_on && vword != vword_in) *vptr = vword;
      const float rew = (float)rew_d;
      ep_ret += rew;
      const bool done = r.terminated || r.truncated;
      p.reward[e] = rew;
      p.terminated[e] = r.terminated ? 1 : 0;
      p.truncated[e] = r.truncated ? 1 : 0;
      if (p.out_ep_return) p.out_ep_return[e] = done ? ep_ret : 0.f;
      if (p.out_ep_length) p.out_ep_length[e] = done ? s.step_count : 0;
      if (p.out_stuck) p.out_stuck[e] = stuck ? 1 : 0;
      if (p.out_done) p.out_done[e] = done ? 1.f : 0.f;
      record_first_episode<STEP == 2>(p, e, done, r.terminated && r.reward > 0.0, ep_ret, s.step_count);
      restart = done && f.auto_reset;
    }
  } else {
    restart = active && (p.reset_mask == nullptr || p.reset_mask[e] != 0);
    render = restart;
  }

  // (re)start: pose from the pool, counters cleared, cursor advanced; mutable grids / visited maps are
  // re-initialised by the whole warp with coalesced copies
  const unsigned restart_mask = __ballot_sync(0xffffffffu, restart);
  if (restart_mask) {
    // pool index this env loads: pending first layout, else the next (PPO) or the same (FOMAML) one
    const int load_cur = s.layout < 0 ? ~s.layout
                                      : (f.advance ? (int)(((unsigned)s.layout + (unsigned)p.cursor_stride) % (unsigned)p.n_layouts) : s.layout);
    if (restart) {
      const uint32_t a = p.pool_agent[load_cur];
      s.x = a & 0xff; s.y = (a >> 8) & 0xff; s.dir = (a >> 16) & 3; s.carry = 0;
      s.step_count = 0; s.stay = 0; s.last_x = s.x; s.last_y = s.y;
      ep_ret = 0.f;
    }
    if (f.mutable_grid || f.explore_on) {
      unsigned m = restart_mask;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int cur = __shfl_sync(0xffffffffu, load_cur, src);
        const int sx = __shfl_sync(0xffffffffu, s.x, src), sy = __shfl_sync(0xffffffffu, s.y, src);
        const size_t ee = (size_t)(e0 + src);
        if (f.mutable_grid) {
          const int4* from = reinterpret_cast<const int4*>(p.pool_cells + (size_t)cur * p.cell_stride);
          int4* to = reinterpret_cast<int4*>(p.cells + ee * p.cell_stride);
          for (int i = lane; i < p.cell_stride / 16; i += 32) to[i] = from[i];
        }
        if (f.explore_on) {
          const int cell = sy * p.W + sx;
          for (int i = lane; i < p.vis_words; i += 32)
            p.visited[ee * p.vis_words + i] = (i == (cell >> 5)) ? (1u << (cell & 31)) : 0u;
        }
      }
      __syncwarp();
    }
    if (restart) s.layout = load_cur;
  }

  if (active && (STEP || restart)) {
    int4 st;
    pack_state(s, st.x, st.y, st.z, st.w);
    p.state[e] = st;
    p.ep_return[e] = ep_ret;
  }

  // observation, part 1 (per lane): window gather -> visibility -> tile kinds (+ symbolic bytes) in smem
  if ((f.want_rgb || f.want_sym) && render) {
    const uint8_t* grid = f.mutable_grid ? p.cells + (size_t)e * p.cell_stride
                                         : p.pool_cells + (size_t)s.layout * p.cell_stride;
    uint8_t* kind = kinds_s + lane * kKindStride;
    if (SWAR) {
      uint64_t g[kView], seen[kView];
      observe_swar(s, grid, p.W, p.H, g, seen, f.doors, f.pol);
      if (f.want_rgb) {
        uint32_t kw[13];
        kind_words(g, s.carry, kw);
        uint32_t* kdst = reinterpret_cast<uint32_t*>(kind);  // kKindStride = 52: word-aligned rows
#pragma unroll
        for (int i = 0; i < 13; ++i) kdst[i] = kw[i];
      }
      if (f.want_sym) store_sym_row(sym_s + lane * kSymBytes, g, f.doors);
    } else {
      const uint64_t pol = f.pol;
      const uint64_t transp = gather_view(s, p.W, p.H, [&](int idx) -> uint32_t { return ld_cell(grid + idx, pol); }, kind);
      const uint64_t vis = visibility(transp);
      uint8_t* sym = sym_s + lane * kSymBytes;
#pragma unroll
      for (int vi = 0; vi < kView; ++vi) {
#pragma unroll
        for (int vj = 0; vj < kView; ++vj) {
          const int c = vi * kView + vj;
          const bool seen = (vis >> (vj * kView + vi)) & 1;
          uint32_t code = kind[c];
          const bool agent_cell = (vi == kView / 2 && vj == kView - 1);
          if (agent_cell) code = s.carry ? s.carry : CODE_EMPTY;
          kind[c] = (uint8_t)(agent_cell ? agent_kind(s.carry) : (seen ? code : KIND_UNSEEN));
          if (f.want_sym) {
            uint8_t t = 0, col = 0, stt = 0;
            if (seen) sym_of_code(code, t, col, stt);
            sym[c * 3 + 0] = t; sym[c * 3 + 1] = col; sym[c * 3 + 2] = stt;
          }
        }
      }
    }
  }
  __syncwarp();
  return __ballot_sync(0xffffffffu, render);
}

// ---------------------------------------------------------------------------------------------------
// env_kernel<G, STEP>: a warp owns G consecutive envs.
template <int G, int STEP>
__global__ void __launch_bounds__(kThreads, 1) env_kernel(const EnvParams p, const int n_groups) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const Flags f(p);
  uint8_t* atlas_s = smem;
  uint8_t* warp_s = smem + kAtlasBytes + warp * warp_smem_bytes(G);
  uint8_t* kinds_s = warp_s;                       // [G][kKindStride]
  uint8_t* sym_s = warp_s + G * kKindStride;       // [G][147] contiguous, same layout as the output rows

  uint32_t lut[kChunksPerLane];
  if (f.want_rgb) {
    stage_atlas(p, atlas_s);
    load_lut(p, lane, lut);
  }
  __syncthreads();

  const int warps_per_cta = blockDim.x >> 5;
  for (int g = blockIdx.x * warps_per_cta + warp; g < n_groups; g += gridDim.x * warps_per_cta) {
    const int e0 = g * G;
    const unsigned render_mask = state_phase<G, STEP>(p, f, e0, lane, kinds_s, sym_s);
    if (f.want_sym && render_mask)
      emit_sym_rows(p.obs_sym + (size_t)e0 * kSymBytes, sym_s, min(G, p.N - e0), render_mask, lane, 32);
    if (f.want_rgb) {
      unsigned m = render_mask;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        blit_frame(atlas_s, kinds_s + i * kKindStride, lut, p.obs_rgb + (size_t)(e0 + i) * kImgBytes, lane);
      }
    }
    __syncwarp();  // smem rows are reused by the next group
  }
}

// ---------------------------------------------------------------------------------------------------
// env_kernel_ordered<G, STEP, THREADS, MINB>: env_kernel's mapping (a warp owns G consecutive envs for both phases)
// with the groups handed out IN ORDER, one ticket per warp, from the same self-resetting counter the tile kernel
// uses.  Every warp of the SM runs its own state phase (no warp-0 bottleneck, no CTA barrier in the loop): the
// dependent-load chain of a state phase (~10 us) is hidden by all resident warps instead of one per CTA, while the
// write fronts still advance together.  The next ticket is drawn before the work so its latency is off the path.
// MEASURED AND NOT ADOPTED (B200, 1M envs, RGB, fraction of the HBM copy peak; env_kernel_tile<16>: 1.08):
//   G=8, 128 thr x 3 CTAs  1.046    G=16, 128 x 3  1.045    G=32, 128 x 3  1.041    G=8, 128 x 4 (spills)  1.033
//   G=8, 256 x 2  1.034    G=16, 256 x 2  1.024    G=16, 128 x 4  1.021    G=4, 128 x 4  0.920
// i.e. in-order hand-out lifts the group mapping from 0.92 (static assignment) to 1.05, but more concurrent state phases
// buy nothing over the tile kernel: its limit is the store stream, not the state-phase chain.  Kept selectable (choice 6).
template <int G, int STEP, int THREADS, int MINB, bool SWAR = false>
__global__ void __launch_bounds__(THREADS, MINB) env_kernel_ordered(const EnvParams p, const int n_groups) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  Flags f(p);
  if (SWAR) f.doors = pool_has_doors(p.tile_present);
  uint8_t* atlas_s = smem;
  uint8_t* warp_s = smem + kAtlasBytes + warp * warp_smem_bytes(G);
  uint8_t* kinds_s = warp_s;
  uint8_t* sym_s = warp_s + G * kKindStride;

  uint32_t lut[kChunksPerLane];
  if (f.want_rgb) {
    stage_atlas(p, atlas_s);
    load_lut(p, lane, lut);
  }
  __syncthreads();

  int g = 0;
  if (lane == 0) g = (int)atomicAdd(&p.sched[0], 1u);
  g = __shfl_sync(0xffffffffu, g, 0);
  while (g < n_groups) {
    int next = 0;
    if (lane == 0) next = (int)atomicAdd(&p.sched[0], 1u);
    const int e0 = g * G;
    const unsigned render_mask = state_phase<G, STEP, SWAR>(p, f, e0, lane, kinds_s, sym_s);
    if (f.want_sym && render_mask)
      emit_sym_rows(p.obs_sym + (size_t)e0 * kSymBytes, sym_s, min(G, p.N - e0), render_mask, lane, 32);
    if (f.want_rgb) {
      unsigned m = render_mask;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        blit_frame(atlas_s, kinds_s + i * kKindStride, lut, p.obs_rgb + (size_t)(e0 + i) * kImgBytes, lane);
      }
    }
    __syncwarp();  // smem rows are reused by the next group
    g = __shfl_sync(0xffffffffu, next, 0);
  }
  // every warp of the grid has drawn its last ticket once all of them have passed here: the last one rearms
  if (lane == 0 && atomicAdd(&p.sched[1], 1u) == gridDim.x * (THREADS / 32) - 1) {
    p.sched[0] = 0;
    p.sched[1] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// env_kernel_tile<T, STEP>: a CTA owns a tile of T <= 32 consecutive envs; warp 0 runs the state phase, every
// warp of the CTA takes frames of the tile; co-resident CTAs overlap one tile's state phase with others' frames.
// Shape (B200, 1M envs, RGB), all with tiles handed out in order (see the ticket scheduler below):
//   T=16, 128 threads x 4 CTAs/SM   1.08 of the measured HBM copy peak (7.3e8 env-steps/s)   <- used
//   T=32, 128 x 3: 1.08    T=32, 256 x 2: 1.06    T=32, 128 x 4: 1.06    T=16, 64 x 6: 1.07    T=8, 64 x 8: 1.04
//   T=16, 128 x 5 (96 registers): 1.03    T=16, 128 x 6 (80 registers): 0.95    T=16, 256 x 2: 0.87
// With the static `tile += gridDim.x` assignment the best shape (T=32, 256 x 2) reached 0.99 and every other one
// 0.69-0.96.  Also slower: the blit map in shared memory (-7 %), overlapping the next tile's state phase inside the
// CTA, all warps writing ONE frame at a time, plain / .cg / 256-bit stores instead of st.global.cs.v4, and a separate
// state kernel + high-occupancy frame kernel with static assignment.  Padding the shared-memory atlas slots to remove
// the 17 % bank conflicts changes nothing (the store stream, not the LSU, is the limit).  A plain vectorised fill reaches 7.4-7.6 TB/s on
// this part and frames streamed in order without any env logic 7.4 TB/s (tools/cuda/write_pattern_bench.cu): the fused
// kernel's 7.1 TB/s is 95 % of that.
constexpr int kTileThreads = 128;
constexpr int kTileCtasPerSm = 4;
__host__ __device__ constexpr int tile_smem_bytes(int T) {
  return kAtlasBytes + ((T * kKindStride + T * kSymBytes + 15) & ~15) + 16;
}

template <int T, int STEP, int THREADS = kTileThreads, int MINB = kTileCtasPerSm, bool SWAR = false>
__global__ void __launch_bounds__(THREADS, MINB) env_kernel_tile(const EnvParams p, const int n_tiles) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  Flags f(p);
  if (SWAR) f.doors = pool_has_doors(p.tile_present);
  uint8_t* atlas_s = smem;
  uint8_t* kinds_s = smem + kAtlasBytes;   // [T][kKindStride]
  uint8_t* sym_s = kinds_s + T * kKindStride;              // [T][147]
  unsigned* mask_s = reinterpret_cast<unsigned*>(smem + tile_smem_bytes(T) - 16);

  uint32_t lut[kChunksPerLane];
  if (f.want_rgb) {
    stage_atlas(p, atlas_s);
    load_lut(p, lane, lut);
  }

  // Tiles are handed out IN ORDER from a ticket counter rather than round-robin by CTA index: the CTAs' write fronts
  // then stay inside one narrow, advancing window of the observation buffer, which is what keeps HBM writes near the
  // plain-fill rate (measured with tools/cuda/write_pattern_bench.cu: 7.4 TB/s in order vs 6.4 TB/s with the static
  // `tile += gridDim.x` assignment, whose CTAs drift apart).  Thread 32 draws the next ticket while warp 0 runs the
  // state phase; the last CTA to finish rearms the counters for the next launch (also under CUDA-graph replay).
  __shared__ int s_next;
  if (threadIdx.x == 0) s_next = (int)atomicAdd(&p.sched[0], 1u);
  __syncthreads();
  int tile = s_next;
  while (tile < n_tiles) {
    const int e0 = tile * T;
    if (warp == 0) {
      const unsigned m = state_phase<T, STEP, SWAR>(p, f, e0, lane, kinds_s, sym_s);
      if (lane == 0) *mask_s = m;
    } else if (threadIdx.x == 32) {
      s_next = (int)atomicAdd(&p.sched[0], 1u);
    }
    __syncthreads();  // kinds / sym / mask of this tile, the next ticket (and, first time round, the atlas) are in smem
    const unsigned render_mask = *mask_s;
    const int next = s_next;
    if (f.want_sym && render_mask)
      emit_sym_rows(p.obs_sym + (size_t)e0 * kSymBytes, sym_s, min(T, p.N - e0), render_mask, threadIdx.x, blockDim.x);
    if (f.want_rgb) {
      for (int i = warp; i < T; i += warps_per_cta)
        if ((render_mask >> i) & 1)
          blit_frame(atlas_s, kinds_s + i * kKindStride, lut, p.obs_rgb + (size_t)(e0 + i) * kImgBytes, lane);
    }
    __syncthreads();  // the tile buffers and the ticket slot are rewritten in the next round
    tile = next;
  }
  if (threadIdx.x == 0 && atomicAdd(&p.sched[1], 1u) == gridDim.x - 1) {
    p.sched[0] = 0;  // every CTA has drawn its last ticket: safe to rearm
    p.sched[1] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// env_kernel_tile_tma<T, STEP, THREADS, MINB, NBUF>: env_kernel_tile with the frame phase routed through the TMA unit.
// A warp assembles a frame in one of its NBUF shared-memory staging buffers (the same atlas reads, 16-byte shared
// stores instead of global ones), makes it visible to the async proxy and ONE lane issues a single 9408-byte
// cp.async.bulk.global.shared::cta for it; the buffer is reused once its bulk group has been read.  The copy engine
// streams whole frames to HBM while warp 0 is already in the next tile's state phase, and the LSU no longer carries
// the 10 GB/launch store stream (ncu on env_kernel_tile: L1/TEX 78 % busy next to 85 % DRAM).
// MEASURED AND NOT ADOPTED (B200, 1M envs, RGB, fraction of the HBM copy peak; env_kernel_tile<16>: 1.08):
//   T=32, 128 thr x 3 CTAs, 1 buffer/warp  0.96      T=32, 128 x 2, 2 buffers  0.87      T=16, 128 x 3, 1 buffer  0.82
//   T=32, 256 x 2, 1 buffer  0.82    T=16, 64 x 3, 2 buffers  0.78    T=16, 128 x 2, 2 buffers  0.70    T=16, 256 x 1, 2 buffers  0.48
// The copy engine itself is not the problem -- tools/cuda/tma_store_bench.cu streams staged frames at 7.56-7.60 TB/s
// with as little as ONE 64-thread CTA per SM (per-lane st.global.cs.v4: 7.49) -- but every frame byte now crosses
// shared memory three times (atlas read, staging write, engine read) instead of once, and the staging buffers
// (9.4 KB per frame in flight) cost a resident CTA per SM.  Kept selectable (kernel choice 4) with its parity tests.
__device__ __forceinline__ void bulk_store_frame(uint8_t* gdst, const uint8_t* ssrc) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "n"(kImgBytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}

// Frame phase into shared memory: `stage` receives the frame exactly as blit_frame would write it to global memory.
__device__ __forceinline__ void blit_frame_smem(const uint8_t* atlas_s, const uint8_t* kp, const uint32_t (&lut)[kChunksPerLane],
                                                uint8_t* stage, int lane) {
  const uint2* atlas64 = reinterpret_cast<const uint2*>(atlas_s);
#pragma unroll
  for (int k = 0; k < kChunksPerLane; ++k) {
    const int c = lane + 32 * k;
    if (c < kChunks) {
      const uint32_t q = lut[k];
      const uint32_t k0 = kp[q & 0xff], k1 = kp[(q >> 16) & 0xff];
      const uint2 a = atlas64[k0 * (kTileBytes / 8) + ((q >> 8) & 0xff)];
      const uint2 b = atlas64[k1 * (kTileBytes / 8) + (q >> 24)];
      *reinterpret_cast<uint4*>(stage + c * 16) = make_uint4(a.x, a.y, b.x, b.y);
    }
  }
}

__host__ __device__ constexpr int tile_tma_smem_bytes(int T, int threads, int nbuf) {
  return tile_smem_bytes(T) + (threads / 32) * nbuf * kImgBytes;
}

template <int T, int STEP, int THREADS, int MINB, int NBUF>
__global__ void __launch_bounds__(THREADS, MINB) env_kernel_tile_tma(const EnvParams p, const int n_tiles) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int warps_per_cta = THREADS / 32;
  const Flags f(p);
  uint8_t* atlas_s = smem;
  uint8_t* kinds_s = smem + kAtlasBytes;
  uint8_t* sym_s = kinds_s + T * kKindStride;
  unsigned* mask_s = reinterpret_cast<unsigned*>(smem + tile_smem_bytes(T) - 16);
  uint8_t* stage_s = smem + tile_smem_bytes(T) + warp * NBUF * kImgBytes;   // this warp's NBUF frame buffers

  uint32_t lut[kChunksPerLane];
  if (f.want_rgb) {
    stage_atlas(p, atlas_s);
    load_lut(p, lane, lut);
  }
  int buf = 0;
  __shared__ int s_next;
  if (threadIdx.x == 0) s_next = (int)atomicAdd(&p.sched[0], 1u);
  __syncthreads();
  int tile = s_next;
  while (tile < n_tiles) {
    const int e0 = tile * T;
    if (warp == 0) {
      const unsigned m = state_phase<T, STEP>(p, f, e0, lane, kinds_s, sym_s);
      if (lane == 0) *mask_s = m;
    } else if (threadIdx.x == 32) {
      s_next = (int)atomicAdd(&p.sched[0], 1u);
    }
    __syncthreads();
    const unsigned render_mask = *mask_s;
    const int next = s_next;
    if (f.want_sym && render_mask)
      emit_sym_rows(p.obs_sym + (size_t)e0 * kSymBytes, sym_s, min(T, p.N - e0), render_mask, threadIdx.x, blockDim.x);
    if (f.want_rgb) {
      for (int i = warp; i < T; i += warps_per_cta) {
        if (!((render_mask >> i) & 1)) continue;
        uint8_t* stage = stage_s + buf * kImgBytes;
        if (lane == 0) bulk_wait_read<NBUF - 1>();   // the bulk group that last read this buffer is done with it
        __syncwarp();
        blit_frame_smem(atlas_s, kinds_s + i * kKindStride, lut, stage, lane);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the copy engine
        __syncwarp();
        if (lane == 0) bulk_store_frame(p.obs_rgb + (size_t)(e0 + i) * kImgBytes, stage);
        buf = (buf + 1 == NBUF) ? 0 : buf + 1;
      }
    }
    __syncthreads();
    tile = next;
  }
  if (lane == 0) bulk_wait_read<0>();   // shared memory must outlive the engine's reads
  if (threadIdx.x == 0 && atomicAdd(&p.sched[1], 1u) == gridDim.x - 1) {
    p.sched[0] = 0;
    p.sched[1] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// env_kernel_sym<STEP>: symbolic observations only (no frame phase): one warp per 32 envs, plain grid.  Without the
// blit map and the frame loop the state phase fits in ~72 registers, so six to seven 128-thread CTAs are resident per
// SM instead of one 256-thread CTA of env_kernel<32>: this mode is instruction/latency-bound (449 B per env-step), and
// occupancy is what it needs.
__host__ __device__ constexpr int sym_kernel_warp_smem(bool swar) {
  return swar ? 32 * kSymBytes : warp_smem_bytes(32);
}

#ifndef MERLIN_SYM_MINB
#define MERLIN_SYM_MINB 10  // 48 registers, 40 warps/SM: 1.16e10 env-steps/s at 1M envs (unconstrained, 56 registers: 1.11e10; 12 CTAs, 40 registers + spills: 1.04e10)
#endif
template <int STEP, bool SWAR = false>
__global__ void __launch_bounds__(128, SWAR ? MERLIN_SYM_MINB : 1) env_kernel_sym(const EnvParams p, const int n_groups) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  Flags f(p);
  f.want_rgb = false;  // no frame phase here: tile kinds are not produced either
  if (SWAR) f.doors = pool_has_doors(p.tile_present);
  // the row-parallel form keeps the window in registers; only the per-cell form stages the 49 codes per env
  uint8_t* warp_s = smem + warp * sym_kernel_warp_smem(SWAR);
  uint8_t* kinds_s = SWAR ? nullptr : warp_s;
  uint8_t* sym_s = SWAR ? warp_s : warp_s + 32 * kKindStride;
  const int g = blockIdx.x * (blockDim.x >> 5) + warp;
  if (g >= n_groups) return;
  const int e0 = g * 32;
  const unsigned render_mask = state_phase<32, STEP, SWAR>(p, f, e0, lane, kinds_s, sym_s);
  if (f.want_sym && render_mask)
    emit_sym_rows(p.obs_sym + (size_t)e0 * kSymBytes, sym_s, min(32, p.N - e0), render_mask, lane, 32);
}

// Observation path.  Measured on B200 at 1M envs: the row-parallel form (obs_swar.cuh, 30 % fewer instructions, 48
// instead of 72 registers) lifts the symbolic-only kernel from 7.6e9 to 1.19e10 env-steps/s (0.52 -> 0.82 of the HBM
// roofline of its 449 B/step): that kernel is ALU-bound and runs 40 warps per SM.  The frame kernels are indifferent
// (tile kernel 1.083 per-cell vs 1.076 row form, ordered-group kernel 1.05 either way): their limit is the store
// stream.  (With the window loads behind per-row branches the tile kernel dropped to 1.01 -- its single state-phase
// warp paid one L2 round trip per row; window_rows() is straight-line for that reason.)  Hence: 0 = automatic =
// row-parallel in the symbolic-only kernel only, 1 = per-cell everywhere, 2 = row-parallel in every kernel that has it
// (symbolic-only, tile, ordered; tests, A/B).
static bool use_swar(const EnvParams& p, const LaunchCtx& ctx, bool frame_kernel) {
  if (p.W < kView || ctx.observation_path == 1) return false;
  return ctx.observation_path == 2 || !frame_kernel;
}

template <int STEP>
static cudaError_t launch_sym_kernel(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  constexpr int threads = 128, warps = threads / 32;
  const int n_groups = (p.N + 31) / 32;
  const int grid = (n_groups + warps - 1) / warps;
  if (use_swar(p, ctx, false))
    env_kernel_sym<STEP, true><<<grid, threads, warps * sym_kernel_warp_smem(true), stream>>>(p, n_groups);
  else
    env_kernel_sym<STEP, false><<<grid, threads, warps * sym_kernel_warp_smem(false), stream>>>(p, n_groups);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// env_kernel_warp<STEP, PDL, LEAN>: one WARP per environment -- the small-batch mapping (N <= 24 576: every env of the
// batch is resident at once or nearly so, and the run time is launch latency + one env's dependent chain + its stores).
//   * state / action / forward cell are loaded at warp-uniform addresses (one broadcast transaction each) and the
//     step logic runs redundantly on all lanes; lane 0 alone writes the per-env outputs and the new state.  The first
//     env's state and action are requested BEFORE the atlas is staged, so both round trips overlap.
//   * LEAN (three actions, no shaping wrapper => immutable grids; launch_warp_kernel): the step needs TWO dependent
//     round trips instead of three -- see the block at `if (LEAN)`.
//   * the 49-cell window is gathered two cells per lane; transparency goes through two warp ballots into the same
//     49-bit mask the visibility routine consumes; every lane then knows the visibility of its own cells.
//   * frame phase without a blit map: the frame is 28 pairs of pixel rows; a pair is 336 bytes = 21 16-byte chunks, and
//     chunk l of EVERY pair covers the same (cell column, tile part) twice over -- so lane l < 21 owns chunk l of all 28
//     pairs, its two (vi, part) are loop constants, the tile row (vj, py) is the loop counter, and after unrolling
//     every shared-memory and global address is `lane register + immediate`: 3 instructions per chunk (two 8-byte
//     atlas reads, one 16-byte store) plus 4 per tile row for the two kinds, ~115 per frame instead of ~480 with the
//     588-entry chunk map (ncu at 4096 envs, profiles/r02_warp_n4096_v2_ncu_details.csv: 1387 -> 791 warp instructions per env).
//     The kinds are kept premultiplied (kind * 192, the tile's byte offset in the atlas) as 16-bit words.
//   * atlas staging touches only the slots the pool can show: thread t tests tile t / 2 and copies half of it.
#ifndef MERLIN_WARP_THREADS
#define MERLIN_WARP_THREADS 256
#define MERLIN_WARP_CTAS 4
#endif
// RGB batches of at least this many envs launch the warp kernel with programmatic dependent launch (see the kernel)
#ifndef MERLIN_WARP_PDL_MIN_ENVS
#define MERLIN_WARP_PDL_MIN_ENVS 4096
#endif
constexpr int kWarpKernelThreads = MERLIN_WARP_THREADS;
constexpr int kWarpKindBytes = 128;          // 49 premultiplied kinds (u16) per warp, padded
constexpr int kPairChunks = 2 * kUnitsPerRow / 2;   // 21 16-byte chunks per pair of pixel rows

template <int STEP, bool PDL, bool LEAN = false>
__global__ void __launch_bounds__(kWarpKernelThreads, MERLIN_WARP_CTAS) env_kernel_warp(const EnvParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const Flags f(p);
  uint8_t* atlas_s = smem;
  uint16_t* kq = reinterpret_cast<uint16_t*>(smem + kAtlasBytes + warp * kWarpKindBytes);   // this warp's 49 kinds * 192

  int e = blockIdx.x * warps_per_cta + warp;
  int4 st_next = make_int4(0, 0, 0, 0);
  long long act_next = 0;
  if (PDL) {
    // Programmatic dependent launch (RGB batches of >= 4096 envs): the kernel is launched with
    // cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may be scheduled while the PREVIOUS kernel of the
    // stream is still draining its stores.  Everything up to griddepcontrol.wait touches only data no kernel writes (the
    // atlas and the tile mask are written by synchronous uploads): launch latency and the two-round-trip atlas staging
    // overlap the predecessor's tail.  After the wait the predecessor has completed and its writes (state, the policy's
    // logits, ...) are visible.  launch_dependents at once: a following env step may start its own prologue under THIS
    // kernel's store phase (all CTAs of a <= 24 576-env batch are resident together: early arrivals cannot starve it).
    // Measured (B200, CUDA-graph replay / eager back to back, us per step): 4096 envs 10.4 -> 9.7 / 12.3 -> 10.0,
    // 16 384 envs 28.1 -> 27.2 / 29.9 -> 27.5, 24 576 envs 40.9 -> 40.0; at 1024 envs graph replay gets SLOWER (4.9 ->
    // 8.0: the programmatic edge costs more than it hides), hence the size threshold.
    asm volatile("griddepcontrol.launch_dependents;");
  } else if (e < p.N) {
    // first env's state / action: in flight while the atlas is staged
    st_next = p.state[e];
    if (STEP == 1) act_next = p.actions[e];
  }
  if (f.want_rgb) {
    // thread t tests tile t / 2 and copies half of it (6 int4); a 128-thread CTA takes two rounds
    for (int t = threadIdx.x; t < 2 * kAtlasTiles; t += blockDim.x) {
      const int tile = t >> 1, half = t & 1;
      if (tile_bit(p.tile_present, tile)) {
        const int4* src = reinterpret_cast<const int4*>(p.atlas) + tile * (kTileBytes / 16) + half * 6;
        int4* dst = reinterpret_cast<int4*>(atlas_s) + tile * (kTileBytes / 16) + half * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) dst[i] = __ldg(src + i);
      }
    }
  }
  if (PDL) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (e < p.N) {
      st_next = p.state[e];
      if (STEP == 1) act_next = p.actions[e];
    }
  }
  __syncthreads();

  // this lane's two window cells, in mask order c = vj*7 + vi
  const int c0 = lane, c1 = lane + 32;
  const int vj0 = (c0 * 37) >> 8, vi0 = c0 - vj0 * kView;                       // c / 7 for c < 64
  const int vj1 = c1 < kCells ? (c1 * 37) >> 8 : 0, vi1 = c1 < kCells ? c1 - vj1 * kView : 0;
  const int a0 = (kView - 1) - vj0, b0 = vi0 - kView / 2, a1 = (kView - 1) - vj1, b1 = vi1 - kView / 2;
  // lean step: second cell of the 8-row window, ext index lane + 32 < 56 (row j = x / 7 counted from the far end)
  const int vj1x = (c1 * 37) >> 8, a1x = (kView - 1) - vj1x, b1x = c1 - vj1x * kView - kView / 2;
  // frame phase constants of this lane: chunk `lane` of every row pair = units 2*lane and 2*lane + 1 of 42
  const int u0 = 2 * lane, u1 = 2 * lane + 1;
  const int r0 = u0 >= kUnitsPerRow, r1 = u1 >= kUnitsPerRow;          // second row of the pair?
  const int w0 = u0 - r0 * kUnitsPerRow, w1 = u1 - r1 * kUnitsPerRow;
  const int cv0 = (w0 * 11) >> 5, cv1 = (w1 * 11) >> 5;                 // w / 3 for w < 21: the cell column vi
  const uint16_t* kq0 = kq + cv0 * kView;                               // + vj
  const uint16_t* kq1 = kq + cv1 * kView;
  const uint8_t* at0 = atlas_s + (w0 - 3 * cv0) * 8 + r0 * 24;          // + kind * 192 + (py & ~1) * 24
  const uint8_t* at1 = atlas_s + (w1 - 3 * cv1) * 8 + r1 * 24;

  for (; e < p.N; e += gridDim.x * warps_per_cta) {
    EnvState s{};
    const int4 st = st_next;
    const long long act_in = act_next;
    {
      const int en = e + gridDim.x * warps_per_cta;   // prefetch the next env of this warp
      if (en < p.N) {
        st_next = p.state[en];
        if (STEP == 1) act_next = p.actions[en];
      }
    }
    unpack_state(st.x, st.y, st.z, st.w, s);
    float ep_ret = p.ep_return[e];
    bool restart = false, render = true;
    uint32_t code0, code1 = CODE_WALL;   // this lane's two window cells (mask order c0 = lane, c1 = lane + 32)

    if (LEAN) {
      // Lean step (the reference's own configuration: ThreeActionWrapper, no reward-shaping wrapper => immutable grids):
      // TWO dependent round trips instead of three.  The action is known together with the state, hence the new
      // heading; the agent then stands on its old cell or one cell ahead of it.  Both windows lie inside 8 rows x 7
      // columns in front of the OLD cell (rows a = 7..0 ahead, ext index x = (7 - a) * 7 + vi): loaded at once, two
      // cells per lane (56 <= 64), BEFORE the forward cell has been looked at -- it is ext cell 45 itself.  The window
      // of an agent that moved is ext rows 0..6 (cell c = x: the lane's own loads); of one that did not, rows 1..7
      // (c = x - 7: one lane rotation by 7).
      ActionDraw act;
      if (STEP == 2) act = draw_action<true>(p, 3, e);
      else { act.action = act_in; act.logp = 0.f; act.draw = 0; }
      const bool bad = act.action < 0 || act.action >= 3;
      const int a = bad ? A_DONE : (int)act.action;
      s.dir = a == A_LEFT ? (s.dir + 3) & 3 : (a == A_RIGHT ? (s.dir + 1) & 3 : s.dir);
      const uint8_t* grid = p.pool_cells + (size_t)s.layout * p.cell_stride;
      const int fx = dir_dx(s.dir), fy = dir_dy(s.dir);
      uint32_t e0, e1 = CODE_WALL;
      {
        const int wx = s.x + (a0 + 1) * fx - b0 * fy, wy = s.y + (a0 + 1) * fy + b0 * fx;
        e0 = ((unsigned)wx < (unsigned)p.W && (unsigned)wy < (unsigned)p.H) ? ld_cell(grid + wy * p.W + wx, f.pol) : CODE_WALL;
      }
      if (lane + 32 < kCells + kView) {
        const int wx = s.x + (a1x + 1) * fx - b1x * fy, wy = s.y + (a1x + 1) * fy + b1x * fx;
        e1 = ((unsigned)wx < (unsigned)p.W && (unsigned)wy < (unsigned)p.H) ? ld_cell(grid + wy * p.W + wx, f.pol) : CODE_WALL;
      }
      const uint32_t ft = __shfl_sync(0xffffffffu, e1, 45 - 32) & 0xf;   // the cell ahead: ext row 6, column 3
      s.step_count += 1;
      const bool fwd_act = a == A_FORWARD;
      const bool moved = fwd_act && ((M_OVERLAP >> ft) & 1u);
      const bool goal = fwd_act && ft == T_GOAL;
      const bool terminated = goal || (fwd_act && ft == T_LAVA);
      const bool truncated = s.step_count >= p.max_steps;
      float rew = 0.f;
      if (goal) rew = (float)(1 - 0.9 * ((double)s.step_count / (double)p.max_steps));   // MiniGridEnv._reward in float64
      if (moved) { s.x += fx; s.y += fy; }
      ep_ret += rew;
      const bool done = terminated || truncated;
      if (lane == 0) {
        commit_action<STEP == 2>(p, e, act);
        record_first_episode<STEP == 2>(p, e, done, goal && rew > 0.f, ep_ret, s.step_count);
        if (bad) atomicAdd(p.bad_actions, 1ull);
        p.reward[e] = rew;
        p.terminated[e] = terminated ? 1 : 0;
        p.truncated[e] = truncated ? 1 : 0;
        if (p.out_ep_return) p.out_ep_return[e] = done ? ep_ret : 0.f;
        if (p.out_ep_length) p.out_ep_length[e] = done ? s.step_count : 0;
        if (p.out_stuck) p.out_stuck[e] = 0;
        if (p.out_done) p.out_done[e] = done ? 1.f : 0.f;
      }
      restart = done && f.auto_reset;
      const int rot = (lane + kView) & 31;
      const uint32_t r0 = __shfl_sync(0xffffffffu, e0, rot), r1 = __shfl_sync(0xffffffffu, e1, rot);
      code0 = moved ? e0 : (lane + kView < 32 ? r0 : r1);
      code1 = moved ? e1 : r1;
    } else if (STEP) {
      const uint8_t* grid = f.mutable_grid ? p.cells + (size_t)e * p.cell_stride
                                           : p.pool_cells + (size_t)s.layout * p.cell_stride;
      const int fx = s.x + dir_dx(s.dir), fy = s.y + dir_dy(s.dir);
      const bool inb = (unsigned)fx < (unsigned)p.W && (unsigned)fy < (unsigned)p.H;
      const int fidx = fy * p.W + fx;
      const uint32_t fwd = inb ? ld_cell(grid + fidx, f.pol) : CODE_WALL;
      ActionDraw act;
      if (STEP == 2) act = draw_action<true>(p, f.n_actions, e);
      else { act.action = act_in; act.logp = 0.f; act.draw = 0; }
      StepResult r = step_logic(s, act.action, f.n_actions, fwd, inb, fidx, p.max_steps);
      uint32_t vword = 0;
      const int cell = s.y * p.W + s.x;
      uint32_t* vptr = nullptr;
      if (f.explore_on) { vptr = p.visited + (size_t)e * p.vis_words + (cell >> 5); vword = *vptr; }
      bool stuck = false;
      const uint32_t vword_in = vword;
      const double rew_d = shape_reward(s, r.reward, f.stuck_on, p.stuck_max_stay, p.stuck_penalty, f.explore_on,
                                        p.explore_bonus, vword, cell & 31, stuck);
      const float rew = (float)rew_d;
      ep_ret += rew;
      const bool done = r.terminated || r.truncated;
      __syncwarp();  // every lane has read the old grid cell / visited word / draw counter before lane 0 overwrites them
      if (lane == 0) {
        commit_action<STEP == 2>(p, e, act);
        record_first_episode<STEP == 2>(p, e, done, r.terminated && r.reward > 0.0, ep_ret, s.step_count);
        if (r.bad_action) atomicAdd(p.bad_actions, 1ull);
        if (r.write_idx >= 0 && f.mutable_grid) p.cells[(size_t)e * p.cell_stride + r.write_idx] = (uint8_t)r.write_code;
        if (f.explore_on && vword != vword_in) *vptr = vword;
        p.reward[e] = rew;
        p.terminated[e] = r.terminated ? 1 : 0;
        p.truncated[e] = r.truncated ? 1 : 0;
        if (p.out_ep_return) p.out_ep_return[e] = done ? ep_ret : 0.f;
        if (p.out_ep_length) p.out_ep_length[e] = done ? s.step_count : 0;
        if (p.out_stuck) p.out_stuck[e] = stuck ? 1 : 0;
        if (p.out_done) p.out_done[e] = done ? 1.f : 0.f;
      }
      restart = done && f.auto_reset;
    } else {
      restart = p.reset_mask == nullptr || p.reset_mask[e] != 0;
      render = restart;
    }

    if (restart) {  // warp-uniform
      const int load_cur = s.layout < 0 ? ~s.layout
                                        : (f.advance ? (int)(((unsigned)s.layout + (unsigned)p.cursor_stride) % (unsigned)p.n_layouts) : s.layout);
      const uint32_t a = p.pool_agent[load_cur];
      s.x = a & 0xff; s.y = (a >> 8) & 0xff; s.dir = (a >> 16) & 3; s.carry = 0;
      s.step_count = 0; s.stay = 0; s.last_x = s.x; s.last_y = s.y;
      ep_ret = 0.f;
      s.layout = load_cur;
      if (!LEAN && f.mutable_grid) {
        __syncwarp();
        const int4* from = reinterpret_cast<const int4*>(p.pool_cells + (size_t)load_cur * p.cell_stride);
        int4* to = reinterpret_cast<int4*>(p.cells + (size_t)e * p.cell_stride);
        for (int i = lane; i < p.cell_stride / 16; i += 32) to[i] = from[i];
      }
      if (!LEAN && f.explore_on) {
        __syncwarp();
        const int cell = s.y * p.W + s.x;
        for (int i = lane; i < p.vis_words; i += 32)
          p.visited[(size_t)e * p.vis_words + i] = (i == (cell >> 5)) ? (1u << (cell & 31)) : 0u;
      }
    }
    if (lane == 0 && (STEP || restart)) {
      int4 o;
      pack_state(s, o.x, o.y, o.z, o.w);
      p.state[e] = o;
      p.ep_return[e] = ep_ret;
    }
    if (!render || !(f.want_rgb || f.want_sym)) continue;
    __syncwarp();  // grid writes of this step (pickup/drop/toggle, restart copy) are visible to the gather below

    // observation, part 1: two window cells per lane -> ballot -> visibility -> tile kinds (+ symbolic bytes)
    if (!LEAN || restart) {   // the lean step holds the window already unless the env restarted (warp-uniform)
      const uint8_t* grid = (!LEAN && f.mutable_grid) ? p.cells + (size_t)e * p.cell_stride
                                           : p.pool_cells + (size_t)s.layout * p.cell_stride;
      const int fx = dir_dx(s.dir), fy = dir_dy(s.dir);
      const int rx = -fy, ry = fx;
      code1 = CODE_WALL;
      {
        const int wx = s.x + a0 * fx + b0 * rx, wy = s.y + a0 * fy + b0 * ry;
        code0 = ((unsigned)wx < (unsigned)p.W && (unsigned)wy < (unsigned)p.H) ? ld_cell(grid + wy * p.W + wx, f.pol) : CODE_WALL;
      }
      if (c1 < kCells) {
        const int wx = s.x + a1 * fx + b1 * rx, wy = s.y + a1 * fy + b1 * ry;
        code1 = ((unsigned)wx < (unsigned)p.W && (unsigned)wy < (unsigned)p.H) ? ld_cell(grid + wy * p.W + wx, f.pol) : CODE_WALL;
      }
    }
    const unsigned t0 = __ballot_sync(0xffffffffu, !((M_OPAQUE >> (code0 & 0xf)) & 1u));
    const unsigned t1 = __ballot_sync(0xffffffffu, c1 < kCells && !((M_OPAQUE >> (code1 & 0xf)) & 1u));
    const uint64_t vis = visibility((uint64_t)t0 | ((uint64_t)t1 << 32));
    uint8_t* sym_out = f.want_sym ? p.obs_sym + (size_t)e * kSymBytes : nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = h ? c1 : c0;
      if (c >= kCells) break;
      const int vi = h ? vi1 : vi0, vj = h ? vj1 : vj0;
      uint32_t code = h ? code1 : code0;
      const bool seen = (vis >> c) & 1;
      const bool agent_cell = (vi == kView / 2 && vj == kView - 1);
      if (agent_cell) code = s.carry ? s.carry : CODE_EMPTY;
      const int k = vi * kView + vj;
      kq[k] = (uint16_t)((agent_cell ? agent_kind(s.carry) : (seen ? code : KIND_UNSEEN)) * kTileBytes);
      if (f.want_sym) {
        uint8_t t = 0, col = 0, stt = 0;
        if (seen) sym_of_code(code, t, col, stt);
        sym_out[k * 3 + 0] = t; sym_out[k * 3 + 1] = col; sym_out[k * 3 + 2] = stt;
      }
    }
    __syncwarp();
    if (f.want_rgb && lane < kPairChunks) {
      uint8_t* out = p.obs_rgb + (size_t)e * kImgBytes + lane * 16;
#pragma unroll
      for (int vj = 0; vj < kView; ++vj) {
        const uint8_t* t0p = at0 + kq0[vj];
        const uint8_t* t1p = at1 + kq1[vj];
#pragma unroll
        for (int q = 0; q < kTile / 2; ++q) {   // row pair vj*4 + q: pixel rows py = 2q, 2q + 1 of tile row vj
          const uint2 a = *reinterpret_cast<const uint2*>(t0p + q * 48);
          const uint2 b = *reinterpret_cast<const uint2*>(t1p + q * 48);
          st_stream_v4(out + (vj * (kTile / 2) + q) * (2 * kRowBytes), a.x, a.y, b.x, b.y);
        }
      }
    }
    __syncwarp();  // kq is reused by this warp's next env
  }
}

// ---------------------------------------------------------------------------------------------------
// env_kernel_quad<STEP> (kernel choice 7): one warp per FOUR environments, eight lanes per env -- a small-batch mapping
// for the reference's own configuration (ThreeActionWrapper => immutable grids, no reward-shaping wrapper; RGB frames).
// MEASURED AND NOT ADOPTED as a default (profiles/r02_quad_vs_warp.txt): it executes 328 instead of 773 instructions per
// env, but the small-batch step is bound by the latency of one env's dependent chain, not by issue slots (ncu at 4096
// envs: env_kernel_warp 33 % of the issue slots busy, 15.8 cycles per issued instruction per warp) -- fewer, longer-
// running warps hide less of it: 8.8 vs 8.3 us per step at 4096 envs, 22.4 vs 20.6 at 12 288, faster only around 8192
// (14.6 vs 15.7).  Kept selectable and held to the parity bar (tests/test_gpu_variants.py).
// In env_kernel_warp one instruction stream serves one env and every lane repeats the warp-uniform step logic.  Here one
// stream serves four envs -- everything below is uniform within a group of eight lanes and differs between the groups:
//   * lane r of a group loads ROW r of the 8-row x 7-column region in front of the agent's old cell (seven byte loads,
//     all in flight together with the other rows'): both candidate windows -- the agent stays, or moves one cell ahead --
//     lie inside it, so the step needs two dependent round trips (state + action, then cells), and the cell ahead is
//     byte 3 of row 6 (one shuffle);
//   * the lane's row arrives as 7 codes in a 64-bit register with its 7-bit transparency mask computed in place (no
//     ballots); the window row vj is ext row vj (agent moved) or vj + 1 (not moved): one lane-shifted shuffle;
//   * the seven row masks meet in shared memory (one byte each, one 8-byte read per lane); every lane runs
//     `visibility_rows` on them, takes its own row of the result and writes its seven tile kinds;
//   * frame phase as in env_kernel_warp (map-free row pairs), one env after the other.
// Quads are dealt to warps CTA-minor (quad q -> warp q / gridDim of CTA q % gridDim), so a batch that needs 1.3 rounds
// leaves every SM with the same share of second-round quads.
#ifndef MERLIN_QUAD_CTAS
#define MERLIN_QUAD_CTAS 4
#endif
constexpr int kQuadKindBytes = 4 * 128 + 32;   // per warp: 4 envs x 49 premultiplied kinds (u16, padded to 64) + 4 x 8 row masks

// Row `r` (0 = farthest, 7 = the agent's own row) of the 8 x 7 region in front of pose (x, y, heading f): codes of its
// seven cells, byte vi, and the transparency mask, bit vi.  Outside the grid = wall (Grid.slice).
__device__ __forceinline__ void load_ext_row(const EnvParams& p, const uint8_t* grid, uint64_t pol, int x, int y, int fx,
                                             int fy, int r, uint64_t& codes, uint32_t& transp) {
  const int ae = kView - r;
  int wx = x + ae * fx + (kView / 2) * fy, wy = y + ae * fy - (kView / 2) * fx;
  uint32_t c[kView];
#pragma unroll
  for (int vi = 0; vi < kView; ++vi) {
    c[vi] = ((unsigned)wx < (unsigned)p.W && (unsigned)wy < (unsigned)p.H) ? ld_cell(grid + wy * p.W + wx, pol) : CODE_WALL;
    wx -= fy; wy += fx;
  }
  codes = 0; transp = 0;
#pragma unroll
  for (int vi = 0; vi < kView; ++vi) {
    codes |= (uint64_t)c[vi] << (8 * vi);
    transp |= (((M_OPAQUE >> (c[vi] & 0xf)) & 1u) ^ 1u) << vi;
  }
}

template <int STEP, bool PDL>
__global__ void __launch_bounds__(kWarpKernelThreads, MERLIN_QUAD_CTAS) env_kernel_quad(const EnvParams p) {
  static_assert(STEP == 1 || STEP == 2, "the quad kernel steps; resets run the warp kernel");
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int g = lane >> 3, r = lane & 7;
  uint8_t* atlas_s = smem;
  uint16_t* kq = reinterpret_cast<uint16_t*>(smem + kAtlasBytes + warp * kQuadKindBytes);   // [4][64]
  uint8_t* tmask = smem + kAtlasBytes + warp * kQuadKindBytes + 4 * 128;                    // [4][8] row masks
  const uint64_t pol = grid_policy(true);
  const bool auto_reset = p.flags & MERLIN_F_AUTO_RESET, advance = !(p.flags & MERLIN_F_RESET_SAME);
  const int n_quads = (p.N + 3) >> 2;
  const int stride = gridDim.x * warps_per_cta;

  int q = warp * gridDim.x + blockIdx.x;
  int4 st_next = make_int4(0, 0, 0, 0);
  long long act_next = 0;
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;");   // see env_kernel_warp
  } else if (q < n_quads) {
    const int e = min(4 * q + g, p.N - 1);
    st_next = p.state[e];
    if (STEP == 1) act_next = p.actions[e];
  }
  for (int t = threadIdx.x; t < 2 * kAtlasTiles; t += blockDim.x) {
    const int tile = t >> 1, half = t & 1;
    if (tile_bit(p.tile_present, tile)) {
      const int4* src = reinterpret_cast<const int4*>(p.atlas) + tile * (kTileBytes / 16) + half * 6;
      int4* dst = reinterpret_cast<int4*>(atlas_s) + tile * (kTileBytes / 16) + half * 6;
#pragma unroll
      for (int i = 0; i < 6; ++i) dst[i] = __ldg(src + i);
    }
  }
  if (PDL) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (q < n_quads) {
      const int e = min(4 * q + g, p.N - 1);
      st_next = p.state[e];
      if (STEP == 1) act_next = p.actions[e];
    }
  }
  __syncthreads();

  // frame phase constants of this lane (see env_kernel_warp): chunk `lane` of every row pair
  const int u0 = 2 * lane, u1 = 2 * lane + 1;
  const int r0 = u0 >= kUnitsPerRow, r1 = u1 >= kUnitsPerRow;
  const int w0 = u0 - r0 * kUnitsPerRow, w1 = u1 - r1 * kUnitsPerRow;
  const int cv0 = (w0 * 11) >> 5, cv1 = (w1 * 11) >> 5;
  const int ko0 = cv0 * kView, ko1 = cv1 * kView;
  const uint8_t* at0 = atlas_s + (w0 - 3 * cv0) * 8 + r0 * 24;
  const uint8_t* at1 = atlas_s + (w1 - 3 * cv1) * 8 + r1 * 24;

  for (; q < n_quads; q += stride) {
    const bool valid = 4 * q + g < p.N;           // a ragged last quad: its surplus groups shadow env N - 1, store nothing
    const int e = valid ? 4 * q + g : p.N - 1;
    const int4 st = st_next;
    const long long act_in = act_next;
    if (q + stride < n_quads) {
      const int en = min(4 * (q + stride) + g, p.N - 1);
      st_next = p.state[en];
      if (STEP == 1) act_next = p.actions[en];
    }
    int x = st.x & 0xff, y = (st.x >> 8) & 0xff, dir = (st.x >> 16) & 3;
    uint32_t carry = ((uint32_t)st.x >> 24) & 0x7f;
    int step_count = st.y, layout = st.z, sw = st.w;
    float ep_ret = p.ep_return[e];

    ActionDraw act;
    if (STEP == 2) act = draw_action<true>(p, 3, e);
    else { act.action = act_in; act.logp = 0.f; act.draw = 0; }
    const bool bad = act.action < 0 || act.action >= 3;
    const int a = bad ? A_DONE : (int)act.action;
    dir = a == A_LEFT ? (dir + 3) & 3 : (a == A_RIGHT ? (dir + 1) & 3 : dir);
    int fx = dir_dx(dir), fy = dir_dy(dir);
    uint64_t codes;
    uint32_t transp_row;
    load_ext_row(p, p.pool_cells + (size_t)layout * p.cell_stride, pol, x, y, fx, fy, r, codes, transp_row);
    const uint32_t ft = __shfl_sync(0xffffffffu, (uint32_t)(codes >> 24), (lane & 24) | (kView - 1)) & 0xf;  // row 6, column 3

    step_count += 1;
    const bool fwd_act = a == A_FORWARD;
    bool moved = fwd_act && ((M_OVERLAP >> ft) & 1u);
    const bool goal = fwd_act && ft == T_GOAL;
    const bool terminated = goal || (fwd_act && ft == T_LAVA);
    const bool truncated = step_count >= p.max_steps;
    float rew = 0.f;
    if (goal) rew = (float)(1 - 0.9 * ((double)step_count / (double)p.max_steps));   // MiniGridEnv._reward, float64
    if (moved) { x += fx; y += fy; }
    ep_ret += rew;
    const bool done = terminated || truncated;
    if (r == 0 && valid) {
      commit_action<STEP == 2>(p, e, act);
      record_first_episode<STEP == 2>(p, e, done, goal && rew > 0.f, ep_ret, step_count);
      if (bad) atomicAdd(p.bad_actions, 1ull);
      p.reward[e] = rew;
      p.terminated[e] = terminated ? 1 : 0;
      p.truncated[e] = truncated ? 1 : 0;
      if (p.out_ep_return) p.out_ep_return[e] = done ? ep_ret : 0.f;
      if (p.out_ep_length) p.out_ep_length[e] = done ? step_count : 0;
      if (p.out_stuck) p.out_stuck[e] = 0;
      if (p.out_done) p.out_done[e] = done ? 1.f : 0.f;
    }
    if (done && auto_reset) {   // uniform within the group
      const int load_cur = layout < 0 ? ~layout
                                      : (advance ? (int)(((unsigned)layout + (unsigned)p.cursor_stride) % (unsigned)p.n_layouts) : layout);
      const uint32_t pa = p.pool_agent[load_cur];
      x = pa & 0xff; y = (pa >> 8) & 0xff; dir = (pa >> 16) & 3; carry = 0;
      step_count = 0; sw = (x << 16) | (int)((uint32_t)y << 24);
      ep_ret = 0.f;
      layout = load_cur;
      fx = dir_dx(dir); fy = dir_dy(dir);
      load_ext_row(p, p.pool_cells + (size_t)layout * p.cell_stride, pol, x, y, fx, fy, r, codes, transp_row);
      moved = false;
    }
    if (r == 0 && valid) {
      p.state[e] = make_int4(x | (y << 8) | (dir << 16) | (int)(carry << 24), step_count, layout, sw);
      p.ep_return[e] = ep_ret;
    }

    // window row vj = r (r < 7): ext row r when the agent moved, r + 1 when it did not
    const int src = (lane + (moved ? 0 : 1)) & 31;
    const uint32_t wlo = __shfl_sync(0xffffffffu, (uint32_t)codes, src);
    const uint32_t whi = __shfl_sync(0xffffffffu, (uint32_t)(codes >> 32), src);
    const uint32_t wt = __shfl_sync(0xffffffffu, transp_row, src);
    // the group's seven row masks, one byte each, through shared memory (a redux.sync on an 8-lane member mask is not
    // one instruction: REDUX reduces the whole warp, sub-warp masks take a loop)
    tmask[lane] = (uint8_t)wt;
    __syncwarp();
    const uint64_t vis = visibility_rows(*reinterpret_cast<const uint64_t*>(tmask + (lane & 24)));
    if (r < kView) {
      const uint32_t visrow = (uint32_t)(vis >> (8 * r)) & 0x7f;
      const uint64_t w = (uint64_t)wlo | ((uint64_t)whi << 32);
      uint16_t* kd = kq + g * 64 + r;
      uint8_t* so = (p.obs_sym != nullptr && valid) ? p.obs_sym + (size_t)e * kSymBytes + r * 3 : nullptr;
#pragma unroll
      for (int vi = 0; vi < kView; ++vi) {
        const bool seen = (visrow >> vi) & 1u;
        const bool agent_cell = vi == kView / 2 && r == kView - 1;
        uint32_t code = (uint32_t)(w >> (8 * vi)) & 0xff;
        if (agent_cell) code = carry ? carry : CODE_EMPTY;
        kd[vi * kView] = (uint16_t)((agent_cell ? agent_kind(carry) : (seen ? code : KIND_UNSEEN)) * kTileBytes);
        if (so) {   // Grid.encode of the cell, (0, 0, 0) when not visible: bytes (vi * 7 + vj) * 3 ..
          uint8_t t = 0, col = 0, stt = 0;
          if (seen) sym_of_code(code, t, col, stt);
          so[vi * kView * 3 + 0] = t; so[vi * kView * 3 + 1] = col; so[vi * kView * 3 + 2] = stt;
        }
      }
    }
    __syncwarp();
    if (lane < kPairChunks) {
      const int n_here = min(4, p.N - 4 * q);
      uint8_t* out = p.obs_rgb + (size_t)(4 * q) * kImgBytes + lane * 16;
      const uint16_t* kf = kq;
#pragma unroll 1
      for (int fr = 0; fr < n_here; ++fr, out += kImgBytes, kf += 64) {
#pragma unroll
        for (int vj = 0; vj < kView; ++vj) {
          const uint8_t* t0p = at0 + kf[ko0 + vj];
          const uint8_t* t1p = at1 + kf[ko1 + vj];
#pragma unroll
          for (int h = 0; h < kTile / 2; ++h) {
            const uint2 a2 = *reinterpret_cast<const uint2*>(t0p + h * 48);
            const uint2 b2 = *reinterpret_cast<const uint2*>(t1p + h * 48);
            st_stream_v4(out + (vj * (kTile / 2) + h) * (2 * kRowBytes), a2.x, a2.y, b2.x, b2.y);
          }
        }
      }
    }
    __syncwarp();  // kq is reused by this warp's next quad
  }
}

// ---------------------------------------------------------------------------------------------------
// render_kernel<BLOCKED>: frames from stored symbolic observations, one warp per frame, optional row gather.
// Rollouts can then keep 147 B per step instead of 9408 B and expand minibatches on read.
template <bool BLOCKED>
__global__ void __launch_bounds__(256, 2) render_kernel(const RenderParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  uint8_t* atlas_s = smem;
  uint8_t* kp = smem + kAtlasBytes + warp * kWarpKindStride;
  {
    const int4* src = reinterpret_cast<const int4*>(p.atlas);
    int4* dst = reinterpret_cast<int4*>(atlas_s);
    for (int i = threadIdx.x; i < kAtlasBytes / 16; i += blockDim.x) {
      const int tile = i / (kTileBytes / 16);
      if (tile_bit(p.tile_present, tile)) dst[i] = __ldg(src + i);
    }
  }
  uint32_t lut[kChunksPerLane];
#pragma unroll
  for (int k = 0; k < kChunksPerLane; ++k) lut[k] = __ldg(p.lut + k * 32 + lane);
  __syncthreads();

  // groups of consecutive frames are drawn in order from a ticket counter (see env_kernel_tile): 32 frames per ticket
  // for large batches (the write fronts of all CTAs stay in one narrow window), 8 -- one per warp -- for minibatch-sized
  // ones (16 384 frames are 512 groups of 32 on 296 CTAs: two rounds, the second 73 % full; 2048 groups of 8 are 6.9)
  const int kRenderGroup = p.group_frames;
  __shared__ int s_next;
  const int n_groups = (p.M + kRenderGroup - 1) / kRenderGroup;
  if (threadIdx.x == 0) s_next = (int)atomicAdd(&p.sched[0], 1u);
  __syncthreads();
  int group = s_next;
  while (group < n_groups) {
    __syncthreads();  // everyone has read the ticket
    if (threadIdx.x == 0) s_next = (int)atomicAdd(&p.sched[0], 1u);
    for (int m = group * kRenderGroup + warp; m < min(p.M, (group + 1) * kRenderGroup); m += warps_per_cta) {
      long long row = p.index ? p.index[m] : m;
      if (p.n_rows && (row < 0 || row >= p.n_rows)) row = 0;  // never read outside the buffer; callers validate indices
      const uint8_t* sym = p.sym + (size_t)row * kSymBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;  // cell index vi*7 + vj, the order Grid.encode stores them
        if (k < kCells) {
          const uint32_t t = sym[3 * k], c = sym[3 * k + 1], st = sym[3 * k + 2];
          kp[k] = (uint8_t)kind_of_sym(t, c, st, k == (kView / 2) * kView + (kView - 1));
        }
      }
      __syncwarp();
      uint8_t* frame = p.out + (size_t)m * kImgBytes;
      if (BLOCKED) {
        const uint4* atlas128 = reinterpret_cast<const uint4*>(atlas_s);
#pragma unroll
        for (int k = 0; k < kChunksPerLane; ++k) {
          const int c = lane + 32 * k;
          if (c < kChunks) {
            const uint32_t q = lut[k];
            const uint4 v = atlas128[kp[q & 0xff] * (kTileBytes / 16) + (q >> 8)];
            st_stream_v4(frame + c * 16, v.x, v.y, v.z, v.w);
          }
        }
      } else {
        blit_frame(atlas_s, kp, lut, frame, lane);
      }
      __syncwarp();  // kp is reused by this warp's next frame
    }
    __syncthreads();  // the next ticket is in shared memory
    group = s_next;
  }
  if (threadIdx.x == 0 && atomicAdd(&p.sched[1], 1u) == gridDim.x - 1) {
    p.sched[0] = 0;
    p.sched[1] = 0;
  }
}

#ifndef MERLIN_RENDER_GROUP32_MIN_FRAMES
#define MERLIN_RENDER_GROUP32_MIN_FRAMES 65536
#endif
cudaError_t launch_render(const RenderParams& p, bool blocked, int sm_count, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  constexpr int threads = 256, warps = threads / 32;
  const size_t smem = kAtlasBytes + warps * kWarpKindStride;
  RenderParams q = p;
  q.group_frames = p.M >= MERLIN_RENDER_GROUP32_MIN_FRAMES ? 32 : warps;
  const int grid = min(sm_count * 2, (p.M + q.group_frames - 1) / q.group_frames);
  if (blocked) render_kernel<true><<<grid, threads, smem, stream>>>(q);
  else render_kernel<false><<<grid, threads, smem, stream>>>(q);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// render_f32_kernel: the blocked frame as float32 -- f32[M][14][14][48], optionally pixel / 255.0f -- i.e. the very
// tensor the actor-critic's first layer reads (reference src/actor_critic.py:21 forms `x / 255.0` from a float32 copy
// of the frame on every evaluation).  Writing it here replaces three passes of the learner's minibatch path (u8 frame
// write, u8 read + f32 write of the cast, and the layout copy) by one 37 632-byte streaming write per frame.
//
// One warp per frame.  The atlas slots the layout pool can show are converted ONCE per CTA into a float atlas in
// shared memory (compacted: slot_of_kind[128]; 768 B per staged tile, `cap_tiles` of them), so a 16-byte output chunk
// is one u16 map read (shared by 4 lanes), one slot read and one 16-byte shared-memory read.  A kind that is not
// staged (a CUDA graph replayed after a re-upload brought new tile kinds) is converted on the fly from the u8 atlas.
// HBM-bound: 147 (+8) B read, 37 632 B written per frame.
constexpr int kF32Chunks = kImgBytes / 4;                  // 2352 float4 chunks per frame
constexpr int kF32Iters = (kF32Chunks + 31) / 32;          // 74
constexpr int kRenderF32Group = 8;                          // frames per ticket: 301 KB, like the u8 kernels' groups
constexpr int kRenderF32Threads = 256;
#ifndef MERLIN_RENDER_F32_CTA_FRAMES_MAX
#define MERLIN_RENDER_F32_CTA_FRAMES_MAX 8192
#endif

__host__ __device__ constexpr size_t render_f32_smem(int cap_tiles) {
  return (size_t)cap_tiles * kTileBytes * 4 + 128 + 592 * 2 + (kRenderF32Threads / 32) * 128;
}

// normalise: 0 = the pixel value, 1 = pixel / 255.0f (IEEE division: what torch's CPU kernels compute for `x / 255.0`),
// 2 = pixel * (1.0f / 255.0f) (what torch's CUDA kernel computes for a tensor divided by a Python scalar)
__device__ __forceinline__ float pixel_f32(uint32_t b, int normalise) {
  const float v = (float)b;
  if (normalise == 1) return __fdiv_rn(v, 255.0f);
  if (normalise == 2) return __fmul_rn(v, __fdiv_rn(1.0f, 255.0f));
  return v;
}

__global__ void __launch_bounds__(kRenderF32Threads, 3) render_f32_kernel(const RenderParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int warps_per_cta = kRenderF32Threads / 32;
  float* atlas_f = reinterpret_cast<float*>(smem);
  uint8_t* slot_s = smem + (size_t)p.cap_tiles * kTileBytes * 4;
  uint16_t* map_s = reinterpret_cast<uint16_t*>(slot_s + 128);
  uint8_t* kp = reinterpret_cast<uint8_t*>(map_s + 592) + warp * 128;   // [0..48] slots, [64..112] kinds
  const int normalise = p.normalise;

  if (threadIdx.x < kAtlasTiles) {  // compact the present tiles: slot = number of present tiles below this one
    const int t = threadIdx.x;
    int below = 0;
    for (int w = 0; w < (t >> 5); ++w) below += __popc(__ldg(p.tile_present + w));
    below += __popc(__ldg(p.tile_present + (t >> 5)) & ((1u << (t & 31)) - 1u));
    slot_s[t] = (tile_bit(p.tile_present, t) && below < p.cap_tiles) ? (uint8_t)below : (uint8_t)255;
  }
  for (int c = threadIdx.x; c < kChunks; c += blockDim.x) map_s[c] = (uint16_t)chunk_lut_blocked(c);
  __syncthreads();
  for (int t = warp; t < kAtlasTiles; t += warps_per_cta) {
    const uint32_t slot = slot_s[t];
    if (slot == 255) continue;
    for (int i = lane; i < kTileBytes; i += 32)
      atlas_f[slot * kTileBytes + i] = pixel_f32(__ldg(p.atlas + t * kTileBytes + i), normalise);
  }
  __syncthreads();

  // one frame's float4 chunk f: one u16 map read (shared by 4 lanes), one slot read, one 16-byte atlas read, one store
  auto emit_chunk = [&](const uint8_t* kpw, float* frame, int f) {
    const uint32_t q = map_s[f >> 2];
    const uint32_t cell = q & 0xff, off = (q >> 8) * 16 + (f & 3) * 4;   // element offset inside the tile
    const uint32_t slot = kpw[cell];
    float4 v;
    if (slot != 255) {
      v = *reinterpret_cast<const float4*>(atlas_f + slot * kTileBytes + off);
    } else {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p.atlas + (uint32_t)kpw[64 + cell] * kTileBytes + off));
      v = make_float4(pixel_f32(w & 0xff, normalise), pixel_f32((w >> 8) & 0xff, normalise),
                      pixel_f32((w >> 16) & 0xff, normalise), pixel_f32(w >> 24, normalise));
    }
    st_stream_v4(frame + f * 4, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
  };
  auto load_kinds = [&](uint8_t* kpw, int m, int t) {   // thread t < 64 of the caller's group handles cells t (and < 49)
    long long row = p.index ? p.index[m] : m;
    if (p.n_rows && (row < 0 || row >= p.n_rows)) row = 0;
    const uint8_t* sym = p.sym + (size_t)row * kSymBytes;
    if (t < kCells) {
      const uint32_t kind = kind_of_sym(sym[3 * t], sym[3 * t + 1], sym[3 * t + 2], t == (kView / 2) * kView + (kView - 1));
      kpw[t] = slot_s[kind];
      kpw[64 + t] = (uint8_t)kind;
    }
  };

  if (p.frame_per_cta) {
    // few frames (a policy-input render for a small batch: launch latency is what counts): one CTA per frame, all
    // eight warps share its 2352 chunks -- 10 chunk rounds per thread instead of 74 per lane
    for (int m = blockIdx.x; m < p.M; m += gridDim.x) {
      __syncthreads();  // kp of the previous frame has been consumed
      load_kinds(kp - warp * 128, m, threadIdx.x);
      __syncthreads();
      float* frame = p.out_f32 + (size_t)m * kImgBytes;
      for (int f = threadIdx.x; f < kF32Chunks; f += kRenderF32Threads) emit_chunk(kp - warp * 128, frame, f);
    }
    return;
  }

  __shared__ int s_next;
  const int n_groups = (p.M + kRenderF32Group - 1) / kRenderF32Group;
  if (threadIdx.x == 0) s_next = (int)atomicAdd(&p.sched[0], 1u);
  __syncthreads();
  int group = s_next;
  while (group < n_groups) {
    __syncthreads();  // everyone has read the ticket
    if (threadIdx.x == 0) s_next = (int)atomicAdd(&p.sched[0], 1u);
    for (int m = group * kRenderF32Group + warp; m < min(p.M, (group + 1) * kRenderF32Group); m += warps_per_cta) {
      load_kinds(kp, m, lane);
      load_kinds(kp, m, lane + 32);
      __syncwarp();
      float* frame = p.out_f32 + (size_t)m * kImgBytes;
#pragma unroll 4
      for (int k = 0; k < kF32Iters; ++k) {
        const int f = lane + 32 * k;
        if (f < kF32Chunks) emit_chunk(kp, frame, f);
      }
      __syncwarp();  // kp is reused by this warp's next frame
    }
    __syncthreads();  // the next ticket is in shared memory
    group = s_next;
  }
  if (threadIdx.x == 0 && atomicAdd(&p.sched[1], 1u) == gridDim.x - 1) {
    p.sched[0] = 0;
    p.sched[1] = 0;
  }
}

cudaError_t launch_render_f32(const RenderParams& p, int sm_count, cudaStream_t stream) {
  if (p.M <= 0) return cudaSuccess;
  const size_t smem = render_f32_smem(p.cap_tiles);
  if (smem > 48 * 1024) {  // every atlas slot staged (7-action handles): opt in, per device (the call is cheap)
    cudaError_t err = cudaFuncSetAttribute(render_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)render_f32_smem(kAtlasTiles));
    if (err != cudaSuccess) return err;
  }
  const int per_sm = smem > 72 * 1024 ? 2 : 3;
  RenderParams q = p;
  // Few frames -- up to a few rounds of (resident CTAs x 8 warps): a CTA renders a frame with all its warps and strides
  // over the frames (no ticket counter involved).  With a warp per frame, 4096 frames are 1.15 rounds of the 3552 resident
  // warps, i.e. two rounds of which the second is 15 % full; with a CTA per frame they are 9.2 rounds of 444 CTAs, i.e. ten.
  // Above: groups of 8 frames per ticket, one warp per frame.
  q.frame_per_cta = p.M <= MERLIN_RENDER_F32_CTA_FRAMES_MAX ? 1 : 0;
  const int grid = q.frame_per_cta ? min(p.M, sm_count * per_sm)
                                   : min(sm_count * per_sm, (p.M + kRenderF32Group - 1) / kRenderF32Group);
  render_f32_kernel<<<grid, kRenderF32Threads, smem, stream>>>(q);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// full_obs_kernel: the fully observable symbolic observation (minigrid FullyObsWrapper; selected by
// `observation.fully_observable: true` in the reference's scenario.yaml, src/scenario_creator/scenario_creator.py:45-46).
// One thread per output cell, output-order indexing (coalesced 3-byte cells; the 256-byte grids are read through L1).
__global__ void __launch_bounds__(256) full_obs_kernel(const EnvParams p, uint8_t* __restrict__ out) {
  const long long cells_per_env = (long long)p.W * p.H;
  const long long total = cells_per_env * p.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / cells_per_env);
    const int c = (int)(i - (long long)e * cells_per_env);
    const int x = c / p.H, y = c - x * p.H;  // output index [x][y]
    const int4 st = p.state[e];
    EnvState s{};
    unpack_state(st.x, st.y, st.z, st.w, s);
    const uint8_t* grid = p.cells ? p.cells + (size_t)e * p.cell_stride
                                  : p.pool_cells + (size_t)(s.layout < 0 ? ~s.layout : s.layout) * p.cell_stride;
    uint8_t t, col, stt;
    if (x == s.x && y == s.y) { t = (uint8_t)T_AGENT; col = 0; stt = (uint8_t)s.dir; }
    else sym_of_code(grid[y * p.W + x], t, col, stt);
    out[i * 3 + 0] = t; out[i * 3 + 1] = col; out[i * 3 + 2] = stt;
  }
}

cudaError_t launch_full_obs(const EnvParams& p, uint8_t* out, int sm_count, cudaStream_t stream) {
  const long long total = (long long)p.W * p.H * p.N;
  const int grid = (int)min((long long)sm_count * 8, (total + 255) / 256);
  full_obs_kernel<<<grid, 256, 0, stream>>>(p, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// launch helpers: persistent grids of (SMs x resident CTAs), capped by the work available.  The resident-CTA count of
// every kernel instance is looked up once per HANDLE (LaunchCtx::occ, one slot per instance): a process may drive
// several handles on several devices from several threads, so nothing here is process-wide.
template <typename Kernel>
static cudaError_t resident_ctas(Kernel kernel, int threads, size_t smem, int& blocks_per_sm) {
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, threads, smem);
  if (err != cudaSuccess) return err;
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  return cudaSuccess;
}

// occupancy-cache slots (LaunchCtx::occ): one per kernel instance
enum : int { kSlotGroup = 0 /* + 3*log2(32/G) + STEP: 12 */, kSlotTile = 12 /* + (T==8)*6 + SWAR*3 + STEP: 12 */,
             kSlotOrdered = 24 /* + SWAR*3 + STEP: 6 */, kSlotTma = 30 /* + STEP: 3 */, kSlotWarp = 33 /* + LEAN*3 + STEP: 6 */, kSlotQuad = 39 /* + STEP - 1: 2 */ };
static_assert(kSlotQuad + 2 <= kOccSlots, "occupancy cache too small");

template <int G, int STEP>
static cudaError_t launch_group_kernel(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  const int n_groups = (p.N + G - 1) / G;
  const size_t smem = cta_smem_bytes(G);
  int& blocks_per_sm = ctx.occ[kSlotGroup + 3 * (G == 32 ? 0 : G == 16 ? 1 : G == 8 ? 2 : 3) + STEP];
  if (!blocks_per_sm) {
    cudaError_t err = resident_ctas(env_kernel<G, STEP>, kThreads, smem, blocks_per_sm);
    if (err != cudaSuccess) return err;
  }
  const int grid = min(ctx.sm_count * blocks_per_sm, (n_groups + kWarps - 1) / kWarps);
  env_kernel<G, STEP><<<grid, kThreads, smem, stream>>>(p, n_groups);
  return cudaGetLastError();
}

template <int T, int STEP, bool SWAR, int THREADS = kTileThreads, int MINB = kTileCtasPerSm>
static cudaError_t launch_tile_kernel_impl(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  const int n_tiles = (p.N + T - 1) / T;
  const size_t smem = tile_smem_bytes(T);
  int& blocks_per_sm = ctx.occ[kSlotTile + (T == 8 ? 6 : 0) + (SWAR ? 3 : 0) + STEP];
  if (!blocks_per_sm) {
    cudaError_t err = resident_ctas(env_kernel_tile<T, STEP, THREADS, MINB, SWAR>, THREADS, smem, blocks_per_sm);
    if (err != cudaSuccess) return err;
    if (MINB < blocks_per_sm) blocks_per_sm = MINB;
  }
  const int grid = min(ctx.sm_count * blocks_per_sm, n_tiles);
  env_kernel_tile<T, STEP, THREADS, MINB, SWAR><<<grid, THREADS, smem, stream>>>(p, n_tiles);
  return cudaGetLastError();
}
template <int T, int STEP>
static cudaError_t launch_tile_kernel(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  return use_swar(p, ctx, true) ? launch_tile_kernel_impl<T, STEP, true>(p, ctx, stream)
                                : launch_tile_kernel_impl<T, STEP, false>(p, ctx, stream);
}

#ifndef MERLIN_ORD_G
#define MERLIN_ORD_G 8
#define MERLIN_ORD_THREADS 128
#define MERLIN_ORD_CTAS 3
#endif
template <int G, int STEP, bool SWAR, int THREADS = MERLIN_ORD_THREADS, int MINB = MERLIN_ORD_CTAS>
static cudaError_t launch_ordered_kernel_impl(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  constexpr int warps = THREADS / 32;
  const int n_groups = (p.N + G - 1) / G;
  const size_t smem = kAtlasBytes + warps * warp_smem_bytes(G);
  int& blocks_per_sm = ctx.occ[kSlotOrdered + (SWAR ? 3 : 0) + STEP];
  if (!blocks_per_sm) {
    cudaError_t err = resident_ctas(env_kernel_ordered<G, STEP, THREADS, MINB, SWAR>, THREADS, smem, blocks_per_sm);
    if (err != cudaSuccess) return err;
    if (MINB < blocks_per_sm) blocks_per_sm = MINB;
  }
  const int grid = min(ctx.sm_count * blocks_per_sm, (n_groups + warps - 1) / warps);
  env_kernel_ordered<G, STEP, THREADS, MINB, SWAR><<<grid, THREADS, smem, stream>>>(p, n_groups);
  return cudaGetLastError();
}
template <int G, int STEP>
static cudaError_t launch_ordered_kernel(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  return use_swar(p, ctx, true) ? launch_ordered_kernel_impl<G, STEP, true>(p, ctx, stream)
                                : launch_ordered_kernel_impl<G, STEP, false>(p, ctx, stream);
}

#ifndef MERLIN_TMA_T
#define MERLIN_TMA_T 32
#define MERLIN_TMA_THREADS 128
#define MERLIN_TMA_CTAS 3
#define MERLIN_TMA_NBUF 1
#endif
template <int T, int STEP, int THREADS = MERLIN_TMA_THREADS, int MINB = MERLIN_TMA_CTAS, int NBUF = MERLIN_TMA_NBUF>
static cudaError_t launch_tile_tma_kernel(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  const int n_tiles = (p.N + T - 1) / T;
  const size_t smem = tile_tma_smem_bytes(T, THREADS, NBUF);
  int& blocks_per_sm = ctx.occ[kSlotTma + STEP];
  if (!blocks_per_sm) {
    cudaError_t err = resident_ctas(env_kernel_tile_tma<T, STEP, THREADS, MINB, NBUF>, THREADS, smem, blocks_per_sm);
    if (err != cudaSuccess) return err;
    if (MINB < blocks_per_sm) blocks_per_sm = MINB;
  }
  const int grid = min(ctx.sm_count * blocks_per_sm, n_tiles);
  env_kernel_tile_tma<T, STEP, THREADS, MINB, NBUF><<<grid, THREADS, smem, stream>>>(p, n_tiles);
  return cudaGetLastError();
}

// Shape of a warp-kernel launch.  Batches that fit the machine in one round (N <= SMs x resident CTAs x 8 warps) are
// BALANCED: with 8-warp CTAs, 4096 envs are 512 CTAs on 148 SMs -- 68 SMs get four CTAs (32 envs), 80 get three (24),
// and the launch lasts as long as the SMs with 32.  Instead the CTA shape follows the batch: the fewest envs per SM
// that cover it, ceil(N / SMs), split into c <= resident CTAs of w <= 8 warps (4096 envs: 4 x 7 warps, 28 envs on
// (almost) every SM).  Larger batches loop over rounds of full CTAs, where the imbalance is a few percent at most.
#ifndef MERLIN_WARP_BALANCE
#define MERLIN_WARP_BALANCE 1
#endif
static void warp_kernel_shape(int N, int sm_count, int blocks_per_sm, int& grid, int& threads, bool few_ctas = false) {
  constexpr int max_warps = kWarpKernelThreads / 32;
  threads = kWarpKernelThreads;
  grid = min(sm_count * blocks_per_sm, (N + max_warps - 1) / max_warps);
  if (!MERLIN_WARP_BALANCE || N > sm_count * blocks_per_sm * max_warps) return;
  const int per_sm = (N + sm_count - 1) / sm_count;
  int best_w = max_warps, best_cost = 1 << 30;
  for (int c = blocks_per_sm; c >= 1; --c) {   // ties: more, smaller CTAs (each clears its staging barrier sooner)
    const int w = (per_sm + c - 1) / c;
    if (w > max_warps) continue;
    if (c * w < best_cost || (few_ctas && c * w == best_cost)) { best_cost = c * w; best_w = w; }
  }
  threads = best_w * 32;
  grid = (N + best_w - 1) / best_w;
}

template <int STEP, bool LEAN>
static cudaError_t launch_warp_kernel_impl(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  constexpr int warps = kWarpKernelThreads / 32;
  const size_t smem = kAtlasBytes + warps * kWarpKindBytes;
  int& blocks_per_sm = ctx.occ[kSlotWarp + (LEAN ? 3 : 0) + STEP];
  if (!blocks_per_sm) {
    cudaError_t err = resident_ctas(env_kernel_warp<STEP, false, LEAN>, kWarpKernelThreads, smem, blocks_per_sm);
    if (err == cudaSuccess) {
      int same = 0;
      err = resident_ctas(env_kernel_warp<STEP, true, LEAN>, kWarpKernelThreads, smem, same);
    }
    if (err != cudaSuccess) { blocks_per_sm = 0; return err; }
  }
  int grid, threads;
  warp_kernel_shape(p.N, ctx.sm_count, blocks_per_sm, grid, threads);
  if (p.obs_rgb != nullptr && p.N >= MERLIN_WARP_PDL_MIN_ENVS) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, env_kernel_warp<STEP, true, LEAN>, p);
  }
  env_kernel_warp<STEP, false, LEAN><<<grid, threads, smem, stream>>>(p);
  return cudaGetLastError();
}
// The lean step (two dependent round trips, see the kernel) serves the reference's own configuration: three actions
// (=> immutable grids) and no reward-shaping wrapper.
#ifndef MERLIN_WARP_LEAN
#define MERLIN_WARP_LEAN 1
#endif
// ... while every env of the batch is resident at once or nearly so: from 16 384 envs up (3.5 rounds of resident warps,
// chains hidden behind other warps' stores) the lean form's 56-cell region and shuffles cost 1.5 % (27.5 vs 27.0 us)
#ifndef MERLIN_WARP_LEAN_MAX_ENVS
#define MERLIN_WARP_LEAN_MAX_ENVS 12288
#endif
template <int STEP>
static cudaError_t launch_warp_kernel(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  constexpr uint32_t kNotLean = MERLIN_F_SEVEN_ACTIONS | MERLIN_F_STUCK_PENALTY | MERLIN_F_EXPLORE_BONUS;
  if (MERLIN_WARP_LEAN && p.N <= MERLIN_WARP_LEAN_MAX_ENVS && STEP != 0 && (p.flags & kNotLean) == 0 && p.cells == nullptr)
    return launch_warp_kernel_impl<STEP, STEP != 0>(p, ctx, stream);
  return launch_warp_kernel_impl<STEP, false>(p, ctx, stream);
}

// env_kernel_quad serves steps of the lean configuration with RGB frames; anything else asked of choice 7 runs the warp kernel.
static bool quad_eligible(const EnvParams& p) {
  constexpr uint32_t kNotLean = MERLIN_F_SEVEN_ACTIONS | MERLIN_F_STUCK_PENALTY | MERLIN_F_EXPLORE_BONUS;
  return (p.flags & kNotLean) == 0 && p.cells == nullptr && p.obs_rgb != nullptr;
}
// The quad kernel is launched with programmatic dependent launch only when its CTAs fill every resident slot of the
// machine: a grid that leaves room lets its successors (which signal launch_dependents on entry themselves) pile up
// resident behind it, and a step then takes twice as long (4096 envs = 147 CTAs: 16 us instead of 9).
#ifndef MERLIN_QUAD_PDL_MIN_ENVS
#define MERLIN_QUAD_PDL_MIN_ENVS 16384
#endif
template <typename Kernel>
static cudaError_t launch_maybe_pdl(Kernel kernel, bool pdl, int grid, int threads, size_t smem, cudaStream_t stream,
                                    const EnvParams& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}
template <int STEP>
static cudaError_t launch_quad_kernel(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  constexpr int warps = kWarpKernelThreads / 32;
  const size_t smem = kAtlasBytes + warps * kQuadKindBytes;
  int& blocks_per_sm = ctx.occ[kSlotQuad + STEP - 1];
  if (!blocks_per_sm) {
    cudaError_t err = resident_ctas(env_kernel_quad<STEP, false>, kWarpKernelThreads, smem, blocks_per_sm);
    if (err == cudaSuccess) {
      int same = 0;
      err = resident_ctas(env_kernel_quad<STEP, true>, kWarpKernelThreads, smem, same);
    }
    if (err != cudaSuccess) { blocks_per_sm = 0; return err; }
    if (MERLIN_QUAD_CTAS < blocks_per_sm) blocks_per_sm = MERLIN_QUAD_CTAS;
  }
  int grid, threads;
  warp_kernel_shape((p.N + 3) >> 2, ctx.sm_count, blocks_per_sm, grid, threads, /*few_ctas=*/true);   // one warp per quad
  if (p.N >= MERLIN_QUAD_PDL_MIN_ENVS)
    return launch_maybe_pdl(env_kernel_quad<STEP, true>, true, grid, threads, smem, stream, p);
  return launch_maybe_pdl(env_kernel_quad<STEP, false>, false, grid, threads, smem, stream, p);
}

// tiles of 16 envs from this batch size up, tiles of 8 below
#ifndef MERLIN_TILE16_MIN_ENVS
#define MERLIN_TILE16_MIN_ENVS(SMS) ((SMS) * kTileCtasPerSm * 16)
#endif
constexpr int kSymWarpMaxEnvs = 2048;   // symbolic-only batches up to this size run the warp-per-env kernel

template <int STEP>
static cudaError_t launch_sized(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  const int sm_count = ctx.sm_count;
  int choice = ctx.kernel_choice;
  if (choice == 0) {
    // symbolic-only observations are instruction-bound: the state phase alone, at high occupancy -- except for batches so
    // small that one env's chain of ~1800 dependent instructions IS the run time: there a warp per env (the same chain
    // spread over 32 lanes) is 1.3 us faster (2.9 vs 4.2 us at 32 envs, 3.4 vs 4.7 us at 1024; equal at ~3000)
    if (p.obs_rgb == nullptr) choice = p.N <= kSymWarpMaxEnvs ? 2 : 5;
    else choice = MERLIN_AUTO_RGB_CHOICE(p.N, sm_count);
  }
  if (choice == 2 && ctx.kernel_choice == 0 && MERLIN_AUTO_QUAD(p.N)) choice = 7;   // quads where they apply and win
  if (choice == 7) {
    if (STEP != 0 && quad_eligible(p)) return launch_quad_kernel<STEP == 0 ? 1 : STEP>(p, ctx, stream);
    choice = 2;
  }
  if (choice == 2) return launch_warp_kernel<STEP>(p, ctx, stream);
  if (choice == 5) return launch_sym_kernel<STEP>(p, ctx, stream);
  if (choice == 6) return launch_ordered_kernel<MERLIN_ORD_G, STEP>(p, ctx, stream);
  if (choice == 4) {
    if ((reinterpret_cast<uintptr_t>(p.obs_rgb) & 15) == 0) return launch_tile_tma_kernel<MERLIN_TMA_T, STEP>(p, ctx, stream);
    choice = 3;  // bulk copies need a 16-byte aligned destination
  }
  if (choice == 3) {
    // tiles of 16 envs once every resident CTA gets one; smaller tiles spread a small batch over more CTAs
    if (p.N >= MERLIN_TILE16_MIN_ENVS(sm_count)) return launch_tile_kernel<16, STEP>(p, ctx, stream);
    return launch_tile_kernel<8, STEP>(p, ctx, stream);
  }
  // pick the largest group size that still yields >= ~8 warps per SM; tiny batches use smaller groups
  const long long want_warps = (long long)sm_count * 8;
  if (p.N / 32 >= want_warps) return launch_group_kernel<32, STEP>(p, ctx, stream);
  if (p.N / 16 >= want_warps) return launch_group_kernel<16, STEP>(p, ctx, stream);
  if (p.N / 8 >= want_warps) return launch_group_kernel<8, STEP>(p, ctx, stream);
  return launch_group_kernel<4, STEP>(p, ctx, stream);
}

const char* step_kernel_name(int n_envs, bool rgb, bool quad_ok, const LaunchCtx& ctx) {
  const int sm_count = ctx.sm_count;
  int choice = ctx.kernel_choice;
  if (choice == 0) {
    choice = rgb ? MERLIN_AUTO_RGB_CHOICE(n_envs, sm_count) : (n_envs <= kSymWarpMaxEnvs ? 2 : 5);
    if (choice == 2 && rgb && MERLIN_AUTO_QUAD(n_envs)) choice = 7;
  }
  if (choice == 7) {
    if (quad_ok && rgb) return "merlin::env_kernel_quad<true>";
    choice = 2;
  }
  if (choice == 5) return "merlin::env_kernel_sym<true>";
  if (choice == 2) return "merlin::env_kernel_warp<true>";
  if (choice == 4) return "merlin::env_kernel_tile_tma<true>";
  if (choice == 6) return "merlin::env_kernel_ordered<true>";
  if (choice == 3)
    return n_envs >= MERLIN_TILE16_MIN_ENVS(sm_count) ? "merlin::env_kernel_tile<16,true>" : "merlin::env_kernel_tile<8,true>";
  const long long want_warps = (long long)sm_count * 8;
  if (n_envs / 32 >= want_warps) return "merlin::env_kernel<32,true>";
  if (n_envs / 16 >= want_warps) return "merlin::env_kernel<16,true>";
  if (n_envs / 8 >= want_warps) return "merlin::env_kernel<8,true>";
  return "merlin::env_kernel<4,true>";
}

cudaError_t launch_env_step(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  return p.logits != nullptr ? launch_sized<2>(p, ctx, stream) : launch_sized<1>(p, ctx, stream);
}
cudaError_t launch_env_reset(const EnvParams& p, const LaunchCtx& ctx, cudaStream_t stream) {
  return launch_sized<0>(p, ctx, stream);
}

}  // namespace merlin
