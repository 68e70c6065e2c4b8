// env_kernels.cu -- fused MERLIN env kernels for sm_100a.
//
//   env_kernel<G, STEP=true >  : step + wrappers + auto-reset + gen_obs(+process_vis) + symbolic encode + RGB blit
//   env_kernel<G, STEP=false>  : (masked) reset + the same observation path
//
// Mapping (HBM-bound streaming writer; tensor cores are not involved):
//   * one WARP owns G consecutive environments (G = 32 at scale; smaller G only to spread tiny batches
//     over more SMs).  Phase A runs one env per lane: 128-bit coalesced load of the packed state,
//     coalesced action read, step logic, reward shaping, 49-cell window gather, bitmask visibility.
//     The 49 tile kinds (and the 147 symbolic bytes) of each env go to shared memory.
//   * Phase B: the warp walks its G envs; per env the 32 lanes emit the 9408-byte frame as 588
//     coalesced 16-byte streaming stores, each assembled from two 8-byte reads of the tile atlas
//     held in shared memory.  Per-lane chunk->(cell, tile offset) maps are computed once per kernel
//     and live in registers.
//   * All per-env scalars (reward, flags, episode stats, state) are written by lane=env, i.e. coalesced.
//   * Finished envs are restarted by the whole warp (coalesced 16-byte copies of the pool layout) when the
//     grid is mutable; with the 3-action set grids are immutable and envs read the pool in place.
//
// Algorithmic HBM bytes per env-step (16x16, RGB): 9408 obs + 256 grid + 32 state + 8 action + 6 = 9710.
#include <cuda_runtime.h>
#include <stdint.h>

#include "env_kernels.cuh"
#include "env_logic.cuh"

namespace merlin {

__device__ __forceinline__ void st_stream_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int G, bool STEP>
__global__ void __launch_bounds__(kThreads, MERLIN_MIN_BLOCKS) env_kernel(const EnvParams p, const int n_groups) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool want_rgb = p.obs_rgb != nullptr;
  const bool want_sym = p.obs_sym != nullptr;

  uint8_t* atlas_s = smem;
  uint8_t* warp_s = smem + kAtlasBytes + kLutBytes + warp * warp_smem_bytes(G);
  uint8_t* kinds_s = warp_s;                       // [G][kKindStride]
  uint8_t* sym_s = warp_s + G * kKindStride;       // [G][147] contiguous, same layout as the output rows

  if (want_rgb) {  // stage the 24 KB tile atlas once per CTA
    const int4* src = reinterpret_cast<const int4*>(p.atlas);
    int4* dst = reinterpret_cast<int4*>(atlas_s);
    for (int i = threadIdx.x; i < kAtlasBytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  // blit map: chunk c = lane + 32*k  ->  (cell0, off0, cell1, off1); per lane in registers, or one copy in smem
#if MERLIN_LUT_SMEM
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem + kAtlasBytes);   // [k][lane]: conflict-free
  for (int c = threadIdx.x; c < kChunksPerLane * 32; c += blockDim.x) lut[c] = c < kChunks ? chunk_lut(c) : 0u;
  __syncthreads();
#define MERLIN_LUT(k) lut[(k) * 32 + lane]
#else
  __syncthreads();
  uint32_t lut[kChunksPerLane];
#pragma unroll
  for (int k = 0; k < kChunksPerLane; ++k) {
    const int c = lane + 32 * k;
    lut[k] = c < kChunks ? chunk_lut(c) : 0u;
  }
#define MERLIN_LUT(k) lut[k]
#endif

  const int n_actions = (p.flags & MERLIN_F_SEVEN_ACTIONS) ? 7 : 3;
  const bool mutable_grid = p.cells != nullptr;
  const bool stuck_on = p.flags & MERLIN_F_STUCK_PENALTY;
  const bool explore_on = p.flags & MERLIN_F_EXPLORE_BONUS;
  const bool auto_reset = p.flags & MERLIN_F_AUTO_RESET;
  const bool advance = !(p.flags & MERLIN_F_RESET_SAME);
  const int warps_per_cta = blockDim.x >> 5;

  for (int g = blockIdx.x * warps_per_cta + warp; g < n_groups; g += gridDim.x * warps_per_cta) {
    const int e0 = g * G;
    const int e = e0 + lane;
    const bool active = lane < G && e < p.N;

    // ------------------------------------------------------------------ phase A: one env per lane
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
        const uint8_t* grid = mutable_grid ? p.cells + (size_t)e * p.cell_stride
                                           : p.pool_cells + (size_t)s.layout * p.cell_stride;
        const int fx = s.x + dir_dx(s.dir), fy = s.y + dir_dy(s.dir);
        const bool inb = (unsigned)fx < (unsigned)p.W && (unsigned)fy < (unsigned)p.H;
        const int fidx = fy * p.W + fx;
        const uint32_t fwd = inb ? grid[fidx] : CODE_WALL;
        StepResult r = step_logic(s, p.actions[e], n_actions, fwd, inb, fidx, p.max_steps);
        if (r.bad_action) atomicAdd(p.bad_actions, 1ull);
        if (r.write_idx >= 0 && mutable_grid) p.cells[(size_t)e * p.cell_stride + r.write_idx] = (uint8_t)r.write_code;

        uint32_t vword = 0;
        const int cell = s.y * p.W + s.x;
        uint32_t* vptr = nullptr;
        if (explore_on) { vptr = p.visited + (size_t)e * p.vis_words + (cell >> 5); vword = *vptr; }
        bool stuck = false;
        const uint32_t vword_in = vword;
        const double rew_d = shape_reward(s, r.reward, stuck_on, p.stuck_max_stay, p.stuck_penalty, explore_on,
                                          p.explore_bonus, vword, cell & 31, stuck);
        if (explore_on && vword != vword_in) *vptr = vword;
        const float rew = (float)rew_d;
        ep_ret += rew;
        const bool done = r.terminated || r.truncated;
        p.reward[e] = rew;
        p.terminated[e] = r.terminated ? 1 : 0;
        p.truncated[e] = r.truncated ? 1 : 0;
        if (p.out_ep_return) p.out_ep_return[e] = done ? ep_ret : 0.f;
        if (p.out_ep_length) p.out_ep_length[e] = done ? s.step_count : 0;
        if (p.out_stuck) p.out_stuck[e] = stuck ? 1 : 0;
        restart = done && auto_reset;
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
                                        : (advance ? (int)(((long long)s.layout + p.N) % p.n_layouts) : s.layout);
      if (restart) {
        const uint32_t a = p.pool_agent[load_cur];
        s.x = a & 0xff; s.y = (a >> 8) & 0xff; s.dir = (a >> 16) & 3; s.carry = 0;
        s.step_count = 0; s.stay = 0; s.last_x = s.x; s.last_y = s.y;
        ep_ret = 0.f;
      }
      if (mutable_grid || explore_on) {
        unsigned m = restart_mask;
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int cur = __shfl_sync(0xffffffffu, load_cur, src);
          const int sx = __shfl_sync(0xffffffffu, s.x, src), sy = __shfl_sync(0xffffffffu, s.y, src);
          const size_t ee = (size_t)(e0 + src);
          if (mutable_grid) {
            const int4* from = reinterpret_cast<const int4*>(p.pool_cells + (size_t)cur * p.cell_stride);
            int4* to = reinterpret_cast<int4*>(p.cells + ee * p.cell_stride);
            for (int i = lane; i < p.cell_stride / 16; i += 32) to[i] = from[i];
          }
          if (explore_on) {
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
    if ((want_rgb || want_sym) && render) {
      const uint8_t* grid = mutable_grid ? p.cells + (size_t)e * p.cell_stride
                                         : p.pool_cells + (size_t)s.layout * p.cell_stride;
      uint8_t* kind = kinds_s + lane * kKindStride;
      const uint64_t transp = gather_view(s, p.W, p.H, [&](int idx) -> uint32_t { return grid[idx]; }, kind);
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
          if (want_sym) {
            uint8_t t = 0, col = 0, stt = 0;
            if (seen) sym_of_code(code, t, col, stt);
            sym[c * 3 + 0] = t; sym[c * 3 + 1] = col; sym[c * 3 + 2] = stt;
          }
        }
      }
    }
    __syncwarp();

    const unsigned render_mask = __ballot_sync(0xffffffffu, render);

    // observation, part 2 (whole warp): symbolic rows out, coalesced
    if (want_sym && render_mask) {
      const int n_here = min(G, p.N - e0);
      uint8_t* out = p.obs_sym + (size_t)e0 * kSymBytes;
      const unsigned full = n_here >= 32 ? 0xffffffffu : ((1u << n_here) - 1u);
      if (render_mask == full && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (n_here * kSymBytes) % 16 == 0) {
        const int4* src = reinterpret_cast<const int4*>(sym_s);
        int4* dst = reinterpret_cast<int4*>(out);
        for (int i = lane; i < n_here * kSymBytes / 16; i += 32) dst[i] = src[i];
      } else {
        for (int i = 0; i < n_here; ++i) {
          if (!((render_mask >> i) & 1)) continue;
          for (int b = lane; b < kSymBytes; b += 32) out[i * kSymBytes + b] = sym_s[i * kSymBytes + b];
        }
      }
    }

    // observation, part 3 (whole warp): RGB frames, 588 x 16-byte streaming stores per env
    if (want_rgb && render_mask) {
      const uint2* atlas64 = reinterpret_cast<const uint2*>(atlas_s);
      unsigned m = render_mask;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        const uint8_t* kp = kinds_s + i * kKindStride;
        uint8_t* frame = p.obs_rgb + (size_t)(e0 + i) * kImgBytes;
#pragma unroll
        for (int k = 0; k < kChunksPerLane; ++k) {
          const int c = lane + 32 * k;
          if (c < kChunks) {
            const uint32_t q = MERLIN_LUT(k);
            const uint32_t k0 = kp[q & 0xff], k1 = kp[(q >> 16) & 0xff];
            const uint2 a = atlas64[k0 * (kTileBytes / 8) + ((q >> 8) & 0xff)];
            const uint2 b = atlas64[k1 * (kTileBytes / 8) + (q >> 24)];
            st_stream_v4(frame + c * 16, a.x, a.y, b.x, b.y);
          }
        }
      }
    }
    __syncwarp();  // smem rows are reused by the next group
  }
}

// ---------------------------------------------------------------------------------------------------
template <int G, bool STEP>
static cudaError_t launch_one(const EnvParams& p, int sm_count, cudaStream_t stream) {
  const int n_groups = (p.N + G - 1) / G;
  const size_t smem = cta_smem_bytes(G);
  static bool configured_dev[64] = {};  // per template instance and device
  static int blocks_per_sm_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  bool& configured = configured_dev[dev];
  int& blocks_per_sm = blocks_per_sm_dev[dev];
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(env_kernel<G, STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, env_kernel<G, STEP>, kThreads, smem);
    if (err != cudaSuccess) return err;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  const int ctas_needed = (n_groups + kWarps - 1) / kWarps;
  int grid = sm_count * blocks_per_sm;
  if (grid > ctas_needed) grid = ctas_needed;
  env_kernel<G, STEP><<<grid, kThreads, smem, stream>>>(p, n_groups);
  return cudaGetLastError();
}

template <bool STEP>
static cudaError_t launch_sized(const EnvParams& p, int sm_count, cudaStream_t stream) {
  // pick the largest group size that still yields >= ~8 warps per SM; tiny batches use smaller groups
  const long long want_warps = (long long)sm_count * 8;
  if (p.N / 32 >= want_warps) return launch_one<32, STEP>(p, sm_count, stream);
  if (p.N / 16 >= want_warps) return launch_one<16, STEP>(p, sm_count, stream);
  if (p.N / 8 >= want_warps) return launch_one<8, STEP>(p, sm_count, stream);
  return launch_one<4, STEP>(p, sm_count, stream);
}

cudaError_t launch_env_step(const EnvParams& p, int sm_count, cudaStream_t stream) {
  return launch_sized<true>(p, sm_count, stream);
}
cudaError_t launch_env_reset(const EnvParams& p, int sm_count, cudaStream_t stream) {
  return launch_sized<false>(p, sm_count, stream);
}

}  // namespace merlin
