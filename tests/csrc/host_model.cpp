// TEST VEHICLE (tests/ only) -- compiles the product's per-environment logic header
// (ppo-2dgrid_b200/csrc/env_logic.cuh) for the HOST so the CPU test tier can check the exact functions the
// CUDA kernels call (step_logic, shape_reward, gather_view, visibility, agent_kind, sym_of_code, chunk_lut)
// against the oracle without a GPU.  It is not part of the product and is not shipped or loaded by it.
#include <stdint.h>
#include <string.h>

#include "env_logic.cuh"
#include "obs_swar.cuh"

using namespace merlin;

extern "C" {

// One step (no auto-reset) + observation for N envs. state: int32[N][4] packed like the device state.
// cells: u8[N][cell_stride] private grids (mutated by pickup/drop/toggle). visited: u32[N][vis_words] or NULL.
int hm_step(int N, int W, int H, int max_steps, int cell_stride, int n_actions, int do_step, int stuck_on,
            int max_stay, double penalty, int explore_on, double bonus, int vis_words, int32_t* state, uint8_t* cells,
            uint32_t* visited, const int64_t* actions, const uint8_t* atlas, uint8_t* obs_rgb, uint8_t* obs_sym,
            float* reward, uint8_t* terminated, uint8_t* truncated, uint8_t* stuck_out) {
  for (int e = 0; e < N; ++e) {
    EnvState s{};
    int32_t* st = state + 4 * e;
    unpack_state(st[0], st[1], st[2], st[3], s);
    uint8_t* grid = cells + (size_t)e * cell_stride;
    if (do_step) {
      const int fx = s.x + dir_dx(s.dir), fy = s.y + dir_dy(s.dir);
      const bool inb = (unsigned)fx < (unsigned)W && (unsigned)fy < (unsigned)H;
      const int fidx = fy * W + fx;
      const uint32_t fwd = inb ? grid[fidx] : CODE_WALL;
      StepResult r = step_logic(s, actions[e], n_actions, fwd, inb, fidx, max_steps);
      if (r.write_idx >= 0) grid[r.write_idx] = (uint8_t)r.write_code;
      uint32_t vword = 0;
      const int cell = s.y * W + s.x;
      if (explore_on) vword = visited[(size_t)e * vis_words + (cell >> 5)];
      bool stuck = false;
      const double rd = shape_reward(s, r.reward, stuck_on, max_stay, penalty, explore_on, bonus, vword, cell & 31, stuck);
      if (explore_on) visited[(size_t)e * vis_words + (cell >> 5)] = vword;
      reward[e] = (float)rd;
      terminated[e] = r.terminated;
      truncated[e] = r.truncated;
      stuck_out[e] = stuck;
      pack_state(s, st[0], st[1], st[2], st[3]);
    }
    uint8_t kind[kCells];
    const uint64_t transp = gather_view(s, W, H, [&](int idx) -> uint32_t { return grid[idx]; }, kind);
    const uint64_t vis = visibility(transp);
    for (int vi = 0; vi < kView; ++vi)
      for (int vj = 0; vj < kView; ++vj) {
        const int c = vi * kView + vj;
        const bool seen = (vis >> (vj * kView + vi)) & 1;
        uint32_t code = kind[c];
        const bool agent_cell = (vi == kView / 2 && vj == kView - 1);
        if (agent_cell) code = s.carry ? s.carry : CODE_EMPTY;
        kind[c] = (uint8_t)(agent_cell ? agent_kind(s.carry) : (seen ? code : KIND_UNSEEN));
        uint8_t t = 0, col = 0, stt = 0;
        if (seen) sym_of_code(code, t, col, stt);
        uint8_t* sym = obs_sym + (size_t)e * kSymBytes + c * 3;
        sym[0] = t; sym[1] = col; sym[2] = stt;
      }
    uint8_t* frame = obs_rgb + (size_t)e * kImgBytes;
    for (int c = 0; c < kChunks; ++c) {
      const uint32_t q = chunk_lut(c);
      const uint32_t k0 = kind[q & 0xff], k1 = kind[(q >> 16) & 0xff];
      memcpy(frame + c * 16, atlas + k0 * kTileBytes + ((q >> 8) & 0xff) * 8, 8);
      memcpy(frame + c * 16 + 8, atlas + k1 * kTileBytes + (q >> 24) * 8, 8);
    }
  }
  return 0;
}

// The row-parallel observation (obs_swar.cuh) for N envs: symbolic image u8[N][147] and tile kinds u8[N][49], to be
// compared with what hm_step produced by the per-cell form.  W >= 7; `cells` needs 8 bytes of slack after the last grid.
int hm_observe_swar(int N, int W, int H, int cell_stride, const int32_t* state, const uint8_t* cells,
                    const uint8_t* atlas, uint8_t* obs_rgb, uint8_t* obs_sym, int doors) {
  for (int e = 0; e < N; ++e) {
    EnvState s{};
    const int32_t* st = state + 4 * e;
    unpack_state(st[0], st[1], st[2], st[3], s);
    uint64_t g[kView], seen[kView];
    observe_swar(s, cells + (size_t)e * cell_stride, W, H, g, seen, doors != 0);
    for (int vi = 0; vi < kView; ++vi) {
      uint32_t w[6];
      encode_group(g[vi], w, doors != 0);
      memcpy(obs_sym + (size_t)e * kSymBytes + vi * 21, w, 21);
    }
    uint32_t kw[13];
    kind_words(g, s.carry, kw);
    const uint8_t* kind = reinterpret_cast<const uint8_t*>(kw);
    uint8_t* frame = obs_rgb + (size_t)e * kImgBytes;
    for (int c = 0; c < kChunks; ++c) {
      const uint32_t q = chunk_lut(c);
      const uint32_t k0 = kind[q & 0xff], k1 = kind[(q >> 16) & 0xff];
      memcpy(frame + c * 16, atlas + k0 * kTileBytes + ((q >> 8) & 0xff) * 8, 8);
      memcpy(frame + c * 16 + 8, atlas + k1 * kTileBytes + (q >> 24) * 8, 8);
    }
  }
  return 0;
}

// The in-kernel action sampler (env_logic.cuh: philox4x32_10_x0, sampler_uniform, sample_policy) for N envs.
void hm_philox(const uint32_t* counter4, const uint32_t* key2, uint32_t* out4) {
  philox4x32_10_x0(counter4[0], counter4[1], counter4[2], counter4[3], key2[0], key2[1], out4);
}
int hm_sample(int N, int n_actions, const float* logits, uint64_t seed, const uint32_t* draws, int greedy,
              int64_t* action, float* logp, float* uniform) {
  for (int e = 0; e < N; ++e) {
    float lg[kMaxActions];
    for (int a = 0; a < kMaxActions; ++a) lg[a] = a < n_actions ? logits[(size_t)e * n_actions + a] : 0.f;
    const float u = greedy ? 0.f : sampler_uniform((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)e, draws[e]);
    const PolicySample s = sample_policy(lg, n_actions, u, greedy != 0);
    action[e] = s.action; logp[e] = s.logp; uniform[e] = u;
  }
  return 0;
}

// Grid.process_vis on a 49-bit transparency mask (bit vj*7 + vi), for the property tests.
unsigned long long hm_visibility(unsigned long long transp) { return merlin::visibility(transp); }
unsigned long long hm_visibility_literal(unsigned long long transp) { return merlin::visibility_literal(transp); }
// The byte-per-row form env_kernel_quad's lanes exchange (row vj in bits 8*vj .. 8*vj+6, in and out).
unsigned long long hm_visibility_rows(unsigned long long transp_rows) { return merlin::visibility_rows(transp_rows); }
// One process_vis row, both forms: returns lit | next_seed << 8.
unsigned hm_vis_row(unsigned seed, unsigned T, int literal) {
  uint32_t lit = 0;
  const uint32_t next = literal ? merlin::vis_row_literal(seed, T, lit) : merlin::vis_row(seed, T, lit);
  return lit | (next << 8);
}

}  // extern "C"
