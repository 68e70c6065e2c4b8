/* cabi_smoke.c -- the C ABI used from plain C (no Python, no torch): create a handle, upload two hand-made layouts,
 * reset, step with CUDA-runtime device buffers, read results back, GAE on a tiny rollout.
 *   gcc tests/csrc/cabi_smoke.c -Iinclude -I/usr/local/cuda/include -Lppo-2dgrid_b200/lib -lmerlin_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,ppo-2dgrid_b200/lib -o tests/csrc/_build/cabi_smoke
 * Exit code 0 and "cabi_smoke ok" on success.  Known answers: an agent at (1,1) facing right in an empty
 * 8x8 room with the goal two cells ahead reaches it on step 2: reward = 1 - 0.9 * 2 / 256, terminated. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "merlin_b200.h"

#define CHECK(call)                                                                         \
  do {                                                                                      \
    int rc_ = (call);                                                                       \
    if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, merlin_last_error()); return 1; } \
  } while (0)
#define CUDA(call)                                                                          \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } \
  } while (0)

enum { W = 8, H = 8, N = 3, L = 2 };

int main(void) {
  merlin_env_config_t cfg;
  merlin_env_default_config(&cfg);
  cfg.n_envs = N; cfg.width = W; cfg.height = H;  /* max_steps 0 -> 4*W*H = 256 */
  merlin_env_t* h = NULL;
  CHECK(merlin_env_create(&cfg, &h));

  uint8_t cells[L][W * H];
  int32_t agent[L][3] = {{1, 1, 0}, {1, 1, 1}};
  const uint8_t wall = merlin_pack_cell(2, 5, 0), empty = merlin_pack_cell(1, 0, 0), goal = merlin_pack_cell(8, 1, 0);
  for (int l = 0; l < L; ++l)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        cells[l][y * W + x] = (x == 0 || y == 0 || x == W - 1 || y == H - 1) ? wall : empty;
  cells[0][1 * W + 3] = goal;  /* layout 0: goal two cells to the right of the agent */
  cells[1][3 * W + 1] = goal;  /* layout 1: goal two cells below the agent (which faces down) */
  CHECK(merlin_env_upload_layouts(h, &cells[0][0], &agent[0][0], L));

  uint8_t *d_sym, *d_term, *d_trunc;
  int64_t* d_act;
  float* d_rew;
  CUDA(cudaMalloc((void**)&d_sym, N * 147));
  CUDA(cudaMalloc((void**)&d_term, N));
  CUDA(cudaMalloc((void**)&d_trunc, N));
  CUDA(cudaMalloc((void**)&d_act, N * sizeof(int64_t)));
  CUDA(cudaMalloc((void**)&d_rew, N * sizeof(float)));
  CHECK(merlin_env_reset(h, NULL, NULL, d_sym, NULL));

  uint8_t sym[N][7][7][3];
  CUDA(cudaMemcpy(sym, d_sym, sizeof sym, cudaMemcpyDeviceToHost));
  /* view cell (3,6) is the agent's own cell (encoded as empty), (3,4) is two cells ahead: the goal (8,1,0) */
  if (sym[0][3][6][0] != 1 || sym[0][3][4][0] != 8 || sym[0][3][4][1] != 1 || sym[1][3][4][0] != 8) {
    fprintf(stderr, "unexpected reset observation\n");
    return 1;
  }
  const int64_t forward[N] = {2, 2, 2};
  float rew[N];
  uint8_t term[N], trunc[N];
  for (int t = 1; t <= 2; ++t) {
    CUDA(cudaMemcpy(d_act, forward, sizeof forward, cudaMemcpyHostToDevice));
    CHECK(merlin_env_step(h, d_act, NULL, d_sym, d_rew, d_term, d_trunc, NULL, NULL));
    CUDA(cudaMemcpy(rew, d_rew, sizeof rew, cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(term, d_term, sizeof term, cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(trunc, d_trunc, sizeof trunc, cudaMemcpyDeviceToHost));
    for (int e = 0; e < N; ++e) {
      const float want = t == 2 ? (float)(1.0 - 0.9 * (2.0 / 256.0)) : 0.0f;
      if (rew[e] != want || term[e] != (t == 2) || trunc[e] != 0) {
        fprintf(stderr, "step %d env %d: reward %.9g terminated %d truncated %d (want %.9g)\n", t, e, rew[e], term[e], trunc[e], want);
        return 1;
      }
    }
  }
  int32_t state[N][4];
  CHECK(merlin_env_read_state(h, &state[0][0], NULL, NULL));
  /* auto-reset: env 0 advanced its cursor by N = 3 -> layout (0 + 3) % 2 = 1; fresh episode, step_count 0 */
  if (state[0][2] != 1 || state[0][1] != 0 || state[1][2] != 0) { fprintf(stderr, "unexpected cursors after auto-reset\n"); return 1; }

  /* GAE, T = 2, one env, no dones: adv1 = r1 + g*last - v1 ; adv0 = r0 + g*v1 - v0 + g*lam*adv1 */
  const float r[2] = {0.5f, 1.0f}, v[2] = {0.25f, -0.5f}, d[2] = {0.f, 0.f}, last = 2.0f;
  float *d_r, *d_v, *d_d, *d_l, *d_a, *d_ret, adv[2], ret[2];
  CUDA(cudaMalloc((void**)&d_r, 8)); CUDA(cudaMalloc((void**)&d_v, 8)); CUDA(cudaMalloc((void**)&d_d, 8));
  CUDA(cudaMalloc((void**)&d_l, 4)); CUDA(cudaMalloc((void**)&d_a, 8)); CUDA(cudaMalloc((void**)&d_ret, 8));
  CUDA(cudaMemcpy(d_r, r, 8, cudaMemcpyHostToDevice)); CUDA(cudaMemcpy(d_v, v, 8, cudaMemcpyHostToDevice));
  CUDA(cudaMemcpy(d_d, d, 8, cudaMemcpyHostToDevice)); CUDA(cudaMemcpy(d_l, &last, 4, cudaMemcpyHostToDevice));
  CHECK(merlin_gae(d_r, d_v, d_d, d_l, d_a, d_ret, 2, 1, 0.99, 0.95, NULL));
  CUDA(cudaMemcpy(adv, d_a, 8, cudaMemcpyDeviceToHost)); CUDA(cudaMemcpy(ret, d_ret, 8, cudaMemcpyDeviceToHost));
  const double a1 = 1.0 + 0.99 * 2.0 + 0.5, a0 = 0.5 + 0.99 * -0.5 - 0.25 + 0.99 * 0.95 * a1;
  if (fabs(adv[1] - a1) > 1e-5 || fabs(adv[0] - a0) > 1e-5 || fabs(ret[0] - (0.25 + a0)) > 1e-5) {
    fprintf(stderr, "GAE mismatch: %g %g vs %g %g\n", adv[0], adv[1], a0, a1);
    return 1;
  }
  if (merlin_env_step(h, NULL, NULL, NULL, d_rew, d_term, d_trunc, NULL, NULL) != MERLIN_EINVAL) { fprintf(stderr, "NULL actions accepted\n"); return 1; }

  /* The fused transition: policy outputs in, sampled action + log-probability + value rows out, env stepped, all in one
   * launch.  Greedy on logits (0, 0, 5) = `forward` twice from a fresh reset: the goal again on step 2; log-probability
   * = -log(1 + 2 e^-5); the first-episode record holds return / length / goal flag of that episode. */
  {
    float *d_logits, *d_val, *d_logp, *d_vout, *d_fret;
    int32_t* d_flen;
    uint8_t *d_fin, *d_fgoal;
    const float logits[N][3] = {{0, 0, 5}, {0, 0, 5}, {0, 0, 5}}, val[N] = {0.5f, 1.5f, 2.5f};
    CUDA(cudaMalloc((void**)&d_logits, sizeof logits)); CUDA(cudaMalloc((void**)&d_val, sizeof val));
    CUDA(cudaMalloc((void**)&d_logp, N * 4)); CUDA(cudaMalloc((void**)&d_vout, N * 4)); CUDA(cudaMalloc((void**)&d_fret, N * 4));
    CUDA(cudaMalloc((void**)&d_flen, N * 4)); CUDA(cudaMalloc((void**)&d_fin, N)); CUDA(cudaMalloc((void**)&d_fgoal, N));
    CUDA(cudaMemcpy(d_logits, logits, sizeof logits, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(d_val, val, sizeof val, cudaMemcpyHostToDevice));
    CUDA(cudaMemset(d_fin, 0, N));
    merlin_policy_io_t io;
    memset(&io, 0, sizeof io);
    io.logits = d_logits; io.value = d_val; io.action = d_act; io.logprob = d_logp; io.value_out = d_vout; io.greedy = 1;
    io.finished = d_fin; io.first_return = d_fret; io.first_length = d_flen; io.first_goal = d_fgoal;
    CHECK(merlin_env_seed_sampler(h, 42));
    const int32_t cursors[N] = {0, 1, 0};
    CHECK(merlin_env_set_cursors(h, cursors));
    CHECK(merlin_env_reset(h, NULL, NULL, d_sym, NULL));
    for (int t = 1; t <= 2; ++t)
      CHECK(merlin_env_policy_step(h, &io, NULL, d_sym, d_rew, d_term, d_trunc, NULL, NULL));
    int64_t act[N]; float logp[N], vout[N], fret[N]; int32_t flen[N]; uint8_t fin[N], fgoal[N];
    CUDA(cudaMemcpy(act, d_act, sizeof act, cudaMemcpyDeviceToHost)); CUDA(cudaMemcpy(logp, d_logp, sizeof logp, cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(vout, d_vout, sizeof vout, cudaMemcpyDeviceToHost)); CUDA(cudaMemcpy(fret, d_fret, sizeof fret, cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(flen, d_flen, sizeof flen, cudaMemcpyDeviceToHost)); CUDA(cudaMemcpy(fin, d_fin, N, cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(fgoal, d_fgoal, N, cudaMemcpyDeviceToHost)); CUDA(cudaMemcpy(term, d_term, N, cudaMemcpyDeviceToHost));
    const double want_lp = -log(1.0 + 2.0 * exp(-5.0));
    for (int e = 0; e < N; ++e)
      if (act[e] != 2 || fabs(logp[e] - want_lp) > 1e-6 || vout[e] != val[e] || !term[e] || !fin[e] || !fgoal[e] || flen[e] != 2 ||
          fret[e] != (float)(1.0 - 0.9 * (2.0 / 256.0))) {
        fprintf(stderr, "policy_step env %d: action %lld logp %.9g value %g finished %d goal %d length %d return %.9g\n", e,
                (long long)act[e], logp[e], vout[e], fin[e], fgoal[e], flen[e], fret[e]);
        return 1;
      }
    io.logits = NULL;
    if (merlin_env_policy_step(h, &io, NULL, d_sym, d_rew, d_term, d_trunc, NULL, NULL) != MERLIN_EINVAL) { fprintf(stderr, "NULL logits accepted\n"); return 1; }
    CHECK(merlin_env_rearm(h, NULL));
  }
  CHECK(merlin_env_destroy(h));
  printf("cabi_smoke ok (%s)\n", merlin_version());
  return 0;
}
