/* ORACLE (test infrastructure, NOT product code) -- plain-C restatement of the per-step
 * arithmetic of the rollout hot path, batched over independent environments.
 *
 * It follows the SAME literal algorithm as oracle/minigrid_restated.py (slice -> rotate_left x (dir+1)
 * -> process_vis sweeps -> erase -> agent cell -> encode / tile blit), cell by cell, so that it can
 * check the CUDA kernels (which use closed-form index maps and bitmask visibility) at sizes the
 * Python oracle cannot reach.  tests/test_oracle_fast.py pins it to the Python oracle and to the
 * golden fixtures.  PARITY STATUS of the upstream (minigrid 3.0.0) semantics: unpinned, see
 * oracle/minigrid_restated.py.
 *
 * Reference anchors: MiniGridEnv.step / gen_obs / get_frame are reached from src/ppo.py:76 and
 * src/fomaml.py:71 through the wrapper stack of src/scenario_creator/scenario_creator.py:43-55;
 * env constants from src/custom_envs/base_env.py:32-41.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T_UNSEEN 0
#define T_EMPTY 1
#define T_WALL 2
#define T_FLOOR 3
#define T_DOOR 4
#define T_KEY 5
#define T_BALL 6
#define T_BOX 7
#define T_GOAL 8
#define T_LAVA 9
#define T_AGENT 10

#define S_OPEN 0
#define S_CLOSED 1
#define S_LOCKED 2

#define MAXV 15

typedef struct { uint8_t type, color, state; } cell_t; /* type==T_EMPTY means "None" */

static const int DIRX[4] = {1, 0, -1, 0};
static const int DIRY[4] = {0, 1, 0, -1};

static int can_overlap(cell_t c) {
  return c.type == T_GOAL || c.type == T_FLOOR || c.type == T_LAVA || (c.type == T_DOOR && c.state == S_OPEN);
}
static int can_pickup(cell_t c) { return c.type == T_KEY || c.type == T_BALL || c.type == T_BOX; }
static int see_behind(cell_t c) {
  if (c.type == T_WALL) return 0;
  if (c.type == T_DOOR) return c.state == S_OPEN;
  return 1;
}
static int is_none(cell_t c) { return c.type == T_EMPTY; }

/* tile atlas: dense [11 types][6 colors][3 states][2 agent][2 highlight][T*T*3], with a presence map */
static inline long tile_index(int type, int color, int state, int agent, int hl) {
  return ((((long)type * 6 + color) * 3 + state) * 2 + agent) * 2 + hl;
}

typedef struct {
  int W, H, max_steps, V, T;
  const uint8_t *tiles;         /* may be NULL when rgb not requested */
  const uint8_t *tiles_present; /* [792] */
} cfg_t;

/* One environment: dynamics (MiniGridEnv.step). Returns 0 ok, <0 on invalid action. */
static int env_step(const cfg_t *cfg, uint8_t *gt, uint8_t *gc, uint8_t *gs, int32_t *ax, int32_t *ay,
                    int32_t *adir, int32_t *stepc, uint8_t *carry_t, uint8_t *carry_c, int64_t action,
                    double *reward, uint8_t *terminated, uint8_t *truncated) {
  const int W = cfg->W, H = cfg->H;
  *stepc += 1;
  *reward = 0.0;
  *terminated = 0;
  *truncated = 0;
  int fx = *ax + DIRX[*adir], fy = *ay + DIRY[*adir];
  int inb = fx >= 0 && fx < W && fy >= 0 && fy < H;
  cell_t fwd = {T_WALL, 5, 0}; /* upstream would assert; treat out-of-grid as a wall */
  long fi = (long)fy * W + fx;
  if (inb) { fwd.type = gt[fi]; fwd.color = gc[fi]; fwd.state = gs[fi]; }
  switch (action) {
    case 0: *adir -= 1; if (*adir < 0) *adir += 4; break;
    case 1: *adir = (*adir + 1) % 4; break;
    case 2:
      if (is_none(fwd) || can_overlap(fwd)) { *ax = fx; *ay = fy; }
      if (!is_none(fwd) && fwd.type == T_GOAL) {
        *terminated = 1;
        *reward = 1 - 0.9 * ((double)*stepc / (double)cfg->max_steps);
      }
      if (!is_none(fwd) && fwd.type == T_LAVA) *terminated = 1;
      break;
    case 3:
      if (!is_none(fwd) && can_pickup(fwd) && inb) {
        if (*carry_t == 0) { *carry_t = fwd.type; *carry_c = fwd.color; gt[fi] = T_EMPTY; gc[fi] = 0; gs[fi] = 0; }
      }
      break;
    case 4:
      if (is_none(fwd) && *carry_t != 0 && inb) { gt[fi] = *carry_t; gc[fi] = *carry_c; gs[fi] = 0; *carry_t = 0; *carry_c = 0; }
      break;
    case 5:
      if (!is_none(fwd) && inb) {
        if (fwd.type == T_DOOR) {
          if (fwd.state == S_LOCKED) {
            if (*carry_t == T_KEY && *carry_c == fwd.color) gs[fi] = S_OPEN;
          } else {
            gs[fi] = (fwd.state == S_OPEN) ? S_CLOSED : S_OPEN;
          }
        } else if (fwd.type == T_BOX) { /* boxes here never contain anything: replaced by None */
          gt[fi] = T_EMPTY; gc[fi] = 0; gs[fi] = 0;
        }
      }
      break;
    case 6: break;
    default: return -1;
  }
  if (*stepc >= cfg->max_steps) *truncated = 1;
  return 0;
}

/* gen_obs_grid: literal slice + rotate_left + process_vis + agent cell. view[i][j], vis[i][j]. */
static void env_view(const cfg_t *cfg, const uint8_t *gt, const uint8_t *gc, const uint8_t *gs, int ax, int ay,
                     int adir, uint8_t carry_t, uint8_t carry_c, cell_t view[MAXV][MAXV], uint8_t vis[MAXV][MAXV]) {
  const int W = cfg->W, H = cfg->H, V = cfg->V;
  int topX, topY;
  switch (adir) { /* get_view_exts */
    case 0: topX = ax; topY = ay - V / 2; break;
    case 1: topX = ax - V / 2; topY = ay; break;
    case 2: topX = ax - V + 1; topY = ay - V / 2; break;
    default: topX = ax - V / 2; topY = ay - V + 1; break;
  }
  cell_t a[MAXV][MAXV], b[MAXV][MAXV];
  for (int j = 0; j < V; j++)
    for (int i = 0; i < V; i++) { /* Grid.slice */
      int x = topX + i, y = topY + j;
      cell_t c = {T_WALL, 5, 0};
      if (x >= 0 && x < W && y >= 0 && y < H) {
        long k = (long)y * W + x;
        c.type = gt[k]; c.color = gc[k]; c.state = gs[k];
      }
      a[i][j] = c;
    }
  for (int r = 0; r < adir + 1; r++) { /* Grid.rotate_left: new(j, V-1-i) = old(i,j) */
    for (int i = 0; i < V; i++)
      for (int j = 0; j < V; j++) b[j][V - 1 - i] = a[i][j];
    memcpy(a, b, sizeof(a));
  }
  /* Grid.process_vis */
  memset(vis, 0, MAXV * MAXV);
  vis[V / 2][V - 1] = 1;
  for (int j = V - 1; j >= 0; j--) {
    for (int i = 0; i < V - 1; i++) {
      if (!vis[i][j]) continue;
      if (!is_none(a[i][j]) && !see_behind(a[i][j])) continue;
      vis[i + 1][j] = 1;
      if (j > 0) { vis[i + 1][j - 1] = 1; vis[i][j - 1] = 1; }
    }
    for (int i = V - 1; i >= 1; i--) {
      if (!vis[i][j]) continue;
      if (!is_none(a[i][j]) && !see_behind(a[i][j])) continue;
      vis[i - 1][j] = 1;
      if (j > 0) { vis[i - 1][j - 1] = 1; vis[i][j - 1] = 1; }
    }
  }
  for (int j = 0; j < V; j++)
    for (int i = 0; i < V; i++)
      if (!vis[i][j]) { a[i][j].type = T_EMPTY; a[i][j].color = 0; a[i][j].state = 0; }
  /* agent cell shows what is carried, else nothing */
  if (carry_t) { a[V / 2][V - 1].type = carry_t; a[V / 2][V - 1].color = carry_c; a[V / 2][V - 1].state = 0; }
  else { a[V / 2][V - 1].type = T_EMPTY; a[V / 2][V - 1].color = 0; a[V / 2][V - 1].state = 0; }
  memcpy(view, a, sizeof(a));
}

/* returns 0 ok, -2 if a needed tile is not in the atlas */
static int env_obs(const cfg_t *cfg, const uint8_t *gt, const uint8_t *gc, const uint8_t *gs, int ax, int ay, int adir,
                   uint8_t carry_t, uint8_t carry_c, uint8_t *sym /* [V][V][3] or NULL */,
                   uint8_t *rgb /* [V*T][V*T][3] or NULL */) {
  const int V = cfg->V, T = cfg->T;
  cell_t view[MAXV][MAXV];
  uint8_t vis[MAXV][MAXV];
  env_view(cfg, gt, gc, gs, ax, ay, adir, carry_t, carry_c, view, vis);
  if (sym) { /* Grid.encode(vis_mask): array[i][j][c] */
    for (int i = 0; i < V; i++)
      for (int j = 0; j < V; j++) {
        uint8_t *p = sym + ((long)i * V + j) * 3;
        if (vis[i][j]) { p[0] = view[i][j].type; p[1] = view[i][j].color; p[2] = view[i][j].state; }
        else { p[0] = 0; p[1] = 0; p[2] = 0; }
      }
  }
  if (rgb) { /* second gen_obs_grid + Grid.render(tile, agent_pos=(V/2,V-1), agent_dir=3, highlight=vis) */
    const long rowbytes = (long)V * T * 3;
    for (int j = 0; j < V; j++)
      for (int i = 0; i < V; i++) {
        int agent = (i == V / 2 && j == V - 1);
        long ti = tile_index(view[i][j].type, view[i][j].color, view[i][j].state, agent, vis[i][j] ? 1 : 0);
        if (!cfg->tiles_present[ti]) return -2;
        const uint8_t *tile = cfg->tiles + ti * (long)T * T * 3;
        for (int py = 0; py < T; py++)
          memcpy(rgb + ((long)j * T + py) * rowbytes + (long)i * T * 3, tile + (long)py * T * 3, (size_t)T * 3);
      }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * exported entry points (ctypes)
 * ---------------------------------------------------------------------------------------- */
int fo_step(int N, int W, int H, int max_steps, int V, int T, uint8_t *gt, uint8_t *gc, uint8_t *gs, int32_t *ax,
            int32_t *ay, int32_t *adir, int32_t *stepc, uint8_t *carry_t, uint8_t *carry_c, const int64_t *actions,
            double *reward, uint8_t *terminated, uint8_t *truncated) {
  if (V > MAXV) return -3;
  cfg_t cfg = {W, H, max_steps, V, T, NULL, NULL};
  int err = 0;
#pragma omp parallel for schedule(static) reduction(min : err)
  for (int e = 0; e < N; e++) {
    long g = (long)e * W * H;
    int r = env_step(&cfg, gt + g, gc + g, gs + g, ax + e, ay + e, adir + e, stepc + e, carry_t + e, carry_c + e,
                     actions[e], reward + e, terminated + e, truncated + e);
    if (r < err) err = r;
  }
  return err;
}

int fo_obs(int N, int W, int H, int V, int T, const uint8_t *gt, const uint8_t *gc, const uint8_t *gs,
           const int32_t *ax, const int32_t *ay, const int32_t *adir, const uint8_t *carry_t, const uint8_t *carry_c,
           const uint8_t *tiles, const uint8_t *tiles_present, uint8_t *sym, uint8_t *rgb) {
  if (V > MAXV) return -3;
  cfg_t cfg = {W, H, 0, V, T, tiles, tiles_present};
  int err = 0;
  const long symsz = (long)V * V * 3, rgbsz = (long)V * T * V * T * 3;
#pragma omp parallel for schedule(static) reduction(min : err)
  for (int e = 0; e < N; e++) {
    long g = (long)e * W * H;
    int r = env_obs(&cfg, gt + g, gc + g, gs + g, ax[e], ay[e], adir[e], carry_t[e], carry_c[e],
                    sym ? sym + e * symsz : NULL, rgb ? rgb + e * rgbsz : NULL);
    /* the reference wrapper stack evaluates gen_obs_grid twice per step (gen_obs, then get_frame);
       the second evaluation is identical, so it is not repeated here */
    if (r < err) err = r;
  }
  return err;
}

int fo_set_threads(int n) {
#ifdef _OPENMP
  extern void omp_set_num_threads(int);
  if (n > 0) omp_set_num_threads(n);
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* GAE, literal fp32 reverse loop of src/ppo.py:107-120 for N independent columns of a [T][N] rollout.
 * gamma*last_value and gamma*lam are formed in double exactly as python does before meeting an f32. */
void fo_gae(int T, int N, const float *rew, const float *val, const float *done, const float *last_val, double gamma,
            double lam, float *adv, float *ret) {
  const float g = (float)gamma, gl = (float)(gamma * lam);
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) {
    float gae = 0.0f;
    for (int t = T - 1; t >= 0; t--) {
      long k = (long)t * N + n;
      float mask = 1.0f - done[k];
      volatile float nv = (t == T - 1) ? (float)(gamma * (double)last_val[n]) : g * val[k + N];
      volatile float a = nv * mask;
      volatile float b = rew[k] + a;
      volatile float delta = b - val[k];
      volatile float c = gl * mask;
      volatile float d = c * gae;
      gae = delta + d;
      adv[k] = gae;
      ret[k] = val[k] + gae;
    }
  }
}
