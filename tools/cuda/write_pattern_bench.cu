// write_pattern_bench.cu -- how fast can 9408-byte frames be streamed to HBM under different work mappings?
// (development aid for the frame phase of env_kernels.cu; no env logic, constant data, st.global.cs.v4 stores)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cuda/wpb tools/cuda/write_pattern_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

constexpr int kImg = 9408, kChunks = 588;

__device__ __forceinline__ void st_cs(void* p, uint32_t a) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(a) : "memory");
}
__device__ __forceinline__ void st_plain(void* p, uint32_t a) {
  asm volatile("st.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(a) : "memory");
}

// V0: flat -- thread i writes chunk i (what a vectorised fill does)
template <bool CS> __global__ void k_flat(uint8_t* out, size_t n_chunks) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_chunks) { if (CS) st_cs(out + i * 16, (uint32_t)i); else st_plain(out + i * 16, (uint32_t)i); }
}
// V1: one warp per frame, persistent grid
__global__ void k_warp_frame(uint8_t* out, int n_frames) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n_frames; f += warps) {
    uint8_t* frame = out + (size_t)f * kImg;
#pragma unroll
    for (int k = 0; k < 19; ++k) { const int c = lane + 32 * k; if (c < kChunks) st_cs(frame + c * 16, f); }
  }
}
// V2: CTA tile of T frames, warps share them; optional busy-wait by warp 0 before each tile (models the state phase)
__global__ void k_tile(uint8_t* out, int n_frames, int T, int delay_cycles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int n_tiles = (n_frames + T - 1) / T;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (delay_cycles > 0) {
      if (warp == 0) { const long long t0 = clock64(); while (clock64() - t0 < delay_cycles) {} }
      __syncthreads();
    }
    for (int i = warp; i < T; i += wpc) {
      const int f = tile * T + i;
      if (f >= n_frames) break;
      uint8_t* frame = out + (size_t)f * kImg;
#pragma unroll
      for (int k = 0; k < 19; ++k) { const int c = lane + 32 * k; if (c < kChunks) st_cs(frame + c * 16, f); }
    }
    if (delay_cycles > 0) __syncthreads();
  }
}
// V3: non-persistent: one CTA (8 warps) per 8 frames, in launch order
__global__ void k_cta8(uint8_t* out, int n_frames) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.x * 8 + warp;
  if (f >= n_frames) return;
  uint8_t* frame = out + (size_t)f * kImg;
#pragma unroll
  for (int k = 0; k < 19; ++k) { const int c = lane + 32 * k; if (c < kChunks) st_cs(frame + c * 16, f); }
}

// V4: tile per CTA, NON-persistent (grid = n_tiles, in launch order), optional warp-0 delay
__global__ void k_tile_np(uint8_t* out, int n_frames, int T, int delay_cycles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int tile = blockIdx.x;
  if (delay_cycles > 0) {
    if (warp == 0) { const long long t0 = clock64(); while (clock64() - t0 < delay_cycles) {} }
    __syncthreads();
  }
  for (int i = warp; i < T; i += wpc) {
    const int f = tile * T + i;
    if (f >= n_frames) break;
    uint8_t* frame = out + (size_t)f * kImg;
#pragma unroll
    for (int k = 0; k < 19; ++k) { const int c = lane + 32 * k; if (c < kChunks) st_cs(frame + c * 16, f); }
  }
}
// V5: persistent CTAs fetching tiles IN ORDER from an atomic counter (self-resetting), optional warp-0 delay
__global__ void k_tile_dyn(uint8_t* out, int n_frames, int T, int delay_cycles, unsigned* sched) {
  __shared__ int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int n_tiles = (n_frames + T - 1) / T;
  for (;;) {
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(&sched[0], 1u);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    if (delay_cycles > 0) {
      if (warp == 0) { const long long t0 = clock64(); while (clock64() - t0 < delay_cycles) {} }
      __syncthreads();
    }
    for (int i = warp; i < T; i += wpc) {
      const int f = tile * T + i;
      if (f >= n_frames) break;
      uint8_t* frame = out + (size_t)f * kImg;
#pragma unroll
      for (int k = 0; k < 19; ++k) { const int c = lane + 32 * k; if (c < kChunks) st_cs(frame + c * 16, f); }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && atomicAdd(&sched[1], 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
}

// V6: in-order tiles + the per-env side traffic of the real step kernel: warp 0 reads state (16 B), action (8 B) and
// episode return (4 B) per env and writes state, return, reward and two flags back.  HINT: 0 plain, 1 = L2 evict_last
// policy on those side accesses (keep the 30 MB of per-env state resident in L2 across launches).
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <int HINT>
__global__ void k_tile_side(uint8_t* out, int n_frames, int T, unsigned* sched, int4* state, const long long* actions,
                            float* ep_ret, float* reward, uint8_t* term, uint8_t* trunc) {
  __shared__ int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int n_tiles = (n_frames + T - 1) / T;
  const uint64_t pol = policy_evict_last();
  for (;;) {
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(&sched[0], 1u);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    if (warp == 0 && lane < T) {
      const int e = tile * T + lane;
      if (e < n_frames) {
        int4 st; long long a; float r;
        if (HINT) {
          asm volatile("ld.global.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(st.x), "=r"(st.y), "=r"(st.z), "=r"(st.w) : "l"(state + e), "l"(pol));
          asm volatile("ld.global.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(a) : "l"(actions + e), "l"(pol));
          asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(ep_ret + e), "l"(pol));
        } else { st = state[e]; a = actions[e]; r = ep_ret[e]; }
        st.y += (int)a; r += 1.0f;
        if (HINT) {
          asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(state + e), "r"(st.x), "r"(st.y), "r"(st.z), "r"(st.w), "l"(pol) : "memory");
          asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(ep_ret + e), "f"(r), "l"(pol) : "memory");
        } else { state[e] = st; ep_ret[e] = r; }
        reward[e] = r; term[e] = 0; trunc[e] = (uint8_t)(st.y & 1);
      }
    }
    __syncthreads();
    for (int i = warp; i < T; i += wpc) {
      const int f = tile * T + i;
      if (f >= n_frames) break;
      uint8_t* frame = out + (size_t)f * kImg;
#pragma unroll
      for (int k = 0; k < 19; ++k) { const int c = lane + 32 * k; if (c < kChunks) st_cs(frame + c * 16, f); }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && atomicAdd(&sched[1], 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
}

template <typename F> float time_ms(F launch, int reps = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 1 << 20;
  const size_t bytes = (size_t)N * kImg;
  uint8_t* out; cudaMalloc(&out, bytes);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto report = [&](const char* name, float ms) { printf("%-46s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
  const size_t n_chunks = (size_t)N * kChunks;
  report("flat, plain st", time_ms([&] { k_flat<false><<<(unsigned)((n_chunks + 255) / 256), 256>>>(out, n_chunks); }));
  report("flat, st.cs", time_ms([&] { k_flat<true><<<(unsigned)((n_chunks + 255) / 256), 256>>>(out, n_chunks); }));
  for (int wps : {8, 16, 24, 32, 48, 64}) {
    char name[96]; snprintf(name, sizeof name, "warp per frame, persistent, %d warps/SM", wps);
    report(name, time_ms([&] { k_warp_frame<<<sms * wps / 8, 256>>>(out, N); }));
  }
  report("CTA(8 warps) per 8 frames, non-persistent", time_ms([&] { k_cta8<<<(N + 7) / 8, 256>>>(out, N); }));
  for (int T : {8, 16, 32, 64}) for (int cps : {1, 2, 3, 4}) {
    char name[96]; snprintf(name, sizeof name, "tile T=%d, %d CTAs/SM x 8 warps, no delay", T, cps);
    report(name, time_ms([&] { k_tile<<<sms * cps, 256>>>(out, N, T, 0); }));
  }
  for (int delay : {2000, 5000, 10000}) for (int cps : {2, 3, 4}) {
    char name[96]; snprintf(name, sizeof name, "tile T=32, %d CTAs/SM, warp0 busy %d cycles/tile", cps, delay);
    report(name, time_ms([&] { k_tile<<<sms * cps, 256>>>(out, N, 32, delay); }));
  }
  unsigned* sched; cudaMalloc(&sched, 8); cudaMemset(sched, 0, 8);
  for (int delay : {0, 2000, 5000, 10000}) {
    char name[96];
    for (int T : {16, 32}) {
      snprintf(name, sizeof name, "tile T=%d NON-persistent, warp0 busy %d", T, delay);
      report(name, time_ms([&] { k_tile_np<<<(N + T - 1) / T, 256>>>(out, N, T, delay); }));
    }
    for (int cps : {2, 3, 4}) {
      snprintf(name, sizeof name, "tile T=32 dynamic in-order, %d CTAs/SM, busy %d", cps, delay);
      report(name, time_ms([&] { k_tile_dyn<<<sms * cps, 256>>>(out, N, 32, delay, sched); }));
    }
  }
  int4* state; long long* actions; float *ep_ret, *reward; uint8_t *term, *trunc;
  cudaMalloc(&state, (size_t)N * 16); cudaMalloc(&actions, (size_t)N * 8); cudaMalloc(&ep_ret, (size_t)N * 4);
  cudaMalloc(&reward, (size_t)N * 4); cudaMalloc(&term, N); cudaMalloc(&trunc, N);
  cudaMemset(state, 0, (size_t)N * 16); cudaMemset(actions, 0, (size_t)N * 8); cudaMemset(ep_ret, 0, (size_t)N * 4);
  report("tile T=16 in-order 128thr x4 + side traffic, plain", time_ms([&] { k_tile_side<0><<<sms * 4, 128>>>(out, N, 16, sched, state, actions, ep_ret, reward, term, trunc); }));
  report("tile T=16 in-order 128thr x4 + side traffic, evict_last", time_ms([&] { k_tile_side<1><<<sms * 4, 128>>>(out, N, 16, sched, state, actions, ep_ret, reward, term, trunc); }));
  report("tile T=16 in-order 128thr x4, no side traffic", time_ms([&] { k_tile_dyn<<<sms * 4, 128>>>(out, N, 16, 0, sched); }));
  // the same with the 16 MB state array pinned in L2 by a persisting access-policy window on the launching stream
  // (state is re-read and re-written every step; ep_ret / actions / outputs stay ordinary traffic).
  // MEASURED (B200): setting aside a 48 MB persisting carve-out drops BOTH this run and a plain run after it from
  // 7.18 TB/s to 5.19 TB/s -- the carve-out takes L2 away from the write stream -- so the library never touches
  // cudaLimitPersistingL2CacheSize; evict_last hints (no carve-out) are worth +0.6 % here and were not adopted either.
  {
    cudaStream_t st; cudaStreamCreate(&st);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)48 << 20);
    cudaStreamAttrValue attr = {};
    attr.accessPolicyWindow.base_ptr = state;
    attr.accessPolicyWindow.num_bytes = (size_t)N * 16;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaError_t err = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    printf("persisting window on state: %s\n", cudaGetErrorString(err));
    report("tile T=16 in-order 128thr x4 + side traffic, state persisting in L2",
           time_ms([&] { k_tile_side<0><<<sms * 4, 128, 0, st>>>(out, N, 16, sched, state, actions, ep_ret, reward, term, trunc); }));
    report("tile T=16 in-order 128thr x4 + side traffic, plain (again, same stream type)",
           time_ms([&] { k_tile_side<0><<<sms * 4, 128>>>(out, N, 16, sched, state, actions, ep_ret, reward, term, trunc); }));
  }
  cudaFree(out);
  return 0;
}
