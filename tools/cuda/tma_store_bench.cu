// tma_store_bench.cu -- frames streamed to HBM through the TMA unit (cp.async.bulk.global.shared::cta) instead of
// per-lane st.global.cs.v4: does the bulk-copy engine write 9408-byte frames faster than the LSU path?
// (development aid for the frame phase of env_kernels.cu; same in-order tile scheduler as env_kernel_tile, no env logic)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cuda/tsb tools/cuda/tma_store_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

constexpr int kImg = 9408, kChunks = 588;

__device__ __forceinline__ void st_cs(void* p, uint32_t a) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// baseline: in-order tiles, per-lane streaming stores (the shipped frame phase without the atlas reads)
__global__ void k_tile_lsu(uint8_t* out, int n_frames, int T, unsigned* sched) {
  __shared__ int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int n_tiles = (n_frames + T - 1) / T;
  for (;;) {
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(&sched[0], 1u);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) break;
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

// TMA: every warp owns NBUF staging buffers of one frame; build (19 x 16-byte shared stores per lane), fence, one lane
// issues the bulk store, the buffer is reused once its group has been READ.  BUILD = false skips the shared stores
// (pure engine throughput).  HINT: L2 evict_first policy on the bulk store.
template <int NBUF, bool BUILD, bool HINT>
__global__ void k_tile_tma(uint8_t* out, int n_frames, int T, unsigned* sched) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  uint8_t* my = smem + (size_t)warp * NBUF * kImg;
  const int n_tiles = (n_frames + T - 1) / T;
  const uint64_t pol = policy_evict_first();
  int buf = 0;
  for (;;) {
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(&sched[0], 1u);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    for (int i = warp; i < T; i += wpc) {
      const int f = tile * T + i;
      if (f >= n_frames) break;
      uint8_t* stage = my + buf * kImg;
      if (lane == 0) bulk_wait_read<NBUF - 1>();   // the group that last used this buffer has been read
      __syncwarp();
      if (BUILD) {
#pragma unroll
        for (int k = 0; k < 19; ++k) {
          const int c = lane + 32 * k;
          if (c < kChunks) *reinterpret_cast<uint4*>(stage + c * 16) = make_uint4(f, f, f, f);
        }
        fence_async_smem();
      }
      __syncwarp();
      if (lane == 0) {
        if (HINT) bulk_store_hint(out + (size_t)f * kImg, stage, kImg, pol);
        else bulk_store(out + (size_t)f * kImg, stage, kImg);
        bulk_commit();
      }
      buf = (buf + 1) % NBUF;
    }
    __syncthreads();
  }
  if (lane == 0) bulk_wait_read<0>();
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

template <int NBUF, bool BUILD, bool HINT>
void run_tma(const char* tag, uint8_t* out, int N, size_t bytes, int sms, unsigned* sched) {
  for (int threads : {64, 128, 256}) for (int cps : {1, 2, 3, 4, 6}) for (int T : {16, 32}) {
    const size_t smem = (size_t)(threads / 32) * NBUF * kImg;
    if (smem * cps > 220 * 1024 || smem > 227 * 1024) continue;
    cudaFuncSetAttribute(k_tile_tma<NBUF, BUILD, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const float ms = time_ms([&] { k_tile_tma<NBUF, BUILD, HINT><<<sms * cps, threads, smem>>>(out, N, T, sched); });
    cudaError_t err = cudaDeviceSynchronize();
    printf("%-28s nbuf=%d T=%2d %3d thr x %d CTAs/SM  %8.1f us  %7.1f GB/s  %s\n", tag, NBUF, T, threads, cps, ms * 1e3,
           bytes / ms / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 1 << 20;
  const size_t bytes = (size_t)N * kImg;
  uint8_t* out; cudaMalloc(&out, bytes);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* sched; cudaMalloc(&sched, 8); cudaMemset(sched, 0, 8);
  for (int T : {16, 32}) for (int cps : {3, 4}) {
    const float ms = time_ms([&] { k_tile_lsu<<<sms * cps, 128>>>(out, N, T, sched); });
    printf("LSU st.cs in-order           T=%2d 128 thr x %d CTAs/SM  %8.1f us  %7.1f GB/s\n", T, cps, ms * 1e3, bytes / ms / 1e6);
  }
  run_tma<1, false, false>("TMA no build", out, N, bytes, sms, sched);
  run_tma<2, false, false>("TMA no build", out, N, bytes, sms, sched);
  run_tma<1, true, false>("TMA build", out, N, bytes, sms, sched);
  run_tma<2, true, false>("TMA build", out, N, bytes, sms, sched);
  run_tma<2, true, true>("TMA build evict_first", out, N, bytes, sms, sched);
  // spot check: the last frame holds its own number
  uint32_t v = 0;
  cudaMemcpy(&v, out + (size_t)(N - 1) * kImg + 64, 4, cudaMemcpyDeviceToHost);
  printf("check: last frame word = %u (expect %d)\n", v, N - 1);
  return 0;
}
