// small_fill_bench.cu -- the floor for a SMALL batch: N 9408-byte frames written by one warp per frame with the step
// kernel's own store instruction (st.global.cs.v4) and grid shape, constant data, no env logic, 64 launches replayed
// from one CUDA graph (as tools/sweep.py replays the step kernel).  The gap between this and env_kernel_warp is what
// the dependent chain (state -> window -> visibility -> kinds) and the frame assembly cost at that batch size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cuda/sfb tools/cuda/small_fill_bench.cu
//   tools/cuda/sfb 4096 8192 16384 24576
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int kImg = 9408;

__device__ __forceinline__ void st_cs(void* p, uint32_t a) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(a) : "memory");
}
// 19 rounds of 32 lanes (the chunk-map order of the tile kernel)
__global__ void __launch_bounds__(256) k_rounds19(uint8_t* out, int n) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n; f += warps) {
    uint8_t* frame = out + (size_t)f * kImg;
#pragma unroll
    for (int k = 0; k < 19; ++k) { const int c = lane + 32 * k; if (c < 588) st_cs(frame + c * 16, f); }
  }
}
// 28 row pairs of 21 lanes (the map-free order of env_kernel_warp)
__global__ void __launch_bounds__(256) k_pairs28(uint8_t* out, int n) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n; f += warps) {
    uint8_t* frame = out + (size_t)f * kImg + lane * 16;
    if (lane < 21) {
#pragma unroll
      for (int k = 0; k < 28; ++k) st_cs(frame + k * 336, f);
    }
  }
}

template <typename F> static float graph_us(F launch, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t ge;
  launch(); cudaStreamSynchronize(s);
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < 64; ++i) launch();
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, s);
  for (int i = 0; i < 8; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(b, s); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1e3f / 512.f;
}

int main(int argc, char** argv) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaStream_t s; cudaStreamCreate(&s);
  for (int i = 1; i < argc; ++i) {
    const int n = atoi(argv[i]);
    uint8_t* out; cudaMalloc(&out, (size_t)n * kImg);
    const int grid = (n + 7) / 8 < sms * 4 ? (n + 7) / 8 : sms * 4;
    const float a = graph_us([&] { k_rounds19<<<grid, 256, 0, s>>>(out, n); }, s);
    const float b = graph_us([&] { k_pairs28<<<grid, 256, 0, s>>>(out, n); }, s);
    const float e = graph_us([&] { k_rounds19<<<1, 32, 0, s>>>(out, 1); }, s);
    printf("N=%6d  %7.1f MB  19x32 lanes %6.2f us (%5.0f GB/s)   28x21 lanes %6.2f us (%5.0f GB/s)   one-warp launch %5.2f us\n", n,
           n * (double)kImg / 1e6, a, n * (double)kImg / a / 1e3, b, n * (double)kImg / b / 1e3, e);
    cudaFree(out);
  }
  return 0;
}
