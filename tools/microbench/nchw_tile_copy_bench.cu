// Micro-benchmark: copy a [C][HW] fp32 tensor (C = 256 planes of 721x1440) the way a pixel-tile kernel touches NCHW:
// a CTA takes a tile of TILE consecutive pixels and walks all C channel rows (TILE*4 contiguous bytes per row, rows
// 4 MB apart).  Reports GB/s (read + write) per tile width -- how much of the HBM rate survives the access pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nchw_tile_copy_bench nchw_tile_copy_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int TILE, bool WRITE>
__global__ void k(const float* __restrict__ x, float* __restrict__ y, long long HW, int C, float* sink) {
  constexpr int V = TILE / 4;              // float4 per channel row of the tile
  const int ntiles = (int)(HW / TILE);
  float acc = 0.f;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long p0 = (long long)t * TILE;
    for (int i = threadIdx.x; i < C * V; i += blockDim.x) {
      const int c = i / V, v = i - c * V;
      const float4 a = __ldcs(reinterpret_cast<const float4*>(x + (long long)c * HW + p0) + v);
      if (WRITE) __stcs(reinterpret_cast<float4*>(y + (long long)c * HW + p0) + v, a);
      else acc += a.x + a.y + a.z + a.w;
    }
  }
  if (!WRITE && acc == 1.2345f) *sink = acc;
}

template <int TILE, bool WRITE>
static void run(const float* x, float* y, long long HW, int C, float* sink, int grid) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<TILE, WRITE><<<grid, 512>>>(x, y, HW, C, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) k<TILE, WRITE><<<grid, 512>>>(x, y, HW, C, sink);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  const double bytes = (double)C * (HW / TILE * TILE) * 4 * (WRITE ? 2 : 1);
  printf(" {\"tile_px\": %d, \"mode\": \"%s\", \"grid\": %d, \"ms\": %.4f, \"GBps\": %.0f},\n", TILE, WRITE ? "copy" : "read", grid, ms, bytes / ms / 1e6);
}

int main() {
  const long long HW = 721LL * 1440; const int C = 256;
  float *x, *y, *sink; cudaMalloc(&x, C * HW * 4); cudaMalloc(&y, C * HW * 4); cudaMalloc(&sink, 4);
  cudaMemset(x, 0, C * HW * 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("[\n");
  for (int g : {sms, 2 * sms, 4 * sms}) {
    run<32, true>(x, y, HW, C, sink, g); run<128, true>(x, y, HW, C, sink, g); run<256, true>(x, y, HW, C, sink, g);
    run<512, true>(x, y, HW, C, sink, g); run<1440, true>(x, y, HW, C, sink, g);
    run<128, false>(x, y, HW, C, sink, g); run<256, false>(x, y, HW, C, sink, g); run<1440, false>(x, y, HW, C, sink, g);
  }
  printf(" {}\n]\n");
  return 0;
}
