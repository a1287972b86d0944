// How long does a tiny kernel on a HIGH-priority stream wait while a low-priority grid that owns every SM (one fat CTA
// per SM: 512 threads x 128 registers, 200 KB shared memory) is running?  Mimics the look-ahead pass vs the rank-n
// score corrections.  nvcc -arch=sm_100a -O3 prio_probe.cu -o prio_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <algorithm>
#include <chrono>
#include <vector>
__global__ void __launch_bounds__(512, 1) fat(long long cycles_per_tile, const int* tiles, double* sink) {
  extern __shared__ double sm[];
  const int n = tiles[blockIdx.x];
  double acc = threadIdx.x;
  for (int t = 0; t < n; ++t) {
    long long t0 = clock64();
    while (clock64() - t0 < cycles_per_tile) acc = acc * 1.0000001 + sm[threadIdx.x];
  }
  if (acc == 12345.678) sink[0] = acc;
}
__global__ void tiny(double* out) { out[threadIdx.x] = threadIdx.x; }
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStream_t s_lo, s_hi; cudaStreamCreateWithPriority(&s_lo, cudaStreamNonBlocking, lo); cudaStreamCreateWithPriority(&s_hi, cudaStreamNonBlocking, hi);
  cudaFuncSetAttribute(fat, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  double* d; cudaMalloc(&d, 1 << 20);
  const int nb = 41, slots = 30, items = nb * slots;
  for (int order = 0; order < 2; ++order) {
    std::vector<int> tiles(items);
    for (int z = 0; z < nb; ++z)
      for (int s = 0; s < slots; ++s) tiles[z * slots + s] = order == 0 ? nb - z : ((z & 1) ? (z >> 1) + 1 : nb - (z >> 1));
    int* dt; cudaMalloc(&dt, items * sizeof(int)); cudaMemcpy(dt, tiles.data(), items * sizeof(int), cudaMemcpyHostToDevice);
    for (int load = 0; load < 2; ++load) {
      std::vector<double> lat;
      for (int rep = 0; rep < 40; ++rep) {
        if (load) fat<<<items, 512, 200 * 1024, s_lo>>>(2400, dt, d);  // ~1.2 us per tile: the whole grid ~ 300 us
        for (int k = 0; k < 6; ++k) {
          auto t0 = std::chrono::steady_clock::now();
          while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 30e-6) {}
          t0 = std::chrono::steady_clock::now();
          tiny<<<1, 32, 0, s_hi>>>(d + 1024); tiny<<<1, 32, 0, s_hi>>>(d + 2048); tiny<<<1, 32, 0, s_hi>>>(d + 4096);
          cudaStreamSynchronize(s_hi);
          lat.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6);
        }
        cudaDeviceSynchronize();
      }
      std::sort(lat.begin(), lat.end());
      printf("order %s, %s: 3 dependent tiny kernels + sync: median %.1f us, p90 %.1f us, max %.1f us\n", order ? "zigzag" : "longest-first",
             load ? "fat grid running (low priority)" : "idle GPU", lat[lat.size() / 2], lat[lat.size() * 9 / 10], lat.back());
    }
  }
  return 0;
}
