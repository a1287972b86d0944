// FP64 peak probes for the roofline denominators of the FP64-bound kernels (MEASURED_PEAKS.json has no FP64 entry):
// register-resident DMMA.8x8x4 and DFMA loops, no memory traffic.  nvcc -arch=sm_100a -O3 fp64_peak.cu -o fp64_peak
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double acc[CH][2];
  for (int c = 0; c < CH; ++c) acc[c][0] = acc[c][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) dmma(acc[c][0], acc[c][1], a, b);
  }
  double s = 0.0;
  for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double acc[CH];
  for (int c = 0; c < CH; ++c) acc[c] = c;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0.0;
  for (int c = 0; c < CH; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Are the tensor (DMMA) and vector (DFMA) FP64 paths the same units?  Even warps run the DMMA loop, odd warps the DFMA
// loop, in the same CTA: if the two were independent pipes the combined rate would exceed either peak.
template <int CH>
__global__ void k_mix(double* out, int iters_mma, int iters_fma, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0, s = 0.0;
  if ((threadIdx.x >> 5) & 1) {
    double acc[CH];
    for (int c = 0; c < CH; ++c) acc[c] = c;
    for (int i = 0; i < iters_fma; ++i) {
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
    }
    for (int c = 0; c < CH; ++c) s += acc[c];
  } else {
    double acc[CH][2];
    for (int c = 0; c < CH; ++c) acc[c][0] = acc[c][1] = 0.0;
    for (int i = 0; i < iters_mma; ++i) {
#pragma unroll
      for (int c = 0; c < CH; ++c) dmma(acc[c][0], acc[c][1], a, b);
    }
    for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lat(double* out, int iters, double a, double b) {  // dependent DFMA chain: latency per op
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / iters;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_dmma<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = (double)sms * warps * iters * 8 * 512.0;
      if (rep) printf("DMMA.8x8x4  %2d warps/SM, 8 chains: %.2f TFLOP/s  (%.2f clk per DMMA per SM at 1.965 GHz)\n", warps, flops / ms / 1e9, ms * 1e-3 * 1.965e9 / (warps * iters * 8.0));
    }
  }
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_dfma<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 0.9999999);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = (double)sms * warps * 32 * iters * 8 * 2.0;
      if (rep) printf("DFMA        %2d warps/SM, 8 chains: %.2f TFLOP/s\n", warps, flops / ms / 1e9);
    }
  }
  for (int warps = 8; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      // a DMMA warp instruction is 512 flops, a DFMA warp instruction 64: 8x the DFMA iterations = equal flops per warp
      cudaEventRecord(e0);
      k_mix<8><<<sms, warps * 32>>>(out, iters, iters * 8, 1.0000001, 0.9999999);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = (double)sms * (warps / 2) * iters * 8 * 512.0 * 2.0;
      if (rep) printf("DMMA + DFMA %2d warps/SM (half each, equal flops): %.2f TFLOP/s combined\n", warps, flops / ms / 1e9);
    }
  }
  k_lat<<<1, 32>>>(out, 100000, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  double lat; cudaMemcpy(&lat, out + 64, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.1f cycles\n", lat);
  return 0;
}
