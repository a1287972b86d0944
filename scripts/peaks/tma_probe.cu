// tma_probe.cu — validates the tensor map the scoring kernel uses for its A operand before it goes into the engine:
// a 2-D map over a row-major FP64 matrix (leading dimension ld), box = 64 rows x 72 doubles (the padded shared-memory
// row of the wide scoring kernel, LD_K = 72: 64 columns of the tile + 8 spill columns), no swizzle, read through
// cp.async.bulk.tensor.2d with an mbarrier (SASS: UTMALDG).  Checks (a) a map passed as a __grid_constant__ parameter,
// (b) the same map read from GLOBAL memory (the engine keeps one per cluster slot in a device table), (c) zero fill of
// the spill columns beyond the matrix edge.
//   nvcc -arch=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int use_global, int col0, int row0,
                        double* out) {
  __shared__ __align__(128) double tile[64 * 72];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap* tm = use_global ? gmap : &pmap;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(64 * 72 * 8) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(tile)),
        "l"(tm), "r"(col0), "r"(row0), "r"(smem_u32(&bar))
        : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(&bar)), "r"(0)
                 : "memory");
  }
  for (int i = threadIdx.x; i < 64 * 72; i += blockDim.x) out[i] = tile[i];
}

int main() {
  const int64_t ld = 3072;
  std::vector<double> hM(ld * ld);
  for (int64_t i = 0; i < ld; ++i)
    for (int64_t j = 0; j < ld; ++j) hM[i * ld + j] = (double)(i * 10000 + j);
  double *dM, *dout;
  cudaMalloc(&dM, sizeof(double) * ld * ld);
  cudaMalloc(&dout, sizeof(double) * 64 * 72);
  cudaMemcpy(dM, hM.data(), sizeof(double) * ld * ld, cudaMemcpyHostToDevice);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)ld};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {72, 64}, estr[2] = {1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dM, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed: %d\n", (int)r);
    return 1;
  }
  CUtensorMap* gmap;
  cudaMalloc(&gmap, sizeof(CUtensorMap));
  cudaMemcpy(gmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  std::vector<double> ho(64 * 72);
  int bad = 0;
  const int cases[4][2] = {{0, 0}, {640, 1280}, {3008, 3008}, {2944, 64}};  // (col0, row0); 3008 + 72 runs past the edge
  for (int ug = 0; ug < 2; ++ug)
    for (auto& c : cases) {
      cudaMemset(dout, 0xFF, sizeof(double) * 64 * 72);
      k_probe<<<1, 128>>>(map, gmap, ug, c[0], c[1], dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(ho.data(), dout, sizeof(double) * 64 * 72, cudaMemcpyDeviceToHost);
      int nb = 0;
      for (int rr = 0; rr < 64; ++rr)
        for (int cc = 0; cc < 72; ++cc) {
          const int64_t gi = c[1] + rr, gj = c[0] + cc;
          const double want = gj < ld ? hM[gi * ld + gj] : 0.0;
          if (ho[rr * 72 + cc] != want) ++nb;
        }
      printf("map in %s memory, tile (col %d, row %d): %d wrong of %d\n", ug ? "global" : "param ", c[0], c[1], nb, 64 * 72);
      bad += nb;
    }
  printf(bad ? "TMA PROBE FAILED\n" : "TMA PROBE OK\n");
  return bad ? 1 : 0;
}
