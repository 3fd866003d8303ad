// micro test: 2-D TMA box loads of a uint16 matrix without swizzle (which descriptor settings work on sm_100a?)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, int c0, int c1, int bytes,
                  uint16_t* out, int sel) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    if (MODE == 0) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(s32(smem)), "l"(&m0), "r"(c0), "r"(c1), "r"(s32(&bar)) : "memory");
    } else {
      if (sel == 0)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(smem)), "l"(&m0), "r"(c0), "r"(c1), "r"(s32(&bar)) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(smem)), "l"(&m1), "r"(c0), "r"(c1), "r"(s32(&bar)) : "memory");
    }
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1; int idx = 0;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int W = 512, R = 1024;
  std::vector<uint16_t> h(W * R);
  for (int i = 0; i < W * R; ++i) h[i] = (uint16_t)(i * 7 + (i >> 9));
  uint16_t *d, *o; cudaMalloc(&d, W * R * 2); cudaMalloc(&o, 65536); cudaMemcpy(d, h.data(), W * R * 2, cudaMemcpyHostToDevice);
  struct Cfg { CUtensorMapDataType dt; int bw, br; CUtensorMapL2promotion l2; const char* name; };
  const int c0_override = argc > 2 ? atoi(argv[2]) : 37;
  Cfg cfgs[] = {{CU_TENSOR_MAP_DATA_TYPE_UINT16, 64, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "u16 64x8 l2_128"},
                {CU_TENSOR_MAP_DATA_TYPE_UINT16, 256, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "u16 256x8 l2_128"},
                {CU_TENSOR_MAP_DATA_TYPE_UINT16, 256, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE, "u16 256x8 l2_none"},
                {CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 256, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "bf16 256x8 l2_256"},
                {CU_TENSOR_MAP_DATA_TYPE_UINT16, 192, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "u16 192x8 l2_128"}};
  for (int mode = 0; mode < 2; ++mode)
    for (auto& c : cfgs) {
      if (only >= 0 && idx++ != only) continue;
      CUtensorMap m0, m1;
      cuuint64_t dims[2] = {W, R}, strides[1] = {W * 2};
      cuuint32_t box[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.br}, es[2] = {1, 1};
      CUresult r0 = enc(&m0, c.dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      m1 = m0;
      const int bytes = c.bw * c.br * 2, c0 = c0_override, c1 = 101;
      cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      if (mode == 0) k<0><<<1, 128, 65536>>>(m0, m1, c0, c1, bytes, o, 1); else k<1><<<1, 128, 65536>>>(m0, m1, c0, c1, bytes, o, 1);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<uint16_t> got(bytes / 2); cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < c.br; ++r) for (int x = 0; x < c.bw; ++x) {
        uint16_t want = (c0 + x < W) ? h[(c1 + r) * W + c0 + x] : 0;
        bad += got[r * c.bw + x] != want;
      }
      printf("mode %d %-20s encode %d run %s mismatches %d\n", mode, c.name, (int)r0, cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
