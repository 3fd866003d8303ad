// micro benchmark: latency / throughput of 2-D TMA box loads of a uint16 [R, 512] matrix (one CTA per SM)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t par) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
}
// depth = boxes in flight; each CTA streams `n` boxes of [rows x 256] starting at its own row offset
__global__ void k(const __grid_constant__ CUtensorMap m, int box_bytes, int box_rows, int n, int depth, int rows_total, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int row0 = (int)(((long long)blockIdx.x * 7919 * box_rows * n) % (rows_total - box_rows * n - 8)) & ~7;
    long long t0 = clock64();
    for (int i = 0; i < n + depth; ++i) {
      if (i >= depth) wait(&bar[(i - depth) % depth], ((i - depth) / depth) & 1);
      if (i < n) {
        const int s = i % depth;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s32(&bar[s])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(smem + s * box_bytes)), "l"(&m), "r"(128), "r"(row0 + i * box_rows), "r"(s32(&bar[s])) : "memory");
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main(int argc, char** argv) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int W = 512; const long long R = 1024LL * 512;   // 512 MiB of uint16
  uint16_t* d; cudaMalloc(&d, W * R * 2); cudaMemset(d, 1, W * R * 2);
  long long* o; cudaMalloc(&o, 4096 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int cfgs[][4] = {{4, 1, 1, 64}, {4, 4, 1, 64}, {4, 8, 1, 64}, {16, 1, 1, 16}, {16, 2, 1, 16}, {16, 4, 1, 16}, {4, 4, 148, 64}, {4, 8, 148, 64}, {16, 2, 148, 16}, {16, 4, 296, 16}, {4, 8, 296, 64}, {64, 2, 148, 8}};
  for (auto& c : cfgs) {
    const int rows = c[0], depth = c[1], grid = c[2], n = c[3];
    CUtensorMap m; cuuint64_t dims[2] = {W, (cuuint64_t)R}, strides[1] = {W * 2}; cuuint32_t box[2] = {256, (cuuint32_t)rows}, es[2] = {1, 1};
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int bytes = 256 * 2 * rows;
    k<<<grid, 32, depth * bytes>>>(m, bytes, rows, n, depth, (int)R, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid); cudaMemcpy(h.data(), o, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= grid;
    printf("box %2dx256 (%5d B) depth %d grid %3d: %8.0f clk per box, %6.2f B/clk per CTA  (%s)\n", rows, bytes, depth, grid, avg / n, bytes / (avg / n), cudaGetErrorString(e));
  }
  return 0;
}
