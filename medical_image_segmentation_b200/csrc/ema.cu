// EMA of the momentum encoder as ONE multi-tensor kernel.
//
// Reference: BYOL.momentum_update (train/model/byol_pytorch.py:291-296), called after every training batch (:253-255):
//     for po, pm in zip(online.parameters(), momentum.parameters()):  pm.data.mul_(m).add_(po.data, alpha=1.0 - m)
// i.e. two elementwise torch kernels per parameter tensor (~120 launches for a ResNet-18).  Here one launch walks a
// device table of (online, momentum, length) triples; every CTA owns one 16 KB chunk of one tensor.
// Arithmetic as ATen does it: t = fl(pm * m); pm = fma(po, float(1 - m), t)  (mul_ rounds, add_(alpha) is a fused a+alpha*b).
// HBM-bound: 12 bytes per element (two reads, one write).
#include "common.cuh"

namespace mis {
namespace ema {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;      // elements per CTA

__global__ void __launch_bounds__(kThreads) ema_kernel(const MisEmaEntry* __restrict__ table, int n_tensors, float m, float om) {
  // which tensor owns this chunk: binary search over the exclusive prefix of chunk counts
  const long long c = blockIdx.x;
  int lo = 0, hi = n_tensors - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].chunk0 <= c) lo = mid;
    else hi = mid - 1;
  }
  const MisEmaEntry e = table[lo];
  const long long base = (c - e.chunk0) * kChunk;
  const long long n = e.n - base < kChunk ? e.n - base : kChunk;
  const float* __restrict__ po = static_cast<const float*>(e.online) + base;
  float* __restrict__ pm = static_cast<float*>(e.momentum) + base;
  const bool vec = ((reinterpret_cast<uintptr_t>(po) | reinterpret_cast<uintptr_t>(pm)) & 15) == 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += kThreads) {
      const float4 a = reinterpret_cast<const float4*>(po)[i];
      float4 b = reinterpret_cast<float4*>(pm)[i];
      b.x = fmaf(a.x, om, b.x * m);
      b.y = fmaf(a.y, om, b.y * m);
      b.z = fmaf(a.z, om, b.z * m);
      b.w = fmaf(a.w, om, b.w * m);
      reinterpret_cast<float4*>(pm)[i] = b;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) pm[i] = fmaf(po[i], om, pm[i] * m);
  } else {
    for (long long i = threadIdx.x; i < n; i += kThreads) pm[i] = fmaf(po[i], om, pm[i] * m);
  }
}

}  // namespace ema
}  // namespace mis

using namespace mis;

extern "C" int64_t mis_ema_chunks(int64_t n_elements) {
  return n_elements <= 0 ? 0 : (n_elements + mis::ema::kChunk - 1) / mis::ema::kChunk;
}

extern "C" int mis_ema_update(const MisEmaEntry* table_dev, int n_tensors, int64_t total_chunks, float m, float one_minus_m,
                              void* stream) {
  MIS_REQUIRE(table_dev && n_tensors > 0, MIS_ERR_INVALID_ARG, "mis_ema_update: empty table");
  MIS_REQUIRE(total_chunks > 0 && total_chunks < (1ll << 31), MIS_ERR_INVALID_ARG, "mis_ema_update: %lld chunks",
              (long long)total_chunks);
  MIS_REQUIRE(m >= 0.f && m <= 1.f, MIS_ERR_INVALID_ARG, "mis_ema_update: momentum %g outside [0, 1]", (double)m);
  const float om = one_minus_m;
  mis::ema::ema_kernel<<<dim3((unsigned)total_chunks), mis::ema::kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table_dev, n_tensors, m, om);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}
