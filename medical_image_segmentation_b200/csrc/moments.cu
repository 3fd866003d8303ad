// N3 -- exact per-channel sum / sum-of-squares of uint16 slices (the statistics behind the normalisation constants).
//
// Replaces the float64 streaming reduction of compute_mean_and_std
// (medical_image_segmentation/analyze_data/compute_dataset_metrics.py:12-29): sum_ = sum(images),
// sum_squared = sum(images**2) over (batch, H, W) per channel.  For uint16 pixels both sums are exact integers
// (sum x^2 <= 2^32 * n fits uint64 for n <= 2^32 pixels per channel), so the device accumulates in uint64 and the
// result is independent of the reduction order; mean / population-std are formed on the host in float64 exactly as
// the reference does (mean_of_squares - mean**2).
//
// Pure HBM-bound scan: 16-byte streaming loads (ld.global.nc.L1::no_allocate), 4 loads in flight per thread,
// grid = a multiple of the SM count, one pair of 64-bit atomics per CTA.
#include "common.cuh"

namespace mis {
namespace mom {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 8;                    // 16-byte vectors per thread per chunk
constexpr int kChunkElems = kThreads * kVecPerThread * 8;

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void acc_word(uint32_t w, uint32_t& s1, unsigned long long& s2) {
  const uint32_t lo = w & 0xffffu, hi = w >> 16;
  s1 += lo + hi;
  s2 += (unsigned long long)(lo * lo) + (unsigned long long)(hi * hi);   // each square < 2^32: exact in 32 bits
}

// planes = n_images * C planes of plane_elems pixels; CTA (chunk, plane) scans one chunk of one plane
__global__ void __launch_bounds__(kThreads) moments_kernel(const uint16_t* __restrict__ src, long long plane_elems, int C,
                                                           unsigned long long* __restrict__ sums) {
  const long long plane = blockIdx.y;
  const uint16_t* base = src + plane * plane_elems;
  const long long e0 = (long long)blockIdx.x * kChunkElems;
  const long long e1 = min(e0 + (long long)kChunkElems, plane_elems);
  unsigned long long s1 = 0, s2 = 0;
  // vector body (the plane base is 16-byte aligned when plane_elems % 8 == 0 and src is aligned; checked on the host)
  const long long nvec = (e1 - e0) >> 3;
  const uint4* vp = reinterpret_cast<const uint4*>(base + e0);
  for (long long v = threadIdx.x; v < nvec; v += (long long)kThreads * 4) {
    uint4 r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      r[k] = (v + (long long)k * kThreads < nvec) ? ld_stream(vp + v + (long long)k * kThreads) : make_uint4(0, 0, 0, 0);
    uint32_t t1 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc_word(r[k].x, t1, s2);
      acc_word(r[k].y, t1, s2);
      acc_word(r[k].z, t1, s2);
      acc_word(r[k].w, t1, s2);
    }
    s1 += t1;
  }
  for (long long e = e0 + (nvec << 3) + threadIdx.x; e < e1; e += kThreads) {   // tail
    const uint32_t x = base[e];
    s1 += x;
    s2 += (unsigned long long)(x * x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  __shared__ unsigned long long red[2][kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0, b = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    const int ch = (int)(plane % C);
    atomicAdd(&sums[2 * ch], a);
    atomicAdd(&sums[2 * ch + 1], b);
  }
}

}  // namespace mom
}  // namespace mis

using namespace mis;

extern "C" int mis_u16_moments(const uint16_t* src, long long n_images, int C, long long plane_elems,
                               unsigned long long* sums, void* stream) {
  MIS_REQUIRE(src && sums, MIS_ERR_INVALID_ARG, "mis_u16_moments: null pointer");
  MIS_REQUIRE(n_images > 0 && C > 0 && C <= 4 && plane_elems > 0, MIS_ERR_INVALID_ARG, "mis_u16_moments: bad sizes");
  MIS_REQUIRE(n_images * plane_elems <= (1ll << 32), MIS_ERR_UNSUPPORTED,
              "mis_u16_moments: more than 2^32 pixels per channel would overflow the exact uint64 sum of squares; "
              "accumulate over several calls");
  MIS_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (plane_elems & 7) == 0, MIS_ERR_UNSUPPORTED,
              "mis_u16_moments: src must be 16-byte aligned and plane_elems a multiple of 8");
  MIS_REQUIRE(n_images * C <= 65535, MIS_ERR_UNSUPPORTED, "mis_u16_moments: at most 65535 planes per call");
  const unsigned chunks = (unsigned)((plane_elems + mom::kChunkElems - 1) / mom::kChunkElems);
  mom::moments_kernel<<<dim3(chunks, (unsigned)(n_images * C)), mom::kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, plane_elems, C, sums);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}
