// Arithmetic shared by the K1 variants (aug.cu: band kernels, aug_tile.cu: warp-tile kernel).
//
// Restated from torchvision 0.26 / ATen (see oracle/aug_oracle.py):
//   taps : _upsample_bilinear2d_aa -- triangle filter, support = max(scale, 1), weights normalised after summation
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace mis {
namespace aug {

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2) ---------------------------------------------
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// two packed uint16 -> two exact floats without the (slow, XU-pipe) I2F:
// 0x4B000000 | v is the float 2^23 + v; subtracting 2^23 is exact.
__device__ __forceinline__ uint64_t u16x2_to_f32x2(uint32_t p) {
  const uint32_t lo = __byte_perm(p, 0x4B000000u, 0x7610);
  const uint32_t hi = __byte_perm(p, 0x4B000000u, 0x7632);
  return fadd2(pack2(__uint_as_float(lo), __uint_as_float(hi)), pack2(-8388608.f, -8388608.f));
}

// Window of one output index.  n = input size, scale = n/m in fp32.
__device__ __forceinline__ void aa_window(int i, int n, float scale, float support, int& lo, int& hi, float& center) {
  center = (float)((double)scale * ((double)i + 0.5));
  lo = (int)((double)center - (double)support + 0.5);
  lo = lo < 0 ? 0 : lo;
  hi = (int)((double)center + (double)support + 0.5);
  hi = hi > n ? n : hi;
}
// un-normalised triangle weight of source index `idx` for an output whose window centre is `center`
__device__ __forceinline__ float aa_tri(int idx, float center, float invscale) {
  const float arg = ((float)idx - center + 0.5f) * invscale;
  return fmaxf(0.f, 1.f - fabsf(arg));
}

// ---- write RUN consecutive output pixels of one row (normalised values in v[0..RUN)), mirrored when flipped -------
template <int RUN>
__device__ __forceinline__ void store_run(const float* __restrict__ v, void* out_base, size_t row_off, int xs, int s,
                                          bool flip, bool f32) {
  const bool full = (xs + RUN <= s) && ((s & 7) == 0);
  if (f32) {
    float* out = reinterpret_cast<float*>(out_base) + row_off;
    if (full) {
      if (!flip) {
        float4* dst = reinterpret_cast<float4*>(out + xs);
#pragma unroll
        for (int i = 0; i < RUN / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
        float4* dst = reinterpret_cast<float4*>(out + (s - xs - RUN));
#pragma unroll
        for (int i = 0; i < RUN / 4; ++i)
          dst[i] = make_float4(v[RUN - 1 - 4 * i], v[RUN - 2 - 4 * i], v[RUN - 3 - 4 * i], v[RUN - 4 - 4 * i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < RUN; ++i)
        if (xs + i < s) out[flip ? (s - 1 - xs - i) : (xs + i)] = v[i];
    }
  } else {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_base) + row_off;
    if (full) {
      uint32_t pk[RUN / 2];
#pragma unroll
      for (int i = 0; i < RUN / 2; ++i) {
        __nv_bfloat162 h = flip ? __floats2bfloat162_rn(v[RUN - 1 - 2 * i], v[RUN - 2 - 2 * i])
                                : __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        pk[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      uint4* dst = reinterpret_cast<uint4*>(out + (flip ? (s - xs - RUN) : xs));
#pragma unroll
      for (int i = 0; i < RUN / 8; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < RUN; ++i)
        if (xs + i < s) out[flip ? (s - 1 - xs - i) : (xs + i)] = __float2bfloat16_rn(v[i]);
    }
  }
}

}  // namespace aug
}  // namespace mis
