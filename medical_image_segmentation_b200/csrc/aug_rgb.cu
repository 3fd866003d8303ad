// K1c -- colour chain of the two-view transform for 3-channel input (sm_100a).
//
// The colour ops of ColorJitter / RandomGrayscale mix the three channels of a pixel, so for C == 3 the strip kernel
// (aug_strip.cu, raw_all) only crops, resamples and flips and leaves every plane as uint16 (round(x * 65535)) in the
// first 2*s*s bytes of its output plane; this kernel -- one CTA per view -- stages the three planes in shared memory and
// applies, in the view's random op order (torchvision v2/_color.py:146-171):
//   brightness  x * b, clamp                                            functional/_color.py:114-125
//   contrast    blend(x, mean(gray(x)), c), mean over the whole view    :190-205, _blend :92-97
//   saturation  blend(x, gray(x), s)                                    :151-166
//   hue         rgb -> hsv, h = (h + hue) mod 1, hsv -> rgb             :300-396
//   gray(x) = 0.2989 r + 0.587 g + 0.114 b                              :31-48
// then RandomGrayscale (gray replicated), RandomSolarize, Normalize -- or, for a view that drew a GaussianBlur, the
// post-colour image as uint16 again for mis_aug_blur_views.  The contrast mean needs the ops in front of it applied
// to every pixel: those are evaluated twice (once for the mean, once for the result).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mis {
namespace augc {

constexpr int kThreads = 512;

struct RgbArgs {
  void* out;
  const MisViewParams* params;
  int s, out_f32;
  int stream;                // 1: planes too large for shared memory (s > 192), read and finished in place
  float mean[4], inv_std[4];
};

struct Px {
  float r, g, b;
};
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float gray_of(const Px& p) { return fmaf(p.b, 0.114f, fmaf(p.g, 0.587f, p.r * 0.2989f)); }

__device__ __forceinline__ Px adjust_hue(const Px& p, float hue) {
  // _rgb_to_hsv (functional/_color.py:300-343)
  const float maxc = fmaxf(p.r, fmaxf(p.g, p.b)), minc = fminf(p.r, fminf(p.g, p.b));
  const bool eqc = maxc == minc;
  const float range = maxc - minc;
  const float sat = range / (eqc ? 1.f : maxc);
  const float div = eqc ? 1.f : range;
  const float rc = (maxc - p.r) / div, gc = (maxc - p.g) / div, bc = (maxc - p.b) / div;
  const bool max_neq_r = maxc != p.r, max_eq_g = maxc == p.g;
  const float hg = (max_eq_g && max_neq_r) ? (rc + 2.f) - bc : 0.f;
  const float hr = (!max_neq_r) ? bc - gc : 0.f;
  const float hb = (max_neq_r && !max_eq_g) ? (gc + 4.f) - rc : 0.f;
  float h = fmodf((hr + hg + hb) * (1.f / 6.f) + 1.f, 1.f);
  // h.add_(hue).remainder_(1.0): python-style remainder (result in [0, 1))
  h = h + hue;
  h = h - floorf(h);
  // _hsv_to_rgb (:346-372)
  const float h6 = h * 6.f;
  const float fi = floorf(h6);
  const float f = h6 - fi;
  int i = (int)fi % 6;
  if (i < 0) i += 6;
  const float v = maxc;
  const float sxf = sat * f, oms = 1.f - sat;
  const float q = clamp01((1.f - sxf) * v), t = clamp01((sxf + oms) * v), pp = clamp01(oms * v);
  Px o;
  switch (i) {
    case 0: o = {v, t, pp}; break;
    case 1: o = {q, v, pp}; break;
    case 2: o = {pp, v, t}; break;
    case 3: o = {pp, q, v}; break;
    case 4: o = {t, pp, v}; break;
    default: o = {v, pp, q}; break;
  }
  return o;
}

// ops order[k0 .. k1) of the jitter chain on one pixel; `cadd` = mean * (1 - c) of the contrast op if it is in range
__device__ __forceinline__ Px apply_ops(Px p, const MisViewParams& P, int k0, int k1, float cadd) {
  for (int k = k0; k < k1; ++k) {
    switch (P.order[k]) {
      case 0:
        p = {clamp01(p.r * P.brightness), clamp01(p.g * P.brightness), clamp01(p.b * P.brightness)};
        break;
      case 1:
        p = {clamp01(fmaf(p.r, P.contrast, cadd)), clamp01(fmaf(p.g, P.contrast, cadd)), clamp01(fmaf(p.b, P.contrast, cadd))};
        break;
      case 2: {
        const float ga = gray_of(p) * (1.f - P.saturation);
        p = {clamp01(fmaf(p.r, P.saturation, ga)), clamp01(fmaf(p.g, P.saturation, ga)), clamp01(fmaf(p.b, P.saturation, ga))};
        break;
      }
      default:
        p = adjust_hue(p, P.hue);
        break;
    }
  }
  return p;
}

__global__ void __launch_bounds__(kThreads, 1) rgb_color_kernel(const RgbArgs a) {
  extern __shared__ __align__(16) uint16_t sm16[];
  __shared__ float red[kThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int view = blockIdx.x;
  const int s = a.s, n = s * s;
  const MisViewParams P = a.params[view];
  const size_t esz = a.out_f32 ? 4 : 2;
  uint8_t* const base = static_cast<uint8_t*>(a.out) + (size_t)view * 3 * n * esz;

  // ---- the three uint16 planes: staged in shared memory (the output overwrites them in global memory), or -- crops
  //      above 192, whose planes exceed one SM's shared memory -- read in place (a.stream) --------------------------------
  const uint16_t* pl[3];
  if (!a.stream) {
    for (int c = 0; c < 3; ++c) {
      const uint4* src = reinterpret_cast<const uint4*>(base + (size_t)c * n * esz);
      uint4* dst = reinterpret_cast<uint4*>(sm16 + (size_t)c * n);
      for (int i = tid; i < n / 8; i += kThreads) dst[i] = src[i];
      pl[c] = sm16 + (size_t)c * n;
    }
    __syncthreads();
  } else {
    for (int c = 0; c < 3; ++c) pl[c] = reinterpret_cast<const uint16_t*>(base + (size_t)c * n * esz);
  }
  auto load_px = [&](int i) {
    Px p;
    p.r = (float)pl[0][i] * (1.f / 65535.f);
    p.g = (float)pl[1][i] * (1.f / 65535.f);
    p.b = (float)pl[2][i] * (1.f / 65535.f);
    return p;
  };

  const bool jitter = (P.flags & MIS_VIEW_JITTER) != 0;
  int pos_c = 4;
  float cadd = 0.f;
  if (jitter) {
    for (int k = 0; k < 4; ++k)
      if (P.order[k] == 1) pos_c = k;
    // contrast mean: grayscale mean of the image as it stands when the contrast op runs
    float acc = 0.f;
    for (int i = tid; i < n; i += kThreads) acc += gray_of(apply_ops(load_px(i), P, 0, pos_c, 0.f));
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
    cadd = tot / (float)n * (1.f - P.contrast);
  }
  const bool gray = (P.flags & MIS_VIEW_GRAY) != 0, sol = (P.flags & MIS_VIEW_SOLARIZE) != 0;
  const bool raw = (P.flags & MIS_VIEW_BLUR) != 0;
  auto finish = [&](int i, Px p) {
    if (jitter) p = apply_ops(p, P, 0, 4, cadd);
    if (gray) {
      const float gsc = gray_of(p);
      p = {gsc, gsc, gsc};
    }
    float v[3] = {p.r, p.g, p.b};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint8_t* plane = base + (size_t)c * n * esz;
      if (raw) {        // post-colour image for mis_aug_blur_views (blur -> solarize -> normalise)
        reinterpret_cast<uint16_t*>(plane)[i] = (uint16_t)__float2uint_rn(v[c] * 65535.f);
      } else {
        float x = v[c];
        if (sol) x = x >= MIS_SOLARIZE_THRESHOLD ? 1.f - x : x;
        x = (x - a.mean[c]) * a.inv_std[c];
        if (a.out_f32) reinterpret_cast<float*>(plane)[i] = x;
        else reinterpret_cast<__nv_bfloat16*>(plane)[i] = __float2bfloat16_rn(x);
      }
    }
  };
  if (!a.stream || !a.out_f32 || raw) {
    // staged planes, or a result as wide as the uint16 it replaces: pixel i only overwrites its own three inputs
    for (int i = tid; i < n; i += kThreads) finish(i, load_px(i));
  } else {
    // fp32 result in place: pixel i's result covers the uint16 inputs 2i and 2i+1 of its plane.  Chunks of kThreads
    // pixels from the top down, every pixel of a chunk loaded before any is stored: a chunk [a, b) then only overwrites
    // inputs in [a, 2b), its own (already in registers) and those of chunks finished before.
    for (int c0 = ((n - 1) / kThreads) * kThreads; c0 >= 0; c0 -= kThreads) {
      const int i = c0 + tid;
      Px p = {0.f, 0.f, 0.f};
      if (i < n) p = load_px(i);
      __syncthreads();
      if (i < n) finish(i, p);
    }
  }
}

bool rgb_supported(int s) { return (s & 7) == 0 && s <= 256; }

int launch_rgb_color(void* out, int out_f32, const MisViewParams* params, int n_views, int s, const float* mean,
                     const float* inv_std, cudaStream_t stream) {
  RgbArgs a = {};
  a.out = out;
  a.params = params;
  a.s = s;
  a.out_f32 = out_f32;
  for (int c = 0; c < 3; ++c) {
    a.mean[c] = mean[c];
    a.inv_std[c] = inv_std[c];
  }
  a.stream = (size_t)6 * s * s > 227 * 1024 - 256;
  const size_t smem = a.stream ? 0 : (size_t)6 * s * s;
  MIS_CUDA_TRY(cudaFuncSetAttribute(rgb_color_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rgb_color_kernel<<<dim3((unsigned)n_views), kThreads, smem, stream>>>(a);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}

}  // namespace augc
}  // namespace mis
