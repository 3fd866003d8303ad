// K1b -- GaussianBlur(23) + RandomSolarize + Normalize for the views of the two-view chain that drew a blur (sm_100a).
//
// Reference: RandomApply([GaussianBlur(kernel_size=23)], p) then RandomSolarize(128, p) then ToDtype / Normalize
// (train/data_loaders/lightning_module.py:53-57); torchvision gaussian_blur_image (v2/functional/_misc.py:104-165):
// kernel1d = softmax(-(linspace(-lim, lim, 23) / sigma)^2), lim = 22 / (2 sqrt 2); reflect padding by 11; conv2d with
// the outer product of the two 1-D kernels.  The blur acts on the image AFTER the colour jitter, whose clamps do not
// commute with it, so K1 (aug_strip.cu) leaves the post-colour image of such a view as uint16 (round(x * 65535)) in
// the first 2*s*s bytes of the view's own output plane and this kernel finishes it in place:
//
//   one CTA per plane: uint16 plane -> fp32 in shared memory (rows padded by the reflected 11 columns) ->
//   horizontal 23-tap pass in place (one warp per row, 8 outputs per lane from 8 16-byte loads) ->
//   vertical 23-tap pass (8 rows x 4 columns per thread, every loaded row feeds all the outputs it touches) ->
//   solarize (x >= 128/255 -> 1 - x, functional/_color.py:497-501) -> (x - mean) / std -> bf16 / fp32 store.
//
// The separable order differs from torchvision's single 2-D convolution only by fp32 rounding (~1e-7).
// Views without MIS_VIEW_BLUR exit at once.  s must be a multiple of 8 with 16 <= s <= 256; crops above 224 (whose fp32
// plane exceeds one SM's 227 KB) are processed as two bands of 128 output rows.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mis {
namespace augb {

constexpr int kTaps = 23;
constexpr int kR = 11;
constexpr int kPad = 12;            // left padding of a shared-memory row (>= kR, multiple of 4 floats)
constexpr int kThreads = 512;

struct BlurArgs {
  void* out;
  const MisViewParams* params;
  int C, s, out_f32;
  int pitch;                        // floats per shared-memory row: s + 2 * kPad
  int nbands, band_h, band_rows;    // 1 band of s rows (s <= 224) or 2 bands of 128 output rows (+ 11 halo rows each side)
  float mean[4], inv_std[4];
};

__global__ void __launch_bounds__(kThreads, 1) blur_kernel(const BlurArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane = blockIdx.x;
  const int view = a.C == 1 ? plane : plane / a.C;
  const int chan = plane - view * a.C;
  const MisViewParams P = a.params[view];
  if (!(P.flags & MIS_VIEW_BLUR)) return;
  const int s = a.s, pitch = a.pitch;
  const size_t esz = a.out_f32 ? 4 : 2;
  uint8_t* const plane_ptr = static_cast<uint8_t*>(a.out) + (size_t)plane * s * s * esz;
  // shared memory: [band_rows][pitch] working rows | [kR][pitch] carry rows (two bands only) | 32 floats of weights
  float* const carry = sm + (size_t)a.band_rows * pitch;
  float* const wsm = carry + (a.nbands > 1 ? (size_t)kR * pitch : 0);

  // ---- 1-D kernel (v2/functional/_misc.py:86-90): fp32 linspace as ATen builds it, softmax over the 23 taps --------
  if (warp == 0) {
    const float lim = (float)((kTaps - 1) / (2.0 * 1.4142135623730951));
    const float step = (lim - (-lim)) / (float)(kTaps - 1);
    float e = 0.f;
    if (lane < kTaps) {
      const float x = lane < kTaps / 2 ? -lim + step * (float)lane : lim - step * (float)(kTaps - 1 - lane);
      const float q = x / P.blur_sigma;
      e = expf(-(q * q));             // the maximum of -(x/sigma)^2 is 0 (centre tap)
    }
    const float tot = warp_sum(e);
    if (lane < kTaps) wsm[lane] = e / tot;
  }

  const int chunks_per_row = s >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(plane_ptr);
  // u16 rows [r0, r1) of the plane -> fp32 rows of `dst` (row r lands at dst row r - r0), 8 pixels per thread and step
  auto load_rows = [&](float* dst, int r0, int r1) {
    for (int c = tid; c < (r1 - r0) * chunks_per_row; c += kThreads) {
      const int y = c / chunks_per_row, x8 = (c - y * chunks_per_row) * 8;
      const uint4 q = src[(size_t)(r0 + y) * chunks_per_row + (x8 >> 3)];
      const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] = (float)(qq[i] & 0xffffu) * (1.f / 65535.f);
        v[2 * i + 1] = (float)(qq[i] >> 16) * (1.f / 65535.f);
      }
      float4* d4 = reinterpret_cast<float4*>(dst + (size_t)y * pitch + kPad + x8);
      d4[0] = make_float4(v[0], v[1], v[2], v[3]);
      d4[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  };

  // Two bands (s > 224: the fp32 plane does not fit one SM): the output overwrites the uint16 input in place, so the rows
  // of the lower band that the upper band's windows reach into are saved first, the LOWER band is finished first (its
  // bf16 rows overwrite exactly its own input rows; its fp32 rows lie beyond the uint16 image), then the upper band.
  const int bh = a.nbands > 1 ? a.band_h : s;
  if (a.nbands > 1) load_rows(carry, bh, bh + kR);
  __syncthreads();
  float w[kTaps];
#pragma unroll
  for (int k = 0; k < kTaps; ++k) w[k] = wsm[k];
  const bool sol = (P.flags & MIS_VIEW_SOLARIZE) != 0;
  const float mean = a.mean[chan], inv_std = a.inv_std[chan];

  for (int band = a.nbands - 1; band >= 0; --band) {
    const int y0 = band * bh, y1 = min(s, y0 + bh);            // output rows of this band
    const int lo = max(0, y0 - kR), hi = min(s, y1 + kR);      // source rows it reads (reflection stays inside)
    const int nr = hi - lo;
    if (a.nbands > 1 && band == 0) {
      load_rows(sm, lo, y1);
      for (int i = tid; i < (hi - y1) * (s >> 2); i += kThreads) {   // rows [y1, hi) were overwritten: take the saved copy
        const int y = i / (s >> 2), x4 = (i - y * (s >> 2)) * 4;
        *reinterpret_cast<float4*>(sm + (size_t)(y1 - lo + y) * pitch + kPad + x4) =
            *reinterpret_cast<const float4*>(carry + (size_t)y * pitch + kPad + x4);
      }
    } else {
      load_rows(sm, lo, hi);
    }
    __syncthreads();
    // reflected columns: -j <-> j and s-1+j <-> s-1-j (torch_pad mode="reflect")
    for (int i = tid; i < nr * 2 * kPad; i += kThreads) {
      const int y = i / (2 * kPad), j = i - y * (2 * kPad);
      float* row = sm + (size_t)y * pitch + kPad;
      if (j < kPad) row[-1 - j] = row[min(1 + j, s - 1)];
      else row[s + (j - kPad)] = row[max(s - 2 - (j - kPad), 0)];
    }
    __syncthreads();

    // ---- horizontal pass, in place: one warp per row, lane = 8 consecutive outputs ---------------------------------
    for (int y = warp; y < nr; y += kThreads / 32) {
      float* row = sm + (size_t)y * pitch;
      float v[32];
      const bool act = 8 * lane < s;
      if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = *reinterpret_cast<const float4*>(row + 8 * lane + 4 * i);
          v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
      }
      __syncwarp();                               // every window of the row is in registers before it is overwritten
      if (act) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < kTaps; ++k) acc = fmaf(w[k], v[j + k + 1], acc);   // padded column 8*lane + j + k + 1
          o[j] = acc;
        }
        float4* dst = reinterpret_cast<float4*>(row + kPad + 8 * lane);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    __syncthreads();

    // ---- vertical pass + solarize + normalise + store: 8 rows x 4 columns per thread -------------------------------
    const int cg = s >> 2;                        // column groups of 4
    const int rb = (y1 - y0 + 7) >> 3;            // row blocks of 8
    for (int task = tid; task < cg * rb; task += kThreads) {
      const int yb = y0 + (task / cg) * 8, x0 = (task - (task / cg) * cg) * 4;
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int i = 0; i < 8 + 2 * kR; ++i) {      // source rows yb - 11 .. yb + 18, reflected at the plane's borders
        int r = yb - kR + i;
        r = r < 0 ? -r : r;
        r = r >= s ? 2 * (s - 1) - r : r;
        r = min(max(r, lo), hi - 1);              // (rows beyond the band's range only feed output rows >= y1: dropped)
        const float4 t = *reinterpret_cast<const float4*>(sm + (size_t)(r - lo) * pitch + kPad + x0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = i - j;                    // tap of output row yb + j that reads source row yb - 11 + i
          if (k >= 0 && k < kTaps) {
            acc[j][0] = fmaf(w[k], t.x, acc[j][0]);
            acc[j][1] = fmaf(w[k], t.y, acc[j][1]);
            acc[j][2] = fmaf(w[k], t.z, acc[j][2]);
            acc[j][3] = fmaf(w[k], t.w, acc[j][3]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int y = yb + j;
        if (y < y1) {
          float o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float x = acc[j][c];
            if (sol) x = x >= MIS_SOLARIZE_THRESHOLD ? 1.f - x : x;
            o[c] = (x - mean) * inv_std;
          }
          if (a.out_f32) {
            *reinterpret_cast<float4*>(plane_ptr + ((size_t)y * s + x0) * 4) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
            *reinterpret_cast<uint2*>(plane_ptr + ((size_t)y * s + x0) * 2) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          }
        }
      }
    }
    __syncthreads();                              // the band's rows are rewritten by the next band's load
  }
}

}  // namespace augb
}  // namespace mis

using namespace mis;

extern "C" int mis_aug_blur_views(void* out, int out_dtype, const MisViewParams* params, int n_views, int C, int s,
                                  const float* mean, const float* std, void* stream) {
  using namespace mis::augb;
  MIS_REQUIRE(out && params && mean && std, MIS_ERR_INVALID_ARG, "mis_aug_blur_views: null pointer");
  MIS_REQUIRE(out_dtype == MIS_DTYPE_BF16 || out_dtype == MIS_DTYPE_F32, MIS_ERR_INVALID_ARG,
              "mis_aug_blur_views: out_dtype %d", out_dtype);
  MIS_REQUIRE(n_views >= 0 && C >= 1 && C <= 4, MIS_ERR_INVALID_ARG, "mis_aug_blur_views: bad sizes");
  MIS_REQUIRE(s >= 12 && s <= 256 && (s & 7) == 0, MIS_ERR_UNSUPPORTED,
              "mis_aug_blur_views: crop size %d (GaussianBlur(23) is fused for multiples of 8 in [16, 256]: reflect padding "
              "needs s > 11, two 128-row bands of the fp32 plane must fit one SM's shared memory)", s);
  MIS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, MIS_ERR_INVALID_ARG, "mis_aug_blur_views: out must be 16-byte aligned");
  for (int c = 0; c < C; ++c) MIS_REQUIRE(std[c] != 0.f, MIS_ERR_INVALID_ARG, "mis_aug_blur_views: std[%d] == 0", c);
  if (n_views == 0) return MIS_OK;
  BlurArgs a = {};
  a.out = out;
  a.params = params;
  a.C = C;
  a.s = s;
  a.out_f32 = out_dtype == MIS_DTYPE_F32 ? 1 : 0;
  a.pitch = s + 2 * kPad;
  a.nbands = s <= 224 ? 1 : 2;
  a.band_h = 128;
  a.band_rows = a.nbands == 1 ? s : a.band_h + 2 * kR;
  for (int c = 0; c < C; ++c) {
    a.mean[c] = mean[c];
    a.inv_std[c] = 1.0f / std[c];
  }
  const size_t smem = ((size_t)(a.band_rows + (a.nbands > 1 ? kR : 0)) * a.pitch + 32) * sizeof(float);
  MIS_CUDA_TRY(cudaFuncSetAttribute(blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  blur_kernel<<<dim3((unsigned)(n_views * C)), kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}
