// K1 -- fused two-view augmentation for 16-bit slices (sm_100a).
//
// One thread-block CLUSTER per output view plane, one CTA per band of 32 output rows:
//
//   producer warp : streams the crop-window rows the band needs from HBM into a shared-memory
//                   ring with 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP), one mbarrier
//                   per 8-row chunk, released by the consumers chunk by chunk;
//   V pass        : 8 consumer warps, lanes over source columns (u16x2 per lane, conflict-free),
//                   vertical antialias taps with warp-uniform weights -> fp32 tile in smem;
//   H pass        : lanes over the band's 32 output rows (conflict-free, odd row stride), each
//                   warp owns 32 consecutive output columns -> 32 results stay in registers;
//   colour        : brightness / contrast in the per-view fn_idx order; the contrast mean over the
//                   whole view is reduced warp -> CTA -> cluster through distributed shared
//                   memory (st.shared::cluster + barrier.cluster), so the view is written once;
//   store         : normalise, flip, convert to bf16 (or fp32) and write NCHW with 16-byte stores.
//
// HBM traffic per view = crop window (u16) once + output once: the resampled tile, the tap
// tables and the mean never leave the SM / cluster.
//
// Arithmetic restated from torchvision 0.26 / ATen (see oracle/aug_oracle.py):
//   taps   : _upsample_bilinear2d_aa (triangle filter, support = max(scale,1), weights normalised)
//   colour : functional/_color.py:114-125 (brightness), :190-205 + _blend :92-97 (contrast)
//   output : (x - mean) / std, functional/_misc.py:37-67
#include <cuda_bf16.h>

#include "common.cuh"

namespace mis {
namespace aug {

constexpr int kBandRows = 32;
constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kChunkRows = 8;
constexpr int kMaxChunks = 16;
constexpr int kMaxBands = 8;
constexpr int kSeg = 32;  // output columns per H-pass warp

struct Args {
  const uint16_t* src;
  int64_t img_stride;
  int C, H, W;
  const MisViewParams* params;
  float win_lo, win_scale;
  float mean[4], inv_std[4];
  void* out;
  int s;
  int nbands;
  int kv, kh;       // tap-table strides (>= max taps on that axis)
  int nch;          // ring depth in chunks
  int pitch;        // ring row pitch in bytes (multiple of 16)
  int pstr;         // tmp-plane row stride in words (odd)
  // shared-memory byte offsets
  int off_vw, off_hw, off_tmp, off_ring;
};

struct SmemHeader {
  uint64_t full[kMaxChunks];
  uint64_t empty[kMaxChunks];
  float part[kMaxBands];      // per-CTA partial sums of the contrast mean (written by peers)
  float red[kConsumerWarps];
  int v_min[kBandRows];
  int v_size[kBandRows];
  int h_min[256];
  int h_size[256];
};

// Tap table of one output index.  n = input size, scale = n/m in fp32.
__device__ __forceinline__ void aa_taps(int i, int n, float scale, float support, float invscale, int kmax,
                                        int& lo_out, int& size_out, float* w) {
  const float center = (float)((double)scale * ((double)i + 0.5));
  int lo = (int)((double)center - (double)support + 0.5);
  lo = lo < 0 ? 0 : lo;
  int hi = (int)((double)center + (double)support + 0.5);
  hi = hi > n ? n : hi;
  int size = hi - lo;
  size = size < 0 ? 0 : (size > kmax ? kmax : size);
  float total = 0.f;
  for (int j = 0; j < size; ++j) {
    const float arg = ((float)(j + lo) - center + 0.5f) * invscale;
    const float wj = fmaxf(0.f, 1.f - fabsf(arg));
    w[j] = wj;
    total += wj;
  }
  if (total != 0.f)
    for (int j = 0; j < size; ++j) w[j] = w[j] / total;
  for (int j = size; j < kmax; ++j) w[j] = 0.f;
  lo_out = lo;
  size_out = size;
}

template <bool kBulk, bool kWindow, bool kOutF32>
__global__ void __launch_bounds__(kThreads, 2) aug_kernel(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemHeader& sh = *reinterpret_cast<SmemHeader*>(smem);
  float* v_w = reinterpret_cast<float*>(smem + a.off_vw);
  float* h_w = reinterpret_cast<float*>(smem + a.off_hw);
  float* tmp = reinterpret_cast<float*>(smem + a.off_tmp);
  uint8_t* ring = smem + a.off_ring;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int band = blockIdx.x % a.nbands;       // == rank in cluster
  const int plane = blockIdx.x / a.nbands;      // view * C + c
  const int view = plane / a.C;
  const int chan = plane - view * a.C;
  const int s = a.s;

  cluster_arrive_relaxed();   // phase 1: "every CTA of the cluster is running" (waited before DSMEM use)

  const MisViewParams P = a.params[view];
  const int y0 = band * kBandRows;
  const int nrows = min(kBandRows, s - y0);
  const int lp = P.left & 1;                    // pair alignment of the crop's first column
  const int npairs = (lp + P.w + 1) >> 1;
  const int64_t plane_base = (int64_t)P.img * a.img_stride + (int64_t)chan * a.H * a.W;
  const int64_t e0 = plane_base + (int64_t)P.top * a.W + P.left;   // element index of crop (0,0)
  const int e0_lo = (int)(e0 & 7);

  // ---- tap tables + barriers ------------------------------------------------------------
  {
    const float vscale = (float)P.h / (float)s;
    const float hscale = (float)P.w / (float)s;
    const float vsup = vscale >= 1.f ? vscale : 1.f, vinv = vscale >= 1.f ? 1.f / vscale : 1.f;
    const float hsup = hscale >= 1.f ? hscale : 1.f, hinv = hscale >= 1.f ? 1.f / hscale : 1.f;
    for (int idx = tid; idx < nrows + s; idx += kThreads) {
      if (idx < nrows)
        aa_taps(y0 + idx, P.h, vscale, vsup, vinv, a.kv, sh.v_min[idx], sh.v_size[idx], v_w + idx * a.kv);
      else {
        const int x = idx - nrows;
        aa_taps(x, P.w, hscale, hsup, hinv, a.kh, sh.h_min[x], sh.h_size[x], h_w + x * a.kh);
      }
    }
    if (kBulk && tid == 0) {
      for (int i = 0; i < a.nch; ++i) {
        mbar_init(&sh.full[i], 1);
        mbar_init(&sh.empty[i], kConsumerWarps);
      }
      mbar_fence_init();
    }
  }
  __syncthreads();

  const int r_lo = sh.v_min[0];
  const int r_hi = sh.v_min[nrows - 1] + sh.v_size[nrows - 1];

  float o[kSeg];     // this thread's 32 output pixels (row y0+lane, columns 32*warp ..)
#pragma unroll
  for (int i = 0; i < kSeg; ++i) o[i] = 0.f;

  if (warp == kConsumerWarps) {
    // ================================ producer warp =======================================
    if (kBulk && lane == 0) {
      const int total_chunks = (r_hi - r_lo + kChunkRows - 1) / kChunkRows;
      for (int ci = 0; ci < total_chunks; ++ci) {
        const int slot = ci % a.nch;
        if (ci >= a.nch) mbar_wait(&sh.empty[slot], ((ci / a.nch) - 1) & 1);
        const int rbeg = r_lo + ci * kChunkRows;
        const int rend = min(rbeg + kChunkRows, r_hi);
        uint32_t total = 0;
        for (int r = rbeg; r < rend; ++r) {
          const int ph = (e0_lo + r * a.W) & 7;
          total += (uint32_t)((ph + P.w + 7) >> 3) << 4;
        }
        mbar_arrive_expect_tx(&sh.full[slot], total);
        for (int r = rbeg; r < rend; ++r) {
          const int64_t e = e0 + (int64_t)r * a.W;
          const int ph = (int)(e & 7);
          const uint32_t nb = (uint32_t)((ph + P.w + 7) >> 3) << 4;
          bulk_g2s(ring + (size_t)(slot * kChunkRows + (r - rbeg)) * a.pitch, a.src + (e - ph), nb, &sh.full[slot]);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ V pass ==============================================
    const int ring_rows = a.nch * kChunkRows;
    float* tmp_e = tmp;
    float* tmp_o = tmp + kBandRows * a.pstr;
    const uint16_t* gplane = a.src + (e0 - lp);     // 4-byte aligned (W even)
    int loaded = 0, released = 0;
    for (int yy = 0; yy < nrows; ++yy) {
      const int ymin = sh.v_min[yy];
      const int n = sh.v_size[yy];
      int sr0 = 0;
      if (kBulk) {
        const int need = (ymin + n - 1 - r_lo) / kChunkRows + 1;
        while (loaded < need) {
          mbar_wait(&sh.full[loaded % a.nch], (loaded / a.nch) & 1);
          ++loaded;
        }
        const int first = (ymin - r_lo) / kChunkRows;
        while (released < first) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh.empty[released % a.nch]);
          ++released;
        }
        sr0 = (ymin - r_lo) % ring_rows;
      }
      const float* wrow = v_w + yy * a.kv;
      for (int q = tid; q < npairs; q += kConsumerThreads) {
        float a0 = 0.f, a1 = 0.f;
        int sr = sr0;
        for (int j = 0; j < n; ++j) {
          const int r = ymin + j;
          uint32_t p;
          if (kBulk) {
            const int orow = ((e0_lo + r * a.W) & 7) - lp;      // element offset of pair 0 in the slot row
            p = *reinterpret_cast<const uint32_t*>(ring + (size_t)sr * a.pitch + 2 * orow + 4 * q);
            if (++sr == ring_rows) sr = 0;
          } else {
            p = __ldg(reinterpret_cast<const uint32_t*>(gplane + (int64_t)r * a.W) + q);
          }
          float f0 = (float)(p & 0xffffu);
          float f1 = (float)(p >> 16);
          if (kWindow) {
            f0 = fminf(fmaxf((f0 - a.win_lo) * a.win_scale, 0.f), 1.f);
            f1 = fminf(fmaxf((f1 - a.win_lo) * a.win_scale, 0.f), 1.f);
          }
          const float w = wrow[j];
          a0 = fmaf(f0, w, a0);
          a1 = fmaf(f1, w, a1);
        }
        tmp_e[yy * a.pstr + q] = a0;
        tmp_o[yy * a.pstr + q] = a1;
      }
    }
    bar_sync(1, kConsumerThreads);

    // ================================ H pass ==============================================
    // lane = band row, warp = 32-column segment; crop column k lives in plane (k+lp)&1 at (k+lp)>>1
    const int x0 = warp * kSeg;
    if (x0 < s) {
      const float* trow_e = tmp_e + lane * a.pstr;
      const float* trow_o = tmp_o + lane * a.pstr;
      const float post = kWindow ? 1.f : (1.f / 65535.f);
#pragma unroll
      for (int i = 0; i < kSeg; ++i) {
        const int x = x0 + i;
        if (x < s) {
          const int n = sh.h_size[x];
          int k = sh.h_min[x] + lp;
          const float* wrow = h_w + x * a.kh;
          float acc = 0.f;
          for (int j = 0; j < n; ++j, ++k) {
            const float v = (k & 1) ? trow_o[k >> 1] : trow_e[k >> 1];
            acc = fmaf(v, wrow[j], acc);
          }
          o[i] = acc * post;
        }
      }
    }
  }

  // ================================ colour ops =============================================
  cluster_wait_acquire();   // phase 1 done: all CTAs of the cluster are resident
  const bool row_ok = (warp < kConsumerWarps) && (lane < nrows);
  const int x0 = warp * kSeg;
  if (P.flags & MIS_VIEW_JITTER) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int op = P.order[k];
      if (op == 0) {
        const float b = P.brightness;
#pragma unroll
        for (int i = 0; i < kSeg; ++i) o[i] = fminf(fmaxf(o[i] * b, 0.f), 1.f);
      } else if (op == 1) {
        // mean over the whole view: thread -> warp -> CTA -> cluster (DSMEM)
        float part = 0.f;
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < kSeg; ++i)
            if (x0 + i < s) part += o[i];
        }
        part = warp_sum(part);
        if (warp < kConsumerWarps && lane == 0) sh.red[warp] = part;
        __syncthreads();
        if (tid == 0) {
          float tot = 0.f;
          for (int i = 0; i < kConsumerWarps; ++i) tot += sh.red[i];
          for (int r = 0; r < a.nbands; ++r) st_cluster_f32(&sh.part[band], (uint32_t)r, tot);
        }
        cluster_arrive_release();
        cluster_wait_acquire();
        float tot = 0.f;
        for (int r = 0; r < a.nbands; ++r) tot += sh.part[r];
        const float mu = tot / (float)(s * s);
        const float c = P.contrast;
        const float add = mu * (1.f - c);
#pragma unroll
        for (int i = 0; i < kSeg; ++i) o[i] = fminf(fmaxf(fmaf(o[i], c, add), 0.f), 1.f);
      }
      // op 2 (saturation) and 3 (hue) are identities for single-channel slices
    }
  }

  // ================================ normalise + store ======================================
  if (row_ok && x0 < s) {
    const float mean = a.mean[chan], inv_std = a.inv_std[chan];
#pragma unroll
    for (int i = 0; i < kSeg; ++i) o[i] = (o[i] - mean) * inv_std;
    const bool flip = (P.flags & MIS_VIEW_FLIP) != 0;
    const size_t row_off = ((size_t)plane * s + (y0 + lane)) * s;
    const bool full = (x0 + kSeg <= s) && ((s & 7) == 0);
    if (kOutF32) {
      float* out = reinterpret_cast<float*>(a.out) + row_off;
      if (full) {
        if (!flip) {
          float4* dst = reinterpret_cast<float4*>(out + x0);
#pragma unroll
          for (int i = 0; i < kSeg / 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        } else {
          float4* dst = reinterpret_cast<float4*>(out + (s - x0 - kSeg));
#pragma unroll
          for (int i = 0; i < kSeg / 4; ++i)
            dst[i] = make_float4(o[kSeg - 1 - 4 * i], o[kSeg - 2 - 4 * i], o[kSeg - 3 - 4 * i], o[kSeg - 4 - 4 * i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < kSeg; ++i)
          if (x0 + i < s) out[flip ? (s - 1 - x0 - i) : (x0 + i)] = o[i];
      }
    } else {
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out) + row_off;
      if (full) {
        uint32_t pk[kSeg / 2];
        if (!flip) {
#pragma unroll
          for (int i = 0; i < kSeg / 2; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
            pk[i] = *reinterpret_cast<uint32_t*>(&h);
          }
        } else {
#pragma unroll
          for (int i = 0; i < kSeg / 2; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(o[kSeg - 1 - 2 * i], o[kSeg - 2 - 2 * i]);
            pk[i] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (flip ? (s - x0 - kSeg) : x0));
#pragma unroll
        for (int i = 0; i < kSeg / 8; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < kSeg; ++i)
          if (x0 + i < s) out[flip ? (s - 1 - x0 - i) : (x0 + i)] = __float2bfloat16_rn(o[i]);
      }
    }
  }
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

template <bool kBulk, bool kWindow, bool kOutF32>
static int launch(const Args& a, int grid, size_t smem, cudaStream_t stream) {
  auto* fn = &aug_kernel<kBulk, kWindow, kOutF32>;
  MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.nbands;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, a));
  return MIS_OK;
}

}  // namespace aug
}  // namespace mis

using namespace mis;

extern "C" int mis_aug_two_view(const uint16_t* src, int n_images, int C, int H, int W, int64_t img_stride,
                                const MisViewParams* params, int n_views, float win_lo, float win_hi,
                                const float* mean, const float* std, void* out, int s, int out_dtype, int use_tma,
                                void* stream) {
  using namespace mis::aug;
  MIS_REQUIRE(src && params && mean && std && out, MIS_ERR_INVALID_ARG, "mis_aug_two_view: null pointer");
  MIS_REQUIRE(n_images > 0 && n_views >= 0 && H > 0 && W > 0, MIS_ERR_INVALID_ARG,
              "mis_aug_two_view: sizes must be positive (n_images=%d H=%d W=%d)", n_images, H, W);
  MIS_REQUIRE(out_dtype == MIS_DTYPE_BF16 || out_dtype == MIS_DTYPE_F32, MIS_ERR_INVALID_ARG,
              "mis_aug_two_view: out_dtype %d", out_dtype);
  MIS_REQUIRE(win_hi > win_lo, MIS_ERR_INVALID_ARG, "mis_aug_two_view: empty window [%g,%g]", win_lo, win_hi);
  MIS_REQUIRE(C == 1, MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: C=%d; only single-channel slices are implemented (3-channel saturation/hue "
              "are a SURVEY 8f 'next' row)", C);
  MIS_REQUIRE(s >= 8 && s <= kBandRows * kMaxBands, MIS_ERR_UNSUPPORTED, "mis_aug_two_view: crop size %d not in [8,256]", s);
  MIS_REQUIRE((W & 1) == 0 && (img_stride & 1) == 0, MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: W (%d) and img_stride must be even", W);
  MIS_REQUIRE(img_stride >= (int64_t)C * H * W, MIS_ERR_INVALID_ARG, "mis_aug_two_view: img_stride too small");
  MIS_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              MIS_ERR_INVALID_ARG, "mis_aug_two_view: src/out must be 16-byte aligned");
  for (int c = 0; c < C; ++c)
    MIS_REQUIRE(std[c] != 0.f, MIS_ERR_INVALID_ARG, "mis_aug_two_view: std[%d] == 0", c);
  if (n_views == 0) return MIS_OK;

  Args a = {};
  a.src = src;
  a.img_stride = img_stride;
  a.C = C;
  a.H = H;
  a.W = W;
  a.params = params;
  a.win_lo = win_lo;
  a.win_scale = 1.0f / (win_hi - win_lo);
  for (int c = 0; c < C; ++c) {
    a.mean[c] = mean[c];
    a.inv_std[c] = 1.0f / std[c];
  }
  a.out = out;
  a.s = s;
  a.nbands = (s + kBandRows - 1) / kBandRows;
  // worst-case taps per axis: support = max(size/s, 1), K = 2*ceil(support) + 1
  auto kmax = [&](int n) { int sup = (n + s - 1) / s; if (sup < 1) sup = 1; return 2 * sup + 1; };
  a.kv = kmax(H);
  a.kh = kmax(W);
  a.nch = (a.kv + kChunkRows - 1) / kChunkRows + 3;
  MIS_REQUIRE(a.nch <= kMaxChunks, MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: H/s = %d/%d needs a %d-chunk ring (max %d)", H, s, a.nch, kMaxChunks);
  a.pitch = align_up(2 * W + 32, 16);
  a.pstr = (W / 2 + 1) | 1;
  int off = align_up((int)sizeof(SmemHeader), 16);
  a.off_vw = off;
  off += align_up(kBandRows * a.kv * 4, 16);
  a.off_hw = off;
  off += align_up(s * a.kh * 4, 16);
  a.off_tmp = off;
  off += align_up(2 * kBandRows * a.pstr * 4, 128);
  a.off_ring = off;
  if (use_tma) off += a.nch * kChunkRows * a.pitch;
  const size_t smem = (size_t)off;
  MIS_REQUIRE(smem <= 227 * 1024, MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: needs %zu B of shared memory per CTA (H=%d W=%d s=%d)", smem, H, W, s);

  const bool window = !(win_lo == 0.f && win_hi == 65535.f);
  const bool f32 = out_dtype == MIS_DTYPE_F32;
  const int grid = a.nbands * n_views * C;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define MIS_AUG_DISPATCH(B, Wd, F) return launch<B, Wd, F>(a, grid, smem, st)
  if (use_tma) {
    if (window) { if (f32) MIS_AUG_DISPATCH(true, true, true); else MIS_AUG_DISPATCH(true, true, false); }
    else        { if (f32) MIS_AUG_DISPATCH(true, false, true); else MIS_AUG_DISPATCH(true, false, false); }
  } else {
    if (window) { if (f32) MIS_AUG_DISPATCH(false, true, true); else MIS_AUG_DISPATCH(false, true, false); }
    else        { if (f32) MIS_AUG_DISPATCH(false, false, true); else MIS_AUG_DISPATCH(false, false, false); }
  }
#undef MIS_AUG_DISPATCH
}

extern "C" int64_t mis_aug_algorithmic_bytes(const MisViewParams* p, int n_views, int C, int s, int out_dtype) {
  if (!p || n_views < 0) return -1;
  const int64_t ob = out_dtype == MIS_DTYPE_F32 ? 4 : 2;
  int64_t total = 0;
  for (int v = 0; v < n_views; ++v) total += 2 * (int64_t)C * p[v].h * p[v].w + ob * C * s * s;
  return total;
}
