// K1 -- fused two-view augmentation for 16-bit slices (sm_100a).
//
// One thread-block CLUSTER per output view plane, one CTA per band of 32 output rows:
//
//   staging       : (use_tma=1) a producer warp streams the band's crop rows into a shared-memory ring with 2-D TMA
//                   tensor-map boxes (cp.async.bulk.tensor.2d -> SASS UTMALDG; 8 rows x <=256 columns, inner start
//                   coordinate 16-byte aligned), one full/empty mbarrier pair per slot, started at kernel entry so it
//                   overlaps the table prologue; (use_tma=0) every consumer thread fetches its own four columns with
//                   8-byte cp.async into a private 8-deep ring after an early L2 prefetch of the band, two 16-row
//                   streams per band, no producer and no barriers;
//   V pass        : input-stationary: thread = 4 adjacent source columns, every source pixel is loaded and converted
//                   once (exact magic-number u16->f32, no I2F) and scattered into the <=3 output rows whose window
//                   contains it with packed fma.rn.f32x2 (FFMA2); output-stationary unrolled taps as fallback
//                   (upscaling / very wide images);
//   H pass        : lanes over the band's 32 output rows (conflict-free, odd row stride), each warp owns 32 consecutive
//                   output columns, per-view tap count unrolled -> 32 results stay in registers;
//   colour        : brightness / contrast in the per-view fn_idx order; the contrast mean over the whole view is
//                   reduced warp -> CTA -> cluster through distributed shared memory (st.shared::cluster +
//                   barrier.cluster), so the view is written once;
//   store         : normalise, flip, convert to bf16 (or fp32) and write NCHW with 16-byte stores.
//
// HBM traffic per view = crop window (u16) once + output once: the resampled tile, the tap tables and the mean never
// leave the SM / cluster.
//
// Arithmetic restated from torchvision 0.26 / ATen (see oracle/aug_oracle.py):
//   taps   : _upsample_bilinear2d_aa (triangle filter, support = max(scale,1), weights normalised)
//   colour : functional/_color.py:114-125 (brightness), :190-205 + _blend :92-97 (contrast)
//   output : (x - mean) / std, functional/_misc.py:37-67
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <mutex>

#include "aug_math.cuh"
#include "aug_strip.cuh"

namespace mis {
namespace augc {
bool rgb_supported(int s);
int launch_rgb_color(void* out, int out_f32, const MisViewParams* params, int n_views, int s, const float* mean,
                     const float* inv_std, cudaStream_t stream);
}  // namespace augc
}  // namespace mis

#include "aug_tile.cuh"
#include <cstring>

#include "common.cuh"

namespace mis {
namespace aug {

constexpr int kBandRows = 32;
constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kChunkRows = 8;     // rows per TMA box: one thread issues ~1 box per 455 clk whatever its size (measured)
constexpr int kRowUnroll = 4;     // rows whose loads are hoisted together in the V pass
constexpr int kBoxCols = 256;                           // widest TMA box (tensor-map limit)
constexpr int kBoxBytes = kBoxCols * 2 * kChunkRows;      // bytes one box occupies in a ring slot
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
  int out_f32;
  int nbands;
  int kstride;      // tap-table stride in floats (multiple of 4, >= unrolled tap count)
  int nch;          // ring depth in chunks (power of two)
  int nch_log2;
  int slot_bytes;   // bytes of one ring slot (= boxes per row * kBoxBytes)
  int64_t plane_rows;  // H (rows per plane in the [planes*H, W] tensor-map view)
  int pstr;         // tmp row stride in words (odd)
  int rmax;         // capacity of the input-stationary schedule (source rows per band)
  // shared-memory byte offsets
  int off_vw, off_hw, off_tmp, off_ring, off_sched;
  long long* dbg;   // optional [grid][8] clock stamps of thread 0 (debug / profiling aid), may be null
};

struct SmemHeader {
  uint64_t full[kMaxChunks];
  uint64_t empty[kMaxChunks];
  float part[kMaxBands];      // per-CTA partial sums of the contrast mean (written by peers)
  float red[kConsumerWarps];
  int4 v_info[kBandRows];     // {first source row, taps, chunks that must have landed, first chunk still needed}
  int h_min[256];
  int kv_max, kh_max;         // largest tap count of this band / this view
  int m_max;                  // most output rows any single source row of this band feeds
};

// ---- TMA-staged ring --------------------------------------------------------------------------
// A ring slot holds kChunkRows source rows of the crop.  The crop's columns are fetched as 2-D TMA boxes of
// up to 256 columns (box b covers crop columns [256b, 256b + wb)), wb rounded up to 64 so that one of four
// tensor maps (box widths 64/128/192/256) fits; box b lives at byte b*kBoxBytes of the slot, row pitch 2*wb.
// Tiled TMA needs a 16-byte aligned inner start coordinate (measured: an unaligned one raises 'illegal
// instruction' on sm_100a), so boxes start at column left & ~7 and crop column 0 sits at slot column left & 7.
// `w` in the helpers below is that total width (left & 7) + crop width.
__device__ __forceinline__ int box_width(int w, int b) {
  const int rem = w - b * kBoxCols;
  return rem >= kBoxCols ? kBoxCols : ((rem + 63) & ~63);
}
struct ColAddr {
  int off;     // byte offset of this thread's first column inside a slot
  int pitch;   // row pitch of the box holding it
};
__device__ __forceinline__ ColAddr col_addr(int col, int w) {
  const int b = col / kBoxCols;
  ColAddr ca;
  ca.off = b * kBoxBytes + (col - b * kBoxCols) * 2;
  ca.pitch = box_width(w, b) * 2;
  return ca;
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_box_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// Tap table of one output index (window arithmetic in aug_math.cuh).
// Writes kstride weights (zero padded) with element stride `wstep` (2 = duplicated pairs).
__device__ __forceinline__ void aa_taps(int i, int n, float scale, float support, float invscale, int kstride,
                                        int wstep, int& lo_out, int& size_out, float* w) {
  float center;
  int lo, hi;
  aa_window(i, n, scale, support, lo, hi, center);
  int size = hi - lo;
  size = size < 0 ? 0 : (size > kstride ? kstride : size);
  float total = 0.f;
  for (int j = 0; j < size; ++j) {
    const float arg = ((float)(j + lo) - center + 0.5f) * invscale;
    const float wj = fmaxf(0.f, 1.f - fabsf(arg));
    w[j * wstep] = wj;
    total += wj;
  }
  for (int j = 0; j < kstride; ++j) {
    float wj = 0.f;
    if (j < size) wj = (total != 0.f) ? w[j * wstep] / total : w[j * wstep];
    for (int t = 0; t < wstep; ++t) w[j * wstep + t] = wj;
  }
  lo_out = lo;
  size_out = size;
}

struct Ctx {
  const Args* a;
  SmemHeader* sh;
  const float* v_w;     // [32][kstride][2]  duplicated pairs
  const float* h_w;     // [s][kstride]
  float* tmp;           // [32][pstr]
  const uint8_t* ring;
  const uint16_t* gplane;   // crop (0, -lp) in global memory
  int tid, lane, warp;
  int nrows, npairs, lp, w, h, r_lo;
  int coff;             // tmp column of crop column 0 (left&1 for the pair paths, left&3 for the quad path)
  int ring_byte0;       // unused by the TMA ring (boxes start at the crop's first column)
};

// ---- V pass: lanes over source column pairs, K taps unrolled, packed FMAs -----------------------
template <int K, bool kBulk, bool kWindow>
__device__ __forceinline__ void v_pass(const Ctx& c, int ya, int yb, int ysub0) {
  const Args& a = *c.a;
  SmemHeader& sh = *c.sh;
  const int ring_mask = a.nch * kChunkRows - 1;
  const int npad = min(c.npairs + a.kstride / 2 + 1, a.pstr / 2);  // zero columns the H pass may touch
  int loaded = 0, released = 0;
  const uint64_t wsc = pack2(a.win_scale, a.win_scale);
  const uint64_t wof = pack2(-a.win_lo * a.win_scale, -a.win_lo * a.win_scale);
  for (int yy = ya; yy < yb; ++yy) {
    const int4 info = sh.v_info[yy];     // {ymin, n, need, first}
    if (kBulk) {
      while (loaded < info.z) {
        mbar_wait(&sh.full[loaded & (a.nch - 1)], (loaded >> a.nch_log2) & 1);
        ++loaded;
      }
      while (released < info.w) {
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&sh.empty[released & (a.nch - 1)]);
        ++released;
      }
    }
    const uint64_t* wrow = reinterpret_cast<const uint64_t*>(c.v_w + yy * a.kstride * 2);
    float* trow = c.tmp + (yy - ysub0) * a.pstr;
    for (int q = c.tid; q < npad; q += kConsumerThreads) {
      uint64_t acc = 0ull;   // (+0.f, +0.f)
      if (q < c.npairs) {
        uint32_t p[K];
        if (kBulk) {
          const ColAddr ca = col_addr(2 * q, c.coff + c.w);
          const int sr0 = info.x - c.r_lo;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const int sr = (sr0 + j) & ring_mask;
            p[j] = *reinterpret_cast<const uint32_t*>(c.ring + (sr / kChunkRows) * a.slot_bytes + ca.off +
                                                      (sr % kChunkRows) * ca.pitch);
          }
        } else {
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const int r = min(info.x + j, c.h - 1);
            p[j] = __ldg(reinterpret_cast<const uint32_t*>(c.gplane + (int64_t)r * a.W) + q);
          }
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
          uint64_t f = u16x2_to_f32x2(p[j]);
          if (kWindow) {
            f = ffma2(f, wsc, wof);
            float f0, f1;
            unpack2(f, f0, f1);
            f = pack2(fminf(fmaxf(f0, 0.f), 1.f), fminf(fmaxf(f1, 0.f), 1.f));
          }
          acc = ffma2(f, wrow[j], acc);
        }
      }
      float a0, a1;
      unpack2(acc, a0, a1);
      trow[2 * q] = a0;
      trow[2 * q + 1] = a1;
    }
  }
}


// Input-stationary schedule: source row rr feeds output rows first .. first+2 with (duplicated, FFMA2-ready) weights.
// TMA path: w[k] belongs to output row first + k and the three accumulators rotate when a row completes.
// cp.async path: output row ya + j of a stream accumulates in the FIXED accumulator j % 3 and w[j % 3] holds its weight.
// Rows before `first` are complete when rr is reached.
struct __align__(16) SchedRow {
  float w[3][2];
  int first;
  int pad;
};

// ---- V pass, input-stationary: every source pixel is loaded and converted ONCE ------------------------
// Thread = one source column pair, streaming down the band's source rows in ring order with three running
// accumulators (the output rows whose vertical window contains the current source row).
template <bool kBulk, bool kWindow>
__device__ __forceinline__ void v_pass_is(const Ctx& c, const SchedRow* __restrict__ sched, int nsrc) {
  const Args& a = *c.a;
  SmemHeader& sh = *c.sh;
  // thread = FOUR adjacent source columns (one 8-byte load per source row); column group 0 starts at the crop's
  // first column rounded down to a multiple of 4 (c.coff = left & 3 columns of slack on the left)
  const int ngroups = (c.coff + c.w + 3) >> 2;
  const int q = c.tid;
  const bool active = q < ngroups;
  const bool warp_idle = (c.warp * 32) >= ngroups;
  const int pstr = a.pstr;
  const int wq = a.W >> 2;
  // zero the few columns right of the crop that the unrolled H-pass taps may touch (weights there are 0)
  {
    const int c0 = 4 * ngroups, per = min(a.kstride + 2, pstr - c0);
    for (int i = c.tid; i < c.nrows * per; i += kConsumerThreads) c.tmp[(i / per) * pstr + c0 + (i % per)] = 0.f;
  }
  const uint64_t wsc = pack2(a.win_scale, a.win_scale);
  const uint64_t wof = pack2(-a.win_lo * a.win_scale, -a.win_lo * a.win_scale);
  const int qa = active ? q : 0;                         // inactive lanes compute on a valid address, never store
  const ColAddr ca = col_addr(4 * qa, c.coff + c.w);
  const int pitch = ca.pitch;
  const uint8_t* rbase = c.ring + ca.off;
  const uint2* gp = reinterpret_cast<const uint2*>(c.gplane + (int64_t)c.r_lo * a.W) + qa;
  float* tp = c.tmp + 4 * qa;                            // next output row of this column group
  uint64_t a0 = 0ull, a1 = 0ull, b0 = 0ull, b1 = 0ull, c0 = 0ull, c1 = 0ull;   // 3 open output rows x 4 columns
  int ycur = 0;
  const int nchunks = (nsrc + kChunkRows - 1) / kChunkRows;
  const SchedRow* sp = sched;

  auto flush = [&](int target) {                           // uniform: output rows [ycur, target) are complete
#pragma unroll 1
    for (; ycur < target; ++ycur) {
      if (active) {
        float v0, v1, v2, v3;
        unpack2(a0, v0, v1);
        unpack2(a1, v2, v3);
        tp[0] = v0;
        tp[1] = v1;
        tp[2] = v2;
        tp[3] = v3;
      }
      tp += pstr;
      a0 = b0; a1 = b1;
      b0 = c0; b1 = c1;
      c0 = 0ull; c1 = 0ull;
    }
  };
  auto conv = [&](uint32_t p) {
    uint64_t f = u16x2_to_f32x2(p);
    if (kWindow) {
      f = ffma2(f, wsc, wof);
      float f0, f1;
      unpack2(f, f0, f1);
      f = pack2(fminf(fmaxf(f0, 0.f), 1.f), fminf(fmaxf(f1, 0.f), 1.f));
    }
    return f;
  };
  auto row = [&](uint2 p, const float4& s0, const float4& s1) {
    flush(__float_as_int(s1.z));
    const uint64_t f0 = conv(p.x), f1 = conv(p.y);
    const uint64_t w0 = pack2(s0.x, s0.y), w1 = pack2(s0.z, s0.w), w2 = pack2(s1.x, s1.y);
    a0 = ffma2(f0, w0, a0);
    a1 = ffma2(f1, w0, a1);
    b0 = ffma2(f0, w1, b0);
    b1 = ffma2(f1, w1, b1);
    c0 = ffma2(f0, w2, c0);
    c1 = ffma2(f1, w2, c1);
  };

#pragma unroll 1
  for (int ci = 0; ci < nchunks; ++ci) {
    if (kBulk) mbar_wait(&sh.full[ci & (a.nch - 1)], (ci >> a.nch_log2) & 1);
    if (!warp_idle) {
      const int rr0 = ci * kChunkRows;
      const uint8_t* rp = rbase + (size_t)(ci & (a.nch - 1)) * a.slot_bytes;
      const int nrow = min(kChunkRows, nsrc - rr0);
      int k0 = 0;
      for (; k0 + kRowUnroll <= nrow; k0 += kRowUnroll) {  // groups of 4 rows: loads hoisted, no per-row loop overhead
        uint2 p[kRowUnroll];
        float4 s0[kRowUnroll], s1[kRowUnroll];
#pragma unroll
        for (int k = 0; k < kRowUnroll; ++k) {
          if (kBulk) p[k] = *reinterpret_cast<const uint2*>(rp + (k0 + k) * pitch);
          else p[k] = __ldg(gp + (k0 + k) * wq);
          s0[k] = *reinterpret_cast<const float4*>(&sp[k0 + k].w[0][0]);
          s1[k] = *reinterpret_cast<const float4*>(&sp[k0 + k].w[2][0]);
        }
#pragma unroll
        for (int k = 0; k < kRowUnroll; ++k) row(p[k], s0[k], s1[k]);
      }
      for (; k0 < nrow; ++k0) {
        uint2 p;
        if (kBulk) p = *reinterpret_cast<const uint2*>(rp + k0 * pitch);
        else p = __ldg(gp + k0 * wq);
        row(p, *reinterpret_cast<const float4*>(&sp[k0].w[0][0]), *reinterpret_cast<const float4*>(&sp[k0].w[2][0]));
      }
      gp += kChunkRows * wq;
      sp += kChunkRows;
    }
    if (kBulk) {
      __syncwarp();
      if (c.lane == 0) mbar_arrive(&sh.empty[ci & (a.nch - 1)]);
    }
  }
  if (!warp_idle) flush(c.nrows);
}


// ---- V pass without TMA: per-thread cp.async ring, two independent row streams -----------------------------
// Measured on B200: one thread issues at most ~1 TMA box per 455 clk, so a producer-fed ring delivers only
// ~4.5-9 B/clk per CTA with the 2-4 KB boxes that fit next to the 68 KB transposition tile.  Here every
// consumer thread instead fetches ITS OWN four columns with 8-byte cp.async (LDGSTS) into a private 8-deep
// ring: no producer, no mbarriers, no cross-thread hazards (a thread only reads what it copied itself).
// The band's 32 output rows are split in two streams of 16 rows (warps 0-3 / 4-7), each walking its own source
// row range, which halves the serial dependency chain per warp.
constexpr int kCpDepth = 6;      // rows in flight per thread (the band is L2-prefetched at kernel entry)
constexpr int kCpSlotBytes = 16; // bytes one thread stages per source row (8 columns; the 4-column variant uses half)
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// VEC = source columns per thread (4 -> 8-byte copies, 8 -> 16-byte copies); `t` = thread index inside the stream.
template <int VEC, bool kWindow>
__device__ __forceinline__ void v_pass_cp(const Ctx& c, const SchedRow* __restrict__ sched, int nsrc, int r_first,
                                          int ya, int yb, int ysub0, uint8_t* __restrict__ ring, int t) {
  constexpr int NP = VEC / 2;                                  // packed pairs per thread
  constexpr int kStreamRows = kBandRows / 4;                   // output rows per stream (8)
  const Args& a = *c.a;
  const int ngroups = (c.coff + c.w + VEC - 1) / VEC;
  const bool active = t < ngroups;
  const int ta = active ? t : 0;
  const int pstr = a.pstr;
  const size_t row_bytes = (size_t)a.W * 2;
  const uint64_t wsc = pack2(a.win_scale, a.win_scale);
  const uint64_t wof = pack2(-a.win_lo * a.win_scale, -a.win_lo * a.win_scale);
  const uint8_t* gnext = reinterpret_cast<const uint8_t*>(c.gplane + (int64_t)r_first * a.W) + (size_t)ta * VEC * 2;
  uint8_t* myr = ring + (size_t)c.tid * (VEC * 2);             // lanes contiguous: conflict-free LDGSTS / LDS
  constexpr int kSlotStride = kConsumerThreads * kCpSlotBytes;  // ring slot k of this thread: myr + k * kSlotStride
  float* tp = c.tmp + (ya - ysub0) * pstr + VEC * ta;    // the tile holds one sub-band starting at output row ysub0
  uint64_t acc[3][NP];                                         // output row ya + j accumulates in acc[j % 3] (static)
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[k][i] = 0ull;

  auto conv = [&](uint32_t p) {
    uint64_t f = u16x2_to_f32x2(p);
    if (kWindow) {
      f = ffma2(f, wsc, wof);
      float f0, f1;
      unpack2(f, f0, f1);
      f = pack2(fminf(fmaxf(f0, 0.f), 1.f), fminf(fmaxf(f1, 0.f), 1.f));
    }
    return f;
  };
  int to_fetch = nsrc;                                         // rows not yet requested
#pragma unroll
  for (int k = 0; k < kCpDepth; ++k) {
    if (to_fetch > 0) cp_async<VEC * 2>(myr + k * kSlotStride, gnext);
    cp_async_commit();
    gnext += row_bytes;
    --to_fetch;
  }
  const SchedRow* sp = sched;
  const SchedRow* const sp_end = sched + nsrc;
  uint8_t* slot = myr;                                         // ring slot of the next source row
  uint8_t* const slot_end = myr + kCpDepth * kSlotStride;
  // Static unroll over the stream's output rows: while the next source row still belongs to output row ya + j
  // (its `first` open row is <= ya + j) it is accumulated into the three fixed accumulators; then row j is stored.
#pragma unroll
  for (int j = 0; j < kStreamRows; ++j) {
    if (ya + j < yb) {                                          // uniform
#pragma unroll 1
      while (sp < sp_end) {
        const float4 s1 = *reinterpret_cast<const float4*>(&sp->w[2][0]);
        if (__float_as_int(s1.z) > ya + j) break;              // uniform: this source row opens a later output row
        const float4 s0 = *reinterpret_cast<const float4*>(&sp->w[0][0]);
        cp_async_wait<kCpDepth - 1>();                          // the copy of this row has landed
        uint32_t p[NP];
        if (VEC == 8) {
          const uint4 v = *reinterpret_cast<const uint4*>(slot);
          p[0] = v.x; p[1] = v.y; p[NP - 2] = v.z; p[NP - 1] = v.w;
        } else {
          const uint2 v = *reinterpret_cast<const uint2*>(slot);
          p[0] = v.x; p[NP - 1] = v.y;
        }
        if (to_fetch > 0) cp_async<VEC * 2>(slot, gnext);       // refill the slot just read
        cp_async_commit();
        gnext += row_bytes;
        --to_fetch;
        ++sp;
        slot += kSlotStride;
        if (slot == slot_end) slot = myr;
        const uint64_t w0 = pack2(s0.x, s0.y), w1 = pack2(s0.z, s0.w), w2 = pack2(s1.x, s1.y);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const uint64_t f = conv(p[i]);
          acc[0][i] = ffma2(f, w0, acc[0][i]);
          acc[1][i] = ffma2(f, w1, acc[1][i]);
          acc[2][i] = ffma2(f, w2, acc[2][i]);
        }
      }
      if (active) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          float v0, v1;
          unpack2(acc[j % 3][i], v0, v1);
          tp[j * pstr + 2 * i] = v0;
          tp[j * pstr + 2 * i + 1] = v1;
        }
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) acc[j % 3][i] = 0ull;
    }
  }
  cp_async_wait<0>();
}

// dynamic-length fallback (tap counts outside the unrolled set)
template <bool kBulk, bool kWindow>
__device__ __forceinline__ void v_pass_dyn(const Ctx& c, int ya, int yb, int ysub0) {
  const Args& a = *c.a;
  SmemHeader& sh = *c.sh;
  const int ring_mask = a.nch * kChunkRows - 1;
  const int npad = min(c.npairs + a.kstride / 2 + 1, a.pstr / 2);
  int loaded = 0, released = 0;
  for (int yy = ya; yy < yb; ++yy) {
    const int4 info = sh.v_info[yy];
    if (kBulk) {
      while (loaded < info.z) {
        mbar_wait(&sh.full[loaded & (a.nch - 1)], (loaded >> a.nch_log2) & 1);
        ++loaded;
      }
      while (released < info.w) {
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&sh.empty[released & (a.nch - 1)]);
        ++released;
      }
    }
    const float* wrow = c.v_w + yy * a.kstride * 2;
    float* trow = c.tmp + (yy - ysub0) * a.pstr;
    for (int q = c.tid; q < npad; q += kConsumerThreads) {
      float a0 = 0.f, a1 = 0.f;
      if (q < c.npairs) {
        for (int j = 0; j < info.y; ++j) {
          uint32_t p;
          if (kBulk)
          {
            const ColAddr ca = col_addr(2 * q, c.coff + c.w);
            const int sr = (info.x - c.r_lo + j) & ring_mask;
            p = *reinterpret_cast<const uint32_t*>(c.ring + (sr / kChunkRows) * a.slot_bytes + ca.off +
                                                   (sr % kChunkRows) * ca.pitch);
          }
          else
            p = __ldg(reinterpret_cast<const uint32_t*>(c.gplane + (int64_t)(info.x + j) * a.W) + q);
          float f0, f1;
          unpack2(u16x2_to_f32x2(p), f0, f1);
          if (kWindow) {
            f0 = fminf(fmaxf((f0 - a.win_lo) * a.win_scale, 0.f), 1.f);
            f1 = fminf(fmaxf((f1 - a.win_lo) * a.win_scale, 0.f), 1.f);
          }
          const float w = wrow[2 * j];
          a0 = fmaf(f0, w, a0);
          a1 = fmaf(f1, w, a1);
        }
      }
      trow[2 * q] = a0;
      trow[2 * q + 1] = a1;
    }
  }
}

// ---- H pass: lanes over the band's rows, one warp per 32 output columns, K taps unrolled --------
template <int K>
__device__ __forceinline__ void h_pass(const Ctx& c, float (&o)[kSeg], float post) {
  const Args& a = *c.a;
  const int x0 = c.warp * kSeg;
  const float* trow = c.tmp + c.lane * a.pstr + c.coff;
  constexpr int K4 = (K + 3) / 4;
#pragma unroll
  for (int i = 0; i < kSeg; ++i) {
    const int x = x0 + i;
    if (x < a.s) {
      const float* tp = trow + c.sh->h_min[x];
      const float4* wp = reinterpret_cast<const float4*>(c.h_w + x * a.kstride);
      float acc = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < K4; ++j4) {
        const float4 w = wp[j4];
        acc = fmaf(tp[4 * j4], w.x, acc);
        if (4 * j4 + 1 < K) acc = fmaf(tp[4 * j4 + 1], w.y, acc);
        if (4 * j4 + 2 < K) acc = fmaf(tp[4 * j4 + 2], w.z, acc);
        if (4 * j4 + 3 < K) acc = fmaf(tp[4 * j4 + 3], w.w, acc);
      }
      o[i] = acc * post;
    }
  }
}

__device__ __forceinline__ void h_pass_dyn(const Ctx& c, float (&o)[kSeg], float post) {
  const Args& a = *c.a;
  const int x0 = c.warp * kSeg;
  const float* trow = c.tmp + c.lane * a.pstr + c.coff;
#pragma unroll
  for (int i = 0; i < kSeg; ++i) {
    const int x = x0 + i;
    if (x < a.s) {
      const float* tp = trow + c.sh->h_min[x];
      const float* wp = c.h_w + x * a.kstride;
      float acc = 0.f;
      for (int j = 0; j < c.sh->kh_max; ++j) acc = fmaf(tp[j], wp[j], acc);
      o[i] = acc * post;
    }
  }
}


// ---- H pass for 16-row sub-bands: lane = (row 0..15, column half 0..1); a warp still owns 32 output columns --------
template <int K>
__device__ __forceinline__ void h_pass16(const Ctx& c, float* __restrict__ o, float post) {
  const Args& a = *c.a;
  const int xh = c.warp * kSeg + 16 * (c.lane >> 4);            // first column of this thread's 16-column run
  const float* trow = c.tmp + (c.lane & 15) * a.pstr + c.coff;
  constexpr int K4 = (K + 3) / 4;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int x = xh + i;
    if (x < a.s) {
      const float* tp = trow + c.sh->h_min[x];
      const float4* wp = reinterpret_cast<const float4*>(c.h_w + x * a.kstride);
      float acc = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < K4; ++j4) {
        const float4 w = wp[j4];
        acc = fmaf(tp[4 * j4], w.x, acc);
        if (4 * j4 + 1 < K) acc = fmaf(tp[4 * j4 + 1], w.y, acc);
        if (4 * j4 + 2 < K) acc = fmaf(tp[4 * j4 + 2], w.z, acc);
        if (4 * j4 + 3 < K) acc = fmaf(tp[4 * j4 + 3], w.w, acc);
      }
      o[i] = acc * post;
    }
  }
}
__device__ __forceinline__ void h_pass16_dyn(const Ctx& c, float* __restrict__ o, float post) {
  const Args& a = *c.a;
  const int xh = c.warp * kSeg + 16 * (c.lane >> 4);
  const float* trow = c.tmp + (c.lane & 15) * a.pstr + c.coff;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int x = xh + i;
    if (x < a.s) {
      const float* tp = trow + c.sh->h_min[x];
      const float* wp = c.h_w + x * a.kstride;
      float acc = 0.f;
      for (int j = 0; j < c.sh->kh_max; ++j) acc = fmaf(tp[j], wp[j], acc);
      o[i] = acc * post;
    }
  }
}

template <bool kBulk, bool kWindow>
__global__ void __launch_bounds__(kThreads, kBulk ? 2 : 3)
aug_kernel(const __grid_constant__ CUtensorMap map64, const __grid_constant__ CUtensorMap map128,
           const __grid_constant__ CUtensorMap map192, const __grid_constant__ CUtensorMap map256, const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  SmemHeader& sh = *reinterpret_cast<SmemHeader*>(smem);
  float* v_w = reinterpret_cast<float*>(smem + a.off_vw);
  float* h_w = reinterpret_cast<float*>(smem + a.off_hw);
  float* tmp = reinterpret_cast<float*>(smem + a.off_tmp);
  uint8_t* ring = smem + a.off_ring;
  SchedRow* sched = reinterpret_cast<SchedRow*>(smem + a.off_sched);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int band = blockIdx.x % a.nbands;       // == rank in cluster
  const int plane = blockIdx.x / a.nbands;      // view * C + c
  const int view = plane / a.C;
  const int chan = plane - view * a.C;
  const int s = a.s;

  cluster_arrive_relaxed();   // phase 1: "every CTA of the cluster is running" (waited before DSMEM use)
#ifdef MIS_DEBUG
#define MIS_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[(size_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)
#else
#define MIS_STAMP(i) do { } while (0)
#endif
  MIS_STAMP(0);

  const MisViewParams P = a.params[view];
  const int y0 = band * kBandRows;
  const int nrows = min(kBandRows, s - y0);
  const int lp = P.left & 1;                    // pair alignment of the crop's first column
  const int64_t plane_base = (int64_t)P.img * a.img_stride + (int64_t)chan * a.H * a.W;
  const int64_t e0 = plane_base + (int64_t)P.top * a.W + P.left;   // element index of crop (0,0)

  // ---- producer head start + tap tables ------------------------------------------------------
  // The producer lane derives the band's source-row range on its own (two window evaluations), prefetches the
  // whole range into L2 and fills the ring, all while the 8 consumer warps build the tap tables.
  const float vscale = (float)P.h / (float)s;
  const float hscale = (float)P.w / (float)s;
  const float vsup = vscale >= 1.f ? vscale : 1.f, vinv = vscale >= 1.f ? 1.f / vscale : 1.f;
  const float hsup = hscale >= 1.f ? hscale : 1.f, hinv = hscale >= 1.f ? 1.f / hscale : 1.f;
  // one chunk = kChunkRows rows x the crop's columns, fetched as ceil(w/256) TMA boxes into ring slot `slot`
  const int wtot = (P.left & 7) + P.w;                            // slot columns: aligned box start .. crop end
  const int nbox = (wtot + kBoxCols - 1) / kBoxCols;
  uint32_t chunk_bytes = 0;
  for (int b = 0; b < nbox; ++b) chunk_bytes += (uint32_t)box_width(wtot, b) * 2 * kChunkRows;
  const int row_base = (P.img * a.C + chan) * (int)a.plane_rows + P.top;    // tensor-map row of crop row 0
  // (macro, not a lambda: the tensor maps must be addressed in param space, a closure would copy them to local memory)
#define MIS_ISSUE_CHUNK(slot_, r_)                                                                                  \
  do {                                                                                                              \
    mbar_arrive_expect_tx(&sh.full[slot_], chunk_bytes);                                                            \
    for (int b_ = 0; b_ < nbox; ++b_) {                                                                             \
      const int wb_ = box_width(wtot, b_);                                                                           \
      void* dst_ = ring + (size_t)(slot_) * a.slot_bytes + b_ * kBoxBytes;                                          \
      const int c0_ = (P.left & ~7) + b_ * kBoxCols, c1_ = row_base + (r_);                                                \
      if (wb_ == 64) tma_box_2d(dst_, &map64, c0_, c1_, &sh.full[slot_]);                                           \
      else if (wb_ == 128) tma_box_2d(dst_, &map128, c0_, c1_, &sh.full[slot_]);                                    \
      else if (wb_ == 192) tma_box_2d(dst_, &map192, c0_, c1_, &sh.full[slot_]);                                    \
      else tma_box_2d(dst_, &map256, c0_, c1_, &sh.full[slot_]);                                                    \
    }                                                                                                               \
  } while (0)
#define MIS_PREFETCH_CHUNK(r_)                                                                                      \
  do {                                                                                                              \
    for (int b_ = 0; b_ < nbox; ++b_) {                                                                             \
      const int wb_ = box_width(wtot, b_);                                                                          \
      const int c0_ = (P.left & ~7) + b_ * kBoxCols, c1_ = row_base + (r_);                                         \
      if (wb_ == 64) tma_prefetch_2d(&map64, c0_, c1_);                                                             \
      else if (wb_ == 128) tma_prefetch_2d(&map128, c0_, c1_);                                                      \
      else if (wb_ == 192) tma_prefetch_2d(&map192, c0_, c1_);                                                      \
      else tma_prefetch_2d(&map256, c0_, c1_);                                                                      \
    }                                                                                                               \
  } while (0)
  int issued = 0;                                                // chunks already issued by the head start
  if (!kBulk) {
    // non-TMA path: pull the band's crop rows into L2 right away (every thread derives the row range itself and
    // prefetches a few 128-byte lines), so the cp.async stream of the V pass, which starts ~3 us later, hits L2
    int lo0, hi0, lo1, hi1;
    float ctr;
    aa_window(y0, P.h, vscale, vsup, lo0, hi0, ctr);
    aa_window(y0 + nrows - 1, P.h, vscale, vsup, lo1, hi1, ctr);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a.src + e0);
    const int head = (int)(reinterpret_cast<uintptr_t>(base) & 127);
    const int lines = (head + 2 * P.w + 127) >> 7;               // 128-byte lines per crop row
    const int total = (hi1 - lo0) * lines;
    for (int i = tid; i < total; i += kThreads) {
      const int r = lo0 + i / lines, l = i - (i / lines) * lines;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)r * a.W * 2 - head + l * 128));
    }
  }
  if (tid == 0) {
    sh.kv_max = 0;
    sh.kh_max = 0;
    sh.m_max = 0;
  }
  if (kBulk && warp == kConsumerWarps) {
    if (lane == 0) {
      for (int i = 0; i < a.nch; ++i) {
        mbar_init(&sh.full[i], 1);
        mbar_init(&sh.empty[i], kConsumerWarps);
      }
      mbar_fence_init();
      int lo0, hi0, lo1, hi1;
      float ctr;
      aa_window(y0, P.h, vscale, vsup, lo0, hi0, ctr);
      aa_window(y0 + nrows - 1, P.h, vscale, vsup, lo1, hi1, ctr);
      const int total = (hi1 - lo0 + kChunkRows - 1) / kChunkRows;
      issued = min(a.nch, total);
      for (int ci = 0; ci < issued; ++ci) MIS_ISSUE_CHUNK(ci, lo0 + ci * kChunkRows);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp < kConsumerWarps) {
    for (int idx = tid; idx < nrows + s; idx += kConsumerThreads) {
      int lo, size;
      if (idx < nrows) {
        aa_taps(y0 + idx, P.h, vscale, vsup, vinv, a.kstride, 2, lo, size, v_w + idx * a.kstride * 2);
        sh.v_info[idx] = make_int4(lo, size, 0, 0);
        atomicMax(&sh.kv_max, size);
      } else {
        const int x = idx - nrows;
        aa_taps(x, P.w, hscale, hsup, hinv, a.kstride, 1, lo, size, h_w + x * a.kstride);
        sh.h_min[x] = lo;
        atomicMax(&sh.kh_max, size);
      }
    }
  }
  __syncthreads();
  const int r_lo = sh.v_info[0].x;
  const int r_hi = sh.v_info[nrows - 1].x + sh.v_info[nrows - 1].y;
  if (tid < nrows) {   // chunk bookkeeping of each band row (uses the unrolled tap count, see below)
    int4 info = sh.v_info[tid];
    info.w = (info.x - r_lo) / kChunkRows;
    sh.v_info[tid] = info;
  }
  // unrolled tap counts of this view (uniform over the CTA)
  auto round_k = [](int k) { return k <= 3 ? 3 : k <= 5 ? 5 : k <= 7 ? 7 : k <= 9 ? 9 : k <= 13 ? 13 : 0; };
  const int KV = round_k(sh.kv_max);
  const int KH = round_k(sh.kh_max);
  const int total_chunks = (r_hi - r_lo + kChunkRows - 1) / kChunkRows;
  __syncthreads();
  if (tid < nrows) {
    int4 info = sh.v_info[tid];
    // rows [ymin, ymin+KV) are read (zero weights beyond the window): they only need to be *loaded*
    // when they carry weight, but they must never be overwritten mid-read -> wait for the real window.
    const int last = info.x + info.y - 1 - r_lo;
    info.z = min(last / kChunkRows + 1, total_chunks);
    sh.v_info[tid] = info;
  }
  // input-stationary schedules (windows are monotone in y): one entry per source row feeding output rows [ya, yb)
  auto build_sched = [&](SchedRow* out, int ya, int yb, int cap) {
    const int rlo = sh.v_info[ya].x;
    const int n = sh.v_info[yb - 1].x + sh.v_info[yb - 1].y - rlo;
    for (int rr = tid; rr < min(n, cap); rr += kThreads) {
      const int r = rlo + rr;
      int first = yb, last = -1;
      for (int yy = ya; yy < yb; ++yy) {
        const int4 info = sh.v_info[yy];
        if (info.x <= r && r < info.x + info.y) {
          first = min(first, yy);
          last = yy;
        }
      }
      if (last < 0) first = ya;   // cannot happen (windows overlap); keeps the flush logic monotone anyway
      SchedRow e;
#pragma unroll
      for (int k = 0; k < 3; ++k) e.w[k][0] = e.w[k][1] = 0.f;
      for (int yy = first; yy <= last && yy < first + 3; ++yy) {     // output row yy accumulates in slot (yy - ya) % 3
        const float w = v_w[(yy * a.kstride + (r - sh.v_info[yy].x)) * 2];
        const int slot = kBulk ? (yy - first) : ((yy - ya) % 3);   // TMA path rotates accumulators, cp path keeps them fixed
        e.w[slot][0] = w;
        e.w[slot][1] = w;
      }
      e.first = first;
      e.pad = 0;
      out[rr] = e;
      atomicMax(&sh.m_max, last - first + 1);
    }
    return n;
  };
  const int nsrc = r_hi - r_lo;
  // TMA path: one stream over the whole band.  cp.async path: NS independent streams of 32/NS output rows, each
  // owned by 256/NS threads: 4 streams x 8 columns per thread when the crop fits 64 x 8 columns and rows are
  // 16-byte aligned (W % 8 == 0), else 2 streams x 4 columns.
  // (the 8-columns-per-thread variant is kept for experiments; the sub-band layout below uses 4 columns)
  const bool wide8 = false;
  const int ns = 4;                                              // 2 sub-bands x 2 streams of 8 output rows
  const int rows_per_stream = kBandRows / ns;
  const int cap_s = a.rmax / 4;
  int nsrc_s[4] = {0, 0, 0, 0};
  if (kBulk) {
    build_sched(sched, 0, nrows, a.rmax);
  } else {
    for (int st = 0; st < ns; ++st) {
      const int ya = st * rows_per_stream, yb = min(ya + rows_per_stream, nrows);
      if (ya < nrows) nsrc_s[st] = build_sched(sched + st * cap_s, ya, yb, cap_s);
    }
  }
  __syncthreads();
  const int nsrc_max = max(max(nsrc_s[0], nsrc_s[1]), max(nsrc_s[2], nsrc_s[3]));
  const bool use_is = kBulk ? ((sh.m_max <= 3) && (nsrc <= a.rmax) && ((a.W & 3) == 0) &&
                               ((((P.left & 7) + P.w + 3) >> 2) <= kConsumerThreads))
                            : ((sh.m_max <= 3) && (nsrc_max <= cap_s) && ((a.W & 3) == 0) && ((a.img_stride & 3) == 0) &&
                               ((((P.left & 3) + P.w + 3) >> 2) <= kConsumerThreads / 2));

  MIS_STAMP(1);   // tables + schedule done
  float o[kSeg];     // this thread's 32 output pixels (row y0+lane, columns 32*warp ..)
#pragma unroll
  for (int i = 0; i < kSeg; ++i) o[i] = 0.f;

  if (warp == kConsumerWarps) {
    // ================================ producer warp =======================================
    if (kBulk && lane == 0) {
      for (int ci = issued; ci < total_chunks; ++ci) {
        const int slot = ci & (a.nch - 1);
        if (ci >= a.nch) mbar_spin_wait(&sh.empty[slot], ((ci >> a.nch_log2) - 1) & 1);
        MIS_ISSUE_CHUNK(slot, r_lo + ci * kChunkRows);
      }
    }
    __syncwarp();
  } else {
    Ctx c;
    c.a = &a;
    c.sh = &sh;
    c.v_w = v_w;
    c.h_w = h_w;
    c.tmp = tmp;
    c.ring = ring;
    c.gplane = a.src + (e0 - lp);
    c.tid = tid;
    c.lane = lane;
    c.warp = warp;
    c.nrows = nrows;
    c.lp = kBulk ? (P.left & 7) : lp;
    c.npairs = (c.lp + P.w + 1) >> 1;
    c.w = P.w;
    c.h = P.h;
    c.r_lo = r_lo;
    c.coff = kBulk ? (P.left & 7) : (use_is ? (wide8 ? (P.left & 7) : (P.left & 3)) : lp);   // tmp column of crop column 0
    c.gplane = a.src + (e0 - c.coff);
    c.ring_byte0 = 0;
    if (kBulk) {
      if (use_is) v_pass_is<kBulk, kWindow>(c, sched, nsrc);
      else switch (KV) {
        case 3: v_pass<3, kBulk, kWindow>(c, 0, nrows, 0); break;
        case 5: v_pass<5, kBulk, kWindow>(c, 0, nrows, 0); break;
        case 7: v_pass<7, kBulk, kWindow>(c, 0, nrows, 0); break;
        case 9: v_pass<9, kBulk, kWindow>(c, 0, nrows, 0); break;
        case 13: v_pass<13, kBulk, kWindow>(c, 0, nrows, 0); break;
        default: v_pass_dyn<kBulk, kWindow>(c, 0, nrows, 0); break;
      }
    } else {
      // non-TMA path: the band is processed as two 16-row sub-bands through a 16-row tile (34 KB instead of 68 KB,
      // which lets three CTAs share an SM); each sub-band = two 8-row cp.async streams of 128 threads (or the
      // output-stationary fallback), then its H pass.
      if (use_is) {   // zero the columns right of the crop that the unrolled H-pass taps may touch (weights are 0)
        const int c0z = 4 * ((c.coff + c.w + 3) >> 2), per = min(a.kstride + 2, a.pstr - c0z);
        for (int i = tid; i < 16 * per; i += kConsumerThreads) tmp[(i / per) * a.pstr + c0z + (i % per)] = 0.f;
      }
      const float post = kWindow ? 1.f : (1.f / 65535.f);
#pragma unroll
      for (int sb = 0; sb < 2; ++sb) {
        if (sb * 16 < nrows) {                                  // uniform
          const int yb16 = min(sb * 16 + 16, nrows);
          if (use_is) {
            const int st = sb * 2 + (tid >> 7), t = tid & 127;
            const int ya = st * rows_per_stream, yb = min(ya + rows_per_stream, nrows);
            if (ya < nrows)
              v_pass_cp<4, kWindow>(c, sched + st * cap_s, nsrc_s[st], sh.v_info[ya].x, ya, yb, sb * 16, ring, t);
          } else switch (KV) {
            case 3: v_pass<3, false, kWindow>(c, sb * 16, yb16, sb * 16); break;
            case 5: v_pass<5, false, kWindow>(c, sb * 16, yb16, sb * 16); break;
            case 7: v_pass<7, false, kWindow>(c, sb * 16, yb16, sb * 16); break;
            case 9: v_pass<9, false, kWindow>(c, sb * 16, yb16, sb * 16); break;
            case 13: v_pass<13, false, kWindow>(c, sb * 16, yb16, sb * 16); break;
            default: v_pass_dyn<false, kWindow>(c, sb * 16, yb16, sb * 16); break;
          }
          bar_sync(1, kConsumerThreads);
          if (warp * kSeg < s) {
            switch (KH) {
              case 3: h_pass16<3>(c, &o[16 * sb], post); break;
              case 5: h_pass16<5>(c, &o[16 * sb], post); break;
              case 7: h_pass16<7>(c, &o[16 * sb], post); break;
              case 9: h_pass16<9>(c, &o[16 * sb], post); break;
              case 13: h_pass16<13>(c, &o[16 * sb], post); break;
              default: h_pass16_dyn(c, &o[16 * sb], post); break;
            }
          }
          bar_sync(1, kConsumerThreads);                        // the tile is reused by the next sub-band
        }
      }
    }
    MIS_STAMP(2);   // V pass done (this warp)
    if (kBulk) bar_sync(1, kConsumerThreads);
    MIS_STAMP(3);   // all V warps done
    if (kBulk && warp * kSeg < s) {
      const float post = kWindow ? 1.f : (1.f / 65535.f);
      switch (KH) {
        case 3: h_pass<3>(c, o, post); break;
        case 5: h_pass<5>(c, o, post); break;
        case 7: h_pass<7>(c, o, post); break;
        case 9: h_pass<9>(c, o, post); break;
        case 13: h_pass<13>(c, o, post); break;
        default: h_pass_dyn(c, o, post); break;
      }
    }
  }

  MIS_STAMP(4);   // H pass done
  // ================================ colour ops =============================================
  cluster_wait_acquire();   // phase 1 done: all CTAs of the cluster are resident
  // element idx of o[] sits at output row y0 + elem_row(idx), column elem_x(idx):
  //   TMA path      : lane = row, one run of 32 columns per thread
  //   non-TMA path  : idx = 16*sub-band + i, lane = (row & 15, column half), two runs of 16 columns per thread
  const int x0 = warp * kSeg;
  auto elem_ok = [&](int idx) {
    const int row = kBulk ? lane : ((idx >> 4) * 16 + (lane & 15));
    const int x = kBulk ? (x0 + idx) : (x0 + 16 * (lane >> 4) + (idx & 15));
    return (warp < kConsumerWarps) && row < nrows && x < s;
  };
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
#pragma unroll
        for (int i = 0; i < kSeg; ++i)
          if (elem_ok(i)) part += o[i];
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

  MIS_STAMP(5);   // colour ops (incl. cluster reduction) done
  // ================================ normalise + store ======================================
  if (warp < kConsumerWarps && x0 < s) {
    const float mean = a.mean[chan], inv_std = a.inv_std[chan];
#pragma unroll
    for (int i = 0; i < kSeg; ++i) o[i] = (o[i] - mean) * inv_std;
    const bool flip = (P.flags & MIS_VIEW_FLIP) != 0;
    if (kBulk) {
      if (lane < nrows) store_run<kSeg>(o, a.out, ((size_t)plane * s + (y0 + lane)) * s, x0, s, flip, a.out_f32 != 0);
    } else {
#pragma unroll
      for (int sb = 0; sb < 2; ++sb) {
        const int row = sb * 16 + (lane & 15), xs = x0 + 16 * (lane >> 4);
        if (row < nrows && xs < s)
          store_run<16>(&o[16 * sb], a.out, ((size_t)plane * s + (y0 + row)) * s, xs, s, flip, a.out_f32 != 0);
      }
    }
  }
  MIS_STAMP(6);
#undef MIS_STAMP
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
// uint16 [rows, W] row-major, box [kChunkRows x box_w], no swizzle, out-of-bounds elements read as 0
static int make_map(CUtensorMap* map, const uint16_t* base, uint64_t W, uint64_t rows, uint32_t box_w) {
  EncodeTiledFn fn = encode_fn();
  MIS_REQUIRE(fn, MIS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {W, rows};
  const cuuint64_t strides[1] = {W * sizeof(uint16_t)};
  const cuuint32_t box[2] = {box_w, (cuuint32_t)kChunkRows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MIS_REQUIRE(r == CUDA_SUCCESS, MIS_ERR_CUDA, "cuTensorMapEncodeTiled(uint16, box %u) failed with CUresult %d", box_w, (int)r);
  return MIS_OK;
}

template <bool kBulk, bool kWindow>
static int launch(const Args& a, const CUtensorMap* maps, int grid, size_t smem, cudaStream_t stream) {
  auto* fn = &aug_kernel<kBulk, kWindow>;
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
  MIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, maps[0], maps[1], maps[2], maps[3], a));
  return MIS_OK;
}

}  // namespace aug
}  // namespace mis

using namespace mis;

#ifdef MIS_DEBUG
// profiling aid of -DMIS_DEBUG builds only (not part of the ABI): device buffer of [grid][8] int64 clock stamps, or NULL
static long long* g_dbg = nullptr;
extern "C" void mis_debug_set_stamp_buffer(void* p) { g_dbg = static_cast<long long*>(p); }
#endif

extern "C" int mis_aug_two_view(const uint16_t* src, int n_images, int C, int H, int W, int64_t img_stride,
                                const MisViewParams* params, int n_views, float win_lo, float win_hi,
                                const float* mean, const float* std, void* out, int s, int out_dtype, int use_tma,
                                void* stream) {
  return mis_aug_two_view_ordered(src, n_images, C, H, W, img_stride, params, n_views, nullptr, win_lo, win_hi, mean, std,
                                  out, s, out_dtype, use_tma, stream);
}

extern "C" int mis_aug_two_view_ordered(const uint16_t* src, int n_images, int C, int H, int W, int64_t img_stride,
                                        const MisViewParams* params, int n_views, const int32_t* view_order, float win_lo,
                                        float win_hi, const float* mean, const float* std, void* out, int s,
                                        int out_dtype, int use_tma, void* stream) {
  using namespace mis::aug;
  MIS_REQUIRE(src && params && mean && std && out, MIS_ERR_INVALID_ARG, "mis_aug_two_view: null pointer");
  MIS_REQUIRE(n_images > 0 && n_views >= 0 && H > 0 && W > 0, MIS_ERR_INVALID_ARG,
              "mis_aug_two_view: sizes must be positive (n_images=%d H=%d W=%d)", n_images, H, W);
  MIS_REQUIRE(out_dtype == MIS_DTYPE_BF16 || out_dtype == MIS_DTYPE_F32, MIS_ERR_INVALID_ARG,
              "mis_aug_two_view: out_dtype %d", out_dtype);
  MIS_REQUIRE(win_hi > win_lo, MIS_ERR_INVALID_ARG, "mis_aug_two_view: empty window [%g,%g]", win_lo, win_hi);
  MIS_REQUIRE(C == 1 || C == 3, MIS_ERR_UNSUPPORTED, "mis_aug_two_view: C=%d; 1 or 3 channels", C);
  MIS_REQUIRE(C == 1 || (use_tma == 0 && mis::augc::rgb_supported(s) && mis::augs::strip_supported(C, H, W, img_stride, s)),
              MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: 3-channel input runs on the strip kernel + colour kernel only: variant 0, crop a multiple "
              "of 8 up to 256, even W, at most 5.5x downscaling (got s=%d, H=%d, W=%d, variant %d)", s, H, W, use_tma);
  MIS_REQUIRE(s >= 8 && s <= kBandRows * kMaxBands, MIS_ERR_UNSUPPORTED, "mis_aug_two_view: crop size %d not in [8,256]", s);
  MIS_REQUIRE((W & 1) == 0 && (img_stride & 1) == 0, MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: W (%d) and img_stride must be even", W);
  MIS_REQUIRE(img_stride >= (int64_t)C * H * W, MIS_ERR_INVALID_ARG, "mis_aug_two_view: img_stride too small");
  MIS_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              MIS_ERR_INVALID_ARG, "mis_aug_two_view: src/out must be 16-byte aligned");
  for (int c = 0; c < C; ++c)
    MIS_REQUIRE(std[c] != 0.f, MIS_ERR_INVALID_ARG, "mis_aug_two_view: std[%d] == 0", c);
  if (n_views == 0) return MIS_OK;

  // variant 0 (default): strip kernel (aug_strip.cu) wherever it applies, else the round-1 warp-tile kernel (3), else
  // the cp.async band kernel (2); 1: TMA band kernel
  const bool window = !(win_lo == 0.f && win_hi == 65535.f);
  if (use_tma == 0 && mis::augs::strip_supported(C, H, W, img_stride, s)) {
    mis::augs::StripArgs t = {};
    t.src = src;
    t.img_stride = img_stride;
    t.C = C;
    t.H = H;
    t.W = W;
    t.params = params;
    t.order = view_order;
    t.win_lo = win_lo;
    t.win_scale = 1.0f / (win_hi - win_lo);
    for (int c = 0; c < C; ++c) {
      t.mean[c] = mean[c];
      t.inv_std[c] = 1.0f / std[c];
    }
    t.out = out;
    t.s = s;
    t.out_f32 = out_dtype == MIS_DTYPE_F32 ? 1 : 0;
    t.raw_all = C == 3 ? 1 : 0;
    if (int rc = mis::augs::launch_strip(t, n_views, window, reinterpret_cast<cudaStream_t>(stream))) return rc;
    if (C == 3)      // colour ops mix the channels: a second kernel over the three parked planes of every view
      return mis::augc::launch_rgb_color(out, t.out_f32, params, n_views, s, t.mean, t.inv_std, reinterpret_cast<cudaStream_t>(stream));
    return MIS_OK;
  }
  if ((use_tma == 0 || use_tma == 3) && mis::augt::tile_supported(C, H, W, img_stride, s)) {
    mis::augt::TileArgs t = {};
    t.src = src;
    t.img_stride = img_stride;
    t.C = C;
    t.H = H;
    t.W = W;
    t.params = params;
    t.win_lo = win_lo;
    t.win_scale = 1.0f / (win_hi - win_lo);
    for (int c = 0; c < C; ++c) {
      t.mean[c] = mean[c];
      t.inv_std[c] = 1.0f / std[c];
    }
    t.out = out;
    t.s = s;
    t.out_f32 = out_dtype == MIS_DTYPE_F32 ? 1 : 0;
    t.nbands = (s + kBandRows - 1) / kBandRows;
    return mis::augt::launch_tile(t, n_views, window, reinterpret_cast<cudaStream_t>(stream));
  }
  use_tma = (use_tma == 1) ? 1 : 0;

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
  a.out_f32 = out_dtype == MIS_DTYPE_F32 ? 1 : 0;
  // worst-case taps per axis: support = max(size/s, 1), K = 2*ceil(support) + 1
  auto kmax = [&](int n) { int sup = (n + s - 1) / s; if (sup < 1) sup = 1; return 2 * sup + 1; };
  const int kbound = kmax(H) > kmax(W) ? kmax(H) : kmax(W);
  const int kunroll = kbound <= 3 ? 3 : kbound <= 5 ? 5 : kbound <= 7 ? 7 : kbound <= 9 ? 9 : kbound <= 13 ? 13 : kbound;
  a.kstride = align_up(kunroll, 4);
  // ring: the widest read window (kunroll rows) plus the chunk being filled; power of two for cheap wrap
  // ring: 4 slots of 8 rows (32 rows in flight); the unrolled fallback's read window must fit next to the slot
  // being filled
  int nch = 4;
  while (nch * kChunkRows < kunroll + 2 * kChunkRows) nch *= 2;
  a.nch = nch;
  a.nch_log2 = 0;
  while ((1 << a.nch_log2) < nch) ++a.nch_log2;
  MIS_REQUIRE(a.nch <= kMaxChunks, MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: H/s = %d/%d needs a %d-chunk ring (max %d)", H, s, a.nch, kMaxChunks);
  // TMA staging needs one 16-byte phase for all rows of a crop (W % 8 == 0); otherwise plain loads
  // TMA staging views the batch as one [n_images*C*H, W] uint16 matrix: needs 16-byte row pitch and dense images
  const bool bulk = use_tma && (W % 8 == 0) && (img_stride == (int64_t)C * H * W);
  a.plane_rows = H;
  a.slot_bytes = ((W + kBoxCols - 1) / kBoxCols) * kBoxBytes;
  a.pstr = (W + 6 + a.kstride + 2) | 1;
  int off = align_up((int)sizeof(SmemHeader), 16);
  a.off_vw = off;
  off += align_up(kBandRows * a.kstride * 2 * 4, 16);
  a.off_hw = off;
  off += align_up(s * a.kstride * 4, 16);
  a.off_tmp = off;
  off += align_up((bulk ? kBandRows : kBandRows / 2) * a.pstr * 4, 128);   // transposition tile: band or 16-row sub-band
  // schedule capacity: one band-long schedule (TMA path) or four quarter-band schedules (cp.async path)
  {
    const int sup = (H + s - 1) / s;
    const bool bulk_path = use_tma && (W % 8 == 0) && (img_stride == (int64_t)C * H * W);
    a.rmax = bulk_path ? (kBandRows * sup + a.kstride + 4) : 4 * ((kBandRows / 4) * sup + a.kstride + 4);
  }
  a.off_sched = off;
  off += align_up(a.rmax * (int)sizeof(SchedRow), 128);
  off = align_up(off, 128);
  a.off_ring = off;
  if (bulk) off += a.nch * a.slot_bytes;
  else off += kCpDepth * kConsumerThreads * kCpSlotBytes;     // per-thread cp.async ring of the non-TMA path
  const size_t smem = (size_t)off;
  MIS_REQUIRE(smem <= 227 * 1024, MIS_ERR_UNSUPPORTED,
              "mis_aug_two_view: needs %zu B of shared memory per CTA (H=%d W=%d s=%d)", smem, H, W, s);

#ifdef MIS_DEBUG
  a.dbg = g_dbg;
#endif
  const int grid = a.nbands * n_views * C;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap maps[4] = {};
  if (bulk) {
    for (int i = 0; i < 4; ++i)
      if (int rc = make_map(&maps[i], src, (uint64_t)W, (uint64_t)n_images * C * H, 64u * (i + 1))) return rc;
    return window ? launch<true, true>(a, maps, grid, smem, st) : launch<true, false>(a, maps, grid, smem, st);
  }
  return window ? launch<false, true>(a, maps, grid, smem, st) : launch<false, false>(a, maps, grid, smem, st);
}

// Which K1 variant mis_aug_two_view runs for this shape and selector (0 strip, 3 warp-tile, 1 TMA band, 2 cp.async band).
extern "C" int mis_aug_kernel_variant(int C, int H, int W, int64_t img_stride, int s, int use_tma) {
  using namespace mis::aug;
  if (use_tma == 0 && mis::augs::strip_supported(C, H, W, img_stride, s)) return 0;
  if ((use_tma == 0 || use_tma == 3) && mis::augt::tile_supported(C, H, W, img_stride, s)) return 3;
  const bool bulk = use_tma == 1 && (W % 8 == 0) && (img_stride == (int64_t)C * H * W);
  return bulk ? 1 : 2;
}

extern "C" int64_t mis_aug_algorithmic_bytes(const MisViewParams* p, int n_views, int C, int s, int out_dtype) {
  if (!p || n_views < 0) return -1;
  const int64_t ob = out_dtype == MIS_DTYPE_F32 ? 4 : 2;
  int64_t total = 0;
  for (int v = 0; v < n_views; ++v) total += 2 * (int64_t)C * p[v].h * p[v].w + ob * C * s * s;
  return total;
}

// One host round trip per batch: check the table, copy it and its launch order into the caller's pinned block, one H2D
// copy, K1, and the blur kernel when a record asks for it (the Python side of a step is as long as its GPU side at small
// per-GPU batches, so every separate ctypes / torch call counts).
extern "C" int mis_aug_two_view_staged(const uint16_t* src, int n_images, int C, int H, int W, int64_t img_stride,
                                       const MisViewParams* params_host, int n_views, void* staging_pinned,
                                       void* staging_dev, int64_t staging_bytes, float win_lo, float win_hi,
                                       const float* mean, const float* std, void* out, int s, int out_dtype, int use_tma,
                                       uint32_t* flags_or, int* bad_index, int* n_launches, void* stream) {
  MIS_REQUIRE((params_host || n_views == 0) && staging_pinned && staging_dev && flags_or && bad_index && n_launches,
              MIS_ERR_INVALID_ARG, "mis_aug_two_view_staged: null pointer");
  MIS_REQUIRE(n_views >= 0 && staging_bytes >= (int64_t)n_views * (int64_t)(sizeof(MisViewParams) + 4), MIS_ERR_INVALID_ARG,
              "mis_aug_two_view_staged: staging block of %lld bytes for %d views", (long long)staging_bytes, n_views);
  *n_launches = 0;
  if (int rc = mis_view_params_check(params_host, n_views, n_images, H, W, flags_or, bad_index)) return rc;
  if (*bad_index >= 0 || n_views == 0) return MIS_OK;          // the caller reports the offending record
  const uint32_t extra = *flags_or & (MIS_VIEW_BLUR | MIS_VIEW_SOLARIZE);
  MIS_REQUIRE(!extra || mis_aug_kernel_variant(C, H, W, img_stride, s, use_tma) == 0, MIS_ERR_UNSUPPORTED,
              "GaussianBlur / RandomSolarize are fused into the strip kernel only (use_tma=0, 8 <= crop <= 256, at most "
              "5.5x downscaling); this call would run K1 variant %d", mis_aug_kernel_variant(C, H, W, img_stride, s, use_tma));
  const size_t rec_bytes = (size_t)n_views * sizeof(MisViewParams);
  uint8_t* hp = static_cast<uint8_t*>(staging_pinned);
  memcpy(hp, params_host, rec_bytes);
  if (int rc = mis_view_cost_order(params_host, n_views, reinterpret_cast<int32_t*>(hp + rec_bytes))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MIS_CUDA_TRY(cudaMemcpyAsync(staging_dev, staging_pinned, rec_bytes + (size_t)n_views * 4, cudaMemcpyHostToDevice, st));
  const MisViewParams* pd = static_cast<const MisViewParams*>(staging_dev);
  const int32_t* od = reinterpret_cast<const int32_t*>(static_cast<const uint8_t*>(staging_dev) + rec_bytes);
  if (int rc = mis_aug_two_view_ordered(src, n_images, C, H, W, img_stride, pd, n_views, od, win_lo, win_hi, mean, std, out,
                                        s, out_dtype, use_tma, stream))
    return rc;
  *n_launches = 1 + (C == 3 ? 1 : 0);
  if (*flags_or & MIS_VIEW_BLUR) {
    if (int rc = mis_aug_blur_views(out, out_dtype, pd, n_views, C, s, mean, std, stream)) return rc;
    ++*n_launches;
  }
  return MIS_OK;
}
