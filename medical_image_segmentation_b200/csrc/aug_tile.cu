// K1 (tile variant) -- fused two-view augmentation for 16-bit slices, warp-autonomous tiles (sm_100a).
//
// One thread-block CLUSTER per output view plane, one CTA per band of 32 output rows, and inside the CTA one WARP per
// 32 x 32 output tile.  After a short CTA prologue (vertical tap tables / schedule) a warp never
// synchronises with another warp until the contrast mean:
//
//   V pass   : lane = NL pairs of adjacent source columns (columns 2*lane + 64*i), read straight from global memory
//              with 4-byte loads (a warp reads 128 contiguous bytes per source row and pair), G rows in flight per lane
//              in register slots that are refilled as soon as they are consumed, after an L2 prefetch of the band;
//              input-stationary: every source pixel is loaded and converted once (exact magic-number u16 -> f32) and
//              scattered with packed FFMA2 into the <= 3 output rows whose window contains it; the accumulators rotate
//              through the FMA operands when an output row completes (a precomputed bit mask says when);
//   H pass   : a completed intermediate row (<= 192 floats) goes to a warp-private double-buffered row in shared
//              memory; lane = one output column, its tap weights live in REGISTERS for the whole tile, laid out
//              against the 16-byte-aligned window start so an output costs NS 16-byte shared loads + 2*NS FFMA2;
//              results are parked in a warp-private 32 x 32 fp32 tile (row pitch 36 floats);
//   colour   : 1/65535 and a brightness that precedes contrast are one saturating multiply on the fly; the contrast
//              mean over the whole view is reduced lane -> warp -> CTA -> cluster through distributed shared memory,
//              so each view is written exactly once;
//   store    : the tile is re-read 8 pixels per lane (conflict-free), contrast / brightness / normalise / flip,
//              16-byte stores.
//
// Classes (warp-uniform, from the strip's source span): (NL, NS) = (1,2) up to ~1.85x, (2,3) up to ~3x, (3,4) up to
// 5.5x downscaling.  Vertical upscaling uses an output-stationary three-tap pass, any window that would feed more than
// three output rows a generic output-stationary pass, both over the same H pass.  (A 64-column strip with two outputs
// per lane executes 16 % fewer instructions but needs ~118 registers: 16 warps/SM, measured 0.79 ms vs 0.51 ms.)
//
// Arithmetic restated from torchvision 0.26 / ATen (see oracle/aug_oracle.py):
//   taps   : _upsample_bilinear2d_aa (triangle filter, support = max(scale,1), weights normalised)
//   colour : functional/_color.py:114-125 (brightness), :190-205 + _blend :92-97 (contrast)
//   output : (x - mean) / std, functional/_misc.py:37-67
#include <cuda_bf16.h>

#include "aug_math.cuh"
#include "aug_tile.cuh"
#include "common.cuh"

namespace mis {
namespace augt {

using namespace mis::aug;

constexpr int kBand = 32;          // output rows per CTA
constexpr int kStrip = 32;         // output columns per warp
constexpr int kSchedCap = 200;     // source rows of one band (31*5.5 + window + group padding); <= 256 mask bits
constexpr int kVK = 16;            // widest vertical window kept in the weight table (2*ceil(5.5)+1 = 13)
constexpr int kRowBuf = 208;       // floats per intermediate-row buffer: 3*64 columns + slack for the aligned H reads
constexpr int kOPitch = kStrip + 4; // output-tile row pitch in floats: 16-byte aligned rows, conflict-free 8-pixel reads
constexpr int kMaxWarps = 256 / kStrip;   // s <= 256

// Input-stationary schedule entry of one source row: it feeds output rows first .. first+2 of the band with the
// (duplicated, FFMA2-ready) weights w[0..2].
struct __align__(16) Sched {
  float w[3][2];
  int first;
  int pad;
};

struct Smem {
  Sched sched[kSchedCap];
  float vw[kBand][kVK];     // normalised vertical weights of the band's output rows
  int2 vinfo[kBand];        // {first source row, taps}
  float part[8];            // per-CTA partial sums of the contrast mean (written by the cluster peers)
  float red[kMaxWarps];
  uint32_t fmask[8];        // bit rr: source row rr completes an output row (the one before its first open row)
  int m_max;                // most output rows any single source row of this band feeds (4 = schedule unusable)
  int kv_max;               // widest vertical window of the band
  int pad[2];
};
struct WarpSmem {
  float row[2][kRowBuf];
  float o[kBand * kOPitch];
};
static_assert(sizeof(Smem) % 16 == 0 && sizeof(WarpSmem) % 16 == 0, "16-byte aligned shared-memory blocks");

// what a warp needs to know about its strip of 32 output columns
struct Tile {
  float win_lo, win_scale;
  Smem* sh;
  WarpSmem* ws;
  const uint16_t* crop;    // crop (0, 0) of this plane in global memory
  int W;                   // source row pitch in elements
  int ca;                  // first staged source column (crop coordinates, 4-byte aligned address)
  int lane, nthreads;
  int nrows;
  int span;                // source columns this warp stages, from `ca`
  int hlo, hsize;          // this lane's output: first source column, taps
  float hctr;
  float hinv;
  float out_scale;         // 1/65535, times the brightness factor when brightness precedes contrast
  bool down;               // vertical downscaling and the schedule fits
};

template <bool kWindow>
__device__ __forceinline__ uint64_t conv_px(uint32_t p, uint64_t wsc, uint64_t wof) {
  uint64_t f = u16x2_to_f32x2(p);
  if (kWindow) {
    f = ffma2(f, wsc, wof);
    float f0, f1;
    unpack2(f, f0, f1);
    f = pack2(fminf(fmaxf(f0, 0.f), 1.f), fminf(fmaxf(f1, 0.f), 1.f));
  }
  return f;
}
__device__ __forceinline__ uint64_t ptr_add(uint64_t p, uint32_t bytes) {     // one IMAD.WIDE instead of an add pair
  uint64_t r;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(bytes), "l"(p));
  return r;
}

// One strip of one warp: 32 output columns, all rows of the band.  Returns this lane's share of the pixel sum.
template <int NL, int NS, bool kWindow>
__device__ __forceinline__ float run_tile(const Tile& t) {
  Smem& sh = *t.sh;
  WarpSmem& ws = *t.ws;
  const int lane = t.lane;

  // ---- this lane's horizontal taps, in registers, aligned to the 16-byte window start ------------------
  uint64_t hw[2 * NS];
  uint32_t rbase;
  {
    const int off = (t.hlo - t.ca) & 3;
    const int xa = t.hsize > 0 ? ((t.hlo - t.ca) & ~3) : 0;
    // taps at their aligned positions (zero outside the window); their sum in ascending order is the reference's total
    float w[4 * NS];
    float tot = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4 * NS; ++jj) {
      const int j = jj - off;
      w[jj] = (j >= 0 && j < t.hsize) ? aa_tri(j + t.hlo, t.hctr, t.hinv) : 0.f;
      tot += w[jj];
    }
    // (one reciprocal instead of a division per tap: weights differ from w / total by at most one ulp)
    const float rtot = tot != 0.f ? __frcp_rn(tot) : 1.f;
#pragma unroll
    for (int i = 0; i < 2 * NS; ++i) hw[i] = pack2(w[2 * i] * rtot, w[2 * i + 1] * rtot);
    rbase = opaque(smem_u32(&ws.row[0][0]) + 4u * xa);             // aligned window start in buffer 0
  }
  // zero-weight taps may read up to 15 columns behind the staged span: those must hold finite values
  if (lane < 24) {
    const int c = (t.span & ~1) + lane;
    if (c < kRowBuf) ws.row[0][c] = ws.row[1][c] = 0.f;
  }
  __syncwarp();

  bool act[NL];
  uint32_t wb[NL];
  uint64_t gl[NL];
  const int64_t rowe = t.W;
  // (idle lanes read a valid address and park their finite garbage in the zero-weight slack behind the span)
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    act[i] = (2 * lane + 64 * i) < t.span;
    wb[i] = opaque(smem_u32(&ws.row[0][0]) + (act[i] ? 8u * lane + 256u * i : 4u * (kRowBuf - 8)));
  }
  const uint64_t wsc = pack2(t.win_scale, t.win_scale);
  const uint64_t wof = pack2(-t.win_lo * t.win_scale, -t.win_lo * t.win_scale);
  const uint32_t rowb = 2u * (uint32_t)t.W;                          // source row pitch in bytes

  const int r_lo = sh.vinfo[0].x;
  const int nsrc = sh.vinfo[t.nrows - 1].x + sh.vinfo[t.nrows - 1].y - r_lo;
#pragma unroll
  for (int i = 0; i < NL; ++i)
    gl[i] = reinterpret_cast<uint64_t>(t.crop + (int64_t)r_lo * rowe + t.ca + (act[i] ? 2 * lane + 64 * i : 0));

  // G register slots per lane hold the next G source rows of the input-stationary stream; issued before the
  // schedule barrier so their latency overlaps it
  constexpr int G = NL == 1 ? 8 : 4;
  uint32_t p[G][NL];
  uint64_t gq[NL];                                  // next row to fetch, this lane's columns
  if (t.down) {
#pragma unroll
    for (int i = 0; i < NL; ++i) {
#pragma unroll
      for (int k = 0; k < G; ++k) p[k][i] = ldg_nc_u32(gl[i] + (uint64_t)((uint32_t)min(k, nsrc - 1) * rowb));
      gq[i] = gl[i] + (uint64_t)((uint32_t)G * rowb);
    }
  }
  bar_sync(1, t.nthreads);                                  // the schedule (built by the CTA's first threads) is complete
  const bool use_is = t.down && sh.m_max <= 3;

  uint64_t A[NL], B[NL], Cc[NL];    // the three open output rows of this lane's column pair(s)
#pragma unroll
  for (int i = 0; i < NL; ++i) A[i] = B[i] = Cc[i] = 0ull;
  float sum = 0.f;
  // running shared-memory cursors (32-bit shared addresses): the intermediate row alternates between two buffers,
  // the output tile advances one row per completed output row
  uint32_t sel = 0;                                                    // 0 / kRowBuf * 4: buffer in use
  const uint32_t op0 = smem_u32(ws.o) + 4u * lane;
  uint32_t op = opaque(op0);

  // the oldest open output row is complete in A: H pass over the intermediate row (first half: park the row, read the
  // windows, accumulate; second half: scale, clamp, sum, park the results in the tile)
  uint64_t hacc;
  auto hrow_a = [&]() {
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      float v0, v1;
      unpack2(A[i], v0, v1);
      sts64(wb[i] + sel, v0, v1);
    }
    __syncwarp();
    uint64_t acc = 0ull;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const float4 v = lds128(rbase + sel + 16u * j);
      acc = ffma2(pack2(v.x, v.y), hw[2 * j], acc);
      acc = ffma2(pack2(v.z, v.w), hw[2 * j + 1], acc);
    }
    hacc = acc;
  };
  auto hrow_b = [&]() {
    float lo, hi;
    unpack2(hacc, lo, hi);
    // 1/65535 (and a brightness factor that precedes contrast) in one saturating multiply; without brightness the
    // clamp only trims the one-ulp overshoot a normalised filter can produce
    const float val = __saturatef((lo + hi) * t.out_scale);
    sum += val;                                   // lanes beyond the view's last column carry zero weights
    sts32(op, val);
    op += kOPitch * 4;
    sel ^= (uint32_t)(kRowBuf * 4);
  };
  auto hrow = [&]() {
    hrow_a();
    hrow_b();
  };

  if (use_is) {
    // ---- input-stationary stream over the band's source rows --------------------------------------------
    // A slot is refilled (row + G) as soon as it is consumed, so the loads stay G rows ahead without a second
    // buffer.  Whether a source row completes an output row comes from a precomputed bit mask (the branch does not
    // wait for a shared-memory load); the schedule guarantees at most one completed output row per source row.
    const Sched* sp = sh.sched;
#pragma unroll 1
    for (int rr0 = 0; rr0 < nsrc; rr0 += G) {
      const uint32_t m = sh.fmask[rr0 >> 5] >> (rr0 & 31);
      const int rem = nsrc - G - rr0;                 // slot k is refilled while k < rem
#pragma unroll
      for (int k = 0; k < G; ++k) {
        uint64_t f[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) {
          f[i] = conv_px<kWindow>(p[k][i], wsc, wof);
          if (k < rem) p[k][i] = ldg_nc_u32(gq[i]);
          gq[i] = ptr_add(gq[i], rowb);
        }
        if (m & (1u << k)) {                            // warp-uniform: the oldest open output row is complete
          // the row's weights are requested between the two halves of the H pass: not live across its window loads,
          // yet their latency hides behind its tail
          hrow_a();
          const float4 s0 = *reinterpret_cast<const float4*>(&sp[k].w[0][0]);
          const float2 s1 = *reinterpret_cast<const float2*>(&sp[k].w[2][0]);
          hrow_b();
          const uint64_t w0 = pack2(s0.x, s0.y), w1 = pack2(s0.z, s0.w), w2 = pack2(s1.x, s1.y);
#pragma unroll
          for (int i = 0; i < NL; ++i) {                // rotate the accumulators through the FMA operands
            A[i] = ffma2(f[i], w0, B[i]);
            B[i] = ffma2(f[i], w1, Cc[i]);
            Cc[i] = ffma2(f[i], w2, 0ull);
          }
        } else {
          const float4 s0 = *reinterpret_cast<const float4*>(&sp[k].w[0][0]);
          const float2 s1 = *reinterpret_cast<const float2*>(&sp[k].w[2][0]);
          const uint64_t w0 = pack2(s0.x, s0.y), w1 = pack2(s0.z, s0.w), w2 = pack2(s1.x, s1.y);
#pragma unroll
          for (int i = 0; i < NL; ++i) {
            A[i] = ffma2(f[i], w0, A[i]);
            B[i] = ffma2(f[i], w1, B[i]);
            Cc[i] = ffma2(f[i], w2, Cc[i]);
          }
        }
      }
      sp += G;
    }
    const uint32_t op_end = op0 + (uint32_t)t.nrows * (kOPitch * 4);
#pragma unroll 1
    while (op != op_end) {
      hrow();
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        A[i] = B[i];
        B[i] = Cc[i];
        Cc[i] = 0ull;
      }
    }
  } else if (sh.kv_max <= 3) {
    // ---- output-stationary, three taps (vertical upscaling): the taps of the next output row are fetched while
    // the current one is accumulated
    uint32_t q[3][NL];
    auto fetch = [&](int y) {
      const int2 info = sh.vinfo[y];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint32_t rr = (uint32_t)min(info.x - r_lo + k, nsrc - 1);
#pragma unroll
        for (int i = 0; i < NL; ++i) q[k][i] = ldg_nc_u32(gl[i] + (uint64_t)(rr * rowb));
      }
    };
    fetch(0);
#pragma unroll 1
    for (int y = 0; y < t.nrows; ++y) {
      uint64_t f[3][NL];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < NL; ++i) f[k][i] = conv_px<kWindow>(q[k][i], wsc, wof);
      const float4 wv = *reinterpret_cast<const float4*>(sh.vw[y]);      // zero beyond the window
      if (y + 1 < t.nrows) fetch(y + 1);
      const uint64_t w0 = pack2(wv.x, wv.x), w1 = pack2(wv.y, wv.y), w2 = pack2(wv.z, wv.z);
#pragma unroll
      for (int i = 0; i < NL; ++i) A[i] = ffma2(f[2][i], w2, ffma2(f[1][i], w1, ffma2(f[0][i], w0, 0ull)));
      hrow();
    }
  } else {
    // ---- output-stationary, any window (a source row feeding more than three output rows) --------------------
#pragma unroll 1
    for (int y = 0; y < t.nrows; ++y) {
      const int2 info = sh.vinfo[y];
      const float* wv = sh.vw[y];
#pragma unroll
      for (int i = 0; i < NL; ++i) A[i] = 0ull;
#pragma unroll 1
      for (int k = 0; k < info.y; ++k) {
        const float w = wv[k];
        const uint64_t wp = pack2(w, w);
        const uint64_t ro = (uint64_t)((uint32_t)(info.x - r_lo + k) * rowb);
#pragma unroll
        for (int i = 0; i < NL; ++i) A[i] = ffma2(conv_px<kWindow>(ldg_nc_u32(gl[i] + ro), wsc, wof), wp, A[i]);
      }
      hrow();
    }
  }
  return sum;
}

template <bool kWindow>
__global__ void __maxnreg__(64) aug_tile_kernel(const __grid_constant__ TileArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  Smem& sh = *reinterpret_cast<Smem*>(smem);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int nthreads = blockDim.x;
  WarpSmem& ws = *reinterpret_cast<WarpSmem*>(smem + sizeof(Smem) + (size_t)warp * sizeof(WarpSmem));
  const int band = blockIdx.x;                  // grid = (bands, planes): band == rank in cluster, no integer divisions
  const int plane = blockIdx.y;                 // view * C + c
  const int view = a.C == 1 ? plane : plane / a.C;
  const int chan = plane - view * a.C;
  const int s = a.s;

  cluster_arrive_relaxed();   // phase 1: "every CTA of the cluster is running" (waited before DSMEM use)

  const MisViewParams P = a.params[view];
  const int y0 = band * kBand;
  const int nrows = min(kBand, s - y0);
  const int64_t plane_base = (int64_t)P.img * a.img_stride + (int64_t)chan * a.H * a.W;
  const int64_t e0 = plane_base + (int64_t)P.top * a.W + P.left;   // element index of crop (0,0)
  const float vscale = (float)P.h / (float)s;
  const float hscale = (float)P.w / (float)s;
  const float vsup = vscale >= 1.f ? vscale : 1.f, vinv = vscale >= 1.f ? 1.f / vscale : 1.f;
  const float hsup = hscale >= 1.f ? hscale : 1.f, hinv = hscale >= 1.f ? 1.f / hscale : 1.f;

  // ---- the band's source rows ---------------------------------------------------------------------------------
  int lo0, hi0, lo1, hi1;
  {
    float ctr;
    aa_window(y0, P.h, vscale, vsup, lo0, hi0, ctr);
    aa_window(y0 + nrows - 1, P.h, vscale, vsup, lo1, hi1, ctr);
  }
  // ---- pull the band's crop rows into L2 (every thread derives the row range itself) --------------------------
  {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a.src + e0);
    const int head = (int)(reinterpret_cast<uintptr_t>(base) & 127);
    const int lines = (head + 2 * P.w + 127) >> 7;               // 128-byte lines per crop row
    if (lane < lines)                                            // one source row per warp and step, one line per lane
      for (int r = lo0 + warp; r < hi1; r += (nthreads >> 5))
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)r * a.W * 2 - head + lane * 128));
  }

  // ---- this warp's strip: horizontal windows and class (independent of the tables above) --------------------
  Tile t;
  t.win_lo = a.win_lo;
  t.win_scale = a.win_scale;
  t.sh = &sh;
  t.ws = &ws;
  t.crop = a.src + e0;
  t.W = a.W;
  t.lane = lane;
  t.nthreads = nthreads;
  t.nrows = nrows;
  t.hinv = hinv;
  const bool jitter = (P.flags & MIS_VIEW_JITTER) != 0;
  // brightness (op 0) before contrast (op 1) is applied while the tile is produced: the contrast mean needs it
  int pos_b = 0, pos_c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (P.order[k] == 0) pos_b = k;
    if (P.order[k] == 1) pos_c = k;
  }
  const float post = kWindow ? 1.f : (1.f / 65535.f);
  t.out_scale = opaque((jitter && pos_b < pos_c) ? post * P.brightness : post);
  const bool has_post = jitter && pos_b > pos_c;

  const int x0 = warp * kStrip;
  {
    const int x = x0 + lane;
    int lo = 0, hi = 0;
    float ctr = 0.f;
    if (x < s) aa_window(x, P.w, hscale, hsup, lo, hi, ctr);
    const int kcap = 2 * (int)ceilf(hsup) + 1;
    int size = hi - lo;
    size = size < 0 ? 0 : (size > kcap ? kcap : size);
    t.hlo = lo;
    t.hsize = (x < s) ? size : 0;
    t.hctr = ctr;
  }
  // ---- vertical tap tables of the band: warp 0, one lane per output row.  First thing after the parameters: every
  // other warp needs them at the first barrier and has the prefetch and its strip windows to do meanwhile ---------
  if (tid == 0) sh.m_max = 0;
  if (tid < 8) sh.fmask[tid] = 0u;
  if (warp == 0) {
    int size = 0;
    if (tid < nrows) {
      int lo, hi;
      float ctr;
      aa_window(y0 + tid, P.h, vscale, vsup, lo, hi, ctr);
      const int kcap = min(2 * (int)ceilf(vsup) + 1, kVK);
      size = hi - lo;
      size = size < 0 ? 0 : (size > kcap ? kcap : size);
      float4* w4 = reinterpret_cast<float4*>(sh.vw[tid]);
      if (size <= 8) {                                  // the usual case: taps in registers, no loops, two 16-byte stores
        float wj[8];
        float total = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          wj[j] = (j < size) ? aa_tri(j + lo, ctr, vinv) : 0.f;
          total += wj[j];                               // same order as the reference (zeros beyond the window)
        }
        const float rtot = total != 0.f ? __frcp_rn(total) : 1.f;
        w4[0] = make_float4(wj[0] * rtot, wj[1] * rtot, wj[2] * rtot, wj[3] * rtot);
        w4[1] = make_float4(wj[4] * rtot, wj[5] * rtot, wj[6] * rtot, wj[7] * rtot);
        w4[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        w4[3] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        float* w = sh.vw[tid];
        float total = 0.f;
        for (int j = 0; j < size; ++j) {
          const float t0 = aa_tri(j + lo, ctr, vinv);
          w[j] = t0;
          total += t0;
        }
        if (total != 0.f) {
          const float rtot = __frcp_rn(total);
          for (int j = 0; j < size; ++j) w[j] = w[j] * rtot;
        }
        for (int j = size; j < kVK; ++j) w[j] = 0.f;
      }
      sh.vinfo[tid] = make_int2(lo, size);
    }
    size = __reduce_max_sync(0xffffffffu, size);
    if (lane == 0) sh.kv_max = size;
  }

  __syncthreads();                                  // tables of the band are complete
  const int r_lo = sh.vinfo[0].x;
  const int nsrc = sh.vinfo[nrows - 1].x + sh.vinfo[nrows - 1].y - r_lo;
  t.down = (vscale >= 1.f) && (nsrc + 8 <= kSchedCap) && (nsrc > 0);
  if (t.down) {
    // input-stationary schedule (windows are monotone in y): one entry per source row, padded to a whole group.
    // Source row r can only lie in the windows of the output rows around (r + 0.5) / vscale - 0.5.
    const int npad = (nsrc + 7) & ~7;
    const float inv_vs = 1.f / vscale;
    for (int rr = tid; rr < npad; rr += nthreads) {
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;
      int first = 0;
      if (rr < nsrc) {
        const int r = r_lo + rr;
        const int yc = (int)floorf(((float)r + 0.5f) * inv_vs - 0.5f) - y0;
        int last = -1, prev = nrows;
        first = nrows;
#pragma unroll
        for (int d = -3; d <= 3; ++d) {
          const int yy = yc + d;
          if (yy >= 0 && yy < nrows) {
            const int2 info = sh.vinfo[yy];
            if (info.x <= r && r < info.x + info.y) {
              first = min(first, yy);
              last = max(last, yy);
            }
            if (info.x <= r - 1 && r - 1 < info.x + info.y) prev = min(prev, yy);
          }
        }
        // the stream handles at most three open output rows and completes at most ONE output row per source row;
        // anything else (never seen for downscaling windows) sends the band to the output-stationary pass
        int m = last - first + 1;
        if (last < 0 || (rr == 0 ? (first != 0) : (prev >= nrows || first - prev > 1 || first < prev))) {
          m = 4;
          first = 0;
          last = -1;
        }
        if (first <= last) w0 = sh.vw[first][r - sh.vinfo[first].x];
        if (first + 1 <= last) w1 = sh.vw[first + 1][r - sh.vinfo[first + 1].x];
        if (first + 2 <= last) w2 = sh.vw[first + 2][r - sh.vinfo[first + 2].x];
        atomicMax(&sh.m_max, m);
        if (rr > 0 && first > prev) atomicOr(&sh.fmask[rr >> 5], 1u << (rr & 31));
      }
      Sched e;
      e.w[0][0] = e.w[0][1] = w0;
      e.w[1][0] = e.w[1][1] = w1;
      e.w[2][0] = e.w[2][1] = w2;
      e.first = first;
      e.pad = 0;
      sh.sched[rr] = e;
    }
  }

  // ---- the tile: staged span (from the 4-byte aligned column at or before the strip's first window), class dispatch
  // (the schedule barrier sits inside, behind the per-warp set-up) ------------------------------------------------
  float sum;
  {
    const int c_lo = __shfl_sync(0xffffffffu, t.hlo, 0);
    t.ca = c_lo - (int)((e0 + c_lo) & 1);
    t.span = __reduce_max_sync(0xffffffffu, t.hsize > 0 ? t.hlo + t.hsize : 0) - t.ca;
    const int need = __reduce_max_sync(0xffffffffu, t.hsize > 0 ? ((t.hlo - t.ca) & 3) + t.hsize : 0);
    if (t.span <= 64 && need <= 8) sum = run_tile<1, 2, kWindow>(t);
    else if (t.span <= 128 && need <= 12) sum = run_tile<2, 3, kWindow>(t);
    else sum = run_tile<3, 4, kWindow>(t);
  }

  // ================================ contrast mean over the view ===========================================
  cluster_wait_acquire();   // phase 1 done: all CTAs of the cluster are resident
  float cadd = 0.f;
  const float cf = P.contrast;
  if (jitter) {
    sum = warp_sum(sum);
    if (lane == 0) sh.red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int i = 0; i < (nthreads >> 5); ++i) tot += sh.red[i];
      for (int r = 0; r < a.nbands; ++r) st_cluster_f32(&sh.part[band], (uint32_t)r, tot);
      fence_acq_rel_cluster();          // the writer releases; everybody else arrives relaxed (no CTA-wide membar)
    }
    cluster_arrive_relaxed();
    cluster_wait_acquire();
    float tot = 0.f;
    for (int r = 0; r < a.nbands; ++r) tot += sh.part[r];
    const float mu = tot / (float)(s * s);
    cadd = mu * (1.f - cf);
  } else {
    __syncwarp();
  }

  // ================================ colour, normalise, store ==============================================
  const float mean = a.mean[chan], inv_std = a.inv_std[chan];
  const bool flip = (P.flags & MIS_VIEW_FLIP) != 0;
  const float pb = P.brightness;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int id = it * 32 + lane;
    const int row = id >> 2, xs = x0 + 8 * (id & 3);
    if (row < nrows && xs < s) {
      const float* src = ws.o + row * kOPitch + 8 * (id & 3);
      const float4 v0 = *reinterpret_cast<const float4*>(src);
      const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
      float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      if (jitter) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __saturatef(fmaf(v[i], cf, cadd));
        if (has_post) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __saturatef(v[i] * pb);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * inv_std;
      store_run<8>(v, a.out, ((size_t)plane * s + (y0 + row)) * s, xs, s, flip, a.out_f32 != 0);
    }
  }
}

bool tile_supported(int C, int H, int W, int64_t img_stride, int s) {
  // class (3,4) covers 5.5x downscaling per axis: 31*5.5 + 2*5.5 + 2 <= 192 staged columns, 3 + 13 <= 16 aligned
  // taps, 31*5.5 + 13 + 8 <= kSchedCap schedule rows
  return C == 1 && s >= 8 && s <= kStrip * kMaxWarps && (W & 1) == 0 && (img_stride & 1) == 0 && 2 * W <= 11 * s &&
         2 * H <= 11 * s;
}

int launch_tile(const TileArgs& a, int n_views, bool window, cudaStream_t stream) {
  const int nxw = (a.s + kStrip - 1) / kStrip;
  const size_t smem = sizeof(Smem) + (size_t)nxw * sizeof(WarpSmem);
  auto* fn = window ? &aug_tile_kernel<true> : &aug_tile_kernel<false>;
  MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // grid = (bands, planes): gridDim.y holds at most 65535 planes, larger batches go out in several launches
  const int max_views = 65535 / a.C;
  const size_t esize = a.out_f32 ? 4 : 2;
  for (int v0 = 0; v0 < n_views; v0 += max_views) {
    const int nv = n_views - v0 < max_views ? n_views - v0 : max_views;
    TileArgs b = a;
    b.params = a.params + v0;
    b.out = static_cast<uint8_t*>(a.out) + (size_t)v0 * a.C * a.s * a.s * esize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.nbands, (unsigned)(nv * a.C));
    cfg.blockDim = dim3((unsigned)(32 * nxw));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.nbands;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, b));
  }
  return MIS_OK;
}

}  // namespace augt
}  // namespace mis
