// K1 (tile variant) -- fused two-view augmentation for 16-bit slices, warp-autonomous tiles (sm_100a).
//
// One thread-block CLUSTER per output view plane, one CTA per band of 32 output rows, and inside the CTA one WARP per
// 32 x 32 output tile.  After a short CTA prologue (vertical tap tables / schedule, one barrier) a warp never
// synchronises with another warp until the contrast mean:
//
//   V pass   : lane = two adjacent source columns (x NL when the crop is wide), read straight from global memory with
//              4-byte loads (a warp reads 128 contiguous bytes per source row), G rows in flight per lane in registers
//              after an L2 prefetch of the band; input-stationary: every source pixel is loaded and converted once
//              (exact magic-number u16 -> f32) and scattered with packed FFMA2 into the <= 3 output rows whose window
//              contains it; the three accumulators rotate when an output row completes;
//   H pass   : a completed intermediate row (<= 192 floats) goes to a warp-private double-buffered row in shared
//              memory; lane = one output column, its tap weights live in REGISTERS for the whole tile, laid out
//              against the 16-byte-aligned window start so a row costs NS 16-byte shared loads + 2*NS FFMA2;
//              the result is parked in a warp-private 32 x 32 fp32 tile (row pitch 36 floats);
//   colour   : brightness before contrast is applied on the fly; the contrast mean over the whole view is reduced
//              lane -> warp -> CTA -> cluster through distributed shared memory, so each view is written exactly once;
//   store    : the tile is re-read 8 pixels per lane (conflict-free), contrast / brightness / normalise / flip,
//              16-byte stores.
//
// Classes (warp-uniform, chosen from the crop's horizontal scale): (NL, NS) = (1,2) up to ~1.85x, (2,3) up to 3x,
// (3,4) up to 5.5x downscaling.  Vertical upscaling (and any window that would feed more than three output rows)
// uses an output-stationary V pass over the same H pass.
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
constexpr int kSchedCap = 200;     // source rows of one band (31*5.5 + window + group padding); <= 256 mask bits
constexpr int kVK = 16;            // widest vertical window kept in the weight table (2*ceil(5.5)+1 = 13)
constexpr int kRowBuf = 208;       // floats per intermediate-row buffer: 3*64 columns + slack for the aligned H reads
constexpr int kOPitch = 36;        // output-tile row pitch in floats: 16-byte aligned rows, conflict-free 8-pixel reads
constexpr int kMaxWarps = 8;       // s <= 256

// Input-stationary schedule entry of one source row: it feeds output rows first .. first+2 of the band with the
// (duplicated, FFMA2-ready) weights w[0..2]; rows before `first` are complete when it is reached.
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
  int m_max;                // most output rows any single source row of this band feeds
  int pad[3];
};
struct WarpSmem {
  float row[2][kRowBuf];
  float o[kBand * kOPitch];
};
static_assert(sizeof(Smem) % 16 == 0 && sizeof(WarpSmem) % 16 == 0, "16-byte aligned shared-memory blocks");

struct Tile {
  float win_lo, win_scale;
  Smem* sh;
  WarpSmem* ws;
  const uint32_t* g0;      // this lane's first column pair of source row r_lo (crop coordinates)
  int wq;                  // source row pitch in 4-byte words
  int lane;
  int nrows, nsrc, r_lo;
  int span;                // source columns this warp stages (from the even-aligned column `ca`)
  int xa;                  // float index of this lane's 16-byte-aligned window start inside the row buffer
  int off, hlo, hsize;     // window start relative to xa, first source column, taps
  float hctr, hinv;
  bool valid, use_is;
  float post, pre_b;
  bool has_pre;
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

// The whole tile of one warp: returns this lane's share of the tile's pixel sum (for the contrast mean).
template <int NL, int NS, bool kWindow>
__device__ __forceinline__ float run_tile(const Tile& t) {
  Smem& sh = *t.sh;
  WarpSmem& ws = *t.ws;
  const int lane = t.lane;

  // ---- this lane's horizontal taps, in registers, aligned to the 16-byte window start ------------------
  uint64_t hw[2 * NS];
  {
    float tot = 0.f;
    for (int j = 0; j < t.hsize; ++j) tot += aa_tri(j + t.hlo, t.hctr, t.hinv);
    float w[4 * NS];
#pragma unroll
    for (int jj = 0; jj < 4 * NS; ++jj) {
      const int j = jj - t.off;
      const float wj = aa_tri(j + t.hlo, t.hctr, t.hinv);
      w[jj] = (j >= 0 && j < t.hsize) ? (tot != 0.f ? wj / tot : wj) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 2 * NS; ++i) hw[i] = pack2(w[2 * i], w[2 * i + 1]);
  }
  // the row buffers must hold finite values where zero-weight taps may read
  if (lane < 24) {
    const int c = (t.span & ~1) + lane;             // columns at and behind the staged span (active lanes write < span + 1)
    if (c < kRowBuf) ws.row[0][c] = ws.row[1][c] = 0.f;
  }
  __syncwarp();

  bool act[NL];
  int lofs[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    act[i] = (2 * lane + 64 * i) < t.span;
    lofs[i] = act[i] ? lane + 32 * i : 0;      // idle lanes read a valid address and never store
  }
  const uint64_t wsc = pack2(t.win_scale, t.win_scale);
  const uint64_t wof = pack2(-t.win_lo * t.win_scale, -t.win_lo * t.win_scale);

  uint64_t A[NL], B[NL], Cc[NL];    // output rows ycur, ycur+1, ycur+2 of this lane's column pair(s)
#pragma unroll
  for (int i = 0; i < NL; ++i) A[i] = B[i] = Cc[i] = 0ull;
  float sum = 0.f;
  const float vscale_out = t.has_pre ? t.post * t.pre_b : t.post;
  // running shared-memory cursors (32-bit shared addresses): the intermediate row alternates between two buffers,
  // the output tile advances one row per completed output row
  // (idle lanes park their finite garbage in the zero-weight slack columns behind the staged span)
  uint32_t wb[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i)
    wb[i] = opaque(smem_u32(&ws.row[0][0]) + (act[i] ? 8u * lane + 256u * i : 4u * (kRowBuf - 8)));
  const uint32_t rbase = opaque(smem_u32(&ws.row[0][0]) + 4u * t.xa);  // aligned window start
  uint32_t sel = 0;                                                    // 0 / kRowBuf * 4: buffer in use
  uint32_t op = opaque(smem_u32(ws.o) + 4u * lane);

  // output row ycur is complete in A: H pass over the intermediate row, park the result in the output tile
  auto hrow = [&]() {
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
    float lo, hi;
    unpack2(acc, lo, hi);
    // 1/65535 (and a brightness factor that precedes contrast) in one saturating multiply; without brightness the
    // clamp only trims the one-ulp overshoot a normalised filter can produce
    const float val = __saturatef((lo + hi) * vscale_out);
    sum += val;                                   // lanes beyond the view's last column carry zero weights
    sts32(op, val);
    op += kOPitch * 4;
    sel ^= (uint32_t)(kRowBuf * 4);
  };

  if (t.use_is) {
    // ---- input-stationary stream over the band's source rows, G rows in flight per lane ------------------
    // (the schedule guarantees at most one completed output row per source row)
    // G register slots per lane hold the next G source rows; a slot is refilled (row + G) as soon as it is consumed,
    // so the loads stay G rows ahead without a second buffer.  Whether a source row completes an output row comes
    // from a precomputed bit mask (no dependence of the branch on a shared-memory load).
    constexpr int G = NL == 1 ? 8 : 4;
    const int nsrc = t.nsrc;
    const uint32_t rowb = 4u * (uint32_t)t.wq;          // source row pitch in bytes
    uint64_t gq[NL];                                  // next row to fetch, this lane's columns
    uint32_t p[G][NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const uint64_t gl = reinterpret_cast<uint64_t>(t.g0 + lofs[i]);
#pragma unroll
      for (int k = 0; k < G; ++k) p[k][i] = ldg_nc_u32(gl + (uint64_t)((uint32_t)min(k, nsrc - 1) * rowb));
      gq[i] = gl + (uint64_t)((uint32_t)G * rowb);
    }
    const Sched* sp = sh.sched;
#pragma unroll 1
    for (int rr0 = 0; rr0 < nsrc; rr0 += G) {
      const uint32_t m = sh.fmask[rr0 >> 5] >> (rr0 & 31);
      const int rem = nsrc - G - rr0;                 // slot k is refilled while k < rem
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const float4 s0 = *reinterpret_cast<const float4*>(&sp[k].w[0][0]);
        const float2 s1 = *reinterpret_cast<const float2*>(&sp[k].w[2][0]);
        const uint64_t w0 = pack2(s0.x, s0.y), w1 = pack2(s0.z, s0.w), w2 = pack2(s1.x, s1.y);
        uint64_t f[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) {
          f[i] = conv_px<kWindow>(p[k][i], wsc, wof);
          if (k < rem) p[k][i] = ldg_nc_u32(gq[i]);
          gq[i] += rowb;
        }
        if (m & (1u << k)) {                            // warp-uniform: output row ycur is complete
          hrow();
#pragma unroll
          for (int i = 0; i < NL; ++i) {                // rotate the accumulators through the FMA operands
            A[i] = ffma2(f[i], w0, B[i]);
            B[i] = ffma2(f[i], w1, Cc[i]);
            Cc[i] = ffma2(f[i], w2, 0ull);
          }
        } else {
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
    const uint32_t op_end = smem_u32(ws.o) + 4u * lane + (uint32_t)t.nrows * (kOPitch * 4);
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
  } else {
    // ---- output-stationary fallback: vertical upscaling, or a source row feeding more than three output rows ------
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
        const uint32_t ro = (uint32_t)(info.x - t.r_lo + k) * (uint32_t)t.wq;
#pragma unroll
        for (int i = 0; i < NL; ++i) A[i] = ffma2(conv_px<kWindow>(__ldg(t.g0 + ro + lofs[i]), wsc, wof), wp, A[i]);
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
  const int band = blockIdx.x % a.nbands;       // == rank in cluster
  const int plane = blockIdx.x / a.nbands;      // view * C + c
  const int view = plane / a.C;
  const int chan = plane - view * a.C;
  const int s = a.s;

  const bool nocl = (a.debug_no_cluster & 1) != 0;
  if (!nocl) cluster_arrive_relaxed();   // phase 1: "every CTA of the cluster is running" (waited before DSMEM use)

  const MisViewParams P = a.params[view];
  const int y0 = band * kBand;
  const int nrows = min(kBand, s - y0);
  const int64_t plane_base = (int64_t)P.img * a.img_stride + (int64_t)chan * a.H * a.W;
  const int64_t e0 = plane_base + (int64_t)P.top * a.W + P.left;   // element index of crop (0,0)
  const float vscale = (float)P.h / (float)s;
  const float hscale = (float)P.w / (float)s;
  const float vsup = vscale >= 1.f ? vscale : 1.f, vinv = vscale >= 1.f ? 1.f / vscale : 1.f;
  const float hsup = hscale >= 1.f ? hscale : 1.f, hinv = hscale >= 1.f ? 1.f / hscale : 1.f;

  // ---- pull the band's crop rows into L2 right away (every thread derives the row range itself) -------------
  if (!(a.debug_no_cluster & 2)) {
    int lo0, hi0, lo1, hi1;
    float ctr;
    aa_window(y0, P.h, vscale, vsup, lo0, hi0, ctr);
    aa_window(y0 + nrows - 1, P.h, vscale, vsup, lo1, hi1, ctr);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a.src + e0);
    const int head = (int)(reinterpret_cast<uintptr_t>(base) & 127);
    const int lines = (head + 2 * P.w + 127) >> 7;               // 128-byte lines per crop row
    const int total = (hi1 - lo0) * lines;
    for (int i = tid; i < total; i += nthreads) {
      const int r = lo0 + i / lines, l = i - (i / lines) * lines;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)r * a.W * 2 - head + l * 128));
    }
  }

  // ---- vertical tap tables of the band (one thread per output row) ------------------------------------------
  if (tid == 0) sh.m_max = 0;
  if (tid < 8) sh.fmask[tid] = 0u;
  if (tid < nrows) {
    int lo, hi;
    float ctr;
    aa_window(y0 + tid, P.h, vscale, vsup, lo, hi, ctr);
    const int kcap = min(2 * (int)ceilf(vsup) + 1, kVK);
    int size = hi - lo;
    size = size < 0 ? 0 : (size > kcap ? kcap : size);
    float* w = sh.vw[tid];
    float total = 0.f;
    for (int j = 0; j < size; ++j) {
      const float wj = aa_tri(j + lo, ctr, vinv);
      w[j] = wj;
      total += wj;
    }
    for (int j = 0; j < kVK; ++j) w[j] = (j < size) ? (total != 0.f ? w[j] / total : w[j]) : 0.f;
    sh.vinfo[tid] = make_int2(lo, size);
  }
  __syncthreads();
  const int r_lo = sh.vinfo[0].x;
  const int r_hi = sh.vinfo[nrows - 1].x + sh.vinfo[nrows - 1].y;
  const int nsrc = r_hi - r_lo;
  const bool down = (vscale >= 1.f) && (nsrc + 8 <= kSchedCap) && (nsrc > 0);
  if (down) {
    // input-stationary schedule (windows are monotone in y): one entry per source row, padded to a whole group
    const int npad = (nsrc + 7) & ~7;
    for (int rr = tid; rr < npad; rr += nthreads) {
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;
      int first = 0;
      if (rr < nsrc) {
        const int r = r_lo + rr;
        int last = -1, prev = nrows;
        first = nrows;
        for (int yy = 0; yy < nrows; ++yy) {
          const int2 info = sh.vinfo[yy];
          if (info.x <= r && r < info.x + info.y) {
            first = min(first, yy);
            last = yy;
          }
          if (info.x <= r - 1 && r - 1 < info.x + info.y) prev = min(prev, yy);
        }
        if (last < 0) first = 0;       // cannot happen (windows overlap); keeps the stream monotone anyway
        if (first <= last) w0 = sh.vw[first][r - sh.vinfo[first].x];
        if (first + 1 <= last) w1 = sh.vw[first + 1][r - sh.vinfo[first + 1].x];
        if (first + 2 <= last) w2 = sh.vw[first + 2][r - sh.vinfo[first + 2].x];
        // the stream handles at most three open output rows and completes at most ONE output row per source row
        int m = last - first + 1;
        if (rr == 0 ? (first != 0) : (first - prev > 1)) m = 4;
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
  __syncthreads();

  // ---- this warp's tile -------------------------------------------------------------------------------------
  Tile t;
  t.win_lo = a.win_lo;
  t.win_scale = a.win_scale;
  t.sh = &sh;
  t.ws = &ws;
  t.lane = lane;
  t.nrows = nrows;
  t.nsrc = nsrc;
  t.r_lo = r_lo;
  t.use_is = down && sh.m_max <= 3;
  t.post = kWindow ? 1.f : (1.f / 65535.f);
  t.hinv = hinv;
  const bool jitter = (P.flags & MIS_VIEW_JITTER) != 0;
  // brightness (op 0) before contrast (op 1) is applied while the tile is produced: the contrast mean needs it
  int pos_b = 0, pos_c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (P.order[k] == 0) pos_b = k;
    if (P.order[k] == 1) pos_c = k;
  }
  t.has_pre = jitter && pos_b < pos_c;
  t.pre_b = P.brightness;
  const bool has_post = jitter && pos_b > pos_c;

  const int x0 = warp * 32;
  const int x = x0 + lane;
  t.valid = x < s;
  int hlo = 0, hhi = 0;
  float hctr = 0.f;
  if (t.valid) aa_window(x, P.w, hscale, hsup, hlo, hhi, hctr);
  {
    const int kcap = 2 * (int)ceilf(hsup) + 1;
    int size = hhi - hlo;
    size = size < 0 ? 0 : (size > kcap ? kcap : size);
    t.hsize = t.valid ? size : 0;
  }
  t.hlo = hlo;
  t.hctr = hctr;
  const int c_lo = __shfl_sync(0xffffffffu, hlo, 0);
  const int ca = c_lo - (int)((e0 + c_lo) & 1);                       // 4-byte aligned first staged column
  const int c_hi = __reduce_max_sync(0xffffffffu, t.valid ? hlo + t.hsize : 0);
  t.span = c_hi - ca;
  t.off = t.valid ? ((hlo - ca) & 3) : 0;
  t.xa = t.valid ? ((hlo - ca) & ~3) : 0;
  const int need = __reduce_max_sync(0xffffffffu, t.off + t.hsize);
  t.wq = a.W >> 1;
  t.g0 = reinterpret_cast<const uint32_t*>(a.src + e0 + (int64_t)r_lo * a.W + ca);

  float sum;
  if (t.span <= 64 && need <= 8) sum = run_tile<1, 2, kWindow>(t);
  else if (t.span <= 128 && need <= 12) sum = run_tile<2, 3, kWindow>(t);
  else sum = run_tile<3, 4, kWindow>(t);

  // ================================ contrast mean over the view ===========================================
  if (!nocl) cluster_wait_acquire();   // phase 1 done: all CTAs of the cluster are resident
  float cadd = 0.f;
  const float cf = P.contrast;
  if (jitter) {
    sum = warp_sum(sum);
    if (lane == 0) sh.red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int i = 0; i < (nthreads >> 5); ++i) tot += sh.red[i];
      if (nocl) { for (int r = 0; r < a.nbands; ++r) sh.part[r] = tot; }
      else for (int r = 0; r < a.nbands; ++r) st_cluster_f32(&sh.part[band], (uint32_t)r, tot);
    }
    if (nocl) __syncthreads();
    else {
      cluster_arrive_release();
      cluster_wait_acquire();
    }
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
  // class (3,4) covers 5.5x downscaling per axis: 31*5.5 + 2*5.5 + 2 <= 192 staged columns, 3 + 13 <= 16 aligned taps,
  // 31*5.5 + 13 + 8 <= kSchedCap schedule rows
  return C == 1 && s >= 8 && s <= kBand * kMaxWarps && (W & 1) == 0 && (img_stride & 1) == 0 && 2 * W <= 11 * s &&
         2 * H <= 11 * s;
}

int launch_tile(const TileArgs& a, int n_views, bool window, cudaStream_t stream) {
  const int nxw = (a.s + 31) / 32;
  const size_t smem = sizeof(Smem) + (size_t)nxw * sizeof(WarpSmem);
  auto* fn = window ? &aug_tile_kernel<true> : &aug_tile_kernel<false>;
  MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.nbands * n_views * a.C));
  cfg.blockDim = dim3((unsigned)(32 * nxw));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (a.debug_no_cluster & 1) ? 1u : (unsigned)a.nbands;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, a));
  return MIS_OK;
}

}  // namespace augt
}  // namespace mis
