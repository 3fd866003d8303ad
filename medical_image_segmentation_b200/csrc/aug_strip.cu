// K1 (strip variant, the default) -- fused two-view augmentation for 16-bit slices (sm_100a).
//
// One CTA per output view plane.  The plane is cut into strips of 32 output columns and nsy vertical parts; every
// (strip, part) is one WARP that streams its own source rows and never synchronises with another warp between the
// CTA-wide table build and the contrast mean:
//
//   tables   : all threads build the vertical tap table (one output row per thread) and from it the
//              input-stationary schedule of the WHOLE crop (one 16-byte entry per source row: the weights of the
//              three output rows it feeds, plus a bit mask "this row completes an output row") -- once per view
//              instead of once per 32-row band as in the round-1 tile kernel;
//   V pass   : lane = NL pairs of adjacent source columns, read straight from global memory with 4-byte loads
//              through G register slots refilled on consumption; every pixel is converted once (exact
//              magic-number u16 -> f32) and scattered with FFMA2 (scalar-broadcast weight operand) into the three
//              open output rows, which rotate through the FMA operands;
//   H pass   : a completed row goes through a warp-private row buffer; lane = one output column with its taps in
//              registers (pre-scaled so the result is in uint16 units); the result is PARKED AS UINT16 (round to
//              nearest of x*65535, |error| <= 7.7e-6 of full scale) in a warp-private tile -- 2*s*s bytes per view
//              instead of 4*s*s, which is what lets a whole 224x224 view (and two CTAs) live on one SM;
//   colour   : the contrast mean is a CTA reduction (no cluster, no DSMEM); views without jitter skip the barrier;
//   store    : 8 pixels per lane from the parked tile: contrast / brightness / normalise / flip, 16-byte stores.
//
// Vertical upscaling uses an output-stationary three-tap pass, windows the stream cannot express (a source row
// feeding more than three output rows -- never seen for downscaling) a generic output-stationary pass.
//
// Arithmetic restated from torchvision 0.26 / ATen (see oracle/aug_oracle.py):
//   taps   : _upsample_bilinear2d_aa (triangle filter, support = max(scale,1), weights normalised)
//   colour : functional/_color.py:114-125 (brightness), :190-205 + _blend :92-97 (contrast)
//   output : (x - mean) / std, functional/_misc.py:37-67
#include <cuda_bf16.h>

#include "aug_math.cuh"
#include "aug_strip.cuh"
#include "common.cuh"

namespace mis {
namespace augs {

using namespace mis::aug;

constexpr int kStrip = 32;         // output columns per warp
constexpr int kVK = 16;            // widest vertical window kept in the weight table (2*ceil(5.5)+1 = 13)
constexpr int kTilePitch = 64;     // bytes per parked row of a warp tile (32 x uint16)
constexpr int kMaxParts = 8;
constexpr float kMagic = 8388608.f;   // 2^23: float(2^23 + q) carries the integer q in its low mantissa bits

struct Misc {
  float red[32];
  int part_lo[kMaxParts];    // first source row of a part's stream (crop coordinates)
  int part_end[kMaxParts];   // one past its last source row
  int part_e0[kMaxParts];    // source row whose flush completes the part's first output row
  int m_max;                 // most output rows a single source row feeds (4 = the schedule cannot express the view)
  int pad[3];
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts64u(uint32_t addr, uint64_t v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t ptr_add(uint64_t p, uint32_t bytes) {     // one IMAD.WIDE instead of an add pair
  uint64_t r;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(bytes), "l"(p));
  return r;
}

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

// what a warp needs to know about its (strip, part)
struct Part {
  float win_lo, win_scale;
  uint32_t sched;          // shared address of the schedule / upscaling table
  uint32_t fmask;          // shared address of the completion bit mask
  uint32_t row;            // shared address of this warp's row buffer(s)
  uint32_t tile;           // shared address of this warp's parked tile
  int rowbuf;              // floats per row buffer
  const uint16_t* crop;    // crop (0, 0) of this plane in global memory
  int W;                   // source row pitch in elements
  int h;                   // crop height
  int ca;                  // first staged source column (crop coordinates, 4-byte aligned address)
  int lane;
  int y0, nrows;           // output rows of this part
  int rp;                  // output rows of a full part (kernel argument)
  int span;                // source columns this warp stages, from `ca`
  int hlo, hsize;          // this lane's output: first source column, taps
  float hctr, hinv;
  float out_k;             // folded into the horizontal taps: result in uint16 units (times an early brightness)
  int pf_groups;           // L2 prefetch distance of the stream, in groups of G rows
  int mode;                // 0 input-stationary stream, 1 three-tap upscaling, 2 generic
  int r_lo, r_end;         // mode 0: source rows [r_lo, r_end) of the stream
  int r_e0;                // mode 0: source row whose flush completes the part's first output row
  float vscale, vsup, vinv;
};

// One (strip, part) of one warp: 32 output columns, nrows output rows.  Returns this lane's share of the pixel sum
// (uint16 units).
template <int NL, int NS, bool kWindow, bool kDbl, int kW>
__device__ __forceinline__ float run_part(const Part& t) {
  const int lane = t.lane;
  const uint32_t toggle = kDbl ? 4u * (uint32_t)t.rowbuf : 0u;

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
    const float rtot = (tot != 0.f ? __frcp_rn(tot) : 1.f);
#pragma unroll
    for (int i = 0; i < 2 * NS; ++i) hw[i] = pack2(w[2 * i] * rtot * t.out_k, w[2 * i + 1] * rtot * t.out_k);
    rbase = opaque(t.row + 4u * xa);             // aligned window start in buffer 0
  }
  // zero-weight taps may read up to 15 columns behind the staged span: those must hold finite values
  if (lane < 24) {
    const int c = (t.span & ~1) + lane;
    if (c < t.rowbuf) {
      sts32(t.row + 4u * c, 0.f);
      sts32(t.row + toggle + 4u * c, 0.f);
    }
  }
  __syncwarp();

  bool act[NL];
  uint32_t wb[NL];
  uint64_t gl[NL];
  // (idle lanes read a valid address and park their finite garbage in the last two floats of the buffer)
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    act[i] = (2 * lane + 64 * i) < t.span;
    wb[i] = opaque(t.row + (act[i] ? 8u * lane + 256u * i : 4u * (t.rowbuf - 2)));
  }
  const uint64_t wsc = pack2(t.win_scale, t.win_scale);
  const uint64_t wof = pack2(-t.win_lo * t.win_scale, -t.win_lo * t.win_scale);
  const uint32_t rowb = kW > 0 ? 2u * (uint32_t)kW : 2u * (uint32_t)t.W;   // source row pitch in bytes
#pragma unroll
  for (int i = 0; i < NL; ++i)
    gl[i] = reinterpret_cast<uint64_t>(t.crop + t.ca + (act[i] ? 2 * lane + 64 * i : 0));

  uint64_t A[NL], B[NL], Cc[NL];    // the three open output rows of this lane's column pair(s)
#pragma unroll
  for (int i = 0; i < NL; ++i) A[i] = B[i] = Cc[i] = 0ull;
  float sum = 0.f;
  uint32_t sel = 0;                                  // 0 / toggle: row buffer in use
  uint32_t op = opaque(t.tile + 2u * lane);          // where this lane parks its next result

  // the oldest open output row is complete in A: H pass over the intermediate row, result parked as uint16
  // first half: park the intermediate row, read this lane's aligned window
  float4 hv[NS];
  auto hrow_a = [&]() {
    if (!kDbl) __syncwarp();                         // single buffer: the previous row's window reads are done
#pragma unroll
    for (int i = 0; i < NL; ++i) sts64u(wb[i] + sel, A[i]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NS; ++j) hv[j] = lds128(rbase + sel + 16u * j);
  };
  // second half: taps, clamp, sum, park the result as uint16
  auto hrow_b = [&]() {
    uint64_t acc = 0ull;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      acc = ffma2(pack2(hv[j].x, hv[j].y), hw[2 * j], acc);
      acc = ffma2(pack2(hv[j].z, hv[j].w), hw[2 * j + 1], acc);
    }
    float lo, hi;
    unpack2(acc, lo, hi);
    // uint16 units; the clamp is the upper half of clamp(x * brightness, 0, 1) (everything here is >= 0) and trims
    // the one-ulp overshoot a normalised filter can produce
    const float val = fminf(lo + hi, 65535.f);
    sum += val;                                   // lanes beyond the view's last column carry zero weights
    sts_u16(op, __float_as_uint(val + kMagic));   // low 16 mantissa bits of 2^23 + val = round-to-nearest(val)
    op += kTilePitch;
    if (kDbl) sel ^= toggle;
  };
  auto hrow = [&]() {
    hrow_a();
    hrow_b();
  };

  if (t.mode == 0) {
    // ---- input-stationary stream over the part's source rows --------------------------------------------
    // The stream starts at the group-aligned row at or before the part's first row: the rows in front of it only
    // feed output rows of the previous part, whose (partial) results are dropped: output row y is complete when the
    // stream reaches source row e(y) = one past its window, so this part emits exactly at the flushing rows in
    // [e(y0), r_end].
    constexpr int G = NL == 1 ? 8 : 4;
    // (warp-uniform by construction; the reductions tell the compiler so: uniform registers and plain branches
    // instead of convergence barriers around every per-row decision)
    const int r_start = __reduce_max_sync(0xffffffffu, t.r_lo & ~(G - 1));
    const int r_end = __reduce_max_sync(0xffffffffu, t.r_end);
    const int r_e0 = __reduce_max_sync(0xffffffffu, t.r_e0);
    int cur = 0;                                      // output rows emitted so far
    uint32_t p[G][NL];
    uint64_t gq[NL];                                  // kW == 0: next row to fetch; kW > 0: first row of the current group
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const uint64_t g0 = gl[i] + (uint64_t)((uint32_t)r_start * rowb);
#pragma unroll
      for (int k = 0; k < G; ++k) p[k][i] = ldg_nc_u32(g0 + (uint64_t)((uint32_t)min(k, r_end - 1 - r_start) * rowb));
      gq[i] = kW > 0 ? g0 : g0 + (uint64_t)((uint32_t)G * rowb);
    }
    // L2 prefetch pf_groups groups ahead of the register slots: one instruction per group, lane = (row of the group,
    // 128-byte line of the strip's staged span).  The slots alone keep G rows (< 1 us of work) in flight, less than the
    // DRAM latency under load.
    constexpr int kLinesLog2 = G == 8 ? 2 : 3;
    uint64_t pfa;
    int pf_row;                                      // crop row this lane prefetches next
    bool pf_lane;
    {
      const uint64_t first = reinterpret_cast<uint64_t>(t.crop + t.ca);
      const uint32_t head = (uint32_t)(first & 127u);
      const int line = lane & ((1 << kLinesLog2) - 1);
      pf_lane = t.pf_groups > 0 && (uint32_t)(line * 128) < head + 2u * (uint32_t)t.span;
      pf_row = r_start + G + (lane >> kLinesLog2);
      pfa = first - head + (uint64_t)((uint32_t)pf_row * rowb) + (uint64_t)(line * 128);
#pragma unroll 1
      for (int j = 0; j < t.pf_groups - 1; ++j) {
        if (pf_lane && pf_row < r_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(pfa));
        pf_row += G;
        pfa += (uint64_t)((uint32_t)G * rowb);
      }
    }
    uint32_t sp = t.sched + 16u * r_start;
    // schedule entries and mask words are requested one row / one group ahead of their use
    float4 sn = lds128(sp);
    uint32_t mw = lds32(t.fmask + 4u * (r_start >> 5));
#pragma unroll 1
    for (int rr0 = r_start; rr0 < r_end; rr0 += G) {
      const uint32_t m = __reduce_or_sync(0xffffffffu, mw >> (rr0 & 31));
      // rows rr0 + k of this group that emit: flushing rows with e(y0) <= row <= r_end
      const int lo_c = min(max(r_e0 - rr0, 0), G), hi_c = min(max(r_end + 1 - rr0, 0), G);
      const uint32_t me = m & ((1u << hi_c) - 1u) & ~((1u << lo_c) - 1u);
      cur += __popc(me);
      mw = lds32(t.fmask + 4u * ((rr0 + G) >> 5));
      const int rem = r_end - G - rr0;                 // slot k is refilled while k < rem
      if (pf_lane && pf_row < r_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(pfa));
      pf_row += G;
      pfa += (uint64_t)((uint32_t)G * rowb);
#pragma unroll
      for (int k = 0; k < G; ++k) {
        uint64_t f[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) {
          f[i] = conv_px<kWindow>(p[k][i], wsc, wof);
          if (kW > 0) {
            if (k < rem) p[k][i] = ldg_nc_u32(gq[i] + (uint64_t)((uint32_t)(G + k) * rowb));
          } else {
            if (k < rem) p[k][i] = ldg_nc_u32(gq[i]);
            gq[i] = ptr_add(gq[i], rowb);
          }
        }
        const float4 s0 = sn;                          // {w0, w1, w2, info}: weights of the three open rows
        sn = lds128(sp + 16u * (k + 1));               // (the schedule is padded by a whole group)
        const uint64_t w0 = pack2(s0.x, s0.x), w1 = pack2(s0.y, s0.y), w2 = pack2(s0.z, s0.z);
        if (m & (1u << k)) {                            // warp-uniform: the oldest open output row is complete
          if (me & (1u << k)) hrow();
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
      if (kW > 0) {
#pragma unroll
        for (int i = 0; i < NL; ++i) gq[i] += (uint64_t)((uint32_t)G * rowb);
      }
      sp += 16u * G;
    }
#pragma unroll 1
    while (cur < t.nrows) {
      hrow();
      ++cur;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        A[i] = B[i];
        B[i] = Cc[i];
        Cc[i] = 0ull;
      }
    }
  } else if (t.mode == 1) {
    // ---- three taps (vertical upscaling): table entry y = {w0, w1, w2, first source row}.  The loop runs over the
    // window position (one source row per step, unrolled over a ring of four raw rows in flight, so the ring index
    // is a compile-time constant); the output rows whose window starts there -- one or two when upscaling -- are
    // produced before the window slides.  The L2 prefetch runs 16 rows ahead of the ring.
    uint64_t f0[NL], f1[NL], f2[NL];
    uint32_t q[4][NL];
    float4 e = lds128(t.sched + 16u * t.y0);
    const int base0 = __reduce_max_sync(0xffffffffu, __float_as_int(e.w));     // (warp-uniform, see mode 0)
    const int hm1 = t.h - 1;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      auto at = [&](int r) { return ldg_nc_u32(gl[i] + (uint64_t)((uint32_t)min(r, hm1) * rowb)); };
      const uint32_t r0 = at(base0), r1 = at(base0 + 1), r2 = at(base0 + 2);
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k][i] = at(base0 + 3 + k);
      f0[i] = conv_px<kWindow>(r0, wsc, wof);
      f1[i] = conv_px<kWindow>(r1, wsc, wof);
      f2[i] = conv_px<kWindow>(r2, wsc, wof);
    }
    uint64_t pfb;
    bool pf_lane;
    {
      const uint64_t first = reinterpret_cast<uint64_t>(t.crop + t.ca);
      const uint32_t head = (uint32_t)(first & 127u);
      pf_lane = (uint32_t)(lane * 128) < head + 2u * (uint32_t)t.span;
      pfb = first - head + (uint64_t)(lane * 128);
#pragma unroll 1
      for (int r = base0 + 7; r < base0 + 23; ++r)
        if (pf_lane && r <= hm1) asm volatile("prefetch.global.L2 [%0];" ::"l"(pfb + (uint64_t)((uint32_t)r * rowb)));
    }
    int y = 0;
    int lo_e = base0;                                 // first source row of the pending output row
    const int nrows = t.nrows;
#pragma unroll 1
    for (int b0 = base0; y < nrows; b0 += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll 1
        while (y < nrows && lo_e <= b0 + k) {          // (== by monotonicity; <= so that nothing can make the loop spin)
          const uint64_t w0 = pack2(e.x, e.x), w1 = pack2(e.y, e.y), w2 = pack2(e.z, e.z);
          ++y;
          e = lds128(t.sched + 16u * (t.y0 + y));     // (the table is padded: reading one entry past the part is harmless)
#pragma unroll
          for (int i = 0; i < NL; ++i) A[i] = ffma2(f2[i], w2, ffma2(f1[i], w1, ffma2(f0[i], w0, 0ull)));
          hrow();
          lo_e = __reduce_max_sync(0xffffffffu, __float_as_int(e.w));
        }
        const int rn = b0 + k + 7;                    // the ring slot is re-armed with the row four behind the newest
#pragma unroll
        for (int i = 0; i < NL; ++i) {
          f0[i] = f1[i];
          f1[i] = f2[i];
          f2[i] = conv_px<kWindow>(q[k][i], wsc, wof);
          q[k][i] = ldg_nc_u32(gl[i] + (uint64_t)((uint32_t)min(rn, hm1) * rowb));
        }
        if (pf_lane && rn + 16 <= hm1) asm volatile("prefetch.global.L2 [%0];" ::"l"(pfb + (uint64_t)((uint32_t)(rn + 16) * rowb)));
      }
    }
  } else {
    // ---- output-stationary, any window: the taps are recomputed per output row (warp-uniform arithmetic) --------
#pragma unroll 1
    for (int y = 0; y < t.nrows; ++y) {
      int lo, hi;
      float ctr;
      aa_window(t.y0 + y, t.h, t.vscale, t.vsup, lo, hi, ctr);
      const int kcap = min(2 * (int)ceilf(t.vsup) + 1, kVK);
      int size = hi - lo;
      size = size < 0 ? 0 : (size > kcap ? kcap : size);
      float total = 0.f;
      for (int k = 0; k < size; ++k) total += aa_tri(lo + k, ctr, t.vinv);
      const float rtot = total != 0.f ? __frcp_rn(total) : 1.f;
#pragma unroll
      for (int i = 0; i < NL; ++i) A[i] = 0ull;
#pragma unroll 1
      for (int k = 0; k < size; ++k) {
        const float w = aa_tri(lo + k, ctr, t.vinv) * rtot;
        const uint64_t wp = pack2(w, w);
        const uint64_t ro = (uint64_t)((uint32_t)(lo + k) * rowb);
#pragma unroll
        for (int i = 0; i < NL; ++i) A[i] = ffma2(conv_px<kWindow>(ldg_nc_u32(gl[i] + ro), wsc, wof), wp, A[i]);
      }
      hrow();
    }
  }
  return sum;
}

// 8 pixels per lane from the parked tile: contrast / late brightness / solarize / normalise / flip, one 16-byte store
// (bf16).  kRaw: the view drew a GaussianBlur -- the post-colour image is left as uint16 (round(x * 65535)) in the first
// 2*s*s bytes of its output plane for mis_aug_blur_views (aug_blur.cu), which blurs, solarizes and normalises it.
template <bool kFlip, bool kF32, bool kRaw>
__device__ __forceinline__ void store_tile(uint32_t tile, int nrows, int lane, int x0, int s, uint8_t* plane_ptr, int y0,
                                           bool jitter, bool has_post, bool sol, float cs, float cadd, float pb,
                                           float mean, float inv_std) {
  const int chunk = lane & 3, r0 = lane >> 2;
  const int xs = x0 + 8 * chunk;
  if (xs >= s) return;
  const int col = kFlip ? (s - xs - 8) : xs;
  uint32_t ta = tile + (uint32_t)(r0 * kTilePitch + 16 * chunk);
  constexpr size_t esz = kRaw ? 2 : (kF32 ? 4 : 2);
  uint8_t* op = plane_ptr + ((size_t)(y0 + r0) * (size_t)s + col) * esz;
  const size_t step = (size_t)8 * s * esz;
  const uint64_t nmagic = pack2(-kMagic, -kMagic);
  const uint64_t nmean = pack2(-mean, -mean), istd = pack2(inv_std, inv_std);
  const uint64_t unit = pack2(1.f / 65535.f, 1.f / 65535.f);
#pragma unroll 1
  for (int row = r0; row < nrows; row += 8, ta += 8 * kTilePitch, op += step) {
    const uint4 q = lds128u(ta);
    const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
    uint64_t u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t lo = __byte_perm(qq[i], 0x4B000000u, 0x7610);
      const uint32_t hi = __byte_perm(qq[i], 0x4B000000u, 0x7632);
      u[i] = fadd2(pack2(__uint_as_float(lo), __uint_as_float(hi)), nmagic);
    }
    if (jitter) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a, b;
        unpack2(u[i], a, b);
        a = __saturatef(fmaf(a, cs, cadd));
        b = __saturatef(fmaf(b, cs, cadd));
        if (has_post) {
          a = __saturatef(a * pb);
          b = __saturatef(b * pb);
        }
        u[i] = pack2(a, b);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] = fmul2(u[i], unit);
    }
    float v[8];
    if (kRaw) {
#pragma unroll
      for (int i = 0; i < 4; ++i) unpack2(u[i], v[2 * i], v[2 * i + 1]);
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e0 = kFlip ? v[7 - 2 * i] : v[2 * i], e1 = kFlip ? v[6 - 2 * i] : v[2 * i + 1];
        const uint32_t b0 = __float_as_uint(fmaf(e0, 65535.f, kMagic)), b1 = __float_as_uint(fmaf(e1, 65535.f, kMagic));
        pk[i] = __byte_perm(b0, b1, 0x5410);          // low 16 mantissa bits = round-to-nearest(x * 65535)
      }
      *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      continue;
    }
    if (sol) {        // RandomSolarize(128) on the [0,1] scale: x >= 128/255 -> 1 - x (functional/_color.py:497-501)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a, b;
        unpack2(u[i], a, b);
        a = a >= MIS_SOLARIZE_THRESHOLD ? 1.f - a : a;
        b = b >= MIS_SOLARIZE_THRESHOLD ? 1.f - b : b;
        u[i] = pack2(a, b);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = fmul2(fadd2(u[i], nmean), istd);
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2(u[i], v[2 * i], v[2 * i + 1]);
    if (kF32) {
      float4* dst = reinterpret_cast<float4*>(op);
      if (!kFlip) {
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        dst[0] = make_float4(v[7], v[6], v[5], v[4]);
        dst[1] = make_float4(v[3], v[2], v[1], v[0]);
      }
    } else {
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = kFlip ? __floats2bfloat162_rn(v[7 - 2 * i], v[6 - 2 * i])
                                       : __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

template <bool kWindow, int kMaxThreads, int kMinBlocks, bool kDbl, int kW>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) aug_strip_kernel(const __grid_constant__ StripArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int nthreads = blockDim.x;
  const int nwarps = nthreads >> 5;
  const int sy = warp / a.nsx;
  const int sx = warp - sy * a.nsx;
  const int vb = a.C == 1 ? (int)blockIdx.x : (int)blockIdx.x / a.C;
  const int chan = (int)blockIdx.x - vb * a.C;
  const int view = a.order ? a.order[vb] : vb;   // most expensive views first (mis_view_cost_order)
  const int plane = view * a.C + chan;
  const int s = a.s;
  Misc& misc = *reinterpret_cast<Misc*>(smem + a.off_misc);
  float4* const sched = reinterpret_cast<float4*>(smem + a.off_sched);
  uint32_t* const fmask = reinterpret_cast<uint32_t*>(smem + a.off_fmask);
  // vertical tap table of the view: lives in the (not yet used) parked-tile area until the schedule is built
  float* const vw = reinterpret_cast<float*>(smem);                       // [s][kVK]
  int2* const vinfo = reinterpret_cast<int2*>(smem + (size_t)s * kVK * 4);  // [s] {first source row, taps}

  const MisViewParams P = a.params[view];
  const int64_t plane_base = (int64_t)P.img * a.img_stride + (int64_t)chan * a.H * a.W;
  const int64_t e0 = plane_base + (int64_t)P.top * a.W + P.left;   // element index of crop (0,0)
  const float vscale = (float)P.h / (float)s;
  const float hscale = (float)P.w / (float)s;
  const float vsup = vscale >= 1.f ? vscale : 1.f, vinv = vscale >= 1.f ? 1.f / vscale : 1.f;
  const float hsup = hscale >= 1.f ? hscale : 1.f, hinv = hscale >= 1.f ? 1.f / hscale : 1.f;
  const bool vdown = vscale >= 1.f;
  const int y0 = sy * a.rp;
  const int nrows = max(0, min(a.rp, s - y0));

  // ---- vertical taps: one output row per thread -------------------------------------------------------------
  const int hpad = (P.h + 7) & ~7;
  if (tid == 0) misc.m_max = 0;
  for (int i = tid; i < (hpad >> 5) + 1; i += nthreads) fmask[i] = 0u;
  if (tid < s) {
    int lo, hi;
    float ctr;
    aa_window(tid, P.h, vscale, vsup, lo, hi, ctr);
    const int kcap = min(2 * (int)ceilf(vsup) + 1, kVK);
    int size = hi - lo;
    size = size < 0 ? 0 : (size > kcap ? kcap : size);
    float wj[kVK];
    float total = 0.f;
    if (size <= 8) {                                  // the usual case: no loops
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wj[j] = (j < size) ? aa_tri(j + lo, ctr, vinv) : 0.f;
        total += wj[j];                               // same order as the reference (zeros beyond the window)
      }
#pragma unroll
      for (int j = 8; j < kVK; ++j) wj[j] = 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < kVK; ++j) {
        wj[j] = (j < size) ? aa_tri(j + lo, ctr, vinv) : 0.f;
        total += wj[j];
      }
    }
    const float rtot = total != 0.f ? __frcp_rn(total) : 1.f;
    if (vdown) {
      float4* w4 = reinterpret_cast<float4*>(vw + (size_t)tid * kVK);
#pragma unroll
      for (int j = 0; j < kVK / 4; ++j)
        w4[j] = make_float4(wj[4 * j] * rtot, wj[4 * j + 1] * rtot, wj[4 * j + 2] * rtot, wj[4 * j + 3] * rtot);
      vinfo[tid] = make_int2(lo, size);
    } else {
      // upscaling: at most three taps; the table entry IS the pass's working set
      sched[tid] = make_float4(wj[0] * rtot, wj[1] * rtot, wj[2] * rtot, __int_as_float(lo));
    }
    // stream bounds of the parts (the table above is recycled as tile storage once the first warp parks a row)
    const int py = tid / a.rp;
    if (tid == py * a.rp) {
      misc.part_lo[py] = lo;
      misc.part_e0[py] = lo + size;
    }
    if (tid == min(s, (py + 1) * a.rp) - 1) misc.part_end[py] = lo + size;
  }

  // ---- this warp's strip: horizontal windows (independent of the tables above) -----------------------------
  Part t;
  t.win_lo = a.win_lo;
  t.win_scale = a.win_scale;
  t.sched = smem_u32(sched);
  t.fmask = smem_u32(fmask);
  t.rowbuf = a.rowbuf;
  t.row = smem_u32(smem + a.off_row) + (uint32_t)warp * (kDbl ? 8u : 4u) * (uint32_t)a.rowbuf;
  t.tile = smem_u32(smem) + (uint32_t)warp * (uint32_t)(a.rp * kTilePitch);
  t.crop = a.src + e0;
  t.W = a.W;
  t.h = P.h;
  t.lane = lane;
  t.y0 = y0;
  t.nrows = nrows;
  t.rp = a.rp;
  t.hinv = hinv;
  t.vscale = vscale;
  t.pf_groups = a.pf_groups;
  t.vsup = vsup;
  t.vinv = vinv;
  // raw_all (3-channel input): the colour ops mix channels, so this kernel only resamples and flips and leaves every
  // plane as uint16 for the colour kernel (aug_rgb.cu)
  const bool jitter = !a.raw_all && (P.flags & MIS_VIEW_JITTER) != 0;
  // brightness (op 0) before contrast (op 1) is applied while the tile is produced: the contrast mean needs it
  int pos_b = 0, pos_c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (P.order[k] == 0) pos_b = k;
    if (P.order[k] == 1) pos_c = k;
  }
  const float unit = kWindow ? 65535.f : 1.f;          // the windowed pixel is already in [0, 1]
  t.out_k = (jitter && pos_b < pos_c) ? unit * P.brightness : unit;
  const bool has_post = jitter && pos_b > pos_c;

  const int x0 = sx * kStrip;
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

  // ---- the strip's staged span (from the 4-byte aligned column at or before its first window) and class ---------
  const int c_lo = __shfl_sync(0xffffffffu, t.hlo, 0);
  t.ca = c_lo - (int)((e0 + c_lo) & 1);
  t.span = __reduce_max_sync(0xffffffffu, t.hsize > 0 ? t.hlo + t.hsize : 0) - t.ca;
  const int need = __reduce_max_sync(0xffffffffu, t.hsize > 0 ? ((t.hlo - t.ca) & 3) + t.hsize : 0);
  // pull the first 16 rows of this warp's stream towards L2 while the tables are built (lane = row, 128-byte line)
  if (nrows > 0) {
    int lo, hi;
    float ctr;
    aa_window(y0, P.h, vscale, vsup, lo, hi, ctr);
    const uint64_t first = reinterpret_cast<uint64_t>(t.crop + t.ca);
    const uint32_t head = (uint32_t)(first & 127u);
    const int line = lane & 3;
    if ((uint32_t)(line * 128) < head + 2u * (uint32_t)t.span) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = (vdown ? (lo & ~7) : lo) + (lane >> 2) + 8 * j;
        if (r < P.h) asm volatile("prefetch.global.L2 [%0];" ::"l"(first - head + (uint64_t)((uint32_t)r * 2u * (uint32_t)a.W) + (uint64_t)(line * 128)));
      }
    }
  }

  __syncthreads();                                  // tap table complete
  if (vdown) {
    // input-stationary schedule (windows are monotone in y): one entry per source row of the crop.
    // Source row r can only lie in the windows of the output rows around (r + 0.5) / vscale - 0.5.
    const float inv_vs = 1.f / vscale;
    for (int r = tid; r < hpad; r += nthreads) {
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;
      int first = s;
      uint32_t flush = 0;
      if (r < P.h) {
        const int yc = (int)floorf(((float)r + 0.5f) * inv_vs - 0.5f);
        int last = -1, prev = s;
#pragma unroll
        for (int d = -3; d <= 3; ++d) {
          const int yy = yc + d;
          if (yy >= 0 && yy < s) {
            const int2 info = vinfo[yy];
            if (info.x <= r && r < info.x + info.y) {
              first = min(first, yy);
              last = max(last, yy);
            }
            if (info.x <= r - 1 && r - 1 < info.x + info.y) prev = min(prev, yy);
          }
        }
        // the stream handles at most three open output rows and completes at most ONE output row per source row;
        // anything else (never seen for downscaling windows) sends the view to the generic pass
        int m = last - first + 1;
        if (last < 0 || (r == 0 ? (first != 0) : (prev >= s || first - prev > 1 || first < prev))) {
          m = 4;
          first = 0;
          last = -1;
        }
        if (first <= last) w0 = vw[(size_t)first * kVK + (r - vinfo[first].x)];
        if (first + 1 <= last) w1 = vw[(size_t)(first + 1) * kVK + (r - vinfo[first + 1].x)];
        if (first + 2 <= last) w2 = vw[(size_t)(first + 2) * kVK + (r - vinfo[first + 2].x)];
        if (m > 3) atomicMax(&misc.m_max, m);
        if (r > 0 && first > prev) {
          flush = 1;
          atomicOr(&fmask[r >> 5], 1u << (r & 31));
        }
      }
      sched[r] = make_float4(w0, w1, w2, __uint_as_float((uint32_t)first | (flush << 16)));
    }
  }

  __syncthreads();                                  // schedule complete; vw / vinfo are dead from here on
  t.mode = !vdown ? 1 : (misc.m_max <= 3 ? 0 : 2);
  t.r_lo = misc.part_lo[sy];
  t.r_end = misc.part_end[sy];
  t.r_e0 = misc.part_e0[sy];

  float sum = 0.f;
  if (nrows > 0) {
    if (t.span <= 64 && need <= 8) sum = run_part<1, 2, kWindow, kDbl, kW>(t);
    else if (t.span <= 128 && need <= 12) sum = run_part<2, 3, kWindow, kDbl, kW>(t);
    else sum = run_part<3, 4, kWindow, kDbl, kW>(t);
  }

  // ================================ contrast mean over the view ===========================================
  float cadd = 0.f;
  const float cf = P.contrast;
  if (jitter) {
    sum = warp_sum(sum);
    if (lane == 0) misc.red[warp] = sum;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < nwarps; ++i) tot += misc.red[i];
    const float mu = tot / (float)(s * s) * (1.f / 65535.f);
    cadd = mu * (1.f - cf);
  } else {
    __syncwarp();
  }

  // ================================ colour, normalise, store ==============================================
  const float mean = a.mean[chan], inv_std = a.inv_std[chan];
  const bool flip = (P.flags & MIS_VIEW_FLIP) != 0;
  const float pb = P.brightness;
  const float cs = cf * (1.f / 65535.f);
  const bool sol = (P.flags & MIS_VIEW_SOLARIZE) != 0;
  const bool raw = a.raw_all || (P.flags & MIS_VIEW_BLUR) != 0;       // finished by the colour / blur kernels
  uint8_t* const plane_ptr = static_cast<uint8_t*>(a.out) + (size_t)plane * s * s * (a.out_f32 ? 4 : 2);
  if ((s & 7) == 0) {
    const bool f32 = a.out_f32 != 0;
#define MIS_STORE(F, T, R) store_tile<F, T, R>(t.tile, nrows, lane, x0, s, plane_ptr, y0, jitter, has_post, sol, cs, cadd, pb, mean, inv_std)
    if (raw) {
      if (flip) MIS_STORE(true, false, true);
      else MIS_STORE(false, false, true);
    } else if (!f32) {
      if (flip) MIS_STORE(true, false, false);
      else MIS_STORE(false, false, false);
    } else {
      if (flip) MIS_STORE(true, true, false);
      else MIS_STORE(false, true, false);
    }
#undef MIS_STORE
    return;
  }
  // crop sizes that are not a multiple of 8: element-wise stores
  const uint64_t nmagic = pack2(-kMagic, -kMagic);
  const int iters = (nrows + 7) >> 3;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const int id = it * 32 + lane;
    const int row = id >> 2, xs = x0 + 8 * (id & 3);
    if (row < nrows && xs < s) {
      const uint4 q = lds128u(t.tile + (uint32_t)(row * kTilePitch + 16 * (id & 3)));
      const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t lo = __byte_perm(qq[i], 0x4B000000u, 0x7610);
        const uint32_t hi = __byte_perm(qq[i], 0x4B000000u, 0x7632);
        unpack2(fadd2(pack2(__uint_as_float(lo), __uint_as_float(hi)), nmagic), v[2 * i], v[2 * i + 1]);
      }
      if (jitter) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __saturatef(fmaf(v[i], cs, cadd));
        if (has_post) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __saturatef(v[i] * pb);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] * (1.f / 65535.f);
      }
      if (sol) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] >= MIS_SOLARIZE_THRESHOLD ? 1.f - v[i] : v[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * inv_std;
      store_run<8>(v, a.out, ((size_t)plane * s + (y0 + row)) * s, xs, s, flip, a.out_f32 != 0);
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct Plan {
  int nsx, nsy, rp, rowbuf, dbl, threads, big;
  uint32_t off_sched, off_fmask, off_row, off_misc;
  size_t smem;
  int ctas;
};

static inline size_t al16(size_t v) { return (v + 15) & ~size_t(15); }

static bool make_plan(int H, int W, int s, Plan* best) {
  const int nsx = (s + kStrip - 1) / kStrip;
  // widest staged span of a strip: 31 output steps + one window + alignment
  const double sm = (double)W / s > 1.0 ? (double)W / s : 1.0;
  const int span_max = (int)(33.0 * sm + 3.0) + 2;
  if (span_max > 192) return false;
  const int rowbuf = (span_max + 18 + 7) & ~7;
  const int sched_cap = (((H > s ? H : s) + 7) & ~7) + 8;
  const size_t kSmPerSm = 233472, kSmPerBlock = 232448;
  bool found = false;
  int best_score = 0;
  for (int nsy = 1; nsy <= kMaxParts; ++nsy) {
    const int threads = 32 * nsx * nsy;
    if (threads > 1024) break;
    const int rp = (s + nsy - 1) / nsy;
    if (rp < 8 && nsy > 1) break;
    if ((s + rp - 1) / rp != nsy) continue;            // an empty trailing part: skip this split
    const int nwarps = nsx * nsy;
    size_t park = (size_t)nwarps * rp * kTilePitch;
    const size_t vtab = (size_t)s * (kVK * 4 + 8);
    if (park < vtab) park = vtab;
    for (int dbl = 1; dbl >= 0; --dbl) {
      Plan p = {};
      p.nsx = nsx; p.nsy = nsy; p.rp = rp; p.rowbuf = rowbuf; p.dbl = dbl; p.threads = threads;
      p.big = threads > 768 ? 2 : (threads > 448 ? 1 : 0);   // 0: (448, 2) 72 regs; 1: (768, 1) 85 regs; 2: (1024, 1) 64 regs
      size_t off = al16(park);
      p.off_sched = (uint32_t)off; off += (size_t)sched_cap * 16;
      p.off_fmask = (uint32_t)off; off += al16((size_t)(sched_cap / 32 + 2) * 4);
      p.off_row = (uint32_t)off; off += (size_t)nwarps * (dbl ? 2 : 1) * rowbuf * 4;
      p.off_misc = (uint32_t)off; off += sizeof(Misc);
      p.smem = off;
      if (p.smem > kSmPerBlock) continue;
      const int regs = p.big == 2 ? 64 : (p.big == 1 ? 85 : 72);
      int ctas = (int)(kSmPerSm / (p.smem + 1024));
      const int by_threads = 2048 / threads, by_regs = 65536 / (regs * threads);
      ctas = ctas < by_threads ? ctas : by_threads;
      ctas = ctas < by_regs ? ctas : by_regs;
      if (ctas > 32) ctas = 32;
      if (ctas < 1) continue;
      p.ctas = ctas;
      const int warps = ctas * nwarps;
      // More resident warps win; on a tie fewer parts (less halo, fewer redundant set-ups), then two row buffers.
      // The 1024-thread instantiation has 64 registers per thread and spills the load slots of the wide classes
      // (a spilled slot serialises the prefetch), so it only competes when the 72- / 85-register ones cannot reach 20
      // warps (256^2: one CTA per SM either way; 24 warps without spills beat 32 with).
      const int score = p.big == 2 ? warps : warps + 1000;
      if (!found || score > best_score) {
        if (p.big != 2 && warps < 20) continue;
        *best = p;
        best_score = score;
        found = true;
      }
    }
  }
  if (found) return true;
  // second pass without the 20-warp floor (small crops on small slices)
  for (int nsy = 1; nsy <= kMaxParts && !found; ++nsy) {
    const int threads = 32 * nsx * nsy;
    if (threads > 448) break;
    const int rp = (s + nsy - 1) / nsy;
    if ((s + rp - 1) / rp != nsy) continue;
    const int nwarps = nsx * nsy;
    size_t park = (size_t)nwarps * rp * kTilePitch;
    const size_t vtab = (size_t)s * (kVK * 4 + 8);
    if (park < vtab) park = vtab;
    Plan p = {};
    p.nsx = nsx; p.nsy = nsy; p.rp = rp; p.rowbuf = rowbuf; p.dbl = 0; p.threads = threads; p.big = 0;
    size_t off = al16(park);
    p.off_sched = (uint32_t)off; off += (size_t)sched_cap * 16;
    p.off_fmask = (uint32_t)off; off += al16((size_t)(sched_cap / 32 + 2) * 4);
    p.off_row = (uint32_t)off; off += (size_t)nwarps * rowbuf * 4;
    p.off_misc = (uint32_t)off; off += sizeof(Misc);
    p.smem = off;
    if (p.smem > kSmPerBlock) continue;
    p.ctas = 1;
    *best = p;
    found = true;
  }
  return found;
}

bool strip_supported(int C, int H, int W, int64_t img_stride, int s) {
  // class (3,4) covers 5.5x downscaling per axis: 31*5.5 + 2*5.5 + 2 <= 192 staged columns, 3 + 13 <= 16 aligned taps
  if (!((C == 1 || (C == 3 && (s & 7) == 0)) && s >= 8 && s <= 256 && (W & 1) == 0 && (img_stride & 1) == 0 &&
        ((int64_t)H * W & 1) == 0 && 2 * W <= 11 * s && 2 * H <= 11 * s))
    return false;
  Plan p;
  return make_plan(H, W, s, &p);
}

template <bool kWindow, int kMaxThreads, int kMinBlocks, bool kDbl, int kW>
static int launch_one(const StripArgs& a, int n_planes, const Plan& p, cudaStream_t stream) {
  auto* fn = &aug_strip_kernel<kWindow, kMaxThreads, kMinBlocks, kDbl, kW>;
  MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  fn<<<dim3((unsigned)n_planes), dim3((unsigned)p.threads), p.smem, stream>>>(a);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}
template <bool kWindow, int kMaxThreads, int kMinBlocks>
static int launch_shape(const StripArgs& a, int n_planes, const Plan& p, cudaStream_t stream) {
  // the row pitch of the common 512-wide slice is a compile-time constant of the hot loop (load offsets become
  // immediates); any other width takes the generic instantiation
  if (a.W == 512)
    return p.dbl ? launch_one<kWindow, kMaxThreads, kMinBlocks, true, 512>(a, n_planes, p, stream)
                 : launch_one<kWindow, kMaxThreads, kMinBlocks, false, 512>(a, n_planes, p, stream);
  return p.dbl ? launch_one<kWindow, kMaxThreads, kMinBlocks, true, 0>(a, n_planes, p, stream)
               : launch_one<kWindow, kMaxThreads, kMinBlocks, false, 0>(a, n_planes, p, stream);
}

int launch_strip(StripArgs a, int n_views, bool window, cudaStream_t stream) {
  Plan p;
  MIS_REQUIRE(make_plan(a.H, a.W, a.s, &p), MIS_ERR_UNSUPPORTED, "mis_aug_two_view: no strip plan for H=%d W=%d s=%d",
              a.H, a.W, a.s);
  a.nsx = p.nsx; a.nsy = p.nsy; a.rp = p.rp; a.rowbuf = p.rowbuf; a.dbl = p.dbl;
  a.pf_groups = 2;     // measured over 0..16 groups at 96^2 / 224^2 / 256^2: 2 is best everywhere, none costs 6-17 %

  a.off_sched = p.off_sched; a.off_fmask = p.off_fmask; a.off_row = p.off_row; a.off_misc = p.off_misc;
  const int n_planes = n_views * a.C;
  if (p.big == 2) return window ? launch_shape<true, 1024, 1>(a, n_planes, p, stream) : launch_shape<false, 1024, 1>(a, n_planes, p, stream);
  if (p.big == 1) return window ? launch_shape<true, 768, 1>(a, n_planes, p, stream) : launch_shape<false, 768, 1>(a, n_planes, p, stream);
  return window ? launch_shape<true, 448, 2>(a, n_planes, p, stream) : launch_shape<false, 448, 2>(a, n_planes, p, stream);
}

}  // namespace augs
}  // namespace mis
