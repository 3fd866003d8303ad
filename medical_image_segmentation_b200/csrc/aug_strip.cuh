// Host-side interface of the strip K1 variant (aug_strip.cu), used by mis_aug_two_view (aug.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mis_b200.h"

namespace mis {
namespace augs {

struct StripArgs {
  const uint16_t* src;
  int64_t img_stride;
  int C, H, W;
  const MisViewParams* params;
  const int32_t* order;    // launch order (CTA b works on view order[b]) or null
  float win_lo, win_scale;
  float mean[4], inv_std[4];
  void* out;
  int s;
  int out_f32;
  int nsx, nsy, rp;        // strips of 32 output columns, vertical parts, output rows per part
  int rowbuf;              // floats per intermediate-row buffer of a warp
  int dbl;                 // 1: two row buffers per warp (no second warp barrier per output row)
  int pf_groups;           // L2 prefetch distance of the streams, in groups of G source rows
  int raw_all;             // 1 (C == 3): resample + flip only; every plane is left as uint16 for mis rgb colour kernel
  uint32_t off_sched, off_fmask, off_row, off_misc;   // byte offsets into dynamic shared memory (parked tiles at 0)
};

// shapes the strip kernel covers: 1 or 3 channels (3: resample + flip only, crop a multiple of 8), 8 <= s <= 256, even
// W / img_stride, at most 5.5x downscaling of the whole slice per axis, and the shared-memory plan fits one SM
bool strip_supported(int C, int H, int W, int64_t img_stride, int s);
int launch_strip(StripArgs a, int n_views, bool window, cudaStream_t stream);

}  // namespace augs
}  // namespace mis
