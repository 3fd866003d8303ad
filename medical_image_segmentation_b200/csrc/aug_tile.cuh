// Host-side interface of the warp-tile K1 variant (aug_tile.cu), used by mis_aug_two_view (aug.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mis_b200.h"

namespace mis {
namespace augt {

struct TileArgs {
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
};

// shapes the tile kernel covers (single channel, s <= 256, at most 5.5x downscaling of the whole slice per axis)
bool tile_supported(int C, int H, int W, int64_t img_stride, int s);
int launch_tile(const TileArgs& a, int n_views, bool window, cudaStream_t stream);

}  // namespace augt
}  // namespace mis
