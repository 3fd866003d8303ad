// Host -> device staging of a batch of slices for K1: only the rows the parameter table touches.
//
// The reference decodes whole slices on the CPU and the augmentation reads a crop window of each (on average 48 % of
// the area per view, lightning_module.py:47-48 RandomResizedCrop(scale=(0.08, 1))).  Per slice the rows needed by all of
// its views form ONE contiguous byte range of the source (full-width rows [lo, hi)); ranges are copied into the same
// offsets of a full-size device buffer, so K1 is unchanged.  Measured on B200 (PCIe Gen5, 52 GB/s): one cudaMemcpyAsync
// per slice is SLOWER than one copy of the whole batch (1024 copies of ~430 KB: 35 GB/s effective, 12.6 ms vs 10.2 ms
// although only 83 % of the bytes move) -- a copy costs ~4 us of set-up, i.e. ~200 KB of wire time.  Neighbouring ranges
// are therefore merged whenever the gap between them is below `min_gap_bytes`; with the default (256 KB) the two-view
// chain on 512x512 slices degenerates to a handful of large copies, while sparse tables (small crops, unused slices)
// skip what they do not need.
#include <vector>

#include "common.cuh"

using namespace mis;

extern "C" int mis_h2d_needed_rows(const uint16_t* src_host, uint16_t* dst_dev, int n_images, int C, int H, int W,
                                   int64_t img_stride, const MisViewParams* params_host, int n_views,
                                   int64_t min_gap_bytes, void* stream, int64_t* bytes_copied) {
  MIS_REQUIRE(src_host && dst_dev && params_host, MIS_ERR_INVALID_ARG, "mis_h2d_needed_rows: null pointer");
  MIS_REQUIRE(n_images > 0 && n_views >= 0 && C > 0 && H > 0 && W > 0 && img_stride >= (int64_t)C * H * W,
              MIS_ERR_INVALID_ARG, "mis_h2d_needed_rows: bad sizes");
  std::vector<int> lo(n_images, H), hi(n_images, 0);
  for (int v = 0; v < n_views; ++v) {
    const MisViewParams& p = params_host[v];
    MIS_REQUIRE(p.img >= 0 && p.img < n_images && p.top >= 0 && p.h > 0 && p.top + p.h <= H, MIS_ERR_INVALID_ARG,
                "mis_h2d_needed_rows: record %d out of range (img %d, rows [%d, %d) of %d)", v, p.img, p.top, p.top + p.h, H);
    if (p.top < lo[p.img]) lo[p.img] = p.top;
    if (p.top + p.h > hi[p.img]) hi[p.img] = p.top + p.h;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t total = 0;
  int64_t run_lo = -1, run_hi = -1;                                // pending merged range, in elements
  auto flush = [&]() -> cudaError_t {
    if (run_hi <= run_lo) return cudaSuccess;
    total += (run_hi - run_lo) * (int64_t)sizeof(uint16_t);
    return cudaMemcpyAsync(dst_dev + run_lo, src_host + run_lo, (size_t)(run_hi - run_lo) * sizeof(uint16_t),
                           cudaMemcpyHostToDevice, st);
  };
  const int64_t gap_elems = min_gap_bytes < 0 ? 0 : min_gap_bytes / (int64_t)sizeof(uint16_t);
  for (int i = 0; i < n_images; ++i) {
    if (hi[i] <= lo[i]) continue;                                  // no view reads this slice
    for (int c = 0; c < C; ++c) {
      const int64_t a = (int64_t)i * img_stride + (int64_t)c * H * W + (int64_t)lo[i] * W;
      const int64_t b = (int64_t)i * img_stride + (int64_t)c * H * W + (int64_t)hi[i] * W;
      if (run_hi >= 0 && a - run_hi <= gap_elems) {
        run_hi = b;                                                // close enough: one copy, the gap rides along
      } else {
        MIS_CUDA_TRY(flush());
        run_lo = a;
        run_hi = b;
      }
    }
  }
  MIS_CUDA_TRY(flush());
  if (bytes_copied) *bytes_copied = total;
  return MIS_OK;
}
