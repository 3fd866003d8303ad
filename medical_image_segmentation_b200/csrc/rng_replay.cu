// Host-side replay of the random draws of the reference's two-view transform chain.
//
// The reference draws every augmentation parameter with per-sample Python calls into torch's
// global CPU generator (lightning_module.py:47-61 -> torchvision v2: RandomResizedCrop.make_params
// _geometry.py:272-308, _RandomApplyTransform.forward _transform.py:181, RandomApply.forward
// _container.py:104, ColorJitter.make_params _color.py:146-154).  ~20 tensor ops per view is three
// orders of magnitude slower than the GPU consumes views, so the same stream is consumed here
// natively: the caller hands over the bytes of torch.get_rng_state(), this code advances the
// mt19937 exactly as torch would and the caller puts the state back with torch.set_rng_state().
//
// Restated torch CPU generator semantics (ATen CPUGeneratorImpl / DistributionsHelper.h):
//   random()             one tempered mt19937 word
//   uniform_(a,b) f32    fma((random() & (2^24-1)) * 2^-24, b - a, a)          (float, fused)
//   rand(1)              uniform_(0,1)
//   randint(0,n)         random() % n                                         (n < 2^32)
//   randperm(n)          Fisher-Yates: for i < n-1: z = random() % (n-i); swap(r[i], r[i+z])
//
// The one operation that cannot be restated bit-for-bit is torch.exp on a float32 tensor (SLEEF
// vector expf, not libm).  A crop box only depends on it through round(sqrt(area*ratio)); whenever
// the box would differ for exp(x) one ulp up or down, the image is reported back to the caller
// (n_done < n_images) who draws that single image with torch itself -- the result is bit-exact
// always, and native for all but ~3e-4 of the images.  mis_draw_two_view_params_cb instead asks the
// caller for that one float32 exp() through a callback (torch.exp on a one-element tensor, exactly
// the reference's call) and carries on: no image is ever handed back, no 0.6 ms Python replay.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace mis {
namespace rng {

constexpr int kN = 624, kM = 397;
constexpr int64_t kStateLen = 5056;   // sizeof(at::CPUGeneratorImplState)

struct Engine {
  int32_t left;
  uint64_t next;
  uint32_t st[kN];
  uint32_t prev[kN];        // the state array before the last refill (so that a rewind across a refill is cheap)
  uint32_t refills = 0;

  static inline uint32_t twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  void next_state() {
    memcpy(prev, st, sizeof(st));
    ++refills;
    uint32_t* p = st;
    left = kN;
    next = 0;
    for (int j = kN - kM + 1; --j; p++) *p = p[kM] ^ twist(p[0], p[1]);
    for (int j = kM; --j; p++) *p = p[kM - kN] ^ twist(p[0], p[1]);
    *p = p[kM - kN] ^ twist(p[0], st[0]);
  }
  inline uint32_t random() {
    if (--left == 0) next_state();
    uint32_t y = st[next++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  inline float uniform(float from, float to) {
    const float x = (float)(random() & ((1u << 24) - 1)) * (1.0f / 16777216.0f);
    return fmaf(x, to - from, from);   // torch's kernel is built with FMA contraction: one rounding
  }
};

// layout of at::CPUGeneratorImplStateLegacy: seed u64 | left i32 | seeded i32 | next u64 | state u64[624] | ...
static void load(const uint8_t* blob, Engine& e) {
  memcpy(&e.left, blob + 8, 4);
  memcpy(&e.next, blob + 16, 8);
  for (int i = 0; i < kN; ++i) {
    uint64_t v;
    memcpy(&v, blob + 24 + 8 * i, 8);
    e.st[i] = (uint32_t)v;
  }
}
static void store(uint8_t* blob, const Engine& e) {
  memcpy(blob + 8, &e.left, 4);
  memcpy(blob + 16, &e.next, 8);
  for (int i = 0; i < kN; ++i) {
    const uint64_t v = e.st[i];
    memcpy(blob + 24 + 8 * i, &v, 8);
  }
}

// constants of the reference chain (lightning_module.py:44,49-54)
constexpr float kScaleLo = 0.08f, kScaleHi = 1.0f;
constexpr float kLogRatioLo = -0.28768208622932434f;   // torch.log(torch.tensor(3/4))  (float32)
constexpr float kLogRatioHi = 0.28768211603164673f;    // torch.log(torch.tensor(4/3))  (float32)
constexpr double kRatioLo = 3.0 / 4.0, kRatioHi = 4.0 / 3.0;
constexpr float kFlipP = 0.5f, kJitterP = 0.8f, kGrayP = 0.2f;

struct Box {
  int top, left, h, w;
  bool ok;
};

// w = round(sqrt(area*ar)), h = round(sqrt(area/ar)) as Python computes them (double sqrt, round-half-even).
// `safe` is false when either value is so close to a rounding boundary that the last bit of exp() could flip it:
// a one-ulp change of the float32 ratio moves sqrt(...) by < 1e-7 relative, i.e. < 1e-4 px for boxes < 1000 px.
static inline void box_dims(double target_area, float ar, int& w, int& h, bool& safe) {
  const double ws = std::sqrt(target_area * (double)ar);
  const double hs = std::sqrt(target_area / (double)ar);
  w = (int)std::nearbyint(ws);
  h = (int)std::nearbyint(hs);
  const double fw = std::fabs(ws - std::floor(ws) - 0.5), fh = std::fabs(hs - std::floor(hs) - 0.5);
  const double eps = 1e-6 * (ws > hs ? ws : hs) + 1e-9;
  safe = fw > eps && fh > eps;
}

typedef float (*ExpFn)(float, void*);

// returns false when the result depends on the last bit of exp() and no callback is there to supply torch's value
static bool draw_view(Engine& e, int H, int W, float blur_p, float sol_p, MisViewParams& out, ExpFn exp_cb, void* exp_ctx) {
  const double area = (double)H * (double)W;
  bool have = false;
  int top = 0, left = 0, h = 0, w = 0;
  for (int t = 0; t < 10; ++t) {
    const double target_area = area * (double)e.uniform(kScaleLo, kScaleHi);
    const float lr = e.uniform(kLogRatioLo, kLogRatioHi);
    const float ar = expf(lr);
    bool safe;
    box_dims(target_area, ar, w, h, safe);
    if (!safe) {   // rare: decide with the neighbouring float32 ratios; if they disagree, hand the image back
      int w1, h1, w2, h2;
      bool s1, s2;
      // libm expf and SLEEF's vector expf are each within 1 ulp of exp(): they differ by at most 2 ulp
      box_dims(target_area, std::nextafterf(std::nextafterf(ar, 0.f), 0.f), w1, h1, s1);
      box_dims(target_area, std::nextafterf(std::nextafterf(ar, 4.f), 4.f), w2, h2, s2);
      if (w1 != w || w2 != w || h1 != h || h2 != h) {
        if (!exp_cb) return false;
        bool s3;
        box_dims(target_area, exp_cb(lr, exp_ctx), w, h, s3);   // torch's own float32 exp of this draw decides
      }
    }
    if (0 < w && w <= W && 0 < h && h <= H) {
      top = (int)(e.random() % (uint32_t)(H - h + 1));
      left = (int)(e.random() % (uint32_t)(W - w + 1));
      have = true;
      break;
    }
  }
  if (!have) {   // central-crop fallback, _geometry.py:293-306
    const double in_ratio = (double)W / (double)H;
    if (in_ratio < kRatioLo) {
      w = W;
      h = (int)std::nearbyint((double)w / kRatioLo);
    } else if (in_ratio > kRatioHi) {
      h = H;
      w = (int)std::nearbyint((double)h * kRatioHi);
    } else {
      w = W;
      h = H;
    }
    top = (H - h) / 2;
    left = (W - w) / 2;
  }
  out.top = top;
  out.left = left;
  out.h = h;
  out.w = w;
  out.flags = 0;
  out.order[0] = 0; out.order[1] = 1; out.order[2] = 2; out.order[3] = 3;
  out.brightness = 1.f; out.contrast = 1.f; out.saturation = 1.f; out.hue = 0.f;
  out.blur_sigma = 0.f;
  if (!(e.uniform(0.f, 1.f) >= kFlipP)) out.flags |= MIS_VIEW_FLIP;
  if (!(e.uniform(0.f, 1.f) >= kJitterP)) {
    out.flags |= MIS_VIEW_JITTER;
    uint8_t perm[4] = {0, 1, 2, 3};
    for (int i = 0; i < 3; ++i) {
      const uint32_t z = e.random() % (uint32_t)(4 - i);
      const uint8_t sav = perm[i];
      perm[i] = perm[i + z];
      perm[i + z] = sav;
    }
    memcpy(out.order, perm, 4);
    out.brightness = e.uniform(0.6f, 1.4f);
    out.contrast = e.uniform(0.6f, 1.4f);
    out.saturation = e.uniform(0.8f, 1.2f);
    out.hue = e.uniform(-0.1f, 0.1f);
  }
  if (!(e.uniform(0.f, 1.f) >= kGrayP)) out.flags |= MIS_VIEW_GRAY;       // RandomGrayscale(p=0.2): identity at C == 1
  if (!(e.uniform(0.f, 1.f) >= blur_p)) {                                 // RandomApply([GaussianBlur(23)])
    out.flags |= MIS_VIEW_BLUR;
    out.blur_sigma = e.uniform(0.1f, 2.0f);                               //   sigma draw, v2/_misc.py:209-211
  }
  if (!(e.uniform(0.f, 1.f) >= sol_p)) out.flags |= MIS_VIEW_SOLARIZE;    // RandomSolarize(128)
  return true;
}

}  // namespace rng
}  // namespace mis

extern "C" int mis_draw_two_view_params_cb(uint8_t* rng_state, int64_t rng_state_len, int n_images, int img0, int H, int W,
                                           const float* blur_prob, const float* solarize_prob, MisViewParams* out,
                                           int* n_done, float (*exp_f32)(float, void*), void* exp_ctx) {
  using namespace mis;
  using namespace mis::rng;
  MIS_REQUIRE(rng_state && out && blur_prob && solarize_prob && n_done, MIS_ERR_INVALID_ARG,
              "mis_draw_two_view_params: null pointer");
  MIS_REQUIRE(rng_state_len == kStateLen, MIS_ERR_INVALID_ARG,
              "mis_draw_two_view_params: rng state is %lld bytes, expected %lld (torch CPUGeneratorImpl)",
              (long long)rng_state_len, (long long)kStateLen);
  MIS_REQUIRE(n_images >= 0 && H > 0 && W > 0, MIS_ERR_INVALID_ARG, "mis_draw_two_view_params: bad sizes");
  for (int v = 0; v < 2; ++v)
    MIS_REQUIRE(blur_prob[v] >= 0.f && blur_prob[v] <= 1.f && solarize_prob[v] >= 0.f && solarize_prob[v] <= 1.f,
                MIS_ERR_INVALID_ARG, "mis_draw_two_view_params: probabilities must be in [0,1]");
  Engine e;
  load(rng_state, e);
  MIS_REQUIRE(e.left >= 1 && e.left <= kN && e.next <= (uint64_t)kN, MIS_ERR_INVALID_ARG,
              "mis_draw_two_view_params: corrupt generator state (left=%d next=%llu)", e.left,
              (unsigned long long)e.next);
  *n_done = 0;
  for (int i = 0; i < n_images; ++i) {
    // checkpoint = the two cursors; the state array itself only changes at a refill, which keeps its predecessor
    const int32_t ck_left = e.left;
    const uint64_t ck_next = e.next;
    const uint32_t ck_refills = e.refills;
    bool exact = true;
    for (int v = 0; v < 2 && exact; ++v) {
      MisViewParams& p = out[2 * (size_t)i + v];
      p.img = img0 + i;
      exact = draw_view(e, H, W, blur_prob[v], solarize_prob[v], p, exp_f32, exp_ctx);
    }
    if (!exact) {   // hand this image back to the caller, generator rewound to its first draw
      // (an image draws at most 2 * (10 * 4 + 13) words, far fewer than the 624 of a refill: at most one refill to undo)
      if (e.refills != ck_refills) memcpy(e.st, e.prev, sizeof(e.st));
      e.left = ck_left;
      e.next = ck_next;
      break;
    }
    *n_done = i + 1;
  }
  store(rng_state, e);
  return MIS_OK;
}

extern "C" int mis_draw_two_view_params(uint8_t* rng_state, int64_t rng_state_len, int n_images, int img0, int H, int W,
                                        const float* blur_prob, const float* solarize_prob, MisViewParams* out,
                                        int* n_done) {
  return mis_draw_two_view_params_cb(rng_state, rng_state_len, n_images, img0, H, W, blur_prob, solarize_prob, out, n_done,
                                     nullptr, nullptr);
}

// Single-view "Resize + ColorJitter(brightness, contrast)" parameters: the Decathlon flavour of the chain
// (lightning_module.py:684-693: Resize((s,s)) -> ColorJitter(brightness=0.2, contrast=0.2) -> ToDtype -> Normalize).
// torchvision ColorJitter.make_params (v2/_color.py:146-154): randperm(4), then one uniform per non-None factor;
// a magnitude of 0 makes the factor None (no draw), v2/_color.py _check_input.
extern "C" int mis_draw_resize_jitter_params(uint8_t* rng_state, int64_t rng_state_len, int n_images, int img0, int H, int W,
                                             float brightness, float contrast, MisViewParams* out) {
  using namespace mis;
  using namespace mis::rng;
  MIS_REQUIRE(rng_state && out, MIS_ERR_INVALID_ARG, "mis_draw_resize_jitter_params: null pointer");
  MIS_REQUIRE(rng_state_len == kStateLen, MIS_ERR_INVALID_ARG,
              "mis_draw_resize_jitter_params: rng state is %lld bytes, expected %lld", (long long)rng_state_len,
              (long long)kStateLen);
  MIS_REQUIRE(n_images >= 0 && H > 0 && W > 0 && brightness >= 0.f && contrast >= 0.f, MIS_ERR_INVALID_ARG,
              "mis_draw_resize_jitter_params: bad arguments");
  Engine e;
  load(rng_state, e);
  for (int i = 0; i < n_images; ++i) {
    MisViewParams& p = out[i];
    p.img = img0 + i;
    p.top = 0;
    p.left = 0;
    p.h = H;
    p.w = W;
    p.flags = MIS_VIEW_JITTER;
    uint8_t perm[4] = {0, 1, 2, 3};
    for (int k = 0; k < 3; ++k) {
      const uint32_t z = e.random() % (uint32_t)(4 - k);
      const uint8_t sav = perm[k];
      perm[k] = perm[k + z];
      perm[k + z] = sav;
    }
    memcpy(p.order, perm, 4);
    const float blo = 1.f - brightness < 0.f ? 0.f : 1.f - brightness;
    p.brightness = brightness > 0.f ? e.uniform(blo, 1.f + brightness) : 1.f;
    const float clo = 1.f - contrast < 0.f ? 0.f : 1.f - contrast;
    p.contrast = contrast > 0.f ? e.uniform(clo, 1.f + contrast) : 1.f;
    p.saturation = 1.f;
    p.hue = 0.f;
    p.blur_sigma = 0.f;
  }
  store(rng_state, e);
  return MIS_OK;
}

// [2*i + v] (image-major, the draw order) -> [v*n_images + i] (view-major: the row order of cat([view1, view2]),
// byol_pytorch.py:207).  Plain host copy of 48-byte records.
extern "C" int mis_params_to_view_major(const MisViewParams* in, int n_images, MisViewParams* out) {
  MIS_REQUIRE(in && out && n_images >= 0 && in != out, MIS_ERR_INVALID_ARG, "mis_params_to_view_major: bad arguments");
  for (int i = 0; i < n_images; ++i) {
    out[i] = in[2 * i];
    out[n_images + i] = in[2 * i + 1];
  }
  return MIS_OK;
}

// K1 trusts its table; this is the host-side check of one (a few microseconds for thousands of records, in place of a
// dozen numpy passes): every record addresses a slice of the batch and a box inside it, a blurred view has a usable sigma.
extern "C" int mis_view_params_check(const MisViewParams* p, int n_views, int n_images, int H, int W, uint32_t* flags_or,
                                     int* bad_index) {
  MIS_REQUIRE((p || n_views == 0) && n_views >= 0 && flags_or && bad_index, MIS_ERR_INVALID_ARG,
              "mis_view_params_check: bad arguments");
  uint32_t f = 0;
  *bad_index = -1;
  for (int k = 0; k < n_views; ++k) {
    const MisViewParams& r = p[k];
    const bool box_ok = r.img >= 0 && r.img < n_images && r.top >= 0 && r.left >= 0 && r.h >= 1 && r.w >= 1 &&
                        (int64_t)r.top + r.h <= H && (int64_t)r.left + r.w <= W;
    const bool blur_ok = !(r.flags & MIS_VIEW_BLUR) || (r.blur_sigma > 0.f && r.blur_sigma <= 3.4e38f);
    if (!box_ok || !blur_ok) {
      *bad_index = k;
      *flags_or = f;
      return MIS_OK;
    }
    f |= r.flags;
  }
  *flags_or = f;
  return MIS_OK;
}

// Launch order of the views of one K1 call: the most expensive first (cost ~ crop area h * w, which is what the vertical
// pass reads), so that the last CTAs the hardware scheduler hands out are the cheap ones and the SMs drain together
// (longest-processing-time-first; measured -10 % at 512 slices per GPU, -3.5 % at 4096).  Stable counting sort, host only.
extern "C" int mis_view_cost_order(const MisViewParams* p, int n_views, int32_t* order) {
  MIS_REQUIRE((p && order) || n_views == 0, MIS_ERR_INVALID_ARG, "mis_view_cost_order: null pointer");
  MIS_REQUIRE(n_views >= 0, MIS_ERR_INVALID_ARG, "mis_view_cost_order: n_views < 0");
  constexpr int kBuckets = 1024;
  int64_t amax = 1;
  for (int k = 0; k < n_views; ++k) {
    const int64_t a = (int64_t)p[k].h * p[k].w;
    if (a > amax) amax = a;
  }
  int count[kBuckets + 1] = {};
  auto bucket = [&](int k) {
    int64_t a = (int64_t)p[k].h * p[k].w;
    if (a < 0) a = 0;
    return (kBuckets - 1) - (int)(a * (kBuckets - 1) / amax);      // large area -> small bucket index
  };
  for (int k = 0; k < n_views; ++k) ++count[bucket(k) + 1];
  for (int b = 0; b < kBuckets; ++b) count[b + 1] += count[b];
  for (int k = 0; k < n_views; ++k) order[count[bucket(k)]++] = k;
  return MIS_OK;
}
