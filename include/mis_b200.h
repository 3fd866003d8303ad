/*
 * mis_b200.h -- C ABI of the B200-native SSL hot path (libmis_b200.so).
 *
 * The reference (EthanHaque/medical_image_segmentation) is pure Python and has no FFI or
 * operator registry for this path: its "interface" is two Python callables,
 *   BYOLRGBDataTransforms.__call__(x) -> [view1, view2]
 *       medical_image_segmentation/train/data_loaders/lightning_module.py:39-64
 *   BYOL.cosine_similarity_loss(preds, targets) -> scalar
 *       medical_image_segmentation/train/model/byol_pytorch.py:181-198 (called :217)
 * The entry points below are what a ctypes binding of those two call sites binds
 * (INTEGRATION.md shows the stub); each one names the reference arithmetic it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a raw device pointer unless it says "host";
 *   - the caller allocates and owns every buffer (including scratch);
 *   - every launch goes to the caller's stream (a cudaStream_t / CUstream passed as void*),
 *     asynchronously, with no internal synchronisation;
 *   - return value 0 = success, otherwise one of MIS_ERR_*; mis_last_error() gives the text
 *     (thread-local).  Nothing throws across the ABI.
 */
#ifndef MIS_B200_H_
#define MIS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIS_ABI_VERSION 1

enum {
  MIS_OK = 0,
  MIS_ERR_INVALID_ARG = 1,   /* null pointer, non-positive size, bad enum            */
  MIS_ERR_UNSUPPORTED = 2,   /* valid request outside what the kernels implement     */
  MIS_ERR_CUDA = 3           /* a CUDA runtime / driver call failed                  */
};

enum { MIS_DTYPE_BF16 = 0, MIS_DTYPE_F32 = 1 };

/* flags of MisViewParams */
#define MIS_VIEW_FLIP     1u   /* RandomHorizontalFlip fired   (torchvision v2/_transform.py:181)   */
#define MIS_VIEW_JITTER   2u   /* RandomApply(ColorJitter) fired (v2/_container.py:104)             */
#define MIS_VIEW_GRAY     4u   /* RandomGrayscale fired (v2/_color.py:33-55); identity for C == 1   */
#define MIS_VIEW_BLUR     8u   /* RandomApply([GaussianBlur(23)]) fired; sigma in blur_sigma        */
#define MIS_VIEW_SOLARIZE 16u  /* RandomSolarize fired (v2/_color.py:312-337)                       */

/* Solarize threshold of the reference chain, RandomSolarize(128) on the 0..255 scale
 * (lightning_module.py:54), expressed on [0,1]: x >= 128/255 -> 1 - x (functional/_color.py:497-501). */
#define MIS_SOLARIZE_THRESHOLD (128.0f / 255.0f)

/*
 * One augmented view = one record (48 bytes).  Filled on the host by the RNG replay of
 * torchvision's draw order (RandomResizedCrop.make_params v2/_geometry.py:272-308,
 * ColorJitter.make_params v2/_color.py:146-154) -- see mis_draw_two_view_params().
 */
typedef struct MisViewParams {
  int32_t img;            /* index of the source slice in the batch                         */
  int32_t top, left;      /* crop box origin  (rows, cols)                                  */
  int32_t h, w;           /* crop box size                                                  */
  uint32_t flags;         /* MIS_VIEW_*                                                     */
  uint8_t order[4];       /* ColorJitter fn_idx: 0 brightness, 1 contrast, 2 sat, 3 hue     */
  float brightness;       /* factors; meaningful only when MIS_VIEW_JITTER is set           */
  float contrast;
  float saturation;       /* drawn (RNG parity) but identity for C == 1                     */
  float hue;
  float blur_sigma;       /* GaussianBlur sigma ~ U(0.1, 2) (v2/_misc.py:209-211); with MIS_VIEW_BLUR */
} MisViewParams;

int mis_version(void);
const char* mis_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Host-side RNG replay (no GPU work).
 *
 * Replaces the per-sample Python RNG calls of the transform chain
 * (lightning_module.py:47-61 -> torchvision make_params / RandomApply / _RandomApplyTransform):
 * consumes a torch CPU mt19937 generator state exactly as `n_images` calls of
 * BYOLRGBDataTransforms.__call__ would (view 1 completely, then view 2) and writes
 * 2*n_images records, image-major: out[2*i + v].
 *
 *   rng_state      host, the byte blob of torch.get_rng_state() (read and advanced in place)
 *   rng_state_len  its length (5056 for torch's CPUGeneratorImpl)
 *   blur_prob / solarize_prob  host float[2]: per-view probabilities of RandomApply([GaussianBlur(23)]) and
 *                  RandomSolarize(128) (lightning_module.py:53-54; reference defaults (1.0, 0.1) / (0.0, 0.2))
 *   n_done         host, out: number of images completed.  n_done < n_images means image
 *                  img0+n_done has a crop box that depends on the last bit of torch.exp (SLEEF
 *                  vs libm expf); the generator is left at that image's first draw and the caller
 *                  draws that one image with torch itself, then calls again for the rest.
 * ------------------------------------------------------------------------------------------ */
int mis_draw_two_view_params(uint8_t* rng_state, int64_t rng_state_len, int n_images, int img0,
                             int H, int W, const float* blur_prob, const float* solarize_prob,
                             MisViewParams* out, int* n_done);

/* The same draw with a callback for the one value this library cannot restate bit for bit: when a crop box depends on
 * the last bit of the reference's float32 torch.exp, exp_f32(x, exp_ctx) is asked for it (the caller evaluates torch.exp
 * on a one-element float32 tensor, exactly the reference's call, v2/_geometry.py:284) and the draw carries on -- *n_done
 * is then always n_images.  exp_f32 == NULL gives the hand-back protocol of mis_draw_two_view_params. */
int mis_draw_two_view_params_cb(uint8_t* rng_state, int64_t rng_state_len, int n_images, int img0,
                                int H, int W, const float* blur_prob, const float* solarize_prob,
                                MisViewParams* out, int* n_done, float (*exp_f32)(float, void*), void* exp_ctx);

/* Reorders the records of mis_draw_two_view_params from image-major [2*i + v] to view-major [v*n_images + i], the row
 * order of torch.cat([view1, view2]) (byol_pytorch.py:207) that mis_aug_two_view's output planes follow.  Host only. */
int mis_params_to_view_major(const MisViewParams* in, int n_images, MisViewParams* out);

/* Host-side check of a table before it is handed to mis_aug_two_view (the kernel trusts it): every record must address a
 * slice 0 <= img < n_images and a box inside H x W, and a record with MIS_VIEW_BLUR a finite positive blur_sigma.
 * *bad_index = first offending record or -1; *flags_or = OR of the flag words (of the records before bad_index). */
int mis_view_params_check(const MisViewParams* params, int n_views, int n_images, int H, int W, uint32_t* flags_or,
                          int* bad_index);

/* Single-view "Resize((s,s)) + ColorJitter(brightness, contrast)" records (the Decathlon flavour of the chain,
 * lightning_module.py:684-693): box = whole image, no flip, jitter always applied; consumes randperm(4) and one
 * uniform per non-zero magnitude per image exactly like torchvision's ColorJitter.make_params
 * (v2/_color.py:146-154).  out[i] for i < n_images. */
int mis_draw_resize_jitter_params(uint8_t* rng_state, int64_t rng_state_len, int n_images, int img0,
                                  int H, int W, float brightness, float contrast, MisViewParams* out);

/* ------------------------------------------------------------------------------------------
 * Fused two-view augmentation (kernel K1).
 *
 * Replaces, per view, the chain RandomResizedCrop (crop_image + antialiased bilinear
 * resize_image, torchvision v2/functional/_geometry.py:1785-1800, 271-340) ->
 * horizontal flip (:56-57) -> ColorJitter brightness/contrast in fn_idx order
 * (functional/_color.py:114-125, 190-205, _blend :92-97) -> ToDtype(float32, scale=True)
 * (functional/_misc.py:304) -> Normalize (functional/_misc.py:37-67), i.e. lightning_module.py:47-58.
 * RandomSolarize (MIS_VIEW_SOLARIZE) is applied in the store epilogue of variant 0; a view with MIS_VIEW_BLUR is left
 * for mis_aug_blur_views (below), which must follow on the same stream.  Variants 1-3 implement neither.
 *
 *   src        uint16 [n_images, C, H, W], plane stride H*W, image stride img_stride elements.
 *              W must be even; the allocation must be readable up to the next 16-byte boundary
 *              past its last element (any cudaMalloc / torch allocation is).
 *   params     device array of n_views records (params[v].img selects the source slice)
 *   win_lo, win_hi   intensity window: x = clamp((u16 - lo) / (hi - lo), 0, 1); (0, 65535) is
 *              the reference's plain 1/65535 scaling
 *   mean,std   host float[C]
 *   out        [n_views, C, s, s] in out_dtype (MIS_DTYPE_BF16 or MIS_DTYPE_F32), NCHW
 *   s          output crop size, 8 <= s <= 256
 *   use_tma    kernel variant (the name is historical):
 *              0: strip kernel (csrc/aug_strip.cu; default, fastest measured on B200: one CTA per view, one warp per
 *                 strip of 32 output columns and vertical part, rows straight from global memory through
 *                 refill-on-consume register slots behind an L2 prefetch, the pre-colour tile parked as uint16 in
 *                 shared memory) for C == 1, s <= 256 and at most 5.5x downscaling of the whole slice per axis;
 *                 other shapes fall through to variant 3, then 2
 *              3: warp-tile kernel (csrc/aug_tile.cu, round 1: cluster per view, one warp per 32x32 output tile)
 *              1: band kernel, a producer warp stages crop rows with 2-D TMA tensor-map boxes (cp.async.bulk.tensor)
 *                 into a shared-memory ring (needs W % 8 == 0 and dense images, otherwise variant 2)
 *              2: band kernel, every thread stages its own columns with cp.async into a private ring, two 16-row
 *                 sub-bands per band
 * ------------------------------------------------------------------------------------------ */
int mis_aug_two_view(const uint16_t* src, int n_images, int C, int H, int W, int64_t img_stride,
                     const MisViewParams* params, int n_views, float win_lo, float win_hi,
                     const float* mean, const float* std, void* out, int s, int out_dtype,
                     int use_tma, void* stream);

/* The variant mis_aug_two_view runs for this shape and selector: 0 strip, 3 warp-tile, 1 TMA band, 2 cp.async band. */
/* mis_aug_two_view with an explicit launch order: CTA b of the kernel works on view view_order[b] (a permutation of
 * 0..n_views-1 in device memory; NULL = identity).  The output does not depend on it; mis_view_cost_order (host) returns
 * the most-expensive-first order that lets the SMs drain together.  Honoured by kernel variant 0, ignored by the others. */
int mis_aug_two_view_ordered(const uint16_t* src, int n_images, int C, int H, int W, int64_t img_stride,
                             const MisViewParams* params, int n_views, const int32_t* view_order, float win_lo,
                             float win_hi, const float* mean, const float* std, void* out, int s, int out_dtype,
                             int use_tma, void* stream);
int mis_view_cost_order(const MisViewParams* params_host, int n_views, int32_t* order_host);

/* The whole host side of one batch in one call: mis_view_params_check on the host table -> table + launch order into
 * the caller's pinned block (>= n_views * (sizeof(MisViewParams) + 4) bytes) -> one cudaMemcpyAsync to `staging_dev` ->
 * mis_aug_two_view_ordered -> mis_aug_blur_views when a record carries MIS_VIEW_BLUR.  *bad_index >= 0 (and nothing
 * launched) when a record is invalid; *n_launches = kernels enqueued.  The caller must not reuse the pinned block
 * before the copy has completed (record an event on `stream` after the call). */
int mis_aug_two_view_staged(const uint16_t* src, int n_images, int C, int H, int W, int64_t img_stride,
                            const MisViewParams* params_host, int n_views, void* staging_pinned, void* staging_dev,
                            int64_t staging_bytes, float win_lo, float win_hi, const float* mean, const float* std,
                            void* out, int s, int out_dtype, int use_tma, uint32_t* flags_or, int* bad_index,
                            int* n_launches, void* stream);

int mis_aug_kernel_variant(int C, int H, int W, int64_t img_stride, int s, int use_tma);

/* GaussianBlur(23) + RandomSolarize(128) + Normalize for the views whose record carries MIS_VIEW_BLUR
 * (RandomApply([GaussianBlur(kernel_size=23)], p) -> RandomSolarize(128, p) -> ToDtype -> Normalize,
 * lightning_module.py:53-57; torchvision v2/functional/_misc.py:104-165: softmax kernel, reflect padding).
 * The blur follows the colour jitter, whose clamps do not commute with it: mis_aug_two_view (variant 0) leaves the
 * post-colour image of such a view as uint16 (round(x*65535)) in the first 2*s*s bytes of the view's output plane and
 * this call finishes those planes in place; other views are not touched.  Call it right after mis_aug_two_view on the
 * same stream with the same `out`, params, C, s, out_dtype, mean, std whenever a record has MIS_VIEW_BLUR.
 * s a multiple of 8 in [16, 256] (above 224 the plane is processed as two bands of 128 output rows). */
int mis_aug_blur_views(void* out, int out_dtype, const MisViewParams* params, int n_views, int C, int s,
                       const float* mean, const float* std, void* stream);

/* Host -> device staging for mis_aug_two_view: copies only the full-width rows [lo, hi) of each slice that the slice's
 * records read, into the same offsets of `dst_dev` (the other rows keep whatever they held and are never read by K1
 * for this table).  Neighbouring ranges whose gap is at most `min_gap_bytes` are merged into one cudaMemcpyAsync
 * (a copy costs ~4 us of set-up = ~200 KB of PCIe time on B200; 0 = one copy per slice).  src_host should be pinned.
 * params_host: HOST copy of the table (any order).  bytes_copied (may be NULL): bytes put on the wire. */
int mis_h2d_needed_rows(const uint16_t* src_host, uint16_t* dst_dev, int n_images, int C, int H, int W,
                        int64_t img_stride, const MisViewParams* params_host, int n_views,
                        int64_t min_gap_bytes, void* stream, int64_t* bytes_copied);

/* Algorithmic bytes K1 moves for a host copy of the params table (crop window read + output
 * write).  Pure host arithmetic. */
int64_t mis_aug_algorithmic_bytes(const MisViewParams* params_host, int n_views, int C, int s,
                                  int out_dtype);

/* ------------------------------------------------------------------------------------------
 * NT-Xent (SimCLR InfoNCE) forward / backward over all-gathered embeddings (kernels K2/K3).
 *
 * Slots in where the reference calls its SSL loss (byol_pytorch.py:217); the reference's own
 * loss is BYOL (see mis_byol_loss_*), NT-Xent is specified by the north star.
 *
 * Layout: a rank owns `rows` = 2*B_local consecutive embedding rows [v1_local; v2_local] that
 * sit at row offset row0 of the rank-major all-gathered matrix of `cols` = 2N rows.  The
 * positive of local row i is local row (i + rows/2) mod rows.
 *
 *   mis_ntxent_prep   z [rows, D] (f32 or bf16) -> u = z/max(|z|,1e-12) rounded to TF32 (f32
 *                     container) and rinv = 1/max(|z|,1e-12).  u is what gets all-gathered.
 *   mis_ntxent_fwd    lse_i = log sum_{j != g(i)} exp(<u_i,u_j>/T) over all `cols` columns (tcgen05
 *                     kind::tf32, accumulators in TMEM; S never leaves the SM);
 *                     loss[0] = mean_i (lse_i - <u_i,u_p(i)>/T) over the local rows.  Two launches: the tile kernel
 *                     and a rows kernel (lse, positives, mean; its last block publishes).
 *   mis_ntxent_bwd    dz_local = grad_out[0] * grad_scale * sum_r' dL_r'/dz_local  (SURVEY A.5,
 *                     option L: uses the all-gathered lse instead of a D-wide gradient exchange).
 *                     S is recomputed tile by tile and W = P + P^T (minus the positives) is rounded to TF32.
 *                     D <= 256: W stays in TMEM as the A operand of the second MMA and dU (128 x D fp32) stays in
 *                     TMEM for the whole column walk.  D > 256 (a multiple of 256): the W tiles are written to the
 *                     scratch ([rows, cols] fp32 -- 4 bytes against 2 * D flops per element) and dU = W . U_all is
 *                     one tcgen05 GEMM over all of D, so S is computed once whatever the width.
 *                     grad_out may be NULL (= 1).  dz has z's dtype.
 *
 * `rows` is the number of embedding rows the rank really has (any even number: a per-GPU batch of 96 gives 192).
 * Every rank's block of the gathered matrix is padded to P = mis_ntxent_padded_rows(rows) (the next multiple of 128):
 * u / u_all hold P (resp. world * P) rows, rinv / lse P (world * P) floats, and `cols`, `row0` are in PADDED
 * coordinates (cols = world * P, row0 = rank * P).  The padding rows are zero vectors, masked as columns inside the
 * tile kernels and skipped as rows; z and dz hold the valid rows only.
 * D a multiple of 32 (backward: D <= 256, or a multiple of 256); T >= 0.025.  `scratch` must hold
 * mis_ntxent_scratch_bytes(rows, cols, D) bytes (for D > 256 that includes the rows x cols W matrix: 0.5 GB per rank
 * at 4096 x 32768).
 * ------------------------------------------------------------------------------------------ */
int mis_ntxent_padded_rows(int rows);
int64_t mis_ntxent_scratch_bytes(int rows, int cols, int D);

int mis_ntxent_prep(const void* z, int z_dtype, int rows, int D, float* u, float* rinv, void* stream);

int mis_ntxent_fwd(const float* u_all, int cols, int D, int row0, int rows, float inv_T,
                   float* lse_rows, float* loss, void* scratch, int64_t scratch_bytes, void* stream);

int mis_ntxent_bwd(const float* u_all, const float* lse_all, const void* z_rows, int z_dtype,
                   const float* rinv_rows, int cols, int D, int row0, int rows, float inv_T,
                   float grad_scale, const float* grad_out, void* dz, void* scratch,
                   int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * NT-Xent over NVLink peer memory (one node, 2 <= world <= 8): the two all-gathers of the loss (normalised rows in the
 * forward, log-sum-exp scalars for the backward; reference template: concat_all_gather, train/callback/knn.py:143-144)
 * are fused into the kernels that PRODUCE the data, and the kernels that CONSUME it wait for it themselves.
 * Every rank passes the same-shaped buffers of all ranks (peer pointers of one symmetric allocation):
 *
 *   u_peers0/1[r]    rank r's gathered matrix [world*rows, D] f32, buffer parity 0 / 1
 *   lse_peers0/1[r]  rank r's gathered lse [world*rows] f32, buffer parity 0 / 1
 *   ctl_peers[r]     rank r's control block (256 bytes, zeroed once): uint32 flag[2][8] (slot 0: rows, slot 1: lse;
 *                    entry = source rank, value = epoch), producer counter, EPOCH (forwards completed on that rank),
 *                    abort word
 *
 * forward  = prep kernel (normalise, round to TF32, store the local rows into every rank's matrix, raise
 *            flag[0][rank] on every rank) + the tile kernel, whose TMA producers wait per column tile for the flag of the
 *            rank that owns those rows (tiles are walked starting at the local rows) + a rows kernel that forms lse /
 *            positives / row losses, stores the lse rows into every rank's vector, and whose last block forms the mean
 *            loss, raises flag[1][rank] on every rank and publishes the epoch.
 * backward = transpose + the tile kernel (one launch per 256 columns of D), whose epilogue warps wait once for the lse
 *            flags of all ranks (S tiles are already being computed meanwhile) + the normalisation-Jacobian kernel.
 *            (Folding the row kernels into the tile kernels' last CTAs was tried: the tails -- one CTA of 13 warps per
 *            128 rows chasing L2 round trips -- cost more than the two launches they saved, DESIGN.md.)
 * Epoch and buffer parity are read from the DEVICE-side counter, so both calls can be captured into CUDA graphs and
 * replayed.  Buffers alternate with the epoch's parity; that is sufficient for safe reuse as long as a rank's backward of
 * epoch k is enqueued before its forward of epoch k+1 (loss.py issues the backward kernels right behind the forward's
 * whenever a gradient is required, so this holds by construction; forwards without a backward may follow each other).
 * A consumer waits `timeout_s` seconds for a peer's flag; after that the kernel sets the abort word and TRAPS (the step
 * fails with a CUDA error instead of continuing on stale rows).  Use a time-out of the order of the NCCL watchdog's.
 * rows: any even number (padded per rank as for mis_ntxent_fwd; at most 414 * 128 per rank); every peer buffer is
 * sized with the padded rows; D as for mis_ntxent_fwd / mis_ntxent_bwd.
 * ------------------------------------------------------------------------------------------ */
int mis_ntxent_fwd_peer(const void* z, int z_dtype, int rows, int D, float inv_T, int world, int rank,
                        void* const* u_peers0, void* const* u_peers1, void* const* lse_peers0,
                        void* const* lse_peers1, void* const* ctl_peers, double timeout_s, float* rinv,
                        float* loss, void* scratch, int64_t scratch_bytes, void* stream);

int mis_ntxent_bwd_peer(const void* z_rows, int z_dtype, const float* rinv_rows, int rows, int D, float inv_T,
                        float grad_scale, const float* grad_out, void* dz, int world, int rank,
                        void* const* u_peers0, void* const* u_peers1, void* const* lse_peers0,
                        void* const* lse_peers1, void* const* ctl_peers, double timeout_s, void* scratch,
                        int64_t scratch_bytes, void* stream);

/* Single-rank NT-Xent (cols == rows, row0 == 0): prep, forward and backward with grad_out = 1 in ONE call --
 * the loss slot of byol_pytorch.py:217 when no cross-GPU gather is involved.  loss[0] and dz (z's dtype) are the
 * outputs; `workspace` (256-byte aligned, mis_ntxent_fwd_bwd_workspace_bytes(rows, D) bytes) holds u, rinv, lse
 * and the kernels' scratch and may be reused by the next call on the same stream. */
int64_t mis_ntxent_fwd_bwd_workspace_bytes(int rows, int D);

int mis_ntxent_fwd_bwd(const void* z, int z_dtype, int rows, int D, float inv_T, float* loss, void* dz,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * BYOL cosine loss, fused forward + backward (the loss the reference actually trains with):
 *   loss[0] = 2 - 2 * mean_i <p_i/|p_i|, t_i/|t_i|>        byol_pytorch.py:196-198
 *   dpreds  = dloss/dpreds (may be NULL; targets carry no gradient, byol_pytorch.py:212-214)
 *   scratch_rows: `rows` floats.
 * ------------------------------------------------------------------------------------------ */
int mis_byol_loss_fwd_bwd(const float* preds, const float* targets, int rows, int D, float* loss,
                          float* dpreds, float* scratch_rows, void* stream);

/* ------------------------------------------------------------------------------------------
 * EMA update of the momentum encoder as ONE multi-tensor kernel.
 * Replaces BYOL.momentum_update (byol_pytorch.py:291-296, called at :253-255):
 *     for po, pm in zip(online.parameters(), momentum.parameters()): pm.data.mul_(m).add_(po.data, alpha=1.0 - m)
 * with pm = fma(po, one_minus_m, fl(pm * m)) over a device table of float32 tensors (bit-identical to the two ATen
 * kernels when m = float(m64), one_minus_m = float(1.0 - m64): the reference forms 1 - m in double).  table_dev[i].chunk0 = sum_{j<i} mis_ema_chunks(table[j].n); total_chunks = that sum over all tensors.
 * ------------------------------------------------------------------------------------------ */
typedef struct MisEmaEntry {
  const void* online;     /* float32 [n], device  */
  void* momentum;         /* float32 [n], device, updated in place */
  int64_t n;
  int64_t chunk0;
} MisEmaEntry;

int64_t mis_ema_chunks(int64_t n_elements);
int mis_ema_update(const MisEmaEntry* table_dev, int n_tensors, int64_t total_chunks, float m, float one_minus_m,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weighted kNN prediction of the online evaluator -- KNNOnlineEvaluator.predict (train/callback/knn.py:38-70):
 * sim = query . bank^T (tcgen05 TF32 GEMM over hi/lo-split operands: fp32-grade similarities, fp32 accumulate) ->
 * the k most similar bank rows per query (exact radix
 * select; ties at the k-th similarity go to the lower bank index) -> votes exp(sim / T) summed per class ->
 * pred_labels[b, :] = the classes by descending score (equal scores: ascending class), what `argsort(descending=True)`
 * returns up to the order of equal scores.  pred_scores ([n_query, num_classes], may be NULL) receives the scores.
 * query [n_query, D], bank [n_bank, D]: fp32 row-major, 16-byte aligned, L2-normalised by the caller (knn.py:100,129);
 * bank_labels [n_bank] int64; D a multiple of 32; 1 <= k <= min(n_bank, 1024); num_classes <= 8192.
 * scratch: mis_knn_scratch_bytes(n_query, n_bank, D) bytes (the padded similarity matrix and the split operands).
 * ------------------------------------------------------------------------------------------ */
int64_t mis_knn_scratch_bytes(int n_query, int n_bank, int D);
int mis_knn_predict(const float* query, const float* bank, const int64_t* bank_labels, int n_query, int n_bank, int D,
                    int k, float inv_T, int num_classes, int64_t* pred_labels, float* pred_scores, void* scratch,
                    int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Exact per-channel moments of uint16 slices: the statistics behind the normalisation constants.
 * Replaces the float64 streaming sums of compute_mean_and_std
 * (medical_image_segmentation/analyze_data/compute_dataset_metrics.py:12-29).
 *
 *   src    uint16 [n_images, C, plane_elems] (16-byte aligned, plane_elems % 8 == 0, n_images*C <= 65535)
 *   sums   device uint64 [C][2], ACCUMULATED into (caller zeroes it): sums[c][0] += sum x, sums[c][1] += sum x^2
 * Both sums are exact integers; mean = s0/n, std = sqrt(s1/n - mean^2) are formed by the caller in float64.
 * ------------------------------------------------------------------------------------------ */
int mis_u16_moments(const uint16_t* src, long long n_images, int C, long long plane_elems,
                    unsigned long long* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIS_B200_H_ */
