// K2/K3 -- NT-Xent (SimCLR InfoNCE) forward and backward on tcgen05 tensor cores (sm_100a).
//
// The 2N x 2N similarity matrix is never written to HBM.  One CTA owns a 128-row tile of the
// local rows and walks a range of 128-column tiles of the all-gathered matrix:
//
//   warps 0, 10-12  TMA producers: 128B-swizzled K-major boxes of U (and, for the backward, of U^T) into a
//            shared-memory ring, one mbarrier pair per stage; in a multi-rank forward they wait, per column
//            tile, for the flag of the rank that owns those rows (stored over NVLink by its prep kernel);
//   warp 1   MMA issuer (one thread): S = U_I . U_J^T with tcgen05.mma kind::tf32, accumulator in
//            TMEM (double buffered, 2 x 128 columns); backward, D <= 256: dU_I += W . U_J with the A
//            operand W read straight from TMEM and dU_I (128 x D fp32) resident in TMEM columns
//            256.. for the whole column walk;
//   warps 2-9 two epilogue groups (alternate tiles), lane == row (tcgen05.ld 32x32b): exp2 with the fixed max
//            1/T, diagonal / padding masks; forward: thread-local row sums (no shuffles);  backward:
//            W_ij = E_ij (c_i + c_j) rounded to TF32 and stored back over S with tcgen05.st, so the tile never
//            leaves the SM -- or, for D > 256 (kWOut), written to HBM once, after which wu_gemm_kernel computes
//            dU = W . U_all as one tcgen05 GEMM over all of D (no per-slice recompute of S).
// Small kernels around them: prep (normalise, TF32 rounding, peer stores), fwd_rows (lse, positives, mean loss, lse peer
// stores + flags), transpose (U^T for the K-major boxes of the backward), bwd_finalize (split sum + normalisation Jacobian).
//
// Math (oracle/loss_oracle.py):  u = z/max(|z|,1e-12), S = u u^T / T,
//   lse_i = log sum_{j != i} exp S_ij,   L_r = mean_{i in rank r} (lse_i - S_{i,p(i)}),
//   dU_i = (g/(T rows)) [ sum_j (exp(S_ij - lse_i) + exp(S_ij - lse_j)) u_j  -  2 u_p(i) ],
//   dz_i = (dU_i - u_i <u_i,dU_i>) / max(|z_i|,1e-12).
// With E_ij = exp(S_ij - 1/T) and c_j = exp(1/T - lse_j):  exp(S_ij - lse_i) + exp(S_ij - lse_j)
// = E_ij (c_i + c_j): one exp per element.  S_ij <= 1/T always, so 1/T is a valid fixed maximum
// (requires exp(-2/T) to stay normal in fp32: T >= 0.025).
//
// Operand precision: U is rounded to TF32 (round-to-nearest) once by the prep kernel, products
// accumulate in fp32; positives, the one-hot term and the normalisation Jacobian are done in
// plain fp32.  Measured gradient error vs the fp64 oracle ~2e-4 relative (bf16 operands would
// give 1.6e-3 and miss the 1e-3 gate).
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "common.cuh"

namespace mis {
namespace ntx {

constexpr int kTile = 128;             // rows per CTA tile, columns per S tile
constexpr int kKBlock = 32;            // fp32 elements per 128-byte swizzle row
constexpr int kRingBytes = 192 * 1024;  // operand ring (+ resident row tile); see smem plan in the kernel
constexpr int kMaxStages = 8;
constexpr int kThreads = 416;           // warps 0, 10, 11, 12: TMA producers; warp 1: MMA issuer; warps 2-5 / 6-9: epilogue groups
constexpr int kProducers = 4;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kDuCol = 256;            // TMEM column where the dU accumulator starts
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kMaxPeers = 8;

// Control block at the start of every rank's symmetric (peer-mapped) buffer.  flag[slot][r] = epoch is raised by
// rank r (st.release.sys) when its rows (slot 0) / log-sum-exp scalars (slot 1) of that epoch have landed here.
struct PeerCtl {
  uint32_t flag[2][kMaxPeers];
  uint32_t prep_counter;     // CTAs of the prep kernel that have finished (self-resetting)
  uint32_t epoch;            // forwards completed on this rank: kernels derive epoch and buffer parity from it, so
                             // a captured CUDA graph needs no host-side value
  uint32_t abort;            // set before a kernel traps on a peer time-out
  uint32_t pad;
};

// arguments of the kernel that finishes the forward rows (lse, positives, row losses, mean) and publishes them
struct RowsArgs {
  const float* partial;      // [nparts][rows] row-sum partials of the tile kernel
  int nparts, rows, rows_valid, D, row0;
  float inv_T;
  const float* u_all[2];     // per parity: gathered matrix (positives)
  float* lse_dst[2][kMaxPeers];  // per parity, per rank: gathered lse vector (dst[.][rank] is the local copy)
  int lse_off;               // row offset of this rank inside the gathered vectors
  float* row_loss;           // [rows] scratch
  float* loss;               // [1]
  unsigned int* counter;     // blocks that have finished (zero on entry, left zero)
  PeerCtl* ctl;              // local control block, null for a single rank
  PeerCtl* ctl_peers[kMaxPeers];
  int world, rank;
};

struct TileArgs {
  int rows, cols, D, row0;   // rows / cols / row0 are PADDED to multiples of 128
  int rows_valid;            // embedding rows a rank really has; rows [rows_valid, rows) of every rank's block are padding:
                             // masked as columns, skipped as rows
  int col_tiles, tiles_per_split;
  int split_fast;          // 1: grid = (splits, row tiles) -- the column splits of a row tile are co-scheduled, so fewer row
                           // tiles are live at a time (wide embeddings: their streamed operands then stay inside the L2)
  float k1;                // log2(e) / T
  int d0, ds;              // backward: columns [d0, d0 + ds) of dU are produced by this launch (ds <= 256)
  const float* lse[2];     // per parity: [cols] all-gathered log-sum-exp (backward): c_j = exp(1/T - lse_j) on the fly
  float inv_T;
  float* partial;          // fwd: [nsplit][rows]   bwd: [nsplit][rows][D]
  float* w_out;            // backward for D > 256 (kWOut): W = P + P^T - positives, TF32-rounded, [rows][cols] row-major
  // multi-rank exchange (world == 1: ctl == nullptr, parity 0)
  PeerCtl* ctl;            // local control block
  PeerCtl* ctl_peers[kMaxPeers];
  int world, rank, rows_per_rank;
  int epoch_add;           // forward: 1 (the epoch being produced), backward: 0 (the epoch completed last)
  long long timeout_clk;   // SM clocks a consumer waits for a peer before it traps
};

struct Bars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t a_full;          // resident row tile U_I has landed
  uint64_t tmem_full[2];
  uint64_t epi_done[2];
  uint64_t du_full;
  uint32_t tmem_ptr;
  uint32_t pad;
  float cj[2][kTile];
};

// ---- tcgen05 / TMA wrappers -----------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);        // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset       bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// fp32 accumulate, TF32 x TF32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until a peer has raised `*f` to `epoch`.  A peer that is merely late (checkpointing, a data-loader stall) is
// waited for like NCCL would; after `timeout_clk` SM clocks the rank is declared dead: the kernel records it in
// ctl->abort and TRAPS, so the step fails with a CUDA error instead of continuing on stale rows.
__device__ __forceinline__ void wait_peer_flag(const uint32_t* f, uint32_t epoch, long long timeout_clk, uint32_t* abort) {
  if ((int32_t)(ld_acquire_sys(f) - epoch) >= 0) return;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
    __nanosleep(100);
    if (clock64() - t0 > timeout_clk) {
      *abort = 1u;
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ float ldcg_f32(const float* p) { return __ldcg(p); }

// kWOut (backward, D > 256): the W tiles are written to HBM instead of feeding the dU MMAs from TMEM; dU = W . U_all is
// then one plain GEMM (wu_gemm_kernel) -- at D = 2048 a W element costs 4 bytes of traffic against 2 * D flops, so
// materialising it is cheap, whereas walking dU in 256-column slices recomputed S once per slice.
template <bool kBwd, bool kWOut = false>
__global__ void __launch_bounds__(kThreads, 1)
ntxent_tile_kernel(const __grid_constant__ CUtensorMap map_u0, const __grid_constant__ CUtensorMap map_u1,
                   const __grid_constant__ CUtensorMap map_ut, const __grid_constant__ TileArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars& bars = *reinterpret_cast<Bars*>(smem + kRingBytes);
  // smem plan (192 KB).  D <= 128: the row tile U_I stays RESIDENT (D/32 k-blocks of 16 KB, loaded once) and the ring
  // has 2 stages of 64 KB, each a WHOLE operand tile (all k-blocks of U_J, or all four U^T boxes): the MMA thread then
  // issues 16 MMAs (~1k clk of tensor work) per barrier round trip -- with one 4-MMA k-block per stage its own
  // wait/commit loop (~700 clk) left the tensor pipe idle 60 % of the time.  Wider D: 6 stages of 32 KB holding one
  // (U_I, U_J) k-block pair or one U^T box of 256 rows.
  const bool res_a = a.D <= 128;
  const int kStages = res_a ? 2 : 6;
  const int kStageBytes = res_a ? 64 * 1024 : 32 * 1024;
  const int kps_s = res_a ? a.D / kKBlock : 1;                   // k-blocks per stage, S phase
  const int kps_g = res_a ? kTile / kKBlock : 1;                 // k-blocks per stage, dU phase
  uint8_t* const smem_a = smem;                                  // resident U_I (only when res_a)
  uint8_t* const ring = smem + (res_a ? 64 * 1024 : 0);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row_tile = a.split_fast ? blockIdx.y : blockIdx.x, split = a.split_fast ? blockIdx.x : blockIdx.y;
  const int t_begin = split * a.tiles_per_split;
  const int T = min(a.tiles_per_split, a.col_tiles - t_begin);
  const int KB = a.D / kKBlock;
  const int g_row_tile0 = a.row0 + row_tile * kTile;   // global row index of this tile's first row
  // epoch / buffer parity of this evaluation (single rank: one buffer, nothing to wait for)
  const uint32_t epoch = a.ctl ? a.ctl->epoch + (uint32_t)a.epoch_add : 0u;
  const int par = (int)(epoch & 1u);
  const CUtensorMap* const map_u = par ? &map_u1 : &map_u0;
  // column tiles are walked starting at this rank's own rows, so the tiles of late peers come last
  const int rot = a.world > 1 ? (a.rank * a.rows_per_rank) / kTile : 0;
  auto col_tile = [&](int t) {
    int c = t_begin + t + rot;
    return c >= a.col_tiles ? c - a.col_tiles : c;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(map_u);
    if (kBwd && !kWOut) prefetch_tmap(&map_ut);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kMaxStages; ++i) {
        mbar_init(&bars.full[i], 1);
        mbar_init(&bars.empty[i], 1);
      }
      mbar_init(&bars.a_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars.tmem_full[i], 1);
        mbar_init(&bars.epi_done[i], kEpiThreads);
      }
      mbar_init(&bars.du_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_ptr)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_ptr;

  if (warp == 0 || warp >= 10) {
    // ===================================== TMA producers =====================================
    // One thread sustains only ~1 TMA box per 455 clk (measured, scripts/micro/tma_lat.cu) and a stage is refilled
    // only after its MMAs retire, so the refill latency (boxes x 455 clk + memory latency) bounds the tile rate.  Four
    // producer threads in different warps therefore walk the same stage sequence and each issues one box of every
    // stage; producer 0 also posts the stage's arrive.expect_tx.
    if (lane == 0) {
      const int pidx = warp == 0 ? 0 : warp - 9;               // 0..3
      int stage = 0, phase = 0;
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      if (res_a) {    // resident row tile, loaded once
        if (pidx == 0) mbar_arrive_expect_tx(&bars.a_full, (uint32_t)KB * kTile * 128);
        for (int kb = pidx; kb < KB; kb += kProducers)
          tma_load_2d(smem_a + kb * kTile * 128, map_u, kb * kKBlock, g_row_tile0, &bars.a_full);
      }
      // forward of a multi-rank evaluation: the rows of rank r are complete once flag[0][r] carries this epoch
      // (written by r's prep kernel over NVLink); every producer thread checks before its first box of that rank
      uint32_t ready = (!kBwd && a.world > 1) ? (1u << a.rank) : 0xffffffffu;
      auto push_s = [&](int t) {   // operands of S = U_I . U_J^T
        const int ct = col_tile(t);
        if (!kBwd && a.world > 1) {
          const int r = (ct * kTile) / a.rows_per_rank;
          if (!((ready >> r) & 1u)) {
            wait_peer_flag(&a.ctl->flag[0][r], epoch, a.timeout_clk, &a.ctl->abort);
            asm volatile("fence.proxy.async.global;" ::: "memory");   // peer stores (generic proxy) before TMA reads
            ready |= 1u << r;
          }
        }
        for (int kb = 0; kb < KB; kb += kps_s) {
          mbar_wait(&bars.empty[stage], phase ^ 1);
          uint8_t* sa = ring + stage * kStageBytes;
          if (res_a) {             // the whole U_J tile in one stage, one box per producer
            if (pidx == 0) mbar_arrive_expect_tx(&bars.full[stage], (uint32_t)KB * kTile * 128);
            for (int k2 = pidx; k2 < KB; k2 += kProducers)
              tma_load_2d(sa + k2 * kTile * 128, map_u, k2 * kKBlock, ct * kTile, &bars.full[stage]);
          } else {
            if (pidx == 0) {
              mbar_arrive_expect_tx(&bars.full[stage], 2 * kTile * 128);
              tma_load_2d(sa, map_u, kb * kKBlock, g_row_tile0, &bars.full[stage]);
            } else if (pidx == 1) {
              tma_load_2d(sa + kTile * 128, map_u, kb * kKBlock, ct * kTile, &bars.full[stage]);
            }
          }
          advance();
        }
      };
      auto push_g = [&](int t) {   // U^T boxes [ds x 32 columns] for dU += W . U_J
        for (int kb = 0; kb < kTile / kKBlock; kb += kps_g) {
          mbar_wait(&bars.empty[stage], phase ^ 1);
          uint8_t* sa = ring + stage * kStageBytes;
          if (pidx == 0) mbar_arrive_expect_tx(&bars.full[stage], (uint32_t)kps_g * a.ds * 128);
          for (int k2 = pidx; k2 < kps_g; k2 += kProducers)
            tma_load_2d(sa + k2 * a.ds * 128, &map_ut, col_tile(t) * kTile + (kb + k2) * kKBlock, a.d0, &bars.full[stage]);
          advance();
        }
      };
      for (int t = 0; t < T; ++t) {
        push_s(t);
        if (kBwd && !kWOut && t >= 1) push_g(t - 1);
      }
      if (kBwd && !kWOut) push_g(T - 1);
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    if (lane == 0) {
      int stage = 0, phase = 0;
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      const uint32_t idesc_s = make_idesc(kTile, kTile);
      const uint32_t idesc_g = make_idesc(kTile, a.ds);
      auto mma_g = [&](int t) {    // dU += W(t) . U_J(t);  W lives in TMEM where S(t) was
        const int pb = t & 1;
        mbar_wait(&bars.epi_done[pb], (t >> 1) & 1);
        tc_fence_after();
        for (int kb = 0; kb < kTile / kKBlock; kb += kps_g) {
          mbar_wait(&bars.full[stage], phase);
          tc_fence_after();
          for (int k2 = 0; k2 < kps_g; ++k2) {
            const uint64_t bd = make_desc(ring + stage * kStageBytes + k2 * a.ds * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_ts(tmem + kDuCol, tmem + pb * kTile + (kb + k2) * kKBlock + k * 8, bd + (uint64_t)(k * 2), idesc_g,
                     (uint32_t)((t | (kb + k2) | k) != 0));
          }
          tc_commit(&bars.empty[stage]);
          advance();
        }
      };
      if (res_a) {
        mbar_wait(&bars.a_full, 0);
        tc_fence_after();
      }
      for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        if ((!kBwd || kWOut) && t >= 2) {     // no dU MMAs: S[buf] is free once the epilogue of tile t-2 has read it
          mbar_wait(&bars.epi_done[buf], ((t - 2) >> 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < KB; kb += kps_s) {
          mbar_wait(&bars.full[stage], phase);
          tc_fence_after();
          const uint8_t* sa = ring + stage * kStageBytes;
          for (int k2 = 0; k2 < kps_s; ++k2) {
            const uint64_t ad = res_a ? make_desc(smem_a + (kb + k2) * kTile * 128) : make_desc(sa);
            const uint64_t bd = res_a ? make_desc(sa + k2 * kTile * 128) : make_desc(sa + kTile * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_ss(tmem + buf * kTile, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_s,
                     (uint32_t)(((kb + k2) | k) != 0));
          }
          tc_commit(&bars.empty[stage]);
          advance();
        }
        tc_commit(&bars.tmem_full[buf]);
        if (kBwd && !kWOut && t >= 1) mma_g(t - 1);
      }
      if (kBwd && !kWOut) {
        mma_g(T - 1);
        tc_commit(&bars.du_full);
      }
    }
  } else if (warp >= 2 && warp <= 9) {
    // ===================================== epilogue ==========================================
    // Two groups of four warps: group g owns the tiles t with t & 1 == g, i.e. always TMEM buffer g, so the
    // exp-bound epilogue of tile t overlaps the epilogue of tile t+1 (and the MMAs of both).
    const int grp = (warp - 2) >> 2;
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int r_in = quad * 32 + lane;               // row inside the tile == TMEM lane
    const int etid = (tid - 64) & 127;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    const float* const lse_all = a.lse[par];
    if (kBwd && a.world > 1) {
      // the log-sum-exp scalars of rank r are complete once flag[1][r] carries this epoch (stored over NVLink by the
      // last CTAs of r's forward); S tiles are already being computed while this group waits
      if (etid < a.world && etid != a.rank) wait_peer_flag(&a.ctl->flag[1][etid], epoch, a.timeout_clk, &a.ctl->abort);
      bar_sync(2 + grp, kEpiThreads);
    }
    const float ci = kBwd ? __expf(a.inv_T - lse_all[g_row_tile0 + r_in]) : 0.f;
    const float k1 = a.k1;
    const int il = row_tile * kTile + r_in;                       // local row index
    const int pos_col = a.row0 + (il + (a.rows_valid >> 1)) % a.rows_valid;   // global column of this row's positive
    float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;         // independent partial sums: no serial FADD chain
    for (int t = grp; t < T; t += 2) {
      const int buf = grp;
      const int col0 = col_tile(t) * kTile;
      if (kBwd) {
        bar_sync(2 + grp, kEpiThreads);                         // previous tile's readers of cj[buf] are done
        bars.cj[buf][etid] = __expf(a.inv_T - lse_all[col0 + etid]);
        bar_sync(2 + grp, kEpiThreads);
      }
      mbar_wait(&bars.tmem_full[buf], (t >> 1) & 1);
      tc_fence_after();
      // special tiles: the diagonal (self-similarity is excluded) and, for the backward, the tile(s)
      // holding this row's positive -- its one-hot is folded into W *before* the TF32 rounding, so
      // the cancellation P_ip + P_pi - 2 keeps full relative precision.
      const bool diag = (col0 == g_row_tile0);
      const int jpos = kBwd ? (pos_col - col0) : -1;
      // columns of a rank's block beyond its valid rows are padding: they must not enter any sum
      const int jpad = a.rows_valid - (col0 % a.rows_per_rank);
      const bool special = diag || (jpos >= 0 && jpos < kTile) || jpad < kTile;
#pragma unroll 1
      for (int c = 0; c < kTile / 32; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem + tlane + (uint32_t)(buf * kTile + c * 32);
        tmem_ld32(taddr, v);
        tmem_wait_ld();
        if (!special) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = ex2f(fmaf(__uint_as_float(v[j]), k1, -k1));
            if (kBwd) v[j] = to_tf32(e * (ci + bars.cj[buf][c * 32 + j]));
            else if ((j & 3) == 0) rs0 += e;
            else if ((j & 3) == 1) rs1 += e;
            else if ((j & 3) == 2) rs2 += e;
            else rs3 += e;
          }
        } else {
          const int jd = diag ? (r_in - c * 32) : -1;   // position of the diagonal inside this chunk
          const int jp = jpos - c * 32;                 // position of the positive inside this chunk
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e = ex2f(fmaf(__uint_as_float(v[j]), k1, -k1));
            if (j == jd || c * 32 + j >= jpad) e = 0.f;
            if (kBwd) {
              float w = e * (ci + bars.cj[buf][c * 32 + j]);
              if (j == jp) w -= 2.f;
              v[j] = to_tf32(w);
            } else {
              rs0 += e;
            }
          }
        }
        if (kBwd && !kWOut) tmem_st32(taddr, v);
        if (kWOut) {               // one 128-byte line of this thread's row
          uint4* dst = reinterpret_cast<uint4*>(a.w_out + (size_t)(row_tile * kTile + r_in) * a.cols + col0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (kBwd && !kWOut) tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars.epi_done[buf]);
    }
    if (!kBwd) {
      // each group writes its own partial row sums: partial[(2*split + grp)][row]
      a.partial[(size_t)(2 * split + grp) * a.rows + row_tile * kTile + r_in] = (rs0 + rs1) + (rs2 + rs3);
    } else if (!kWOut) {
      mbar_wait(&bars.du_full, 0);
      tc_fence_after();
      float* dst = a.partial + ((size_t)split * a.rows + row_tile * kTile + r_in) * a.D + a.d0;
      for (int c = grp; c < a.ds / 32; c += 2) {                // the two groups read alternate 32-column chunks
        uint32_t v[32];
        tmem_ld32(tmem + tlane + (uint32_t)(kDuCol + c * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<uint4*>(dst + c * 32)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }

}

// ---- dU = W . U_all for wide embeddings (D > 256) -----------------------------------------------------------------
// Plain TF32 GEMM on tcgen05: one CTA = 128 rows x 256 columns of dU, accumulator in TMEM, K = a range of the gathered
// rows walked in 32-wide k-blocks through a 4-stage ring (A = W box 128 x 32, B = U^T box 256 x 32, both K-major,
// 128-byte swizzle).  warp 0 / warp 6: TMA producers (A / B; one thread sustains one box per ~455 clk, a k-block of MMAs
// lasts ~740), warp 1: MMA issuer, warps 2-5: epilogue (TMEM -> partial[split][rows][D]).
struct GemmArgs {
  int rows, D;               // padded rows of this rank, embedding width
  int kblocks, kb_per_split; // k-blocks (32 gathered rows each) in total / per grid.z slice
  float* partial;            // [ksplit][rows][D]
};
constexpr int kGemmThreads = 224;
constexpr int kGemmN = 256;
constexpr int kGemmStages = 4;
constexpr int kGemmStageBytes = (kTile + kGemmN) * 128;   // 48 KB

__global__ void __launch_bounds__(kGemmThreads, 1)
wu_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_ut,
               const __grid_constant__ GemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars& bars = *reinterpret_cast<Bars*>(smem + kRingBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x, row_tile = blockIdx.y, split = blockIdx.z;   // the n-tiles of a row tile are co-scheduled:
                                                                              // they read the same W boxes (L2 hits)
  const int kb0 = split * a.kb_per_split;
  const int nkb = min(a.kb_per_split, a.kblocks - kb0);
  if (warp == 0 && lane == 0) prefetch_tmap(&map_w);
  if (warp == 6 && lane == 0) prefetch_tmap(&map_ut);
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kGemmStages; ++i) {
        mbar_init(&bars.full[i], 1);
        mbar_init(&bars.empty[i], 1);
      }
      mbar_init(&bars.du_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_ptr)),
                 "r"(kGemmN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_ptr;

  if (warp == 0 || warp == 6) {
    if (lane == 0) {
      const bool is_a = warp == 0;
      int stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&bars.empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kGemmStageBytes;
        const int k0 = (kb0 + kb) * kKBlock;
        if (is_a) {
          mbar_arrive_expect_tx(&bars.full[stage], (uint32_t)kGemmStageBytes);
          tma_load_2d(sa, &map_w, k0, row_tile * kTile, &bars.full[stage]);
        } else {
          tma_load_2d(sa + kTile * 128, &map_ut, k0, n_tile * kGemmN, &bars.full[stage]);
        }
        if (++stage == kGemmStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kTile, kGemmN);
      int stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&bars.full[stage], phase);
        tc_fence_after();
        const uint8_t* sa = smem + stage * kGemmStageBytes;
        const uint64_t ad = make_desc(sa), bd = make_desc(sa + kTile * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_ss(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
        tc_commit(&bars.empty[stage]);
        if (++stage == kGemmStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit(&bars.du_full);
    }
  } else {
    const int quad = warp & 3;
    const int r_in = quad * 32 + lane;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    mbar_wait(&bars.du_full, 0);
    tc_fence_after();
    float* dst = a.partial + ((size_t)split * a.rows + row_tile * kTile + r_in) * a.D + n_tile * kGemmN;
#pragma unroll 1
    for (int c = 0; c < kGemmN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem + tlane + (uint32_t)(c * 32), v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<uint4*>(dst + c * 32)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kGemmN) : "memory");
  }
}

// ---- small kernels ------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_as_f32(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_as_f32<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_f32<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }

// ---- peer-memory exchange (NVLink stores, no NCCL call on the data path) --------------------------------------------
// A producer writes its rows straight into the same buffer of EVERY rank of the node (peer pointers of one symmetric
// allocation; dst[rank] is the local copy) and its last CTA raises flag[slot][rank] = epoch in every rank's control
// block with a system-scope release; the CONSUMING tile kernels wait for the flags themselves (wait_peer_flag: the
// forward's TMA producers per column tile, the backward's epilogue warps once).  Epoch and buffer parity come from the
// device-side counter PeerCtl::epoch.  world == 1 degenerates to a plain local store with no signalling.
struct PrepPeers {
  float* dst[2][kMaxPeers];     // per parity, per rank: gathered matrix
  PeerCtl* ctl[kMaxPeers];      // per rank: control block; ctl[rank] is local; null when world == 1
  int world, rank;
};

// one warp per row: rinv = 1/max(|z|,1e-12), u = tf32(z * rinv), written to every rank's gathered matrix at row row0 + i
template <typename T>
__global__ void prep_kernel(const T* __restrict__ z, int rows, int rows_pad, int D, int row0, float* __restrict__ rinv,
                            const PrepPeers pe) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  PeerCtl* const ctl = pe.world > 1 ? pe.ctl[pe.rank] : nullptr;
  const uint32_t epoch = ctl ? ctl->epoch + 1u : 0u;
  const int par = (int)(epoch & 1u);
  if (row >= rows && row < rows_pad) {            // padding rows of the 128-row tiles: zero vectors (masked as columns)
    for (int d = lane; d < D; d += 32)
      for (int q = 0; q < pe.world; ++q) pe.dst[par][q][(size_t)(row0 + row) * D + d] = 0.f;
    if (lane == 0) rinv[row] = 0.f;
  }
  if (row < rows) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = load_as_f32(z, (size_t)row * D + d);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float r = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    for (int d = lane; d < D; d += 32) {
      const float v = __uint_as_float(to_tf32(load_as_f32(z, (size_t)row * D + d) * r));
      for (int q = 0; q < pe.world; ++q) pe.dst[par][q][(size_t)(row0 + row) * D + d] = v;
    }
    if (lane == 0) rinv[row] = r;
  }
  if (!ctl) return;
  // the last CTA to arrive signals all ranks: flag[0][rank] = epoch
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&ctl->prep_counter, 1u) == gridDim.x - 1) {
      __threadfence_system();
      ctl->prep_counter = 0u;
      for (int q = 0; q < pe.world; ++q) st_release_sys(&pe.ctl[q]->flag[0][pe.rank], epoch);
    }
  }
}

// U [cols, D] -> U^T [D, cols] (the gathered matrix of the current parity when a control block is given)
__global__ void transpose_kernel(const float* __restrict__ u0, const float* __restrict__ u1, const PeerCtl* ctl,
                                 float* __restrict__ ut, int cols, int D) {
  __shared__ float tile[32][33];
  const float* u = (ctl && (ctl->epoch & 1u)) ? u1 : u0;
  const int c0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = u[(size_t)(c0 + i) * D + d0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) ut[(size_t)(d0 + i) * cols + c0 + threadIdx.x] = tile[threadIdx.x][i];
}

// Finishes the forward rows: lse_i = 1/T + log(sum of the split partials); pos_i = <u_i,u_p(i)>/T; row loss
// lse_i - pos_i; the lse rows go to every rank's gathered vector.  One warp per row, all loads of a row issued before
// the first use.  The LAST block forms the mean loss (fixed order: deterministic), raises flag[1][rank] = epoch on
// every rank, publishes the epoch (PeerCtl::epoch: the backward and the next forward read it) and re-arms the counter.
__global__ void __launch_bounds__(256) fwd_rows_kernel(const RowsArgs a) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const uint32_t epoch = a.ctl ? a.ctl->epoch + 1u : 0u;
  const int par = (int)(epoch & 1u);
  if (i < a.rows) {
    if (i >= a.rows_valid) {        // padding row: a finite lse for the peers' c_j, no loss term
      if (lane == 0) {
        for (int q = 0; q < a.world; ++q) a.lse_dst[par][q][a.lse_off + i] = 0.f;
        a.row_loss[i] = 0.f;
      }
    } else {
      const float* u_all = a.u_all[par];
      const float* ui = u_all + (size_t)(a.row0 + i) * a.D;
      const float* up = u_all + (size_t)(a.row0 + (i + (a.rows_valid >> 1)) % a.rows_valid) * a.D;
      float sacc = 0.f;
      for (int k = lane; k < a.nparts; k += 32) sacc += a.partial[(size_t)k * a.rows + i];
      float dot = 0.f;
      for (int d0 = lane; d0 < a.D; d0 += 32 * 4) {
        float x[4], y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = d0 + 32 * e;
          x[e] = d < a.D ? ui[d] : 0.f;
          y[e] = d < a.D ? up[d] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) dot = fmaf(x[e], y[e], dot);
      }
      sacc = warp_sum(sacc);
      dot = warp_sum(dot);
      if (lane == 0) {
        const float lse = a.inv_T + logf(sacc);
        for (int q = 0; q < a.world; ++q) a.lse_dst[par][q][a.lse_off + i] = lse;     // every rank's gathered lse
        a.row_loss[i] = lse - dot * a.inv_T;
      }
    }
  }
  __shared__ bool last;
  __shared__ float red[8];
  if (a.world > 1) __threadfence_system();
  else __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  float acc = 0.f;
  for (int r0 = threadIdx.x; r0 < a.rows; r0 += 8 * 256) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = r0 + e * 256 < a.rows ? __ldcg(a.row_loss + r0 + e * 256) : 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += v[e];
  }
  acc = warp_sum(acc);
  if (lane == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    a.loss[0] = tot / (float)a.rows_valid;
    *a.counter = 0u;                                   // ready for the next call on this scratch buffer
    if (a.ctl) {
      __threadfence_system();
      for (int q = 0; q < a.world; ++q) st_release_sys(&a.ctl_peers[q]->flag[1][a.rank], epoch);
      a.ctl->epoch = epoch;                            // this forward is complete
    }
  }
}

// dz_i = (dU_i - u_i <u_i,dU_i>) * rinv_i with dU_i = scale * sum_splits partial (the -2 u_p(i) one-hot term is already
// inside W); one warp per row, eight elements per lane in flight per step.
template <typename T>
__global__ void bwd_finalize_kernel(const float* __restrict__ partial, int nsplit, const T* __restrict__ z_rows,
                                    const float* __restrict__ rinv, int D, int rows, int rows_valid, float scale,
                                    const float* __restrict__ grad_out, T* __restrict__ dz) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= rows_valid) return;
  const float g = scale * (grad_out ? grad_out[0] : 1.f);
  const float r = rinv[i];
  constexpr int kE = 8;
  float dot = 0.f;
  if (D <= 32 * kE) {                 // the common case: the whole row stays in registers, one pass
    float du[kE], zz[kE];
#pragma unroll
    for (int e = 0; e < kE; ++e) {
      const int d = lane + 32 * e;
      du[e] = 0.f;
      zz[e] = d < D ? load_as_f32(z_rows, (size_t)i * D + d) : 0.f;
      if (d < D)
        for (int sp = 0; sp < nsplit; ++sp) du[e] += partial[((size_t)sp * rows + i) * D + d];
    }
#pragma unroll
    for (int e = 0; e < kE; ++e) {
      du[e] *= g;
      dot = fmaf(zz[e] * r, du[e], dot);
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int e = 0; e < kE; ++e) {
      const int d = lane + 32 * e;
      if (d < D) {
        const float v = (du[e] - zz[e] * r * dot) * r;
        if constexpr (sizeof(T) == 4) dz[(size_t)i * D + d] = v;
        else dz[(size_t)i * D + d] = __float2bfloat16_rn(v);
      }
    }
    return;
  }
  auto du_at = [&](int d) {
    float sacc = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) sacc += partial[((size_t)sp * rows + i) * D + d];
    return g * sacc;
  };
  for (int d = lane; d < D; d += 32) dot = fmaf(load_as_f32(z_rows, (size_t)i * D + d) * r, du_at(d), dot);
  dot = warp_sum(dot);
  for (int d = lane; d < D; d += 32) {
    const float v = (du_at(d) - load_as_f32(z_rows, (size_t)i * D + d) * r * dot) * r;
    if constexpr (sizeof(T) == 4) dz[(size_t)i * D + d] = v;
    else dz[(size_t)i * D + d] = __float2bfloat16_rn(v);
  }
}

// BYOL: loss = 2 - 2 mean cos(p_i,t_i); dp_i = -(2/n) (t^_i - cos_i p^_i)/|p_i|.  One warp per row.
__global__ void byol_rows_kernel(const float* __restrict__ p, const float* __restrict__ t, int rows, int D,
                                 float* __restrict__ cos_rows, float* __restrict__ dp) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= rows) return;
  float pp = 0.f, tt = 0.f, pt = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float a = p[(size_t)i * D + d], b = t[(size_t)i * D + d];
    pp = fmaf(a, a, pp);
    tt = fmaf(b, b, tt);
    pt = fmaf(a, b, pt);
  }
  pp = warp_sum(pp);
  tt = warp_sum(tt);
  pt = warp_sum(pt);
  const float np = fmaxf(sqrtf(pp), 1e-12f), nt = fmaxf(sqrtf(tt), 1e-12f);
  const float c = pt / (np * nt);
  if (lane == 0) cos_rows[i] = c;
  if (dp) {
    const float k = -2.f / (float)rows;
    for (int d = lane; d < D; d += 32) {
      const float ph = p[(size_t)i * D + d] / np, th = t[(size_t)i * D + d] / nt;
      dp[(size_t)i * D + d] = k * (th - c * ph) / np;
    }
  }
}
__global__ void byol_reduce_kernel(const float* __restrict__ cos_rows, int rows, float* __restrict__ loss) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) acc += cos_rows[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    loss[0] = 2.f - 2.f * tot / (float)rows;
  }
}

// ---- host helpers -----------------------------------------------------------------------------
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

// fp32 row-major [outer, inner] matrix, box [box_outer x 32 floats], 128-byte swizzle
static int make_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  MIS_REQUIRE(fn, MIS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {inner * sizeof(float)};
  const cuuint32_t box[2] = {kKBlock, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MIS_REQUIRE(r == CUDA_SUCCESS, MIS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MIS_OK;
}

struct Plan {
  int row_tiles, col_tiles, tiles_per_split, nsplit;
};
static Plan make_plan(int rows, int cols, int D) {
  Plan p;
  p.row_tiles = rows / kTile;
  p.col_tiles = cols / kTile;
  int target = 148 / p.row_tiles;
  if (target < 1) target = 1;
  // wide embeddings stream the row tile as well (128 x D x 4 bytes per column tile): keep the row tiles that are live at
  // once (148 / splits of them) within ~40 MB so that they and the column tiles they share stay in the L2
  if (D > 256) {
    const int need = (int)((148.0 * kTile * D * 4 + 40e6 - 1) / 40e6);
    if (target < need) target = need;
  }
  if (target > p.col_tiles) target = p.col_tiles;
  p.tiles_per_split = (p.col_tiles + target - 1) / target;
  p.nsplit = (p.col_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}
static inline size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

// wide embeddings (D > 256): K-splits of the dU = W . U_all GEMM, enough for one CTA per SM
static int gemm_ksplit(int rows, int cols, int D) {
  const int ctas = (rows / kTile) * (D / kGemmN);
  int ks = ctas >= 148 ? 1 : (148 + ctas - 1) / ctas;
  if (ks > 16) ks = 16;
  const int kblocks = cols / kKBlock;
  if (ks > kblocks) ks = kblocks;
  return ks;
}

static inline int pad_rows(int rows) { return (rows + kTile - 1) / kTile * kTile; }

// `rows` = the embedding rows a rank really has (any even number); its block of the gathered matrix is padded to
// pad_rows(rows); cols and row0 are in padded coordinates
static int check_shapes(const char* who, int rows, int cols, int D, int row0, float inv_T, bool bwd) {
  MIS_REQUIRE(rows > 0 && cols > 0 && D > 0, MIS_ERR_INVALID_ARG, "%s: sizes must be positive", who);
  MIS_REQUIRE(rows % 2 == 0, MIS_ERR_INVALID_ARG, "%s: rows (%d) = [view 1; view 2] must be even", who, rows);
  const int rp = pad_rows(rows);
  MIS_REQUIRE(cols % rp == 0 && row0 % rp == 0, MIS_ERR_INVALID_ARG,
              "%s: cols (%d) and row0 (%d) must be multiples of the padded rows per rank (%d = %d rows padded to %d)", who,
              cols, row0, rp, rows, kTile);
  rows = rp;
  MIS_REQUIRE(rows / kTile < kThreads - 1, MIS_ERR_UNSUPPORTED, "%s: at most %d rows per rank", who, (kThreads - 2) * kTile);
  MIS_REQUIRE(D % kKBlock == 0 && D <= 8192, MIS_ERR_UNSUPPORTED, "%s: D=%d must be a multiple of 32 (<= 8192)", who, D);
  MIS_REQUIRE(!bwd || D <= 256 || D % 256 == 0, MIS_ERR_UNSUPPORTED,
              "%s: D=%d > 256 must be a multiple of 256 (the backward walks dU in 256-column slices)", who, D);
  MIS_REQUIRE(row0 >= 0 && row0 + rows <= cols, MIS_ERR_INVALID_ARG, "%s: rows [%d,%d) outside [0,%d)", who, row0,
              row0 + rows, cols);
  MIS_REQUIRE(inv_T > 0.f && inv_T <= 40.f, MIS_ERR_UNSUPPORTED,
              "%s: temperature %g outside (0.025, inf): the fixed-max log-sum-exp needs exp(-2/T) to stay normal", who,
              1.0 / inv_T);
  return MIS_OK;
}

constexpr size_t kSmemBytes = (size_t)kRingBytes + sizeof(Bars) + 1024;
constexpr size_t kCounterBytes = 2048;   // [row_tiles + 1] uint32 arrival counters at the start of the scratch

// everything a multi-rank call adds to the single-rank one (world == 1: all null)
struct Exchange {
  int world = 1, rank = 0;
  PeerCtl* ctl_peers[kMaxPeers] = {};
  float* u_peers[2][kMaxPeers] = {};
  float* lse_peers[2][kMaxPeers] = {};
  long long timeout_clk = 0;
};

}  // namespace ntx
}  // namespace mis

using namespace mis;
using namespace mis::ntx;

extern "C" int mis_ntxent_padded_rows(int rows) { return rows <= 0 ? 0 : pad_rows(rows); }

// out[pad128(M), pad256(N)] = A[M, K] . B[N, K]^T (fp32 containers, TF32 MMA, fp32 accumulate) with wu_gemm_kernel;
// rows / columns beyond M / N are read as zeros by TMA and come out as zeros.  Used by the kNN evaluator (knn.cu).
namespace mis {
int launch_gemm_tf32_nt(const float* A, int M, const float* B, int N, int K, float* out, cudaStream_t st) {
  MIS_REQUIRE(A && B && out && M > 0 && N > 0 && K >= kKBlock && K % kKBlock == 0, MIS_ERR_INVALID_ARG,
              "gemm_tf32_nt: bad arguments (M %d N %d K %d)", M, N, K);
  MIS_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              MIS_ERR_INVALID_ARG, "gemm_tf32_nt: operands must be 16-byte aligned");
  const int Mp = pad_rows(M), Np = (N + kGemmN - 1) / kGemmN * kGemmN;
  MIS_REQUIRE(Mp / kTile <= 65535, MIS_ERR_UNSUPPORTED, "gemm_tf32_nt: at most %d rows per call", 65535 * kTile);
  CUtensorMap map_a, map_b;
  if (int rc = make_map(&map_a, A, (uint64_t)K, (uint64_t)M, kTile)) return rc;
  if (int rc = make_map(&map_b, B, (uint64_t)K, (uint64_t)N, kGemmN)) return rc;
  GemmArgs g = {};
  g.rows = Mp; g.D = Np;
  g.kblocks = K / kKBlock;
  g.kb_per_split = g.kblocks;
  g.partial = out;
  MIS_CUDA_TRY(cudaFuncSetAttribute(wu_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  wu_gemm_kernel<<<dim3(Np / kGemmN, Mp / kTile, 1), kGemmThreads, kSmemBytes, st>>>(map_a, map_b, g);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}
}  // namespace mis

extern "C" int64_t mis_ntxent_scratch_bytes(int rows, int cols, int D) {
  if (rows <= 0 || cols <= 0 || D <= 0) return -1;
  rows = pad_rows(rows);
  cols = pad_rows(cols);
  const Plan p = make_plan(rows, cols, D);
  size_t b = 0;
  b += al256(kCounterBytes + (size_t)cols * 4);       // arrival counters + per-row loss terms
  b += al256((size_t)D * cols * 4);                   // U^T
  size_t np = (size_t)p.nsplit;
  if (D > 256) {                                      // wide embeddings: W is materialised, dU partials per GEMM K-split
    np = (size_t)gemm_ksplit(rows, cols, D);
    if (np * D < (size_t)2 * p.nsplit) np = ((size_t)2 * p.nsplit + D - 1) / D;   // forward row-sum partials [2 nsplit][rows]
  }
  b += al256(np * rows * D * 4);                      // dU partials (also covers the forward's row-sum partials)
  if (D > 256) b += al256((size_t)rows * cols * 4);   // W
  return (int64_t)b;
}

static int launch_prep(const void* z, int z_dtype, int rows, int D, int row0, float* rinv, const PrepPeers& pe, cudaStream_t st) {
  const int wpb = 8;
  const int rp = pad_rows(rows);
  const dim3 grid((rp + wpb - 1) / wpb), block(wpb * 32);
  if (z_dtype == MIS_DTYPE_F32)
    prep_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(z), rows, rp, D, row0, rinv, pe);
  else
    prep_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(z), rows, rp, D, row0, rinv, pe);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}

// peer tables of one call -> Exchange (validated before any CUDA call)
static int make_exchange(Exchange* ex, const char* who, int world, int rank, void* const* u0, void* const* u1,
                         void* const* l0, void* const* l1, void* const* ctl, double timeout_s) {
  MIS_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, MIS_ERR_INVALID_ARG,
              "%s: world %d / rank %d (2 to %d ranks of one node)", who, world, rank, kMaxPeers);
  MIS_REQUIRE(u0 && u1 && l0 && l1 && ctl, MIS_ERR_INVALID_ARG, "%s: null peer table", who);
  MIS_REQUIRE(timeout_s > 0.0, MIS_ERR_INVALID_ARG, "%s: peer time-out must be positive", who);
  *ex = Exchange{};
  for (int r = 0; r < world; ++r) {
    MIS_REQUIRE(u0[r] && u1[r] && l0[r] && l1[r] && ctl[r], MIS_ERR_INVALID_ARG, "%s: null peer pointer for rank %d", who, r);
    ex->u_peers[0][r] = static_cast<float*>(u0[r]);
    ex->u_peers[1][r] = static_cast<float*>(u1[r]);
    ex->lse_peers[0][r] = static_cast<float*>(l0[r]);
    ex->lse_peers[1][r] = static_cast<float*>(l1[r]);
    ex->ctl_peers[r] = static_cast<PeerCtl*>(ctl[r]);
  }
  ex->world = world;
  ex->rank = rank;
  int dev = 0, khz = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0) khz = 1965000;
  ex->timeout_clk = (long long)(timeout_s * 1e3 * (double)khz);
  return MIS_OK;
}

extern "C" int mis_ntxent_prep(const void* z, int z_dtype, int rows, int D, float* u, float* rinv, void* stream) {
  MIS_REQUIRE(z && u && rinv, MIS_ERR_INVALID_ARG, "mis_ntxent_prep: null pointer");
  MIS_REQUIRE(rows > 0 && D > 0, MIS_ERR_INVALID_ARG, "mis_ntxent_prep: sizes must be positive");
  MIS_REQUIRE(z_dtype == MIS_DTYPE_F32 || z_dtype == MIS_DTYPE_BF16, MIS_ERR_INVALID_ARG, "mis_ntxent_prep: dtype %d", z_dtype);
  PrepPeers pe = {};
  pe.dst[0][0] = u;
  pe.world = 1;
  return launch_prep(z, z_dtype, rows, D, 0, rinv, pe, reinterpret_cast<cudaStream_t>(stream));
}

// forward over the gathered matrix: ONE tile-kernel launch (+ a memset node for its arrival counters)
static int ntxent_fwd_impl(const float* u0, const float* u1, int cols, int D, int row0, int rows, float inv_T,
                           const Exchange& ex, float* lse_local, float* loss, void* scratch, int64_t scratch_bytes,
                           cudaStream_t st) {
  MIS_REQUIRE(u0 && loss && scratch, MIS_ERR_INVALID_ARG, "mis_ntxent_fwd: null pointer");
  if (int rc = check_shapes("mis_ntxent_fwd", rows, cols, D, row0, inv_T, false)) return rc;
  MIS_REQUIRE(scratch_bytes >= mis_ntxent_scratch_bytes(rows, cols, D), MIS_ERR_INVALID_ARG,
              "mis_ntxent_fwd: scratch too small (%lld < %lld)", (long long)scratch_bytes,
              (long long)mis_ntxent_scratch_bytes(rows, cols, D));
  const int rows_valid = rows;
  rows = pad_rows(rows);
  const Plan p = make_plan(rows, cols, D);
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  float* partial = reinterpret_cast<float*>(sc + al256(kCounterBytes + (size_t)cols * 4) + al256((size_t)D * cols * 4));

  CUtensorMap map0, map1;
  if (int rc = make_map(&map0, u0, (uint64_t)D, (uint64_t)cols, kTile)) return rc;
  if (int rc = make_map(&map1, u1 ? u1 : u0, (uint64_t)D, (uint64_t)cols, kTile)) return rc;
  TileArgs a = {};
  a.rows = rows; a.cols = cols; a.D = D; a.row0 = row0;
  a.rows_valid = rows_valid;
  a.col_tiles = p.col_tiles; a.tiles_per_split = p.tiles_per_split;
  a.k1 = kLog2e * inv_T;
  a.inv_T = inv_T;
  a.partial = partial;
  a.world = ex.world; a.rank = ex.rank; a.rows_per_rank = rows;      // (padded)
  a.epoch_add = 1;
  a.timeout_clk = ex.timeout_clk;
  if (ex.world > 1) {
    a.ctl = ex.ctl_peers[ex.rank];
    for (int r = 0; r < ex.world; ++r) a.ctl_peers[r] = ex.ctl_peers[r];
  }
  auto* fn = &ntxent_tile_kernel<false>;
  MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  a.split_fast = D > 256;
  fn<<<a.split_fast ? dim3(p.nsplit, p.row_tiles) : dim3(p.row_tiles, p.nsplit), kThreads, kSmemBytes, st>>>(map0, map1, map0, a);
  MIS_CUDA_TRY(cudaGetLastError());

  RowsArgs ra = {};
  ra.partial = partial;
  ra.nparts = 2 * p.nsplit;
  ra.rows = rows; ra.rows_valid = rows_valid; ra.D = D; ra.row0 = row0;
  ra.inv_T = inv_T;
  ra.u_all[0] = u0;
  ra.u_all[1] = u1 ? u1 : u0;
  ra.row_loss = reinterpret_cast<float*>(sc + kCounterBytes);
  ra.loss = loss;
  ra.counter = reinterpret_cast<unsigned int*>(sc);
  ra.world = ex.world; ra.rank = ex.rank;
  if (ex.world > 1) {
    ra.ctl = ex.ctl_peers[ex.rank];
    for (int r = 0; r < ex.world; ++r) {
      ra.ctl_peers[r] = ex.ctl_peers[r];
      ra.lse_dst[0][r] = ex.lse_peers[0][r];
      ra.lse_dst[1][r] = ex.lse_peers[1][r];
    }
    ra.lse_off = row0;
  } else {
    ra.lse_dst[0][0] = ra.lse_dst[1][0] = lse_local;
    ra.lse_off = 0;
  }
  MIS_CUDA_TRY(cudaMemsetAsync(sc, 0, sizeof(unsigned int), st));                 // scratch arrives uninitialised
  fwd_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ra);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}

extern "C" int mis_ntxent_fwd(const float* u_all, int cols, int D, int row0, int rows, float inv_T, float* lse_rows,
                              float* loss, void* scratch, int64_t scratch_bytes, void* stream) {
  MIS_REQUIRE(lse_rows, MIS_ERR_INVALID_ARG, "mis_ntxent_fwd: null pointer");
  return ntxent_fwd_impl(u_all, nullptr, cols, D, row0, rows, inv_T, Exchange{}, lse_rows, loss, scratch, scratch_bytes,
                         reinterpret_cast<cudaStream_t>(stream));
}

// backward: transpose + ONE tile-kernel launch per 256-column slice of dU (+ the finalize kernel when D > 256)
static int ntxent_bwd_impl(const float* u0, const float* u1, const float* lse0, const float* lse1, const void* z_rows,
                           int z_dtype, const float* rinv_rows, int cols, int D, int row0, int rows, float inv_T,
                           float grad_scale, const float* grad_out, void* dz, const Exchange& ex, void* scratch,
                           int64_t scratch_bytes, cudaStream_t st) {
  MIS_REQUIRE(u0 && lse0 && z_rows && rinv_rows && dz && scratch, MIS_ERR_INVALID_ARG, "mis_ntxent_bwd: null pointer");
  MIS_REQUIRE(z_dtype == MIS_DTYPE_F32 || z_dtype == MIS_DTYPE_BF16, MIS_ERR_INVALID_ARG, "mis_ntxent_bwd: dtype %d", z_dtype);
  if (int rc = check_shapes("mis_ntxent_bwd", rows, cols, D, row0, inv_T, true)) return rc;
  MIS_REQUIRE(scratch_bytes >= mis_ntxent_scratch_bytes(rows, cols, D), MIS_ERR_INVALID_ARG,
              "mis_ntxent_bwd: scratch too small (%lld < %lld)", (long long)scratch_bytes,
              (long long)mis_ntxent_scratch_bytes(rows, cols, D));
  const int rows_valid = rows;
  rows = pad_rows(rows);
  const Plan p = make_plan(rows, cols, D);
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  float* ut = reinterpret_cast<float*>(sc + al256(kCounterBytes + (size_t)cols * 4));
  float* partial = reinterpret_cast<float*>(sc + al256(kCounterBytes + (size_t)cols * 4) + al256((size_t)D * cols * 4));
  PeerCtl* ctl = ex.world > 1 ? ex.ctl_peers[ex.rank] : nullptr;

  transpose_kernel<<<dim3(cols / 32, D / 32), dim3(32, 8), 0, st>>>(u0, u1 ? u1 : u0, ctl, ut, cols, D);
  MIS_CUDA_TRY(cudaGetLastError());

  CUtensorMap map0, map1, map_ut;
  if (int rc = make_map(&map0, u0, (uint64_t)D, (uint64_t)cols, kTile)) return rc;
  if (int rc = make_map(&map1, u1 ? u1 : u0, (uint64_t)D, (uint64_t)cols, kTile)) return rc;
  const int ds = D <= 256 ? D : 256;     // dU lives in TMEM columns 256..511: at most 256 columns in the fused kernel;
                                         // wider embeddings take the W-materialised path below (256-column GEMM tiles)
  if (int rc = make_map(&map_ut, ut, (uint64_t)cols, (uint64_t)D, (uint32_t)ds)) return rc;
  const float scale = grad_scale * inv_T / (float)rows_valid;
  TileArgs a = {};
  a.rows = rows; a.cols = cols; a.D = D; a.row0 = row0;
  a.rows_valid = rows_valid;
  a.col_tiles = p.col_tiles; a.tiles_per_split = p.tiles_per_split;
  a.k1 = kLog2e * inv_T;
  a.lse[0] = lse0;
  a.lse[1] = lse1 ? lse1 : lse0;
  a.inv_T = inv_T;
  a.partial = partial;
  a.ds = ds;
  a.world = ex.world; a.rank = ex.rank; a.rows_per_rank = rows;
  a.epoch_add = 0;
  a.timeout_clk = ex.timeout_clk;
  a.ctl = ctl;
  int nsum = p.nsplit;                   // partial buffers the Jacobian kernel adds up
  if (D <= 256) {
    auto* fn = &ntxent_tile_kernel<true, false>;
    MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    a.d0 = 0;
    fn<<<dim3(p.row_tiles, p.nsplit), kThreads, kSmemBytes, st>>>(map0, map1, map_ut, a);
    MIS_CUDA_TRY(cudaGetLastError());
  } else {
    // wide embeddings: W tiles to HBM (S computed once), then dU = W . U_all as one GEMM over all of D
    const int ks = gemm_ksplit(rows, cols, D);
    size_t np = (size_t)ks;
    if (np * D < (size_t)2 * p.nsplit) np = ((size_t)2 * p.nsplit + D - 1) / D;
    float* wbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(partial) + al256(np * rows * D * 4));
    a.w_out = wbuf;
    a.split_fast = 1;
    auto* fn = &ntxent_tile_kernel<true, true>;
    MIS_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    fn<<<dim3(p.nsplit, p.row_tiles), kThreads, kSmemBytes, st>>>(map0, map1, map_ut, a);
    MIS_CUDA_TRY(cudaGetLastError());
    CUtensorMap map_w;
    if (int rc = make_map(&map_w, wbuf, (uint64_t)cols, (uint64_t)rows, kTile)) return rc;
    GemmArgs g = {};
    g.rows = rows; g.D = D;
    g.kblocks = cols / kKBlock;
    g.kb_per_split = (g.kblocks + ks - 1) / ks;
    g.partial = partial;
    nsum = (g.kblocks + g.kb_per_split - 1) / g.kb_per_split;
    MIS_CUDA_TRY(cudaFuncSetAttribute(wu_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    wu_gemm_kernel<<<dim3(D / kGemmN, p.row_tiles, nsum), kGemmThreads, kSmemBytes, st>>>(map_w, map_ut, g);
    MIS_CUDA_TRY(cudaGetLastError());
  }
  {
    const int wpb = 8;
    const dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
    if (z_dtype == MIS_DTYPE_F32)
      bwd_finalize_kernel<float><<<grid, block, 0, st>>>(partial, nsum, static_cast<const float*>(z_rows), rinv_rows, D,
                                                          rows, rows_valid, scale, grad_out, static_cast<float*>(dz));
    else
      bwd_finalize_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(partial, nsum, static_cast<const __nv_bfloat16*>(z_rows),
                                                                  rinv_rows, D, rows, rows_valid, scale, grad_out,
                                                                  static_cast<__nv_bfloat16*>(dz));
    MIS_CUDA_TRY(cudaGetLastError());
  }
  return MIS_OK;
}

extern "C" int mis_ntxent_bwd(const float* u_all, const float* lse_all, const void* z_rows, int z_dtype,
                              const float* rinv_rows, int cols, int D, int row0, int rows, float inv_T, float grad_scale,
                              const float* grad_out, void* dz, void* scratch, int64_t scratch_bytes, void* stream) {
  return ntxent_bwd_impl(u_all, nullptr, lse_all, nullptr, z_rows, z_dtype, rinv_rows, cols, D, row0, rows, inv_T,
                         grad_scale, grad_out, dz, Exchange{}, scratch, scratch_bytes, reinterpret_cast<cudaStream_t>(stream));
}

// ---- multi-rank: one ABI call per autograd phase; epoch and buffer parity live on the device (PeerCtl::epoch), so both
// calls can be captured into CUDA graphs and replayed ------------------------------------------------------------------
extern "C" int mis_ntxent_fwd_peer(const void* z, int z_dtype, int rows, int D, float inv_T, int world, int rank,
                                   void* const* u_peers0, void* const* u_peers1, void* const* lse_peers0,
                                   void* const* lse_peers1, void* const* ctl_peers, double timeout_s, float* rinv,
                                   float* loss, void* scratch, int64_t scratch_bytes, void* stream) {
  MIS_REQUIRE(z && rinv && loss && scratch, MIS_ERR_INVALID_ARG, "mis_ntxent_fwd_peer: null pointer");
  MIS_REQUIRE(z_dtype == MIS_DTYPE_F32 || z_dtype == MIS_DTYPE_BF16, MIS_ERR_INVALID_ARG, "mis_ntxent_fwd_peer: dtype %d", z_dtype);
  Exchange ex;
  if (int rc = make_exchange(&ex, "mis_ntxent_fwd_peer", world, rank, u_peers0, u_peers1, lse_peers0, lse_peers1, ctl_peers,
                             timeout_s)) return rc;
  const int rp = pad_rows(rows > 0 ? rows : 1);
  if (int rc = check_shapes("mis_ntxent_fwd_peer", rows, world * rp, D, rank * rp, inv_T, false)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PrepPeers pe = {};
  pe.world = world;
  pe.rank = rank;
  for (int r = 0; r < world; ++r) {
    pe.dst[0][r] = ex.u_peers[0][r];
    pe.dst[1][r] = ex.u_peers[1][r];
    pe.ctl[r] = ex.ctl_peers[r];
  }
  if (int rc = launch_prep(z, z_dtype, rows, D, rank * rp, rinv, pe, st)) return rc;
  return ntxent_fwd_impl(ex.u_peers[0][rank], ex.u_peers[1][rank], world * rp, D, rank * rp, rows, inv_T, ex, nullptr,
                         loss, scratch, scratch_bytes, st);
}

extern "C" int mis_ntxent_bwd_peer(const void* z_rows, int z_dtype, const float* rinv_rows, int rows, int D, float inv_T,
                                   float grad_scale, const float* grad_out, void* dz, int world, int rank,
                                   void* const* u_peers0, void* const* u_peers1, void* const* lse_peers0,
                                   void* const* lse_peers1, void* const* ctl_peers, double timeout_s, void* scratch,
                                   int64_t scratch_bytes, void* stream) {
  Exchange ex;
  if (int rc = make_exchange(&ex, "mis_ntxent_bwd_peer", world, rank, u_peers0, u_peers1, lse_peers0, lse_peers1, ctl_peers,
                             timeout_s)) return rc;
  const int rp = pad_rows(rows > 0 ? rows : 1);
  return ntxent_bwd_impl(ex.u_peers[0][rank], ex.u_peers[1][rank], ex.lse_peers[0][rank], ex.lse_peers[1][rank], z_rows,
                         z_dtype, rinv_rows, world * rp, D, rank * rp, rows, inv_T, grad_scale, grad_out, dz, ex, scratch,
                         scratch_bytes, reinterpret_cast<cudaStream_t>(stream));
}

// ---- single-rank convenience: prep -> forward -> backward in one call (4 kernels + 2 memsets, one host round trip) ----
static inline size_t ws_u_bytes(int rows, int D) { return al256((size_t)pad_rows(rows) * D * 4); }
static inline size_t ws_row_bytes(int rows) { return al256((size_t)pad_rows(rows) * 4); }

extern "C" int64_t mis_ntxent_fwd_bwd_workspace_bytes(int rows, int D) {
  if (rows <= 0 || D <= 0) return -1;
  const int64_t sc = mis_ntxent_scratch_bytes(rows, rows, D);
  if (sc < 0) return -1;
  return (int64_t)(ws_u_bytes(rows, D) + 2 * ws_row_bytes(rows)) + sc;
}

extern "C" int mis_ntxent_fwd_bwd(const void* z, int z_dtype, int rows, int D, float inv_T, float* loss, void* dz,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  MIS_REQUIRE(z && loss && dz && workspace, MIS_ERR_INVALID_ARG, "mis_ntxent_fwd_bwd: null pointer");
  const int64_t need = mis_ntxent_fwd_bwd_workspace_bytes(rows, D);
  MIS_REQUIRE(need > 0 && workspace_bytes >= need, MIS_ERR_INVALID_ARG,
              "mis_ntxent_fwd_bwd: workspace of %lld bytes, %lld needed", (long long)workspace_bytes, (long long)need);
  MIS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, MIS_ERR_INVALID_ARG,
              "mis_ntxent_fwd_bwd: workspace must be 256-byte aligned");
  uint8_t* w = static_cast<uint8_t*>(workspace);
  float* u = reinterpret_cast<float*>(w);
  float* rinv = reinterpret_cast<float*>(w + ws_u_bytes(rows, D));
  float* lse = reinterpret_cast<float*>(w + ws_u_bytes(rows, D) + ws_row_bytes(rows));
  void* scratch = w + ws_u_bytes(rows, D) + 2 * ws_row_bytes(rows);
  const int64_t sc = mis_ntxent_scratch_bytes(rows, rows, D);
  const int rp = pad_rows(rows);
  if (int rc = mis_ntxent_prep(z, z_dtype, rows, D, u, rinv, stream)) return rc;
  if (int rc = mis_ntxent_fwd(u, rp, D, 0, rows, inv_T, lse, loss, scratch, sc, stream)) return rc;
  return mis_ntxent_bwd(u, lse, z, z_dtype, rinv, rp, D, 0, rows, inv_T, 1.0f, nullptr, dz, scratch, sc, stream);
}

extern "C" int mis_byol_loss_fwd_bwd(const float* preds, const float* targets, int rows, int D, float* loss,
                                     float* dpreds, float* scratch_rows, void* stream) {
  MIS_REQUIRE(preds && targets && loss && scratch_rows, MIS_ERR_INVALID_ARG, "mis_byol_loss_fwd_bwd: null pointer");
  MIS_REQUIRE(rows > 0 && D > 0, MIS_ERR_INVALID_ARG, "mis_byol_loss_fwd_bwd: sizes must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;
  byol_rows_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, st>>>(preds, targets, rows, D, scratch_rows, dpreds);
  MIS_CUDA_TRY(cudaGetLastError());
  byol_reduce_kernel<<<1, 1024, 0, st>>>(scratch_rows, rows, loss);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}
