// Weighted kNN prediction of the online evaluator.
//
// Reference: KNNOnlineEvaluator.predict (train/callback/knn.py:38-70):
//     sim = query @ bank.T                       [B, N]   (features are L2-normalised by the caller, :100, :129)
//     w, idx = sim.topk(k)                       [B, k]
//     labels = target_bank[idx];  w = exp(w / T)
//     scores[b, c] = sum_j w[b, j] * (labels[b, j] == c)
//     return scores.argsort(dim=-1, descending=True)
//
// Two steps on the GPU:
//  1. sim = Q . Bank^T with the tcgen05 TF32 GEMM of the loss path (wu_gemm_kernel, ntxent.cu): 128 x 256 tiles, accumulator
//     in TMEM, written once to a padded [pad128(B), pad256(N)] fp32 matrix.  A plain TF32 product is not enough here:
//     the votes are exp(sim / T) with T = 0.07, so a similarity error of 3e-4 is a vote error of 4e-3.  Each operand is
//     therefore split into hi = tf32(x) and lo = tf32(x - hi) and the GEMM runs over K' = 3 D on the concatenations
//     [hi | hi | lo] x [hi | lo | hi] (the lo.lo term, 2^-22 relative, is dropped): fp32-grade similarities from the
//     TF32 tensor cores, at three times a GEMM cost that is negligible next to reading the bank.
//  2. knn_vote_kernel, one CTA per query: exact selection of the k largest similarities in two reads of the row (per-thread
//     maxima give a lower bound of the k-th largest that only ~1.4 k elements reach; those are gathered and sorted; rows
//     with masses of equal similarities take an MSB-first radix select instead), the k (index, similarity) pairs in shared memory
//     and ordered by bank index (so that the result does not depend on thread timing), exp(sim / T) votes accumulated per
//     class in that order, and a bitonic sort of the (score, class) pairs: score descending, class ascending among equal
//     scores (torch's argsort leaves the order of equal scores -- e.g. all the classes without a vote -- unspecified).
//     Ties at the k-th similarity are broken towards the lower bank index.
#include <stdint.h>

#include "common.cuh"

namespace mis {
int launch_gemm_tf32_nt(const float* A, int M, const float* B, int N, int K, float* out, cudaStream_t st);

namespace knn {

constexpr int kThreads = 512;
constexpr int kMaxK = 1024;
constexpr int kMaxClasses = 8192;
constexpr int kCand = 4096;          // candidate list of the fast selection path (>= 2048: it first holds the thread maxima)

__device__ __forceinline__ uint32_t order_key(float s) {     // larger similarity -> larger key
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}

// src [rows, D] -> dst [rows, 3 D]: [hi | hi | lo] (second_lo == 0, the queries) or [hi | lo | hi] (the bank)
__global__ void __launch_bounds__(256) knn_split_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                                        int D, int second_lo) {
  const int64_t n4 = rows * (D >> 2);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / (D >> 2);
    const int c4 = (int)(i - r * (D >> 2));
    const float4 x = reinterpret_cast<const float4*>(src)[i];
    float4 hi, lo;
    hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
    lo.x = tf32_rna(x.x - hi.x); lo.y = tf32_rna(x.y - hi.y); lo.z = tf32_rna(x.z - hi.z); lo.w = tf32_rna(x.w - hi.w);
    float4* d = reinterpret_cast<float4*>(dst + r * 3 * D) + c4;
    d[0] = hi;
    d[D >> 2] = second_lo ? lo : hi;
    d[2 * (D >> 2)] = second_lo ? hi : lo;
  }
}

struct VoteArgs {
  const float* sim;          // [B_pad, ld] similarities (row b = query b)
  int64_t ld;
  int n_bank, k, k_pad, num_classes, c_pad;
  float inv_T;
  const int64_t* labels;     // [n_bank]
  int64_t* pred;             // [B, num_classes]
  float* scores;             // [B, num_classes] or null
};

__global__ void __launch_bounds__(kThreads) knn_vote_kernel(const VoteArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem);                       // [256]
  uint32_t* ctl = hist + 256;                                               // [8]: prefix, remaining, n_sel, n_tie, ...
  uint64_t* sel = reinterpret_cast<uint64_t*>(ctl + 8);                     // [k_pad]  (index << 32) | similarity bits
  uint64_t* sc = sel + a.k_pad;                                             // [c_pad]  sortable (score, class) pairs
  int* lab_s = reinterpret_cast<int*>(sc + a.c_pad);                        // [k_pad]  label of the j-th selected neighbour
  float* w_s = reinterpret_cast<float*>(lab_s + a.k_pad);                   // [k_pad]  exp(sim / T)
  uint64_t* cand = reinterpret_cast<uint64_t*>(w_s + a.k_pad);              // [kCand]  (key << 32) | ~index, sortable
  const int tid = threadIdx.x;
  const float* row = a.sim + (size_t)blockIdx.x * a.ld;
  const int n = a.n_bank;

  // ---- selection, fast path: two reads of the row, no atomics on the row --------------------------------------------
  // Each thread keeps the m largest keys of its share (m = ceil(2k / threads), so the k-th largest of those
  // threads*m values, T0, is a lower bound of the true k-th largest that only ~1.4 k elements of the row reach); the
  // elements >= T0 are gathered and sorted (key descending, bank index ascending) and the first k are the neighbours.
  // A row with more than kCand elements >= T0 (masses of equal similarities) takes the radix select below instead.
  for (int i = tid; i < a.k_pad; i += kThreads) sel[i] = ~0ull;             // padding sorts last
  bool fast = false;
  {
    const int m = (2 * a.k + kThreads - 1) / kThreads;                      // 1 .. 4
    uint32_t top[4] = {0u, 0u, 0u, 0u};
    for (int i = tid; i < n; i += kThreads) {
      uint32_t key = order_key(row[i]);
      if (key > top[3]) {                                                   // insertion into the sorted quadruple
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t hi = max(top[j], key);
          key = min(top[j], key);
          top[j] = hi;
        }
      }
    }
    const int nv = kThreads * m;                                            // <= 2048 values, a power of two times m
    const int nv_pad = 2048;
    uint32_t* vals = reinterpret_cast<uint32_t*>(cand);                     // the candidate list is not in use yet
    for (int i = tid; i < nv_pad; i += kThreads) vals[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < m) vals[tid * m + j] = top[j];
    __syncthreads();
    for (int size = 2; size <= nv_pad; size <<= 1) {                        // descending bitonic sort of the thread maxima
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < nv_pad; i += kThreads) {
          const int j = i ^ stride;
          if (j > i) {
            const uint32_t x = vals[i], y = vals[j];
            const bool down = (i & size) == 0;
            if ((x < y) == down) {
              vals[i] = y;
              vals[j] = x;
            }
          }
        }
        __syncthreads();
      }
    }
    const uint32_t t0 = vals[a.k - 1];                                      // (k <= n guarantees k real values, nv >= 2k)
    (void)nv;
    if (tid == 0) ctl[2] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {
      const uint32_t key = order_key(row[i]);
      if (key >= t0) {
        const uint32_t slot = atomicAdd(&ctl[2], 1u);
        if (slot < (uint32_t)kCand) cand[slot] = ((uint64_t)key << 32) | (uint32_t)(~(uint32_t)i);
      }
    }
    __syncthreads();
    const int nc = (int)ctl[2];
    fast = nc <= kCand;                                                     // CTA-uniform
    if (fast) {
      int nc_pad = 1;
      while (nc_pad < nc) nc_pad <<= 1;
      for (int i = nc + tid; i < nc_pad; i += kThreads) cand[i] = 0ull;     // padding sorts last (descending)
      __syncthreads();
      for (int size = 2; size <= nc_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int i = tid; i < nc_pad; i += kThreads) {
            const int j = i ^ stride;
            if (j > i) {
              const uint64_t x = cand[i], y = cand[j];
              const bool down = (i & size) == 0;
              if ((x < y) == down) {
                cand[i] = y;
                cand[j] = x;
              }
            }
          }
          __syncthreads();
        }
      }
      for (int j = tid; j < a.k; j += kThreads) {                           // larger key first, lower bank index among equal keys
        const uint32_t idx = ~(uint32_t)cand[j];
        sel[j] = ((uint64_t)idx << 32) | __float_as_uint(row[idx]);
      }
    }
    __syncthreads();
  }

  if (!fast) {
  // ---- k-th largest key: MSB-first radix select -------------------------------------------------------------------
  uint32_t prefix = 0, mask = 0;
  int remaining = a.k;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += kThreads) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {
      const uint32_t key = order_key(row[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, b = 255;
      for (; b > 0; --b) {
        if (cum + (int)hist[b] >= remaining) break;
        cum += (int)hist[b];
      }
      ctl[0] = prefix | ((uint32_t)b << shift);
      ctl[1] = (uint32_t)(remaining - cum);     // still to take among the keys that share the new prefix
    }
    __syncthreads();
    prefix = ctl[0];
    remaining = (int)ctl[1];
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t thr = prefix;                  // key of the k-th largest similarity; `remaining` of its ties are taken

  // ---- gather the selected pairs ----------------------------------------------------------------------------------
  if (tid == 0) ctl[2] = 0, ctl[3] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {
    const float s = row[i];
    const uint32_t key = order_key(s);
    if (key > thr) {
      const uint32_t slot = atomicAdd(&ctl[2], 1u);
      sel[slot] = ((uint64_t)(uint32_t)i << 32) | __float_as_uint(s);
    } else if (key == thr) {
      atomicAdd(&ctl[3], 1u);
    }
  }
  __syncthreads();
  const int n_above = (int)ctl[2], n_tie = (int)ctl[3];
  if (n_tie == remaining) {                     // the usual case: every element at the threshold is taken
    for (int i = tid; i < n; i += kThreads) {
      const float s = row[i];
      if (order_key(s) == thr) {
        const uint32_t slot = atomicAdd(&ctl[2], 1u);
        sel[slot] = ((uint64_t)(uint32_t)i << 32) | __float_as_uint(s);
      }
    }
  } else if (tid < 32) {                        // more ties than places: the lowest bank indices win (one warp, in order)
    int taken = 0;
    for (int i0 = 0; i0 < n && taken < remaining; i0 += 32) {
      const int i = i0 + tid;
      const float s = i < n ? row[i] : 0.f;
      const bool tie = i < n && order_key(s) == thr;
      const uint32_t m = __ballot_sync(0xffffffffu, tie);
      const int rank = taken + __popc(m & ((1u << tid) - 1u));
      if (tie && rank < remaining) sel[n_above + rank] = ((uint64_t)(uint32_t)i << 32) | __float_as_uint(s);
      taken += __popc(m);
    }
  }
  __syncthreads();

  }

  // ---- order the pairs by bank index (bitonic, ascending; padding = ~0 sorts last) ----------------------------------
  for (int size = 2; size <= a.k_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < a.k_pad; i += kThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const uint64_t x = sel[i], y = sel[j];
          const bool up = (i & size) == 0;
          if ((x > y) == up) {
            sel[i] = y;
            sel[j] = x;
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- votes: thread c sums the weights of its classes in index order (deterministic) ------------------------------
  for (int j = tid; j < a.k; j += kThreads) {
    const uint64_t e = sel[j];
    const int64_t lab = a.labels[(uint32_t)(e >> 32)];
    lab_s[j] = (lab >= 0 && lab < a.num_classes) ? (int)lab : -1;          // a label outside [0, C) votes for nothing
    w_s[j] = expf(__uint_as_float((uint32_t)e) * a.inv_T);
  }
  __syncthreads();
  for (int c = tid; c < a.c_pad; c += kThreads) {
    float acc = 0.f;
    if (c < a.num_classes) {
      for (int j = 0; j < a.k; ++j)
        if (lab_s[j] == c) acc += w_s[j];
      if (a.scores) a.scores[(size_t)blockIdx.x * a.num_classes + c] = acc;
      // sortable pair: score descending, class ascending <=> ascending in (~score_bits, class); scores are >= 0
      sc[c] = ((uint64_t)(~__float_as_uint(acc)) << 32) | (uint32_t)c;
    } else {
      sc[c] = ~0ull;
    }
  }
  __syncthreads();
  for (int size = 2; size <= a.c_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < a.c_pad; i += kThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const uint64_t x = sc[i], y = sc[j];
          const bool up = (i & size) == 0;
          if ((x > y) == up) {
            sc[i] = y;
            sc[j] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int c = tid; c < a.num_classes; c += kThreads)
    a.pred[(size_t)blockIdx.x * a.num_classes + c] = (int64_t)(uint32_t)sc[c];
}

static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
static inline int64_t pad_to(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

}  // namespace knn
}  // namespace mis

using namespace mis;

// scratch layout: similarity matrix [pad128(B), pad256(N)] | split queries [B, 3 D] | split bank [N, 3 D]
extern "C" int64_t mis_knn_scratch_bytes(int n_query, int n_bank, int D) {
  if (n_query <= 0 || n_bank <= 0 || D <= 0) return -1;
  return knn::pad_to(knn::pad_to(n_query, 128) * knn::pad_to(n_bank, 256) * 4, 256) +
         knn::pad_to((int64_t)n_query * 3 * D * 4, 256) + knn::pad_to((int64_t)n_bank * 3 * D * 4, 256);
}

extern "C" int mis_knn_predict(const float* query, const float* bank, const int64_t* bank_labels, int n_query, int n_bank,
                               int D, int k, float inv_T, int num_classes, int64_t* pred_labels, float* pred_scores,
                               void* scratch, int64_t scratch_bytes, void* stream) {
  using namespace mis::knn;
  MIS_REQUIRE(query && bank && bank_labels && pred_labels && scratch, MIS_ERR_INVALID_ARG, "mis_knn_predict: null pointer");
  MIS_REQUIRE(n_query > 0 && n_bank > 0, MIS_ERR_INVALID_ARG, "mis_knn_predict: sizes must be positive");
  MIS_REQUIRE(D >= 32 && D % 32 == 0, MIS_ERR_UNSUPPORTED, "mis_knn_predict: D=%d must be a multiple of 32", D);
  MIS_REQUIRE(k >= 1 && k <= n_bank, MIS_ERR_INVALID_ARG, "mis_knn_predict: k=%d outside [1, %d] (torch.topk raises too)", k,
              n_bank);
  MIS_REQUIRE(k <= kMaxK, MIS_ERR_UNSUPPORTED, "mis_knn_predict: k=%d > %d", k, kMaxK);
  MIS_REQUIRE(num_classes >= 1 && num_classes <= kMaxClasses, MIS_ERR_UNSUPPORTED,
              "mis_knn_predict: num_classes=%d outside [1, %d]", num_classes, kMaxClasses);
  MIS_REQUIRE(inv_T > 0.f, MIS_ERR_INVALID_ARG, "mis_knn_predict: temperature must be positive");
  MIS_REQUIRE(scratch_bytes >= mis_knn_scratch_bytes(n_query, n_bank, D), MIS_ERR_INVALID_ARG,
              "mis_knn_predict: scratch too small (%lld < %lld)", (long long)scratch_bytes,
              (long long)mis_knn_scratch_bytes(n_query, n_bank, D));
  MIS_REQUIRE(((reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(bank) | reinterpret_cast<uintptr_t>(scratch)) & 15) == 0,
              MIS_ERR_INVALID_ARG, "mis_knn_predict: query, bank and scratch must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* sc8 = static_cast<uint8_t*>(scratch);
  float* sim = reinterpret_cast<float*>(sc8);
  float* q3 = reinterpret_cast<float*>(sc8 + pad_to(pad_to(n_query, 128) * pad_to(n_bank, 256) * 4, 256));
  float* b3 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(q3) + pad_to((int64_t)n_query * 3 * D * 4, 256));
  {
    const int64_t nq4 = (int64_t)n_query * (D >> 2), nb4 = (int64_t)n_bank * (D >> 2);
    const unsigned gq = (unsigned)((nq4 + 255) / 256 < 4096 ? (nq4 + 255) / 256 : 4096);
    const unsigned gb = (unsigned)((nb4 + 255) / 256 < 4096 ? (nb4 + 255) / 256 : 4096);
    knn_split_kernel<<<gq, 256, 0, st>>>(query, q3, n_query, D, 0);
    knn_split_kernel<<<gb, 256, 0, st>>>(bank, b3, n_bank, D, 1);
    MIS_CUDA_TRY(cudaGetLastError());
  }
  if (int rc = launch_gemm_tf32_nt(q3, n_query, b3, n_bank, 3 * D, sim, st)) return rc;
  VoteArgs a = {};
  a.sim = sim;
  a.ld = pad_to(n_bank, 256);
  a.n_bank = n_bank;
  a.k = k;
  a.k_pad = next_pow2(k);
  a.num_classes = num_classes;
  a.c_pad = next_pow2(num_classes);
  a.inv_T = inv_T;
  a.labels = bank_labels;
  a.pred = pred_labels;
  a.scores = pred_scores;
  const size_t smem = (256 + 8) * 4 + (size_t)a.k_pad * 8 + (size_t)a.c_pad * 8 + (size_t)a.k_pad * 8 + (size_t)kCand * 8;
  MIS_CUDA_TRY(cudaFuncSetAttribute(knn_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn_vote_kernel<<<dim3((unsigned)n_query), kThreads, smem, st>>>(a);
  MIS_CUDA_TRY(cudaGetLastError());
  return MIS_OK;
}
