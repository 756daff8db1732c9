// K3  : approximate (centroid-only) document scores                       (search.rs:554-592)
//         approx[d] = sum_{q<Q}^{fp32}  max_{t<len(d)}  S[b][code[d,t]][q]     (fp16 max)
// K3b : pruning to the n_full_scores/4 best candidates                     (search.rs:602-619)
//
// The reference gathers S rows into a [tokens, Q] tensor, pads it to [2000, maxlen, Q], masks, maxes and
// sums, 128 times per query with two host syncs each.  A one-pass GPU formulation (one warp walks one
// candidate, every token gathers its Qp*2-byte S row) is bound by the L1 data pipe: one wavefront per
// gathered row, 4.87 G rows per batch on cfg-3 = 16.8 ms at one wavefront per clock per SM, and the kernel
// sat at 97 % of that (profiles/r01c_k3_approx_nsh_raw.csv).  Going faster needs FEWER ROWS, exactly:
//
//   1. tau[b,q]   = a quantile of the per-128-centroid-tile column maxima K1 already produces
//                   (any value is correct; it only moves work between the two passes)
//   2. hi[b,c]    = exists q: S[b,c,q] >= tau[b,q]          (a K-bit map per query, in shared memory)
//   3. bound pass : walk every candidate, test one bit per token, gather ONLY the rows of high centroids
//                   (~16 % of the tokens at the median tile maximum).  With m_q = max over the gathered rows:
//                     m_q >= tau_q  =>  m_q is the true column maximum (every skipped row is < tau_q <= m_q)
//                     otherwise     =>  m_q <= true maximum < tau_q
//                   so   lb = sum_q m_q  <=  approx  <=  sum_q max(m_q, tau_q) = ub,  with lb == ub == approx
//                   bit for bit when every column is resolved (fp32 addition is monotone in each operand and
//                   both sums use the summation order of the exact kernel).
//   4. threshold  : T = a value such that at least n_full_scores/4 candidates have lb >= T.
//   5. exact pass : the unresolved candidates with ub >= T are re-scored with all their rows.
//   A candidate left with an upper bound has approx <= ub < T <= the scores of >= n_full_scores/4 others:
//   it cannot enter the pruned list whatever the tie rule, and K3b (which only orders by value) never
//   selects it.  The pruned list and everything after it are bit-identical to scoring every candidate.
#include <math.h>
#include <stdlib.h>

#include "kernels.h"

namespace {

constexpr int64_t K3_TWO_PASS_MIN_TOKENS = 200000000;  // B * index tokens below which one pass is used by default
constexpr int K3_THREADS = 256;
constexpr int K3_DOCS_PER_CHUNK = 64;    // exact pass: work-queue granule
constexpr int K3A_DOCS_PER_CHUNK = 128;  // bound pass
// bound pass: per-warp ring of high codes waiting for their gather (a power of two >= one group + one batch)
constexpr int K3_WQ_FOR(int W, int FLUSH) { int n = 64; while (n < 32 * W + FLUSH) n <<= 1; return n; }

// chunk prefix of the dynamic work queue: work[b] = first chunk of query b, work[B] = total, work[B+1] = counter
__global__ void k3_prefix_kernel(const int32_t* __restrict__ n_items, int B, int per_chunk,
                                 int32_t* __restrict__ work) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      work[b] = acc;
      acc += (n_items[b] + per_chunk - 1) / per_chunk;
    }
    work[B] = acc;
    work[B + 1] = 0;
  }
}

// next (query, chunk) of the queue; s_b = -1 when it is empty.  Called by every thread of the CTA.
__device__ __forceinline__ void k3_next_chunk(int32_t* work, int B, int* s_b, int* s_c) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const int c = atomicAdd(&work[B + 1], 1);
    if (c >= work[B]) {
      *s_b = -1;
    } else {
      int lo = 0, hi = B - 1;  // largest b with work[b] <= c
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (work[mid] <= c) lo = mid; else hi = mid - 1;
      }
      *s_b = lo;
      *s_c = c - work[lo];
    }
  }
  __syncthreads();
}

// maxima of the four half2 registers across the lane groups, then the fp32 sum over the real query tokens
// (sum_dim_intlist(.., Kind::Float), search.rs:401).  The order of the additions is part of the contract between
// the bound pass and the exact pass: both call this.
template <int LPR>
__device__ __forceinline__ void k3_reduce_groups(__half2& m0, __half2& m1, __half2& m2, __half2& m3) {
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1) {
    m0 = __hmax2(m0, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m0), off)));
    m1 = __hmax2(m1, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m1), off)));
    m2 = __hmax2(m2, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m2), off)));
    m3 = __hmax2(m3, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m3), off)));
  }
}
template <int LPR, bool FULLQ = false>
__device__ __forceinline__ float k3_sum_columns(__half2 m0, __half2 m1, __half2 m2, __half2 m3, int col0, int Q) {
  float s = 0.f;
  const float2 f0 = __half22float2(m0), f1 = __half22float2(m1), f2 = __half22float2(m2), f3 = __half22float2(m3);
  if constexpr (FULLQ) {  // Q == Qp: every column is a real query token; same additions in the same order
    s += f0.x; s += f0.y; s += f1.x; s += f1.y; s += f2.x; s += f2.y; s += f3.x; s += f3.y;
#pragma unroll
    for (int off = 1; off < LPR; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
  }
  if (col0 + 0 < Q) s += f0.x;
  if (col0 + 1 < Q) s += f0.y;
  if (col0 + 2 < Q) s += f1.x;
  if (col0 + 3 < Q) s += f1.y;
  if (col0 + 4 < Q) s += f2.x;
  if (col0 + 5 < Q) s += f2.y;
  if (col0 + 6 < Q) s += f3.x;
  if (col0 + 7 < Q) s += f3.y;
#pragma unroll
  for (int off = 1; off < LPR; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

// ---------------------------------------------------------------------------------------
// Exact pass / one-pass scoring.  One warp walks one candidate; each S row (Qp fp16 = LPR x 16 B) is fetched
// by LPR adjacent lanes, so a warp-wide load touches 32/LPR distinct rows, and the running maxima stay in
// registers.  `list` (the bound pass's refine list) selects the candidates; NULL = all of them.
// Two code-distribution schemes: shuffles (any Qp) or, for Qp <= 32, shuffle-free vector loads of the
// LPR consecutive codes a lane group gathers (the __shfl_sync run through the same L1 data pipe as the gathers).
// ---------------------------------------------------------------------------------------
template <int LPR, int UNROLL, int MINB>
__global__ void __launch_bounds__(K3_THREADS, MINB)
k3_approx_kernel(const __half* __restrict__ S, int64_t K, int Q, const int64_t* __restrict__ doc_offsets,
                 const int32_t* __restrict__ codes, const int32_t* __restrict__ cand, int cand_cap,
                 const int32_t* __restrict__ n_cand, const int32_t* __restrict__ list,
                 const int32_t* __restrict__ n_list, int32_t* __restrict__ work, int B,
                 float* __restrict__ approx, unsigned long long* __restrict__ stats) {
  constexpr int QP = LPR * 8;
  constexpr int TPI = 32 / LPR;
  __shared__ int s_b, s_c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const __half2 sentinel = __float2half2_rn(FPB_PAD_SENTINEL);
  unsigned long long rows = 0;

  for (;;) {
    k3_next_chunk(work, B, &s_b, &s_c);
    const int b = s_b;
    if (b < 0) break;
    const int n = list ? n_list[b] : n_cand[b];
    const uint4* Sb = reinterpret_cast<const uint4*>(S + int64_t(b) * K * QP);
    const int32_t* cb = cand + int64_t(b) * cand_cap;
    const int32_t* lb = list ? list + int64_t(b) * cand_cap : nullptr;
    float* ab = approx + int64_t(b) * cand_cap;

    for (int i = 0; i < K3_DOCS_PER_CHUNK / 8; ++i) {
      const int j = s_c * K3_DOCS_PER_CHUNK + i * 8 + warp;
      if (j >= n) break;
      const int idx = lb ? lb[j] : j;
      const int d = cb[idx];
      const int64_t o0 = doc_offsets[d];
      const int len = int(doc_offsets[d + 1] - o0);
      rows += unsigned(len);
      __half2 m0 = sentinel, m1 = sentinel, m2 = sentinel, m3 = sentinel;
      for (int base = 0; base < len; base += 32 * UNROLL) {
        int code[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int t = base + u * 32 + lane;
          code[u] = (t < len) ? __ldg(codes + o0 + t) : -1;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
          for (int jj = 0; jj < LPR; ++jj) {
            const int c = __shfl_sync(0xffffffffu, code[u], jj * TPI + grp);
            if (c >= 0) {
              const uint4 v = __ldg(Sb + int64_t(c) * LPR + sub);
              m0 = __hmax2(m0, u32_as_half2(v.x));
              m1 = __hmax2(m1, u32_as_half2(v.y));
              m2 = __hmax2(m2, u32_as_half2(v.z));
              m3 = __hmax2(m3, u32_as_half2(v.w));
            }
          }
        }
      }
      k3_reduce_groups<LPR>(m0, m1, m2, m3);
      const float s = k3_sum_columns<LPR>(m0, m1, m2, m3, sub * 8, Q);
      if (lane == 0) ab[idx] = s;
    }
  }
  if (stats && lane == 0 && rows) atomicAdd(stats + 2, rows);
}

// Shuffle-free variant: the LPR lanes of a group load "their" LPR consecutive codes themselves with one
// vector load (the lanes of a group read the same 4*LPR bytes: a broadcast); the document is walked in
// 32-token windows aligned to absolute multiples of 32 tokens, so the vector loads are aligned whatever the
// document offset; tokens outside [o0, o0+len) are masked.  max() is order-free: bit-identical values.
template <int LPR>
__device__ __forceinline__ void load_codes(int (&c)[LPR], const int32_t* p) {
  if constexpr (LPR == 2) {
    const int2 v = __ldg(reinterpret_cast<const int2*>(p));
    c[0] = v.x; c[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < LPR / 4; ++i) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(p) + i);
      c[4 * i + 0] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w;
    }
  }
}

template <int LPR, int UNROLL, int MINB>
__global__ void __launch_bounds__(K3_THREADS, MINB)
k3_approx_nsh_kernel(const __half* __restrict__ S, int64_t K, int Q, const int64_t* __restrict__ doc_offsets,
                     const int32_t* __restrict__ codes, int64_t n_codes, const int32_t* __restrict__ cand,
                     int cand_cap, const int32_t* __restrict__ n_cand, const int32_t* __restrict__ list,
                     const int32_t* __restrict__ n_list, int32_t* __restrict__ work, int B,
                     float* __restrict__ approx, unsigned long long* __restrict__ stats) {
  constexpr int QP = LPR * 8;
  __shared__ int s_b, s_c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const __half2 sentinel = __float2half2_rn(FPB_PAD_SENTINEL);
  unsigned long long rows = 0;

  for (;;) {
    k3_next_chunk(work, B, &s_b, &s_c);
    const int b = s_b;
    if (b < 0) break;
    const int n = list ? n_list[b] : n_cand[b];
    const uint4* Sb = reinterpret_cast<const uint4*>(S + int64_t(b) * K * QP) + sub;
    const int32_t* cb = cand + int64_t(b) * cand_cap;
    const int32_t* lb = list ? list + int64_t(b) * cand_cap : nullptr;
    float* ab = approx + int64_t(b) * cand_cap;

    for (int i = 0; i < K3_DOCS_PER_CHUNK / 8; ++i) {
      const int j = s_c * K3_DOCS_PER_CHUNK + i * 8 + warp;
      if (j >= n) break;
      const int idx = lb ? lb[j] : j;
      const int d = cb[idx];
      const int64_t o0 = doc_offsets[d];
      const int len = int(doc_offsets[d + 1] - o0);
      rows += unsigned(len);
      // frame: token f of the frame is absolute token w0 + f; the document is [lo, hi)
      const int64_t w0 = o0 & ~int64_t(31);
      const int lo = int(o0 - w0), hi = lo + len;
      const int32_t* cw = codes + w0 + grp * LPR;
      // the last window may reach past the end of the code array: lanes whose LPR codes are not all
      // inside it take the scalar path (at most once per index)
      const int64_t readable = n_codes - (w0 + grp * LPR);
      __half2 m0 = sentinel, m1 = sentinel, m2 = sentinel, m3 = sentinel;
      for (int f0 = 0; f0 < hi; f0 += 32 * UNROLL) {
        int c[UNROLL][LPR];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int fo = f0 + u * 32;
          if (fo + LPR <= readable) {
            load_codes<LPR>(c[u], cw + fo);
          } else {
#pragma unroll
            for (int jj = 0; jj < LPR; ++jj) c[u][jj] = (fo + jj < readable) ? __ldg(cw + fo + jj) : 0;
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
          for (int jj = 0; jj < LPR; ++jj) {
            const int f = f0 + u * 32 + grp * LPR + jj;
            if (unsigned(f - lo) < unsigned(len)) {
              const uint4 v = __ldg(Sb + int64_t(c[u][jj]) * LPR);
              m0 = __hmax2(m0, u32_as_half2(v.x));
              m1 = __hmax2(m1, u32_as_half2(v.y));
              m2 = __hmax2(m2, u32_as_half2(v.z));
              m3 = __hmax2(m3, u32_as_half2(v.w));
            }
          }
        }
      }
      k3_reduce_groups<LPR>(m0, m1, m2, m3);
      const float s = k3_sum_columns<LPR>(m0, m1, m2, m3, sub * 8, Q);
      if (lane == 0) ab[idx] = s;
    }
  }
  if (stats && lane == 0 && rows) atomicAdd(stats + 2, rows);
}

// ---------------------------------------------------------------------------------------
// Two-pass scheme, step 1: tau[b,q].  Any value is correct; what it should be is decided by the candidates, not by
// the centroid table: with tau_q at the level that a candidate document holds on average LAMBDA tokens at or above
// it, a column stays unresolved with probability ~exp(-LAMBDA) and a row is gathered with probability
// ~Q*LAMBDA/len, whatever the score distribution (uniform codes, clustered topics, short documents).  So tau_q is
// estimated from a sample: up to K3_TAU_DOCS candidates per query (evenly spaced), every token's S row, and per
// column the (LAMBDA * #sampled docs)-th largest value, found with a two-level 256-bin radix select on the 16-bit
// order-preserving keys.  One CTA per query; columns are handled 32 at a time.
// ---------------------------------------------------------------------------------------
constexpr int K3_TAU_DOCS = 64;
constexpr int K3_TAU_THREADS = 1024;

// The level-1 histogram costs one shared-memory atomic per sampled value; almost all of them fall far below the
// answer.  The smallest of the column's tile maxima (K1 writes one per 128 centroids) sits near the 94th percentile
// of the column, well below any useful tau, so values under it are not counted; if that ever leaves fewer than the
// wanted rank (tau would be below the floor), the level is redone without a floor.
template <int LPR>
__global__ void __launch_bounds__(K3_TAU_THREADS)
k3_tau_kernel(const __half* __restrict__ S, int64_t K, int Q, const int64_t* __restrict__ doc_offsets,
              const int32_t* __restrict__ codes, const int32_t* __restrict__ cand, int cand_cap,
              const int32_t* __restrict__ n_cand, const __half* __restrict__ tmax, int n_tiles, float lambda,
              __half* __restrict__ tau) {
  constexpr int QP = LPR * 8, TPI = 32 / LPR;
  constexpr int SUBS = LPR < 4 ? LPR : 4;  // lane subs (8 columns each) handled per round
  __shared__ int hist[32][256];
  __shared__ int s_bin[32], s_rem[32];
  __shared__ uint32_t s_floor[32];
  __shared__ int s_redo;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  uint16_t* out = reinterpret_cast<uint16_t*>(tau) + int64_t(b) * QP;
  const int n = n_cand[b];
  const int ns = min(n, K3_TAU_DOCS);
  const int32_t* cb = cand + int64_t(b) * cand_cap;
  const uint4* Sb = reinterpret_cast<const uint4*>(S + int64_t(b) * K * QP);
  const int want = max(1, __float2int_rn(lambda * float(ns)));  // rank (from the top) of the selected value

  for (int g0 = 0; g0 < LPR; g0 += SUBS) {
    const bool mine = sub >= g0 && sub < g0 + SUBS;
    const int qrow = (sub - g0) * 8;  // first of this lane's 8 histogram rows
    __syncthreads();  // the previous round's last reads of s_redo / s_floor / s_bin are done
    // floor of the round's columns: the smallest tile maximum (one warp per column)
    for (int col = warp; col < SUBS * 8; col += K3_TAU_THREADS / 32) {
      const int q = g0 * 8 + col;
      uint32_t mn = 0xffffu;
      if (q < Q) {
        const uint16_t* tm = reinterpret_cast<const uint16_t*>(tmax) + (int64_t(b) * QP + q) * n_tiles;
        for (int i = lane; i < n_tiles; i += 32) mn = min(mn, f16_key(tm[i]));
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      if (lane == 0) s_floor[col] = mn;
    }
    if (tid == 0) s_redo = 0;
    __syncthreads();
    for (int level = 0; level < 2; ++level) {
      for (int i = tid; i < 32 * 256; i += K3_TAU_THREADS) (&hist[0][0])[i] = 0;
      __syncthreads();
      uint32_t fl[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) fl[e] = (mine && !s_redo) ? s_floor[qrow + e] : 0u;
      for (int j = warp; j < ns; j += K3_TAU_THREADS / 32) {
        const int d = cb[int64_t(j) * n / ns];
        const int64_t o0 = doc_offsets[d];
        const int len = int(doc_offsets[d + 1] - o0);
        for (int base = 0; base < len; base += 32) {
          const int t = base + lane;
          const int code = (t < len) ? __ldg(codes + o0 + t) : -1;
          // all LPR row gathers of the window go out before the first histogram update (the kernel is a chain of
          // dependent latencies otherwise: 0.33 ms with the loads issued one per update round)
          constexpr int JB = LPR < 8 ? LPR : 8;  // gathers in flight per lane
#pragma unroll 1
          for (int j0 = 0; j0 < LPR; j0 += JB) {
          uint4 vv[JB];
          int cc[JB];
#pragma unroll
          for (int jj = 0; jj < JB; ++jj) {
            cc[jj] = __shfl_sync(0xffffffffu, code, (j0 + jj) * TPI + grp);
            vv[jj] = make_uint4(0u, 0u, 0u, 0u);
            if (cc[jj] >= 0 && mine) vv[jj] = __ldg(Sb + int64_t(cc[jj]) * LPR + sub);
          }
#pragma unroll
          for (int jj = 0; jj < JB; ++jj) {
            const int c = cc[jj];
            if (c >= 0 && mine) {
              const uint4 v = vv[jj];
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t key = f16_key(uint16_t(w[e >> 1] >> ((e & 1) * 16)));
                if (level == 0) {
                  if (key >= fl[e]) atomicAdd(&hist[qrow + e][key >> 8], 1);
                } else if (int(key >> 8) == s_bin[qrow + e]) {
                  atomicAdd(&hist[qrow + e][key & 255u], 1);
                }
              }
            }
          }
          }
        }
      }
      __syncthreads();
      if (tid < 32) {  // one thread per column of the round: walk the bins from the top
        const int need = level == 0 ? want : s_rem[tid];
        int cum = 0, bin = 255;
        for (; bin > 0; --bin) {
          if (cum + hist[tid][bin] >= need) break;
          cum += hist[tid][bin];
        }
        if (level == 0) {
          s_bin[tid] = bin;
          s_rem[tid] = need - cum;
          // fewer counted values than the wanted rank although a floor was applied: count everything once more
          if (bin == 0 && cum + hist[tid][0] < need && !s_redo && tid < SUBS * 8 && g0 * 8 + tid < Q &&
              s_floor[tid] != 0u)
            atomicExch(&s_redo, 2);
        } else {
          const int q = g0 * 8 + tid;
          if (q < QP) {
            uint32_t key = (uint32_t(s_bin[tid]) << 8) | uint32_t(bin);
            // fewer sampled values than `want` (tiny documents): everything is "high" -> every column resolves
            if (s_bin[tid] == 0 && bin == 0) key = f16_key(0xFC00u);  // -inf
            uint16_t h = uint16_t((key & 0x8000u) ? (key & 0x7fffu) : (~key & 0xffffu));  // inverse of f16_key
            if (q >= Q) h = 0x7C00u;  // padded column: +inf, never "high", excluded from every sum
            if (q < QP && tid < SUBS * 8) out[q] = h;
          }
        }
      }
      __syncthreads();
      if (level == 0 && s_redo == 2) {  // CTA-uniform: the floor hid too much, redo level 0 without it
        __syncthreads();
        if (tid == 0) s_redo = 1;
        __syncthreads();
        level = -1;
      }
    }
  }
}

// step 2: hi[b,c] = exists q: S[b,c,q] >= tau[b,q].  One warp per 32 consecutive centroids = one bitmap word;
// LPR lanes per row, so every warp-wide load is 512 contiguous bytes of S.
template <int LPR>
__global__ void __launch_bounds__(256)
k3_hibits_kernel(const __half* __restrict__ S, int64_t K, const __half* __restrict__ tau,
                 uint32_t* __restrict__ hibits, int hb_words) {
  constexpr int QP = LPR * 8, TPI = 32 / LPR;
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int64_t word = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t c0 = word * 32;
  if (c0 >= K) return;  // warp-uniform
  const int sub = lane % LPR, grp = lane / LPR;
  const uint4 t = *reinterpret_cast<const uint4*>(tau + int64_t(b) * QP + sub * 8);
  const uint4* Sb = reinterpret_cast<const uint4*>(S + int64_t(b) * K * QP);
  constexpr uint32_t GMASK = (LPR == 32) ? 0xffffffffu : ((1u << (LPR & 31)) - 1u);
  bool mine = false;
#pragma unroll
  for (int j = 0; j < LPR; ++j) {
    const int64_t c = c0 + j * TPI + grp;
    bool h = false;
    if (c < K) {
      const uint4 v = ldg_nc_na(Sb + c * LPR + sub);
      const unsigned any = __hge2_mask(u32_as_half2(v.x), u32_as_half2(t.x)) |
                           __hge2_mask(u32_as_half2(v.y), u32_as_half2(t.y)) |
                           __hge2_mask(u32_as_half2(v.z), u32_as_half2(t.z)) |
                           __hge2_mask(u32_as_half2(v.w), u32_as_half2(t.w));
      h = any != 0u;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, h);
    // centroid c0 + l (l = this lane) was handled in iteration l / TPI by lane group l % TPI
    if (lane / TPI == j) mine = ((bal >> ((lane % TPI) * LPR)) & GMASK) != 0u;
  }
  const unsigned w = __ballot_sync(0xffffffffu, mine);
  if (lane == 0) hibits[int64_t(b) * hb_words + word] = w;
}

// step 3: the bound pass.  Shared memory: the query's K-bit map, one ring of high codes per warp, and the offsets /
// lengths of the chunk's documents (staged by the first K3A_DOCS_PER_CHUNK threads so that the dependent
// candidate -> offset loads are paid once per chunk, not once per document).  A warp walks its documents as a stream
// of W-window groups and always has the NEXT group's code loads in flight while it tests, queues and gathers the
// current one.  The first version of this kernel was ISSUE-bound (profiles/r02_k3_bound_v1_raw.csv: 15.3 G warp
// instructions, 91 per 32-token window, issue slots 81 % busy; a third of them branches and generic-address
// arithmetic around the shared-memory accesses), so the per-window path is written branch-free with explicit 32-bit
// shared addresses: predicated code load, one LDS of the bitmap word, a wrap-around funnel shift for the bit, ballot,
// predicated STS into the ring.  Tokens outside the document carry the code K3_INV whose bit is a spare zero word.
__device__ __forceinline__ uint32_t k3_lds(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// One window's queue step in a single block so that the bit is tested ONCE: p = bit c of the bitmap word, ballot,
// rank of this lane among the set lanes, predicated store of the code into the ring.  The ring is aligned to its
// size, so "position modulo the ring, plus its base" is one logic op.  Returns the ballot; tail_bytes is the
// warp-uniform fill pointer in bytes.
__device__ __forceinline__ unsigned k3_push(uint32_t word, uint32_t code, uint32_t ring_base, uint32_t ring_mask_bytes,
                                            uint32_t tail_bytes, uint32_t lt_mask) {
  unsigned mask;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 t, r;\n"
      "shf.r.wrap.b32 t, %1, 0, %2;\n"   // word >> (code & 31)
      "and.b32 t, t, 1;\n"
      "setp.ne.u32 p, t, 0;\n"
      "vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
      "and.b32 r, %0, %6;\n"
      "popc.b32 r, r;\n"
      "shl.b32 r, r, 2;\n"
      "add.u32 r, r, %5;\n"
      "and.b32 r, r, %4;\n"
      "or.b32 r, r, %3;\n"
      "@p st.shared.u32 [r], %2;\n"
      "}"
      : "=r"(mask)
      : "r"(word), "r"(code), "r"(ring_base), "r"(ring_mask_bytes), "r"(tail_bytes), "r"(lt_mask)
      : "memory");
  return mask;
}

template <int LPR, int MINB, int W, int U, bool FULLQ, bool PAIR>
__global__ void __launch_bounds__(K3_THREADS, MINB)
k3_bound_kernel(const __half* __restrict__ S, int64_t K, int Q, const int64_t* __restrict__ doc_offsets,
                const int32_t* __restrict__ codes, const int32_t* __restrict__ cand, int cand_cap,
                const int32_t* __restrict__ n_cand, int32_t* __restrict__ work, int B,
                const __half* __restrict__ tau, const uint32_t* __restrict__ hibits, int hb_words,
                float* __restrict__ ub_out, float* __restrict__ lb_out, unsigned long long* __restrict__ stats) {
  constexpr int QP = LPR * 8;
  constexpr int TPI = 32 / LPR;   // rows per warp-wide gather
  constexpr int FLUSH = U * TPI;  // rows per batch of gathers (<= 64)
  constexpr int DPW = K3A_DOCS_PER_CHUNK / (K3_THREADS / 32);  // documents per warp and chunk
  constexpr int WQ = K3_WQ_FOR(W, FLUSH);  // ring entries per warp: one group of pushes on top of an unflushed rest
  static_assert(FLUSH + 32 * W <= WQ, "ring too small");
  extern __shared__ __align__(16) uint32_t k3_smem[];  // [hb_words + 4] bitmap (+ a zero word) | [warps][WQ] rings
  __shared__ int s_b, s_c;
  __shared__ int64_t s_o0[K3A_DOCS_PER_CHUNK];
  __shared__ int s_len[K3A_DOCS_PER_CHUNK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const uint32_t sb_bm = smem_u32(k3_smem);
  // rings: aligned to their size (WQ * 4 bytes) so that k3_push can OR the base in
  const uint32_t sb_wq = ((sb_bm + uint32_t(hb_words + 4) * 4u + (WQ * 4u - 1u)) & ~(WQ * 4u - 1u)) + uint32_t(warp) * (WQ * 4u);
  const int INV = hb_words * 32;  // a code whose bit lives in the spare zero word
  unsigned lt_mask;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  const __half2 sentinel = __float2half2_rn(FPB_PAD_SENTINEL);
  unsigned rows = 0, toks = 0;  // per warp over the CTA's life: far below 2^32
  int cur_b = -1;
  if (tid < 4) k3_smem[hb_words + tid] = 0u;  // never overwritten: the bitmap copy below covers hb_words words

  for (;;) {
    k3_next_chunk(work, B, &s_b, &s_c);
    const int b = s_b;
    if (b < 0) break;
    const int n = n_cand[b];
    const int32_t* cb = cand + int64_t(b) * cand_cap;
    if (tid < K3A_DOCS_PER_CHUNK) {  // stage the chunk's document extents
      const int idx = s_c * K3A_DOCS_PER_CHUNK + tid;
      int64_t o0 = 0;
      int len = -1;  // -1: no such document
      if (idx < n) {
        const int d = cb[idx];
        o0 = doc_offsets[d];
        len = int(doc_offsets[d + 1] - o0);
      }
      s_o0[tid] = o0;
      s_len[tid] = len;
    }
    if (b != cur_b) {  // CTA-uniform; every warp is past the previous chunk (barrier in k3_next_chunk)
      const uint4* src = reinterpret_cast<const uint4*>(hibits + int64_t(b) * hb_words);
      for (int i = tid; i < hb_words / 4; i += K3_THREADS) reinterpret_cast<uint4*>(k3_smem)[i] = src[i];
      cur_b = b;
    }
    __syncthreads();
    const uint4* Sb = reinterpret_cast<const uint4*>(S + int64_t(b) * K * QP) + sub;
    const uint4* tqp = reinterpret_cast<const uint4*>(tau + int64_t(b) * QP + sub * 8);
    float* ub_chunk = ub_out + int64_t(b) * cand_cap + int64_t(s_c) * K3A_DOCS_PER_CHUNK;
    float* lb_chunk = lb_out + int64_t(b) * cand_cap + int64_t(s_c) * K3A_DOCS_PER_CHUNK;

    // ---- the warp's documents are slots warp*DPW .. warp*DPW + DPW - 1 of the chunk, walked group by group ----
    int slot = warp * DPW;
    const int slot_end = slot + DPW;
    int len = s_len[slot];
    if (len < 0) continue;  // warp-uniform: this warp has no document in the (last, partial) chunk
    // rel = index inside the document of this lane's token in window 0 of the group (negative before the start);
    // nwin = windows of the frame [0, lo + len) aligned to absolute multiples of 32 tokens (one 128-byte line each)
    int rel, left;  // left = windows of the document not yet walked (including the current group's)
    const int32_t* cw;
    {
      const int64_t o0 = s_o0[slot];
      const int lo = int(o0 & 31);
      rel = lane - lo;
      left = (lo + len + 31) >> 5;
      cw = codes + (o0 - lo) + lane;
    }
    int c[W];
#pragma unroll
    for (int u = 0; u < W; ++u) {
      c[u] = INV;
      if (unsigned(rel + 32 * u) < unsigned(len)) c[u] = __ldg(cw + 32 * u);
    }
    uint32_t head = 0, tail = 0;
    __half2 m0 = sentinel, m1 = sentinel, m2 = sentinel, m3 = sentinel;
    // PAIR: the end-of-document work (maxima across the lane groups, two column sums) is done for two documents at
    // once -- the first of a pair is parked in a0..a3, then the lower half of the warp finishes one document and the
    // upper half the other: half the shuffles, conversions and additions per document
    __half2 a0 = sentinel, a1 = sentinel, a2 = sentinel, a3 = sentinel;
    int slot_a = -1;

    for (;;) {
      // ---- the next group: same document or the next slot; its code loads go out now ----
      const bool last_of_doc = left <= W;
      int nlen = len, nrel = rel + 32 * W, nleft = left - W;
      const int32_t* ncw = cw + 32 * W;
      if (last_of_doc) {
        nlen = (slot + 1 < slot_end) ? s_len[slot + 1] : -1;
        if (nlen >= 0) {
          const int64_t o0 = s_o0[slot + 1];
          const int lo = int(o0 & 31);
          nrel = lane - lo;
          nleft = (lo + nlen + 31) >> 5;
          ncw = codes + (o0 - lo) + lane;
        }
      }
      int nx[W];
      {
        const unsigned nlen_u = nlen < 0 ? 0u : unsigned(nlen);
#pragma unroll
        for (int u = 0; u < W; ++u) {
          nx[u] = INV;
          if (unsigned(nrel + 32 * u) < nlen_u) nx[u] = __ldg(ncw + 32 * u);
        }
      }
      // ---- current group: bit of every token (W independent shared loads), queue the high codes ----
      uint32_t wbit[W];
#pragma unroll
      for (int u = 0; u < W; ++u) wbit[u] = k3_lds(sb_bm + ((uint32_t(c[u]) >> 5) << 2));
#pragma unroll
      for (int u = 0; u < W; ++u) {
        const unsigned mask = k3_push(wbit[u], uint32_t(c[u]), sb_wq, WQ * 4u - 1u, tail << 2, lt_mask);
        tail += __popc(mask);
      }
      __syncwarp();
      // ---- gather in full batches ----
      while (tail - head >= FLUSH) {
        uint4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const uint32_t code = k3_lds(sb_wq + (((head + k * TPI + grp) & (WQ - 1)) << 2));
          v[k] = __ldg(Sb + int64_t(code) * LPR);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          m0 = __hmax2(m0, u32_as_half2(v[k].x));
          m1 = __hmax2(m1, u32_as_half2(v[k].y));
          m2 = __hmax2(m2, u32_as_half2(v[k].z));
          m3 = __hmax2(m3, u32_as_half2(v[k].w));
        }
        head += FLUSH;
      }
      if (last_of_doc) {
        {  // drain: fewer than FLUSH codes left
          const uint32_t rem = tail - head;
          uint4 v[U];
#pragma unroll
          for (int k = 0; k < U; ++k) {
            const uint32_t jj = k * TPI + grp;
            v[k] = make_uint4(half2_as_u32(sentinel), half2_as_u32(sentinel), half2_as_u32(sentinel),
                              half2_as_u32(sentinel));
            if (jj < rem) {
              const uint32_t code = k3_lds(sb_wq + (((head + jj) & (WQ - 1)) << 2));
              v[k] = __ldg(Sb + int64_t(code) * LPR);
            }
          }
#pragma unroll
          for (int k = 0; k < U; ++k) {
            m0 = __hmax2(m0, u32_as_half2(v[k].x));
            m1 = __hmax2(m1, u32_as_half2(v[k].y));
            m2 = __hmax2(m2, u32_as_half2(v[k].z));
            m3 = __hmax2(m3, u32_as_half2(v[k].w));
          }
        }
        rows += tail;
        toks += unsigned(len);
        const uint4 tq = __ldg(tqp);
        const int col0 = sub * 8;
        // lower bound: the maxima over the gathered rows; upper bound: unresolved columns raised to tau.  Both sums
        // use the exact kernel's order, fp32 addition is monotone, so lb <= approx <= ub; when no column was raised
        // the two are the same additions of the same values, and whenever lb == ub the score is pinned between them:
        // "resolved" needs no flag, it IS lb == ub.
        if (PAIR && LPR <= 8 && slot_a < 0 && nlen >= 0) {
          a0 = m0; a1 = m1; a2 = m2; a3 = m3;  // park: the next document completes the pair
          slot_a = slot;
        } else if (PAIR && LPR <= 8 && slot_a >= 0) {
          const bool up = lane >= 16;  // lower half: the parked document, upper half: this one
          __half2 x0 = up ? m0 : a0, x1 = up ? m1 : a1, x2 = up ? m2 : a2, x3 = up ? m3 : a3;
          const __half2 y0 = up ? a0 : m0, y1 = up ? a1 : m1, y2 = up ? a2 : m2, y3 = up ? a3 : m3;
          x0 = __hmax2(x0, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(y0), 16)));
          x1 = __hmax2(x1, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(y1), 16)));
          x2 = __hmax2(x2, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(y2), 16)));
          x3 = __hmax2(x3, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(y3), 16)));
#pragma unroll
          for (int off = LPR; off < 16; off <<= 1) {
            x0 = __hmax2(x0, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(x0), off)));
            x1 = __hmax2(x1, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(x1), off)));
            x2 = __hmax2(x2, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(x2), off)));
            x3 = __hmax2(x3, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(x3), off)));
          }
          const float lb = k3_sum_columns<LPR, FULLQ>(x0, x1, x2, x3, col0, Q);
          const float ub = k3_sum_columns<LPR, FULLQ>(__hmax2(x0, u32_as_half2(tq.x)), __hmax2(x1, u32_as_half2(tq.y)),
                                                      __hmax2(x2, u32_as_half2(tq.z)), __hmax2(x3, u32_as_half2(tq.w)),
                                                      col0, Q);
          if ((lane & 15) == 0) {
            const int at = up ? slot : slot_a;
            ub_chunk[at] = ub;
            lb_chunk[at] = lb;
          }
          slot_a = -1;
        } else {
          k3_reduce_groups<LPR>(m0, m1, m2, m3);
          const float lb = k3_sum_columns<LPR, FULLQ>(m0, m1, m2, m3, col0, Q);
          const float ub = k3_sum_columns<LPR, FULLQ>(__hmax2(m0, u32_as_half2(tq.x)), __hmax2(m1, u32_as_half2(tq.y)),
                                                      __hmax2(m2, u32_as_half2(tq.z)), __hmax2(m3, u32_as_half2(tq.w)),
                                                      col0, Q);
          if (lane == 0) {
            ub_chunk[slot] = ub;
            lb_chunk[slot] = lb;
          }
        }
        if (nlen < 0) break;  // no further document for this warp in the chunk
        head = tail = 0;
        m0 = m1 = m2 = m3 = sentinel;
        ++slot;
      }
      __syncwarp();  // the ring slots read above may be overwritten by the next group's pushes
#pragma unroll
      for (int u = 0; u < W; ++u) c[u] = nx[u];
      len = nlen;
      rel = nrel;
      left = nleft;
      cw = ncw;
    }
  }
  if (stats && lane == 0) {
    if (rows) atomicAdd(stats + 0, (unsigned long long)rows);
    if (toks) atomicAdd(stats + 1, (unsigned long long)toks);
  }
}

// step 4: per query the pruning threshold T (at least n_dec candidates have lb >= T; found with two levels of
// 2048 value buckets, so outliers only cost resolution) and the list of unresolved candidates with ub >= T.
// warp 0: bucket t with count(bucket > t) < need <= count(bucket >= t) over a 2048-bin histogram
__device__ __forceinline__ void k3_find_bucket(const int* hist, int need, int lane, int* s_t, int* s_need) {
  int mine = 0;
  for (int k = 0; k < 64; ++k) mine += hist[lane * 64 + k];
  int above = 0;
  for (int l = 31; l >= 0; --l) {
    const int c = __shfl_sync(0xffffffffu, mine, l);
    if (l > lane) above += c;
  }
  if (above < need && above + mine >= need) {
    int cum = above, d = lane * 64 + 63;
    for (; d > lane * 64; --d) {
      const int h = hist[d];
      if (cum + h >= need) break;
      cum += h;
    }
    *s_t = d;
    *s_need = need - cum;  // still to take from bucket d
  }
}

__global__ void __launch_bounds__(1024)
k3_refine_list_kernel(const float* __restrict__ ub, const float* __restrict__ lb, int cand_cap,
                      const int32_t* __restrict__ n_cand, int n_dec, int refine_all, int32_t* __restrict__ list,
                      int32_t* __restrict__ n_list, float* __restrict__ thresh) {
  __shared__ int hist[2048];
  __shared__ float s_red[64];
  __shared__ int s_t, s_need, s_cnt;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int n = n_cand[b];
  const float* ubb = ub + int64_t(b) * cand_cap;
  const float* lbb = lb + int64_t(b) * cand_cap;
  int32_t* out = list + int64_t(b) * cand_cap;
  constexpr int FV = 8;
  float T = -INFINITY;
  if (!refine_all && n > n_dec) {  // n <= n_dec: nothing is pruned (search.rs:605 / :615), every score is needed
    float mn = INFINITY, mx = -INFINITY;
    for (int i0 = tid; i0 < n; i0 += 1024 * FV) {
      float v[FV];
#pragma unroll
      for (int u = 0; u < FV; ++u) v[u] = (i0 + u * 1024 < n) ? lbb[i0 + u * 1024] : NAN;  // fmin/fmax skip NaN
#pragma unroll
      for (int u = 0; u < FV; ++u) {
        mn = fminf(mn, v[u]);
        mx = fmaxf(mx, v[u]);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) {
      s_red[tid >> 5] = mn;
      s_red[32 + (tid >> 5)] = mx;
    }
    __syncthreads();
    mn = s_red[lane];
    mx = s_red[32 + lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    const float range = mx - mn;
    if (range > 0.f && range < 3.0e38f) {
      // level 1: 2048 buckets over [mn, mx]; bucket() is monotone in v
      const float scale1 = 2047.0f / range;
      auto bucket1 = [&](float v) { return min(2047, max(0, __float2int_rz((v - mn) * scale1))); };
      for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
      __syncthreads();
      for (int i0 = tid; i0 < n; i0 += 1024 * FV) {
        float v[FV];
#pragma unroll
        for (int u = 0; u < FV; ++u) v[u] = (i0 + u * 1024 < n) ? lbb[i0 + u * 1024] : 0.f;
#pragma unroll
        for (int u = 0; u < FV; ++u)
          if (i0 + u * 1024 < n && v[u] == v[u]) atomicAdd(&hist[bucket1(v[u])], 1);
      }
      if (tid == 0) s_t = -1;
      __syncthreads();
      if (tid < 32) k3_find_bucket(hist, n_dec, lane, &s_t, &s_need);
      __syncthreads();
      const int t1 = s_t, need1 = s_need;
      __syncthreads();
      if (t1 >= 0) {  // (t1 < 0 only if NaNs leave fewer than n_dec comparable values: T stays -inf)
        // level 2: 2048 buckets inside bucket t1
        const float lo1 = mn + float(t1) / scale1;
        const float scale2 = 2047.0f * scale1;
        auto bucket2 = [&](float v) { return min(2047, max(0, __float2int_rz((v - lo1) * scale2))); };
        for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
        __syncthreads();
        for (int i0 = tid; i0 < n; i0 += 1024 * FV) {
          float v[FV];
#pragma unroll
          for (int u = 0; u < FV; ++u) v[u] = (i0 + u * 1024 < n) ? lbb[i0 + u * 1024] : 0.f;
#pragma unroll
          for (int u = 0; u < FV; ++u)
            if (i0 + u * 1024 < n && v[u] == v[u] && bucket1(v[u]) == t1) atomicAdd(&hist[bucket2(v[u])], 1);
        }
        if (tid == 0) s_t = -1;
        __syncthreads();
        if (tid < 32) k3_find_bucket(hist, need1, lane, &s_t, &s_need);
        __syncthreads();
        const int t2 = s_t;
        // T = the smallest value of the selected upper set {b1 > t1} u {b1 == t1, b2 >= t2}: it holds >= n_dec values
        float tmin = INFINITY;
        if (t2 >= 0) {
          for (int i0 = tid; i0 < n; i0 += 1024 * FV) {
            float v[FV];
#pragma unroll
            for (int u = 0; u < FV; ++u) v[u] = (i0 + u * 1024 < n) ? lbb[i0 + u * 1024] : NAN;
#pragma unroll
            for (int u = 0; u < FV; ++u) {
              if (v[u] == v[u]) {
                const int b1 = bucket1(v[u]);
                if (b1 > t1 || (b1 == t1 && bucket2(v[u]) >= t2)) tmin = fminf(tmin, v[u]);
              }
            }
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, off));
        __syncthreads();
        if (lane == 0) s_red[tid >> 5] = tmin;
        __syncthreads();
        tmin = s_red[lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, off));
        if (t2 >= 0 && tmin < INFINITY) T = tmin;
      }
    } else if (range == 0.f) {
      T = mn;  // every lower bound is the same value: all n > n_dec candidates have lb >= mn
    }
  }
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  const int n_up = (n + 1023) / 1024 * 1024;
  for (int i = tid; i < n_up; i += 1024) {
    bool take = false;
    if (i < n) {
      const float u = ubb[i], l = lbb[i];
      take = (l < u) && (u >= T);
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&s_cnt, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) out[base + __popc(m & ((1u << lane) - 1u))] = i;
  }
  __syncthreads();
  if (tid == 0) {
    n_list[b] = s_cnt;
    thresh[b] = T;
  }
}

// ---------------------------------------------------------------------------------------
// K3b: top-n_dec by (approx desc, candidate index asc) -- candidate index order is doc id
// order, so this is the canonical rule "larger score, then smaller doc id".  Equivalent to
// the reference's topk(n_full) followed by topk(n_full/4) up to tie order.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t approx_key(float a, uint32_t i) {
  return (uint64_t(f32_key(a)) << 32) | uint64_t(0xffffffffu - i);
}

// Digits of the 64-bit key, most significant first: 11+11+10 bits cover the score, the rest
// only matters when scores tie at the threshold.
__constant__ int K3B_LO[6] = {53, 42, 32, 21, 10, 0};
__constant__ int K3B_W[6] = {11, 11, 10, 11, 11, 10};
constexpr int K3B_VPT = 4;  // independent loads in flight per thread
constexpr int K3B_BCAP = 2048;  // capacity of the threshold bucket on the fast path

__global__ void __launch_bounds__(1024)
k3b_select_kernel(const float* __restrict__ approx, const int32_t* __restrict__ cand, int cand_cap,
                  const int32_t* __restrict__ n_cand, int n_dec, int Rp2, int32_t* __restrict__ rerank,
                  float* __restrict__ rerank_approx, int32_t* __restrict__ n_rerank) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ int hist[2048];
  __shared__ int s_need, s_hd, s_cnt;
  __shared__ uint64_t s_prefix, s_mask;
  __shared__ uint64_t bkeys[K3B_BCAP];  // fast path: keys of the threshold bucket
  __shared__ float s_red[64];
  __shared__ int s_cnt2, s_fast;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int n = n_cand[b];
  const float* ab = approx + int64_t(b) * cand_cap;
  const int32_t* cb = cand + int64_t(b) * cand_cap;
  int32_t* rr = rerank + int64_t(b) * n_dec;
  float* ra = rerank_approx + int64_t(b) * n_dec;
  if (n <= n_dec) {  // search.rs:605 / :615 conditions false: nothing is pruned
    for (int i = tid; i < n; i += 1024) {
      rr[i] = cb[i];
      ra[i] = ab[i];
    }
    if (tid == 0) n_rerank[b] = n;
    return;
  }
  // ---- fast path: 2048 buckets over the VALUE range [min, max] of this query's scores ----
  // The radix passes below start from the top bits of the float key, where the scores of one query share
  // sign, exponent and the leading mantissa bits: a handful of hot bins, so every element pays a ballot +
  // match_any + contended shared atomic, three passes long (0.49 ms on cfg-3).  A linear bucketisation of the
  // actual value range spreads the scores, one histogram pass isolates the threshold bucket, and only that
  // bucket (typically n / 2048 elements) is ordered by the exact 64-bit key.  bucket(v) is monotone in v, so
  // every element of a higher bucket is strictly larger: the selection is exactly the same.
  {
    constexpr int FV = 8;  // independent loads in flight per thread
    float mn = INFINITY, mx = -INFINITY;
    for (int i0 = tid; i0 < n; i0 += 1024 * FV) {
      float v[FV];
#pragma unroll
      for (int u = 0; u < FV; ++u) v[u] = (i0 + u * 1024 < n) ? ab[i0 + u * 1024] : NAN;  // fmin/fmax skip NaN
#pragma unroll
      for (int u = 0; u < FV; ++u) {
        mn = fminf(mn, v[u]);
        mx = fmaxf(mx, v[u]);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) {
      s_red[tid >> 5] = mn;
      s_red[32 + (tid >> 5)] = mx;
    }
    for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
    if (tid == 0) {
      s_cnt = 0;
      s_cnt2 = 0;
      s_fast = 0;
    }
    __syncthreads();
    mn = s_red[lane];
    mx = s_red[32 + lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    const float range = mx - mn;
    // finite, non-degenerate range (NaN / inf scores or all-equal scores take the radix path)
    const bool usable = range > 0.f && range < 3.0e38f;
    const float scale = usable ? 2047.0f / range : 0.f;
    auto bucket = [&](float v) { return min(2047, max(0, __float2int_rz((v - mn) * scale))); };
    if (usable) {
      for (int i0 = tid; i0 < n; i0 += 1024 * FV) {
        float v[FV];
#pragma unroll
        for (int u = 0; u < FV; ++u) v[u] = (i0 + u * 1024 < n) ? ab[i0 + u * 1024] : 0.f;
#pragma unroll
        for (int u = 0; u < FV; ++u)
          if (i0 + u * 1024 < n) atomicAdd(&hist[bucket(v[u])], 1);
      }
      __syncthreads();
      if (tid < 32) {
        // warp 0: bucket t with  count(bucket > t) < n_dec <= count(bucket >= t)
        int mine = 0;
        for (int k = 0; k < 64; ++k) mine += hist[lane * 64 + k];
        int above = 0;
        for (int l = 31; l >= 0; --l) {
          const int c = __shfl_sync(0xffffffffu, mine, l);
          if (l > lane) above += c;
        }
        if (above < n_dec && above + mine >= n_dec) {
          int cum = above, d = lane * 64 + 63;
          for (; d > lane * 64; --d) {
            const int h = hist[d];
            if (cum + h >= n_dec) break;
            cum += h;
          }
          s_need = n_dec - cum;  // still to take from bucket d
          s_hd = d;
          s_fast = hist[d] <= K3B_BCAP ? 1 : 0;
        }
      }
      __syncthreads();
      if (s_fast) {
        const int t = s_hd;
        for (int i0 = tid; i0 < n; i0 += 1024 * FV) {
          float v[FV];
#pragma unroll
          for (int u = 0; u < FV; ++u) v[u] = (i0 + u * 1024 < n) ? ab[i0 + u * 1024] : 0.f;
#pragma unroll
          for (int u = 0; u < FV; ++u) {
            const int i = i0 + u * 1024;
            if (i < n) {
              const int bk = bucket(v[u]);
              if (bk > t) {
                keys[atomicAdd(&s_cnt, 1)] = approx_key(v[u], uint32_t(i));    // fewer than n_dec of these
              } else if (bk == t) {
                bkeys[atomicAdd(&s_cnt2, 1)] = approx_key(v[u], uint32_t(i));  // at most K3B_BCAP of these
              }
            }
          }
        }
        __syncthreads();
        const int c2 = s_cnt2, need2 = s_need;
        for (int i = c2 + tid; i < K3B_BCAP; i += 1024) bkeys[i] = 0ull;
        __syncthreads();
        for (int k = 2; k <= K3B_BCAP; k <<= 1) {  // bitonic sort, descending
          for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < K3B_BCAP; i += 1024) {
              const int ixj = i ^ j;
              if (ixj > i) {
                const bool up = (i & k) == 0;
                const uint64_t x = bkeys[i], y = bkeys[ixj];
                if ((x < y) == up) {
                  bkeys[i] = y;
                  bkeys[ixj] = x;
                }
              }
            }
            __syncthreads();
          }
        }
        const int c1 = s_cnt;
        for (int i = tid; i < need2; i += 1024) keys[c1 + i] = bkeys[i];
        __syncthreads();
        if (tid == 0) s_cnt = c1 + need2;  // == n_dec
        __syncthreads();
      }
    }
  }
  const bool fast_done = s_fast != 0;
  if (!fast_done) {
  if (tid == 0) {
    s_need = n_dec;
    s_prefix = 0;
    s_mask = 0;
  }
  }
  const int stride = 1024 * K3B_VPT;
  const int n_up = (n + stride - 1) / stride * stride;
  if (!fast_done) {
  for (int pass = 0; pass < 6; ++pass) {
    const int lo = K3B_LO[pass], width = K3B_W[pass];
    const uint64_t dmask = (1ull << width) - 1ull;
    for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
    __syncthreads();
    const uint64_t prefix = s_prefix, mask = s_mask;
    for (int i0 = tid; i0 < n_up; i0 += stride) {
      float v[K3B_VPT];
#pragma unroll
      for (int u = 0; u < K3B_VPT; ++u) {
        const int i = i0 + u * 1024;
        v[u] = (i < n) ? ab[i] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < K3B_VPT; ++u) {
        const int i = i0 + u * 1024;
        const bool valid = i < n;
        const uint64_t key = valid ? approx_key(v[u], uint32_t(i)) : 0ull;
        const bool in = valid && ((key & mask) == prefix);
        const unsigned act = __ballot_sync(0xffffffffu, in);
        if (in) {
          const int bin = int((key >> lo) & dmask);
          const unsigned peers = __match_any_sync(act, bin);
          if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
        }
      }
    }
    __syncthreads();
    if (tid < 32) {
      // warp 0: find the digit d with  count(digit > d) < need <= count(digit >= d)
      const int nbins = 1 << width;
      const int per = nbins / 32;  // bins per lane, lane 31 owns the top bins
      int mine = 0;
      for (int k = 0; k < per; ++k) mine += hist[lane * per + k];
      // suffix sums over lanes (lanes above me)
      int above = 0;
      for (int l = 31; l >= 0; --l) {
        const int c = __shfl_sync(0xffffffffu, mine, l);
        if (l > lane) above += c;
      }
      const int need = s_need;
      const bool here = (above < need) && (above + mine >= need);
      if (here) {
        int cum = above, d = lane * per + per - 1;
        for (; d > lane * per; --d) {
          const int h = hist[d];
          if (cum + h >= need) break;
          cum += h;
        }
        s_need = need - cum;
        s_hd = hist[d];
        s_prefix = prefix | (uint64_t(d) << lo);
        s_mask = mask | (dmask << lo);
      }
    }
    __syncthreads();
    if (s_hd == s_need) break;  // the whole bucket is selected
  }
  const uint64_t T = s_prefix;  // unprocessed low bits are zero
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int i0 = tid; i0 < n_up; i0 += stride) {
    float v[K3B_VPT];
#pragma unroll
    for (int u = 0; u < K3B_VPT; ++u) {
      const int i = i0 + u * 1024;
      v[u] = (i < n) ? ab[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < K3B_VPT; ++u) {
      const int i = i0 + u * 1024;
      if (i < n) {
        const uint64_t key = approx_key(v[u], uint32_t(i));
        if (key >= T) {
          const int pos = atomicAdd(&s_cnt, 1);
          if (pos < Rp2) keys[pos] = key;
        }
      }
    }
  }
  }  // radix path
  __syncthreads();
  const int cnt = min(s_cnt, Rp2);
  for (int i = cnt + tid; i < Rp2; i += 1024) keys[i] = 0ull;
  __syncthreads();
  for (int k = 2; k <= Rp2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < Rp2; i += 1024) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const uint64_t x = keys[i], y = keys[ixj];
          if ((x < y) == up) {
            keys[i] = y;
            keys[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < n_dec; r += 1024) {
    const uint32_t idx = 0xffffffffu - uint32_t(keys[r]);
    rr[r] = cb[idx];
    ra[r] = ab[idx];
  }
  if (tid == 0) n_rerank[b] = n_dec;
}

// exact scoring of `list` (NULL: every candidate) into approx
template <int LPR>
int launch_k3_exact(const fpb_index* ix, const Ws& ws, const int32_t* list, const int32_t* n_list, int32_t* work,
                    cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  const int blocks = ix->sm_count * 8;
  if constexpr (LPR <= 4) {
    // shuffle-free up to Qp = 32 (the vector code loads need a 16-byte aligned code array; any torch allocation
    // is); at Qp = 64 the shuffle kernel is faster (cfg-5: 10.55 vs 11.31 ms)
    if ((reinterpret_cast<uintptr_t>(ix->doc_codes) & 15u) == 0) {
      k3_approx_nsh_kernel<LPR, 2, 6><<<blocks, K3_THREADS, 0, st>>>(
          ws.S(), ix->K, L.Q, ix->doc_offsets, ix->doc_codes, ix->E, ws.cand(), L.cand_cap, ws.n_cand(), list, n_list,
          work, L.B, ws.approx(), ws.stats());
      FPB_LAUNCH_CHECK("k3_approx_nsh");
      return FPB_OK;
    }
  }
  k3_approx_kernel<LPR, 2, 6><<<blocks, K3_THREADS, 0, st>>>(ws.S(), ix->K, L.Q, ix->doc_offsets, ix->doc_codes,
                                                             ws.cand(), L.cand_cap, ws.n_cand(), list, n_list, work,
                                                             L.B, ws.approx(), ws.stats());
  FPB_LAUNCH_CHECK("k3_approx");
  return FPB_OK;
}

// LAMBDA of k3_tau_kernel: the expected number of tokens per candidate and column at or above tau.  A document is
// resolved when all of its Q columns are, so the useful level grows with log Q: LAMBDA = ln(Q) - 1.45 (2.0 at
// Q = 32, 2.7 at Q = 64; measured optimum on cfg-3, within 10 % of it on cfg-5).  FPB_K3_LAMBDA overrides it
// (tuning only: every value gives the same results).
float k3_tau_lambda(int Q) {
  static const float pinned = [] {
    const char* e = getenv("FPB_K3_LAMBDA");
    return e ? float(atof(e)) : 0.f;
  }();
  const float x = pinned > 0.f ? pinned : logf(float(Q < 2 ? 2 : Q)) - 1.45f;
  return x < 0.5f ? 0.5f : (x > 64.f ? 64.f : x);
}

template <int LPR>
int launch_k3_t(const fpb_index* ix, const Ws& ws, int flags, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  // one-pass scoring: asked for, or the K-bit map of a query does not fit next to a second CTA's, or (nothing asked
  // for) the job is too small to repay the fixed cost of the two passes -- results are identical in every case
  const bool forced = (flags & (FPB_FLAG_APPROX_TWO_PASS | FPB_FLAG_APPROX_EXACT_ALL)) != 0;
  const bool small_job = int64_t(L.B) * ix->E < K3_TWO_PASS_MIN_TOKENS;
  if ((flags & FPB_FLAG_APPROX_DIRECT) || size_t(L.hb_words) * 4 > 96 * 1024 || (!forced && small_job)) {
    FPB_CUDA_CHECK(cudaMemsetAsync(ws.n_refine(), 0, size_t(L.B) * 4, st));  // nothing was re-scored
    k3_prefix_kernel<<<1, 32, 0, st>>>(ws.n_cand(), L.B, K3_DOCS_PER_CHUNK, ws.work());
    FPB_LAUNCH_CHECK("k3_prefix");
    return launch_k3_exact<LPR>(ix, ws, nullptr, nullptr, ws.work(), st);
  }
  k3_tau_kernel<LPR><<<L.B, K3_TAU_THREADS, 0, st>>>(ws.S(), ix->K, L.Q, ix->doc_offsets, ix->doc_codes, ws.cand(),
                                                     L.cand_cap, ws.n_cand(), ws.tmax(), L.n_tiles, k3_tau_lambda(L.Q),
                                                     ws.tau());
  FPB_LAUNCH_CHECK("k3_tau");
  {
    dim3 grid(unsigned((ix->K + 255) / 256), unsigned(L.B));
    k3_hibits_kernel<LPR><<<grid, 256, 0, st>>>(ws.S(), ix->K, ws.tau(), ws.hibits(), L.hb_words);
    FPB_LAUNCH_CHECK("k3_hibits");
  }
  k3_prefix_kernel<<<1, 32, 0, st>>>(ws.n_cand(), L.B, K3A_DOCS_PER_CHUNK, ws.work());
  FPB_LAUNCH_CHECK("k3_prefix");
  {
    // resident CTAs per SM: limited by the bitmap (228 KB of shared memory per SM, 1 KB reserved per CTA)
    // shape of the walk: 6 32-token windows per group (all their code loads in flight at once, the next group's
    // prefetched), 4 gathers per lane and batch, 4 CTAs per SM at 64 registers, paired epilogues.  Measured
    // alternatives (profiles/r02_summary.md): 4 / 8 / 11 / 12 windows, 2 / 8 gathers, 5-6 CTAs at 48 / 40 registers,
    // one document per epilogue -- all within 3 % or slower.
    constexpr int TPI = 32 / LPR;
    constexpr int W = 6, U = 4, MINB = 4;
    const bool fullq = L.Q == L.Qp;
    auto kern = fullq ? k3_bound_kernel<LPR, MINB, W, U, true, true> : k3_bound_kernel<LPR, MINB, W, U, false, true>;
    const int minb = MINB, wq = K3_WQ_FOR(W, U * TPI);
    const size_t smem = size_t(L.hb_words + 4) * 4 + size_t(K3_THREADS / 32 + 1) * wq * 4;  // + alignment slack of the rings
    FPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    // resident CTAs per SM: limited by the bitmap (228 KB of shared memory per SM, 1 KB reserved per CTA)
    int per_sm = int((227 * 1024) / (smem + 1024 + 2048));
    per_sm = per_sm < 1 ? 1 : (per_sm > minb ? minb : per_sm);
    kern<<<ix->sm_count * per_sm, K3_THREADS, smem, st>>>(ws.S(), ix->K, L.Q, ix->doc_offsets, ix->doc_codes, ws.cand(),
                                                          L.cand_cap, ws.n_cand(), ws.work(), L.B, ws.tau(),
                                                          ws.hibits(), L.hb_words, ws.approx(), ws.lb(), ws.stats());
    FPB_LAUNCH_CHECK("k3_bound");
  }
  k3_refine_list_kernel<<<L.B, 1024, 0, st>>>(ws.approx(), ws.lb(), L.cand_cap, ws.n_cand(), L.R,
                                              (flags & FPB_FLAG_APPROX_EXACT_ALL) ? 1 : 0, ws.refine(), ws.n_refine(),
                                              ws.thresh());
  FPB_LAUNCH_CHECK("k3_refine_list");
  k3_prefix_kernel<<<1, 32, 0, st>>>(ws.n_refine(), L.B, K3_DOCS_PER_CHUNK, ws.work2());
  FPB_LAUNCH_CHECK("k3_prefix");
  return launch_k3_exact<LPR>(ix, ws, ws.refine(), ws.n_refine(), ws.work2(), st);
}

}  // namespace

int launch_approx(const fpb_index* ix, const Ws& ws, int flags, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  switch (L.Qp / 8) {
    case 2: return launch_k3_t<2>(ix, ws, flags, st);
    case 4: return launch_k3_t<4>(ix, ws, flags, st);
    case 8: return launch_k3_t<8>(ix, ws, flags, st);
    case 16: return launch_k3_t<16>(ix, ws, flags, st);
    case 32: return launch_k3_t<32>(ix, ws, flags, st);
    default:
      fpb_set_error("approx scoring: unsupported padded query length %d", L.Qp);
      return FPB_ERR_UNSUPPORTED;
  }
}

int launch_select(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  (void)ix;
  const fpb_layout& L = *ws.L;
  const int Rp2 = fpb_next_pow2(L.R);
  // dynamic keys[] (8 B x Rp2, 32 KB at the maximum R = 4096) on top of 25 KB of static shared memory: opt in
  // (per device: cudaFuncSetAttribute applies to the current device only, and the call is cheap)
  FPB_CUDA_CHECK(cudaFuncSetAttribute(k3b_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Rp2 * 8));
  k3b_select_kernel<<<L.B, 1024, size_t(Rp2) * 8, st>>>(ws.approx(), ws.cand(), L.cand_cap, ws.n_cand(),
                                                       L.R, Rp2, ws.rerank(), ws.rerank_approx(),
                                                       ws.n_rerank());
  FPB_LAUNCH_CHECK("k3b_select");
  return FPB_OK;
}
