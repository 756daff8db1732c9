// K3  : approximate (centroid-only) document scores                       (search.rs:554-592)
//         approx[d] = sum_{q<Q}^{fp32}  max_{t<len(d)}  S[b][code[d,t]][q]     (fp16 max)
// K3b : pruning to the n_full_scores/4 best candidates                     (search.rs:602-619)
//
// The reference gathers S rows into a [tokens, Q] tensor, pads it to [2000, maxlen, Q],
// masks, maxes and sums, 128 times per query with two host syncs each.  Here one warp walks
// one candidate document: 32 codes are read with one coalesced load, each S row (Qp fp16 =
// LPR x 16 B) is fetched by LPR adjacent lanes so a warp-wide load touches 32/LPR distinct
// rows (one L1 wavefront per row instead of four), and the running maxima stay in registers.
// The stage is bound by the L1 data pipe (one wavefront per gathered row; 97 % busy at Qp = 32) or, at
// Qp = 64, by L2 bandwidth; the codes stream from HBM once per (query, candidate).  Two kernels: the
// shuffle-free one below (default up to Qp = 32) and this one (Qp >= 64, unaligned code arrays).
#include <stdlib.h>

#include "kernels.h"

namespace {

constexpr int K3_THREADS = 256;
constexpr int K3_DOCS_PER_CHUNK = 64;

__global__ void k3_prefix_kernel(const int32_t* __restrict__ n_cand, int B, int32_t* __restrict__ work) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      work[b] = acc;
      acc += (n_cand[b] + K3_DOCS_PER_CHUNK - 1) / K3_DOCS_PER_CHUNK;
    }
    work[B] = acc;
    work[B + 1] = 0;  // dynamic work counter
  }
}

template <int LPR, int UNROLL, int MINB>
__global__ void __launch_bounds__(K3_THREADS, MINB)
k3_approx_kernel(const __half* __restrict__ S, int64_t K, int Q, const int64_t* __restrict__ doc_offsets,
                 const int32_t* __restrict__ codes, const int32_t* __restrict__ cand, int cand_cap,
                 const int32_t* __restrict__ n_cand, int32_t* __restrict__ work, int B,
                 float* __restrict__ approx) {
  constexpr int QP = LPR * 8;
  constexpr int TPI = 32 / LPR;
  __shared__ int s_b, s_c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const __half2 sentinel = __float2half2_rn(FPB_PAD_SENTINEL);

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const int c = atomicAdd(&work[B + 1], 1);
      if (c >= work[B]) {
        s_b = -1;
      } else {
        int lo = 0, hi = B - 1;  // largest b with work[b] <= c
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (work[mid] <= c) lo = mid; else hi = mid - 1;
        }
        s_b = lo;
        s_c = c - work[lo];
      }
    }
    __syncthreads();
    const int b = s_b;
    if (b < 0) return;
    const int n = n_cand[b];
    const uint4* Sb = reinterpret_cast<const uint4*>(S + int64_t(b) * K * QP);
    const int32_t* cb = cand + int64_t(b) * cand_cap;
    float* ab = approx + int64_t(b) * cand_cap;

    for (int i = 0; i < K3_DOCS_PER_CHUNK / 8; ++i) {
      const int idx = s_c * K3_DOCS_PER_CHUNK + i * 8 + warp;
      if (idx >= n) break;
      const int d = cb[idx];
      const int64_t o0 = doc_offsets[d];
      const int len = int(doc_offsets[d + 1] - o0);
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
          for (int j = 0; j < LPR; ++j) {
            const int c = __shfl_sync(0xffffffffu, code[u], j * TPI + grp);
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
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1) {
        m0 = __hmax2(m0, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m0), off)));
        m1 = __hmax2(m1, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m1), off)));
        m2 = __hmax2(m2, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m2), off)));
        m3 = __hmax2(m3, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m3), off)));
      }
      // fp32 sum over the real query tokens (sum_dim_intlist(.., Kind::Float), search.rs:401)
      const int col0 = sub * 8;
      float s = 0.f;
      const float2 f0 = __half22float2(m0), f1 = __half22float2(m1), f2 = __half22float2(m2),
                   f3 = __half22float2(m3);
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
      if (lane == 0) ab[idx] = s;
    }
  }
}

// Shuffle-free variant.  The ncu capture of the kernel above (profiles/r01b_top3_raw.csv) shows the
// L1TEX data pipe 96.5 % busy: 32.1 M of its 38.6 M wavefronts per SM are the row gathers (one per
// row, the floor of this formulation) and 6.5 M are the __shfl_sync that hand each code to the LPR
// lanes fetching its row -- shuffles run through the same pipe.  Here the LPR lanes of a group
// load "their" LPR consecutive codes themselves with one vector load (the lanes of a group read
// the same 4*LPR bytes: a broadcast, one wavefront for the warp): the document is walked in
// 32-token windows aligned to absolute multiples of 32 tokens, so the vector loads are aligned
// whatever the document offset; tokens outside [o0, o0+len) are masked.  max() is order-free,
// so the values are bit-identical to the shuffle kernel.
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
                     int cand_cap, const int32_t* __restrict__ n_cand, int32_t* __restrict__ work, int B,
                     float* __restrict__ approx) {
  constexpr int QP = LPR * 8;
  __shared__ int s_b, s_c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const __half2 sentinel = __float2half2_rn(FPB_PAD_SENTINEL);

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const int c = atomicAdd(&work[B + 1], 1);
      if (c >= work[B]) {
        s_b = -1;
      } else {
        int lo = 0, hi = B - 1;  // largest b with work[b] <= c
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (work[mid] <= c) lo = mid; else hi = mid - 1;
        }
        s_b = lo;
        s_c = c - work[lo];
      }
    }
    __syncthreads();
    const int b = s_b;
    if (b < 0) return;
    const int n = n_cand[b];
    const uint4* Sb = reinterpret_cast<const uint4*>(S + int64_t(b) * K * QP) + sub;
    const int32_t* cb = cand + int64_t(b) * cand_cap;
    float* ab = approx + int64_t(b) * cand_cap;

    for (int i = 0; i < K3_DOCS_PER_CHUNK / 8; ++i) {
      const int idx = s_c * K3_DOCS_PER_CHUNK + i * 8 + warp;
      if (idx >= n) break;
      const int d = cb[idx];
      const int64_t o0 = doc_offsets[d];
      const int len = int(doc_offsets[d + 1] - o0);
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
            for (int j = 0; j < LPR; ++j) c[u][j] = (fo + j < readable) ? __ldg(cw + fo + j) : 0;
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
          for (int j = 0; j < LPR; ++j) {
            const int f = f0 + u * 32 + grp * LPR + j;
            if (unsigned(f - lo) < unsigned(len)) {
              const uint4 v = __ldg(Sb + int64_t(c[u][j]) * LPR);
              m0 = __hmax2(m0, u32_as_half2(v.x));
              m1 = __hmax2(m1, u32_as_half2(v.y));
              m2 = __hmax2(m2, u32_as_half2(v.z));
              m3 = __hmax2(m3, u32_as_half2(v.w));
            }
          }
        }
      }
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1) {
        m0 = __hmax2(m0, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m0), off)));
        m1 = __hmax2(m1, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m1), off)));
        m2 = __hmax2(m2, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m2), off)));
        m3 = __hmax2(m3, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m3), off)));
      }
      const int col0 = sub * 8;
      float s = 0.f;
      const float2 f0 = __half22float2(m0), f1 = __half22float2(m1), f2 = __half22float2(m2),
                   f3 = __half22float2(m3);
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
      if (lane == 0) ab[idx] = s;
    }
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

template <int LPR>
int launch_k3_t(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  const int blocks = ix->sm_count * 8;
  // FPB_K3_VARIANT selects a tuning variant (A/B measurements; all compute identical values)
  static const int variant = getenv("FPB_K3_VARIANT") ? atoi(getenv("FPB_K3_VARIANT")) : 0;
#define K3_LAUNCH(U, M)                                                                                     \
  k3_approx_kernel<LPR, U, M><<<blocks, K3_THREADS, 0, st>>>(ws.S(), ix->K, L.Q, ix->doc_offsets,            \
                                                             ix->doc_codes, ws.cand(), L.cand_cap, ws.n_cand(), \
                                                             ws.work(), L.B, ws.approx())
  // measured on cfg-3 (ms): U1/M6 26.2, U2/M6 20.4, U1/M8 23.5, U2/M8 20.8, U4/M4 22.4
#define K3_LAUNCH_NSH(U, M)                                                                                  \
  k3_approx_nsh_kernel<LPR, U, M><<<blocks, K3_THREADS, 0, st>>>(ws.S(), ix->K, L.Q, ix->doc_offsets,         \
                                                                 ix->doc_codes, ix->E, ws.cand(), L.cand_cap,  \
                                                                 ws.n_cand(), ws.work(), L.B, ws.approx())
  if constexpr (LPR <= 8) {
    // the vector code loads need a 16-byte aligned code array (any torch allocation is)
    const bool aligned = (reinterpret_cast<uintptr_t>(ix->doc_codes) & 15u) == 0;
    // default: shuffle-free up to Qp = 32; at Qp = 64 (8 codes per lane) the shuffle kernel is faster
    // (cfg-5: 10.55 vs 11.31 ms)
    if (aligned && ((variant == 0 && LPR <= 4) || variant >= 10)) {
      switch (variant) {
        case 0: K3_LAUNCH_NSH(2, 6); break;  // default: 20.0-20.2 ms on cfg-3 (shuffle kernel: 20.9-21.0 in the same session)
        case 10: K3_LAUNCH_NSH(1, 6); break;
        case 11: K3_LAUNCH_NSH(2, 6); break;
        case 12: K3_LAUNCH_NSH(2, 5); break;
        case 13: K3_LAUNCH_NSH(1, 8); break;
        case 14: K3_LAUNCH_NSH(2, 7); break;
        case 15: K3_LAUNCH_NSH(2, 8); break;
        case 16: K3_LAUNCH_NSH(3, 6); break;
        case 17: K3_LAUNCH_NSH(4, 6); break;
        default: K3_LAUNCH_NSH(3, 5); break;
      }
      FPB_LAUNCH_CHECK("k3_approx_nsh");
      return FPB_OK;
    }
  }
  switch (variant) {  // shuffle kernel: Qp > 64, unaligned code arrays, or FPB_K3_VARIANT=1..5
    case 1: K3_LAUNCH(1, 6); break;
    case 2: K3_LAUNCH(2, 7); break;
    case 3: K3_LAUNCH(3, 5); break;
    case 4: K3_LAUNCH(3, 6); break;
    default: K3_LAUNCH(2, 6); break;  // also FPB_K3_VARIANT=5
  }
#undef K3_LAUNCH
#undef K3_LAUNCH_NSH
  FPB_LAUNCH_CHECK("k3_approx");
  return FPB_OK;
}

}  // namespace

int launch_approx(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  k3_prefix_kernel<<<1, 32, 0, st>>>(ws.n_cand(), L.B, ws.work());
  FPB_LAUNCH_CHECK("k3_prefix");
  switch (L.Qp / 8) {
    case 2: return launch_k3_t<2>(ix, ws, st);
    case 4: return launch_k3_t<4>(ix, ws, st);
    case 8: return launch_k3_t<8>(ix, ws, st);
    case 16: return launch_k3_t<16>(ix, ws, st);
    case 32: return launch_k3_t<32>(ix, ws, st);
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
  static int attr_bytes = 0;
  if (Rp2 * 8 > attr_bytes) {
    FPB_CUDA_CHECK(cudaFuncSetAttribute(k3b_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Rp2 * 8));
    attr_bytes = Rp2 * 8;
  }
  k3b_select_kernel<<<L.B, 1024, size_t(Rp2) * 8, st>>>(ws.approx(), ws.cand(), L.cand_cap, ws.n_cand(),
                                                       L.R, Rp2, ws.rerank(), ws.rerank_approx(),
                                                       ws.n_rerank());
  FPB_LAUNCH_CHECK("k3b_select");
  return FPB_OK;
}
