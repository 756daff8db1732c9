// K1  : centroid scores  S[b][k][q] = fp16( sum_d C[k][d] * q_b[q][d] )     (search.rs:491)
// K1b : per query token, the n_ivf_probe best centroids                      (search.rs:520-528)
//
// v1 data path: legacy tensor-core MMA (mma.sync m16n8k16, fp16 in / fp32 accumulate), one
// rounding to fp16 at the end exactly like ATen's half matmul.  Every CTA keeps a tile of 128
// centroid rows resident (fragments in registers) and streams all query tokens of the batch
// past it, so the centroid table is read from HBM once per batch instead of once per query
// as the reference does.  The epilogue writes S in the [b][k][q] layout the approximate stage
// gathers from (one contiguous Qp*2-byte row per centroid and query) and, per 128-row tile,
// the column maxima that let K1b find the top-n cells by touching ~n tiles instead of all K.
#include <stdlib.h>

#include "kernels.h"

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_ROWS = 128;

template <int D, int QC>
struct K1Smem {
  static constexpr int LDS = D + 8;   // +16 B pad: conflict-free ldmatrix
  static constexpr int STG = QC + 8;
  static constexpr int bytes = (K1_ROWS * LDS + QC * LDS + 8 * 16 * STG + 8 * QC) * 2;
};

template <int D, int QC>
__global__ void __launch_bounds__(K1_THREADS)
k1_centroid_scores_kernel(const __half* __restrict__ C, int K, const __half* __restrict__ Qpad,
                          int B, int Qp, __half* __restrict__ S, __half* __restrict__ tmax,
                          int n_tiles, int b_per_cta) {
  constexpr int LDS = K1Smem<D, QC>::LDS;
  constexpr int STG = K1Smem<D, QC>::STG;
  constexpr int NT = QC / 8;
  constexpr int KS = D / 16;
  static_assert(NT % 2 == 0, "QC must be a multiple of 16");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* Cs = reinterpret_cast<__half*>(smem_raw);
  __half* Qs = Cs + K1_ROWS * LDS;
  __half* stage = Qs + QC * LDS;
  __half* cmax = stage + 8 * 16 * STG;

  const int tile = blockIdx.x;
  const int row0 = tile * K1_ROWS;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < K1_ROWS * (D / 8); i += K1_THREADS) {
    const int r = i / (D / 8), c8 = i % (D / 8);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < K) v = *reinterpret_cast<const uint4*>(C + int64_t(row0 + r) * D + c8 * 8);
    *reinterpret_cast<uint4*>(Cs + r * LDS + c8 * 8) = v;
  }
  __syncthreads();

  uint32_t a[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint32_t addr = smem_u32(Cs + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
    ldmatrix_x4(a[ks][0], a[ks][1], a[ks][2], a[ks][3], addr);
  }

  const int g = lane >> 2, t = lane & 3;
  const bool valid0 = (row0 + warp * 16 + g) < K;
  const bool valid1 = (row0 + warp * 16 + g + 8) < K;
  const __half2 ninf2 = __half2half2(__ushort_as_half(0xFC00));
  __half* st = stage + warp * 16 * STG;

  const int b_begin = blockIdx.y * b_per_cta;
  const int b_end = min(B, b_begin + b_per_cta);
  for (int b = b_begin; b < b_end; ++b) {
    for (int qc0 = 0; qc0 < Qp; qc0 += QC) {
      __syncthreads();  // previous round done with Qs / cmax
      for (int i = tid; i < QC * (D / 8); i += K1_THREADS) {
        const int n = i / (D / 8), c8 = i % (D / 8);
        *reinterpret_cast<uint4*>(Qs + n * LDS + c8 * 8) =
            *reinterpret_cast<const uint4*>(Qpad + (int64_t(b) * Qp + qc0 + n) * D + c8 * 8);
      }
      __syncthreads();

      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int nt = 0; nt < NT; nt += 2) {
          const int mat = lane >> 3;
          const int n = (nt + (mat >> 1)) * 8 + (lane & 7);
          const int k = ks * 16 + (mat & 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(b0, b1, b2, b3, smem_u32(Qs + n * LDS + k));
          mma_16816(acc[nt], a[ks], b0, b1);
          mma_16816(acc[nt + 1], a[ks], b2, b3);
        }
      }

#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint32_t p0 = pack_half2_rn(acc[nt][0], acc[nt][1]);
        const uint32_t p1 = pack_half2_rn(acc[nt][2], acc[nt][3]);
        *reinterpret_cast<uint32_t*>(st + g * STG + nt * 8 + 2 * t) = p0;
        *reinterpret_cast<uint32_t*>(st + (g + 8) * STG + nt * 8 + 2 * t) = p1;
        __half2 m = __hmax2(valid0 ? u32_as_half2(p0) : ninf2, valid1 ? u32_as_half2(p1) : ninf2);
        m = __hmax2(m, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m), 4)));
        m = __hmax2(m, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m), 8)));
        m = __hmax2(m, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m), 16)));
        if (g == 0) *reinterpret_cast<uint32_t*>(cmax + warp * QC + nt * 8 + 2 * t) = half2_as_u32(m);
      }
      __syncwarp();
      // the warp's 16 x QC block: rows are contiguous in S when QC == Qp
      for (int i = lane; i < 16 * (QC / 8); i += 32) {
        const int r = i / (QC / 8), c8 = i % (QC / 8);
        const int row = row0 + warp * 16 + r;
        if (row < K)
          *reinterpret_cast<uint4*>(S + (int64_t(b) * K + row) * Qp + qc0 + c8 * 8) =
              *reinterpret_cast<const uint4*>(st + r * STG + c8 * 8);
      }
      __syncthreads();
      if (tid < QC) {
        __half m = cmax[tid];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = __hmax(m, cmax[w * QC + tid]);
        tmax[(int64_t(b) * Qp + qc0 + tid) * n_tiles + tile] = m;
      }
    }
  }
}

// ---- warp-held sorted top-n list (lane i holds the i-th best key; 0 = empty) -----------
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, uint32_t(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, uint32_t(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v) {
  uint32_t lo = __shfl_up_sync(0xffffffffu, uint32_t(v), 1);
  uint32_t hi = __shfl_up_sync(0xffffffffu, uint32_t(v >> 32), 1);
  return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ void topn_insert(uint64_t& mine, int n, uint64_t x, int lane) {
  const unsigned gt = __ballot_sync(0xffffffffu, mine > x);
  const int pos = __popc(gt);
  const uint64_t up = shfl_up64(mine);
  if (pos < n) {
    if (lane == pos) mine = x;
    else if (lane > pos) mine = up;
    if (lane >= n) mine = 0;
  }
}
// Offer the keys held by the lanes (one each); keys equal to 0 are ignored.
__device__ __forceinline__ void topn_offer(uint64_t& mine, int n, uint64_t key, int lane) {
  uint64_t thr = shfl64(mine, n - 1);
  unsigned bits = __ballot_sync(0xffffffffu, key > thr);
  while (bits) {
    const int src = __ffs(bits) - 1;
    bits &= bits - 1;
    const uint64_t x = shfl64(key, src);
    if (x > thr) {  // uniform: x and thr are warp-uniform
      topn_insert(mine, n, x, lane);
      thr = shfl64(mine, n - 1);
    }
  }
}

// One warp per (query, query token).  Canonical tie rule: larger score first, then smaller
// centroid id (the reference's topk(sorted=false) leaves ties implementation-defined).
__global__ void __launch_bounds__(256)
k1b_probe_kernel(const __half* __restrict__ S, const __half* __restrict__ tmax, int K, int B, int Q,
                 int Qp, int n_tiles, int n_probe, int32_t* __restrict__ cells) {
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wg >= B * Q) return;
  const int b = wg / Q, q = wg % Q;
  const uint16_t* tm = reinterpret_cast<const uint16_t*>(tmax) + (int64_t(b) * Qp + q) * n_tiles;

  // pass 1: n-th largest tile maximum -> tau (any top-n element lives in a tile whose max >= tau)
  uint64_t mine = 0;
  for (int base = 0; base < n_tiles; base += 32) {
    const int tix = base + lane;
    uint64_t key = 0;
    if (tix < n_tiles) key = (uint64_t(f16_key(tm[tix])) << 32) | uint64_t(0xffffffffu - uint32_t(tix));
    topn_offer(mine, n_probe, key, lane);
  }
  const uint32_t tau = uint32_t(shfl64(mine, n_probe - 1) >> 32);  // 0 when fewer than n tiles

  // pass 2: exact top-n over the rows of the qualifying tiles
  mine = 0;
  const uint16_t* Sb = reinterpret_cast<const uint16_t*>(S) + int64_t(b) * K * Qp + q;
  for (int base = 0; base < n_tiles; base += 32) {
    const int tix = base + lane;
    bool qual = false;
    if (tix < n_tiles) qual = f16_key(tm[tix]) >= tau;
    unsigned tb = __ballot_sync(0xffffffffu, qual);
    while (tb) {
      const int tl = __ffs(tb) - 1;
      tb &= tb - 1;
      const int r0 = (base + tl) * K1_ROWS;
#pragma unroll
      for (int j = 0; j < K1_ROWS / 32; ++j) {
        const int row = r0 + j * 32 + lane;
        uint64_t key = 0;
        if (row < K) {
          const uint32_t k16 = f16_key(Sb[int64_t(row) * Qp]);
          if (k16 >= tau) key = (uint64_t(k16) << 32) | uint64_t(0xffffffffu - uint32_t(row));
        }
        topn_offer(mine, n_probe, key, lane);
      }
    }
  }
  if (lane < n_probe) {
    int32_t c = -1;
    if (mine != 0) c = int32_t(0xffffffffu - uint32_t(mine));
    cells[(int64_t(b) * Q + q) * n_probe + lane] = c;
  }
}

// Subset variant (search.rs:494-517): the top-n is taken over the centroids that occur in the
// subset's documents only (clist, ascending), n = min(n_ivf_probe, #such centroids).
__global__ void __launch_bounds__(256)
k1b_probe_subset_kernel(const __half* __restrict__ S, int K, int B, int Q, int Qp, int n_probe,
                        const int32_t* __restrict__ clist, const int32_t* __restrict__ n_clist,
                        int32_t* __restrict__ cells) {
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wg >= B * Q) return;
  const int b = wg / Q, q = wg % Q;
  const int nc = n_clist[b];
  const int n = min(n_probe, nc);
  const int32_t* cl = clist + int64_t(b) * K;
  const uint16_t* Sb = reinterpret_cast<const uint16_t*>(S) + int64_t(b) * K * Qp + q;
  uint64_t mine = 0;
  if (n > 0) {
    for (int base = 0; base < nc; base += 32) {
      const int i = base + lane;
      uint64_t key = 0;
      if (i < nc) {
        const int c = cl[i];
        key = (uint64_t(f16_key(Sb[int64_t(c) * Qp])) << 32) | uint64_t(0xffffffffu - uint32_t(c));
      }
      topn_offer(mine, n, key, lane);
    }
  }
  if (lane < n_probe) {
    int32_t c = -1;
    if (lane < n && mine != 0) c = int32_t(0xffffffffu - uint32_t(mine));
    cells[(int64_t(b) * Q + q) * n_probe + lane] = c;
  }
}

__global__ void pad_queries_kernel(const __half* __restrict__ q, __half* __restrict__ out, int B, int Q,
                                   int Qp, int D) {
  const int64_t n8 = int64_t(B) * Qp * (D / 8);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
    const int c8 = int(i % (D / 8));
    const int64_t row = i / (D / 8);
    const int qq = int(row % Qp);
    const int64_t b = row / Qp;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (qq < Q) v = *reinterpret_cast<const uint4*>(q + (b * Q + qq) * D + c8 * 8);
    *reinterpret_cast<uint4*>(out + row * D + c8 * 8) = v;
  }
}

template <int D, int QC>
int launch_k1_t(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  auto kern = k1_centroid_scores_kernel<D, QC>;
  constexpr int smem = K1Smem<D, QC>::bytes;
  // opt in on every launch: the attribute is per device and the call costs about a microsecond
  FPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  // split the batch over gridDim.y only when the centroid tiles alone cannot fill the chip
  int ysplit = 1;
  const int want = 2 * ix->sm_count;
  if (L.n_tiles < want) ysplit = min(L.B, (want + L.n_tiles - 1) / L.n_tiles);
  const int b_per_cta = (L.B + ysplit - 1) / ysplit;
  ysplit = (L.B + b_per_cta - 1) / b_per_cta;
  dim3 grid(L.n_tiles, ysplit);
  kern<<<grid, K1_THREADS, smem, st>>>(ix->centroids, int(ix->K), ws.queries(), L.B, L.Qp, ws.S(),
                                       ws.tmax(), L.n_tiles, b_per_cta);
  FPB_LAUNCH_CHECK("k1_centroid_scores");
  return FPB_OK;
}

template <int D>
int launch_k1_d(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  const int Qp = ws.L->Qp;
  if (Qp == 16) return launch_k1_t<D, 16>(ix, ws, st);
  if (Qp == 32) return launch_k1_t<D, 32>(ix, ws, st);
  return launch_k1_t<D, 64>(ix, ws, st);  // Qp in {64,128,256}: 64-column chunks
}

}  // namespace

int launch_pad_queries(const fpb_index* ix, const Ws& ws, const __half* d_queries, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  const int64_t n8 = int64_t(L.B) * L.Qp * (ix->dim / 8);
  const int blocks = int(((n8 + 255) / 256) < 4096 ? ((n8 + 255) / 256) : 4096);
  pad_queries_kernel<<<blocks, 256, 0, st>>>(d_queries, ws.queries(), L.B, L.Q, L.Qp, ix->dim);
  FPB_LAUNCH_CHECK("pad_queries");
  return FPB_OK;
}

int launch_centroid_scores(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  // FPB_K1=v1 pins the mma.sync kernel; default is the tcgen05 kernel where it applies.
  static const char* pin = getenv("FPB_K1");
  if (!pin || pin[1] != '1') {
    bool handled = false;
    const int rc = launch_centroid_scores_v2(ix, ws, st, &handled);
    if (rc != FPB_OK || handled) return rc;
  }
  switch (ix->dim) {
    case 64: return launch_k1_d<64>(ix, ws, st);
    case 128: return launch_k1_d<128>(ix, ws, st);
    default:
      fpb_set_error("centroid scoring: unsupported dim %d", ix->dim);
      return FPB_ERR_UNSUPPORTED;
  }
}

int launch_probe(const fpb_index* ix, const Ws& ws, bool subset, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  const int warps = L.B * L.Q;
  const int blocks = (warps + 7) / 8;
  if (subset) {
    k1b_probe_subset_kernel<<<blocks, 256, 0, st>>>(ws.S(), int(ix->K), L.B, L.Q, L.Qp, L.n_probe, ws.clist(),
                                                    ws.n_clist(), ws.cells());
    FPB_LAUNCH_CHECK("k1b_probe_subset");
    return FPB_OK;
  }
  k1b_probe_kernel<<<blocks, 256, 0, st>>>(ws.S(), ws.tmax(), int(ix->K), L.B, L.Q, L.Qp, L.n_tiles,
                                           L.n_probe, ws.cells());
  FPB_LAUNCH_CHECK("k1b_probe");
  return FPB_OK;
}
