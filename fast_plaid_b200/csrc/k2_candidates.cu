// K2 : candidate documents = sorted-unique union of the IVF lists of the probed cells
//      (search.rs:531-541: unique cells -> ivf lookup -> sort -> unique_consecutive).
//
// The reference concatenates the lists, radix-sorts ~300k int64 ids and compacts, per query.
// Here every probed list is OR-ed into a per-query bitmap over the local doc ids and the
// bitmap is compacted in index order, which yields the sorted-unique id list directly.
#include "kernels.h"

namespace {

// grid (Q*n_probe, B): one CTA per probe slot.  A slot whose cell already appears in an
// earlier slot of the same query is skipped (unique_dim of search.rs:531-532).
__global__ void __launch_bounds__(128)
k2_mark_kernel(const int32_t* __restrict__ cells, int slots, const int64_t* __restrict__ ivf_offsets,
               const int32_t* __restrict__ ivf_pids, uint32_t* __restrict__ bitmap, int bitmap_words) {
  const int b = blockIdx.y, s = blockIdx.x;
  const int32_t* cq = cells + int64_t(b) * slots;
  const int32_t c = cq[s];
  if (c < 0) return;
  int dup = 0;
  for (int i = threadIdx.x; i < s; i += blockDim.x) dup |= (cq[i] == c);
  if (__syncthreads_or(dup)) return;
  const int64_t o0 = ivf_offsets[c], o1 = ivf_offsets[c + 1];
  uint32_t* bm = bitmap + int64_t(b) * bitmap_words;
  for (int64_t i = o0 + threadIdx.x; i < o1; i += blockDim.x) {
    const int32_t pid = __ldg(ivf_pids + i);
    atomicOr(bm + (pid >> 5), 1u << (pid & 31));
  }
}

// one CTA per query: ordered compaction of the bitmap into cand[b][0..n_cand[b])
__global__ void __launch_bounds__(1024)
k2_compact_kernel(const uint32_t* __restrict__ bitmap, int bitmap_words, int32_t* __restrict__ cand,
                  int cand_cap, int32_t* __restrict__ n_cand) {
  __shared__ int warp_sums[32];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const uint32_t* bm = bitmap + int64_t(b) * bitmap_words;
  int32_t* out = cand + int64_t(b) * cand_cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int w0 = 0; w0 < bitmap_words; w0 += 1024) {
    const int wi = w0 + tid;
    uint32_t w = (wi < bitmap_words) ? bm[wi] : 0u;
    const int c = __popc(w);
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int ws = warp_sums[lane];
      int wincl = ws;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, wincl, off);
        if (lane >= off) wincl += v;
      }
      warp_sums[lane] = wincl - ws;  // exclusive
    }
    __syncthreads();
    const int base = s_base;
    int pos = base + warp_sums[warp] + incl - c;
    while (w) {
      const int bit = __ffs(w) - 1;
      w &= w - 1;
      out[pos++] = wi * 32 + bit;
    }
    __syncthreads();
    if (tid == 1023) s_base = base + warp_sums[31] + incl;  // total of this round
    __syncthreads();
  }
  if (tid == 0) n_cand[b] = s_base;
}

}  // namespace

int launch_candidates(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  FPB_CUDA_CHECK(cudaMemsetAsync(ws.bitmap(), 0, size_t(L.B) * L.bitmap_words * 4, st));
  const int slots = L.Q * L.n_probe;
  dim3 grid(slots, L.B);
  k2_mark_kernel<<<grid, 128, 0, st>>>(ws.cells(), slots, ix->ivf_offsets, ix->ivf_pids, ws.bitmap(),
                                       L.bitmap_words);
  FPB_LAUNCH_CHECK("k2_mark");
  k2_compact_kernel<<<L.B, 1024, 0, st>>>(ws.bitmap(), L.bitmap_words, ws.cand(), L.cand_cap,
                                          ws.n_cand());
  FPB_LAUNCH_CHECK("k2_compact");
  return FPB_OK;
}
