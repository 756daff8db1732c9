// K6 : final ranking (search.rs:659-666) and the document-sharded merge (new).
//
// Canonical order: larger exact score first, then smaller doc id.  (The reference's
// non-stable sort(descending) leaves equal scores in an implementation-defined order.)
#include "kernels.h"

namespace {

__device__ __forceinline__ void bitonic_desc(uint64_t* keys, int P, int tid, int nthreads) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += nthreads) {
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
}

__device__ __forceinline__ void bitonic_desc_payload(uint64_t* keys, uint32_t* pay, int P, int tid, int nthreads) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += nthreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const uint64_t x = keys[i], y = keys[ixj];
          if ((x < y) == up) {
            keys[i] = y;
            keys[ixj] = x;
            const uint32_t px = pay[i];
            pay[i] = pay[ixj];
            pay[ixj] = px;
          }
        }
      }
      __syncthreads();
    }
  }
}

// one CTA per query
__global__ void __launch_bounds__(1024)
k6_rank_kernel(const float* __restrict__ exact, const int32_t* __restrict__ rerank,
               const int32_t* __restrict__ n_rerank, int R, int Rp2, int top_k, int64_t doc_id_base,
               int64_t* __restrict__ out_ids, float* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = n_rerank[b];
  for (int i = tid; i < Rp2; i += 1024) {
    uint64_t k = 0;
    if (i < n) k = (uint64_t(f32_key(exact[int64_t(b) * R + i])) << 32) |
                   uint64_t(0xffffffffu - uint32_t(rerank[int64_t(b) * R + i]));
    keys[i] = k;
  }
  __syncthreads();
  bitonic_desc(keys, Rp2, tid, 1024);
  const int cnt = min(top_k, n);
  for (int i = tid; i < top_k; i += 1024) {
    int64_t id = -1;
    float sc = -INFINITY;
    if (i < cnt) {
      id = doc_id_base + int64_t(0xffffffffu - uint32_t(keys[i]));
      sc = f32_unkey(uint32_t(keys[i] >> 32));
    }
    out_ids[int64_t(b) * top_k + i] = id;
    out_scores[int64_t(b) * top_k + i] = sc;
  }
  if (tid == 0) out_counts[b] = cnt;
}

__global__ void emit_records_kernel(const float* __restrict__ exact, const float* __restrict__ rerank_approx,
                                    const int32_t* __restrict__ rerank, const int32_t* __restrict__ n_rerank,
                                    int B, int R, int64_t doc_id_base, fpb_record* __restrict__ rec) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= int64_t(B) * R) return;
  const int b = int(i / R), r = int(i % R);
  fpb_record o;
  if (r < n_rerank[b]) {
    o.approx = rerank_approx[i];
    o.exact = exact[i];
    o.doc_id = doc_id_base + rerank[i];
  } else {
    o.approx = -INFINITY;
    o.exact = -INFINITY;
    o.doc_id = -1;
  }
  rec[i] = o;
}

// one CTA per query: re-apply the global pruning rule over the gathered records, then rank.
//   keep the R best by (approx desc, doc id asc)   -- search.rs:605-619 on the whole index
//   order them by  (exact desc, doc id asc)        -- search.rs:659
__global__ void __launch_bounds__(1024)
k6_merge_kernel(const fpb_record* __restrict__ all_groups, int n_shards, int B, int R, int P, int Rp2, int top_k,
                int64_t* __restrict__ out_ids, float* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [P]
  uint64_t* keys2 = keys + P;                               // [Rp2]
  uint32_t* pay = reinterpret_cast<uint32_t*>(keys2 + Rp2);  // [P] record index carried through the sort
  __shared__ int s_valid;
  // query blockIdx.x of the whole batch = query `b` of query group blockIdx.x / B, whose records are the
  // [n_shards, B, R] block of that group (one group: the block index is the query)
  const int b = blockIdx.x % B, tid = threadIdx.x;
  const fpb_record* all = all_groups + int64_t(blockIdx.x / B) * n_shards * B * R;
  out_ids += int64_t(blockIdx.x - b) * top_k;
  out_scores += int64_t(blockIdx.x - b) * top_k;
  out_counts += blockIdx.x - b;
  const int total = n_shards * R;
  if (tid == 0) s_valid = 0;
  __syncthreads();
  int local_valid = 0;
  for (int i = tid; i < P; i += 1024) {
    uint64_t k = 0;
    if (i < total) {
      const int s = i / R, r = i % R;
      const fpb_record rec = all[(int64_t(s) * B + b) * R + r];
      if (rec.doc_id >= 0) {
        k = (uint64_t(f32_key(rec.approx)) << 32) | uint64_t(0xffffffffu - uint32_t(rec.doc_id));
        ++local_valid;
      }
    }
    keys[i] = k;
    pay[i] = uint32_t(i);
  }
  if (local_valid) atomicAdd(&s_valid, local_valid);
  __syncthreads();
  bitonic_desc_payload(keys, pay, P, tid, 1024);
  const int keep = min(s_valid, R);
  for (int i = tid; i < Rp2; i += 1024) {
    uint64_t k2 = 0;
    if (i < keep) {
      const int j = int(pay[i]);
      const int s = j / R, r = j % R;
      const fpb_record rec = all[(int64_t(s) * B + b) * R + r];
      k2 = (uint64_t(f32_key(rec.exact)) << 32) | uint64_t(0xffffffffu - uint32_t(rec.doc_id));
    }
    keys2[i] = k2;
  }
  __syncthreads();
  bitonic_desc(keys2, Rp2, tid, 1024);
  const int cnt = min(top_k, keep);
  for (int i = tid; i < top_k; i += 1024) {
    int64_t id = -1;
    float sc = -INFINITY;
    if (i < cnt) {
      id = int64_t(0xffffffffu - uint32_t(keys2[i]));
      sc = f32_unkey(uint32_t(keys2[i] >> 32));
    }
    out_ids[int64_t(b) * top_k + i] = id;
    out_scores[int64_t(b) * top_k + i] = sc;
  }
  if (tid == 0) out_counts[b] = cnt;
}

// two-step sharded search, step 1: 64-bit keys of the local pruned list
__global__ void emit_keys_kernel(const float* __restrict__ rerank_approx, const int32_t* __restrict__ rerank,
                                 const int32_t* __restrict__ n_rerank, int B, int R, int64_t doc_id_base,
                                 uint64_t* __restrict__ keys) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= int64_t(B) * R) return;
  const int b = int(i / R), r = int(i % R);
  uint64_t k = 0;
  if (r < n_rerank[b])
    k = (uint64_t(f32_key(rerank_approx[i])) << 32) | uint64_t(0xffffffffu - uint32_t(doc_id_base + rerank[i]));
  keys[i] = k;
}

// step 2: global R-th best key per query; the local list keeps (in order) the entries at or above
// it.  One CTA per query.
__global__ void __launch_bounds__(1024)
apply_threshold_kernel(const uint64_t* __restrict__ all, int n_shards, int rank, int B, int R, int P,
                       int32_t* __restrict__ n_rerank, int32_t* __restrict__ rerank,
                       float* __restrict__ rerank_approx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ int warp_sums[32];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = n_shards * R;
  for (int i = tid; i < P; i += 1024) {
    uint64_t k = 0;
    if (i < total) k = all[(int64_t(i / R) * B + b) * R + (i % R)];
    keys[i] = k;
  }
  __syncthreads();
  bitonic_desc(keys, P, tid, 1024);
  const uint64_t T = keys[R - 1];  // 0 when the whole index has fewer than R candidates
  // ordered compaction of the local list (it is in id order when nothing was pruned locally):
  // thread t owns entries [t*PER, (t+1)*PER), one block-wide exclusive scan gives the positions
  const uint64_t* mine = all + (int64_t(rank) * B + b) * R;
  int32_t* rr = rerank + int64_t(b) * R;
  float* ra = rerank_approx + int64_t(b) * R;
  int32_t* tmp_id = reinterpret_cast<int32_t*>(keys);   // reuse the sort buffer (P*8 >= R*8 bytes)
  float* tmp_ap = reinterpret_cast<float*>(tmp_id + R);
  const int n_old = n_rerank[b];
  constexpr int PER = 4;  // R <= 4096
  int32_t id[PER];
  float ap[PER];
  bool keep[PER];
  int c = 0;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int i = tid * PER + u;
    keep[u] = false;
    if (i < n_old && i < R) {
      const uint64_t k = mine[i];
      keep[u] = (k != 0) && (k >= T);
      id[u] = rr[i];
      ap[u] = ra[i];
    }
    c += keep[u] ? 1 : 0;
  }
  int incl = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();  // also: everybody has read T / keys[] before tmp_* overwrites the buffer
  if (warp == 0) {
    const int ws = warp_sums[lane];
    int wincl = ws;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wincl, off);
      if (lane >= off) wincl += v;
    }
    warp_sums[lane] = wincl - ws;
    if (lane == 31) s_base = wincl;
  }
  __syncthreads();
  int pos = warp_sums[warp] + incl - c;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    if (keep[u]) {
      tmp_id[pos] = id[u];
      tmp_ap[pos] = ap[u];
      ++pos;
    }
  }
  __syncthreads();
  const int n_new = s_base;
  for (int i = tid; i < n_new; i += 1024) {
    rr[i] = tmp_id[i];
    ra[i] = tmp_ap[i];
  }
  if (tid == 0) n_rerank[b] = n_new;
}

}  // namespace

int launch_emit_keys(const fpb_index* ix, const Ws& ws, uint64_t* d_keys, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  const int64_t n = int64_t(L.B) * L.R;
  emit_keys_kernel<<<int((n + 255) / 256), 256, 0, st>>>(ws.rerank_approx(), ws.rerank(), ws.n_rerank(), L.B, L.R,
                                                        ix->doc_id_base, d_keys);
  FPB_LAUNCH_CHECK("emit_keys");
  return FPB_OK;
}

int launch_apply_threshold(const Ws& ws, const uint64_t* d_all_keys, int n_shards, int rank, cudaStream_t st,
                           int b_stride) {
  const fpb_layout& L = *ws.L;
  if (b_stride <= 0) b_stride = L.B;  // queries per shard in the gathered key array
  const int P = fpb_next_pow2(n_shards * L.R);
  const size_t smem = size_t(P) * 8;
  if (smem > 200 * 1024) {
    fpb_set_error("apply_threshold: n_shards*R=%d keys per query exceed the shared-memory sort", n_shards * L.R);
    return FPB_ERR_UNSUPPORTED;
  }
  // opt in on every launch: the attribute is per device and the call costs about a microsecond
  FPB_CUDA_CHECK(cudaFuncSetAttribute(apply_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  apply_threshold_kernel<<<L.B, 1024, smem, st>>>(d_all_keys, n_shards, rank, b_stride, L.R, P, ws.n_rerank(),
                                                  ws.rerank(), ws.rerank_approx());
  FPB_LAUNCH_CHECK("apply_threshold");
  return FPB_OK;
}

int launch_rank(const fpb_index* ix, const Ws& ws, int top_k, int64_t* d_out_ids, float* d_out_scores,
                int32_t* d_out_counts, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  const int Rp2 = fpb_next_pow2(L.R);
  k6_rank_kernel<<<L.B, 1024, size_t(Rp2) * 8, st>>>(ws.exact(), ws.rerank(), ws.n_rerank(), L.R, Rp2, top_k,
                                                    ix->doc_id_base, d_out_ids, d_out_scores, d_out_counts);
  FPB_LAUNCH_CHECK("k6_rank");
  return FPB_OK;
}

int launch_emit_records(const fpb_index* ix, const Ws& ws, fpb_record* d_records, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  const int64_t n = int64_t(L.B) * L.R;
  emit_records_kernel<<<int((n + 255) / 256), 256, 0, st>>>(ws.exact(), ws.rerank_approx(), ws.rerank(),
                                                           ws.n_rerank(), L.B, L.R, ix->doc_id_base, d_records);
  FPB_LAUNCH_CHECK("emit_records");
  return FPB_OK;
}

extern "C" int fpb_merge_shards(const fpb_record* d_all_records, int n_shards, int B, int R, int top_k,
                                int64_t* d_out_ids, float* d_out_scores, int32_t* d_out_counts, void* stream) {
  return launch_merge(d_all_records, n_shards, B, B, R, top_k, d_out_ids, d_out_scores, d_out_counts,
                      static_cast<cudaStream_t>(stream));
}

// merge of `n_queries` queries of gathered records laid out [n_groups][n_shards, b_stride, R]: query q is query
// q % b_stride of group q / b_stride (n_queries <= b_stride: one group)
int launch_merge(const fpb_record* d_all_records, int n_shards, int b_stride, int n_queries, int R, int top_k,
                 int64_t* d_out_ids, float* d_out_scores, int32_t* d_out_counts, cudaStream_t stream) {
  const int B = b_stride;
  if (!d_all_records || n_shards < 1 || B < 1 || R < 1 || top_k < 1 || n_queries < 0) {
    fpb_set_error("fpb_merge_shards: bad arguments");
    return FPB_ERR_INVALID;
  }
  if (n_queries == 0) return FPB_OK;
  const int P = fpb_next_pow2(n_shards * R);
  const int Rp2 = fpb_next_pow2(R);
  const size_t smem = size_t(P + Rp2) * 8 + size_t(P) * 4;
  if (smem > 200 * 1024) {
    fpb_set_error("fpb_merge_shards: n_shards*R=%d records per query exceed the shared-memory sort", n_shards * R);
    return FPB_ERR_UNSUPPORTED;
  }
  // opt in on every launch: the attribute is per device and the call costs about a microsecond
  FPB_CUDA_CHECK(cudaFuncSetAttribute(k6_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k6_merge_kernel<<<n_queries, 1024, smem, stream>>>(d_all_records, n_shards, B, R, P, Rp2, top_k, d_out_ids,
                                                    d_out_scores, d_out_counts);
  FPB_LAUNCH_CHECK("k6_merge");
  return FPB_OK;
}
