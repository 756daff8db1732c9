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
k6_merge_kernel(const fpb_record* __restrict__ all, int n_shards, int B, int R, int P, int Rp2, int top_k,
                int64_t* __restrict__ out_ids, float* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [P]
  uint64_t* keys2 = keys + P;                               // [Rp2]
  uint32_t* pay = reinterpret_cast<uint32_t*>(keys2 + Rp2);  // [P] record index carried through the sort
  __shared__ int s_valid;
  const int b = blockIdx.x, tid = threadIdx.x;
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

}  // namespace

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
  if (!d_all_records || n_shards < 1 || B < 1 || R < 1 || top_k < 1) {
    fpb_set_error("fpb_merge_shards: bad arguments");
    return FPB_ERR_INVALID;
  }
  const int P = fpb_next_pow2(n_shards * R);
  const int Rp2 = fpb_next_pow2(R);
  const size_t smem = size_t(P + Rp2) * 8 + size_t(P) * 4;
  if (smem > 200 * 1024) {
    fpb_set_error("fpb_merge_shards: n_shards*R=%d records per query exceed the shared-memory sort", n_shards * R);
    return FPB_ERR_UNSUPPORTED;
  }
  static bool attr_done = false;
  if (!attr_done) {
    FPB_CUDA_CHECK(cudaFuncSetAttribute(k6_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  k6_merge_kernel<<<B, 1024, smem, static_cast<cudaStream_t>(stream)>>>(d_all_records, n_shards, B, R, P, Rp2,
                                                                       top_k, d_out_ids, d_out_scores, d_out_counts);
  FPB_LAUNCH_CHECK("k6_merge");
  return FPB_OK;
}
