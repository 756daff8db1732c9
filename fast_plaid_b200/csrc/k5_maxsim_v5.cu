// K5 v5 : fused residual decompression + exact MaxSim on tcgen05, many decode warps
// (dim=128, nbits=4, Qp <= 128; the default for Qp > 32).  Replaces search.rs:626-656 + :53-107 like v1..v4.
//
// v4 (register-resident mma.sync) is latency-bound: its 64 query-fragment registers limit an SM to
// 12 warps and the dependent chain  reduce -> sqrt -> rcp -> divide -> 8 chained HMMA  is not hidden
// (issue 55 %).  v3 (tcgen05) had the right division of labour -- the query tile is the A operand in
// shared memory, the accumulator lives in TMEM, so a decode warp needs neither query registers nor
// accumulators -- but only 14 decode warps and the old scalar decode.  v5 combines them:
//
//   * 800 threads: 20 decode warps at <= 80 registers, one MMA-issuing warp, 4 epilogue warps;
//   * decode is v4's: packed FFMA2/FMUL2, branch-free sqrt/rcp, one raw buffer whose loads for the
//     next pass are issued as soon as the current pass has been decoded;
//   * a tile is no longer "up to 112 tokens of ONE document": the chunk's documents are cut into
//     8-token passes, pass g goes to decode slot g % n_dec of tile g / n_dec, so every tile is full
//     (except the last of a chunk) whatever the document lengths.  A partially filled pass repeats
//     the document's last token, which cannot change a maximum, so no column masks are needed.
//   * D[q][t] = sum_k Q[q][k] E[t][k] with M = 128, N = 160, K = 128 as 8 tcgen05.mma.kind::f16 per
//     tile; 3 shared-memory stages, 2 TMEM accumulators.  The query tile is REPLICATED over the 128
//     accumulator rows (4 x 32 or 2 x 64) so every TMEM lane quarter holds all query tokens and the four
//     epilogue warps split the columns; per-document maxima are merged in shared memory (atomicMax on
//     order-preserving keys) and summed once per chunk.  (A first version with one epilogue warp walking
//     all 176 columns was epilogue-bound: 2.78 ms.)
#include "kernels.h"
#include "tc05.cuh"

namespace {

constexpr int V5_THREADS = 800;
constexpr int V5_NDEC = 20;                // decode warps 0..19 = pass slots of a tile
constexpr int V5_EPI0 = 20;                // epilogue warps 20..23: TMEM lane quarter = warp id % 4
constexpr int V5_MMA_WARP = 24;
static_assert(V5_THREADS == 32 * (V5_MMA_WARP + 1), "warp roles");
constexpr int V5_STAGES = 3;
constexpr int V5_ROWS = V5_NDEC * 8;       // 160 token rows per B stage = MMA N
constexpr int V5_ACC_STRIDE = 256;         // TMEM columns between the two accumulators
constexpr int V5_TMEM_COLS = 512;
constexpr int V5_MAX_DOCS = 32;
constexpr int V5_MAX_PASS = 2048;          // passes per chunk (host picks docs per chunk accordingly)
constexpr int V5_A_KBLOCK = 128 * 128;     // A operand: 128 rows x 128 B per K block
constexpr int V5_A_BYTES = 2 * V5_A_KBLOCK;
constexpr int V5_B_KBLOCK = V5_ROWS * 128;
constexpr int V5_B_BYTES = 2 * V5_B_KBLOCK;

struct DocMeta5 {
  int64_t o0;  // first token row of the document
  int len;
  int r;       // slot in the re-rank list
  int pfx;     // passes of the chunk before this document
  int pad;
};

struct V5Smem {
  static constexpr int a_off = 0;
  static constexpr int b_off = a_off + V5_A_BYTES;
  static constexpr int lut_off = b_off + V5_STAGES * V5_B_BYTES;
  static constexpr int prow_off = lut_off + 256 * 32 * 4;        // int64 first token row of every pass
  static constexpr int pnv_off = prow_off + V5_MAX_PASS * 8;     // uint8 valid tokens of every pass
  static constexpr int dmax_off = pnv_off + V5_MAX_PASS;         // [docs][128] running maxima (ordered keys)
  static constexpr int bar_off = dmax_off + V5_MAX_DOCS * 128 * 4;
  static constexpr int meta_off = bar_off + 128;
  static constexpr int meta_bytes = (V5_MAX_DOCS + 1) * int(sizeof(DocMeta5)) + 64;
  static constexpr int bytes = meta_off + meta_bytes + 1024;  // + slack for the 1024-byte alignment
};
static_assert(V5Smem::bytes <= 227 * 1024, "K5 v5 shared memory");

struct Raw5 {
  uint32_t w[4];  // residual words j, j+4, j+8, j+12 of the token
  uint4 c[4];     // centroid chunks j, j+4, j+8, j+12 (8 halves each)
};

__device__ __forceinline__ void v5_load_raw(Raw5& raw, const uint8_t* __restrict__ residuals,
                                            const __half* __restrict__ C, int64_t row, int code, int j) {
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(residuals + row * 64) + j;
  const uint4* cc = reinterpret_cast<const uint4*>(C + int64_t(code) * 128) + j;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    raw.w[k] = __ldg(rw + 4 * k);
    raw.c[k] = __ldg(cc + 4 * k);
  }
}

__device__ __forceinline__ float2 v5_fmul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 v5_ffma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
// IEEE fp32 e/n for both halves (q = e*r; rem = e - q*n exactly; q + rem*r), one rounding to fp16
__device__ __forceinline__ uint32_t v5_div2_pack(float2 e, float2 nneg, float2 r) {
  const float2 q = v5_fmul2(e, r);
  const float2 rem = v5_ffma2(q, nneg, e);
  const float2 res = v5_ffma2(rem, r, q);
  return pack_half2_rn(res.x, res.y);
}
// sqrt.rn / rcp.rn fast paths (see k5_maxsim_v4.cu and tools/check_sqrt_rcp.cu)
__device__ __forceinline__ float v5_sqrt_rn(float x) {
  float y, s, h;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(s) : "f"(x), "f"(y));
  asm("mul.rn.ftz.f32 %0, %1, 0f3F000000;" : "=f"(h) : "f"(y));
  const float r = __fmaf_rn(-s, s, x);
  return __fmaf_rn(r, h, s);
}
__device__ __forceinline__ float v5_rcp_rn(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float e = __fmaf_rn(x, y, -1.0f);
  return __fmaf_rn(y, -e, y);
}

__global__ void __launch_bounds__(V5_THREADS, 1)
k5_maxsim_v5_kernel(const __half* __restrict__ C, const int64_t* __restrict__ doc_offsets,
                    const int32_t* __restrict__ codes, const uint8_t* __restrict__ residuals,
                    const __half* __restrict__ norms, WPerm wp,
                    const __half* __restrict__ Qpad, int Q, int Qp, int B, int R, int docs_per_chunk,
                    const int32_t* __restrict__ n_rerank, const int32_t* __restrict__ rerank,
                    float* __restrict__ exact, int* __restrict__ counter) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t dyn_addr = smem_u32(smem_dyn);
  unsigned char* base = smem_dyn + ((1024u - (dyn_addr & 1023u)) & 1023u);  // SWIZZLE_128B atoms are 1024 B
  unsigned char* smA = base + V5Smem::a_off;
  unsigned char* smB = base + V5Smem::b_off;
  uint32_t* lut = reinterpret_cast<uint32_t*>(base + V5Smem::lut_off);
  int64_t* pass_row = reinterpret_cast<int64_t*>(base + V5Smem::prow_off);
  uint8_t* pass_nv = base + V5Smem::pnv_off;
  uint32_t* dmax = reinterpret_cast<uint32_t*>(base + V5Smem::dmax_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + V5Smem::bar_off);
  DocMeta5* docs = reinterpret_cast<DocMeta5*>(base + V5Smem::meta_off);
  int* misc = reinterpret_cast<int*>(base + V5Smem::meta_off + (V5_MAX_DOCS + 1) * sizeof(DocMeta5));
  // misc[0] chunk id, [1] tiles in chunk, [2] docs in chunk, [3] TMEM base, [4] passes in chunk
  const uint32_t bar_full = smem_u32(bars);        // [3]  decode -> MMA
  const uint32_t bar_empty = smem_u32(bars + 3);   // [3]  MMA -> decode
  const uint32_t bar_tfull = smem_u32(bars + 6);   // [2]  MMA -> epilogue
  const uint32_t bar_tempty = smem_u32(bars + 8);  // [2]  epilogue -> MMA

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_mma = warp == V5_MMA_WARP;
  const bool is_epi = warp >= V5_EPI0 && warp < V5_EPI0 + 4;
  const bool is_dec = warp < V5_NDEC;
  // The query tile is replicated over the 128 accumulator rows (4 copies of 32 rows, 2 of 64, 1 of 128) so that
  // every TMEM lane quarter holds all query tokens and the four epilogue warps can split the COLUMNS.
  const int qrep = Qp <= 32 ? 32 : (Qp <= 64 ? 64 : 128);

  // ---- one-time setup ----
  for (int i = tid; i < 256 * 32; i += V5_THREADS) {
    const int v = i >> 5;
    lut[i] = uint32_t(wp.v[v >> 4]) | (uint32_t(wp.v[v & 15]) << 16);
  }
  for (int i = tid; i < V5_A_BYTES / 16; i += V5_THREADS)
    reinterpret_cast<uint4*>(smA)[i] = make_uint4(0u, 0u, 0u, 0u);  // rows of padded query tokens stay zero
  if (tid == 0) {
    for (int s = 0; s < V5_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, V5_NDEC);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_tfull + 8 * t, 1);
      mbar_init(bar_tempty + 8 * t, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (is_mma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc[3])),
                 "n"(V5_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uint32_t(misc[3]);

  const int chunks_per_query = (R + docs_per_chunk - 1) / docs_per_chunk;
  const int total_chunks = B * chunks_per_query;
  int cur_b = -1;
  uint32_t gtile = 0;  // tiles processed by this CTA so far (same value in every thread)

  for (;;) {
    __syncthreads();  // all roles are done with the previous chunk
    if (tid == 0) misc[0] = atomicAdd(counter, 1);
    __syncthreads();
    const int chunk = misc[0];
    if (chunk >= total_chunks) break;
    const int b = chunk / chunks_per_query;
    const int r0 = (chunk % chunks_per_query) * docs_per_chunk;
    const int nr = n_rerank[b];
    if (r0 >= nr) continue;
    const int nd = min(docs_per_chunk, nr - r0);

    // ---- chunk metadata: documents, their passes (warp 0) ----
    if (warp == 0) {
      int np = 0, len = 0;
      int64_t o0 = 0;
      if (lane < nd) {  // one lane per document: the dependent loads run in parallel
        const int d = rerank[int64_t(b) * R + r0 + lane];
        o0 = doc_offsets[d];
        len = int(doc_offsets[d + 1] - o0);
        np = (len + 7) >> 3;
      }
      int incl = np;  // inclusive scan of the pass counts
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      if (lane < nd) {
        docs[lane].o0 = o0;
        docs[lane].len = len;
        docs[lane].r = r0 + lane;
        docs[lane].pfx = incl - np;
        for (int p = 0; p < np; ++p) {
          pass_row[incl - np + p] = o0 + 8 * p;
          pass_nv[incl - np + p] = uint8_t(min(8, len - 8 * p));
        }
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (lane == 0) {
        docs[nd].pfx = total;  // sentinel entry
        misc[1] = (total + V5_NDEC - 1) / V5_NDEC;
        misc[4] = total;
      }
    }
    for (int i = tid; i < nd * 128; i += V5_THREADS) dmax[i] = 0u;  // below the key of every float
    if (b != cur_b) {
      // Q tile, K-major SWIZZLE_128B: row r, 16-byte chunk c -> K block c/8, chunk (c%8) xor (r%8);
      // row r holds query token r % qrep (zero rows of the padded query stay zero)
      for (int i = tid; i < 128 * 16; i += V5_THREADS) {
        const int r = i >> 4, c = i & 15, q = r & (qrep - 1);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (q < Qp) v = *reinterpret_cast<const uint4*>(Qpad + (int64_t(b) * Qp + q) * 128 + c * 8);
        const int kb = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(smA + kb * V5_A_KBLOCK + (r >> 3) * 1024 + (r & 7) * 128 + ((cc ^ (r & 7)) << 4)) = v;
      }
      cur_b = b;
      fence_proxy_async();
    }
    __syncthreads();
    const int n_tiles = misc[1];
    const int n_pass = misc[4];

    if (is_dec) {
      // =========================== decode warps ===========================
      const int j = lane & 3, tslot = lane >> 2;
      const int prow = (tslot >> 1) + 4 * (tslot & 1);  // token of the pass handled by this lane group
      const uint32_t lut_lane = smem_u32(lut) + lane * 4;
      // a partially filled pass repeats the document's last token: a duplicate cannot change a maximum
      auto row_of = [&](int g) -> int64_t { return pass_row[g] + min(prow, int(pass_nv[g]) - 1); };
      int g = warp;
      Raw5 raw;
      int code_nxt = 0;
      __half nrm = __float2half(1.0f);  // the token's fp16 norm from the per-token table (derived at index load)
      if (g < n_pass) {
        const int64_t row = row_of(g);
        v5_load_raw(raw, residuals, C, row, __ldg(codes + row), j);
        nrm = __ldg(norms + row);
        if (g + V5_NDEC < n_pass) code_nxt = __ldg(codes + row_of(g + V5_NDEC));
      }
      for (int T = 0; T < n_tiles; ++T, g += V5_NDEC) {
        const uint32_t gt = gtile + T;
        const uint32_t stage = gt % V5_STAGES;
        if (g < n_pass) {
          // ---- decode the lane's 32 elements: e = fp16(w_perm[nibble] + centroid) ----
          float2 f[16];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t word = raw.w[k];
            const uint32_t cw[4] = {raw.c[k].x, raw.c[k].y, raw.c[k].z, raw.c[k].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t byte = (word >> (8 * i)) & 0xffu;
              uint32_t lv;
              asm("ld.shared.u32 %0, [%1];" : "=r"(lv) : "r"(lut_lane + byte * 128u));
              f[k * 4 + i] = __half22float2(__hadd2(u32_as_half2(lv), u32_as_half2(cw[i])));
            }
          }
          // ---- raw is dead: fetch pass g + 20, and the code of pass g + 40 ----
          const float nf = __half2float(nrm);
          if (g + V5_NDEC < n_pass) {
            const int64_t row = row_of(g + V5_NDEC);
            v5_load_raw(raw, residuals, C, row, code_nxt, j);
            nrm = __ldg(norms + row);
            if (g + 2 * V5_NDEC < n_pass) code_nxt = __ldg(codes + row_of(g + 2 * V5_NDEC));
          }
          // ---- norm from the per-token table, exact division ----
          const float rcp = v5_rcp_rn(nf);
          const float2 r2 = make_float2(rcp, rcp), nneg = make_float2(-nf, -nf);

          mbar_wait(bar_empty + 8 * stage, ((gt / V5_STAGES) & 1) ^ 1);
          unsigned char* st = smB + stage * V5_B_BYTES + warp * 1024 + prow * 128;  // row = slot*8 + prow
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = v5_div2_pack(f[k * 4 + i], nneg, r2);
            const int c = j + 4 * k, kb = c >> 3, cc = c & 7;
            *reinterpret_cast<uint4*>(st + kb * V5_B_KBLOCK + ((cc ^ prow) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          }
          fence_proxy_async();
          __syncwarp();
        } else {
          // the last tile of a chunk may have no pass for this slot; its arrival is still needed, in order
          mbar_wait(bar_empty + 8 * stage, ((gt / V5_STAGES) & 1) ^ 1);
        }
        if (lane == 0) mbar_arrive(bar_full + 8 * stage);
      }
    } else if (is_mma) {
      // =========================== MMA issuer ===========================
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smA);
        const uint32_t idesc = umma_idesc(V5_ROWS);
        for (int T = 0; T < n_tiles; ++T) {
          const uint32_t gt = gtile + T;
          const uint32_t stage = gt % V5_STAGES, acc = gt & 1;
          mbar_wait(bar_full + 8 * stage, (gt / V5_STAGES) & 1);
          mbar_wait(bar_tempty + 8 * acc, ((gt >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smB + stage * V5_B_BYTES);
          const uint32_t d_tmem = tmem_base + acc * V5_ACC_STRIDE;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a_off = (ks >> 2) * V5_A_KBLOCK + (ks & 3) * 32;
            const uint32_t b_off = (ks >> 2) * V5_B_KBLOCK + (ks & 3) * 32;
            umma_f16(d_tmem, umma_desc(a_addr + a_off), umma_desc(b_addr + b_off), idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * stage);  // stage reusable once these MMAs have read it
          umma_commit(bar_tfull + 8 * acc);    // accumulator ready
        }
      }
      __syncwarp();
    } else if (is_epi) {
      // =========================== epilogue ===========================
      // warp e reads TMEM lanes 32e..32e+31 = query tokens (32e + lane) % qrep of copy `copy`; the copies
      // split the 20 pass groups (8 accumulator columns each) of a tile between them.  Running maxima of
      // a document are merged into dmax[doc][q] (order-preserving keys, atomicMax) when the walk leaves it.
      const int e = warp - V5_EPI0;
      const int q = (32 * e + lane) & (qrep - 1);
      const int copies = 128 / qrep;                      // 4, 2 or 1
      const int copy = (32 * e) / qrep;
      const int gpc = V5_NDEC / copies;                   // groups per copy: 5, 10 or 20
      const int s_lo = copy * gpc;
      int di = 0, d_end = docs[1].pfx - 1;               // current document and its last pass
      float m = -INFINITY;
      bool dirty = false;
      for (int T = 0; T < n_tiles; ++T) {
        const uint32_t gt = gtile + T;
        const uint32_t acc = gt & 1;
        mbar_wait(bar_tfull + 8 * acc, (gt >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(32 * e) << 16) + acc * V5_ACC_STRIDE;
        const int g0 = T * V5_NDEC + s_lo;
        const int ng = min(gpc, n_pass - g0);  // may be <= 0 in the last tile
        for (int u = 0; u < ng; ++u) {
          const int g = g0 + u;
          if (g > d_end) {  // the walk leaves document di (possibly skipping documents of other copies)
            if (dirty) atomicMax(&dmax[di * 128 + q], f32_key(m));
            while (g >= docs[di + 1].pfx) ++di;
            d_end = docs[di + 1].pfx - 1;
            m = -INFINITY;
            dirty = false;
          }
          float v[8];
          tmem_ld8(taddr + (s_lo + u) * 8, v);
          m = fmaxf(fmaxf(m, fmaxf(v[0], v[1])), fmaxf(fmaxf(v[2], v[3]), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]))));
          dirty = true;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      }
      if (dirty) atomicMax(&dmax[di * 128 + q], f32_key(m));
      // ---- all four epilogue warps have merged: one score per document ----
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = e; i < nd; i += 4) {
        float sc;
        if (docs[i].len == 0) {
          sc = float(Q) * FPB_PAD_SENTINEL;  // no token: Q times the padding sentinel (search.rs:395)
        } else {
          sc = 0.f;
          for (int qq = lane; qq < Q; qq += 32) sc += __half2float(__float2half_rn(f32_unkey(dmax[i * 128 + qq])));
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, off);
        }
        if (lane == 0) exact[int64_t(b) * R + docs[i].r] = sc;
      }
    }
    gtile += uint32_t(n_tiles);
  }

  tc_fence_before();
  __syncthreads();
  if (is_mma) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(V5_TMEM_COLS));
  }
}

}  // namespace

int launch_maxsim_v5(const fpb_index* ix, const Ws& ws, cudaStream_t st, bool* handled) {
  *handled = false;
  const fpb_layout& L = *ws.L;
  if (ix->dim != 128 || ix->nbits != 4 || L.Qp > 128) return FPB_OK;
  *handled = true;
  // opt in on every launch: the attribute is per device and the call costs about a microsecond
  FPB_CUDA_CHECK(cudaFuncSetAttribute(k5_maxsim_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V5Smem::bytes));
  WPerm wp;
  for (int i = 0; i < 16; ++i) wp.v[i] = ix->w_perm_bits[i];
  int* counter = ws.work() + L.B + 3;
  FPB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), st));
  // documents per chunk: enough tiles to amortise the pipeline fill/drain, enough chunks to balance the SMs
  const int64_t total_docs = int64_t(L.B) * L.R;
  const int64_t passes_per_doc = (ix->max_doc_len + 7) / 8;
  if (passes_per_doc < 1 || passes_per_doc > V5_MAX_PASS) {
    *handled = false;  // a single document does not fit the pass table: the generic kernel takes it
    return FPB_OK;
  }
  int docs_per_chunk = V5_MAX_DOCS;
  while (docs_per_chunk > 1 && docs_per_chunk * passes_per_doc > V5_MAX_PASS) docs_per_chunk >>= 1;
  while (docs_per_chunk > 4 && total_docs / docs_per_chunk < int64_t(ix->sm_count) * 8) docs_per_chunk >>= 1;
  const int chunks = L.B * ((L.R + docs_per_chunk - 1) / docs_per_chunk);
  const int blocks = chunks < ix->sm_count ? chunks : ix->sm_count;
  k5_maxsim_v5_kernel<<<blocks, V5_THREADS, V5Smem::bytes, st>>>(
      ix->centroids, ix->doc_offsets, ix->doc_codes, ix->doc_residuals, ix->token_norms, wp, ws.queries(), L.Q, L.Qp,
      L.B, L.R,
      docs_per_chunk, ws.n_rerank(), ws.rerank(), ws.exact(), counter);
  FPB_LAUNCH_CHECK("k5_maxsim_v5");
  return FPB_OK;
}
