// K5 v3 : fused residual decompression + exact MaxSim on the 5th-generation tensor cores
// (tcgen05.mma, accumulator in TMEM) for dim=128, nbits=4, Qp <= 128.
//
// Orientation.  D[q][t] = sum_k Q[q][k] * E[t][k]: the query tile is the A operand (M = 128 rows,
// rows >= Qp are zero), the decompressed tokens of ONE document tile are the B operand
// (N = up to 128 tokens, K = 128).  The accumulator row of a query token lives in one TMEM lane,
// so the MaxSim reduction max_t is a per-thread running maximum over the columns an epilogue
// thread reads back with tcgen05.ld -- no cross-lane shuffles, and the tile never touches the
// LSU pipe again after the decode warps have written it (v2 spent 6 of its 17.5 L1/shared
// wavefronts per token on ldmatrix and 3.5 of its 51 instructions per token on HMMA/LDSM).
//
// Warp roles in a 512-thread CTA (one CTA per SM, 160 KB of shared memory):
//   warps 0..10  decode: 4 lanes per token, 8 tokens per pass; bank-replicated LUT lookup,
//                fp16 add of the centroid, fp32 norm, exact division (same arithmetic as v1/v2),
//                16-byte stores into the B stage in the canonical K-major SWIZZLE_128B layout;
//                passes are dealt round-robin to the 11 warps across tiles
//   warp  11     one elected lane issues 8 x tcgen05.mma (kind::f16, M=128, N=tile, K=16) per tile
//                and commits to the "stage free" and "accumulator ready" mbarriers
//   warps 12..15 epilogue: warp 12+i owns TMEM lanes 32i..32i+31 (query tokens); running max in
//                fp32 (rounding to fp16 is monotone, so max-then-round == round-then-max),
//                fp32 sum over query tokens, one score per document
// Pipelines: 3 shared-memory stages (decode -> MMA), 2 TMEM accumulators (MMA -> epilogue).
#include "kernels.h"

namespace {

constexpr int V3_THREADS = 512;
constexpr int V3_MMA_WARP = 11;
constexpr int V3_EPI_WARP0 = 12;
constexpr int V3_STAGES = 3;
constexpr int V3_TILE_N = 128;       // accumulator slot width in TMEM columns (tiles hold <= 112 tokens)
constexpr int V3_MAX_TILES = 256;    // tiles per chunk (host picks docs per chunk accordingly)
constexpr int V3_MAX_DOCS = 32;
constexpr int V3_KBLOCK_BYTES = 128 * 128;        // 128 rows x 128 B (64 fp16) per K block
constexpr int V3_TILE_BYTES = 2 * V3_KBLOCK_BYTES;  // K = 128 = two K blocks
constexpr int V3_TMEM_COLS = 256;                  // 2 accumulators x 128 fp32 columns

struct V3Smem {
  // offsets from a 1024-byte aligned base
  static constexpr int a_off = 0;                                   // Q tile (A operand)
  static constexpr int b_off = a_off + V3_TILE_BYTES;               // 3 token stages (B operand)
  static constexpr int lut_off = b_off + V3_STAGES * V3_TILE_BYTES;  // 256 x 32 x u32
  static constexpr int bar_off = lut_off + 256 * 32 * 4;             // mbarriers
  static constexpr int meta_off = bar_off + 128;
  static constexpr int meta_bytes = V3_MAX_TILES * 16 + V3_MAX_DOCS * 16 + 64;
  static constexpr int bytes = meta_off + meta_bytes + 1024;        // + slack for the alignment
};

struct Raw3 {
  uint32_t w[4];
  uint4 c[4];
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// and offsets in 16-byte units, LBO = 1 (unused for swizzled K-major), SBO = 1024 B between 8-row
// groups, version = 1 (Blackwell), layout_type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFFu) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float v3_div_rn(float e, float n, float r) {
  const float q = __fmul_rn(e, r);
  const float rem = __fmaf_rn(-q, n, e);
  return __fmaf_rn(rem, r, q);
}

__device__ __forceinline__ void v3_load_raw(Raw3& raw, const uint8_t* __restrict__ residuals,
                                            const __half* __restrict__ C, int64_t tok_global, int code, int j) {
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(residuals + tok_global * 64) + j;
  const uint4* cc = reinterpret_cast<const uint4*>(C + int64_t(code) * 128) + j;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    raw.w[k] = __ldg(rw + 4 * k);
    raw.c[k] = __ldg(cc + 4 * k);
  }
}

// per-chunk metadata in shared memory
struct TileMeta {
  int doc;      // index of the document inside the chunk
  int tok0;     // first token of the tile inside the document
  int nvalid;   // tokens in the tile (1..128)
  int pass0;    // number of 8-token passes of the chunk before this tile
};
struct DocMeta {
  int64_t o0;   // first token row of the document
  int len;
  int r;        // slot in the re-rank list
};

__global__ void __launch_bounds__(V3_THREADS, 1)
k5_maxsim_v3_kernel(const __half* __restrict__ C, const int64_t* __restrict__ doc_offsets,
                    const int32_t* __restrict__ codes, const uint8_t* __restrict__ residuals, WPerm wp,
                    const __half* __restrict__ Qpad, int Q, int Qp, int B, int R, int docs_per_chunk, int n_dec,
                    const int32_t* __restrict__ n_rerank, const int32_t* __restrict__ rerank,
                    float* __restrict__ exact, int* __restrict__ counter) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned base (SWIZZLE_128B atoms are 1024 B)
  const uint32_t dyn_addr = smem_u32(smem_dyn);
  unsigned char* base = smem_dyn + ((1024u - (dyn_addr & 1023u)) & 1023u);
  unsigned char* smA = base + V3Smem::a_off;
  unsigned char* smB = base + V3Smem::b_off;
  uint32_t* lut = reinterpret_cast<uint32_t*>(base + V3Smem::lut_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + V3Smem::bar_off);
  TileMeta* tiles = reinterpret_cast<TileMeta*>(base + V3Smem::meta_off);
  DocMeta* docs = reinterpret_cast<DocMeta*>(base + V3Smem::meta_off + V3_MAX_TILES * 16);
  int* misc = reinterpret_cast<int*>(base + V3Smem::meta_off + V3_MAX_TILES * 16 + V3_MAX_DOCS * 16);
  // misc[0] = chunk id, misc[1] = tiles in chunk, misc[2] = docs in chunk, misc[3] = TMEM base, misc[4..7] epilogue partial sums
  const uint32_t bar_full = smem_u32(bars);             // [3]
  const uint32_t bar_empty = smem_u32(bars + 3);        // [3]
  const uint32_t bar_tfull = smem_u32(bars + 6);        // [2]
  const uint32_t bar_tempty = smem_u32(bars + 8);       // [2]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_epi_warps = (Qp + 31) / 32;
  // decode slots: warps 0..10 plus the epilogue warps this query length leaves idle; pass s of every
  // tile belongs to slot s, so a tile holds n_dec passes = 8*n_dec tokens
  const int dec_slot = (warp < V3_MMA_WARP) ? warp
                       : (warp >= V3_EPI_WARP0 + n_epi_warps ? V3_MMA_WARP + (warp - V3_EPI_WARP0 - n_epi_warps) : -1);
  const bool is_decoder = dec_slot >= 0 && dec_slot < n_dec;
  const int tile_tokens = 8 * n_dec;

  // ---- one-time setup ----
  for (int i = tid; i < 256 * 32; i += V3_THREADS) {
    const int v = i >> 5;
    lut[i] = uint32_t(wp.v[v >> 4]) | (uint32_t(wp.v[v & 15]) << 16);
  }
  for (int i = tid; i < V3_TILE_BYTES / 16; i += V3_THREADS)
    reinterpret_cast<uint4*>(smA)[i] = make_uint4(0u, 0u, 0u, 0u);  // query rows >= Qp stay zero
  if (tid == 0) {
    for (int s = 0; s < V3_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, n_dec);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_tfull + 8 * t, 1);
      mbar_init(bar_tempty + 8 * t, n_epi_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == V3_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc[3])),
                 "n"(V3_TMEM_COLS));
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

    // ---- chunk metadata + (on a query change) the A tile ----
    if (warp == 0) {
      const int nd = min(docs_per_chunk, nr - r0);
      if (lane < nd) {  // one lane per document: the dependent loads run in parallel
        const int d = rerank[int64_t(b) * R + r0 + lane];
        const int64_t o0 = doc_offsets[d];
        docs[lane].o0 = o0;
        docs[lane].len = int(doc_offsets[d + 1] - o0);
        docs[lane].r = r0 + lane;
      }
      __syncwarp();
      if (lane == 0) {
        int nt = 0, np = 0;
        for (int i = 0; i < nd; ++i) {
          const int len = docs[i].len;
          for (int t0 = 0; t0 < len; t0 += tile_tokens) {
            tiles[nt].doc = i;
            tiles[nt].tok0 = t0;
            tiles[nt].nvalid = min(tile_tokens, len - t0);
            tiles[nt].pass0 = np;
            np += (tiles[nt].nvalid + 7) >> 3;
            ++nt;
          }
        }
        misc[1] = nt;
        misc[2] = nd;
      }
    }
    if (b != cur_b) {
      // Q tile, K-major SWIZZLE_128B: row q, 16-byte chunk c -> kb = c/8, cc = c%8, xor with q%8
      for (int i = tid; i < Qp * 16; i += V3_THREADS) {
        const int q = i >> 4, c = i & 15;
        const uint4 v = *reinterpret_cast<const uint4*>(Qpad + (int64_t(b) * Qp + q) * 128 + c * 8);
        const int kb = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(smA + kb * V3_KBLOCK_BYTES + (q >> 3) * 1024 + (q & 7) * 128 + ((cc ^ (q & 7)) << 4)) = v;
      }
      cur_b = b;
      fence_proxy_async();
    }
    __syncthreads();
    const int n_tiles = misc[1];
    const int n_docs = misc[2];

    if (is_decoder) {
      // =========================== decode warps ===========================
      const int j = lane & 3, tslot = lane >> 2;
      const int prow = (tslot >> 1) + 4 * (tslot & 1);
      const uint32_t lut_lane = smem_u32(lut) + lane * 4;
      const int p = dec_slot;  // this warp's pass inside every tile
      auto has_pass = [&](int T) { return p * 8 < tiles[T].nvalid; };
      auto next_tile = [&](int T) {  // first tile >= T where this slot has a pass
        while (T < n_tiles && !has_pass(T)) ++T;
        return T;
      };
      // software pipeline: codes are fetched two tiles ahead, residual + centroid rows one tile ahead
      auto tok_of = [&](int T, int64_t& row) {
        const TileMeta tm = tiles[T];
        const DocMeta dm = docs[tm.doc];
        row = dm.o0 + min(tm.tok0 + p * 8 + prow, dm.len - 1);
      };

      int T = next_tile(0);
      int T1 = (T < n_tiles) ? next_tile(T + 1) : n_tiles;
      int done = 0;  // tiles [0, done) have received this warp's arrival
      int64_t row_cur = 0, row_1 = 0;
      int code_1 = 0;
      Raw3 cur;
      if (T < n_tiles) {
        tok_of(T, row_cur);
        const int code_0 = __ldg(codes + row_cur);
        if (T1 < n_tiles) {
          tok_of(T1, row_1);
          code_1 = __ldg(codes + row_1);
        }
        v3_load_raw(cur, residuals, C, row_cur, code_0, j);
      }
      while (T < n_tiles) {
        const int T2 = (T1 < n_tiles) ? next_tile(T1 + 1) : n_tiles;
        int64_t row_2 = 0;
        int code_2 = 0;
        if (T2 < n_tiles) {
          tok_of(T2, row_2);
          code_2 = __ldg(codes + row_2);
        }
        Raw3 nxt;
        if (T1 < n_tiles) v3_load_raw(nxt, residuals, C, row_1, code_1, j);

        // tiles without a pass of this slot still need its arrival, in order
        while (done < T) {
          const uint32_t g = gtile + done;
          mbar_wait(bar_empty + 8 * (g % V3_STAGES), ((g / V3_STAGES) & 1) ^ 1);
          if (lane == 0) mbar_arrive(bar_full + 8 * (g % V3_STAGES));
          ++done;
        }
        const uint32_t g = gtile + T;
        const uint32_t stage = g % V3_STAGES;
        mbar_wait(bar_empty + 8 * stage, ((g / V3_STAGES) & 1) ^ 1);

        // ---- decode: e = fp16(w_perm[nibble] + centroid), n = fp16(sqrt(sum e^2)), e_hat = fp16(e / n) ----
        __half2 e[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t word = cur.w[k];
          const uint32_t cw[4] = {cur.c[k].x, cur.c[k].y, cur.c[k].z, cur.c[k].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t byte = (word >> (8 * i)) & 0xffu;
            uint32_t lv;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lv) : "r"(lut_lane + byte * 128u));
            e[k * 4 + i] = __hadd2(u32_as_half2(lv), u32_as_half2(cw[i]));
          }
        }
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 f = __half22float2(e[i]);
          ss = __fmaf_rn(f.x, f.x, ss);
          ss = __fmaf_rn(f.y, f.y, ss);
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        const float nf = __half2float(__float2half_rn(sqrtf(ss)));
        const float rcp = __frcp_rn(nf);
        unsigned char* st = smB + stage * V3_TILE_BYTES + p * 1024 + prow * 128;  // row = p*8 + prow
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(e[k * 4 + i]);
            o[i] = pack_half2_rn(v3_div_rn(f.x, nf, rcp), v3_div_rn(f.y, nf, rcp));
          }
          const int c = j + 4 * k, kb = c >> 3, cc = c & 7;
          *reinterpret_cast<uint4*>(st + kb * V3_KBLOCK_BYTES + ((cc ^ prow) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * stage);
        done = T + 1;
        T = T1;
        T1 = T2;
        row_1 = row_2;
        code_1 = code_2;
        cur = nxt;
      }
      while (done < n_tiles) {
        const uint32_t g = gtile + done;
        mbar_wait(bar_empty + 8 * (g % V3_STAGES), ((g / V3_STAGES) & 1) ^ 1);
        if (lane == 0) mbar_arrive(bar_full + 8 * (g % V3_STAGES));
        ++done;
      }
    } else if (warp == V3_MMA_WARP) {
      // =========================== MMA issuer ===========================
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smA);
        for (int T = 0; T < n_tiles; ++T) {
          const uint32_t g = gtile + T;
          const uint32_t stage = g % V3_STAGES, acc = g & 1;
          mbar_wait(bar_full + 8 * stage, (g / V3_STAGES) & 1);
          mbar_wait(bar_tempty + 8 * acc, ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          const int n = (tiles[T].nvalid + 15) & ~15;
          const uint32_t idesc = umma_idesc(n);
          const uint32_t b_addr = smem_u32(smB + stage * V3_TILE_BYTES);
          const uint32_t d_tmem = tmem_base + acc * V3_TILE_N;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (ks >> 2) * V3_KBLOCK_BYTES + (ks & 3) * 32;
            umma_f16(d_tmem, umma_desc(a_addr + off), umma_desc(b_addr + off), idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * stage);   // stage reusable once these MMAs have read it
          umma_commit(bar_tfull + 8 * acc);     // accumulator ready
        }
      }
      __syncwarp();
    } else if (warp >= V3_EPI_WARP0 && warp - V3_EPI_WARP0 < n_epi_warps) {
      // =========================== epilogue ===========================
      const int ew = warp - V3_EPI_WARP0;
      const int q = ew * 32 + lane;
      int T = 0;
      for (int i = 0; i < n_docs; ++i) {
        float m0 = FPB_PAD_SENTINEL, m1 = FPB_PAD_SENTINEL, m2 = FPB_PAD_SENTINEL, m3 = FPB_PAD_SENTINEL;
        while (T < n_tiles && tiles[T].doc == i) {
          const uint32_t g = gtile + T;
          const uint32_t acc = g & 1;
          const int nvalid = tiles[T].nvalid;
          mbar_wait(bar_tfull + 8 * acc, (g >> 1) & 1);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + acc * V3_TILE_N;
          for (int c0 = 0; c0 < nvalid; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
            if (c0 + 32 <= nvalid) {
#pragma unroll
              for (int x = 0; x < 32; x += 4) {
                m0 = fmaxf(m0, v[x]);
                m1 = fmaxf(m1, v[x + 1]);
                m2 = fmaxf(m2, v[x + 2]);
                m3 = fmaxf(m3, v[x + 3]);
              }
            } else {
#pragma unroll
              for (int x = 0; x < 32; ++x)
                if (c0 + x < nvalid) m0 = fmaxf(m0, v[x]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
          ++T;
        }
        // fp16 rounding of the maximum, fp32 sum over the real query tokens
        const float mq = __half2float(__float2half_rn(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3))));
        float s = (q < Q) ? mq : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (n_epi_warps == 1) {
          if (lane == 0) exact[int64_t(b) * R + docs[i].r] = s;
        } else {
          float* part = reinterpret_cast<float*>(&misc[4]);
          if (lane == 0) part[ew] = s;
          asm volatile("bar.sync 1, %0;" ::"r"(n_epi_warps * 32) : "memory");
          if (ew == 0 && lane == 0) {
            float tot = 0.f;
            for (int x = 0; x < n_epi_warps; ++x) tot += part[x];
            exact[int64_t(b) * R + docs[i].r] = tot;
          }
          asm volatile("bar.sync 1, %0;" ::"r"(n_epi_warps * 32) : "memory");
        }
      }
    }
    gtile += uint32_t(n_tiles);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == V3_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(V3_TMEM_COLS));
  }
}

}  // namespace

int launch_maxsim_v3(const fpb_index* ix, const Ws& ws, cudaStream_t st, bool* handled) {
  *handled = false;
  const fpb_layout& L = *ws.L;
  if (ix->dim != 128 || ix->nbits != 4 || L.Qp > 128) return FPB_OK;
  const int n_epi = (L.Qp + 31) / 32;
  const int n_dec = (11 + (4 - n_epi)) & ~1;  // decode slots (even, so a tile is a multiple of 16 tokens): 14/12/12/10
  const int tile_tokens = 8 * n_dec;
  const int64_t tiles_per_doc = (ix->max_doc_len + tile_tokens - 1) / tile_tokens;
  if (tiles_per_doc < 1 || tiles_per_doc > V3_MAX_TILES) return FPB_OK;
  int docs_per_chunk = int(V3_MAX_TILES / tiles_per_doc);
  if (docs_per_chunk > V3_MAX_DOCS) docs_per_chunk = V3_MAX_DOCS;
  *handled = true;
  static bool attr_done = false;
  if (!attr_done) {
    FPB_CUDA_CHECK(cudaFuncSetAttribute(k5_maxsim_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V3Smem::bytes));
    attr_done = true;
  }
  WPerm wp;
  for (int i = 0; i < 16; ++i) wp.v[i] = ix->w_perm_bits[i];
  int* counter = ws.work() + L.B + 3;
  FPB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), st));
  const int chunks = L.B * ((L.R + docs_per_chunk - 1) / docs_per_chunk);
  const int blocks = chunks < ix->sm_count ? chunks : ix->sm_count;
  k5_maxsim_v3_kernel<<<blocks, V3_THREADS, V3Smem::bytes, st>>>(
      ix->centroids, ix->doc_offsets, ix->doc_codes, ix->doc_residuals, wp, ws.queries(), L.Q, L.Qp, L.B, L.R,
      docs_per_chunk, n_dec, ws.n_rerank(), ws.rerank(), ws.exact(), counter);
  FPB_LAUNCH_CHECK("k5_maxsim_v3");
  return FPB_OK;
}
