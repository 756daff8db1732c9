// K5 v2 : fused residual decompression + exact MaxSim for (dim=128, nbits=4, Qp in {32,64}).
// Superseded as the default by v4 (Qp <= 32) and v5 (Qp > 32); still the path for Qp = 64 indexes whose
// longest document does not fit v5's pass table, and the A/B reference (FPB_K5=v2).
// Same arithmetic, bit for bit, as k5_maxsim.cu (v1, kept for the other shapes); what changed
// is the data movement, guided by the round-1 ncu capture (profiles/r01_summary.md: v1 was
// bound by L1/shared wavefronts at 75 % and issued only 35 % of its slots):
//
//  * warp-autonomous: one warp owns one document and a private 16-row A tile, so the steady
//    state has no block barrier (v1: two __syncthreads per 64-token tile);
//  * the 4 lanes of a token read the centroid row as 4 x 64 contiguous bytes per load
//    (lane j owns 16-byte chunks j, j+4, j+8, j+12) instead of one 64-byte slice per lane:
//    4 L1 wavefronts per token instead of 8;
//  * the 256-entry half2 LUT is replicated once per bank (32 KB): a lookup is always one
//    conflict-free wavefront (v1: ~3.5-way conflicts);
//  * token->row permutation inside a pass makes the 16-byte A-tile stores conflict-free;
//  * loads of pass p+1 (codes two passes ahead) are issued before pass p is decoded;
//  * CTAs pull chunks of 32 documents of ONE query from a dynamic queue, so the query tile
//    is loaded once per chunk.
#include "kernels.h"

namespace {

constexpr int V2_THREADS = 256;
constexpr int V2_WARPS = 8;
constexpr int V2_CHUNK = 32;  // documents per work item (4 per warp)
constexpr int V2_D = 128;
constexpr int V2_LDS = V2_D + 8;  // halves; 272-byte rows: conflict-free ldmatrix

template <int QP>
struct V2Smem {
  static constexpr int q_bytes = QP * V2_LDS * 2;
  static constexpr int a_bytes = V2_WARPS * 16 * V2_LDS * 2;
  static constexpr int lut_bytes = 256 * 32 * 4;
  static constexpr int bytes = q_bytes + a_bytes + lut_bytes;
};

struct Raw {
  uint32_t w[4];  // residual words j, j+4, j+8, j+12 of the token
  uint4 c[4];     // centroid chunks j, j+4, j+8, j+12 (8 halves each)
};

__device__ __forceinline__ float v2_div_rn(float e, float n, float r) {
  const float q = __fmul_rn(e, r);
  const float rem = __fmaf_rn(-q, n, e);
  return __fmaf_rn(rem, r, q);
}

__device__ __forceinline__ void load_raw(Raw& raw, const uint8_t* __restrict__ residuals,
                                         const __half* __restrict__ C, int64_t tok_global, int code, int j) {
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(residuals + tok_global * 64) + j;
  const uint4* cc = reinterpret_cast<const uint4*>(C + int64_t(code) * V2_D) + j;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    raw.w[k] = __ldg(rw + 4 * k);
    raw.c[k] = __ldg(cc + 4 * k);
  }
}

template <int QP>
__global__ void __launch_bounds__(V2_THREADS, 2)
k5_maxsim_v2_kernel(const __half* __restrict__ C, const int64_t* __restrict__ doc_offsets,
                    const int32_t* __restrict__ codes, const uint8_t* __restrict__ residuals, WPerm wp,
                    const __half* __restrict__ Qpad, int Q, int B, int R, const int32_t* __restrict__ n_rerank,
                    const int32_t* __restrict__ rerank, float* __restrict__ exact, int* __restrict__ counter) {
  constexpr int LDS = V2_LDS;
  constexpr int NT = QP / 8;
  constexpr int KS = V2_D / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* Qs = reinterpret_cast<__half*>(smem_raw);
  __half* As = reinterpret_cast<__half*>(smem_raw + V2Smem<QP>::q_bytes);
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + V2Smem<QP>::q_bytes + V2Smem<QP>::a_bytes);
  __shared__ int s_chunk;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = lane & 3, tslot = lane >> 2;
  const int prow = (tslot >> 1) + 4 * (tslot & 1);  // row of this lane's token inside a pass
  const int g = lane >> 2, t = lane & 3;            // mma fragment coordinates
  __half* Aw = As + warp * 16 * LDS;

  // bank-replicated LUT: entry for byte v and lane l lives at word v*32 + l
  for (int i = tid; i < 256 * 32; i += V2_THREADS) {
    const int v = i >> 5;
    lut[i] = uint32_t(wp.v[v >> 4]) | (uint32_t(wp.v[v & 15]) << 16);
  }
  const uint32_t lut_lane = smem_u32(lut) + lane * 4;
  const __half2 sentinel = __float2half2_rn(FPB_PAD_SENTINEL);
  const int chunks_per_query = (R + V2_CHUNK - 1) / V2_CHUNK;
  const int total_chunks = B * chunks_per_query;
  int cur_b = -1;

  for (;;) {
    __syncthreads();  // everybody is done with the previous chunk (and with the LUT build)
    if (tid == 0) s_chunk = atomicAdd(counter, 1);
    __syncthreads();
    const int chunk = s_chunk;
    if (chunk >= total_chunks) break;
    const int b = chunk / chunks_per_query;
    const int r0 = (chunk % chunks_per_query) * V2_CHUNK;
    const int nr = n_rerank[b];
    if (r0 >= nr) continue;
    if (b != cur_b) {
      for (int i = tid; i < QP * (V2_D / 8); i += V2_THREADS) {
        const int n = i / (V2_D / 8), c8 = i % (V2_D / 8);
        *reinterpret_cast<uint4*>(Qs + n * LDS + c8 * 8) =
            *reinterpret_cast<const uint4*>(Qpad + (int64_t(b) * QP + n) * V2_D + c8 * 8);
      }
      cur_b = b;
      __syncthreads();
    }

    for (int di = 0; di < V2_CHUNK / V2_WARPS; ++di) {
      const int r = r0 + di * V2_WARPS + warp;
      if (r >= nr) break;
      const int d = rerank[int64_t(b) * R + r];
      const int64_t o0 = doc_offsets[d];
      const int len = int(doc_offsets[d + 1] - o0);
      __half2 mx[NT];
#pragma unroll
      for (int i = 0; i < NT; ++i) mx[i] = sentinel;

      if (len > 0) {
        const int npass = (len + 7) >> 3;
        const int last = len - 1;
        int code_cur = __ldg(codes + o0 + min(prow, last));
        int code_nxt = __ldg(codes + o0 + min(8 + prow, last));
        Raw cur;
        load_raw(cur, residuals, C, o0 + min(prow, last), code_cur, j);

        for (int p = 0; p < npass; ++p) {
          Raw nxt;
          int code_nn = 0;
          if (p + 1 < npass) load_raw(nxt, residuals, C, o0 + min((p + 1) * 8 + prow, last), code_nxt, j);
          if (p + 2 < npass) code_nn = __ldg(codes + o0 + min((p + 2) * 8 + prow, last));

          // ---- decode the lane's 32 elements: e = fp16(w_perm[nibble] + centroid) ----
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
          __half* arow = Aw + ((p & 1) * 8 + prow) * LDS;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(e[k * 4 + i]);
              o[i] = pack_half2_rn(v2_div_rn(f.x, nf, rcp), v2_div_rn(f.y, nf, rcp));
            }
            *reinterpret_cast<uint4*>(arow + 8 * (j + 4 * k)) = make_uint4(o[0], o[1], o[2], o[3]);
          }

          // ---- every second pass (or at the end): ts = A(16 x 128) . Q^T ----
          if ((p & 1) || p == npass - 1) {
            __syncwarp();
            const int t0 = (p & ~1) * 8;
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              uint32_t a[4];
              ldmatrix_x4(a[0], a[1], a[2], a[3], smem_u32(Aw + (lane & 15) * LDS + ks * 16 + (lane >> 4) * 8));
#pragma unroll
              for (int nt = 0; nt < NT; nt += 2) {
                const int mat = lane >> 3;
                const int n = (nt + (mat >> 1)) * 8 + (lane & 7);
                const int kk = ks * 16 + (mat & 1) * 8;
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3, smem_u32(Qs + n * LDS + kk));
                mma_16816(acc[nt], a, b0, b1);
                mma_16816(acc[nt + 1], a, b2, b3);
              }
            }
            const bool v0 = (t0 + g) < len;
            const bool v1 = (t0 + g + 8) < len;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const __half2 h0 = v0 ? __floats2half2_rn(acc[nt][0], acc[nt][1]) : sentinel;
              const __half2 h1 = v1 ? __floats2half2_rn(acc[nt][2], acc[nt][3]) : sentinel;
              mx[nt] = __hmax2(mx[nt], __hmax2(h0, h1));
            }
            __syncwarp();  // the A tile is rewritten by the next pass
          }
          cur = nxt;
          code_cur = code_nxt;
          code_nxt = code_nn;
        }
      }
      // ---- column maxima over the 8 row groups, fp32 sum over the real query tokens ----
      float s = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        __half2 m = mx[nt];
        m = __hmax2(m, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m), 4)));
        m = __hmax2(m, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m), 8)));
        m = __hmax2(m, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(m), 16)));
        const float2 f = __half22float2(m);
        const int col = nt * 8 + 2 * t;
        if (col < Q) s += f.x;
        if (col + 1 < Q) s += f.y;
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (lane == 0) exact[int64_t(b) * R + r] = s;
    }
  }
}

template <int QP>
int launch_v2_t(const fpb_index* ix, const Ws& ws, cudaStream_t st) {
  const fpb_layout& L = *ws.L;
  auto kern = k5_maxsim_v2_kernel<QP>;
  constexpr int smem = V2Smem<QP>::bytes;
  // opt in on every launch: the attribute is per device and the call costs about a microsecond
  FPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  WPerm wp;
  for (int i = 0; i < 16; ++i) wp.v[i] = ix->w_perm_bits[i];
  int* counter = ws.work() + L.B + 2;
  FPB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), st));
  const int chunks = L.B * ((L.R + V2_CHUNK - 1) / V2_CHUNK);
  const int blocks = chunks < ix->sm_count * 2 ? chunks : ix->sm_count * 2;
  kern<<<blocks, V2_THREADS, smem, st>>>(ix->centroids, ix->doc_offsets, ix->doc_codes, ix->doc_residuals, wp,
                                         ws.queries(), L.Q, L.B, L.R, ws.n_rerank(), ws.rerank(), ws.exact(),
                                         counter);
  FPB_LAUNCH_CHECK("k5_maxsim_v2");
  return FPB_OK;
}

}  // namespace

// Returns FPB_ERR_UNSUPPORTED (without setting an error) when the shape is not covered, so the
// caller falls through to the generic v1 kernel.
int launch_maxsim_v2(const fpb_index* ix, const Ws& ws, cudaStream_t st, bool* handled) {
  *handled = false;
  if (ix->dim != 128 || ix->nbits != 4) return FPB_OK;
  if (ws.L->Qp == 32) {
    *handled = true;
    return launch_v2_t<32>(ix, ws, st);
  }
  if (ws.L->Qp == 64) {
    *handled = true;
    return launch_v2_t<64>(ix, ws, st);
  }
  return FPB_OK;
}
