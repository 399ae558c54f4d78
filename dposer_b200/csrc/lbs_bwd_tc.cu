// LBS backward, the transposed blend on tcgen05:  dL/d[feat | beta] [B, P+S] = g_vposed [B, 3V] . [posedirs ; shapedirs^T]^T
// -- the adjoint of the blend-shape step of smplx 0.1.28 lbs() (reached from the fitting loops
// run/motion_denoising.py:255-268 through autograd in the reference).  It is the largest contraction of the backward
// (62 GFLOP per 1920 SMPL-X poses) and ran as an fp32 FFMA SGEMM in round 1 (bwd_gemm_kernel, 14.4 ms per 15 360 poses).
//
//   D[k, b] = sum_r basisT[k, r] * g[b, r]        M = features (tiles of 128 rows), N = 128 poses, K = r = 3v + c
//
// Both operands are K-major fp16 [hi | lo] pairs (basisT built once per model, g written by the skinning adjoint), three
// products hi.hi + hi.lo + lo.hi with fp32 accumulation in TMEM (~1e-6 relative, same scheme as the forward).  One CTA
// owns 128 poses x ALL feature tiles (MT x 128 TMEM columns), so every g slab is read once; the r range can be split
// across blockIdx.y for small batches (partials are summed in a fixed order by the unpack kernel: deterministic).
//
//   warp 0  TMA producer: g slabs (hi + lo, double buffered) and basis tiles (ring)
//   warp 1  TMEM allocator + MMA issuer
//   warps 2-5  epilogue: TMEM -> out[split, b, k]   (lane = feature row: 32 consecutive k per store -> full lines)
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include <vector>

#include "lbs.h"
#include "ptx.cuh"

namespace dpb {

int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t rows,
                 uint32_t box_inner, uint32_t box_rows, size_t elem_bytes);  // score_tc.cu

namespace lbt {

constexpr int BK = 64;
constexpr int TILE = 128 * BK * 2;      // 16 KB: [128 rows x 64 k] fp16, SWIZZLE_128B
constexpr int BST = 2;                  // g stages (hi + lo each)
constexpr int AST = 6;                  // basis-tile stages
constexpr int NUM_THREADS = 192;
constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + BST * 2 * TILE;
constexpr int OFF_BAR = OFF_A + AST * TILE;
constexpr int NBARS = 2 * BST + 2 * AST + 1;
constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;

struct Params {
  int Rp;            // columns of one half (3V padded to 64)
  int n_slabs;       // Rp / 64
  int Kp;            // feature rows (MT * 128)
  int splits;
  int64_t B;
  float* out;        // [splits, B, Kp]
};

template <int MT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lbs_blendT_tc_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap tm_bT,
                     const __grid_constant__ CUtensorMap tm_g) {
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, 128, 0);
  constexpr uint32_t TCOLS = MT * 128 <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t b_base = sb + OFF_B, a_base = sb + OFF_A, bar = sb + OFF_BAR;
  auto bfull = [&](uint32_t s) { return bar + 8u * s; };
  auto bempty = [&](uint32_t s) { return bar + 8u * (BST + s); };
  auto afull = [&](uint32_t s) { return bar + 8u * (2 * BST + s); };
  auto aempty = [&](uint32_t s) { return bar + 8u * (2 * BST + AST + s); };
  const uint32_t dfull = bar + 8u * (2 * BST + 2 * AST);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < BST; ++s) { ptx::mbar_init(bfull(s), 1); ptx::mbar_init(bempty(s), 1); }
    for (int s = 0; s < AST; ++s) { ptx::mbar_init(afull(s), 1); ptx::mbar_init(aempty(s), 1); }
    ptx::mbar_init(dfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), TCOLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n0 = blockIdx.x * 128;                                   // first pose of this CTA
  const int s0 = (int)((long long)p.n_slabs * blockIdx.y / p.splits);
  const int s1 = (int)((long long)p.n_slabs * (blockIdx.y + 1) / p.splits);

  if (warp == 0) {
    if (lane == 0) { ptx::prefetch_tmap(&tm_bT); ptx::prefetch_tmap(&tm_g); }
    __syncwarp();
    uint32_t bs = 0, bph = 0, as = 0, aph = 0;
    for (int s = s0; s < s1; ++s) {
      ptx::mbar_wait(bempty(bs), bph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(bfull(bs), 2 * TILE);
        ptx::tma_load_2d(b_base + (bs * 2) * TILE, &tm_g, bfull(bs), s * BK, n0);
        ptx::tma_load_2d(b_base + (bs * 2 + 1) * TILE, &tm_g, bfull(bs), p.Rp + s * BK, n0);
      }
      __syncwarp();
      if (++bs == BST) { bs = 0; bph ^= 1; }
#pragma unroll 1
      for (int t = 0; t < 2 * MT; ++t) {                               // tile t: feature tile t / 2, half t % 2
        ptx::mbar_wait(aempty(as), aph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(afull(as), TILE);
          ptx::tma_load_2d(a_base + as * TILE, &tm_bT, afull(as), (t & 1) * p.Rp + s * BK, (t >> 1) * 128);
        }
        __syncwarp();
        if (++as == AST) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 1) {
    uint32_t bs = 0, bph = 0, as = 0, aph = 0;
    for (int s = s0; s < s1; ++s) {
      ptx::mbar_wait(bfull(bs), bph);
      ptx::tc_fence_after();
      const uint64_t bhi = ptx::umma_desc_sw128(b_base + (bs * 2) * TILE), blo = bhi + (uint64_t)(TILE >> 4);
      const uint32_t first = (s == s0) ? 0u : 1u;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const uint32_t taddr = tmem_base + m * 128;
        // basis_hi x (g_hi + g_lo)
        ptx::mbar_wait(afull(as), aph);
        ptx::tc_fence_after();
        {
          const uint64_t ad = ptx::umma_desc_sw128(a_base + as * TILE);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              ptx::mma_f16_ss(taddr, ad + 2 * j, bhi + 2 * j, IDESC, j == 0 ? first : 1u);
              ptx::mma_f16_ss(taddr, ad + 2 * j, blo + 2 * j, IDESC, 1u);
            }
            ptx::mma_commit(aempty(as));
          }
          __syncwarp();
          if (++as == AST) { as = 0; aph ^= 1; }
        }
        // basis_lo x g_hi
        ptx::mbar_wait(afull(as), aph);
        ptx::tc_fence_after();
        {
          const uint64_t ad = ptx::umma_desc_sw128(a_base + as * TILE);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) ptx::mma_f16_ss(taddr, ad + 2 * j, bhi + 2 * j, IDESC, 1u);
            ptx::mma_commit(aempty(as));
            if (m == MT - 1) ptx::mma_commit(bempty(bs));
          }
          __syncwarp();
          if (++as == AST) { as = 0; aph ^= 1; }
        }
      }
      if (++bs == BST) { bs = 0; bph ^= 1; }
    }
    if (ptx::elect_one()) ptx::mma_commit(dfull);
    __syncwarp();
  } else {
    // ---- epilogue: warp % 4 = TMEM lane quarter, lane = feature row of the tile
    const int q = warp & 3;
    ptx::mbar_wait(dfull, 0);
    ptx::tc_fence_after();
    float* out = p.out + (size_t)blockIdx.y * p.B * p.Kp;
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int k = m * 128 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + m * 128 + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int64_t b = (int64_t)n0 + c * 32 + i;
          if (b < p.B) out[(size_t)b * p.Kp + k] = __uint_as_float(v[i]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TCOLS);
  }
}

// gfeat[b, k] += sum_s C[s, b, k] (k < P);  gbeta[b, k - P] += ... (P <= k < P + S): splits added in a fixed order
__global__ void __launch_bounds__(256) blendT_unpack_kernel(const float* __restrict__ C, int splits, int Kp, int P, int S,
                                                            const float* __restrict__ scale, float* __restrict__ gfeat,
                                                            float* __restrict__ gbt, int64_t B) {
  const int64_t b = blockIdx.x;                         // one pose row per block: no per-thread 64-bit division
  const float inv = scale ? 1.0f / scale[b] : 1.0f;     // the operand was scaled into fp16 range by a power of two (exact)
  for (int k = threadIdx.x; k < P + S; k += 256) {
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += C[((size_t)s * B + b) * Kp + k];
    v *= inv;
    if (k < P) gfeat[b * P + k] += v;
    else gbt[b * (S + 3) + (k - P)] += v;
  }
}

}  // namespace lbt

// basisT16 [Kp, 2*Rp]: row k < P = posedirs[k, :], row P + s = shapedirs[:, s]; columns r = 3v + c as fp16 hi | lo
int lbs_bwd_tc_prepare(dpb_lbs* h, const dpb_body_tensors* m) {
  const int V = h->V, P = h->P, S = h->S;
  h->bt_rp = (3 * V + 63) / 64 * 64;
  h->bt_kp = (P + S <= 256) ? 256 : 512;                // feature tiles of 128 rows: 2 (SMPL) or 4 (SMPL-X) per CTA
  if (P + S > 512) return DPB_OK;                       // more feature rows than one CTA's TMEM holds: SGEMM path stays
  const size_t ld = (size_t)2 * h->bt_rp;
  std::vector<__half> buf((size_t)h->bt_kp * ld, __float2half_rn(0.f));
  auto put = [&](int k, int r, float x) {
    const __half hi = __float2half_rn(x);
    buf[(size_t)k * ld + r] = hi;
    buf[(size_t)k * ld + h->bt_rp + r] = __float2half_rn(x - __half2float(hi));
  };
  for (int k = 0; k < P; ++k)
    for (int r = 0; r < 3 * V; ++r) put(k, r, m->posedirs[(size_t)k * 3 * V + r]);
  for (int r = 0; r < 3 * V; ++r)
    for (int s = 0; s < S; ++s) put(P + s, r, m->shapedirs[(size_t)r * S + s]);
  DPB_CUDA_CHECK(cudaMalloc((void**)&h->basisT16, buf.size() * sizeof(__half)));
  DPB_CUDA_CHECK(cudaMemcpy(h->basisT16, buf.data(), buf.size() * sizeof(__half), cudaMemcpyHostToDevice));
  int rc = make_tmap_2d(&h->tm_bT, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, h->basisT16, ld, (uint64_t)h->bt_kp, lbt::BK, 128, 2);
  if (rc != DPB_OK) return rc;
  DPB_CUDA_CHECK(cudaFuncSetAttribute(lbt::lbs_blendT_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      lbt::SMEM_BYTES));
  DPB_CUDA_CHECK(cudaFuncSetAttribute(lbt::lbs_blendT_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      lbt::SMEM_BYTES));
  h->bt_ready = true;
  return DPB_OK;
}

void lbs_bwd_tc_release(dpb_lbs* h) {
  if (h->basisT16) cudaFree(h->basisT16);
  h->basisT16 = nullptr;
  h->bt_ready = false;
}

int lbs_blendT_splits(const dpb_lbs* h, int64_t B) {
  const int64_t nt = (B + 127) / 128;
  int64_t s = h->sm_count / (nt < 1 ? 1 : nt);
  if (s > 8) s = 8;
  if (s > h->bt_rp / lbt::BK) s = h->bt_rp / lbt::BK;   // every split owns at least one slab (an empty one would drain an
  if (s < 1) s = 1;                                     // accumulator that no MMA ever wrote)
  return (int)s;
}

// gvp16 [B, 2*Rp] fp16 [hi | lo] (pad columns zero; row b scaled by scale[b] if given) -> gfeat [B,P] += ..., gbeta [B,S+3][:S] += ...; cpart [splits,B,Kp]
int lbs_blendT_tc(dpb_lbs* h, const __half* gvp16, const float* scale, float* cpart, float* gfeat, float* gbeta,
                  int64_t B, cudaStream_t st) {
  CUtensorMap tm_g;
  int rc = make_tmap_2d(&tm_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, gvp16, (uint64_t)2 * h->bt_rp, (uint64_t)B, lbt::BK, 128, 2);
  if (rc != DPB_OK) return rc;
  lbt::Params p{};
  p.Rp = h->bt_rp;
  p.n_slabs = h->bt_rp / lbt::BK;
  p.Kp = h->bt_kp;
  p.splits = lbs_blendT_splits(h, B);
  p.B = B;
  p.out = cpart;
  dim3 grid((unsigned)((B + 127) / 128), (unsigned)p.splits);
  if (h->bt_kp <= 256)
    lbt::lbs_blendT_tc_kernel<2><<<grid, lbt::NUM_THREADS, lbt::SMEM_BYTES, st>>>(p, h->tm_bT, tm_g);
  else
    lbt::lbs_blendT_tc_kernel<4><<<grid, lbt::NUM_THREADS, lbt::SMEM_BYTES, st>>>(p, h->tm_bT, tm_g);
  const int64_t n = B * (h->P + h->S);
  lbt::blendT_unpack_kernel<<<(unsigned)B, 256, 0, st>>>(cpart, p.splits, p.Kp, h->P, h->S, scale, gfeat, gbeta, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

}  // namespace dpb
