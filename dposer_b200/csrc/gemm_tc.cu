// Split-fp16 GEMM on tcgen05 for the training step of the score network (reference: the nn.Linear forward / backward
// that autograd runs under lib/algorithms/advanced/losses.py:61-137,187-275 -- y = x W^T + b, dX = dY W, dW = dY^T X).
//
//   C[m, n] = sum_k A[m, k] * B[n, k]  (+ bias1[n] + bias2[n] + add[m, n])
//
// Both operands are K-major fp16 [hi | lo] pairs (x = hi + lo, lo = fp16(x - hi)); three products hi.hi + lo.hi + hi.lo
// accumulate in one fp32 TMEM tile (~1e-6 relative, the scheme every tensor-core kernel of this library uses), so the
// result stands in for the reference's fp32 matmul.  Transposed products are formed from transposed operand copies
// (split_kernel writes both layouts in one pass), which keeps one kernel and one shared-memory layout (SWIZZLE_128B).
//
//   warp 0     TMA producer: per k-block the four tiles A_hi, A_lo, B_hi, B_lo (64 KB stage, 3 stages)
//   warp 1     TMEM allocator + MMA issuer: 12 UMMAs (M = 128, N = 128, K = 16) per stage
//   warps 2-5  epilogue: TMEM -> registers -> 32 x 32 transpose in shared memory -> C (a warp stores whole row segments)
// Persistent: a CTA walks tiles blockIdx.x, + gridDim.x, ... with two TMEM accumulators, so that the epilogue of one tile
// overlaps the main loop of the next (matters once a GEMM has more tiles than the chip has SMs).
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include "score.h"
#include "ptx.cuh"
#include "train.h"

namespace dpb {

int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t rows,
                 uint32_t box_inner, uint32_t box_rows, size_t elem_bytes);  // score_tc.cu

namespace gtc {

constexpr int BK = 64;
constexpr int TILE = 128 * BK * 2;      // 16 KB: [128 rows x 64 k] fp16, SWIZZLE_128B
constexpr int ST = 3;                   // stages of four tiles
constexpr int NUM_THREADS = 192;
constexpr int OFF_TR = ST * 4 * TILE;            // 4 x [32][33] fp32 transpose buffers of the epilogue warps
constexpr int OFF_BAR = OFF_TR + 4 * 32 * 33 * 4 + 512;
constexpr int NBARS = 2 * ST + 4;                // operand stages full / empty, accumulator full / empty x 2
constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;
static_assert(OFF_BAR % 8 == 0 && SMEM_BYTES <= 232448, "shared memory layout");

struct Params {
  int M, N, kt;              // valid output extent, k-blocks of 64 per product
  int a_r0, a_k0, a_lo;      // A operand: first row, first column of the hi half, column distance of the lo half
  int b_r0, b_k0, b_lo;
  float* C;
  int64_t ldc;
  const float* bias1;        // [N] or null
  const float* bias2;        // [N] or null
  const float* add;          // [M, ldadd] or null (may alias C)
  int64_t ldadd;
  int bias_rows;             // the biases apply to rows < bias_rows (stacked [primal ; tangent] rows of the JVP)
  int mt, n_tiles;           // m-tiles, all tiles: a CTA walks tiles blockIdx.x, + gridDim.x, ... (persistent)
};

template <int BN>      // output tile 128 x BN: BN = 64 doubles the CTA count of the small GEMMs (a step at batch 1280 is latency-bound)
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_split_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap tm_a,
                  const __grid_constant__ CUtensorMap tm_b) {
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, BN, 0);
  constexpr uint32_t BTILE = BN * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR;
  auto full = [&](uint32_t s) { return bar + 8u * s; };
  auto empty = [&](uint32_t s) { return bar + 8u * (ST + s); };
  auto dfull = [&](uint32_t b) { return bar + 8u * (2 * ST + b); };
  auto dempty = [&](uint32_t b) { return bar + 8u * (2 * ST + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(dfull(b), 1); ptx::mbar_init(dempty(b), 4); }   // dempty: the four epilogue warps
    ptx::fence_barrier_init();
  }
  // Two accumulators: the epilogue of tile i (TMEM -> transpose -> global) overlaps the main loop of tile i + 1 when a CTA
  // owns several tiles (large batches); with one tile per CTA the second accumulator is simply never used.
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 2 * BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto tile_m0 = [&](int t) { return (t % p.mt) * 128; };
  auto tile_n0 = [&](int t) { return (t / p.mt) * BN; };

  if (warp == 0) {
    if (lane == 0) { ptx::prefetch_tmap(&tm_a); ptx::prefetch_tmap(&tm_b); }
    __syncwarp();
    uint32_t s = 0, ph = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
    const int m0 = tile_m0(t), n0 = tile_n0(t);
    for (int k = 0; k < p.kt; ++k) {
      ptx::mbar_wait(empty(s), ph ^ 1);
      if (ptx::elect_one()) {
        const uint32_t base = sb + s * 4 * TILE;
        ptx::mbar_arrive_expect_tx(full(s), 2 * TILE + 2 * BTILE);
        ptx::tma_load_2d(base, &tm_a, full(s), p.a_k0 + k * BK, p.a_r0 + m0);
        ptx::tma_load_2d(base + TILE, &tm_a, full(s), p.a_k0 + p.a_lo + k * BK, p.a_r0 + m0);
        ptx::tma_load_2d(base + 2 * TILE, &tm_b, full(s), p.b_k0 + k * BK, p.b_r0 + n0);
        ptx::tma_load_2d(base + 3 * TILE, &tm_b, full(s), p.b_k0 + p.b_lo + k * BK, p.b_r0 + n0);
      }
      __syncwarp();
      if (++s == ST) { s = 0; ph ^= 1; }
    }
    }
  } else if (warp == 1) {
    uint32_t s = 0, ph = 0, it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
    const uint32_t buf = it & 1, acc = tmem_base + buf * BN;
    ptx::mbar_wait(dempty(buf), ((it >> 1) & 1) ^ 1);          // the epilogue has read this accumulator's previous tile
    ptx::tc_fence_after();
    for (int k = 0; k < p.kt; ++k) {
      ptx::mbar_wait(full(s), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t ahi = ptx::umma_desc_sw128(sb + s * 4 * TILE), alo = ahi + (uint64_t)(TILE >> 4);
        const uint64_t bhi = alo + (uint64_t)(TILE >> 4), blo = bhi + (uint64_t)(TILE >> 4);
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          ptx::mma_f16_ss(acc, alo + 2 * j, bhi + 2 * j, IDESC, (k == 0 && j == 0) ? 0u : 1u);
          ptx::mma_f16_ss(acc, ahi + 2 * j, blo + 2 * j, IDESC, 1u);
          ptx::mma_f16_ss(acc, ahi + 2 * j, bhi + 2 * j, IDESC, 1u);
        }
        ptx::mma_commit(empty(s));
      }
      __syncwarp();
      if (++s == ST) { s = 0; ph ^= 1; }
    }
    if (ptx::elect_one()) ptx::mma_commit(dfull(buf));
    __syncwarp();
    }
  } else {
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    // Each 32 x 32 block is transposed through shared memory (a buffer of its own: the operand stages already hold the next
    // tile) so that a warp writes
    // 128 contiguous bytes of one output row per instruction.  Everything the epilogue reads from global memory (bias,
    // the added matrix) is requested BEFORE the accumulator is waited for / loaded: the loads of a block are independent.
    float* tr = reinterpret_cast<float*>(smem + OFF_TR) + q * (32 * 33);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
    const int m0 = tile_m0(t), n0 = tile_n0(t);
    const uint32_t buf = it & 1, acc = tmem_base + buf * BN;
    const int mrow0 = m0 + q * 32;
    const int rmax = min(32, p.M - mrow0);
    float bsum[BN / 32];
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      const int n = n0 + c * 32 + lane;
      bsum[c] = 0.f;
      if (n < p.N) {
        if (p.bias1) bsum[c] += p.bias1[n];
        if (p.bias2) bsum[c] += p.bias2[n];
      }
    }
    auto load_add = [&](int c, float (&a)[32]) {
      const int n = n0 + c * 32 + lane;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        a[r] = (p.add && n < p.N && r < rmax) ? p.add[(size_t)(mrow0 + r) * p.ldadd + n] : 0.f;
    };
    float a[32];
    load_add(0, a);
    ptx::mbar_wait(dfull(buf), (it >> 1) & 1);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(acc + ((uint32_t)(q * 32) << 16) + c * 32, v);
      ptx::tmem_ld_wait();
      if (c == BN / 32 - 1) {                  // this warp's last read of the accumulator: hand it back to the issuer
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(dempty(buf));
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 32; ++i) tr[lane * 33 + i] = __uint_as_float(v[i]);
      __syncwarp();
      const int n = n0 + c * 32 + lane;
      float o[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) o[r] = tr[r * 33 + lane] + (mrow0 + r < p.bias_rows ? bsum[c] : 0.f) + a[r];
      if (c + 1 < BN / 32) load_add(c + 1, a);            // next block's loads fly while this one is stored
      if (n < p.N) {
#pragma unroll
        for (int r = 0; r < 32; ++r)
          if (r < rmax) p.C[(size_t)(mrow0 + r) * p.ldc + n] = o[r];
      }
    }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * BN);
  }
}

// src fp32 [R, C] (row stride ld)  ->  row form  dst[(r0 + r), c0 + c | c0 + lo + c]   (stride dld, if dst)
//                                      col form  dstT[(tr0 + c), tc0 + r | tc0 + tlo + r] (stride tld, if dstT)
__global__ void split_kernel(const float* __restrict__ src, int R, int Cc, int64_t ld, __half* __restrict__ dst,
                             int64_t dld, int r0, int c0, int lo, __half* __restrict__ dstT, int64_t tld, int tr0,
                             int tc0, int tlo) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;           // bx: column block, by: row block
  const int tx = threadIdx.x, ty = threadIdx.y;                   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = by + i, c = bx + tx;
    float x = 0.f;
    if (r < R && c < Cc) {
      x = src[(size_t)r * ld + c];
      if (dst) {
        const __half hi = __float2half_rn(x);
        __half* d = dst + (size_t)(r0 + r) * dld + c0 + c;
        d[0] = hi;
        d[lo] = __float2half_rn(x - __half2float(hi));
      }
    }
    tile[i][tx] = x;
  }
  if (!dstT) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = bx + i, r = by + tx;
    if (r < R && c < Cc) {
      const float x = tile[tx][i];
      const __half hi = __float2half_rn(x);
      __half* d = dstT + (size_t)(tr0 + c) * tld + tc0 + r;
      d[0] = hi;
      d[tlo] = __float2half_rn(x - __half2float(hi));
    }
  }
}

}  // namespace gtc

int gemm_tc_init() {
  DPB_CUDA_CHECK(cudaFuncSetAttribute(gtc::gemm_split_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc::SMEM_BYTES));
  DPB_CUDA_CHECK(cudaFuncSetAttribute(gtc::gemm_split_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc::SMEM_BYTES));
  return DPB_OK;
}

int gemm_tc(const Op16& A, const Op16& B, int M, int N, int K, float* C, int64_t ldc, const float* bias1,
            const float* bias2, const float* add, int64_t ldadd, cudaStream_t st, int bias_rows) {
  if (M <= 0 || N <= 0) return DPB_OK;
  CUtensorMap ta, tb;
  const int mt = (M + 127) / 128;
  const int bn = (mt * ((N + 127) / 128) < 48) ? 64 : 128;     // under a third of a wave of 128 x 128 tiles: halve the tile
  // (measured: at 80 CTAs the 128-wide tile is as fast -- the kernel is bound by the bytes each SM pulls from L2 per k-block)
  int rc = make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, A.ptr, (uint64_t)A.ld, (uint64_t)A.rows, gtc::BK, 128, 2);
  if (rc != DPB_OK) return rc;
  rc = make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, B.ptr, (uint64_t)B.ld, (uint64_t)B.rows, gtc::BK, (uint32_t)bn, 2);
  if (rc != DPB_OK) return rc;
  gtc::Params p{};
  p.M = M; p.N = N; p.kt = (K + gtc::BK - 1) / gtc::BK;
  p.a_r0 = A.r0; p.a_k0 = A.k0; p.a_lo = A.lo;
  p.b_r0 = B.r0; p.b_k0 = B.k0; p.b_lo = B.lo;
  p.C = C; p.ldc = ldc; p.bias1 = bias1; p.bias2 = bias2; p.add = add; p.ldadd = ldadd;
  p.bias_rows = bias_rows;
  p.mt = mt;
  p.n_tiles = mt * ((N + bn - 1) / bn);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  dim3 grid((unsigned)(p.n_tiles < sm_count ? p.n_tiles : sm_count));
  if (bn == 64)
    gtc::gemm_split_kernel<64><<<grid, gtc::NUM_THREADS, gtc::SMEM_BYTES, st>>>(p, ta, tb);
  else
    gtc::gemm_split_kernel<128><<<grid, gtc::NUM_THREADS, gtc::SMEM_BYTES, st>>>(p, ta, tb);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

int split16(const float* src, int R, int Cc, int64_t ld, const Op16* row, const Op16* col, cudaStream_t st) {
  if (R <= 0 || Cc <= 0) return DPB_OK;
  dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((R + 31) / 32)), block(32, 8);
  gtc::split_kernel<<<grid, block, 0, st>>>(src, R, Cc, ld, row ? row->ptr : nullptr, row ? row->ld : 0, row ? row->r0 : 0,
                                            row ? row->k0 : 0, row ? row->lo : 0, col ? col->ptr : nullptr,
                                            col ? col->ld : 0, col ? col->r0 : 0, col ? col->k0 : 0, col ? col->lo : 0);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

}  // namespace dpb
