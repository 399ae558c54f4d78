// LBS backward, the skinning adjoint with the joint-transform cotangent on tcgen05 (deterministic, no atomics).
// Adjoint of  v_out = sum_j w[v,j] (R_j v_posed + t_j)  (smplx 0.1.28 lbs()) for dense vertex
// cotangents g [B,V,3] (run/motion_denoising.py:255-268 differentiates through it):
//
//   g_vposed[b,v,:] = (sum_j w[v,j] R_j[b])^T g[b,v,:]      ltc::lbs_skin_tc_kernel in adjoint mode (lbs_tc.cu): T = W A on the
//                                                            tensor cores as in the forward, thread = vertex applies T_R^T
//   dL/dA[b,j,(x,c)] = sum_v w[v,j] g[b,v,x] [v_posed[b,v,c] | 1]                       this kernel, tensor cores:
//       D[j, (pose,e)] = sum_v W^T[j,v] * Q[(pose,e), v],   Q = g (x) [v_posed | 1]  (12 entries per vertex and pose)
//
// Round 1 accumulated dL/dA with 48 shared-memory atomics per (vertex, pose) plus global atomics (15.5 ms per 15 360
// SMPL-X poses, order-dependent sums).  Here a CTA owns 16 poses and sweeps the vertices in slabs of 64: sixteen compute
// warps form Q for the slab and write it -- fp16 [hi | lo], SWIZZLE_128B K-major -- straight into shared memory as the
// B operand (generic-proxy stores + fence.proxy.async), an issuing warp accumulates W^T . Q over all slabs in 192 TMEM
// columns (hi.hi + hi.lo + lo.hi), and the epilogue writes dL/dA once.  Row J of W^T is all ones, so D[J, 9..11] is the
// translation cotangent sum_v g.  (The first version also rebuilt T_R per vertex and pose from the sparse weights with a
// shared-memory gather of the joint transforms: 673 M shared-memory wavefronts per launch, 60 % of the LSU pipe,
// profiles/r2_ncu_skin_bwd_tc.md -- that half now runs on the tensor cores in lbs_tc.cu.)
//
//   warp 0  TMA: W^T slabs (hi + lo, double buffered);  warp 1  TMEM allocator + MMA issuer
//   warps 2-17  compute (thread = vertex of the slab x 2 poses);  warps 2-5 also drain the accumulator at the end
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <vector>

#include "lbs.h"
#include "ptx.cuh"

namespace dpb {

int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t rows,
                 uint32_t box_inner, uint32_t box_rows, size_t elem_bytes);  // score_tc.cu

namespace lsb {

__global__ void lbs_rowscale_kernel(const float* __restrict__ g, int64_t n, const float* __restrict__ g2, int64_t n2,
                                    float* __restrict__ scale);

constexpr int BK = 64;                       // vertices per slab = K of one operand tile
constexpr int NPG = 16;                      // poses per CTA
constexpr int NQ = NPG * 12;                 // 192 = N of the MMAs
constexpr int Q_TILE = NQ * BK * 2;          // 24 KB: [192 rows x 64 k] fp16, SWIZZLE_128B
constexpr int W_TILE = 128 * BK * 2;         // 16 KB: [128 joint rows x 64 k]
constexpr int CW = 16;                       // compute warps: 4 per scheduler (8 ran one dependent instruction at a time)
constexpr int NUM_THREADS = 64 + CW * 32;    // 576
constexpr int OFF_Q = 0;                     // 2 buffers x (hi, lo)
constexpr int OFF_W = OFF_Q + 4 * Q_TILE;    // 2 stages x (hi, lo)
constexpr int OFF_BAR = OFF_W + 4 * W_TILE;
constexpr int NBARS = 9;
constexpr int OFF_A = OFF_BAR + NBARS * 8 + 16;   // fp32 [NPG][J][12] skinning transforms of the CTA's poses
static_assert(OFF_W % 1024 == 0 && OFF_BAR % 1024 == 0, "operand tiles are 1024-byte aligned");

struct Params {
  const float* vposed;       // [B,V,3]
  const float* g_verts;      // [B,V,3]
  const float* gextra;       // [B,n_need,3] or nullptr
  const int32_t* need_index; // [V] or nullptr
  float* gA;                 // [B,J,12]
  float* gbt;                // [B,S+3]
  const float* scale;        // [B] power-of-two scale of each pose's cotangents (fp16 range), see lbs_rowscale_kernel
  int V, J, S, n_need, Vp;   // Vp = V padded to 64 (columns of one half of W^T)
  int64_t B;
};

// byte offset of element (row r, column k) in a K-major SWIZZLE_128B tile of 64 fp16 columns
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
lbs_skin_bwd_tc_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap tm_wT) {
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, NQ, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t q_base = sb + OFF_Q, w_base = sb + OFF_W, bar = sb + OFF_BAR;
  auto qfull = [&](uint32_t b) { return bar + 8u * b; };
  auto qempty = [&](uint32_t b) { return bar + 8u * (2 + b); };
  auto wfull = [&](uint32_t s) { return bar + 8u * (4 + s); };
  auto wempty = [&](uint32_t s) { return bar + 8u * (6 + s); };
  const uint32_t dfull = bar + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBARS * 8);
  float* sc_s = reinterpret_cast<float*>(smem + OFF_A);   // [NPG] scale, [NPG] 1 / scale
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = p.J, V = p.V;
  const int64_t b0 = (int64_t)blockIdx.x * NPG;
  const int np = (int)min((int64_t)NPG, p.B - b0);
  const int n_slabs = p.Vp / BK;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(qfull(b), CW); ptx::mbar_init(qempty(b), 1);
      ptx::mbar_init(wfull(b), 1); ptx::mbar_init(wempty(b), 1);
    }
    ptx::mbar_init(dfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 256);
  if (threadIdx.x < NPG) {
    const float sc = threadIdx.x < np ? p.scale[b0 + threadIdx.x] : 1.f;
    sc_s[threadIdx.x] = sc;
    sc_s[NPG + threadIdx.x] = 1.f / sc;              // a power of two: exact
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) ptx::prefetch_tmap(&tm_wT);
    __syncwarp();
    for (int s = 0; s < n_slabs; ++s) {
      const uint32_t st = s & 1;
      ptx::mbar_wait(wempty(st), ((s >> 1) & 1) ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(wfull(st), 2 * W_TILE);
        ptx::tma_load_2d(w_base + (st * 2) * W_TILE, &tm_wT, wfull(st), s * BK, 0);
        ptx::tma_load_2d(w_base + (st * 2 + 1) * W_TILE, &tm_wT, wfull(st), p.Vp + s * BK, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    for (int s = 0; s < n_slabs; ++s) {
      const uint32_t st = s & 1, ph = (s >> 1) & 1;
      ptx::mbar_wait(qfull(st), ph);
      ptx::mbar_wait(wfull(st), ph);
      ptx::tc_fence_after();
      const uint64_t whi = ptx::umma_desc_sw128(w_base + (st * 2) * W_TILE), wlo = whi + (uint64_t)(W_TILE >> 4);
      const uint64_t qhi = ptx::umma_desc_sw128(q_base + (st * 2) * Q_TILE), qlo = qhi + (uint64_t)(Q_TILE >> 4);
      if (ptx::elect_one()) {
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          ptx::mma_f16_ss(tmem_base, whi + 2 * j, qhi + 2 * j, IDESC, (s == 0 && j == 0) ? 0u : 1u);
          ptx::mma_f16_ss(tmem_base, whi + 2 * j, qlo + 2 * j, IDESC, 1u);
          ptx::mma_f16_ss(tmem_base, wlo + 2 * j, qhi + 2 * j, IDESC, 1u);
        }
        ptx::mma_commit(wempty(st));
        ptx::mma_commit(qempty(st));
        if (s == n_slabs - 1) ptx::mma_commit(dfull);
      }
      __syncwarp();
    }
  } else {
    // ---- compute warps: thread = (two adjacent vertices of the slab, one pose): the two fp16 values of a Q row land in
    // one 32-bit shared-memory store (a warp writes a whole 128-byte row; with one vertex per thread the kernel was bound
    // by the count of 16-bit STS instructions, 768 per slab)
    const int t = threadIdx.x - 64;
    const int vp2 = t & 31, pl = t >> 5;                  // vertex pair within the slab, pose within the CTA
    static_assert(CW * 32 == NPG * (BK / 2), "one thread per (vertex pair, pose)");
    float gn_[2][3], xn_[2][3];
    auto fetch = [&](int s, float (*gg)[3], float (*xx)[3]) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int v = s * BK + 2 * vp2 + i;
        gg[i][0] = gg[i][1] = gg[i][2] = 0.f;
        xx[i][0] = xx[i][1] = xx[i][2] = 0.f;
        if (v < V && pl < np) {
          const float* gv = p.g_verts + ((size_t)(b0 + pl) * V + v) * 3;
          const float* xp = p.vposed + ((size_t)(b0 + pl) * V + v) * 3;
          gg[i][0] = gv[0]; gg[i][1] = gv[1]; gg[i][2] = gv[2];
          xx[i][0] = xp[0]; xx[i][1] = xp[1]; xx[i][2] = xp[2];
          const int qn = p.need_index ? p.need_index[v] : -1;
          if (qn >= 0 && p.gextra) {
            const float* ge = p.gextra + ((size_t)(b0 + pl) * p.n_need + qn) * 3;
            gg[i][0] += ge[0]; gg[i][1] += ge[1]; gg[i][2] += ge[2];
          }
        }
      }
    };
    fetch(0, gn_, xn_);
    const float sc = sc_s[pl];                            // into fp16's normal range (exact: a power of two)
    for (int s = 0; s < n_slabs; ++s) {
      const uint32_t st = s & 1;
      float q[2][12];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float g[3] = {gn_[i][0] * sc, gn_[i][1] * sc, gn_[i][2] * sc};
        const float x[3] = {xn_[i][0], xn_[i][1], xn_[i][2]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          q[i][3 * a + 0] = g[a] * x[0];
          q[i][3 * a + 1] = g[a] * x[1];
          q[i][3 * a + 2] = g[a] * x[2];
          q[i][9 + a] = g[a];
        }
      }
      if (s + 1 < n_slabs) fetch(s + 1, gn_, xn_);        // next slab's operands in flight across the stores / barrier
      ptx::mbar_wait(qempty(st), ((s >> 1) & 1) ^ 1);     // the MMAs of slab s - 2 are done with this buffer
      uint8_t* qh = smem + OFF_Q + (st * 2) * Q_TILE;
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        const uint32_t off = sw128_off(pl * 12 + e, 2 * vp2);
        const __half2 hi = __floats2half2_rn(q[0][e], q[1][e]);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(q[0][e] - hf.x, q[1][e] - hf.y);
        *reinterpret_cast<__half2*>(qh + off) = hi;
        *reinterpret_cast<__half2*>(qh + Q_TILE + off) = lo;
      }
      ptx::fence_proxy_async_smem();                      // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(qfull(st));
    }
    if (warp < 6) {
      // ---- drain: warp % 4 = TMEM lane quarter, lane = joint row (row J: the all-ones row = translation cotangent)
      const int qd = warp & 3;
      const int j = qd * 32 + lane;
      ptx::mbar_wait(dfull, 0);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < NQ / 32; ++c) {
        uint32_t d[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + c * 32, d);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i, pl = col / 12, e = col % 12;
          if (pl < np) {
            const float val = __uint_as_float(d[i]) * sc_s[NPG + pl];
            if (j < J) p.gA[((size_t)(b0 + pl) * J + j) * 12 + e] = val;
            else if (j == J && e >= 9) p.gbt[(size_t)(b0 + pl) * (p.S + 3) + p.S + (e - 9)] += val;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace lsb

// W^T [128, 2*Vp] fp16 [hi | lo]: row j < J = skinning weights of joint j over the vertices, row J = ones, rest zero
int lbs_skin_bwd_tc_prepare(dpb_lbs* h, const dpb_body_tensors* m) {
  const int V = h->V, J = h->J;
  h->sb_ready = false;
  if (V <= 1024) {   // small vertex sets: per-joint vertex lists for lbs_skin_bwd_small_kernel
    std::vector<int32_t> ptr(J + 1, 0), vs;
    std::vector<float> ws;
    for (int j = 0; j < J; ++j) {
      for (int v = 0; v < V; ++v) {
        const float x = m->lbs_weights[(size_t)v * J + j];
        if (x != 0.f) { vs.push_back(v); ws.push_back(x); }
      }
      ptr[j + 1] = (int32_t)vs.size();
    }
    DPB_CUDA_CHECK(cudaMalloc((void**)&h->csr_ptr, ptr.size() * 4));
    DPB_CUDA_CHECK(cudaMalloc((void**)&h->csr_v, std::max<size_t>(vs.size(), 1) * 4));
    DPB_CUDA_CHECK(cudaMalloc((void**)&h->csr_w, std::max<size_t>(ws.size(), 1) * 4));
    DPB_CUDA_CHECK(cudaMemcpy(h->csr_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
    if (!vs.empty()) {
      DPB_CUDA_CHECK(cudaMemcpy(h->csr_v, vs.data(), vs.size() * 4, cudaMemcpyHostToDevice));
      DPB_CUDA_CHECK(cudaMemcpy(h->csr_w, ws.data(), ws.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  if (J >= 128) return DPB_OK;                          // needs a spare row for the ones
  h->sb_vp = (V + 63) / 64 * 64;
  const size_t ld = (size_t)2 * h->sb_vp;
  std::vector<__half> buf((size_t)128 * ld, __float2half_rn(0.f));
  for (int v = 0; v < V; ++v) {
    for (int j = 0; j < J; ++j) {
      const float x = m->lbs_weights[(size_t)v * J + j];
      const __half hi = __float2half_rn(x);
      buf[(size_t)j * ld + v] = hi;
      buf[(size_t)j * ld + h->sb_vp + v] = __float2half_rn(x - __half2float(hi));
    }
    buf[(size_t)J * ld + v] = __float2half_rn(1.f);
  }
  DPB_CUDA_CHECK(cudaMalloc((void**)&h->wT16, buf.size() * sizeof(__half)));
  DPB_CUDA_CHECK(cudaMemcpy(h->wT16, buf.data(), buf.size() * sizeof(__half), cudaMemcpyHostToDevice));
  int rc = make_tmap_2d(&h->tm_wT, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, h->wT16, ld, 128, lsb::BK, 128, 2);
  if (rc != DPB_OK) return rc;
  const size_t smem = lsb::OFF_A + 2 * lsb::NPG * 4 + 1024;
  if (smem > 232448) return DPB_OK;
  DPB_CUDA_CHECK(cudaFuncSetAttribute(lsb::lbs_skin_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  h->sb_smem = (int)smem;
  h->sb_ready = true;
  return DPB_OK;
}

namespace lsb {
// dL/dA for SMALL vertex sets (the compact set of the joints-only mode: 21 / 174 vertices): the GEMM above is all
// prologue and drain there (2.1 ms per 131 072 poses for three slabs per CTA).  One block per pose; the pose's cotangents
// and blended vertices sit in shared memory; thread = one of the J x 12 outputs walks the joint's vertex list (CSR of the
// sparse weights, ascending vertex order: deterministic).
// The kernel is issue-bound (ncu: 7.7 k warp instructions per pose at 87 % issue utilisation), so the per-list-element
// work is cut down: the outer products q_v = g_v (x) [x_v | 1] are formed once per vertex, a block serves PB poses with
// one walk of the index / weight lists, and the walk is one shared-memory load and one FMA per pose.
template <int PB>
__global__ void __launch_bounds__(256) lbs_skin_bwd_small_kernel(
    const float* __restrict__ vposed, const float* __restrict__ g_verts, const int32_t* __restrict__ csr_ptr,
    const int32_t* __restrict__ csr_v, const float* __restrict__ csr_w, int V, int J, int S, float* __restrict__ gA,
    float* __restrict__ gbt, int64_t B) {
  extern __shared__ float sm[];                     // q [PB][V,12]
  const int64_t b0 = (int64_t)blockIdx.x * PB;
  for (int i = threadIdx.x; i < PB * V; i += 256) {
    const int pb = i / V, v = i - pb * V;
    const int64_t b = b0 + pb;
    float* q = sm + ((size_t)pb * V + v) * 12;
    if (b < B) {
      const float* gp = g_verts + ((size_t)b * V + v) * 3;
      const float* xp = vposed + ((size_t)b * V + v) * 3;
      const float g0 = gp[0], g1 = gp[1], g2 = gp[2], x0 = xp[0], x1 = xp[1], x2 = xp[2];
      q[0] = g0 * x0; q[1] = g0 * x1; q[2] = g0 * x2;
      q[3] = g1 * x0; q[4] = g1 * x1; q[5] = g1 * x2;
      q[6] = g2 * x0; q[7] = g2 * x1; q[8] = g2 * x2;
      q[9] = g0; q[10] = g1; q[11] = g2;
    } else {
#pragma unroll
      for (int e = 0; e < 12; ++e) q[e] = 0.f;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < J * 12; o += 256) {
    const int j = o / 12, e = o - j * 12;
    float acc[PB];
#pragma unroll
    for (int pb = 0; pb < PB; ++pb) acc[pb] = 0.f;
    const int k1 = csr_ptr[j + 1];
    for (int k = csr_ptr[j]; k < k1; ++k) {          // ascending vertex order: deterministic
      const float w = csr_w[k];
      const float* q = sm + (size_t)csr_v[k] * 12 + e;
#pragma unroll
      for (int pb = 0; pb < PB; ++pb) acc[pb] = fmaf(w, q[(size_t)pb * V * 12], acc[pb]);
    }
#pragma unroll
    for (int pb = 0; pb < PB; ++pb)
      if (b0 + pb < B) gA[((size_t)(b0 + pb) * J) * 12 + o] = acc[pb];
  }
  if (threadIdx.x < 3 * PB) {  // translation cotangent = sum_v g (ascending vertex order)
    const int pb = threadIdx.x / 3, c = threadIdx.x - pb * 3;
    if (b0 + pb < B) {
      float acc = 0.f;
      for (int v = 0; v < V; ++v) acc += sm[((size_t)pb * V + v) * 12 + 9 + c];
      gbt[(size_t)(b0 + pb) * (S + 3) + S + c] += acc;
    }
  }
}
}  // namespace lsb

int lbs_skin_bwd_small(dpb_lbs* h, const float* vposed, const float* g_verts, float* gA, float* gbt, int64_t B,
                       cudaStream_t st) {
  constexpr int PB = 2;                                       // poses per block: one walk of the joint lists serves both
  const size_t smem = (size_t)PB * h->V * 12 * 4;             // V <= 1024 (prepare): at most 96 KB
  DPB_CUDA_CHECK(cudaFuncSetAttribute(lsb::lbs_skin_bwd_small_kernel<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lsb::lbs_skin_bwd_small_kernel<PB><<<(unsigned)((B + PB - 1) / PB), 256, smem, st>>>(vposed, g_verts, h->csr_ptr, h->csr_v,
                                                                                   h->csr_w, h->V, h->J, h->S, gA, gbt, B);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

void lbs_skin_bwd_tc_release(dpb_lbs* h) {
  if (h->csr_ptr) cudaFree(h->csr_ptr);
  if (h->csr_v) cudaFree(h->csr_v);
  if (h->csr_w) cudaFree(h->csr_w);
  h->csr_ptr = h->csr_v = nullptr;
  h->csr_w = nullptr;
  if (h->wT16) cudaFree(h->wT16);
  h->wT16 = nullptr;
  h->sb_ready = false;
}

namespace lsb {
// scale[b] = 2^k with max |cotangent of pose b| * scale in [32, 64): fp16 operands keep their full significand and the
// [hi | lo] split its second half (cotangents of mean-reduced losses are ~1e-5: fp16-subnormal without it)
__global__ void __launch_bounds__(256) lbs_rowscale_kernel(const float* __restrict__ g, int64_t n, const float* __restrict__ g2,
                                                           int64_t n2, float* __restrict__ scale) {
  const int64_t b = blockIdx.x;
  float m = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) m = fmaxf(m, fabsf(g[b * n + i]));
  float m2 = 0.f;
  if (g2)
    for (int64_t i = threadIdx.x; i < n2; i += 256) m2 = fmaxf(m2, fabsf(g2[b * n2 + i]));
  m += m2;                                               // the two are added on some vertices
  __shared__ float part[8];
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, part[i]);
    int e = 0;
    float sc = 1.f;
    if (m > 0.f && m < 3.0e38f) {                          // finite and non-zero (NaN fails both comparisons)
      frexpf(m, &e);                                       // m = f * 2^e, f in [0.5, 1)
      e = 6 - e;
      e = e < -100 ? -100 : (e > 100 ? 100 : e);
      sc = ldexpf(1.f, e);
    }
    scale[b] = sc;
  }
}
}  // namespace lsb

// scale [B] (scratch) = power-of-two scale of every pose's cotangents, shared by both halves of the skinning adjoint
int lbs_bwd_rowscale(dpb_lbs* h, const float* g_verts, const float* gextra, bool have_extra, float* scale, int64_t B,
                     cudaStream_t st) {
  lsb::lbs_rowscale_kernel<<<(unsigned)B, 256, 0, st>>>(g_verts, (int64_t)h->V * 3, have_extra ? gextra : nullptr,
                                                        (int64_t)h->n_need * 3, scale);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

// gA [B,J,12] is WRITTEN (no zeroing needed), gbt[:, S:S+3] += translation cotangent
int lbs_skin_bwd_tc(dpb_lbs* h, const float* vposed, const float* g_verts, const float* gextra, bool have_extra,
                    float* gA, float* gbt, const float* scale, int64_t B, cudaStream_t st) {
  lsb::Params p{};
  p.scale = scale;
  p.vposed = vposed; p.g_verts = g_verts;
  p.gextra = have_extra ? gextra : nullptr;
  p.need_index = have_extra ? h->need_index : nullptr;
  p.gA = gA; p.gbt = gbt;
  p.V = h->V; p.J = h->J; p.S = h->S; p.n_need = h->n_need; p.Vp = h->sb_vp;
  p.B = B;
  lsb::lbs_skin_bwd_tc_kernel<<<(unsigned)((B + lsb::NPG - 1) / lsb::NPG), lsb::NUM_THREADS, h->sb_smem, st>>>(p, h->tm_wT);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

}  // namespace dpb
