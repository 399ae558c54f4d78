// Small-batch engine of the fused sampler / score forward (configs[0]: 500 poses, run/demo.py:127-145).
// Replaces the same reference code as score_tc.cu (ScoreModelFC.forward model.py:141-196, pc_sampler's hot loop
// sampling.py:456-461 with the Euler-Maruyama family of predictors :182-259 in affine form) for B <= 1024 rows.
//
// Why: score_tc_kernel gives one CTA a whole 128-row tile, so 500 rows keep 4 of 148 SMs busy and a sampler step is six
// dependent 128 x 1024 x 1024 GEMMs on one SM each (99 us per step measured).  Here the 1024 output features of every
// layer are split over NSPLIT = 16 CTAs per row tile (64 CTAs for 500 rows):
//   * CTA (tile, s) owns columns [64 s, 64 s + 64) of every hidden layer: its weight slice [64 x K] streams by TMA, its
//     accumulator is 64 TMEM columns, GroupNorm groups (32 consecutive channels) stay thread-local (thread = row), and
//     the residual stream of its columns never leaves shared memory (fp32);
//   * a layer's output slice goes to the L2-resident activation scratch (fp16, the handle's act_h / act_t buffers), the
//     16 CTAs of a row tile meet at a counter barrier, and the next layer's A operand -- the full [128 x 1024] tile -- comes
//     back by TMA; post_dense (N = 64) and the sampler update run on split 0, which publishes the new x to the others;
//   * six barriers per step; weights 0.54 MB + activations 1.3 MB from L2 per CTA and step.
// Operand formats and arithmetic are those of score_tc.cu (fp16 operands / fp32 accumulate, bf16 hi/lo split of the 63-wide
// input as a K extension, Philox keyed by (row, step, slot)); SiLU is the exact form.
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "ptx.cuh"
#include "score.h"

namespace dpb {

int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t rows,
                 uint32_t box_inner, uint32_t box_rows, size_t elem_bytes);  // score_tc.cu

namespace tcs {

constexpr int TILE_M = 128, BK = 64, NW = 64, NSPLIT = H / NW;   // 16 column slices of 64
constexpr int A_BYTES = TILE_M * BK * 2;     // 16 KB
constexpr int W_BYTES = NW * BK * 2;         // 8 KB
constexpr int STAGE = A_BYTES + W_BYTES;
constexpr int STAGES = 5;
constexpr int XA_SLABS = 3;                  // [x_hi | x_lo | x_hi] bf16, 64 columns each
constexpr int OFF_XA = STAGES * STAGE;       // 122880
constexpr int OFF_RES = OFF_XA + XA_SLABS * A_BYTES;          // fp32 residual stream [64 cols][128 rows]
constexpr int OFF_BAR = OFF_RES + NW * TILE_M * 4;
constexpr int NBARS = 2 * STAGES + 3;        // full, empty, dfull, xa_ready, go
constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;
constexpr int NUM_THREADS = 192;             // warp 0 TMA, warp 1 TMEM + MMA, warps 2-5 epilogue (thread = row)
static_assert(OFF_XA % 1024 == 0 && OFF_RES % 1024 == 0 && SMEM_BYTES <= 232448, "shared memory layout");
constexpr uint32_t IDESC_F16 = ptx::umma_idesc_f16(TILE_M, NW, 0);
constexpr uint32_t IDESC_BF16 = ptx::umma_idesc_f16(TILE_M, NW, 1);

struct Params {
  int mode, n_steps, impute, noise_k, n_tiles;
  long long B;
  const float* x_in;        // mode 0
  float* x_io;              // mode 1 (state)
  const float* table;       // [n_steps,5,1024]
  const float* coef;        // [n_steps,8]
  const float* gn;          // [5][2][1024] gamma | beta
  const float* post_b;      // [64]
  const float* row_scale;
  float scale;
  float* out;
  const float* obs;
  const float* mask;
  const float* noise;
  unsigned long long seed, step_offset;
  float* traj;
  float* x_mean;
  __half* act_p;            // [n_tiles*128, 1024] ping
  __half* act_q;            // pong
  int* counters;            // [n_tiles] barrier counters (zeroed before the launch)
};

__device__ __forceinline__ void draw_row(const float* plane, long long row, unsigned long long seed, uint32_t step,
                                         uint32_t slot, float* z) {   // 64 values, columns >= 63 unused
  if (plane) {
#pragma unroll
    for (int c = 0; c < 64; ++c) z[c] = c < D ? plane[row * D + c] : 0.f;
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) normal4(seed, (uint64_t)row, step, slot, (uint32_t)q, z + 4 * q);
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
score_small_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap tm_p,
                   const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_pre,
                   const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
                   const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_w3,
                   const __grid_constant__ CUtensorMap tm_post) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR;
  auto full = [&](uint32_t s) { return bar + 8u * s; };
  auto empty = [&](uint32_t s) { return bar + 8u * (STAGES + s); };
  const uint32_t dfull = bar + 8u * (2 * STAGES), xa_ready = dfull + 8, go = dfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBARS * 8);
  float* res = reinterpret_cast<float*>(smem + OFF_RES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / NSPLIT, ns = blockIdx.x % NSPLIT;
  const int row0 = tile * TILE_M;
  const bool lead = ns == 0;                                    // runs post_dense + the update for this row tile
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    ptx::mbar_init(dfull, 1);
    ptx::mbar_init(xa_ready, 4);
    ptx::mbar_init(go, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 64);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // layers of one step as every role walks them: 0 = pre_dense (A from xa), 1..4 = hidden (A from the scratch),
  // 5 = post_dense (split 0 only).  The activations alternate between act_p (layers 0, 2, 4 write it) and act_q.
  const int n_layers = lead ? 6 : 5;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_p); ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_pre);
      ptx::prefetch_tmap(&tm_w0); ptx::prefetch_tmap(&tm_w1); ptx::prefetch_tmap(&tm_w2); ptx::prefetch_tmap(&tm_w3);
      ptx::prefetch_tmap(&tm_post);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0, gph = 0;
    // `go` completes one phase per tile barrier that is followed by loads: after hidden layers 0..3 (and 4 on split 0,
    // before post_dense) and at the end of a step that is not the last.  Every phase is consumed exactly once.
    auto wait_go = [&]() { ptx::mbar_wait(go, gph); gph ^= 1; };
    for (int step = 0; step < p.n_steps; ++step) {
      if (step > 0) wait_go();                                    // end of the previous step (the new x is published)
      for (int layer = 0; layer < n_layers; ++layer) {
        if (layer > 0) wait_go();                                 // layer - 1's slices of all 16 CTAs are in L2
        const CUtensorMap* tw = layer == 0 ? &tm_pre : layer == 1 ? &tm_w0 : layer == 2 ? &tm_w1 : layer == 3 ? &tm_w2
                              : layer == 4 ? &tm_w3 : &tm_post;
        const CUtensorMap* ta = (layer & 1) ? &tm_p : &tm_q;      // layer l reads what layer l-1 wrote (0, 2, 4 -> act_p)
        const int n_slabs = layer == 0 ? XA_SLABS : H / BK;
        const int wrow = layer == 5 ? 0 : ns * NW;
        for (int ks = 0; ks < n_slabs; ++ks) {
          ptx::mbar_wait(empty(stage), phase ^ 1);
          if (ptx::elect_one()) {
            const uint32_t dst = sb + stage * STAGE;
            if (layer == 0) {
              ptx::mbar_arrive_expect_tx(full(stage), W_BYTES);
            } else {
              ptx::mbar_arrive_expect_tx(full(stage), STAGE);
              ptx::tma_load_2d(dst, ta, full(stage), ks * BK, row0);
            }
            ptx::tma_load_2d(dst + A_BYTES, tw, full(stage), ks * BK, wrow);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    uint32_t stage = 0, phase = 0, xph = 0;
    for (int step = 0; step < p.n_steps; ++step)
      for (int layer = 0; layer < n_layers; ++layer) {
        const int n_slabs = layer == 0 ? XA_SLABS : H / BK;
        if (layer == 0) {
          ptx::mbar_wait(xa_ready, xph);                          // the epilogue wrote this step's x operand
          xph ^= 1;
        }
        for (int ks = 0; ks < n_slabs; ++ks) {
          ptx::mbar_wait(full(stage), phase);
          ptx::tc_fence_after();
          const uint64_t ad = ptx::umma_desc_sw128(layer == 0 ? sb + OFF_XA + ks * A_BYTES : sb + stage * STAGE);
          const uint64_t wd = ptx::umma_desc_sw128(sb + stage * STAGE + A_BYTES);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j)
              ptx::mma_f16_ss(tmem_base, ad + 2 * j, wd + 2 * j, layer == 0 ? IDESC_BF16 : IDESC_F16,
                              (ks == 0 && j == 0) ? 0u : 1u);
            ptx::mma_commit(empty(stage));
            if (ks == n_slabs - 1) ptx::mma_commit(dfull);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
  } else {
    // ======================= epilogue: thread = row =======================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const long long row = (long long)row0 + r;
    const bool valid = row < p.B;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int et = threadIdx.x - 64;                              // 0..127
    uint32_t dph = 0, bar_count = 0;
    const size_t plane = (size_t)p.B * D;
    // counter barrier of the 16 CTAs of this row tile; then release the producer (layer operands are in L2)
    auto tile_barrier = [&](bool release_producer) {
      __threadfence();
      ptx::named_bar_sync(1, 128);
      ++bar_count;
      if (et == 0) {
        atomicAdd(p.counters + tile, 1);
        const int target = (int)(bar_count * NSPLIT);
        while (ptx::ld_acquire_gpu(p.counters + tile) < target) {
        }
        __threadfence();
      }
      ptx::named_bar_sync(1, 128);
      ptx::fence_proxy_async_global();                            // generic-proxy writes of the peers -> TMA reads
      if (et == 0 && release_producer) ptx::mbar_arrive(go);
    };
    const float* xsrc = p.mode == 0 ? p.x_in : p.x_io;
    if (p.mode == 1 && p.impute) {
      // imputation that follows the (none) corrector of step 0 (sampling.py:459): split 0 applies it to x_io
      if (lead && valid) {
        float zc[64];
        draw_row(p.noise, row, p.seed, (uint32_t)p.step_offset, 0, zc);
        const float al = p.coef[3], sd = p.coef[4];
        for (int c = 0; c < D; ++c) {
          const float m = p.mask[row * D + c];
          p.x_io[row * D + c] = p.x_io[row * D + c] * (1.0f - m) + (al * p.obs[row * D + c] + zc[c] * sd) * m;
        }
      }
      __threadfence();
      ptx::named_bar_sync(1, 128);
      ++bar_count;
      if (et == 0) {
        atomicAdd(p.counters + tile, 1);
        while (ptx::ld_acquire_gpu(p.counters + tile) < (int)(bar_count * NSPLIT)) {
        }
        __threadfence();
      }
      ptx::named_bar_sync(1, 128);
    }
    for (int step = 0; step < p.n_steps; ++step) {
      // ---- this step's x: the bf16 [hi | lo | hi] operand of pre_dense, rows of 128 bytes, SWIZZLE_128B
      float x[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) x[c] = (valid && c < D) ? __ldcg(xsrc + row * D + c) : 0.f;
      {
        uint8_t* xa = smem + OFF_XA;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {                          // 16-byte chunks of 8 columns
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = x[ch * 8 + 2 * i], b = x[ch * 8 + 2 * i + 1];
            const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(a - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(b - __bfloat162float(h1));
            hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(xa + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(xa + A_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(xa + 2 * A_BYTES + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(xa_ready);
      }
      // ---- hidden layers 0..4: bias, GroupNorm (two thread-local groups), SiLU, residual, fp16 slice to the scratch
      for (int layer = 0; layer < 5; ++layer) {
        const float* tb = p.table + ((size_t)step * NL + layer) * H + ns * NW;
        const float* gam = p.gn + ((size_t)layer * 2) * H + ns * NW;
        const float* bet = gam + H;
        ptx::mbar_wait(dfull, dph);
        dph ^= 1;
        ptx::tc_fence_after();
        __half* dst = ((layer & 1) ? p.act_q : p.act_p) + (size_t)row * H + ns * NW;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + lane_addr + g * 32, v);
          ptx::tmem_ld_wait();
          float y[32];
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) { y[i] = __uint_as_float(v[i]) + __ldg(tb + g * 32 + i); s += y[i]; }
          const float mean = s * (1.0f / 32.0f);
          float qv = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) { y[i] -= mean; qv = fmaf(y[i], y[i], qv); }
          const float rstd = rsqrtf(qv * (1.0f / 32.0f) + 1e-5f);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float t = y[i] * rstd * __ldg(gam + g * 32 + i) + __ldg(bet + g * 32 + i);
            float a = t / (1.0f + __expf(-t));
            float* rp = res + (g * 32 + i) * TILE_M + r;
            if (layer == 0) *rp = a;                              // h = act(gn(pre_dense))            model.py:166-169
            else if (layer == 2 || layer == 4) { a += *rp; *rp = a; }   // h = h + block(h)                  model.py:187
            y[i] = a;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
            pk[i] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          uint4* d4 = reinterpret_cast<uint4*>(dst + g * 32);     // rows beyond B carry x = 0: harmless, the slot exists
#pragma unroll
          for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
        ptx::tc_fence_before();
        tile_barrier(layer < 4 || lead);                          // after layer 4 only split 0 loads again (post_dense)
      }
      // ---- post_dense + tail on split 0; the others wait for the new x
      if (lead) {
        ptx::mbar_wait(dfull, dph);
        dph ^= 1;
        ptx::tc_fence_after();
        float raw[64];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + lane_addr + g * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[g * 32 + i] = __uint_as_float(v[i]) + __ldg(p.post_b + g * 32 + i);
        }
        ptx::tc_fence_before();
        if (valid) {
          if (p.mode == 0) {
            const float sc = p.row_scale ? p.row_scale[row] : p.scale;
#pragma unroll
            for (int c = 0; c < D; ++c) p.out[row * D + c] = raw[c] * sc;
          } else {
            const bool last = (step + 1 == p.n_steps);
            const float* cf = p.coef + (size_t)step * DPB_COEF_STRIDE;
            const uint32_t gstep = (uint32_t)(p.step_offset + (unsigned long long)step);
            const float* nz = p.noise ? p.noise + (size_t)step * p.noise_k * plane : nullptr;
            const float ca = cf[0], cb = cf[1], cc = cf[2], al = cf[3], sd = cf[4];
            float z[64];
            draw_row(nz ? nz + (p.noise_k == 3 ? plane : 0) : nullptr, row, p.seed, gstep, 1, z);
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const float xm = ca * x[c] + cb * raw[c];           // sampling.py:185-186 in affine form
              x[c] = xm + cc * z[c];
              if (last && p.x_mean) p.x_mean[row * D + c] = xm;
            }
            if (p.impute && (last || p.traj)) {   // else overwritten by the next step's pre-imputation before any read
              draw_row(nz ? nz + 2 * plane : nullptr, row, p.seed, gstep, 2, z);
#pragma unroll
              for (int c = 0; c < D; ++c) {
                const float m = p.mask[row * D + c];
                x[c] = x[c] * (1.0f - m) + (al * p.obs[row * D + c] + z[c] * sd) * m;
              }
            }
            if (p.traj) {
#pragma unroll
              for (int c = 0; c < D; ++c) p.traj[((size_t)step * p.B + row) * D + c] = x[c];
            }
            if (!last && p.impute) {   // imputation in the corrector slot of the NEXT step precedes its score eval
              const float* nz1 = p.noise ? p.noise + (size_t)(step + 1) * p.noise_k * plane : nullptr;
              draw_row(nz1, row, p.seed, gstep + 1, 0, z);
              const float al1 = cf[DPB_COEF_STRIDE + 3], sd1 = cf[DPB_COEF_STRIDE + 4];
#pragma unroll
              for (int c = 0; c < D; ++c) {
                const float m = p.mask[row * D + c];
                x[c] = x[c] * (1.0f - m) + (al1 * p.obs[row * D + c] + z[c] * sd1) * m;
              }
            }
#pragma unroll
            for (int c = 0; c < D; ++c) p.x_io[row * D + c] = x[c];
          }
        }
      }
      if (step + 1 < p.n_steps) tile_barrier(true);                 // the new x is visible to every split (no TMA follows: the
                                                                  // producer's `go` is consumed by layer 1 of the next step)
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace tcs

int tcs_prepare(dpb_score* h) {
  int rc = DPB_OK;
  for (int l = 0; l < 4 && rc == DPB_OK; ++l)
    rc = make_tmap_2d(&h->tms_w[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, h->w16[l], H, H, tcs::BK, tcs::NW, 2);
  if (rc == DPB_OK) rc = make_tmap_2d(&h->tms_post, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, h->post16, H, DP, tcs::BK, tcs::NW, 2);
  if (rc == DPB_OK) rc = make_tmap_2d(&h->tms_pre, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, h->pre_split, 192, H, tcs::BK, tcs::NW, 2);
  if (rc != DPB_OK) return rc;
  DPB_CUDA_CHECK(cudaFuncSetAttribute(tcs::score_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcs::SMEM_BYTES));
  h->tcs_ready = true;
  return DPB_OK;
}

bool tcs_wanted(const dpb_score* h, const TcJob& j) {
  if (!h->tcs_ready || (j.mode != 0 && j.mode != 1)) return false;
  // MEASURED (profiles/r2_c1_small_batch.md): 100.8 us per sampler step against 98.6 us for the whole-tile kernel at
  // 128..1024 rows -- a counter barrier plus a cold TMA chain per layer costs what one SM's MMAs cost -- so this engine is
  // opt-in (DPB_TC_SMALL=1); the Langevin corrector's single forward passes are 7 % faster with it.
  const char* e = getenv("DPB_TC_SMALL");
  if (!(e && atoi(e) == 1)) return false;
  const int64_t n_tiles = (j.B + tcs::TILE_M - 1) / tcs::TILE_M;
  return n_tiles * tcs::NSPLIT <= (h->sm_count / tcs::NSPLIT) * tcs::NSPLIT && n_tiles <= h->tc_slots;   // all CTAs co-resident
}

int tcs_launch(dpb_score* h, const TcJob& j, cudaStream_t st) {
  tcs::Params p{};
  p.mode = j.mode;
  p.n_steps = j.mode == 0 ? 1 : j.n_steps;
  p.impute = j.impute;
  p.noise_k = j.noise_k;
  p.B = j.B;
  p.n_tiles = (int)((j.B + tcs::TILE_M - 1) / tcs::TILE_M);
  p.x_in = j.x_in; p.x_io = j.x_io; p.table = j.table; p.coef = j.coef;
  p.gn = h->gn_packed; p.post_b = h->post_b;
  p.row_scale = j.row_scale; p.scale = j.scale; p.out = j.out;
  p.obs = j.obs; p.mask = j.mask; p.noise = j.noise;
  p.seed = j.seed; p.step_offset = j.step_offset; p.traj = j.traj; p.x_mean = j.x_mean;
  p.act_p = h->act_h; p.act_q = h->act_t;
  p.counters = h->tc_flags;
  DPB_CUDA_CHECK(cudaMemsetAsync(h->tc_flags, 0, sizeof(int) * p.n_tiles, st));
  tcs::score_small_kernel<<<p.n_tiles * tcs::NSPLIT, tcs::NUM_THREADS, tcs::SMEM_BYTES, st>>>(
      p, h->tm_act_h, h->tm_act_t, h->tms_pre, h->tms_w[0], h->tms_w[1], h->tms_w[2], h->tms_w[3], h->tms_post);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

}  // namespace dpb
