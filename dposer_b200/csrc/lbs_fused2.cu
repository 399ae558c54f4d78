// Fused LBS forward, second generation: blend + skinning in ONE tcgen05 kernel on CTA PAIRS (cta_group::2), for
// SMPL and (with the constant hand / face tail folded away) SMPL-X.  Replaces smplx 0.1.28 lbs() as called from
// lib/body_model/body_model.py:75-88 and lib/body_model/smpl.py:67-78.
//
// Why a second kernel (profiles/r1_ncu_full_summary_final.md, lbs_fused_tc_kernel): tensor pipe 49 % active, the
// MMA warp waited 63 % of the time for a T buffer and the eight epilogue warps ran one dependent instruction every
// ~5 cycles.  The tensor work itself is fixed by the 3-product fp16 split (hi.hi + hi.lo + lo.hi), so the whole gap
// is overlap.  Changes:
//   * CTA pair, M = 256: each CTA keeps its own 128-vertex tile of the blend basis / skinning weights (A operands)
//     and only HALF of every pose-side B operand, so an N = 96 MMA fetches 32 + 12 cycles of shared memory for 48
//     cycles of math (N = 128 on one CTA: 32 + 32 for 64 -- the pipe and the epilogue's staging compete for the same
//     128 B/clk).  The pose operand of a group (48 rows) and a deep basis ring fit next to each other.
//   * 96 poses per group: D_x | D_y | D_z take 288 TMEM columns, leaving THREE 64-column T buffers (5 poses x 12
//     entries): the MMA -> epilogue -> MMA round trip (~650 cycles) is covered by two chunks in flight.
//   * 12 epilogue warps in three sets; a set owns every third chunk, so chunks are in their math / store tail
//     while the next ones are being loaded from TMEM.
//   * v_template rides in a spare K slot of the blend (feature = 1), the translation in the spare joint slot of the
//     skinning GEMM: the epilogue is 9 FFMA + 3 stores per (vertex, pose).
//   * work = (pose group, tile pair) items laid end to end and cut into equal contiguous pieces per CTA pair.
//
//   warp 0  TMA: pose-group operand F (per group), blend-basis slabs (ring)        [both CTAs, own tile / own half]
//   warp 3  TMA: skinning weights (per tile, double buffered), transform chunks (ring)
//   warp 1  MMA issuer (leader CTA only);   warp 2  TMEM allocator
//   warps 4-15 epilogue: set = (warp - 4) / 4, TMEM lane quarter = warp % 4, thread = vertex
#include <cudaTypedefs.h>

#include <cstdlib>

#include "lbs.h"
#include "ptx.cuh"

namespace dpb {

int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t rows,
                 uint32_t box_inner, uint32_t box_rows, size_t elem_bytes);  // score_tc.cu

namespace lt2 {

constexpr int TILE_V = 128;
constexpr int BK = 64;
constexpr int A_SLAB = TILE_V * BK * 2;        // 16 KB: [128 rows x 64 k] fp16, SWIZZLE_128B
constexpr int NP = 96;                         // poses per group = N of the blend MMAs
constexpr int CP = 5;                          // poses per skinning chunk
constexpr int NS = 64;                         // N of the skinning MMAs (CP * 12 = 60 used)
constexpr int NCH = (NP + CP - 1) / CP;        // 20 chunks per group (the last one holds a single pose)
constexpr int NT = 3;                          // T buffers
constexpr int NSETS = 3;                       // epilogue warp sets (12 warps: 84 TMEM values live per thread need ~150 registers)
constexpr int NUM_THREADS = 128 + NSETS * 128; // 512
constexpr int F_SLAB = (NP / 2) * BK * 2;      // 6 KB: this CTA's 48 poses x 64 k
constexpr int S_SLAB = (NS / 2) * BK * 2;      // 4 KB: this CTA's 32 transform rows x 64 k
constexpr int T_COL0 = 3 * NP;                 // 288: first T column
constexpr int ROLE_REGS = 72, EPI_REGS = 144;  // setmaxnreg: 128 * 72 + 384 * 144 <= 65536
constexpr uint32_t IDESC_BLEND = ptx::umma_idesc_f16(2 * TILE_V, NP, 0);
constexpr uint32_t IDESC_SKIN = ptx::umma_idesc_f16(2 * TILE_V, NS, 0);
static_assert(NCH >= NSETS, "every epilogue set handles at least one chunk per tile");
static_assert(T_COL0 + NT * NS <= 512 && F_SLAB % 1024 == 0 && S_SLAB % 1024 == 0, "TMEM / SWIZZLE_128B layout");

struct Params {
  int V, V_pad, n_tp;        // vertices, padded vertices, tile PAIRS (n_vt / 2)
  int64_t B;
  long long n_items;         // pose groups x tile pairs
  float* verts;              // [B,V,3]
};

template <int JSLABS, int ASTAGES, int SSTAGES>
struct Smem {
  static constexpr int NSLABS_MAX = 7;
  static constexpr int OFF_F = 0;
  static constexpr int OFF_A = OFF_F + NSLABS_MAX * F_SLAB;                 // 43008 (1024-aligned)
  static constexpr int OFF_W = OFF_A + ASTAGES * A_SLAB;
  static constexpr int OFF_S = OFF_W + 2 * JSLABS * A_SLAB;
  static constexpr int OFF_STG = OFF_S + SSTAGES * JSLABS * S_SLAB;         // per-warp store staging: 16 x 2 x 384 B
  static constexpr int OFF_BAR = OFF_STG + NSETS * 4 * 2 * 384;
  static constexpr int NBARS = 2 * ASTAGES + 2 * SSTAGES + 4 + 2 + 2 + 2 * NT;
  static constexpr int BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;             // + tmem slot + alignment slack
  static_assert(OFF_A % 1024 == 0 && OFF_W % 1024 == 0 && OFF_S % 1024 == 0, "operand tiles are 1024-byte aligned");
  static_assert(BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

// KH16 = K16 steps of the hi half of the blend K, NSLABS = 64-wide slabs of [hi | lo]; JS = K16 steps of the hi half of
// the skinning K (joints + translation slot, padded), JSLABS = its 64-wide slabs.  Compile-time: the issuer's loops
// unroll and every descriptor is base + constant.  STAGED: transpose each pose's 32 x (x,y,z) through shared memory
// so that the global stores are three full 128-byte lines per warp instead of three stride-12 scatters.
template <int KH16, int NSLABS, int JS, int JSLABS, int ASTAGES, int SSTAGES, bool STAGED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
lbs_fused2_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap tm_dirs,
                  const __grid_constant__ CUtensorMap tm_feat, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_s) {
  using L = Smem<JSLABS, ASTAGES, SSTAGES>;
  static_assert(NSLABS <= L::NSLABS_MAX && 2 * KH16 <= 4 * NSLABS && 2 * JS <= 4 * JSLABS, "K geometry");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t f_base = sb + L::OFF_F, a_base = sb + L::OFF_A, w_base = sb + L::OFF_W, s_base = sb + L::OFF_S;
  const uint32_t bar = sb + L::OFF_BAR;
  auto afull = [&](uint32_t s) { return bar + 8u * s; };
  auto aempty = [&](uint32_t s) { return bar + 8u * (ASTAGES + s); };
  auto sfull = [&](uint32_t s) { return bar + 8u * (2 * ASTAGES + s); };
  auto sempty = [&](uint32_t s) { return bar + 8u * (2 * ASTAGES + SSTAGES + s); };
  const uint32_t b2 = bar + 8u * (2 * ASTAGES + 2 * SSTAGES);
  auto wfull = [&](uint32_t b) { return b2 + 8u * b; };
  auto wempty = [&](uint32_t b) { return b2 + 8u * (2 + b); };
  const uint32_t ffull = b2 + 32, fempty = b2 + 40, dfull = b2 + 48, dempty = b2 + 56;
  auto tfull = [&](uint32_t b) { return b2 + 64 + 8u * b; };
  auto tempty = [&](uint32_t b) { return b2 + 64 + 8u * (NT + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_BAR + L::NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // "full" barriers are waited on by the pair's leader only and collect one arrive.expect_tx from EACH CTA's
    // producer; "empty" / TMEM-full barriers exist in both CTAs and are signalled by multicast tcgen05.commit
    for (int s = 0; s < ASTAGES; ++s) { ptx::mbar_init(afull(s), 2); ptx::mbar_init(aempty(s), 1); }
    for (int s = 0; s < SSTAGES; ++s) { ptx::mbar_init(sfull(s), 2); ptx::mbar_init(sempty(s), 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(wfull(b), 2); ptx::mbar_init(wempty(b), 1); }
    ptx::mbar_init(ffull, 2); ptx::mbar_init(fempty, 1);
    ptx::mbar_init(dfull, 1); ptx::mbar_init(dempty, 2 * NSETS * 4);        // every epilogue warp of both CTAs
    for (int b = 0; b < NT; ++b) { ptx::mbar_init(tfull(b), 1); ptx::mbar_init(tempty(b), 2 * 4); }  // one set, both CTAs
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc_2sm(ptx::smem_u32(tmem_slot), 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();       // the peer's barriers are initialised before anyone arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = ptx::cluster_ctarank();
  constexpr uint16_t CMASK = 3;
  const long long worker = blockIdx.x >> 1, n_workers = gridDim.x >> 1;
  const long long i0 = p.n_items * worker / n_workers, i1 = p.n_items * (worker + 1) / n_workers;
  const int n_my = (int)(i1 - i0);                 // items of this pair
  const int grp0 = (int)(i0 / p.n_tp), tp0 = (int)(i0 % p.n_tp);
  // walk (group, tile pair) without a division per item
  auto advance = [&](int& grp, int& tp) { if (++tp == p.n_tp) { tp = 0; ++grp; } };

  if (warp < 4) {
    ptx::setmaxnreg_dec<ROLE_REGS>();
    if (warp == 0) {
      // ---- pose-group operand + blend-basis slabs
      if (lane == 0) { ptx::prefetch_tmap(&tm_dirs); ptx::prefetch_tmap(&tm_feat); }
      __syncwarp();
      uint32_t stage = 0, phase = 0, fph = 0;
      int grp = grp0, tp = tp0;
      for (int it = 0; it < n_my; ++it, advance(grp, tp)) {
        const int tile = 2 * tp + (int)crank;
        if (it == 0 || tp == 0) {
          ptx::mbar_wait(fempty, fph ^ 1);   // the previous group's blends are done with the operand
          fph ^= 1;
          const uint32_t lbar = ptx::mapa(ffull, 0);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx_cluster(lbar, NSLABS * F_SLAB);
            for (int i = 0; i < NSLABS; ++i)
              ptx::tma_load_2d_2sm(f_base + i * F_SLAB, &tm_feat, lbar, i * BK, grp * NP + (int)crank * (NP / 2));
          }
        }
        for (int c = 0; c < 3; ++c)
          for (int i = 0; i < NSLABS; ++i) {
            ptx::mbar_wait(aempty(stage), phase ^ 1);
            const uint32_t lbar = ptx::mapa(afull(stage), 0);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx_cluster(lbar, A_SLAB);
              ptx::tma_load_2d_2sm(a_base + stage * A_SLAB, &tm_dirs, lbar, i * BK, c * p.V_pad + tile * TILE_V);
            }
            if (++stage == ASTAGES) { stage = 0; phase ^= 1; }
          }
      }
    } else if (warp == 3) {
      // ---- skinning weights (per tile) + transform chunks
      if (lane == 0) { ptx::prefetch_tmap(&tm_w); ptx::prefetch_tmap(&tm_s); }
      __syncwarp();
      uint32_t stage = 0, phase = 0, tc = 0;
      int grp = grp0, tp = tp0;
      for (int it = 0; it < n_my; ++it, ++tc, advance(grp, tp)) {
        const int tile = 2 * tp + (int)crank;
        const uint32_t wb = tc & 1;
        ptx::mbar_wait(wempty(wb), ((tc >> 1) & 1) ^ 1);
        {
          const uint32_t lbar = ptx::mapa(wfull(wb), 0);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx_cluster(lbar, JSLABS * A_SLAB);
            for (int i = 0; i < JSLABS; ++i)
              ptx::tma_load_2d_2sm(w_base + (wb * JSLABS + i) * A_SLAB, &tm_w, lbar, i * BK, tile * TILE_V);
          }
        }
        for (int ch = 0; ch < NCH; ++ch) {
          ptx::mbar_wait(sempty(stage), phase ^ 1);
          const uint32_t lbar = ptx::mapa(sfull(stage), 0);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx_cluster(lbar, JSLABS * S_SLAB);
            for (int i = 0; i < JSLABS; ++i)
              ptx::tma_load_2d_2sm(s_base + (stage * JSLABS + i) * S_SLAB, &tm_s, lbar, i * BK,
                                   (grp * NP + ch * CP) * 12 + (int)crank * (NS / 2));
          }
          if (++stage == SSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && crank == 0) {
      // ---- MMA issuer (the pair's leader issues for both CTAs; warp-wide loop, the elected lane issues)
      uint32_t astage = 0, aphase = 0, sstage = 0, sphase = 0, fph = 0, dph = 0, tc = 0, cc = 0;
      const uint64_t adesc0 = ptx::umma_desc_sw128(a_base), fdesc0 = ptx::umma_desc_sw128(f_base);
      auto fdesc = [&](int step) { return fdesc0 + (uint64_t)((step >> 2) * (F_SLAB >> 4) + 2 * (step & 3)); };
      int grp = grp0, tp = tp0;
      for (int it = 0; it < n_my; ++it, ++tc, advance(grp, tp)) {
        if (it == 0 || tp == 0) {
          ptx::mbar_wait(ffull, fph);
          fph ^= 1;
        }
        ptx::mbar_wait(dempty, dph ^ 1);   // both CTAs' epilogues have read all of the previous tile's D
        dph ^= 1;
        ptx::tc_fence_after();
        const bool last_of_group = (it == n_my - 1) || (tp == p.n_tp - 1);
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          const uint32_t taddr = tmem_base + c * NP;
#pragma unroll
          for (int i = 0; i < NSLABS; ++i) {
            ptx::mbar_wait(afull(astage), aphase);
            ptx::tc_fence_after();
            const uint64_t adesc = adesc0 + (uint64_t)(astage * (A_SLAB >> 4));
            if (ptx::elect_one()) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                const int g = i * (BK / 16) + j;   // K16 step inside [hi | lo], compile-time after unrolling
                if (g < KH16) {                    // basis_hi x (feat_hi + feat_lo)
                  ptx::mma_f16_ss_2sm(taddr, adesc + 2 * j, fdesc(g), IDESC_BLEND, g != 0 ? 1u : 0u);
                  ptx::mma_f16_ss_2sm(taddr, adesc + 2 * j, fdesc(KH16 + g), IDESC_BLEND, 1u);
                } else if (g < 2 * KH16) {         // basis_lo x feat_hi
                  ptx::mma_f16_ss_2sm(taddr, adesc + 2 * j, fdesc(g - KH16), IDESC_BLEND, 1u);
                }
              }
              ptx::mma_commit_2sm_mcast(aempty(astage), CMASK);
              if (c == 2 && i == NSLABS - 1) {
                ptx::mma_commit_2sm_mcast(dfull, CMASK);
                if (last_of_group) ptx::mma_commit_2sm_mcast(fempty, CMASK);
              }
            }
            __syncwarp();
            if (++astage == ASTAGES) { astage = 0; aphase ^= 1; }
          }
        }
        const uint32_t wb = tc & 1;
        ptx::mbar_wait(wfull(wb), (tc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t wa = w_base + wb * JSLABS * A_SLAB;
        auto wdesc = [&](int step) { return ptx::umma_desc_sw128(wa + (step >> 2) * A_SLAB) + 2 * (step & 3); };
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch, ++cc) {
          const uint32_t buf = cc % NT;
          ptx::mbar_wait(tempty(buf), ((cc / NT) & 1) ^ 1);
          ptx::mbar_wait(sfull(sstage), sphase);
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + T_COL0 + buf * NS;
          const uint32_t sa = s_base + sstage * JSLABS * S_SLAB;
          auto sdesc = [&](int step) { return ptx::umma_desc_sw128(sa + (step >> 2) * S_SLAB) + 2 * (step & 3); };
          if (ptx::elect_one()) {
#pragma unroll
            for (int g = 0; g < JS; ++g) {   // w_hi x (A_hi + A_lo), then w_lo x A_hi
              ptx::mma_f16_ss_2sm(taddr, wdesc(g), sdesc(g), IDESC_SKIN, g ? 1u : 0u);
              ptx::mma_f16_ss_2sm(taddr, wdesc(g), sdesc(JS + g), IDESC_SKIN, 1u);
              ptx::mma_f16_ss_2sm(taddr, wdesc(JS + g), sdesc(g), IDESC_SKIN, 1u);
            }
            ptx::mma_commit_2sm_mcast(sempty(sstage), CMASK);
            ptx::mma_commit_2sm_mcast(tfull(buf), CMASK);
            if (ch == NCH - 1) ptx::mma_commit_2sm_mcast(wempty(wb), CMASK);
          }
          __syncwarp();
          if (++sstage == SSTAGES) { sstage = 0; sphase ^= 1; }
        }
      }
    }
  } else {
    // ---- epilogue: thread = vertex.  T (60 columns) and the chunk's 5 x (x,y,z) of D -> skinned vertex -> store
    ptx::setmaxnreg_inc<EPI_REGS>();
    const int q = warp & 3;
    const int set = (warp - 4) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const size_t pstride = (size_t)p.V * 3;
    const uint32_t l_tempty0 = ptx::mapa(tempty(0), 0), l_dempty = ptx::mapa(dempty, 0);
    float* stg = reinterpret_cast<float*>(smem + L::OFF_STG) + (warp - 4) * 2 * 96;
    uint32_t dph = 0, tc = 0, sflip = 0;
    int grp = grp0, tp = tp0;
    for (int it = 0; it < n_my; ++it, ++tc, advance(grp, tp)) {
      const int tile = 2 * tp + (int)crank;
      const int v0 = tile * TILE_V + q * 32;
      const int v = v0 + lane;
      const int n_floats = max(0, min(32, p.V - v0)) * 3;
      ptx::mbar_wait(dfull, dph);
      dph ^= 1;
      ptx::tc_fence_after();
#pragma unroll 1
      for (int ch = set; ch < NCH; ch += NSETS) {
        const uint32_t cc = tc * NCH + ch, buf = cc % NT;
        ptx::mbar_wait(tfull(buf), (cc / NT) & 1);
        ptx::tc_fence_after();
        uint32_t t[64], dx[8], dy[8], dz[8];
        const uint32_t d0 = tmem_base + lane_addr + ch * CP;
        ptx::tmem_ld_32x64(tmem_base + lane_addr + T_COL0 + buf * NS, t);
        ptx::tmem_ld_32x8(d0, dx);
        ptx::tmem_ld_32x8(d0 + NP, dy);
        ptx::tmem_ld_32x8(d0 + 2 * NP, dz);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_cluster(l_tempty0 + 8u * buf);
          if (ch + NSETS >= NCH) ptx::mbar_arrive_cluster(l_dempty);   // this warp's last read of the tile's D
        }
        const int64_t b0 = (int64_t)grp * NP + ch * CP;
        int n_ok = NP - ch * CP;                                        // poses of this chunk inside the group ...
        if (n_ok > CP) n_ok = CP;
        if (b0 + n_ok > p.B) n_ok = (int)max((int64_t)0, p.B - b0);     // ... and inside the batch
        if (STAGED) {
          float* dst = p.verts + (size_t)b0 * pstride + (size_t)v0 * 3;
#pragma unroll
          for (int i = 0; i < CP; ++i) {
            if (i < n_ok) {
              const float* T = reinterpret_cast<const float*>(t) + i * 12;
              const float x = __uint_as_float(dx[i]), y = __uint_as_float(dy[i]), z = __uint_as_float(dz[i]);
              float* sg = stg + sflip * 96;
              sflip ^= 1;
              sg[3 * lane + 0] = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[9])));
              sg[3 * lane + 1] = fmaf(T[3], x, fmaf(T[4], y, fmaf(T[5], z, T[10])));
              sg[3 * lane + 2] = fmaf(T[6], x, fmaf(T[7], y, fmaf(T[8], z, T[11])));
              __syncwarp();
              float* w = dst + (size_t)i * pstride;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const int f = 32 * k + lane;
                if (f < n_floats) w[f] = sg[f];
              }
            }
          }
        } else if (v < p.V) {
          // each lane stores its vertex's 12 bytes; the warp's 32 records are one contiguous 384-byte run
          float* dst = p.verts + (size_t)b0 * pstride + (size_t)v * 3;
#pragma unroll
          for (int i = 0; i < CP; ++i) {
            if (i < n_ok) {
              const float* T = reinterpret_cast<const float*>(t) + i * 12;
              const float x = __uint_as_float(dx[i]), y = __uint_as_float(dy[i]), z = __uint_as_float(dz[i]);
              float* w = dst + (size_t)i * pstride;
              w[0] = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[9])));
              w[1] = fmaf(T[3], x, fmaf(T[4], y, fmaf(T[5], z, T[10])));
              w[2] = fmaf(T[6], x, fmaf(T[7], y, fmaf(T[8], z, T[11])));
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();       // no CTA leaves while the peer may still arrive on its barriers / read its operands
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace lt2

bool lbs_fused2_fits(const dpb_lbs* h, const LbsVariant& v) {
  if (!h->tc_ready || !v.dirs16) return false;
  if (v.kext != 448) return false;                      // instantiated for a 224-wide blend K (SMPL; SMPL-X const tail)
  if (h->J >= h->jp) return false;                      // needs the spare joint slot for the translation
  if (h->jp != 32 && h->jp != 64) return false;
  if ((h->n_cols_pad / lt2::TILE_V) % 2) return false;  // tile pairs
  return true;
}

// verts[B,V,3] = skinned vertices.  featop [B_pad, kext] and skinop [B_pad*12, 2*jp] were written by the pose kernel.
int lbs_fused2(dpb_lbs* h, const LbsVariant& v, __half* featop, __half* skinop, float* verts, int64_t B,
               cudaStream_t st) {
  const int K2 = v.kext, Jp = h->jp;
  const int64_t B_pad = (B + 127) / 128 * 128;          // rows that physically exist (lbs_tc_ws_bytes); TMA zero-fills beyond
  CUtensorMap tm_feat, tm_s;
  int rc = make_tmap_2d(&tm_feat, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, featop, K2, (uint64_t)B_pad, lt2::BK, lt2::NP / 2, 2);
  if (rc == DPB_OK)
    rc = make_tmap_2d(&tm_s, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, skinop, 2 * Jp, (uint64_t)B_pad * 12, lt2::BK, lt2::NS / 2, 2);
  if (rc != DPB_OK) return rc;
  lt2::Params p{};
  p.V = h->V;
  p.V_pad = h->n_cols_pad;
  p.n_tp = h->n_cols_pad / lt2::TILE_V / 2;
  p.B = B;
  p.n_items = (long long)((B + lt2::NP - 1) / lt2::NP) * p.n_tp;
  p.verts = verts;
  const int staged = getenv("DPB_LBS_STAGED") ? atoi(getenv("DPB_LBS_STAGED")) : 1;   // A/B timing only (read per call)
  void (*kern)(lt2::Params, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap) = nullptr;
  size_t smem = 0;
  if (Jp == 32) {
    kern = staged ? lt2::lbs_fused2_kernel<14, 7, 2, 1, 6, 4, true> : lt2::lbs_fused2_kernel<14, 7, 2, 1, 6, 4, false>;
    smem = lt2::Smem<1, 6, 4>::BYTES;
  } else {
    kern = staged ? lt2::lbs_fused2_kernel<14, 7, 4, 2, 4, 4, true> : lt2::lbs_fused2_kernel<14, 7, 4, 2, 4, 4, false>;
    smem = lt2::Smem<2, 4, 4>::BYTES;
  }
  DPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = h->sm_count & ~1;                                   // whole CTA pairs
  const long long max_pairs = p.n_items < 1 ? 1 : p.n_items;
  if (grid / 2 > max_pairs) grid = (int)(2 * max_pairs);
  kern<<<grid, lt2::NUM_THREADS, smem, st>>>(p, v.tm_dirs, tm_feat, h->tm_wop, tm_s);
  DPB_CUDA_CHECK(cudaGetLastError());
  return DPB_OK;
}

}  // namespace dpb
