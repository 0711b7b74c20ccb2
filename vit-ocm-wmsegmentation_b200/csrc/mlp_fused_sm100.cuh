// Fused transformer MLP for sm_100a (embed dim D <= 384, inference engines with single 16-bit operands):
//     X[M, D] += gelu(XN[M, D] . W1[Hd, D]^T + b1) . W2[D, Hd]^T + b2
// Replaces Mlp.forward (SSS/dino/vision_transformer.py:57-63: fc1 -> GELU -> fc2, dropout p = 0) together with the residual
// add of Block.forward (:111).  The hidden activations [M, Hd] never leave the SM: the separate fc1 / fc2 GEMMs wrote and
// re-read them through HBM (6 KB per token row and block at ViT-S -- a quarter of the whole forward's HBM traffic), and
// the fc1 kernel was bound by its GELU epilogue, not by the tensor core.  Here the GELU of hidden chunk c runs on the
// epilogue warps while the tensor core works on fc1 of chunk c + 1 and fc2 of chunk c - 1.
//
// One work item = a 256-row tile owned by a CTA pair (cta_group::2: each CTA holds its own 128 rows of every A operand
// and half of every B operand, so weights cross the L2 -> SM link once per 256 rows).  Per CTA:
//   A     = XN tile [128 x D] bf16, resident for the whole item (KB1 = D / 64 SWIZZLE_128B k-blocks, TMA)
//   for each chunk c of 128 hidden columns:
//     fc1:  Hacc[128 x 128] (TMEM)  = A . W1[c]^T                         KB1 x 4 MMAs  (M 256, N 128, K 16)
//     GELU: Hacc -> registers -> + b1 -> gelu -> 16-bit -> smem H[128 x 128] (two SWIZZLE_128B k-blocks: fc2's A operand)
//     fc2:  OUT[128 x D] (TMEM)   += H . W2[:, c]^T                       2 x NP2 x 4 MMAs  (M 256, N D/NP2 = 192 at D = 384)
//   OUT + b2 -> TMA reduce-add into the fp32 residual stream (performed by the L2)
// Weights stream through a ring of four 16 KB slots, in the order the MMA thread consumes them: fc1(0), fc1(1), fc2(0), fc1(2),
// fc1(3), fc2(1), fc1(4), fc2(2), ...  A slot holds two fc1 k-blocks ([64 rows x 64 k] per CTA = its half of a [128 x 64] B tile) or one D/NP2-wide
// output part of one fc2 k-block: ONE barrier wait and ONE commit per 4-8 MMAs.  (With one 8 KB tile per slot the issuing thread
// spent ~290 clk per slot on the wait / commit round trip against 256 clk of tensor work: the kernel was MMA-issue bound --
// profiles/r02_mlp_fused.txt.)  TMEM: OUT in columns [0, D), Hacc in [384, 512).
// Warp roles: 0 = weight TMA, 1 = MMA issuer (leader CTA), 2 = TMEM allocator, 3 = A-tile TMA, 4.. = 16 epilogue warps in TWO
// GROUPS of 8 (lane quadrant = warp % 4, column half = ((warp - 4) / 4) % 2, group = (warp - 4) / 8).  The groups take the hidden
// chunks in turn (group = chunk parity); each has its OWN gelu buffer H[g] in shared memory and its own set of barriers, and fc1
// runs TWO chunks ahead of fc2.  History (profiles/r02_mlp_fused.txt): with all 16 warps on the same chunk and one H buffer the
// kernel ran at ~4 700 clk per chunk against 3 072 clk of tensor work, with OR without the MMAs and with OR without the GELU
// arithmetic -- a latency chain (TMEM read, arithmetic, wait for fc2 of the previous chunk to release H, stores, proxy fence,
// cluster-scope hand-over, fc2, commit), not a throughput limit.  Two groups on one H buffer still serialised on it (stores of
// chunk n + 1 only after fc2(n) retired, fc2(n + 1) only after the stores: ~4 300 clk).  With a buffer per group the stores of chunk
// n + 1 overlap fc2(n), and the other group's arithmetic fills the pipes while one group is in the latency-bound part of its chunk.
// The group that handled an item's LAST chunk drains OUT (32 x 32 fp32 boxes inside its own H buffer) while the
// other group already works on the next item's first chunk.
#pragma once
#include "gemm_sm100.cuh"

namespace vitocm {

struct MlpArgs {
  int M;               // token rows
  int hidden;          // Hd (multiple of 128)
  int f16;             // 16-bit operand format: 0 = bf16, 1 = IEEE fp16
  int gelu_mode;       // 0 = three-coefficient sigmoid form (bf16 engines), 2 = five-coefficient sigmoid form (fp16 engines)
  const float* bias1;  // [Hd]
  const float* bias2;  // [D]
  // diagnostics (vitocm_mlp_fused_timeline) or nullptr: 64 SM-clock stamps (low 32 bits) of the leader CTA of cluster 0 on its work
  // item `timeline_item`, collected in shared memory (global stores would sit in front of the cluster-scope barrier arrives) and
  // written out when the kernel ends: [3 c + e], c < 12: epilogue warp 0, e = 0 fc1(c) complete, 1 gelu arithmetic done, 2 gelu(c)
  // handed over; [36 + 2 c + e]: MMA thread, e = 0 fc1(c) issued, 1 gelu(c) available; [60] item start, [61] OUT complete,
  // [62] item epilogue done
  long long* timeline;
  int timeline_item;
  // Start stagger: every cluster does the same work in the same time, so without it all of them write their 128 x D output
  // tiles in the same instant (a burst the L2 absorbs at ~1/3 of the kernel's average rate) and fetch the same weights in the
  // same instant.  Clusters with index >= stagger_from -- those that have one work item fewer than the busiest, so the delay
  // costs nothing -- start up to stagger_clk SM clocks late, evenly spread.  stagger_from >= the cluster count disables it.
  int stagger_from;
  int stagger_clk;
  int debug;           // diagnostics (VITOCM_MLP_DEBUG): bit 0 = the MMA thread issues no MMAs (barrier traffic only), bit 1 = the
                       // epilogue skips the GELU arithmetic, bit 2 = the epilogue skips the shared-memory stores of gelu(chunk), bit 3 = the
                       // output drain issues no TMA reduce-adds (TMEM reads, shared-memory stores and fences only)
};
__device__ __forceinline__ void mlp_stamp(bool on, uint32_t smem_tl, int idx) {
  if (on) asm volatile("{\n\t.reg .b32 t;\n\tmov.u32 t, %%clock;\n\tst.shared.b32 [%0], t;\n\t}" ::"r"(smem_tl + 4u * idx) : "memory");
}

constexpr int MLP_HC = 128;                 // hidden columns per chunk
constexpr int MLP_SLOT_BYTES = 2 * 64 * 64 * 2;   // one ring slot: two [64 rows x 64 k] fc1 weight tiles or one fc2 output part (16 KB), one mbarrier round trip
constexpr int MLP_KB_BYTES = 128 * 64 * 2;  // one [128 x 64] k-block of an A operand
constexpr int MLP_H_COL = 384;              // TMEM column of the hidden-chunk accumulator
constexpr int MLP_MAX_SLOTS = 4;

constexpr int MLP_EW = 16;                  // epilogue warps: two groups of 8 (4 lane quadrants x 2 column halves of a chunk)
constexpr int MLP_GW = MLP_EW / 2;          // warps per group
template <int KB1, int CL>
struct MlpCfg {
  static constexpr int EW = MLP_EW;
  static_assert(CL == 2 || CL == 4, "fused MLP: clusters of one or two CTA pairs");
  static constexpr int D = KB1 * 64;
  static constexpr int NP2 = (D + 255) / 256;           // output parts, one fc2 MMA each per k step: N = BN2 = D / NP2 (192 at D = 384)
  static constexpr int BN2 = D / NP2;
  static constexpr int T1 = KB1 % 2 == 0 ? 2 : 1;       // fc1 k-blocks per ring slot
  static constexpr int W2_TILE_BYTES = (BN2 / 2) * 128;  // one CTA's half of a [BN2 x 64] fc2 weight tile: one per ring slot
  static_assert(W2_TILE_BYTES <= MLP_SLOT_BYTES && T1 * 8192 <= MLP_SLOT_BYTES, "fused MLP: ring slot too small");
  static_assert(D % 128 == 0 && D <= 384, "fused MLP: D must be 128, 256 or 384");
  static constexpr int THREADS = (4 + EW) * 32;
  static constexpr int A_BYTES = KB1 * MLP_KB_BYTES;
  static constexpr int H_BYTES = 2 * MLP_KB_BYTES;      // one group's gelu buffer [128 x 128]; between items: its 8 output staging boxes (32 x 32 fp32, SWIZZLE_128B)
  static constexpr int BAR_BYTES = 512;
  static constexpr int FIXED = A_BYTES + 2 * H_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
  static constexpr int SLOTS_FIT = (GEMM_SMEM_LIMIT - FIXED) / MLP_SLOT_BYTES;
  static constexpr int SLOTS = SLOTS_FIT > MLP_MAX_SLOTS ? MLP_MAX_SLOTS : SLOTS_FIT;
  static_assert(SLOTS >= 3, "fused MLP: weight ring too shallow");
  static constexpr int SMEM_BYTES = FIXED + SLOTS * MLP_SLOT_BYTES;
  // setmaxnreg budgets, EW = 16 only: the pool is what the CTA was launched with (640 threads x 96 registers), so
  // 128 x REGS_CTRL + 512 x REGS_EPI <= 640 x 96 -- a larger sum leaves setmaxnreg.inc waiting for ever.  An epilogue thread holds 64
  // accumulator values of its row per chunk.
  static constexpr int REGS_CTRL = 48;
  static constexpr int REGS_EPI = 104;
  static_assert(4 * REGS_CTRL + EW * REGS_EPI <= (4 + EW) * 96, "fused MLP: setmaxnreg budgets exceed the launch allocation");
};

template <int KB1, int CL>
__global__ void __launch_bounds__(MlpCfg<KB1, CL>::THREADS, 1)
mlp_fused_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
                         const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_x, const MlpArgs args) {
  using Cfg = MlpCfg<KB1, CL>;
  constexpr int D = Cfg::D, NP2 = Cfg::NP2, BN2 = Cfg::BN2, T1 = Cfg::T1, SLOTS = Cfg::SLOTS, EW = Cfg::EW;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_h = smem_a + Cfg::A_BYTES;
  const uint32_t smem_w = smem_h + 2 * Cfg::H_BYTES;   // smem_h: H[0] | H[1]
  const uint32_t bars = smem_w + SLOTS * MLP_SLOT_BYTES;
  const uint32_t w_full = bars;                       // [SLOTS] TMA -> MMA (leader's copy counts both CTAs' bytes)
  const uint32_t w_empty = bars + 8 * MLP_MAX_SLOTS;  // [SLOTS] MMA -> TMA (both CTAs)
  const uint32_t a_full = bars + 16 * MLP_MAX_SLOTS;  // A tile landed (leader's copy)
  const uint32_t a_empty = a_full + 8;                // last fc1 of the item retired (both CTAs)
  // per epilogue group g (= chunk parity): every barrier below completes once per chunk of that group
  const uint32_t h_full = a_full + 16;                // [2] fc1 chunk complete in TMEM (both CTAs)
  const uint32_t h_tmem_empty = a_full + 32;          // [2] the group's warps of both CTAs read the chunk out of TMEM (leader's copy)
  const uint32_t h_smem_full = a_full + 48;           // [2] the group's warps of both CTAs wrote gelu(chunk) to smem (leader's copy)
  const uint32_t h_smem_empty = a_full + 64;          // [2] fc2 of the chunk retired (both CTAs)
  const uint32_t out_full = a_full + 80;              // [2] last fc2 of an item drained by group g retired (both CTAs)
  const uint32_t out_empty = a_full + 96;             // [2] group g's warps of both CTAs read OUT out of TMEM (leader's copy)
  const uint32_t tmem_ptr_smem = a_full + 112;
  const uint32_t smem_tl = bars + 256;                // [64] diagnostics stamps

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(ptx::cluster_ctarank());
  const int rank = crank & 1;            // rank inside the CTA pair (0 = leader: issues the MMAs)
  const int pairid = crank >> 1;         // pair inside the cluster (CL == 4: two pairs share every weight slot by multicast)
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pairid));
  const uint16_t all_mask = CL == 4 ? 0xF : 0x3;
  // work items: 256-row tiles; cluster k walks tile groups k, k + gridDim / CL, ... (CL / 2 consecutive tiles per group, one per pair:
  // the pairs of a cluster run in lockstep on the shared weight ring, so a pair whose tile lies beyond M still goes through the
  // motions -- its A rows are zero-filled and its stores clipped by the tensor maps)
  const int tiles_m = (args.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int groups = (tiles_m + CL / 2 - 1) / (CL / 2);
  const int group0 = static_cast<int>(blockIdx.x) / CL;
  const int gstep = static_cast<int>(gridDim.x) / CL;
  const int NC = args.hidden / MLP_HC;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w1);
    ptx::prefetch_tmap(&tmap_w2);
    ptx::prefetch_tmap(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SLOTS; ++s) {
      ptx::mbar_init(w_full + 8 * s, 1);
      ptx::mbar_init(w_empty + 8 * s, CL / 2);   // one commit per pair sharing the slot
    }
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(h_full + 8 * g, 1);
      ptx::mbar_init(h_tmem_empty + 8 * g, 2 * MLP_GW);
      ptx::mbar_init(h_smem_full + 8 * g, 2 * MLP_GW);
      ptx::mbar_init(h_smem_empty + 8 * g, 1);
    }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(out_full + 8 * g, 1);
      ptx::mbar_init(out_empty + 8 * g, 2 * MLP_GW);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_ptr_smem, 512);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = ptx::lds_u32(tmem_ptr_smem);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_CTRL));
    if (warp == 0) {
      // ===================== weight TMA producer (both CTAs: each stages its half of every B tile) =====================
      if (lane == 0) {
        // start stagger (see MlpArgs::stagger_clk): everything downstream waits for the first weights
        if (group0 >= args.stagger_from) {
          const long long wait_clk = static_cast<long long>(args.stagger_clk) * (group0 - args.stagger_from + 1) / (gstep - args.stagger_from + 1);
          const long long t0 = clock64();
          while (clock64() - t0 < wait_clk) {}
        }
        int slot = 0;
        uint32_t phase = 0;
        int n = 0;   // loads so far: the pairs of a 4-CTA cluster take turns issuing the multicast load of a slot
        // one ring slot = `ntiles` boxes: acquire it, arm the leader's barrier with both CTAs' bytes, then the boxes (CL == 4: only the
        // pair whose turn it is issues them, multicast to the CTA of the same rank in the other pair)
        auto acquire = [&](int bytes) {
          ptx::mbar_wait(w_empty + 8 * slot, phase ^ 1, 31);   // every pair sharing the slot has consumed it
          if (rank == 0) ptx::mbar_arrive_expect_tx(w_full + 8 * slot, 2 * bytes);
          return CL == 2 || (n & 1) == pairid;
        };
        auto box = [&](const CUtensorMap* tm, int off, int c0, int c1) {
          if (CL == 2) ptx::tma_load_2d_2cta(smem_w + slot * MLP_SLOT_BYTES + off, tm, w_full + 8 * slot, c0, c1);
          else ptx::tma_load_2d_2cta_mc(smem_w + slot * MLP_SLOT_BYTES + off, tm, w_full + 8 * slot, c0, c1, static_cast<uint16_t>(5u << rank));
        };
        auto release = [&]() {
          ++n;
          if (++slot == SLOTS) { slot = 0; phase ^= 1; }
        };
        auto load_fc1 = [&](int c) {
          for (int kb = 0; kb < KB1; kb += T1) {
            if (acquire(T1 * 8192))
              for (int kk = 0; kk < T1; ++kk) box(&tmap_w1, kk * 8192, (kb + kk) * GEMM_BK, c * MLP_HC + rank * 64);
            release();
          }
        };
        auto load_fc2 = [&](int c) {
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int np = 0; np < NP2; ++np) {
              if (acquire(Cfg::W2_TILE_BYTES)) box(&tmap_w2, 0, c * MLP_HC + kb2 * GEMM_BK, np * BN2 + rank * (BN2 / 2));
              release();
            }
        };
        // the order the MMA thread consumes them in: fc1 runs two chunks ahead of fc2
        for (int grp = group0; grp < groups; grp += gstep) {
          load_fc1(0);
          if (NC > 1) load_fc1(1);
          load_fc2(0);
          if (NC > 2) load_fc1(2);
          for (int c = 1; c < NC; ++c) {
            if (c + 2 < NC) load_fc1(c + 2);
            load_fc2(c);
          }
        }
      }
    } else if (warp == 3) {
      // ===================== A-tile TMA producer =====================
      if (lane == 0) {
        int t = 0;
        for (int grp = group0; grp < groups; grp += gstep, ++t) {
          const int tile = grp * (CL / 2) + pairid;
          ptx::mbar_wait(a_empty, (t & 1) ^ 1, 32);   // the previous item's fc1 MMAs no longer read the tile
          if (rank == 0) ptx::mbar_arrive_expect_tx(a_full, 2 * Cfg::A_BYTES);
          for (int kb = 0; kb < KB1; ++kb)
            ptx::tma_load_2d_2cta(smem_a + kb * MLP_KB_BYTES, &tmap_a, a_full, kb * GEMM_BK, tile * 2 * GEMM_BM + rank * GEMM_BM);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (leader CTA) =====================
      if (rank == 0 && ptx::elect_one()) {
        const uint32_t idesc = ptx::make_idesc(2 * GEMM_BM, MLP_HC, false, false, args.f16 ? 0u : 1u);    // fc1: N = one hidden chunk
        const uint32_t idesc2 = ptx::make_idesc(2 * GEMM_BM, BN2, false, false, args.f16 ? 0u : 1u);      // fc2: N = one output part
        const uint64_t a_desc0 = ptx::make_smem_desc_sw128(smem_a, 1024, 0);
        const uint64_t h_desc0 = ptx::make_smem_desc_sw128(smem_h, 1024, 0);
        const uint64_t w_desc0 = ptx::make_smem_desc_sw128(smem_w, 1024, 0);
        const uint32_t h_tmem = tmem_base + MLP_H_COL;
        int slot = 0;
        uint32_t phase = 0;
        int g1 = 0, g2 = 0;   // fc1 / fc2 chunks issued so far (all items): phase counters
        int t = 0;
        bool tl = false;
        auto issue_fc1 = [&](int c) {
          if (g1 > 0) {   // the previous chunk has been read out of the TMEM accumulator (by the group of its parity)
            ptx::mbar_wait(h_tmem_empty + 8 * ((g1 - 1) & 1), ((g1 - 1) >> 1) & 1, 33);
            ptx::tc_fence_after();
          }
#pragma unroll 1
          for (int kb = 0; kb < KB1; kb += T1) {
            ptx::mbar_wait(w_full + 8 * slot, phase, 34);
            ptx::tc_fence_after();
            const uint64_t adesc = ptx::desc_advance(a_desc0, kb * MLP_KB_BYTES);
            const uint64_t bdesc = ptx::desc_advance(w_desc0, slot * MLP_SLOT_BYTES);
            if (!(args.debug & 1)) {
#pragma unroll
              for (int kk = 0; kk < T1; ++kk)
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k)
                  ptx::umma_bf16_ss_2cta(h_tmem, ptx::desc_advance(adesc, kk * MLP_KB_BYTES + k * 32), ptx::desc_advance(bdesc, kk * 8192 + k * 32), idesc,
                                         (kb > 0 || kk > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit_2cta_mask(w_empty + 8 * slot, all_mask);
            if (++slot == SLOTS) { slot = 0; phase ^= 1; }
          }
          ptx::umma_commit_2cta_mask(h_full + 8 * (g1 & 1), pair_mask);
          if (c < 12) mlp_stamp(tl, smem_tl, 36 + 2 * c);
          if (c + 1 == NC) ptx::umma_commit_2cta_mask(a_empty, pair_mask);   // last fc1 of the item: the A tile may be replaced once it retires
          ++g1;
        };
        int drains[2] = {0, 0};   // items drained so far by each epilogue group
        int pend_g = 0, pend_par = 0;   // the previous item's drain: group and parity of its out_empty completion
        auto issue_fc2 = [&](int c) {
          ptx::mbar_wait_cluster(h_smem_full + 8 * (g2 & 1), (g2 >> 1) & 1, 35);   // gelu(chunk) sits in both CTAs' shared memory
          ptx::tc_fence_after();
          if (c < 12) mlp_stamp(tl, smem_tl, 36 + 2 * c + 1);
          if (c == 0 && t > 0) {   // OUT still holds the previous item until both CTAs' draining warps have read it
            ptx::mbar_wait(out_empty + 8 * pend_g, pend_par, 36);
            ptx::tc_fence_after();
          }
          const uint64_t hdesc = ptx::desc_advance(h_desc0, (g2 & 1) * Cfg::H_BYTES);   // this chunk's group's buffer
#pragma unroll 1
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            const uint64_t adesc = ptx::desc_advance(hdesc, kb2 * MLP_KB_BYTES);
#pragma unroll 1
            for (int np = 0; np < NP2; ++np) {
              ptx::mbar_wait(w_full + 8 * slot, phase, 37);
              ptx::tc_fence_after();
              const uint64_t bdesc = ptx::desc_advance(w_desc0, slot * MLP_SLOT_BYTES);
              if (!(args.debug & 1)) {
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k)
                  ptx::umma_bf16_ss_2cta(tmem_base + np * BN2, ptx::desc_advance(adesc, k * 32), ptx::desc_advance(bdesc, k * 32), idesc2,
                                         (c > 0 || kb2 > 0 || k > 0) ? 1u : 0u);
              }
              ptx::umma_commit_2cta_mask(w_empty + 8 * slot, all_mask);
              if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
          }
          ptx::umma_commit_2cta_mask(h_smem_empty + 8 * (g2 & 1), pair_mask);
          ++g2;
        };
        for (int grp = group0; grp < groups; grp += gstep, ++t) {
          tl = args.timeline != nullptr && blockIdx.x == 0 && t == args.timeline_item;
          ptx::mbar_wait(a_full, t & 1, 38);
          ptx::tc_fence_after();
          mlp_stamp(tl, smem_tl, 60);
          // fc1 two chunks ahead of fc2 (one chunk per epilogue group in flight), except at the start of an item: fc1(2) needs the
          // group that is still draining the previous item (it reads chunk 1 afterwards), fc2(0) only that drain's TMEM reads
          issue_fc1(0);
          if (NC > 1) issue_fc1(1);
          issue_fc2(0);
          if (NC > 2) issue_fc1(2);
          for (int c = 1; c < NC; ++c) {
            if (c + 2 < NC) issue_fc1(c + 2);
            issue_fc2(c);
          }
          // the group that handled the item's last chunk drains it
          pend_g = (g2 - 1) & 1;
          pend_par = drains[pend_g] & 1;
          ++drains[pend_g];
          ptx::umma_commit_2cta_mask(out_full + 8 * pend_g, pair_mask);
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_EPI));
    const int ew = warp - 4;
    const int q = warp & 3;          // TMEM lane quadrant
    const int ch = (ew >> 2) & 1;    // column half of a chunk (GELU) / of OUT (drain)
    const int grp = ew >> 3;         // epilogue group: takes the chunks whose global index has this parity
    constexpr int CPW = MLP_HC / 2;          // hidden columns per warp and chunk
    constexpr int NSUB = CPW / 32;
    constexpr int OCW = D / 2;               // output columns per draining warp
    constexpr int NOS = OCW / 32;            // 32-column steps of the drain
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int row = q * 32 + lane;   // row of the CTA's 128
    const uint32_t my_h = smem_h + grp * Cfg::H_BYTES;   // this group's gelu buffer
    // output staging: between items that buffer holds one 32 x 32 fp32 box (SWIZZLE_128B) for each warp of the group
    const uint32_t box = my_h + (ew & 7) * 4096;
    int my_drains = 0;
    const uint32_t my_h_full = h_full + 8 * grp, my_tmem_empty = h_tmem_empty + 8 * grp;
    const uint32_t my_smem_full = ptx::mapa(h_smem_full + 8 * grp, crank & ~1);
    int n0 = 0;   // global index of the item's first chunk (all items of this CTA)
    int t = 0;
    for (int item = group0; item < groups; item += gstep, ++t, n0 += NC) {
      const int tile = item * (CL / 2) + pairid;
      const bool tl = args.timeline != nullptr && blockIdx.x == 0 && t == args.timeline_item && (ew & 7) == 0 && lane == 0;
      for (int c = (n0 + grp) & 1; c < NC; c += 2) {
        const int n = n0 + c;          // (n & 1) == grp
        ptx::mbar_wait(my_h_full, (n >> 1) & 1, 40);
        ptx::tc_fence_after();
        if (c < 12) mlp_stamp(tl, smem_tl, 3 * c);
        uint32_t r[NSUB][32];
#pragma unroll
        for (int s = 0; s < NSUB; ++s) ptx::tmem_ld_32x32b_x32(lane_taddr + MLP_H_COL + ch * CPW + s * 32, r[s]);
#pragma unroll
        for (int s = 0; s < NSUB; ++s) ptx::tmem_ld_wait(r[s]);
        // the chunk is in registers: the accumulator goes back to the MMA thread (fc1 of the next chunk)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_leader(my_tmem_empty);
        uint32_t pk[NSUB][16];
#pragma unroll
        for (int s = 0; s < NSUB; ++s) {
          const float* bp = args.bias1 + c * MLP_HC + ch * CPW + s * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);   // same address in every lane
            v[4 * j] = __uint_as_float(r[s][4 * j]) + b4.x;
            v[4 * j + 1] = __uint_as_float(r[s][4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(r[s][4 * j + 2]) + b4.z;
            v[4 * j + 3] = __uint_as_float(r[s][4 * j + 3]) + b4.w;
          }
          if (args.debug & 2) {
          } else if (args.gelu_mode == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_sigmoid_x2(v[j], v[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_sigmoid5_x2(v[j], v[j + 1]);
          }
          if (args.f16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[s][j] = ptx::pack_f16x2(v[2 * j], v[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[s][j] = ptx::pack_bf16x2(v[2 * j], v[2 * j + 1]);
          }
        }
        if (args.timeline != nullptr) {   // diagnostics: pin the arithmetic above the stamp
#pragma unroll
          for (int s = 0; s < NSUB; ++s)
#pragma unroll
            for (int j = 0; j < 16; ++j) asm volatile("" : "+r"(pk[s][j]));
        }
        if (c < 12) mlp_stamp(tl, smem_tl, 3 * c + 1);
        // this group's H buffer is free once fc2 of its previous chunk has retired
        if (n > 1) ptx::mbar_wait(h_smem_empty + 8 * grp, ((n >> 1) - 1) & 1, 41);
        if (!(args.debug & 4)) {
          // this thread's 64 columns are one full 128-byte row of k-block `ch` of the H tile
          const uint32_t tile_addr = my_h + ch * MLP_KB_BYTES + row * 128;
#pragma unroll
          for (int s = 0; s < NSUB; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ptx::sts_v4(tile_addr + (((s * 4 + j) ^ (row & 7)) << 4), pk[s][4 * j], pk[s][4 * j + 1], pk[s][4 * j + 2], pk[s][4 * j + 3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote(my_smem_full);
        if (c < 12) mlp_stamp(tl, smem_tl, 3 * c + 2);
      }
      // ---- item epilogue: OUT + b2 -> fp32 boxes -> TMA reduce-add into the residual stream, by the group that handled the item's
      // LAST chunk (its H buffer is free once that chunk's fc2 -- the item's last MMA -- has retired); the other group goes straight
      // on to the next item's first chunk
      if (grp == ((n0 + NC - 1) & 1)) {
        ptx::mbar_wait(out_full + 8 * grp, my_drains & 1, 42);
        ++my_drains;
        ptx::tc_fence_after();
        mlp_stamp(tl, smem_tl, 61);
        const int row_g = tile * 2 * GEMM_BM + rank * GEMM_BM + q * 32;
        // 32-column boxes (128-byte rows): the TMA unit's reduce-add runs at a fixed rate per box ROW (~4-5 clk), so 64-byte rows
        // (double-buffered 32 x 16 boxes) made the drain slower, not faster (11.4 k against 8.5 k clk per item)
        const int sw = lane & 7;
#pragma unroll 1
        for (int s = 0; s < NOS; ++s) {
          const int col = ch * OCW + s * 32;
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_taddr + col, r);
          float bv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias2 + col) + j);
            bv[4 * j] = b4.x; bv[4 * j + 1] = b4.y; bv[4 * j + 2] = b4.z; bv[4 * j + 3] = b4.w;
          }
          ptx::tmem_ld_wait(r);
          if (s == NOS - 1) {   // this warp's part of OUT is in registers
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(out_empty + 8 * grp);
          }
          if (lane == 0) ptx::bulk_wait_read0();   // the previous reduce-add has read the box
          __syncwarp();
          const uint32_t rowaddr = box + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::sts_v4(rowaddr + ((j ^ sw) << 4), __float_as_uint(__uint_as_float(r[4 * j]) + bv[4 * j]),
                        __float_as_uint(__uint_as_float(r[4 * j + 1]) + bv[4 * j + 1]), __float_as_uint(__uint_as_float(r[4 * j + 2]) + bv[4 * j + 2]),
                        __float_as_uint(__uint_as_float(r[4 * j + 3]) + bv[4 * j + 3]));
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(args.debug & 8)) {
            ptx::tma_reduce_add_2d(&tmap_x, box, col, row_g);
            ptx::bulk_commit();
          }
        }
        // every staging box inside the H buffer has been read before any warp of this group writes its next chunk there
        if (lane == 0) ptx::bulk_wait_read0();
        __syncwarp();
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(MLP_GW * 32) : "memory");
        mlp_stamp(tl, smem_tl, 62);
      }
    }
    if (lane == 0) ptx::bulk_wait_all0();   // global writes complete before the CTA exits
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (args.timeline != nullptr && blockIdx.x == 0 && threadIdx.x < 64) args.timeline[threadIdx.x] = ptx::lds_u32(smem_tl + 4 * threadIdx.x);
  ptx::cluster_sync_all();   // no CTA leaves while the peer could still address its shared memory / TMEM
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace vitocm
