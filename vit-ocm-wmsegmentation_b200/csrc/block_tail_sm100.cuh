// The second half of a transformer block in ONE kernel for sm_100a (embed dim D <= 384, inference engines with single 16-bit
// operands):
//     X  += CTX . Wp^T + bp                                   attention output projection + residual   (vit.py:88, :110)
//     XN2 = LayerNorm(X; ln2)                                 norm2                                    (vit.py:111)
//     X  += gelu(XN2 . W1^T + b1) . W2^T + b2                 Mlp.forward + residual                   (vit.py:57-63, :111)
//     XN  = LayerNorm(X; next block's norm1)                  (optional) the LayerNorm that reads X next (vit.py:107)
// Replaces the proj + LayerNorm GEMM (gemm_sm100.cuh, EPI_RESID_LN), the fused MLP kernel (mlp_fused_sm100.cuh) and the stand-alone
// LayerNorm kernel of the next block.  Why: those three kernels are bound by the residual stream, not by arithmetic -- the proj GEMM
// moved 4.6 KB per token row for 0.3 MFLOP (4.6 TB/s, 320 TFLOP/s), the MLP's reduce-add drain and the LayerNorm pass touched
// the fp32 rows again (profiles/r02_*).  Here a row of X is read ONCE and written ONCE per block, and norm2's output never exists
// in HBM: per row 768 B of CTX + 1536 B of X in, 1536 B of X + 768 B of XN out (4.6 KB against 10.8 KB before).
//
// One work item = a 256-row tile owned by a CTA pair (cta_group::2), as in mlp_fused_sm100.cuh.  Per CTA (its 128 rows):
//   A      = CTX tile [128 x D], 16-bit, by TMA into smem_a
//   proj:    OUT[128 x D] (TMEM, fp32) = A . Wp^T                          KB1 x NP2 x 4 MMAs (M 256, N D / NP2, K 16)
//   ep 1:    y = OUT + bp + X (fp32 rows staged by TMA boxes) -> written BACK to the OUT columns of TMEM; LayerNorm statistics of
//            the row (thread = row = TMEM lane; the four warps sharing a lane quadrant split the columns and exchange (mean, M2)
//            through shared memory, Chan's formula); normalised 16-bit rows -> smem_a, in place of CTX: fc1's A operand
//   chunks:  fc1 -> GELU -> fc2 exactly as in the fused MLP kernel, except that EVERY fc2 MMA accumulates: OUT already holds the
//            residual rows, so the tensor core performs the residual add in fp32 and X never enters shared memory again
//   ep 2:    x = OUT + b2 -> statistics -> fp32 boxes -> TMA store of X;  (x - mean) rstd gamma + beta -> 16-bit boxes -> TMA store
//   qkv:     (optional, TailArgs::n_qkv_chunks > 0) the NEXT block's QKV projection: ep 2 then leaves the normalised rows in smem_a
//            instead of HBM, and QKV[128 x 3D] = XN . Wqkv^T + b runs as 3D / 128 more chunks through the hidden-chunk accumulator
//            (fc1-shaped MMAs; epilogue: bias, 16-bit, the warp's own 4 KB region of the GELU buffer, one TMA store per warp) --
//            the stand-alone QKV GEMM, its launch and the XN round trip through HBM (1.5 KB per row) disappear
// The item boundary is serial (ep 2 of item t, proj and ep 1 of item t + 1 all need the OUT columns: 384 + 128 of 512 TMEM
// columns are in use), so all 16 epilogue warps work on it; between the boundaries the two groups of 8 take the hidden chunks in
// turn as before.  Weights stream through the same ring: proj's [D / NP2 x 64] tiles have the shape of fc2's.
#pragma once
#include "mlp_fused_sm100.cuh"

namespace vitocm {

struct TailArgs {
  int M;               // token rows
  int hidden;          // Hd (multiple of 128)
  int gelu5;           // 1 = five-coefficient sigmoid form of GELU (fp16 engines), 0 = three coefficients (bf16): a RUN-TIME switch on purpose --
                       // the branch keeps the two halves of a chunk in separate basic blocks; as straight-line code ptxas interleaves them
                       // and the chunk loop runs 25 % slower (530 vs 428 us per 175-tile launch)
  const float* bias_p; // [D] proj bias
  const float* ln2_w;  // [D] norm2
  const float* ln2_b;
  const float* bias1;  // [Hd]
  const float* bias2;  // [D]
  const float* lnn_w;  // [D] the LayerNorm that reads X next, or nullptr (then no XN is written)
  const float* lnn_b;
  const float* bias_qkv;   // [n_qkv_chunks * 128] bias of the next block's QKV projection (n_qkv_chunks > 0)
  int n_qkv_chunks;        // 0: XN = LayerNorm(X; lnn) goes to HBM (tmap_xn); 3 D / 128: it stays in shared memory and QKV = XN . Wqkv^T + bias_qkv
                           // is written instead (tmap_wqkv, tmap_qkv); needs lnn_w
  float ln_eps;
  // LayerNorm affine parameters folded into the weights that read the normalised rows (W' = W diag(gamma), b' = b + W beta, prepared by
  // vitocm_finalize_weights): the epilogue then only forms (x - mean) rstd = fma(x, rstd, -mean rstd) -- one instruction per element
  // instead of three and no gamma / beta loads in phases that are instruction-issue bound.
  int fold2;           // norm2 is folded into W1 / bias1 (ln2_w / ln2_b are not read)
  int foldn;           // the next norm1 is folded into Wqkv / bias_qkv (n_qkv_chunks > 0 only; lnn_w is only tested against nullptr)
  // diagnostics (vitocm_block_tail_timeline) or nullptr: 64 SM-clock stamps (low 32 bits) of the leader CTA of pair 0 on its work item
  // `timeline_item`: [3 c + e], c < 5: epilogue warp 0 (group 0) on chunk c as MlpArgs; [36 + 2 c + e], c < 5: MMA thread as MlpArgs;
  // MMA thread: [60] item start (CTX landed), [56] OUT free, proj issued, [59] ep 1 done (mid_ready seen);
  // epilogue warp 0: [57] proj complete, [58] ep 1 pass 1 done, [55] ep 1 handed over, [61] OUT complete, [54] ep 2 statistics known,
  // [62] ep 2 done; [25 + 4 k + e]: the group's k-th QKV chunk (k < 2): accumulator complete, packed, staging free, stored; [46 + c]: MMA thread, QKV chunk c issued; [15 + s] ep 1 pass 1 step s done, [18] ep 1 statistics combined, [19] epilogue warp 0 has handed its rows of the next norm1 over, [63] MMA thread: xn_ready seen, [20 + s] ep 2 pass 2 step s stored, [24] ep 2 stores read
  long long* timeline;
  int timeline_item;
  int debug;           // bit 0: no MMAs (barrier traffic only), bit 1: ep 1 neither loads nor awaits the fp32 rows, bit 2: ep 2 issues no TMA stores,
                       // bit 3: the weight ring runs without its TMA loads (results are garbage; timing only), bit 4: fine stamps of hidden
                       // chunks 4 / 5 instead of the QKV-phase stamps (tools/tail_timeline.py), bit 5: .release.cluster arrives on the
                       // shared-memory hand-overs (as before; A/B)
  // Start stagger: every CTA pair does the same work in the same time, so without it all of them reach the item boundary -- 290 KB
  // of rows in, 290 KB out per CTA -- in the same instant, a burst the L2 serves at a fraction of the kernel's average rate.
  // Pair k starts k / pairs x stagger_clk SM clocks late.
  int stagger_clk;
};

template <int KB1>
struct TailCfg {
  static constexpr int EW = MLP_EW;
  static constexpr int D = KB1 * 64;
  static_assert(D % 128 == 0 && D <= 384, "block tail: D must be 128, 256 or 384");
  static constexpr int NP2 = (D + 255) / 256;           // output parts of proj / fc2: N = BN2 = D / NP2 per MMA
  static constexpr int BN2 = D / NP2;
  static constexpr int T1 = KB1 % 2 == 0 ? 2 : 1;       // fc1 k-blocks per ring slot
  static constexpr int W2_TILE_BYTES = (BN2 / 2) * 128;  // one CTA's half of a [BN2 x 64] proj / fc2 weight tile
  static_assert(W2_TILE_BYTES <= MLP_SLOT_BYTES && T1 * 8192 <= MLP_SLOT_BYTES, "block tail: ring slot too small");
  static constexpr int CW = D / 4;                       // columns of a row per epilogue warp at the item boundary
  static constexpr int NS = CW / 32;                     // 32-column steps per warp
  static constexpr int THREADS = (4 + EW) * 32;
  static constexpr int A_BYTES = KB1 * MLP_KB_BYTES;
  static constexpr int H_BYTES = 2 * MLP_KB_BYTES;       // one group's gelu buffer; at the item boundary: one 4 KB box per warp
  static constexpr int STAGE_A = A_BYTES / EW;            // each epilogue warp's share of smem_a at the item boundary: a 2 KB box (16-bit rows out)
  static_assert(STAGE_A % 1024 == 0 && STAGE_A >= (NS == 1 ? 2048 : 6144), "block tail: staging boxes in smem_a");   // + a 4 KB box (fp32 rows) when NS > 1
  static constexpr int BAR_BYTES = 1024;
  static constexpr int FIXED = A_BYTES + 2 * H_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
  static constexpr int SLOTS_FIT = (GEMM_SMEM_LIMIT - FIXED) / MLP_SLOT_BYTES;
  static constexpr int SLOTS = SLOTS_FIT > MLP_MAX_SLOTS ? MLP_MAX_SLOTS : SLOTS_FIT;
  static_assert(SLOTS >= 3, "block tail: weight ring too shallow");
  static constexpr int SMEM_BYTES = FIXED + SLOTS * MLP_SLOT_BYTES;
  static constexpr int REGS_CTRL = 48;
  static constexpr int REGS_EPI = 104;
  static_assert(4 * REGS_CTRL + EW * REGS_EPI <= (4 + EW) * 96, "block tail: setmaxnreg budgets exceed the launch allocation");
};

// exact (mean, M2) of 32 values.  (Measured and dropped: sum y and sum y^2 in one sweep, M2 = sum y^2 - 32 mean^2 with a fall-back to
// this two-pass form when the subtraction cancels -- two instructions per element instead of three, all tests green, and no change
// of the launch time at all (3 467 / 3 454 / 3 485 against 3 440 / 3 482 / 3 476 us per 1 225-tile launch): the statistics passes wait
// for their TMEM loads, not for issue slots.  profiles/r02_gpu_call_au_fast_stats.log)
__device__ __forceinline__ void tail_stats32(const float (&y)[32], float& mean, float& m2) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 4) { s0 += y[j]; s1 += y[j + 1]; s2 += y[j + 2]; s3 += y[j + 3]; }
  mean = ((s0 + s1) + (s2 + s3)) * (1.0f / 32.0f);
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float d0 = y[j] - mean, d1 = y[j + 1] - mean;
    a = fmaf(d0, d0, a);
    b = fmaf(d1, d1, b);
  }
  m2 = a + b;
}
// Chan: (n_a = 32 s values so far) + (32 new values)
__device__ __forceinline__ void tail_chan_step(int s, float& mean, float& m2, float mean_b, float m2_b) {
  if (s == 0) { mean = mean_b; m2 = m2_b; return; }
  const float na = 32.0f * s, nb = 32.0f, inv = 1.0f / (na + nb);
  const float d = mean_b - mean;
  mean = fmaf(d, nb * inv, mean);
  m2 = m2 + m2_b + d * d * (na * nb * inv);
}

template <int KB1, bool F16>
__global__ void __launch_bounds__(TailCfg<KB1>::THREADS, 1)
block_tail_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a /*CTX*/, const __grid_constant__ CUtensorMap tmap_wp,
                          const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                          const __grid_constant__ CUtensorMap tmap_x /*fp32 rows: 32 x 32 boxes, loads and stores*/,
                          const __grid_constant__ CUtensorMap tmap_xn /*16-bit normalised rows out: 32 x 32 boxes*/,
                          const __grid_constant__ CUtensorMap tmap_wqkv /*next block's QKV weight [3D][D]: 64 x 64 boxes*/,
                          const __grid_constant__ CUtensorMap tmap_qkv /*16-bit QKV rows out: 64 x 32 boxes*/, const TailArgs args) {
  using Cfg = TailCfg<KB1>;
  constexpr int D = Cfg::D, NP2 = Cfg::NP2, BN2 = Cfg::BN2, T1 = Cfg::T1, SLOTS = Cfg::SLOTS, EW = Cfg::EW, CW = Cfg::CW, NS = Cfg::NS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_h = smem_a + Cfg::A_BYTES;
  const uint32_t smem_w = smem_h + 2 * Cfg::H_BYTES;   // smem_h: H[0] | H[1]
  const uint32_t bars = smem_w + SLOTS * MLP_SLOT_BYTES;
  const uint32_t w_full = bars;                       // [SLOTS] TMA -> MMA (leader's copy counts both CTAs' bytes)
  const uint32_t w_empty = bars + 8 * MLP_MAX_SLOTS;  // [SLOTS] MMA -> TMA (both CTAs)
  const uint32_t a_full = bars + 16 * MLP_MAX_SLOTS;  // CTX tile landed (leader's copy)
  const uint32_t a_empty = a_full + 8;                // last fc1 of the item retired (both CTAs)
  const uint32_t h_full = a_full + 16;                // [2] per epilogue group, as in mlp_fused_sm100.cuh
  const uint32_t h_tmem_empty = a_full + 32;          // [2]
  const uint32_t h_smem_full = a_full + 48;           // [2]
  const uint32_t h_smem_empty = a_full + 64;          // [2]
  const uint32_t out_full = a_full + 80;              // last fc2 of the item retired (both CTAs)
  const uint32_t out_empty = a_full + 88;             // every epilogue warp of both CTAs has read OUT for the last time (leader's copy)
  const uint32_t proj_full = a_full + 96;             // proj complete in TMEM (both CTAs)
  const uint32_t mid_ready = a_full + 104;            // ep 1 done in both CTAs: residual rows in TMEM, norm2 rows in smem_a (leader's copy)
  const uint32_t tmem_ptr_smem = a_full + 112;
  const uint32_t xn_ready = a_full + 128;             // ep 2 done in both CTAs with the next norm's rows in smem_a (QKV chunks; leader's copy)
  const uint32_t stage_free = a_full + 120;           // this CTA's epilogue warps no longer use smem_a as staging space (ep 2 stores have been read)
  const uint32_t xbox_bar = bars + 256;               // [EW][2] fp32 row boxes landed (per warp: box in H, box in smem_a)
  const uint32_t smem_tl = bars + 512;                // [64] diagnostics stamps
  static_assert(512 + 256 <= Cfg::BAR_BYTES && 256 + EW * 16 <= 512, "block tail: barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(ptx::cluster_ctarank()) & 1;   // 0 = leader: issues the MMAs
  const uint16_t pair_mask = 3;
  const int tiles_m = (args.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int tile0 = static_cast<int>(blockIdx.x) / 2;
  const int tstep = static_cast<int>(gridDim.x) / 2;
  const int NC = args.hidden / MLP_HC;
  const int NQ = args.n_qkv_chunks;   // QKV chunks behind the MLP chunks of every item (0: none)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_wp);
    ptx::prefetch_tmap(&tmap_w1);
    ptx::prefetch_tmap(&tmap_w2);
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_xn);
    ptx::prefetch_tmap(&tmap_wqkv);
    ptx::prefetch_tmap(&tmap_qkv);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SLOTS; ++s) {
      ptx::mbar_init(w_full + 8 * s, 1);
      ptx::mbar_init(w_empty + 8 * s, 1);
    }
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(h_full + 8 * g, 1);
      ptx::mbar_init(h_tmem_empty + 8 * g, 2 * MLP_GW);
      ptx::mbar_init(h_smem_full + 8 * g, 2 * MLP_GW);
      ptx::mbar_init(h_smem_empty + 8 * g, 1);
    }
    ptx::mbar_init(out_full, 1);
    ptx::mbar_init(out_empty, 2 * EW);
    ptx::mbar_init(proj_full, 1);
    ptx::mbar_init(mid_ready, 2 * EW);
    ptx::mbar_init(stage_free, EW);
    ptx::mbar_init(xn_ready, 2 * EW);
    for (int i = 0; i < 2 * EW; ++i) ptx::mbar_init(xbox_bar + 8 * i, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_ptr_smem, 512);
    ptx::tmem_relinquish_2cta();
  }
  if (warp == 3 && lane == 0 && args.stagger_clk > 0) {   // start stagger (TailArgs::stagger_clk)
    const long long wait_clk = static_cast<long long>(args.stagger_clk) * tile0 / tstep;
    const long long t0 = clock64();
    while (clock64() - t0 < wait_clk) {}
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
        int slot = 0;
        uint32_t phase = 0;
        const bool no_loads = (args.debug & 8) != 0;   // diagnostics: the ring runs without the weight loads (is the chunk loop bound by them?)
        auto acquire = [&](int bytes) {
          ptx::mbar_wait(w_empty + 8 * slot, phase ^ 1, 31);
          if (rank == 0) { if (no_loads) ptx::mbar_arrive(w_full + 8 * slot); else ptx::mbar_arrive_expect_tx(w_full + 8 * slot, 2 * bytes); }
        };
        auto box = [&](const CUtensorMap* tm, int off, int c0, int c1) {
          if (!no_loads) ptx::tma_load_2d_2cta(smem_w + slot * MLP_SLOT_BYTES + off, tm, w_full + 8 * slot, c0, c1);
        };
        auto release = [&]() {
          if (++slot == SLOTS) { slot = 0; phase ^= 1; }
        };
        auto load_proj = [&]() {
          for (int kb = 0; kb < KB1; ++kb)
            for (int np = 0; np < NP2; ++np) {
              acquire(Cfg::W2_TILE_BYTES);
              box(&tmap_wp, 0, kb * GEMM_BK, np * BN2 + rank * (BN2 / 2));
              release();
            }
        };
        auto load_fc1 = [&](int c) {
          for (int kb = 0; kb < KB1; kb += T1) {
            acquire(T1 * 8192);
            for (int kk = 0; kk < T1; ++kk) box(&tmap_w1, kk * 8192, (kb + kk) * GEMM_BK, c * MLP_HC + rank * 64);
            release();
          }
        };
        auto load_fc2 = [&](int c) {
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int np = 0; np < NP2; ++np) {
              acquire(Cfg::W2_TILE_BYTES);
              box(&tmap_w2, 0, c * MLP_HC + kb2 * GEMM_BK, np * BN2 + rank * (BN2 / 2));
              release();
            }
        };
        // the order the MMA thread consumes them in
        for (int tile = tile0; tile < tiles_m; tile += tstep) {
          load_proj();
          load_fc1(0);
          if (NC > 1) load_fc1(1);
          load_fc2(0);
          if (NC > 2) load_fc1(2);
          for (int c = 1; c < NC; ++c) {
            if (c + 2 < NC) load_fc1(c + 2);
            load_fc2(c);
          }
          for (int c = 0; c < NQ; ++c)   // the next block's QKV weight: fc1-shaped chunks of 128 output columns
            for (int kb = 0; kb < KB1; kb += T1) {
              acquire(T1 * 8192);
              for (int kk = 0; kk < T1; ++kk) box(&tmap_wqkv, kk * 8192, (kb + kk) * GEMM_BK, c * MLP_HC + rank * 64);
              release();
            }
        }
      }
    } else if (warp == 3) {
      // ===================== CTX-tile TMA producer =====================
      if (lane == 0) {
        int t = 0;
        for (int tile = tile0; tile < tiles_m; tile += tstep, ++t) {
          ptx::mbar_wait(a_empty, (t & 1) ^ 1, 32);   // the previous item's fc1 MMAs no longer read smem_a
          // (NQ == 0: smem_a is also the previous item's output staging space until its stores have been read, ep 2; with QKV chunks
          // a_empty already lies behind ep 2)
          if (t > 0 && NQ == 0) ptx::mbar_wait(stage_free, (t - 1) & 1, 29);
          if (rank == 0) ptx::mbar_arrive_expect_tx(a_full, 2 * Cfg::A_BYTES);
          for (int kb = 0; kb < KB1; ++kb)
            ptx::tma_load_2d_2cta(smem_a + kb * MLP_KB_BYTES, &tmap_a, a_full, kb * GEMM_BK, tile * 2 * GEMM_BM + rank * GEMM_BM);
          if (tile + tstep < tiles_m) {
            // the next item's tile on its way to L2 (its load is issued late: behind this item's last use of smem_a) -- but not before
            // this item's hidden chunks are done: a whole item ahead, the kernel's own output stream (~100 MB through a 126 MB L2)
            // evicts the prefetched lines again and they are read from HBM twice (ncu: +200 MB of DRAM reads per 175-tile launch)
            ptx::mbar_wait(out_full, t & 1, 26);
            for (int kb = 0; kb < KB1; ++kb) ptx::tma_prefetch_2d(&tmap_a, kb * GEMM_BK, (tile + tstep) * 2 * GEMM_BM + rank * GEMM_BM);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (leader CTA) =====================
      // (ONE issuing thread, inside the elected branch, on purpose.  It spends ~3 000 of a chunk's ~3 700 clk issuing -- it shares its
      // scheduler with four epilogue warps and every ring slot costs ~60 instructions, a third of them R2UR moves.  Measured and dropped:
      // a second issuer for the fc2 chunks on the same weight ring runs more than one ring wrap ahead of the fills of the slots it
      // skips, so its parity waits succeed on the wrong phase -- a hang (two issuers need two rings); the whole warp running the loop
      // with only the tcgen05 instructions under elect.sync keeps slot / phase in vector registers all the same: no fewer
      // instructions, 3 730 against 3 640 us per 1 225-tile launch.  profiles/r02_gpu_call_ai_aj_tail_chunks.log)
      if (rank == 0 && ptx::elect_one()) {
        const uint32_t idesc = ptx::make_idesc(2 * GEMM_BM, MLP_HC, false, false, F16 ? 0u : 1u);    // fc1: N = one hidden chunk
        const uint32_t idesc2 = ptx::make_idesc(2 * GEMM_BM, BN2, false, false, F16 ? 0u : 1u);      // proj, fc2: N = one output part
        const uint64_t a_desc0 = ptx::make_smem_desc_sw128(smem_a, 1024, 0);
        const uint64_t h_desc0 = ptx::make_smem_desc_sw128(smem_h, 1024, 0);
        const uint64_t w_desc0 = ptx::make_smem_desc_sw128(smem_w, 1024, 0);
        int slot = 0;
        uint32_t phase = 0;
        int g1 = 0;           // fc1-shaped chunks (hidden chunks AND QKV chunks) issued so far, all items: phase counter of h_full / h_tmem_empty,
                              // chunk g1 belongs to epilogue group g1 & 1
        int f2[2] = {0, 0};   // fc2 chunks consumed from each group's gelu buffer so far: phases of h_smem_full / h_smem_empty
        int t = 0;
        bool tl = false;
        auto issue_proj = [&]() {
#pragma unroll 1
          for (int kb = 0; kb < KB1; ++kb) {
            const uint64_t adesc = ptx::desc_advance(a_desc0, kb * MLP_KB_BYTES);
#pragma unroll 1
            for (int np = 0; np < NP2; ++np) {
              ptx::mbar_wait(w_full + 8 * slot, phase, 30);
              ptx::tc_fence_after();
              const uint64_t bdesc = ptx::desc_advance(w_desc0, slot * MLP_SLOT_BYTES);
              if (!(args.debug & 1)) {
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k)
                  ptx::umma_bf16_ss_2cta(tmem_base + np * BN2, ptx::desc_advance(adesc, k * 32), ptx::desc_advance(bdesc, k * 32), idesc2,
                                         (kb > 0 || k > 0) ? 1u : 0u);
              }
              ptx::umma_commit_2cta_mask(w_empty + 8 * slot, pair_mask);
              if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
          }
          ptx::umma_commit_2cta_mask(proj_full, pair_mask);
        };
        // one fc1-shaped chunk (128 output columns, K = D): hidden chunk c of W1 into the hidden-chunk accumulator (acc_col = MLP_H_COL,
        // behind = 1: the chunk before it used the same columns), or a QKV chunk -- those alternate between the hidden-chunk accumulator
        // and the first 128 OUT columns (free between ep 2 and the next item's proj), so the tensor core never waits for the epilogue
        // of the chunk it has just finished (behind = 2: the columns' previous user is the chunk before last; 0: nobody to wait for)
        auto issue_chunk = [&](int c, bool last_use_of_a, uint32_t acc_col = MLP_H_COL, int behind = 1) {
          const uint32_t h_tmem = tmem_base + acc_col;
          const bool fine = tl && (args.debug & 16) && (c == 6 || c == 7);   // diagnostics: fc1(6) -> slots 31 / 32, fc1(7) -> 35 / 52
          if (behind > 0 && g1 >= behind) {   // the accumulator's previous user has been read out of TMEM (by the group of its parity)
            const int u = g1 - behind;
            ptx::mbar_wait(h_tmem_empty + 8 * (u & 1), (u >> 1) & 1, 33);
            ptx::tc_fence_after();
          }
          mlp_stamp(fine, smem_tl, c == 6 ? 31 : 35);   // accumulator free
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
            ptx::umma_commit_2cta_mask(w_empty + 8 * slot, pair_mask);
            if (++slot == SLOTS) { slot = 0; phase ^= 1; }
          }
          ptx::umma_commit_2cta_mask(h_full + 8 * (g1 & 1), pair_mask);
          mlp_stamp(fine, smem_tl, c == 6 ? 32 : 52);   // issued
          if (c < 5) mlp_stamp(tl, smem_tl, 36 + 2 * c);
          if (c >= 100 && c < 108) mlp_stamp(tl && !(args.debug & 16), smem_tl, 46 + (c - 100));   // QKV chunk c - 100 issued
          if (last_use_of_a) ptx::umma_commit_2cta_mask(a_empty, pair_mask);   // smem_a may take the next CTX tile once this chunk retires
          ++g1;
        };
        int gbase = 0;   // g1 of the item's first hidden chunk
        auto issue_fc1 = [&](int c) { issue_chunk(c, NQ == 0 && c + 1 == NC); };
        auto issue_fc2 = [&](int c) {
          const int g = (gbase + c) & 1;   // the epilogue group (and gelu buffer) of hidden chunk c
          ptx::mbar_wait_cluster(h_smem_full + 8 * g, f2[g] & 1, 35);   // gelu(chunk) sits in both CTAs' shared memory
          ptx::tc_fence_after();
          if (c < 5) mlp_stamp(tl, smem_tl, 36 + 2 * c + 1);
          if ((args.debug & 16) && c == 5) mlp_stamp(tl, smem_tl, 34);   // diagnostics: gelu(5) seen
          const uint64_t hdesc = ptx::desc_advance(h_desc0, g * Cfg::H_BYTES);
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
                for (int k = 0; k < GEMM_BK / 16; ++k)   // always accumulating: OUT holds the residual rows (ep 1)
                  ptx::umma_bf16_ss_2cta(tmem_base + np * BN2, ptx::desc_advance(adesc, k * 32), ptx::desc_advance(bdesc, k * 32), idesc2, 1u);
              }
              ptx::umma_commit_2cta_mask(w_empty + 8 * slot, pair_mask);
              if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
          }
          ptx::umma_commit_2cta_mask(h_smem_empty + 8 * g, pair_mask);
          if (args.debug & 16) { if (c == 4) mlp_stamp(tl, smem_tl, 33); if (c == 5) mlp_stamp(tl, smem_tl, 53); }   // diagnostics: fc2(4) / fc2(5) issued
          ++f2[g];
        };
        for (int tile = tile0; tile < tiles_m; tile += tstep, ++t) {
          tl = args.timeline != nullptr && blockIdx.x == 0 && t == args.timeline_item;
          ptx::mbar_wait(a_full, t & 1, 38);
          ptx::tc_fence_after();
          mlp_stamp(tl, smem_tl, 60);
          if (t > 0) {   // OUT holds the previous item until every epilogue warp of both CTAs has read it for the last time
            ptx::mbar_wait(out_empty, (t - 1) & 1, 36);
            ptx::tc_fence_after();
          }
          issue_proj();
          mlp_stamp(tl, smem_tl, 56);
          ptx::mbar_wait_cluster(mid_ready, t & 1, 39);   // residual rows in TMEM, norm2 rows in both CTAs' smem_a
          ptx::tc_fence_after();
          mlp_stamp(tl, smem_tl, 59);
          gbase = g1;
          issue_fc1(0);
          if (NC > 1) issue_fc1(1);
          issue_fc2(0);
          if (NC > 2) issue_fc1(2);
          for (int c = 1; c < NC; ++c) {
            if (c + 2 < NC) issue_fc1(c + 2);
            issue_fc2(c);
          }
          ptx::umma_commit_2cta_mask(out_full, pair_mask);
          if (NQ > 0) {   // the next block's QKV projection on the rows ep 2 leaves in smem_a
            ptx::mbar_wait_cluster(xn_ready, t & 1, 28);
            ptx::tc_fence_after();
            mlp_stamp(tl, smem_tl, 63);   // the next norm1's rows seen in both CTAs' smem_a
            for (int c = 0; c < NQ; ++c) issue_chunk(c + 100, c + 1 == NQ, (c & 1) ? 0u : static_cast<uint32_t>(MLP_H_COL), c == 0 ? 1 : (c == 1 ? 0 : 2));
            // both accumulators have been read out before the next item's proj (OUT columns) and first hidden chunk are issued
            for (int u = g1 - 1; u >= 0 && u >= g1 - 2; --u) ptx::mbar_wait(h_tmem_empty + 8 * (u & 1), (u >> 1) & 1, 27);
            ptx::tc_fence_after();
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_EPI));
    const int ew = warp - 4;
    const int q = warp & 3;          // TMEM lane quadrant
    const int ch = (ew >> 2) & 1;    // column half of a hidden chunk (GELU)
    const int grp = ew >> 3;         // epilogue group: takes the chunks whose global index has this parity
    const int cg = ew >> 2;          // column quarter of a row at the item boundary
    constexpr int CPW = MLP_HC / 2;  // hidden columns per warp and chunk
    constexpr int NSUB = CPW / 32;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int row = q * 32 + lane;   // row of the CTA's 128
    const uint32_t my_h = smem_h + grp * Cfg::H_BYTES;   // this group's gelu buffer
    const uint32_t box_h = smem_h + ew * 4096;           // this warp's 4 KB box inside the gelu buffers (free at the item boundary)
    const uint32_t xn_box = smem_a + ew * Cfg::STAGE_A;  // ... its share of smem_a (free between proj and the norm2 rows, and during ep 2):
    const uint32_t box_a = xn_box + 2048;                //     a 2 KB box for 16-bit rows and (NS > 1) a second 4 KB box
    const uint32_t xb_h = xbox_bar + ew * 16, xb_a = xb_h + 8;
    const uint32_t stat_addr = box_h + lane * 8;
    const int col0 = cg * CW;
    const int sw = lane & 7;
    const uint32_t my_h_full = h_full + 8 * grp, my_tmem_empty = h_tmem_empty + 8 * grp;
    const uint32_t my_smem_full = ptx::mapa(h_smem_full + 8 * grp, 0);
    const uint32_t mid_ready_leader = ptx::mapa(mid_ready, 0);
    const bool strong_arrive = (args.debug & 32) != 0;   // diagnostics / A-B: hand-overs with .release.cluster arrives (ptx::mbar_arrive_remote)
    int cnt_h = 0, cnt_a = 0;   // boxes consumed so far: phases of xb_h / xb_a
    int n0 = 0;                 // global index of the item's first fc1-shaped chunk (hidden + QKV chunks of all items of this CTA)
    int k2 = 0;                 // hidden chunks this group has handed to fc2 so far: phase of its h_smem_empty
    int t = 0;
    auto load_xbox = [&](uint32_t dst, uint32_t bar, int col, int row_g) {   // lane 0
      ptx::mbar_arrive_expect_tx(bar, 4096);
      ptx::tma_load_2d(dst, &tmap_x, bar, col, row_g);
    };
    // all four column quarters of this thread's row: Chan's combination of four groups of CW values
    auto combine4 = [&](float& mean, float& rstd) {
      float mp[4];
      float m2 = 0.f;
      mean = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 v = ptx::lds_v2f(smem_h + (c * 4 + q) * 4096 + lane * 8);
        mp[c] = v.x;
        mean += v.x;
        m2 += v.y;
      }
      mean *= 0.25f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float d = mp[c] - mean;
        m2 = fmaf(static_cast<float>(CW) * d, d, m2);
      }
      rstd = rsqrtf(m2 * (1.0f / static_cast<float>(D)) + args.ln_eps);
    };
    if (tile0 < tiles_m && lane == 0 && !(args.debug & 2)) load_xbox(box_h, xb_h, col0, tile0 * 2 * GEMM_BM + rank * GEMM_BM + q * 32);
    for (int tile = tile0; tile < tiles_m; tile += tstep, ++t, n0 += NC + NQ) {
      const bool tl = args.timeline != nullptr && blockIdx.x == 0 && t == args.timeline_item && ew == 0 && lane == 0;
      const int row_g = tile * 2 * GEMM_BM + rank * GEMM_BM + q * 32;   // first row of this warp's boxes
      // ------------------------------------------------------------------ ep 1: residual add + norm2
      ptx::mbar_wait(proj_full, t & 1, 43);
      ptx::tc_fence_after();
      mlp_stamp(tl, smem_tl, 57);
      if (NS > 1 && lane == 0 && !(args.debug & 2)) load_xbox(box_a, xb_a, col0 + 32, row_g);   // proj has retired: smem_a is free until the norm2 rows
      {
        float mean_w = 0.f, m2_w = 0.f;
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
          const int col = col0 + s * 32;
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_taddr + col, r);
          const uint32_t xbox = (s & 1) ? box_a : box_h;
          if (!(args.debug & 2)) {
            if (s & 1) { ptx::mbar_wait(xb_a, cnt_a & 1, 44); ++cnt_a; }
            else { ptx::mbar_wait(xb_h, cnt_h & 1, 45); ++cnt_h; }
          }
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 xo = ptx::lds_v4f(xbox + lane * 128 + ((j ^ sw) << 4));   // SWIZZLE_128B box row
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias_p + col) + j);
            y[4 * j] = xo.x + b4.x; y[4 * j + 1] = xo.y + b4.y; y[4 * j + 2] = xo.z + b4.z; y[4 * j + 3] = xo.w + b4.w;
          }
          ptx::tmem_ld_wait(r);
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(r[j]);
          float mean_s, m2_s;
          tail_stats32(y, mean_s, m2_s);
          tail_chan_step(s, mean_w, m2_w, mean_s, m2_s);
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(y[j]);
          ptx::tmem_st_32x32b_x32(lane_taddr + col, r);   // the residual rows: every fc2 MMA accumulates on them
          if (s + 2 < NS) {   // this box takes the rows of step s + 2
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(args.debug & 2)) load_xbox(xbox, (s & 1) ? xb_a : xb_h, col + 64, row_g);
          }
          mlp_stamp(tl, smem_tl, 15 + s);
        }
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stat_addr), "f"(mean_w), "f"(m2_w) : "memory");
      }
      ptx::tmem_st_wait();
      mlp_stamp(tl, smem_tl, 58);
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");   // partial statistics published; every staging box in smem_a has been read
      {
        float mean, rstd;
        combine4(mean, rstd);
        const float nmr = -mean * rstd;   // (x - mean) rstd = fma(x, rstd, nmr)
        mlp_stamp(tl, smem_tl, 18);
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
          const int col = col0 + s * 32;
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_taddr + col, r);
          ptx::tmem_ld_wait(r);
          uint32_t pk[16];
          if (args.fold2) {   // gamma / beta live in W1 / bias1
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = ptx::pack_h2<F16>(fmaf(__uint_as_float(r[2 * j]), rstd, nmr), fmaf(__uint_as_float(r[2 * j + 1]), rstd, nmr));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(args.ln2_w + col) + j);
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.ln2_b + col) + j);
              const float v0 = fmaf(fmaf(__uint_as_float(r[4 * j]), rstd, nmr), g4.x, b4.x), v1 = fmaf(fmaf(__uint_as_float(r[4 * j + 1]), rstd, nmr), g4.y, b4.y);
              const float v2 = fmaf(fmaf(__uint_as_float(r[4 * j + 2]), rstd, nmr), g4.z, b4.z), v3 = fmaf(fmaf(__uint_as_float(r[4 * j + 3]), rstd, nmr), g4.w, b4.w);
              pk[2 * j] = ptx::pack_h2<F16>(v0, v1); pk[2 * j + 1] = ptx::pack_h2<F16>(v2, v3);
            }
          }
          // 32 columns = half of a 128-byte row of k-block col / 64 of the A tile (SWIZZLE_128B)
          const uint32_t rowaddr = smem_a + (col >> 6) * MLP_KB_BYTES + row * 128;
          const int c16 = (col & 63) >> 3;   // first 16-byte chunk
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ptx::sts_v4(rowaddr + (((c16 + j) ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (strong_arrive) ptx::mbar_arrive_remote(mid_ready_leader); else ptx::mbar_arrive_remote_cta(mid_ready_leader); }
      mlp_stamp(tl, smem_tl, 55);
      // ------------------------------------------------------------------ hidden chunks (mlp_fused_sm100.cuh)
      for (int c = (n0 + grp) & 1; c < NC; c += 2) {
        const int n = n0 + c;          // (n & 1) == grp
        // diagnostics (debug bit 4): one steady-state chunk of each group in detail -- chunk 4 by warp 0 (slots 25..30), chunk 5 by
        // warp 8 (slots 46..51): accumulator complete, in registers, GELU done, gelu buffer free, stored + fenced, handed over
        const bool fine = (args.debug & 16) && args.timeline != nullptr && blockIdx.x == 0 && t == args.timeline_item && lane == 0 &&
                          ((ew == 0 && c == 4) || (ew == 8 && c == 5));
        const int fs = ew == 0 ? 25 : 46;
        ptx::mbar_wait(my_h_full, (n >> 1) & 1, 40);
        ptx::tc_fence_after();
        mlp_stamp(fine, smem_tl, fs);
        if (c < 5) mlp_stamp(tl, smem_tl, 3 * c);
        uint32_t r[NSUB][32];
        static_assert(NSUB == 2, "block tail: a warp reads its 64 columns of a hidden chunk in one load");
        ptx::tmem_ld_32x32b_x64_wait(lane_taddr + MLP_H_COL + ch * CPW, r[0], r[1]);
        mlp_stamp(fine, smem_tl, fs + 1);
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
          if (args.gelu5) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_sigmoid5_x2(v[j], v[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_sigmoid_x2(v[j], v[j + 1]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[s][j] = ptx::pack_h2<F16>(v[2 * j], v[2 * j + 1]);
        }
        if (c < 5) mlp_stamp(tl, smem_tl, 3 * c + 1);
        mlp_stamp(fine, smem_tl, fs + 2);
        // this group's H buffer is free once fc2 of its previous chunk has retired
        if (k2 > 0) ptx::mbar_wait(h_smem_empty + 8 * grp, (k2 - 1) & 1, 41);
        ++k2;
        mlp_stamp(fine, smem_tl, fs + 3);
        if (NQ > 0) {   // ... and once the TMA store of this warp's last QKV chunk has read its region
          if (lane == 0) ptx::bulk_wait_read0();
          __syncwarp();
        }
        const uint32_t tile_addr = my_h + ch * MLP_KB_BYTES + row * 128;
#pragma unroll
        for (int s = 0; s < NSUB; ++s)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ptx::sts_v4(tile_addr + (((s * 4 + j) ^ (row & 7)) << 4), pk[s][4 * j], pk[s][4 * j + 1], pk[s][4 * j + 2], pk[s][4 * j + 3]);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        mlp_stamp(fine, smem_tl, fs + 4);
        if (lane == 0) { if (strong_arrive) ptx::mbar_arrive_remote(my_smem_full); else ptx::mbar_arrive_remote_cta(my_smem_full); }
        mlp_stamp(fine, smem_tl, fs + 5);
        if (c < 5) mlp_stamp(tl, smem_tl, 3 * c + 2);
      }
      // ------------------------------------------------------------------ ep 2: X out (+ the next LayerNorm)
      ptx::mbar_wait(out_full, t & 1, 42);   // the item's last fc2 has retired: OUT is complete, both gelu buffers are free
      ptx::tc_fence_after();
      mlp_stamp(tl, smem_tl, 61);
      const bool want_xn = args.lnn_w != nullptr;
      float mean = 0.f, rstd = 0.f;
      // (Measured and dropped: the statistics pass writing OUT + b2 back into the OUT columns so that the second pass needs no bias
      // loads / adds -- 1.25 instructions per element fewer on paper, but ptxas answers the extra tcgen05.st and the branch around the
      // bias loads with 480 instead of 196 bytes of spills: 3 975 against 3 475 us per 1 225-tile launch.  profiles/r02_gpu_call_as_ep2_writeback.log)
      if (want_xn) {
        float mean_w = 0.f, m2_w = 0.f;
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
          const int col = col0 + s * 32;
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_taddr + col, r);
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias2 + col) + j);
            y[4 * j] = b4.x; y[4 * j + 1] = b4.y; y[4 * j + 2] = b4.z; y[4 * j + 3] = b4.w;
          }
          ptx::tmem_ld_wait(r);
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(r[j]);
          float mean_s, m2_s;
          tail_stats32(y, mean_s, m2_s);
          tail_chan_step(s, mean_w, m2_w, mean_s, m2_s);
        }
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stat_addr), "f"(mean_w), "f"(m2_w) : "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        combine4(mean, rstd);
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");   // every partial has been read: the boxes may be overwritten
      }
      mlp_stamp(tl, smem_tl, 54);
      const float nmr2 = -mean * rstd;
      const int sw2 = (lane >> 1) & 3;
      // staging: fp32 boxes alternate between box_h and box_a, 16-bit boxes go through xn_box; one bulk group per store, in the order
      // X0, XN0, X1, XN1, X2, XN2 (X0, X1, X2 without the trailing LayerNorm)
#pragma unroll 1
      for (int s = 0; s < NS; ++s) {
        const int col = col0 + s * 32;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(lane_taddr + col, r);
        float y[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias2 + col) + j);
          y[4 * j] = b4.x; y[4 * j + 1] = b4.y; y[4 * j + 2] = b4.z; y[4 * j + 3] = b4.w;
        }
        ptx::tmem_ld_wait(r);
        if (s == NS - 1) {   // this warp has read OUT for the last time
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_leader(out_empty);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(r[j]);
        const bool to_smem = NQ > 0;   // the normalised rows feed the QKV chunks from smem_a: no staging space there, fp32 boxes through box_h only
        const uint32_t xbox = (!to_smem && (s & 1)) ? box_a : box_h;
        if (to_smem ? s >= 1 : s >= 2) {   // the store that used this box last has read it
          if (lane == 0) { if (to_smem) ptx::bulk_wait_read0(); else if (want_xn) ptx::bulk_wait_read2(); else ptx::bulk_wait_read1(); }
          __syncwarp();
        }
        const uint32_t rowaddr = xbox + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ptx::sts_v4(rowaddr + ((j ^ sw) << 4), __float_as_uint(y[4 * j]), __float_as_uint(y[4 * j + 1]), __float_as_uint(y[4 * j + 2]),
                      __float_as_uint(y[4 * j + 3]));
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(args.debug & 4)) {
          ptx::tma_store_2d(&tmap_x, xbox, col, row_g);
          ptx::bulk_commit();
        }
        mlp_stamp(tl, smem_tl, 20 + s);
        if (want_xn) {
          uint32_t pk[16];
          if (args.foldn) {   // gamma / beta live in Wqkv / bias_qkv
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_h2<F16>(fmaf(y[2 * j], rstd, nmr2), fmaf(y[2 * j + 1], rstd, nmr2));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(args.lnn_w + col) + j);
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.lnn_b + col) + j);
              const float v0 = fmaf(fmaf(y[4 * j], rstd, nmr2), g4.x, b4.x), v1 = fmaf(fmaf(y[4 * j + 1], rstd, nmr2), g4.y, b4.y);
              const float v2 = fmaf(fmaf(y[4 * j + 2], rstd, nmr2), g4.z, b4.z), v3 = fmaf(fmaf(y[4 * j + 3], rstd, nmr2), g4.w, b4.w);
              pk[2 * j] = ptx::pack_h2<F16>(v0, v1); pk[2 * j + 1] = ptx::pack_h2<F16>(v2, v3);
            }
          }
          if (to_smem) {   // half of a 128-byte row of k-block col / 64 of the A tile (SWIZZLE_128B), as the norm2 rows in ep 1
            const uint32_t arow = smem_a + (col >> 6) * MLP_KB_BYTES + row * 128;
            const int c16 = (col & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ptx::sts_v4(arow + (((c16 + j) ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          } else {
            if (s >= 1) {   // the previous 16-bit store has read xn_box (only this step's fp32 store may still be pending)
              if (lane == 0) ptx::bulk_wait_read1();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)   // 32 x 32 16-bit box, SWIZZLE_64B
              ptx::sts_v4(xn_box + lane * 64 + ((j ^ sw2) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(args.debug & 4)) {
              ptx::tma_store_2d(&tmap_xn, xn_box, col, row_g);
              ptx::bulk_commit();
            }
          }
        }
      }
      if (lane == 0 && tile + tstep < tiles_m) {   // the next item's fp32 rows on their way to L2 (no earlier: see the CTX-tile producer)
        const int next_row = (tile + tstep) * 2 * GEMM_BM + rank * GEMM_BM + q * 32;
#pragma unroll
        for (int s = 1; s < NS; ++s) ptx::tma_prefetch_2d(&tmap_x, col0 + s * 32, next_row);
      }
      if (NQ > 0) {
        // ------------------------------------------------------------------ the next block's QKV chunks
        ptx::fence_proxy_async_smem();   // the normalised rows in smem_a, for the tensor core
        __syncwarp();
        if (lane == 0) { if (strong_arrive) ptx::mbar_arrive_remote(ptx::mapa(xn_ready, 0)); else ptx::mbar_arrive_remote_cta(ptx::mapa(xn_ready, 0)); }
        mlp_stamp(tl, smem_tl, 19);   // this warp's share of the next norm1's rows handed over
        const int nq0 = n0 + NC;
        for (int c = (nq0 + grp) & 1; c < NQ; c += 2) {
          const int n = nq0 + c;          // (n & 1) == grp
          ptx::mbar_wait(my_h_full, (n >> 1) & 1, 46);
          ptx::tc_fence_after();
          const int ts = c >> 1;   // diagnostics: this group's first two QKV chunks
          if (ts < 2) mlp_stamp(tl && !(args.debug & 16), smem_tl, 25 + 4 * ts);
          uint32_t r[NSUB][32];
          ptx::tmem_ld_32x32b_x64_wait(lane_taddr + ((c & 1) ? 0 : MLP_H_COL) + ch * CPW, r[0], r[1]);   // (QKV chunks alternate between two accumulators)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_leader(my_tmem_empty);
          uint32_t pk[NSUB][16];
#pragma unroll
          for (int s = 0; s < NSUB; ++s) {
            const float* bp = args.bias_qkv + c * MLP_HC + ch * CPW + s * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
              pk[s][2 * j] = ptx::pack_h2<F16>(__uint_as_float(r[s][4 * j]) + b4.x, __uint_as_float(r[s][4 * j + 1]) + b4.y);
              pk[s][2 * j + 1] = ptx::pack_h2<F16>(__uint_as_float(r[s][4 * j + 2]) + b4.z, __uint_as_float(r[s][4 * j + 3]) + b4.w);
            }
          }
          // staging = this warp's own 4 KB region of its group's gelu buffer (32 rows x 64 columns, SWIZZLE_128B = box_h): free once
          // the warp's previous store has read it (the fp32 boxes of ep 2, the QKV chunk before this one)
          if (ts < 2) mlp_stamp(tl && !(args.debug & 16), smem_tl, 26 + 4 * ts);
          if (lane == 0) ptx::bulk_wait_read0();
          __syncwarp();
          if (ts < 2) mlp_stamp(tl && !(args.debug & 16), smem_tl, 27 + 4 * ts);
#pragma unroll
          for (int s = 0; s < NSUB; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ptx::sts_v4(box_h + lane * 128 + (((s * 4 + j) ^ sw) << 4), pk[s][4 * j], pk[s][4 * j + 1], pk[s][4 * j + 2], pk[s][4 * j + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(args.debug & 4)) {
            ptx::tma_store_2d(&tmap_qkv, box_h, c * MLP_HC + ch * CPW, row_g);
            ptx::bulk_commit();
          }
          if (ts < 2) mlp_stamp(tl && !(args.debug & 16), smem_tl, 28 + 4 * ts);
        }
        if (lane == 0) {   // box_h goes on to receive the next item's first fp32 rows
          ptx::bulk_wait_read0();
          mlp_stamp(tl, smem_tl, 24);
          if (tile + tstep < tiles_m && !(args.debug & 2)) load_xbox(box_h, xb_h, col0, (tile + tstep) * 2 * GEMM_BM + rank * GEMM_BM + q * 32);
        }
        __syncwarp();
      } else {
        // every staging box has been read: smem_a goes to the CTX-tile producer, box_h receives the next item's first fp32 rows
        if (lane == 0) {
          ptx::bulk_wait_read0();
          mlp_stamp(tl, smem_tl, 24);
          ptx::mbar_arrive(stage_free);
          if (tile + tstep < tiles_m && !(args.debug & 2)) load_xbox(box_h, xb_h, col0, (tile + tstep) * 2 * GEMM_BM + rank * GEMM_BM + q * 32);
        }
        __syncwarp();
      }
      mlp_stamp(tl, smem_tl, 62);
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
