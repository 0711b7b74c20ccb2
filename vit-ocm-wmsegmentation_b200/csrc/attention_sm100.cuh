// Flash-style multi-head self-attention for sm_100a (head_dim 64), tcgen05 + TMEM + TMA:
//     ctx[b, i, h*64:(h+1)*64] = softmax_j( scale * q_i . k_j ) @ v        (never materialises N x N)
// Replaces SSS/dino/vision_transformer.py:83-87 (q@k^T*scale, softmax, attn@v, transpose/reshape)
// for the blocks whose attention matrix is not returned.
//
// One work item = one (image b, head h, 128-query tile); CTAs are persistent and walk the items.  q/k/v are read straight out of the fused
// QKV activation [B*N, ld] (bf16, columns [3][H][64]) by one 2-D TMA tensor map (box 64 x 128,
// SWIZZLE_128B).  Warp roles (256 threads): warps 0..3 = softmax (one thread per query row; TMEM lane
// quadrant = warp % 4), warp 4 = TMA producer, warp 5 = MMA issuer + TMEM allocator, warps 6..7 idle.
// setmaxnreg moves registers from warpgroup 1 to the softmax warpgroup (208 vs 48 per thread in bf16 mode).
//   S = Q K_j^T      : tcgen05.mma  M128 x N(<=128) x K64, both operands K-major, into TMEM
//   softmax          : one tcgen05.ld pass of S into registers; scale/shift on the packed f32x2 pipe
//                      (FFMA2), MUFU.EX2, row sums on FADD2, P -> TMEM (tcgen05.st, packed bf16x2 columns):
//                      no shared-memory round trip, no generic->async proxy fence
//   O += P V_j       : tcgen05.mma  M128 x N64 x K(<=128), A = P from TMEM, B = V (smem, MN-major),
//                      accumulating in TMEM
// The running maximum is only raised when a row exceeds it by more than 2^8 ("lazy rescale"): softmax
// is shift invariant, so a stale maximum changes nothing but keeps O in TMEM untouched in the
// common case; when it is raised the softmax warps rescale their O rows in TMEM.
// Two CTAs are co-resident per SM (64 KB smem, 256 TMEM columns each: S 128 | O 64 | P 64) so one CTA's MMAs overlap
// the other's exponentials.  SPLIT = true is the fp32-parity mode: every operand is a bf16
// (hi, lo) pair and each product is hi*hi + hi*lo + lo*hi (fp32 accumulate in TMEM).
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace vitocm {

struct AttnArgs {
  int n_items;       // work items = query tiles x heads x images (persistent CTAs walk them with stride gridDim.x)
  int n_qtiles, heads;
  int n_tokens;      // N per image (785 for 224^2 / patch 8)
  int embed_dim;     // D = H * 64
  int lo_col_off;    // SPLIT: column offset of the lo halves inside the qkv activation (= 3D)
  float scale_log2;  // qk scale * log2(e)
  __nv_bfloat16* out;  // ctx [B*N, ldo]
  long long ldo;
  int out_lo_off;    // SPLIT: column offset of the lo half of ctx
  float* lse2;       // training: [B][H][Npad] (Npad = N rounded up to 128) log2-sum-exp of the scaled logits
                     // (max * scale_log2 + log2 l); +inf in the pad rows; or nullptr
  int timeline_item;    // diagnostics: which of a CTA's work items (0, 1, ...) the stamps are taken on
  long long* timeline;  // diagnostics (vitocm_attention_timeline) or nullptr: clock64 stamps of CTAs (0,0,0) and (1,0,0)
};

// timeline layout: [cta 0..1][role 0 = softmax warp 0, 1 = MMA thread][kv block j < 16][event < 8]
constexpr int ATT_TL_EVENTS = 8;
constexpr int ATT_TL_BLOCKS = 16;
__device__ __forceinline__ void att_stamp(const AttnArgs& args, bool on, int role, int j, int ev, int pipe = -1) {
  if (pipe < 0) pipe = blockIdx.x;
  if (on && j < ATT_TL_BLOCKS) args.timeline[((pipe * 2 + role) * ATT_TL_BLOCKS + j) * ATT_TL_EVENTS + ev] = clock64();
}

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_DH = 64;
constexpr int ATT_THREADS = 256;   // warpgroup 0 = softmax (warps 0..3), warpgroup 1 = TMA (warp 4) + MMA (warp 5)
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr int ATT_RING = 3;
constexpr int ATT_S_COL = 0;      // S: 128 columns (fp32)
constexpr int ATT_O_COL = 128;    // O: 64 columns (fp32)
constexpr int ATT_P_COL = 192;    // P: 64 columns of packed bf16x2 (128 keys); split mode: lo part in the next 64
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units: raise the running maximum (rescale O, l) before the next block
constexpr float ATT_REDO_THRESHOLD = 60.0f;    // log2 units: exp2 of the current block may overflow -> redo it now
constexpr int ATT_POLY_DEFAULT = 0;            // see run_attention (0: all MUFU, 1: 4/16 polynomial, 2: 7/16)

template <bool SPLIT>
struct AttnCfg {
  static constexpr int NPART = SPLIT ? 2 : 1;                     // hi (+ lo)
  static constexpr int SLOT_BYTES = ATT_TILE_BYTES * NPART;       // one K or V block
  static constexpr int Q_BYTES = ATT_TILE_BYTES * NPART;
  static constexpr int SMEM_BYTES = Q_BYTES + ATT_RING * SLOT_BYTES + 1024 + 128;
  static constexpr int TMEM_COLS = SPLIT ? 512 : 256;
  // setmaxnreg budgets: 2 CTAs/SM x 128 x (208 + 48) = 64 K registers (bf16); one CTA/SM in split mode
  static constexpr int REGS_SOFTMAX = SPLIT ? 240 : 208;
  static constexpr int REGS_OTHER = SPLIT ? 64 : 48;
};

// POLY_MASK: bit i set = pair i of every 16 pairs of a 32-key chunk takes the FMA-pipe exp2 polynomial
template <bool SPLIT, uint32_t POLY_MASK>
__global__ void __launch_bounds__(ATT_THREADS, SPLIT ? 1 : 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs args) {
  using Cfg = AttnCfg<SPLIT>;
  constexpr int NPART = Cfg::NPART;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem;
  const uint32_t smem_ring = smem_q + Cfg::Q_BYTES;
  const uint32_t bars = smem_ring + ATT_RING * Cfg::SLOT_BYTES;
  const uint32_t q_full = bars;             // [1]
  const uint32_t kv_full = bars + 8;        // [3]
  const uint32_t kv_empty = bars + 32;      // [3]
  const uint32_t s_full = bars + 56;        // MMA -> softmax
  const uint32_t s_empty = bars + 64;       // softmax -> MMA   (4 warps)
  const uint32_t p_full = bars + 72;        // softmax -> MMA   (4 warps)
  const uint32_t o_full = bars + 80;        // MMA -> softmax
  const uint32_t q_empty = bars + 88;       // MMA -> producer: all S MMAs of the work item retired (Q tile reusable)
  const uint32_t o_empty = bars + 96;       // softmax -> MMA: O of the finished work item has been read (4 warps)
  const uint32_t tmem_ptr_smem = bars + 104;

  // Persistent CTA: work items (query tile, head, image), query tile fastest so that the CTAs running side by side share
  // one image-head's K / V through L2.  Barriers, the K/V ring and TMEM live across items (all phase counters run on), so
  // the next item's Q / K loads and its first S MMA overlap the current item's last exponentials and its epilogue.
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = args.n_tokens;
  const int D = args.embed_dim;
  const int n_kv = (N + ATT_BKV - 1) / ATT_BKV;
  const bool tl0 = args.timeline != nullptr && blockIdx.x < 2 && (threadIdx.x & 31) == 0;

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < ATT_RING; ++i) {
      ptx::mbar_init(kv_full + 8 * i, 1);
      ptx::mbar_init(kv_empty + 8 * i, 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 4);
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(q_empty, 1);
    ptx::mbar_init(o_empty, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ptx::lds_u32(tmem_ptr_smem);
  if (tl0 && threadIdx.x == 0) args.timeline[((blockIdx.x * 2 + 0) * ATT_TL_BLOCKS + 15) * ATT_TL_EVENTS + 6] = clock64();   // CTA start

  if (warp >= 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_OTHER));
   if (warp == 4) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int item = 0;   // K/V ring position, running across work items
      int w = 0;
      for (int it = blockIdx.x; it < args.n_items; it += gridDim.x, ++w) {
      const int qt = it % args.n_qtiles, h = (it / args.n_qtiles) % args.heads, b = it / (args.n_qtiles * args.heads);
      const int row_base = b * N;  // first row of this image in the [B*N, ld] activation
      ptx::mbar_wait(q_empty, (w & 1) ^ 1, 16);   // the previous item's S MMAs have retired
      ptx::mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
      for (int part = 0; part < NPART; ++part)
        ptx::tma_load_2d(smem_q + part * ATT_TILE_BYTES, &tmap_qkv, q_full, part * args.lo_col_off + h * ATT_DH,
                         row_base + qt * ATT_BQ);
      // ring order = consumption order of the MMA warp: K0, K1, V0, K2, V1, ..., V_{n-1}
      auto load = [&](int which /*1 = K, 2 = V*/, int j) {
        const int slot = item % ATT_RING;
        const uint32_t parity = ((item / ATT_RING) & 1) ^ 1;
        ptx::mbar_wait(kv_empty + 8 * slot, parity, 10);
        ptx::mbar_arrive_expect_tx(kv_full + 8 * slot, Cfg::SLOT_BYTES);
        for (int part = 0; part < NPART; ++part)
          ptx::tma_load_2d(smem_ring + slot * Cfg::SLOT_BYTES + part * ATT_TILE_BYTES, &tmap_qkv, kv_full + 8 * slot,
                           part * args.lo_col_off + which * D + h * ATT_DH, row_base + j * ATT_BKV);
        ++item;
      };
      load(1, 0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) load(1, j + 1);
        load(2, j);
      }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    // One elected thread issues (elect.sync lets ptxas emit the tcgen05 instructions without a per-instruction
    // leader-election loop); descriptors are advanced by compile-time constants in fully unrolled loops.
    if (ptx::elect_one()) {
      const uint32_t s_tmem = tmem_base + ATT_S_COL;
      const uint32_t o_tmem = tmem_base + ATT_O_COL;
      const uint64_t q_desc = ptx::make_smem_desc_sw128(smem_q, 1024, 0);
      int item = 0;   // K/V ring position, running across work items
      int g = 0;      // KV blocks processed so far (all work items): phase counter of s_full / s_empty / p_full / o_full
      int w = 0;
      for (int it = blockIdx.x; it < args.n_items; it += gridDim.x, ++w) {
      const bool tl = tl0 && w == args.timeline_item;
      auto kv_len_mma = [&](int j) {  // keys of block j rounded up to the MMA granularity (16)
        int len = N - j * ATT_BKV;
        len = len > ATT_BKV ? ATT_BKV : len;
        return (len + 15) & ~15;
      };
      auto issue_s = [&](int j) {
        const int slot = item % ATT_RING;
        ptx::mbar_wait(kv_full + 8 * slot, (item / ATT_RING) & 1, 11);
        ptx::tc_fence_after();
        const uint64_t k_desc = ptx::make_smem_desc_sw128(smem_ring + slot * Cfg::SLOT_BYTES, 1024, 0);
        const uint32_t idesc = ptx::make_idesc(ATT_BQ, kv_len_mma(j), false, false);
        // terms: (Qhi,Khi) [, (Qhi,Klo), (Qlo,Khi)]; K-major operands advance 32 B per 16-wide k step
#pragma unroll
        for (int t = 0; t < (SPLIT ? 3 : 1); ++t) {
          const uint64_t qa = ptx::desc_advance(q_desc, t == 2 ? ATT_TILE_BYTES : 0);
          const uint64_t ka = ptx::desc_advance(k_desc, t == 1 ? ATT_TILE_BYTES : 0);
#pragma unroll
          for (int k = 0; k < ATT_DH / 16; ++k)
            ptx::umma_bf16_ss(s_tmem, ptx::desc_advance(qa, k * 32), ptx::desc_advance(ka, k * 32), idesc, (t | k) ? 1u : 0u);
        }
        ptx::umma_commit(kv_empty + 8 * slot);
        ptx::umma_commit(s_full);
        att_stamp(args, tl, 1, j, 0);   // S_j issued
        ++item;
      };
      auto issue_pv = [&](int j) {
        const int slot = item % ATT_RING;
        ptx::mbar_wait(kv_full + 8 * slot, (item / ATT_RING) & 1, 12);
        ptx::tc_fence_after();
        // O accumulates across KV blocks in TMEM.
        // A = P in TMEM: 16 keys = 8 packed columns per step
        // B = V: MN-major [keys x 64]; 16 keys = two 8-row groups of 1024 B
        const uint64_t v_desc = ptx::make_smem_desc_sw128(smem_ring + slot * Cfg::SLOT_BYTES, 1024, 1024);
        constexpr uint32_t idesc = ptx::make_idesc(ATT_BQ, ATT_DH, false, /*B = V is MN-major*/ true);
        const int ksteps = kv_len_mma(j) / 16;
        const uint32_t acc0 = j > 0 ? 1u : 0u;
        if (j == 0 && w > 0) {   // O still holds the previous work item until the softmax warps have read it out
          ptx::mbar_wait(o_empty, (w - 1) & 1, 17);
          ptx::tc_fence_after();
        }
        // terms: (Phi,Vhi) [, (Phi,Vlo), (Plo,Vhi)]
#pragma unroll
        for (int t = 0; t < (SPLIT ? 3 : 1); ++t) {
          const uint32_t pa = tmem_base + ATT_P_COL + (t == 2 ? 64 : 0);
          const uint64_t va = ptx::desc_advance(v_desc, t == 1 ? ATT_TILE_BYTES : 0);
          if (ksteps == ATT_BKV / 16) {
#pragma unroll
            for (int k = 0; k < ATT_BKV / 16; ++k)
              ptx::umma_bf16_ts(o_tmem, pa + k * 8, ptx::desc_advance(va, k * 2048), idesc, (t | k) ? 1u : acc0);
          } else {
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k)
              ptx::umma_bf16_ts(o_tmem, pa + k * 8, ptx::desc_advance(va, k * 2048), idesc, (t | k) ? 1u : acc0);
          }
        }
        ptx::umma_commit(kv_empty + 8 * slot);
        ptx::umma_commit(o_full);
        att_stamp(args, tl, 1, j, 1);   // PV_j issued
        ++item;
      };
      ptx::mbar_wait(q_full, w & 1, 13);
      if (g > 0) ptx::mbar_wait(s_empty, (g - 1) & 1, 14);   // the previous item's last S has been read out of TMEM
      ptx::tc_fence_after();
      issue_s(0);
      if (n_kv == 1) ptx::umma_commit(q_empty);
      for (int j = 0; j < n_kv; ++j, ++g) {
        if (j + 1 < n_kv) {
          ptx::mbar_wait(s_empty, g & 1, 14);  // softmax has read S_j out of TMEM
          ptx::tc_fence_after();
          issue_s(j + 1);
          if (j + 2 == n_kv) ptx::umma_commit(q_empty);   // last S MMA of the item: Q tile reusable once it retires
        }
        ptx::mbar_wait(p_full, g & 1, 15);     // P_j in TMEM, O rescaled if needed
        ptx::tc_fence_after();
        issue_pv(j);
      }
      }
    }
   }
  } else {
    // ===================== softmax / output (warps 0..3) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_SOFTMAX));
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float sl2 = args.scale_log2;
    const uint64_t sl2_2 = ptx::dup_f32x2(sl2);
    int g = 0;                    // KV blocks processed so far (all work items): barrier phase counter
    int w = 0;
    for (int it = blockIdx.x; it < args.n_items; it += gridDim.x, ++w) {
    const int qt = it % args.n_qtiles, h = (it / args.n_qtiles) % args.heads, b = it / (args.n_qtiles * args.heads);
    const int row_base = b * N;
    const bool tl = tl0 && w == args.timeline_item;
    float m_used = -INFINITY;     // the row maximum the exponentials are taken against
    float m_next = -INFINITY;     // a larger maximum seen in the previous block (lazy rescale pending)
    float l_run = 0.f;            // running row sum (same units as O in TMEM)

    for (int j = 0; j < n_kv; ++j, ++g) {
      int kv_len = N - j * ATT_BKV;
      kv_len = kv_len > ATT_BKV ? ATT_BKV : kv_len;
      const int nchunks = (((kv_len + 15) & ~15) + 31) >> 5;   // 32-column chunks the MMA produced
      att_stamp(args, tl && warp == 0, 0, j, 0);   // waiting for S_j
      ptx::mbar_wait(s_full, g & 1, 21);
      ptx::tc_fence_after();
      att_stamp(args, tl && warp == 0, 0, j, 1);   // S_j complete
      // ---- S_j -> registers (one pass), then hand the TMEM columns back to the MMA warp
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nchunks) ptx::tmem_ld_32x32b_x32(lane_addr + ATT_S_COL + c * 32, s[c]);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nchunks) ptx::tmem_ld_wait(s[c]);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_empty);
      att_stamp(args, tl && warp == 0, 0, j, 2);   // S_j in registers
      // ---- ragged last block only: columns beyond the sequence (and chunks the MMA never wrote) -> -inf,
      //      so the common path below carries no masks (exp2(-inf) = 0)
      if (kv_len < ATT_BKV) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= kv_len) s[c][i] = 0xff800000u;
      }
      // ---- row maximum.  Block 0 needs it before any exponential (it sets the scale).  For the later blocks the
      // exponentials are taken against the maximum already in use and this block's maximum is computed inside the
      // same instruction stream (off the critical path): softmax is shift invariant and O / l live in fp32, so a
      // stale maximum costs nothing until a row exceeds it by 2^ATT_RESCALE_THRESHOLD -- then O and l are rescaled
      // before the NEXT block -- or by 2^ATT_REDO_THRESHOLD inside one block -- then this block is redone.
      auto row_max = [&]() {
        float mx4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float m0 = fmaxf(__uint_as_float(s[c][0]), __uint_as_float(s[c][1]));
#pragma unroll
          for (int i = 2; i < 32; i += 2) m0 = fmaxf(fmaxf(m0, __uint_as_float(s[c][i])), __uint_as_float(s[c][i + 1]));
          mx4[c] = m0;
        }
        return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      };
      if (j == 0) m_used = row_max();
      // ---- a maximum raised by the previous block: rescale l now, O once PV_{j-1} has completed
      float alpha = 1.0f;
      if (m_next > m_used) {
        alpha = ptx::ex2_approx((m_used - m_next) * sl2);
        m_used = m_next;
        l_run *= alpha;
      }
      att_stamp(args, tl && warp == 0, 0, j, 3);   // scale known
      auto rescale_o = [&](float a) {
#pragma unroll
        for (int c = 0; c < ATT_DH; c += 32) {
          uint32_t t[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + ATT_O_COL + c, t);
          ptx::tmem_ld_wait(t);
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a);
          ptx::tmem_st_32x32b_x32(lane_addr + ATT_O_COL + c, t);
        }
      };
      if (j > 0) {
        ptx::mbar_wait(o_full, (g - 1) & 1, 20);   // PV_{j-1} done: the P columns are free, O is complete up to j-1
        ptx::tc_fence_after();
        att_stamp(args, tl && warp == 0, 0, j, 4); // PV_{j-1} complete
        if (__any_sync(0xffffffffu, alpha != 1.0f)) rescale_o(alpha);
      }
      // ---- p = exp2(s*sl2 - m*sl2) -> bf16 pairs -> TMEM columns P_COL + key/2 of this thread's lane
      // (the A operand of the PV MMA).  Optionally some pairs take the FMA-pipe polynomial instead of MUFU.EX2
      // (full blocks only: the ragged block carries -inf).
      uint64_t sum2[2];
      auto exp_chunks = [&](auto poly_tag) {
        constexpr bool POLY = decltype(poly_tag)::value;
        const uint64_t nm2 = ptx::dup_f32x2(-m_used * sl2);
        sum2[0] = 0ull;
        sum2[1] = 0ull;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nchunks) {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t a2 = ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), sl2_2, nm2);
              float e0, e1;
              if (POLY && ((POLY_MASK >> i) & 1u)) {
                ptx::ex2_poly_x2(a2, e0, e1);
              } else {
                float a0, a1;
                ptx::unpack_f32x2(a2, a0, a1);
                e0 = ptx::ex2_approx(a0);
                e1 = ptx::ex2_approx(a1);
              }
              sum2[i & 1] = ptx::add_f32x2(sum2[i & 1], ptx::pack_f32x2(e0, e1));
              ph[i] = ptx::pack_bf16x2(e0, e1);
              if (SPLIT) pl[i] = ptx::pack_bf16x2(e0 - ptx::bf16_round(e0), e1 - ptx::bf16_round(e1));
            }
            ptx::tmem_st_32x32b_x16(lane_addr + ATT_P_COL + c * 16, ph);
            if (SPLIT) ptx::tmem_st_32x32b_x16(lane_addr + ATT_P_COL + 64 + c * 16, pl);
          }
        }
      };
      auto run_exps = [&]() {
        if (!SPLIT && POLY_MASK != 0 && kv_len == ATT_BKV) exp_chunks(std::true_type{}); else exp_chunks(std::false_type{});
      };
      run_exps();
      if (j > 0) {
        const float excess = (row_max() - m_used) * sl2;       // independent of the exponentials above: overlaps them
        if (__any_sync(0xffffffffu, excess > ATT_REDO_THRESHOLD)) {
          // rare: a row jumped so far above the running maximum that exp2 may have overflowed -> raise the maximum
          // now (O is complete up to block j-1 and may be rescaled here) and redo this block's exponentials
          const float m_new = excess > ATT_REDO_THRESHOLD ? m_used + excess / sl2 : m_used;
          const float a = ptx::ex2_approx((m_used - m_new) * sl2);
          m_used = m_new;
          l_run *= a;
          rescale_o(a);
          run_exps();
        } else if (excess > ATT_RESCALE_THRESHOLD) {
          m_next = m_used + excess / sl2;                        // applied before the next block
        }
      }
      {
        float a0, a1, b0, b1;
        ptx::unpack_f32x2(sum2[0], a0, a1);
        ptx::unpack_f32x2(sum2[1], b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
      att_stamp(args, tl && warp == 0, 0, j, 5);   // exponentials issued
      ptx::tmem_st_wait();             // P (and a rescaled O) are in TMEM
      ptx::tc_fence_before();          // ... and ordered before the MMA that reads / accumulates on them
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
      att_stamp(args, tl && warp == 0, 0, j, 6);   // P_j handed to the MMA warp
    }
    // ---- epilogue: ctx = O / l
    ptx::mbar_wait(o_full, (g - 1) & 1, 22);
    ptx::tc_fence_after();
    const float inv = 1.0f / l_run;
    const int qrow = qt * ATT_BQ + r;
    if (args.lse2 != nullptr)   // pad rows get +inf: the backward turns that into P = 0 without a bounds test
      args.lse2[(static_cast<long long>(b) * args.heads + h) * (args.n_qtiles * ATT_BQ) + qrow] = qrow < N ? m_used * sl2 + log2f(l_run) : INFINITY;
    __nv_bfloat16* o = args.out + static_cast<long long>(row_base + qrow) * args.ldo + h * ATT_DH;
    uint32_t t[ATT_DH / 32][32];
#pragma unroll
    for (int c = 0; c < ATT_DH / 32; ++c) ptx::tmem_ld_32x32b_x32(lane_addr + ATT_O_COL + c * 32, t[c]);
#pragma unroll
    for (int c = 0; c < ATT_DH / 32; ++c) ptx::tmem_ld_wait(t[c]);
    // O is in registers: the MMA warp may start accumulating the next work item into the same columns
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(o_empty);
    if (qrow < N) {
#pragma unroll
      for (int c = 0; c < ATT_DH / 32; ++c) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(t[c][8 * q4 + i]) * inv;
          reinterpret_cast<uint4*>(o + c * 32)[q4] = make_uint4(ptx::pack_bf16x2(v[0], v[1]), ptx::pack_bf16x2(v[2], v[3]),
                                                               ptx::pack_bf16x2(v[4], v[5]), ptx::pack_bf16x2(v[6], v[7]));
          if (SPLIT) {
            float lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) lo[i] = v[i] - ptx::bf16_round(v[i]);
            reinterpret_cast<uint4*>(o + args.out_lo_off + c * 32)[q4] =
                make_uint4(ptx::pack_bf16x2(lo[0], lo[1]), ptx::pack_bf16x2(lo[2], lo[3]), ptx::pack_bf16x2(lo[4], lo[5]),
                           ptx::pack_bf16x2(lo[6], lo[7]));
          }
        }
      }
    }
    }   // work items
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (tl0 && threadIdx.x == 0) args.timeline[((blockIdx.x * 2 + 0) * ATT_TL_BLOCKS + 15) * ATT_TL_EVENTS + 7] = clock64();   // CTA end
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace vitocm
