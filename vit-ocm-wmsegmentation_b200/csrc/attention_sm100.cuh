// Flash-style multi-head self-attention for sm_100a (head_dim 64), tcgen05 + TMEM + TMA:
//     ctx[b, i, h*64:(h+1)*64] = softmax_j( scale * q_i . k_j ) @ v        (never materialises N x N)
// Replaces SSS/dino/vision_transformer.py:83-87 (q@k^T*scale, softmax, attn@v, transpose/reshape)
// for the blocks whose attention matrix is not returned.
//
// One CTA = one (image b, head h, 128-query tile).  q/k/v are read straight out of the fused
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
#include "ptx.cuh"

namespace vitocm {

struct AttnArgs {
  int n_tokens;      // N per image (785 for 224^2 / patch 8)
  int embed_dim;     // D = H * 64
  int lo_col_off;    // SPLIT: column offset of the lo halves inside the qkv activation (= 3D)
  float scale_log2;  // qk scale * log2(e)
  __nv_bfloat16* out;  // ctx [B*N, ldo]
  long long ldo;
  int out_lo_off;    // SPLIT: column offset of the lo half of ctx
};

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_DH = 64;
constexpr int ATT_THREADS = 256;   // warpgroup 0 = softmax (warps 0..3), warpgroup 1 = TMA (warp 4) + MMA (warp 5)
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr int ATT_RING = 3;
constexpr int ATT_S_COL = 0;      // S: 128 columns (fp32)
constexpr int ATT_O_COL = 128;    // O: 64 columns (fp32)
constexpr int ATT_P_COL = 192;    // P: 64 columns of packed bf16x2 (128 keys); split mode: lo part in the next 64
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units

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

template <bool SPLIT>
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
  const uint32_t tmem_ptr_smem = bars + 88;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int N = args.n_tokens;
  const int D = args.embed_dim;
  const int n_kv = (N + ATT_BKV - 1) / ATT_BKV;
  const int row_base = b * N;  // first row of this image in the [B*N, ld] activation

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

  if (warp >= 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_OTHER));
   if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
      for (int part = 0; part < NPART; ++part)
        ptx::tma_load_2d(smem_q + part * ATT_TILE_BYTES, &tmap_qkv, q_full, part * args.lo_col_off + h * ATT_DH,
                         row_base + qt * ATT_BQ);
      // ring order = consumption order of the MMA warp: K0, K1, V0, K2, V1, ..., V_{n-1}
      int item = 0;
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
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t s_tmem = tmem_base + ATT_S_COL;
      const uint32_t o_tmem = tmem_base + ATT_O_COL;
      int item = 0;
      auto kv_len_mma = [&](int j) {  // keys of block j rounded up to the MMA granularity (16)
        int len = N - j * ATT_BKV;
        len = len > ATT_BKV ? ATT_BKV : len;
        return (len + 15) & ~15;
      };
      auto issue_s = [&](int j) {
        const int slot = item % ATT_RING;
        ptx::mbar_wait(kv_full + 8 * slot, (item / ATT_RING) & 1, 11);
        ptx::tc_fence_after();
        const uint32_t k_addr = smem_ring + slot * Cfg::SLOT_BYTES;
        const uint32_t idesc = ptx::make_idesc(ATT_BQ, kv_len_mma(j), false, false);
        uint32_t acc = 0;
        // terms: (Qhi,Khi) [, (Qhi,Klo), (Qlo,Khi)]
#pragma unroll 1
        for (int t = 0; t < (SPLIT ? 3 : 1); ++t) {
          const uint32_t qa = smem_q + (t == 2 ? ATT_TILE_BYTES : 0);
          const uint32_t ka = k_addr + (t == 1 ? ATT_TILE_BYTES : 0);
#pragma unroll
          for (int k = 0; k < ATT_DH / 16; ++k) {
            ptx::umma_bf16_ss(s_tmem, ptx::make_smem_desc_sw128(qa + k * 32, 1024, 0),
                              ptx::make_smem_desc_sw128(ka + k * 32, 1024, 0), idesc, acc);
            acc = 1;
          }
        }
        ptx::umma_commit(kv_empty + 8 * slot);
        ptx::umma_commit(s_full);
        ++item;
      };
      auto issue_pv = [&](int j) {
        const int slot = item % ATT_RING;
        ptx::mbar_wait(kv_full + 8 * slot, (item / ATT_RING) & 1, 12);
        ptx::tc_fence_after();
        const uint32_t v_addr = smem_ring + slot * Cfg::SLOT_BYTES;
        constexpr uint32_t idesc = ptx::make_idesc(ATT_BQ, ATT_DH, false, /*B = V is MN-major*/ true);
        const int ksteps = kv_len_mma(j) / 16;
        // O accumulates across KV blocks in TMEM.  Loops are kept rolled: this warp runs on a small
        // register budget.
        // A = P in TMEM: 16 keys = 8 packed columns per step
        // B = V: MN-major [keys x 64]; 16 keys = two 8-row groups of 1024 B
        uint32_t acc = j > 0 ? 1u : 0u;
        // terms: (Phi,Vhi) [, (Phi,Vlo), (Plo,Vhi)]
#pragma unroll 1
        for (int t = 0; t < (SPLIT ? 3 : 1); ++t) {
          const uint32_t pa = tmem_base + ATT_P_COL + (t == 2 ? 64 : 0);
          const uint32_t va = v_addr + (t == 1 ? ATT_TILE_BYTES : 0);
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k) {
            ptx::umma_bf16_ts(o_tmem, pa + k * 8, ptx::make_smem_desc_sw128(va + k * 2048, 1024, 1024), idesc, acc);
            acc = 1;
          }
        }
        ptx::umma_commit(kv_empty + 8 * slot);
        ptx::umma_commit(o_full);
        ++item;
      };
      ptx::mbar_wait(q_full, 0, 13);
      ptx::tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          ptx::mbar_wait(s_empty, j & 1, 14);  // softmax has read S_j out of TMEM
          ptx::tc_fence_after();
          issue_s(j + 1);
        }
        ptx::mbar_wait(p_full, j & 1, 15);     // P_j in smem, O rescaled if needed
        ptx::tc_fence_after();
        issue_pv(j);
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
    float m_used = -INFINITY;     // the row maximum the exponentials are taken against
    float l_run = 0.f;            // running row sum (same units as O in TMEM)

    for (int j = 0; j < n_kv; ++j) {
      int kv_len = N - j * ATT_BKV;
      kv_len = kv_len > ATT_BKV ? ATT_BKV : kv_len;
      const int nchunks = (((kv_len + 15) & ~15) + 31) >> 5;   // 32-column chunks the MMA produced
      ptx::mbar_wait(s_full, j & 1, 21);
      ptx::tc_fence_after();
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
      // ---- ragged last block only: columns beyond the sequence (and chunks the MMA never wrote) -> -inf,
      //      so the common path below carries no masks (exp2(-inf) = 0)
      if (kv_len < ATT_BKV) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= kv_len) s[c][i] = 0xff800000u;
      }
      // ---- row maximum: four independent 3-input chains
      float mx4[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float m0 = fmaxf(__uint_as_float(s[c][0]), __uint_as_float(s[c][1]));
#pragma unroll
        for (int i = 2; i < 32; i += 2) m0 = fmaxf(fmaxf(m0, __uint_as_float(s[c][i])), __uint_as_float(s[c][i + 1]));
        mx4[c] = m0;
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // ---- lazy rescale decision
      float alpha = 1.0f;
      if (j == 0) {
        m_used = mx;
      } else if ((mx - m_used) * sl2 > ATT_RESCALE_THRESHOLD) {
        alpha = ptx::ex2_approx((m_used - mx) * sl2);
        m_used = mx;
        l_run *= alpha;
      }
      const uint64_t nm2 = ptx::dup_f32x2(-m_used * sl2);
      // ---- previous PV done: P buffer free, O valid -> rescale it if this warp raised a maximum
      if (j > 0) {
        ptx::mbar_wait(o_full, (j - 1) & 1, 20);
        ptx::tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int c = 0; c < ATT_DH; c += 32) {
            uint32_t t[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + ATT_O_COL + c, t);
            ptx::tmem_ld_wait(t);
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            ptx::tmem_st_32x32b_x32(lane_addr + ATT_O_COL + c, t);
          }
        }
      }
      // ---- p = exp2(s*sl2 - m*sl2) -> bf16 pairs -> TMEM columns P_COL + key/2 of this thread's lane
      // (the A operand of the PV MMA)
      uint64_t sum2[2] = {0ull, 0ull};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunks) {
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t a2 = ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), sl2_2, nm2);
            float a0, a1;
            ptx::unpack_f32x2(a2, a0, a1);
            const float e0 = ptx::ex2_approx(a0), e1 = ptx::ex2_approx(a1);
            sum2[i & 1] = ptx::add_f32x2(sum2[i & 1], ptx::pack_f32x2(e0, e1));
            ph[i] = ptx::pack_bf16x2(e0, e1);
            if (SPLIT) pl[i] = ptx::pack_bf16x2(e0 - ptx::bf16_round(e0), e1 - ptx::bf16_round(e1));
          }
          ptx::tmem_st_32x32b_x16(lane_addr + ATT_P_COL + c * 16, ph);
          if (SPLIT) ptx::tmem_st_32x32b_x16(lane_addr + ATT_P_COL + 64 + c * 16, pl);
        }
      }
      {
        float a0, a1, b0, b1;
        ptx::unpack_f32x2(sum2[0], a0, a1);
        ptx::unpack_f32x2(sum2[1], b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
      ptx::tmem_st_wait();             // P (and a rescaled O) are in TMEM
      ptx::tc_fence_before();          // ... and ordered before the MMA that reads / accumulates on them
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    // ---- epilogue: ctx = O / l
    ptx::mbar_wait(o_full, (n_kv - 1) & 1, 22);
    ptx::tc_fence_after();
    const float inv = 1.0f / l_run;
    const int qrow = qt * ATT_BQ + r;
    __nv_bfloat16* o = args.out + static_cast<long long>(row_base + qrow) * args.ldo + h * ATT_DH;
#pragma unroll
    for (int c = 0; c < ATT_DH; c += 32) {
      uint32_t t[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + ATT_O_COL + c, t);
      ptx::tmem_ld_wait(t);
      if (qrow < N) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(t[8 * g + i]) * inv;
          reinterpret_cast<uint4*>(o + c)[g] = make_uint4(ptx::pack_bf16x2(v[0], v[1]), ptx::pack_bf16x2(v[2], v[3]),
                                                          ptx::pack_bf16x2(v[4], v[5]), ptx::pack_bf16x2(v[6], v[7]));
          if (SPLIT) {
            float w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = v[i] - ptx::bf16_round(v[i]);
            reinterpret_cast<uint4*>(o + args.out_lo_off + c)[g] =
                make_uint4(ptx::pack_bf16x2(w[0], w[1]), ptx::pack_bf16x2(w[2], w[3]), ptx::pack_bf16x2(w[4], w[5]),
                           ptx::pack_bf16x2(w[6], w[7]));
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace vitocm
