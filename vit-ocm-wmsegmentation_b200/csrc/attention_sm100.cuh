// Flash-style multi-head self-attention for sm_100a (head_dim 64), tcgen05 + TMEM + TMA:
//     ctx[b, i, h*64:(h+1)*64] = softmax_j( scale * q_i . k_j ) @ v        (never materialises N x N)
// Replaces SSS/dino/vision_transformer.py:83-87 (q@k^T*scale, softmax, attn@v, transpose/reshape)
// for the blocks whose attention matrix is not returned.
//
// One CTA = one (image b, head h, 128-query tile).  q/k/v are read straight out of the fused
// QKV activation [B*N, ld] (bf16, columns [3][H][64]) by one 2-D TMA tensor map (box 64 x 128,
// SWIZZLE_128B).  Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM
// allocator, warps 2..5 = softmax (one thread per query row; TMEM lane quadrant = warp % 4).
//   S = Q K_j^T      : tcgen05.mma  M128 x N(<=128) x K64, both operands K-major, into TMEM
//   softmax          : tcgen05.ld S (pass 1 row max, pass 2 exp2), P -> smem (bf16, swizzled)
//   O_j = P V_j      : tcgen05.mma  M128 x N64 x K(<=128), A = P (K-major), B = V (MN-major)
//   running output   : registers (fp32), rescaled by exp2(m_old - m_new) per KV block
// Two CTAs are co-resident per SM (96 KB smem, 256 TMEM columns each) so one CTA's MMAs overlap
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
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr int ATT_RING = 3;
constexpr int ATT_TMEM_COLS = 256;  // S: 128 cols, O: 64 cols
constexpr int ATT_S_COL = 0;
constexpr int ATT_O_COL = 128;

template <bool SPLIT>
struct AttnCfg {
  static constexpr int NPART = SPLIT ? 2 : 1;                     // hi (+ lo)
  static constexpr int SLOT_BYTES = ATT_TILE_BYTES * NPART;       // one K or V block
  static constexpr int Q_BYTES = ATT_TILE_BYTES * NPART;
  static constexpr int P_BYTES = 2 * ATT_TILE_BYTES * NPART;      // [128 x 128] bf16 (two 64-key halves)
  static constexpr int SMEM_BYTES = Q_BYTES + ATT_RING * SLOT_BYTES + P_BYTES + 1024 + 128;
};

template <bool SPLIT>
__global__ void __launch_bounds__(ATT_THREADS, SPLIT ? 1 : 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs args) {
  using Cfg = AttnCfg<SPLIT>;
  constexpr int NPART = Cfg::NPART;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_ring = smem_q + Cfg::Q_BYTES;
  uint8_t* smem_p = smem_ring + ATT_RING * Cfg::SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + Cfg::P_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]
  uint64_t* s_full = bars + 7;        // MMA -> softmax
  uint64_t* s_empty = bars + 8;       // softmax -> MMA   (4 warps)
  uint64_t* p_full = bars + 9;        // softmax -> MMA   (4 warps)
  uint64_t* o_full = bars + 10;       // MMA -> softmax
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int N = args.n_tokens;
  const int D = args.embed_dim;
  const int n_kv = (N + ATT_BKV - 1) / ATT_BKV;
  const int row_base = b * N;  // first row of this image in the [B*N, ld] activation

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < ATT_RING; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 4);
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, ATT_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
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
        ptx::mbar_wait(&kv_empty[slot], parity, 10);
        ptx::mbar_arrive_expect_tx(&kv_full[slot], Cfg::SLOT_BYTES);
        for (int part = 0; part < NPART; ++part)
          ptx::tma_load_2d(smem_ring + slot * Cfg::SLOT_BYTES + part * ATT_TILE_BYTES, &tmap_qkv, &kv_full[slot],
                           part * args.lo_col_off + which * D + h * ATT_DH, row_base + j * ATT_BKV);
        ++item;
      };
      load(1, 0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) load(1, j + 1);
        load(2, j);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t q_addr = ptx::smem_u32(smem_q);
      const uint32_t p_addr = ptx::smem_u32(smem_p);
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
        ptx::mbar_wait(&kv_full[slot], (item / ATT_RING) & 1, 11);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(smem_ring + slot * Cfg::SLOT_BYTES);
        const uint32_t idesc = ptx::make_idesc(ATT_BQ, kv_len_mma(j), false, false);
        uint32_t acc = 0;
        // terms: (Qhi,Khi) [, (Qhi,Klo), (Qlo,Khi)]
        for (int t = 0; t < (SPLIT ? 3 : 1); ++t) {
          const uint32_t qa = q_addr + (t == 2 ? ATT_TILE_BYTES : 0);
          const uint32_t ka = k_addr + (t == 1 ? ATT_TILE_BYTES : 0);
#pragma unroll
          for (int k = 0; k < ATT_DH / 16; ++k) {
            ptx::umma_bf16_ss(s_tmem, ptx::make_smem_desc_sw128(qa + k * 32, 1024, 0),
                              ptx::make_smem_desc_sw128(ka + k * 32, 1024, 0), idesc, acc);
            acc = 1;
          }
        }
        ptx::umma_commit(&kv_empty[slot]);
        ptx::umma_commit(s_full);
        ++item;
      };
      auto issue_pv = [&](int j) {
        const int slot = item % ATT_RING;
        ptx::mbar_wait(&kv_full[slot], (item / ATT_RING) & 1, 12);
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(smem_ring + slot * Cfg::SLOT_BYTES);
        constexpr uint32_t idesc = ptx::make_idesc(ATT_BQ, ATT_DH, false, /*B = V is MN-major*/ true);
        const int ksteps = kv_len_mma(j) / 16;
        uint32_t acc = 0;
        // terms: (Phi,Vhi) [, (Phi,Vlo), (Plo,Vhi)]
        for (int t = 0; t < (SPLIT ? 3 : 1); ++t) {
          const uint32_t pa = p_addr + (t == 2 ? 2 * ATT_TILE_BYTES : 0);
          const uint32_t va = v_addr + (t == 1 ? ATT_TILE_BYTES : 0);
          for (int k = 0; k < ksteps; ++k) {
            // A = P: K-major, 64-key halves of 16 KB, 32 B per 16-key step inside the 128 B swizzle row
            const uint64_t adesc = ptx::make_smem_desc_sw128(pa + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 1024, 0);
            // B = V: MN-major [keys x 64]; 16 keys = two 8-row groups of 1024 B
            const uint64_t bdesc = ptx::make_smem_desc_sw128(va + k * 2048, 1024, 1024);
            ptx::umma_bf16_ss(o_tmem, adesc, bdesc, idesc, acc);
            acc = 1;
          }
        }
        ptx::umma_commit(&kv_empty[slot]);
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
        ptx::mbar_wait(p_full, j & 1, 15);     // P_j in smem, O_{j-1} consumed
        ptx::tc_fence_after();
        issue_pv(j);
      }
    }
  } else {
    // ===================== softmax / output (warps 2..5) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float sl2 = args.scale_log2;
    float o_acc[ATT_DH];
#pragma unroll
    for (int d = 0; d < ATT_DH; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    auto fetch_o = [&](int j) {  // o_acc += O_j  (PV of block j, in units of the current m_run)
      ptx::mbar_wait(o_full, j & 1, 20);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < ATT_DH; c += 32) {
        uint32_t t[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + lane_addr + ATT_O_COL + c, t);
        ptx::tmem_ld_wait(t);
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c + i] += __uint_as_float(t[i]);
      }
    };

    for (int j = 0; j < n_kv; ++j) {
      int kv_len = N - j * ATT_BKV;
      kv_len = kv_len > ATT_BKV ? ATT_BKV : kv_len;
      const int ncols = (kv_len + 15) & ~15;       // columns the MMA produced
      const int nchunks = (ncols + 31) >> 5;
      ptx::mbar_wait(s_full, j & 1, 21);
      ptx::tc_fence_after();
      // ---- pass 1: row maximum
      float mx = m_run;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t t[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + lane_addr + ATT_S_COL + c * 32, t);
        ptx::tmem_ld_wait(t);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kv_len) mx = fmaxf(mx, __uint_as_float(t[i]));
      }
      const float alpha = exp2f((m_run - mx) * sl2);
      // ---- previous block's PV result (also: P buffer is free again)
      if (j > 0) fetch_o(j - 1);
#pragma unroll
      for (int d = 0; d < ATT_DH; ++d) o_acc[d] *= alpha;
      l_run *= alpha;
      m_run = mx;
      const float m_sl2 = mx * sl2;
      // ---- pass 2: p = exp2(s*sl2 - m*sl2) -> bf16 -> swizzled smem (A operand of the PV MMA)
      float psum = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t t[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + lane_addr + ATT_S_COL + c * 32, t);
        ptx::tmem_ld_wait(t);
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = exp2f(fmaf(__uint_as_float(t[i]), sl2, -m_sl2));
          p[i] = (c * 32 + i < kv_len) ? e : 0.f;
          psum += p[i];
        }
        // 32 keys = 64 B = four 16-byte chunks of this row; chunk index inside the 128 B row is
        // XOR-swizzled with (row % 8) (SWIZZLE_128B, tile base 1024-aligned)
        uint8_t* half_base = smem_p + (c >> 1) * ATT_TILE_BYTES + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk16 = ((c & 1) * 4 + g) ^ (r & 7);
          uint4 hi;
          hi.x = ptx::pack_bf16x2(p[8 * g + 0], p[8 * g + 1]);
          hi.y = ptx::pack_bf16x2(p[8 * g + 2], p[8 * g + 3]);
          hi.z = ptx::pack_bf16x2(p[8 * g + 4], p[8 * g + 5]);
          hi.w = ptx::pack_bf16x2(p[8 * g + 6], p[8 * g + 7]);
          *reinterpret_cast<uint4*>(half_base + chunk16 * 16) = hi;
          if (SPLIT) {
            uint4 lo;
            lo.x = ptx::pack_bf16x2(p[8 * g + 0] - ptx::bf16_round(p[8 * g + 0]), p[8 * g + 1] - ptx::bf16_round(p[8 * g + 1]));
            lo.y = ptx::pack_bf16x2(p[8 * g + 2] - ptx::bf16_round(p[8 * g + 2]), p[8 * g + 3] - ptx::bf16_round(p[8 * g + 3]));
            lo.z = ptx::pack_bf16x2(p[8 * g + 4] - ptx::bf16_round(p[8 * g + 4]), p[8 * g + 5] - ptx::bf16_round(p[8 * g + 5]));
            lo.w = ptx::pack_bf16x2(p[8 * g + 6] - ptx::bf16_round(p[8 * g + 6]), p[8 * g + 7] - ptx::bf16_round(p[8 * g + 7]));
            *reinterpret_cast<uint4*>(half_base + 2 * ATT_TILE_BYTES + chunk16 * 16) = lo;
          }
        }
      }
      l_run += psum;
      // S_j fully read: the MMA warp may overwrite it with S_{j+1}
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_empty);
      // P_j visible to the async proxy (tcgen05.mma reads smem through it)
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    fetch_o(n_kv - 1);
    const int qrow = qt * ATT_BQ + r;
    if (qrow < N) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* o = args.out + static_cast<long long>(row_base + qrow) * args.ldo + h * ATT_DH;
#pragma unroll
      for (int g = 0; g < ATT_DH / 8; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o_acc[8 * g + i] * inv;
        reinterpret_cast<uint4*>(o)[g] = make_uint4(ptx::pack_bf16x2(v[0], v[1]), ptx::pack_bf16x2(v[2], v[3]),
                                                    ptx::pack_bf16x2(v[4], v[5]), ptx::pack_bf16x2(v[6], v[7]));
        if (SPLIT) {
          float w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = v[i] - ptx::bf16_round(v[i]);
          reinterpret_cast<uint4*>(o + args.out_lo_off)[g] =
              make_uint4(ptx::pack_bf16x2(w[0], w[1]), ptx::pack_bf16x2(w[2], w[3]), ptx::pack_bf16x2(w[4], w[5]),
                         ptx::pack_bf16x2(w[6], w[7]));
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace vitocm
