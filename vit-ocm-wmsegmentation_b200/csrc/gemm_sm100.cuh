// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//     C[M,N] = epilogue( sum_t A[:, a_koff[t] : +K] * B[:, b_koff[t] : +K]^T )
// A [M, lda] and B [N, ldb] are bf16, K-major (row-major with K contiguous), staged by TMA into
// 128B-swizzled shared-memory tiles; accumulators live in TMEM (2 stages, so the epilogue of
// tile i overlaps the main loop of tile i+1).  `nterms` = 3 with hi/lo operand halves gives the
// split-bf16 ("fp32 mode") product  hi*hi + hi*lo + lo*hi  in the same kernel.
//
// Replaces, on the reference path, nn.Linear at SSS/dino/vision_transformer.py:58,61 (fc1/fc2),
// :80 (qkv), :88 (proj) with their bias / GELU (:59) / residual (:110-111) fused as epilogues.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 = idle, warps 4..11 = epilogue (lane quadrant = warp % 4, column half = (warp - 4) / 4).
#pragma once
#include "ptx.cuh"

namespace vitocm {

enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,       // out_bf16 = acc + bias                (qkv / k projection)
  EPI_BIAS_GELU_BF16 = 1,  // out_bf16 = gelu_erf(acc + bias)      (fc1)
  EPI_BIAS_RESID_F32 = 2,  // resid_f32 += acc + bias              (proj, fc2)
  EPI_BIAS_F32 = 3,        // out_f32 = acc + bias                 (generic / decoder)
};

struct GemmArgs {
  int M, N;
  int kblocks;        // K / 64 per term
  int nterms;         // 1 (bf16 mode) or 3 (split mode: hi*hi + hi*lo + lo*hi)
  int lo_k;           // split mode: column where the lo halves of A and B start (= K)
  const float* bias;  // [N] or nullptr
  void* out;          // bf16 or f32, row-major, leading dimension ldo
  long long ldo;
  int split_out;      // bf16 outputs only: also write lo = bf16(v - hi) at column offset lo_off
  int lo_off;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARP0 = 4;

constexpr int GEMM_NUM_EPI_WARPS = 8;
constexpr int GEMM_SMEM_LIMIT = 227 * 1024;

template <int BN, int EPI>
struct GemmCfg {
  static constexpr bool OUT_BF16 = (EPI == 0 || EPI == 1);
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;       // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue staging: one [32 rows][32 cols] chunk per warp, rows padded by 16 B (conflict-free 128-bit stores)
  static constexpr int STG_PITCH = (OUT_BF16 ? 64 : 128) + 16;
  static constexpr int STG_WARP_BYTES = 32 * STG_PITCH;
  static constexpr int STG_BYTES = GEMM_NUM_EPI_WARPS * STG_WARP_BYTES;
  static constexpr int BIAS_WARP_FLOATS = 128;                 // >= BN / 2
  static constexpr int BIAS_BYTES = GEMM_NUM_EPI_WARPS * BIAS_WARP_FLOATS * 4;
  static constexpr int FIXED_BYTES = STG_BYTES + BIAS_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int STAGES_FIT = (GEMM_SMEM_LIMIT - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
  static_assert(STAGES >= 3, "not enough shared memory for a 3-stage pipeline");
};

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() default (approximate='none'): x * Phi(x)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmArgs args) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_stg = smem + STAGES * Cfg::STAGE_BYTES;
  float* smem_bias = reinterpret_cast<float*>(smem_stg + Cfg::STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + Cfg::STG_BYTES + Cfg::BIAS_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]       epilogue -> MMA
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (args.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = args.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int k_iters = args.kblocks * args.nterms;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 8);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * GEMM_BM;
        const int n0 = (tile % tiles_n) * BN;
        for (int it = 0; it < k_iters; ++it) {
          const int term = it / args.kblocks;
          const int kb = it - term * args.kblocks;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          // split mode terms: (hi,hi) (hi,lo) (lo,hi); lo halves start at column K of each operand
          const int a_off = (term == 2 ? args.lo_k : 0) + kb * GEMM_BK;
          const int b_off = (term == 1 ? args.lo_k : 0) + kb * GEMM_BK;
          ptx::tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], a_off, m0);
          ptx::tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmap_b, &full_bar[stage], b_off, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc(GEMM_BM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1, 2);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int it = 0; it < k_iters; ++it) {
          ptx::mbar_wait(&full_bar[stage], phase, 3);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr + k * 32, 1024, 0);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 32, 1024, 0);
            ptx::umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[as]);  // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ===================== epilogue =====================
    // TMEM -> registers (lane = row) -> +bias/GELU -> per-warp smem transpose -> global stores in which
    // consecutive lanes cover consecutive 16-byte pieces of a row (full 32-byte sectors, no partial writes).
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    const int ew = warp - GEMM_EPI_WARP0;           // 0..7
    const int half = ew >> 2;                       // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    constexpr int PITCH = Cfg::STG_PITCH;
    uint8_t* stg = smem_stg + ew * Cfg::STG_WARP_BYTES;
    float* bias_s = smem_bias + ew * Cfg::BIAS_WARP_FLOATS;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * GEMM_BM;
      const int n0 = (tile % tiles_n) * BN;
      const int row_base = m0 + q * 32;
      const int col_base = n0 + half * COLS_PER_WARP;
      // this warp's slice of the bias, fetched before waiting on the accumulator
      for (int j = lane; j < COLS_PER_WARP; j += 32) bias_s[j] = (args.bias != nullptr) ? __ldg(args.bias + col_base + j) : 0.f;
      __syncwarp();
      ptx::mbar_wait(&tfull_bar[as], aphase, 4);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        // residual epilogue: the 8 row-pieces of x this lane will update are fetched first, so their
        // latency hides behind the TMEM load and the smem transpose (lane -> row it*4 + lane/8, piece lane%8)
        float4 xres[8];
        if (EPI == EPI_BIAS_RESID_F32) {
          const float* xbase = reinterpret_cast<const float*>(args.out) + col_base + c + (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = row_base + it * 4 + (lane >> 3);
            xres[it] = (row < args.M) ? *reinterpret_cast<const float4*>(xbase + static_cast<long long>(row) * args.ldo)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN + half * COLS_PER_WARP + c), r);
        ptx::tmem_ld_wait(r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __uint_as_float(r[j]) + bias_s[c + j];
          if (EPI == EPI_BIAS_GELU_BF16) v[j] = gelu_erf(v[j]);
        }
        const int col = col_base + c;
        if (Cfg::OUT_BF16) {
          const int npass = args.split_out ? 2 : 1;
          for (int pass = 0; pass < npass; ++pass) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              w[j] = pass == 0 ? ptx::pack_bf16x2(v[2 * j], v[2 * j + 1])
                               : ptx::pack_bf16x2(v[2 * j] - ptx::bf16_round(v[2 * j]), v[2 * j + 1] - ptx::bf16_round(v[2 * j + 1]));
            uint4* srow = reinterpret_cast<uint4*>(stg + lane * PITCH);
#pragma unroll
            for (int j = 0; j < 4; ++j) srow[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            __syncwarp();
            __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(args.out) + col + (pass ? args.lo_off : 0);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rr = it * 8 + (lane >> 2), piece = lane & 3;
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * PITCH + piece * 16);
              const int row = row_base + rr;
              if (row < args.M) *reinterpret_cast<uint4*>(obase + static_cast<long long>(row) * args.ldo + piece * 8) = val;
            }
            __syncwarp();
          }
        } else {
          float4* srow = reinterpret_cast<float4*>(stg + lane * PITCH);
#pragma unroll
          for (int j = 0; j < 8; ++j) srow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          float* obase = reinterpret_cast<float*>(args.out) + col;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3), piece = lane & 7;
            float4 val = *reinterpret_cast<const float4*>(stg + rr * PITCH + piece * 16);
            const int row = row_base + rr;
            if (row < args.M) {
              float4* o = reinterpret_cast<float4*>(obase + static_cast<long long>(row) * args.ldo + piece * 4);
              if (EPI == EPI_BIAS_RESID_F32) {
                val.x += xres[it].x; val.y += xres[it].y; val.z += xres[it].z; val.w += xres[it].w;
              }
              *o = val;
            }
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace vitocm
