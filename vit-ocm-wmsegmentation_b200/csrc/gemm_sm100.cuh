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
  int nterms;         // 1 (bf16 mode) or 3 (split mode)
  int a_koff[3];      // element offset along K into A for each term
  int b_koff[3];      // element offset along K into B for each term
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

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;       // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192 ? 5 : 6);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() default (approximate='none'): x * Phi(x)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
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
          ptx::tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], args.a_koff[term] + kb * GEMM_BK, m0);
          ptx::tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmap_b, &full_bar[stage], args.b_koff[term] + kb * GEMM_BK, n0);
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
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    const int half = (warp - GEMM_EPI_WARP0) >> 2;  // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * GEMM_BM;
      const int n0 = (tile % tiles_n) * BN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < args.M;
      ptx::mbar_wait(&tfull_bar[as], aphase, 4);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        const int col_t = half * COLS_PER_WARP + c;  // column inside the tile
        const int col = n0 + col_t;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN + col_t), r);
        ptx::tmem_ld_wait(r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __uint_as_float(r[j]);
          if (args.bias != nullptr) v[j] += __ldg(args.bias + col + j);
          if (EPI == EPI_BIAS_GELU_BF16) v[j] = gelu_erf(v[j]);
        }
        if (row_ok) {
          if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + static_cast<long long>(row) * args.ldo + col;
            uint32_t hi[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) hi[j] = ptx::pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              reinterpret_cast<uint4*>(o)[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            if (args.split_out) {
              uint32_t lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j)
                lo[j] = ptx::pack_bf16x2(v[2 * j] - ptx::bf16_round(v[2 * j]), v[2 * j + 1] - ptx::bf16_round(v[2 * j + 1]));
#pragma unroll
              for (int j = 0; j < 4; ++j)
                reinterpret_cast<uint4*>(o + args.lo_off)[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          } else if (EPI == EPI_BIAS_RESID_F32) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + static_cast<long long>(row) * args.ldo + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 x = o[j];
              x.x += v[4 * j]; x.y += v[4 * j + 1]; x.z += v[4 * j + 2]; x.w += v[4 * j + 3];
              o[j] = x;
            }
          } else {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + static_cast<long long>(row) * args.ldo + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
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
