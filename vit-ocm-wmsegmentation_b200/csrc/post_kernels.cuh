// Attention post-processing, thresholding and sliding-window stitching (HBM-bound byte/float
// work; warp-shuffle reductions, shared-memory histograms, no tensor cores):
//   head_mean_kernel        SSS/utils.py:232-233 + SSS/eval.py:142 / SSS/sw_processing.py:245,253-254
//   tile_threshold_kernel   SSS/eval.py:169-173 + SSS/utils.py:62-115 (per-image "ours"/otsu/heatmap masks)
//   extract_tiles_kernel    SSS/sw_processing.py:151-163 + ToTensor (:236-237)
//   stitch_gray_kernel      SSS/sw_processing.py:113-149 on the uint8 image crops (:225)
//   stitch_* kernels        SSS/sw_processing.py:255-259 (resize pair + concat_crops on the maps) and
//                           SSS/sw_processing.py:37-61 (global min-max, img*att, Otsu x2)
//   otsu_kernel             cv2.threshold(THRESH_OTSU) scan (OpenCV, restated in oracle/post_oracle.py)
#pragma once
#include "ptx.cuh"

namespace vitocm {

// order-preserving float <-> int map for atomicMin/atomicMax
__device__ __forceinline__ int f2ord(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// cv2.resize INTER_LINEAR source coordinate / weight for destination index d (scale = src/dst)
__device__ __forceinline__ void linear_coeff(int d, double scale, int src, int& s0, int& s1, float& f) {
  double fx = (d + 0.5) * scale - 0.5;
  int s = static_cast<int>(floor(fx));
  fx -= s;
  if (s < 0) { s = 0; fx = 0.0; }
  if (s >= src - 1) { s = src - 1; fx = 0.0; }
  s0 = s;
  s1 = min(s + 1, src - 1);
  f = static_cast<float>(fx);
}
// bilinear sample of a low-res map at full-res pixel (y, x): horizontal pass then vertical pass,
// fp32 with separate rounding of every product and sum (matches the numpy/cv2 oracle bit for bit)
__device__ __forceinline__ float bilinear_up(const float* __restrict__ lo, int lh, int lw, int y, int x, double scale) {
  int x0, x1, y0, y1;
  float fx, fy;
  linear_coeff(x, scale, lw, x0, x1, fx);
  linear_coeff(y, scale, lh, y0, y1, fy);
  const float a0 = __fsub_rn(1.0f, fx), b0 = __fsub_rn(1.0f, fy);
  const float r0 = __fadd_rn(__fmul_rn(lo[y0 * lw + x0], a0), __fmul_rn(lo[y0 * lw + x1], fx));
  const float r1 = __fadd_rn(__fmul_rn(lo[y1 * lw + x0], a0), __fmul_rn(lo[y1 * lw + x1], fx));
  return __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, fy));
}

// ---------------------------------------------------------------------------------------
// CLS rows [T][H][N] -> low-res map [T][n]: mean over heads of columns 1..N-1 (sequential fp32
// adds then a divide, as numpy reduces over a leading axis); mode 1 additionally applies the
// per-tile min-max * 255 of SSS/sw_processing.py:253-254.  One block per tile.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_mean_kernel(const float* __restrict__ rows, float* __restrict__ lowres, int heads, int N, int mode) {
  __shared__ float red_mn[8], red_mx[8];
  const int t = blockIdx.x, n = N - 1;
  const float* rt = rows + static_cast<long long>(t) * heads * N;
  float* lt = lowres + static_cast<long long>(t) * n;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = rt[1 + i];
    for (int h = 1; h < heads; ++h) s = __fadd_rn(s, rt[static_cast<long long>(h) * N + 1 + i]);
    s = __fdiv_rn(s, static_cast<float>(heads));
    lt[i] = s;
    mn = fminf(mn, s);
    mx = fmaxf(mx, s);
  }
  if (mode == 0) return;
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { red_mn[threadIdx.x >> 5] = mn; red_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  mn = red_mn[0]; mx = red_mx[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) { mn = fminf(mn, red_mn[w]); mx = fmaxf(mx, red_mx[w]); }
  const float range = __fsub_rn(mx, mn);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    lt[i] = __fmul_rn(__fdiv_rn(__fsub_rn(lt[i], mn), range), 255.0f);
}

// ---------------------------------------------------------------------------------------
// Otsu threshold from a 256-bin histogram (OpenCV's scan, fp64, first maximum wins).
// ---------------------------------------------------------------------------------------
__device__ inline int otsu_from_hist(const unsigned long long* hist) {
  double total = 0.0, wsum = 0.0;
  for (int i = 0; i < 256; ++i) { total += static_cast<double>(hist[i]); wsum += static_cast<double>(i) * static_cast<double>(hist[i]); }
  if (total <= 0.0) return 0;
  const double scale = 1.0 / total;
  const double mu = __dmul_rn(wsum, scale);
  double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
  int max_val = 0;
  const double eps = 1.1920928955078125e-07;
  for (int i = 0; i < 256; ++i) {
    const double p_i = __dmul_rn(static_cast<double>(hist[i]), scale);
    mu1 = __dmul_rn(mu1, q1);
    q1 = __dadd_rn(q1, p_i);
    const double q2 = __dsub_rn(1.0, q1);
    if (fmin(q1, q2) < eps || fmax(q1, q2) > 1.0 - eps) continue;
    mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn(static_cast<double>(i), p_i)), q1);
    const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
    const double dm = __dsub_rn(mu1, mu2);
    const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), dm), dm);
    if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
  }
  return max_val;
}

__global__ void otsu_kernel(const unsigned long long* __restrict__ hists, int nhist, int* __restrict__ thresholds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nhist) thresholds[i] = otsu_from_hist(hists + static_cast<long long>(i) * 256);
}

// ---------------------------------------------------------------------------------------
// Per-image threshold (eval / PGT flavour).  One block per tile, three passes over the tile:
//   1. att = bilinear_up(lowres) ; block min / max
//   2. att_u8 = trunc((att-min)/(max-min)*255) ; result = trunc((img/2)*0.6 + (att_u8/2)*0.4) (fp64)
//      -> three 256-bin histograms (result, img, att_u8) in shared memory -> Otsu x3
//   3. masks th / th2 / th3
// img = PIL "L" of ToPILImage(x): floor(x*255) per channel, then (19595 R + 38470 G + 7471 B + 32768) >> 16.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int gray_u8_from_x(const float* __restrict__ xt, int C, long long plane, long long off) {
  const int r = static_cast<int>(xt[off] * 255.0f);
  if (C < 3) return r;
  const int g = static_cast<int>(xt[plane + off] * 255.0f);
  const int b = static_cast<int>(xt[2 * plane + off] * 255.0f);
  return (19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16;
}

__global__ void __launch_bounds__(512)
tile_threshold_kernel(const float* __restrict__ lowres /*[T][lh][lw]*/, const float* __restrict__ x /*[T][C][S][S]*/,
                      int C, int S, int lh, int lw, uint8_t* __restrict__ masks /*[T][3][S][S]*/,
                      int* __restrict__ thresholds /*[T][3]*/, float* __restrict__ att_out /*[T][S][S] or null*/,
                      const float* __restrict__ att_in /*[T][S][S] or null: use instead of upsampling lowres*/,
                      const uint8_t* __restrict__ img_in /*[T][S][S] or null: use instead of deriving from x*/,
                      uint8_t* __restrict__ aux /*[T][2][S][S] or null: the blended image `result` and att_u8 (the images utils.threshold saves)*/) {
  __shared__ float red_mn[16], red_mx[16];
  __shared__ unsigned int hist[3][256];
  __shared__ unsigned long long hist64[256];
  __shared__ int thr[3];
  const int t = blockIdx.x;
  const float* lo = lowres + static_cast<long long>(t) * lh * lw;
  const long long plane = static_cast<long long>(S) * S;
  const float* xt = x + static_cast<long long>(t) * C * plane;
  const float* ain = att_in != nullptr ? att_in + static_cast<long long>(t) * plane : nullptr;
  const uint8_t* iin = img_in != nullptr ? img_in + static_cast<long long>(t) * plane : nullptr;
  const double scale = static_cast<double>(lw) / static_cast<double>(S);
  const int npx = S * S;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {
    const float a = ain != nullptr ? ain[i] : bilinear_up(lo, lh, lw, i / S, i % S, scale);
    if (att_out != nullptr) att_out[static_cast<long long>(t) * plane + i] = a;
    mn = fminf(mn, a);
    mx = fmaxf(mx, a);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { red_mn[threadIdx.x >> 5] = mn; red_mx[threadIdx.x >> 5] = mx; }
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) (&hist[0][0])[i] = 0u;
  __syncthreads();
  mn = red_mn[0]; mx = red_mx[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) { mn = fminf(mn, red_mn[w]); mx = fmaxf(mx, red_mx[w]); }
  const bool flat = (mx == mn);  // min_max_normalize returns its input when the map is flat
  const float range = __fsub_rn(mx, mn);
  auto classify = [&](int i, int& img, int& att_u8, int& res) {
    const float a = ain != nullptr ? ain[i] : bilinear_up(lo, lh, lw, i / S, i % S, scale);
    const float an = flat ? a : __fdiv_rn(__fsub_rn(a, mn), range);
    att_u8 = static_cast<int>(static_cast<uint8_t>(static_cast<int>(__fmul_rn(an, 255.0f))));
    img = iin != nullptr ? static_cast<int>(iin[i]) : gray_u8_from_x(xt, C, plane, i);
    const double r = __dadd_rn(__dmul_rn(static_cast<double>(img) / 2.0, 1.0 - 0.4), __dmul_rn(static_cast<double>(att_u8) / 2.0, 0.4));
    res = static_cast<int>(r);
  };
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {
    int img, au, res;
    classify(i, img, au, res);
    atomicAdd(&hist[0][res], 1u);
    atomicAdd(&hist[1][img], 1u);
    atomicAdd(&hist[2][au], 1u);
  }
  __syncthreads();
  for (int k = 0; k < 3; ++k) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist64[i] = hist[k][i];
    __syncthreads();
    if (threadIdx.x == 0) {
      thr[k] = otsu_from_hist(hist64);
      thresholds[t * 3 + k] = thr[k];
    }
    __syncthreads();
  }
  uint8_t* m0 = masks + static_cast<long long>(t) * 3 * plane;
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {
    int img, au, res;
    classify(i, img, au, res);
    m0[i] = res > thr[0] ? 255 : 0;
    m0[plane + i] = img > thr[1] ? 255 : 0;
    m0[2 * plane + i] = au > thr[2] ? 255 : 0;
    if (aux != nullptr) {
      aux[static_cast<long long>(t) * 2 * plane + i] = static_cast<uint8_t>(res);
      aux[static_cast<long long>(t) * 2 * plane + plane + i] = static_cast<uint8_t>(au);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Sliding-window geometry shared by the mosaic kernels.  n x n tiles of size W at stride S
// (origins range(0, size - 2S, S)); stitched extent E = (n-1) S + W; step = W - S.
// wtab[k] = numpy.linspace(1, 0, step)[k] (computed on the host so it is bit-identical).
// ---------------------------------------------------------------------------------------
struct StitchGeom {
  int n, W, S, step, E;
  int lh, lw;        // low-res map size (W / patch)
  double scale;      // lw / W
};

// sliding_window + ToTensor: mosaic u8 gray [E0][pitch] -> x [T][C][W][W] fp32 = v / 255 (zero padded
// outside the mosaic like PIL's crop); tiles t0 .. t0+T-1 in row-major order.
__global__ void extract_tiles_kernel(const uint8_t* __restrict__ mosaic, int mos_h, int mos_w, long long pitch, int n,
                                     int W, int S, int t0, int T, int C, float* __restrict__ x) {
  const long long per_tile = static_cast<long long>(W) * W;
  const long long total = per_tile * T;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tl = static_cast<int>(i / per_tile);
    const int r = static_cast<int>(i - tl * per_tile);
    const int y = r / W, xx = r - y * W;
    const int t = t0 + tl;
    const int gy = (t / n) * S + y, gx = (t % n) * S + xx;
    const float v = (gy < mos_h && gx < mos_w) ? __fdiv_rn(static_cast<float>(mosaic[gy * pitch + gx]), 255.0f) : 0.f;
    float* xt = x + static_cast<long long>(tl) * C * per_tile + r;
    for (int c = 0; c < C; ++c) xt[c * per_tile] = v;
  }
}

// uint8 blend of the image crops: sequential pairwise blend with truncation, first along x inside
// each strip, then along y across strips -- what concat_crops does to the uint8 crops at :225.
__device__ __forceinline__ uint8_t blend_u8(uint8_t a, uint8_t b, double w) {
  return static_cast<uint8_t>(static_cast<int>(__dadd_rn(__dmul_rn(static_cast<double>(a), w), __dmul_rn(static_cast<double>(b), __dsub_rn(1.0, w)))));
}
__global__ void stitch_gray_kernel(const uint8_t* __restrict__ mosaic, int mos_h, int mos_w, long long pitch, StitchGeom g,
                                   const double* __restrict__ wtab, int y_begin, int y_end, uint8_t* __restrict__ out /*[E][E]*/) {
  const long long total = static_cast<long long>(y_end - y_begin) * g.E;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int Y = y_begin + static_cast<int>(idx / g.E), X = static_cast<int>(idx % g.E);
    // every crop holds the same source pixel (zero beyond the mosaic)
    const uint8_t src = (Y < mos_h && X < mos_w) ? mosaic[Y * pitch + X] : 0;
    int j0 = (X - g.W + g.S) / g.S; if (X - g.W + 1 <= 0) j0 = 0; if (j0 < 0) j0 = 0;
    int j1 = min(X / g.S, g.n - 1);
    int i0 = (Y - g.W + g.S) / g.S; if (Y - g.W + 1 <= 0) i0 = 0; if (i0 < 0) i0 = 0;
    int i1 = min(Y / g.S, g.n - 1);
    // horizontal sequence is identical for every strip (same source value)
    uint8_t hv = src;
    for (int j = j0 + 1; j <= j1; ++j) {
      const int kx = X - j * g.S;
      hv = (kx < g.step) ? blend_u8(hv, src, wtab[kx]) : src;
    }
    uint8_t v = hv;
    for (int i = i0 + 1; i <= i1; ++i) {
      const int ky = Y - i * g.S;
      v = (ky < g.step) ? blend_u8(v, hv, wtab[ky]) : hv;
    }
    out[static_cast<long long>(Y) * g.E + X] = v;
  }
}

// value of the stitched attention map at (Y, X): sequential blends of the bilinearly upsampled
// per-tile maps, fp64 products/sum rounded to fp32 at every seam (the crops are float32 arrays).
__device__ __forceinline__ float blend_f32(float a, float b, double w) {
  return static_cast<float>(__dadd_rn(__dmul_rn(static_cast<double>(a), w), __dmul_rn(static_cast<double>(b), __dsub_rn(1.0, w))));
}
__device__ __forceinline__ float stitched_value(const float* __restrict__ lowres, const StitchGeom& g,
                                                const double* __restrict__ wtab, int Y, int X) {
  int j0 = (X - g.W + g.S) / g.S; if (X - g.W + 1 <= 0) j0 = 0; if (j0 < 0) j0 = 0;
  const int j1 = min(X / g.S, g.n - 1);
  int i0 = (Y - g.W + g.S) / g.S; if (Y - g.W + 1 <= 0) i0 = 0; if (i0 < 0) i0 = 0;
  const int i1 = min(Y / g.S, g.n - 1);
  const int lsz = g.lh * g.lw;
  float v = 0.f;
  for (int i = i0; i <= i1; ++i) {
    const int ky = Y - i * g.S;
    float hv = 0.f;
    for (int j = j0; j <= j1; ++j) {
      const int kx = X - j * g.S;
      const float tv = bilinear_up(lowres + static_cast<long long>(i * g.n + j) * lsz, g.lh, g.lw, ky, kx, g.scale);
      hv = (j > j0 && kx < g.step) ? blend_f32(hv, tv, wtab[kx]) : tv;
    }
    v = (i > i0 && ky < g.step) ? blend_f32(v, hv, wtab[ky]) : hv;
  }
  return v;
}

// concat_crops (SSS/sw_processing.py:113-134) on full-resolution float32 crops [n*n][W][W] -> out [E][E]
__global__ void concat_crops_f32_kernel(const float* __restrict__ crops, StitchGeom g, const double* __restrict__ wtab,
                                        float* __restrict__ out) {
  const long long total = static_cast<long long>(g.E) * g.E;
  const long long tsz = static_cast<long long>(g.W) * g.W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int Y = static_cast<int>(idx / g.E), X = static_cast<int>(idx % g.E);
    int j0 = (X - g.W + g.S) / g.S; if (X - g.W + 1 <= 0) j0 = 0;
    const int j1 = min(X / g.S, g.n - 1);
    int i0 = (Y - g.W + g.S) / g.S; if (Y - g.W + 1 <= 0) i0 = 0;
    const int i1 = min(Y / g.S, g.n - 1);
    float v = 0.f;
    for (int i = i0; i <= i1; ++i) {
      const int ky = Y - i * g.S;
      float hv = 0.f;
      for (int j = j0; j <= j1; ++j) {
        const int kx = X - j * g.S;
        const float tv = crops[(i * g.n + j) * tsz + static_cast<long long>(ky) * g.W + kx];
        hv = (j > j0 && kx < g.step) ? blend_f32(hv, tv, wtab[kx]) : tv;
      }
      v = (i > i0 && ky < g.step) ? blend_f32(v, hv, wtab[ky]) : hv;
    }
    out[idx] = v;
  }
}
// same on uint8 crops [n*n][W][W][C] (HWC, as PIL / numpy image crops) -> out [E][E][C]
__global__ void concat_crops_u8_kernel(const uint8_t* __restrict__ crops, StitchGeom g, int C, const double* __restrict__ wtab,
                                       uint8_t* __restrict__ out) {
  const long long total = static_cast<long long>(g.E) * g.E * C;
  const long long tsz = static_cast<long long>(g.W) * g.W * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const long long px = idx / C;
    const int Y = static_cast<int>(px / g.E), X = static_cast<int>(px % g.E);
    int j0 = (X - g.W + g.S) / g.S; if (X - g.W + 1 <= 0) j0 = 0;
    const int j1 = min(X / g.S, g.n - 1);
    int i0 = (Y - g.W + g.S) / g.S; if (Y - g.W + 1 <= 0) i0 = 0;
    const int i1 = min(Y / g.S, g.n - 1);
    uint8_t v = 0;
    for (int i = i0; i <= i1; ++i) {
      const int ky = Y - i * g.S;
      uint8_t hv = 0;
      for (int j = j0; j <= j1; ++j) {
        const int kx = X - j * g.S;
        const uint8_t tv = crops[(i * g.n + j) * tsz + (static_cast<long long>(ky) * g.W + kx) * C + c];
        hv = (j > j0 && kx < g.step) ? blend_u8(hv, tv, wtab[kx]) : tv;
      }
      v = (i > i0 && ky < g.step) ? blend_u8(v, hv, wtab[ky]) : hv;
    }
    out[idx] = v;
  }
}
// sliding_window on a uint8 HWC image -> crops [ny*nx][W][W][C], zero padded outside the image.
// Window (iy, ix) starts at (oy[iy], ox[ix]) = (iy*S, ix*S).
__global__ void crop_u8_kernel(const uint8_t* __restrict__ img, int img_h, int img_w, int C, int ny, int nx, int W, int S,
                               uint8_t* __restrict__ crops) {
  const long long tsz = static_cast<long long>(W) * W * C;
  const long long total = tsz * ny * nx;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(idx / tsz);
    const long long r = idx - t * tsz;
    const int c = static_cast<int>(r % C);
    const int xx = static_cast<int>((r / C) % W), yy = static_cast<int>(r / (static_cast<long long>(C) * W));
    const int gy = (t / nx) * S + yy, gx = (t % nx) * S + xx;
    crops[idx] = (gy < img_h && gx < img_w) ? img[(static_cast<long long>(gy) * img_w + gx) * C + c] : 0;
  }
}

// ---------------------------------------------------------------------------------------
// Crop variants of the eval / analyse scripts.
// ---------------------------------------------------------------------------------------
// concat_crops_overlap(crops, stride) (SSS/utils.py:319-347): n x n crops of size W overlapping by
// V = 2 * stride; inside an overlap the running image and the next crop are each floor-halved and added
// (`a // 2 + b // 2`), first along x inside a strip, then along y across strips -- except that the LAST
// strip is appended without blending: its overlap rows keep the running image (:337-339).
// Gather form: the value of an output pixel is the fold, in crop order, over the crops that cover it.
struct OverlapGeom {
  int n, W, V, step, E;   // step = W - V, E = W + (n-1) * step
};
__device__ __forceinline__ float half_floor(float a) { return floorf(__fmul_rn(a, 0.5f)); }
__device__ __forceinline__ float avg_halves(float a, float b) { return __fadd_rn(half_floor(a), half_floor(b)); }
__device__ __forceinline__ uint8_t avg_halves(uint8_t a, uint8_t b) { return static_cast<uint8_t>((a >> 1) + (b >> 1)); }

template <typename T>
__global__ void concat_crops_overlap_kernel(const T* __restrict__ crops /*[n*n][W][W][C]*/, OverlapGeom g, int C,
                                            T* __restrict__ out /*[E][E][C]*/) {
  const long long total = static_cast<long long>(g.E) * g.E * C;
  const long long tsz = static_cast<long long>(g.W) * g.W * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const long long px = idx / C;
    const int Y = static_cast<int>(px / g.E), X = static_cast<int>(px % g.E);
    // crops covering X: j*step <= X < j*step + W
    int j0 = X < g.W ? 0 : (X - g.W) / g.step + 1;
    const int j1 = min(X / g.step, g.n - 1);
    int i0 = Y < g.W ? 0 : (Y - g.W) / g.step + 1;
    const int i1 = min(Y / g.step, g.n - 1);
    T v = T(0);
    for (int i = i0; i <= i1; ++i) {
      const int ky = Y - i * g.step;
      T hv = T(0);
      for (int j = j0; j <= j1; ++j) {
        const int kx = X - j * g.step;
        const T tv = crops[(i * g.n + j) * tsz + (static_cast<long long>(ky) * g.W + kx) * C + c];
        hv = (j > j0 && kx < g.V) ? avg_halves(hv, tv) : tv;
      }
      if (i > i0 && ky < g.V) {
        if (i != g.n - 1) v = avg_halves(v, hv);   // last strip: the running image is kept as it is
      } else {
        v = hv;
      }
    }
    out[idx] = v;
  }
}

// plain n x n tiling (SSS/utils.py:304-317 `concat_crops`; SSS/eval.py:160-161) of one channel of batched crops:
// src [B][cr*cr][C][h][w] fp32 -> dst [B][cr*h][cr*w], channel c0.  Pure data movement.
__global__ void concat_grid_f32_kernel(const float* __restrict__ src, int B, int cr, int C, int c0, int h, int w,
                                       float* __restrict__ dst) {
  const int EH = cr * h, EW = cr * w;
  const long long total = static_cast<long long>(B) * EH * EW;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % EW);
    const int Y = static_cast<int>((idx / EW) % EH);
    const int b = static_cast<int>(idx / (static_cast<long long>(EW) * EH));
    const int ci = Y / h, cj = X / w;
    dst[idx] = src[(((static_cast<long long>(b) * cr * cr + ci * cr + cj) * C + c0) * h + (Y - ci * h)) * w + (X - cj * w)];
  }
}

// pass 1: global min / max of the stitched map over rows [y_begin, y_end); minmax_ord[0] = min, [1] = max
// (order-preserving int keys; initialise to INT_MAX / INT_MIN); optionally store the map.
__global__ void __launch_bounds__(256)
stitch_minmax_kernel(const float* __restrict__ lowres, StitchGeom g, const double* __restrict__ wtab, int y_begin, int y_end,
                     int* __restrict__ minmax_ord, float* __restrict__ map_out /*[E][E] or null*/,
                     const float* __restrict__ map_in /*[E][E] or null: use instead of stitching lowres*/) {
  __shared__ float red_mn[8], red_mx[8];
  const long long total = static_cast<long long>(y_end - y_begin) * g.E;
  float mn = INFINITY, mx = -INFINITY;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int Y = y_begin + static_cast<int>(idx / g.E), X = static_cast<int>(idx % g.E);
    const float v = map_in != nullptr ? map_in[static_cast<long long>(Y) * g.E + X] : stitched_value(lowres, g, wtab, Y, X);
    if (map_out != nullptr) map_out[static_cast<long long>(Y) * g.E + X] = v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { red_mn[threadIdx.x >> 5] = mn; red_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) { mn = fminf(mn, red_mn[w]); mx = fmaxf(mx, red_mx[w]); }
    if (mn <= mx) {
      atomicMin(&minmax_ord[0], f2ord(mn));
      atomicMax(&minmax_ord[1], f2ord(mx));
    }
  }
}

// per-pixel classification of the mosaic flavour (SSS/sw_processing.py:43-48)
__device__ __forceinline__ void sw_classify(float v, float mn, float range, bool flat, float att_max, int img, int& res, int& au) {
  const float an = flat ? v : __fdiv_rn(__fsub_rn(v, mn), range);
  res = static_cast<int>(static_cast<uint8_t>(static_cast<int>(__fdiv_rn(__fmul_rn(static_cast<float>(img), an), att_max))));
  au = static_cast<int>(static_cast<uint8_t>(static_cast<int>(__fmul_rn(an, 255.0f))));
}

// pass 2: histograms of result (= img * att), of the stitched gray image and of att_u8 -> hists[3][256]
__global__ void __launch_bounds__(256)
stitch_hist_kernel(const float* __restrict__ lowres, StitchGeom g, const double* __restrict__ wtab,
                   const uint8_t* __restrict__ gray /*[E][E]*/, const int* __restrict__ minmax_ord, int y_begin, int y_end,
                   unsigned long long* __restrict__ hists, const float* __restrict__ map_in) {
  __shared__ unsigned int h[3][256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) (&h[0][0])[i] = 0u;
  __syncthreads();
  const float mn = ord2f(minmax_ord[0]), mx = ord2f(minmax_ord[1]);
  const bool flat = (mx == mn);
  const float range = __fsub_rn(mx, mn);
  const float att_max = flat ? mx : 1.0f;  // np.max(attention) after min_max_normalize
  const long long total = static_cast<long long>(y_end - y_begin) * g.E;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int Y = y_begin + static_cast<int>(idx / g.E), X = static_cast<int>(idx % g.E);
    const float v = map_in != nullptr ? map_in[static_cast<long long>(Y) * g.E + X] : stitched_value(lowres, g, wtab, Y, X);
    const int img = gray[static_cast<long long>(Y) * g.E + X];
    int res, au;
    sw_classify(v, mn, range, flat, att_max, img, res, au);
    atomicAdd(&h[0][res], 1u);
    atomicAdd(&h[1][img], 1u);
    atomicAdd(&h[2][au], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const unsigned int c = (&h[0][0])[i];
    if (c) atomicAdd(&hists[i], static_cast<unsigned long long>(c));
  }
}

// pass 3: masks th (result > t0), th2 (gray > t1), th3 (att_u8 > t2); rows [y_begin, y_end) written at
// out + (Y - y_begin) * E so that a rank can hold only its own band.
__global__ void __launch_bounds__(256)
stitch_mask_kernel(const float* __restrict__ lowres, StitchGeom g, const double* __restrict__ wtab,
                   const uint8_t* __restrict__ gray, const int* __restrict__ minmax_ord, const int* __restrict__ thr,
                   int y_begin, int y_end, uint8_t* __restrict__ th, uint8_t* __restrict__ th2, uint8_t* __restrict__ th3,
                   const float* __restrict__ map_in) {
  const float mn = ord2f(minmax_ord[0]), mx = ord2f(minmax_ord[1]);
  const bool flat = (mx == mn);
  const float range = __fsub_rn(mx, mn);
  const float att_max = flat ? mx : 1.0f;
  const int t0 = thr[0], t1 = thr[1], t2 = thr[2];
  const long long total = static_cast<long long>(y_end - y_begin) * g.E;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int Y = y_begin + static_cast<int>(idx / g.E), X = static_cast<int>(idx % g.E);
    const float v = map_in != nullptr ? map_in[static_cast<long long>(Y) * g.E + X] : stitched_value(lowres, g, wtab, Y, X);
    const int img = gray[static_cast<long long>(Y) * g.E + X];
    int res, au;
    sw_classify(v, mn, range, flat, att_max, img, res, au);
    if (th != nullptr) th[idx] = res > t0 ? 255 : 0;
    if (th2 != nullptr) th2[idx] = img > t1 ? 255 : 0;
    if (th3 != nullptr) th3[idx] = au > t2 ? 255 : 0;
  }
}

// the weighted image `result = (img * att / max(att)).astype(u8)` of the mosaic flavour (SSS/sw_processing.py:44-46, the
// "weighted_iamge_attention.png" it saves at :75) and att_u8, for rows [y_begin, y_end)
__global__ void __launch_bounds__(256)
stitch_result_kernel(const float* __restrict__ lowres, StitchGeom g, const double* __restrict__ wtab, const uint8_t* __restrict__ gray,
                     const int* __restrict__ minmax_ord, int y_begin, int y_end, uint8_t* __restrict__ result, uint8_t* __restrict__ att_u8,
                     const float* __restrict__ map_in) {
  const float mn = ord2f(minmax_ord[0]), mx = ord2f(minmax_ord[1]);
  const bool flat = (mx == mn);
  const float range = __fsub_rn(mx, mn);
  const float att_max = flat ? mx : 1.0f;
  const long long total = static_cast<long long>(y_end - y_begin) * g.E;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int Y = y_begin + static_cast<int>(idx / g.E), X = static_cast<int>(idx % g.E);
    const float v = map_in != nullptr ? map_in[static_cast<long long>(Y) * g.E + X] : stitched_value(lowres, g, wtab, Y, X);
    int res, au;
    sw_classify(v, mn, range, flat, att_max, gray[static_cast<long long>(Y) * g.E + X], res, au);
    if (result != nullptr) result[idx] = static_cast<uint8_t>(res);
    if (att_u8 != nullptr) att_u8[idx] = static_cast<uint8_t>(au);
  }
}

}  // namespace vitocm
