// fp32 FFMA implicit-GEMM convolution (forward / data-gradient / weight-gradient), sm_100a.
//
// This is the EXACT-fp32 path of the library (north-star: "fp32 voxel probabilities within
// 1e-3"): operands are read as stored (fp32, or bf16 widened to fp32), every product and sum
// is an fp32 FMA.  It is also the in-GPU cross-check for the tcgen05 kernels in conv_umma.cu.
// Replaces tensor.nnet.conv2d (lib/layers.py:322), conv3d2d.conv3d (lib/layers.py:441,504) and
// tensor.dot (lib/layers.py:159,503) of the reference, plus their tensor.grad (models/net.py:58).
#include "common.cuh"

namespace r2n2 {

struct ConvGeom {
  int N, D, H, W, Cin, Cout, kd, kh, kw;
  int taps, K;     // K = taps * Cin (GEMM reduction length)
  long long P;     // N*D*H*W output pixels (GEMM M)
};

constexpr int kMaxTaps = 64;

// ------------------------------------------------------------------------------------ fwd
// Tile 128 pixels x 64 out-channels x 16 k ; 256 threads, 8x4 register tile each.
// kVec: Cin % 4 == 0 -> 4-wide vector loads along the channel axis; otherwise scalar loads
// (only the data gradient of conv11, whose "input" is the 2-channel dlogits, needs that).
template <typename TIn, typename TOut, bool kVec>
__global__ void __launch_bounds__(256)
conv_fwd_simt_kernel(const TIn* __restrict__ x, const TIn* __restrict__ w,
                     const float* __restrict__ bias, const TOut* mask,
                     const TOut* res, TOut* y, ConvGeom g, int flags) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  __shared__ __align__(16) float As[2][16][128];
  __shared__ __align__(16) float Bs[2][16][64];
  __shared__ int tap_dz[kMaxTaps], tap_dy[kMaxTaps], tap_dx[kMaxTaps];

  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * 128;
  const int n0 = blockIdx.y * 64;

  if (tid < g.taps) {
    int t = tid;
    int tz = t / (g.kh * g.kw);
    int r = t - tz * g.kh * g.kw;
    int ty = r / g.kw;
    int tx = r - ty * g.kw;
    tap_dz[t] = tz - (g.kd - 1) / 2;
    tap_dy[t] = ty - (g.kh - 1) / 2;
    tap_dx[t] = tx - (g.kw - 1) / 2;
  }
  __syncthreads();

  // ---- A-load role: one pixel row per thread, 8 consecutive k
  const int a_row = tid & 127;
  const int a_k8 = (tid >> 7) * 8;
  const long long a_p = m0 + a_row;
  const bool a_valid = a_p < g.P;
  int a_n = 0, a_z = 0, a_y = 0, a_x = 0;
  if (a_valid) {
    long long t = a_p;
    a_x = (int)(t % g.W); t /= g.W;
    a_y = (int)(t % g.H); t /= g.H;
    a_z = (int)(t % g.D);
    a_n = (int)(t / g.D);
  }
  // ---- B-load role: one out-channel per thread, 4 consecutive k
  const int b_co = tid & 63;
  const int b_k4 = (tid >> 6) * 4;
  const bool b_valid = (n0 + b_co) < g.Cout;
  const TIn* b_ptr = w + (long long)(n0 + b_co) * g.K;

  float4 ra[2], rb;
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int kk = k0 + a_k8 + 4 * j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!kVec) {
        float t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int ke = kk + e;
          if (a_valid && ke < g.K) {
            int tap = ke / g.Cin;
            int ci = ke - tap * g.Cin;
            int sz = a_z + tap_dz[tap], sy = a_y + tap_dy[tap], sx = a_x + tap_dx[tap];
            if ((unsigned)sz < (unsigned)g.D && (unsigned)sy < (unsigned)g.H &&
                (unsigned)sx < (unsigned)g.W)
              t4[e] = to_float<TIn>(
                  x[((((long long)a_n * g.D + sz) * g.H + sy) * g.W + sx) * g.Cin + ci]);
          }
        }
        v = make_float4(t4[0], t4[1], t4[2], t4[3]);
      } else if (a_valid && kk < g.K) {
        int tap = kk / g.Cin;
        int ci = kk - tap * g.Cin;
        int sz = a_z + tap_dz[tap], sy = a_y + tap_dy[tap], sx = a_x + tap_dx[tap];
        if ((unsigned)sz < (unsigned)g.D && (unsigned)sy < (unsigned)g.H &&
            (unsigned)sx < (unsigned)g.W) {
          long long off = ((((long long)a_n * g.D + sz) * g.H + sy) * g.W + sx) * g.Cin + ci;
          v = Vec4<TIn>::load(x + off);
        }
      }
      ra[j] = v;
    }
    {
      int kk = k0 + b_k4;
      if (kVec) {
        rb = (b_valid && kk < g.K) ? Vec4<TIn>::load(b_ptr + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        float t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (b_valid && kk + e < g.K) t4[e] = to_float<TIn>(b_ptr[kk + e]);
        rb = make_float4(t4[0], t4[1], t4[2], t4[3]);
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      As[buf][a_k8 + 4 * j + 0][a_row] = ra[j].x;
      As[buf][a_k8 + 4 * j + 1][a_row] = ra[j].y;
      As[buf][a_k8 + 4 * j + 2][a_row] = ra[j].z;
      As[buf][a_k8 + 4 * j + 3][a_row] = ra[j].w;
    }
    Bs[buf][b_k4 + 0][b_co] = rb.x;
    Bs[buf][b_k4 + 1][b_co] = rb.y;
    Bs[buf][b_k4 + 2][b_co] = rb.z;
    Bs[buf][b_k4 + 3][b_co] = rb.w;
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (g.K + 15) / 16;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: [+residual (RES_PRE)] -> bias -> lrelu -> [+residual] -> *mask
  const bool res_pre = (flags & R2N2_EPI_RESIDUAL) && (flags & R2N2_EPI_RES_PRE);
  const bool res_post = (flags & R2N2_EPI_RESIDUAL) && !(flags & R2N2_EPI_RES_PRE);
  const int co0 = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (flags & R2N2_EPI_BIAS) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co0 + j < g.Cout) bv[j] = bias[co0 + j];
  }
  const bool vec_ok = ((g.Cout & 3) == 0) && (co0 + 3 < g.Cout);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long p = m0 + ty * 8 + i;
    if (p >= g.P) continue;
    long long off = p * g.Cout + co0;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j];
      if (res_pre && co0 + j < g.Cout) v[j] += to_float<TOut>(res[off + j]);
      v[j] += bv[j];
      if (flags & R2N2_EPI_LRELU) v[j] = lrelu(v[j]);
    }
    if (vec_ok) {
      if (res_post) {
        float4 r = Vec4<TOut>::load(res + off);
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
      if (flags & R2N2_EPI_MASK) {
        float4 m = Vec4<TOut>::load(mask + off);
        v[0] *= lrelu_slope(m.x); v[1] *= lrelu_slope(m.y);
        v[2] *= lrelu_slope(m.z); v[3] *= lrelu_slope(m.w);
      }
      Vec4<TOut>::store(y + off, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (co0 + j >= g.Cout) continue;
        float o = v[j];
        if (res_post) o += to_float<TOut>(res[off + j]);
        if (flags & R2N2_EPI_MASK) o *= lrelu_slope(to_float<TOut>(mask[off + j]));
        y[off + j] = from_float<TOut>(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ wgrad
// dw[co][kcol] += sum_p dy[p][co] * x[p + off(tap)][ci],  kcol = tap*Cin + ci.
// Tile 64 co x 64 kcol x 16 pixels; grid.z splits the pixel range (fp32 atomics combine).
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                       ConvGeom g, long long p_per_split, int use_atomics) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  __shared__ __align__(16) float As[2][16][64];
  __shared__ __align__(16) float Bs[2][16][64];
  const int tid = threadIdx.x;
  const int kc0 = blockIdx.x * 64;
  const int co0 = blockIdx.y * 64;
  const long long p_begin = (long long)blockIdx.z * p_per_split;
  long long p_end = p_begin + p_per_split;
  if (p_end > g.P) p_end = g.P;

  const int pr = tid >> 4;
  const int v4 = (tid & 15) * 4;
  // A role (dy)
  const int a_co = co0 + v4;
  const bool a_vec = ((g.Cout & 3) == 0) && (a_co + 3 < g.Cout);
  // B role (shifted x)
  const int b_kcol = kc0 + v4;
  const bool b_valid = b_kcol < g.K;
  int b_ci = 0, dz = 0, dyy = 0, dxx = 0;
  if (b_valid) {
    int tap = b_kcol / g.Cin;
    b_ci = b_kcol - tap * g.Cin;
    int tz = tap / (g.kh * g.kw);
    int r = tap - tz * g.kh * g.kw;
    int ty = r / g.kw;
    int tx = r - ty * g.kw;
    dz = tz - (g.kd - 1) / 2;
    dyy = ty - (g.kh - 1) / 2;
    dxx = tx - (g.kw - 1) / 2;
  }

  float4 ra, rb;
  auto load_tiles = [&](long long p0) {
    long long p = p0 + pr;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (p < p_end) {
      if (a_vec) {
        ra = Vec4<T>::load(dy + p * g.Cout + a_co);
      } else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j)
          if (a_co + j < g.Cout) t[j] = to_float<T>(dy[p * g.Cout + a_co + j]);
        ra = make_float4(t[0], t[1], t[2], t[3]);
      }
      if (b_valid) {
        long long t = p;
        int xx = (int)(t % g.W); t /= g.W;
        int yy = (int)(t % g.H); t /= g.H;
        int zz = (int)(t % g.D);
        int n = (int)(t / g.D);
        int sz = zz + dz, sy = yy + dyy, sx = xx + dxx;
        if ((unsigned)sz < (unsigned)g.D && (unsigned)sy < (unsigned)g.H &&
            (unsigned)sx < (unsigned)g.W) {
          long long off = ((((long long)n * g.D + sz) * g.H + sy) * g.W + sx) * g.Cin + b_ci;
          rb = Vec4<T>::load(x + off);
        }
      }
    }
  };
  auto store_tiles = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][pr][v4]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][pr][v4]) = rb;
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const long long np = p_end - p_begin;
  const int nk = (int)((np + 15) / 16);
  if (nk > 0) {
    load_tiles(p_begin);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
      int buf = kt & 1;
      if (kt + 1 < nk) load_tiles(p_begin + (long long)(kt + 1) * 16);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w};
        float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (kt + 1 < nk) {
        store_tiles(buf ^ 1);
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kc = kc0 + tx * 4 + j;
      if (kc >= g.K) continue;
      float* dst = dw + (long long)co * g.K + kc;
      if (use_atomics) atomicAdd(dst, acc[i][j]);
      else *dst = acc[i][j];
    }
  }
}

static int make_geom(ConvGeom& g, int N, int D, int H, int W, int Cin, int Cout, int kd, int kh,
                     int kw, const char* who, bool need_vec = true) {
  R2N2_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, R2N2_ERR_BAD_SHAPE,
                 "%s: non-positive dimension", who);
  R2N2_CHECK_ARG(kd > 0 && kh > 0 && kw > 0 && (kd & 1) && (kh & 1) && (kw & 1),
                 R2N2_ERR_BAD_SHAPE, "%s: kernel extents must be odd ('same' convolution)", who);
  R2N2_CHECK_ARG(kd * kh * kw <= kMaxTaps, R2N2_ERR_BAD_SHAPE, "%s: more than %d taps", who,
                 kMaxTaps);
  R2N2_CHECK_ARG(!need_vec || (Cin & 3) == 0, R2N2_ERR_BAD_ALIGN,
                 "%s: Cin (%d) must be a multiple of 4 (pad the channel axis)", who, Cin);
  g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
  g.kd = kd; g.kh = kh; g.kw = kw;
  g.taps = kd * kh * kw;
  g.K = g.taps * Cin;
  g.P = (long long)N * D * H * W;
  return R2N2_OK;
}

int conv_fwd_simt(const void* x, const void* w, const float* bias, const void* mask,
                  const void* res, void* y, int N, int D, int H, int W, int Cin, int Cout, int kd,
                  int kh, int kw, int flags, int in_dtype, int out_dtype, cudaStream_t st) {
  ConvGeom g;
  int rc = make_geom(g, N, D, H, W, Cin, Cout, kd, kh, kw, "conv_fwd[simt]", false);
  if (rc) return rc;
  const bool vec = (Cin & 3) == 0;
  dim3 grid((unsigned)((g.P + 127) / 128), (Cout + 63) / 64);
  R2N2_CHECK_ARG(grid.y <= 65535, R2N2_ERR_BAD_SHAPE, "conv_fwd[simt]: Cout too large");
#define R2N2_LAUNCH_FWD(TI, TO)                                                              \
  do {                                                                                       \
    if (vec)                                                                                 \
      r2n2::launch_k((conv_fwd_simt_kernel<TI, TO, true>), dim3(grid), dim3(256), 0, (cudaStream_t)(st), 0,                              \
          (const TI*)x, (const TI*)w, bias, (const TO*)mask, (const TO*)res, (TO*)y, g, flags); \
    else                                                                                     \
      r2n2::launch_k((conv_fwd_simt_kernel<TI, TO, false>), dim3(grid), dim3(256), 0, (cudaStream_t)(st), 0,                             \
          (const TI*)x, (const TI*)w, bias, (const TO*)mask, (const TO*)res, (TO*)y, g, flags); \
  } while (0)
  if (in_dtype == R2N2_F32 && out_dtype == R2N2_F32) R2N2_LAUNCH_FWD(float, float);
  else if (in_dtype == R2N2_BF16 && out_dtype == R2N2_BF16) R2N2_LAUNCH_FWD(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == R2N2_BF16 && out_dtype == R2N2_F32) R2N2_LAUNCH_FWD(__nv_bfloat16, float);
  else if (in_dtype == R2N2_F32 && out_dtype == R2N2_BF16) R2N2_LAUNCH_FWD(float, __nv_bfloat16);
  else R2N2_CHECK_ARG(false, R2N2_ERR_BAD_DTYPE, "conv_fwd[simt]: bad dtype codes %d/%d", in_dtype, out_dtype);
#undef R2N2_LAUNCH_FWD
  R2N2_CHECK_LAUNCH("conv_fwd[simt]");
  return R2N2_OK;
}

int conv_wgrad_simt(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin,
                    int Cout, int kd, int kh, int kw, int dtype, int accumulate, cudaStream_t st) {
  ConvGeom g;
  int rc = make_geom(g, N, D, H, W, Cin, Cout, kd, kh, kw, "conv_wgrad[simt]");
  if (rc) return rc;
  int gx = (g.K + 63) / 64, gy = (Cout + 63) / 64;
  long long tiles = (long long)gx * gy;
  long long want = ((long long)num_sms() * 4 + tiles - 1) / tiles;
  long long max_splits = (g.P + 255) / 256;
  long long splits = want < max_splits ? want : max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long pps = (g.P + splits - 1) / splits;
  pps = (pps + 15) / 16 * 16;
  splits = (g.P + pps - 1) / pps;
  int use_atomics = (splits > 1 || accumulate) ? 1 : 0;
  if (!accumulate && use_atomics)
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * g.K, st);
  dim3 grid(gx, gy, (unsigned)splits);
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((conv_wgrad_simt_kernel<T>), dim3(grid), dim3(256), 0, (cudaStream_t)(st), 0, (const T*)x, (const T*)dy, dw, g, pps,
                                                   use_atomics);
  })
  R2N2_CHECK_LAUNCH("conv_wgrad[simt]");
  return R2N2_OK;
}

// implemented in conv_umma.cu
int conv_fwd_umma(const void* x, const void* w, const float* bias, const void* mask,
                  const void* res, void* y, int N, int D, int H, int W, int Cin, int Cout, int kd,
                  int kh, int kw, int flags, int in_dtype, int out_dtype, float* colsum, cudaStream_t st,
                  const GruEpi* gru = nullptr);
int conv_wgrad_umma(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin,
                    int Cout, int kd, int kh, int kw, int dtype, int accumulate, cudaStream_t st);

}  // namespace r2n2

extern "C" {

static int conv_fwd_dispatch(const void* x, const void* w, const float* bias, const void* mask_src,
                             const void* residual, void* y, int N, int D, int H, int W, int Cin, int Cout,
                             int kd, int kh, int kw, int flags, int in_dtype, int out_dtype, int algo,
                             float* colsum, void* stream) {
  using namespace r2n2;
  R2N2_CHECK_ARG(x && w && y, R2N2_ERR_BAD_SHAPE, "conv_fwd: null pointer");
  R2N2_CHECK_ARG(!(flags & R2N2_EPI_BIAS) || bias, R2N2_ERR_BAD_SHAPE, "conv_fwd: BIAS flag without bias");
  R2N2_CHECK_ARG(!(flags & R2N2_EPI_RESIDUAL) || residual, R2N2_ERR_BAD_SHAPE, "conv_fwd: RESIDUAL flag without residual");
  R2N2_CHECK_ARG(!(flags & R2N2_EPI_MASK) || mask_src, R2N2_ERR_BAD_SHAPE, "conv_fwd: MASK flag without mask_src");
  R2N2_CHECK_ARG(algo == R2N2_ALGO_UMMA || !(flags & (R2N2_CONV_UP2 | R2N2_CONV_DOWN2)), R2N2_ERR_UNSUPPORTED,
                 "conv_fwd: UP2 / DOWN2 need R2N2_ALGO_UMMA (materialise the unpooled tensor for ALGO_SIMT)");
  if (algo == R2N2_ALGO_SIMT) {
    int rc = conv_fwd_simt(x, w, bias, mask_src, residual, y, N, D, H, W, Cin, Cout, kd, kh, kw, flags,
                           in_dtype, out_dtype, (cudaStream_t)stream);
    if (rc == R2N2_OK && colsum)
      rc = r2n2_bias_grad(y, colsum, (long long)N * D * H * W, Cout, out_dtype, 1, stream);
    return rc;
  }
  if (algo == R2N2_ALGO_UMMA)
    return conv_fwd_umma(x, w, bias, mask_src, residual, y, N, D, H, W, Cin, Cout, kd, kh, kw, flags,
                         in_dtype, out_dtype, colsum, (cudaStream_t)stream);
  R2N2_CHECK_ARG(false, R2N2_ERR_UNSUPPORTED, "conv_fwd: unknown algo %d", algo);
}

int r2n2_conv_fwd(const void* x, const void* w, const float* bias, const void* mask_src,
                  const void* residual, void* y, int N, int D, int H, int W, int Cin, int Cout,
                  int kd, int kh, int kw, int flags, int in_dtype, int out_dtype, int algo,
                  void* stream) {
  return conv_fwd_dispatch(x, w, bias, mask_src, residual, y, N, D, H, W, Cin, Cout, kd, kh, kw, flags, in_dtype,
                           out_dtype, algo, nullptr, stream);
}

int r2n2_conv_fwd_colsum(const void* x, const void* w, const float* bias, const void* mask_src,
                         const void* residual, void* y, int N, int D, int H, int W, int Cin, int Cout,
                         int kd, int kh, int kw, int flags, int in_dtype, int out_dtype, int algo,
                         float* colsum, void* stream) {
  R2N2_CHECK_ARG(colsum != nullptr, R2N2_ERR_BAD_SHAPE, "conv_fwd_colsum: null colsum");
  return conv_fwd_dispatch(x, w, bias, mask_src, residual, y, N, D, H, W, Cin, Cout, kd, kh, kw, flags, in_dtype,
                           out_dtype, algo, colsum, stream);
}

int r2n2_conv_fwd_gru(int mode, const void* x, const void* w, const float* bias, const void* in0, const void* in1,
                      const void* in2, const void* in3, void* out0, void* out1, void* out2, int N, int D, int H,
                      int W, int Cin, int Cout, int kd, int kh, int kw, int dtype, void* stream) {
  using namespace r2n2;
  R2N2_CHECK_ARG(mode >= R2N2_GRU_UR && mode <= R2N2_GRU_OBWD, R2N2_ERR_UNSUPPORTED, "conv_fwd_gru: unknown mode %d", mode);
  R2N2_CHECK_ARG(dtype == R2N2_BF16, R2N2_ERR_BAD_DTYPE, "conv_fwd_gru: bf16 tensors only (the fp32 path keeps the gate kernels)");
  R2N2_CHECK_ARG(x && w && in0 && in1 && out0 && out1, R2N2_ERR_BAD_SHAPE, "conv_fwd_gru: null pointer");
  const bool fwd = mode == R2N2_GRU_UR || mode == R2N2_GRU_BLEND;
  R2N2_CHECK_ARG(!fwd || bias, R2N2_ERR_BAD_SHAPE, "conv_fwd_gru: forward modes need the gate bias");
  R2N2_CHECK_ARG(mode == R2N2_GRU_RBWD || (mode == R2N2_GRU_BLEND ? in2 != nullptr : out2 != nullptr), R2N2_ERR_BAD_SHAPE,
                 "conv_fwd_gru: null pointer");
  R2N2_CHECK_ARG(mode != R2N2_GRU_OBWD || in2, R2N2_ERR_BAD_SHAPE, "conv_fwd_gru: null pointer");
  R2N2_CHECK_ARG(Cout % 32 == 0 && Cin % 16 == 0, R2N2_ERR_BAD_SHAPE, "conv_fwd_gru: Cout %d must be a multiple of 32", Cout);
  const void* ptrs[7] = {in0, in1, in2, in3, out0, out1, out2};
  for (int i = 0; i < 7; ++i)
    R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(ptrs[i]) & 31) == 0, R2N2_ERR_BAD_ALIGN,
                   "conv_fwd_gru: gate tensors must be 32-byte aligned");
  GruEpi g;
  g.mode = mode;
  g.in[0] = in0; g.in[1] = in1; g.in[2] = in2; g.in[3] = in3;
  g.out[0] = out0; g.out[1] = out1; g.out[2] = out2;
  return conv_fwd_umma(x, w, bias, nullptr, nullptr, out0, N, D, H, W, Cin, Cout, kd, kh, kw, fwd ? R2N2_EPI_BIAS : 0,
                       R2N2_BF16, R2N2_BF16, nullptr, (cudaStream_t)stream, &g);
}

int r2n2_conv_wgrad(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin,
                    int Cout, int kd, int kh, int kw, int dtype, int algo, int accumulate,
                    void* stream) {
  using namespace r2n2;
  R2N2_CHECK_ARG(x && dy && dw, R2N2_ERR_BAD_SHAPE, "conv_wgrad: null pointer");
  R2N2_CHECK_ARG(algo == R2N2_ALGO_UMMA || !(accumulate & R2N2_WGRAD_UP2), R2N2_ERR_UNSUPPORTED,
                 "conv_wgrad: UP2 needs R2N2_ALGO_UMMA");
  if (algo == R2N2_ALGO_SIMT)
    return conv_wgrad_simt(x, dy, dw, N, D, H, W, Cin, Cout, kd, kh, kw, dtype, accumulate & R2N2_WGRAD_ACCUMULATE,
                           (cudaStream_t)stream);
  if (algo == R2N2_ALGO_UMMA)
    return conv_wgrad_umma(x, dy, dw, N, D, H, W, Cin, Cout, kd, kh, kw, dtype, accumulate,
                           (cudaStream_t)stream);
  R2N2_CHECK_ARG(false, R2N2_ERR_UNSUPPORTED, "conv_wgrad: unknown algo %d", algo);
}

}  // extern "C"
