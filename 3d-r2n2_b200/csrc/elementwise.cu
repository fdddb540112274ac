// Bandwidth-bound kernels of the 3D-R2N2 hot path (sm_100a): layout transforms, 2x2 max-pool,
// Unpool3D, LeakyReLU/residual helpers, conv-GRU / conv-LSTM gate math, fused voxel
// softmax-cross-entropy (fwd+bwd), bias gradients, Adam/SGD, voxel IoU counts.
// All kernels: channels-last tensors, 128-bit (fp32) / 64-bit (bf16) vector accesses along the
// channel axis, grid-stride loops sized in multiples of the SM count, fp32 math.
#include <stdarg.h>
#include "common.cuh"

namespace r2n2 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline int grid_for(long long work_items, int block) {
  long long b = (work_items + block - 1) / block;
  long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ layout: NCHW -> NHWC(pad)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int N, int C,
                                    int H, int W, int Cpad, int Wpad, int xoff) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long total = (long long)N * H * Wpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int wp = (int)(i % Wpad);
    long long t = i / Wpad;
    int h = (int)(t % H);
    int n = (int)(t / H);
    int w = wp - xoff;
    T* out = y + i * Cpad;
    bool in = (w >= 0 && w < W);
    for (int c = 0; c < Cpad; ++c) {
      float v = 0.f;
      if (in && c < C) v = x[(((long long)n * C + c) * H + h) * W + w];
      out[c] = from_float<T>(v);
    }
  }
}

// Row-packed first-layer input: y[n,h,w, kx*C + c] = x[n,c,h, w + kx - (kw-1)/2] (0 outside the
// image, 0 for j >= kw*C).  A kh x kw convolution over C channels becomes a kh x 1 convolution over
// Cpack channels, so the 7x7x3 first layer (K = 147) runs as 7 tensor-core K blocks of 32 instead of
// 49 blocks of 16 zero-padded channels.
template <typename T>
__global__ void __launch_bounds__(128)
nchw_to_rowpack_kernel(const float* __restrict__ x, T* __restrict__ y, int N, int C, int H, int W,
                       int kw, int Cpack) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  // one block per image row (n, h): the C channel rows are staged in shared memory with (kw-1)/2
  // zero columns either side (coalesced fp32 reads, each input element read once), then every
  // thread assembles 16-byte groups of packed channels from a per-block offset table and stores them
  // (fully coalesced 16-byte stores: this kernel is bound by the 5x larger packed output).
  extern __shared__ float rp_smem[];
  constexpr int PN = Pack<T>::N;
  const int pad = (kw - 1) / 2;
  const int pitch = W + 2 * pad;
  float* row = rp_smem;                                  // [C][pitch]
  int* lut = reinterpret_cast<int*>(rp_smem + C * pitch);  // [Cpack] offset into row, or -1
  const int groups = Cpack / PN;
  for (int j = threadIdx.x; j < Cpack; j += blockDim.x) {
    const int kx = j / C, c = j - kx * C;
    lut[j] = (kx < kw) ? c * pitch + kx : (j == kw * C ? -2 : -1);      // -2: the constant-one channel
  }
  for (long long rh = blockIdx.x; rh < (long long)N * H; rh += gridDim.x) {
    const int h = (int)(rh % H);
    const long long n = rh / H;
    __syncthreads();
    for (int i = threadIdx.x; i < C * pitch; i += blockDim.x) {
      const int c = i / pitch, sw = i - c * pitch - pad;
      row[i] = (sw >= 0 && sw < W) ? x[((n * C + c) * H + h) * W + sw] : 0.f;
    }
    __syncthreads();
    T* out = y + rh * W * Cpack;
    for (int i = threadIdx.x; i < W * groups; i += blockDim.x) {
      const int w = i / groups, gq = i - w * groups;
      Pack<T> o;
#pragma unroll
      for (int e = 0; e < PN; ++e) {
        const int off = lut[gq * PN + e];
        o.v[e] = off >= 0 ? row[off + w] : (off == -2 ? 1.f : 0.f);
      }
      o.store(out + (long long)i * PN);
    }
  }
}

// ------------------------------------------------------------------ max pool 2x2/2 pad 1
// raw 16-byte vector -> floats (4 x fp32 / 8 x bf16): lets a kernel keep many loads in flight in 4
// registers each and widen them only when they are consumed
template <typename T>
__device__ __forceinline__ Pack<T> unpack16(const uint4& raw);
template <>
__device__ __forceinline__ Pack<float> unpack16<float>(const uint4& raw) {
  Pack<float> r;
  r.v[0] = __uint_as_float(raw.x); r.v[1] = __uint_as_float(raw.y);
  r.v[2] = __uint_as_float(raw.z); r.v[3] = __uint_as_float(raw.w);
  return r;
}
template <>
__device__ __forceinline__ Pack<__nv_bfloat16> unpack16<__nv_bfloat16>(const uint4& raw) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  Pack<__nv_bfloat16> r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    r.v[2 * j] = f.x; r.v[2 * j + 1] = f.y;
  }
  return r;
}

// the value a float becomes when it is stored as T and read back
template <typename T>
__device__ __forceinline__ float round_as(float x);
template <>
__device__ __forceinline__ float round_as<float>(float x) { return x; }
template <>
__device__ __forceinline__ float round_as<__nv_bfloat16>(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// CODE: additionally one byte per output element for the backward pass (r2n2_maxpool2d_bwd_code): bit k (k = 2 dh + dw,
// window position (2 ho - 1 + dh, 2 wo - 1 + dw)) = position k attains the maximum (ties: every such bit is set, like
// the `==` of maxpool_bwd_kernel); bit 4 + k = the mask source (b if given, else a) is > 0 at position k.  The
// backward pass then reads N*Ho*Wo*C bytes instead of re-reading a (and b): 1/8 (1/16) of their bf16 bytes.
// flat index -> (channel group, wo, ho, image).  32-bit divisions where the index fits (always, for this network):
// a 64-bit division by a run-time value is ~80 instructions, three of them per 16-byte vector rivalled the payload
__device__ __forceinline__ void pool_decode(long long i, int CG, int Wo, int Ho, int& cg, int& wo, int& ho, int& n) {
  if (i <= 0x7fffffffLL) {
    const unsigned u = (unsigned)i;
    const unsigned q1 = u / (unsigned)CG;
    cg = (int)(u - q1 * (unsigned)CG);
    const unsigned q2 = q1 / (unsigned)Wo;
    wo = (int)(q1 - q2 * (unsigned)Wo);
    const unsigned q3 = q2 / (unsigned)Ho;
    ho = (int)(q2 - q3 * (unsigned)Ho);
    n = (int)q3;
  } else {
    cg = (int)(i % CG);
    long long t = i / CG;
    wo = (int)(t % Wo);
    t /= Wo;
    ho = (int)(t % Ho);
    n = (int)(t / Ho);
  }
}

template <typename T, bool HAS_B, bool CODE = false>
__global__ void __launch_bounds__(256, (CODE && HAS_B) ? 3 : 4)
maxpool_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int N, int H, int W, int C,
                   int Ho, int Wo, unsigned char* __restrict__ code = nullptr) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  // one thread per (output pixel, 16-byte channel group); the window's vectors are loaded raw, all before
  // the first use (memory-level parallelism), and widened when compared
  constexpr int PN = Pack<T>::N;
  int CG = C / PN;
  long long total = (long long)N * Ho * Wo * CG;
  float sa[PN], sb[PN];
#pragma unroll
  for (int e = 0; e < PN; ++e) { sa[e] = 0.f; sb[e] = 0.f; }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cg, wo, ho, n;
    pool_decode(i, CG, Wo, Ho, cg, wo, ho, n);
    uint4 ra[4], rb[4];
    bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int h = 2 * ho - 1 + (k >> 1), w = 2 * wo - 1 + (k & 1);
      valid[k] = (h >= 0 && h < H && w >= 0 && w < W);
      ra[k] = make_uint4(0, 0, 0, 0);
      rb[k] = make_uint4(0, 0, 0, 0);
      if (valid[k]) {
        const long long off = (((long long)n * H + h) * W + w) * C + cg * PN;
        ra[k] = *reinterpret_cast<const uint4*>(a + off);
        if (HAS_B) rb[k] = *reinterpret_cast<const uint4*>(b + off);
      }
    }
    Pack<T> m;
#pragma unroll
    for (int e = 0; e < PN; ++e) m.v[e] = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!valid[k]) continue;
      Pack<T> v = unpack16<T>(ra[k]);
      if (HAS_B) {
        Pack<T> u = unpack16<T>(rb[k]);
#pragma unroll
        for (int e = 0; e < PN; ++e) v.v[e] += u.v[e];
      }
#pragma unroll
      for (int e = 0; e < PN; ++e) m.v[e] = fmaxf(m.v[e], v.v[e]);
    }
    m.store(y + i * PN);
    if (CODE) {
      if constexpr (sizeof(T) == 2) {
        // packed: word j of t holds the code of channel 2j in bits 0..7 and of channel 2j+1 in bits 16..23
        uint32_t t4[4] = {0u, 0u, 0u, 0u};
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
        if constexpr (!HAS_B) {
          // a alone: the maximum of bf16 values is a bf16 value -> compare the raw pairs (2 channels per instruction)
          uint32_t m2[4];
          {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(m.v[0], m.v[1]);   // (exact: m holds bf16 values)
            m2[0] = *reinterpret_cast<const uint32_t*>(&h2);
          }
#pragma unroll
          for (int j = 1; j < 4; ++j) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(m.v[2 * j], m.v[2 * j + 1]);
            m2[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (!valid[k]) continue;
            const uint32_t aw[4] = {ra[k].x, ra[k].y, ra[k].z, ra[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 av2 = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
              const uint32_t eq = __heq2_mask(av2, *reinterpret_cast<const __nv_bfloat162*>(&m2[j]));
              const uint32_t sg = __hgt2_mask(av2, zero2);
              t4[j] |= (eq & (0x00010001u << k)) | (sg & (0x00100010u << k));
            }
          }
        } else {
          // a + b is compared in fp32 (the sums are not bf16 values); the signs of b pairwise
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (!valid[k]) continue;
            const Pack<T> av = unpack16<T>(ra[k]);
            const Pack<T> bv = unpack16<T>(rb[k]);
            const uint32_t bw[4] = {rb[k].x, rb[k].y, rb[k].z, rb[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t sg = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&bw[j]), zero2);
              uint32_t w = sg & (0x00100010u << k);
              w |= (av.v[2 * j] + bv.v[2 * j] == m.v[2 * j]) ? (1u << k) : 0u;
              w |= (av.v[2 * j + 1] + bv.v[2 * j + 1] == m.v[2 * j + 1]) ? (0x10000u << k) : 0u;
              t4[j] |= w;
            }
          }
        }
        *reinterpret_cast<uint2*>(code + i * PN) = make_uint2(__byte_perm(t4[0], t4[1], 0x6420), __byte_perm(t4[2], t4[3], 0x6420));
      } else {
        uint32_t cb[PN];
#pragma unroll
        for (int e = 0; e < PN; ++e) cb[e] = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!valid[k]) continue;
          Pack<T> v = unpack16<T>(ra[k]);
          Pack<T> src = v;
          if (HAS_B) {
            src = unpack16<T>(rb[k]);
#pragma unroll
            for (int e = 0; e < PN; ++e) v.v[e] += src.v[e];
          }
#pragma unroll
          for (int e = 0; e < PN; ++e)
            cb[e] |= (v.v[e] == m.v[e] ? (1u << k) : 0u) | (src.v[e] > 0.f ? (16u << k) : 0u);
        }
        *reinterpret_cast<uint32_t*>(code + i * PN) = cb[0] | (cb[1] << 8) | (cb[2] << 16) | (cb[3] << 24);
      }
    }
  }
}

// Max-pool backward from the forward pass's code bytes (see maxpool_fwd_kernel<.., CODE>): dx = dy where the position
// attained the maximum, times LeakyReLU' of the recorded sign when lrelu_mask (code recorded WITHOUT b); dxb (code
// recorded WITH b) = dx * LeakyReLU'(b); column sums as in maxpool_bwd_kernel<.., EXTRA>.  Reads dy + code only.
template <typename T, bool EXTRA>
__global__ void __launch_bounds__(256, 3)
maxpool_bwd_code_kernel(const unsigned char* __restrict__ code, const T* __restrict__ dy, T* __restrict__ dx,
                        int N, int H, int W, int C, int Ho, int Wo, int lrelu_mask, T* __restrict__ dxb,
                        float* __restrict__ colsum_a, float* __restrict__ colsum_b) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  constexpr int PN = Pack<T>::N;
  int CG = C / PN;
  long long total = (long long)N * Ho * Wo * CG;
  float sa[PN], sb[PN];
#pragma unroll
  for (int e = 0; e < PN; ++e) { sa[e] = 0.f; sb[e] = 0.f; }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cg, wo, ho, n;
    pool_decode(i, CG, Wo, Ho, cg, wo, ho, n);
    const uint4 rg = *reinterpret_cast<const uint4*>(dy + i * PN);
    uint32_t cw[2];
    if (PN == 8) {
      const uint2 c2 = *reinterpret_cast<const uint2*>(code + i * PN);
      cw[0] = c2.x; cw[1] = c2.y;
    } else {
      cw[0] = *reinterpret_cast<const uint32_t*>(code + i * PN); cw[1] = 0u;
    }
    if constexpr (sizeof(T) == 2) {
      // packed: two channels per 32-bit word.  t4[j]: code of channel 2j in bits 0..7, of channel 2j+1 in bits 16..23
      const uint32_t gw[4] = {rg.x, rg.y, rg.z, rg.w};
      const uint32_t t4[4] = {__byte_perm(cw[0], 0u, 0x4140), __byte_perm(cw[0], 0u, 0x4342),
                              __byte_perm(cw[1], 0u, 0x4140), __byte_perm(cw[1], 0u, 0x4342)};
      const bool need_slope = lrelu_mask || dxb != nullptr;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int h = 2 * ho - 1 + (k >> 1), w = 2 * wo - 1 + (k & 1);
        if (!(h >= 0 && h < H && w >= 0 && w < W)) continue;
        const long long off = (((long long)n * H + h) * W + w) * C + cg * PN;
        uint32_t ow[4], sw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ow[j] = gw[j] & (((t4[j] >> k) & 0x00010001u) * 0xFFFFu);         // dy where the position attained the maximum
          sw[j] = ow[j];
        }
        if (need_slope) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // LeakyReLU': x 0.01 (fp32 product rounded to bf16, like the unpacked kernels) where the recorded sign is <= 0
            const float lo = __uint_as_float(ow[j] << 16) * kLeak, hi = __uint_as_float(ow[j] & 0xFFFF0000u) * kLeak;
            const __nv_bfloat162 sc = __floats2bfloat162_rn(lo, hi);
            const uint32_t pos = ((t4[j] >> (4 + k)) & 0x00010001u) * 0xFFFFu;
            sw[j] = (ow[j] & pos) | (*reinterpret_cast<const uint32_t*>(&sc) & ~pos);
          }
        }
        const uint32_t* xw = lrelu_mask ? sw : ow;
        *reinterpret_cast<uint4*>(dx + off) = make_uint4(xw[0], xw[1], xw[2], xw[3]);
        if (dxb) *reinterpret_cast<uint4*>(dxb + off) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
        if (EXTRA) {
          if (colsum_a) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              sa[2 * j] += __uint_as_float(xw[j] << 16);
              sa[2 * j + 1] += __uint_as_float(xw[j] & 0xFFFF0000u);
            }
          }
          if (colsum_b) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              sb[2 * j] += __uint_as_float(sw[j] << 16);
              sb[2 * j + 1] += __uint_as_float(sw[j] & 0xFFFF0000u);
            }
          }
        }
      }
    } else {
    const Pack<T> g = unpack16<T>(rg);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int h = 2 * ho - 1 + (k >> 1), w = 2 * wo - 1 + (k & 1);
      if (!(h >= 0 && h < H && w >= 0 && w < W)) continue;
      const long long off = (((long long)n * H + h) * W + w) * C + cg * PN;
      Pack<T> o, ob;
#pragma unroll
      for (int e = 0; e < PN; ++e) {
        const uint32_t cb = (cw[e >> 2] >> (8 * (e & 3))) & 0xffu;
        const float slope = ((cb >> (4 + k)) & 1u) ? 1.f : kLeak;
        o.v[e] = ((cb >> k) & 1u) ? g.v[e] : 0.f;
        ob.v[e] = o.v[e] * slope;
        if (lrelu_mask) o.v[e] = ob.v[e];
      }
      o.store(dx + off);
      if (dxb) ob.store(dxb + off);
      if (EXTRA) {
        if (colsum_a) {
#pragma unroll
          for (int e = 0; e < PN; ++e) sa[e] += round_as<T>(o.v[e]);
        }
        if (colsum_b) {
#pragma unroll
          for (int e = 0; e < PN; ++e) sb[e] += round_as<T>(ob.v[e]);
        }
      }
    }
    }
  }
  if (EXTRA) {
    __shared__ float red[2][512];
    for (int c = threadIdx.x; c < 2 * 512; c += blockDim.x) (&red[0][0])[c] = 0.f;
    __syncthreads();
    const int cg = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % CG);
#pragma unroll
    for (int e = 0; e < PN; ++e) {
      if (colsum_a) atomicAdd(&red[0][cg * PN + e], sa[e]);
      if (colsum_b) atomicAdd(&red[1][cg * PN + e], sb[e]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (colsum_a) atomicAdd(colsum_a + c, red[0][c]);
      if (colsum_b) atomicAdd(colsum_b + c, red[1][c]);
    }
  }
}

// EXTRA: additionally (each optional) dxb = dx * LeakyReLU'(b) -- the pre-activation gradient of the branch b =
// LeakyReLU(z_b) of pool(a + b), written here instead of by a separate pass over dx and b -- and the column
// sums of dx (after the lrelu_mask, if any) / of dxb into colsum_a / colsum_b (fp32 [C], atomically added):
// the bias gradients of the convolutions that produced a and b.  The launch makes gridDim * blockDim a
// multiple of C / PN, so a thread keeps its channel group and sums in registers.
template <typename T, bool HAS_B, bool EXTRA>
__global__ void __launch_bounds__(256, HAS_B ? 2 : 3)        // (8 raw vectors in flight with b: 3 blocks/SM would spill)
maxpool_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ dy,
                   T* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo, int lrelu_mask,
                   T* __restrict__ dxb, float* __restrict__ colsum_a, float* __restrict__ colsum_b) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  // one thread per (output pixel, 16-byte channel group): the up to 8 input vectors of the window stay
  // RAW in registers (all loads issued before the first use), the window maximum is formed first, then
  // each position is re-widened, compared and stored (HBM-bound: registers buy loads in flight)
  constexpr int PN = Pack<T>::N;
  int CG = C / PN;
  long long total = (long long)N * Ho * Wo * CG;
  float sa[PN], sb[PN];
#pragma unroll
  for (int e = 0; e < PN; ++e) { sa[e] = 0.f; sb[e] = 0.f; }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cg, wo, ho, n;
    pool_decode(i, CG, Wo, Ho, cg, wo, ho, n);
    uint4 ra[4], rb[4];
    long long off[4];
    bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int h = 2 * ho - 1 + (k >> 1), w = 2 * wo - 1 + (k & 1);
      valid[k] = (h >= 0 && h < H && w >= 0 && w < W);
      off[k] = (((long long)n * H + h) * W + w) * C + cg * PN;
      ra[k] = make_uint4(0, 0, 0, 0);
      rb[k] = make_uint4(0, 0, 0, 0);
      if (valid[k]) {
        ra[k] = *reinterpret_cast<const uint4*>(a + off[k]);
        if (HAS_B) rb[k] = *reinterpret_cast<const uint4*>(b + off[k]);
      }
    }
    const uint4 rg = *reinterpret_cast<const uint4*>(dy + i * PN);
    Pack<T> m;
#pragma unroll
    for (int e = 0; e < PN; ++e) m.v[e] = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!valid[k]) continue;
      Pack<T> s = unpack16<T>(ra[k]);
      if (HAS_B) {
        Pack<T> u = unpack16<T>(rb[k]);
#pragma unroll
        for (int e = 0; e < PN; ++e) s.v[e] += u.v[e];      // same fp32 sums the forward compared
      }
#pragma unroll
      for (int e = 0; e < PN; ++e) m.v[e] = fmaxf(m.v[e], s.v[e]);
    }
    const Pack<T> g = unpack16<T>(rg);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!valid[k]) continue;
      Pack<T> av = unpack16<T>(ra[k]);
      Pack<T> s = av;
      if (HAS_B) {
        Pack<T> u = unpack16<T>(rb[k]);
#pragma unroll
        for (int e = 0; e < PN; ++e) s.v[e] += u.v[e];
      }
      Pack<T> o;
#pragma unroll
      for (int e = 0; e < PN; ++e) {
        o.v[e] = (s.v[e] == m.v[e]) ? g.v[e] : 0.f;
        if (lrelu_mask) o.v[e] *= lrelu_slope(av.v[e]);
      }
      o.store(dx + off[k]);
      if (EXTRA) {
        // (the sums are formed from the values as STORED, i.e. rounded to T, like a separate pass would see them)
        if (colsum_a) {
#pragma unroll
          for (int e = 0; e < PN; ++e) sa[e] += round_as<T>(o.v[e]);
        }
        if (HAS_B && dxb) {
          const Pack<T> u = unpack16<T>(rb[k]);
          Pack<T> ob;
#pragma unroll
          for (int e = 0; e < PN; ++e) ob.v[e] = o.v[e] * lrelu_slope(u.v[e]);
          ob.store(dxb + off[k]);
          if (colsum_b) {
#pragma unroll
            for (int e = 0; e < PN; ++e) sb[e] += round_as<T>(ob.v[e]);
          }
        }
      }
    }
  }
  if (EXTRA) {
    __shared__ float red[2][512];
    for (int c = threadIdx.x; c < 2 * 512; c += blockDim.x) (&red[0][0])[c] = 0.f;
    __syncthreads();
    // the thread's channel group is fixed: (gridDim * blockDim) % CG == 0
    const int cg = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % CG);
#pragma unroll
    for (int e = 0; e < PN; ++e) {
      if (colsum_a) atomicAdd(&red[0][cg * PN + e], sa[e]);
      if (colsum_b) atomicAdd(&red[1][cg * PN + e], sb[e]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (colsum_a) atomicAdd(colsum_a + c, red[0][c]);
      if (colsum_b) atomicAdd(colsum_b + c, red[1][c]);
    }
  }
}

// ------------------------------------------------------------------ Unpool3D (zero stuffing x2)
// flat index -> (channel group, x, y, z, image): 32-bit divisions where the index fits (see pool_decode)
__device__ __forceinline__ void grid_decode(long long i, int CG, int W, int H, int D, int& cg, int& x, int& y, int& z, int& n) {
  if (i <= 0x7fffffffLL) {
    const unsigned u = (unsigned)i;
    const unsigned q1 = u / (unsigned)CG;
    cg = (int)(u - q1 * (unsigned)CG);
    const unsigned q2 = q1 / (unsigned)W;
    x = (int)(q1 - q2 * (unsigned)W);
    const unsigned q3 = q2 / (unsigned)H;
    y = (int)(q2 - q3 * (unsigned)H);
    const unsigned q4 = q3 / (unsigned)D;
    z = (int)(q3 - q4 * (unsigned)D);
    n = (int)q4;
  } else {
    cg = (int)(i % CG);
    long long t = i / CG;
    x = (int)(t % W); t /= W;
    y = (int)(t % H); t /= H;
    z = (int)(t % D);
    n = (int)(t / D);
  }
}

template <typename T>
__global__ void unpool_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                  T* __restrict__ y, int B, int D, int H, int W, int C) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  constexpr int PN = Pack<T>::N;
  int CG = C / PN;
  int D2 = 2 * D, H2 = 2 * H, W2 = 2 * W;
  long long total = (long long)B * D2 * H2 * W2 * CG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cg, x, yy, z, n;
    grid_decode(i, CG, W2, H2, D2, cg, x, yy, z, n);
    Pack<T> v;
#pragma unroll
    for (int e = 0; e < PN; ++e) v.v[e] = 0.f;
    if (((x | yy | z) & 1) == 0) {
      long long off = ((((long long)n * D + (z >> 1)) * H + (yy >> 1)) * W + (x >> 1)) * C + cg * PN;
      v = Pack<T>::load(a + off);
      if (b) {
        Pack<T> u = Pack<T>::load(b + off);
#pragma unroll
        for (int e = 0; e < PN; ++e) v.v[e] += u.v[e];
      }
    }
    v.store(y + i * PN);
  }
}

template <typename T>
__global__ void unpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int D,
                                  int H, int W, int C) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  constexpr int PN = Pack<T>::N;
  int CG = C / PN;
  long long total = (long long)B * D * H * W * CG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cg, x, yy, z, n;
    grid_decode(i, CG, W, H, D, cg, x, yy, z, n);
    long long off =
        ((((long long)n * 2 * D + 2 * z) * 2 * H + 2 * yy) * 2 * W + 2 * x) * C + cg * PN;
    Pack<T>::load(dy + off).store(dx + i * PN);
  }
}

// ------------------------------------------------------------------ pointwise helpers
template <typename T>
__global__ void add_kernel(const T* a, const T* b, T* y,
                           long long n4) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 u = Vec4<T>::load(a + i * 4), v = Vec4<T>::load(b + i * 4);
    Vec4<T>::store(y + i * 4, make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w));
  }
}

template <typename T>
__global__ void lrelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n4) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = Vec4<T>::load(x + i * 4);
    Vec4<T>::store(y + i * 4, make_float4(lrelu(v.x), lrelu(v.y), lrelu(v.z), lrelu(v.w)));
  }
}

template <typename T>
__global__ void lrelu_bwd_kernel(const T* a, const T* dy,
                                 const T* addend, T* dx, long long n4) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 av = Vec4<T>::load(a + i * 4), g = Vec4<T>::load(dy + i * 4);
    float4 o = make_float4(g.x * lrelu_slope(av.x), g.y * lrelu_slope(av.y),
                           g.z * lrelu_slope(av.z), g.w * lrelu_slope(av.w));
    if (addend) {
      float4 e = Vec4<T>::load(addend + i * 4);
      o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
    }
    Vec4<T>::store(dx + i * 4, o);
  }
}

// lrelu_bwd over a [rows, C] tensor that also accumulates the column sums of dx (as stored) into colsum[C] (fp32): dx is
// the pre-activation gradient of a convolution output, its column sums are that convolution's bias gradient -- no
// separate r2n2_bias_grad pass over dx.  16-byte vectors; gridDim * blockDim is a multiple of C / PN, so a thread
// keeps its channel group and sums in registers (one resident wave, per-block shared-memory reduction, C atomics per block).
template <typename T>
__global__ void __launch_bounds__(256)
lrelu_bwd_colsum_kernel(const T* __restrict__ a, const T* __restrict__ dy, const T* __restrict__ addend, T* __restrict__ dx,
                        long long nvec, int C, float* __restrict__ colsum) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  constexpr int PN = Pack<T>::N;
  const int CG = C / PN;
  float sa[PN];
#pragma unroll
  for (int e = 0; e < PN; ++e) sa[e] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const Pack<T> av = unpack16<T>(*reinterpret_cast<const uint4*>(a + i * PN));
    const Pack<T> g = unpack16<T>(*reinterpret_cast<const uint4*>(dy + i * PN));
    Pack<T> o;
#pragma unroll
    for (int e = 0; e < PN; ++e) o.v[e] = g.v[e] * lrelu_slope(av.v[e]);
    if (addend) {
      const Pack<T> ad = unpack16<T>(*reinterpret_cast<const uint4*>(addend + i * PN));
#pragma unroll
      for (int e = 0; e < PN; ++e) o.v[e] += ad.v[e];
    }
    o.store(dx + i * PN);
#pragma unroll
    for (int e = 0; e < PN; ++e) sa[e] += round_as<T>(o.v[e]);
  }
  __shared__ float red[512];
  for (int c = threadIdx.x; c < 512; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  const int cg = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % CG);
#pragma unroll
  for (int e = 0; e < PN; ++e) atomicAdd(&red[cg * PN + e], sa[e]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(colsum + c, red[c]);
}

// SigmoidLayer / TanhLayer / ComplementLayer / EltwiseMultiplyLayer (lib/layers.py:217-225,649-676)
template <typename T>
__global__ void eltwise_kernel(int op, const T* __restrict__ a, const T* __restrict__ b,
                               T* __restrict__ y, long long n4) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = Vec4<T>::load(a + i * 4);
    float4 o;
    if (op == R2N2_ELT_SIGMOID) {
      o = make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w));
    } else if (op == R2N2_ELT_TANH) {
      o = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
    } else if (op == R2N2_ELT_COMPLEMENT) {
      o = make_float4(1.f - v.x, 1.f - v.y, 1.f - v.z, 1.f - v.w);
    } else if (op == R2N2_ELT_NEG) {
      o = make_float4(-v.x, -v.y, -v.z, -v.w);
    } else {
      float4 u = Vec4<T>::load(b + i * 4);
      if (op == R2N2_ELT_MUL) {
        o = make_float4(v.x * u.x, v.y * u.y, v.z * u.z, v.w * u.w);
      } else if (op == R2N2_ELT_SIGMOID_BWD) {      // a = sigmoid output, b = dL/dy
        o = make_float4(u.x * v.x * (1.f - v.x), u.y * v.y * (1.f - v.y), u.z * v.z * (1.f - v.z),
                        u.w * v.w * (1.f - v.w));
      } else {                                      // R2N2_ELT_TANH_BWD: a = tanh output
        o = make_float4(u.x * (1.f - v.x * v.x), u.y * (1.f - v.y * v.y), u.z * (1.f - v.z * v.z),
                        u.w * (1.f - v.w * v.w));
      }
    }
    Vec4<T>::store(y + i * 4, o);
  }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ s, T* __restrict__ d, long long n) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x)
    Vec4<T>::store(d + i * 4, *reinterpret_cast<const float4*>(s + i * 4));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    long long j = (n4 << 2) + threadIdx.x;
    d[j] = from_float<T>(s[j]);
  }
}

// ------------------------------------------------------------------ hi / lo operand split (COMPUTE = 'bf16x3')
// x = hi + lo + O(2^-17 |x|):  hi = bf16_rn(x), lo = bf16_rn(x - hi).  A contraction of split operands
// (hi*hi + lo*hi + hi*lo, fp32 accumulation) carries ~16 mantissa bits instead of bf16's 8.
__global__ void split_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  const long long n8 = n >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(s)[2 * i], b = reinterpret_cast<const float4*>(s)[2 * i + 1];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    Pack<__nv_bfloat16> ph, pl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = __bfloat162float(__float2bfloat16_rn(v[j]));
      ph.v[j] = h;
      pl.v[j] = v[j] - h;          // exact in fp32 (Sterbenz), rounded to bf16 by the store
    }
    ph.store(hi + 8 * i);
    pl.store(lo + 8 * i);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long j = (n8 << 3) + threadIdx.x;
    const __nv_bfloat16 h = __float2bfloat16_rn(s[j]);
    hi[j] = h;
    lo[j] = __float2bfloat16_rn(s[j] - __bfloat162float(h));
  }
}

// ------------------------------------------------------------------ conv-GRU gate math
// hidden grid rows: pr = b*64 + pos ; 128 channels -> 32 float4 groups per row
template <typename T>
__global__ void gru_ur_fwd_kernel(const T* __restrict__ conv_ur, const T* __restrict__ xp_ur,
                                  const float* __restrict__ bias, const T* __restrict__ h,
                                  T* __restrict__ u, T* __restrict__ r, T* __restrict__ rh,
                                  long long rows) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long total = rows * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pr = i >> 5;
    int c = (int)(i & 31) * 4;
    float4 au = Vec4<T>::load(xp_ur + pr * 256 + c);
    float4 ar = Vec4<T>::load(xp_ur + pr * 256 + 128 + c);
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (conv_ur) {
      float4 cu = Vec4<T>::load(conv_ur + pr * 256 + c);
      float4 cr = Vec4<T>::load(conv_ur + pr * 256 + 128 + c);
      au.x += cu.x; au.y += cu.y; au.z += cu.z; au.w += cu.w;
      ar.x += cr.x; ar.y += cr.y; ar.z += cr.z; ar.w += cr.w;
      hv = Vec4<T>::load(h + pr * 128 + c);
    }
    float4 bu = *reinterpret_cast<const float4*>(bias + c);
    float4 br = *reinterpret_cast<const float4*>(bias + 128 + c);
    float4 uo = make_float4(sigmoidf_(au.x + bu.x), sigmoidf_(au.y + bu.y),
                            sigmoidf_(au.z + bu.z), sigmoidf_(au.w + bu.w));
    float4 ro = make_float4(sigmoidf_(ar.x + br.x), sigmoidf_(ar.y + br.y),
                            sigmoidf_(ar.z + br.z), sigmoidf_(ar.w + br.w));
    Vec4<T>::store(u + pr * 128 + c, uo);
    Vec4<T>::store(r + pr * 128 + c, ro);
    Vec4<T>::store(rh + pr * 128 + c,
                   make_float4(ro.x * hv.x, ro.y * hv.y, ro.z * hv.z, ro.w * hv.w));
  }
}

template <typename T>
__global__ void gru_out_fwd_kernel(const T* __restrict__ conv_c, const T* __restrict__ xp_c,
                                   const float* __restrict__ bias, const T* __restrict__ u,
                                   const T* __restrict__ h, T* __restrict__ cc,
                                   T* __restrict__ hn, long long rows) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long total = rows * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long off = i * 4;
    int c = (int)(i & 31) * 4;
    float4 a = Vec4<T>::load(xp_c + off);
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (conv_c) {
      float4 cv = Vec4<T>::load(conv_c + off);
      a.x += cv.x; a.y += cv.y; a.z += cv.z; a.w += cv.w;
      hv = Vec4<T>::load(h + off);
    }
    float4 bv = *reinterpret_cast<const float4*>(bias + c);
    float4 uv = Vec4<T>::load(u + off);
    float4 cv = make_float4(tanhf(a.x + bv.x), tanhf(a.y + bv.y), tanhf(a.z + bv.z),
                            tanhf(a.w + bv.w));
    Vec4<T>::store(cc + off, cv);
    Vec4<T>::store(hn + off, make_float4(uv.x * hv.x + (1.f - uv.x) * cv.x,
                                         uv.y * hv.y + (1.f - uv.y) * cv.y,
                                         uv.z * hv.z + (1.f - uv.z) * cv.z,
                                         uv.w * hv.w + (1.f - uv.w) * cv.w));
  }
}

template <typename T>
__global__ void gru_out_bwd_kernel(const T* __restrict__ dh, const T* __restrict__ u,
                                   const T* __restrict__ cc, const T* __restrict__ h,
                                   T* __restrict__ da_ur, T* __restrict__ da_c,
                                   T* __restrict__ dh_prev, long long rows) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long total = rows * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pr = i >> 5;
    int c = (int)(i & 31) * 4;
    long long off = i * 4;
    float4 g = Vec4<T>::load(dh + off), uv = Vec4<T>::load(u + off), cv = Vec4<T>::load(cc + off);
    float4 hv = h ? Vec4<T>::load(h + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dac, dau, dhp;
#define R2N2_GRU_BWD(e)                                   \
  dac.e = g.e * (1.f - uv.e) * (1.f - cv.e * cv.e);       \
  dau.e = g.e * (hv.e - cv.e) * uv.e * (1.f - uv.e);      \
  dhp.e = g.e * uv.e;
    R2N2_GRU_BWD(x) R2N2_GRU_BWD(y) R2N2_GRU_BWD(z) R2N2_GRU_BWD(w)
#undef R2N2_GRU_BWD
    Vec4<T>::store(da_c + off, dac);
    Vec4<T>::store(da_ur + pr * 256 + c, dau);
    Vec4<T>::store(dh_prev + off, dhp);
  }
}

template <typename T>
__global__ void gru_r_bwd_kernel(const T* __restrict__ d_rh, const T* __restrict__ h,
                                 const T* __restrict__ r, T* __restrict__ da_ur,
                                 T* __restrict__ dh_prev, long long rows) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long total = rows * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pr = i >> 5;
    int c = (int)(i & 31) * 4;
    long long off = i * 4;
    float4 g = Vec4<T>::load(d_rh + off), hv = Vec4<T>::load(h + off), rv = Vec4<T>::load(r + off);
    float4 dp = Vec4<T>::load(dh_prev + off);
    float4 dar = make_float4(g.x * hv.x * rv.x * (1.f - rv.x), g.y * hv.y * rv.y * (1.f - rv.y),
                             g.z * hv.z * rv.z * (1.f - rv.z), g.w * hv.w * rv.w * (1.f - rv.w));
    dp.x += g.x * rv.x; dp.y += g.y * rv.y; dp.z += g.z * rv.z; dp.w += g.w * rv.w;
    Vec4<T>::store(da_ur + pr * 256 + 128 + c, dar);
    Vec4<T>::store(dh_prev + off, dp);
  }
}

// ------------------------------------------------------------------ conv-LSTM gate math
template <typename T>
__global__ void lstm_fwd_kernel(const T* __restrict__ conv, const T* __restrict__ xp,
                                const float* __restrict__ bias, const T* __restrict__ s,
                                T* __restrict__ fig, T* __restrict__ sn, T* __restrict__ hn,
                                long long rows) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long total = rows * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pr = i >> 5;
    int c = (int)(i & 31) * 4;
    float4 a[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a[k] = Vec4<T>::load(xp + pr * 384 + k * 128 + c);
      if (conv) {
        float4 cv = Vec4<T>::load(conv + pr * 384 + k * 128 + c);
        a[k].x += cv.x; a[k].y += cv.y; a[k].z += cv.z; a[k].w += cv.w;
      }
      float4 bv = *reinterpret_cast<const float4*>(bias + k * 128 + c);
      a[k].x += bv.x; a[k].y += bv.y; a[k].z += bv.z; a[k].w += bv.w;
    }
    float4 sv = s ? Vec4<T>::load(s + pr * 128 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 f = make_float4(sigmoidf_(a[0].x), sigmoidf_(a[0].y), sigmoidf_(a[0].z), sigmoidf_(a[0].w));
    float4 ig = make_float4(sigmoidf_(a[1].x), sigmoidf_(a[1].y), sigmoidf_(a[1].z), sigmoidf_(a[1].w));
    float4 g = make_float4(tanhf(a[2].x), tanhf(a[2].y), tanhf(a[2].z), tanhf(a[2].w));
    float4 so = make_float4(f.x * sv.x + ig.x * g.x, f.y * sv.y + ig.y * g.y,
                            f.z * sv.z + ig.z * g.z, f.w * sv.w + ig.w * g.w);
    Vec4<T>::store(fig + pr * 384 + c, f);
    Vec4<T>::store(fig + pr * 384 + 128 + c, ig);
    Vec4<T>::store(fig + pr * 384 + 256 + c, g);
    Vec4<T>::store(sn + pr * 128 + c, so);
    Vec4<T>::store(hn + pr * 128 + c, make_float4(tanhf(so.x), tanhf(so.y), tanhf(so.z), tanhf(so.w)));
  }
}

template <typename T>
__global__ void lstm_bwd_kernel(const T* __restrict__ dh, const T* __restrict__ ds_in,
                                const T* __restrict__ fig, const T* __restrict__ s,
                                const T* __restrict__ hn, T* __restrict__ da,
                                T* __restrict__ ds_prev, long long rows) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long total = rows * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pr = i >> 5;
    int c = (int)(i & 31) * 4;
    long long off = pr * 128 + c;
    float4 g = Vec4<T>::load(dh + off), hv = Vec4<T>::load(hn + off);
    float4 dsi = ds_in ? Vec4<T>::load(ds_in + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 sv = s ? Vec4<T>::load(s + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 f = Vec4<T>::load(fig + pr * 384 + c), ig = Vec4<T>::load(fig + pr * 384 + 128 + c),
           gg = Vec4<T>::load(fig + pr * 384 + 256 + c);
    float4 daf, dai, dag, dsp;
#define R2N2_LSTM_BWD(e)                                        \
  {                                                             \
    float dst = g.e * (1.f - hv.e * hv.e) + dsi.e;              \
    daf.e = dst * sv.e * f.e * (1.f - f.e);                     \
    dai.e = dst * gg.e * ig.e * (1.f - ig.e);                   \
    dag.e = dst * ig.e * (1.f - gg.e * gg.e);                   \
    dsp.e = dst * f.e;                                          \
  }
    R2N2_LSTM_BWD(x) R2N2_LSTM_BWD(y) R2N2_LSTM_BWD(z) R2N2_LSTM_BWD(w)
#undef R2N2_LSTM_BWD
    Vec4<T>::store(da + pr * 384 + c, daf);
    Vec4<T>::store(da + pr * 384 + 128 + c, dai);
    Vec4<T>::store(da + pr * 384 + 256 + c, dag);
    Vec4<T>::store(ds_prev + off, dsp);
  }
}

// ------------------------------------------------------------------ voxel softmax + CE (fwd+bwd)
template <typename TD>
__global__ void softmax_ce3d_kernel(const float* __restrict__ logits, const float* __restrict__ y,
                                    float* __restrict__ pred, float* __restrict__ loss_sum,
                                    TD* __restrict__ dlogits, long long nvox_total, int nv,
                                    float grad_scale, int cstride) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  const long long plane = (long long)nv * nv;  // one (y,x) plane
  float local = 0.f;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox_total;
       v += (long long)gridDim.x * blockDim.x) {
    // v = ((b*nv + z)*nv + yy)*nv + xx ; reference layout offset of channel c:
    long long bz = v / plane;
    long long yx = v - bz * plane;
    long long ref0 = (bz * 2) * plane + yx;
    float2 x = *reinterpret_cast<const float2*>(logits + v * cstride);
    float e0 = expf(x.x), e1 = expf(x.y);  // no max subtraction: lib/layers.py:585-589
    float s = e0 + e1;
    float inv = 1.f / s;
    float p0 = e0 * inv, p1 = e1 * inv;
    float y0 = 0.f, y1 = 0.f;
    if (y) {
      y0 = y[ref0];
      y1 = y[ref0 + plane];
    }
    if (pred) {
      pred[ref0] = p0;
      pred[ref0 + plane] = p1;
    }
    if (loss_sum) local += -(y0 * x.x + y1 * x.y) + logf(s);
    if (dlogits) {
      TD* dl = dlogits + v * cstride;
      constexpr int PN = Pack<TD>::N;
      if ((cstride % PN) == 0) {                       // padded logit channels: 16-byte stores
        Pack<TD> o;
#pragma unroll
        for (int e = 0; e < PN; ++e) o.v[e] = 0.f;
        for (int c = PN; c < cstride; c += PN) o.store(dl + c);
        o.v[0] = (p0 - y0) * grad_scale;
        o.v[1] = (p1 - y1) * grad_scale;
        o.store(dl);
      } else {
        dl[0] = from_float<TD>((p0 - y0) * grad_scale);
        dl[1] = from_float<TD>((p1 - y1) * grad_scale);
        for (int c = 2; c < cstride; ++c) dl[c] = from_float<TD>(0.f);
      }
    }
  }
  if (loss_sum) {
    __shared__ float red[32];
    local = warp_sum(local);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = local;
    __syncthreads();
    if (wid == 0) {
      float t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) atomicAdd(loss_sum, t);
    }
  }
}

// ------------------------------------------------------------------ bias gradient (column sums)
// grid.x covers channel groups of 32*PN (a warp reads 512 contiguous bytes of one row), grid.y
// splits the rows; 8 row-lanes per block, fp32 register accumulation, smem reduce, one atomic per
// (block, channel).  Scalar tail path for C not a multiple of the vector width (conv11: C = 2).
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ dy, float* __restrict__ db, long long rows,
                                 int C, long long rows_per_block) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  constexpr int PN = Pack<T>::N;
  __shared__ float red[8][32 * PN + 1];
  const int cx = threadIdx.x, ry = threadIdx.y;
  long long r0 = blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[PN];
#pragma unroll
  for (int e = 0; e < PN; ++e) acc[e] = 0.f;
  const bool vec = (C % PN) == 0;
  const int c0 = (blockIdx.x * 32 + cx) * PN;
  if (vec) {
    if (c0 < C)
      for (long long r = r0 + ry; r < r1; r += 8) {
        Pack<T> v = Pack<T>::load(dy + r * C + c0);
#pragma unroll
        for (int e = 0; e < PN; ++e) acc[e] += v.v[e];
      }
  } else {
    for (int e = 0; e < PN; ++e)
      if (c0 + e < C)
        for (long long r = r0 + ry; r < r1; r += 8) acc[e] += to_float<T>(dy[r * C + c0 + e]);
  }
#pragma unroll
  for (int e = 0; e < PN; ++e) red[ry][cx * PN + e] = acc[e];
  __syncthreads();
  if (ry == 0) {
#pragma unroll
    for (int e = 0; e < PN; ++e) {
      if (c0 + e >= C) continue;
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][cx * PN + e];
      atomicAdd(db + c0 + e, t);
    }
  }
}

// Flat variant for C a multiple of the vector width with at most 256 vectors per row (every conv
// layer): the block walks the [rows][C] matrix as ONE contiguous stream -- thread t owns the
// channel group t % (C/PN) of rows t / (C/PN) + k * rows_per_iter, so a warp always reads 512
// contiguous bytes whatever C is (the kernel above leaves 28 of 32 lanes idle at C = 32).
template <typename T>
__global__ void __launch_bounds__(256)
bias_grad_flat_kernel(const T* __restrict__ dy, float* __restrict__ db, long long rows, int C,
                      long long rows_per_block) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  constexpr int PN = Pack<T>::N;
  __shared__ float red[256 * PN];
  const int CV = C / PN;
  const int rpi = 256 / CV;                       // rows per block iteration
  const int tid = threadIdx.x;
  const int rr = tid / CV, cg = tid - rr * CV;
  float acc[PN];
#pragma unroll
  for (int e = 0; e < PN; ++e) acc[e] = 0.f;
  if (rr < rpi) {
    const long long r0 = blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    long long r = r0 + rr;
    const T* p = dy + r * C + cg * PN;
    const long long step = (long long)rpi * C;
    for (; r + 3 * rpi < r1; r += 4 * rpi, p += 4 * step) {
      Pack<T> v0 = Pack<T>::load(p), v1 = Pack<T>::load(p + step), v2 = Pack<T>::load(p + 2 * step),
              v3 = Pack<T>::load(p + 3 * step);
#pragma unroll
      for (int e = 0; e < PN; ++e) acc[e] += (v0.v[e] + v1.v[e]) + (v2.v[e] + v3.v[e]);
    }
    for (; r < r1; r += rpi, p += step) {
      Pack<T> v = Pack<T>::load(p);
#pragma unroll
      for (int e = 0; e < PN; ++e) acc[e] += v.v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < PN; ++e) red[tid * PN + e] = acc[e];   // = red[rr * C + cg * PN + e]
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float t = 0.f;
    for (int k = 0; k < rpi; ++k) t += red[k * C + c];
    atomicAdd(db + c, t);
  }
}

// ------------------------------------------------------------------ weight transpose (+cast)
// w[Cout][taps][Cin] -> wt[Cin][taps-1-t][Cout]
template <typename T>
__global__ void weight_transpose_kernel(const float* __restrict__ w, T* __restrict__ wt, int Cout,
                                        int taps, int Cin) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  __shared__ float tile[32][33];
  int t = blockIdx.z;
  int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int co = co0 + j, ci = ci0 + threadIdx.x;
    tile[j][threadIdx.x] =
        (co < Cout && ci < Cin) ? w[((long long)co * taps + t) * Cin + ci] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int ci = ci0 + j, co = co0 + threadIdx.x;
    if (ci < Cin && co < Cout)
      wt[((long long)ci * taps + (taps - 1 - t)) * Cout + co] = from_float<T>(tile[threadIdx.x][j]);
  }
}

// All weight tensors of a network in ONE launch: table[e] = {src offset (floats), dst offset
// (elements), Cout, taps, Cin, first tile}; a block finds its entry by scanning the (<= 64-entry) table.
template <typename T>
__global__ void weight_transpose_batch_kernel(const float* __restrict__ base, T* __restrict__ wt_base,
                                              const long long* __restrict__ table, int n) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  // tile = 64 output channels x 32 input channels: 128-byte row reads of w, and the transposed rows are written
  // as 64 consecutive co (128 bytes of bf16) per ci -- full lines both ways
  __shared__ float tile[64][33];
  const long long bid = blockIdx.x;
  int e = 0;
  while (e + 1 < n && __ldg(table + (e + 1) * 6 + 5) <= bid) ++e;
  const long long* row = table + e * 6;
  const float* w = base + __ldg(row + 0);
  T* wt = wt_base + __ldg(row + 1);
  const int Cout = (int)__ldg(row + 2), taps = (int)__ldg(row + 3), Cin = (int)__ldg(row + 4);
  long long local = bid - __ldg(row + 5);
  const int tiles_ci = (Cin + 31) / 32, tiles_co = (Cout + 63) / 64;
  const int ci0 = (int)(local % tiles_ci) * 32; local /= tiles_ci;
  const int co0 = (int)(local % tiles_co) * 64; local /= tiles_co;
  const int t = (int)local;
  for (int j = threadIdx.y; j < 64; j += 8) {
    int co = co0 + j, ci = ci0 + threadIdx.x;
    tile[j][threadIdx.x] = (co < Cout && ci < Cin) ? w[((long long)co * taps + t) * Cin + ci] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int ci = ci0 + j, co = co0 + 2 * threadIdx.x;
    if (ci >= Cin) continue;
    T* dst = wt + ((long long)ci * taps + (taps - 1 - t)) * Cout + co;
    if (co + 1 < Cout && (Cout & 1) == 0) {
      if constexpr (sizeof(T) == 2) {
        __nv_bfloat162 h = __floats2bfloat162_rn(tile[2 * threadIdx.x][j], tile[2 * threadIdx.x + 1][j]);
        *reinterpret_cast<__nv_bfloat162*>(dst) = h;
      } else {
        *reinterpret_cast<float2*>(dst) = make_float2(tile[2 * threadIdx.x][j], tile[2 * threadIdx.x + 1][j]);
      }
    } else {
      if (co < Cout) dst[0] = from_float<T>(tile[2 * threadIdx.x][j]);
      if (co + 1 < Cout) dst[1] = from_float<T>(tile[2 * threadIdx.x + 1][j]);
    }
  }
}

// Same, from the bf16 operand mirror (identical layout and offsets as the fp32 master buffer): half the read
// traffic, 64 x 64 tiles (128-byte segments both ways, a quarter of the blocks of the fp32 version).
__global__ void weight_transpose_batch_bf16_kernel(const __nv_bfloat16* __restrict__ base,
                                                   __nv_bfloat16* __restrict__ wt_base,
                                                   const long long* __restrict__ table, int n) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  __shared__ __nv_bfloat16 tile[64][66];
  const long long bid = blockIdx.x;
  int e = 0;
  while (e + 1 < n && __ldg(table + (e + 1) * 6 + 5) <= bid) ++e;
  const long long* row = table + e * 6;
  const __nv_bfloat16* w = base + __ldg(row + 0);
  __nv_bfloat16* wt = wt_base + __ldg(row + 1);
  const int Cout = (int)__ldg(row + 2), taps = (int)__ldg(row + 3), Cin = (int)__ldg(row + 4);
  long long local = bid - __ldg(row + 5);
  const int tiles_ci = (Cin + 63) / 64, tiles_co = (Cout + 63) / 64;
  const int ci0 = (int)(local % tiles_ci) * 64; local /= tiles_ci;
  const int co0 = (int)(local % tiles_co) * 64; local /= tiles_co;
  const int t = (int)local;
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  // 16-byte path (every layer of the networks: channel counts and offsets are multiples of 8): 2 vector loads and 2
  // vector stores per thread instead of 8 + 8 four-byte accesses -- the 4-byte version moved 144 MB at 1.9 TB/s
  if ((Cin & 7) == 0 && (Cout & 7) == 0 && ((__ldg(row + 0) | __ldg(row + 1)) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (reinterpret_cast<uintptr_t>(wt_base) & 15) == 0) {
    const int tid = threadIdx.y * 32 + threadIdx.x;
#pragma unroll
    for (int q = tid; q < 512; q += 256) {
      const int j = q >> 3, vi = q & 7;
      const int co = co0 + j, ci = ci0 + vi * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (co < Cout && ci < Cin) v = *reinterpret_cast<const uint4*>(w + ((long long)co * taps + t) * Cin + ci);
      uint32_t* tp = reinterpret_cast<uint32_t*>(&tile[j][vi * 8]);      // (row pitch 132 bytes: 4-byte aligned)
      tp[0] = v.x; tp[1] = v.y; tp[2] = v.z; tp[3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int q = tid; q < 512; q += 256) {
      const int j = q >> 3, vo = q & 7;
      const int ci = ci0 + j, co = co0 + vo * 8;
      if (ci >= Cin || co >= Cout) continue;
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = __halves2bfloat162(tile[vo * 8 + 2 * e][j], tile[vo * 8 + 2 * e + 1][j]);
        o[e] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(wt + ((long long)ci * taps + (taps - 1 - t)) * Cout + co) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    return;
  }
  for (int j = threadIdx.y; j < 64; j += 8) {
    const int co = co0 + j, ci = ci0 + 2 * threadIdx.x;
    __nv_bfloat162 v = __halves2bfloat162(zero, zero);
    if (co < Cout && ci + 1 < Cin) {
      v = *reinterpret_cast<const __nv_bfloat162*>(w + ((long long)co * taps + t) * Cin + ci);   // (Cin even, ci even)
    } else if (co < Cout && ci < Cin) {
      v = __halves2bfloat162(w[((long long)co * taps + t) * Cin + ci], zero);
    }
    tile[j][2 * threadIdx.x] = __low2bfloat16(v);
    tile[j][2 * threadIdx.x + 1] = __high2bfloat16(v);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 64; j += 8) {
    const int ci = ci0 + j, co = co0 + 2 * threadIdx.x;
    if (ci >= Cin || co >= Cout) continue;
    __nv_bfloat16* dst = wt + ((long long)ci * taps + (taps - 1 - t)) * Cout + co;
    if (co + 1 < Cout) *reinterpret_cast<__nv_bfloat162*>(dst) = __halves2bfloat162(tile[2 * threadIdx.x][j], tile[2 * threadIdx.x + 1][j]);
    else dst[0] = tile[2 * threadIdx.x][j];
  }
}

// ------------------------------------------------------------------ optimisers
template <typename TM, bool kMirror>
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, TM* __restrict__ mirror,
                            long long n, long long n_decay, float lr_t, float b1, float b2,
                            float eps, float wd, float gscale) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  long long n4 = (n + 3) >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    long long base = i * 4;
    if (base + 3 < n) {
      float4 pv = *reinterpret_cast<float4*>(p + base);
      float4 gv = *reinterpret_cast<const float4*>(g + base);
      float4 mv = *reinterpret_cast<float4*>(m + base);
      float4 vv = *reinterpret_cast<float4*>(v + base);
      float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float rg = gp[e] * gscale;
        if (base + e < n_decay) rg += wd * pp[e];
        mp[e] = b1 * mp[e] + (1.f - b1) * rg;
        vp[e] = b2 * vp[e] + (1.f - b2) * rg * rg;
        pp[e] = pp[e] - lr_t * mp[e] / (sqrtf(vp[e]) + eps);
      }
      *reinterpret_cast<float4*>(p + base) = pv;
      *reinterpret_cast<float4*>(m + base) = mv;
      *reinterpret_cast<float4*>(v + base) = vv;
      if (kMirror) Vec4<TM>::store(mirror + base, pv);
    } else {
      for (long long j = base; j < n; ++j) {
        float rg = g[j] * gscale;
        if (j < n_decay) rg += wd * p[j];
        float mm = b1 * m[j] + (1.f - b1) * rg;
        float vv = b2 * v[j] + (1.f - b2) * rg * rg;
        float pn = p[j] - lr_t * mm / (sqrtf(vv) + eps);
        m[j] = mm; v[j] = vv; p[j] = pn;
        if (kMirror) mirror[j] = from_float<TM>(pn);
      }
    }
  }
}

template <typename TM, bool kMirror>
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                           float* __restrict__ vel, TM* __restrict__ mirror, long long n,
                           long long n_decay, float lr, float mom, float wd, float gscale) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    float rg = g[j] * gscale;
    if (j < n_decay) rg += wd * p[j];
    float add = mom * vel[j] - lr * rg;   // lib/solver.py:68-70
    vel[j] = add;
    float pn = p[j] + add;
    p[j] = pn;
    if (kMirror) mirror[j] = from_float<TM>(pn);
  }
}

// ------------------------------------------------------------------ voxel IoU counts
__global__ void voxel_eval_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                  float thresh, unsigned long long* __restrict__ counts, int nv) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  int b = blockIdx.y;
  long long plane = (long long)nv * nv;
  long long nvox = plane * nv;
  unsigned int c[5] = {0, 0, 0, 0, 0};
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox;
       v += (long long)gridDim.x * blockDim.x) {
    long long z = v / plane, yx = v - z * plane;
    long long o0 = ((b * (long long)nv + z) * 2) * plane + yx;
    bool occ = pred[o0 + plane] >= thresh;
    bool g1 = gt[o0 + plane] != 0.f, g0 = gt[o0] != 0.f;
    c[0] += (occ != g1); c[1] += (occ && g1); c[2] += (occ || g1);
    c[3] += (occ && g0); c[4] += (!occ && g1);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    unsigned int t = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(counts + b * 5 + k, (unsigned long long)t);
  }
}

// ------------------------------------------------------------------ input pipeline (lib/data_augmentation.py:6-70)
// One thread per output pixel: background compositing on alpha == 0 (add_random_color_background, 40-54), crop
// at (cr, cc) and horizontal flip (image_transform, 6-28 / crop_center, 31-38), /255 in fp32 (66), written
// channels-first like batch_img (lib/data_process.py:145-148).  params[n] = {cr, cc, flip, r, g, b}.
template <typename T>
__global__ void preprocess_views_kernel(const uchar4* __restrict__ rgba, const int* __restrict__ params,
                                        T* __restrict__ out, int n_img, int Hs, int Ws, int Ho, int Wo) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  const long long total = (long long)n_img * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    long long t = i / Wo;
    const int y = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const int* pr = params + n * 6;
    const int cr = pr[0], cc = pr[1], flip = pr[2];
    const int sx = cc + (flip ? (Wo - 1 - x) : x), sy = cr + y;
    const uchar4 px = rgba[((long long)n * Hs + sy) * Ws + sx];
    const bool bg = px.w == 0;
    const float r = (float)(bg ? pr[3] : (int)px.x), g = (float)(bg ? pr[4] : (int)px.y),
                b = (float)(bg ? pr[5] : (int)px.z);
    const long long plane = (long long)Ho * Wo;
    T* o = out + (long long)n * 3 * plane + (long long)y * Wo + x;
    o[0] = from_float<T>(__fdiv_rn(r, 255.f));
    o[plane] = from_float<T>(__fdiv_rn(g, 255.f));
    o[2 * plane] = from_float<T>(__fdiv_rn(b, 255.f));
  }
}

// ------------------------------------------------------------------ voxel labels
// lib/data_process.py:153-154: batch_voxel[b, :, 0] = voxel_data < 1, batch_voxel[b, :, 1] = voxel_data for
// a boolean occupancy grid; occ [B,nv,nv,nv] bytes (0 = empty, 1 = occupied, anything else = slot of a short
// last batch that was never filled: both channels stay 0 like the reference's np.zeros rows), y [B,nv,2,nv,nv] fp32.
__global__ void voxel_labels_kernel(const unsigned char* __restrict__ occ, float* __restrict__ y, int nv,
                                    long long total) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  const int plane = nv * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long slab = i / plane;                        // b * nv + i0
    const int jk = (int)(i - slab * plane);
    const unsigned char v = occ[i];
    y[(slab * 2) * plane + jk] = v == 0 ? 1.f : 0.f;
    y[(slab * 2 + 1) * plane + jk] = v == 1 ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------ average precision per sample
// lib/test_net.py:69-70: sklearn.metrics.average_precision_score(gt[:,1].flatten(), pred[:,1].flatten()):
//   AP = sum over DISTINCT score values s (descending) of (R(s) - R(prev)) * P(s),
//   P(s) = TP(>=s) / #(>=s), R(s) = TP(>=s) / #positives  (0 if there is no positive).
// One CTA per sample: the nv^3 (<= 32768) scores are sorted in shared memory (bitonic network on the
// order-preserving 32-bit integer image of the floats, the 0/1 labels are swapped along in a byte array),
// then two block scans give TP at every position and TP at the end of
// the previous run of equal scores; the sum is formed in double.
constexpr int kApThreads = 1024;
constexpr int kApMax = 32768;

__device__ __forceinline__ unsigned int float_order_key(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // ascending integer order == ascending float order
}

__global__ void __launch_bounds__(kApThreads, 1)
average_precision_kernel(const float* __restrict__ pred, const float* __restrict__ gt, double* __restrict__ ap,
                         int nv) {
  r2n2::pdl_trigger();
  r2n2::pdl_wait();
  extern __shared__ unsigned int ap_smem[];                 // [npad] 32-bit score keys, then [npad] 8-bit labels
  __shared__ unsigned int part[kApThreads];
  __shared__ double red[kApThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int plane = nv * nv, n = plane * nv;
  int npad = 1;
  while (npad < n) npad <<= 1;
  unsigned int* keys = ap_smem;                             // sorted DESCENDING; padding (0) sorts last
  unsigned char* lab = reinterpret_cast<unsigned char*>(ap_smem + npad);
  for (int v = tid; v < npad; v += kApThreads) {
    unsigned int k = 0;
    unsigned char l = 0;
    if (v < n) {
      const int z = v / plane, yx = v - z * plane;
      const long long o1 = ((long long)(b * nv + z) * 2 + 1) * plane + yx;
      k = float_order_key(pred[o1]);
      if (k == 0) k = 1;                                    // (only a NaN pattern maps to 0)
      l = gt[o1] != 0.f ? 1 : 0;
    }
    keys[v] = k;
    lab[v] = l;
  }
  __syncthreads();
  // bitonic sort, descending (ties keep whatever label order: AP only looks at whole runs of equal scores)
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += kApThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned int a = keys[i], c = keys[l];
          const bool desc = (i & k) == 0;
          if (desc ? (a < c) : (a > c)) {
            keys[i] = c; keys[l] = a;
            const unsigned char t = lab[i]; lab[i] = lab[l]; lab[l] = t;
          }
        }
      }
      __syncthreads();
    }
  // each thread owns a contiguous chunk; inclusive scan of the labels
  const int per = npad / kApThreads > 0 ? npad / kApThreads : 1;
  const int i0 = tid * per;
  unsigned int local = 0;
  for (int i = i0; i < i0 + per && i < n; ++i) local += lab[i];
  part[tid] = local;
  __syncthreads();
  for (int o = 1; o < kApThreads; o <<= 1) {                // Hillis-Steele inclusive scan of the partial sums
    unsigned int v = part[tid];
    if (tid >= o) v += part[tid - o];
    __syncthreads();
    part[tid] = v;
    __syncthreads();
  }
  const unsigned int total_pos = part[kApThreads - 1];
  const unsigned int tp0 = part[tid] - local;               // positives before this chunk
  // a run of equal scores ends at i when keys[i] != keys[i+1] (or i = n-1).  Pass 1: TP at the LAST run end
  // of every chunk (0xffffffff if it has none), so that each chunk can find TP at the previous run end.
  unsigned int run_tp = tp0, last_end_tp = 0xffffffffu;
  for (int i = i0; i < i0 + per && i < n; ++i) {
    run_tp += lab[i];
    if (i == n - 1 || keys[i] != keys[i + 1]) last_end_tp = run_tp;
  }
  __syncthreads();
  part[tid] = last_end_tp;
  __syncthreads();
  unsigned int prev_tp = 0;
  for (int t = tid - 1; t >= 0; --t) {
    const unsigned int v = part[t];
    if (v != 0xffffffffu) { prev_tp = v; break; }
  }
  double acc = 0.0;
  run_tp = tp0;
  for (int i = i0; i < i0 + per && i < n; ++i) {
    run_tp += lab[i];
    if (i == n - 1 || keys[i] != keys[i + 1]) {
      if (run_tp != prev_tp && total_pos)
        acc += ((double)(run_tp - prev_tp) / (double)total_pos) * ((double)run_tp / (double)(i + 1));
      prev_tp = run_tp;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kApThreads / 32; ++w) t += red[w];
    ap[b] = t;
  }
}

}  // namespace r2n2

using namespace r2n2;

// ====================================================================== C ABI
extern "C" {

#ifndef R2N2_SRC_HASH
#define R2N2_SRC_HASH "unhashed"
#endif
const char* r2n2_version(void) { return "r2n2_b200 0.2 (sm_100a) src " R2N2_SRC_HASH; }
const char* r2n2_last_error(void) { return r2n2::g_err; }

int r2n2_fill_zero(void* p, long long bytes, void* stream) {
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
  R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "fill_zero: %s", cudaGetErrorString(e));
  return R2N2_OK;
}

int r2n2_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int Cpad, int Wpad,
                      int xoff, int out_dtype, void* stream) {
  R2N2_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C && Wpad >= W + xoff && xoff >= 0,
                 R2N2_ERR_BAD_SHAPE, "nchw_to_nhwc: bad shape");
  long long total = (long long)N * H * Wpad;
  R2N2_DISPATCH_DTYPE(out_dtype, T, {
    r2n2::launch_k((nchw_to_nhwc_kernel<T>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        x, (T*)y, N, C, H, W, Cpad, Wpad, xoff);
  })
  R2N2_CHECK_LAUNCH("nchw_to_nhwc");
  return R2N2_OK;
}

int r2n2_nchw_to_rowpack(const float* x, void* y, int N, int C, int H, int W, int kw, int Cpack,
                         int out_dtype, void* stream) {
  R2N2_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && (kw & 1) && kw * C <= Cpack && (Cpack % 8) == 0,
                 R2N2_ERR_BAD_SHAPE, "nchw_to_rowpack: bad shape");
  R2N2_DISPATCH_DTYPE(out_dtype, T, {
    const size_t smem = sizeof(float) * (size_t)C * (W + kw - 1) + sizeof(int) * (size_t)Cpack;
    long long blocks = (long long)N * H;
    if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
    r2n2::launch_k((nchw_to_rowpack_kernel<T>), dim3((unsigned)blocks), dim3(128), smem, (cudaStream_t)((cudaStream_t)stream), 0,
        x, (T*)y, N, C, H, W, kw, Cpack);
  })
  R2N2_CHECK_LAUNCH("nchw_to_rowpack");
  return R2N2_OK;
}

int r2n2_maxpool2d_fwd(const void* a, const void* b, void* y, int N, int H, int W, int C, int dtype,
                       void* stream) {
  R2N2_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && (C % (dtype == R2N2_BF16 ? 8 : 4)) == 0,
                 R2N2_ERR_BAD_SHAPE, "maxpool2d_fwd: C must be a multiple of 4 (fp32) / 8 (bf16), got %d", C);
  int Ho = H / 2 + 1, Wo = W / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / (dtype == R2N2_BF16 ? 8 : 4));
  R2N2_DISPATCH_DTYPE(dtype, T, {
    if (b) {
      r2n2::launch_k((maxpool_fwd_kernel<T, true>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)a, (const T*)b, (T*)y, N, H, W, C, Ho, Wo, (unsigned char*)nullptr);
    } else {
      r2n2::launch_k((maxpool_fwd_kernel<T, false>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)a, (const T*)b, (T*)y, N, H, W, C, Ho, Wo, (unsigned char*)nullptr);
    }
  })
  R2N2_CHECK_LAUNCH("maxpool2d_fwd");
  return R2N2_OK;
}

int r2n2_maxpool2d_fwd_code(const void* a, const void* b, void* y, unsigned char* code, int N, int H, int W, int C,
                            int dtype, void* stream) {
  R2N2_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && (C % (dtype == R2N2_BF16 ? 8 : 4)) == 0,
                 R2N2_ERR_BAD_SHAPE, "maxpool2d_fwd_code: C must be a multiple of 4 (fp32) / 8 (bf16), got %d", C);
  R2N2_CHECK_ARG(code != nullptr && (reinterpret_cast<uintptr_t>(code) & 7) == 0, R2N2_ERR_BAD_ALIGN,
                 "maxpool2d_fwd_code: code must be 8-byte aligned");
  int Ho = H / 2 + 1, Wo = W / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / (dtype == R2N2_BF16 ? 8 : 4));
  R2N2_DISPATCH_DTYPE(dtype, T, {
    if (b) {
      r2n2::launch_k((maxpool_fwd_kernel<T, true, true>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)a, (const T*)b, (T*)y, N, H, W, C, Ho, Wo, code);
    } else {
      r2n2::launch_k((maxpool_fwd_kernel<T, false, true>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)a, (const T*)b, (T*)y, N, H, W, C, Ho, Wo, code);
    }
  })
  R2N2_CHECK_LAUNCH("maxpool2d_fwd_code");
  return R2N2_OK;
}

int r2n2_maxpool2d_bwd_code(const unsigned char* code, const void* dy, void* dx, void* dxb, float* colsum_a,
                            float* colsum_b, int N, int H, int W, int C, int lrelu_mask, int dtype, void* stream) {
  const int pn = dtype == R2N2_BF16 ? 8 : 4;
  R2N2_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && (C % pn) == 0,
                 R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd_code: C must be a multiple of 4 (fp32) / 8 (bf16), got %d", C);
  R2N2_CHECK_ARG(code && dy && dx, R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd_code: null pointer");
  R2N2_CHECK_ARG(!(lrelu_mask && dxb), R2N2_ERR_UNSUPPORTED,
                 "maxpool2d_bwd_code: the code holds ONE sign per position (a's without b, b's with b)");
  R2N2_CHECK_ARG(dxb || !colsum_b, R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd_code: colsum_b is the column sum of dxb");
  int Ho = H / 2 + 1, Wo = W / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / pn);
  const bool extra = colsum_a || colsum_b;
  unsigned grid = grid_for(total, 256);
  if (extra) {
    R2N2_CHECK_ARG(C <= 512, R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd_code: fused column sums support C <= 512, got %d", C);
    int cg = C / pn, g = 256, r = cg;
    while (r) { int t = g % r; g = r; r = t; }                 // g = gcd(256, cg)
    const unsigned m = (unsigned)(cg / g);
    const unsigned wave = (unsigned)num_sms() * 3u;            // one resident wave: see r2n2_maxpool2d_bwd_fused
    if (grid > wave) grid = wave;
    grid = (grid + m - 1) / m * m;
  }
  R2N2_DISPATCH_DTYPE(dtype, T, {
    if (extra)
      r2n2::launch_k((maxpool_bwd_code_kernel<T, true>), dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        code, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, lrelu_mask, (T*)dxb, colsum_a, colsum_b);
    else
      r2n2::launch_k((maxpool_bwd_code_kernel<T, false>), dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        code, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, lrelu_mask, (T*)dxb, nullptr, nullptr);
  })
  R2N2_CHECK_LAUNCH("maxpool2d_bwd_code");
  return R2N2_OK;
}

int r2n2_maxpool2d_bwd_fused(const void* a, const void* b, const void* dy, void* dx, void* dxb, float* colsum_a,
                             float* colsum_b, int N, int H, int W, int C, int lrelu_mask, int dtype, void* stream) {
  const int pn = dtype == R2N2_BF16 ? 8 : 4;
  R2N2_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && (C % pn) == 0,
                 R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd: C must be a multiple of 4 (fp32) / 8 (bf16), got %d", C);
  R2N2_CHECK_ARG(b || (!dxb && !colsum_b), R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd: dxb / colsum_b need the branch b");
  R2N2_CHECK_ARG(dxb || !colsum_b, R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd: colsum_b is the column sum of dxb");
  int Ho = H / 2 + 1, Wo = W / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C / pn);
  const bool extra = dxb || colsum_a || colsum_b;
  unsigned grid = grid_for(total, 256);
  if (extra) {
    R2N2_CHECK_ARG(C <= 512, R2N2_ERR_BAD_SHAPE, "maxpool2d_bwd: fused column sums support C <= 512, got %d", C);
    // every thread must keep its channel group across the grid-stride loop: grid * 256 % (C / pn) == 0
    int cg = C / pn, g = 256, r = cg;
    while (r) { int t = g % r; g = r; r = t; }                 // g = gcd(256, cg)
    const unsigned m = (unsigned)(cg / g);
    // one wave of resident blocks (grid-stride): every block ends with C atomic adds per sum on the SAME C
    // addresses, so 16 blocks per SM would serialise ~2400 atomics per address at L2
    const unsigned wave = (unsigned)num_sms() * (b ? 2u : 3u);
    if (grid > wave) grid = wave;
    grid = (grid + m - 1) / m * m;
  }
  R2N2_DISPATCH_DTYPE(dtype, T, {
    if (b) {
      if (extra)
        r2n2::launch_k((maxpool_bwd_kernel<T, true, true>), dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
          (const T*)a, (const T*)b, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, lrelu_mask, (T*)dxb, colsum_a, colsum_b);
      else
        r2n2::launch_k((maxpool_bwd_kernel<T, true, false>), dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
          (const T*)a, (const T*)b, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, lrelu_mask, nullptr, nullptr, nullptr);
    } else {
      if (extra)
        r2n2::launch_k((maxpool_bwd_kernel<T, false, true>), dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
          (const T*)a, (const T*)b, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, lrelu_mask, nullptr, colsum_a, nullptr);
      else
        r2n2::launch_k((maxpool_bwd_kernel<T, false, false>), dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
          (const T*)a, (const T*)b, (const T*)dy, (T*)dx, N, H, W, C, Ho, Wo, lrelu_mask, nullptr, nullptr, nullptr);
    }
  })
  R2N2_CHECK_LAUNCH("maxpool2d_bwd");
  return R2N2_OK;
}

int r2n2_maxpool2d_bwd(const void* a, const void* b, const void* dy, void* dx, int N, int H, int W,
                       int C, int lrelu_mask, int dtype, void* stream) {
  return r2n2_maxpool2d_bwd_fused(a, b, dy, dx, nullptr, nullptr, nullptr, N, H, W, C, lrelu_mask, dtype, stream);
}

int r2n2_unpool3d_fwd(const void* a, const void* b, void* y, int B, int D, int H, int W, int C,
                      int dtype, void* stream) {
  R2N2_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && (C % (dtype == R2N2_BF16 ? 8 : 4)) == 0,
                 R2N2_ERR_BAD_SHAPE, "unpool3d_fwd: bad shape");
  long long total = (long long)B * D * H * W * 8 * (C / (dtype == R2N2_BF16 ? 8 : 4));
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((unpool_fwd_kernel<T>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)a, (const T*)b, (T*)y, B, D, H, W, C);
  })
  R2N2_CHECK_LAUNCH("unpool3d_fwd");
  return R2N2_OK;
}

int r2n2_unpool3d_bwd(const void* dy, void* dx, int B, int D, int H, int W, int C, int dtype,
                      void* stream) {
  R2N2_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && (C % (dtype == R2N2_BF16 ? 8 : 4)) == 0,
                 R2N2_ERR_BAD_SHAPE, "unpool3d_bwd: bad shape");
  long long total = (long long)B * D * H * W * (C / (dtype == R2N2_BF16 ? 8 : 4));
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((unpool_bwd_kernel<T>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)dy, (T*)dx, B, D, H, W, C);
  })
  R2N2_CHECK_LAUNCH("unpool3d_bwd");
  return R2N2_OK;
}

int r2n2_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream) {
  R2N2_CHECK_ARG(n > 0 && (n & 3) == 0, R2N2_ERR_BAD_SHAPE, "add: n must be a multiple of 4");
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((add_kernel<T>), dim3(grid_for(n >> 2, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0, (const T*)a, (const T*)b,
                                                                           (T*)y, n >> 2);
  })
  R2N2_CHECK_LAUNCH("add");
  return R2N2_OK;
}

int r2n2_lrelu_fwd(const void* x, void* y, long long n, int dtype, void* stream) {
  R2N2_CHECK_ARG(n > 0 && (n & 3) == 0, R2N2_ERR_BAD_SHAPE, "lrelu_fwd: n must be a multiple of 4");
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((lrelu_fwd_kernel<T>), dim3(grid_for(n >> 2, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0, (const T*)x, (T*)y,
                                                                                 n >> 2);
  })
  R2N2_CHECK_LAUNCH("lrelu_fwd");
  return R2N2_OK;
}

int r2n2_lrelu_bwd(const void* a, const void* dy, const void* addend, void* dx, long long n,
                   int dtype, void* stream) {
  R2N2_CHECK_ARG(n > 0 && (n & 3) == 0, R2N2_ERR_BAD_SHAPE, "lrelu_bwd: n must be a multiple of 4");
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((lrelu_bwd_kernel<T>), dim3(grid_for(n >> 2, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)a, (const T*)dy, (const T*)addend, (T*)dx, n >> 2);
  })
  R2N2_CHECK_LAUNCH("lrelu_bwd");
  return R2N2_OK;
}

int r2n2_lrelu_bwd_colsum(const void* a, const void* dy, const void* addend, void* dx, long long rows, int C,
                          float* colsum, int dtype, void* stream) {
  const int pn = dtype == R2N2_BF16 ? 8 : 4;
  R2N2_CHECK_ARG(rows > 0 && C > 0 && C % pn == 0 && C <= 512, R2N2_ERR_BAD_SHAPE,
                 "lrelu_bwd_colsum: C (%d) must be a multiple of %d and <= 512", C, pn);
  R2N2_CHECK_ARG(a && dy && dx && colsum, R2N2_ERR_BAD_SHAPE, "lrelu_bwd_colsum: null pointer");
  const long long nvec = rows * (C / pn);
  unsigned grid = grid_for(nvec, 256);
  {
    int cg = C / pn, g = 256, r = cg;
    while (r) { int t = g % r; g = r; r = t; }                 // g = gcd(256, cg)
    const unsigned m = (unsigned)(cg / g);
    const unsigned wave = (unsigned)num_sms() * 4u;
    if (grid > wave) grid = wave;
    grid = (grid + m - 1) / m * m;
  }
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((lrelu_bwd_colsum_kernel<T>), dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)a, (const T*)dy, (const T*)addend, (T*)dx, nvec, C, colsum);
  })
  R2N2_CHECK_LAUNCH("lrelu_bwd_colsum");
  return R2N2_OK;
}

int r2n2_eltwise(int op, const void* a, const void* b, void* y, long long n, int dtype,
                 void* stream) {
  R2N2_CHECK_ARG(n > 0 && (n & 3) == 0, R2N2_ERR_BAD_SHAPE, "eltwise: n must be a multiple of 4");
  R2N2_CHECK_ARG(op >= R2N2_ELT_SIGMOID && op <= R2N2_ELT_NEG, R2N2_ERR_UNSUPPORTED, "eltwise: bad op %d", op);
  R2N2_CHECK_ARG(!(op == R2N2_ELT_MUL || op == R2N2_ELT_SIGMOID_BWD || op == R2N2_ELT_TANH_BWD) || b,
                 R2N2_ERR_BAD_SHAPE, "eltwise: op %d needs two operands", op);
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((eltwise_kernel<T>), dim3(grid_for(n >> 2, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        op, (const T*)a, (const T*)b, (T*)y, n >> 2);
  })
  R2N2_CHECK_LAUNCH("eltwise");
  return R2N2_OK;
}

int r2n2_cast(const float* src, void* dst, long long n, int out_dtype, void* stream) {
  R2N2_CHECK_ARG(n > 0, R2N2_ERR_BAD_SHAPE, "cast: n <= 0");
  R2N2_DISPATCH_DTYPE(out_dtype, T, {
    r2n2::launch_k((cast_kernel<T>), dim3(grid_for((n >> 2) + 1, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0, src, (T*)dst, n);
  })
  R2N2_CHECK_LAUNCH("cast");
  return R2N2_OK;
}

int r2n2_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream) {
  R2N2_CHECK_ARG(n > 0 && src && hi && lo, R2N2_ERR_BAD_SHAPE, "split_bf16: bad arguments");
  R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(lo) & 15) == 0,
                 R2N2_ERR_BAD_ALIGN, "split_bf16: pointers must be 16-byte aligned");
  r2n2::launch_k((split_bf16_kernel), dim3(grid_for((n >> 3) + 1, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
      src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  R2N2_CHECK_LAUNCH("split_bf16");
  return R2N2_OK;
}

int r2n2_gru_gates_ur_fwd(const void* conv_ur, const void* xp_ur, const float* bias_ur,
                          const void* h, void* u, void* r, void* rh, int B, int dtype,
                          void* stream) {
  R2N2_CHECK_ARG(B > 0, R2N2_ERR_BAD_SHAPE, "gru_gates_ur_fwd: B <= 0");
  long long rows = (long long)B * 64;
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((gru_ur_fwd_kernel<T>), dim3(grid_for(rows * 32, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)conv_ur, (const T*)xp_ur, bias_ur, (const T*)h, (T*)u, (T*)r, (T*)rh, rows);
  })
  R2N2_CHECK_LAUNCH("gru_gates_ur_fwd");
  return R2N2_OK;
}

int r2n2_gru_gates_out_fwd(const void* conv_c, const void* xp_c, const float* bias_c, const void* u,
                           const void* h, void* c, void* h_new, int B, int dtype, void* stream) {
  R2N2_CHECK_ARG(B > 0, R2N2_ERR_BAD_SHAPE, "gru_gates_out_fwd: B <= 0");
  long long rows = (long long)B * 64;
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((gru_out_fwd_kernel<T>), dim3(grid_for(rows * 32, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)conv_c, (const T*)xp_c, bias_c, (const T*)u, (const T*)h, (T*)c, (T*)h_new, rows);
  })
  R2N2_CHECK_LAUNCH("gru_gates_out_fwd");
  return R2N2_OK;
}

int r2n2_gru_gates_out_bwd(const void* dh, const void* u, const void* c, const void* h, void* da_ur,
                           void* da_c, void* dh_prev, int B, int dtype, void* stream) {
  R2N2_CHECK_ARG(B > 0, R2N2_ERR_BAD_SHAPE, "gru_gates_out_bwd: B <= 0");
  long long rows = (long long)B * 64;
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((gru_out_bwd_kernel<T>), dim3(grid_for(rows * 32, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)dh, (const T*)u, (const T*)c, (const T*)h, (T*)da_ur, (T*)da_c, (T*)dh_prev, rows);
  })
  R2N2_CHECK_LAUNCH("gru_gates_out_bwd");
  return R2N2_OK;
}

int r2n2_gru_gates_r_bwd(const void* d_rh, const void* h, const void* r, void* da_ur,
                         void* dh_prev, int B, int dtype, void* stream) {
  R2N2_CHECK_ARG(B > 0, R2N2_ERR_BAD_SHAPE, "gru_gates_r_bwd: B <= 0");
  long long rows = (long long)B * 64;
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((gru_r_bwd_kernel<T>), dim3(grid_for(rows * 32, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)d_rh, (const T*)h, (const T*)r, (T*)da_ur, (T*)dh_prev, rows);
  })
  R2N2_CHECK_LAUNCH("gru_gates_r_bwd");
  return R2N2_OK;
}

int r2n2_lstm_gates_fwd(const void* conv_fig, const void* xp_fig, const float* bias_fig,
                        const void* s, void* fig, void* s_new, void* h_new, int B, int dtype,
                        void* stream) {
  R2N2_CHECK_ARG(B > 0, R2N2_ERR_BAD_SHAPE, "lstm_gates_fwd: B <= 0");
  long long rows = (long long)B * 64;
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((lstm_fwd_kernel<T>), dim3(grid_for(rows * 32, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)conv_fig, (const T*)xp_fig, bias_fig, (const T*)s, (T*)fig, (T*)s_new, (T*)h_new,
        rows);
  })
  R2N2_CHECK_LAUNCH("lstm_gates_fwd");
  return R2N2_OK;
}

int r2n2_lstm_gates_bwd(const void* dh, const void* ds_in, const void* fig, const void* s,
                        const void* h_new, void* da_fig, void* ds_prev, int B, int dtype,
                        void* stream) {
  R2N2_CHECK_ARG(B > 0, R2N2_ERR_BAD_SHAPE, "lstm_gates_bwd: B <= 0");
  long long rows = (long long)B * 64;
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((lstm_bwd_kernel<T>), dim3(grid_for(rows * 32, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)dh, (const T*)ds_in, (const T*)fig, (const T*)s, (const T*)h_new, (T*)da_fig,
        (T*)ds_prev, rows);
  })
  R2N2_CHECK_LAUNCH("lstm_gates_bwd");
  return R2N2_OK;
}

int r2n2_softmax_ce3d(const float* logits, const float* y, float* pred, float* loss_sum,
                      void* dlogits, int B, int nv, float grad_scale, int ddtype, int cstride,
                      void* stream) {
  R2N2_CHECK_ARG(B > 0 && nv > 0 && cstride >= 2 && (cstride & 1) == 0, R2N2_ERR_BAD_SHAPE,
                 "softmax_ce3d: bad shape");
  long long total = (long long)B * nv * nv * nv;
  R2N2_DISPATCH_DTYPE(ddtype, T, {
    r2n2::launch_k((softmax_ce3d_kernel<T>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        logits, y, pred, loss_sum, (T*)dlogits, total, nv, grad_scale, cstride);
  })
  R2N2_CHECK_LAUNCH("softmax_ce3d");
  return R2N2_OK;
}

int r2n2_bias_grad(const void* dy, float* db, long long rows, int C, int dtype, int accumulate,
                   void* stream) {
  R2N2_CHECK_ARG(rows > 0 && C > 0, R2N2_ERR_BAD_SHAPE, "bias_grad: bad shape");
  if (!accumulate) cudaMemsetAsync(db, 0, sizeof(float) * C, (cudaStream_t)stream);
  const int pn = (dtype == R2N2_BF16) ? 8 : 4;
  if (C % pn == 0 && C / pn <= 256 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    const int rpi = 256 / (C / pn);
    long long blocks = (rows + 8LL * rpi - 1) / (8LL * rpi);       // >= 8 iterations per block
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    if (blocks < 1) blocks = 1;
    const long long rpb = (rows + blocks - 1) / blocks;
    blocks = (rows + rpb - 1) / rpb;
    R2N2_DISPATCH_DTYPE(dtype, T, {
      r2n2::launch_k((bias_grad_flat_kernel<T>), dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0, (const T*)dy, db, rows, C, rpb);
    })
    R2N2_CHECK_LAUNCH("bias_grad");
    return R2N2_OK;
  }
  int gx = (C + 32 * pn - 1) / (32 * pn);
  long long want = ((long long)num_sms() * 8 + gx - 1) / gx;
  long long gy = rows / 64;
  if (gy < 1) gy = 1;
  if (gy > want) gy = want;
  long long rpb = (rows + gy - 1) / gy;
  gy = (rows + rpb - 1) / rpb;
  R2N2_DISPATCH_DTYPE(dtype, T, {
    r2n2::launch_k((bias_grad_kernel<T>), dim3(dim3(gx, (unsigned)gy)), dim3(dim3(32, 8)), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        (const T*)dy, db, rows, C, rpb);
  })
  R2N2_CHECK_LAUNCH("bias_grad");
  return R2N2_OK;
}

int r2n2_weight_transpose(const float* w, void* wt, int Cout, int taps, int Cin, int out_dtype,
                          void* stream) {
  R2N2_CHECK_ARG(Cout > 0 && taps > 0 && Cin > 0 && taps <= 65535, R2N2_ERR_BAD_SHAPE,
                 "weight_transpose: bad shape");
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, taps);
  R2N2_CHECK_ARG(grid.y <= 65535, R2N2_ERR_BAD_SHAPE, "weight_transpose: Cout too large");
  R2N2_DISPATCH_DTYPE(out_dtype, T, {
    r2n2::launch_k((weight_transpose_kernel<T>), dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)((cudaStream_t)stream), 0, w, (T*)wt, Cout,
                                                                                taps, Cin);
  })
  R2N2_CHECK_LAUNCH("weight_transpose");
  return R2N2_OK;
}

int r2n2_weight_transpose_batch(const float* base, void* wt_base, const long long* table, int n_entries,
                                long long total_tiles, int out_dtype, void* stream) {
  R2N2_CHECK_ARG(base && wt_base && table && n_entries > 0 && total_tiles > 0 && total_tiles < (1LL << 31),
                 R2N2_ERR_BAD_SHAPE, "weight_transpose_batch: bad arguments");
  R2N2_DISPATCH_DTYPE(out_dtype, T, {
    r2n2::launch_k((weight_transpose_batch_kernel<T>), dim3((unsigned)total_tiles), dim3(dim3(32, 8)), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        base, (T*)wt_base, table, n_entries);
  })
  R2N2_CHECK_LAUNCH("weight_transpose_batch");
  return R2N2_OK;
}

int r2n2_weight_transpose_batch_bf16(const void* mirror_base, void* wt_base, const long long* table, int n_entries,
                                     long long total_tiles, void* stream) {
  R2N2_CHECK_ARG(mirror_base && wt_base && table && n_entries > 0 && total_tiles > 0 && total_tiles < (1LL << 31),
                 R2N2_ERR_BAD_SHAPE, "weight_transpose_batch_bf16: bad arguments");
  R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(mirror_base) & 3) == 0 && (reinterpret_cast<uintptr_t>(wt_base) & 3) == 0,
                 R2N2_ERR_BAD_ALIGN, "weight_transpose_batch_bf16: buffers must be 4-byte aligned");
  r2n2::launch_k((weight_transpose_batch_bf16_kernel), dim3((unsigned)total_tiles), dim3(dim3(32, 8)), 0, (cudaStream_t)((cudaStream_t)stream), 0,
      (const __nv_bfloat16*)mirror_base, (__nv_bfloat16*)wt_base, table, n_entries);
  R2N2_CHECK_LAUNCH("weight_transpose_batch_bf16");
  return R2N2_OK;
}

int r2n2_adam(float* p, const float* g, float* m, float* v, void* p_mirror, long long n,
              long long n_decay, float lr_t, float beta1, float beta2, float eps, float wd,
              float grad_scale, int mirror_dtype, void* stream) {
  R2N2_CHECK_ARG(n > 0 && n_decay >= 0 && n_decay <= n, R2N2_ERR_BAD_SHAPE, "adam: bad sizes");
  int grid = grid_for((n + 3) >> 2, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (!p_mirror) {
    r2n2::launch_k((adam_kernel<float, false>), dim3(grid), dim3(256), 0, (cudaStream_t)(st), 0, p, g, m, v, nullptr, n, n_decay, lr_t, beta1,
                                                    beta2, eps, wd, grad_scale);
  } else if (mirror_dtype == R2N2_BF16) {
    r2n2::launch_k((adam_kernel<__nv_bfloat16, true>), dim3(grid), dim3(256), 0, (cudaStream_t)(st), 0, p, g, m, v, (__nv_bfloat16*)p_mirror, n,
                                                           n_decay, lr_t, beta1, beta2, eps, wd,
                                                           grad_scale);
  } else {
    R2N2_CHECK_ARG(false, R2N2_ERR_BAD_DTYPE, "adam: mirror dtype must be bf16");
  }
  R2N2_CHECK_LAUNCH("adam");
  return R2N2_OK;
}

int r2n2_sgd(float* p, const float* g, float* vel, void* p_mirror, long long n, long long n_decay,
             float lr, float momentum, float wd, float grad_scale, int mirror_dtype, void* stream) {
  R2N2_CHECK_ARG(n > 0 && n_decay >= 0 && n_decay <= n, R2N2_ERR_BAD_SHAPE, "sgd: bad sizes");
  int grid = grid_for(n, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (!p_mirror) {
    r2n2::launch_k((sgd_kernel<float, false>), dim3(grid), dim3(256), 0, (cudaStream_t)(st), 0, p, g, vel, nullptr, n, n_decay, lr, momentum, wd,
                                                   grad_scale);
  } else if (mirror_dtype == R2N2_BF16) {
    r2n2::launch_k((sgd_kernel<__nv_bfloat16, true>), dim3(grid), dim3(256), 0, (cudaStream_t)(st), 0, p, g, vel, (__nv_bfloat16*)p_mirror, n,
                                                          n_decay, lr, momentum, wd, grad_scale);
  } else {
    R2N2_CHECK_ARG(false, R2N2_ERR_BAD_DTYPE, "sgd: mirror dtype must be bf16");
  }
  R2N2_CHECK_LAUNCH("sgd");
  return R2N2_OK;
}

int r2n2_voxel_eval(const float* pred, const float* gt, float thresh, long long* counts, int B,
                    int nv, void* stream) {
  R2N2_CHECK_ARG(B > 0 && nv > 0, R2N2_ERR_BAD_SHAPE, "voxel_eval: bad shape");
  cudaMemsetAsync(counts, 0, sizeof(long long) * 5 * B, (cudaStream_t)stream);
  long long nvox = (long long)nv * nv * nv;
  int gx = (int)((nvox + 255) / 256);
  if (gx > 64) gx = 64;
  r2n2::launch_k((voxel_eval_kernel), dim3(dim3(gx, B)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
      pred, gt, thresh, (unsigned long long*)counts, nv);
  R2N2_CHECK_LAUNCH("voxel_eval");
  return R2N2_OK;
}

int r2n2_preprocess_views(const unsigned char* rgba, const int* params, void* out, int n_img, int Hs, int Ws,
                          int Ho, int Wo, int out_dtype, void* stream) {
  R2N2_CHECK_ARG(rgba && params && out, R2N2_ERR_BAD_SHAPE, "preprocess_views: null pointer");
  R2N2_CHECK_ARG(n_img > 0 && Ho > 0 && Wo > 0 && Ho <= Hs && Wo <= Ws, R2N2_ERR_BAD_SHAPE,
                 "preprocess_views: bad shape (output %dx%d from %dx%d)", Ho, Wo, Hs, Ws);
  R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(rgba) & 3) == 0, R2N2_ERR_BAD_ALIGN, "preprocess_views: rgba must be 4-byte aligned");
  const long long total = (long long)n_img * Ho * Wo;
  R2N2_DISPATCH_DTYPE(out_dtype, T, {
    r2n2::launch_k((preprocess_views_kernel<T>), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0,
        reinterpret_cast<const uchar4*>(rgba), params, (T*)out, n_img, Hs, Ws, Ho, Wo);
  })
  R2N2_CHECK_LAUNCH("preprocess_views");
  return R2N2_OK;
}

int r2n2_voxel_labels(const unsigned char* occ, float* y, int B, int nv, void* stream) {
  R2N2_CHECK_ARG(occ && y, R2N2_ERR_BAD_SHAPE, "voxel_labels: null pointer");
  R2N2_CHECK_ARG(B > 0 && nv > 0, R2N2_ERR_BAD_SHAPE, "voxel_labels: bad shape B=%d nv=%d", B, nv);
  const long long total = (long long)B * nv * nv * nv;
  r2n2::launch_k((voxel_labels_kernel), dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), 0, occ, y, nv, total);
  R2N2_CHECK_LAUNCH("voxel_labels");
  return R2N2_OK;
}

int r2n2_average_precision(const float* pred, const float* gt, double* ap, int B, int nv, void* stream) {
  R2N2_CHECK_ARG(B > 0 && nv > 0 && (long long)nv * nv * nv <= kApMax, R2N2_ERR_BAD_SHAPE,
                 "average_precision: need 0 < nv^3 <= %d", kApMax);
  int npad = 1;
  while (npad < nv * nv * nv) npad <<= 1;
  if (npad < kApThreads) npad = kApThreads;
  const size_t smem = (size_t)npad * 5;                     // 4-byte keys + 1-byte labels
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(average_precision_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kApMax * 5);
    attr_set = true;
  }
  r2n2::launch_k((average_precision_kernel), dim3(B), dim3(kApThreads), smem, (cudaStream_t)((cudaStream_t)stream), 0, pred, gt, ap, nv);
  R2N2_CHECK_LAUNCH("average_precision");
  return R2N2_OK;
}

}  // extern "C"
