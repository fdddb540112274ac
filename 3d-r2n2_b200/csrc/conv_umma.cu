// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate).
//
//   y[p, co] = epilogue( sum_{tap, ci} x[p + off(tap), ci] * w[co, tap, ci] )
//
// GEMM view: M = output pixels, N = Cout, K = taps * Cin.  One CTA computes a 128 x BN tile:
//   * the M tile is a SPATIAL BOX (bw x bh x bd x bn pixels, <= 128) of the channels-last
//     activation tensor; for every filter tap the A operand is ONE tiled-TMA box load of the
//     shifted box -- TMA's out-of-bounds zero fill implements the 'same' zero padding, so no
//     im2col buffer and no padded copy ever exists (the reference materialises both,
//     lib/layers.py:304-313, 427-437);
//   * the B operand (packed weights [Cout][taps*Cin], K contiguous) is a 2-D TMA box;
//   * both land in shared memory in the canonical K-major swizzled layout that tcgen05.mma
//     reads through shared-memory descriptors; accumulators live in TMEM;
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-5 = epilogue
//     (tcgen05.ld -> bias / LeakyReLU / residual / LeakyReLU' mask -> global), connected by an
//     mbarrier ring (full/empty per stage) and a tmem_full barrier.
// The same kernel computes data gradients (dgrad) when fed dy and the transposed weights.
#include "umma.cuh"

namespace r2n2 {

struct UmmaConvParams {
  int N, D, H, W, Cin, Cout;
  int kd, kh, kw;
  int bw, bh, bd, bn;                      // spatial box of the M tile
  int tiles_w, tiles_h, tiles_d, tiles_n;  // number of boxes per axis
  int BN;                                  // N tile (multiple of 16, <= 256)
  int kbox;                                // channels per K block (16 / 32 / 64)
  int kchunks;                             // ceil(Cin / kbox)
  int stages;
  int a_stage_bytes, b_stage_bytes;        // 1024-aligned
  int tmem_cols;
  int flags;
  int* err_word;
  long long* dbg;                          // optional [grid][8] clock64 timestamps (R2N2_UMMA_DBG=1)
};

template <typename TOut>
__global__ void __launch_bounds__(kUmmaThreads)
conv_fwd_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const float* __restrict__ bias, const TOut* mask, const TOut* res, TOut* y,
                     const UmmaConvParams p) {
  r2n2::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] [full[stages]] [empty[stages]] [tmem_full] [tmem_ptr] [abort]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  R2N2_WS_BEGIN(bars);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full_bar = bars + 2 * p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* dbg = p.dbg ? p.dbg + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  long long* tl = (p.dbg && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0)
                      ? p.dbg + (long long)gridDim.x * gridDim.y * 8 : nullptr;   // [kb][4] timeline
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();

  // tile coordinates
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h; t /= p.tiles_h;
  const int td = t % p.tiles_d; t /= p.tiles_d;
  const int tn = t;
  const int x0 = tw * p.bw, y0 = th * p.bh, z0 = td * p.bd, n0 = tn * p.bn;
  const int col0 = blockIdx.y * p.BN;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  r2n2::pdl_wait();            // prologue done; from here on global memory of the preceding kernels may be touched
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  const int taps = p.kd * p.kh * p.kw;
  const int nkb = taps * p.kchunks;
  (void)nkb;
  const int row_bytes = p.kbox * 2;
  const int rows = p.bw * p.bh * p.bd * p.bn;

  if (warp == 0) {
    // ================================ TMA producer =========================================
    // A lone thread runs this loop: every instruction's latency is exposed, so the loop carries
    // only counters (no divisions): nested (tz, ty, tx, chunk) iteration, incremental stage/phase.
    {
      const uint32_t tx_bytes = (uint32_t)(rows * row_bytes + p.BN * row_bytes);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const int cz = z0 - (p.kd - 1) / 2, cy = y0 - (p.kh - 1) / 2, cx = x0 - (p.kw - 1) / 2;
      int s = 0;
      uint32_t ph = 0;
      int kbi = 0;
      int kcol = 0;                                   // tap * Cin + c0 (weights K coordinate)
      for (int tz = 0; tz < p.kd; ++tz)
        for (int ty = 0; ty < p.kh; ++ty)
          for (int tx = 0; tx < p.kw; ++tx) {
            for (int c0 = 0; c0 < p.Cin; c0 += p.kbox) {
              mbar_wait(empty0 + 8u * s, ph ^ 1u, abort_flag, p.err_word);
              if (tl) { tl[kbi * 4 + 0] = clock64(); }
              const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes);
              const uint32_t bar = full0 + 8u * s;
              const int skip = (p.flags >> 16) & 3;        // perf experiment: 1 = no A loads, 2 = no B loads
              if (elect_one()) {
                mbar_expect_tx(bar, skip == 1 ? (uint32_t)(p.BN * row_bytes)
                                              : (skip == 2 ? (uint32_t)(rows * row_bytes) : tx_bytes));
                if (skip != 1) tma_load_5d(a_dst, &map_a, bar, c0, cx + tx, cy + ty, cz + tz, n0);
                if (skip != 2)
                  tma_load_2d(a_dst + (uint32_t)p.a_stage_bytes, &map_b, bar, kcol + c0, col0);
              }
              __syncwarp();
              if (tl) { tl[kbi * 4 + 1] = clock64(); ++kbi; }
              if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            kcol += p.Cin;
          }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ===========================================
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t layout_type = (row_bytes == 128) ? 2u : (row_bytes == 64 ? 4u : 6u);
      const uint32_t sbo = 8u * (uint32_t)row_bytes;
      // descriptor templates: only the 14-bit start-address field changes per MMA
      const uint64_t desc_hi = make_kmajor_desc(0, sbo, layout_type);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const int last_nk16 = (p.Cin - (p.kchunks - 1) * p.kbox) >> 4;
      const int full_nk16 = p.kbox >> 4;
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int tap = 0; tap < taps; ++tap)
        for (int cc = 0; cc < p.kchunks; ++cc) {
          mbar_wait(full0 + 8u * s, ph, abort_flag, p.err_word);
          tc_fence_after();
          if (dbg && acc == 0) dbg[2] = clock64();
          if (tl) tl[(tap * p.kchunks + cc) * 4 + 2] = clock64();
          const int nk16 = (cc == p.kchunks - 1) ? last_nk16 : full_nk16;
          const uint32_t a_addr = smem_base + (uint32_t)(s * stage_bytes);
          const uint64_t da0 = desc_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
          const uint64_t db0 = desc_hi | (uint64_t)(((a_addr + (uint32_t)p.a_stage_bytes) & 0x3FFFF) >> 4);
          if (elect_one()) {
            uint32_t a2 = acc;
            for (int k = 0; k < nk16; ++k) {
              umma_bf16(tmem_base, da0 + 2u * k, db0 + 2u * k, idesc, a2);
              a2 = 1;
            }
            umma_commit(empty0 + 8u * s);            // frees the smem stage when the MMAs retire
          }
          __syncwarp();
          acc = 1;
          if (tl) tl[(tap * p.kchunks + cc) * 4 + 3] = clock64();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      if (dbg) dbg[3] = clock64();
      if (elect_one()) umma_commit(smem_u32(tmem_full_bar));   // accumulator complete
    }
    __syncwarp();
  } else {
    // ================================ epilogue (warps 2..5) =================================
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    int r = row;
    const int xi = r % p.bw; r /= p.bw;
    const int yi = r % p.bh; r /= p.bh;
    const int zi = r % p.bd; r /= p.bd;
    const int ni = r;
    const int xx = x0 + xi, yy = y0 + yi, zz = z0 + zi, nn = n0 + ni;
    const bool valid = (row < rows) && xx < p.W && yy < p.H && zz < p.D && nn < p.N;
    const long long pix = (((long long)nn * p.D + zz) * p.H + yy) * p.W + xx;
    mbar_wait(smem_u32(tmem_full_bar), 0u, abort_flag, p.err_word);
    tc_fence_after();
    if (dbg && threadIdx.x == 64) dbg[4] = clock64();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c = 0; c < p.BN; c += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c, v);
      const int co = col0 + c;
      if (!valid || co >= p.Cout) continue;
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
      if (p.flags & R2N2_EPI_BIAS) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 b4 = *reinterpret_cast<const float4*>(bias + co + j);
          f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
        }
      }
      if (p.flags & R2N2_EPI_LRELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = lrelu(f[j]);
      }
      const long long off = pix * p.Cout + co;
      if (p.flags & R2N2_EPI_RESIDUAL) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 r4 = Vec4<TOut>::load(res + off + j);
          f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
        }
      }
      if (p.flags & R2N2_EPI_MASK) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 m4 = Vec4<TOut>::load(mask + off + j);
          f[j] *= lrelu_slope(m4.x); f[j + 1] *= lrelu_slope(m4.y);
          f[j + 2] *= lrelu_slope(m4.z); f[j + 3] *= lrelu_slope(m4.w);
        }
      }
      store16(y + off, f);
    }
  }
  R2N2_WS_END(p.err_word);
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[5] = clock64();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
  if (dbg && threadIdx.x == 0) {
    dbg[6] = clock64();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    dbg[7] = smid;
  }
}

// --------------------------------------------------------------------------------- host side
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// one sticky error word PER DEVICE (a process may drive several GPUs; the word must live on the
// device whose kernels write it)
int* get_err_word() {
  static int* words[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!words[dev]) {
    int* w = nullptr;
#ifdef R2N2_WAITSTAT
    const size_t bytes = 64 + (size_t)512 * 16 * 40 * sizeof(long long);     // error word + wait statistics
#else
    const size_t bytes = sizeof(int);
#endif
    if (cudaMalloc(&w, bytes) != cudaSuccess) return nullptr;
    cudaMemset(w, 0, bytes);
    words[dev] = w;
  }
  return words[dev];
}

// choose the spatial box (bw,bh,bd,bn), product <= 128, minimising the number of M tiles
static void choose_box(int N, int D, int H, int W, int& bw, int& bh, int& bd, int& bn) {
  long long best = -1;
  int best_rows = 0;
  for (int w = 1; w <= W && w <= 128; ++w) {
    for (int h = 1; h <= H && w * h <= 128; ++h) {
      if (h > 1 && w < W && false) continue;
      for (int d = 1; d <= D && w * h * d <= 128; ++d) {
        int n = 128 / (w * h * d);
        if (n > N) n = N;
        if (n < 1) continue;
        long long tiles = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((D + d - 1) / d) *
                          ((N + n - 1) / n);
        int rows = w * h * d * n;
        if (best < 0 || tiles < best || (tiles == best && rows < best_rows)) {
          best = tiles; best_rows = rows;
          bw = w; bh = h; bd = d; bn = n;
        }
      }
    }
  }
}

CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static void umma_debug_report(const UmmaConvParams& p, dim3 grid, long long* dbg);

int conv_fwd_umma_v1(const void* x, const void* w, const float* bias, const void* mask,
                     const void* res, void* y, int N, int D, int H, int W, int Cin, int Cout, int kd,
                     int kh, int kw, int flags, int in_dtype, int out_dtype, cudaStream_t st) {
  R2N2_CHECK_ARG(in_dtype == R2N2_BF16, R2N2_ERR_BAD_DTYPE, "conv_fwd[umma]: operands must be bf16");
  R2N2_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, R2N2_ERR_BAD_SHAPE,
                 "conv_fwd[umma]: non-positive dimension");
  R2N2_CHECK_ARG((kd & 1) && (kh & 1) && (kw & 1), R2N2_ERR_BAD_SHAPE,
                 "conv_fwd[umma]: kernel extents must be odd");
  R2N2_CHECK_ARG(Cin % 16 == 0 && Cout % 16 == 0, R2N2_ERR_UNSUPPORTED,
                 "conv_fwd[umma]: Cin (%d) and Cout (%d) must be multiples of 16 (use ALGO_SIMT)", Cin, Cout);
  R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 R2N2_ERR_BAD_ALIGN, "conv_fwd[umma]: x / w must be 16-byte aligned");
  R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 31) == 0 &&
                     (!(flags & R2N2_EPI_RESIDUAL) || (reinterpret_cast<uintptr_t>(res) & 31) == 0) &&
                     (!(flags & R2N2_EPI_MASK) || (reinterpret_cast<uintptr_t>(mask) & 31) == 0),
                 R2N2_ERR_BAD_ALIGN, "conv_fwd[umma]: y / residual / mask must be 32-byte aligned");
  EncodeTiledFn encode = get_encode();
  R2N2_CHECK_ARG(encode != nullptr, R2N2_ERR_CUDA, "conv_fwd[umma]: cuTensorMapEncodeTiled unavailable");

  UmmaConvParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.kd = kd; p.kh = kh; p.kw = kw;
  choose_box(N, D, H, W, p.bw, p.bh, p.bd, p.bn);
  p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_d = (D + p.bd - 1) / p.bd; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.BN = Cout >= 128 ? 128 : Cout;
  p.kbox = Cin >= 64 ? 64 : (Cin >= 32 ? 32 : 16);
  p.kchunks = (Cin + p.kbox - 1) / p.kbox;
  const int row_bytes = p.kbox * 2;
  p.a_stage_bytes = (128 * row_bytes + 1023) / 1024 * 1024;
  p.b_stage_bytes = (p.BN * row_bytes + 1023) / 1024 * 1024;
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  p.stages = stage_bytes >= 32768 ? 3 : 4;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.BN) p.tmem_cols *= 2;
  p.flags = flags;
  p.err_word = get_err_word();
  p.dbg = nullptr;
  const bool want_dbg = getenv("R2N2_UMMA_DBG") != nullptr;
  if (const char* e = getenv("R2N2_UMMA_STAGES")) p.stages = atoi(e);          // perf experiments
  if (const char* e = getenv("R2N2_UMMA_SKIP")) {
    int v = atoi(e);
    if (v == 1 || v == 2) p.flags |= v << 16;
  }
  if (const char* e = getenv("R2N2_UMMA_BOX")) {
    sscanf(e, "%d,%d,%d,%d", &p.bw, &p.bh, &p.bd, &p.bn);
    p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh;
    p.tiles_d = (D + p.bd - 1) / p.bd; p.tiles_n = (N + p.bn - 1) / p.bn;
  }

  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kbox, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd,
                         (cuuint32_t)p.bn};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(row_bytes),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA,
                   "conv_fwd[umma]: cuTensorMapEncodeTiled(A) failed (%d) dims %d,%d,%d,%d,%d box %d,%d,%d,%d,%d",
                   (int)r, Cin, W, H, D, N, p.kbox, p.bw, p.bh, p.bd, p.bn);
  }
  {
    const long long K = (long long)kd * kh * kw * Cin;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kbox, (cuuint32_t)p.BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(row_bytes),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA,
                   "conv_fwd[umma]: cuTensorMapEncodeTiled(B) failed (%d) K %lld Cout %d box %d,%d", (int)r,
                   K, Cout, p.kbox, p.BN);
  }
  const size_t smem_bytes = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
  dim3 grid((unsigned)((long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n),
            (unsigned)((Cout + p.BN - 1) / p.BN));
  cudaError_t e;
  if (want_dbg) {
    cudaMalloc(&p.dbg, ((size_t)grid.x * grid.y * 8 + 4 * 512) * sizeof(long long));
    cudaMemset(p.dbg, 0, ((size_t)grid.x * grid.y * 8 + 4 * 512) * sizeof(long long));
  }
  if (out_dtype == R2N2_BF16) {
    static bool attr_set = false;
    if (!attr_set) {
      e = cudaFuncSetAttribute(conv_fwd_umma_kernel<__nv_bfloat16>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "conv_fwd[umma]: smem attribute: %s",
                     cudaGetErrorString(e));
      attr_set = true;
    }
    r2n2::launch_k((conv_fwd_umma_kernel<__nv_bfloat16>), dim3(grid), dim3(kUmmaThreads), smem_bytes, (cudaStream_t)(st), 0,
        map_a, map_b, bias, (const __nv_bfloat16*)mask, (const __nv_bfloat16*)res, (__nv_bfloat16*)y, p);
  } else if (out_dtype == R2N2_F32) {
    static bool attr_set = false;
    if (!attr_set) {
      e = cudaFuncSetAttribute(conv_fwd_umma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               200 * 1024);
      R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "conv_fwd[umma]: smem attribute: %s",
                     cudaGetErrorString(e));
      attr_set = true;
    }
    r2n2::launch_k((conv_fwd_umma_kernel<float>), dim3(grid), dim3(kUmmaThreads), smem_bytes, (cudaStream_t)(st), 0,
        map_a, map_b, bias, (const float*)mask, (const float*)res, (float*)y, p);
  } else {
    R2N2_CHECK_ARG(false, R2N2_ERR_BAD_DTYPE, "conv_fwd[umma]: bad out dtype %d", out_dtype);
  }
  R2N2_CHECK_LAUNCH("conv_fwd[umma]");
  if (want_dbg) umma_debug_report(p, grid, p.dbg);
  return R2N2_OK;
}

// debugging aid for tools/bench_conv.py: per-CTA phase timestamps of ONE launch, summarised
static void umma_debug_report(const UmmaConvParams& p, dim3 grid, long long* dbg) {
  cudaDeviceSynchronize();
  size_t n = (size_t)grid.x * grid.y;
  long long* h = (long long*)malloc(n * 8 * sizeof(long long));
  cudaMemcpy(h, dbg, n * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  double s[6] = {0, 0, 0, 0, 0, 0};
  long long tmin = h[0], tmax = 0;
  for (size_t i = 0; i < n; ++i) {
    long long* d = h + i * 8;
    s[0] += d[1] - d[0];   // setup (barrier init + TMEM alloc + sync)
    s[1] += d[2] - d[1];   // first k-block arrives
    s[2] += d[3] - d[2];   // main loop (first full -> last full)
    s[3] += d[4] - d[3];   // last MMAs -> accumulator ready
    s[4] += d[5] - d[4];   // epilogue
    s[5] += d[6] - d[0];   // total CTA
    if (d[0] < tmin) tmin = d[0];
    if (d[6] > tmax) tmax = d[6];
  }
  fprintf(stderr,
          "[umma dbg] grid %ux%u box %dx%dx%dx%d BN %d kbox %d stages %d nkb %d | avg cycles: setup %.0f "
          "first-load %.0f mainloop %.0f drain %.0f epilogue %.0f total %.0f | span %lld\n",
          grid.x, grid.y, p.bw, p.bh, p.bd, p.bn, p.BN, p.kbox, p.stages, p.kd * p.kh * p.kw * p.kchunks,
          s[0] / n, s[1] / n, s[2] / n, s[3] / n, s[4] / n, s[5] / n, tmax - tmin);
  {
    int nkb = p.kd * p.kh * p.kw * p.kchunks;
    if (nkb > 64) nkb = 64;
    long long tlh[4 * 64];
    cudaMemcpy(tlh, dbg + n * 8, sizeof(long long) * 4 * nkb, cudaMemcpyDeviceToHost);
    long long t0 = tlh[0];
    fprintf(stderr, "[umma timeline] kb: empty-ok  tma-issued  full-seen  commit-issued (cycles from first)\n");
    for (int i = 0; i < nkb; ++i)
      fprintf(stderr, "[umma timeline] %2d: %7lld %7lld %7lld %7lld\n", i, tlh[i * 4] - t0, tlh[i * 4 + 1] - t0,
              tlh[i * 4 + 2] - t0, tlh[i * 4 + 3] - t0);
  }
  free(h);
  cudaFree(dbg);
}

int umma_error_flag() {
  int* w = get_err_word();
  int h = 0;
  if (w) cudaMemcpy(&h, w, sizeof(int), cudaMemcpyDeviceToHost);
  return h;
}

// =====================================================================================
// Weight gradient:  dw[co][tap][ci] += sum_p dy[p, co] * x[p + off(tap), ci]
//
// GEMM view per tap: M = co (TMEM lanes), N = ci (TMEM columns), K = pixels.  Both operands are
// "MN-major" in memory (channels contiguous), so the TMA boxes [pixels][64 channels] land in the
// canonical MN-major swizzled layout and the K loop runs over pixel boxes; the x box is loaded
// shifted by the tap (zero fill = 'same' padding).  One CTA owns a 128(co) x ci_tile x G(taps)
// block of dw (G accumulators side by side in TMEM, G * ci_cols <= 512 columns), loops over its
// share of the pixel boxes and finally reduces into dw with red.global.add.f32 (fp32 atomics
// combine the pixel splits).
struct UmmaWgradParams {
  int N, D, H, W, Cin, Cout;
  int kd, kh, kw, taps;
  int bw, bh, bd, bn, rows;                // pixel box = K block (rows % 16 == 0)
  int tiles_w, tiles_h, tiles_d, tiles_n;
  long long nbox, box_per_cta;
  int ci_tiles, ci_tile, ci_cols;          // N tile (<=256), its TMEM column footprint
  int G;                                   // taps per CTA
  int cw_a, cw_b;                          // channels per smem chunk (16/32/64) for dy / x
  int a_bytes, b_tap_bytes, stage_bytes;   // per stage
  int stages, tmem_cols;
  int up2;                                 // x is the small tensor of an UP2 conv: dy is read with stride 2,
  int a_tap_bytes, b_off;                  //   one dy tile PER TAP (A side), one shared x tile (B side)
  int dxh;                                 // dx-halo mode: the CTA owns the kw=3 taps of ONE (dz,dy) filter row and loads
  int P, segs, chunk_bx;                   //   ONE x box that is 2 pixels wider (pitch P = bw+2); tap dx = view shifted by dx rows
  int ystack;                              // kh x 1 filter over one channel chunk (row-packed conv1a): ONE x box with kh-1 halo
                                           //   rows; the kh taps are stacked along N (B chunk stride = one image row): 1 MMA per K16
  int plain_store;                         // one pixel split and no accumulation: every dw element has exactly one writer ->
                                           //   plain 16-byte stores instead of memset + red.global.add (FC / x-projection layers)
  int* err_word;
};

// NKT / GT / SEGT > 0: K16 steps per stage, taps per CTA and (dx-halo mode) K16 steps per image row are compile-time, so
// the MMA issue sequence of a stage is straight-line code with hoisted descriptor offsets.  Round 2 (tools/waitstat.py):
// the generic loop kept the issuing thread busy 96 % of the kernel at ~85 cycles per N = 96 MMA (56 back to back) --
// like the forward kernels before their issue sequences were specialised, the weight gradients were issue-bound.
template <int NKT, int GT, int SEGT>
__global__ void __launch_bounds__(kUmmaThreads)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                       float* __restrict__ dw, const UmmaWgradParams p) {
  r2n2::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  R2N2_WS_BEGIN(bars);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full_bar = bars + 2 * p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_ptr_smem + 1);
  int* tap_off = reinterpret_cast<int*>(tmem_ptr_smem + 2);      // [G] packed (dx,dy,dz)+8

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int co_tile = blockIdx.y / p.ci_tiles;
  const int ci_t = blockIdx.y - co_tile * p.ci_tiles;
  const int co0 = co_tile * 128;
  const int ci0 = ci_t * p.ci_tile;
  const int tap0 = blockIdx.z * p.G;
  int ntap = p.taps - tap0;
  if (ntap > p.G) ntap = p.G;
  int nci = p.Cin - ci0;                       // valid ci in this tile (multiple of 16)
  if (nci > p.ci_tile) nci = p.ci_tile;
  if (p.ystack) { ntap = 1; nci = p.taps * p.Cin; }   // all taps live in ONE accumulator of taps*Cin columns
  int nco = p.Cout - co0;
  if (nco > 128) nco = 128;
  const int na = (nco + p.cw_a - 1) / p.cw_a;  // dy chunks to load
  const int nb = (nci + p.cw_b - 1) / p.cw_b;  // x chunks to load per tap
  const long long box_begin = (long long)blockIdx.x * p.box_per_cta;
  long long box_end = box_begin + p.box_per_cta;
  if (box_end > p.nbox) box_end = p.nbox;
  const int nkb = (int)(box_end > box_begin ? box_end - box_begin : 0);

  if (threadIdx.x < ntap) {
    const int tap = tap0 + threadIdx.x;
    const int tz = tap / (p.kh * p.kw);
    const int r2 = tap - tz * p.kh * p.kw;
    const int ty = r2 / p.kw;
    const int tx = r2 - ty * p.kw;
    tap_off[threadIdx.x] = (tx - (p.kw - 1) / 2 + 8) | ((ty - (p.kh - 1) / 2 + 8) << 8) |
                           ((tz - (p.kd - 1) / 2 + 8) << 16);
  }
  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  r2n2::pdl_wait();            // prologue done; from here on global memory of the preceding kernels may be touched

  const int rb_a = p.cw_a * 2, rb_b = p.cw_b * 2;        // smem row pitch (bytes) of dy / x chunks
  const int chunk_a = p.rows * rb_a, chunk_b = p.rows * rb_b;

  if (warp == 0) {
    if (nkb > 0) {
      // warp-uniform loop: incremental box coordinates, tap offsets precomputed in smem
      const uint32_t tx_bytes = p.ystack ? (uint32_t)(na * chunk_a + p.chunk_bx)
                                : p.dxh ? (uint32_t)(na * chunk_a + nb * p.chunk_bx)
                                : p.up2 ? (uint32_t)(ntap * na * chunk_a + nb * chunk_b)
                                        : (uint32_t)(na * chunk_a + ntap * nb * chunk_b);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      long long b = box_begin;
      int tw = (int)(b % p.tiles_w); b /= p.tiles_w;
      int th = (int)(b % p.tiles_h); b /= p.tiles_h;
      int td = (int)(b % p.tiles_d); b /= p.tiles_d;
      int tn = (int)b;
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty0 + 8u * s, ph ^ 1u, abort_flag, p.err_word);
        const int x0 = tw * p.bw, y0 = th * p.bh, z0 = td * p.bd, n0 = tn * p.bn;
        const uint32_t stage = smem_base + (uint32_t)(s * p.stage_bytes);
        const uint32_t bar = full0 + 8u * s;
        if (elect_one()) {
          mbar_expect_tx(bar, tx_bytes);
          if (p.ystack) {
            for (int c = 0; c < na; ++c)
              tma_load_5d(stage + c * chunk_a, &map_dy, bar, co0 + c * p.cw_a, x0, y0, z0, n0);
            tma_load_5d(stage + p.b_off, &map_x, bar, 0, x0, y0 - (p.kh - 1) / 2, z0, n0);
          } else if (p.dxh) {
            for (int c = 0; c < na; ++c)
              tma_load_5d(stage + c * chunk_a, &map_dy, bar, co0 + c * p.cw_a, x0, y0, z0, n0);
            const int o = tap_off[0];
            const int oy = ((o >> 8) & 0xff) - 8, oz = ((o >> 16) & 0xff) - 8;
            const uint32_t bdst = stage + p.b_off;
            for (int c = 0; c < nb; ++c)
              tma_load_5d(bdst + c * p.chunk_bx, &map_x, bar, ci0 + c * p.cw_b, x0 - 1, y0 + oy, z0 + oz, n0);
          } else if (!p.up2) {
            for (int c = 0; c < na; ++c)
              tma_load_5d(stage + c * chunk_a, &map_dy, bar, co0 + c * p.cw_a, x0, y0, z0, n0);
            uint32_t bdst = stage + p.b_off;
            for (int g = 0; g < ntap; ++g) {
              const int o = tap_off[g];
              const int ox = (o & 0xff) - 8, oy = ((o >> 8) & 0xff) - 8, oz = ((o >> 16) & 0xff) - 8;
              for (int c = 0; c < nb; ++c)
                tma_load_5d(bdst + c * chunk_b, &map_x, bar, ci0 + c * p.cw_b, x0 + ox, y0 + oy, z0 + oz, n0);
              bdst += p.b_tap_bytes;
            }
          } else {
            // zero-stuffed input: xs[o + tap] is stored only where o + tap = 2i, i.e. the small pixel i
            // meets dy[2i - tap] -> one stride-2 dy box per tap, one shared box of the small x
            uint32_t adst = stage;
            for (int g = 0; g < ntap; ++g) {
              const int o = tap_off[g];
              const int ox = (o & 0xff) - 8, oy = ((o >> 8) & 0xff) - 8, oz = ((o >> 16) & 0xff) - 8;
              for (int c = 0; c < na; ++c)
                tma_load_5d(adst + c * chunk_a, &map_dy, bar, co0 + c * p.cw_a, 2 * x0 - ox, 2 * y0 - oy,
                            2 * z0 - oz, n0);
              adst += p.a_tap_bytes;
            }
            const uint32_t bdst = stage + p.b_off;
            for (int c = 0; c < nb; ++c)
              tma_load_5d(bdst + c * chunk_b, &map_x, bar, ci0 + c * p.cw_b, x0, y0, z0, n0);
          }
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            if (++td == p.tiles_d) { td = 0; ++tn; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (nkb > 0) {
      const int nmma = (nci + 15) & ~15;                 // UMMA N
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lt_a = (rb_a == 128) ? 2u : (rb_a == 64 ? 4u : 6u);
      const uint32_t lt_b = (rb_b == 128) ? 2u : (rb_b == 64 ? 4u : 6u);
      const uint64_t da_hi = make_mnmajor_desc(0, chunk_a, 8 * rb_a, lt_a);
      // ystack: chunk j of the N side = the x tile shifted by j image rows (bw pixels) = filter tap j
      const uint64_t db_hi = make_mnmajor_desc(0, p.ystack ? p.bw * rb_b : (p.dxh ? p.chunk_bx : chunk_b), 8 * rb_b, lt_b);
      const uint32_t ka = (uint32_t)(16 * rb_a) >> 4, kbs = (uint32_t)(16 * rb_b) >> 4;  // desc step / K16
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const int nk = p.rows >> 4;
      const uint32_t a_step = p.up2 ? (uint32_t)p.a_tap_bytes : 0u;       // the operand that changes per tap
      const uint32_t b_step = (p.up2 || p.dxh) ? 0u : (uint32_t)p.b_tap_bytes;
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full0 + 8u * s, ph, abort_flag, p.err_word);
        tc_fence_after();
        if (NKT > 0 && ntap == GT) {
          // compile-time issue sequence (every tap group of this CTA is full, NKT K16 steps per stage)
          if (elect_one()) {
            const uint32_t a_addr = smem_base + (uint32_t)(s * p.stage_bytes);
            const uint64_t da0 = da_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
            const uint64_t db0 = db_hi | (uint64_t)(((a_addr + p.b_off) & 0x3FFFF) >> 4);
            const uint32_t a_step16 = a_step >> 4, b_step16 = (p.dxh ? (uint32_t)rb_b : b_step) >> 4;
            const uint32_t row_skip16 = (uint32_t)(2 * rb_b) >> 4;         // dx-halo: the x tile is 2 pixels wider per image row
#pragma unroll
            for (int g = 0; g < (GT > 0 ? GT : 1); ++g) {
#pragma unroll
              for (int k = 0; k < (NKT > 0 ? NKT : 1); ++k) {
                const uint32_t koff = kbs * (uint32_t)k + (SEGT > 0 ? (uint32_t)(k / (SEGT > 0 ? SEGT : 1)) * row_skip16 : 0u);
                umma_bf16(tmem_base + (uint32_t)(g * p.ci_cols), da0 + (uint64_t)(ka * (uint32_t)k + a_step16 * (uint32_t)g),
                          db0 + (uint64_t)(koff + b_step16 * (uint32_t)g), idesc, k == 0 ? acc : 1u);
              }
            }
            umma_commit(empty0 + 8u * s);
          }
        } else if (elect_one()) {
          uint32_t a_addr = smem_base + (uint32_t)(s * p.stage_bytes);
          uint32_t b_addr = a_addr + p.b_off;
          uint32_t d_tmem = tmem_base;
          for (int g = 0; g < ntap; ++g) {
            const uint64_t da0 = da_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
            const uint64_t db0 = db_hi | (uint64_t)((b_addr & 0x3FFFF) >> 4);
            uint32_t a2 = acc;
            if (p.dxh) {
              // K16 step = 16 consecutive pixels of ONE image row (bw % 16 == 0): in the wider x tile the same
              // pixels start 2 rows further per image row, and tap g looks g pixels to the right
              uint32_t koff = 0;
              int seg = 0;
              for (int k = 0; k < nk; ++k) {
                umma_bf16(d_tmem, da0 + (uint64_t)(ka * k), db0 + (uint64_t)koff, idesc, a2);
                a2 = 1;
                koff += kbs;
                if (++seg == p.segs) { seg = 0; koff += (uint32_t)(2 * rb_b) >> 4; }
              }
              b_addr += (uint32_t)rb_b;                      // next dx tap: one pixel to the right
            } else {
              for (int k = 0; k < nk; ++k) {
                umma_bf16(d_tmem, da0 + (uint64_t)(ka * k), db0 + (uint64_t)(kbs * k), idesc, a2);
                a2 = 1;
              }
            }
            a_addr += a_step;
            b_addr += b_step;
            d_tmem += (uint32_t)p.ci_cols;
          }
          umma_commit(empty0 + 8u * s);
        }
        __syncwarp();
        acc = 1;
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(smem_u32(tmem_full_bar));
    }
    __syncwarp();
  } else if (nkb > 0) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                    // co within the tile
    const bool valid = row < nco;
    mbar_wait(smem_u32(tmem_full_bar), 0u, abort_flag, p.err_word);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int g = 0; g < ntap; ++g) {
      float* dst = dw + ((long long)(co0 + row) * p.taps + (tap0 + g)) * p.Cin + ci0;
      for (int c = 0; c < nci; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)(g * p.ci_cols + c), v);
        if (!valid) continue;
        if (p.plain_store) {
          float f16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f16[j] = __uint_as_float(v[j]);
          store16(dst + c, f16);                           // two 256-bit stores: whole 32-byte sectors
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            red_add_v4(dst + c + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  }
  R2N2_WS_END(p.err_word);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

template <int NKT, int GT, int SEGT>
static int launch_wgrad(const CUtensorMap& map_dy, const CUtensorMap& map_x, float* dw, const UmmaWgradParams& p, dim3 grid,
                        size_t smem_bytes, cudaStream_t st) {
  auto kern = conv_wgrad_umma_kernel<NKT, GT, SEGT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "conv_wgrad[umma]: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaError_t le = r2n2::launch_k((kern), grid, dim3(kUmmaThreads), smem_bytes, st, 0, map_dy, map_x, dw, p);
  R2N2_CHECK_ARG(le == cudaSuccess, R2N2_ERR_CUDA, "conv_wgrad[umma]: launch: %s", cudaGetErrorString(le));
  return R2N2_OK;
}

// pixel box for the K loop: rows = bw*bh*bd*bn must be a multiple of 16 and <= max_rows;
// maximise the fraction of in-bounds pixels (out-of-bounds ones are zero-filled by TMA and only
// cost time), then the box size.
static void choose_kbox(int N, int D, int H, int W, int max_rows, int& bw, int& bh, int& bd, int& bn) {
  double best_eff = -1.0;
  int best_rows = 0;
  bw = bh = bd = 1; bn = 16;
  for (int w = 1; w <= W && w <= max_rows; ++w)
    for (int h = 1; h <= H && w * h <= max_rows; ++h)
      for (int d = 1; d <= D && w * h * d <= max_rows; ++d)
        for (int n = 1; n <= N && w * h * d * n <= max_rows; ++n) {
          int rows = w * h * d * n;
          if (rows % 16) continue;
          double cover = (double)((W + w - 1) / w * w) * ((H + h - 1) / h * h) * ((D + d - 1) / d * d) *
                         ((N + n - 1) / n * n);
          double eff = ((double)W * H * D * N) / cover;
          // max in-bounds fraction, then the biggest K block, then the box that follows the memory
          // order (long runs along x, then y): contiguous TMA rows and L2 reuse between shifted taps
          const bool better = eff > best_eff + 1e-9 ||
                              (eff > best_eff - 1e-9 &&
                               (rows > best_rows ||
                                (rows == best_rows && (w > bw || (w == bw && (h > bh || (h == bh && d > bd)))))));
          if (better) {
            best_eff = eff; best_rows = rows;
            bw = w; bh = h; bd = d; bn = n;
          }
        }
}

// dx-halo variant: box width a multiple of 16 (may overhang the image: TMA zero-fills), so that every K16
// step lies inside one image row
static void choose_kbox_dx(int N, int D, int H, int W, int max_rows, int& bw, int& bh, int& bd, int& bn) {
  double best_eff = -1.0;
  int best_rows = 0;
  bw = 0;
  for (int w = 16; w <= 240 && w <= (W + 15) / 16 * 16 && w <= max_rows; w += 16)
    for (int h = 1; h <= H && w * h <= max_rows; ++h)
      for (int d = 1; d <= D && w * h * d <= max_rows; ++d)
        for (int n = 1; n <= N && w * h * d * n <= max_rows; ++n) {
          int rows = w * h * d * n;
          double cover = (double)((W + w - 1) / w * w) * ((H + h - 1) / h * h) * ((D + d - 1) / d * d) *
                         ((N + n - 1) / n * n);
          double eff = ((double)W * H * D * N) / cover;
          const bool better = eff > best_eff + 1e-9 ||
                              (eff > best_eff - 1e-9 &&
                               (rows > best_rows || (rows == best_rows && (w > bw || (w == bw && h > bh)))));
          if (better) { best_eff = eff; best_rows = rows; bw = w; bh = h; bd = d; bn = n; }
        }
  if (best_eff < 0.8) bw = 0;                // too much overhang (e.g. W = 33): keep the per-tap loads
}

// channels per TMA box / swizzle span; 96 (160) = 3 (5) x 32: no zero-filled half chunk -> 25 % less smem + traffic
static int pow2_chunk(int c) {
  if (c > 64 && c % 64 == 32 && !getenv("R2N2_WGRAD_CHUNK64")) return 32;
  return c >= 64 ? 64 : (c >= 32 ? 32 : 16);
}

int conv_wgrad2_umma(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin,
                     int Cout, int kd, int kh, int kw, int accumulate, cudaStream_t st);
bool conv_wgrad3_supported(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw);
int conv_wgrad3_umma(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin, int Cout,
                     int accumulate, cudaStream_t st);

int conv_wgrad_umma(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin,
                    int Cout, int kd, int kh, int kw, int dtype, int wflags, cudaStream_t st) {
  R2N2_CHECK_ARG(dtype == R2N2_BF16, R2N2_ERR_BAD_DTYPE, "conv_wgrad[umma]: operands must be bf16");
  const int accumulate = wflags & R2N2_WGRAD_ACCUMULATE;
  const int up2 = (wflags & R2N2_WGRAD_UP2) ? 1 : 0;
  // 1x1 convolutions with few output channels: co on the N side, channels stacked along M
  // (conv_umma_wgrad2.cu; measured faster only there -- with 27 taps the per-tap reloads dominate)
  if (!up2 && Cout <= 64 && kd * kh * kw == 1 && Cin % 16 == 0 && Cout % 16 == 0 && N > 0 && D > 0 && H > 0 &&
      W > 0 && !getenv("R2N2_UMMA_WGRAD_V1"))
    return conv_wgrad2_umma(x, dy, dw, N, D, H, W, Cin, Cout, kd, kh, kw, accumulate, st);
  // 3x3x3 with narrow channels: flat halo tiles, shifted descriptor views (conv_umma_wgrad3.cu)
  if (!up2 && conv_wgrad3_supported(N, D, H, W, Cin, Cout, kd, kh, kw))
    return conv_wgrad3_umma(x, dy, dw, N, D, H, W, Cin, Cout, accumulate, st);
  R2N2_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, R2N2_ERR_BAD_SHAPE,
                 "conv_wgrad[umma]: non-positive dimension");
  R2N2_CHECK_ARG((kd & 1) && (kh & 1) && (kw & 1), R2N2_ERR_BAD_SHAPE,
                 "conv_wgrad[umma]: kernel extents must be odd");
  R2N2_CHECK_ARG(Cin % 16 == 0 && Cout % 16 == 0, R2N2_ERR_UNSUPPORTED,
                 "conv_wgrad[umma]: Cin (%d) and Cout (%d) must be multiples of 16 (use ALGO_SIMT)", Cin, Cout);
  EncodeTiledFn encode = get_encode();
  R2N2_CHECK_ARG(encode != nullptr, R2N2_ERR_CUDA, "conv_wgrad[umma]: cuTensorMapEncodeTiled unavailable");

  UmmaWgradParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.kd = kd; p.kh = kh; p.kw = kw; p.taps = kd * kh * kw;
  p.cw_a = pow2_chunk(Cout);
  p.cw_b = pow2_chunk(Cin);
  p.ci_tile = Cin > 256 ? 256 : Cin;
  p.ci_tiles = (Cin + p.ci_tile - 1) / p.ci_tile;
  p.ci_cols = (p.ci_tile + 31) / 32 * 32;
  p.G = 512 / p.ci_cols;
  if (p.G > p.taps) p.G = p.taps;
  // same number of tap groups, evenly filled: 9 taps at G = 4 would be groups of 4, 4, 1 (the CTAs of the
  // last group idle 3/4 of the time); 3, 3, 3 is the same grid with 25 % less work on the critical CTAs
  if (!getenv("R2N2_WGRAD_NO_BALANCE")) {
    const int groups = (p.taps + p.G - 1) / p.G;
    p.G = (p.taps + groups - 1) / groups;
  }
  const int co_tiles = (Cout + 127) / 128;
  // bytes per K row (pixel) in one stage
  const int a_row = ((Cout < 128 ? Cout : 128) + p.cw_a - 1) / p.cw_a * p.cw_a * 2;
  const int b_row = (p.ci_tile + p.cw_b - 1) / p.cw_b * p.cw_b * 2;
  p.up2 = up2;
  // dx-halo mode: 3x3(x3) filters whose rows can be cut into 16-pixel segments; the CTA takes the 3 dx taps
  // of one filter row from ONE x box (3x less L2 -> SM traffic for x, the operand that bounded these layers)
  p.dxh = 0;
  p.stages = 3;
  int max_rows = 0;
  if (!up2 && kw == 3 && p.ci_tile <= 160 && !getenv("R2N2_WGRAD_NO_DXH")) {
    const int per_row_dx = a_row + b_row * 5 / 4;
    int mr = (180 * 1024 / p.stages) / per_row_dx / 16 * 16;
    if (mr > 256) mr = 256;
    if (mr >= 64) {
      choose_kbox_dx(N, D, H, W, mr, p.bw, p.bh, p.bd, p.bn);
      if (p.bw > 0 && p.bw * p.bh * p.bd * p.bn >= 64) { p.dxh = 1; p.G = 3; max_rows = mr; }
    }
  }
  p.ystack = 0;
  if (!up2 && !p.dxh && kd == 1 && kw == 1 && kh > 1 && Cin == p.cw_b && kh * Cin <= 256 && Cout <= 128 &&
      !getenv("R2N2_WGRAD_NO_YSTACK")) {
    // box (bw multiple of 16) x bh image rows of one sample; the x tile carries kh-1 halo rows: pick the box
    // with the least overhang, then the least halo traffic per pixel, within a 60 KB stage
    double best = -1.0;
    int bw_ = 0, bh_ = 0;
    for (int w = 16; w <= 240 && w <= (W + 15) / 16 * 16; w += 16)
      for (int h = 1; h <= H && w * h <= 256; ++h) {
        const int stage = w * h * a_row + w * (h + kh - 1) * b_row;
        if (stage > 60 * 1024 || w * h < 64) continue;
        const double cover = (double)((W + w - 1) / w * w) * ((H + h - 1) / h * h);
        const double score = ((double)W * H / cover) * ((double)h / (h + kh - 1) * 0.5 + 0.5);
        if (score > best + 1e-9) { best = score; bw_ = w; bh_ = h; }
      }
    if (bw_ > 0 && (double)W / ((W + bw_ - 1) / bw_ * bw_) >= 0.8) {
      p.ystack = 1; p.G = 1; max_rows = 256;
      p.bw = bw_; p.bh = bh_; p.bd = 1; p.bn = 1;
      p.ci_cols = (kh * Cin + 31) / 32 * 32;
    }
  }
  if (!p.dxh && !p.ystack) {
    const int per_row = up2 ? p.G * a_row + (b_row > 256 - a_row ? b_row : 256 - a_row) : a_row + p.G * b_row;
    max_rows = (180 * 1024 / p.stages) / per_row / 16 * 16;
    if (max_rows > 128) max_rows = 128;
    if (max_rows < 16) {
      p.stages = 2;
      max_rows = (180 * 1024 / p.stages) / per_row / 16 * 16;
    }
    R2N2_CHECK_ARG(max_rows >= 16, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma]: tile does not fit shared memory");
    choose_kbox(N, D, H, W, max_rows, p.bw, p.bh, p.bd, p.bn);
  }
  const int tap_groups = p.ystack ? 1 : (p.taps + p.G - 1) / p.G;
  p.rows = p.bw * p.bh * p.bd * p.bn;
  p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_d = (D + p.bd - 1) / p.bd; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.nbox = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  p.a_bytes = (p.rows * a_row + 1023) / 1024 * 1024;
  p.b_tap_bytes = p.rows * b_row;
  p.a_tap_bytes = (p.rows * a_row + 1023) / 1024 * 1024;
  if (up2) {
    // [G dy tiles][x tile]; an M=128 MMA on a narrower dy tile reads on into the next tile / the x
    // tile (finite data, rows discarded by the epilogue): the x tile is padded to cover that span
    p.b_off = p.G * p.a_tap_bytes;
    // (an M=128 MN-major A view always spans rows * 256 bytes from the start of its tile)
    const int span = p.rows * 256 - p.a_tap_bytes;
    const int xb = p.rows * b_row > span ? p.rows * b_row : span;
    p.stage_bytes = (p.b_off + xb + 1023) / 1024 * 1024;
  } else if (p.ystack) {
    p.chunk_bx = p.bw * (p.bh + kh - 1) * p.cw_b * 2;
    p.b_off = p.a_bytes;
    p.stage_bytes = (p.a_bytes + p.chunk_bx + 1023) / 1024 * 1024;
  } else if (p.dxh) {
    p.P = p.bw + 2;
    p.segs = p.bw / 16;
    p.chunk_bx = p.P * p.bh * p.bd * p.bn * p.cw_b * 2;                 // one channel chunk of the wider x tile
    p.b_off = p.a_bytes;
    // (+2 rows: the dx = 2 view of the last K16 step runs two pixels past the tile)
    p.stage_bytes = (p.a_bytes + ((p.ci_tile + p.cw_b - 1) / p.cw_b) * p.chunk_bx + 2 * p.cw_b * 2 + 1023) / 1024 * 1024;
  } else {
    p.b_off = p.a_bytes;
    p.stage_bytes = (p.a_bytes + p.G * p.b_tap_bytes + 1023) / 1024 * 1024;
  }
  p.tmem_cols = 32;
  while (p.tmem_cols < p.G * p.ci_cols) p.tmem_cols *= 2;
  p.err_word = get_err_word();
  // pixel splits: enough CTAs to fill the machine, but at least ~4 K blocks each
  const long long units = (long long)co_tiles * p.ci_tiles * tap_groups;
  // resident CTAs per SM (shared memory and the 512 TMEM columns): the grid should be ONE wave of them --
  // 297 CTAs at one CTA per SM is three waves with a single CTA in the last
  const size_t smem_est = (size_t)p.stages * p.stage_bytes + 2048 + (p.rows * 256 > p.stage_bytes ? p.rows * 256 - p.stage_bytes : 0);
  int occ = (int)((size_t)227 * 1024 / smem_est);
  if (occ > 512 / p.tmem_cols) occ = 512 / p.tmem_cols;
  if (occ < 1) occ = 1;
  if (occ > 2) occ = 2;
  long long splits = ((long long)num_sms() * occ) / units;
  if (splits > p.nbox / 4) splits = p.nbox / 4;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.box_per_cta = (p.nbox + splits - 1) / splits;
  splits = (p.nbox + p.box_per_cta - 1) / p.box_per_cta;

  CUtensorMap map_dy, map_x;
  {
    const int sc = up2 ? 2 : 1;
    const long long Wy = (long long)W * sc, Hy = (long long)H * sc, Dy = (long long)D * sc;
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)Wy, (cuuint64_t)Hy, (cuuint64_t)Dy, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wy * Cout * 2, (cuuint64_t)Hy * Wy * Cout * 2,
                             (cuuint64_t)Dy * Hy * Wy * Cout * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.cw_a, (cuuint32_t)(p.bw * sc), (cuuint32_t)(p.bh * sc),
                         (cuuint32_t)(p.bd * sc), (cuuint32_t)p.bn};
    cuuint32_t es[5] = {1, (cuuint32_t)sc, (cuuint32_t)sc, (cuuint32_t)sc, 1};
    CUresult r = encode(&map_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.cw_a * 2),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_wgrad[umma]: cuTensorMapEncodeTiled(dy) failed (%d)",
                   (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.cw_b, (cuuint32_t)(p.dxh ? p.bw + 2 : p.bw),
                         (cuuint32_t)(p.ystack ? p.bh + kh - 1 : p.bh), (cuuint32_t)p.bd, (cuuint32_t)p.bn};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.cw_b * 2),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_wgrad[umma]: cuTensorMapEncodeTiled(x) failed (%d)",
                   (int)r);
  }
  p.plain_store = (!accumulate && splits == 1 && (reinterpret_cast<uintptr_t>(dw) & 31) == 0 &&
                   !getenv("R2N2_WGRAD_NO_PLAIN_STORE")) ? 1 : 0;
  if (!accumulate && !p.plain_store)
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * p.taps * Cin, st);
  // the M=128 A view of the LAST stage may run past its tile (narrow Cout): keep it inside the window
  const int a_span_slack = p.rows * 256 > p.stage_bytes ? p.rows * 256 - p.stage_bytes : 0;
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + (2 * p.stages + 1) * 8 + 16 + 256 + 1024 + a_span_slack;
  dim3 grid((unsigned)splits, (unsigned)(co_tiles * p.ci_tiles), (unsigned)tap_groups);
  if (getenv("R2N2_UMMA_DBG"))
    fprintf(stderr, "[umma wgrad] Cin %d Cout %d taps %d G %d box %dx%dx%dx%d rows %d dxh %d ystack %d up2 %d stage %d x %d "
            "smem %zu grid %ux%ux%u tmem %d ci_cols %d\n", Cin, Cout, p.taps, p.G, p.bw, p.bh, p.bd, p.bn, p.rows, p.dxh,
            p.ystack, p.up2, p.stage_bytes, p.stages, smem_bytes, grid.x, grid.y, grid.z, p.tmem_cols, p.ci_cols);
  R2N2_CHECK_ARG(smem_bytes <= 200 * 1024, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma]: smem %zu too large", smem_bytes);
  // compile-time issue sequences for the configurations of the training step (anything else: the generic loop)
  const int nk = p.rows >> 4, seg = p.dxh ? p.segs : 0;
  const bool gen = getenv("R2N2_UMMA_WGRAD_GENERIC") != nullptr;
  int rc;
#define R2N2_WG(NK_, G_, SEG_) (!gen && nk == NK_ && p.G == G_ && seg == SEG_) rc = launch_wgrad<NK_, G_, SEG_>(map_dy, map_x, dw, p, grid, smem_bytes, st)
  if R2N2_WG(8, 3, 8);          // conv1b (128-pixel rows)
  else if R2N2_WG(6, 3, 2);     // conv2a / conv2b
  else if R2N2_WG(6, 3, 1);     // conv8a / conv8b
  else if R2N2_WG(8, 3, 2);     // conv9a
  else if R2N2_WG(9, 3, 1);     // conv9a data-gradient shape (UP2 layers)
  else if R2N2_WG(3, 3, 0);     // conv3a
  else if R2N2_WG(3, 2, 0);     // conv3b, conv4, conv5, conv6
  else if R2N2_WG(3, 4, 0);     // recurrent unit, conv7
  else if R2N2_WG(13, 1, 0);    // row-packed conv1a (y-stacked taps)
  else if R2N2_WG(8, 1, 0);     // 1x1 convolutions, 128-pixel boxes
  else if R2N2_WG(4, 1, 0);     // FC layers / x-projections (64-sample boxes)
  else rc = launch_wgrad<0, 0, 0>(map_dy, map_x, dw, p, grid, smem_bytes, st);
#undef R2N2_WG
  if (rc != R2N2_OK) return rc;
  R2N2_CHECK_LAUNCH("conv_wgrad[umma]");
  return R2N2_OK;
}

}  // namespace r2n2

extern "C" int r2n2_umma_error_flag(void) { return r2n2::umma_error_flag(); }
extern "C" int r2n2_set_sm_limit(int n_sms) { r2n2::sm_limit_ref() = n_sms; return R2N2_OK; }

#ifdef R2N2_WAITSTAT
// diagnostic build only (tools/waitstat.py): copy the wait statistics of the last tcgen05 launch to the host and clear them
extern "C" int r2n2_waitstat_read(long long* dst, long long n_words) {
  int* w = r2n2::get_err_word();
  if (!w) return 1;
  const long long cap = 512LL * 16 * 40;
  if (n_words > cap) n_words = cap;
  if (cudaDeviceSynchronize() != cudaSuccess) return 2;
  if (cudaMemcpy(dst, reinterpret_cast<long long*>(w) + 8, (size_t)n_words * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
  cudaMemset(reinterpret_cast<long long*>(w) + 8, 0, (size_t)cap * sizeof(long long));
  return 0;
}
#endif
