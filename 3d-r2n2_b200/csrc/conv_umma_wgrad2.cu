// tcgen05 weight gradient for NARROW output-channel counts (Cout <= 64), sm_100a.
//
//   dw[co][tap][ci] += sum_p dy[p, co] * x[p + off(tap), ci]
//
// conv_wgrad_umma_kernel (conv_umma.cu) puts co on the 128 TMEM lanes: with Cout = 64/32/16 (the
// 32^3 decoder stages conv9..conv11, 1.18 M pixels each) 50-87 % of every MMA is wasted.  Here the
// roles are swapped and taps are STACKED along M:
//     M (lanes)   = (tap_local, ci)   TS = 128 / Cin_tile taps per MMA   -> all 128 lanes useful
//     N (columns) = co                                                     -> Cout/2 cycles per K16
//     K           = pixels (TMA boxes [pixels][channels], MN-major operands for both sides)
// so one MMA covers TS taps and costs Cout/2 cycles instead of TS * Cin/2: a Cout/128 reduction.
// Every tap group owns Cout (rounded to 32) TMEM columns; a CTA handles up to 512/cols groups.
#include "umma.cuh"

namespace r2n2 {

struct Wgrad2Params {
  int N, D, H, W, Cin, Cout;
  int kd, kh, kw, taps;
  int bw, bh, bd, bn, rows;                // pixel box = K block (rows % 16 == 0)
  int tiles_w, tiles_h, tiles_d, tiles_n;
  long long nbox, box_per_cta;
  int ci_tile, ci_tiles;                   // channels of x per tap in the M stack (<= 128)
  int TS;                                  // taps per MMA (128 / ci_tile)
  int groups_per_cta, n_cols;              // tap groups per CTA, TMEM columns per group
  int cw_x, cw_y;                          // smem chunk widths (channels) of x / dy
  int y_bytes, tap_bytes, stage_bytes;     // per stage: dy tile, one tap's x tile
  int stages, tmem_cols;
  int* err_word;
};

__global__ void __launch_bounds__(kUmmaThreads)
conv_wgrad2_umma_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                        float* __restrict__ dw, const Wgrad2Params p) {
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
  int* tap_off = reinterpret_cast<int*>(tmem_ptr_smem + 2);      // [taps of this CTA] packed offsets

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int ci0 = blockIdx.y * p.ci_tile;
  const int tap0 = blockIdx.z * p.groups_per_cta * p.TS;
  int ntap = p.taps - tap0;
  if (ntap > p.groups_per_cta * p.TS) ntap = p.groups_per_cta * p.TS;
  const int ngroups = (ntap + p.TS - 1) / p.TS;
  const long long box_begin = (long long)blockIdx.x * p.box_per_cta;
  long long box_end = box_begin + p.box_per_cta;
  if (box_end > p.nbox) box_end = p.nbox;
  const int nkb = (int)(box_end > box_begin ? box_end - box_begin : 0);

  if ((int)threadIdx.x < ntap) {
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

  const int rb_x = p.cw_x * 2, rb_y = p.cw_y * 2;          // smem row pitch of x / dy chunks
  const int chunk_x = p.rows * rb_x, chunk_y = p.rows * rb_y;
  const int nx = (p.ci_tile + p.cw_x - 1) / p.cw_x;        // x chunks per tap
  const int ny = (p.Cout + p.cw_y - 1) / p.cw_y;           // dy chunks
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);

  if (warp == 0) {
    if (nkb > 0) {
      const uint32_t tx_bytes = (uint32_t)(ny * chunk_y + ntap * nx * chunk_x);
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
          for (int c = 0; c < ny; ++c)
            tma_load_5d(stage + c * chunk_y, &map_dy, bar, c * p.cw_y, x0, y0, z0, n0);
          uint32_t dst = stage + p.y_bytes;
          for (int g = 0; g < ntap; ++g) {
            const int o = tap_off[g];
            const int ox = (o & 0xff) - 8, oy = ((o >> 8) & 0xff) - 8, oz = ((o >> 16) & 0xff) - 8;
            for (int c = 0; c < nx; ++c)
              tma_load_5d(dst + c * chunk_x, &map_x, bar, ci0 + c * p.cw_x, x0 + ox, y0 + oy, z0 + oz, n0);
            dst += p.tap_bytes;
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
  } else if (warp == 1) {
    if (nkb > 0) {
      // A = stacked x taps (MN-major, M = 128), B = dy (MN-major, N = Cout)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lt_x = (rb_x == 128) ? 2u : (rb_x == 64 ? 4u : 6u);
      const uint32_t lt_y = (rb_y == 128) ? 2u : (rb_y == 64 ? 4u : 6u);
      const uint64_t dx_hi = make_mnmajor_desc(0, chunk_x, 8 * rb_x, lt_x);
      const uint64_t dy_hi = make_mnmajor_desc(0, chunk_y, 8 * rb_y, lt_y);
      const uint32_t kx = (uint32_t)(16 * rb_x) >> 4, ky = (uint32_t)(16 * rb_y) >> 4;
      const int nk = p.rows >> 4;
      const uint32_t group_bytes = (uint32_t)(p.TS * p.tap_bytes);
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full0 + 8u * s, ph, abort_flag, p.err_word);
        tc_fence_after();
        const uint32_t stage = smem_base + (uint32_t)(s * p.stage_bytes);
        if (elect_one()) {
          const uint64_t db0 = dy_hi | (uint64_t)((stage & 0x3FFFF) >> 4);
          uint32_t a_addr = stage + p.y_bytes;
          uint32_t d_tmem = tmem_base;
          for (int g = 0; g < ngroups; ++g) {
            const uint64_t da0 = dx_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
            uint32_t a2 = acc;
            for (int k = 0; k < nk; ++k) {
              umma_bf16(d_tmem, da0 + (uint64_t)(kx * k), db0 + (uint64_t)(ky * k), idesc, a2);
              a2 = 1;
            }
            a_addr += group_bytes;
            d_tmem += (uint32_t)p.n_cols;
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
    const int row = quad * 32 + lane;                 // (tap_local, ci) in the M stack
    const int tl = row / p.ci_tile;
    const int ci = row - tl * p.ci_tile;
    mbar_wait(smem_u32(tmem_full_bar), 0u, abort_flag, p.err_word);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int g = 0; g < ngroups; ++g) {
      const int tap = tap0 + g * p.TS + tl;
      const bool valid = (tl < p.TS) && (g * p.TS + tl < ntap) && (ci0 + ci < p.Cin);
      // dw[co][tap][ci]: lanes of a warp hit consecutive ci -> each red instruction is one 128-B line
      float* dst = dw + ((long long)tap) * p.Cin + ci0 + ci;
      const long long co_stride = (long long)p.taps * p.Cin;
      for (int c = 0; c < p.Cout; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)(g * p.n_cols + c), v);
        if (!valid) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(dst + (long long)(c + j) * co_stride, __uint_as_float(v[j]));
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

static void choose_kbox2(int N, int D, int H, int W, int max_rows, int& bw, int& bh, int& bd, int& bn) {
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

static int pow2_chunk2(int c) { return c >= 64 ? 64 : (c >= 32 ? 32 : 16); }

// returns R2N2_ERR_UNSUPPORTED (without setting an error) when the v1 kernel should be used
int conv_wgrad2_umma(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin,
                     int Cout, int kd, int kh, int kw, int accumulate, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  R2N2_CHECK_ARG(encode != nullptr, R2N2_ERR_CUDA, "conv_wgrad[umma2]: cuTensorMapEncodeTiled unavailable");
  Wgrad2Params p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.kd = kd; p.kh = kh; p.kw = kw; p.taps = kd * kh * kw;
  p.ci_tile = Cin >= 128 ? 128 : Cin;
  p.TS = (128 % p.ci_tile == 0) ? 128 / p.ci_tile : 1;
  if (p.TS > p.taps) p.TS = p.taps;
  p.ci_tiles = (Cin + p.ci_tile - 1) / p.ci_tile;
  p.cw_x = pow2_chunk2(p.ci_tile);
  p.cw_y = pow2_chunk2(Cout);
  p.n_cols = (Cout + 31) / 32 * 32;
  const int total_groups = (p.taps + p.TS - 1) / p.TS;
  int gpc = 512 / p.n_cols;
  if (gpc > total_groups) gpc = total_groups;
  // bytes per K row (pixel): dy chunks + all taps of this CTA
  const int y_row = (Cout + p.cw_y - 1) / p.cw_y * p.cw_y * 2;
  const int x_row = (p.ci_tile + p.cw_x - 1) / p.cw_x * p.cw_x * 2;
  p.stages = 3;
  int max_rows;
  for (;;) {
    const int per_row = y_row + gpc * p.TS * x_row;
    max_rows = (190 * 1024 / p.stages) / per_row / 16 * 16;
    if (max_rows >= 32 || gpc == 1) break;
    --gpc;                                  // fewer tap groups per CTA -> deeper K blocks
  }
  if (max_rows > 128) max_rows = 128;
  R2N2_CHECK_ARG(max_rows >= 16, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma2]: tile does not fit shared memory");
  p.groups_per_cta = gpc;
  const int tap_group_ctas = (total_groups + gpc - 1) / gpc;
  choose_kbox2(N, D, H, W, max_rows, p.bw, p.bh, p.bd, p.bn);
  p.rows = p.bw * p.bh * p.bd * p.bn;
  p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_d = (D + p.bd - 1) / p.bd; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.nbox = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  p.y_bytes = (p.rows * y_row + 1023) / 1024 * 1024;
  p.tap_bytes = p.rows * x_row;
  p.stage_bytes = (p.y_bytes + gpc * p.TS * p.tap_bytes + 1023) / 1024 * 1024;
  p.tmem_cols = 32;
  while (p.tmem_cols < gpc * p.n_cols) p.tmem_cols *= 2;
  p.err_word = get_err_word();
  const long long units = (long long)p.ci_tiles * tap_group_ctas;
  const int ctas_per_sm = (p.tmem_cols <= 256 && p.stages * p.stage_bytes <= 100 * 1024) ? 2 : 1;
  long long splits = ((long long)num_sms() * ctas_per_sm + units - 1) / units;
  if (splits > p.nbox / 4) splits = p.nbox / 4;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.box_per_cta = (p.nbox + splits - 1) / splits;
  splits = (p.nbox + p.box_per_cta - 1) / p.box_per_cta;

  CUtensorMap map_dy, map_x;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2,
                             (cuuint64_t)D * H * W * Cout * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.cw_y, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd,
                         (cuuint32_t)p.bn};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.cw_y * 2),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_wgrad[umma2]: cuTensorMapEncodeTiled(dy) failed (%d)",
                   (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.cw_x, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd,
                         (cuuint32_t)p.bn};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.cw_x * 2),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_wgrad[umma2]: cuTensorMapEncodeTiled(x) failed (%d)",
                   (int)r);
  }
  if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * p.taps * Cin, st);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "conv_wgrad[umma2]: smem attribute: %s",
                   cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + (2 * p.stages + 1) * 8 + 16 + 512 + 1024;
  R2N2_CHECK_ARG(smem_bytes <= 200 * 1024, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma2]: smem %zu too large", smem_bytes);
  if (getenv("R2N2_UMMA_DBG"))
    fprintf(stderr, "[umma wgrad2] Cin %d Cout %d taps %d TS %d groups/cta %d box %dx%dx%dx%d rows %d stage %d smem %zu "
            "grid %lldx%dx%d tmem %d\n", Cin, Cout, p.taps, p.TS, gpc, p.bw, p.bh, p.bd, p.bn, p.rows, p.stage_bytes,
            smem_bytes, splits, p.ci_tiles, tap_group_ctas, p.tmem_cols);
  dim3 grid((unsigned)splits, (unsigned)p.ci_tiles, (unsigned)tap_group_ctas);
  r2n2::launch_k((conv_wgrad2_umma_kernel), dim3(grid), dim3(kUmmaThreads), smem_bytes, (cudaStream_t)(st), 0, map_dy, map_x, dw, p);
  R2N2_CHECK_LAUNCH("conv_wgrad[umma2]");
  return R2N2_OK;
}

}  // namespace r2n2
