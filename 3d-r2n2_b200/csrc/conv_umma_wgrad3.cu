// tcgen05 weight gradient of 3x3x3 convolutions with NARROW channel counts (Cin, Cout <= 64): the
// 32^3 decoder stages conv9b..conv11 (1.18 M pixels, 16..64 channels), sm_100a.
//
//   dw[co][tap][ci] += sum_p dy[p, co] * x[p + off(tap), ci]
//
// The per-tap kernels (conv_umma.cu / conv_umma_wgrad2.cu) re-load a shifted x box for every one of
// the 27 taps (27x L2->SM traffic) and run one narrow MMA per tap; at N <= 64 an SS-mode MMA is bound
// by the 128 B/clk shared-memory operand read (32 + N/4 cycles per K16, tools/exp_desc.cu), not by
// the tensor pipe.  This kernel removes both costs:
//   * FLAT HALO TILES.  A work unit is a band of Y image rows of one (n, z) plane.  TMA loads the dy
//     band and the x band (+1 halo row above/below, plane z+dz) as boxes that include one zero
//     column left and right (x = -1 .. W, out-of-bounds zero fill), so inside shared memory a pixel
//     is the flat row r = y * (W+2) + x + 1 and EVERY (dy, dx) tap is a plain row offset.  Both
//     tensors are read once per dz (3x instead of 27x).
//   * SHIFTED VIEWS.  A tcgen05 shared-memory descriptor may start on any row of a swizzled tile
//     (the swizzle is a function of the absolute address; verified in tools/exp_desc.cu), so the
//     tap-shifted operand is a descriptor, not a copy.
//   * STACKED dx TAPS.  Operands are MN-major (channels contiguous).  The narrower tensor S goes on
//     the M side with the descriptor's leading-dimension stride set to ONE ROW: M chunk j (cs
//     channels) then reads the tile shifted by j pixels, i.e. one MMA computes 128/cs dx-taps at
//     once with all 128 lanes busy.  N = channels of the other tensor.
//   * STACKED dy TAPS (round 2, `nstack`).  The other tensor O = x carries the halo rows, and the N
//     operand's leading-dimension stride is ONE IMAGE ROW (P flat rows): N chunk k reads the x tile
//     shifted by k image rows, so ONE MMA with N = 3 * Cin yields the 3 dy x (128/cs) dx taps of a
//     filter plane: 32 + 3*Cin/4 cycles per K16 instead of 3 x (32 + Cin/4) (SS-operand read bound).
// grid = (pixel splits, 3 dz planes); each CTA accumulates 3 dy x 3 dx taps in TMEM over all its
// units and reduces into dw with fp32 red.global.add at the end.
#include "umma.cuh"

namespace r2n2 {

struct Wgrad3Params {
  int N, D, H, W, Cin, Cout;
  int P, Y, bands, R, nk;                  // row pitch W+2, band height, bands per plane, Y*P, K16 steps
  long long units, units_per_cta;          // units = N * D * bands
  int s_is_x;                              // stacked (M side) tensor: 1 = x, 0 = dy
  int cs, co;                              // channels of the stacked / the other tensor
  int nshift, nmma;                        // dx shifts per MMA (128/cs), MMAs per dy (ceil(3/nshift))
  int nstack;                              // 1: the 3 dy taps are chunks of the N operand (O = x, chunk stride = P rows)
  int acc_cols, tmem_cols;
  int rb_s, rb_o;                          // row bytes
  int s_off, o_off, stage_bytes, stages;   // S tile / O tile offsets inside a stage
  int s_box_bytes, o_box_bytes;
  int* err_word;
};

__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_wgrad3_umma_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                        float* __restrict__ dw, const Wgrad3Params p) {
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int dzi = blockIdx.y;                             // 0..2 -> dz = dzi - 1

  // every byte an MMA can touch must be finite: the slack rows around the tiles stay zero for the
  // whole kernel (TMA only ever writes the tiles themselves)
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = p.stages * p.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

  const long long u_begin = (long long)blockIdx.x * p.units_per_cta;
  long long u_end = u_begin + p.units_per_cta;
  if (u_end > p.units) u_end = p.units;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  // units whose x plane z+dz lies outside the volume contribute nothing: all roles skip them alike
  int n_valid = 0;
  {
    long long u = u_begin;
    int band = (int)(u % p.bands);
    int z = (int)((u / p.bands) % p.D);
    for (; u < u_end; ++u) {
      const int zz = z + dzi - 1;
      if (zz >= 0 && zz < p.D) ++n_valid;
      if (++band == p.bands) { band = 0; if (++z == p.D) z = 0; }
    }
  }

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    long long u = u_begin;
    int band = (int)(u % p.bands);
    long long t = u / p.bands;
    int z = (int)(t % p.D);
    int n = (int)(t / p.D);
    int s = 0;
    uint32_t ph = 0;
    const uint32_t tx = (uint32_t)(p.s_box_bytes + p.o_box_bytes);
    for (; u < u_end; ++u) {
      const int zz = z + dzi - 1;
      if (zz >= 0 && zz < p.D) {
        mbar_wait(empty0 + 8u * s, ph ^ 1u, abort_flag, p.err_word);
        const uint32_t stage = smem_base + (uint32_t)(s * p.stage_bytes);
        const uint32_t bar = full0 + 8u * s;
        const int y0 = band * p.Y;
        if (elect_one()) {
          mbar_expect_tx(bar, tx);
          const uint32_t xdst = stage + (uint32_t)(p.s_is_x ? p.s_off : p.o_off);
          const uint32_t ydst = stage + (uint32_t)(p.s_is_x ? p.o_off : p.s_off);
          tma_load_5d(ydst, &map_dy, bar, 0, -1, y0, z, n);
          tma_load_5d(xdst, &map_x, bar, 0, -1, y0 - 1, zz, n);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (++band == p.bands) { band = 0; if (++z == p.D) { z = 0; ++n; } }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    if (n_valid > 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)((p.nstack ? 3 * p.co : p.co) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int nd = p.nstack ? 1 : 3;                                   // MMA groups per stage along dy
      const uint32_t lt_s = (p.rb_s == 128) ? 2u : (p.rb_s == 64 ? 4u : 6u);
      const uint32_t lt_o = (p.rb_o == 128) ? 2u : (p.rb_o == 64 ? 4u : 6u);
      // stacked operand: leading-dimension (chunk) stride = ONE ROW -> chunk j is the tile shifted by j pixels
      const uint64_t ds_hi = make_mnmajor_desc(0, (uint32_t)p.rb_s, 8u * (uint32_t)p.rb_s, lt_s);
      // other operand, dy-stacked: chunk k of N is the tile shifted by k image rows (P flat rows)
      const uint64_t do_hi = make_mnmajor_desc(0, (p.nstack ? (uint32_t)p.P : 8u) * (uint32_t)p.rb_o,
                                               8u * (uint32_t)p.rb_o, lt_o);
      const uint32_t ks = (uint32_t)p.rb_s, ko = (uint32_t)p.rb_o;       // descriptor step per K16 (16 rows >> 4)
      // row offsets of the views for dy index dyi: the x tile starts one image row above the dy tile
      const int s_dy = p.s_is_x ? p.P : 0, o_dy = p.s_is_x ? 0 : p.P;
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc = 0;
      for (int it = 0; it < n_valid; ++it) {
        mbar_wait(full0 + 8u * s, ph, abort_flag, p.err_word);
        tc_fence_after();
        const uint32_t stage = smem_base + (uint32_t)(s * p.stage_bytes);
        if (elect_one()) {
          for (int dyi = 0; dyi < nd; ++dyi) {
            const uint32_t o_addr = stage + (uint32_t)p.o_off + (uint32_t)(dyi * o_dy * p.rb_o);
            const uint64_t db0 = do_hi | (uint64_t)((o_addr & 0x3FFFF) >> 4);
            for (int m = 0; m < p.nmma; ++m) {
              // first stacked shift of this MMA is (m * nshift - 1) pixels
              const uint32_t s_addr = stage + (uint32_t)p.s_off +
                                      (uint32_t)((dyi * s_dy + m * p.nshift) * p.rb_s) - (uint32_t)p.rb_s;
              const uint64_t da0 = ds_hi | (uint64_t)((s_addr & 0x3FFFF) >> 4);
              const uint32_t d_tmem = tmem_base + (uint32_t)((dyi * p.nmma + m) * p.acc_cols);
              uint32_t a2 = acc;
              for (int k = 0; k < p.nk; ++k) {
                umma_bf16(d_tmem, da0 + (uint64_t)(ks * k), db0 + (uint64_t)(ko * k), idesc, a2);
                a2 = 1;
              }
            }
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
  } else if (n_valid > 0) {
    // ================================ epilogue: TMEM -> red.global.add ========================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int j = row / p.cs, ch = row - j * p.cs;        // stacked chunk (shift within the MMA), channel
    mbar_wait(smem_u32(tmem_full_bar), 0u, abort_flag, p.err_word);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int dyi = 0; dyi < 3; ++dyi)
      for (int m = 0; m < p.nmma; ++m) {
        // accumulator columns of (dy, m): separate accumulators, or the dy-th chunk of MMA m's 3*co columns
        const int acol = p.nstack ? m * p.acc_cols + dyi * p.co : (dyi * p.nmma + m) * p.acc_cols;
        const int sg = m * p.nshift + j;                  // shift index: pixel offset sg - 1 of the stacked tensor
        const bool valid = sg < 3;
        // stacked x: x[p + sg - 1] meets dy[p] -> dx = sg - 1; stacked dy: dy[p + sg - 1] meets x[p] -> dx = 1 - sg
        const int tx = p.s_is_x ? sg : 2 - sg;
        const int tap = (dzi * 3 + dyi) * 3 + tx;
        for (int c = 0; c < p.co; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)(acol + c), v);
          if (!valid) continue;
          if (p.s_is_x) {
            // lanes = ci (contiguous in dw), columns = co
            float* dst = dw + ((long long)c * 27 + tap) * p.Cin + ch;
#pragma unroll
            for (int q = 0; q < 16; ++q) atomicAdd(dst + (long long)q * 27 * p.Cin, __uint_as_float(v[q]));
          } else {
            // lanes = co, columns = ci (contiguous)
            float* dst = dw + ((long long)ch * 27 + tap) * p.Cin + c;
#pragma unroll
            for (int q = 0; q < 16; q += 4)
              red_add_v4(dst + q, __uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                         __uint_as_float(v[q + 3]));
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

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// layout of one stage for a given band height; returns stage bytes
static int plan_stage(Wgrad3Params& p, int Y) {
  p.Y = Y;
  p.R = Y * p.P;
  p.nk = (p.R + 15) / 16;
  const int rpad = p.nk * 16;
  const int x_rows = (Y + 2) * p.P;
  const int s_rows = p.s_is_x ? x_rows : p.R;
  const int o_rows = p.s_is_x ? p.R : x_rows;
  // rows touched (exclusive upper bound, relative to the tile start)
  const int s_reach = (p.s_is_x ? 2 * p.P : 0) - 1 + (p.nmma * p.nshift - 1) + rpad;
  const int o_reach = (p.s_is_x ? 0 : 2 * p.P) + rpad;
  p.s_box_bytes = s_rows * p.rb_s;
  p.o_box_bytes = o_rows * p.rb_o;
  p.s_off = 1024;                                         // zero rows in front of the stacked tile (shift -1)
  p.o_off = p.s_off + round_up((s_reach > s_rows ? s_reach : s_rows) * p.rb_s, 1024);
  p.stage_bytes = p.o_off + round_up((o_reach > o_rows ? o_reach : o_rows) * p.rb_o, 1024);
  return p.stage_bytes;
}

bool conv_wgrad3_supported(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw) {
  auto ok = [](int c) { return c == 16 || c == 32 || c == 64; };
  return kd == 3 && kh == 3 && kw == 3 && ok(Cin) && ok(Cout) && W + 2 <= 256 && N > 0 && D > 0 && H > 0 && W > 0 &&
         !getenv("R2N2_UMMA_NO_WGRAD3");
}

int conv_wgrad3_umma(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin, int Cout,
                     int accumulate, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  R2N2_CHECK_ARG(encode != nullptr, R2N2_ERR_CUDA, "conv_wgrad[umma3]: cuTensorMapEncodeTiled unavailable");
  Wgrad3Params p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.P = W + 2;
  // the narrower tensor is stacked (more dx taps per MMA); ties: x (no band-boundary subtleties)
  p.s_is_x = Cin <= Cout ? 1 : 0;
  if (const char* e = getenv("R2N2_WGRAD3_STACK")) p.s_is_x = atoi(e) ? 1 : 0;
  p.cs = p.s_is_x ? Cin : Cout;
  p.co = p.s_is_x ? Cout : Cin;
  p.rb_s = p.cs * 2;
  p.rb_o = p.co * 2;
  p.nshift = 128 / p.cs;
  p.nmma = (3 + p.nshift - 1) / p.nshift;
  // dy taps stacked on N (one MMA per filter plane and dx group): needs the halo rows on the N side (O = x)
  p.nstack = getenv("R2N2_WGRAD3_NO_NSTACK") ? 0 : 1;
  if (p.nstack && p.s_is_x) {
    p.s_is_x = 0;
    p.cs = Cout; p.co = Cin;
    p.rb_s = p.cs * 2; p.rb_o = p.co * 2;
    p.nshift = 128 / p.cs;
    p.nmma = (3 + p.nshift - 1) / p.nshift;
  }
  p.acc_cols = p.nstack ? 3 * p.co : (p.co < 32 ? 32 : p.co);
  p.tmem_cols = 32;
  while (p.tmem_cols < (p.nstack ? 1 : 3) * p.nmma * p.acc_cols) p.tmem_cols *= 2;
  R2N2_CHECK_ARG(p.tmem_cols <= 512, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma3]: accumulators exceed TMEM");
  // band height: fewest wasted MMA rows / image rows, then the tallest band whose stage is <= 56 KB
  int bestY = 0;
  double best = -1.0;
  for (int Y = 1; Y <= H && Y <= 64; ++Y) {
    if ((Y + 2) > 256) break;
    Wgrad3Params q = p;
    // two CTAs per SM need two 2-stage pipelines in 111 KB; a layer whose accumulators take the whole TMEM (conv9b) is
    // alone on its SM anyway and may use two 100 KB stages (measured: Y = 8 instead of 4, 0.240 -> 0.222 ms)
    if (plan_stage(q, Y) > (p.tmem_cols > 256 ? 100 : 56) * 1024) break;
    const int bands = (H + Y - 1) / Y;
    const double eff = ((double)H / (bands * Y)) * ((double)q.R / (q.nk * 16)) * ((double)Y / (Y + 2) * 0.25 + 0.75);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && Y > bestY)) { best = eff; bestY = Y; }
  }
  if (const char* e = getenv("R2N2_WGRAD3_Y")) { int v = atoi(e); if (v >= 1 && v <= H) bestY = v; }
  R2N2_CHECK_ARG(bestY >= 1, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma3]: no band fits shared memory");
  plan_stage(p, bestY);
  p.bands = (H + p.Y - 1) / p.Y;
  p.units = (long long)N * D * p.bands;
  // two CTAs per SM (two MMA issuers, two TMEM allocations of <= 256 columns) when two 2-stage pipelines fit:
  // measured +26..29 % on conv10a / conv10b / conv11 (one CTA: the issuer idles while a stage is refilled)
  const bool two = p.tmem_cols <= 256 && 2 * p.stage_bytes + 2048 <= 113 * 1024 && !getenv("R2N2_WGRAD3_ONE_CTA");
  p.stages = ((two ? 111 : 200) * 1024) / p.stage_bytes;
  if (p.stages > 4) p.stages = 4;
  if (const char* e = getenv("R2N2_WGRAD3_STAGES")) { int v = atoi(e); if (v >= 2 && v <= p.stages) p.stages = v; }
  R2N2_CHECK_ARG(p.stages >= 2, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma3]: stage too large");
  p.err_word = get_err_word();
  long long splits = (long long)num_sms() * (two ? 2 : 1) / 3;
  if (const char* e = getenv("R2N2_WGRAD3_SPLITS")) { int v = atoi(e); if (v >= 1) splits = v; }
  if (splits > p.units) splits = p.units;
  if (splits < 1) splits = 1;
  p.units_per_cta = (p.units + splits - 1) / splits;
  splits = (p.units + p.units_per_cta - 1) / p.units_per_cta;

  CUtensorMap map_dy, map_x;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2,
                             (cuuint64_t)D * H * W * Cout * 2};
    cuuint32_t box[5] = {(cuuint32_t)Cout, (cuuint32_t)p.P, (cuuint32_t)p.Y, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(Cout * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_wgrad[umma3]: cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)Cin, (cuuint32_t)p.P, (cuuint32_t)(p.Y + 2), 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(Cin * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_wgrad[umma3]: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  }
  if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * 27 * Cin, st);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad3_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "conv_wgrad[umma3]: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
  R2N2_CHECK_ARG(smem_bytes <= 227 * 1024, R2N2_ERR_UNSUPPORTED, "conv_wgrad[umma3]: smem %zu too large", smem_bytes);
  if (getenv("R2N2_UMMA_DBG"))
    fprintf(stderr, "[umma wgrad3] Cin %d Cout %d nstack %d stack %s cs %d nshift %d nmma %d Y %d bands %d R %d nk %d stage %d x %d "
            "units %lld per cta %lld grid %lldx3 tmem %d\n", Cin, Cout, p.nstack, p.s_is_x ? "x" : "dy", p.cs, p.nshift, p.nmma,
            p.Y, p.bands, p.R, p.nk, p.stage_bytes, p.stages, p.units, p.units_per_cta, splits, p.tmem_cols);
  dim3 grid((unsigned)splits, 3);
  r2n2::launch_k((conv_wgrad3_umma_kernel), dim3(grid), dim3(kUmmaThreads), smem_bytes, (cudaStream_t)(st), 0, map_dy, map_x, dw, p);
  R2N2_CHECK_LAUNCH("conv_wgrad[umma3]");
  return R2N2_OK;
}

}  // namespace r2n2
