// Flat-halo tcgen05 implicit-GEMM convolution for 3x3 / 3x3x3 filters (forward and data gradient), sm_100a.
//
//   y[p, co] = epilogue( sum_{tap, ci} x[p + off(tap), ci] * w[co, tap, ci] )       bf16 x bf16 -> fp32
//
// conv_fwd_umma_persistent (conv_umma_fwd.cu) loads one shifted 128-pixel box per filter tap: every
// input element crosses L2 -> shared memory 9 (27) times, and ncu shows the big 3x3 layers bound by
// that traffic (lts throughput 66 %, tensor pipe 30 % on conv1b).  Here the input is loaded ONCE:
//   * a work step is an output block of Yb rows x Xb columns of one (n, z) plane.  TMA loads the
//     input block with a one-pixel halo, (Yb+2) x (Xb+2) pixels, per 64-channel chunk; in shared
//     memory a pixel is the flat row  r = y * P + x  (P = Xb + 2), K-major (channels contiguous);
//   * the outputs of the block are the flat rows 0 .. Yb*P-1 (the two rows-ends x >= Xb are junk and
//     never stored), cut into T tiles of 128 rows.  For tap (ty, tx) the A operand of tile t is the
//     SAME shared-memory plane viewed from row 128 t + ty P + tx: a tcgen05 descriptor may start on
//     any row of a swizzled tile (tools/exp_desc.cu), so a tap costs a descriptor, not a load;
//   * 3-D: the kd = 3 input planes of output plane z are three ring slots; stepping to z+1 loads only
//     plane z+2 (rolling ring, `zc` consecutive planes per work item);
//   * the weights of all taps are streamed through their own small TMA ring once per step and shared
//     by the T tiles of the step (T accumulators side by side in TMEM, double buffered so that the
//     epilogue of step i overlaps the MMAs of step i+1).
// dx-stacked mode (narrow Cout, STK): an SS-mode MMA costs 32 + N/4 cycles (operand reads), so at N = 32 a
// tap runs at 40 % of the tensor peak however well it is fed.  The three taps (ty, 0..2) of a filter row read
// the SAME plane rows shifted by 0/1/2 pixels; their weight tiles lie back to back in shared memory, i.e. they
// ARE one K-major B operand of 3 BN rows.  So ONE MMA with N = 3 BN on the unshifted view computes
//   D_j[r] = sum_{tz,ty,ci} x[r + ty P (+ plane tz), ci] w[co, (tz,ty,j), ci]      j = 0, 1, 2  (column groups)
// and the output is  y[r] = D_0[r] + D_1[r+1] + D_2[r+2]:  the shift moves to the epilogue, where it is a
// warp shuffle across TMEM lanes (+ a 3-value-per-column exchange through shared memory at the 32-row
// warp boundaries).  N = 96 costs 56 cycles instead of 3 x 40.
// Warp roles: 0 = input-plane TMA, 1 = MMA issuer, 2..9 = epilogue, 10 = weight TMA.
#include "umma.cuh"

namespace r2n2 {

constexpr int kFwd3Threads = 96 + 32 * kEpiWarps;   // planes, MMA, 8 epilogue warps, weights

struct Fwd3Params {
  int N, D, H, W, Cin, Cout;
  int kd, kh, kw;                          // kd 1 or 3; (kh, kw) = (3, 3) or (7, 1)
  int w_bytes;                             // bytes of the weight region (ring or resident filter)
  int w_res;                               // all filter taps stay resident in shared memory (loaded once per CTA)
  int Xb, Yb, P, bx, by;                   // block, row pitch Xb+2, blocks per row / column
  int R, T;                                // output rows of a block, 128-row tiles
  int zc, zitems;                          // planes per work item, items along z
  long long items;
  int kbox, nchunk, rb, last_nk16, full_nk16;
  int plane_rows, chunk_bytes, plane_bytes, nring;
  int BN, bn_cols, w_sub_bytes, wg, wt, wstages, w_off;   // wt taps (wg = wt * chunks sub-tiles) per weight stage
  int tmem_cols, flags;
  int no_rot;                              // debugging: all CTAs stream the filter taps in the same order
  int tma_epi, SC, slab_bytes, epi_off;    // epilogue through a dense (Yb x Xb pixels) x SC channels smem slab + TMA
  int nslab;                               // output slabs in rotation (2 when no residual / mask slab is loaded)
  int stk, xch_off, xch_bytes;             // dx-stacked mode: the 3 taps of a filter row share ONE MMA (N = 3 BN)
  int out_bytes;                           // sizeof(output element)
  int ew;                                  // epilogue warps (8, or 16 on the warp-staged bf16 epilogue)
  unsigned P_magic;                        // ceil(2^32 / P): r / P == __umulhi(r, P_magic) for r, P < 2^16
  int bias_off;                            // smem offset of the epilogue warps' bias copy (BN floats)
  int stage_epi, stg_off;                  // warp-staged epilogue: per-warp 32 pixels x 64 bytes staging tiles at this smem offset
  int pair;                                // CTA-pair mode (cluster of 2, cta_group::2 MMAs with M = 256, each CTA holds half of the filter rows)
  int skip;                                // TIMING EXPERIMENTS ONLY (R2N2_FWD3_SKIP bit mask): 1 no plane loads, 2 no weight loads,
                                           // 4 no epilogue work, 8 no MMAs -- results are garbage, the barrier protocol is kept
  int* err_word;
};

struct Item3 {
  int n, y0, x0, z0, nz;
};
static __device__ __forceinline__ Item3 decode_item(const Fwd3Params& p, long long it) {
  Item3 c;
  const int xb = (int)(it % p.bx); it /= p.bx;
  const int yb = (int)(it % p.by); it /= p.by;
  const int zi = (int)(it % p.zitems); it /= p.zitems;
  c.n = (int)it;
  c.x0 = xb * p.Xb;
  c.y0 = yb * p.Yb;
  c.z0 = zi * p.zc;
  c.nz = p.D - c.z0 < p.zc ? p.D - c.z0 : p.zc;
  return c;
}

// CTA-pair mode: the pair walks the items two at a time in lockstep (leader = even item, peer = odd item); an odd
// tail gives the peer a GHOST item (n = N: its TMA loads are zero-filled, its stores clipped, its rows invalid)
template <int PAIR>
static __device__ __forceinline__ Item3 decode_item_pair(const Fwd3Params& p, long long itp, uint32_t crank) {
  if (!PAIR) return decode_item(p, itp);
  const long long it = 2 * itp + crank;
  const bool ghost = it >= p.items;
  Item3 c = decode_item(p, ghost ? p.items - 1 : it);
  if (ghost) c.n = p.N;
  return c;
}
template <int PAIR>
static __device__ __forceinline__ void mma_issue(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR) umma_bf16_2cta(d, a, b, idesc, acc); else umma_bf16(d, a, b, idesc, acc);
}
template <int PAIR>
static __device__ __forceinline__ void mma_commit(uint32_t bar) {
  if (PAIR) umma_commit_2cta(bar); else umma_commit(bar);
}

// KB = K16 steps per full channel chunk (row bytes = 32 KB), NCH = channel chunks, LASTK = K16 steps of
// the last chunk: compile-time so that the MMA issue sequence of one tap is straight-line code with
// immediate descriptor offsets (a single thread issues every MMA of the SM: each extra instruction per
// MMA costs more than the MMA itself at N <= 64).
// TT / WTT / RESS / KDD > 0 (RESS >= 0): tiles per step, taps per weight stage, resident filter and kd are
// compile-time too (3x3 taps) -- the hot configurations get a fully unrolled, branch-free issue sequence
// (ncu: the generic loop spent 158 instructions per tap for 12 MMAs and the issuing thread, not the tensor
// pipe, set the pace of conv1b); 0 / -1 = read from the parameter block.
template <typename TOut, int KB, int NCH, int LASTK, int TT, int WTT, int RESS, int KDD, int STK, int PAIR, int EW>
__global__ void __launch_bounds__(96 + 32 * EW, 1)
conv_fwd3_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_res,
                      const __grid_constant__ CUtensorMap map_mask,
                      const float* __restrict__ bias, const TOut* mask, const TOut* res, TOut* y,
                      const Fwd3Params p) {
  r2n2::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ring_bytes = p.nring * p.plane_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.w_off + p.w_bytes);
  R2N2_WS_BEGIN(bars);
  uint64_t* pfull = bars;                              // [nring]   plane TMA -> MMA
  uint64_t* pempty = pfull + p.nring;                  // [nring]   MMA -> plane TMA
  uint64_t* wfull = pempty + p.nring;                  // [wstages] weight TMA -> MMA
  uint64_t* wempty = wfull + p.wstages;                // [wstages]
  uint64_t* tfull = wempty + p.wstages;                // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;                        // [2]       epilogue -> MMA
  uint64_t* efull = tempty + 2;                        // [1]       residual / mask slab landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(efull + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(!(STK != 0 && PAIR != 0), "the dx-stacked mode has no CTA-pair variant");
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;             // 0 = leader (issues the MMAs of the pair)
  const unsigned cid = PAIR ? (blockIdx.x >> 1) : blockIdx.x, ncta = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const long long pitems = PAIR ? (p.items + 1) / 2 : p.items;
  const int bn_local = PAIR ? (p.BN >> 1) : p.BN;                     // filter rows held by this CTA
  const int T_ = TT > 0 ? TT : p.T;
  const int WT_ = WTT > 0 ? WTT : p.wt;
  const bool RES_ = RESS >= 0 ? (RESS != 0) : (p.w_res != 0);
  // KDD = 71: the 1 x 7 x 1 filter of the row-packed first layer
  const int KD_ = KDD == 71 ? 1 : (KDD > 0 ? KDD : p.kd);
  const int KH_ = KDD == 71 ? 7 : (KDD > 0 ? 3 : p.kh), KW_ = KDD == 71 ? 1 : (KDD > 0 ? 3 : p.kw);

  // rows past the loaded halo block are read by the junk rows of the last tile: keep them finite (zero)
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < ring_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if constexpr (STK != 0) {
    uint4* z = reinterpret_cast<uint4*>(smem + p.xch_off + 2 * p.T * 4 * p.BN * 16);
    for (int i = threadIdx.x; i < p.BN; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    *abort_flag = 0;
    // pair mode: the leader's `full` barriers collect one expect_tx arrival + the TMA bytes of EACH CTA, its
    // `tempty` barriers the epilogue warps of both; `empty` / `tfull` are per CTA (multicast tcgen05.commit)
    const uint32_t np = PAIR ? 2u : 1u;
    for (int s = 0; s < p.nring; ++s) { mbar_init(smem_u32(&pfull[s]), np); mbar_init(smem_u32(&pempty[s]), 1); }
    for (int s = 0; s < p.wstages; ++s) { mbar_init(smem_u32(&wfull[s]), np); mbar_init(smem_u32(&wempty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull[b]), 1); mbar_init(smem_u32(&tempty[b]), np * (uint32_t)EW); }
    mbar_init(smem_u32(efull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();          // both CTAs' barriers initialised and rings zero-filled before any remote arrive / MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  r2n2::pdl_wait();            // prologue done; from here on global memory of the preceding kernels may be touched
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t pfull0 = smem_u32(&pfull[0]), pempty0 = smem_u32(&pempty[0]);
  const uint32_t wfull0 = smem_u32(&wfull[0]), wempty0 = smem_u32(&wempty[0]);
  const uint32_t tfull0 = smem_u32(&tfull[0]), tempty0 = smem_u32(&tempty[0]);
  const int pz = (KD_ - 1) / 2;
  // rotation of the tap order (streamed filters only), in units of one issue group
  const int grp_taps = WT_ * ((WTT == 1 && RESS == 0) ? 3 : 1);       // taps issued per elected region (never split)
  const int rot_taps = (!RES_ && !p.no_rot) ? (int)(cid % (unsigned)((KD_ * KH_ * KW_) / grp_taps)) * grp_taps : 0;

  if (warp == 0) {
    // ================================ input planes (TMA) ====================================
    int slot = 0;
    uint32_t ph = 0;
    const uint32_t tx = (uint32_t)(p.nchunk * p.plane_rows * p.rb);
    for (long long it = cid; it < pitems; it += ncta) {
      const Item3 c = decode_item_pair<PAIR>(p, it, crank);
      const int nplanes = c.nz + KD_ - 1;
      for (int i = 0; i < nplanes; ++i) {
        mbar_wait(pempty0 + 8u * slot, ph ^ 1u, abort_flag, p.err_word);
        const uint32_t bar = PAIR ? mapa_shared(pfull0 + 8u * slot, 0) : pfull0 + 8u * slot;
        const uint32_t dst = smem_base + (uint32_t)(slot * p.plane_bytes);
        if (elect_one()) {
          if (PAIR) mbar_expect_tx_cluster(bar, (p.skip & 1) ? 0u : tx); else mbar_expect_tx(bar, (p.skip & 1) ? 0u : tx);
          if (!(p.skip & 1)) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            if (PAIR)
              tma_load_5d_2sm(dst + (uint32_t)(ch * p.chunk_bytes), &map_a, bar, ch * p.kbox, c.x0 - (KW_ - 1) / 2,
                              c.y0 - (KH_ - 1) / 2, c.z0 + i - pz, c.n);
            else
              tma_load_5d(dst + (uint32_t)(ch * p.chunk_bytes), &map_a, bar, ch * p.kbox, c.x0 - (KW_ - 1) / 2, c.y0 - (KH_ - 1) / 2,
                          c.z0 + i - pz, c.n);
          }
          }
        }
        __syncwarp();
        if (++slot == p.nring) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 2 + EW) {
    // ================================ weights (TMA): wt filter taps per stage ==================
    int slot = 0;
    uint32_t ph = 0;
    const uint32_t stage_tx = (uint32_t)(bn_local * p.rb * NCH * WT_);
    const int wrow0 = PAIR ? (int)crank * bn_local : 0;          // this CTA's half of the filter rows
    const int ntaps = KD_ * KH_ * KW_;
    if (RES_) {
      // small filters: every tap is loaded once and stays in shared memory for the whole kernel
      uint32_t dst = smem_base + (uint32_t)p.w_off;
      if (elect_one()) {
        const uint32_t wbar = PAIR ? mapa_shared(wfull0, 0) : wfull0;
        if (PAIR) mbar_expect_tx_cluster(wbar, (uint32_t)(bn_local * p.rb * NCH * ntaps));
        else mbar_expect_tx(wbar, (uint32_t)(bn_local * p.rb * NCH * ntaps));
        for (int tap = 0; tap < ntaps; ++tap) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            if (PAIR) tma_load_2d_2sm(dst, &map_b, wbar, tap * p.Cin + ch * p.kbox, wrow0);
            else tma_load_2d(dst, &map_b, wbar, tap * p.Cin + ch * p.kbox, 0);
            dst += (uint32_t)p.w_sub_bytes;
          }
        }
      }
      __syncwarp();
    } else
    for (long long it = cid; it < pitems; it += ncta) {
      const Item3 c = decode_item_pair<PAIR>(p, it, crank);
      for (int j = 0; j < c.nz; ++j) {
        for (int tap_o = 0; tap_o < ntaps; tap_o += WT_) {
          // filter-row groups are visited in an order rotated by the CTA index: at any moment the 148 CTAs
          // stream DIFFERENT taps, which spreads the reads of the (small, shared) weight tensor over L2
          int tap = tap_o + rot_taps;
          if (tap >= ntaps) tap -= ntaps;
          mbar_wait(wempty0 + 8u * slot, ph ^ 1u, abort_flag, p.err_word);
          const uint32_t bar = PAIR ? mapa_shared(wfull0 + 8u * slot, 0) : wfull0 + 8u * slot;
          uint32_t dst = smem_base + (uint32_t)(p.w_off + slot * p.wg * p.w_sub_bytes);
          if (elect_one()) {
            if (PAIR) mbar_expect_tx_cluster(bar, (p.skip & 2) ? 0u : stage_tx); else mbar_expect_tx(bar, (p.skip & 2) ? 0u : stage_tx);
            for (int u = 0; u < ((p.skip & 2) ? 0 : WT_); ++u) {
#pragma unroll
              for (int ch = 0; ch < NCH; ++ch) {
                if (PAIR) tma_load_2d_2sm(dst, &map_b, bar, (tap + u) * p.Cin + ch * p.kbox, wrow0);
                else tma_load_2d(dst, &map_b, bar, (tap + u) * p.Cin + ch * p.kbox, 0);
                dst += (uint32_t)p.w_sub_bytes;
              }
            }
          }
          __syncwarp();
          if (++slot == p.wstages) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (pair mode: the leader CTA only) ============
    if (PAIR == 0 || crank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                           ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    constexpr uint32_t kRb = 32u * KB;
    constexpr uint32_t layout_type = (kRb == 128) ? 2u : (kRb == 64 ? 4u : 6u);
    const uint64_t desc_hi = make_kmajor_desc(0, 8u * kRb, layout_type);
    // everything in units of 16 bytes (the descriptor's address granule); smem < 256 KB, so adding
    // offsets to the 14-bit start-address field never carries into the next field
    const uint64_t a_ring = desc_hi | (uint64_t)((smem_base & 0x3FFFF) >> 4);
    const uint64_t b_ring = desc_hi | (uint64_t)(((smem_base + (uint32_t)p.w_off) & 0x3FFFF) >> 4);
    const uint32_t plane16 = (uint32_t)p.plane_bytes >> 4, chunk16 = (uint32_t)p.chunk_bytes >> 4;
    const uint32_t wstage16 = (uint32_t)(p.wg * p.w_sub_bytes) >> 4, wsub16 = (uint32_t)p.w_sub_bytes >> 4;
    constexpr uint32_t tile16 = 128u * kRb / 16u;
    const uint32_t rowP16 = (uint32_t)p.P * kRb / 16u;      // one image row of the block
    constexpr uint32_t row16 = kRb / 16u;                   // one pixel
    int pslot = 0;                 // ring slot of the first plane of the current step
    uint32_t pph = 0;              // phase of the next plane to wait for
    int wait_slot = 0;             // next plane slot to wait on
    int wslot = 0;
    uint32_t wph = 0;
    int step = 0;
    for (long long it = cid; it < pitems; it += ncta) {
      const Item3 c = decode_item_pair<PAIR>(p, it, crank);
      int waited = 0;              // planes of this item already seen full
      pslot = wait_slot;
      for (int j = 0; j < c.nz; ++j, ++step) {
        const int buf = step & 1;
        mbar_wait(tempty0 + 8u * buf, ((uint32_t)(step >> 1) & 1u) ^ 1u, abort_flag, p.err_word);
        while (waited < j + KD_) {
          mbar_wait(pfull0 + 8u * wait_slot, pph, abort_flag, p.err_word);
          if (++wait_slot == p.nring) { wait_slot = 0; pph ^= 1u; }
          ++waited;
        }
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * T_ * p.bn_cols);
        uint32_t acc = 0;
        const int ntaps = KD_ * KH_ * KW_;
        if (RES_ && step == 0) {
          mbar_wait(wfull0, 0u, abort_flag, p.err_word);
          tc_fence_after();
        }
        // GR weight stages per elected region: with one tap per stage (big filters) a whole filter row of three
        // stages is waited for and issued at once, so the per-region overhead (elect, fence, R2UR of the slot
        // addresses, reconvergence) is paid once per 36 MMAs instead of once per 12
        constexpr int GR = (WTT == 1 && RESS == 0) ? 3 : 1;
        int tz = 0, ty = 0, tx = 0;                        // first tap of the group (no divisions in this loop)
        if (rot_taps) {
          tz = rot_taps / (KH_ * KW_);
          const int rem = rot_taps - tz * KH_ * KW_;
          ty = rem / KW_; tx = rem - ty * KW_;
        }
        if constexpr (STK != 0) {
          // ---- dx-stacked issue: one MMA (N = 3 BN) per filter row (tz, ty), K16 step and tile ----
          static_assert(NCH == 1, "stacked mode: one channel chunk");
          const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * p.BN) >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
          const int ngrp = KD_ * 3;                          // filter rows
          const int gps = RES_ ? ngrp : WT_ / 3;             // filter rows per elected region (= weight stage)
          int g0 = rot_taps / 3;
          for (int go = 0; go < ngrp; go += gps) {
            int slot_w = 0;
            if (!RES_) {
              slot_w = wslot;
              mbar_wait(wfull0 + 8u * wslot, wph, abort_flag, p.err_word);
              if (++wslot == p.wstages) { wslot = 0; wph ^= 1u; }
              tc_fence_after();
            }
            if (elect_one()) {
              int g = g0;
              uint64_t b_g = b_ring + (uint64_t)(RES_ ? (uint32_t)g * 3u * wsub16 : (uint32_t)slot_w * wstage16);
              for (int u = 0; u < gps; ++u) {
                const int tz = g / 3, ty = g - tz * 3;
                int ps = pslot + tz;
                if (ps >= p.nring) ps -= p.nring;
                uint64_t a_t = a_ring + (uint64_t)((uint32_t)ps * plane16 + (uint32_t)ty * rowP16);
                uint32_t d_t = d0;
                for (int t = 0; t < T_; ++t, a_t += tile16, d_t += (uint32_t)p.bn_cols) {
#pragma unroll
                  for (int k = 0; k < LASTK; ++k)
                    umma_bf16(d_t, a_t + (uint64_t)(2u * k), b_g + (uint64_t)(2u * k), idesc_s, k == 0 ? acc : 1u);
                }
                acc = 1;
                b_g += (uint64_t)(3u * wsub16);
                if (++g == ngrp) { g = 0; if (RES_) b_g = b_ring; }
              }
              if (!RES_) umma_commit(wempty0 + 8u * slot_w);
            }
            __syncwarp();
            acc = 1;
            g0 += gps;
            if (g0 >= ngrp) g0 -= ngrp;
          }
        } else
        for (int tap_o = 0; tap_o < ntaps; tap_o += WT_ * GR) {  // one weight stage = wt taps of ONE plane
          int tap = tap_o + rot_taps;
          if (tap >= ntaps) tap -= ntaps;
          int ps = pslot + tz;
          if (ps >= p.nring) ps -= p.nring;
          int slots[GR];
          if (!RES_) {
#pragma unroll
            for (int g = 0; g < GR; ++g) {
              slots[g] = wslot;
              mbar_wait(wfull0 + 8u * wslot, wph, abort_flag, p.err_word);
              if (++wslot == p.wstages) { wslot = 0; wph ^= 1u; }
            }
            tc_fence_after();
          }
          if (elect_one()) {
            uint64_t a_tap = a_ring + (uint64_t)((uint32_t)ps * plane16 + (uint32_t)ty * rowP16 + (uint32_t)tx * row16);
            int txu = tx;
#pragma unroll
            for (int g = 0; g < GR; ++g) {
              uint64_t b_tap = b_ring + (uint64_t)(RES_ ? (uint32_t)tap * (NCH * wsub16) : (uint32_t)slots[g] * wstage16);
              for (int u = 0; u < WT_; ++u) {
                uint64_t a_t = a_tap;
                uint32_t d_t = d0;
                for (int t = 0; t < ((p.skip & 8) ? 0 : T_); ++t, a_t += tile16, d_t += (uint32_t)p.bn_cols) {
#pragma unroll
                  for (int ch = 0; ch < NCH; ++ch) {
                    const int nk = (ch == NCH - 1) ? LASTK : KB;
#pragma unroll
                    for (int k = 0; k < KB; ++k) {
                      if (k < nk)
                        mma_issue<PAIR>(d_t, a_t + (uint64_t)(ch * chunk16 + 2u * k), b_tap + (uint64_t)(ch * wsub16 + 2u * k),
                                        idesc, (ch == 0 && k == 0) ? acc : 1u);
                    }
                  }
                }
                acc = 1;
                b_tap += (uint64_t)(NCH * wsub16);
                a_tap += row16;
                if (++txu == KW_) { txu = 0; a_tap += (uint64_t)(rowP16 - (uint32_t)KW_ * row16); }
              }
              if (!RES_) mma_commit<PAIR>(wempty0 + 8u * slots[g]);
            }
          }
          __syncwarp();
          acc = 1;
          tx += WT_ * GR;
          while (tx >= KW_) { tx -= KW_; ++ty; }
          if (ty >= KH_) { ty = 0; ++tz; }
          if (tz >= KD_) tz = 0;                           // (rotated order wraps around)
        }
        if (elect_one()) {
          mma_commit<PAIR>(tfull0 + 8u * buf);            // accumulators of this step complete
          mma_commit<PAIR>(pempty0 + 8u * pslot);         // plane z-1 is not needed by later steps
          if (j == c.nz - 1)
            for (int e = 1; e < KD_; ++e) {
              int ps = pslot + e;
              if (ps >= p.nring) ps -= p.nring;
              mma_commit<PAIR>(pempty0 + 8u * ps);
            }
        }
        __syncwarp();
        if (++pslot == p.nring) pslot = 0;
      }
    }
    }
  } else if (warp >= 2 && warp < 2 + EW) {
    // ================================ epilogue ==============================================
    const int quad = warp & 3;
    // bias in shared memory (every epilogue variant): with ~220 KB of the SM carved out as shared memory the L1 keeps
    // nothing, so the bias loads of every chunk / round came from L2 (~250 cycles; ncu source view of conv1a: 11 % of the
    // kernel's stall samples sat on the first FADD behind them).  conv1a 0.264 -> 0.248 ms, conv1b 0.431 -> 0.422 ms.
    float* bias_s = reinterpret_cast<float*>(smem + p.bias_off);
    if (p.flags & R2N2_EPI_BIAS)
      for (int i = (warp - 2) * 32 + lane; i < p.BN; i += 32 * EW) bias_s[i] = bias[i];
    named_bar_sync(3, 32 * EW);
    if (p.tma_epi) {
      // TMA epilogue (see conv_umma_fwd.cu): the block's outputs are gathered DENSELY (junk columns dropped)
      // into a swizzled smem slab of Yb x Xb pixels x SC channels; residual / mask slabs arrive by TMA load,
      // the result leaves by one TMA store.  Warp (quadrant, half) takes the tiles t = half, half+2, ...
      if constexpr (sizeof(TOut) == 2) {
        const int half = (warp - 2) >> 2;
        const bool leader = warp == 2 && lane == 0;
        const uint32_t bufY = smem_base + (uint32_t)p.epi_off;
        const uint32_t bufM = bufY + (uint32_t)p.slab_bytes;       // (mask slab; second output slab when nslab == 2)
        const uint32_t ef = smem_u32(efull);
        const uint32_t rb = (uint32_t)p.SC * 2u;
        const uint32_t smask = rb == 128 ? 7u : (rb == 64 ? 3u : 1u);
        const bool has_in = (p.flags & (R2N2_EPI_RESIDUAL | R2N2_EPI_MASK)) != 0;
        const uint32_t in_tx = (uint32_t)(p.Xb * p.Yb * p.SC * 2) *
                               (((p.flags & R2N2_EPI_RESIDUAL) ? 1u : 0u) + ((p.flags & R2N2_EPI_MASK) ? 1u : 0u));
        uint32_t eph = 0;
        bool pending = false;
        int step = 0;
        // without input slabs two output slabs rotate: the wait is for the store issued TWO slabs ago, so the
        // shared-memory read of one slab's store overlaps the next slab's TMEM loads and packing
        const bool two = p.nslab == 2;
        const uint32_t bufY0 = bufY;
        uint32_t sidx = 0;
        for (long long it = cid; it < pitems; it += ncta) {
          const Item3 c = decode_item_pair<PAIR>(p, it, crank);
          for (int j = 0; j < c.nz; ++j, ++step) {
            const int buf = step & 1;
            const int zz = c.z0 + j;
            for (int sc = 0; sc < p.BN; sc += p.SC, ++sidx) {
              const uint32_t bufY = two ? bufY0 + (sidx & 1u) * (uint32_t)p.slab_bytes : bufY0;
              if (leader && pending) {
                if (two) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else tma_store_wait_read();
              }
              named_bar_sync(1, 32 * EW);
              if (has_in && leader) {
                mbar_expect_tx(ef, in_tx);
                if (p.flags & R2N2_EPI_RESIDUAL) tma_load_5d(bufY, &map_res, ef, sc, c.x0, c.y0, zz, c.n);
                if (p.flags & R2N2_EPI_MASK) tma_load_5d(bufM, &map_mask, ef, sc, c.x0, c.y0, zz, c.n);
              }
              if (sc == 0) {
                mbar_wait(tfull0 + 8u * buf, (uint32_t)(step >> 1) & 1u, abort_flag, p.err_word);
                tc_fence_after();
              }
              if (has_in) {
                mbar_wait(ef, eph, abort_flag, p.err_word);
                eph ^= 1u;
              }
              for (int t = half; t < ((p.skip & 4) ? 0 : T_); t += 2) {
                const int r = t * 128 + quad * 32 + lane;
                const int yy = r / p.P, xx = r - yy * p.P;
                const bool valid = r < p.R && xx < p.Xb;
                const uint32_t drow = (uint32_t)(yy * p.Xb + xx) * rb;       // dense row in the slab
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * T_ + t) * p.bn_cols);
                for (int cc = 0; cc < p.SC; cc += 16) {
                  uint32_t v[16];
                  tmem_ld16(taddr + (uint32_t)(sc + cc), v);
                  if (!valid) continue;
                  float f[16];
#pragma unroll
                  for (int q = 0; q < 16; ++q) f[q] = __uint_as_float(v[q]);
                  if (p.flags & R2N2_EPI_BIAS) {
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + sc + cc + q);
                      f[q] += b4.x; f[q + 1] += b4.y; f[q + 2] += b4.z; f[q + 3] += b4.w;
                    }
                  }
                  if (p.flags & R2N2_EPI_LRELU) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] = lrelu(f[q]);
                  }
                  const uint32_t o0 = swz_off(drow + (uint32_t)cc * 2u, smask);
                  const uint32_t o1 = swz_off(drow + (uint32_t)cc * 2u + 16u, smask);
                  if (p.flags & R2N2_EPI_RESIDUAL) {
                    Row16<__nv_bfloat16> r16;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r16.a.x), "=r"(r16.a.y), "=r"(r16.a.z), "=r"(r16.a.w) : "r"(bufY + o0));
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r16.b.x), "=r"(r16.b.y), "=r"(r16.b.z), "=r"(r16.b.w) : "r"(bufY + o1));
                    float rr[16];
                    r16.unpack(rr);
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] += rr[q];
                  }
                  if (p.flags & R2N2_EPI_MASK) {
                    Row16<__nv_bfloat16> m16;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(m16.a.x), "=r"(m16.a.y), "=r"(m16.a.z), "=r"(m16.a.w) : "r"(bufM + o0));
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(m16.b.x), "=r"(m16.b.y), "=r"(m16.b.z), "=r"(m16.b.w) : "r"(bufM + o1));
                    float mm[16];
                    m16.unpack(mm);
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] *= lrelu_slope(mm[q]);
                  }
                  uint32_t w[8];
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
                    w[q] = *reinterpret_cast<uint32_t*>(&h);
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(bufY + o0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(bufY + o1), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
                }
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              named_bar_sync(1, 32 * EW);
              if (leader && !(p.skip & 4)) {
                tma_store_5d(&map_y, bufY, sc, c.x0, c.y0, zz, c.n);
                pending = true;
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(mapa_shared(tempty0 + 8u * buf, 0));
              else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * buf) : "memory");
            }
          }
        }
        if (leader && pending) tma_store_wait_all();
      }
    } else if constexpr (STK != 0) {
    // ---- dx-stacked register epilogue: y[r] = D_0[r] + D_1[r+1] + D_2[r+2] -------------------------------
    const int nch = p.BN >> 4, half = (warp - 2) >> 2;
    const int c_begin = half ? ((nch + 1) >> 1) << 4 : 0, c_end = half ? p.BN : ((nch + 1) >> 1) << 4;
    // exchange buffer [2 step parities][T][4 quadrants][BN] x 16 bytes {D2[row 1], D1[row 0], D2[row 0], -}, then
    // one all-zero block [BN] x 16 bytes (the "neighbour" of the last quadrant of the last tile):
    // lane 31 reads the first pair (e0 = its r+2, e1 = its r+1 neighbour), lane 30 the second (e0 = its r+2)
    const uint32_t xch = smem_base + (uint32_t)p.xch_off;
    const uint32_t xstep = (uint32_t)(T_ * 4 * p.BN) * 16u;
    const uint32_t lane_sel = lane == 30 ? 8u : 0u;
    int step = 0;
    for (long long it = cid; it < pitems; it += ncta) {
      const Item3 c = decode_item_pair<PAIR>(p, it, crank);
      for (int j = 0; j < c.nz; ++j, ++step) {
        const int buf = step & 1;
        const int zz = c.z0 + j;
        const uint32_t xb = xch + (uint32_t)(step & 1) * xstep;
        mbar_wait(tfull0 + 8u * buf, (uint32_t)(step >> 1) & 1u, abort_flag, p.err_word);
        tc_fence_after();
        // phase A: rows 0 and 1 of every (tile, quadrant) are the r+1 / r+2 neighbours of rows 30, 31 of the
        // quadrant before: they go through shared memory (buffer of this step's parity)
        for (int t = 0; t < T_; ++t) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * T_ + t) * p.bn_cols);
          for (int cc = c_begin; cc < c_end; cc += 16) {
            uint32_t v1[16], v2[16];
            tmem_ld16_nowait(taddr + (uint32_t)(p.BN + cc), v1);
            tmem_ld16_nowait(taddr + (uint32_t)(2 * p.BN + cc), v2);
            tmem_wait_ld();
            // lane 1 stores D2[row 1] at +0, lane 0 stores D1[row 0] at +4 and D2[row 0] at +8
            const uint32_t dst = xb + (uint32_t)((t * 4 + quad) * p.BN + cc) * 16u + (lane == 0 ? 4u : 0u);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              if (lane < 2)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 16u * q), "r"(lane == 0 ? v1[q] : v2[q]) : "memory");
              if (lane == 0)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 16u * q + 4u), "r"(v2[q]) : "memory");
            }
          }
        }
        named_bar_sync(1, 32 * EW);
        for (int t = 0; t < T_; ++t) {
          const int r = t * 128 + quad * 32 + lane;
          const int yy = r / p.P, xx = r - yy * p.P;
          const int gy = c.y0 + yy, gx = c.x0 + xx;
          const bool valid = r < p.R && xx < p.Xb && gx < p.W && gy < p.H;
          const long long off = ((((long long)c.n * p.D + zz) * p.H + gy) * p.W + gx) * p.Cout;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * T_ + t) * p.bn_cols);
          int nq = quad + 1, nt = t;
          if (nq == 4) { nq = 0; ++nt; }
          const bool has_next = nt < T_;                   // (rows past the last tile are junk rows' neighbours)
          const uint32_t nb = (has_next ? xb + (uint32_t)((nt * 4 + nq) * p.BN) * 16u : xch + 2u * xstep) + lane_sel;
          const bool use1 = lane == 31, use2 = lane >= 30;
          for (int cc = c_begin; cc < c_end; cc += 16) {
            EpiPrefetch<TOut> pf;
            epi_prefetch(pf, p.flags, valid, res + off, mask + off, cc, p.Cout);
            uint32_t v0[16], v1[16], v2[16];
            tmem_ld16_nowait(taddr + (uint32_t)cc, v0);
            tmem_ld16_nowait(taddr + (uint32_t)(p.BN + cc), v1);
            tmem_ld16_nowait(taddr + (uint32_t)(2 * p.BN + cc), v2);
            tmem_wait_ld();
            float f[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              // straight-line: every lane loads a pair (volatile asm: the compiler must not wrap it in a branch)
              float e0, e1;
              asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(e0), "=f"(e1) : "r"(nb + (uint32_t)(cc + q) * 16u));
              float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[q]), 1);
              float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[q]), 2);
              a1 = use1 ? e1 : a1;
              a2 = use2 ? e0 : a2;
              f[q] = __uint_as_float(v0[q]) + a1 + a2;
            }
            if (!valid) continue;
            if ((p.flags & (R2N2_EPI_RESIDUAL | R2N2_EPI_RES_PRE)) == (R2N2_EPI_RESIDUAL | R2N2_EPI_RES_PRE)) {
              float rr[16];
              pf.r.unpack(rr);
#pragma unroll
              for (int q = 0; q < 16; ++q) f[q] += rr[q];
            }
            if (p.flags & R2N2_EPI_BIAS) {
#pragma unroll
              for (int q = 0; q < 16; q += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc + q);
                f[q] += b4.x; f[q + 1] += b4.y; f[q + 2] += b4.z; f[q + 3] += b4.w;
              }
            }
            if (p.flags & R2N2_EPI_LRELU) {
#pragma unroll
              for (int q = 0; q < 16; ++q) f[q] = lrelu(f[q]);
            }
            if ((p.flags & (R2N2_EPI_RESIDUAL | R2N2_EPI_RES_PRE)) == R2N2_EPI_RESIDUAL) {
              float rr[16];
              pf.r.unpack(rr);
#pragma unroll
              for (int q = 0; q < 16; ++q) f[q] += rr[q];
            }
            if (p.flags & R2N2_EPI_MASK) {
              float mm[16];
              pf.m.unpack(mm);
#pragma unroll
              for (int q = 0; q < 16; ++q) f[q] *= lrelu_slope(mm[q]);
            }
            store16(y + off + cc, f);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(tempty0 + 8u * buf, 0));
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * buf) : "memory");
        }
      }
    }
    } else if (p.stage_epi) {
    // ---- warp-staged epilogue ---------------------------------------------------------------------------
    // A thread owns one accumulator row (= one pixel); writing its channels with 16-byte accesses makes every
    // store / load instruction of the warp touch 32 different 128-byte lines (32 L1 wavefronts).  Here the warp
    // exchanges a ROUND of RB bytes per pixel through a private, XOR-swizzled shared-memory tile (32 pixels x RB)
    // and accesses global memory as (32 / NU) pixels x RB contiguous bytes per instruction (full 32-byte sectors,
    // 1/4 - 1/2 of the wavefronts).  Warp-local only: no CTA barrier, no TMA store whose shared-memory read would
    // have to drain before the next slab (what bounded the TMA epilogue).
    // EW = 8: RB = 64 (32 bf16 / 16 fp32 channels), two warps per TMEM lane quadrant.  EW = 16 (bf16 outputs): RB = 32
    // (16 channels, one tcgen05.ld), four warps per quadrant -- the epilogue is latency-bound (IPC ~0.4 of 4 with
    // two warps per scheduler), so twice the warps on half the registers each move twice the data.
    constexpr int RB = EW == 16 ? 32 : 64;                // bytes per pixel and round
    constexpr int RC = RB / (int)sizeof(TOut);            // channels per round
    constexpr int NU = RB / 16;                           // 16-byte units per pixel and round = global accesses per round
    constexpr int PPA = 32 / NU;                          // pixels per global access
    static_assert(EW == 8 || sizeof(TOut) == 2, "16 epilogue warps: bf16 outputs");
    const int nrounds = p.BN / RC;
    const unsigned nr_magic = (unsigned)((0x100000000ULL + (unsigned)nrounds - 1) / (unsigned)nrounds);
    const int sub = (warp - 2) >> 2;
    constexpr int nsubw = EW / 4;                         // this warp among the warps of its quadrant
    const uint32_t stg = smem_base + (uint32_t)p.stg_off + (uint32_t)(warp - 2) * (32u * RB);
    // XOR swizzle of the 16-byte unit index: conflict-free for the owner pattern (32 rows, fixed unit) and for the
    // global-access pattern (PPA rows x NU units)
    auto swz = [&](uint32_t row, uint32_t u) -> uint32_t {
      return RB == 64 ? (u ^ ((row >> 1) & 3u)) : (u ^ ((row >> 2) & 1u));
    };
    const uint32_t my_row = stg + (uint32_t)lane * RB;
    const int q_lo = lane / NU;                           // pixel (of PPA) this lane moves in each global access
    const uint32_t u_g = (uint32_t)(lane % NU);           // its 16-byte unit
    const bool has_res = (p.flags & R2N2_EPI_RESIDUAL) != 0, has_mask = (p.flags & R2N2_EPI_MASK) != 0;
    const bool res_pre = (p.flags & R2N2_EPI_RES_PRE) != 0;
    int step = 0;
    for (long long it = cid; it < pitems; it += ncta) {
      const Item3 c = decode_item_pair<PAIR>(p, it, crank);
      for (int j = 0; j < c.nz; ++j, ++step) {
        const int buf = step & 1;
        const int zz = c.z0 + j;
        bool first = true;
        // work items (tile, round) dealt round-robin to the warps of a quadrant; whole tiles when that balances
        const int nitems = T_ * nrounds;
        const bool by_tile = (T_ % nsubw) == 0;
        for (int wi0 = sub; wi0 < ((p.skip & 4) ? 0 : nitems); wi0 += nsubw) {
          // by_tile: warp `sub` takes the tiles t = sub, sub + nsubw, ... with all their rounds
          // (x / nrounds by multiply-high: the run-time divisions were ~5 % of conv1a's stall samples in ncu's source view)
          const int q_ = by_tile ? wi0 / nsubw : wi0;                            // nsubw is a compile-time power of two
          const int qd = nrounds == 1 ? q_ : (int)__umulhi((unsigned)q_, nr_magic);   // q_ / nrounds, exact for q_ < 2^16
                                                                                // (nrounds = 1: the magic number is 2^32)
          const int t = by_tile ? qd * nsubw + sub : qd;
          const int rd = q_ - qd * nrounds;
          R2N2_WS_T0();
          // the NU pixels whose 16-byte unit u_g this lane loads / stores: rows q = PPA k + q_lo of the warp's 32
          long long goff0[NU];
          bool gok[NU];
#pragma unroll
          for (int k = 0; k < NU; ++k) {
            const unsigned r = (unsigned)(t * 128 + quad * 32 + PPA * k + q_lo);
            const int yy = (int)__umulhi(r, (unsigned)p.P_magic), xx = (int)r - yy * p.P;     // r / P (exact for r, P < 2^16)
            const int gy = c.y0 + yy, gx = c.x0 + xx;
            gok[k] = (int)r < p.R && xx < p.Xb && gx < p.W && gy < p.H && c.n < p.N;
            goff0[k] = ((((long long)c.n * p.D + zz) * p.H + gy) * p.W + gx) * (long long)(p.Cout * (int)sizeof(TOut)) +
                       (long long)(rd * RB) + (long long)u_g * 16;
          }
          R2N2_WS_ACC(30);                       // geometry
          uint4 gin[2][NU];
          if (has_res || has_mask) {
#pragma unroll
            for (int k = 0; k < NU; ++k) {
              if (gok[k]) {
                if (has_res) gin[0][k] = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(res) + goff0[k]);
                if (has_mask) gin[1][k] = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(mask) + goff0[k]);
              }
            }
          }
          R2N2_WS_ACC(31);                       // input loads issued
          if (first) {
            mbar_wait(tfull0 + 8u * buf, (uint32_t)(step >> 1) & 1u, abort_flag, p.err_word);
            tc_fence_after();
            first = false;
          }
          R2N2_WS_ACC(32);                       // accumulator wait
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * T_ + t) * p.bn_cols + rd * RC);
          float f[RC];
          if (p.skip & 16) {
#pragma unroll
            for (int q = 0; q < RC; ++q) f[q] = 0.f;
          } else {
            uint32_t v[16];
            tmem_ld16_nowait(taddr, v);
            if (RC == 32) {
              uint32_t v2[16];
              tmem_ld16_nowait(taddr + 16u, v2);
              tmem_wait_ld();
#pragma unroll
              for (int q = 0; q < 16; ++q) f[(RC == 32 ? 16 : 0) + q] = __uint_as_float(v2[q]);
            } else {
              tmem_wait_ld();
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) f[q] = __uint_as_float(v[q]);
          }
          R2N2_WS_ACC(33);                       // TMEM loads
          // an input of this round: global (PPA pixels x RB bytes per access, issued above) -> staging tile -> the pixel's
          // owner, folded into f at once (op 0: f += v, op 1: f *= LeakyReLU'(v)) so that no second row stays live
          auto fold = [&](const uint4 (&g4)[NU], int op) {
#pragma unroll
            for (int k = 0; k < NU; ++k) {
              const uint32_t row = (uint32_t)(PPA * k + q_lo);
              const uint32_t a = stg + row * RB + (swz(row, u_g) << 4);
              const uint4 g = gok[k] ? g4[k] : make_uint4(0, 0, 0, 0);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(g.x), "r"(g.y), "r"(g.z), "r"(g.w) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < NU; ++u) {
              uint4 o;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                           : "r"(my_row + (swz((uint32_t)lane, (uint32_t)u) << 4)));
              const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
              if (sizeof(TOut) == 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                  const int q = (u * 8 + 2 * e) % RC;
                  if (op == 0) { f[q] += t2.x; f[q + 1] += t2.y; }
                  else { f[q] *= lrelu_slope(t2.x); f[q + 1] *= lrelu_slope(t2.y); }
                }
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float tv = __uint_as_float(w4[e]);
                  const int q = (u * 4 + e) % RC;
                  if (op == 0) f[q] += tv; else f[q] *= lrelu_slope(tv);
                }
              }
            }
            __syncwarp();
          };
          if (has_res && res_pre) fold(gin[0], 0);
          if (p.flags & R2N2_EPI_BIAS) {
#pragma unroll
            for (int q = 0; q < RC; q += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + rd * RC + q);
              f[q] += b4.x; f[q + 1] += b4.y; f[q + 2] += b4.z; f[q + 3] += b4.w;
            }
          }
          if (p.flags & R2N2_EPI_LRELU) {
#pragma unroll
            for (int q = 0; q < RC; ++q) f[q] = lrelu(f[q]);
          }
          if (has_res && !res_pre) fold(gin[0], 0);
          if (has_mask) fold(gin[1], 1);
          R2N2_WS_ACC(34);                       // inputs folded, bias, LeakyReLU
          // results: owner -> staging tile -> global
          uint32_t w[4 * NU];
          if (sizeof(TOut) == 2) {
#pragma unroll
            for (int q = 0; q < 4 * NU; ++q) {
              __nv_bfloat162 h = __floats2bfloat162_rn(f[(2 * q) % RC], f[(2 * q + 1) % RC]);
              w[q] = *reinterpret_cast<uint32_t*>(&h);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4 * NU; ++q) w[q] = __float_as_uint(f[q % RC]);
          }
#pragma unroll
          for (int u = 0; u < NU; ++u)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my_row + (swz((uint32_t)lane, (uint32_t)u) << 4)), "r"(w[4 * u]),
                         "r"(w[4 * u + 1]), "r"(w[4 * u + 2]), "r"(w[4 * u + 3])
                         : "memory");
          __syncwarp();
#pragma unroll
          for (int k = 0; k < NU; ++k) {
            const uint32_t row = (uint32_t)(PPA * k + q_lo);
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(stg + row * RB + (swz(row, u_g) << 4)));
            if (gok[k] && !(p.skip & 32)) *reinterpret_cast<uint4*>(reinterpret_cast<char*>(y) + goff0[k]) = o;
          }
          __syncwarp();
          R2N2_WS_ACC(35);                       // pack, staging exchange, global stores
        }
        if (first) {          // (no work item for this warp in this step: still a consumer of the accumulator barrier)
          mbar_wait(tfull0 + 8u * buf, (uint32_t)(step >> 1) & 1u, abort_flag, p.err_word);
          tc_fence_after();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(tempty0 + 8u * buf, 0));
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * buf) : "memory");
        }
      }
    }
    } else {
    const int nch = p.BN >> 4, half = (warp - 2) >> 2;     // two warps per lane quadrant, half the columns each
    const int c_begin = half ? ((nch + 1) >> 1) << 4 : 0, c_end = half ? p.BN : ((nch + 1) >> 1) << 4;
    int step = 0;
    for (long long it = cid; it < pitems; it += ncta) {
      const Item3 c = decode_item_pair<PAIR>(p, it, crank);
      for (int j = 0; j < c.nz; ++j, ++step) {
        const int buf = step & 1;
        const int zz = c.z0 + j;
        bool first = true;
        if (p.skip & 4) { mbar_wait(tfull0 + 8u * buf, (uint32_t)(step >> 1) & 1u, abort_flag, p.err_word); first = false; }
        for (int t = 0; t < ((p.skip & 4) ? 0 : T_); ++t) {
          const int r = t * 128 + quad * 32 + lane;
          const int yy = r / p.P, xx = r - yy * p.P;
          const int gy = c.y0 + yy, gx = c.x0 + xx;
          const bool valid = r < p.R && xx < p.Xb && gx < p.W && gy < p.H && c.n < p.N;
          const long long off = ((((long long)c.n * p.D + zz) * p.H + gy) * p.W + gx) * p.Cout;
          EpiPrefetch<TOut> pf;
          epi_prefetch(pf, p.flags, valid, res + off, mask + off, c_begin, p.Cout);
          if (first) {
            mbar_wait(tfull0 + 8u * buf, (uint32_t)(step >> 1) & 1u, abort_flag, p.err_word);
            tc_fence_after();
            first = false;
          }
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * T_ + t) * p.bn_cols);
          epi_row(pf, taddr, c_begin, c_end, 0, p.Cout, valid, false, p.flags, bias_s, res + off, mask + off, y + off);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(tempty0 + 8u * buf, 0));
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * buf) : "memory");
        }
      }
    }
    }
  }
  R2N2_WS_END(p.err_word);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();          // no CTA leaves (or frees TMEM) while its peer may still signal it / read its smem
  if (warp == 0) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

static inline int ru(int v, int m) { return (v + m - 1) / m * m; }

// Choose the block (Xb, Yb), ring depth and weight staging.  Returns false when the layer does not fit
// (caller falls back to the per-tap kernel).
// want_tma_epi: 0 never, 1 if it pays (heuristic), 2 always (bf16 outputs only; decided by the caller)
static bool plan_fwd3(Fwd3Params& p, int want_tma_epi = 0, int want_stk = 1) {
  int budget = 222 * 1024 - 512;          // (512 bytes: the epilogue warps' bias copy, Fwd3Params::bias_off)
  p.tma_epi = 0; p.SC = 32; p.slab_bytes = 0; p.epi_off = 0; p.nslab = 1;
  // channel chunk = TMA box / swizzle width; 96 = 3 x 32 (no zero-filled half chunk: 25 % less smem + traffic)
  p.kbox = p.Cin % 64 == 0 ? 64 : (p.Cin % 32 == 0 ? 32 : 16);
  p.nchunk = (p.Cin + p.kbox - 1) / p.kbox;
  p.rb = p.kbox * 2;
  p.full_nk16 = p.kbox >> 4;
  p.last_nk16 = (p.Cin - (p.nchunk - 1) * p.kbox) >> 4;
  p.BN = p.Cout;
  p.w_sub_bytes = ru((p.pair ? p.BN / 2 : p.BN) * p.rb, 1024);     // pair mode: each CTA holds half of the filter rows
  // dx-stacked mode (see the header): 3x3 rows, one channel chunk, the three weight tiles of a filter row
  // contiguous in shared memory (no padding between them), 3 BN accumulator columns per tile
  p.stk = (p.kh == 3 && p.kw == 3 && p.nchunk == 1 && 3 * p.BN <= 256 && p.BN * p.rb == p.w_sub_bytes &&
           !getenv("R2N2_FWD3_NO_STK") && !p.pair) ? 1 : 0;
  // Measured (B200, 36 x 32^3): the stacked epilogue (2 shuffles + 1 shared load per element, boundary exchange)
  // costs ~2x the plain one, so stacking pays where a tile carries enough MMA work to hide it: Cin = 64
  // (conv10a 0.282 -> 0.185 ms, conv9b 0.341 -> 0.323 ms) and the N = 16 layer (conv11 0.115 -> 0.102 ms);
  // at Cin <= 32 with N = 32 the epilogue binds (conv10b 0.134 -> 0.149 ms).  R2N2_FWD3_STK=0/1 forces it.
  if (p.stk && !(p.Cin >= 64 || (p.BN <= 16 && p.Cin >= 32))) p.stk = 0;
  if (const char* e = getenv("R2N2_FWD3_STK")) {
    if (!atoi(e)) p.stk = 0;
    else p.stk = (p.kh == 3 && p.kw == 3 && p.nchunk == 1 && 3 * p.BN <= 256 && p.BN * p.rb == p.w_sub_bytes && !p.pair) ? 1 : 0;
  }
  if (want_stk == 0) p.stk = 0;
  p.xch_off = 0; p.xch_bytes = 0;
  if (p.stk) { budget -= 10 * 1024; want_tma_epi = 0; }     // (exchange buffer of the stacked epilogue)
  // warp-staged epilogue (8 warps x 2 KB): whenever a pixel's channels are whole 64-byte rounds and no other epilogue is chosen
  p.stage_epi = 0; p.stg_off = 0;
  {
    const int out_bytes = p.out_bytes;
    // (measured: it pays at 192 bytes per pixel -- conv1b forward 0.51 -> 0.44 ms, conv1a 0.30 -> 0.28 ms; the 64 / 128-byte
    //  layers of the decoder run 2-5 % faster on the thread-per-pixel epilogue, whose column split balances the two warps)
    // (with 256-bit thread-per-pixel accesses the other widths are as fast or faster without staging: conv2a 0.171 vs 0.179 ms)
    int want_stage = (!p.stk && want_tma_epi != 2 && p.BN * out_bytes == 192) ? 1 : 0;
    if (const char* e = getenv("R2N2_FWD3_STAGE_EPI")) want_stage = (atoi(e) != 0 && !p.stk && (p.BN * out_bytes) % 64 == 0) ? 1 : 0;
    if (want_stage) { p.stage_epi = 1; want_tma_epi = 0; budget -= 16 * 1024; }
    p.ew = (p.stage_epi && out_bytes == 2 && !(getenv("R2N2_FWD3_EW") && atoi(getenv("R2N2_FWD3_EW")) == 8)) ? 16 : 8;
  }
  p.bn_cols = ru(p.stk ? 3 * p.BN : p.BN, 32);
  const int Tmax = 256 / p.bn_cols;                       // two accumulator sets in 512 TMEM columns
  if (Tmax < 1) return false;
  // weight ring: one filter tap (all channel chunks) per stage; depth fixed after the block is chosen
  // (9, 3 or 1 taps: each barrier round costs the issuing thread ~300 cycles, so small taps are grouped)
  const int tpp = p.kh * p.kw, ntaps = p.kd * tpp;
  const int all_w = ntaps * p.nchunk * p.w_sub_bytes;
  // (pair mode: half the filter per CTA -- conv1b's 9 x 96 x 96 taps are 81 KB per CTA and stay resident, which removes
  //  the weight ring whose refill round trip bounded that layer: tools/r02 R2N2_FWD3_SKIP experiments)
  p.w_res = (all_w <= (p.pair ? 96 : 64) * 1024 && !getenv("R2N2_FWD3_NO_RES")) ? 1 : 0;
  p.wt = tpp;                                            // a stage never straddles two planes
  if (!p.w_res) {
    while (p.wt > 1 && p.wt * p.nchunk * p.w_sub_bytes > 24 * 1024) p.wt = (p.wt % 3 == 0) ? p.wt / 3 : 1;
    if (const char* e = getenv("R2N2_FWD3_WT")) { int v = atoi(e); if (v >= 1 && tpp % v == 0) p.wt = v; }
  }
  if (p.stk && !p.w_res && p.wt % 3 != 0) p.wt = 3;       // a weight stage holds whole filter rows
  p.wg = p.wt * p.nchunk;
  const int wstage_bytes = p.wg * p.w_sub_bytes;
  p.wstages = p.w_res ? 1 : 2;
  const int w_bytes = p.w_res ? all_w : p.wstages * wstage_bytes;
  p.nring = p.kd == 1 ? 2 : 4;
  double best = -1.0;
  int bX = 0, bY = 0;
  for (int Xb = (p.W < 8 ? p.W : 8); Xb + p.kw - 1 <= 256 && Xb <= p.W; ++Xb) {
    const int bx = (p.W + Xb - 1) / Xb;
    const int P = Xb + p.kw - 1;
    for (int Yb = 1; Yb <= p.H && Yb <= 64; ++Yb) {
      const int R = Yb * P;
      const int T = (R + 127) / 128;
      if (T > Tmax) break;
      const int reach = T * 128 + (p.kh - 1) * P + p.kw - 1;
      const int rows = (Yb + p.kh - 1) * P;
      if (Yb + p.kh - 1 > 256) break;
      const int chunk_bytes = ru((reach > rows ? reach : rows) * p.rb, 1024);
      if ((long long)p.nring * p.nchunk * chunk_bytes + w_bytes + 256 > budget) break;
      const int by = (p.H + Yb - 1) / Yb;
      // useful MMA rows / issued rows, and (weakly) the halo overhead of the loads
      const double eff = ((double)p.W / (bx * Xb)) * ((double)Xb * Yb / (T * 128)) * ((double)p.H / (by * Yb)) *
                         (0.7 + 0.3 * ((double)Xb * Yb / rows));
      if (eff > best + 1e-9) { best = eff; bX = Xb; bY = Yb; }
    }
  }
  if (bX == 0 || best < 0.6) return false;
  if (const char* e = getenv("R2N2_FWD3_BLOCK")) { int a, b; if (sscanf(e, "%d,%d", &a, &b) == 2) { bX = a; bY = b; } }
  p.Xb = bX; p.Yb = bY; p.P = bX + p.kw - 1;
  p.bx = (p.W + bX - 1) / bX; p.by = (p.H + bY - 1) / bY;
  p.R = bY * p.P;
  p.P_magic = (unsigned)((0x100000000ULL + (unsigned long long)p.P - 1) / (unsigned long long)p.P);
  p.T = (p.R + 127) / 128;
  p.plane_rows = (bY + p.kh - 1) * p.P;
  const int reach = p.T * 128 + (p.kh - 1) * p.P + p.kw - 1;
  p.chunk_bytes = ru((reach > p.plane_rows ? reach : p.plane_rows) * p.rb, 1024);
  p.plane_bytes = p.nchunk * p.chunk_bytes;
  // TMA epilogue: worth its shared memory where thread-per-pixel global accesses (~66 L1 wavefront cycles per
  // 16-byte instruction: 2 stores + 2 loads per input tensor per 16-column chunk per warp and tile) would take
  // more than half of the step's MMA time
  if (want_tma_epi) {
    const int nin = ((p.flags & R2N2_EPI_RESIDUAL) ? 1 : 0) + ((p.flags & R2N2_EPI_MASK) ? 1 : 0);
    const double epi_cycles = (double)p.T * 4.0 * (p.BN / 16) * (2 + 2 * nin) * 66.0;
    const double mma_cycles = (double)p.T * ntaps * (p.Cin / 16) * (p.BN >= 128 ? p.BN / 2 : 32 + p.BN / 4);
    if (want_tma_epi == 2 || epi_cycles > 0.5 * mma_cycles) {
      p.SC = p.BN % 32 == 0 ? 32 : 16;
      p.slab_bytes = ru(p.Xb * p.Yb * p.SC * 2, 1024);
      int need = ((p.flags & R2N2_EPI_MASK) ? 2 : 1) * p.slab_bytes;
      // the slabs come out of the ring / weight-stage budget; skip if the pipeline would starve
      const int w_min = p.w_res ? all_w : 2 * wstage_bytes;
      // (streamed filters: a second slab would cost a weight stage, measured slower on conv1b: 0.52 -> 0.55 ms)
      if (nin == 0 && p.w_res && p.nring * p.plane_bytes + w_min + 2 * p.slab_bytes + 1024 <= budget &&
          !getenv("R2N2_FWD3_ONE_SLAB")) {
        p.nslab = 2;
        need = 2 * p.slab_bytes;
      }
      if (p.nring * p.plane_bytes + w_min + need + 1024 <= budget) {
        p.tma_epi = 1;
        budget -= need;
      }
    }
  }
  // deeper input ring when shared memory allows: a plane load is ~3-4K cycles of latency, a small step
  // only ~1.5K cycles of MMA, so one plane of prefetch is not enough to keep the tensor pipe fed
  {
    const int w_need = p.w_res ? all_w : 3 * wstage_bytes;
    int extra = (budget - 256 - w_need) / p.plane_bytes - p.nring;
    const int max_extra = p.kd == 1 ? 6 : 2;
    if (extra > max_extra) extra = max_extra;
    if (extra > 0) p.nring += extra;
    if (const char* e = getenv("R2N2_FWD3_RING")) { int v = atoi(e); if (v >= p.kd + 1 && v <= 12) p.nring = v; }
  }
  p.w_off = p.nring * p.plane_bytes;
  // as many weight stages as fit (the taps of a step stream through them back to back)
  if (!p.w_res) {
    p.wstages = (budget - 256 - p.w_off) / wstage_bytes;
    if (p.wstages > 9) p.wstages = 9;
    if (p.wstages < 2) return false;
  }
  p.w_bytes = p.w_res ? all_w : p.wstages * wstage_bytes;
  p.epi_off = ru(p.w_off + p.w_bytes + (2 * p.nring + 2 * p.wstages + 5) * 8 + 16, 1024);
  p.stg_off = p.epi_off;
  if (p.stk) {
    p.xch_off = ru(p.w_off + p.w_bytes + (2 * p.nring + 2 * p.wstages + 5) * 8 + 16, 16);   // float4 quadruples
    p.xch_bytes = 2 * p.T * 4 * p.BN * 16 + p.BN * 16;     // + the all-zero block
  }
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.T * p.bn_cols) p.tmem_cols *= 2;
  if (p.kd == 1) { p.zc = 1; }
  else {
    // planes per item: amortise the 2 halo planes, keep enough items to balance the SMs
    p.zc = p.D < 8 ? p.D : 8;
    if (const char* e = getenv("R2N2_FWD3_ZC")) { int v = atoi(e); if (v >= 1 && v <= p.D) p.zc = v; }
  }
  p.zitems = (p.D + p.zc - 1) / p.zc;
  p.items = (long long)p.N * p.zitems * p.by * p.bx;
  return p.T >= 1 && p.T <= Tmax;
}

bool conv_fwd3_applicable(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw) {
  if (getenv("R2N2_UMMA_NO_FWD3")) return false;
  const bool k33 = kh == 3 && kw == 3 && (kd == 1 || kd == 3);
  const bool k71 = kd == 1 && kh == 7 && kw == 1;           // row-packed first layer
  if (!k33 && !k71) return false;
  if (kd == 1 && D != 1) return false;
  if (Cout % 16 || Cout > 128) return false;
  if (Cin != 16 && Cin != 32 && Cin != 64 && Cin != 96 && Cin != 128) return false;   // instantiated chunk patterns
  if (!getenv("R2N2_FWD3_FORCE")) {
    if ((long long)N * D * H * W < 128 * 148) return false;   // small layers: the per-tap kernel is fine
    // Measured (tools/bench_conv.py, B200): the win needs the streamed weights of a step to be small next to
    // its MMA time.  3-D layers with Cin <= 32 (conv10b/c, conv11 and their data gradients) run 1.6-2.3x
    // faster, conv1b (96 -> 96, 127^2) 1.25x; at Cin >= 64 only two weight stages fit beside the plane ring
    // and the weight TMA latency is exposed (conv9b, conv10a, conv2a/b: 0.9-1.0x) -> per-tap kernel.
    // (Cin = 64: only in dx-stacked mode, i.e. Cout <= 64 -- see plan_fwd3)
    const bool small3d = kd == 3 && (Cin <= 32 || (Cin == 64 && (3 * Cout <= 256 || getenv("R2N2_FWD3_C64")) &&
                                                   !getenv("R2N2_FWD3_NO_C64")));
    const bool conv1b_like = kd == 1 && Cin == 96 && Cout == 96;
    // round 2, with the warp-staged epilogue (>= 192 output bytes per pixel): conv2a (96 -> 128) 0.220 -> 0.192 ms and the
    // row-packed 7x1 first layer 0.333 -> 0.279 ms (resident filter, input loaded once instead of once per tap);
    // conv2b / the 128 -> 96 data gradient stay on the per-tap kernel (0.216 vs 0.316 ms, 0.171 vs 0.222 ms: their
    // streamed filters leave too few weight stages).  R2N2_FWD3_K71=0 / R2N2_FWD3_C96=0 restore the round-1 routing.
    const bool conv2a_like = kd == 1 && Cin == 96 && Cout == 128 && !(getenv("R2N2_FWD3_C96") && !atoi(getenv("R2N2_FWD3_C96")));
    const bool k71_on = k71 && Cin <= 32 && Cout * 2 >= 192 && !(getenv("R2N2_FWD3_K71") && !atoi(getenv("R2N2_FWD3_K71")));
    if (!small3d && !conv1b_like && !conv2a_like && !k71_on) return false;
  }
  Fwd3Params p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = kh; p.kw = kw;
  p.pair = 0; p.skip = 0; p.flags = 0; p.out_bytes = 2;
  return plan_fwd3(p);
}

struct Fwd3Launch {
  const CUtensorMap& map_a;
  const CUtensorMap& map_b;
  const CUtensorMap& map_y;
  const CUtensorMap& map_res;
  const CUtensorMap& map_mask;
  const float* bias;
  const void* mask;
  const void* res;
  void* y;
  const Fwd3Params& p;
  unsigned grid;
  size_t smem;
  cudaStream_t st;
};

template <typename TOut, int KB, int NCH, int LASTK, int TT = 0, int WTT = 0, int RESS = -1, int KDD = 0, int STK = 0, int PAIR = 0, int EW = 8>
static int launch_fwd3(const Fwd3Launch& L) {
  auto kern = conv_fwd3_umma_kernel<TOut, KB, NCH, LASTK, TT, WTT, RESS, KDD, STK, PAIR, EW>;
  constexpr int kThreads = 96 + 32 * EW;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "conv_fwd[umma3]: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  // (PAIR: clusters of two CTAs, one TPC -- tcgen05 cta_group::2)
  cudaError_t le = r2n2::launch_k((kern), dim3(L.grid), dim3(kThreads), L.smem, (cudaStream_t)(L.st), PAIR ? 2 : 0,
      L.map_a, L.map_b, L.map_y, L.map_res, L.map_mask, L.bias, (const TOut*)L.mask, (const TOut*)L.res, (TOut*)L.y, L.p);
  R2N2_CHECK_ARG(le == cudaSuccess, R2N2_ERR_CUDA, "conv_fwd[umma3]: launch: %s", cudaGetErrorString(le));
  return R2N2_OK;
}
template <int KB, int NCH, int LASTK, int TT = 0, int WTT = 0, int RESS = -1, int KDD = 0, int STK = 0, int PAIR = 0>
static int launch_fwd3_dt(const Fwd3Launch& L, int out_dtype) {
  return out_dtype == R2N2_BF16 ? launch_fwd3<__nv_bfloat16, KB, NCH, LASTK, TT, WTT, RESS, KDD, STK, PAIR>(L)
                                : launch_fwd3<float, KB, NCH, LASTK, TT, WTT, RESS, KDD, STK, PAIR>(L);
}
// bf16 outputs through the warp-staged epilogue with 16 epilogue warps
template <int KB, int NCH, int LASTK, int TT, int WTT, int RESS, int KDD, int PAIR>
static int launch_fwd3_ew16(const Fwd3Launch& L) {
  return launch_fwd3<__nv_bfloat16, KB, NCH, LASTK, TT, WTT, RESS, KDD, 0, PAIR, 16>(L);
}

int conv_fwd3_umma(const void* x, const void* w, const float* bias, const void* mask, const void* res, void* y,
                   int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int flags, int out_dtype,
                   cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  R2N2_CHECK_ARG(encode != nullptr, R2N2_ERR_CUDA, "conv_fwd[umma3]: cuTensorMapEncodeTiled unavailable");
  Fwd3Params p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kd = kd; p.kh = kh; p.kw = kw;
  p.flags = flags & (15 | R2N2_EPI_RES_PRE);
  p.out_bytes = out_dtype == R2N2_F32 ? 4 : 2;
  R2N2_CHECK_ARG(!(flags & R2N2_EPI_RES_PRE) || (out_dtype == R2N2_F32 && (flags & R2N2_EPI_RESIDUAL)),
                 R2N2_ERR_UNSUPPORTED, "conv_fwd[umma3]: RES_PRE needs RESIDUAL and an fp32 output");
  p.no_rot = getenv("R2N2_FWD3_NO_ROT") ? 1 : 0;
  // CTA-pair mode (cta_group::2): where half of the filter per CTA fits next to the plane ring and N >= 96 makes the
  // B-operand read worth halving.  R2N2_FWD3_PAIR=0/1 forces it off / on (any 2-D or 3-D layer with Cout % 32 == 0).
  // (also conv10a's data gradient, 32 -> 64 channels at 32^3: 55 KB of half filters resident, 0.168 -> 0.147 ms)
  p.pair = (((kd == 1 && Cin == 96 && Cout == 96) || (kd == 3 && Cin == 32 && Cout == 64)) && (num_sms() % 2) == 0) ? 1 : 0;
  if (const char* e = getenv("R2N2_FWD3_PAIR")) p.pair = (atoi(e) != 0 && Cout % 32 == 0 && Cout >= 32 && (num_sms() % 2) == 0) ? 1 : 0;
  p.skip = getenv("R2N2_FWD3_SKIP") ? atoi(getenv("R2N2_FWD3_SKIP")) : 0;
  // Measured: on this kernel the smem-slab + TMA epilogue pays only without input slabs and where the register
  // epilogue binds: conv1b forward (96 -> 96, bias + LeakyReLU) 0.553 -> 0.520 ms; with a mask / residual slab
  // every slab's load -> compute -> store latency is exposed (one team): masked data gradient 0.66 -> 0.74 ms.
  // R2N2_FWD3_TMA_EPI=0/1 forces it off / on (both parity-tested).
  const bool has_in = (flags & (R2N2_EPI_RESIDUAL | R2N2_EPI_MASK)) != 0;
  // (round 2: the warp-staged epilogue -- plan_fwd3 -- replaces it by default: no slab hand-over between CTA-wide barriers
  //  and no wait for the TMA store's shared-memory read; the TMA path stays selectable and tested)
  int want = 0;
  (void)has_in;
  if (const char* e = getenv("R2N2_FWD3_TMA_EPI")) want = out_dtype == R2N2_BF16 ? (atoi(e) ? 2 : 0) : 0;
  R2N2_CHECK_ARG(plan_fwd3(p, want), R2N2_ERR_UNSUPPORTED, "conv_fwd[umma3]: layer does not fit the flat-halo kernel");
  p.err_word = get_err_word();
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kbox, (cuuint32_t)p.P, (cuuint32_t)(p.Yb + kh - 1), 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.rb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_fwd[umma3]: cuTensorMapEncodeTiled(A) failed (%d)", (int)r);
  }
  {
    const long long K = (long long)kd * kh * kw * Cin;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kbox, (cuuint32_t)(p.pair ? p.BN / 2 : p.BN)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.rb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_fwd[umma3]: cuTensorMapEncodeTiled(B) failed (%d)", (int)r);
  }
  const size_t used_bytes = p.tma_epi
      ? (size_t)p.epi_off + (size_t)((p.flags & R2N2_EPI_MASK) || p.nslab == 2 ? 2 : 1) * p.slab_bytes
      : (p.stk ? (size_t)p.xch_off + (size_t)p.xch_bytes
               : (p.stage_epi ? (size_t)p.stg_off + 16 * 1024
                              : (size_t)p.w_off + (size_t)p.w_bytes + (2 * p.nring + 2 * p.wstages + 5) * 8 + 16));
  p.bias_off = (int)((used_bytes + 15) / 16 * 16);
  const size_t smem_bytes = (size_t)p.bias_off + 512 + 1024;
  CUtensorMap map_y = map_a, map_res = map_a, map_mask = map_a;
  if (p.tma_epi) {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2,
                             (cuuint64_t)D * H * W * Cout * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.SC, (cuuint32_t)p.Xb, (cuuint32_t)p.Yb, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const void* ptrs[3] = {y, (flags & R2N2_EPI_RESIDUAL) ? res : y, (flags & R2N2_EPI_MASK) ? mask : y};
    CUtensorMap* maps[3] = {&map_y, &map_res, &map_mask};
    for (int i = 0; i < 3; ++i) {
      R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, R2N2_ERR_BAD_ALIGN,
                     "conv_fwd[umma3]: y / residual / mask must be 16-byte aligned");
      CUresult r = encode(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptrs[i]), dims, strides, box,
                          es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.SC * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_fwd[umma3]: cuTensorMapEncodeTiled(epilogue %d) failed (%d)",
                     i, (int)r);
    }
  }
  R2N2_CHECK_ARG(smem_bytes <= 227 * 1024, R2N2_ERR_UNSUPPORTED, "conv_fwd[umma3]: smem %zu too large", smem_bytes);
  long long grid = num_sms();
  if (p.pair) {
    const long long pitems = (p.items + 1) / 2;
    if (grid > 2 * pitems) grid = 2 * pitems;
    grid &= ~1LL;
  } else if (grid > p.items) grid = p.items;
  if (getenv("R2N2_UMMA_DBG"))
    fprintf(stderr, "[umma fwd3] %dx%dx%dx%d Cin %d Cout %d kd %d block %dx%d P %d R %d T %d bands %dx%d zc %d items %lld "
            "kbox %d nchunk %d plane %d ring %d wsub %d wt %d wstages %d resident %d tma_epi %d stk %d pair %d tmem %d smem %zu\n", N, D, H, W, Cin, Cout, kd,
            p.Xb, p.Yb, p.P, p.R, p.T, p.bx, p.by, p.zc, p.items, p.kbox, p.nchunk, p.plane_bytes, p.nring,
            p.w_sub_bytes, p.wt, p.wstages, p.w_res, p.tma_epi, p.stk, p.pair, p.tmem_cols, smem_bytes);
  R2N2_CHECK_ARG(out_dtype == R2N2_BF16 || out_dtype == R2N2_F32, R2N2_ERR_BAD_DTYPE,
                 "conv_fwd[umma3]: bad out dtype %d", out_dtype);
  const Fwd3Launch L{map_a, map_b, map_y, map_res, map_mask, bias, mask, res, y, p, (unsigned)grid, smem_bytes, st};
  int rc;
  const bool k33 = kh == 3 && kw == 3 && !getenv("R2N2_FWD3_GENERIC");
  const bool k71s = kd == 1 && kh == 7 && kw == 1 && !getenv("R2N2_FWD3_GENERIC");
  // specialised (compile-time issue sequence) configurations of the training step
  if (p.pair) {
    // CTA-pair configurations (cta_group::2)
    if (k33 && Cin == 96 && kd == 1 && p.T == 2 && p.wt == 9 && p.w_res && p.ew == 16) rc = launch_fwd3_ew16<2, 3, 2, 2, 9, 1, 1, 1>(L);
    else if (k33 && Cin == 96 && kd == 1 && p.T == 2 && p.wt == 9 && p.w_res) rc = launch_fwd3_dt<2, 3, 2, 2, 9, 1, 1, 0, 1>(L, out_dtype);
    else if (Cin == 96) rc = launch_fwd3_dt<2, 3, 2, 0, 0, -1, 0, 0, 1>(L, out_dtype);
    else if (Cin == 128) rc = launch_fwd3_dt<4, 2, 4, 0, 0, -1, 0, 0, 1>(L, out_dtype);
    else if (Cin == 64) rc = launch_fwd3_dt<4, 1, 4, 0, 0, -1, 0, 0, 1>(L, out_dtype);
    else if (Cin == 32) rc = launch_fwd3_dt<2, 1, 2, 0, 0, -1, 0, 0, 1>(L, out_dtype);
    else rc = launch_fwd3_dt<1, 1, 1, 0, 0, -1, 0, 0, 1>(L, out_dtype);
  }
  else if (p.stk) {
    // dx-stacked configurations (generic issue loop unless listed)
    if (k33 && Cin == 32 && kd == 3 && p.T == 2 && p.w_res) rc = launch_fwd3_dt<2, 1, 2, 2, 9, 1, 3, 1>(L, out_dtype);
    else if (k33 && Cin == 32 && kd == 3 && p.T == 4 && p.w_res) rc = launch_fwd3_dt<2, 1, 2, 4, 9, 1, 3, 1>(L, out_dtype);
    else if (k33 && Cin == 16 && kd == 3 && p.T == 2 && p.w_res) rc = launch_fwd3_dt<1, 1, 1, 2, 9, 1, 3, 1>(L, out_dtype);
    else if (k33 && Cin == 32 && kd == 3 && p.T == 1 && p.wt == 3 && !p.w_res) rc = launch_fwd3_dt<2, 1, 2, 1, 3, 0, 3, 1>(L, out_dtype);
    else if (k33 && Cin == 64 && kd == 3 && p.T == 2 && p.wt == 3 && !p.w_res) rc = launch_fwd3_dt<4, 1, 4, 2, 3, 0, 3, 1>(L, out_dtype);
    else if (k33 && Cin == 64 && kd == 3 && p.T == 1 && p.wt == 3 && !p.w_res) rc = launch_fwd3_dt<4, 1, 4, 1, 3, 0, 3, 1>(L, out_dtype);
    else if (Cin == 16) rc = launch_fwd3_dt<1, 1, 1, 0, 0, -1, 0, 1>(L, out_dtype);
    else if (Cin == 32) rc = launch_fwd3_dt<2, 1, 2, 0, 0, -1, 0, 1>(L, out_dtype);
    else rc = launch_fwd3_dt<4, 1, 4, 0, 0, -1, 0, 1>(L, out_dtype);
  }
  else if (k71s && Cin == 32 && p.T == 2 && p.wt == 7 && p.w_res && p.ew == 16) rc = launch_fwd3_ew16<2, 1, 2, 2, 7, 1, 71, 0>(L);
  else if (k71s && Cin == 32 && p.T == 2 && p.wt == 7 && p.w_res) rc = launch_fwd3_dt<2, 1, 2, 2, 7, 1, 71>(L, out_dtype);
  else if (k33 && Cin == 96 && kd == 1 && p.T == 2 && p.wt == 1 && !p.w_res) rc = launch_fwd3_dt<2, 3, 2, 2, 1, 0, 1>(L, out_dtype);
  else if (k33 && Cin == 32 && kd == 3 && p.T == 3 && p.wt == 9 && p.w_res) rc = launch_fwd3_dt<2, 1, 2, 3, 9, 1, 3>(L, out_dtype);
  else if (k33 && Cin == 16 && kd == 3 && p.T == 3 && p.wt == 9 && p.w_res) rc = launch_fwd3_dt<1, 1, 1, 3, 9, 1, 3>(L, out_dtype);
  else if (k33 && Cin == 64 && kd == 3 && p.T == 2 && p.wt == 3 && !p.w_res) rc = launch_fwd3_dt<4, 1, 4, 2, 3, 0, 3>(L, out_dtype);
  else if (k33 && Cin == 32 && kd == 3 && p.T == 2 && p.wt == 3 && !p.w_res) rc = launch_fwd3_dt<2, 1, 2, 2, 3, 0, 3>(L, out_dtype);
  else if (k33 && Cin == 32 && kd == 3 && p.T == 3 && p.wt == 3 && !p.w_res) rc = launch_fwd3_dt<2, 1, 2, 3, 3, 0, 3>(L, out_dtype);
  else if (Cin == 16) rc = launch_fwd3_dt<1, 1, 1>(L, out_dtype);
  else if (Cin == 32) rc = launch_fwd3_dt<2, 1, 2>(L, out_dtype);
  else if (Cin == 64) rc = launch_fwd3_dt<4, 1, 4>(L, out_dtype);
  else if (Cin == 96) rc = launch_fwd3_dt<2, 3, 2>(L, out_dtype);
  else rc = launch_fwd3_dt<4, 2, 4>(L, out_dtype);
  if (rc != R2N2_OK) return rc;
  R2N2_CHECK_LAUNCH("conv_fwd[umma3]");
  return R2N2_OK;
}

}  // namespace r2n2
