// Persistent tcgen05 / TMEM / TMA implicit-GEMM convolution (forward and data gradient), sm_100a.
//
//   y[p, co] = epilogue( sum_{tap, ci} x[p + off(tap), ci] * w[co, tap, ci] )     bf16 x bf16 -> fp32
//
// One CTA per SM loops over (M tile, N tile) pairs:
//   * M tile = 128-pixel SPATIAL BOX of the channels-last tensor; for each filter tap the A operand
//     is one tiled-TMA load of the shifted box (out-of-bounds zero fill = 'same' padding);
//   * N tile = BN output channels (<= 256); weights [Cout][taps*Cin] come in through a 2-D TMA;
//   * a pipeline STAGE holds NSUB (tap, channel-chunk) sub-blocks, so that one full/empty barrier
//     round feeds NSUB * kbox/16 MMAs (the per-round wait/fence/commit overhead of the single
//     issuing thread is ~300 cycles: it has to be amortised over >= 512 cycles of tensor work);
//   * TWO accumulators live in TMEM (2 x BN columns): the epilogue warps drain tile i
//     (tcgen05.ld -> bias / LeakyReLU / residual / LeakyReLU' -> 16-byte global stores) while the
//     MMA warp already accumulates tile i+1 and the TMA warp prefetches further ahead.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue (two warps per TMEM lane quadrant).  All role loops are executed by
// the whole warp (elect.sync around the issuing instructions) so operands stay in uniform registers.
#include "umma.cuh"

namespace r2n2 {

struct FwdParams {
  int N, D, H, W, Cin, Cout;
  int kd, kh, kw;
  int bw, bh, bd, bn;                      // spatial box of the M tile
  int tiles_w, tiles_h, tiles_d, tiles_n;  // boxes per axis
  int m_tiles, n_tiles, total_tiles;
  int BN, bn_cols;                         // N tile and its TMEM column footprint
  int kbox, kchunks;                       // channels per sub-block, sub-blocks per tap
  int nsub;                                // sub-blocks per pipeline stage
  int stages;
  int a_bytes, b_bytes, stage_bytes;       // per sub-block A / B bytes (1024-aligned), per stage
  int tmem_cols;
  int flags;
  int tma_epi;                             // epilogue through shared memory + TMA loads/stores (bf16 out, mode != UP2)
  int SC, slab_bytes, epi_off;             // slab = 128 rows x SC output channels; smem offset of the staging buffers
  float* colsum;                           // TMA epilogue, one N tile: colsum[c] += sum over pixels of y[p, c] (fp32), or NULL
  int csum_off;                            //   smem offset of the per-team partial sums (2 x BN floats)
  int mc;                                  // multicast cluster: the mc CTAs of a cluster compute mc N tiles of ONE M tile; each loads
  int mc_rows;                             //   mc_rows rows of every A box and multicasts them to all (slice r at smem row r * mc_rows)
  signed char mc_dn[8], mc_dd[8];          //   box-relative (image, plane) coordinates of the slice of cluster rank r
  int mode;                                // 0 plain, 1 UP2 (zero-stuffed x2 input, 8 parity classes), 2 DOWN2
  int cls_tiles;                           // tiles per parity class (= m_tiles * n_tiles)
  int* err_word;
  // conv-GRU gate math fused into the register epilogue (R2N2_GRU_*; r2n2_conv_fwd_gru): no y tensor, the accumulator row of
  // a hidden-grid position feeds the gate formulas directly (models/res_gru_net.py:130-151 and their gradients)
  int bias_off;                            // smem offset of the epilogue warps' bias copy (Cout floats), or -1: read global
  int gru;
  const void* gin[4];
  void* gout[3];
};

// Filter taps that hit non-zero input along one axis.  Plain / DOWN2: all k taps, input offset
// j - (k-1)/2.  UP2 (input = zero-stuffed x2 upsampling of the small tensor, output position 2i+par):
// only taps t with (par + t - centre) even touch a stored value, which is small[i + (par+t-centre)/2].
// Returns count; tap index of entry j is t0 + j*ts, input offset o0 + j.
static __device__ __forceinline__ int tap_spec(int mode, int k, int par, int& t0, int& ts, int& o0) {
  const int c = (k - 1) / 2;
  if (mode != 1) { t0 = 0; ts = 1; o0 = -c; return k; }
  t0 = (c + par) & 1; ts = 2; o0 = (par + t0 - c) / 2;
  return (k - t0 + 1) / 2;
}

struct TileCoord {
  int nt, tw, th, td, tn, cls;
};
static __device__ __forceinline__ TileCoord decode_tile(const FwdParams& p, int tile) {
  TileCoord c;
  c.cls = 0;
  if (p.mode == 1) {                       // heavy classes (8 taps) first: better tail balance
    const int q = tile / p.cls_tiles;
    tile -= q * p.cls_tiles;
    c.cls = 7 - q;
  }
  c.nt = tile % p.n_tiles;
  int t = tile / p.n_tiles;
  c.tw = t % p.tiles_w; t /= p.tiles_w;
  c.th = t % p.tiles_h; t /= p.tiles_h;
  c.td = t % p.tiles_d; t /= p.tiles_d;
  c.tn = t;
  return c;
}

// Gate epilogue of one accumulator row (hidden-grid position pr) over the 16-column chunks [c_begin, c_end) of a tile.
// HC = hidden channels.  Inputs are fetched BEFORE the TMEM load of the chunk; every lane executes the tcgen05.ld.
//   UR    (Cout = 2 HC): s = sigmoid(acc + xp[pr,co] + b[co]); co < HC: u = s; else r = s, rh = s * h
//   BLEND (Cout = HC):   c = tanh(acc + xp[pr,co] + b[co]); h' = u h + (1 - u) c
//   RBWD  (Cout = HC):   g = acc (= d(r h)); da_r = g h r (1 - r) -> da_ur[pr, HC + co]; dh_prev += g r
//   OBWD  (Cout = HC):   dh = acc + dh_in; da_c = dh (1 - u)(1 - c^2); da_u = dh (h - c) u (1 - u) -> da_ur[pr, co]; dh_prev = dh u
template <typename T>
struct GruIn {
  Row16<T> a0, a1, a2, a3;
};
// operand rows of one 16-column chunk (issued BEFORE the accumulator-ready wait for the first chunk of a tile)
template <typename T>
static __device__ __forceinline__ void gru_fetch(GruIn<T>& g, const FwdParams& p, int co, bool ok, long long pr) {
  const int mode = p.gru;
  const int HC = mode == R2N2_GRU_UR ? (p.Cout >> 1) : p.Cout;
  g.a0.zero(); g.a1.zero(); g.a2.zero(); g.a3.zero();
  if (!ok || co >= p.Cout) return;
  const T* in0 = static_cast<const T*>(p.gin[0]);
  const T* in1 = static_cast<const T*>(p.gin[1]);
  const T* in2 = static_cast<const T*>(p.gin[2]);
  const T* in3 = static_cast<const T*>(p.gin[3]);
  const int ch = (mode == R2N2_GRU_UR && co >= HC) ? co - HC : co;      // hidden channel
  const long long hoff = pr * HC + ch;
  if (mode == R2N2_GRU_UR) {
    g.a0.load(in0 + pr * p.Cout + co);
    if (co >= HC) g.a1.load(in1 + hoff);
  } else if (mode == R2N2_GRU_BLEND) {
    g.a0.load(in0 + hoff); g.a1.load(in1 + hoff); g.a2.load(in2 + hoff);
  } else if (mode == R2N2_GRU_RBWD) {
    g.a0.load(in0 + hoff); g.a1.load(in1 + hoff); g.a2.load(static_cast<const T*>(p.gout[1]) + hoff);
  } else {
    g.a0.load(in0 + hoff); g.a1.load(in1 + hoff); g.a2.load(in2 + hoff);
    if (in3) g.a3.load(in3 + hoff);
  }
}
template <typename T>
static __device__ __forceinline__ void epi_row_gru(GruIn<T>& pf, const FwdParams& p, uint32_t taddr, int c_begin, int c_end,
                                                    int co0, bool ok, long long pr, const float* __restrict__ bias) {
  const int mode = p.gru;
  const int HC = mode == R2N2_GRU_UR ? (p.Cout >> 1) : p.Cout;
  T* out0 = static_cast<T*>(p.gout[0]);
  T* out1 = static_cast<T*>(p.gout[1]);
  T* out2 = static_cast<T*>(p.gout[2]);
  for (int c = c_begin; c < c_end; c += 16) {
    const int co = co0 + c;
    const bool act = ok && co < p.Cout;
    const int ch = (mode == R2N2_GRU_UR && co >= HC) ? co - HC : co;      // hidden channel
    const long long hoff = pr * HC + ch;
    const GruIn<T> cur = pf;
    if (c + 16 < c_end) gru_fetch(pf, p, co + 16, ok, pr);
    uint32_t v[16];
    tmem_ld16(taddr + (uint32_t)c, v);
    if (!act) continue;
    float f[16], q0[16], q1[16], q2[16], q3[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
    cur.a0.unpack(q0); cur.a1.unpack(q1); cur.a2.unpack(q2); cur.a3.unpack(q3);
    if (mode == R2N2_GRU_UR || mode == R2N2_GRU_BLEND) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + co + j);
        f[j] += q0[j] + b4.x; f[j + 1] += q0[j + 1] + b4.y; f[j + 2] += q0[j + 2] + b4.z; f[j + 3] += q0[j + 3] + b4.w;
      }
    }
    if (mode == R2N2_GRU_UR) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = sigmoidf_(f[j]);
      if (co < HC) {
        store16(out0 + hoff, f);
      } else {
        store16(out1 + hoff, f);
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] *= q1[j];
        store16(out2 + hoff, f);
      }
    } else if (mode == R2N2_GRU_BLEND) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = tanhf(f[j]);
      store16(out0 + hoff, f);
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = q1[j] * q2[j] + (1.f - q1[j]) * f[j];
      store16(out1 + hoff, f);
    } else if (mode == R2N2_GRU_RBWD) {
      float d[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { d[j] = f[j] * q0[j] * q1[j] * (1.f - q1[j]); f[j] = q2[j] + f[j] * q1[j]; }
      store16(out0 + pr * (2 * HC) + HC + ch, d);
      store16(out1 + hoff, f);
    } else {
      float d[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { f[j] += q0[j]; d[j] = f[j] * (1.f - q1[j]) * (1.f - q2[j] * q2[j]); }
      store16(out0 + hoff, d);
#pragma unroll
      for (int j = 0; j < 16; ++j) d[j] = f[j] * (q3[j] - q2[j]) * q1[j] * (1.f - q1[j]);
      store16(out1 + pr * (2 * HC) + ch, d);
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] *= q1[j];
      store16(out2 + hoff, f);
    }
  }
}

constexpr int kFwdThreads = 64 + 32 * kEpiWarps;

// NK16 / NSUB > 0: K16 steps per sub-block and sub-blocks per stage are compile-time (every stage full, every
// channel chunk full): the MMA issue sequence of a stage is straight-line code with immediate descriptor
// offsets -- the issuing thread, not the tensor pipe, sets the pace when it spends ~17 instructions per MMA.
template <typename TOut, int NK16, int NSUB, int MC, int GRU = 0>
__global__ void __launch_bounds__(kFwdThreads, (MC || GRU) ? 1 : 2)
conv_fwd_umma_persistent(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_res,
                         const __grid_constant__ CUtensorMap map_mask,
                         const float* __restrict__ bias, const TOut* mask, const TOut* res, TOut* y,
                         const FwdParams p) {
  r2n2::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  R2N2_WS_BEGIN(bars);
  uint64_t* full_bar = bars;                       // [stages]  TMA -> MMA
  uint64_t* empty_bar = bars + p.stages;           // [stages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * p.stages;       // [2]       MMA -> epilogue (accumulator ready)
  uint64_t* tempty_bar = bars + 2 * p.stages + 2;  // [2]       epilogue -> MMA (accumulator drained)
  uint64_t* efull_bar = bars + 2 * p.stages + 4;   // [2]       residual / mask slab landed (one per epilogue team)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 6);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      // multicast: a stage may be refilled (by every CTA's slice) only when ALL CTAs of the cluster have consumed it
      mbar_init(smem_u32(&empty_bar[s]), MC ? (uint32_t)p.mc : 1u);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull_bar[b]), 1);
      // one arrive per epilogue warp that drains the buffer (TMA epilogue: the 4 warps of one team)
      mbar_init(smem_u32(&tempty_bar[b]), p.tma_epi ? kEpiWarps / 2 : kEpiWarps);
      mbar_init(smem_u32(&efull_bar[b]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (p.colsum) {
    float* cs = reinterpret_cast<float*>(smem + p.csum_off);
    for (int i = threadIdx.x; i < 2 * p.BN; i += blockDim.x) cs[i] = 0.f;
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
  if (MC) cluster_sync_all();            // every CTA's barriers exist before the first multicast load / commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  r2n2::pdl_wait();            // prologue done; from here on global memory of the preceding kernels may be touched
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  const uint16_t cmask = MC ? (uint16_t)((1u << p.mc) - 1u) : (uint16_t)1;

  const int taps = p.kd * p.kh * p.kw;
  const int nq = taps * p.kchunks;                  // sub-blocks per tile
  const int row_bytes = p.kbox * 2;
  const int rows = p.bw * p.bh * p.bd * p.bn;
  const int sub_bytes = p.a_bytes + p.b_bytes;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);

  if (warp == 0) {
    // ================================ TMA producer =========================================
    const uint32_t sub_tx = (uint32_t)(rows * row_bytes + p.BN * row_bytes);
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      int t0x, tsx, o0x, t0y, tsy, o0y, t0z, tsz, o0z;
      const int cntx = tap_spec(p.mode, p.kw, tc.cls & 1, t0x, tsx, o0x);
      const int cnty = tap_spec(p.mode, p.kh, (tc.cls >> 1) & 1, t0y, tsy, o0y);
      const int cntz = tap_spec(p.mode, p.kd, (tc.cls >> 2) & 1, t0z, tsz, o0z);
      const int nq_t = cntx * cnty * cntz * p.kchunks;
      const int sc = (p.mode == 2) ? 2 : 1;              // DOWN2: output i reads input 2i + tap
      const int cx = sc * tc.tw * p.bw + o0x, cy = sc * tc.th * p.bh + o0y, cz = sc * tc.td * p.bd + o0z;
      const int n0 = tc.tn * p.bn;
      const int col0 = tc.nt * p.BN;
      int jz = 0, jy = 0, jx = 0, c0 = 0;                 // current sub-block (tap, channel chunk)
      int kcol = (((t0z * p.kh) + t0y) * p.kw + t0x) * p.Cin;
      for (int q0 = 0; q0 < nq_t; q0 += p.nsub) {
        const int ns = (nq_t - q0 < p.nsub) ? nq_t - q0 : p.nsub;
        mbar_wait(empty0 + 8u * s, ph ^ 1u, abort_flag, p.err_word);
        const uint32_t bar = full0 + 8u * s;
        const bool leader = elect_one();
        if (leader) mbar_expect_tx(bar, sub_tx * (uint32_t)ns);
        uint32_t dst = smem_base + (uint32_t)(s * p.stage_bytes);
        for (int u = 0; u < ns; ++u) {
          if (leader) {
            if (MC)       // this CTA's slice of the box, to every CTA of the cluster (each full barrier expects the whole box)
              tma_load_5d_mc(dst + (uint32_t)((int)crank * p.mc_rows * row_bytes), &map_a, bar, cmask, c0, cx + jx, cy + jy,
                             cz + jz + p.mc_dd[crank], n0 + p.mc_dn[crank]);
            else
              tma_load_5d(dst, &map_a, bar, c0, cx + jx, cy + jy, cz + jz, n0);
            tma_load_2d(dst + (uint32_t)p.a_bytes, &map_b, bar, kcol + c0, col0);
          }
          dst += (uint32_t)sub_bytes;
          c0 += p.kbox;
          if (c0 >= p.Cin) {
            c0 = 0;
            if (++jx == cntx) { jx = 0; if (++jy == cnty) { jy = 0; ++jz; } }
            kcol = ((((t0z + jz * tsz) * p.kh) + (t0y + jy * tsy)) * p.kw + (t0x + jx * tsx)) * p.Cin;
          }
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===========================================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t layout_type = (row_bytes == 128) ? 2u : (row_bytes == 64 ? 4u : 6u);
    const uint64_t desc_hi = make_kmajor_desc(0, 8u * (uint32_t)row_bytes, layout_type);
    const int last_nk16 = (p.Cin - (p.kchunks - 1) * p.kbox) >> 4;
    const int full_nk16 = p.kbox >> 4;
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(tempty0 + 8u * buf, (use & 1u) ^ 1u, abort_flag, p.err_word);   // accumulator drained
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.bn_cols);
      uint32_t acc = 0;
      int cc = 0;                                                          // channel chunk counter
      int nq_t = nq;
      if (p.mode == 1) {
        const TileCoord tc = decode_tile(p, tile);
        int a_, b_, c_;
        nq_t = tap_spec(1, p.kw, tc.cls & 1, a_, b_, c_) * tap_spec(1, p.kh, (tc.cls >> 1) & 1, a_, b_, c_) *
               tap_spec(1, p.kd, (tc.cls >> 2) & 1, a_, b_, c_) * p.kchunks;
      }
      for (int q0 = 0; q0 < nq_t; q0 += p.nsub) {
        const int ns = (nq_t - q0 < p.nsub) ? nq_t - q0 : p.nsub;
        mbar_wait(full0 + 8u * s, ph, abort_flag, p.err_word);
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)(s * p.stage_bytes);
        if (NK16 > 0) {
          if (elect_one()) {
            const uint64_t da0 = desc_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
            const uint64_t db0 = da0 + (uint64_t)((uint32_t)p.a_bytes >> 4);
            const uint32_t sub16 = (uint32_t)sub_bytes >> 4;
#pragma unroll
            for (int u = 0; u < (NSUB > 0 ? NSUB : 1); ++u) {
#pragma unroll
              for (int k = 0; k < (NK16 > 0 ? NK16 : 1); ++k)
                umma_bf16(d_tmem, da0 + (uint64_t)(u * sub16 + 2u * k), db0 + (uint64_t)(u * sub16 + 2u * k), idesc,
                          (u == 0 && k == 0) ? acc : 1u);
            }
            if (MC) umma_commit_mc(empty0 + 8u * s, cmask); else umma_commit(empty0 + 8u * s);
          }
        } else if (elect_one()) {
          uint32_t addr = a_addr;
          int c = cc;
          uint32_t a2 = acc;
          for (int u = 0; u < ns; ++u) {
            const int nk16 = (c == p.kchunks - 1) ? last_nk16 : full_nk16;
            const uint64_t da0 = desc_hi | (uint64_t)((addr & 0x3FFFF) >> 4);
            const uint64_t db0 = desc_hi | (uint64_t)(((addr + (uint32_t)p.a_bytes) & 0x3FFFF) >> 4);
            for (int k = 0; k < nk16; ++k) {
              umma_bf16(d_tmem, da0 + 2u * k, db0 + 2u * k, idesc, a2);
              a2 = 1;
            }
            addr += (uint32_t)sub_bytes;
            if (++c == p.kchunks) c = 0;
          }
          if (MC) umma_commit_mc(empty0 + 8u * s, cmask); else umma_commit(empty0 + 8u * s);   // stage free once these MMAs retire
        }
        __syncwarp();
        acc = 1;
        cc += ns;
        while (cc >= p.kchunks) cc -= p.kchunks;
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(tfull0 + 8u * buf);  // accumulator complete
      __syncwarp();
    }
  } else {
    // ================================ epilogue (warps 2..9) =================================
    const int quad = warp & 3;                          // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;
    // bias copy in shared memory: the L1 keeps nothing beside ~200 KB of shared memory, every 16-column chunk paid an
    // L2 round trip for its bias values (measured on the flat-halo kernel's staged epilogue: conv1a -6 %)
    const float* bsrc = bias;                           // (generic pointer: global, or the shared-memory copy)
    if (p.bias_off >= 0) {
      float* bs = reinterpret_cast<float*>(smem + p.bias_off);
      for (int i = (warp - 2) * 32 + lane; i < p.Cout; i += 32 * kEpiWarps) bs[i] = bias[i];
      named_bar_sync(3, 32 * kEpiWarps);
      bsrc = bs;
    }
    if (p.tma_epi) {
      // Two teams of 4 warps; team t drains accumulator buffer t (every other tile).  A thread-per-pixel
      // global access touches 32 different 128-byte lines per instruction (32 L1 wavefronts): on layers
      // with K <= ~1000 that, not the MMAs, bounded the tile time.  Here the tile goes through a swizzled
      // shared-memory slab (128 pixels x SC channels): residual / mask slabs arrive by TMA load, the result
      // leaves by ONE TMA tensor store (clipped at the tensor bounds), the threads only touch TMEM + smem.
      if constexpr (sizeof(TOut) == 2) {
        const int team = (warp - 2) >> 2;
        const bool leader = ((warp - 2) & 3) == 0 && lane == 0;     // always the same thread: it owns the bulk groups
        const uint32_t bufY = smem_base + (uint32_t)(p.epi_off + team * p.slab_bytes);
        const uint32_t bufM = smem_base + (uint32_t)(p.epi_off + (2 + team) * p.slab_bytes);
        const uint32_t efull = smem_u32(&efull_bar[team]);
        const uint32_t rb = (uint32_t)p.SC * 2u;
        const uint32_t smask = rb == 128 ? 7u : (rb == 64 ? 3u : 1u);
        const bool has_in = (p.flags & (R2N2_EPI_RESIDUAL | R2N2_EPI_MASK)) != 0;
        const uint32_t in_tx = (uint32_t)(rows * p.SC * 2) *
                               (((p.flags & R2N2_EPI_RESIDUAL) ? 1u : 0u) + ((p.flags & R2N2_EPI_MASK) ? 1u : 0u));
        uint32_t eph = 0;
        bool pending = false;
        int it = 0;
        // this thread's pixel inside the box (for the column sums: out-of-bounds rows must hold zeros)
        int rr = row;
        const int xi = rr % p.bw; rr /= p.bw;
        const int yi = rr % p.bh; rr /= p.bh;
        const int zi = rr % p.bd; rr /= p.bd;
        const int ni = rr;
        float* csum = reinterpret_cast<float*>(smem + p.csum_off) + team * p.BN;
        const int tt = ((warp - 2) & 3) * 32 + lane;          // thread index inside the team
        const int ncp = p.SC >> 1;                            // column pairs of a slab
        const int cp = tt % ncp, rg = tt / ncp, rows_per_rg = 128 / (128 / ncp);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
          if ((it & 1) != team) continue;
          const int buf = team;
          const uint32_t use = (uint32_t)(it >> 1);
          const TileCoord tc = decode_tile(p, tile);
          const int x0 = tc.tw * p.bw, y0 = tc.th * p.bh, z0 = tc.td * p.bd, n0 = tc.tn * p.bn;
          const bool inb = (row < rows) && x0 + xi < p.W && y0 + yi < p.H && z0 + zi < p.D && n0 + ni < p.N;
          const int col0 = tc.nt * p.BN;
          int ncol = p.Cout - col0;
          if (ncol > p.BN) ncol = p.BN;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.bn_cols);
          for (int sc = 0; sc < ncol; sc += p.SC) {
            if (leader && pending) tma_store_wait_read();          // previous store has read its slab
            named_bar_sync(1 + team, 128);
            if (has_in && leader) {
              mbar_expect_tx(efull, in_tx);
              if (p.flags & R2N2_EPI_RESIDUAL) tma_load_5d(bufY, &map_res, efull, col0 + sc, x0, y0, z0, n0);
              if (p.flags & R2N2_EPI_MASK) tma_load_5d(bufM, &map_mask, efull, col0 + sc, x0, y0, z0, n0);
            }
            if (sc == 0) {
              mbar_wait(tfull0 + 8u * buf, use & 1u, abort_flag, p.err_word);
              tc_fence_after();
            }
            if (has_in) {
              mbar_wait(efull, eph, abort_flag, p.err_word);
              eph ^= 1u;
            }
            for (int c = 0; c < p.SC; c += 16) {
              uint32_t v[16];
              tmem_ld16(taddr + (uint32_t)(sc + c), v);
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
              const int co = col0 + sc + c;
              if (p.flags & R2N2_EPI_BIAS) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bsrc + co + j);
                  f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                }
              }
              if (p.flags & R2N2_EPI_LRELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = lrelu(f[j]);
              }
              const uint32_t o0 = swz_off((uint32_t)row * rb + (uint32_t)c * 2u, smask);
              const uint32_t o1 = swz_off((uint32_t)row * rb + (uint32_t)c * 2u + 16u, smask);
              if (p.flags & R2N2_EPI_RESIDUAL) {
                Row16<__nv_bfloat16> r16;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r16.a.x), "=r"(r16.a.y), "=r"(r16.a.z), "=r"(r16.a.w) : "r"(bufY + o0));
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r16.b.x), "=r"(r16.b.y), "=r"(r16.b.z), "=r"(r16.b.w) : "r"(bufY + o1));
                float r[16];
                r16.unpack(r);
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += r[j];
              }
              if (p.flags & R2N2_EPI_MASK) {
                Row16<__nv_bfloat16> m16;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(m16.a.x), "=r"(m16.a.y), "=r"(m16.a.z), "=r"(m16.a.w) : "r"(bufM + o0));
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(m16.b.x), "=r"(m16.b.y), "=r"(m16.b.z), "=r"(m16.b.w) : "r"(bufM + o1));
                float m[16];
                m16.unpack(m);
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] *= lrelu_slope(m[j]);
              }
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                w[j] = (p.colsum && !inb) ? 0u : *reinterpret_cast<uint32_t*>(&h);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(bufY + o0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(bufY + o1), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            named_bar_sync(1 + team, 128);
            if (leader) {
              tma_store_5d(&map_y, bufY, col0 + sc, x0, y0, z0, n0);
              pending = true;
            }
            if (p.colsum) {
              // column sums of the finished slab (the bf16 values that were stored): thread (rg, cp) adds the
              // rows of its group for one column pair; per-CTA partial sums live in shared memory
              float s0 = 0.f, s1 = 0.f;
              const int r0 = rg * rows_per_rg;
              for (int r = r0; r < r0 + rows_per_rg; ++r) {
                const uint32_t o = swz_off((uint32_t)r * rb + ((uint32_t)(cp * 4) & ~15u), smask) + ((uint32_t)(cp * 4) & 15u);
                uint32_t wv;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(bufY + o));
                float2 fv = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&wv));
                s0 += fv.x; s1 += fv.y;
              }
              atomicAdd(csum + sc + 2 * cp, s0);
              atomicAdd(csum + sc + 2 * cp + 1, s1);
            }
          }
          // this team has finished reading the accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * buf) : "memory");
        }
        if (leader && pending) tma_store_wait_all();               // smem must outlive the last store
        if (p.colsum) {
          named_bar_sync(1 + team, 128);
          for (int c = tt; c < p.Cout; c += 128) atomicAdd(p.colsum + c, csum[c]);
        }
      }
    } else {
    int r = row;
    const int xi = r % p.bw; r /= p.bw;
    const int yi = r % p.bh; r /= p.bh;
    const int zi = r % p.bd; r /= p.bd;
    const int ni = r;
    // two warps per lane quadrant: each takes half of the tile's 16-column chunks
    const int nch = p.BN >> 4, half = (warp - 2) >> 2;
    const int c_begin = half ? ((nch + 1) >> 1) << 4 : 0, c_end = half ? p.BN : ((nch + 1) >> 1) << 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const TileCoord tc = decode_tile(p, tile);
      const int xx = tc.tw * p.bw + xi, yy = tc.th * p.bh + yi, zz = tc.td * p.bd + zi, nn = tc.tn * p.bn + ni;
      const bool valid = (row < rows) && xx < p.W && yy < p.H && zz < p.D && nn < p.N;
      long long pix;
      bool empty_acc = false;                              // UP2 class without any non-zero tap
      if (p.mode == 1) {
        const int px = tc.cls & 1, py = (tc.cls >> 1) & 1, pz = (tc.cls >> 2) & 1;
        pix = (((long long)nn * (2 * p.D) + 2 * zz + pz) * (2 * p.H) + 2 * yy + py) * (2 * p.W) + 2 * xx + px;
        int a_, b_, c_;
        empty_acc = tap_spec(1, p.kw, px, a_, b_, c_) * tap_spec(1, p.kh, py, a_, b_, c_) *
                        tap_spec(1, p.kd, pz, a_, b_, c_) == 0;
      } else {
        pix = (((long long)nn * p.D + zz) * p.H + yy) * p.W + xx;
      }
      const int col0 = tc.nt * p.BN;
      const long long off = pix * p.Cout;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.bn_cols);
      if constexpr (GRU) {
        GruIn<TOut> gpf;
        gru_fetch(gpf, p, col0 + c_begin, valid, pix);
        mbar_wait(tfull0 + 8u * buf, use & 1u, abort_flag, p.err_word);
        tc_fence_after();
        epi_row_gru<TOut>(gpf, p, taddr, c_begin, c_end, col0, valid, pix, bsrc);
      } else {
      EpiPrefetch<TOut> pf;
      epi_prefetch(pf, p.flags, valid, res + off, mask + off, col0 + c_begin, p.Cout);
      mbar_wait(tfull0 + 8u * buf, use & 1u, abort_flag, p.err_word);
      tc_fence_after();
      epi_row(pf, taddr, c_begin, c_end, col0, p.Cout, valid, empty_acc, p.flags, bsrc, res + off, mask + off, y + off);
      }
      // this warp has finished reading the accumulator: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * buf) : "memory");
    }
    }
  }
  R2N2_WS_END(p.err_word);
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();            // no CTA exits while a peer may still multicast into its shared memory / barriers
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

struct FwdLaunch {
  const CUtensorMap& map_a; const CUtensorMap& map_b; const CUtensorMap& map_y; const CUtensorMap& map_res;
  const CUtensorMap& map_mask;
  const float* bias; const void* mask; const void* res; void* y;
  const FwdParams& p;
  int grid; size_t smem; cudaStream_t st;
};
template <typename TOut, int NK16, int NSUB, int MC, int GRU = 0>
static int launch_fwd_k(const FwdLaunch& L) {
  auto kern = conv_fwd_umma_persistent<TOut, NK16, NSUB, MC, GRU>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    R2N2_CHECK_ARG(e == cudaSuccess, R2N2_ERR_CUDA, "conv_fwd[umma]: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaError_t le = r2n2::launch_k((kern), dim3(L.grid), dim3(kFwdThreads), L.smem, (cudaStream_t)(L.st), MC ? L.p.mc : 0,
      L.map_a, L.map_b, L.map_y, L.map_res, L.map_mask, L.bias, (const TOut*)L.mask, (const TOut*)L.res, (TOut*)L.y, L.p);
  R2N2_CHECK_ARG(le == cudaSuccess, R2N2_ERR_CUDA, "conv_fwd[umma]: launch: %s", cudaGetErrorString(le));
  return R2N2_OK;
}
template <typename TOut, int NK16, int NSUB>
static int launch_fwd(const FwdLaunch& L) {
  // fused gate epilogues: the generic and the (4, 3) issue sequence (the hidden convolutions at small batch) are instantiated
  if (L.p.gru) {
    if constexpr (sizeof(TOut) == 2) {
      if constexpr (NK16 == 4 && NSUB == 3) return launch_fwd_k<TOut, 4, 3, 0, 1>(L);
      else return launch_fwd_k<TOut, 0, 0, 0, 1>(L);
    } else {
      R2N2_CHECK_ARG(false, R2N2_ERR_BAD_DTYPE, "conv_fwd_gru[umma]: bf16 tensors only");
    }
  }
  // multicast clusters: only the generic and the (4, 2) / (4, 3) / (4, 4) issue sequences are instantiated
  if (L.p.mc > 1) {
    if constexpr (NK16 == 4 && (NSUB == 2 || NSUB == 3 || NSUB == 4)) return launch_fwd_k<TOut, NK16, NSUB, 1>(L);
    else return launch_fwd_k<TOut, 0, 0, 1>(L);
  }
  return launch_fwd_k<TOut, NK16, NSUB, 0>(L);
}

// resident clusters of `c` CTAs (one CTA per SM, full shared memory) the device can hold at once
static int max_active_clusters(int c) {
  static int cache[9] = {0};
  if (c < 2 || c > 8) return 0;
  if (cache[c]) return cache[c];
  auto kern = conv_fwd_umma_persistent<__nv_bfloat16, 0, 0, 1>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms() / c * c, 1, 1);
  cfg.blockDim = dim3(kFwdThreads, 1, 1);
  cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[c] = n > 0 ? n : -1;
  return cache[c];
}
template <int NK16, int NSUB>
static int launch_fwd_dt(const FwdLaunch& L, int out_dtype) {
  return out_dtype == R2N2_BF16 ? launch_fwd<__nv_bfloat16, NK16, NSUB>(L) : launch_fwd<float, NK16, NSUB>(L);
}

// choose the spatial box (bw,bh,bd,bn), product <= 128, minimising the number of M tiles
static void choose_box_fwd(int N, int D, int H, int W, int& bw, int& bh, int& bd, int& bn) {
  long long best = -1;
  for (int w = 1; w <= W && w <= 128; ++w)
    for (int h = 1; h <= H && w * h <= 128; ++h)
      for (int d = 1; d <= D && w * h * d <= 128; ++d) {
        int n = 128 / (w * h * d);
        if (n > N) n = N;
        if (n < 1) continue;
        long long tiles = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((D + d - 1) / d) *
                          ((N + n - 1) / n);
        // fewer tiles first; then the most compact box (short TMA rows hurt nothing, but a box that
        // follows the memory order keeps the L2 footprint of the shifted taps small)
        const bool better = best < 0 || tiles < best ||
                            (tiles == best && (w > bw || (w == bw && (h > bh || (h == bh && d > bd)))));
        if (better) {
          best = tiles;
          bw = w; bh = h; bd = d; bn = n;
        }
      }
}

int conv_fwd_umma_v1(const void* x, const void* w, const float* bias, const void* mask,
                     const void* res, void* y, int N, int D, int H, int W, int Cin, int Cout, int kd,
                     int kh, int kw, int flags, int in_dtype, int out_dtype, cudaStream_t st);

bool conv_fwd3_applicable(int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw);
int conv_fwd3_umma(const void* x, const void* w, const float* bias, const void* mask, const void* res, void* y,
                   int N, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int flags, int out_dtype,
                   cudaStream_t st);

// colsum (may be NULL): colsum[c] += sum over output pixels of y[p, c] -- the bias gradient of the layer
// whose output gradient this call produces; fused into the TMA epilogue where possible, else a second pass
int conv_fwd_umma(const void* x, const void* w, const float* bias, const void* mask,
                  const void* res, void* y, int N, int D, int H, int W, int Cin, int Cout, int kd,
                  int kh, int kw, int flags, int in_dtype, int out_dtype, float* colsum, cudaStream_t st,
                  const GruEpi* gru) {
  const int mode = (flags & R2N2_CONV_UP2) ? 1 : ((flags & R2N2_CONV_DOWN2) ? 2 : 0);
  R2N2_CHECK_ARG(!((flags & R2N2_CONV_UP2) && (flags & R2N2_CONV_DOWN2)), R2N2_ERR_BAD_SHAPE,
                 "conv_fwd[umma]: UP2 and DOWN2 are exclusive");
  const long long out_rows = (long long)N * D * H * W * (mode == 1 ? 8 : 1);
  if (getenv("R2N2_UMMA_V1") && mode == 0 && !colsum)   // A/B comparison with the one-tile-per-CTA kernel
    return conv_fwd_umma_v1(x, w, bias, mask, res, y, N, D, H, W, Cin, Cout, kd, kh, kw, flags, in_dtype,
                            out_dtype, st);
  R2N2_CHECK_ARG(in_dtype == R2N2_BF16, R2N2_ERR_BAD_DTYPE, "conv_fwd[umma]: operands must be bf16");
  R2N2_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, R2N2_ERR_BAD_SHAPE,
                 "conv_fwd[umma]: non-positive dimension");
  R2N2_CHECK_ARG((kd & 1) && (kh & 1) && (kw & 1), R2N2_ERR_BAD_SHAPE,
                 "conv_fwd[umma]: kernel extents must be odd");
  R2N2_CHECK_ARG(Cin % 16 == 0 && Cout % 16 == 0, R2N2_ERR_UNSUPPORTED,
                 "conv_fwd[umma]: Cin (%d) and Cout (%d) must be multiples of 16 (use ALGO_SIMT)", Cin, Cout);
  R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 R2N2_ERR_BAD_ALIGN, "conv_fwd[umma]: x / w must be 16-byte aligned");
  // the register epilogues move 32 bytes per instruction (256-bit global loads / stores)
  R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 31) == 0 &&
                     (!(flags & R2N2_EPI_RESIDUAL) || (reinterpret_cast<uintptr_t>(res) & 31) == 0) &&
                     (!(flags & R2N2_EPI_MASK) || (reinterpret_cast<uintptr_t>(mask) & 31) == 0),
                 R2N2_ERR_BAD_ALIGN, "conv_fwd[umma]: y / residual / mask must be 32-byte aligned");
  // 3x3 / 3x3x3 layers with <= 128 output channels: input loaded once, taps as descriptor views
  if (gru) {
    R2N2_CHECK_ARG(mode == 0 && !colsum && out_dtype == R2N2_BF16 && (flags & ~R2N2_EPI_BIAS) == 0, R2N2_ERR_UNSUPPORTED,
                   "conv_fwd_gru[umma]: plain geometry, bf16 tensors, no other epilogue flags");
  }
  if (!gru && mode == 0 && N > 0 && D > 0 && H > 0 && W > 0 && conv_fwd3_applicable(N, D, H, W, Cin, Cout, kd, kh, kw)) {
    int rc = conv_fwd3_umma(x, w, bias, mask, res, y, N, D, H, W, Cin, Cout, kd, kh, kw, flags, out_dtype, st);
    if (rc == R2N2_OK && colsum) rc = r2n2_bias_grad(y, colsum, out_rows, Cout, out_dtype, 1, (void*)st);
    return rc;
  }
  EncodeTiledFn encode = get_encode();
  R2N2_CHECK_ARG(encode != nullptr, R2N2_ERR_CUDA, "conv_fwd[umma]: cuTensorMapEncodeTiled unavailable");

  FwdParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.kd = kd; p.kh = kh; p.kw = kw;
  choose_box_fwd(N, D, H, W, p.bw, p.bh, p.bd, p.bn);
  p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_d = (D + p.bd - 1) / p.bd; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  p.BN = Cout >= 256 ? 256 : Cout;
  // few M tiles (recurrent unit: 18, FC layers: 2): split N further so that the tiles cover the SMs; the
  // narrower MMAs are less efficient but the layer is latency-bound by its longest CTA, not throughput-bound
  while (p.BN >= 64 && p.BN % 32 == 0 && (long long)p.m_tiles * ((Cout + p.BN - 1) / p.BN) * 2 <= num_sms())
    p.BN /= 2;
  if (const char* e = getenv("R2N2_UMMA_BN")) { int v = atoi(e); if (v >= 16 && v <= 256 && v % 16 == 0 && v <= Cout) p.BN = v; }
  p.n_tiles = (Cout + p.BN - 1) / p.BN;
  // Multicast clusters for the layers that were split along N to cover the SMs (recurrent unit, FC layers, conv6): the N
  // tiles of one M tile re-load the same A boxes, and their CTAs are bound by the ~55 B/clk of TMA traffic one SM takes
  // (tools/waitstat.py: the producer warp is busy 93 % of the kernel, nobody waits).  In a cluster of C such CTAs each
  // loads 1/C of every A box and multicasts it: per-CTA traffic A/C + B.  Needs all clusters resident at once (one
  // tile per CTA, lockstep) -- C = 8 fits 16 clusters on a B200, so N is re-merged where that gives a resident shape.
  p.mc = 0; p.mc_rows = 0;
  int mc_sd = 0, mc_sn = 0;
  // MEASURED (CUDA-graph replay, no host launch time): no gain -- GRU hidden conv 13 us with or without (only C = 2 leaves
  // all 72 clusters resident), fc7 10 -> 11 us with C = 8, x-projection 17 -> 18 us: at ~10 us these kernels sit on their
  // launch + prologue + first-load latency floor, not on the A traffic.  Kept as an opt-in (R2N2_UMMA_MC=1), parity-tested.
  if (!gru && mode == 0 && getenv("R2N2_UMMA_MC") && atoi(getenv("R2N2_UMMA_MC")) && !getenv("R2N2_UMMA_BN")) {
    const int rows = p.bw * p.bh * p.bd * p.bn, c3 = p.bw * p.bh * p.bd, c2 = p.bw * p.bh;
    double best_traffic = 1e30;
    int bBN = 0, bC = 0;
    for (int BN = p.BN; BN <= 256 && BN <= Cout; BN *= 2) {
      if (Cout % BN) continue;
      const int nt = Cout / BN;
      if ((long long)p.m_tiles * nt > num_sms()) continue;
      for (int C = 8; C >= 2; C >>= 1) {
        if (nt % C || rows % C || (rows / C) % 8) continue;
        const int sr = rows / C;
        if (!(sr % c3 == 0 || (c3 % sr == 0 && sr % c2 == 0))) continue;
        if (p.m_tiles * nt / C > max_active_clusters(C)) continue;
        const double traffic = 128.0 / C + BN;              // rows of A + rows of B per K element and CTA
        if (traffic < best_traffic) { best_traffic = traffic; bBN = BN; bC = C; }
      }
    }
    if (bC && best_traffic < 0.8 * (128.0 + p.BN)) {
      p.BN = bBN;
      p.n_tiles = Cout / p.BN;
      p.mc = bC;
      p.mc_rows = rows / bC;
      if (p.mc_rows % c3 == 0) { mc_sn = p.mc_rows / c3; mc_sd = p.bd; }
      else { mc_sn = 1; mc_sd = p.mc_rows / c2; }
      for (int r = 0; r < 8; ++r) {
        if (p.mc_rows % c3 == 0) { p.mc_dn[r] = (signed char)(r * mc_sn); p.mc_dd[r] = 0; }
        else { const int per = c3 / p.mc_rows; p.mc_dn[r] = (signed char)(r / per); p.mc_dd[r] = (signed char)((r % per) * mc_sd); }
      }
    }
  }
  p.mode = mode;
  p.cls_tiles = p.m_tiles * p.n_tiles;
  p.total_tiles = p.cls_tiles * (mode == 1 ? 8 : 1);
  p.bn_cols = (p.BN + 31) / 32 * 32;
  p.kbox = Cin >= 64 ? 64 : (Cin >= 32 ? 32 : 16);
  if (Cin > 64 && Cin % 64 == 32 && Cin <= 160) p.kbox = 32;   // 96 = 3 x 32: no zero-filled half K block
  p.kchunks = (Cin + p.kbox - 1) / p.kbox;
  const int row_bytes = p.kbox * 2;
  p.a_bytes = (128 * row_bytes + 1023) / 1024 * 1024;
  p.b_bytes = (p.BN * row_bytes + 1023) / 1024 * 1024;
  const int sub_bytes = p.a_bytes + p.b_bytes;
  const int nq = kd * kh * kw * p.kchunks;
  // narrow N tiles are bound by the single MMA-issuing thread, not by the tensor pipe: run two
  // CTAs (two issuers) per SM there; wide tiles get the whole SM (deeper stages, 2 x 256 TMEM cols)
  // (measured, profiles/r01_umma_ctas_nsub_sweep.txt: N<=64 +40%, N=128 +5..18%; N=96 prefers one
  // CTA with two sub-blocks per stage because a single 28 KB sub-block is < 256 MMA cycles)
  const int sub_cycles = (p.kbox / 16) * (p.BN / 2);
  // (re-measured after the issue-sequence work: two CTAs are never slower up to BN = 128 -- conv2a 0.25 -> 0.23 ms)
  int ctas = p.BN <= 128 ? 2 : 1;
  (void)sub_cycles;
  if (const char* e = getenv("R2N2_UMMA_CTAS")) { int v = atoi(e); if (v == 1 || v == 2) ctas = v; }
  if (2 * p.bn_cols * ctas > 512) ctas = 1;
  // no more tiles than SMs (recurrent unit, FC layers, conv5/6): every CTA is alone on its SM anyway, so it gets the
  // whole shared memory -- 5 pipeline stages instead of 2 hide the TMA round trip of these latency-bound layers
  if (p.total_tiles <= num_sms() && !getenv("R2N2_UMMA_CTAS")) ctas = 1;
  if (p.mc > 1) ctas = 1;
  // epilogue through shared memory + TMA (bf16 outputs; UP2 writes strided positions and keeps the register path)
  // Worth its shared memory only where the register epilogue would be the critical path: thread-per-pixel
  // accesses cost ~66 L1 wavefront cycles per 16-byte instruction (2 stores + 2 loads per input tensor per
  // 16-column chunk per warp), to be compared with the MMA time of a tile.
  {
    const int nin = ((flags & R2N2_EPI_RESIDUAL) ? 1 : 0) + ((flags & R2N2_EPI_MASK) ? 1 : 0);
    const double epi_cycles = 4.0 * (p.BN / 16) * (2 + 2 * nin) * 66.0;
    const double mma_cycles = (double)nq * (p.kbox / 16) * (p.BN >= 128 ? p.BN / 2 : 32 + p.BN / 4);
    p.tma_epi = (out_dtype == R2N2_BF16 && mode != 1 && epi_cycles > 0.5 * mma_cycles) ? 1 : 0;
    if (const char* e = getenv("R2N2_TMA_EPI")) p.tma_epi = (atoi(e) != 0 && out_dtype == R2N2_BF16 && mode != 1) ? 1 : 0;
    // (measured: forcing the TMA epilogue just to fuse a column sum costs more than the separate bias-gradient
    // pass it saves -- conv1b dgrad 0.68 + 0.20 -> 0.97 ms, conv3b dgrad 0.19 + 0.03 -> 0.26 ms -- so the sum
    // is fused only where the TMA epilogue is chosen anyway)
    if (getenv("R2N2_FORCE_COLSUM_FUSE") && colsum && out_dtype == R2N2_BF16 && mode != 1 && Cout <= p.BN) p.tma_epi = 1;
    if (gru) p.tma_epi = 0;                 // the gate math lives in the register epilogue
  }
  p.gru = gru ? gru->mode : 0;
  for (int i = 0; i < 4; ++i) p.gin[i] = gru ? gru->in[i] : nullptr;
  for (int i = 0; i < 3; ++i) p.gout[i] = gru ? gru->out[i] : nullptr;
  const bool fuse_colsum = colsum && p.tma_epi && Cout <= p.BN && !getenv("R2N2_NO_COLSUM_FUSE");
  p.SC = (p.BN % 64 == 0 && ctas == 1) ? 64 : (p.BN % 32 == 0 ? 32 : 16);
  p.slab_bytes = 128 * p.SC * 2;
  const int slab_region = p.tma_epi ? ((flags & R2N2_EPI_MASK) ? 4 : 2) * p.slab_bytes : 0;
  const int bias_bytes = ((flags & R2N2_EPI_BIAS) && Cout <= 1024 && !getenv("R2N2_NO_SMEM_BIAS")) ? (Cout * 4 + 15) / 16 * 16 : 0;
  const int epi_bytes = slab_region + (fuse_colsum ? 2 * p.BN * 4 : 0) + bias_bytes;
  p.colsum = fuse_colsum ? colsum : nullptr;
  const int smem_budget = (ctas == 2 ? 104 : 208) * 1024 - epi_bytes;
  const int stage_cap = ctas == 2 ? (smem_budget / 2 < 48 * 1024 ? smem_budget / 2 : 48 * 1024) : 64 * 1024;
  // sub-blocks per stage: aim at >= 512 cycles of MMA per barrier round, stage <= stage_cap
  const int mma_cycles_per_sub = (p.kbox / 16) * (p.BN / 2);
  int nsub = (512 + mma_cycles_per_sub - 1) / mma_cycles_per_sub;
  if (nsub > 8) nsub = 8;
  while (nsub > 1 && nsub * sub_bytes > stage_cap) --nsub;
  if (nsub > nq) nsub = nq;
  // a stage count that divides the tile's sub-blocks keeps every stage full, i.e. on the compile-time issue sequence
  // (the generic loop costs the issuing thread ~100 cycles per MMA on the N = 32 / 64 tiles of the small layers)
  if (mode != 1 && Cin % p.kbox == 0 && nq % nsub != 0 && !getenv("R2N2_UMMA_NO_DIV_NSUB"))
    for (int v = nsub; v >= 1; --v)
      if (nq % v == 0 && (v <= 4 || v == 6 || v == 8)) { if (2 * v >= nsub || v >= 3) nsub = v; break; }
  if (const char* e = getenv("R2N2_UMMA_NSUB")) { int v = atoi(e); if (v >= 1 && v <= 8) nsub = v < nq ? v : nq; }
  p.nsub = nsub;
  p.stage_bytes = nsub * sub_bytes;
  int stages = smem_budget / p.stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  if (const char* e = getenv("R2N2_UMMA_STAGES")) { int v = atoi(e); if (v >= 2 && v <= 12) stages = v; }
  p.stages = stages;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.bn_cols) p.tmem_cols *= 2;
  p.flags = flags & (15 | R2N2_EPI_RES_PRE);
  if (flags & R2N2_EPI_RES_PRE) {          // partial-sum accumulation (fp32 outputs): register epilogue only
    R2N2_CHECK_ARG(out_dtype == R2N2_F32 && (flags & R2N2_EPI_RESIDUAL), R2N2_ERR_UNSUPPORTED,
                   "conv_fwd[umma]: RES_PRE needs RESIDUAL and an fp32 output");
    p.tma_epi = 0;
  }
  p.err_word = get_err_word();
  // [stages][barriers ...][1024-aligned epilogue slabs]
  p.epi_off = (int)(((size_t)p.stages * p.stage_bytes + (2 * p.stages + 6) * 8 + 16 + 1023) / 1024 * 1024);
  p.csum_off = p.epi_off + slab_region;
  p.bias_off = bias_bytes ? p.csum_off + (fuse_colsum ? 2 * p.BN * 4 : 0) : -1;
  const size_t smem_bytes = (size_t)p.epi_off + epi_bytes + 1024;
  R2N2_CHECK_ARG(smem_bytes <= 227 * 1024, R2N2_ERR_UNSUPPORTED, "conv_fwd[umma]: pipeline does not fit smem");

  CUtensorMap map_a, map_b;
  {
    // DOWN2 reads the (2D,2H,2W) input with a traversal stride of 2: a box of 2*b elements lands as b rows
    const int sc = (mode == 2) ? 2 : 1;
    const long long Wi = (long long)W * sc, Hi = (long long)H * sc, Di = (long long)D * sc;
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)Wi * Cin * 2, (cuuint64_t)Hi * Wi * Cin * 2,
                             (cuuint64_t)Di * Hi * Wi * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kbox, (cuuint32_t)(p.bw * sc), (cuuint32_t)(p.bh * sc),
                         (cuuint32_t)((p.mc > 1 ? mc_sd : p.bd) * sc), (cuuint32_t)(p.mc > 1 ? mc_sn : p.bn)};
    cuuint32_t es[5] = {1, (cuuint32_t)sc, (cuuint32_t)sc, (cuuint32_t)sc, 1};
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
  CUtensorMap map_y = map_a, map_res = map_a, map_mask = map_a;
  if (p.tma_epi) {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2,
                             (cuuint64_t)D * H * W * Cout * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.SC, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd, (cuuint32_t)p.bn};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const void* ptrs[3] = {y, (flags & R2N2_EPI_RESIDUAL) ? res : y, (flags & R2N2_EPI_MASK) ? mask : y};
    CUtensorMap* maps[3] = {&map_y, &map_res, &map_mask};
    for (int i = 0; i < 3; ++i) {
      R2N2_CHECK_ARG((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, R2N2_ERR_BAD_ALIGN,
                     "conv_fwd[umma]: y / residual / mask must be 16-byte aligned");
      CUresult r = encode(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptrs[i]), dims, strides, box,
                          es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.SC * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      R2N2_CHECK_ARG(r == CUDA_SUCCESS, R2N2_ERR_CUDA, "conv_fwd[umma]: cuTensorMapEncodeTiled(epilogue %d) failed (%d)",
                     i, (int)r);
    }
  }
  int grid = num_sms() * ctas;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (getenv("R2N2_UMMA_DBG"))
    fprintf(stderr, "[umma fwd] mc %d(%d rows) spec %d tiles %d (m %d x n %d) box %dx%dx%dx%d BN %d kbox %d nsub %d stages %d ctas %d tma_epi %d SC %d smem %zu\n",
            p.mc, p.mc_rows, (int)(mode != 1 && Cin % p.kbox == 0 && nq % p.nsub == 0),
            p.total_tiles, p.m_tiles, p.n_tiles, p.bw, p.bh, p.bd, p.bn, p.BN, p.kbox, p.nsub, p.stages, ctas,
            p.tma_epi, p.SC, smem_bytes);
  R2N2_CHECK_ARG(out_dtype == R2N2_BF16 || out_dtype == R2N2_F32, R2N2_ERR_BAD_DTYPE, "conv_fwd[umma]: bad out dtype %d",
                 out_dtype);
  const FwdLaunch L{map_a, map_b, map_y, map_res, map_mask, bias, mask, res, y, p, grid, smem_bytes, st};
  // compile-time issue sequence when every stage holds nsub full sub-blocks (not for UP2: taps vary per class)
  const bool spec = mode != 1 && Cin % p.kbox == 0 && nq % p.nsub == 0 && !getenv("R2N2_UMMA_GENERIC");
  const int nk16 = p.kbox / 16;
  int rc;
  if (spec && nk16 == 4 && p.nsub == 1) rc = launch_fwd_dt<4, 1>(L, out_dtype);
  else if (spec && nk16 == 4 && p.nsub == 2) rc = launch_fwd_dt<4, 2>(L, out_dtype);
  else if (spec && nk16 == 2 && p.nsub == 3) rc = launch_fwd_dt<2, 3>(L, out_dtype);
  else if (spec && nk16 == 2 && p.nsub == 4) rc = launch_fwd_dt<2, 4>(L, out_dtype);
  else if (spec && nk16 == 4 && p.nsub == 3) rc = launch_fwd_dt<4, 3>(L, out_dtype);
  else if (spec && nk16 == 4 && p.nsub == 4) rc = launch_fwd_dt<4, 4>(L, out_dtype);
  else if (spec && nk16 == 4 && p.nsub == 6) rc = launch_fwd_dt<4, 6>(L, out_dtype);
  else if (spec && nk16 == 4 && p.nsub == 8) rc = launch_fwd_dt<4, 8>(L, out_dtype);
  else if (spec && nk16 == 2 && p.nsub == 6) rc = launch_fwd_dt<2, 6>(L, out_dtype);
  else if (spec && nk16 == 2 && p.nsub == 8) rc = launch_fwd_dt<2, 8>(L, out_dtype);
  else rc = launch_fwd_dt<0, 0>(L, out_dtype);
  if (rc != R2N2_OK) return rc;
  R2N2_CHECK_LAUNCH("conv_fwd[umma]");
  if (colsum && !fuse_colsum) return r2n2_bias_grad(y, colsum, out_rows, Cout, out_dtype, 1, (void*)st);
  return R2N2_OK;
}

}  // namespace r2n2
