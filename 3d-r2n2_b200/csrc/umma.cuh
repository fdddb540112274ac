// Shared device/host helpers of the tcgen05 / TMA kernels (conv_umma.cu, conv_umma_fwd.cu).
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace r2n2 {

// --------------------------------------------------------------------------------- PTX wrappers
static __device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
static __device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must never hang the GPU.  On timeout the CTA-wide abort flag is
// raised (every later wait returns at once) and a global error word is set for the host.
// The abort flag lives in shared memory.  Read through the generic `volatile int*` it became LD.E.STRONG.SYS, a LONG-scoreboard
// operation: an epilogue warp that had issued its residual / mask loads (global, ~1 us) and then blocked on the
// accumulator barrier could not evaluate `if (*abort_flag)` before those loads had landed too (ncu source view of the conv1b
// data gradient: 26 % of the stall samples behind these branches and the first instruction after the wait) -- the
// prefetch was serialised with the wait it was meant to overlap.  An explicit ld.volatile.shared is a short-scoreboard op.
static __device__ __forceinline__ int abort_read(volatile int* abort_flag) {
  int v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];"
               : "=r"(v)
               : "r"((uint32_t)__cvta_generic_to_shared(const_cast<int*>(abort_flag)))
               : "memory");
  return v;
}
static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag,
                                          int* err_word) {
  if (mbar_try_wait(bar, parity)) return;
  if (abort_read(abort_flag)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL || abort_read(abort_flag)) {
      *abort_flag = 1;
      if (err_word) atomicExch(err_word, 1);
      return;
    }
  }
}
#ifdef R2N2_WAITSTAT
// DIAGNOSTIC BUILD ONLY (libr2n2_b200_ws.so, tools/waitstat.py): cycles each warp spends inside mbar_wait, per
// barrier slot, written at kernel exit to the device buffer behind the error word:
//   out[(block * 16 + warp) * 40 + i]: i < 38 = wait cycles on barrier i (8-byte slots from the kernel's first
//   mbarrier), 38 = cycles from kernel entry to exit, 39 = number of waits that actually blocked.
struct WaitStat {
  long long bin[38];
  long long t0, blocked;
  uint32_t base;
};
static __device__ __forceinline__ void ws_begin(WaitStat& w, const void* bars) {
  for (int i = 0; i < 38; ++i) w.bin[i] = 0;
  w.blocked = 0;
  w.base = smem_u32(bars);
  w.t0 = clock64();
}
static __device__ __forceinline__ void ws_end(const WaitStat& w, int* err_word) {
  if ((threadIdx.x & 31) != 0 || err_word == nullptr) return;
  const unsigned blk = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const unsigned warp = threadIdx.x >> 5;
  if (blk >= 512 || warp >= 16) return;
  long long* out = reinterpret_cast<long long*>(err_word) + 8 + (size_t)(blk * 16 + warp) * 40;
  for (int i = 0; i < 38; ++i) out[i] = w.bin[i];
  out[38] = clock64() - w.t0;
  out[39] = w.blocked;
}
static __device__ __forceinline__ void mbar_wait_ws(WaitStat& w, uint32_t bar, uint32_t parity, volatile int* abort_flag,
                                             int* err_word) {
  // (mbarrier.try_wait may itself suspend the thread for a while before it returns true: time the whole call)
  const long long t = clock64();
  mbar_wait(bar, parity, abort_flag, err_word);
  const long long dt = clock64() - t;
  uint32_t idx = (bar - w.base) >> 3;
  if (idx > 37) idx = 37;
  w.bin[idx] += dt;
  if (dt > 200) ++w.blocked;
}
#define R2N2_WS_BEGIN(bars_ptr) WaitStat ws_; ws_begin(ws_, bars_ptr)
#define R2N2_WS_END(err_word_ptr) ws_end(ws_, err_word_ptr)
// section timers (bins 30..36 by convention): R2N2_WS_T0() ... R2N2_WS_ACC(slot) adds the cycles since the last mark
#define R2N2_WS_T0() long long ws_t_ = clock64()
#define R2N2_WS_ACC(slot) do { const long long ws_n_ = clock64(); ws_.bin[slot] += ws_n_ - ws_t_; ws_t_ = ws_n_; } while (0)
#define mbar_wait(bar, parity, af, ew) mbar_wait_ws(ws_, bar, parity, af, ew)
#else
#define R2N2_WS_BEGIN(bars_ptr)
#define R2N2_WS_END(err_word_ptr)
#define R2N2_WS_T0()
#define R2N2_WS_ACC(slot)
#endif
static __device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
static __device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// TMA tensor store (shared -> global, out-of-bounds elements are clipped) and its bulk-group bookkeeping
static __device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(map)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
static __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
static __device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
static __device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// byte offset inside a 1024-aligned TMA tile with 128/64/32-byte rows -> swizzled offset (mask 7/3/1)
static __device__ __forceinline__ uint32_t swz_off(uint32_t off, uint32_t mask) {
  return off ^ (((off >> 7) & mask) << 4);
}

static __device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
static __device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// ---- CTA pair (cluster of 2, tcgen05 cta_group::2): one M = 256 MMA spans both SMs' tensor cores, each CTA
// supplies its own 128 A rows and HALF of the B (N) rows from its own shared memory at the same offsets.
static __device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same location in CTA `rank` of the cluster
static __device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
static __device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
static __device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes)
               : "memory");
}
static __device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // relaxed: the arrivals that use this (accumulator hand-backs of the pair kernel's epilogue warps) order TMEM reads, which
  // tcgen05.wait::ld + tcgen05.fence::before_thread_sync already did; `.release.cluster` made every hand-back a MEMBAR that
  // waited for the warp's outstanding global STORES (ncu source view: 10 % of the conv1b data gradient's stall samples)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA loads into THIS CTA's shared memory whose completion is signalled on an mbarrier of the pair's leader CTA
static __device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0,
                                                int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
static __device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
static __device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TMA multicast: the box lands at the same shared-memory offset in every CTA of `mask`, and each of them gets the
// complete_tx on its own mbarrier at the same offset
static __device__ __forceinline__ void tma_load_5d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, uint16_t mask, int c0,
                                               int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%4, %5, %6, %7, %8}], [%2], %3;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// tcgen05.commit of a single-CTA MMA stream that arrives on the barrier at this offset in every CTA of `mask`
static __device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs
static __device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
static __device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
static __device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15])
      : "r"(taddr));
}
static __device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// one elected lane of a fully converged warp (the loops below are executed by the WHOLE warp so
// that ptxas keeps descriptors / barrier addresses in uniform registers; running them under
// `if (lane == 0)` made every operand a vector register that needed R2UR moves per instruction)
static __device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
static __device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
static __device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// shared-memory matrix descriptor, K-major, swizzled (canonical layouts of the tcgen05 ISA):
//   start address >>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
static __device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t sbo_bytes,
                                                     uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// 16 consecutive output channels of one pixel: 2 x 16-byte stores (bf16) / 4 x 16-byte (fp32)
static __device__ __forceinline__ void store16(__nv_bfloat16* dst, const float (&f)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  // one 256-bit store (sm_100: STG.256): a warp's thread-per-pixel accesses cost one L1 wavefront per lane and
  // instruction, so 32 bytes per instruction halve the epilogue's LSU time (dst is 32-byte aligned: channel counts
  // are multiples of 16 and the dispatchers check the base pointers)
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
static __device__ __forceinline__ void store16(float* dst, const float (&f)[16]) {
#pragma unroll
  for (int j = 0; j < 2; ++j)
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 8 * j), "f"(f[8 * j]), "f"(f[8 * j + 1]),
                 "f"(f[8 * j + 2]), "f"(f[8 * j + 3]), "f"(f[8 * j + 4]), "f"(f[8 * j + 5]), "f"(f[8 * j + 6]), "f"(f[8 * j + 7])
                 : "memory");
}


// 16 consecutive channels of one pixel as stored in memory (raw 16-byte loads; converted on use)
template <typename T>
struct Row16;
template <>
struct Row16<__nv_bfloat16> {
  uint4 a, b;
  __device__ __forceinline__ void zero() { a = make_uint4(0, 0, 0, 0); b = a; }
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
  }
  __device__ __forceinline__ void unpack(float (&f)[16]) const {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
      float2 t = __bfloat1622float2(h);
      f[2 * j] = t.x; f[2 * j + 1] = t.y;
    }
  }
};
template <>
struct Row16<float> {
  float4 q[4];
  __device__ __forceinline__ void zero() { q[0] = q[1] = q[2] = q[3] = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float* p) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(q[2 * j].x), "=f"(q[2 * j].y), "=f"(q[2 * j].z), "=f"(q[2 * j].w), "=f"(q[2 * j + 1].x),
                     "=f"(q[2 * j + 1].y), "=f"(q[2 * j + 1].z), "=f"(q[2 * j + 1].w)
                   : "l"(p + 8 * j));
  }
  __device__ __forceinline__ void unpack(float (&f)[16]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[4 * j] = q[j].x; f[4 * j + 1] = q[j].y; f[4 * j + 2] = q[j].z; f[4 * j + 3] = q[j].w; }
  }
};

// Epilogue of one accumulator row (one pixel) over the 16-column chunks [c_begin, c_end) of a tile:
// TMEM -> bias -> LeakyReLU -> + residual -> x LeakyReLU'(mask) -> 16-byte stores.  The residual / mask
// rows of chunk c+1 are fetched while chunk c is processed, and those of the first chunk are issued
// by epi_prefetch() BEFORE the accumulator-ready wait, so the global-load latency overlaps the MMAs.
// res / mask / y point at channel 0 of this pixel; `ok` = pixel in bounds; co0 = first channel of the tile.
template <typename TOut>
struct EpiPrefetch {
  Row16<TOut> r, m;
};
template <typename TOut>
static __device__ __forceinline__ void epi_prefetch(EpiPrefetch<TOut>& pf, int flags, bool ok, const TOut* res,
                                                     const TOut* mask, int co, int Cout) {
  if (ok && co < Cout) {
    if (flags & R2N2_EPI_RESIDUAL) pf.r.load(res + co);
    if (flags & R2N2_EPI_MASK) pf.m.load(mask + co);
  }
}
template <typename TOut>
static __device__ __forceinline__ void epi_row(EpiPrefetch<TOut>& pf, uint32_t taddr, int c_begin, int c_end, int co0,
                                                int Cout, bool ok, bool empty_acc, int flags,
                                                const float* __restrict__ bias, const TOut* res, const TOut* mask,
                                                TOut* y) {
  for (int c = c_begin; c < c_end; c += 16) {
    const EpiPrefetch<TOut> cur = pf;
    if (c + 16 < c_end) epi_prefetch(pf, flags, ok, res, mask, co0 + c + 16, Cout);
    uint32_t v[16];
    tmem_ld16(taddr + (uint32_t)c, v);
    const int co = co0 + c;
    if (!ok || co >= Cout) continue;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = empty_acc ? 0.f : __uint_as_float(v[j]);
    if ((flags & (R2N2_EPI_RESIDUAL | R2N2_EPI_RES_PRE)) == (R2N2_EPI_RESIDUAL | R2N2_EPI_RES_PRE)) {
      float r[16];
      cur.r.unpack(r);
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] += r[j];
    }
    if (flags & R2N2_EPI_BIAS) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + co + j);     // (generic: the kernels pass a shared-memory copy)
        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
      }
    }
    if (flags & R2N2_EPI_LRELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = lrelu(f[j]);
    }
    if ((flags & (R2N2_EPI_RESIDUAL | R2N2_EPI_RES_PRE)) == R2N2_EPI_RESIDUAL) {
      float r[16];
      cur.r.unpack(r);
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] += r[j];
    }
    if (flags & R2N2_EPI_MASK) {
      float m[16];
      cur.m.unpack(m);
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] *= lrelu_slope(m[j]);
    }
    store16(y + co, f);
  }
}

static __device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                      uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

static __device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}


constexpr int kUmmaThreads = 192;
constexpr int kEpiWarps = 8;                 // conv_umma_fwd.cu / conv_umma_fwd3.cu: two epilogue warps per TMEM lane quadrant

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
int* get_err_word();
CUtensorMapSwizzle swizzle_for(int row_bytes);

}  // namespace r2n2
