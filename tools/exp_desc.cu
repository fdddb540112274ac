// Hardware experiment (sm_100a): can a tcgen05 shared-memory descriptor START on a row that is not a
// multiple of the swizzle atom (8 rows)?  That is what lets ONE halo tile in shared memory serve all
// filter taps of a convolution (tap shift = row offset of the operand view).  Also measures the
// issue rate of back-to-back SS-mode MMAs for several N.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o gpurun_out/exp_desc tools/exp_desc.cu
//   (always link the runtime dynamically and keep the binary under gpurun_out/, never in the tree)
//
// Operands are written by hand with the TMA swizzle (byte offset bits [4:6] ^= bits [7:9], masked to the
// swizzle width), values are small integers so fp32 accumulation is exact.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static __device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
static __device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
static __device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (clock64() - t0 > 400000000LL) break;
  }
}
static __device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
  int mode;          // 0: K-major A shifted by r0 rows; 1: MN-major B shifted by r0 rows (wgrad style);
                     // 2: MN-major A with LBO = lbo_rows rows (taps stacked along M), B MN-major unshifted
  int rb;            // row bytes: 128 / 64 / 32 (= swizzle width)
  int r0;            // row shift
  int base_off;      // descriptor base_offset field
  int lbo_rows;      // mode 2
  int N;             // MMA N
  int reps;          // timing mode: MMAs issued back to back (mode 3)
  int M;             // MMA M for mode 4 (64 / 128)
};

__host__ __device__ inline uint32_t swz(uint32_t off, int rb) {
  uint32_t mask = rb == 128 ? 7u : (rb == 64 ? 3u : 1u);
  return off ^ (((off >> 7) & mask) << 4);
}

constexpr int kRows = 512;    // rows in each operand buffer

__global__ void __launch_bounds__(192) exp_kernel(const __nv_bfloat16* ga, const __nv_bfloat16* gb, float* out,
                                                   long long* cycles, Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // kRows x 128 B max
  uint8_t* sb = smem + kRows * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kRows * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int epr = p.rb / 2;                 // elements per row
  // global [row][epr] -> smem row-major with swizzle
  for (int i = threadIdx.x; i < kRows * epr; i += blockDim.x) {
    int r = i / epr, k = i % epr;
    uint32_t off = (uint32_t)(r * p.rb + k * 2);
    *reinterpret_cast<__nv_bfloat16*>(sa + swz(off, p.rb)) = ga[i];
    *reinterpret_cast<__nv_bfloat16*>(sb + swz(off, p.rb)) = gb[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  const uint32_t lt = p.rb == 128 ? 2u : (p.rb == 64 ? 4u : 6u);
  const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
  if (warp == 1 && p.mode == 4) {
    // issue-rate test without any per-iteration ALU work: whole warp runs the loop, one elected lane issues
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | (((uint32_t)p.M >> 4) << 24);
    uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)((8u * p.rb) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
    uint64_t da = hi | (uint64_t)(((a0 + p.r0 * p.rb) & 0x3FFFF) >> 4);
    uint64_t db = hi | (uint64_t)((b0 & 0x3FFFF) >> 4);
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    long long t0 = clock64();
    if (pred) {
      for (int i = 0; i < p.reps; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma(tmem, da + 2u * (j & 3), db + 2u * (j & 3), idesc, 1u);
      }
      commit(smem_u32(bar));
    }
    __syncwarp();
    wait_bar(smem_u32(bar), 0);
    long long t1 = clock64();
    if (lane == 0) cycles[0] = t1 - t0;
  } else if (warp == 1) {
    if (lane == 0) {
      long long t0 = 0, t1 = 0;
      if (p.mode == 0) {
        // K-major: A rows = M, row = K elements.  D[m][n] = sum_k A[m + r0][k] * B[n][k]
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)((8u * p.rb) >> 4) << 32) | ((uint64_t)1 << 46) |
                      ((uint64_t)lt << 61);
        uint64_t da = hi | (uint64_t)(((a0 + p.r0 * p.rb) & 0x3FFFF) >> 4) | ((uint64_t)(p.base_off & 7) << 49);
        uint64_t db = hi | (uint64_t)((b0 & 0x3FFFF) >> 4);
        for (int k = 0; k < epr / 16; ++k) umma(tmem, da + 2u * k, db + 2u * k, idesc, k > 0);
      } else if (p.mode == 1) {
        // MN-major both: A = dy [K rows][128 co in chunks of epr], B = x [K rows + shift][N ci in chunks of epr]
        // buffers: chunk c of A lives at rows [c*64, c*64+64) of sa (64 K rows per chunk); same for B.
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t chunk = 64u * p.rb;
        uint64_t hi = ((uint64_t)((chunk >> 4) & 0x3FFF) << 16) | ((uint64_t)((8u * p.rb) >> 4) << 32) |
                      ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
        uint64_t da = hi | (uint64_t)((a0 & 0x3FFFF) >> 4);
        uint64_t db = hi | (uint64_t)(((b0 + p.r0 * p.rb) & 0x3FFFF) >> 4) | ((uint64_t)(p.base_off & 7) << 49);
        uint32_t kstep = (16u * p.rb) >> 4;
        for (int k = 0; k < 2; ++k) umma(tmem, da + (uint64_t)(kstep * k), db + (uint64_t)(kstep * k), idesc, k > 0);
      } else if (p.mode == 2) {
        // MN-major A with LBO = lbo_rows * rb: M chunk c reads the rows shifted by c*lbo_rows (stacked taps)
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t chunk = 64u * p.rb;
        uint64_t hia = ((uint64_t)(((p.lbo_rows * p.rb) >> 4) & 0x3FFF) << 16) | ((uint64_t)((8u * p.rb) >> 4) << 32) |
                       ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
        uint64_t hib = ((uint64_t)((chunk >> 4) & 0x3FFF) << 16) | ((uint64_t)((8u * p.rb) >> 4) << 32) |
                       ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
        uint64_t da = hia | (uint64_t)(((a0 + p.r0 * p.rb) & 0x3FFFF) >> 4) | ((uint64_t)(p.base_off & 7) << 49);
        uint64_t db = hib | (uint64_t)((b0 & 0x3FFFF) >> 4);
        uint32_t kstep = (16u * p.rb) >> 4;
        for (int k = 0; k < 2; ++k) umma(tmem, da + (uint64_t)(kstep * k), db + (uint64_t)(kstep * k), idesc, k > 0);
      } else {
        // timing: reps x (K-major, M=128, N) MMAs back to back on the same operands
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)((8u * p.rb) >> 4) << 32) | ((uint64_t)1 << 46) |
                      ((uint64_t)lt << 61);
        uint64_t da = hi | (uint64_t)((a0 & 0x3FFFF) >> 4);
        uint64_t db = hi | (uint64_t)((b0 & 0x3FFFF) >> 4);
        t0 = clock64();
        for (int i = 0; i < p.reps; ++i) umma(tmem, da + 2u * (i & 3) + 64u * ((i >> 2) & 7), db + 2u * (i & 3), idesc, i > 0);
      }
      commit(smem_u32(bar));
      if (p.mode == 3) {
        wait_bar(smem_u32(bar), 0);
        t1 = clock64();
        cycles[0] = t1 - t0;
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    const int quad = warp & 3;
    wait_bar(smem_u32(bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = quad * 32 + lane;
    for (int c = 0; c < p.N; c += 16) {
      uint32_t v[16];
      ld16(tmem + ((uint32_t)(quad * 32) << 16) + c, v);
      for (int j = 0; j < 16; ++j) out[row * 256 + c + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

static float bf(int v) { return (float)v; }

int main() {
  const size_t smem_bytes = 2 * kRows * 128 + 64 + 1024;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  __nv_bfloat16 *ga, *gb;
  float* out;
  long long* cyc;
  cudaMalloc(&ga, kRows * 64 * 2);
  cudaMalloc(&gb, kRows * 64 * 2);
  cudaMalloc(&out, 128 * 256 * 4);
  cudaMalloc(&cyc, 8);
  std::vector<float> hout(128 * 256);
  for (int rb : {128, 64, 32}) {
    const int epr = rb / 2;
    std::vector<int> A(kRows * epr), B(kRows * epr);
    std::vector<__nv_bfloat16> hA(kRows * epr), hB(kRows * epr);
    srand(1234 + rb);
    for (int i = 0; i < kRows * epr; ++i) {
      A[i] = rand() % 7 - 3; B[i] = rand() % 5 - 2;
      hA[i] = __float2bfloat16(bf(A[i])); hB[i] = __float2bfloat16(bf(B[i]));
    }
    cudaMemcpy(ga, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(gb, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 3; ++mode) {
      const int N = (mode == 0) ? 64 : epr * 2 > 64 ? 64 : epr * 2;   // MN-major N: up to 2 chunks
      for (int r0 : {0, 1, 2, 3, 5, 7, 8, 9, 34, 35}) {
        for (int bo_sel = 0; bo_sel < 2; ++bo_sel) {
          for (int lbo_rows : {1, 2, 34}) {
            if (mode != 2 && lbo_rows != 1) continue;
            Params p;
            p.mode = mode; p.rb = rb; p.r0 = r0; p.base_off = bo_sel ? ((r0 * rb) >> 7) & 7 : 0;
            if (bo_sel && p.base_off == 0) continue;
            p.lbo_rows = lbo_rows; p.N = N; p.reps = 0; p.M = 128;
            cudaMemset(out, 0, 128 * 256 * 4);
            exp_kernel<<<1, 192, smem_bytes>>>(ga, gb, out, cyc, p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hout.data(), out, hout.size() * 4, cudaMemcpyDeviceToHost);
            long long bad = 0;
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
              for (int n = 0; n < N; ++n) {
                double ref = 0;
                if (mode == 0) {
                  for (int k = 0; k < epr; ++k) ref += (double)A[(m + r0) * epr + k] * B[n * epr + k];
                } else if (mode == 1) {
                  // A element (m,k): chunk m/epr at rows [64*chunk ...], row k, col m%epr
                  for (int k = 0; k < 32; ++k)
                    ref += (double)A[((m / epr) * 64 + k) * epr + m % epr] *
                           B[((n / epr) * 64 + k + r0) * epr + n % epr];
                } else {
                  for (int k = 0; k < 32; ++k)
                    ref += (double)A[(r0 + (m / epr) * lbo_rows + k) * epr + m % epr] *
                           B[((n / epr) * 64 + k) * epr + n % epr];
                }
                double err = fabs(ref - hout[m * 256 + n]);
                if (err > 1e-3) ++bad;
                if (err > maxerr) maxerr = err;
              }
            printf("rb %3d mode %d r0 %2d base_off %d lbo_rows %2d N %3d : %s (bad %lld / %d, max err %.1f)\n", rb, mode,
                   r0, p.base_off, lbo_rows, N, bad ? "MISMATCH" : "ok", bad, 128 * N, maxerr);
          }
        }
      }
    }
  }
  // MMA issue rate
  for (int rb : {128, 64, 32})
    for (int N : {16, 32, 64, 96, 128, 256}) {
      for (int reps : {256, 1024}) {
        Params p;
        p.mode = 3; p.rb = rb; p.r0 = 0; p.base_off = 0; p.lbo_rows = 1; p.N = N; p.reps = reps; p.M = 128;
        exp_kernel<<<1, 192, smem_bytes>>>(ga, gb, out, cyc, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("issue rate rb %3d N %3d reps %4d : %lld cycles = %.1f / MMA\n", rb, N, reps, h, (double)h / reps);
      }
    }
  for (int rbb : {128, 64, 32})
    for (int r0 : {0, 1, 3, 4, 8})
      for (int N : {32, 64, 128}) {
        Params p;
        p.mode = 4; p.rb = rbb; p.r0 = r0; p.base_off = 0; p.lbo_rows = 1; p.N = N; p.reps = 1024; p.M = 128;
        exp_kernel<<<1, 192, smem_bytes>>>(ga, gb, out, cyc, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("K-major A start row %d rb %3d N %3d : %.1f cycles / MMA\n", r0, rbb, N, (double)h / p.reps);
      }
  for (int M : {128, 64})
    for (int N : {16, 32, 64, 128, 256}) {
      Params p;
      p.mode = 4; p.rb = 128; p.r0 = 0; p.base_off = 0; p.lbo_rows = 1; p.N = N; p.reps = 1024; p.M = M;
      exp_kernel<<<1, 192, smem_bytes>>>(ga, gb, out, cyc, p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long h;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("unrolled issue rate M %3d N %3d reps %4d : %lld cycles = %.1f / MMA\n", M, N, p.reps, h, (double)h / p.reps);
    }
  return 0;
}
