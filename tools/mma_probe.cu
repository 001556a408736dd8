// mma_probe.cu — raw tcgen05.mma.kind::tf32 issue-rate probe for sm_100a (diagnosis tool, not product code).
// One CTA per SM; one thread issues `iters` groups of MMAs on zero-filled shared-memory operands with no
// TMA / mbarrier traffic in the loop, then a single commit.  Prints clocks per MMA for several issue patterns,
// which separates "what the tensor pipe can do with SS operands" from pipeline/barrier overheads in the GEMM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 28); ++it) if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lt << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode bits: see main()
struct Params { int mode, iters, N, commit_every; long long* out; };

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the operand area (160 KB)
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int N = p.N;
  const bool b_mn = (p.mode & 1) != 0;        // B operand MN-major (SW128 base-32B layout), else K-major
  const bool three = (p.mode & 2) != 0;       // 3xTF32 pattern: (lo,hi)->cross, (hi,lo)->cross, (hi,hi)->main
  const bool merged = (p.mode & 4) != 0;      // all three into ONE accumulator
  const bool bf16 = (p.mode & 8) != 0;        // kind::f16 (bf16 operands), K = 16 per MMA
  const bool a_mn = (p.mode & 16) != 0;       // A operand MN-major
  const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768, b_lo = base + 32768 + 32768;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, N, a_mn, b_mn, bf16 ? 1 : 2);
      const uint32_t A_LBO = a_mn ? 32 * 128 : 16, A_SBO = a_mn ? 512 : 1024, A_LT = a_mn ? 1 : 2, A_KS = a_mn ? 1024 : 32;
      const uint32_t B_LBO = b_mn ? 32 * 128 : 16, B_SBO = b_mn ? 512 : 1024, B_LT = b_mn ? 1 : 2, B_KS = b_mn ? 1024 : 32;
      const uint32_t t_small = tmem, t_main = tmem + (uint32_t)N;
      t0 = clock64();
      int since = 0;
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dah = make_smem_desc(a_hi + j * A_KS, A_LBO, A_SBO, A_LT);
          const uint64_t dal = make_smem_desc(a_lo + j * A_KS, A_LBO, A_SBO, A_LT);
          const uint64_t dbh = make_smem_desc(b_hi + j * B_KS, B_LBO, B_SBO, B_LT);
          const uint64_t dbl = make_smem_desc(b_lo + j * B_KS, B_LBO, B_SBO, B_LT);
          if (bf16) {
            umma_bf16(t_main, dah, dbh, idesc, 1);
          } else if (three) {
            umma_tf32(merged ? t_main : t_small, dal, dbh, idesc, 1);
            umma_tf32(merged ? t_main : t_small, dah, dbl, idesc, 1);
            umma_tf32(t_main, dah, dbh, idesc, 1);
          } else {
            umma_tf32(t_main, dah, dbh, idesc, 1);
          }
        }
        if (p.commit_every > 0 && ++since == p.commit_every) { umma_commit(smem_u32(&bars[1])); since = 0; }
      }
      umma_commit(smem_u32(&bars[0]));
      mbar_wait(smem_u32(&bars[0]), 0);
      t1 = clock64();
      p.out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}


// ---- cta_group::2 probe: cluster of 2 CTAs, leader issues M=256 x N MMAs; each CTA holds 128 A rows and N/2 B rows.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int N = p.N;
  const bool b_mn = (p.mode & 1) != 0, three = (p.mode & 2) != 0, merged = (p.mode & 4) != 0, a_mn = (p.mode & 16) != 0;
  const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768, b_lo = base + 32768 + 16384;   // B halves: <= 128 rows
  if (warp == 1) {
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = make_idesc(256, N, a_mn, b_mn, 2);
      const uint32_t A_LBO = a_mn ? 32 * 128 : 16, A_SBO = a_mn ? 512 : 1024, A_LT = a_mn ? 1 : 2, A_KS = a_mn ? 1024 : 32;
      const uint32_t B_LBO = b_mn ? 32 * 128 : 16, B_SBO = b_mn ? 512 : 1024, B_LT = b_mn ? 1 : 2, B_KS = b_mn ? 1024 : 32;
      const uint32_t t_small = tmem, t_main = merged ? tmem : tmem + (uint32_t)N;
      long long t0 = clock64();
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dah = make_smem_desc(a_hi + j * A_KS, A_LBO, A_SBO, A_LT);
          const uint64_t dal = make_smem_desc(a_lo + j * A_KS, A_LBO, A_SBO, A_LT);
          const uint64_t dbh = make_smem_desc(b_hi + j * B_KS, B_LBO, B_SBO, B_LT);
          const uint64_t dbl = make_smem_desc(b_lo + j * B_KS, B_LBO, B_SBO, B_LT);
          if (three) {
            umma_tf32_2sm(t_small, dal, dbh, idesc, 1);
            umma_tf32_2sm(t_small, dah, dbl, idesc, 1);
            umma_tf32_2sm(t_main, dah, dbh, idesc, 1);
          } else {
            umma_tf32_2sm(t_main, dah, dbh, idesc, 1);
          }
        }
      }
      umma_commit_2sm(smem_u32(&bars[0]), 0x3);
      mbar_wait(smem_u32(&bars[0]), 0);
      p.out[blockIdx.x] = clock64() - t0;
    } else if (lane == 0) {
      mbar_wait(smem_u32(&bars[0]), 0);     // peer: wait for the leader's multicast commit
      p.out[blockIdx.x] = 0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* out;
  cudaMalloc(&out, sizeof(long long) * sms);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Case { const char* name; int mode, N, commit_every; };
  Case cases[] = {
      {"tf32 1xMMA  K/K   N=128", 0, 128, 0},
      {"tf32 1xMMA  K/MN  N=128", 1, 128, 0},
      {"tf32 1xMMA  MN/MN N=128", 17, 128, 0},
      {"tf32 3xMMA  K/K   N=128 (2 acc)", 2, 128, 0},
      {"tf32 3xMMA  K/MN  N=128 (2 acc)", 3, 128, 0},
      {"tf32 3xMMA  MN/MN N=128 (2 acc)", 19, 128, 0},
      {"tf32 3xMMA  K/K   N=128 (merged)", 6, 128, 0},
      {"tf32 3xMMA  K/K   N=128 (2 acc) commit/k-block", 2, 128, 1},
      {"tf32 1xMMA  K/K   N=256", 0, 256, 0},
      {"tf32 3xMMA  K/K   N=256 (merged)", 6, 256, 0},
      {"tf32 3xMMA  K/MN  N=256 (merged)", 7, 256, 0},
      {"tf32 1xMMA  K/K   N=64", 0, 64, 0},
      {"bf16 1xMMA  K/K   N=128", 8, 128, 0},
      {"bf16 1xMMA  K/K   N=256", 8, 256, 0},
  };
  for (const Case& c : cases) {
    Params p{c.mode, iters, c.N, c.commit_every, out};
    for (int rep = 0; rep < 2; ++rep) {
      probe_kernel<<<sms, 128, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    long long h[256];
    cudaMemcpy(h, out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 62;
    for (int i = 0; i < sms; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    const bool three = (c.mode & 2) != 0, bf16 = (c.mode & 8) != 0;
    const double mmas = (double)iters * 4 * (three && !bf16 ? 3 : 1);
    const double kper = bf16 ? 16 : 8;
    printf("%-50s clk/MMA min %.1f max %.1f   -> %.0f flop/clk/SM (peak 128x%dx%g per 64*%d/128 clk)\n", c.name, mn / mmas, mx / mmas,
           2.0 * 128 * c.N * kper * mmas / mx, c.N, kper, c.N);
  }

  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Case cases2[] = {
      {"2cta tf32 1xMMA K/K   M=256 N=128", 0, 128, 0},
      {"2cta tf32 3xMMA K/K   M=256 N=128 (2 acc)", 2, 128, 0},
      {"2cta tf32 3xMMA K/MN  M=256 N=128 (2 acc)", 3, 128, 0},
      {"2cta tf32 3xMMA MN/MN M=256 N=128 (2 acc)", 19, 128, 0},
      {"2cta tf32 3xMMA K/K   M=256 N=256 (2 acc)", 2, 256, 0},
      {"2cta tf32 3xMMA K/K   M=256 N=256 (merged)", 6, 256, 0},
      {"2cta tf32 3xMMA K/MN  M=256 N=256 (merged)", 7, 256, 0},
      {"2cta tf32 3xMMA MN/MN M=256 N=256 (merged)", 23, 256, 0},
      {"2cta tf32 3xMMA K/K   M=256 N=192 (merged)", 6, 192, 0},
  };
  for (const Case& c : cases2) {
    Params p{c.mode, iters, c.N, c.commit_every, out};
    for (int rep = 0; rep < 2; ++rep) {
      probe2_kernel<<<sms / 2 * 2, 128, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    long long h[256];
    cudaMemcpy(h, out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const bool three = (c.mode & 2) != 0;
    const double mmas = (double)iters * 4 * (three ? 3 : 1);
    printf("%-50s clk/MMA %.1f   -> %.0f flop/clk/SM (ideal %d clk/MMA)\n", c.name, mx / mmas,
           2.0 * 256 * c.N * 8 * mmas / mx / 2, c.N / 2);
  }
  return 0;
}
