// Which coding of the UMMA issue loop is cheap?  (sm_100a; companion of umma_probe.cu)
//   MODE 0: `if (lane == 0)` around a runtime loop (the round-1 production coding)
//   MODE 1: converged warp, elect.sync inside the asm, runtime loop
//   MODE 2: converged warp, elect.sync inside the asm, loop unrolled x8, descriptor advanced by constants
//   MODE 3: `if (lane == 0)`, loop unrolled x8
//   MODE 4: converged warp, ONE elect.sync per 8 UMMAs (single asm block issuing 8 tcgen05.mma)
// CH = number of accumulators cycled through.  Reports cycles per UMMA on the issuing thread.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fastvocoder_b200/csrc/fv_tc.cuh"

namespace fv {
std::atomic<long long> g_launches{0};
std::atomic<long long> g_tc_launches{0};
}
using namespace fv;

__device__ __forceinline__ void umma_elect(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
// 8 UMMAs, A descriptor advanced by `astep` (16-B units) each, all accumulating
__device__ __forceinline__ void umma_elect8(uint32_t d, uint64_t ad, uint64_t astep, uint64_t bd, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b64 a;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 a, %1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t"
      "add.u64 a, a, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t"
      "add.u64 a, a, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t"
      "add.u64 a, a, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t"
      "add.u64 a, a, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t"
      "add.u64 a, a, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t"
      "add.u64 a, a, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t"
      "add.u64 a, a, %2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, %3, %4, 1;\n\t}" ::"r"(d),
      "l"(ad), "l"(astep), "l"(bd), "r"(idesc)
      : "memory");
}

struct Args {
  int N, reps;
  long long* out;   // [3]: issue cycles, total cycles, ummas
};

template <int MODE, int CH>
__global__ void __launch_bounds__(256, 1) probe_kernel(const Args p) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  const int rows = 1024;
  uint8_t* A = tc_smem;
  uint8_t* B = A + 2 * rows * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(B + 2 * 512 * 16);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < (2 * rows * 16 + 2 * 512 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(A)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(smem_u32(bars), 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 7) tmem_alloc(smem_u32(slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint64_t a_t = make_kmajor_desc(smem_u32(A), rows * 16, 128);
    const uint64_t b_t = make_kmajor_desc(smem_u32(B), 512 * 16, 128);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | (8u << 24);
    long long t0 = 0, t1 = 0;
    const int reps = p.reps;   // multiple of 8
    if (MODE == 0) {
      if (lane == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
          for (int c = 0; c < CH; ++c) umma_f16(tmem + c * p.N, a_t + (uint64_t)((r & 7) * 4), b_t, idesc, 1u);
        t1 = clock64();
      }
    } else if (MODE == 1) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int c = 0; c < CH; ++c) umma_elect(tmem + c * p.N, a_t + (uint64_t)((r & 7) * 4), b_t, idesc, 1u);
      t1 = clock64();
    } else if (MODE == 2) {
      t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int c = 0; c < CH; ++c) umma_elect(tmem + c * p.N, a_t + (uint64_t)(u * 4), b_t, idesc, 1u);
      }
      t1 = clock64();
    } else if (MODE == 3) {
      if (lane == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) umma_f16(tmem + c * p.N, a_t + (uint64_t)(u * 4), b_t, idesc, 1u);
        }
        t1 = clock64();
      }
    } else {
      t0 = clock64();
      for (int r = 0; r < reps; r += 8)
#pragma unroll
        for (int c = 0; c < CH; ++c) umma_elect8(tmem + c * p.N, a_t, 4ull, b_t, idesc);
      t1 = clock64();
    }
    __syncwarp();
    if (lane == 0) {
      umma_commit(smem_u32(bars));
      mbar_wait(smem_u32(bars), 0, 900);
      const long long t2 = clock64();
      p.out[0] = t1 - t0; p.out[1] = t2 - t0; p.out[2] = (long long)reps * CH;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) tmem_dealloc(tmem, 512);
}

template <int MODE, int CH>
void run(int N, long long* dout) {
  const size_t smem = 2 * 1024 * 16 + 2 * 512 * 16 + 128;
  cudaFuncSetAttribute(probe_kernel<MODE, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  Args p{N, 1024, dout};
  for (int rep = 0; rep < 2; ++rep) probe_kernel<MODE, CH><<<1, 256, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[3];
  cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost);
  printf("mode %d chains %d N %3d | issue %7.1f total %7.1f clk/umma (math %5.1f)\n", MODE, CH, N, (double)h[0] / h[2],
         (double)h[1] / h[2], N / 2.0);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  const int Ns[] = {32, 64, 128, 256};
  for (int N : Ns) {
    run<0, 1>(N, dout); run<0, 2>(N, dout);
    run<1, 1>(N, dout); run<1, 2>(N, dout);
    run<2, 1>(N, dout); run<2, 2>(N, dout);
    run<3, 1>(N, dout); run<3, 2>(N, dout);
    run<4, 1>(N, dout); run<4, 2>(N, dout);
  }
  return 0;
}
