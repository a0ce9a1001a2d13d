// UMMA issue/throughput probe (sm_100a): how many cycles does one tcgen05.mma kind::f16 (M=128, K=16, SWIZZLE_NONE
// K-major operands in shared memory) cost as a function of N, of how many independent accumulators ("chains") one
// issuing thread cycles through, of how many warps issue concurrently, and of an unaligned A start row (the tap shift
// of the implicit-GEMM conv)?   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
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

struct ProbeArgs {
  int N, chains, issuers, shift_rows, reps, vary_a, pair;   // pair: alternate N=2N / N=N UMMAs on the same accumulator
  long long* out;   // [grid][4 issuers][3]: issue cycles, total cycles, ummas
};

__global__ void __launch_bounds__(256, 1) probe_kernel(const ProbeArgs p) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  const int rows = 1024;
  uint8_t* A = tc_smem;                       // [2][rows][16 B]
  uint8_t* B = A + 2 * rows * 16;             // [2][512][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(B + 2 * 512 * 16);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (2 * rows * 16 + 2 * 512 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(A)[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(bars + i), 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 7) tmem_alloc(smem_u32(slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp < p.issuers && lane == 0) {
    const uint64_t a_t = make_kmajor_desc(smem_u32(A) + p.shift_rows * 16, rows * 16, 128);
    const uint64_t b_t = make_kmajor_desc(smem_u32(B), 512 * 16, 128);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | (8u << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * p.N) >> 3) << 17) | (8u << 24);
    const int ncol = p.pair ? 2 * p.N : p.N;
    const long long t0 = clock64();
    long long n = 0;
    for (int r = 0; r < p.reps; ++r) {
      for (int c = 0; c < p.chains; ++c) {
        const uint32_t d = tmem + (uint32_t)((warp * p.chains + c) * ncol);
        const uint64_t ad = a_t + (uint64_t)(p.vary_a ? ((r * 7 + c * 128) & 511) : c * 128);
        if (p.pair) {
          umma_f16(d, ad, b_t, idesc2, r > 0);
          umma_f16(d, ad + 4, b_t, idesc, 1u);
          n += 2;
        } else {
          umma_f16(d, ad, b_t, idesc, r > 0);
          n += 1;
        }
      }
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(bars + warp));
    mbar_wait(smem_u32(bars + warp), 0, 900 + warp);
    const long long t2 = clock64();
    long long* o = p.out + ((long long)blockIdx.x * 4 + warp) * 3;
    o[0] = t1 - t0; o[1] = t2 - t0; o[2] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) tmem_dealloc(tmem, 512);
}

int main() {
  long long* dout;
  const int max_grid = 148;
  cudaMalloc(&dout, max_grid * 4 * 3 * sizeof(long long));
  const size_t smem = 2 * 1024 * 16 + 2 * 512 * 16 + 128;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("%4s %6s %7s %5s %6s %4s %4s | %10s %10s\n", "N", "chains", "issuers", "shift", "vary_a", "pair", "grid", "issue/umma", "total/umma");
  auto run = [&](int N, int chains, int issuers, int shift, int vary, int pair, int grid) {
    const int ncol = pair ? 2 * N : N;
    if (issuers * chains * ncol > 512 || (pair && 2 * N > 256)) return;
    ProbeArgs p{N, chains, issuers, shift, 512, vary, pair, dout};
    cudaMemset(dout, 0, max_grid * 4 * 3 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) probe_kernel<<<grid, 256, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(max_grid * 4 * 3);
    cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double is = 0, tot = 0; int cnt = 0;
    for (int b = 0; b < grid; ++b)
      for (int w = 0; w < issuers; ++w) {
        const long long* o = &h[(b * 4 + w) * 3];
        is += (double)o[0] / o[2]; tot += (double)o[1] / o[2]; ++cnt;
      }
    // per-SM cost of one UMMA = per-issuer cycles / issuers
    printf("%4d %6d %7d %5d %6d %4d %4d | %10.1f %10.1f   (per SM: %.1f clk/umma, ideal %.1f)\n", N, chains, issuers, shift, vary,
           pair, grid, is / cnt, tot / cnt, tot / cnt / issuers, pair ? 0.75 * N : N / 2.0);
  };
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int N : Ns) {
    run(N, 1, 1, 0, 0, 0, 1);
    run(N, 2, 1, 0, 0, 0, 1);
    run(N, 4, 1, 0, 0, 0, 1);
    run(N, 1, 2, 0, 0, 0, 1);
    run(N, 1, 4, 0, 0, 0, 1);
    run(N, 2, 2, 0, 0, 0, 1);
    run(N, 1, 1, 3, 0, 0, 1);
    run(N, 1, 1, 0, 1, 0, 1);
    run(N, 4, 1, 0, 1, 0, 1);
    run(N, 1, 1, 0, 0, 1, 1);
    run(N, 2, 1, 0, 0, 1, 1);
    run(N, 1, 2, 0, 1, 1, 1);
    run(N, 1, 4, 0, 1, 1, 1);
    run(N, 1, 1, 0, 0, 0, 148);
    run(N, 1, 4, 0, 1, 1, 148);
  }
  return 0;
}
