// TMA probe (sm_100a): validates the tensor-map encodings of fv_tma.cuh on hardware and measures the streaming rate of
// the split-format box shape ({2*128 x 8 B, 1 plane} = 2 KB per cp.async.bulk.tensor).
//   1. split format [planes][L][8 halfs]: tile rows [g0, g0+rows) of 3 planes with g0 < 0 and g0+rows > L -> zero fill
//   2. fp32 rows [R][L]: box {256, 4} at a negative column
//   3. throughput: 148 CTAs x persistent loop, double-buffered 40 KB tiles
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_probe.cu
#include <cuda_fp16.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fastvocoder_b200/csrc/fv_tc.cuh"
#include "../../fastvocoder_b200/csrc/fv_tma.cuh"

namespace fv {
std::atomic<long long> g_launches{0};
std::atomic<long long> g_tc_launches{0};
}
using namespace fv;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void split_tile_kernel(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_tail,
                                  int g0, int rows, int rows_alloc, int plane0, int nplanes, uint4* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)nplanes * rows_alloc * 16);
  const uint32_t b = smem_u32(bar);
  if (threadIdx.x == 0) { mbar_init(b, 1); fence_mbar_init(); }
  __syncthreads();
  const int nfull = rows / TMA_SPLIT_RB, tail = rows - nfull * TMA_SPLIT_RB;
  if (threadIdx.x == 0) mbar_expect_tx(b, (uint32_t)(nplanes * rows * 16));
  __syncwarp();
  const int nops = nplanes * (nfull + (tail ? 1 : 0));
  for (int i = threadIdx.x; i < nops; i += 32) {
    const int pl = i / (nfull + (tail ? 1 : 0)), rb = i - pl * (nfull + (tail ? 1 : 0));
    const uint32_t dst = smem_u32(smem + ((size_t)pl * rows_alloc + rb * TMA_SPLIT_RB) * 16);
    tma_load_2d(dst, rb < nfull ? &tm_main : &tm_tail, 2 * (g0 + rb * TMA_SPLIT_RB), plane0 + pl, b);
  }
  mbar_wait(b, 0, 1);
  __syncwarp();
  for (int i = threadIdx.x; i < nplanes * rows_alloc; i += 32) out[i] = reinterpret_cast<uint4*>(smem)[i];
}

__global__ void f32_tile_kernel(const __grid_constant__ CUtensorMap tm, int c0, int r0, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 256 * 4);
  const uint32_t b = smem_u32(bar);
  if (threadIdx.x == 0) { mbar_init(b, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(b, 4 * 256 * 4);
    tma_load_2d(smem_u32(smem), &tm, c0, r0, b);
  }
  mbar_wait(b, 0, 2);
  for (int i = threadIdx.x; i < 4 * 256; i += 32) out[i] = reinterpret_cast<float*>(smem)[i];
}

// persistent streaming: each CTA walks tiles of `nplanes` x `rows` (rows multiple of 128), two buffers in flight
__global__ void stream_kernel(const __grid_constant__ CUtensorMap tm, int L, int nplanes_total, int nplanes, int rows,
                              int tiles_per_group, int total_tiles, unsigned long long* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const size_t tile_bytes = (size_t)nplanes * rows * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * tile_bytes);
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) { mbar_init(b0, 1); mbar_init(b0 + 8, 1); fence_mbar_init(); }
  __syncthreads();
  const int nrb = rows / TMA_SPLIT_RB;
  auto issue = [&](int tile, int s) {
    const int grp = tile / tiles_per_group, tt = tile - grp * tiles_per_group;
    if (threadIdx.x == 0) mbar_expect_tx(b0 + 8 * s, (uint32_t)tile_bytes);
    __syncwarp();
    for (int i = threadIdx.x; i < nplanes * nrb; i += 32) {
      const int pl = i / nrb, rb = i - pl * nrb;
      tma_load_2d(smem_u32(smem + s * tile_bytes + ((size_t)pl * rows + rb * TMA_SPLIT_RB) * 16), &tm,
                  2 * (tt * rows + rb * TMA_SPLIT_RB), grp * nplanes + pl, b0 + 8 * s);
    }
  };
  unsigned long long acc = 0;
  int it = 0;
  if ((int)blockIdx.x < total_tiles) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    if (tile + (int)gridDim.x < total_tiles) issue(tile + gridDim.x, s ^ 1);
    mbar_wait(b0 + 8 * s, (uint32_t)((it >> 1) & 1), 3);
    acc += reinterpret_cast<unsigned long long*>(smem + s * tile_bytes)[threadIdx.x];
    __syncwarp();
  }
  if (acc == 0x1234567ull) *sink = acc;
}

int main() {
  if (!tma_encode_fn()) { printf("FAIL: cuTensorMapEncodeTiled entry point not found\n"); return 1; }
  int fails = 0;
  {  // ---- 1. split format
    const int L = 1000, planes = 12, rows = 308, rows_alloc = 312, nplanes = 3, plane0 = 5;
    std::vector<__half> h((size_t)planes * L * 8);
    for (int p = 0; p < planes; ++p)
      for (int t = 0; t < L; ++t)
        for (int e = 0; e < 8; ++e) h[((size_t)p * L + t) * 8 + e] = __float2half((float)((p * 131 + t * 7 + e) % 2039));
    __half* d;
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tm_main, tm_tail;
    const int tail = rows % TMA_SPLIT_RB;
    if (!tma_encode_split(&tm_main, d, L, planes, TMA_SPLIT_RB) || !tma_encode_split(&tm_tail, d, L, planes, tail)) {
      printf("FAIL: split tensor-map encode\n");
      return 1;
    }
    uint4* out;
    CK(cudaMalloc(&out, (size_t)nplanes * rows_alloc * 16));
    for (int g0 : {-5, 0, L - 100, L - rows + 3, -400}) {
      CK(cudaMemset(out, 0xff, (size_t)nplanes * rows_alloc * 16));
      const size_t sm = (size_t)nplanes * rows_alloc * 16 + 64;
      CK(cudaFuncSetAttribute(split_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      split_tile_kernel<<<1, 32, sm>>>(tm_main, tm_tail, g0, rows, rows_alloc, plane0, nplanes, out);
      CK(cudaDeviceSynchronize());
      std::vector<__half> o((size_t)nplanes * rows_alloc * 8);
      CK(cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int pl = 0; pl < nplanes; ++pl)
        for (int r = 0; r < rows; ++r)
          for (int e = 0; e < 8; ++e) {
            const int t = g0 + r;
            const float want = (t >= 0 && t < L) ? (float)(((plane0 + pl) * 131 + t * 7 + e) % 2039) : 0.f;
            const float got = __half2float(o[((size_t)pl * rows_alloc + r) * 8 + e]);
            if (want != got && bad++ < 5) printf("  split mismatch g0=%d pl=%d r=%d e=%d want %g got %g\n", g0, pl, r, e, want, got);
          }
      printf("split tile g0=%d: %s (%d mismatches)\n", g0, bad ? "FAIL" : "ok", bad);
      fails += bad != 0;
    }
    cudaFree(d); cudaFree(out);
  }
  {  // ---- 2. fp32 rows
    const int L = 1000, R = 16;
    std::vector<float> h((size_t)R * L);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *out;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, 4 * 256 * 4));
    CUtensorMap tm;
    if (!tma_encode_f32_rows(&tm, d, L, R, 256, 4)) { printf("FAIL: f32 tensor-map encode\n"); return 1; }
    for (int c0 : {-8, 0, L - 60}) {
      f32_tile_kernel<<<1, 32, 4 * 256 * 4 + 64>>>(tm, c0, 8, out);
      CK(cudaDeviceSynchronize());
      std::vector<float> o(4 * 256);
      CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 256; ++c) {
          const int t = c0 + c;
          const float want = (t >= 0 && t < L) ? h[(size_t)(8 + r) * L + t] : 0.f;
          if (o[r * 256 + c] != want && bad++ < 5) printf("  f32 mismatch c0=%d r=%d c=%d want %g got %g\n", c0, r, c, want, o[r * 256 + c]);
        }
      printf("f32 tile c0=%d: %s (%d mismatches)\n", c0, bad ? "FAIL" : "ok", bad);
      fails += bad != 0;
    }
    cudaFree(d); cudaFree(out);
  }
  {  // ---- 3. streaming rate: 2 GB of split data (> L2), tiles of 4 planes x 640 rows = 40 KB
    const int nplanes = 4, rows = 640, tiles_per_group = 64;
    const int L = rows * tiles_per_group;                 // 40960 rows per plane
    const int groups = 192;                               // 192 * 4 planes * 40960 * 16 B = 503 MB
    const long long planes = (long long)groups * nplanes;
    uint8_t* d;
    CK(cudaMalloc(&d, (size_t)planes * L * 16));
    CK(cudaMemset(d, 1, (size_t)planes * L * 16));
    CUtensorMap tm;
    if (!tma_encode_split(&tm, d, L, planes, TMA_SPLIT_RB)) { printf("FAIL: stream tensor-map encode\n"); return 1; }
    unsigned long long* sink;
    CK(cudaMalloc(&sink, 8));
    const int total_tiles = groups * tiles_per_group;
    const size_t sm = 2 * (size_t)nplanes * rows * 16 + 64;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      stream_kernel<<<148, 32, sm>>>(tm, L, (int)planes, nplanes, rows, tiles_per_group, total_tiles, sink);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("stream: %.1f MB in %.3f ms = %.0f GB/s (148 CTAs x 1 warp, 2 x 40 KB in flight per CTA)\n",
             (double)planes * L * 16 / 1e6, ms, (double)planes * L * 16 / ms / 1e6);
    }
    cudaFree(d);
  }
  printf(fails ? "TMA PROBE FAIL\n" : "TMA PROBE PASS\n");
  return fails ? 1 : 0;
}
