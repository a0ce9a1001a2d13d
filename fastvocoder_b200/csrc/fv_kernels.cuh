// CUDA-core (exact fp32 FFMA) kernels of the generator forward path + small utility kernels.
// One generic "GEMM conv over time" kernel serves Conv1d, polyphase ConvTranspose1d and the
// Basis-MelGAN linear+overlap-add (see Layer in fv_model.h for the three views).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>

namespace fv {

extern std::atomic<long long> g_launches;

enum { PAD_ZERO = 0, PAD_REFLECT = 1 };
// ACC_ADD / ACC_ADD_DIV: read-modify-write of the running MRF sum exactly as hifigan.py:98-103 (xs += ...; xs / num_kernels).
// ACC_STORE_SCALE / ACC_RED_SCALE: the same sum with the 1/num_kernels folded into every branch — the first branch stores
// v/div, the others ADD v/div with a fire-and-forget reduction (red.global.add.f32, no read of the running sum in the
// epilogue); every element is touched by exactly one thread per launch, so the result is deterministic and differs from
// the divide-at-the-end form by rounding only (~1e-7 relative).
enum { ACC_STORE = 0, ACC_ADD = 1, ACC_ADD_DIV = 2, ACC_STORE_SCALE = 3, ACC_RED_SCALE = 4 };
__host__ __device__ inline bool acc_reads_y(int m) { return m == ACC_ADD || m == ACC_ADD_DIV; }
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// OUT_PHASE_SPLIT (tensor-core kernel only): the polyphase ConvTranspose output, LeakyReLU(out_slope)-activated and split
// into fp16 hi / lo in the blocked "split" activation format the fused ResBlock units fetch by TMA (fv_tma.cuh).
// OUT_BCL_SPLIT: the same for a stride-1 conv output [B, N, L] (h of an unfused ResBlock unit / the unit output that the
// next unit consumes).
enum { OUT_BCL = 0, OUT_BLC = 1, OUT_PHASE = 2, OUT_PHASE_SPLIT = 3, OUT_BCL_SPLIT = 4 };

struct ConvArgs {
  const float* x;     // [B, Cin, Lin]
  const float* w;     // derived image [Cin][K][N]
  const float* bias;  // [bias_mod] or nullptr
  const float* res;   // residual, same layout/strides as y, or nullptr
  float* y;
  int B, Cin, N, Lin, Lpos, K, dil, pad_left;
  int pad_mode;     // PAD_ZERO / PAD_REFLECT
  float pre_slope;  // < 0: no pre-activation; >= 0: leaky_relu(x, slope) (0 = ReLU)
  int acc_mode;     // ACC_*
  float acc_div;
  int post_tanh;
  int out_layout;   // OUT_*
  int bias_mod;
  // OUT_PHASE (ConvTranspose1d): n = r*ph_cout + co ; t = pos*ph_stride + r - ph_pad in [0, ph_lout)
  int ph_stride, ph_pad, ph_cout, ph_lout;
  float out_slope;               // OUT_*_SPLIT: LeakyReLU slope baked into the split copy (the consumers' pre-activation)
  // tensor-core kernel only: x / res are split-format buffers (fv_tma.cuh) instead of fp32 [B, C, L].  x_split: the A
  // tiles are fetched by TMA (zero padding only; the copy is already activated with pre_slope); res_split: the residual
  // is rebuilt from the split copy, whose baked slope is 1 / res_inv_slope.
  int x_split, res_split;
  float res_inv_slope;
  // two-input (pair) layers on the tensor-core kernel: the FIRST input (channels < cin_split: h of the ResidualStack, written
  // pre-activated in the split format by the dilated conv) is fetched by TMA, the second (the raw stack input) by the loaders
  int x1_split;
  // OUT_BCL_SPLIT with acc_mode ACC_STORE_SCALE: the LAST branch of an MRF stage emits the stage result itself,
  // split(lrelu(ysum + v / acc_div)), where ysum is the fp32 running sum the other branches stored / red-added.
  const float* ysum;
  long long x_bs, y_bs, res_bs;  // batch strides in floats
  // two-input form (L_PAIR): input channels >= cin_split come from x2 (own pre-activation); 0 = single input
  const float* x2;
  int cin_split;
  float pre_slope2;
  long long x2_bs;
  // ragged batches: valid input length of utterance b (<= Lin; Lin stays the row stride). nullptr = all Lin.
  // Samples at or beyond lens[b] are treated exactly like the padding beyond the end of the sequence.
  const int* lens;
};

// LeakyReLU with 0 <= slope <= 1 (slope 0 = ReLU) is max(v, slope*v); slope < 0 means "no activation".
__device__ __forceinline__ float pre_act(float v, float slope) {
  return slope >= 0.f ? fmaxf(v, v * slope) : v;
}

// ---------------------------------------------------------------------------------------------
// Generic fp32 conv: CTA tile = CO_T outputs x TT positions, 8 warps, thread tile 8 (n) x 8 (pos).
// lanes own consecutive positions (conflict-free shared reads, coalesced stores), warps own groups
// of 8 output channels (weight reads are warp-broadcast LDS.128).
// ---------------------------------------------------------------------------------------------
template <int CO_T>
struct ConvTile {
  static constexpr int CPT = 8, TPT = 8, CI_C = 8;
  static constexpr int NWC = CO_T / CPT;       // warps along N
  static constexpr int NWT = 8 / NWC;          // warps along positions
  static constexpr int TT = NWT * 32 * TPT;    // positions per CTA
};

template <int CO_T>
__global__ void __launch_bounds__(256) conv_ffma_kernel(const ConvArgs a) {
  using T = ConvTile<CO_T>;
  constexpr int CPT = T::CPT, TPT = T::TPT, CI_C = T::CI_C, NWC = T::NWC, TT = T::TT;
  extern __shared__ float smem[];
  const int halo = (a.K - 1) * a.dil;
  const int XW = TT + halo;
  float* sx = smem;                 // [CI_C][XW]
  float* sw = smem + CI_C * XW;     // [CI_C][K][CO_T]
  const int b = blockIdx.z;
  const int n0 = blockIdx.y * CO_T;
  const int t0 = blockIdx.x * TT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wc = warp % NWC, wt = warp / NWC;

  float acc[CPT][TPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i)
#pragma unroll
    for (int j = 0; j < TPT; ++j) acc[i][j] = 0.f;

  const float* xb = a.x + (long long)b * a.x_bs;
  const int wrow = a.K * CO_T;
  const int Lb = a.lens ? __ldg(a.lens + b) : a.Lin;

  for (int ci0 = 0; ci0 < a.Cin; ci0 += CI_C) {
    {  // x tile: warp `warp` stages channel ci0+warp (CI_C == 8 warps)
      const int ci = ci0 + warp;
      float* row = sx + warp * XW;
      const bool cok = ci < a.Cin;
      const bool second = a.cin_split > 0 && ci >= a.cin_split;
      const float* xr = second ? a.x2 + (long long)b * a.x2_bs + (long long)(ci - a.cin_split) * a.Lin
                               : xb + (long long)ci * a.Lin;
      const float slope = second ? a.pre_slope2 : a.pre_slope;
      for (int s = lane; s < XW; s += 32) {
        int g = t0 - a.pad_left + s;
        if (a.pad_mode == PAD_REFLECT) {
          if (g < 0) g = -g;
          if (g >= Lb) g = 2 * (Lb - 1) - g;
        }
        float v = 0.f;
        if (cok && g >= 0 && g < Lb) v = pre_act(__ldg(xr + g), slope);
        row[s] = v;
      }
    }
    {  // weight tile rows (c, j) are contiguous in the derived image
      const int rows = CI_C * a.K;
      const long long base = (long long)ci0 * a.K;
      const int rows_ok = (a.Cin - ci0) * a.K;  // rows beyond Cin are zero
      for (int idx = threadIdx.x; idx < rows * CO_T; idx += 256) {
        const int r = idx / CO_T, co = idx - r * CO_T;
        float v = 0.f;
        if (r < rows_ok && n0 + co < a.N) v = __ldg(a.w + (base + r) * a.N + n0 + co);
        sw[idx] = v;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CI_C; ++c) {
      const float* sxr = sx + c * XW + wt * (32 * TPT) + lane;
      const float* swr = sw + c * wrow + wc * CPT;
#pragma unroll 1
      for (int j = 0; j < a.K; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(swr + j * CO_T);
        const float4 w1 = *reinterpret_cast<const float4*>(swr + j * CO_T + 4);
        const float wv[CPT] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        float xv[TPT];
#pragma unroll
        for (int i = 0; i < TPT; ++i) xv[i] = sxr[j * a.dil + 32 * i];
#pragma unroll
        for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
          for (int i = 0; i < TPT; ++i) acc[cc][i] = fmaf(wv[cc], xv[i], acc[cc][i]);
      }
    }
    __syncthreads();
  }

  // epilogue: bias, residual, MRF accumulate, tanh, layout
  float* yb = a.y + (long long)b * a.y_bs;
  const float* rb = a.res ? a.res + (long long)b * a.res_bs : nullptr;
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc) {
    const int n = n0 + wc * CPT + cc;
    if (n >= a.N) break;
    const float bv = a.bias ? __ldg(a.bias + (n % a.bias_mod)) : 0.f;
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      const int pos = t0 + wt * (32 * TPT) + lane + 32 * i;
      if (pos >= a.Lpos) continue;
      long long o;
      if (a.out_layout == OUT_BCL) {
        o = (long long)n * a.Lpos + pos;
      } else if (a.out_layout == OUT_BLC) {
        o = (long long)pos * a.N + n;
      } else {
        const int r = n / a.ph_cout, co = n - r * a.ph_cout;
        const int t = pos * a.ph_stride + r - a.ph_pad;
        if (t < 0 || t >= a.ph_lout) continue;
        o = (long long)co * a.ph_lout + t;
      }
      float v = acc[cc][i] + bv;
      if (rb) v += rb[o];
      if (a.acc_mode == ACC_ADD) v = yb[o] + v;
      else if (a.acc_mode == ACC_ADD_DIV) v = (yb[o] + v) / a.acc_div;
      else if (a.acc_mode == ACC_STORE_SCALE) v = v * (1.0f / a.acc_div);
      else if (a.acc_mode == ACC_RED_SCALE) v = yb[o] + v * (1.0f / a.acc_div);
      if (a.post_tanh) v = tanhf(v);
      yb[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Narrow-output conv (conv_post 16->1 / 64->4, MelGAN LastLayer 32->1): one thread per output position,
// all (<= 4) output channels in registers, weights in shared memory, input through L1 (each element is
// reused by K neighbouring threads).  Pure streaming kernel: 4*Cin bytes in, 4*Cout bytes out per position.
// ---------------------------------------------------------------------------------------------
template <int NOUT>
__global__ void __launch_bounds__(256) conv_narrow_kernel(const ConvArgs a) {
  extern __shared__ float smem[];   // [Cin][K][NOUT]
  const int wn = a.Cin * a.K * NOUT;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) {
    const int n = i % NOUT, r = i / NOUT;   // derived image is [Cin][K][N]
    smem[i] = n < a.N ? __ldg(a.w + (long long)r * a.N + n) : 0.f;
  }
  __syncthreads();
  const int b = blockIdx.y;
  const float* xb = a.x + (long long)b * a.x_bs;
  float* yb = a.y + (long long)b * a.y_bs;
  const int Lb = a.lens ? __ldg(a.lens + b) : a.Lin;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.Lpos; t += gridDim.x * blockDim.x) {
    float acc[NOUT];
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[n] = 0.f;
    const int g0 = t - a.pad_left;
    const bool interior = g0 >= 0 && g0 + (a.K - 1) * a.dil < Lb;   // no padding involved: skip the checks
    for (int ci = 0; ci < a.Cin; ++ci) {
      const float* xr = xb + (long long)ci * a.Lin;
      const float* wr = smem + ci * a.K * NOUT;
      if (interior) {
        for (int j = 0; j < a.K; ++j) {
          const float xv = pre_act(__ldg(xr + g0 + j * a.dil), a.pre_slope);
#pragma unroll
          for (int n = 0; n < NOUT; ++n) acc[n] = fmaf(wr[j * NOUT + n], xv, acc[n]);
        }
      } else {
        for (int j = 0; j < a.K; ++j) {
          int g = g0 + j * a.dil;
          if (a.pad_mode == PAD_REFLECT) {
            if (g < 0) g = -g;
            if (g >= Lb) g = 2 * (Lb - 1) - g;
          }
          const float xv = (g >= 0 && g < Lb) ? pre_act(__ldg(xr + g), a.pre_slope) : 0.f;
#pragma unroll
          for (int n = 0; n < NOUT; ++n) acc[n] = fmaf(wr[j * NOUT + n], xv, acc[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
      if (n >= a.N) break;
      float v = acc[n] + (a.bias ? __ldg(a.bias + n) : 0.f);
      if (a.post_tanh) v = tanhf(v);
      yb[(long long)n * a.Lpos + t] = v;
    }
  }
}

// Vectorised variant (dilation 1, zero padding, K <= 8): each thread produces 4 consecutive positions from three
// aligned float4 loads per input channel (12-sample window) instead of 4*K scalar loads.
template <int NOUT>
__global__ void __launch_bounds__(256) conv_narrow_v4_kernel(const ConvArgs a) {
  extern __shared__ float smem[];   // [Cin][K][NOUT]
  const int wn = a.Cin * a.K * NOUT;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) {
    const int n = i % NOUT, r = i / NOUT;
    smem[i] = n < a.N ? __ldg(a.w + (long long)r * a.N + n) : 0.f;
  }
  __syncthreads();
  const int b = blockIdx.y;
  const float* xb = a.x + (long long)b * a.x_bs;
  float* yb = a.y + (long long)b * a.y_bs;
  const int off = 4 - a.pad_left;   // window index of tap 0 for output 0
  const int nquads = a.Lpos >> 2;
  for (int qd = blockIdx.x * blockDim.x + threadIdx.x; qd < nquads; qd += gridDim.x * blockDim.x) {
    const int t0 = qd << 2;
    float acc[NOUT][4];
#pragma unroll
    for (int n = 0; n < NOUT; ++n)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[n][q] = 0.f;
    const bool okL = t0 - 4 >= 0, okR = t0 + 4 < a.Lin;
    for (int ci = 0; ci < a.Cin; ++ci) {
      const float4* xr = reinterpret_cast<const float4*>(xb + (long long)ci * a.Lin + t0);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v0 = okL ? __ldg(xr - 1) : z, v1 = __ldg(xr), v2 = okR ? __ldg(xr + 1) : z;
      float w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
      for (int i = 0; i < 12; ++i) w[i] = pre_act(w[i], a.pre_slope);
      const float* wr = smem + ci * a.K * NOUT;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j >= a.K) break;
#pragma unroll
        for (int n = 0; n < NOUT; ++n) {
          const float wt = wr[j * NOUT + n];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[n][q] = fmaf(wt, w[q + j + off], acc[n][q]);   // off is 0..4: see launcher
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
      if (n >= a.N) break;
      const float bv = a.bias ? __ldg(a.bias + n) : 0.f;
      float4 o;
      o.x = acc[n][0] + bv; o.y = acc[n][1] + bv; o.z = acc[n][2] + bv; o.w = acc[n][3] + bv;
      if (a.post_tanh) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
      *reinterpret_cast<float4*>(yb + (long long)n * a.Lpos + t0) = o;
    }
  }
}

inline bool conv_narrow_v4_ok(const ConvArgs& a) {
  // window index q + j + off must stay in [0, 12): off = 4 - pad_left in [0,4], 3 + (K-1) + off <= 11
  return a.lens == nullptr && a.dil == 1 && a.pad_mode == PAD_ZERO && a.K <= 8 && a.pad_left <= 4 && (a.K - 1) + (4 - a.pad_left) <= 8 &&
         a.Lin % 4 == 0 && a.Lpos == a.Lin && a.x_bs % 4 == 0 && a.y_bs % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0;
}

// Streaming kernel for the k = 7 narrow-output convs every config ends with (conv_post 16->1 / 64->4, MelGAN LastLayer
// 32->1; padding 3, dilation 1): 4*Cin bytes in and 4*Cout bytes out per position, so HBM is the roofline.  Each thread
// produces 4 consecutive positions from three aligned 16-byte loads per input channel (window t0-4 .. t0+7, all indices
// compile-time -> registers), the 7*NOUT weights of the channel come from shared memory as broadcast 16-byte reads, and
// tiles that touch the padding / the ragged end take the scalar path of conv_narrow_kernel's arithmetic.
// Q = quads (4 outputs) per thread: Q = 2 loads 4 aligned float4 per channel for 8 outputs (2x instead of 3x re-reads through
// L1, 14 instead of 2 x 10 activations) — same per-output FMA order, so Q = 1 and Q = 2 are bit-identical.
template <int NOUT, int Q>
__global__ void __launch_bounds__(256) conv_narrow7_kernel(const ConvArgs a) {
  constexpr int K = 7, PADL = 3, T = 4 * Q;
  extern __shared__ __align__(16) float smem[];   // [Cin][NOUT][8] (tap 7 = 0)
  for (int i = threadIdx.x; i < a.Cin * NOUT * 8; i += blockDim.x) {
    const int j = i & 7, n = (i >> 3) % NOUT, ci = i / (8 * NOUT);
    smem[i] = (j < K && n < a.N) ? __ldg(a.w + ((long long)ci * K + j) * a.N + n) : 0.f;   // derived image is [Cin][K][N]
  }
  __syncthreads();
  const int b = blockIdx.y;
  const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
  float* __restrict__ yb = a.y + (long long)b * a.y_bs;
  const int Lb = a.lens ? __ldg(a.lens + b) : a.Lin;
  const int ngroups = (a.Lpos + T - 1) / T;
  const float slope = a.pre_slope;
  const uint32_t uL = (uint32_t)a.Lin;
  for (int qd = blockIdx.x * blockDim.x + threadIdx.x; qd < ngroups; qd += gridDim.x * blockDim.x) {
    const int t0 = qd * T;
    float acc[NOUT][T];
#pragma unroll
    for (int n = 0; n < NOUT; ++n)
#pragma unroll
      for (int q = 0; q < T; ++q) acc[n][q] = 0.f;
    if (t0 - 4 >= 0 && t0 + T + 4 <= Lb && t0 + T <= a.Lpos) {   // interior: the whole (T + 8)-sample window is real data
      const float* px = xb + t0;
#pragma unroll 2
      for (int ci = 0; ci < a.Cin; ++ci) {
        const float4* xr = reinterpret_cast<const float4*>(px + (uint32_t)ci * uL);
        float w[T + 8];
#pragma unroll
        for (int v = 0; v < Q + 2; ++v) {
          const float4 f = __ldg(xr + (v - 1));
          w[4 * v] = f.x; w[4 * v + 1] = f.y; w[4 * v + 2] = f.z; w[4 * v + 3] = f.w;
        }
#pragma unroll
        for (int i = 1; i < T + 7; ++i) w[i] = pre_act(w[i], slope);   // w[0], w[T+7] are never used (window t0-3 .. t0+T+2)
#pragma unroll
        for (int n = 0; n < NOUT; ++n) {
          const float4 c0 = *reinterpret_cast<const float4*>(smem + (ci * NOUT + n) * 8);
          const float4 c1 = *reinterpret_cast<const float4*>(smem + (ci * NOUT + n) * 8 + 4);
          const float cw[7] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z};
#pragma unroll
          for (int j = 0; j < K; ++j)
#pragma unroll
            for (int q = 0; q < T; ++q) acc[n][q] = fmaf(cw[j], w[q + j + (4 - PADL)], acc[n][q]);
        }
      }
    } else {   // padding / ragged end / tail: scalar taps with the generic index rules
      for (int ci = 0; ci < a.Cin; ++ci) {
        const float* xr = xb + (long long)ci * a.Lin;
        for (int j = 0; j < K; ++j) {
#pragma unroll
          for (int q = 0; q < T; ++q) {
            int g = t0 + q - PADL + j;
            if (a.pad_mode == PAD_REFLECT) {
              if (g < 0) g = -g;
              if (g >= Lb) g = 2 * (Lb - 1) - g;
            }
            const float xv = (g >= 0 && g < Lb) ? pre_act(__ldg(xr + g), slope) : 0.f;
#pragma unroll
            for (int n = 0; n < NOUT; ++n) acc[n][q] = fmaf(smem[(ci * NOUT + n) * 8 + j], xv, acc[n][q]);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
      if (n >= a.N) break;
      const float bv = a.bias ? __ldg(a.bias + n) : 0.f;
      float o[T];
#pragma unroll
      for (int q = 0; q < T; ++q) {
        o[q] = acc[n][q] + bv;
        if (a.post_tanh) o[q] = tanhf(o[q]);
      }
      float* yo = yb + (long long)n * a.Lpos + t0;
#pragma unroll
      for (int v = 0; v < Q; ++v) {
        if (t0 + 4 * v + 4 <= a.Lpos) {
          *reinterpret_cast<float4*>(yo + 4 * v) = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        } else {
          for (int q = 4 * v; q < 4 * v + 4 && t0 + q < a.Lpos; ++q) yo[q] = o[q];
        }
      }
    }
  }
}
// shapes the streaming kernel takes (both the tensor-core and the exact-fp32 path route them here: an HBM-bound op gains
// nothing from a zero-padded N = 16 UMMA)
inline bool conv_narrow7_ok(const ConvArgs& a) {
  return a.cin_split == 0 && a.N <= 4 && a.K == 7 && a.dil == 1 && a.pad_left == 3 && a.out_layout == OUT_BCL &&
         a.res == nullptr && a.acc_mode == ACC_STORE && a.Lpos == a.Lin && a.Lin % 4 == 0 && a.x_bs % 4 == 0 && a.y_bs % 4 == 0 &&
         (long long)a.Cin * a.Lin < 0x7fffffffLL && (size_t)a.Cin * 4 * 8 * sizeof(float) <= 40 * 1024 &&
         (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0;
}
inline cudaError_t launch_conv_narrow7(const ConvArgs& a, cudaStream_t st) {
  // outputs per thread / 4 (single-channel output).  Measured (gpurun r2s): Q = 2 is SLOWER (conv_post 0.148 -> 0.160 ms, MelGAN
  // LastLayer 0.271 -> 0.320 ms: half the threads, fewer loads in flight) -> Q = 1 stays the default
  static const int q_env = getenv("FV_NARROW7_Q") ? atoi(getenv("FV_NARROW7_Q")) : 1;
  const int Qn = (a.N == 1 && q_env >= 2) ? 2 : 1;
  long long g4 = ((a.Lpos + 4 * Qn - 1) / (4 * Qn) + 255) / 256;
  if (g4 > 148 * 16) g4 = 148 * 16;
  if (g4 < 1) g4 = 1;
  dim3 grid((unsigned)g4, a.B);
  if (a.N == 1 && Qn == 2) conv_narrow7_kernel<1, 2><<<grid, 256, (size_t)a.Cin * 1 * 8 * sizeof(float), st>>>(a);
  else if (a.N == 1) conv_narrow7_kernel<1, 1><<<grid, 256, (size_t)a.Cin * 1 * 8 * sizeof(float), st>>>(a);
  else if (a.N == 2) conv_narrow7_kernel<2, 1><<<grid, 256, (size_t)a.Cin * 2 * 8 * sizeof(float), st>>>(a);
  else conv_narrow7_kernel<4, 1><<<grid, 256, (size_t)a.Cin * 4 * 8 * sizeof(float), st>>>(a);
  g_launches++;
  return cudaGetLastError();
}

inline bool conv_narrow_ok(const ConvArgs& a) {
  return a.cin_split == 0 && a.N <= 4 && a.K <= 16 && a.out_layout == OUT_BCL && a.res == nullptr && a.acc_mode == ACC_STORE &&
         (size_t)a.Cin * a.K * 4 * sizeof(float) <= 40 * 1024;
}

inline cudaError_t launch_conv_narrow(const ConvArgs& a, cudaStream_t st) {
  long long gx = (a.Lpos + 255) / 256;
  if (gx > 148 * 32) gx = 148 * 32;
  dim3 grid((unsigned)gx, a.B);
  if (conv_narrow_v4_ok(a)) {
    long long g4 = (a.Lpos / 4 + 255) / 256;
    if (g4 > 148 * 32) g4 = 148 * 32;
    if (g4 < 1) g4 = 1;
    dim3 grid4((unsigned)g4, a.B);
    if (a.N == 1) conv_narrow_v4_kernel<1><<<grid4, 256, (size_t)a.Cin * a.K * 1 * sizeof(float), st>>>(a);
    else conv_narrow_v4_kernel<4><<<grid4, 256, (size_t)a.Cin * a.K * 4 * sizeof(float), st>>>(a);
  } else if (a.N == 1) conv_narrow_kernel<1><<<grid, 256, (size_t)a.Cin * a.K * 1 * sizeof(float), st>>>(a);
  else conv_narrow_kernel<4><<<grid, 256, (size_t)a.Cin * a.K * 4 * sizeof(float), st>>>(a);
  g_launches++;
  return cudaGetLastError();
}

inline size_t conv_ffma_smem(int co_t, int K, int dil) {
  int tt = (co_t == 16) ? ConvTile<16>::TT : (co_t == 32) ? ConvTile<32>::TT : ConvTile<64>::TT;
  return (size_t)(8 * (tt + (K - 1) * dil) + 8 * K * co_t) * sizeof(float);
}

inline cudaError_t launch_conv_ffma(const ConvArgs& a, cudaStream_t st) {
  if (conv_narrow7_ok(a)) return launch_conv_narrow7(a, st);
  if (conv_narrow_ok(a)) return launch_conv_narrow(a, st);
  const int co_t = a.N <= 16 ? 16 : (a.N <= 32 ? 32 : 64);
  const size_t smem = conv_ffma_smem(co_t, a.K, a.dil);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e;
  dim3 block(256);
#define FV_LAUNCH(COT)                                                                              \
  {                                                                                                 \
    dim3 grid((a.Lpos + ConvTile<COT>::TT - 1) / ConvTile<COT>::TT, (a.N + COT - 1) / COT, a.B);    \
    if (smem > 48 * 1024) {                                                                         \
      e = cudaFuncSetAttribute(conv_ffma_kernel<COT>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                               (int)smem);                                                          \
      if (e != cudaSuccess) return e;                                                               \
    }                                                                                               \
    conv_ffma_kernel<COT><<<grid, block, smem, st>>>(a);                                            \
  }
  if (co_t == 16) FV_LAUNCH(16) else if (co_t == 32) FV_LAUNCH(32) else FV_LAUNCH(64)
#undef FV_LAUNCH
  g_launches++;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Weight image derivation (bind time)
// ---------------------------------------------------------------------------------------------
// Conv1d [Cout][Cin][K] -> [Cin][K][Cout]
__global__ void derive_conv_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cout, int Cin, int K) {
  const long long n = (long long)Cout * Cin * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long r = i / Cout;
    const int j = (int)(r % K);
    const int ci = (int)(r / K);
    wd[i] = w[((long long)co * Cin + ci) * K + j];
  }
}
// ConvTranspose1d [Cin][Cout][K] -> [Cin][M][stride*Cout], tap m' <-> input offset pos-(M-1)+m' <-> kk = r + (M-1-m')*stride
__global__ void derive_convt_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cin, int Cout, int K,
                                    int stride, int M) {
  const int N = stride * Cout;
  const long long n = (long long)Cin * M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int nn = (int)(i % N);
    const long long q = i / N;
    const int mp = (int)(q % M);
    const int ci = (int)(q / M);
    const int r = nn / Cout, co = nn - r * Cout;
    const int kk = r + (M - 1 - mp) * stride;
    wd[i] = kk < K ? w[((long long)ci * Cout + co) * K + kk] : 0.f;
  }
}
// UpsampleLayer conv [Cout][Cin][K] -> [Cin][Kd][u*Cout]: taps hitting the same input sample are summed (in double)
__global__ void derive_upconv_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cin, int Cout, int K, int u,
                                     int pad, int dmin, int Kd) {
  const int N = u * Cout;
  const long long n = (long long)Cin * Kd * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int nn = (int)(i % N);
    const long long q = i / N;
    const int jj = (int)(q % Kd);
    const int ci = (int)(q / Kd);
    const int r = nn / Cout, co = nn - r * Cout;
    double acc = 0.0;
    for (int j = 0; j < K; ++j) {
      const int v = r - pad + j;
      const int d = v >= 0 ? v / u : -((-v + u - 1) / u);   // floor division
      if (d == dmin + jj) acc += (double)w[((long long)co * Cin + ci) * K + j];
    }
    wd[i] = (float)acc;
  }
}
// Pair of 1x1 convs (ResidualStack stack.4 + skip_layer): [C][C][1] x2 -> [2C][1][C] and summed bias
__global__ void derive_pair_kernel(const float* __restrict__ wa, const float* __restrict__ wb, const float* __restrict__ ba,
                                   const float* __restrict__ bb, float* __restrict__ wd, float* __restrict__ bias, int C) {
  const long long n = 2LL * C * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % C);
    const int ci = (int)(i / C);
    wd[i] = ci < C ? wa[(long long)co * C + ci] : wb[(long long)co * C + (ci - C)];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C; i += gridDim.x * blockDim.x)
    bias[i] = (ba ? ba[i] : 0.f) + (bb ? bb[i] : 0.f);
}
// Basis Linear [L][C] -> [C][2][hop]: tap 0 <-> frame n-1 second half (rows hop..L-1), tap 1 <-> frame n first half
__global__ void derive_basis_kernel(const float* __restrict__ w, float* __restrict__ wd, int L, int C) {
  const int hop = L / 2;
  const long long n = (long long)C * 2 * hop;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % hop);
    const long long q = i / hop;
    const int j = (int)(q % 2);
    const int ci = (int)(q / 2);
    wd[i] = w[(long long)((j == 0 ? hop : 0) + co) * C + ci];
  }
}

// ---------------------------------------------------------------------------------------------
// PQMF (pqmf.py:108-135).  Index arithmetic is exact: for impulse inputs exactly one product is
// non-zero per output sample, so results are bit-identical to the reference.
// synthesis: y[t] = sum_k sum_{j : (t+j-P) % S == 0, 0 <= (t+j-P)/S < Lb} (S*x[k,(t+j-P)/S]) * h[k,j],  P = taps/2
// ---------------------------------------------------------------------------------------------
__global__ void pqmf_synthesis_kernel(const float* __restrict__ x, const float* __restrict__ h, float* __restrict__ y,
                                      int S, int taps, int Lb, const int* __restrict__ lens = nullptr) {
  extern __shared__ float sh[];  // [S][taps+1]
  const int nt = taps + 1;
  for (int i = threadIdx.x; i < S * nt; i += blockDim.x) sh[i] = h[i];
  __syncthreads();
  const int b = blockIdx.y;
  const long long L = (long long)Lb * S;
  const int P = taps / 2;
  const float scale = (float)S;
  const int Lv = lens ? __ldg(lens + b) : Lb;   // ragged batches: sub-band samples beyond the utterance are zero
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L; t += (long long)gridDim.x * blockDim.x) {
    // smallest j >= 0 with (t + j - P) % S == 0
    long long u = t - P;
    int j0 = (int)(((-u) % S + S) % S);
    float acc = 0.f;
    for (int k = 0; k < S; ++k) {
      const float* xk = x + ((long long)b * S + k) * Lb;
      const float* hk = sh + k * nt;
      for (int j = j0; j < nt; j += S) {
        const long long n = (u + j) / S;
        if (n >= 0 && n < Lv) acc = fmaf(scale * __ldg(xk + n), hk[j], acc);
      }
    }
    y[(long long)b * L + t] = acc;
  }
}
// Polyphase form of the same sum for the shipped filterbank shape (S sub-bands, NT = taps + 1 coefficients): output
// t = S*q + r reads x[k, q + m] * h[k, S*m - r + P] for the m with 0 <= S*m - r + P <= taps.  One thread owns one q:
// it reads the window x[k, q + MLO .. q + MHI] of each band once (registers) and produces the S outputs of that
// position group (one 16-byte store for S = 4) — no per-tap index division, 4 B in + 4 B out per output sample.
// Per output the products are added in the same order as above (k ascending, then j ascending), so both kernels
// return identical bits; out-of-range window samples contribute an exact 0 * h.
template <int S, int NT, int QT>
__global__ void __launch_bounds__(256) pqmf_synthesis_poly_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                                  float* __restrict__ y, int Lb,
                                                                  const int* __restrict__ lens) {
  constexpr int P = (NT - 1) / 2;
  constexpr int MLO = -(P / S);
  constexpr int MHI = (NT - 1 + S - 1 - P) / S;
  constexpr int W = MHI - MLO + 1;
  __shared__ float sh[S * NT];
  // (S*x) * h == x * (S*h) exactly (S is a power of two here), so the up-sampling gain is folded into the coefficients
  for (int i = threadIdx.x; i < S * NT; i += blockDim.x) sh[i] = (float)S * h[i];
  __syncthreads();
  const int b = blockIdx.y;
  // QT position groups per thread, blockDim apart (coalesced loads / 16-byte stores): every filter coefficient read
  // from shared memory feeds QT FMAs.  The kernel is FP32-bound, not HBM-bound: 63 FMAs per 8 bytes of traffic.
  const int q0 = blockIdx.x * (blockDim.x * QT) + threadIdx.x;
  const int Lv = lens ? __ldg(lens + b) : Lb;
  // every window of this thread inside [0, Lv): no per-load bounds predicates (they were half of the instruction stream)
  const bool interior = q0 + MLO >= 0 && q0 + (QT - 1) * (int)blockDim.x + MHI < Lv;
  float acc[QT][S];
#pragma unroll
  for (int u = 0; u < QT; ++u)
#pragma unroll
    for (int r = 0; r < S; ++r) acc[u][r] = 0.f;
#pragma unroll 1   // one band's windows in registers at a time (unrolled: 240 registers, one CTA per SM)
  for (int k = 0; k < S; ++k) {
    const float* xk = x + ((long long)b * S + k) * Lb;
    float xw[QT][W];
    if (interior) {
#pragma unroll
      for (int u = 0; u < QT; ++u) {
        const float* xq = xk + (q0 + u * (int)blockDim.x + MLO);
#pragma unroll
        for (int w = 0; w < W; ++w) xw[u][w] = __ldg(xq + w);
      }
    } else {
#pragma unroll
      for (int u = 0; u < QT; ++u) {
        const int q = q0 + u * (int)blockDim.x;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const int n = q + MLO + w;
          xw[u][w] = (n >= 0 && n < Lv) ? __ldg(xk + n) : 0.f;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < S; ++r) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const int j = S * (MLO + w) - r + P;   // compile-time after unrolling
        if (j >= 0 && j < NT) {
          const float c = sh[k * NT + j];
#pragma unroll
          for (int u = 0; u < QT; ++u) acc[u][r] = fmaf(xw[u][w], c, acc[u][r]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < QT; ++u) {
    const int q = q0 + u * (int)blockDim.x;
    if (q >= Lb) continue;
    float* yo = y + ((long long)b * Lb + q) * S;
    if (S == 4) {
      *reinterpret_cast<float4*>(yo) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
    } else {
#pragma unroll
      for (int r = 0; r < S; ++r) yo[r] = acc[u][r];
    }
  }
}
// 256-bit global store (sm_100: STG.E.ENL2.256): eight consecutive floats of one thread in one fully used 32-byte sector
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
// Same polyphase sum with the memory side vectorised: one thread owns G = 4 CONSECUTIVE position groups (16 output samples).
// Its window of a band is x[k, q0-7 .. q0+11], fetched as five aligned 16-byte loads (the kernel above issues 64 scalar loads
// per band for the same 4 groups and re-reads every sample 16x through L1), the four coefficients of a (band, window tap)
// pair — one per output phase r — are ONE 16-byte shared-memory read, and the 16 outputs leave as two 32-byte stores.
// Per output the products are still added k ascending, then j ascending: bit-identical to both kernels above.
// Requires Lb % 4 == 0 (16-byte aligned band rows, whole groups) and a 32-byte aligned y.
template <int S, int NT>
__global__ void __launch_bounds__(128) pqmf_synthesis_poly_v4_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                                     float* __restrict__ y, int Lb,
                                                                     const int* __restrict__ lens) {
  static_assert(S == 4 && NT == 63, "shipped PQMF(subbands=4, taps=62) only");
  constexpr int P = (NT - 1) / 2;                 // 31
  constexpr int MLO = -(P / S);                   // -7
  constexpr int MHI = (NT - 1 + S - 1 - P) / S;   // 8
  constexpr int W = MHI - MLO + 1;                // 16 window taps per output
  constexpr int G = 4;                            // position groups per thread
  __shared__ float4 sh[S * W];                    // [k][w] -> coefficients of phases r = 0..3 (0 where the tap does not exist)
  for (int i = threadIdx.x; i < S * W; i += blockDim.x) {
    const int k = i / W, w = i - k * W;
    float c[S];
#pragma unroll
    for (int r = 0; r < S; ++r) {
      const int j = S * (MLO + w) - r + P;
      c[r] = (j >= 0 && j < NT) ? (float)S * h[k * NT + j] : 0.f;   // (S*x)*h == x*(S*h) exactly: S is a power of two
    }
    sh[i] = make_float4(c[0], c[1], c[2], c[3]);
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int q0 = (blockIdx.x * blockDim.x + threadIdx.x) * G;
  if (q0 >= Lb) return;
  const int Lv = lens ? __ldg(lens + b) : Lb;
  const bool interior = q0 - 8 >= 0 && q0 + 12 <= Lv;
  float acc[G][S];
#pragma unroll
  for (int u = 0; u < G; ++u)
#pragma unroll
    for (int r = 0; r < S; ++r) acc[u][r] = 0.f;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    const float* xk = x + ((long long)b * S + k) * Lb;
    float xw[20];                                 // x[k, q0 - 8 + i]
    if (interior) {
      const float4* xv = reinterpret_cast<const float4*>(xk + q0 - 8);
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        const float4 f = __ldg(xv + v);
        xw[4 * v] = f.x; xw[4 * v + 1] = f.y; xw[4 * v + 2] = f.z; xw[4 * v + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 20; ++i) {
        const int n = q0 - 8 + i;
        xw[i] = (n >= 0 && n < Lv) ? __ldg(xk + n) : 0.f;   // out-of-range samples contribute an exact 0 * h
      }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const float4 c = sh[k * W + w];
      const float cr[S] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int r = 0; r < S; ++r) {
        if (S * (MLO + w) - r + P >= NT) continue;          // the one (w, r) pair without a tap: skipped, as above (compile time)
#pragma unroll
        for (int u = 0; u < G; ++u) acc[u][r] = fmaf(xw[u + w + 1], cr[r], acc[u][r]);   // n = q0 + u + MLO + w
      }
    }
  }
  float* yo = y + ((long long)b * Lb + q0) * S;
#pragma unroll
  for (int u = 0; u < G; u += 2) {
    const float v[8] = {acc[u][0], acc[u][1], acc[u][2], acc[u][3], acc[u + 1][0], acc[u + 1][1], acc[u + 1][2], acc[u + 1][3]};
    st_global_v8(yo + u * S, v);
  }
}
// host dispatch: the polyphase kernel for the reference's PQMF(subbands=4, taps=62), the generic kernel otherwise
inline cudaError_t launch_pqmf_synthesis(const float* x, const float* h, float* y, int B, int S, int taps, int Lb,
                                         const int* lens, cudaStream_t st) {
  static const bool v4_env = getenv("FV_PQMF_V4") == nullptr || atoi(getenv("FV_PQMF_V4")) != 0;   // A/B switch
  if (v4_env && S == 4 && taps == 62 && Lb % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 31) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 grid((Lb / 4 + 127) / 128, B);
    pqmf_synthesis_poly_v4_kernel<4, 63><<<grid, 128, 0, st>>>(x, h, y, Lb, lens);
  } else if (S == 4 && taps == 62 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    constexpr int QT = 4;
    dim3 grid((Lb + 256 * QT - 1) / (256 * QT), B);
    pqmf_synthesis_poly_kernel<4, 63, QT><<<grid, 256, 0, st>>>(x, h, y, Lb, lens);
  } else {
    long long n = (long long)Lb * S;
    long long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    dim3 grid((unsigned)g, B);
    pqmf_synthesis_kernel<<<grid, 256, S * (taps + 1) * sizeof(float), st>>>(x, h, y, S, taps, Lb, lens);
  }
  return cudaGetLastError();
}
// analysis: a[k,n] = sum_j xpad[S*n + j] * h[k,j], xpad = zero pad P each side; n < L/S (conv1d stride S floor)
__global__ void pqmf_analysis_kernel(const float* __restrict__ x, const float* __restrict__ h, float* __restrict__ y,
                                     int S, int taps, long long L, long long Lb) {
  extern __shared__ float sh[];
  const int nt = taps + 1;
  for (int i = threadIdx.x; i < S * nt; i += blockDim.x) sh[i] = h[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int P = taps / 2;
  const float* xb = x + (long long)b * L;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < Lb; n += (long long)gridDim.x * blockDim.x) {
    for (int k = 0; k < S; ++k) {
      const float* hk = sh + k * nt;
      float acc = 0.f;
      for (int j = 0; j < nt; ++j) {
        const long long g = n * S + j - P;
        if (g >= 0 && g < L) acc = fmaf(__ldg(xb + g), hk[j], acc);
      }
      y[((long long)b * S + k) * Lb + n] = acc;
    }
  }
}

// Same sum for the shipped shape: tap-major loop, every x sample read once and used by the S bands, every coefficient
// vector (one 16-byte shared read) used by QT outputs (the generic kernel re-reads x per band through 64-bit index
// arithmetic).  Same addition order per output (j ascending) -> identical bits.
template <int S, int NT, int QT>
__global__ void __launch_bounds__(256) pqmf_analysis_win_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                                float* __restrict__ y, int L, int Lb) {
  constexpr int P = (NT - 1) / 2;
  __shared__ float sh[NT * S];   // transposed [j][k]: the S coefficients of tap j are one 16-byte read
  for (int i = threadIdx.x; i < S * NT; i += blockDim.x) sh[(i % NT) * S + i / NT] = h[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * (blockDim.x * QT) + threadIdx.x;   // QT outputs per thread, blockDim apart
  const float* xb = x + (long long)b * L;
  float acc[QT][S];
#pragma unroll
  for (int u = 0; u < QT; ++u)
#pragma unroll
    for (int k = 0; k < S; ++k) acc[u][k] = 0.f;
  const bool interior = n0 * S - P >= 0 && (n0 + (QT - 1) * (int)blockDim.x) * S + (NT - 1) - P < L;
  if (interior) {
    const float* xq[QT];
#pragma unroll
    for (int u = 0; u < QT; ++u) xq[u] = xb + ((n0 + u * (int)blockDim.x) * S - P);
#pragma unroll 9
    for (int j = 0; j < NT; ++j) {
      float c[S];
#pragma unroll
      for (int k = 0; k < S; ++k) c[k] = sh[j * S + k];
#pragma unroll
      for (int u = 0; u < QT; ++u) {
        const float xv = __ldg(xq[u] + j);
#pragma unroll
        for (int k = 0; k < S; ++k) acc[u][k] = fmaf(xv, c[k], acc[u][k]);
      }
    }
  } else {
#pragma unroll 9
    for (int j = 0; j < NT; ++j) {
      float c[S];
#pragma unroll
      for (int k = 0; k < S; ++k) c[k] = sh[j * S + k];
#pragma unroll
      for (int u = 0; u < QT; ++u) {
        const int g = (n0 + u * (int)blockDim.x) * S + j - P;
        const float xv = (g >= 0 && g < L) ? __ldg(xb + g) : 0.f;
#pragma unroll
        for (int k = 0; k < S; ++k) acc[u][k] = fmaf(xv, c[k], acc[u][k]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < QT; ++u) {
    const int n = n0 + u * (int)blockDim.x;
    if (n >= Lb) continue;
#pragma unroll
    for (int k = 0; k < S; ++k) y[((long long)b * S + k) * Lb + n] = acc[u][k];
  }
}
// Vectorised memory side of the same sum: one thread owns 4 CONSECUTIVE outputs n0..n0+3 of all S bands.  Their inputs are the
// 75 samples x[4*n0 - 31 .. 4*n0 + 43], fetched as nineteen aligned 16-byte loads (the kernel above: 63 scalar loads per output),
// each coefficient vector is one 16-byte shared read feeding 16 FMAs, and every band row gets one 16-byte store.
// Same addition order per output (j ascending) -> identical bits.  Requires L % 4 == 0 and Lb % 4 == 0.
template <int S, int NT>
__global__ void __launch_bounds__(128) pqmf_analysis_v4_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                               float* __restrict__ y, int L, int Lb) {
  static_assert(S == 4 && NT == 63, "shipped PQMF(subbands=4, taps=62) only");
  constexpr int P = (NT - 1) / 2;
  constexpr int G = 4;
  __shared__ float4 sh[NT];   // [j] -> (h[0][j], h[1][j], h[2][j], h[3][j])
  for (int j = threadIdx.x; j < NT; j += blockDim.x) sh[j] = make_float4(h[j], h[NT + j], h[2 * NT + j], h[3 * NT + j]);
  __syncthreads();
  const int b = blockIdx.y;
  const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * G;
  if (n0 >= Lb) return;
  const float* xb = x + (long long)b * L;
  const int g0 = n0 * S - (P + 1);                 // window base: x[g0 + i], i in [0, 76)
  float xw[76];
  if (g0 >= 0 && g0 + 76 <= L) {
    const float4* xv = reinterpret_cast<const float4*>(xb + g0);
#pragma unroll
    for (int v = 0; v < 19; ++v) {
      const float4 f = __ldg(xv + v);
      xw[4 * v] = f.x; xw[4 * v + 1] = f.y; xw[4 * v + 2] = f.z; xw[4 * v + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 76; ++i) {
      const int g = g0 + i;
      xw[i] = (g >= 0 && g < L) ? __ldg(xb + g) : 0.f;
    }
  }
  float acc[G][S];
#pragma unroll
  for (int u = 0; u < G; ++u)
#pragma unroll
    for (int k = 0; k < S; ++k) acc[u][k] = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const float4 c4 = sh[j];
    const float c[S] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int u = 0; u < G; ++u)
#pragma unroll
      for (int k = 0; k < S; ++k) acc[u][k] = fmaf(xw[S * u + j + 1], c[k], acc[u][k]);   // x index 4(n0+u) + j - P
  }
#pragma unroll
  for (int k = 0; k < S; ++k)
    *reinterpret_cast<float4*>(y + ((long long)b * S + k) * Lb + n0) = make_float4(acc[0][k], acc[1][k], acc[2][k], acc[3][k]);
}
inline cudaError_t launch_pqmf_analysis(const float* x, const float* h, float* y, int B, int S, int taps, long long L,
                                        long long Lb, cudaStream_t st) {
  static const bool v4_env = getenv("FV_PQMF_V4") == nullptr || atoi(getenv("FV_PQMF_V4")) != 0;   // A/B switch
  if (v4_env && S == 4 && taps == 62 && L < 0x7fffffffLL && L % 4 == 0 && Lb % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    dim3 grid((unsigned)((Lb / 4 + 127) / 128), B);
    pqmf_analysis_v4_kernel<4, 63><<<grid, 128, 0, st>>>(x, h, y, (int)L, (int)Lb);
  } else if (S == 4 && taps == 62 && L < 0x7fffffffLL) {
    constexpr int QT = 4;
    dim3 grid((unsigned)((Lb + 256 * QT - 1) / (256 * QT)), B);
    pqmf_analysis_win_kernel<4, 63, QT><<<grid, 256, 0, st>>>(x, h, y, (int)L, (int)Lb);
  } else {
    long long g = (Lb + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    dim3 grid((unsigned)g, B);
    pqmf_analysis_kernel<<<grid, 256, S * (taps + 1) * sizeof(float), st>>>(x, h, y, S, taps, L, Lb);
  }
  return cudaGetLastError();
}

// overlap_and_add for frame_length == 2*step: out[n*step + j] = f[n][j] + f[n-1][step + j]  (modules.py:34-73)
__global__ void overlap_add_kernel(const float* __restrict__ f, float* __restrict__ out, int frames, int step) {
  const int b = blockIdx.y;
  const long long n_out = (long long)(frames + 1) * step;
  const float* fb = f + (long long)b * frames * 2 * step;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / step;
    const int j = (int)(i - n * step);
    float v = 0.f;
    if (n >= 1) v += fb[(n - 1) * 2 * step + step + j];  // index_add_ order: frame n-1's second half first
    if (n < frames) v += fb[n * 2 * step + j];
    out[(long long)b * n_out + i] = v;
  }
}

// General overlap_and_add (modules.py:34-73, any frame_length / frame_step): out[i] = sum over the frames n that cover
// sample i of f[n, i - n*step], added in increasing n — the order index_add_ visits the gcd sub-frames on the CPU.
__global__ void overlap_add_general_kernel(const float* __restrict__ f, float* __restrict__ out, int frames, int flen,
                                           int step) {
  const int b = blockIdx.y;
  const long long n_out = (long long)(frames - 1) * step + flen;
  const float* fb = f + (long long)b * frames * flen;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
    long long n_hi = i / step;
    if (n_hi > frames - 1) n_hi = frames - 1;
    long long n_lo = i - flen + 1 <= 0 ? 0 : (i - flen + 1 + step - 1) / step;
    float v = 0.f;
    for (long long n = n_lo; n <= n_hi; ++n) v += fb[n * flen + (i - n * step)];
    out[(long long)b * n_out + i] = v;
  }
}

// Basis forward(): est[b, i] = full[b, i] - full[zero, i], i < Ltrunc (full rows have Lfull samples)
__global__ void sub_broadcast_kernel(const float* __restrict__ full, const float* __restrict__ zero, float* __restrict__ out,
                                     long long Lfull, long long Ltrunc) {
  const int b = blockIdx.y;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < Ltrunc; i += (long long)gridDim.x * blockDim.x)
    out[(long long)b * Ltrunc + i] = full[(long long)b * Lfull + i] - zero[i];
}
__global__ void copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
// fp32 [B, C, L] -> "split" activation format (fv_tma.cuh): LeakyReLU(slope) applied, value = hi + lo with hi = the fp32
// value truncated to 11 significant bits (exactly an fp16), lo = fp16(value - hi); uint4 out[B][2][C/8][L] (8 halfs each).
// Same arithmetic as the in-kernel split of fv_tc.cuh (split_f16x2) so that both producers of the format agree bit for bit.
__global__ void pack_split_kernel(const float* __restrict__ x, uint4* __restrict__ out, int C, int L, float slope) {
  const int nkc = C >> 3;
  const int b = blockIdx.z, kc = blockIdx.y;
  const float* xb = x + ((long long)b * C + kc * 8) * L;
  uint4* oh = out + ((long long)(b * 2) * nkc + kc) * L;
  uint4* ol = out + ((long long)(b * 2 + 1) * nkc + kc) * L;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      const float a0 = xb[(long long)c * L + t], a1 = xb[(long long)(c + 1) * L + t];
      const float v0 = fmaxf(a0, a0 * slope), v1 = fmaxf(a1, a1 * slope);
      const float h0 = __uint_as_float(__float_as_uint(v0) & 0xFFFFE000u);
      const float h1 = __uint_as_float(__float_as_uint(v1) & 0xFFFFE000u);
      const float l0 = v0 - h0, l1 = v1 - h1;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hp[c >> 1]) : "f"(h1), "f"(h0));
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lp[c >> 1]) : "f"(l1), "f"(l0));
    }
    oh[t] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    ol[t] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}
// weight[b, l, c] = relu(x[b, c, l]) - relu(x0[c, l])   (basis_melgan.py:154-160)
// relu = 0 (use_final_nonlinear_activation=False, basis_melgan.py:120-121): the raw predictor output, no ReLU
__global__ void relu_transpose_sub_kernel(const float* __restrict__ x, const float* __restrict__ x0, float* __restrict__ out,
                                          int C, long long L, int relu) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long l0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* xb = x + (long long)b * C * L;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long long l = l0 + threadIdx.x;
    float v = 0.f;
    if (c < C && l < L) {
      const float a = xb[(long long)c * L + l], z = x0 ? x0[(long long)c * L + l] : 0.f;
      v = relu ? fmaxf(a, 0.f) - (x0 ? fmaxf(z, 0.f) : 0.f) : a - z;
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long l = l0 + r;
    const int c = c0 + threadIdx.x;
    if (c < C && l < L) out[((long long)b * L + l) * C + c] = tile[threadIdx.x][r];
  }
}

// save_wav quantiser (data/audio.py:12-14): peak -> scale -> int16 truncation
// Both passes are HBM streams: 16-byte loads (8-byte int16 stores), four independent vectors in flight per thread; the scalar
// loops take unaligned buffers and the tail.  Same arithmetic per element as the scalar form (max / one fp32 multiply + truncation).
__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  long long done = 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long n4 = n >> 2;
    long long i = tid;
    for (; i + 3 * nth < n4; i += 4 * nth) {
      const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + nth), c = __ldg(x4 + i + 2 * nth), d = __ldg(x4 + i + 3 * nth);
      m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                         fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
      m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))),
                         fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)))));
    }
    for (; i < n4; i += nth) {
      const float4 a = __ldg(x4 + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += nth) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order like uints
}
__global__ void encode16_kernel(const float* __restrict__ x, long long n, const unsigned int* __restrict__ peak_bits,
                                float rescale, short* __restrict__ out) {
  // numpy: x *= 32767 / max(0.01, max|x|) * rescale_out  (python float64 scalar, applied to a float32 array)
  const double scale = 32767.0 / fmax(0.01, (double)__uint_as_float(*peak_bits)) * (double)rescale;
  const float s = (float)scale;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  long long done = 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    short4* o4 = reinterpret_cast<short4*>(out);
    const long long n4 = n >> 2;
    auto q = [s](const float4& v) {   // astype(int16): truncation toward zero
      return make_short4((short)(int)(v.x * s), (short)(int)(v.y * s), (short)(int)(v.z * s), (short)(int)(v.w * s));
    };
    long long i = tid;
    for (; i + 3 * nth < n4; i += 4 * nth) {
      const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + nth), c = __ldg(x4 + i + 2 * nth), d = __ldg(x4 + i + 3 * nth);
      o4[i] = q(a); o4[i + nth] = q(b); o4[i + 2 * nth] = q(c); o4[i + 3 * nth] = q(d);
    }
    for (; i < n4; i += nth) o4[i] = q(__ldg(x4 + i));
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += nth) out[i] = (short)(int)(x[i] * s);
}

}  // namespace fv
