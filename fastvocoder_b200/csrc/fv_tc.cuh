// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a.
//
//   D[pos, n] += sum_{tap j} sum_{ci} A_j[pos, ci] * W_j[ci, n]        (one UMMA per 16 channels of one tap)
//
//   M = 128 time positions per UMMA (TMEM lane = position), N = output-channel tile (16..256),
//   K = 16 input channels.  The activation tile lives ONCE in shared memory as
//   [channel-chunk of 8][row][8 x fp16] (K-major, SWIZZLE_NONE canonical layout with SBO = 128 B), so
//   the operand of tap j is the same tile viewed through a descriptor whose start address is shifted by
//   (j * dilation) rows * 16 B — no im2col, no per-tap copies, any dilation.
//
// Precision: fp32 parity (<= 1e-4 on the waveform) rules out single-pass TF32/fp16/bf16 (measured 1e-3, see
// DESIGN.md), so each operand is split x = hi + lo (both fp16, hi = rn(x), lo = rn(x - hi)) and each logical
// product is three kind::f16 UMMAs (hi*hi + hi*lo + lo*hi) accumulated in fp32 TMEM: error ~2e-6.
//
// Kernels in this file: conv_tc2_kernel (persistent warp-specialised GEMM-conv, K-chunked, weight ring or resident
// weights), conv_tc3_fused_kernel (fused ResBlock1 unit).  The first, non-persistent kernel of round 1 is in git history.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "fv_kernels.cuh"
#include "fv_model.h"
#include "fv_tma.cuh"

namespace fv {

extern std::atomic<long long> g_tc_launches;

constexpr int TC_MAX_STAGES = 8;

struct TcLayer {
  bool eligible = false;
  int NT = 0;              // N tile (divides N, multiple of 16, <= 256)
  int n_tiles = 0;
  int n_pad = 0;           // N rounded up to a multiple of 16 (zero weight columns beyond the layer's N)
  int64_t img_offset = 0;  // byte offset of this layer's image in TcWeights::buf
  const uint8_t* image = nullptr;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // suspend-time hint: sleep, do not poll
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ int g_wait_hint = 1;   // 1: try_wait with a suspend-time hint, 0: plain try_wait polling, 2: hint + nanosleep back-off (FV_WAIT_HINT)
// Bounded wait: a broken pipeline traps (launch failure reported to the host) instead of hanging the GPU.  The bound
// is an iteration count, not clock64(): the polling loop of the waiting roles was a third of all executed instructions
// of the fused-unit kernel (ncu source page), competing for issue slots with the loaders / epilogues on the same SMSP.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_spin(bar, parity)) return;
  const int mode = g_wait_hint;
  uint32_t spins = 0;
  while (!(mode ? mbar_try_wait(bar, parity) : mbar_try_wait_spin(bar, parity))) {
    if (mode == 2) __nanosleep(128);
    if (++spins > (1u << 24)) {
      printf("fv_tc: mbarrier timeout tag=%d block=(%d,%d,%d) thread=%d\n", tag, blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
// Stall accounting (FV_STALL_DEBUG=1 builds of the persistent kernels): per role, cycles spent in each class of
// mbarrier wait, total cycles in the role loop and tiles walked -> dbg[(cta * 8 + role) * 8 + slot].
template <bool DBG>
struct WaitAcc {
  long long w[5];
  long long t0;
  __device__ __forceinline__ void begin() {
    if (DBG) {
#pragma unroll
      for (int i = 0; i < 5; ++i) w[i] = 0;
      t0 = clock64();
    }
  }
  __device__ __forceinline__ void wait(int k, uint32_t bar, uint32_t parity, int tag) {
    if (DBG) {
      const long long t = clock64();
      mbar_wait(bar, parity, tag);
      w[k] += clock64() - t;
    } else {
      mbar_wait(bar, parity, tag);
    }
  }
  __device__ __forceinline__ void end(long long* dbg, int role, bool writer, long long tiles) {
    if (DBG) {
      if (dbg && writer) {
        long long* o = dbg + (((long long)blockIdx.y * gridDim.x + blockIdx.x) * 8 + role) * 8;
#pragma unroll
        for (int i = 0; i < 5; ++i) o[i] = w[i];
        o[5] = clock64() - t0;
        o[6] = tiles;
      }
    }
  }
};
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Converged-warp issue: every lane of the (fully converged) issuer warp executes this, one elected lane issues.
// With warp-uniform operands the compiler keeps descriptors in uniform registers and emits a bare UTCHMMA — no
// per-instruction ELECT / BRA.U.ANY "waterfall" and no divergent-branch bookkeeping (scripts/probes/umma_issue_probe.cu:
// 40 clk per N=32 UMMA, the shared-memory operand floor, against 150-250 clk for a loop with per-UMMA index math).
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A-operand collector hints: `fill` keeps the fetched A tile in the tensor core's collector buffer, `lastuse` takes A
// from there instead of re-reading 4 KB of shared memory (SASS: UTCHMMA ...A_KEEP / ...A_REUSE).  Only valid when
// the two UMMAs are adjacent in the tensor pipe's queue -> single-issuer plans only (tc2_plan).
__device__ __forceinline__ void umma_f16_elect_fill(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_f16_elect_lastuse(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Programmatic dependent launch: the next kernel of the layer chain may start its prologue (barriers, TMEM, weight
// image) while this one drains; it touches activations only after pdl_wait() (= all prerequisite grids complete).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// tcgen05.commit tracks the MMAs of the EXECUTING thread: it must come from the same elected lane (elect.sync is
// deterministic for a given member mask).
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Two 16-column loads in flight, ONE wait (the epilogues always need the [hi*hi + lo*hi | hi*lo] pair of the dual layout)
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, uint32_t (&r)[16], uint32_t (&q)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr0)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr1)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (sm_100 "version 1" format):
//   [0,14) start>>4 | [16,30) LBO>>4 (K-chunk stride) | [32,46) SBO>>4 (8-row group stride) | [46,48) = 1
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor, kind::f16: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
inline uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  // satfinite: huge activations clamp to +-65504 instead of becoming inf
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  hi = __ushort_as_half(h);
  lo = __float2half_rn(v - __half2float(hi));
}

// Two values at once: hi = x with the low 13 mantissa bits cleared (exactly an fp16 value inside the fp16 normal
// range), lo = fp16(x - hi) where the subtraction is exact in fp32.  hi + lo carries >= 21 mantissa bits, one
// packed cvt per pair for hi and one for lo (no convert-back).  satfinite clamps |x| > 65504.
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi2, uint32_t& lo2) {
  const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
  const float h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
  const float l0 = x0 - h0, l1 = x1 - h1;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(h1), "f"(h0));   // d.hi = first src, d.lo = second
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(l1), "f"(l0));
}

// hi / lo rows (8 channels each) of the split copy -> the un-activated fp32 values: v = hi + lo (exact in fp32),
// x = min(v, v / slope) undoes LeakyReLU for 0 < slope <= 1.
__device__ __forceinline__ void unsplit8(const uint4& h, const uint4& l, float inv_slope, float* out) {
  const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hh[j]));
    const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ll[j]));
    const float v0 = fh.x + fl.x, v1 = fh.y + fl.y;
    out[2 * j] = fminf(v0, v0 * inv_slope);
    out[2 * j + 1] = fminf(v1, v1 * inv_slope);
  }
}

// Flattened fill of one A stage.  The (8-channel group kc, row r) items of the stage are dealt round-robin over the
// 256 loader threads: item i = kc * rows + r, thread ltid takes i = ltid + 256 t.  Every thread then has `per` items =
// 8 * per independent global loads in flight per round, and a stage whose row count is not a multiple of 128 (every
// tile with a halo) no longer pays a whole extra DRAM round trip for a few leftover rows.  `src(kc, ptr, slope)`
// yields the channel-group base pointer (row stride `lstride` floats between channels) and its pre-activation.
// LeakyReLU for slopes known to lie in [0, 1] (the fused-unit kernel: always an activation): one FMUL + one FMNMX
__device__ __forceinline__ float lrelu01(float v, float slope) { return fmaxf(v, v * slope); }

// EXPERIMENTAL, compile-time opt-in (-DFV_PACKED_F32=1, see fastvocoder_b200/build.py --defs): the epilogue / loader
// arithmetic on packed fp32 pairs (FADD2 / FMUL2 / FFMA2, sm_100a).  Every packed operation is the same IEEE rn operation on
// each half, in the same order as the scalar code, so results are bit-identical; the point is fewer issue slots per element
// in the instruction-bound epilogues.  Timed in one call against the default binary (gpurun r2aa): HiFi-GAN 15.56 / 15.59 vs
// 15.69 / 15.78 ms, MB-HiFi-GAN 20.96 vs 20.39 ms — inside the clock noise of the box -> the default build does not contain it.
#ifndef FV_PACKED_F32
#define FV_PACKED_F32 0
#endif
#ifndef FV_AB_OLD_LD
#define FV_AB_OLD_LD 0   // A/B switch: 1 = one tcgen05.wait::ld per 16-column TMEM load in the fused-unit epilogues
#endif
#if FV_PACKED_F32
// (a + b) + c on two lanes
__device__ __forceinline__ float2 add3_x2(float a0, float a1, float b0, float b1, float c0, float c1) {
  return __fadd2_rn(__fadd2_rn(make_float2(a0, a1), make_float2(b0, b1)), make_float2(c0, c1));
}
// LeakyReLU (slope in [0, 1]) + hi/lo fp16 split of a pair: FMUL2, 2 FMNMX, 2 LOP3, FFMA2 (x - hi), 2 F2FP
__device__ __forceinline__ void lrelu_split_x2(float2 v, float slope, uint32_t& hi2, uint32_t& lo2) {
  const float2 m = __fmul2_rn(v, make_float2(slope, slope));
  const float x0 = fmaxf(v.x, m.x), x1 = fmaxf(v.y, m.y);
  const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
  const float h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
  const float2 l = __ffma2_rn(make_float2(h0, h1), make_float2(-1.f, -1.f), make_float2(x0, x1));   // x - h, exact
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(h1), "f"(h0));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(l.y), "f"(l.x));
}
#endif

// Materialise a base pointer so that `base + int_index` becomes one IMAD.WIDE (the compiler otherwise re-associates
// the 64-bit element offsets of the whole expression: five integer instructions per global access).
template <class T>
__device__ __forceinline__ T* opaque_ptr(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}
constexpr int LD_MAX = 5;
// All offsets inside one utterance's [C, L] plane are 32-bit element indices (the planners reject planes >= 2^31
// elements): one IMAD.WIDE per global address instead of a chain of 64-bit multiplies and adds — these kernels are
// bound by the instruction stream of their loader / epilogue warps (ncu: 60-65 % issue-slot utilisation, of which the
// address arithmetic was the largest part).
// RAW (fused ResidualStack): rows [raw_lo, raw_lo + raw_rows) are ALSO stored un-activated (split hi / lo) at
// raw_hi / raw_lo_p + (kc * raw_rows + r - raw_lo) * 16 — the same loaded registers feed both copies.
template <bool ACT01 = false, bool RAW = false, class Src>
__device__ __forceinline__ void fill_stage_flat(uint8_t* A_hi, uint8_t* A_lo, int rows, int nkc, int ltid, int per,
                                                int rounds, int g0, int Lb, int lstride, bool reflect, Src src,
                                                int plane_rows = 0, uint8_t* raw_hi = nullptr, uint8_t* raw_lo_p = nullptr,
                                                int raw_lo = 0, int raw_rows = 0) {
  if (plane_rows == 0) plane_rows = rows;   // row stride of the 8-channel planes in shared memory (>= rows)
  const int nitems = nkc * rows;
  const uint32_t ls = (uint32_t)lstride;   // unsigned: keeps c * ls a 32-bit multiply feeding one IMAD.WIDE.U32
  const int kstep = 256 / rows, rstep = 256 - kstep * rows;   // item + 256 -> (kc + kstep, r + rstep) with one carry
  int i0 = ltid;
  int kc0 = i0 / rows, r0 = i0 - kc0 * rows;
  for (int rd = 0; rd < rounds; ++rd) {
    float v[LD_MAX][8];
    float sl[LD_MAX];
    int ofs[LD_MAX];
    int rofs[LD_MAX];
#pragma unroll
    for (int t = 0; t < LD_MAX; ++t) {
      const bool live = t < per && i0 < nitems;
      const int kc = live ? kc0 : 0;
      const float* xc;
      src(kc, xc, sl[t]);
      int g = g0 + r0;
      if (reflect) {
        if (g < 0) g = -g;
        if (g >= Lb) g = 2 * (Lb - 1) - g;
      }
      const bool ok = live && g >= 0 && g < Lb;
      const float* pg = opaque_ptr(xc + (ok ? g : 0));
#pragma unroll
      for (int c = 0; c < 8; ++c) v[t][c] = ok ? __ldg(pg + (uint32_t)c * ls) : 0.f;
      ofs[t] = live ? (kc * plane_rows + r0) * 16 : -1;
      if (RAW) rofs[t] = (live && r0 >= raw_lo && r0 < raw_lo + raw_rows) ? (kc * raw_rows + (r0 - raw_lo)) * 16 : -1;
      if (t < per) {   // advance to item i0 + 256
        i0 += 256; kc0 += kstep; r0 += rstep;
        if (r0 >= rows) { r0 -= rows; ++kc0; }
      }
    }
#pragma unroll
    for (int t = 0; t < LD_MAX; ++t) {
      if (ofs[t] < 0) continue;
      uint32_t hp[4], lp[4];
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        if (ACT01) split_f16x2(lrelu01(v[t][c], sl[t]), lrelu01(v[t][c + 1], sl[t]), hp[c >> 1], lp[c >> 1]);
        else split_f16x2(pre_act(v[t][c], sl[t]), pre_act(v[t][c + 1], sl[t]), hp[c >> 1], lp[c >> 1]);
      }
      *reinterpret_cast<uint4*>(A_hi + ofs[t]) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *reinterpret_cast<uint4*>(A_lo + ofs[t]) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
      if (RAW && rofs[t] >= 0) {
#pragma unroll
        for (int c = 0; c < 8; c += 2) split_f16x2(v[t][c], v[t][c + 1], hp[c >> 1], lp[c >> 1]);
        *reinterpret_cast<uint4*>(raw_hi + rofs[t]) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(raw_lo_p + rofs[t]) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
      }
    }
  }
}
// rounds / items per thread per round for a stage of `nitems` items
inline void flat_plan(int nitems, int& per, int& rounds) {
  rounds = (nitems + 256 * LD_MAX - 1) / (256 * LD_MAX);
  if (rounds < 1) rounds = 1;
  per = (nitems + 256 * rounds - 1) / (256 * rounds);
  if (per < 1) per = 1;
}

// ------------------------------------------------------------------------------------------------
// Weight image (bind time): fp32 derived image [Cin][Kd][N] -> fp16 hi/lo UMMA-canonical blocks
// ------------------------------------------------------------------------------------------------
__global__ void tc_pack_weights_kernel(const float* __restrict__ wd, uint8_t* __restrict__ img, int Cin, int Kd, int N,
                                       int Npad, int NT, int* __restrict__ range_flag) {
  const int ksteps = Cin / 16;
  const long long total = (long long)Cin * Kd * Npad;  // one thread per (padded) weight
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Npad);
    const long long q = i / Npad;
    const int j = (int)(q % Kd);
    const int ci = (int)(q / Kd);
    __half hi, lo;
    const float wv = n < N ? wd[q * N + n] : 0.f;
    if (!(fabsf(wv) <= 65504.f) && range_flag) *range_flag = 1;   // outside the fp16 hi/lo range (or NaN): the split would clamp it
    split_f16(wv, hi, lo);
    const int nt = n / NT, nn = n - nt * NT;
    const int ks = ci >> 4, kc2 = (ci >> 3) & 1, e = ci & 7;
    const long long kb = (long long)j * ksteps + ks;
    const long long blk = ((long long)nt * Kd * ksteps + kb) * ((long long)NT * 64);
    const long long off_hi = blk + ((long long)kc2 * 2 * NT + nn) * 16 + e * 2;   // [kc2][hi|lo][NT][8]
    *reinterpret_cast<__half*>(img + off_hi) = hi;
    *reinterpret_cast<__half*>(img + off_hi + (long long)NT * 16) = lo;
  }
}

inline int tc_pick_nt(int N) {
  if (N % 16) return 0;
  static const int nt_max = getenv("FV_NT_MAX") ? atoi(getenv("FV_NT_MAX")) : 256;   // tuning knob: widest N tile
  const int cap = (nt_max >= 16 && nt_max <= 256) ? nt_max / 16 * 16 : 256;
  if (N <= cap) return N;
  for (int nt = cap; nt >= 16; nt -= 16)
    if (N % nt == 0) return nt;
  return 0;
}

struct TcWeights {
  std::vector<TcLayer> layers;
  uint8_t* buf = nullptr;
  int* range_flag = nullptr;   // device int behind the images: set by the pack kernel when a weight does not fit the fp16 split

  int build(const std::vector<Layer>& ls, const float* derived, cudaStream_t st) {
    release();
    layers.assign(ls.size(), TcLayer{});
    int64_t total = 0;
    for (size_t i = 0; i < ls.size(); ++i) {
      const Layer& l = ls[i];
      TcLayer& t = layers[i];
      int npad = l.N;
      // zero-padded output columns, never stored: Basis hop (15) and the narrow output convs (conv_post 1 / 4 channels)
      if ((l.type == L_BASIS || (l.type == L_CONV && l.N < 16)) && l.N % 16) npad = (l.N + 15) / 16 * 16;
      int nt = (l.Cin % 16 == 0) ? tc_pick_nt(npad) : 0;
      if ((l.type == L_CONVT || l.type == L_UPCONV) && l.Cout % 16) nt = 0;  // a 16-column epilogue chunk must stay inside one phase
      if (nt == 0) continue;
      t.eligible = true;
      t.NT = nt;
      t.n_tiles = npad / nt;
      t.n_pad = npad;
      t.img_offset = total;
      total += ((int64_t)l.Cin * l.Kd * npad * 4 + 255) / 256 * 256;
    }
    if (total == 0) return 0;
    if (cudaMalloc(&buf, (size_t)total + 256) != cudaSuccess) return -1;
    range_flag = reinterpret_cast<int*>(buf + total);
    if (cudaMemsetAsync(range_flag, 0, sizeof(int), st) != cudaSuccess) return -1;
    for (size_t i = 0; i < ls.size(); ++i) {
      if (!layers[i].eligible) continue;
      layers[i].image = buf + layers[i].img_offset;
      const Layer& l = ls[i];
      const long long n = (long long)l.Cin * l.Kd * layers[i].n_pad;
      long long g = (n + 255) / 256;
      if (g > 148 * 16) g = 148 * 16;
      tc_pack_weights_kernel<<<(int)g, 256, 0, st>>>(derived + l.wd_offset, buf + layers[i].img_offset, l.Cin, l.Kd,
                                                     l.N, layers[i].n_pad, layers[i].NT, range_flag);
      g_launches++;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  const TcLayer* layer(int i) const { return (buf && i < (int)layers.size()) ? &layers[i] : nullptr; }
  // after the stream has been synchronised: did every weight fit the split?  (false -> the caller drops the images: exact fp32 path)
  bool in_range() const {
    int f = 0;
    if (range_flag && cudaMemcpy(&f, range_flag, sizeof f, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    return f == 0;
  }
  void release() {
    if (buf) cudaFree(buf);
    buf = nullptr;
    range_flag = nullptr;
    layers.clear();
  }
};

// ---- thread-block-cluster helpers (weight-stream multicast) ------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> the same smem offset in every CTA of `mask`; each destination CTA's mbarrier (same offset) gets complete_tx
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of the cluster (plain remote arrive, release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// =================================================================================================
// v2: persistent, warp-specialised pipeline
//   warps 0-7   loaders   : global fp32 -> act/pad -> fp16 hi/lo -> A[stage]          (a_full / a_empty)
//   warps 8-11  UMMA issue: `n_issuers` of them each own the M tiles mt = id, id + n, ... (own accumulators), so
//                           the scalar issue cost (~60-100 cycles per UMMA) is spread over several threads when
//                           the UMMAs are short (narrow layers);  tcgen05.commit per issuer    (acc_full / acc_empty)
//               weights   : warp 11: cp.async.bulk; whole layer image resident in smem when it fits (then warp 11
//                           may also issue), else a ring it keeps feeding
//   warps 12-15 epilogue  : tcgen05.ld -> bias/residual/accumulate/tanh -> global      (one TMEM lane quarter each)
// Each CTA walks tiles (utterance b, time tile) with stride gridDim.x, so the load of tile i+1, the MMAs of
// tile i and the epilogue of tile i-1 overlap.
// =================================================================================================
constexpr int TC2_LOADER_WARPS = 8;
constexpr int TC2_ISSUE_WARPS = 4;    // warps 8..11: UMMA issuers (warp 11 doubles as the weight producer)
constexpr int TC2_THREADS = (TC2_LOADER_WARPS + TC2_ISSUE_WARPS + 4) * 32;  // 8 loaders + 4 issue/producer + 4 epilogue

// `nkb` consecutive 16-channel k-blocks (A advances ks_step16, weights advance kb16 per block) on MT M tiles of this
// issuer (A rows mt_step16 / accumulator columns d_step apart).  Everything is warp-uniform; MT is unrolled so the
// per-UMMA overhead is one or two uniform adds.  DUAL: B = [B_hi | B_lo] as one N = 2*NT operand (2 UMMAs per block,
// hi*lo terms in their own accumulator columns); else three N = NT UMMAs into the same columns.
template <bool DUAL, int MT, bool REUSE = false>
__device__ __forceinline__ void issue_kblocks(uint64_t ad, uint64_t bd, uint32_t d0, int nkb, uint32_t accum,
                                              uint64_t ks_step16, uint64_t kb16, uint64_t mt_step16, uint32_t d_step,
                                              uint64_t lo_delta16, uint64_t nt16, uint32_t idesc, uint32_t idesc2) {
  // Issue order: a UMMA that accumulates into the columns the previous one wrote waits for it to drain (~150-250 clk for
  // these short N <= 128 instructions, profiles/r02_notes.md), so the products of one k-block are issued product-major over
  // the M tiles (independent accumulators back to back) instead of tile-major.
  for (int q = 0; q < nkb; ++q, ad += ks_step16, bd += kb16) {
    if (REUSE) {   // single issuer: the hi tile is fetched once for its two products (needs the pair adjacent)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const uint64_t am = ad + (uint64_t)m * mt_step16;
        const uint32_t d = d0 + (uint32_t)m * d_step;
        umma_f16_elect_fill(d, am, bd, idesc, accum);
        umma_f16_elect_lastuse(d, am, bd + nt16, idesc, 1u);
        umma_f16_elect(d, am + lo_delta16, bd, idesc, 1u);
      }
    } else if (DUAL) {
#pragma unroll
      for (int m = 0; m < MT; ++m)   // [hi*hi | hi*lo] -> cols [0,NT) | [NT,2NT)
        umma_f16_elect(d0 + (uint32_t)m * d_step, ad + (uint64_t)m * mt_step16, bd, idesc2, accum);
#pragma unroll
      for (int m = 0; m < MT; ++m)   // lo*hi -> cols [0,NT)
        umma_f16_elect(d0 + (uint32_t)m * d_step, ad + (uint64_t)m * mt_step16 + lo_delta16, bd, idesc, 1u);
    } else {
#pragma unroll
      for (int m = 0; m < MT; ++m)
        umma_f16_elect(d0 + (uint32_t)m * d_step, ad + (uint64_t)m * mt_step16, bd, idesc, accum);
#pragma unroll
      for (int m = 0; m < MT; ++m)   // + NT*16 bytes: the lo half of the block
        umma_f16_elect(d0 + (uint32_t)m * d_step, ad + (uint64_t)m * mt_step16, bd + nt16, idesc, 1u);
#pragma unroll
      for (int m = 0; m < MT; ++m)
        umma_f16_elect(d0 + (uint32_t)m * d_step, ad + (uint64_t)m * mt_step16 + lo_delta16, bd, idesc, 1u);
    }
    accum = 1u;
  }
}

struct Tc2Args {
  ConvArgs a;
  const uint8_t* wimg;
  int NT, m_tiles, rows, ksteps, kblocks;
  int a_stages, acc_stages, w_resident, kb_per_stage, w_stages, stage_bytes;
  int tmem_cols, acc_cols;
  int tiles_per_batch, total_tiles;
  int ck, nck;       // K-chunking: an A stage holds `ck` input channels of the tile; nck = Cin / ck stages per tile
  int acc_per_mt;    // single accumulator set, one M tile per issuer: accumulators are handed back per M tile (epilogue M-tile-major)
  int rows_alloc;    // row stride of the A planes (= rows; rounded up to 8 with a split / TMA-fed input: 128-B aligned boxes)
  int epi_groups;    // split input: warps 4-7 are a second epilogue group (no loader warps then)
  int cluster_mode;  // experimental (FV_CLUSTER): 1 = pairs, private weight copies; 2 = each CTA multicasts its half; 3 = rank 0 multicasts all
  int n_issuers;     // UMMA issuer warps in use (1..4; at most 3 when the weight ring needs warp 11)
  int dual;          // 1: B = [B_hi | B_lo] as one N = 2*NT operand (2 UMMAs / k-block), 0: three N = NT UMMAs
  int a_reuse;       // three-UMMA form, one issuer: A_hi stays in the collector for its second product
  int ld_per, ld_rounds;   // flattened loader: items per thread per round / rounds per A stage (0 = legacy pair loop)
  int pdl;           // launched with programmatic stream serialization
  uint32_t idesc;    // M=128, N=NT
  uint32_t idesc2;   // M=128, N=2*NT (dual)
  long long* dbg;    // stall-accounting buffer (FV_STALL_DEBUG) or nullptr
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Epilogue of one CTA tile for layers WITHOUT a residual / running-sum addend (any output layout):
// tcgen05.ld -> + bias -> [tanh] -> coalesced stores (ConvTranspose: phase interleave; narrow layers: masked tail columns).
template <int LAYOUT, bool DBG>
__device__ __forceinline__ void tc2_epilogue_tile(const Tc2Args& p, uint32_t tmem_acc, int q, int lane, int b, int t0,
                                                  int nt, WaitAcc<DBG>& wa, int grp = 0, int ngrp = 1, int mt_lo = 0,
                                                  int mt_hi = -1) {
  const ConvArgs& a = p.a;
  float* __restrict__ yb = a.y + (long long)b * a.y_bs;
  const int nchunks = p.NT >> 4;
  if (mt_hi < 0) mt_hi = p.m_tiles;
  for (int c = 0; c < nchunks; ++c) {
    const int nbase = nt * p.NT + c * 16;
    constexpr bool PHASED = (LAYOUT == OUT_PHASE || LAYOUT == OUT_PHASE_SPLIT);
    int r = 0, co0 = nbase;
    if (PHASED) { r = nbase / a.ph_cout; co0 = nbase - r * a.ph_cout; }
    float bias[16];
    if (a.bias && (PHASED || nbase + 16 <= a.N)) {
      const float4* bp = reinterpret_cast<const float4*>(a.bias + (PHASED ? co0 : nbase));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(bp + i);
        bias[4 * i] = v.x; bias[4 * i + 1] = v.y; bias[4 * i + 2] = v.z; bias[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) bias[i] = 0.f;
    }
    for (int mt = mt_lo; mt < mt_hi; ++mt) {
      if (ngrp == 2 && ((c * p.m_tiles + mt) & 1) != grp) continue;   // the lane quarter's other epilogue warp takes it
      const int pos = t0 + mt * 128 + q * 32 + lane;
      uint32_t rr[16];
      const uint32_t tcol = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.NT * (p.dual ? 2 : 1) + c * 16);
      long long tdbg = 0;
      if (DBG) tdbg = clock64();
      tmem_ld16(tcol, rr);
      if (p.dual) {  // columns [NT, 2NT) hold the separately accumulated hi*lo terms
        uint32_t r2[16];
        tmem_ld16(tcol + (uint32_t)p.NT, r2);
#pragma unroll
        for (int i = 0; i < 16; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) + __uint_as_float(r2[i]));
      }
      if (DBG) { const long long t1 = clock64(); wa.w[1] += t1 - tdbg; tdbg = t1; }   // "tmem": TMEM read + wait::ld
      int o0, ostride;   // 32-bit offsets inside the utterance's output plane (tc2_plan rejects planes >= 2^31 elements)
      bool ok = pos < a.Lpos;
      if (LAYOUT == OUT_BCL || LAYOUT == OUT_BCL_SPLIT) {
        o0 = nbase * a.Lpos + pos; ostride = a.Lpos;
      } else if (LAYOUT == OUT_BLC) {
        o0 = pos * a.N + nbase; ostride = 1;
      } else {
        const int t = pos * a.ph_stride + r - a.ph_pad;
        ok = ok && t >= 0 && t < a.ph_lout;
        o0 = co0 * a.ph_lout + t; ostride = a.ph_lout;
      }
      if (!ok) continue;
      if (LAYOUT == OUT_BCL_SPLIT) {   // h of an unfused ResBlock unit: lrelu -> fp16 hi / lo rows of the blocked planes [2][N/8][Lpos]
        uint32_t hp[8], lp[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          split_f16x2(lrelu01(__uint_as_float(rr[i]) + bias[i], a.out_slope),
                      lrelu01(__uint_as_float(rr[i + 1]) + bias[i + 1], a.out_slope), hp[i >> 1], lp[i >> 1]);
        const uint32_t ul = (uint32_t)a.Lpos, lo_off = (uint32_t)(a.N >> 3) * ul;
        uint4* yq = opaque_ptr(reinterpret_cast<uint4*>(yb) + ((uint32_t)(nbase >> 3) * ul + (uint32_t)pos));
        yq[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        yq[ul] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
        yq[lo_off] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        yq[lo_off + ul] = make_uint4(lp[4], lp[5], lp[6], lp[7]);
        if (DBG) wa.w[4] += clock64() - tdbg;
        continue;
      }
      if (LAYOUT == OUT_PHASE_SPLIT) {
        // split copy for the TMA-fed fused units: lrelu(out_slope) -> fp16 hi / lo -> rows of 16 B in the blocked planes
        // uint4 [2 (hi, lo)][Cout/8][Lout]; a 16-column chunk covers planes co0/8 and co0/8 + 1 of both halves
        uint32_t hp[8], lp[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          split_f16x2(lrelu01(__uint_as_float(rr[i]) + bias[i], a.out_slope),
                      lrelu01(__uint_as_float(rr[i + 1]) + bias[i + 1], a.out_slope), hp[i >> 1], lp[i >> 1]);
        uint4* yq = opaque_ptr(reinterpret_cast<uint4*>(yb) + ((co0 >> 3) * a.ph_lout + (o0 - co0 * a.ph_lout)));
        const uint32_t ul = (uint32_t)a.ph_lout, lo_off = (uint32_t)(a.ph_cout >> 3) * ul;
        yq[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        yq[ul] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
        yq[lo_off] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        yq[lo_off + ul] = make_uint4(lp[4], lp[5], lp[6], lp[7]);
        if (DBG) wa.w[4] += clock64() - tdbg;
        continue;
      }
      float* py = opaque_ptr(yb + o0);
      if (!PHASED && nbase + 16 > a.N) {   // zero-padded tail columns (Basis 15 of 16, conv_post 1 / 4 of 16)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (nbase + i < a.N) {
            float vv = __uint_as_float(rr[i]) + (a.bias ? __ldg(a.bias + nbase + i) : 0.f);
            if (a.post_tanh) vv = tanhf(vv);
            py[(uint32_t)i * (uint32_t)ostride] = vv;
          }
        }
        continue;
      }
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]) + bias[i];
      if (a.post_tanh) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) py[(uint32_t)i * (uint32_t)ostride] = v[i];
      if (DBG) wa.w[4] += clock64() - tdbg;      // "store": bias add + issuing the 16 stores
    }
  }
}

// Epilogue for [B, C, L] layers WITH an addend from global memory: y = conv + bias + residual [+ running MRF sum,
// / num_kernels].  The addend does not depend on the accumulators, so its loads are issued before the TMEM read:
// one memory latency per (chunk, M tile) instead of two or three back to back.  (No padded-N layers here: run_layer.)
// RS: the residual is a split-format copy (4 x 16-byte loads per 16 channels, rebuilt with unsplit8);
// OS: the output is written LeakyReLU(out_slope)-activated in the split format (the next unit's TMA-fed input).
template <bool DBG, bool RS, bool OS>
__device__ __forceinline__ void tc2_epilogue_tile_add(const Tc2Args& p, uint32_t tmem_acc, int q, int lane, int b, int t0,
                                                      int nt, WaitAcc<DBG>& wa, int grp = 0, int ngrp = 1, int mt_lo = 0,
                                                      int mt_hi = -1) {
  const ConvArgs& a = p.a;
  float* __restrict__ yb = a.y + (long long)b * a.y_bs;
  const float* __restrict__ rb = a.res ? a.res + (long long)b * a.res_bs : nullptr;
  const int nchunks = p.NT >> 4;
  if (mt_hi < 0) mt_hi = p.m_tiles;
  const int mts = mt_hi - mt_lo;
  const int n_it = nchunks * mts;                // iteration = (chunk c, M tile mt), mt fastest
  // xs / num_kernels (hifigan.py:103) as a multiply by the fp32 reciprocal: <= 1 ulp from the division, far inside 1e-4
  const float inv = (a.acc_mode == ACC_ADD_DIV || a.acc_mode >= ACC_STORE_SCALE) ? 1.0f / a.acc_div : 1.0f;
  const bool red = a.acc_mode == ACC_RED_SCALE;
  const int ostride = a.Lpos;   // 32-bit offsets inside the utterance plane (tc2_plan)
  const uint32_t uos = (uint32_t)ostride;
  const uint32_t q_lo = (uint32_t)(a.N >> 3) * uos;   // split planes: lo half starts N/8 planes after the hi half
  const int row = q * 32 + lane;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  // The residual of this group's NEXT iteration is requested before the current one is processed (software pipelining:
  // two sets of loads in flight per warp; these warps are latency-bound, not bandwidth-bound).
  float nxt[RS ? 1 : 16];
  uint4 nq[RS ? 4 : 1];
  auto fetch = [&](int it2) {
    const int c2 = it2 / mts, mt2 = mt_lo + (it2 - c2 * mts);
    const int pos2 = t0 + mt2 * 128 + row;
    const bool ok2 = rb != nullptr && pos2 < a.Lpos;
    const int nb2 = nt * p.NT + c2 * 16;
    if (RS) {
      const uint4* pr = opaque_ptr(reinterpret_cast<const uint4*>(rb) + (ok2 ? (uint32_t)(nb2 >> 3) * uos + (uint32_t)pos2 : 0u));
      nq[0] = ok2 ? __ldg(pr) : z4; nq[1] = ok2 ? __ldg(pr + uos) : z4;
      nq[2] = ok2 ? __ldg(pr + q_lo) : z4; nq[3] = ok2 ? __ldg(pr + q_lo + uos) : z4;
    } else {
      const float* pr = opaque_ptr(rb + (ok2 ? nb2 * a.Lpos + pos2 : 0));
#pragma unroll
      for (int i = 0; i < 16; ++i) nxt[i] = ok2 ? __ldg(pr + (uint32_t)i * uos) : 0.f;
    }
  };
  if (grp < n_it) fetch(grp);
  for (int it = grp; it < n_it; it += ngrp) {
    const int c = it / mts, mt = mt_lo + (it - c * mts);
    const int nbase = nt * p.NT + c * 16;
    const int pos = t0 + mt * 128 + row;
    const bool ok = pos < a.Lpos;
    const int o0 = nbase * a.Lpos + pos;
    float addend[16];
    long long tld = 0;
    if (DBG) tld = clock64();
    if (RS) {
      unsplit8(nq[0], nq[2], a.res_inv_slope, addend);
      unsplit8(nq[1], nq[3], a.res_inv_slope, addend + 8);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) addend[i] = nxt[i];
    }
    if (it + ngrp < n_it) fetch(it + ngrp);
    float* py = opaque_ptr(yb + (ok ? o0 : 0));
    if (OS && a.ysum) {   // last MRF branch: (v + ysum * num_kernels) / num_kernels, loads in flight before the TMEM read
      const float* ps = opaque_ptr(a.ysum + (long long)b * a.res_bs + (ok ? o0 : 0));
#pragma unroll
      for (int i = 0; i < 16; ++i) addend[i] = fmaf(ok ? __ldg(ps + (uint32_t)i * uos) : 0.f, a.acc_div, addend[i]);
    }
    if (!OS && acc_reads_y(a.acc_mode)) {
#pragma unroll
      for (int i = 0; i < 16; ++i) addend[i] += ok ? py[(uint32_t)i * uos] : 0.f;
    }
    uint32_t rr[16];
    const uint32_t tcol = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.NT * (p.dual ? 2 : 1) + c * 16);
    long long tdbg = 0;
    if (DBG) { tdbg = clock64(); wa.w[3] += tdbg - tld; }   // "ldissue": issuing the residual / running-sum loads
    if (p.dual) {
      uint32_t r2[16];
      tmem_ld16x2(tcol, tcol + (uint32_t)p.NT, rr, r2);
#pragma unroll
      for (int i = 0; i < 16; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) + __uint_as_float(r2[i]));
    } else {
      tmem_ld16(tcol, rr);
    }
    if (DBG) {
      const long long t1 = clock64();
      wa.w[1] += t1 - tdbg;                      // "tmem": TMEM read + wait::ld
      float sink = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) sink += addend[i];
      if (sink == 123.456f) wa.w[4] += 1;         // forces the addend loads to have landed
      tdbg = clock64();
      wa.w[2] += tdbg - t1;                      // "addwait": residual / running-sum loads still in flight after the TMEM read
    }
    if (ok) {
      if (a.bias) {
        const float4* bp = reinterpret_cast<const float4*>(a.bias + nbase);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 bv = __ldg(bp + i4);
          addend[4 * i4] += bv.x; addend[4 * i4 + 1] += bv.y; addend[4 * i4 + 2] += bv.z; addend[4 * i4 + 3] += bv.w;
        }
      }
      if (OS) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (__uint_as_float(rr[i]) + addend[i]) * inv;
        uint32_t hp[8], lp[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          split_f16x2(lrelu01(v[i], a.out_slope), lrelu01(v[i + 1], a.out_slope), hp[i >> 1], lp[i >> 1]);
        uint4* yq = opaque_ptr(reinterpret_cast<uint4*>(yb) + ((uint32_t)(nbase >> 3) * uos + (uint32_t)pos));
        yq[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        yq[uos] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
        yq[q_lo] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        yq[q_lo + uos] = make_uint4(lp[4], lp[5], lp[6], lp[7]);
      } else if (red) {
#pragma unroll
        for (int i = 0; i < 16; ++i) red_add_f32(py + (uint32_t)i * uos, (__uint_as_float(rr[i]) + addend[i]) * inv);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) py[(uint32_t)i * uos] = (__uint_as_float(rr[i]) + addend[i]) * inv;
      }
    }
    if (DBG) wa.w[4] += clock64() - tdbg;        // "store": bias add + issuing the 16 stores
  }
}

// XS: the input is a split-format buffer (fv_tma.cuh) fetched by TMA — no loader warps: warp 0 is the TMA producer and
// warps 4-7 a second epilogue group.
template <bool DBG, bool XS>
__global__ void __launch_bounds__(TC2_THREADS, 1) conv_tc2_kernel(const Tc2Args p, const __grid_constant__ CUtensorMap tm_main,
                                                                  const __grid_constant__ CUtensorMap tm_tail) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  const ConvArgs& a = p.a;
  const int rows = p.rows;
  const uint32_t a_bytes = (uint32_t)p.rows_alloc * p.ck * 2;   // hi or lo of one A stage (one channel chunk of the tile)
  const int kblock_bytes = p.NT * 64;
  const int kpc = p.ck >> 4;                              // 16-channel k-steps per chunk
  const int runs_per_grp = (kpc + p.kb_per_stage - 1) / p.kb_per_stage;   // ring stages per (chunk, tap)
  uint8_t* Abuf = smem;                                   // [a_stages][hi|lo][a_bytes]
  uint8_t* Wbuf = smem + (size_t)p.a_stages * 2 * a_bytes;
  const size_t w_bytes = p.w_resident ? (size_t)p.kblocks * kblock_bytes : (size_t)p.w_stages * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wbuf + w_bytes);
  // barrier map: [0,2) a_full  [2,4) a_empty  [4,6) acc_full  [6,8) acc_empty  [8,16) w_full  [16,24) w_empty
  //              [24,32) w_free (cluster mode: every CTA's issuers are done with the slot)
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  //              [32,40) acc_mt_empty[acc stage][M tile] (acc_per_mt plans: the epilogue hands the accumulators back per M tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform -> uniform-register code
  const int nt = blockIdx.y;
  const int M = p.m_tiles * 128;
  if (p.pdl) pdl_launch_dependents();   // the next layer's CTAs may take over SMs as soon as ours retire
  const uint8_t* wsrc = p.wimg + (size_t)nt * p.kblocks * kblock_bytes;
  // Weight-ring multicast: the CTAs of a cluster (launched as pairs for ring-mode layers) walk the same number of
  // ring iterations; each fetches 1/cs of every stage and multicasts it to all of them.
  const uint32_t cs = cluster_nctarank(), cr = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  const int cl_base = (int)blockIdx.x - (int)cr;                      // lowest blockIdx.x of my cluster
  const int ring_tiles = (p.total_tiles > cl_base) ? (p.total_tiles - cl_base + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  const int epi_groups = (XS && p.epi_groups == 2) ? 2 : 1;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(0 + s), XS ? 1 : TC2_LOADER_WARPS);   // a_full: one arrive per loader warp / one expect_tx + the TMA bytes
      mbar_init(BAR(2 + s), p.n_issuers);   // a_empty: one tcgen05.commit per issuer
      mbar_init(BAR(4 + s), p.n_issuers);   // acc_full: one tcgen05.commit per issuer
      mbar_init(BAR(6 + s), 4 * epi_groups);   // acc_empty: one arrive per epilogue warp
    }
    for (int s = 0; s < 8; ++s) mbar_init(BAR(32 + s), 4 * epi_groups);   // acc_mt_empty: one arrive per epilogue warp
    for (int s = 0; s < 8; ++s) {
      mbar_init(BAR(8 + s), 1);                // w_full: expect_tx by the producer
      mbar_init(BAR(16 + s), p.n_issuers);     // w_empty: one (local) tcgen05.commit per issuer
      mbar_init(BAR(24 + s), cs);               // w_free: one remote arrive per CTA of the cluster
    }
    fence_mbar_init();
  }
  if (warp == TC2_LOADER_WARPS) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // peers' barriers must be initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool is_epi = warp >= TC2_LOADER_WARPS + TC2_ISSUE_WARPS || (epi_groups == 2 && warp >= 4 && warp < TC2_LOADER_WARPS);
  if (XS && warp == 0) {
    // ------------------------------------------------------------------ TMA producer: split planes -> A[stage]
    const int nkc = p.ck >> 3, planes_b = a.Cin >> 3;
    const int nfull = rows / TMA_SPLIT_RB, tail = rows - nfull * TMA_SPLIT_RB;
    const int nrb = nfull + (tail ? 1 : 0);
    const int per_half = nkc * nrb, nops = 2 * per_half;
    const uint32_t stage_tx = (uint32_t)(2 * nkc * rows * 16);
    if (lane == 0) { tma_prefetch_desc(&tm_main); tma_prefetch_desc(&tm_tail); }
    int u = 0;
    WaitAcc<DBG> wa;
    wa.begin();
    if (p.pdl) pdl_wait();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_batch;
      const int g0 = (tile - b * p.tiles_per_batch) * M - a.pad_left;
      for (int ch = 0; ch < p.nck; ++ch, ++u) {
        const int s = u % p.a_stages;
        if (u >= p.a_stages) wa.wait(0, BAR(2 + s), (uint32_t)((u / p.a_stages - 1) & 1), 400 + s);
        if (lane == 0) mbar_expect_tx(BAR(0 + s), stage_tx);
        __syncwarp();
        const uint32_t a_hi = smem_u32(Abuf + (size_t)s * 2 * a_bytes);
        for (int i = lane; i < nops; i += 32) {
          const int half = i >= per_half ? 1 : 0;
          const int rem = i - half * per_half;
          const int kc = rem / nrb, rb = rem - kc * nrb;
          const uint32_t dst = a_hi + (uint32_t)half * a_bytes + (uint32_t)(kc * p.rows_alloc + rb * TMA_SPLIT_RB) * 16u;
          tma_load_2d(dst, rb < nfull ? &tm_main : &tm_tail, 2 * (g0 + rb * TMA_SPLIT_RB),
                      (b * 2 + half) * planes_b + ch * nkc + kc, BAR(0 + s));
        }
      }
    }
    wa.end(p.dbg, 0, lane == 0, u);
  } else if (!XS && warp < TC2_LOADER_WARPS) {
    // ------------------------------------------------------------------ loaders
    const int nkc = p.ck >> 3;                  // 8-channel groups per chunk
    const int nrb = (rows + 127) >> 7;          // row blocks of 128 (4 rows per lane)
    const int npairs = nkc * nrb;
    int u = 0;                                  // A-stage unit counter: (tile, channel chunk)
    WaitAcc<DBG> wa;
    wa.begin();
    if (p.pdl) pdl_wait();                      // activations of the previous layer are complete and visible
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_batch;
      const int t0 = (tile - b * p.tiles_per_batch) * M;
      const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
      const int Lb = a.lens ? __ldg(a.lens + b) : a.Lin;   // ragged batches: valid length of this utterance
      for (int ch = 0; ch < p.nck; ++ch, ++u) {
        const int s = u % p.a_stages;
        if (u >= p.a_stages) wa.wait(0, BAR(2 + s), (uint32_t)((u / p.a_stages - 1) & 1), 400 + s);
        uint8_t* A_hi = Abuf + (size_t)s * 2 * a_bytes;
        uint8_t* A_lo = A_hi + a_bytes;
        if (a.x1_split && ch * p.ck < a.cin_split) {
          // hybrid stage: this chunk of the pair layer's first input (h, split format) arrives by TMA — warp 0 issues the
          // boxes, its expect_tx is its arrival; the other loader warps just arrive (the barrier still counts 8 + the bytes)
          if (warp == 0) {
            const int nfull = rows / TMA_SPLIT_RB, tail = rows - nfull * TMA_SPLIT_RB;
            const int nrb = nfull + (tail ? 1 : 0);
            const int per_half = nkc * nrb, nops = 2 * per_half, planes_b = a.cin_split >> 3;
            const int g0 = t0 - a.pad_left;
            if (lane == 0) mbar_expect_tx(BAR(0 + s), (uint32_t)(2 * nkc * rows * 16));
            __syncwarp();
            const uint32_t a_hi = smem_u32(A_hi);
            for (int i = lane; i < nops; i += 32) {
              const int half = i >= per_half ? 1 : 0;
              const int rem = i - half * per_half;
              const int kc = rem / nrb, rb = rem - kc * nrb;
              const uint32_t dst = a_hi + (uint32_t)half * a_bytes + (uint32_t)(kc * p.rows_alloc + rb * TMA_SPLIT_RB) * 16u;
              tma_load_2d(dst, rb < nfull ? &tm_main : &tm_tail, 2 * (g0 + rb * TMA_SPLIT_RB),
                          (b * 2 + half) * planes_b + ch * nkc + kc, BAR(0 + s));
            }
          } else {
            if (lane == 0) mbar_arrive(BAR(0 + s));
          }
          continue;
        }
        if (p.ld_per > 0) {
          fill_stage_flat(A_hi, A_lo, rows, nkc, warp * 32 + lane, p.ld_per, p.ld_rounds, t0 - a.pad_left, Lb,
                          a.Lin, a.pad_mode == PAD_REFLECT, [&](int kc, const float*& xc, float& slope) {
                            const int cg = ch * p.ck + kc * 8;
                            const bool second = a.cin_split > 0 && cg >= a.cin_split;
                            xc = second ? a.x2 + (long long)b * a.x2_bs + (long long)(cg - a.cin_split) * a.Lin
                                        : xb + (long long)cg * a.Lin;
                            slope = second ? a.pre_slope2 : a.pre_slope;
                          }, p.rows_alloc);
        } else
        for (int pr = warp; pr < npairs; pr += TC2_LOADER_WARPS) {
          const int kc = pr / nrb, rbk = pr - kc * nrb;
          const int cg = ch * p.ck + kc * 8;                     // first of 8 global input channels of this group
          const bool second = a.cin_split > 0 && cg >= a.cin_split;   // two-input (pair) layers: channels >= split come from x2
          const float* __restrict__ xc = second ? a.x2 + (long long)b * a.x2_bs + (long long)(cg - a.cin_split) * a.Lin
                                                : xb + (long long)cg * a.Lin;
          const float slope = second ? a.pre_slope2 : a.pre_slope;
          float v[4][8];
          int rrow[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int r = rbk * 128 + t * 32 + lane;
            rrow[t] = r;
            int g = t0 - a.pad_left + r;
            if (a.pad_mode == PAD_REFLECT) {
              if (g < 0) g = -g;
              if (g >= Lb) g = 2 * (Lb - 1) - g;
            }
            const bool ok = r < rows && g >= 0 && g < Lb;
            const float* pg = opaque_ptr(xc + (ok ? g : 0));
#pragma unroll
            for (int c = 0; c < 8; ++c) v[t][c] = ok ? __ldg(pg + (uint32_t)c * (uint32_t)a.Lin) : 0.f;
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (rrow[t] >= rows) continue;
            uint32_t hp[4], lp[4];
#pragma unroll
            for (int c = 0; c < 8; c += 2) split_f16x2(pre_act(v[t][c], slope), pre_act(v[t][c + 1], slope),
                                                       hp[c >> 1], lp[c >> 1]);
            const uint32_t off = ((uint32_t)kc * p.rows_alloc + rrow[t]) * 16;
            *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(0 + s));
      }
      // The epilogue of this tile (one or two tiles from now) reads the residual and, for the MRF sum, the running
      // output.  Its 4 warps cannot keep enough DRAM requests in flight, so the loaders pull those lines into L2 now.
      if (a.out_layout == OUT_BCL && (a.res != nullptr || a.acc_mode != ACC_STORE)) {
        const int lines_per_row = (M * 4 + 127) >> 7;
        const int total = p.NT * lines_per_row;
        const int ltid = warp * 32 + lane;
        for (int i = ltid; i < total; i += TC2_LOADER_WARPS * 32) {
          const int n = i / lines_per_row, ln = i - n * lines_per_row;
          const long long pos = (long long)t0 + ln * 32;
          if (pos >= a.Lpos) continue;
          const long long o = (long long)(nt * p.NT + n) * a.Lpos + pos;
          if (a.res) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + (long long)b * a.res_bs + o));
          if (acc_reads_y(a.acc_mode)) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.y + (long long)b * a.y_bs + o));
        }
      }
    }
    wa.end(p.dbg, 0, warp == 0 && lane == 0, u);
  } else if (warp >= TC2_LOADER_WARPS && warp < TC2_LOADER_WARPS + TC2_ISSUE_WARPS) {
    // ------------------------------------------------------------------ weight producer + UMMA issuers
    const int wid = warp - TC2_LOADER_WARPS;                 // 0..3
    const bool is_producer = (wid == TC2_ISSUE_WARPS - 1);
    const bool L0 = (lane == 0);   // the whole warp walks the loops (uniform control flow), lane 0 issues
    {
      if (is_producer) {
        WaitAcc<DBG> wa;
        wa.begin();
        if (p.w_resident) {
          const uint32_t total = (uint32_t)p.kblocks * kblock_bytes;
          if (L0) mbar_expect_tx(BAR(8), total);
          for (uint32_t off = 0; off < total; off += 32768) {
            const uint32_t n = min(32768u, total - off);
            if (L0) bulk_g2s(smem_u32(Wbuf + off), wsrc + off, n, BAR(8));
          }
        } else {
          int g = 0;
          for (int rt = 0; rt < ring_tiles; ++rt) {
            for (int ch = 0; ch < p.nck; ++ch)
              for (int j = 0; j < a.K; ++j)
                for (int run = 0; run < runs_per_grp; ++run, ++g) {   // k-blocks (tap j, k-steps of chunk ch) are contiguous
                  const int slot = g % p.w_stages;
                  if (g >= p.w_stages) {
                    const uint32_t ph = (uint32_t)((g / p.w_stages - 1) & 1);
                    wa.wait(0, BAR(16 + slot), ph, 500 + slot);          // my issuers are done with the slot
                    if (cs > 1 && p.cluster_mode >= 2) {                 // ... tell every CTA, then wait for all of them
                      if (L0) for (uint32_t r = 0; r < cs; ++r) mbar_arrive_remote(BAR(24 + slot), r);
                      wa.wait(1, BAR(24 + slot), ph, 520 + slot);
                    }
                  }
                  const int kb0 = j * p.ksteps + ch * kpc + run * p.kb_per_stage;
                  const int nkb = min(p.kb_per_stage, kpc - run * p.kb_per_stage);
                  const uint32_t bytes = (uint32_t)nkb * kblock_bytes;
                  if (L0) {
                    mbar_expect_tx(BAR(8 + slot), bytes);   // the full stage lands here: my slice + the peers' slices
                    const uint32_t dst = smem_u32(Wbuf + (size_t)slot * p.stage_bytes);
                    const uint8_t* src = wsrc + (size_t)kb0 * kblock_bytes;
                    if (cs == 1 || p.cluster_mode == 1) {
                      bulk_g2s(dst, src, bytes, BAR(8 + slot));
                    } else if (p.cluster_mode == 2) {
                      const uint32_t slice = bytes / cs;     // k-block bytes are a multiple of 1024
                      bulk_g2s_mcast(dst + cr * slice, src + (size_t)cr * slice, slice, BAR(8 + slot), cmask);
                    } else if (cr == 0) {
                      bulk_g2s_mcast(dst, src, bytes, BAR(8 + slot), cmask);
                    }
                  }
                }
          }
        }
        wa.end(p.dbg, 1, L0, ring_tiles);
      }
      if (wid < p.n_issuers) {   // (in ring mode n_issuers <= 3, so the producer never gets here)
        WaitAcc<DBG> wa;
        wa.begin();
        const uint32_t a_lbo = (uint32_t)p.rows_alloc * 16;
        const uint32_t b_lbo = (uint32_t)p.NT * 32;
        const uint32_t wbase = smem_u32(Wbuf);
        if (p.w_resident) wa.wait(0, BAR(8), 0, 600);
        // descriptor templates: only the 14-bit start-address field (units of 16 B) changes between UMMAs
        const uint64_t a_tmpl = make_kmajor_desc(0, a_lbo, 128);
        const uint64_t b_tmpl = make_kmajor_desc(0, b_lbo, 128);
        const uint32_t acc_mt_cols = (uint32_t)(p.NT * (p.dual ? 2 : 1));
        const uint32_t a_lo_delta = a_bytes >> 4;
        const uint32_t mt_step16 = 128u * (uint32_t)p.n_issuers;          // A rows between my consecutive M tiles
        const uint32_t d_step = acc_mt_cols * (uint32_t)p.n_issuers;
        const uint32_t idesc = p.idesc, idesc2 = p.idesc2;
        const int K = a.K, dil = a.dil, m_tiles = p.m_tiles, n_iss = p.n_issuers;
        const bool dual = p.dual != 0, resident = p.w_resident != 0, reuse = p.a_reuse != 0;
        const uint32_t kb16 = (uint32_t)kblock_bytes >> 4;                  // one weight k-block, in 16-B units
        const uint64_t nt16 = (uint64_t)p.NT;
        const uint64_t ks_step16 = 2ull * (uint64_t)p.rows_alloc;           // next 16-channel k-step of the A stage
        const int my_mts = (m_tiles - wid + n_iss - 1) / n_iss;             // M tiles wid, wid + n_iss, ... (<= 4, see tc2_plan)
        __syncwarp();                                                       // elect.sync needs the full warp converged
        int it = 0, g = 0, u = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
          const int as = it % p.acc_stages;
          // acc_per_mt: this issuer owns exactly one M tile and restarts as soon as the epilogue has drained THAT tile's columns
          if (it >= p.acc_stages)
            wa.wait(1, p.acc_per_mt ? BAR(32 + as * 4 + wid) : BAR(6 + as), (uint32_t)((it / p.acc_stages - 1) & 1), 620 + as);
          const uint32_t acc = tmem_base + (uint32_t)(as * p.acc_cols) + (uint32_t)wid * acc_mt_cols;
          for (int ch = 0; ch < p.nck; ++ch, ++u) {
            const int s = u % p.a_stages;
            wa.wait(2, BAR(0 + s), (uint32_t)((u / p.a_stages) & 1), 610 + s);
            tc_fence_after();
            const uint32_t a_hi0 = smem_u32(Abuf + (size_t)s * 2 * a_bytes);
            const uint64_t ad_mine = a_tmpl + (uint64_t)((a_hi0 >> 4) & 0x3FFF) + (uint64_t)(wid * 128);   // 14-bit field: in a cluster
            // the shared-window address carries the CTA rank in its upper bits, which must not spill into the LBO field
            // one 16-channel k-block of tap j (k-step ksl inside this chunk) on all my M tiles
            // Tight, warp-converged issue (umma_f16_elect): descriptors advance incrementally in uniform registers.
            // One 16-channel k-block (A rows at `ad_hi`, weights at `bd_hi`) on all my M tiles:
            auto do_kblocks = [&](uint64_t ad, uint64_t bd, int nkb, uint32_t accum) {
#define FV_ISSUE(D, M) issue_kblocks<D, M>(ad, bd, acc, nkb, accum, ks_step16, (uint64_t)kb16, (uint64_t)mt_step16, d_step, \
                                           (uint64_t)a_lo_delta, nt16, idesc, idesc2)
              if (dual) {
                switch (my_mts) {
                  case 1: FV_ISSUE(true, 1); break;
                  case 2: FV_ISSUE(true, 2); break;
                  case 3: FV_ISSUE(true, 3); break;
                  default: FV_ISSUE(true, 4); break;
                }
              } else if (reuse) {
#define FV_ISSUE_R(M) issue_kblocks<false, M, true>(ad, bd, acc, nkb, accum, ks_step16, (uint64_t)kb16, (uint64_t)mt_step16, d_step, \
                                                    (uint64_t)a_lo_delta, nt16, idesc, idesc2)
                switch (my_mts) {
                  case 1: FV_ISSUE_R(1); break;
                  case 2: FV_ISSUE_R(2); break;
                  case 3: FV_ISSUE_R(3); break;
                  default: FV_ISSUE_R(4); break;
                }
#undef FV_ISSUE_R
              } else {
                switch (my_mts) {
                  case 1: FV_ISSUE(false, 1); break;
                  case 2: FV_ISSUE(false, 2); break;
                  case 3: FV_ISSUE(false, 3); break;
                  default: FV_ISSUE(false, 4); break;
                }
              }
#undef FV_ISSUE
            };
            uint32_t accum = ch ? 1u : 0u;     // the very first k-block of the tile overwrites the accumulators
            for (int j = 0; j < K; ++j) {
              uint64_t ad = ad_mine + (uint64_t)(j * dil);                       // tap shift: j*dil rows of 16 B
              if (resident) {
                const uint64_t bd = b_tmpl + (uint64_t)(((wbase >> 4) + (uint32_t)(j * p.ksteps + ch * kpc) * kb16) & 0x3FFF);
                do_kblocks(ad, bd, kpc, accum);
                accum = 1u;
              } else {
                for (int run = 0; run < runs_per_grp; ++run, ++g) {
                  const int slot = g % p.w_stages;
                  wa.wait(3, BAR(8 + slot), (uint32_t)((g / p.w_stages) & 1), 630 + slot);
                  tc_fence_after();
                  const int nkb = min(p.kb_per_stage, kpc - run * p.kb_per_stage);
                  const uint64_t bd = b_tmpl + (uint64_t)(((wbase + (uint32_t)slot * p.stage_bytes) >> 4) & 0x3FFF);
                  do_kblocks(ad, bd, nkb, accum);
                  accum = 1u;
                  ad += (uint64_t)nkb * ks_step16;
                  umma_commit_elect(BAR(16 + slot));   // local; the producers exchange "slot free" across the cluster
                }
              }
            }
            umma_commit_elect(BAR(2 + s));    // this A stage may be overwritten once these UMMAs have read it
          }
          umma_commit_elect(BAR(4 + as));     // my accumulators of this tile are complete
        }
        // a cluster peer may have one more tile than I do: keep consuming / releasing the shared weight ring
        if (!p.w_resident) {
          const int iters_per_tile = p.nck * a.K * runs_per_grp;
          for (int rt = it; rt < ring_tiles; ++rt) {
            for (int wi = 0; wi < iters_per_tile; ++wi, ++g) {
              const int slot = g % p.w_stages;
              wa.wait(4, BAR(8 + slot), (uint32_t)((g / p.w_stages) & 1), 640 + slot);
              umma_commit_elect(BAR(16 + slot));
            }
          }
        }
        wa.end(p.dbg, 2, wid == 0 && L0, it);
      }
    }
    __syncwarp();
  } else if (is_epi) {
    // ------------------------------------------------------------------ epilogue warps (one or two per TMEM lane quarter)
    const int q = warp & 3;
    const int grp = warp < TC2_LOADER_WARPS ? 1 : 0;
    int it = 0;
    WaitAcc<DBG> wa;
    wa.begin();
    if (p.pdl) pdl_wait();   // residual / running-sum reads and all stores come after the previous layer
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it % p.acc_stages;
      wa.wait(0, BAR(4 + as), (uint32_t)((it / p.acc_stages) & 1), 700 + as);
      tc_fence_after();
      const int b = tile / p.tiles_per_batch;
      const int t0 = (tile - b * p.tiles_per_batch) * M;
      const uint32_t acc = tmem_base + (uint32_t)(as * p.acc_cols);
      const bool has_add = a.res != nullptr || a.acc_mode != ACC_STORE;
      auto run_epi = [&](int m0, int m1) {
        if (a.out_layout == OUT_BCL_SPLIT && has_add) tc2_epilogue_tile_add<DBG, true, true>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
        else if (a.out_layout == OUT_BCL_SPLIT) tc2_epilogue_tile<OUT_BCL_SPLIT, DBG>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
        else if (a.out_layout == OUT_BCL && has_add && a.res_split) tc2_epilogue_tile_add<DBG, true, false>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
        else if (a.out_layout == OUT_BCL && has_add) tc2_epilogue_tile_add<DBG, false, false>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
        else if (a.out_layout == OUT_BCL) tc2_epilogue_tile<OUT_BCL, DBG>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
        else if (a.out_layout == OUT_BLC) tc2_epilogue_tile<OUT_BLC, DBG>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
        else if (a.out_layout == OUT_PHASE_SPLIT) tc2_epilogue_tile<OUT_PHASE_SPLIT, DBG>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
        else tc2_epilogue_tile<OUT_PHASE, DBG>(p, acc, q, lane, b, t0, nt, wa, grp, epi_groups, m0, m1);
      };
      if (p.acc_per_mt) {   // M tile by M tile: each tile's issuer restarts while the next tile is still being drained
        for (int m = 0; m < p.m_tiles; ++m) {
          run_epi(m, m + 1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(32 + as * 4 + m));
        }
        continue;
      }
      run_epi(0, p.m_tiles);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(6 + as));
    }
    wa.end(p.dbg, 3, q == 0 && lane == 0 && grp == 0, it);
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // no CTA may exit while a peer can still multicast into it
  if (warp == TC2_LOADER_WARPS) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

inline size_t tc2_smem_bytes(const Tc2Args& p) {
  const size_t a_bytes = (size_t)p.rows_alloc * p.ck * 2;
  const size_t w_bytes = p.w_resident ? (size_t)p.kblocks * p.NT * 64 : (size_t)p.w_stages * p.stage_bytes;
  return p.a_stages * 2 * a_bytes + w_bytes + 41 * 8;
}

// Choose tile shape / buffering for one layer launch.  Preference order: weights resident in smem (no L2
// re-streaming), double-buffered A, as many 128-row M tiles per CTA tile as TMEM (2 accumulator sets) allows.
inline bool tc2_plan(const ConvArgs& a, const TcLayer& L, Tc2Args& p, int num_sms) {
  const int NT = L.NT;
  const int halo = (a.K - 1) * a.dil;
  const int kblock_bytes = NT * 64;
  const int ksteps = a.Cin / 16;
  const int kblocks = a.K * ksteps;
  const long long w_total = (long long)kblocks * kblock_bytes;
  const long long BUDGET = 225 * 1024;
  static const int force_dual = getenv("FV_TC2_DUAL") ? atoi(getenv("FV_TC2_DUAL")) : -1;   // tuning knob: 0 / 1
  static const int force_mt_xs = getenv("FV_TC2_MT_XS") ? atoi(getenv("FV_TC2_MT_XS")) : 0;   // tuning knob: M tiles per CTA tile, split-input layers
  if ((a.res != nullptr || a.acc_mode != ACC_STORE) && ((a.out_layout != OUT_BCL && a.out_layout != OUT_BCL_SPLIT) || a.post_tanh)) return false;
  if (a.out_layout == OUT_BCL_SPLIT &&
      (!(a.acc_mode == ACC_STORE || (a.acc_mode == ACC_STORE_SCALE && a.res != nullptr)) || a.N % 16 || L.n_pad != a.N || a.post_tanh || !(a.out_slope >= 0.f && a.out_slope <= 1.f) ||
       (a.res != nullptr && !a.res_split)))
    return false;
  if (a.res_split && (a.res == nullptr || a.N % 16 || !(a.res_inv_slope >= 1.f))) return false;
  // split (TMA-fed) input: zero padding only (OOB rows of the tensor map), single dense input, no ragged batch
  if (a.x_split && (a.pad_mode != PAD_ZERO || a.cin_split != 0 || a.lens != nullptr || a.Cin % 16 ||
                    a.x_bs != (long long)a.Cin * a.Lin || !tma_encode_fn()))
    return false;
  if (a.x1_split && (a.x_split || a.cin_split <= 0 || a.cin_split % 16 || a.pad_mode != PAD_ZERO || a.pad_left != 0 || a.K != 1 ||
                     a.lens != nullptr || a.x_bs != (long long)a.cin_split * a.Lin || !tma_encode_fn()))
    return false;
  auto xr = [&](long long rows) { return (a.x_split || a.x1_split) ? (rows + 7) / 8 * 8 : rows; };   // 128-byte aligned TMA boxes
  {  // the kernels index inside one utterance's input / output plane with 32-bit element offsets
    const long long lim = 0x7fffffffLL - 65536;
    const long long out_plane = (a.out_layout == OUT_PHASE || a.out_layout == OUT_PHASE_SPLIT) ? (long long)a.ph_cout * a.ph_lout
                                                                                              : (long long)L.n_pad * a.Lpos;
    if (a.out_layout == OUT_PHASE_SPLIT && (a.ph_cout % 16 || !(a.out_slope >= 0.f && a.out_slope <= 1.f) || a.post_tanh)) return false;
    if ((long long)a.Cin * a.Lin >= lim || out_plane >= lim) return false;
  }
  struct Cand { int mt, a_st, res, w_st, ck, kbps, dual; double score; };
  Cand best{0, 0, 0, 0, 0, 0, 0, -1.0};
  Cand best2{0, 0, 0, 0, 0, 0, 0, -1.0};   // best candidate with TWO accumulator sets
  // Upsample layers (phase-interleaved outputs) are bound by their epilogue: with a single accumulator set the MMAs of the next
  // tile wait for it (issuer 62-86 % in acc_empty, epilogue 21-32 % in acc_full, profiles/r02_stall_hifigan_b32.txt), and the
  // cycle model underrates that.  Measured in one call (gpurun r2ab, FV_TC2_MT_XS=1): 256 -> 8 x 128: 0.205 -> 0.140 ms,
  // 64 -> 3 x 32: 0.277 -> 0.222 ms with one M tile and two sets -> those layouts take the best two-set plan when there is one.
  // The M-tile count does not change the accumulation order (ck does), so results are unchanged.  FV_TC2_UPS_ACC2=0: off.
  static const bool ups_acc2_env = getenv("FV_TC2_UPS_ACC2") == nullptr || atoi(getenv("FV_TC2_UPS_ACC2")) != 0;
  const bool prefer_acc2 = ups_acc2_env && (a.out_layout == OUT_PHASE || a.out_layout == OUT_PHASE_SPLIT);
  // Cycle model per CTA tile, calibrated on B200 (profiles/r01_notes.md, scripts/probes/umma_issue_probe.cu and the
  // FV_STALL_DEBUG role accounting):
  //   UMMA M=128 K=16: max(N/2, (4096 + 32 N)/128) clk — math vs the 128 B/clk shared-memory operand fetch; a tight
  //     issue loop costs ~20 clk per UMMA per issuer thread,
  //   loaders ~8 B/clk of fp32 input (DRAM-latency bound, 8 warps x 32 loads in flight),
  //   weight ring: bytes in flight / ~2500-cycle bulk-copy round trip, capped at ~14 B/clk per SM (all SMs stream the
  //     same image from L2) -> ring-mode layers want the largest M per weight pass,
  //   epilogue: 0.34 clk per output element, +0.47 with a residual read, +0.74 with the MRF read-modify-write
  //     (latency-bound global accesses of 4 warps); it overlaps the MMAs only with two accumulator sets.
  const int ck_opts[4] = {a.Cin, 128, 64, 32};
  // The K-chunk size fixes the ORDER in which the tensor core accumulates (chunk-major vs tap-major), i.e. the rounding of
  // every output.  It is therefore chosen from the layer shape alone (phase 0: canonical problem size) and only the
  // scheduling parameters (tile height, buffering, ring depth) follow the actual batch / length (phase 1) — a batched call
  // and per-utterance calls produce bit-identical samples whatever plans they get.
  int ck_fixed = 0;
  for (int phase = 0; phase < 2; ++phase) {
  const int pl_Lpos = phase == 0 ? (1 << 20) : a.Lpos;
  const int pl_B = phase == 0 ? 8 : a.B;
  const int need_mt = (pl_Lpos + 127) / 128;
  best = Cand{0, 0, 0, 0, 0, 0, 0, -1.0};
  best2 = Cand{0, 0, 0, 0, 0, 0, 0, -1.0};
  for (int df = 2; df >= 1; --df) {
    if (df == 2 && NT > 128) continue;
    if (force_dual >= 0 && NT <= 128 && (df == 2) != (force_dual != 0)) continue;
    // measured on B200 (history #12): [B_hi|B_lo] N-doubling wins for NT < 128 (fewer, longer UMMAs against the smem operand
    // floor); for NT = 128 the three-UMMA form costs the same tensor time and leaves room for two accumulator sets
    if (force_dual < 0 && (df == 2) != (NT < 128)) continue;
    const double c_pipe = df == 2
        ? std::max((double)NT, (4096.0 + NT * 64.0) / 128.0) + std::max(NT / 2.0, (4096.0 + NT * 32.0) / 128.0)
        : 3.0 * std::max(NT / 2.0, (4096.0 + NT * 32.0) / 128.0);
    const double n_umma = df == 2 ? 2.0 : 3.0;
    for (int cki = 0; cki < 4; ++cki) {
      const int ck = ck_opts[cki];
      if (ck > a.Cin || a.Cin % ck || ck % 16 || (cki > 0 && ck == a.Cin)) continue;
      if (phase == 1 && ck != ck_fixed) continue;
      if (a.x1_split && a.cin_split % ck) continue;   // a TMA-fed chunk must not straddle the two inputs
      const int nck = a.Cin / ck, kpc = ck / 16;
      int kbps = 16384 / kblock_bytes;
      if (kbps < 1) kbps = 1;
      if (kbps > kpc) kbps = kpc;
      const int stage_bytes = kbps * kblock_bytes;
      for (int res = 1; res >= 0; --res)
        for (int a_st = 2; a_st >= 1; --a_st)
          for (int mt = 8; mt >= 1; --mt) {
            if (mt * NT * df > 512) continue;
            if (force_mt_xs > 0 && a.x_split && phase == 1 && mt != std::min(force_mt_xs, std::min(need_mt, 512 / (NT * df)))) continue;
            if (mt > need_mt && mt > 1) continue;
            if (nck > 1 && a_st < 2) continue;    // chunking only pays with double-buffered stages
            for (int w_st = (res ? 1 : 8); w_st >= (res ? 1 : 2); --w_st) {
              const long long wb = res ? w_total : (long long)w_st * stage_bytes;
              const long long a_stage = 2LL * xr(mt * 128 + halo) * ck * 2;
              if (a_st * a_stage + wb + 512 > BUDGET) continue;
              const bool acc2 = 2 * mt * NT * df <= 512;
              int ni = res ? TC2_ISSUE_WARPS : TC2_ISSUE_WARPS - 1;
              if (ni > mt) ni = mt;
              if (NT * df >= 256 && ni > 2) ni = 2;
              const int my_mts = (mt + ni - 1) / ni;
              if (my_mts > 4) continue;
              const double t_mma = std::max((double)kblocks * mt * c_pipe, (double)kblocks * my_mts * n_umma * 20.0);
              const int rows_t = mt * 128 + halo;
              const int pairs = (ck / 8) * ((rows_t + 127) / 128);
              double t_load = a.x_split ? nck * (double)rows_t * ck * 4.0 / 40.0   // TMA: bandwidth only, no loader rounds
                                        : nck * (((pairs + 7) / 8) * 2500.0 + (double)rows_t * ck * 4.0 / 40.0);
              if (a.x1_split) t_load *= (double)(a.Cin - a.cin_split) / a.Cin;   // only the second input goes through the loaders
              const double ring_bw = std::min(14.0, (double)wb / 2500.0);
              const double t_w = res ? 0.0 : (double)w_total / ring_bw;
              const double t_epi = (double)mt * 128 * NT *
                                       (0.34 + (a.res != nullptr ? 0.47 : 0.0) + (acc_reads_y(a.acc_mode) ? 0.74 : 0.0)) +
                                   600.0;
              const double t_core = std::max(t_mma, t_w);
              double t_tile = (a_st == 2) ? std::max(t_core, t_load) : (t_core + t_load);
              t_tile = acc2 ? std::max(t_tile, t_epi) : (t_tile + t_epi);
              t_tile += 3000.0;   // measured fixed cost per tile (barrier hand-offs, pipeline bubbles)
              const long long tiles = (long long)((pl_Lpos + mt * 128 - 1) / (mt * 128)) * pl_B;
              long long gx = std::max(1, num_sms / L.n_tiles);
              if (gx > tiles) gx = tiles;
              const long long waves = (tiles + gx - 1) / gx;
              const double tail_eff = (double)tiles / (double)(waves * gx);
              const double useful = std::min<double>(mt * 128.0, (double)pl_Lpos);
              const double sc = useful / t_tile * tail_eff;
              if (sc > best.score * 1.02) best = Cand{mt, a_st, res, w_st, ck, kbps, df, sc};
              if (acc2 && sc > best2.score * 1.02) best2 = Cand{mt, a_st, res, w_st, ck, kbps, df, sc};
            }
          }
    }
  }
  if (best.score < 0) return false;
  if (prefer_acc2 && best2.score > 0) best = best2;
  ck_fixed = best.ck;
  }   // phase
  const int dualf = best.dual;
  p.a = a;
  p.NT = NT;
  p.m_tiles = best.mt;
  p.rows = best.mt * 128 + halo;
  p.rows_alloc = (int)xr(p.rows);
  static const int epi2_env = getenv("FV_TC2_EPI") ? atoi(getenv("FV_TC2_EPI")) : 2;
  p.epi_groups = (a.x_split && epi2_env >= 2) ? 2 : 1;
  p.ksteps = ksteps;
  p.kblocks = kblocks;
  p.a_stages = best.a_st;
  p.w_resident = best.res;
  p.kb_per_stage = best.kbps;
  p.w_stages = best.w_st;
  p.stage_bytes = best.kbps * kblock_bytes;
  p.ck = best.ck;
  p.nck = a.Cin / best.ck;
  {  // issuers: each owns the M tiles wid, wid + n, ... (its own accumulators)
    int ni = best.res ? TC2_ISSUE_WARPS : TC2_ISSUE_WARPS - 1;
    if (ni > best.mt) ni = best.mt;
    if (NT * dualf >= 256 && ni > 2) ni = 2;   // long UMMAs: the tensor pipe, not the issue slot, is the limit
    p.n_issuers = ni < 1 ? 1 : ni;
  }
  p.dual = dualf == 2 ? 1 : 0;
  // A-collector reuse (three-UMMA form): needs the fill / lastuse pair adjacent in the tensor pipe -> one issuer, which
  // is enough when every UMMA runs >= 64 clk (NT >= 128) against ~20 clk of issue cost
  static const int reuse_env = getenv("FV_A_REUSE") ? atoi(getenv("FV_A_REUSE")) : 0;
  p.a_reuse = (reuse_env && !p.dual && NT >= 128 && best.mt <= 4) ? 1 : 0;
  if (p.a_reuse) p.n_issuers = 1;
  static const int flat_env = getenv("FV_LOADER_FLAT") ? atoi(getenv("FV_LOADER_FLAT")) : 1;
  p.ld_per = p.ld_rounds = 0;
  // conv_tc2: measured mixed (C=64/128 k=11 -7 %, Cin=512 pair layers +9 %) -> opt-in with FV_LOADER_FLAT=2
  if (flat_env >= 2) flat_plan((best.ck / 8) * (best.mt * 128 + halo), p.ld_per, p.ld_rounds);
  p.acc_cols = best.mt * NT * dualf;
  p.acc_stages = (2 * p.acc_cols <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.acc_stages * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  // measured (gpurun r2r): the issuers drift apart by half an epilogue and hold the shared weight-ring slots longer — Basis-MelGAN
  // 11.2 -> 11.8-12.0 ms, MelGAN 16.9 -> 17.3 -> opt-in only
  static const int per_mt_env = getenv("FV_TC2_ACC_PER_MT") ? atoi(getenv("FV_TC2_ACC_PER_MT")) : 0;
  p.acc_per_mt = (per_mt_env && p.acc_stages == 1 && p.m_tiles == p.n_issuers && p.m_tiles <= 4 && p.m_tiles > 1) ? 1 : 0;
  p.idesc = make_idesc_f16(128, NT);
  p.idesc2 = make_idesc_f16(128, 2 * NT);
  p.tiles_per_batch = (a.Lpos + best.mt * 128 - 1) / (best.mt * 128);
  p.total_tiles = p.tiles_per_batch * a.B;
  return true;
}

inline void tc_apply_env_once() {
  static std::once_flag once;
  std::call_once(once, []() {
    const char* e = getenv("FV_WAIT_HINT");
    if (e) {
      int v = atoi(e);
      cudaMemcpyToSymbol(g_wait_hint, &v, sizeof(int));
    }
  });
}

// FV_PDL=1: launch the tensor-core kernels of the layer chain with programmatic stream serialization (the next
// layer's prologue overlaps this layer's tail).  Correct (GPU tests pass with it) but measured neutral-to-slower on the
// B=32 step (19.7 -> 20.1 ms, profiles/r01_notes.md "ab2"), so it stays opt-in.
// Default (FV_PDL unset): on for LATENCY-bound launches only — few tiles per CTA, i.e. small batches / the batch-1 path, where
// the next kernel's prologue (barrier init, TMEM alloc, resident weight images) is a visible share of a ~20-30 us launch
// (measured, T = 1000 batch 1: HiFi-GAN 1.52 -> 1.42 ms eager, 1.45 -> 1.40 ms graph replay; profiles/r02_notes.md).
inline bool tc_pdl_enabled(long long total_tiles = 0, int num_sms = 148) {
  static const int mode = getenv("FV_PDL") ? atoi(getenv("FV_PDL")) : -1;
  if (mode >= 0) return mode != 0;
  return total_tiles > 0 && total_tiles <= 4LL * num_sms;
}

// FV_STALL_DEBUG=1: launch the instrumented instantiation, synchronise and print where each role waited (stderr).
inline bool tc_stall_debug() {
  static const bool on = getenv("FV_STALL_DEBUG") != nullptr && atoi(getenv("FV_STALL_DEBUG")) != 0;
  return on;
}
struct StallReport {
  long long* dev = nullptr;
  int ctas = 0;
  bool begin(int n_ctas) {
    ctas = n_ctas;
    if (cudaMalloc(&dev, (size_t)ctas * 64 * sizeof(long long)) != cudaSuccess) return false;
    return cudaMemset(dev, 0, (size_t)ctas * 64 * sizeof(long long)) == cudaSuccess;
  }
  // role_names[r] / slot_names[r][k]: nullptr-terminated labels
  void finish(cudaStream_t st, const char* title, const char* const* role_names, const char* const (*slot_names)[5]) {
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)ctas * 64);
    cudaMemcpy(h.data(), dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    dev = nullptr;
    fprintf(stderr, "[stall] %s ctas=%d\n", title, ctas);
    for (int r = 0; r < 8; ++r) {
      if (!role_names[r]) continue;
      double tot = 0, tiles = 0, w[5] = {0, 0, 0, 0, 0};
      int n = 0;
      for (int c = 0; c < ctas; ++c) {
        const long long* o = &h[((size_t)c * 8 + r) * 8];
        if (o[5] == 0) continue;
        ++n; tot += (double)o[5]; tiles += (double)o[6];
        for (int k = 0; k < 5; ++k) w[k] += (double)o[k];
      }
      if (!n) continue;
      fprintf(stderr, "[stall]   %-9s total %9.0f clk  units %6.1f", role_names[r], tot / n, tiles / n);
      double ws = 0;
      for (int k = 0; k < 5; ++k) {
        if (!slot_names[r][k]) continue;
        fprintf(stderr, "  %s %5.1f%%", slot_names[r][k], 100.0 * w[k] / tot);
        ws += w[k];
      }
      fprintf(stderr, "  busy %5.1f%% (%.0f clk/unit)\n", 100.0 * (tot - ws) / tot, tiles > 0 ? (tot - ws) / tiles : 0.0);
    }
  }
};

inline int launch_conv_tc2(const ConvArgs& a, const TcLayer& L, cudaStream_t st) {
  tc_apply_env_once();
  static std::mutex mu;          // per-device launch attributes: concurrent fv_forward calls on distinct streams are allowed
  static int num_sms[64] = {};
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!num_sms[dev]) {
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
      num_sms[dev] = prop.multiProcessorCount;
    }
    if (!attr_set[dev]) {
      const int mx = 227 * 1024;
      if (cudaFuncSetAttribute(conv_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess)
        return -1;
      attr_set[dev] = true;
    }
  }
  Tc2Args p{};
  if (!L.eligible || !L.image || !tc2_plan(a, L, p, num_sms[dev])) return 1;
  p.wimg = L.image;
  const size_t smem = tc2_smem_bytes(p);
  CUtensorMap tm_main, tm_tail;
  memset(&tm_main, 0, sizeof tm_main);
  memset(&tm_tail, 0, sizeof tm_tail);
  if (a.x_split || a.x1_split) {
    const long long planes = (long long)a.B * 2 * ((a.x1_split ? a.cin_split : a.Cin) / 8);
    const int tail = p.rows % TMA_SPLIT_RB;
    if (!tma_encode_split(&tm_main, a.x, a.Lin, planes, TMA_SPLIT_RB)) return 1;
    if (tail) {
      if (!tma_encode_split(&tm_tail, a.x, a.Lin, planes, tail)) return 1;
    } else {
      tm_tail = tm_main;
    }
  }
  int gx = num_sms[dev] / L.n_tiles;
  if (gx < 1) gx = 1;
  if (gx > p.total_tiles) gx = p.total_tiles;
  // EXPERIMENTAL (round 1): the multicast ring produces wrong results on hardware, so it is opt-in until debugged.
  static const int cluster_mode = getenv("FV_CLUSTER") ? atoi(getenv("FV_CLUSTER")) : 0;
  const int cs = (!p.w_resident && cluster_mode > 0 && !a.x_split && num_sms[dev] / L.n_tiles >= 2) ? 2 : 1;
  p.cluster_mode = cs == 2 ? cluster_mode : 0;
  if (cs == 2) gx = (gx + 1) & ~1;   // pairs; an odd tile count leaves one CTA with ring duty only
  if (cs == 2 && gx > num_sms[dev] / L.n_tiles) gx -= 2;
  if (gx < cs) gx = cs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(gx, L.n_tiles, 1);
  cfg.blockDim = dim3(TC2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  p.pdl = (tc_pdl_enabled(p.total_tiles, num_sms[dev]) && cs == 1 && !tc_stall_debug()) ? 1 : 0;
  if (p.pdl) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t le;
  if (tc_stall_debug()) {
    StallReport rep;
    if (!rep.begin(gx * L.n_tiles)) return -1;
    p.dbg = rep.dev;
    le = a.x_split ? cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, true>, p, tm_main, tm_tail)
                   : cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, false>, p, tm_main, tm_tail);
    static const char* const roles[8] = {"loader", "producer", "issuer0", "epilogue", nullptr, nullptr, nullptr, nullptr};
    static const char* const slots[8][5] = {{"a_empty", nullptr, nullptr, nullptr, nullptr},
                                            {"w_empty", "w_free", nullptr, nullptr, nullptr},
                                            {"w_res", "acc_empty", "a_full", "w_full", "w_drain"},
                                            {"acc_full", "tmem", "addwait", "ldissue", "store"}};
    char title[256];
    snprintf(title, sizeof title,
             "tc2 Cin=%d N=%d K=%d dil=%d Lpos=%d B=%d layout=%d res=%d acc=%d | NT=%d mt=%d ck=%d a_st=%d acc_st=%d "
             "resident=%d w_st=%d kbps=%d issuers=%d dual=%d tiles=%d grid=%dx%d",
             a.Cin, a.N, a.K, a.dil, a.Lpos, a.B, a.out_layout, a.res != nullptr, a.acc_mode, p.NT, p.m_tiles, p.ck,
             p.a_stages, p.acc_stages, p.w_resident, p.w_stages, p.kb_per_stage, p.n_issuers, p.dual, p.total_tiles, gx,
             L.n_tiles);
    rep.finish(st, title, roles, slots);
  } else {
    le = a.x_split ? cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, true>, p, tm_main, tm_tail)
                   : cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, false>, p, tm_main, tm_tail);
  }
  g_launches++;
  g_tc_launches++;
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : -1;
}


// =================================================================================================
// v3: fused ResBlock1 unit (modules.py:224-229)
//        y = conv2( lrelu( conv1_dil(lrelu(x)) + b1 ) ) + b2 + x          [+ MRF accumulate]
// in ONE persistent kernel: the intermediate h never leaves the SM.  conv1's accumulators are read back by the
// epilogue warps, biased, activated, split to fp16 hi/lo and written straight into shared memory in the UMMA
// operand layout (A2); conv2 then runs on A2.  Per unit this moves 4C B in + 4C B out (+4C accumulate) per
// position instead of 4C·5 (+4C) for the two-kernel form — the narrow layers were DRAM-bound on exactly that.
// Both weight images stay resident in shared memory (C <= 64).  Tile: 128m rows of h -> M_out = 128m-(K-1)
// outputs (conv2's halo is recomputed), x rows = 128m + (K-1)*dil.
// =================================================================================================
constexpr int TC3_THREADS = (TC2_LOADER_WARPS + TC2_ISSUE_WARPS + 8) * 32;   // + 4 epiA warps + 4 epiB warps
// Split (TMA-fed) units have no loader warps: warp 0 is the TMA producer and warps 4-7 are a SECOND epiB group (a TMEM lane
// quarter is served by two warps that take alternate (chunk, M tile) iterations; epiB — residual loads, activation split,
// global stores — is the heavier epilogue: 80-96 % busy against 45-60 % for epiA on the k = 3 units).  A 21st warp would put
// six warps on one SM sub-partition and cap the kernel at 80 registers; taking an issuer warp for the TMA duty leaves an
// uneven M-tile split (profiles/r02_notes.md).

struct Tc3Args {
  const float* x;      // unit input [B, C, L] (also the residual)
  float* y;            // [B, C, L]
  const float* bias1;  // conv1 bias or nullptr
  const float* bias2;
  const int* lens;     // ragged batches: valid length per utterance (<= L), or nullptr
  const uint8_t* w1img;
  const uint8_t* w2img;
  int B, C, L, K, dil;
  float slope;         // LeakyReLU slope applied to x and to h (0.1)
  int acc_mode;
  float acc_div;
  int m_tiles, x_rows, h_rows_alloc, m_out;
  int a1_stages, acc1_stages, n_issuers;
  int acc_cols;        // columns of one accumulator set = m_tiles * 2C
  int tmem_cols;
  int ksteps, kblocks;
  int tiles_per_batch, total_tiles;
  uint32_t idesc, idesc2;
  int ld_per, ld_rounds;   // flattened loader (fill_stage_flat); 0 = legacy pair loop
  int pdl;                 // launched with programmatic stream serialization
  // Split ("TMA-native") activation I/O (template parameter IO of the kernel, fv_tma.cuh): the unit input arrives
  // pre-activated and pre-split as fp16 hi / lo planes xs[B][2][C/8][L] (uint4 rows) and is fetched by TMA straight into
  // the A1 stage (no loader warps, OOB rows zero-filled = the conv's zero padding); the residual is rebuilt from the
  // same planes (x = unlrelu(hi + lo)); IO_SPLIT_SPLIT units write their output in the same format for the next unit.
  const void* xs;
  void* ys;
  const float* ysum;       // IO_SPLIT_SPLIT only: fp32 running MRF sum added after the 1/num_kernels scaling (last branch)
  float inv_slope;         // 1 / slope (slope > 0): undoes the LeakyReLU baked into the split copy
  int x_rows_alloc;        // row stride of the A1 planes (x_rows rounded up to 8 in split mode: 128-B aligned TMA boxes)
  int epi_groups;          // split mode: 1 or 2 warps per TMEM lane quarter in each of epiA / epiB
  int pair_issue;          // resident, non-ping-pong plans with one M tile per issuer: conv1(i+1) and conv2(i) issued interleaved
  // Streamed weights (the two images do not fit next to the tiles: C = 64, k = 7 / 11): warp 11 feeds a ring of
  // `w_stages` slots, one slot = one tap (ksteps k-blocks = stage_bytes), in the order the issuers consume them
  // (conv1 of tile 0, then per tile conv2(i), conv1(i+1)).  Resident mode: w_resident = 1.
  int w_resident, w_stages, stage_bytes;
  int acc2_stages;         // accumulator sets of conv2 (2 unless TMEM is needed for larger tiles)
  // Ping-pong tiles: each tile in flight owns ONE A buffer and ONE accumulator set for both of its convs — epiA writes h
  // in place over the (dead) x tile and conv2 accumulates over the conv1 columns epiA has drained.  Two tiles in flight
  // with half the TMEM / no separate A2 buffer; the issuers walk pairs: conv1(a) conv1(b) conv2(a) conv2(b).
  int pp;
  // Fused ResidualStack (template parameter IO = IO_STACK, MelGAN family, modules.py:353-382): conv1 = the dilated k-tap conv on
  // lrelu(c) with reflect padding, conv2 = ONE tap over 2C channels [h | c] against the concatenated image [W_1x1 ; W_skip]
  // (the stack's 1x1 conv and its skip conv as one GEMM), no residual in epiB.  Stage layout per half (hi / lo):
  // [C/8 activated planes x x_rows_alloc rows (halo included)] [C/8 raw planes x m_out rows (row r = position t0 + r)] — both
  // written by the loaders from the same loaded registers; epiA writes h in place over the activated planes and conv2 is issued
  // as two descriptor walks (h planes, then the raw planes, accumulating) over the two halves of its image.  Always ping-pong
  // tiles with resident weights; h never leaves the SM and c is read once.
  int stack;
  int k2, ksteps2, kblocks2;   // conv2: taps, k-steps per tap, k-blocks of its image (HiFi units: K, ksteps, kblocks)
  int reflect;                 // loaders: ReflectionPad1d instead of zero padding (per utterance length)
  long long* dbg;      // stall-accounting buffer (FV_STALL_DEBUG) or nullptr
};

// One issuer's UMMAs of a whole K-tap conv on ONE M tile, k-steps unrolled (KS = Cin/16): per tap the loop body is
// 2*KS UTCHMMAs plus two uniform adds.  dual layout: [hi*hi | hi*lo] with N = 2*NT, then lo*hi with N = NT.
template <int KS>
__device__ __forceinline__ void issue_conv_1mt(uint64_t ad, uint64_t bd, uint32_t d, int K, uint64_t tap_step16,
                                               uint64_t ks_step16, uint64_t kb_step16, uint64_t lo_delta16,
                                               uint32_t idesc, uint32_t idesc2, uint32_t accum = 0u) {
  for (int j = 0; j < K; ++j, ad += tap_step16, bd += KS * kb_step16) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      umma_f16_elect(d, ad + ks * ks_step16, bd + ks * kb_step16, idesc2, accum);
      umma_f16_elect(d, ad + ks * ks_step16 + lo_delta16, bd + ks * kb_step16, idesc, 1u);
      accum = 1u;
    }
  }
}

enum { IO_F32 = 0, IO_SPLIT_SPLIT = 1, IO_SPLIT_F32 = 2, IO_STACK = 3 };

// conv1 of the NEXT tile and conv2 of the current one are independent (different A buffers, weights, accumulators): issued
// interleaved, tap by tap, an issuer that owns one M tile drives TWO accumulator chains instead of one, i.e. each dependent
// UMMA waits for half as long (k = 3 units are bound by exactly that chain latency).
template <int KS>
__device__ __forceinline__ void issue_conv_pair_1mt(uint64_t ad1, uint64_t bd1, uint32_t d1, uint64_t tap1_16, uint64_t ks1_16,
                                                    uint64_t lo1_16, uint64_t ad2, uint64_t bd2, uint32_t d2, uint64_t ks2_16,
                                                    uint64_t lo2_16, int K, uint64_t kb_step16, uint32_t idesc, uint32_t idesc2) {
  uint32_t accum = 0u;
  for (int j = 0; j < K; ++j, ad1 += tap1_16, ad2 += 1, bd1 += KS * kb_step16, bd2 += KS * kb_step16) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      umma_f16_elect(d1, ad1 + ks * ks1_16, bd1 + ks * kb_step16, idesc2, accum);
      umma_f16_elect(d2, ad2 + ks * ks2_16, bd2 + ks * kb_step16, idesc2, accum);
      umma_f16_elect(d1, ad1 + ks * ks1_16 + lo1_16, bd1 + ks * kb_step16, idesc, 1u);
      umma_f16_elect(d2, ad2 + ks * ks2_16 + lo2_16, bd2 + ks * kb_step16, idesc, 1u);
      accum = 1u;
    }
  }
}

template <bool DBG, int IO>
__global__ void __launch_bounds__(TC3_THREADS, 1) conv_tc3_fused_kernel(const Tc3Args p, const __grid_constant__ CUtensorMap tm_main,
                                                                        const __grid_constant__ CUtensorMap tm_tail) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  constexpr bool STK = (IO == IO_STACK);                              // fused ResidualStack (see Tc3Args::stack)
  constexpr bool SPLIT_IN = (IO == IO_SPLIT_SPLIT || IO == IO_SPLIT_F32);   // x tile fetched by TMA from a split copy
  const int C = p.C, NT = p.C;
  const uint32_t a1_bytes = (uint32_t)(p.x_rows_alloc + (STK ? p.m_out : 0)) * C * 2;  // hi or lo of one A1 stage (STK: + raw planes)
  const uint32_t a2_bytes = (uint32_t)p.h_rows_alloc * C * 2;  // hi or lo of A2
  const int kblock_bytes = NT * 64;
  const uint32_t w_bytes = (uint32_t)p.kblocks * kblock_bytes;
  const uint32_t w2_bytes = (uint32_t)p.kblocks2 * kblock_bytes;
  uint8_t* A1 = smem;                                            // [a1_stages][hi|lo]
  uint8_t* A2 = A1 + (size_t)p.a1_stages * 2 * a1_bytes;        // [hi|lo]
  uint8_t* W1 = A2 + (p.pp ? 0 : 2 * (size_t)a2_bytes);         // resident: [W1 | W2]; ring: w_stages slots (pp: no A2 buffer)
  uint8_t* W2 = W1 + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(W1 + (p.w_resident ? (size_t)w_bytes + w2_bytes : (size_t)p.w_stages * p.stage_bytes));
  // [0,2) a1_full [2,4) a1_empty [4,6) acc1_full [6,8) acc1_empty  8 a2_full  9 a2_empty  [10,12) acc2_full  [12,14) acc2_empty  14 w_full
  // [15,19) ring slot full  [19,23) ring slot empty
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int p2 = (p.k2 - 1) / 2, p1 = (p.K - 1) * p.dil / 2;
  const int epi_groups = (SPLIT_IN && p.epi_groups == 2) ? 2 : 1;
  // epilogue roles: group 0 = warps 12-15 (epiA) / 16-19 (epiB); split mode with two groups: + warps 0-3 / 4-7
  const bool is_epiA = (warp >= 12 && warp < 16);
  const bool is_epiB = (warp >= 16 && warp < 20) || (epi_groups == 2 && warp >= 4 && warp < 8);
  const int epi_grp = warp < TC2_LOADER_WARPS ? 1 : 0;
  const int epiA_groups = 1, epiB_groups = epi_groups;
  if (p.pdl) pdl_launch_dependents();

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(0 + s), SPLIT_IN ? 1 : TC2_LOADER_WARPS);   // split mode: one expect_tx arrive + the TMA bytes
      mbar_init(BAR(2 + s), p.n_issuers);
      mbar_init(BAR(4 + s), p.n_issuers);
      mbar_init(BAR(6 + s), 4 * epiA_groups);
    }
    mbar_init(BAR(8), 4 * epiA_groups);
    mbar_init(BAR(9), p.pp ? 4 * epiA_groups : p.n_issuers);   // pp: h_full of the odd buffer; else a2_empty
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(10 + s), p.n_issuers);
      mbar_init(BAR(12 + s), 4 * epiB_groups);
    }
    mbar_init(BAR(14), 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(BAR(15 + s), 1);              // ring slot full: expect_tx by the producer
      mbar_init(BAR(19 + s), p.n_issuers);    // ring slot empty: one tcgen05.commit per issuer
    }
    fence_mbar_init();
  }
  if (warp == TC2_LOADER_WARPS) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc2_base = tmem_base + (uint32_t)((p.pp ? 0 : p.acc1_stages) * p.acc_cols);   // the acc2 set(s) follow (pp: aliased)
  // single-buffered A1: conv2(i) is issued before conv1(i+1) so that it never waits behind the load of the next tile
  const bool conv2_first = p.a1_stages == 1;

  // ------------------------------------------------------------------ TMA producer loop (one warp, see TC3 role notes)
  auto tma_producer = [&]() {
      const int rows = p.x_rows, nkc = C >> 3;
      const int nfull = rows / TMA_SPLIT_RB, tail = rows - nfull * TMA_SPLIT_RB;
      const int nrb = nfull + (tail ? 1 : 0);
      const int per_half = nkc * nrb, nops = 2 * per_half;
      const uint32_t tile_bytes = (uint32_t)(2 * nkc * rows * 16);
      if (lane == 0) { tma_prefetch_desc(&tm_main); tma_prefetch_desc(&tm_tail); }
      int it = 0;
      WaitAcc<DBG> wa;
      wa.begin();
      if (p.pdl) pdl_wait();
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it % p.a1_stages;
        if (it >= p.a1_stages) wa.wait(0, BAR(2 + s), (uint32_t)((it / p.a1_stages - 1) & 1), 800 + s);
        const int b = tile / p.tiles_per_batch;
        const int g0 = (tile - b * p.tiles_per_batch) * p.m_out - p2 - p1;
        if (lane == 0) mbar_expect_tx(BAR(0 + s), tile_bytes);
        __syncwarp();
        const uint32_t a_hi = smem_u32(A1 + (size_t)s * 2 * a1_bytes);
        for (int i = lane; i < nops; i += 32) {
          const int half = i >= per_half ? 1 : 0;
          const int rem = i - half * per_half;
          const int kc = rem / nrb, rb = rem - kc * nrb;
          const uint32_t dst = a_hi + (uint32_t)half * a1_bytes + (uint32_t)(kc * p.x_rows_alloc + rb * TMA_SPLIT_RB) * 16u;
          tma_load_2d(dst, rb < nfull ? &tm_main : &tm_tail, 2 * (g0 + rb * TMA_SPLIT_RB), (b * 2 + half) * nkc + kc, BAR(0 + s));
        }
      }
      wa.end(p.dbg, 0, lane == 0, it);
  };
  if (SPLIT_IN && warp == 0) {
    tma_producer();
  } else if (!SPLIT_IN && warp < TC2_LOADER_WARPS) {
    // ------------------------------------------------------------------ loaders: x tile -> A1[stage]
    const int rows = p.x_rows;
    const int nkc = C >> 3;
    const int nrb = (rows + 127) >> 7;
    const int npairs = nkc * nrb;
    int it = 0;
    WaitAcc<DBG> wa;
    wa.begin();
    if (p.pdl) pdl_wait();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it % p.a1_stages;
      if (it >= p.a1_stages) wa.wait(0, BAR(2 + s), (uint32_t)((it / p.a1_stages - 1) & 1), 800 + s);
      const int b = tile / p.tiles_per_batch;
      const int t0 = (tile - b * p.tiles_per_batch) * p.m_out;
      const float* __restrict__ xb = p.x + (long long)b * C * p.L;
      const int Lb = p.lens ? __ldg(p.lens + b) : p.L;
      uint8_t* A_hi = A1 + (size_t)s * 2 * a1_bytes;
      uint8_t* A_lo = A_hi + a1_bytes;
      if (STK) {   // activated tile (reflect padding) + the un-activated central rows behind it (conv2's skip operand)
        // (central rows beyond a ragged utterance's end hold reflected samples in both copies: they only feed outputs >= Lb)
        const float slope1 = p.slope;
        const uint32_t raw_off = (uint32_t)nkc * (uint32_t)p.x_rows_alloc * 16u;
        fill_stage_flat<true, true>(A_hi, A_lo, rows, nkc, warp * 32 + lane, p.ld_per, p.ld_rounds, t0 - p2 - p1, Lb, p.L,
                        p.reflect != 0, [&](int kc, const float*& xc, float& slope) {
                          xc = xb + (long long)(kc * 8) * p.L;
                          slope = slope1;
                        }, p.x_rows_alloc, A_hi + raw_off, A_lo + raw_off, p1, p.m_out);
      } else if (p.ld_per > 0) {
        const float slope1 = p.slope;
        fill_stage_flat<true>(A_hi, A_lo, rows, nkc, warp * 32 + lane, p.ld_per, p.ld_rounds, t0 - p2 - p1, Lb, p.L,
                        false, [&](int kc, const float*& xc, float& slope) {
                          xc = xb + (long long)(kc * 8) * p.L;
                          slope = slope1;
                        });
      } else
      for (int pr = warp; pr < npairs; pr += TC2_LOADER_WARPS) {
        const int kc = pr / nrb, rbk = pr - kc * nrb;
        const float* __restrict__ xc = xb + (long long)(kc * 8) * p.L;
        float v[4][8];
        int rrow[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int r = rbk * 128 + t * 32 + lane;
          rrow[t] = r;
          const int g = t0 - p2 - p1 + r;
          const bool ok = r < rows && g >= 0 && g < Lb;
          const float* pg = opaque_ptr(xc + (ok ? g : 0));
#pragma unroll
          for (int c = 0; c < 8; ++c) v[t][c] = ok ? __ldg(pg + (uint32_t)c * (uint32_t)p.L) : 0.f;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (rrow[t] >= rows) continue;
          uint32_t hp[4], lp[4];
#pragma unroll
          for (int c = 0; c < 8; c += 2)
            split_f16x2(lrelu01(v[t][c], p.slope), lrelu01(v[t][c + 1], p.slope), hp[c >> 1], lp[c >> 1]);
          const uint32_t off = ((uint32_t)kc * rows + rrow[t]) * 16;
          *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(0 + s));
      // epiB of this tile (two tiles from now) adds the running MRF sum y: a DRAM read its 4 warps cannot hide.
      // Pull those lines into L2 now (x itself was just read by this loop, so the residual is an L2 hit already).
      if (acc_reads_y(p.acc_mode)) {
        const int lines_per_row = (p.m_out * 4 + 127) >> 7;
        const int total = C * lines_per_row;
        const float* __restrict__ yb = p.y + (long long)b * C * p.L;
        for (int i = warp * 32 + lane; i < total; i += TC2_LOADER_WARPS * 32) {
          const int n = i / lines_per_row, ln = i - n * lines_per_row;
          const long long pos = (long long)t0 + ln * 32;
          if (pos < p.L) asm volatile("prefetch.global.L2 [%0];" ::"l"(yb + (long long)n * p.L + pos));
        }
      }
    }
    wa.end(p.dbg, 0, warp == 0 && lane == 0, it);
  } else if (warp >= TC2_LOADER_WARPS && warp < TC2_LOADER_WARPS + TC2_ISSUE_WARPS) {
    // ------------------------------------------------------------------ weights (once) + UMMA issuers
    const int wid = warp - TC2_LOADER_WARPS;
    const bool L0 = (lane == 0);
    {
      if (wid == TC2_ISSUE_WARPS - 1 && p.w_resident) {
        if (L0) {
          mbar_expect_tx(BAR(14), w_bytes + w2_bytes);
          for (uint32_t off = 0; off < w_bytes; off += 32768)
            bulk_g2s(smem_u32(W1 + off), p.w1img + off, min(32768u, w_bytes - off), BAR(14));
          for (uint32_t off = 0; off < w2_bytes; off += 32768)
            bulk_g2s(smem_u32(W2 + off), p.w2img + off, min(32768u, w2_bytes - off), BAR(14));
        }
      } else if (wid == TC2_ISSUE_WARPS - 1) {
        // ring producer (n_issuers <= 3 in ring mode): one slot per tap, in the issuers' consumption order
        WaitAcc<DBG> wa;
        wa.begin();
        int n_my = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) ++n_my;
        int g = 0;
        auto stream_conv = [&](const uint8_t* img) {
          for (int j = 0; j < p.K; ++j, ++g) {
            const int slot = g % p.w_stages;
            if (g >= p.w_stages) wa.wait(0, BAR(19 + slot), (uint32_t)((g / p.w_stages - 1) & 1), 890 + slot);
            if (L0) {
              mbar_expect_tx(BAR(15 + slot), (uint32_t)p.stage_bytes);
              bulk_g2s(smem_u32(W1 + (size_t)slot * p.stage_bytes), img + (size_t)j * p.stage_bytes, (uint32_t)p.stage_bytes,
                       BAR(15 + slot));
            }
          }
        };
        if (p.pp) {
          for (int i = 0; i < n_my; i += 2) {
            stream_conv(p.w1img);
            if (i + 1 < n_my) stream_conv(p.w1img);
            stream_conv(p.w2img);
            if (i + 1 < n_my) stream_conv(p.w2img);
          }
        } else if (n_my > 0) stream_conv(p.w1img);
        for (int i = 0; i < n_my && !p.pp; ++i) {
          if (conv2_first) {
            stream_conv(p.w2img);
            if (i + 1 < n_my) stream_conv(p.w1img);
          } else {
            if (i + 1 < n_my) stream_conv(p.w1img);
            stream_conv(p.w2img);
          }
        }
        wa.end(p.dbg, 1, L0, n_my);
      }
      if (wid < p.n_issuers) {
        WaitAcc<DBG> wa;
        wa.begin();
        if (p.w_resident) wa.wait(0, BAR(14), 0, 810);
        int gw = 0;                                   // ring mode: taps consumed so far
        const uint32_t b_lbo = (uint32_t)NT * 32;
        const uint64_t b_tmpl = make_kmajor_desc(0, b_lbo, 128);
        const uint64_t a1_tmpl = make_kmajor_desc(0, (uint32_t)p.x_rows_alloc * 16, 128);
        const uint64_t a2_tmpl = make_kmajor_desc(0, (uint32_t)p.h_rows_alloc * 16, 128);
        const uint32_t mt_cols = (uint32_t)(2 * NT);
        const uint32_t mt_step16 = 128u * (uint32_t)p.n_issuers;
        const uint32_t d_step = mt_cols * (uint32_t)p.n_issuers;
        const uint32_t w1s = smem_u32(W1), w2s = smem_u32(W2);
        // one GEMM-conv over a resident weight image: A rows [row0 + j*dil], accumulators at `acc`
        // Tight issue loop: all descriptor arithmetic is incremental and warp-uniform (uniform registers), the whole
        // warp walks it converged and one elected lane issues (umma_f16_elect).
        const uint32_t idesc = p.idesc, idesc2 = p.idesc2;
        const int K = p.K, ksteps = p.ksteps, m_tiles = p.m_tiles, n_iss = p.n_issuers;
        const uint64_t kb_step16 = (uint64_t)(kblock_bytes >> 4);
        auto run_conv = [&](uint64_t a_tmpl, uint32_t a_hi_addr, uint32_t a_rows, uint32_t a_lo_delta16, uint32_t wsm,
                            int dil, uint32_t acc, int K, int ksteps, uint32_t accum0 = 0u) {   // K taps of `ksteps` k-blocks (shadow the unit's)
          const uint64_t ad0 = a_tmpl + (uint64_t)((a_hi_addr >> 4) & 0x3FFF) + (uint64_t)(wid * 128);
          uint64_t bd = b_tmpl + (uint64_t)((wsm >> 4) & 0x3FFF);
          const uint32_t d0 = acc + (uint32_t)wid * mt_cols;
          const uint64_t ks_step16 = (uint64_t)(2u * a_rows);
          const uint64_t lo_delta = (uint64_t)a_lo_delta16;
          if (!p.w_resident) {      // streamed weights: tap j of this conv sits in ring slot gw % w_stages
            uint32_t accum = 0u;
            for (int j = 0; j < K; ++j, ++gw) {
              const int slot = gw % p.w_stages;
              wa.wait(0, BAR(15 + slot), (uint32_t)((gw / p.w_stages) & 1), 895 + slot);
              tc_fence_after();
              uint64_t bdj = b_tmpl + (uint64_t)(((w1s + (uint32_t)slot * (uint32_t)p.stage_bytes) >> 4) & 0x3FFF);
              uint64_t ad = ad0 + (uint64_t)(j * dil);
              if (m_tiles <= n_iss) {   // one M tile per issuer: the tight two-UMMA body (issue cost is on the chain's critical path)
                for (int ks = 0; ks < ksteps; ++ks, ad += ks_step16, bdj += kb_step16) {
                  umma_f16_elect(d0, ad, bdj, idesc2, accum);
                  umma_f16_elect(d0, ad + lo_delta, bdj, idesc, 1u);
                  accum = 1u;
                }
              } else
              for (int ks = 0; ks < ksteps; ++ks, ad += ks_step16, bdj += kb_step16) {
                uint64_t a = ad;
                uint32_t d = d0;
                for (int mt = wid; mt < m_tiles; mt += n_iss, a += mt_step16, d += d_step) umma_f16_elect(d, a, bdj, idesc2, accum);
                a = ad + lo_delta; d = d0;   // product-major: independent accumulators back to back (see issue_kblocks)
                for (int mt = wid; mt < m_tiles; mt += n_iss, a += mt_step16, d += d_step) umma_f16_elect(d, a, bdj, idesc, 1u);
                accum = 1u;
              }
              umma_commit_elect(BAR(19 + slot));      // the slot may be refilled once these UMMAs have read it
            }
            return;
          }
          if (m_tiles <= n_iss) {   // one M tile per issuer (every plan tc3_plan makes): k-steps unrolled
            if (ksteps == 1) { issue_conv_1mt<1>(ad0, bd, d0, K, (uint64_t)dil, ks_step16, kb_step16, lo_delta, idesc, idesc2, accum0); return; }
            if (ksteps == 2) { issue_conv_1mt<2>(ad0, bd, d0, K, (uint64_t)dil, ks_step16, kb_step16, lo_delta, idesc, idesc2, accum0); return; }
            if (ksteps == 4) { issue_conv_1mt<4>(ad0, bd, d0, K, (uint64_t)dil, ks_step16, kb_step16, lo_delta, idesc, idesc2, accum0); return; }
          }
          uint32_t accum = accum0;
          for (int j = 0; j < K; ++j) {
            uint64_t ad = ad0 + (uint64_t)(j * dil);
            for (int ks = 0; ks < ksteps; ++ks, ad += ks_step16, bd += kb_step16) {
              uint64_t a = ad;
              uint32_t d = d0;
              for (int mt = wid; mt < m_tiles; mt += n_iss, a += mt_step16, d += d_step) umma_f16_elect(d, a, bd, idesc2, accum);
              a = ad + lo_delta; d = d0;
              for (int mt = wid; mt < m_tiles; mt += n_iss, a += mt_step16, d += d_step) umma_f16_elect(d, a, bd, idesc, 1u);
              accum = 1u;
            }
          }
        };
        int n_my = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) ++n_my;
        auto conv1 = [&](int i) {
          const int s = i % p.a1_stages, as = i % p.acc1_stages;
          wa.wait(1, BAR(0 + s), (uint32_t)((i / p.a1_stages) & 1), 820 + s);
          if (p.pp) {   // the set is free once epiB of the tile two back has drained it
            if (i >= 2) wa.wait(2, BAR(12 + as), (uint32_t)((i / 2 - 1) & 1), 830 + as);
          } else if (i >= p.acc1_stages) {
            wa.wait(2, BAR(6 + as), (uint32_t)((i / p.acc1_stages - 1) & 1), 830 + as);
          }
          tc_fence_after();
          run_conv(a1_tmpl, smem_u32(A1 + (size_t)s * 2 * a1_bytes), (uint32_t)p.x_rows_alloc, a1_bytes >> 4, w1s, p.dil,
                   tmem_base + (uint32_t)(as * p.acc_cols), K, ksteps);
          if (!p.pp) umma_commit_elect(BAR(2 + s));   // pp: the buffer stays busy (h in place) until conv2 is done
          umma_commit_elect(BAR(4 + as));
        };
        auto conv2 = [&](int i) {
          const int bs = i % p.acc2_stages;   // acc2 double buffered (when TMEM allows): epiB(i-1) overlaps conv2(i)
          if (p.pp) {   // h sits in the tile's own buffer (row stride x_rows), the accumulators are the drained conv1 set
            wa.wait(3, BAR(8 + bs), (uint32_t)((i / 2) & 1), 840 + bs);
            tc_fence_after();
            if (STK) {   // one tap over 2C channels: h (in place over the x planes), then the raw c planes behind them
              const uint32_t stage = smem_u32(A1 + (size_t)bs * 2 * a1_bytes);
              const uint32_t acc = tmem_base + (uint32_t)(bs * p.acc_cols);
              run_conv(a1_tmpl, stage, (uint32_t)p.x_rows_alloc, a1_bytes >> 4, w2s, 1, acc, 1, ksteps);
              run_conv(make_kmajor_desc(0, (uint32_t)p.m_out * 16, 128), stage + (uint32_t)(C >> 3) * (uint32_t)p.x_rows_alloc * 16u,
                       (uint32_t)p.m_out, a1_bytes >> 4, w2s + (uint32_t)ksteps * (uint32_t)kblock_bytes, 1, acc, 1, ksteps, 1u);
            } else
            run_conv(a1_tmpl, smem_u32(A1 + (size_t)bs * 2 * a1_bytes), (uint32_t)p.x_rows_alloc, a1_bytes >> 4, w2s, 1,
                     tmem_base + (uint32_t)(bs * p.acc_cols), K, ksteps);
            umma_commit_elect(BAR(2 + bs));   // buffer free for the loader
            umma_commit_elect(BAR(10 + bs));
            return;
          }
          wa.wait(3, BAR(8), (uint32_t)(i & 1), 840);
          if (i >= p.acc2_stages) wa.wait(4, BAR(12 + bs), (uint32_t)((i / p.acc2_stages - 1) & 1), 850 + bs);
          tc_fence_after();
          run_conv(a2_tmpl, smem_u32(A2), (uint32_t)p.h_rows_alloc, a2_bytes >> 4, w2s, 1,
                   acc2_base + (uint32_t)(bs * p.acc_cols), K, ksteps);
          umma_commit_elect(BAR(9));
          umma_commit_elect(BAR(10 + bs));
        };
        // conv1(i + 1) and conv2(i) interleaved (issue_conv_pair_1mt): resident weights, one M tile per issuer, no ping-pong
        auto conv_pair = [&](int i1, int i2) {
          const int s = i1 % p.a1_stages, as = i1 % p.acc1_stages, bs = i2 % p.acc2_stages;
          wa.wait(1, BAR(0 + s), (uint32_t)((i1 / p.a1_stages) & 1), 820 + s);
          if (i1 >= p.acc1_stages) wa.wait(2, BAR(6 + as), (uint32_t)((i1 / p.acc1_stages - 1) & 1), 830 + as);
          wa.wait(3, BAR(8), (uint32_t)(i2 & 1), 840);
          if (i2 >= p.acc2_stages) wa.wait(4, BAR(12 + bs), (uint32_t)((i2 / p.acc2_stages - 1) & 1), 850 + bs);
          tc_fence_after();
          const uint64_t ad1 = a1_tmpl + (uint64_t)((smem_u32(A1 + (size_t)s * 2 * a1_bytes) >> 4) & 0x3FFF) + (uint64_t)(wid * 128);
          const uint64_t ad2 = a2_tmpl + (uint64_t)((smem_u32(A2) >> 4) & 0x3FFF) + (uint64_t)(wid * 128);
          const uint64_t bd1 = b_tmpl + (uint64_t)((w1s >> 4) & 0x3FFF), bd2 = b_tmpl + (uint64_t)((w2s >> 4) & 0x3FFF);
          const uint32_t d1 = tmem_base + (uint32_t)(as * p.acc_cols) + (uint32_t)wid * mt_cols;
          const uint32_t d2 = acc2_base + (uint32_t)(bs * p.acc_cols) + (uint32_t)wid * mt_cols;
          const uint64_t ks1 = 2ull * (uint64_t)p.x_rows_alloc, ks2 = 2ull * (uint64_t)p.h_rows_alloc;
          const uint64_t lo1 = (uint64_t)(a1_bytes >> 4), lo2 = (uint64_t)(a2_bytes >> 4);
          if (ksteps == 1) issue_conv_pair_1mt<1>(ad1, bd1, d1, (uint64_t)p.dil, ks1, lo1, ad2, bd2, d2, ks2, lo2, K, kb_step16, idesc, idesc2);
          else if (ksteps == 2) issue_conv_pair_1mt<2>(ad1, bd1, d1, (uint64_t)p.dil, ks1, lo1, ad2, bd2, d2, ks2, lo2, K, kb_step16, idesc, idesc2);
          else issue_conv_pair_1mt<4>(ad1, bd1, d1, (uint64_t)p.dil, ks1, lo1, ad2, bd2, d2, ks2, lo2, K, kb_step16, idesc, idesc2);
          umma_commit_elect(BAR(2 + s));
          umma_commit_elect(BAR(4 + as));
          umma_commit_elect(BAR(9));
          umma_commit_elect(BAR(10 + bs));
        };
        const bool pair_issue = p.pair_issue && !p.pp && p.w_resident && p.a1_stages == 2 && m_tiles <= n_iss &&
                                (ksteps == 1 || ksteps == 2 || ksteps == 4);
        if (p.pp) {
          for (int i = 0; i < n_my; i += 2) {
            conv1(i);
            if (i + 1 < n_my) conv1(i + 1);
            conv2(i);
            if (i + 1 < n_my) conv2(i + 1);
          }
        } else if (n_my > 0) conv1(0);
        if (pair_issue) {
          for (int i = 0; i < n_my; ++i) {
            if (i + 1 < n_my) conv_pair(i + 1, i);
            else conv2(i);
          }
        } else
        for (int i = 0; i < n_my && !p.pp; ++i) {
          if (conv2_first) {
            conv2(i);
            if (i + 1 < n_my) conv1(i + 1);
          } else {
            if (i + 1 < n_my) conv1(i + 1);
            conv2(i);
          }
        }
        wa.end(p.dbg, 2, wid == 0 && L0, n_my);
      }
    }
    __syncwarp();
  } else if (is_epiA) {
    // ------------------------------------------------------------------ epiA warps: acc1 -> A2 (shared memory)
    const int q = warp & 3;
    const int nchunks = NT >> 4;
    uint8_t* A2_hi = A2;
    uint8_t* A2_lo = A2 + a2_bytes;
    uint32_t h_rows = (uint32_t)p.h_rows_alloc;   // row stride of the h planes
    int it = 0;
    WaitAcc<DBG> wa;
    wa.begin();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_batch;
      const int t0 = (tile - b * p.tiles_per_batch) * p.m_out;
      const int Lb = p.lens ? __ldg(p.lens + b) : p.L;
      {  // ---- epiA: acc1 -> +b1 -> lrelu -> fp16 split -> A2 (zero rows outside the sequence: conv2's zero padding)
        const int as = it % p.acc1_stages;
        wa.wait(0, BAR(4 + as), (uint32_t)((it / p.acc1_stages) & 1), 860 + as);
        if (p.pp) {   // in place: conv1 (whose completion acc1_full signals) was the last reader of this buffer's x tile
          A2_hi = A1 + (size_t)as * 2 * a1_bytes;
          A2_lo = A2_hi + a1_bytes;
          h_rows = (uint32_t)p.x_rows_alloc;
        } else if (it >= 1) {
          wa.wait(1, BAR(9), (uint32_t)((it - 1) & 1), 870);
        }
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(as * p.acc_cols);
        for (int c = 0; c < nchunks; ++c) {
          float b1v[16];
          if (p.bias1) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias1 + c * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = __ldg(bp + i);
              b1v[4 * i] = bv.x; b1v[4 * i + 1] = bv.y; b1v[4 * i + 2] = bv.z; b1v[4 * i + 3] = bv.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) b1v[i] = 0.f;
          }
          for (int mt = 0; mt < p.m_tiles; ++mt) {
            const int r = mt * 128 + q * 32 + lane;
            const int gpos = t0 - p2 + r;
            const bool inside = gpos >= 0 && gpos < Lb;
            uint32_t rr[16], r2[16];
            const uint32_t tcol = acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 2 * NT + c * 16);
#if FV_AB_OLD_LD
            tmem_ld16(tcol, rr);
            tmem_ld16(tcol + (uint32_t)NT, r2);
#else
            tmem_ld16x2(tcol, tcol + (uint32_t)NT, rr, r2);
#endif
            uint32_t hp[8], lp[8];
            const uint32_t keep = inside ? 0xffffffffu : 0u;   // rows outside the sequence are conv2's zero padding
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
#if FV_PACKED_F32
              lrelu_split_x2(add3_x2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1]), __uint_as_float(r2[i]),
                                     __uint_as_float(r2[i + 1]), b1v[i], b1v[i + 1]),
                             p.slope, hp[i >> 1], lp[i >> 1]);
#else
              const float v0 = __uint_as_float(rr[i]) + __uint_as_float(r2[i]) + b1v[i];
              const float v1 = __uint_as_float(rr[i + 1]) + __uint_as_float(r2[i + 1]) + b1v[i + 1];
              split_f16x2(lrelu01(v0, p.slope), lrelu01(v1, p.slope), hp[i >> 1], lp[i >> 1]);
#endif
              hp[i >> 1] &= keep;
              lp[i >> 1] &= keep;
            }
            const uint32_t off0 = ((uint32_t)(2 * c) * h_rows + r) * 16;
            const uint32_t off1 = ((uint32_t)(2 * c + 1) * h_rows + r) * 16;
            *reinterpret_cast<uint4*>(A2_hi + off0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            *reinterpret_cast<uint4*>(A2_hi + off1) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
            *reinterpret_cast<uint4*>(A2_lo + off0) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
            *reinterpret_cast<uint4*>(A2_lo + off1) = make_uint4(lp[4], lp[5], lp[6], lp[7]);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (p.pp) {
            mbar_arrive(BAR(8 + as));   // h ready AND the set drained: conv2 may overwrite the columns
          } else {
            mbar_arrive(BAR(6 + as));   // acc1 set drained
            mbar_arrive(BAR(8));        // A2 ready for conv2
          }
        }
      }
    }
    wa.end(p.dbg, 4, q == 0 && lane == 0 && epi_grp == 0, it);
  } else if (is_epiB) {
    // ------------------------------------------------------------------ epiB warps: acc2 -> global
    const int q = warp & 3;
    const int nchunks = NT >> 4;
    int it = 0;
    WaitAcc<DBG> wa;
    wa.begin();
    if (p.pdl) pdl_wait();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_batch;
      const int t0 = (tile - b * p.tiles_per_batch) * p.m_out;
      {  // ---- epiB: acc2 -> +b2 + x (+ MRF accumulate) -> y
        const int bs = it % p.acc2_stages;
        const float* __restrict__ xb = p.x + (long long)b * C * p.L;
        float* __restrict__ yb = p.y + (long long)b * C * p.L;
        const float inv = 1.0f / p.acc_div;
        const uint32_t uL = (uint32_t)p.L;
        const uint32_t acc2 = acc2_base + (uint32_t)(bs * p.acc_cols);
        bool waited = false;
        for (int c = 0; c < nchunks; ++c) {
          float bias[16];
          if (p.bias2) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias2 + c * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = __ldg(bp + i);
              bias[4 * i] = bv.x; bias[4 * i + 1] = bv.y; bias[4 * i + 2] = bv.z; bias[4 * i + 3] = bv.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) bias[i] = 0.f;
          }
          for (int mt = 0; mt < p.m_tiles; ++mt) {
            if (epiB_groups == 2 && ((c * p.m_tiles + mt) & 1) != epi_grp) continue;   // the quarter's other warp takes it
            const int r = mt * 128 + q * 32 + lane;
            const int t = t0 + r;
            const bool ok = r < p.m_out && t < p.L;
            const int o0 = (c * 16) * p.L + t;   // 32-bit offset inside the utterance plane (tc3_plan: C*L < 2^31)
            // the residual does not depend on the accumulators: issue its loads before waiting for the UMMAs
            float xv[16];
            float* py = opaque_ptr(yb + (ok ? o0 : 0));
            // split copies: plane kc of half hf at uint4 index (hf * C/8 + kc) * L + t inside the utterance
            const uint32_t q_h0 = (uint32_t)(2 * c) * uL + (uint32_t)(ok ? t : 0), q_lo = (uint32_t)(C >> 3) * uL;
            if (STK) {   // the skip path is part of conv2's GEMM: nothing to add
#pragma unroll
              for (int i = 0; i < 16; ++i) xv[i] = 0.f;
            } else if (IO == IO_F32) {
              const float* px = opaque_ptr(xb + (ok ? o0 : 0));
#pragma unroll
              for (int i = 0; i < 16; ++i) xv[i] = ok ? __ldg(px + (uint32_t)i * uL) : 0.f;
            } else {
              const uint4* pq = opaque_ptr(reinterpret_cast<const uint4*>(p.xs) + (long long)b * (2 * (C >> 3)) * p.L + q_h0);
              const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
              const uint4 h0 = ok ? __ldg(pq) : z4, h1 = ok ? __ldg(pq + uL) : z4;
              const uint4 l0 = ok ? __ldg(pq + q_lo) : z4, l1 = ok ? __ldg(pq + q_lo + uL) : z4;
              unsplit8(h0, l0, p.inv_slope, xv);
              unsplit8(h1, l1, p.inv_slope, xv + 8);
            }
            if (acc_reads_y(p.acc_mode)) {   // running MRF sum: fold it into the prefetched addend (x + xs)
#pragma unroll
              for (int i = 0; i < 16; ++i) xv[i] += ok ? py[(uint32_t)i * uL] : 0.f;
            }
            if (IO == IO_SPLIT_SPLIT && p.ysum) {
              // last branch of the stage: result = ysum + v / num_kernels = (v + ysum * num_kernels) / num_kernels — folded into
              // the addend here so that the loads are in flight before the accumulator wait (no extra registers)
              const float* ps = opaque_ptr(p.ysum + (long long)b * C * p.L + (ok ? o0 : 0));
#pragma unroll
              for (int i = 0; i < 16; ++i) xv[i] = fmaf(ok ? __ldg(ps + (uint32_t)i * uL) : 0.f, p.acc_div, xv[i]);
            }
            if (!waited) {
              wa.wait(0, BAR(10 + bs), (uint32_t)((it / p.acc2_stages) & 1), 880 + bs);
              tc_fence_after();
              waited = true;
            }
            uint32_t rr[16], r2[16];
            const uint32_t tcol = acc2 + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 2 * NT + c * 16);
#if FV_AB_OLD_LD
            tmem_ld16(tcol, rr);
            tmem_ld16(tcol + (uint32_t)NT, r2);
#else
            tmem_ld16x2(tcol, tcol + (uint32_t)NT, rr, r2);
#endif
            if (!ok) continue;
            float v[16];
#if FV_PACKED_F32
            const bool scaled = p.acc_mode == ACC_ADD_DIV || p.acc_mode >= ACC_STORE_SCALE;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float2 t = __fadd2_rn(add3_x2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1]), __uint_as_float(r2[i]),
                                            __uint_as_float(r2[i + 1]), bias[i], bias[i + 1]),
                                    make_float2(xv[i], xv[i + 1]));
              if (scaled) t = __fmul2_rn(t, make_float2(inv, inv));
              v[i] = t.x; v[i + 1] = t.y;
            }
#else
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v[i] = (__uint_as_float(rr[i]) + __uint_as_float(r2[i]) + bias[i]) + xv[i];
            if (p.acc_mode == ACC_ADD_DIV || p.acc_mode >= ACC_STORE_SCALE) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= inv;
            }
#endif
            if (IO == IO_SPLIT_SPLIT) {   // next unit's input: lrelu -> fp16 hi / lo rows of the blocked planes
              uint32_t hp[8], lp[8];
#pragma unroll
              for (int i = 0; i < 16; i += 2)
                split_f16x2(lrelu01(v[i], p.slope), lrelu01(v[i + 1], p.slope), hp[i >> 1], lp[i >> 1]);
              uint4* pw = opaque_ptr(reinterpret_cast<uint4*>(p.ys) + (long long)b * (2 * (C >> 3)) * p.L + q_h0);
              pw[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
              pw[uL] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
              pw[q_lo] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
              pw[q_lo + uL] = make_uint4(lp[4], lp[5], lp[6], lp[7]);
            } else if (p.acc_mode == ACC_RED_SCALE) {   // y += v / num_kernels without reading y (one thread per element)
#pragma unroll
              for (int i = 0; i < 16; ++i) red_add_f32(py + (uint32_t)i * uL, v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) py[(uint32_t)i * uL] = v[i];
            }
          }
        }
        if (!waited) {   // a warp without an iteration in this tile still follows the accumulator hand-shake
          wa.wait(0, BAR(10 + bs), (uint32_t)((it / p.acc2_stages) & 1), 880 + bs);
          tc_fence_after();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(12 + bs));
      }
    }
    wa.end(p.dbg, 5, q == 0 && lane == 0 && epi_grp == 0, it);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC2_LOADER_WARPS) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

inline size_t tc3_smem_bytes(const Tc3Args& p) {
  const size_t w = p.w_resident ? (size_t)(p.kblocks + p.kblocks2) * p.C * 64 : (size_t)p.w_stages * p.stage_bytes;
  return (size_t)p.a1_stages * 2 * (p.x_rows_alloc + (p.stack ? p.m_out : 0)) * p.C * 2 +
         (p.pp ? 0 : 2ULL * p.h_rows_alloc * p.C * 2) + w + 24 * 8;
}

// conv1/conv2 must be same-shape C->C convs with K taps (conv2 dilation 1) whose images are single-N-tile (C <= 64).
// Both images resident in shared memory when they fit; otherwise (C = 64, k = 7 / 11) they are streamed through a ring,
// one tap per slot, and the tile is made as tall as TMEM allows (single accumulator sets) because every tile re-streams
// all 2*K taps from L2: positions per weight pass is what bounds those units.
inline bool tc3_plan(int B, int C, int L, int K, int dil, Tc3Args& p, bool split = false) {
  if (C % 16 || C > 64 || K % 2 == 0) return false;
  // split (TMA-fed) units: A1 plane stride rounded up to 8 rows so that every TMA box lands 128-byte aligned
  auto xr = [split](long long rows) { return split ? (rows + 7) / 8 * 8 : rows; };
  if ((long long)C * L >= 0x7fffffffLL - 65536) return false;   // 32-bit offsets inside the utterance plane
  const int ksteps = C / 16, kblocks = K * ksteps;
  const long long BUDGET = 225 * 1024;
  int best_m = 0, best_a1 = 0, best_acc1 = 0, best_acc2 = 2, best_res = 1, best_wst = 0, best_pp = 0;
  double best_sc = -1;
  static const int force_m_env = getenv("FV_TC3_M") ? atoi(getenv("FV_TC3_M")) : 0;   // tuning knob
  static const int ring_env = getenv("FV_TC3_RING") ? atoi(getenv("FV_TC3_RING")) : 1;   // 0: never stream weights
  int force_m = force_m_env;
  const long long stage_bytes = (long long)ksteps * C * 64;   // one tap
retry:
  for (int acc1 = 2; acc1 >= 1; --acc1)
    for (int a1 = 2; a1 >= 1; --a1)
      for (int m = 8; m >= 1; --m) {
        if (force_m > 0 && m != force_m) continue;
        if ((acc1 + 2) * m * 2 * C > 512) continue;
        const long long x_rows = xr(128LL * m + (long long)(K - 1) * dil), h_alloc = 128LL * m + (K - 1);
        const long long sm = a1 * 2 * x_rows * C * 2 + 2 * h_alloc * C * 2 + 2LL * kblocks * C * 64 + 256;
        if (sm > BUDGET) continue;
        const int m_out = 128 * m - (K - 1);
        if (m_out <= 0) continue;
        const long long tiles = (long long)((L + m_out - 1) / m_out) * B;
        const long long gx = std::min<long long>(148, tiles);
        const long long waves = (tiles + gx - 1) / gx;
        const double tail = (double)tiles / (double)(waves * gx);
        double sc = (double)m_out / (128.0 * m) * tail;          // useful fraction of computed rows x wave balance
        sc *= (a1 == 2 ? 1.0 : 0.8) * (acc1 == 2 ? 1.0 : 0.85);  // overlap bonuses
        sc *= (m >= 2 ? 1.0 : 0.9);
        if (sc > best_sc) { best_sc = sc; best_m = m; best_a1 = a1; best_acc1 = acc1; best_acc2 = 2; best_res = 1; best_wst = 0; }
      }
  // FV_TC3_PP: 0 = no ping-pong tiles, 1 = streamed-weight units only, 2 = also resident units with k >= 7 (default),
  // 3 = every resident unit
  static const int pp_env = getenv("FV_TC3_PP") ? atoi(getenv("FV_TC3_PP")) : 2;
  // split (TMA-fed) units at C = 16: ping-pong tiles also for k = 3 (m 3 -> 8: eight independent accumulator chains per
  // tile keep the tensor pipe fed, profiles/r02_notes.md: 311 -> 276 kclk); C = 32 k = 3 measured neutral then (294 vs 297 kclk, before the second epiB group)
  // re-measured on the final binary (gpurun r2ab, FV_TC3_PP=3): C = 32 k = 3 1.109-1.125 -> 1.015 ms per three units (m 2 -> 4: four
  // accumulator chains per tile instead of two; the issuers of the m = 2 plan sat 91 % busy at ~210 clk per dependent UMMA)
  if (best_sc >= 0 && (pp_env >= 3 || (pp_env == 2 && (K >= 7 || (split && C <= 32))))) {
    // Resident weights + ping-pong tiles: the same two tiles in flight need half the TMEM and no A2 buffer, so the tile can be
    // taller (C=32: m 2 -> 4, C=16: m 3-4 -> 8): less conv2 halo recompute and fewer per-tile hand-offs.  Measured in one call
    // (gpurun_out/pp2_*): C=32 k=11 0.676 -> 0.60 ms, C=16 k=11 0.61 -> 0.52, k=7 -3 %, k=3 +-3 % (left on the old plans);
    // HiFi-GAN step 17.17 -> 16.88 ms.
    static const int pp_m_env = getenv("FV_TC3_PP_M") ? atoi(getenv("FV_TC3_PP_M")) : 0;
    for (int m = 8; m > best_m; --m) {
      if (pp_m_env > 0 && m > pp_m_env) continue;
      if (2 * m * 2 * C > 512) continue;
      const long long x_rows = xr(128LL * m + (long long)(K - 1) * dil);
      const long long sm = 2 * 2 * x_rows * C * 2 + 2LL * kblocks * C * 64 + 256;
      if (sm > BUDGET) continue;
      const int m_out = 128 * m - (K - 1);
      const long long tiles = (long long)((L + m_out - 1) / m_out) * B;
      if (tiles < 148 * 4) continue;            // keep enough tiles per CTA for the two-tile pipeline to fill
      best_m = m; best_a1 = 2; best_acc1 = 2; best_acc2 = 2; best_res = 1; best_wst = 0; best_pp = 1;
      break;
    }
  }
  if (best_sc < 0 && ring_env && stage_bytes <= 32768) {   // streamed weights
    static const int ring_m_env = getenv("FV_TC3_RING_M") ? atoi(getenv("FV_TC3_RING_M")) : 0;   // tuning knob
    // ping-pong tiles first: two tiles in flight, each with one in-place A buffer and one accumulator set
    for (int m = 4; m >= 1 && best_sc < 0 && pp_env; --m) {
      if (force_m > 0 && m != force_m) continue;
      if (ring_m_env > 0 && m != ring_m_env) continue;
      if (2 * m * 2 * C > 512 || 128 * m - (K - 1) <= 0) continue;
      for (int wst = 4; wst >= 3; --wst) {
        const long long x_rows = xr(128LL * m + (long long)(K - 1) * dil);
        const long long sm = 2 * 2 * x_rows * C * 2 + wst * stage_bytes + 256;
        if (sm > BUDGET) continue;
        best_sc = 1.0; best_m = m; best_a1 = 2; best_acc1 = 2; best_acc2 = 2; best_res = 0; best_wst = wst; best_pp = 1;
        break;
      }
    }
    for (int m = 4; m >= 1 && best_sc < 0; --m) {
      if (force_m > 0 && m != force_m) continue;
      if (ring_m_env > 0 && m != ring_m_env) continue;
      for (int acc = 2; acc >= 1 && best_sc < 0; --acc) {        // accumulator sets of conv1 and of conv2
        if (2 * acc * m * 2 * C > 512) continue;
        for (int a1 = 2; a1 >= 1 && best_sc < 0; --a1)
          for (int wst = 4; wst >= 3; --wst) {
            const long long x_rows = xr(128LL * m + (long long)(K - 1) * dil), h_alloc = 128LL * m + (K - 1);
            const long long sm = a1 * 2 * x_rows * C * 2 + 2 * h_alloc * C * 2 + wst * stage_bytes + 256;
            if (sm > BUDGET || 128 * m - (K - 1) <= 0) continue;
            best_sc = 1.0; best_m = m; best_a1 = a1; best_acc1 = acc; best_acc2 = acc; best_res = 0; best_wst = wst;
            break;
          }
      }
    }
  }
  if (best_sc < 0 && force_m > 0) { force_m = 0; goto retry; }   // forced tile count infeasible for this width
  if (best_sc < 0) return false;
  p.B = B; p.C = C; p.L = L; p.K = K; p.dil = dil;
  p.m_tiles = best_m;
  p.x_rows = 128 * best_m + (K - 1) * dil;
  p.x_rows_alloc = (int)xr(p.x_rows);
  p.h_rows_alloc = 128 * best_m + (K - 1);
  p.m_out = 128 * best_m - (K - 1);
  p.a1_stages = best_a1;
  p.acc1_stages = best_acc1;
  p.acc2_stages = best_acc2;
  p.w_resident = best_res;
  p.pp = best_pp;
  p.w_stages = best_wst;
  p.stage_bytes = (int)stage_bytes;
  // ring mode: warp 11 is the weight producer
  p.n_issuers = std::min(best_m, best_res ? TC2_ISSUE_WARPS : TC2_ISSUE_WARPS - 1);
  static const int force_iss = getenv("FV_TC3_ISSUERS") ? atoi(getenv("FV_TC3_ISSUERS")) : 0;   // tuning knob
  if (force_iss > 0) p.n_issuers = std::max(1, std::min(p.n_issuers, force_iss));
  p.acc_cols = best_m * 2 * C;
  int cols = 32;
  while (cols < (best_pp ? 2 : best_acc1 + best_acc2) * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  p.ksteps = ksteps;
  p.kblocks = kblocks;
  p.stack = 0; p.k2 = K; p.ksteps2 = ksteps; p.kblocks2 = kblocks; p.reflect = 0;
  p.tiles_per_batch = (L + p.m_out - 1) / p.m_out;
  p.total_tiles = p.tiles_per_batch * B;
  p.idesc = make_idesc_f16(128, C);
  p.idesc2 = make_idesc_f16(128, 2 * C);
  static const int flat_env = getenv("FV_LOADER_FLAT") ? atoi(getenv("FV_LOADER_FLAT")) : 1;
  p.ld_per = p.ld_rounds = 0;
  // measured (profiles/r01_notes.md, ab2 / ab3): flat wins where the unit is loader-bound and the pair loop needs a second
  // round (k = 3 with more than 8 pairs: C=32 -6 %); it loses where the pair loop is one round (C=16 k=3 +15 %) and on the
  // issue-bound k = 7 / 11 units (C=32 k=11 +18 %)
  if ((flat_env >= 2 || (flat_env && K <= 3)) && (C / 8) * ((p.x_rows + 127) / 128) > TC2_LOADER_WARPS)
    flat_plan((C / 8) * p.x_rows, p.ld_per, p.ld_rounds);
  return true;
}

// Is the split (TMA-native) activation path usable at all on this machine?  (driver exports cuTensorMapEncodeTiled)
inline bool tc3_split_available() { return tma_encode_fn() != nullptr; }

// Launch of a planned fused kernel (ResBlock1 unit or ResidualStack): per-device attributes once, persistent grid, optional PDL,
// FV_STALL_DEBUG instrumentation.  returns 0 launched, -1 CUDA error.
inline int tc3_launch(Tc3Args& p, int io, const CUtensorMap& tm_main, const CUtensorMap& tm_tail, cudaStream_t st,
                      const char* title) {
  static std::mutex mu;
  static bool attr_set[64] = {};
  static int num_sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_set[dev]) {
      const int mx = 227 * 1024;
      if (cudaFuncSetAttribute(conv_tc3_fused_kernel<false, IO_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc3_fused_kernel<true, IO_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc3_fused_kernel<false, IO_SPLIT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc3_fused_kernel<true, IO_SPLIT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc3_fused_kernel<false, IO_SPLIT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc3_fused_kernel<true, IO_SPLIT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc3_fused_kernel<false, IO_STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess ||
          cudaFuncSetAttribute(conv_tc3_fused_kernel<true, IO_STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx) != cudaSuccess)
        return -1;
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
      num_sms[dev] = prop.multiProcessorCount;
      attr_set[dev] = true;
    }
  }
  int gx = std::min(num_sms[dev], p.total_tiles);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(gx, 1, 1);
  cfg.blockDim = dim3(TC3_THREADS, 1, 1);
  cfg.dynamicSmemBytes = tc3_smem_bytes(p);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  p.pdl = (tc_pdl_enabled(p.total_tiles, num_sms[dev]) && !tc_stall_debug()) ? 1 : 0;
  if (p.pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaError_t le;
  if (tc_stall_debug()) {
    StallReport rep;
    if (!rep.begin(gx)) return -1;
    p.dbg = rep.dev;
    if (io == IO_F32) le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<true, IO_F32>, p, tm_main, tm_tail);
    else if (io == IO_SPLIT_SPLIT) le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<true, IO_SPLIT_SPLIT>, p, tm_main, tm_tail);
    else if (io == IO_STACK) le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<true, IO_STACK>, p, tm_main, tm_tail);
    else le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<true, IO_SPLIT_F32>, p, tm_main, tm_tail);
    static const char* const roles[8] = {"loader", "producer", "issuer0", nullptr, "epiA", "epiB", nullptr, nullptr};
    static const char* const slots[8][5] = {{"a1_empty", nullptr, nullptr, nullptr, nullptr},
                                            {"w_empty", nullptr, nullptr, nullptr, nullptr},
                                            {"w_full", "a1_full", "acc1_empty", "a2_full", "acc2_empty"},
                                            {nullptr, nullptr, nullptr, nullptr, nullptr},
                                            {"acc1_full", "a2_empty", nullptr, nullptr, nullptr},
                                            {"acc2_full", nullptr, nullptr, nullptr, nullptr}};
    char full[320];
    snprintf(full, sizeof full, "%s grid=%d", title, gx);
    rep.finish(st, full, roles, slots);
  } else {
    if (io == IO_F32) le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<false, IO_F32>, p, tm_main, tm_tail);
    else if (io == IO_SPLIT_SPLIT) le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<false, IO_SPLIT_SPLIT>, p, tm_main, tm_tail);
    else if (io == IO_STACK) le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<false, IO_STACK>, p, tm_main, tm_tail);
    else le = cudaLaunchKernelEx(&cfg, conv_tc3_fused_kernel<false, IO_SPLIT_F32>, p, tm_main, tm_tail);
  }
  if (le != cudaSuccess) return -1;
  g_launches++;
  g_tc_launches++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// returns 0 launched, 1 not applicable, -1 CUDA error.
// io = IO_F32: x / y are fp32 [B, C, L].  IO_SPLIT_SPLIT / IO_SPLIT_F32: x is a split-format buffer (fv_tma.cuh) and, for
// IO_SPLIT_SPLIT, so is y (acc_mode must be ACC_STORE); lens (ragged batches) is not supported in split mode.
inline int launch_fused_unit(const float* x, float* y, const float* b1, const float* b2, const TcLayer& l1,
                             const TcLayer& l2, int B, int C, int L, int K, int dil, float slope, int acc_mode,
                             float acc_div, cudaStream_t st, const int* lens = nullptr, int io = IO_F32,
                             const float* ysum = nullptr) {
  if (!l1.eligible || !l2.eligible || !l1.image || !l2.image || l1.n_tiles != 1 || l2.n_tiles != 1) return 1;
  tc_apply_env_once();
  Tc3Args p{};
  if (!(slope >= 0.f && slope <= 1.f)) return 1;   // lrelu01
  if (io != IO_F32 && (lens != nullptr || !(slope > 0.f) || !tc3_split_available())) return 1;
  if (io == IO_SPLIT_SPLIT && !(acc_mode == ACC_STORE || (acc_mode == ACC_STORE_SCALE && ysum != nullptr))) return 1;
  if (!tc3_plan(B, C, L, K, dil, p, io != IO_F32)) return 1;
  p.x = x; p.y = y; p.bias1 = b1; p.bias2 = b2; p.lens = lens;
  p.xs = x; p.ys = y; p.ysum = (io == IO_SPLIT_SPLIT) ? ysum : nullptr;
  p.inv_slope = slope > 0.f ? 1.0f / slope : 1.0f;
  static const int epi_env = getenv("FV_TC3_EPI") ? atoi(getenv("FV_TC3_EPI")) : 2;   // epilogue warp groups in split mode
  p.epi_groups = (io != IO_F32 && epi_env >= 2) ? 2 : 1;
  // measured (gpurun r2q): the interleaved pair must wait for epiA(i) AND the load of tile i+1 before either conv starts,
  // which costs more than the second accumulator chain gains (C=64 k=3 units +8..+20 %) -> opt-in only
  static const int pair_env = getenv("FV_TC3_PAIR") ? atoi(getenv("FV_TC3_PAIR")) : 0;
  p.pair_issue = pair_env ? 1 : 0;
  p.w1img = l1.image; p.w2img = l2.image;
  p.slope = slope; p.acc_mode = acc_mode; p.acc_div = acc_div;
  CUtensorMap tm_main, tm_tail;
  memset(&tm_main, 0, sizeof tm_main);
  memset(&tm_tail, 0, sizeof tm_tail);
  if (io != IO_F32) {
    const long long planes = (long long)B * 2 * (C / 8);
    const int tail = p.x_rows % TMA_SPLIT_RB;
    if (!tma_encode_split(&tm_main, x, L, planes, TMA_SPLIT_RB)) return 1;
    if (tail) {
      if (!tma_encode_split(&tm_tail, x, L, planes, tail)) return 1;
    } else {
      tm_tail = tm_main;
    }
  }
  char title[256];
  snprintf(title, sizeof title,
           "tc3 C=%d K=%d dil=%d L=%d B=%d acc=%d io=%d | mt=%d m_out=%d a1_st=%d acc1_st=%d acc2_st=%d resident=%d w_st=%d pp=%d issuers=%d "
           "tiles=%d", C, K, dil, L, B, acc_mode, io, p.m_tiles, p.m_out, p.a1_stages, p.acc1_stages, p.acc2_stages,
           p.w_resident, p.w_stages, p.pp, p.n_issuers, p.total_tiles);
  return tc3_launch(p, io, tm_main, tm_tail, st, title);
}

// Fused ResidualStack (Tc3Args::stack): y = W_1x1 * lrelu(conv_dil(reflect_pad(lrelu(c))) + b_dil) + W_skip * c + (b_1x1 + b_skip)
// (modules.py:372-382) in one launch.  l1 = the dilated conv's image, l2 = the pair image [W_1x1 ; W_skip] (Cin = 2C, one tap),
// b2 = the summed bias.  returns 0 launched, 1 not applicable (the caller runs the two-launch path), -1 CUDA error.
inline bool tc3_plan_stack(int B, int C, int L, int K, int dil, Tc3Args& p) {
  if (C % 16 || C > 64 || K % 2 == 0 || K < 3) return false;
  if ((long long)C * L >= 0x7fffffffLL - 65536) return false;   // 32-bit offsets inside the utterance plane
  const int ksteps = C / 16, kblocks = K * ksteps, kblocks2 = 2 * ksteps;
  const long long BUDGET = 225 * 1024;
  static const int force_m = getenv("FV_STACK_M") ? atoi(getenv("FV_STACK_M")) : 0;   // tuning knob
  int best_m = 0;
  double best_sc = -1;
  for (int m = 8; m >= 1; --m) {
    if (force_m > 0 && m != force_m) continue;
    if (2 * m * 2 * C > 512) continue;                           // two accumulator sets (ping-pong tiles)
    const long long x_rows = 128LL * m + (long long)(K - 1) * dil;
    const long long sm = 2 * 2 * (x_rows + 128LL * m) * C * 2 + (long long)(kblocks + kblocks2) * C * 64 + 256;
    if (sm > BUDGET) continue;
    const long long tiles = (long long)((L + 128 * m - 1) / (128 * m)) * B;
    const long long gx = std::min<long long>(148, tiles);
    const long long waves = (tiles + gx - 1) / gx;
    double sc = (double)tiles / (double)(waves * gx);            // wave balance
    sc *= (double)(128 * m) / (double)x_rows;                    // halo re-read
    if (waves < 4) sc *= 0.8;                                    // the two-tile pipeline needs a few tiles per CTA to fill
    if (sc > best_sc + 1e-9) { best_sc = sc; best_m = m; }
  }
  if (best_m == 0) return false;
  p.B = B; p.C = C; p.L = L; p.K = K; p.dil = dil;
  p.m_tiles = best_m;
  p.x_rows = 128 * best_m + (K - 1) * dil;
  p.x_rows_alloc = p.x_rows;
  p.h_rows_alloc = p.x_rows;
  p.m_out = 128 * best_m;
  p.a1_stages = 2; p.acc1_stages = 2; p.acc2_stages = 2;
  p.w_resident = 1; p.pp = 1; p.w_stages = 0; p.stage_bytes = 0;
  p.n_issuers = std::min(best_m, TC2_ISSUE_WARPS);
  p.acc_cols = best_m * 2 * C;
  int cols = 32;
  while (cols < 2 * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  p.ksteps = ksteps; p.kblocks = kblocks;
  p.stack = 1; p.k2 = 1; p.ksteps2 = 2 * ksteps; p.kblocks2 = kblocks2;
  p.tiles_per_batch = (L + p.m_out - 1) / p.m_out;
  p.total_tiles = p.tiles_per_batch * B;
  p.idesc = make_idesc_f16(128, C);
  p.idesc2 = make_idesc_f16(128, 2 * C);
  flat_plan((C / 8) * p.x_rows, p.ld_per, p.ld_rounds);
  return true;
}

inline int launch_fused_stack(const float* c, float* y, const float* b1, const float* b2, const TcLayer& l1, const TcLayer& l2,
                              int B, int C, int L, int K, int dil, float slope, bool reflect, cudaStream_t st,
                              const int* lens = nullptr) {
  if (!l1.eligible || !l2.eligible || !l1.image || !l2.image || l1.n_tiles != 1 || l2.n_tiles != 1) return 1;
  if (l1.NT != C || l2.NT != C) return 1;
  if (!(slope >= 0.f && slope <= 1.f)) return 1;   // lrelu01
  tc_apply_env_once();
  Tc3Args p{};
  if (!tc3_plan_stack(B, C, L, K, dil, p)) return 1;
  p.x = c; p.y = y; p.bias1 = b1; p.bias2 = b2; p.lens = lens;
  p.inv_slope = 1.f; p.epi_groups = 1; p.pair_issue = 0;
  p.w1img = l1.image; p.w2img = l2.image;
  p.slope = slope; p.acc_mode = ACC_STORE; p.acc_div = 1.f;
  p.reflect = reflect ? 1 : 0;
  CUtensorMap tm_main, tm_tail;
  memset(&tm_main, 0, sizeof tm_main);
  memset(&tm_tail, 0, sizeof tm_tail);
  char title[256];
  snprintf(title, sizeof title, "tc3-stack C=%d K=%d dil=%d L=%d B=%d | mt=%d x_rows=%d issuers=%d tiles=%d", C, K, dil, L, B,
           p.m_tiles, p.x_rows, p.n_issuers, p.total_tiles);
  return tc3_launch(p, IO_STACK, tm_main, tm_tail, st, title);
}

// ------------------------------------------------------------------------------------------------
// conv_narrow7_staged_kernel: the HBM-bound k = 7 output convs (conv_post 16 -> 1, LastLayer 32 -> 1, MB conv_post 64 -> 4) with
// the input streamed through shared memory by the bulk-copy engine.  conv_narrow7_kernel (fv_kernels.cuh) reads every sample
// three times through L1 with at most six 16-byte loads in flight per thread and reaches 0.53 of the HBM peak; here a persistent
// CTA keeps N7_STAGES stages of [16 channels][N7_TL + 8 samples] in flight (cp.async.bulk per channel row, mbarrier
// complete_tx), every sample crosses L1 once, and the threads read their 12-sample windows from shared memory (conflict-free
// 16-byte reads).  Channels are walked in groups of 16 per stage with the accumulators kept in registers across the groups of a
// tile, so any Cin % 16 == 0 fits.  Same sums in the same order as conv_narrow7_kernel: identical bits.
// Algorithmic bytes per position: 4 * (Cin + N).
// ------------------------------------------------------------------------------------------------
constexpr int N7_TL = 1024;      // positions per tile = 4 per thread x 256 threads
constexpr int N7_STAGES = 3;
constexpr int N7_ROW = N7_TL + 8;
template <int NOUT>
__global__ void __launch_bounds__(256, 1) conv_narrow7_staged_kernel(const ConvArgs a, int tiles_per_b, int total_tiles) {
  constexpr int K = 7, PADL = 3, T = 4;
  extern __shared__ __align__(128) uint8_t n7_smem[];
  float* stage0 = reinterpret_cast<float*>(n7_smem);                                   // [N7_STAGES][16][N7_ROW]
  float* wsm = stage0 + N7_STAGES * 16 * N7_ROW;                                       // [Cin][NOUT][8] (tap 7 = 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + a.Cin * NOUT * 8);
  const uint32_t bar0 = smem_u32(bars);
  const int tid = threadIdx.x;
  for (int i = tid; i < a.Cin * NOUT * 8; i += blockDim.x) {
    const int j = i & 7, n = (i >> 3) % NOUT, ci = i / (8 * NOUT);
    wsm[i] = (j < K && n < a.N) ? __ldg(a.w + ((long long)ci * K + j) * a.N + n) : 0.f;   // derived image is [Cin][K][N]
  }
  if (tid == 0) {
    for (int s = 0; s < N7_STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int ngroups = a.Cin >> 4;
  int n_my = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) ++n_my;
  const int n_iters = n_my * ngroups;
  // stage `it`: channels [16*cg, 16*cg + 16) of tile (it / ngroups), samples [t0 - 4, t0 + N7_TL + 4) clipped to the row
  auto issue = [&](int it) {   // warp 0, converged
    const int lt = it / ngroups, cg = it - lt * ngroups;
    const int tile = blockIdx.x + lt * (int)gridDim.x;
    const int b = tile / tiles_per_b;
    const int t0 = (tile - b * tiles_per_b) * N7_TL;
    const int lo = max(t0 - 4, 0), hi = min(t0 + N7_TL + 4, a.Lin);
    const uint32_t bytes = (uint32_t)(hi - lo) * 4u;
    const int s = it % N7_STAGES;
    const uint32_t bar = bar0 + 8u * s;
    const int lane = tid & 31;
    if (lane == 0) mbar_expect_tx(bar, 16u * bytes);
    __syncwarp();
    if (lane < 16) {
      const float* src = a.x + (long long)b * a.x_bs + (long long)(cg * 16 + lane) * a.Lin + lo;
      bulk_g2s(smem_u32(stage0 + ((size_t)s * 16 + lane) * N7_ROW + (lo - (t0 - 4))), src, bytes, bar);
    }
  };
  if (tid < 32)
    for (int it = 0; it < N7_STAGES && it < n_iters; ++it) issue(it);
  const float slope = a.pre_slope;
  float acc[NOUT][T];
  for (int it = 0; it < n_iters; ++it) {
    const int lt = it / ngroups, cg = it - lt * ngroups;
    const int tile = blockIdx.x + lt * (int)gridDim.x;
    const int b = tile / tiles_per_b;
    const int tt = (tile - b * tiles_per_b) * N7_TL;
    const int t0 = tt + T * tid;                         // this thread's first output
    const int Lb = a.lens ? __ldg(a.lens + b) : a.Lin;
    const int s = it % N7_STAGES;
    if (cg == 0) {
#pragma unroll
      for (int n = 0; n < NOUT; ++n)
#pragma unroll
        for (int q = 0; q < T; ++q) acc[n][q] = 0.f;
    }
    mbar_wait(bar0 + 8u * s, (uint32_t)((it / N7_STAGES) & 1), 700 + s);
    if (t0 < a.Lpos) {
      if (t0 - 4 >= 0 && t0 + T + 4 <= Lb && t0 + T <= a.Lpos) {   // interior: the whole 12-sample window is real data
        const float* st = stage0 + (size_t)s * 16 * N7_ROW + T * tid;
#pragma unroll 4
        for (int c = 0; c < 16; ++c) {
          const float4* xr = reinterpret_cast<const float4*>(st + c * N7_ROW);
          float w[T + 8];
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const float4 f = xr[v];
            w[4 * v] = f.x; w[4 * v + 1] = f.y; w[4 * v + 2] = f.z; w[4 * v + 3] = f.w;
          }
#pragma unroll
          for (int i = 1; i < T + 7; ++i) w[i] = pre_act(w[i], slope);   // w[0], w[11] are never used (window t0-3 .. t0+6)
          const int ci = cg * 16 + c;
#pragma unroll
          for (int n = 0; n < NOUT; ++n) {
            const float4 c0 = *reinterpret_cast<const float4*>(wsm + (ci * NOUT + n) * 8);
            const float4 c1 = *reinterpret_cast<const float4*>(wsm + (ci * NOUT + n) * 8 + 4);
            const float cw[7] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z};
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
              for (int q = 0; q < T; ++q) acc[n][q] = fmaf(cw[j], w[q + j + (4 - PADL)], acc[n][q]);
          }
        }
      } else {   // padding / ragged end / tail: scalar taps with the generic index rules, straight from global memory
        const float* xb = a.x + (long long)b * a.x_bs;
        for (int c = 0; c < 16; ++c) {
          const int ci = cg * 16 + c;
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
              for (int n = 0; n < NOUT; ++n) acc[n][q] = fmaf(wsm[(ci * NOUT + n) * 8 + j], xv, acc[n][q]);
            }
          }
        }
      }
      if (cg == ngroups - 1) {
        float* yb = a.y + (long long)b * a.y_bs;
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
          if (t0 + 4 <= a.Lpos) {
            *reinterpret_cast<float4*>(yo) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
            for (int q = 0; q < 4 && t0 + q < a.Lpos; ++q) yo[q] = o[q];
          }
        }
      }
    }
    __syncthreads();                                     // every thread is done with stage s
    if (tid < 32 && it + N7_STAGES < n_iters) issue(it + N7_STAGES);
  }
}
inline size_t n7_staged_smem(const ConvArgs& a, int nout) {
  return (size_t)N7_STAGES * 16 * N7_ROW * sizeof(float) + (size_t)a.Cin * nout * 8 * sizeof(float) + N7_STAGES * 8 + 128;
}
inline bool conv_narrow7_staged_ok(const ConvArgs& a) {
  // Measured (gpurun r2v, B = 32, T = 1000): conv_post 16 -> 1 0.150 -> 0.20 ms (SLOWER: one 256-thread CTA per SM cannot hide
  // the shared-memory / FMA latency the 32-warp occupancy of conv_narrow7_kernel hides), LastLayer 32 -> 1 0.400 -> 0.385 ms,
  // MB conv_post 64 -> 4 0.81 ms against 0.48 ms on tcgen05 -> opt-in only (FV_NARROW7_STAGED=1); bit-identical either way (test).
  static const bool env = getenv("FV_NARROW7_STAGED") != nullptr && atoi(getenv("FV_NARROW7_STAGED")) != 0;
  return env && conv_narrow7_ok(a) && a.Cin % 16 == 0 && a.Lin >= N7_TL && n7_staged_smem(a, a.N == 1 ? 1 : (a.N == 2 ? 2 : 4)) <= 225 * 1024;
}
// returns cudaSuccess when launched
inline cudaError_t launch_conv_narrow7_staged(const ConvArgs& a, cudaStream_t st) {
  static std::mutex mu;
  static bool attr_set[64] = {};
  static int num_sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_set[dev]) {
      const int mx = 227 * 1024;
      cudaError_t e;
      if ((e = cudaFuncSetAttribute(conv_narrow7_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(conv_narrow7_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(conv_narrow7_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
      cudaDeviceProp prop;
      if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) return e;
      num_sms[dev] = prop.multiProcessorCount;
      attr_set[dev] = true;
    }
  }
  const int tiles_per_b = (a.Lpos + N7_TL - 1) / N7_TL;
  const long long total = (long long)tiles_per_b * a.B;
  if (total >= 0x7fffffffLL) return cudaErrorInvalidValue;
  const int grid = (int)std::min<long long>(num_sms[dev], total);
  const int nout = a.N == 1 ? 1 : (a.N == 2 ? 2 : 4);
  const size_t smem = n7_staged_smem(a, nout);
  if (nout == 1) conv_narrow7_staged_kernel<1><<<grid, 256, smem, st>>>(a, tiles_per_b, (int)total);
  else if (nout == 2) conv_narrow7_staged_kernel<2><<<grid, 256, smem, st>>>(a, tiles_per_b, (int)total);
  else conv_narrow7_staged_kernel<4><<<grid, 256, smem, st>>>(a, tiles_per_b, (int)total);
  g_launches++;
  return cudaGetLastError();
}

}  // namespace fv
