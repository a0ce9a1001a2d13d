// C ABI of fastvocoder_b200 (see include/fastvocoder_b200.h) + the host-side executor that walks the
// layer graph of fv_model.h and launches the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fastvocoder_b200.h"
#include "fv_kernels.cuh"
#include "fv_model.h"
#include "fv_tc.cuh"

namespace fv {
std::atomic<long long> g_launches{0};
std::atomic<long long> g_tc_launches{0};
thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define FV_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fv::fail(FV_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                                  \
  } while (0)

static inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}
}  // namespace fv

struct fv_handle {
  fv::Model model;
  const float* packed = nullptr;  // caller-owned canonical weights
  float* derived = nullptr;       // library-owned kernel-side images (fp32)
  fv::TcWeights tc;               // library-owned fp16 hi/lo images for the tcgen05 path
  const float* pqmf_ana = nullptr;
  const float* pqmf_syn = nullptr;
  bool bound = false;
};

namespace fv {

// ---- weight derivation ----------------------------------------------------------------------------
static inline float* pair_bias_ptr(const Layer& l, float* wd) {   // summed bias lives right after the image
  return wd + ((int64_t)l.Cin * l.N + 63) / 64 * 64;
}
static int derive_pair(const Layer& l, const float* wa, const float* wb, const float* ba, const float* bb, float* wd,
                       cudaStream_t st) {
  derive_pair_kernel<<<grid_for(2LL * l.Cout * l.Cout), 256, 0, st>>>(wa, wb, ba, bb, wd, pair_bias_ptr(l, wd), l.Cout);
  g_launches++;
  FV_CUDA(cudaGetLastError());
  return FV_OK;
}

static int derive_layer(const Layer& l, const float* w_canon, float* wd, cudaStream_t st) {
  const long long n = (long long)l.Cin * l.Kd * l.N;
  const int g = grid_for(n);
  if (l.type == L_CONV) derive_conv_kernel<<<g, 256, 0, st>>>(w_canon, wd, l.Cout, l.Cin, l.K);
  else if (l.type == L_CONVT) derive_convt_kernel<<<g, 256, 0, st>>>(w_canon, wd, l.Cin, l.Cout, l.K, l.stride, l.Kd);
  else if (l.type == L_UPCONV)
    derive_upconv_kernel<<<g, 256, 0, st>>>(w_canon, wd, l.Cin, l.Cout, l.K, l.stride, l.padding, l.dmin, l.Kd);
  else derive_basis_kernel<<<g, 256, 0, st>>>(w_canon, wd, l.Cout, l.Cin);
  g_launches++;
  FV_CUDA(cudaGetLastError());
  return FV_OK;
}

// ---- one layer ---------------------------------------------------------------------------------------
struct LayerCall {
  const float* x = nullptr;
  float* y = nullptr;
  const float* res = nullptr;
  int B = 0;
  long long Lin = 0;
  float pre_slope = -1.f;
  int pad_mode = PAD_ZERO;
  int acc_mode = ACC_STORE;
  float acc_div = 1.f;
  int post_tanh = 0;
  long long x_bs = -1, y_bs = -1, res_bs = -1;  // -1: dense
  bool allow_tc = true;
  const float* x2 = nullptr;   // L_PAIR: second input (the un-activated stack input for the skip layer)
  float pre_slope2 = -1.f;
  const int* lens = nullptr;   // ragged batches: valid input length per utterance (device), or nullptr
  // ConvTranspose / UpsampleLayer only: write the output LeakyReLU(out_slope)-activated in the split activation format
  // (fv_tma.cuh) instead of fp32.  Tensor-core kernel only: run_layer returns FV_NOT_APPLICABLE when it cannot.
  bool out_split = false;
  float out_slope = 0.f;
  // Conv1d on the tensor-core kernel: x (and res) are split-format buffers; x is fetched by TMA (pre_slope must be the slope
  // baked into the copy), the residual is rebuilt from its copy (res_slope = the slope baked into it).
  bool in_split = false, res_split = false;
  float res_slope = 0.f;
  const float* ysum = nullptr;   // split output of the last MRF branch: fp32 running sum of the other branches
  bool x1_split = false;         // L_PAIR: the first input (h) is a split copy fetched by TMA, the second stays fp32
};
constexpr int FV_NOT_APPLICABLE = 1;   // positive: not an error, the caller takes its fallback

static long long layer_out_len(const Layer& l, long long Lin) {
  if (l.type == L_CONV || l.type == L_PAIR) return Lin;
  if (l.type == L_CONVT || l.type == L_UPCONV) return Model::convt_out_len(l, Lin);
  return (Lin + 1) * l.N;  // basis: samples
}

static int run_layer(const Layer& l, const float* wd, const float* bias, const TcLayer* tcl, const LayerCall& c,
                     cudaStream_t st, int* used_tc = nullptr) {
  if (used_tc) *used_tc = 0;
  ConvArgs a{};
  a.x = c.x; a.w = wd; a.bias = bias; a.res = c.res; a.y = c.y;
  a.B = c.B; a.Cin = l.Cin; a.N = l.N; a.Lin = (int)c.Lin; a.K = l.Kd; a.dil = l.dil;
  a.pad_mode = c.pad_mode; a.pre_slope = c.pre_slope;
  a.acc_mode = c.acc_mode; a.acc_div = c.acc_div; a.post_tanh = c.post_tanh;
  a.lens = c.lens;
  a.x_bs = c.x_bs >= 0 ? c.x_bs : (long long)l.Cin * c.Lin;
  long long out_per_b;
  if (l.type == L_CONV) {
    a.Lpos = (int)c.Lin;
    // get_padding (modules.py:186) == ReflectionPad1d((k-1)//2*d); CausalConv1d: everything on the left (modules.py:282,297)
    a.pad_left = l.causal ? (l.K - 1) * l.dil : (l.K - 1) * l.dil / 2;
    a.out_layout = OUT_BCL;
    a.bias_mod = l.Cout;
    out_per_b = (long long)l.Cout * c.Lin;
  } else if (l.type == L_PAIR) {   // two 1x1 convs summed: channels [0,C) from x, [C,2C) from x2
    a.Lpos = (int)c.Lin;
    a.pad_left = 0;
    a.out_layout = OUT_BCL;
    a.bias_mod = l.Cout;
    out_per_b = (long long)l.Cout * c.Lin;
    a.x2 = c.x2;
    a.cin_split = l.Cout;
    a.pre_slope2 = c.pre_slope2;
    a.x_bs = c.x_bs >= 0 ? c.x_bs : (long long)l.Cout * c.Lin;
    a.x2_bs = (long long)l.Cout * c.Lin;
    if (!c.x2) return fail(FV_EINVAL, "pair layer needs two inputs");
  } else if (l.type == L_CONVT) {
    const long long Lout = Model::convt_out_len(l, c.Lin);
    a.Lpos = (int)((Lout - 1 + l.padding) / l.stride + 1);
    a.pad_left = l.Kd - 1;
    a.out_layout = OUT_PHASE;
    a.ph_stride = l.stride; a.ph_pad = l.padding; a.ph_cout = l.Cout; a.ph_lout = (int)Lout;
    a.bias_mod = l.Cout;
    out_per_b = (long long)l.Cout * Lout;
  } else if (l.type == L_UPCONV) {   // nearest stretch + conv, polyphase: t = pos*u + r, taps at x[pos + dmin + jj]
    const long long Lout = Model::convt_out_len(l, c.Lin);
    a.Lpos = (int)((Lout + l.stride - 1) / l.stride);
    a.pad_left = -l.dmin;
    a.out_layout = OUT_PHASE;
    a.ph_stride = l.stride; a.ph_pad = 0; a.ph_cout = l.Cout; a.ph_lout = (int)Lout;
    a.bias_mod = l.Cout;
    out_per_b = (long long)l.Cout * Lout;
  } else {
    a.Lpos = (int)c.Lin + 1;
    a.pad_left = 1;
    a.out_layout = OUT_BLC;
    a.bias_mod = l.N;
    out_per_b = (long long)a.Lpos * l.N;
  }
  a.y_bs = c.y_bs >= 0 ? c.y_bs : out_per_b;
  a.res_bs = c.res_bs >= 0 ? c.res_bs : out_per_b;
  if (c.Lin <= 0 || c.B <= 0) return fail(FV_EINVAL, "empty batch or sequence");
  if (c.pad_mode == PAD_REFLECT && a.pad_left >= c.Lin)
    return fail(FV_EINVAL, "ReflectionPad1d needs pad (%d) < length (%lld)", a.pad_left, c.Lin);
  static const bool tc_disabled = getenv("FV_DISABLE_TC") != nullptr;  // operational kill switch
  // zero-padded-N layers (N < 16) have no residual / accumulate support in the masked epilogue
  const bool padded_ok = !tcl || tcl->n_pad == l.N || (c.res == nullptr && c.acc_mode == ACC_STORE);
  static const bool narrow7_env = getenv("FV_NARROW7") == nullptr || atoi(getenv("FV_NARROW7")) != 0;
  // HBM-bound k = 7 single-channel output convs (conv_post 16->1, LastLayer 32->1): the streaming kernel on both paths
  // (measured 0.305 -> 0.148 ms / 0.457 -> 0.265 ms = 3.5-3.8 TB/s).  The 64->4 conv_post of Multiband-HiFi-GAN is FMA-bound
  // there (0.56 ms) and stays on tcgen05 (0.47 ms) when tensor cores are allowed.
  if (c.out_split || c.in_split || c.res_split || c.x1_split) {
    if (!c.allow_tc || tc_disabled || !tcl || !tcl->eligible || c.lens) return FV_NOT_APPLICABLE;
    if (c.x1_split && l.type != L_PAIR) return FV_NOT_APPLICABLE;
    a.x1_split = c.x1_split ? 1 : 0;
    if (a.out_layout != OUT_PHASE && a.out_layout != OUT_BCL) return FV_NOT_APPLICABLE;
    if (c.res_split && a.out_layout != OUT_BCL) return FV_NOT_APPLICABLE;
    if (c.out_split) a.out_layout = a.out_layout == OUT_PHASE ? OUT_PHASE_SPLIT : OUT_BCL_SPLIT;
    a.ysum = c.ysum;
    a.out_slope = c.out_slope;
    a.x_split = c.in_split ? 1 : 0;
    a.res_split = c.res_split ? 1 : 0;
    a.res_inv_slope = (c.res_split && c.res_slope > 0.f) ? 1.0f / c.res_slope : 0.f;
    int rc = launch_conv_tc2(a, *tcl, st);
    if (rc < 0) return fail(FV_ECUDA, "tcgen05 conv launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == 0 && used_tc) *used_tc = 1;
    return rc == 0 ? FV_OK : FV_NOT_APPLICABLE;
  }
  // FV_NARROW7_MB=1: also the 64 -> 4 conv_post of Multiband-HiFi-GAN on the staged streaming kernel instead of tcgen05 (A/B)
  static const bool narrow7_mb_env = getenv("FV_NARROW7_MB") != nullptr && atoi(getenv("FV_NARROW7_MB")) != 0;
  if (narrow7_env && conv_narrow7_ok(a) &&
      (a.N <= 2 || (narrow7_mb_env && conv_narrow7_staged_ok(a)) || !(c.allow_tc && !tc_disabled && tcl && tcl->eligible))) {
    if (conv_narrow7_staged_ok(a)) FV_CUDA(launch_conv_narrow7_staged(a, st));   // bulk-copy staged streaming form (long rows)
    else FV_CUDA(launch_conv_narrow7(a, st));
    return FV_OK;
  }
  if (c.allow_tc && !tc_disabled && tcl && tcl->eligible && padded_ok) {
    int rc = launch_conv_tc2(a, *tcl, st);
    if (rc == 0) {
      if (used_tc) *used_tc = 1;
      return FV_OK;
    }
    if (rc < 0) return fail(FV_ECUDA, "tcgen05 conv launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    // rc > 0: shape not handled by the tensor-core kernel -> exact fp32 kernel (same GPU, not a CPU fallback)
  }
  FV_CUDA(launch_conv_ffma(a, st));
  return FV_OK;
}

// Would conv_tc2 take this ResBlock conv on the split / TMA activation chain?  (same ConvArgs run_layer builds)
static bool tc2_split_ok(const Layer& l, const TcLayer* t, int nb, long long L, bool with_res, bool out_split) {
  if (!t || !t->eligible || l.type != L_CONV) return false;
  ConvArgs a{};
  a.B = nb; a.Cin = l.Cin; a.N = l.N; a.Lin = (int)L; a.Lpos = (int)L; a.K = l.Kd; a.dil = l.dil;
  a.pad_mode = PAD_ZERO; a.pad_left = (l.K - 1) * l.dil / 2; a.pre_slope = 0.1f;
  a.out_layout = out_split ? OUT_BCL_SPLIT : OUT_BCL; a.out_slope = 0.1f; a.bias_mod = l.Cout;
  a.x_split = 1; a.x_bs = (long long)l.Cin * L; a.y_bs = a.res_bs = (long long)l.Cout * L;
  a.res = with_res ? reinterpret_cast<const float*>(0x10) : nullptr;   // only tested for null-ness by the planner
  a.res_split = with_res ? 1 : 0; a.res_inv_slope = 10.f;
  a.acc_mode = ACC_STORE;
  Tc2Args p{};
  return tc2_plan(a, *t, p, 148);
}

// Would conv_tc2 take this upsample layer (ConvTranspose1d / UpsampleLayer, polyphase) with a split / TMA-fed input?
static bool tc2_up_split_in_ok(const Layer& l, const TcLayer* t, int nb, long long Lin, bool out_split) {
  if (!t || !t->eligible || (l.type != L_CONVT && l.type != L_UPCONV)) return false;
  ConvArgs a{};
  const long long Lout = Model::convt_out_len(l, Lin);
  a.B = nb; a.Cin = l.Cin; a.N = l.N; a.Lin = (int)Lin; a.K = l.Kd; a.dil = l.dil;
  a.Lpos = l.type == L_CONVT ? (int)((Lout - 1 + l.padding) / l.stride + 1) : (int)((Lout + l.stride - 1) / l.stride);
  a.pad_left = l.type == L_CONVT ? l.Kd - 1 : -l.dmin;
  a.pad_mode = PAD_ZERO; a.pre_slope = 0.1f;
  a.out_layout = out_split ? OUT_PHASE_SPLIT : OUT_PHASE; a.out_slope = 0.1f; a.bias_mod = l.Cout;
  a.ph_stride = l.stride; a.ph_pad = l.type == L_CONVT ? l.padding : 0; a.ph_cout = l.Cout; a.ph_lout = (int)Lout;
  a.x_split = 1; a.x_bs = (long long)l.Cin * Lin; a.y_bs = a.res_bs = (long long)l.Cout * Lout;
  a.acc_mode = ACC_STORE;
  Tc2Args p{};
  return a.pad_left >= 0 && tc2_plan(a, *t, p, 148);
}

// ResidualStack on the hybrid split path: would conv_tc2 take the dilated conv with a split (pre-activated) OUTPUT, and the
// fused pair 1x1 with its first input (that h) fetched by TMA?
static bool tc2_stack_split_ok(const Model& m, const Stack& sk, const TcLayer* td, const TcLayer* tp, int nb, long long L,
                               float slope) {
  if (sk.pair < 0 || !td || !tp || !td->eligible || !tp->eligible) return false;
  const Layer& d = m.layers[sk.dil_conv];
  const Layer& pl = m.layers[sk.pair];
  if (d.Cout % 16 || td->n_pad != d.N || !(slope >= 0.f && slope <= 1.f)) return false;
  Tc2Args p{};
  ConvArgs a{};
  a.B = nb; a.Cin = d.Cin; a.N = d.N; a.Lin = (int)L; a.Lpos = (int)L; a.K = d.Kd; a.dil = d.dil;
  a.pad_mode = PAD_REFLECT; a.pad_left = d.causal ? (d.K - 1) * d.dil : (d.K - 1) * d.dil / 2; a.pre_slope = slope;
  a.out_layout = OUT_BCL_SPLIT; a.out_slope = slope; a.bias_mod = d.Cout;
  a.x_bs = (long long)d.Cin * L; a.y_bs = a.res_bs = (long long)d.Cout * L; a.acc_mode = ACC_STORE;
  if (!tc2_plan(a, *td, p, 148)) return false;
  ConvArgs q{};
  q.B = nb; q.Cin = pl.Cin; q.N = pl.N; q.Lin = (int)L; q.Lpos = (int)L; q.K = 1; q.dil = 1;
  q.pad_mode = PAD_ZERO; q.pad_left = 0; q.pre_slope = slope; q.pre_slope2 = -1.f;
  q.out_layout = OUT_BCL; q.bias_mod = pl.Cout; q.cin_split = pl.Cout; q.x1_split = 1;
  q.x2 = reinterpret_cast<const float*>(0x10);
  q.x_bs = q.x2_bs = (long long)pl.Cout * L; q.y_bs = q.res_bs = (long long)pl.Cout * L; q.acc_mode = ACC_STORE;
  Tc2Args p2{};
  return tc2_plan(q, *tp, p2, 148);
}

// ---- whole-model forward ----------------------------------------------------------------------------
struct Bufs {
  float *a, *b, *h, *u0, *u1, *mel_ext;
  size_t each_floats;
};

static size_t max_act_floats(const Model& m, int B, int T) {
  size_t mx = (size_t)B * m.cfg.channels[0] * T;
  for (size_t s = 0; s < m.stages.size(); ++s) {
    size_t v = (size_t)B * m.stages[s].Cout * (size_t)m.len_after(T, (int)s);
    if (v > mx) mx = v;
  }
  if (m.basis >= 0) {
    size_t v = (size_t)B * (size_t)((m.len_after(T, (int)m.stages.size() - 1) + 1) * (m.cfg.basis_L / 2));
    if (v > mx) mx = v;
    v = (size_t)B * m.cfg.out_channels * (size_t)m.len_after(T, (int)m.stages.size() - 1);   // LastLinear output
    if (v > mx) mx = v;
  }
  return (mx + 63) / 64 * 64;
}

static int eff_batch(const Model& m, int B, int flags) {
  return (m.cfg.kind == FV_BASIS_MELGAN && !(flags & FV_FWD_BASIS_INFERENCE)) ? B + 1 : B;
}

struct Profiler {
  struct Rec {
    int layer, used_tc;   // used_tc: 0 fp32 conv, 1 tcgen05 conv, 2 tcgen05 fused ResBlock1 unit (layer = conv1, + conv2),
                          // 3 tcgen05 fused ResidualStack (layer = the dilated conv, + the 1x1 pair)
    long long Lin;
    int B;
    cudaEvent_t e0, e1;
  };
  std::vector<Rec> recs;
};

static size_t lens_floats(int Be) { return ((size_t)(FV_MAX_STAGES + 2) * Be + 63) / 64 * 64; }

static int forward_impl(fv_handle* h, const float* mel, int B, int T, float* out, float* out2, void* ws,
                        size_t ws_bytes, int flags, cudaStream_t st, Profiler* prof = nullptr,
                        const int32_t* lens_host = nullptr) {
  const Model& m = h->model;
  const fv_config& c = m.cfg;
  static const bool tc_disabled_env = getenv("FV_DISABLE_TC") != nullptr;
  static const bool fuse_disabled_env = getenv("FV_NO_FUSE") != nullptr;
  const bool tc_ok = !(flags & FV_FWD_NO_TENSOR_CORES) && !tc_disabled_env && h->tc.buf != nullptr;   // (images dropped: fv_tc_usable)
  const bool fuse_ok = !fuse_disabled_env;
  // MRF sum on the tensor-core path: 1/num_kernels folded into every branch and accumulated with red.global.add (no read of
  // the running sum in the epilogues); the exact-fp32 path keeps the reference's add-then-divide order.  FV_MRF_RED=0: off.
  static const bool mrf_red_env = getenv("FV_MRF_RED") == nullptr || atoi(getenv("FV_MRF_RED")) != 0;
  const bool mrf_red = tc_ok && mrf_red_env;
  const bool pair_ok = !fuse_disabled_env;   // ResidualStack: fuse the two 1x1 convs (both kernels support two inputs)
  const int Be = eff_batch(m, B, flags);
  const float mslope = m.mel_slope();
  const size_t each = max_act_floats(m, Be, T);
  const size_t mel_ext_floats = (Be != B) ? ((size_t)Be * c.in_channels * T + 63) / 64 * 64 : 0;
  const size_t need = (6 * each + mel_ext_floats + (lens_host ? lens_floats(Be) : 0)) * sizeof(float);
  if (ws_bytes < need) return fail(FV_ENOMEM, "workspace too small: have %zu need %zu", ws_bytes, need);
  float* base = (float*)ws;
  float* bufA = base;
  float* bufB = base + each;
  float* bufH = base + 2 * each;
  float* bufU0 = base + 3 * each;
  float* bufU1 = base + 4 * each;
  float* bufY = base + 5 * each;
  float* mel_ext = base + 6 * each;
  // Ragged batch: per-utterance valid lengths at the input (row 0) and after every upsample stage (rows 1..), on the
  // device.  Every kernel treats samples at or beyond its utterance's length as sequence padding, so utterance b
  // gets exactly what a B=1 call with T = lens[b] computes; the tail of each row is computed but meaningless.
  int* lens_dev = nullptr;
  if (lens_host) {
    if (Be != B) return fail(FV_EINVAL, "ragged Basis-MelGAN batches need FV_FWD_BASIS_INFERENCE (one pass, no zero-input subtraction)");
    std::vector<int> hl((m.stages.size() + 1) * (size_t)B);
    for (int b = 0; b < B; ++b) {
      long long l = lens_host[b];
      if (l <= 0 || l > T || (!m.is_hifi() && l <= (c.pre_kernel_size - 1) / 2))
        return fail(FV_EINVAL, "lens[%d] = %lld out of range for T = %d", b, l, T);
      if (c.kind == FV_MELGAN && m.len_after((int)l, (int)m.stages.size() - 1) <= (c.post_kernel_size - 1) / 2)
        return fail(FV_EINVAL, "lens[%d] = %lld too short for the LastLayer reflection pad", b, l);
      hl[b] = (int)l;
      for (size_t s = 0; s < m.stages.size(); ++s) {
        l = Model::convt_out_len(m.layers[m.stages[s].up], l);
        if (l <= 0) return fail(FV_EINVAL, "lens[%d] too short for this architecture", b);
        // ReflectionPad1d of the widest ResidualStack of this stage needs pad < length (the reference raises there)
        for (const Stack& sk : m.stages[s].stacks) {
          const Layer& dc = m.layers[sk.dil_conv];
          const long long pad = dc.causal ? (long long)(dc.K - 1) * dc.dil : (long long)(dc.K - 1) / 2 * dc.dil;
          if (pad >= l)
            return fail(FV_EINVAL, "lens[%d] = %d too short: stage %d has %lld samples but a reflection pad of %lld", b,
                        (int)lens_host[b], (int)s, l, pad);
        }
        hl[(s + 1) * B + b] = (int)l;
      }
    }
    lens_dev = reinterpret_cast<int*>(mel_ext + mel_ext_floats);
    FV_CUDA(cudaMemcpyAsync(lens_dev, hl.data(), hl.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  auto lens_at = [&](int stage_out, int b0) -> const int* {   // stage_out = -1: mel / conv_pre; s: after upsample stage s
    return lens_dev ? lens_dev + (size_t)(stage_out + 1) * B + b0 : nullptr;
  };

  auto wd = [&](int li) { return h->derived + m.layers[li].wd_offset; };
  auto bias = [&](int li) -> const float* {
    if (m.layers[li].type == L_PAIR) return pair_bias_ptr(m.layers[li], h->derived + m.layers[li].wd_offset);
    return m.layers[li].b_param >= 0 ? h->packed + m.params[m.layers[li].b_param].offset : nullptr;
  };
  auto tcl = [&](int li) -> const TcLayer* { return h->tc.layer(li); };
  auto call = [&](int li, LayerCall lc) -> int {
    lc.allow_tc = tc_ok;
    if (!prof) return run_layer(m.layers[li], wd(li), bias(li), tcl(li), lc, st);
    Profiler::Rec r{};
    r.layer = li; r.Lin = lc.Lin; r.B = lc.B;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess)
      return fail(FV_ECUDA, "cudaEventCreate failed");
    cudaEventRecord(r.e0, st);
    int rc = run_layer(m.layers[li], wd(li), bias(li), tcl(li), lc, st, &r.used_tc);
    cudaEventRecord(r.e1, st);
    if (rc == FV_NOT_APPLICABLE) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); return rc; }
    prof->recs.push_back(r);
    return rc;
  };
  auto pack_split = [&](const float* src, float* dst, int nb, int C, long long Ls, float slope) -> int {
    dim3 grid((unsigned)std::min<long long>((Ls + 255) / 256, 64), (unsigned)(C / 8), (unsigned)nb);
    pack_split_kernel<<<grid, 256, 0, st>>>(src, reinterpret_cast<uint4*>(dst), C, (int)Ls, slope);
    g_launches++;
    FV_CUDA(cudaGetLastError());
    return FV_OK;
  };
  // Split (TMA-native) activation path of a HiFi-family stage: every ResBlock1 unit of the stage runs as a fused unit whose
  // input is fetched by TMA from a pre-activated fp16 hi / lo copy.  FV_SPLIT=0 turns it off (A/B and debugging).
  static const bool split_env = getenv("FV_SPLIT") == nullptr || atoi(getenv("FV_SPLIT")) != 0;

  const float* x_in = mel;
  if (Be != B) {  // Basis forward(): append the all-zero utterance (basis_melgan.py:148-151)
    const long long n = (long long)B * c.in_channels * T;
    copy_rows_kernel<<<grid_for(n), 256, 0, st>>>(mel, mel_ext, n);
    g_launches++;
    FV_CUDA(cudaMemsetAsync(mel_ext + n, 0, (size_t)c.in_channels * T * sizeof(float), st));
    x_in = mel_ext;
  }

  int rc;
  long long L = T;
  float* cur = bufA;
  float* other = bufB;
  // Split (TMA-native) activation chain across stages: `cur_split` = the stage input in `cur` is a split copy with LeakyReLU 0.1
  // baked in (written by conv_pre / by the last MRF branch of the previous stage) and the upsample layer fetches it by TMA.
  static const bool split_env0 = getenv("FV_SPLIT") == nullptr || atoi(getenv("FV_SPLIT")) != 0;
  static const bool split_final_env = getenv("FV_SPLIT_FINAL") == nullptr || atoi(getenv("FV_SPLIT_FINAL")) != 0;
  const bool chain_ok = m.is_hifi() && tc_ok && split_env0 && split_final_env && !lens_host && tc3_split_available() &&
                        !m.stages.empty();
  bool cur_split = false;
  {  // conv_pre (hifigan.py:93, zero pad) / ReflectionPad1d + Conv1d (melgan.py:68-71)
    LayerCall lc;
    lc.x = x_in; lc.y = cur; lc.B = Be; lc.Lin = L; lc.lens = lens_at(-1, 0);
    lc.pad_mode = m.is_hifi() ? PAD_ZERO : PAD_REFLECT;
    rc = FV_NOT_APPLICABLE;
    if (chain_ok && m.layers[m.pre].Cout % 16 == 0 &&
        tc2_up_split_in_ok(m.layers[m.stages[0].up], tcl(m.stages[0].up), Be, L, false)) {
      LayerCall ls = lc;
      ls.out_split = true; ls.out_slope = 0.1f;   // the slope of the LeakyReLU in front of ups[0] (hifigan.py:95)
      rc = call(m.pre, ls);
      if (rc == FV_OK) cur_split = true;
    }
    if (rc == FV_NOT_APPLICABLE) rc = call(m.pre, lc);
    if (rc) return rc;
  }
  // Each stage can run in micro-batches of utterances (FV_L2_BUDGET_MB) sized so that the stage's intermediate
  // tensors stay resident in the 126 MB L2.  Utterances are independent, so this is pure scheduling.  Measured on
  // B200 (profiles/r01_notes.md): with one persistent launch per conv the extra ~2000 launches/step cost more
  // than the L2 hits save (35 -> 66 ms), so the default budget is "whole batch"; the knob stays for fused kernels.
  static const long long l2_budget = []() {
    const char* e = getenv("FV_L2_BUDGET_MB");
    return (long long)(e ? atoi(e) : (1 << 20)) * 1024 * 1024;   // default: off (whole batch per launch)
  }();
  for (size_t s = 0; s < m.stages.size(); ++s) {
    const Stage& sg = m.stages[s];
    const Layer& up = m.layers[sg.up];
    const long long Lout = Model::convt_out_len(up, L);
    const long long in_per_utt = (long long)up.Cin * L, out_per_utt = (long long)sg.Cout * Lout;
    long long mb = l2_budget / (4 * out_per_utt * (long long)sizeof(float));
    if (mb < 1) mb = 1;
    if (mb > Be) mb = Be;
    bool stage_fin_split = false;
    for (int b0 = 0; b0 < Be; b0 += (int)mb) {
      const int nb = (int)std::min<long long>(mb, Be - b0);
      const float* x_in_mb = cur + (long long)b0 * in_per_utt;
      float* s_out = other + (long long)b0 * out_per_utt;   // this micro-batch's slice of the stage output
      // split path: all units of the stage are fused ResBlock1 units that the planner accepts in split mode
      bool split_stage = m.is_hifi() && tc_ok && fuse_ok && split_env && !lens_dev && tc3_split_available() &&
                         !sg.branches.empty() && Lout < (1LL << 30);
      // split_wide: the stage is too wide for the fused-unit kernel (C = 128 ...): the same split / TMA activation chain
      // through conv_tc2 — conv1 (split in -> split h), conv2 (split h in, residual from the split x, split or fp32 out)
      static const bool split_wide_env = getenv("FV_SPLIT_WIDE") == nullptr || atoi(getenv("FV_SPLIT_WIDE")) != 0;
      bool split_wide = false;
      for (size_t j = 0; split_stage && j < sg.branches.size(); ++j)
        for (const ResUnit& ru : sg.branches[j].units) {
          Tc3Args probe{};
          const Layer& la = m.layers[ru.c1];
          const TcLayer *t1 = tcl(ru.c1), *t2 = ru.c2 >= 0 ? tcl(ru.c2) : nullptr;
          if (ru.c2 < 0 || !t1 || !t2 || !t1->eligible || !t2->eligible || t1->n_tiles != 1 || t2->n_tiles != 1 ||
              !tc3_plan(nb, la.Cin, (int)Lout, la.K, la.dil, probe, true)) { split_stage = false; break; }
        }
      if (!split_stage && m.is_hifi() && tc_ok && split_env && split_wide_env && !lens_dev && tc3_split_available() &&
          !sg.branches.empty() && Lout < (1LL << 30) && sg.Cout % 16 == 0) {
        split_wide = true;
        for (size_t j = 0; split_wide && j < sg.branches.size(); ++j)
          for (const ResUnit& ru : sg.branches[j].units) {
            const TcLayer *t1 = tcl(ru.c1), *t2 = ru.c2 >= 0 ? tcl(ru.c2) : nullptr;
            if (ru.c2 < 0 || !t1 || !t2 || t1->n_pad != sg.Cout || t2->n_pad != sg.Cout ||
                !tc2_split_ok(m.layers[ru.c1], t1, nb, Lout, false, true) ||
                !tc2_split_ok(m.layers[ru.c2], t2, nb, Lout, true, true) ||
                !tc2_split_ok(m.layers[ru.c2], t2, nb, Lout, true, false)) {
              split_wide = false;
              break;
            }
          }
        split_stage = split_wide;   // the upsample layer writes the split copy either way
      }
      {  // LeakyReLU + ConvTranspose1d
        LayerCall lc;
        lc.x = x_in_mb; lc.y = bufY; lc.B = nb; lc.Lin = L; lc.lens = lens_at((int)s - 1, b0);
        lc.pre_slope = m.is_hifi() ? 0.1f : mslope;  // LRELU_SLOPE modules.py:9 / negative_slope melgan.py:30
        lc.in_split = cur_split;                     // the previous stage / conv_pre left a split copy: TMA-fed, no loader warps
        if (split_stage) {   // the upsample layer's epilogue writes lrelu(y) pre-split for the three branches' first convs
          LayerCall ls = lc;
          ls.out_split = true; ls.out_slope = 0.1f;
          rc = call(sg.up, ls);
          if (rc == FV_NOT_APPLICABLE) {   // not on the tensor-core kernel: fp32 output, then one packing pass
            lc.y = bufH;
            rc = call(sg.up, lc);
            if (rc == FV_NOT_APPLICABLE) return fail(FV_ESTATE, "upsample layer refused the split input it was planned for");
            if (rc) return rc;
            if ((rc = pack_split(bufH, bufY, nb, sg.Cout, Lout, 0.1f))) return rc;
          } else if (rc) return rc;
        } else {
          rc = call(sg.up, lc);
          if (rc == FV_NOT_APPLICABLE) return fail(FV_ESTATE, "upsample layer refused the split input it was planned for");
          if (rc) return rc;
        }
      }
      // fin_split: the last MRF branch emits the stage result itself (running sum + its share, LeakyReLU 0.1, split) into the
      // dead stage-input buffer, so the next stage's upsample layer is TMA-fed too.
      const bool fin_split = chain_ok && split_stage && s + 1 < m.stages.size() && nb == Be && b0 == 0 &&
                             (sg.branches.size() == 1 || mrf_red) &&
                             tc2_up_split_in_ok(m.layers[m.stages[s + 1].up], tcl(m.stages[s + 1].up), nb, Lout, true) &&
                             tc2_up_split_in_ok(m.layers[m.stages[s + 1].up], tcl(m.stages[s + 1].up), nb, Lout, false);
      stage_fin_split = fin_split;
      const int* sl = lens_at((int)s, b0);   // lengths of this stage's tensors
      if (m.is_hifi()) {
        // MRF: xs = sum_j resblock_j(y); x = xs / num_kernels (hifigan.py:97-103); s_out accumulates xs.
        const int nb_br = (int)sg.branches.size();
        for (int j = 0; j < nb_br; ++j) {
          const Branch& br = sg.branches[j];
          const float* bc = bufY;
          const int nu = (int)br.units.size();
          for (int u = 0; u < nu; ++u) {
            const bool last = (u == nu - 1);
            float* dst = last ? s_out : (u % 2 ? bufU1 : bufU0);
            int acc = ACC_STORE;
            float div = 1.f;
            if (last && mrf_red && nb_br > 1) {   // 1/num_kernels folded into every branch, no read of the running sum
              acc = (j == 0) ? ACC_STORE_SCALE : ACC_RED_SCALE; div = (float)nb_br;
            } else if (last && j > 0) { acc = (j == nb_br - 1) ? ACC_ADD_DIV : ACC_ADD; div = (float)nb_br; }
            const bool emit_fin = fin_split && last && j == nb_br - 1;   // this launch writes the stage result (split) into `cur`
            const float* ysum = nullptr;
            if (emit_fin) {
              dst = cur;
              acc = nb_br > 1 ? ACC_STORE_SCALE : ACC_STORE; div = (float)nb_br;
              ysum = nb_br > 1 ? s_out : nullptr;
            }
            if (br.units[u].c2 >= 0 && split_wide) {   // unfused unit on the split / TMA chain (conv_tc2 x 2)
              LayerCall l1;
              l1.x = bc; l1.y = bufH; l1.B = nb; l1.Lin = Lout; l1.pre_slope = 0.1f;
              l1.in_split = true; l1.out_split = true; l1.out_slope = 0.1f;
              rc = call(br.units[u].c1, l1);
              if (rc == FV_NOT_APPLICABLE) return fail(FV_ESTATE, "split-chain conv1 was not applicable after planning");
              if (rc) return rc;
              LayerCall l2;
              l2.x = bufH; l2.y = dst; l2.res = bc; l2.B = nb; l2.Lin = Lout; l2.pre_slope = 0.1f;
              l2.in_split = true; l2.res_split = true; l2.res_slope = 0.1f;
              l2.out_split = !last || emit_fin; l2.out_slope = 0.1f;
              l2.acc_mode = acc; l2.acc_div = div; l2.ysum = ysum;
              rc = call(br.units[u].c2, l2);
              if (rc == FV_NOT_APPLICABLE) return fail(FV_ESTATE, "split-chain conv2 was not applicable after planning");
              if (rc) return rc;
              bc = dst;
              continue;
            }
            if (br.units[u].c2 >= 0) {  // ResBlock1 unit (modules.py:224-229)
              if (tc_ok && fuse_ok && !split_wide) {   // one kernel: conv1 -> lrelu -> conv2 -> +x, h stays in shared memory
                const Layer& la = m.layers[br.units[u].c1];
                const TcLayer* t1 = tcl(br.units[u].c1);
                const TcLayer* t2 = tcl(br.units[u].c2);
                if (t1 && t2) {
                  Profiler::Rec r{};
                  if (prof) {
                    r.layer = br.units[u].c1; r.Lin = Lout; r.B = nb; r.used_tc = 2;
                    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
                    cudaEventRecord(r.e0, st);
                  }
                  const int io = !split_stage ? IO_F32 : ((last && !emit_fin) ? IO_SPLIT_F32 : IO_SPLIT_SPLIT);
                  const int frc = launch_fused_unit(bc, dst, bias(br.units[u].c1), bias(br.units[u].c2), *t1, *t2, nb,
                                                    la.Cin, (int)Lout, la.K, la.dil, 0.1f, acc, div, st, sl, io, ysum);
                  if (prof) {
                    if (frc == 0) { cudaEventRecord(r.e1, st); prof->recs.push_back(r); }
                    else { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
                  }
                  if (frc < 0) return fail(FV_ECUDA, "fused unit launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                  if (frc == 0) { bc = dst; continue; }
                }
              }
              if (split_stage) return fail(FV_ESTATE, "split-path fused unit was not applicable after planning");
              LayerCall l1;
              l1.x = bc; l1.y = bufH; l1.B = nb; l1.Lin = Lout; l1.pre_slope = 0.1f; l1.lens = sl;
              if ((rc = call(br.units[u].c1, l1))) return rc;
              LayerCall l2;
              l2.x = bufH; l2.y = dst; l2.res = bc; l2.B = nb; l2.Lin = Lout; l2.pre_slope = 0.1f; l2.lens = sl;
              l2.acc_mode = acc; l2.acc_div = div;
              if ((rc = call(br.units[u].c2, l2))) return rc;
            } else {  // ResBlock2 unit (modules.py:248-251)
              LayerCall l1;
              l1.x = bc; l1.y = dst; l1.res = bc; l1.B = nb; l1.Lin = Lout; l1.pre_slope = 0.1f; l1.lens = sl;
              l1.acc_mode = acc; l1.acc_div = div;
              if ((rc = call(br.units[u].c1, l1))) return rc;
            }
            bc = dst;
          }
        }
      } else {
        const float* sc_in = bufY;
        const int ns = (int)sg.stacks.size();
        if (ns == 0) {
          FV_CUDA(cudaMemcpyAsync(s_out, bufY, sizeof(float) * (size_t)nb * out_per_utt, cudaMemcpyDeviceToDevice, st));
        }
        for (int k = 0; k < ns; ++k) {  // ResidualStack (modules.py:372-382)
          const Stack& sk = sg.stacks[k];
          float* dst = (k == ns - 1) ? s_out : (k % 2 ? bufY : bufU1);
          // hybrid split path: the dilated conv writes h pre-activated in the split format and the pair layer fetches those
          // chunks by TMA (bit-identical A operand: the loader computed the same lrelu + hi/lo split); FV_STACK_SPLIT=0: off
          static const bool stack_split_env = getenv("FV_STACK_SPLIT") == nullptr || atoi(getenv("FV_STACK_SPLIT")) != 0;
          const bool h_split = pair_ok && tc_ok && split_env0 && stack_split_env && !lens_dev && tc3_split_available() &&
                               tc2_stack_split_ok(m, sk, tcl(sk.dil_conv), sk.pair >= 0 ? tcl(sk.pair) : nullptr, nb, Lout, mslope);
          // One launch for the whole stack where the fused kernel takes it (C <= 64): h stays in shared memory, c is read once
          // (these stacks are HBM-bound as two launches: 640 B -> 256 B per position at C = 32).  FV_STACK_FUSED=0: off.
          static const bool stack_fused_env = getenv("FV_STACK_FUSED") == nullptr || atoi(getenv("FV_STACK_FUSED")) != 0;
          if (pair_ok && tc_ok && fuse_ok && stack_fused_env && sk.pair >= 0) {
            const Layer& d = m.layers[sk.dil_conv];
            const TcLayer* td = tcl(sk.dil_conv);
            const TcLayer* tp = tcl(sk.pair);
            const int pad = (d.K - 1) * d.dil / 2;
            if (td && tp && !d.causal && d.Cin == d.Cout && m.layers[sk.pair].Cout == d.Cout && m.layers[sk.pair].Cin == 2 * d.Cout) {
              if (pad >= Lout) return fail(FV_EINVAL, "ReflectionPad1d needs pad (%d) < length (%lld)", pad, Lout);
              Profiler::Rec r{};
              if (prof) {
                r.layer = sk.dil_conv; r.Lin = Lout; r.B = nb; r.used_tc = 3;
                cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
                cudaEventRecord(r.e0, st);
              }
              const int frc = launch_fused_stack(sc_in, dst, bias(sk.dil_conv), bias(sk.pair), *td, *tp, nb, d.Cout, (int)Lout,
                                                 d.K, d.dil, mslope, true, st, sl);
              if (prof) {
                if (frc == 0) { cudaEventRecord(r.e1, st); prof->recs.push_back(r); }
                else { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
              }
              if (frc < 0) return fail(FV_ECUDA, "fused stack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
              if (frc == 0) { sc_in = dst; continue; }
            }
          }
          LayerCall l1;
          l1.x = sc_in; l1.y = bufH; l1.B = nb; l1.Lin = Lout; l1.pre_slope = mslope; l1.pad_mode = PAD_REFLECT; l1.lens = sl;
          l1.out_split = h_split; l1.out_slope = mslope;
          rc = call(sk.dil_conv, l1);
          if (rc == FV_NOT_APPLICABLE) return fail(FV_ESTATE, "stack conv refused the split output it was planned for");
          if (rc) return rc;
          if (pair_ok && sk.pair >= 0) {   // stack.4(lrelu(h)) + skip_layer(c) as one two-input 1x1 GEMM-conv
            LayerCall lp;
            lp.x = bufH; lp.pre_slope = mslope; lp.x2 = sc_in; lp.pre_slope2 = -1.f;
            lp.y = dst; lp.B = nb; lp.Lin = Lout; lp.lens = sl;
            lp.x1_split = h_split;
            rc = call(sk.pair, lp);
            if (rc == FV_NOT_APPLICABLE) return fail(FV_ESTATE, "pair layer refused the split input it was planned for");
            if (rc) return rc;
          } else {
            LayerCall ls;
            ls.x = sc_in; ls.y = bufU0; ls.B = nb; ls.Lin = Lout; ls.lens = sl;
            if ((rc = call(sk.skip, ls))) return rc;
            LayerCall l2;
            l2.x = bufH; l2.y = dst; l2.res = bufU0; l2.B = nb; l2.Lin = Lout; l2.pre_slope = mslope; l2.lens = sl;
            if ((rc = call(sk.conv1x1, l2))) return rc;
          }
          sc_in = dst;
        }
      }
    }
    L = Lout;
    if (stage_fin_split) {
      cur_split = true;     // `cur` already holds the stage result as a split copy (written by the last MRF branch)
    } else {
      std::swap(cur, other);  // cur = stage output (fp32)
      cur_split = false;
    }
  }

  if (m.is_hifi()) {  // F.leaky_relu(x) [slope 0.01!] -> conv_post -> tanh (hifigan.py:104-106)
    LayerCall lc;
    lc.x = cur; lc.y = out; lc.B = Be; lc.Lin = L; lc.pre_slope = 0.01f; lc.post_tanh = 1;
    lc.lens = lens_at((int)m.stages.size() - 1, 0);
    if ((rc = call(m.post, lc))) return rc;
    if (c.kind == FV_MB_HIFIGAN && out2) {
      if (!h->pqmf_syn) return fail(FV_ESTATE, "PQMF synthesis filter not bound");
      FV_CUDA(launch_pqmf_synthesis(out, h->pqmf_syn, out2, B, c.pqmf_subbands, c.pqmf_taps, (int)L,
                                    lens_at((int)m.stages.size() - 1, 0), st));
      g_launches++;
    }
  } else if (c.kind == FV_MELGAN) {  // LastLayer (modules.py:85-89) + Tanh (melgan.py:109-110)
    LayerCall lc;
    lc.x = cur; lc.y = out; lc.B = Be; lc.Lin = L; lc.pre_slope = mslope; lc.pad_mode = PAD_REFLECT;
    lc.lens = lens_at((int)m.stages.size() - 1, 0);
    lc.post_tanh = c.use_final_activation ? 1 : 0;
    if ((rc = call(m.post, lc))) return rc;
  } else {  // ReLU -> Linear(C->L) -> overlap_and_add(L/2)  (basis_melgan.py:121, modules.py:264-267)
    if (m.ll1 >= 0) {   // LastLinear (modules.py:124-131), eval-mode BatchNorm folded into the two 1x1 convs
      LayerCall l1;
      l1.x = cur; l1.y = other; l1.B = Be; l1.Lin = L; l1.pre_slope = 0.2f; l1.lens = lens_at((int)m.stages.size() - 1, 0);
      if ((rc = call(m.ll1, l1))) return rc;
      LayerCall l2 = l1;
      l2.x = other; l2.y = cur;
      if ((rc = call(m.ll2, l2))) return rc;
    }
    const Layer& bl = m.layers[m.basis];
    const long long full_len = (L + 1) * bl.N;
    LayerCall lc;
    lc.x = cur; lc.B = Be; lc.Lin = L; lc.pre_slope = c.use_final_activation ? 0.f : -1.f;
    lc.lens = lens_at((int)m.stages.size() - 1, 0);
    if (flags & FV_FWD_BASIS_INFERENCE) {
      lc.y = out;
      if ((rc = call(m.basis, lc))) return rc;
    } else {
      lc.y = other;
      if ((rc = call(m.basis, lc))) return rc;
      const long long trunc = L * bl.N;  // [:, :weight.size(1) * (L // 2)]
      dim3 grid(grid_for(trunc), B);
      sub_broadcast_kernel<<<grid, 256, 0, st>>>(other, other + (long long)B * full_len, out, full_len, trunc);
      g_launches++;
      FV_CUDA(cudaGetLastError());
      if (out2) {
        const int C = bl.Cin;
        dim3 g2((unsigned)((L + 31) / 32), (C + 31) / 32, B), b2(32, 8);
        relu_transpose_sub_kernel<<<g2, b2, 0, st>>>(cur, cur + (long long)B * C * L, out2, C, L, c.use_final_activation ? 1 : 0);
        g_launches++;
        FV_CUDA(cudaGetLastError());
      }
    }
  }
  return FV_OK;
}

// temp derived weights for the per-op test entry points
struct TempW {
  float* p = nullptr;
  ~TempW() { if (p) cudaFree(p); }
  int make(const Layer& l, const float* w, cudaStream_t st) {
    FV_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)l.Cin * l.Kd * l.N));
    return derive_layer(l, w, p, st);
  }
};

static Layer make_conv_layer(int Cin, int Cout, int K, int dil) {
  Layer l;
  l.type = L_CONV; l.Cin = Cin; l.Cout = Cout; l.K = K; l.dil = dil; l.N = Cout; l.Kd = K;
  return l;
}

// run one conv with raw canonical weights (test entry points); derives fp32 (and, if asked, fp16 hi/lo) images
static int conv_raw(const Layer& l, const float* w, const float* bias, const LayerCall& lc, int use_tc,
                    cudaStream_t st) {
  TempW tw;
  int rc = tw.make(l, w, st);
  if (rc) return rc;
  TcWeights tcw;
  const TcLayer* tl = nullptr;
  if (use_tc) {
    std::vector<Layer> one{l};
    one[0].wd_offset = 0;
    if ((rc = tcw.build(one, tw.p, st))) return fail(FV_ECUDA, "tc weight build failed");
    tl = tcw.layer(0);
  }
  LayerCall c2 = lc;
  c2.allow_tc = use_tc != 0;
  rc = run_layer(l, tw.p, bias, tl, c2, st);
  if (rc) return rc;
  FV_CUDA(cudaStreamSynchronize(st));  // temp weights are freed on return
  return FV_OK;
}

}  // namespace fv

using namespace fv;

extern "C" {

const char* fv_last_error(void) { return g_err.c_str(); }
int fv_abi_version(void) { return FV_ABI_VERSION; }
int64_t fv_launch_count(void) { return (int64_t)g_launches.load(); }
int64_t fv_tc_launch_count(void) { return (int64_t)g_tc_launches.load(); }

int fv_create(const fv_config* cfg, fv_handle** out) {
  if (!cfg || !out) return fail(FV_EINVAL, "null argument");
  fv_handle* h = new fv_handle();
  if (!h->model.build(*cfg)) {
    int rc = fail(FV_EINVAL, "%s", h->model.err.c_str());
    delete h;
    return rc;
  }
  *out = h;
  return FV_OK;
}

void fv_destroy(fv_handle* h) {
  if (!h) return;
  if (h->derived) cudaFree(h->derived);
  h->tc.release();
  delete h;
}

int fv_num_params(const fv_handle* h) { return h ? (int)h->model.params.size() : FV_EINVAL; }

int fv_param_info(const fv_handle* h, int index, char* name, int name_cap, int64_t shape[4], int* ndim,
                  int64_t* offset_floats) {
  if (!h || index < 0 || index >= (int)h->model.params.size()) return fail(FV_EINVAL, "bad param index");
  const Param& p = h->model.params[index];
  if (name && name_cap > 0) {
    strncpy(name, p.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = p.shape[i];
  if (ndim) *ndim = p.ndim;
  if (offset_floats) *offset_floats = p.offset;
  return FV_OK;
}

int64_t fv_param_total_floats(const fv_handle* h) { return h ? h->model.total_floats : FV_EINVAL; }

int fv_bind_weights(fv_handle* h, const float* packed_dev, int64_t n_floats, const float* pqmf_analysis_dev,
                    const float* pqmf_synthesis_dev, void* stream) {
  if (!h || !packed_dev) return fail(FV_EINVAL, "null argument");
  if (n_floats < h->model.total_floats)
    return fail(FV_EINVAL, "packed weights too short: %lld < %lld", (long long)n_floats,
                (long long)h->model.total_floats);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->derived) { cudaFree(h->derived); h->derived = nullptr; }
  h->tc.release();
  FV_CUDA(cudaMalloc(&h->derived, sizeof(float) * (size_t)h->model.derived_floats));
  for (const Layer& l : h->model.layers) {
    int rc;
    if (l.type == L_PAIR) {
      auto P = [&](int idx) -> const float* { return idx >= 0 ? packed_dev + h->model.params[idx].offset : nullptr; };
      rc = derive_pair(l, P(l.w_param), P(l.w_param2), P(l.b_param), P(l.b_param2), h->derived + l.wd_offset, st);
    } else {
      rc = derive_layer(l, packed_dev + h->model.params[l.w_param].offset, h->derived + l.wd_offset, st);
    }
    if (rc) return rc;
  }
  if (h->tc.build(h->model.layers, h->derived, st)) return fail(FV_ECUDA, "tensor-core weight image build failed");
  h->packed = packed_dev;
  h->pqmf_ana = pqmf_analysis_dev;
  h->pqmf_syn = pqmf_synthesis_dev;
  FV_CUDA(cudaStreamSynchronize(st));
  if (!h->tc.in_range()) {
    // |w| > 65504 (or a non-finite weight): cvt.satfinite would clamp it silently in the fp16 hi/lo images.  Drop them: every
    // layer of this handle then runs on the exact-fp32 kernels (still on the GPU), and fv_tc_usable() reports it.
    h->tc.release();
    fprintf(stderr, "fastvocoder_b200: a weight exceeds the fp16 split range (|w| > 65504 or non-finite); "
                    "this handle runs on the exact-fp32 kernels\n");
  }
  h->bound = true;
  return FV_OK;
}

int fv_tc_usable(const fv_handle* h) { return (h && h->bound && h->tc.buf) ? 1 : 0; }

int fv_out_length(const fv_handle* h, int T, int flags, int64_t* out_len) {
  if (!h || !out_len || T <= 0) return fail(FV_EINVAL, "bad argument");
  *out_len = h->model.out_length(T, flags);
  return FV_OK;
}

int fv_workspace_bytes(const fv_handle* h, int B, int T, size_t* bytes) {
  if (!h || !bytes || B <= 0 || T <= 0) return fail(FV_EINVAL, "bad argument");
  const int Be = h->model.cfg.kind == FV_BASIS_MELGAN ? B + 1 : B;
  const size_t each = max_act_floats(h->model, Be, T);
  const size_t mel_ext = ((size_t)Be * h->model.cfg.in_channels * T + 63) / 64 * 64;
  *bytes = (6 * each + mel_ext + lens_floats(Be)) * sizeof(float);   // incl. room for a ragged batch's length table
  return FV_OK;
}

int fv_forward(fv_handle* h, const float* mel, int B, int T, float* out, float* out2, void* workspace,
               size_t workspace_bytes, int flags, void* stream) {
  if (!h || !mel || !out || !workspace) return fail(FV_EINVAL, "null argument");
  if (!h->bound) return fail(FV_ESTATE, "fv_forward before fv_bind_weights");
  if (B <= 0 || T <= 0) return fail(FV_EINVAL, "B and T must be > 0");
  if (!h->model.is_hifi() && T <= (h->model.cfg.pre_kernel_size - 1) / 2)
    return fail(FV_EINVAL, "ReflectionPad1d needs T > %d", (h->model.cfg.pre_kernel_size - 1) / 2);
  if (h->model.out_length(T, flags) <= 0) return fail(FV_EINVAL, "T too small for this architecture");
  return forward_impl(h, mel, B, T, out, out2, workspace, workspace_bytes, flags, (cudaStream_t)stream);
}

int fv_forward_ragged(fv_handle* h, const float* mel, int B, int T, const int32_t* lens_host, float* out, float* out2,
                      void* workspace, size_t workspace_bytes, int flags, void* stream) {
  if (!h || !mel || !out || !workspace || !lens_host) return fail(FV_EINVAL, "null argument");
  if (!h->bound) return fail(FV_ESTATE, "fv_forward_ragged before fv_bind_weights");
  if (B <= 0 || T <= 0) return fail(FV_EINVAL, "B and T must be > 0");
  if (h->model.out_length(T, flags) <= 0) return fail(FV_EINVAL, "T too small for this architecture");
  return forward_impl(h, mel, B, T, out, out2, workspace, workspace_bytes, flags, (cudaStream_t)stream, nullptr,
                      lens_host);
}

int fv_forward_profile(fv_handle* h, const float* mel, int B, int T, float* out, float* out2, void* workspace,
                       size_t workspace_bytes, int flags, void* stream, fv_profile_entry* entries, int cap,
                       int* count) {
  if (!h || !mel || !out || !workspace || !entries || !count) return fail(FV_EINVAL, "null argument");
  if (!h->bound) return fail(FV_ESTATE, "fv_forward_profile before fv_bind_weights");
  Profiler prof;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = forward_impl(h, mel, B, T, out, out2, workspace, workspace_bytes, flags, st, &prof);
  cudaError_t e = cudaStreamSynchronize(st);
  int n = 0;
  for (auto& r : prof.recs) {
    if (rc == 0 && e == cudaSuccess && n < cap) {
      const Layer& l = h->model.layers[r.layer];
      fv_profile_entry& o = entries[n++];
      memset(&o, 0, sizeof o);
      strncpy(o.name, h->model.params[l.w_param].name.c_str(), sizeof(o.name) - 1);
      o.kernel = r.used_tc;
      o.Cin = l.Cin; o.N = l.N; o.K = l.Kd; o.dil = l.dil;
      o.positions = (int64_t)r.B * r.Lin;
      // algorithmic MACs: every input sample meets every (Cout, tap) of the reference op
      o.flops = 2.0 * (double)l.Cin * l.Cout * l.K * (double)r.Lin * r.B * (r.used_tc == 2 ? 2.0 : 1.0);
      if (r.used_tc == 3) o.flops += 2.0 * 2.0 * (double)l.Cout * l.Cout * (double)r.Lin * r.B;   // + stack.4 and skip_layer (1x1)
      // algorithmic HBM bytes if nothing were cached: read x, write y (+ read residual for half the convs, ignored)
      o.bytes = 4.0 * r.B * ((double)l.Cin * r.Lin + (double)l.Cout * (double)layer_out_len(l, r.Lin) /
                                                         (l.type == L_BASIS ? l.Cout : 1));
      cudaEventElapsedTime(&o.ms, r.e0, r.e1);
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  *count = n;
  if (rc) return rc;
  FV_CUDA(e);
  return FV_OK;
}

int fv_forward_flops(const fv_handle* h, int B, int T, int flags, double* flops) {
  if (!h || !flops) return fail(FV_EINVAL, "null argument");
  *flops = 2.0 * h->model.macs_per_utt(T) * eff_batch(h->model, B, flags);
  return FV_OK;
}

int fv_conv1d(const float* x, const float* w, const float* bias, const float* residual, float* y, int B, int Cin,
              int Cout, int L, int K, int dilation, int pad_mode, float pre_slope, int post_tanh, int use_tc,
              void* stream) {
  if (!x || !w || !y || B <= 0 || Cin <= 0 || Cout <= 0 || L <= 0 || K <= 0 || K % 2 == 0 || dilation <= 0)
    return fail(FV_EINVAL, "fv_conv1d: bad argument (K must be odd: 'same' convolution)");
  Layer l = make_conv_layer(Cin, Cout, K, dilation);
  LayerCall lc;
  lc.x = x; lc.y = y; lc.res = residual; lc.B = B; lc.Lin = L;
  lc.pre_slope = pre_slope; lc.pad_mode = pad_mode; lc.post_tanh = post_tanh;
  return conv_raw(l, w, bias, lc, use_tc, (cudaStream_t)stream);
}

int fv_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                        int Lin, int K, int stride, int padding, int output_padding, float pre_slope, int use_tc,
                        void* stream) {
  if (!x || !w || !y || B <= 0 || Cin <= 0 || Cout <= 0 || Lin <= 0 || K <= 0 || stride <= 0 || K < stride)
    return fail(FV_EINVAL, "fv_conv_transpose1d: bad argument");
  Layer l;
  l.type = L_CONVT; l.Cin = Cin; l.Cout = Cout; l.K = K; l.dil = 1;
  l.stride = stride; l.padding = padding; l.output_padding = output_padding;
  l.N = stride * Cout; l.Kd = (K + stride - 1) / stride;
  if (Model::convt_out_len(l, Lin) <= 0) return fail(FV_EINVAL, "fv_conv_transpose1d: empty output");
  LayerCall lc;
  lc.x = x; lc.y = y; lc.B = B; lc.Lin = Lin; lc.pre_slope = pre_slope;
  return conv_raw(l, w, bias, lc, use_tc, (cudaStream_t)stream);
}

int fv_resblock1(const float* x, const float* const* w1, const float* const* b1, const float* const* w2,
                 const float* const* b2, const int* dilations, int num_dilations, float* y, float* scratch,
                 int B, int C, int L, int K, int use_tc, void* stream) {
  if (!x || !y || !scratch || num_dilations <= 0) return fail(FV_EINVAL, "fv_resblock1: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* hbuf = scratch;                          // [B,C,L]
  float* pp[2] = {scratch + (size_t)B * C * L, y};  // ping-pong so the last unit lands in y
  const float* cur = x;
  if (use_tc == 3) {   // split (TMA-native) chain: pack x once, units hand the split format to each other, the last writes fp32
    if (!tc3_split_available() || C % 8) return fail(FV_EINVAL, "fv_resblock1: split path unavailable");
    dim3 grid((unsigned)std::min((L + 255) / 256, 64), (unsigned)(C / 8), (unsigned)B);
    pack_split_kernel<<<grid, 256, 0, st>>>(x, reinterpret_cast<uint4*>(hbuf), C, L, 0.1f);
    g_launches++;
    FV_CUDA(cudaGetLastError());
    cur = hbuf;
  }
  for (int u = 0; u < num_dilations; ++u) {
    float* dst = pp[(num_dilations - 1 - u) % 2 ? 0 : 1];
    Layer l1 = make_conv_layer(C, C, K, dilations[u]);
    if (use_tc >= 2) {  // fused-unit kernel (conv1 -> lrelu -> conv2 -> +x in one launch)
      const int io = use_tc == 3 ? (u == num_dilations - 1 ? IO_SPLIT_F32 : IO_SPLIT_SPLIT) : IO_F32;
      Layer l2f = make_conv_layer(C, C, K, 1);
      TempW t1, t2;
      int rc = t1.make(l1, w1[u], st);
      if (rc) return rc;
      if ((rc = t2.make(l2f, w2[u], st))) return rc;
      // one contiguous fp32 derived buffer is what TcWeights::build expects: pack the two images side by side
      float* both = nullptr;
      const size_t n1 = (size_t)C * K * C;
      FV_CUDA(cudaMalloc(&both, 2 * n1 * sizeof(float)));
      cudaMemcpyAsync(both, t1.p, n1 * sizeof(float), cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(both + n1, t2.p, n1 * sizeof(float), cudaMemcpyDeviceToDevice, st);
      std::vector<Layer> two{l1, l2f};
      two[0].wd_offset = 0;
      two[1].wd_offset = (int64_t)n1;
      TcWeights tcw;
      int frc = 1;
      if (tcw.build(two, both, st) == 0 && tcw.layer(0) && tcw.layer(1))
        frc = launch_fused_unit(cur, dst, b1 ? b1[u] : nullptr, b2 ? b2[u] : nullptr, *tcw.layer(0), *tcw.layer(1), B, C,
                                L, K, dilations[u], 0.1f, ACC_STORE, 1.f, st, nullptr, io);
      cudaError_t se = cudaStreamSynchronize(st);
      cudaFree(both);
      tcw.release();
      if (frc < 0 || se != cudaSuccess) return fail(FV_ECUDA, "fused unit failed: %s", cudaGetErrorString(se));
      if (frc == 0) { cur = dst; continue; }
      if (use_tc == 3) return fail(FV_EINVAL, "fv_resblock1: shape not handled by the split-path fused unit");
    }
    LayerCall c1;
    c1.x = cur; c1.y = hbuf; c1.B = B; c1.Lin = L; c1.pre_slope = 0.1f;
    int rc = conv_raw(l1, w1[u], b1 ? b1[u] : nullptr, c1, use_tc, st);
    if (rc) return rc;
    Layer l2 = make_conv_layer(C, C, K, 1);
    LayerCall c2;
    c2.x = hbuf; c2.y = dst; c2.res = cur; c2.B = B; c2.Lin = L; c2.pre_slope = 0.1f;
    rc = conv_raw(l2, w2[u], b2 ? b2[u] : nullptr, c2, use_tc, st);
    if (rc) return rc;
    cur = dst;
  }
  return FV_OK;
}

int fv_residual_stack(const float* c, const float* w_dil, const float* b_dil, const float* w_1x1,
                      const float* b_1x1, const float* w_skip, const float* b_skip, float* y, float* scratch,
                      int B, int C, int L, int K, int dilation, int use_tc, void* stream) {
  if (!c || !y || !scratch) return fail(FV_EINVAL, "fv_residual_stack: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* hbuf = scratch;
  float* ubuf = scratch + (size_t)B * C * L;
  Layer ld = make_conv_layer(C, C, K, dilation);
  if (use_tc >= 2) {   // fused ResidualStack kernel (one launch); falls through to the layer-by-layer path when not applicable
    if (B <= 0 || C <= 0 || L <= 0 || K <= 0 || K % 2 == 0 || dilation <= 0) return fail(FV_EINVAL, "fv_residual_stack: bad argument");
    if ((K - 1) * dilation / 2 >= L) return fail(FV_EINVAL, "ReflectionPad1d needs pad (%d) < length (%d)", (K - 1) * dilation / 2, L);
    Layer lp;
    lp.type = L_PAIR; lp.Cin = 2 * C; lp.Cout = C; lp.K = 1; lp.dil = 1; lp.N = C; lp.Kd = 1;
    const size_t n1 = ((size_t)C * K * C + 63) / 64 * 64, n2 = ((size_t)2 * C * C + 63) / 64 * 64;
    float* both = nullptr;
    FV_CUDA(cudaMalloc(&both, (n1 + n2 + 64 + (size_t)C) * sizeof(float)));
    std::vector<Layer> two{ld, lp};
    two[0].wd_offset = 0;
    two[1].wd_offset = (int64_t)n1;
    int rc = derive_layer(ld, w_dil, both, st);
    if (!rc) rc = derive_pair(lp, w_1x1, w_skip, b_1x1, b_skip, both + n1, st);
    TcWeights tcw;
    int frc = 1;
    if (!rc && tcw.build(two, both, st) == 0 && tcw.layer(0) && tcw.layer(1))
      frc = launch_fused_stack(c, y, b_dil, pair_bias_ptr(lp, both + n1), *tcw.layer(0), *tcw.layer(1), B, C, L, K, dilation, 0.2f,
                               true, st);
    cudaError_t se = cudaStreamSynchronize(st);
    cudaFree(both);
    tcw.release();
    if (rc) return rc;
    if (frc < 0 || se != cudaSuccess) return fail(FV_ECUDA, "fused stack failed: %s", cudaGetErrorString(se));
    if (frc == 0) return FV_OK;
    if (use_tc == 3) return fail(FV_EINVAL, "fv_residual_stack: shape not handled by the fused stack kernel");
  }
  LayerCall c1;
  c1.x = c; c1.y = hbuf; c1.B = B; c1.Lin = L; c1.pre_slope = 0.2f; c1.pad_mode = PAD_REFLECT;
  int rc = conv_raw(ld, w_dil, b_dil, c1, use_tc, st);
  if (rc) return rc;
  Layer l1 = make_conv_layer(C, C, 1, 1);
  LayerCall cs;
  cs.x = c; cs.y = ubuf; cs.B = B; cs.Lin = L;
  if ((rc = conv_raw(l1, w_skip, b_skip, cs, use_tc, st))) return rc;
  LayerCall c2;
  c2.x = hbuf; c2.y = y; c2.res = ubuf; c2.B = B; c2.Lin = L; c2.pre_slope = 0.2f;
  return conv_raw(l1, w_1x1, b_1x1, c2, use_tc, st);
}

int fv_basis_signal(fv_handle* h, const float* weight_bcl, int B, int frames, float* out, int use_tc, void* stream) {
  if (!h || !weight_bcl || !out || B <= 0 || frames <= 0) return fail(FV_EINVAL, "fv_basis_signal: bad argument");
  if (!h->bound) return fail(FV_ESTATE, "fv_basis_signal before fv_bind_weights");
  const Model& m = h->model;
  if (m.basis < 0) return fail(FV_EINVAL, "fv_basis_signal: not a Basis-MelGAN handle");
  LayerCall lc;
  lc.x = weight_bcl; lc.y = out; lc.B = B; lc.Lin = frames; lc.pre_slope = -1.f;   // BasisSignalLayer has no activation
  lc.allow_tc = use_tc != 0;
  return run_layer(m.layers[m.basis], h->derived + m.layers[m.basis].wd_offset, nullptr, h->tc.layer(m.basis), lc,
                   (cudaStream_t)stream);
}

int fv_overlap_add(const float* frames, int B, int num_frames, int frame_length, int frame_step, float* out,
                   void* stream) {
  if (!frames || !out || B <= 0 || num_frames <= 0) return fail(FV_EINVAL, "fv_overlap_add: bad argument");
  if (frame_length <= 0 || frame_step <= 0) return fail(FV_EINVAL, "fv_overlap_add: frame_length / frame_step must be > 0");
  if (frame_length == 2 * frame_step) {   // the Basis-MelGAN shape (L = 30, hop = 15): two addends per sample
    dim3 grid(grid_for((long long)(num_frames + 1) * frame_step), B);
    overlap_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, out, num_frames, frame_step);
  } else {                                // any other ratio (the reference's gcd sub-frame path, modules.py:57-72)
    dim3 grid(grid_for((long long)(num_frames - 1) * frame_step + frame_length), B);
    overlap_add_general_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, out, num_frames, frame_length, frame_step);
  }
  g_launches++;
  FV_CUDA(cudaGetLastError());
  return FV_OK;
}

int fv_pqmf_synthesis(const float* x, const float* synthesis_filter, int B, int subbands, int taps, int Lband,
                      float* y, void* stream) {
  if (!x || !synthesis_filter || !y || B <= 0 || subbands <= 0 || taps <= 0 || taps % 2 || Lband <= 0)
    return fail(FV_EINVAL, "fv_pqmf_synthesis: bad argument");
  FV_CUDA(launch_pqmf_synthesis(x, synthesis_filter, y, B, subbands, taps, Lband, nullptr, (cudaStream_t)stream));
  g_launches++;
  return FV_OK;
}

int fv_pqmf_analysis(const float* x, const float* analysis_filter, int B, int subbands, int taps, int L, float* y,
                     void* stream) {
  if (!x || !analysis_filter || !y || B <= 0 || subbands <= 0 || taps <= 0 || taps % 2 || L < subbands)
    return fail(FV_EINVAL, "fv_pqmf_analysis: bad argument");
  const long long Lb = (L - subbands) / subbands + 1;  // conv1d(stride=S, kernel=S) output length
  FV_CUDA(launch_pqmf_analysis(x, analysis_filter, y, B, subbands, taps, (long long)L, Lb, (cudaStream_t)stream));
  g_launches++;
  return FV_OK;
}

int fv_encode_16bits(const float* x, int64_t n, float rescale_out, int16_t* out, float* scratch1, void* stream) {
  if (!x || !out || !scratch1 || n <= 0) return fail(FV_EINVAL, "fv_encode_16bits: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FV_CUDA(cudaMemsetAsync(scratch1, 0, sizeof(float), st));
  absmax_kernel<<<grid_for((n + 3) / 4), 256, 0, st>>>(x, n, (unsigned int*)scratch1);
  encode16_kernel<<<grid_for((n + 3) / 4), 256, 0, st>>>(x, n, (const unsigned int*)scratch1, rescale_out, (short*)out);
  g_launches += 2;
  FV_CUDA(cudaGetLastError());
  return FV_OK;
}

}  // extern "C"
