// Host-side model description: fv_config -> parameter table + layer graph.
// Pure C++ (no CUDA): testable without a GPU.
//
// Parameter names are exactly the reference's state_dict keys after
// remove_weight_norm() (hifigan.py:58-67), e.g. "conv_pre.weight",
// "ups.0.weight", "resblocks.3.convs1.2.bias", "melgan.4.stack.2.weight",
// "melgan.4.skip_layer.bias", "melgan.22.conv.weight",
// "basis_signal.layer.weight".
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/fastvocoder_b200.h"

namespace fv {

struct Param {
  std::string name;
  int ndim = 0;
  int64_t shape[4] = {0, 0, 0, 0};
  int64_t offset = 0;  // in floats, 64-float (256 B) aligned
  int64_t numel() const {
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    return n;
  }
};

enum LayerType { L_CONV = 0, L_CONVT = 1, L_BASIS = 2, L_PAIR = 3, L_UPCONV = 4 };

// One learned layer as the kernels see it (a stride-1 "GEMM conv" over time):
//   y[n, pos] = bias[n % bias_mod] + sum_{ci, j} Wd[ci][j][n] * act(x)[ci, pos - pad_left + j*dil]
// Conv1d      : N = Cout,        Kd = K,            positions = L
// ConvTranspose1d (polyphase): N = stride*Cout (n = r*Cout + co), Kd = ceil(K/stride),
//               positions i' = (t+p)/stride, output sample t = i'*stride + r - p
// Basis linear + overlap-add (L = 2*hop): N = hop, Kd = 2, positions = frames+1, output time-major
// UpsampleLayer (nearest stretch by u, then Conv1d(k, padding p), modules.py:160-177) in polyphase form: output sample
//               t = pos*u + r reads xu[t - p + j] = x[pos + floor((r - p + j)/u)], so N = u*Cout (n = r*Cout + co), the
//               taps that land on the same input sample are pre-summed:  Wd[ci][jj][n] = sum_{j: floor((r-p+j)/u) = dmin+jj} w[co][ci][j],
//               dmin = floor(-p/u), Kd = floor((u-1-p+k-1)/u) - dmin + 1, pad_left = -dmin
struct Layer {
  LayerType type = L_CONV;
  int w_param = -1, b_param = -1;
  int w_param2 = -1, b_param2 = -1;  // L_PAIR: second 1x1 conv (the skip layer) fused along the input-channel axis
  int Cin = 0, Cout = 0, K = 0, dil = 1;
  int stride = 1, padding = 0, output_padding = 0;  // ConvTranspose1d; UpsampleLayer: stride = stretch u, padding = conv padding
  int dmin = 0;                                     // UpsampleLayer: smallest input offset (<= 0)
  bool causal = false;                              // CausalConv1d: all (K-1)*dil samples of padding on the left
  // derived GEMM view
  int N = 0, Kd = 0;
  int64_t wd_offset = 0;  // float offset of the [Cin][Kd][N] image in the derived buffer
};

struct ResUnit {  // ResBlock1 unit (conv1 dilated, conv2) or ResBlock2 unit (conv only; c2 = -1)
  int c1 = -1, c2 = -1;
};
struct Branch {
  int kernel_size = 0;
  std::vector<ResUnit> units;
};
struct Stack {  // MelGAN ResidualStack
  int dil_conv = -1, conv1x1 = -1, skip = -1, dilation = 1;
  int pair = -1;  // fused  W_1x1 * lrelu(h) + W_skip * c + (b_1x1 + b_skip): one two-input 1x1 GEMM-conv (Cin = 2C)
};
struct Stage {
  int up = -1;
  int Cout = 0;
  std::vector<Branch> branches;  // HiFi family
  std::vector<Stack> stacks;     // MelGAN family
};

struct Model {
  fv_config cfg{};
  std::vector<Param> params;
  int64_t total_floats = 0;
  std::vector<Layer> layers;
  int64_t derived_floats = 0;
  int pre = -1, post = -1, basis = -1;
  int ll1 = -1, ll2 = -1;   // Basis-MelGAN LastLinear (BatchNorm folded on the host): two 1x1 convs after the last stage
  std::vector<Stage> stages;
  std::string err;

  bool is_hifi() const { return cfg.kind == FV_HIFIGAN || cfg.kind == FV_MB_HIFIGAN; }
  // LeakyReLU slope of the MelGAN family (nonlinear_activation_params["negative_slope"], melgan.py:30; default 0.2)
  float mel_slope() const { return cfg.negative_slope_set ? cfg.negative_slope : 0.2f; }

  int add_param(const std::string& name, std::initializer_list<int64_t> shape) {
    Param p;
    p.name = name;
    p.ndim = (int)shape.size();
    int i = 0;
    for (auto s : shape) p.shape[i++] = s;
    p.offset = total_floats;
    total_floats += (p.numel() + 63) / 64 * 64;
    params.push_back(p);
    return (int)params.size() - 1;
  }

  int add_conv(const std::string& prefix, int Cin, int Cout, int K, int dil, bool bias) {
    Layer l;
    l.type = L_CONV;
    l.Cin = Cin; l.Cout = Cout; l.K = K; l.dil = dil;
    if (bias) l.b_param = add_param(prefix + ".bias", {Cout});
    l.w_param = add_param(prefix + ".weight", {Cout, Cin, K});
    l.N = Cout; l.Kd = K;
    return push_layer(l);
  }
  int add_convt(const std::string& prefix, int Cin, int Cout, int K, int stride, bool bias) {
    Layer l;
    l.type = L_CONVT;
    l.Cin = Cin; l.Cout = Cout; l.K = K; l.dil = 1;
    l.stride = stride;
    l.padding = stride / 2 + stride % 2;  // hifigan.py:42, melgan.py:83
    l.output_padding = stride % 2;
    if (bias) l.b_param = add_param(prefix + ".bias", {Cout});
    l.w_param = add_param(prefix + ".weight", {Cin, Cout, K});
    l.N = stride * Cout;
    l.Kd = (K + stride - 1) / stride;
    return push_layer(l);
  }
  int add_basis(const std::string& name, int C, int L) {
    Layer l;
    l.type = L_BASIS;
    l.Cin = C; l.Cout = L; l.K = 1;
    l.w_param = add_param(name, {L, C});
    l.N = L / 2; l.Kd = 2;
    return push_layer(l);
  }
  int push_layer(Layer& l) {
    l.wd_offset = derived_floats;
    derived_floats += ((int64_t)l.Cin * l.Kd * l.N + 63) / 64 * 64;
    layers.push_back(l);
    return (int)layers.size() - 1;
  }

  static int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
  int add_upconv(const std::string& prefix, int Cin, int Cout, int K, int u, int pad, bool bias) {
    Layer l;
    l.type = L_UPCONV;
    l.Cin = Cin; l.Cout = Cout; l.K = K; l.dil = 1;
    l.stride = u; l.padding = pad;
    if (bias) l.b_param = add_param(prefix + ".conv.bias", {Cout});
    l.w_param = add_param(prefix + ".conv.weight", {Cout, Cin, K});
    l.N = u * Cout;
    l.dmin = floor_div(-pad, u);
    l.Kd = floor_div(u - 1 - pad + K - 1, u) - l.dmin + 1;
    return push_layer(l);
  }
  // output length of an upsampling layer (ConvTranspose1d or UpsampleLayer)
  static int64_t convt_out_len(const Layer& l, int64_t Lin) {
    if (l.type == L_UPCONV) return Lin * l.stride + 2 * l.padding - l.K + 1;
    return (Lin - 1) * l.stride - 2 * l.padding + l.K + l.output_padding;
  }
  // time length after stage s (s = -1: after conv_pre)
  int64_t len_after(int T, int s) const {
    int64_t L = T;
    for (int i = 0; i <= s; ++i) L = convt_out_len(layers[stages[i].up], L);
    return L;
  }

  bool build(const fv_config& c) {
    cfg = c;
    char buf[256];
    auto fail = [&](const char* m) { err = m; return false; };
    if (c.kind < FV_HIFIGAN || c.kind > FV_BASIS_MELGAN) return fail("unknown model kind (no model find!)");
    if (c.in_channels <= 0) return fail("in_channels must be > 0");
    if (c.num_upsamples <= 0 || c.num_upsamples > FV_MAX_STAGES) return fail("num_upsamples out of range");
    for (int i = 0; i <= c.num_upsamples; ++i)
      if (c.channels[i] <= 0) return fail("channels[] must be > 0");
    for (int i = 0; i < c.num_upsamples; ++i) {
      if (c.upsample_rates[i] <= 0 || c.upsample_kernel_sizes[i] <= 0) return fail("bad upsample rate/kernel");
      if (!c.upsample_layer && c.upsample_kernel_sizes[i] < c.upsample_rates[i]) return fail("upsample kernel < rate unsupported");
    }
    if (c.pre_kernel_size <= 0 || c.pre_kernel_size % 2 == 0) return fail("Not support even number kernel size.");
    if (c.negative_slope_set && !(c.negative_slope >= 0.f && c.negative_slope <= 1.f))
      return fail("negative_slope must lie in [0, 1]");
    if (c.use_causal_conv && is_hifi()) return fail("use_causal_conv is a MelGAN-family switch");
    if (c.lastlinear && c.kind != FV_BASIS_MELGAN) return fail("lastlinear is a BasisMelGANGenerator option (basis_melgan.py:38)");
    if (c.upsample_layer && c.kind == FV_MELGAN) return fail("transposedconv=False is not a MelGANGenerator option (melgan.py:20-36)");
    const bool bias = c.bias != 0;
    if (is_hifi()) {
      if (c.num_kernels <= 0 || c.num_kernels > FV_MAX_BRANCH) return fail("num_kernels out of range");
      if (c.resblock_type != 1 && c.resblock_type != 2) return fail("resblock_type must be 1 or 2");
      if (c.post_kernel_size <= 0 || c.post_kernel_size % 2 == 0) return fail("post kernel must be odd");
      pre = add_conv("conv_pre", c.in_channels, c.channels[0], c.pre_kernel_size, 1, bias);
      // parameter order mirrors the reference module order: conv_pre, ups.*, resblocks.*, conv_post
      std::vector<int> ups;
      for (int i = 0; i < c.num_upsamples; ++i) {
        snprintf(buf, sizeof buf, "ups.%d", i);
        if (c.upsample_layer)   // hifigan.py:32-38: UpsampleLayer(.., upsample_rate=u, kernel_size=k, stride=1, padding=k//2)
          ups.push_back(add_upconv(buf, c.channels[i], c.channels[i + 1], c.upsample_kernel_sizes[i], c.upsample_rates[i],
                                   c.upsample_kernel_sizes[i] / 2, bias));
        else
          ups.push_back(add_convt(buf, c.channels[i], c.channels[i + 1], c.upsample_kernel_sizes[i],
                                  c.upsample_rates[i], bias));
      }
      for (int i = 0; i < c.num_upsamples; ++i) {
        Stage st;
        st.up = ups[i];
        st.Cout = c.channels[i + 1];
        for (int j = 0; j < c.num_kernels; ++j) {
          Branch br;
          br.kernel_size = c.resblock_kernel_sizes[j];
          if (br.kernel_size <= 0 || br.kernel_size % 2 == 0) return fail("resblock kernel must be odd");
          int nd = c.resblock_num_dilations[j];
          if (nd <= 0 || nd > FV_MAX_DIL) return fail("resblock dilation count out of range");
          for (int u = 0; u < nd; ++u) {
            ResUnit ru;
            int d = c.resblock_dilations[j][u];
            if (d <= 0) return fail("dilation must be > 0");
            if (c.resblock_type == 1) {
              snprintf(buf, sizeof buf, "resblocks.%d.convs1.%d", i * c.num_kernels + j, u);
              ru.c1 = add_conv(buf, st.Cout, st.Cout, br.kernel_size, d, bias);
            } else {
              snprintf(buf, sizeof buf, "resblocks.%d.convs.%d", i * c.num_kernels + j, u);
              ru.c1 = add_conv(buf, st.Cout, st.Cout, br.kernel_size, d, bias);
            }
            br.units.push_back(ru);
          }
          if (c.resblock_type == 1) {
            for (int u = 0; u < nd; ++u) {
              snprintf(buf, sizeof buf, "resblocks.%d.convs2.%d", i * c.num_kernels + j, u);
              br.units[u].c2 = add_conv(buf, st.Cout, st.Cout, br.kernel_size, 1, bias);
            }
          }
          st.branches.push_back(br);
        }
        stages.push_back(st);
      }
      post = add_conv("conv_post", c.channels[c.num_upsamples], c.out_channels, c.post_kernel_size, 1, bias);
      if (c.kind == FV_MB_HIFIGAN) {
        if (c.pqmf_subbands != c.out_channels) return fail("pqmf_subbands must equal out_channels");
        if (c.pqmf_taps <= 0 || c.pqmf_taps % 2) return fail("The number of taps mush be even number.");
      }
    } else {
      if (c.stacks < 0 || c.stack_kernel_size <= 0 || (!c.use_causal_conv && (c.stack_kernel_size - 1) % 2 != 0))
        return fail("Not support even number kernel size.");
      int idx = 1;
      snprintf(buf, sizeof buf, "melgan.%d", idx);
      pre = add_conv(buf, c.in_channels, c.channels[0], c.pre_kernel_size, 1, bias);
      idx = 2;
      for (int i = 0; i < c.num_upsamples; ++i) {
        Stage st;
        st.Cout = c.channels[i + 1];
        snprintf(buf, sizeof buf, "melgan.%d", idx + 1);
        if (c.upsample_layer)   // basis_melgan.py:82-88: UpsampleLayer(.., kernel_size=2*scale+1, stride=1, padding=scale)
          st.up = add_upconv(buf, c.channels[i], c.channels[i + 1], 2 * c.upsample_rates[i] + 1, c.upsample_rates[i],
                             c.upsample_rates[i], bias);
        else
          st.up = add_convt(buf, c.channels[i], c.channels[i + 1], c.upsample_kernel_sizes[i],
                            c.upsample_rates[i], bias);
        idx += 2;
        int dil = 1;
        for (int j = 0; j < c.stacks; ++j) {
          Stack sk;
          sk.dilation = dil;
          // non-causal: stack = [act, pad, conv, act, conv1x1] -> stack.2 / stack.4;  causal: [act, CausalConv1d, act,
          // conv1x1] -> stack.1.conv / stack.3 (modules.py:345-361)
          snprintf(buf, sizeof buf, c.use_causal_conv ? "melgan.%d.stack.1.conv" : "melgan.%d.stack.2", idx);
          sk.dil_conv = add_conv(buf, st.Cout, st.Cout, c.stack_kernel_size, dil, bias);
          layers[sk.dil_conv].causal = c.use_causal_conv != 0;
          snprintf(buf, sizeof buf, c.use_causal_conv ? "melgan.%d.stack.3" : "melgan.%d.stack.4", idx);
          sk.conv1x1 = add_conv(buf, st.Cout, st.Cout, 1, 1, bias);
          snprintf(buf, sizeof buf, "melgan.%d.skip_layer", idx);
          sk.skip = add_conv(buf, st.Cout, st.Cout, 1, 1, bias);
          {  // virtual fused layer over the two 1x1 convs (no parameters of its own)
            Layer pl;
            pl.type = L_PAIR;
            pl.Cin = 2 * st.Cout; pl.Cout = st.Cout; pl.K = 1; pl.dil = 1;
            pl.w_param = layers[sk.conv1x1].w_param; pl.b_param = layers[sk.conv1x1].b_param;
            pl.w_param2 = layers[sk.skip].w_param; pl.b_param2 = layers[sk.skip].b_param;
            pl.N = st.Cout; pl.Kd = 1;
            pl.wd_offset = derived_floats;
            derived_floats += ((int64_t)pl.Cin * pl.N + 63) / 64 * 64 + (pl.N + 63) / 64 * 64;  // image + summed bias
            layers.push_back(pl);
            sk.pair = (int)layers.size() - 1;
          }
          st.stacks.push_back(sk);
          dil *= c.stack_kernel_size;
          idx += 1;
        }
        stages.push_back(st);
      }
      if (c.kind == FV_MELGAN) {
        if (c.post_kernel_size <= 0 || c.post_kernel_size % 2 == 0) return fail("post kernel must be odd");
        snprintf(buf, sizeof buf, "melgan.%d.conv", idx);
        post = add_conv(buf, c.channels[c.num_upsamples], c.out_channels, c.post_kernel_size, 1, bias);
      } else {
        if (c.basis_L <= 0 || c.basis_L % 2) return fail("basis L must be even (hop = L/2)");
        const int Cl = c.channels[c.num_upsamples];
        if (c.lastlinear) {   // basis_melgan.py:117-118, modules.py:116-132
          if (c.out_channels <= 0) return fail("out_channels must be > 0");
          if (!bias) return fail("lastlinear with bias=False is not supported (the folded BatchNorm needs the bias slot)");
          snprintf(buf, sizeof buf, "melgan.%d.linear_1", idx);
          ll1 = add_conv(buf, Cl, Cl, 1, 1, true);
          snprintf(buf, sizeof buf, "melgan.%d.linear_2", idx);
          ll2 = add_conv(buf, Cl, c.out_channels, 1, 1, true);
        } else if (c.out_channels != Cl) {
          return fail("basis out_channels != channels[-1]");
        }
        basis = add_basis("basis_signal.layer.weight", c.out_channels, c.basis_L);
      }
    }
    return true;
  }

  // samples per utterance written to `out` by fv_forward
  int64_t out_length(int T, int flags) const {
    int64_t L = len_after(T, (int)stages.size() - 1);
    if (cfg.kind == FV_BASIS_MELGAN) {
      int hop = cfg.basis_L / 2;
      return (flags & FV_FWD_BASIS_INFERENCE) ? (L + 1) * hop : L * hop;
    }
    return L;
  }

  // dense-conv MACs of one pass over one utterance of T frames (SURVEY.md §8a accounting)
  double macs_per_utt(int T) const {
    double m = 0;
    auto conv = [&](const Layer& l, double L) { return (double)l.Cout * l.Cin * l.K * L; };
    m += conv(layers[pre], T);
    double L = T;
    for (size_t s = 0; s < stages.size(); ++s) {
      const Layer& up = layers[stages[s].up];
      const double Lup = (double)convt_out_len(up, (int64_t)L);
      // ConvTranspose1d: every input sample meets every tap; UpsampleLayer: a dense conv on the stretched signal
      m += (double)up.Cin * up.Cout * up.K * (up.type == L_UPCONV ? Lup : L);
      L = Lup;
      for (auto& br : stages[s].branches)
        for (auto& u : br.units) {
          m += conv(layers[u.c1], L);
          if (u.c2 >= 0) m += conv(layers[u.c2], L);
        }
      for (auto& sk : stages[s].stacks)
        m += conv(layers[sk.dil_conv], L) + conv(layers[sk.conv1x1], L) + conv(layers[sk.skip], L);
    }
    if (post >= 0) m += conv(layers[post], L);
    if (ll1 >= 0) m += conv(layers[ll1], L) + conv(layers[ll2], L);
    if (basis >= 0) m += (double)layers[basis].Cin * layers[basis].Cout * L;
    return m;
  }
};

}  // namespace fv
