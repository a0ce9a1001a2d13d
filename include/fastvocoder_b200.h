/*
 * fastvocoder_b200 — C ABI of the B200-native generator forward path.
 *
 * The reference (xcmyz/FastVocoder) has no FFI: its boundary for this path is
 * the Python class API of model/generator/{hifigan,multiband_hifigan,melgan,
 * basis_melgan}.py.  This header is the C boundary the drop-in Python classes
 * (fastvocoder_b200/generators.py) bind with ctypes; each entry point cites the
 * reference interface it replaces.  Plain pointers and sizes only: device
 * pointers are raw `float*` (e.g. torch `tensor.data_ptr()`), streams are raw
 * `cudaStream_t` passed as `void*`.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FV_E* code on failure;
 *     fv_last_error() returns a thread-local message for the last failure.
 *   - activations are fp32, contiguous, reference layout [B, C, L] (time fastest).
 *   - outputs are caller-allocated; nothing returned is owned by the library
 *     except the handle and its derived weight images.
 *   - a handle is immutable after fv_bind_weights(): concurrent fv_forward()
 *     calls on distinct streams with distinct workspaces are safe.
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef FASTVOCODER_B200_H_
#define FASTVOCODER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FV_ABI_VERSION 1

#define FV_MAX_STAGES 8   /* upsample stages */
#define FV_MAX_BRANCH 8   /* MRF branches (resblock_kernel_sizes) */
#define FV_MAX_DIL 8      /* dilations per ResBlock */

/* error codes */
#define FV_OK 0
#define FV_EINVAL (-1)    /* bad argument / unsupported configuration */
#define FV_ECUDA (-2)     /* CUDA runtime error (message has the cudaError string) */
#define FV_ESTATE (-3)    /* call order violated (e.g. forward before bind) */
#define FV_ENOMEM (-4)    /* workspace too small */

/* model kinds — names as accepted by bin/synthesize.py:25-68 (`--model_name`) */
#define FV_HIFIGAN 0            /* "hifigan"            model/generator/hifigan.py:13 */
#define FV_MB_HIFIGAN 1         /* "multiband-hifigan"  model/generator/multiband_hifigan.py:14 */
#define FV_MELGAN 2             /* "melgan"             model/generator/melgan.py:17 */
#define FV_BASIS_MELGAN 3       /* "basis-melgan"       model/generator/basis_melgan.py:19 */

/* fv_forward flags */
#define FV_FWD_DEFAULT 0
#define FV_FWD_BASIS_INFERENCE 1   /* Basis-MelGAN: one pass, no zero-input subtraction, untruncated
                                      (16T+1)*15 samples (basis_melgan.py:196-208) instead of forward()
                                      semantics (basis_melgan.py:140-162) */
#define FV_FWD_NO_TENSOR_CORES 2   /* force the exact-fp32 CUDA-core path for every layer */

/* Architecture description == the reference constructor kwargs (YAML keys of conf/<model>/<size>.yaml). */
typedef struct fv_config {
  int32_t kind;                         /* FV_HIFIGAN ... */
  int32_t in_channels;                  /* 80 mel bins */
  int32_t bias;                         /* conv bias on/off (`bias` kwarg of all four generators) */
  int32_t num_upsamples;                /* len(upsample_rates) / len(upsample_scales) */
  int32_t upsample_rates[FV_MAX_STAGES];
  int32_t upsample_kernel_sizes[FV_MAX_STAGES]; /* HiFi: upsample_kernel_sizes; MelGAN: 2*scale */
  int32_t channels[FV_MAX_STAGES + 1];  /* channels after conv_pre and after each upsample stage */
  int32_t pre_kernel_size;              /* 7 (HiFi conv_pre hifigan.py:26; MelGAN `kernel_size`) */
  int32_t post_kernel_size;             /* 7 (conv_post hifigan.py:52; LastLayer modules.py:76) */
  int32_t out_channels;                 /* HiFi 1, MB 4, MelGAN `out_channels`, Basis = channels[last] */
  /* HiFi family (hifigan.py:14-21) */
  int32_t num_kernels;                  /* len(resblock_kernel_sizes) */
  int32_t resblock_type;                /* 1 (ResBlock1 modules.py:190) or 2 (ResBlock2 modules.py:233) */
  int32_t resblock_kernel_sizes[FV_MAX_BRANCH];
  int32_t resblock_num_dilations[FV_MAX_BRANCH];
  int32_t resblock_dilations[FV_MAX_BRANCH][FV_MAX_DIL];
  /* MelGAN family (melgan.py:20-36, basis_melgan.py:22-42) */
  int32_t stacks;                       /* ResidualStacks per stage, dilation = stack_kernel_size**j */
  int32_t stack_kernel_size;
  int32_t use_final_activation;         /* tanh (MelGAN) / ReLU (Basis) */
  /* Basis-MelGAN */
  int32_t basis_L;                      /* basis length L (30); hop = L/2 */
  /* PQMF (pqmf.py:61): only used by FV_MB_HIFIGAN */
  int32_t pqmf_subbands;                /* 4 */
  int32_t pqmf_taps;                    /* 62 */
  /* non-default architecture switches (YAML keys no shipped config turns on) */
  int32_t upsample_layer;               /* 1: `transposedconv: False` -> UpsampleLayer = nearest-neighbour stretch + Conv1d
                                           (modules.py:160-177; HiFi: k = upsample_kernel_sizes[i], padding k/2,
                                           hifigan.py:32-38; Basis: k = 2*scale+1, padding scale, basis_melgan.py:82-88) */
  int32_t use_causal_conv;              /* 1: MelGAN family `use_causal_conv: True` -> the ResidualStack's dilated conv is
                                           CausalConv1d (modules.py:273-297, 355-361): (k-1)*d samples of padding on the
                                           LEFT only (reflected, the stack's `pad` class) */
  int32_t lastlinear;                   /* 1: Basis-MelGAN `lastlinear: True` -> LastLinear after the last stage
                                           (modules.py:116-132, basis_melgan.py:117-118): LReLU.2 -> BN -> 1x1(C->C) ->
                                           LReLU.2 -> BN -> 1x1(C->out_channels).  Inference (eval-mode) semantics: the
                                           host folds each BatchNorm1d's running statistics into the 1x1 conv that follows
                                           it, so the packed `melgan.N.linear_{1,2}.{weight,bias}` are the folded ones */
  int32_t negative_slope_set;           /* MelGAN family: 1 = `negative_slope` below replaces the default 0.2 of           */
  float negative_slope;                 /* nonlinear_activation_params (melgan.py:30; LeakyReLU of every stack / upsample / */
                                        /* LastLayer; LastLinear keeps its hard-coded 0.2, modules.py:119)                  */
  int32_t reserved[3];
} fv_config;

typedef struct fv_handle fv_handle;

/* ---- diagnostics ------------------------------------------------------------------------------ */
const char* fv_last_error(void);
int fv_abi_version(void);
/* number of kernels this library launched on this process since load (bench `gpu_launches`) */
int64_t fv_launch_count(void);
/* how many of those were tcgen05 (tensor-core) convolution launches */
int64_t fv_tc_launch_count(void);

/* ---- model life cycle (host only until fv_bind_weights) ----------------------------------------
 * fv_create      <- Generator.__init__(**yaml)            hifigan.py:14 melgan.py:20 basis_melgan.py:22
 * fv_param_*     <- Generator.state_dict() key/shape set after remove_weight_norm()  (hifigan.py:58-67)
 * fv_bind_weights<- load_state_dict + .to(device) + remove_weight_norm   bin/synthesize.py:69-71
 *                   `packed_dev` holds every parameter, fp32, in its reference layout
 *                   (Conv1d [Cout,Cin,K]; ConvTranspose1d [Cin,Cout,K]; Linear [out,in]) at the float
 *                   offsets fv_param_info reports.  The library derives its own kernel-side images
 *                   (transposed / phase-split / fp16 hi-lo split) into memory it owns; `packed_dev`
 *                   must stay alive and unchanged while the handle is in use.
 *                   For FV_MB_HIFIGAN pass the PQMF filters too (pqmf_*: [S,1,taps+1], [1,S,taps+1]
 *                   fp32 device pointers as designed host-side in float64, pqmf.py:61-92); else NULL. */
int fv_create(const fv_config* cfg, fv_handle** out);
void fv_destroy(fv_handle* h);
int fv_num_params(const fv_handle* h);
int fv_param_info(const fv_handle* h, int index, char* name, int name_cap, int64_t shape[4], int* ndim,
                  int64_t* offset_floats);
int64_t fv_param_total_floats(const fv_handle* h);
int fv_bind_weights(fv_handle* h, const float* packed_dev, int64_t n_floats, const float* pqmf_analysis_dev,
                    const float* pqmf_synthesis_dev, void* stream);
/* 1 when the bound handle holds tensor-core (fp16 hi/lo split) weight images.  fv_bind_weights drops them, and every layer
 * then runs on the exact-fp32 kernels, when a weight does not fit the split (|w| > 65504 or non-finite: cvt.satfinite would
 * clamp it silently).  Activations are assumed to stay below 65504 in magnitude on the tensor-core path. */
int fv_tc_usable(const fv_handle* h);

/* ---- forward -------------------------------------------------------------------------------------
 * fv_out_length  : samples per utterance of `out` for T mel frames (prod(rates)*T; MB: per-band length;
 *                  Basis forward: 16T*15, Basis inference flag: (16T+1)*15).
 * fv_forward     <- Generator.forward(x[B,80,T])   hifigan.py:92 multiband_hifigan.py:101 melgan.py:125
 *                                                   basis_melgan.py:140 (and .inference via B=1)
 *   mel   [B, in_channels, T] fp32 device
 *   out   HiFi/MelGAN: [B, Lout];  MB: sub-bands [B, 4, Lband];  Basis: est_source [B, Lout]
 *   out2  MB: optional PQMF-synthesised waveform [B, 1, 4*Lband] (multiband_hifigan.py:136), or NULL
 *         Basis forward(): optional `weight - zero_weight` [B, 16T, C] (basis_melgan.py:160), or NULL
 *   workspace: >= fv_workspace_bytes(h, B, T) bytes of device memory, 256-byte aligned. */
int fv_out_length(const fv_handle* h, int T, int flags, int64_t* out_len);
int fv_workspace_bytes(const fv_handle* h, int B, int T, size_t* bytes);
int fv_forward(fv_handle* h, const float* mel, int B, int T, float* out, float* out2, void* workspace,
               size_t workspace_bytes, int flags, void* stream);
/* Ragged batch <- the per-file loop of bin/test.py:123-131 / bin/synthesize.py:74-80 done as ONE launch chain:
 * B utterances of lens_host[b] <= T mel frames, packed in mel [B, in_channels, T] (contents at or beyond lens_host[b]
 * are ignored).  Utterance b receives exactly what fv_forward(B = 1, T = lens_host[b]) computes for it: the first
 * fv_out_length(lens_host[b]) samples of its row of `out` (rows are fv_out_length(T) long; the tail is unspecified).
 * lens_host is a HOST array (copied to the device on `stream`).  Basis-MelGAN: FV_FWD_BASIS_INFERENCE semantics only. */
int fv_forward_ragged(fv_handle* h, const float* mel, int B, int T, const int32_t* lens_host, float* out, float* out2,
                      void* workspace, size_t workspace_bytes, int flags, void* stream);
/* Same as fv_forward, but brackets every learned layer with CUDA events on `stream` and reports, per layer
 * launch: which kernel ran (0 = fp32 CUDA-core conv, 1 = tcgen05 conv), its algorithmic FLOPs / bytes and its
 * device time.  Synchronises the stream.  bench.py uses it for the live roofline numbers. */
typedef struct fv_profile_entry {
  char name[64];       /* parameter name of the layer's weight, e.g. "resblocks.2.convs1.1.weight" */
  int32_t kernel;      /* 0 fp32 FFMA conv, 1 tcgen05 split-fp16 conv, 2 tcgen05 fused ResBlock1 unit (conv1+conv2) */
  int32_t Cin, N, K, dil;
  int64_t positions;   /* B * input length */
  double flops;        /* 2 * MAC of the reference op */
  double bytes;        /* fp32 input + output bytes of the op (no reuse assumed) */
  float ms;            /* device time between the two events */
  float reserved;
} fv_profile_entry;
int fv_forward_profile(fv_handle* h, const float* mel, int B, int T, float* out, float* out2, void* workspace,
                       size_t workspace_bytes, int flags, void* stream, fv_profile_entry* entries, int cap,
                       int* count);
/* algorithmic FLOPs (2*MAC, dense conv math as the reference executes it) of one fv_forward(B,T) */
int fv_forward_flops(const fv_handle* h, int B, int T, int flags, double* flops);

/* ---- per-op entry points (unit parity tests; reference layouts, weights passed raw) ---------------
 * All pointers are device pointers.  `pad_mode`: 0 zero, 1 reflect.  `pre_slope` < 0 disables the
 * LeakyReLU applied to the input before the convolution (0 = ReLU).  `residual` may be NULL.
 * `use_tc` != 0 routes through the tcgen05 path when the shape is eligible (fv_resblock1: 2 = fused-unit kernel,
 * 3 = fused units chained through the TMA-fed split fp16 hi/lo activation format; fv_residual_stack: 2 = the fused
 * ResidualStack kernel when the shape is eligible, else layer by layer, 3 = the fused kernel or FV_EINVAL).
 *
 * fv_conv1d            <- torch.nn.Conv1d call sites (modules.py:193-220,364,366,377; hifigan.py:93,105)
 * fv_conv_transpose1d  <- torch.nn.ConvTranspose1d call sites (hifigan.py:39-44, melgan.py:77-85)
 * fv_resblock1         <- ResBlock1.forward        modules.py:223-230
 * fv_residual_stack    <- ResidualStack.forward    modules.py:372-382
 * fv_overlap_add       <- overlap_and_add          modules.py:34-73 (any frame_length / frame_step; out [B, (frames-1)*step + length])
 * fv_pqmf_synthesis    <- PQMF.synthesis           pqmf.py:121-135
 * fv_pqmf_analysis     <- PQMF.analysis            pqmf.py:108-119
 * fv_encode_16bits     <- data/audio.py:12-14      (save_wav's peak-normalise + int16 cast) */
int fv_conv1d(const float* x, const float* w, const float* bias, const float* residual, float* y, int B, int Cin,
              int Cout, int L, int K, int dilation, int pad_mode, float pre_slope, int post_tanh, int use_tc,
              void* stream);
int fv_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                        int Lin, int K, int stride, int padding, int output_padding, float pre_slope, int use_tc,
                        void* stream);
int fv_resblock1(const float* x, const float* const* w1, const float* const* b1, const float* const* w2,
                 const float* const* b2, const int* dilations, int num_dilations, float* y, float* scratch,
                 int B, int C, int L, int K, int use_tc, void* stream);
int fv_residual_stack(const float* c, const float* w_dil, const float* b_dil, const float* w_1x1,
                      const float* b_1x1, const float* w_skip, const float* b_skip, float* y, float* scratch,
                      int B, int C, int L, int K, int dilation, int use_tc, void* stream);
/* BasisMelGANGenerator.test(weight) / BasisSignalLayer.forward   basis_melgan.py:210-212, modules.py:255-267
 *   weight_bcl [B, C, frames] (channels first — the Python binding transposes the reference's [B, frames, C]),
 *   out [B, (frames+1)*hop]: Linear(C -> L, no bias, no activation) + overlap_and_add(hop = L/2) of the bound handle. */
int fv_basis_signal(fv_handle* h, const float* weight_bcl, int B, int frames, float* out, int use_tc, void* stream);
int fv_overlap_add(const float* frames, int B, int num_frames, int frame_length, int frame_step, float* out,
                   void* stream);
int fv_pqmf_synthesis(const float* x, const float* synthesis_filter, int B, int subbands, int taps, int Lband,
                      float* y, void* stream);
int fv_pqmf_analysis(const float* x, const float* analysis_filter, int B, int subbands, int taps, int L, float* y,
                     void* stream);
int fv_encode_16bits(const float* x, int64_t n, float rescale_out, int16_t* out, float* scratch1, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTVOCODER_B200_H_ */
