/*
 * libw2s_b200 - C ABI of the B200-native wav2sleep forward path.
 *
 * The reference (joncarter1/wav2sleep) has no FFI: its seam is the Python class contract of
 * wav2sleep.models.wav2sleep.Wav2Sleep (models/wav2sleep.py:16-80).  This header is the boundary a binding
 * (ctypes / cffi / pybind) attaches to; every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every pointer is a DEVICE pointer unless stated otherwise; descriptors (structs) live on the host.
 *   - functions never allocate device memory and never synchronise: the caller passes workspaces sized by
 *     the matching *_workspace_bytes() call and a cudaStream_t (as void*; NULL = legacy default stream).
 *   - return value 0 = ok, non-zero = error; w2s_last_error() gives a thread-local message.
 *   - activations are fp16, channels-last ([B, L, C]); accumulation and statistics are fp32.
 */
#ifndef W2S_B200_H_
#define W2S_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2S_MAX_BLOCKS 12
#define W2S_MAX_MIXER_LAYERS 8
#define W2S_MAX_SIGNALS 4
#define W2S_MAX_SEQ_BLOCKS 4
#define W2S_MAX_DILATIONS 8

/* ABI version of this header; bumped on any struct change. */
int w2s_abi_version(void);
/* Thread-local message of the last failing call ("" if none). */
const char* w2s_last_error(void);

/* ---------------------------------------------------------------------------------------------------------
 * Weight packing (host passes fp32 parameters exactly as they sit in the reference state_dict).
 * ------------------------------------------------------------------------------------------------------- */

/* nn.Conv1d weight [cout, cin, taps] fp32 -> fp16 UMMA operand layout [taps][cin/8][cout][8].
 * Also used for nn.Linear(4*C -> F) of SignalEncoder (models/wav2sleep.py:230,264) by viewing its
 * weight [F, 4*C] as [F, taps=4, C] -> pass taps_major = 1 (input index = tap*cin + c). */
int w2s_pack_conv_weight(const float* w, int cout, int cin, int taps, int taps_major, int split, void* out_fp16,
                         void* stream);
size_t w2s_packed_conv_weight_bytes(int cout, int cin, int taps, int split);
/* 1 if the (cin, cout) encoder conv kernels carry operands as fp16 hi + fp16 lo pairs; their weights must then be
 * packed with split = 1 (hi block followed by lo block, twice the bytes).  True for cin <= 32 and cout <= 32. */
int w2s_conv_uses_split(int cin, int cout);

/* nn.Linear weight [n, k] fp32 -> fp16 mma.sync B-fragment order [n/8][k/16][32 lanes][4] (epoch mixer). */
int w2s_pack_linear_frag(const float* w, int n, int k, void* out_fp16, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * One generic implicit-GEMM conv layer (kernel-level entry, used by the tests and by the stage calls below).
 * Replaces ConvLayer1D.forward (models/blocks.py:173-186) + the consumer-side InstanceNorm/GELU of its input.
 * ------------------------------------------------------------------------------------------------------- */
enum { W2S_PRO_NONE = 0, W2S_PRO_NORM = 1, W2S_PRO_NORM_RES = 2 };
enum { W2S_EPI_STATS = 0, W2S_EPI_BIAS_GELU = 1, W2S_EPI_LN_GELU = 2, W2S_EPI_LN_GELU_RES = 3 };

typedef struct w2s_conv_call {
  int32_t cin, cout, taps, stride, dilation, pad;
  int32_t prologue, epilogue, has_ds;
  int32_t B, L_in, L_out;
  const void* in;         /* fp16 [B, L_in, cin] */
  const void* in_res;     /* fp16 [B, L_in, cin]      (W2S_PRO_NORM_RES) */
  const double* in_stats; /* [B, cin, 2] sum, sumsq   (W2S_PRO_NORM*)    */
  const void* w;          /* packed fp16 */
  const void* w_ds;       /* packed fp16 1x1          (has_ds) */
  void* out;              /* fp16 [B, L_out, cout] */
  void* out_ds;           /* fp16 [B, L_out/2, cout]  (has_ds) */
  double* out_stats;      /* [B, cout, 2], pre-zeroed (W2S_EPI_STATS) */
  const uint8_t* row_mask;/* [B] or NULL */
  const float* bias;      /* [cout] (W2S_EPI_BIAS_GELU) */
  const float* ln_w;      /* [cout] (W2S_EPI_LN_*) */
  const float* ln_b;
  const void* res;        /* fp16 [B, L_out, cout] (W2S_EPI_LN_GELU_RES) */
  const float* head_w;    /* [n_classes, cout] or NULL */
  const float* head_b;
  float* logits;          /* [B, L_out, n_classes] */
  int32_t n_classes;
  float in_eps, ln_eps;
} w2s_conv_call;

int w2s_conv1d_fwd(const w2s_conv_call* call, void* stream);
/* Kernel selection for the encoder convs (testing / A-B measurement): 0 = auto (persistent warp-specialised
 * streaming kernel, conv_stream.cuh), 1 = tile-per-CTA kernel only (conv_igemm.cuh).  Same results either way. */
int w2s_set_conv_impl(int impl);

/* ---------------------------------------------------------------------------------------------------------
 * Stage 1: signal encoder.  Replaces SignalEncoders.forward for one signal (models/wav2sleep.py:146-161)
 * = SignalEncoder.forward non-causal branch (:235-267) = Sequential(ConvBlock1D) (models/blocks.py:57-71)
 * with whole-night InstanceNorm1d(eps=1e-2) (models/utils.py:89-92) + Linear(4C->F) + GELU.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct w2s_encoder_desc {
  int32_t n_blocks;                         /* log2(samples_per_epoch) - 2 */
  int32_t channels[W2S_MAX_BLOCKS];         /* out channels of block i: min(16 * 2^(i/2), 128) */
  int32_t feature_dim;                      /* 128 */
  float norm_eps;                           /* 1e-2 */
  const float* w_first;                     /* fp32 [16, 3]   cnn.0.conv1.conv.weight */
  const float* w_first_ds;                  /* fp32 [16]      cnn.0.downsample.weight */
  const void* w_conv[W2S_MAX_BLOCKS][3];    /* packed fp16 conv1..conv3 ([0][0] unused) */
  const void* w_ds[W2S_MAX_BLOCKS];         /* packed fp16 downsample ([0] unused) */
  const void* w_lin;                        /* packed fp16 linear.weight as taps=4 */
  const float* b_lin;                       /* fp32 [feature_dim] */
} w2s_encoder_desc;

/* keep_activations = 0: inference, tensors rotate through a few slots; 1: every layer output is kept
 * (layout = w2s_encoder_layout) for a backward pass. */
size_t w2s_encoder_workspace_bytes(const w2s_encoder_desc* d, int B, int64_t T, int keep_activations);

/* x: fp32 [B, T] raw (z-scored) signal, rows of -inf = missing signal (data/dataset.py:170-173).
 * z_out: fp16 [B, T / samples_per_epoch, feature_dim]; rows of masked samples are left untouched.
 * row_mask: [B] written with 1 for missing samples. */
int w2s_encoder_fwd(const w2s_encoder_desc* d, const float* x, int B, int64_t T, void* workspace,
                    size_t workspace_bytes, int keep_activations, void* z_out, uint8_t* row_mask, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Stage 2: epoch mixer.  Replaces MultiModalAttentionEmbedder.forward (models/wav2sleep.py:301-346).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct w2s_mixer_layer {
  const void* in_w;   /* frag-packed in_proj_weight [384,128] */
  const void* out_w;  /* frag-packed out_proj.weight [128,128] */
  const void* ff1_w;  /* frag-packed linear1.weight [512,128] */
  const void* ff2_w;  /* frag-packed linear2.weight [128,512] */
  const float *in_b, *out_b, *ff1_b, *ff2_b;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
} w2s_mixer_layer;

typedef struct w2s_mixer_desc {
  int32_t n_layers;
  int32_t feature_dim; /* 128 */
  int32_t n_heads;     /* 8   */
  int32_t dim_ff;      /* 512 */
  float ln_eps;        /* 1e-5 */
  const float* cls;    /* fp32 [128] = register_tokens[0,0,:,0] */
  w2s_mixer_layer layer[W2S_MAX_MIXER_LAYERS];
} w2s_mixer_desc;

/* z[i]: fp16 [B*S, 128] features of signal i (signals sorted by name, wav2sleep.py:311);
 * row_mask[i]: [B] 1 = signal i missing for that sample (may be NULL = all present);
 * out: fp16 [B*S, 128] CLS token after the last layer. */
int w2s_epoch_mixer_fwd(const w2s_mixer_desc* d, const void* const* z, const uint8_t* const* row_mask, int n_signals,
                        int B, int S, void* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Stage 3: sequence mixer + classifier.  Replaces SequenceCNN.forward (models/wav2sleep.py:379-390) =
 * DilatedConvBlock x n (models/blocks.py:115-126) with ConvLayerNorm (models/utils.py:9-23) and
 * Wav2Sleep.classifier (models/wav2sleep.py:41,66).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct w2s_seq_desc {
  int32_t n_blocks;      /* num_layers (2) */
  int32_t n_dilations;   /* 6 -> dilations 1..32 */
  int32_t kernel_size;   /* 7 */
  int32_t feature_dim;   /* 128 */
  int32_t n_classes;
  float ln_eps;          /* 1e-5 */
  const void* w[W2S_MAX_SEQ_BLOCKS][W2S_MAX_DILATIONS];     /* packed fp16 conv weights */
  const float* ln_w[W2S_MAX_SEQ_BLOCKS][W2S_MAX_DILATIONS]; /* fp32 [128] */
  const float* ln_b[W2S_MAX_SEQ_BLOCKS][W2S_MAX_DILATIONS];
  const float* head_w;   /* fp32 [n_classes, 128] classifier.weight */
  const float* head_b;   /* fp32 [n_classes] */
} w2s_seq_desc;

size_t w2s_seqmixer_workspace_bytes(const w2s_seq_desc* d, int B, int S, int keep_activations);

/* x: fp16 [B, S, 128] epoch-mixer output; feat_out: fp16 [B, S, 128] (may be NULL -> workspace);
 * logits: fp32 [B, S, n_classes]. */
int w2s_seqmixer_head_fwd(const w2s_seq_desc* d, const void* x, int B, int S, void* workspace, size_t workspace_bytes,
                          int keep_activations, void* feat_out, float* logits, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py).  Not part of the data path.
 * ------------------------------------------------------------------------------------------------------- */
/* Number of kernels this library has launched in this process (monotonic). */
long long w2s_launch_count(void);
/* on = 1: every later launch is bracketed by CUDA events on its stream and recorded; on = 0: stop and clear. */
int w2s_profile_enable(int on);
int w2s_profile_count(void);
/* Record i: label (host buffer), elapsed ms (synchronises on the record's end event), algorithmic bytes / flops. */
int w2s_profile_get(int i, char* label, int label_cap, float* ms, double* bytes, double* flops);

/* argmax over classes (Wav2Sleep.predict, models/wav2sleep.py:69-80): logits fp32 [n, c] -> int64 [n]. */
int w2s_argmax(const float* logits, int64_t n, int n_classes, int64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* W2S_B200_H_ */
