/*
 * aclip_b200.h -- C ABI of the B200-native AnomalyCLIP inference hot path.
 *
 * The reference (lucazanella/AnomalyCLIP) is pure Python and has no FFI of its own; these entry
 * points are what its modules bind instead of their torch.nn bodies:
 *
 *   aclip_vit_forward        <- VisionTransformer.forward      src/models/components/clip/model.py:266-290
 *   aclip_temporal_forward   <- AnomalyCLIP.forward(test_mode) src/models/components/anomaly_clip.py:132-154
 *                               (SelectorModel.forward          selector_model.py:32-69,
 *                                TemporalModel.forward          temporal_model.py:42-77,
 *                                ClassificationHead.forward     classification_head.py:11-15,
 *                                softmax(sim)*score             src/models/anomaly_clip_module.py:473-477)
 *
 * plus the building-block operators (split, layer norm, GEMM, attention ...) those two are made
 * of, exported so that each one can be parity-tested on its own.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name ends in _host. The caller owns all memory,
 *     including the workspace; the library allocates nothing on the device and keeps no per-call state.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - Return value: ACLIP_OK (0) or a negative AclipStatus; aclip_last_error() returns the text
 *     of the calling thread's last failure. Nothing throws across this boundary.
 *   - "split" tensors carry an fp32 value as two bf16 planes (hi, lo): hi = bf16(x),
 *     lo = bf16(x - hi); plane 0 is hi, plane 1 is lo, `plane_stride` elements apart.
 *   - "f16f8" tensors (GEMM passes = 2) carry an fp32 value v as three planes in the same 4 bytes
 *     per element: H = fp16(v * 2^e_main) at byte 0, L = e4m3((v * 2^e_main - H) * 2^e_res) at byte
 *     2 * plane_stride, C = e4m3(v * 2^e_coarse) at byte 3 * plane_stride.  The GEMM accumulates
 *     x_H w_H (fp16 MMA) + x_L w_C + x_C w_L (e4m3 MMAs at twice the rate) = 2^(ex+ew) x w, i.e. two
 *     pass-equivalents instead of the three bf16 passes at the same ~1e-5 end-to-end error.
 *     Activations use (e_main, e_res, e_coarse) = (4, 7, 0); a weight tensor packed with exponent
 *     ew uses (ew, 4, ew - 7) with ew chosen so that max|w| * 2^ew lies in (2^14, 2^15].
 *   - "f16" tensors (GEMM passes = 4, out_enc = 2) are the H plane alone: a plain fp16 matrix
 *     [rows][ld] holding fp16(v * 2^4); weights reuse the H plane of their f16f8 pack.  One MMA
 *     pass per product, ~2.5e-4 relative on the ViT-B/16 features instead of ~1e-5: the host
 *     selects it per checkpoint after a calibration against the f16f8 mode.
 *   - The fp16 conversions saturate; aclip_saturation_count() reports how many threads stored a
 *     value at or beyond the fp16 range since the last reset (0 on a healthy run).
 */
#ifndef ACLIP_B200_H_
#define ACLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum AclipStatus {
  ACLIP_OK = 0,
  ACLIP_ERR_INVALID = -1, /* bad argument (shape, alignment, null pointer) */
  ACLIP_ERR_CUDA = -2,    /* a CUDA runtime/driver call failed              */
  ACLIP_ERR_WORKSPACE = -3 /* workspace too small                           */
} AclipStatus;

typedef enum AclipAct { ACLIP_ACT_NONE = 0, ACLIP_ACT_QUICKGELU = 1, ACLIP_ACT_LEAKYRELU = 2 } AclipAct;

int aclip_version(void);
const char* aclip_last_error(void);
/* Number of kernels this library has launched in the calling process (all streams). */
long long aclip_launch_count(void);
/* Add n to the launch counter: a caller that replays a captured CUDA graph holding n kernels of
 * this library reports them here (the launchers only run at capture time). Returns the new count. */
long long aclip_note_launches(long long n);
/* Threads that stored an activation beyond the fp16 range of the f16f8 / f16 encodings on the
 * CURRENT device since the last reset (synchronises the device); reset != 0 clears the counter. */
long long aclip_saturation_count(int reset);

/* Optional device timing of every kernel this library launches (cudaEvent pairs on the launch
 * stream), grouped by kernel kind, with the ALGORITHMIC flops / bytes of the launches: bench.py
 * derives its roofline numbers from this.  Enable, run, then collect (collect synchronises on
 * the recorded events and resets the table).  Returns the number of rows written. */
typedef struct AclipTimingRow {
  char name[32];
  long long launches;
  double ms, flops, bytes;
} AclipTimingRow;
int aclip_timing_enable(int on);
int aclip_timing_collect(AclipTimingRow* rows, int max_rows);

/* ------------------------------------------------------------------------------------------
 * Building-block operators
 * ---------------------------------------------------------------------------------------- */

/* fp32 [rows][cols] (pitch ld_in) -> split bf16 [2][rows][ld_out]; columns cols..ld_out-1 are
 * zero-filled so that the result can feed a GEMM whose K is padded. */
int aclip_split_f32(const float* in, long long rows, int cols, int ld_in, void* out_split,
                    int ld_out, long long plane_stride, void* stream);

/* fp32 [rows][cols] -> f16f8 planes [rows][ld_out] (ld_out and plane_stride multiples of 16). */
int aclip_encode_f16f8(const float* in, long long rows, int cols, int ld_in, void* out,
                       int ld_out, long long plane_stride, int e_main, int e_res, int e_coarse,
                       void* stream);

/* fp32 [rows][cols] -> an f16mx tensor [rows][ld_out] (ld_out % 64 == 0): fp16 plane of
 * v * 2^e_main, two MXFP4 planes (residual of the fp16 plane; coarse copy: e2m1 elements, one
 * UE8M0 scale per 32 values along a row) and the scale bytes in the chunked layout the block-scaled
 * tensor-core instruction reads (csrc/mx.cuh).  `out` holds aclip_f16mx_bytes(rows, ld_out) bytes
 * and must be zero-filled by the caller before the first use (scale bytes of rows that do not
 * exist stay 0).  Operand of aclip_gemm with passes = 7. */
long long aclip_f16mx_bytes(long long rows, int ld);
int aclip_encode_f16mx(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out,
                       int e_main, void* stream);

/* (x - centroid) of fp32 feature rows [rows][D] -> split-bf16 rows of pitch ld_out, regrouped from
 * the caller's "(b n s l)" order to sub-video order "(b s) n l" (temporal_model.py:46-53);
 * n = s = l = 1 keeps the order.  Replaces the two centroid subtractions of the reference
 * (selector_model.py:54, anomaly_clip.py:143). */
int aclip_center_regroup(const float* feats, long long rows, int D, const float* centroid,
                         int num_segments, int segment_size, int seg_length, void* out_split,
                         int ld_out, long long plane_stride, void* stream);

/* Frames (B,3,R,R), fp32 normalised or uint8 (then ToTensor + Normalize run here), -> im2col rows
 * [2][B*(R/P)^2][3*P*P] (split-bf16) for the patch-embedding GEMM (clip/model.py:246-252,267). */
int aclip_patchify(const void* frames, int frames_are_u8, int B, int R, int P,
                   const float* mean3_host, const float* std3_host, void* out_split,
                   long long plane_stride, int out_enc, void* stream);

/* Pillow-exact bicubic resize + centre crop of decoded frames (H x W x 3 uint8) to planar
 * (3, size, size) uint8: the reference's GroupScale(224, BICUBIC) + GroupCenterCrop(224)
 * (src/utils/augmentations.py:25-29) without the per-frame PIL calls.  The tap tables (device
 * int32: bounds [size][2] = first source index, tap count; coeffs [size][k] 22-bit fixed point)
 * come from the host plan (anomalyclip_b200/data.py:resize_crop_plan); vertical bounds are relative
 * to row0.  tmp: [num_frames][rows][size][3] scratch.  Bit-exact with Pillow. */
int aclip_resize_crop_u8(const uint8_t* frames_hwc, int num_frames, int H, int W, int row0, int rows,
                         int size, const int* hbounds, const int* hcoeffs, int hk,
                         const int* vbounds, const int* vcoeffs, int vk, uint8_t* tmp,
                         uint8_t* out_chw, void* stream);

struct AclipPeerGather;   /* defined below: fused all-gather over NVLink peer memory */

typedef struct AclipGemmArgs {
  /* operands (bf16 split planes) */
  const void* a;  /* linear: [planes][M][lda]; conv3x3: [planes][S][H][W][C]            */
  const void* w;  /* [planes][N][ldw]  (nn.Linear weight layout: out_features x in_features) */
  int M, N, K;    /* C[M,N] = A[M,K] * W[N,K]^T ; conv3x3: M = S*H*W, K = 9*C            */
  int lda, ldw;   /* row pitches in elements, multiples of 8                              */
  long long a_plane_stride, w_plane_stride; /* elements between hi and lo plane          */
  int passes;     /* 3 = split-bf16 (fp32-faithful), 1 = plain bf16 (hi plane only),
                     2 = f16f8 operands (fp32-faithful, CTA-pair kernel, N % 256 == 0),
                     4 = fp16 operands, one pass,
                     6 = f16f8 operands without the weight-residual cross term (1.5
                         pass-equivalents; the weights are then effectively fp16)            */
  int a_mode;     /* 0 linear, 1 conv3x3 (zero padding 1, stride 1)                      */
  int conv_c, conv_h, conv_w, conv_s; /* conv3x3: channels, grid height, width, images  */
  /* epilogue: out = act(acc + bias) + residual */
  const float* bias;     /* [N] or NULL */
  const float* residual; /* fp32 [*][ldr] or NULL */
  int res_mod;           /* >0: residual row = m % res_mod; 0: residual row = output row */
  int ldr;
  int act;               /* AclipAct */
  float* out_f32;        /* fp32 [*][ldc] or NULL */
  void* out_split;       /* bf16 [2][*][ldc] or NULL */
  long long split_plane_stride;
  int ldc;               /* pitch of out_f32 (and of out_split when ld_split == 0) */
  int ld_split;          /* pitch of out_split, 0 = ldc */
  /* output row remap (0,0,0 = identity):
   * out_row = (m / row_group) * row_group_stride + (m % row_group) + row_offset */
  int row_group, row_group_stride, row_offset;
  int max_ctas;          /* 0 = one persistent CTA per SM */
  int kernel;            /* 0 = auto, 1 = single-CTA tiles (128 x N), 2 = CTA-pair tiles (256 x 256) */
  float out_scale;       /* passes = 2: 2^-(e_act + e_weight) applied to the accumulator; 0 = 1 */
  int out_enc;           /* encoding of out_split: 0 = bf16 hi/lo, 1 = f16f8 activation planes,
                            2 = fp16 plane [*][ld_split] */
  /* optional fused all-gather of the fp32 output (NULL = off): every row stored to out_f32 also
   * goes to row (rank * rows_per_rank + gather_row0 + m) of every rank's gathered buffer
   * (gather->width must equal ldc); gather_signal != 0 publishes this rank's flag when the launch
   * has stored everything (set it on the launch that completes the rank's block). */
  const struct AclipPeerGather* gather;
  int gather_signal;
  long long gather_row0;
  int tile;              /* single-CTA kernel: 0 = auto, 1 = 128-row tiles, 2 = 64 x 32 tiles (the
                            small-problem tile; passes 3 / 4, no gather); results do not depend on it */
} AclipGemmArgs;

/* tcgen05 / TMA GEMM with fused epilogue. N must be a multiple of 32, K a multiple of 8. */
int aclip_gemm(const AclipGemmArgs* args, void* stream);

/* out_enc (here and below): encoding of the split output, 0 = bf16 hi/lo planes, 1 = f16f8
 * activation planes (the A operand of a passes = 2 GEMM), 2 = fp16 plane (passes = 4).
 * nn.LayerNorm (mode 0, clip/model.py:174-180) or axial_attention's ChanLayerNorm (mode 1:
 * (x-mean)/(std+eps)) over rows of D fp32 values; writes fp32 and/or split-bf16 rows. */
int aclip_layernorm(const float* x, long long rows, int D, long long ldx, const float* gamma,
                    const float* beta, float eps, int mode, float* out_f32, long long ld_f32,
                    void* out_split, long long ld_split, long long plane_stride, int out_enc,
                    void* stream);

/* softmax(Q K^T / 8) V for B frames of L tokens, `heads` heads of 64 dims; qkv_split is the
 * split-bf16 [2][B*L][ld_in] output of the in_proj GEMM (q | k | v), out_split [2][B*L][ld_out].
 * Replaces nn.MultiheadAttention inside ResidualAttentionBlock.attention (clip/model.py:206-212).
 * out_enc = 2: qkv_split and out_split are fp16 matrices (the "f16" encoding) and every product is
 * one MMA pass.  kernel: 0 or 2 = the tcgen05/TMEM kernel (the only one). */
int aclip_vit_attention(const void* qkv_split, long long in_plane_stride, int ld_in, int B, int L,
                        int heads, void* out_split, long long out_plane_stride, int ld_out,
                        int kernel, int out_enc, void* stream);

/* Axial self-attention over fp32 qkv rows [sub_videos*n*l][3E] in sub-video order; axis 0 = along
 * the n segments, axis 1 = along the l frames.  Output split-bf16 [2][rows][E].
 * Replaces axial_attention.SelfAttention under PermuteToFrom (temporal_model.py:32-39,64). */
int aclip_axial_attention(const float* qkv, long long sub_videos, int n, int l, int E, int heads,
                          int axis, void* out_split, long long plane_stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * The two hot-path entry points.  Weight structs hold DEVICE pointers to tensors the caller
 * packed once per checkpoint ("split" = bf16 [2][rows][cols], planes contiguous); the structs
 * themselves and the per-layer arrays they point to live in HOST memory.
 * ---------------------------------------------------------------------------------------- */

typedef struct AclipVitBlock {      /* one ResidualAttentionBlock, clip/model.py:188-217 */
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const void* qkv_w;  const float* qkv_b;   /* attn.in_proj_weight  split [2][3W][W]  */
  const void* out_w;  const float* out_b;   /* attn.out_proj.weight split [2][W][W]   */
  const void* fc_w;   const float* fc_b;    /* mlp.c_fc.weight      split [2][4W][W]  */
  const void* proj_w; const float* proj_b;  /* mlp.c_proj.weight    split [2][W][4W]  */
  /* passes = 2 only: qkv_w, fc_w and proj_w are f16f8 planes and *_s = 2^-(4 + e_weight) is the
   * accumulator scale of that GEMM (4 = the activations' e_main); out_w stays bf16 hi/lo (the
   * attention output feeds out_proj in that form) and out_s is unused.  Ignored otherwise. */
  float qkv_s, out_s, fc_s, proj_s;
  /* passes = 4 only: attn.out_proj.weight as f16f8 / fp16 planes (its fp16 plane is read) with
   * accumulator scale out_s; NULL otherwise. */
  const void* out_w16;
  /* passes = 7 only: mlp.c_fc.weight / mlp.c_proj.weight as f16mx tensors (aclip_encode_f16mx with
   * the same per-tensor exponent as the f16f8 pack, so fc_s / proj_s apply); NULL otherwise. */
  const void* fc_wmx;
  const void* proj_wmx;
} AclipVitBlock;

typedef struct AclipVitWeights {    /* VisionTransformer, clip/model.py:233-264 */
  int width, layers, heads, patch, resolution, output_dim;
  const void* conv1_w;                /* conv1.weight.reshape(W, 3*P*P) split [2][W][3PP] */
  const float* class_embedding;       /* [W] */
  const float* positional_embedding;  /* [(R/P)^2 + 1][W] */
  const float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
  const void* proj_w;                 /* proj^T split [2][output_dim][W] */
  const AclipVitBlock* blocks;        /* host array [layers] */
  float conv1_s, proj_s;              /* passes = 2: accumulator scales of conv1_w / proj_w */
} AclipVitWeights;

size_t aclip_vit_workspace_bytes(const AclipVitWeights* w, int micro_batch);

/* features_out[num_frames][output_dim] = VisionTransformer.forward(frames)  (clip/model.py:266-290).
 * frames: (num_frames,3,R,R) fp32 already normalised, or uint8 (frames_are_u8 != 0), in which case
 * ToTensor + Normalize(mean3_host, std3_host) (src/utils/augmentations.py:21-34) run on the GPU.
 * Frames are processed in micro-batches of `micro_batch` through `workspace`
 * (>= aclip_vit_workspace_bytes(w, micro_batch) bytes, 1024-byte aligned).
 * passes = 3: split-bf16 GEMMs (fp32-faithful, the parity mode); passes = 1: plain bf16 GEMMs;
 * passes = 2: f16f8 operands (fp32-faithful at two pass-equivalents; the weights in `w` must have
 * been packed with aclip_encode_f16f8 and width / output_dim must be multiples of 256);
 * passes = 4: fp16 operands end to end, one pass per product (same packed weights plus out_w16);
 * passes = 5: "mixed" -- in_proj, attention and out_proj as passes = 4, the MLP pair, the patch
 * embedding and the output projection as passes = 2 (~1e-4 on the features);
 * passes = 6: as 5, with c_proj issued without its weight-residual cross term (~2e-4);
 * passes = 7: as 5, with the MLP pair on f16mx operands (fp16 main product + two MXFP4 cross terms:
 * 1.5 instead of 2 pass-equivalents at the same ~1e-4; width must be a multiple of 768). */
int aclip_vit_forward(const AclipVitWeights* w, const void* frames, int frames_are_u8,
                      long long num_frames, int micro_batch, const float* mean3_host,
                      const float* std3_host, float* features_out, void* workspace,
                      size_t workspace_bytes, int passes, void* stream);
/* Same, for an encoder whose frames are sharded over the ranks of one NVSwitch domain (frames of a
 * video are independent, anomaly_clip.py:119-123): with `gather` set (rows_per_rank = num_frames,
 * width = output_dim) the epilogue of the output projection also stores every feature row into ALL
 * ranks' gathered buffers [world * num_frames][output_dim] over NVLink peer memory and the last
 * micro-batch's launch raises this rank's flag -- the all-gather of the features is fused into the
 * GEMM that produces them.  aclip_peer_wait holds the consumer's stream until every rank's rows
 * have arrived.  gather == NULL: identical to aclip_vit_forward. */
struct AclipPeerGather;
int aclip_vit_forward_ex(const AclipVitWeights* w, const void* frames, int frames_are_u8,
                         long long num_frames, int micro_batch, const float* mean3_host,
                         const float* std3_host, float* features_out, void* workspace,
                         size_t workspace_bytes, int passes, const struct AclipPeerGather* gather,
                         void* stream);

typedef struct AclipAxialAttnWeights { /* PreNorm(SelfAttention), one per axis per depth */
  const float *norm_g, *norm_b;       /* nn.LayerNorm(E) */
  const void* qkv_w;                  /* [to_q.weight ; to_kv.weight] split [2][3E][E], no bias */
  const void* out_w; const float* out_b; /* to_out split [2][E][E], bias [E] */
} AclipAxialAttnWeights;

typedef struct AclipConvFFWeights {   /* ChanLayerNorm -> conv3x3 -> LeakyReLU -> conv3x3 */
  const float *g, *b;                 /* [E] */
  const void* conv1_w; const float* conv1_b; /* split [2][4E][9*E], k = (ky*3+kx)*E + c */
  const void* conv2_w; const float* conv2_b; /* split [2][E][9*4E] */
  /* optional (NULL = absent): the same two weights as f16f8 planes and their accumulator scales
   * 2^-(4 + e_weight); used by passes = 2 calls whose chunk has >= 4096 rows and E % 256 == 0
   * (all three planes) and by passes = 4 calls with E % 64 == 0 (the fp16 plane, one pass) */
  const void* conv1_w8; const void* conv2_w8;
  float conv1_s, conv2_s;
} AclipConvFFWeights;

typedef struct AclipTemporalWeights { /* SelectorModel (test branch) + TemporalModel + head */
  int feature_dim;      /* 512 */
  int num_dirs;         /* C - 1 similarity columns (<= 32) */
  int emb, depth, heads, num_segments, seg_length;
  int concat;           /* concat_features: temporal input = [similarity, x - ncentroid] */
  int ldf;              /* pitch of the packed feature rows: feature_dim (+ 32 if concat) */
  const float* ncentroid;   /* [feature_dim] */
  const void* selector_w;   /* split [2][32][feature_dim]: normalised, re-centred text directions
                               scaled by the BatchNorm eval factor; rows >= num_dirs are zero */
  const float* selector_b;  /* [32]: -running_mean / sqrt(running_var + eps), zero padded */
  const void* proj_w;       /* projection.weight with columns reordered to [x | sim | 0],
                               split [2][emb][ldf] */
  const float* proj_b;      /* [emb] */
  const float* pos;         /* [n*l][emb]: pos_emb.param_0[i] + pos_emb.param_1[k] */
  const AclipAxialAttnWeights* attn; /* host array [2*depth]: (axis n, axis l) per depth */
  const AclipConvFFWeights* ff;      /* host array [2*depth]: (f, g) per depth */
  const float *head_ln_g, *head_ln_b, *head_w; /* classifier.layer_norm, classifier.linear.weight */
  float head_bias;
} AclipTemporalWeights;

size_t aclip_temporal_workspace_bytes(const AclipTemporalWeights* w, long long sub_videos);

/* Optional fused all-gather of the result rows over NVLink peer memory (one process per GPU).
 * Every rank owns a buffer rows[world * rows_per_rank][width] (width = 1 + num_dirs: score, then
 * the class probabilities) and a flag array flags[world], both mapped into every peer (e.g. with
 * torch.distributed._symmetric_memory).  The head kernel of aclip_temporal_forward_ex stores each
 * of its rows into ALL peers' buffers at row (rank * rows_per_rank + r) while it computes them,
 * and its last CTA publishes flags[rank] = epoch on every peer (system-scope release).
 * aclip_peer_wait then blocks the STREAM (not the host) until all `world` flags of the local array
 * have reached `epoch`.  This replaces the NCCL all-gather of per-frame scores. */
typedef struct AclipPeerGather {
  int world, rank;
  long long rows_per_rank;
  int width;
  float* rows[8];            /* peer-mapped base of each rank's gathered buffer (index = peer rank) */
  unsigned int* flags[8];    /* peer-mapped flag arrays [world] */
  unsigned int epoch;        /* > 0, increasing by one per call on every rank */
  unsigned int* counter;     /* LOCAL zero-initialised device word (CTA completion counter) */
} AclipPeerGather;

/* local_flags: [2 * world] words: arrival flags, then timeout markers (set to 1 for a rank whose
 * flag did not arrive within ~5 s; the wait never hangs the device). */
int aclip_peer_wait(unsigned int* local_flags, int world, unsigned int epoch, void* stream);
/* Raise this rank's flag on every peer without storing rows: a rank that has no unit in a call
 * (fewer units than ranks) still takes part in the exchange of that epoch. */
int aclip_peer_signal(const AclipPeerGather* gather, void* stream);

/* AnomalyCLIP.forward(test_mode=True) after the image encoder (anomaly_clip.py:132-154) fused with
 * test_step's softmax(similarity) * score (anomaly_clip_module.py:473-477).
 * features: fp32 [N][feature_dim], rows in the caller's "(b n s l)" order, N = sub_videos*n*l with
 * sub_videos = b*segment_size.  Outputs in the same row order: similarity_out [N][num_dirs],
 * scores_out [N], class_probs_out [N][num_dirs] (may be NULL).  Sub-videos are processed in as
 * large chunks as `workspace` allows.  passes: 3 (split-bf16) or 1 (bf16) for every GEMM; 2 runs the
 * 3x3 conv GEMMs (94 % of the work) on f16f8 operands where AclipConvFFWeights carries them and the
 * chunk is large enough for the CTA-pair kernel, everything else as passes = 3; 4 runs the conv
 * GEMMs on fp16 operands in one pass at any chunk size (~9e-5 on the scores; class indices cannot
 * change: the score scales every class alike), everything else as passes = 3. */
int aclip_temporal_forward(const AclipTemporalWeights* w, const float* features,
                           long long sub_videos, int segment_size, float* similarity_out,
                           float* scores_out, float* class_probs_out, void* workspace,
                           size_t workspace_bytes, int passes, void* stream);
/* TemporalModel.forward on its own (temporal_model.py:42-77) after the projection: `projected` is
 * fp32 [sub_videos*n*l][emb] = projection(features) + axial positional embedding, rows in sub-video
 * order (overwritten); scores_out [N] comes back in the caller's "(b n s l)" order.  Only the
 * transformer / classifier fields of `w` are read.  Workspace as aclip_temporal_workspace_bytes. */
int aclip_temporal_core_forward(const AclipTemporalWeights* w, float* projected, long long sub_videos,
                                int segment_size, float* scores_out, void* workspace,
                                size_t workspace_bytes, int passes, void* stream);
/* Same as aclip_temporal_forward, with the fused peer all-gather (gather may be NULL). */
int aclip_temporal_forward_ex(const AclipTemporalWeights* w, const float* features,
                              long long sub_videos, int segment_size, float* similarity_out,
                              float* scores_out, float* class_probs_out, void* workspace,
                              size_t workspace_bytes, int passes, const AclipPeerGather* gather,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACLIP_B200_H_ */
