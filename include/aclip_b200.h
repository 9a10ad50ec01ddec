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

/* ------------------------------------------------------------------------------------------
 * Building-block operators
 * ---------------------------------------------------------------------------------------- */

/* fp32 [rows][cols] (pitch ld_in) -> split bf16 [2][rows][ld_out]; columns cols..ld_out-1 are
 * zero-filled so that the result can feed a GEMM whose K is padded. */
int aclip_split_f32(const float* in, long long rows, int cols, int ld_in, void* out_split,
                    int ld_out, long long plane_stride, void* stream);

typedef struct AclipGemmArgs {
  /* operands (bf16 split planes) */
  const void* a;  /* linear: [planes][M][lda]; conv3x3: [planes][S][H][W][C]            */
  const void* w;  /* [planes][N][ldw]  (nn.Linear weight layout: out_features x in_features) */
  int M, N, K;    /* C[M,N] = A[M,K] * W[N,K]^T ; conv3x3: M = S*H*W, K = 9*C            */
  int lda, ldw;   /* row pitches in elements, multiples of 8                              */
  long long a_plane_stride, w_plane_stride; /* elements between hi and lo plane          */
  int passes;     /* 3 = split-bf16 (fp32-faithful), 1 = plain bf16 (hi plane only)      */
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
  int ldc;
  /* output row remap (0,0,0 = identity):
   * out_row = (m / row_group) * row_group_stride + (m % row_group) + row_offset */
  int row_group, row_group_stride, row_offset;
  int max_ctas;          /* 0 = one persistent CTA per SM */
} AclipGemmArgs;

/* tcgen05 / TMA GEMM with fused epilogue. N must be a multiple of 32, K a multiple of 8. */
int aclip_gemm(const AclipGemmArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACLIP_B200_H_ */
