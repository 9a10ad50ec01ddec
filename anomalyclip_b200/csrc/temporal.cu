// aclip_temporal_forward: everything AnomalyCLIP.forward(test_mode=True) does after the image
// encoder (/root/reference/src/models/components/anomaly_clip.py:132-154), for feature rows that
// are already in HBM:
//
//   centre + regroup rows to sub-video order                 (selector_model.py:54, anomaly_clip.py:143,
//                                                             temporal_model.py:46-53)
//   GEMM(selector directions) -> similarity (BatchNorm eval folded into the operand)   (selector_model.py:44-65)
//   GEMM(projection) + axial positional embedding             (temporal_model.py:43; axial pos-emb)
//   depth x [ LN -> GEMM(q|kv) -> axial attention (n) -> GEMM(to_out)+b+residual
//             LN -> GEMM(q|kv) -> axial attention (l) -> GEMM(to_out)+b+residual
//             ChanLN -> conv3x3 GEMM + LeakyReLU -> conv3x3 GEMM + residual   (x2: f and g) ]
//   head: mean of halves -> LN -> Linear -> sigmoid; softmax(similarity) * score; rows written back
//   in the caller's order                                      (classification_head.py:11-15,
//                                                             anomaly_clip_module.py:473-477)
//
// The reversible network's two streams are kept as two fp32 buffers (x1 in A1, x2 in P).
#include "common.h"
#include "rowmap.cuh"

namespace aclip {

namespace {

constexpr size_t kAlign = 1024;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct TemporalPlan {
  size_t rows;
  size_t f_plane, h_plane, mid_plane;  // elements
  size_t off_f, off_sim, off_p, off_a1, off_h, off_qkv, off_mid, total;
};

TemporalPlan plan_temporal(const AclipTemporalWeights& w, long long sub_videos) {
  TemporalPlan p{};
  const size_t unit = static_cast<size_t>(w.num_segments) * w.seg_length;
  p.rows = static_cast<size_t>(sub_videos) * unit;
  const size_t E = w.emb;
  p.f_plane = p.rows * w.ldf;
  p.h_plane = p.rows * E;
  p.mid_plane = p.rows * 4 * E;
  size_t off = 0;
  p.off_f = off;   off = align_up(off + 2 * p.f_plane * 2);
  p.off_sim = off; off = align_up(off + p.rows * 32 * sizeof(float));
  p.off_p = off;   off = align_up(off + p.rows * E * sizeof(float));
  p.off_a1 = off;  off = align_up(off + p.rows * E * sizeof(float));
  p.off_h = off;   off = align_up(off + 2 * p.h_plane * 2);
  p.off_qkv = off; off = align_up(off + p.rows * 3 * E * sizeof(float));
  p.off_mid = off; off = align_up(off + 2 * p.mid_plane * 2);
  p.total = off;
  return p;
}

AclipGemmArgs linear(const void* a, long long a_plane, long long M, int K, int lda, const void* w,
                     int N, int passes) {
  AclipGemmArgs g{};
  g.a = a; g.w = w;
  g.M = static_cast<int>(M); g.N = N; g.K = K;
  g.lda = lda; g.ldw = K;
  g.a_plane_stride = a_plane;
  g.w_plane_stride = static_cast<long long>(N) * K;
  g.passes = passes;
  return g;
}

int check_weights(const AclipTemporalWeights& w) {
  ACLIP_REQUIRE(w.feature_dim > 0 && w.feature_dim % 64 == 0, "temporal: feature_dim=%d", w.feature_dim);
  ACLIP_REQUIRE(w.num_dirs >= 1 && w.num_dirs <= 32, "temporal: num_dirs=%d (1..32)", w.num_dirs);
  ACLIP_REQUIRE(w.emb % 64 == 0 && w.emb <= 256, "temporal: emb=%d must be 64, 128, 192 or 256", w.emb);
  ACLIP_REQUIRE(w.depth >= 0 && w.heads > 0, "temporal: depth=%d heads=%d", w.depth, w.heads);
  ACLIP_REQUIRE(w.seg_length > 0 && 128 % w.seg_length == 0 &&
                    (w.num_segments * w.seg_length) % 128 == 0,
                "temporal: grid %dx%d unsupported by the conv GEMM (need l | 128 and 128 | n*l)",
                w.num_segments, w.seg_length);
  ACLIP_REQUIRE(w.ldf == w.feature_dim + (w.concat ? 32 : 0), "temporal: ldf=%d inconsistent", w.ldf);
  ACLIP_REQUIRE(w.ncentroid && w.selector_w && w.selector_b && w.proj_w && w.proj_b && w.pos &&
                    w.head_ln_g && w.head_ln_b && w.head_w && (w.depth == 0 || (w.attn && w.ff)),
                "temporal: null weight pointer");
  return ACLIP_OK;
}


// The reversible axial transformer + head on projected rows P (fp32 [rows][E], sub-video order;
// overwritten).  sim may be NULL (no class probabilities: TemporalModel.forward on its own).
int run_core(const AclipTemporalWeights& w, long long cs, float* P, float* A1, void* H, long long hp,
             float* QKV, void* MID, long long mp, const float* SIM, const RowMap& map,
             float* scores_out, float* similarity_out, float* class_probs_out,
             const AclipPeerGather* gather, int signal, int passes, cudaStream_t stream) {
  const int n = w.num_segments, l = w.seg_length, E = w.emb;
  const long long rows = cs * n * l;
  // passes = 2: f16f8 operands for the conv GEMMs when the CTA-pair kernel applies (N = 4E and E
  // multiples of 256, enough rows to fill it); passes = 4: fp16 operands in ONE pass for the conv
  // GEMMs (94 % of this stage's flops; ~9e-5 on the scores, which scale every class alike and so
  // cannot change a class index) at any chunk size; every other GEMM of this stage runs three passes
  const bool want8 = passes == 2 && E % 256 == 0 && rows >= 4096;
  const bool want16 = passes == 4 && E % 64 == 0;   // single-CTA tiles serve any N % 32 (XD: E = 128)
  if (passes == 2 || passes == 4) passes = 3;
  const float* x1 = P;  // both reversible streams start as the same tensor
  for (int d = 0; d < w.depth; ++d) {
    for (int axis = 0; axis < 2; ++axis) {  // y1 = x1 + Attn_n(LN(x2)); y2 = x2 + Attn_l(LN(y1))
      const AclipAxialAttnWeights& a = w.attn[2 * d + axis];
      ACLIP_REQUIRE(a.norm_g && a.norm_b && a.qkv_w && a.out_w && a.out_b,
                    "temporal_forward: attention %d/%d has a null weight", d, axis);
      const float* src = axis == 0 ? P : A1;
      ACLIP_TRY(layernorm(src, rows, E, E, a.norm_g, a.norm_b, 1e-5f, 0, nullptr, 0, H, E, hp, 0, stream));
      {
        AclipGemmArgs g = linear(H, hp, rows, E, E, a.qkv_w, 3 * E, passes);
        g.out_f32 = QKV; g.ldc = 3 * E;
        ACLIP_TRY(gemm(g, stream));
      }
      ACLIP_TRY(axial_attention(QKV, cs, n, l, E, w.heads, axis, H, hp, stream));
      {
        AclipGemmArgs g = linear(H, hp, rows, E, E, a.out_w, E, passes);
        g.bias = a.out_b;
        g.residual = axis == 0 ? x1 : P; g.ldr = E;
        g.out_f32 = axis == 0 ? A1 : P; g.ldc = E;
        ACLIP_TRY(gemm(g, stream));
      }
    }
    x1 = A1;
    for (int fg = 0; fg < 2; ++fg) {  // y1 = x1 + FF_f(x2); y2 = x2 + FF_g(y1)
      const AclipConvFFWeights& c = w.ff[2 * d + fg];
      ACLIP_REQUIRE(c.g && c.b && c.conv1_w && c.conv1_b && c.conv2_w && c.conv2_b,
                    "temporal_forward: feed-forward %d/%d has a null weight", d, fg);
      const float* src = fg == 0 ? P : A1;
      float* dst = fg == 0 ? A1 : P;
      const bool have8 = c.conv1_w8 != nullptr && c.conv2_w8 != nullptr;
      const bool ff16 = want16 && have8;             // fp16 planes (of the f16f8 weight pack), one pass
      const bool ff8 = (want8 && have8) || ff16;     // fp16-based operands: accumulator scale applies
      const int ff_passes = ff16 ? 4 : ff8 ? 2 : passes;
      const int ff_enc = ff16 ? 2 : ff8 ? 1 : 0;
      ACLIP_TRY(layernorm(src, rows, E, E, c.g, c.b, 1e-5f, 1, nullptr, 0, H, E, hp, ff_enc, stream));
      {
        AclipGemmArgs g = linear(H, hp, rows, 9 * E, E, ff8 ? c.conv1_w8 : c.conv1_w, 4 * E, ff_passes);
        g.a_mode = 1; g.conv_c = E; g.conv_h = n; g.conv_w = l; g.conv_s = static_cast<int>(cs);
        g.bias = c.conv1_b;
        g.act = ACLIP_ACT_LEAKYRELU;
        g.out_split = MID; g.split_plane_stride = mp; g.ld_split = 4 * E;
        g.out_scale = ff8 ? c.conv1_s : 0.0f;
        g.out_enc = ff_enc;
        ACLIP_TRY(gemm(g, stream));
      }
      {
        AclipGemmArgs g = linear(MID, mp, rows, 36 * E, 4 * E, ff8 ? c.conv2_w8 : c.conv2_w, E, ff_passes);
        g.out_scale = ff8 ? c.conv2_s : 0.0f;
        g.a_mode = 1; g.conv_c = 4 * E; g.conv_h = n; g.conv_w = l; g.conv_s = static_cast<int>(cs);
        g.bias = c.conv2_b;
        g.residual = dst; g.ldr = E;
        g.out_f32 = dst; g.ldc = E;
        ACLIP_TRY(gemm(g, stream));
      }
    }
  }
  ACLIP_TRY(score_head(x1, P, rows, E, w.head_ln_g, w.head_ln_b, 1e-5f, w.head_w, w.head_bias,
                       SIM, 32, SIM != nullptr ? w.num_dirs : 0, map, scores_out, similarity_out, class_probs_out,
                       gather, signal, stream));
  return ACLIP_OK;
}

}  // namespace

}  // namespace aclip

extern "C" size_t aclip_temporal_workspace_bytes(const AclipTemporalWeights* w, long long sub_videos) {
  if (w == nullptr || sub_videos <= 0) return 0;
  return aclip::plan_temporal(*w, sub_videos).total;
}

extern "C" int aclip_temporal_forward(const AclipTemporalWeights* wp, const float* features,
                                      long long sub_videos, int segment_size,
                                      float* similarity_out, float* scores_out,
                                      float* class_probs_out, void* workspace,
                                      size_t workspace_bytes, int passes, void* stream_) {
  return aclip_temporal_forward_ex(wp, features, sub_videos, segment_size, similarity_out,
                                   scores_out, class_probs_out, workspace, workspace_bytes, passes,
                                   nullptr, stream_);
}

extern "C" int aclip_temporal_forward_ex(const AclipTemporalWeights* wp, const float* features,
                                         long long sub_videos, int segment_size,
                                         float* similarity_out, float* scores_out,
                                         float* class_probs_out, void* workspace,
                                         size_t workspace_bytes, int passes,
                                         const AclipPeerGather* gather, void* stream_) {
  using namespace aclip;
  ACLIP_REQUIRE(wp != nullptr, "temporal_forward: weights are NULL");
  const AclipTemporalWeights& w = *wp;
  ACLIP_TRY(check_weights(w));
  ACLIP_REQUIRE(features && similarity_out && scores_out, "temporal_forward: null input/output");
  ACLIP_REQUIRE(sub_videos >= 0 && segment_size >= 1 && sub_videos % segment_size == 0,
                "temporal_forward: sub_videos=%lld must be a multiple of segment_size=%d",
                sub_videos, segment_size);
  ACLIP_REQUIRE(passes >= 1 && passes <= 4, "temporal_forward: passes must be 1, 2, 3 or 4");
  if (sub_videos == 0) return ACLIP_OK;
  // a rank may contribute fewer rows than its block holds (uneven unit counts over the ranks)
  ACLIP_REQUIRE(gather == nullptr ||
                    gather->rows_per_rank >= sub_videos * w.num_segments * w.seg_length,
                "temporal_forward: peer gather holds %lld rows per rank, fewer than this call's",
                gather ? gather->rows_per_rank : 0ll);
  ACLIP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) == 0,
                "temporal_forward: workspace must be 1024-byte aligned");
  // largest chunk of sub-videos the workspace can hold
  long long chunk = sub_videos;
  const size_t per_one = plan_temporal(w, 1).total;
  if (plan_temporal(w, chunk).total > workspace_bytes) {
    chunk = static_cast<long long>(workspace_bytes / per_one);
    while (chunk > 0 && plan_temporal(w, chunk).total > workspace_bytes) --chunk;
  }
  if (chunk <= 0)
    return fail(ACLIP_ERR_WORKSPACE, "temporal_forward: workspace %zu < %zu bytes (one sub-video)",
                workspace_bytes, per_one);
  const long long max_chunk = (1ll << 31) / (static_cast<long long>(w.num_segments) * w.seg_length) - 1;
  if (chunk > max_chunk) chunk = max_chunk;
  cudaStream_t stream = as_stream(stream_);

  const int n = w.num_segments, l = w.seg_length, E = w.emb, D = w.feature_dim;
  const long long unit = static_cast<long long>(n) * l;
  const TemporalPlan pl = plan_temporal(w, chunk);
  auto* base = static_cast<uint8_t*>(workspace);
  void* F = base + pl.off_f;
  float* SIM = reinterpret_cast<float*>(base + pl.off_sim);
  float* P = reinterpret_cast<float*>(base + pl.off_p);
  float* A1 = reinterpret_cast<float*>(base + pl.off_a1);
  void* H = base + pl.off_h;
  float* QKV = reinterpret_cast<float*>(base + pl.off_qkv);
  void* MID = base + pl.off_mid;
  const long long fp = static_cast<long long>(pl.f_plane), hp = static_cast<long long>(pl.h_plane);
  const long long mp = static_cast<long long>(pl.mid_plane);

  const int lin_passes = (passes == 2 || passes == 4) ? 3 : passes;  // selector / projection: split-bf16
  for (long long u0 = 0; u0 < sub_videos; u0 += chunk) {
    const long long cs = sub_videos - u0 < chunk ? sub_videos - u0 : chunk;
    const long long rows = cs * unit;
    const RowMap map{n, segment_size, l, u0 * unit};

    ACLIP_TRY(center_regroup(features, rows, D, w.ncentroid, map, F, w.ldf, fp, stream));
    {  // similarity = BatchNorm_eval((x - m) @ directions^T)
      AclipGemmArgs g = linear(F, fp, rows, D, w.ldf, w.selector_w, 32, lin_passes);
      g.bias = w.selector_b;
      g.out_f32 = SIM; g.ldc = 32;
      if (w.concat) {  // similarity columns follow the centred features in the packed rows
        g.out_split = static_cast<uint8_t*>(F) + static_cast<size_t>(D) * 2;
        g.split_plane_stride = fp;
        g.ld_split = w.ldf;
      }
      ACLIP_TRY(gemm(g, stream));
    }
    {  // projection + axial positional embedding
      AclipGemmArgs g = linear(F, fp, rows, w.ldf, w.ldf, w.proj_w, E, lin_passes);
      g.bias = w.proj_b;
      g.residual = w.pos; g.res_mod = static_cast<int>(unit); g.ldr = E;
      g.out_f32 = P; g.ldc = E;
      ACLIP_TRY(gemm(g, stream));
    }

    ACLIP_TRY(run_core(w, cs, P, A1, H, hp, QKV, MID, mp, SIM, map, scores_out, similarity_out,
                       class_probs_out, gather, u0 + cs >= sub_videos ? 1 : 0, passes, stream));
  }
  return ACLIP_OK;
}

extern "C" int aclip_temporal_core_forward(const AclipTemporalWeights* wp, float* projected,
                                           long long sub_videos, int segment_size,
                                           float* scores_out, void* workspace,
                                           size_t workspace_bytes, int passes, void* stream_) {
  using namespace aclip;
  ACLIP_REQUIRE(wp != nullptr && projected != nullptr && scores_out != nullptr,
                "temporal_core_forward: null argument");
  const AclipTemporalWeights& w = *wp;
  ACLIP_REQUIRE(w.emb % 64 == 0 && w.emb <= 256 && w.depth >= 0 && w.heads > 0 &&
                    w.seg_length > 0 && 128 % w.seg_length == 0 &&
                    (w.num_segments * w.seg_length) % 128 == 0 && w.head_ln_g && w.head_ln_b &&
                    w.head_w && (w.depth == 0 || (w.attn && w.ff)),
                "temporal_core_forward: unsupported configuration or null weight");
  ACLIP_REQUIRE(sub_videos >= 0 && segment_size >= 1 && sub_videos % segment_size == 0,
                "temporal_core_forward: sub_videos must be a multiple of segment_size");
  ACLIP_REQUIRE(passes >= 1 && passes <= 4, "temporal_core_forward: passes must be 1, 2, 3 or 4");
  if (sub_videos == 0) return ACLIP_OK;
  ACLIP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) == 0,
                "temporal_core_forward: workspace must be 1024-byte aligned");
  const TemporalPlan pl = plan_temporal(w, sub_videos);
  if (workspace_bytes < pl.total)
    return fail(ACLIP_ERR_WORKSPACE, "temporal_core_forward: workspace %zu < %zu bytes",
                workspace_bytes, pl.total);
  auto* base = static_cast<uint8_t*>(workspace);
  const RowMap map{w.num_segments, segment_size, w.seg_length, 0};
  return run_core(w, sub_videos, projected, reinterpret_cast<float*>(base + pl.off_a1), base + pl.off_h,
                  static_cast<long long>(pl.h_plane), reinterpret_cast<float*>(base + pl.off_qkv),
                  base + pl.off_mid, static_cast<long long>(pl.mid_plane), nullptr, map, scores_out,
                  nullptr, nullptr, nullptr, 0, passes, as_stream(stream_));
}
