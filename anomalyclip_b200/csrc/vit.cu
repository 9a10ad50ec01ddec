// aclip_vit_forward: the CLIP ViT image encoder (VisionTransformer.forward,
// /root/reference/src/models/components/clip/model.py:266-290) as a fixed sequence of the
// library's kernels.  Per micro-batch of Bm frames (M = Bm * tokens residual rows):
//
//   patchify -> GEMM(conv1) + pos-emb, rows remapped past the CLS slot -> CLS rows -> ln_pre
//   12 x [ LN1 -> GEMM(in_proj)+b -> attention -> GEMM(out_proj)+b+residual
//          LN2 -> GEMM(c_fc)+b+QuickGELU -> GEMM(c_proj)+b+residual ]
//   ln_post on the CLS rows -> GEMM(proj)
//
// The residual stream X stays fp32; GEMM operands travel in the encoding of the requested mode
// (gemm.cuh / split.cuh): passes = 3 split-bf16, 2 f16f8 (out_proj and the attention stay
// split-bf16), 4 fp16 planes end to end (every GEMM and the attention in one pass), 5 "mixed":
// the attention side of a block (in_proj, attention, out_proj: a third of its flops, a tenth of its
// error sensitivity -- scripts/numerics_passes.py) on fp16 operands, the MLP pair, the patch
// embedding and the output projection on f16f8 operands.  No allocation,
// no synchronisation: everything is enqueued on the caller's stream into the caller's workspace.
#include "common.h"

namespace aclip {

namespace {

constexpr size_t kAlign = 1024;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct VitPlan {
  int tokens, grid2, k0;
  size_t x_bytes, h_plane, big_plane;  // planes in elements
  size_t off_x, off_h, off_big, total;
};

VitPlan plan_vit(const AclipVitWeights& w, int mb) {
  VitPlan p{};
  const int G = w.resolution / w.patch;
  p.grid2 = G * G;
  p.tokens = p.grid2 + 1;
  p.k0 = 3 * w.patch * w.patch;
  const size_t rows = static_cast<size_t>(mb) * p.tokens;
  p.x_bytes = rows * w.width * sizeof(float);
  p.h_plane = rows * w.width;
  const size_t patch_elems = static_cast<size_t>(mb) * p.grid2 * p.k0;
  p.big_plane = rows * 4 * w.width;
  if (patch_elems > p.big_plane) p.big_plane = patch_elems;
  p.big_plane = (p.big_plane + 15) / 16 * 16;
  p.off_x = 0;
  p.off_h = align_up(p.off_x + p.x_bytes);
  p.off_big = align_up(p.off_h + 2 * p.h_plane * 2);
  p.total = align_up(p.off_big + 2 * p.big_plane * 2);
  return p;
}

AclipGemmArgs linear(const void* a, long long a_plane, int M, int K, int lda, const void* w, int N,
                     int passes, float w_scale = 0.0f) {
  AclipGemmArgs g{};
  g.out_scale = (passes == 2 || passes == 4) ? w_scale : 0.0f;   // fp16-based operands: 2^-(4 + e_w)
  g.a = a; g.w = w;
  g.M = M; g.N = N; g.K = K;
  g.lda = lda; g.ldw = K;
  g.a_plane_stride = a_plane;
  g.w_plane_stride = static_cast<long long>(N) * K;
  g.passes = passes;
  return g;
}

int check_weights(const AclipVitWeights& w) {
  ACLIP_REQUIRE(w.width > 0 && w.width % 64 == 0 && w.heads * 64 == w.width,
                "vit: width=%d heads=%d (head dim must be 64)", w.width, w.heads);
  ACLIP_REQUIRE(w.layers >= 0 && w.patch % 8 == 0 && w.resolution % w.patch == 0 &&
                    w.resolution % 8 == 0,
                "vit: patch=%d resolution=%d unsupported", w.patch, w.resolution);
  ACLIP_REQUIRE(w.output_dim % 32 == 0, "vit: output_dim=%d must be a multiple of 32", w.output_dim);
  ACLIP_REQUIRE(w.conv1_w && w.class_embedding && w.positional_embedding && w.ln_pre_g &&
                    w.ln_pre_b && w.ln_post_g && w.ln_post_b && w.proj_w &&
                    (w.layers == 0 || w.blocks),
                "vit: null weight pointer");
  return ACLIP_OK;
}

}  // namespace

}  // namespace aclip

extern "C" size_t aclip_vit_workspace_bytes(const AclipVitWeights* w, int micro_batch) {
  if (w == nullptr || micro_batch <= 0 || w->patch <= 0) return 0;
  return aclip::plan_vit(*w, micro_batch).total;
}

extern "C" int aclip_vit_forward(const AclipVitWeights* wp, const void* frames, int frames_are_u8,
                                 long long num_frames, int micro_batch, const float* mean3_host,
                                 const float* std3_host, float* features_out, void* workspace,
                                 size_t workspace_bytes, int passes, void* stream_) {
  return aclip_vit_forward_ex(wp, frames, frames_are_u8, num_frames, micro_batch, mean3_host,
                              std3_host, features_out, workspace, workspace_bytes, passes, nullptr,
                              stream_);
}

extern "C" int aclip_vit_forward_ex(const AclipVitWeights* wp, const void* frames, int frames_are_u8,
                                    long long num_frames, int micro_batch, const float* mean3_host,
                                    const float* std3_host, float* features_out, void* workspace,
                                    size_t workspace_bytes, int passes,
                                    const AclipPeerGather* gather, void* stream_) {
  using namespace aclip;
  ACLIP_REQUIRE(wp != nullptr, "vit_forward: weights are NULL");
  const AclipVitWeights& w = *wp;
  ACLIP_TRY(check_weights(w));
  ACLIP_REQUIRE(frames != nullptr && features_out != nullptr, "vit_forward: null frames/output");
  ACLIP_REQUIRE(num_frames >= 0 && micro_batch > 0, "vit_forward: bad frame count / micro-batch");
  ACLIP_REQUIRE(passes >= 1 && passes <= 7, "vit_forward: passes must be 1 .. 7");
  // 7 = 5 with the MLP pair on f16mx operands (fp16 main product + two MXFP4 cross terms, 1.5
  // pass-equivalents instead of 2; gemm_mx.cuh): ln_2 writes f16mx, c_fc reads and writes it
  const bool mlp_mx = passes == 7;
  if (mlp_mx) {
    ACLIP_REQUIRE(w.width % 192 == 0, "vit_forward: passes=7 needs a width that is a multiple of 192 (and 256)");
    passes = 5;
  }
  ACLIP_REQUIRE(passes == 1 || passes == 3 || (w.width % 256 == 0 && w.output_dim % 256 == 0 &&
                                               (3 * w.patch * w.patch) % 16 == 0),
                "vit_forward: passes=%d (fp16-based operands) needs width and output_dim multiples of 256",
                passes);
  // operand mode of the attention side (in_proj, attention, out_proj) and of everything else
  // 6 = 5 with c_proj issued WITHOUT its weight-residual cross term (x_H w_H + x_L w_C, gemm.cuh)
  const int p_cproj = passes == 6 ? 6 : 0;
  if (passes == 6) passes = 5;
  const int p_att = passes == 5 ? 4 : passes;
  const int p_mlp = passes == 5 ? 2 : passes;
  const bool f16 = p_att == 4;                       // fp16 q | k | v, one-pass attention and out_proj
  const int enc_att = f16 ? 2 : p_att == 2 ? 1 : 0;  // encoding of ln_1's output (in_proj's A operand)
  const int enc = p_mlp == 4 ? 2 : p_mlp == 2 ? 1 : 0;  // encoding of the other GEMM A operands
  passes = p_mlp;
  ACLIP_REQUIRE(gather == nullptr || (gather->width == w.output_dim && num_frames == gather->rows_per_rank),
                "vit_forward: the feature gather must be built for %lld rows of %d values", num_frames,
                w.output_dim);
  if (num_frames == 0) return ACLIP_OK;
  if (micro_batch > num_frames) micro_batch = static_cast<int>(num_frames);
  const VitPlan pl = plan_vit(w, micro_batch);
  ACLIP_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) == 0,
                "vit_forward: workspace must be 1024-byte aligned");
  if (workspace_bytes < pl.total)
    return fail(ACLIP_ERR_WORKSPACE, "vit_forward: workspace %zu < %zu bytes", workspace_bytes, pl.total);
  cudaStream_t stream = as_stream(stream_);

  auto* base = static_cast<uint8_t*>(workspace);
  float* X = reinterpret_cast<float*>(base + pl.off_x);
  void* H = base + pl.off_h;
  void* BIG = base + pl.off_big;
  const long long hp = static_cast<long long>(pl.h_plane), bp = static_cast<long long>(pl.big_plane);
  const int W = w.width, T = pl.tokens;
  const size_t frame_elems = static_cast<size_t>(3) * w.resolution * w.resolution;

  for (long long f0 = 0; f0 < num_frames; f0 += micro_batch) {
    const int Bm = static_cast<int>(num_frames - f0 < micro_batch ? num_frames - f0 : micro_batch);
    const int M = Bm * T;
    const void* fr = frames_are_u8
                         ? static_cast<const void*>(static_cast<const uint8_t*>(frames) + f0 * frame_elems)
                         : static_cast<const void*>(static_cast<const float*>(frames) + f0 * frame_elems);
    // patch embedding (:267-269) + positional embedding of the patch tokens (:278)
    ACLIP_TRY(patchify(fr, frames_are_u8, Bm, w.resolution, w.patch, mean3_host, std3_host, BIG, bp, enc, stream));
    {
      AclipGemmArgs g = linear(BIG, bp, Bm * pl.grid2, pl.k0, pl.k0, w.conv1_w, W, passes, w.conv1_s);
      g.residual = w.positional_embedding + W;  // rows 1.. of the table
      g.res_mod = pl.grid2;
      g.ldr = W;
      g.out_f32 = X;
      g.ldc = W;
      g.row_group = pl.grid2; g.row_group_stride = T; g.row_offset = 1;
      ACLIP_TRY(gemm(g, stream));
    }
    ACLIP_TRY(cls_rows(X, Bm, T, W, w.class_embedding, w.positional_embedding, stream));  // :270-278
    ACLIP_TRY(layernorm(X, M, W, W, w.ln_pre_g, w.ln_pre_b, 1e-5f, 0, X, W, nullptr, 0, 0, 0, stream));  // :279

    for (int l = 0; l < w.layers; ++l) {  // :214-217
      const AclipVitBlock& b = w.blocks[l];
      ACLIP_REQUIRE(b.ln1_g && b.ln1_b && b.ln2_g && b.ln2_b && b.qkv_w && b.qkv_b &&
                        (f16 ? b.out_w16 != nullptr : b.out_w != nullptr) &&
                        b.out_b && b.fc_w && b.fc_b && b.proj_w && b.proj_b,
                    "vit_forward: block %d has a null weight", l);
      ACLIP_TRY(layernorm(X, M, W, W, b.ln1_g, b.ln1_b, 1e-5f, 0, nullptr, 0, H, W, hp, enc_att, stream));
      {
        AclipGemmArgs g = linear(H, hp, M, W, W, b.qkv_w, 3 * W, p_att, b.qkv_s);
        g.bias = b.qkv_b;
        g.out_split = BIG; g.split_plane_stride = bp; g.ld_split = 3 * W;
        g.out_enc = f16 ? 2 : 0;
        ACLIP_TRY(gemm(g, stream));
      }
      // q | k | v and the attention output stay bf16 hi/lo planes in every mode: out_proj is the
      // smallest GEMM of the block and runs the three-pass kernel also when passes = 2 (the f16f8
      // encode in the attention epilogue costs more than the two-pass out_proj saves)
      // (passes = 4: fp16 q | k | v in, fp16 out, one MMA pass, and out_proj on its fp16 weight)
      ACLIP_TRY(vit_attention(BIG, bp, 3 * W, Bm, T, w.heads, H, hp, W, 0, f16 ? 2 : 0, stream));
      {
        AclipGemmArgs g = f16 ? linear(H, hp, M, W, W, b.out_w16, W, 4, b.out_s)
                              : linear(H, hp, M, W, W, b.out_w, W, p_att == 2 ? 3 : p_att);
        g.bias = b.out_b;
        g.residual = X; g.ldr = W;
        g.out_f32 = X; g.ldc = W;
        ACLIP_TRY(gemm(g, stream));
      }
      if (mlp_mx) {
        ACLIP_REQUIRE(b.fc_wmx != nullptr && b.proj_wmx != nullptr, "vit_forward: block %d lacks its f16mx weights", l);
        const long long mw = static_cast<long long>(M) * W;
        ACLIP_TRY(layernorm(X, M, W, W, b.ln2_g, b.ln2_b, 1e-5f, 0, nullptr, 0, H, W, mw, 3, stream));
        AclipGemmArgs g = linear(H, mw, M, W, W, b.fc_wmx, 4 * W, 7, b.fc_s);
        g.out_scale = b.fc_s;
        g.bias = b.fc_b;
        g.act = ACLIP_ACT_QUICKGELU;
        g.out_enc = 3;
        g.out_split = BIG; g.split_plane_stride = 4 * mw; g.ld_split = 4 * W;
        ACLIP_TRY(gemm(g, stream));
        AclipGemmArgs q = linear(BIG, 4 * mw, M, 4 * W, 4 * W, b.proj_wmx, W, 7, b.proj_s);
        q.out_scale = b.proj_s;
        q.bias = b.proj_b;
        q.residual = X; q.ldr = W;
        q.out_f32 = X; q.ldc = W;
        ACLIP_TRY(gemm(q, stream));
        continue;
      }
      ACLIP_TRY(layernorm(X, M, W, W, b.ln2_g, b.ln2_b, 1e-5f, 0, nullptr, 0, H, W, hp, enc, stream));
      {
        AclipGemmArgs g = linear(H, hp, M, W, W, b.fc_w, 4 * W, passes, b.fc_s);
        g.bias = b.fc_b;
        g.act = ACLIP_ACT_QUICKGELU;
        g.out_enc = enc;
        g.out_split = BIG; g.split_plane_stride = bp; g.ld_split = 4 * W;
        ACLIP_TRY(gemm(g, stream));
      }
      {
        AclipGemmArgs g = linear(BIG, bp, M, 4 * W, 4 * W, b.proj_w, W, passes, b.proj_s);
        if (p_cproj != 0) g.passes = p_cproj;
        g.bias = b.proj_b;
        g.residual = X; g.ldr = W;
        g.out_f32 = X; g.ldc = W;
        ACLIP_TRY(gemm(g, stream));
      }
    }
    // ln_post on the CLS token (:285) and the output projection (:287-288)
    ACLIP_TRY(layernorm(X, Bm, W, static_cast<long long>(T) * W, w.ln_post_g, w.ln_post_b, 1e-5f, 0,
                        nullptr, 0, H, W, hp, enc, stream));
    {
      AclipGemmArgs g = linear(H, hp, Bm, W, W, w.proj_w, w.output_dim, passes, w.proj_s);
      g.out_f32 = features_out + f0 * w.output_dim;
      g.ldc = w.output_dim;
      // frame-sharded encoder: the projection's epilogue stores the feature rows straight into
      // every rank's gathered buffer; the launch of the last micro-batch raises this rank's flag
      g.gather = gather;
      g.gather_row0 = f0;
      g.gather_signal = f0 + Bm >= num_frames ? 1 : 0;
      ACLIP_TRY(gemm(g, stream));
    }
  }
  return ACLIP_OK;
}
