// ViT / DeiT transformer-attribution engine: forward with saved activations, activation-gradient
// backward (attention gradients only — no dW), LRP relprop through every block, aggregation and
// rollout.  Host-side orchestration of the kernels in te_gemm.cu / te_elementwise.cu /
// te_rollout.cu; O(1) launches per block per BATCH, never per sample.
//
// Reference wiring: baselines/ViT/ViT_LRP.py (forward :305-322, relprop :324-369, Block :196-213,
// Attention :132-177, Mlp :61-74), baselines/ViT/ViT_explanation_generator.py:25-41.
#include <string.h>
#include <string>
#include <vector>

#include "../../include/te_b200.h"
#include "te_engine_util.h"
#include "te_kernels.h"
#include "te_rollout.h"
#include "te_zplus.h"
#include "te_gemm_tc.h"

namespace {

constexpr int kMaxDepth = 64;

struct Dims {
    int B, N, NP, D, H, dh, F, C, L, P, img, Cin, npatch, prefix, KP;
    long long M;
};

static bool make_dims(const te_vit_config* c, int B, Dims& d) {
    if (!c || c->depth <= 0 || c->depth > kMaxDepth || c->heads <= 0 || c->dim % c->heads != 0 || c->dim % 4 != 0 ||
        c->mlp_dim % 4 != 0 || c->patch_size % 4 != 0 || c->img_size % c->patch_size != 0 || c->num_classes <= 0) {
        te_set_last_error("te_vit: invalid config");
        return false;
    }
    d.B = B; d.D = c->dim; d.H = c->heads; d.dh = c->dim / c->heads; d.F = c->mlp_dim; d.C = c->num_classes;
    d.L = c->depth; d.P = c->patch_size; d.img = c->img_size; d.Cin = c->in_chans;
    d.npatch = (c->img_size / c->patch_size) * (c->img_size / c->patch_size);
    d.prefix = c->distilled ? 2 : 1;
    d.N = d.npatch + d.prefix;
    d.NP = (d.N + 3) & ~3;
    d.KP = d.Cin * d.P * d.P;
    d.M = (long long)B * d.N;
    if (d.dh % 4 != 0) { te_set_last_error("te_vit: head_dim % 4 != 0"); return false; }
    return true;
}

// ---- flat weight buffer ------------------------------------------------------------------------
struct WEntry { std::string name; long long numel; long long offset; };

static std::vector<WEntry> weight_table(const te_vit_config* c) {
    Dims d;
    std::vector<WEntry> t;
    if (!make_dims(c, 1, d)) return t;
    long long off = 0;
    auto add = [&](const std::string& n, long long numel) {
        t.push_back({n, numel, off});
        off += (numel + 31) & ~31LL;
    };
    add("patch_embed.proj.weight", (long long)d.D * d.KP);
    add("patch_embed.proj.bias", d.D);
    add("cls_token", d.D);
    if (c->distilled) add("dist_token", d.D);
    add("pos_embed", (long long)d.N * d.D);
    for (int i = 0; i < d.L; ++i) {
        const std::string p = "blocks." + std::to_string(i) + ".";
        add(p + "norm1.weight", d.D);
        add(p + "norm1.bias", d.D);
        add(p + "attn.qkv.weight", 3LL * d.D * d.D);
        add(p + "attn.qkv.bias", 3LL * d.D);
        add(p + "attn.proj.weight", (long long)d.D * d.D);
        add(p + "attn.proj.bias", d.D);
        add(p + "norm2.weight", d.D);
        add(p + "norm2.bias", d.D);
        add(p + "mlp.fc1.weight", (long long)d.F * d.D);
        add(p + "mlp.fc1.bias", d.F);
        add(p + "mlp.fc2.weight", (long long)d.D * d.F);
        add(p + "mlp.fc2.bias", d.D);
    }
    add("norm.weight", d.D);
    add("norm.bias", d.D);
    add("head.weight", (long long)d.C * d.D);
    add("head.bias", d.C);
    if (c->distilled) {
        add("head_dist.weight", (long long)d.C * d.D);
        add("head_dist.bias", d.C);
    }
    t.push_back({"", 0, off});   // sentinel: total
    return t;
}

struct BlockW {
    const float *n1w, *n1b, *qkvw, *qkvb, *projw, *projb, *n2w, *n2b, *fc1w, *fc1b, *fc2w, *fc2b;
};
struct Weights {
    const float *patchw, *patchb, *cls, *dist, *pos, *normw, *normb, *headw, *headb, *headdw, *headdb;
    BlockW blk[kMaxDepth];
};

static void bind_weights(const te_vit_config* c, const float* base, Weights& w) {
    const std::vector<WEntry> t = weight_table(c);
    size_t i = 0;
    auto next = [&]() { return base + t[i++].offset; };
    w.patchw = next(); w.patchb = next(); w.cls = next();
    w.dist = c->distilled ? next() : nullptr;
    w.pos = next();
    for (int l = 0; l < c->depth; ++l) {
        BlockW& b = w.blk[l];
        b.n1w = next(); b.n1b = next(); b.qkvw = next(); b.qkvb = next(); b.projw = next(); b.projb = next();
        b.n2w = next(); b.n2b = next(); b.fc1w = next(); b.fc1b = next(); b.fc2w = next(); b.fc2b = next();
    }
    w.normw = next(); w.normb = next(); w.headw = next(); w.headb = next();
    w.headdw = c->distilled ? next() : nullptr;
    w.headdb = c->distilled ? next() : nullptr;
}

// ---- derived (tensor-core) weight copies: per block qkv | proj | fc1 | fc2, 4 copies each ---------
struct DerivedW { const float *qkv, *proj, *fc1, *fc2; };
static long long derived_block_floats(const Dims& d) {
    return te_tc_derived_floats(d.D, 3 * d.D) + te_tc_derived_floats(d.D, d.D) + te_tc_derived_floats(d.D, d.F) +
           te_tc_derived_floats(d.F, d.D);
}
static DerivedW bind_derived(const Dims& d, const float* base, int l) {
    DerivedW w = {nullptr, nullptr, nullptr, nullptr};
    if (!base) return w;
    const float* p = base + (long long)l * derived_block_floats(d);
    w.qkv = p;  p += te_tc_derived_floats(d.D, 3 * d.D);
    w.proj = p; p += te_tc_derived_floats(d.D, d.D);
    w.fc1 = p;  p += te_tc_derived_floats(d.D, d.F);
    w.fc2 = p;
    return w;
}

// ---- workspace ---------------------------------------------------------------------------------
struct LayerAct {
    float *x_in, *xn1, *mean1, *rstd1, *qkv, *P, *ctx, *attn_out, *x_mid, *xn2, *mean2, *rstd2, *h, *g, *mlp_out,
        *G, *cam;
};
struct Workspace {
    LayerAct layer[kMaxDepth];
    float *x_last, *xf, *logits, *logits2, *seed, *dpool, *rhead0, *rhead1, *shead;
    float *tD[4], *tF[2], *t3D[2], *tA;
    float *mats, *joint[2];
    float* pix;                       // scratch of the first-layer (pixel) relprop, method="full"
    double* addpart;
    int* index_tmp;
    long long bytes;
};

static void carve(const Dims& d, char* base, Workspace& ws) {
    long long off = 0;
    auto take = [&](long long nfloat) -> float* {
        float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
        off += ((nfloat * 4 + 255) / 256) * 256;
        return p;
    };
    const long long MD = d.M * d.D, MF = d.M * d.F, M3D = d.M * 3LL * d.D;
    const long long AT = (long long)d.B * d.H * d.N * d.NP;
    for (int l = 0; l < d.L; ++l) {
        LayerAct& a = ws.layer[l];
        a.x_in = take(MD); a.xn1 = take(MD); a.mean1 = take(d.M); a.rstd1 = take(d.M);
        a.qkv = take(M3D); a.P = take(AT); a.ctx = take(MD); a.attn_out = take(MD); a.x_mid = take(MD);
        a.xn2 = take(MD); a.mean2 = take(d.M); a.rstd2 = take(d.M); a.h = take(MF); a.g = take(MF);
        a.mlp_out = take(MD); a.G = take(AT); a.cam = take(AT);
    }
    ws.x_last = take(MD); ws.xf = take(MD);
    ws.logits = take((long long)d.B * d.C); ws.logits2 = take((long long)d.B * d.C);
    ws.seed = take((long long)d.B * d.C); ws.shead = take((long long)d.B * d.C);
    ws.dpool = take((long long)d.B * d.D); ws.rhead0 = take((long long)d.B * d.D);
    ws.rhead1 = take((long long)d.B * d.D);
    for (int i = 0; i < 4; ++i) ws.tD[i] = take(MD);
    const long long patches = (long long)d.B * d.npatch * d.KP;
    ws.tF[0] = take(MF > patches ? MF : patches);
    ws.tF[1] = take(MF);
    ws.t3D[0] = take(M3D); ws.t3D[1] = take(M3D);
    ws.tA = take(AT);
    ws.mats = take((long long)d.L * d.B * d.N * d.NP);
    ws.joint[0] = take((long long)d.B * d.N * d.NP);
    ws.joint[1] = take((long long)d.B * d.N * d.NP);
    ws.pix = take(te_patch_relprop_scratch_floats(d.B, d.Cin, d.img, d.P, d.D));
    ws.addpart = reinterpret_cast<double*>(take((long long)d.B * TE_ADD_SPLIT * 3 * 2));
    ws.index_tmp = reinterpret_cast<int*>(take(d.B));
    ws.bytes = off;
}

// ---- GEMM parameter helpers (shared builders live in te_engine_util.h) ----------------------------
using te_util::HeadOp;
using te_util::head_rows;
using te_util::linear_bwd;
using te_util::linear_fwd;

static int check_ws(const te_vit_config* cfg, int batch, void* workspace, long long bytes, Dims& d, Workspace& ws) {
    if (batch <= 0 || !workspace) { te_set_last_error("te_vit: batch <= 0 or null workspace"); return TE_ERR_ARG; }
    if (!make_dims(cfg, batch, d)) return TE_ERR_ARG;
    if (((uintptr_t)workspace & 255u) != 0) { te_set_last_error("te_vit: workspace must be 256-byte aligned"); return TE_ERR_ARG; }
    carve(d, reinterpret_cast<char*>(workspace), ws);
    if (ws.bytes > bytes) { te_set_last_error("te_vit: workspace too small"); return TE_ERR_WORKSPACE; }
    return TE_OK;
}

}  // namespace

// ================================================================================================
// public: weights / workspace description
// ================================================================================================
extern "C" int te_vit_num_weights(const te_vit_config* cfg) {
    const auto t = weight_table(cfg);
    return t.empty() ? TE_ERR_ARG : (int)t.size() - 1;
}
extern "C" const char* te_vit_weight_name(const te_vit_config* cfg, int i) {
    static thread_local std::string s;
    const auto t = weight_table(cfg);
    if (i < 0 || i + 1 >= (int)t.size()) return nullptr;
    s = t[i].name;
    return s.c_str();
}
extern "C" long long te_vit_weight_numel(const te_vit_config* cfg, int i) {
    const auto t = weight_table(cfg);
    if (i < 0 || i + 1 >= (int)t.size()) return TE_ERR_ARG;
    return t[i].numel;
}
extern "C" long long te_vit_weight_offset(const te_vit_config* cfg, int i) {
    const auto t = weight_table(cfg);
    if (i < 0 || i + 1 >= (int)t.size()) return TE_ERR_ARG;
    return t[i].offset;
}
extern "C" long long te_vit_weight_total(const te_vit_config* cfg) {
    const auto t = weight_table(cfg);
    return t.empty() ? TE_ERR_ARG : t.back().offset;
}
extern "C" long long te_vit_workspace_bytes(const te_vit_config* cfg, int batch) {
    Dims d;
    if (batch <= 0 || !make_dims(cfg, batch, d)) return TE_ERR_ARG;
    Workspace ws;
    carve(d, nullptr, ws);
    return ws.bytes;
}

// ================================================================================================
// forward   (ViT_LRP.py:305-322)
// ================================================================================================
extern "C" int te_vit_forward(const te_vit_config* cfg, const float* weights, const float* derived, const float* images,
                              int batch, unsigned flags, float* logits, void* workspace, long long workspace_bytes,
                              void* stream) {
    Dims d; Workspace ws;
    TE_TRY(check_ws(cfg, batch, workspace, workspace_bytes, d, ws));
    if (!weights || !images) { te_set_last_error("te_vit_forward: null pointer"); return TE_ERR_ARG; }
    if ((flags & TE_FLAG_LINEAR_TENSOR_CORES) && !derived) {
        te_set_last_error("te_vit_forward: TE_FLAG_LINEAR_TENSOR_CORES needs the derived weight buffer");
        return TE_ERR_ARG;
    }
    const float* lbase = (flags & TE_FLAG_LINEAR_TENSOR_CORES) ? derived : nullptr;
    // fp16-split forward Linears (te_tc_fwd16.cu).  The block-scaled split of every Linear input lives in buffers that are idle
    // until the backward pass: A = tD[1] (+ scales tD[0]) for the D-wide inputs, B = tF[1] (+ scales tD[2]) for the GELU output.
    // LayerNorm emits the split of what it produces; the attention output and the GELU output go through the pre-pass (emitting
    // the split from the fc1 GELU epilogue was measured: 1.01 ms against 0.56 + 0.2 ms, profiles/r02_results.md).
    const bool f16 = lbase && (flags & TE_FLAG_LINEAR_F16_SPLIT) && d.F >= d.D && te_tc_fwd16_supported(d.M, d.D, 3 * d.D, d.D) &&
                     te_tc_fwd16_supported(d.M, d.D, d.D, d.D) && te_tc_fwd16_supported(d.M, d.D, d.F, d.D) &&
                     te_tc_fwd16_supported(d.M, d.F, d.D, d.F);
    const te_util::F16Split fsA_ready = {ws.tD[1], ws.tD[0], true};
    const te_util::F16Split fsA_pre = {ws.tD[1], ws.tD[0], false};
    const bool gsf = te_engine_gelu_split();
    const te_util::F16Split fsA_fc1 = {ws.tD[1], ws.tD[0], true, gsf ? ws.tF[1] : nullptr, gsf ? ws.tD[2] : nullptr};
    const te_util::F16Split fsB = {ws.tF[1], ws.tD[2], gsf};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Weights w;
    bind_weights(cfg, weights, w);
    const float scale = 1.0f / sqrtf((float)d.dh);

    // patch embedding: conv k=s=P  ==  im2col + GEMM   (PatchEmbed.forward :230-236)
    float* patches = ws.tF[0];
    float* patch_out = ws.tD[3];
    TE_TRY(te_launch_im2col(images, patches, d.B, d.Cin, d.img, d.img, d.P, st));
    TE_TRY(linear_fwd(patches, d.KP, w.patchw, w.patchb, patch_out, nullptr, nullptr, (long long)d.B * d.npatch, d.KP,
                      d.D, TE_EPI_BIAS, st));
    TE_TRY(te_launch_assemble_tokens(patch_out, w.cls, w.dist, w.pos, ws.layer[0].x_in, d.B, d.N, d.D, d.prefix, st));

    for (int l = 0; l < d.L; ++l) {
        LayerAct& a = ws.layer[l];
        const BlockW& bw = w.blk[l];
        float* x_next = (l + 1 < d.L) ? ws.layer[l + 1].x_in : ws.x_last;
        const DerivedW lw = bind_derived(d, lbase, l);
        if (f16)
            TE_TRY(te_launch_layernorm_split(a.x_in, bw.n1w, bw.n1b, a.xn1, a.mean1, a.rstd1, d.M, d.D, cfg->eps_block, ws.tD[1],
                                             ws.tD[0], st));
        else
            TE_TRY(te_launch_layernorm(a.x_in, bw.n1w, bw.n1b, a.xn1, a.mean1, a.rstd1, d.M, d.D, cfg->eps_block, st));
        TE_TRY(te_util::linear_fwd_tc(lw.qkv, a.xn1, d.D, bw.qkvw, bw.qkvb, a.qkv, nullptr, nullptr, d.M, d.D, 3 * d.D,
                                      TE_EPI_BIAS, st, f16 ? &fsA_ready : nullptr));
        const HeadOp q = head_rows(a.qkv, 3 * d.D, d.N, d.dh);
        const HeadOp k = head_rows(a.qkv + d.D, 3 * d.D, d.N, d.dh);
        const HeadOp v = head_rows(a.qkv + 2 * d.D, 3 * d.D, d.N, d.dh);
        // dots = q k^T * scale ; attn = softmax(dots)        (:139-141)
        TE_TRY(te_util::attn_probs((flags & TE_FLAG_ATTN_TENSOR_CORES) != 0, d.B, d.H, d.N, d.NP, d.dh, a.qkv, 3 * d.D,
                                   a.qkv + d.D, 3 * d.D, a.P, scale, st));
        // out = attn v -> 'b h n d -> b n (h d)'              (:147-148)
        TE_TRY(te_util::attn_nk((flags & TE_FLAG_ATTN_TENSOR_CORES) != 0, d.B, d.H, d.N, d.NP, d.dh, a.P, 0, a.qkv + 2 * d.D,
                                3 * d.D, a.ctx, d.D, nullptr, 1.f, TE_EPI_STORE, st));
        // proj + residual add1                                  (:150, :198)
        TE_TRY(te_util::linear_fwd_tc(lw.proj, a.ctx, d.D, bw.projw, bw.projb, a.attn_out, a.x_mid, a.x_in, d.M, d.D, d.D,
                                      TE_EPI_BIAS_ADD, st, f16 ? &fsA_pre : nullptr));
        if (f16)
            TE_TRY(te_launch_layernorm_split(a.x_mid, bw.n2w, bw.n2b, a.xn2, a.mean2, a.rstd2, d.M, d.D, cfg->eps_block, ws.tD[1],
                                             ws.tD[0], st));
        else
            TE_TRY(te_launch_layernorm(a.x_mid, bw.n2w, bw.n2b, a.xn2, a.mean2, a.rstd2, d.M, d.D, cfg->eps_block, st));
        TE_TRY(te_util::linear_fwd_tc(lw.fc1, a.xn2, d.D, bw.fc1w, bw.fc1b, a.h, a.g, nullptr, d.M, d.D, d.F,
                                      TE_EPI_BIAS_GELU, st, f16 ? &fsA_fc1 : nullptr));
        TE_TRY(te_util::linear_fwd_tc(lw.fc2, a.g, d.F, bw.fc2w, bw.fc2b, a.mlp_out, x_next, a.x_mid, d.M, d.F, d.D,
                                      TE_EPI_BIAS_ADD, st, f16 ? &fsB : nullptr));
    }
    // final norm, pool token 0 (and 1), head(s)                (:318-321)
    TE_TRY(te_launch_layernorm(ws.x_last, w.normw, w.normb, ws.xf, nullptr, nullptr, d.M, d.D, cfg->eps_final, st));
    TE_TRY(linear_fwd(ws.xf, d.N * d.D, w.headw, w.headb, ws.logits, nullptr, nullptr, d.B, d.D, d.C, TE_EPI_BIAS, st));
    if (cfg->distilled) {
        TE_TRY(linear_fwd(ws.xf + d.D, d.N * d.D, w.headdw, w.headdb, ws.logits2, nullptr, nullptr, d.B, d.D, d.C,
                          TE_EPI_BIAS, st));
        TE_TRY(te_launch_average2(ws.logits, ws.logits2, ws.logits, (long long)d.B * d.C, st));
    }
    if (logits) {
        if (cudaMemcpyAsync(logits, ws.logits, sizeof(float) * d.B * d.C, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            te_set_last_error("te_vit_forward: logits copy failed");
            return TE_ERR_CUDA;
        }
    }
    return TE_OK;
}

// ================================================================================================
// attribute = class-gradient backward + relprop + aggregation + rollout
// ================================================================================================
extern "C" long long te_vit_derived_total(const te_vit_config* cfg) {
    Dims d;
    if (!make_dims(cfg, 1, d)) return TE_ERR_ARG;
    return (long long)d.L * derived_block_floats(d);
}

extern "C" int te_vit_prepare_derived(const te_vit_config* cfg, const float* weights, float* derived, void* stream) {
    Dims d;
    if (!make_dims(cfg, 1, d)) return TE_ERR_ARG;
    if (!weights || !derived) { te_set_last_error("te_vit_prepare_derived: null pointer"); return TE_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Weights w;
    bind_weights(cfg, weights, w);
    for (int l = 0; l < d.L; ++l) {
        const DerivedW dw = bind_derived(d, derived, l);
        TE_TRY(te_tc_prepare_weights(w.blk[l].qkvw, const_cast<float*>(dw.qkv), d.D, 3 * d.D, st));
        TE_TRY(te_tc_prepare_weights(w.blk[l].projw, const_cast<float*>(dw.proj), d.D, d.D, st));
        TE_TRY(te_tc_prepare_weights(w.blk[l].fc1w, const_cast<float*>(dw.fc1), d.D, d.F, st));
        TE_TRY(te_tc_prepare_weights(w.blk[l].fc2w, const_cast<float*>(dw.fc2), d.F, d.D, st));
    }
    return TE_OK;
}

extern "C" int te_vit_attribute(const te_vit_config* cfg, const float* weights, const float* derived, int batch,
                                int* index, int start_layer, unsigned flags, float* maps, void* workspace,
                                long long workspace_bytes, void* stream) {
    Dims d; Workspace ws;
    TE_TRY(check_ws(cfg, batch, workspace, workspace_bytes, d, ws));
    if (!weights || !index || (!maps && !(flags & TE_FLAG_GRADIENTS_ONLY))) { te_set_last_error("te_vit_attribute: null pointer"); return TE_ERR_ARG; }
    if (start_layer < 0 || start_layer >= d.L) { te_set_last_error("te_vit_attribute: start_layer out of range"); return TE_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Weights w;
    bind_weights(cfg, weights, w);
    const float scale = 1.0f / sqrtf((float)d.dh);
    const long long MD = d.M * d.D;
    const int low = (flags & (TE_FLAG_KEEP_ALL_CAMS | TE_FLAG_RELPROP_TO_INPUT)) ? 0 : start_layer;   // lowest block the relprop must reach
    const float* dbase = (flags & TE_FLAG_ZPLUS_TENSOR_CORES) ? derived : nullptr;
    if ((flags & (TE_FLAG_ZPLUS_TENSOR_CORES | TE_FLAG_LINEAR_TENSOR_CORES)) && !derived) {
        te_set_last_error("te_vit_attribute: tensor-core flags need the derived weight buffer");
        return TE_ERR_ARG;
    }
    const float* lbase = (flags & TE_FLAG_LINEAR_TENSOR_CORES) ? derived : nullptr;
    const bool atc = (flags & TE_FLAG_ATTN_TENSOR_CORES) != 0;
    const bool btf = (flags & TE_FLAG_BACKWARD_TF32) != 0;       // single-pass TF32 backward Linears
    // single-pass fp16 backward Linears: hi-only split of the incoming gradient in tF[1], block scales in t3D[1] (idle until the relprop)
    const te_util::F16Split bfs_v = {ws.tF[1], ws.t3D[1], false};
    const te_util::F16Split* bfs = (lbase && (flags & TE_FLAG_BACKWARD_F16)) ? &bfs_v : nullptr;
    const bool rtf = (flags & TE_FLAG_RELPROP_TF32) != 0;        // single-pass TF32 relevance-side attention contractions
    const bool lrpv = (flags & TE_FLAG_RULES_LRP) != 0;         // rule library of modules/layers_lrp.py (ViT_orig_LRP.py)
    const int zb = ((flags & TE_FLAG_ZPLUS_BF16) ? 1 : 0) | ((flags & TE_FLAG_ZPLUS_S1_BF16) ? 2 : 0) |
                   ((flags & TE_FLAG_ZPLUS_R_F16) ? 4 : 0);                                               // bf16 / fp16 variants of the z+ rule

    // ---- class index and seeds  (ViT_explanation_generator.py:28-35) ---------------------------
    TE_TRY(te_launch_argmax(ws.logits, index, d.B, d.C, /*only_negative=*/1, st));
    const float seedv = cfg->distilled ? 0.5f : 1.0f;   // averaged heads: each gets half of the seed
    TE_TRY(te_launch_onehot(index, ws.seed, d.B, d.C, seedv, st));

    // ---- backward: d logit_c / d attn_l for every block  (what :145's hook captures) -------------
    float* dxa = ws.tD[0]; float* dxb = ws.tD[1]; float* dctx = ws.tD[2]; float* dxn = ws.tD[3];
    float* dF = ws.tF[0]; float* dqkv = ws.t3D[0]; float* dS = ws.tA;
    TE_TRY(te_launch_fill(dxa, 0.f, MD, st));
    {
        // d pooled = seed * W_head ; final LayerNorm backward touches only the pooled token rows
        TE_TRY(linear_bwd(ws.seed, w.headw, ws.dpool, nullptr, d.B, d.D, d.C, TE_EPI_STORE, st));
        TE_TRY(te_launch_layernorm_bwd_strided(ws.dpool, d.D, ws.x_last, (long long)d.N * d.D, w.normw, cfg->eps_final,
                                               dxa, (long long)d.N * d.D, d.B, d.D, st));
        if (cfg->distilled) {
            TE_TRY(linear_bwd(ws.seed, w.headdw, ws.dpool, nullptr, d.B, d.D, d.C, TE_EPI_STORE, st));
            TE_TRY(te_launch_layernorm_bwd_strided(ws.dpool, d.D, ws.x_last + d.D, (long long)d.N * d.D, w.normw,
                                                   cfg->eps_final, dxa + d.D, (long long)d.N * d.D, d.B, d.D, st));
        }
    }
    for (int l = d.L - 1; l >= start_layer; --l) {
        LayerAct& a = ws.layer[l];
        const BlockW& bw = w.blk[l];
        const HeadOp q = head_rows(a.qkv, 3 * d.D, d.N, d.dh);
        const HeadOp k = head_rows(a.qkv + d.D, 3 * d.D, d.N, d.dh);
        const HeadOp v = head_rows(a.qkv + 2 * d.D, 3 * d.D, d.N, d.dh);
        const DerivedW lw = bind_derived(d, lbase, l);
        // mlp branch
        TE_TRY(te_util::linear_bwd_tc(lw.fc2, dxa, bw.fc2w, dF, a.h, d.M, d.F, d.D, TE_EPI_GELU_BWD, st, btf, bfs));
        TE_TRY(te_util::linear_bwd_tc(lw.fc1, dF, bw.fc1w, dxn, nullptr, d.M, d.D, d.F, TE_EPI_STORE, st, btf, bfs));
        TE_TRY(te_launch_layernorm_bwd(dxn, a.x_mid, bw.n2w, a.mean2, a.rstd2, dxa, dxb, d.M, d.D, st));
        // attention branch
        TE_TRY(te_util::linear_bwd_tc(lw.proj, dxb, bw.projw, dctx, nullptr, d.M, d.D, d.D, TE_EPI_STORE, st, btf, bfs));
        TE_TRY(te_util::attn_nn(atc, d.B, d.H, d.N, d.NP, d.dh, dctx, d.D, a.qkv + 2 * d.D, 3 * d.D, a.G, nullptr, 1.f,
                                TE_EPI_STORE, st, btf));                                 // G = dctx v^T
        if (l == start_layer) break;                                                // lower gradients are never read
        TE_TRY(te_util::attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, a.P, 1, dctx, d.D, dqkv + 2 * d.D, 3 * d.D, nullptr, 1.f,
                                TE_EPI_STORE, st, btf));                                 // dV = P^T dctx
        TE_TRY(te_launch_softmax_bwd(a.P, a.G, dS, (long long)d.B * d.H * d.N, d.N, d.NP, scale, st));
        TE_TRY(te_util::attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, dS, 0, a.qkv + d.D, 3 * d.D, dqkv, 3 * d.D, nullptr, 1.f,
                                TE_EPI_STORE, st, btf));                                 // dQ = dS k
        TE_TRY(te_util::attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, dS, 1, a.qkv, 3 * d.D, dqkv + d.D, 3 * d.D, nullptr, 1.f,
                                TE_EPI_STORE, st, btf));                                 // dK = dS^T q
        TE_TRY(te_util::linear_bwd_tc(lw.qkv, dqkv, bw.qkvw, dxn, nullptr, d.M, d.D, 3 * d.D, TE_EPI_STORE, st, btf, bfs));
        TE_TRY(te_launch_layernorm_bwd(dxn, a.x_in, bw.n1w, a.mean1, a.rstd1, dxb, dxa, d.M, d.D, st));
    }

    if (flags & TE_FLAG_GRADIENTS_ONLY) return TE_OK;      // attention-GradCAM baseline: gradients are all it reads

    // ---- relprop  (VisionTransformer.relprop :324-331) --------------------------------------------
    float* R = ws.tD[0]; float* R1 = ws.tD[1]; float* R2 = ws.tD[2]; float* R3 = ws.tD[3];
    float* RF = ws.tF[0]; float* SF = ws.tF[1]; float* S = ws.t3D[0]; float* Rqkv = ws.t3D[1]; float* S1 = ws.tA;
    // head.relprop (z+), pool.relprop (IndexSelect), norm.relprop (identity)
    // z+ rule / Add rule of the selected rule library (layers_ours, or layers_lrp with TE_FLAG_RULES_LRP)
    auto zrule = [&](const float* x, long long ldx, const float* wt, const float* dwt, const float* r, long long ldr, float* out,
                     float* sbuf, long long rows, int in, int outf, const float* y, long long ldy, const float* bias,
                     long long ld_out, float* xabs) -> int {
        if (lrpv) return te_zplus_linear_relprop_lrp(x, ldx, wt, r, ldr, out, sbuf, rows, in, outf, st);
        return te_zplus_linear_relprop_ldr(x, ldx, wt, dwt, r, ldr, out, sbuf, rows, in, outf, st, y, ldy, bias, zb, ld_out, xabs);
    };
    auto addrule = [&](const float* x1, const float* x2, const float* r, float* r1, float* r2) -> int {
        return te_launch_add_relprop(x1, x2, r, r1, r2, lrpv ? nullptr : ws.addpart, d.B, (long long)d.N * d.D, st);
    };
    TE_TRY(zrule(ws.xf, (long long)d.N * d.D, w.headw, nullptr, ws.seed, d.C, ws.rhead0, ws.shead, d.B, d.D, d.C, nullptr, 0,
                 nullptr, 0, nullptr));
    if (cfg->distilled)
        TE_TRY(zrule(ws.xf + d.D, (long long)d.N * d.D, w.headdw, nullptr, ws.seed, d.C, ws.rhead1, ws.shead, d.B, d.D, d.C,
                     nullptr, 0, nullptr, 0, nullptr));
    TE_TRY(te_launch_index_select_relprop(ws.xf, ws.rhead0, cfg->distilled ? ws.rhead1 : nullptr, R, d.B, d.N, d.D, st));

    for (int l = d.L - 1; l >= low; --l) {
        LayerAct& a = ws.layer[l];
        const BlockW& bw = w.blk[l];
        const DerivedW dw = bind_derived(d, dbase, l);
        const HeadOp q = head_rows(a.qkv, 3 * d.D, d.N, d.dh);
        const HeadOp k = head_rows(a.qkv + d.D, 3 * d.D, d.N, d.dh);
        const HeadOp v = head_rows(a.qkv + 2 * d.D, 3 * d.D, d.N, d.dh);
        // Block.relprop :203-213
        // In the TOP block the relevance that enters is non-zero only in the pooled token's row (IndexSelect.relprop, a7),
        // and every rule down to the proj rule is row-wise: Add / Clone map a zero row to a zero row, the z+ rule computes
        // each output row from the same input row.  So the three z+ rules of the top block run on the B pooled rows only
        // (row stride N*D / N*F) — exact (SURVEY.md 8a "structural savings"), bit-identical rows, 1/N of the work.
        const bool top = (l == d.L - 1) && !cfg->distilled && !lrpv && te_engine_cls_rows();
        const long long zr = top ? d.B : d.M;                          // rows the z+ rules of this block touch
        const long long sD = top ? (long long)d.N * d.D : d.D, sF = top ? (long long)d.N * d.F : d.F;
        TE_TRY(addrule(a.x_mid, a.mlp_out, R, R1, R2));                                                            // add2
        TE_TRY(zrule(a.g, sF, bw.fc2w, dw.fc2, R2, sD, RF, S, zr, d.F, d.D, a.mlp_out, sD, bw.fc2b, sF, SF));      // fc2 ; GELU id
        TE_TRY(zrule(a.xn2, sD, bw.fc1w, dw.fc1, RF, sF, R2, SF, zr, d.D, d.F, a.h, sF, bw.fc1b, sD, S));          // fc1 ; norm2 id
        TE_TRY(te_launch_clone_relprop(a.x_mid, R1, R2, nullptr, R, MD, st));                                      // clone2
        TE_TRY(addrule(a.x_in, a.attn_out, R, R1, R2));                                                            // add1
        // Attention.relprop :154-177
        if (top) TE_TRY(te_launch_fill(R3, 0.f, MD, st));                                   // rows the strided rule does not write
        TE_TRY(zrule(a.ctx, sD, bw.projw, dw.proj, R2, sD, R3, S, zr, d.D, d.D, a.attn_out, sD, bw.projb, sD, S + MD));   // proj
        // matmul2 rule: Z = attn v is the saved ctx itself (bit-identical recomputation in the reference)
        TE_TRY(te_launch_sd(R3, a.ctx, S, MD, st));
        TE_TRY(te_util::attn_nn(atc, d.B, d.H, d.N, d.NP, d.dh, S, d.D, a.qkv + 2 * d.D, 3 * d.D, a.cam, a.P, 0.5f,
                                TE_EPI_MUL, st, rtf));                                   // attn_cam = (P * (S v^T)) / 2   :160-165
        if (l == low && !(flags & TE_FLAG_RELPROP_TO_INPUT)) break;                                                        // nothing below is consumed
        TE_TRY(te_util::attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, a.P, 1, S, d.D, Rqkv + 2 * d.D, 3 * d.D, a.qkv + 2 * d.D, 0.5f,
                                TE_EPI_MUL, st, rtf));                                   // cam_v
        // matmul1 rule (unscaled Z = q k^T)  :170-173
        TE_TRY(te_util::attn_nn(atc, d.B, d.H, d.N, d.NP, d.dh, a.qkv, 3 * d.D, a.qkv + d.D, 3 * d.D, S1, a.cam, 1.f,
                                TE_EPI_SD, st));
        TE_TRY(te_util::attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, S1, 0, a.qkv + d.D, 3 * d.D, Rqkv, 3 * d.D, a.qkv, 0.5f,
                                TE_EPI_MUL, st, rtf));                                   // cam_q
        TE_TRY(te_util::attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, S1, 1, a.qkv, 3 * d.D, Rqkv + d.D, 3 * d.D, a.qkv + d.D, 0.5f,
                                TE_EPI_MUL, st, rtf));                                   // cam_k
        TE_TRY(zrule(a.xn1, d.D, bw.qkvw, dw.qkv, Rqkv, 3 * d.D, R2, S, d.M, d.D, 3 * d.D, a.qkv, 3 * d.D, bw.qkvb, 0, RF));   // qkv ; norm1 id
        TE_TRY(te_launch_clone_relprop(a.x_in, R1, R2, nullptr, R, MD, st));                                       // clone1
    }

    // ---- aggregation + rollout  (:357-368) ---------------------------------------------------------
    TE_TRY(te_rollout_layers(ws.layer[0].G, ws.layer[0].cam,
                             d.L > 1 ? (long long)(ws.layer[1].G - ws.layer[0].G) : 0, d.L, d.B, d.H, d.N, d.NP, d.NP,
                             start_layer, /*normalize=*/0, flags, ws.mats, ws.joint[0], ws.joint[1], nullptr, maps,
                             d.prefix, /*bert_fix=*/0, st));
    return TE_OK;
}

// ================================================================================================
// method="full": relevance of every input pixel   (ViT_LRP.py:337-343)
//   (cam, _) = self.add.relprop(cam) ; cam = cam[:, 1:] ; cam = self.patch_embed.relprop(cam) ; cam.sum(dim=1)
// Precondition: te_vit_forward + te_vit_attribute(flags | TE_FLAG_RELPROP_TO_INPUT) on this workspace and these images.
// ================================================================================================
extern "C" int te_vit_relprop_pixels(const te_vit_config* cfg, const float* weights, const float* images, int batch,
                                     float* pixel_maps, float* pixel_relevance, void* workspace, long long workspace_bytes,
                                     void* stream) {
    return te_vit_relprop_pixels_ex(cfg, weights, images, batch, 0u, pixel_maps, pixel_relevance, workspace, workspace_bytes,
                                    stream);
}

extern "C" int te_vit_relprop_pixels_ex(const te_vit_config* cfg, const float* weights, const float* images, int batch,
                                        unsigned flags, float* pixel_maps, float* pixel_relevance, void* workspace,
                                        long long workspace_bytes, void* stream) {
    Dims d; Workspace ws;
    TE_TRY(check_ws(cfg, batch, workspace, workspace_bytes, d, ws));
    if (!weights || !images || (!pixel_maps && !pixel_relevance)) { te_set_last_error("te_vit_relprop_pixels: null pointer"); return TE_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Weights w;
    bind_weights(cfg, weights, w);
    float* R = ws.tD[0];                       // relevance at the encoder input (clone1 of block 0)
    float* tokens = ws.tD[1];                  // cat(cls[,dist], patch_embed(x)): the first operand of self.add (:311)
    float* Rtok = ws.tD[2];
    float* patches = ws.tF[0];
    float* patch_out = ws.tD[3];
    TE_TRY(te_launch_im2col(images, patches, d.B, d.Cin, d.img, d.img, d.P, st));
    TE_TRY(linear_fwd(patches, d.KP, w.patchw, w.patchb, patch_out, nullptr, nullptr, (long long)d.B * d.npatch, d.KP,
                      d.D, TE_EPI_BIAS, st));
    TE_TRY(te_launch_assemble_tokens(patch_out, w.cls, w.dist, nullptr, tokens, d.B, d.N, d.D, d.prefix, st));
    // self.add.relprop: x2 = pos_embed, shared by every sample; only the tokens' share is consumed
    TE_TRY(te_launch_add_relprop_ex(tokens, w.pos, 0, R, Rtok, nullptr, (flags & TE_FLAG_RULES_LRP) ? nullptr : ws.addpart, d.B,
                                    (long long)d.N * d.D, st));
    // cam[:, 1:] -> PatchEmbed.relprop (:238-242) -> Conv2d z^B rule -> sum over channels
    return te_patch_relprop_run(images, w.patchw, Rtok + (long long)d.prefix * d.D, (long long)d.N * d.D, d.B, d.Cin, d.img,
                                d.P, d.D, ws.pix, pixel_relevance, pixel_maps, st);
}

extern "C" int te_vit_explain(const te_vit_config* cfg, const float* weights, const float* derived, const float* images,
                              int batch, int* index, int start_layer, unsigned flags, float* maps, float* logits,
                              void* workspace, long long workspace_bytes, void* stream) {
    TE_TRY(te_vit_forward(cfg, weights, derived, images, batch, flags, logits, workspace, workspace_bytes, stream));
    return te_vit_attribute(cfg, weights, derived, batch, index, start_layer, flags, maps, workspace, workspace_bytes,
                            stream);
}

extern "C" int te_vit_tensor(const te_vit_config* cfg, int batch, void* workspace, const char* name, int layer,
                             float** ptr, long long dims[4], long long strides[4]) {
    Dims d; Workspace ws;
    if (!workspace || !name || !ptr) return TE_ERR_ARG;
    if (batch <= 0 || !make_dims(cfg, batch, d)) return TE_ERR_ARG;
    carve(d, reinterpret_cast<char*>(workspace), ws);
    const std::string n(name);
    auto set = [&](float* p, long long d0, long long d1, long long d2, long long d3, long long s0, long long s1,
                   long long s2, long long s3) {
        *ptr = p; dims[0] = d0; dims[1] = d1; dims[2] = d2; dims[3] = d3;
        strides[0] = s0; strides[1] = s1; strides[2] = s2; strides[3] = s3;
        return TE_OK;
    };
    if (n == "logits") return set(ws.logits, d.B, d.C, 1, 1, d.C, 1, 1, 1);
    if (n == "relevance_in") return set(ws.tD[0], d.B, d.N, d.D, 1, (long long)d.N * d.D, d.D, 1, 1);
    if (n == "rollout_mats") return set(ws.mats, d.L, d.B, d.N, d.N, (long long)d.B * d.N * d.NP, (long long)d.N * d.NP, d.NP, 1);
    // scratch of the last attribute() call (debug / diagnostics): tmp_d0..3 [B,N,D], tmp_f0..1 [B,N,F], tmp_3d0..1 [B,N,3D]
    if (n.rfind("tmp_d", 0) == 0 && n.size() == 6 && n[5] >= '0' && n[5] <= '3')
        return set(ws.tD[n[5] - '0'], d.B, d.N, d.D, 1, (long long)d.N * d.D, d.D, 1, 1);
    if (n.rfind("tmp_f", 0) == 0 && n.size() == 6 && n[5] >= '0' && n[5] <= '1')
        return set(ws.tF[n[5] - '0'], d.B, d.N, d.F, 1, (long long)d.N * d.F, d.F, 1, 1);
    if (n.rfind("tmp_3d", 0) == 0 && n.size() == 7 && n[6] >= '0' && n[6] <= '1')
        return set(ws.t3D[n[6] - '0'], d.B, d.N, 3LL * d.D, 1, (long long)d.N * 3 * d.D, 3LL * d.D, 1, 1);
    if (layer < 0 || layer >= d.L) { te_set_last_error("te_vit_tensor: layer out of range"); return TE_ERR_ARG; }
    LayerAct& a = ws.layer[layer];
    const long long hs = (long long)d.N * d.NP, bs = hs * d.H;
    if (n == "attn") return set(a.P, d.B, d.H, d.N, d.N, bs, hs, d.NP, 1);
    if (n == "attn_grad") return set(a.G, d.B, d.H, d.N, d.N, bs, hs, d.NP, 1);
    if (n == "attn_cam") return set(a.cam, d.B, d.H, d.N, d.N, bs, hs, d.NP, 1);
    if (n == "qkv") return set(a.qkv, d.B, d.N, 3LL * d.D, 1, (long long)d.N * 3 * d.D, 3LL * d.D, 1, 1);
    if (n == "x_in") return set(a.x_in, d.B, d.N, d.D, 1, (long long)d.N * d.D, d.D, 1, 1);
    if (n == "ctx") return set(a.ctx, d.B, d.N, d.D, 1, (long long)d.N * d.D, d.D, 1, 1);
    te_set_last_error("te_vit_tensor: unknown tensor name");
    return TE_ERR_ARG;
}
