// BERT (post-LN encoder + pooler + classifier) transformer-attribution engine.
//
// Reference wiring: BERT_explainability/modules/BERT/BERT.py (BertEmbeddings :61-85, BertSelfAttention :307-409,
// BertSelfOutput :420-434, BertIntermediate :446-456, BertOutput :467-487, BertLayer :498-530, BertPooler :169-190,
// BertModel.relprop :645-651), BertForSequenceClassification.py:23-88, ExplanationGenerator.py:7-59.
// The additive attention mask (1-mask)*-10000 and head_mask = None are transformers==3.5.1 behaviour
// (un-vendored dependency; call sites BERT.py:598,616) restated here.
//
// q/k/v Linears are packed into one [3D, D] weight so that the forward is one GEMM and the per-head slices are
// addressed in place exactly like the ViT engine; their three z+ rules stay separate (Clone(3) needs them apart).
#include <string.h>
#include <string>
#include <vector>

#include "../../include/te_b200.h"
#include "te_engine_util.h"
#include "te_gemm_tc.h"
#include "te_kernels.h"
#include "te_rollout.h"
#include "te_zplus.h"

using namespace te_util;

namespace {

constexpr int kMaxDepth = 64;

struct Dims {
    int B, N, NP, D, H, dh, F, C, L, V, P, T;
    long long M;
    float eps;
};

static bool make_dims(const te_bert_config* c, int B, int S, Dims& d) {
    if (!c || c->layers <= 0 || c->layers > kMaxDepth || c->heads <= 0 || c->hidden % c->heads != 0 || c->hidden % 8 != 0 ||
        c->intermediate % 4 != 0 || c->num_labels <= 0 || c->vocab_size <= 0 || c->max_position <= 0 ||
        c->type_vocab <= 0) {
        te_set_last_error("te_bert: invalid config");
        return false;
    }
    if (S <= 0 || S > c->max_position) { te_set_last_error("te_bert: sequence length out of range"); return false; }
    d.B = B; d.N = S; d.NP = (S + 3) & ~3; d.D = c->hidden; d.H = c->heads; d.dh = c->hidden / c->heads;
    d.F = c->intermediate; d.C = c->num_labels; d.L = c->layers; d.V = c->vocab_size; d.P = c->max_position;
    d.T = c->type_vocab; d.M = (long long)B * S; d.eps = c->layer_norm_eps;
    if (d.dh % 4 != 0) { te_set_last_error("te_bert: head_dim % 4 != 0"); return false; }
    return true;
}

// ---- flat weight buffer (HF state_dict keys) -------------------------------------------------------
struct WEntry { std::string name; long long numel; long long offset; };

static std::vector<WEntry> weight_table(const te_bert_config* c) {
    std::vector<WEntry> t;
    Dims d;
    if (!make_dims(c, 1, 1, d)) return t;
    long long off = 0;
    auto add = [&](const std::string& n, long long numel, bool pad = true) {
        t.push_back({n, numel, off});
        off += pad ? ((numel + 31) & ~31LL) : numel;
    };
    const std::string E = "bert.embeddings.";
    add(E + "word_embeddings.weight", (long long)d.V * d.D);
    add(E + "position_embeddings.weight", (long long)d.P * d.D);
    add(E + "token_type_embeddings.weight", (long long)d.T * d.D);
    add(E + "LayerNorm.weight", d.D);
    add(E + "LayerNorm.bias", d.D);
    for (int i = 0; i < d.L; ++i) {
        const std::string L = "bert.encoder.layer." + std::to_string(i) + ".";
        // query | key | value stored back to back (no padding): one packed [3D, D] weight, [3D] bias
        add(L + "attention.self.query.weight", (long long)d.D * d.D, false);
        add(L + "attention.self.key.weight", (long long)d.D * d.D, false);
        add(L + "attention.self.value.weight", (long long)d.D * d.D, true);
        add(L + "attention.self.query.bias", d.D, false);
        add(L + "attention.self.key.bias", d.D, false);
        add(L + "attention.self.value.bias", d.D, true);
        add(L + "attention.output.dense.weight", (long long)d.D * d.D);
        add(L + "attention.output.dense.bias", d.D);
        add(L + "attention.output.LayerNorm.weight", d.D);
        add(L + "attention.output.LayerNorm.bias", d.D);
        add(L + "intermediate.dense.weight", (long long)d.F * d.D);
        add(L + "intermediate.dense.bias", d.F);
        add(L + "output.dense.weight", (long long)d.D * d.F);
        add(L + "output.dense.bias", d.D);
        add(L + "output.LayerNorm.weight", d.D);
        add(L + "output.LayerNorm.bias", d.D);
    }
    add("bert.pooler.dense.weight", (long long)d.D * d.D);
    add("bert.pooler.dense.bias", d.D);
    add("classifier.weight", (long long)d.C * d.D);
    add("classifier.bias", d.C);
    t.push_back({"", 0, off});
    return t;
}

struct LayerW {
    const float *qkvw, *qkvb, *ow, *ob, *ln1w, *ln1b, *w1, *b1, *w2, *b2, *ln2w, *ln2b;
};
struct Weights {
    const float *word, *pos, *type, *elnw, *elnb, *poolw, *poolb, *clsw, *clsb;
    LayerW layer[kMaxDepth];
};

static void bind_weights(const te_bert_config* c, const float* base, Weights& w) {
    const std::vector<WEntry> t = weight_table(c);
    size_t i = 0;
    auto next = [&]() { return base + t[i++].offset; };
    w.word = next(); w.pos = next(); w.type = next(); w.elnw = next(); w.elnb = next();
    for (int l = 0; l < c->layers; ++l) {
        LayerW& y = w.layer[l];
        y.qkvw = next(); next(); next();
        y.qkvb = next(); next(); next();
        y.ow = next(); y.ob = next(); y.ln1w = next(); y.ln1b = next();
        y.w1 = next(); y.b1 = next(); y.w2 = next(); y.b2 = next(); y.ln2w = next(); y.ln2b = next();
    }
    w.poolw = next(); w.poolb = next(); w.clsw = next(); w.clsb = next();
}

// ---- derived tensor-core copies: per layer q | k | v | o | w1 | w2 ---------------------------------------
// (q, k, v: operands of their three z+ rules; qkv: the packed [3D, D] weight for the forward / backward GEMMs)
struct DerivedW { const float *q, *k, *v, *o, *w1, *w2, *qkv; };
static long long derived_layer_floats(const Dims& d) {
    return 4 * te_tc_derived_floats(d.D, d.D) + te_tc_derived_floats(d.D, d.F) + te_tc_derived_floats(d.F, d.D) +
           te_tc_derived_floats(d.D, 3 * d.D);
}
static DerivedW bind_derived(const Dims& d, const float* base, int l) {
    DerivedW w = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!base) return w;
    const float* p = base + (long long)l * derived_layer_floats(d);
    const long long dd = te_tc_derived_floats(d.D, d.D);
    w.q = p; w.k = p + dd; w.v = p + 2 * dd; w.o = p + 3 * dd;
    w.w1 = p + 4 * dd;
    w.w2 = w.w1 + te_tc_derived_floats(d.D, d.F);
    w.qkv = w.w2 + te_tc_derived_floats(d.F, d.D);
    return w;
}

// ---- workspace ---------------------------------------------------------------------------------------
struct LayerAct {
    float *h, *qkv, *P, *ctx, *d1, *s1, *ao, *mean1, *rstd1, *hpre, *g, *d2, *s2, *mean2, *rstd2, *G, *cam;
};
struct Workspace {
    LayerAct layer[kMaxDepth];
    float *h_last, *maskadd, *pd, *pooled, *logits, *seed, *dpool, *dpd, *dfirst, *rpool, *rfirst, *shead;
    float *tD[4], *tF[2], *t3D[2], *tA[2];
    float *mats, *joint[2];
    double* addpart;
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
        a.h = take(MD); a.qkv = take(M3D); a.P = take(AT); a.ctx = take(MD); a.d1 = take(MD); a.s1 = take(MD);
        a.ao = take(MD); a.mean1 = take(d.M); a.rstd1 = take(d.M); a.hpre = take(MF); a.g = take(MF); a.d2 = take(MD);
        a.s2 = take(MD); a.mean2 = take(d.M); a.rstd2 = take(d.M); a.G = take(AT); a.cam = take(AT);
    }
    ws.h_last = take(MD);
    ws.maskadd = take((long long)d.B * d.N);
    const long long BD = (long long)d.B * d.D, BC = (long long)d.B * d.C;
    ws.pd = take(BD); ws.pooled = take(BD); ws.dpool = take(BD); ws.dpd = take(BD); ws.dfirst = take(BD);
    ws.rpool = take(BD); ws.rfirst = take(BD);
    ws.logits = take(BC); ws.seed = take(BC); ws.shead = take(BC > BD ? BC : BD);
    for (int i = 0; i < 4; ++i) ws.tD[i] = take(MD);
    ws.tF[0] = take(MF); ws.tF[1] = take(MF);
    ws.t3D[0] = take(M3D); ws.t3D[1] = take(M3D);
    ws.tA[0] = take(AT); ws.tA[1] = take(AT);
    ws.mats = take((long long)d.L * d.B * d.N * d.NP);
    ws.joint[0] = take((long long)d.B * d.N * d.NP);
    ws.joint[1] = take((long long)d.B * d.N * d.NP);
    ws.addpart = reinterpret_cast<double*>(take((long long)d.B * TE_ADD_SPLIT * 3 * 2));
    ws.bytes = off;
}

static int check_ws(const te_bert_config* cfg, int batch, int seq, void* workspace, long long bytes, Dims& d,
                    Workspace& ws) {
    if (batch <= 0 || !workspace) { te_set_last_error("te_bert: batch <= 0 or null workspace"); return TE_ERR_ARG; }
    if (!make_dims(cfg, batch, seq, d)) return TE_ERR_ARG;
    if (((uintptr_t)workspace & 255u) != 0) { te_set_last_error("te_bert: workspace must be 256-byte aligned"); return TE_ERR_ARG; }
    carve(d, reinterpret_cast<char*>(workspace), ws);
    if (ws.bytes > bytes) { te_set_last_error("te_bert: workspace too small"); return TE_ERR_WORKSPACE; }
    return TE_OK;
}

}  // namespace

// =====================================================================================================
extern "C" int te_bert_num_weights(const te_bert_config* cfg) {
    const auto t = weight_table(cfg);
    return t.empty() ? TE_ERR_ARG : (int)t.size() - 1;
}
extern "C" const char* te_bert_weight_name(const te_bert_config* cfg, int i) {
    static thread_local std::string s;
    const auto t = weight_table(cfg);
    if (i < 0 || i + 1 >= (int)t.size()) return nullptr;
    s = t[i].name;
    return s.c_str();
}
extern "C" long long te_bert_weight_numel(const te_bert_config* cfg, int i) {
    const auto t = weight_table(cfg);
    if (i < 0 || i + 1 >= (int)t.size()) return TE_ERR_ARG;
    return t[i].numel;
}
extern "C" long long te_bert_weight_offset(const te_bert_config* cfg, int i) {
    const auto t = weight_table(cfg);
    if (i < 0 || i + 1 >= (int)t.size()) return TE_ERR_ARG;
    return t[i].offset;
}
extern "C" long long te_bert_weight_total(const te_bert_config* cfg) {
    const auto t = weight_table(cfg);
    return t.empty() ? TE_ERR_ARG : t.back().offset;
}
extern "C" long long te_bert_workspace_bytes(const te_bert_config* cfg, int batch, int seq) {
    Dims d;
    if (batch <= 0 || !make_dims(cfg, batch, seq, d)) return TE_ERR_ARG;
    Workspace ws;
    carve(d, nullptr, ws);
    return ws.bytes;
}
extern "C" long long te_bert_derived_total(const te_bert_config* cfg) {
    Dims d;
    if (!make_dims(cfg, 1, 1, d)) return TE_ERR_ARG;
    return (long long)d.L * derived_layer_floats(d);
}
extern "C" int te_bert_prepare_derived(const te_bert_config* cfg, const float* weights, float* derived, void* stream) {
    Dims d;
    if (!make_dims(cfg, 1, 1, d)) return TE_ERR_ARG;
    if (!weights || !derived) { te_set_last_error("te_bert_prepare_derived: null pointer"); return TE_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Weights w;
    bind_weights(cfg, weights, w);
    const long long DD = (long long)d.D * d.D;
    for (int l = 0; l < d.L; ++l) {
        const DerivedW dw = bind_derived(d, derived, l);
        TE_TRY(te_tc_prepare_weights(w.layer[l].qkvw, const_cast<float*>(dw.q), d.D, d.D, st));
        TE_TRY(te_tc_prepare_weights(w.layer[l].qkvw + DD, const_cast<float*>(dw.k), d.D, d.D, st));
        TE_TRY(te_tc_prepare_weights(w.layer[l].qkvw + 2 * DD, const_cast<float*>(dw.v), d.D, d.D, st));
        TE_TRY(te_tc_prepare_weights(w.layer[l].ow, const_cast<float*>(dw.o), d.D, d.D, st));
        TE_TRY(te_tc_prepare_weights(w.layer[l].w1, const_cast<float*>(dw.w1), d.D, d.F, st));
        TE_TRY(te_tc_prepare_weights(w.layer[l].w2, const_cast<float*>(dw.w2), d.F, d.D, st));
        TE_TRY(te_tc_prepare_weights(w.layer[l].qkvw, const_cast<float*>(dw.qkv), d.D, 3 * d.D, st));
    }
    return TE_OK;
}

// =====================================================================================================
// forward  (BertForSequenceClassification.forward -> BertModel.forward)
// =====================================================================================================
extern "C" int te_bert_forward(const te_bert_config* cfg, const float* weights, const float* derived,
                               const long long* input_ids, const long long* attention_mask, int batch, int seq,
                               unsigned flags, float* logits, void* workspace, long long workspace_bytes, void* stream) {
    Dims d; Workspace ws;
    TE_TRY(check_ws(cfg, batch, seq, workspace, workspace_bytes, d, ws));
    if (!weights || !input_ids || !attention_mask) { te_set_last_error("te_bert_forward: null pointer"); return TE_ERR_ARG; }
    if ((flags & TE_FLAG_LINEAR_TENSOR_CORES) && !derived) {
        te_set_last_error("te_bert_forward: TE_FLAG_LINEAR_TENSOR_CORES needs the derived weight buffer");
        return TE_ERR_ARG;
    }
    const float* lbase = (flags & TE_FLAG_LINEAR_TENSOR_CORES) ? derived : nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // fp16-split forward Linears (te_tc_fwd16.cu): split of the D-wide inputs in A = tD[1] (+ scales tD[2]), of the GELU output in
    // B = tF[1] (+ scales tD[3]); all idle until the backward pass.  Every LayerNorm emits the split of its output (the hidden
    // state feeds the next layer's qkv); the attention context and the GELU output go through the pre-pass.
    const bool f16 = lbase && (flags & TE_FLAG_LINEAR_F16_SPLIT) && d.F >= d.D && te_tc_fwd16_supported(d.M, d.D, 3 * d.D, d.D) &&
                     te_tc_fwd16_supported(d.M, d.D, d.D, d.D) && te_tc_fwd16_supported(d.M, d.D, d.F, d.D) &&
                     te_tc_fwd16_supported(d.M, d.F, d.D, d.F);
    const te_util::F16Split fsA_ready = {ws.tD[1], ws.tD[2], true};
    const te_util::F16Split fsA_pre = {ws.tD[1], ws.tD[2], false};
    const bool gsf = te_engine_gelu_split();
    const te_util::F16Split fsA_w1 = {ws.tD[1], ws.tD[2], true, gsf ? ws.tF[1] : nullptr, gsf ? ws.tD[3] : nullptr};
    const te_util::F16Split fsB = {ws.tF[1], ws.tD[3], gsf};
    auto layernorm = [&](const float* x, const float* g, const float* b, float* y, float* mean, float* rstd) {
        return f16 ? te_launch_layernorm_split(x, g, b, y, mean, rstd, d.M, d.D, d.eps, ws.tD[1], ws.tD[2], st)
                   : te_launch_layernorm(x, g, b, y, mean, rstd, d.M, d.D, d.eps, st);
    };
    Weights w;
    bind_weights(cfg, weights, w);
    const float scale = 1.0f / sqrtf((float)d.dh);

    TE_TRY(te_launch_bert_embed(input_ids, w.word, w.pos, w.type, ws.tD[0], d.B, d.N, d.D, d.V, st));
    TE_TRY(layernorm(ws.tD[0], w.elnw, w.elnb, ws.layer[0].h, nullptr, nullptr));
    TE_TRY(te_launch_bert_mask(attention_mask, ws.maskadd, (long long)d.B * d.N, st));

    for (int l = 0; l < d.L; ++l) {
        LayerAct& a = ws.layer[l];
        const LayerW& lw = w.layer[l];
        float* h_next = (l + 1 < d.L) ? ws.layer[l + 1].h : ws.h_last;
        const DerivedW tw = bind_derived(d, lbase, l);
        TE_TRY(linear_fwd_tc(tw.qkv, a.h, d.D, lw.qkvw, lw.qkvb, a.qkv, nullptr, nullptr, d.M, d.D, 3 * d.D, TE_EPI_BIAS, st,
                             f16 ? &fsA_ready : nullptr));
        const HeadOp q = head_rows(a.qkv, 3 * d.D, d.N, d.dh);
        const HeadOp k = head_rows(a.qkv + d.D, 3 * d.D, d.N, d.dh);
        const HeadOp v = head_rows(a.qkv + 2 * d.D, 3 * d.D, d.N, d.dh);
        // scores = q k^T / sqrt(d) ; + extended mask ; softmax      (:338-345)
        TE_TRY(attn_nn((flags & TE_FLAG_ATTN_TENSOR_CORES) != 0, d.B, d.H, d.N, d.NP, d.dh, a.qkv, 3 * d.D, a.qkv + d.D,
                       3 * d.D, a.P, nullptr, scale, TE_EPI_STORE, st));
        TE_TRY(te_launch_softmax_masked(a.P, (long long)d.B * d.H * d.N, d.N, d.NP, ws.maskadd, (long long)d.H * d.N, st));
        TE_TRY(attn_nk((flags & TE_FLAG_ATTN_TENSOR_CORES) != 0, d.B, d.H, d.N, d.NP, d.dh, a.P, 0, a.qkv + 2 * d.D, 3 * d.D,
                       a.ctx, d.D, nullptr, 1.f, TE_EPI_STORE, st));
        // BertSelfOutput: dense -> add([dense, input]) -> LayerNorm
        TE_TRY(linear_fwd_tc(tw.o, a.ctx, d.D, lw.ow, lw.ob, a.d1, a.s1, a.h, d.M, d.D, d.D, TE_EPI_BIAS_ADD, st,
                             f16 ? &fsA_pre : nullptr));
        TE_TRY(layernorm(a.s1, lw.ln1w, lw.ln1b, a.ao, a.mean1, a.rstd1));
        // BertIntermediate (dense + GELU), BertOutput (dense -> add -> LayerNorm)
        TE_TRY(linear_fwd_tc(tw.w1, a.ao, d.D, lw.w1, lw.b1, a.hpre, a.g, nullptr, d.M, d.D, d.F, TE_EPI_BIAS_GELU, st,
                             f16 ? &fsA_w1 : nullptr));
        TE_TRY(linear_fwd_tc(tw.w2, a.g, d.F, lw.w2, lw.b2, a.d2, a.s2, a.ao, d.M, d.F, d.D, TE_EPI_BIAS_ADD, st,
                             f16 ? &fsB : nullptr));
        TE_TRY(layernorm(a.s2, lw.ln2w, lw.ln2b, h_next, a.mean2, a.rstd2));
    }
    // pooler (first token -> dense -> tanh), classifier
    TE_TRY(linear_fwd(ws.h_last, d.N * d.D, w.poolw, w.poolb, ws.pd, nullptr, nullptr, d.B, d.D, d.D, TE_EPI_BIAS, st));
    TE_TRY(te_launch_tanh(ws.pd, ws.pooled, (long long)d.B * d.D, st));
    TE_TRY(linear_fwd(ws.pooled, d.D, w.clsw, w.clsb, ws.logits, nullptr, nullptr, d.B, d.D, d.C, TE_EPI_BIAS, st));
    if (logits && cudaMemcpyAsync(logits, ws.logits, sizeof(float) * d.B * d.C, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        te_set_last_error("te_bert_forward: logits copy failed");
        return TE_ERR_CUDA;
    }
    return TE_OK;
}

// =====================================================================================================
// attribute = class-gradient backward + relprop + normalised rollout   (Generator.generate_LRP :33-59)
// =====================================================================================================
extern "C" int te_bert_attribute(const te_bert_config* cfg, const float* weights, const float* derived, int batch,
                                 int seq, int* index, int start_layer, unsigned flags, float* maps, void* workspace,
                                 long long workspace_bytes, void* stream) {
    Dims d; Workspace ws;
    TE_TRY(check_ws(cfg, batch, seq, workspace, workspace_bytes, d, ws));
    if (!weights || !index || (!maps && !(flags & TE_FLAG_GRADIENTS_ONLY))) { te_set_last_error("te_bert_attribute: null pointer"); return TE_ERR_ARG; }
    if (start_layer < 0 || start_layer >= d.L) { te_set_last_error("te_bert_attribute: start_layer out of range"); return TE_ERR_ARG; }
    if ((flags & (TE_FLAG_ZPLUS_TENSOR_CORES | TE_FLAG_LINEAR_TENSOR_CORES)) && !derived) {
        te_set_last_error("te_bert_attribute: tensor-core flags need the derived weight buffer");
        return TE_ERR_ARG;
    }
    const float* lbase = (flags & TE_FLAG_LINEAR_TENSOR_CORES) ? derived : nullptr;
    const bool atc = (flags & TE_FLAG_ATTN_TENSOR_CORES) != 0;
    const bool btf = (flags & TE_FLAG_BACKWARD_TF32) != 0;       // single-pass TF32 backward Linears
    // single-pass fp16 backward Linears: hi-only split of the incoming gradient in tF[1], block scales in t3D[1] (idle until the relprop)
    const te_util::F16Split bfs_v = {ws.tF[1], ws.t3D[1], false};
    const te_util::F16Split* bfs = (lbase && (flags & TE_FLAG_BACKWARD_F16)) ? &bfs_v : nullptr;
    const bool rtf = (flags & TE_FLAG_RELPROP_TF32) != 0;        // single-pass TF32 relevance-side attention contractions
    const int zb = ((flags & TE_FLAG_ZPLUS_BF16) ? 1 : 0) | ((flags & TE_FLAG_ZPLUS_S1_BF16) ? 2 : 0) |
                   ((flags & TE_FLAG_ZPLUS_R_F16) ? 4 : 0);                                               // bf16 / fp16 variants of the z+ rule
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Weights w;
    bind_weights(cfg, weights, w);
    const float* dbase = (flags & TE_FLAG_ZPLUS_TENSOR_CORES) ? derived : nullptr;
    const float scale = 1.0f / sqrtf((float)d.dh);
    const long long MD = d.M * d.D, DD = (long long)d.D * d.D;
    const int low = (flags & (TE_FLAG_KEEP_ALL_CAMS | TE_FLAG_RELPROP_TO_INPUT)) ? 0 : start_layer;

    TE_TRY(te_launch_argmax(ws.logits, index, d.B, d.C, 1, st));
    TE_TRY(te_launch_onehot(index, ws.seed, d.B, d.C, 1.0f, st));

    // ---- backward: d logit_c / d attention_probs of every layer --------------------------------------------
    float* dxa = ws.tD[0]; float* dsx = ws.tD[1]; float* dctx = ws.tD[2]; float* dxn = ws.tD[3];
    float* dF = ws.tF[0]; float* dqkv = ws.t3D[0]; float* dS = ws.tA[0];
    TE_TRY(linear_bwd(ws.seed, w.clsw, ws.dpool, nullptr, d.B, d.D, d.C, TE_EPI_STORE, st));         // classifier
    TE_TRY(te_launch_tanh_bwd(ws.dpool, ws.pooled, ws.dpd, (long long)d.B * d.D, st));               // pooler tanh
    TE_TRY(linear_bwd(ws.dpd, w.poolw, ws.dfirst, nullptr, d.B, d.D, d.D, TE_EPI_STORE, st));        // pooler dense
    TE_TRY(te_launch_fill(dxa, 0.f, MD, st));
    if (cudaMemcpy2DAsync(dxa, sizeof(float) * d.N * d.D, ws.dfirst, sizeof(float) * d.D, sizeof(float) * d.D, d.B,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        te_set_last_error("te_bert_attribute: scatter of the pooled-token gradient failed");
        return TE_ERR_CUDA;
    }
    for (int l = d.L - 1; l >= start_layer; --l) {
        LayerAct& a = ws.layer[l];
        const LayerW& lw = w.layer[l];
        const HeadOp q = head_rows(a.qkv, 3 * d.D, d.N, d.dh);
        const HeadOp k = head_rows(a.qkv + d.D, 3 * d.D, d.N, d.dh);
        const HeadOp v = head_rows(a.qkv + 2 * d.D, 3 * d.D, d.N, d.dh);
        TE_TRY(te_launch_layernorm_bwd(dxa, a.s2, lw.ln2w, a.mean2, a.rstd2, nullptr, dsx, d.M, d.D, st));    // d s2
        const DerivedW tw = bind_derived(d, lbase, l);
        TE_TRY(linear_bwd_tc(tw.w2, dsx, lw.w2, dF, a.hpre, d.M, d.F, d.D, TE_EPI_GELU_BWD, st, btf, bfs));
        TE_TRY(linear_bwd_tc(tw.w1, dF, lw.w1, dxn, nullptr, d.M, d.D, d.F, TE_EPI_STORE, st, btf, bfs));
        TE_TRY(te_launch_add2(dxn, dsx, dxn, MD, st));                                                          // d ao
        TE_TRY(te_launch_layernorm_bwd(dxn, a.s1, lw.ln1w, a.mean1, a.rstd1, nullptr, dsx, d.M, d.D, st));    // d s1
        TE_TRY(linear_bwd_tc(tw.o, dsx, lw.ow, dctx, nullptr, d.M, d.D, d.D, TE_EPI_STORE, st, btf, bfs));
        TE_TRY(attn_nn(atc, d.B, d.H, d.N, d.NP, d.dh, dctx, d.D, a.qkv + 2 * d.D, 3 * d.D, a.G, nullptr, 1.f,
                       TE_EPI_STORE, st, btf));                                                                      // G = dctx v^T
        if (l == start_layer) break;
        TE_TRY(attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, a.P, 1, dctx, d.D, dqkv + 2 * d.D, 3 * d.D, nullptr, 1.f,
                       TE_EPI_STORE, st, btf));                                                                 // dV = P^T dctx
        TE_TRY(te_launch_softmax_bwd(a.P, a.G, dS, (long long)d.B * d.H * d.N, d.N, d.NP, scale, st));
        TE_TRY(attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, dS, 0, a.qkv + d.D, 3 * d.D, dqkv, 3 * d.D, nullptr, 1.f, TE_EPI_STORE, st, btf));   // dQ
        TE_TRY(attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, dS, 1, a.qkv, 3 * d.D, dqkv + d.D, 3 * d.D, nullptr, 1.f, TE_EPI_STORE, st, btf));   // dK
        TE_TRY(linear_bwd_tc(tw.qkv, dqkv, lw.qkvw, dxn, nullptr, d.M, d.D, 3 * d.D, TE_EPI_STORE, st, btf, bfs));
        TE_TRY(te_launch_add2(dxn, dsx, dxa, MD, st));                                                          // d h
    }

    if (flags & TE_FLAG_GRADIENTS_ONLY) return TE_OK;      // attention-GradCAM baseline: gradients are all it reads

    // ---- relprop -----------------------------------------------------------------------------------------------
    float* R = ws.tD[0]; float* R1 = ws.tD[1]; float* R2 = ws.tD[2]; float* R3 = ws.tD[3];
    float* RF = ws.tF[0]; float* SF = ws.tF[1]; float* S = ws.t3D[0]; float* Rqkv = ws.t3D[1];
    // classifier.relprop (X = pooled) ; dropout / Tanh identity ; pooler.dense.relprop (X = first token) ; pool
    TE_TRY(te_zplus_linear_relprop(ws.pooled, d.D, w.clsw, nullptr, ws.seed, ws.rpool, ws.shead, d.B, d.D, d.C, st));
    TE_TRY(te_zplus_linear_relprop(ws.h_last, (long long)d.N * d.D, w.poolw, nullptr, ws.rpool, ws.rfirst, ws.shead, d.B,
                                   d.D, d.D, st));
    TE_TRY(te_launch_index_select_relprop(ws.h_last, ws.rfirst, nullptr, R, d.B, d.N, d.D, st));

    for (int l = d.L - 1; l >= low; --l) {
        LayerAct& a = ws.layer[l];
        const LayerW& lw = w.layer[l];
        const DerivedW dw = bind_derived(d, dbase, l);
        const HeadOp q = head_rows(a.qkv, 3 * d.D, d.N, d.dh);
        const HeadOp k = head_rows(a.qkv + d.D, 3 * d.D, d.N, d.dh);
        const HeadOp v = head_rows(a.qkv + 2 * d.D, 3 * d.D, d.N, d.dh);
        // BertOutput.relprop :474-487 ; BertIntermediate.relprop :451-456 ; BertLayer.clone
        // top layer: relevance is non-zero only in the first token's row (pooler, BERT.py:181-190) and every rule down
        // to the attention-output dense rule is row-wise -> its three z+ rules run on the B first-token rows only (exact)
        const bool top = (l == d.L - 1) && te_engine_cls_rows();
        const long long zr = top ? d.B : d.M;
        const long long sD = top ? (long long)d.N * d.D : d.D, sF = top ? (long long)d.N * d.F : d.F;
        TE_TRY(te_launch_add_relprop(a.d2, a.ao, R, R1, R2, ws.addpart, d.B, (long long)d.N * d.D, st));
        TE_TRY(te_zplus_linear_relprop_ldr(a.g, sF, lw.w2, dw.w2, R1, sD, RF, S, zr, d.F, d.D, st, a.d2, sD, lw.b2, zb, sF, SF));
        TE_TRY(te_zplus_linear_relprop_ldr(a.ao, sD, lw.w1, dw.w1, RF, sF, R1, SF, zr, d.D, d.F, st, a.hpre, sF, lw.b1, zb, sD, S));
        TE_TRY(te_launch_clone_relprop(a.ao, R1, R2, nullptr, R, MD, st));
        // BertSelfOutput.relprop :427-434
        TE_TRY(te_launch_add_relprop(a.d1, a.h, R, R1, R2, ws.addpart, d.B, (long long)d.N * d.D, st));
        if (top) TE_TRY(te_launch_fill(R3, 0.f, MD, st));
        TE_TRY(te_zplus_linear_relprop_ldr(a.ctx, sD, lw.ow, dw.o, R1, sD, R3, S, zr, d.D, d.D, st, a.d1, sD, lw.ob, zb, sD, S + MD));
        // BertSelfAttention.relprop :367-409
        TE_TRY(te_launch_sd(R3, a.ctx, S, MD, st));                                       // matmul2: Z == saved ctx
        TE_TRY(attn_nn(atc, d.B, d.H, d.N, d.NP, d.dh, S, d.D, a.qkv + 2 * d.D, 3 * d.D, a.cam, a.P, 0.5f, TE_EPI_MUL,
                       st, rtf));                                                              // attn_cam   :380
        if (l == low && !(flags & TE_FLAG_RELPROP_TO_INPUT)) break;
        TE_TRY(attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, a.P, 1, S, d.D, Rqkv + 2 * d.D, 3 * d.D, a.qkv + 2 * d.D, 0.5f, TE_EPI_MUL,
                       st, rtf));
        // add([scores, mask]).relprop : scores = q k^T / sqrt(d) recomputed ; relevance renormalised  :386-388
        TE_TRY(attn_nn(atc, d.B, d.H, d.N, d.NP, d.dh, a.qkv, 3 * d.D, a.qkv + d.D, 3 * d.D, ws.tA[0], nullptr, scale,
                       TE_EPI_STORE, st));
        TE_TRY(te_launch_add_relprop_keymask(ws.tA[0], ws.maskadd, a.cam, ws.tA[1], ws.addpart, d.B, d.H, d.N, d.NP, st));
        // matmul1 rule on the unscaled product
        TE_TRY(attn_nn(atc, d.B, d.H, d.N, d.NP, d.dh, a.qkv, 3 * d.D, a.qkv + d.D, 3 * d.D, ws.tA[0], ws.tA[1], 1.f,
                       TE_EPI_SD, st));
        TE_TRY(attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, ws.tA[0], 0, a.qkv + d.D, 3 * d.D, Rqkv, 3 * d.D, a.qkv, 0.5f, TE_EPI_MUL, st, rtf));
        TE_TRY(attn_nk(atc, d.B, d.H, d.N, d.NP, d.dh, ws.tA[0], 1, a.qkv, 3 * d.D, Rqkv + d.D, 3 * d.D, a.qkv + d.D, 0.5f,
                       TE_EPI_MUL, st, rtf));
        // query / key / value z+ rules (separate Linears), Clone(3), Clone(2)
        TE_TRY(te_zplus_linear_relprop_ldr(a.h, d.D, lw.qkvw, dw.q, Rqkv, 3 * d.D, R, S, d.M, d.D, d.D, st, a.qkv, 3 * d.D, lw.qkvb, zb, 0, S + MD));
        TE_TRY(te_zplus_linear_relprop_ldr(a.h, d.D, lw.qkvw + DD, dw.k, Rqkv + d.D, 3 * d.D, R1, S, d.M, d.D, d.D, st, a.qkv + d.D, 3 * d.D,
                                           lw.qkvb + d.D, zb, 0, S + MD));
        TE_TRY(te_zplus_linear_relprop_ldr(a.h, d.D, lw.qkvw + 2 * DD, dw.v, Rqkv + 2 * d.D, 3 * d.D, R3, S, d.M, d.D, d.D, st,
                                           a.qkv + 2 * d.D, 3 * d.D, lw.qkvb + 2 * d.D, zb, 0, S + MD));
        TE_TRY(te_launch_clone_relprop(a.h, R, R1, R3, SF, MD, st));                      // self.clone (3-way)
        TE_TRY(te_launch_clone_relprop(a.h, SF, R2, nullptr, R, MD, st));                 // attention.clone
    }

    // ---- aggregation + normalised rollout, row 0 with [0] = min   (ExplanationGenerator.py:47-59) ---------------
    TE_TRY(te_rollout_layers(ws.layer[0].G, ws.layer[0].cam, d.L > 1 ? (long long)(ws.layer[1].G - ws.layer[0].G) : 0,
                             d.L, d.B, d.H, d.N, d.NP, d.NP, start_layer, /*normalize=*/1, flags, ws.mats, ws.joint[0],
                             ws.joint[1], nullptr, maps, /*first=*/0, /*bert_fix=*/1, st));
    return TE_OK;
}

extern "C" int te_bert_explain(const te_bert_config* cfg, const float* weights, const float* derived,
                               const long long* input_ids, const long long* attention_mask, int batch, int seq,
                               int* index, int start_layer, unsigned flags, float* maps, float* logits, void* workspace,
                               long long workspace_bytes, void* stream) {
    TE_TRY(te_bert_forward(cfg, weights, derived, input_ids, attention_mask, batch, seq, flags, logits, workspace,
                           workspace_bytes, stream));
    return te_bert_attribute(cfg, weights, derived, batch, seq, index, start_layer, flags, maps, workspace, workspace_bytes,
                             stream);
}

extern "C" int te_bert_tensor(const te_bert_config* cfg, int batch, int seq, void* workspace, const char* name, int layer,
                              float** ptr, long long dims[4], long long strides[4]) {
    Dims d; Workspace ws;
    if (!workspace || !name || !ptr) return TE_ERR_ARG;
    if (batch <= 0 || !make_dims(cfg, batch, seq, d)) return TE_ERR_ARG;
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
    if (layer < 0 || layer >= d.L) { te_set_last_error("te_bert_tensor: layer out of range"); return TE_ERR_ARG; }
    LayerAct& a = ws.layer[layer];
    const long long hs = (long long)d.N * d.NP, bs = hs * d.H;
    if (n == "attn") return set(a.P, d.B, d.H, d.N, d.N, bs, hs, d.NP, 1);
    if (n == "attn_grad") return set(a.G, d.B, d.H, d.N, d.N, bs, hs, d.NP, 1);
    if (n == "attn_cam") return set(a.cam, d.B, d.H, d.N, d.N, bs, hs, d.NP, 1);
    if (n == "hidden") return set(a.h, d.B, d.N, d.D, 1, (long long)d.N * d.D, d.D, 1, 1);
    te_set_last_error("te_bert_tensor: unknown tensor name");
    return TE_ERR_ARG;
}
