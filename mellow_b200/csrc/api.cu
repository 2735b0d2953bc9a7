// C ABI + orchestration of the Mellow inference path (see include/mellow_b200.h for the reference call sites each
// entry point replaces).  Host code here only sequences kernel launches; all arithmetic is in the .cu kernels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mellow_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"

namespace mb {

namespace { thread_local int g_pdl_off = 0; }
bool pdl_enabled() {
    static const bool on = getenv("MB_NO_PDL") == nullptr;
    return on && g_pdl_off == 0;
}
// Scope in which kernels are launched with a full (non-programmatic) dependency: the first kernel after a stream
// fork / join of the row-group decode step, whose predecessors are event edges rather than one kernel.
struct PdlOff {
    bool on;
    explicit PdlOff(bool enable = true) : on(enable) { if (on) ++g_pdl_off; }
    ~PdlOff() { if (on) --g_pdl_off; }
};

cudaError_t launch_gemm_umma(const GemmArgs& g, int epi, cudaStream_t st, bool* handled);   // gemm_umma.cu

namespace {

constexpr int kMaxSplitK = 12;
constexpr int kMaxGroups = 4;
constexpr int kGuSplitMax = 3;
constexpr int kDepths[4] = {2, 2, 6, 2};
constexpr int kHeadsPerStage[4] = {4, 8, 16, 32};
inline int stage_dim(int i) { return kEmbed << i; }
inline int stage_res(int i) { return kGrid0 >> i; }

struct Plane { const bf16* hi; const bf16* lo; };

struct SwinBlockW {
    const float *ln1_w, *ln1_b, *qkv_b, *relbias, *proj_b, *ln2_w, *ln2_b, *fc1_b, *fc2_b;
    Plane qkv, proj, fc1, fc2;
};
struct MergeW { const float *ln_w, *ln_b; Plane red; };
struct LmLayerW { const float *ln1, *ln2; Plane qkv, o, gu, down; };

struct Weights {
    // front end
    const float *window, *twiddle, *melW, *bn_scale, *bn_shift;
    const int *mel_lo, *mel_hi;
    const float *pe_w, *pe_b, *pe_ln_w, *pe_ln_b;
    SwinBlockW blk[4][6];
    MergeW merge[3];
    const float *tail_ln_w, *tail_ln_b, *tscam_b, *c2l_b, *proj_ln_w, *proj_ln_b;
    Plane tscam, c2l, proj1, proj2;
    // LM
    const float *embed, *lm_norm, *rope_cos, *rope_sin;
    Plane head;
    LmLayerW layer[kLayers];
};

struct Entry { std::string name; size_t bytes; size_t offset; const void** slot; };

struct Table {
    std::vector<Entry> entries;
    size_t total = 0;
    void add(const std::string& name, size_t bytes, const void** slot) {
        entries.push_back({name, bytes, total, slot});
        total += (bytes + 255) / 256 * 256;
    }
    void f32(const std::string& n, size_t count, const float** slot) { add(n, count * 4, (const void**)slot); }
    void i32(const std::string& n, size_t count, const int** slot) { add(n, count * 4, (const void**)slot); }
    void plane(const std::string& n, size_t count, Plane* p) {
        add(n + ".hi", count * 2, (const void**)&p->hi);
        add(n + ".lo", count * 2, (const void**)&p->lo);
    }
};

// The arena layout: single source of truth for the host packer (mellow_b200/weights.py asks by name).
void build_table(Table& t, Weights& w) {
    t.f32("fe.window", kNfft, &w.window);
    t.f32("fe.twiddle", 1024, &w.twiddle);
    t.f32("fe.melW", (size_t)kBins * kMels, &w.melW);
    t.i32("fe.mel_lo", kMels, &w.mel_lo);
    t.i32("fe.mel_hi", kMels, &w.mel_hi);
    t.f32("fe.bn_scale", kMels, &w.bn_scale);
    t.f32("fe.bn_shift", kMels, &w.bn_shift);
    t.f32("pe.w", kEmbed * 16, &w.pe_w);
    t.f32("pe.b", kEmbed, &w.pe_b);
    t.f32("pe.ln_w", kEmbed, &w.pe_ln_w);
    t.f32("pe.ln_b", kEmbed, &w.pe_ln_b);
    for (int i = 0; i < 4; ++i) {
        const size_t C = stage_dim(i);
        for (int b = 0; b < kDepths[i]; ++b) {
            SwinBlockW& k = w.blk[i][b];
            const std::string p = "s" + std::to_string(i) + ".b" + std::to_string(b) + ".";
            t.f32(p + "ln1_w", C, &k.ln1_w);
            t.f32(p + "ln1_b", C, &k.ln1_b);
            t.plane(p + "qkv_w", 3 * C * C, &k.qkv);
            t.f32(p + "qkv_b", 3 * C, &k.qkv_b);
            t.f32(p + "relbias", (size_t)kHeadsPerStage[i] * 64 * 64, &k.relbias);
            t.plane(p + "proj_w", C * C, &k.proj);
            t.f32(p + "proj_b", C, &k.proj_b);
            t.f32(p + "ln2_w", C, &k.ln2_w);
            t.f32(p + "ln2_b", C, &k.ln2_b);
            t.plane(p + "fc1_w", 4 * C * C, &k.fc1);
            t.f32(p + "fc1_b", 4 * C, &k.fc1_b);
            t.plane(p + "fc2_w", 4 * C * C, &k.fc2);
            t.f32(p + "fc2_b", C, &k.fc2_b);
        }
        if (i < 3) {
            const std::string p = "m" + std::to_string(i) + ".";
            t.f32(p + "ln_w", 4 * C, &w.merge[i].ln_w);
            t.f32(p + "ln_b", 4 * C, &w.merge[i].ln_b);
            t.plane(p + "red_w", 8 * C * C, &w.merge[i].red);
        }
    }
    t.f32("tail.ln_w", kEncOut, &w.tail_ln_w);
    t.f32("tail.ln_b", kEncOut, &w.tail_ln_b);
    t.plane("tail.tscam_w", (size_t)kClasses * 6 * kEncOut, &w.tscam);
    t.f32("tail.tscam_b", kClasses, &w.tscam_b);
    t.plane("tail.c2l_w", (size_t)kEncOut * kClassesPad, &w.c2l);
    t.f32("tail.c2l_b", kEncOut, &w.c2l_b);
    t.plane("proj.w1", (size_t)kProj * kEncOut, &w.proj1);
    t.plane("proj.w2", (size_t)kProj * kProj, &w.proj2);
    t.f32("proj.ln_w", kProj, &w.proj_ln_w);
    t.f32("proj.ln_b", kProj, &w.proj_ln_b);
    t.f32("lm.embed", (size_t)kVocab * kHidden, &w.embed);
    t.plane("lm.head_w", (size_t)kVocab * kHidden, &w.head);
    t.f32("lm.norm", kHidden, &w.lm_norm);
    t.f32("lm.rope_cos", (size_t)kMaxPos * 32, &w.rope_cos);
    t.f32("lm.rope_sin", (size_t)kMaxPos * 32, &w.rope_sin);
    for (int l = 0; l < kLayers; ++l) {
        LmLayerW& k = w.layer[l];
        const std::string p = "lm.l" + std::to_string(l) + ".";
        t.f32(p + "ln1", kHidden, &k.ln1);
        t.plane(p + "qkv_w", (size_t)kQkvDim * kHidden, &k.qkv);
        t.plane(p + "o_w", (size_t)kHidden * kHidden, &k.o);
        t.f32(p + "ln2", kHidden, &k.ln2);
        t.plane(p + "gu_w", (size_t)2 * kInter * kHidden, &k.gu);
        t.plane(p + "down_w", (size_t)kHidden * kInter, &k.down);
    }
}

struct Layout {
    Weights w;
    Table t;
    Layout() { build_table(t, w); }
};
Layout& layout() {
    static Layout L;
    return L;
}

thread_local char g_create_error[512] = "";

struct Handle {
    int device = 0, max_batch = 0, max_new = 0, policy = 0, engine = 1;   // engine 1: tcgen05 where a tile fits
    int t_max = 0;
    bool bound = false;
    Weights w;
    Table t;
    char err[512] = "";
    long long launches = 0;
    size_t ws_bytes = 0;
    std::vector<void*> allocs;
    // encoder workspace (N = 2*max_batch clips)
    float *bn = nullptr, *xa = nullptr, *xb = nullptr, *qkv = nullptr;
    bf16 *a_hi = nullptr, *a_lo = nullptr, *h_hi = nullptr, *h_lo = nullptr;
    float *ytail = nullptr, *latent = nullptr, *frames = nullptr, *e1 = nullptr, *s33 = nullptr, *rows33 = nullptr;
    bf16 *col_hi = nullptr, *col_lo = nullptr, *cls_hi = nullptr, *cls_lo = nullptr;
    bf16 *a33_hi = nullptr, *a33_lo = nullptr, *g33_hi = nullptr, *g33_lo = nullptr;
    // LM workspace (M = max_batch*389 rows)
    float *x = nullptr, *q = nullptr, *logits = nullptr;
    bf16 *la_hi = nullptr, *la_lo = nullptr, *lh_hi = nullptr, *lh_lo = nullptr;
    void *kcache = nullptr, *vcache = nullptr;
    size_t kv_layer_elems = 0;
    float *part_acc = nullptr, *part_ml = nullptr, *gemm_partial = nullptr, *cand_val = nullptr, *qkv_part = nullptr, *gu_part = nullptr;
    float* rope_cur = nullptr;           // cos | sin row of the position the current decode step writes (step_advance_kernel)
    unsigned* fix_counter = nullptr;     // split-K fix-up tickets of the gate/up tiles (self-resetting)
    int* cand_idx = nullptr;
    bool logits_fused = false;           // the last lm_head wrote argmax candidates instead of logits
    unsigned* chain_bar = nullptr;       // [kLayers][8] grid-barrier counters of the fused decode chain
    int *d_tokens = nullptr, *d_done = nullptr, *d_step = nullptr, *d_stop = nullptr, *d_ids = nullptr;
    float *wave_stage = nullptr;
    int prefix_B = 0;
    int enc_clips = 0;                   // clips of the last full encoder pass (its tail buffers feed mb_encode_heads)
    // decode graph cache
    cudaGraphExec_t graph = nullptr;
    int g_B = 0, g_max_len = 0, g_eos = 0, g_launches = 0;
    float g_temp = 0.f;
    cudaStream_t g_stream = nullptr;
    cudaStream_t own_stream = nullptr;   // used when the caller passes NULL (the legacy stream cannot be graph-captured)
    // row-group decode: groups 1.. run on their own streams, forked from / joined to the caller's stream by events
    cudaStream_t grp_stream[kMaxGroups - 1] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxGroups - 1] = {}, ev_attn[kMaxGroups] = {};
    int kv_fmt = kKvF32;                 // KV-cache row format (common.cuh)
    int qkv_split = -1;                  // mb_set_decode_qkv_split: -1 = default (MB_DEC_QKV_SPLIT / 0)
    int groups = 0;                      // mb_set_decode_groups: 0 = automatic
    TraceBuf* trace = nullptr;           // mb_set_trace: optional in-kernel timeline of the decode kernels
};

inline cudaStream_t pick_stream(Handle* h, void* stream) { return stream ? (cudaStream_t)stream : h->own_stream; }

#define MB_CK(h, expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            snprintf((h)->err, sizeof((h)->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                                                \
            return 1;                                                                                    \
        }                                                                                                \
    } while (0)

#define MB_TRY(expr)            \
    do {                        \
        int r__ = (expr);       \
        if (r__ != 0) return r__; \
    } while (0)

int fail(Handle* h, const char* msg) {
    snprintf(h->err, sizeof(h->err), "%s", msg);
    return 1;
}

template <typename T>
int dev_alloc(Handle* h, T** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
    MB_CK(h, cudaMalloc(&q, bytes));
    h->allocs.push_back(q);
    h->ws_bytes += bytes;
    *p = reinterpret_cast<T*>(q);
    return 0;
}

inline const bf16* lo_of(const Handle* h, const bf16* lo) { return h->policy == kPolicySplit ? lo : nullptr; }
inline bf16* lo_of(const Handle* h, bf16* lo) { return h->policy == kPolicySplit ? lo : nullptr; }

GemmArgs gemm_base(const Handle* h, const bf16* a_hi, const bf16* a_lo, int lda, Plane wgt, int ldw, int M, int N, int K) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A_hi = a_hi; g.A_lo = lo_of(h, a_lo); g.lda = lda;
    g.W_hi = wgt.hi; g.W_lo = lo_of(h, wgt.lo); g.ldw = ldw;
    g.M = M; g.N = N; g.K = K;
    g.passes = h->policy == kPolicySplit ? 3 : 1;
    static const int epi_sleep = getenv("MB_EPI_SLEEP") ? atoi(getenv("MB_EPI_SLEEP")) : 128;
    g.epi_sleep = epi_sleep;
    return g;
}

int run_gemm(Handle* h, const GemmArgs& g, int epi, cudaStream_t st) {
    if (h->engine == 1 && g.resident) {
        const int bn = g.bn_hint ? g.bn_hint : (g.N >= 2048 ? 32 : 16);
        cudaError_t e = launch_gemm_skinny(g, epi, bn, st);
        if (e == cudaSuccess) { h->launches++; return 0; }
        if (e != cudaErrorNotSupported) MB_CK(h, e);
        (void)cudaGetLastError();
    }
    if (h->engine == 1) {
        bool handled = false;
        MB_CK(h, launch_gemm_umma(g, epi, st, &handled));
        if (handled) { h->launches++; return 0; }
    }
    if (epi == EPI_ARGMAX) return fail(h, "argmax epilogue needs the tcgen05 engine");
    MB_CK(h, launch_gemm_mma(g, epi, st));
    h->launches++;
    return 0;
}

int run_norm(Handle* h, int kind, const float* x, const float* w, const float* b, int rows, int C, bf16* hi, bf16* lo,
             float* out_f32, int row_stride, int row_off, int res, cudaStream_t st) {
    NormArgs a;
    a.x = x; a.w = w; a.b = b;
    a.out_hi = hi; a.out_lo = hi ? lo_of(h, lo) : nullptr; a.out_f32 = out_f32;
    a.rows = rows; a.C = C; a.row_stride = row_stride; a.row_off = row_off; a.res = res;
    MB_CK(h, launch_norm(a, kind, st));
    h->launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------------ encoder
int swin_block(Handle* h, float* x, int n_clips, int stage, int b, cudaStream_t st) {
    const int C = stage_dim(stage), R = stage_res(stage), nH = kHeadsPerStage[stage];
    const int M = n_clips * R * R;
    const int shift = (b % 2 == 1 && R > kWin) ? kWin / 2 : 0;      // htsat.py:539, 368-371
    const SwinBlockW& k = h->w.blk[stage][b];
    MB_TRY(run_norm(h, NORM_LN, x, k.ln1_w, k.ln1_b, M, C, h->a_hi, h->a_lo, nullptr, 1, 0, 0, st));
    {
        GemmArgs g = gemm_base(h, h->a_hi, h->a_lo, C, k.qkv, C, M, 3 * C, C);
        g.bias = k.qkv_b; g.out_f32 = h->qkv; g.ldo = 3 * C;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    static const bool fp32_win = getenv("MB_ATTN_FP32") != nullptr;            // CUDA-core fp32 kernel, kept for A/B checks
    if (fp32_win)
        MB_CK(h, launch_window_attention(h->qkv, k.relbias, h->a_hi, lo_of(h, h->a_lo), n_clips, R, C, nH, shift, st));
    else
        MB_CK(h, launch_window_attention_mma(h->qkv, k.relbias, h->a_hi, lo_of(h, h->a_lo), n_clips, R, C, nH, shift, st));
    h->launches++;
    {
        GemmArgs g = gemm_base(h, h->a_hi, h->a_lo, C, k.proj, C, M, C, C);
        g.bias = k.proj_b; g.residual = x; g.ldr = C; g.out_f32 = x; g.ldo = C;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    MB_TRY(run_norm(h, NORM_LN, x, k.ln2_w, k.ln2_b, M, C, h->a_hi, h->a_lo, nullptr, 1, 0, 0, st));
    {
        GemmArgs g = gemm_base(h, h->a_hi, h->a_lo, C, k.fc1, C, M, 4 * C, C);
        g.bias = k.fc1_b; g.act = ACT_GELU; g.out_hi = h->h_hi; g.out_lo = lo_of(h, h->h_lo); g.ldp = 4 * C;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    {
        GemmArgs g = gemm_base(h, h->h_hi, h->h_lo, 4 * C, k.fc2, 4 * C, M, C, 4 * C);
        g.bias = k.fc2_b; g.residual = x; g.ldr = C; g.out_f32 = x; g.ldo = C;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    return 0;
}

// bn [n,1001,64] -> rows33 (or a tap).  Returns the buffer holding the residual stream in *x_final.
int encoder_body(Handle* h, int n_clips, int stop_stage, float* tap_out, float* rows_out, cudaStream_t st) {
    PatchW pw{h->w.pe_w, h->w.pe_b, h->w.pe_ln_w, h->w.pe_ln_b};
    MB_CK(h, launch_patch_embed(h->bn, n_clips, pw, h->xa, st));
    h->launches++;
    float* x = h->xa;
    float* other = h->xb;
    auto tap = [&](int stage_id, const float* src, size_t count) -> int {
        if (tap_out && stop_stage == stage_id) {
            MB_CK(h, cudaMemcpyAsync(tap_out, src, count * sizeof(float), cudaMemcpyDeviceToDevice, st));
            return 2;
        }
        return 0;
    };
    int r = tap(0, x, (size_t)n_clips * 4096 * kEmbed);
    if (r) return r == 2 ? 0 : r;
    for (int i = 0; i < 4; ++i) {
        const int C = stage_dim(i), R = stage_res(i);
        for (int b = 0; b < kDepths[i]; ++b) MB_TRY(swin_block(h, x, n_clips, i, b, st));
        size_t count = (size_t)n_clips * R * R * C;
        if (i < 3) {
            const int M4 = n_clips * (R / 2) * (R / 2);
            MB_TRY(run_norm(h, NORM_LN_MERGE, x, h->w.merge[i].ln_w, h->w.merge[i].ln_b, M4, 4 * C, h->a_hi, h->a_lo,
                            nullptr, 1, 0, R, st));
            GemmArgs g = gemm_base(h, h->a_hi, h->a_lo, 4 * C, h->w.merge[i].red, 4 * C, M4, 2 * C, 4 * C);
            g.out_f32 = other; g.ldo = 2 * C;
            MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
            float* tmp = x; x = other; other = tmp;
            count = (size_t)M4 * 2 * C;
        }
        r = tap(i + 1, x, count);
        if (r) return r == 2 ? 0 : r;
    }
    // tail: final LN, latent mean, TSCAM conv as GEMM (+sigmoid), c2l, projection -- on the 33 unique rows per clip
    const int Mt = n_clips * 64;
    MB_TRY(run_norm(h, NORM_LN, x, h->w.tail_ln_w, h->w.tail_ln_b, Mt, kEncOut, nullptr, nullptr, h->ytail, 1, 0, 0, st));
    MB_CK(h, launch_tail_gather(h->ytail, n_clips, h->latent, h->col_hi, lo_of(h, h->col_lo), st));
    h->launches++;
    {
        GemmArgs g = gemm_base(h, h->col_hi, h->col_lo, 6 * kEncOut, h->w.tscam, 6 * kEncOut, n_clips * 32, kClasses, 6 * kEncOut);
        g.bias = h->w.tscam_b; g.act = ACT_SIGMOID; g.out_hi = h->cls_hi; g.out_lo = lo_of(h, h->cls_lo); g.ldp = kClassesPad;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    {
        GemmArgs g = gemm_base(h, h->cls_hi, h->cls_lo, kClassesPad, h->w.c2l, kClassesPad, n_clips * 32, kEncOut, kClassesPad);
        g.bias = h->w.c2l_b; g.out_f32 = h->frames; g.ldo = kEncOut;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    if (tap_out && stop_stage == 5) {
        MB_CK(h, cudaMemcpyAsync(tap_out, h->latent, (size_t)n_clips * kEncOut * 4, cudaMemcpyDeviceToDevice, st));
        MB_CK(h, cudaMemcpyAsync(tap_out + (size_t)n_clips * kEncOut, h->frames, (size_t)n_clips * 32 * kEncOut * 4,
                                 cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    MB_CK(h, launch_assemble33(h->latent, h->frames, n_clips, h->a33_hi, lo_of(h, h->a33_lo), st));
    h->launches++;
    h->enc_clips = n_clips;
    const int M33 = n_clips * kAudioRows;
    {
        GemmArgs g = gemm_base(h, h->a33_hi, h->a33_lo, kEncOut, h->w.proj1, kEncOut, M33, kProj, kEncOut);
        g.out_f32 = h->e1; g.ldo = kProj;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    MB_CK(h, launch_gelu_planes(h->e1, (size_t)M33 * kProj, h->g33_hi, lo_of(h, h->g33_lo), st));
    h->launches++;
    {
        GemmArgs g = gemm_base(h, h->g33_hi, h->g33_lo, kProj, h->w.proj2, kProj, M33, kProj, kProj);
        g.residual = h->e1; g.ldr = kProj; g.out_f32 = h->s33; g.ldo = kProj;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    MB_TRY(run_norm(h, NORM_LN, h->s33, h->w.proj_ln_w, h->w.proj_ln_b, M33, kProj, nullptr, nullptr, h->rows33, 1, 0, 0, st));
    if (rows_out)
        MB_CK(h, cudaMemcpyAsync(rows_out, h->rows33, (size_t)M33 * kProj * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int frontend(Handle* h, const float* wave, int n_clips, float* logmel_out, float* bn_out, cudaStream_t st) {
    FrontendW fw{h->w.window, h->w.twiddle, h->w.melW, h->w.mel_lo, h->w.mel_hi, h->w.bn_scale, h->w.bn_shift};
    MB_CK(h, launch_logmel(wave, n_clips, fw, logmel_out, bn_out, st));
    h->launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------------ LM
inline void* kv_layer(const Handle* h, void* base, int l) {
    return reinterpret_cast<char*>(base) + (size_t)l * (h->kv_layer_elems / kHeadDim) * kv_row_bytes(h->kv_fmt);
}

constexpr int kMaxAttnSplit = 16;
inline int decode_nsplit(const Handle* h, int B) {
    // One CTA per (row, kv head) streams the whole context through a 2-tile cp.async ring and writes the final
    // output itself when that already gives >= ~2.5 CTAs per SM (B >= 128); smaller batches split the keys so the
    // 148 SMs stay busy and merge the partial softmax states in decode_combine_kernel.
    const int tiles_max = (h->t_max + 63) / 64;
    int ns = (384 + 3 * B - 1) / (3 * B);
    if (ns > tiles_max) ns = tiles_max;
    return ns < 1 ? 1 : (ns > kMaxAttnSplit ? kMaxAttnSplit : ns);
}

// Tiling of the split-K decode GEMMs (gemm_umma.cu launch_epi): N-tile width x K split.  The default follows the
// measured rule "a decode GEMM costs what its CTAs' tcgen05.mma COUNT costs": wide tiles, 1-2 k-blocks per CTA.
// MB_DEC_TILING=0 restores the 16-column tiles with split 3 / 4.
struct DecodeTiling { int o_bn, o_split, down_bn, down_split, resident, qkv_split, gu_split; };
const DecodeTiling& decode_tiling() {
    static const DecodeTiling t = [] {
        const char* e = getenv("MB_DEC_TILING");
        const int mode = e ? atoi(e) : 6;
        const char* r = getenv("MB_DEC_RESIDENT");
        const int res = r ? atoi(r) : 1;
        // MB_DEC_QKV_SPLIT=9: QKV as 9 one-k-block slices of 64-column tiles (135 CTAs x 8 MMAs instead of 60 CTAs x 72); the
        // partial sums are reduced, roped and appended to the KV cache by the decode-attention kernel that consumes them.
        // It paid off while the MMA issue loop was slow (1.59 -> 1.55 ms/step); with the elect.sync loop the unsplit GEMM
        // with its in-epilogue RoPE / KV write is ahead again (1.356 vs 1.383 ms/step) and attention has no prologue
        const char* q = getenv("MB_DEC_QKV_SPLIT");
        const int qs = res ? (q ? atoi(q) : 0) : 0;       // off by default: see below
        // gate/up as 3 K slices of 64-column tiles (144 CTAs x 24 MMAs instead of 96 x 72); SwiGLU needs the complete
        // sums, so the split that arrives last at its tile finishes it inside the same kernel (gemm_skinny.cu fix-up)
        const char* gq = getenv("MB_DEC_GU_SPLIT");
        const int gs = res ? (gq ? atoi(gq) : 0) : 0;   // measured slower (fence + ticket + re-read cost more than the MMAs saved): off
        if (mode == 1) return DecodeTiling{48, 9, 48, 12, res, qs, gs};
        if (mode == 2) return DecodeTiling{64, 9, 64, 12, res, qs, gs};
        if (mode == 3) return DecodeTiling{48, 9, 48, 6, res, qs, gs};
        if (mode == 4) return DecodeTiling{48, 3, 48, 6, res, qs, gs};
        if (mode == 5) return DecodeTiling{0, 3, 32, 8, res, qs, gs};
        if (mode == 6) return DecodeTiling{32, 3, 32, 8, res, qs, gs};
        if (mode == 7) return DecodeTiling{0, 3, 48, 12, res, qs, gs};
        if (mode == 8) return DecodeTiling{32, 3, 0, 4, res, qs, gs};
        if (mode == 0) return DecodeTiling{0, 3, 0, 4, res, qs, gs};
        return DecodeTiling{32, 3, 32, 8, res, qs, gs};
    }();
    return t;
}

// Row groups of the decode step (see decode_step).  MB_DECODE_GROUPS / MB_DECODE_COMPACT override the defaults for
// A/B measurements.
int decode_groups(const Handle* h, int B) {
    static const int forced = getenv("MB_DECODE_GROUPS") ? atoi(getenv("MB_DECODE_GROUPS")) : 0;
    int G = h->groups > 0 ? h->groups : (forced > 0 ? forced : 1);
    if (G > kMaxGroups) G = kMaxGroups;
    if (G > B) G = B;
    if (h->engine != 1) G = 1;
    return G < 1 ? 1 : G;
}
int decode_stagger(int G) {
    static const int forced = getenv("MB_DECODE_STAGGER") ? atoi(getenv("MB_DECODE_STAGGER")) : -1;
    return G > 1 ? (forced >= 0 ? forced : 0) : 0;
}
int decode_compact(int G) {
    static const int forced = getenv("MB_DECODE_COMPACT") ? atoi(getenv("MB_DECODE_COMPACT")) : -1;
    return forced >= 0 ? forced : 1;
}

// Rows [r0, r0+n) of the batch; n_all = rows of the whole step (all row groups), which sets the key split.
int run_decode_attention(Handle* h, int l, int r0, int n, int n_all, cudaStream_t st, bool skip_done = true,
                         const float* qkv_part = nullptr, int qkv_nsplit = 0) {
    DecodeAttnArgs a;
    a.qkv_part = qkv_part; a.qkv_nsplit = qkv_nsplit; a.rope_cur = h->rope_cur;
    const size_t kv_off = (size_t)r0 * kKvHeads * h->t_max * kv_row_bytes(h->kv_fmt);
    a.q = h->q + (size_t)r0 * kHidden;
    a.kc = reinterpret_cast<char*>(kv_layer(h, h->kcache, l)) + kv_off;
    a.vc = reinterpret_cast<char*>(kv_layer(h, h->vcache, l)) + kv_off;
    a.kv_fmt = h->kv_fmt; a.B = n; a.t_max = h->t_max;
    a.nsplit = decode_nsplit(h, n_all);
    a.tps = ((h->t_max + 63) / 64 + a.nsplit - 1) / a.nsplit;
    a.ctx_base = kPrefix; a.d_step = h->d_step;
    a.done = skip_done ? h->d_done + r0 : nullptr;
    a.part_acc = h->part_acc + (size_t)r0 * kHeads * a.nsplit * kHeadDim;
    a.part_ml = h->part_ml + (size_t)r0 * kHeads * a.nsplit * 2;
    a.out_hi = h->la_hi + (size_t)r0 * kHidden; a.out_lo = lo_of(h, h->la_lo + (size_t)r0 * kHidden);
    a.trace = h->trace; a.trace_id = 2000 + (r0 ? 100 : 0) + l;
    MB_CK(h, launch_decode_attention(a, st));
    h->launches += a.nsplit == 1 ? 1 : 2;
    return 0;
}
int run_decode_attention(Handle* h, int l, int B, cudaStream_t st, bool skip_done = true) {
    return run_decode_attention(h, l, 0, B, B, st, skip_done);
}

// Decode layer for one row group [r0, r0+n), n <= 128 rows (one M tile).  The residual stream update and the next
// RMSNorm are fused into add_rmsnorm_kernel, which also reduces the split-K partial sums of o_proj / down_proj in a
// fixed order.  On entry la_hi/la_lo hold RMSNorm(x) with this layer's input_layernorm; on exit they hold RMSNorm(x)
// with `next_norm` (the next layer's input_layernorm, or the final model norm).  Every row is independent of every
// other row, so the values are identical however the batch is cut into groups.
int lm_layer_decode_fused(Handle* h, int l, int gi, int r0, int n, int n_all, int compact,
                          const float* next_norm, cudaStream_t st, cudaEvent_t attn_wait = nullptr,
                          cudaEvent_t attn_record = nullptr) {
    float* partial = h->gemm_partial + (size_t)gi * kMaxSplitK * 128 * kHidden;           // split-K partials of this row group
    float* qkv_part = h->qkv_part + (size_t)gi * kQkvSplitMax * 128 * kQkvDim;
    const LmLayerW& k = h->w.layer[l];
    // Policy fast keeps the 16-column tiles and the in-GEMM QKV epilogue: with the 32-column split-K tiles its first
    // decode on a fresh handle produced non-finite logits on 3 of ~12 boxes (B=2; never with policy split, never
    // under compute-sanitizer memcheck / initcheck / racecheck) -- unexplained, see DESIGN.md section 10.
    static const DecodeTiling conservative{0, 3, 0, 4, decode_tiling().resident, 0, 0};
    const DecodeTiling& tl = h->policy == kPolicyFast ? conservative : decode_tiling();
    const size_t kv_off = (size_t)r0 * kKvHeads * h->t_max * kv_row_bytes(h->kv_fmt);
    float* x = h->x + (size_t)r0 * kHidden;
    bf16* la_hi = h->la_hi + (size_t)r0 * kHidden; bf16* la_lo = h->la_lo + (size_t)r0 * kHidden;
    bf16* lh_hi = h->lh_hi + (size_t)r0 * kInter; bf16* lh_lo = h->lh_lo + (size_t)r0 * kInter;
    const int qkv_nsplit = h->policy == kPolicyFast ? 0 : (h->qkv_split >= 0 ? h->qkv_split : tl.qkv_split);
    const bool qkv_split = qkv_nsplit > 1 && h->engine == 1 && tl.resident;
    if (qkv_split) {
        GemmArgs g = gemm_base(h, la_hi, la_lo, kHidden, k.qkv, kHidden, n, kQkvDim, kHidden);
        g.resident = 1; g.bn_hint = 64;
        g.trace = h->trace; g.trace_id = 1000 + (r0 ? 100 : 0) + l;
        g.split_k = qkv_nsplit; g.partial = qkv_part;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    } else {
        GemmArgs g = gemm_base(h, la_hi, la_lo, kHidden, k.qkv, kHidden, n, kQkvDim, kHidden);
        g.compact = compact; g.resident = tl.resident;
        g.trace = h->trace; g.trace_id = 1000 + (r0 ? 100 : 0) + l;
        g.q_out = h->q + (size_t)r0 * kHidden;
        g.k_cache = reinterpret_cast<char*>(kv_layer(h, h->kcache, l)) + kv_off;
        g.v_cache = reinterpret_cast<char*>(kv_layer(h, h->vcache, l)) + kv_off;
        g.rope_cos = h->w.rope_cos; g.rope_sin = h->w.rope_sin;
        g.rows_per_seq = 1; g.pos_base = kPrefix - 1; g.d_pos = h->d_step;
        g.t_max = h->t_max; g.kv_fmt = h->kv_fmt;
        MB_TRY(run_gemm(h, g, EPI_QKV_ROPE, st));
    }
    if (attn_wait) MB_CK(h, cudaStreamWaitEvent(st, attn_wait, 0));
    {
        PdlOff off(attn_wait != nullptr);                  // two predecessors: the QKV kernel and the other group's event
        MB_TRY(run_decode_attention(h, l, r0, n, n_all, st, true, qkv_split ? qkv_part : nullptr, qkv_nsplit));
    }
    if (attn_record) MB_CK(h, cudaEventRecord(attn_record, st));
    {
        GemmArgs g = gemm_base(h, la_hi, la_lo, kHidden, k.o, kHidden, n, kHidden, kHidden);
        g.compact = compact; g.resident = tl.resident;
        g.trace = h->trace; g.trace_id = 3000 + (r0 ? 100 : 0) + l;
        g.split_k = tl.o_split; g.bn_hint = tl.o_bn; g.partial = partial;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    MB_CK(h, launch_add_rmsnorm(x, partial, tl.o_split, n, k.ln2, la_hi, lo_of(h, la_lo), st, h->trace, 4000 + (r0 ? 100 : 0) + l));
    h->launches++;
    {
        GemmArgs g = gemm_base(h, la_hi, la_lo, kHidden, k.gu, kHidden, n, 2 * kInter, kHidden);
        g.compact = compact; g.resident = tl.resident;
        g.trace = h->trace; g.trace_id = 5000 + (r0 ? 100 : 0) + l;
        if (tl.gu_split > 1 && h->engine == 1) {
            g.split_k = tl.gu_split; g.bn_hint = 64;
            g.partial = h->gu_part + (size_t)gi * kGuSplitMax * 128 * 2 * kInter;
            g.fix_counter = h->fix_counter + (size_t)gi * 64;
        }
        g.out_hi = lh_hi; g.out_lo = lo_of(h, lh_lo); g.ldp = kInter;
        MB_TRY(run_gemm(h, g, EPI_SWIGLU, st));
    }
    {
        GemmArgs g = gemm_base(h, lh_hi, lh_lo, kInter, k.down, kInter, n, kHidden, kInter);
        g.compact = compact; g.resident = tl.resident;
        g.trace = h->trace; g.trace_id = 6000 + (r0 ? 100 : 0) + l;
        g.split_k = tl.down_split; g.bn_hint = tl.down_bn; g.partial = partial;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    MB_CK(h, launch_add_rmsnorm(x, partial, tl.down_split, n, next_norm, la_hi, lo_of(h, la_lo), st, h->trace, 7000 + (r0 ? 100 : 0) + l));
    h->launches++;
    return 0;
}

// Fused chain after the attention of layer l (decode_chain.cu): o_proj, +norm, gate/up, down, +norm, next layer's QKV.
int lm_layer_decode_chain(Handle* h, int l, int B, cudaStream_t st) {
    const LmLayerW& k = h->w.layer[l];
    const bool last = l + 1 == kLayers;
    ChainMaps maps;
    ChainArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.split = h->policy == kPolicySplit;
    ca.bar = h->chain_bar + (size_t)l * 8;
    auto mk = [&](int idx, const bf16* hi, const bf16* lo, int rows, int K, int ld, int box) -> int {
        MB_CK(h, build_chain_map(&maps.m[idx], hi, rows, K, ld, box));
        if (ca.split) MB_CK(h, build_chain_map(&maps.m[idx + 1], lo, rows, K, ld, box));
        else maps.m[idx + 1] = maps.m[idx];
        return 0;
    };
    MB_TRY(mk(0, h->la_hi, h->la_lo, B, kHidden, kHidden, 128));
    MB_TRY(mk(2, h->lh_hi, h->lh_lo, B, kInter, kInter, 128));
    MB_TRY(mk(4, k.o.hi, k.o.lo, kHidden, kHidden, kHidden, 16));
    MB_TRY(mk(6, k.gu.hi, k.gu.lo, 2 * kInter, kHidden, kHidden, 32));
    MB_TRY(mk(8, k.down.hi, k.down.lo, kHidden, kInter, kInter, 16));
    if (!last) MB_TRY(mk(10, h->w.layer[l + 1].qkv.hi, h->w.layer[l + 1].qkv.lo, kQkvDim, kHidden, kHidden, 16));
    else { maps.m[10] = maps.m[0]; maps.m[11] = maps.m[0]; }
    int n = 0;
    {   // o_proj, split-K 3 -> partial
        ChainOp& op = ca.op[n++];
        op.kind = CH_GEMM; op.map_a = 0; op.map_b = 4; op.bn = 16; op.epi = EPI_GENERIC;
        op.g = gemm_base(h, h->la_hi, h->la_lo, kHidden, k.o, kHidden, B, kHidden, kHidden);
        op.g.split_k = 3; op.g.partial = h->gemm_partial;
    }
    {   // x += partials; planes = RMSNorm(x) * post_attention_layernorm
        ChainOp& op = ca.op[n++];
        op.kind = CH_ADDNORM; op.g.M = B; op.x = h->x; op.partial = h->gemm_partial; op.n_partial = 3; op.w = k.ln2;
        op.hi = h->la_hi; op.lo = lo_of(h, h->la_lo);
    }
    {   // gate/up + SwiGLU -> hidden planes
        ChainOp& op = ca.op[n++];
        op.kind = CH_GEMM; op.map_a = 0; op.map_b = 6; op.bn = 32; op.epi = EPI_SWIGLU;
        op.g = gemm_base(h, h->la_hi, h->la_lo, kHidden, k.gu, kHidden, B, 2 * kInter, kHidden);
        op.g.out_hi = h->lh_hi; op.g.out_lo = lo_of(h, h->lh_lo); op.g.ldp = kInter;
    }
    {   // down, split-K 4 -> partial
        ChainOp& op = ca.op[n++];
        op.kind = CH_GEMM; op.map_a = 2; op.map_b = 8; op.bn = 16; op.epi = EPI_GENERIC;
        op.g = gemm_base(h, h->lh_hi, h->lh_lo, kInter, k.down, kInter, B, kHidden, kInter);
        op.g.split_k = 4; op.g.partial = h->gemm_partial;
    }
    {   // x += partials; planes = RMSNorm(x) * (next input_layernorm | final norm)
        ChainOp& op = ca.op[n++];
        op.kind = CH_ADDNORM; op.g.M = B; op.x = h->x; op.partial = h->gemm_partial; op.n_partial = 4;
        op.w = last ? h->w.lm_norm : h->w.layer[l + 1].ln1;
        op.hi = h->la_hi; op.lo = lo_of(h, h->la_lo);
    }
    if (!last) {   // next layer's QKV + RoPE + KV-cache write
        ChainOp& op = ca.op[n++];
        op.kind = CH_GEMM; op.map_a = 0; op.map_b = 10; op.bn = 16; op.epi = EPI_QKV_ROPE;
        GemmArgs& g = op.g;
        g = gemm_base(h, h->la_hi, h->la_lo, kHidden, h->w.layer[l + 1].qkv, kHidden, B, kQkvDim, kHidden);
        g.q_out = h->q;
        g.k_cache = kv_layer(h, h->kcache, l + 1); g.v_cache = kv_layer(h, h->vcache, l + 1);
        g.rope_cos = h->w.rope_cos; g.rope_sin = h->w.rope_sin;
        g.rows_per_seq = 1; g.pos_base = kPrefix - 1; g.d_pos = h->d_step;
        g.t_max = h->t_max; g.kv_fmt = h->kv_fmt;
    }
    ca.n_ops = n;
    MB_CK(h, launch_decode_chain(maps, ca, st));
    h->launches++;
    return 0;
}

// one transformer layer over M rows of h->x.  prefill: rows_per_seq = 389; decode: rows_per_seq = 1.
int lm_layer(Handle* h, int l, int B, int rows_per_seq, bool decode, cudaStream_t st) {
    const LmLayerW& k = h->w.layer[l];
    const int M = B * rows_per_seq;
    MB_TRY(run_norm(h, NORM_RMS, h->x, k.ln1, nullptr, M, kHidden, h->la_hi, h->la_lo, nullptr, 1, 0, 0, st));
    {
        GemmArgs g = gemm_base(h, h->la_hi, h->la_lo, kHidden, k.qkv, kHidden, M, kQkvDim, kHidden);
        g.q_out = h->q;
        g.k_cache = kv_layer(h, h->kcache, l); g.v_cache = kv_layer(h, h->vcache, l);
        g.rope_cos = h->w.rope_cos; g.rope_sin = h->w.rope_sin;
        g.rows_per_seq = rows_per_seq;
        g.pos_base = decode ? kPrefix - 1 : 0;
        g.d_pos = decode ? h->d_step : nullptr;
        g.t_max = h->t_max; g.kv_fmt = h->kv_fmt;
        MB_TRY(run_gemm(h, g, EPI_QKV_ROPE, st));
    }
    if (decode) {
        MB_TRY(run_decode_attention(h, l, B, st));
    } else {
        static const bool fp32_attn = getenv("MB_ATTN_FP32") != nullptr;      // CUDA-core fp32 kernel, kept for A/B checks
        if (fp32_attn)
            MB_CK(h, launch_prefill_attention(h->q, kv_layer(h, h->kcache, l), kv_layer(h, h->vcache, l),
                                              h->kv_fmt, B, rows_per_seq, h->t_max, h->la_hi,
                                              lo_of(h, h->la_lo), st));
        else
            MB_CK(h, launch_prefill_attention_mma(h->q, kv_layer(h, h->kcache, l), kv_layer(h, h->vcache, l),
                                                  h->kv_fmt, B, rows_per_seq, h->t_max, h->la_hi,
                                                  lo_of(h, h->la_lo), st));
        h->launches++;
    }
    {
        GemmArgs g = gemm_base(h, h->la_hi, h->la_lo, kHidden, k.o, kHidden, M, kHidden, kHidden);
        g.residual = h->x; g.ldr = kHidden; g.out_f32 = h->x; g.ldo = kHidden;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    MB_TRY(run_norm(h, NORM_RMS, h->x, k.ln2, nullptr, M, kHidden, h->la_hi, h->la_lo, nullptr, 1, 0, 0, st));
    {
        GemmArgs g = gemm_base(h, h->la_hi, h->la_lo, kHidden, k.gu, kHidden, M, 2 * kInter, kHidden);
        g.out_hi = h->lh_hi; g.out_lo = lo_of(h, h->lh_lo); g.ldp = kInter;
        MB_TRY(run_gemm(h, g, EPI_SWIGLU, st));
    }
    {
        GemmArgs g = gemm_base(h, h->lh_hi, h->lh_lo, kInter, k.down, kInter, M, kHidden, kInter);
        g.residual = h->x; g.ldr = kHidden; g.out_f32 = h->x; g.ldo = kHidden;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
    }
    return 0;
}

// lm_head over the planes in la_hi/la_lo.  fused: write per-16-column argmax candidates for sample_kernel instead of
// the [B,49152] logits (the decode loop); otherwise the full fp32 logits (parity dumps, mb_prefill).
int lm_head_gemm(Handle* h, int B, bool fused, cudaStream_t st) {
    GemmArgs g = gemm_base(h, h->la_hi, h->la_lo, kHidden, h->w.head, kHidden, B, kVocab, kHidden);
    g.trace = h->trace; g.trace_id = 8000;
    fused = fused && h->engine == 1;                      // the argmax epilogue exists in the row-per-thread engine only
    h->logits_fused = fused;
    if (fused) {
        g.cand_val = h->cand_val; g.cand_idx = h->cand_idx;
        return run_gemm(h, g, EPI_ARGMAX, st);
    }
    g.out_f32 = h->logits; g.ldo = kVocab;
    return run_gemm(h, g, EPI_GENERIC, st);
}

// final RMSNorm of one row per sequence + lm_head
int lm_head(Handle* h, int B, int row_stride, int row_off, bool fused, cudaStream_t st) {
    MB_TRY(run_norm(h, NORM_RMS, h->x, h->w.lm_norm, nullptr, B, kHidden, h->la_hi, h->la_lo, nullptr, row_stride,
                    row_off, 0, st));
    return lm_head_gemm(h, B, fused, st);
}

int decode_step(Handle* h, int B, bool fused, cudaStream_t st) {
    if (B <= 128 && h->engine == 1 && getenv("MB_DECODE_UNFUSED") == nullptr && getenv("MB_CHAIN") != nullptr) {
        // Experimental (MB_CHAIN=1): 2 kernels per layer, decode attention + one persistent chain kernel whose phases
        // are separated by software grid barriers.  Parity-green, but measured 10 % slower than the PDL-linked
        // per-phase kernels below (1.94 vs 1.76 ms/step at B=128): a grid barrier plus the TMA round trip after it
        // costs as much as a programmatic kernel boundary.  Kept as the starting point of a deeper fusion.
        MB_CK(h, cudaMemsetAsync(h->chain_bar, 0, sizeof(unsigned) * kLayers * 8, st));
        MB_CK(h, launch_add_rmsnorm(h->x, nullptr, 0, B, h->w.layer[0].ln1, h->la_hi, lo_of(h, h->la_lo), st));
        h->launches++;
        {
            const LmLayerW& k0 = h->w.layer[0];
            GemmArgs g = gemm_base(h, h->la_hi, h->la_lo, kHidden, k0.qkv, kHidden, B, kQkvDim, kHidden);
            g.q_out = h->q;
            g.k_cache = kv_layer(h, h->kcache, 0); g.v_cache = kv_layer(h, h->vcache, 0);
            g.rope_cos = h->w.rope_cos; g.rope_sin = h->w.rope_sin;
            g.rows_per_seq = 1; g.pos_base = kPrefix - 1; g.d_pos = h->d_step;
            g.t_max = h->t_max; g.kv_fmt = h->kv_fmt;
            MB_TRY(run_gemm(h, g, EPI_QKV_ROPE, st));
        }
        for (int l = 0; l < kLayers; ++l) {
            MB_TRY(run_decode_attention(h, l, B, st));
            MB_TRY(lm_layer_decode_chain(h, l, B, st));
        }
        return lm_head_gemm(h, B, fused, st);
    }
    if (B <= 128 && getenv("MB_DECODE_UNFUSED") == nullptr) {
        // Row groups: the batch is cut into G contiguous row groups whose per-layer kernel chains (7 PDL-linked
        // kernels per layer) run on G streams.  Each chain is latency-bound (a dependent kernel every ~6 us) except
        // its HBM-bound attention kernel, so while one group streams its K/V the other groups' GEMM / norm phases
        // run; the compact GEMM variants keep two CTAs per SM so that the chains do not queue behind each other's
        // shared memory.  Rows never interact, so the result does not depend on G.  lm_head runs once over all rows
        // after the join (its 113 MB weight stream is the cost, not the rows).
        const int G = decode_groups(h, B);
        const int compact = decode_compact(G);
        // Chains that start together stay in lock-step (all groups stream K/V at the same time, then all wait on
        // their GEMM phases), which gains nothing.  stagger 1: group g starts when group g-1 has finished its first
        // attention, half a layer period later.  stagger 2: the attention kernels of a layer are chained through
        // events (g waits for g-1, group 0 of the next layer for the last group), so exactly one group streams K/V at
        // any time while the others are in their GEMM / norm phases.
        const int stagger = decode_stagger(G);
        if (G > 1) MB_CK(h, cudaEventRecord(h->ev_fork, st));
        for (int gi = 1; gi < G; ++gi) MB_CK(h, cudaStreamWaitEvent(h->grp_stream[gi - 1], h->ev_fork, 0));
        for (int l = 0; l < kLayers; ++l) {
            for (int gi = 0; gi < G; ++gi) {
                const int r0 = (int)(((long long)B * gi) / G), n = (int)(((long long)B * (gi + 1)) / G) - r0;
                cudaStream_t sg = gi == 0 ? st : h->grp_stream[gi - 1];
                if (l == 0) {
                    if (stagger == 1 && gi > 0) MB_CK(h, cudaStreamWaitEvent(sg, h->ev_attn[gi - 1], 0));
                    PdlOff off(gi > 0);                     // a forked chain starts behind event edges, not a kernel
                    MB_CK(h, launch_add_rmsnorm(h->x + (size_t)r0 * kHidden, nullptr, 0, n, h->w.layer[0].ln1,
                                                h->la_hi + (size_t)r0 * kHidden, lo_of(h, h->la_lo + (size_t)r0 * kHidden), sg));
                    h->launches++;
                }
                cudaEvent_t wait_ev = nullptr, rec_ev = nullptr;
                if (stagger == 2) {
                    wait_ev = gi > 0 ? h->ev_attn[gi - 1] : (l > 0 ? h->ev_attn[G - 1] : nullptr);
                    rec_ev = h->ev_attn[gi];
                } else if (stagger == 1 && l == 0 && gi + 1 < G) {
                    rec_ev = h->ev_attn[gi];
                }
                MB_TRY(lm_layer_decode_fused(h, l, gi, r0, n, B, compact,
                                             l + 1 < kLayers ? h->w.layer[l + 1].ln1 : h->w.lm_norm, sg, wait_ev, rec_ev));
            }
        }
        for (int gi = 1; gi < G; ++gi) MB_CK(h, cudaEventRecord(h->ev_join[gi - 1], h->grp_stream[gi - 1]));
        for (int gi = 1; gi < G; ++gi) MB_CK(h, cudaStreamWaitEvent(st, h->ev_join[gi - 1], 0));
        PdlOff off(G > 1);                                  // after the join lm_head has G predecessors
        return lm_head_gemm(h, B, fused, st);
    }
    for (int l = 0; l < kLayers; ++l) MB_TRY(lm_layer(h, l, B, 1, true, st));
    return lm_head(h, B, 1, 0, fused, st);
}

int sample_and_advance(Handle* h, int B, int max_len, float temperature, float top_p, int eos_id, float* logits_dump,
                       const int* forced, cudaStream_t st) {
    SampleArgs a;
    a.logits = h->logits_fused ? nullptr : h->logits;
    a.cand_val = h->cand_val; a.cand_idx = h->cand_idx; a.n_cand = kVocab / 16;
    a.embed = h->w.embed; a.B = B; a.max_len = max_len; a.eos_id = eos_id;
    a.temperature = temperature; a.top_p = top_p; a.d_step = h->d_step; a.tokens_out = h->d_tokens;
    a.forced = forced; a.x_next = h->x; a.done = h->d_done; a.logits_dump = logits_dump;
    MB_CK(h, launch_sample(a, st));
    MB_CK(h, launch_step_advance(h->d_step, h->d_done, B, h->d_stop, h->w.rope_cos, h->w.rope_sin, kPrefix - 1, h->rope_cur, st));
    h->launches += 2;
    return 0;
}

int check_ready(Handle* h, int B) {
    if (!h->bound) return fail(h, "weights not bound (call mb_bind_weights first)");
    if (B < 1 || B > h->max_batch) return fail(h, "batch size out of range for this handle");
    return 0;
}

int do_prefill(Handle* h, int B, float* logits_out, cudaStream_t st) {
    for (int l = 0; l < kLayers; ++l) MB_TRY(lm_layer(h, l, B, kPrefix, false, st));
    MB_TRY(lm_head(h, B, kPrefix, kPrefix - 1, /*fused=*/false, st));   // step-0 logits stay available to mb_prefill callers
    if (logits_out)
        MB_CK(h, cudaMemcpyAsync(logits_out, h->logits, (size_t)B * kVocab * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int do_decode(Handle* h, int B, int max_len, float temperature, float top_p, int eos_id, int* tokens_out,
              int tokens_on_host, int* steps_out_host, float* logits_dump, const int* forced, cudaStream_t st) {
    if (max_len < 1 || max_len > h->max_new) return fail(h, "max_len out of range for this handle");
    MB_CK(h, cudaMemsetAsync(h->d_step, 0, sizeof(int), st));
    MB_CK(h, cudaMemsetAsync(h->d_done, 0, sizeof(int) * B, st));
    MB_CK(h, cudaMemsetAsync(h->d_stop, 0xff, sizeof(int), st));
    MB_CK(h, cudaMemsetAsync(h->d_tokens, 0, sizeof(int) * (size_t)B * max_len, st));
    // step 0: logits come from the prefill
    MB_TRY(sample_and_advance(h, B, max_len, temperature, top_p, eos_id, logits_dump, forced, st));
    const bool use_graph = !logits_dump && !forced && getenv("MB_NO_GRAPH") == nullptr;
    int stop = -1;
    if (use_graph && max_len > 1) {
        const bool hit = h->graph && h->g_B == B && h->g_max_len == max_len && h->g_eos == eos_id &&
                         h->g_temp == temperature && h->g_stream == st;
        if (!hit) {
            if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
            cudaGraph_t graph = nullptr;
            const long long before = h->launches;
            MB_CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc = decode_step(h, B, /*fused=*/true, st);
            if (rc == 0) rc = sample_and_advance(h, B, max_len, temperature, top_p, eos_id, nullptr, nullptr, st);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
            MB_CK(h, ce);
            h->g_launches = (int)(h->launches - before);
            h->launches = before;
            MB_CK(h, cudaGraphInstantiate(&h->graph, graph, 0));
            cudaGraphDestroy(graph);
            h->g_B = B; h->g_max_len = max_len; h->g_eos = eos_id; h->g_temp = temperature; h->g_stream = st;
        }
    }
    for (int s = 1; s < max_len; ++s) {
        if (use_graph) {
            MB_CK(h, cudaGraphLaunch(h->graph, st));
            h->launches += h->g_launches;
        } else {
            MB_TRY(decode_step(h, B, /*fused=*/logits_dump == nullptr, st));
            MB_TRY(sample_and_advance(h, B, max_len, temperature, top_p, eos_id, logits_dump, forced, st));
        }
        if ((s & 15) == 0) {          // the reference syncs every step (wrapper.py:248); poll the stop flag sparsely
            MB_CK(h, cudaMemcpyAsync(&stop, h->d_stop, sizeof(int), cudaMemcpyDeviceToHost, st));
            MB_CK(h, cudaStreamSynchronize(st));
            if (stop >= 0) break;
        }
    }
    MB_CK(h, cudaMemcpyAsync(&stop, h->d_stop, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (tokens_out)
        MB_CK(h, cudaMemcpyAsync(tokens_out, h->d_tokens, sizeof(int) * (size_t)B * max_len,
                                 tokens_on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    MB_CK(h, cudaStreamSynchronize(st));
    if (steps_out_host) *steps_out_host = stop >= 0 ? stop : max_len;
    return 0;
}

int do_generate(Handle* h, const float* w1, const float* w2, const int* ids, int B, int max_len, float temperature,
                float top_p, int eos_id, int* tokens_out, int tokens_on_host, int* steps_out_host, cudaStream_t st) {
    MB_TRY(frontend(h, w1, B, nullptr, h->bn, st));
    MB_TRY(frontend(h, w2, B, nullptr, h->bn + (size_t)B * kFrames * kMels, st));
    MB_TRY(encoder_body(h, 2 * B, -1, nullptr, nullptr, st));
    MB_CK(h, launch_prefix(h->rows33, ids, h->w.embed, B, h->x, st));
    h->launches++;
    h->prefix_B = B;
    MB_TRY(do_prefill(h, B, nullptr, st));
    return do_decode(h, B, max_len, temperature, top_p, eos_id, tokens_out, tokens_on_host, steps_out_host, nullptr,
                     nullptr, st);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" {

int mb_version(void) { return 1; }
int mb_weight_entry_count(void) { return (int)layout().t.entries.size(); }
const char* mb_weight_entry_name(int i) { return layout().t.entries[i].name.c_str(); }
long long mb_weight_entry_offset(int i) { return (long long)layout().t.entries[i].offset; }
long long mb_weight_entry_bytes(int i) { return (long long)layout().t.entries[i].bytes; }
long long mb_weights_size(void) { return (long long)layout().t.total; }

const char* mb_last_error(void* hv) { return hv ? reinterpret_cast<Handle*>(hv)->err : g_create_error; }

void mb_destroy(void* hv) {
    if (!hv) return;
    Handle* h = reinterpret_cast<Handle*>(hv);
    cudaSetDevice(h->device);
    if (h->graph) cudaGraphExecDestroy(h->graph);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (int i = 0; i < kMaxGroups - 1; ++i) {
        if (h->grp_stream[i]) cudaStreamDestroy(h->grp_stream[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int i = 0; i < kMaxGroups; ++i)
        if (h->ev_attn[i]) cudaEventDestroy(h->ev_attn[i]);
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

static int create_body(Handle* h) {
    const size_t N = 2 * (size_t)h->max_batch;           // clips
    const size_t M0 = N * 4096;                          // stage-0 tokens
    MB_TRY(dev_alloc(h, &h->bn, N * kFrames * kMels));
    MB_TRY(dev_alloc(h, &h->xa, M0 * kEmbed));
    MB_TRY(dev_alloc(h, &h->xb, M0 * kEmbed / 2));
    MB_TRY(dev_alloc(h, &h->qkv, M0 * 3 * kEmbed));
    MB_TRY(dev_alloc(h, &h->a_hi, M0 * kEmbed));
    MB_TRY(dev_alloc(h, &h->a_lo, M0 * kEmbed));
    MB_TRY(dev_alloc(h, &h->h_hi, M0 * 4 * kEmbed));
    MB_TRY(dev_alloc(h, &h->h_lo, M0 * 4 * kEmbed));
    MB_TRY(dev_alloc(h, &h->ytail, N * 64 * kEncOut));
    MB_TRY(dev_alloc(h, &h->latent, N * kEncOut));
    MB_TRY(dev_alloc(h, &h->frames, N * 32 * kEncOut));
    MB_TRY(dev_alloc(h, &h->col_hi, N * 32 * 6 * kEncOut));
    MB_TRY(dev_alloc(h, &h->col_lo, N * 32 * 6 * kEncOut));
    MB_TRY(dev_alloc(h, &h->cls_hi, N * 32 * kClassesPad));
    MB_TRY(dev_alloc(h, &h->cls_lo, N * 32 * kClassesPad));
    MB_CK(h, cudaMemset(h->cls_hi, 0, N * 32 * kClassesPad * 2));
    MB_CK(h, cudaMemset(h->cls_lo, 0, N * 32 * kClassesPad * 2));
    MB_TRY(dev_alloc(h, &h->a33_hi, N * kAudioRows * kEncOut));
    MB_TRY(dev_alloc(h, &h->a33_lo, N * kAudioRows * kEncOut));
    MB_TRY(dev_alloc(h, &h->g33_hi, N * kAudioRows * kProj));
    MB_TRY(dev_alloc(h, &h->g33_lo, N * kAudioRows * kProj));
    MB_TRY(dev_alloc(h, &h->e1, N * kAudioRows * kProj));
    MB_TRY(dev_alloc(h, &h->s33, N * kAudioRows * kProj));
    MB_TRY(dev_alloc(h, &h->rows33, N * kAudioRows * kProj));
    const size_t B = h->max_batch, M = B * kPrefix;
    MB_TRY(dev_alloc(h, &h->x, M * kHidden));
    MB_TRY(dev_alloc(h, &h->q, M * kHidden));
    MB_TRY(dev_alloc(h, &h->la_hi, M * kHidden));
    MB_TRY(dev_alloc(h, &h->la_lo, M * kHidden));
    MB_TRY(dev_alloc(h, &h->lh_hi, M * kInter));
    MB_TRY(dev_alloc(h, &h->lh_lo, M * kInter));
    MB_TRY(dev_alloc(h, &h->logits, B * kVocab));
    MB_TRY(dev_alloc(h, &h->cand_val, B * (kVocab / 16)));
    MB_TRY(dev_alloc(h, &h->cand_idx, B * (kVocab / 16)));
    h->t_max = kPrefix + h->max_new;
    h->kv_layer_elems = B * kKvHeads * h->t_max * kHeadDim;
    // KV format: fp32 rows (split policy default), bf16 (fast policy), or 24-bit rows (MB_KV24=1 / policy split24)
    const size_t kv_bytes = (h->kv_layer_elems / kHeadDim) * kLayers * kv_row_bytes(h->kv_fmt);
    char* kc = nullptr; char* vc = nullptr;
    MB_TRY(dev_alloc(h, &kc, kv_bytes));
    MB_TRY(dev_alloc(h, &vc, kv_bytes));
    h->kcache = kc; h->vcache = vc;
    MB_TRY(dev_alloc(h, &h->part_acc, B * kHeads * kMaxAttnSplit * kHeadDim));
    MB_TRY(dev_alloc(h, &h->part_ml, B * kHeads * kMaxAttnSplit * 2));
    MB_TRY(dev_alloc(h, &h->gemm_partial, (size_t)kMaxGroups * kMaxSplitK * 128 * kHidden));
    MB_TRY(dev_alloc(h, &h->qkv_part, (size_t)kMaxGroups * kQkvSplitMax * 128 * kQkvDim));
    MB_TRY(dev_alloc(h, &h->rope_cur, 64));
    MB_TRY(dev_alloc(h, &h->gu_part, (size_t)kMaxGroups * kGuSplitMax * 128 * 2 * kInter));
    MB_TRY(dev_alloc(h, &h->fix_counter, (size_t)kMaxGroups * 64));
    MB_CK(h, cudaMemset(h->fix_counter, 0, sizeof(unsigned) * kMaxGroups * 64));
    for (int i = 0; i < kMaxGroups - 1; ++i) {
        MB_CK(h, cudaStreamCreateWithFlags(&h->grp_stream[i], cudaStreamNonBlocking));
        MB_CK(h, cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
    }
    MB_CK(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < kMaxGroups; ++i) MB_CK(h, cudaEventCreateWithFlags(&h->ev_attn[i], cudaEventDisableTiming));
    MB_TRY(dev_alloc(h, &h->chain_bar, (size_t)kLayers * 8));
    MB_TRY(dev_alloc(h, &h->d_tokens, B * h->max_new));
    MB_TRY(dev_alloc(h, &h->d_done, B));
    MB_TRY(dev_alloc(h, &h->d_step, 4));
    MB_TRY(dev_alloc(h, &h->d_stop, 4));
    MB_TRY(dev_alloc(h, &h->d_ids, B * kTextLen));
    MB_TRY(dev_alloc(h, &h->wave_stage, N * kClipSamples));
    return 0;
}

void* mb_create(int device, int max_batch, int max_new_tokens, int policy) {
    g_create_error[0] = 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        snprintf(g_create_error, sizeof(g_create_error), "no CUDA device available (%s); mellow_b200 has no CPU fallback",
                 cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= count || max_batch < 1 || max_new_tokens < 1 || kPrefix + max_new_tokens > kMaxPos ||
        (policy != kPolicySplit && policy != kPolicyFast && policy != kPolicySplit24)) {
        snprintf(g_create_error, sizeof(g_create_error), "mb_create: invalid argument");
        return nullptr;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        snprintf(g_create_error, sizeof(g_create_error), "device %d is sm_%d%d; this library is built for sm_100a only",
                 device, prop.major, prop.minor);
        return nullptr;
    }
    cudaSetDevice(device);
    Handle* h = new Handle();
    h->device = device; h->max_batch = max_batch; h->max_new = max_new_tokens;
    {
        // kPolicySplit24 = kPolicySplit with the KV cache stored at 24 bits (MB_KV24=1 turns it on for kPolicySplit too)
        const char* e = getenv("MB_KV24");
        const bool kv24 = policy == kPolicySplit24 || (e && atoi(e) != 0);
        h->policy = policy == kPolicySplit24 ? kPolicySplit : policy;
        h->kv_fmt = policy == kPolicyFast ? kKvBf16 : (kv24 ? kKvF24 : kKvF32);
    }
    build_table(h->t, h->w);
    // a blocking stream: implicitly ordered with work the caller issued on the legacy default stream
    if (cudaStreamCreate(&h->own_stream) != cudaSuccess || create_body(h) != 0) {
        snprintf(g_create_error, sizeof(g_create_error), "%s", h->err);
        mb_destroy(h);
        return nullptr;
    }
    return h;
}

int mb_bind_weights(void* hv, const void* dev_arena, long long nbytes) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (nbytes != (long long)h->t.total) return fail(h, "weight arena size mismatch");
    if (((uintptr_t)dev_arena & 255) != 0) return fail(h, "weight arena must be 256-byte aligned");
    for (auto& e : h->t.entries) *e.slot = reinterpret_cast<const char*>(dev_arena) + e.offset;
    h->bound = true;
    return 0;
}

long long mb_workspace_bytes(void* hv) { return (long long)reinterpret_cast<Handle*>(hv)->ws_bytes; }
int mb_set_gemm_engine(void* hv, int engine) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (engine != 0 && engine != 1) return fail(h, "unknown GEMM engine");
    h->engine = engine;
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    return 0;
}
int mb_set_decode_groups(void* hv, int groups) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (groups < 0 || groups > kMaxGroups) return fail(h, "decode row groups must be 0 (automatic) .. 4");
    h->groups = groups;
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    return 0;
}
int mb_set_decode_qkv_split(void* hv, int nsplit) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (nsplit != -1 && nsplit != 0 && nsplit != 3 && nsplit != 9) return fail(h, "decode QKV split must be -1 (default), 0, 3 or 9");
    h->qkv_split = nsplit;
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    return 0;
}
int mb_set_trace(void* hv, void* dev_trace_buf) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    h->trace = reinterpret_cast<TraceBuf*>(dev_trace_buf);
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    return 0;
}
long long mb_kernel_launches(void* hv) { return reinterpret_cast<Handle*>(hv)->launches; }

int mb_frontend(void* hv, const float* wave, int n_clips, float* logmel_out, float* bn_out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (!h->bound) return fail(h, "weights not bound");
    if (n_clips < 1) return fail(h, "n_clips < 1");
    cudaSetDevice(h->device);
    return frontend(h, wave, n_clips, logmel_out, bn_out, pick_stream(h, stream));
}

int mb_encode(void* hv, const float* wave1, const float* wave2, int B, float* rows_out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    cudaSetDevice(h->device);
    cudaStream_t st = pick_stream(h, stream);
    MB_TRY(frontend(h, wave1, B, nullptr, h->bn, st));
    MB_TRY(frontend(h, wave2, B, nullptr, h->bn + (size_t)B * kFrames * kMels, st));
    return encoder_body(h, 2 * B, -1, nullptr, rows_out, st);
}

// SURVEY 8 row f4 (reference mellow.py:100-108 returns od1/od2 = the encoder's output dict; htsat.py:782-796,950-955).
// The TSCAM conv is re-run without its sigmoid on the im2col rows the last encoder pass left in the workspace (a
// 0.12 TFLOP GEMM, only when a caller asks for the heads), then heads_kernel forms the two classification outputs.
int mb_encode_heads(void* hv, int n_clips, float* clipwise_out, float* framewise_rows_out, float* latent_out,
                    float* frame_embed_out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (!h->bound) return fail(h, "weights not bound");
    if (n_clips < 1 || n_clips != h->enc_clips) return fail(h, "mb_encode_heads: call it after mb_encode / mb_generate with n_clips = 2*B of that call");
    cudaSetDevice(h->device);
    cudaStream_t st = pick_stream(h, stream);
    if (clipwise_out || framewise_rows_out) {
        float* logits = h->qkv;                            // encoder scratch, free after the Swin stages
        GemmArgs g = gemm_base(h, h->col_hi, h->col_lo, 6 * kEncOut, h->w.tscam, 6 * kEncOut, n_clips * 32, kClasses, 6 * kEncOut);
        g.bias = h->w.tscam_b; g.out_f32 = logits; g.ldo = kClassesPad;
        MB_TRY(run_gemm(h, g, EPI_GENERIC, st));
        MB_CK(h, launch_heads(logits, kClassesPad, n_clips, clipwise_out, framewise_rows_out, st));
        h->launches++;
    }
    if (latent_out)
        MB_CK(h, cudaMemcpyAsync(latent_out, h->latent, (size_t)n_clips * kEncOut * 4, cudaMemcpyDeviceToDevice, st));
    if (frame_embed_out)
        MB_CK(h, cudaMemcpyAsync(frame_embed_out, h->frames, (size_t)n_clips * 32 * kEncOut * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int mb_encode_tap(void* hv, const float* wave, int n_clips, int stage, float* out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (!h->bound) return fail(h, "weights not bound");
    if (n_clips < 1 || n_clips > 2 * h->max_batch || stage < 0 || stage > 5) return fail(h, "bad tap arguments");
    cudaSetDevice(h->device);
    cudaStream_t st = pick_stream(h, stream);
    MB_TRY(frontend(h, wave, n_clips, nullptr, h->bn, st));
    return encoder_body(h, n_clips, stage, out, nullptr, st);
}

int mb_prefix(void* hv, const int* input_ids, int B, float* prefix_out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    cudaSetDevice(h->device);
    cudaStream_t st = pick_stream(h, stream);
    MB_CK(h, launch_prefix(h->rows33, input_ids, h->w.embed, B, h->x, st));
    h->launches++;
    h->prefix_B = B;
    if (prefix_out)
        MB_CK(h, cudaMemcpyAsync(prefix_out, h->x, (size_t)B * kPrefix * kHidden * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int mb_set_prefix(void* hv, const float* prefix, int B, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    cudaSetDevice(h->device);
    MB_CK(h, cudaMemcpyAsync(h->x, prefix, (size_t)B * kPrefix * kHidden * 4, cudaMemcpyDeviceToDevice, pick_stream(h, stream)));
    h->prefix_B = B;
    return 0;
}

int mb_prefill(void* hv, int B, float* logits_out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    if (h->prefix_B != B) return fail(h, "mb_prefill: no prefix of this batch size (call mb_prefix / mb_set_prefix)");
    cudaSetDevice(h->device);
    return do_prefill(h, B, logits_out, pick_stream(h, stream));
}

int mb_decode(void* hv, int B, int max_len, float temperature, float top_p, int eos_id, int* tokens_out,
              int* steps_out_host, float* logits_dump, const int* forced_tokens, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    cudaSetDevice(h->device);
    return do_decode(h, B, max_len, temperature, top_p, eos_id, tokens_out, 0, steps_out_host, logits_dump,
                     forced_tokens, pick_stream(h, stream));
}

int mb_generate(void* hv, const float* wave1, const float* wave2, const int* input_ids, int B, int max_len,
                float temperature, float top_p, int eos_id, int* tokens_out, int* steps_out_host, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    cudaSetDevice(h->device);
    return do_generate(h, wave1, wave2, input_ids, B, max_len, temperature, top_p, eos_id, tokens_out, 0,
                       steps_out_host, pick_stream(h, stream));
}

int mb_generate_host(void* hv, const float* wave1_host, const float* wave2_host, const int* input_ids_host, int B,
                     int max_len, float temperature, float top_p, int eos_id, int* tokens_out_host,
                     int* steps_out_host, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    cudaSetDevice(h->device);
    cudaStream_t st = pick_stream(h, stream);
    const size_t wbytes = (size_t)B * kClipSamples * sizeof(float);
    float* w1 = h->wave_stage;
    float* w2 = h->wave_stage + (size_t)B * kClipSamples;
    MB_CK(h, cudaMemcpyAsync(w1, wave1_host, wbytes, cudaMemcpyHostToDevice, st));
    MB_CK(h, cudaMemcpyAsync(w2, wave2_host, wbytes, cudaMemcpyHostToDevice, st));
    MB_CK(h, cudaMemcpyAsync(h->d_ids, input_ids_host, (size_t)B * kTextLen * sizeof(int), cudaMemcpyHostToDevice, st));
    return do_generate(h, w1, w2, h->d_ids, B, max_len, temperature, top_p, eos_id, tokens_out_host, 1,
                       steps_out_host, st);
}

int mb_audio_resample(void* hv, const float* pcm, long long n_in, int orig, int new_, const float* kernel, int klen,
                      int width, float* out, long long n_out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (n_in < 1 || orig < 1 || new_ < 1 || klen != 2 * width + orig || n_out < 1) return fail(h, "mb_audio_resample: bad arguments");
    cudaSetDevice(h->device);
    MB_CK(h, launch_resample(pcm, n_in, orig, new_, kernel, klen, width, out, n_out, pick_stream(h, stream)));
    h->launches++;
    return 0;
}

int mb_audio_fit(void* hv, const float* samples, long long total, long long start, float* out, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    if (total < 1 || start < 0 || (total >= kClipSamples && start + kClipSamples > total) || (total < kClipSamples && start != 0))
        return fail(h, "mb_audio_fit: bad arguments");
    cudaSetDevice(h->device);
    MB_CK(h, launch_fit(samples, total, start, out, pick_stream(h, stream)));
    h->launches++;
    return 0;
}

int mb_bench_decode_attention(void* hv, int B, int ctx, int iters, void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    MB_TRY(check_ready(h, B));
    if (ctx < 1 || ctx > h->t_max) return fail(h, "ctx out of range");
    cudaSetDevice(h->device);
    cudaStream_t st = pick_stream(h, stream);
    const int step = ctx - kPrefix;
    MB_CK(h, cudaMemcpyAsync(h->d_step, &step, sizeof(int), cudaMemcpyHostToDevice, st));
    for (int i = 0; i < iters; ++i) {
        MB_TRY(run_decode_attention(h, i % kLayers, B, st, /*skip_done=*/false));
    }
    return 0;
}

int mb_op_gemm(void* hv, const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int act,
               void* stream) {
    Handle* h = reinterpret_cast<Handle*>(hv);
    cudaSetDevice(h->device);
    cudaStream_t st = pick_stream(h, stream);
    if (K % 32 != 0) return fail(h, "mb_op_gemm: K must be a multiple of 32");
    bf16 *ah = nullptr, *al = nullptr, *wh = nullptr, *wl = nullptr;
    const size_t na = (size_t)M * K, nw = (size_t)N * K;
    MB_CK(h, cudaMalloc((void**)&ah, na * 2));
    MB_CK(h, cudaMalloc((void**)&al, na * 2));
    MB_CK(h, cudaMalloc((void**)&wh, nw * 2));
    MB_CK(h, cudaMalloc((void**)&wl, nw * 2));
    int rc = 0;
    do {
        if (launch_split_planes(A, na, ah, al, st) != cudaSuccess || launch_split_planes(W, nw, wh, wl, st) != cudaSuccess) {
            rc = fail(h, "split_planes launch failed");
            break;
        }
        // the GEMM kernels prefetch W before their programmatic-dependency wait (weights are never produced by the
        // preceding kernel on the product path); here W was just written by split_planes, so drain the stream first
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(h, "stream sync failed"); break; }
        Plane wp{wh, wl};
        GemmArgs g = gemm_base(h, ah, al, K, wp, K, M, N, K);
        g.bias = bias; g.act = act; g.out_f32 = C; g.ldo = N;
        rc = run_gemm(h, g, EPI_GENERIC, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (rc == 0 && e != cudaSuccess) { snprintf(h->err, sizeof(h->err), "mb_op_gemm: %s", cudaGetErrorString(e)); rc = 1; }
    } while (0);
    cudaFree(ah); cudaFree(al); cudaFree(wh); cudaFree(wl);
    return rc;
}

}  // extern "C"
