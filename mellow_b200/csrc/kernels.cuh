// Launchers of the non-GEMM kernels (frontend.cu, encoder.cu, lm.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm.cuh"

namespace mb {

// ---- frontend.cu
struct FrontendW {
    const float* window;    // [1024] analysis window (row k=0 of the checkpoint's conv_real basis)
    const float* twiddle;   // [512][2] exp(-2 pi i k / 1024)
    const float* melW;      // [513][64]
    const int* mel_lo;      // [64] first non-zero bin of each mel filter
    const int* mel_hi;      // [64] one past the last non-zero bin
    const float* bn_scale;  // [64] gamma / sqrt(var + eps)
    const float* bn_shift;  // [64] beta - mean * scale
};
cudaError_t launch_logmel(const float* wave, int n_clips, const FrontendW& w, float* logmel_out, float* bn_out,
                          cudaStream_t st);
struct PatchW { const float* w; const float* b; const float* ln_w; const float* ln_b; };
cudaError_t launch_patch_embed(const float* bn, int n_clips, const PatchW& w, float* x_out, cudaStream_t st);

// ---- audio.cu
cudaError_t launch_resample(const float* x, long long n_in, int orig, int nw, const float* kern, int klen, int width,
                            float* y, long long n_out, cudaStream_t st);
cudaError_t launch_fit(const float* src, long long total, long long start, float* out, cudaStream_t st);

// ---- encoder.cu
enum { NORM_LN = 0, NORM_RMS = 1, NORM_LN_MERGE = 2 };
struct NormArgs {
    const float* x; const float* w; const float* b;
    bf16* out_hi; bf16* out_lo; float* out_f32;     // any subset
    int rows, C;
    int row_stride, row_off;                         // source row = r * row_stride + row_off (LN / RMS)
    int res;                                         // NORM_LN_MERGE: input grid side (tokens are res x res, C/4 wide)
};
cudaError_t launch_norm(const NormArgs& a, int kind, cudaStream_t st);
cudaError_t launch_window_attention_mma(const float* qkv, const float* relbias, bf16* out_hi, bf16* out_lo, int n_clips,
                                        int res, int C, int n_heads, int shift, cudaStream_t st);   // attn_mma.cu
cudaError_t launch_tail_gather(const float* y, int n_clips, float* latent, bf16* col_hi, bf16* col_lo, cudaStream_t st);
cudaError_t launch_assemble33(const float* latent, const float* frames, int n_clips, bf16* a_hi, bf16* a_lo,
                              cudaStream_t st);
cudaError_t launch_heads(const float* logits, int ld, int n_clips, float* clipwise, float* framewise, cudaStream_t st);
cudaError_t launch_gelu_planes(const float* x, size_t n, bf16* hi, bf16* lo, cudaStream_t st);
cudaError_t launch_split_planes(const float* x, size_t n, bf16* hi, bf16* lo, cudaStream_t st);

// ---- lm.cu
cudaError_t launch_prefix(const float* rows33, const int* ids, const float* embed, int B, float* prefix, cudaStream_t st);
// causal prefill attention on tcgen05 (attn_umma.cu): operand planes written by the QKV GEMM epilogue (gemm.cuh)
struct PrefillAttnPlanes {
    const bf16 *qp_hi, *qp_lo;      // [B*S][576] roped queries * 64^-0.5
    const bf16 *kp_hi, *kp_lo;      // [B][3][S][64] roped keys
    const bf16 *vt_hi, *vt_lo;      // [B][3][64][vt_ld] values, transposed
    int vt_ld;
};
cudaError_t launch_prefill_attention_umma(const PrefillAttnPlanes& p, int B, int S, bf16* out_hi, bf16* out_lo,
                                          cudaStream_t st);
// causal prefill attention on the legacy tensor path (mma.sync, attn_mma.cu)
cudaError_t launch_prefill_attention_mma(const float* q, const void* kc, const void* vc, int kv_fmt, int B, int S,
                                         int t_max, bf16* out_hi, bf16* out_lo, cudaStream_t st);
constexpr int kAttnDynParts = 4;       // most CTAs that share the keys of one (row, kv head) under DecodeAttnArgs::assign
struct DecodeAttnArgs {
    const float* q;                 // [B,576]
    const void* kc; const void* vc; // layer caches [B][3][t_max][64]
    int kv_fmt, B, t_max, nsplit;
    int tps;                           // key tiles (64 keys) owned by each split: ceil(ceil(t_max/64)/nsplit)
    int ctx_base; const int* d_step;   // ctx = ctx_base + *d_step  (keys 0..ctx-1)
    const int* done;                   // optional [B]: rows that already emitted the stop token skip their K/V stream
    // optional (nsplit == 1, B <= 128): work list written by step_advance_kernel.  assign[slot] = row |
    // part << 8 | nparts << 16 (nparts = 0: idle slot): the CTAs of finished rows take a share of the keys of the rows
    // that are still decoding; the nparts partial states of a (row, kv head) are merged by whichever CTA finishes last
    // (merge_count [B][3], self-resetting), in a fixed order, through part_acc / part_ml (stride kAttnDynParts)
    const int* assign; int* merge_count;
    float* part_acc; float* part_ml;   // [B][9][nsplit][64], [B][9][nsplit][2]
    bf16* out_hi; bf16* out_lo;        // [B,576]
    int variant;                       // 1 = warp-autonomous kernel (default), 0 = 64-key tile kernel
    int pf_keys;                       // (tile kernel) L2 prefetch of the CTA's immutable K/V history before the dependency wait:
                                       // 0 = off, -1 = all of it, n > 0 = keys below n only
    TraceBuf* trace; unsigned trace_id;
};
cudaError_t launch_decode_attention(const DecodeAttnArgs& a, cudaStream_t st);
struct SampleArgs {
    const float* logits;            // [B,V] (null when the candidates below are used)
    const float* cand_val; const int* cand_idx; int n_cand;   // [B][n_cand] per-16-column (max, first argmax) from lm_head
    const float* embed;             // [V,576] fp32 table (next-token embedding gather)
    int B, max_len, eos_id;
    float temperature, top_p;
    const int* d_step;              // current step (column of tokens_out)
    int* tokens_out;                // [B,max_len]
    const int* forced;              // optional teacher-forced tokens [B,max_len]: fed back instead of the argmax
    float* x_next;                  // [B,576] embedding of the token fed to the next step
    int* done;                      // [B] row has emitted eos
    float* logits_dump;             // optional [max_len][B][V] temperature-scaled logits
};
cudaError_t launch_add_rmsnorm(float* x, const float* partial, int n_partial, int M, const float* w, bf16* hi, bf16* lo,
                               cudaStream_t st, TraceBuf* trace = nullptr, unsigned trace_id = 0);
cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st);
cudaError_t launch_step_advance(int* d_step, const int* done, int B, int* d_stop_step, int* assign, int advance, cudaStream_t st);
// out[i] = embed[ids[i]] (fp32 rows of 576): lm.model.embed_tokens of the reference (wrapper.py:237)
cudaError_t launch_embed_rows(const int* ids, int n, const float* embed, float* out, cudaStream_t st);

}  // namespace mb
