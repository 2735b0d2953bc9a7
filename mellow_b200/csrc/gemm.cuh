// C[M,N] = A[M,K] * W[N,K]^T with fused epilogues; A and W are bf16 "planes" (hi, optional lo), fp32 accumulate.
// This header holds the argument block and the epilogues shared by every GEMM engine in the library.
#pragma once
#include "common.cuh"

namespace mb {

enum { EPI_GENERIC = 0, EPI_SWIGLU = 1, EPI_QKV_ROPE = 2 };
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SIGMOID = 2 };

struct GemmArgs {
    const bf16* A_hi; const bf16* A_lo; int lda;      // activations [M,K] (lo may be null when passes==1)
    const bf16* W_hi; const bf16* W_lo; int ldw;      // weights [N,K], K contiguous (nn.Linear layout)
    int M, N, K;                                      // K % 32 == 0
    int passes;                                       // 3: hi*hi + hi*lo + lo*hi ; 1: hi*hi
    int split_k; float* partial;                      // split_k > 1: raw fp32 partial sums [split_k][M][N], no epilogue
    // EPI_GENERIC: v = act(acc + bias) + residual -> out_f32 and/or bf16 planes
    const float* bias;
    const float* residual; int ldr;
    float* out_f32; int ldo;
    bf16* out_hi; bf16* out_lo; int ldp;
    int act;
    // EPI_QKV_ROPE (SmolLM2 attention input, transformers modeling_llama.py:262-270,138-168):
    // columns [0,576) q, [576,768) k, [768,960) v.  q/k head dims are stored pair-interleaved (2i <- i, 2i+1 <- i+32;
    // the weight rows are permuted at pack time), so rotate-half pairs sit in adjacent accumulator columns.
    float* q_out;                                     // [M,576] roped queries
    void* k_cache; void* v_cache;                     // this layer: [B][3][t_max][64], float or bf16
    const float* rope_cos; const float* rope_sin;     // [kMaxPos][32]
    int rows_per_seq;                                 // m = b*rows_per_seq + s
    int pos_base; const int* d_pos;                   // position = pos_base + (d_pos ? *d_pos : 0) + s
    int t_max; int kv_bf16;
};

// (m, n) and (m, n+1); n is even; caller guarantees m < M and n < N.  r0/r1: residual values (EPI_GENERIC), already
// loaded by the caller so that engines can batch those loads ahead of the dependent stores.
template <int EPI>
__device__ __forceinline__ void epilogue_pair_r(const GemmArgs& g, int m, int n, float v0, float v1, float r0, float r1) {
    if (EPI == EPI_GENERIC) {
        const bool has1 = (n + 1 < g.N);
        if (g.bias) { v0 += __ldg(g.bias + n); if (has1) v1 += __ldg(g.bias + n + 1); }
        if (g.act == ACT_GELU) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
        else if (g.act == ACT_SIGMOID) { v0 = sigmoidf_(v0); v1 = sigmoidf_(v1); }
        v0 += r0; v1 += r1;
        if (g.out_f32) {
            float* o = g.out_f32 + (size_t)m * g.ldo + n;
            if (has1 && !(g.ldo & 1)) *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
            else { o[0] = v0; if (has1) o[1] = v1; }
        }
        if (g.out_hi) {
            size_t idx = (size_t)m * g.ldp + n;
            if (has1 && !(g.ldp & 1)) store_planes2(g.out_hi, g.out_lo, idx, v0, v1);
            else { store_planes1(g.out_hi, g.out_lo, idx, v0); if (has1) store_planes1(g.out_hi, g.out_lo, idx + 1, v1); }
        }
    } else if (EPI == EPI_SWIGLU) {
        // weight rows are interleaved (gate_j, up_j): down_proj(silu(gate) * up), modeling_llama.py:182-184
        float h = siluf_(v0) * v1;
        store_planes1(g.out_hi, g.out_lo, (size_t)m * g.ldp + (n >> 1), h);
    } else {
        const int b = m / g.rows_per_seq;
        const int s = m - b * g.rows_per_seq;
        const int pos = g.pos_base + (g.d_pos ? *g.d_pos : 0) + s;
        if (n < kHidden + kKvHeads * kHeadDim) {
            const int i = (n & (kHeadDim - 1)) >> 1;
            const float c = __ldg(g.rope_cos + pos * 32 + i), sn = __ldg(g.rope_sin + pos * 32 + i);
            const float r0 = v0 * c - v1 * sn;
            const float r1 = v1 * c + v0 * sn;
            v0 = r0; v1 = r1;
        }
        if (n < kHidden) {
            *reinterpret_cast<float2*>(g.q_out + (size_t)m * kHidden + n) = make_float2(v0, v1);
        } else {
            const int nn = n - kHidden;
            const bool is_v = nn >= kKvHeads * kHeadDim;
            const int c2 = is_v ? nn - kKvHeads * kHeadDim : nn;
            const int kvh = c2 >> 6, dd = c2 & 63;
            const size_t off = (((size_t)b * kKvHeads + kvh) * g.t_max + pos) * kHeadDim + dd;
            void* base = is_v ? g.v_cache : g.k_cache;
            if (g.kv_bf16) kv_store2(reinterpret_cast<bf16*>(base) + off, v0, v1);
            else kv_store2(reinterpret_cast<float*>(base) + off, v0, v1);
        }
    }
}

template <int EPI>
__device__ __forceinline__ void load_residual_pair(const GemmArgs& g, int m, int n, float& r0, float& r1) {
    r0 = 0.f; r1 = 0.f;
    if (EPI == EPI_GENERIC && g.residual) {
        const float* r = g.residual + (size_t)m * g.ldr + n;
        r0 = r[0];
        if (n + 1 < g.N) r1 = r[1];
    }
}

template <int EPI>
__device__ __forceinline__ void epilogue_pair(const GemmArgs& g, int m, int n, float v0, float v1) {
    float r0, r1;
    load_residual_pair<EPI>(g, m, n, r0, r1);
    epilogue_pair_r<EPI>(g, m, n, v0, v1, r0, r1);
}

// Engine entry points (gemm_mma.cu, gemm_umma.cu)
cudaError_t launch_gemm_mma(const GemmArgs& g, int epi, cudaStream_t st);

}  // namespace mb
