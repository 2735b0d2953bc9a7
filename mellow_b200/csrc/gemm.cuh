// C[M,N] = A[M,K] * W[N,K]^T with fused epilogues; A and W are bf16 "planes" (hi, optional lo), fp32 accumulate.
// This header holds the argument block and the epilogues shared by every GEMM engine in the library.
#pragma once
#include "common.cuh"

namespace mb {

constexpr float kQScaleLog2 = 0.125f * 1.4426950408889634f;   // head_dim^-0.5 * log2(e): attn_umma.cu evaluates exp2(s - m)

enum { EPI_GENERIC = 0, EPI_SWIGLU = 1, EPI_QKV_ROPE = 2, EPI_ARGMAX = 3 };
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SIGMOID = 2 };

struct GemmArgs {
    const bf16* A_hi; const bf16* A_lo; int lda;      // activations [M,K] (lo may be null when passes==1)
    const bf16* W_hi; const bf16* W_lo; int ldw;      // weights [N,K], K contiguous (nn.Linear layout)
    int M, N, K;                                      // K % 32 == 0
    int passes;                                       // 3: hi*hi + hi*lo + lo*hi ; 1: hi*hi
    int split_k; float* partial;                      // split_k > 1: raw fp32 partial sums [split_k][M][N], no epilogue
    TraceBuf* trace; unsigned trace_id;               // optional timeline stamps (common.cuh)
    int epi_sleep;                                    // weight-resident kernel: ns the epilogue warps sleep between polls of the accumulator barrier
    int resident;                                     // decode chain: use the weight-resident kernel (gemm_skinny.cu) when the shape fits
    int bn_hint;                                      // decode-sized GEMMs: N-tile width 16 / 32 (0 = default rule)
    int cta_pairs;                                    // gemm_umma.cu: 1 = cta_group::2 pairs for the 192 / 256-column tiles of large GEMMs
    int epi_rows;                                     // gemm_umma.cu: 1 = keep the row-per-thread global accesses in the plain epilogue (A/B switch)
    // Deferred RMSNorm (LlamaRMSNorm, modeling_llama.py:62-67) across a producer / consumer pair of GEMMs.  RMSNorm(x) * W^T =
    // rstd(x) * ((x * gain) * W^T), and rstd is one scalar per row, so the PRODUCER of x (o_proj / down_proj, EPI_GENERIC
    // with a residual) writes the planes of x * gain (`norm_w`) plus per-row partial sums of x^2 (`ssq_out`, one per
    // 32-column chunk group it owns), and the CONSUMER (gate/up, QKV) multiplies its accumulator rows by
    // rstd = rsqrt(sum(partials) / K + eps) before its own epilogue.  This removes the stand-alone norm kernel and its pass
    // over x; sums run in a fixed order (deterministic).
    const float* norm_w;                              // producer: gain applied to the plane outputs (out_f32 keeps x itself)
    float* ssq_out; int ssq_ld;                       // producer: partial row sums of squares [parts][ssq_ld]
    const float* ssq_in; int ssq_parts;               // consumer: partials written by the producer (same ssq_ld), their count
    // EPI_GENERIC: v = act(acc + bias) + residual -> out_f32 and/or bf16 planes
    const float* bias;
    const float* residual; int ldr;
    float* out_f32; int ldo;
    bf16* out_hi; bf16* out_lo; int ldp;
    int act;
    // EPI_QKV_ROPE (SmolLM2 attention input, transformers modeling_llama.py:262-270,138-168):
    // columns [0,576) q, [576,768) k, [768,960) v.  q/k head dims are stored pair-interleaved (2i <- i, 2i+1 <- i+32;
    // the weight rows are permuted at pack time), so rotate-half pairs sit in adjacent accumulator columns.
    float* q_out;                                     // [M,576] roped queries (may be null when the planes below are written)
    // prefill with the tcgen05 attention kernel (attn_umma.cu): the operands it streams with TMA, as bf16 hi/lo planes
    bf16* qp_hi; bf16* qp_lo;                         // [M,576] roped queries * head_dim^-0.5 * log2(e)
    bf16* kp_hi; bf16* kp_lo;                         // [B][3][rows_per_seq][64] roped keys
    bf16* vt_hi; bf16* vt_lo; int vt_ld;              // [B][3][64][vt_ld] values, TRANSPOSED (keys contiguous)
    void* k_cache; void* v_cache;                     // this layer: [B][3][t_max][64], float or bf16
    const float* rope_cos; const float* rope_sin;     // [kMaxPos][32]
    int rows_per_seq;                                 // m = b*rows_per_seq + s
    int pos_base; const int* d_pos;                   // position = pos_base + (d_pos ? *d_pos : 0) + s
    int t_max; int kv_fmt;
    // EPI_ARGMAX (lm_head fused with the first stage of the sampling step): per row and per 16-column group the
    // (max logit, first arg max) candidate is written instead of the 49152 logits; sample_kernel finishes the scan.
    float* cand_val; int* cand_idx;                   // [M][N/16]
};

// Per-(row, column pair) auxiliary operands an epilogue needs from global memory: the residual pair (EPI_GENERIC) or
// the RoPE (cos, sin) of that row's position (EPI_QKV_ROPE).  Engines load them ahead of time / in batches so the
// epilogue is not a chain of exposed memory latencies.
template <int EPI>
__device__ __forceinline__ void load_aux_pair(const GemmArgs& g, int m, int n, float& r0, float& r1) {
    r0 = 0.f; r1 = 0.f;
    if (EPI == EPI_GENERIC) {
        if (g.residual) {
            const float* r = g.residual + (size_t)m * g.ldr + n;
            r0 = r[0];
            if (n + 1 < g.N) r1 = r[1];
        }
    } else if (EPI == EPI_QKV_ROPE) {
        if (n < kHidden + kKvHeads * kHeadDim) {
            const int b = m / g.rows_per_seq;
            const int pos = g.pos_base + (g.d_pos ? *g.d_pos : 0) + (m - b * g.rows_per_seq);
            const int i = (n & (kHeadDim - 1)) >> 1;
            r0 = __ldg(g.rope_cos + pos * 32 + i);
            r1 = __ldg(g.rope_sin + pos * 32 + i);
        }
    }
}

// (m, n) and (m, n+1); n is even; caller guarantees m < M and n < N.  (r0, r1) come from load_aux_pair.
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
    } else if (EPI == EPI_ARGMAX) {
        // only reachable through epilogue_row16 (N % 16 == 0 is enforced by the launcher)
    } else {
        const int b = m / g.rows_per_seq;
        const int s = m - b * g.rows_per_seq;
        const int pos = g.pos_base + (g.d_pos ? *g.d_pos : 0) + s;
        if (n < kHidden + kKvHeads * kHeadDim) {                   // q and k columns: rotate-half pair (v0, v1)
            const float x0 = v0 * r0 - v1 * r1;
            const float x1 = v1 * r0 + v0 * r1;
            v0 = x0; v1 = x1;
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
            if (g.kv_fmt == kKvF24) {
                unsigned char* row = reinterpret_cast<unsigned char*>(base) + (off - dd) / kHeadDim * 192;
                const uint32_t u0 = f24_bits(v0), u1 = f24_bits(v1);
                *reinterpret_cast<uint32_t*>(row + dd * 2) = (u0 >> 16) | (u1 & 0xFFFF0000u);
                *reinterpret_cast<unsigned short*>(row + 128 + dd) = (unsigned short)(((u0 >> 8) & 0xFFu) | (u1 & 0xFF00u));
            } else if (g.kv_fmt == kKvBf16) kv_store2(reinterpret_cast<bf16*>(base) + off, v0, v1);
            else kv_store2(reinterpret_cast<float*>(base) + off, v0, v1);
        }
    }
}

template <int EPI>
__device__ __forceinline__ void epilogue_pair(const GemmArgs& g, int m, int n, float v0, float v1) {
    float r0, r1;
    load_aux_pair<EPI>(g, m, n, r0, r1);
    epilogue_pair_r<EPI>(g, m, n, v0, v1, r0, r1);
}

// ---------------------------------------------------------------------------------------------------------------
// Row-vector epilogue for engines whose threads own whole accumulator rows (tcgen05.ld gives one row per thread):
// 16 consecutive columns n..n+15 of row m (n % 16 == 0) are finished with 16-byte loads / stores.  Few instructions
// per element matter more here than perfect coalescing: the epilogue warps run one per scheduler.
__device__ __forceinline__ void ld4(const float* p, float* d) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
}
__device__ __forceinline__ void ldg4(const float* p, float* d) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
}
__device__ __forceinline__ void st4(float* p, const float* s) {
    *reinterpret_cast<float4*>(p) = make_float4(s[0], s[1], s[2], s[3]);
}
// 8 floats -> 8 bf16 hi (+ 8 bf16 lo) as one 16-byte store per plane
__device__ __forceinline__ void store_planes8(bf16* hi, bf16* lo, size_t idx, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bf16 ah, al, bh, bl;
        split_bf16(v[2 * j], ah, al);
        split_bf16(v[2 * j + 1], bh, bl);
        __nv_bfloat162 h2, l2;
        h2.x = ah; h2.y = bh; l2.x = al; l2.y = bl;
        h[j] = *reinterpret_cast<uint32_t*>(&h2);
        l[j] = *reinterpret_cast<uint32_t*>(&l2);
    }
    *reinterpret_cast<uint4*>(hi + idx) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + idx) = make_uint4(l[0], l[1], l[2], l[3]);
}

// EPI_QKV_ROPE on 16 columns, in two steps so that an engine can fetch the RoPE factors (a dependent chain of two
// global loads: the step counter, then the table row) while its MMAs are still running.
__device__ __forceinline__ void qkv_rope_load(const GemmArgs& g, int m, int n, float* c, float* sn) {
    if (n < kHidden + kKvHeads * kHeadDim) {                       // q and k columns only
        const int b = m / g.rows_per_seq;
        const int pos = g.pos_base + (g.d_pos ? *g.d_pos : 0) + (m - b * g.rows_per_seq);
        const int i0 = (n & (kHeadDim - 1)) >> 1;                  // multiple of 8
        ldg4(g.rope_cos + pos * 32 + i0, c); ldg4(g.rope_cos + pos * 32 + i0 + 4, c + 4);
        ldg4(g.rope_sin + pos * 32 + i0, sn); ldg4(g.rope_sin + pos * 32 + i0 + 4, sn + 4);
    }
}
__device__ __forceinline__ void epilogue_row16_qkv(const GemmArgs& g, int m, int n, float* v, const float* c, const float* sn) {
    const int b = m / g.rows_per_seq;
    const int s = m - b * g.rows_per_seq;
    const int pos = g.pos_base + (g.d_pos ? *g.d_pos : 0) + s;
    if (n < kHidden + kKvHeads * kHeadDim) {                       // 8 rotate-half pairs of one head
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x0 = v[2 * j] * c[j] - v[2 * j + 1] * sn[j];
            const float x1 = v[2 * j + 1] * c[j] + v[2 * j] * sn[j];
            v[2 * j] = x0; v[2 * j + 1] = x1;
        }
    }
    if (n < kHidden) {
        if (g.q_out) {
            float* o = g.q_out + (size_t)m * kHidden + n;
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(o + j, v + j);
        }
        if (g.qp_hi) {
            float qs[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) qs[j] = v[j] * kQScaleLog2;   // scores come out in log2 units (softmax uses exp2)
            store_planes8(g.qp_hi, g.qp_lo, (size_t)m * kHidden + n, qs);
            store_planes8(g.qp_hi, g.qp_lo, (size_t)m * kHidden + n + 8, qs + 8);
        }
    } else {
        const int nn = n - kHidden;
        const bool is_v = nn >= kKvHeads * kHeadDim;
        const int c2 = is_v ? nn - kKvHeads * kHeadDim : nn;
        const int kvh = c2 >> 6, dd = c2 & 63;
        if (g.kp_hi) {
            if (!is_v) {
                const size_t o = (((size_t)b * kKvHeads + kvh) * g.rows_per_seq + s) * kHeadDim + dd;
                store_planes8(g.kp_hi, g.kp_lo, o, v);
                store_planes8(g.kp_hi, g.kp_lo, o + 8, v + 8);
            } else {
                // transposed: lanes of a warp hold consecutive rows = consecutive keys, so each of these 2-byte stores
                // is one 64-byte run per warp
                const size_t o = (((size_t)b * kKvHeads + kvh) * kHeadDim + dd) * g.vt_ld + s;
#pragma unroll
                for (int j = 0; j < 16; ++j) store_planes1(g.vt_hi, g.vt_lo, o + (size_t)j * g.vt_ld, v[j]);
            }
        }
        const size_t off = (((size_t)b * kKvHeads + kvh) * g.t_max + pos) * kHeadDim + dd;
        void* base = is_v ? g.v_cache : g.k_cache;
        if (g.kv_fmt == kKvF24) {                                    // 16 values: 32 B of upper halves + 16 B of mantissa bytes
            unsigned char* row = reinterpret_cast<unsigned char*>(base) + (off - dd) / kHeadDim * 192;
            uint32_t hw[8], lw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t u0 = f24_bits(v[2 * j]), u1 = f24_bits(v[2 * j + 1]);
                hw[j] = (u0 >> 16) | (u1 & 0xFFFF0000u);
                lw[j >> 1] |= (((u0 >> 8) & 0xFFu) | (u1 & 0xFF00u)) << (16 * (j & 1));
            }
            *reinterpret_cast<uint4*>(row + dd * 2) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(row + dd * 2 + 16) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
            *reinterpret_cast<uint4*>(row + 128 + dd) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        } else if (g.kv_fmt == kKvBf16) {
            bf16* o = reinterpret_cast<bf16*>(base) + off;
            store_planes8(o, nullptr, 0, v);
            store_planes8(o, nullptr, 8, v + 8);
        } else {
            float* o = reinterpret_cast<float*>(base) + off;
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(o + j, v + j);
        }
    }
}

template <int EPI>
__device__ __forceinline__ void epilogue_row16(const GemmArgs& g, int m, int n, float* v) {
    bool vec = (n + 16 <= g.N);
    if (EPI == EPI_GENERIC)
        vec = vec && (!g.out_f32 || (g.ldo & 3) == 0) && (!g.residual || (g.ldr & 3) == 0) && (!g.out_hi || (g.ldp & 7) == 0);
    if (!vec) {                                                    // ragged tail / odd leading dimension
#pragma unroll
        for (int j = 0; j < 16; j += 2)
            if (n + j < g.N) epilogue_pair<EPI>(g, m, n + j, v[j], v[j + 1]);
        return;
    }
    if (EPI == EPI_GENERIC) {
        if (g.bias) {
            float bq[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) ldg4(g.bias + n + j, bq + j);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += bq[j];
        }
        if (g.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
        } else if (g.act == ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = sigmoidf_(v[j]);
        }
        if (g.residual) {
            float r[16];
            const float* rp = g.residual + (size_t)m * g.ldr + n;
#pragma unroll
            for (int j = 0; j < 16; j += 4) ld4(rp + j, r + j);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += r[j];
        }
        if (g.out_f32) {
            float* o = g.out_f32 + (size_t)m * g.ldo + n;
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(o + j, v + j);
        }
        if (g.out_hi) {
            const size_t idx = (size_t)m * g.ldp + n;
            if (g.norm_w) {                                            // planes of x * gain (deferred RMSNorm, see GemmArgs)
                float y[16], gw[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) ldg4(g.norm_w + n + j, gw + j);
#pragma unroll
                for (int j = 0; j < 16; ++j) y[j] = gw[j] * v[j];
                store_planes8(g.out_hi, g.out_lo, idx, y);
                store_planes8(g.out_hi, g.out_lo, idx + 8, y + 8);
            } else {
                store_planes8(g.out_hi, g.out_lo, idx, v);
                store_planes8(g.out_hi, g.out_lo, idx + 8, v + 8);
            }
        }
    } else if (EPI == EPI_SWIGLU) {
        float hv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[j] = siluf_(v[2 * j]) * v[2 * j + 1];
        store_planes8(g.out_hi, g.out_lo, (size_t)m * g.ldp + (n >> 1), hv);
    } else if (EPI == EPI_ARGMAX) {
        float best = v[0];
        int bi = 0;
#pragma unroll
        for (int j = 1; j < 16; ++j)
            if (v[j] > best) { best = v[j]; bi = j; }              // strict >: first index wins ties, like torch.argmax
        const size_t c = (size_t)m * (g.N >> 4) + (n >> 4);
        g.cand_val[c] = best;
        g.cand_idx[c] = n + bi;
    } else {
        float c[8], sn[8];
        qkv_rope_load(g, m, n, c, sn);
        epilogue_row16_qkv(g, m, n, v, c, sn);
    }
}

// consumer side of the deferred RMSNorm: rstd of row m from the producer's partial sums of squares.  All loads are
// issued before the first add (a loop with a running sum serialises one L2 round trip per partial: measured +1.6 us on
// the decode QKV GEMM); the sum runs in a fixed order.
constexpr int kMaxSsqParts = 18;
__device__ __forceinline__ float deferred_rstd(const GemmArgs& g, int m) {
    float p[kMaxSsqParts];
#pragma unroll
    for (int i = 0; i < kMaxSsqParts; ++i) p[i] = i < g.ssq_parts ? __ldcg(g.ssq_in + (size_t)i * g.ssq_ld + m) : 0.f;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxSsqParts; ++i) t += p[i];
    return rsqrtf(t * (1.0f / (float)g.K) + 1e-5f);
}

// Engine entry points (gemm_umma.cu, gemm_skinny.cu; gemm_mma.cu = mma.sync cross-check engine of lab builds)
cudaError_t launch_gemm_mma(const GemmArgs& g, int epi, cudaStream_t st);
cudaError_t launch_gemm_skinny(const GemmArgs& g, int epi, int bn, cudaStream_t st);   // gemm_skinny.cu
cudaError_t launch_gemm_tail(const GemmArgs& g, int cs, cudaStream_t st);               // gemm_skinny.cu (cluster split-K tail)
cudaError_t launch_gemm_cluster3(const GemmArgs& g, int epi, cudaStream_t st);          // gemm_skinny.cu (cluster split-K + fused epilogue)

}  // namespace mb
