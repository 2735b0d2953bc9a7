// HTSAT (Swin) encoder glue kernels: row normalisation feeding the GEMM operand planes and the TSCAM tail gathers
// (the shifted-window attention lives in attn_mma.cu).  Reference: mellow/model/htsat.py:414-455 (block), :301-332 (window attention),
// :478-499 (patch merging), :742-796 (tail), mellow/model/mellow.py:48-52 (projection).
#include "kernels.cuh"

namespace mb {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// One warp per row: LayerNorm / RMSNorm (fp32 statistics) -> bf16 hi/lo planes (GEMM A operand) and/or fp32.
// NORM_LN_MERGE gathers the 2x2 neighbourhood of PatchMerging (htsat.py:489-493) on the fly.
template <int C, int KIND>
__global__ void __launch_bounds__(256) norm_kernel(const NormArgs a) {
    constexpr int PER = C / 32;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    pdl_trigger();
    pdl_wait();
    if (row >= a.rows) return;
    float v[PER];
    if (KIND == NORM_LN_MERGE) {
        constexpr int CI = C / 4;
        const int r2 = a.res >> 1;
        const int n = row / (r2 * r2);
        const int rem = row - n * r2 * r2;
        const int y2 = rem / r2, x2 = rem - y2 * r2;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = lane + 32 * j;
            const int qd = i / CI, ci = i - qd * CI;
            const int dy = qd & 1, dx = qd >> 1;          // x0:(0,0) x1:(1,0) x2:(0,1) x3:(1,1)
            const size_t tok = (size_t)n * a.res * a.res + (size_t)(2 * y2 + dy) * a.res + (2 * x2 + dx);
            v[j] = a.x[tok * CI + ci];
        }
    } else {
        const float* src = a.x + ((size_t)row * a.row_stride + a.row_off) * C;
#pragma unroll
        for (int j = 0; j < PER; ++j) v[j] = src[lane + 32 * j];
    }
    float mean = 0.f, rstd;
    if (KIND == NORM_RMS) {
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) sq += v[j] * v[j];
        rstd = rsqrtf(warp_sum(sq) * (1.0f / C) + 1e-5f);
    } else {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) s += v[j];
        mean = warp_sum(s) * (1.0f / C);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) { const float d = v[j] - mean; sq += d * d; }
        rstd = rsqrtf(warp_sum(sq) * (1.0f / C) + 1e-5f);
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int i = lane + 32 * j;
        float y;
        if (KIND == NORM_RMS) y = a.w[i] * (v[j] * rstd);          // LlamaRMSNorm: weight * (x * rsqrt(ms + eps))
        else y = (v[j] - mean) * rstd * a.w[i] + a.b[i];
        const size_t o = (size_t)row * C + i;
        if (a.out_f32) a.out_f32[o] = y;
        if (a.out_hi) store_planes1(a.out_hi, a.out_lo, o, y);
    }
}

template <int KIND>
cudaError_t launch_norm_kind(const NormArgs& a, cudaStream_t st) {
    const int grid = (a.rows + 7) / 8;
    switch (a.C) {
        case 96: return launch_k(norm_kernel<96, KIND>, dim3(grid), dim3(256), 0, st, a);
        case 192: return launch_k(norm_kernel<192, KIND>, dim3(grid), dim3(256), 0, st, a);
        case 384: return launch_k(norm_kernel<384, KIND>, dim3(grid), dim3(256), 0, st, a);
        case 576: return launch_k(norm_kernel<576, KIND>, dim3(grid), dim3(256), 0, st, a);
        case 768: return launch_k(norm_kernel<768, KIND>, dim3(grid), dim3(256), 0, st, a);
        case 1536: return launch_k(norm_kernel<1536, KIND>, dim3(grid), dim3(256), 0, st, a);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tail (htsat.py:744-757,774): y = LN(final tokens) [N,64,768], token = (cq*2+fb)*8 + tt, time = cq*8+tt.
//   latent[n,c]  = mean over the 64 tokens (order fb-major then time, like avgpool over flatten(x,2))
//   col[(n,t), (dt*2+fb)*768 + c] = y[n, token(fb, t+dt-1), c]  (zero outside 0..31): im2col of the (2,3) TSCAM conv
__global__ void __launch_bounds__(256) tail_gather_kernel(const float* __restrict__ y, float* __restrict__ latent,
                                                          bf16* __restrict__ col_hi, bf16* __restrict__ col_lo) {
    const int n = blockIdx.y, t = blockIdx.x;                               // t in 0..32 ; t == 32 -> latent
    const float* yn = y + (size_t)n * 64 * kEncOut;
    pdl_trigger();
    pdl_wait();
    if (t == 32) {
        for (int c = threadIdx.x; c < kEncOut; c += 256) {
            float s = 0.f;
            for (int fb = 0; fb < 2; ++fb)
                for (int tm = 0; tm < 32; ++tm) s += yn[(size_t)(((tm >> 3) * 2 + fb) * 8 + (tm & 7)) * kEncOut + c];
            latent[(size_t)n * kEncOut + c] = s * (1.0f / 64.0f);
        }
        return;
    }
    const size_t rowo = ((size_t)n * 32 + t) * (6 * kEncOut);
    for (int e = threadIdx.x; e < 6 * kEncOut; e += 256) {
        const int seg = e / kEncOut, c = e - seg * kEncOut;
        const int dt = seg >> 1, fb = seg & 1;
        const int tm = t + dt - 1;
        float v = 0.f;
        if (tm >= 0 && tm < 32) v = yn[(size_t)(((tm >> 3) * 2 + fb) * 8 + (tm & 7)) * kEncOut + c];
        store_planes1(col_hi, col_lo, rowo + e, v);
    }
}

// rows of the projection input: [latent ; 32 c2l frame rows] per clip (htsat.py:952-954 without the 32x repeat)
__global__ void __launch_bounds__(256) assemble33_kernel(const float* __restrict__ latent,
                                                         const float* __restrict__ frames, bf16* __restrict__ a_hi,
                                                         bf16* __restrict__ a_lo) {
    const int n = blockIdx.y, r = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    const float* src = (r == 0) ? latent + (size_t)n * kEncOut : frames + ((size_t)n * 32 + (r - 1)) * kEncOut;
    const size_t o = ((size_t)n * kAudioRows + r) * kEncOut;
    for (int c = threadIdx.x; c < kEncOut; c += 256) store_planes1(a_hi, a_lo, o + c, src[c]);
}

// the classification heads the path computes but generate() discards (htsat.py:774-796, loss_type
// "clip_bce").  logits [N*32, ld] are the TSCAM conv outputs before the sigmoid:
//   framewise[n,t,c] = sigmoid(logit[n,t,c])                 (the 32 unique rows; the reference repeats each 32x)
//   clipwise[n,c]    = sigmoid(mean_t logit[n,t,c])          (avgpool over time, then sigmoid)
__global__ void __launch_bounds__(256) heads_kernel(const float* __restrict__ logits, int ld, float* __restrict__ clipwise,
                                                    float* __restrict__ framewise) {
    const int n = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    for (int c = threadIdx.x; c < kClasses; c += 256) {
        float s = 0.f;
        for (int t = 0; t < 32; ++t) {
            const float v = logits[((size_t)n * 32 + t) * ld + c];
            s += v;
            if (framewise) framewise[((size_t)n * 32 + t) * kClasses + c] = sigmoidf_(v);
        }
        if (clipwise) clipwise[(size_t)n * kClasses + c] = sigmoidf_(s * (1.0f / 32.0f));
    }
}

__global__ void __launch_bounds__(256) gelu_planes_kernel(const float* __restrict__ x, size_t n, bf16* __restrict__ hi,
                                                          bf16* __restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (i < n) store_planes1(hi, lo, i, gelu_erf(x[i]));
}

__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, size_t n, bf16* __restrict__ hi,
                                                           bf16* __restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (i < n) store_planes1(hi, lo, i, x[i]);
}

}  // namespace

cudaError_t launch_norm(const NormArgs& a, int kind, cudaStream_t st) {
    switch (kind) {
        case NORM_LN: return launch_norm_kind<NORM_LN>(a, st);
        case NORM_RMS: return launch_norm_kind<NORM_RMS>(a, st);
        case NORM_LN_MERGE: return launch_norm_kind<NORM_LN_MERGE>(a, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_tail_gather(const float* y, int n_clips, float* latent, bf16* col_hi, bf16* col_lo,
                               cudaStream_t st) {
    dim3 grid(33, n_clips);
    return launch_k(tail_gather_kernel, grid, dim3(256), 0, st, y, latent, col_hi, col_lo);
}

cudaError_t launch_assemble33(const float* latent, const float* frames, int n_clips, bf16* a_hi, bf16* a_lo,
                              cudaStream_t st) {
    dim3 grid(kAudioRows, n_clips);
    return launch_k(assemble33_kernel, grid, dim3(256), 0, st, latent, frames, a_hi, a_lo);
}

cudaError_t launch_heads(const float* logits, int ld, int n_clips, float* clipwise, float* framewise, cudaStream_t st) {
    return launch_k(heads_kernel, dim3(n_clips), dim3(256), 0, st, logits, ld, clipwise, framewise);
}

cudaError_t launch_gelu_planes(const float* x, size_t n, bf16* hi, bf16* lo, cudaStream_t st) {
    return launch_k(gelu_planes_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x, n, hi, lo);
}

cudaError_t launch_split_planes(const float* x, size_t n, bf16* hi, bf16* lo, cudaStream_t st) {
    return launch_k(split_planes_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x, n, hi, lo);
}

}  // namespace mb
