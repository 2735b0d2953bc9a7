// Front end: waveform -> log-mel -> BatchNorm (reference mellow/model/htsat.py:864-870, torchlibrosa
// Spectrogram + LogmelFilterBank constructed at :647-653), and time-stretch + fold + PatchEmbed (:830-845, :108-116).
//
// The reference evaluates the STFT as two 1024-tap strided convolutions against the Hann-windowed DFT basis stored
// in the checkpoint.  Here each frame is one in-shared-memory real FFT (valid because the packer verifies that the
// checkpoint basis *is* window[n] * exp(-2 pi i n k / 1024), see mellow_b200/weights.py), and the power spectrum
// never leaves the SM: mel projection, 10*log10 and the BatchNorm affine are applied before the single store.
// Algorithmic HBM traffic: 1.28 MB read + 0.256 MB written per clip.
#include "kernels.cuh"

namespace mb {

namespace {

// Layout of the kernel (ncu on the radix-2 version of round 1: 444 M warp instructions per 128 clips, 27 % issue-active, a
// __syncwarp and a pass over shared memory per radix-2 stage, a serial L1-latency-bound mel loop):
//   * one warp per frame, kFramesPerWarp frames per warp, no block barrier after the table set-up;
//   * the frame is read straight from global memory (consecutive lanes = consecutive float2; frames overlap 3.2x and hit
//     L1 / L2), windowed and packed as 512 complex points;
//   * 512-point complex FFT as THREE radix-8 decimation-in-frequency passes (sub-transform lengths 512, 64, 8): a lane
//     runs two 8-point butterflies per pass entirely in registers (8 independent loads in flight, twiddles w^2..w^7 by
//     multiplication from one table entry), so the data crosses shared memory 3 times instead of 9; the index padding
//     i + i/8 + i/64 keeps every pass and the digit-reversed read-out at most 2-way bank conflicted;
//   * real-FFT untangle on PAIRS (k, 512 - k): both power bins come from the same two loads and one twiddle;
//   * mel projection from a compact table of the non-zero filter weights (built once per CTA in shared memory, rows padded
//     to 4 for 16-byte loads); lane l owns mel l and mel 63 - l.
constexpr int kWarpsPerCta = 8;
constexpr int kFramesPerWarp = 4;
constexpr int kFramesPerCta = kWarpsPerCta * kFramesPerWarp;
constexpr int kZLen = 512 + 64 + 8;                           // padded complex points per warp
constexpr int kMelCap = 2048;                                 // capacity of the compact mel table (floats)

__device__ __forceinline__ int zpad(int i) { return i + (i >> 3) + (i >> 6); }

struct LogmelSmem {
    float2 z[kWarpsPerCta][kZLen];
    float power[kWarpsPerCta][520];
    float2 tw[512];
    float mel_tab[kMelCap];
    int mel_off[kMels + 1];
    int mel_lo[kMels], mel_hi[kMels];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }                 // a * (-i)

// y_q = sum_m x_m exp(-2 pi i m q / 8), in place
__device__ __forceinline__ void dft8(float2* x) {
    constexpr float c = 0.70710678118654752440f;
    const float2 a0 = cadd(x[0], x[4]), a4 = csub(x[0], x[4]);
    const float2 a1 = cadd(x[1], x[5]), t5 = csub(x[1], x[5]);
    const float2 a2 = cadd(x[2], x[6]), t6 = csub(x[2], x[6]);
    const float2 a3 = cadd(x[3], x[7]), t7 = csub(x[3], x[7]);
    const float2 a5 = make_float2(c * (t5.x + t5.y), c * (t5.y - t5.x));       // * (1 - i) / sqrt 2
    const float2 a6 = mul_mi(t6);                                                // * (-i)
    const float2 a7 = make_float2(c * (t7.y - t7.x), -c * (t7.x + t7.y));      // * (-1 - i) / sqrt 2
    const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = mul_mi(csub(a1, a3));
    const float2 b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = mul_mi(csub(a5, a7));
    x[0] = cadd(b0, b1); x[4] = csub(b0, b1);
    x[2] = cadd(b2, b3); x[6] = csub(b2, b3);
    x[1] = cadd(b4, b5); x[5] = csub(b4, b5);
    x[3] = cadd(b6, b7); x[7] = csub(b6, b7);
}

// one radix-8 DIF pass over sub-transforms of length L (stride = L / 8 between the 8 inputs of a butterfly); the outputs
// are multiplied by w_L^(j q) except in the last pass (L = 8, j = 0)
template <int L>
__device__ __forceinline__ void fft_pass(float2* z, const float2* tw, int lane) {
    constexpr int S = L / 8;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int t = lane + 32 * u;                                // butterfly 0..63
        const int j = t & (S - 1);
        const int base = (t / S) * L + j;
        float2 x[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = z[zpad(base + m * S)];
        dft8(x);
        if (L > 8) {
            const float2 w1 = tw[2 * j * (512 / L)];                // w_L^j = w_1024^(2 j 512 / L)
            const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
            const float2 w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w4, w3);
            x[1] = cmul(x[1], w1); x[2] = cmul(x[2], w2); x[3] = cmul(x[3], w3); x[4] = cmul(x[4], w4);
            x[5] = cmul(x[5], w5); x[6] = cmul(x[6], w6); x[7] = cmul(x[7], w7);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) z[zpad(base + q * S)] = x[q];
    }
    __syncwarp();
}

// position of Z[k] after the three DIF passes: the base-8 digits of k reversed
__device__ __forceinline__ int zrev(int k) { return zpad(((k & 7) << 6) | (k & 0x38) | (k >> 6)); }

__global__ void __launch_bounds__(kWarpsPerCta * 32, 3) logmel_kernel(const float* __restrict__ wave, FrontendW w,
                                                                    float* __restrict__ logmel_out,
                                                                    float* __restrict__ bn_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LogmelSmem& sm = *reinterpret_cast<LogmelSmem*>(smem_raw);
    const int clip = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* x = wave + (size_t)clip * kClipSamples;
    pdl_trigger();

    // tables (weights only: before the dependency wait): twiddles, mel filter extents, compact non-zero weights
    for (int j = tid; j < 512; j += blockDim.x) sm.tw[j] = reinterpret_cast<const float2*>(w.twiddle)[j];
    if (tid < kMels) { sm.mel_lo[tid] = __ldg(w.mel_lo + tid); sm.mel_hi[tid] = __ldg(w.mel_hi + tid); }
    __syncthreads();
    if (warp == 0) {                                                // exclusive prefix sum of the row lengths, padded to 4
        const int l0 = (sm.mel_hi[lane] - sm.mel_lo[lane] + 3) & ~3, l1 = (sm.mel_hi[lane + 32] - sm.mel_lo[lane + 32] + 3) & ~3;
        int s0 = l0, s1 = l1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, s0, o), b = __shfl_up_sync(0xffffffffu, s1, o);
            if (lane >= o) { s0 += a; s1 += b; }
        }
        const int tot0 = __shfl_sync(0xffffffffu, s0, 31);
        sm.mel_off[lane] = s0 - l0;
        sm.mel_off[lane + 32] = tot0 + s1 - l1;
        if (lane == 31) sm.mel_off[kMels] = tot0 + s1;
    }
    __syncthreads();
    const bool use_tab = sm.mel_off[kMels] <= kMelCap;              // otherwise the weights are read from the dense matrix
    if (use_tab) {
        for (int m = warp; m < kMels; m += kWarpsPerCta) {
            const int lo = sm.mel_lo[m], n = sm.mel_hi[m] - lo, o = sm.mel_off[m], np = sm.mel_off[m + 1] - o;
            for (int i = lane; i < np; i += 32) sm.mel_tab[o + i] = i < n ? __ldg(w.melW + (size_t)(lo + i) * kMels + m) : 0.f;
        }
    }
    __syncthreads();
    pdl_wait();

    float2* z = sm.z[warp];
    float* pw = sm.power[warp];
    if (lane < 8) pw[512 + lane] = 0.f;                             // the padded 4-wide mel loads may read past bin 512
    for (int fi = 0; fi < kFramesPerWarp; ++fi) {
        const int frame = blockIdx.x * kFramesPerCta + warp * kFramesPerWarp + fi;
        if (frame >= kFrames) break;                                // warp-uniform
        // center=True, pad_mode='reflect': padded[i] = x[reflect(i - 512)]; windowed frame packed as 512 complex points
        const int s0 = frame * kHop - kNfft / 2;
        if (s0 >= 0 && s0 + kNfft <= kClipSamples) {
#pragma unroll 4
            for (int n = lane; n < 512; n += 32) {
                const float2 xs = __ldg(reinterpret_cast<const float2*>(x + s0) + n);
                const float2 ws = __ldg(reinterpret_cast<const float2*>(w.window) + n);
                z[zpad(n)] = make_float2(xs.x * ws.x, xs.y * ws.y);
            }
        } else {
            for (int n = lane; n < 512; n += 32) {
                int i0 = s0 + 2 * n, i1 = i0 + 1;
                if (i0 < 0) i0 = -i0;
                if (i0 >= kClipSamples) i0 = 2 * (kClipSamples - 1) - i0;
                if (i1 < 0) i1 = -i1;
                if (i1 >= kClipSamples) i1 = 2 * (kClipSamples - 1) - i1;
                z[zpad(n)] = make_float2(__ldg(x + i0) * __ldg(w.window + 2 * n), __ldg(x + i1) * __ldg(w.window + 2 * n + 1));
            }
        }
        __syncwarp();
        fft_pass<512>(z, sm.tw, lane);
        fft_pass<64>(z, sm.tw, lane);
        fft_pass<8>(z, sm.tw, lane);
        // real-FFT untangle: X[k] = E + W^k O, X[512-k] = conj(E - W^k O) with E = (Z[k] + conj Z[512-k]) / 2,
        // O = (Z[k] - conj Z[512-k]) / 2i, W = exp(-2 pi i / 1024); power = |X|^2
#pragma unroll 2
        for (int k = lane + 1; k < 256; k += 32) {
            const float2 a = z[zrev(k)], b = z[zrev(512 - k)];
            const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
            const float2 o = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
            const float2 t = cmul(o, sm.tw[k]);
            const float pr = e.x + t.x, pi = e.y + t.y, qr = e.x - t.x, qi = e.y - t.y;
            pw[k] = pr * pr + pi * pi;
            pw[512 - k] = qr * qr + qi * qi;
        }
        if (lane == 0) {
            const float2 z0 = z[zrev(0)], zm = z[zrev(256)];
            pw[0] = (z0.x + z0.y) * (z0.x + z0.y);
            pw[512] = (z0.x - z0.y) * (z0.x - z0.y);
            pw[256] = zm.x * zm.x + zm.y * zm.y;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int m = r == 0 ? lane : kMels - 1 - lane;
            const int lo = sm.mel_lo[m], hi = sm.mel_hi[m];
            float acc = 0.f;
            if (use_tab) {
                const float* tab = sm.mel_tab + sm.mel_off[m];
                for (int k = lo; k < hi; k += 4) {                  // rows are zero-padded to 4, power is readable up to 519
                    const float4 wv = *reinterpret_cast<const float4*>(tab + (k - lo));
                    acc += pw[k] * wv.x; acc += pw[k + 1] * wv.y; acc += pw[k + 2] * wv.z; acc += pw[k + 3] * wv.w;
                }
            } else {
                for (int k = lo; k < hi; ++k) acc += pw[k] * __ldg(w.melW + k * kMels + m);
            }
            const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
            const size_t o = ((size_t)clip * kFrames + frame) * kMels + m;
            if (logmel_out) logmel_out[o] = db;
            if (bn_out) bn_out[o] = db * __ldg(w.bn_scale + m) + __ldg(w.bn_shift + m);
        }
        __syncwarp();                                               // the next frame overwrites z and the power row
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Patch embedding fused with the bicubic time-stretch (1001 -> 1024 frames, A = -0.75, align_corners) and the
// fold into the 256x256 image: img[c*64+f, t] = stretched[c*256+t, f].  One CTA per (clip, patch row), one warp per
// patch; lanes 0..15 build the 4x4 patch pixels, then every lane produces 3 of the 96 channels and the LayerNorm is
// done with warp shuffles.
__device__ __forceinline__ float cubic1(float x) { return ((1.25f * x - 2.25f) * x) * x + 1.0f; }                  // |x|<=1
__device__ __forceinline__ float cubic2(float x) { return ((-0.75f * x + 3.75f) * x - 6.0f) * x + 3.0f; }          // 1<|x|<2

__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ bn, PatchW w,
                                                          float* __restrict__ x_out) {
    __shared__ float sw[kEmbed * 16];
    __shared__ float sb[kEmbed], sg[kEmbed], sbeta[kEmbed];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    for (int i = tid; i < kEmbed * 16; i += 256) sw[i] = w.w[i];
    for (int i = tid; i < kEmbed; i += 256) { sb[i] = w.b[i]; sg[i] = w.ln_w[i]; sbeta[i] = w.ln_b[i]; }
    __syncthreads();
    pdl_wait();
    const int clip = blockIdx.y, ph = blockIdx.x;
    const float* src = bn + (size_t)clip * kFrames * kMels;
    const float scale = (float)(kFrames - 1) / (float)(kStretch - 1);
    for (int pwi = warp; pwi < kGrid0; pwi += 8) {
        float pix = 0.f;
        if (lane < 16) {
            const int i = lane >> 2, j = lane & 3;
            const int h = 4 * ph + i;
            const int c = h >> 6, f = h & 63;
            const int T = c * kImg + 4 * pwi + j;
            const float s = scale * (float)T;
            const float fl = floorf(s);
            const float t = s - fl;
            const int i0 = (int)fl;
            const float wt[4] = {cubic2(t + 1.0f), cubic1(t), cubic1(1.0f - t), cubic2(2.0f - t)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int r = i0 - 1 + k;
                r = r < 0 ? 0 : (r > kFrames - 1 ? kFrames - 1 : r);
                pix += wt[k] * __ldg(src + r * kMels + f);
            }
        }
        float v[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) v[r] = sb[lane + 32 * r];
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            const float p = __shfl_sync(0xffffffffu, pix, l);
#pragma unroll
            for (int r = 0; r < 3; ++r) v[r] += sw[(lane + 32 * r) * 16 + l] * p;
        }
        const float mean = warp_sum(v[0] + v[1] + v[2]) * (1.0f / kEmbed);
        float sq = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) { const float d = v[r] - mean; sq += d * d; }
        const float rstd = rsqrtf(warp_sum(sq) * (1.0f / kEmbed) + 1e-5f);
        float* o = x_out + ((size_t)clip * (kGrid0 * kGrid0) + ph * kGrid0 + pwi) * kEmbed;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int e = lane + 32 * r;
            o[e] = (v[r] - mean) * rstd * sg[e] + sbeta[e];
        }
    }
}

}  // namespace

cudaError_t launch_logmel(const float* wave, int n_clips, const FrontendW& w, float* logmel_out, float* bn_out,
                          cudaStream_t st) {
    static bool configured[kMaxDevices] = {};
    if (cudaError_t e = ensure_smem(logmel_kernel, sizeof(LogmelSmem), configured); e != cudaSuccess) return e;
    dim3 grid((kFrames + kFramesPerCta - 1) / kFramesPerCta, n_clips);
    return launch_k(logmel_kernel, grid, dim3(kWarpsPerCta * 32), sizeof(LogmelSmem), st, wave, w, logmel_out, bn_out);
}

cudaError_t launch_patch_embed(const float* bn, int n_clips, const PatchW& w, float* x_out, cudaStream_t st) {
    dim3 grid(kGrid0, n_clips);
    return launch_k(patch_embed_kernel, grid, dim3(256), 0, st, bn, w, x_out);
}

}  // namespace mb
