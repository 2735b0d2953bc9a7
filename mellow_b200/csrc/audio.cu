// Audio ingest on the GPU: the polyphase windowed-sinc resampler of
// torchaudio.transforms.Resample (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99 -- the defaults the reference
// relies on at mellow/wrapper.py:146-148) and the tile-or-crop to exactly 320000 samples (wrapper.py:152-167).
//
// y[f*new + p] = sum_j xpad[f*orig + j] * K[p][j],  xpad = x zero-padded by `width` on the left and `width + orig` on
// the right, K = the [new][klen] filter bank (klen = 2*width + orig) built on the host in float64 exactly like
// torchaudio does.  One CTA computes FR consecutive frames x PH phases: the PH filter rows are staged once in shared
// memory (transposed, conflict-free) and re-used by all FR frames, the input window is staged once and re-used by all
// PH phases, so neither the 0.6 MB filter bank nor the samples are re-read per output.
#include "kernels.cuh"

namespace mb {

namespace {

constexpr int PH = 32;      // phases per CTA
constexpr int FR = 8;       // frames per CTA

__global__ void __launch_bounds__(PH * FR) resample_kernel(const float* __restrict__ x, long long n_in, int orig, int nw,
                                                          const float* __restrict__ kern, int klen, int width,
                                                          float* __restrict__ y, long long n_out) {
    extern __shared__ float sm_audio[];
    float* sk = sm_audio;                         // [klen][PH+1] (transposed filter rows of this phase chunk, padded)
    float* sx = sm_audio + (size_t)klen * (PH + 1);     // [(FR-1)*orig + klen] input window
    const int tid = threadIdx.x;
    const int p0 = blockIdx.y * PH;
    const long long f0 = (long long)blockIdx.x * FR;
    pdl_trigger();
    pdl_wait();
    for (int e = tid; e < klen * PH; e += PH * FR) {
        const int pl = e / klen, j = e - pl * klen;        // coalesced read of row p0+pl
        sk[j * (PH + 1) + pl] = (p0 + pl < nw) ? kern[(size_t)(p0 + pl) * klen + j] : 0.f;
    }
    const int span = (FR - 1) * orig + klen;
    const long long base = f0 * orig - width;              // index into x of xpad[f0*orig]
    for (int e = tid; e < span; e += PH * FR) {
        const long long i = base + e;
        sx[e] = (i >= 0 && i < n_in) ? x[i] : 0.f;
    }
    __syncthreads();
    const int pl = tid % PH, fl = tid / PH;
    const long long o = (f0 + fl) * nw + p0 + pl;
    if (p0 + pl >= nw || o >= n_out) return;
    const float* xs = sx + fl * orig;
    float acc = 0.f;
    for (int j = 0; j < klen; ++j) acc += xs[j] * sk[j * (PH + 1) + pl];
    y[o] = acc;
}

// out[i] = src[(start + i) % total]: tile (start = 0, total < 320000 wraps) or crop (start + 320000 <= total)
__global__ void __launch_bounds__(256) fit_kernel(const float* __restrict__ src, long long total, long long start,
                                                  float* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (i < kClipSamples) out[i] = src[(start + i) % total];
}

}  // namespace

cudaError_t launch_resample(const float* x, long long n_in, int orig, int nw, const float* kern, int klen, int width,
                            float* y, long long n_out, cudaStream_t st) {
    const size_t smem = ((size_t)klen * (PH + 1) + (size_t)(FR - 1) * orig + klen) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static size_t configured[kMaxDevices] = {};
    const int dev = current_device();
    if (smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    const long long frames = (n_out + nw - 1) / nw;
    dim3 grid((unsigned)((frames + FR - 1) / FR), (unsigned)((nw + PH - 1) / PH));
    return launch_k(resample_kernel, grid, dim3(PH * FR), smem, st, x, n_in, orig, nw, kern, klen, width, y, n_out);
}

cudaError_t launch_fit(const float* src, long long total, long long start, float* out, cudaStream_t st) {
    return launch_k(fit_kernel, dim3((kClipSamples + 255) / 256), dim3(256), 0, st, src, total, start, out);
}

}  // namespace mb
