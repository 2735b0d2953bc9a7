// Shared device helpers for the mellow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- model constants
// reference mellow/config/v0.yaml, mellow/model/config.py:4-9, mellow/model/htsat.py:599-606,
// SmolLM2-135M
constexpr int kClipSamples = 320000;
constexpr int kNfft = 1024;
constexpr int kHop = 320;
constexpr int kBins = 513;
constexpr int kMels = 64;
constexpr int kFrames = 1001;
constexpr int kStretch = 1024;
constexpr int kImg = 256;
constexpr int kGrid0 = 64;          // 64x64 patches
constexpr int kEmbed = 96;
constexpr int kWin = 8;
constexpr int kWinTok = 64;
constexpr int kClasses = 527;
constexpr int kClassesPad = 544;    // K of the c2l GEMM padded to a multiple of 32
constexpr int kEncOut = 768;
constexpr int kProj = 576;
constexpr int kAudioRows = 33;      // 1 latent + 32 unique frame rows per clip
constexpr int kAudioSlots = 129;
constexpr int kTextLen = 129;
constexpr int kPrefix = 389;
constexpr int kVocab = 49152;
constexpr int kHidden = 576;
constexpr int kLayers = 30;
constexpr int kHeads = 9;
constexpr int kKvHeads = 3;
constexpr int kHeadDim = 64;
constexpr int kQkvDim = 960;        // 576 q + 192 k + 192 v
constexpr int kInter = 1536;
constexpr int kMaxPos = 8192;       // rope table rows = SmolLM2 max_position_embeddings (389 + max_new <= 8192)

// Precision policy (mb_create): how the tensor-core GEMM operands are represented.
//   kPolicySplit: every fp32 operand x is carried as two bf16 planes hi=bf16(x), lo=bf16(x-hi) and each GEMM
//                 issues hi*hi + hi*lo + lo*hi with fp32 accumulation (error ~2^-16 relative; greedy ids match the
//                 fp32 reference); KV cache fp32.
//   kPolicyFast:  single bf16 plane, one MMA pass, KV cache bf16 (logits within the stated bf16 tolerance).
//   kPolicySplit24: kPolicySplit with the KV cache rounded to 24 bits (kKvF24 below): the same 2^-17 the GEMM operands
//                 carry, 25 % fewer bytes for the HBM-bound decode attention.
constexpr int kPolicySplit = 0;
constexpr int kPolicyFast = 1;
constexpr int kPolicySplit24 = 2;

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// hi/lo bf16 split of an fp32 value
__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ void store_planes2(bf16* hi, bf16* lo, size_t idx, float a, float b) {
    bf16 ah, al, bh, bl;
    split_bf16(a, ah, al);
    split_bf16(b, bh, bl);
    __nv_bfloat162 h2, l2;
    h2.x = ah; h2.y = bh;
    l2.x = al; l2.y = bl;
    *reinterpret_cast<__nv_bfloat162*>(hi + idx) = h2;
    if (lo) *reinterpret_cast<__nv_bfloat162*>(lo + idx) = l2;
}
__device__ __forceinline__ void store_planes1(bf16* hi, bf16* lo, size_t idx, float a) {
    bf16 ah, al;
    split_bf16(a, ah, al);
    hi[idx] = ah;
    if (lo) lo[idx] = al;
}

// Exact (erf) GELU of the reference (`approximate='none'`, htsat.py:130-136).  erf through Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, the size of erff's own rounding) costs ~12 instructions instead of erff's ~30: the fc1 GEMMs of
// Swin stages 0-1 (K = 96 / 192) are bound by this epilogue, not by the tensor pipe (ncu: 505 M warp instructions,
// 14 % tensor-active for 232 GFLOP).
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = __expf(-ax * ax);
    const float y = fmaf(-p * t, e, 1.0f);
    return copysignf(y, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Programmatic dependent launch (PDL): every kernel of the library triggers its dependents at entry and waits for its
// predecessor right before the first dependent global access, so launch latency, prologues and weight prefetches of
// kernel N+1 overlap the tail of kernel N (also inside the captured decode graph).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Optional in-kernel timeline (mb_set_trace).  The first and the last CTA of a launch each claim one 16-slot record with
// a single atomic at kernel entry (off the critical path: before the programmatic-dependency wait) and afterwards
// stamp %globaltimer into that record with plain stores, so the stamps do not stall the threads they measure.
// A null buffer (the default) costs one predicate per stamp.
struct TraceEvent { unsigned long long t; unsigned int id; unsigned int sm; };
struct TraceBuf { unsigned int n; unsigned int cap; TraceEvent ev[1]; };   // n, cap count 16-event records
enum { TR_ENTRY = 0, TR_WAITED = 1, TR_EXIT = 2, TR_EXIT_LAST = 3 };
constexpr unsigned kTraceNone = 0xffffffffu;
__device__ __forceinline__ void trace_put(TraceBuf* tb, unsigned rec, unsigned id, unsigned phase) {
    if (tb == nullptr || rec == kTraceNone) return;
    unsigned long long t;
    unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    TraceEvent* e = &tb->ev[(size_t)rec * 16 + phase];
    e->t = t; e->id = (id << 4) | phase; e->sm = sm;
}
__device__ __forceinline__ bool first_cta() { return (blockIdx.x | blockIdx.y | blockIdx.z) == 0; }
__device__ __forceinline__ bool last_cta() {
    return blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1;
}
// called by one thread of the first / last CTA at kernel entry; returns the record index (kTraceNone when not traced)
__device__ __forceinline__ unsigned trace_open(TraceBuf* tb, unsigned id) {
    if (tb == nullptr || !(first_cta() || last_cta())) return kTraceNone;
    const unsigned rec = atomicAdd(&tb->n, 1u);
    if (rec >= tb->cap) return kTraceNone;
    trace_put(tb, rec, id, first_cta() ? TR_ENTRY : 15);
    return rec;
}
// exit stamps: phase TR_EXIT from the first CTA, TR_EXIT_LAST from the last CTA (both when the grid has one CTA)
__device__ __forceinline__ void trace_close(TraceBuf* tb, unsigned rec, unsigned id) {
    if (tb == nullptr || rec == kTraceNone) return;
    if (first_cta()) trace_put(tb, rec, id, TR_EXIT);
    if (last_cta()) trace_put(tb, rec, id, TR_EXIT_LAST);
}

// KV-cache formats (GemmArgs::kv_fmt, DecodeAttnArgs::kv_fmt).
//   kKvF32 : float rows, 256 B per (position, kv head)
//   kKvBf16: bf16 rows, 128 B (policy fast)
//   kKvF24 : the value rounded to 24 bits (sign, 8 exponent, 15 mantissa bits -- relative error 2^-17, the precision
//            the split GEMM operands carry) stored as a bf16-like upper half and one extra mantissa byte, planar inside
//            the row: bytes [0,128) = 64 x u16 (bits 31..16), bytes [128,192) = 64 x u8 (bits 15..8); 192 B per row.
//            Decode attention is HBM-bound on exactly these rows, so the format is worth 25 % of its time.
enum { kKvF32 = 0, kKvBf16 = 1, kKvF24 = 2 };
struct kv24 { unsigned char b[3]; };                               // tag type of the kKvF24 kernels (never dereferenced)
__host__ __device__ constexpr int kv_row_bytes(int fmt) { return fmt == kKvF32 ? 256 : (fmt == kKvBf16 ? 128 : 192); }
template <typename T> struct KvRowBytes { static constexpr int value = kHeadDim * (int)sizeof(T); };
template <> struct KvRowBytes<kv24> { static constexpr int value = 192; };
// round-to-nearest-even to 24 bits; returns the fp32 bit pattern with the low byte cleared
__device__ __forceinline__ uint32_t f24_bits(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u + 0x7Fu + ((u >> 8) & 1u)) & 0xFFFFFF00u;
}
__device__ __forceinline__ float f24_round(float x) { return __uint_as_float(f24_bits(x)); }
// Unpack: `hw` holds the upper halves of two consecutive values (even in the low 16 bits), `lw` their mantissa bytes at
// byte 2*sel (even) and 2*sel+1 (odd).  One PRMT per value; the lowest byte of the result repeats the mantissa byte
// instead of being zero, i.e. the value is exact to 2^-25 relative (256x below the 24-bit rounding).
__device__ __forceinline__ float f24_unpack_even(uint32_t hw, uint32_t lw, int sel) {
    return __uint_as_float(__byte_perm(hw, lw, sel ? 0x1066u : 0x1044u));
}
__device__ __forceinline__ float f24_unpack_odd(uint32_t hw, uint32_t lw, int sel) {
    return __uint_as_float(__byte_perm(hw, lw, sel ? 0x3277u : 0x3255u));
}
// asks the L2 to fetch [p, p + bytes) (p 16-byte aligned, bytes a multiple of 16); no destination, no completion
__device__ __forceinline__ void l2_prefetch(const void* p, int bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// KV-cache element types
__device__ __forceinline__ float kv_load(const float* p) { return *p; }
__device__ __forceinline__ float kv_load(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void kv_cast(float x, float& o) { o = x; }
__device__ __forceinline__ void kv_cast(float x, bf16& o) { o = __float2bfloat16_rn(x); }
__device__ __forceinline__ void kv_store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void kv_store2(bf16* p, float a, float b) {
    __nv_bfloat162 v;
    v.x = __float2bfloat16_rn(a);
    v.y = __float2bfloat16_rn(b);
    *reinterpret_cast<__nv_bfloat162*>(p) = v;
}

// Host: per-device caches of one-time launch setup (cudaFuncSetAttribute is per device; a process may drive several
// GPUs, one handle each).
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}
inline int sm_count() {
    static int n[kMaxDevices] = {};
    const int dev = current_device();
    if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    return n[dev];
}
// once per (kernel, device): opt in to more than 48 KB of dynamic shared memory
template <typename K>
inline cudaError_t ensure_smem(K kern, size_t bytes, bool* configured) {
    const int dev = current_device();
    if (configured[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) configured[dev] = true;
    return e;
}

// Host: launch with the programmatic-stream-serialization attribute (disabled by MB_NO_PDL=1).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace mb
