// Causal prefill attention on the tensor cores (SmolLM2: 9 query heads, 3 kv heads, head_dim 64; reference path:
// transformers modeling_llama.py:276-285 called cache-less from mellow/wrapper.py:217).
//
// Flash-style: CTA = 64 query rows of one (batch, head), 4 warps x 16 rows; K/V are streamed in 64-key tiles from the
// KV cache the QKV GEMM epilogue just wrote, converted to bf16 hi/lo planes in shared memory; S = Q K^T and O += P V
// are mma.sync.m16n8k16 with the same 3-pass operand split as the GEMMs (hi*hi + hi*lo + lo*hi, fp32 accumulate), the
// online softmax runs in fp32 registers.  q/k head dims are pair-interleaved (gemm.cuh), which leaves q.k unchanged.
#include "kernels.cuh"

namespace mb {

namespace {

[[maybe_unused]] constexpr int TQ = 64;
constexpr int TK = 64, LDS = kHeadDim + 8;               // 144-byte rows: conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t* r, const bf16* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, const bf16* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pack_split(float x, float y, uint32_t& hi, uint32_t& lo) {
    bf16 xh, xl, yh, yl;
    split_bf16(x, xh, xl);
    split_bf16(y, yh, yl);
    __nv_bfloat162 h2, l2;
    h2.x = xh; h2.y = yh;
    l2.x = xl; l2.y = yl;
    hi = *reinterpret_cast<uint32_t*>(&h2);
    lo = *reinterpret_cast<uint32_t*>(&l2);
}

struct AttnSmem {
    bf16 kh[TK][LDS], kl[TK][LDS], vh[TK][LDS], vl[TK][LDS];
};

// stage a [64 keys][64 dims] tile of the cache as bf16 hi/lo planes
template <bool SPLIT>
__device__ __forceinline__ void stage_tile(const float* src, int key0, int S, bf16 (*hi)[LDS], bf16 (*lo)[LDS], int tid) {
    for (int e = tid; e < TK * (kHeadDim / 4); e += 128) {
        const int j = e >> 4, c = (e & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key0 + j < S) v = *reinterpret_cast<const float4*>(src + (size_t)(key0 + j) * kHeadDim + c);
        uint32_t h0, l0, h1, l1;
        pack_split(v.x, v.y, h0, l0);
        pack_split(v.z, v.w, h1, l1);
        *reinterpret_cast<uint2*>(&hi[j][c]) = make_uint2(h0, h1);
        if (SPLIT) *reinterpret_cast<uint2*>(&lo[j][c]) = make_uint2(l0, l1);
    }
}
template <bool SPLIT>
__device__ __forceinline__ void stage_tile(const bf16* src, int key0, int S, bf16 (*hi)[LDS], bf16 (*lo)[LDS], int tid) {
    for (int e = tid; e < TK * (kHeadDim / 8); e += 128) {
        const int j = e >> 3, c = (e & 7) * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (key0 + j < S) v = *reinterpret_cast<const uint4*>(src + (size_t)(key0 + j) * kHeadDim + c);
        *reinterpret_cast<uint4*>(&hi[j][c]) = v;
    }
}

// kKvF24 rows (common.cuh): the upper halves are the hi plane as they are; value - hi has at most 8 significant bits, so
// the lo plane is exact as well and hi + lo reproduces the stored 24-bit value.
template <bool SPLIT>
__device__ __forceinline__ void stage_tile(const kv24* src_, int key0, int S, bf16 (*hi)[LDS], bf16 (*lo)[LDS], int tid) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(src_);
    for (int e = tid; e < TK * (kHeadDim / 8); e += 128) {
        const int j = e >> 3, c = (e & 7) * 8;
        uint4 hv = make_uint4(0u, 0u, 0u, 0u);
        uint2 lv = make_uint2(0u, 0u);
        if (key0 + j < S) {
            const unsigned char* row = src + (size_t)(key0 + j) * 192;
            hv = *reinterpret_cast<const uint4*>(row + c * 2);
            lv = *reinterpret_cast<const uint2*>(row + 128 + c);
        }
        *reinterpret_cast<uint4*>(&hi[j][c]) = hv;
        if (SPLIT) {
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
            uint32_t out[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const uint32_t lb = p < 2 ? lv.x >> (16 * p) : lv.y >> (16 * (p - 2));
                const uint32_t h0 = hw[p] << 16, h1 = hw[p] & 0xFFFF0000u;
                const float r0 = __uint_as_float(h0 | ((lb & 0xFFu) << 8)) - __uint_as_float(h0);
                const float r1 = __uint_as_float(h1 | (lb & 0xFF00u)) - __uint_as_float(h1);
                __nv_bfloat162 l2;
                l2.x = __float2bfloat16_rn(r0); l2.y = __float2bfloat16_rn(r1);
                out[p] = *reinterpret_cast<uint32_t*>(&l2);
            }
            *reinterpret_cast<uint4*>(&lo[j][c]) = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
}

#ifdef MB_LAB   // mma.sync causal prefill attention of round 1 (prefill_attn 0): lab builds only, cross-check of attn_umma.cu
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(128) prefill_attention_mma_kernel(const float* __restrict__ q, const T* __restrict__ kc,
                                                                    const T* __restrict__ vc, int S, int t_max,
                                                                    bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
    __shared__ __align__(16) AttnSmem sm;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int kvh = h / (kHeads / kKvHeads);
    const T* kb = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(kc) + ((size_t)b * kKvHeads + kvh) * t_max * KvRowBytes<T>::value);
    const T* vb = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(vc) + ((size_t)b * kKvHeads + kvh) * t_max * KvRowBytes<T>::value);
    const int row0 = qt * TQ + warp * 16 + g, row1 = row0 + 8;
    pdl_trigger();
    pdl_wait();

    // Q fragments (scaled by head_dim^-0.5 = 2^-3, exact) as bf16 hi/lo A operands for the 4 k16 steps
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = (r & 1) ? row1 : row0;
            const int col = ks * 16 + 2 * t + ((r & 2) ? 8 : 0);
            float2 v = make_float2(0.f, 0.f);
            if (row < S) v = *reinterpret_cast<const float2*>(q + ((size_t)b * S + row) * kHidden + h * kHeadDim + col);
            pack_split(v.x * 0.125f, v.y * 0.125f, qh[ks][r], ql[ks][r]);
        }
    }
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    const int n_tiles = min(qt + 1, (S + TK - 1) / TK);
    for (int kt = 0; kt < n_tiles; ++kt) {
        const int k0 = kt * TK;
        __syncthreads();                                             // previous tile fully consumed
        stage_tile<SPLIT>(kb, k0, S, sm.kh, sm.kl, tid);
        stage_tile<SPLIT>(vb, k0, S, sm.vh, sm.vl, tid);
        __syncthreads();

        // S = (Q/8) K^T : 16 x 64 per warp
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const int r = j * 8 + (lane & 7) + (lane >> 4) * 8;
                const int c = ks * 16 + ((lane >> 3) & 1) * 8;
                uint32_t bh[4], bl[4];
                ldsm_x4(bh, &sm.kh[r][c]);
                if (SPLIT) {
                    ldsm_x4(bl, &sm.kl[r][c]);
                    mma16816(s[j], ql[ks], bh[0], bh[1]);
                    mma16816(s[j + 1], ql[ks], bh[2], bh[3]);
                    mma16816(s[j], qh[ks], bl[0], bl[1]);
                    mma16816(s[j + 1], qh[ks], bl[2], bl[3]);
                }
                mma16816(s[j], qh[ks], bh[0], bh[1]);
                mma16816(s[j + 1], qh[ks], bh[2], bh[3]);
            }
        }
        if (kt == qt) {                                              // diagonal tile: causal mask (key > row)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = k0 + j * 8 + 2 * t;
                if (key > row0) s[j][0] = -INFINITY;
                if (key + 1 > row0) s[j][1] = -INFINITY;
                if (key > row1) s[j][2] = -INFINITY;
                if (key + 1 > row1) s[j][3] = -INFINITY;
            }
        }
        // online softmax (rows row0 / row1 live in the 4 lanes of a quad)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);      // finite: key k0 <= row for every row of the tile
        const float a0 = expf(m0 - mn0), a1 = expf(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = expf(s[j][0] - mn0); s[j][1] = expf(s[j][1] - mn0);
            s[j][2] = expf(s[j][2] - mn1); s[j][3] = expf(s[j][3] - mn1);
            ls0 += s[j][0] + s[j][1];
            ls1 += s[j][2] + s[j][3];
            o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1;
        }
        l0 = l0 * a0 + ls0;
        l1 = l1 * a1 + ls1;

        // O += P V : P (16 x 64 keys) re-used from the score accumulators as A fragments
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t ph[4], pl[4];
            pack_split(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
            pack_split(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
            pack_split(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
            pack_split(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int jd = 0; jd < 8; jd += 2) {
                const int blk = lane >> 3;
                const int r = kk * 16 + (blk & 1) * 8 + (lane & 7);
                const int c = (jd + (blk >> 1)) * 8;
                uint32_t bh[4], bl[4];
                ldsm_x4_t(bh, &sm.vh[r][c]);
                if (SPLIT) {
                    ldsm_x4_t(bl, &sm.vl[r][c]);
                    mma16816(o[jd], pl, bh[0], bh[1]);
                    mma16816(o[jd + 1], pl, bh[2], bh[3]);
                    mma16816(o[jd], ph, bl[0], bl[1]);
                    mma16816(o[jd + 1], ph, bl[2], bl[3]);
                }
                mma16816(o[jd], ph, bh[0], bh[1]);
                mma16816(o[jd + 1], ph, bh[2], bh[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int jd = 0; jd < 8; ++jd) {
        const int col = h * kHeadDim + jd * 8 + 2 * t;
        if (row0 < S) store_planes2(out_hi, out_lo, ((size_t)b * S + row0) * kHidden + col, o[jd][0] * i0, o[jd][1] * i0);
        if (row1 < S) store_planes2(out_hi, out_lo, ((size_t)b * S + row1) * kHidden + col, o[jd][2] * i1, o[jd][3] * i1);
    }
}

#endif  // MB_LAB

// ---------------------------------------------------------------------------------------------------------------
// (Shifted-)window attention of the Swin blocks on the tensor cores (reference mellow/model/htsat.py:301-332,
// 414-449).  CTA = (head, window, clip), 4 warps x 16 query tokens; the 64 keys of the window form a single tile, so
// the softmax is a plain one-pass softmax.  head_dim 24 is zero-padded to 32 (two k16 steps / four n8 tiles).
// Shift, window partition / reverse and the -100 mask are index arithmetic exactly as in window_attention_kernel.
constexpr int WHD = 24, WLD = 32 + 8;

struct WinSmem {
    bf16 kh[kWinTok][WLD], kl[kWinTok][WLD], vh[kWinTok][WLD], vl[kWinTok][WLD];
    int tok[kWinTok];
    int lab[kWinTok];
};

__device__ __forceinline__ int shift_band2(int v, int R) { return v < R - kWin ? 0 : (v < R - kWin / 2 ? 1 : 2); }

template <bool SPLIT>
__global__ void __launch_bounds__(128) window_attention_mma_kernel(const float* __restrict__ qkv,
                                                                   const float* __restrict__ relbias,
                                                                   bf16* __restrict__ out_hi, bf16* __restrict__ out_lo,
                                                                   int R, int C, int shift) {
    __shared__ __align__(16) WinSmem sm;
    const int head = blockIdx.x, win = blockIdx.y, clip = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nwx = R / kWin;
    const int wy = win / nwx, wx = win - wy * nwx;
    pdl_trigger();
    if (tid < kWinTok) {
        const int y = wy * kWin + (tid >> 3), x = wx * kWin + (tid & 7);      // coordinates in the shifted frame
        const int sy = (y + shift) % R, sx = (x + shift) % R;
        sm.tok[tid] = clip * R * R + sy * R + sx;
        sm.lab[tid] = shift > 0 ? shift_band2(y, R) * 3 + shift_band2(x, R) : 0;
    }
    pdl_wait();
    __syncthreads();
    // stage K and V of this head: 64 tokens x 24 dims (+8 zero dims) as bf16 hi/lo
    for (int e = tid; e < kWinTok * 8; e += 128) {
        const int j = e >> 3, c = (e & 7) * 4;                                // 8 float4 slots per token row (6 real)
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (c < WHD) {
            const float* row = qkv + (size_t)sm.tok[j] * 3 * C + head * WHD + c;
            kv = *reinterpret_cast<const float4*>(row + C);
            vv = *reinterpret_cast<const float4*>(row + 2 * C);
        }
        uint32_t h0, l0, h1, l1;
        pack_split(kv.x, kv.y, h0, l0); pack_split(kv.z, kv.w, h1, l1);
        *reinterpret_cast<uint2*>(&sm.kh[j][c]) = make_uint2(h0, h1);
        if (SPLIT) *reinterpret_cast<uint2*>(&sm.kl[j][c]) = make_uint2(l0, l1);
        pack_split(vv.x, vv.y, h0, l0); pack_split(vv.z, vv.w, h1, l1);
        *reinterpret_cast<uint2*>(&sm.vh[j][c]) = make_uint2(h0, h1);
        if (SPLIT) *reinterpret_cast<uint2*>(&sm.vl[j][c]) = make_uint2(l0, l1);
    }
    // Q fragments straight from global (q rows of the qkv weight are pre-scaled by head_dim^-0.5 at pack time)
    const int i0 = warp * 16 + g, i1 = i0 + 8;
    uint32_t qh[2][4], ql[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = (r & 1) ? i1 : i0;
            const int col = ks * 16 + 2 * t + ((r & 2) ? 8 : 0);
            float2 v = make_float2(0.f, 0.f);
            if (col < WHD) v = *reinterpret_cast<const float2*>(qkv + (size_t)sm.tok[i] * 3 * C + head * WHD + col);
            pack_split(v.x, v.y, qh[ks][r], ql[ks][r]);
        }
    }
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const int r = j * 8 + (lane & 7) + (lane >> 4) * 8;
            const int c = ks * 16 + ((lane >> 3) & 1) * 8;
            uint32_t bh[4], bl[4];
            ldsm_x4(bh, &sm.kh[r][c]);
            if (SPLIT) {
                ldsm_x4(bl, &sm.kl[r][c]);
                mma16816(s[j], ql[ks], bh[0], bh[1]);
                mma16816(s[j + 1], ql[ks], bh[2], bh[3]);
                mma16816(s[j], qh[ks], bl[0], bl[1]);
                mma16816(s[j + 1], qh[ks], bl[2], bl[3]);
            }
            mma16816(s[j], qh[ks], bh[0], bh[1]);
            mma16816(s[j + 1], qh[ks], bh[2], bh[3]);
        }
    }
    // + relative position bias (+ shifted-window mask), softmax over the 64 keys
    const float* bias = relbias + (size_t)head * kWinTok * kWinTok;
    const int lab0 = sm.lab[i0], lab1 = sm.lab[i1];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int key = j * 8 + 2 * t;
        const float2 b0 = __ldg(reinterpret_cast<const float2*>(bias + i0 * kWinTok + key));
        const float2 b1 = __ldg(reinterpret_cast<const float2*>(bias + i1 * kWinTok + key));
        const int lk0 = sm.lab[key], lk1 = sm.lab[key + 1];
        s[j][0] += b0.x + (lk0 != lab0 ? -100.0f : 0.f);
        s[j][1] += b0.y + (lk1 != lab0 ? -100.0f : 0.f);
        s[j][2] += b1.x + (lk0 != lab1 ? -100.0f : 0.f);
        s[j][3] += b1.y + (lk1 != lab1 ? -100.0f : 0.f);
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j][0] = expf(s[j][0] - mx0); s[j][1] = expf(s[j][1] - mx0);
        s[j][2] = expf(s[j][2] - mx1); s[j][3] = expf(s[j][3] - mx1);
        l0 += s[j][0] + s[j][1];
        l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float r0 = 1.0f / l0, r1 = 1.0f / l1;
    // the reference normalises the probabilities before attn @ v (softmax then matmul, htsat.py:323-329)
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] *= r0; s[j][1] *= r0; s[j][2] *= r1; s[j][3] *= r1; }

    float o[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t ph[4], pl[4];
        pack_split(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
        pack_split(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
        pack_split(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
        pack_split(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
        for (int jd = 0; jd < 4; jd += 2) {
            const int blk = lane >> 3;
            const int r = kk * 16 + (blk & 1) * 8 + (lane & 7);
            const int c = (jd + (blk >> 1)) * 8;
            uint32_t bh[4], bl[4];
            ldsm_x4_t(bh, &sm.vh[r][c]);
            if (SPLIT) {
                ldsm_x4_t(bl, &sm.vl[r][c]);
                mma16816(o[jd], pl, bh[0], bh[1]);
                mma16816(o[jd + 1], pl, bh[2], bh[3]);
                mma16816(o[jd], ph, bl[0], bl[1]);
                mma16816(o[jd + 1], ph, bl[2], bl[3]);
            }
            mma16816(o[jd], ph, bh[0], bh[1]);
            mma16816(o[jd + 1], ph, bh[2], bh[3]);
        }
    }
    const size_t ob0 = (size_t)sm.tok[i0] * C + head * WHD, ob1 = (size_t)sm.tok[i1] * C + head * WHD;
#pragma unroll
    for (int jd = 0; jd < 3; ++jd) {
        const int col = jd * 8 + 2 * t;
        store_planes2(out_hi, out_lo, ob0 + col, o[jd][0], o[jd][1]);
        store_planes2(out_hi, out_lo, ob1 + col, o[jd][2], o[jd][3]);
    }
}

}  // namespace

cudaError_t launch_window_attention_mma(const float* qkv, const float* relbias, bf16* out_hi, bf16* out_lo, int n_clips,
                                        int res, int C, int n_heads, int shift, cudaStream_t st) {
    if (C != n_heads * WHD || res % kWin != 0) return cudaErrorInvalidValue;
    dim3 grid(n_heads, (res / kWin) * (res / kWin), n_clips);
    if (out_lo)
        return launch_k(window_attention_mma_kernel<true>, grid, dim3(128), 0, st, qkv, relbias, out_hi, out_lo, res, C, shift);
    return launch_k(window_attention_mma_kernel<false>, grid, dim3(128), 0, st, qkv, relbias, out_hi, out_lo, res, C, shift);
}

cudaError_t launch_prefill_attention_mma(const float* q, const void* kc, const void* vc, int kv_fmt, int B, int S,
                                         int t_max, bf16* out_hi, bf16* out_lo, cudaStream_t st) {
#ifndef MB_LAB
    (void)q; (void)kc; (void)vc; (void)kv_fmt; (void)B; (void)S; (void)t_max; (void)out_hi; (void)out_lo; (void)st;
    return cudaErrorNotSupported;                       // lab builds only (MB_BUILD_LAB=1)
#else
    dim3 grid((S + TQ - 1) / TQ, kHeads, B);
    if (kv_fmt == kKvBf16)
        return launch_k(prefill_attention_mma_kernel<bf16, false>, grid, dim3(128), 0, st, q, (const bf16*)kc,
                        (const bf16*)vc, S, t_max, out_hi, out_lo);
    if (kv_fmt == kKvF24)
        return launch_k(prefill_attention_mma_kernel<kv24, true>, grid, dim3(128), 0, st, q, (const kv24*)kc,
                        (const kv24*)vc, S, t_max, out_hi, out_lo);
    return launch_k(prefill_attention_mma_kernel<float, true>, grid, dim3(128), 0, st, q, (const float*)kc,
                    (const float*)vc, S, t_max, out_hi, out_lo);
#endif
}

}  // namespace mb
