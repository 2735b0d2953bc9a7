// Decode GEMM that consumes the fp32 residual stream directly ("conv" = converting producer):
//
//     out = epilogue( rstd[m] * ( x[m,:] @ (W o ln_w)^T ) )        rstd[m] = rsqrt(mean(x[m,:]^2) + eps)
//
// i.e. RMSNorm(x) * ln_w followed by the projection, with the norm gain folded into the weight columns at pack time
// and the per-row scale applied to the accumulator.  The eight epilogue warps first act as the A-operand producer:
// they read the 128 x K fp32 rows of x once (the same bytes as the bf16 hi+lo planes a separate norm kernel would have
// written and TMA re-read), accumulate the row sums of squares on the way, split to bf16 hi/lo and write the tiles
// straight into the 128B-swizzled shared-memory ring that tcgen05.mma reads; TMA only streams the weight tiles.
// This removes the standalone residual+RMSNorm kernels from the decode chain: 5 dependent kernels per layer, not 7.
#include "kernels.cuh"
#include "umma.cuh"

namespace mb {

using namespace umma;

namespace {

template <int BN, bool SPLIT>
struct ConvCfg {
    static constexpr uint32_t B_BYTES = BN * BK * 2;
    static constexpr uint32_t STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int STAGES = 5;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + 128 * sizeof(float);
};

__device__ __forceinline__ uint32_t pack2(float a, float b, uint32_t& lo) {
    bf16 ah, al, bh, bl;
    split_bf16(a, ah, al);
    split_bf16(b, bh, bl);
    __nv_bfloat162 h2, l2;
    h2.x = ah; h2.y = bh; l2.x = al; l2.y = bl;
    lo = *reinterpret_cast<uint32_t*>(&l2);
    return *reinterpret_cast<uint32_t*>(&h2);
}

template <int BN, int EPI, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const GemmArgs g) {
    using C = ConvCfg<BN, SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    float* rstd_s = reinterpret_cast<float*>(smem + (size_t)C::STAGES * C::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    const int KB = g.K / BK;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1 + kEpiWarps); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                             // weights only: no dependency on the predecessor
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % C::STAGES;
                mbar_wait(&empty[s], ((kb / C::STAGES) & 1) ^ 1);
                unsigned char* st = smem + (size_t)s * C::STAGE_BYTES;
                mbar_expect_tx(&full[s], (SPLIT ? 2u : 1u) * C::B_BYTES);
                tma_load_2d(st + A_BYTES, &tm_b_hi, &full[s], kb * BK, n0);
                if (SPLIT) tma_load_2d(st + 2 * A_BYTES + C::B_BYTES, &tm_b_lo, &full[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % C::STAGES;
                mbar_wait(&full[s], (kb / C::STAGES) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
                const uint32_t b_hi = a_hi + A_BYTES;
                const uint32_t a_lo = b_hi + C::B_BYTES;
                const uint32_t b_lo = a_lo + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint32_t off = k * 32;
                    umma_bf16(tmem_base, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, (kb | k) != 0);
                    if (SPLIT) {
                        umma_bf16(tmem_base, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, 1);
                        umma_bf16(tmem_base, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, 1);
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(tfull);
        }
    } else {
        // ---- A-operand producer: thread = (row r, 32-column half h) of every 128 x 64 k-block
        const int ct = threadIdx.x - 64;
        const int r = ct >> 1, h = ct & 1;
        const bool row_ok = r < g.M;
        pdl_wait();                                                  // x is complete once the predecessor has finished
        const float* xr = g.a_f32 + (size_t)(row_ok ? r : 0) * g.lda32 + h * 32;
        // software pipeline over the k-blocks: the loads of k-blocks kb+1..kb+3 are in flight while kb is converted
        // (a single block of look-ahead would expose one L2 round trip per k-block)
        constexpr int KBC = kHidden / BK;                            // 9: every x-consuming GEMM of this model has K = 576
        constexpr int DEPTH = 4;
        float4 buf[DEPTH][8];
        auto load_kb = [&](int kb, float4* dst) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                dst[i] = row_ok ? *reinterpret_cast<const float4*>(xr + kb * BK + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
#pragma unroll
        for (int p = 0; p < DEPTH - 1; ++p) load_kb(p, buf[p]);
        float sq = 0.f;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
        for (int kb = 0; kb < KBC; ++kb) {
            if (kb + DEPTH - 1 < KBC) load_kb(kb + DEPTH - 1, buf[(kb + DEPTH - 1) % DEPTH]);
            const float4* cur = buf[kb % DEPTH];
            const int s = kb % C::STAGES;
            mbar_wait(&empty[s], ((kb / C::STAGES) & 1) ^ 1);
            unsigned char* a_hi = smem + (size_t)s * C::STAGE_BYTES;
            unsigned char* a_lo = a_hi + A_BYTES + C::B_BYTES;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {                          // four 16-byte chunks (8 bf16 each) of this half row
                const float4 u = cur[2 * c4], w = cur[2 * c4 + 1];
                sq += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w + w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
                uint32_t l0, l1, l2, l3;
                const uint32_t h0 = pack2(u.x, u.y, l0), h1 = pack2(u.z, u.w, l1), h2 = pack2(w.x, w.y, l2), h3 = pack2(w.z, w.w, l3);
                const uint32_t chunk = (uint32_t)(h * 4 + c4);
                const uint32_t off = row_off + ((chunk ^ (uint32_t)(r & 7)) << 4);      // 128B swizzle: chunk ^= row % 8
                *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h0, h1, h2, h3);
                if (SPLIT) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l0, l1, l2, l3);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");               // generic smem writes -> tcgen05 (async proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        if (h == 0) rstd_s[r] = rsqrtf(sq * (1.0f / (float)g.k_norm) + 1e-5f);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        // ---- epilogue: one accumulator row per thread, scaled by its rstd
        const int q = warp & 3, half = (warp - 2) >> 2;
        mbar_wait(tfull, 0);
        tc_fence_after();
        if (half * 16 < BN) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16), v);
            const int m = q * 32 + lane, n = n0 + half * 16;
            if (m < g.M && n < g.N) {
                const float rs = rstd_s[m];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= rs;
                epilogue_row16<EPI>(g, m, n, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
    }
}

template <int BN, int EPI, bool SPLIT>
cudaError_t launch_conv(const GemmArgs& g, cudaStream_t st) {
    using C = ConvCfg<BN, SPLIT>;
    auto kern = gemm_conv_kernel<BN, EPI, SPLIT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    CUtensorMap tb_hi, tb_lo;
    if (!make_map(&tb_hi, g.W_hi, g.N, g.K, g.ldw, BN)) return cudaErrorInvalidValue;
    if (SPLIT) { if (!make_map(&tb_lo, g.W_lo, g.N, g.K, g.ldw, BN)) return cudaErrorInvalidValue; }
    else tb_lo = tb_hi;
    return launch_k(kern, dim3((g.N + BN - 1) / BN), dim3(kThreads), C::SMEM, st, tb_hi, tb_lo, g);
}

template <int EPI>
cudaError_t launch_conv_epi(const GemmArgs& g, cudaStream_t st) {
    const bool split = g.passes == 3;
    if (g.N >= 2048) return split ? launch_conv<32, EPI, true>(g, st) : launch_conv<32, EPI, false>(g, st);
    return split ? launch_conv<16, EPI, true>(g, st) : launch_conv<16, EPI, false>(g, st);
}

}  // namespace

// x-consuming decode GEMM.  Requires M <= 128, K % 64 == 0, N % 16 == 0 and at most 148 N tiles.
cudaError_t launch_gemm_conv(const GemmArgs& g, int epi, cudaStream_t st) {
    if (g.M < 1 || g.M > BM || g.K != kHidden || (g.N & 15) || !g.a_f32 || (g.lda32 & 3) || g.split_k > 1 || !encode_fn())
        return cudaErrorInvalidValue;
    if ((g.N + (g.N >= 2048 ? 32 : 16) - 1) / (g.N >= 2048 ? 32 : 16) > 148) return cudaErrorInvalidValue;
    switch (epi) {
        case EPI_SWIGLU: return launch_conv_epi<EPI_SWIGLU>(g, st);
        case EPI_QKV_ROPE: return launch_conv_epi<EPI_QKV_ROPE>(g, st);
        case EPI_GENERIC: return launch_conv_epi<EPI_GENERIC>(g, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mb
