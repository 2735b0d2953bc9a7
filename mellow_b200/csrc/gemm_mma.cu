// Bring-up GEMM engine: warp-level mma.sync.m16n8k16 bf16 with fp32 accumulation, cp.async 3-stage pipeline,
// ldmatrix operand fetch.  It is the correctness baseline the tcgen05 engine (gemm_umma.cu) is validated against,
// and the engine used for shapes the tcgen05 tiles do not cover.
#include "gemm.cuh"

namespace mb {

namespace {

constexpr int BK = 32;
constexpr int LDS = BK + 8;      // padded smem row (bf16 elements): 80 B, conflict-free for ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int BM, int BN, int STAGES>
struct Smem {
    bf16 a_hi[STAGES][BM][LDS];
    bf16 a_lo[STAGES][BM][LDS];
    bf16 b_hi[STAGES][BN][LDS];
    bf16 b_lo[STAGES][BN][LDS];
};

template <int BM, int BN, int WM, int WN, int EPI, int STAGES>
__global__ void __launch_bounds__(WM * WN * 32) gemm_mma_kernel(const GemmArgs g) {
    constexpr int NT = WM * WN * 32;
    constexpr int TM = BM / WM, TN = BN / WN;      // warp tile
    constexpr int MI = TM / 16, NI = TN / 8;
    static_assert(NI % 2 == 0, "warp tile N must be a multiple of 16");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<BM, BN, STAGES>& sm = *reinterpret_cast<Smem<BM, BN, STAGES>*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const bool split = g.passes == 3;
    // split-K: blockIdx.z owns a contiguous range of K blocks and writes a raw fp32 partial tile
    const int kt_all = g.K / BK;
    const int nsplit = g.split_k > 1 ? g.split_k : 1;
    const int kt_begin = (int)(((long long)kt_all * blockIdx.z) / nsplit);
    const int KT = (int)(((long long)kt_all * (blockIdx.z + 1)) / nsplit) - kt_begin;

    auto load_a = [&](int stage, int kt) {
        const int k0 = (kt_begin + kt) * BK;
        // A: BM rows x 4 chunks of 16 B per plane
        for (int c = tid; c < BM * 4; c += NT) {
            const int r = c >> 2, ch = c & 3;
            const int m = m0 + r;
            const bool ok = m < g.M;
            const size_t off = (size_t)(ok ? m : 0) * g.lda + k0 + ch * 8;
            cp_async16(&sm.a_hi[stage][r][ch * 8], g.A_hi + off, ok);
            if (split) cp_async16(&sm.a_lo[stage][r][ch * 8], g.A_lo + off, ok);
        }
    };
    auto load_b = [&](int stage, int kt) {
        const int k0 = (kt_begin + kt) * BK;
        for (int c = tid; c < BN * 4; c += NT) {
            const int r = c >> 2, ch = c & 3;
            const int n = n0 + r;
            const bool ok = n < g.N;
            const size_t off = (size_t)(ok ? n : 0) * g.ldw + k0 + ch * 8;
            cp_async16(&sm.b_hi[stage][r][ch * 8], g.W_hi + off, ok);
            if (split) cp_async16(&sm.b_lo[stage][r][ch * 8], g.W_lo + off, ok);
        }
    };

    float acc[MI][NI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    // PDL: weights never depend on the preceding kernel, so the first ring slots of W are requested before waiting
    // on it; the activation tiles follow after the wait (they join the same cp.async groups).
    pdl_trigger();
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s)
        if (s < KT) load_b(s, s);
    pdl_wait();
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_a(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) { load_a(nk % STAGES, nk); load_b(nk % STAGES, nk); }
            cp_async_commit();
        }
        const int st = kt % STAGES;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 16) {
            uint32_t ah[MI][4], al[MI][4];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int r = wm * TM + i * 16 + (lane & 15);
                const int c = kk + (lane >> 4) * 8;
                ldsm_x4(ah[i][0], ah[i][1], ah[i][2], ah[i][3], &sm.a_hi[st][r][c]);
                if (split) ldsm_x4(al[i][0], al[i][1], al[i][2], al[i][3], &sm.a_lo[st][r][c]);
            }
#pragma unroll
            for (int j = 0; j < NI; j += 2) {
                uint32_t bh[4], bl[4];
                const int r = wn * TN + j * 8 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk + ((lane >> 3) & 1) * 8;
                ldsm_x4(bh[0], bh[1], bh[2], bh[3], &sm.b_hi[st][r][c]);
                if (split) ldsm_x4(bl[0], bl[1], bl[2], bl[3], &sm.b_lo[st][r][c]);
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    if (split) {
                        mma_bf16(acc[i][j], al[i], bh[0], bh[1]);
                        mma_bf16(acc[i][j + 1], al[i], bh[2], bh[3]);
                        mma_bf16(acc[i][j], ah[i], bl[0], bl[1]);
                        mma_bf16(acc[i][j + 1], ah[i], bl[2], bl[3]);
                    }
                    mma_bf16(acc[i][j], ah[i], bh[0], bh[1]);
                    mma_bf16(acc[i][j + 1], ah[i], bh[2], bh[3]);
                }
            }
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < MI; ++i) {
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int n = n0 + wn * TN + j * 8 + (lane & 3) * 2;
            if (n >= g.N) continue;
            const int r0 = m0 + wm * TM + i * 16 + (lane >> 2);
            if (nsplit > 1) {
                float* p = g.partial + (size_t)blockIdx.z * g.M * g.N;
                if (r0 < g.M) *reinterpret_cast<float2*>(p + (size_t)r0 * g.N + n) = make_float2(acc[i][j][0], acc[i][j][1]);
                if (r0 + 8 < g.M) *reinterpret_cast<float2*>(p + (size_t)(r0 + 8) * g.N + n) = make_float2(acc[i][j][2], acc[i][j][3]);
                continue;
            }
            if (r0 < g.M) epilogue_pair<EPI>(g, r0, n, acc[i][j][0], acc[i][j][1]);
            if (r0 + 8 < g.M) epilogue_pair<EPI>(g, r0 + 8, n, acc[i][j][2], acc[i][j][3]);
        }
    }
}

template <int BM, int BN, int WM, int WN, int EPI, int STAGES = 3>
cudaError_t launch_cfg(const GemmArgs& g, cudaStream_t st) {
    auto kern = gemm_mma_kernel<BM, BN, WM, WN, EPI, STAGES>;
    const int smem = (int)sizeof(Smem<BM, BN, STAGES>);
    static bool configured = false;     // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.split_k > 1 ? g.split_k : 1);
    return launch_k(kern, grid, dim3(WM * WN * 32), (size_t)smem, st, g);
}

template <int EPI>
cudaError_t launch_epi(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 128 && g.N <= 4096) {
        // decode-sized: one M tile; 16-column tiles (x split-K) so ~100+ SMs stream the weights, 6-deep cp.async ring
        // (every CTA re-reads the whole activation tile from L2, so wide-N GEMMs take 32-column tiles)
        if (g.M <= 64) return launch_cfg<64, 16, 4, 1, EPI, 6>(g, st);
        if (g.N >= 2048) return launch_cfg<128, 32, 4, 2, EPI, 6>(g, st);
        return launch_cfg<128, 16, 8, 1, EPI, 6>(g, st);
    }
    if (g.N % 96 == 0 && g.N % 128 != 0) return launch_cfg<128, 96, 4, 2, EPI>(g, st);
    return launch_cfg<128, 128, 2, 4, EPI>(g, st);
}

}  // namespace

cudaError_t launch_gemm_mma(const GemmArgs& g, int epi, cudaStream_t st) {
    if (g.K % BK != 0 || g.M <= 0 || g.N <= 0) return cudaErrorInvalidValue;
    if (g.split_k > 1 && (epi != EPI_GENERIC || !g.partial || (g.N & 1))) return cudaErrorInvalidValue;
    switch (epi) {
        case EPI_GENERIC: return launch_epi<EPI_GENERIC>(g, st);
        case EPI_SWIGLU: return launch_epi<EPI_SWIGLU>(g, st);
        case EPI_QKV_ROPE: return launch_epi<EPI_QKV_ROPE>(g, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mb
