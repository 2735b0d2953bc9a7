// tcgen05 / TMA GEMM engine for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T, bf16 operand planes, fp32 accumulation in TMEM.
//
//   * operands: TMA (cp.async.bulk.tensor.2d, 128B swizzle) stages 128x64 A tiles and BNx64 W tiles, hi and lo planes,
//     into a 3/4-deep shared-memory ring guarded by full/empty mbarriers;
//   * math: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) from shared-memory
//     descriptors; with the split policy each k-step is three MMAs (hi*hi, hi*lo, lo*hi) into the same TMEM
//     accumulator, which reproduces an fp32 GEMM to ~2^-16 while staying on the tensor pipe;
//   * epilogue: four warps read the accumulator with tcgen05.ld (one row per thread) and run the same fused
//     epilogues as the mma.sync engine (gemm.cuh): bias / GELU / sigmoid / residual, SwiGLU, QKV + RoPE + KV write.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quadrant = warp % 4).
#include <cuda.h>

#include "gemm.cuh"

namespace mb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                  // 64 bf16 = 128 B = one swizzle row
constexpr uint32_t A_BYTES = BM * BK * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();          // ~2 s: a broken pipeline must fail, not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor fields)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);       // start address        bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset  bits [16,30) (16 B units; unused for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN, bool SPLIT>
struct Cfg {
    static constexpr uint32_t B_BYTES = BN * BK * 2;
    static constexpr uint32_t STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int STAGES_FIT = (int)((225u * 1024u - 1280u) / STAGE_BYTES);
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : 128);
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(STAGES >= 2, "tile does not fit");
    static_assert(4 * 32 * 65 * 4 <= STAGES * STAGE_BYTES, "epilogue staging tile must fit in the drained ring");
};

template <int BN, int EPI, bool SPLIT>
__global__ void __launch_bounds__(192, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const GemmArgs g) {
    using C = Cfg<BN, SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tmem_full = empty + C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // split-K: blockIdx.z owns a contiguous range of 64-wide k-blocks and writes a raw fp32 partial tile
    const int kb_all = (g.K + BK - 1) / BK;
    const int nsplit = g.split_k > 1 ? g.split_k : 1;
    const int kb_begin = (int)(((long long)kb_all * blockIdx.z) / nsplit);
    const int KB = (int)(((long long)kb_all * (blockIdx.z + 1)) / nsplit) - kb_begin;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // PDL: the weight (B) tiles of the first ring slots never depend on the preceding kernel, so they are
            // requested before the dependency wait; the activation (A) tiles of those slots follow after it.
            const int npre = KB < C::STAGES ? KB : C::STAGES;
            for (int kb = 0; kb < npre; ++kb) {
                unsigned char* st = smem + (size_t)kb * C::STAGE_BYTES;
                mbar_expect_tx(&full[kb], C::STAGE_BYTES);
                tma_load_2d(st + A_BYTES, &tm_b_hi, &full[kb], (kb_begin + kb) * BK, n0);
                if (SPLIT) tma_load_2d(st + 2 * A_BYTES + C::B_BYTES, &tm_b_lo, &full[kb], (kb_begin + kb) * BK, n0);
            }
            pdl_wait();
            for (int kb = 0; kb < npre; ++kb) {
                unsigned char* st = smem + (size_t)kb * C::STAGE_BYTES;
                tma_load_2d(st, &tm_a_hi, &full[kb], (kb_begin + kb) * BK, m0);
                if (SPLIT) tma_load_2d(st + A_BYTES + C::B_BYTES, &tm_a_lo, &full[kb], (kb_begin + kb) * BK, m0);
            }
            for (int kb = npre; kb < KB; ++kb) {
                const int s = kb % C::STAGES;
                const uint32_t ph = (kb / C::STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char* st = smem + (size_t)s * C::STAGE_BYTES;
                mbar_expect_tx(&full[s], C::STAGE_BYTES);
                tma_load_2d(st, &tm_a_hi, &full[s], (kb_begin + kb) * BK, m0);
                tma_load_2d(st + A_BYTES, &tm_b_hi, &full[s], (kb_begin + kb) * BK, n0);
                if (SPLIT) {
                    tma_load_2d(st + A_BYTES + C::B_BYTES, &tm_a_lo, &full[s], (kb_begin + kb) * BK, m0);
                    tma_load_2d(st + 2 * A_BYTES + C::B_BYTES, &tm_b_lo, &full[s], (kb_begin + kb) * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N>>3, M>>4
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % C::STAGES;
                const uint32_t ph = (kb / C::STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
                const uint32_t b_hi = a_hi + A_BYTES;
                const uint32_t a_lo = b_hi + C::B_BYTES;
                const uint32_t b_lo = a_lo + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint32_t off = k * 32;                     // 16 bf16 = 32 B along the swizzled row
                    umma_bf16(tmem_base, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, (kb | k) != 0);
                    if (SPLIT) {
                        umma_bf16(tmem_base, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, 1);
                        umma_bf16(tmem_base, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, 1);
                    }
                }
                umma_commit(&empty[s]);                              // frees the smem stage once these MMAs retire
            }
            umma_commit(tmem_full);                                  // accumulator complete
        }
    } else {
        // Each thread owns one accumulator row in TMEM, but row-per-thread global stores touch 32 different rows per
        // instruction.  The drained smem ring is reused as a per-warp [32 rows][64+1] fp32 staging tile so that the
        // fused epilogue runs with lane = column pair: every store instruction covers one contiguous row segment.
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may access
        constexpr int LDT = 65;
        float* stage_t = reinterpret_cast<float*>(smem) + (size_t)q * 32 * LDT;
        pdl_wait();                                                  // residual reads / output writes depend on the predecessor
        mbar_wait(tmem_full, 0);                                     // all MMAs retired: accumulator ready, smem ring idle
        tc_fence_after();
        if (BN <= 32) {
            // decode-sized tiles: 16-32 columns per row.  Every thread finishes its own row straight from registers
            // (128 rows in parallel); the staged path below would leave 24 of 32 lanes idle.
            const int m = m0 + q * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (m < g.M) {
                    if (nsplit > 1) {
                        float* pz = g.partial + (size_t)blockIdx.z * g.M * g.N + (size_t)m * g.N + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            if (n0 + c0 + j < g.N) *reinterpret_cast<float4*>(pz + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            if (n0 + c0 + j < g.N) epilogue_pair<EPI>(g, m, n0 + c0 + j, v[j], v[j + 1]);
                    }
                }
            }
        } else
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
#pragma unroll
            for (int cc = 0; cc < 64; cc += 16) {
                if (c0 + cc < BN) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + cc), v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) stage_t[lane * LDT + cc + j] = v[j];
                }
            }
            __syncwarp();
            const int n = n0 + c0 + 2 * lane;
            if (nsplit > 1) {
                if (c0 + 2 * lane < BN && n < g.N) {
                    float* pz = g.partial + (size_t)blockIdx.z * g.M * g.N;
                    for (int r = 0; r < 32; ++r) {
                        const int m = m0 + q * 32 + r;
                        if (m >= g.M) break;
                        *reinterpret_cast<float2*>(pz + (size_t)m * g.N + n) =
                            make_float2(stage_t[r * LDT + 2 * lane], stage_t[r * LDT + 2 * lane + 1]);
                    }
                }
            } else if (c0 + 2 * lane < BN && n < g.N) {
#pragma unroll 1
                for (int rb = 0; rb < 32; rb += 8) {
                    const int mb = m0 + q * 32 + rb;
                    float res[8][2];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                     // residual loads of 8 rows in flight together
                        res[i][0] = 0.f; res[i][1] = 0.f;
                        if (mb + i < g.M) load_residual_pair<EPI>(g, mb + i, n, res[i][0], res[i][1]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (mb + i < g.M)
                            epilogue_pair_r<EPI>(g, mb + i, n, stage_t[(rb + i) * LDT + 2 * lane],
                                                 stage_t[(rb + i) * LDT + 2 * lane + 1], res[i][0], res[i][1]);
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, K] bf16 row-major with leading dimension ld -> 2-D map with a {64, box_rows} box, 128B swizzle, zero OOB fill
bool make_map(CUtensorMap* map, const bf16* ptr, int rows, int K, int ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int EPI, bool SPLIT>
cudaError_t launch_one(const GemmArgs& g, cudaStream_t st) {
    using C = Cfg<BN, SPLIT>;
    auto kern = gemm_umma_kernel<BN, EPI, SPLIT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (!make_map(&ta_hi, g.A_hi, g.M, g.K, g.lda, BM) || !make_map(&tb_hi, g.W_hi, g.N, g.K, g.ldw, BN))
        return cudaErrorInvalidValue;
    if (SPLIT) {
        if (!make_map(&ta_lo, g.A_lo, g.M, g.K, g.lda, BM) || !make_map(&tb_lo, g.W_lo, g.N, g.K, g.ldw, BN))
            return cudaErrorInvalidValue;
    } else {
        ta_lo = ta_hi;
        tb_lo = tb_hi;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.split_k > 1 ? g.split_k : 1);
    return launch_k(kern, grid, dim3(192), C::SMEM, st, ta_hi, ta_lo, tb_hi, tb_lo, g);
}

template <int EPI>
cudaError_t launch_epi(const GemmArgs& g, cudaStream_t st) {
    const bool split = g.passes == 3;
    if (g.M <= 128 && g.N <= 4096) {
        // decode-sized (one 128-row UMMA tile): narrow N tiles (x split-K) so 60-110 SMs stream the weights
        if (g.N >= 2048) return split ? launch_one<32, EPI, true>(g, st) : launch_one<32, EPI, false>(g, st);
        return split ? launch_one<16, EPI, true>(g, st) : launch_one<16, EPI, false>(g, st);
    }
    const bool bn96 = (g.N % 128 != 0) && (g.N % 96 == 0);
    if (bn96) return split ? launch_one<96, EPI, true>(g, st) : launch_one<96, EPI, false>(g, st);
    return split ? launch_one<128, EPI, true>(g, st) : launch_one<128, EPI, false>(g, st);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

cudaError_t launch_gemm_umma(const GemmArgs& g, int epi, cudaStream_t st, bool* handled) {
    *handled = false;
    if (g.M < 1 || g.K < BK || (g.K % 8) != 0 || (g.lda % 8) != 0 || (g.ldw % 8) != 0) return cudaSuccess;
    if (g.split_k > 1 && (epi != EPI_GENERIC || !g.partial || (g.N & 1))) return cudaErrorInvalidValue;
    if (!aligned16(g.A_hi) || !aligned16(g.W_hi) || (g.passes == 3 && (!aligned16(g.A_lo) || !aligned16(g.W_lo)))) return cudaSuccess;
    if (!encode_fn()) return cudaSuccess;
    cudaError_t e;
    switch (epi) {
        case EPI_GENERIC: e = launch_epi<EPI_GENERIC>(g, st); break;
        case EPI_SWIGLU: e = launch_epi<EPI_SWIGLU>(g, st); break;
        case EPI_QKV_ROPE: e = launch_epi<EPI_QKV_ROPE>(g, st); break;
        default: return cudaErrorInvalidValue;
    }
    *handled = (e == cudaSuccess);
    return e;
}

}  // namespace mb
