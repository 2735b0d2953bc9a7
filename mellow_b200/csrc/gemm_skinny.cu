// Decode-sized tcgen05 GEMM (M <= 128 rows, one tile per CTA):  C[M,N] = A[M,K] * W[N,K]^T, same operand planes,
// MMA passes and fused epilogues as gemm_umma.cu, but laid out for the decode step's dependency chain:
//
//   * WEIGHT-RESIDENT: the CTA's whole weight slice (BN rows x its K range, <= KBMAX k-blocks) has its own shared
//     memory region and is requested in full BEFORE griddepcontrol.wait.  The kernel is resident several microseconds
//     before its predecessor finishes (PDL), so no weight byte is waited for after the dependency resolves (the ring
//     version could only prefetch as many k-blocks as it had stages).
//   * the 128-row activation tile streams through a ring of its own (as deep as the remaining shared memory allows),
//     requested right after the wait: these are L2 hits (the predecessor just wrote them); the first stage lands
//     ~0.7 us after the wait, and the re-read of the tile by every N tile (tiles_n x 128 x K x 4 B) is what bounds the
//     k-loop once the issue loop is out of the way.
//   * one elect.sync-elected thread issues the MMAs from a fully unrolled loop; eight epilogue warps drain the TMEM
//     accumulator (one row per thread) into the fused epilogue or the split-K partial buffer.
//   What the in-kernel timeline (tools/decode_timeline.py) showed on the way here: the ring kernel spent 0.8 us per
//   k-block whatever the tile width, stage count, accumulator layout or commit count; an EMPTY k-block iteration of the
//   single issuing thread already cost 0.3 us because every uniform-datapath operand was wrapped in an
//   ELECT / R2UR.BROADCAST loop (`lane == 0` does not tell the compiler that one thread is active) and descriptors
//   were rebuilt in 64-bit arithmetic.  Bodies went 8.2 / 3.4 / 8.8 / 5.6 us -> 4.1 / 2.8 / 5.2 / 3.7 us
//   (QKV / o_proj / gate-up / down at B = 128).
//
// Warp roles (320 threads): warps 0..7 = epilogue, warp 8 = TMA producer, warp 9 = TMEM allocator + MMA issuer.  The two
// single-thread roles sit in the HIGHEST-numbered warps of their schedulers (warp id % 4) and the epilogue warps sleep
// between polls of the accumulator barrier: with the roles the other way round (and a bare polling loop) the waiting
// epilogue warps took most issue slots of the shared schedulers and an EMPTY k-block iteration of the MMA thread cost
// 0.3 us (measured, tools/decode_timeline.py with MB_GEMM_DBG=3).
#include "umma.cuh"

namespace mb {

namespace {

using namespace umma;

template <int BN, bool SPLIT, int KBMAX>
struct SkinnyCfg {
    static constexpr uint32_t PLANES = SPLIT ? 2 : 1;
    static constexpr uint32_t B_BYTES = BN * BK * 2;                    // one plane of one k-block
    static constexpr uint32_t B_REGION = KBMAX * PLANES * B_BYTES;
    static constexpr uint32_t A_STAGE = PLANES * A_BYTES;
    static constexpr uint32_t BAR_BYTES = 512;
    static constexpr uint32_t BUDGET = 227u * 1024u - 1024u - BAR_BYTES - B_REGION;
    static constexpr int FIT = (int)(BUDGET / A_STAGE);
    static constexpr int STAGES = FIT > KBMAX ? KBMAX : FIT;
    static constexpr uint32_t ACC_N = PLANES * BN;                      // accumulator columns: [a*b_hi | a_hi*b_lo]
    static constexpr uint32_t ACC_COLS = ACC_N <= 32 ? 32 : (ACC_N <= 64 ? 64 : 128);
    static constexpr uint32_t TMEM_COLS = ACC_COLS;
    static_assert(ACC_N <= 128 && (BN % 16) == 0, "N tile");
    static constexpr size_t SMEM = (size_t)B_REGION + (size_t)STAGES * A_STAGE + 1024 + BAR_BYTES;
    static_assert(STAGES >= 2, "activation ring does not fit");
    static_assert((B_BYTES % 1024) == 0, "swizzled tiles start on 1024-byte boundaries");
    static_assert((2 * KBMAX + 2 * 8 + 2) * 8 + 8 <= BAR_BYTES, "barrier block too small");
};

template <int BN, int EPI, bool SPLIT, int KBMAX>
__global__ void __launch_bounds__(kThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const GemmArgs g) {
    using C = SkinnyCfg<BN, SPLIT, KBMAX>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* breg = smem;                                       // [KBMAX][PLANES][BN x 64] weight slice
    unsigned char* areg = smem + C::B_REGION;                         // [STAGES][PLANES][128 x 64] activation ring
    uint64_t* bfull = reinterpret_cast<uint64_t*>(areg + (size_t)C::STAGES * C::A_STAGE);
    uint64_t* afull = bfull + KBMAX;
    uint64_t* aempty = afull + C::STAGES;
    uint64_t* tfull = aempty + C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    uint32_t* trace_slot = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsplit = g.split_k > 1 ? g.split_k : 1;
    const int tiles_n = (g.N + BN - 1) / BN;
    const int z = blockIdx.x / tiles_n;                               // K split owned by this CTA
    const int n0 = (blockIdx.x - z * tiles_n) * BN;
    const int kb_all = (g.K + BK - 1) / BK;
    const int kb_begin = (int)(((long long)kb_all * z) / nsplit);
    const int KB = (int)(((long long)kb_all * (z + 1)) / nsplit) - kb_begin;      // <= KBMAX (checked by the launcher)

    constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
    pdl_trigger();
    if (threadIdx.x == kProducerWarp * 32) {
        // descriptor fetches are latency the dependency wait would otherwise expose (the A maps are first used after it)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
        if (SPLIT) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
        }
        *trace_slot = trace_open(g.trace, g.trace_id);
        for (int i = 0; i < KBMAX; ++i) mbar_init(&bfull[i], 1);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // the whole weight slice, before the dependency wait (weights are never produced by a predecessor kernel)
        for (int kb = 0; kb < KB; ++kb) {
            unsigned char* bt = breg + (size_t)kb * C::PLANES * C::B_BYTES;
            mbar_expect_tx(&bfull[kb], C::PLANES * C::B_BYTES);
            tma_load_2d(bt, &tm_b_hi, &bfull[kb], (kb_begin + kb) * BK, n0);
            if (SPLIT) tma_load_2d(bt + C::B_BYTES, &tm_b_lo, &bfull[kb], (kb_begin + kb) * BK, n0);
        }
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const unsigned trec = first_cta() ? *trace_slot : kTraceNone;    // fine-grained stamps come from the first CTA only

    // Both single-thread loops below are fully unrolled over the (compile-time bounded) k-blocks: stage indices,
    // barrier parities and shared-memory offsets become constants, and what remains per MMA is one 32-bit add per
    // descriptor plus the instruction itself.  A lone thread issues dependent instructions ~5 cycles apart, so the
    // generic loop (runtime stage arithmetic, 64-bit descriptor rebuilds, a watchdog clock read per barrier) cost
    // ~0.3 us per k-block before a single MMA ran -- more than the MMAs themselves (tools/decode_timeline.py).
    if (warp == kProducerWarp) {
        if (elect_one()) {
            pdl_wait();
            trace_put(g.trace, trec, g.trace_id, TR_WAITED);
#pragma unroll
            for (int kb = 0; kb < KBMAX; ++kb) {
                if (kb < KB) {
                    const int s = kb % C::STAGES;
                    if (kb >= C::STAGES) mbar_wait_fast(&aempty[s], ((kb / C::STAGES) - 1) & 1);
                    unsigned char* at = areg + (size_t)s * C::A_STAGE;
                    mbar_expect_tx(&afull[s], C::A_STAGE);
                    tma_load_2d(at, &tm_a_hi, &afull[s], (kb_begin + kb) * BK, 0);
                    if (SPLIT) tma_load_2d(at + A_BYTES, &tm_a_lo, &afull[s], (kb_begin + kb) * BK, 0);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (elect_one()) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t a_lo0 = umma_desc_lo(smem_u32(areg));      // descriptor low words of stage 0 / k-block 0
            const uint32_t b_lo0 = umma_desc_lo(smem_u32(breg));
            // Split policy in TWO MMAs per k-step instead of three: the weight slice keeps its hi rows and its lo rows
            // adjacent in shared memory, so one MMA with N = 2*BN forms a_hi*[b_hi | b_lo] in one pass over the
            // 128-row activation tile; a_lo*b_hi accumulates onto the first BN columns.  The epilogue adds columns
            // [0,BN) and [BN,2BN).
#pragma unroll
            for (int kb = 0; kb < KBMAX; ++kb) {
                if (kb < KB) {
                    const int s = kb % C::STAGES;
                    mbar_wait_fast(&bfull[kb], 0);
                    if (kb == 0) trace_put(g.trace, trec, g.trace_id, 4);            // weights of k-block 0 present
                    mbar_wait_fast(&afull[s], (kb / C::STAGES) & 1);
                    if (kb == 0) trace_put(g.trace, trec, g.trace_id, 5);            // first activation stage landed
                    tc_fence_after();
                    const uint32_t ah = a_lo0 + (uint32_t)(s * (C::A_STAGE >> 4));
                    const uint32_t al = ah + (A_BYTES >> 4);
                    const uint32_t bh = b_lo0 + (uint32_t)(kb * ((C::PLANES * C::B_BYTES) >> 4));   // lo rows follow the hi rows
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {              // 16 bf16 = 32 B = 2 descriptor units along the swizzled row
                        const uint64_t da = umma_desc_join(ah + 2 * k), db = umma_desc_join(bh + 2 * k);
                        if (SPLIT) {
                            if (kb == 0 && k == 0) umma_bf16_c<false>(tmem_base, da, db, idesc2);
                            else umma_bf16_c<true>(tmem_base, da, db, idesc2);
                            umma_bf16_c<true>(tmem_base, umma_desc_join(al + 2 * k), db, idesc);
                        } else {
                            if (kb == 0 && k == 0) umma_bf16_c<false>(tmem_base, da, db, idesc);
                            else umma_bf16_c<true>(tmem_base, da, db, idesc);
                        }
                    }
                    // a stage is only handed back when a later k-block will actually be loaded into it
                    if (kb + C::STAGES < KBMAX && kb + C::STAGES < KB) umma_commit(&aempty[s]);
                }
            }
            umma_commit(tfull);                                      // accumulator complete
            trace_put(g.trace, trec, g.trace_id, 6);                 // all MMAs issued
        }
    } else {
        // One accumulator row per thread (TMEM lane = row); the two warps of a lane quadrant take the 16-column units
        // of the tile alternately, so a 32-column tile keeps all eight warps busy.
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may access
        const int half = warp >> 2;
        constexpr int kUnits = BN / 16;
        pdl_wait();                                                  // residual reads / output writes depend on the predecessor
        const int m = q * 32 + lane;
        constexpr int kStgLd = BN + 4;                               // staging row stride (floats): 16-byte aligned, conflict-free
        float* stg = reinterpret_cast<float*>(areg);                 // free once the accumulator is complete (all MMAs retired)
        if (half < kUnits) {
            float rc[8], rs[8];
            if (EPI == EPI_QKV_ROPE && nsplit == 1 && m < g.M)       // RoPE factors of the first unit, fetched while the MMAs run
                qkv_rope_load(g, m, n0 + half * 16, rc, rs);
            mbar_wait_sleep(tfull, 0, (unsigned)g.epi_sleep);
            if (threadIdx.x == 0) trace_put(g.trace, trec, g.trace_id, 7);              // accumulator complete
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int u = half; u < kUnits; u += 2) {
                const int c0 = u * 16;
                float v[16];
                tmem_ld16(tacc + (uint32_t)c0, v);
                if (SPLIT) {                                         // + a_hi * b_lo
                    float w[16];
                    tmem_ld16(tacc + (uint32_t)(BN + c0), w);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += w[j];
                }
                const int n = n0 + c0;
                if (nsplit > 1) {
                    // split-K partial sums: staged in the (idle) activation ring so that the tile leaves the SM as
                    // whole 128-byte row segments (below) instead of 32 scattered 16-byte pieces per store instruction
#pragma unroll
                    for (int j = 0; j < 16; j += 4) st4(stg + m * kStgLd + c0 + j, v + j);
                } else if (m < g.M && n < g.N) {
                    if (EPI == EPI_QKV_ROPE && n + 16 <= g.N) {
                        if (u != half) qkv_rope_load(g, m, n, rc, rs);
                        epilogue_row16_qkv(g, m, n, v, rc, rs);
                    } else {
                        epilogue_row16<EPI>(g, m, n, v);
                    }
                }
            }
        }
        if (nsplit > 1) {
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");        // the eight epilogue warps only
            constexpr int kC4 = BN / 4;                              // float4 pieces per tile row
            float* pz = g.partial + (size_t)z * g.M * g.N + n0;
            for (int idx = threadIdx.x; idx < BM * kC4; idx += kEpiWarps * 32) {
                const int row = idx / kC4, c = (idx - row * kC4) * 4;
                if (row < g.M && n0 + c < g.N)
                    *reinterpret_cast<float4*>(pz + (size_t)row * g.N + c) = *reinterpret_cast<const float4*>(stg + row * kStgLd + c);
            }
        }
    }
    if (threadIdx.x == 0) trace_put(g.trace, trec, g.trace_id, 8);                      // this warp's epilogue stores issued
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == kProducerWarp * 32) trace_close(g.trace, *trace_slot, g.trace_id);
    if (warp == kMmaWarp) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

template <int BN, int EPI, bool SPLIT, int KBMAX>
cudaError_t launch_skinny(const GemmArgs& g, cudaStream_t st) {
    using C = SkinnyCfg<BN, SPLIT, KBMAX>;
    auto kern = gemm_skinny_kernel<BN, EPI, SPLIT, KBMAX>;
    static bool configured[kMaxDevices] = {};
    if (cudaError_t e = ensure_smem(kern, C::SMEM, configured); e != cudaSuccess) return e;
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
    const int total = ((g.N + BN - 1) / BN) * (g.split_k > 1 ? g.split_k : 1);
    return launch_k(kern, dim3((unsigned)total), dim3(kThreads), C::SMEM, st, ta_hi, ta_lo, tb_hi, tb_lo, g);
}

template <int BN, int EPI, int KBMAX>
cudaError_t launch_skinny_p(const GemmArgs& g, cudaStream_t st) {
    return g.passes == 3 ? launch_skinny<BN, EPI, true, KBMAX>(g, st) : launch_skinny<BN, EPI, false, KBMAX>(g, st);
}

template <int EPI>
cudaError_t launch_skinny_epi(const GemmArgs& g, int bn, cudaStream_t st) {
    const int nsplit = g.split_k > 1 ? g.split_k : 1;
    const int kb_all = (g.K + BK - 1) / BK;
    const int kb_max = (kb_all + nsplit - 1) / nsplit;               // largest K slice, in k-blocks
    if constexpr (EPI == EPI_GENERIC) {                              // split-K slices: the whole activation slice stays resident too
        if (bn == 16 && kb_max <= 3) return launch_skinny_p<16, EPI, 3>(g, st);
        if (bn == 16 && kb_max <= 6) return launch_skinny_p<16, EPI, 6>(g, st);
        if (bn == 32 && kb_max <= 3) return launch_skinny_p<32, EPI, 3>(g, st);
    }
    if (bn == 16 && kb_max <= 9) return launch_skinny_p<16, EPI, 9>(g, st);
    if (bn == 32 && kb_max <= 9) return launch_skinny_p<32, EPI, 9>(g, st);
    return cudaErrorNotSupported;
}

}  // namespace

// Decode-sized GEMM with resident weights.  Returns cudaErrorNotSupported (nothing launched) when the shape does not
// fit this kernel (more than one M tile, more tiles than SMs, K slice longer than the weight region); the caller then
// uses the persistent ring kernel of gemm_umma.cu.
cudaError_t launch_gemm_skinny(const GemmArgs& g, int epi, int bn, cudaStream_t st) {
    const int num_sms = sm_count();
    const int total = ((g.N + bn - 1) / bn) * (g.split_k > 1 ? g.split_k : 1);
    if (g.M > BM || total > num_sms) return cudaErrorNotSupported;
    switch (epi) {
        case EPI_GENERIC: return launch_skinny_epi<EPI_GENERIC>(g, bn, st);
        case EPI_SWIGLU: return launch_skinny_epi<EPI_SWIGLU>(g, bn, st);
        case EPI_QKV_ROPE: return launch_skinny_epi<EPI_QKV_ROPE>(g, bn, st);
        case EPI_ARGMAX: return cudaErrorNotSupported;            // lm_head has far more tiles than SMs
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mb
