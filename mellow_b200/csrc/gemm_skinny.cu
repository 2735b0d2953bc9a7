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

// CS > 0: "tail" variant (o_proj / down_proj of a decode layer).  The K range is cut over the CS CTAs of a thread-block
// cluster (K slice = cluster rank); the partial accumulators are exchanged through distributed shared memory as a
// reduce-scatter (CTA r finishes rows [r*RPC, (r+1)*RPC) of the tile), and the CTA that owns a row adds the residual,
// writes the new residual stream, the planes of x * gain and the row's partial sum of squares (deferred RMSNorm,
// gemm.cuh).  This replaces the global split-K partial buffer AND the add + RMSNorm kernel that consumed it.
template <int BN, bool SPLIT, int KBMAX, int CS = 0>
struct SkinnyCfg {
    static constexpr uint32_t PLANES = SPLIT ? 2 : 1;
    static constexpr int RPC = CS > 0 ? (BM + CS - 1) / CS : 0;        // tile rows finished by each CTA of the cluster
    static constexpr int RED_LD = BN + 4;                               // floats per staged row (16-byte aligned, conflict-free)
    static constexpr uint32_t RED_BYTES = CS > 0 ? ((uint32_t)(CS * RPC * RED_LD * 4 + RPC * (BN / 4) * 4) + 1023u) / 1024u * 1024u : 0u;
    static constexpr uint32_t B_BYTES = BN * BK * 2;                    // one plane of one k-block
    static constexpr uint32_t B_REGION = KBMAX * PLANES * B_BYTES;
    static constexpr uint32_t A_STAGE = PLANES * A_BYTES;
    static constexpr uint32_t BAR_BYTES = 512;
    static constexpr uint32_t BUDGET = 227u * 1024u - 1024u - BAR_BYTES - B_REGION - RED_BYTES;
    static constexpr int FIT = (int)(BUDGET / A_STAGE);
    static constexpr int STAGES = FIT > KBMAX ? KBMAX : FIT;
    static constexpr uint32_t ACC_N = PLANES * BN;                      // accumulator columns: [a*b_hi | a_hi*b_lo]
    static constexpr uint32_t ACC_COLS = ACC_N <= 32 ? 32 : (ACC_N <= 64 ? 64 : 128);
    static constexpr uint32_t TMEM_COLS = ACC_COLS;
    static_assert(ACC_N <= 128 && (BN % 16) == 0, "N tile");
    static constexpr size_t SMEM = (size_t)B_REGION + (size_t)STAGES * A_STAGE + RED_BYTES + 1024 + BAR_BYTES;
    static_assert(STAGES >= 2, "activation ring does not fit");
    static_assert((B_BYTES % 1024) == 0, "swizzled tiles start on 1024-byte boundaries");
    static_assert((2 * KBMAX + 2 * 8 + 2) * 8 + 8 <= BAR_BYTES, "barrier block too small");
};

template <int BN, int EPI, bool SPLIT, int KBMAX, int CS = 0>
__global__ void __launch_bounds__(kThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const GemmArgs g) {
    using C = SkinnyCfg<BN, SPLIT, KBMAX, CS>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* breg = smem;                                       // [KBMAX][PLANES][BN x 64] weight slice
    unsigned char* areg = smem + C::B_REGION;                         // [STAGES][PLANES][128 x 64] activation ring
    float* red = reinterpret_cast<float*>(areg + (size_t)C::STAGES * C::A_STAGE);   // [CS][RPC][RED_LD] partials sent by the cluster
    uint64_t* bfull = reinterpret_cast<uint64_t*>(areg + (size_t)C::STAGES * C::A_STAGE + C::RED_BYTES);
    uint64_t* afull = bfull + KBMAX;
    uint64_t* aempty = afull + C::STAGES;
    uint64_t* tfull = aempty + C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    uint32_t* trace_slot = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsplit = CS > 0 ? CS : (g.split_k > 1 ? g.split_k : 1);
    const int tiles_n = (g.N + BN - 1) / BN;
    // K split owned by this CTA: the cluster rank in the tail variant (blockIdx.x = tile * CS + rank), else z-major
    const int z = CS > 0 ? (int)(blockIdx.x % (CS > 0 ? CS : 1)) : blockIdx.x / tiles_n;
    const int n0 = (CS > 0 ? (int)(blockIdx.x / (CS > 0 ? CS : 1)) : (blockIdx.x - z * tiles_n)) * BN;
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
        // deferred RMSNorm, consumer side (gemm.cuh): rstd of this row, fetched while the MMAs run
        const float row_scale = (g.ssq_in != nullptr && m < g.M) ? deferred_rstd(g, m) : 1.0f;
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
                if (g.ssq_in != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] *= row_scale;
                }
                const int n = n0 + c0;
                if constexpr (CS > 0) {
                    // reduce-scatter through distributed shared memory: this row's 16 partial sums go to the CTA of the
                    // cluster that finishes the row, into the slot of this CTA's K slice
                    const int owner = m / C::RPC;
                    const uint32_t la = smem_u32(red + ((size_t)(z * C::RPC + (m - owner * C::RPC)) * C::RED_LD + c0));
                    uint32_t ra;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(owner));
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra + 4u * j), "f"(v[j]), "f"(v[j + 1]),
                                     "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                } else if (nsplit > 1) {
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
        if (CS == 0 && nsplit > 1) {
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
    if constexpr (CS > 0 && EPI != EPI_GENERIC) {
        // cluster split-K with a fused epilogue (gate/up + SwiGLU, QKV + RoPE + KV write): the CTA that owns a row sums the
        // CS partial accumulators (fixed order) and runs the row epilogue on 16-column pieces
        constexpr int kP16 = BN / 16;
        constexpr int kItems = C::RPC * kP16;
        static_assert(kItems <= kEpiWarps * 32, "one piece per epilogue thread");
        const int idx = (int)threadIdx.x;
        const int rowl = idx / kP16, c = (idx - rowl * kP16) * 16;
        const int mr = z * C::RPC + rowl, n = n0 + c;
        const bool act = warp < kEpiWarps && idx < kItems && mr < BM && mr < g.M && n < g.N;
        float rc[8], rs[8];
        if (EPI == EPI_QKV_ROPE && act && n + 16 <= g.N) qkv_rope_load(g, mr, n, rc, rs);   // in flight across the cluster barrier
        __syncwarp();
        cluster_sync_all();
        if (act) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) ld4(red + (size_t)rowl * C::RED_LD + c + j, v + j);
#pragma unroll
            for (int zz = 1; zz < CS; ++zz) {
                float t[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) ld4(red + (size_t)(zz * C::RPC + rowl) * C::RED_LD + c + j, t + j);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += t[j];
            }
            if (EPI == EPI_QKV_ROPE && n + 16 <= g.N) epilogue_row16_qkv(g, mr, n, v, rc, rs);
            else epilogue_row16<EPI>(g, mr, n, v);
        }
    } else if constexpr (CS > 0) {
        constexpr int kC4 = BN / 4;                                   // float4 pieces per row of the tile
        constexpr int kItems = C::RPC * kC4;
        static_assert(kItems <= 2 * kEpiWarps * 32, "two pieces per epilogue thread at most");
        float* sqb = red + (size_t)CS * C::RPC * C::RED_LD;          // [RPC][kC4] squares of the finished pieces
        // operands of the rows this CTA finishes that do not come from the cluster: residual and gain, in flight while
        // the cluster barrier is pending (epilogue threads have passed their dependency wait)
        float4 xr[2], gw[2];
        int rowl[2], col[2];
        bool act[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = (int)threadIdx.x + i * kEpiWarps * 32;
            rowl[i] = idx / kC4; col[i] = (idx - rowl[i] * kC4) * 4;
            const int mr = z * C::RPC + rowl[i];
            act[i] = warp < kEpiWarps && idx < kItems && mr < BM && mr < g.M && n0 + col[i] < g.N;
            xr[i] = make_float4(0.f, 0.f, 0.f, 0.f); gw[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (act[i]) {
                if (g.residual) xr[i] = *reinterpret_cast<const float4*>(g.residual + (size_t)mr * g.ldr + n0 + col[i]);
                if (g.norm_w) gw[i] = __ldg(reinterpret_cast<const float4*>(g.norm_w + n0 + col[i]));
            }
        }
        __syncwarp();
        cluster_sync_all();                                           // every partial of this CTA's rows has landed (release / acquire)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (!act[i]) continue;
            const int mr = z * C::RPC + rowl[i];
            float4 a = *reinterpret_cast<const float4*>(red + (size_t)rowl[i] * C::RED_LD + col[i]);
#pragma unroll
            for (int zz = 1; zz < CS; ++zz) {                         // fixed order: deterministic
                const float4 t = *reinterpret_cast<const float4*>(red + (size_t)(zz * C::RPC + rowl[i]) * C::RED_LD + col[i]);
                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            }
            a.x += xr[i].x; a.y += xr[i].y; a.z += xr[i].z; a.w += xr[i].w;
            if (g.out_f32) *reinterpret_cast<float4*>(g.out_f32 + (size_t)mr * g.ldo + n0 + col[i]) = a;
            if (g.out_hi) {
                const size_t o = (size_t)mr * g.ldp + n0 + col[i];
                store_planes2(g.out_hi, g.out_lo, o, gw[i].x * a.x, gw[i].y * a.y);
                store_planes2(g.out_hi, g.out_lo, o + 2, gw[i].z * a.z, gw[i].w * a.w);
            }
            sqb[rowl[i] * kC4 + (col[i] >> 2)] = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
        if (g.ssq_out != nullptr) {
            if (warp < kEpiWarps) asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            const int mr = z * C::RPC + (int)threadIdx.x;
            if ((int)threadIdx.x < C::RPC && mr < BM && mr < g.M) {
                float t = 0.f;
#pragma unroll
                for (int c = 0; c < kC4; ++c) t += (n0 + 4 * c < g.N) ? sqb[threadIdx.x * kC4 + c] : 0.f;
                g.ssq_out[(size_t)(n0 / BN) * g.ssq_ld + mr] = t;
            }
        }
    }
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

// cluster launch (tail / cluster split-K variants): grid = tiles_n x CS, cluster = the CS K slices of one N tile
template <int BN, int EPI, bool SPLIT, int KBMAX, int CS>
cudaError_t launch_cluster(const GemmArgs& g, cudaStream_t st) {
    using C = SkinnyCfg<BN, SPLIT, KBMAX, CS>;
    auto kern = gemm_skinny_kernel<BN, EPI, SPLIT, KBMAX, CS>;
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
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((g.N + BN - 1) / BN) * CS)); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tb_hi, tb_lo, g);
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

// Tail of a decode layer (o_proj / down_proj): cluster split-K with the residual add, the planes of x * gain and the
// deferred-RMSNorm partials in the kernel (SkinnyCfg).  cs = 4: 32-column tiles (o_proj, K = 576 in 3+2+2+2 k-blocks);
// cs = 8: 48-column tiles (down_proj, K = 1536 in 8 x 3 k-blocks).  g.partial / g.split_k are not used.  `shape` = cluster
// size * 100 + tile columns; the other shapes (experiments on the CTA counts, profiles/r2_decode_ab_tail_shapes.jsonl) are
// compiled into lab builds only (MB_BUILD_LAB=1).
template <int BN, int KBMAX, int CS>
cudaError_t launch_tail_p(const GemmArgs& g, cudaStream_t st) {
    const int kb_all = (g.K + BK - 1) / BK;
    if (g.N % BN != 0 || (kb_all + CS - 1) / CS > KBMAX) return cudaErrorNotSupported;
    return g.passes == 3 ? launch_cluster<BN, EPI_GENERIC, true, KBMAX, CS>(g, st) : launch_cluster<BN, EPI_GENERIC, false, KBMAX, CS>(g, st);
}

cudaError_t launch_gemm_tail(const GemmArgs& g, int shape, cudaStream_t st) {
    if (g.M > BM || !g.out_f32 || (g.N & 3) || (g.ldo & 3) || (g.residual && (g.ldr & 3)) || (g.out_hi && (g.ldp & 3)))
        return cudaErrorInvalidValue;
    switch (shape) {                                                  // cluster size * 100 + tile columns
        case 432: return launch_tail_p<32, 3, 4>(g, st);
        case 848: return launch_tail_p<48, 3, 8>(g, st);
#ifdef MB_LAB
        case 332: return launch_tail_p<32, 3, 3>(g, st);
        case 232: return launch_tail_p<32, 5, 2>(g, st);
        case 448: return launch_tail_p<48, 3, 4>(g, st);
        case 648: return launch_tail_p<48, 4, 6>(g, st);
#endif
        default: return cudaErrorNotSupported;
    }
}

// gate/up (+SwiGLU) and QKV (+RoPE + KV write) of a decode layer as cluster split-K GEMMs: K = 576 in 3 slices of 3
// k-blocks (24 MMAs per CTA instead of 72, a third of the activation bytes per CTA), 64- / 32-column tiles, the fused
// epilogue on the rows each CTA of the cluster owns after the distributed-shared-memory reduce-scatter.
// MEASURED SLOWER (profiles/r2_decode_ab_option_matrix.jsonl: 1.30 vs 1.18 ms per step): the 48 clusters of 3 CTAs of
// gate/up do not become resident together (GPCs of 16-20 SMs hold 5-6 such clusters each), so the grid runs in two
// waves (body 11.3 us instead of 5.2).  Compiled into lab builds only (MB_BUILD_LAB=1).
cudaError_t launch_gemm_cluster3(const GemmArgs& g, int epi, cudaStream_t st) {
#ifndef MB_LAB
    (void)g; (void)epi; (void)st;
    return cudaErrorNotSupported;
#else
    const int kb_all = (g.K + BK - 1) / BK;
    if (g.M > BM || (kb_all + 2) / 3 > 3) return cudaErrorNotSupported;
    if (epi == EPI_SWIGLU && g.N % 64 == 0)
        return g.passes == 3 ? launch_cluster<64, EPI_SWIGLU, true, 3, 3>(g, st) : launch_cluster<64, EPI_SWIGLU, false, 3, 3>(g, st);
    if (epi == EPI_QKV_ROPE && g.N % 32 == 0)
        return g.passes == 3 ? launch_cluster<32, EPI_QKV_ROPE, true, 3, 3>(g, st) : launch_cluster<32, EPI_QKV_ROPE, false, 3, 3>(g, st);
    return cudaErrorNotSupported;
#endif
}

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
