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
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue
// (TMEM lane quadrant = warp % 4, two warps per quadrant alternating 32-column chunks).
#include "umma.cuh"

namespace mb {

namespace {

using namespace umma;

// CG = 2: CTA pairs (cta_group::2, umma.cuh).  A 128 x BN tile needs (128 + BN) x 128 B x planes per k-block for
// 3 x 4 x BN/2 tensor cycles: 62.5 B/clk/SM at BN = 256 and 71 at BN = 192, which is the per-SM L2 -> SM rate (ncu: tensor
// pipe 55-70 % active, profiles/r2_lm_prefill_gemms_b128_ncu_full.json).  A pair works on 256 x BN, each CTA loading its
// 128 rows of A and HALF of B: 41.7 / 46.7 B/clk/SM, and the ring gets a third stage.
template <int BN, bool SPLIT, int CG = 1>
struct Cfg {
    static constexpr uint32_t A_TILE = BM * BK * 2;
    static constexpr uint32_t B_BYTES = (BN / CG) * BK * 2;
    static constexpr uint32_t STAGE_BYTES = (SPLIT ? 2 : 1) * (A_TILE + B_BYTES);
    // epilogue staging: one 32-row x 32-column fp32 block (4 KB, 128-byte rows, 16-byte pieces XOR-swizzled by row) per
    // epilogue warp, through which accumulator rows (one per thread after tcgen05.ld) turn into whole 128-byte row
    // segments per 8 lanes for the global loads / stores
    static constexpr uint32_t STG_BYTES = BN >= 32 ? (uint32_t)kEpiWarps * 4096u : 0u;
    static constexpr uint32_t BUDGET = 227u * 1024u - 1024u - 256u - STG_BYTES;
    static constexpr int STAGES_FIT = (int)(BUDGET / STAGE_BYTES);
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr uint32_t ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;                             // two accumulators: MMA of tile i+1
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + STG_BYTES;   // overlaps epilogue of tile i
    static_assert(STAGES >= 2, "tile does not fit");
};

struct TileCoord { int m0, n0, z, kb_begin, KB; };

template <int BN, int CG = 1>
__device__ __forceinline__ TileCoord tile_coord(const GemmArgs& g, int tile, int tiles_n, int per_split, int nsplit) {
    TileCoord t;
    t.z = tile / per_split;
    const int rem = tile - t.z * per_split;
    const int mt = rem / tiles_n;
    t.m0 = mt * BM * CG;                                              // CG = 2: first row of the pair's 256-row tile
    t.n0 = (rem - mt * tiles_n) * BN;
    const int kb_all = (g.K + BK - 1) / BK;
    t.kb_begin = (int)(((long long)kb_all * t.z) / nsplit);
    t.KB = (int)(((long long)kb_all * (t.z + 1)) / nsplit) - t.kb_begin;
    return t;
}


// 4 floats -> 4 bf16 hi (+ 4 bf16 lo) as one 8-byte store per plane
__device__ __forceinline__ void store_planes4(bf16* hi, bf16* lo, size_t idx, float4 x) {
    bf16 h0, l0, h1, l1, h2, l2, h3, l3;
    split_bf16(x.x, h0, l0); split_bf16(x.y, h1, l1); split_bf16(x.z, h2, l2); split_bf16(x.w, h3, l3);
    __nv_bfloat162 a, b;
    a.x = h0; a.y = h1; b.x = h2; b.y = h3;
    *reinterpret_cast<uint2*>(hi + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    if (lo) {
        a.x = l0; a.y = l1; b.x = l2; b.y = l3;
        *reinterpret_cast<uint2*>(lo + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
}

// EPI_GENERIC on one 32-row x 32-column block of the accumulator, coalesced.  `v` holds this thread's row (lane = row,
// already scaled by the deferred rstd).  The block goes through the warp's swizzled staging buffer once; afterwards 8
// lanes own one 128-byte row segment (lane -> row 4i + lane/8, columns 4*(lane%8)..+3), so the residual loads and the
// fp32 / plane stores are whole lines instead of 16-byte pieces of 32 different rows, and bias / gain are one 16-byte
// load per lane and block.  sq[i] accumulates this lane's share of sum(x^2) of row 4i + lane/8.
__device__ __forceinline__ void epilogue_block32_coalesced(const GemmArgs& g, int m_base, int n0, const float* v, float* stg,
                                                           int lane, float* sq) {
    const int cc = lane & 7, rsub = lane >> 3;
    const int n = n0 + cc * 4;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (g.bias) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
    if (g.norm_w) g4 = __ldg(reinterpret_cast<const float4*>(g.norm_w + n));
    float4 res[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m_base + 4 * i + rsub;
        res[i] = (g.residual && m < g.M) ? *reinterpret_cast<const float4*>(g.residual + (size_t)m * g.ldr + n)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rsub;
        const int m = m_base + r;
        float4 x = *reinterpret_cast<const float4*>(stg + r * 32 + ((cc ^ (r & 7)) << 2));
        x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
        if (g.act == ACT_GELU) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
        else if (g.act == ACT_SIGMOID) { x.x = sigmoidf_(x.x); x.y = sigmoidf_(x.y); x.z = sigmoidf_(x.z); x.w = sigmoidf_(x.w); }
        x.x += res[i].x; x.y += res[i].y; x.z += res[i].z; x.w += res[i].w;
        if (m < g.M) {
            if (g.out_f32) *reinterpret_cast<float4*>(g.out_f32 + (size_t)m * g.ldo + n) = x;
            if (g.out_hi) store_planes4(g.out_hi, g.out_lo, (size_t)m * g.ldp + n, make_float4(x.x * g4.x, x.y * g4.y, x.z * g4.z, x.w * g4.w));
            sq[i] += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
    }
    __syncwarp();                                                    // the next block overwrites the staging buffer
}

// EPI_QKV_ROPE on one 32-row x 32-column block (n0 % 32 == 0, so the block lies inside one head of q, k or v), same
// staging as epilogue_block32_coalesced: RoPE, the fp32 queries or their attention-operand planes, the key planes and
// the K / V cache rows are written as whole row segments.  The TRANSPOSED value planes are the exception: there the
// accumulator layout (lane = row = key) already is the contiguous one, so they are stored from `v` directly.
__device__ __forceinline__ void epilogue_block32_qkv_coalesced(const GemmArgs& g, int m_base, int n0, const float* v, float* stg,
                                                               int lane) {
    const int cc = lane & 7, rsub = lane >> 3;
    const int n = n0 + cc * 4;
    const bool is_q = n0 < kHidden;
    const bool is_v = n0 >= kHidden + kKvHeads * kHeadDim;
    const int dpos = g.pos_base + (g.d_pos ? *g.d_pos : 0);
    if (is_v && g.kp_hi) {
        const int m = m_base + lane;
        if (m < g.M) {
            const int b = m / g.rows_per_seq, s = m - b * g.rows_per_seq;
            const int c2 = n0 - kHidden - kKvHeads * kHeadDim;
            const size_t o = (((size_t)b * kKvHeads + (c2 >> 6)) * kHeadDim + (c2 & 63)) * g.vt_ld + s;
#pragma unroll
            for (int j = 0; j < 32; ++j) store_planes1(g.vt_hi, g.vt_lo, o + (size_t)j * g.vt_ld, v[j]);
        }
    }
    float2 cs[8], sn[8];
    int bb[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m_base + 4 * i + rsub;
        m = m < g.M ? m : g.M - 1;
        bb[i] = m / g.rows_per_seq;
        ss[i] = m - bb[i] * g.rows_per_seq;
        if (!is_v) {
            const int t = (dpos + ss[i]) * 32 + ((n & (kHeadDim - 1)) >> 1);
            cs[i] = __ldg(reinterpret_cast<const float2*>(g.rope_cos + t));
            sn[i] = __ldg(reinterpret_cast<const float2*>(g.rope_sin + t));
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rsub;
        const int m = m_base + r;
        float4 x = *reinterpret_cast<const float4*>(stg + r * 32 + ((cc ^ (r & 7)) << 2));
        if (m >= g.M) continue;
        if (!is_v) {                                                 // two rotate-half pairs (pair-interleaved head dims)
            const float a0 = x.x * cs[i].x - x.y * sn[i].x, a1 = x.y * cs[i].x + x.x * sn[i].x;
            const float a2 = x.z * cs[i].y - x.w * sn[i].y, a3 = x.w * cs[i].y + x.z * sn[i].y;
            x = make_float4(a0, a1, a2, a3);
        }
        if (is_q) {
            if (g.q_out) *reinterpret_cast<float4*>(g.q_out + (size_t)m * kHidden + n) = x;
            if (g.qp_hi)
                store_planes4(g.qp_hi, g.qp_lo, (size_t)m * kHidden + n,
                              make_float4(x.x * kQScaleLog2, x.y * kQScaleLog2, x.z * kQScaleLog2, x.w * kQScaleLog2));
            continue;
        }
        const int c2 = n - kHidden - (is_v ? kKvHeads * kHeadDim : 0);
        const int kvh = c2 >> 6, dd = c2 & 63;
        if (!is_v && g.kp_hi)
            store_planes4(g.kp_hi, g.kp_lo, (((size_t)bb[i] * kKvHeads + kvh) * g.rows_per_seq + ss[i]) * kHeadDim + dd, x);
        const size_t row = ((size_t)bb[i] * kKvHeads + kvh) * g.t_max + dpos + ss[i];
        void* base = is_v ? g.v_cache : g.k_cache;
        if (g.kv_fmt == kKvF24) {                                    // 4 values: 8 B of upper halves + 4 mantissa bytes
            unsigned char* rp = reinterpret_cast<unsigned char*>(base) + row * 192;
            const uint32_t u0 = f24_bits(x.x), u1 = f24_bits(x.y), u2 = f24_bits(x.z), u3 = f24_bits(x.w);
            *reinterpret_cast<uint2*>(rp + dd * 2) = make_uint2((u0 >> 16) | (u1 & 0xFFFF0000u), (u2 >> 16) | (u3 & 0xFFFF0000u));
            *reinterpret_cast<uint32_t*>(rp + 128 + dd) =
                ((u0 >> 8) & 0xFFu) | (u1 & 0xFF00u) | ((((u2 >> 8) & 0xFFu) | (u3 & 0xFF00u)) << 16);
        } else if (g.kv_fmt == kKvBf16) {
            store_planes4(reinterpret_cast<bf16*>(base) + row * kHeadDim + dd, nullptr, 0, x);
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + row * kHeadDim + dd) = x;
        }
    }
    __syncwarp();                                                    // the next block overwrites the staging buffer
}

// Persistent: grid = min(#tiles, #SMs); every role walks the same static tile sequence (n fastest, so the CTAs that run
// concurrently share A rows and the whole W in L2).  The smem ring runs continuously across tiles; the accumulator
// is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
template <int BN, int EPI, bool SPLIT, int CG = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const GemmArgs g) {
    using C = Cfg<BN, SPLIT, CG>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;       // [2] accumulator ready
    uint64_t* tempty = tfull + 2;              // [2] accumulator drained by the 4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsplit = g.split_k > 1 ? g.split_k : 1;
    const int tiles_n = (g.N + BN - 1) / BN;
    const int per_split = tiles_n * ((g.M + BM * CG - 1) / (BM * CG));
    const int total = per_split * nsplit;
    // CG = 2: the two CTAs of a cluster walk the same tile sequence; `rank` picks the CTA's 128 rows of A / of the
    // accumulator and its half of the B tile.  Barriers: full[] lives in the LEADER (rank 0: one arrival, the bytes of both
    // CTAs); empty[] and tfull[] exist in both CTAs and are signalled by the multicast commit; tempty[] lives in the leader
    // and counts the epilogue warps of both CTAs.
    const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
    const int tile0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    pdl_trigger();
    unsigned trec = kTraceNone;                                      // thread 0 only
    if (threadIdx.x == 0) {
        trec = trace_open(g.trace, g.trace_id);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], kEpiWarps * CG); mbar_init(&tempty[1], kEpiWarps * CG);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();                                 // the peer must see initialised barriers before it signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            bool first = true;
            // one stage = this CTA's A rows and its share of the B rows; CG = 2: bytes are counted on the leader's barrier
            auto expect = [&](int s) { if (rank == 0) mbar_expect_tx(&full[s], C::STAGE_BYTES * CG); };
            auto load_b = [&](int s, int kb, int n0) {
                unsigned char* st = smem + (size_t)s * C::STAGE_BYTES;
                const int nb = n0 + (int)rank * (BN / CG);
                if (CG == 2) {
                    const uint32_t bar = cluster_addr(&full[s], 0);
                    tma_load_2d_2sm(st + C::A_TILE, &tm_b_hi, bar, kb * BK, nb);
                    if (SPLIT) tma_load_2d_2sm(st + 2 * C::A_TILE + C::B_BYTES, &tm_b_lo, bar, kb * BK, nb);
                } else {
                    tma_load_2d(st + C::A_TILE, &tm_b_hi, &full[s], kb * BK, nb);
                    if (SPLIT) tma_load_2d(st + 2 * C::A_TILE + C::B_BYTES, &tm_b_lo, &full[s], kb * BK, nb);
                }
            };
            auto load_a = [&](int s, int kb, int m0) {
                unsigned char* st = smem + (size_t)s * C::STAGE_BYTES;
                const int ma = m0 + (int)rank * BM;
                if (CG == 2) {
                    const uint32_t bar = cluster_addr(&full[s], 0);
                    tma_load_2d_2sm(st, &tm_a_hi, bar, kb * BK, ma);
                    if (SPLIT) tma_load_2d_2sm(st + C::A_TILE + C::B_BYTES, &tm_a_lo, bar, kb * BK, ma);
                } else {
                    tma_load_2d(st, &tm_a_hi, &full[s], kb * BK, ma);
                    if (SPLIT) tma_load_2d(st + C::A_TILE + C::B_BYTES, &tm_a_lo, &full[s], kb * BK, ma);
                }
            };
            for (int tile = tile0; tile < total; tile += tile_step) {
                const TileCoord t = tile_coord<BN, CG>(g, tile, tiles_n, per_split, nsplit);
                int kb = 0;
                if (first) {
                    // PDL: the weight (B) tiles of the first ring slots never depend on the preceding kernel, so they
                    // are requested before the dependency wait; the activation (A) tiles of those slots follow it.
                    first = false;
                    const int npre = t.KB < C::STAGES ? t.KB : C::STAGES;
                    for (int i = 0; i < npre; ++i) {
                        expect(i);
                        load_b(i, t.kb_begin + i, t.n0);
                    }
                    pdl_wait();
                    if (first_cta()) trace_put(g.trace, trec, g.trace_id, TR_WAITED);
                    for (int i = 0; i < npre; ++i) load_a(i, t.kb_begin + i, t.m0);
                    kb = npre;
                    it = npre;
                }
                for (; kb < t.KB; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait_fast(&empty[s], ph ^ 1);
                    expect(s);
                    load_b(s, t.kb_begin + kb, t.n0);
                    load_a(s, t.kb_begin + kb, t.m0);
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {                              // CG = 2: the leader issues for the pair
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N>>3, M>>4 (M = 256 for a pair)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
            int it = 0, lt = 0;
            for (int tile = tile0; tile < total; tile += tile_step, ++lt) {
                const TileCoord t = tile_coord<BN, CG>(g, tile, tiles_n, per_split, nsplit);
                const int buf = lt & 1;
                mbar_wait(&tempty[buf], ((lt >> 1) & 1) ^ 1);        // epilogue has drained this accumulator (both CTAs of a pair)
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)buf * C::ACC_COLS;
                for (int kb = 0; kb < t.KB; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait_fast(&full[s], ph);
                    tc_fence_after();
                    // descriptor low words (16-byte units); a k-step advances them by 2 (umma.cuh)
                    const uint32_t a_hi = umma_desc_lo(smem_u32(smem + (size_t)s * C::STAGE_BYTES));
                    const uint32_t b_hi = a_hi + (C::A_TILE >> 4);
                    const uint32_t a_lo = b_hi + (C::B_BYTES >> 4);
                    const uint32_t b_lo = a_lo + (C::A_TILE >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t dah = umma_desc_join(a_hi + 2 * k), dbh = umma_desc_join(b_hi + 2 * k);
                        if (CG == 2) {
                            if (k == 0 && kb == 0) umma2_bf16_c<false>(tacc, dah, dbh, idesc);
                            else umma2_bf16_c<true>(tacc, dah, dbh, idesc);
                            if (SPLIT) {
                                umma2_bf16_c<true>(tacc, dah, umma_desc_join(b_lo + 2 * k), idesc);
                                umma2_bf16_c<true>(tacc, umma_desc_join(a_lo + 2 * k), dbh, idesc);
                            }
                        } else {
                            if (k == 0 && kb == 0) umma_bf16_c<false>(tacc, dah, dbh, idesc);
                            else umma_bf16_c<true>(tacc, dah, dbh, idesc);
                            if (SPLIT) {
                                umma_bf16_c<true>(tacc, dah, umma_desc_join(b_lo + 2 * k), idesc);
                                umma_bf16_c<true>(tacc, umma_desc_join(a_lo + 2 * k), dbh, idesc);
                            }
                        }
                    }
                    // frees the smem stage (in both CTAs of a pair) once these MMAs retire
                    if (CG == 2) umma2_commit(&empty[s]); else umma_commit(&empty[s]);
                }
                if (CG == 2) umma2_commit(&tfull[buf]); else umma_commit(&tfull[buf]);   // accumulator complete
            }
        }
    } else {
        // One accumulator row per thread (TMEM lane = tile row).  Columns are drained 32 at a time into registers
        // and finished by the vectorised row epilogue; after the last tcgen05.ld the accumulator goes back to the
        // MMA warp, so the stores of tile i overlap the MMAs of tile i+1.
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                            // the warps of a quadrant take 32-column chunks round-robin
        constexpr int kSub = kEpiWarps / 4;
        pdl_wait();                                                  // residual reads / output writes depend on the predecessor
        // the accumulator goes back to the MMA issuer: a local arrive, or (CG = 2) an arrive on the leader's barrier
        auto release_acc = [&](int buf) {
            if (CG == 2) mbar_arrive_cluster(cluster_addr(&tempty[buf], 0));
            else mbar_arrive(&tempty[buf]);
        };
        int lt = 0;
        for (int tile = tile0; tile < total; tile += tile_step, ++lt) {
            TileCoord t = tile_coord<BN, CG>(g, tile, tiles_n, per_split, nsplit);
            t.m0 += (int)rank * BM;                                  // this CTA's 128 rows of the pair's tile
            const int buf = lt & 1;
            const int m = t.m0 + q * 32 + lane;
            mbar_wait(&tfull[buf], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)buf * C::ACC_COLS + ((uint32_t)(q * 32) << 16);
            constexpr int kChunks = (BN + 31) / 32;
            if (half * 32 >= BN) {                                   // narrow tile: nothing for the second warp to drain
                if (lane == 0) release_acc(buf);
                continue;
            }
            // deferred RMSNorm (gemm.cuh): consumer scale of this row, producer partial of this warp's chunks
            const float rs = (g.ssq_in != nullptr && m < g.M) ? deferred_rstd(g, m) : 1.0f;
            float sq = 0.f;
            // whole 32-column blocks of a plain epilogue leave through the staging buffer as 128-byte row segments
            // (warp-uniform choice); ragged tiles, odd leading dimensions and the other epilogues keep the row path
            bool coalesced = false;
            float sqc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (EPI == EPI_GENERIC && (BN % 32) == 0 && C::STG_BYTES != 0)
                coalesced = !g.epi_rows && nsplit == 1 && t.n0 + BN <= g.N && (!g.out_f32 || (g.ldo & 3) == 0) &&
                            (!g.residual || (g.ldr & 3) == 0) && (!g.out_hi || (g.ldp & 3) == 0);
            if (EPI == EPI_QKV_ROPE && (BN % 32) == 0 && C::STG_BYTES != 0)
                coalesced = !g.epi_rows && nsplit == 1 && t.n0 + BN <= g.N;
            float* stg = reinterpret_cast<float*>(smem + (size_t)C::STAGES * C::STAGE_BYTES + 256) + (warp - 2) * 1024;
#pragma unroll 1
            for (int ci = half; ci < kChunks; ci += kSub) {
                const int c0 = ci * 32;
                float v[32];
                tmem_ld16(tacc + (uint32_t)c0, v);
                if (c0 + 16 < BN) tmem_ld16(tacc + (uint32_t)(c0 + 16), v + 16);
                if (ci + kSub >= kChunks) {                          // this warp has read its share: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(buf);
                }
                if (EPI == EPI_GENERIC && (BN % 32) == 0 && coalesced) {
                    if (g.ssq_in != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= rs;
                    }
                    epilogue_block32_coalesced(g, t.m0 + q * 32, t.n0 + c0, v, stg, lane, sqc);
                } else if (EPI == EPI_QKV_ROPE && (BN % 32) == 0 && coalesced) {
                    if (g.ssq_in != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= rs;
                    }
                    epilogue_block32_qkv_coalesced(g, t.m0 + q * 32, t.n0 + c0, v, stg, lane);
                } else if (m < g.M) {
#pragma unroll
                    for (int h = 0; h < 32; h += 16) {
                        const int n = t.n0 + c0 + h;
                        if (c0 + h < BN && n < g.N) {
                            if (nsplit > 1) {
                                float* pz = g.partial + (size_t)t.z * g.M * g.N + (size_t)m * g.N + n;
#pragma unroll
                                for (int j = 0; j < 16; j += 4)
                                    if (n + j < g.N) st4(pz + j, v + h + j);
                            } else {
                                if (g.ssq_in != nullptr) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[h + j] *= rs;
                                }
                                epilogue_row16<EPI>(g, m, n, v + h);   // leaves the finished values in v (EPI_GENERIC)
                                if (EPI == EPI_GENERIC && g.ssq_out != nullptr) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) sq += (n + j < g.N) ? v[h + j] * v[h + j] : 0.f;
                                }
                            }
                        }
                    }
                }
            }
            if (EPI == EPI_GENERIC && g.ssq_out != nullptr) {
                if (coalesced) {                                     // 8 lanes share a row: fixed-order butterfly, lane % 8 == 0 writes
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float s = sqc[i];
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        s += __shfl_xor_sync(0xffffffffu, s, 4);
                        const int mr = t.m0 + q * 32 + 4 * i + (lane >> 3);
                        if ((lane & 7) == 0 && mr < g.M) g.ssq_out[(size_t)((t.n0 / BN) * kSub + half) * g.ssq_ld + mr] = s;
                    }
                } else if (m < g.M) {
                    g.ssq_out[(size_t)((t.n0 / BN) * kSub + half) * g.ssq_ld + m] = sq;
                }
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();                                 // no CTA of a pair may leave while its peer still uses its smem / TMEM
    else __syncthreads();
    if (threadIdx.x == 0) trace_close(g.trace, trec, g.trace_id);
    if (warp == 1) {
        __syncwarp();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <int BN, int EPI, bool SPLIT, int CG = 1>
cudaError_t launch_one(const GemmArgs& g, cudaStream_t st) {
    using C = Cfg<BN, SPLIT, CG>;
    auto kern = gemm_umma_kernel<BN, EPI, SPLIT, CG>;
    static bool configured[kMaxDevices] = {};
    if (cudaError_t e = ensure_smem(kern, C::SMEM, configured); e != cudaSuccess) return e;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (!make_map(&ta_hi, g.A_hi, g.M, g.K, g.lda, BM) || !make_map(&tb_hi, g.W_hi, g.N, g.K, g.ldw, BN / CG))
        return cudaErrorInvalidValue;
    if (SPLIT) {
        if (!make_map(&ta_lo, g.A_lo, g.M, g.K, g.lda, BM) || !make_map(&tb_lo, g.W_lo, g.N, g.K, g.ldw, BN / CG))
            return cudaErrorInvalidValue;
    } else {
        ta_lo = ta_hi;
        tb_lo = tb_hi;
    }
    const int num_sms = sm_count();
    const long long total = (long long)((g.N + BN - 1) / BN) * ((g.M + BM * CG - 1) / (BM * CG)) * (g.split_k > 1 ? g.split_k : 1);
    if (CG == 1) {
        dim3 grid((unsigned)(total < num_sms ? total : num_sms));
        return launch_k(kern, grid, dim3(kThreads), C::SMEM, st, ta_hi, ta_lo, tb_hi, tb_lo, g);
    }
    // pairs: one cluster of 2 per TPC
    const long long pairs = num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (total < pairs ? total : pairs))); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tb_hi, tb_lo, g);
}

template <int BN, int EPI>
cudaError_t launch_p(const GemmArgs& g, cudaStream_t st) {
    // CTA pairs for the wide tiles of the large (prefill / encoder) GEMMs; GemmArgs::cta_pairs = 0 keeps single CTAs
    if constexpr (BN == 192 || BN == 256) {
        if (g.cta_pairs && g.M >= 1024 && g.split_k <= 1)
            return g.passes == 3 ? launch_one<BN, EPI, true, 2>(g, st) : launch_one<BN, EPI, false, 2>(g, st);
    }
    return g.passes == 3 ? launch_one<BN, EPI, true>(g, st) : launch_one<BN, EPI, false>(g, st);
}

// Tile width.  Every tcgen05.mma streams the 128-row A tile from shared memory (~65 cycles whatever N is), so narrow
// tiles are A-read bound: N = 96 caps the tensor pipe near 74 %, N >= 192 is compute-paced.  N = 576 (o_proj, down,
// projection) and N = 960 (QKV) therefore run as 3 / 5 tiles of 192 columns instead of 6 / 10 of 96.
inline int tile_bn(int M, int N) {
    if (M <= 128 && N <= 4096) return N >= 2048 ? 32 : 16;           // decode-sized shapes the weight-resident kernel did not take
    if (N % 256 == 0 && M >= 1024) return 256;                       // widest tile: A tile re-used over 256 columns
    if (N % 192 == 0 && M >= 1024) return 192;
    if ((N % 128 != 0) && (N % 96 == 0)) return 96;
    return 128;
}

template <int EPI>
cudaError_t launch_epi(const GemmArgs& g, cudaStream_t st) {
    switch (tile_bn(g.M, g.N)) {
        case 16: return launch_p<16, EPI>(g, st);
        case 32: return launch_p<32, EPI>(g, st);
        case 96: return launch_p<96, EPI>(g, st);
        case 192: return launch_p<192, EPI>(g, st);
        case 256: return launch_p<256, EPI>(g, st);
        default: return launch_p<128, EPI>(g, st);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// partial row sums of squares a producer GEMM of this shape writes per row (GemmArgs::ssq_out): two epilogue warps per
// TMEM lane quadrant, each owning every other 32-column chunk of a tile
int gemm_umma_ssq_parts(int M, int N) {
    const int bn = tile_bn(M, N);
    return bn >= 64 ? 2 * ((N + bn - 1) / bn) : 0;
}

cudaError_t launch_gemm_umma(const GemmArgs& g, int epi, cudaStream_t st, bool* handled) {
    *handled = false;
    if (g.ssq_out && (epi != EPI_GENERIC || gemm_umma_ssq_parts(g.M, g.N) == 0 || (g.N & 15) || g.split_k > 1)) return cudaErrorInvalidValue;
    if (g.norm_w && (epi != EPI_GENERIC || (g.N & 15) || (g.ldp & 7))) return cudaErrorInvalidValue;
    if (g.M < 1 || g.K < BK || (g.K % 8) != 0 || (g.lda % 8) != 0 || (g.ldw % 8) != 0) return cudaSuccess;
    if (g.split_k > 1 && (epi != EPI_GENERIC || !g.partial || (g.N & 1))) return cudaErrorInvalidValue;
    if (!aligned16(g.A_hi) || !aligned16(g.W_hi) || (g.passes == 3 && (!aligned16(g.A_lo) || !aligned16(g.W_lo)))) return cudaSuccess;
    if (!encode_fn()) return cudaSuccess;
    cudaError_t e;
    switch (epi) {
        case EPI_GENERIC: e = launch_epi<EPI_GENERIC>(g, st); break;
        case EPI_SWIGLU: e = launch_epi<EPI_SWIGLU>(g, st); break;
        case EPI_QKV_ROPE: e = launch_epi<EPI_QKV_ROPE>(g, st); break;
        case EPI_ARGMAX:
            if ((g.N & 15) || !g.cand_val || !g.cand_idx) return cudaErrorInvalidValue;
            e = launch_epi<EPI_ARGMAX>(g, st);
            break;
        default: return cudaErrorInvalidValue;
    }
    *handled = (e == cudaSuccess);
    return e;
}

}  // namespace mb
