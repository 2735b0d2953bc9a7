// Causal prefill attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands by TMA).
// SmolLM2: 9 query heads, 3 kv heads (GQA), head_dim 64; reference path: transformers modeling_llama.py:276-285 called
// cache-less from mellow/wrapper.py:217 (pure causal mask, pads attended).
//
// Operands.  The QKV GEMM epilogue (gemm.cuh, EPI_QKV_ROPE) leaves, next to the KV cache, bf16 hi/lo PLANES of the roped
// queries (scaled by 64^-0.5 * log2 e: the softmax runs on exp2), the roped keys and the TRANSPOSED values, so this kernel converts nothing: every tile is
// one cp.async.bulk.tensor into 128B-swizzled shared memory, consumed by UMMA descriptors.
//   * the M = 128 rows of a tile are 42 consecutive queries x the 3 query heads that share a kv head (row = 3*s + hh:
//     a 4-D tensor map over [M][kv head][hh][64] walks them), so K / V are streamed once for the three heads and only
//     2 of the 128 rows are padding (a per-head 128-query tiling would waste 31 % of S = 389);
//   * S = Q K^T: M 128 x N 64 keys x K 64, accumulator DOUBLE-BUFFERED in TMEM columns [0,64) / [64,128): S of tile
//     i+1 is issued before P V of tile i, so it is complete when the softmax warps come back for it;
//   * softmax: EIGHT warps, two per TMEM lane quadrant: a thread owns one accumulator row and one half (32 keys) of
//     the tile.  Its scores are read ONCE (two tcgen05.ld in flight, one wait) and stay in registers for the maximum,
//     the exp2 and the sums (four independent chains each); the two halves of a row exchange their partial maxima
//     through shared memory (one 64-thread named barrier per tile) and keep separate partial sums until the end.  P is
//     written as bf16 hi/lo into shared memory in the K-major 128B-swizzle layout (chunk ^ (row & 7)) the next MMA reads
//     as its A operand.  (ncu on the four-warp version: the softmax warps ran at ~0.2 instructions per cycle each, two
//     per scheduler -- a latency problem, so the row work is spread over twice as many warps.)
//   * O += P V: M 128 x N 64 x K 64 keys, B operand = the V^T tile (keys contiguous), accumulator in TMEM columns
//     [128,192); when the running maximum moves, the softmax warps rescale O in TMEM (tcgen05.ld / tcgen05.st);
//   * split policy: every contraction is 3 MMAs (hi*hi + hi*lo + lo*hi) like the GEMMs.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = softmax / epilogue
// (TMEM lane quadrant = warp % 4, key / head-dim half = (warp - 2) / 4).  A CTA needs 98 KB of shared memory and 256
// TMEM columns, so TWO CTAs share an SM.
// Causal work per (batch, kv head): 40 tile visits of 64 keys for S = 389.
#include "kernels.cuh"
#include "umma.cuh"

namespace mb {

namespace {

using namespace umma;

constexpr int kQPT = 42;                       // queries per row tile (x 3 heads = 126 rows)
constexpr int kKT = 64;                        // keys per tile = one 128-byte swizzle row of the PV operands
constexpr uint32_t kTile = 128 * 128;          // 16 KB: 128 rows x 64 bf16 (Q tile, P tile)
constexpr uint32_t kTileK = 64 * 128;          // 8 KB: 64 keys x 64 head dims (K tile), 64 head dims x 64 keys (V^T tile)
constexpr uint32_t kQBytes = 64 * 3 * kQPT * 2;
constexpr int kAttnThreads = 320;
constexpr int kSoftmaxWarps = 8;
constexpr int kAttnTmemCols = 256;             // S0 [0,64), S1 [64,128), O0 [128,192), O1 [192,256)

struct AttnUmmaArgs {
    int S, B;
    bf16* out_hi; bf16* out_lo;                // [B*S][576]
};

template <bool SPLIT>
struct AttnCfg {
    static constexpr uint32_t P = SPLIT ? 2 : 1;
    static constexpr uint32_t OFF_Q = 0;
    static constexpr uint32_t OFF_K = OFF_Q + P * kTile;
    static constexpr uint32_t OFF_V = OFF_K + P * kTileK;
    static constexpr uint32_t OFF_P = OFF_V + P * kTileK;
    static constexpr uint32_t OFF_XM = OFF_P + P * kTile;           // [2 tiles in flight][2 halves][128 rows] partial row maxima
    static constexpr uint32_t OFF_BAR = OFF_XM + 2 * 2 * 128 * 4;
    static constexpr size_t SMEM = OFF_BAR + 128 + 1024;
};

// The softmax warps are the kernel's critical resource (ncu: 44 % issue-active, tensor pipe 19 %), so the per-element
// work is kept to a handful of instructions: one MUFU for exp2 (arguments <= 0, results in [0,1]: the approximate
// instruction's 2^-22 relative error is far inside the bf16x2 operand precision), packed bf16x2 conversions for the
// hi / lo split.
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo_elem, float hi_elem) {      // {hi_elem, lo_elem} -> one register
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void pack8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        h[j] = cvt_bf16x2(a, b);
        const float ra = a - __uint_as_float(h[j] << 16), rb = b - __uint_as_float(h[j] & 0xFFFF0000u);
        l[j] = cvt_bf16x2(ra, rb);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Work item = (query tile, kv head, batch row).  The grid is PERSISTENT (two CTAs per SM): CTA c takes items c, c + grid,
// c + 2 grid, ... of a list ordered longest tile first, so every CTA sees the same mix of tile lengths.  What the
// in-kernel clock showed for one-CTA-per-item (14.5 us per CTA): 3.9 us from entry to the first S (barrier / TMEM set-up,
// Q and K in flight), 7.1 us in the tile loop, 1.0 us for the last P V and 2.2 us of output stores -- half of a CTA's life
// outside the loop.  Here barriers and TMEM are set up once, and the pipelines run ACROSS items: the Q / K tiles of the
// next item are requested as soon as the last S of the current item has retired, its first S is issued right behind the
// last P V, and O is double-buffered in TMEM so that the output stores of item i overlap the first tiles of item i + 1.
struct AttnItem { int s0, s_last, n_kt, kvh, b; };
__device__ __forceinline__ AttnItem attn_item(int idx, int n_tiles, int S, int B) {
    AttnItem w;
    const int per_tile = kKvHeads * B;
    const int tile = n_tiles - 1 - idx / per_tile;               // long tiles (late queries) first
    const int rem = idx - (idx / per_tile) * per_tile;
    w.b = rem / kKvHeads;
    w.kvh = rem - w.b * kKvHeads;
    w.s0 = tile * kQPT;
    w.s_last = min(S - 1, w.s0 + kQPT - 1);
    w.n_kt = w.s_last / kKT + 1;                                  // causal: keys 0..s_last
    return w;
}

template <bool SPLIT>
__global__ void __launch_bounds__(kAttnThreads, 2)
prefill_attention_umma_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                              const __grid_constant__ CUtensorMap tm_k_hi, const __grid_constant__ CUtensorMap tm_k_lo,
                              const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                              const AttnUmmaArgs a) {
    using C = AttnCfg<SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* q_s = smem + C::OFF_Q;      // [P][128 x 64]     rows = 3*s_local + hh
    unsigned char* k_s = smem + C::OFF_K;      // [P][64 keys x 64]
    unsigned char* v_s = smem + C::OFF_V;      // [P][64 dims x 64 keys]
    unsigned char* p_s = smem + C::OFF_P;      // [P][128 rows x 64 keys]; between items: staging of the output rows
    float* xm = reinterpret_cast<float*>(smem + C::OFF_XM);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t *q_full = bars, *k_full = bars + 1, *v_full = bars + 2, *s_full = bars + 3 /* [2] */, *p_full = bars + 5, *o_full = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (a.S + kQPT - 1) / kQPT;
    const int n_items = n_tiles * kKvHeads * a.B;

    pdl_trigger();
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1);
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(p_full, kSoftmaxWarps); mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v_hi) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kAttnTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: S of global tile g in [64 (g & 1), +64), O of item n in [128 + 64 (n & 1), +64)
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

    // Every role walks the same item list and counts the same GLOBAL tile index g (tiles of all its items in order):
    // barrier parities follow g (k_full, v_full, p_full, o_full: g & 1; s_full[g & 1]: (g >> 1) & 1) and the item count n
    // (q_full: n & 1), so the pipelines never drain between items.
    if (warp == 0) {
        if (elect_one()) {
            pdl_wait();                                          // the planes are written by the preceding QKV GEMM
            int g = 0, n = 0;
            for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
                const AttnItem w = attn_item(idx, n_tiles, a.S, a.B);
                for (int it = 0; it < w.n_kt; ++it, ++g) {
                    if (g > 0) mbar_wait(&s_full[(g - 1) & 1], ((g - 1) >> 1) & 1);   // S(g-1) retired: K tile (and, at it = 0, Q tile) free
                    if (it == 0) {
                        mbar_expect_tx(q_full, C::P * kQBytes);
                        tma_load_4d(q_s, &tm_q_hi, q_full, 0, 0, w.kvh, w.b * a.S + w.s0);
                        if (SPLIT) tma_load_4d(q_s + kTile, &tm_q_lo, q_full, 0, 0, w.kvh, w.b * a.S + w.s0);
                    }
                    mbar_expect_tx(k_full, C::P * kTileK);
                    tma_load_3d(k_s, &tm_k_hi, k_full, 0, it * kKT, w.b * kKvHeads + w.kvh);
                    if (SPLIT) tma_load_3d(k_s + kTileK, &tm_k_lo, k_full, 0, it * kKT, w.b * kKvHeads + w.kvh);
                    if (g > 0) mbar_wait(o_full, (g - 1) & 1);   // P V(g-1) retired: the V^T tile is free
                    mbar_expect_tx(v_full, C::P * kTileK);
                    tma_load_3d(v_s, &tm_v_hi, v_full, it * kKT, 0, w.b * kKvHeads + w.kvh);
                    if (SPLIT) tma_load_3d(v_s + kTileK, &tm_v_lo, v_full, it * kKT, 0, w.b * kKvHeads + w.kvh);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 64, M = 128 (the same for S and for O)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t qh = umma_desc_lo(smem_u32(q_s)), ql = qh + (kTile >> 4);
            const uint32_t kh = umma_desc_lo(smem_u32(k_s)), kl = kh + (kTileK >> 4);
            const uint32_t vh = umma_desc_lo(smem_u32(v_s)), vl = vh + (kTileK >> 4);
            const uint32_t ph = umma_desc_lo(smem_u32(p_s)), pl = ph + (kTile >> 4);
            auto issue_s = [&](int g, int n, bool first) {       // S(g) = (Q/8) K(g)^T over the 64 head dims -> S buffer g & 1
                if (first) mbar_wait(q_full, n & 1);
                mbar_wait(k_full, g & 1);
                tc_fence_after();
                const uint32_t ts = tmem_s + 64u * (uint32_t)(g & 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t dqh = umma_desc_join(qh + 2 * k), dkh = umma_desc_join(kh + 2 * k);
                    umma_bf16(ts, dqh, dkh, idesc, k > 0 ? 1u : 0u);
                    if (SPLIT) {
                        umma_bf16(ts, dqh, umma_desc_join(kl + 2 * k), idesc, 1u);
                        umma_bf16(ts, umma_desc_join(ql + 2 * k), dkh, idesc, 1u);
                    }
                }
                umma_commit(&s_full[g & 1]);
            };
            auto issue_pv = [&](int g, int n, bool first) {      // O(n) (+)= P(g) V(g) over the 64 keys of the tile
                mbar_wait(v_full, g & 1);
                mbar_wait(p_full, g & 1);                        // P written (and O rescaled) by the softmax warps
                tc_fence_after();
                const uint32_t to = tmem_o + 64u * (uint32_t)(n & 1);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dph = umma_desc_join(ph + 2 * ks), dvh = umma_desc_join(vh + 2 * ks);
                    umma_bf16(to, dph, dvh, idesc, (!first || ks > 0) ? 1u : 0u);
                    if (SPLIT) {
                        umma_bf16(to, dph, umma_desc_join(vl + 2 * ks), idesc, 1u);
                        umma_bf16(to, umma_desc_join(pl + 2 * ks), dvh, idesc, 1u);
                    }
                }
                umma_commit(o_full);
            };
            int g = 0, n = 0;
            if ((int)blockIdx.x < n_items) issue_s(0, 0, true);
            for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
                const AttnItem w = attn_item(idx, n_tiles, a.S, a.B);
                for (int it = 0; it < w.n_kt; ++it, ++g) {
                    // S runs one tile ahead of P V.  Inside an item S(g+1) goes first (its buffer was drained by the softmax
                    // of tile g-1, its K tile was requested when S(g) retired).  At the end of an item the last P V goes
                    // first: the next item's Q / K were only requested when S(g) retired and must not hold it up.
                    if (it + 1 < w.n_kt) {
                        issue_s(g + 1, n, false);
                        issue_pv(g, n, it == 0);
                    } else {
                        issue_pv(g, n, it == 0);
                        if (idx + (int)gridDim.x < n_items) issue_s(g + 1, n + 1, true);
                    }
                }
            }
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;                        // which 32 keys of a tile / which 32 head dims of O
        constexpr int kHK = kKT / 2;
        const int r = quad * 32 + lane;                          // accumulator row = TMEM lane
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        pdl_wait();                                              // the output planes are an operand of the preceding GEMM
        int g = 0, n = 0, x = 0;                                 // x: exchanges through xm so far (double-buffered by parity)
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x, ++n) {
            const AttnItem w = attn_item(idx, n_tiles, a.S, a.B);
            const int sq = w.s0 + r / 3;
            const bool row_ok = r < 3 * kQPT && sq < a.S;
            const int s_eff = row_ok ? sq : w.s0;                // padding rows follow the tile's first query (finite values)
            // tcgen05.ld is warp-collective: loop bounds follow the warp's LAST row, the per-lane causal mask is applied inside
            const int s_hi = min(w.s_last, w.s0 + (quad * 32 + 31) / 3);
            const int s_lo = w.s0 + (quad * 32) / 3;             // the warp's FIRST query: keys up to it need no causal mask
            const uint32_t to = tmem_o + 64u * (uint32_t)(n & 1) + lane_addr + (uint32_t)(half * 32);
            float m_run = -INFINITY, l_run = 0.f;                // l_run: this half's share of the row sum
            for (int it = 0; it < w.n_kt; ++it, ++g) {
                mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc_fence_after();
                const int key0 = it * kKT + half * kHK;          // first key of this thread's half of the tile
                const bool none = key0 > s_hi;                   // warp-uniform: no row of this warp sees these keys
                const bool full = key0 + kHK - 1 <= s_lo;        // warp-uniform: every row sees every key, no mask
                float v[kHK];
                float mt = -INFINITY;
                if (!none) {
                    // the score half-row in registers: two loads in flight, then the waits (each names its registers, so
                    // nothing that uses them can be scheduled ahead of it)
                    uint32_t rr[kHK];
                    const uint32_t ts = tmem_s + 64u * (uint32_t)(g & 1) + (uint32_t)(half * kHK) + lane_addr;
#pragma unroll
                    for (int g4 = 0; g4 < kHK / 16; ++g4) tmem_ld16_issue(ts + 16u * g4, rr + 16 * g4);
#pragma unroll
                    for (int g4 = 0; g4 < kHK / 16; ++g4) tmem_ld16_wait(rr + 16 * g4);
#pragma unroll
                    for (int j = 0; j < kHK; ++j) v[j] = __uint_as_float(rr[j]);
                    if (!full) {
#pragma unroll
                        for (int j = 0; j < kHK; ++j) v[j] = key0 + j <= s_eff ? v[j] : -INFINITY;
                    }
                    float mx[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
                    for (int j = 4; j < kHK; ++j) mx[j & 3] = fmaxf(mx[j & 3], v[j]);
                    mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                }
                // both halves of the row must use the same maximum: exchange the partial maxima (double-buffered by
                // exchange parity; the 64 threads of the two warps of this lane quadrant meet at named barrier 1 + quad)
                float* xt = xm + (x & 1) * 256;
                ++x;
                xt[half * 128 + r] = mt;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
                const float m_new = fmaxf(m_run, fmaxf(mt, xt[(half ^ 1) * 128 + r]));   // finite: key 0 <= every query
                const float alpha = ex2_fast(m_run - m_new);     // scores are in log2 units (the q planes carry log2(e) / 8)
                float lsum = 0.f;
                if (!none) {
                    float sm4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < kHK; ++j) { v[j] = ex2_fast(v[j] - m_new); sm4[j & 3] += v[j]; }   // masked: exp2(-inf) = 0
                    lsum = (sm4[0] + sm4[1]) + (sm4[2] + sm4[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < kHK; ++j) v[j] = 0.f;
                }
                m_run = m_new;
                l_run = l_run * alpha + lsum;
                if (g > 0) {
                    // P V of the previous tile has retired: the P tile may be overwritten and O is stable.  (S of THIS tile
                    // was issued before that P V, so s_full does not imply it.)
                    mbar_wait(o_full, (g - 1) & 1);
                    tc_fence_after();
                }
                if (it > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {   // rescale this warp's 32 head dims of the running output
#pragma unroll
                    for (int c0 = 0; c0 < kHeadDim / 2; c0 += 16) {
                        float o[16];
                        tmem_ld16(to + (uint32_t)c0, o);
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] *= alpha;
                        tmem_st16(to + (uint32_t)c0, o);
                    }
                    tmem_st_wait();
                }
                // P = exp2(S - m) as bf16 hi/lo in the K-major 128B-swizzle layout of the PV A operand
#pragma unroll
                for (int cc = 0; cc < kHK / 8; ++cc) {           // 16-byte chunk (8 keys) of the 128-byte row
                    const int ch = half * (kHK / 8) + cc;
                    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
                    uint4 hi, lo;
                    pack8(v + 8 * cc, hi, lo);
                    *reinterpret_cast<uint4*>(p_s + off) = hi;
                    if (SPLIT) *reinterpret_cast<uint4*>(p_s + kTile + off) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P (generic-proxy stores) -> UMMA (async proxy)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
            }
            // row sum = the two halves' shares (same reference maximum); exchanged like the maxima
            float* xl = xm + (x & 1) * 256;
            ++x;
            xl[half * 128 + r] = l_run;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
            const float inv = 1.0f / (l_run + xl[(half ^ 1) * 128 + r]);
            mbar_wait(o_full, (g - 1) & 1);                      // last P V of the item: O complete, the P tile is free
            tc_fence_after();
            // Output: this thread's 32 head dims of its row, normalised, as bf16 hi / lo.  The rows go through the (idle)
            // P tile -- 128-byte rows, 16-byte chunks XOR-swizzled by row like P itself -- so that 4 lanes write one
            // 64-byte run of a row and a store instruction covers 8 rows, instead of 32 scattered 16-byte pieces.
#pragma unroll
            for (int c0 = 0; c0 < kHeadDim / 2; c0 += 16) {
                float o[16];
                tmem_ld16(to + (uint32_t)c0, o);
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] *= inv;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ch = half * 4 + (c0 >> 3) + e;
                    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
                    uint4 hi, lo;
                    pack8(o + 8 * e, hi, lo);
                    *reinterpret_cast<uint4*>(p_s + off) = hi;
                    if (SPLIT) *reinterpret_cast<uint4*>(p_s + kTile + off) = lo;
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // row of the tile; this lane's 16-byte chunk of this warp's half.  The two rows a quarter-warp reads are 4
                // apart, so the row XOR puts their 64-byte halves in different bank groups.
                const int j8 = lane >> 2;
                const int rl = quad * 32 + 8 * i + (((j8 & 1) << 2) | (j8 >> 1));
                const int ch = half * 4 + (lane & 3);
                const int sr = w.s0 + rl / 3;
                if (rl < 3 * kQPT && sr < a.S) {
                    const uint32_t off = (uint32_t)rl * 128u + (uint32_t)((ch ^ (rl & 7)) << 4);
                    const size_t ob = ((size_t)w.b * a.S + sr) * kHidden + (size_t)(w.kvh * 3 + rl % 3) * kHeadDim + ch * 8;
                    *reinterpret_cast<uint4*>(a.out_hi + ob) = *reinterpret_cast<const uint4*>(p_s + off);
                    if (SPLIT) *reinterpret_cast<uint4*>(a.out_lo + ob) = *reinterpret_cast<const uint4*>(p_s + kTile + off);
                }
            }
            __syncwarp();                                        // the next item's P stores reuse the staging rows
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAttnTmemCols) : "memory");
    }
}

template <bool SPLIT>
cudaError_t launch_attn(const PrefillAttnPlanes& p, int B, int S, bf16* out_hi, bf16* out_lo, cudaStream_t st) {
    using C = AttnCfg<SPLIT>;
    auto kern = prefill_attention_umma_kernel<SPLIT>;
    static bool configured[kMaxDevices] = {};
    if (cudaError_t e = ensure_smem(kern, C::SMEM, configured); e != cudaSuccess) return e;
    const long long M = (long long)B * S;
    CUtensorMap tq[2], tk[2], tv[2];
    for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {
        {   // queries: [M][kv head 3][hh 3][64]
            const long long dims[4] = {kHeadDim, 3, kKvHeads, M};
            const long long strides[4] = {1, kHeadDim, 3 * kHeadDim, kHidden};
            const int box[4] = {kHeadDim, 3, 1, kQPT};
            if (!make_map_nd(&tq[pl], pl ? p.qp_lo : p.qp_hi, 4, dims, strides, box)) return cudaErrorInvalidValue;
        }
        {   // keys: [B*3][S][64]
            const long long dims[3] = {kHeadDim, S, (long long)B * kKvHeads};
            const long long strides[3] = {1, kHeadDim, (long long)S * kHeadDim};
            const int box[3] = {kHeadDim, kKT, 1};               // 64 keys x 64 head dims
            if (!make_map_nd(&tk[pl], pl ? p.kp_lo : p.kp_hi, 3, dims, strides, box)) return cudaErrorInvalidValue;
        }
        {   // values, transposed: [B*3][64][S (row pitch vt_ld)]
            const long long dims[3] = {S, kHeadDim, (long long)B * kKvHeads};
            const long long strides[3] = {1, p.vt_ld, (long long)kHeadDim * p.vt_ld};
            const int box[3] = {64, kHeadDim, 1};
            if (!make_map_nd(&tv[pl], pl ? p.vt_lo : p.vt_hi, 3, dims, strides, box)) return cudaErrorInvalidValue;
        }
    }
    if (!SPLIT) { tq[1] = tq[0]; tk[1] = tk[0]; tv[1] = tv[0]; }
    AttnUmmaArgs a{S, B, out_hi, out_lo};
    const long long items = (long long)((S + kQPT - 1) / kQPT) * kKvHeads * B;
    const long long slots = 2LL * sm_count();                    // two resident CTAs per SM (shared memory, TMEM columns)
    dim3 grid((unsigned)(items < slots ? items : slots));
    return launch_k(kern, grid, dim3(kAttnThreads), C::SMEM, st, tq[0], tq[1], tk[0], tk[1], tv[0], tv[1], a);
}

}  // namespace

cudaError_t launch_prefill_attention_umma(const PrefillAttnPlanes& p, int B, int S, bf16* out_hi, bf16* out_lo,
                                          cudaStream_t st) {
    if (!encode_fn() || (p.vt_ld & 7) || S < 1) return cudaErrorInvalidValue;
    return out_lo ? launch_attn<true>(p, B, S, out_hi, out_lo, st) : launch_attn<false>(p, B, S, out_hi, out_lo, st);
}

}  // namespace mb
