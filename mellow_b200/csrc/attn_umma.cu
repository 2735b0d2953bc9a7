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
constexpr int kAttnTmemCols = 256;             // S0 [0,64), S1 [64,128), O [128,192)

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
    unsigned char* p_s = smem + C::OFF_P;      // [P][128 rows x 64 keys]
    float* xm = reinterpret_cast<float*>(smem + C::OFF_XM);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t *q_full = bars, *k_full = bars + 1, *v_full = bars + 2, *s_full = bars + 3 /* [2] */, *p_full = bars + 5, *o_full = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (a.S + kQPT - 1) / kQPT;
    const int tile = n_tiles - 1 - (int)blockIdx.x;             // long tiles (late queries) first
    const int kvh = blockIdx.y, b = blockIdx.z;
    const int s0 = tile * kQPT;
    const int s_last = min(a.S - 1, s0 + kQPT - 1);
    const int n_kt = s_last / kKT + 1;                          // causal: keys 0..s_last

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
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;   // S buffer of tile i: tmem_s + 64 * (i & 1)

    if (warp == 0) {
        if (elect_one()) {
            pdl_wait();                                          // the planes are written by the preceding QKV GEMM
            mbar_expect_tx(q_full, C::P * kQBytes);
            tma_load_4d(q_s, &tm_q_hi, q_full, 0, 0, kvh, b * a.S + s0);
            if (SPLIT) tma_load_4d(q_s + kTile, &tm_q_lo, q_full, 0, 0, kvh, b * a.S + s0);
            auto load_k = [&](int it) {
                mbar_expect_tx(k_full, C::P * kTileK);
                tma_load_3d(k_s, &tm_k_hi, k_full, 0, it * kKT, b * kKvHeads + kvh);
                if (SPLIT) tma_load_3d(k_s + kTileK, &tm_k_lo, k_full, 0, it * kKT, b * kKvHeads + kvh);
            };
            auto load_v = [&](int it) {
                mbar_expect_tx(v_full, C::P * kTileK);
                tma_load_3d(v_s, &tm_v_hi, v_full, it * kKT, 0, b * kKvHeads + kvh);
                if (SPLIT) tma_load_3d(v_s + kTileK, &tm_v_lo, v_full, it * kKT, 0, b * kKvHeads + kvh);
            };
            load_k(0);
            load_v(0);
            for (int it = 0; it + 1 < n_kt; ++it) {
                mbar_wait(&s_full[it & 1], (it >> 1) & 1);       // S of tile `it` has retired: the K tile is free
                load_k(it + 1);
                mbar_wait(o_full, it & 1);                       // PV of tile `it` has retired: the V^T tile is free
                load_v(it + 1);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 64, M = 128 (the same for S and for O)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t idesc_s = idesc, idesc_o = idesc;
            const uint32_t qh = umma_desc_lo(smem_u32(q_s)), ql = qh + (kTile >> 4);
            const uint32_t kh = umma_desc_lo(smem_u32(k_s)), kl = kh + (kTileK >> 4);
            const uint32_t vh = umma_desc_lo(smem_u32(v_s)), vl = vh + (kTileK >> 4);
            const uint32_t ph = umma_desc_lo(smem_u32(p_s)), pl = ph + (kTile >> 4);
            mbar_wait(q_full, 0);
            auto issue_s = [&](int it) {                         // S(it) = (Q/8) K(it)^T over the 64 head dims -> S buffer it & 1
                mbar_wait(k_full, it & 1);
                tc_fence_after();
                const uint32_t ts = tmem_s + 64u * (uint32_t)(it & 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t dqh = umma_desc_join(qh + 2 * k), dkh = umma_desc_join(kh + 2 * k);
                    umma_bf16(ts, dqh, dkh, idesc_s, k > 0 ? 1u : 0u);
                    if (SPLIT) {
                        umma_bf16(ts, dqh, umma_desc_join(kl + 2 * k), idesc_s, 1u);
                        umma_bf16(ts, umma_desc_join(ql + 2 * k), dkh, idesc_s, 1u);
                    }
                }
                umma_commit(&s_full[it & 1]);
            };
            issue_s(0);
            for (int it = 0; it < n_kt; ++it) {
                // S of the NEXT tile first: its buffer was drained by the softmax of tile it-1 (p_full(it-1), waited for
                // below in the previous iteration), its K tile was requested when S(it) retired
                if (it + 1 < n_kt) issue_s(it + 1);
                mbar_wait(v_full, it & 1);
                mbar_wait(p_full, it & 1);                       // P written (and O rescaled) by the softmax warps
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                 // O += P V over the 64 keys of the tile
                    const uint64_t dph = umma_desc_join(ph + 2 * ks), dvh = umma_desc_join(vh + 2 * ks);
                    umma_bf16(tmem_o, dph, dvh, idesc_o, (it > 0 || ks > 0) ? 1u : 0u);
                    if (SPLIT) {
                        umma_bf16(tmem_o, dph, umma_desc_join(vl + 2 * ks), idesc_o, 1u);
                        umma_bf16(tmem_o, umma_desc_join(pl + 2 * ks), dvh, idesc_o, 1u);
                    }
                }
                umma_commit(o_full);
            }
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;                        // which 32 keys of a tile / which 32 head dims of O
        constexpr int kHK = kKT / 2;
        const int r = quad * 32 + lane;                          // accumulator row = TMEM lane
        const int hh = r % 3;
        const int sq = s0 + r / 3;
        const bool row_ok = r < 3 * kQPT && sq < a.S;
        const int s_eff = row_ok ? sq : s0;                      // padding rows follow the tile's first query (finite values)
        // tcgen05.ld is warp-collective: loop bounds follow the warp's LAST row, the per-lane causal mask is applied inside
        const int s_hi = min(s_last, s0 + (quad * 32 + 31) / 3);
        const int s_lo = s0 + (quad * 32) / 3;                   // the warp's FIRST query: keys up to it need no causal mask
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        pdl_wait();                                              // the output planes are an operand of the preceding GEMM
        float m_run = -INFINITY, l_run = 0.f;                    // l_run: this half's share of the row sum
        for (int it = 0; it < n_kt; ++it) {
            mbar_wait(&s_full[it & 1], (it >> 1) & 1);
            tc_fence_after();
            const int key0 = it * kKT + half * kHK;              // first key of this thread's half of the tile
            const bool none = key0 > s_hi;                       // warp-uniform: no row of this warp sees these keys
            const bool full = key0 + kHK - 1 <= s_lo;            // warp-uniform: every row sees every key, no mask
            float v[kHK];
            float mt = -INFINITY;
            if (!none) {
                // the score half-row in registers: two loads in flight, then the waits (each names its registers, so
                // nothing that uses them can be scheduled ahead of it)
                uint32_t rr[kHK];
                const uint32_t ts = tmem_s + 64u * (uint32_t)(it & 1) + (uint32_t)(half * kHK) + lane_addr;
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
            // both halves of the row must use the same maximum: exchange the partial maxima (double-buffered by tile
            // parity; the 64 threads of the two warps of this lane quadrant meet at named barrier 1 + quad)
            float* xt = xm + (it & 1) * 256;
            xt[half * 128 + r] = mt;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
            const float m_new = fmaxf(m_run, fmaxf(mt, xt[(half ^ 1) * 128 + r]));   // finite: key 0 <= every query
            const float alpha = ex2_fast(m_run - m_new);         // scores are in log2 units (the q planes carry log2(e) / 8)
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
            if (it > 0) {
                // P V of the previous tile has retired: the P tile may be overwritten and O is stable.  (S of THIS tile
                // was issued before that P V, so s_full no longer implies it.)
                mbar_wait(o_full, (it - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {    // rescale this warp's 32 head dims of the running output
#pragma unroll
                    for (int c0 = 0; c0 < kHeadDim / 2; c0 += 16) {
                        float o[16];
                        tmem_ld16(tmem_o + lane_addr + (uint32_t)(half * 32 + c0), o);
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] *= alpha;
                        tmem_st16(tmem_o + lane_addr + (uint32_t)(half * 32 + c0), o);
                    }
                    tmem_st_wait();
                }
            }
            // P = exp2(S - m) as bf16 hi/lo in the K-major 128B-swizzle layout of the PV A operand
#pragma unroll
            for (int cc = 0; cc < kHK / 8; ++cc) {               // 16-byte chunk (8 keys) of the 128-byte row
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
        float* xl = xm + (n_kt & 1) * 256;
        xl[half * 128 + r] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float inv = 1.0f / (l_run + xl[(half ^ 1) * 128 + r]);
        mbar_wait(o_full, (n_kt - 1) & 1);
        tc_fence_after();
        const size_t ob = ((size_t)b * a.S + (row_ok ? sq : 0)) * kHidden + (size_t)(kvh * 3 + hh) * kHeadDim + half * 32;
#pragma unroll
        for (int c0 = 0; c0 < kHeadDim / 2; c0 += 16) {
            float o[16];
            tmem_ld16(tmem_o + lane_addr + (uint32_t)(half * 32 + c0), o);     // warp-collective: every lane loads, valid rows store
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] *= inv;
                store_planes8(a.out_hi, a.out_lo, ob + c0, o);
                store_planes8(a.out_hi, a.out_lo, ob + c0 + 8, o + 8);
            }
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
    dim3 grid((unsigned)((S + kQPT - 1) / kQPT), kKvHeads, (unsigned)B);
    return launch_k(kern, grid, dim3(kAttnThreads), C::SMEM, st, tq[0], tq[1], tk[0], tk[1], tv[0], tv[1], a);
}

}  // namespace

cudaError_t launch_prefill_attention_umma(const PrefillAttnPlanes& p, int B, int S, bf16* out_hi, bf16* out_lo,
                                          cudaStream_t st) {
    if (!encode_fn() || (p.vt_ld & 7) || S < 1) return cudaErrorInvalidValue;
    return out_lo ? launch_attn<true>(p, B, S, out_hi, out_lo, st) : launch_attn<false>(p, B, S, out_hi, out_lo, st);
}

}  // namespace mb
