// SmolLM2 decoder-side kernels that are not GEMMs: prefix assembly (reference mellow/model/decoder.py:14-55),
// causal prefill attention and split-KV decode attention (transformers modeling_llama.py:276-285; the reference has
// no KV cache, wrapper.py:216-217), and the sampling step (wrapper.py:218-249).
#include <cstdlib>

#include "kernels.cuh"

namespace mb {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// prefix[b] = [lat1, 128 pooled a1, sep, lat2, 128 pooled a2, sep, 129 text embeddings]   (B,389,576)
// The 128 pooled slots are avg_pool(kernel 8) over 32 copies of each of the 32 frame rows, i.e. each unique row 4
// times; the 8-term fp32 running sum and the divide are kept so the value is bit-identical to the reference's pool.
__global__ void __launch_bounds__(144) prefix_kernel(const float* __restrict__ rows33, const int* __restrict__ ids,
                                                     const float* __restrict__ embed, int B, float* __restrict__ prefix) {
    const int p = blockIdx.x, b = blockIdx.y;
    const int c4 = threadIdx.x;                                   // 144 float4 = 576 floats
    pdl_trigger();
    pdl_wait();
    float4 v;
    if (p < 2 * (kAudioSlots + 1)) {
        const int which = p / (kAudioSlots + 1);                  // 0: audio1, 1: audio2
        const int slot = p - which * (kAudioSlots + 1);
        if (slot == kAudioSlots) {
            v = reinterpret_cast<const float4*>(embed)[c4];       // separator = embed_tokens[0], decoder.py:49-50
        } else {
            const int row = slot == 0 ? 0 : 1 + (slot - 1) / 4;
            const float4 r = reinterpret_cast<const float4*>(rows33 + (((size_t)which * B + b) * kAudioRows + row) * kProj)[c4];
            if (slot == 0) {
                v = r;
            } else {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 8; ++k) { s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w; }
                v = make_float4(s.x / 8.0f, s.y / 8.0f, s.z / 8.0f, s.w / 8.0f);
            }
        }
    } else {
        const int id = ids[b * kTextLen + (p - 2 * (kAudioSlots + 1))];
        v = reinterpret_cast<const float4*>(embed + (size_t)id * kHidden)[c4];
    }
    reinterpret_cast<float4*>(prefix + ((size_t)b * kPrefix + p) * kHidden)[c4] = v;
}

#ifdef MB_LAB   // round 1's 64-key tile kernel (attn_variant 0): lab builds only, kept as a cross-check
// ---------------------------------------------------------------------------------------------------------------
// Decode attention: one new query per row, ctx keys.  CTA = (split, kv head, row): the 3 query heads that share the
// kv head (GQA, modeling_llama.py:187-196) are processed together so K/V are read once.  128 threads, 64-key tiles,
// cp.async double buffering; partial (m, l, acc) per split are merged by decode_combine_kernel.
template <typename T>
struct DecodeSmem {
    static constexpr int ROWB = KvRowBytes<T>::value;             // bytes of one cached row (64 values)
    static constexpr int KROWB = ROWB + 16;                       // +16 B: conflict-free 16 B row reads
    __align__(16) unsigned char k[2][64][KROWB];
    __align__(16) unsigned char v[2][64][ROWB];
    __align__(16) float sc[3][64];
    float alpha[3], m[3], l[3];
    float red[4][3][kHeadDim];
};

// Three CTAs per SM (<= 170 registers).  Four CTAs of the 24-bit variant (128 registers, unpadded rows) were measured
// slower in round 1: 23.4 vs 18.7 us.
template <typename T>
__global__ void __launch_bounds__(128, 3) decode_attention_kernel(const DecodeAttnArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DecodeSmem<T>& sm = *reinterpret_cast<DecodeSmem<T>*>(smem_raw);
    constexpr bool F24 = sizeof(T) == 3;          // kKvF24 rows: 64 x u16 upper halves, then 64 x u8 mantissa bytes
    constexpr int ROWB = DecodeSmem<T>::ROWB;
    constexpr int NCHB = ROWB / 16;               // 16-byte chunks per cached row
    constexpr int E = F24 ? 8 : 16 / sizeof(T);   // values per 16 B chunk of the part the score role walks (the upper halves for f24)
    constexpr int NCH = kHeadDim / E;             // such chunks per row
    const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Tile ownership is static (independent of the step counter): split s owns tiles [s*tps, (s+1)*tps).
    const int t_begin = split * a.tps;
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(a.kc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;
    const unsigned char* vb = reinterpret_cast<const unsigned char*>(a.vc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;

    auto load_tile = [&](int buf, int tile, int ctx_limit) {
        const int key0 = tile << 6;
        for (int c = tid; c < 64 * NCHB; c += 128) {
            const int j = c / NCHB, ch = c - j * NCHB;
            const int key = key0 + j;
            const bool ok = key < ctx_limit;
            const size_t off = (size_t)(ok ? key : 0) * ROWB + ch * 16;
            cp_async16(&sm.k[buf][j][ch * 16], kb + off, ok);
            cp_async16(&sm.v[buf][j][ch * 16], vb + off, ok);
        }
    };
    // PDL: tiles that lie entirely inside the prefill prefix (keys < ctx_base - 1) are immutable history for every
    // decode step, so they are requested before waiting on the predecessor.  Nothing that the predecessor chain
    // writes (the step counter, the newest K/V row, q) is touched before pdl_wait().
    pdl_trigger();
    unsigned trec = kTraceNone;
    if (tid == 0) trec = trace_open(a.trace, a.trace_id);
    const bool early0 = (t_begin + 1) * 64 < a.ctx_base;
    const bool early1 = a.tps > 1 && (t_begin + 2) * 64 < a.ctx_base;
    if (early0) load_tile(0, t_begin, a.ctx_base);
    if (early1) load_tile(1, t_begin + 1, a.ctx_base);
    if (a.pf_keys != 0 && warp == 1) {
        // Everything this CTA will stream except the row the running QKV GEMM writes (key ctx-1) is immutable during
        // this step (the step counter was advanced by the previous step's last kernel, a full dependency back), so
        // while the CTA waits for its predecessor it asks the L2 to fetch its K/V history: the wait lasts 5-8 us
        // (the QKV GEMM body), during which HBM is otherwise idle.
        const int hist = a.ctx_base + (a.d_step ? *reinterpret_cast<const volatile int*>(a.d_step) : 0) - 1;
        int k_lo = (t_begin + (early0 ? 1 : 0) + (early1 ? 1 : 0)) * 64;
        int k_hi = min(hist, (t_begin + a.tps) * 64);
        if (a.pf_keys > 0) k_hi = min(k_hi, a.pf_keys);
        const int bytes = (k_hi - k_lo) * ROWB;
        for (int off = lane * 4096; off < bytes; off += 32 * 4096) {
            const int sz = min(4096, bytes - off);
            l2_prefetch(kb + (size_t)k_lo * ROWB + off, sz);
            l2_prefetch(vb + (size_t)k_lo * ROWB + off, sz);
        }
    }
    pdl_wait();
    if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, TR_WAITED);
    const int step_now = a.d_step ? *a.d_step : 0;
    // a row that has emitted the stop token is finished (the reference cuts its text there,
    // wrapper.py:254); its later tokens are never read, so its K/V stream -- the dominant decode traffic -- is skipped.
    // The uniform exit happens after the cp.async groups above are drained by the hardware at CTA exit.
    if (a.done && a.done[b]) { cp_async_commit(); cp_async_wait<0>(); return; }
    const int ctx = a.ctx_base + step_now;
    const int ntiles = (ctx + 63) >> 6;
    const int t_end = min(ntiles, t_begin + a.tps);
    const int ctx_ld = ctx;
    if (!early0 && t_begin < t_end) load_tile(0, t_begin, ctx_ld);
    cp_async_commit();
    if (!early1 && t_begin + 1 < t_end) load_tile(1, t_begin + 1, ctx_ld);
    cp_async_commit();
    // score role: thread = (key j, half); each half owns every other 16 B chunk of the key row.  The matching
    // slices of the three query heads live in registers.
    const int sj = tid >> 1, shalf = tid & 1;
    float qreg[3][kHeadDim / 2];
    {
        const float* qb = a.q + (size_t)b * kHidden + (kvh * 3) * kHeadDim;
#pragma unroll
        for (int h = 0; h < 3; ++h)
#pragma unroll
            for (int i = 0; i < NCH / 2; ++i)
#pragma unroll
                for (int e = 0; e < E; e += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(qb + h * kHeadDim + (2 * i + shalf) * E + e);
                    qreg[h][i * E + e] = t4.x; qreg[h][i * E + e + 1] = t4.y;
                    qreg[h][i * E + e + 2] = t4.z; qreg[h][i * E + e + 3] = t4.w;
                }
    }
    if (tid < 3) { sm.m[tid] = -INFINITY; sm.l[tid] = 0.f; }

    // PV role: thread = (dim pair, key quarter)
    const int dp = (tid & 31) * 2, kq = tid >> 5;
    float acc[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    for (int t = t_begin; t < t_end; ++t) {
        const int buf = (t - t_begin) & 1;
        if (t + 1 < t_end) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
            for (int i = 0; i < NCH / 2; ++i) {
                if constexpr (F24) {
                    // 8 values: one 16 B read of upper halves, one 8 B read of mantissa bytes; one PRMT per value
                    const unsigned char* rowp = sm.k[buf][sj];
                    const uint4 hv = *reinterpret_cast<const uint4*>(rowp + (2 * i + shalf) * 16);
                    const uint2 lv = *reinterpret_cast<const uint2*>(rowp + 128 + (2 * i + shalf) * 8);
                    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float k0 = f24_unpack_even(hw[p], p < 2 ? lv.x : lv.y, p & 1);
                        const float k1 = f24_unpack_odd(hw[p], p < 2 ? lv.x : lv.y, p & 1);
                        p0 += qreg[0][i * E + 2 * p] * k0; p1 += qreg[1][i * E + 2 * p] * k0; p2 += qreg[2][i * E + 2 * p] * k0;
                        p0 += qreg[0][i * E + 2 * p + 1] * k1; p1 += qreg[1][i * E + 2 * p + 1] * k1; p2 += qreg[2][i * E + 2 * p + 1] * k1;
                    }
                } else {
                    const T* kp = reinterpret_cast<const T*>(sm.k[buf][sj]) + (2 * i + shalf) * E;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const float kvv = kv_load(kp + e);
                        p0 += qreg[0][i * E + e] * kvv; p1 += qreg[1][i * E + e] * kvv; p2 += qreg[2][i * E + e] * kvv;
                    }
                }
            }
            p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
            p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
            p2 += __shfl_xor_sync(0xffffffffu, p2, 1);
            if (shalf == 0) {
                const bool ok = ((t << 6) + sj) < ctx;
                sm.sc[0][sj] = ok ? p0 * 0.125f : -INFINITY;
                sm.sc[1][sj] = ok ? p1 * 0.125f : -INFINITY;
                sm.sc[2][sj] = ok ? p2 * 0.125f : -INFINITY;
            }
        }
        __syncthreads();
        if (warp < 3) {                                              // online softmax bookkeeping, one warp per head
            const float s0 = sm.sc[warp][lane], s1 = sm.sc[warp][lane + 32];
            const float mt = warp_max(fmaxf(s0, s1));                // finite: every tile in range has >= 1 valid key
            const float m_old = sm.m[warp];
            const float m_new = fmaxf(m_old, mt);
            const float e0 = expf(s0 - m_new), e1 = expf(s1 - m_new);
            const float ls = warp_sum(e0 + e1);
            sm.sc[warp][lane] = e0;
            sm.sc[warp][lane + 32] = e1;
            if (lane == 0) {
                const float al = expf(m_old - m_new);
                sm.alpha[warp] = al;
                sm.l[warp] = sm.l[warp] * al + ls;
                sm.m[warp] = m_new;
            }
        }
        __syncthreads();
        {
            float x[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
                const int j = kq * 16 + jj;
                float4 p[3];
#pragma unroll
                for (int h = 0; h < 3; ++h) p[h] = *reinterpret_cast<const float4*>(&sm.sc[h][j]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float v0, v1;
                    if constexpr (F24) {
                        const unsigned char* rowp = sm.v[buf][j + u];
                        const uint32_t hw = *reinterpret_cast<const uint32_t*>(rowp + dp * 2);
                        const uint32_t lb = *reinterpret_cast<const unsigned short*>(rowp + 128 + dp);
                        v0 = f24_unpack_even(hw, lb, 0);
                        v1 = f24_unpack_odd(hw, lb, 0);
                    } else {
                        const T* vp = reinterpret_cast<const T*>(sm.v[buf][j + u]);
                        v0 = kv_load(vp + dp); v1 = kv_load(vp + dp + 1);
                    }
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        const float pv = u == 0 ? p[h].x : (u == 1 ? p[h].y : (u == 2 ? p[h].z : p[h].w));
                        x[h][0] += pv * v0; x[h][1] += pv * v1;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const float al = sm.alpha[h];
                acc[h][0] = acc[h][0] * al + x[h][0];
                acc[h][1] = acc[h][1] * al + x[h][1];
            }
        }
        __syncthreads();                                             // tile buffer and sc are reused
        if (t + 2 < t_end) load_tile(buf, t + 2, ctx_ld);
        cp_async_commit();
    }
    cp_async_wait<0>();
    if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, 5);           // all key tiles consumed
#pragma unroll
    for (int h = 0; h < 3; ++h) { sm.red[kq][h][dp] = acc[h][0]; sm.red[kq][h][dp + 1] = acc[h][1]; }
    __syncthreads();
    for (int e = tid; e < 3 * kHeadDim; e += 128) {
        const int h = e >> 6, d = e & 63;
        const float acc_hd = sm.red[0][h][d] + sm.red[1][h][d] + sm.red[2][h][d] + sm.red[3][h][d];
        if (a.nsplit == 1) {
            // this CTA saw every key of its (row, kv head): finish here, no partials, no combine kernel
            store_planes1(a.out_hi, a.out_lo, (size_t)b * kHidden + (kvh * 3 + h) * kHeadDim + d, acc_hd / sm.l[h]);
        } else {
            const size_t o = (((size_t)b * kHeads + kvh * 3 + h) * a.nsplit + split);
            a.part_acc[o * kHeadDim + d] = acc_hd;
            if (d == 0) { a.part_ml[o * 2] = sm.m[h]; a.part_ml[o * 2 + 1] = sm.l[h]; }
        }
    }
    if (tid == 0) trace_close(a.trace, trec, a.trace_id);
}


#endif  // MB_LAB

// ---------------------------------------------------------------------------------------------------------------
// Decode attention, warp-autonomous version (the product kernel).  What ncu showed on round 1's tile kernel (lab
// builds) at B = 128 with 24-bit rows (profiles/r2_decode_attention_kv24_*): 38 % of the DRAM peak, 0.42 issued
// instructions per scheduler cycle, 12 warps per SM, and stalls spread over the block barriers (four per 64-key tile),
// the cp.async scoreboard and shared-memory latency -- a latency problem, not a bandwidth one.  Here the four warps of
// the CTA never synchronise inside the loop: warp w streams the 16-key chunks w, w+4, w+8, ... of the CTA's key range
// through a cp.async ring of its OWN, scores them (lane = key x half of the head dim, the three GQA query heads in
// registers), runs the online softmax with shuffles, and accumulates P.V with lane = pair of head dims.  The four
// (m, l, acc) states are merged once at the end.  About half the instructions per key of the tile kernel and no
// block-wide barrier until the merge.
constexpr int kAttnChunk = 16;                                    // keys per warp-private chunk
// BULK = false: 16-byte cp.async pieces into padded key rows.  BULK = true (variant 2): a chunk is ONE bulk copy per
// operand (cp.async.bulk: 16 keys x ROWB contiguous bytes, completion on a per-(warp, slot) mbarrier), which removes the
// per-piece address arithmetic (a fifth of the kernel's instructions); the rows are then unpadded, and the score role
// stays free of bank conflicts by walking the 16-byte pieces of its key row in an order rotated by the key index (the
// query registers are loaded in the matching order once).
template <typename T, int ST, bool BULK>
struct DecodeSmemW {
    static constexpr int ROWB = KvRowBytes<T>::value;
    static constexpr int KROWB = BULK ? ROWB : ROWB + 16;         // +16 B: conflict-free 16 B reads with lane = key
    __align__(128) unsigned char k[4][ST][kAttnChunk][KROWB];
    __align__(128) unsigned char v[4][ST][kAttnChunk][ROWB];
    __align__(16) float p[4][3][kAttnChunk];
    __align__(16) float red[4][3][kHeadDim];
    float ml[4][3][2];
    __align__(8) unsigned long long full[4][ST];                  // BULK: "chunk landed" barriers
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void bar_wait_parity(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();    // a broken pipeline must fail, not hang the GPU
    }
}

template <typename T, int ST, bool BULK>
__global__ void __launch_bounds__(128, 3) decode_attention_warp_kernel(const DecodeAttnArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = DecodeSmemW<T, ST, BULK>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    constexpr bool F24 = sizeof(T) == 3;
    constexpr int ROWB = SM::ROWB;
    constexpr int NCHB = ROWB / 16;               // 16-byte pieces per cached row
    constexpr int E = F24 ? 8 : 16 / sizeof(T);   // values per 16 B piece of the part the score role walks
    constexpr int NCH = kHeadDim / E;
    constexpr int NST = NCH / 2;                  // pieces each half of a lane pair walks
    const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(a.kc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;
    const unsigned char* vb = reinterpret_cast<const unsigned char*>(a.vc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;
    const int k_begin = split * a.tps * 64;       // static ownership of the key range, like the tile kernel

    // chunk i of this warp = keys [k_begin + 16 * (warp + 4 i), +16), staged in ring slot i % ST
    auto load_chunk = [&](int i, int ctx_limit) {
        const int key0 = k_begin + kAttnChunk * (warp + 4 * i);
        if constexpr (BULK) {
            if (lane == 0) {
                // rows at and beyond ctx_limit are not copied: their slots keep zeros / older finite rows and are masked
                const uint32_t bytes = (uint32_t)(min(kAttnChunk, ctx_limit - key0) * ROWB);
                unsigned long long* bar = &sm.full[warp][i % ST];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the slot -> async writes
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                             ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(2u * bytes) : "memory");
                bulk_load(&sm.k[warp][i % ST][0][0], kb + (size_t)key0 * ROWB, bytes, bar);
                bulk_load(&sm.v[warp][i % ST][0][0], vb + (size_t)key0 * ROWB, bytes, bar);
            }
        } else {
            unsigned char (*kd)[SM::KROWB] = sm.k[warp][i % ST];
            unsigned char (*vd)[ROWB] = sm.v[warp][i % ST];
            for (int c = lane; c < kAttnChunk * NCHB; c += 32) {
                const int j = c / NCHB, ch = c - j * NCHB;
                const bool ok = key0 + j < ctx_limit;
                const size_t off = (size_t)(ok ? key0 + j : 0) * ROWB + ch * 16;
                cp_async16(&kd[j][ch * 16], kb + off, ok);
                cp_async16(&vd[j][ch * 16], vb + off, ok);
            }
        }
    };
    pdl_trigger();
    unsigned trec = kTraceNone;
    if (tid == 0) trec = trace_open(a.trace, a.trace_id);
    if constexpr (BULK) {
        if (lane < ST) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm.full[warp][lane])) : "memory");
        }
        // value slots start as zeros: a partially filled last chunk leaves rows the copy did not touch, and 0 * NaN would
        // poison the accumulators (their probabilities are exactly zero)
        for (int c = lane; c < ST * kAttnChunk * ROWB / 16; c += 32)
            reinterpret_cast<uint4*>(&sm.v[warp][0][0][0])[c] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    // PDL: chunks that lie inside the prefill prefix are immutable history and are requested before the wait
    int n_early = 0;
#pragma unroll
    for (int i = 0; i < ST; ++i) {
        if (n_early == i && k_begin + kAttnChunk * (warp + 4 * i + 1) <= a.ctx_base && warp + 4 * i < a.tps * 4) {
            load_chunk(i, a.ctx_base);
            if (!BULK) cp_async_commit();
            n_early = i + 1;
        }
    }
    pdl_wait();
    if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, TR_WAITED);
    const int step_now = a.d_step ? *a.d_step : 0;
    if (a.done && a.done[b]) {                                    // finished row: no K/V stream
        if constexpr (BULK) { for (int i = 0; i < n_early; ++i) bar_wait_parity(&sm.full[warp][i], 0); }
        else cp_async_wait<0>();
        return;
    }
    const int ctx = a.ctx_base + step_now;
    const int k_end = min(ctx, k_begin + a.tps * 64);
    const int n_chunks = k_end > k_begin ? (k_end - k_begin + kAttnChunk - 1) / kAttnChunk : 0;
    const int n_mine = n_chunks > warp ? (n_chunks - warp + 3) / 4 : 0;
#pragma unroll
    for (int i = 0; i < ST; ++i) {                                // fill the ring: one commit group per slot, empty or not
        if (i >= n_early) {
            if (i < n_mine) load_chunk(i, ctx);
            if (!BULK) cp_async_commit();
        }
    }
    // score role: lane = (key j, half); each half owns every other 16 B piece of the key row.  BULK: the pieces are
    // walked in an order rotated by the key index so that unpadded rows are read without bank conflicts.
    const int sj = lane >> 1, shalf = lane & 1;
    const int rot = !BULK ? 0 : (F24 ? (sj >> 1) & (NST - 1) : sj & (NST - 1));
    float qreg[3][kHeadDim / 2];
    {
        const float* qb = a.q + (size_t)b * kHidden + (kvh * 3) * kHeadDim;
#pragma unroll
        for (int h = 0; h < 3; ++h)
#pragma unroll
            for (int i = 0; i < NST; ++i) {
                const int piece = 2 * ((i + rot) & (NST - 1)) + shalf;
#pragma unroll
                for (int e = 0; e < E; e += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(qb + h * kHeadDim + piece * E + e);
                    qreg[h][i * E + e] = t4.x * 0.125f; qreg[h][i * E + e + 1] = t4.y * 0.125f;      // head_dim^-0.5, exact
                    qreg[h][i * E + e + 2] = t4.z * 0.125f; qreg[h][i * E + e + 3] = t4.w * 0.125f;
                }
            }
    }
    float m_run[3] = {-INFINITY, -INFINITY, -INFINITY}, l_run[3] = {0.f, 0.f, 0.f};
    float acc[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    const int dp = lane * 2;                                      // PV role: this lane's pair of head dims
    for (int i = 0; i < n_mine; ++i) {
        const int slot = i % ST;
        if constexpr (BULK) {
            bar_wait_parity(&sm.full[warp][slot], (uint32_t)(i / ST) & 1u);
        } else {
            cp_async_wait<ST - 1>();                              // chunk i has landed (this thread's pieces) ...
            __syncwarp();                                         // ... and everybody else's
        }
        const int key0 = k_begin + kAttnChunk * (warp + 4 * i);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        {
            const unsigned char* rowp = sm.k[warp][slot][sj];
#pragma unroll
            for (int c = 0; c < NST; ++c) {
                const int piece = 2 * ((c + rot) & (NST - 1)) + shalf;
                if constexpr (F24) {
                    const uint4 hv = *reinterpret_cast<const uint4*>(rowp + piece * 16);
                    const uint2 lv = *reinterpret_cast<const uint2*>(rowp + 128 + piece * 8);
                    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float k0 = f24_unpack_even(hw[p], p < 2 ? lv.x : lv.y, p & 1);
                        const float k1 = f24_unpack_odd(hw[p], p < 2 ? lv.x : lv.y, p & 1);
                        s0 += qreg[0][c * E + 2 * p] * k0; s1 += qreg[1][c * E + 2 * p] * k0; s2 += qreg[2][c * E + 2 * p] * k0;
                        s0 += qreg[0][c * E + 2 * p + 1] * k1; s1 += qreg[1][c * E + 2 * p + 1] * k1; s2 += qreg[2][c * E + 2 * p + 1] * k1;
                    }
                } else {
                    const T* kp = reinterpret_cast<const T*>(rowp) + piece * E;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const float kvv = kv_load(kp + e);
                        s0 += qreg[0][c * E + e] * kvv; s1 += qreg[1][c * E + e] * kvv; s2 += qreg[2][c * E + e] * kvv;
                    }
                }
            }
        }
        float sc[3] = {s0, s1, s2};
        const bool ok = key0 + sj < ctx;
        float al[3];
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            float s = sc[h] + __shfl_xor_sync(0xffffffffu, sc[h], 1);         // the two halves of the head dim
            s = ok ? s : -INFINITY;
            float mt = s;
#pragma unroll
            for (int o = 2; o < 32; o <<= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
            const float m_new = fmaxf(m_run[h], mt);                          // finite: the chunk holds >= 1 valid key
            const float e = expf(s - m_new);
            float ls = e;
#pragma unroll
            for (int o = 2; o < 32; o <<= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
            al[h] = expf(m_run[h] - m_new);
            l_run[h] = l_run[h] * al[h] + ls;
            m_run[h] = m_new;
            if (shalf == 0) sm.p[warp][h][sj] = e;
        }
        __syncwarp();
        {
            float x[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
            for (int jj = 0; jj < kAttnChunk; jj += 4) {
                float4 p[3];
#pragma unroll
                for (int h = 0; h < 3; ++h) p[h] = *reinterpret_cast<const float4*>(&sm.p[warp][h][jj]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float v0, v1;
                    const unsigned char* rowp = sm.v[warp][slot][jj + u];
                    if constexpr (F24) {
                        const uint32_t hw = *reinterpret_cast<const uint32_t*>(rowp + dp * 2);
                        const uint32_t lb = *reinterpret_cast<const unsigned short*>(rowp + 128 + dp);
                        v0 = f24_unpack_even(hw, lb, 0);
                        v1 = f24_unpack_odd(hw, lb, 0);
                    } else {
                        const T* vp = reinterpret_cast<const T*>(rowp);
                        v0 = kv_load(vp + dp); v1 = kv_load(vp + dp + 1);
                    }
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        const float pv = u == 0 ? p[h].x : (u == 1 ? p[h].y : (u == 2 ? p[h].z : p[h].w));
                        x[h][0] += pv * v0; x[h][1] += pv * v1;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                acc[h][0] = acc[h][0] * al[h] + x[h][0];
                acc[h][1] = acc[h][1] * al[h] + x[h][1];
            }
        }
        __syncwarp();                                             // the slot and sm.p are reused
        if (i + ST < n_mine) load_chunk(i + ST, ctx);
        if (!BULK) cp_async_commit();
    }
    if (!BULK) cp_async_wait<0>();
    if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, 5);           // this warp's keys consumed
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        sm.red[warp][h][dp] = acc[h][0]; sm.red[warp][h][dp + 1] = acc[h][1];
        if (lane == 0) { sm.ml[warp][h][0] = m_run[h]; sm.ml[warp][h][1] = l_run[h]; }
    }
    __syncthreads();
    for (int e = tid; e < 3 * kHeadDim; e += 128) {               // merge the four warp states (fixed order)
        const int h = e >> 6, d = e & 63;
        float m = sm.ml[0][h][0];
#pragma unroll
        for (int w = 1; w < 4; ++w) m = fmaxf(m, sm.ml[w][h][0]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float mw = sm.ml[w][h][0];
            const float sc_w = mw == -INFINITY ? 0.f : expf(mw - m);          // a warp without keys contributes nothing
            num += sc_w * sm.red[w][h][d];
            den += sc_w * sm.ml[w][h][1];
        }
        if (a.nsplit == 1) {
            store_planes1(a.out_hi, a.out_lo, (size_t)b * kHidden + (kvh * 3 + h) * kHeadDim + d, num / den);
        } else {
            const size_t o = (((size_t)b * kHeads + kvh * 3 + h) * a.nsplit + split);
            a.part_acc[o * kHeadDim + d] = num;
            if (d == 0) { a.part_ml[o * 2] = m; a.part_ml[o * 2 + 1] = den; }
        }
    }
    if (tid == 0) trace_close(a.trace, trec, a.trace_id);
}

// The same kernel following a WORK LIST (DecodeAttnArgs::assign): launched instead of the plain kernel
// above once two thirds of the rows have finished, so that the CTAs of finished rows take key shares of the rows
// still decoding and the last CTA of a (row, kv head) merges the shares.  A separate function on purpose: carrying the
// work-list code in the plain kernel, even as a dead template branch, cost it 0.8 - 3 % (builds A/B'd against each
// other inside one gpurun call, profiles/r2_lib_ab_*.jsonl).  Only instantiated with BULK = DYN = true.
template <typename T, int ST, bool BULK, bool DYN>
__global__ void __launch_bounds__(128, 3) decode_attention_share_kernel(const DecodeAttnArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = DecodeSmemW<T, ST, BULK>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    constexpr bool F24 = sizeof(T) == 3;
    constexpr int ROWB = SM::ROWB;
    constexpr int NCHB = ROWB / 16;               // 16-byte pieces per cached row
    constexpr int E = F24 ? 8 : 16 / sizeof(T);   // values per 16 B piece of the part the score role walks
    constexpr int NCH = kHeadDim / E;
    constexpr int NST = NCH / 2;                  // pieces each half of a lane pair walks
    const int split = blockIdx.x, kvh = blockIdx.y, slot_b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // What this CTA works on: row b, share `part` of `nparts` of its keys (nparts = 0: nothing).  Plain kernel: its own
    // row, the static split.  DYN: the entry of this slot in the work list (DecodeAttnArgs::assign).
    int b = slot_b, part = 0, nparts = 1, tps = a.tps;
    int k_begin = split * a.tps * 64;             // static ownership of the key range, like the tile kernel
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(a.kc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;
    const unsigned char* vb = reinterpret_cast<const unsigned char*>(a.vc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;
    auto take = [&](int v, int ctx_now) {
        b = v & 0xff; part = (v >> 8) & 0xff; nparts = (v >> 16) & 0xff;
        if (b >= a.B || nparts > kAttnDynParts || part >= nparts) nparts = 0;   // only a stale speculative read can look like this
        const bool shared = nparts > 1;
        tps = shared ? ((ctx_now + 63) / 64 + nparts - 1) / nparts : a.tps;
        k_begin = (shared ? part : split) * tps * 64;
        kb = reinterpret_cast<const unsigned char*>(a.kc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;
        vb = reinterpret_cast<const unsigned char*>(a.vc) + ((size_t)b * kKvHeads + kvh) * a.t_max * ROWB;
    };

    // chunk i of this warp = keys [k_begin + 16 * (warp + 4 i), +16), staged in ring slot i % ST.  DYN: ring position
    // ring0 + i, ring0 = number of early requests that were dropped (the plain kernel keeps the static arithmetic)
    [[maybe_unused]] int ring0 = 0;
    auto ring = [&](int i) { if constexpr (DYN) return ring0 + i; else return i; };
    auto load_chunk = [&](int i, int ctx_limit) {
        const int key0 = k_begin + kAttnChunk * (warp + 4 * i);
        const int slot = ring(i) % ST;
        if constexpr (BULK) {
            if (lane == 0) {
                // rows at and beyond ctx_limit are not copied: their slots keep zeros / older finite rows and are masked
                const uint32_t bytes = (uint32_t)(min(kAttnChunk, ctx_limit - key0) * ROWB);
                unsigned long long* bar = &sm.full[warp][slot];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the slot -> async writes
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                             ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(2u * bytes) : "memory");
                bulk_load(&sm.k[warp][slot][0][0], kb + (size_t)key0 * ROWB, bytes, bar);
                bulk_load(&sm.v[warp][slot][0][0], vb + (size_t)key0 * ROWB, bytes, bar);
            }
        } else {
            unsigned char (*kd)[SM::KROWB] = sm.k[warp][slot];
            unsigned char (*vd)[ROWB] = sm.v[warp][slot];
            for (int c = lane; c < kAttnChunk * NCHB; c += 32) {
                const int j = c / NCHB, ch = c - j * NCHB;
                const bool ok = key0 + j < ctx_limit;
                const size_t off = (size_t)(ok ? key0 + j : 0) * ROWB + ch * 16;
                cp_async16(&kd[j][ch * 16], kb + off, ok);
                cp_async16(&vd[j][ch * 16], vb + off, ok);
            }
        }
    };
    pdl_trigger();
    unsigned trec = kTraceNone;
    if (tid == 0) trec = trace_open(a.trace, a.trace_id);
    if constexpr (BULK) {
        if (lane < ST) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm.full[warp][lane])) : "memory");
        }
        // value slots start as zeros: a partially filled last chunk leaves rows the copy did not touch, and 0 * NaN would
        // poison the accumulators (their probabilities are exactly zero)
        for (int c = lane; c < ST * kAttnChunk * ROWB / 16; c += 32)
            reinterpret_cast<uint4*>(&sm.v[warp][0][0][0])[c] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    // PDL: chunks that lie inside the prefill prefix are immutable history and are requested before the wait -- as if
    // this slot worked on its own row from its first key, which is what every row that is still decoding does (a row's
    // own slot keeps share 0 of a work list): nothing here waits for a load.
    int n_early = 0;
    auto request_early = [&]() {
        n_early = 0;
#pragma unroll
        for (int i = 0; i < ST; ++i) {
            if (n_early == i && k_begin + kAttnChunk * (warp + 4 * i + 1) <= a.ctx_base && warp + 4 * i < tps * 4) {
                load_chunk(i, a.ctx_base);
                if (!BULK) cp_async_commit();
                n_early = i + 1;
            }
        }
    };
    auto drop_early = [&]() {                                     // the requested chunks are not this CTA's keys: let them land
        if constexpr (BULK) {
            for (int i = 0; i < n_early; ++i)
                bar_wait_parity(&sm.full[warp][(ring0 + i) % ST], (uint32_t)((ring0 + i) / ST) & 1u);
        } else cp_async_wait<0>();
        __syncwarp();
        ring0 += n_early; n_early = 0;                            // the ring goes on behind them
    };
    request_early();
    [[maybe_unused]] int b_early = b, k_early = k_begin;
    if constexpr (DYN) {
        // The slot of a finished row works on a key share of another row.  Its work list entry and the step counter are
        // read SPECULATIVELY here (their producers are several kernels back, but the chain of early launches gives no
        // guarantee) and checked after the wait; a stale guess only costs the early requests.
        const int v_spec = (int)__ldcg(a.assign + slot_b);
        if (((v_spec >> 16) & 0xff) > 1 && ((v_spec & 0xff) != slot_b || ((v_spec >> 8) & 0xff) != 0)) {
            const int ctx_spec = a.ctx_base + (a.d_step ? (int)__ldcg(a.d_step) : 0);
            drop_early();
            take(v_spec, ctx_spec);
            if (nparts > 1) { request_early(); b_early = b; k_early = k_begin; }
        }
    }
    pdl_wait();
    if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, TR_WAITED);
    const int step_now = a.d_step ? *a.d_step : 0;
    const int ctx = a.ctx_base + step_now;
    if constexpr (DYN) {
        // the slots of finished rows are handed a share of the keys of the rows still decoding
        take(a.assign[slot_b], ctx);
        if (n_early > 0 && (nparts == 0 || b != b_early || k_begin != k_early)) drop_early();
        if (nparts == 0) return;                                  // idle slot
    } else if (a.done && a.done[b]) {                             // finished row: no K/V stream
        if constexpr (BULK) { for (int i = 0; i < n_early; ++i) bar_wait_parity(&sm.full[warp][i], 0); }
        else cp_async_wait<0>();
        return;
    }
    const int k_end = min(ctx, k_begin + tps * 64);
    const int n_chunks = k_end > k_begin ? (k_end - k_begin + kAttnChunk - 1) / kAttnChunk : 0;
    const int n_mine = n_chunks > warp ? (n_chunks - warp + 3) / 4 : 0;
#pragma unroll
    for (int i = 0; i < ST; ++i) {                                // fill the ring: one commit group per slot, empty or not
        if (i >= n_early) {
            if (i < n_mine) load_chunk(i, ctx);
            if (!BULK) cp_async_commit();
        }
    }
    // score role: lane = (key j, half); each half owns every other 16 B piece of the key row.  BULK: the pieces are
    // walked in an order rotated by the key index so that unpadded rows are read without bank conflicts.
    const int sj = lane >> 1, shalf = lane & 1;
    const int rot = !BULK ? 0 : (F24 ? (sj >> 1) & (NST - 1) : sj & (NST - 1));
    float qreg[3][kHeadDim / 2];
    {
        const float* qb = a.q + (size_t)b * kHidden + (kvh * 3) * kHeadDim;
#pragma unroll
        for (int h = 0; h < 3; ++h)
#pragma unroll
            for (int i = 0; i < NST; ++i) {
                const int piece = 2 * ((i + rot) & (NST - 1)) + shalf;
#pragma unroll
                for (int e = 0; e < E; e += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(qb + h * kHeadDim + piece * E + e);
                    qreg[h][i * E + e] = t4.x * 0.125f; qreg[h][i * E + e + 1] = t4.y * 0.125f;      // head_dim^-0.5, exact
                    qreg[h][i * E + e + 2] = t4.z * 0.125f; qreg[h][i * E + e + 3] = t4.w * 0.125f;
                }
            }
    }
    float m_run[3] = {-INFINITY, -INFINITY, -INFINITY}, l_run[3] = {0.f, 0.f, 0.f};
    float acc[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    const int dp = lane * 2;                                      // PV role: this lane's pair of head dims
    for (int i = 0; i < n_mine; ++i) {
        const int slot = ring(i) % ST;
        if constexpr (BULK) {
            bar_wait_parity(&sm.full[warp][slot], (uint32_t)(ring(i) / ST) & 1u);
        } else {
            cp_async_wait<ST - 1>();                              // chunk i has landed (this thread's pieces) ...
            __syncwarp();                                         // ... and everybody else's
        }
        const int key0 = k_begin + kAttnChunk * (warp + 4 * i);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        {
            const unsigned char* rowp = sm.k[warp][slot][sj];
#pragma unroll
            for (int c = 0; c < NST; ++c) {
                const int piece = 2 * ((c + rot) & (NST - 1)) + shalf;
                if constexpr (F24) {
                    const uint4 hv = *reinterpret_cast<const uint4*>(rowp + piece * 16);
                    const uint2 lv = *reinterpret_cast<const uint2*>(rowp + 128 + piece * 8);
                    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float k0 = f24_unpack_even(hw[p], p < 2 ? lv.x : lv.y, p & 1);
                        const float k1 = f24_unpack_odd(hw[p], p < 2 ? lv.x : lv.y, p & 1);
                        s0 += qreg[0][c * E + 2 * p] * k0; s1 += qreg[1][c * E + 2 * p] * k0; s2 += qreg[2][c * E + 2 * p] * k0;
                        s0 += qreg[0][c * E + 2 * p + 1] * k1; s1 += qreg[1][c * E + 2 * p + 1] * k1; s2 += qreg[2][c * E + 2 * p + 1] * k1;
                    }
                } else {
                    const T* kp = reinterpret_cast<const T*>(rowp) + piece * E;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const float kvv = kv_load(kp + e);
                        s0 += qreg[0][c * E + e] * kvv; s1 += qreg[1][c * E + e] * kvv; s2 += qreg[2][c * E + e] * kvv;
                    }
                }
            }
        }
        float sc[3] = {s0, s1, s2};
        const bool ok = key0 + sj < ctx;
        float al[3];
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            float s = sc[h] + __shfl_xor_sync(0xffffffffu, sc[h], 1);         // the two halves of the head dim
            s = ok ? s : -INFINITY;
            float mt = s;
#pragma unroll
            for (int o = 2; o < 32; o <<= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
            const float m_new = fmaxf(m_run[h], mt);                          // finite: the chunk holds >= 1 valid key
            const float e = expf(s - m_new);
            float ls = e;
#pragma unroll
            for (int o = 2; o < 32; o <<= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
            al[h] = expf(m_run[h] - m_new);
            l_run[h] = l_run[h] * al[h] + ls;
            m_run[h] = m_new;
            if (shalf == 0) sm.p[warp][h][sj] = e;
        }
        __syncwarp();
        {
            float x[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
            for (int jj = 0; jj < kAttnChunk; jj += 4) {
                float4 p[3];
#pragma unroll
                for (int h = 0; h < 3; ++h) p[h] = *reinterpret_cast<const float4*>(&sm.p[warp][h][jj]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float v0, v1;
                    const unsigned char* rowp = sm.v[warp][slot][jj + u];
                    if constexpr (F24) {
                        const uint32_t hw = *reinterpret_cast<const uint32_t*>(rowp + dp * 2);
                        const uint32_t lb = *reinterpret_cast<const unsigned short*>(rowp + 128 + dp);
                        v0 = f24_unpack_even(hw, lb, 0);
                        v1 = f24_unpack_odd(hw, lb, 0);
                    } else {
                        const T* vp = reinterpret_cast<const T*>(rowp);
                        v0 = kv_load(vp + dp); v1 = kv_load(vp + dp + 1);
                    }
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        const float pv = u == 0 ? p[h].x : (u == 1 ? p[h].y : (u == 2 ? p[h].z : p[h].w));
                        x[h][0] += pv * v0; x[h][1] += pv * v1;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                acc[h][0] = acc[h][0] * al[h] + x[h][0];
                acc[h][1] = acc[h][1] * al[h] + x[h][1];
            }
        }
        __syncwarp();                                             // the slot and sm.p are reused
        if (i + ST < n_mine) load_chunk(i + ST, ctx);
        if (!BULK) cp_async_commit();
    }
    if (!BULK) cp_async_wait<0>();
    if constexpr (BULK && DYN) {                                  // early requests beyond a shortened share: let them land
        for (int i = n_mine; i < n_early; ++i) bar_wait_parity(&sm.full[warp][ring(i) % ST], (uint32_t)(ring(i) / ST) & 1u);
    }
    if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, 5);           // this warp's keys consumed
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        sm.red[warp][h][dp] = acc[h][0]; sm.red[warp][h][dp + 1] = acc[h][1];
        if (lane == 0) { sm.ml[warp][h][0] = m_run[h]; sm.ml[warp][h][1] = l_run[h]; }
    }
    __syncthreads();
    for (int e = tid; e < 3 * kHeadDim; e += 128) {               // merge the four warp states (fixed order)
        const int h = e >> 6, d = e & 63;
        float m = sm.ml[0][h][0];
#pragma unroll
        for (int w = 1; w < 4; ++w) m = fmaxf(m, sm.ml[w][h][0]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float mw = sm.ml[w][h][0];
            const float sc_w = mw == -INFINITY ? 0.f : expf(mw - m);          // a warp without keys contributes nothing
            num += sc_w * sm.red[w][h][d];
            den += sc_w * sm.ml[w][h][1];
        }
        if (DYN && nparts > 1) {                                  // key share of a work list: partial state of this share
            const size_t o = ((size_t)b * kHeads + kvh * 3 + h) * kAttnDynParts + part;
            a.part_acc[o * kHeadDim + d] = num;
            if (d == 0) { a.part_ml[o * 2] = m; a.part_ml[o * 2 + 1] = den; }
        } else if (a.nsplit == 1) {
            store_planes1(a.out_hi, a.out_lo, (size_t)b * kHidden + (kvh * 3 + h) * kHeadDim + d, num / den);
        } else {                                                  // static split: decode_combine_kernel merges
            const size_t o = (((size_t)b * kHeads + kvh * 3 + h) * a.nsplit + split);
            a.part_acc[o * kHeadDim + d] = num;
            if (d == 0) { a.part_ml[o * 2] = m; a.part_ml[o * 2 + 1] = den; }
        }
    }
    if constexpr (DYN) {
        if (nparts > 1) {
            // whichever of the nparts CTAs of this (row, kv head) finishes last merges their states, always in share order.
            // (The same in-kernel merge for the STATIC splits of small batches instead of the combine kernel was measured
            // 3 % slower per step -- fence + atomic + re-read cost more than a dependent launch -- and is not built:
            // profiles/r2_decode_ab_self_merge.jsonl.)
            __shared__ int s_last;
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                int* cnt = a.merge_count + b * kKvHeads + kvh;
                const int old = atomicAdd(cnt, 1);
                s_last = old == nparts - 1;
                if (s_last) *cnt = 0;                              // ready for the next layer's launch
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                for (int e = tid; e < 3 * kHeadDim; e += 128) {
                    const int h = e >> 6, d = e & 63;
                    const size_t base = ((size_t)b * kHeads + kvh * 3 + h) * kAttnDynParts;
                    float m = -INFINITY;
                    for (int s = 0; s < nparts; ++s) m = fmaxf(m, __ldcg(a.part_ml + (base + s) * 2));
                    float num = 0.f, den = 0.f;
                    for (int s = 0; s < nparts; ++s) {
                        const float ms = __ldcg(a.part_ml + (base + s) * 2);
                        if (ms == -INFINITY) continue;             // a share without keys
                        const float w = expf(ms - m);
                        num += w * __ldcg(a.part_acc + (base + s) * kHeadDim + d);
                        den += w * __ldcg(a.part_ml + (base + s) * 2 + 1);
                    }
                    store_planes1(a.out_hi, a.out_lo, (size_t)b * kHidden + (kvh * 3 + h) * kHeadDim + d, num / den);
                }
            }
        }
    }
    if (tid == 0) trace_close(a.trace, trec, a.trace_id);
}

__global__ void __launch_bounds__(64) decode_combine_kernel(const DecodeAttnArgs a) {
    const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
    const size_t base = ((size_t)b * kHeads + h) * a.nsplit;
    pdl_trigger();
    pdl_wait();
    if (a.done && a.done[b]) return;                               // finished row: partials were not produced
    float m = -INFINITY;
    for (int s = 0; s < a.nsplit; ++s) m = fmaxf(m, a.part_ml[(base + s) * 2]);
    float num = 0.f, den = 0.f;
    for (int s = 0; s < a.nsplit; ++s) {
        const float ms = a.part_ml[(base + s) * 2];
        if (ms == -INFINITY) continue;                             // empty split
        const float w = expf(ms - m);
        num += w * a.part_acc[(base + s) * kHeadDim + d];
        den += w * a.part_ml[(base + s) * 2 + 1];
    }
    store_planes1(a.out_hi, a.out_lo, (size_t)b * kHidden + h * kHeadDim + d, num / den);
}

// ---------------------------------------------------------------------------------------------------------------
// Sampling step of the reference (wrapper.py:218-249): logits / temperature, top-p filter, ARGMAX, next embedding,
// stop bookkeeping.  The reference's filter never removes the top-1 entry (the shifted mask clears index 0,
// wrapper.py:224-226) and the decision is an argmax over what is kept, so the kept-set argmax equals the global
// argmax for every top_p and every temperature > 0; the kernel therefore scans once for (max, first index).
// One CTA per row.
__global__ void __launch_bounds__(1024) sample_kernel(const SampleArgs a) {
    __shared__ float smax[32];
    __shared__ int sidx[32];
    __shared__ int stoken;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    pdl_wait();
    const int step = *a.d_step;
    const float* lg = a.logits ? a.logits + (size_t)b * kVocab : nullptr;
    const float tdiv = a.temperature > 0.f ? a.temperature : 1.0f;
    float* dump = a.logits_dump ? a.logits_dump + ((size_t)step * a.B + b) * kVocab : nullptr;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (a.logits) {
        for (int i = tid; i < kVocab; i += 1024) {
            const float v = lg[i] / tdiv;
            if (dump) dump[i] = v;
            if (v > best) { best = v; bi = i; }                    // strided scan keeps the smallest index per thread
        }
    } else {
        // lm_head already reduced every 16-column group to its (max, first argmax); dividing by a positive
        // temperature cannot change the order, so the candidates are compared unscaled
        for (int i = tid; i < a.n_cand; i += 1024) {
            const float v = a.cand_val[(size_t)b * a.n_cand + i];
            const int ix = a.cand_idx[(size_t)b * a.n_cand + i];
            if (v > best || (v == best && ix < bi)) { best = v; bi = ix; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { smax[warp] = best; sidx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        best = smax[lane]; bi = sidx[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            if (bi < 0 || bi >= kVocab) bi = 0;                    // all-NaN row: never index the embedding table out of range
            a.tokens_out[(size_t)b * a.max_len + step] = bi;
            const int fed = a.forced ? a.forced[(size_t)b * a.max_len + step] : bi;   // what the row continues with
            if (fed == a.eos_id) a.done[b] = 1;
            stoken = fed;
        }
    }
    __syncthreads();
    const int tok = stoken;
    if (tid < kHidden / 4)
        reinterpret_cast<float4*>(a.x_next + (size_t)b * kHidden)[tid] =
            reinterpret_cast<const float4*>(a.embed + (size_t)tok * kHidden)[tid];
}

// Decode-path fusion: x[row] += sum_s partial[s][row] (split-K partial sums of the preceding o_proj / down_proj GEMM,
// fixed summation order => deterministic), then RMSNorm(x[row]) -> bf16 hi/lo planes for the next GEMM.
// n_partial == 0 gives a plain RMSNorm.
// One CTA per row and 16-byte accesses (144 threads x float4 = 576 columns): up to 8 split-K partials are all in flight
// before the first add, the sum order is fixed (s = 0..S-1), and the 128 rows of a decode step spread over 128 SMs.
template <int S>
__global__ void __launch_bounds__(160) add_rmsnorm_row_kernel(float* __restrict__ x, const float* __restrict__ partial,
                                                              int M, const float* __restrict__ w,
                                                              bf16* __restrict__ hi, bf16* __restrict__ lo,
                                                              TraceBuf* trace, unsigned trace_id) {
    __shared__ float red[5];
    const int row = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool act = t < kHidden / 4;
    unsigned trec = kTraceNone;
    if (t == 0) trec = trace_open(trace, trace_id);
    pdl_trigger();
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) wv = __ldg(reinterpret_cast<const float4*>(w) + t);   // gains do not depend on the predecessor
    pdl_wait();
    if (t == 0 && first_cta()) trace_put(trace, trec, trace_id, TR_WAITED);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 p[S > 0 ? S : 1];
    if (act) {
        v = reinterpret_cast<const float4*>(x + (size_t)row * kHidden)[t];
#pragma unroll
        for (int s = 0; s < S; ++s)
            p[s] = reinterpret_cast<const float4*>(partial + ((size_t)s * M + row) * kHidden)[t];
#pragma unroll
        for (int s = 0; s < S; ++s) { v.x += p[s].x; v.y += p[s].y; v.z += p[s].z; v.w += p[s].w; }
    }
    const float sq = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    const float tot = (((red[0] + red[1]) + red[2]) + red[3]) + red[4];
    const float rstd = rsqrtf(tot * (1.0f / kHidden) + 1e-5f);
    if (act) {
        const size_t o = (size_t)row * kHidden + 4 * t;
        if (S > 0) reinterpret_cast<float4*>(x + (size_t)row * kHidden)[t] = v;
        store_planes2(hi, lo, o, wv.x * (v.x * rstd), wv.y * (v.y * rstd));
        store_planes2(hi, lo, o + 2, wv.z * (v.z * rstd), wv.w * (v.w * rstd));
    }
    if (t == 0) trace_close(trace, trec, trace_id);
}

// step += 1; records the first step count at which every row has emitted eos (wrapper.py:247-249) and, for the host's
// sparse poll, the number of finished rows (d_stop_step[1]).  assign != nullptr (B <= 128 = blockDim): also the work
// list of the next step's decode attention (DecodeAttnArgs::assign).  With n rows still decoding and m = B - n
// finished, every live row is cut into P = min(kAttnDynParts, B / n) key shares (P >= 3, else none: two shares
// measured no faster than just dropping the finished rows' streams): the row's own slot keeps share 0 (its keys from 0
// on: what it requests before its dependency wait stays valid), the j-th finished slot takes share 1 + j / n of the
// (j % n)-th live row, the remaining finished slots idle.  advance = 0: only (re)build the list (the decode loop does
// this once when it switches to the work-list attention kernel).
__global__ void __launch_bounds__(128) step_advance_kernel(int* d_step, const int* done, int B, int* d_stop_step, int* assign,
                                                           int advance) {
    __shared__ int wcnt[4];
    __shared__ int act[128];
    pdl_trigger();
    pdl_wait();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int unfinished = 0;
    for (int b = t; b < B; b += blockDim.x) unfinished |= !done[b];
    const int n_live = __syncthreads_count(unfinished);         // B <= 128: the number of live rows; else only 0 / non-0 matters
    if (assign != nullptr) {
        const bool live = t < B && !done[t];
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int pos = __popc(bal & ((1u << lane) - 1u));             // live rows before this one
        for (int w = 0; w < warp; ++w) pos += wcnt[w];
        const int n = n_live;
        if (live) act[pos] = t;
        __syncthreads();
        if (t < B) {
            int P = n > 0 ? B / n : 0;
            if (P > kAttnDynParts) P = kAttnDynParts;
            if (P < 3) P = 1;                                    // two shares do not pay for the merge (measured): plain skip
            const int j = t - pos;                               // finished rows before this one
            int v = 0;
            if (live) v = t | (P << 16);
            else if (n > 0 && 1 + j / n < P) v = act[j % n] | ((1 + j / n) << 8) | (P << 16);
            assign[t] = v;
        }
    }
    if (t == 0 && advance) {
        const int s = *d_step + 1;
        *d_step = s;
        if (n_live == 0 && *d_stop_step < 0) *d_stop_step = s;
        d_stop_step[1] = B <= 128 ? B - n_live : 0;
    }
}

// lm.model.embed_tokens (wrapper.py:237): one CTA per id, 144 x float4
__global__ void __launch_bounds__(144) embed_rows_kernel(const int* __restrict__ ids, const float* __restrict__ embed,
                                                         float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    int id = ids[blockIdx.x];
    id = id < 0 ? 0 : (id >= kVocab ? kVocab - 1 : id);
    reinterpret_cast<float4*>(out + (size_t)blockIdx.x * kHidden)[threadIdx.x] =
        reinterpret_cast<const float4*>(embed + (size_t)id * kHidden)[threadIdx.x];
}

}  // namespace

cudaError_t launch_embed_rows(const int* ids, int n, const float* embed, float* out, cudaStream_t st) {
    return launch_k(embed_rows_kernel, dim3(n), dim3(144), 0, st, ids, embed, out);
}

cudaError_t launch_prefix(const float* rows33, const int* ids, const float* embed, int B, float* prefix,
                          cudaStream_t st) {
    dim3 grid(kPrefix, B);
    return launch_k(prefix_kernel, grid, dim3(144), 0, st, rows33, ids, embed, B, prefix);
}

template <typename T, int ST>
cudaError_t launch_decode_attention_share(const DecodeAttnArgs& a, dim3 grid, cudaStream_t st) {
    static bool configured[kMaxDevices] = {};
    auto kern = decode_attention_share_kernel<T, ST, true, true>;
    if (cudaError_t e = ensure_smem(kern, sizeof(DecodeSmemW<T, ST, true>), configured); e != cudaSuccess) return e;
    return launch_k(kern, grid, dim3(128), sizeof(DecodeSmemW<T, ST, true>), st, a);
}
template <typename T, int ST, bool BULK>
cudaError_t launch_decode_attention_warp(const DecodeAttnArgs& a, dim3 grid, cudaStream_t st) {
    if (a.assign != nullptr) {                            // the work-list kernel exists for the product kernel (bulk copies) only
        if constexpr (BULK) return launch_decode_attention_share<T, ST>(a, grid, st);
        else return cudaErrorNotSupported;
    }
    static bool configured[kMaxDevices] = {};
    auto kern = decode_attention_warp_kernel<T, ST, BULK>;
    if (cudaError_t e = ensure_smem(kern, sizeof(DecodeSmemW<T, ST, BULK>), configured); e != cudaSuccess) return e;
    return launch_k(kern, grid, dim3(128), sizeof(DecodeSmemW<T, ST, BULK>), st, a);
}

cudaError_t launch_decode_attention(const DecodeAttnArgs& a, cudaStream_t st) {
    dim3 grid(a.nsplit, kKvHeads, a.B);
    cudaError_t e;
    if (a.variant == 2) {                                 // warp-autonomous, bulk copies; ring depth 2 = 3 CTAs per SM
        e = a.kv_fmt == kKvBf16 ? launch_decode_attention_warp<bf16, 3, true>(a, grid, st)
          : a.kv_fmt == kKvF24 ? launch_decode_attention_warp<kv24, 2, true>(a, grid, st)
                               : launch_decode_attention_warp<float, 2, true>(a, grid, st);
    } else {
#ifdef MB_LAB
        if (a.variant == 1) {                             // warp-autonomous, cp.async pieces
            e = a.kv_fmt == kKvBf16 ? launch_decode_attention_warp<bf16, 3, false>(a, grid, st)
              : a.kv_fmt == kKvF24 ? launch_decode_attention_warp<kv24, 2, false>(a, grid, st)
                                   : launch_decode_attention_warp<float, 2, false>(a, grid, st);
        } else {
            static bool c0[kMaxDevices] = {}, c1[kMaxDevices] = {}, c2[kMaxDevices] = {};
            if (cudaError_t e0 = ensure_smem(decode_attention_kernel<float>, sizeof(DecodeSmem<float>), c0); e0 != cudaSuccess) return e0;
            if (cudaError_t e0 = ensure_smem(decode_attention_kernel<bf16>, sizeof(DecodeSmem<bf16>), c1); e0 != cudaSuccess) return e0;
            if (cudaError_t e0 = ensure_smem(decode_attention_kernel<kv24>, sizeof(DecodeSmem<kv24>), c2); e0 != cudaSuccess) return e0;
            e = a.kv_fmt == kKvBf16 ? launch_k(decode_attention_kernel<bf16>, grid, dim3(128), sizeof(DecodeSmem<bf16>), st, a)
              : a.kv_fmt == kKvF24 ? launch_k(decode_attention_kernel<kv24>, grid, dim3(128), sizeof(DecodeSmem<kv24>), st, a)
                                   : launch_k(decode_attention_kernel<float>, grid, dim3(128), sizeof(DecodeSmem<float>), st, a);
        }
#else
        return cudaErrorNotSupported;                     // variants 0 / 1 exist in lab builds only (MB_BUILD_LAB=1)
#endif
    }
    if (e != cudaSuccess || a.nsplit == 1) return e;
    return launch_k(decode_combine_kernel, dim3(kHeads, a.B), dim3(64), 0, st, a);
}

cudaError_t launch_add_rmsnorm(float* x, const float* partial, int n_partial, int M, const float* w, bf16* hi, bf16* lo,
                               cudaStream_t st, TraceBuf* trace, unsigned trace_id) {
#define MB_ROWNORM(S) launch_k(add_rmsnorm_row_kernel<S>, dim3(M), dim3(160), 0, st, x, partial, M, w, hi, lo, trace, trace_id)
    switch (n_partial) {
        case 0: return MB_ROWNORM(0);
        case 3: return MB_ROWNORM(3);
        case 4: return MB_ROWNORM(4);
        case 8: return MB_ROWNORM(8);
        default: return cudaErrorInvalidValue;
    }
#undef MB_ROWNORM
}

cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st) {
    return launch_k(sample_kernel, dim3(a.B), dim3(1024), 0, st, a);
}

cudaError_t launch_step_advance(int* d_step, const int* done, int B, int* d_stop_step, int* assign, int advance, cudaStream_t st) {
    if (assign != nullptr && B > 128) return cudaErrorInvalidValue;
    return launch_k(step_advance_kernel, dim3(1), dim3(128), 0, st, d_step, done, B, d_stop_step, assign, advance);
}

}  // namespace mb
