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

// ---------------------------------------------------------------------------------------------------------------
// Causal prefill attention, fp32 on CUDA cores: one CTA = 64 query rows of one (batch, head); one thread per query.
// K/V tiles of 32 keys are staged in shared memory (converted to fp32) and read as warp-wide broadcasts.
template <typename T>
__global__ void __launch_bounds__(64) prefill_attention_kernel(const float* __restrict__ q, const T* __restrict__ kc,
                                                               const T* __restrict__ vc, int S, int t_max,
                                                               bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
    constexpr int TK = 32;
    __shared__ __align__(16) float sk[TK][kHeadDim];
    __shared__ __align__(16) float sv[TK][kHeadDim];
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x;
    const int r = qt * 64 + tid;
    const bool active = r < S;
    const int kvh = h / (kHeads / kKvHeads);
    const T* kb = kc + ((size_t)b * kKvHeads + kvh) * t_max * kHeadDim;
    const T* vb = vc + ((size_t)b * kKvHeads + kvh) * t_max * kHeadDim;
    float qr[kHeadDim], acc[kHeadDim];
    const float* qp = q + ((size_t)b * S + (active ? r : 0)) * kHidden + h * kHeadDim;
    pdl_trigger();
    pdl_wait();
#pragma unroll
    for (int d = 0; d < kHeadDim; d += 4) {
        const float4 a = *reinterpret_cast<const float4*>(qp + d);
        qr[d] = a.x; qr[d + 1] = a.y; qr[d + 2] = a.z; qr[d + 3] = a.w;
        acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
    }
    float m_run = -INFINITY, l_run = 0.f;
    const int kend = min(S, qt * 64 + 64);
    for (int k0 = 0; k0 < kend; k0 += TK) {
        __syncthreads();
        for (int e = tid; e < TK * kHeadDim; e += 64) {
            const int j = e >> 6, d = e & 63;
            const int key = k0 + j;
            sk[j][d] = key < S ? kv_load(kb + (size_t)key * kHeadDim + d) : 0.f;
            sv[j][d] = key < S ? kv_load(vb + (size_t)key * kHeadDim + d) : 0.f;
        }
        __syncthreads();
        float s[TK];
        float mt = -INFINITY;
#pragma unroll
        for (int j = 0; j < TK; ++j) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < kHeadDim; d += 4) {
                const float4 k4 = *reinterpret_cast<const float4*>(&sk[j][d]);
                a += qr[d] * k4.x; a += qr[d + 1] * k4.y; a += qr[d + 2] * k4.z; a += qr[d + 3] * k4.w;
            }
            a *= 0.125f;                                            // head_dim^-0.5
            if (k0 + j > r) a = -INFINITY;                          // causal mask
            s[j] = a;
            mt = fmaxf(mt, a);
        }
        if (mt == -INFINITY) continue;                              // whole tile is in this row's future (uniform bar count kept above)
        const float m_new = fmaxf(m_run, mt);
        const float alpha = expf(m_run - m_new);
        float lsum = 0.f;
#pragma unroll
        for (int j = 0; j < TK; ++j) { s[j] = expf(s[j] - m_new); lsum += s[j]; }
        l_run = l_run * alpha + lsum;
        m_run = m_new;
#pragma unroll
        for (int d = 0; d < kHeadDim; ++d) acc[d] *= alpha;
#pragma unroll
        for (int j = 0; j < TK; ++j) {
            const float p = s[j];
#pragma unroll
            for (int d = 0; d < kHeadDim; d += 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(&sv[j][d]);
                acc[d] += p * v4.x; acc[d + 1] += p * v4.y; acc[d + 2] += p * v4.z; acc[d + 3] += p * v4.w;
            }
        }
    }
    if (!active) return;
    const float inv = 1.0f / l_run;
    const size_t ob = ((size_t)b * S + r) * kHidden + h * kHeadDim;
#pragma unroll
    for (int d = 0; d < kHeadDim; d += 2) store_planes2(out_hi, out_lo, ob + d, acc[d] * inv, acc[d + 1] * inv);
}

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
    __align__(16) float qs[3 * kHeadDim];       // fused-QKV mode: roped queries of the 3 heads, new key / value row
    __align__(16) float knew[kHeadDim];
    __align__(16) float vnew[kHeadDim];
};

// Three CTAs per SM (<= 170 registers): the fused-QKV prologue pushed the fp32 variant to 207 registers = two CTAs per SM
// without the bound.  Four CTAs of the 24-bit variant (128 registers, unpadded rows) were measured slower: 23.4 vs 18.7 us.
template <typename T>
__global__ void __launch_bounds__(128, 3) decode_attention_kernel(const DecodeAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    pdl_wait();
    if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, TR_WAITED);
    // fused-QKV mode: this row's q | k | v columns (320 = 80 float4) arrive as split-K partial sums.  Their addresses, and
    // that of the RoPE row of the new position (rope_cur), do not depend on the step counter, so all these loads are in
    // flight together with the counter / stop-flag loads: one L2 round trip instead of three.
    const bool fused = a.qkv_part != nullptr;
    float4 pv[kQkvSplitMax];
    float2 c2 = make_float2(1.f, 1.f), s2 = make_float2(0.f, 0.f);
    int qcol = 0;
    if (fused && tid < 80) {
        qcol = tid < 48 ? kvh * 192 + tid * 4
                        : (tid < 64 ? kHidden + kvh * kHeadDim + (tid - 48) * 4
                                    : kHidden + kKvHeads * kHeadDim + kvh * kHeadDim + (tid - 64) * 4);
        const float* pp = a.qkv_part + (size_t)b * kQkvDim + qcol;
        const size_t zs = (size_t)a.B * kQkvDim;
#pragma unroll
        for (int z = 0; z < kQkvSplitMax; ++z)
            pv[z] = z < a.qkv_nsplit ? *reinterpret_cast<const float4*>(pp + z * zs) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < 64) {                                             // RoPE angle index of this thread's two pairs
            const int i = (qcol & (kHeadDim - 1)) >> 1;
            c2 = *reinterpret_cast<const float2*>(a.rope_cur + i);
            s2 = *reinterpret_cast<const float2*>(a.rope_cur + 32 + i);
        }
    }
    const int step_now = a.d_step ? *a.d_step : 0;
    // a row that has emitted the stop token is finished (the reference cuts its text there,
    // wrapper.py:254); its later tokens are never read, so its K/V stream -- the dominant decode traffic -- is skipped.
    // The uniform exit happens after the cp.async groups above are drained by the hardware at CTA exit.
    if (a.done && a.done[b]) { cp_async_commit(); cp_async_wait<0>(); return; }
    const int ctx = a.ctx_base + step_now;
    const int ntiles = (ctx + 63) >> 6;
    const int t_end = min(ntiles, t_begin + a.tps);
    // fused-QKV mode: key ctx-1 is produced by this kernel, so the cache is only read up to ctx-2
    const int ctx_ld = fused ? ctx - 1 : ctx;
    if (!early0 && t_begin < t_end) load_tile(0, t_begin, ctx_ld);
    cp_async_commit();
    if (!early1 && t_begin + 1 < t_end) load_tile(1, t_begin + 1, ctx_ld);
    cp_async_commit();
    const int t_new = (ctx - 1) >> 6;                               // tile that holds the new key
    if (fused) {
        if (tid < 80) {
            float4 acc4 = pv[0];                                    // fixed summation order z = 0..nsplit-1
#pragma unroll
            for (int z = 1; z < kQkvSplitMax; ++z)
                if (z < a.qkv_nsplit) { acc4.x += pv[z].x; acc4.y += pv[z].y; acc4.z += pv[z].z; acc4.w += pv[z].w; }
            if (tid < 64) {                                         // RoPE on q and k: pairs (2i, 2i+1) rotate by angle i
                const float x0 = acc4.x * c2.x - acc4.y * s2.x, x1 = acc4.y * c2.x + acc4.x * s2.x;
                const float x2 = acc4.z * c2.y - acc4.w * s2.y, x3 = acc4.w * c2.y + acc4.z * s2.y;
                acc4 = make_float4(x0, x1, x2, x3);
            }
            if (tid < 48) {
                *reinterpret_cast<float4*>(&sm.qs[tid * 4]) = acc4;
            } else {
                const int dd = (tid < 64 ? tid - 48 : tid - 64) * 4;
                float* dst = tid < 64 ? &sm.knew[dd] : &sm.vnew[dd];
                const bool owner = t_new >= t_begin && t_new < t_begin + a.tps;   // the split that owns the new key appends it to the cache
                unsigned char* crow = const_cast<unsigned char*>(tid < 64 ? kb : vb) + (size_t)(ctx - 1) * ROWB;
                // knew / vnew hold what later steps will read back from the cache (rounded to the cache format)
                if constexpr (F24) {
                    const uint32_t u0 = f24_bits(acc4.x), u1 = f24_bits(acc4.y), u2 = f24_bits(acc4.z), u3 = f24_bits(acc4.w);
                    dst[0] = __uint_as_float(u0); dst[1] = __uint_as_float(u1); dst[2] = __uint_as_float(u2); dst[3] = __uint_as_float(u3);
                    if (owner) {
                        *reinterpret_cast<uint2*>(crow + dd * 2) = make_uint2((u0 >> 16) | (u1 & 0xFFFF0000u), (u2 >> 16) | (u3 & 0xFFFF0000u));
                        *reinterpret_cast<uint32_t*>(crow + 128 + dd) =
                            ((u0 >> 8) & 0xFFu) | (u1 & 0xFF00u) | ((u2 << 8) & 0xFF0000u) | ((u3 << 16) & 0xFF000000u);
                    }
                } else {
                    T r0, r1, r2, r3;
                    kv_cast(acc4.x, r0); kv_cast(acc4.y, r1); kv_cast(acc4.z, r2); kv_cast(acc4.w, r3);
                    dst[0] = kv_load(&r0); dst[1] = kv_load(&r1); dst[2] = kv_load(&r2); dst[3] = kv_load(&r3);
                    if (owner) {
                        T* cb = reinterpret_cast<T*>(crow) + dd;
                        cb[0] = r0; cb[1] = r1; cb[2] = r2; cb[3] = r3;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && first_cta()) trace_put(a.trace, trec, a.trace_id, 4);       // q / k / v of this step ready
    }

    // score role: thread = (key j, half); each half owns every other 16 B chunk of the key row.  The matching
    // slices of the three query heads live in registers.
    const int sj = tid >> 1, shalf = tid & 1;
    float qreg[3][kHeadDim / 2];
    {
        const float* qb = fused ? sm.qs : a.q + (size_t)b * kHidden + (kvh * 3) * kHeadDim;
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
        if (fused && t == t_new) {                                   // uniform: place the new key / value row into the tile
            const int j = (ctx - 1) & 63;
            const int d = tid & (kHeadDim - 1);
            unsigned char* row = tid < kHeadDim ? sm.k[buf][j] : sm.v[buf][j];
            const float val = tid < kHeadDim ? sm.knew[d] : sm.vnew[d];
            if constexpr (F24) {
                const uint32_t u = __float_as_uint(val);             // already rounded to 24 bits
                reinterpret_cast<unsigned short*>(row)[d] = (unsigned short)(u >> 16);
                row[128 + d] = (unsigned char)(u >> 8);
            } else {
                kv_cast(val, reinterpret_cast<T*>(row)[d]);
            }
            __syncthreads();
        }
        {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
            for (int i = 0; i < NCH / 2; ++i) {
                if constexpr (F24) {
                    // 8 values: one 16 B read of upper halves, one 8 B read of mantissa bytes
                    const unsigned char* rowp = sm.k[buf][sj];
                    const uint4 hv = *reinterpret_cast<const uint4*>(rowp + (2 * i + shalf) * 16);
                    const uint2 lv = *reinterpret_cast<const uint2*>(rowp + 128 + (2 * i + shalf) * 8);
                    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const uint32_t lb = p < 2 ? lv.x >> (16 * p) : lv.y >> (16 * (p - 2));
                        const float k0 = __uint_as_float((hw[p] << 16) | ((lb & 0xFFu) << 8));
                        const float k1 = __uint_as_float((hw[p] & 0xFFFF0000u) | (lb & 0xFF00u));
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
                        v0 = __uint_as_float((hw << 16) | ((lb & 0xFFu) << 8));
                        v1 = __uint_as_float((hw & 0xFFFF0000u) | (lb & 0xFF00u));
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
            if (bi == a.eos_id) a.done[b] = 1;
            stoken = a.forced ? a.forced[(size_t)b * a.max_len + step] : bi;
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
// n_partial == 0 gives a plain RMSNorm.  One warp per row.
template <int S>
__global__ void __launch_bounds__(128) add_rmsnorm_kernel(float* __restrict__ x, const float* __restrict__ partial,
                                                          int M, const float* __restrict__ w,
                                                          bf16* __restrict__ hi, bf16* __restrict__ lo,
                                                          TraceBuf* trace, unsigned trace_id) {
    constexpr int PER = kHidden / 32;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    unsigned trec = kTraceNone;
    if (threadIdx.x == 0) trec = trace_open(trace, trace_id);
    if (row >= M) return;
    float v[PER], p[S > 0 ? S : 1][PER], wv[PER];
    float* xr = x + (size_t)row * kHidden;
    pdl_trigger();
#pragma unroll
    for (int j = 0; j < PER; ++j) wv[j] = w[lane + 32 * j];        // weights do not depend on the predecessor
    pdl_wait();
    if (threadIdx.x == 0 && first_cta()) trace_put(trace, trec, trace_id, TR_WAITED);
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = xr[lane + 32 * j];
#pragma unroll
    for (int s = 0; s < S; ++s)                                    // all loads in flight before the first add
#pragma unroll
        for (int j = 0; j < PER; ++j) p[s][j] = partial[((size_t)s * M + row) * kHidden + lane + 32 * j];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < PER; ++j) v[j] += p[s][j];
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) sq += v[j] * v[j];
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / kHidden) + 1e-5f);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int i = lane + 32 * j;
        if (S > 0) xr[i] = v[j];
        store_planes1(hi, lo, (size_t)row * kHidden + i, wv[j] * (v[j] * rstd));
    }
    if (threadIdx.x == 0) trace_close(trace, trec, trace_id);
}

// Same operation with one CTA per row and 16-byte accesses (144 threads x float4 = 576 columns): up to 12 split-K
// partials are all in flight before the first add, the sum order is fixed (s = 0..S-1), and the 128 rows of a decode
// step spread over 128 SMs instead of 32.
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

// step += 1; records the first step count at which every row has emitted eos (wrapper.py:247-249)
__global__ void step_advance_kernel(int* d_step, const int* done, int B, int* d_stop_step, const float* rope_cos,
                                    const float* rope_sin, int pos_base, float* rope_cur) {
    __shared__ int all;
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) all = 1;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x)
        if (!done[b]) all = 0;
    __syncthreads();
    const int s = *d_step + 1;                                     // every thread reads the old value before thread 0 writes
    if (rope_cur && threadIdx.x < 64 && pos_base + s < kMaxPos)
        rope_cur[threadIdx.x] = threadIdx.x < 32 ? rope_cos[(pos_base + s) * 32 + threadIdx.x]
                                                 : rope_sin[(pos_base + s) * 32 + threadIdx.x - 32];
    __syncthreads();
    if (threadIdx.x == 0) {
        *d_step = s;
        if (all && *d_stop_step < 0) *d_stop_step = s;
    }
}

}  // namespace

cudaError_t launch_prefix(const float* rows33, const int* ids, const float* embed, int B, float* prefix,
                          cudaStream_t st) {
    dim3 grid(kPrefix, B);
    return launch_k(prefix_kernel, grid, dim3(144), 0, st, rows33, ids, embed, B, prefix);
}

cudaError_t launch_prefill_attention(const float* q, const void* kc, const void* vc, int kv_fmt, int B, int S,
                                     int t_max, bf16* out_hi, bf16* out_lo, cudaStream_t st) {
    dim3 grid((S + 63) / 64, kHeads, B);
    if (kv_fmt == kKvF24) return cudaErrorNotSupported;            // the CUDA-core A/B kernel reads element-wise formats only
    if (kv_fmt)
        return launch_k(prefill_attention_kernel<bf16>, grid, dim3(64), 0, st, q, (const bf16*)kc, (const bf16*)vc, S, t_max, out_hi, out_lo);
    return launch_k(prefill_attention_kernel<float>, grid, dim3(64), 0, st, q, (const float*)kc, (const float*)vc, S, t_max, out_hi, out_lo);
}

cudaError_t launch_decode_attention(const DecodeAttnArgs& a, cudaStream_t st) {
    dim3 grid(a.nsplit, kKvHeads, a.B);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(decode_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(DecodeSmem<float>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_attention_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(DecodeSmem<bf16>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_attention_kernel<kv24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(DecodeSmem<kv24>));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaError_t e = a.kv_fmt == kKvBf16 ? launch_k(decode_attention_kernel<bf16>, grid, dim3(128), sizeof(DecodeSmem<bf16>), st, a)
                  : a.kv_fmt == kKvF24 ? launch_k(decode_attention_kernel<kv24>, grid, dim3(128), sizeof(DecodeSmem<kv24>), st, a)
                                       : launch_k(decode_attention_kernel<float>, grid, dim3(128), sizeof(DecodeSmem<float>), st, a);
    if (e != cudaSuccess || a.nsplit == 1) return e;
    return launch_k(decode_combine_kernel, dim3(kHeads, a.B), dim3(64), 0, st, a);
}

cudaError_t launch_add_rmsnorm(float* x, const float* partial, int n_partial, int M, const float* w, bf16* hi, bf16* lo,
                               cudaStream_t st, TraceBuf* trace, unsigned trace_id) {
    static const bool v1 = getenv("MB_NORM_V1") != nullptr;       // warp-per-row version, kept for A/B measurements
    if (!v1 || n_partial > 4) {
#define MB_ROWNORM(S) launch_k(add_rmsnorm_row_kernel<S>, dim3(M), dim3(160), 0, st, x, partial, M, w, hi, lo, trace, trace_id)
        switch (n_partial) {
            case 0: return MB_ROWNORM(0);
            case 3: return MB_ROWNORM(3);
            case 4: return MB_ROWNORM(4);
            case 6: return MB_ROWNORM(6);
            case 8: return MB_ROWNORM(8);
            case 9: return MB_ROWNORM(9);
            case 12: return MB_ROWNORM(12);
            default: return cudaErrorInvalidValue;
        }
#undef MB_ROWNORM
    }
    const int grid = (M + 3) / 4;
    switch (n_partial) {
        case 0: return launch_k(add_rmsnorm_kernel<0>, dim3(grid), dim3(128), 0, st, x, partial, M, w, hi, lo, trace, trace_id);
        case 3: return launch_k(add_rmsnorm_kernel<3>, dim3(grid), dim3(128), 0, st, x, partial, M, w, hi, lo, trace, trace_id);
        case 4: return launch_k(add_rmsnorm_kernel<4>, dim3(grid), dim3(128), 0, st, x, partial, M, w, hi, lo, trace, trace_id);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st) {
    return launch_k(sample_kernel, dim3(a.B), dim3(1024), 0, st, a);
}

cudaError_t launch_step_advance(int* d_step, const int* done, int B, int* d_stop_step, const float* rope_cos,
                                const float* rope_sin, int pos_base, float* rope_cur, cudaStream_t st) {
    return launch_k(step_advance_kernel, dim3(1), dim3(128), 0, st, d_step, done, B, d_stop_step, rope_cos, rope_sin,
                    pos_base, rope_cur);
}

}  // namespace mb
