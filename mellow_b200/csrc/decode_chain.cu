// Fused decode layer chain for B <= 128 rows (one 128-row UMMA tile): everything between two attention kernels,
//
//     o_proj (split-K)  ->  residual + RMSNorm  ->  gate/up (+SwiGLU)  ->  down (split-K)  ->  residual + RMSNorm
//                       ->  next layer's QKV (+RoPE, +KV-cache write)
//
// runs as ONE persistent kernel (one CTA per SM) instead of six dependent launches.  The phases are separated by
// software grid barriers (a release/acquire counter per phase) that cost ~1 us instead of a kernel boundary, and the
// TMA producer uses the wait to prefetch the next GEMM's weight tiles, which never depend on the previous phase.
// GEMM phases are the same tcgen05 pipeline as gemm_umma.cu (TMA ring -> tcgen05.mma -> TMEM -> row epilogue), with
// the tile width (16 / 32 columns), split-K factor and epilogue chosen per phase at run time.
#include "kernels.cuh"
#include "umma.cuh"

namespace mb {

using namespace umma;

namespace {

constexpr int kStages = 5;
constexpr uint32_t kBMax = 32 * BK * 2;                                  // widest weight tile: 32 rows x 128 B
constexpr uint32_t kStageBytes = 2 * (A_BYTES + kBMax);                  // hi + lo planes of A and W
constexpr uint32_t kAccCols = 32;
constexpr size_t kChainSmem = (size_t)kStages * kStageBytes + 1024 + 256;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one thread: publish this CTA's writes of the phase and count it in
__device__ __forceinline__ void grid_arrive(unsigned* c) {
    asm volatile("fence.proxy.async;" ::: "memory");                     // later TMA (async proxy) reads see generic writes
    __threadfence();
    atomicAdd(c, 1u);
}
__device__ __forceinline__ void grid_wait(const unsigned* c, unsigned target) {
    const long long t0 = clock64();
    while (ld_acquire(c) < target) {
        if (clock64() - t0 > 4000000000LL) __trap();                     // a broken chain must fail, not hang the GPU
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }

struct Tile { bool has; int n0, z, kb_begin, KB; };

__device__ __forceinline__ Tile chain_tile(const ChainOp& op) {
    Tile t;
    const int tiles_n = (op.g.N + op.bn - 1) / op.bn;
    const int nsplit = op.g.split_k > 1 ? op.g.split_k : 1;
    t.has = (int)blockIdx.x < tiles_n * nsplit;
    t.z = blockIdx.x / tiles_n;
    t.n0 = ((int)blockIdx.x - t.z * tiles_n) * op.bn;
    const int kb_all = (op.g.K + BK - 1) / BK;
    t.kb_begin = (int)(((long long)kb_all * t.z) / nsplit);
    t.KB = (int)(((long long)kb_all * (t.z + 1)) / nsplit) - t.kb_begin;
    return t;
}

__global__ void __launch_bounds__(kThreads, 1)
decode_chain_kernel(const __grid_constant__ ChainMaps tm, const __grid_constant__ ChainArgs ca) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* tfull = empty + kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool split = ca.split != 0;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], kEpiWarps); mbar_init(&tempty[1], kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * kAccCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int o = 0; o < ca.n_ops; ++o) {
                const ChainOp& op = ca.op[o];
                if (op.kind != CH_GEMM) continue;
                const Tile t = chain_tile(op);
                if (!t.has) continue;
                const uint32_t b_bytes = (uint32_t)op.bn * BK * 2;
                const uint32_t stage_tx = (split ? 2u : 1u) * (A_BYTES + b_bytes);
                const CUtensorMap* ma = &tm.m[op.map_a];
                const CUtensorMap* mb_ = &tm.m[op.map_b];
                // weight tiles first (they never depend on the previous phase), then the dependency, then activations
                const int npre = t.KB < kStages ? t.KB : kStages;
                for (int i = 0; i < npre; ++i) {
                    const int s = (it + i) % kStages;
                    mbar_wait(&empty[s], (((it + i) / kStages) & 1) ^ 1);
                    unsigned char* st = smem + (size_t)s * kStageBytes;
                    mbar_expect_tx(&full[s], stage_tx);
                    tma_load_2d(st + A_BYTES, mb_, &full[s], (t.kb_begin + i) * BK, t.n0);
                    if (split) tma_load_2d(st + 2 * A_BYTES + kBMax, mb_ + 1, &full[s], (t.kb_begin + i) * BK, t.n0);
                }
                if (o == 0) pdl_wait(); else grid_wait(&ca.bar[o - 1], gridDim.x);
                for (int i = 0; i < npre; ++i) {
                    const int s = (it + i) % kStages;
                    unsigned char* st = smem + (size_t)s * kStageBytes;
                    tma_load_2d(st, ma, &full[s], (t.kb_begin + i) * BK, 0);
                    if (split) tma_load_2d(st + A_BYTES + kBMax, ma + 1, &full[s], (t.kb_begin + i) * BK, 0);
                }
                for (int kb = npre; kb < t.KB; ++kb) {
                    const int s = (it + kb) % kStages;
                    mbar_wait(&empty[s], (((it + kb) / kStages) & 1) ^ 1);
                    unsigned char* st = smem + (size_t)s * kStageBytes;
                    mbar_expect_tx(&full[s], stage_tx);
                    tma_load_2d(st, ma, &full[s], (t.kb_begin + kb) * BK, 0);
                    tma_load_2d(st + A_BYTES, mb_, &full[s], (t.kb_begin + kb) * BK, t.n0);
                    if (split) {
                        tma_load_2d(st + A_BYTES + kBMax, ma + 1, &full[s], (t.kb_begin + kb) * BK, 0);
                        tma_load_2d(st + 2 * A_BYTES + kBMax, mb_ + 1, &full[s], (t.kb_begin + kb) * BK, t.n0);
                    }
                }
                it += t.KB;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int it = 0, lt = 0;
            for (int o = 0; o < ca.n_ops; ++o) {
                const ChainOp& op = ca.op[o];
                if (op.kind != CH_GEMM) continue;
                const Tile t = chain_tile(op);
                if (!t.has) continue;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(op.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
                const int buf = lt & 1;
                mbar_wait(&tempty[buf], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)buf * kAccCols;
                for (int kb = 0; kb < t.KB; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&full[s], (it / kStages) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes);
                    const uint32_t b_hi = a_hi + A_BYTES;
                    const uint32_t a_lo = b_hi + kBMax;
                    const uint32_t b_lo = a_lo + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t off = k * 32;
                        umma_bf16(tacc, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, (kb | k) != 0);
                        if (split) {
                            umma_bf16(tacc, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, 1);
                            umma_bf16(tacc, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, 1);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tfull[buf]);
                ++lt;
            }
        }
    } else {
        const int ew = warp - 2;                                      // 0..7
        const int q = warp & 3;                                       // TMEM lane quadrant
        const int half = ew >> 2;                                     // which 16-column half of a 32-wide tile
        const bool leader = (threadIdx.x == 64);
        pdl_wait();
        int lt = 0;
        for (int o = 0; o < ca.n_ops; ++o) {
            const ChainOp& op = ca.op[o];
            if (op.kind == CH_GEMM) {
                const Tile t = chain_tile(op);
                if (t.has) {
                    const GemmArgs& g = op.g;
                    const int buf = lt & 1;
                    mbar_wait(&tfull[buf], (lt >> 1) & 1);
                    tc_fence_after();
                    const bool mine = half * 16 < op.bn;
                    float v[16];
                    if (mine) tmem_ld16(tmem_base + (uint32_t)buf * kAccCols + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16), v);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[buf]);
                    const int m = q * 32 + lane, n = t.n0 + half * 16;
                    if (mine && m < g.M && n < g.N) {
                        if (g.split_k > 1) {
                            float* pz = g.partial + (size_t)t.z * g.M * g.N + (size_t)m * g.N + n;
#pragma unroll
                            for (int j = 0; j < 16; j += 4) st4(pz + j, v + j);
                        } else if (op.epi == EPI_SWIGLU) {
                            epilogue_row16<EPI_SWIGLU>(g, m, n, v);
                        } else if (op.epi == EPI_QKV_ROPE) {
                            epilogue_row16<EPI_QKV_ROPE>(g, m, n, v);
                        } else {
                            epilogue_row16<EPI_GENERIC>(g, m, n, v);
                        }
                    }
                    ++lt;
                }
            } else {
                // residual + split-K reduction + RMSNorm: x[row] += sum_s partial[s][row]; planes = RMSNorm(x[row]) * w
                if (leader) grid_wait(&ca.bar[o - 1], gridDim.x);
                epi_bar_sync();
                const int row = (int)blockIdx.x * kEpiWarps + ew;
                const int M = op.g.M;
                if (row < M) {
                    constexpr int PER = kHidden / 32;
                    float xv[PER], wv[PER], pv[4][PER];
                    float* xr = op.x + (size_t)row * kHidden;
#pragma unroll
                    for (int j = 0; j < PER; ++j) { xv[j] = xr[lane + 32 * j]; wv[j] = op.w[lane + 32 * j]; }
#pragma unroll
                    for (int s = 0; s < 4; ++s)
#pragma unroll
                        for (int j = 0; j < PER; ++j)
                            pv[s][j] = s < op.n_partial ? op.partial[((size_t)s * M + row) * kHidden + lane + 32 * j] : 0.f;
#pragma unroll
                    for (int s = 0; s < 4; ++s)
#pragma unroll
                        for (int j = 0; j < PER; ++j)
                            if (s < op.n_partial) xv[j] += pv[s][j];
                    float sq = 0.f;
#pragma unroll
                    for (int j = 0; j < PER; ++j) sq += xv[j] * xv[j];
                    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / kHidden) + 1e-5f);
#pragma unroll
                    for (int j = 0; j < PER; ++j) {
                        const int i = lane + 32 * j;
                        xr[i] = xv[j];
                        store_planes1(op.hi, op.lo, (size_t)row * kHidden + i, wv[j] * (xv[j] * rstd));
                    }
                }
            }
            // end of phase: all epilogue warps of this CTA are done with their global writes
            if (o + 1 < ca.n_ops) {
                __threadfence();
                epi_bar_sync();
                if (leader) grid_arrive(&ca.bar[o]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kAccCols) : "memory");
    }
}

}  // namespace

cudaError_t build_chain_map(CUtensorMap* map, const bf16* ptr, int rows, int K, int ld, int box_rows) {
    return make_map(map, ptr, rows, K, ld, box_rows) ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_decode_chain(const ChainMaps& maps, const ChainArgs& args, cudaStream_t st) {
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmem);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    // every phase's tile count must fit one tile per CTA
    for (int o = 0; o < args.n_ops; ++o) {
        const ChainOp& op = args.op[o];
        if (op.kind != CH_GEMM) continue;
        const int tiles = ((op.g.N + op.bn - 1) / op.bn) * (op.g.split_k > 1 ? op.g.split_k : 1);
        if (tiles > num_sms || op.g.M > BM || (op.bn != 16 && op.bn != 32)) return cudaErrorInvalidValue;
    }
    return launch_k(decode_chain_kernel, dim3(num_sms), dim3(kThreads), kChainSmem, st, maps, args);
}

}  // namespace mb
