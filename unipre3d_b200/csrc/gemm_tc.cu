// gemm_tc.cu -- the dense per-token Linear layers of the point backbone on the 5th-generation tensor cores.
//
// Reference arithmetic: nn.Linear inside Attention / Mlp / Encoder of
// /root/reference/openpoints/models/backbone/transformer.py:10-77 (qkv, proj, fc1 -> GELU -> fc2) and :210-243
// (Conv1d(k=1) of the mini-PointNet), forward and the dX half of the backward.
//
// One CTA computes a 128 x BN tile of  D = A * op(B) (+ bias) (+ epilogue):
//   * operands are bf16, staged global -> shared by TMA tensor copies (cp.async.bulk.tensor.2d, 128-byte swizzle) into a
//     4-deep ring guarded by full/empty mbarriers;
//   * ONE elected thread issues tcgen05.mma (cta_group::1, kind::f16, M = 128, N = BN, K = 16 per instruction) reading
//     the shared-memory tiles through matrix descriptors; the fp32 accumulator lives in TMEM (BN columns x 128 lanes);
//   * tcgen05.commit releases ring slots back to the producer and finally signals the four epilogue warps, which read
//     the accumulator with tcgen05.ld (32 lanes x 32 columns per instruction: one thread = one output row), apply the
//     fused epilogue into swizzled shared-memory slabs and hand them to TMA stores (rows past T are clipped by the
//     tensor map, so ragged row counts need no guards).
//   * launched with programmatic stream serialization: barrier/TMEM set-up and the weight tiles of the first ring pass
//     are in flight while the previous kernel of the stream drains; griddepcontrol.wait precedes the first access to
//     anything that kernel produced.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue (two per lane quarter).
//
// op(B):  B_KMAJOR   B is (N, K) row-major  (nn.Linear weight as stored: the forward  y = x W^T)
//         B_NMAJOR   B is (K, N) row-major  (the same weight read for the backward  dx = dy W; "MN-major" operand)
// Epilogues (all in fp32 on the accumulator, one rounding at the store):
//   EPI_NONE      out = acc (+ bias)
//   EPI_GELU      aux_out = bf16(acc + bias) ; out = gelu(aux_out)                (fc1 -> GELU of Mlp.forward)
//   EPI_GELU_BWD  out = acc * gelu'(aux_in)                                      (dX of fc2 chained with GELU backward)
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace up3d {
namespace tc {

constexpr int BM = 128;      // rows per CTA tile = TMEM lanes
constexpr int BK = 64;       // K elements per ring stage = one 128-byte swizzle atom of bf16
constexpr int UMMA_K = 16;   // K per tcgen05.mma (32 bytes / sizeof(bf16))
// warp 0 TMA, warp 1 MMA, then EPI_WPQ epilogue warps per TMEM lane quarter.  The epilogue (TMEM -> bias / GELU -> bf16 ->
// staging -> TMA store) is the longest serial piece of a one-tile-per-SM launch: with two warps per quarter a thread
// walked 64 columns of GELU (3.7 us of an isolated 10.7 us fc1 + GELU launch, tools/bench_tc_instep.py); four per quarter
// give every 32-column chunk of a 96/128-wide tile its own warp.
#ifndef UP3D_TC_EPI_WPQ
#define UP3D_TC_EPI_WPQ 4
#endif
constexpr int EPI_WPQ = UP3D_TC_EPI_WPQ;
constexpr int EPI_THREADS = 128 * EPI_WPQ;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int A_STAGE_BYTES = BM * BK * 2;

enum { B_KMAJOR = 0, B_NMAJOR = 1 };
enum { EPI_NONE = 0, EPI_GELU = 1, EPI_GELU_BWD = 2 };

struct Args {
    int T, N, K;
    const __nv_bfloat16 *bias;     // (N) or NULL
    const __nv_bfloat16 *aux_in;   // (T, N) EPI_GELU_BWD: the GELU input saved by the forward
    __nv_bfloat16 *aux_out;        // (T, N) EPI_GELU: the GELU input
    void *out;                     // (T, N) bf16 or fp32
    int out_f32;
    int b_static;                  // B is not written by any kernel that may still be in flight (a weight)
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the device.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single-thread issue
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive once all tcgen05.mma issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread `lane` of the warp receives row (lane base + lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (tcgen05 "SmemDescriptor"): start address >> 4 in bits [0,14), leading byte offset >> 4
// in [16,30), stride byte offset >> 4 in [32,46), descriptor version 1 in [46,48), layout type in [61,64) (2 = 128-byte
// swizzle).  K-major tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart): SBO = 1024, LBO unused.
// N-major tile (64 N-elements = 128 B per k-row, 64 k-rows per 8 KB atom column): SBO = 1024 between 8-k-row groups,
// LBO = BK*128 between 64-wide N atoms.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (kind::f16): D fp32 (1 << 4), A and B bf16 (1 << 7, 1 << 10), transpose bits 15/16
// (1 = MN-major), N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of the value it feeds): one
// reciprocal, one exp and five FMAs instead of erff's ~40 instructions -- the GELU epilogues are issue-bound on the
// epilogue warps (ncu: 12.6 us with erff vs 7.9 us for the same tile without an activation).
__device__ __forceinline__ float erf_fast(float x, float e /* = exp(-x^2) */) {
    const float ax = fabsf(x);
    const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float r = 1.f - p * t * e;
    return copysignf(r, x);
}
__device__ __forceinline__ float gelu_f(float x) {
    const float h = x * 0.70710678118654752440f;
    return 0.5f * x * (1.f + erf_fast(h, __expf(-h * h)));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float h = x * 0.70710678118654752440f;
    const float e = __expf(-h * h);                    // = exp(-x^2 / 2): shared by the cdf and the pdf
    const float cdf = 0.5f * (1.f + erf_fast(h, e));
    return fmaf(x * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&p);
}
__device__ __forceinline__ void unpack8(const uint4 r, float *f) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}

template <int BN>
constexpr uint32_t tmem_cols() { return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512; }

// Shared-memory plan: [ring: NST x (A 16 KB | B BN*128 B)] [aux staging 128 x BN x 2 B] [barriers]; the output staging
// tile (128 x BN x 2|4 B) and the split-K partial tile reuse the ring, which is idle once the accumulator is complete.
// The staging tiles are cut into 32-column chunks of 128 rows; a chunk row is 64 B (bf16, 64-byte
// swizzle) or 128 B (fp32, 128-byte swizzle), which is the layout the TMA store / load boxes use.
template <int BN, int NST>
constexpr size_t ring_bytes() { return (size_t)NST * (A_STAGE_BYTES + BN * BK * 2); }
template <int BN, int NST>
constexpr size_t smem_bytes(bool aux) {
    return ring_bytes<BN, NST>() + (aux ? (size_t)BM * BN * 2 : 0) + 1024 + 256;
}
constexpr size_t SMEM_LIMIT = 232448;     // 227 KB opt-in maximum per CTA

// ---- epilogue staging helpers: chunk = 32 columns x 128 rows; `r` = row within the tile, `j` = 16-byte piece of the row
__device__ __forceinline__ uint32_t stage_off_bf16(int chunk, int r, int j) { return chunk * 8192 + r * 64 + 16 * (j ^ ((r >> 1) & 3)); }
__device__ __forceinline__ uint32_t stage_off_f32(int chunk, int r, int j) { return chunk * 16384 + r * 128 + 16 * (j ^ (r & 7)); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- the kernel
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes of CTA `rank`'s shared memory at the address that `local` has in this CTA (distributed shared memory)
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
    return v;
}

// KS > 1: split-K over a thread-block cluster (1,1,KS).  CTA z accumulates k-blocks [z*nk/KS, (z+1)*nk/KS) of the same
// output tile in its own TMEM; the partial tiles are parked in shared memory, and after one cluster barrier CTA z sums
// rows [z*128/KS, (z+1)*128/KS) of all KS partials through distributed shared memory (a reduce-scatter: every SM pulls
// only its slice), adds the bias and stores.  The deep-K Linears (fc2: K = 1536; the dX of qkv / fc1) thereby keep
// ~150 KB of operand bytes per SM in flight instead of ~500 KB on a quarter of the SMs.
template <int BN, int BMAJ, int EPI, int NST, int KS>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux, const Args args) {
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "tcgen05.mma with M = 128 needs N % 16 == 0, 16 <= N <= 256");
    static_assert(BMAJ == B_KMAJOR || BN % 64 == 0, "an N-major B tile is made of 64-wide swizzle atoms");
    constexpr int B_STAGE_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static_assert(STAGE_BYTES % 1024 == 0, "ring stages must keep the 1024-byte alignment of the swizzle atoms");
    static_assert(KS == 1 || EPI == EPI_NONE, "split-K is only wired for the plain epilogue");
    static_assert(KS == 1 || NST * STAGE_BYTES >= BM * BN * 4, "the partial tile is parked in the ring");
    constexpr bool HAS_AUX = EPI != EPI_NONE;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t *smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    static_assert(NST * STAGE_BYTES >= BM * BN * 4, "the output staging tile lives in the ring");
    uint8_t *out_stage = smem;                      // valid after accum_bar: every TMA load landed, every MMA retired
    uint8_t *aux_stage = smem + NST * STAGE_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(aux_stage + (HAS_AUX ? BM * BN * 2 : 0));
    uint64_t *empty_bar = full_bar + NST;
    uint64_t *accum_bar = empty_bar + NST;
    uint64_t *aux_bar = accum_bar + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(aux_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int nk_total = (args.K + BK - 1) / BK;
    const int z = KS > 1 ? (int)cluster_ctarank() : 0;
    const int kb0 = z * nk_total / KS;
    const int nk = (z + 1) * nk_total / KS - kb0;

    pdl_trigger();      // the next kernel of the stream may start its own prologue now
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        if (HAS_AUX) tma_prefetch_desc(&tmAux);
#pragma unroll
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_init(aux_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // The first ring pass needs no consumer hand-shake: issue it before the CTA-wide barrier.  B (a weight: not
        // written by any kernel in flight) is fetched even before the programmatic dependency on the previous kernel
        // resolves; A and the auxiliary tile wait for it.
        const int npre = nk < NST ? nk : NST;
        auto load_b = [&](int kb, int s) {
            uint8_t *sb = smem + s * STAGE_BYTES + A_STAGE_BYTES;
            if (BMAJ == B_KMAJOR) {
                tma_load_2d(sb, &tmB, &full_bar[s], (kb0 + kb) * BK, n0);
            } else {
#pragma unroll
                for (int a = 0; a < BN / 64; ++a) tma_load_2d(sb + a * (BK * 128), &tmB, &full_bar[s], n0 + 64 * a, (kb0 + kb) * BK);
            }
        };
        for (int kb = 0; kb < npre; ++kb) {
            mbar_expect_tx(&full_bar[kb], STAGE_BYTES);
            if (args.b_static) load_b(kb, kb);
        }
        pdl_wait();
        for (int kb = 0; kb < npre; ++kb) {
            if (!args.b_static) load_b(kb, kb);
            tma_load_2d(smem + kb * STAGE_BYTES, &tmA, &full_bar[kb], (kb0 + kb) * BK, m0);
        }
        if (EPI == EPI_GELU_BWD) {
            mbar_expect_tx(aux_bar, BM * BN * 2);
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) tma_load_2d(aux_stage + c * 8192, &tmAux, aux_bar, n0 + 32 * c, m0);
        }
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols<BN>());
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (one lane): the rest of the K loop
        if (lane == 0) {
            for (int kb = NST; kb < nk; ++kb) {
                const int s = kb % NST;
                mbar_wait(&empty_bar[s], ((kb / NST) & 1) ^ 1);
                uint8_t *sa = smem + s * STAGE_BYTES, *sb = sa + A_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_2d(sa, &tmA, &full_bar[s], (kb0 + kb) * BK, m0);
                if (BMAJ == B_KMAJOR) {
                    tma_load_2d(sb, &tmB, &full_bar[s], (kb0 + kb) * BK, n0);
                } else {
#pragma unroll
                    for (int a = 0; a < BN / 64; ++a) tma_load_2d(sb + a * (BK * 128), &tmB, &full_bar[s], n0 + 64 * a, (kb0 + kb) * BK);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one lane)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN, 0, BMAJ == B_NMAJOR);
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % NST;
                mbar_wait(&full_bar[s], (kb / NST) & 1);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adesc = make_desc(sa + k * (UMMA_K * 2), 0, 1024);
                    const uint64_t bdesc = BMAJ == B_KMAJOR ? make_desc(sb + k * (UMMA_K * 2), 0, 1024)
                                                            : make_desc(sb + k * (UMMA_K * 128), BK * 128, 1024);
                    umma_f16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0);
                }
                umma_commit(&empty_bar[s]);          // slot reusable once these MMAs have read it
            }
            umma_commit(accum_bar);                  // accumulator complete
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes 32*(w % 4) .. +31; thread = one output row
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;            // the EPI_WPQ warps of a lane quarter take the 32-column chunks in turn
        const int r = 32 * q + lane;
        if (EPI == EPI_GELU_BWD) mbar_wait(aux_bar, 0);
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (KS > 1) {
            // park this CTA's partial tile (fp32, swizzled 32-column chunks) in the now idle ring
#pragma unroll 1
            for (int c = half; c < BN / 32; c += EPI_WPQ) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * c), v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4 *>(smem + stage_off_f32(c, r, j)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        } else
#pragma unroll 1
        for (int c = half; c < BN / 32; c += EPI_WPQ) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * c), v);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (args.bias) {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float b[8];
                    unpack8(*reinterpret_cast<const uint4 *>(args.bias + n0 + 32 * c + 8 * j8), b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[8 * j8 + j] += b[j];
                }
            }
            if (EPI == EPI_GELU) {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float a = f[8 * j8 + 2 * j], b = f[8 * j8 + 2 * j + 1];
                        w[j] = pack_bf16(a, b);
                        f[8 * j8 + 2 * j] = gelu_f(bf16_round(a));
                        f[8 * j8 + 2 * j + 1] = gelu_f(bf16_round(b));
                    }
                    *reinterpret_cast<uint4 *>(aux_stage + stage_off_bf16(c, r, j8)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else if (EPI == EPI_GELU_BWD) {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float p[8];
                    unpack8(*reinterpret_cast<const uint4 *>(aux_stage + stage_off_bf16(c, r, j8)), p);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[8 * j8 + j] *= gelu_grad_f(p[j]);
                }
            }
            if (args.out_f32) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(out_stage + stage_off_f32(c, r, j)) =
                        make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4 *>(out_stage + stage_off_bf16(c, r, j)) =
                        make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                   pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
            }
            // this warp's 32 x 32 slab of the chunk -> global through TMA (rows past T are clipped by the tensor map)
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                const uint32_t slab = args.out_f32 ? (c * 16384 + q * 4096) : (c * 8192 + q * 2048);
                tma_store_2d(&tmOut, out_stage + slab, n0 + 32 * c, m0 + 32 * q);
                if (EPI == EPI_GELU) tma_store_2d(&tmAux, aux_stage + c * 8192 + q * 2048, n0 + 32 * c, m0 + 32 * q);
            }
        }
        if (KS == 1 && lane == 0) tma_store_commit_wait();      // shared memory must outlive the bulk reads
    }
    if (KS > 1) {
        __syncwarp();
        cluster_sync_all();                 // every CTA's partial tile is parked and visible cluster-wide
        if (warp >= 2) {
            constexpr int R = BM / KS, PIECES = BN / 4;      // rows of this CTA's slice; 16-byte pieces per row
            const int te = threadIdx.x - 64;
#pragma unroll 1
            for (int idx = te; idx < R * PIECES; idx += EPI_THREADS) {
                const int rl = idx / PIECES, p = idx % PIECES;
                const int row = z * R + rl;
                const uint32_t local = smem_u32(smem + stage_off_f32(p >> 3, row, p & 7));
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int src = 0; src < KS; ++src) {                 // fixed order: deterministic sums
                    const float4 v = ld_dsmem_f4(local, (uint32_t)src);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                const int col = n0 + 4 * p, grow = m0 + row;
                if (args.bias) {
                    const uint2 braw = *reinterpret_cast<const uint2 *>(args.bias + col);
                    const float2 b0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&braw.x));
                    const float2 b1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&braw.y));
                    acc.x += b0.x; acc.y += b0.y; acc.z += b1.x; acc.w += b1.y;
                }
                if (grow < args.T) {
                    const size_t o = (size_t)grow * args.N + col;
                    if (args.out_f32) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(args.out) + o) = acc;
                    else *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(args.out) + o) =
                             make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
                }
            }
        }
        __syncwarp();
        cluster_sync_all();                 // no CTA may exit while a peer still reads its shared memory
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols<BN>());
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

struct MapKey {
    const void *ptr; uint64_t rows, cols; uint32_t box_rows, box_cols, elem;
    bool operator==(const MapKey &o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows && box_cols == o.box_cols && elem == o.elem;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        size_t h = (size_t)k.ptr;
        h ^= k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
        return h ^ ((size_t)k.box_rows << 17) ^ ((size_t)k.box_cols << 29) ^ ((size_t)k.elem << 41);
    }
};

// Row-major (rows, cols) matrix of bf16 (elem 2) or fp32 (elem 4); box = (box_cols, box_rows); the swizzle width is the
// byte width of a box row (64 or 128); out-of-bounds reads give zeros, out-of-bounds writes are dropped.
// Encodings are cached per (pointer, shape): inside a captured step graph and with the caching allocator they recur.
static int tensor_map_2d(CUtensorMap *out, const void *ptr, uint64_t rows, uint64_t cols, uint32_t elem, uint32_t box_cols,
                         uint32_t box_rows) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, rows, cols, box_rows, box_cols, elem};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error("up3d_tc_linear: cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * elem};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estride[2] = {1, 1};
    const uint32_t row_bytes = box_cols * elem;
    if (row_bytes != 64 && row_bytes != 128) return set_error("up3d_tc_linear: internal: box rows must be 64 or 128 bytes");
    const CUresult r = fn(out, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                          const_cast<void *>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("up3d_tc_linear: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

template <int BN, int BMAJ, int EPI, int NST, int KS>
static int launch(const Args &a, const void *A, const void *B, cudaStream_t stream) {
    constexpr size_t smem = smem_bytes<BN, NST>(EPI != EPI_NONE);
    static_assert(smem <= SMEM_LIMIT, "tile configuration exceeds the shared memory of an SM");
    static bool attr_set = false;     // benign race: the attribute is idempotent
    if (!attr_set) {
        UP3D_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, BMAJ, EPI, NST, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    CUtensorMap tmA, tmB, tmOut, tmAux;
    if (tensor_map_2d(&tmA, A, a.T, a.K, 2, BK, BM)) return 1;
    if (BMAJ == B_KMAJOR) { if (tensor_map_2d(&tmB, B, a.N, a.K, 2, BK, BN)) return 1; }
    else                  { if (tensor_map_2d(&tmB, B, a.K, a.N, 2, 64, BK)) return 1; }
    if (tensor_map_2d(&tmOut, a.out, a.T, a.N, a.out_f32 ? 4 : 2, 32, 32)) return 1;
    tmAux = tmOut;
    if (EPI == EPI_GELU) { if (tensor_map_2d(&tmAux, a.aux_out, a.T, a.N, 2, 32, 32)) return 1; }
    if (EPI == EPI_GELU_BWD) { if (tensor_map_2d(&tmAux, a.aux_in, a.T, a.N, 2, 32, BM)) return 1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.N / BN, div_up(a.T, BM), KS);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (KS > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = KS;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<BN, BMAJ, EPI, NST, KS>, tmA, tmB, tmOut, tmAux, a);
    if (e != cudaSuccess) return set_error("launch of tc::gemm_kernel failed: %s", cudaGetErrorString(e));
    return 0;
}

// Ring depth: deep enough to have the whole K extent of the 384-wide layers in flight (6 stages) when the staging tiles
// leave room for it; 4 otherwise and for many-wave problems, where two co-resident CTAs per SM overlap one CTA's
// epilogue with the other's main loop.
template <int BN, int BMAJ, int EPI>
static int dispatch_stages(int ks, const Args &a, const void *A, const void *B, cudaStream_t stream) {
    const int nk = div_up(a.K, BK);
    constexpr bool aux = EPI != EPI_NONE;
    const long long ctas = (long long)(a.N / BN) * div_up(a.T, BM) * ks;
    if (EPI == EPI_NONE && ks > 1) {
        if (nk < ks) return set_error("up3d_tc_linear: split-K %d exceeds the %d K blocks", ks, nk);
        if (ks == 2) return launch<BN, BMAJ, EPI_NONE, 6, 2>(a, A, B, stream);
        if (ks == 4) return launch<BN, BMAJ, EPI_NONE, 6, 4>(a, A, B, stream);
        return set_error("up3d_tc_linear: split-K must be 1, 2 or 4");
    }
    if (ks != 1) return set_error("up3d_tc_linear: split-K needs the plain epilogue");
    // ring depth: as many K blocks in flight as the problem has (latency-bound single-wave problems), 4 for many-wave ones
    if (ctas <= 2 * UP3D_NUM_SMS) {
        if constexpr (smem_bytes<BN, 10>(aux) <= SMEM_LIMIT) { if (nk > 8) return launch<BN, BMAJ, EPI, 10, 1>(a, A, B, stream); }
        if constexpr (smem_bytes<BN, 8>(aux) <= SMEM_LIMIT) { if (nk > 6) return launch<BN, BMAJ, EPI, 8, 1>(a, A, B, stream); }
        if constexpr (smem_bytes<BN, 6>(aux) <= SMEM_LIMIT) { if (nk > 4) return launch<BN, BMAJ, EPI, 6, 1>(a, A, B, stream); }
    }
    return launch<BN, BMAJ, EPI, 4, 1>(a, A, B, stream);
}

template <int BMAJ, int EPI>
static int dispatch_bn(int bn, int ks, const Args &a, const void *A, const void *B, cudaStream_t stream) {
    switch (bn) {
        case 64: return dispatch_stages<64, BMAJ, EPI>(ks, a, A, B, stream);
        case 128: return dispatch_stages<128, BMAJ, EPI>(ks, a, A, B, stream);
        case 32: if (BMAJ == B_KMAJOR) return dispatch_stages<32, B_KMAJOR, EPI>(ks, a, A, B, stream); break;
        case 96: if (BMAJ == B_KMAJOR) return dispatch_stages<96, B_KMAJOR, EPI>(ks, a, A, B, stream); break;
        default: break;
    }
    return set_error("up3d_tc_linear: unsupported tile width %d", bn);
}

// Tile width: the widest BN that still yields about one wave of CTAs on the 148 SMs (these are 1032-row problems: the
// GEMMs are latency-, not throughput-bound, so more, narrower CTAs win until the A tile re-reads dominate).
// plus split-K (2 or 4 CTAs per tile) when K is deep.  Cost model (microseconds, fitted to the bench in
// tools/bench_tc_linear.py): a wave costs ~1 us of fixed latency plus its per-CTA operand bytes at ~100 KB/us of
// per-SM ingress; the split-K exchange adds ~0.8 us.
static void pick_tile(int T, int N, int K, int b_major, int epilogue, int *bn_out, int *ks_out) {
    const int mt = div_up(T, BM), nk = div_up(K, BK);
    const int cands_k[4] = {128, 96, 64, 32}, cands_n[2] = {128, 64};
    const int *c = b_major == B_KMAJOR ? cands_k : cands_n;
    const int nc = b_major == B_KMAJOR ? 4 : 2;
    double best = 1e30;
    *bn_out = 0; *ks_out = 1;
    for (int i = 0; i < nc; ++i) {
        if (N % c[i]) continue;
        for (int ks = 1; ks <= 4; ks *= 2) {
            if (ks > 1 && (epilogue != EPI_NONE || nk / ks < 3)) continue;
            const long long ctas = (long long)mt * (N / c[i]) * ks;
            const double waves = (double)((ctas + UP3D_NUM_SMS - 1) / UP3D_NUM_SMS);
            const double cta_kb = (BM + c[i]) * (double)(K / ks) * 2.0 / 1024.0;
            const double est = waves * (1.0 + cta_kb / 100.0) + (ks > 1 ? 0.8 : 0.0);
            if (est < best) { best = est; *bn_out = c[i]; *ks_out = ks; }
        }
    }
}

static bool g_pdl = [] {
    const char *e = getenv("UP3D_PDL");
    return !(e && e[0] == '0');
}();

}  // namespace tc

bool pdl_enabled() { return tc::g_pdl; }

}  // namespace up3d

extern "C" int up3d_set_pdl(int enabled) {
    up3d::tc::g_pdl = enabled != 0;
    return 0;
}

extern "C" int up3d_tc_linear(int T, int N, int K, const void *A, const void *B, int b_major, const void *bias, int epilogue,
                              const void *aux_in, void *aux_out, void *out, int out_f32, int tile_n, up3d_stream_t stream) {
    using namespace up3d;
    using namespace up3d::tc;
    UP3D_CHECK_ARG(T >= 0 && N > 0 && K > 0, "up3d_tc_linear: bad sizes T=%d N=%d K=%d", T, N, K);
    if (T == 0) return 0;
    UP3D_CHECK_ARG(A && B && out, "up3d_tc_linear: NULL operand");
    const int bm = b_major & 1, b_dynamic = (b_major >> 1) & 1;
    UP3D_CHECK_ARG((b_major & ~3) == 0, "up3d_tc_linear: b_major must be 0 (N,K) or 1 (K,N) [+2: B is produced by a kernel in flight]");
    UP3D_CHECK_ARG(K % 8 == 0 && N % 32 == 0, "up3d_tc_linear: K must be a multiple of 8 and N of 32 (got K=%d N=%d)", K, N);
    UP3D_CHECK_ARG((((uintptr_t)A | (uintptr_t)B | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)aux_in | (uintptr_t)aux_out) & 15) == 0,
                   "up3d_tc_linear: pointers must be 16-byte aligned");
    UP3D_CHECK_ARG(epilogue != EPI_GELU || (aux_out && !out_f32), "up3d_tc_linear: the GELU epilogue needs aux_out and a bf16 output");
    UP3D_CHECK_ARG(epilogue != EPI_GELU_BWD || aux_in, "up3d_tc_linear: the GELU-backward epilogue needs aux_in");
    int bn = tile_n & 0xFFFF, ks = (tile_n >> 16) & 0xFF;      // 0 = let the cost model choose
    if (bn == 0 || ks == 0) {
        int pbn, pks;
        pick_tile(T, N, K, bm, epilogue, &pbn, &pks);
        if (bn == 0) { bn = pbn; if (ks == 0) ks = pks; }
        if (ks == 0) ks = 1;
    }
    UP3D_CHECK_ARG(bn > 0 && N % bn == 0, "up3d_tc_linear: no tile width divides N=%d", N);
    Args a{T, N, K, (const __nv_bfloat16 *)bias, (const __nv_bfloat16 *)aux_in, (__nv_bfloat16 *)aux_out, out, out_f32, !b_dynamic};
    cudaStream_t st = (cudaStream_t)stream;
    if (bm == B_KMAJOR) {
        switch (epilogue) {
            case EPI_NONE: return dispatch_bn<B_KMAJOR, EPI_NONE>(bn, ks, a, A, B, st);
            case EPI_GELU: return dispatch_bn<B_KMAJOR, EPI_GELU>(bn, ks, a, A, B, st);
            case EPI_GELU_BWD: return dispatch_bn<B_KMAJOR, EPI_GELU_BWD>(bn, ks, a, A, B, st);
        }
    } else {
        switch (epilogue) {
            case EPI_NONE: return dispatch_bn<B_NMAJOR, EPI_NONE>(bn, ks, a, A, B, st);
            case EPI_GELU: return dispatch_bn<B_NMAJOR, EPI_GELU>(bn, ks, a, A, B, st);
            case EPI_GELU_BWD: return dispatch_bn<B_NMAJOR, EPI_GELU_BWD>(bn, ks, a, A, B, st);
        }
    }
    return set_error("up3d_tc_linear: unknown epilogue %d", epilogue);
}
