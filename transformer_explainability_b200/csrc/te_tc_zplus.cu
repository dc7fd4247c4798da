// tcgen05 path of the z+ Linear rule (modules/layers_ours.py:207-230, alpha=1) for sm_100a.
//
//   kernel 1 (MODE_S):   S    = safe_divide(R, x+ W+^T + x- W-^T)            [rows, out]
//   kernel 2 (MODE_R):   R_in = x+ * (S W+) + x- * (S W-)                     [rows, in]
//
// Both are "two-pass" 128x256 tiled GEMMs on the 5th-generation tensor cores:
//   * operands are K-major fp32 tiles of 128 B rows (32 floats) staged by TMA (SWIZZLE_128B) into a
//     4-stage shared-memory ring; tcgen05.mma.kind::tf32 (M=128, N=256, K=8) is issued by one thread,
//     accumulators live in TMEM (256 columns for kernel 1, 2 x 256 for kernel 2);
//   * pass 0 multiplies by W+ (pre-clamped, pre-rounded to TF32, K-major copy made once per frozen
//     weight by te_tc_prepare_weights), pass 1 by W-;
//   * kernel 1 clamps the activation tile IN PLACE in shared memory between the TMA arrival and the
//     MMA (max(.,0) in pass 0, min(.,0) in pass 1, round-to-nearest TF32) — an elementwise pass that
//     is independent of the swizzled layout — done by the four warps that later run the epilogue;
//   * epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused safe_divide / x+- recombination,
//     128-bit global stores.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = tile transform
// (kernel 1) and epilogue.  Pipelines: full[s] (TMA -> transform/MMA), xf[s] (transform -> MMA),
// empty[s] (tcgen05.commit -> TMA), accum (last commit -> epilogue).
//
// Numerics: TF32 (10-bit mantissa) operands, fp32 accumulation.  Z is a sum of non-negative products,
// so this is well conditioned; SURVEY.md §7b measured TF32 on exactly these GEMMs as indistinguishable
// from the fp32 reference's own noise.  Everything that feeds an ill-conditioned denominator stays on
// the fp32 SIMT path.
#include "te_tc_common.cuh"

namespace {

constexpr int STAGES = 4;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;             // 48 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int MODE_S = 1, MODE_R = 2, MODE_S1 = 3;   // MODE_S1: single-pass S kernel using the saved forward output

struct TcParams {
    int M, N, K;                 // C[M,N] = sum over two passes of A[M,K] * B_pass[N,K]^T
    const float* E; long long lde;   // MODE_S / MODE_S1: R [M,N] ; MODE_R: x [M,N]
    float* C; long long ldc;
    const float* Y; long long ldy; const float* bias;   // MODE_S1: forward output y = x W^T + bias
    int out_bf16;                                        // MODE_S / MODE_S1: write S as bf16 (C is then a bf16 [M, ldc] buffer)
};

// Ring depth / residency per kernel variant.  The single-pass S kernel has a heavy epilogue (reads R and y, safe_divide,
// writes S) and only needs 256 TMEM columns: with 2 stages of 48 KiB two CTAs share an SM and one CTA's prologue /
// epilogue overlaps the other's main loop.  The R kernel owns all 512 TMEM columns, so it stays alone with 4 stages.
template <int MODE> struct ZpCfg {
    static constexpr int STAGES = (MODE == MODE_S1) ? 2 : 4;
    static constexpr int MIN_CTAS = (MODE == MODE_S1) ? 2 : 1;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

// BF (MODE_R only): A (= S, written as bf16 by the S kernel) and B (bf16 weight copies) are 2-byte operands:
// one 128-byte swizzle row holds 64 elements, tcgen05.mma.kind::f16, half the shared-memory traffic per flop.
template <int MODE, bool BF = false>
__global__ void __launch_bounds__(NUM_THREADS, ZpCfg<MODE>::MIN_CTAS)
te_tc_zplus_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmB1, const TcParams p) {
    constexpr int KELEMS = BF ? 64 : 32;              // elements per k-block (one 128-byte row)
    constexpr int NST = ZpCfg<MODE>::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + NST * STAGE_BYTES;            // 8-byte barriers
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xf_bar = [&](int s) { return bars + 8u * (NST + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * NST + s); };
    const uint32_t accum_bar = bars + 8u * (3 * NST);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + NST * STAGE_BYTES + 8 * (3 * NST + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb = p.K / KELEMS, iters = (MODE == MODE_S1) ? kb : 2 * kb;
    constexpr uint32_t TMEM_COLS = (MODE == MODE_R) ? 512u : 256u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
        for (int s = 0; s < NST; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(xf_bar(s), XF_THREADS / 32);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), STAGE_BYTES);
                const int pass = (it >= kb) ? 1 : 0;
                const int k0 = (it - pass * kb) * KELEMS;
                const uint32_t sa = smem_base + s * STAGE_BYTES;
                tma_load_2d(sa, &tmA, full_bar(s), k0, m0);
                tma_load_2d(sa + A_BYTES, pass ? &tmB1 : &tmB0, full_bar(s), k0, n0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1u;
                mbar_wait(MODE != MODE_R ? xf_bar(s) : full_bar(s), ph);
                tcgen05_fence_after();
                const int pass = (it >= kb) ? 1 : 0;
                const uint32_t sa = smem_base + s * STAGE_BYTES;
                const uint64_t adesc = make_smem_desc(sa);
                const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
                const uint32_t d = tmem_base + ((MODE == MODE_R && pass) ? (uint32_t)BN : 0u);
                const bool first = (MODE != MODE_R) ? (it == 0) : (it == 0 || it == kb);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    // advance 8 tf32 (32 bytes) along K inside the 128-byte swizzle row: +2 in 16-byte units
                    if (BF) umma_bf16(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdescBf16, (first && k == 0) ? 0u : 1u);
                    else umma_tf32(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (first && k == 0) ? 0u : 1u);
                }
                umma_commit(empty_bar(s));          // frees the smem stage when these MMAs retire
            }
            umma_commit(accum_bar);                 // accumulators complete
        }
        __syncwarp();
    } else {
        // ================= tile transform (kernel 1) + epilogue: warps 2..5 =================
        const int et = threadIdx.x - 64;            // 0..127
        if (MODE != MODE_R) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1u;
                mbar_wait(full_bar(s), ph);
                const int pass = (it >= kb) ? 1 : 0;
                float4* a4 = reinterpret_cast<float4*>(smem_al + s * STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / XF_THREADS; ++i) {
                    float4 v = a4[et + i * XF_THREADS];
                    if (MODE == MODE_S1) { v.x = fabsf(v.x); v.y = fabsf(v.y); v.z = fabsf(v.z); v.w = fabsf(v.w); }
                    else if (pass == 0) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    else { v.x = fminf(v.x, 0.f); v.y = fminf(v.y, 0.f); v.z = fminf(v.z, 0.f); v.w = fminf(v.w, 0.f); }
                    v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
                    a4[et + i * XF_THREADS] = v;
                }
                fence_proxy_async();                // generic-proxy writes -> visible to the tensor-core (async) proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(xf_bar(s));   // one arrive per warp: 128 arrives on one mbarrier serialise
            }
        }
        mbar_wait(accum_bar, 0);
        tcgen05_fence_after();
        const int q = warp & 3;                     // TMEM lane quarter this warp may read
        const int row = m0 + q * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool live = row < p.M;
        const float* erow = p.E + (long long)row * p.lde + n0;
        float* crow = p.C + (long long)row * p.ldc + n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t acc[32];
            tmem_ld32(tlane + (uint32_t)(c * 32), acc);
            if (MODE != MODE_R) {
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 r = *reinterpret_cast<const float4*>(erow + c * 32 + j);
                        float z[4] = {__uint_as_float(acc[j + 0]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                      __uint_as_float(acc[j + 3])};
                        if (MODE == MODE_S1) {
                            // x+ W+^T + x- W-^T == ( x W^T + |x| |W|^T ) / 2 ,  x W^T = y - bias (saved forward output)
                            const float4 y = *reinterpret_cast<const float4*>(p.Y + (long long)row * p.ldy + n0 + c * 32 + j);
                            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + j));
                            // the true value is a sum of non-negative terms: a negative result is cancellation noise (the persistent
                            // kernel of te_tc_pair.cu additionally recomputes cancelled elements exactly)
                            z[0] = fmaxf(0.5f * ((y.x - bb.x) + z[0]), 0.f); z[1] = fmaxf(0.5f * ((y.y - bb.y) + z[1]), 0.f);
                            z[2] = fmaxf(0.5f * ((y.z - bb.z) + z[2]), 0.f); z[3] = fmaxf(0.5f * ((y.w - bb.w) + z[3]), 0.f);
                        }
                        if (p.out_bf16) {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(te_sd(r.x, z[0]), te_sd(r.y, z[1]));
                            __nv_bfloat162 hi = __floats2bfloat162_rn(te_sd(r.z, z[2]), te_sd(r.w, z[3]));
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&lo);
                            pk.y = *reinterpret_cast<uint32_t*>(&hi);
                            __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)row * p.ldc + n0 + c * 32 + j;
                            *reinterpret_cast<uint2*>(cb) = pk;
                        } else {
                            float4 o;
                            o.x = to_tf32(te_sd(r.x, z[0]));
                            o.y = to_tf32(te_sd(r.y, z[1]));
                            o.z = to_tf32(te_sd(r.z, z[2]));
                            o.w = to_tf32(te_sd(r.w, z[3]));
                            *reinterpret_cast<float4*>(crow + c * 32 + j) = o;
                        }
                    }
                }
            } else {
                uint32_t accn[32];
                tmem_ld32(tlane + (uint32_t)(BN + c * 32), accn);
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 x = *reinterpret_cast<const float4*>(erow + c * 32 + j);
                        float4 o;
                        o.x = fmaxf(x.x, 0.f) * __uint_as_float(acc[j + 0]) + fminf(x.x, 0.f) * __uint_as_float(accn[j + 0]);
                        o.y = fmaxf(x.y, 0.f) * __uint_as_float(acc[j + 1]) + fminf(x.y, 0.f) * __uint_as_float(accn[j + 1]);
                        o.z = fmaxf(x.z, 0.f) * __uint_as_float(acc[j + 2]) + fminf(x.z, 0.f) * __uint_as_float(accn[j + 2]);
                        o.w = fmaxf(x.w, 0.f) * __uint_as_float(acc[j + 3]) + fminf(x.w, 0.f) * __uint_as_float(accn[j + 3]);
                        *reinterpret_cast<float4*>(crow + c * 32 + j) = o;
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// CTA-pair version of the z+ kernels (tcgen05 cta_group::2): two CTAs of one cluster (adjacent 128-row tiles, same
// 256-column tile) execute ONE 256 x 256 x 8 MMA per k-step, issued by the leader CTA.  Each CTA stages its own
// 128 x 32 activation tile and only HALF of the weight tile (128 of the 256 rows), so a stage is 32 KiB instead of
// 48 KiB: 6 stages fit where 4 did and every byte brought into shared memory feeds 1.5x the flops — the kernels are
// bound by bytes in flight (L2 -> smem latency x ring size), not by the tensor pipe (ncu: 46 % / 36 % tensor active).
//   full[s]   local   TMA bytes of this CTA's A tile + B half
//   ready[s]  leader  S1 only: one arrive per transform warp of both CTAs (8) after |.| / TF32 rounding of its share
//                     of A — remote arrive through the cluster address of rank 0.  R has no transform: both CTAs'
//                     TMA bytes are counted directly on the leader's full[s] (cp.async.bulk.tensor .cta_group::2)
//   empty[s]  local   tcgen05.commit.cta_group::2 multicast from the leader to both CTAs
//   accum     local   same multicast commit after the last MMA; each CTA's epilogue reads its own 128 TMEM lanes
// =====================================================================================================================
constexpr int STAGES2 = 6;
constexpr int STAGE2_BYTES = A_BYTES + BH_BYTES;                  // 32 KiB
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
te_tc_zplus2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                    const __grid_constant__ CUtensorMap tmB1, const TcParams p) {
    static_assert(MODE == MODE_S1 || MODE == MODE_R, "pair kernel: single-pass S and R only");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES2 * STAGE2_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto ready_bar = [&](int s) { return bars + 8u * (STAGES2 + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES2 + s); };
    const uint32_t accum_bar = bars + 8u * (3 * STAGES2);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + STAGES2 * STAGE2_BYTES + 8 * (3 * STAGES2 + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    // 1-D grid of CTA pairs: the column tile runs fastest over the pairs, so the pairs that share an activation row
    // block are co-resident (L2 reuse of x); the two CTAs of a pair take adjacent 128-row tiles
    const int ntn = p.N / BN;
    const int pair = blockIdx.x >> 1;
    const int m0 = ((pair / ntn) * 2 + (int)rank) * BM, n0 = (pair % ntn) * BN;
    const int kb = p.K / BK, iters = (MODE == MODE_S1) ? kb : 2 * kb;
    constexpr uint32_t TMEM_COLS = (MODE == MODE_R) ? 512u : 256u;
    constexpr uint32_t READY_COUNT = 2u * (XF_THREADS / 32);        // S1: one arrive per transform warp of both CTAs

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), READY_COUNT);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();                                // barriers of both CTAs initialised, TMEM allocated
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own A tile + own half of the weight tile) =================
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES2;
                const uint32_t ph = (it / STAGES2) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                const int pass = (it >= kb) ? 1 : 0;
                const int k0 = (it - pass * kb) * BK;
                const uint32_t sa = smem_base + s * STAGE2_BYTES;
                if (MODE == MODE_S1) {
                    // the tile is clamped by this CTA's own warps first: bytes are counted on the local barrier
                    mbar_arrive_expect_tx(full_bar(s), STAGE2_BYTES);
                    tma_load_2d(sa, &tmA, full_bar(s), k0, m0);
                    tma_load_2d(sa + A_BYTES, pass ? &tmB1 : &tmB0, full_bar(s), k0, n0 + (int)rank * (BN / 2));
                } else {
                    // no transform: both CTAs' bytes are counted directly on the LEADER's barrier (cta_group::2 TMA)
                    if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * STAGE2_BYTES);
                    tma2_load_2d(sa, &tmA, full_bar(s), k0, m0);
                    tma2_load_2d(sa + A_BYTES, pass ? &tmB1 : &tmB0, full_bar(s), k0, n0 + (int)rank * (BN / 2));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES2;
                const uint32_t ph = (it / STAGES2) & 1u;
                if (MODE == MODE_S1) mbar_wait_cluster(ready_bar(s), ph);
                else mbar_wait(full_bar(s), ph);
                tcgen05_fence_after();
                const int pass = (it >= kb) ? 1 : 0;
                const uint32_t sa = smem_base + s * STAGE2_BYTES;
                const uint64_t adesc = make_smem_desc(sa);
                const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
                const uint32_t d = tmem_base + ((MODE == MODE_R && pass) ? (uint32_t)BN : 0u);
                const bool first = (MODE != MODE_R) ? (it == 0) : (it == 0 || it == kb);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k)
                    umma2_tf32(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc2, (first && k == 0) ? 0u : 1u);
                umma2_commit_both(empty_bar(s));     // frees this stage in BOTH CTAs when the MMAs retire
            }
            umma2_commit_both(accum_bar);
        }
        __syncwarp();
    } else {
        // ================= tile transform / relay + epilogue: warps 2..5 =================
        const int et = threadIdx.x - 64;            // 0..127
        if (MODE == MODE_S1) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES2;
                const uint32_t ph = (it / STAGES2) & 1u;
                mbar_wait(full_bar(s), ph);
                float4* a4 = reinterpret_cast<float4*>(smem_al + s * STAGE2_BYTES);
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / XF_THREADS; ++i) {
                    float4 v = a4[et + i * XF_THREADS];
                    v.x = to_tf32(fabsf(v.x)); v.y = to_tf32(fabsf(v.y)); v.z = to_tf32(fabsf(v.z)); v.w = to_tf32(fabsf(v.w));
                    a4[et + i * XF_THREADS] = v;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_rank0(ready_bar(s)));
            }
        }
        __syncwarp();
        mbar_wait(accum_bar, 0);
        tcgen05_fence_after();
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool live = row < p.M;
        const float* erow = p.E + (long long)row * p.lde + n0;
        float* crow = p.C + (long long)row * p.ldc + n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t acc[32];
            tmem_ld32(tlane + (uint32_t)(c * 32), acc);
            if (MODE == MODE_S1) {
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 r = *reinterpret_cast<const float4*>(erow + c * 32 + j);
                        const float4 y = *reinterpret_cast<const float4*>(p.Y + (long long)row * p.ldy + n0 + c * 32 + j);
                        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + j));
                        // x+ W+^T + x- W-^T == ( x W^T + |x| |W|^T ) / 2 ,  x W^T = y - bias (saved forward output)
                        const float z0 = fmaxf(0.5f * ((y.x - bb.x) + __uint_as_float(acc[j + 0])), 0.f);
                        const float z1 = fmaxf(0.5f * ((y.y - bb.y) + __uint_as_float(acc[j + 1])), 0.f);
                        const float z2 = fmaxf(0.5f * ((y.z - bb.z) + __uint_as_float(acc[j + 2])), 0.f);
                        const float z3 = fmaxf(0.5f * ((y.w - bb.w) + __uint_as_float(acc[j + 3])), 0.f);
                        float4 o;
                        o.x = to_tf32(te_sd(r.x, z0)); o.y = to_tf32(te_sd(r.y, z1));
                        o.z = to_tf32(te_sd(r.z, z2)); o.w = to_tf32(te_sd(r.w, z3));
                        *reinterpret_cast<float4*>(crow + c * 32 + j) = o;
                    }
                }
            } else {
                uint32_t accn[32];
                tmem_ld32(tlane + (uint32_t)(BN + c * 32), accn);
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 x = *reinterpret_cast<const float4*>(erow + c * 32 + j);
                        float4 o;
                        o.x = fmaxf(x.x, 0.f) * __uint_as_float(acc[j + 0]) + fminf(x.x, 0.f) * __uint_as_float(accn[j + 0]);
                        o.y = fmaxf(x.y, 0.f) * __uint_as_float(acc[j + 1]) + fminf(x.y, 0.f) * __uint_as_float(accn[j + 1]);
                        o.z = fmaxf(x.z, 0.f) * __uint_as_float(acc[j + 2]) + fminf(x.z, 0.f) * __uint_as_float(accn[j + 2]);
                        o.w = fmaxf(x.w, 0.f) * __uint_as_float(acc[j + 3]) + fminf(x.w, 0.f) * __uint_as_float(accn[j + 3]);
                        *reinterpret_cast<float4*>(crow + c * 32 + j) = o;
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();                                // nobody leaves while the peer may still touch this CTA
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- weight preparation: W [out,in] -> W+ , W- (K-major for kernel 1) and W+^T , W-^T (K-major for kernel 2),
//      all rounded to TF32 once (weights are frozen) -------------------------------------------------------
__global__ void prepare_weights_kernel(const float* __restrict__ w, float* __restrict__ d, int out_f, int in_f) {
    // d = [ W+ | W- | W+^T | W-^T | W_hi | W_lo | W^T_hi | W^T_lo | |W| ] (fp32, in*out floats each)
    //     [ bf16(W+^T) | bf16(W-^T) ]  (2-byte elements: in*out/2 floats each)
    //     [ bf16(W_hi) | bf16(W_lo) ]  (K-major [out,in], 2-byte elements): the correction operands of the mixed-kind forward
    const long long n = (long long)out_f * in_f;
    float *wp = d, *wn = d + n, *wpt = d + 2 * n, *wnt = d + 3 * n, *wh = d + 4 * n, *wl = d + 5 * n, *wth = d + 6 * n,
          *wtl = d + 7 * n, *wa = d + 8 * n;
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;      // bx: in index, by: out index
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int o = by + i, c = bx + threadIdx.x;
        float v = 0.f;
        if (o < out_f && c < in_f) {
            const long long idx = (long long)o * in_f + c;
            v = w[idx];
            wp[idx] = to_tf32(fmaxf(v, 0.f));
            wn[idx] = to_tf32(fminf(v, 0.f));
            const float hi = to_tf32(v);
            wh[idx] = hi;
            wl[idx] = to_tf32(v - hi);
            wa[idx] = to_tf32(fabsf(v));
            __nv_bfloat16* mb = reinterpret_cast<__nv_bfloat16*>(d + 10 * n);        // [ bf16(W_hi) | bf16(W_lo) ]: mixed-kind forward
            mb[idx] = __float2bfloat16_rn(hi);
            mb[n + idx] = __float2bfloat16_rn(v - hi);
            reinterpret_cast<__nv_bfloat16*>(d + 11 * n)[idx] = __float2bfloat16_rn(fabsf(v));      // bf16(|W|): bf16 S1 kernel
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = bx + i, o = by + threadIdx.x;
        if (o < out_f && c < in_f) {
            const float v = tile[threadIdx.x][i];
            const long long idx = (long long)c * out_f + o;
            wpt[idx] = to_tf32(fmaxf(v, 0.f));
            wnt[idx] = to_tf32(fminf(v, 0.f));
            __nv_bfloat16* bp = reinterpret_cast<__nv_bfloat16*>(d + 9 * n);
            bp[idx] = __float2bfloat16_rn(fmaxf(v, 0.f));
            bp[n + idx] = __float2bfloat16_rn(fminf(v, 0.f));
            const float hi = to_tf32(v);
            wth[idx] = hi;
            wtl[idx] = to_tf32(v - hi);
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------------------
template <int MODE>
int launch(const float* A, long long lda, const float* B0, const float* B1, const float* E, long long lde, float* C,
           long long ldc, long long M, int N, int K, cudaStream_t st, const float* Y = nullptr, long long ldy = 0,
           const float* bias = nullptr, int out_bf16 = 0) {
    CUtensorMap tmA, tmB0, tmB1;
    if (!make_map(&tmA, A, M, K, lda, BM) || !make_map(&tmB0, B0, N, K, K, BN) || !make_map(&tmB1, B1, N, K, K, BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;          // per-device attribute: one bit per device
    if (!smem_optin(te_tc_zplus_kernel<MODE>, ZpCfg<MODE>::SMEM, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    TcParams p;
    p.M = (int)M; p.N = N; p.K = K; p.E = E; p.lde = lde; p.C = C; p.ldc = ldc; p.Y = Y; p.ldy = ldy; p.bias = bias;
    p.out_bf16 = out_bf16;
    dim3 grid(N / BN, (unsigned)((M + BM - 1) / BM));
    te_tc_zplus_kernel<MODE><<<grid, NUM_THREADS, ZpCfg<MODE>::SMEM, st>>>(tmA, tmB0, tmB1, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

// CTA-pair (cta_group::2) launch of the single-pass S kernel / the R kernel.  Opt-in (TE_B200_ZPLUS_2CTA=1, or
// te_tc_set_pair_kernels): parity-tested, but as NON-persistent kernels they measured slower than the single-CTA
// kernels (fc2-shaped rule 2.26 ms vs 1.89 ms; whole step 936 vs 968 expl/s) — a pair can only start when both SMs of
// a TPC are free and pays two cluster barriers per tile.  They are the base for a persistent version.
int g_pair_kernels = -1;
bool use_pair_kernels() {
    if (g_pair_kernels < 0) {
        const char* e = getenv("TE_B200_ZPLUS_2CTA");
        g_pair_kernels = (e && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
    }
    return g_pair_kernels >= 1;
}
bool use_pair_s_kernel() { return use_pair_kernels() && g_pair_kernels == 1; }   // 2: pair form for the R kernel only

template <int MODE>
int launch2(const float* A, long long lda, const float* B0, const float* B1, const float* E, long long lde, float* C,
            long long ldc, long long M, int N, int K, cudaStream_t st, const float* Y = nullptr, long long ldy = 0,
            const float* bias = nullptr) {
    CUtensorMap tmA, tmB0, tmB1;
    if (!make_map(&tmA, A, M, K, lda, BM) || !make_map(&tmB0, B0, N, K, K, BN / 2) || !make_map(&tmB1, B1, N, K, K, BN / 2)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;          // per-device attribute: one bit per device
    if (!smem_optin(te_tc_zplus2_kernel<MODE>, SMEM2_BYTES, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    TcParams p;
    p.M = (int)M; p.N = N; p.K = K; p.E = E; p.lde = lde; p.C = C; p.ldc = ldc; p.Y = Y; p.ldy = ldy; p.bias = bias;
    p.out_bf16 = 0;
    const unsigned mtiles = (unsigned)((M + BM - 1) / BM);
    dim3 grid((unsigned)(N / BN) * ((mtiles + 1u) & ~1u));   // whole CTA pairs: an odd last tile gets an all-padding partner
    te_tc_zplus2_kernel<MODE><<<grid, NUM_THREADS, SMEM2_BYTES, st>>>(tmA, tmB0, tmB1, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

// R kernel with bf16 operands: A = S (bf16 [M, K], row stride K), B0/B1 = bf16 [N, K]
int launch_r_bf16(const void* A, const void* B0, const void* B1, const float* E, long long lde, float* C, long long ldc,
                  long long M, int N, int K, cudaStream_t st) {
    CUtensorMap tmA, tmB0, tmB1;
    if (!make_map_t(&tmA, A, M, K, K, BM, true) || !make_map_t(&tmB0, B0, N, K, K, BN, true) ||
        !make_map_t(&tmB1, B1, N, K, K, BN, true)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (bf16)");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;          // per-device attribute: one bit per device
    if (!smem_optin(te_tc_zplus_kernel<MODE_R, true>, SMEM_BYTES, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.M = (int)M; p.N = N; p.K = K; p.E = E; p.lde = lde; p.C = C; p.ldc = ldc;
    dim3 grid(N / BN, (unsigned)((M + BM - 1) / BM));
    te_tc_zplus_kernel<MODE_R, true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB0, tmB1, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

}  // namespace

bool te_tc_zplus_supported(long long rows, int in_features, int out_features, long long ldx) {
    return rows > 0 && rows < (1LL << 31) && in_features % BN == 0 && out_features % BN == 0 && ldx % 4 == 0 &&
           get_encode() != nullptr;
}

void te_tc_set_pair_kernels(int on) { g_pair_kernels = (on == 2) ? 2 : (on ? 1 : 0); }

static int g_zplus_persistent = -1;            // persistent CTA-pair kernels of te_tc_pair.cu (default on; TE_B200_ZPLUS_PERSISTENT=0)
static bool use_persistent() {
    if (g_zplus_persistent < 0) {
        const char* e = getenv("TE_B200_ZPLUS_PERSISTENT");
        g_zplus_persistent = (e && e[0] == '0') ? 0 : 1;
    }
    return g_zplus_persistent == 1;
}
void te_tc_set_zplus_persistent(int on) { g_zplus_persistent = on ? 1 : 0; }

long long te_tc_derived_floats(int in_features, int out_features) { return 16LL * in_features * out_features; }

int te_tc_prepare_weights(const float* w, float* derived, int in_features, int out_features, cudaStream_t st) {
    dim3 grid((in_features + 31) / 32, (out_features + 31) / 32), block(32, 8);
    prepare_weights_kernel<<<grid, block, 0, st>>>(w, derived, out_features, in_features);
    TE_CUDA_CHECK_LAUNCH();
    const long long n = (long long)in_features * out_features;
    if (in_features % 8 == 0 && in_features >= 8 && a16(w))        // row-scaled fp16 split: [hi | lo | 2^-f] from 11.5 n
        TE_TRY(te_tc_rowsplit_f16(w, in_features, out_features, in_features, derived + 11 * n + n / 2, derived + 12 * n,
                                  derived + 12 * n + n / 2, st));
    if (out_features % 8 == 0 && in_features >= 2) {               // single-pass fp16 operands from the TF32-rounded transposes [in, out]
        TE_TRY(te_tc_rowsplit_f16(derived + 6 * n, out_features, in_features, out_features, derived + 13 * n, nullptr,
                                  derived + 13 * n + n / 2, st));
        TE_TRY(te_tc_rowsplit_f16(derived + 2 * n, out_features, in_features, out_features, derived + 14 * n, nullptr,
                                  derived + 15 * n, st));
        TE_TRY(te_tc_rowsplit_f16(derived + 3 * n, out_features, in_features, out_features, derived + 14 * n + n / 2, nullptr,
                                  derived + 15 * n + in_features, st));
    }
    return TE_OK;
}

int te_tc_zplus_linear_relprop(const float* x, long long ldx, const float* derived, const float* r, long long ldr,
                               float* out, float* s_scratch, long long rows, int in_features, int out_features, cudaStream_t st,
                               const float* y, long long ldy, const float* bias, int bf16, long long ld_out, float* xabs) {
    if (ld_out == 0) ld_out = in_features;
    if (!a16(x) || !a16(derived) || !a16(r) || !a16(out) || !a16(s_scratch)) {
        te_set_last_error("te_gemm_tc: operands must be 16-byte aligned");
        return TE_ERR_ARG;
    }
    const long long n = (long long)in_features * out_features;
    const float *wp = derived, *wn = derived + n, *wpt = derived + 2 * n, *wnt = derived + 3 * n;
    // S = sd(R, x+ W+^T + x- W-^T)          A = x [rows, in] ; B = W+/- [out, in]
    const bool rb = (bf16 & 1) && (out_features % 64 == 0);          // S as bf16, R kernel with bf16 operands (kind::f16)
    if (!rb && use_persistent() && xabs && a16(xabs) && y && a16(y) && ldy % 4 == 0 && (!bias || a16(bias)) &&
        te_tc_pair_supported(rows, in_features, out_features, ldx) && te_tc_pair_supported(rows, out_features, in_features, out_features)) {
        // persistent CTA-pair kernels (te_tc_pair.cu): single-pass S, then R with the A operand shared by both products
        if ((bf16 & 4) && te_tc_f16_single_supported(rows, out_features, in_features, out_features)) {
            // second contraction on kind::f16 (te_tc_fwd16.cu, FM_R).  Its A operand — S as hi-only block-scaled fp16 — is written
            // by the S kernel's epilogue straight into s_scratch ([rows, out] fp16, then the [rows, out/128] scales): rows*out
            // floats hold both (out/2 + out/128 <= out).
            float* s16_scale = s_scratch + ((rows * out_features / 2 + 63) & ~63LL);
            TE_TRY(te_tc_pair_zplus_s1(x, ldx, xabs, derived, r, ldr, y, ldy, bias, nullptr, rows, in_features, out_features, st,
                                       (bf16 & 2) != 0, s_scratch, s16_scale));
            return te_tc_zplus_r16(nullptr, s_scratch, s16_scale, derived, x, ldx, out, ld_out, rows, in_features, out_features, st);
        }
        TE_TRY(te_tc_pair_zplus_s1(x, ldx, xabs, derived, r, ldr, y, ldy, bias, s_scratch, rows, in_features, out_features, st,
                                   (bf16 & 2) != 0));
        return te_tc_pair_zplus_r(s_scratch, derived, x, ldx, out, ld_out, rows, in_features, out_features, st);
    }
    if (ld_out != in_features) {
        te_set_last_error("te_gemm_tc: a strided output needs the persistent pair kernels");
        return TE_ERR_UNSUPPORTED;
    }
    if (y && a16(y) && ldy % 4 == 0 && (!bias || a16(bias))) {
        // single pass: Z = ((y - bias) + |x| |W|^T) / 2 with the saved forward output y = x W^T + bias
        const float* wabs = derived + 8 * n;
        if (!rb && use_pair_s_kernel())
            TE_TRY(launch2<MODE_S1>(x, ldx, wabs, wabs, r, ldr, s_scratch, out_features, rows, out_features, in_features, st, y,
                                    ldy, bias));
        else
            TE_TRY(launch<MODE_S1>(x, ldx, wabs, wabs, r, ldr, s_scratch, out_features, rows, out_features, in_features, st, y,
                                   ldy, bias, rb ? 1 : 0));
    } else {
        TE_TRY(launch<MODE_S>(x, ldx, wp, wn, r, ldr, s_scratch, out_features, rows, out_features, in_features, st, nullptr, 0,
                              nullptr, rb ? 1 : 0));
    }
    if (rb) {
        const __nv_bfloat16* wb = reinterpret_cast<const __nv_bfloat16*>(derived + 9 * n);
        return launch_r_bf16(s_scratch, wb, wb + n, x, ldx, out, in_features, rows, in_features, out_features, st);
    }
    // R_in = x+ (S W+) + x- (S W-)          A = S [rows, out] ; B = W+/-^T [in, out]
    if (use_pair_kernels())
        return launch2<MODE_R>(s_scratch, out_features, wpt, wnt, x, ldx, out, in_features, rows, in_features, out_features, st);
    TE_TRY(launch<MODE_R>(s_scratch, out_features, wpt, wnt, x, ldx, out, in_features, rows, in_features, out_features, st));
    return TE_OK;
}
