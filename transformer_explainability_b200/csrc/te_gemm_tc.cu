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
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "te_gemm_tc.h"

namespace {

constexpr int BM = 128, BN = 256, BK = 32;                 // BK floats = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;                       // 16 KiB
constexpr int B_BYTES = BN * BK * 4;                       // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;             // 48 KiB
constexpr int NUM_THREADS = 192;
constexpr int XF_THREADS = 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int MODE_S = 1, MODE_R = 2, MODE_S1 = 3;   // MODE_S1: single-pass S kernel using the saved forward output

struct TcParams {
    int M, N, K;                 // C[M,N] = sum over two passes of A[M,K] * B_pass[N,K]^T
    const float* E; long long lde;   // MODE_S / MODE_S1: R [M,N] ; MODE_R: x [M,N]
    float* C; long long ldc;
    const float* Y; long long ldy; const float* bias;   // MODE_S1: forward output y = x W^T + bias
    int out_bf16;                                        // MODE_S / MODE_S1: write S as bf16 (C is then a bf16 [M, ldc] buffer)
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//  [0,14) start address >> 4 ; [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1) ;
//  [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B) ; [46,48) version = 1 ; [61,64) layout = 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format TF32 [7,10)/[10,13)=2,
// a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
// same with BF16 operands (a/b format = 1), kind::f16: K = 16 elements (32 bytes) per MMA, 64 elements per 128-byte row
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

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
                            z[0] = 0.5f * ((y.x - bb.x) + z[0]); z[1] = 0.5f * ((y.y - bb.y) + z[1]);
                            z[2] = 0.5f * ((y.z - bb.z) + z[2]); z[3] = 0.5f * ((y.w - bb.w) + z[3]);
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
constexpr int BH_BYTES = B_BYTES / 2;                             // 16 KiB: this CTA's half of the weight tile
constexpr int STAGE2_BYTES = A_BYTES + BH_BYTES;                  // 32 KiB
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;
constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank0(uint32_t addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAITC_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAITC_DONE;\n\t"
        "bra WAITC_LOOP;\n\t"
        "WAITC_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// cta_group::2 TMA load: the bytes are counted on the barrier at the same offset in the LEADER CTA (peer bit cleared)
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

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
                        const float z0 = 0.5f * ((y.x - bb.x) + __uint_as_float(acc[j + 0]));
                        const float z1 = 0.5f * ((y.y - bb.y) + __uint_as_float(acc[j + 1]));
                        const float z2 = 0.5f * ((y.z - bb.z) + __uint_as_float(acc[j + 2]));
                        const float z3 = 0.5f * ((y.w - bb.w) + __uint_as_float(acc[j + 3]));
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

// =====================================================================================================================
// fp32-grade Linear GEMM on tensor cores: 3xTF32 error-compensated split
//   C = A B^T  ~=  A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T ,  x_hi = tf32(x), x_lo = tf32(x - x_hi)
// A (activations) is split in shared memory by the transform warps; B (frozen weights) is pre-split.
// Tile 128 x 256 x 32, 2 stages of [A_hi 16K | A_lo 16K | B_hi 32K | B_lo 32K], 12 MMAs per k-block.
// =====================================================================================================================
constexpr int STAGES3 = 2;
constexpr int STAGE3_BYTES = 2 * A_BYTES + 2 * B_BYTES;          // 96 KiB
constexpr int SMEM3_BYTES = STAGES3 * STAGE3_BYTES + 1024 + 256;
constexpr int NUM_THREADS3 = 320;                                // TMA, MMA, 4 transform+drain warps, 4 drain warps
constexpr int DRAIN_THREADS = 256;
constexpr int CHUNK = 4;                                         // k-blocks (of 32) accumulated inside the tensor core
enum { EP_STORE = 0, EP_BIAS = 1, EP_BIAS_GELU = 2, EP_BIAS_ADD = 3, EP_GELU_BWD = 4 };

struct Tc3Params {
    int M, N, K;
    const float* bias; const float* E; long long lde;
    float* C; long long ldc; float* C2; long long ldc2;
};

// epilogue of both 3xTF32 Linear kernels: one output row x 128 columns per thread, from the fp32 register sums
template <int EPI>
__device__ __forceinline__ void gemm3x_epilogue(const Tc3Params& p, const float (&sum)[128], int row, int cbase) {
        if (row < p.M) {
            const float* erow = p.E ? p.E + (long long)row * p.lde + cbase : nullptr;
            float* crow = p.C + (long long)row * p.ldc + cbase;
            float* c2row = p.C2 ? p.C2 + (long long)row * p.ldc2 + cbase : nullptr;
#pragma unroll
            for (int j = 0; j < 128; j += 4) {
                float o[4], o2[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
                if (EPI == EP_BIAS || EPI == EP_BIAS_GELU || EPI == EP_BIAS_ADD) {
                    if (p.bias) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
                        bb[0] = t.x; bb[1] = t.y; bb[2] = t.z; bb[3] = t.w;
                    }
                }
                if (EPI == EP_BIAS_ADD || EPI == EP_GELU_BWD) {
                    const float4 t = *reinterpret_cast<const float4*>(erow + j);
                    e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float a = sum[j + u];
                    if (EPI == EP_STORE) o[u] = a;
                    else if (EPI == EP_BIAS) o[u] = a + bb[u];
                    else if (EPI == EP_BIAS_GELU) { o[u] = a + bb[u]; o2[u] = te_gelu(o[u]); }
                    else if (EPI == EP_BIAS_ADD) { o[u] = a + bb[u]; o2[u] = e[u] + o[u]; }
                    else o[u] = a * te_gelu_grad(e[u]);
                }
                *reinterpret_cast<float4*>(crow + j) = make_float4(o[0], o[1], o[2], o[3]);
                if (EPI == EP_BIAS_GELU || EPI == EP_BIAS_ADD)
                    *reinterpret_cast<float4*>(c2row + j) = make_float4(o2[0], o2[1], o2[2], o2[3]);
            }
        }
}

// The tensor core accumulates in fp32 with truncation (round-toward-zero) at every MMA, so a long reduction drifts
// by ~7e-9*K relative (measured: 2e-5 at K = 3072).  To stay at fp32 grade the reduction is cut into chunks of
// CHUNK*32 = 128 elements: each chunk accumulates in one of two TMEM accumulators (2 x 256 columns), and while the
// MMAs of the next chunk run, 8 warps drain the finished accumulator with tcgen05.ld and add it into fp32 register
// sums with round-to-nearest CUDA-core adds (128 sums per thread: one row x half of the 256 columns).
template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS3, 1)      // 10 warps are register-allocated as 12: 168 regs / thread
te_tc_gemm3x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                    const __grid_constant__ CUtensorMap tmBl, const Tc3Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES3 * STAGE3_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xf_bar = [&](int s) { return bars + 8u * (2 + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (4 + s); };
    auto accfull_bar = [&](int b) { return bars + 8u * (6 + b); };
    auto accfree_bar = [&](int b) { return bars + 8u * (8 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + STAGES3 * STAGE3_BYTES + 8 * 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb = p.K / BK;
    const int nchunks = (kb + CHUNK - 1) / CHUNK;
    constexpr uint32_t TMEM_COLS = 512u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
        for (int s = 0; s < STAGES3; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(xf_bar(s), XF_THREADS / 32);
            mbar_init(empty_bar(s), 1);
            mbar_init(accfull_bar(s), 1);
            mbar_init(accfree_bar(s), DRAIN_THREADS);
        }
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
        if (lane == 0) {
            for (int it = 0; it < kb; ++it) {
                const int s = it % STAGES3;
                const uint32_t ph = (it / STAGES3) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
                const uint32_t sa = smem_base + s * STAGE3_BYTES;
                tma_load_2d(sa, &tmA, full_bar(s), it * BK, m0);                         // raw A -> A_hi slot
                tma_load_2d(sa + 2 * A_BYTES, &tmBh, full_bar(s), it * BK, n0);
                tma_load_2d(sa + 2 * A_BYTES + B_BYTES, &tmBl, full_bar(s), it * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < kb; ++it) {
                const int c = it / CHUNK, b = c & 1;
                const bool chunk_start = (it % CHUNK) == 0;
                if (chunk_start && c >= 2) {                        // accumulator b must have been drained (chunk c-2)
                    mbar_wait(accfree_bar(b), (uint32_t)(((c >> 1) & 1) ^ 1));
                    tcgen05_fence_after();
                }
                const int s = it % STAGES3;
                const uint32_t ph = (it / STAGES3) & 1u;
                mbar_wait(xf_bar(s), ph);
                tcgen05_fence_after();
                const uint32_t sa = smem_base + s * STAGE3_BYTES;
                const uint64_t ah = make_smem_desc(sa), al = make_smem_desc(sa + A_BYTES);
                const uint64_t bh = make_smem_desc(sa + 2 * A_BYTES), bl = make_smem_desc(sa + 2 * A_BYTES + B_BYTES);
                const uint32_t d = tmem_base + (uint32_t)(b * BN);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint64_t o = (uint64_t)(2 * k);
                    umma_tf32(d, al + o, bh + o, kIdesc, (chunk_start && k == 0) ? 0u : 1u);        // small terms first
                    umma_tf32(d, ah + o, bl + o, kIdesc, 1u);
                    umma_tf32(d, ah + o, bh + o, kIdesc, 1u);
                }
                umma_commit(empty_bar(s));
                if ((it % CHUNK) == CHUNK - 1 || it == kb - 1) umma_commit(accfull_bar(b));
            }
        }
        __syncwarp();
    } else {
        // ---- warps 2..9: row = lane quarter (warp & 3), column half = 0 for warps 2-5, 1 for warps 6-9 ----
        const int q = warp & 3;
        const int half = (warp >= 6) ? 1 : 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
        float sum[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) sum[j] = 0.f;

        auto drain = [&](int c) {
            const int b = c & 1;
            mbar_wait(accfull_bar(b), (uint32_t)((c >> 1) & 1));
            tcgen05_fence_after();
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                uint32_t v[16];
                tmem_ld16(tlane + (uint32_t)(b * BN + cc * 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[cc * 16 + j] += __uint_as_float(v[j]);
            }
            tcgen05_fence_before();
            mbar_arrive(accfree_bar(b));
        };

        if (warp < 6) {
            const int et = threadIdx.x - 64;                        // 0..127: the four transform warps
            for (int it = 0; it < kb; ++it) {
                const int s = it % STAGES3;
                const uint32_t ph = (it / STAGES3) & 1u;
                mbar_wait(full_bar(s), ph);
                float4* a4 = reinterpret_cast<float4*>(smem_al + s * STAGE3_BYTES);
                float4* l4 = reinterpret_cast<float4*>(smem_al + s * STAGE3_BYTES + A_BYTES);
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / XF_THREADS; ++i) {
                    const float4 v = a4[et + i * XF_THREADS];
                    float4 h, l;
                    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
                    l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
                    a4[et + i * XF_THREADS] = h;
                    l4[et + i * XF_THREADS] = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(xf_bar(s));          // one arrive per transform warp
                // the last stage of chunk c has just been handed to the MMA warp: drain chunk c-1 meanwhile
                if (((it % CHUNK) == CHUNK - 1 || it == kb - 1) && it / CHUNK >= 1) drain(it / CHUNK - 1);
            }
            drain(nchunks - 1);
        } else {
            for (int c = 0; c < nchunks; ++c) drain(c);
        }

        // ---- epilogue from the register sums ----
        gemm3x_epilogue<EPI>(p, sum, m0 + q * 32 + lane, n0 + half * 128);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// CTA-pair version of the 3xTF32 Linear GEMM (tcgen05 cta_group::2): the two CTAs of a cluster own adjacent 128-row
// tiles of the same 256-column tile and execute ONE 256 x 256 x 8 MMA per issue (leader CTA).  Each CTA stages its own
// activation tile (raw -> hi, lo) and only its HALF of the pre-split weight tile (128 of the 256 rows of W_hi and
// W_lo): a stage is 64 KiB instead of 96 KiB, so the ring is 3 deep instead of 2, and per k-block an SM moves
// 48 + 48 + 96 KiB through shared memory (TMA in, split, MMA operand reads) instead of 80 + 48 + 144 KiB.
//   full[s]     local    TMA bytes of this CTA's A tile and weight halves
//   ready[s]    leader   one arrive per transform warp of BOTH CTAs (8) after the hi/lo split (remote arrive)
//   empty[s]    local    tcgen05.commit.cta_group::2 multicast
//   accfull[b]  local    same multicast commit at the end of a 128-element chunk
//   accfree[b]  leader   one arrive per drain warp of BOTH CTAs (16) once accumulator b has been read out
// =====================================================================================================================
constexpr int STAGES3P = 3;
constexpr int STAGE3P_BYTES = 2 * A_BYTES + 2 * BH_BYTES;        // 64 KiB
constexpr int SMEM3P_BYTES = STAGES3P * STAGE3P_BYTES + 1024 + 256;

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS3, 1)
te_tc_gemm3x2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                     const __grid_constant__ CUtensorMap tmBl, const Tc3Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + STAGES3P * STAGE3P_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto ready_bar = [&](int s) { return bars + 8u * (STAGES3P + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES3P + s); };
    auto accfull_bar = [&](int b) { return bars + 8u * (3 * STAGES3P + b); };
    auto accfree_bar = [&](int b) { return bars + 8u * (3 * STAGES3P + 2 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + STAGES3P * STAGE3P_BYTES + 8 * (3 * STAGES3P + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int ntn = p.N / BN;
    const int pair = blockIdx.x >> 1;                 // column tile fastest over the pairs (L2 reuse of the activations)
    const int m0 = ((pair / ntn) * 2 + (int)rank) * BM, n0 = (pair % ntn) * BN;
    const int kb = p.K / BK;
    const int nchunks = (kb + CHUNK - 1) / CHUNK;
    constexpr uint32_t TMEM_COLS = 512u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
        for (int s = 0; s < STAGES3P; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), 2 * (XF_THREADS / 32));
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accfull_bar(b), 1);
            mbar_init(accfree_bar(b), 2 * (DRAIN_THREADS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < kb; ++it) {
                const int s = it % STAGES3P;
                const uint32_t ph = (it / STAGES3P) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), A_BYTES + 2 * BH_BYTES);
                const uint32_t sa = smem_base + s * STAGE3P_BYTES;
                tma_load_2d(sa, &tmA, full_bar(s), it * BK, m0);                                    // raw A -> A_hi slot
                tma_load_2d(sa + 2 * A_BYTES, &tmBh, full_bar(s), it * BK, n0 + (int)rank * (BN / 2));
                tma_load_2d(sa + 2 * A_BYTES + BH_BYTES, &tmBl, full_bar(s), it * BK, n0 + (int)rank * (BN / 2));
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            for (int it = 0; it < kb; ++it) {
                const int c = it / CHUNK, b = c & 1;
                const bool chunk_start = (it % CHUNK) == 0;
                if (chunk_start && c >= 2) {                        // accumulator b drained (chunk c-2) in BOTH CTAs
                    mbar_wait_cluster(accfree_bar(b), (uint32_t)(((c >> 1) & 1) ^ 1));
                    tcgen05_fence_after();
                }
                const int s = it % STAGES3P;
                const uint32_t ph = (it / STAGES3P) & 1u;
                mbar_wait_cluster(ready_bar(s), ph);
                tcgen05_fence_after();
                const uint32_t sa = smem_base + s * STAGE3P_BYTES;
                const uint64_t ah = make_smem_desc(sa), al = make_smem_desc(sa + A_BYTES);
                const uint64_t bh = make_smem_desc(sa + 2 * A_BYTES), bl = make_smem_desc(sa + 2 * A_BYTES + BH_BYTES);
                const uint32_t d = tmem_base + (uint32_t)(b * BN);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint64_t o = (uint64_t)(2 * k);
                    umma2_tf32(d, al + o, bh + o, kIdesc2, (chunk_start && k == 0) ? 0u : 1u);        // small terms first
                    umma2_tf32(d, ah + o, bl + o, kIdesc2, 1u);
                    umma2_tf32(d, ah + o, bh + o, kIdesc2, 1u);
                }
                umma2_commit_both(empty_bar(s));
                if ((it % CHUNK) == CHUNK - 1 || it == kb - 1) umma2_commit_both(accfull_bar(b));
            }
        }
        __syncwarp();
    } else {
        // ---- warps 2..9: row = lane quarter (warp & 3), column half = 0 for warps 2-5, 1 for warps 6-9 ----
        const int q = warp & 3;
        const int half = (warp >= 6) ? 1 : 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
        float sum[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) sum[j] = 0.f;

        auto drain = [&](int c) {
            const int b = c & 1;
            mbar_wait(accfull_bar(b), (uint32_t)((c >> 1) & 1));
            tcgen05_fence_after();
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                uint32_t v[16];
                tmem_ld16(tlane + (uint32_t)(b * BN + cc * 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[cc * 16 + j] += __uint_as_float(v[j]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_rank0(accfree_bar(b)));
        };

        if (warp < 6) {
            const int et = threadIdx.x - 64;                        // 0..127: the four transform warps
            for (int it = 0; it < kb; ++it) {
                const int s = it % STAGES3P;
                const uint32_t ph = (it / STAGES3P) & 1u;
                mbar_wait(full_bar(s), ph);
                float4* a4 = reinterpret_cast<float4*>(smem_al + s * STAGE3P_BYTES);
                float4* l4 = reinterpret_cast<float4*>(smem_al + s * STAGE3P_BYTES + A_BYTES);
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / XF_THREADS; ++i) {
                    const float4 v = a4[et + i * XF_THREADS];
                    float4 h, l;
                    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
                    l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
                    a4[et + i * XF_THREADS] = h;
                    l4[et + i * XF_THREADS] = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_rank0(ready_bar(s)));
                if (((it % CHUNK) == CHUNK - 1 || it == kb - 1) && it / CHUNK >= 1) drain(it / CHUNK - 1);
            }
            drain(nchunks - 1);
        } else {
            for (int c = 0; c < nchunks; ++c) drain(c);
        }
        gemm3x_epilogue<EPI>(p, sum, m0 + q * 32 + lane, n0 + half * 128);
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Attention-shaped N x N contractions on tensor cores (fp32-grade 3xTF32):
//   out[b,h,i,j] = epi( alpha * sum_d A[b,i,h,d] * B[b,j,h,d] )        (Q K^T, dctx V^T, S2 V^T: K = head_dim)
// A and B are head slices of packed [batch*N, ld] activations, addressed in place by 2-D tensor maps
// (column = h*dh + kblock*32, row = b*N + tile row).  Rows of a tile that fall into the next sample (N is not a
// multiple of 128 / 256) only produce output rows / columns that the epilogue masks.  Both operands are activations,
// so both are split into hi/lo in shared memory.  One CTA per (b, h, 128-row tile), K (<= 64) streamed one 32-element
// k-block at a time through a single operand buffer, 3 MMAs per 8-wide k-step, one 128 x 256 fp32 accumulator in TMEM,
// fused epilogue (scale / multiply by E / safe_divide).
// =====================================================================================================================
// One k-block (32 of the <= 64 head-dim elements) is resident at a time: 96 KiB of operands (hi + lo of A and B), so
// TWO CTAs share an SM and the TMA wait / split / epilogue of one overlaps the MMAs of the other (a CTA's whole
// reduction is only 1-2 k-blocks; with everything resident — 192 KiB — the SM ran one CTA at a time, start to end).
constexpr int AT_SMEM = 2 * (A_BYTES + B_BYTES) + 1024 + 256;
enum { AT_STORE = 0, AT_MUL = 1, AT_SD = 2, AT_RESID = 3, AT_SOFTMAX = 4 };

struct AtParams {
    int N, H, dh, ld_out;            // tokens, heads, head_dim, row stride of out / E
    const float* E; float* out; float alpha;
};

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 2)
te_tc_attn_nn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AtParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    // layout (one k-block): A_hi | B_hi | A_lo | B_lo
    constexpr uint32_t OFF_AH = 0, OFF_BH = A_BYTES, OFF_AL = OFF_BH + B_BYTES, OFF_BL = OFF_AL + A_BYTES;
    constexpr uint32_t TOTAL = OFF_BL + B_BYTES;
    const uint32_t bars = smem_base + TOTAL;
    const uint32_t full_bar = bars, xf_bar = bars + 8, accum_bar = bars + 16, empty_bar = bars + 24;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + TOTAL + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.z * BN;                    // key-column tile (N > 256: BERT-512 has two)
    const int kb = p.dh / BK;
    constexpr uint32_t TMEM_COLS = 256u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        mbar_init(full_bar, 1);
        mbar_init(xf_bar, XF_THREADS / 32);
        mbar_init(accum_bar, 1);
        mbar_init(empty_bar, 1);
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
        if (lane == 0) {
            for (int k = 0; k < kb; ++k) {
                if (k > 0) mbar_wait(empty_bar, (uint32_t)((k - 1) & 1));        // MMAs of the previous k-block retired
                mbar_arrive_expect_tx(full_bar, (uint32_t)(A_BYTES + B_BYTES));
                tma_load_2d(smem_base + OFF_AH, &tmA, full_bar, h * p.dh + k * BK, b * p.N + m0);
                tma_load_2d(smem_base + OFF_BH, &tmB, full_bar, h * p.dh + k * BK, b * p.N + n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint64_t ah = make_smem_desc(smem_base + OFF_AH), al = make_smem_desc(smem_base + OFF_AL);
            const uint64_t bh_ = make_smem_desc(smem_base + OFF_BH), bl = make_smem_desc(smem_base + OFF_BL);
            for (int kk = 0; kk < kb; ++kk) {
                mbar_wait(xf_bar, (uint32_t)(kk & 1));
                tcgen05_fence_after();
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint64_t o = (uint64_t)(2 * k);
                    umma_tf32(tmem_base, al + o, bh_ + o, kIdesc, (kk == 0 && k == 0) ? 0u : 1u);
                    umma_tf32(tmem_base, ah + o, bl + o, kIdesc, 1u);
                    umma_tf32(tmem_base, ah + o, bh_ + o, kIdesc, 1u);
                }
                umma_commit(empty_bar);              // the operand buffer may be refilled when these MMAs retire
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        const int et = threadIdx.x - 64;
        // split A (16 KiB) and B (32 KiB) of every k-block: hi in place, lo to the *_lo regions (same swizzled offsets)
        for (int kk = 0; kk < kb; ++kk) {
            mbar_wait(full_bar, (uint32_t)(kk & 1));
            float4* a4 = reinterpret_cast<float4*>(smem_al + OFF_AH);
            float4* l4 = reinterpret_cast<float4*>(smem_al + OFF_AL);
#pragma unroll
            for (int i = et; i < A_BYTES / 16; i += XF_THREADS) {
                const float4 v = a4[i];
                float4 hh, l;
                hh.x = to_tf32(v.x); hh.y = to_tf32(v.y); hh.z = to_tf32(v.z); hh.w = to_tf32(v.w);
                l.x = to_tf32(v.x - hh.x); l.y = to_tf32(v.y - hh.y); l.z = to_tf32(v.z - hh.z); l.w = to_tf32(v.w - hh.w);
                a4[i] = hh; l4[i] = l;
            }
            float4* b4 = reinterpret_cast<float4*>(smem_al + OFF_BH);
            float4* m4 = reinterpret_cast<float4*>(smem_al + OFF_BL);
#pragma unroll 4
            for (int i = et; i < B_BYTES / 16; i += XF_THREADS) {
                const float4 v = b4[i];
                float4 hh, l;
                hh.x = to_tf32(v.x); hh.y = to_tf32(v.y); hh.z = to_tf32(v.z); hh.w = to_tf32(v.w);
                l.x = to_tf32(v.x - hh.x); l.y = to_tf32(v.y - hh.y); l.z = to_tf32(v.z - hh.z); l.w = to_tf32(v.w - hh.w);
                b4[i] = hh; m4[i] = l;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(xf_bar);
        }

        const int q = warp & 3;
        const int i = m0 + q * 32 + lane;                       // query row inside the sample
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool live = i < p.N;
        const long long rowoff = ((long long)bh * p.N + i) * p.ld_out + n0;
        const int ncols = min(p.N - n0, BN);              // valid key columns of this tile
        const int nchunks = (ncols + 31) / 32;
        // E (attention probabilities / attn_cam) comes from HBM: its loads are issued one 32-column chunk ahead so
        // that their latency overlaps the TMEM read, the math and the stores of the previous chunk.  Reading a
        // full float4 whose tail lies in the row padding is memory-safe (ld_out % 4 == 0); the tail is masked.
        if (EPI == AT_SOFTMAX) {
            // softmax(alpha * A B^T) over the key axis, fused: every thread owns one query row whose N <= 256 scores sit
            // in its TMEM lane, so the row maximum, the sum of exponentials and the normalised probabilities come from
            // three passes over TMEM (the exponentials are written back with tcgen05.st) — the scores never travel to HBM (attn = dots.softmax(dim=-1), ViT_LRP.py:139-141)
            mbar_wait(accum_bar, 0);
            tcgen05_fence_after();
            float mx = -INFINITY;
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t acc[32];
                tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (cc * 32 + j < ncols) mx = fmaxf(mx, p.alpha * __uint_as_float(acc[j]));
            }
            // pass 2: e = exp(score - max), summed, and written back over the scores in TMEM (one expf per element)
            float sum = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t acc[32];
                tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float e = (cc * 32 + j < ncols) ? expf(p.alpha * __uint_as_float(acc[j]) - mx) : 0.f;
                    sum += e;
                    acc[j] = __float_as_uint(e);
                }
                tmem_st32(tlane + (uint32_t)(cc * 32), acc);
            }
            tmem_st_wait();
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t acc[32];
                tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = cc * 32 + j * 4;
                        if (col < ncols) {
                            float o[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) o[u] = __uint_as_float(acc[j * 4 + u]) / sum;   // padding holds e = 0
                            *reinterpret_cast<float4*>(p.out + rowoff + col) = make_float4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
            }
        } else {
        float4 ebuf[2][8];
        auto load_e = [&](int c, float4 (&buf)[8]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = c * 32 + j * 4;
                buf[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (EPI != AT_STORE && live && col < ncols) buf[j] = __ldcs(reinterpret_cast<const float4*>(p.E + rowoff + col));
            }
        };
        load_e(0, ebuf[0]);
        mbar_wait(accum_bar, 0);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < nchunks; c += 2) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int cc = c + half;
                if (cc < nchunks) {
                    if (cc + 1 < nchunks) load_e(cc + 1, ebuf[half ^ 1]);
                    uint32_t acc[32];
                    tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
                    tmem_ld_wait();
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int col = cc * 32 + j * 4;
                            if (col < ncols) {
                                const float e[4] = {ebuf[half][j].x, ebuf[half][j].y, ebuf[half][j].z, ebuf[half][j].w};
                                float o[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float a = __uint_as_float(acc[j * 4 + u]);
                                    if (EPI == AT_STORE) o[u] = p.alpha * a;
                                    else if (EPI == AT_MUL) o[u] = p.alpha * a * e[u];
                                    else o[u] = te_sd(e[u], p.alpha * a);
                                }
                                if (col + 3 < ncols) *reinterpret_cast<float4*>(p.out + rowoff + col) = make_float4(o[0], o[1], o[2], o[3]);
                                else {
#pragma unroll
                                    for (int u = 0; u < 4; ++u) p.out[rowoff + col + u] = (col + u < ncols) ? o[u] : 0.f;   // zero the row padding
                                }
                            }
                        }
                    }
                }
            }
        }
        }   // EPI != AT_SOFTMAX
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Attention-shaped N x d contractions with the reduction over the TOKEN axis (fp32-grade 3xTF32):
//   out[b, m, h, :] = epi( alpha * sum_k A_h[m,k] * X[b, k, h, :] )
//     AMN = 0:  A_h[m,k] = map[b,h,m,k]   (attn v, dS k, S1 k)           -> A is K-major
//     AMN = 1:  A_h[m,k] = map[b,h,k,m]   (attn^T dctx, attn^T S2, dS^T q, S1^T q: the transposed map) -> A is MN-major
//   X (a head slice of a packed activation, [token, feature]) is always MN-major for this product.
// MN-major tf32 operands use the SWIZZLE_128B_BASE32B layout: 32 consecutive M/N elements (128 B) per K row, 4 K rows
// per 512-byte atom (SBO), 32-element M/N blocks LBO apart; a TMA box of 32 elements x 32 K rows with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B lands exactly as eight such atoms.  3-D tensor maps (col, token, batch*head | batch) make every row past the N tokens of
// a head read as zero, so K (= N = 197) is padded to 224 for free.  4-stage ring, both operands split hi/lo in smem.
// =====================================================================================================================
// NB = number of 32-wide output-column blocks: 2 for the head_dim-64 attention products, 7 (N <= 224) for the dense
// rollout product J <- (M_l + I) J, which is the same contraction with H = 1 (A = M_l K-major, B = J MN-major).
constexpr int NK_A = A_BYTES;                                      // 16 KiB
template <int NB> struct NkCfg {
    static constexpr int BN = NB * 32;
    static constexpr int B_BYTES_ = BN * BK * 4;                   // 4 KiB per block
    static constexpr int STAGE = 2 * NK_A + 2 * B_BYTES_;
    // attention shape (NB = 2): 2 stages of 48 KiB so that TWO CTAs share an SM — a CTA's whole reduction is only 7
    // k-blocks, and its prologue / epilogue then overlap the other CTA's main loop
    static constexpr int STAGES = 2;
    static constexpr int MIN_CTAS = (NB <= 2) ? 2 : 1;
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
    static constexpr int XF4 = (NK_A + B_BYTES_) / 16;
    // wide tiles (the dense rollout product) keep the lo*hi + hi*lo correction terms in a second accumulator at column
    // 256: the tensor core truncates on every accumulate, so the fewer (and the smaller) the addends an accumulator
    // sees after it holds a large value, the smaller the drift
    static constexpr bool SPLIT_ACC = NB >= 7;
    static constexpr uint32_t TMEM_COLS = SPLIT_ACC ? 512u : ((BN <= 64) ? 64u : (BN <= 128 ? 128u : 256u));
};

struct NkParams {
    int N, H, ld_out;                 // tokens (= reduction length and output rows), heads, row stride of out / E
    int n_out;                        // valid output columns of the whole row (all heads / column tiles); columns in
    int n_pad;                        // [n_out, n_pad) are written as zero, columns >= n_pad are not touched
    int a_shared;                     // 1: A is indexed by the batch only (dense product, "heads" are column tiles)
    const float* rowscale;            // AT_RESID: out = acc + rowscale[b*N + m] * E  (null: 1)
    const float* E; float* out; float alpha;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// MN-major descriptor for 32-bit (tf32) operands.  The only layout the tensor core accepts for MN-major tf32 is
// SWIZZLE_128B_BASE32B (cute::UMMA::Layout_MN_SW128_32B_Atom: 32 M/N elements = one 128-byte row per K row, atoms of
// 4 K rows = 512 B, 32-byte chunks XOR-swizzled by (K row % 4)); TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
// LBO = byte distance between 32-element M/N blocks, SBO = 512 B between 4-row K atoms, layout type 1.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

template <int AMN, int EPI, int NB>
__global__ void __launch_bounds__(NUM_THREADS, NkCfg<NB>::MIN_CTAS)
te_tc_attn_nk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const NkParams p) {
    using C = NkCfg<NB>;
    constexpr int NK_BN = C::BN, NK_B = C::B_BYTES_, NK_STAGE = C::STAGE, NK_STAGES = C::STAGES, NK_XF4 = C::XF4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + NK_STAGES * NK_STAGE;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xf_bar = [&](int s) { return bars + 8u * (NK_STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * NK_STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (3 * NK_STAGES);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + NK_STAGES * NK_STAGE + 8 * (3 * NK_STAGES + 1));
    constexpr uint32_t OFF_AL = NK_A, OFF_BH = 2 * NK_A, OFF_BL = 2 * NK_A + NK_B;
    // instruction descriptor: tf32, M = 128, N = 64, A major = AMN, B major = MN
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)AMN << 15) | (1u << 16) |
                               ((uint32_t)(NK_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
    const int m0 = blockIdx.x * BM;
    const int kb = (p.N + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = C::TMEM_COLS;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < NK_STAGES; ++s) {
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
        if (lane == 0) {
            for (int it = 0; it < kb; ++it) {
                const int s = it % NK_STAGES;
                const uint32_t ph = (it / NK_STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), NK_A + NK_B);
                const uint32_t sa = smem_base + s * NK_STAGE;
                const int k0 = it * BK;
                if (AMN == 0) {
                    tma_load_3d(sa, &tmA, full_bar(s), k0, m0, p.a_shared ? b : bh);     // [128 rows m] x [32 k], K-major
                } else {
#pragma unroll
                    for (int mb = 0; mb < BM / 32; ++mb)                                 // four [32 m] x [32 k rows] blocks
                        tma_load_3d(sa + mb * 4096, &tmA, full_bar(s), m0 + mb * 32, k0, p.a_shared ? b : bh);
                }
#pragma unroll
                for (int nb = 0; nb < NK_BN / 32; ++nb)                                  // two [32 d] x [32 k rows] blocks
                    tma_load_3d(sa + OFF_BH + nb * 4096, &tmB, full_bar(s), h * NK_BN + nb * 32, k0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < kb; ++it) {
                const int s = it % NK_STAGES;
                const uint32_t ph = (it / NK_STAGES) & 1u;
                mbar_wait(xf_bar(s), ph);
                tcgen05_fence_after();
                const uint32_t sa = smem_base + s * NK_STAGE;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    uint64_t ah, al;
                    if (AMN == 0) {
                        ah = make_smem_desc(sa) + (uint64_t)(2 * k);                      // +32 B along the K-major row
                        al = make_smem_desc(sa + OFF_AL) + (uint64_t)(2 * k);
                    } else {
                        ah = make_smem_desc_mn(sa + k * 1024, 4096);                      // +8 K rows = one 1 KiB atom
                        al = make_smem_desc_mn(sa + OFF_AL + k * 1024, 4096);
                    }
                    const uint64_t bhd = make_smem_desc_mn(sa + OFF_BH + k * 1024, 4096);
                    const uint64_t bld = make_smem_desc_mn(sa + OFF_BL + k * 1024, 4096);
                    const uint32_t first = (it == 0 && k == 0) ? 0u : 1u;
                    if (C::SPLIT_ACC) {
                        umma_tf32(tmem_base + 256u, al, bhd, idesc, first);
                        umma_tf32(tmem_base + 256u, ah, bld, idesc, 1u);
                        umma_tf32(tmem_base, ah, bhd, idesc, first);
                    } else {
                        umma_tf32(tmem_base, al, bhd, idesc, first);
                        umma_tf32(tmem_base, ah, bld, idesc, 1u);
                        umma_tf32(tmem_base, ah, bhd, idesc, 1u);
                    }
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        const int et = threadIdx.x - 64;
        for (int it = 0; it < kb; ++it) {
            const int s = it % NK_STAGES;
            const uint32_t ph = (it / NK_STAGES) & 1u;
            mbar_wait(full_bar(s), ph);
            // split A_hi (16 KiB) and B_hi (8 KiB) slots -> hi in place, lo into the matching *_lo slot
            float4* base4 = reinterpret_cast<float4*>(smem_al + s * NK_STAGE);
            for (int i = et; i < NK_XF4; i += XF_THREADS) {
                const bool isA = i < NK_A / 16;
                float4* src = isA ? base4 + i : base4 + (OFF_BH / 16) + (i - NK_A / 16);
                float4* dst = isA ? base4 + (OFF_AL / 16) + i : base4 + (OFF_BL / 16) + (i - NK_A / 16);
                const float4 v = *src;
                float4 hh, l;
                hh.x = to_tf32(v.x); hh.y = to_tf32(v.y); hh.z = to_tf32(v.z); hh.w = to_tf32(v.w);
                l.x = to_tf32(v.x - hh.x); l.y = to_tf32(v.y - hh.y); l.z = to_tf32(v.z - hh.z); l.w = to_tf32(v.w - hh.w);
                *src = hh; *dst = l;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(xf_bar(s));              // one arrive per transform warp
        }
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        const bool live = m < p.N;
        const long long off = ((long long)b * p.N + m) * p.ld_out + (long long)h * NK_BN;
        float4 ebuf[(EPI == AT_MUL) ? NB * 8 : 1];
        if (EPI == AT_MUL && live) {
#pragma unroll
            for (int j = 0; j < NB * 8; ++j) ebuf[j] = *reinterpret_cast<const float4*>(p.E + off + j * 4);
        }
        mbar_wait(accum_bar, 0);
        tcgen05_fence_after();
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float rscale = (EPI == AT_RESID && live && p.rowscale) ? p.rowscale[(long long)b * p.N + m] : 1.f;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            uint32_t acc[32];
            tmem_ld32(tlane + (uint32_t)(c * 32), acc);
            float4 rbuf[(EPI == AT_RESID) ? 8 : 1];
            if (EPI == AT_RESID && live) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    rbuf[(EPI == AT_RESID) ? j : 0] = (h * NK_BN + c * 32 + j * 4 < p.n_pad)
                                                         ? *reinterpret_cast<const float4*>(p.E + off + c * 32 + j * 4)
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            tmem_ld_wait();
            if (C::SPLIT_ACC) {
                uint32_t acc2[32];
                tmem_ld32(tlane + 256u + (uint32_t)(c * 32), acc2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(acc2[j]));
            }
            if (live) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = c * 32 + j * 4;
                    const int gcol = h * NK_BN + col;
                    if (gcol < p.n_pad) {
                        float o[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float a = __uint_as_float(acc[j * 4 + u]);
                            float v = p.alpha * a;
                            if (EPI == AT_MUL) {
                                const float4 e4 = ebuf[(EPI == AT_MUL) ? c * 8 + j : 0];
                                const float e = (u == 0) ? e4.x : (u == 1) ? e4.y : (u == 2) ? e4.z : e4.w;
                                v *= e;
                            }
                            if (EPI == AT_RESID) {
                                const float4 e4 = rbuf[(EPI == AT_RESID) ? j : 0];
                                const float e = (u == 0) ? e4.x : (u == 1) ? e4.y : (u == 2) ? e4.z : e4.w;
                                v += rscale * e;
                            }
                            o[u] = (gcol + u < p.n_out) ? v : 0.f;                    // zero the row padding
                        }
                        *reinterpret_cast<float4*>(p.out + off + col) = make_float4(o[0], o[1], o[2], o[3]);
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

// ---- host: tensor maps ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// row-major [rows, cols] with row stride ld (elements); box = [box_rows, one 128-byte row], 128-byte swizzle
bool make_map_t(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows, bool bf16) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const int esz = bf16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool make_map(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_rows) {
    return make_map_t(m, base, rows, cols, ld, box_rows, false);
}

template <int MODE>
int launch(const float* A, long long lda, const float* B0, const float* B1, const float* E, long long lde, float* C,
           long long ldc, long long M, int N, int K, cudaStream_t st, const float* Y = nullptr, long long ldy = 0,
           const float* bias = nullptr, int out_bf16 = 0) {
    CUtensorMap tmA, tmB0, tmB1;
    if (!make_map(&tmA, A, M, K, lda, BM) || !make_map(&tmB0, B0, N, K, K, BN) || !make_map(&tmB1, B1, N, K, K, BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static bool attr_set[4] = {false, false, false, false};
    if (!attr_set[MODE]) {
        if (cudaFuncSetAttribute(te_tc_zplus_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ZpCfg<MODE>::SMEM) != cudaSuccess) {
            te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
            return TE_ERR_CUDA;
        }
        attr_set[MODE] = true;
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
int g_pair_linear = -1;                 // 3xTF32 Linear GEMMs as CTA pairs (TE_B200_LINEAR_2CTA=1 / te_set_option)
bool use_pair_linear() {
    if (g_pair_linear < 0) {
        const char* e = getenv("TE_B200_LINEAR_2CTA");
        g_pair_linear = (e && e[0] == '1') ? 1 : 0;
    }
    return g_pair_linear == 1;
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
    static bool attr_set[4] = {false, false, false, false};
    if (!attr_set[MODE]) {
        if (cudaFuncSetAttribute(te_tc_zplus2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES) != cudaSuccess) {
            te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
            return TE_ERR_CUDA;
        }
        attr_set[MODE] = true;
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
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(te_tc_zplus_kernel<MODE_R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
            te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
            return TE_ERR_CUDA;
        }
        attr_set = true;
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.M = (int)M; p.N = N; p.K = K; p.E = E; p.lde = lde; p.C = C; p.ldc = ldc;
    dim3 grid(N / BN, (unsigned)((M + BM - 1) / BM));
    te_tc_zplus_kernel<MODE_R, true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB0, tmB1, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

inline bool a16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

bool te_tc_zplus_supported(long long rows, int in_features, int out_features, long long ldx) {
    return rows > 0 && rows < (1LL << 31) && in_features % BN == 0 && out_features % BN == 0 && ldx % 4 == 0 &&
           get_encode() != nullptr;
}

void te_tc_set_pair_linear(int on) { g_pair_linear = on ? 1 : 0; }
void te_tc_set_pair_kernels(int on) { g_pair_kernels = (on == 2) ? 2 : (on ? 1 : 0); }

long long te_tc_derived_floats(int in_features, int out_features) { return 10LL * in_features * out_features; }

int te_tc_prepare_weights(const float* w, float* derived, int in_features, int out_features, cudaStream_t st) {
    dim3 grid((in_features + 31) / 32, (out_features + 31) / 32), block(32, 8);
    prepare_weights_kernel<<<grid, block, 0, st>>>(w, derived, out_features, in_features);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

bool te_tc_gemm3x_supported(long long rows, int K, int N, long long lda) {
    return rows > 0 && rows < (1LL << 31) && K % BK == 0 && N % BN == 0 && lda % 4 == 0 && get_encode() != nullptr;
}

namespace {
template <int EPI>
int launch3(const float* A, long long lda, const float* Bh, const float* Bl, const Tc3Params& p, cudaStream_t st) {
    CUtensorMap tmA, tmBh, tmBl;
    if (!make_map(&tmA, A, p.M, p.K, lda, BM) || !make_map(&tmBh, Bh, p.N, p.K, p.K, BN) ||
        !make_map(&tmBl, Bl, p.N, p.K, p.K, BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(te_tc_gemm3x_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3_BYTES) != cudaSuccess) {
            te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
            return TE_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(p.N / BN, (unsigned)((p.M + BM - 1) / BM));
    te_tc_gemm3x_kernel<EPI><<<grid, NUM_THREADS3, SMEM3_BYTES, st>>>(tmA, tmBh, tmBl, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

template <int EPI>
int launch3_pair(const float* A, long long lda, const float* Bh, const float* Bl, const Tc3Params& p, cudaStream_t st) {
    CUtensorMap tmA, tmBh, tmBl;
    if (!make_map(&tmA, A, p.M, p.K, lda, BM) || !make_map(&tmBh, Bh, p.N, p.K, p.K, BN / 2) ||
        !make_map(&tmBl, Bl, p.N, p.K, p.K, BN / 2)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(te_tc_gemm3x2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3P_BYTES) != cudaSuccess) {
            te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
            return TE_ERR_CUDA;
        }
        attr_set = true;
    }
    const unsigned mtiles = (unsigned)((p.M + BM - 1) / BM);
    dim3 grid((unsigned)(p.N / BN) * ((mtiles + 1u) & ~1u));
    te_tc_gemm3x2_kernel<EPI><<<grid, NUM_THREADS3, SMEM3P_BYTES, st>>>(tmA, tmBh, tmBl, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

int dispatch3(int epi, const float* A, long long lda, const float* Bh, const float* Bl, const Tc3Params& p, cudaStream_t st) {
    if (use_pair_linear()) {
        switch (epi) {
            case TE_TC_EPI_STORE: return launch3_pair<EP_STORE>(A, lda, Bh, Bl, p, st);
            case TE_TC_EPI_BIAS: return launch3_pair<EP_BIAS>(A, lda, Bh, Bl, p, st);
            case TE_TC_EPI_BIAS_GELU: return launch3_pair<EP_BIAS_GELU>(A, lda, Bh, Bl, p, st);
            case TE_TC_EPI_BIAS_ADD: return launch3_pair<EP_BIAS_ADD>(A, lda, Bh, Bl, p, st);
            case TE_TC_EPI_GELU_BWD: return launch3_pair<EP_GELU_BWD>(A, lda, Bh, Bl, p, st);
        }
    }
    switch (epi) {
        case TE_TC_EPI_STORE: return launch3<EP_STORE>(A, lda, Bh, Bl, p, st);
        case TE_TC_EPI_BIAS: return launch3<EP_BIAS>(A, lda, Bh, Bl, p, st);
        case TE_TC_EPI_BIAS_GELU: return launch3<EP_BIAS_GELU>(A, lda, Bh, Bl, p, st);
        case TE_TC_EPI_BIAS_ADD: return launch3<EP_BIAS_ADD>(A, lda, Bh, Bl, p, st);
        case TE_TC_EPI_GELU_BWD: return launch3<EP_GELU_BWD>(A, lda, Bh, Bl, p, st);
    }
    te_set_last_error("te_gemm_tc: unsupported epilogue");
    return TE_ERR_UNSUPPORTED;
}
}  // namespace

bool te_tc_attn_supported(int N, int dh, long long lda, long long ldb, int ld_out) {
    return N >= 1 && (dh == 32 || dh == 64) && lda % 4 == 0 && ldb % 4 == 0 && ld_out % 4 == 0 && get_encode() != nullptr;
}

namespace {
template <int EPI>
int launch_attn(const float* A, long long lda, const float* B, long long ldb, long long total_rows, const AtParams& p,
                int batch, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, A, total_rows, (long long)p.H * p.dh, lda, BM) || !make_map(&tmB, B, total_rows, (long long)p.H * p.dh, ldb, BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (attention)");
        return TE_ERR_CUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(te_tc_attn_nn_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) {
            te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
            return TE_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid((p.N + BM - 1) / BM, batch * p.H, (p.N + BN - 1) / BN);
    if (grid.y > 65535) { te_set_last_error("te_gemm_tc: batch*heads too large for one launch"); return TE_ERR_ARG; }
    te_tc_attn_nn_kernel<EPI><<<grid, NUM_THREADS, AT_SMEM, st>>>(tmA, tmB, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}
}  // namespace

// out[b,h,i,j] = epi(alpha * sum_d A[b*N+i, h*dh+d] * B[b*N+j, h*dh+d]);  out / E are [batch,H,N,ld_out]
int te_tc_attn_nn(const float* A, long long lda, const float* B, long long ldb, int batch, int H, int N, int dh,
                  float* out, int ld_out, const float* E, float alpha, int epi, cudaStream_t st) {
    AtParams p;
    p.N = N; p.H = H; p.dh = dh; p.ld_out = ld_out; p.E = E; p.out = out; p.alpha = alpha;
    const long long rows = (long long)batch * N;
    switch (epi) {
        case TE_TC_ATTN_STORE: return launch_attn<AT_STORE>(A, lda, B, ldb, rows, p, batch, st);
        case TE_TC_ATTN_MUL: return launch_attn<AT_MUL>(A, lda, B, ldb, rows, p, batch, st);
        case TE_TC_ATTN_SD: return launch_attn<AT_SD>(A, lda, B, ldb, rows, p, batch, st);
        case TE_TC_ATTN_SOFTMAX:
            if (N > BN) break;                        // the whole key axis must sit in one accumulator
            return launch_attn<AT_SOFTMAX>(A, lda, B, ldb, rows, p, batch, st);
    }
    te_set_last_error("te_gemm_tc: unsupported attention epilogue");
    return TE_ERR_UNSUPPORTED;
}

bool te_tc_attn_nk_supported(int N, int dh, int NP, long long ldx, long long ld_out) {
    return N >= 1 && dh == 64 && NP % 4 == 0 && ldx % 4 == 0 && ld_out % 4 == 0 && get_encode() != nullptr;
}
bool te_tc_bmm_nk_supported(int N, int ld) { return N >= 1 && ld % 4 == 0 && ld >= N && get_encode() != nullptr; }

namespace {
// rank-3 fp32 map: dims {cols, rows, batch}, box {bc, br, 1}, 128-byte swizzle, zero fill outside
bool make_map3(CUtensorMap* m, const float* base, long long cols, long long rows, long long batch, long long row_stride,
               long long batch_stride, int box_cols, int box_rows, bool mn_major) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)batch_stride * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int AMN, int EPI, int NB>
int launch_nk(const float* map, int NP, const float* X, long long ldx, const NkParams& p, int batch, cudaStream_t st) {
    constexpr int NK_SMEM = NkCfg<NB>::SMEM;
    CUtensorMap tmA, tmB;
    // attention-shaped map [batch*H, N, NP] ; activation [batch, N, ldx]
    if (!make_map3(&tmA, map, NP, p.N, p.a_shared ? (long long)batch : (long long)batch * p.H, NP, (long long)p.N * NP, 32,
                   AMN ? 32 : BM, AMN != 0) ||
        !make_map3(&tmB, X, ldx, p.N, batch, ldx, (long long)p.N * ldx, 32, 32, true)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (attention nk)");
        return TE_ERR_CUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(te_tc_attn_nk_kernel<AMN, EPI, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, NK_SMEM) != cudaSuccess) {
            te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
            return TE_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid((p.N + BM - 1) / BM, batch * p.H);
    if (grid.y > 65535) { te_set_last_error("te_gemm_tc: batch*heads too large for one launch"); return TE_ERR_ARG; }
    te_tc_attn_nk_kernel<AMN, EPI, NB><<<grid, NUM_THREADS, NK_SMEM, st>>>(tmA, tmB, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}
}  // namespace

// out[b, m, h, :] = epi(alpha * sum_k A_h[m,k] X[b,k,h,:]) ; A_h = map[b,h] (amn = 0) or its transpose (amn = 1);
// X, out, E: packed activations [batch, N, ld] (head h at columns h*64..); epi: TE_TC_ATTN_STORE / TE_TC_ATTN_MUL
int te_tc_attn_nk(const float* map, int NP, int amn, const float* X, long long ldx, int batch, int H, int N, float* out,
                  int ld_out, const float* E, float alpha, int epi, cudaStream_t st) {
    NkParams p;
    p.N = N; p.H = H; p.ld_out = ld_out; p.n_out = H * 64; p.n_pad = H * 64; p.a_shared = 0; p.rowscale = nullptr;
    p.E = E; p.out = out; p.alpha = alpha;
    if (epi == TE_TC_ATTN_STORE) return amn ? launch_nk<1, AT_STORE, 2>(map, NP, X, ldx, p, batch, st) : launch_nk<0, AT_STORE, 2>(map, NP, X, ldx, p, batch, st);
    if (epi == TE_TC_ATTN_MUL) return amn ? launch_nk<1, AT_MUL, 2>(map, NP, X, ldx, p, batch, st) : launch_nk<0, AT_MUL, 2>(map, NP, X, ldx, p, batch, st);
    te_set_last_error("te_gemm_tc: unsupported attention nk epilogue");
    return TE_ERR_UNSUPPORTED;
}

// One step of the rollout chain in residual form: out[b] = A[b] * J[b] + diag(rowscale[b]) * J[b], all [batch, N, ld]
// (fp32-grade 3xTF32; A K-major, J MN-major).  A is the layer matrix WITHOUT its identity part (mean_h relu(G*cam),
// divided by the row sum for BERT) and rowscale the identity's weight (null: 1; BERT: 1 / rowsum) — the large I * J term
// is added in fp32 in the epilogue instead of being pushed through the truncating tensor-core accumulator.
// One 128 x 224 tile per CTA when N <= 224 (ViT / DeiT), 128 x 256 column tiles otherwise (BERT-512: 4 x 2 CTAs per
// sample).  The padding columns of out are zeroed so that it can be the next J.
int te_tc_bmm_nk_resid(const float* A, const float* J, const float* rowscale, float* out, int batch, int N, int ld,
                       cudaStream_t st) {
    NkParams p;
    p.N = N; p.ld_out = ld; p.n_out = N; p.n_pad = ld; p.a_shared = 1; p.rowscale = rowscale; p.E = J; p.out = out;
    p.alpha = 1.f;
    if (ld <= 224) {
        p.H = 1;
        return launch_nk<0, AT_RESID, 7>(A, ld, J, ld, p, batch, st);
    }
    p.H = (ld + 255) / 256;
    return launch_nk<0, AT_RESID, 8>(A, ld, J, ld, p, batch, st);
}

// y[rows,out] = x[rows,in] W^T (+ epilogue)   — fp32-grade (3xTF32) on tcgen05
int te_tc_linear_fwd(const float* x, long long ldx, const float* derived, int in_features, int out_features,
                     const float* bias, float* y, float* y2, const float* e0, long long rows, int epi, cudaStream_t st) {
    const long long n = (long long)in_features * out_features;
    Tc3Params p;
    p.M = (int)rows; p.N = out_features; p.K = in_features; p.bias = bias; p.E = e0; p.lde = out_features;
    p.C = y; p.ldc = out_features; p.C2 = y2; p.ldc2 = out_features;
    return dispatch3(epi, x, ldx, derived + 4 * n, derived + 5 * n, p, st);
}
// dx[rows,in] = dy[rows,out] W (+ epilogue)
int te_tc_linear_bwd(const float* dy, const float* derived, int in_features, int out_features, float* dx, const float* e0,
                     long long rows, int epi, cudaStream_t st) {
    const long long n = (long long)in_features * out_features;
    Tc3Params p;
    p.M = (int)rows; p.N = in_features; p.K = out_features; p.bias = nullptr; p.E = e0; p.lde = in_features;
    p.C = dx; p.ldc = in_features; p.C2 = nullptr; p.ldc2 = 0;
    return dispatch3(epi, dy, out_features, derived + 6 * n, derived + 7 * n, p, st);
}

int te_tc_zplus_linear_relprop(const float* x, long long ldx, const float* derived, const float* r, long long ldr,
                               float* out, float* s_scratch, long long rows, int in_features, int out_features, cudaStream_t st,
                               const float* y, long long ldy, const float* bias, bool bf16) {
    if (!a16(x) || !a16(derived) || !a16(r) || !a16(out) || !a16(s_scratch)) {
        te_set_last_error("te_gemm_tc: operands must be 16-byte aligned");
        return TE_ERR_ARG;
    }
    const long long n = (long long)in_features * out_features;
    const float *wp = derived, *wn = derived + n, *wpt = derived + 2 * n, *wnt = derived + 3 * n;
    // S = sd(R, x+ W+^T + x- W-^T)          A = x [rows, in] ; B = W+/- [out, in]
    const bool rb = bf16 && (out_features % 64 == 0);          // S as bf16, R kernel with bf16 operands (kind::f16)
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
