// fp32-grade (3xTF32) Linear GEMMs of the forward pass and of the activation-gradient backward on tcgen05:
// single-CTA kernel (default) and CTA-pair kernel (opt-in).  Shared PTX wrappers: te_tc_common.cuh.
#include "te_tc_common.cuh"

namespace {

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

// epilogue of both 3xTF32 Linear kernels from the fp32 register sums (one tile row x 128 columns per thread).  Global
// memory is accessed in the transposed layout of epi_read_t (te_tc_common.cuh: 4 rows x 128 contiguous bytes per warp
// instruction instead of 32 rows x 16 bytes): each 32-column chunk of the sums goes through the warp's staging buffer.
// stage: 4608 bytes private to the warp — the idle operand ring (every MMA of the tile has retired when this runs).
template <int EPI>
__device__ __forceinline__ void gemm3x_epilogue(const Tc3Params& p, const float (&sum)[128], float* stage, int lane, int row0,
                                                int cbase) {
    const int tr = lane >> 3, tc = 4 * (lane & 7);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(sum[cc * 32 + j]);
        epi_stage_rows(stage, lane, v);
        const int col = cbase + cc * 32 + tc;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((EPI == EP_BIAS || EPI == EP_BIAS_GELU || EPI == EP_BIAS_ADD) && p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row0 + 4 * i + tr;
            if (row >= p.M) continue;
            const float4 a = epi_read_t(stage, lane, i);
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI == EP_BIAS_ADD || EPI == EP_GELU_BWD) e = *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col);
            float4 o, o2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI == EP_STORE) o = a;
            else if (EPI == EP_BIAS) o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
            else if (EPI == EP_BIAS_GELU) {
                o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
                o2 = make_float4(te_gelu(o.x), te_gelu(o.y), te_gelu(o.z), te_gelu(o.w));
            } else if (EPI == EP_BIAS_ADD) {
                o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
                o2 = make_float4(e.x + o.x, e.y + o.y, e.z + o.z, e.w + o.w);
            } else {
                o = make_float4(a.x * te_gelu_grad(e.x), a.y * te_gelu_grad(e.y), a.z * te_gelu_grad(e.z), a.w * te_gelu_grad(e.w));
            }
            *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = o;
            if (EPI == EP_BIAS_GELU || EPI == EP_BIAS_ADD) *reinterpret_cast<float4*>(p.C2 + (long long)row * p.ldc2 + col) = o2;
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
        gemm3x_epilogue<EPI>(p, sum, reinterpret_cast<float*>(smem_al + (warp - 2) * EPI_STAGE_BYTES), lane, m0 + q * 32,
                             n0 + half * 128);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Mixed-kind fp32-grade Linear GEMM: the two CORRECTION terms of the error-compensated split run as bf16 MMAs.
//   C = A B^T ~= A_hi B_hi^T (kind::tf32)  +  bf16(A_lo) bf16(B_hi)^T  +  bf16(A_hi) bf16(B_lo)^T (kind::f16, fp32 accumulate)
// The correction terms are 2^-11 of the main term, so the 2^-8 relative error of a bf16 x bf16 product contributes 2^-19 —
// below the fp32 rounding of the result (simulated: 7.0e-7 of the result maximum against 6.2e-7 for a plain fp32 GEMM and
// 7.5e-8 for the exact three-term sum, K = 768 and 3072).  A bf16 MMA covers K = 16 per issue at the cycle cost of a K = 8 TF32
// MMA, so a 32-element k-block takes 4 + 2 + 2 = 8 MMA slots instead of 12: the 3xTF32 kernel sits at 0.81 of the TF32 issue
// roof (0.27 algorithmic), this form needs two thirds of its tensor cycles.
// Stage (96 KiB, 2 stages): A_hi tf32 16K | bf16(A_hi) 8K | bf16(A_lo) 8K | B_hi tf32 32K | bf16(B_hi) 16K | bf16(B_lo) 16K.
// The bf16 tiles are K-major with 64-byte rows in the SWIZZLE_64B layout (16-byte chunk index XOR (row / 2) % 4): TMA writes
// the weight tiles that way (CU_TENSOR_MAP_SWIZZLE_64B), the transform warps write the activation tiles.
// =====================================================================================================================
constexpr int A16_BYTES = BM * BK * 2;                            // 8 KiB
constexpr int B16_BYTES = BN * BK * 2;                            // 16 KiB
constexpr uint32_t M_OFF_AH = 0, M_OFF_A16H = A_BYTES, M_OFF_A16L = A_BYTES + A16_BYTES, M_OFF_BH = A_BYTES + 2 * A16_BYTES,
                   M_OFF_B16H = M_OFF_BH + B_BYTES, M_OFF_B16L = M_OFF_B16H + B16_BYTES;
static_assert(M_OFF_B16L + B16_BYTES == STAGE3_BYTES, "mixed stage = 96 KiB");

// K-major SWIZZLE_64B descriptor: 64-byte rows, 8-row (512 B) atoms
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS3, 1)
te_tc_gemm3m_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                    const __grid_constant__ CUtensorMap tmB16h, const __grid_constant__ CUtensorMap tmB16l, const Tc3Params p) {
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
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB16h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB16l) : "memory");
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
                mbar_arrive_expect_tx(full_bar(s), A_BYTES + B_BYTES + 2 * B16_BYTES);
                const uint32_t sa = smem_base + s * STAGE3_BYTES;
                tma_load_2d(sa + M_OFF_AH, &tmA, full_bar(s), it * BK, m0);                  // raw A -> A_hi slot
                tma_load_2d(sa + M_OFF_BH, &tmBh, full_bar(s), it * BK, n0);
                tma_load_2d(sa + M_OFF_B16H, &tmB16h, full_bar(s), it * BK, n0);
                tma_load_2d(sa + M_OFF_B16L, &tmB16l, full_bar(s), it * BK, n0);
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
                const uint64_t ah = make_smem_desc(sa + M_OFF_AH), bh = make_smem_desc(sa + M_OFF_BH);
                const uint64_t a16h = make_smem_desc_sw64(sa + M_OFF_A16H), a16l = make_smem_desc_sw64(sa + M_OFF_A16L);
                const uint64_t b16h = make_smem_desc_sw64(sa + M_OFF_B16H), b16l = make_smem_desc_sw64(sa + M_OFF_B16L);
                const uint32_t d = tmem_base + (uint32_t)(b * BN);
                // small terms first: two bf16 MMAs (K = 16 each) per correction term, then four TF32 MMAs (K = 8) of the main term
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t o = (uint64_t)(2 * k);
                    umma_bf16(d, a16l + o, b16h + o, kIdescBf16, (chunk_start && k == 0) ? 0u : 1u);
                    umma_bf16(d, a16h + o, b16l + o, kIdescBf16, 1u);
                }
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) umma_tf32(d, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), kIdesc, 1u);
                umma_commit(empty_bar(s));
                if ((it % CHUNK) == CHUNK - 1 || it == kb - 1) umma_commit(accfull_bar(b));
            }
        }
        __syncwarp();
    } else {
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
                uint8_t* st8 = smem_al + s * STAGE3_BYTES;
                float4* a4 = reinterpret_cast<float4*>(st8 + M_OFF_AH);
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / XF_THREADS; ++i) {
                    const int idx = et + i * XF_THREADS;            // float4 index inside the SWIZZLE_128B tile
                    const int r = idx >> 3, lc = (idx & 7) ^ (r & 7);   // row, LOGICAL 16-byte chunk (4 k values)
                    const float4 v = a4[idx];
                    float4 h;
                    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
                    a4[idx] = h;
                    const __nv_bfloat162 h01 = __floats2bfloat162_rn(h.x, h.y), h23 = __floats2bfloat162_rn(h.z, h.w);
                    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - h.x, v.y - h.y), l23 = __floats2bfloat162_rn(v.z - h.z, v.w - h.w);
                    // bf16 tile: row r (64 bytes), logical 8-byte slot lc -> 16-byte chunk (lc >> 1) ^ ((r >> 1) & 3), half lc & 1
                    const uint32_t off = (uint32_t)r * 64u + ((uint32_t)((lc >> 1) ^ ((r >> 1) & 3)) << 4) + ((uint32_t)(lc & 1) << 3);
                    uint2 hb, lb;
                    hb.x = *reinterpret_cast<const uint32_t*>(&h01); hb.y = *reinterpret_cast<const uint32_t*>(&h23);
                    lb.x = *reinterpret_cast<const uint32_t*>(&l01); lb.y = *reinterpret_cast<const uint32_t*>(&l23);
                    *reinterpret_cast<uint2*>(st8 + M_OFF_A16H + off) = hb;
                    *reinterpret_cast<uint2*>(st8 + M_OFF_A16L + off) = lb;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(xf_bar(s));
                if (((it % CHUNK) == CHUNK - 1 || it == kb - 1) && it / CHUNK >= 1) drain(it / CHUNK - 1);
            }
            drain(nchunks - 1);
        } else {
            for (int c = 0; c < nchunks; ++c) drain(c);
        }
        gemm3x_epilogue<EPI>(p, sum, reinterpret_cast<float*>(smem_al + (warp - 2) * EPI_STAGE_BYTES), lane, m0 + q * 32,
                             n0 + half * 128);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Persistent CTA-pair form of the mixed-kind fp32-grade Linear GEMM.  The single-CTA kernels above stage 80 KiB per k-block
// (raw A + the weight tiles) = 53-73 B/clk/SM against the chip-wide L2 -> SM cap of ~42 B/clk/SM: they are L2-bound (the
// mixed-kind kernel removed a third of the MMA cycles and its time did not change).  Here a CTA pair (tcgen05 cta_group::2,
// 256 x 256 tile) stages per CTA its own A tile and HALF of every weight tile: 16 K raw A + 16 K W_hi half + 8 K + 8 K bf16
// halves = 48 KiB per k-block of 8 MMA slots (4 TF32 + 2 + 2 bf16, 256 x 256) = 47 B/clk/SM.
//   warp 0 TMA (both CTAs) · warp 1 MMA (leader) · warps 2-3 hi / bf16 split of the A tile · warps 4-11 chunk drain + epilogue
//   full[s] local TMA bytes · ready[s] leader, one remote arrive per split warp of both CTAs (4) · empty[s] local multicast
//   commit · accfull[b] local multicast commit per 128-element chunk · accfree[b] leader, one remote arrive per drain warp (16)
// Persistent: static round-robin over 256 x 256 tiles (column tile fastest); barriers / TMEM set up once; chunks alternate
// between the two 256-column TMEM accumulators ACROSS tile boundaries, so while the drain warps store tile i from their
// register sums the MMAs of tile i+1's first two chunks already run.
// =====================================================================================================================
constexpr int MP_STAGES = 3;
constexpr int MP_STAGE = A_BYTES + 2 * A16_BYTES + BH_BYTES + B16_BYTES;     // 64 KiB: A_hi | a16h | a16l | Bh half | b16h half | b16l half
constexpr uint32_t MP_OFF_AH = 0, MP_OFF_A16H = A_BYTES, MP_OFF_A16L = A_BYTES + A16_BYTES, MP_OFF_BH = A_BYTES + 2 * A16_BYTES,
                   MP_OFF_B16H = MP_OFF_BH + BH_BYTES, MP_OFF_B16L = MP_OFF_B16H + B16_BYTES / 2;
constexpr int MP_THREADS = 384, MP_XF_THREADS = 64, MP_DRAIN_WARPS = 8;
constexpr int MP_SMEM = MP_STAGES * MP_STAGE + MP_DRAIN_WARPS * EPI16_STAGE_BYTES + 1024 + 256;

template <int EPI>
__device__ __forceinline__ void gemm3x_epilogue16(const Tc3Params& p, float (&sum)[128], float* stage, int lane, int row0, int cbase) {
    const int tr = lane >> 2, tc = 4 * (lane & 3);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j] = sum[cc * 16 + j]; sum[cc * 16 + j] = 0.f; }
        epi16_stage_rows(stage, lane, v);
        const int col = cbase + cc * 16 + tc;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((EPI == EP_BIAS || EPI == EP_BIAS_GELU || EPI == EP_BIAS_ADD) && p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + 8 * i + tr;
            if (row >= p.M) continue;
            const float4 a = epi16_read_t(stage, lane, i);
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI == EP_BIAS_ADD || EPI == EP_GELU_BWD) e = *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col);
            float4 o, o2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI == EP_STORE) o = a;
            else if (EPI == EP_BIAS) o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
            else if (EPI == EP_BIAS_GELU) {
                o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
                o2 = make_float4(te_gelu(o.x), te_gelu(o.y), te_gelu(o.z), te_gelu(o.w));
            } else if (EPI == EP_BIAS_ADD) {
                o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
                o2 = make_float4(e.x + o.x, e.y + o.y, e.z + o.z, e.w + o.w);
            } else {
                o = make_float4(a.x * te_gelu_grad(e.x), a.y * te_gelu_grad(e.y), a.z * te_gelu_grad(e.z), a.w * te_gelu_grad(e.w));
            }
            *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = o;
            if (EPI == EP_BIAS_GELU || EPI == EP_BIAS_ADD) *reinterpret_cast<float4*>(p.C2 + (long long)row * p.ldc2 + col) = o2;
        }
    }
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MP_THREADS, 1)
te_tc_gemm3mp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                     const __grid_constant__ CUtensorMap tmB16h, const __grid_constant__ CUtensorMap tmB16l, const Tc3Params p,
                     int tiles_m, int tiles_n) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + MP_STAGES * MP_STAGE + MP_DRAIN_WARPS * EPI16_STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto ready_bar = [&](int s) { return bars + 8u * (MP_STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * MP_STAGES + s); };
    auto accfull_bar = [&](int b) { return bars + 8u * (3 * MP_STAGES + b); };
    auto accfree_bar = [&](int b) { return bars + 8u * (3 * MP_STAGES + 2 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + MP_STAGES * MP_STAGE + MP_DRAIN_WARPS * EPI16_STAGE_BYTES + 8 * (3 * MP_STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int ntiles = tiles_m * tiles_n;
    const int kb = p.K / BK;
    const int nchunks = (kb + CHUNK - 1) / CHUNK;
    constexpr uint32_t TMEM_COLS = 512u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB16h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB16l) : "memory");
        for (int s = 0; s < MP_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), 2u * (MP_XF_THREADS / 32));
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accfull_bar(b), 1);
            mbar_init(accfree_bar(b), 2u * MP_DRAIN_WARPS);
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
            uint32_t it = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                const int m0 = ((t / tiles_n) * 2 + (int)rank) * BM, n0 = (t % tiles_n) * BN + (int)rank * (BN / 2);
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const int s = (int)(it % MP_STAGES);
                    const uint32_t ph = (it / MP_STAGES) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), A_BYTES + BH_BYTES + B16_BYTES);
                    const uint32_t sa = smem_base + s * MP_STAGE;
                    tma_load_2d(sa + MP_OFF_AH, &tmA, full_bar(s), kk * BK, m0);                    // raw A -> A_hi slot
                    tma_load_2d(sa + MP_OFF_BH, &tmBh, full_bar(s), kk * BK, n0);
                    tma_load_2d(sa + MP_OFF_B16H, &tmB16h, full_bar(s), kk * BK, n0);
                    tma_load_2d(sa + MP_OFF_B16L, &tmB16l, full_bar(s), kk * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            uint32_t it = 0, gc = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const bool chunk_start = (kk % CHUNK) == 0;
                    const uint32_t b = gc & 1u;
                    if (chunk_start && gc >= 2) {                       // accumulator b drained (chunk gc-2) in BOTH CTAs
                        mbar_wait_cluster(accfree_bar(b), ((gc >> 1) & 1u) ^ 1u);
                        tcgen05_fence_after();
                    }
                    const int s = (int)(it % MP_STAGES);
                    const uint32_t ph = (it / MP_STAGES) & 1u;
                    mbar_wait_cluster(ready_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + s * MP_STAGE;
                    const uint64_t ah = make_smem_desc(sa + MP_OFF_AH), bh = make_smem_desc(sa + MP_OFF_BH);
                    const uint64_t a16h = make_smem_desc_sw64(sa + MP_OFF_A16H), a16l = make_smem_desc_sw64(sa + MP_OFF_A16L);
                    const uint64_t b16h = make_smem_desc_sw64(sa + MP_OFF_B16H), b16l = make_smem_desc_sw64(sa + MP_OFF_B16L);
                    const uint32_t d = tmem_base + b * (uint32_t)BN;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t o = (uint64_t)(2 * k);
                        umma2_bf16(d, a16l + o, b16h + o, kIdesc2Bf16, (chunk_start && k == 0) ? 0u : 1u);
                        umma2_bf16(d, a16h + o, b16l + o, kIdesc2Bf16, 1u);
                    }
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) umma2_tf32(d, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), kIdesc2, 1u);
                    umma2_commit_both(empty_bar(s));
                    if ((kk % CHUNK) == CHUNK - 1 || kk == kb - 1) { umma2_commit_both(accfull_bar(b)); ++gc; }
                }
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // ---- A tile: hi (tf32, in place) + bf16(hi) + bf16(lo): warps 2-3 of both CTAs ----
        const int et = threadIdx.x - 64;                                // 0..63
        uint32_t it = 0;
        for (int t = cluster_id; t < ntiles; t += nclusters) {
            for (int kk = 0; kk < kb; ++kk, ++it) {
                const int s = (int)(it % MP_STAGES);
                const uint32_t ph = (it / MP_STAGES) & 1u;
                mbar_wait(full_bar(s), ph);
                uint8_t* st8 = smem_al + s * MP_STAGE;
                float4* a4 = reinterpret_cast<float4*>(st8 + MP_OFF_AH);
#pragma unroll 4
                for (int i = 0; i < A_BYTES / 16 / MP_XF_THREADS; ++i) {
                    const int idx = et + i * MP_XF_THREADS;
                    const int r = idx >> 3, lc = (idx & 7) ^ (r & 7);
                    const float4 v = a4[idx];
                    float4 h;
                    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
                    a4[idx] = h;
                    const __nv_bfloat162 h01 = __floats2bfloat162_rn(h.x, h.y), h23 = __floats2bfloat162_rn(h.z, h.w);
                    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - h.x, v.y - h.y), l23 = __floats2bfloat162_rn(v.z - h.z, v.w - h.w);
                    const uint32_t off = (uint32_t)r * 64u + ((uint32_t)((lc >> 1) ^ ((r >> 1) & 3)) << 4) + ((uint32_t)(lc & 1) << 3);
                    uint2 hb, lb;
                    hb.x = *reinterpret_cast<const uint32_t*>(&h01); hb.y = *reinterpret_cast<const uint32_t*>(&h23);
                    lb.x = *reinterpret_cast<const uint32_t*>(&l01); lb.y = *reinterpret_cast<const uint32_t*>(&l23);
                    *reinterpret_cast<uint2*>(st8 + MP_OFF_A16H + off) = hb;
                    *reinterpret_cast<uint2*>(st8 + MP_OFF_A16L + off) = lb;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_rank0(ready_bar(s)));
            }
        }
    } else {
        // ---- chunk drain + epilogue: warps 4..11 (lane quarter = warp & 3, column half = (warp - 4) / 4) ----
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
        float* stage = reinterpret_cast<float*>(smem_al + MP_STAGES * MP_STAGE + (warp - 4) * EPI16_STAGE_BYTES);
        float sum[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) sum[j] = 0.f;
        uint32_t gc = 0;
        for (int t = cluster_id; t < ntiles; t += nclusters) {
            const int m0 = ((t / tiles_n) * 2 + (int)rank) * BM, n0 = (t % tiles_n) * BN;
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const uint32_t b = gc & 1u;
                mbar_wait(accfull_bar(b), (gc >> 1) & 1u);
                tcgen05_fence_after();
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    uint32_t v[16];
                    tmem_ld16(tlane + b * (uint32_t)BN + (uint32_t)(cc * 16), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[cc * 16 + j] += __uint_as_float(v[j]);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_rank0(accfree_bar(b)));
            }
            gemm3x_epilogue16<EPI>(p, sum, stage, lane, m0 + q * 32, n0 + half * 128);      // also zeroes the sums
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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
        gemm3x_epilogue<EPI>(p, sum, reinterpret_cast<float*>(smem_al + (warp - 2) * EPI_STAGE_BYTES), lane, m0 + q * 32,
                             n0 + half * 128);
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host ---------------------------------------------------------------------------------------------------
int g_pair_linear = -1;                 // 3xTF32 Linear GEMMs as CTA pairs (TE_B200_LINEAR_2CTA=1 / te_set_option)
bool use_pair_linear() {
    if (g_pair_linear < 0) {
        const char* e = getenv("TE_B200_LINEAR_2CTA");
        g_pair_linear = (e && e[0] == '1') ? 1 : 0;
    }
    return g_pair_linear == 1;
}
template <int EPI>
int launch3(const float* A, long long lda, const float* Bh, const float* Bl, const Tc3Params& p, cudaStream_t st) {
    CUtensorMap tmA, tmBh, tmBl;
    if (!make_map(&tmA, A, p.M, p.K, lda, BM) || !make_map(&tmBh, Bh, p.N, p.K, p.K, BN) ||
        !make_map(&tmBl, Bl, p.N, p.K, p.K, BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;          // per-device attribute: one bit per device
    if (!smem_optin(te_tc_gemm3x_kernel<EPI>, SMEM3_BYTES, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
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
    static unsigned long long optin = 0;          // per-device attribute: one bit per device
    if (!smem_optin(te_tc_gemm3x2_kernel<EPI>, SMEM3P_BYTES, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
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

// bf16 [rows, cols] K-major, 64-byte rows, SWIZZLE_64B (box = 32 elements x box_rows)
bool make_map_bf16_sw64(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_mixed_linear = -1;                // forward Linears with bf16 correction terms (TE_B200_LINEAR_MIXED=0/1, te_set_option)
bool use_mixed_linear() {
    if (g_mixed_linear < 0) {
        const char* e = getenv("TE_B200_LINEAR_MIXED");
        g_mixed_linear = (e && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
    }
    return g_mixed_linear >= 1;
}

template <int EPI>
int launch3m(const float* A, long long lda, const float* Bh, const void* B16h, const void* B16l, const Tc3Params& p, cudaStream_t st) {
    CUtensorMap tmA, tmBh, tm16h, tm16l;
    if (!make_map(&tmA, A, p.M, p.K, lda, BM) || !make_map(&tmBh, Bh, p.N, p.K, p.K, BN) ||
        !make_map_bf16_sw64(&tm16h, B16h, p.N, p.K, p.K, BN) || !make_map_bf16_sw64(&tm16l, B16l, p.N, p.K, p.K, BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (mixed)");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;
    if (!smem_optin(te_tc_gemm3m_kernel<EPI>, SMEM3_BYTES, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    dim3 grid(p.N / BN, (unsigned)((p.M + BM - 1) / BM));
    te_tc_gemm3m_kernel<EPI><<<grid, NUM_THREADS3, SMEM3_BYTES, st>>>(tmA, tmBh, tm16h, tm16l, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}
int dispatch3m(int epi, const float* A, long long lda, const float* Bh, const void* B16h, const void* B16l, const Tc3Params& p,
               cudaStream_t st) {
    switch (epi) {
        case TE_TC_EPI_STORE: return launch3m<EP_STORE>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_BIAS: return launch3m<EP_BIAS>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_BIAS_GELU: return launch3m<EP_BIAS_GELU>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_BIAS_ADD: return launch3m<EP_BIAS_ADD>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_GELU_BWD: return launch3m<EP_GELU_BWD>(A, lda, Bh, B16h, B16l, p, st);
    }
    te_set_last_error("te_gemm_tc: unsupported epilogue");
    return TE_ERR_UNSUPPORTED;
}

int mp_sm_pairs() {
    static int cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int& c = cache[dev & 63];
    if (c == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        c = n / 2;
    }
    return c;
}
template <int EPI>
int launch3mp(const float* A, long long lda, const float* Bh, const void* B16h, const void* B16l, const Tc3Params& p, cudaStream_t st) {
    CUtensorMap tmA, tmBh, tm16h, tm16l;
    if (!make_map(&tmA, A, p.M, p.K, lda, BM) || !make_map(&tmBh, Bh, p.N, p.K, p.K, BN / 2) ||
        !make_map_bf16_sw64(&tm16h, B16h, p.N, p.K, p.K, BN / 2) || !make_map_bf16_sw64(&tm16l, B16l, p.N, p.K, p.K, BN / 2)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (mixed pair)");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;
    if (!smem_optin(te_tc_gemm3mp_kernel<EPI>, MP_SMEM, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    const int mt = (p.M + BM - 1) / BM;
    const int tiles_m = (mt + 1) / 2, tiles_n = p.N / BN;
    int pairs = mp_sm_pairs();
    if (pairs <= 0) { te_set_last_error("te_gemm_tc: cannot query the SM count"); return TE_ERR_CUDA; }
    if (pairs > tiles_m * tiles_n) pairs = tiles_m * tiles_n;
    te_tc_gemm3mp_kernel<EPI><<<dim3(2u * (unsigned)pairs), MP_THREADS, MP_SMEM, st>>>(tmA, tmBh, tm16h, tm16l, p, tiles_m, tiles_n);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}
int dispatch3mp(int epi, const float* A, long long lda, const float* Bh, const void* B16h, const void* B16l, const Tc3Params& p,
                cudaStream_t st) {
    switch (epi) {
        case TE_TC_EPI_STORE: return launch3mp<EP_STORE>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_BIAS: return launch3mp<EP_BIAS>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_BIAS_GELU: return launch3mp<EP_BIAS_GELU>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_BIAS_ADD: return launch3mp<EP_BIAS_ADD>(A, lda, Bh, B16h, B16l, p, st);
        case TE_TC_EPI_GELU_BWD: return launch3mp<EP_GELU_BWD>(A, lda, Bh, B16h, B16l, p, st);
    }
    te_set_last_error("te_gemm_tc: unsupported epilogue");
    return TE_ERR_UNSUPPORTED;
}

}  // namespace

void te_tc_set_pair_linear(int on) { g_pair_linear = on ? 1 : 0; }
void te_tc_set_mixed_linear(int on) { g_mixed_linear = (on == 2) ? 2 : (on ? 1 : 0); }

bool te_tc_gemm3x_supported(long long rows, int K, int N, long long lda) {
    return rows > 0 && rows < (1LL << 31) && K % BK == 0 && N % BN == 0 && lda % 4 == 0 && get_encode() != nullptr;
}

// y[rows,out] = x[rows,in] W^T (+ epilogue)   — fp32-grade (3xTF32) on tcgen05
int te_tc_linear_fwd(const float* x, long long ldx, const float* derived, int in_features, int out_features,
                     const float* bias, float* y, float* y2, const float* e0, long long rows, int epi, cudaStream_t st) {
    const long long n = (long long)in_features * out_features;
    Tc3Params p;
    p.M = (int)rows; p.N = out_features; p.K = in_features; p.bias = bias; p.E = e0; p.lde = out_features;
    p.C = y; p.ldc = out_features; p.C2 = y2; p.ldc2 = out_features;
    if (use_mixed_linear()) {           // main term TF32, correction terms bf16: [bf16(W_hi) | bf16(W_lo)] live at derived + 10 n
        const __nv_bfloat16* w16 = reinterpret_cast<const __nv_bfloat16*>(derived + 10 * n);
        if (g_mixed_linear == 2) return dispatch3mp(epi, x, ldx, derived + 4 * n, w16, w16 + n, p, st);     // persistent CTA pair
        return dispatch3m(epi, x, ldx, derived + 4 * n, w16, w16 + n, p, st);
    }
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

