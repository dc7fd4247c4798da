// Attention-shaped contractions (N x N and token-reduced N x d) and the dense rollout product on tcgen05, 3xTF32.
// Shared PTX wrappers: te_tc_common.cuh.
#include "te_tc_common.cuh"

namespace {

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

// 2^x for x <= 0 in one MUFU instruction (ex2.approx.ftz: 2^-22 relative; results below the normal range flush to 0)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AtParams {
    int N, H, dh, ld_out;            // tokens, heads, head_dim, row stride of out / E
    const float* E; float* out; float alpha;
};

// Epilogue of the N x N attention kernels (shared by the one-tile-per-CTA and the persistent kernel): the finished 128 x ncols
// accumulator sits at TMEM address tlane (this warp's lane quarter q); rows m0 + q*32 .. of head bh, key columns n0 .. n0+ncols.
template <int EPI, bool SP>
__device__ __forceinline__ void attn_nn_epilogue(const AtParams& p, uint32_t tlane, float* stage, int lane, int q, int bh, int m0,
                                                 int n0, int ncols) {
        const int nchunks = (ncols + 31) / 32;
        const int tr = lane >> 3, tc = 4 * (lane & 7);
        if (EPI == AT_SOFTMAX) {
            // softmax(alpha * A B^T) over the key axis, fused: every thread owns one query row whose N <= 256 scores sit
            // in its TMEM lane, so the row maximum, the sum of exponentials and the normalised probabilities come from
            // three passes over TMEM (the exponentials are written back with tcgen05.st) — the scores never travel to HBM (attn = dots.softmax(dim=-1), ViT_LRP.py:139-141)
            float mx = -INFINITY;
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t acc[32];
                tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
                tmem_ld_wait();
                if (cc * 32 + 32 <= ncols) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, p.alpha * __uint_as_float(acc[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (cc * 32 + j < ncols) mx = fmaxf(mx, p.alpha * __uint_as_float(acc[j]));
                }
            }
            // pass 2: e = exp(score - max) = 2^(alpha log2e * acc - max log2e): one fma (single rounding of the argument) + ex2.approx
            // (2^-22 relative) per element, summed, and written back over the scores in TMEM
            float sum = 0.f;
            const float a2 = p.alpha * 1.4426950408889634f, m2 = -mx * 1.4426950408889634f;
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t acc[32];
                tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
                tmem_ld_wait();
                if (cc * 32 + 32 <= ncols) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float e = ex2_approx(fmaf(a2, __uint_as_float(acc[j]), m2));
                        sum += e;
                        acc[j] = __float_as_uint(e);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float e = (cc * 32 + j < ncols) ? ex2_approx(fmaf(a2, __uint_as_float(acc[j]), m2)) : 0.f;
                        sum += e;
                        acc[j] = __float_as_uint(e);
                    }
                }
                tmem_st32(tlane + (uint32_t)(cc * 32), acc);
            }
            tmem_st_wait();
            const float inv = 1.0f / sum;
            // pass 3: normalise in the row layout, store in the transposed (coalesced) layout of epi_read_t
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                uint32_t acc[32];
                tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) * inv);   // padding holds e = 0
                epi_stage_rows(stage, lane, acc);
                const int col = cc * 32 + tc;
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2) {
                    const int r = m0 + q * 32 + 4 * i2 + tr;
                    if (r < p.N && col < ncols)
                        *reinterpret_cast<float4*>(p.out + ((long long)bh * p.N + r) * p.ld_out + n0 + col) = epi_read_t(stage, lane, i2);
                }
            }
        } else {
        // E and out are accessed in the transposed layout of epi_read_t (4 rows x 128 B per warp instruction); the loads of
        // E are issued before the TMEM read completes.  A float4 whose tail lies in the row padding is memory-safe
        // (ld_out % 4 == 0); the padding columns [ncols, ...) of the last float4 are written as zeros.
#pragma unroll 1
        for (int cc = 0; cc < nchunks; ++cc) {
            uint32_t acc[32];
            tmem_ld32(tlane + (uint32_t)(cc * 32), acc);
            const int col = cc * 32 + tc;
            float4 e4[8];
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
                const int r = m0 + q * 32 + 4 * i2 + tr;
                e4[i2] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (EPI != AT_STORE && r < p.N && col < ncols)
                    e4[i2] = __ldcs(reinterpret_cast<const float4*>(p.E + ((long long)bh * p.N + r) * p.ld_out + n0 + col));
            }
            tmem_ld_wait();
            epi_stage_rows(stage, lane, acc);
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
                const int r = m0 + q * 32 + 4 * i2 + tr;
                if (r >= p.N || col >= ncols) continue;
                const float4 a4 = epi_read_t(stage, lane, i2);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float e[4] = {e4[i2].x, e4[i2].y, e4[i2].z, e4[i2].w};
                float o[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    // SP: both raw operands were truncated to TF32 by the tensor core (mean shrink 3.4e-4 each): compensated
                    const float av = p.alpha * a[u] * (SP ? 1.00068f : 1.0f);
                    float v;
                    if (EPI == AT_STORE) v = av;
                    else if (EPI == AT_MUL) v = av * e[u];
                    else v = te_sd_fast(e[u], av);
                    o[u] = (col + u < ncols) ? v : 0.f;                                  // zero the row padding
                }
                *reinterpret_cast<float4*>(p.out + ((long long)bh * p.N + r) * p.ld_out + n0 + col) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        }   // EPI != AT_SOFTMAX
}

// SP (single pass): the raw fp32 operands go straight from TMA to one TF32 MMA per k-step (no hi/lo split); both k-blocks
// of the head dimension are resident at once (the lo regions hold the second one).  For the contractions whose result is
// relevance or a gradient — G = dctx V^T under TE_FLAG_BACKWARD_TF32, attn_cam = P * (S V^T) / 2 under
// TE_FLAG_RELPROP_TF32 — never for a safe_divide denominator (Q K^T stays 3xTF32).
template <int EPI, bool SP = false>
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
        mbar_init(xf_bar, SP ? 1 : XF_THREADS / 32);           // SP: TMA barrier of the second k-block
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
            if (SP) {
                for (int k = 0; k < kb && k < 2; ++k) {                          // dh <= 64: at most two k-blocks, both resident
                    const uint32_t bar = k ? xf_bar : full_bar;
                    mbar_arrive_expect_tx(bar, (uint32_t)(A_BYTES + B_BYTES));
                    tma_load_2d(smem_base + (k ? OFF_AL : OFF_AH), &tmA, bar, h * p.dh + k * BK, b * p.N + m0);
                    tma_load_2d(smem_base + (k ? OFF_BL : OFF_BH), &tmB, bar, h * p.dh + k * BK, b * p.N + n0);
                }
            }
            for (int k = 0; k < (SP ? 0 : kb); ++k) {
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
                if (SP) {
                    mbar_wait(kk ? xf_bar : full_bar, 0u);
                    tcgen05_fence_after();
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)
                        umma_tf32(tmem_base, (kk ? al : ah) + (uint64_t)(2 * k), (kk ? bl : bh_) + (uint64_t)(2 * k), kIdesc,
                                  (kk == 0 && k == 0) ? 0u : 1u);
                    continue;
                }
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
        if (EPI == AT_MUL || EPI == AT_SD) {
            // the epilogue operand E (attention probabilities / attn_cam) is streamed from HBM exactly once: pull this tile's rows
            // into L2 now, while the operands are loaded and the MMAs run, so that the epilogue's loads are L2 hits
            const int pr = m0 + et;                                  // one row per thread
            if (pr < p.N) {
                const float* e = p.E + ((long long)bh * p.N + pr) * p.ld_out + n0;
                const int nc = min(p.N - n0, BN);
                for (int j = 0; j < nc; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(e + j));
            }
        }
        // split A (16 KiB) and B (32 KiB) of every k-block: hi in place, lo to the *_lo regions (same swizzled offsets)
        for (int kk = 0; kk < (SP ? 0 : kb); ++kk) {
            mbar_wait_sleep(full_bar, (uint32_t)(kk & 1));
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
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const int ncols = min(p.N - n0, BN);              // valid key columns of this tile
        // per-warp staging buffer of the coalesced epilogue: the operand buffer is idle once the accumulator is complete
        float* stage = reinterpret_cast<float*>(smem_al + (warp - 2) * EPI_STAGE_BYTES);
        mbar_wait_sleep(accum_bar, 0);
        tcgen05_fence_after();
        attn_nn_epilogue<EPI, SP>(p, tlane, stage, lane, q, bh, m0, n0, ncols);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Persistent form of the fp32-grade (3xTF32) N x N attention kernel for N <= 224 (ViT / DeiT): the one-tile-per-CTA kernel
// above runs the chain TMA -> hi/lo split -> MMA -> epilogue of a tile start to end (two CTAs per SM overlap a little of it) and
// sits 4-8x above the HBM bound of its outputs.  Here one CTA per SM loops over (batch*head, row tile) work items with dedicated
// warps per role, a 2-stage operand ring (one 32-element k-block of A and of the 224-row B tile per stage, hi + lo: 88 KiB) that
// the producer fills ahead across tile boundaries, and the 256-column accumulator DOUBLE-BUFFERED in TMEM, so the epilogue of
// item i (scale / * E / safe_divide / the three-pass softmax) overlaps the split and the MMAs of item i+1.
//   warp 0 TMA producer · warp 1 MMA issuer · warps 2-5 hi/lo split · warps 6-9 epilogue (lane quarter = warp & 3)
//   full[s] TMA bytes · xf[s] split done (4 warps) · empty[s] tcgen05.commit · accfull[b] last commit of an item ·
//   accfree[b] accumulator read out (4 warps)
// =====================================================================================================================
constexpr int PN_BN = 224;                                         // B rows (keys) per tile: N <= 224
constexpr int PN_B_BYTES = PN_BN * BK * 4;                         // 28 KiB
constexpr int PN_STAGE = 2 * A_BYTES + 2 * PN_B_BYTES;             // 88 KiB
constexpr int PN_STAGES = 2;
constexpr int PN_THREADS = 320;
constexpr int PN_SMEM = PN_STAGES * PN_STAGE + 4 * EPI_STAGE_BYTES + 1024 + 256;
constexpr uint32_t kIdescN224 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PN_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

template <int EPI>
__global__ void __launch_bounds__(PN_THREADS, 1)
te_tc_attn_nn_p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AtParams p,
                       int items) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr uint32_t OFF_AH = 0, OFF_AL = A_BYTES, OFF_BH = 2 * A_BYTES, OFF_BL = 2 * A_BYTES + PN_B_BYTES;
    const uint32_t bars = smem_base + PN_STAGES * PN_STAGE + 4 * EPI_STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xf_bar = [&](int s) { return bars + 8u * (2 + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (4 + s); };
    auto accfull_bar = [&](int b) { return bars + 8u * (6 + b); };
    auto accfree_bar = [&](int b) { return bars + 8u * (8 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + PN_STAGES * PN_STAGE + 4 * EPI_STAGE_BYTES + 8 * 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb = p.dh / BK;
    const int mtiles = (p.N + BM - 1) / BM;
    constexpr uint32_t TMEM_COLS = 512u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < 2; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(xf_bar(s), 4);
            mbar_init(empty_bar(s), 1);
            mbar_init(accfull_bar(s), 1);
            mbar_init(accfree_bar(s), 4);
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
            uint32_t it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x) {
                const int bh = w / mtiles, m0 = (w % mtiles) * BM, b = bh / p.H, h = bh % p.H;
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const int s = (int)(it & 1u);
                    const uint32_t ph = (it >> 1) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), (uint32_t)(A_BYTES + PN_B_BYTES));
                    const uint32_t sa = smem_base + s * PN_STAGE;
                    tma_load_2d(sa + OFF_AH, &tmA, full_bar(s), h * p.dh + kk * BK, b * p.N + m0);
                    tma_load_2d(sa + OFF_BH, &tmB, full_bar(s), h * p.dh + kk * BK, b * p.N);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, wi = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++wi) {
                const uint32_t ab = wi & 1u;
                if (wi >= 2) {                                   // accumulator buffer ab read out by the epilogue of item wi-2
                    mbar_wait(accfree_bar(ab), ((wi >> 1) & 1u) ^ 1u);
                    tcgen05_fence_after();
                }
                const uint32_t d = tmem_base + ab * 256u;
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const int s = (int)(it & 1u);
                    const uint32_t ph = (it >> 1) & 1u;
                    mbar_wait(xf_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + s * PN_STAGE;
                    const uint64_t ah = make_smem_desc(sa + OFF_AH), al = make_smem_desc(sa + OFF_AL);
                    const uint64_t bh_ = make_smem_desc(sa + OFF_BH), bl = make_smem_desc(sa + OFF_BL);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t o = (uint64_t)(2 * k);
                        umma_tf32(d, al + o, bh_ + o, kIdescN224, (kk == 0 && k == 0) ? 0u : 1u);
                        umma_tf32(d, ah + o, bl + o, kIdescN224, 1u);
                        umma_tf32(d, ah + o, bh_ + o, kIdescN224, 1u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(accfull_bar(ab));
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ---- hi / lo split of every staged k-block: hi in place, lo into the *_lo regions (same swizzled offsets) ----
        const int et = threadIdx.x - 64;
        uint32_t it = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            for (int kk = 0; kk < kb; ++kk, ++it) {
                const int s = (int)(it & 1u);
                const uint32_t ph = (it >> 1) & 1u;
                mbar_wait(full_bar(s), ph);
                uint8_t* st8 = smem_al + s * PN_STAGE;
                float4* a4 = reinterpret_cast<float4*>(st8 + OFF_AH);
                float4* l4 = reinterpret_cast<float4*>(st8 + OFF_AL);
#pragma unroll
                for (int i = et; i < A_BYTES / 16; i += XF_THREADS) {
                    const float4 v = a4[i];
                    float4 hh, l;
                    hh.x = to_tf32(v.x); hh.y = to_tf32(v.y); hh.z = to_tf32(v.z); hh.w = to_tf32(v.w);
                    l.x = to_tf32(v.x - hh.x); l.y = to_tf32(v.y - hh.y); l.z = to_tf32(v.z - hh.z); l.w = to_tf32(v.w - hh.w);
                    a4[i] = hh; l4[i] = l;
                }
                float4* b4 = reinterpret_cast<float4*>(st8 + OFF_BH);
                float4* m4 = reinterpret_cast<float4*>(st8 + OFF_BL);
#pragma unroll 2
                for (int i = et; i < PN_B_BYTES / 16; i += XF_THREADS) {
                    const float4 v = b4[i];
                    float4 hh, l;
                    hh.x = to_tf32(v.x); hh.y = to_tf32(v.y); hh.z = to_tf32(v.z); hh.w = to_tf32(v.w);
                    l.x = to_tf32(v.x - hh.x); l.y = to_tf32(v.y - hh.y); l.z = to_tf32(v.z - hh.z); l.w = to_tf32(v.w - hh.w);
                    b4[i] = hh; m4[i] = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(xf_bar(s));
            }
        }
    } else {
        // ---- epilogue: warps 6..9 ----
        const int q = warp & 3;
        float* stage = reinterpret_cast<float*>(smem_al + PN_STAGES * PN_STAGE + (warp - 6) * EPI_STAGE_BYTES);
        uint32_t wi = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++wi) {
            const int bh = w / mtiles, m0 = (w % mtiles) * BM;
            const uint32_t ab = wi & 1u;
            mbar_wait(accfull_bar(ab), (wi >> 1) & 1u);
            tcgen05_fence_after();
            const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256u;
            attn_nn_epilogue<EPI, false>(p, tlane, stage, lane, q, bh, m0, 0, p.N);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accfree_bar(ab));
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
// SP (single pass): raw fp32 operands straight from TMA to one TF32 MMA per k-step — no hi/lo split, no transform warps in
// the chain, half the stage, twice the ring depth.  Used for the activation-gradient contractions (attn^T dctx, dS k,
// dS^T q) under TE_FLAG_BACKWARD_TF32: the gradients enter the result linearly (relu(G * cam)), never a denominator.
template <int NB, bool SP = false> struct NkCfg {
    static constexpr int BN = NB * 32;
    static constexpr int B_BYTES_ = BN * BK * 4;                   // 4 KiB per block
    static constexpr int STAGE = SP ? (NK_A + B_BYTES_) : (2 * NK_A + 2 * B_BYTES_);
    // attention shape (NB = 2): 2 stages of 48 KiB (SP: 4 of 24 KiB) so that TWO CTAs share an SM — a CTA's whole
    // reduction is only 7 k-blocks, and its prologue / epilogue then overlap the other CTA's main loop
    static constexpr int STAGES = SP ? 4 : 2;
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

template <int AMN, int EPI, int NB, bool SP = false>
__global__ void __launch_bounds__(NUM_THREADS, NkCfg<NB, SP>::MIN_CTAS)
te_tc_attn_nk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const NkParams p) {
    using C = NkCfg<NB, SP>;
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
    constexpr uint32_t OFF_AL = NK_A, OFF_BH = SP ? NK_A : 2 * NK_A, OFF_BL = 2 * NK_A + NK_B;
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
                mbar_wait(SP ? full_bar(s) : xf_bar(s), ph);
                tcgen05_fence_after();
                const uint32_t sa = smem_base + s * NK_STAGE;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    uint64_t ah, al;
                    if (SP) {                     // one TF32 MMA per k-step on the raw operands
                        ah = (AMN == 0) ? make_smem_desc(sa) + (uint64_t)(2 * k) : make_smem_desc_mn(sa + k * 1024, 4096);
                        umma_tf32(tmem_base, ah, make_smem_desc_mn(sa + OFF_BH + k * 1024, 4096), idesc, (it == 0 && k == 0) ? 0u : 1u);
                        continue;
                    }
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
        for (int it = 0; it < (SP ? 0 : kb); ++it) {
            const int s = it % NK_STAGES;
            const uint32_t ph = (it / NK_STAGES) & 1u;
            mbar_wait_sleep(full_bar(s), ph);
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
        mbar_wait_sleep(accum_bar, 0);
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
                            // SP: both raw operands were truncated to TF32 by the tensor core (mean shrink 3.4e-4 each)
                            float v = p.alpha * a * (SP ? 1.00068f : 1.0f);
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

}  // namespace

bool te_tc_attn_supported(int N, int dh, long long lda, long long ldb, int ld_out) {
    return N >= 1 && (dh == 32 || dh == 64) && lda % 4 == 0 && ldb % 4 == 0 && ld_out % 4 == 0 && get_encode() != nullptr;
}

namespace {
// persistent N x N kernel for N <= 224: opt-in (TE_B200_ATTN_PERSISTENT=1 / te_set_option).  Measured SLOWER than two resident
// one-tile CTAs per SM (0.79 / 0.81 ms against 0.54 / 0.57 ms per launch for the softmax / S1 epilogues, step 1484 vs 1524
// expl/s): with one CTA per SM only 4 epilogue warps work on an epilogue of ~37 k cycles per item (three TMEM passes with
// expf, or HBM-latency-bound E loads), where the resident pair has 8; kept as the base of a version with 8-12 epilogue warps.
int g_attn_persistent = -1;
bool use_attn_persistent() {
    if (g_attn_persistent < 0) {
        const char* e = getenv("TE_B200_ATTN_PERSISTENT");
        g_attn_persistent = (e && e[0] == '1') ? 1 : 0;
    }
    return g_attn_persistent == 1;
}
int attn_sm_count() {
    static int cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int& c = cache[dev & 63];
    if (c == 0 && cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return c;
}

template <int EPI>
int launch_attn_p(const float* A, long long lda, const float* B, long long ldb, long long total_rows, const AtParams& p,
                  int batch, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, A, total_rows, (long long)p.H * p.dh, lda, BM) || !make_map(&tmB, B, total_rows, (long long)p.H * p.dh, ldb, PN_BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (attention, persistent)");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;
    if (!smem_optin(te_tc_attn_nn_p_kernel<EPI>, PN_SMEM, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    const int items = batch * p.H * ((p.N + BM - 1) / BM);
    int grid = attn_sm_count();
    if (grid <= 0) { te_set_last_error("te_gemm_tc: cannot query the SM count"); return TE_ERR_CUDA; }
    if (grid > items) grid = items;
    te_tc_attn_nn_p_kernel<EPI><<<grid, PN_THREADS, PN_SMEM, st>>>(tmA, tmB, p, items);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

template <int EPI, bool SP = false>
int launch_attn(const float* A, long long lda, const float* B, long long ldb, long long total_rows, const AtParams& p,
                int batch, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, A, total_rows, (long long)p.H * p.dh, lda, BM) || !make_map(&tmB, B, total_rows, (long long)p.H * p.dh, ldb, BN)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (attention)");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;          // per-device attribute: one bit per device
    if (!smem_optin(te_tc_attn_nn_kernel<EPI, SP>, AT_SMEM, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    dim3 grid((p.N + BM - 1) / BM, batch * p.H, (p.N + BN - 1) / BN);
    if (grid.y > 65535) { te_set_last_error("te_gemm_tc: batch*heads too large for one launch"); return TE_ERR_ARG; }
    te_tc_attn_nn_kernel<EPI, SP><<<grid, NUM_THREADS, AT_SMEM, st>>>(tmA, tmB, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}
}  // namespace

// out[b,h,i,j] = epi(alpha * sum_d A[b*N+i, h*dh+d] * B[b*N+j, h*dh+d]);  out / E are [batch,H,N,ld_out]
int te_tc_attn_nn(const float* A, long long lda, const float* B, long long ldb, int batch, int H, int N, int dh,
                  float* out, int ld_out, const float* E, float alpha, int epi, cudaStream_t st, bool single_pass) {
    AtParams p;
    p.N = N; p.H = H; p.dh = dh; p.ld_out = ld_out; p.E = E; p.out = out; p.alpha = alpha;
    const long long rows = (long long)batch * N;
    if (single_pass && epi == TE_TC_ATTN_STORE) return launch_attn<AT_STORE, true>(A, lda, B, ldb, rows, p, batch, st);
    if (single_pass && epi == TE_TC_ATTN_MUL) return launch_attn<AT_MUL, true>(A, lda, B, ldb, rows, p, batch, st);
    if (N <= PN_BN && use_attn_persistent()) {
        switch (epi) {
            case TE_TC_ATTN_STORE: return launch_attn_p<AT_STORE>(A, lda, B, ldb, rows, p, batch, st);
            case TE_TC_ATTN_MUL: return launch_attn_p<AT_MUL>(A, lda, B, ldb, rows, p, batch, st);
            case TE_TC_ATTN_SD: return launch_attn_p<AT_SD>(A, lda, B, ldb, rows, p, batch, st);
            case TE_TC_ATTN_SOFTMAX: return launch_attn_p<AT_SOFTMAX>(A, lda, B, ldb, rows, p, batch, st);
        }
    }
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

void te_tc_set_attn_persistent(int on) { g_attn_persistent = on ? 1 : 0; }

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

template <int AMN, int EPI, int NB, bool SP = false>
int launch_nk(const float* map, int NP, const float* X, long long ldx, const NkParams& p, int batch, cudaStream_t st) {
    constexpr int NK_SMEM = NkCfg<NB, SP>::SMEM;
    CUtensorMap tmA, tmB;
    // attention-shaped map [batch*H, N, NP] ; activation [batch, N, ldx]
    if (!make_map3(&tmA, map, NP, p.N, p.a_shared ? (long long)batch : (long long)batch * p.H, NP, (long long)p.N * NP, 32,
                   AMN ? 32 : BM, AMN != 0) ||
        !make_map3(&tmB, X, ldx, p.N, batch, ldx, (long long)p.N * ldx, 32, 32, true)) {
        te_set_last_error("te_gemm_tc: cuTensorMapEncodeTiled failed (attention nk)");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;          // per-device attribute: one bit per device
    if (!smem_optin(te_tc_attn_nk_kernel<AMN, EPI, NB, SP>, NK_SMEM, optin)) {
        te_set_last_error("te_gemm_tc: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    dim3 grid((p.N + BM - 1) / BM, batch * p.H);
    if (grid.y > 65535) { te_set_last_error("te_gemm_tc: batch*heads too large for one launch"); return TE_ERR_ARG; }
    te_tc_attn_nk_kernel<AMN, EPI, NB, SP><<<grid, NUM_THREADS, NK_SMEM, st>>>(tmA, tmB, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}
}  // namespace

// out[b, m, h, :] = epi(alpha * sum_k A_h[m,k] X[b,k,h,:]) ; A_h = map[b,h] (amn = 0) or its transpose (amn = 1);
// X, out, E: packed activations [batch, N, ld] (head h at columns h*64..); epi: TE_TC_ATTN_STORE / TE_TC_ATTN_MUL
int te_tc_attn_nk(const float* map, int NP, int amn, const float* X, long long ldx, int batch, int H, int N, float* out,
                  int ld_out, const float* E, float alpha, int epi, cudaStream_t st, bool single_pass) {
    NkParams p;
    p.N = N; p.H = H; p.ld_out = ld_out; p.n_out = H * 64; p.n_pad = H * 64; p.a_shared = 0; p.rowscale = nullptr;
    p.E = E; p.out = out; p.alpha = alpha;
    if (single_pass && epi == TE_TC_ATTN_STORE)
        return amn ? launch_nk<1, AT_STORE, 2, true>(map, NP, X, ldx, p, batch, st) : launch_nk<0, AT_STORE, 2, true>(map, NP, X, ldx, p, batch, st);
    if (single_pass && epi == TE_TC_ATTN_MUL)
        return amn ? launch_nk<1, AT_MUL, 2, true>(map, NP, X, ldx, p, batch, st) : launch_nk<0, AT_MUL, 2, true>(map, NP, X, ldx, p, batch, st);
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

