// fp32-grade forward Linear GEMM on tcgen05 kind::f16: block-scaled fp16 (hi, lo) split of BOTH operands, three MMAs per k-step
// (hi*hi + lo*hi + hi*lo), persistent CTA pairs, chunked accumulation.
//
// Why: the 3xTF32 forward Linears are the largest family of the step (29 %, profiles/r02_results.md) and sit on two limits at
// once: the tensor pipe (3 TF32 MMAs per k-step = 0.81 of the measured TF32 rate) and the L2 -> SM staging cap (96 KiB per
// 32-element k-block = 73 B/clk/SM against ~42).  fp16 has the SAME 11-bit significand as TF32, runs at twice the MMA rate and
// takes half the bytes, so the same error-compensated split
//     x W^T  ~  x_hi W_hi^T + x_lo W_hi^T + x_hi W_lo^T          (dropped: x_lo W_lo^T ~ 2^-22)
// costs 1.5 TF32-equivalents per k-step instead of 3 and — as a CTA pair, each CTA staging its own activation rows and HALF of
// the weight tile — 1 KiB per k element instead of 3 KiB.
//
// What fp16 lacks is range (2^-24 .. 65504): the operands are block floating point.  A block with largest magnitude m is scaled
// by its own power of two so that 2^e m lands in [2^14, 2^15):  x = 2^-e (hi + lo),  hi = fp16(2^e x),  lo = fp16(2^e x - hi).
// Powers of two make the scaling exact; hi keeps 11 bits of every element down to 2^-28 of the block maximum, hi + lo keeps 22
// bits down to 2^-17 of it and degrades gracefully below (absolute error <= 2^-40 of the block maximum).  Products hi*hi are
// exact in the fp32 accumulator (<= 2^30 each).
//   weights      one block per ROW of W [out, in] (split once, te_tc_prepare_weights): the epilogue multiplies column c by 2^-f_c
//   activations  one block per (row, 128 consecutive k): exactly one accumulation chunk (below), whose drain multiplies the
//                thread's row by 2^-e — so the producer of x can emit the split from what ONE warp sees: the LayerNorm kernel
//                (a warp iteration = 128 columns), this kernel's own GELU epilogue (a drain warp owns 32 rows x 128 columns), or
//                the stand-alone pre-pass te_tc_blocksplit_f16 (one read of x, 4 bytes written per element).
//
// Accumulation: the tensor core truncates the fp32 accumulator at every MMA, a systematic drift of ~2e-8 per MMA and element.
// As in the 3xTF32 kernels the reduction is cut into chunks (2 stages = 128 elements = 24 MMAs): chunks alternate between two
// 256-column TMEM accumulators and 8 warps drain the finished one into fp32 register sums (one fma with the block scale,
// round-to-nearest) while the MMAs of the next chunk run — ACROSS tile boundaries, so the epilogue of tile i overlaps the first
// chunks of tile i+1.
//
// Warp roles (both CTAs), FM_FWD3: warpgroup 0 = warp 0 TMA producer · warp 1 TMEM allocator + (leader) MMA issuer · warps 2-3
// idle; warpgroups 1-2 = warps 4-11 drain + epilogue.  The producer warpgroup gives registers back (setmaxnreg 40) and the drain
// warpgroups take 232 each, so the 128 register sums plus the epilogue's staging do not spill (FCfg).  The single-pass modes
// (FM_LIN1, FM_R; opt-in, see the enum) run 2 + 16 warps without the reallocation.
// Barriers: full[s] LEADER's (both CTAs' TMA bytes) · empty[s] local, multicast tcgen05.commit · accfull[b] local, multicast
// commit at the end of a chunk · accfree[b] leader, one remote arrive per drain warp of both CTAs (16).
#include <cuda_fp16.h>

#include "te_tc_common.cuh"

namespace {

enum { FP_STORE = 0, FP_BIAS = 1, FP_BIAS_GELU = 2, FP_BIAS_ADD = 3, FP_GELU_BWD = 4 };
// FM_FWD3  fp32-grade forward Linear: A = (hi, lo), B = (hi, lo), three MMAs per k-step, 256 x 256 pair tiles
// FM_LIN1  single-pass product (activation-gradient backward Linear): A = hi, B = hi, one MMA per k-step — fp16 keeps the 11
//          significant bits of TF32 (rounded to nearest instead of truncated) at twice the tensor rate and half the bytes
// FM_R     second contraction of the z+ rule, R_in = x+ (S W+) + x- (S W-): A = hi(S), B = W+^T and W-^T, two 256 x 128
//          accumulators per pair tile (128 output columns of both products), one MMA per product and k-step
enum { FM_FWD3 = 0, FM_LIN1 = 1, FM_R = 2 };

constexpr int F16_K = 64;                                  // fp16 elements per 128-byte swizzle row = one stage of K
constexpr int F16_TILE = 128 * 128;                        // 16 KiB: 128 rows x 64 fp16
constexpr int F16_CHUNK = 2;                               // stages per TMEM accumulation chunk (128 elements = one scale block)

template <int MODE> struct FCfg {
    static constexpr int NA = (MODE == FM_FWD3) ? 2 : 1;                  // A tiles per stage (hi, lo)
    static constexpr int NB = (MODE == FM_LIN1) ? 1 : 2;                  // B tiles per stage (hi, lo | W+, W-)
    static constexpr int TN = (MODE == FM_R) ? 128 : 256;                 // output columns of a pair tile
    static constexpr int BT = (TN / 2) * 128;                             // one CTA's half of a B tile: TN/2 rows x 128 bytes
    static constexpr int STAGE = NA * F16_TILE + NB * BT;                 // 64 / 32 / 32 KiB
    static constexpr int NST = (MODE == FM_FWD3) ? 3 : 5;
    // drain warps: a chunk of the single-pass modes is only 1024 tensor cycles (FWD3: 3072) and must be drained within one chunk
    // time, so 16 warps share it (64 accumulator columns per thread instead of 128)
    static constexpr int DW = (MODE == FM_FWD3) ? 8 : 16;
    static constexpr int CW = 256 / (DW / 4);                             // accumulator columns (= register sums) per drain thread
    // FWD3: 12 warps = 3 warpgroups — {TMA, MMA, 2 idle} gives registers back (setmaxnreg 40) and the two drain warpgroups take
    // them (232 each: 128 sums + the epilogue's staging without spilling; ncu of the 168-register build showed the GELU epilogue
    // waiting on local-memory loads that missed the 8 KiB L1 left beside 218 KiB of shared memory).  Single-pass modes: 2 + 16 warps.
    static constexpr int W0 = (MODE == FM_FWD3) ? 4 : 2;                  // first drain warp
    static constexpr int THREADS = 32 * (W0 + DW);
    static constexpr int NBARS = 2 * NST + 4;
    static constexpr int SMEM = NST * STAGE + DW * EPI16_STAGE_BYTES + 1024 + 8 * NBARS + 16;
    // cta_group::2, fp16 operands (a/b format 0), fp32 accumulate, M = 256, N = TN
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
};

struct F16Params {
    int M, N, K;
    int tiles_m, tiles_n;
    const float* rs; int rs_ld;           // [M, rs_ld] 2^-e of the activation blocks (row, 128 k)
    const float* cs;                      // [N] 2^-f_c of the weight rows (FM_R: of W+^T)
    const float* cs1;                     // FM_R: [N] of W-^T
    const float* bias; const float* E; long long lde;      // E: residual (BIAS_ADD) / pre-activation (GELU_BWD) / x (FM_R)
    float* C; long long ldc; float* C2; long long ldc2;
    __half* hi2; __half* lo2; float* rs2;  // SPLIT epilogue: block-scaled split of C2 [M, N] and its scales [M, N/128]
};

// sum: one accumulator row x CW columns per thread.  Global memory is accessed in the transposed layout of epi16_read_t (8 rows x
// 64 contiguous bytes per warp instruction).
// SPLIT (GELU epilogue of the 232-register FWD3 build, CW = 128): the warp's 32 rows x 128 columns of C2 = gelu(...) are one scale
// block per row of the NEXT Linear's A operand: the values stay in registers, the row maxima are reduced over the 4 lanes that share
// a row, and hi / lo / scale are written next to the fp32 tensors — the next GEMM needs no pre-pass.
template <int EPI, int CW, bool SPLIT = false>
__device__ __forceinline__ void fwd16_epilogue(const F16Params& p, const float (&sum)[CW], float* stage, int lane, int row0, int cbase) {
    const int tr = lane >> 2, tc = 4 * (lane & 3);
    float4 g[SPLIT ? CW / 4 : 1];
    float rmax[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int cc = 0; cc < CW / 16; ++cc) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = sum[cc * 16 + j];
        epi16_stage_rows(stage, lane, v);
        const int col = cbase + cc * 16 + tc;
        const float4 cs = __ldg(reinterpret_cast<const float4*>(p.cs + col));
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI != FP_STORE && EPI != FP_GELU_BWD && p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + 8 * i + tr;
            if (!SPLIT && row >= p.M) continue;
            float4 a = epi16_read_t(stage, lane, i);
            a.x *= cs.x; a.y *= cs.y; a.z *= cs.z; a.w *= cs.w;                         // exact (powers of two)
            float4 o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w), o2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EPI == FP_BIAS_GELU) {
                o2 = make_float4(te_gelu(o.x), te_gelu(o.y), te_gelu(o.z), te_gelu(o.w));
                if (SPLIT) {
                    g[cc * 4 + i] = o2;
                    rmax[i] = fmaxf(rmax[i], te_absmax4(o2));
                    if (row >= p.M) continue;
                }
            } else if (EPI == FP_BIAS_ADD) {
                const float4 e = *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col);
                o2 = make_float4(e.x + o.x, e.y + o.y, e.z + o.z, e.w + o.w);
            } else if (EPI == FP_GELU_BWD) {
                const float4 e = *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col);
                o = make_float4(a.x * te_gelu_grad(e.x), a.y * te_gelu_grad(e.y), a.z * te_gelu_grad(e.z), a.w * te_gelu_grad(e.w));
            }
            *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = o;
            if (EPI == FP_BIAS_GELU || EPI == FP_BIAS_ADD) *reinterpret_cast<float4*>(p.C2 + (long long)row * p.ldc2 + col) = o2;
        }
    }
    if (SPLIT) {
        const int nblk = p.N / 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float m = rmax[i];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            float sc, si;
            te_f16_block_scale(m, sc, si);
            const int row = row0 + 8 * i + tr;
            if (row >= p.M) continue;
            if ((lane & 3) == 0) p.rs2[(long long)row * nblk + cbase / 128] = si;
#pragma unroll
            for (int cc = 0; cc < CW / 16; ++cc) {
                uint2 h, l;
                te_f16_split4(g[cc * 4 + i], sc, h, l);
                const long long off = (long long)row * p.N + cbase + cc * 16 + tc;
                *reinterpret_cast<uint2*>(p.hi2 + off) = h;
                *reinterpret_cast<uint2*>(p.lo2 + off) = l;
            }
        }
    }
}

// FM_R: sum[0..31] = (S W+) and sum[32..63] = (S W-) of one row x 32 columns;  R_in = x+ * (S W+) + x- * (S W-)   (layers_ours.py:207-230)
__device__ __forceinline__ void r16_epilogue(const F16Params& p, const float (&sum)[64], float* stage, int lane, int row0, int cbase) {
    const int tr = lane >> 2, tc = 4 * (lane & 3);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int col = cbase + cc * 16 + tc;
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + 8 * i + tr;
            x[i] = (row < p.M) ? *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float4 cp = __ldg(reinterpret_cast<const float4*>(p.cs + col));
        const float4 cn = __ldg(reinterpret_cast<const float4*>(p.cs1 + col));
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = sum[cc * 16 + j];
        epi16_stage_rows(stage, lane, v);
        float4 ap[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ap[i] = epi16_read_t(stage, lane, i);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = sum[32 + cc * 16 + j];
        epi16_stage_rows(stage, lane, v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + 8 * i + tr;
            if (row >= p.M) continue;
            const float4 an = epi16_read_t(stage, lane, i);
            float4 o;
            o.x = fmaxf(x[i].x, 0.f) * (ap[i].x * cp.x) + fminf(x[i].x, 0.f) * (an.x * cn.x);
            o.y = fmaxf(x[i].y, 0.f) * (ap[i].y * cp.y) + fminf(x[i].y, 0.f) * (an.y * cn.y);
            o.z = fmaxf(x[i].z, 0.f) * (ap[i].z * cp.z) + fminf(x[i].z, 0.f) * (an.z * cn.z);
            o.w = fmaxf(x[i].w, 0.f) * (ap[i].w * cp.w) + fminf(x[i].w, 0.f) * (an.w * cn.w);
            *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = o;
        }
    }
}

template <int MODE, int EPI, bool SPLIT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FCfg<MODE>::THREADS, 1)
te_tc_fwd16_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const F16Params p) {
    using Cfg = FCfg<MODE>;
    constexpr int F16_NST = Cfg::NST, F16_STAGE = Cfg::STAGE, F16_NBARS = Cfg::NBARS, TN = Cfg::TN, CW = Cfg::CW;
    constexpr int F16_DRAIN_WARPS = Cfg::DW;
    constexpr uint32_t OFF_B0 = Cfg::NA * F16_TILE, OFF_B1 = OFF_B0 + Cfg::BT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr int RING = F16_NST * F16_STAGE, EPIB = F16_DRAIN_WARPS * EPI16_STAGE_BYTES;
    const uint32_t bars = smem_base + RING + EPIB;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (F16_NST + s); };
    auto accfull_bar = [&](int b) { return bars + 8u * (2 * F16_NST + b); };
    auto accfree_bar = [&](int b) { return bars + 8u * (2 * F16_NST + 2 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + RING + EPIB + 8 * F16_NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int ntiles = p.tiles_m * p.tiles_n;
    const int kb = p.K / F16_K;
    const int nchunks = (kb + F16_CHUNK - 1) / F16_CHUNK;
    constexpr uint32_t TMEM_COLS = 512u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
        if (Cfg::NA == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
        if (Cfg::NB == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBl) : "memory");
        for (int s = 0; s < F16_NST; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accfull_bar(b), 1);
            mbar_init(accfree_bar(b), 2u * F16_DRAIN_WARPS);
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

    if (warp < Cfg::W0) {
    // register reallocation between the warpgroups (see FCfg): issued inside the role branches so that the allocator sees which
    // budget governs which code
    if (MODE == FM_FWD3) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ================= TMA producer (both CTAs: own activation rows + own half of the weight tile) =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                const int m0 = ((t / p.tiles_n) * 2 + (int)rank) * BM, n0 = (t % p.tiles_n) * TN + (int)rank * (TN / 2);
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const int s = (int)(it % F16_NST);
                    const uint32_t ph = (it / F16_NST) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = smem_base + s * F16_STAGE;
                    if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * F16_STAGE);
                    tma2_load_2d(sa, &tmAh, full_bar(s), kk * F16_K, m0);
                    tma2_load_2d(sa + OFF_B0, &tmBh, full_bar(s), kk * F16_K, n0);
                    if (Cfg::NA == 2) tma2_load_2d(sa + F16_TILE, &tmAl, full_bar(s), kk * F16_K, m0);
                    if (Cfg::NB == 2) tma2_load_2d(sa + OFF_B1, &tmBl, full_bar(s), kk * F16_K, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && lane == 0) {
            uint32_t it = 0, gc = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const bool chunk_start = (kk % F16_CHUNK) == 0;
                    const uint32_t b = gc & 1u;
                    if (chunk_start && gc >= 2) {                       // accumulator b drained (chunk gc-2) in BOTH CTAs
                        mbar_wait_cluster(accfree_bar(b), ((gc >> 1) & 1u) ^ 1u);
                        tcgen05_fence_after();
                    }
                    const int s = (int)(it % F16_NST);
                    const uint32_t ph = (it / F16_NST) & 1u;
                    mbar_wait_cluster(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + s * F16_STAGE;
                    const uint64_t ah = make_smem_desc(sa), al = make_smem_desc(sa + F16_TILE);
                    const uint64_t bh = make_smem_desc(sa + OFF_B0), bl = make_smem_desc(sa + OFF_B1);
                    const uint32_t d = tmem_base + b * (uint32_t)BN;
#pragma unroll
                    for (int k = 0; k < F16_K / 16; ++k) {
                        const uint64_t o = (uint64_t)(2 * k);
                        const uint32_t acc = (chunk_start && k == 0) ? 0u : 1u;
                        umma2_bf16(d, ah + o, bh + o, Cfg::IDESC, acc);
                        if (MODE == FM_FWD3) {
                            umma2_bf16(d, al + o, bh + o, Cfg::IDESC, 1u);
                            umma2_bf16(d, ah + o, bl + o, Cfg::IDESC, 1u);
                        } else if (MODE == FM_R) {
                            umma2_bf16(d + 128u, ah + o, bl + o, Cfg::IDESC, acc);      // second product: S W-
                        }
                    }
                    umma2_commit_both(empty_bar(s));
                    if ((kk % F16_CHUNK) == F16_CHUNK - 1 || kk == kb - 1) { umma2_commit_both(accfull_bar(b)); ++gc; }
                }
            }
        }
        __syncwarp();
    }
    } else {
        if (MODE == FM_FWD3) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ================= chunk drain + epilogue: warps W0 .. W0 + DW =================
        // lane quarter q = warp % 4 (the TMEM lanes a warp may read); column group cg: FWD3 / LIN1 own CW consecutive accumulator
        // columns, FM_R owns 32 columns of BOTH 128-column products (sum[0..31] = S W+, sum[32..63] = S W-)
        const int q = warp & 3;
        const int cg = (warp - Cfg::W0) >> 2;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * (MODE == FM_R ? 32 : CW));
        float* stage = reinterpret_cast<float*>(smem_al + RING + (warp - Cfg::W0) * EPI16_STAGE_BYTES);
        float sum[CW];
        uint32_t gc = 0;
        for (int t = cluster_id; t < ntiles; t += nclusters) {
            const int m0 = ((t / p.tiles_n) * 2 + (int)rank) * BM, n0 = (t % p.tiles_n) * TN;
            const int myrow = m0 + q * 32 + lane;                      // the accumulator row (TMEM lane) of this thread
            const float* rsr = p.rs + (long long)(myrow < p.M ? myrow : 0) * p.rs_ld;
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const uint32_t b = gc & 1u;
                const float bs = (myrow < p.M) ? __ldg(rsr + c) : 0.f;      // 2^-e of this row's block c
                mbar_wait(accfull_bar(b), (gc >> 1) & 1u);
                tcgen05_fence_after();
#pragma unroll
                for (int cc = 0; cc < CW / 16; ++cc) {
                    uint32_t v[16];
                    // FM_R: chunks 0-1 of the first product, 2-3 of the second (128 columns further)
                    tmem_ld16(tlane + b * (uint32_t)BN + (uint32_t)(MODE == FM_R ? (cc >> 1) * 128 + (cc & 1) * 16 : cc * 16), v);
                    tmem_ld_wait();
                    if (c == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[cc * 16 + j] = __uint_as_float(v[j]) * bs;
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[cc * 16 + j] = fmaf(__uint_as_float(v[j]), bs, sum[cc * 16 + j]);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_rank0(accfree_bar(b)));
            }
            if constexpr (MODE == FM_R) r16_epilogue(p, sum, stage, lane, m0 + q * 32, n0 + cg * 32);
            else fwd16_epilogue<EPI, CW, SPLIT>(p, sum, stage, lane, m0 + q * 32, n0 + cg * CW);
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- fp16 split pre-passes --------------------------------------------------------------------------------------
// Activations: one warp per row, one scale block per warp iteration (128 columns): single pass, one read of x.
__global__ void __launch_bounds__(256) blocksplit_f16_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols,
                                                             __half* __restrict__ hi, __half* __restrict__ lo,
                                                             float* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    const int nblk = (cols + 127) / 128;
    for (long long row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const float* xr = x + row * ldx;
#pragma unroll 2
        for (int base = 0; base < cols; base += 128) {
            const int i = base + lane * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < cols) v = *reinterpret_cast<const float4*>(xr + i);
            float s, si;
            te_f16_block_scale(te_warp_max(te_absmax4(v)), s, si);
            if (i < cols) {
                uint2 h, l;
                te_f16_split4(v, s, h, l);
                *reinterpret_cast<uint2*>(hi + row * cols + i) = h;
                if (lo) *reinterpret_cast<uint2*>(lo + row * cols + i) = l;
            }
            if (lane == 0) inv[row * nblk + base / 128] = si;
        }
    }
}
// Weights: one scale per row of W (two passes over the row; the second hits L1 / L2).
__global__ void __launch_bounds__(256) rowsplit_f16_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols4,
                                                           __half* __restrict__ hi, __half* __restrict__ lo,
                                                           float* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    for (long long row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
        float m = 0.f;
#pragma unroll 4
        for (int c = lane; c < cols4; c += 32) m = fmaxf(m, te_absmax4(xr[c]));
        float s, si;
        te_f16_block_scale(te_warp_max(m), s, si);
        uint2* hr = reinterpret_cast<uint2*>(hi + row * (long long)cols4 * 4);
        uint2* lr = reinterpret_cast<uint2*>(lo + row * (long long)cols4 * 4);
#pragma unroll 4
        for (int c = lane; c < cols4; c += 32) {
            uint2 h, l;
            te_f16_split4(xr[c], s, h, l);
            hr[c] = h;
            if (lo) lr[c] = l;
        }
        if (lane == 0) inv[row] = si;
    }
}

bool make_map_f16(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)F16_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int f16_sm_pairs() {
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

template <int MODE, int EPI, bool SPLIT = false>
int launch_f16(const __half* ah, const __half* al, const __half* b0, const __half* b1, F16Params p, cudaStream_t st) {
    using Cfg = FCfg<MODE>;
    CUtensorMap tmAh, tmAl, tmB0, tmB1;
    if (!make_map_f16(&tmAh, ah, p.M, p.K, BM) || !make_map_f16(&tmAl, al ? al : ah, p.M, p.K, BM) ||
        !make_map_f16(&tmB0, b0, p.N, p.K, Cfg::TN / 2) || !make_map_f16(&tmB1, b1 ? b1 : b0, p.N, p.K, Cfg::TN / 2)) {
        te_set_last_error("te_tc_fwd16: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;
    if (!smem_optin(te_tc_fwd16_kernel<MODE, EPI, SPLIT>, Cfg::SMEM, optin)) {
        te_set_last_error("te_tc_fwd16: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    const int mt = (p.M + BM - 1) / BM;
    p.tiles_m = (mt + 1) / 2;
    p.tiles_n = p.N / Cfg::TN;
    const int ntiles = p.tiles_m * p.tiles_n;
    int pairs = f16_sm_pairs();
    if (pairs <= 0) { te_set_last_error("te_tc_fwd16: cannot query the SM count"); return TE_ERR_CUDA; }
    if (pairs > ntiles) pairs = ntiles;
    te_tc_fwd16_kernel<MODE, EPI, SPLIT><<<dim3(2u * (unsigned)pairs), Cfg::THREADS, Cfg::SMEM, st>>>(tmAh, tmAl, tmB0, tmB1, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

}  // namespace

bool te_tc_fwd16_supported(long long rows, int K, int N, long long lda) {
    return rows > 0 && rows < (1LL << 31) && K % F16_K == 0 && N % BN == 0 && lda % 4 == 0 && get_encode() != nullptr;
}


// weights: W [rows, cols] (row stride ldx) -> hi, lo fp16 [rows, cols] and ONE 2^-f per row
int te_tc_rowsplit_f16(const float* x, long long ldx, long long rows, int cols, void* hi, void* lo, float* scale_inv,
                       cudaStream_t st) {
    if (cols % 4 != 0 || ldx % 4 != 0 || !a16(x) || ((uintptr_t)hi & 7u) || (lo && ((uintptr_t)lo & 7u))) {
        te_set_last_error("te_tc_rowsplit_f16: alignment");
        return TE_ERR_ARG;
    }
    long long blocks = (rows + 7) / 8;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    if (blocks < 1) blocks = 1;
    rowsplit_f16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ldx, rows, cols / 4, reinterpret_cast<__half*>(hi),
                                                          reinterpret_cast<__half*>(lo), scale_inv);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

// activations: x [rows, cols] (row stride ldx) -> split = [hi | lo] fp16 [rows, cols] and one 2^-e per (row, 128 columns):
// scale_inv [rows, ceil(cols / 128)]
int te_tc_blocksplit_f16(const float* x, long long ldx, long long rows, int cols, float* split, float* scale_inv, cudaStream_t st,
                         bool hi_only) {
    if (cols % 4 != 0 || ldx % 4 != 0 || !a16(x) || !a16(split) || !scale_inv) {
        te_set_last_error("te_tc_blocksplit_f16: alignment");
        return TE_ERR_ARG;
    }
    long long blocks = (rows + 7) / 8;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    if (blocks < 1) blocks = 1;
    __half* hi = reinterpret_cast<__half*>(split);
    blocksplit_f16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ldx, rows, cols, hi, hi_only ? nullptr : hi + rows * cols, scale_inv);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

// y[rows,out] = x[rows,in] W^T (+ epilogue), fp32-grade.
//   split / scale: the block-scaled split of x ([hi | lo] fp16 [rows, in] = rows*in floats; [rows, ceil(in/128)] floats).  x != NULL:
//   scratch, filled here by the pre-pass.  x == NULL: already filled by the producer of x (te_launch_layernorm_split, or the
//   split_out / scale_out of the previous call).
//   split_out / scale_out (TE_TC_EPI_BIAS_GELU only, may be NULL): receive the split of y2 = gelu(y) for the next Linear.
// The weight split lives in the derived buffer (te_gemm_tc.h).
int te_tc_linear_fwd16(const float* x, long long ldx, float* split, float* scale, const float* derived, int in_features,
                       int out_features, const float* bias, float* y, float* y2, const float* e0, long long rows, int epi,
                       cudaStream_t st, float* split_out, float* scale_out) {
    const long long n = (long long)in_features * out_features;
    if (!a16(split) || !scale || !a16(derived) || !a16(y) || (y2 && !a16(y2)) || (e0 && !a16(e0)) || (bias && !a16(bias)) ||
        (split_out && (!a16(split_out) || !scale_out || epi != TE_TC_EPI_BIAS_GELU))) {
        te_set_last_error("te_tc_linear_fwd16: bad operands");
        return TE_ERR_ARG;
    }
    __half* ah = reinterpret_cast<__half*>(split);
    __half* al = ah + rows * in_features;
    if (x) TE_TRY(te_tc_blocksplit_f16(x, ldx, rows, in_features, split, scale, st));
    const __half* bh = reinterpret_cast<const __half*>(derived + 11 * n + n / 2);
    const __half* bl = reinterpret_cast<const __half*>(derived + 12 * n);
    F16Params p;
    memset(&p, 0, sizeof(p));
    p.M = (int)rows; p.N = out_features; p.K = in_features;
    p.rs = scale; p.rs_ld = (in_features + 127) / 128; p.cs = derived + 12 * n + n / 2;
    p.bias = bias; p.E = e0; p.lde = out_features; p.C = y; p.ldc = out_features; p.C2 = y2; p.ldc2 = out_features;
    if (split_out) {
        p.hi2 = reinterpret_cast<__half*>(split_out);
        p.lo2 = p.hi2 + rows * out_features;
        p.rs2 = scale_out;
        return launch_f16<FM_FWD3, FP_BIAS_GELU, true>(ah, al, bh, bl, p, st);
    }
    switch (epi) {
        case TE_TC_EPI_STORE: return launch_f16<FM_FWD3, FP_STORE>(ah, al, bh, bl, p, st);
        case TE_TC_EPI_BIAS: return launch_f16<FM_FWD3, FP_BIAS>(ah, al, bh, bl, p, st);
        case TE_TC_EPI_BIAS_GELU: return launch_f16<FM_FWD3, FP_BIAS_GELU>(ah, al, bh, bl, p, st);
        case TE_TC_EPI_BIAS_ADD: return launch_f16<FM_FWD3, FP_BIAS_ADD>(ah, al, bh, bl, p, st);
    }
    te_set_last_error("te_tc_linear_fwd16: unsupported epilogue");
    return TE_ERR_UNSUPPORTED;
}

// ---- single-pass products -------------------------------------------------------------------------------------------
// dx[rows, in] = epi(dy[rows, out] W)  (activation-gradient backward Linear; TE_FLAG_F16_SINGLE_PASS).  split / scale: hi-only
// block-scaled split of dy ([rows, out] fp16 = rows*out/2 floats; [rows, ceil(out/128)]); dy != NULL: filled here by the pre-pass,
// dy == NULL: already filled by the producer.  fp16(W^T) [in, out] + its row scales live in the derived buffer at 13 n.
int te_tc_linear_bwd16(const float* dy, long long lddy, float* split, float* scale, const float* derived, int in_features,
                       int out_features, float* dx, const float* e0, long long rows, int epi, cudaStream_t st) {
    const long long n = (long long)in_features * out_features;
    if (!a16(split) || !scale || !a16(derived) || !a16(dx) || (e0 && !a16(e0))) {
        te_set_last_error("te_tc_linear_bwd16: bad operands");
        return TE_ERR_ARG;
    }
    if (dy) TE_TRY(te_tc_blocksplit_f16(dy, lddy, rows, out_features, split, scale, st, true));
    F16Params p;
    memset(&p, 0, sizeof(p));
    p.M = (int)rows; p.N = in_features; p.K = out_features;
    p.rs = scale; p.rs_ld = (out_features + 127) / 128; p.cs = derived + 13 * n + n / 2;
    p.E = e0; p.lde = in_features; p.C = dx; p.ldc = in_features;
    const __half* ah = reinterpret_cast<const __half*>(split);
    const __half* bt = reinterpret_cast<const __half*>(derived + 13 * n);
    if (epi == TE_TC_EPI_GELU_BWD) return launch_f16<FM_LIN1, FP_GELU_BWD>(ah, nullptr, bt, nullptr, p, st);
    if (epi == TE_TC_EPI_STORE) return launch_f16<FM_LIN1, FP_STORE>(ah, nullptr, bt, nullptr, p, st);
    te_set_last_error("te_tc_linear_bwd16: unsupported epilogue");
    return TE_ERR_UNSUPPORTED;
}

// R_in[rows, in] = x+ * (S W+) + x- * (S W-)  (second contraction of the z+ rule).  split / scale: hi-only block-scaled split of
// S [rows, out]; s != NULL: filled here by the pre-pass.  fp16(W+^T), fp16(W-^T) [in, out] + row scales: derived buffer at 14 n.
int te_tc_zplus_r16(const float* s, float* split, float* scale, const float* derived, const float* x, long long ldx, float* out,
                    long long ld_out, long long rows, int in_features, int out_features, cudaStream_t st) {
    const long long n = (long long)in_features * out_features;
    if (!a16(split) || !scale || !a16(derived) || !a16(x) || !a16(out) || ldx % 4 != 0 || ld_out % 4 != 0) {
        te_set_last_error("te_tc_zplus_r16: bad operands");
        return TE_ERR_ARG;
    }
    if (s) TE_TRY(te_tc_blocksplit_f16(s, out_features, rows, out_features, split, scale, st, true));
    F16Params p;
    memset(&p, 0, sizeof(p));
    p.M = (int)rows; p.N = in_features; p.K = out_features;
    p.rs = scale; p.rs_ld = (out_features + 127) / 128; p.cs = derived + 15 * n; p.cs1 = derived + 15 * n + in_features;
    p.E = x; p.lde = ldx; p.C = out; p.ldc = ld_out;
    const __half* ah = reinterpret_cast<const __half*>(split);
    const __half* bp = reinterpret_cast<const __half*>(derived + 14 * n);
    return launch_f16<FM_R, FP_STORE>(ah, nullptr, bp, bp + n, p, st);
}
bool te_tc_f16_single_supported(long long rows, int K, int N, long long lda) {
    return rows > 0 && rows < (1LL << 31) && K % F16_K == 0 && N % BN == 0 && lda % 4 == 0 && get_encode() != nullptr;
}
