// Persistent CTA-pair (tcgen05 cta_group::2) GEMM family: the z+ rule's two contractions and the single-pass TF32
// activation-gradient Linear, all as ONE kernel template.
//
// Why pairs: ncu of the round-1 single-CTA kernels (profiles/r01_ncu_summary.md, prof_zplus_s1.ncu-rep) shows them at the
// L2 -> SM bandwidth cap, not at the tensor pipe: lts__t_sectors = 11.1 TB/s (R kernel) / 12.8 TB/s (S1 kernel) =
// 5700 / 6580 bytes per clock against a measured chip cap of ~6300 B/clk, with the tensor pipe 36 % active.  A 128x256
// tile stages 48 KiB per 4 MMAs (512 tensor cycles) = 96 B/clk/SM = 14 200 B/clk chip-wide at full tensor rate: the
// cap allows 44 %.  The fix is flops per staged byte:
//   * a CTA pair executes one 256 x 256 x 8 MMA per issue and each CTA stages only HALF of the weight tile;
//   * the R kernel's two products (S W+ and S W-) share their A operand (S): one A tile + two weight halves per stage,
//     8 MMAs per 48 KiB instead of 4  ->  47 B/clk/SM = 6 940 B/clk chip-wide: the cap allows 91 %.
// Why persistent: the non-persistent pair kernels of round 1 lost to the single-CTA ones (a pair only starts when both
// SMs of a TPC are free, TMEM alloc + two cluster barriers + an exposed prologue per tile).  Here each cluster loops
// over tiles (static round-robin, column tile fastest so concurrently running clusters share activation rows in L2);
// barriers / TMEM are set up once, the TMA producer runs ahead into the next tile during the epilogue, and the modes
// with one 256-column accumulator double-buffer it in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  The R mode owns all 512 columns (two accumulators): its epilogue is shortened instead (8 warps,
// x rows prefetched into L2 at tile start, register-double-buffered loads).
//
//   mode PM_R    R_in = x+ * (S W+) + x- * (S W-)          A = S [M,K]      B0/B1 = W+^T / W-^T [N,K]   (layers_ours.py:207-230)
//   mode PM_S1   S = sd(R, ((y - b) + |x| |W|^T) / 2)       A = tf32(|x|)    B0 = |W| [N,K]
//   mode PM_LIN  C = epi(A B^T)  single-pass TF32            A = dy           B0 = tf32(W)^T  (activation-gradient backward)
//
// The |x| operand of PM_S1 is produced by a separate elementwise pre-pass (te_tc_abs_tf32, one read + one write of x):
// a first version transformed the tile in shared memory (4 warps between the TMA arrival and the MMA, remote arrives onto
// the leader's barrier) and ran at 26 % tensor-memory-pipe activity against 78-82 % for the transform-free modes at the
// same shape (profiles/r02_ncu_pair.md) — the cvt/fence/arrive chain per 512-cycle k-block was the critical path.
//
// Warp roles (both CTAs of the pair): warp 0 TMA producer, warp 1 TMEM allocator + (leader CTA only) MMA issuer,
// warps 2-9 epilogue (lane quarter = warp % 4, column half = (warp - 2) / 4).
// Barriers per CTA:
//   full[s]     LEADER's: both CTAs' cp.async.bulk.tensor .cta_group::2 count on the leader's barrier (expect_tx = both
//               CTAs' bytes)
//   empty[s]    local, tcgen05.commit.cta_group::2 multicast from the leader when the MMAs of the stage retire
//   accfull[b]  local, multicast commit after the last k-block of a tile into accumulator buffer b
//   accfree[b]  leader, one remote arrive per epilogue warp of both CTAs (16) once buffer b has been read out
#include <cuda_fp16.h>

#include "te_tc_common.cuh"

namespace {

enum { PM_R = 0, PM_S1 = 1, PM_LIN = 2 };
enum { PE_STORE = 0, PE_F16 = 1 /* PM_S1: S as block-scaled fp16 for the fp16 R kernel */, PE_GELU_BWD = 4 };

struct PairParams {
    int M, N, K;
    int tiles_m, tiles_n;                 // tiles_m: 256-row pair tiles
    const float* E; long long lde;        // PM_R: x [M,N] ; PM_S1: R [M,N] ; PM_LIN/GELU_BWD: h [M,N]
    float* C; long long ldc;
    const float* Y; long long ldy; const float* bias;                   // PM_S1: saved forward output y = x W^T + b
    const float* X; long long ldx; const float* Wp; const float* Wn;    // PM_S1: exact fallback of a cancelled denominator
    __half* H; float* HS;                 // PM_S1 / PE_F16: S as hi-only block-scaled fp16 [M, N] + 2^-e per (row, 128 columns) [M, N/128]
};

template <int MODE> struct PairCfg {
    static constexpr int NB = (MODE == PM_R) ? 2 : 1;                         // weight operands per stage
    static constexpr int STAGE = A_BYTES + NB * BH_BYTES;                     // 48 KiB / 32 KiB
    static constexpr int NST = (MODE == PM_R) ? 3 : 5;                        // 144 / 160 KiB ring (+ 36 KiB epilogue staging)
    static constexpr int ACC_BUFS = (MODE == PM_R) ? 1 : 2;                   // TMEM accumulator buffers of NB x 256 columns
    static constexpr int THREADS = 320;
    static constexpr int NBARS = 2 * NST + 4;
    static constexpr int SMEM = NST * STAGE + 8 * EPI_STAGE_BYTES + 1024 + 8 * NBARS + 16;
};

constexpr int EPI_WARPS = 8;
constexpr float kTruncComp = 1.0f + 3.4e-4f;       // PM_LIN: compensation of the A-operand truncation bias (see the epilogue)

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// exact z+ denominator of one element from the TF32-rounded W+ / W- copies (rare path, see pair_epilogue_s1)
__device__ __noinline__ float zplus_exact(const float* __restrict__ xrow, const float* __restrict__ wp,
                                          const float* __restrict__ wn, int K) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const float x = xrow[k];
        acc += fmaxf(x, 0.f) * wp[k] + fminf(x, 0.f) * wn[k];
    }
    return acc;
}

// BF (PM_S1 only): |x| and |W| are bf16 (kind::f16, 64 elements per 128-byte row): the denominator term |x||W|^T is a sum of K
// non-negative products whose independent 2^-9 rounding errors average out (relative error ~2^-9 / sqrt(K)), at half the staged
// bytes and half the tensor cycles per flop of the TF32 form — the S1 kernel sits at its L2 cap (62 B/clk/SM, 68 %).
template <int MODE, int EPI, bool BF = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg<MODE>::THREADS, 1)
te_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                  const __grid_constant__ CUtensorMap tmB1, const PairParams p) {
    static_assert(!BF || MODE == PM_S1, "bf16 operands: S1 kernel only");
    constexpr int KELEMS = BF ? 64 : BK;
    using Cfg = PairCfg<MODE>;
    constexpr int NST = Cfg::NST, NB = Cfg::NB, STAGE = Cfg::STAGE, ACC_BUFS = Cfg::ACC_BUFS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + NST * STAGE + 8 * EPI_STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (NST + s); };
    auto accfull_bar = [&](int b) { return bars + 8u * (2 * NST + b); };
    auto accfree_bar = [&](int b) { return bars + 8u * (2 * NST + 2 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + NST * STAGE + 8 * EPI_STAGE_BYTES + 8 * Cfg::NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int ntiles = p.tiles_m * p.tiles_n;
    const int kb = p.K / KELEMS;
    constexpr uint32_t TMEM_COLS = 512u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
        if (NB == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
        for (int s = 0; s < NST; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accfull_bar(b), 1);
            mbar_init(accfree_bar(b), 2u * EPI_WARPS);
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
    cluster_sync_all();                                // barriers of both CTAs initialised, TMEM allocated
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own A tile + own half of every weight tile) =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                const int m0 = ((t / p.tiles_n) * 2 + (int)rank) * BM, n0 = (t % p.tiles_n) * BN + (int)rank * (BN / 2);
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const int s = (int)(it % NST);
                    const uint32_t ph = (it / NST) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = smem_base + s * STAGE;
                    if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * STAGE);
                    tma2_load_2d(sa, &tmA, full_bar(s), kk * KELEMS, m0);
                    tma2_load_2d(sa + A_BYTES, &tmB0, full_bar(s), kk * KELEMS, n0);
                    if (NB == 2) tma2_load_2d(sa + A_BYTES + BH_BYTES, &tmB1, full_bar(s), kk * KELEMS, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && lane == 0) {
            uint32_t it = 0, ti = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters, ++ti) {
                const uint32_t b = ti % ACC_BUFS;
                if (ti >= (uint32_t)ACC_BUFS) {                 // buffer b read out by the epilogues of BOTH CTAs
                    mbar_wait_cluster(accfree_bar(b), ((ti / ACC_BUFS) & 1u) ^ 1u);
                    tcgen05_fence_after();
                }
                const uint32_t d0 = tmem_base + b * (uint32_t)(NB * BN);
                for (int kk = 0; kk < kb; ++kk, ++it) {
                    const int s = (int)(it % NST);
                    const uint32_t ph = (it / NST) & 1u;
                    mbar_wait_cluster(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + s * STAGE;
                    const uint64_t adesc = make_smem_desc(sa);
                    const uint64_t b0desc = make_smem_desc(sa + A_BYTES);
                    const uint64_t b1desc = make_smem_desc(sa + A_BYTES + BH_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint32_t acc = (kk == 0 && k == 0) ? 0u : 1u;
                        if (BF) { umma2_bf16(d0, adesc + (uint64_t)(2 * k), b0desc + (uint64_t)(2 * k), kIdesc2Bf16, acc); continue; }
                        umma2_tf32(d0, adesc + (uint64_t)(2 * k), b0desc + (uint64_t)(2 * k), kIdesc2, acc);
                        if (NB == 2) umma2_tf32(d0 + (uint32_t)BN, adesc + (uint64_t)(2 * k), b1desc + (uint64_t)(2 * k), kIdesc2, acc);
                    }
                    umma2_commit_both(empty_bar(s));        // frees this stage in BOTH CTAs when the MMAs retire
                }
                umma2_commit_both(accfull_bar(b));
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: warps 2..9 =================
        // All global traffic of the epilogue is issued in the TRANSPOSED layout of epi_read_t (4 rows x 128 B per warp
        // instruction): the accumulator chunk goes TMEM -> registers (lane = row) -> per-warp staging buffer -> registers
        // (lane = 4 columns of row 4i + lane/8), the operands (x / R, y / h) are loaded and the result stored in that layout.
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;            // column half of the 256-column tile
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        float* stage = reinterpret_cast<float*>(smem_al + NST * STAGE + (warp - 2) * EPI_STAGE_BYTES);
        const int tr = lane >> 3, tc = 4 * (lane & 7);         // transposed coordinates inside a 32 x 32 chunk (row 4i + tr)
        uint32_t ti = 0;
        for (int t = cluster_id; t < ntiles; t += nclusters, ++ti) {
            const int m0 = ((t / p.tiles_n) * 2 + (int)rank) * BM + q * 32, n0 = (t % p.tiles_n) * BN + half * (BN / 2);
            const uint32_t b = ti % ACC_BUFS;
            // the epilogue operands of this tile are streamed from HBM exactly once: pull them into L2 while the MMAs run
            if (m0 + lane < p.M) {
                if (p.E) {
                    const float* e = p.E + (long long)(m0 + lane) * p.lde + n0;
#pragma unroll
                    for (int j = 0; j < BN / 2; j += 32) prefetch_l2(e + j);
                }
                if (MODE == PM_S1) {
                    const float* y = p.Y + (long long)(m0 + lane) * p.ldy + n0;
#pragma unroll
                    for (int j = 0; j < BN / 2; j += 32) prefetch_l2(y + j);
                }
            }
            // PM_R: its accumulators fill TMEM, so this epilogue is NOT overlapped by the next tile's MMAs — keep it short.  The x operand
            // does not depend on the accumulators: chunk 0's loads are issued before the wait for the last MMA, and every chunk reloads
            // its x registers for the NEXT chunk as soon as it has consumed them (software pipeline without extra registers).
            float4 xr[MODE == PM_R ? 8 : 1];
            if (MODE == PM_R) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = m0 + 4 * i + tr;
                    xr[i] = (row < p.M) ? *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + n0 + tc) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            mbar_wait(accfull_bar(b), (ti / ACC_BUFS) & 1u);
            tcgen05_fence_after();
            const uint32_t tcol = tlane + b * (uint32_t)(NB * BN) + (uint32_t)(half * (BN / 2));
            float rmax[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // PE_F16: max |S| of rows 4i + tr over this warp's 128 columns
#pragma unroll 1
            for (int c = 0; c < BN / 2 / 32; ++c) {
                const int col = n0 + c * 32 + tc;
                uint32_t acc[32];
                tmem_ld32(tcol + (uint32_t)(c * 32), acc);
                if (MODE == PM_R) {
                    uint32_t accn[32];
                    tmem_ld32(tcol + (uint32_t)(BN + c * 32), accn);
                    tmem_ld_wait();
                    epi_stage_rows(stage, lane, acc);
                    float4 ap[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) ap[i] = epi_read_t(stage, lane, i);
                    epi_stage_rows(stage, lane, accn);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 an = epi_read_t(stage, lane, i);
                        const int row = m0 + 4 * i + tr;
                        const float4 x = xr[MODE == PM_R ? i : 0];
                        float4 o;
                        o.x = fmaxf(x.x, 0.f) * ap[i].x + fminf(x.x, 0.f) * an.x;
                        o.y = fmaxf(x.y, 0.f) * ap[i].y + fminf(x.y, 0.f) * an.y;
                        o.z = fmaxf(x.z, 0.f) * ap[i].z + fminf(x.z, 0.f) * an.z;
                        o.w = fmaxf(x.w, 0.f) * ap[i].w + fminf(x.w, 0.f) * an.w;
                        if (c + 1 < BN / 2 / 32 && row < p.M)                 // next chunk's x into the register just consumed
                            xr[MODE == PM_R ? i : 0] = *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col + 32);
                        if (row < p.M) *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = o;
                    }
                } else if (MODE == PM_S1) {
                    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                    // every global load of the chunk is in flight before the TMEM read completes
                    float4 r[8], y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = m0 + 4 * i + tr;
                        const bool ok = row < p.M;
                        r[i] = ok ? *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                        y[i] = ok ? *reinterpret_cast<const float4*>(p.Y + (long long)row * p.ldy + col) : bb;
                    }
                    tmem_ld_wait();
                    epi_stage_rows(stage, lane, acc);
                    // x+ W+^T + x- W-^T == ( x W^T + |x| |W|^T ) / 2 ,  x W^T = y - bias (saved forward output).  The true
                    // value is a sum of non-negative products; the identity cancels when almost every product is negative
                    // (x W^T ~ -|x||W|^T): then Z carries an absolute error of ~2^-11 * a and is recomputed exactly in a
                    // second, warp-voted pass (rare; all-zero rows / columns give an exact 0 on both sides).
                    unsigned redo = 0u;                                  // bit 4i+u: element u of iteration i is cancelled
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 a4 = epi_read_t(stage, lane, i);
                        const float yy[4] = {y[i].x - bb.x, y[i].y - bb.y, y[i].z - bb.z, y[i].w - bb.w};
                        const float rr[4] = {r[i].x, r[i].y, r[i].z, r[i].w};
                        const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
                        float o[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float z = 0.5f * (yy[u] + aa[u]);
                            if (z < aa[u] * 0.0078125f && aa[u] > 0.f) redo |= 1u << (4 * i + u);
                            const float sv = te_sd_fast_nonneg(rr[u], fmaxf(z, 0.f));      // 2 ulp; S is rounded to 11 bits right below
                            o[u] = (EPI == PE_F16) ? sv : to_tf32(sv);
                        }
                        r[i] = make_float4(o[0], o[1], o[2], o[3]);          // r[] now holds S
                    }
                    if (EPI == PE_F16) {
                        // S leaves as block-scaled fp16: one power of two per (row, this warp's 128 columns).  The chunk's S values are
                        // parked in the TMEM columns the accumulator chunk just vacated (a thread reads back exactly the registers it
                        // stored, so the transposed layout needs no second transposition) until the row maxima of all 4 chunks exist.
                        if (__any_sync(0xffffffffu, redo != 0u)) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) {
                                if (!((redo >> e) & 1u)) continue;
                                const int i = e >> 2, u = e & 3;
                                const int row = m0 + 4 * i + tr;
                                if (row >= p.M) continue;
                                const float z = zplus_exact(p.X + (long long)row * p.ldx, p.Wp + (long long)(col + u) * p.K,
                                                            p.Wn + (long long)(col + u) * p.K, p.K);
                                const float sv = te_sd(p.E[(long long)row * p.lde + col + u], fmaxf(z, 0.f));
                                if (u == 0) r[i].x = sv; else if (u == 1) r[i].y = sv; else if (u == 2) r[i].z = sv; else r[i].w = sv;
                            }
                        }
                        uint32_t sv32[32];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            rmax[i] = fmaxf(rmax[i], te_absmax4(r[i]));
                            sv32[4 * i] = __float_as_uint(r[i].x); sv32[4 * i + 1] = __float_as_uint(r[i].y);
                            sv32[4 * i + 2] = __float_as_uint(r[i].z); sv32[4 * i + 3] = __float_as_uint(r[i].w);
                        }
                        tmem_st32(tcol + (uint32_t)(c * 32), sv32);
                        continue;
                    }
                    if (__any_sync(0xffffffffu, redo != 0u)) {
                        // r[] was overwritten: reload the relevance of the few cancelled elements
                        for (int e = 0; e < 32; ++e) {
                            if (!((redo >> e) & 1u)) continue;
                            const int i = e >> 2, u = e & 3;
                            const int row = m0 + 4 * i + tr;
                            if (row >= p.M) continue;
                            const float z = zplus_exact(p.X + (long long)row * p.ldx, p.Wp + (long long)(col + u) * p.K,
                                                        p.Wn + (long long)(col + u) * p.K, p.K);
                            const float sv = to_tf32(te_sd(p.E[(long long)row * p.lde + col + u], fmaxf(z, 0.f)));
                            p.C[(long long)row * p.ldc + col + u] = sv;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = m0 + 4 * i + tr;
                        if (row >= p.M) continue;
                        float4 o = r[i];
                        if (redo & (0xFu << (4 * i))) {                      // keep the exactly recomputed elements
                            float* cp = p.C + (long long)row * p.ldc + col;
                            if (redo & (1u << (4 * i + 0))) o.x = cp[0];
                            if (redo & (1u << (4 * i + 1))) o.y = cp[1];
                            if (redo & (1u << (4 * i + 2))) o.z = cp[2];
                            if (redo & (1u << (4 * i + 3))) o.w = cp[3];
                        }
                        *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = o;
                    }
                } else {
                    float4 e[8];
                    if (EPI == PE_GELU_BWD) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = m0 + 4 * i + tr;
                            e[i] = (row < p.M) ? *reinterpret_cast<const float4*>(p.E + (long long)row * p.lde + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    tmem_ld_wait();
                    epi_stage_rows(stage, lane, acc);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = m0 + 4 * i + tr;
                        float4 o = epi_read_t(stage, lane, i);
                        // the tensor core TRUNCATES the raw fp32 A operand to TF32 (the weights are rounded once, unbiased):
                        // every |a| shrinks by a relative 2^-11 * E[1/m] ~ 3.4e-4 on average (mantissa m in [1,2)).  Undo the
                        // systematic part; what is left is a zero-mean error of the same size as round-to-nearest would give.
                        o.x *= kTruncComp; o.y *= kTruncComp; o.z *= kTruncComp; o.w *= kTruncComp;
                        if (EPI == PE_GELU_BWD) {
                            o.x *= te_gelu_grad_fast(e[i].x); o.y *= te_gelu_grad_fast(e[i].y);
                            o.z *= te_gelu_grad_fast(e[i].z); o.w *= te_gelu_grad_fast(e[i].w);
                        }
                        if (row < p.M) *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = o;
                    }
                }
            }
            if (MODE == PM_S1 && EPI == PE_F16) {
                tmem_st_wait();
                float sc[8];
                const int nblk = p.N / 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float m = rmax[i];
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                    float si;
                    te_f16_block_scale(m, sc[i], si);
                    const int row = m0 + 4 * i + tr;
                    if ((lane & 7) == 0 && row < p.M) p.HS[(long long)row * nblk + n0 / 128] = si;
                }
#pragma unroll 1
                for (int c = 0; c < BN / 2 / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tcol + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = m0 + 4 * i + tr;
                        if (row >= p.M) continue;
                        const __half2 h01 = __floats2half2_rn(__uint_as_float(v[4 * i]) * sc[i], __uint_as_float(v[4 * i + 1]) * sc[i]);
                        const __half2 h23 = __floats2half2_rn(__uint_as_float(v[4 * i + 2]) * sc[i], __uint_as_float(v[4 * i + 3]) * sc[i]);
                        uint2 h;
                        h.x = *reinterpret_cast<const uint32_t*>(&h01);
                        h.y = *reinterpret_cast<const uint32_t*>(&h23);
                        *reinterpret_cast<uint2*>(p.H + (long long)row * p.N + n0 + c * 32 + tc) = h;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_rank0(accfree_bar(b)));
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();                                // nobody leaves while the peer may still touch this CTA
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// |x| rounded to TF32 (rna): the A operand of the single-pass S kernel.  x rows at stride ldx -> compact [rows, cols].
__global__ void abs_tf32_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out, long long rows, int cols4) {
    const long long total = rows * cols4;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cols4;
        const int c = (int)(t - r * cols4);
        float4 v = *reinterpret_cast<const float4*>(x + r * ldx + 4 * c);
        v.x = to_tf32(fabsf(v.x)); v.y = to_tf32(fabsf(v.y)); v.z = to_tf32(fabsf(v.z)); v.w = to_tf32(fabsf(v.w));
        *reinterpret_cast<float4*>(out + t * 4) = v;
    }
}

// same with a bf16 result: the A operand of the bf16 S1 kernel
__global__ void abs_bf16_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long rows, int cols4) {
    const long long total = rows * cols4;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cols4;
        const int c = (int)(t - r * cols4);
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + 4 * c);
        const __nv_bfloat162 a = __floats2bfloat162_rn(fabsf(v.x), fabsf(v.y)), b = __floats2bfloat162_rn(fabsf(v.z), fabsf(v.w));
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&a);
        pk.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(out + t * 4) = pk;
    }
}

// ---- host ---------------------------------------------------------------------------------------------------
int sm_pairs() {
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

template <int MODE, int EPI, bool BF = false>
int launch_pair(const float* A, long long lda, const float* B0, const float* B1, PairParams p, cudaStream_t st) {
    using Cfg = PairCfg<MODE>;
    CUtensorMap tmA, tmB0, tmB1;
    if (!make_map_t(&tmA, A, p.M, p.K, lda, BM, BF) || !make_map_t(&tmB0, B0, p.N, p.K, p.K, BN / 2, BF) ||
        !make_map_t(&tmB1, B1 ? B1 : B0, p.N, p.K, p.K, BN / 2, BF)) {
        te_set_last_error("te_tc_pair: cuTensorMapEncodeTiled failed");
        return TE_ERR_CUDA;
    }
    static unsigned long long optin = 0;
    if (!smem_optin(te_tc_pair_kernel<MODE, EPI, BF>, Cfg::SMEM, optin)) {
        te_set_last_error("te_tc_pair: cannot raise dynamic shared memory");
        return TE_ERR_CUDA;
    }
    const int mt = (p.M + BM - 1) / BM;
    p.tiles_m = (mt + 1) / 2;
    p.tiles_n = p.N / BN;
    const int ntiles = p.tiles_m * p.tiles_n;
    int pairs = sm_pairs();
    if (pairs <= 0) { te_set_last_error("te_tc_pair: cannot query the SM count"); return TE_ERR_CUDA; }
    if (pairs > ntiles) pairs = ntiles;
    te_tc_pair_kernel<MODE, EPI, BF><<<dim3(2u * (unsigned)pairs), Cfg::THREADS, Cfg::SMEM, st>>>(tmA, tmB0, tmB1, p);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

}  // namespace

bool te_tc_pair_supported(long long rows, int K, int N, long long lda) {
    return rows > 0 && rows < (1LL << 31) && K % BK == 0 && N % BN == 0 && lda % 4 == 0 && get_encode() != nullptr;
}

int te_tc_abs_tf32(const float* x, long long ldx, float* out, long long rows, int cols, cudaStream_t st) {
    if (cols % 4 != 0 || ldx % 4 != 0 || !a16(x) || !a16(out)) { te_set_last_error("te_tc_abs_tf32: alignment"); return TE_ERR_ARG; }
    const long long total = rows * (cols / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    abs_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ldx, out, rows, cols / 4);
    TE_CUDA_CHECK_LAUNCH();
    return TE_OK;
}

// S = sd(R, ((y - bias) + |x||W|^T)/2)  [rows, out]   (single-pass denominator, see te_tc_zplus.cu)
// xabs: scratch [rows, in] that receives tf32(|x|)
int te_tc_pair_zplus_s1(const float* x, long long ldx, float* xabs, const float* derived, const float* r, long long ldr,
                        const float* y, long long ldy, const float* bias, float* s_out, long long rows, int in_features,
                        int out_features, cudaStream_t st, bool bf16, float* s16, float* s16_scale) {
    const long long n = (long long)in_features * out_features;
    if (bf16 && in_features % 64 == 0) {
        PairParams p;
        memset(&p, 0, sizeof(p));
        p.M = (int)rows; p.N = out_features; p.K = in_features;
        p.E = r; p.lde = ldr; p.C = s_out; p.ldc = out_features; p.Y = y; p.ldy = ldy; p.bias = bias;
        p.X = x; p.ldx = ldx; p.Wp = derived; p.Wn = derived + n;
        if (ldx % 4 != 0 || !a16(x) || !a16(xabs)) { te_set_last_error("te_tc_pair_zplus_s1: alignment"); return TE_ERR_ARG; }
        const long long total = rows * (in_features / 4);
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 16) blocks = 148LL * 16;
        abs_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(xabs), rows, in_features / 4);
        TE_CUDA_CHECK_LAUNCH();
        // bf16(|W|) [out, in] lives at derived + 11 n (te_tc_prepare_weights)
        if (s16) {                     // S straight to the fp16 R kernel's operand format: no fp32 S, no pre-pass
            p.H = reinterpret_cast<__half*>(s16); p.HS = s16_scale;
            return launch_pair<PM_S1, PE_F16, true>(xabs, in_features, derived + 11 * n, nullptr, p, st);
        }
        return launch_pair<PM_S1, PE_STORE, true>(xabs, in_features, derived + 11 * n, nullptr, p, st);
    }
    TE_TRY(te_tc_abs_tf32(x, ldx, xabs, rows, in_features, st));
    PairParams p;
    memset(&p, 0, sizeof(p));
    p.M = (int)rows; p.N = out_features; p.K = in_features;
    p.E = r; p.lde = ldr; p.C = s_out; p.ldc = out_features; p.Y = y; p.ldy = ldy; p.bias = bias;
    p.X = x; p.ldx = ldx; p.Wp = derived; p.Wn = derived + n;
    if (s16) {
        p.H = reinterpret_cast<__half*>(s16); p.HS = s16_scale;
        return launch_pair<PM_S1, PE_F16>(xabs, in_features, derived + 8 * n, nullptr, p, st);
    }
    return launch_pair<PM_S1, PE_STORE>(xabs, in_features, derived + 8 * n, nullptr, p, st);
}

// R_in = x+ * (S W+) + x- * (S W-)  [rows, in]
int te_tc_pair_zplus_r(const float* s, const float* derived, const float* x, long long ldx, float* out, long long ld_out,
                       long long rows, int in_features, int out_features, cudaStream_t st) {
    const long long n = (long long)in_features * out_features;
    PairParams p;
    memset(&p, 0, sizeof(p));
    p.M = (int)rows; p.N = in_features; p.K = out_features;
    p.E = x; p.lde = ldx; p.C = out; p.ldc = ld_out;
    return launch_pair<PM_R, PE_STORE>(s, out_features, derived + 2 * n, derived + 3 * n, p, st);
}

// dx[rows, in] = epi(dy[rows, out] W)   single-pass TF32 (dy truncated to TF32 by the tensor core, W rounded once)
int te_tc_pair_linear_bwd(const float* dy, long long lddy, const float* derived, int in_features, int out_features, float* dx,
                          const float* e0, long long rows, int epi, cudaStream_t st) {
    const long long n = (long long)in_features * out_features;
    PairParams p;
    memset(&p, 0, sizeof(p));
    p.M = (int)rows; p.N = in_features; p.K = out_features;
    p.E = e0; p.lde = in_features; p.C = dx; p.ldc = in_features;
    if (epi == TE_TC_EPI_GELU_BWD) return launch_pair<PM_LIN, PE_GELU_BWD>(dy, lddy, derived + 6 * n, nullptr, p, st);
    if (epi == TE_TC_EPI_STORE) { p.E = nullptr; return launch_pair<PM_LIN, PE_STORE>(dy, lddy, derived + 6 * n, nullptr, p, st); }
    te_set_last_error("te_tc_pair_linear_bwd: unsupported epilogue");
    return TE_ERR_UNSUPPORTED;
}
