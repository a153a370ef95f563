// Stress shape (BASELINE configs[4]: 4x longer windows, 4x wider LSTM): TRAINING of the wide decoders (H = 96 / 144 / 192)
// on the tensor cores, 16-bit tier (IEEE fp16 operands, fp32 accumulate / cell state, 2e-2 contract).
//
// At H = 192 one layer's W_hh is 288 KB of fp16 -- it cannot stay in shared memory ("W_hh SMEM residency limit") -- so, as
// in the wide inference kernel, the recurrent STATE is resident and the WEIGHTS stream: every step re-reads the weight
// image of the layer from L2 through a TMA ring, one 6 KB K16 slice at a time, in exactly the order the MMAs consume it.
// Only the serial part of a layer runs here; everything that is parallel over time is a plain GEMM and goes to cuBLAS
// (ops.py): the input projection Gx = in . W_ih^T + b for all steps, din = dG . W_ih, dW = dG^T . [in | h_prev].
//
//   lstm_wide_fwd_kernel<NCH>   per step and 48-unit task k:  D = h_{t-1} . W_hh[task]^T  (K = H, N = 192 gate columns),
//                               epilogue: + Gx_t, activations, cell update; saves the ACTIVATED gates (fp16), c (fp32),
//                               h (fp16); h_t goes to the other of two resident operand buffers.  Two task accumulators
//                               in TMEM: the MMAs of task k+1 run under the epilogue of task k.
//   lstm_wide_bwd_kernel<NCH>   per step (descending) and task: epilogue: d(gates) of 48 units from dh_t = dh_in + dh_rec,
//                               the saved gates, c_t, c_{t-1} -> shared memory (A operand, double-buffered) + HBM;
//                               D_R += dG[task] . W_hh[task] (K = 192 gate columns, N = H); two D_R accumulators alternate
//                               between steps.  The running d(cell) lives in an L2-resident per-CTA workspace.
//
// Per-thread vectors live in the "wide tile layout" WTL [T][NT][NCH][3][P][128][16 bytes]: thread (row, 16-unit group g) of
// task k owns P 16-byte pieces (gates / d(gates): 64 fp16 = 8 pieces, in TMEM column order q*16 + gate*4 + u%4, q = u/4;
// h: 16 fp16 = 2 pieces; c, dh: 16 fp32 = 4 pieces), and the 32 lanes of a warp touch 512 CONSECUTIVE bytes per piece.
// (First version: [..][128][E], each thread's vector contiguous -- every 16-byte load of a warp hit 32 different cache lines:
// 30 us per step, long-scoreboard stall 15 per issue, tensor pipe 8 % busy.)  ops.py converts to / from row-major for the GEMMs.
#include "na_tc_common.cuh"

namespace na {
namespace tc {

constexpr int kWtThreads = 14 * 32;
constexpr int kWtMmaWarp = 12, kWtTmaWarp = 13;
// The weight ring moves GROUPS of K16 slices: one TMA bulk copy and one tcgen05.commit per group.  A commit after every
// MMA (first version: 6 KB stages) serialises the tensor pipe -- 0.56 us per slice, 27 us per step, however deep the ring;
// with half a task per group a step issues 8 commits instead of 48.
constexpr int kWtRingBytes = 108 * 1024;

// element offset of piece `piece` (of `np` pieces of `epp` elements = 16 bytes) of thread (row, g) of task k
__device__ __forceinline__ int64_t wtl_off(int t, int ntiles, int tile, int nch, int k, int g, int np, int piece, int row, int epp) {
    return (((((((int64_t)t * ntiles + tile) * nch + k) * 3 + g) * np + piece) * kRows) + row) * epp;
}

__device__ __forceinline__ void wt_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// forward recurrence of one layer
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
struct WfSmem {
    static constexpr int kKC = 6 * NCH;                  // 8-unit K chunks of h
    static constexpr int kSlice = 2 * kN * 16;           // one K16 slice of a task's B operand: [2][192 rows][8] fp16
    static constexpr int kKS = 3 * NCH;                  // K16 slices per task (K = H)
    static constexpr int kGroup = (kKS % 2 == 0) ? kKS / 2 : kKS;       // slices per ring stage
    static constexpr int kStageBytes = kGroup * kSlice;
    static constexpr int kStages = kWtRingBytes / kStageBytes;
    alignas(128) unsigned char h[2][kKC * kAChunk];
    alignas(128) unsigned char ring[kStages][kStageBytes];
    alignas(8) uint64_t ring_full[kStages], ring_empty[kStages];
    uint64_t d_full[2], d_empty[2], h_ready;
    uint32_t tmem_base;
};

template <int NCH>
__global__ void __launch_bounds__(kWtThreads, 1)
lstm_wide_fwd_kernel(const __half* __restrict__ gx,                // WTL E=64: input projection + bias (pre-activations)
                     const unsigned char* __restrict__ wimg,       // [NCH tasks][H/16 slices][2][192][8] fp16
                     __half* __restrict__ gates_out,               // WTL E=64: activated i, f, g, o
                     __half* __restrict__ h_out,                   // WTL E=16
                     float* __restrict__ c_out,                    // WTL E=16
                     int T, int ntiles) {
    using SM = WfSmem<NCH>;
    constexpr int kKS = SM::kKS, kGroup = SM::kGroup, kStages = SM::kStages, kGroupsPerStep = NCH * kKS / kGroup;
    constexpr uint32_t kIdesc = make_idesc(kN, kFmtVal, kFmtVal);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.ring_full[s], 1); mbar_init(&S.ring_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&S.d_full[b], 1); mbar_init(&S.d_empty[b], 12 * 32); }
        mbar_init(&S.h_ready, 12 * 32 * NCH);
        fence_mbar_init();
    }
    if (warp == kWtTmaWarp) tmem_alloc_all(&S.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, S.tmem_base, 0);

    uint32_t sc = 0;          // producer / MMA: ring slices produced / consumed
    uint32_t cnt = 0;         // MMA / epilogue: tasks issued / consumed
    uint32_t hr = 0;          // MMA: h_ready phases consumed
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        {   // h_{-1} = 0 (both buffers)
            uint4* z = reinterpret_cast<uint4*>(S.h[0]);
            for (int i = tid; i < 2 * SM::kKC * kAChunk / 16; i += kWtThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
            fence_proxy_async_smem();
            __syncthreads();
        }
        if (warp == kWtTmaWarp) {
            if (lane == 0)
                for (int t = 0; t < T; ++t)
                    for (int i = 0; i < kGroupsPerStep; ++i, ++sc) {
                        const uint32_t s = sc % kStages, u = sc / kStages;
                        mbar_wait(&S.ring_empty[s], (u & 1) ^ 1);
                        mbar_arrive_expect_tx(&S.ring_full[s], SM::kStageBytes);
                        bulk_load(S.ring[s], wimg + (size_t)i * SM::kStageBytes, SM::kStageBytes, &S.ring_full[s]);
                    }
        } else if (warp == kWtMmaWarp) {
            const bool leader = elect_one();
            const uint64_t d_h[2] = {umma_desc(smem_u32(S.h[0]), kAChunk, 128), umma_desc(smem_u32(S.h[1]), kAChunk, 128)};
            for (int t = 0; t < T; ++t) {
                if (t >= 1) { mbar_wait(&S.h_ready, hr & 1); ++hr; tc_fence_after(); }
                const uint64_t hp = d_h[(t + 1) & 1];                 // h_{t-1}
                for (int k = 0; k < NCH; ++k, ++cnt) {
                    const uint32_t b = cnt & 1;
                    mbar_wait(&S.d_empty[b], ((cnt >> 1) & 1) ^ 1);
                    tc_fence_after();
                    for (int gr = 0; gr < kKS / kGroup; ++gr, ++sc) {
                        const uint32_t s = sc % kStages, u = sc / kStages;
                        mbar_wait(&S.ring_full[s], u & 1);
                        tc_fence_after();
                        const uint64_t d_w = umma_desc(smem_u32(S.ring[s]), kBChunk, 128);
#pragma unroll
                        for (int j = 0; j < kGroup; ++j) {
                            const int ks = gr * kGroup + j;
                            if (leader) umma_bf16_i(tmem + b * kN, desc_adv(hp, 2 * ks * kAChunk), desc_adv(d_w, j * SM::kSlice), kIdesc, ks == 0 ? 0u : 1u);
                        }
                        if (leader) umma_commit(&S.ring_empty[s]);
                    }
                    if (leader) umma_commit(&S.d_full[b]);
                }
            }
            mbar_wait(&S.h_ready, hr & 1); ++hr;                       // the last step's epilogue has finished
        } else {
            const int q = warp & 3, g = warp >> 2;
            const int row = q * 32 + lane;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            for (int t = 0; t < T; ++t) {
                for (int k = 0; k < NCH; ++k, ++cnt) {
                    const uint32_t b = cnt & 1;
                    float cprev[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t > 0) c4 = *reinterpret_cast<const float4*>(c_out + wtl_off(t - 1, ntiles, tile, NCH, k, g, 4, j / 4, row, 4));
                        cprev[j] = c4.x; cprev[j + 1] = c4.y; cprev[j + 2] = c4.z; cprev[j + 3] = c4.w;
                    }
                    uint4 gxv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) gxv[i] = *reinterpret_cast<const uint4*>(gx + wtl_off(t, ntiles, tile, NCH, k, g, 8, i, row, 8));
                    mbar_wait(&S.d_full[b], (cnt >> 1) & 1);
                    tc_fence_after();
                    float cn[16], hn[16];
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        uint32_t v[32];
                        tmem_ld32(tmem + b * kN + lane_base + (4 * g + 2 * pr) * 16, v);
                        uint32_t act[16];                               // activated gates of the two granules, fp16x2
#pragma unroll
                        for (int gi = 0; gi < 2; ++gi) {
                            const int qq = 2 * pr + gi;                 // granule within the 16-unit group
                            float pre[16];
#pragma unroll
                            for (int e = 0; e < 16; e += 2) {
                                const uint32_t w = reinterpret_cast<const uint32_t*>(gxv)[qq * 8 + e / 2];
                                pre[e] = __uint_as_float(v[gi * 16 + e]) + val_lo(w);
                                pre[e + 1] = __uint_as_float(v[gi * 16 + e + 1]) + val_hi(w);
                            }
                            float a[16];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int j = qq * 4 + u;
                                const float gi_ = sigmoid_apx(pre[u]), gf = sigmoid_apx(pre[4 + u]);
                                const float gg = tanh_apx(pre[8 + u]), go = sigmoid_apx(pre[12 + u]);
                                cn[j] = fmaf(gf, cprev[j], gi_ * gg);
                                hn[j] = go * tanh_apx(cn[j]);
                                a[u] = gi_; a[4 + u] = gf; a[8 + u] = gg; a[12 + u] = go;
                            }
#pragma unroll
                            for (int e = 0; e < 8; ++e) act[gi * 8 + e] = pack_val(a[2 * e], a[2 * e + 1]);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *reinterpret_cast<uint4*>(gates_out + wtl_off(t, ntiles, tile, NCH, k, g, 8, pr * 4 + i, row, 8)) =
                                make_uint4(act[4 * i], act[4 * i + 1], act[4 * i + 2], act[4 * i + 3]);
                        // h of these 8 units: HBM (WTL) + the operand buffer of step t + 1 (K chunk 6 k + 2 g + pr)
                        const uint32_t p0 = pack_val(hn[pr * 8], hn[pr * 8 + 1]), p1 = pack_val(hn[pr * 8 + 2], hn[pr * 8 + 3]);
                        const uint32_t p2 = pack_val(hn[pr * 8 + 4], hn[pr * 8 + 5]), p3 = pack_val(hn[pr * 8 + 6], hn[pr * 8 + 7]);
                        st_shared_v4(S.h[t & 1] + (6 * k + 2 * g + pr) * kAChunk + row * 16, p0, p1, p2, p3);
                        *reinterpret_cast<uint4*>(h_out + wtl_off(t, ntiles, tile, NCH, k, g, 2, pr, row, 8)) = make_uint4(p0, p1, p2, p3);
                    }
                    tc_fence_before();
                    mbar_arrive(&S.d_empty[b]);
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(c_out + wtl_off(t, ntiles, tile, NCH, k, g, 4, j / 4, row, 4)) = make_float4(cn[j], cn[j + 1], cn[j + 2], cn[j + 3]);
                    fence_proxy_async_smem();
                    mbar_arrive(&S.h_ready);
                }
            }
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWtTmaWarp) { tc_fence_after(); tmem_free_all(tmem); }
}

// ---------------------------------------------------------------------------------------------------------------
// BPTT of one layer
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
struct WbSmem {
    static constexpr int kSlice = 2 * 48 * NCH * 16;     // one K16 slice of W_hh^T for a task: [2][H rows][8] fp16
    static constexpr int kGroup = 6;                     // slices per ring stage (12 per task: K = 192 gate columns)
    static constexpr int kStageBytes = kGroup * kSlice;
    static constexpr int kStages = kWtRingBytes / kStageBytes;
    alignas(128) unsigned char dg[2][24 * kAChunk];      // d(gates) of one task (192 columns), double-buffered
    alignas(128) unsigned char ring[kStages][kStageBytes];
    alignas(8) uint64_t ring_full[kStages], ring_empty[kStages];
    uint64_t dg_ready[2], dg_free[2], r_full;
    uint32_t tmem_base;
};

template <int NCH>
__global__ void __launch_bounds__(kWtThreads, 1)
lstm_wide_bwd_kernel(const __half* __restrict__ gates,             // WTL E=64 (forward)
                     const float* __restrict__ cstate,             // WTL E=16 (forward)
                     const float* __restrict__ dh_in,              // WTL E=16: gradient arriving from above
                     const unsigned char* __restrict__ wimg,       // [NCH tasks][12 slices][2][H][8] fp16: W_hh, K = the task's gate columns
                     __half* __restrict__ dg_out,                  // WTL E=64
                     float* __restrict__ dc_ws,                    // [grid][NCH][3][128][16] running d(cell)
                     int T, int ntiles) {
    using SM = WbSmem<NCH>;
    constexpr int kH = 48 * NCH, kGroup = SM::kGroup, kStages = SM::kStages;
    constexpr uint32_t kIdescR = make_idesc(kH, kFmtVal, kFmtVal);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.ring_full[s], 1); mbar_init(&S.ring_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&S.dg_ready[b], 12 * 32); mbar_init(&S.dg_free[b], 1); }
        mbar_init(&S.r_full, 1);
        fence_mbar_init();
    }
    if (warp == kWtTmaWarp) tmem_alloc_all(&S.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, S.tmem_base, 0);
    // D_R[0] at column 0, D_R[1] at column 256 (kH <= 192)
    uint32_t sc = 0, cnt = 0, rf = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (warp == kWtTmaWarp) {
            if (lane == 0)
                for (int i = 0; i < T; ++i)
                    for (int j = 0; j < NCH * 12 / kGroup; ++j, ++sc) {
                        const uint32_t s = sc % kStages, u = sc / kStages;
                        mbar_wait(&S.ring_empty[s], (u & 1) ^ 1);
                        mbar_arrive_expect_tx(&S.ring_full[s], SM::kStageBytes);
                        bulk_load(S.ring[s], wimg + (size_t)j * SM::kStageBytes, SM::kStageBytes, &S.ring_full[s]);
                    }
        } else if (warp == kWtMmaWarp) {
            const bool leader = elect_one();
            const uint64_t d_dg[2] = {umma_desc(smem_u32(S.dg[0]), kAChunk, 128), umma_desc(smem_u32(S.dg[1]), kAChunk, 128)};
            for (int i = 0; i < T; ++i) {
                const uint32_t tr = tmem + (i & 1) * 256;
                for (int k = 0; k < NCH; ++k, ++cnt) {
                    const uint32_t b = cnt & 1;
                    mbar_wait(&S.dg_ready[b], (cnt >> 1) & 1);
                    tc_fence_after();
                    for (int gr = 0; gr < 12 / kGroup; ++gr, ++sc) {
                        const uint32_t s = sc % kStages, u = sc / kStages;
                        mbar_wait(&S.ring_full[s], u & 1);
                        tc_fence_after();
                        const uint64_t d_w = umma_desc(smem_u32(S.ring[s]), kH * 16, 128);
#pragma unroll
                        for (int j = 0; j < kGroup; ++j) {
                            const int ks = gr * kGroup + j;
                            if (leader) umma_bf16_i(tr, desc_adv(d_dg[b], 2 * ks * kAChunk), desc_adv(d_w, j * SM::kSlice), kIdescR,
                                                    (k == 0 && ks == 0) ? 0u : 1u);
                        }
                        if (leader) umma_commit(&S.ring_empty[s]);
                    }
                    if (leader) umma_commit(&S.dg_free[b]);
                }
                if (leader) umma_commit(&S.r_full);
            }
        } else {
            const int q = warp & 3, g = warp >> 2;
            const int row = q * 32 + lane;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            float* dcw = dc_ws + (((size_t)blockIdx.x * NCH * 3) * kRows) * 16;      // [k][g][4 pieces][128][4]
            for (int i = 0; i < T; ++i) {
                const int t = T - 1 - i;
                for (int k = 0; k < NCH; ++k, ++cnt) {
                    const uint32_t b = cnt & 1;
                    float* dcp = dcw + (((size_t)k * 3 + g) * 4 * kRows + row) * 4;     // piece p at dcp + p * 128 * 4
                    // ---- loads that do not depend on the tensor pipe -------------------------------------------------
                    uint4 gv[8];      // (fetching these one task ahead, as the forward does with gx, costs 230 B of spills here and is slower)
#pragma unroll
                    for (int e = 0; e < 8; ++e) gv[e] = *reinterpret_cast<const uint4*>(gates + wtl_off(t, ntiles, tile, NCH, k, g, 8, e, row, 8));
                    float ct[16], cp[16], dh[16], dc[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(cstate + wtl_off(t, ntiles, tile, NCH, k, g, 4, j / 4, row, 4));
                        ct[j] = a.x; ct[j + 1] = a.y; ct[j + 2] = a.z; ct[j + 3] = a.w;
                        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t > 0) p = *reinterpret_cast<const float4*>(cstate + wtl_off(t - 1, ntiles, tile, NCH, k, g, 4, j / 4, row, 4));
                        cp[j] = p.x; cp[j + 1] = p.y; cp[j + 2] = p.z; cp[j + 3] = p.w;
                        const float4 d = *reinterpret_cast<const float4*>(dh_in + wtl_off(t, ntiles, tile, NCH, k, g, 4, j / 4, row, 4));
                        dh[j] = d.x; dh[j + 1] = d.y; dh[j + 2] = d.z; dh[j + 3] = d.w;
                        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i > 0) c4 = *reinterpret_cast<const float4*>(dcp + (j / 4) * kRows * 4);
                        dc[j] = c4.x; dc[j + 1] = c4.y; dc[j + 2] = c4.z; dc[j + 3] = c4.w;
                    }
                    // ---- dh_rec of this step = D_R of the previous iteration (complete when its last task has committed) ----
                    if (i >= 1) {
                        if (k == 0) { mbar_wait(&S.r_full, rf & 1); ++rf; tc_fence_after(); }
                        uint32_t r[16];
                        wt_tmem_ld16(tmem + ((i - 1) & 1) * 256 + lane_base + 48 * k + 16 * g, r);
#pragma unroll
                        for (int j = 0; j < 16; ++j) dh[j] += __uint_as_float(r[j]);
                    }
                    uint32_t out[32];                                   // d(gates) fp16x2, TMEM column order
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        float pg_[16];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = qq * 4 + u;
                            const uint32_t* gw = reinterpret_cast<const uint32_t*>(gv) + qq * 8;
                            const uint32_t wi = gw[u / 2], wf = gw[2 + u / 2], wg = gw[4 + u / 2], wo = gw[6 + u / 2];
                            const float gi_ = (u & 1) ? val_hi(wi) : val_lo(wi), gf = (u & 1) ? val_hi(wf) : val_lo(wf);
                            const float gg = (u & 1) ? val_hi(wg) : val_lo(wg), go = (u & 1) ? val_hi(wo) : val_lo(wo);
                            const float tcv = tanh_apx(ct[j]);
                            const float d_o = dh[j] * tcv;
                            const float dct = fmaf(dh[j] * go, 1.0f - tcv * tcv, dc[j]);
                            dc[j] = dct * gf;
                            pg_[u] = dct * gg * gi_ * (1.0f - gi_);
                            pg_[4 + u] = dct * cp[j] * gf * (1.0f - gf);
                            pg_[8 + u] = dct * gi_ * (1.0f - gg * gg);
                            pg_[12 + u] = d_o * go * (1.0f - go);
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) out[qq * 8 + e] = pack_val(pg_[2 * e], pg_[2 * e + 1]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(dcp + (j / 4) * kRows * 4) = make_float4(dc[j], dc[j + 1], dc[j + 2], dc[j + 3]);
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        *reinterpret_cast<uint4*>(dg_out + wtl_off(t, ntiles, tile, NCH, k, g, 8, e, row, 8)) =
                            make_uint4(out[4 * e], out[4 * e + 1], out[4 * e + 2], out[4 * e + 3]);
                    // A operand of R: columns n = (4 g + qq) * 16 + e -> K chunks (4 g + qq) * 2 and + 1 of the task's 24
                    mbar_wait(&S.dg_free[b], ((cnt >> 1) & 1) ^ 1);     // the MMAs that read this buffer two tasks ago are done
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        unsigned char* d0 = S.dg[b] + ((4 * g + qq) * 2) * kAChunk + row * 16;
                        st_shared_v4(d0, out[qq * 8], out[qq * 8 + 1], out[qq * 8 + 2], out[qq * 8 + 3]);
                        st_shared_v4(d0 + kAChunk, out[qq * 8 + 4], out[qq * 8 + 5], out[qq * 8 + 6], out[qq * 8 + 7]);
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();
                    mbar_arrive(&S.dg_ready[b]);
                }
            }
            // the last step's D_R is not needed; consume its phase so the next tile starts aligned
            mbar_wait(&S.r_full, rf & 1); ++rf;
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWtTmaWarp) { tc_fence_after(); tmem_free_all(tmem); }
}

static int wt_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <int NCH>
static int launch_wide_fwd(const void* gx, const void* wimg, void* gates, void* h, float* c, int T, int ntiles, cudaStream_t st) {
    const size_t smem = sizeof(WfSmem<NCH>);
    cudaError_t e = cudaFuncSetAttribute(lstm_wide_fwd_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "na_lstm_wide_fwd_train: shared memory opt-in failed (%s)", cudaGetErrorString(e));
    const int grid = ntiles < wt_sms() ? ntiles : wt_sms();
    lstm_wide_fwd_kernel<NCH><<<grid, kWtThreads, smem, st>>>(reinterpret_cast<const __half*>(gx), reinterpret_cast<const unsigned char*>(wimg),
                                                              reinterpret_cast<__half*>(gates), reinterpret_cast<__half*>(h), c, T, ntiles);
    return 0;
}

template <int NCH>
static int launch_wide_bwd(const void* gates, const float* c, const float* dh, const void* wimg, void* dg, float* ws, int T, int ntiles,
                           cudaStream_t st) {
    const size_t smem = sizeof(WbSmem<NCH>);
    cudaError_t e = cudaFuncSetAttribute(lstm_wide_bwd_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "na_lstm_wide_bwd: shared memory opt-in failed (%s)", cudaGetErrorString(e));
    const int grid = ntiles < wt_sms() ? ntiles : wt_sms();
    lstm_wide_bwd_kernel<NCH><<<grid, kWtThreads, smem, st>>>(reinterpret_cast<const __half*>(gates), c, dh,
                                                              reinterpret_cast<const unsigned char*>(wimg), reinterpret_cast<__half*>(dg), ws, T,
                                                              ntiles);
    return 0;
}

}  // namespace tc
}  // namespace na

extern "C" int64_t na_wide_train_ws_floats(int64_t H) {
    return (int64_t)na::tc::wt_sms() * (H / 48) * 3 * na::tc::kRows * 16;
}

extern "C" int na_lstm_wide_fwd_train(const void* gx, const void* w_image, void* gates, void* h, float* c, int64_t T, int64_t Bp,
                                      int64_t H, na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(H == 96 || H == 144 || H == 192, NA_EUNSUPPORTED, "na_lstm_wide_fwd_train: hidden_size=%lld (96, 144 or 192)", (long long)H);
    NA_REQUIRE(T >= 1 && T < (1 << 20) && Bp >= tc::kRows && Bp % tc::kRows == 0, NA_EINVAL,
               "na_lstm_wide_fwd_train: bad shape T=%lld Bp=%lld (Bp must be a multiple of 128)", (long long)T, (long long)Bp);
    NA_REQUIRE_PTR(gx); NA_REQUIRE_PTR(w_image); NA_REQUIRE_PTR(gates); NA_REQUIRE_PTR(h); NA_REQUIRE_PTR(c);
    const int ntiles = (int)(Bp / tc::kRows);
    int rc;
    if (H == 96) rc = tc::launch_wide_fwd<2>(gx, w_image, gates, h, c, (int)T, ntiles, as_stream(stream));
    else if (H == 144) rc = tc::launch_wide_fwd<3>(gx, w_image, gates, h, c, (int)T, ntiles, as_stream(stream));
    else rc = tc::launch_wide_fwd<4>(gx, w_image, gates, h, c, (int)T, ntiles, as_stream(stream));
    if (rc) return rc;
    count_launch();
    return check_launch("na_lstm_wide_fwd_train");
}

extern "C" int na_lstm_wide_bwd(const void* gates, const float* c, const float* dh_in, const void* w_image, void* dg, float* workspace,
                                int64_t T, int64_t Bp, int64_t H, na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(H == 96 || H == 144 || H == 192, NA_EUNSUPPORTED, "na_lstm_wide_bwd: hidden_size=%lld (96, 144 or 192)", (long long)H);
    NA_REQUIRE(T >= 1 && T < (1 << 20) && Bp >= tc::kRows && Bp % tc::kRows == 0, NA_EINVAL,
               "na_lstm_wide_bwd: bad shape T=%lld Bp=%lld (Bp must be a multiple of 128)", (long long)T, (long long)Bp);
    NA_REQUIRE_PTR(gates); NA_REQUIRE_PTR(c); NA_REQUIRE_PTR(dh_in); NA_REQUIRE_PTR(w_image); NA_REQUIRE_PTR(dg); NA_REQUIRE_PTR(workspace);
    const int ntiles = (int)(Bp / tc::kRows);
    int rc;
    if (H == 96) rc = tc::launch_wide_bwd<2>(gates, c, dh_in, w_image, dg, workspace, (int)T, ntiles, as_stream(stream));
    else if (H == 144) rc = tc::launch_wide_bwd<3>(gates, c, dh_in, w_image, dg, workspace, (int)T, ntiles, as_stream(stream));
    else rc = tc::launch_wide_bwd<4>(gates, c, dh_in, w_image, dg, workspace, (int)T, ntiles, as_stream(stream));
    if (rc) return rc;
    count_launch();
    return check_launch("na_lstm_wide_bwd");
}
