// Tensor-core tier, TRAINING: 2-layer LSTM forward that saves what the backward needs, and a
// fused BPTT + weight-gradient kernel per layer, both on tcgen05 / TMEM / TMA (bf16 operands, fp32
// accumulate / cell state / gradients of the cell state).
//
// Layouts
//   TMP  fp32 [T][Bp][48]                      (c, fp32 copy of h1 for the head kernels, dh, din)
//   TCL  bf16 [T][Bp/128][F/8][128][8]         "tile-chunk": the [128 x F] slab of a tile at step t is
//        contiguous AND already in the UMMA core-matrix layout, so it is one TMA bulk copy into
//        an operand buffer (x: F = 8, identical to TMP bf16; h0, h0-after-dropout, h1: F = 48)
//   mask u8  [T][Bp][48]                        inter-layer dropout keep-mask (lstm_eeg_model.py:21)
//
// Backward of one layer, per step t (descending) and 128-window tile:
//   G: gates = [in_t | 1 | h_{t-1}] . B            (recompute; same MMA as the forward)       D_G
//   E: thread = window: activations, d(gates) from dh_t = dh_out_t + dh_rec, c_t, c_{t-1};
//      d(gates) -> shared memory (bf16) in the A-operand layout                                 DG
//   R: [din | dh_rec] = DG . [W_ih | W_hh]          (K = 192)                                   D_R
//   W: dW += DG^T . [in_t | 1 | h_{t-1}]            (K = 128 windows; DG and the operand buffer are
//      re-used as MN-major operands -- no transposes, no dgates round trip through HBM)          D_W
//   dW (+ db via the ones column) accumulates in TMEM in fp32 over ALL steps and tiles of the CTA;
//   per-CTA partials are reduced in a fixed order by a second tiny kernel (deterministic).
// Pipeline of a step (round 2, second pass): only R is on the critical path d(gates) -> dh_rec -> d(gates).  The MMA warp issues
// R(t+1) as soon as d(gates) of step t+1 are in shared memory (dg_ready) and commits r_full; W(t+1) follows and releases the
// stage (act_free) and d(gates) (w_done); G(t-1) is issued once the epilogue has drained G(t) out of TMEM (g_free) and runs
// under the epilogue of step t.  The [in | 1 | h] stages form a 2- or 3-deep TMA ring (bwd_stages); with the fused head
// backward the epilogue threads read h_t out of the stage of step t+1, which is therefore freed by the W commit AND 384 thread
// arrivals.  Half tiles additionally evaluate the activations before waiting for R.
#include "na_tc_common.cuh"

namespace na {
int reduce_partials(const float* partial, float* out, int nchunks, int64_t n, cudaStream_t st);   // na_reduce.cu
namespace tc {

constexpr int kTrainThreads = 576;     // forward: 8 + 8 epilogue warps (2 per TMEM quarter and layer), MMA warp, TMA warp
constexpr int kBwdEG = 3;               // backward: epilogue warps per TMEM lane quarter (each: 6 / kBwdEG unit blocks)
constexpr int kBwdUB = 6 / kBwdEG;      // 8-unit blocks per epilogue thread
constexpr int kBwdU = 8 * kBwdUB;       // units per epilogue thread
constexpr int kBwdEW = 4 * kBwdEG;      // epilogue warps; warp kBwdEW = MMA issuer, kBwdEW + 1 = TMA producer
constexpr int kBwdThreads = (kBwdEW + 2) * 32;
constexpr int kXStagesT = 4;
// backward: [in_t | 1 | h_{t-1}] stages in flight.  The gate recompute of step t-1 is issued early in step t; with two stages its
// TMA load, which has to wait for W(t+1), was 35 % of the MMA warp's time in half-tile mode (three stages: -15 % / -21 %
// kernel time).  Full tiles of layer 1 are bound by instruction issue instead and lose L1 capacity to a third stage.
constexpr int bwd_stages(int KI, bool half) { return (KI == 48 && !half) ? 2 : 3; }

__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }

// cell update of 8 units; returns fp32 h and leaves c updated.
__device__ __forceinline__ void cell_block_f(const uint32_t (&v)[32], float* c, float (&h)[8]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float gi = sigmoid_apx(__uint_as_float(v[u]));
        const float gf = sigmoid_apx(__uint_as_float(v[8 + u]));
        const float gg = tanh_apx(__uint_as_float(v[16 + u]));
        const float go = sigmoid_apx(__uint_as_float(v[24 + u]));
        c[u] = fmaf(gf, c[u], gi * gg);
        h[u] = go * tanh_apx(c[u]);
    }
}

__device__ __forceinline__ void st_global_v4f(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// TCL32: fp32 tile-chunk layout [T][Bp/128][12][128][4] for per-window-row tensors the epilogue threads
// (thread = window) read and write 16 bytes at a time: lanes of a warp touch 32 consecutive 16-B pieces
// (fully coalesced) instead of 16 B out of every 192-B row.  Used for c0, c1 and din.
__device__ __forceinline__ int64_t tcl32_off(int t, int ntiles, int tile, int chunk, int row) {
    return (((((int64_t)t * ntiles + tile) * 12 + chunk) * kRows) + row) * 4;
}

// =================================================================================================
// forward (training): x -> h0, h0 after dropout, c0, h1 (bf16 + fp32), c1
// =================================================================================================
struct FwdSmem {
    alignas(128) unsigned char b0[8 * kBChunk];
    alignas(128) unsigned char b1[14 * kBChunk];
    alignas(128) unsigned char x[kXStagesT][2 * kAChunk];     // [x | ones]
    alignas(128) unsigned char h0[2][6 * kAChunk];            // raw h0_t        (layer-0 recurrence)
    alignas(128) unsigned char h0d[2][6 * kAChunk];           // dropped h0_t    (layer-1 input)
    alignas(128) unsigned char h1[6 * kAChunk];
    alignas(128) unsigned char onez[2 * kAChunk];
    alignas(8) uint64_t x_full[kXStagesT], x_empty[kXStagesT];
    uint64_t d0_full, d1_full, h0_ready[2], h0_free[2], h1_ready;
    uint32_t tmem_base;
    float wa[kH];                                             // attention weights (fused pooling)
    float ba;
    float score_part[2][2][kRows];                            // per-step partial scores of the two half-row warps
};

__global__ void __launch_bounds__(kTrainThreads, 1)
lstm2_fwd_train_bf16_kernel(const __nv_bfloat16* __restrict__ x, const unsigned char* __restrict__ packed,
                            const unsigned char* __restrict__ mask, uint64_t seed, uint32_t thresh16, float drop_scale,
                            __nv_bfloat16* __restrict__ h0_out, __nv_bfloat16* __restrict__ h0d_out,
                            float* __restrict__ c0_out, __nv_bfloat16* __restrict__ h1_out,
                            float* __restrict__ h1f_out, float* __restrict__ c1_out,
                            const float* __restrict__ attn_w, const float* __restrict__ attn_b,   // fused attention pooling
                            float* __restrict__ zpool_out, float* __restrict__ stats_out, int64_t B,   // (all NULL: off)
                            int T, int64_t Bp, int ntiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform (role branches = uniform control flow)
    {
        const uint4* src = reinterpret_cast<const uint4*>(packed);
        uint4* dst = reinterpret_cast<uint4*>(S.b0);
        for (int i = tid; i < 22 * kBChunk / 16; i += kTrainThreads) dst[i] = src[i];
        const uint4 ones = make_uint4(kValOnes2, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < kRows; i += kTrainThreads) {
#pragma unroll
            for (int s = 0; s < kXStagesT; ++s) reinterpret_cast<uint4*>(S.x[s] + kAChunk)[i] = ones;
            reinterpret_cast<uint4*>(S.onez)[i] = ones;
            reinterpret_cast<uint4*>(S.onez + kAChunk)[i] = zero;
        }
        if (tid == 0) {
            for (int s = 0; s < kXStagesT; ++s) { mbar_init(&S.x_full[s], 1); mbar_init(&S.x_empty[s], 1); }
            mbar_init(&S.d0_full, 1); mbar_init(&S.d1_full, 1);
            mbar_init(&S.h0_ready[0], 256); mbar_init(&S.h0_ready[1], 256);
            mbar_init(&S.h0_free[0], 1); mbar_init(&S.h0_free[1], 1);
            mbar_init(&S.h1_ready, 256);
            fence_mbar_init();
        }
        if (zpool_out != nullptr) {
            for (int i = tid; i < kH; i += kTrainThreads) S.wa[i] = attn_w[i];
            if (tid == 0) S.ba = attn_b[0];
        }
        if (warp == 17) tmem_alloc_all(&S.tmem_base);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = S.tmem_base, tmem_d0 = tmem, tmem_d1 = tmem + kN;
    const bool drop = mask != nullptr || thresh16 < 65536u;
    const bool pool = zpool_out != nullptr;      // explicit mask tensor, or in-kernel counter-based RNG

    int n0 = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n0 += T) {
        const int64_t b0 = (int64_t)tile * kRows;
        {
            uint4* z0 = reinterpret_cast<uint4*>(S.h0[(n0 + 1) & 1]);
            uint4* z1 = reinterpret_cast<uint4*>(S.h1);
            for (int i = tid; i < 6 * kAChunk / 16; i += kTrainThreads) { z0[i] = make_uint4(0, 0, 0, 0); z1[i] = make_uint4(0, 0, 0, 0); }
            fence_proxy_async_smem();
            __syncthreads();
        }
        if (warp == 17) {
            if (lane == 0)
                for (int t = 0; t < T; ++t) {
                    const int n = n0 + t, s = n % kXStagesT, u = n / kXStagesT;
                    mbar_wait(&S.x_empty[s], (u & 1) ^ 1);
                    mbar_arrive_expect_tx(&S.x_full[s], kAChunk);
                    bulk_load(S.x[s], x + ((int64_t)t * Bp + b0) * 8, kAChunk, &S.x_full[s]);
                }
        } else if (warp == 16) {
            {   // whole warp, warp-uniform control flow; one elected lane issues (see na_decoder_wide.cu)
                const bool leader = elect_one();
                const uint64_t d_b0 = umma_desc(smem_u32(S.b0), kBChunk, 128), d_b1 = umma_desc(smem_u32(S.b1), kBChunk, 128);
                const uint64_t d_x0 = umma_desc(smem_u32(S.x[0]), kAChunk, 128);
                const uint64_t d_h0[2] = {umma_desc(smem_u32(S.h0[0]), kAChunk, 128), umma_desc(smem_u32(S.h0[1]), kAChunk, 128)};
                const uint64_t d_h0d[2] = {umma_desc(smem_u32(S.h0d[0]), kAChunk, 128), umma_desc(smem_u32(S.h0d[1]), kAChunk, 128)};
                const uint64_t d_h1 = umma_desc(smem_u32(S.h1), kAChunk, 128), d_onez = umma_desc(smem_u32(S.onez), kAChunk, 128);
                for (int t = 0; t <= T; ++t) {
                    if (t < T) {
                        const int n = n0 + t, s = n % kXStagesT, u = n / kXStagesT;
                        mbar_wait(&S.x_full[s], u & 1);
                        mbar_wait(&S.h0_ready[(n + 1) & 1], ((n - 1) >> 1) & 1);
                        tc_fence_after();
                        const uint64_t hprev = d_h0[(n + 1) & 1];
                        if (leader) umma_bf16(tmem_d0, desc_adv(d_x0, s * 2 * kAChunk), d_b0, 0u);
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            if (leader) umma_bf16(tmem_d0, desc_adv(hprev, 2 * i * kAChunk), desc_adv(d_b0, (2 + 2 * i) * kBChunk), 1u);
                        if (leader) umma_commit(&S.x_empty[s]);
                        if (leader) umma_commit(&S.d0_full);
                    }
                    if (t >= 1) {
                        const int m = n0 + t - 1;
                        mbar_wait(&S.h0_ready[m & 1], (m >> 1) & 1);
                        mbar_wait(&S.h1_ready, (m - 1) & 1);
                        tc_fence_after();
                        const uint64_t hin = drop ? d_h0d[m & 1] : d_h0[m & 1];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            if (leader) umma_bf16(tmem_d1, desc_adv(hin, 2 * i * kAChunk), desc_adv(d_b1, 2 * i * kBChunk), i == 0 ? 0u : 1u);
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            if (leader) umma_bf16(tmem_d1, desc_adv(d_h1, 2 * i * kAChunk), desc_adv(d_b1, (6 + 2 * i) * kBChunk), 1u);
                        if (leader) umma_bf16(tmem_d1, d_onez, desc_adv(d_b1, 12 * kBChunk), 1u);
                        if (leader) umma_commit(&S.d1_full);
                        if (leader) umma_commit(&S.h0_free[m & 1]);
                    }
                }
            }
        } else if (warp < 8) {
            // ---- layer-0 epilogue: thread = window x half of the units (3 blocks of 8) -----------------
            const int q = warp & 3, hf = warp >> 2;
            const int row = q * 32 + lane;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            float c[24];
#pragma unroll
            for (int j = 0; j < 24; ++j) c[j] = 0.f;
            for (int t = 0; t < T; ++t) {
                const int n = n0 + t;
                const int64_t grow = (int64_t)t * Bp + b0 + row;                 // TMP row
                const int64_t tcl = ((int64_t)t * ntiles + tile) * 6 * (kAChunk / 2) + row * 8;   // TCL element offset of chunk 0
                uint32_t keep[3];
                if (drop) {
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) {
                        const int blk = hf * 3 + bb;
                        keep[bb] = mask ? mask_keep8(*reinterpret_cast<const uint2*>(mask + grow * kH + blk * 8))
                                        : dropout_keep8(seed, grow, blk, thresh16);
                    }
                }
                mbar_wait(&S.d0_full, n & 1);
                mbar_wait(&S.h0_free[n & 1], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                unsigned char* dst = S.h0[n & 1] + row * 16;
                unsigned char* dstd = S.h0d[n & 1] + row * 16;
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    const int blk = hf * 3 + bb;
                    uint32_t v[32];
                    float h[8];
                    tmem_ld32(tmem_d0 + lane_base + blk * 32, v);
                    cell_block_f(v, c + bb * 8, h);
                    const uint32_t p0 = pack_val(h[0], h[1]), p1 = pack_val(h[2], h[3]);
                    const uint32_t p2 = pack_val(h[4], h[5]), p3 = pack_val(h[6], h[7]);
                    st_shared_v4(dst + blk * kAChunk, p0, p1, p2, p3);
                    *reinterpret_cast<uint4*>(h0_out + tcl + blk * (kAChunk / 2)) = make_uint4(p0, p1, p2, p3);
                    st_global_v4f(c0_out + tcl32_off(t, ntiles, tile, 2 * blk, row), c[bb * 8], c[bb * 8 + 1], c[bb * 8 + 2], c[bb * 8 + 3]);
                    st_global_v4f(c0_out + tcl32_off(t, ntiles, tile, 2 * blk + 1, row), c[bb * 8 + 4], c[bb * 8 + 5], c[bb * 8 + 6], c[bb * 8 + 7]);
                    if (drop) {
                        float hd[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) hd[u] = ((keep[bb] >> u) & 1u) ? h[u] * drop_scale : 0.f;
                        const uint32_t q0 = pack_val(hd[0], hd[1]), q1 = pack_val(hd[2], hd[3]);
                        const uint32_t q2 = pack_val(hd[4], hd[5]), q3 = pack_val(hd[6], hd[7]);
                        st_shared_v4(dstd + blk * kAChunk, q0, q1, q2, q3);
                        *reinterpret_cast<uint4*>(h0d_out + tcl + blk * (kAChunk / 2)) = make_uint4(q0, q1, q2, q3);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&S.h0_ready[n & 1]);
            }
        } else {
            // ---- layer-1 epilogue --------------------------------------------------------------------
            const int q = (warp - 8) & 3, hf = (warp - 8) >> 2;
            const int row = q * 32 + lane;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            float c[24], z[24];
#pragma unroll
            for (int j = 0; j < 24; ++j) { c[j] = 0.f; z[j] = 0.f; }
            float mx = -INFINITY, l = 0.f;
            for (int t = 0; t < T; ++t) {
                const int m = n0 + t;
                const int64_t grow = (int64_t)t * Bp + b0 + row;
                const int64_t tcl = ((int64_t)t * ntiles + tile) * 6 * (kAChunk / 2) + row * 8;
                mbar_wait(&S.d1_full, m & 1);
                tc_fence_after();
                unsigned char* dst = S.h1 + row * 16;
                uint32_t hb[12];
                float score = 0.f;
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    const int blk = hf * 3 + bb;
                    uint32_t v[32];
                    float h[8];
                    tmem_ld32(tmem_d1 + lane_base + blk * 32, v);
                    cell_block_f(v, c + bb * 8, h);
                    const uint32_t p0 = pack_val(h[0], h[1]), p1 = pack_val(h[2], h[3]);
                    const uint32_t p2 = pack_val(h[4], h[5]), p3 = pack_val(h[6], h[7]);
                    st_shared_v4(dst + blk * kAChunk, p0, p1, p2, p3);
                    *reinterpret_cast<uint4*>(h1_out + tcl + blk * (kAChunk / 2)) = make_uint4(p0, p1, p2, p3);
                    if (pool) {       // attention score on the fp16-rounded h (what the backward re-reads)
                        hb[bb * 4] = p0; hb[bb * 4 + 1] = p1; hb[bb * 4 + 2] = p2; hb[bb * 4 + 3] = p3;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            score = fmaf(val_lo(hb[bb * 4 + u]), S.wa[blk * 8 + 2 * u], score);
                            score = fmaf(val_hi(hb[bb * 4 + u]), S.wa[blk * 8 + 2 * u + 1], score);
                        }
                    }
                    if (h1f_out) {
                        st_global_v4f(h1f_out + grow * kH + blk * 8, h[0], h[1], h[2], h[3]);
                        st_global_v4f(h1f_out + grow * kH + blk * 8 + 4, h[4], h[5], h[6], h[7]);
                    }
                    st_global_v4f(c1_out + tcl32_off(t, ntiles, tile, 2 * blk, row), c[bb * 8], c[bb * 8 + 1], c[bb * 8 + 2], c[bb * 8 + 3]);
                    st_global_v4f(c1_out + tcl32_off(t, ntiles, tile, 2 * blk + 1, row), c[bb * 8 + 4], c[bb * 8 + 5], c[bb * 8 + 6], c[bb * 8 + 7]);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&S.h1_ready);
                if (pool) {
                    // online softmax over time (lstm_eeg_model.py:35-37); the two half-row warps add their
                    // partial scores in a fixed order, so both carry bit-identical (max, sum)
                    S.score_part[t & 1][hf][row] = score;
                    named_bar_sync(1 + q, 64);
                    score = S.score_part[t & 1][0][row] + S.score_part[t & 1][1][row] + S.ba;
                    if (score > mx) {
                        const float sc = __expf(mx - score);
                        l *= sc;
#pragma unroll
                        for (int j = 0; j < 24; ++j) z[j] *= sc;
                        mx = score;
                    }
                    const float e = __expf(score - mx);
                    l += e;
#pragma unroll
                    for (int u = 0; u < 12; ++u) {
                        z[2 * u] = fmaf(e, val_lo(hb[u]), z[2 * u]);
                        z[2 * u + 1] = fmaf(e, val_hi(hb[u]), z[2 * u + 1]);
                    }
                }
            }
            if (pool && b0 + row < B) {
                const float inv_l = 1.0f / l;
                float* zo = zpool_out + (b0 + row) * kH + hf * 24;
#pragma unroll
                for (int j = 0; j < 24; j += 4) st_global_v4f(zo + j, z[j] * inv_l, z[j + 1] * inv_l, z[j + 2] * inv_l, z[j + 3] * inv_l);
                if (hf == 0) { stats_out[2 * (b0 + row)] = mx; stats_out[2 * (b0 + row) + 1] = l; }
            }
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) { tc_fence_after(); tmem_free_all(tmem); }
}

// =================================================================================================
// backward of one layer (KI = 8: layer 0, no din;  KI = 48: layer 1, din with the dropout mask)
// =================================================================================================
template <int KI>
struct BwdCfg {
    static constexpr int kInChunks = KI / 8;
    static constexpr int kStageChunks = KI == 8 ? 8 : 14;       // L0: x|ones|h x6   L1: in x6|h x6|ones|zero
    static constexpr int kHprevChunk = KI == 8 ? 2 : 6;         // first h_{t-1} chunk inside the stage
    static constexpr int kOnesChunk = KI == 8 ? 1 : 12;
    static constexpr int kNW = kStageChunks * 8;                // 64 / 112 features of the stage = dW columns
    static constexpr int kNR = KI == 8 ? 48 : 96;               // columns of D_R: [din (KI) |] dh_rec (48)
    static constexpr int kRecCol = KI == 8 ? 0 : 48;            // first dh_rec column of D_R
    static constexpr int kColG = 0, kColR = 192, kColW1 = KI == 8 ? 256 : 288, kColW2 = kColW1 + kNW;
    static_assert(kColW2 + kNW <= 512, "TMEM budget");
};

template <int KI, int kBwdStages>
struct BwdSmem {
    using C = BwdCfg<KI>;
    alignas(128) unsigned char bg[C::kStageChunks * kBChunk];          // forward B operand (recompute)
    alignas(128) unsigned char br[24 * C::kNR * 16];                   // [W_ih | W_hh]^T, K = 192 gate columns
    alignas(128) unsigned char act[kBwdStages][C::kStageChunks * kAChunk];      // [in_t | 1 | h_{t-1}] stages
    alignas(128) unsigned char dg[24 * kAChunk];                       // d(gates) bf16, A operand of R and W
    alignas(8) uint64_t act_full[kBwdStages], act_free[kBwdStages];
    uint64_t g_full, g_free, r_full, dg_ready, w_done;
    uint32_t tmem_base;
    // fused head backward (layer 1): attention weights, per-step exchange of the two half-row warps'
    // partial (score, dz.h), the per-tile dz.z exchange, and the final d(attn) reduction
    float wa[kH];
    float ba;
    float2 xch[2][kBwdEG][kRows];
    float xch0[kBwdEG][kRows];
    float red[kBwdEW][kBwdU + 1];
};

static_assert(sizeof(BwdSmem<48, 3>) + 1024 <= 232448 && sizeof(BwdSmem<8, 3>) + 1024 <= 232448, "shared-memory budget (227 KB)");

// HALF = half tiles (see lstm2_fwd_train_v2_kernel): 64 distinct windows per tile, rows 64..127 mirror rows 0..63.  Copy
// rp = q / 2 of a window takes unit block 2 hf + rp (8 units per thread instead of 16) and writes its d(gates) to BOTH row
// copies, so D_R is complete on every row; the weight-gradient MMAs then run over 64 window rows (K = 4 x 16).
template <int KI, bool HALF>
__global__ void __launch_bounds__(kBwdThreads, 1)
lstm_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ act_in,     // TCL [T][NT][KI/8][128][8]
                     const __nv_bfloat16* __restrict__ h,          // TCL [T][NT][6][128][8]  (this layer's raw h)
                     const float* __restrict__ cstate,             // TMP
                     const float* __restrict__ dh_out,             // TMP
                     const unsigned char* __restrict__ packed_g,   // forward B operand of this layer
                     const unsigned char* __restrict__ packed_r,   // [24 chunks][kNR][8] bf16
                     const __nv_bfloat16* __restrict__ zeros,      // >= 6*2048 B of zeros (h_{-1})
                     const unsigned char* __restrict__ mask, uint64_t seed, uint32_t thresh16,
                     float drop_scale,                             // KI == 48: dropout of this layer's input
                     float* __restrict__ din,                      // TMP, KI == 48 only
                     float* __restrict__ dw_partial,               // [grid][192][kNW]
                     int dh_chunked,                               // dh_out layout: 1 = TCL32 (din of the layer above), 0 = TMP
                     // fused head backward (KI == 48, dz != NULL): dh_out is not read; dh_t = alpha_t dz + ds_t w_a is
                     // rebuilt per step from dz [B,48], the softmax stats [B,2], the pooled z [B,48] and h (= h1)
                     const float* __restrict__ dz, const float* __restrict__ stats, const float* __restrict__ zpool,
                     const float* __restrict__ attn_w, const float* __restrict__ attn_b, int64_t B,
                     float* __restrict__ attn_partial,             // [grid][49]: d attn_w | d attn_b
                     int T, int64_t Bp, int ntiles, int64_t drop_stride) {
    using C = BwdCfg<KI>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int kBwdStages = bwd_stages(KI, HALF);
    BwdSmem<KI, kBwdStages>& S = *reinterpret_cast<BwdSmem<KI, kBwdStages>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform (role branches = uniform control flow)
    {
        const uint4* sg = reinterpret_cast<const uint4*>(packed_g);
        uint4* dgp = reinterpret_cast<uint4*>(S.bg);
        for (int i = tid; i < C::kStageChunks * kBChunk / 16; i += kBwdThreads) dgp[i] = sg[i];
        const uint4* sr = reinterpret_cast<const uint4*>(packed_r);
        uint4* drp = reinterpret_cast<uint4*>(S.br);
        for (int i = tid; i < 24 * C::kNR; i += kBwdThreads) drp[i] = sr[i];
        const uint4 ones = make_uint4(kValOnes2, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < kRows; i += kBwdThreads)
#pragma unroll
            for (int s = 0; s < kBwdStages; ++s) {
                reinterpret_cast<uint4*>(S.act[s] + C::kOnesChunk * kAChunk)[i] = ones;
                if constexpr (KI == 48) reinterpret_cast<uint4*>(S.act[s] + 13 * kAChunk)[i] = zero;
            }
        if (tid == 0) {
            // fused head backward: the epilogue threads read h_t out of the stage of step t+1, so a stage is free once the W
            // MMAs (one commit) AND every epilogue thread have finished with it
            const uint32_t free_cnt = (KI == 48 && dz != nullptr) ? 1 + 32 * kBwdEW : 1;
            for (int s = 0; s < kBwdStages; ++s) { mbar_init(&S.act_full[s], 1); mbar_init(&S.act_free[s], free_cnt); }
            mbar_init(&S.g_full, 1);
            mbar_init(&S.r_full, 1);
            mbar_init(&S.g_free, 32 * kBwdEW);
            mbar_init(&S.w_done, 1);
            mbar_init(&S.dg_ready, 32 * kBwdEW);
            fence_mbar_init();
        }
        if (dz != nullptr) {
            for (int i = tid; i < kH; i += kBwdThreads) S.wa[i] = attn_w[i];
            if (tid == 0) S.ba = attn_b[0];
        }
        if (warp == kBwdEW + 1) tmem_alloc_all(&S.tmem_base);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncthreads();
        tc_fence_after();
    }
    const bool head = (KI == 48) && dz != nullptr;
    float dwa[kBwdU], dba = 0.f;                  // d attn_w (this thread's units), d attn_b; epilogue warps only
#pragma unroll
    for (int j = 0; j < kBwdU; ++j) dwa[j] = 0.f;
    const uint32_t tmem = S.tmem_base;
    const uint32_t tm_g = tmem + C::kColG, tm_r = tmem + C::kColR, tm_w1 = tmem + C::kColW1, tm_w2 = tmem + C::kColW2;
    constexpr uint32_t kIdescR = make_idesc(C::kNR, kFmtGrad, kFmtVal);               // d(gates) x W^T, K-major x K-major
    constexpr uint32_t kIdescW = make_idesc(C::kNW, kFmtGrad, kFmtVal, true, true);   // d(gates)^T x act^T, MN-major x MN-major
    constexpr uint32_t kStageBytes = (C::kInChunks + 6) * kAChunk;

    uint32_t k0 = 0;                              // running step counter (stage / phase bookkeeping)
    uint32_t gphase = 0, rphase = 0, wphase = 0;  // phases of g_full / r_full / w_done consumed so far (epilogue warps)
    bool first_w = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t b0 = (int64_t)tile * kRows;
        if (warp == kBwdEW + 1) {
            // ================= TMA producer: [in_t | h_{t-1}] for t = T-1 .. 0 ==========================
            if (lane == 0)
                for (int i = 0; i < T; ++i) {
                    const int t = T - 1 - i;
                    const uint32_t k = k0 + i, s = k % kBwdStages, u = k / kBwdStages;
                    mbar_wait(&S.act_free[s], (u & 1) ^ 1);
                    mbar_arrive_expect_tx(&S.act_full[s], kStageBytes);
                    bulk_load(S.act[s], act_in + (((int64_t)t * ntiles + tile) * C::kInChunks) * (kAChunk / 2),
                              C::kInChunks * kAChunk, &S.act_full[s]);
                    const __nv_bfloat16* hsrc = t > 0 ? h + (((int64_t)(t - 1) * ntiles + tile) * 6) * (kAChunk / 2) : zeros;
                    bulk_load(S.act[s] + C::kHprevChunk * kAChunk, hsrc, 6 * kAChunk, &S.act_full[s]);
                }
        } else if (warp == kBwdEW) {
            // ================= MMA issuer ================================================================
            // per iteration i (step t = T-1-i):  R(t+1) -> commit r_full -> W(t+1) -> commit act_free, w_done -> G(t-1) -> commit
            // g_full.  Only R is on the step's critical path (d(gates) -> dh_rec -> d(gates)): the gate recompute of the NEXT
            // step is issued as soon as the epilogue has drained this step's gates out of TMEM (g_free) and runs, like the
            // dW accumulation, under the epilogue.
            {   // whole warp, warp-uniform control flow; one elected lane issues
                const bool leader = elect_one();
                // base descriptors, built once (see desc_adv)
                const uint64_t d_bg = umma_desc(smem_u32(S.bg), kBChunk, 128);                 // forward B, K-major
                const uint64_t d_br = umma_desc(smem_u32(S.br), C::kNR * 16, 128);             // [W_ih|W_hh]^T, K-major
                const uint64_t d_dgk = umma_desc(smem_u32(S.dg), kAChunk, 128);                // d(gates) as K-major A (R)
                const uint64_t d_dgm1 = umma_desc(smem_u32(S.dg), 128, kAChunk);               // d(gates)^T, MN-major A (W, rows 0..127)
                const uint64_t d_dgm2 = umma_desc(smem_u32(S.dg) + 8 * kAChunk, 128, kAChunk); // rows 64..191
                const uint64_t d_actk0 = umma_desc(smem_u32(S.act[0]), kAChunk, 128);          // stage s: desc_adv(.., s * stage bytes)
                const uint64_t d_actm0 = umma_desc(smem_u32(S.act[0]), 128, kAChunk);
                constexpr uint32_t kStageSmem = C::kStageChunks * kAChunk;
                uint32_t dgp = k0, gfp = k0;      // dg_ready / g_free phases consumed (T of each per tile)
                auto issue_g = [&](const uint32_t k) {
                    const uint32_t s = k % kBwdStages, u = k / kBwdStages;
                    mbar_wait(&S.act_full[s], u & 1);
                    tc_fence_after();
                    const uint64_t da = desc_adv(d_actk0, s * kStageSmem);
#pragma unroll
                    for (int ks = 0; ks < C::kStageChunks / 2; ++ks)
                        if (leader) umma_bf16(tm_g, desc_adv(da, 2 * ks * kAChunk), desc_adv(d_bg, 2 * ks * kBChunk), ks == 0 ? 0u : 1u);
                    if (leader) umma_commit(&S.g_full);
                };
                issue_g(k0);                      // G(T-1)
                for (int i = 0; i <= T; ++i) {
                    const uint32_t sp = (k0 + i - 1) % kBwdStages;
                    if (i >= 1) {
                        mbar_wait(&S.dg_ready, dgp & 1);
                        ++dgp;
                        tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 12; ++ks)
                            if (leader) umma_bf16_i(tm_r, desc_adv(d_dgk, 2 * ks * kAChunk), desc_adv(d_br, 2 * ks * C::kNR * 16), kIdescR,
                                        ks == 0 ? 0u : 1u);
                    }
                    if (leader) umma_commit(&S.r_full);        // R(t+1) done (i == 0: nothing pending)
                    if (i >= 1) {
                        const uint64_t db = desc_adv(d_actm0, sp * kStageSmem);
#pragma unroll
                        for (int ks = 0; ks < (HALF ? 4 : 8); ++ks) {          // half tiles: rows 64..127 are copies
                            const uint32_t acc = (first_w && ks == 0) ? 0u : 1u;
                            const uint64_t bdesc = desc_adv(db, ks * 256);
                            if (leader) umma_bf16_i(tm_w1, desc_adv(d_dgm1, ks * 256), bdesc, kIdescW, acc);
                            if (leader) umma_bf16_i(tm_w2, desc_adv(d_dgm2, ks * 256), bdesc, kIdescW, acc);
                        }
                        first_w = false;
                        if (leader) umma_commit(&S.act_free[sp]);
                        if (leader) umma_commit(&S.w_done);
                    }
                    if (i < T) {
                        mbar_wait(&S.g_free, gfp & 1);         // the epilogue has read G(t) out of TMEM
                        ++gfp;
                        tc_fence_after();
                        if (i + 1 < T) issue_g(k0 + i + 1);    // G(t-1)
                    }
                }
            }
        } else if (warp < kBwdEW) {
            // ================= epilogue: thread = window x 1/kBwdEG of the units (kBwdUB blocks of 8) =======
            const int q = warp & 3, hf = warp >> 2;
            const int row = q * 32 + lane;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            constexpr int kNB = HALF ? 1 : kBwdUB, kU = 8 * kNB;            // unit blocks / units of this thread
            const int rp = HALF ? (q >> 1) : 0;
            const int crow = HALF ? (row & 63) : row;                       // canonical row (first copy)
            const int blk0 = HALF ? hf * kBwdUB + rp : hf * kBwdUB;         // first 8-unit block
            const int u0 = 8 * blk0;                                        // first hidden unit
            const int64_t bwin = HALF ? (int64_t)tile * 64 + crow : b0 + row;
            const uint32_t xbar = HALF ? 1 + (q & 1) : 1 + q;               // the warps that share a window
            const uint32_t xcnt = HALF ? 64 * kBwdEG : 32 * kBwdEG;
            float dc[kU], ccur[kU];
#pragma unroll
            for (int j = 0; j < kU; ++j) dc[j] = 0.f;
            {   // c_{T-1} of this tile (afterwards c_t is carried over from the previous iteration's c_{t-1})
#pragma unroll
                for (int j = 0; j < kU; j += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(cstate + tcl32_off(T - 1, ntiles, tile, 2 * blk0 + j / 4, row));
                    ccur[j] = a.x; ccur[j + 1] = a.y; ccur[j + 2] = a.z; ccur[j + 3] = a.w;
                }
            }
            unsigned char* dgrow = S.dg + row * 16;
            float dzr[kU], dzz = 0.f, sm_m = 0.f, inv_l = 1.f;
            if (head) {
                const int64_t b = bwin;
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < kU; ++j) {
                    dzr[j] = (b < B) ? dz[b * kH + u0 + j] : 0.f;
                    part = fmaf(dzr[j], (b < B) ? zpool[b * kH + u0 + j] : 0.f, part);
                }
                if (b < B) { sm_m = stats[2 * b]; inv_l = 1.0f / stats[2 * b + 1]; }
                S.xch0[hf][row] = part;
                named_bar_sync(xbar, xcnt);
                dzz = 0.f;
#pragma unroll
                for (int e = 0; e < kBwdEG; ++e) {                          // dz . z  (all unit groups, fixed order)
                    dzz += S.xch0[e][crow];
                    if (HALF) dzz += S.xch0[e][crow + 64];
                }
            }
            for (int i = 0; i <= T; ++i) {
                const int t = T - 1 - i;                                  // step whose gates are in D_G (i < T)
                // ---- prefetch this step's c_{t-1} and dh_out BEFORE waiting for the tensor pipe -----------
                float cp[kU], dh[kU];
                if (i < T && head) {
                    // head backward fused here: the loads and the exchange overlap the tensor pipe's R/G of this step
                    // h_t: step t+1's stage holds it as h_{(t+1)-1} (already in shared memory: no global-load latency at the
                    // top of the step); only the first step of a tile reads it from HBM
                    uint4 hp[kNB];
                    if (i == 0) {
#pragma unroll
                        for (int bb = 0; bb < kNB; ++bb)
                            hp[bb] = *reinterpret_cast<const uint4*>(h + ((((int64_t)t * ntiles + tile) * 6 + blk0 + bb) * kRows + row) * 8);
                    } else {
                        const uint32_t kp = k0 + i - 1, sp = kp % kBwdStages;
                        mbar_wait(&S.act_full[sp], (kp / kBwdStages) & 1);          // completed long ago: acquires the TMA's writes
#pragma unroll
                        for (int bb = 0; bb < kNB; ++bb)
                            hp[bb] = *reinterpret_cast<const uint4*>(S.act[sp] + (C::kHprevChunk + blk0 + bb) * kAChunk + row * 16);
                    }
#pragma unroll
                    for (int j = 0; j < kU; j += 4) {
                        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t > 0) c4 = *reinterpret_cast<const float4*>(cstate + tcl32_off(t - 1, ntiles, tile, 2 * blk0 + j / 4, row));
                        cp[j] = c4.x; cp[j + 1] = c4.y; cp[j + 2] = c4.z; cp[j + 3] = c4.w;
                    }
                    float hv[kU];
#pragma unroll
                    for (int bb = 0; bb < kNB; ++bb) {
                        const uint32_t w4[4] = {hp[bb].x, hp[bb].y, hp[bb].z, hp[bb].w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) { hv[bb * 8 + 2 * u] = val_lo(w4[u]); hv[bb * 8 + 2 * u + 1] = val_hi(w4[u]); }
                    }
                    float sp = 0.f, gp = 0.f;
#pragma unroll
                    for (int j = 0; j < kU; ++j) { sp = fmaf(S.wa[u0 + j], hv[j], sp); gp = fmaf(dzr[j], hv[j], gp); }
                    if (i >= 1) mbar_arrive(&S.act_free[(k0 + i - 1) % kBwdStages]);          // this thread is done with the stage
                    S.xch[i & 1][hf][row] = make_float2(sp, gp);
                    named_bar_sync(xbar, xcnt);
                    float xs = 0.f, xg = 0.f;
#pragma unroll
                    for (int e = 0; e < kBwdEG; ++e) {                      // fixed order
                        const float2 xe = S.xch[i & 1][e][crow];
                        xs += xe.x; xg += xe.y;
                        if (HALF) { const float2 xf = S.xch[i & 1][e][crow + 64]; xs += xf.x; xg += xf.y; }
                    }
                    const float alpha = __expf(xs + S.ba - sm_m) * inv_l;
                    const float ds = alpha * (xg - dzz);
#pragma unroll
                    for (int j = 0; j < kU; ++j) {
                        dh[j] = fmaf(alpha, dzr[j], ds * S.wa[u0 + j]);
                        dwa[j] = fmaf(ds, hv[j], dwa[j]);
                    }
                    dba += ds;
                } else if (i < T) {
                    const int64_t grow = (int64_t)t * Bp + b0 + row;
                    const float* dhrow = dh_out + grow * kH + u0;                   // row-major TMP (separate head backward; full tiles only)
#pragma unroll
                    for (int j = 0; j < kU; j += 4) {
                        const float4 d = dh_chunked ? *reinterpret_cast<const float4*>(dh_out + tcl32_off(t, ntiles, tile, 2 * blk0 + j / 4, row))
                                                    : *reinterpret_cast<const float4*>(dhrow + j);
                        dh[j] = d.x; dh[j + 1] = d.y; dh[j + 2] = d.z; dh[j + 3] = d.w;
                        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t > 0) c4 = *reinterpret_cast<const float4*>(cstate + tcl32_off(t - 1, ntiles, tile, 2 * blk0 + j / 4, row));
                        cp[j] = c4.x; cp[j + 1] = c4.y; cp[j + 2] = c4.z; cp[j + 3] = c4.w;
                    }
                }
                // ---- phase A (does not need dh_rec): the gates of step t have been in TMEM since the previous epilogue; everything
                // of the first unit block that does not depend on dh -- the five activations and the products around them --
                // is evaluated while the tensor pipe runs R(t+1).  po = dh A,  dct = dh Bc + dc,  (pi, pf, pg) = dct (Ci, Cf, Cg)
                // (half tiles only: with 16 cells per thread the 48 coefficients do not fit the 128-register budget)
                constexpr bool kPre = HALF;
                float cA[8], cB[8], cI[8], cF[8], cG[8], cGf[8];
                if (kPre && i < T) {
                    mbar_wait(&S.g_full, gphase & 1);
                    ++gphase;
                    tc_fence_after();
                    uint32_t v[32];
                    tmem_ld32(tm_g + lane_base + blk0 * 32, v);
                    if (kNB == 1) {                                       // this thread's gates are in registers: G(t-1) may overwrite
                        tc_fence_before();
                        mbar_arrive(&S.g_free);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float gi = sigmoid_apx(__uint_as_float(v[u]));
                        const float gf = sigmoid_apx(__uint_as_float(v[8 + u]));
                        const float gg = tanh_apx(__uint_as_float(v[16 + u]));
                        const float go = sigmoid_apx(__uint_as_float(v[24 + u]));
                        const float tcv = tanh_apx(ccur[u]);
                        cA[u] = tcv * (go * (1.0f - go));
                        cB[u] = go * (1.0f - tcv * tcv);
                        cI[u] = gg * (gi * (1.0f - gi));
                        cF[u] = cp[u] * (gf * (1.0f - gf));
                        cG[u] = gi * (1.0f - gg * gg);
                        cGf[u] = gf;
                        ccur[u] = cp[u];                                  // c_{t-1} is next iteration's c_t
                    }
                }
                mbar_wait(&S.r_full, rphase & 1);
                ++rphase;
                if (!kPre && i < T) {
                    mbar_wait(&S.g_full, gphase & 1);
                    ++gphase;
                }
                tc_fence_after();
                if (KI == 48 && i >= 1) {
                    // din of step t+1 = D_R[:, 0:48] * mask * scale  -> dh_out of the layer below
                    const int64_t grow = (int64_t)(t + 1) * Bp + b0 + crow;                               // mask-tensor row
                    const int64_t gkey = HALF ? (int64_t)(t + 1) * drop_stride + bwin : grow;             // counter-based generator
#pragma unroll
                    for (int bb = 0; bb < kNB; ++bb) {
                        const int blk = blk0 + bb;
                        uint32_t r[8];
                        tmem_ld8(tm_r + lane_base + blk * 8, r);
                        float o[8];
                        if (mask || thresh16 < 65536u) {
                            const uint32_t keep = mask ? mask_keep8(*reinterpret_cast<const uint2*>(mask + grow * kH + blk * 8))
                                                       : dropout_keep8(seed, gkey, blk, thresh16);
#pragma unroll
                            for (int u = 0; u < 8; ++u) o[u] = ((keep >> u) & 1u) ? __uint_as_float(r[u]) * drop_scale : 0.f;
                        } else {
#pragma unroll
                            for (int u = 0; u < 8; ++u) o[u] = __uint_as_float(r[u]);
                        }
                        // (deferring these stores until after the fence + arrive below costs registers: 1.31 -> 1.80 ms, measured)
                        st_global_v4f(din + tcl32_off(t + 1, ntiles, tile, 2 * blk, row), o[0], o[1], o[2], o[3]);
                        st_global_v4f(din + tcl32_off(t + 1, ntiles, tile, 2 * blk + 1, row), o[4], o[5], o[6], o[7]);
                    }
                }
                if (i == T) {                                             // tail: all W of this tile must be done
                    if (head) mbar_arrive(&S.act_free[(k0 + T - 1) % kBwdStages]);  // keeps the stage's arrival count (no h to read)
                    mbar_wait(&S.w_done, wphase & 1);
                    ++wphase;
                    break;
                }
#pragma unroll
                for (int bb = 0; bb < kNB; ++bb) {
                    const int blk = blk0 + bb;
                    float pi[8], pf[8], pg[8], po[8];
                    if (kPre && bb == 0) {
                        // ---- phase B of the first block: only the products with dh_t = dh_out_t + dh_rec are left ----
                        if (i >= 1) {
                            uint32_t r[8];
                            tmem_ld8(tm_r + lane_base + C::kRecCol + blk * 8, r);
#pragma unroll
                            for (int u = 0; u < 8; ++u) dh[u] += __uint_as_float(r[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float dct = fmaf(dh[u], cB[u], dc[u]);
                            po[u] = dh[u] * cA[u];
                            dc[u] = dct * cGf[u];
                            pi[u] = dct * cI[u];
                            pf[u] = dct * cF[u];
                            pg[u] = dct * cG[u];
                        }
                        if (i >= 1) {                                     // W of the previous step still reads d(gates)
                            mbar_wait(&S.w_done, wphase & 1);
                            ++wphase;
                        }
                    } else {
                        uint32_t v[32];
                        tmem_ld32(tm_g + lane_base + blk * 32, v);
                        if (i >= 1) {
                            uint32_t r[8];
                            tmem_ld8(tm_r + lane_base + C::kRecCol + blk * 8, r);
#pragma unroll
                            for (int u = 0; u < 8; ++u) dh[bb * 8 + u] += __uint_as_float(r[u]);
                        }
                        if (bb == kNB - 1) {                              // this thread's gates are in registers: G(t-1) may overwrite
                            tc_fence_before();
                            mbar_arrive(&S.g_free);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = bb * 8 + u;
                            const float gi = sigmoid_apx(__uint_as_float(v[u]));
                            const float gf = sigmoid_apx(__uint_as_float(v[8 + u]));
                            const float gg = tanh_apx(__uint_as_float(v[16 + u]));
                            const float go = sigmoid_apx(__uint_as_float(v[24 + u]));
                            const float tcv = tanh_apx(ccur[j]);
                            const float d_o = dh[j] * tcv;
                            const float dct = fmaf(dh[j] * go, 1.0f - tcv * tcv, dc[j]);
                            dc[j] = dct * gf;
                            pi[u] = dct * gg * gi * (1.0f - gi);
                            pf[u] = dct * cp[j] * gf * (1.0f - gf);
                            pg[u] = dct * gi * (1.0f - gg * gg);
                            po[u] = d_o * go * (1.0f - go);
                            ccur[j] = cp[j];                              // c_{t-1} is next iteration's c_t
                        }
                        if (bb == 0 && i >= 1) {                          // W of the previous step still reads d(gates)
                            mbar_wait(&S.w_done, wphase & 1);
                            ++wphase;
                        }
                    }
                    unsigned char* d4 = dgrow + (blk * 4) * kAChunk;
                    st_shared_v4(d4, pack_val(pi[0], pi[1]), pack_val(pi[2], pi[3]), pack_val(pi[4], pi[5]), pack_val(pi[6], pi[7]));
                    st_shared_v4(d4 + kAChunk, pack_val(pf[0], pf[1]), pack_val(pf[2], pf[3]), pack_val(pf[4], pf[5]), pack_val(pf[6], pf[7]));
                    st_shared_v4(d4 + 2 * kAChunk, pack_val(pg[0], pg[1]), pack_val(pg[2], pg[3]), pack_val(pg[4], pg[5]), pack_val(pg[6], pg[7]));
                    st_shared_v4(d4 + 3 * kAChunk, pack_val(po[0], po[1]), pack_val(po[2], po[3]), pack_val(po[4], po[5]), pack_val(po[6], po[7]));
                    if (HALF) {                                           // the other row copy: D_R must be complete on every row
                        unsigned char* e4 = S.dg + (row ^ 64) * 16 + (blk * 4) * kAChunk;
                        st_shared_v4(e4, pack_val(pi[0], pi[1]), pack_val(pi[2], pi[3]), pack_val(pi[4], pi[5]), pack_val(pi[6], pi[7]));
                        st_shared_v4(e4 + kAChunk, pack_val(pf[0], pf[1]), pack_val(pf[2], pf[3]), pack_val(pf[4], pf[5]), pack_val(pf[6], pf[7]));
                        st_shared_v4(e4 + 2 * kAChunk, pack_val(pg[0], pg[1]), pack_val(pg[2], pg[3]), pack_val(pg[4], pg[5]), pack_val(pg[6], pg[7]));
                        st_shared_v4(e4 + 3 * kAChunk, pack_val(po[0], po[1]), pack_val(po[2], po[3]), pack_val(po[4], po[5]), pack_val(po[6], po[7]));
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&S.dg_ready);
            }
        }
        k0 += (uint32_t)T;       // every thread tracks the running step count; gphase / first_w are role-private
        __syncthreads();
    }

    // ---- per-CTA partial of d attn_w / d attn_b (fused head backward) ----------------------------------------
    if (head) {
        if (warp < kBwdEW) {
#pragma unroll
            for (int j = 0; j < kBwdU; ++j) {
                const float v = warp_sum(j < (HALF ? 8 : kBwdU) ? dwa[j] : 0.f);
                if (lane == 0) S.red[warp][j] = v;
            }
            const float vb = warp_sum(dba);
            if (lane == 0) S.red[warp][kBwdU] = vb;
        }
        __syncthreads();
        if (!HALF) {
            if (tid < kH) {
                const int hfj = tid / kBwdU, j = tid % kBwdU;
                attn_partial[(size_t)blockIdx.x * (kH + 1) + tid] =
                    S.red[4 * hfj][j] + S.red[4 * hfj + 1][j] + S.red[4 * hfj + 2][j] + S.red[4 * hfj + 3][j];
            } else if (tid == kH) {
                attn_partial[(size_t)blockIdx.x * (kH + 1) + kH] = S.red[0][kBwdU] + S.red[1][kBwdU] + S.red[2][kBwdU] + S.red[3][kBwdU];
            }
        } else {
            // unit u = 8 blk + j belongs to the warps (q, hf) with hf = blk / 2 and q / 2 = blk % 2 (both quarters of that copy);
            // ds is identical in every warp of a window: count it once (hf = 0, copy 0)
            if (tid < kH) {
                const int blk = tid / 8, j = tid % 8, w0 = 4 * (blk / kBwdUB) + 2 * (blk % kBwdUB);
                attn_partial[(size_t)blockIdx.x * (kH + 1) + tid] = S.red[w0][j] + S.red[w0 + 1][j];
            } else if (tid == kH) {
                attn_partial[(size_t)blockIdx.x * (kH + 1) + kH] = S.red[0][kBwdU] + S.red[1][kBwdU];
            }
        }
    }
    // ---- per-CTA weight-gradient partial: D_W1 rows = gate columns 0..127, D_W2 rows 64..127 = 128..191 ----
    tc_fence_after();
    if (warp < 4) {
        const int row = warp * 32 + lane;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float* out1 = dw_partial + ((size_t)blockIdx.x * kN + row) * C::kNW;
#pragma unroll 1
        for (int cc = 0; cc < C::kNW; cc += 8) {
            uint32_t r[8];
            tmem_ld8(tm_w1 + lane_base + cc, r);
            st_global_v4f(out1 + cc, __uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
            st_global_v4f(out1 + cc + 4, __uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
        }
        if (warp >= 2) {
            float* out2 = dw_partial + ((size_t)blockIdx.x * kN + 64 + row) * C::kNW;
#pragma unroll 1
            for (int cc = 0; cc < C::kNW; cc += 8) {
                uint32_t r[8];
                tmem_ld8(tm_w2 + lane_base + cc, r);
                st_global_v4f(out2 + cc, __uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                st_global_v4f(out2 + cc + 4, __uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kBwdEW + 1) { tc_fence_after(); tmem_free_all(tmem); }
}

// [W_ih | W_hh]^T as the B operand of R: out[chunk = n/8][o][n%8], n = permuted gate column (K index),
// o = output column (din columns first when KI == 48, then dh_rec).
__global__ void pack_bwd_r_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, int KI,
                                  __nv_bfloat16* __restrict__ out) {
    const int NR = KI == 8 ? 48 : 96;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 24 * NR * 8; idx += gridDim.x * blockDim.x) {
        const int n = (idx / (NR * 8)) * 8 + (idx % 8);
        const int o = (idx / 8) % NR;
        const int j = (n / 32) * 8 + (n % 8), q = (n % 32) / 8, col = q * kH + j;
        float v;
        if (KI == 8) v = w_hh[col * kH + o];
        else v = o < 48 ? w_ih[col * kH + o] : w_hh[col * kH + (o - 48)];
        reinterpret_cast<uint16_t*>(out)[idx] = val16(v);
    }
}

// Sum the per-CTA partials in a fixed order and scatter to torch layouts.
__global__ void reduce_dw_kernel(const float* __restrict__ partial, int nparts, int KI, float* __restrict__ dw_ih,
                                 float* __restrict__ dw_hh, float* __restrict__ db) {
    const int NW = KI == 8 ? 64 : 112;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kN * NW) return;
    const int n = idx / NW, f = idx % NW;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * kN * NW + idx];
    const int j = (n / 32) * 8 + (n % 8), q = (n % 32) / 8, col = q * kH + j;
    if (KI == 8) {
        if (f < 8) dw_ih[col * 8 + f] = s * kF16InScaleInv;      // the stored input is x / 16
        else if (f == 8) db[col] = s;
        else if (f >= 16) dw_hh[col * kH + (f - 16)] = s;
    } else {
        if (f < 48) dw_ih[col * kH + f] = s;
        else if (f < 96) dw_hh[col * kH + (f - 48)] = s;
        else if (f == 96) db[col] = s;
    }
}

// Materialise the counter-based dropout mask (tests: the RNG mode must equal the mask-tensor mode bit for bit).
__global__ void dropout_mask_kernel(uint64_t seed, uint32_t thresh16, int64_t rows, unsigned char* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // one (row, block) per thread
    if (i >= rows * 6) return;
    const uint32_t keep = dropout_keep8(seed, i / 6, (int)(i % 6), thresh16);
    uint2 v;
    v.x = ((keep >> 0) & 1u) | (((keep >> 1) & 1u) << 8) | (((keep >> 2) & 1u) << 16) | (((keep >> 3) & 1u) << 24);
    v.y = ((keep >> 4) & 1u) | (((keep >> 5) & 1u) << 8) | (((keep >> 6) & 1u) << 16) | (((keep >> 7) & 1u) << 24);
    reinterpret_cast<uint2*>(out)[i] = v;
}

static int tc_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <int KI>
static int launch_bwd(const void* act_in, const void* h, const float* c, const float* dh_out, const void* packed_g,
                      const void* packed_r, const void* zeros, const unsigned char* mask, uint64_t seed, uint32_t thresh16,
                      float scale, float* din, float* partial, const float* dz, const float* stats, const float* zpool,
                      const float* attn_w, const float* attn_b, int64_t B, float* attn_partial, int64_t T, int64_t Bp,
                      cudaStream_t st, int* grid_out, int64_t half_stride) {
    const size_t smem = (half_stride > 0 ? sizeof(BwdSmem<KI, bwd_stages(KI, true)>) : sizeof(BwdSmem<KI, bwd_stages(KI, false)>)) + 1024;
    auto kern = half_stride > 0 ? lstm_bwd_bf16_kernel<KI, true> : lstm_bwd_bf16_kernel<KI, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "lstm_bwd_bf16: shared memory opt-in (%zu B) failed: %s", smem, cudaGetErrorString(e));
    const int ntiles = (int)(Bp / kRows);
    const int grid = train_grid_cap(ntiles < tc_sms() ? ntiles : tc_sms());
    *grid_out = grid;
    kern<<<grid, kBwdThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(act_in), reinterpret_cast<const __nv_bfloat16*>(h),
                                          c, dh_out, reinterpret_cast<const unsigned char*>(packed_g),
                                          reinterpret_cast<const unsigned char*>(packed_r),
                                          reinterpret_cast<const __nv_bfloat16*>(zeros), mask, seed, thresh16, scale, din, partial,
                                          KI == 8 ? 1 : 0, dz, stats, zpool, attn_w, attn_b, B, attn_partial, (int)T, Bp, ntiles,
                                          half_stride > 0 ? half_stride : Bp);
    count_launch();
    return check_launch("na_lstm_bwd_bf16");
}

}  // namespace tc
}  // namespace na

// ---- C ABI ---------------------------------------------------------------------------------------
namespace na { namespace tc {
bool train_fwd_v2_enabled();
int launch_train_fwd_v2(const void*, const unsigned char*, const unsigned char*, uint64_t, uint32_t, float, void*, void*, float*, void*,
                        float*, const float*, const float*, float*, float*, int64_t, int, int64_t, int, cudaStream_t, int64_t);
} }

// scratch floats of na_lstm_bwd_bf16 after its 36,864-byte operand image: per-CTA attention partials [grid][52] then
// per-CTA weight-gradient partials [grid][192][112]; grid <= the SM count of the current device
extern "C" int64_t na_train_bf16_partial_floats(void) {
    const int64_t sms = na::tc::tc_sms();
    return sms * na::tc::kN * 112 + sms * 52;
}

extern "C" int na_lstm2_fwd_train_bf16(const void* x_bf16_tmp, const void* packed, const unsigned char* mask,
                                       uint64_t seed, int64_t thresh16, float drop_scale, void* h0, void* h0d, float* c0, void* h1, float* h1f,
                                       float* c1, const float* attn_w, const float* attn_b, float* zpool, float* stats,
                                       int64_t B, int64_t T, int64_t Bp, int64_t half_stride, na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(T >= 1 && T < (1 << 20) && Bp >= tc::kRows && Bp % tc::kRows == 0, NA_EINVAL,
               "na_lstm2_fwd_train_bf16: bad shape T=%lld Bp=%lld (Bp must be a multiple of 128)", (long long)T, (long long)Bp);
    NA_REQUIRE_PTR(x_bf16_tmp); NA_REQUIRE_PTR(packed); NA_REQUIRE_PTR(h0); NA_REQUIRE_PTR(c0);
    NA_REQUIRE_PTR(h1); NA_REQUIRE_PTR(c1);
    NA_OPTIONAL_PTR(mask); NA_OPTIONAL_PTR(h0d); NA_OPTIONAL_PTR(h1f); NA_OPTIONAL_PTR(zpool);
    NA_REQUIRE(half_stride >= 0, NA_EINVAL, "na_lstm2_fwd_train_bf16: half_stride < 0");
    NA_REQUIRE(half_stride == 0 || (tc::train_fwd_v2_enabled() && zpool != nullptr && h1f == nullptr && B <= Bp / 2), NA_EUNSUPPORTED,
               "na_lstm2_fwd_train_bf16: half tiles need the generation-2 kernel with fused pooling and B <= Bp / 2");
    NA_REQUIRE(zpool == nullptr || (attn_w && attn_b && stats && B >= 1 && B <= Bp), NA_EINVAL,
               "na_lstm2_fwd_train_bf16: fused pooling needs attn_w, attn_b, stats and 1 <= B <= Bp");
    NA_REQUIRE(zpool != nullptr || h1f != nullptr, NA_EINVAL, "na_lstm2_fwd_train_bf16: give h1f (separate head) or zpool (fused pooling)");
    NA_REQUIRE(thresh16 >= 0 && thresh16 <= 65536, NA_EINVAL, "na_lstm2_fwd_train_bf16: thresh16 outside [0,65536]");
    NA_REQUIRE((mask != nullptr || thresh16 < 65536) == (h0d != nullptr), NA_EINVAL,
               "na_lstm2_fwd_train_bf16: h0d must be given exactly when dropout is on (mask tensor or thresh16 < 65536)");
    if (tc::train_fwd_v2_enabled() && zpool != nullptr && h1f == nullptr)      // generation 2 (na_train_fwd2.cuh): fused pooling only
        return tc::launch_train_fwd_v2(x_bf16_tmp, reinterpret_cast<const unsigned char*>(packed) + na_decoder_packed_bf16_bytes() / 2, mask,
                                       seed, (uint32_t)thresh16, drop_scale, h0, h0d, c0, h1, c1, attn_w, attn_b, zpool, stats, B, (int)T,
                                       Bp, tc::tc_sms(), as_stream(stream), half_stride);
    const size_t smem = sizeof(tc::FwdSmem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc::lstm2_fwd_train_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "na_lstm2_fwd_train_bf16: shared memory opt-in failed (%s)", cudaGetErrorString(e));
    const int ntiles = (int)(Bp / tc::kRows);
    const int grid = train_grid_cap(ntiles < tc::tc_sms() ? ntiles : tc::tc_sms());
    tc::lstm2_fwd_train_bf16_kernel<<<grid, tc::kTrainThreads, smem, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_bf16_tmp), reinterpret_cast<const unsigned char*>(packed), mask, seed,
        (uint32_t)thresh16, drop_scale,
        reinterpret_cast<__nv_bfloat16*>(h0), reinterpret_cast<__nv_bfloat16*>(h0d), c0, reinterpret_cast<__nv_bfloat16*>(h1),
        h1f, c1, attn_w, attn_b, zpool, stats, B, (int)T, Bp, ntiles);
    count_launch();
    return check_launch("na_lstm2_fwd_train_bf16");
}

extern "C" int na_dropout_mask_u8(uint64_t seed, int64_t thresh16, int64_t T, int64_t Bp, unsigned char* out, na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(T >= 1 && Bp >= 1 && thresh16 >= 0 && thresh16 <= 65536, NA_EINVAL, "na_dropout_mask_u8: bad arguments");
    NA_REQUIRE_PTR(out);
    const int64_t n = T * Bp * 6;
    tc::dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(seed, (uint32_t)thresh16, T * Bp, out);
    count_launch();
    return check_launch("na_dropout_mask_u8");
}

extern "C" int na_lstm_bwd_bf16(int64_t layer, const void* act_in, const void* h, const float* cstate, const float* dh_out,
                                const void* packed_fwd, const float* w_ih, const float* w_hh, const void* zeros,
                                const unsigned char* in_mask, uint64_t seed, int64_t thresh16, float drop_scale,
                                float* din, float* dw_ih, float* dw_hh,
                                float* db, void* scratch, const float* dz, const float* stats, const float* zpool,
                                const float* attn_w, const float* attn_b, int64_t B, float* d_attn,
                                int64_t T, int64_t Bp, int64_t half_stride, na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(layer == 0 || layer == 1, NA_EINVAL, "na_lstm_bwd_bf16: layer must be 0 or 1");
    NA_REQUIRE(half_stride >= 0 && (half_stride == 0 || layer == 0 || dz != nullptr), NA_EUNSUPPORTED,
               "na_lstm_bwd_bf16: half tiles need the fused head backward on layer 1");
    NA_REQUIRE(T >= 1 && T < (1 << 20) && Bp >= tc::kRows && Bp % tc::kRows == 0, NA_EINVAL,
               "na_lstm_bwd_bf16: bad shape T=%lld Bp=%lld", (long long)T, (long long)Bp);
    NA_REQUIRE_PTR(act_in); NA_REQUIRE_PTR(h); NA_REQUIRE_PTR(cstate); NA_REQUIRE_PTR(packed_fwd);
    NA_OPTIONAL_PTR(dh_out); NA_OPTIONAL_PTR(dz);
    NA_REQUIRE((dz != nullptr) != (dh_out != nullptr), NA_EINVAL, "na_lstm_bwd_bf16: give exactly one of dh_out and dz");
    NA_REQUIRE(dz == nullptr || (layer == 1 && stats && zpool && attn_w && attn_b && d_attn && B >= 1 && B <= Bp), NA_EINVAL,
               "na_lstm_bwd_bf16: the fused head backward is for layer 1 and needs stats, zpool, attn_w, attn_b, d_attn, B");
    NA_REQUIRE_PTR(w_ih); NA_REQUIRE_PTR(w_hh); NA_REQUIRE_PTR(zeros); NA_REQUIRE_PTR(dw_ih); NA_REQUIRE_PTR(dw_hh);
    NA_REQUIRE_PTR(db); NA_REQUIRE_PTR(scratch);
    NA_OPTIONAL_PTR(in_mask); NA_OPTIONAL_PTR(din);
    NA_REQUIRE(layer == 0 || din != nullptr, NA_EINVAL, "na_lstm_bwd_bf16: layer 1 needs din");
    NA_REQUIRE(thresh16 >= 0 && thresh16 <= 65536, NA_EINVAL, "na_lstm_bwd_bf16: thresh16 outside [0,65536]");
    cudaStream_t st = as_stream(stream);
    const int KI = layer == 0 ? 8 : 48;
    // scratch: [packed_r bf16: 24*96*8*2 B = 36,864 B][partials fp32]
    __nv_bfloat16* packed_r = reinterpret_cast<__nv_bfloat16*>(scratch);
    float* attn_partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 36864);        // [sms][49 (52)]
    float* partial = attn_partial + (size_t)tc::tc_sms() * 52;
    tc::pack_bwd_r_kernel<<<32, 256, 0, st>>>(w_ih, w_hh, KI, packed_r);
    count_launch();
    // the forward pack holds B0 (8 chunks) then B1 (14 chunks)
    const unsigned char* pg = reinterpret_cast<const unsigned char*>(packed_fwd) + (layer == 0 ? 0 : 8 * tc::kBChunk);
    int grid = 0, rc;
    if (layer == 0)
        rc = tc::launch_bwd<8>(act_in, h, cstate, dh_out, pg, packed_r, zeros, nullptr, 0, 65536u, 1.0f, nullptr, partial,
                               nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, T, Bp, st, &grid, half_stride);
    else
        rc = tc::launch_bwd<48>(act_in, h, cstate, dh_out, pg, packed_r, zeros, in_mask, seed, (uint32_t)thresh16, drop_scale, din,
                                partial, dz, stats, zpool, attn_w, attn_b, B, attn_partial, T, Bp, st, &grid, half_stride);
    if (rc) return rc;
    if (dz != nullptr)
        if ((rc = reduce_partials(attn_partial, d_attn, grid, tc::kH + 1, st))) return rc;
    const int NW = layer == 0 ? 64 : 112;
    tc::reduce_dw_kernel<<<(tc::kN * NW + 255) / 256, 256, 0, st>>>(partial, grid, KI, dw_ih, dw_hh, db);
    count_launch();
    return check_launch("na_lstm_bwd_bf16(reduce)");
}
