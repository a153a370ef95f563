// Training forward, generation 2: the software-pipelined structure of decoder_infer_v2_kernel (each epilogue warp owns
// both layers of (lane quarter, 16-unit group); attention score from the tensor core one step late; 0.5 scalings folded
// into the packed weights, H = 2h carried in shared memory) + the saves the backward needs.
// Included by na_decoder_tc2.cu (shares its helpers and the v2 weight section).
//   saves (same layouts as lstm2_fwd_train_bf16_kernel): h0, h0d = h0 * mask * scale, h1 (fp16 TCL: [T][NT][6][128][8]),
//   c0, c1 (fp32 TCL32: [T][NT][12][128][4]); fused attention pooling: zpool [B][48], stats [B][2] = (max, sum).
#pragma once

namespace na {
namespace tc {

struct SmemT2 {
    alignas(128) unsigned char b0[kV2K0Chunks * kBChunk];
    alignas(128) unsigned char b1[kV2K1Chunks * kB1Chunk];
    alignas(128) unsigned char x[kV2XStages][2 * kAChunk];         // [x chunk | ones chunk]
    alignas(128) unsigned char h0[2][6 * kAChunk];                 // H0_t = 2 h0_t          (layer-0 recurrence)
    alignas(128) unsigned char h0d[2][6 * kAChunk];                // dropped H0_t           (layer-1 input)
    alignas(128) unsigned char h1[6 * kAChunk];
    alignas(128) unsigned char onez[2 * kAChunk];
    alignas(8) uint64_t x_full[kV2XStages], x_empty[kV2XStages];
    uint64_t d0_full, d1_full, h0_ready[2], h1_ready;
    uint32_t tmem_base;
};

// cell update of 4 units returning fp32 H = 2h (the caller derives the fp16 operand, the saved h and the dropped copy)
__device__ __forceinline__ void cell_granule_f(const uint32_t* v, float* c, float* H) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float ti = tanh_apx(__uint_as_float(v[u]));
        const float tf = tanh_apx(__uint_as_float(v[4 + u]));
        const float tg = tanh_apx(__uint_as_float(v[8 + u]));
        const float to = tanh_apx(__uint_as_float(v[12 + u]));
        const float w = fmaf(tf, c[u], c[u]);
        const float uu = fmaf(ti, tg, tg);
        c[u] = 0.5f * (w + uu);
        const float tcell = tanh_apx(c[u]);
        H[u] = fmaf(to, tcell, tcell);
    }
}

__device__ __forceinline__ uint32_t half2_halve(uint32_t p) {        // exact: fp16 x 0.5 (H = 2h -> h)
    __half2 v = *reinterpret_cast<__half2*>(&p);
    v = __hmul2(v, __floats2half2_rn(0.5f, 0.5f));
    return *reinterpret_cast<uint32_t*>(&v);
}

// HALF = half tiles (strong scaling at small batches): a tile holds 64 distinct windows and rows 64..127 of every operand
// mirror rows 0..63 (the host replicates the input rows, this kernel stores h to both copies, in shared memory and in
// HBM).  The MMAs are unchanged; the two copies split the hidden units -- copy rp = q / 2 takes K chunk 2 g + rp (8 units
// per thread instead of 16) -- so a tile costs roughly half a step and a batch of B windows occupies B / 64 SMs.
template <bool HALF>
__global__ void __launch_bounds__(kV2Threads, 1)
lstm2_fwd_train_v2_kernel(const __nv_bfloat16* __restrict__ x,        // TMP [T][Bp][8] fp16 bits
                          const unsigned char* __restrict__ packed,   // v2 section (pack_decoder_v2_kernel)
                          const unsigned char* __restrict__ mask, uint64_t seed, uint32_t thresh16, float drop_scale,
                          __nv_bfloat16* __restrict__ h0_out, __nv_bfloat16* __restrict__ h0d_out, float* __restrict__ c0_out,
                          __nv_bfloat16* __restrict__ h1_out, float* __restrict__ c1_out,
                          const float* __restrict__ attn_w, const float* __restrict__ attn_b,
                          float* __restrict__ zpool_out, float* __restrict__ stats_out, int64_t B,
                          int T, int64_t Bp, int ntiles, int64_t drop_stride) {
    constexpr int kMmaWarp = 12, kTmaWarp = 13;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SmemT2& S = *reinterpret_cast<SmemT2*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    {
        const uint4* src = reinterpret_cast<const uint4*>(packed);
        uint4* d0 = reinterpret_cast<uint4*>(S.b0);
        constexpr int n0_16 = kV2K0Chunks * kBChunk / 16;
        for (int i = tid; i < n0_16; i += kV2Threads) d0[i] = src[i];
        for (int i = tid; i < kV2K1Chunks * kN; i += kV2Threads) {
            const int ch = i / kN, r = i % kN;
            reinterpret_cast<uint4*>(S.b1 + ch * kB1Chunk)[r] = src[n0_16 + i];
        }
        for (int i = tid; i < kV2K1Chunks * 16; i += kV2Threads) {          // attention score columns (rows 192..207)
            const int ch = i / 16, r = i % 16;
            uint16_t w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = ch * 8 + e;
                float v = 0.f;
                if (r < 2) {
                    float full = 0.f;
                    if (k >= 48 && k < 96) full = 0.5f * attn_w[k - 48];
                    else if (k == 96) full = attn_b[0];
                    const float hi = val16_to_float(val16(full));
                    v = (r == 0) ? hi : full - hi;
                }
                w[e] = val16(v);
            }
            uint4 pk;
            pk.x = w[0] | ((uint32_t)w[1] << 16); pk.y = w[2] | ((uint32_t)w[3] << 16);
            pk.z = w[4] | ((uint32_t)w[5] << 16); pk.w = w[6] | ((uint32_t)w[7] << 16);
            reinterpret_cast<uint4*>(S.b1 + ch * kB1Chunk)[kN + r] = pk;
        }
        const uint4 ones = make_uint4(kValOnes2, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < kRows; i += kV2Threads) {
#pragma unroll
            for (int s = 0; s < kV2XStages; ++s) reinterpret_cast<uint4*>(S.x[s] + kAChunk)[i] = ones;
            reinterpret_cast<uint4*>(S.onez)[i] = ones;
            reinterpret_cast<uint4*>(S.onez + kAChunk)[i] = zero;
        }
        if (tid == 0) {
            for (int s = 0; s < kV2XStages; ++s) { mbar_init(&S.x_full[s], 1); mbar_init(&S.x_empty[s], 1); }
            mbar_init(&S.d0_full, 1); mbar_init(&S.d1_full, 1);
            mbar_init(&S.h0_ready[0], 384); mbar_init(&S.h0_ready[1], 384);
            mbar_init(&S.h1_ready, 384);
            fence_mbar_init();
        }
        if (warp == kTmaWarp) tmem_alloc_all(&S.tmem_base);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = __shfl_sync(0xffffffffu, S.tmem_base, 0);
    const uint32_t tmem_d0 = tmem, tmem_d1 = tmem + kN;
    const bool drop = mask != nullptr || thresh16 < 65536u;

    int n0 = 0;
    uint32_t k1 = 0;                                       // running d1_full phase index (T + 1 per tile)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n0 += T) {
        const int64_t b0 = (int64_t)tile * kRows;
        if (warp == kTmaWarp) {
            if (lane == 0)
                for (int t = 0; t < T; ++t) {
                    const int n = n0 + t, s = n % kV2XStages, u = n / kV2XStages;
                    mbar_wait(&S.x_empty[s], (u & 1) ^ 1);
                    mbar_arrive_expect_tx(&S.x_full[s], kAChunk);
                    bulk_load(S.x[s], x + ((int64_t)t * Bp + b0) * 8, kAChunk, &S.x_full[s]);
                }
        } else if (warp == kMmaWarp) {
            const bool leader = elect_one();
            const uint64_t d_b0 = umma_desc(smem_u32(S.b0), kBChunk, 128), d_b1 = umma_desc(smem_u32(S.b1), kB1Chunk, 128);
            const uint64_t d_x0 = umma_desc(smem_u32(S.x[0]), kAChunk, 128);
            const uint64_t d_h0[2] = {umma_desc(smem_u32(S.h0[0]), kAChunk, 128), umma_desc(smem_u32(S.h0[1]), kAChunk, 128)};
            const uint64_t d_h0d[2] = {umma_desc(smem_u32(S.h0d[0]), kAChunk, 128), umma_desc(smem_u32(S.h0d[1]), kAChunk, 128)};
            const uint64_t d_h1 = umma_desc(smem_u32(S.h1), kAChunk, 128), d_onez = umma_desc(smem_u32(S.onez), kAChunk, 128);
            for (int t = 0; t <= T; ++t) {
                const int n = n0 + t;
                if (t >= 1) {                                  // H0_{t-1} (and its dropped copy) written, D0 drained
                    mbar_wait(&S.h0_ready[(n - 1) & 1], ((n - 1) >> 1) & 1);
                    tc_fence_after();
                }
                if (t < T) {                                   // layer 0, step t
                    const int s = n % kV2XStages, u = n / kV2XStages;
                    mbar_wait(&S.x_full[s], u & 1);
                    tc_fence_after();
                    if (leader) umma_bf16(tmem_d0, desc_adv(d_x0, s * 2 * kAChunk), d_b0, 0u);
                    if (t >= 1) {
                        const uint64_t hprev = d_h0[(n - 1) & 1];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            if (leader) umma_bf16(tmem_d0, desc_adv(hprev, 2 * i * kAChunk), desc_adv(d_b0, (2 + 2 * i) * kBChunk), 1u);
                    }
                    if (leader) umma_commit(&S.d0_full);
                }
                if (t >= 1) {                                  // layer 1, step m = t - 1
                    const int m = n - 1;
                    if (t >= 2) {
                        mbar_wait(&S.h1_ready, (m - 1) & 1);
                        tc_fence_after();
                    }
                    const uint64_t hin = drop ? d_h0d[m & 1] : d_h0[m & 1];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (leader) umma_bf16_i(tmem_d1, desc_adv(hin, 2 * i * kAChunk), desc_adv(d_b1, 2 * i * kB1Chunk), kIdescL1, i == 0 ? 0u : 1u);
                    if (t >= 2) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            if (leader) umma_bf16_i(tmem_d1, desc_adv(d_h1, 2 * i * kAChunk), desc_adv(d_b1, (6 + 2 * i) * kB1Chunk), kIdescL1, 1u);
                    }
                    if (leader) umma_bf16_i(tmem_d1, d_onez, desc_adv(d_b1, 12 * kB1Chunk), kIdescL1, 1u);
                    if (leader) umma_commit(&S.d1_full);
                }
            }
            {   // flush: score of the last step into the 16 score columns
                const int m = n0 + T - 1;
                mbar_wait(&S.h1_ready, m & 1);
                tc_fence_after();
                const uint64_t d_b1s = desc_adv(d_b1, kN * 16);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (leader) umma_bf16_i(tmem_d1 + kN, desc_adv(d_h1, 2 * i * kAChunk), desc_adv(d_b1s, (6 + 2 * i) * kB1Chunk), kIdescFlush, i == 0 ? 0u : 1u);
                if (leader) umma_bf16_i(tmem_d1 + kN, d_onez, desc_adv(d_b1s, 12 * kB1Chunk), kIdescFlush, 1u);
                if (leader) umma_commit(&S.d1_full);
            }
        } else {
            // ================= epilogue: both layers of (quarter q, unit group g: K chunks 2g, 2g+1) ==================
            const int q = warp & 3, g = warp >> 2;
            const int row = q * 32 + lane;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            constexpr int kNB = HALF ? 1 : 2;              // 8-unit blocks per thread
            const int rp = HALF ? (q >> 1) : 0;            // row copy of this warp
            const int crow = HALF ? (row & 63) : row;      // canonical row (first copy)
            const int blk0 = HALF ? 2 * g + rp : 2 * g;    // first K chunk of this thread
            const int64_t bwin = HALF ? (int64_t)tile * 64 + crow : b0 + row;       // window index
            float c0[8 * kNB], c1[8 * kNB], z[8 * kNB];
#pragma unroll
            for (int j = 0; j < 8 * kNB; ++j) { c0[j] = 0.f; c1[j] = 0.f; z[j] = 0.f; }
            uint32_t hprev[4 * kNB];
#pragma unroll
            for (int j = 0; j < 4 * kNB; ++j) hprev[j] = 0u;
            float mx = -INFINITY, l = 0.f;
            auto pool = [&](float score) {                 // online softmax over time (lstm_eeg_model.py:35-37)
                if (score > mx) {
                    const float sc = __expf(mx - score);
                    l *= sc;
#pragma unroll
                    for (int j = 0; j < 8 * kNB; ++j) z[j] *= sc;
                    mx = score;
                }
                const float e = __expf(score - mx);
                l += e;
#pragma unroll
                for (int u = 0; u < 4 * kNB; ++u) {
                    z[2 * u] = fmaf(e, val_lo(hprev[u]), z[2 * u]);
                    z[2 * u + 1] = fmaf(e, val_hi(hprev[u]), z[2 * u + 1]);
                }
            };
            for (int t = 0; t <= T; ++t) {
                const int n = n0 + t;
                if (t < T) {                               // ---- layer 0, step t
                    const int64_t grow = (int64_t)t * Bp + b0 + crow;                      // mask-tensor row
                    const int64_t gkey = HALF ? (int64_t)t * drop_stride + bwin : grow;   // counter-based generator: layout-independent
                    const int64_t tcl = ((int64_t)t * ntiles + tile) * 6 * (kAChunk / 2) + row * 8;
                    uint32_t keep[kNB];
#pragma unroll
                    for (int pr = 0; pr < kNB; ++pr) keep[pr] = 0xFFu;
                    if (drop) {
#pragma unroll
                        for (int pr = 0; pr < kNB; ++pr) {
                            const int blk = blk0 + pr;
                            keep[pr] = mask ? mask_keep8(*reinterpret_cast<const uint2*>(mask + grow * kH + blk * 8))
                                            : dropout_keep8(seed, gkey, blk, thresh16);
                        }
                    }
                    mbar_wait(&S.d0_full, n & 1);
                    if (q == 0 && g == 0 && lane == 0) mbar_arrive(&S.x_empty[n % kV2XStages]);
                    tc_fence_after();
                    // the saves go to HBM only AFTER the fence + arrive that release h0_t to the tensor pipe: the fence
                    // (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC) otherwise waits for the global stores, on the step's critical path
                    uint32_t ph[4 * kNB], pd[4 * kNB];
#pragma unroll
                    for (int pr = 0; pr < kNB; ++pr) {
                        const int blk = blk0 + pr;
                        uint32_t v[32];
                        float H[8];
                        tmem_ld32(tmem_d0 + lane_base + blk * 32, v);
                        cell_granule_f(v, c0 + pr * 8, H);
                        cell_granule_f(v + 16, c0 + pr * 8 + 4, H + 4);
                        ph[pr * 4] = pack_val(H[0], H[1]); ph[pr * 4 + 1] = pack_val(H[2], H[3]);
                        ph[pr * 4 + 2] = pack_val(H[4], H[5]); ph[pr * 4 + 3] = pack_val(H[6], H[7]);
                        st_shared_v4(S.h0[n & 1] + blk * kAChunk + row * 16, ph[pr * 4], ph[pr * 4 + 1], ph[pr * 4 + 2], ph[pr * 4 + 3]);
                        if (HALF)                          // the other row copy
                            st_shared_v4(S.h0[n & 1] + blk * kAChunk + (row ^ 64) * 16, ph[pr * 4], ph[pr * 4 + 1], ph[pr * 4 + 2], ph[pr * 4 + 3]);
                        if (drop) {
                            float Hd[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) Hd[u] = ((keep[pr] >> u) & 1u) ? H[u] * drop_scale : 0.f;
                            pd[pr * 4] = pack_val(Hd[0], Hd[1]); pd[pr * 4 + 1] = pack_val(Hd[2], Hd[3]);
                            pd[pr * 4 + 2] = pack_val(Hd[4], Hd[5]); pd[pr * 4 + 3] = pack_val(Hd[6], Hd[7]);
                            st_shared_v4(S.h0d[n & 1] + blk * kAChunk + row * 16, pd[pr * 4], pd[pr * 4 + 1], pd[pr * 4 + 2], pd[pr * 4 + 3]);
                            if (HALF)
                                st_shared_v4(S.h0d[n & 1] + blk * kAChunk + (row ^ 64) * 16, pd[pr * 4], pd[pr * 4 + 1], pd[pr * 4 + 2], pd[pr * 4 + 3]);
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();
                    mbar_arrive(&S.h0_ready[n & 1]);
#pragma unroll
                    for (int pr = 0; pr < kNB; ++pr) {
                        const int blk = blk0 + pr;
                        const uint4 hs = make_uint4(half2_halve(ph[pr * 4]), half2_halve(ph[pr * 4 + 1]), half2_halve(ph[pr * 4 + 2]), half2_halve(ph[pr * 4 + 3]));
                        *reinterpret_cast<uint4*>(h0_out + tcl + blk * (kAChunk / 2)) = hs;
                        if (HALF) *reinterpret_cast<uint4*>(h0_out + tcl + blk * (kAChunk / 2) + ((row ^ 64) - row) * 8) = hs;
                        st_global_v4f(c0_out + tcl32_off(t, ntiles, tile, 2 * blk, row), c0[pr * 8], c0[pr * 8 + 1], c0[pr * 8 + 2], c0[pr * 8 + 3]);
                        st_global_v4f(c0_out + tcl32_off(t, ntiles, tile, 2 * blk + 1, row), c0[pr * 8 + 4], c0[pr * 8 + 5], c0[pr * 8 + 6], c0[pr * 8 + 7]);
                        if (drop) {
                            const uint4 ds = make_uint4(half2_halve(pd[pr * 4]), half2_halve(pd[pr * 4 + 1]), half2_halve(pd[pr * 4 + 2]), half2_halve(pd[pr * 4 + 3]));
                            *reinterpret_cast<uint4*>(h0d_out + tcl + blk * (kAChunk / 2)) = ds;
                            if (HALF) *reinterpret_cast<uint4*>(h0d_out + tcl + blk * (kAChunk / 2) + ((row ^ 64) - row) * 8) = ds;
                        }
                    }
                }
                if (t >= 1) {                              // ---- layer 1, step t-1 (+ pooling of step t-2)
                    const int tt = t - 1;
                    const int64_t tcl = ((int64_t)tt * ntiles + tile) * 6 * (kAChunk / 2) + row * 8;
                    mbar_wait(&S.d1_full, k1 & 1); ++k1;
                    tc_fence_after();
                    uint32_t sc2[2];
                    tmem_ld2(tmem_d1 + lane_base + kN, sc2);
                    uint32_t hb[4 * kNB];
#pragma unroll
                    for (int pr = 0; pr < kNB; ++pr) {
                        const int blk = blk0 + pr;
                        uint32_t v[32];
                        float H[8];
                        tmem_ld32(tmem_d1 + lane_base + blk * 32, v);
                        cell_granule_f(v, c1 + pr * 8, H);
                        cell_granule_f(v + 16, c1 + pr * 8 + 4, H + 4);
                        hb[pr * 4] = pack_val(H[0], H[1]); hb[pr * 4 + 1] = pack_val(H[2], H[3]);
                        hb[pr * 4 + 2] = pack_val(H[4], H[5]); hb[pr * 4 + 3] = pack_val(H[6], H[7]);
                        st_shared_v4(S.h1 + blk * kAChunk + row * 16, hb[pr * 4], hb[pr * 4 + 1], hb[pr * 4 + 2], hb[pr * 4 + 3]);
                        if (HALF) st_shared_v4(S.h1 + blk * kAChunk + (row ^ 64) * 16, hb[pr * 4], hb[pr * 4 + 1], hb[pr * 4 + 2], hb[pr * 4 + 3]);
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();
                    mbar_arrive(&S.h1_ready);
#pragma unroll
                    for (int pr = 0; pr < kNB; ++pr) {             // saves: after the release (see layer 0)
                        const int blk = blk0 + pr;
                        const uint4 hs = make_uint4(half2_halve(hb[pr * 4]), half2_halve(hb[pr * 4 + 1]), half2_halve(hb[pr * 4 + 2]), half2_halve(hb[pr * 4 + 3]));
                        *reinterpret_cast<uint4*>(h1_out + tcl + blk * (kAChunk / 2)) = hs;
                        if (HALF) *reinterpret_cast<uint4*>(h1_out + tcl + blk * (kAChunk / 2) + ((row ^ 64) - row) * 8) = hs;
                        st_global_v4f(c1_out + tcl32_off(tt, ntiles, tile, 2 * blk, row), c1[pr * 8], c1[pr * 8 + 1], c1[pr * 8 + 2], c1[pr * 8 + 3]);
                        st_global_v4f(c1_out + tcl32_off(tt, ntiles, tile, 2 * blk + 1, row), c1[pr * 8 + 4], c1[pr * 8 + 5], c1[pr * 8 + 6], c1[pr * 8 + 7]);
                    }
                    if (t >= 2) pool(__uint_as_float(sc2[0]) + __uint_as_float(sc2[1]));
#pragma unroll
                    for (int u = 0; u < 4 * kNB; ++u) hprev[u] = hb[u];
                }
            }
            {                                              // flush: score of the last step
                mbar_wait(&S.d1_full, k1 & 1); ++k1;
                tc_fence_after();
                uint32_t sc2[2];
                tmem_ld2(tmem_d1 + lane_base + kN, sc2);
                tc_fence_before();
                pool(__uint_as_float(sc2[0]) + __uint_as_float(sc2[1]));
            }
            if (bwin < B) {                                // pooled vector (H = 2h accumulated) and the softmax stats
                const float inv_l = 0.5f / l;
                float* zo = zpool_out + bwin * kH + blk0 * 8;
#pragma unroll
                for (int j = 0; j < 8 * kNB; j += 4) st_global_v4f(zo + j, z[j] * inv_l, z[j + 1] * inv_l, z[j + 2] * inv_l, z[j + 3] * inv_l);
                if (g == 0 && rp == 0) { stats_out[2 * bwin] = mx; stats_out[2 * bwin + 1] = l; }
            }
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTmaWarp) { tc_fence_after(); tmem_free_all(tmem); }
}

int g_train_fwd_v2 = 1;      // na_set_tuning("tc_train_fwd_v2", 0 | 1)
void set_train_fwd_v2(int v) { g_train_fwd_v2 = v ? 1 : 0; }
bool train_fwd_v2_enabled() { return g_train_fwd_v2 != 0; }

int launch_train_fwd_v2(const void* x, const unsigned char* packed_v2, const unsigned char* mask, uint64_t seed, uint32_t thresh16,
                        float drop_scale, void* h0, void* h0d, float* c0, void* h1, float* c1, const float* attn_w, const float* attn_b,
                        float* zpool, float* stats, int64_t B, int T, int64_t Bp, int sms, cudaStream_t stream, int64_t half_stride) {
    const size_t smem = sizeof(SmemT2) + 1024;
    auto kern = half_stride > 0 ? lstm2_fwd_train_v2_kernel<true> : lstm2_fwd_train_v2_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "na_lstm2_fwd_train_bf16: shared memory opt-in failed (%s)", cudaGetErrorString(e));
    const int ntiles = (int)(Bp / kRows);
    const int grid = train_grid_cap(ntiles < sms ? ntiles : sms);
    kern<<<grid, kV2Threads, smem, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), packed_v2, mask, seed, thresh16, drop_scale, reinterpret_cast<__nv_bfloat16*>(h0),
        reinterpret_cast<__nv_bfloat16*>(h0d), c0, reinterpret_cast<__nv_bfloat16*>(h1), c1, attn_w, attn_b, zpool, stats, B, T, Bp, ntiles,
        half_stride > 0 ? half_stride : Bp);
    count_launch();
    return check_launch("na_lstm2_fwd_train_bf16 (v2)");
}

}  // namespace tc
}  // namespace na
