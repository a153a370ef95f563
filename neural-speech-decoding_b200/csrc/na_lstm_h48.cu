// K3 (specialised tier, exact fp32): one LSTM layer with H = 48 for all timesteps.
//
// Design (B200):
//   * one CTA per SM, NG independent 128-thread groups per CTA; each group owns a tile of
//     32 windows and walks the T steps on-chip.  The [W_ih | W_hh]^T matrix (K+48 rows x 192)
//     is staged ONCE per CTA in shared memory and shared by the groups; it is never re-read
//     from HBM/L2 inside the time loop.
//   * register tiling: thread = 4 windows x 3 hidden units x 4 gates = 48 fp32 accumulators;
//     per k it issues 1/4 + 3 LDS.128 for 48 FFMA.  Weight rows are laid out
//     [k][unit-group][gate][3] so the three LDS.128 of a thread are contiguous and the 16
//     unit-groups of a half-warp sweep all 32 banks exactly twice (the 2-wavefront minimum);
//     the two window-groups of a warp sit 48 words apart = 16 banks -> conflict-free.
//   * I/O is time-major: the tile's step-t input and its outputs are contiguous blocks,
//     moved by 1-D TMA bulk copies (cp.async.bulk, mbarrier complete_tx for loads, bulk
//     async-groups for stores) issued by one elected thread per group; a 3-stage input ring
//     hides the load latency behind the ~2.5 us step.
//   * activations: MUFU.EX2 + MUFU.RCP forms (abs. error ~2e-7, same order as the fp32
//     rounding of the pre-activation) -- no tanh.approx, no TF32.
//   * two named barriers per group and step; groups never wait on each other.
#include "na_common.cuh"
#include "na_sm100.cuh"

namespace na {

constexpr int kH = 48;
constexpr int kG = 4 * kH;          // 192 gate columns
constexpr int kTB = 32;             // windows per group tile
constexpr int kGroupThreads = 128;
constexpr int kStages = 3;          // input ring depth

template <int KIN>
struct H48Smem {
    static constexpr int kK = KIN + kH;
    static constexpr size_t weights = sizeof(float) * kK * kG;           // shared by all groups
    static constexpr size_t bias = sizeof(float) * kG;
    static constexpr size_t in_stage = sizeof(float) * kTB * KIN;
    static constexpr size_t h_buf = sizeof(float) * kTB * kH;
    static constexpr size_t c_buf = sizeof(float) * kTB * kH;
    static constexpr size_t g_buf = sizeof(float) * kTB * kG;
    static constexpr size_t bars = 64;                                    // kStages mbarriers, padded
    static constexpr size_t per_group(bool save) {
        return kStages * in_stage + h_buf + (save ? c_buf + g_buf : 0) + bars;
    }
    static constexpr size_t total(int ng, bool save) { return weights + bias + ng * per_group(save); }
};

template <int KIN, int NG, bool SAVE>
__global__ void __launch_bounds__(NG * kGroupThreads, 1)
lstm_fwd_h48_kernel(const float* __restrict__ in, const float* __restrict__ wt, const float* __restrict__ bias,
                    float* __restrict__ hout, float* __restrict__ cout, float* __restrict__ gates,
                    int T, int64_t Bp, int ntiles) {
    using L = H48Smem<KIN>;
    constexpr int K = L::kK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* w_s = reinterpret_cast<float*>(smem_raw);                       // [K][16][4][3]
    float* bias_s = w_s + K * kG;                                          // [16][4][3] (same order)
    unsigned char* gbase = smem_raw + L::weights + L::bias;

    const int tid = threadIdx.x;
    const int g = tid / kGroupThreads;           // group within the CTA
    const int gt = tid % kGroupThreads;          // thread within the group
    const int ug = gt & 15;                      // units 3*ug .. 3*ug+2
    const int wg = gt >> 4;                      // windows wg, wg+8, wg+16, wg+24

    unsigned char* mine = gbase + (size_t)g * L::per_group(SAVE);
    float* in_s = reinterpret_cast<float*>(mine);                          // [kStages][kTB][KIN]
    float* h_s = in_s + kStages * kTB * KIN;                               // [kTB][kH]
    float* c_s = h_s + kTB * kH;                                           // [kTB][kH]      (SAVE)
    float* g_s = c_s + kTB * kH;                                           // [kTB][kG]      (SAVE)
    uint64_t* full = reinterpret_cast<uint64_t*>(mine + L::per_group(SAVE) - L::bars);

    // ---- one-time: weights + bias into shared memory, in the per-thread order ----------------
    for (int idx = tid; idx < K * kG; idx += NG * kGroupThreads) {
        const int k = idx / kG, col = idx % kG;          // col = q*48 + j  (packed layout)
        const int q = col / kH, j = col % kH;
        w_s[((k * 16 + j / 3) * 4 + q) * 3 + (j % 3)] = wt[idx];
    }
    for (int col = tid; col < kG; col += NG * kGroupThreads) {
        const int q = col / kH, j = col % kH;
        bias_s[((j / 3) * 4 + q) * 3 + (j % 3)] = bias[col];
    }
    if (gt == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint32_t bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    constexpr uint32_t in_bytes = (uint32_t)L::in_stage;
    uint32_t it = 0;                                      // running step counter -> ring stage / parity

    for (int tile = blockIdx.x * NG + g; tile < ntiles; tile += gridDim.x * NG) {
        const int64_t b0 = (int64_t)tile * kTB;
        // reset the recurrent state of the tile
        for (int idx = gt; idx < kTB * kH; idx += kGroupThreads) h_s[idx] = 0.f;
        float c[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int u = 0; u < 3; ++u) c[i][u] = 0.f;
        if (gt == 0) {
            const int pre = T < kStages ? T : kStages;
            for (int s = 0; s < pre; ++s) {
                const uint32_t st = (it + s) % kStages;
                mbar_arrive_expect_tx(&full[st], in_bytes);
                bulk_load(in_s + st * kTB * KIN, in + ((int64_t)s * Bp + b0) * KIN, in_bytes, &full[st]);
            }
        }
        named_bar_sync(bar_b, kGroupThreads);             // h_s zeroed before the first GEMM phase

        for (int t = 0; t < T; ++t, ++it) {
            const uint32_t st = it % kStages;
            mbar_wait(&full[st], (it / kStages) & 1);
            const float* xin = in_s + st * kTB * KIN;

            float acc[4][12];
            {
                const float4* bp = reinterpret_cast<const float4*>(bias_s + ug * 12);
                const float4 b0v = bp[0], b1v = bp[1], b2v = bp[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = b0v.x; acc[i][1] = b0v.y; acc[i][2] = b0v.z; acc[i][3] = b0v.w;
                    acc[i][4] = b1v.x; acc[i][5] = b1v.y; acc[i][6] = b1v.z; acc[i][7] = b1v.w;
                    acc[i][8] = b2v.x; acc[i][9] = b2v.y; acc[i][10] = b2v.z; acc[i][11] = b2v.w;
                }
            }
            // ---- GEMM phase: gates[4 windows][12] += [x_t | h_{t-1}] . Wt --------------------
            auto fma_block = [&](const float* a_base, int a_stride, int kc, const float* w_rows) {
                float4 a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[i] = *reinterpret_cast<const float4*>(a_base + (wg + 8 * i) * a_stride + 4 * kc);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4* wp = reinterpret_cast<const float4*>(w_rows + ((4 * kc + kk) * 16 + ug) * 12);
                    const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                        acc[i][0] = fmaf(av, w0.x, acc[i][0]);
                        acc[i][1] = fmaf(av, w0.y, acc[i][1]);
                        acc[i][2] = fmaf(av, w0.z, acc[i][2]);
                        acc[i][3] = fmaf(av, w0.w, acc[i][3]);
                        acc[i][4] = fmaf(av, w1.x, acc[i][4]);
                        acc[i][5] = fmaf(av, w1.y, acc[i][5]);
                        acc[i][6] = fmaf(av, w1.z, acc[i][6]);
                        acc[i][7] = fmaf(av, w1.w, acc[i][7]);
                        acc[i][8] = fmaf(av, w2.x, acc[i][8]);
                        acc[i][9] = fmaf(av, w2.y, acc[i][9]);
                        acc[i][10] = fmaf(av, w2.z, acc[i][10]);
                        acc[i][11] = fmaf(av, w2.w, acc[i][11]);
                    }
                }
            };
#pragma unroll 2
            for (int kc = 0; kc < KIN / 4; ++kc) fma_block(xin, KIN, kc, w_s);
#pragma unroll 2
            for (int kc = 0; kc < kH / 4; ++kc) fma_block(h_s, kH, kc, w_s + KIN * kG);

            if (gt == 0) bulk_wait_read_all();            // stores of step t-1 no longer read smem
            named_bar_sync(bar_a, kGroupThreads);         // everyone is done with in_s[st] and h_s

            if (gt == 0 && t + kStages < T) {             // refill the stage we just drained
                mbar_arrive_expect_tx(&full[st], in_bytes);
                bulk_load(in_s + st * kTB * KIN, in + ((int64_t)(t + kStages) * Bp + b0) * KIN, in_bytes, &full[st]);
            }

            // ---- activation phase: acc layout per window [i0 i1 i2 f0 | f1 f2 g0 g1 | g2 o0 o1 o2]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int w = wg + 8 * i;
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const float gi = sigmoid_fast(acc[i][u]);
                    const float gf = sigmoid_fast(acc[i][3 + u]);
                    const float gg = tanh_fast(acc[i][6 + u]);
                    const float go = sigmoid_fast(acc[i][9 + u]);
                    c[i][u] = fmaf(gf, c[i][u], gi * gg);
                    const float h = go * tanh_fast(c[i][u]);
                    const int j = 3 * ug + u;
                    h_s[w * kH + j] = h;
                    if (SAVE) {
                        c_s[w * kH + j] = c[i][u];
                        g_s[w * kG + j] = gi;
                        g_s[w * kG + kH + j] = gf;
                        g_s[w * kG + 2 * kH + j] = gg;
                        g_s[w * kG + 3 * kH + j] = go;
                    }
                }
            }
            fence_proxy_async_smem();                     // STS above -> visible to the TMA store
            named_bar_sync(bar_b, kGroupThreads);
            if (gt == 0) {
                const int64_t row0 = (int64_t)t * Bp + b0;
                bulk_store(hout + row0 * kH, h_s, (uint32_t)L::h_buf);
                if (SAVE) {
                    bulk_store(cout + row0 * kH, c_s, (uint32_t)L::c_buf);
                    bulk_store(gates + row0 * kG, g_s, (uint32_t)L::g_buf);
                }
                bulk_commit();
            }
        }
        if (gt == 0) bulk_wait_read_all();                // before h_s is zeroed for the next tile
        named_bar_sync(bar_a, kGroupThreads);
    }
    if (gt == 0) bulk_wait_all();
}

template <int KIN, int NG, bool SAVE>
static int launch_h48(const float* in, const float* wt, const float* bias, float* hout, float* cout, float* gates,
                      int64_t T, int64_t Bp, int sms, cudaStream_t st) {
    const size_t smem = H48Smem<KIN>::total(NG, SAVE);
    auto kern = lstm_fwd_h48_kernel<KIN, NG, SAVE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "lstm_fwd_h48: cannot opt in to %zu B of shared memory (%s)", smem,
                                      cudaGetErrorString(e));
    const int ntiles = (int)(Bp / kTB);
    int grid = (ntiles + NG - 1) / NG;
    if (grid > sms) grid = sms;
    kern<<<grid, NG * kGroupThreads, smem, st>>>(in, wt, bias, hout, cout, gates, (int)T, Bp, ntiles);
    count_launch();
    return check_launch("na_lstm_layer_fwd_f32(h48)");
}

static int g_force_ng = 0;   // test / tuning hook: na_set_tuning("h48_ng", n)

static int device_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// Pick the number of groups per CTA: as many as shared memory allows, but prefer the count that
// wastes the fewest group-slots in the last round (all tiles cost the same).
static int pick_ng(int ntiles, int sms, int max_ng) {
    if (g_force_ng >= 1 && g_force_ng <= max_ng) return g_force_ng;
    int best = 1;
    double best_score = -1.0;
    for (int ng = 1; ng <= max_ng; ++ng) {
        const int64_t slots = (int64_t)sms * ng;
        const int64_t rounds = (ntiles + slots - 1) / slots;
        const double eff = (double)ntiles / (double)(rounds * slots);         // busy fraction of slots
        // more resident groups hide latency better; saturating model of per-SM throughput
        const double per_sm = ng >= 4 ? 1.0 : ng == 3 ? 0.95 : ng == 2 ? 0.80 : 0.45;
        const double used_sms = ntiles >= slots ? 1.0 : (double)((ntiles + ng - 1) / ng) / sms;
        const double score = ntiles >= slots ? eff * per_sm : used_sms * per_sm * ((double)ntiles / (((ntiles + ng - 1) / ng) * ng));
        if (score > best_score + 1e-9) { best_score = score; best = ng; }
    }
    return best;
}

template <int KIN>
static int dispatch_h48(const float* in, const float* wt, const float* bias, float* hout, float* cout,
                        float* gates, int64_t T, int64_t Bp, cudaStream_t st) {
    const bool save = cout != nullptr;
    const int sms = device_sms();
    const int ntiles = (int)(Bp / kTB);
    constexpr size_t kMaxSmem = 227 * 1024;
    int max_ng = 4;
    while (max_ng > 1 && H48Smem<KIN>::total(max_ng, save) > kMaxSmem) --max_ng;
    const int ng = pick_ng(ntiles, sms, max_ng);
#define NA_H48_CASE(NGV)                                                                                  \
    case NGV:                                                                                             \
        return save ? launch_h48<KIN, NGV, true>(in, wt, bias, hout, cout, gates, T, Bp, sms, st)         \
                    : launch_h48<KIN, NGV, false>(in, wt, bias, hout, cout, gates, T, Bp, sms, st);
    switch (ng) {
        NA_H48_CASE(1)
        NA_H48_CASE(2)
        NA_H48_CASE(3)
        NA_H48_CASE(4)
    }
#undef NA_H48_CASE
    return fail(NA_EINVAL, "lstm_fwd_h48: bad group count %d", ng);
}

// Entry used by na_api.cu.  Returns NA_EUNSUPPORTED (without touching the error string) when the
// shape is not one this tier implements, so the caller can fall through to the generic tier.
bool lstm_h48_supported(int64_t K, int64_t H, bool save_c, bool save_g) {
    return H == kH && (K == 8 || K == 48) && (save_c == save_g);
}

int lstm_layer_fwd_h48(const float* in, const float* wt, const float* bias, float* hout, float* cout,
                       float* gates, int64_t T, int64_t Bp, int64_t K, cudaStream_t st) {
    if (K == 8) return dispatch_h48<8>(in, wt, bias, hout, cout, gates, T, Bp, st);
    return dispatch_h48<48>(in, wt, bias, hout, cout, gates, T, Bp, st);
}

void set_h48_groups(int ng) { g_force_ng = ng; }

// out = h * mask * scale  (inter-layer dropout, lstm_eeg_model.py:21), vectorised
__global__ void mask_scale_kernel(const float4* __restrict__ h, const float4* __restrict__ m, float scale,
                                  float4* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = h[i], b = m[i];
        out[i] = make_float4(a.x * b.x * scale, a.y * b.y * scale, a.z * b.z * scale, a.w * b.w * scale);
    }
}

int mask_scale(const float* h, const float* mask, float scale, float* out, int64_t n, cudaStream_t st) {
    const int64_t n4 = n / 4;   // n = T*Bp*48, a multiple of 4
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    mask_scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(h),
                                                        reinterpret_cast<const float4*>(mask), scale,
                                                        reinterpret_cast<float4*>(out), n4);
    count_launch();
    return check_launch("mask_scale");
}

}  // namespace na
