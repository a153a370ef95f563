// the collector-side filter chain (Neural_decoding_data_collector.py:109-127) on the GPU.
// Per (window, channel) series: detrend(CONSTANT), then for every filter of the chain a "zero phase" application
// = forward DirectFormII biquad cascade, time reversal, the same cascade again (BrainFlow re-uses the filter object,
// so by default the state of the forward pass is carried into the backward pass), reversal; finally np.round(., d).
//
// One thread per series, float64 like BrainFlow.  An IIR is serial in time, so the parallelism is the B x C series;
// the 2 x nfilt passes read and write the series in a per-CTA tiled [T][128 series] scratch (coalesced, contiguous per CTA),
// the first pass reads the fp32 window directly, the last one writes the fp32 result.  HBM-bound byte work:
// algorithmic bytes per series = T x (4 + 4 + (2 nfilt - 1) x 16).
// PARITY UNPINNED (brainflow==5.19.0 is absent from this image): checked against oracle/filter_chain.py, a scipy
// restatement of BrainFlow's published algorithm.
#include "na_common.cuh"

namespace na {

constexpr int kIirMaxF = 8, kIirMaxS = 8;

__global__ void __launch_bounds__(128)
iir_chain_kernel(const float* __restrict__ x, float* __restrict__ y, double* __restrict__ scratch,
                 const double* __restrict__ coef, const int* __restrict__ nsec, int nfilt, int64_t S, int T, int C,
                 int detrend, int round_decimals, int carry_state) {
    __shared__ double s_coef[kIirMaxF * kIirMaxS * 5];
    __shared__ int s_nsec[kIirMaxF];
    int total = 0;
    for (int f = 0; f < nfilt; ++f) total += nsec[f];
    for (int i = threadIdx.x; i < total * 5; i += blockDim.x) s_coef[i] = coef[i];
    if (threadIdx.x < nfilt) s_nsec[threadIdx.x] = nsec[threadIdx.x];
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int64_t b = s / C;
    const int c = (int)(s % C);
    const float* xs = x + (b * T) * C + c;           // element t at xs[t * C]
    float* ys = y + (b * T) * C + c;
    // scratch is tiled per CTA: [S / 128][T][128] -- a CTA walks ONE contiguous T x 1 KB region forwards and backwards
    // (DRAM-page and L2 friendly; the first layout, [T][S], scattered 256-byte accesses over the whole buffer)
    double* sc = scratch + (int64_t)blockIdx.x * T * 128 + threadIdx.x;
    double mean = 0.0;
    if (detrend) {
        for (int t = 0; t < T; ++t) mean += (double)xs[(int64_t)t * C];
        mean /= (double)T;
    }
    double scale = 1.0;
    for (int k = 0; k < round_decimals; ++k) scale *= 10.0;
    if (nfilt == 0) {
        for (int t = 0; t < T; ++t) {
            double v = (double)xs[(int64_t)t * C] - mean;
            if (round_decimals >= 0) v = rint(v * scale) / scale;
            ys[(int64_t)t * C] = (float)(v == 0.0 ? 0.0 : v);
        }
        return;
    }
    int cbase = 0;
    for (int f = 0; f < nfilt; ++f) {
        const int ns = s_nsec[f];
        double b0[kIirMaxS], b1[kIirMaxS], b2[kIirMaxS], a1[kIirMaxS], a2[kIirMaxS], w1[kIirMaxS], w2[kIirMaxS];
#pragma unroll
        for (int k = 0; k < kIirMaxS; ++k) {
            const bool on = k < ns;
            const double* cf = s_coef + (cbase + (on ? k : 0)) * 5;
            b0[k] = on ? cf[0] : 1.0; b1[k] = on ? cf[1] : 0.0; b2[k] = on ? cf[2] : 0.0;
            a1[k] = on ? cf[3] : 0.0; a2[k] = on ? cf[4] : 0.0;
            w1[k] = 0.0; w2[k] = 0.0;
        }
        cbase += ns;
        const bool last = (f == nfilt - 1);
        // forward pass (8 samples fetched ahead of the serial recurrence: the loads are independent of it)
        for (int t0 = 0; t0 < T; t0 += 8) {
            double buf[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = t0 + j;
                buf[j] = t < T ? ((f == 0) ? (double)xs[(int64_t)t * C] - mean : sc[(int64_t)t * 128]) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = t0 + j;
                if (t >= T) break;
                double v = buf[j];
#pragma unroll
                for (int k = 0; k < kIirMaxS; ++k) {
                    if (k < ns) {                               // DirectFormII: w = v - a1 w1 - a2 w2; out = b0 w + b1 w1 + b2 w2
                        const double w = v - a1[k] * w1[k] - a2[k] * w2[k];
                        v = b0[k] * w + b1[k] * w1[k] + b2[k] * w2[k];
                        w2[k] = w1[k];
                        w1[k] = w;
                    }
                }
                sc[(int64_t)t * 128] = v;
            }
        }
        if (!carry_state) {
#pragma unroll
            for (int k = 0; k < kIirMaxS; ++k) { w1[k] = 0.0; w2[k] = 0.0; }
        }
        // backward pass (the reversed series through the same cascade)
        for (int t0 = T - 1; t0 >= 0; t0 -= 8) {
            double buf[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = t0 - j;
                buf[j] = t >= 0 ? sc[(int64_t)t * 128] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = t0 - j;
                if (t < 0) break;
                double v = buf[j];
#pragma unroll
                for (int k = 0; k < kIirMaxS; ++k) {
                    if (k < ns) {
                        const double w = v - a1[k] * w1[k] - a2[k] * w2[k];
                        v = b0[k] * w + b1[k] * w1[k] + b2[k] * w2[k];
                        w2[k] = w1[k];
                        w1[k] = w;
                    }
                }
                if (last) {
                    if (round_decimals >= 0) v = rint(v * scale) / scale;      // np.round: multiply, rint, divide
                    ys[(int64_t)t * C] = (float)(v == 0.0 ? 0.0 : v);          // the collector also clears negative zero
                } else {
                    sc[(int64_t)t * 128] = v;
                }
            }
        }
    }
}

}  // namespace na

extern "C" int na_iir_chain(const float* x, float* y, double* scratch, const double* coef, const int* nsec,
                            int64_t nfilt, int64_t B, int64_t T, int64_t C, int detrend, int round_decimals,
                            int carry_state, na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(B >= 1 && T >= 1 && T < (1 << 30) && C >= 1 && C < (1 << 20), NA_EINVAL,
               "na_iir_chain: bad shape B=%lld T=%lld C=%lld", (long long)B, (long long)T, (long long)C);
    NA_REQUIRE(nfilt >= 0 && nfilt <= kIirMaxF, NA_EUNSUPPORTED, "na_iir_chain: %lld filters (at most %d)", (long long)nfilt, kIirMaxF);
    NA_REQUIRE(round_decimals <= 22, NA_EINVAL, "na_iir_chain: round_decimals=%d", round_decimals);
    NA_REQUIRE_PTR(x); NA_REQUIRE_PTR(y);
    NA_REQUIRE(nfilt == 0 || (scratch != nullptr && coef != nullptr && nsec != nullptr), NA_EINVAL, "na_iir_chain: null pointer");
    const int64_t S = B * C;
    iir_chain_kernel<<<(unsigned)((S + 127) / 128), 128, 0, as_stream(stream)>>>(x, y, scratch, coef, nsec, (int)nfilt, S, (int)T,
                                                                              (int)C, detrend, round_decimals, carry_state);
    count_launch();
    return check_launch("na_iir_chain");
}
