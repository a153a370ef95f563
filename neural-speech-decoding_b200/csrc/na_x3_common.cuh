// Shared pieces of the fp32-accurate tensor-core kernels (operands split into fp16 hi + lo, 3 MMAs per product):
// the inference kernel (na_decoder_x3.cu) and the training kernels (na_train_x3.cu).
#pragma once
#include "na_tc_common.cuh"

namespace na {
namespace tc {

constexpr int kX3Threads = 14 * 32;
constexpr int kX3XStages = 3;
constexpr int kX3B0Chunks = 15, kX3B1Chunks = 25;
constexpr int kX3N1 = 208;
constexpr int kX3B1Chunk = kX3N1 * 16;
constexpr int kX3Fc = NA_FC_HIDDEN;
// Input range: raw EEG can carry DC offsets of 1e5 uV, beyond fp16's 65,504.  The producer stores x / 16 and the packed
// W_ih of layer 0 is multiplied by 16 (both exact powers of two): samples up to 1e6 stay finite and values as small as
// 1e-3 keep an absolute error below 5e-7 (fp16 subnormal spacing x 16) after the hi + lo split.
constexpr float kX3XScale = 0.0625f, kX3XScaleInv = 16.0f;
constexpr uint32_t kX3IdescL1 = make_idesc(kX3N1, kFmtVal, kFmtVal);
constexpr uint32_t kX3IdescFlush = make_idesc(16, kFmtVal, kFmtVal);

__device__ __forceinline__ void split16(float v, uint16_t& hi, uint16_t& lo) {
    hi = val16(v);
    lo = val16(v - val16_to_float(hi));
}

__device__ __forceinline__ void x3_tmem_ld2(uint32_t taddr, uint32_t (&v)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// cell update of 4 units at fp32 accuracy; v = [i x4 | f x4 | g x4 | o x4] pre-activations.
// The transcendental pipe bounds this kernel, so the reciprocals are combined algebraically: with E_i = e^-i, E_f = e^-f,
// E_g = e^-2g (sigmoid(x) = 1 / (1 + e^-x), tanh(x) = (1 - e^-2x) / (1 + e^-2x))
//     c' = f c + i g = [ c (1+E_i)(1+E_g) + (1-E_g)(1+E_f) ] / [ (1+E_f)(1+E_i)(1+E_g) ]          3 ex2 + 1 rcp
//     h  = o tanh(c') = (1 - E_c) / [ (1+E_o)(1+E_c) ],  E_c = e^-2c'                              2 ex2 + 1 rcp
// = 7 MUFU per cell instead of 10 (ex2 + rcp per activation).  Pre-activations are clamped where the functions are
// already saturated in fp32 (|x| <= 28 for the sigmoids, <= 14 for the tanh arguments) so the products stay below 2^127.
// The MUFU pipe (16 results / clk / SM) bounds these kernels while the FMA pipe idles, so a reciprocal can be taken off it:
// rcp_fma = an exponent-trick seed (12 % off) refined by two cubically convergent steps r <- r (1 + e + e^2), e = 1 - d r
// (0.125^9 = 7e-9, then fp32 rounding: as accurate as rcp.approx) -- 1 integer op + 6 FFMA instead of 1 MUFU.
// The denominators here are products of (1 + E) terms: positive, normal, <= 2^120; a NaN stays a NaN.
__device__ __forceinline__ float rcp_fma(float d) {
    float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
    float e = fmaf(-d, r, 1.0f);
    r = fmaf(r, fmaf(e, e, e), r);
    e = fmaf(-d, r, 1.0f);
    r = fmaf(r, fmaf(e, e, e), r);
    return r;
}
// NR = how many of the two reciprocals of a cell update run on the FMA pipe: 0 = none (7 MUFU per cell), 1 = the one of h
// (6 MUFU), 2 = both (5 MUFU), 3 = alternate 1 / 2 over the units (5.5 MUFU).  Runtime knob "x3_rcp_fma" (A/B timing).
// MEASURED (B200, round 2, scripts/time_offload.py, 18,944 windows = one full round): NR = 0: 2.359 ms, 1: 2.599, 2: 2.664,
// 3: 2.708 -- the offload LOSES 10-15 %: with 3 epilogue warps per SM sub-partition the step is bound by the dependent chain
// of each warp (ex2 -> products -> rcp -> ex2 -> rcp), not by MUFU issue, and the 7 serially dependent instructions of the
// FMA-pipe reciprocal lengthen exactly that chain.  Same accuracy (8.9e-7 vs 9.2e-7 against the FFMA kernels).  Default: off.
constexpr int kX3DefaultNR = 0;

template <int NR = kX3DefaultNR>
__device__ __forceinline__ void cell_granule_exact(const uint32_t* v, float* c, float* h) {
    // an activation costs the scale (FMUL), one NaN-propagating min and one ex2: the cap keeps the products below 2^127,
    // the lower side needs none (E -> 0).  cap 40: e^-x <= 2^40 <=> x >= -27.7, where sigmoid is 9e-13 and tanh is -1 to
    // fp32 precision.  (Folding the scale into the packed weights saves the FMUL and 3 % of the time, but the rounding of
    // the scaled weights pushed the worst case of 70,000 random short windows to 1.03e-5 against the FFMA kernels.)
    constexpr float kL2e = 1.4426950408889634f;
    auto capped_ex2 = [](float a) {
        float r;
        asm("min.NaN.f32 %0, %1, 0f42200000;" : "=f"(r) : "f"(a));        // min(a, 40.0f), NaN stays NaN
        return ex2_approx(r);
    };
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float ei = capped_ex2(-kL2e * __uint_as_float(v[u])), ef = capped_ex2(-kL2e * __uint_as_float(v[4 + u]));
        const float eg = capped_ex2(-2.0f * kL2e * __uint_as_float(v[8 + u])), eo = capped_ex2(-kL2e * __uint_as_float(v[12 + u]));
        const float dig = (1.0f + ei) * (1.0f + eg);
        const float df = 1.0f + ef;
        const float num = fmaf(c[u], dig, (1.0f - eg) * df);
        const bool c_on_fma = NR == 2 || (NR == 3 && (u & 1));
        const float cn = num * (c_on_fma ? rcp_fma(df * dig) : rcp_approx(df * dig));
        c[u] = cn;
        const float ec = capped_ex2(-2.0f * kL2e * cn);
        const float dh = (1.0f + eo) * (1.0f + ec);
        h[u] = (1.0f - ec) * (NR >= 1 ? rcp_fma(dh) : rcp_approx(dh));
    }
}

// 8 fp32 values -> hi chunk entry and lo chunk entry (4 x fp16x2 each)
__device__ __forceinline__ void split_pack8(const float* h, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        hi[u] = pack_val(h[2 * u], h[2 * u + 1]);
        lo[u] = pack_val(h[2 * u] - val_lo(hi[u]), h[2 * u + 1] - val_hi(hi[u]));
    }
}


}  // namespace tc
}  // namespace na
