// Fused Adam over a list of small tensors: ONE launch for the whole decoder (16 tensors, 31,764 parameters) instead of
// torch's per-op foreach kernels.  Follows torch.optim.Adam (no amsgrad, no weight decay unless given):
//   m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g g;  p <- p - (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
// `table` (device, int64): per tensor {param ptr, grad ptr, exp_avg ptr, exp_avg_sq ptr, numel}; one CTA per tensor chunk.
#include "na_common.cuh"

namespace na {

__global__ void adam_multi_kernel(const int64_t* __restrict__ table, int n_tensors, float lr, float b1, float b2, float eps,
                                  float weight_decay, float bc1, float rsqrt_bc2, const float* __restrict__ grad_scale) {
    const int ti = blockIdx.y;
    if (ti >= n_tensors) return;
    float* p = reinterpret_cast<float*>(table[5 * ti]);
    const float* g = reinterpret_cast<const float*>(table[5 * ti + 1]);
    float* m = reinterpret_cast<float*>(table[5 * ti + 2]);
    float* v = reinterpret_cast<float*>(table[5 * ti + 3]);
    const int64_t n = table[5 * ti + 4];
    const float gs = grad_scale ? grad_scale[0] : 1.0f;
    const float step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gs;
        const float pi = p[i];
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = fmaf(1.0f - b1, gi - m[i], m[i]);
        const float vi = fmaf(1.0f - b2, gi * gi, b2 * v[i]);
        m[i] = mi;
        v[i] = vi;
        const float denom = fmaf(sqrtf(vi), rsqrt_bc2, eps);
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace na

extern "C" int na_adam_multi(const int64_t* table, int64_t n_tensors, int64_t max_numel, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float bias_correction1, float bias_correction2, const float* grad_scale,
                             na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(n_tensors >= 1 && n_tensors <= 65535 && max_numel >= 1, NA_EINVAL, "na_adam_multi: bad sizes");
    NA_REQUIRE(table != nullptr, NA_EINVAL, "na_adam_multi: null table");
    NA_REQUIRE(bias_correction1 > 0.f && bias_correction2 > 0.f, NA_EINVAL, "na_adam_multi: bias corrections must be positive");
    const unsigned gx = (unsigned)((max_numel + 1023) / 1024 < 64 ? (max_numel + 1023) / 1024 : 64);
    adam_multi_kernel<<<dim3(gx, (unsigned)n_tensors), 256, 0, as_stream(stream)>>>(table, (int)n_tensors, lr, beta1, beta2, eps, weight_decay,
                                                                                   bias_correction1, 1.0f / sqrtf(bias_correction2), grad_scale);
    count_launch();
    return check_launch("na_adam_multi");
}
