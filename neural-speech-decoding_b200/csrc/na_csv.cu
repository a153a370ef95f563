// ingestion of the collector's CSV windows ON the GPU.
//
// The reference writes every window as text, np.savetxt(fmt="%.7f", delimiter=",") of a [625, 8] array
// (Neural_decoding_data_collector.py:129-139), and reads it back with np.loadtxt(dtype=float32).  Here the raw file
// BYTES go to the device in one copy (an offsets table delimits the files) and one kernel turns them into
// fp32 [N][fields]: one CTA per file, the file staged in shared memory, every thread owns a byte range, a block
// scan of the delimiter counts gives each range its first field index, and the thread parses the fields that START
// in its range.
//
// Bit-exactness with np.loadtxt: numpy parses the decimal string to a correctly rounded double (strtod) and then
// rounds to float32.  A "%.7f" field has <= 18 significant digits here, so mantissa (an exact integer below 2^63,
// exact in double below 2^53 for every value the collector can write) / 10^frac_digits evaluated with ONE IEEE
// double division is that same correctly rounded double; __double2float_rn is the same second rounding.
// Fields with more than 15 significant digits, exponents, nan/inf or any other character are counted in
// status[2n+1] and the host wrapper raises -- there is no silent approximation and no CPU fallback.
// HBM-bound byte work: algorithmic bytes = text bytes read + 4 bytes per value written.
#include "na_common.cuh"

namespace na {

constexpr int kCsvThreads = 256;
__constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};     // all exact doubles

__device__ __forceinline__ bool csv_is_delim(unsigned char c) { return c == ',' || c == '\n'; }

__global__ void __launch_bounds__(kCsvThreads)
csv_parse_kernel(const unsigned char* __restrict__ text, const int64_t* __restrict__ offsets, float* __restrict__ out,
                 int* __restrict__ status, int fields_per_file, int starts_off) {
    extern __shared__ __align__(16) unsigned char smem_csv[];
    __shared__ int warp_tot[kCsvThreads / 32];
    __shared__ int s_bad;
    const int n = blockIdx.x, tid = threadIdx.x;
    const int64_t beg = offsets[n];
    const int len = (int)(offsets[n + 1] - beg);
    if (tid == 0) s_bad = 0;
    // stage the file with 16-byte loads: copy the aligned span that covers [beg, beg + len) and keep the
    // misalignment as an index shift (files start at arbitrary byte offsets; the over-read stays inside the 16-byte
    // granules that hold the file's first and last byte, i.e. inside the caller's allocation granularity)
    const int shift = (int)(beg & 15);
    const uint4* src16 = reinterpret_cast<const uint4*>(text + (beg - shift));
    const int n16 = (shift + len + 15) >> 4;
    for (int i = tid; i < n16; i += kCsvThreads) reinterpret_cast<uint4*>(smem_csv)[i] = __ldg(src16 + i);
    const unsigned char* buf = smem_csv + shift;
    __syncthreads();

    // a field starts at byte 0 and after every delimiter, provided a non-delimiter, non-blank byte follows
    // (np.loadtxt ignores blank lines / a trailing newline).
    // Pass 1 (byte-parallel, uniform trip counts): every thread counts the field starts in its contiguous byte range,
    // a block scan turns the counts into first-field indices, the thread writes the start positions of its fields.
    // Pass 2 (field-parallel): thread f parses field f -- neighbouring lanes parse fields of (nearly) the same length,
    // so the warp stays converged; the first version parsed inside the byte ranges and ran 8x slower.
    uint32_t* starts = reinterpret_cast<uint32_t*>(smem_csv + starts_off);
    const int per = (len + kCsvThreads - 1) / kCsvThreads;
    const int lo = min(len, tid * per), hi = min(len, lo + per);
    auto starts_field = [&](int i) -> bool {
        if (i > 0 && !csv_is_delim(buf[i - 1])) return false;
        const unsigned char c = buf[i];
        return !(csv_is_delim(c) || c == '\r' || c == ' ');       // empty field / blank line / CR-LF
    };
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += starts_field(i) ? 1 : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
    int total = 0;
    for (int w = 0; w < kCsvThreads / 32; ++w) total += warp_tot[w];
    int field = base + incl - cnt;
    for (int i = lo; i < hi; ++i)
        if (starts_field(i)) {
            if (field < fields_per_file) starts[field] = (uint32_t)i;
            ++field;
        }
    __syncthreads();

    int bad = 0;
    const int nf = min(total, fields_per_file);
    for (int f = tid; f < nf; f += kCsvThreads) {
        // parse [sign] digits [. digits]
        int p = (int)starts[f];
        bool neg = false;
        if (buf[p] == '-') { neg = true; ++p; } else if (buf[p] == '+') ++p;
        unsigned long long mant = 0;
        int sig = 0, frac = 0, ndig = 0;
        bool dot = false, ok = true;
        for (; p < len && !csv_is_delim(buf[p]); ++p) {
            const unsigned char c = buf[p];
            if (c >= '0' && c <= '9') {
                ++ndig;
                if (sig > 0 || c != '0') ++sig;
                if (sig <= 18) { mant = mant * 10ull + (unsigned)(c - '0'); if (dot) ++frac; }
                else ok = false;                              // more digits than one exact integer can hold
            } else if (c == '.' && !dot) dot = true;
            else if (c == '\r' && (p + 1 == len || buf[p + 1] == '\n')) { /* CR of a CR-LF line end */ }
            else ok = false;
        }
        if (ndig == 0 || sig > 15 || frac > 22) ok = false;       // mantissa must be exact in double, 10^frac too
        float v = 0.f;
        if (ok) {
            const double d = __ddiv_rn((double)mant, kPow10[frac]);
            v = __double2float_rn(neg ? -d : d);
        } else ++bad;
        out[(int64_t)n * fields_per_file + f] = v;
    }
    if (bad) atomicAdd(&s_bad, bad);
    __syncthreads();
    if (tid == 0) { status[2 * n] = total; status[2 * n + 1] = s_bad; }
}

}  // namespace na

extern "C" int na_csv_parse_f32(const void* text, const int64_t* offsets, float* out, int* status, int64_t n_files,
                                int64_t fields_per_file, int64_t max_file_bytes, na_stream_t stream) {
    using namespace na;
    NA_REQUIRE(n_files >= 1 && n_files < (1ll << 31) && fields_per_file >= 1 && fields_per_file < (1ll << 31), NA_EINVAL,
               "na_csv_parse_f32: bad shape n_files=%lld fields_per_file=%lld", (long long)n_files, (long long)fields_per_file);
    NA_REQUIRE(max_file_bytes >= 1 && max_file_bytes + 4 * fields_per_file <= 200 * 1024, NA_EUNSUPPORTED,
               "na_csv_parse_f32: a file of %lld bytes with %lld fields does not fit the 200 KB shared-memory stage",
               (long long)max_file_bytes, (long long)fields_per_file);
    NA_REQUIRE(text != nullptr && offsets != nullptr && status != nullptr, NA_EINVAL, "na_csv_parse_f32: null pointer");
    NA_REQUIRE_PTR(out);
    const size_t text_smem = (size_t)((max_file_bytes + 15 + 15) / 16 * 16);      // + the start misalignment
    const size_t smem = text_smem + 4 * (size_t)fields_per_file;                  // + the field start positions
    NA_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15u) == 0, NA_EALIGN, "na_csv_parse_f32: text not 16-byte aligned");
    cudaError_t e = cudaFuncSetAttribute(csv_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "na_csv_parse_f32: shared memory opt-in failed (%s)", cudaGetErrorString(e));
    csv_parse_kernel<<<(unsigned)n_files, kCsvThreads, smem, as_stream(stream)>>>(
        reinterpret_cast<const unsigned char*>(text), offsets, out, status, (int)fields_per_file, (int)text_smem);
    count_launch();
    return check_launch("na_csv_parse_f32");
}
