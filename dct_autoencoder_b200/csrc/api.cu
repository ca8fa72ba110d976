// Error state and library identification for libdcta.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dcta {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool g_profile_on = false;
cudaStream_t g_last_stream = nullptr;

}  // namespace dcta

extern "C" const char* dcta_last_error(void) { return dcta::g_err; }
extern "C" int dcta_abi_version(void) { return 3; }
extern "C" int dcta_compiled_arch(void) { return 100; }

// ------------------------------------------------------------------------------ basis tables (host side)
// The orthonormal DCT-II matrix C_n[q, m] = s_q cos(pi (2m + 1) q / 2n), s_0 = sqrt(1/n), s_q = sqrt(2/n)
// (the definition behind util.py:333-338 / torch_dct "ortho"), evaluated in double and laid out as each GEMM path
// reads it.  Pure host arithmetic into caller-provided HOST buffers: the caller uploads them once per
// (n, k, layout) and keeps the device copies (the "basis-table cache handle" of the boundary is the caller's).
#include <cuda_fp16.h>
#include <math.h>

#include <vector>

namespace dcta {
static const double kBasisScale = 1024.0;      // kScaleBasis (gemm_tc.cu) == kFScaleBasis (dct_fold.cu)

static double basis_at(int n, int q, int m) {
    if (q == 0) return sqrt(1.0 / n);
    return cos(M_PI * (2.0 * m + 1.0) * q / (2.0 * n)) * sqrt(2.0 / n);
}
static void split_store(double v, __half* hi, __half* lo, int64_t i) {
    const __half h = __double2half(v);
    hi[i] = h;
    lo[i] = __double2half(v - (double)__half2float(h));
}
static int64_t round8(int64_t v) { return (v + 7) / 8 * 8; }
}  // namespace dcta

extern "C" int64_t dcta_basis_elems(int layout, int n, int k) {
    using namespace dcta;
    switch (layout) {
        case DCTA_BASIS_F32: return (int64_t)k * n;
        case DCTA_BASIS_SPLIT_FWD: return (int64_t)k * round8(n);
        case DCTA_BASIS_SPLIT_INV: return (int64_t)n * round8(k);
        case DCTA_BASIS_FOLD_FWD: return 2ll * (k / 2) * round8((n + 1) / 2);
        case DCTA_BASIS_FOLD_INV: return 2ll * ((n + 1) / 2) * round8(k / 2);
    }
    return -1;
}

extern "C" int dcta_basis_init(int layout, int n, int k, void* hi_host, void* lo_host, float* row_scale_host) {
    using namespace dcta;
    DCTA_REQUIRE(n > 0 && k > 0 && k <= n, "basis_init: needs 0 < k <= n");
    DCTA_REQUIRE(hi_host != nullptr, "basis_init: null output");
    if (layout == DCTA_BASIS_F32) {
        float* out = static_cast<float*>(hi_host);
        for (int q = 0; q < k; ++q)
            for (int m = 0; m < n; ++m) out[(int64_t)q * n + m] = (float)basis_at(n, q, m);
        return DCTA_OK;
    }
    DCTA_REQUIRE(lo_host != nullptr, "basis_init: null lo plane");
    __half* hi = static_cast<__half*>(hi_host);
    __half* lo = static_cast<__half*>(lo_host);
    const int64_t elems = dcta_basis_elems(layout, n, k);
    DCTA_REQUIRE(elems > 0, "basis_init: unknown layout %d", layout);
    for (int64_t i = 0; i < elems; ++i) { hi[i] = __double2half(0.0); lo[i] = hi[i]; }
    switch (layout) {
        case DCTA_BASIS_SPLIT_FWD: {          // (k, round8(n)): C * 2^10, row 0 stored as the exact constant 32
            DCTA_REQUIRE(row_scale_host != nullptr, "basis_init: forward layouts need row_scale");
            const int64_t ld = round8(n);
            for (int q = 0; q < k; ++q) {
                for (int m = 0; m < n; ++m) split_store(q == 0 ? 32.0 : basis_at(n, q, m) * kBasisScale, hi, lo, q * ld + m);
                row_scale_host[q] = (float)(q == 0 ? sqrt(1.0 / n) / 32.0 : 1.0 / kBasisScale);
            }
            return DCTA_OK;
        }
        case DCTA_BASIS_SPLIT_INV: {          // (n, round8(k)): (C * 2^10)^T
            const int64_t ld = round8(k);
            for (int m = 0; m < n; ++m)
                for (int q = 0; q < k; ++q) split_store(basis_at(n, q, m) * kBasisScale, hi, lo, m * ld + q);
            return DCTA_OK;
        }
        case DCTA_BASIS_FOLD_FWD: {           // (2, k/2, round8(ceil(n/2))): group g = rows of parity g, first half of the
            // samples (for odd n including the middle one, which pairs with itself: its odd rows are zero)
            DCTA_REQUIRE(row_scale_host != nullptr, "basis_init: forward layouts need row_scale");
            DCTA_REQUIRE(k % 2 == 0, "basis_init: the folded layouts need an even k");
            const int k2 = k / 2, n2 = (n + 1) / 2;
            const int64_t ldn = round8(n2);
            for (int g = 0; g < 2; ++g)
                for (int j = 0; j < k2; ++j) {
                    const int q = 2 * j + g;
                    for (int m = 0; m < n2; ++m) {
                        const bool zero = (n & 1) && (q & 1) && m == n2 - 1;      // cos(pi q / 2) of an odd q
                        split_store(q == 0 ? 32.0 : (zero ? 0.0 : basis_at(n, q, m) * kBasisScale), hi, lo,
                                    ((int64_t)g * k2 + j) * ldn + m);
                    }
                    row_scale_host[g * k2 + j] = (float)(q == 0 ? sqrt(1.0 / n) / 32.0 : 1.0 / kBasisScale);
                }
            return DCTA_OK;
        }
        case DCTA_BASIS_FOLD_INV: {           // (2, ceil(n/2), round8(k/2)): the transposes
            DCTA_REQUIRE(k % 2 == 0, "basis_init: the folded layouts need an even k");
            const int k2 = k / 2, n2 = (n + 1) / 2;
            const int64_t ld = round8(k2);
            for (int g = 0; g < 2; ++g)
                for (int m = 0; m < n2; ++m)
                    for (int j = 0; j < k2; ++j) {
                        const bool zero = (n & 1) && g == 1 && m == n2 - 1;
                        split_store(zero ? 0.0 : basis_at(n, 2 * j + g, m) * kBasisScale, hi, lo, ((int64_t)g * n2 + m) * ld + j);
                    }
            return DCTA_OK;
        }
    }
    return DCTA_ERR_INVALID_ARG;
}


// ------------------------------------------------------------------------------ per-launch device times
// dcta_profile_begin(stream) records a first event on `stream`; from then on every launch group the library checks
// (check_launch) is followed by an event on its stream.  dcta_profile_end() waits for the last event and returns the
// elapsed milliseconds between consecutive events together with the names given to check_launch: the device time of
// each launch group, measured with CUDA events on the launching stream, without a profiler attached.
#include <string>

namespace dcta {
static std::vector<cudaEvent_t> g_prof_events;
static std::vector<std::string> g_prof_names;

void profile_mark(const char* what) {
    if (g_prof_events.size() >= 4096) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, g_last_stream);
    g_prof_events.push_back(ev);
    g_prof_names.push_back(what);
}
}  // namespace dcta

extern "C" int dcta_profile_begin(void* stream) {
    using namespace dcta;
    for (cudaEvent_t ev : g_prof_events) cudaEventDestroy(ev);
    g_prof_events.clear();
    g_prof_names.clear();
    g_last_stream = reinterpret_cast<cudaStream_t>(stream);
    g_profile_on = true;
    profile_mark("begin");
    return DCTA_OK;
}

// names_out: the names joined by '\n' (truncated to names_cap bytes, NUL-terminated); ms_out: up to max_n durations.
// Returns the number of launch groups recorded (<= max_n written), or a negative status.
extern "C" int dcta_profile_end(char* names_out, int names_cap, float* ms_out, int max_n) {
    using namespace dcta;
    g_profile_on = false;
    const int n = (int)g_prof_events.size() - 1;
    if (n < 0) return 0;
    cudaError_t e = cudaEventSynchronize(g_prof_events.back());
    if (e != cudaSuccess) { set_error("profile_end: %s", cudaGetErrorString(e)); return DCTA_ERR_LAUNCH; }
    std::string names;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_prof_events[i], g_prof_events[i + 1]);
        if (ms_out && i < max_n) ms_out[i] = ms;
        if (i) names += '\n';
        names += g_prof_names[i + 1];
    }
    if (names_out && names_cap > 0) {
        const size_t len = names.size() < (size_t)names_cap - 1 ? names.size() : (size_t)names_cap - 1;
        memcpy(names_out, names.data(), len);
        names_out[len] = 0;
    }
    for (cudaEvent_t ev : g_prof_events) cudaEventDestroy(ev);
    g_prof_events.clear();
    g_prof_names.clear();
    return n;
}
