// Status / error plumbing of the C ABI.
#include <cstdarg>

#include "common.cuh"

namespace sanerf {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return status;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("SANERF_PDL");
        return e == nullptr || e[0] != '0';
    }();
    return on;
}

}  // namespace sanerf

extern "C" int sanerf_abi_version(void) { return SANERF_ABI_VERSION; }

extern "C" const char* sanerf_last_error(void) { return sanerf::error_buffer(); }

extern "C" const char* sanerf_status_string(int status) {
    switch (status) {
        case SANERF_OK: return "ok";
        case SANERF_ERR_INVALID_ARG: return "invalid argument";
        case SANERF_ERR_NULL_POINTER: return "null pointer";
        case SANERF_ERR_CUDA: return "CUDA error";
        case SANERF_ERR_MISALIGNED: return "misaligned buffer";
        default: return "unknown status";
    }
}

#ifdef SANERF_HEAD_TRACE
#include <vector>
namespace sanerf {
static std::vector<void (*)(StampBuf*)>& stamp_setters() { static std::vector<void (*)(StampBuf*)> v; return v; }
void register_stamp_tu(void (*set)(StampBuf*)) { stamp_setters().push_back(set); }
}  // namespace sanerf
// op 0: allocate the stamp buffer and hand it to every translation unit; 1: reset; 2: copy {count, stamps} to `out`
extern "C" __attribute__((visibility("default"))) int sanerf_debug_stamps(int op, void* out) {
    static sanerf::StampBuf* buf = nullptr;
    if (op == 0) {
        if (buf == nullptr && cudaMalloc(&buf, sizeof(sanerf::StampBuf)) != cudaSuccess) return 1;
        cudaMemset(buf, 0, sizeof(sanerf::StampBuf));
        for (auto set : sanerf::stamp_setters()) set(buf);
        return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
    }
    if (buf == nullptr) return 1;
    if (op == 1) return cudaMemset(buf, 0, 8) == cudaSuccess ? 0 : 1;
    return cudaMemcpy(out, buf, sizeof(sanerf::StampBuf), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}
#endif
