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
