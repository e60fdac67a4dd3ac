// b2048_capi.cu — error reporting and version of the C ABI (include/b2048.h).
#include "b2048_internal.h"

namespace b2 {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(b2048_status st, const std::string& msg) {
    set_error(msg);
    return (int)st;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return B2048_OK;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return (int)B2048_ERR_CUDA;
}

}  // namespace b2

extern "C" const char* b2048_last_error(void) { return b2::g_last_error.c_str(); }

extern "C" int b2048_version(void) { return 100; }

extern "C" int b2048_debug_set(b2048_handle* h, int32_t option, int32_t value) {
    B2_REQUIRE(h != nullptr, "b2048_debug_set: handle is NULL");
    if (option == B2048_DBG_PARAM_PIPE_SPLIT) {
        h->pipe_split = (unsigned)value;
        return B2048_OK;
    }
    B2_REQUIRE(option >= 0 && option < B2048_DBG_COUNT, "b2048_debug_set: unknown option");
    if (value) h->debug |= 1u << option;
    else h->debug &= ~(1u << option);
    return B2048_OK;
}
