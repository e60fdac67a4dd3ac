// b2048_internal.h — things shared by the .cu translation units of libb2048 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/b2048.h"

struct b2048_handle {
    uint8_t* d_tables;   // [131072 B lut_left (u16 x 65536)] [65536 B lut_merge (u8 x 65536)] [small tables]
    int device;
    int num_sms;
    int smem_optin;      // max opt-in dynamic shared memory per block
    uint8_t* tc_image;   // bf16 weight image of the tensor-core policy kernel (lazily allocated)
    uint8_t* hp_image;   // split-fp16 weight image of the float32-grade forward kernel (lazily allocated)
    uint8_t* gen_image;  // weight images of the shape-generic tensor-core kernels (b2048_mlp_gen.cu; lazily allocated)
    size_t gen_image_bytes;
    unsigned pipe_split; // B2048_DBG_PARAM_PIPE_SPLIT: CTAs per role of the update pipeline (0 = default split)
    unsigned debug;      // bit B2048_DBG_* set through b2048_debug_set (test / profiling switches; 0 in production)
    unsigned attrs;      // bit k set: the opt-in shared-memory attribute of kernel family k has been set on THIS device
                         // (cudaFuncSetAttribute is per device; a process may hold handles on several)
};

#define B2048_LUT_LEFT_BYTES 131072
#define B2048_LUT_MERGE_BYTES 65536
#define B2048_LUT_BYTES (B2048_LUT_LEFT_BYTES + B2048_LUT_MERGE_BYTES)
// the small tables of b2048_step_fast.cuh follow the row tables in the same allocation
#define B2048_TABLES_BYTES (B2048_LUT_BYTES + 2048 + 128 + 64)

namespace b2 {

void set_error(const std::string& msg);
int fail(b2048_status st, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);

inline const uint16_t* lut_left_ptr(const b2048_handle* h) { return reinterpret_cast<const uint16_t*>(h->d_tables); }
inline const uint8_t* lut_merge_ptr(const b2048_handle* h) { return h->d_tables + B2048_LUT_LEFT_BYTES; }

}  // namespace b2

#define B2_CUDA(expr)                                              \
    do {                                                           \
        int _st = b2::check_cuda((expr), #expr);                   \
        if (_st != B2048_OK) return _st;                           \
    } while (0)

#define B2_REQUIRE(cond, msg)                                      \
    do {                                                           \
        if (!(cond)) return b2::fail(B2048_ERR_INVALID, msg);      \
    } while (0)
