// Shared helpers for the artiboost_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/artiboost_b200.h"

namespace ab {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool profile_enabled();
int profile_begin(int stage, cudaStream_t st);
void profile_end(int token, cudaStream_t st);

// RAII bracket around one kernel launch for ab_profile_collect (no-op unless profiling is enabled)
struct StageTimer {
    int token;
    cudaStream_t st;
    StageTimer(int stage, cudaStream_t s) : token(profile_begin(stage, s)), st(s) {}
    ~StageTimer() { profile_end(token, st); }
};

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return AB_OK;
}

#define AB_REQUIRE(cond, msg)                                  \
    do {                                                       \
        if (!(cond)) {                                         \
            ab::set_error("%s: %s", __func__, msg);            \
            return AB_ERR_ARG;                                 \
        }                                                      \
    } while (0)

#define AB_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            ab::set_error("%s: %s", #call, cudaGetErrorString(e__));           \
            return (int)e__;                                                   \
        }                                                                      \
    } while (0)

__host__ __device__ static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace ab
