// pdecg_internal.h -- host-side plumbing shared by the translation units of libpdecg.so
#pragma once
#include "../../include/pdecg.h"
#include "pdecg_common.cuh"

namespace pdecg {

// Record a thread-local error message and return `code` (printf-style).
int fail(int code, const char* fmt, ...);
// cudaGetLastError() -> PDECG_OK / PDECG_ERR_CUDA (with message naming `what`).
int check_launch(const char* what);
// Bump the per-process launch counter (pdecg_launch_count()).
void count_launch(int n = 1);

}  // namespace pdecg
