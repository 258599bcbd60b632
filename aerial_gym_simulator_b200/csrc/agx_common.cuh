// Error plumbing shared by the translation units of libaerial_gym_b200.so.
#pragma once
#include <cuda_runtime.h>

// records a thread-local message, returns `code`
int agx_set_error(int code, const char* fmt, ...);
// cudaPeekAtLastError after a launch -> AGX_OK or AGX_E_CUDA (message recorded)
int agx_check_launch(const char* what);
int agx_check_cuda(cudaError_t e, const char* what);
