// decode_kernel.cuh -- placeholder until the decode kernels land (next commit).
#pragma once
#include "qb_common.cuh"
namespace qb
{
#ifndef QB_EMU
    inline cudaError_t dec_set_attrs() { return cudaSuccess; }
#endif
}
