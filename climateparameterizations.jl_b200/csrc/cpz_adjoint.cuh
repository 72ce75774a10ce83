#pragma once
#include "cpz_device.cuh"
namespace cpz {
inline size_t adjoint_other_smem(int S, int nbc, int CT) { return ((size_t)4 * S * CT + (size_t)CT * (S + 4) + nbc * CT + CT + 64) * sizeof(float); }
}
