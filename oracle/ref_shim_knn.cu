// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// C-ABI shim around the unmodified reference simple-knn
// (/root/reference/submodules/simple-knn/simple_knn.cu), replacing spatial.cu:15-26.
#include <cstdint>
#include <cfloat>
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_dist2(int P, const float* points_dev, float* mean_dists_dev)
{
    if (P == 0) return 0;
    SimpleKNN::knn(P, (float3*)points_dev, mean_dists_dev);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}
