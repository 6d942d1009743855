// Stub of the few OptiX names RayJoin's non-RT headers touch (test infrastructure;
// lets the reference's grid / lbvh backends compile without the OptiX SDK, which
// has no function on a GPU without RT cores).  Pullers: src/util/exception.h:33-34,
// src/util/helpers.h:7,35,94-95, src/app/query_config.h:20.
#pragma once
typedef unsigned long long OptixTraversableHandle;
typedef enum { OPTIX_SUCCESS = 0 } OptixResult;
struct OptixAabb {
  float minX, minY, minZ, maxX, maxY, maxZ;
};
static inline const char* optixGetErrorName(OptixResult) { return "OPTIX_STUB"; }
#ifdef __CUDACC__
static __device__ __forceinline__ unsigned int optixGetPayload_0() { return 0; }
static __device__ __forceinline__ unsigned int optixGetPayload_1() { return 0; }
#endif
