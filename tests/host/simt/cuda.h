// Host stand-in for <cuda.h> (TEST INFRASTRUCTURE, see cuda_runtime.h in this directory): types only.
#pragma once
struct alignas(64) CUtensorMap { unsigned long long opaque[16]; };
