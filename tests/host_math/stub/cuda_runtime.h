#pragma once
// TEST ONLY stub (tests/host_math/harness.cpp): lets the device arithmetic headers of csrc/ compile as plain host C++.
// They use nothing of the CUDA runtime: only the function-space qualifiers and the overloaded math functions.
#include <cmath>
using std::sqrt;  // CUDA's headers overload sqrt for float like <cmath> does
#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
