// plbm_f32x2.cuh -- two fp32 nodes per instruction: Blackwell's packed fp32 operations (sm_100a).
//
// The fp32 collisions are bound by instruction issue, not by HBM (ncu, round 1, RR fp32 8192^2: issue slots 79 % busy,
// 86 % of the issued instructions are FADD / FMUL -- no FMA contraction is allowed, so ~250 of them per node and step).
// sm_100 executes an fp32 operation on TWO values held in an aligned register pair with one instruction (PTX
// {add,sub,mul,fma}.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2): half the issue slots for the same arithmetic.  F2 is that
// register pair with the arithmetic operators of a scalar, so the collision templates of plbm_math.cuh instantiate
// unchanged with T = F2 and process two vertically adjacent nodes at once; every lane is an individually rounded IEEE
// operation, so the result is BIT-IDENTICAL to the scalar instantiation.
//
// One trap: ptxas (12.9) contracts mul.rn.f32x2 followed by add.rn.f32x2 / sub.rn.f32x2 into one FFMA2 even under
// --fmad=false (it honours the flag for scalar code only; seen in the SASS of a two-line test).  A contracted multiply-add
// rounds once instead of twice and breaks bit parity, so the operators below are written as the FMAs that cannot be
// contracted any further, with the neutral operand read from constant memory so that ptxas cannot simplify them back:
//     a + b = fma(a, 1, b)      a - b = fma(b, -1, a)      a * b = fma(a, b, -0)
// each exact in the neutral operation (x * 1 and x * -1 are exact; p + (-0) = p for every p including both zeros), hence
// rounded once, exactly like the plain operation.
#pragma once
#include <cuda_runtime.h>

namespace plbm {

static __device__ __constant__ unsigned long long c_f2_one = 0x3f8000003f800000ull;      // ( 1.0f,  1.0f)
static __device__ __constant__ unsigned long long c_f2_negone = 0xbf800000bf800000ull;   // (-1.0f, -1.0f)
static __device__ __constant__ unsigned long long c_f2_negzero = 0x8000000080000000ull;  // (-0.0f, -0.0f)

struct F2 {
    unsigned long long r;  // lane 0 in the low half
    F2() = default;
    __host__ __device__ constexpr F2(float s) : r(bits(s) | ((unsigned long long)bits(s) << 32)) {}
    __host__ __device__ constexpr F2(double s) : F2((float)s) {}
    __host__ __device__ constexpr F2(int s) : F2((float)s) {}
    __host__ __device__ constexpr F2(float lo, float hi) : r(bits(lo) | ((unsigned long long)bits(hi) << 32)) {}
    static __host__ __device__ constexpr unsigned bits(float s) { return __builtin_bit_cast(unsigned, s); }
    __host__ __device__ constexpr float lo() const { return __builtin_bit_cast(float, (unsigned)r); }
    __host__ __device__ constexpr float hi() const { return __builtin_bit_cast(float, (unsigned)(r >> 32)); }
};

__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c)
{
    F2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.r) : "l"(a.r), "l"(b.r), "l"(c.r));
    return d;
}
__device__ __forceinline__ F2 f2_const(unsigned long long r)
{
    F2 x;
    x.r = r;
    return x;
}
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return f2_fma(a, f2_const(c_f2_one), b); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return f2_fma(b, f2_const(c_f2_negone), a); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return f2_fma(a, b, f2_const(c_f2_negzero)); }
// no packed division: two IEEE divisions (one per node and collision, as in the scalar code)
__host__ __device__ constexpr F2 operator/(F2 a, F2 b) { return F2(a.lo() / b.lo(), a.hi() / b.hi()); }
__host__ __device__ constexpr F2 operator-(F2 a)
{
    F2 x(0.f);
    x.r = a.r ^ 0x8000000080000000ull;
    return x;
}

// the scalar type a collision template computes its compile-time constants in
template <typename T> struct scalar_of {
    typedef T type;
};
template <> struct scalar_of<F2> {
    typedef float type;
};

}  // namespace plbm
