// hlsl_shim.hpp -- just enough HLSL for g++ to compile the reference compute shader
// (Particles/nBodyGravityCS.hlsl) on the CPU.  TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// The shader source itself is never copied into this repository: oracle/Makefile lowers it with
// lower_hlsl.py (declaration syntax only -- every arithmetic line is compiled verbatim) and pipes the
// result, between this header and ref_harness.cpp, into g++; the output lands in oracle/_ref/.
// Semantics chosen where HLSL leaves latitude: every operation is a separately rounded IEEE binary32
// operation (-ffp-contract=off, no fast-math), dot() sums left to right, 1/sqrt is a correctly
// rounded sqrt followed by a correctly rounded divide.  A GPU driver may contract to mad and use
// rsq -- those variants are what the parity tolerance (1e-5) is for.
#ifndef MAPO_HLSL_SHIM_HPP
#define MAPO_HLSL_SHIM_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>

namespace hlsl {

typedef unsigned int uint;
struct float3;

// `.xyz` of a float3 / float4, as an lvalue sharing the vector's storage
struct swizzle_xyz {
    float x, y, z;
    swizzle_xyz &operator=(const float3 &v);
    swizzle_xyz &operator+=(const float3 &v);
    swizzle_xyz &operator*=(float s);
};

struct float3 {
    union {
        struct { float x, y, z; };
        swizzle_xyz xyz;
    };
    float3() : x(0.f), y(0.f), z(0.f) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(const swizzle_xyz &s) : x(s.x), y(s.y), z(s.z) {}
    float3 &operator+=(const float3 &v) { x += v.x; y += v.y; z += v.z; return *this; }
    float3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

inline swizzle_xyz &swizzle_xyz::operator=(const float3 &v) { x = v.x; y = v.y; z = v.z; return *this; }
inline swizzle_xyz &swizzle_xyz::operator+=(const float3 &v) { x += v.x; y += v.y; z += v.z; return *this; }
inline swizzle_xyz &swizzle_xyz::operator*=(float s) { x *= s; y *= s; z *= s; return *this; }

inline float3 operator-(const float3 &a, const float3 &b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator+(const float3 &a, const float3 &b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator*(const float3 &a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator-(const swizzle_xyz &a, const swizzle_xyz &b) { return float3(a) - float3(b); }
inline float3 operator*(const swizzle_xyz &a, float s) { return float3(a) * s; }

struct float4 {
    union {
        struct { float x, y, z, w; };
        swizzle_xyz xyz;
    };
    float4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(const float3 &v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float4(const swizzle_xyz &v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
};

struct uint3 { uint x, y, z; };
struct uint4 { uint x, y, z, w; };

inline float dot(const float3 &a, const float3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float sqrt(float v) { return std::sqrt(v); }
inline float length(const float3 &v) { return std::sqrt(dot(v, v)); }

// RWStructuredBuffer<T>: a view the harness points at host memory
template <class T> struct RWStructuredBuffer {
    T *data = nullptr;
    T &operator[](uint i) { return data[i]; }
};

}  // namespace hlsl

using namespace hlsl;

#endif  // MAPO_HLSL_SHIM_HPP
