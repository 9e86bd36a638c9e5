// Facade header mirroring the public surface of ch4/v3/src/Vec3.h that client code uses.
// This header is only ever included by host C++ translation units (the aliases double3 / int3 below
// collide with CUDA's built-in vector types, which is why the device code lives behind the C ABI).
#ifndef VEC3_H
#define VEC3_H
#include <cmath>
#include <iomanip>
#include <ostream>
#include "all.h"

template <class T>
class Vec3 {
    T d[3];

public:
    Vec3() noexcept : d{0, 0, 0} {}
    Vec3(T x, T y, T z) noexcept : d{x, y, z} {}
    Vec3(const T o[3]) noexcept : d{o[0], o[1], o[2]} {}
    Vec3(T v) noexcept : d{v, v, v} {}

    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    void clear() { d[0] = d[1] = d[2] = 0; }
    type_calc length() const { return std::sqrt((*this) * (*this)); }
    void normalise() { (*this) /= length(); }
    Vec3 unit() const { return (*this) / length(); }
    Vec3 cross(const Vec3& o) const { return {d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]}; }
    Vec3 elWiseMult(const Vec3& o) const { return {d[0] * o.d[0], d[1] * o.d[1], d[2] * o.d[2]}; }
    T volume() const { return d[0] * d[1] * d[2]; }
    bool isNan() const { return std::isnan((double)d[0]) || std::isnan((double)d[1]) || std::isnan((double)d[2]); }

    Vec3& operator=(const T& v) noexcept { d[0] = d[1] = d[2] = v; return *this; }
    Vec3 operator+(T v) const { return {d[0] + v, d[1] + v, d[2] + v}; }
    Vec3 operator-(T v) const { return {d[0] - v, d[1] - v, d[2] - v}; }
    Vec3 operator*(T v) const { return {d[0] * v, d[1] * v, d[2] * v}; }
    Vec3 operator/(T v) const { T inv = 1.0 / v; return {d[0] * inv, d[1] * inv, d[2] * inv}; }     // multiply by the reciprocal, as the reference does
    void operator+=(T v) { d[0] += v; d[1] += v; d[2] += v; }
    void operator-=(T v) { d[0] -= v; d[1] -= v; d[2] -= v; }
    void operator*=(T v) { d[0] *= v; d[1] *= v; d[2] *= v; }
    void operator/=(T v) { T inv = 1.0 / v; d[0] *= inv; d[1] *= inv; d[2] *= inv; }

    Vec3 operator+(const Vec3& o) const { return {d[0] + o.d[0], d[1] + o.d[1], d[2] + o.d[2]}; }
    Vec3 operator-(const Vec3& o) const { return {d[0] - o.d[0], d[1] - o.d[1], d[2] - o.d[2]}; }
    T operator*(const Vec3& o) const { return d[0] * o.d[0] + d[1] * o.d[1] + d[2] * o.d[2]; }       // dot product
    Vec3 operator/(const Vec3& o) const { return {d[0] / o.d[0], d[1] / o.d[1], d[2] / o.d[2]}; }
    void operator+=(const Vec3& o) { d[0] += o.d[0]; d[1] += o.d[1]; d[2] += o.d[2]; }
    void operator-=(const Vec3& o) { d[0] -= o.d[0]; d[1] -= o.d[1]; d[2] -= o.d[2]; }
    bool operator==(const Vec3& o) const { return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2]; }
};

template <class T> Vec3<T> operator+(const T& v, const Vec3<T>& a) { return a + v; }
template <class T> Vec3<T> operator*(const T& v, const Vec3<T>& a) { return a * v; }
template <class T, class S> Vec3<T> operator*(const S& v, const Vec3<T>& a) { return a * T(v); }
template <class T> std::ostream& operator<<(std::ostream& out, const Vec3<T>& v) {
    return out << std::setw(5) << v[0] << " " << std::setw(5) << v[1] << " " << std::setw(5) << v[2];
}
template <class T> Vec3<T> cross(const Vec3<T>& l, const Vec3<T>& r) { return l.cross(r); }
template <class T> Vec3<T> abs(const Vec3<T> v) { return {std::fabs(v[0]), std::fabs(v[1]), std::fabs(v[2])}; }

using double3 = Vec3<double>;
using type_calc3 = Vec3<type_calc>;
using int3 = Vec3<int>;
#endif
