// TEST INFRASTRUCTURE — not product code.
//
// Minimal stand-in for the GLM 0.9.9.8 header (README.md:109 of the reference
// says "tested with GLM 0.9.9.8"; GLM is not vendored and not installed here).
// Only what the reference uses is provided: dvec3 / vec3, component-wise
// + - unary-, scalar*vec, vec*scalar, += -=, cross, dot, inversesqrt,
// normalize, and the converting constructor dvec3 <-> vec3
// (main/hmap.cpp:963 -> src/Orthographic.cpp:3).
//
// The arithmetic follows GLM 0.9.9.8's scalar (non-SIMD) definitions:
//   dot(a,b)       = (a.x*b.x + a.y*b.y) + a.z*b.z      (detail/func_geometric.inl, compute_dot<vec<3>>)
//   cross(x,y)     = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   inversesqrt(x) = 1 / sqrt(x)
//   normalize(v)   = v * inversesqrt(dot(v,v))
// GLM itself cannot be fetched offline, so these four lines are a restatement
// of its published source ("parity unpinned" at the GLM seam).
#ifndef HMRM_ORACLE_GLM_STANDIN_HPP
#define HMRM_ORACLE_GLM_STANDIN_HPP

#include <cmath>

namespace glm {

template <typename T>
struct tvec3 {
	T x, y, z;

	tvec3() {}
	tvec3(T a, T b, T c) : x(a), y(b), z(c) {}

	template <typename U>
	tvec3(const tvec3<U> &o)
		: x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}

	template <typename U>
	tvec3<T> &operator=(const tvec3<U> &o) {
		x = static_cast<T>(o.x);
		y = static_cast<T>(o.y);
		z = static_cast<T>(o.z);
		return *this;
	}

	template <typename U>
	tvec3<T> &operator+=(const tvec3<U> &o) {
		x += static_cast<T>(o.x);
		y += static_cast<T>(o.y);
		z += static_cast<T>(o.z);
		return *this;
	}

	template <typename U>
	tvec3<T> &operator-=(const tvec3<U> &o) {
		x -= static_cast<T>(o.x);
		y -= static_cast<T>(o.y);
		z -= static_cast<T>(o.z);
		return *this;
	}
};

template <typename T>
inline tvec3<T> operator+(const tvec3<T> &a, const tvec3<T> &b) {
	return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z);
}

template <typename T>
inline tvec3<T> operator-(const tvec3<T> &a, const tvec3<T> &b) {
	return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z);
}

template <typename T>
inline tvec3<T> operator-(const tvec3<T> &a) {
	return tvec3<T>(-a.x, -a.y, -a.z);
}

template <typename T>
inline tvec3<T> operator*(T s, const tvec3<T> &v) {
	return tvec3<T>(s * v.x, s * v.y, s * v.z);
}

template <typename T>
inline tvec3<T> operator*(const tvec3<T> &v, T s) {
	return tvec3<T>(v.x * s, v.y * s, v.z * s);
}

template <typename T>
inline T dot(const tvec3<T> &a, const tvec3<T> &b) {
	tvec3<T> tmp(a.x * b.x, a.y * b.y, a.z * b.z);
	return tmp.x + tmp.y + tmp.z;
}

template <typename T>
inline tvec3<T> cross(const tvec3<T> &x, const tvec3<T> &y) {
	return tvec3<T>(
		x.y * y.z - y.y * x.z,
		x.z * y.x - y.z * x.x,
		x.x * y.y - y.x * x.y);
}

template <typename T>
inline T inversesqrt(T x) {
	return static_cast<T>(1) / std::sqrt(x);
}

template <typename T>
inline tvec3<T> normalize(const tvec3<T> &v) {
	return v * inversesqrt(dot(v, v));
}

typedef tvec3<double> dvec3;
typedef tvec3<float> vec3;

} // namespace glm

#endif
