// Small fixed-size vector / matrix algebra for the per-surfel kernels.
//
// Column-major matrices (m[col][row]) and strictly left-to-right sums, so that
// the scalar expression trees nvcc sees (and therefore its FMA contraction) are
// the ones the reference's GLM expressions expand to.  Bit-exact radii / tile
// rectangles / sort keys depend on this (SURVEY.md §7 "Hard parts").
#pragma once
#include <cuda_runtime.h>

namespace pgs {

#define PGS_HD __host__ __device__ __forceinline__

template <int N> struct vec;
template <> struct vec<2> {
  float x, y;
  PGS_HD vec() {}
  PGS_HD vec(float a, float b) : x(a), y(b) {}
  PGS_HD float& operator[](int i) { return (&x)[i]; }
  PGS_HD const float& operator[](int i) const { return (&x)[i]; }
};
template <> struct vec<3> {
  float x, y, z;
  PGS_HD vec() {}
  PGS_HD vec(float a, float b, float c) : x(a), y(b), z(c) {}
  PGS_HD float& operator[](int i) { return (&x)[i]; }
  PGS_HD const float& operator[](int i) const { return (&x)[i]; }
};
template <> struct vec<4> {
  float x, y, z, w;
  PGS_HD vec() {}
  PGS_HD vec(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  PGS_HD vec(const vec<3>& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  PGS_HD float& operator[](int i) { return (&x)[i]; }
  PGS_HD const float& operator[](int i) const { return (&x)[i]; }
};
typedef vec<2> v2;
typedef vec<3> v3;
typedef vec<4> v4;

PGS_HD v3 xyz(const v4& v) { return v3(v.x, v.y, v.z); }

#define PGS_VOP(op)                                                                   \
  template <int N> PGS_HD vec<N> operator op(const vec<N>& a, const vec<N>& b) {      \
    vec<N> r;                                                                         \
    for (int i = 0; i < N; i++) r[i] = a[i] op b[i];                                  \
    return r;                                                                         \
  }                                                                                   \
  template <int N> PGS_HD vec<N> operator op(const vec<N>& a, float b) {              \
    vec<N> r;                                                                         \
    for (int i = 0; i < N; i++) r[i] = a[i] op b;                                     \
    return r;                                                                         \
  }                                                                                   \
  template <int N> PGS_HD vec<N> operator op(float a, const vec<N>& b) {              \
    vec<N> r;                                                                         \
    for (int i = 0; i < N; i++) r[i] = a op b[i];                                     \
    return r;                                                                         \
  }                                                                                   \
  template <int N> PGS_HD vec<N>& operator op##=(vec<N>& a, const vec<N>& b) {        \
    for (int i = 0; i < N; i++) a[i] = a[i] op b[i];                                  \
    return a;                                                                         \
  }                                                                                   \
  template <int N> PGS_HD vec<N>& operator op##=(vec<N>& a, float b) {                \
    for (int i = 0; i < N; i++) a[i] = a[i] op b;                                     \
    return a;                                                                         \
  }
PGS_VOP(+)
PGS_VOP(-)
PGS_VOP(*)
PGS_VOP(/)
#undef PGS_VOP

template <int N> PGS_HD vec<N> operator-(const vec<N>& a) {
  vec<N> r;
  for (int i = 0; i < N; i++) r[i] = -a[i];
  return r;
}
template <int N> PGS_HD float dot(const vec<N>& a, const vec<N>& b) {
  float s = a[0] * b[0];
  for (int i = 1; i < N; i++) s += a[i] * b[i];
  return s;
}
template <int N> PGS_HD float length(const vec<N>& a) { return sqrtf(dot(a, a)); }
template <int N> PGS_HD vec<N> vmax(const vec<N>& a, const vec<N>& b) {
  vec<N> r;
  for (int i = 0; i < N; i++) r[i] = fmaxf(a[i], b[i]);
  return r;
}
template <int N> PGS_HD vec<N> vmax(const vec<N>& a, float b) {
  vec<N> r;
  for (int i = 0; i < N; i++) r[i] = fmaxf(a[i], b);
  return r;
}
template <int N> PGS_HD vec<N> vsqrt(const vec<N>& a) {
  vec<N> r;
  for (int i = 0; i < N; i++) r[i] = sqrtf(a[i]);
  return r;
}

// C columns of R rows.
template <int C, int R> struct mat {
  vec<R> c[C];
  PGS_HD mat() {}
  PGS_HD vec<R>& operator[](int i) { return c[i]; }
  PGS_HD const vec<R>& operator[](int i) const { return c[i]; }
};
typedef mat<3, 3> m3;
typedef mat<4, 4> m4;
typedef mat<3, 4> m3x4;  // 3 columns, 4 rows
typedef mat<4, 3> m4x3;  // 4 columns, 3 rows

PGS_HD m3 make_m3(const v3& a, const v3& b, const v3& d) {
  m3 r;
  r[0] = a; r[1] = b; r[2] = d;
  return r;
}
PGS_HD m3x4 make_m3x4(const v4& a, const v4& b, const v4& d) {
  m3x4 r;
  r[0] = a; r[1] = b; r[2] = d;
  return r;
}
PGS_HD m3 diag3(float d) {
  m3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r[i][j] = (i == j) ? d : 0.f;
  return r;
}
template <int C, int R> PGS_HD mat<R, C> transpose(const mat<C, R>& m) {
  mat<R, C> r;
  for (int i = 0; i < C; i++)
    for (int j = 0; j < R; j++) r[j][i] = m[i][j];
  return r;
}
template <int C, int R> PGS_HD vec<R> operator*(const mat<C, R>& m, const vec<C>& v) {
  vec<R> r = m[0] * v[0];
  for (int i = 1; i < C; i++) r += m[i] * v[i];
  return r;
}
template <int K, int R, int C2> PGS_HD mat<C2, R> operator*(const mat<K, R>& a, const mat<C2, K>& b) {
  mat<C2, R> r;
  for (int i = 0; i < C2; i++) r[i] = a * b[i];
  return r;
}

}  // namespace pgs
