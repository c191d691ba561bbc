// TEST INFRASTRUCTURE — not product code.
//
// Minimal GLM stand-in used ONLY to compile the unmodified reference CUDA
// rasteriser (/root/reference/submodules/diff-surfel-rasterization{,_part}) into
// oracle/_ref/.  The reference depends on g-truc/glm as an un-vendored git
// submodule (third_party/glm, no pin recorded anywhere in the reference tree;
// upstream 2DGS pins 5c46b9c).  This header implements exactly the subset the
// reference touches: vec2/3/4, mat3/mat4/mat3x4/mat4x3 (column-major,
// m[col][row]), component-wise arithmetic, dot/length/max/sqrt/transpose and
// matrix products.  Every sum is left-to-right like GLM's scalar expansions,
// so nvcc's FMA contraction sees the same expression trees.
#pragma once
#include <cmath>
#include <type_traits>
#include <cuda_runtime.h>
#define GLM_HD __host__ __device__ inline
namespace glm {
template<int N> struct vec;
template<> struct vec<2>{ float x,y; GLM_HD vec(){} GLM_HD vec(float a):x(a),y(a){}
  template<class A,class B> GLM_HD vec(A a,B b):x((float)a),y((float)b){}
  GLM_HD float& operator[](int i){return (&x)[i];} GLM_HD const float& operator[](int i)const{return (&x)[i];} };
template<> struct vec<4>;
template<> struct vec<3>{ float x,y,z; GLM_HD vec(){} GLM_HD vec(float a):x(a),y(a),z(a){}
  template<class A,class B,class C> GLM_HD vec(A a,B b,C c):x((float)a),y((float)b),z((float)c){}
  GLM_HD explicit vec(const vec<4>& v);
  GLM_HD float& operator[](int i){return (&x)[i];} GLM_HD const float& operator[](int i)const{return (&x)[i];} };
template<> struct vec<4>{ float x,y,z,w; GLM_HD vec(){} GLM_HD vec(float a):x(a),y(a),z(a),w(a){}
  template<class A,class B,class C,class D> GLM_HD vec(A a,B b,C c,D d):x((float)a),y((float)b),z((float)c),w((float)d){}
  template<class D> GLM_HD vec(const vec<3>& v,D d):x(v.x),y(v.y),z(v.z),w((float)d){}
  GLM_HD float& operator[](int i){return (&x)[i];} GLM_HD const float& operator[](int i)const{return (&x)[i];} };
GLM_HD vec<3>::vec(const vec<4>& v):x(v.x),y(v.y),z(v.z){}
typedef vec<2> vec2; typedef vec<3> vec3; typedef vec<4> vec4;
#define GLM_VOP(op) \
 template<int N> GLM_HD vec<N> operator op(const vec<N>&a,const vec<N>&b){vec<N> r; for(int i=0;i<N;i++) r[i]=a[i] op b[i]; return r;} \
 template<int N> GLM_HD vec<N> operator op(const vec<N>&a,float b){vec<N> r; for(int i=0;i<N;i++) r[i]=a[i] op b; return r;} \
 template<int N> GLM_HD vec<N> operator op(float a,const vec<N>&b){vec<N> r; for(int i=0;i<N;i++) r[i]=a op b[i]; return r;} \
 template<int N> GLM_HD vec<N>& operator op##=(vec<N>&a,const vec<N>&b){for(int i=0;i<N;i++) a[i]=a[i] op b[i]; return a;} \
 template<int N> GLM_HD vec<N>& operator op##=(vec<N>&a,float b){for(int i=0;i<N;i++) a[i]=a[i] op b; return a;}
GLM_VOP(+) GLM_VOP(-) GLM_VOP(*) GLM_VOP(/)
template<int N> GLM_HD vec<N> operator-(const vec<N>&a){vec<N> r; for(int i=0;i<N;i++) r[i]=-a[i]; return r;}
template<int N> GLM_HD float dot(const vec<N>&a,const vec<N>&b){float s=a[0]*b[0]; for(int i=1;i<N;i++) s+=a[i]*b[i]; return s;}
template<int N> GLM_HD float length(const vec<N>&a){return sqrtf(dot(a,a));}
template<int N> GLM_HD vec<N> max(const vec<N>&a,const vec<N>&b){vec<N> r; for(int i=0;i<N;i++) r[i]=fmaxf(a[i],b[i]); return r;}
template<int N> GLM_HD vec<N> max(const vec<N>&a,float b){vec<N> r; for(int i=0;i<N;i++) r[i]=fmaxf(a[i],b); return r;}
template<int N> GLM_HD vec<N> sqrt(const vec<N>&a){vec<N> r; for(int i=0;i<N;i++) r[i]=sqrtf(a[i]); return r;}
template<int C,int R> struct mat{ vec<R> c[C]; GLM_HD mat(){}
  GLM_HD explicit mat(float d){for(int i=0;i<C;i++) for(int j=0;j<R;j++) c[i][j]=(i==j)?d:0.f;}
  GLM_HD mat(const vec<R>&a,const vec<R>&b,const vec<R>&d){static_assert(C==3,"");c[0]=a;c[1]=b;c[2]=d;}
  template<class...T, class=typename std::enable_if<sizeof...(T)==C*R && (C*R>3)>::type>
  GLM_HD mat(T...v){float t[]={(float)v...}; for(int i=0;i<C;i++) for(int j=0;j<R;j++) c[i][j]=t[i*R+j];}
  GLM_HD vec<R>& operator[](int i){return c[i];} GLM_HD const vec<R>& operator[](int i)const{return c[i];} };
typedef mat<3,3> mat3; typedef mat<4,4> mat4; typedef mat<3,4> mat3x4; typedef mat<4,3> mat4x3;
template<int C,int R> GLM_HD mat<R,C> transpose(const mat<C,R>&m){mat<R,C> r; for(int i=0;i<C;i++) for(int j=0;j<R;j++) r[j][i]=m[i][j]; return r;}
template<int C,int R> GLM_HD vec<R> operator*(const mat<C,R>&m,const vec<C>&v){vec<R> r=m[0]*v[0]; for(int i=1;i<C;i++) r+=m[i]*v[i]; return r;}
template<int K,int R,int C2> GLM_HD mat<C2,R> operator*(const mat<K,R>&a,const mat<C2,K>&b){mat<C2,R> r; for(int i=0;i<C2;i++) r[i]=a*b[i]; return r;}
}
