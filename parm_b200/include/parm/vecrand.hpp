// parm_b200 drop-in for ParM's src/vecrand.hpp (scalars, vectors, RNG) -- hot-path subset.
// Source compatible with the reference for the names the MD hot path and its drivers use
// (vecrand.hpp:1-62, 186-210; vecrand.cpp:1-39). Eigen and Boost are NOT required: `Vec` is a
// small fixed-size vector with the Eigen spellings ParM code uses (Zero(), dot, norm, ...).
#ifndef PARM_B200_VECRAND_H
#define PARM_B200_VECRAND_H
#ifdef VEC2D
#define NDIM 2
#define DIMROTATIONS 4
#else
#ifndef VEC3D
#define VEC3D
#endif
#endif
#ifdef VEC3D
#define NDIM 3
#define DIMROTATIONS 24
#endif
#ifdef LONGFLOAT
#error "parm_b200: the long double build (-DLONGFLOAT) has no GPU equivalent (DESIGN.md, out of scope)"
#endif
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <complex>
#include <ctime>
#include <iostream>
#include <memory>
#include <random>
#include <stdexcept>
#include <vector>

using namespace std;  // the reference headers do this (vecrand.hpp:26); drivers such as LJatoms.cpp rely on it
typedef unsigned int uint;
typedef double flt;
typedef std::complex<flt> cmplx;

namespace boost {  // ParM spells its smart pointers boost::shared_ptr (box.hpp:6-9)
using std::shared_ptr;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
}

namespace parm_b200 {
// Fixed-size R x C matrix of flt, column-major; R x 1 is the physics vector.
template <int R, int C>
class Mat {
    flt d[R * C];

   public:
    Mat() { for (int i = 0; i < R * C; i++) d[i] = 0; }
    Mat(flt x, flt y) { static_assert(R * C == 2, "2-vector"); d[0] = x; d[1] = y; }
    Mat(flt x, flt y, flt z) { static_assert(R * C == 3, "3-vector"); d[0] = x; d[1] = y; d[2] = z; }
    static Mat Zero() { return Mat(); }
    static Mat Identity() { Mat m; for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1; return m; }
    Mat &setZero() { for (int i = 0; i < R * C; i++) d[i] = 0; return *this; }
    flt *data() { return d; }
    const flt *data() const { return d; }
    int rows() const { return R; }
    int cols() const { return C; }
    int size() const { return R * C; }
    flt &operator()(int i) { return d[i]; }
    const flt &operator()(int i) const { return d[i]; }
    flt &operator[](int i) { return d[i]; }
    const flt &operator[](int i) const { return d[i]; }
    flt &operator()(int i, int j) { return d[i + j * R]; }
    const flt &operator()(int i, int j) const { return d[i + j * R]; }
    Mat operator+(const Mat &o) const { Mat r; for (int i = 0; i < R * C; i++) r.d[i] = d[i] + o.d[i]; return r; }
    Mat operator-(const Mat &o) const { Mat r; for (int i = 0; i < R * C; i++) r.d[i] = d[i] - o.d[i]; return r; }
    Mat operator-() const { Mat r; for (int i = 0; i < R * C; i++) r.d[i] = -d[i]; return r; }
    Mat operator*(flt s) const { Mat r; for (int i = 0; i < R * C; i++) r.d[i] = d[i] * s; return r; }
    Mat operator/(flt s) const { Mat r; for (int i = 0; i < R * C; i++) r.d[i] = d[i] / s; return r; }
    Mat &operator+=(const Mat &o) { for (int i = 0; i < R * C; i++) d[i] += o.d[i]; return *this; }
    Mat &operator-=(const Mat &o) { for (int i = 0; i < R * C; i++) d[i] -= o.d[i]; return *this; }
    Mat &operator*=(flt s) { for (int i = 0; i < R * C; i++) d[i] *= s; return *this; }
    Mat &operator/=(flt s) { for (int i = 0; i < R * C; i++) d[i] /= s; return *this; }
    bool operator==(const Mat &o) const { for (int i = 0; i < R * C; i++) if (!(d[i] == o.d[i])) return false; return true; }
    bool operator!=(const Mat &o) const { return !(*this == o); }
    // reductions associate like the reference's Eigen build: e0 + (e1 + e2)
    flt dot(const Mat &o) const {
        if (R * C == 3) return d[0] * o.d[0] + (d[1] * o.d[1] + d[2 % (R * C)] * o.d[2 % (R * C)]);
        flt s = 0;
        for (int i = 0; i < R * C; i++) s += d[i] * o.d[i];
        return s;
    }
    flt sum() const {
        if (R * C == 3) return d[0] + (d[1] + d[2 % (R * C)]);
        flt s = 0;
        for (int i = 0; i < R * C; i++) s += d[i];
        return s;
    }
    flt squaredNorm() const { return dot(*this); }
    flt norm() const { return std::sqrt(squaredNorm()); }
    Mat normalized() const { return (*this) / norm(); }
    void normalize() { *this /= norm(); }
    Mat cross(const Mat &o) const {
        static_assert(R * C == 3, "cross needs a 3-vector");
        return Mat(d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]);
    }
    Mat<C, R> transpose() const { Mat<C, R> r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r(j, i) = (*this)(i, j); return r; }
    template <int C2>
    Mat<R, C2> operator*(const Mat<C, C2> &o) const {
        Mat<R, C2> r;
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C2; j++) {
                flt s = 0;
                for (int k = 0; k < C; k++) s += (*this)(i, k) * o(k, j);
                r(i, j) = s;
            }
        return r;
    }
    flt trace() const { flt s = 0; for (int i = 0; i < (R < C ? R : C); i++) s += (*this)(i, i); return s; }
};
template <int R, int C>
inline Mat<R, C> operator*(flt s, const Mat<R, C> &m) { return m * s; }
template <int R, int C>
inline std::ostream &operator<<(std::ostream &os, const Mat<R, C> &m) {
    for (int i = 0; i < R; i++) {
        for (int j = 0; j < C; j++) os << (j ? " " : "") << m(i, j);
        if (i + 1 < R) os << "\n";
    }
    return os;
}
}  // namespace parm_b200

typedef parm_b200::Mat<NDIM, 1> Vec;
typedef parm_b200::Mat<2, 1> Vec2;
typedef parm_b200::Mat<3, 1> Vec3;
typedef parm_b200::Mat<NDIM, NDIM> Matrix;
typedef parm_b200::Mat<NDIM, 2> VecPair;

const flt OVERNDIM = ((flt)1.0) / NDIM;

#ifdef VEC3D
inline Vec vec() { return Vec(0, 0, 0); }
#else
inline Vec vec() { return Vec(0, 0); }
#endif
inline Vec2 vec(double x, double y) { return Vec2(x, y); }
inline Vec3 vec(double x, double y, double z) { return Vec3(x, y, z); }

// ---- global RNG (vecrand.cpp:3-39). boost::mt19937 == std::mt19937; the Gaussian stream of
// boost::normal_distribution is Boost-version dependent, so only the engine is pinned.
namespace parm_b200 {
inline std::mt19937 &randengine() {
    static std::mt19937 e;
    return e;
}
inline flt gauss01() {
    static std::normal_distribution<flt> d(0, 1);
    return d(randengine());
}
}  // namespace parm_b200
inline flt rand01() { return flt(parm_b200::randengine()()) / flt(4294967296.0); }
#ifdef VEC2D
inline Vec rand_vec() { flt a = parm_b200::gauss01(), b = parm_b200::gauss01(); return Vec(a, b); }
inline Vec rand_vec_boxed() { flt a = rand01(), b = rand01(); return Vec(a, b); }
#else
inline Vec rand_vec() { flt a = parm_b200::gauss01(), b = parm_b200::gauss01(), c = parm_b200::gauss01(); return Vec(a, b, c); }
inline Vec rand_vec_boxed() { flt a = rand01(), b = rand01(), c = rand01(); return Vec(a, b, c); }
#endif
inline unsigned int seed(unsigned int n) { parm_b200::randengine().seed(n); return n; }
inline unsigned int seed() { unsigned int n = static_cast<unsigned int>(time(0)); parm_b200::randengine().seed(n); return n; }

#endif
