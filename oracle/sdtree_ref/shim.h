// TEST INFRASTRUCTURE.  Type shim that lets lines 25-1008 of the reference's
// mitsuba/src/integrators/path/guided_path.cpp (BlobWriter ... STree) compile
// VERBATIM, outside Mitsuba, straight from /root/reference (never copied into
// this repo; see oracle/Makefile).  Only the handful of Mitsuba core types the
// SD-tree code touches are provided, with Mitsuba's single-precision semantics
// (include/mitsuba/core/{constants,point,vector,aabb,math}.h).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <stack>
#include <string>
#include <vector>
#include <array>
#include <atomic>
#include <chrono>
#include <fstream>
#include <functional>
#include <iomanip>
#include <sstream>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }

#ifdef M_PI
#undef M_PI
#endif
#define M_PI 3.14159265358979323846f   /* core/constants.h:63,80 (SINGLE_PRECISION) */
#define Epsilon 1e-4f                  /* core/constants.h:28 */

namespace mitsuba {
typedef float Float;

#define SAssert(cond) do { } while (0)          /* release build: SAssert compiles away */
enum ELogLevel { EWarn = 300 };
#define SLog(level, ...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)

struct Point2 {
    Float x, y;
    static const int dim = 2;
    Point2() : x(0), y(0) {}
    Point2(Float x_, Float y_) : x(x_), y(y_) {}
    explicit Point2(Float v) : x(v), y(v) {}
    Float &operator[](int i) { return (&x)[i]; }
    const Float &operator[](int i) const { return (&x)[i]; }
    Point2 operator+(const Point2 &o) const { return Point2(x + o.x, y + o.y); }
};
inline Point2 operator*(Float f, const Point2 &p) { return Point2(f * p.x, f * p.y); }

struct Vector {
    Float x, y, z;
    Vector() : x(0), y(0), z(0) {}
    Vector(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    explicit Vector(Float v) : x(v), y(v), z(v) {}
    Float &operator[](int i) { return (&x)[i]; }
    const Float &operator[](int i) const { return (&x)[i]; }
    Vector operator*(Float f) const { return Vector(x * f, y * f, z * f); }
};

struct Point {
    Float x, y, z;
    Point() : x(0), y(0), z(0) {}
    Point(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    explicit Point(const Vector &v) : x(v.x), y(v.y), z(v.z) {}
    Float &operator[](int i) { return (&x)[i]; }
    const Float &operator[](int i) const { return (&x)[i]; }
    Point operator+(const Vector &v) const { return Point(x + v.x, y + v.y, z + v.z); }
    Point operator-(const Vector &v) const { return Point(x - v.x, y - v.y, z - v.z); }
    Vector operator-(const Point &p) const { return Vector(x - p.x, y - p.y, z - p.z); }
};

struct AABB {
    Point min, max;
    Vector getExtents() const { return max - min; }
    Point clip(const Point &p) const {          /* core/aabb.h TAABB::clip */
        Point r;
        for (int i = 0; i < 3; ++i) r[i] = std::min(std::max(p[i], min[i]), max[i]);
        return r;
    }
};

struct Ray {                                     /* core/ray.h: origin + direction is all Vertex::commit (GP:1730-1768) touches */
    Point o; Vector d;
    Ray() {}
    Ray(const Point &o_, const Vector &d_, Float /* time */) : o(o_), d(d_) {}
};
struct Spectrum {                                /* TSpectrum<Float, 3> (core/spectrum.h): the members Vertex uses, with the reference's arithmetic (:467-486) */
    Float s[3];
    Spectrum() { s[0] = s[1] = s[2] = 0; }
    Spectrum(Float v) { s[0] = s[1] = s[2] = v; }
    Float &operator[](int i) { return s[i]; }
    const Float &operator[](int i) const { return s[i]; }
    Spectrum operator*(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * o.s[i]; return r; }
    Spectrum &operator+=(const Spectrum &o) { for (int i = 0; i < 3; ++i) s[i] += o.s[i]; return *this; }
    bool isValid() const { for (int i = 0; i < 3; ++i) if (!std::isfinite(s[i]) || s[i] < 0.0f) return false; return true; }
    Float average() const { Float result = 0.0f; for (int i = 0; i < 3; ++i) result += s[i]; return result * (1.0f / 3); }
};

class Sampler {
public:
    virtual ~Sampler() {}
    virtual Float next1D() = 0;
    virtual Point2 next2D() = 0;
};

namespace math {
inline Float clamp(Float v, Float lo, Float hi) { return std::min(hi, std::max(lo, v)); }   /* core/math.h */
inline void sincos(float theta, float *s, float *c) { ::sincosf(theta, s, c); }             /* core/math.h:219-221 */
}
}  // namespace mitsuba
