// TEST INFRASTRUCTURE -- the few Mitsuba core types / helpers the reference's MicrofacetDistribution (src/bsdfs/microfacet.h:45-721) and math::erf / erfinv /
// hypot2 (src/libcore/math.cpp:20-...) need, so that both compile VERBATIM (piped in by oracle/Makefile from /root/reference, never copied into this repository).
// Single precision build of the reference (Float = float), like sdtree_ref/shim.h.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <sstream>
#include <string>
#include <vector>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_EXPORT_CORE
#define FINLINE inline
#define SAssert(cond) do { } while (0)
#define SLog(level, ...) do { if ((level) >= mitsuba::EError) { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#undef M_PI
#define M_PI         3.14159265358979323846f      // include/mitsuba/core/constants.h (SINGLE_PRECISION)
#define INV_PI       0.31830988618379067154f
#define INV_TWOPI    0.15915494309189533577f
#define RCPOVERFLOW  2.93873587705571876e-39f
#define Epsilon      1e-4f

namespace mitsuba {
typedef float Float;
enum ELogLevel { ETrace = 0, EDebug = 100, EInfo = 200, EWarn = 300, EError = 400 };

struct Vector {
    Float x, y, z;
    Vector() : x(0), y(0), z(0) {}
    Vector(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    explicit Vector(Float v) : x(v), y(v), z(v) {}
    Vector operator*(Float f) const { return Vector(x * f, y * f, z * f); }
    Vector operator-(const Vector &v) const { return Vector(x - v.x, y - v.y, z - v.z); }
    Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); }
    Float length() const { return std::sqrt(x * x + y * y + z * z); }
    Float operator[](int i) const { return (&x)[i]; }
    Float &operator[](int i) { return (&x)[i]; }
};
struct Point {                                                                                  // TPoint3<Float>
    Float x, y, z;
    Point() : x(0), y(0), z(0) {}
    Point(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    Vector operator-(const Point &p) const { return Vector(x - p.x, y - p.y, z - p.z); }
    Float operator[](int i) const { return (&x)[i]; }
    operator Vector() const { return Vector(x, y, z); }                                         // `Vector(A)` in TriAccel::load
};
struct Ray { Point o; Vector d; };
struct Normal : public Vector {
    Normal() {}
    Normal(Float x_, Float y_, Float z_) : Vector(x_, y_, z_) {}
    Normal(const Vector &v) : Vector(v.x, v.y, v.z) {}
};
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float x_, Float y_) : x(x_), y(y_) {} };
struct Vector2 { Float x, y; Vector2() : x(0), y(0) {} Vector2(Float x_, Float y_) : x(x_), y(y_) {} explicit Vector2(Float v) : x(v), y(v) {} };
inline Float dot(const Vector &a, const Vector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float absDot(const Vector &a, const Vector &b) { return std::abs(dot(a, b)); }
inline Vector normalize(const Vector &v) { return v * (1.0f / v.length()); }                    // TVector3::operator/ multiplies by the reciprocal (core/vector.h)

struct Spectrum {                                                                               // TSpectrum<Float, 3> (include/mitsuba/core/spectrum.h): element-wise arithmetic
    Float s[3];
    Spectrum() { s[0] = s[1] = s[2] = 0; }
    explicit Spectrum(Float v) { s[0] = s[1] = s[2] = v; }
    Spectrum operator+(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] + o.s[i]; return r; }
    Spectrum operator-(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] - o.s[i]; return r; }
    Spectrum operator*(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * o.s[i]; return r; }
    Spectrum operator/(const Spectrum &o) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] / o.s[i]; return r; }
    Spectrum operator*(Float f) const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = s[i] * f; return r; }
    Spectrum safe_sqrt() const { Spectrum r; for (int i = 0; i < 3; ++i) r.s[i] = std::sqrt(std::max((Float) 0, s[i])); return r; }
};
inline Spectrum operator*(Float f, const Spectrum &v) { return v * f; }
inline Vector cross(const Vector &a, const Vector &b) { return Vector(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
#define EXPECT_NOT_TAKEN(a) (a)
namespace warp { Point2 squareToUniformDiskConcentric(const Point2 &sample); }                 // include/mitsuba/core/warp.h (used before its definition in warp.cpp)

struct Frame {                                                                                  // include/mitsuba/core/frame.h:60-130 (local shading frame helpers)
    static Float cosTheta(const Vector &v) { return v.z; }
    static Float cosTheta2(const Vector &v) { return v.z * v.z; }
    static Float sinTheta2(const Vector &v) { return 1.0f - v.z * v.z; }
    static Float sinTheta(const Vector &v) { Float t = sinTheta2(v); if (t <= 0.0f) return 0.0f; return std::sqrt(t); }
    static Float tanTheta(const Vector &v) { Float t = 1 - v.z * v.z; if (t <= 0.0f) return 0.0f; return std::sqrt(t) / v.z; }
};
namespace math {
    inline Float safe_sqrt(Float v) { return std::sqrt(std::max(0.0f, v)); }
    inline void sincos(Float theta, Float *s, Float *c) { ::sincosf(theta, s, c); }
    inline Float signum(Float v) { return copysignf(1.0f, v); }                                  // include/mitsuba/core/math.h:269-278 (SINGLE_PRECISION)
    inline float fastexp(float v) { return ::expf(v); }                                         // include/mitsuba/core/math.h:201-215 (non-MSVC branch)
    inline float fastlog(float v) { return ::logf(v); }
    extern Float erf(Float x); extern Float erfinv(Float x); extern float hypot2(float a, float b);
}
struct Properties {                                                                             // the Properties-based constructor compiles, the tests use the explicit one
    bool hasProperty(const std::string &) const { return false; }
    std::string getString(const std::string &) const { return ""; }
    Float getFloat(const std::string &) const { return 0; }
    bool getBoolean(const std::string &, bool d) const { return d; }
};
inline std::string formatString(const char *fmt, ...) { char buf[256]; va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap); return buf; }
}
namespace boost { inline std::string to_lower_copy(std::string s) { for (auto &c : s) c = (char) std::tolower(c); return s; } }
