// TEST INFRASTRUCTURE -- CPU oracle, never shipped, never on the product path.
//
// ppg_cpu_tracer.h: plain-C++ restatement of the reference's GuidedPathTracer
// integrator (mitsuba/src/integrators/path/guided_path.cpp, "GP") and of the Mitsuba
// services its loop calls: perspective camera, triangle meshes and analytic spheres,
// the BSDF models diffuse / dielectric / conductor / roughconductor / roughplastic /
// roughdielectric / plastic / thindielectric with the twosided, mask and bumpmap
// wrappers, bilinear bitmap textures, index-matched (null) transitions, area lights on
// meshes and spheres with light sampling, a lat-long environment emitter (sunsky is
// baked into one on the host, light sampling included), box-filtered film.  Pinned against the authors' render logs
// and images of CBOX, SPACESHIP and KITCHEN (tests/test_oracle_golden.py), and -- function by function, bit for bit --
// against the reference's own code compiled verbatim where that is possible: the microfacet distribution, erf / erfinv,
// the Fresnel functions, coordinateSystem and the cosine-hemisphere warp (oracle/microfacet_ref -> oracle/_ref/
// libmicrofacet_ref.so, tests/test_oracle_bsdf.py), the sky model (libskymodel_ref.so).  Templated on the SD-tree
// backend so that the same tracer runs either on the restated trees
// (sdtree_port.h) or on the reference's own SD-tree code compiled verbatim
// (oracle/sdtree_ref, built into oracle/_ref/).
//
// Deliberate, documented deviations from the reference (none changes the estimator):
//   * sampler: the reference uses one SFMT stream per worker thread and is not
//     reproducible (src/samplers/independent.cpp:41-59).  Here every path owns a
//     PCG32 stream keyed by (seed, global pass, pixel, sample) and consumes numbers
//     in the reference's order; the stochastic spatial filter draws its
//     3 numbers per committed vertex from a second per-vertex stream.  The CUDA
//     path uses the same generator, which is what makes path-level parity testable.
//   * acceleration structure: a BVH instead of the SAH kd-tree (hit set identical;
//     ties on t broken towards the lower triangle index).
//   * film: a sample lands in exactly the pixel containing it with weight 1 (the
//     reference's box filter has radius 0.5+1e-5: a 1e-5 sliver also touches the
//     neighbour; weights cancel in every quantity we reproduce).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <omp.h>

#include "../include/ppg.h"

namespace ppgo {

// ------------------------------------------------------------------ small vector math
struct F3 { float x, y, z; };
static inline F3 f3(float x, float y, float z) { return F3{x, y, z}; }
static inline F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline F3 operator*(F3 a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
static inline F3 operator*(F3 a, F3 b) { return F3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline F3 operator-(F3 a) { return F3{-a.x, -a.y, -a.z}; }
static inline float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline F3 cross(F3 a, F3 b) { return F3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline float length(F3 a) { return std::sqrt(dot(a, a)); }
static inline F3 normalize(F3 a) { return a * (1.0f / length(a)); }   // TVector3::operator/ multiplies by the reciprocal (core/vector.h)
static inline float comp(const F3 &a, int i) { return (&a.x)[i]; }
static inline bool is_zero(F3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }
static inline bool is_valid(F3 a) {   // Spectrum::isValid (core/spectrum.h): finite and non-negative
    return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z) && a.x >= 0 && a.y >= 0 && a.z >= 0;
}
static inline float max3(F3 a) { return std::max(std::max(a.x, a.y), a.z); }

static const float kPiT = 3.14159265358979323846f;    // M_PI (single-precision build)
static const float kEpsilon = 1e-4f;                 // core/constants.h:28
static const float kInvPi = 0.31830988618379067154f; // INV_PI
// coordinateSystem(a, b, c), src/libcore/util.cpp:592-601
static inline void coordinate_system(F3 a, F3 &b, F3 &c) {
    if (std::fabs(a.x) > std::fabs(a.y)) { const float invLen = 1.0f / std::sqrt(a.x * a.x + a.z * a.z); c = f3(a.z * invLen, 0.0f, -a.x * invLen); }
    else { const float invLen = 1.0f / std::sqrt(a.y * a.y + a.z * a.z); c = f3(0.0f, a.z * invLen, -a.y * invLen); }
    b = cross(c, a);
}

// ------------------------------------------------------------------ PCG32 path sampler
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct Pcg32 {
    uint64_t state, inc;
    void seed(uint64_t initstate, uint64_t initseq) {
        state = 0; inc = (initseq << 1) | 1u;
        nextU32(); state += initstate; nextU32();
    }
    uint32_t nextU32() {
        const uint64_t old = state;
        state = old * 6364136223846793005ull + inc;
        const uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t) (old >> 59u);
        return (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    float next1D() { return (float) (nextU32() >> 8) * (1.0f / 16777216.0f); }
};
// path stream: key = (seed, global sample index)
static inline void seed_path_rng(Pcg32 &r, uint64_t seed, uint64_t sampleIndex) {
    r.seed(splitmix64(seed ^ splitmix64(sampleIndex)), sampleIndex);
}
// per-committed-vertex stream for the stochastic spatial filter
static inline void seed_vertex_rng(Pcg32 &r, uint64_t seed, uint64_t sampleIndex, uint32_t vertexOrdinal) {
    r.seed(splitmix64((seed + 0x5851F42D4C957F2Dull) ^ splitmix64(sampleIndex * 64 + vertexOrdinal)), sampleIndex * 64 + vertexOrdinal);
}

// ------------------------------------------------------------------ scene
struct TriAccelP {   // Wald's projection test, restated from include/mitsuba/render/triaccel.h:60-158
    int k; float n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv;
};
static inline bool triaccel_load(TriAccelP &t, F3 A, F3 B, F3 C) {
    static const int mod3[4] = {1, 2, 0, 1};
    const F3 b = C - A, c = B - A, N = cross(c, b);
    int k = 0;
    for (int j = 0; j < 3; ++j) if (std::fabs(comp(N, j)) > std::fabs(comp(N, k))) k = j;
    const int u = mod3[k], v = mod3[k + 1];
    const float n_k = comp(N, k), denom = comp(b, u) * comp(c, v) - comp(b, v) * comp(c, u);
    if (denom == 0) { t.k = 3; return false; }
    t.k = k;
    t.n_u = comp(N, u) / n_k; t.n_v = comp(N, v) / n_k; t.n_d = dot(A, N) / n_k;
    t.b_nu = comp(b, u) / denom; t.b_nv = -comp(b, v) / denom;
    t.a_u = comp(A, u); t.a_v = comp(A, v);
    t.c_nu = comp(c, v) / denom; t.c_nv = -comp(c, u) / denom;
    return true;
}
static inline bool triaccel_intersect(const TriAccelP &tr, F3 o, F3 d, float mint, float maxt, float &u, float &v, float &t) {
    float o_u, o_v, o_k, d_u, d_v, d_k;
    switch (tr.k) {
        case 0: o_u = o.y; o_v = o.z; o_k = o.x; d_u = d.y; d_v = d.z; d_k = d.x; break;
        case 1: o_u = o.z; o_v = o.x; o_k = o.y; d_u = d.z; d_v = d.x; d_k = d.y; break;
        case 2: o_u = o.x; o_v = o.y; o_k = o.z; d_u = d.x; d_v = d.y; d_k = d.z; break;
        default: return false;
    }
    t = (tr.n_d - o_u * tr.n_u - o_v * tr.n_v - o_k) / (d_u * tr.n_u + d_v * tr.n_v + d_k);
    if (!(t >= mint && t <= maxt)) return false;   // NaN-safe form of "t < mint || t > maxt -> miss"
    const float hu = o_u + t * d_u - tr.a_u;
    const float hv = o_v + t * d_v - tr.a_v;
    u = hv * tr.b_nu + hu * tr.b_nv;
    v = hu * tr.c_nu + hv * tr.c_nv;
    return u >= 0 && v >= 0 && u + v <= 1.0f;
}

static const float kInvPiS = 0.31830988618379067154f; // INV_PI
struct BvhNode { float bmin[3], bmax[3]; uint32_t left, count; };   // count>0: leaf, left = first prim slot; else children left, left+1

struct Scene {
    std::vector<F3> P, N; std::vector<float> UV;
    std::vector<uint32_t> idx, triShape;
    std::vector<ppg_shape> shapes; std::vector<ppg_bsdf> bsdfs; std::vector<F3> radiance; std::vector<float> tables;
    std::vector<ppg_sphere> spheres;
    ppg_camera cam; F3 aabbMin, aabbMax;
    std::vector<TriAccelP> accel; std::vector<BvhNode> bvh; std::vector<uint32_t> primOrder;
    // camera derived (src/sensors/perspective.cpp:120-298)
    F3 camO, camLeft, camUp, camDir; float tanX, tanY;
    // emitter sampling (next event estimation): per emitter the area distribution over its triangles
    // (TriMesh::prepareSamplingTable, src/librender/trimesh.cpp:388-403) and the discrete emitter choice (scene.cpp:357-381)
    struct EmitterSampler { uint32_t firstTri, nTris; std::vector<float> cdf; float invArea; int shape; int sphere; };
    std::vector<EmitterSampler> emitterSamplers; std::vector<float> emitterCdf; float emitterNormalization = 0;
    // bitmap textures (level 0, half precision like the reference's storage) and the environment map
    std::vector<ppg_texture> textures; std::vector<uint16_t> texels;
    uint32_t envW = 0, envH = 0; std::vector<uint16_t> envTexels; float envScale = 1.0f; float worldToEnv[9];

    static float halfToFloat(uint16_t h) {
        const uint32_t sgn = (uint32_t) (h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
        uint32_t bits;
        if (e == 0) {
            if (m == 0) bits = sgn;
            else { int sh = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; ++sh; } bits = sgn | ((uint32_t) (113 - sh) << 23) | ((mm & 1023u) << 13); }
        } else if (e == 31) bits = sgn | 0x7f800000u | (m << 13);
        else bits = sgn | ((e + 112u) << 23) | (m << 13);
        float f; std::memcpy(&f, &bits, 4); return f;
    }
    static int wrapIndex(int x, int size, uint32_t mode) {            // TMIPMap::evalTexel boundary handling, render/mipmap.h:503-563
        if (x >= 0 && x < size) return x;
        if (mode == PPG_WRAP_REPEAT) { int r = x % size; return r < 0 ? r + size : r; }          // math::modulo
        if (mode == PPG_WRAP_CLAMP) return std::min(std::max(x, 0), size - 1);
        int r = x % (2 * size); if (r < 0) r += 2 * size;                                        // mirror
        return r >= size ? 2 * size - r - 1 : r;
    }
    static F3 texel(const uint16_t *base, uint32_t W, uint32_t Hh, uint32_t channels, uint32_t wu, uint32_t wv, int x, int y) {
        x = wrapIndex(x, (int) W, wu); y = wrapIndex(y, (int) Hh, wv);
        const uint16_t *px = base + ((size_t) y * W + (size_t) x) * channels;
        if (channels == 1) { const float v = halfToFloat(px[0]); return f3(v, v, v); }
        return f3(halfToFloat(px[0]), halfToFloat(px[1]), halfToFloat(px[2]));
    }
    // TMIPMap::evalBilinear(0, uv), render/mipmap.h:575-596
    static F3 bilinear(const uint16_t *base, uint32_t W, uint32_t Hh, uint32_t channels, uint32_t wu, uint32_t wv, float uu, float vv) {
        if (!std::isfinite(uu) || !std::isfinite(vv)) return f3(0, 0, 0);
        const float u = uu * (float) W - 0.5f, v = vv * (float) Hh - 0.5f;
        const int xPos = (int) std::floor(u), yPos = (int) std::floor(v);
        const float dx1 = u - (float) xPos, dx2 = 1.0f - dx1, dy1 = v - (float) yPos, dy2 = 1.0f - dy1;
        return texel(base, W, Hh, channels, wu, wv, xPos, yPos) * dx2 * dy2 + texel(base, W, Hh, channels, wu, wv, xPos, yPos + 1) * dx2 * dy1
             + texel(base, W, Hh, channels, wu, wv, xPos + 1, yPos) * dx1 * dy2 + texel(base, W, Hh, channels, wu, wv, xPos + 1, yPos + 1) * dx1 * dy1;
    }
    // Texture2D::eval(its) without UV partials (librender/texture.cpp:112-121) -> BitmapTexture::eval(uv) (textures/bitmap.cpp:431-453)
    F3 evalTexture(uint32_t idx, float u, float v) const {
        const ppg_texture &t = textures[idx];
        return bilinear(texels.data() + t.first_texel, t.width, t.height, t.channels, t.wrap_u, t.wrap_v, u * t.uv_scale[0] + t.uv_offset[0], v * t.uv_scale[1] + t.uv_offset[1]);
    }
    // Texture2D::evalGradient(its) (texture.cpp:123-130) -> BitmapTexture::evalGradient(uv) (bitmap.cpp:455-479) -> evalGradientBilinear (mipmap.h:601-626);
    // returns the luminances BumpMap::getFrame uses (bumpmap.cpp:141-144)
    void evalTextureGradientLum(uint32_t idx, float u_, float v_, float &dDu, float &dDv) const {
        const ppg_texture &t = textures[idx];
        const float uu = u_ * t.uv_scale[0] + t.uv_offset[0], vv = v_ * t.uv_scale[1] + t.uv_offset[1];
        F3 g0 = f3(0, 0, 0), g1 = f3(0, 0, 0);
        if (std::isfinite(uu) && std::isfinite(vv)) {
            const uint16_t *base = texels.data() + t.first_texel;
            const float u = uu * (float) t.width - 0.5f, v = vv * (float) t.height - 0.5f;
            const int xPos = (int) std::floor(u), yPos = (int) std::floor(v);
            const float dx = u - (float) xPos, dy = v - (float) yPos;
            const F3 p00 = texel(base, t.width, t.height, t.channels, t.wrap_u, t.wrap_v, xPos, yPos), p10 = texel(base, t.width, t.height, t.channels, t.wrap_u, t.wrap_v, xPos + 1, yPos),
                     p01 = texel(base, t.width, t.height, t.channels, t.wrap_u, t.wrap_v, xPos, yPos + 1), p11 = texel(base, t.width, t.height, t.channels, t.wrap_u, t.wrap_v, xPos + 1, yPos + 1);
            const F3 tmp = p01 + p10 - p11;
            g0 = (p10 + p00 * (dy - 1) - tmp * dy) * (float) t.width;
            g1 = (p01 + p00 * (dx - 1) - tmp * dx) * (float) t.height;
        }
        g0 = g0 * t.uv_scale[0]; g1 = g1 * t.uv_scale[1];
        dDu = g0.x * 0.212671f + g0.y * 0.715160f + g0.z * 0.072169f;
        dDv = g1.x * 0.212671f + g1.y * 0.715160f + g1.z * 0.072169f;
    }
    bool hasEnvironment() const { return envW != 0; }
    // ---- light sampling of the environment emitter (nee != never).  EnvironmentMap::configure (src/emitters/envmap.cpp:260-329): marginal row / conditional
    // column CDFs over texel luminance x sin(theta), in the reference's float / double mix; createShape (:330-335): the scene's bounding sphere x 1.5
    std::vector<float> envCdfRows, envCdfCols, envRowWeights; float envNormalization = 0, envPixelX = 0, envPixelY = 0, envRadius = 0; F3 envCenter; float envToWorld[9];
    static float luminance(F3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }   // Color3::getLuminance, spectrum.h:836-838
    F3 envTexel(int x, int y) const { return texel(envTexels.data(), envW, envH, 3, PPG_WRAP_REPEAT, PPG_WRAP_CLAMP, x, y); }   // evalTexel(0, x, y): u repeats, v clamps
    void buildEnvSampler() {
        const uint32_t Wd = envW, Hd = envH;
        envCdfCols.assign((size_t) (Wd + 1) * Hd, 0.f); envCdfRows.assign((size_t) Hd + 1, 0.f); envRowWeights.assign(Hd, 0.f);
        size_t colPos = 0, rowPos = 0; float rowSum = 0.0f;
        envCdfRows[rowPos++] = 0;
        for (uint32_t y = 0; y < Hd; ++y) {
            float colSum = 0;
            envCdfCols[colPos++] = 0;
            for (uint32_t x = 0; x < Wd; ++x) { colSum += luminance(envTexel((int) x, (int) y)); envCdfCols[colPos++] = colSum; }
            const float normalization = 1.0f / colSum;
            for (uint32_t x = 1; x < Wd; ++x) envCdfCols[colPos - x - 1] *= normalization;
            envCdfCols[colPos - 1] = 1.0f;
            const float weight = (float) std::sin((double) ((float) y + 0.5f) * 3.14159265358979323846 / (double) Hd);
            envRowWeights[y] = weight;
            rowSum += colSum * weight;
            envCdfRows[rowPos++] = rowSum;
        }
        const float normalization = 1.0f / rowSum;
        for (uint32_t y = 1; y < Hd; ++y) envCdfRows[rowPos - y - 1] *= normalization;
        envCdfRows[rowPos - 1] = 1.0f;
        envNormalization = (float) (1.0 / ((double) rowSum * (2 * 3.14159265358979323846 / (double) Wd) * (3.14159265358979323846 / (double) Hd)));
        envPixelX = (float) (2 * 3.14159265358979323846 / (double) Wd); envPixelY = (float) (3.14159265358979323846 / (double) Hd);
        envCenter = (aabbMax + aabbMin) * 0.5f;                                                // AABB::getBSphere, libcore/aabb.cpp:44-47
        envRadius = std::max(kEpsilon, length(envCenter - aabbMax) * 1.5f);
        const float *m = worldToEnv;                                                           // the emitter-to-world rotation back from its inverse
        const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
        const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g), id = 1.0 / det;
        const double inv[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id,
                               (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
        for (int k = 0; k < 9; ++k) envToWorld[k] = (float) inv[k];
    }
    static uint32_t envSampleReuse(const float *cdf, uint32_t size, float &sample) {           // EnvironmentMap::sampleReuse, envmap.cpp:657-662
        const float *entry = std::lower_bound(cdf, cdf + size + 1, sample);
        const uint32_t index = std::min((uint32_t) std::max((ptrdiff_t) 0, entry - cdf - 1), size - 1);
        sample = (sample - cdf[index]) / (cdf[index + 1] - cdf[index]);
        return index;
    }
    static float intervalToTent(float sample) {                                                // libcore/warp.cpp:143-155
        float sign;
        if (sample < 0.5f) { sign = 1; sample *= 2; } else { sign = -1; sample = 2 * (sample - 0.5f); }
        return sign * (1 - std::sqrt(sample));
    }
    // bilinear luminance-weighted density shared by internalSampleDirection / internalPdfDirection (envmap.cpp:577-591, 619-632), before the 1/sin(theta)
    float envDensity(float px, float py, F3 *valueOut) const {
        const int xPos = (int) std::floor(px), yPos = (int) std::floor(py);
        const float dx1 = px - (float) xPos, dx2 = 1.0f - dx1, dy1 = py - (float) yPos, dy2 = 1.0f - dy1;
        const F3 value1 = envTexel(xPos, yPos) * dx2 * dy2 + envTexel(xPos + 1, yPos) * dx1 * dy2;
        const F3 value2 = envTexel(xPos, yPos + 1) * dx2 * dy1 + envTexel(xPos + 1, yPos + 1) * dx1 * dy1;
        if (valueOut) *valueOut = (value1 + value2) * envScale;
        const int H1 = (int) envH - 1;
        return (luminance(value1) * envRowWeights[std::min(std::max(yPos, 0), H1)] + luminance(value2) * envRowWeights[std::min(std::max(yPos + 1, 0), H1)]) * envNormalization;
    }
    // EnvironmentMap::internalSampleDirection, envmap.cpp:567-600 (direction in the emitter's frame)
    void envSampleDirection(float sx, float sy, F3 &d, F3 &value, float &pdf) const {
        const uint32_t row = envSampleReuse(envCdfRows.data(), envH, sy);
        const uint32_t col = envSampleReuse(envCdfCols.data() + (size_t) row * (envW + 1), envW, sx);
        const float px = (float) col + intervalToTent(sx), py = (float) row + intervalToTent(sy);
        pdf = envDensity(px, py, &value);
        float sinPhi, cosPhi, sinTheta, cosTheta;
        sincosf(envPixelX * (px + 0.5f), &sinPhi, &cosPhi);
        sincosf(envPixelY * (py + 0.5f), &sinTheta, &cosTheta);
        d = f3(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta);
        pdf /= std::max(std::fabs(sinTheta), kEpsilon);
    }
    // EnvironmentMap::pdfDirect (solid angle measure) -> internalPdfDirection, envmap.cpp:545-548, 603-633, for a WORLD direction
    float envPdfDirection(F3 dw) const {
        const F3 d = f3(worldToEnv[0] * dw.x + worldToEnv[1] * dw.y + worldToEnv[2] * dw.z, worldToEnv[3] * dw.x + worldToEnv[4] * dw.y + worldToEnv[5] * dw.z,
                        worldToEnv[6] * dw.x + worldToEnv[7] * dw.y + worldToEnv[8] * dw.z);
        const float uu = std::atan2(d.x, -d.z) * 0.15915494309189533577f, vv = std::acos(std::min(1.0f, std::max(-1.0f, d.y))) * kInvPiS;
        if (!std::isfinite(uu) || !std::isfinite(vv)) return 0.0f;
        const float u = uu * (float) envW - 0.5f, v = vv * (float) envH - 0.5f;
        const float sinTheta = std::sqrt(std::max(0.0f, 1 - d.y * d.y));                       // math::safe_sqrt
        return envDensity(u, v, nullptr) / std::max(std::fabs(sinTheta), kEpsilon);
    }
    // EnvironmentMap::evalEnvironment without ray differentials (src/emitters/envmap.cpp:380-410): u repeats, v clamps (:176-178)
    F3 evalEnvironment(F3 d) const {
        const F3 v = f3(worldToEnv[0] * d.x + worldToEnv[1] * d.y + worldToEnv[2] * d.z, worldToEnv[3] * d.x + worldToEnv[4] * d.y + worldToEnv[5] * d.z,
                        worldToEnv[6] * d.x + worldToEnv[7] * d.y + worldToEnv[8] * d.z);
        const float uu = std::atan2(v.x, -v.z) * 0.15915494309189533577f;                  // INV_TWOPI
        const float vv = std::acos(std::min(1.0f, std::max(-1.0f, v.y))) * kInvPiS;          // math::safe_acos * INV_PI
        return bilinear(envTexels.data(), envW, envH, 3, PPG_WRAP_REPEAT, PPG_WRAP_CLAMP, uu, vv) * envScale;
    }

    // DiscreteDistribution::sample (include/mitsuba/core/pmf.h:124-137)
    static size_t cdfSample(const std::vector<float> &cdf, float v) {
        const ptrdiff_t entry = std::lower_bound(cdf.begin(), cdf.end(), v) - cdf.begin();
        size_t index = std::min(cdf.size() - 2, (size_t) std::max((ptrdiff_t) 0, entry - 1));
        while (cdf[index + 1] - cdf[index] == 0 && index < cdf.size() - 1) ++index;
        return index;
    }
    void buildEmitterSamplers() {
        emitterSamplers.clear();
        for (size_t e = 0; e < radiance.size(); ++e) {
            EmitterSampler es; es.shape = -1; es.firstTri = es.nTris = 0; es.invArea = 0; es.sphere = -1;
            for (size_t sidx = 0; sidx < shapes.size(); ++sidx) if (shapes[sidx].emitter == (int) e) { es.shape = (int) sidx; es.firstTri = shapes[sidx].first_triangle; es.nTris = shapes[sidx].n_triangles; }
            es.cdf.assign(1, 0.0f);
            for (uint32_t t = es.firstTri; t < es.firstTri + es.nTris; ++t) {
                const F3 p0 = P[idx[3 * t]], p1 = P[idx[3 * t + 1]], p2 = P[idx[3 * t + 2]];
                es.cdf.push_back(es.cdf.back() + 0.5f * length(cross(p1 - p0, p2 - p0)));        // Triangle::surfaceArea (libcore/triangle.cpp:61-67)
            }
            const float sum = es.cdf.back();                                                     // DiscreteDistribution::normalize (pmf.h:101-114)
            if (sum > 0) { const float nrm = 1.0f / sum; for (size_t i = 1; i < es.cdf.size(); ++i) es.cdf[i] *= nrm; es.cdf.back() = 1.0f; es.invArea = 1.0f / sum; }
            for (size_t k = 0; k < spheres.size(); ++k) if (spheres[k].shape == es.shape) {          // sphere.cpp:128: m_invSurfaceArea
                es.sphere = (int) k; es.invArea = 1 / (4 * kPiT * spheres[k].radius * spheres[k].radius);
            }
            emitterSamplers.push_back(es);
        }
        emitterCdf.assign(1, 0.0f);
        for (size_t e = 0; e < emitterSamplers.size(); ++e) emitterCdf.push_back(emitterCdf.back() + 1.0f);     // getSamplingWeight() == 1
        if (hasEnvironment()) { buildEnvSampler(); emitterCdf.push_back(emitterCdf.back() + 1.0f); }           // the environment emitter is the last entry of m_emitters here
        if (emitterCdf.back() > 0) { emitterNormalization = 1.0f / emitterCdf.back(); for (size_t i = 1; i < emitterCdf.size(); ++i) emitterCdf[i] *= emitterNormalization; emitterCdf.back() = 1.0f; }
    }

    void load(const ppg_scene_desc &d) {
        P.resize(d.n_vertices); N.resize(d.n_vertices); UV.assign(2 * (size_t) d.n_vertices, 0.f);
        for (uint32_t i = 0; i < d.n_vertices; ++i) {
            P[i] = f3(d.positions[3 * i], d.positions[3 * i + 1], d.positions[3 * i + 2]);
            N[i] = d.normals ? f3(d.normals[3 * i], d.normals[3 * i + 1], d.normals[3 * i + 2]) : f3(0, 0, 0);
            if (d.uvs) { UV[2 * i] = d.uvs[2 * i]; UV[2 * i + 1] = d.uvs[2 * i + 1]; }
        }
        idx.assign(d.indices, d.indices + 3 * (size_t) d.n_triangles);
        triShape.assign(d.triangle_shape, d.triangle_shape + d.n_triangles);
        shapes.assign(d.shapes, d.shapes + d.n_shapes);
        bsdfs.assign(d.bsdfs, d.bsdfs + d.n_bsdfs);
        tables.clear();
        if (d.bsdf_tables && d.n_bsdf_tables) tables.assign(d.bsdf_tables, d.bsdf_tables + (size_t)d.n_bsdf_tables * PPG_BSDF_TABLE_SIZE);
        spheres.clear();
        if (d.spheres && d.n_spheres) spheres.assign(d.spheres, d.spheres + d.n_spheres);
        radiance.resize(d.n_emitters);
        for (uint32_t i = 0; i < d.n_emitters; ++i) radiance[i] = f3(d.area_radiance[3 * i], d.area_radiance[3 * i + 1], d.area_radiance[3 * i + 2]);
        textures.clear(); texels.clear();
        if (d.textures && d.n_textures) { textures.assign(d.textures, d.textures + d.n_textures); texels.assign(d.texels, d.texels + d.n_texels); }
        envW = d.envmap.width; envH = d.envmap.height; envTexels.clear();
        if (envW && envH && d.envmap.texels) { envTexels.assign(d.envmap.texels, d.envmap.texels + (size_t) envW * envH * 3); envScale = d.envmap.scale; std::memcpy(worldToEnv, d.envmap.world_to_env, sizeof(worldToEnv)); }
        else envW = envH = 0;
        cam = d.camera;
        aabbMin = f3(d.aabb_min[0], d.aabb_min[1], d.aabb_min[2]); aabbMax = f3(d.aabb_max[0], d.aabb_max[1], d.aabb_max[2]);
        const float *m = cam.to_world;
        camLeft = f3(m[0], m[4], m[8]); camUp = f3(m[1], m[5], m[9]); camDir = f3(m[2], m[6], m[10]); camO = f3(m[3], m[7], m[11]);
        const float aspect = (float) cam.film_width / (float) cam.film_height;
        tanX = std::tan(0.5f * cam.x_fov_deg * (kPiT / 180.0f));
        tanY = tanX / aspect;
        accel.resize(d.n_triangles);
        for (uint32_t t = 0; t < d.n_triangles; ++t) triaccel_load(accel[t], P[idx[3 * t]], P[idx[3 * t + 1]], P[idx[3 * t + 2]]);
        buildBvh();
        buildEmitterSamplers();
    }

    // median-split BVH over triangle centroids (quality irrelevant for parity; hit set is what matters)
    void buildBvh() {
        const uint32_t nt = (uint32_t) triShape.size();
        primOrder.resize(nt);
        for (uint32_t i = 0; i < nt; ++i) primOrder[i] = i;
        std::vector<F3> cen(nt), tmin(nt), tmax(nt);
        for (uint32_t t = 0; t < nt; ++t) {
            F3 a = P[idx[3 * t]], b = P[idx[3 * t + 1]], c = P[idx[3 * t + 2]];
            tmin[t] = f3(std::min(a.x, std::min(b.x, c.x)), std::min(a.y, std::min(b.y, c.y)), std::min(a.z, std::min(b.z, c.z)));
            tmax[t] = f3(std::max(a.x, std::max(b.x, c.x)), std::max(a.y, std::max(b.y, c.y)), std::max(a.z, std::max(b.z, c.z)));
            cen[t] = (tmin[t] + tmax[t]) * 0.5f;
        }
        bvh.clear(); bvh.reserve(2 * nt + 1); bvh.emplace_back();
        struct Job { uint32_t node, first, count; };
        std::vector<Job> jobs; jobs.push_back(Job{0, 0, nt});
        while (!jobs.empty()) {
            Job j = jobs.back(); jobs.pop_back();
            F3 mn = f3(1e30f, 1e30f, 1e30f), mx = f3(-1e30f, -1e30f, -1e30f), cmn = mn, cmx = mx;
            for (uint32_t i = j.first; i < j.first + j.count; ++i) {
                const uint32_t t = primOrder[i];
                mn = f3(std::min(mn.x, tmin[t].x), std::min(mn.y, tmin[t].y), std::min(mn.z, tmin[t].z));
                mx = f3(std::max(mx.x, tmax[t].x), std::max(mx.y, tmax[t].y), std::max(mx.z, tmax[t].z));
                cmn = f3(std::min(cmn.x, cen[t].x), std::min(cmn.y, cen[t].y), std::min(cmn.z, cen[t].z));
                cmx = f3(std::max(cmx.x, cen[t].x), std::max(cmx.y, cen[t].y), std::max(cmx.z, cen[t].z));
            }
            BvhNode nd;
            nd.bmin[0] = mn.x; nd.bmin[1] = mn.y; nd.bmin[2] = mn.z; nd.bmax[0] = mx.x; nd.bmax[1] = mx.y; nd.bmax[2] = mx.z;
            const F3 ext = cmx - cmn;
            int ax = 0; if (ext.y > comp(ext, ax)) ax = 1; if (ext.z > comp(ext, ax)) ax = 2;
            if (j.count <= 2 || comp(ext, ax) <= 0) { nd.left = j.first; nd.count = j.count; bvh[j.node] = nd; continue; }
            const uint32_t mid = j.first + j.count / 2;
            std::nth_element(primOrder.begin() + j.first, primOrder.begin() + mid, primOrder.begin() + j.first + j.count,
                             [&](uint32_t a, uint32_t b) { return comp(cen[a], ax) < comp(cen[b], ax) || (comp(cen[a], ax) == comp(cen[b], ax) && a < b); });
            nd.left = (uint32_t) bvh.size(); nd.count = 0; bvh[j.node] = nd;
            bvh.emplace_back(); bvh.emplace_back();
            jobs.push_back(Job{nd.left, j.first, mid - j.first});
            jobs.push_back(Job{nd.left + 1, mid, j.first + j.count - mid});
        }
    }

    struct Hit { float t, u, v; uint32_t prim; };   // prim: triangle index, or kSphereBit | sphere index
    static const uint32_t kSphereBit = 0x80000000u;
    uint32_t hitShape(const Hit &h) const { return (h.prim & kSphereBit) ? (uint32_t) spheres[h.prim & ~kSphereBit].shape : triShape[h.prim]; }
    // the normal ShapeKDTree::rayIntersect(ray, t, shape, n, uv) reports (skdtree.cpp:165-175: plain face normal; spheres: geoFrame.n)
    F3 hitGeoNormal(const Hit &h, F3 ro, F3 rd) const {
        if (h.prim & kSphereBit) {
            const ppg_sphere &sp = spheres[h.prim & ~kSphereBit];
            const F3 c = f3(sp.center[0], sp.center[1], sp.center[2]);
            F3 p = ro + rd * h.t; p = c + normalize(p - c) * sp.radius;
            F3 n = normalize(p - c); if (sp.flip_normals) n = n * -1.0f;
            return n;
        }
        const F3 p0 = P[idx[3 * h.prim]], p1 = P[idx[3 * h.prim + 1]], p2 = P[idx[3 * h.prim + 2]];
        return normalize(cross(p1 - p0, p2 - p0));
    }

    // Sphere::rayIntersect (src/shapes/sphere.cpp:163-187) with solveQuadraticDouble (src/libcore/util.cpp:487-525): double precision
    static bool sphereIntersect(const ppg_sphere &sp, F3 ro, F3 rd, float mint, float maxt, float &t) {
        const double ox = (double) ro.x - (double) sp.center[0], oy = (double) ro.y - (double) sp.center[1], oz = (double) ro.z - (double) sp.center[2];
        const double dx = rd.x, dy = rd.y, dz = rd.z;
        const double A = dx * dx + dy * dy + dz * dz;
        const double B = 2 * (ox * dx + oy * dy + oz * dz);
        const double C = (ox * ox + oy * oy + oz * oz) - (double) (sp.radius * sp.radius);        // m_radius*m_radius is a float product
        double nearT, farT;
        if (A == 0) { if (B != 0) nearT = farT = -C / B; else return false; }
        else {
            const double discrim = B * B - 4.0f * A * C;
            if (discrim < 0) return false;
            const double sqrtDiscrim = std::sqrt(discrim);
            const double temp = B < 0 ? -0.5f * (B - sqrtDiscrim) : -0.5f * (B + sqrtDiscrim);
            nearT = temp / A; farT = C / temp;
            if (nearT > farT) std::swap(nearT, farT);
        }
        if (!(nearT <= maxt && farT >= mint)) return false;
        if (nearT < mint) { if (farT > maxt) return false; t = (float) farT; }
        else t = (float) nearT;
        return true;
    }

    // nearest hit in [mint, maxt]; ties on t go to the lower triangle index
    bool intersect(F3 o, F3 d, float mint, float maxt, Hit &hit) const {
        const bool any = intersectTriangles(o, d, mint, maxt, hit);
        bool sph = false;
        for (size_t k = 0; k < spheres.size(); ++k) {
            float t;
            if (sphereIntersect(spheres[k], o, d, mint, maxt, t) && t < hit.t) { hit.t = t; hit.u = hit.v = 0; hit.prim = kSphereBit | (uint32_t) k; sph = true; }
        }
        return any || sph;
    }
    bool intersectTriangles(F3 o, F3 d, float mint, float maxt, Hit &hit) const {
        hit.t = std::numeric_limits<float>::infinity(); hit.prim = 0xFFFFFFFFu;
        if (triShape.empty()) return false;
        const F3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        uint32_t stack[64]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const BvhNode &n = bvh[stack[--sp]];
            float t0 = mint, t1 = std::min(maxt, hit.t);
            bool miss = false;
            for (int a = 0; a < 3; ++a) {
                float ta = (n.bmin[a] - comp(o, a)) * comp(inv, a), tb = (n.bmax[a] - comp(o, a)) * comp(inv, a);
                if (ta > tb) std::swap(ta, tb);
                // widen conservatively: degenerate (flat) boxes and NaNs from 0*inf must not cull
                if (!(ta != ta)) t0 = std::max(t0, ta * (ta > 0 ? 1.0f - 1e-6f : 1.0f + 1e-6f));      // multiplicative: inf - inf would be NaN and never cull
                if (!(tb != tb)) t1 = std::min(t1, tb * (tb > 0 ? 1.0f + 1e-6f : 1.0f - 1e-6f));
                if (t0 > t1) { miss = true; break; }
            }
            if (miss) continue;
            if (n.count) {
                for (uint32_t i = n.left; i < n.left + n.count; ++i) {
                    const uint32_t p = primOrder[i];
                    float u, v, t;
                    if (triaccel_intersect(accel[p], o, d, mint, maxt, u, v, t)) {
                        if (t < hit.t || (t == hit.t && p < hit.prim)) { hit.t = t; hit.u = u; hit.v = v; hit.prim = p; }
                    }
                }
            } else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
        }
        return hit.prim != 0xFFFFFFFFu;
    }
};

// Intersection record: the fields of render/shape.h:36 Intersection the path uses
struct Its {
    bool valid; float t; F3 p, geoN, shN, shS, shT, wi; uint32_t shape;
    float uvU = 0, uvV = 0; uint32_t prim = 0xFFFFFFFFu; F3 bary;      // texture coordinates (skdtree.h:398-405) and the triangle hit (for the UV tangents)
    F3 toLocal(F3 v) const { return f3(dot(v, shS), dot(v, shT), dot(v, shN)); }
    F3 toWorld(F3 v) const { return shS * v.x + shT * v.y + shN * v.z; }
};

// ShapeKDTree::rayIntersect incl. the adaptive epsilon (src/librender/skdtree.cpp:112-142)
// + fillIntersectionRecord (include/mitsuba/render/skdtree.h:343-428) + computeShadingFrame (src/libcore/util.cpp:603-608)
static inline bool ray_intersect(const Scene &sc, F3 o, F3 d, float mint, float maxt, Its &its) {
    its.valid = false; its.t = std::numeric_limits<float>::infinity();
    if (mint == kEpsilon)
        mint *= std::max(std::max(std::max(std::fabs(o.x), std::fabs(o.y)), std::fabs(o.z)), kEpsilon);
    Scene::Hit h;
    if (!sc.intersect(o, d, mint, maxt, h)) return false;
    its.valid = true; its.t = h.t;
    if (h.prim & Scene::kSphereBit) {          // Sphere::fillIntersectionRecord (src/shapes/sphere.cpp:209-255), identity rotation
        const ppg_sphere &sp = sc.spheres[h.prim & ~Scene::kSphereBit];
        const F3 c = f3(sp.center[0], sp.center[1], sp.center[2]);
        its.p = o + d * h.t;
        its.p = c + normalize(its.p - c) * sp.radius;        // re-projection (single precision)
        const F3 local = its.p - c;
        const F3 dpdu = f3(-local.y, local.x, 0) * (2 * kPiT);
        its.geoN = normalize(its.p - c);
        if (sp.flip_normals) its.geoN = its.geoN * -1.0f;
        its.shN = its.geoN;
        its.shape = (uint32_t) sp.shape;
        its.shS = normalize(dpdu - its.shN * dot(its.shN, dpdu));
        its.shT = cross(its.shN, its.shS);
        its.wi = its.toLocal(-d);
        its.prim = h.prim; its.uvU = its.uvV = 0;      // (textured / bump-mapped spheres are refused by the loader)
        return true;
    }
    const uint32_t i0 = sc.idx[3 * h.prim], i1 = sc.idx[3 * h.prim + 1], i2 = sc.idx[3 * h.prim + 2];
    const F3 p0 = sc.P[i0], p1 = sc.P[i1], p2 = sc.P[i2];
    const F3 b = f3(1 - h.u - h.v, h.u, h.v);
    its.p = p0 * b.x + p1 * b.y + p2 * b.z;
    const F3 side1 = p1 - p0, side2 = p2 - p0;
    F3 faceN = cross(side1, side2);
    const float len = length(faceN);
    if (!is_zero(faceN)) faceN = faceN * (1.0f / len);   // Normal::operator/= (reciprocal multiply)
    its.shape = sc.triShape[h.prim];
    its.prim = h.prim; its.bary = b;
    if (sc.shapes[its.shape].has_uvs) {            // its.uv = t0 * b.x + t1 * b.y + t2 * b.z (skdtree.h:398-403)
        its.uvU = sc.UV[2 * i0] * b.x + sc.UV[2 * i1] * b.y + sc.UV[2 * i2] * b.z;
        its.uvV = sc.UV[2 * i0 + 1] * b.x + sc.UV[2 * i1 + 1] * b.y + sc.UV[2 * i2 + 1] * b.z;
    } else { its.uvU = b.y; its.uvV = b.z; }
    if (sc.shapes[its.shape].has_normals) {
        its.shN = normalize(sc.N[i0] * b.x + sc.N[i1] * b.y + sc.N[i2] * b.z);
        if (dot(faceN, its.shN) < 0) faceN = -faceN;
    } else its.shN = faceN;
    its.geoN = faceN;
    // dpdu = side1 also for meshes WITH texture coordinates, where the reference uses the UV tangent (skdtree.h:373-380, trimesh.cpp:385): the two
    // frames differ by a rotation about n, which no isotropic BSDF can observe; the bumpmap wrapper, which can, computes the UV tangents itself
    const F3 dpdu = side1;
    its.shS = normalize(dpdu - its.shN * dot(its.shN, dpdu));
    its.shT = cross(its.shN, its.shS);
    its.wi = its.toLocal(-d);
    return true;
}

// ------------------------------------------------------------------ next event estimation
static const float kShadowEpsilon = 1e-3f;   // core/constants.h:29

// Scene::evalTransmittance without media / null surfaces (src/librender/scene.cpp:619-679): 1 if unoccluded, else 0.
// The shadow-ray query scales the epsilon WITHOUT the max(.., Epsilon) clamp (skdtree.cpp:154-158).
static inline bool occluded(const Scene &sc, F3 p1, F3 d, float remaining) {
    float mint = kEpsilon * std::max(std::max(std::fabs(p1.x), std::fabs(p1.y)), std::fabs(p1.z));
    Scene::Hit h;
    return sc.intersect(p1, d, mint, remaining * (1 - kShadowEpsilon), h);
}
static inline bool bsdf_has_null(const ppg_bsdf &b);
static inline F3 bsdf_eval_null(const ppg_bsdf &b, float cosThetaI);
// Scene::evalTransmittance with index-matched surfaces (scene.cpp:619-679, p1 and p2 on surfaces): a null surface multiplies its
// straight-through transmittance (evaluated in the GEOMETRIC frame, :650-655) and the ray continues behind it, at most
// maxInteractions times (maxDepth - depth - 1; negative = unlimited); anything else blocks.
static inline F3 eval_transmittance(const Scene &sc, F3 p1, F3 d, float remaining, int maxInteractions) {
    const float lengthFactor = 1 - kShadowEpsilon;
    F3 ro = p1, transmittance = f3(1, 1, 1);
    int interactions = 0;
    float maxt = remaining * lengthFactor;
    while (remaining > 0) {
        const float mint = kEpsilon * std::max(std::max(std::fabs(ro.x), std::fabs(ro.y)), std::fabs(ro.z));
        Scene::Hit h;
        const bool surface = sc.intersect(ro, d, mint, maxt, h);
        if (!surface) break;
        const ppg_bsdf &b = sc.bsdfs[sc.shapes[sc.hitShape(h)].bsdf];
        if (interactions == maxInteractions || !bsdf_has_null(b)) return f3(0, 0, 0);
        const F3 n = sc.hitGeoNormal(h, ro, d);
        transmittance = transmittance * bsdf_eval_null(b, -dot(n, d));
        if (is_zero(transmittance)) break;
        if (++interactions > 100) break;
        ro = ro + d * h.t; remaining -= h.t; maxt = remaining * lengthFactor;
    }
    return transmittance;
}

struct DirectSample { F3 value, d, n; float dist, pdf; int emitter; };
static const int kEnvEmitter = -2;       // `dRec.object` is the environment emitter
// Scene::sampleAttenuatedEmitterDirect (scene.cpp:876-897) -> AreaLight::sampleDirect (emitters/area.cpp:158-173)
// -> Shape::sampleDirect (shape.cpp:102-115) -> TriMesh::samplePosition (trimesh.cpp:412-423) -> Triangle::sample (libcore/triangle.cpp:24-59)
static inline bool sample_emitter_direct(const Scene &sc, F3 ref, F3 refN, float sx, float sy, DirectSample &out, int maxInteractions = 0) {
    if (sc.emitterCdf.size() < 2) return false;
    const size_t ei = Scene::cdfSample(sc.emitterCdf, sx);
    const float emPdf = sc.emitterCdf[ei + 1] - sc.emitterCdf[ei];
    sx = (sx - sc.emitterCdf[ei]) / (sc.emitterCdf[ei + 1] - sc.emitterCdf[ei]);            // sampleReuse
    if (ei == sc.emitterSamplers.size()) {   // the environment emitter: EnvironmentMap::sampleDirect, src/emitters/envmap.cpp:516-543 (no dRec.refN test there)
        F3 dl, value; float pdf;
        sc.envSampleDirection(sx, sy, dl, value, pdf);
        const float *m = sc.envToWorld;
        const F3 d = f3(m[0] * dl.x + m[1] * dl.y + m[2] * dl.z, m[3] * dl.x + m[4] * dl.y + m[5] * dl.z, m[6] * dl.x + m[7] * dl.y + m[8] * dl.z);
        out.pdf = 0; out.value = f3(0, 0, 0); out.d = d; out.emitter = kEnvEmitter;
        // m_sceneBSphere.rayIntersect (bsphere.h:88-95) -> solveQuadratic (util.cpp:447-485)
        const F3 o = ref - sc.envCenter;
        const float A = dot(d, d), B = 2 * dot(o, d), C = dot(o, o) - sc.envRadius * sc.envRadius;
        float nearT, farT;
        {   if (A == 0) return false;
            const float discrim = B * B - 4.0f * A * C;
            if (discrim < 0) return false;
            const float sq = std::sqrt(discrim), temp = B < 0 ? -0.5f * (B - sq) : -0.5f * (B + sq);
            nearT = temp / A; farT = C / temp; if (nearT > farT) std::swap(nearT, farT); }
        if (is_zero(value) || pdf == 0 || nearT >= 0 || farT <= 0) return false;
        const F3 p = ref + d * farT;
        out.n = normalize(sc.envCenter - p); out.dist = farT;
        out.value = value * (1.0f / pdf);                                                       // Spectrum / Float multiplies by the reciprocal
        out.value = out.value * eval_transmittance(sc, ref, d, farT, maxInteractions);          // isOnSurface(): EOnSurface is set (envmap.cpp:107)
        out.value = out.value * (1.0f / emPdf);
        out.pdf = pdf * emPdf;
        return true;
    }
    const Scene::EmitterSampler &E = sc.emitterSamplers[ei];
    if (E.sphere >= 0) {                     // Sphere::sampleDirect, src/shapes/sphere.cpp:286-355
        const ppg_sphere &sp = sc.spheres[E.sphere];
        const F3 c = f3(sp.center[0], sp.center[1], sp.center[2]);
        const F3 refToCenter = c - ref;
        const float refDist2 = dot(refToCenter, refToCenter);
        const float invRefDist = 1.0f / std::sqrt(refDist2);
        const float sinAlpha = sp.radius * invRefDist;
        F3 d, n; float pdf, dist;
        if (sinAlpha < 1 - kEpsilon) {       // outside: uniform cone
            const float cosAlpha = std::sqrt(std::max(0.0f, 1.0f - sinAlpha * sinAlpha));
            const F3 fn = refToCenter * invRefDist; F3 fs, ft; coordinate_system(fn, fs, ft);
            const float cosTheta = (1 - sx) + sx * cosAlpha, sinTheta = std::sqrt(std::max(0.0f, 1.0f - cosTheta * cosTheta));
            float sinPhi, cosPhi; sincosf(2.0f * kPiT * sy, &sinPhi, &cosPhi);
            const F3 lv = f3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta);
            d = fs * lv.x + ft * lv.y + fn * lv.z;
            pdf = (0.5f * kInvPi) / (1 - cosAlpha);                                            // INV_TWOPI / (1-cosCutoff)
            const float projDist = dot(refToCenter, d);
            const float baseT = refDist2 / projDist;
            const F3 query = ref + d * baseT;
            const F3 queryToCenter = c - query;
            const float queryDist2 = dot(queryToCenter, queryToCenter), queryProjDist = dot(queryToCenter, d);
            const float A = 1.0f, B = -2 * queryProjDist, C = queryDist2 - sp.radius * sp.radius;
            float nearT;
            { const float discrim = B * B - 4.0f * A * C;                                       // solveQuadratic, util.cpp:447-485
              if (discrim < 0) nearT = queryProjDist;
              else { const float sq = std::sqrt(discrim); const float temp = B < 0 ? -0.5f * (B - sq) : -0.5f * (B + sq); float x0 = temp / A, x1 = C / temp; if (x0 > x1) std::swap(x0, x1); nearT = x0; } }
            dist = baseT + nearT;
            n = normalize(d * nearT - queryToCenter);
        } else {                             // inside: uniform sphere
            const float z = 1.0f - 2.0f * sy, r = std::sqrt(std::max(0.0f, 1.0f - z * z));
            float sinPhi, cosPhi; sincosf(2.0f * kPiT * sx, &sinPhi, &cosPhi);
            const F3 v = f3(r * cosPhi, r * sinPhi, z);
            const F3 p = c + v * sp.radius;
            n = v; d = p - ref;
            const float dist2 = dot(d, d);
            dist = std::sqrt(dist2);
            d = d * (1.0f / dist);                                                              // Vector /= Float: reciprocal multiply
            pdf = E.invArea * dist2 / std::fabs(dot(d, n));
        }
        if (sp.flip_normals) n = n * -1.0f;
        out.d = d; out.n = n; out.dist = dist; out.emitter = (int) ei;
        if (dot(d, refN) >= 0 && dot(d, n) < 0 && pdf != 0) out.value = sc.radiance[ei] * (1.0f / pdf);
        else { out.pdf = 0; out.value = f3(0, 0, 0); return false; }
        out.value = out.value * eval_transmittance(sc, ref, d, dist, maxInteractions);          // value *= evalTransmittance(...), scene.cpp:886-889
        out.value = out.value * (1.0f / emPdf);
        out.pdf = pdf * emPdf;
        return true;
    }
    if (E.nTris == 0) return false;
    const size_t ti = Scene::cdfSample(E.cdf, sy);
    sy = (sy - E.cdf[ti]) / (E.cdf[ti + 1] - E.cdf[ti]);
    const uint32_t t = E.firstTri + (uint32_t) ti;
    const F3 p0 = sc.P[sc.idx[3 * t]], p1 = sc.P[sc.idx[3 * t + 1]], p2 = sc.P[sc.idx[3 * t + 2]];
    const float a = std::sqrt(std::max(0.0f, 1.0f - sx));                                      // warp::squareToUniformTriangle
    const float bx = 1 - a, by = a * sy;
    const F3 sideA = p1 - p0, sideB = p2 - p0;
    const F3 p = p0 + (sideA * bx) + (sideB * by);
    F3 n;
    if (sc.shapes[E.shape].has_normals) n = normalize(sc.N[sc.idx[3 * t]] * (1.0f - bx - by) + sc.N[sc.idx[3 * t + 1]] * bx + sc.N[sc.idx[3 * t + 2]] * by);
    else n = normalize(cross(sideA, sideB));
    float pdf = E.invArea;
    F3 d = p - ref;
    const float distSquared = dot(d, d);
    const float dist = std::sqrt(distSquared);
    d = d * (1.0f / dist);
    const float dp = std::fabs(dot(d, n));
    pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
    out.d = d; out.n = n; out.dist = dist; out.emitter = (int) ei;
    if (dot(d, refN) >= 0 && dot(d, n) < 0 && pdf != 0) out.value = sc.radiance[ei] * (1.0f / pdf);
    else { out.pdf = 0; out.value = f3(0, 0, 0); return false; }
    out.value = out.value * eval_transmittance(sc, ref, d, dist, maxInteractions);
    out.value = out.value * (1.0f / emPdf);          // value *= transmittance / emPdf
    out.pdf = pdf * emPdf;
    return true;
}
// Scene::pdfEmitterDirect (scene.cpp:949-952) for an emitter hit found by BSDF / guiding sampling (dRec.setQuery, records.inl:170-178)
static inline float pdf_emitter_direct(const Scene &sc, int emitter, F3 ref, F3 refN, F3 d, F3 n, float dist) {
    if (emitter == kEnvEmitter) return sc.envPdfDirection(d) * (1.0f * sc.emitterNormalization);   // EnvironmentMap::pdfDirect, ESolidAngle (fillDirectSamplingRecord, envmap.cpp:371)
    if (!(dot(d, refN) >= 0 && dot(d, n) < 0)) return 0.0f;                                    // AreaLight::pdfDirect (area.cpp:175-183)
    if (sc.emitterSamplers[emitter].sphere >= 0) {                                             // Sphere::pdfDirect, sphere.cpp:357-392
        const ppg_sphere &sp = sc.spheres[sc.emitterSamplers[emitter].sphere];
        const F3 refToCenter = f3(sp.center[0], sp.center[1], sp.center[2]) - ref;
        const float invRefDist = 1.0f / length(refToCenter);
        const float sinAlpha = sp.radius * invRefDist;
        float pdfSA;
        if (sinAlpha < 1 - kEpsilon) { const float cosAlpha = std::sqrt(std::max(0.0f, 1 - sinAlpha * sinAlpha)); pdfSA = (0.5f * kInvPi) / (1 - cosAlpha); }
        else pdfSA = sc.emitterSamplers[emitter].invArea * dist * dist / std::fabs(dot(d, n));
        return pdfSA * (1.0f * sc.emitterNormalization);
    }
    const float pdfPos = sc.emitterSamplers[emitter].invArea;
    return pdfPos * (dist * dist) / std::fabs(dot(d, n)) * (1.0f * sc.emitterNormalization);   // Shape::pdfDirect (shape.cpp:117-126) * pdfEmitterDiscrete
}
static inline float mi_weight(float pdfA, float pdfB) { pdfA *= pdfA; pdfB *= pdfB; return pdfA / (pdfA + pdfB); }   // GP:2247-2250

// ------------------------------------------------------------------ BSDF: diffuse (+ twosided)
// src/libcore/warp.cpp:81-102 then :43-52
static inline F3 square_to_cosine_hemisphere(float sx, float sy) {
    const float r1 = 2.0f * sx - 1.0f, r2 = 2.0f * sy - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (kPiT / 4.0f) * (r2 / r1); }
    else { r = r2; phi = (kPiT / 2.0f) - (r1 / r2) * (kPiT / 4.0f); }
    const float px = r * std::cos(phi), py = r * std::sin(phi);
    float z = std::sqrt(std::max(0.0f, 1.0f - px * px - py * py));   // math::safe_sqrt
    if (z == 0) z = 1e-10f;
    return f3(px, py, z);
}
struct BsdfSample { F3 wo; float eta; bool delta; bool null = false; };   // null: sampledType == ENull (index-matched transition)
static inline bool bsdf_has_null(const ppg_bsdf &b) { return b.type == PPG_BSDF_THINDIELECTRIC || (b.flags & PPG_BSDF_FLAG_MASK); }                              // type & ENull
static inline bool bsdf_has_smooth(const ppg_bsdf &b) { return b.type == PPG_BSDF_DIFFUSE || b.type == PPG_BSDF_NULL_BLACK || b.type == PPG_BSDF_ROUGHCONDUCTOR || b.type == PPG_BSDF_ROUGHPLASTIC || b.type == PPG_BSDF_ROUGHDIELECTRIC || b.type == PPG_BSDF_PLASTIC; }   // type & ESmooth = diffuse | glossy (bsdf.h:224-285)
static inline bool bsdf_has_transmission_or_backside(const ppg_bsdf &b) { return (b.flags & (PPG_BSDF_FLAG_TWOSIDED | PPG_BSDF_FLAG_MASK)) || b.type == PPG_BSDF_DIELECTRIC || b.type == PPG_BSDF_ROUGHDIELECTRIC || b.type == PPG_BSDF_THINDIELECTRIC; }

// fresnelDielectricExt, src/libcore/util.cpp:651-683
static inline float fresnel_dielectric_ext(float cosThetaI_, float &cosThetaT_, float eta) {
    if (eta == 1) { cosThetaT_ = -cosThetaI_; return 0.0f; }
    const float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    const float cosThetaI = std::fabs(cosThetaI_), cosThetaT = std::sqrt(cosThetaTSqr);
    const float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    const float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
// fresnelConductorExact (scalar form applied per channel), src/libcore/util.cpp:715-738
static inline float fresnel_conductor_exact(float cosThetaI, float eta, float k) {
    const float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    const float temp1 = eta * eta - k * k - sinThetaI2;
    const float a2pb2 = std::sqrt(std::max(0.0f, temp1 * temp1 + k * k * eta * eta * 4));
    const float a = std::sqrt(std::max(0.0f, (a2pb2 + temp1) * 0.5f));
    const float term1 = a2pb2 + cosThetaI2, term2 = a * (2 * cosThetaI);
    const float Rs2 = (term1 - term2) / (term1 + term2);
    const float term3 = a2pb2 * cosThetaI2 + sinThetaI4, term4 = term2 * sinThetaI2;
    const float Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (Rp2 + Rs2);
}
// ---- thindielectric (src/bsdfs/thindielectric.cpp): R' = R + T R T + T R^3 T + ... (:160-165)
static inline float thindielectric_reflectance(float cosThetaI, float eta) {
    float ct; float R = fresnel_dielectric_ext(std::fabs(cosThetaI), ct, eta); const float T = 1 - R;
    if (R < 1) R += T * T * R / (1 - R * R);
    return R;
}
// bsdf->eval(bRec, EDiscrete) with typeMask == ENull and wo == -wi (thindielectric.cpp:153-176): what a straight-through ray keeps
static inline F3 bsdf_eval_null(const ppg_bsdf &b, float cosThetaI) {
    if (b.flags & PPG_BSDF_FLAG_MASK) return f3(1 - b.opacity[0], 1 - b.opacity[1], 1 - b.opacity[2]);   // mask.cpp:118-119: Spectrum(1) - opacity
    if (b.type != PPG_BSDF_THINDIELECTRIC) return f3(0, 0, 0);
    const float R = thindielectric_reflectance(cosThetaI, b.eta[0]);
    return f3(b.specular_transmittance[0], b.specular_transmittance[1], b.specular_transmittance[2]) * (1 - R);
}
// ---- MicrofacetDistribution (isotropic, Beckmann / GGX, visible-normal sampling), restated from src/bsdfs/microfacet.h
static inline float mts_erfinv(float x) {       // math::erfinv, src/libcore/math.cpp:25-53 (Giles)
    float w = -std::log((1.0f - x) * (1.0f + x)), p;
    if (w < 5.0f) {
        w = w - 2.5f; p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w; p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w; p = 1.50140941f + p * w;
    } else {
        w = std::sqrt(w) - 3.0f; p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w; p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w; p = 2.83297682f + p * w;
    }
    return p * x;
}
static inline float mts_erf(float x) {          // math::erf, src/libcore/math.cpp:55-72 (A&S 7.1.26)
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
    const float sign = std::copysign(1.0f, x); x = std::fabs(x);
    const float t = 1.0f / (1.0f + p * x);
    const float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * std::exp(-x * x);
    return sign * y;
}
static inline float mts_hypot2(float a, float b) {   // math::hypot2, src/libcore/math.cpp:74-86
    float r;
    if (std::fabs(a) > std::fabs(b)) { r = b / a; r = std::fabs(a) * std::sqrt(1.0f + r * r); }
    else if (b != 0.0f) { r = a / b; r = std::fabs(b) * std::sqrt(1.0f + r * r); }
    else r = 0.0f;
    return r;
}
struct Microfacet {
    int type; float alpha;
    Microfacet(int t, float a) : type(t), alpha(std::max(a, 1e-4f)) {}                    // microfacet.h:60-67
    float eval(F3 m) const {                                                              // microfacet.h:191-236
        if (m.z <= 0) return 0.0f;
        const float cosTheta2 = m.z * m.z;
        const float beckmannExponent = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
        float result;
        if (type == PPG_MICROFACET_BECKMANN) result = std::exp(-beckmannExponent) / (kPiT * alpha * alpha * cosTheta2 * cosTheta2);
        else { const float root = (1.0f + beckmannExponent) * cosTheta2; result = 1.0f / (kPiT * alpha * alpha * root * root); }
        if (result * m.z < 1e-20f) result = 0;
        return result;
    }
    float smithG1(F3 v, F3 m) const {                                                     // microfacet.h:477-517
        if (dot(v, m) * v.z <= 0) return 0.0f;
        const float temp = 1 - v.z * v.z;
        const float tanTheta = std::fabs(temp <= 0.0f ? 0.0f : std::sqrt(temp) / v.z);   // Frame::tanTheta
        if (tanTheta == 0.0f) return 1.0f;
        if (type == PPG_MICROFACET_BECKMANN) {
            const float a = 1.0f / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0f;
            const float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
        }
        const float root = alpha * tanTheta;
        return 2.0f / (1.0f + mts_hypot2(1.0f, root));
    }
    float pdfVisible(F3 wi, F3 m) const {                                                 // microfacet.h:462-466
        if (wi.z == 0) return 0.0f;
        return smithG1(wi, m) * std::fabs(dot(wi, m)) * eval(m) / std::fabs(wi.z);
    }
    void sampleVisible11(float thetaI, float sx, float sy, float &slopeX, float &slopeY) const {   // microfacet.h:573-690
        const float SQRT_PI_INV = 1 / std::sqrt(kPiT);
        if (type == PPG_MICROFACET_BECKMANN) {
            if (thetaI < 1e-4f) { const float r = std::sqrt(-std::log(1.0f - sx)); const float ph = 2 * kPiT * sy; slopeX = r * std::cos(ph); slopeY = r * std::sin(ph); return; }
            const float tanThetaI = std::tan(thetaI), cotThetaI = 1 / tanThetaI;
            float a = -1, c = mts_erf(cotThetaI);
            const float sample_x = std::max(sx, 1e-6f);
            const float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            float b = c - (1 + c) * std::pow(1 - sample_x, fit);
            const float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * std::exp(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                const float invErf = mts_erfinv(b);
                const float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * std::exp(-invErf * invErf)) - sample_x;
                const float derivative = normalization * (1 - invErf * tanThetaI);
                if (std::fabs(value) < 1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slopeX = mts_erfinv(b);
            slopeY = mts_erfinv(2.0f * std::max(sy, 1e-6f) - 1.0f);
            return;
        }
        if (thetaI < 1e-4f) { const float r = std::sqrt(std::max(0.0f, sx / (1 - sx))); const float ph = 2 * kPiT * sy; slopeX = r * std::cos(ph); slopeY = r * std::sin(ph); return; }
        const float tanThetaI = std::tan(thetaI), a = 1 / tanThetaI;
        const float G1 = 2.0f / (1.0f + std::sqrt(std::max(0.0f, 1.0f + 1.0f / (a * a))));
        float A = 2.0f * sx / G1 - 1.0f;
        if (std::fabs(A) == 1) A -= std::copysign(1.0f, A) * kEpsilon;
        const float tmp = 1.0f / (A * A - 1.0f), B = tanThetaI;
        const float D = std::sqrt(std::max(0.0f, B * B * tmp * tmp - (A * A - B * B) * tmp));
        const float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
        slopeX = (A < 0.0f || slope_x_2 > 1.0f / tanThetaI) ? slope_x_1 : slope_x_2;
        float S;
        if (sy > 0.5f) { S = 1.0f; sy = 2.0f * (sy - 0.5f); } else { S = -1.0f; sy = 2.0f * (0.5f - sy); }
        const float z = (sy * (sy * (sy * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                        (sy * (sy * (sy * (sy * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
        slopeY = S * z * std::sqrt(1.0f + slopeX * slopeX);
    }
    F3 sampleVisible(F3 _wi, float sx, float sy) const {                                  // microfacet.h:421-459
        const F3 wi = normalize(f3(alpha * _wi.x, alpha * _wi.y, _wi.z));
        float theta = 0, phi = 0;
        if (wi.z < 0.99999f) { theta = std::acos(wi.z); phi = std::atan2(wi.y, wi.x); }
        const float sinPhi = std::sin(phi), cosPhi = std::cos(phi);
        float slx, sly; sampleVisible11(theta, sx, sy, slx, sly);
        float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
        rx *= alpha; ry *= alpha;
        const float normalization = 1.0f / std::sqrt(rx * rx + ry * ry + 1.0f);
        return f3(-rx * normalization, -ry * normalization, normalization);
    }
};
static inline F3 fresnel_conductor_rgb(float c, const ppg_bsdf &b) {
    return f3(b.reflectance[0] * fresnel_conductor_exact(c, b.eta[0], b.k[0]), b.reflectance[1] * fresnel_conductor_exact(c, b.eta[1], b.k[1]),
              b.reflectance[2] * fresnel_conductor_exact(c, b.eta[2], b.k[2]));
}
// roughconductor.cpp:257-283 (eval), :285-312 (pdf), :355-404 (sample)
static inline F3 roughconductor_eval(const ppg_bsdf &b, F3 wi, F3 wo) {
    if (wi.z <= 0 || wo.z <= 0) return f3(0, 0, 0);
    const F3 H = normalize(wo + wi);
    const Microfacet distr(b.distribution, b.alpha);
    const float D = distr.eval(H);
    if (D == 0) return f3(0, 0, 0);
    const F3 F = fresnel_conductor_rgb(dot(wi, H), b);
    const float G = distr.smithG1(wi, H) * distr.smithG1(wo, H);
    const float model = D * G / (4.0f * wi.z);
    return F * model;
}
static inline float roughconductor_pdf(const ppg_bsdf &b, F3 wi, F3 wo) {
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    const F3 H = normalize(wo + wi);
    const Microfacet distr(b.distribution, b.alpha);
    return distr.eval(H) * distr.smithG1(wi, H) / (4.0f * wi.z);
}
static inline F3 roughconductor_sample(const ppg_bsdf &b, F3 wi, float sx, float sy, F3 &wo, float &pdf) {
    pdf = 0;
    if (wi.z < 0) return f3(0, 0, 0);
    const Microfacet distr(b.distribution, b.alpha);
    const F3 m = distr.sampleVisible(wi, sx, sy);
    pdf = distr.pdfVisible(wi, m);
    if (pdf == 0) return f3(0, 0, 0);
    wo = m * (2 * dot(wi, m)) - wi;                       // reflect(wi, m) = 2 * dot(wi, m) * Vector(m) - wi
    if (wo.z <= 0) return f3(0, 0, 0);
    const F3 F = fresnel_conductor_rgb(dot(wi, m), b);
    const float weight = distr.smithG1(wo, m);
    pdf /= 4.0f * dot(wo, m);
    return F * weight;
}

// ---- roughplastic (src/bsdfs/roughplastic.cpp). The external rough transmittance of the material (constant eta, alpha) is the 1-D table
// RoughTransmittance::eval reads once alpha and eta are fixed (src/bsdfs/rtrans.h:183-193, 233), interpolated by evalCubicInterp1D
// (src/libcore/spline.cpp:23-60) over cos(theta)^(1/4).
static inline float rough_transmittance(const float *values, float cosTheta) {
    if (!(cosTheta >= 0)) return 0.0f;
    const float x = std::pow(std::fabs(cosTheta), 0.25f);
    const size_t size = PPG_BSDF_TABLE_SIZE;
    float result = 0.0f;
    if (x >= 0.0f && x <= 1.0f) {
        float t = ((x - 0.0f) * (size - 1)) / (1.0f - 0.0f);
        const size_t k = std::max((size_t)0, std::min((size_t)t, size - 2));
        const float f0 = values[k], f1 = values[k + 1];
        const float d0 = k > 0 ? 0.5f * (values[k + 1] - values[k - 1]) : values[k + 1] - values[k];
        const float d1 = k + 2 < size ? 0.5f * (values[k + 2] - values[k]) : values[k + 1] - values[k];
        t = t - (float)k;
        const float t2 = t * t, t3 = t2 * t;
        result = (2 * t3 - 3 * t2 + 1) * f0 + (-2 * t3 + 3 * t2) * f1 + (t3 - 2 * t2 + t) * d0 + (t3 - t2) * d1;
    }
    return std::min(1.0f, std::max(0.0f, result));
}
static inline float roughplastic_prob_specular(const ppg_bsdf &b, const float *lut, float cosThetaI) {   // roughplastic.cpp:403-409 = :446-452
    float probSpecular = 1 - rough_transmittance(lut, cosThetaI);
    return (probSpecular * b.specular_sampling_weight) / (probSpecular * b.specular_sampling_weight + (1 - probSpecular) * (1 - b.specular_sampling_weight));
}
static inline F3 roughplastic_eval(const ppg_bsdf &b, const float *lut, F3 wi, F3 wo) {                  // roughplastic.cpp:326-380
    if (wi.z <= 0 || wo.z <= 0) return f3(0, 0, 0);
    const Microfacet distr(b.distribution, b.alpha);
    const F3 H = normalize(wo + wi);
    const float D = distr.eval(H);
    float cosThetaT; const float F = fresnel_dielectric_ext(dot(wi, H), cosThetaT, b.eta[0]);
    const float G = distr.smithG1(wi, H) * distr.smithG1(wo, H);
    const float value = F * D * G / (4.0f * wi.z);
    F3 result = f3(b.specular_reflectance[0], b.specular_reflectance[1], b.specular_reflectance[2]) * value;
    F3 diff = f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
    const float T12 = rough_transmittance(lut, wi.z), T21 = rough_transmittance(lut, wo.z), Fdr = b.fdr_int;
    if (b.flags & PPG_BSDF_FLAG_NONLINEAR) diff = f3(diff.x / (1.0f - diff.x * Fdr), diff.y / (1.0f - diff.y * Fdr), diff.z / (1.0f - diff.z * Fdr));
    else diff = diff * (1.0f / (1 - Fdr));                   // Spectrum /= Float multiplies by the reciprocal (core/spectrum.h:447-456)
    const float invEta2 = 1 / (b.eta[0] * b.eta[0]);
    return result + diff * (kInvPi * wo.z * T12 * T21 * invEta2);
}
static inline float roughplastic_pdf(const ppg_bsdf &b, const float *lut, F3 wi, F3 wo) {                // roughplastic.cpp:382-430
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    const Microfacet distr(b.distribution, b.alpha);
    const F3 H = normalize(wo + wi);
    const float probSpecular = roughplastic_prob_specular(b, lut, wi.z), probDiffuse = 1 - probSpecular;
    const float dwh_dwo = 1.0f / (4.0f * dot(wo, H));
    const float prob = distr.pdfVisible(wi, H);
    float result = prob * dwh_dwo * probSpecular;
    result += probDiffuse * (kInvPi * wo.z);
    return result;
}
static inline F3 roughplastic_sample(const ppg_bsdf &b, const float *lut, F3 wi, float sx, float sy, F3 &wo, float &pdf) {   // roughplastic.cpp:432-497
    pdf = 0;
    if (wi.z <= 0) return f3(0, 0, 0);
    bool choseSpecular = true;
    const Microfacet distr(b.distribution, b.alpha);
    const float probSpecular = roughplastic_prob_specular(b, lut, wi.z);
    if (sy < probSpecular) sy /= probSpecular;
    else { sy = (sy - probSpecular) / (1 - probSpecular); choseSpecular = false; }
    if (choseSpecular) {
        const F3 m = distr.sampleVisible(wi, sx, sy);
        wo = m * (2 * dot(wi, m)) - wi;
        if (wo.z <= 0) return f3(0, 0, 0);
    } else wo = square_to_cosine_hemisphere(sx, sy);
    pdf = roughplastic_pdf(b, lut, wi, wo);
    if (pdf == 0) return f3(0, 0, 0);
    return roughplastic_eval(b, lut, wi, wo) * (1.0f / pdf);  // Spectrum / Float, core/spectrum.h:415-425
}
// ---- roughdielectric (src/bsdfs/roughdielectric.cpp), visible-normal sampling (m_sampleVisible, the default: no Walter alpha scaling).
// sample() draws ONE extra number from the path's sampler to choose reflection / refraction (EUsesSampler, :536-543).
static inline float mts_signum(float v) { return std::copysign(1.0f, v); }                             // core/math.h:269-278
static inline F3 roughdielectric_eval(const ppg_bsdf &b, F3 wi, F3 wo) {                                 // roughdielectric.cpp:270-350
    if (wi.z == 0) return f3(0, 0, 0);
    const float m_eta = b.eta[0], m_invEta = 1 / m_eta;
    const bool reflect = wi.z * wo.z > 0;
    F3 H;
    if (reflect) H = normalize(wo + wi);
    else { const float eta = wi.z > 0 ? m_eta : m_invEta; H = normalize(wi + wo * eta); }
    H = H * mts_signum(H.z);
    const Microfacet distr(b.distribution, b.alpha);
    const float D = distr.eval(H);
    if (D == 0) return f3(0, 0, 0);
    float cosThetaT; const float F = fresnel_dielectric_ext(dot(wi, H), cosThetaT, m_eta);
    const float G = distr.smithG1(wi, H) * distr.smithG1(wo, H);
    if (reflect) {
        const float value = F * D * G / (4.0f * std::fabs(wi.z));
        return f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]) * value;
    }
    const float eta = wi.z > 0.0f ? m_eta : m_invEta;
    const float sqrtDenom = dot(wi, H) + eta * dot(wo, H);
    const float value = ((1 - F) * D * G * eta * eta * dot(wi, H) * dot(wo, H)) / (wi.z * sqrtDenom * sqrtDenom);
    const float factor = wi.z > 0 ? m_invEta : m_eta;                                                    // ERadiance
    return f3(b.specular_transmittance[0], b.specular_transmittance[1], b.specular_transmittance[2]) * std::fabs(value * factor * factor);
}
static inline float roughdielectric_pdf(const ppg_bsdf &b, F3 wi, F3 wo) {                               // roughdielectric.cpp:352-422
    const float m_eta = b.eta[0], m_invEta = 1 / m_eta;
    const bool reflect = wi.z * wo.z > 0;
    F3 H; float dwh_dwo;
    if (reflect) { H = normalize(wo + wi); dwh_dwo = 1.0f / (4.0f * dot(wo, H)); }
    else {
        const float eta = wi.z > 0 ? m_eta : m_invEta;
        H = normalize(wi + wo * eta);
        const float sqrtDenom = dot(wi, H) + eta * dot(wo, H);
        dwh_dwo = (eta * eta * dot(wo, H)) / (sqrtDenom * sqrtDenom);
    }
    H = H * mts_signum(H.z);
    const Microfacet distr(b.distribution, b.alpha);
    float prob = distr.pdfVisible(wi * mts_signum(wi.z), H);
    float cosThetaT; const float F = fresnel_dielectric_ext(dot(wi, H), cosThetaT, m_eta);
    prob *= reflect ? F : (1 - F);
    return std::fabs(prob * dwh_dwo);
}
static inline F3 roughdielectric_sample(const ppg_bsdf &b, F3 wi, float sx, float sy, float su, F3 &wo, float &etaOut, float &pdf) {   // roughdielectric.cpp:502-600
    pdf = 0;
    const float m_eta = b.eta[0], m_invEta = 1 / m_eta;
    const Microfacet distr(b.distribution, b.alpha);
    const F3 wiUp = wi * mts_signum(wi.z);
    const F3 m = distr.sampleVisible(wiUp, sx, sy);
    const float microfacetPDF = distr.pdfVisible(wiUp, m);
    if (microfacetPDF == 0) return f3(0, 0, 0);
    pdf = microfacetPDF;
    float cosThetaT; const float F = fresnel_dielectric_ext(dot(wi, m), cosThetaT, m_eta);
    F3 weight = f3(1, 1, 1);
    bool sampleReflection = true;
    if (su > F) { sampleReflection = false; pdf *= 1 - F; } else pdf *= F;
    float dwh_dwo;
    if (sampleReflection) {
        wo = m * (2 * dot(wi, m)) - wi; etaOut = 1.0f;
        if (wi.z * wo.z <= 0) return f3(0, 0, 0);
        weight = weight * f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
        dwh_dwo = 1.0f / (4.0f * dot(wo, m));
    } else {
        if (cosThetaT == 0) return f3(0, 0, 0);
        { const float eta = cosThetaT < 0 ? 1 / m_eta : m_eta; wo = m * (dot(wi, m) * eta + cosThetaT) - wi * eta; }   // refract(wi, m, eta, cosThetaT), util.cpp:767-772
        etaOut = cosThetaT < 0 ? m_eta : m_invEta;
        if (wi.z * wo.z >= 0) return f3(0, 0, 0);
        const float factor = cosThetaT < 0 ? m_invEta : m_eta;
        weight = weight * (f3(b.specular_transmittance[0], b.specular_transmittance[1], b.specular_transmittance[2]) * (factor * factor));
        const float sqrtDenom = dot(wi, m) + etaOut * dot(wo, m);
        dwh_dwo = (etaOut * etaOut * dot(wo, m)) / (sqrtDenom * sqrtDenom);
    }
    weight = weight * distr.smithG1(wo, m);
    pdf *= std::fabs(dwh_dwo);
    return weight;
}
// ---- plastic (src/bsdfs/plastic.cpp): delta reflection off the coat + diffuse base; eval / pdf in the solid-angle measure see the diffuse part only
static inline F3 plastic_diffuse(const ppg_bsdf &b) {                                                    // plastic.cpp:266-271
    F3 diff = f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
    if (b.flags & PPG_BSDF_FLAG_NONLINEAR) return f3(diff.x / (1.0f - diff.x * b.fdr_int), diff.y / (1.0f - diff.y * b.fdr_int), diff.z / (1.0f - diff.z * b.fdr_int));
    return diff * (1.0f / (1 - b.fdr_int));
}
static inline float plastic_prob_specular(const ppg_bsdf &b, float Fi) {                                 // plastic.cpp:292-294
    return (Fi * b.specular_sampling_weight) / (Fi * b.specular_sampling_weight + (1 - Fi) * (1 - b.specular_sampling_weight));
}
static inline F3 plastic_eval(const ppg_bsdf &b, F3 wi, F3 wo) {                                         // plastic.cpp:245-278
    if (wo.z <= 0 || wi.z <= 0) return f3(0, 0, 0);
    float ct; const float Fi = fresnel_dielectric_ext(wi.z, ct, b.eta[0]), Fo = fresnel_dielectric_ext(wo.z, ct, b.eta[0]);
    const float invEta2 = 1 / (b.eta[0] * b.eta[0]);
    return plastic_diffuse(b) * ((kInvPi * wo.z) * invEta2 * (1 - Fi) * (1 - Fo));
}
static inline float plastic_pdf(const ppg_bsdf &b, F3 wi, F3 wo) {                                       // plastic.cpp:280-308
    if (wo.z <= 0 || wi.z <= 0) return 0.0f;
    float ct; const float Fi = fresnel_dielectric_ext(wi.z, ct, b.eta[0]);
    return (kInvPi * wo.z) * (1 - plastic_prob_specular(b, Fi));
}
static inline F3 plastic_sample(const ppg_bsdf &b, F3 wi, float sx, float sy, F3 &wo, bool &delta, float &pdf) {   // plastic.cpp:374-441
    pdf = 0; delta = false;
    if (wi.z <= 0) return f3(0, 0, 0);
    float ct; const float Fi = fresnel_dielectric_ext(wi.z, ct, b.eta[0]);
    const float probSpecular = plastic_prob_specular(b, Fi);
    if (sx < probSpecular) {
        delta = true; wo = f3(-wi.x, -wi.y, wi.z); pdf = probSpecular;
        return (f3(b.specular_reflectance[0], b.specular_reflectance[1], b.specular_reflectance[2]) * Fi) * (1.0f / probSpecular);
    }
    wo = square_to_cosine_hemisphere((sx - probSpecular) / (1 - probSpecular), sy);
    const float Fo = fresnel_dielectric_ext(wo.z, ct, b.eta[0]);
    const float invEta2 = 1 / (b.eta[0] * b.eta[0]);
    pdf = (1 - probSpecular) * (kInvPi * wo.z);
    return plastic_diffuse(b) * (invEta2 * (1 - Fi) * (1 - Fo) / (1 - probSpecular));
}
static inline const float *bsdf_table(const ppg_bsdf &b, const float *tables) { return tables ? tables + (size_t)b.table * PPG_BSDF_TABLE_SIZE : nullptr; }

// eval / pdf with the solid-angle measure (delta models return 0); sample per src/bsdfs/{diffuse.cpp:110-150, dielectric.cpp:277-334, conductor.cpp:262-277};
// twosided per src/bsdfs/twosided.cpp:108-184
static inline F3 bsdf_eval_inner(const ppg_bsdf &b, F3 wi, F3 wo, const float *tables) {
    if (!bsdf_has_smooth(b)) return f3(0, 0, 0);
    if (b.flags & PPG_BSDF_FLAG_TWOSIDED) { if (wi.z < 0) { wi.z = -wi.z; wo.z = -wo.z; } }
    if (b.type == PPG_BSDF_ROUGHCONDUCTOR) return roughconductor_eval(b, wi, wo);
    if (b.type == PPG_BSDF_ROUGHDIELECTRIC) return roughdielectric_eval(b, wi, wo);
    if (b.type == PPG_BSDF_PLASTIC) return plastic_eval(b, wi, wo);
    if (b.type == PPG_BSDF_ROUGHPLASTIC) return roughplastic_eval(b, bsdf_table(b, tables), wi, wo);
    if (wi.z <= 0 || wo.z <= 0) return f3(0, 0, 0);
    return f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]) * (kInvPi * wo.z);
}
static inline float bsdf_pdf_inner(const ppg_bsdf &b, F3 wi, F3 wo, const float *tables) {
    if (!bsdf_has_smooth(b)) return 0.0f;
    if (b.flags & PPG_BSDF_FLAG_TWOSIDED) { if (wi.z < 0) { wi.z = -wi.z; wo.z = -wo.z; } }
    if (b.type == PPG_BSDF_ROUGHCONDUCTOR) return roughconductor_pdf(b, wi, wo);
    if (b.type == PPG_BSDF_ROUGHDIELECTRIC) return roughdielectric_pdf(b, wi, wo);
    if (b.type == PPG_BSDF_PLASTIC) return plastic_pdf(b, wi, wo);
    if (b.type == PPG_BSDF_ROUGHPLASTIC) return roughplastic_pdf(b, bsdf_table(b, tables), wi, wo);
    if (wi.z <= 0 || wo.z <= 0) return 0.0f;
    return kInvPi * wo.z;   // warp::squareToCosineHemispherePdf
}
// `rng`: the path's sampler, consumed only by models that draw from it themselves (roughdielectric)
static inline F3 bsdf_sample_inner(const ppg_bsdf &b, F3 wi, float sx, float sy, BsdfSample &s, float &pdf, const float *tables, Pcg32 *rng) {
    bool flip = false;
    if (b.flags & PPG_BSDF_FLAG_TWOSIDED) { if (wi.z < 0) { wi.z = -wi.z; flip = true; } }
    s.eta = 1.0f; s.delta = false; s.null = false; pdf = 0;
    if (b.type == PPG_BSDF_DIELECTRIC) {
        const float eta = b.eta[0], invEta = 1 / eta;
        float cosThetaT; const float F = fresnel_dielectric_ext(wi.z, cosThetaT, eta);
        s.delta = true;
        if (sx <= F) { s.wo = f3(-wi.x, -wi.y, wi.z); s.eta = 1.0f; pdf = F; return f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]); }
        const float scale = -(cosThetaT < 0 ? invEta : eta);
        s.wo = f3(scale * wi.x, scale * wi.y, cosThetaT); s.eta = cosThetaT < 0 ? eta : invEta; pdf = 1 - F;
        const float factor = cosThetaT < 0 ? invEta : eta;      // ERadiance: solid angle compression
        return f3(b.specular_transmittance[0], b.specular_transmittance[1], b.specular_transmittance[2]) * (factor * factor);
    }
    if (b.type == PPG_BSDF_THINDIELECTRIC) {                                                   // thindielectric.cpp:206-240
        const float R = thindielectric_reflectance(wi.z, b.eta[0]);
        s.delta = true; s.eta = 1.0f;
        if (sx <= R) { s.wo = f3(-wi.x, -wi.y, wi.z); pdf = R; return f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]); }
        s.null = true; s.wo = f3(-wi.x, -wi.y, -wi.z); pdf = 1 - R;
        return f3(b.specular_transmittance[0], b.specular_transmittance[1], b.specular_transmittance[2]);
    }
    if (b.type == PPG_BSDF_CONDUCTOR) {
        if (wi.z <= 0) return f3(0, 0, 0);
        s.delta = true; s.wo = f3(-wi.x, -wi.y, wi.z); pdf = 1;
        if (flip) s.wo.z = -s.wo.z;
        return f3(b.reflectance[0] * fresnel_conductor_exact(wi.z, b.eta[0], b.k[0]), b.reflectance[1] * fresnel_conductor_exact(wi.z, b.eta[1], b.k[1]),
                  b.reflectance[2] * fresnel_conductor_exact(wi.z, b.eta[2], b.k[2]));
    }
    if (b.type == PPG_BSDF_ROUGHCONDUCTOR) {
        const F3 w = roughconductor_sample(b, wi, sx, sy, s.wo, pdf);
        if (flip) s.wo.z = -s.wo.z;
        return w;
    }
    if (b.type == PPG_BSDF_ROUGHDIELECTRIC) return roughdielectric_sample(b, wi, sx, sy, rng ? rng->next1D() : 0.5f, s.wo, s.eta, pdf);
    if (b.type == PPG_BSDF_PLASTIC) {
        const F3 w = plastic_sample(b, wi, sx, sy, s.wo, s.delta, pdf);
        if (flip) s.wo.z = -s.wo.z;
        return w;
    }
    if (b.type == PPG_BSDF_ROUGHPLASTIC) {
        const F3 w = roughplastic_sample(b, bsdf_table(b, tables), wi, sx, sy, s.wo, pdf);
        if (flip) s.wo.z = -s.wo.z;
        return w;
    }
    if (wi.z <= 0) return f3(0, 0, 0);
    s.wo = square_to_cosine_hemisphere(sx, sy);
    pdf = kInvPi * s.wo.z;
    if (flip) s.wo.z = -s.wo.z;
    return f3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
}

// ---- mask (src/bsdfs/mask.cpp:113-220), the outermost wrapper: nested model scaled by the opacity, or a straight-through null transition
static inline F3 mask_opacity(const ppg_bsdf &b) { return f3(b.opacity[0], b.opacity[1], b.opacity[2]); }
static inline float mask_prob(const ppg_bsdf &b) { return b.opacity[0] * 0.212671f + b.opacity[1] * 0.715160f + b.opacity[2] * 0.072169f; }   // getLuminance, spectrum.h:725-727
static inline F3 bsdf_eval(const ppg_bsdf &b, F3 wi, F3 wo, const float *tables = nullptr) {
    if (b.flags & PPG_BSDF_FLAG_MASK) return bsdf_eval_inner(b, wi, wo, tables) * mask_opacity(b);
    return bsdf_eval_inner(b, wi, wo, tables);
}
static inline float bsdf_pdf(const ppg_bsdf &b, F3 wi, F3 wo, const float *tables = nullptr) {
    if (b.flags & PPG_BSDF_FLAG_MASK) return bsdf_pdf_inner(b, wi, wo, tables) * mask_prob(b);
    return bsdf_pdf_inner(b, wi, wo, tables);
}
// `rng`: the path's sampler, consumed only by models that draw from it themselves (roughdielectric)
static inline F3 bsdf_sample(const ppg_bsdf &b, F3 wi, float sx, float sy, BsdfSample &s, float &pdf, const float *tables = nullptr, Pcg32 *rng = nullptr) {
    if (b.flags & PPG_BSDF_FLAG_MASK) {                                                       // mask.cpp:186-207
        const F3 opacity = mask_opacity(b); const float prob = mask_prob(b);
        if (sx < prob) {
            sx /= prob;
            F3 result = (bsdf_sample_inner(b, wi, sx, sy, s, pdf, tables, rng) * opacity) * (1.0f / prob);
            pdf *= prob;
            return result;
        }
        s.wo = f3(-wi.x, -wi.y, -wi.z); s.eta = 1.0f; s.delta = true; s.null = true;
        pdf = 1 - prob;
        return (f3(1, 1, 1) - opacity) * (1.0f / pdf);
    }
    return bsdf_sample_inner(b, wi, sx, sy, s, pdf, tables, rng);
}

// ---- the BSDF at a hit: textured parameters looked up at its.uv (Texture2D::eval, librender/texture.cpp:112-121), and the bumpmap wrapper
// (src/bsdfs/bumpmap.cpp): BumpMap::getFrame :139-159, eval :161-175, pdf :177-193, sample :214-236
struct HitBsdf {
    ppg_bsdf b; bool bump = false; F3 bs, bt, bn;          // perturbed frame (world space)
    F3 toPerturbed(const Its &its, F3 wLocal) const { const F3 w = its.toWorld(wLocal); return f3(dot(w, bs), dot(w, bt), dot(w, bn)); }
    F3 fromPerturbed(const Its &its, F3 wP) const { return its.toLocal(bs * wP.x + bt * wP.y + bn * wP.z); }
};
// per-triangle UV tangents, TriMesh::computeUVTangents (src/librender/trimesh.cpp:683-743); meshes without texture coordinates: the two edges
static inline void uv_tangents(const Scene &sc, const Its &its, F3 &dpdu, F3 &dpdv) {
    const uint32_t i0 = sc.idx[3 * its.prim], i1 = sc.idx[3 * its.prim + 1], i2 = sc.idx[3 * its.prim + 2];
    const F3 dP1 = sc.P[i1] - sc.P[i0], dP2 = sc.P[i2] - sc.P[i0];
    if (!sc.shapes[its.shape].has_uvs) { dpdu = dP1; dpdv = dP2; return; }
    const float du1 = sc.UV[2 * i1] - sc.UV[2 * i0], dv1 = sc.UV[2 * i1 + 1] - sc.UV[2 * i0 + 1], du2 = sc.UV[2 * i2] - sc.UV[2 * i0], dv2 = sc.UV[2 * i2 + 1] - sc.UV[2 * i0 + 1];
    const F3 n = cross(dP1, dP2); const float len = length(n);
    if (len == 0) { dpdu = dpdv = f3(0, 0, 0); return; }
    const float determinant = du1 * dv2 - dv1 * du2;
    if (determinant == 0) { coordinate_system(n * (1.0f / len), dpdu, dpdv); return; }
    const float invDet = 1.0f / determinant;
    dpdu = (dP1 * dv2 - dP2 * dv1) * invDet;
    dpdv = (dP1 * (-du2) + dP2 * du1) * invDet;
}
static inline void resolve_bsdf(const Scene &sc, const Its &its, int index, HitBsdf &hb) {
    hb.b = sc.bsdfs[index]; hb.bump = false;
    if (hb.b.reflectance_texture) {
        const F3 c = sc.evalTexture(hb.b.reflectance_texture - 1, its.uvU, its.uvV);
        hb.b.reflectance[0] = c.x; hb.b.reflectance[1] = c.y; hb.b.reflectance[2] = c.z;
    }
    if ((hb.b.flags & PPG_BSDF_FLAG_BUMPMAP) && hb.b.bump_texture && !(its.prim & Scene::kSphereBit)) {
        float dDispDu, dDispDv; sc.evalTextureGradientLum(hb.b.bump_texture - 1, its.uvU, its.uvV, dDispDu, dDispDv);
        F3 dpdu0, dpdv0; uv_tangents(sc, its, dpdu0, dpdv0);
        const F3 dpdu = dpdu0 + its.shN * (dDispDu - dot(its.shN, dpdu0)), dpdv = dpdv0 + its.shN * (dDispDv - dot(its.shN, dpdv0));
        hb.bn = normalize(cross(dpdu, dpdv));
        hb.bs = normalize(dpdu - hb.bn * dot(hb.bn, dpdu));
        hb.bt = cross(hb.bn, hb.bs);
        if (dot(hb.bn, its.geoN) < 0) hb.bn = hb.bn * -1.0f;
        hb.bump = true;
    }
}
static inline F3 hit_eval(const HitBsdf &hb, const Its &its, F3 wi, F3 wo, const float *tables) {
    if (!hb.bump) return bsdf_eval(hb.b, wi, wo, tables);
    const F3 pwi = hb.toPerturbed(its, wi), pwo = hb.toPerturbed(its, wo);
    if (wo.z * pwo.z <= 0) return f3(0, 0, 0);
    return bsdf_eval(hb.b, pwi, pwo, tables);
}
static inline float hit_pdf(const HitBsdf &hb, const Its &its, F3 wi, F3 wo, const float *tables) {
    if (!hb.bump) return bsdf_pdf(hb.b, wi, wo, tables);
    const F3 pwi = hb.toPerturbed(its, wi), pwo = hb.toPerturbed(its, wo);
    if (wo.z * pwo.z <= 0) return 0.0f;
    return bsdf_pdf(hb.b, pwi, pwo, tables);
}
static inline F3 hit_sample(const HitBsdf &hb, const Its &its, F3 wi, float sx, float sy, BsdfSample &s, float &pdf, const float *tables, Pcg32 *rng) {
    if (!hb.bump) return bsdf_sample(hb.b, wi, sx, sy, s, pdf, tables, rng);
    F3 result = bsdf_sample(hb.b, hb.toPerturbed(its, wi), sx, sy, s, pdf, tables, rng);
    if (!is_zero(result)) {
        const F3 pwo = s.wo;
        s.wo = hb.fromPerturbed(its, pwo);
        if (s.wo.z * pwo.z <= 0) return f3(0, 0, 0);
    }
    return result;
}

// ------------------------------------------------------------------ the integrator
struct CommitRec {   // GP:1713-1724 Vertex, minus the tree pointer (kept as backend leaf handle)
    F3 o, d, voxel, throughput, bsdfVal, radiance; float woPdf, bsdfPdf, dTreePdf; bool isDelta;
};

// Vertex::commit, GP:1730-1768 (restated; the reference's own struct Vertex is compiled verbatim into oracle/_ref and compared with this in tests/test_oracle_sdtree.py):
// reject invalid records, radiance / throughput per channel where throughput * woPdf > Epsilon, product with the BSDF value, channel averages, then the spatial filter
// (Backend::record: nearest / stochastic with the three jitter numbers `rnd` / box)
template <class Backend>
static inline void commit_vertex(Backend &tree, typename Backend::Leaf *leaf, const CommitRec &v, float statisticalWeight, int sfilter, int dfilter, int loss, const float rnd[3]) {
    if (!(v.woPdf > 0) || !is_valid(v.radiance) || !is_valid(v.bsdfVal)) return;
    F3 local = f3(0, 0, 0);
    if (v.throughput.x * v.woPdf > kEpsilon) local.x = v.radiance.x / v.throughput.x;
    if (v.throughput.y * v.woPdf > kEpsilon) local.y = v.radiance.y / v.throughput.y;
    if (v.throughput.z * v.woPdf > kEpsilon) local.z = v.radiance.z / v.throughput.z;
    const F3 product = local * v.bsdfVal;
    const float avgLocal = (local.x + local.y + local.z) * (1.0f / 3.0f);      // Spectrum::average(): sum * (1/N)
    const float avgProduct = (product.x + product.y + product.z) * (1.0f / 3.0f);
    tree.record(leaf, &v.o.x, &v.voxel.x, &v.d.x, avgLocal, avgProduct, v.woPdf, v.bsdfPdf, v.dTreePdf, statisticalWeight, v.isDelta, sfilter, dfilter, loss, rnd);
}

template <class Backend> class Tracer {
public:
    ppg_params prm; Scene sc; Backend tree; int nthreads;
    bool isBuilt = false, isFinalIter = false, doNee = false; int iter = 0; int passesRendered = 0;
    int W, H;
    bool doNeeWithSpp(int spp) const { return prm.nee == PPG_NEE_NEVER ? false : (prm.nee == PPG_NEE_KICKSTART ? spp < 128 : true); }   // GP:1331-1340
    std::vector<float> image, sqImage, film;     // W*H*4 (r,g,b,weight)
    std::vector<std::vector<float>> images; std::vector<float> variances;
    uint64_t totalVertices = 0, totalPaths = 0;
    ppg_stats stats;
    // tile sharding for the 2-rank gloo test: this rank renders blocks with blk % world == rank and
    // sums [building sums | building weights] over ranks before every build, the variance numerator and the film
    int shardRank = 0, shardWorld = 1; ppg_allreduce_fn allreduce = nullptr; void *allreduceUser = nullptr;
    // optional per-path capture for parity tests
    std::vector<float> *captureLi = nullptr; std::vector<int32_t> *captureDepth = nullptr;

    Tracer(const ppg_params &p, const ppg_scene_desc &d, int threads)
        : prm(p), tree(d.aabb_min, d.aabb_max), nthreads(threads) {
        sc.load(d); W = d.camera.film_width; H = d.camera.film_height;
        image.assign((size_t) W * H * 4, 0.f); sqImage = image; film = image;
        std::memset(&stats, 0, sizeof(stats));
    }

    // src/sensors/perspective.cpp:271-298 written out for the lookAt camera (see DESIGN.md "camera")
    void sampleRay(float px, float py, F3 &o, F3 &d, float &mint, float &maxt) const {
        const float sx = px * (1.0f / (float) W), sy = py * (1.0f / (float) H);
        const F3 nearP = f3((1.0f - 2.0f * sx) * sc.tanX, (1.0f - 2.0f * sy) * sc.tanY, 1.0f);
        const F3 dl = normalize(nearP);
        const float invZ = 1.0f / dl.z;
        mint = sc.cam.near_clip * invZ; maxt = sc.cam.far_clip * invZ;
        o = sc.camO;
        d = sc.camLeft * dl.x + sc.camUp * dl.y + sc.camDir * dl.z;
    }

    // GP:1712-2157, surface branch (no media: README.md:5-7), nee = never
    F3 Li(F3 o, F3 d, float mint, float maxt, Pcg32 &rng, uint64_t sampleIndex, int &depthOut, uint64_t &nVerticesTraced) {
        typename Backend::Leaf *vLeaf[32]; CommitRec vtx[32]; int nVertices = 0;
        F3 LiAcc = f3(0, 0, 0), throughput = f3(1, 1, 1); float eta = 1.0f;
        bool scattered = false, emittedRadiance = true;   // rRec.type & EEmittedRadiance (ERadiance for sensor rays)
        int depth = 1;
        Its its; ray_intersect(sc, o, d, mint, maxt, its);
        nVerticesTraced++;
        auto recordRadiance = [&](F3 r) { LiAcc = LiAcc + r; for (int i = 0; i < nVertices; ++i) vtx[i].radiance = vtx[i].radiance + r; };   // GP:1791-1796
        while (depth <= prm.max_depth || prm.max_depth < 0) {
            if (!its.valid) {                                                                    // GP:1902-1914: attenuated radiance of the environment emitter
                if (sc.hasEnvironment() && emittedRadiance && (!prm.hide_emitters || scattered)) recordRadiance(throughput * sc.evalEnvironment(d));
                break;
            }
            const ppg_shape &shp = sc.shapes[its.shape];
            if (shp.emitter >= 0 && emittedRadiance && (!prm.hide_emitters || scattered)) {   // GP:1917-1919; area.cpp:104-109
                if (dot(its.shN, -d) > 0) recordRadiance(throughput * sc.radiance[shp.emitter]);
            }
            if (depth >= prm.max_depth && prm.max_depth != -1) break;                          // GP:1925
            const float wiDotGeoN = -dot(its.geoN, d), wiDotShN = its.wi.z;
            if (wiDotGeoN * wiDotShN < 0 && prm.strict_normals) break;                         // GP:1929-1932
            HitBsdf hb; resolve_bsdf(sc, its, shp.bsdf, hb);                                    // its.getBSDF() (GP:1934): no ray differentials
            const ppg_bsdf &bsdf = hb.b;
            float voxel[3] = {0, 0, 0}; typename Backend::Leaf *leaf = nullptr;
            if (bsdf_has_smooth(bsdf)) leaf = tree.lookup(&its.p.x, voxel);                    // GP:1942-1944: only smooth BSDFs are guided
            float frac = prm.bsdf_sampling_fraction;
            if (leaf && prm.bsdf_sampling_fraction_loss != PPG_LOSS_NONE) frac = tree.bsdfSamplingFraction(leaf);   // GP:1946-1949
            // ---- sampleMat, GP:1650-1691
            float woPdf, bsdfPdf, dTreePdf; F3 bsdfWeight; BsdfSample bs;
            float sx = rng.next1D(), sy = rng.next1D();
            if (!isBuilt || !leaf) {
                bsdfWeight = hit_sample(hb, its, its.wi, sx, sy, bs, bsdfPdf, sc.tables.data(), &rng);
                woPdf = bsdfPdf; dTreePdf = 0;
            } else {
                F3 result;
                bool zero = false, deltaEarly = false;
                if (sx < frac) {
                    sx /= frac;
                    result = hit_sample(hb, its, its.wi, sx, sy, bs, bsdfPdf, sc.tables.data(), &rng);
                    if (is_zero(result)) { woPdf = bsdfPdf = dTreePdf = 0; zero = true; }
                    else if (bs.delta) { dTreePdf = 0; woPdf = bsdfPdf * frac; result = result * (1.0f / frac); deltaEarly = true; }   // GP:1670-1676
                    else result = result * bsdfPdf;
                } else {
                    float dw[3]; tree.sample(leaf, rng, dw);
                    bs.wo = its.toLocal(f3(dw[0], dw[1], dw[2])); bs.eta = 1.0f; bs.delta = false;
                    result = hit_eval(hb, its, its.wi, bs.wo, sc.tables.data());
                }
                if (zero) bsdfWeight = f3(0, 0, 0);
                else if (deltaEarly) bsdfWeight = result;
                else {
                    // pdfMat, GP:1693-1710
                    bsdfPdf = hit_pdf(hb, its, its.wi, bs.wo, sc.tables.data());
                    if (!std::isfinite(bsdfPdf)) { woPdf = 0; dTreePdf = 0; }
                    else {
                        const F3 wow = its.toWorld(bs.wo);
                        dTreePdf = tree.pdf(leaf, &wow.x);
                        woPdf = frac * bsdfPdf + (1 - frac) * dTreePdf;
                    }
                    bsdfWeight = (woPdf == 0) ? f3(0, 0, 0) : result * (1.0f / woPdf);   // Spectrum::operator/(Float) multiplies by the reciprocal
                }
            }
            // ---- luminaire sampling, GP:1964-2021 (DirectSamplingRecord dRec(its): refN = shading normal unless the BSDF is two-sided / transmissive, records.inl:160-164)
            const F3 refN = bsdf_has_transmission_or_backside(bsdf) ? f3(0, 0, 0) : its.shN;
            if (doNee && bsdf_has_smooth(bsdf)) {                                               // GP:1967-1969
                const float ex = rng.next1D(), ey = rng.next1D();
                DirectSample ds;
                if (sample_emitter_direct(sc, its.p, refN, ex, ey, ds, prm.max_depth - depth - 1) && !is_zero(ds.value)) {     // interactions, GP:1970
                    const F3 dl = its.toLocal(ds.d);
                    const float woDotGeoN2 = dot(its.geoN, ds.d);
                    if (!prm.strict_normals || woDotGeoN2 * dl.z > 0) {
                        const F3 bsdfVal = hit_eval(hb, its, its.wi, dl, sc.tables.data());
                        float nWoPdf = 0, nBsdfPdf = 0, nDTreePdf = 0;
                        {   // pdfMat (emitter->isOnSurface() && measure == ESolidAngle always hold for area lights)
                            if (!isBuilt || !leaf) { nWoPdf = nBsdfPdf = hit_pdf(hb, its, its.wi, dl, sc.tables.data()); }
                            else {
                                nBsdfPdf = hit_pdf(hb, its, its.wi, dl, sc.tables.data());
                                if (!std::isfinite(nBsdfPdf)) nWoPdf = 0;
                                else { nDTreePdf = tree.pdf(leaf, &ds.d.x); nWoPdf = frac * nBsdfPdf + (1 - frac) * nDTreePdf; }
                            }
                        }
                        const float weight = mi_weight(ds.pdf, nWoPdf);
                        const F3 value = ds.value * bsdfVal;
                        const F3 L = throughput * value * weight;
                        if (!isFinalIter && prm.nee != PPG_NEE_ALWAYS && leaf) {               // GP:1999-2016: half-weight vertex for the sampled light direction
                            CommitRec v; v.o = its.p; v.d = ds.d; v.voxel = f3(voxel[0], voxel[1], voxel[2]);
                            v.throughput = throughput * bsdfVal * (1.0f / ds.pdf); v.bsdfVal = bsdfVal; v.radiance = L;
                            v.woPdf = ds.pdf; v.bsdfPdf = nBsdfPdf; v.dTreePdf = nDTreePdf; v.isDelta = false;
                            commit(leaf, v, 0.5f, isBuilt ? prm.bsdf_sampling_fraction_loss : PPG_LOSS_NONE, sampleIndex, 32u + (uint32_t) depth);
                        }
                        recordRadiance(L);
                    }
                }
            }
            if (is_zero(bsdfWeight)) break;                                                     // GP:2024-2025
            const F3 wo = its.toWorld(bs.wo);
            const float woDotGeoN = dot(its.geoN, wo);
            if (woDotGeoN * bs.wo.z <= 0 && prm.strict_normals) break;                          // GP:2028-2032
            o = its.p; d = wo;
            throughput = throughput * bsdfWeight; eta *= bs.eta;
            if (bs.null) {                                                                      // index-matched transition, GP:2044-2075
                // smooth/null hybrids (mask): the null transition is recorded as a delta vertex for the sampling-fraction optimisation
                if (prm.bsdf_sampling_fraction_loss != PPG_LOSS_NONE && leaf && nVertices < 32 && !isFinalIter) {   // GP:2049-2066
                    if (1 / woPdf > 0) {
                        CommitRec &v = vtx[nVertices]; vLeaf[nVertices] = leaf;
                        v.o = o; v.d = d; v.voxel = f3(voxel[0], voxel[1], voxel[2]); v.throughput = throughput;
                        v.bsdfVal = bsdfWeight * woPdf; v.radiance = f3(0, 0, 0);
                        v.woPdf = woPdf; v.bsdfPdf = bsdfPdf; v.dTreePdf = dTreePdf; v.isDelta = true;
                        ++nVertices;
                    }
                }
                emittedRadiance = !scattered;                                                   // ERadiance : ERadianceNoEmission
                ray_intersect(sc, o, d, kEpsilon, std::numeric_limits<float>::infinity(), its);
                nVerticesTraced++;
                depth++;
                continue;                                                                       // no Russian roulette, `scattered` unchanged
            }
            // ---- next hit + emitter lookup, GP:2078-2091 -> rayIntersectAndLookForEmitter GP:2184-2245: the emitter is looked for THROUGH
            // index-matched surfaces (at most maxDepth - depth - 1 of them), the path itself continues at the first hit
            F3 value = f3(0, 0, 0);
            Its next; ray_intersect(sc, o, d, kEpsilon, std::numeric_limits<float>::infinity(), next);
            nVerticesTraced++;
            int qEmitter = -1; F3 qN = f3(0, 0, 0); float qDist = 0;                            // dRec.setQuery(ray, *its), records.inl:170-178
            {
                const Its *cur = &next; Its its2; F3 ro = o, transmittance = f3(1, 1, 1);
                int interactions = 0; const int maxInteractions = prm.max_depth - depth - 1;
                bool surface, lost = false;
                while (true) {
                    surface = cur->valid;
                    if (surface && (interactions == maxInteractions || !bsdf_has_null(sc.bsdfs[sc.shapes[cur->shape].bsdf]) || sc.shapes[cur->shape].emitter >= 0)) break;
                    if (!surface) break;
                    if (is_zero(transmittance)) { lost = true; break; }
                    const F3 wol = cur->toLocal(d);
                    transmittance = transmittance * bsdf_eval_null(sc.bsdfs[sc.shapes[cur->shape].bsdf], -wol.z);   // bRec(its, -wo, wo), typeMask ENull, EDiscrete
                    ro = ro + d * cur->t;
                    ray_intersect(sc, ro, d, kEpsilon, std::numeric_limits<float>::infinity(), its2); cur = &its2;
                    if (++interactions > 100) { lost = true; break; }
                }
                if (!lost && surface) {
                    const ppg_shape &ns = sc.shapes[cur->shape];
                    if (ns.emitter >= 0) {
                        qEmitter = ns.emitter; qN = cur->shN; qDist = cur->t;                   // dist: the last segment only (quirk of setQuery after ray.o moved)
                        if (dot(cur->shN, -d) > 0) value = transmittance * sc.radiance[ns.emitter];
                    }
                } else if (!lost && sc.hasEnvironment()) {
                    // "Intersected nothing -- perhaps there is an environment map?" (GP:2228-2243); fillDirectSamplingRecord succeeds for any ray
                    // that starts inside the scene's bounding sphere (envmap.cpp:360-378), which every surface point does (radius x 1.5)
                    value = transmittance * sc.evalEnvironment(d);
                    qEmitter = kEnvEmitter;
                }
            }
            const bool isDelta = bs.delta;
            {
                float emitterPdf = 0;                                                            // GP:2084-2087
                if (doNee && !isDelta && !is_zero(value)) emitterPdf = pdf_emitter_direct(sc, qEmitter, its.p, refN, d, qN, qDist);
                const float weight = mi_weight(woPdf, emitterPdf);
                const F3 L = throughput * value * weight;
                if (!is_zero(L)) recordRadiance(L);
                if ((!isDelta || prm.bsdf_sampling_fraction_loss != PPG_LOSS_NONE) && leaf && nVertices < 32 && !isFinalIter) {   // GP:2093-2110
                    if (1 / woPdf > 0) {
                        CommitRec &v = vtx[nVertices]; vLeaf[nVertices] = leaf;
                        v.o = o; v.d = d; v.voxel = f3(voxel[0], voxel[1], voxel[2]); v.throughput = throughput;
                        v.bsdfVal = bsdfWeight * woPdf; v.radiance = (prm.nee == PPG_NEE_ALWAYS) ? f3(0, 0, 0) : L;
                        v.woPdf = woPdf; v.bsdfPdf = bsdfPdf; v.dTreePdf = dTreePdf; v.isDelta = isDelta;
                        ++nVertices;
                    }
                }
            }
            its = next;
            emittedRadiance = false;                                                            // GP:2121 ERadianceNoEmission
            if (depth++ >= prm.rr_depth) {                                                      // GP:2123-2142
                float successProb = 1.0f;
                if (leaf && !isDelta) {
                    if (!isBuilt) successProb = max3(throughput) * eta * eta;
                    successProb = std::max(0.1f, std::min(successProb, 0.99f));
                }
                if (rng.next1D() >= successProb) break;
                throughput = throughput * (1.0f / successProb);
            }
            scattered = true;
        }
        depthOut = depth;
        if (nVertices > 0 && !isFinalIter) {                                                    // GP:2150-2154
            const int loss = isBuilt ? prm.bsdf_sampling_fraction_loss : PPG_LOSS_NONE;
            const float w = (prm.nee == PPG_NEE_KICKSTART && doNee) ? 0.5f : 1.0f;               // GP:2152
            for (int i = 0; i < nVertices; ++i) commit(vLeaf[i], vtx[i], w, loss, sampleIndex, (uint32_t) i);
        }
        return LiAcc;
    }

    // Vertex::commit, GP:1730-1768
    void commit(typename Backend::Leaf *leaf, const CommitRec &v, float statisticalWeight, int loss, uint64_t sampleIndex, uint32_t ordinal) {
        if (!(v.woPdf > 0) || !is_valid(v.radiance) || !is_valid(v.bsdfVal)) return;        // (before the jitter numbers are drawn, like the reference's early return)
        float rnd[3] = {0, 0, 0};
        if (prm.spatial_filter == PPG_SFILTER_STOCHASTIC) {
            Pcg32 r; seed_vertex_rng(r, prm.seed, sampleIndex, ordinal);
            rnd[0] = r.next1D(); rnd[1] = r.next1D(); rnd[2] = r.next1D();
        }
        commit_vertex(tree, leaf, v, statisticalWeight, prm.spatial_filter, prm.directional_filter, loss, rnd);
    }

    // renderBlock (GP:1587-1641) over every 32x32 block of one pass, OpenMP over blocks like the LocalWorkers
    void renderPass(int passGlobal) {
        const int bs = 32, bx = (W + bs - 1) / bs, by = (H + bs - 1) / bs;
        uint64_t verts = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : verts)
        for (int blk = 0; blk < bx * by; ++blk) {
            if (blk % shardWorld != shardRank) continue;
            const int x0 = (blk % bx) * bs, y0 = (blk / bx) * bs;
            for (int y = y0; y < std::min(y0 + bs, H); ++y)
                for (int x = x0; x < std::min(x0 + bs, W); ++x)
                    for (int j = 0; j < prm.spp_per_pass; ++j) {
                        const uint64_t sampleIndex = (((uint64_t) passGlobal * H + y) * W + x) * prm.spp_per_pass + j;
                        Pcg32 rng; seed_path_rng(rng, prm.seed, sampleIndex);
                        const float jx = rng.next1D(), jy = rng.next1D();
                        F3 o, d; float mint, maxt;
                        sampleRay((float) x + jx, (float) y + jy, o, d, mint, maxt);
                        int depth; uint64_t nv = 0;
                        const F3 spec = Li(o, d, mint, maxt, rng, sampleIndex, depth, nv);
                        verts += nv;
                        if (captureLi) {
                            const size_t li = ((size_t) y * W + x) * prm.spp_per_pass + j;
                            (*captureLi)[3 * li] = spec.x; (*captureLi)[3 * li + 1] = spec.y; (*captureLi)[3 * li + 2] = spec.z;
                            if (captureDepth) (*captureDepth)[li] = depth;
                        }
                        if (!is_valid(spec)) continue;                    // ImageBlock::put rejects invalid samples (imageblock.h:150-154)
                        float *px = &image[((size_t) y * W + x) * 4], *sq = &sqImage[((size_t) y * W + x) * 4];
                        px[0] += spec.x; px[1] += spec.y; px[2] += spec.z; px[3] += 1.0f;
                        sq[0] += spec.x * spec.x; sq[1] += spec.y * spec.y; sq[2] += spec.z * spec.z; sq[3] += 1.0f;
                    }
        }
        totalVertices += verts;
        uint64_t px = 0;
        for (int blk = 0; blk < bx * by; ++blk) if (blk % shardWorld == shardRank) {
            const int x0 = (blk % bx) * bs, y0 = (blk / bx) * bs;
            px += (uint64_t) (std::min(x0 + bs, W) - x0) * (std::min(y0 + bs, H) - y0);
        }
        totalPaths += px * prm.spp_per_pass;
    }

    // performRenderPasses, GP:1210-1329 (scheduling elided)
    bool performRenderPasses(float &variance, int numPasses, ppg_iteration_stats &st) {
        std::fill(image.begin(), image.end(), 0.f); std::fill(sqImage.begin(), sqImage.end(), 0.f);
        const auto t0 = std::chrono::steady_clock::now();
        const uint64_t v0 = totalVertices, p0 = totalPaths;
        int local = 0;
        for (int i = 0; i < numPasses; ++i) {
            renderPass(passesRendered);
            ++passesRendered; ++local;
            if (prm.budget_type == PPG_BUDGET_SECONDS && elapsed(startTime) > prm.budget) break;   // GP:1259-1262
        }
        for (size_t i = 0; i < film.size(); ++i) film[i] += image[i];                             // film->put(block), renderproc.cpp:143-151
        if (prm.sample_combination == PPG_COMB_INVERSEVAR) images.push_back(image);              // GP:1292-1296
        // variance estimate with the getPixel() quirk (GP:1300-1313)
        const int N = local * prm.spp_per_pass;
        variance = 0;
        for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y) {
            if (shardWorld > 1 && ((y / 32) * ((W + 31) / 32) + x / 32) % shardWorld != shardRank) continue;
            const float *px = &image[((size_t) y * W + x) * 4], *sq = &sqImage[((size_t) y * W + x) * 4];
            const float iw = px[3] != 0 ? 1.0f / px[3] : 0.0f, isw = sq[3] != 0 ? 1.0f / sq[3] : 0.0f;
            float lv[3];
            for (int c = 0; c < 3; ++c) { const float pix = px[c] * iw; lv[c] = sq[c] * isw - pix * pix / (float) N; }
            const float lum = lv[0] * 0.212671f + lv[1] * 0.715160f + lv[2] * 0.072169f;
            variance += std::min(lum, 10000.0f);
        }
        if (allreduce && shardWorld > 1) allreduce(allreduceUser, &variance, 1);
        variance /= (float) W * H * (N - 1);
        if (prm.sample_combination == PPG_COMB_INVERSEVAR) variances.push_back(variance);
        st.seconds += elapsed(t0); st.passes += local; st.variance = variance; st.total_passes = passesRendered;
        st.vertices += totalVertices - v0; st.paths += totalPaths - p0;
        return true;
    }

    std::chrono::steady_clock::time_point startTime;
    static float elapsed(std::chrono::steady_clock::time_point s) {
        return (float) std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - s).count() / 1000;
    }

    void resetSDTree() {   // GP:1108-1113
        tree.refine((size_t) (std::sqrt(std::pow(2, iter) * prm.spp_per_pass / 4) * prm.s_tree_threshold), prm.sd_tree_max_memory);
        tree.resetAll(20, prm.d_tree_threshold, nthreads);
    }
    void buildSDTree(ppg_iteration_stats &st) {   // GP:1115-1189
        if (allreduce && shardWorld > 1) {
            std::vector<float> buf(tree.packSize());
            tree.packBuilding(buf.data(), false);
            allreduce(allreduceUser, buf.data(), buf.size());
            tree.packBuilding(buf.data(), true);
        }
        tree.buildAll(nthreads);
        tree.statistics(st);
        isBuilt = true;
    }

    void beginIterStats(ppg_iteration_stats &st, int passes) { std::memset(&st, 0, sizeof(st)); st.iteration = iter; (void) passes; }

    // GP:1342-1426
    bool renderSPP() {
        const int nPasses = (int) std::ceil((size_t) prm.budget / (float) prm.spp_per_pass);
        bool result = true; float currentVarAtEnd = std::numeric_limits<float>::infinity();
        while (result && passesRendered < nPasses) {
            const int sppRendered = passesRendered * prm.spp_per_pass;
            doNee = doNeeWithSpp(sppRendered);                                                   // GP:1362
            int remainingPasses = nPasses - passesRendered;
            int passesThisIteration = std::min(remainingPasses, 1 << iter);
            if (remainingPasses - passesThisIteration < 2 * passesThisIteration) passesThisIteration = remainingPasses;
            isFinalIter = passesThisIteration >= remainingPasses;
            ppg_iteration_stats &st = stats.iterations[std::min(iter, PPG_MAX_ITERATIONS - 1)]; beginIterStats(st, passesThisIteration);
            std::fill(film.begin(), film.end(), 0.f);
            auto t0 = std::chrono::steady_clock::now(); resetSDTree(); st.reset_seconds = elapsed(t0);
            float variance;
            if (!performRenderPasses(variance, passesThisIteration, st)) { result = false; break; }
            const float lastVarAtEnd = currentVarAtEnd;
            currentVarAtEnd = passesThisIteration * variance / remainingPasses;
            remainingPasses -= passesThisIteration;
            if (prm.sample_combination == PPG_COMB_AUTOMATIC && remainingPasses > 0 &&
                (remainingPasses < passesThisIteration || (sppRendered > 256 && currentVarAtEnd > lastVarAtEnd))) {
                isFinalIter = true;
                if (!performRenderPasses(variance, remainingPasses, st)) { result = false; break; }
            }
            st.is_final = isFinalIter;
            t0 = std::chrono::steady_clock::now(); buildSDTree(st); st.build_seconds = elapsed(t0);
            ++iter; stats.n_iterations = std::min(iter, PPG_MAX_ITERATIONS);
        }
        return result;
    }

    // GP:1434-1514
    bool renderTime() {
        const float nSeconds = prm.budget; bool result = true;
        float currentVarAtEnd = std::numeric_limits<float>::infinity(), elapsedSeconds = 0;
        while (result && elapsedSeconds < nSeconds) {
            const int sppRendered = passesRendered * prm.spp_per_pass;
            doNee = doNeeWithSpp(sppRendered);                                                   // GP:1452
            float remainingTime = nSeconds - elapsedSeconds;
            const int passesThisIteration = 1 << iter;
            ppg_iteration_stats &st = stats.iterations[std::min(iter, PPG_MAX_ITERATIONS - 1)]; beginIterStats(st, passesThisIteration);
            const auto startIter = std::chrono::steady_clock::now();
            std::fill(film.begin(), film.end(), 0.f);
            resetSDTree(); st.reset_seconds = elapsed(startIter);
            float variance;
            if (!performRenderPasses(variance, passesThisIteration, st)) { result = false; break; }
            const float secondsIter = elapsed(startIter);
            const float lastVarAtEnd = currentVarAtEnd;
            currentVarAtEnd = secondsIter * variance / remainingTime;
            remainingTime -= secondsIter;
            if (prm.sample_combination == PPG_COMB_AUTOMATIC && remainingTime > 0 &&
                (remainingTime < secondsIter || (sppRendered > 256 && currentVarAtEnd > lastVarAtEnd))) {
                isFinalIter = true;
                do {
                    if (!performRenderPasses(variance, passesThisIteration, st)) { result = false; break; }
                    elapsedSeconds = elapsed(startTime);
                } while (elapsedSeconds < nSeconds);
            }
            st.is_final = isFinalIter;
            auto t0 = std::chrono::steady_clock::now(); buildSDTree(st); st.build_seconds = elapsed(t0);
            ++iter; stats.n_iterations = std::min(iter, PPG_MAX_ITERATIONS);
            elapsedSeconds = elapsed(startTime);
        }
        return result;
    }

    // GP:1516-1585; film develop = weight-normalised RGB (hdrfilm)
    bool render(float *rgbOut) {
        iter = 0; isFinalIter = false; passesRendered = 0; images.clear(); variances.clear();
        startTime = std::chrono::steady_clock::now();
        const bool ok = prm.budget_type == PPG_BUDGET_SPP ? renderSPP() : renderTime();
        std::vector<float> out((size_t) W * H * 3, 0.f);
        if (prm.sample_combination == PPG_COMB_INVERSEVAR) {   // GP:1567-1582
            const size_t begin = images.size() - std::min(images.size(), (size_t) 4);
            float totalWeight = 0;
            for (size_t i = begin; i < variances.size(); ++i) totalWeight += 1.0f / variances[i];
            for (size_t i = begin; i < images.size(); ++i) {
                const float wgt = 1.0f / variances[i] / totalWeight;
                for (size_t p = 0; p < (size_t) W * H; ++p) {
                    const float *px = &images[i][p * 4]; const float iw = px[3] != 0 ? 1.0f / px[3] : 0.0f;
                    for (int c = 0; c < 3; ++c) out[p * 3 + c] += px[c] * iw * wgt;
                }
            }
        } else {
            for (size_t p = 0; p < (size_t) W * H; ++p) {
                const float *px = &film[p * 4]; const float iw = px[3] != 0 ? 1.0f / px[3] : 0.0f;
                for (int c = 0; c < 3; ++c) out[p * 3 + c] = px[c] * iw;
            }
        }
        if (allreduce && shardWorld > 1) allreduce(allreduceUser, out.data(), out.size());
        if (rgbOut) std::memcpy(rgbOut, out.data(), out.size() * sizeof(float));
        stats.total_passes = passesRendered; stats.total_paths = totalPaths; stats.total_vertices = totalVertices;
        stats.render_seconds = elapsed(startTime);
        stats.final_variance = stats.n_iterations ? stats.iterations[stats.n_iterations - 1].variance : 0;
        return ok;
    }
};

}  // namespace ppgo
