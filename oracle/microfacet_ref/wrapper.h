// TEST INFRASTRUCTURE -- C entry points over the reference's MicrofacetDistribution compiled verbatim (see oracle/Makefile, target mfref).
// Expects shim.h, math.cpp:25-86 (inside namespace mitsuba::math) and microfacet.h:45-721 (inside namespace mitsuba) to precede it in the translation unit.
#include <cstddef>
extern "C" {
using mitsuba::MicrofacetDistribution; using mitsuba::Vector; using mitsuba::Normal; using mitsuba::Point2;
static inline MicrofacetDistribution mk(int type, float alpha) { return MicrofacetDistribution(type == 1 ? MicrofacetDistribution::EGGX : MicrofacetDistribution::EBeckmann, alpha, true); }
int mfref_eval(int type, float alpha, size_t n, const float *m, float *out) {
    const MicrofacetDistribution d = mk(type, alpha);
    for (size_t i = 0; i < n; ++i) out[i] = d.eval(Vector(m[3 * i], m[3 * i + 1], m[3 * i + 2]));
    return 0;
}
int mfref_smith_g1(int type, float alpha, size_t n, const float *v, const float *m, float *out) {
    const MicrofacetDistribution d = mk(type, alpha);
    for (size_t i = 0; i < n; ++i) out[i] = d.smithG1(Vector(v[3 * i], v[3 * i + 1], v[3 * i + 2]), Vector(m[3 * i], m[3 * i + 1], m[3 * i + 2]));
    return 0;
}
int mfref_pdf(int type, float alpha, size_t n, const float *wi, const float *m, float *out) {
    const MicrofacetDistribution d = mk(type, alpha);
    for (size_t i = 0; i < n; ++i) out[i] = d.pdf(Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), Vector(m[3 * i], m[3 * i + 1], m[3 * i + 2]));
    return 0;
}
int mfref_sample(int type, float alpha, size_t n, const float *wi, const float *sample, float *m_out, float *pdf_out) {
    const MicrofacetDistribution d = mk(type, alpha);
    for (size_t i = 0; i < n; ++i) {
        mitsuba::Float pdf = 0;
        const Normal m = d.sample(Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), Point2(sample[2 * i], sample[2 * i + 1]), pdf);
        m_out[3 * i] = m.x; m_out[3 * i + 1] = m.y; m_out[3 * i + 2] = m.z; pdf_out[i] = pdf;
    }
    return 0;
}
// ---- helpers of src/libcore/util.cpp and warp.cpp compiled verbatim next to the class
int mfref_fresnel_dielectric_ext(size_t n, const float *cosThetaI, float eta, float *f_out, float *cos_t_out) {
    for (size_t i = 0; i < n; ++i) { mitsuba::Float ct = 0; f_out[i] = mitsuba::fresnelDielectricExt(cosThetaI[i], ct, eta); cos_t_out[i] = ct; }
    return 0;
}
int mfref_fresnel_conductor_exact(size_t n, const float *cosThetaI, const float eta[3], const float k[3], float *out /* 3n: the Spectrum overload the conductor BSDFs call */) {
    mitsuba::Spectrum e, kk; for (int c = 0; c < 3; ++c) { e.s[c] = eta[c]; kk.s[c] = k[c]; }
    for (size_t i = 0; i < n; ++i) { const mitsuba::Spectrum f = mitsuba::fresnelConductorExact(cosThetaI[i], e, kk); out[3 * i] = f.s[0]; out[3 * i + 1] = f.s[1]; out[3 * i + 2] = f.s[2]; }
    return 0;
}
int mfref_coordinate_system(size_t n, const float *a, float *b_out, float *c_out) {
    for (size_t i = 0; i < n; ++i) {
        Vector b, c; mitsuba::coordinateSystem(Vector(a[3 * i], a[3 * i + 1], a[3 * i + 2]), b, c);
        b_out[3 * i] = b.x; b_out[3 * i + 1] = b.y; b_out[3 * i + 2] = b.z; c_out[3 * i] = c.x; c_out[3 * i + 1] = c.y; c_out[3 * i + 2] = c.z;
    }
    return 0;
}
int mfref_square_to_cosine_hemisphere(size_t n, const float *sample, float *out) {
    for (size_t i = 0; i < n; ++i) { const Vector v = mitsuba::warp::squareToCosineHemisphere(Point2(sample[2 * i], sample[2 * i + 1])); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
    return 0;
}
// ---- TriAccel (include/mitsuba/render/triaccel.h:27-158) and evalCubicInterp1D (src/libcore/spline.cpp:23-60)
int mfref_triaccel(size_t n, const float *A, const float *B, const float *C, const float *o, const float *d, const float *mint, const float *maxt,
                   int *k_out, float *consts_out /* 9n: n_u n_v n_d a_u a_v b_nu b_nv c_nu c_nv */, unsigned char *hit_out, float *tuv_out /* 3n */) {
    for (size_t i = 0; i < n; ++i) {
        mitsuba::TriAccel t; t.n_u = t.n_v = t.n_d = t.a_u = t.a_v = t.b_nu = t.b_nv = t.c_nu = t.c_nv = 0;
        t.load(mitsuba::Point(A[3 * i], A[3 * i + 1], A[3 * i + 2]), mitsuba::Point(B[3 * i], B[3 * i + 1], B[3 * i + 2]), mitsuba::Point(C[3 * i], C[3 * i + 1], C[3 * i + 2]));
        k_out[i] = (int) t.k;
        const float c9[9] = {t.n_u, t.n_v, t.n_d, t.a_u, t.a_v, t.b_nu, t.b_nv, t.c_nu, t.c_nv};
        for (int j = 0; j < 9; ++j) consts_out[9 * i + j] = c9[j];
        mitsuba::Ray r; r.o = mitsuba::Point(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        mitsuba::Float u = 0, v = 0, tt = 0;
        hit_out[i] = t.rayIntersect(r, mint[i], maxt[i], u, v, tt) ? 1 : 0;
        tuv_out[3 * i] = tt; tuv_out[3 * i + 1] = u; tuv_out[3 * i + 2] = v;
    }
    return 0;
}
int mfref_cubic_interp_1d(size_t n, const float *x, const float *values, size_t size, float mn, float mx, float *out) {
    for (size_t i = 0; i < n; ++i) out[i] = mitsuba::evalCubicInterp1D(x[i], values, size, mn, mx, false);
    return 0;
}
// ---- DiscreteDistribution (include/mitsuba/core/pmf.h:35-210): append / normalize / sample / sampleReuse -- the emitter choice and the per-emitter triangle choice
int mfref_discrete(size_t n_entries, const float *weights, size_t n, const float *sample, float *pdf_out /* n_entries: the normalised entries */, float *sum_out,
                   unsigned *index_out, float *reused_out) {
    mitsuba::DiscreteDistribution d(n_entries);
    for (size_t i = 0; i < n_entries; ++i) d.append(weights[i]);
    *sum_out = d.normalize();
    for (size_t i = 0; i < n_entries; ++i) pdf_out[i] = d[i];
    for (size_t i = 0; i < n; ++i) { mitsuba::Float s = sample[i]; index_out[i] = (unsigned) d.sampleReuse(s); reused_out[i] = s; }
    return 0;
}
int mfref_erf(size_t n, const float *x, float *erf_out, float *erfinv_out) {
    for (size_t i = 0; i < n; ++i) { erf_out[i] = mitsuba::math::erf(x[i]); erfinv_out[i] = mitsuba::math::erfinv(x[i]); }
    return 0;
}
}
