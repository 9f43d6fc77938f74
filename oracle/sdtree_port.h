// TEST INFRASTRUCTURE -- CPU oracle, never shipped, never on the product path.
//
// sdtree_port.h: plain-C++ restatement of the reference's SD-tree
// (mitsuba/src/integrators/path/guided_path.cpp, "GP"): spatial binary tree whose
// leaves hold a building and a sampling directional quadtree plus an Adam-optimised
// BSDF sampling fraction.  Index-based flat arrays and iterative traversals instead
// of the reference's recursive object graph; every function cites the GP lines it
// restates.  Pinned against the verbatim-compiled reference code (oracle/_ref,
// tests/test_oracle_sdtree.py) and the known-answer statistics of the golden EXR logs.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <vector>

namespace ppgo {

static const float kPi = 3.14159265358979323846f;   // M_PI as Float (core/constants.h)
static const float kInv4Pi = 1.0f / (4.0f * kPi);

struct V2 { float x, y; };
struct V3 { float x, y, z; float &operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };

// atomic float add; the reference uses a CAS loop (GP:59-62)
static inline void atomic_add(float &dst, float v) {
#pragma omp atomic
    dst += v;
}

// ---------------------------------------------------------------- direction <-> square
// GP:586-595
static inline V3 canonical_to_dir(V2 p) {
    const float cosTheta = 2 * p.x - 1;
    const float phi = 2 * kPi * p.y;
    const float sinTheta = std::sqrt(1 - cosTheta * cosTheta);
    return V3{sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta};
}
// GP:597-608
static inline V2 dir_to_canonical(const V3 &d) {
    if (!std::isfinite(d.x) || !std::isfinite(d.y) || !std::isfinite(d.z)) return V2{0, 0};
    const float cosTheta = std::min(std::max(d.z, -1.0f), 1.0f);
    float phi = std::atan2(d.y, d.x);
    while (phi < 0) phi += 2.0 * (double) kPi;   // M_PI is a float literal in the single-precision build (core/constants.h:63,80); the add happens in double
    return V2{(cosTheta + 1) / 2, phi / (2 * kPi)};
}

// ---------------------------------------------------------------- Adam (GP:69-133)
struct Adam {
    int iter = 0;
    float m1 = 0, m2 = 0, variable = 0, batchAcc = 0, batchGrad = 0;
    static constexpr float lr = 0.01f, eps = 1e-08f, beta1 = 0.9f, beta2 = 0.999f;
    static constexpr int batchSize = 1;
    void step(float g) {                       // GP:97-109
        ++iter;
        float actualLr = lr * std::sqrt(1 - std::pow(beta2, iter)) / (1 - std::pow(beta1, iter));
        m1 = beta1 * m1 + (1 - beta1) * g;
        m2 = beta2 * m2 + (1 - beta2) * g * g;
        variable -= actualLr * m1 / (std::sqrt(m2) + eps);
        variable = std::min(std::max(variable, -20.0f), 20.0f);
    }
    void append(float g, float w) {            // GP:85-95
        batchGrad += g * w;
        batchAcc += w;
        if (batchAcc > batchSize) {
            step(batchGrad / batchAcc);
            batchGrad = 0; batchAcc = 0;
        }
    }
};

static inline float logistic(float x) { return 1 / (1 + std::exp(-x)); }   // GP:64-66

// ---------------------------------------------------------------- D-tree
struct QNode {                 // GP:158-371 (24 bytes: 4 sums + 4 uint16 children, 0 = leaf)
    float sum[4] = {0, 0, 0, 0};
    uint16_t child[4] = {0, 0, 0, 0};
};

// child slot of p inside a node, rescaling p into it: bit0 = x>=0.5, bit1 = y>=0.5 (GP:205-217)
static inline int quad_child_index(V2 &p) {
    int res = 0;
    if (p.x < 0.5f) p.x *= 2; else { p.x = (p.x - 0.5f) * 2; res |= 1; }
    if (p.y < 0.5f) p.y *= 2; else { p.y = (p.y - 0.5f) * 2; res |= 2; }
    return res;
}

template <class Rng> struct RngOps;   // next1D(rng), next2D(rng) provided by the tracer / tests

struct DTree {                  // GP:374-560
    std::vector<QNode> nodes;
    float sum = 0, weight = 0;  // m_atomic.sum / m_atomic.statisticalWeight
    int maxDepth = 0;

    DTree() { nodes.emplace_back(); }

    float mean() const {        // GP:387-393
        if (weight == 0) return 0;
        const float factor = 1 / (kPi * 4 * weight);
        return factor * sum;
    }

    // GP:415-421 + 232-245 (iterative)
    float pdf(V2 p) const {
        if (!(mean() > 0)) return 1 / (4 * kPi);
        float result = 1.0f;
        uint32_t n = 0;
        // the reference multiplies factor * (recursive result): accumulate the factors
        // in a list and fold from the leaf upward to keep the same rounding order
        float factors[64]; int nf = 0;
        for (;;) {
            const QNode &q = nodes[n];
            const int c = quad_child_index(p);
            if (!(q.sum[c] > 0)) { factors[nf++] = 0; break; }
            const float factor = 4 * q.sum[c] / (q.sum[0] + q.sum[1] + q.sum[2] + q.sum[3]);
            factors[nf++] = factor;
            if (q.child[c] == 0) break;
            n = q.child[c];
        }
        result = factors[nf - 1];
        for (int i = nf - 2; i >= 0; --i) result = factors[i] * result;
        return result / (4 * kPi);
    }

    // GP:423-425 + 247-255
    int depthAt(V2 p) const {
        int d = 1; uint32_t n = 0;
        for (;;) {
            const int c = quad_child_index(p);
            if (nodes[n].child[c] == 0) return d;
            n = nodes[n].child[c]; ++d;
        }
    }

    // GP:431-442 + 257-301.  The reference returns origin + 0.5 * child_sample
    // recursively; unrolled here by collecting the origins and folding from the
    // leaf upward (same arithmetic order).
    template <class Rng> V2 sample(Rng &rng) const {
        if (!(mean() > 0)) return rng.next2D();
        V2 origins[64]; int no = 0;
        uint32_t n = 0;
        V2 res;
        for (;;) {
            const QNode &q = nodes[n];
            int index = 0;
            const float topLeft = q.sum[0], topRight = q.sum[1];
            float partial = topLeft + q.sum[2];
            const float total = partial + topRight + q.sum[3];
            if (!(total > 0.0f)) { res = rng.next2D(); break; }
            float boundary = partial / total;
            V2 origin{0.0f, 0.0f};
            float s = rng.next1D();
            if (s < boundary) {
                s /= boundary;
                boundary = topLeft / partial;
            } else {
                partial = total - partial;
                origin.x = 0.5f;
                s = (s - boundary) / (1.0f - boundary);
                boundary = topRight / partial;
                index |= 1;
            }
            if (s < boundary) {
                s /= boundary;
            } else {
                origin.y = 0.5f;
                s = (s - boundary) / (1.0f - boundary);
                index |= 2;
            }
            origins[no++] = origin;
            if (q.child[index] == 0) { res = rng.next2D(); break; }   // leaf: origin + 0.5 * next2D
            n = q.child[index];
        }
        for (int i = no - 1; i >= 0; --i) { res.x = origins[i].x + 0.5f * res.x; res.y = origins[i].y + 0.5f * res.y; }
        res.x = std::min(std::max(res.x, 0.0f), 1.0f);
        res.y = std::min(std::max(res.y, 0.0f), 1.0f);
        return res;
    }

    // GP:303-312
    void recordNearest(V2 p, float value) {
        uint32_t n = 0;
        for (;;) {
            const int c = quad_child_index(p);
            if (nodes[n].child[c] == 0) { atomic_add(nodes[n].sum[c], value); return; }
            n = nodes[n].child[c];
        }
    }

    // GP:322-338 (explicit stack instead of recursion; same child visiting order 0..3, depth first)
    void recordBox(V2 origin, float size, float value) {
        struct E { uint32_t n; V2 o; float s; };
        E st[4 * 24]; int sp = 0;
        st[sp++] = E{0, V2{0, 0}, 1.0f};
        while (sp) {
            E e = st[--sp];
            const float childSize = e.s / 2;
            // push in reverse so that children pop in 0..3 order with depth-first descent like the recursion
            E pend[4]; int np = 0;
            for (int i = 0; i < 4; ++i) {
                V2 co = e.o;
                if (i & 1) co.x += childSize;
                if (i & 2) co.y += childSize;
                const float lx = std::max(std::min(origin.x + size, co.x + childSize) - std::max(origin.x, co.x), 0.0f);
                const float ly = std::max(std::min(origin.y + size, co.y + childSize) - std::max(origin.y, co.y), 0.0f);
                const float w = lx * ly;          // GP:314-320
                if (w > 0.0f) {
                    if (nodes[e.n].child[i] == 0) atomic_add(nodes[e.n].sum[i], value * w);
                    else pend[np++] = E{nodes[e.n].child[i], co, childSize};
                }
            }
            for (int i = np - 1; i >= 0; --i) st[sp++] = pend[i];
        }
    }

    // GP:395-413
    void recordIrradiance(V2 p, float irradiance, float statisticalWeight, int directionalFilter) {
        if (std::isfinite(statisticalWeight) && statisticalWeight > 0) {
            atomic_add(weight, statisticalWeight);
            if (std::isfinite(irradiance) && irradiance > 0) {
                if (directionalFilter == 0) {
                    recordNearest(p, irradiance * statisticalWeight);
                } else {
                    const int depth = depthAt(p);
                    const float size = std::pow(0.5f, depth);
                    V2 origin = p;
                    origin.x -= size / 2;
                    origin.y -= size / 2;
                    recordBox(origin, size, irradiance * statisticalWeight / (size * size));
                }
            }
        }
    }

    // GP:456-514: new topology = refinement of `prev`; all sums zero afterwards.
    void reset(const DTree &prev, int newMaxDepth, float subdivisionThreshold) {
        sum = 0; weight = 0; maxDepth = 0;
        nodes.clear(); nodes.emplace_back();
        struct S { size_t nodeIndex, otherNodeIndex; bool otherIsPrev; int depth; };
        std::vector<S> stack;
        stack.push_back(S{0, 0, true, 1});
        const float total = prev.sum;
        while (!stack.empty()) {
            S s = stack.back(); stack.pop_back();
            maxDepth = std::max(maxDepth, s.depth);
            for (int i = 0; i < 4; ++i) {
                const QNode otherNode = s.otherIsPrev ? prev.nodes[s.otherNodeIndex] : nodes[s.otherNodeIndex];
                const float fraction = total > 0 ? (otherNode.sum[i] / total) : std::pow(0.25f, s.depth);
                if (s.depth < newMaxDepth && fraction > subdivisionThreshold) {
                    if (s.otherIsPrev && otherNode.child[i] != 0) stack.push_back(S{nodes.size(), otherNode.child[i], true, s.depth + 1});
                    else stack.push_back(S{nodes.size(), nodes.size(), false, s.depth + 1});
                    nodes[s.nodeIndex].child[i] = (uint16_t) nodes.size();
                    nodes.emplace_back();
                    const float quarter = otherNode.sum[i] / 4;
                    for (int j = 0; j < 4; ++j) nodes.back().sum[j] = quarter;
                    if (nodes.size() > std::numeric_limits<uint16_t>::max()) { stack.clear(); break; }
                }
            }
        }
        for (auto &q : nodes) for (int j = 0; j < 4; ++j) q.sum[j] = 0;
    }

    // GP:520-533 + 346-366.  Children always have larger indices than their parent
    // (reset appends), so one reverse sweep equals the recursion.
    void build() {
        for (size_t n = nodes.size(); n-- > 0;) {
            QNode &q = nodes[n];
            for (int i = 0; i < 4; ++i) {
                if (q.child[i] == 0) continue;
                const QNode &c = nodes[q.child[i]];
                float s = 0;
                for (int j = 0; j < 4; ++j) s += c.sum[j];
                q.sum[i] = s;
            }
        }
        float s = 0;
        for (int i = 0; i < 4; ++i) s += nodes[0].sum[i];
        sum = s;
    }
};

struct DTreeRecord {            // GP:562-568
    V3 d; float radiance, product, woPdf, bsdfPdf, dTreePdf, statisticalWeight; bool isDelta;
};

struct DTreeWrapper {           // GP:570-738
    DTree building, sampling;
    Adam opt;
    std::atomic_flag lock = ATOMIC_FLAG_INIT;

    DTreeWrapper() {}
    DTreeWrapper(const DTreeWrapper &o) : building(o.building), sampling(o.sampling), opt(o.opt) {}
    DTreeWrapper &operator=(const DTreeWrapper &o) { building = o.building; sampling = o.sampling; opt = o.opt; return *this; }

    float bsdfSamplingFraction() const { return logistic(opt.variable); }   // GP:659-670

    // GP:672-697
    void optimize(const DTreeRecord &rec, float ratioPower) {
        while (lock.test_and_set(std::memory_order_acquire)) {}
        const float variable = opt.variable;
        const float f = logistic(variable);
        const float mixPdf = f * rec.bsdfPdf + (1 - f) * rec.dTreePdf;
        const float ratio = std::pow(rec.product / mixPdf, ratioPower);
        const float dLoss_df = -ratio / rec.woPdf * (rec.bsdfPdf - rec.dTreePdf);
        const float dLoss_dv = dLoss_df * (f * (1 - f));
        const float l2 = 0.01f * variable;
        opt.append(l2 + dLoss_dv, rec.statisticalWeight);
        lock.clear(std::memory_order_release);
    }

    // GP:575-584
    void record(const DTreeRecord &rec, int directionalFilter, int loss) {
        if (!rec.isDelta) {
            const float irradiance = rec.radiance / rec.woPdf;
            building.recordIrradiance(dir_to_canonical(rec.d), irradiance, rec.statisticalWeight, directionalFilter);
        }
        if (loss != 0 && rec.product > 0) optimize(rec, loss == 1 ? 1.0f : 2.0f);
    }

    void build() { building.build(); sampling = building; }                              // GP:610-613
    void reset(int maxDepth, float thr) { building.reset(sampling, maxDepth, thr); }      // GP:615-617
    template <class Rng> V3 sample(Rng &rng) const { return canonical_to_dir(sampling.sample(rng)); }   // GP:619-621
    float pdf(const V3 &d) const { return sampling.pdf(dir_to_canonical(d)); }            // GP:623-625
};

// ---------------------------------------------------------------- S-tree (GP:740-1007)
struct SNode {
    bool isLeaf = true;
    int axis = 0;
    uint32_t children[2] = {0, 0};
    DTreeWrapper dTree;
};

struct STree {
    std::vector<SNode> nodes;
    V3 amin, amax;

    STree(const V3 &mn, const V3 &mx) {           // GP:850-860: cubify from the min corner
        nodes.emplace_back();
        amin = mn;
        const float sx = mx.x - mn.x, sy = mx.y - mn.y, sz = mx.z - mn.z;
        const float m = std::max(std::max(sx, sy), sz);
        amax = V3{mn.x + m, mn.y + m, mn.z + m};
    }
    V3 extents() const { return V3{amax.x - amin.x, amax.y - amin.y, amax.z - amin.z}; }

    // GP:897-905 + 761-769 (iterative)
    uint32_t lookup(V3 p, V3 *sizeOut = nullptr) const {
        V3 size = extents();
        p = V3{p.x - amin.x, p.y - amin.y, p.z - amin.z};
        p.x /= size.x; p.y /= size.y; p.z /= size.z;
        uint32_t n = 0;
        while (!nodes[n].isLeaf) {
            const int a = nodes[n].axis;
            size[a] /= 2;
            int c;
            if (p[a] < 0.5f) { p[a] *= 2; c = 0; } else { p[a] = (p[a] - 0.5f) * 2; c = 1; }   // GP:747-755
            n = nodes[n].children[c];
        }
        if (sizeOut) *sizeOut = size;
        return n;
    }

    // GP:876-895
    void subdivide(uint32_t idx) {
        nodes.resize(nodes.size() + 2);
        SNode &cur = nodes[idx];
        for (int i = 0; i < 2; ++i) {
            const uint32_t c = (uint32_t) nodes.size() - 2 + i;
            cur.children[i] = c;
            nodes[c].axis = (cur.axis + 1) % 3;
            nodes[c].dTree = cur.dTree;
            nodes[c].dTree.building.weight = nodes[c].dTree.building.weight / 2;
        }
        cur.isLeaf = false;
        cur.dTree = DTreeWrapper();
    }

    // GP:957-998 (memory cap: approximate footprint by node counts, GP:516-518)
    void refine(size_t sTreeThreshold, int maxMB) {
        if (maxMB >= 0) {
            size_t fp = 0;
            for (const auto &n : nodes) fp += (n.dTree.building.nodes.capacity() + n.dTree.sampling.nodes.capacity()) * sizeof(QNode) + 2 * 48;
            if (fp / 1000000 >= (size_t) maxMB) return;
        }
        std::vector<uint32_t> stack;
        stack.push_back(0);
        while (!stack.empty()) {
            const uint32_t i = stack.back(); stack.pop_back();
            if (nodes[i].isLeaf && nodes[i].dTree.building.weight > (float) sTreeThreshold) subdivide(i);   // GP:953-955 (size_t -> Float comparison)
            if (!nodes[i].isLeaf) { stack.push_back(nodes[i].children[0]); stack.push_back(nodes[i].children[1]); }
        }
    }

    // GP:935-943 + 823-839 (explicit stack; children visited 0 then 1, depth first)
    void recordBox(const V3 &p, const V3 &voxel, DTreeRecord rec, int directionalFilter, int loss) {
        const float volume = voxel.x * voxel.y * voxel.z;
        rec.statisticalWeight /= volume;
        const V3 min1{p.x - voxel.x * 0.5f, p.y - voxel.y * 0.5f, p.z - voxel.z * 0.5f};
        const V3 max1{p.x + voxel.x * 0.5f, p.y + voxel.y * 0.5f, p.z + voxel.z * 0.5f};
        struct E { uint32_t n; V3 mn, sz; };
        std::vector<E> st;
        st.push_back(E{0, amin, extents()});
        while (!st.empty()) {
            E e = st.back(); st.pop_back();
            float w = 1;
            for (int i = 0; i < 3; ++i) {
                const float l = std::max(std::min(max1[i], e.mn[i] + e.sz[i]) - std::max(min1[i], e.mn[i]), 0.0f);
                w *= l;                                // GP:815-821: lengths[0]*lengths[1]*lengths[2]
            }
            if (!(w > 0)) continue;
            SNode &nd = nodes[e.n];
            if (nd.isLeaf) {
                DTreeRecord r = rec; r.statisticalWeight = rec.statisticalWeight * w;
                nd.dTree.record(r, directionalFilter, loss);
            } else {
                V3 sz = e.sz; sz[nd.axis] /= 2;
                V3 mn1 = e.mn; mn1[nd.axis] += sz[nd.axis];
                st.push_back(E{nd.children[1], mn1, sz});
                st.push_back(E{nd.children[0], e.mn, sz});
            }
        }
    }

    V3 clip(V3 p) const {   // AABB::clip
        p.x = std::min(std::max(p.x, amin.x), amax.x);
        p.y = std::min(std::max(p.y, amin.y), amax.y);
        p.z = std::min(std::max(p.z, amin.z), amax.z);
        return p;
    }
};

}  // namespace ppgo
