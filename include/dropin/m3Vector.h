// m3Vector.h — float 3-vector with the public surface of the reference's Math3D/m3Vector.h:11-123 (x, y, z members,
// value-semantics arithmetic, TRUE division in operator/), so host code written against the reference compiles
// unchanged.  Layout is three consecutive floats (12 bytes): Particle (Particle.h) depends on it.
// The component-wise operators are generated from one macro each; only the names and semantics follow the reference.
#ifndef SPHSM_DROPIN_M3VECTOR_H
#define SPHSM_DROPIN_M3VECTOR_H

#include <cassert>

#include "m3Real.h"

// v OP= w (component-wise, vector right-hand side) and v OP= f (scalar right-hand side)
#define SPHSM_M3V_COMPOUND_VEC(OP) \
    void operator OP(const m3Vector &w) { x OP w.x; y OP w.y; z OP w.z; }
#define SPHSM_M3V_COMPOUND_SCALAR(OP) \
    void operator OP(m3Real f) { x OP f; y OP f; z OP f; }
// r = v OP w / r = v OP f as value-returning members
#define SPHSM_M3V_BINARY_VEC(OP) \
    m3Vector operator OP(const m3Vector &w) const { return m3Vector(x OP w.x, y OP w.y, z OP w.z); }
#define SPHSM_M3V_BINARY_SCALAR(OP) \
    m3Vector operator OP(m3Real f) const { return m3Vector(x OP f, y OP f, z OP f); }

class m3Vector {
public:
    m3Real x, y, z;

    m3Vector() : x(0.0f), y(0.0f), z(0.0f) {}
    m3Vector(m3Real x0, m3Real y0, m3Real z0) : x(x0), y(y0), z(z0) {}
    // (copy construction / assignment: the implicit member-wise ones)

    void set(m3Real x0, m3Real y0, m3Real z0) { *this = m3Vector(x0, y0, z0); }
    void zero() { set(0.0f, 0.0f, 0.0f); }
    bool isZero() const { return *this == m3Vector(); }

    m3Real &operator[](int axis) {
        assert(axis >= 0 && axis <= 2);
        return axis == 0 ? x : (axis == 1 ? y : z);
    }
    const m3Real &operator[](int axis) const {
        assert(axis >= 0 && axis <= 2);
        return axis == 0 ? x : (axis == 1 ? y : z);
    }
    bool operator==(const m3Vector &w) const { return x == w.x && y == w.y && z == w.z; }

    SPHSM_M3V_BINARY_VEC(+)
    SPHSM_M3V_BINARY_VEC(-)
    SPHSM_M3V_BINARY_SCALAR(*)
    SPHSM_M3V_BINARY_SCALAR(/)  // a division per component, not a multiplication by the reciprocal (m3Vector.h:60)
    m3Vector operator-() const { return m3Vector(-x, -y, -z); }

    SPHSM_M3V_COMPOUND_VEC(+=)
    SPHSM_M3V_COMPOUND_VEC(-=)
    SPHSM_M3V_COMPOUND_VEC(*=)
    SPHSM_M3V_COMPOUND_VEC(/=)
    SPHSM_M3V_COMPOUND_SCALAR(*=)
    SPHSM_M3V_COMPOUND_SCALAR(/=)

    // the reference's cross() ignores *this and returns a x b (m3Vector.h:80-85)
    m3Vector cross(const m3Vector &a, const m3Vector &b) const {
        return m3Vector(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    }
    m3Real dot(const m3Vector &w) const { return x * w.x + y * w.y + z * w.z; }

    // component-wise min / max with another vector, in place
    void minimum(const m3Vector &w) {
        for (int a = 0; a < 3; a++)
            if (w[a] < (*this)[a]) (*this)[a] = w[a];
    }
    void maximum(const m3Vector &w) {
        for (int a = 0; a < 3; a++)
            if (w[a] > (*this)[a]) (*this)[a] = w[a];
    }

    m3Real magnitudeSquared() const { return dot(*this); }  // x*x + y*y + z*z, left to right
    m3Real magnitude() const { return sqrtf(magnitudeSquared()); }
    m3Real distanceSquared(const m3Vector &w) const { return (w - *this).magnitudeSquared(); }
    m3Real distance(const m3Vector &w) const { return (w - *this).magnitude(); }

    void normalize() {
        const m3Real len = magnitude();
        if (len != 0.0f) *this *= 1.0f / len;
    }
};

#undef SPHSM_M3V_COMPOUND_VEC
#undef SPHSM_M3V_COMPOUND_SCALAR
#undef SPHSM_M3V_BINARY_VEC
#undef SPHSM_M3V_BINARY_SCALAR

#endif
