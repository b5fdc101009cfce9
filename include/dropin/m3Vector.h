// m3Vector.h — float 3-vector with the public surface of the reference's Math3D/m3Vector.h:11-123 (x, y, z members,
// value-semantics arithmetic, TRUE division in operator/), so host code written against the reference compiles
// unchanged.  Layout is three consecutive floats (12 bytes): Particle (Particle.h) depends on it.
#ifndef SPHSM_DROPIN_M3VECTOR_H
#define SPHSM_DROPIN_M3VECTOR_H

#include <cassert>

#include "m3Real.h"

class m3Vector {
public:
    m3Real x, y, z;

    m3Vector() : x(0.0f), y(0.0f), z(0.0f) {}
    m3Vector(m3Real x0, m3Real y0, m3Real z0) : x(x0), y(y0), z(z0) {}
    m3Vector(const m3Vector &o) : x(o.x), y(o.y), z(o.z) {}
    m3Vector &operator=(const m3Vector &o) { x = o.x; y = o.y; z = o.z; return *this; }

    void set(m3Real x0, m3Real y0, m3Real z0) { x = x0; y = y0; z = z0; }
    void zero() { x = y = z = 0.0f; }
    bool isZero() const { return x == 0.0f && y == 0.0f && z == 0.0f; }

    m3Real &operator[](int i) { assert(i >= 0 && i <= 2); return (&x)[i]; }
    const m3Real &operator[](int i) const { assert(i >= 0 && i <= 2); return (&x)[i]; }

    bool operator==(const m3Vector &v) const { return x == v.x && y == v.y && z == v.z; }

    m3Vector operator+(const m3Vector &v) const { return m3Vector(x + v.x, y + v.y, z + v.z); }
    m3Vector operator-(const m3Vector &v) const { return m3Vector(x - v.x, y - v.y, z - v.z); }
    m3Vector operator-() const { return m3Vector(-x, -y, -z); }
    m3Vector operator*(m3Real f) const { return m3Vector(x * f, y * f, z * f); }
    m3Vector operator/(m3Real f) const { return m3Vector(x / f, y / f, z / f); }  // division, not reciprocal-multiply

    void operator+=(const m3Vector &v) { x += v.x; y += v.y; z += v.z; }
    void operator-=(const m3Vector &v) { x -= v.x; y -= v.y; z -= v.z; }
    void operator*=(const m3Vector &v) { x *= v.x; y *= v.y; z *= v.z; }
    void operator/=(const m3Vector &v) { x /= v.x; y /= v.y; z /= v.z; }
    void operator*=(m3Real f) { x *= f; y *= f; z *= f; }
    void operator/=(m3Real f) { x /= f; y /= f; z /= f; }

    m3Vector cross(const m3Vector &a, const m3Vector &b) const {
        return m3Vector(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    }
    m3Real dot(const m3Vector &v) const { return x * v.x + y * v.y + z * v.z; }

    void minimum(const m3Vector &v) { if (v.x < x) x = v.x; if (v.y < y) y = v.y; if (v.z < z) z = v.z; }
    void maximum(const m3Vector &v) { if (v.x > x) x = v.x; if (v.y > y) y = v.y; if (v.z > z) z = v.z; }

    m3Real magnitudeSquared() const { return x * x + y * y + z * z; }
    m3Real magnitude() const { return sqrtf(magnitudeSquared()); }
    m3Real distanceSquared(const m3Vector &v) const { return (v - *this).magnitudeSquared(); }
    m3Real distance(const m3Vector &v) const { return (v - *this).magnitude(); }

    void normalize() {
        const m3Real l = magnitude();
        if (l != 0.0f) *this *= 1.0f / l;
    }
};

#endif
