// m3Bounds.h — axis-aligned box with the surface of the reference's Math3D/m3Bounds.h:9-99 that host code can reach
// (the simulation's own bounds clamp, m3Bounds.h:84-88, is part of the GPU integration kernel).
#ifndef SPHSM_DROPIN_M3BOUNDS_H
#define SPHSM_DROPIN_M3BOUNDS_H

#include "math3d.h"

class m3Bounds {
public:
    m3Vector min, max;

    m3Bounds() { setEmpty(); }
    m3Bounds(const m3Vector &lo, const m3Vector &hi) : min(lo), max(hi) {}
    void set(const m3Vector &lo, const m3Vector &hi) { min = lo; max = hi; }
    void setEmpty() { set(m3Vector(m3RealMax, m3RealMax, m3RealMax), m3Vector(m3RealMin, m3RealMin, m3RealMin)); }
    void setInfinite() { set(m3Vector(m3RealMin, m3RealMin, m3RealMin), m3Vector(m3RealMax, m3RealMax, m3RealMax)); }
    bool isEmpty() const { return min.x > max.x || min.y > max.y || min.z > max.z; }
    bool operator==(const m3Bounds &b) const { return min == b.min && max == b.max; }

    void combine(const m3Bounds &b) { min.minimum(b.min); max.maximum(b.max); }
    void operator+=(const m3Bounds &b) { combine(b); }
    m3Bounds operator+(const m3Bounds &b) const { m3Bounds r(*this); r.combine(b); return r; }
    void intersect(const m3Bounds &b) { min.maximum(b.min); max.minimum(b.max); }
    bool intersects(const m3Bounds &b) const {  // the reference tests x and y only (m3Bounds.h:55-59)
        return !(b.min.x > max.x || min.x > b.max.x || b.min.y > max.y || min.y > b.max.y);
    }
    void include(const m3Vector &v) { max.maximum(v); min.minimum(v); }
    void operator+=(const m3Vector &v) { include(v); }
    bool contain(const m3Vector &v) const { return min.x <= v.x && v.x <= max.x && min.y <= v.y && v.y <= max.y; }
    void getCenter(m3Vector &c) const { c = (min + max) * 0.5f; }

    void clamp(m3Vector &p) const {
        if (isEmpty()) return;
        p.maximum(min);
        p.minimum(max);
    }
    void clamp(m3Vector &p, m3Real off) const {
        if (isEmpty()) return;
        if (p.x < min.x + off) p.x = min.x + off;
        if (p.x > max.x - off) p.x = max.x - off;
        if (p.y < min.y + off) p.y = min.y + off;
        if (p.y > max.y - off) p.y = max.y - off;
    }
};

#endif
