// m3Bounds.h — the axis-aligned box type that `SPH_SM_monodomain.h` pulls in (reference Math3D/m3Bounds.h).  The
// simulation's own wall clamp (m3Bounds.h:84-88) lives in the GPU integration kernel; host code of the reference never
// touches the class (main.cpp does not name it), so this header keeps the type and the handful of operations a caller
// could reasonably reach — corners, emptiness, growing, clamping — written over a small per-axis helper.
#ifndef SPHSM_DROPIN_M3BOUNDS_H
#define SPHSM_DROPIN_M3BOUNDS_H

#include "math3d.h"

class m3Bounds {
    // apply f(lo_component, hi_component, axis) to the three axes
    template <class F>
    void eachAxis(F f) {
        for (int a = 0; a < 3; a++) f(min[a], max[a], a);
    }

public:
    m3Vector min, max;  // corners; the box is empty while any min component exceeds its max

    m3Bounds() { setEmpty(); }
    m3Bounds(const m3Vector &lo, const m3Vector &hi) { set(lo, hi); }

    void set(const m3Vector &lo, const m3Vector &hi) {
        min = lo;
        max = hi;
    }
    void setEmpty() {
        eachAxis([](m3Real &lo, m3Real &hi, int) {
            lo = m3RealMax;
            hi = m3RealMin;
        });
    }
    bool isEmpty() const {
        for (int a = 0; a < 3; a++)
            if (min[a] > max[a]) return true;
        return false;
    }

    // grow to contain a point / another box
    void include(const m3Vector &v) {
        eachAxis([&v](m3Real &lo, m3Real &hi, int a) {
            if (v[a] < lo) lo = v[a];
            if (v[a] > hi) hi = v[a];
        });
    }
    void combine(const m3Bounds &other) {
        if (other.isEmpty()) return;
        include(other.min);
        include(other.max);
    }
    m3Vector center() const { return (min + max) * 0.5f; }

    // move a point onto the box (no-op for an empty box), as Update_Properties does at the end of a step
    void clamp(m3Vector &p) const {
        if (isEmpty()) return;
        for (int a = 0; a < 3; a++) {
            if (p[a] < min[a]) p[a] = min[a];
            if (p[a] > max[a]) p[a] = max[a];
        }
    }
};

#endif
