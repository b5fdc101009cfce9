// m3Real.h — scalar type of the drop-in headers (API of the reference's Math3D/m3Real.h:9-67).
// Written for sphsm-b200; the simulation arithmetic itself runs on the GPU (libsphsm_b200.so), these helpers only
// serve host code that includes the reference's header names.
#ifndef SPHSM_DROPIN_M3REAL_H
#define SPHSM_DROPIN_M3REAL_H

#include <cfloat>
#include <cmath>
#include <cstdlib>

typedef float m3Real;  // Math3D/m3Real.h:17

#define m3Pi 3.1415926535897932f
#define m3HalfPi 1.5707963267948966f
#define m3TwoPi 6.2831853071795865f
#define m3RealMax FLT_MAX
#define m3RealMin FLT_MIN
#define m3RadToDeg 57.295779513082321f
#define m3DegToRad 0.0174532925199433f

inline m3Real m2Clamp(m3Real &r, m3Real lo, m3Real hi) { return r < lo ? lo : (r > hi ? hi : r); }
inline m3Real m2Min(m3Real a, m3Real b) { return a <= b ? a : b; }
inline m3Real m2Max(m3Real a, m3Real b) { return a >= b ? a : b; }
inline m3Real m2Abs(m3Real r) { return r < 0.0f ? -r : r; }
inline m3Real m2Random(m3Real lo, m3Real hi) { return lo + ((m3Real)rand() / RAND_MAX) * (hi - lo); }
inline m3Real m2Acos(m3Real r) { return acos(r < -1.0f ? -1.0f : (r > 1.0f ? 1.0f : r)); }

#endif
