// math3d.h — umbrella header kept for source compatibility with the reference (Math3D/math3d.h).
#ifndef SPHSM_DROPIN_MATH3D_H
#define SPHSM_DROPIN_MATH3D_H
#include <cstdlib>

#include "m3Vector.h"
#endif
