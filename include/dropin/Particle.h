// Particle.h — the caller-visible particle record and grid bucket of the reference (SPH_SM_monodomain/Particle.h:7-41).
// Same member names, order, types and therefore the same 132-byte layout: sphsm_upload_aos / sphsm_download_aos
// (include/sphsm_b200.h) exchange exactly this image with the device-side SoA arrays.
#ifndef SPHSM_DROPIN_PARTICLE_H
#define SPHSM_DROPIN_PARTICLE_H

#include <cstddef>
#include <vector>

#include "m3Vector.h"

class Particle {
public:
    m3Vector pos;            // position
    m3Vector vel;            // velocity
    m3Vector predicted_vel;  // after external forces (stage 2a)
    m3Vector inter_vel;      // XSPH-mixed intermediate velocity (stage 3)
    m3Vector corrected_vel;  // shape-matching corrected velocity (stage 2c)
    m3Vector acc;            // pressure + viscosity acceleration (stage 6)
    float mass;

    m3Vector mOriginalPos;   // rest position
    m3Vector mGoalPos;       // shape-matching goal position
    bool mFixed;             // pinned in place

    float dens;
    float pres;

    m3Real Vm;               // transmembrane voltage
    m3Real Inter_Vm;         // dVm/dt * m (stage 6)
    m3Real Iion;             // ionic current
    m3Real stim;             // stimulation current
    m3Real w;                // recovery variable

    m3Real getDisplacement() { return (mOriginalPos - pos).magnitude(); }
};
static_assert(sizeof(Particle) == 132, "Particle must keep the reference's 132-byte layout");
// the byte offsets the device-side AoS <-> SoA kernels use (sphsm_capi.cu: OFF_*)
static_assert(offsetof(Particle, vel) == 12 && offsetof(Particle, predicted_vel) == 24 && offsetof(Particle, inter_vel) == 36 &&
                  offsetof(Particle, corrected_vel) == 48 && offsetof(Particle, acc) == 60 && offsetof(Particle, mass) == 72 &&
                  offsetof(Particle, mOriginalPos) == 76 && offsetof(Particle, mGoalPos) == 88 && offsetof(Particle, mFixed) == 100 &&
                  offsetof(Particle, dens) == 104 && offsetof(Particle, pres) == 108 && offsetof(Particle, Vm) == 112 &&
                  offsetof(Particle, Inter_Vm) == 116 && offsetof(Particle, Iion) == 120 && offsetof(Particle, stim) == 124 &&
                  offsetof(Particle, w) == 128,
              "Particle field offsets differ from the image sphsm_upload_aos / sphsm_download_aos exchange");

class Cell {
public:
    std::vector<Particle *> contained_particles;
};

#endif
