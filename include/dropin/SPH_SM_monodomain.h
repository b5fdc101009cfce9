// SPH_SM_monodomain.h — drop-in replacement for the reference's simulation class
// (SPH_SM_monodomain/SPH_SM_monodomain.h:29-168 of Hagen23/SPH-SM-Monodomain).
//
// Same class name, public methods, public data members and typedefs, so the reference's main.cpp (and any other host
// code written against the reference header) compiles unchanged with `-I include/dropin` in place of
// `-I Math3D/ -I SPH_SM_monodomain/` and links against libsphsm_dropin.so + libsphsm_b200.so.  Every numerical stage
// runs on the B200 through the C-ABI of include/sphsm_b200.h; this class only keeps the caller-visible AoS mirror
// (Particle[], Cell[]) and the run-time toggles.  There is no CPU implementation of the step in here.
//
// Differences a caller can observe (all additive):
//   * a second constructor takes capacity / world size (the reference hard-codes 50000 and 1.5^3, cpp:19,29);
//   * Get_Paticles() returns a host mirror that is refreshed lazily from the device; writes made through the pointer
//     are detected (compared with a shadow copy) and uploaded before the next device operation — see
//     set_accessor_readonly() to skip that check on large particle counts;
//   * Animation() runs the fused two-pass step; the d_* stage timers hold DEVICE time (CUDA events) sampled from every 50th
//     step and filed under the stage that dominates each kernel group (find_neighbors, corrected_velocity,
//     intermediate_velocity = pass A, compute_Force = pass B); set_stage_timing(true) (or SPHSM_STAGE_TIMING=1) runs the
//     reference's seven stages one kernel group each, with an event pair around every one, as the reference times them.
#ifndef __SPH_SM_monodomain_H__
#define __SPH_SM_monodomain_H__

#include <m3Bounds.h>
#include <m3Real.h>
#include <m3Vector.h>

#include <Particle.h>

#include <chrono>
#include <map>
#include <vector>

#define PI 3.141592f
#define INF 1E-12f  // (sic) the reference's "INF" is the near-zero distance threshold, h:24

typedef std::chrono::system_clock::time_point tpoint;
typedef std::chrono::duration<double> duration_d;

struct sphsm_handle;  // include/sphsm_b200.h

class SPH_SM_monodomain {
private:
    sphsm_handle *dev;          // owns all device state
    int Max_Number_Paticles;    // capacity (sic, reference spelling)
    int Number_Particles;
    int Number_Cells;
    m3Vector Grid_Size, World_Size;
    m3Real kernel, Cell_Size, Stand_Density, Time_Delta;
    m3Real Poly6_constant, Spiky_constant, B_spline_constant;

    Particle *Particles;        // host mirror, caller-visible through Get_Paticles()
    Particle *Shadow;           // what the mirror held right after the last download (write detection)
    Cell *Cells;                // host buckets, filled on Get_Cells()
    bool mirror_current;        // Particles[] equals the device state
    bool mirror_handed_out;     // the caller holds the pointer and may have written through it
    bool accessor_readonly;
    bool cells_current;
    bool stage_timing;
    double stage_seen[7];       // device seconds already folded into the d_* members

    void sync_public_tunables();    // voltage_constant / max_pressure / max_voltage -> device params
    void push_host_writes();        // upload the mirror if the caller changed it
    void device_changed();          // device state moved on: mirror and buckets are stale
    void refresh_mirror();
    void run_stage(int stage);
    void collect_stage_times();
    void construct(int capacity, m3Vector world);

    SPH_SM_monodomain(const SPH_SM_monodomain &);             // not copyable (owns a device handle)
    SPH_SM_monodomain &operator=(const SPH_SM_monodomain &);

public:
    SPH_SM_monodomain();                                      // cpp:13-79 defaults: capacity 50000, world 1.5^3
    SPH_SM_monodomain(int capacity, m3Vector world_size);     // extension for the synthetic lattices
    ~SPH_SM_monodomain();

    m3Real voltage_constant = 1;
    m3Real max_pressure = 15000;
    m3Real max_voltage = 200;

    // per-stage time accumulators (h:97-99); device time here, wall-clock in the reference
    tpoint t_start_find_neighbors, t_start_corrected_velocity, t_start_intermediate_velocity, t_start_Density_SingPressure,
        t_start_cell_model, t_start_compute_Force, t_start_Update_Properties;
    duration_d d_find_neighbors, d_corrected_velocity, d_intermediate_velocity, d_Density_SingPressure, d_cell_model,
        d_compute_Force, d_Update_Properties;

    int total_time_steps;

    void Init_Fluid(std::vector<m3Vector> positions);
    void Init_Particle(m3Vector pos, m3Vector vel);
    m3Vector Calculate_Cell_Position(m3Vector pos);
    int Calculate_Cell_Hash(m3Vector pos);
    void print_report(double avg_fps = 0.0f, double avg_step_d = 0.0f);
    void add_viscosity(float value);

    // smoothing kernels (host-side scalar evaluations of the same formulas the device passes use)
    m3Real Poly6(m3Real r2);
    m3Real Spiky(m3Real r);
    m3Real Visco(m3Real r);
    m3Real B_spline(m3Real r);
    m3Real B_spline_1(m3Real r);
    m3Real B_spline_2(m3Real r);

    void Find_neighbors();

    void calculate_corrected_velocity();
    void apply_external_forces(m3Vector *forcesArray = 0, int *indexArray = 0, int size = 0);
    void projectPositions();

    void calculate_cell_model();
    void set_stim(m3Vector center, m3Real radius, m3Real stim_strength);
    void turnOnStim_Cube(std::vector<m3Vector> positions);
    void turnOnStim_Mesh(std::vector<m3Vector> positions);
    void turnOffStim();

    void calculate_intermediate_velocity();
    void Compute_Density_SingPressure();
    void Compute_Force();
    void Update_Properties();

    void compute_SPH_SM_monodomain();
    void Animation();

    inline int Get_Particle_Number() { return Number_Particles; }
    inline m3Vector Get_World_Size() { return World_Size; }
    Particle *Get_Paticles();  // (sic)
    Cell *Get_Cells();
    inline m3Real Get_stand_dens() { return Stand_Density; }

    bool flip_quadratic();
    bool flip_volume();

    inline int pow2roundup(int v) {
        int p = 1;
        while (p < v && p > 0) p <<= 1;
        return v <= 0 ? 0 : p;
    }

    // ---- extensions (no reference counterpart) ----------------------------------------------------------------
    void Animation(int nsteps);                 // nsteps steps with one call (asynchronous on the device stream)
    void set_stage_timing(bool on);             // true: the seven stages run and are timed one by one (default: fused step, sampled timers)
    void set_accessor_readonly(bool on);        // true: Get_Paticles() results are never written back
    const Particle *Get_Paticles_readonly();    // refresh + return without arming the write-back check
    void download_positions(float *xyz);        // 3 floats per particle, what a viewer needs each frame
    void synchronize();
    sphsm_handle *native_handle() { return dev; }
};

#endif
