// SPH_SM_monodomain.cpp — host side of the drop-in class (include/dropin/SPH_SM_monodomain.h).
//
// Mirrors the reference's SPH_SM_monodomain (SPH_SM_monodomain/SPH_SM_monodomain.cpp) method by method, but every
// numerical stage is one call into libsphsm_b200.so (include/sphsm_b200.h); this file holds no simulation arithmetic
// apart from the scalar kernel-function accessors the reference exposes as public methods.  Failures of the device
// layer are reported on stderr and abort: the reference's methods are void and silent, and there is no CPU fallback
// to continue on.
#include <SPH_SM_monodomain.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "../../include/sphsm_b200.h"

using namespace std;

namespace {
void must(sphsm_handle *h, int rc, const char *what) {
    if (rc == SPHSM_OK) return;
    fprintf(stderr, "SPH_SM_monodomain (B200): %s failed (%d): %s\n", what, rc, sphsm_last_error(h));
    abort();
}
std::vector<float> flatten(const std::vector<m3Vector> &v) {
    std::vector<float> out(v.size() * 3);
    for (size_t i = 0; i < v.size(); i++) {
        out[3 * i] = v[i].x;
        out[3 * i + 1] = v[i].y;
        out[3 * i + 2] = v[i].z;
    }
    return out;
}
}  // namespace

void SPH_SM_monodomain::construct(int capacity, m3Vector world) {
    sphsm_params p;
    must(0, sphsm_default_params(&p), "sphsm_default_params");  // the ctor literals of cpp:13-69, bit for bit
    p.capacity = capacity;
    p.world[0] = world.x; p.world[1] = world.y; p.world[2] = world.z;
    p.diagnostics = 1;  // keep all 33 Particle fields current: Get_Paticles() shows what the reference would
    if (const char *d = getenv("SPHSM_DEVICE")) p.device = atoi(d);
    if (const char *st = getenv("SPHSM_STRICT")) p.strict = atoi(st) != 0;  // reference-order arithmetic (bit-level validation)
    dev = 0;
    must(0, sphsm_create(&p, &dev), "sphsm_create");

    Max_Number_Paticles = capacity;
    Number_Particles = 0;
    total_time_steps = 0;
    kernel = Cell_Size = p.kernel_h;
    World_Size = world;
    int g[3];
    must(dev, sphsm_grid_size(dev, g), "sphsm_grid_size");
    Grid_Size = m3Vector((m3Real)g[0], (m3Real)g[1], (m3Real)g[2]);
    Number_Cells = sphsm_num_cells(dev);
    Stand_Density = p.stand_density;
    Time_Delta = p.time_delta;
    Poly6_constant = p.poly6_constant;
    Spiky_constant = p.spiky_constant;
    B_spline_constant = p.bspline_constant;

    Particles = new Particle[Max_Number_Paticles];
    Shadow = 0;  // allocated when the pointer is first handed out
    Cells = new Cell[Number_Cells];
    mirror_current = true;  // nothing on the device yet
    mirror_handed_out = false;
    accessor_readonly = false;
    cells_current = false;
    // Animation() runs the fused step (two neighbour passes, CUDA-graph replay on small sets).  The d_* timers of the reference
    // (h:97-99) are fed by SAMPLING: every SAMPLE_PERIOD-th step runs with per-kernel-group CUDA events and stands for the
    // period (see Animation); set_stage_timing(true) switches to the reference's seven separately timed stages instead.
    stage_timing = getenv("SPHSM_STAGE_TIMING") && atoi(getenv("SPHSM_STAGE_TIMING")) != 0;
    for (int k = 0; k < 7; k++) stage_seen[k] = 0.0;
    d_find_neighbors = d_corrected_velocity = d_intermediate_velocity = d_Density_SingPressure = d_cell_model = d_compute_Force =
        d_Update_Properties = duration_d::zero();
    must(dev, sphsm_enable_stage_timing(dev, stage_timing ? 1 : 0), "sphsm_enable_stage_timing");

    // the reference's banner, cpp:71-78 (it prints Grid_Size.y on the Z line too)
    cout << "SPHSystem" << endl;
    cout << "Grid_Size_X : " << Grid_Size.x << endl;
    cout << "Grid_Size_Y : " << Grid_Size.y << endl;
    cout << "Grid_Size_Z : " << Grid_Size.y << endl;
    cout << "Alpha :" << p.alpha << " Beta :" << p.beta << endl;
    cout << "Volume conservation :" << (p.volume_conservation != 0) << " Quadratic match : " << (p.quadratic_match != 0) << endl;
    cout << "Cell Number : " << Number_Cells << endl;
    cout << "Time Delta : " << Time_Delta << endl;
}

SPH_SM_monodomain::SPH_SM_monodomain() { construct(50000, m3Vector(1.5f, 1.5f, 1.5f)); }
SPH_SM_monodomain::SPH_SM_monodomain(int capacity, m3Vector world_size) { construct(capacity, world_size); }

SPH_SM_monodomain::~SPH_SM_monodomain() {
    sphsm_destroy(dev);
    delete[] Particles;
    delete[] Shadow;
    delete[] Cells;
}

// ---- mirror protocol ------------------------------------------------------------------------------------------
void SPH_SM_monodomain::sync_public_tunables() {
    sphsm_params p;
    must(dev, sphsm_get_params(dev, &p), "sphsm_get_params");
    if (p.voltage_constant != voltage_constant || p.max_pressure != max_pressure || p.max_voltage != max_voltage) {
        p.voltage_constant = voltage_constant;
        p.max_pressure = max_pressure;
        p.max_voltage = max_voltage;
        must(dev, sphsm_set_params(dev, &p), "sphsm_set_params");
    }
}

void SPH_SM_monodomain::push_host_writes() {
    sync_public_tunables();
    if (!mirror_handed_out || accessor_readonly || Number_Particles == 0) return;
    mirror_handed_out = false;
    if (Shadow && memcmp(Shadow, Particles, sizeof(Particle) * (size_t)Number_Particles) == 0) return;  // caller only read
    must(dev, sphsm_upload_aos(dev, Particles, Number_Particles, (int)sizeof(Particle)), "sphsm_upload_aos");
}

void SPH_SM_monodomain::device_changed() {
    mirror_current = false;
    cells_current = false;
}

void SPH_SM_monodomain::refresh_mirror() {
    if (mirror_current || Number_Particles == 0) return;
    must(dev, sphsm_download_aos(dev, Particles, Number_Particles, (int)sizeof(Particle)), "sphsm_download_aos");
    mirror_current = true;
}

Particle *SPH_SM_monodomain::Get_Paticles() {
    push_host_writes();  // a second call must not lose writes made through the first pointer
    refresh_mirror();
    if (!accessor_readonly && Number_Particles > 0) {
        if (!Shadow) Shadow = new Particle[Max_Number_Paticles];
        memcpy(Shadow, Particles, sizeof(Particle) * (size_t)Number_Particles);
        mirror_handed_out = true;
    }
    return Particles;
}

const Particle *SPH_SM_monodomain::Get_Paticles_readonly() {
    push_host_writes();
    refresh_mirror();
    return Particles;
}

Cell *SPH_SM_monodomain::Get_Cells() {
    push_host_writes();
    refresh_mirror();
    if (!cells_current) {
        std::vector<int> start(Number_Cells + 1), idx(Number_Particles > 0 ? Number_Particles : 1);
        must(dev, sphsm_get_cells_csr(dev, start.data(), idx.data()), "sphsm_get_cells_csr");
        for (int c = 0; c < Number_Cells; c++) {
            std::vector<Particle *> &b = Cells[c].contained_particles;
            b.clear();
            for (int k = start[c]; k < start[c + 1]; k++) b.push_back(&Particles[idx[k]]);
        }
        cells_current = true;
    }
    return Cells;
}

void SPH_SM_monodomain::download_positions(float *xyz) {
    push_host_writes();
    if (Number_Particles > 0) must(dev, sphsm_download_positions(dev, xyz, Number_Particles), "sphsm_download_positions");
}

void SPH_SM_monodomain::synchronize() { must(dev, sphsm_sync(dev), "sphsm_sync"); }
void SPH_SM_monodomain::set_accessor_readonly(bool on) { accessor_readonly = on; }

void SPH_SM_monodomain::set_stage_timing(bool on) {
    stage_timing = on;
    must(dev, sphsm_enable_stage_timing(dev, on ? 1 : 0), "sphsm_enable_stage_timing");
}

// ---- initialisation -----------------------------------------------------------------------------------------------
void SPH_SM_monodomain::Init_Fluid(std::vector<m3Vector> positions) {  // cpp:93-99
    push_host_writes();
    const std::vector<float> xyz = flatten(positions);
    must(dev, sphsm_init_fluid(dev, xyz.data(), (int)positions.size()), "sphsm_init_fluid");
    Number_Particles = sphsm_num_particles(dev);  // particles beyond capacity were dropped, cpp:103
    device_changed();
    cout << "Number of Paticles : " << Number_Particles << endl;
}

void SPH_SM_monodomain::Init_Particle(m3Vector pos, m3Vector vel) {  // cpp:101-125
    if (Number_Particles + 1 > Max_Number_Paticles) return;
    push_host_writes();
    const float xyz[3] = {pos.x, pos.y, pos.z};
    must(dev, sphsm_init_fluid(dev, xyz, 1), "sphsm_init_fluid");
    Number_Particles = sphsm_num_particles(dev);
    device_changed();
    if (vel.x != 0.0f || vel.y != 0.0f || vel.z != 0.0f) {  // Init_Fluid always passes zero; honour the general case
        Particle *p = Get_Paticles();
        p[Number_Particles - 1].vel = vel;
    }
}

// ---- scalar helpers (public in the reference, h:105-117) ---------------------------------------------------------------
m3Vector SPH_SM_monodomain::Calculate_Cell_Position(m3Vector pos) {  // cpp:127-134: float division, C truncation
    m3Vector c = pos / Cell_Size;
    return m3Vector((m3Real)(int)c.x, (m3Real)(int)c.y, (m3Real)(int)c.z);
}

int SPH_SM_monodomain::Calculate_Cell_Hash(m3Vector cell) {  // cpp:136-146: -1 outside the grid; float evaluation
    const bool outside = cell.x < 0 || cell.x >= Grid_Size.x || cell.y < 0 || cell.y >= Grid_Size.y || cell.z < 0 || cell.z >= Grid_Size.z;
    if (outside) return -1;
    return (int)(cell.x + Grid_Size.x * (cell.y + Grid_Size.y * cell.z));
}

m3Real SPH_SM_monodomain::Poly6(m3Real r2) {  // cpp:149-152 (double pow)
    const m3Real h2 = kernel * kernel;
    if (!(r2 >= 0 && r2 <= h2)) return 0;
    return (m3Real)((double)Poly6_constant * pow((double)(h2 - r2), 3.0));
}
m3Real SPH_SM_monodomain::Spiky(m3Real r) {  // cpp:155-158
    if (!(r >= 0 && r <= kernel)) return 0;
    const m3Real d = kernel - r;
    return -Spiky_constant * d * d;
}
m3Real SPH_SM_monodomain::Visco(m3Real r) {  // cpp:161-164 (the Spiky constant, sic)
    return (r >= 0 && r <= kernel) ? Spiky_constant * (kernel - r) : 0;
}
m3Real SPH_SM_monodomain::B_spline(m3Real r) {  // cpp:166-175
    const m3Real q = r / kernel;
    if (q >= 0 && q < 1) return B_spline_constant * (1.0f - 1.5f * q * q + 0.75f * q * q * q);
    if (q >= 1 && q < 2) return (m3Real)((double)B_spline_constant * (0.25 * pow((double)(2 - q), 3.0)));
    return 0;
}
m3Real SPH_SM_monodomain::B_spline_1(m3Real r) {  // cpp:177-186
    const m3Real q = r / kernel;
    if (q >= 0 && q < 1) return B_spline_constant * (-3.0f * q + 2.25f * q * q);
    if (q >= 1 && q < 2) return (m3Real)((double)B_spline_constant * (-0.75 * pow((double)(2 - q), 2.0)));
    return 0;
}
m3Real SPH_SM_monodomain::B_spline_2(m3Real r) {  // cpp:188-197
    const m3Real q = r / kernel;
    if (q >= 0 && q < 1) return (m3Real)((double)B_spline_constant * (-3 + 4.5 * (double)q));
    if (q >= 1 && q < 2) return (m3Real)((double)B_spline_constant * (1.5 * (double)(2 - q)));
    return 0;
}

// ---- toggles ----------------------------------------------------------------------------------------------------------
void SPH_SM_monodomain::add_viscosity(float value) {  // cpp:87-91
    sphsm_params p;
    must(dev, sphsm_get_params(dev, &p), "sphsm_get_params");
    p.mu += (p.mu + value) >= 0 ? value : 0;
    must(dev, sphsm_set_params(dev, &p), "sphsm_set_params");
}
bool SPH_SM_monodomain::flip_quadratic() {  // h:154
    sphsm_params p;
    must(dev, sphsm_get_params(dev, &p), "sphsm_get_params");
    p.quadratic_match = !p.quadratic_match;
    must(dev, sphsm_set_params(dev, &p), "sphsm_set_params");
    return p.quadratic_match != 0;
}
bool SPH_SM_monodomain::flip_volume() {  // h:155
    sphsm_params p;
    must(dev, sphsm_get_params(dev, &p), "sphsm_get_params");
    p.volume_conservation = !p.volume_conservation;
    must(dev, sphsm_set_params(dev, &p), "sphsm_set_params");
    return p.volume_conservation != 0;
}

// ---- stimulation ------------------------------------------------------------------------------------------------------
void SPH_SM_monodomain::set_stim(m3Vector c, m3Real radius, m3Real strength) {  // cpp:704-717
    push_host_writes();
    must(dev, sphsm_set_stim(dev, c.x, c.y, c.z, radius, strength), "sphsm_set_stim");
    device_changed();
}
void SPH_SM_monodomain::turnOnStim_Cube(std::vector<m3Vector> positions) {  // cpp:719-743
    push_host_writes();
    const std::vector<float> xyz = flatten(positions);
    must(dev, sphsm_stim_cube(dev, xyz.data(), (int)positions.size()), "sphsm_stim_cube");
    device_changed();
}
void SPH_SM_monodomain::turnOnStim_Mesh(std::vector<m3Vector> positions) {  // cpp:745-762
    push_host_writes();
    const std::vector<float> xyz = flatten(positions);
    must(dev, sphsm_stim_mesh(dev, xyz.data(), (int)positions.size()), "sphsm_stim_mesh");
    device_changed();
}
void SPH_SM_monodomain::turnOffStim() {  // cpp:764-783
    push_host_writes();
    must(dev, sphsm_stim_off(dev), "sphsm_stim_off");
    device_changed();
}

// ---- the stages, individually callable as in the reference (h:120-143) ----------------------------------------------------
void SPH_SM_monodomain::run_stage(int stage) {
    push_host_writes();
    must(dev, sphsm_stage(dev, stage), "sphsm_stage");
    device_changed();
}
void SPH_SM_monodomain::Find_neighbors() { run_stage(SPHSM_STAGE_FIND_NEIGHBORS); }
void SPH_SM_monodomain::calculate_corrected_velocity() { run_stage(SPHSM_STAGE_CORRECTED_VELOCITY); }
void SPH_SM_monodomain::apply_external_forces(m3Vector *, int *, int) {
    // cpp:215-232: the external-force array is applied and then overwritten by the gravity loop (cpp:229), so only
    // gravity has an effect in the reference; the arguments are accepted for source compatibility.
    run_stage(SPHSM_STAGE_EXTERNAL_FORCES);
}
void SPH_SM_monodomain::projectPositions() { run_stage(SPHSM_STAGE_PROJECT_POSITIONS); }
void SPH_SM_monodomain::calculate_intermediate_velocity() { run_stage(SPHSM_STAGE_INTERMEDIATE_VELOCITY); }
void SPH_SM_monodomain::Compute_Density_SingPressure() { run_stage(SPHSM_STAGE_DENSITY_PRESSURE); }
void SPH_SM_monodomain::calculate_cell_model() { run_stage(SPHSM_STAGE_CELL_MODEL); }
void SPH_SM_monodomain::Compute_Force() { run_stage(SPHSM_STAGE_FORCE); }
void SPH_SM_monodomain::Update_Properties() { run_stage(SPHSM_STAGE_UPDATE); }

// ---- the step --------------------------------------------------------------------------------------------------------------
void SPH_SM_monodomain::collect_stage_times() {
    double t[7];
    must(dev, sphsm_get_stage_times(dev, t), "sphsm_get_stage_times");
    duration_d *slots[7] = {&d_find_neighbors,       &d_corrected_velocity, &d_intermediate_velocity, &d_Density_SingPressure,
                            &d_cell_model,           &d_compute_Force,      &d_Update_Properties};
    for (int k = 0; k < 7; k++) {
        *slots[k] += duration_d(t[k] - stage_seen[k]);
        stage_seen[k] = t[k];
    }
}

// Fused step: the kernel groups do not map one to one onto the reference's seven stages, so the sampled group times land in
// the slot of the stage that dominates each group: neighbour grid (hash + sort + gather) -> d_find_neighbors; shape-matching
// sums + solve (+ goal / corrected velocity, applied inside the gather) -> d_corrected_velocity; pass A (intermediate velocity
// AND density / pressure in one sweep) -> d_intermediate_velocity; pass B (cell model + force + Laplacian + integration)
// -> d_compute_Force; d_Density_SingPressure, d_cell_model and d_Update_Properties stay zero.
static const int SAMPLE_PERIOD = 50;
void SPH_SM_monodomain::Animation(int nsteps) {
    push_host_writes();
    if (stage_timing) {
        must(dev, sphsm_step(dev, nsteps), "sphsm_step");
        total_time_steps += nsteps;
        device_changed();
        collect_stage_times();
        return;
    }
    int left = nsteps;
    while (left > 0) {
        if (total_time_steps % SAMPLE_PERIOD == 0) {  // a sampled step: per-group CUDA events, weighted for the whole period
            float ms[SPHSM_NUM_KERNEL_GROUPS];
            must(dev, sphsm_profile_step(dev, 1, ms), "sphsm_profile_step");
            const double w = 1e-3 * SAMPLE_PERIOD;
            d_find_neighbors += duration_d(w * (ms[0] + ms[1] + ms[2]));
            d_corrected_velocity += duration_d(w * (ms[3] + ms[4]));
            d_intermediate_velocity += duration_d(w * ms[5]);
            d_compute_Force += duration_d(w * ms[6]);
            total_time_steps += 1;
            left -= 1;
            continue;
        }
        const int run = std::min(left, SAMPLE_PERIOD - total_time_steps % SAMPLE_PERIOD);
        must(dev, sphsm_step(dev, run), "sphsm_step");
        total_time_steps += run;
        left -= run;
    }
    device_changed();
}
void SPH_SM_monodomain::compute_SPH_SM_monodomain() { Animation(1); }  // cpp:794-824
void SPH_SM_monodomain::Animation() { Animation(1); }                  // cpp:826-829

void SPH_SM_monodomain::print_report(double avg_fps, double avg_step_d) {  // the 23-field line of cpp:785-792
    sphsm_params p;
    must(dev, sphsm_get_params(dev, &p), "sphsm_get_params");
    // (sampled timers cover whole periods: normalise by the steps they stand for)
    const double n = stage_timing ? (double)total_time_steps : (double)(((total_time_steps + SAMPLE_PERIOD - 1) / SAMPLE_PERIOD) * SAMPLE_PERIOD);
    cout << avg_fps << ";" << avg_step_d << ";" << total_time_steps;
    const duration_d *slots[7] = {&d_find_neighbors,       &d_corrected_velocity, &d_intermediate_velocity, &d_Density_SingPressure,
                                  &d_cell_model,           &d_compute_Force,      &d_Update_Properties};
    for (int k = 0; k < 7; k++) cout << ";" << slots[k]->count() / n;
    cout << ";" << p.K << ";" << p.alpha << ";" << p.beta << ";" << p.mu << ";" << p.sigma << ";" << p.stim_strength << ";" << p.FH_Vt << ";"
         << p.FH_Vp << ";" << p.FH_Vr << ";" << p.C1 << ";" << p.C2 << ";" << p.C3 << ";" << p.C4 << endl;
}
