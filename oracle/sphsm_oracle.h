/* TEST INFRASTRUCTURE ONLY — the CPU oracle for the SPH_SM_monodomain per-timestep pipeline.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product path (sph_sm_monodomain_b200/csrc, include/) never does.
 *
 * A plain-C restatement of the reference's algorithm, every function citing the reference
 * file:line it follows ("cpp" = SPH_SM_monodomain/SPH_SM_monodomain.cpp, paths relative to the
 * reference root).  PARITY IS PINNED: the reference itself holds no tests/golden vectors
 * (SURVEY.md §4), so the pin is (a) bit-identity with the genuine reference class compiled here
 * from its own sources (oracle/_ref, tests/test_oracle_vs_ref.py) and (b) the committed
 * tests/golden/ fixtures that tools/make_golden.py generated from that genuine build.
 *
 * Build with `-O2 -ffp-contract=off` (no FMA contraction, no fast-math): the reference's x86-64
 * g++ build evaluates float expressions as separate IEEE mul/add and this file relies on that.
 */
#ifndef SPHSM_ORACLE_H
#define SPHSM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order, sizes and padding as the reference's `Particle` (Particle.h:7-35): 132 bytes. */
typedef struct {
    float pos[3], vel[3], predicted_vel[3], inter_vel[3], corrected_vel[3], acc[3];
    float mass;
    float orig[3], goal[3];
    unsigned char fixed, _pad[3];
    float dens, pres, Vm, Inter_Vm, Iion, stim, w;
} ora_particle;

typedef struct ora_sim ora_sim;

/* world = World_Size (cpp:29); capacity = Max_Number_Paticles (cpp:19). The default reference
 * object is ora_create(50000, 1.5f, 1.5f, 1.5f). */
ora_sim *ora_create(int capacity, float wx, float wy, float wz);
void ora_destroy(ora_sim *s);

/* moments_in_double != 0: accumulate the shape-matching sums (cpp:244-292, 343-386) in double and
 * round once (the oracle of record for N >= 1e5, SURVEY.md §7 hard part 3); 0 = the reference's
 * sequential float accumulation (bit-identical to the genuine class). */
void ora_set_moments_in_double(ora_sim *s, int on);

void ora_init_fluid(ora_sim *s, const float *xyz, int n);                        /* cpp:93-125 */
void ora_set_stim(ora_sim *s, float cx, float cy, float cz, float radius, float strength); /* cpp:704-717 */
void ora_stim_mesh(ora_sim *s, const float *xyz, int n);                         /* cpp:745-762 */
void ora_stim_cube(ora_sim *s, const float *xyz, int n);                         /* cpp:719-743 */
void ora_stim_off(ora_sim *s);                                                   /* cpp:764-783 */
int ora_flip_quadratic(ora_sim *s);                                              /* h:154 */
int ora_flip_volume(ora_sim *s);                                                 /* h:155 */
void ora_add_viscosity(ora_sim *s, float v);                                     /* cpp:87-91 */

int ora_n(ora_sim *s);
ora_particle *ora_particles(ora_sim *s);
int ora_num_cells(ora_sim *s);
void ora_constants(ora_sim *s, float *out16); /* same order as ref_harness.cpp:ref_constants */

/* stage: 0 whole step, 1 Find_neighbors, 2 calculate_corrected_velocity, 3 calculate_intermediate_velocity,
 * 4 Compute_Density_SingPressure, 5 calculate_cell_model, 6 Compute_Force, 7 Update_Properties (cpp:794-824) */
void ora_stage(ora_sim *s, int stage);
void ora_steps(ora_sim *s, int n);

int ora_cells_csr(ora_sim *s, int *cell_start, int *indices);
int ora_cell_hash(ora_sim *s, float x, float y, float z);

/* Neighbour sets of particle i in the reference's visiting order (27 cells dk,dj,di; bucket order):
 * kind 0 = candidate set C(i); 1 = {j : r2 <= h*h} (Poly6 support, cpp:151); 2 = {j : r2 > 1e-12f && r <= h}
 * (Spiky/Visco support, cpp:546,157); 3 = {j : r2 > 1e-12f && r/h < 2} (B_spline_2 support, cpp:190-196).
 * Returns the count; writes up to cap indices. Requires Find_neighbors to have run. */
int ora_neighbors(ora_sim *s, int i, int kind, int *out, int cap);

/* shape-matching internals of the last projectPositions call: cm[3], ocm[3], T or A9 (27 floats, the
 * linear T occupies [0..8]) */
void ora_sm_debug(ora_sim *s, float *cm3, float *ocm3, float *xform27);

float ora_poly6(ora_sim *s, float r2);
float ora_spiky(ora_sim *s, float r);
float ora_visco(ora_sim *s, float r);
float ora_bspline2(ora_sim *s, float r);
void ora_polar3(const float *a9, float *r9);
int ora_invert3(float *a9);
void ora_invert9(float *a81);

#ifdef __cplusplus
}
#endif
#endif
