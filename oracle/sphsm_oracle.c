/* TEST INFRASTRUCTURE ONLY — see sphsm_oracle.h for the rules on who may load this.
 *
 * Plain-C restatement of the reference's per-timestep pipeline.  "cpp" below means
 * SPH_SM_monodomain/SPH_SM_monodomain.cpp of the reference; every block cites the lines it follows.
 * Float/double promotions are restated exactly as g++ resolves them in the reference (SURVEY.md
 * §3.6 Q16): arithmetic is IEEE float with separate mul/add (build with -ffp-contract=off) except
 * where a double literal or pow(float,int) promotes an expression, which is written out here with
 * explicit (double) casts.
 */
#include "sphsm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORA_PI 3.1415926535897932f /* m3Pi, Math3D/m3Real.h:9 */
#define ORA_INF 1E-12f             /* INF, SPH_SM_monodomain.h:24 */

typedef struct { float x, y, z; } v3;

struct ora_sim {
    /* cpp:13-69 / SPH_SM_monodomain.h:33-94 */
    float kernel, cell_size;
    int capacity, n;
    v3 grid, world;
    int num_cells;
    v3 gravity;
    float K, stand_density, dt, wall_hit, mu, velocity_mixing;
    float poly6_c, spiky_c, bspline_c;
    v3 bmin, bmax;
    float alpha, beta;
    int quadratic, volume_conservation, allow_flip;
    float Cm, Beta, sigma, stim_strength;
    float FH_Vt, FH_Vp, FH_Vr, C1, C2, C3, C4;
    float voltage_constant, max_pressure, max_voltage;
    int moments_in_double;

    ora_particle *p;
    /* buckets (Cells[], cpp:52,199-213) as CSR rebuilt by stage 1; order inside a bucket = ascending index */
    int *cell_start; /* num_cells + 1 */
    int *cell_items; /* capacity */
    int *cell_of;    /* scratch, capacity */

    float dbg_cm[3], dbg_ocm[3], dbg_xform[27];
};

/* ---------------------------------------------------------------- construction: cpp:13-79 */
ora_sim *ora_create(int capacity, float wx, float wy, float wz) {
    ora_sim *s = (ora_sim *)calloc(1, sizeof(ora_sim));
    float sigma_i = 0.893, sigma_e = 0.67;                   /* cpp:15 */
    s->kernel = 0.04f;                                       /* cpp:17 */
    s->capacity = capacity;                                  /* cpp:19 */
    s->n = 0;
    s->Cm = 1.f;                                             /* cpp:23 */
    s->Beta = 50;                                            /* cpp:24 */
    s->sigma = sigma_i * sigma_e / (sigma_i + sigma_e);      /* cpp:26 */
    s->stim_strength = 300.0f;                               /* cpp:27 */
    s->world.x = wx; s->world.y = wy; s->world.z = wz;       /* cpp:29 */
    s->cell_size = 0.04;                                     /* cpp:31 */
    s->grid.x = (int)ceilf(s->world.x / s->cell_size);       /* cpp:32-35 */
    s->grid.y = (int)ceilf(s->world.y / s->cell_size);
    s->grid.z = (int)ceilf(s->world.z / s->cell_size);
    s->num_cells = (int)s->grid.x * (int)s->grid.y * (int)s->grid.z; /* cpp:37 */
    s->gravity.x = 0.0f; s->gravity.y = -9.8f; s->gravity.z = 0.0f;  /* cpp:39 */
    s->K = 0.5f;                                             /* cpp:40 */
    s->stand_density = 1112.0f;                              /* cpp:41 */
    s->velocity_mixing = 1.0f;                               /* cpp:43 */
    {   /* cpp:42,47: max_vel=(3,3,3); Time_Delta = 0.4 * kernel / sqrt(magnitudeSquared) — double
           literal 0.4, float sqrt (std::sqrt(float) via <math.h> + using namespace std) */
        float mv2 = 3.0f * 3.0f + 3.0f * 3.0f + 3.0f * 3.0f;
        s->dt = (float)(0.4 * (double)s->kernel / (double)sqrtf(mv2));
    }
    s->wall_hit = -1.0f;                                     /* cpp:48 */
    s->mu = 100.0f;                                          /* cpp:49 */
    /* cpp:54-55: pow(float,int) is the double pow */
    s->poly6_c = (float)((double)315.0f / ((double)(64.0f * ORA_PI) * pow((double)s->kernel, 9.0)));
    s->spiky_c = (float)((double)45.0f / ((double)ORA_PI * pow((double)s->kernel, 6.0)));
    s->bspline_c = 1.0f / (ORA_PI * s->kernel * s->kernel * s->kernel); /* cpp:57 */
    s->bmin.x = s->bmin.y = s->bmin.z = 0.0f;                /* cpp:60-61 */
    s->bmax = s->world;
    s->alpha = 0.3f; s->beta = 0.4f;                         /* cpp:64-65 */
    s->quadratic = 0; s->volume_conservation = 1; s->allow_flip = 0; /* cpp:67-69 */
    s->FH_Vt = -75.0; s->FH_Vp = 15.0; s->FH_Vr = -85.0;     /* h:72-74 */
    s->C1 = 0.175; s->C2 = 0.03; s->C3 = 0.011; s->C4 = 0.55; /* h:76-80 */
    s->voltage_constant = 1; s->max_pressure = 15000; s->max_voltage = 200; /* h:92-94 */
    s->p = (ora_particle *)calloc((size_t)capacity, sizeof(ora_particle));
    s->cell_start = (int *)calloc((size_t)s->num_cells + 1, sizeof(int));
    s->cell_items = (int *)calloc((size_t)capacity, sizeof(int));
    s->cell_of = (int *)calloc((size_t)capacity, sizeof(int));
    return s;
}

void ora_destroy(ora_sim *s) {
    if (!s) return;
    free(s->p); free(s->cell_start); free(s->cell_items); free(s->cell_of); free(s);
}

void ora_set_moments_in_double(ora_sim *s, int on) { s->moments_in_double = on; }
int ora_n(ora_sim *s) { return s->n; }
ora_particle *ora_particles(ora_sim *s) { return s->p; }
int ora_num_cells(ora_sim *s) { return s->num_cells; }
int ora_flip_quadratic(ora_sim *s) { s->quadratic = !s->quadratic; return s->quadratic; }
int ora_flip_volume(ora_sim *s) { s->volume_conservation = !s->volume_conservation; return s->volume_conservation; }
void ora_add_viscosity(ora_sim *s, float v) { s->mu += (s->mu + v) >= 0 ? v : 0; } /* cpp:87-91 */

void ora_constants(ora_sim *s, float *o) {
    o[0] = s->K; o[1] = s->stand_density; o[2] = s->dt; o[3] = s->mu; o[4] = s->poly6_c; o[5] = s->spiky_c;
    o[6] = s->bspline_c; o[7] = s->sigma; o[8] = s->alpha; o[9] = s->beta; o[10] = s->kernel;
    o[11] = s->stim_strength; o[12] = s->velocity_mixing; o[13] = s->wall_hit; o[14] = s->Cm; o[15] = s->Beta;
}

/* ---------------------------------------------------------------- init: cpp:93-125 */
void ora_init_fluid(ora_sim *s, const float *xyz, int n) {
    for (int i = 0; i < n; i++) {
        if (s->n + 1 > s->capacity) continue;                /* cpp:103 (silently dropped) */
        ora_particle *p = &s->p[s->n];
        memset(p, 0, sizeof(*p));                            /* m3Vector() zero-initialises, Particle fields */
        for (int a = 0; a < 3; a++) p->pos[a] = p->orig[a] = p->goal[a] = xyz[3 * i + a];
        p->fixed = 0;
        p->dens = s->stand_density;                          /* cpp:115 */
        p->mass = 0.2f;                                      /* cpp:116 */
        s->n++;
    }
}

/* cpp:704-717 — note: squared distance compared with `radius` (Q11) */
void ora_set_stim(ora_sim *s, float cx, float cy, float cz, float radius, float strength) {
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        float dx = p->pos[0] - cx, dy = p->pos[1] - cy, dz = p->pos[2] - cz;
        if ((dx * dx + dy * dy + dz * dz) <= radius) p->stim = strength;
    }
}

/* cpp:745-762 — double-literal comparisons */
void ora_stim_mesh(ora_sim *s, const float *xyz, int n) {
    for (int i = 0; i < n; i++) ora_set_stim(s, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.01f, s->stim_strength);
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        double x = p->pos[0], y = p->pos[1];
        if ((x >= 0.0 && x <= 0.07) || (x >= 0.90 && y >= 0.80)) p->fixed = 1;
    }
}

/* cpp:719-743 */
void ora_stim_cube(ora_sim *s, const float *xyz, int n) {
    for (int i = 0; i < n; i++) {
        float px = xyz[3 * i], pz = xyz[3 * i + 2];
        if (((double)px >= 0.45 && (double)px <= 0.48) || ((double)px > 1.0 && pz <= 1.05f))
            ora_set_stim(s, px, xyz[3 * i + 1], pz, 0.001f, s->stim_strength);
    }
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        if ((p->pos[1] == 0.0f && p->pos[0] <= 0.48f) || (p->pos[1] == 0.0f && (double)p->pos[0] >= 1.0)) p->fixed = 1;
    }
}

/* cpp:764-783 */
void ora_stim_off(ora_sim *s) {
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        p->stim = -10000.0f; p->Vm = 0.0f; p->Inter_Vm = 0.0f; p->Iion = 0.0f; p->pres = -10000.0f; p->w = 0.0f;
    }
}

/* ---------------------------------------------------------------- grid: cpp:127-146 */
static v3 cell_position(const ora_sim *s, const float *pos) {
    v3 c;
    c.x = (int)(pos[0] / s->cell_size);
    c.y = (int)(pos[1] / s->cell_size);
    c.z = (int)(pos[2] / s->cell_size);
    return c;
}
static int cell_hash(const ora_sim *s, v3 c) {
    if ((c.x < 0) || (c.x >= s->grid.x) || (c.y < 0) || (c.y >= s->grid.y) || (c.z < 0) || (c.z >= s->grid.z)) return -1;
    return (int)(c.x + s->grid.x * (c.y + s->grid.y * c.z)); /* evaluated in float, cpp:142 */
}
int ora_cell_hash(ora_sim *s, float x, float y, float z) {
    float pos[3] = {x, y, z};
    return cell_hash(s, cell_position(s, pos));
}

/* ---------------------------------------------------------------- kernels: cpp:149-164,188-197 */
static float poly6(const ora_sim *s, float r2) {
    float h2 = s->kernel * s->kernel;
    if (r2 >= 0 && r2 <= h2) return (float)((double)s->poly6_c * pow((double)(h2 - r2), 3.0));
    return 0;
}
static float spiky(const ora_sim *s, float r) {
    return (r >= 0 && r <= s->kernel) ? -s->spiky_c * (s->kernel - r) * (s->kernel - r) : 0;
}
static float visco(const ora_sim *s, float r) {
    return (r >= 0 && r <= s->kernel) ? s->spiky_c * (s->kernel - r) : 0;
}
static float bspline2(const ora_sim *s, float r) {
    float q = r / s->kernel;
    if (q >= 0 && q < 1) return (float)((double)s->bspline_c * (-3 + 4.5 * (double)q));
    else if (q >= 1 && q < 2) return (float)((double)s->bspline_c * (1.5 * (double)(2 - q)));
    return 0;
}
float ora_poly6(ora_sim *s, float r2) { return poly6(s, r2); }
float ora_spiky(ora_sim *s, float r) { return spiky(s, r); }
float ora_visco(ora_sim *s, float r) { return visco(s, r); }
float ora_bspline2(ora_sim *s, float r) { return bspline2(s, r); }

/* ---------------------------------------------------------------- stage 1: cpp:199-213 */
static void find_neighbors(ora_sim *s) {
    int nc = s->num_cells;
    memset(s->cell_start, 0, sizeof(int) * ((size_t)nc + 1));
    for (int i = 0; i < s->n; i++) {
        int h = cell_hash(s, cell_position(s, s->p[i].pos));
        /* the reference indexes Cells[-1] here (UB, Q15); unreachable while stage 7 clamps. Park such
           particles in no bucket. */
        s->cell_of[i] = h;
        if (h >= 0 && h < nc) s->cell_start[h + 1]++;
    }
    for (int c = 0; c < nc; c++) s->cell_start[c + 1] += s->cell_start[c];
    /* stable fill: bucket order = ascending particle index, as push_back yields */
    int *cursor = (int *)malloc(sizeof(int) * (size_t)nc);
    memcpy(cursor, s->cell_start, sizeof(int) * (size_t)nc);
    for (int i = 0; i < s->n; i++) {
        int h = s->cell_of[i];
        if (h >= 0 && h < nc) s->cell_items[cursor[h]++] = i;
    }
    free(cursor);
}

int ora_cells_csr(ora_sim *s, int *cell_start, int *indices) {
    int tot = s->cell_start[s->num_cells];
    if (cell_start) memcpy(cell_start, s->cell_start, sizeof(int) * ((size_t)s->num_cells + 1));
    if (indices) memcpy(indices, s->cell_items, sizeof(int) * (size_t)tot);
    return tot;
}

/* iterate the 27-cell stencil in the reference's order (k outer, i inner; cpp:462-464) */
#define FOR_CANDIDATES(s, P, J, BODY)                                                  \
    do {                                                                               \
        v3 cp_ = cell_position((s), (P)->pos);                                         \
        for (int kk_ = -1; kk_ <= 1; kk_++)                                            \
            for (int jj_ = -1; jj_ <= 1; jj_++)                                        \
                for (int ii_ = -1; ii_ <= 1; ii_++) {                                  \
                    v3 np_;                                                            \
                    np_.x = cp_.x + (float)ii_; np_.y = cp_.y + (float)jj_; np_.z = cp_.z + (float)kk_; \
                    int h_ = cell_hash((s), np_);                                      \
                    if (h_ == -1) continue;                                            \
                    for (int e_ = (s)->cell_start[h_]; e_ < (s)->cell_start[h_ + 1]; e_++) { \
                        int J = (s)->cell_items[e_];                                   \
                        BODY                                                           \
                    }                                                                  \
                }                                                                      \
    } while (0)

static inline float dist2(const ora_particle *a, const ora_particle *b, float *d) {
    d[0] = a->pos[0] - b->pos[0]; d[1] = a->pos[1] - b->pos[1]; d[2] = a->pos[2] - b->pos[2];
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2]; /* m3Vector::magnitudeSquared, m3Vector.h:93 */
}

int ora_neighbors(ora_sim *s, int i, int kind, int *out, int cap) {
    int cnt = 0;
    const ora_particle *p = &s->p[i];
    float h2 = s->kernel * s->kernel;
    FOR_CANDIDATES(s, p, j, {
        float d[3];
        float r2 = dist2(p, &s->p[j], d);
        int in;
        if (kind == 0) in = 1;
        else if (kind == 1) in = (r2 >= 0 && r2 <= h2);
        else if (kind == 2) { float r = sqrtf(r2); in = (r2 > ORA_INF) && (r >= 0 && r <= s->kernel); }
        else { float q = sqrtf(r2) / s->kernel; in = (r2 > ORA_INF) && (q >= 0 && q < 2); }
        if (in) { if (cnt < cap && out) out[cnt] = j; cnt++; }
    });
    return cnt;
}

/* ---------------------------------------------------------------- 3x3 math (Math3D/m3Matrix.{h,cpp}) */
#define M(a, i, j) (a)[(i) * 3 + (j)]

static float det3(const float *m) { /* m3Matrix.h:288-291 */
    return M(m,0,0) * (M(m,1,1) * M(m,2,2) - M(m,2,1) * M(m,1,2)) - M(m,0,1) * (M(m,1,0) * M(m,2,2) - M(m,2,0) * M(m,1,2)) +
           M(m,0,2) * (M(m,1,0) * M(m,2,1) - M(m,1,1) * M(m,2,0));
}
static int invert3(float *m) { /* m3Matrix.h:293-318 */
    float d = det3(m);
    if (d == 0.0) return 0;
    d = (float)1.0 / d;
    float r[9];
    r[0] = (M(m,1,1) * M(m,2,2) - M(m,1,2) * M(m,2,1)) * d;
    r[1] = -(M(m,0,1) * M(m,2,2) - M(m,0,2) * M(m,2,1)) * d;
    r[2] = (M(m,0,1) * M(m,1,2) - M(m,0,2) * M(m,1,1)) * d;
    r[3] = -(M(m,1,0) * M(m,2,2) - M(m,1,2) * M(m,2,0)) * d;
    r[4] = (M(m,0,0) * M(m,2,2) - M(m,0,2) * M(m,2,0)) * d;
    r[5] = -(M(m,0,0) * M(m,1,2) - M(m,0,2) * M(m,1,0)) * d;
    r[6] = (M(m,1,0) * M(m,2,1) - M(m,1,1) * M(m,2,0)) * d;
    r[7] = -(M(m,0,0) * M(m,2,1) - M(m,0,1) * M(m,2,0)) * d;
    r[8] = (M(m,0,0) * M(m,1,1) - M(m,0,1) * M(m,1,0)) * d;
    memcpy(m, r, sizeof(r));
    return 1;
}
static void mul3(float *out, const float *l, const float *r) { /* m3Matrix.h:223-239 */
    float t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[i * 3 + j] = M(l,i,0) * M(r,0,j) + M(l,i,1) * M(r,1,j) + M(l,i,2) * M(r,2,j);
    memcpy(out, t, sizeof(t));
}
static void mul3_tl(float *out, const float *l, const float *r) { /* multiplyTransposedLeft, m3Matrix.h:241-257 */
    float t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[i * 3 + j] = M(l,0,i) * M(r,0,j) + M(l,1,i) * M(r,1,j) + M(l,2,i) * M(r,2,j);
    memcpy(out, t, sizeof(t));
}

/* Jacobi rotation shared by the 3x3 and 9x9 code (m3Matrix.cpp:3-35, m9Matrix.cpp:10-43). In those
   two files unqualified fabs/sqrt on floats are the DOUBLE C functions (Q16). */
static void jacobi_rotate(float *A, float *R, int n, int p, int q) {
#define E(a, i, j) (a)[(i) * n + (j)]
    float d = (E(A,p,p) - E(A,q,q)) / (2.0f * E(A,p,q));
    float t = (float)((double)1.0f / (fabs((double)d) + sqrt((double)(d * d + 1.0f))));
    if (d < 0.0f) t = -t;
    float c = (float)((double)1.0f / sqrt((double)(t * t + 1)));
    float sn = t * c;
    E(A,p,p) += t * E(A,p,q);
    E(A,q,q) -= t * E(A,p,q);
    E(A,p,q) = E(A,q,p) = 0.0f;
    for (int k = 0; k < n; k++) {
        if (k != p && k != q) {
            float Akp = c * E(A,k,p) + sn * E(A,k,q);
            float Akq = -sn * E(A,k,p) + c * E(A,k,q);
            E(A,k,p) = E(A,p,k) = Akp;
            E(A,k,q) = E(A,q,k) = Akq;
        }
    }
    for (int k = 0; k < n; k++) {
        float Rkp = c * E(R,k,p) + sn * E(R,k,q);
        float Rkq = -sn * E(R,k,p) + c * E(R,k,q);
        E(R,k,p) = Rkp;
        E(R,k,q) = Rkq;
    }
}
/* m3Matrix.cpp:38-70 / m9Matrix.cpp:47-76: at most 20 rotations, pivot = first max |off-diagonal| */
static void eigen_decomposition(float *A, float *R, int n) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) E(R,i,j) = (i == j) ? 1.0f : 0.0f;
    int iter = 0;
    while (iter < 20) {
        int p = 0, q = 0;
        float a, max = -1.0f;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                a = (float)fabs((double)E(A,i,j));
                if (max < 0.0f || a > max) { p = i; q = j; max = a; }
            }
        if (max <= 0.0f) break;
        jacobi_rotate(A, R, n, p, q);
        iter++;
    }
#undef E
}
static void polar3(const float *A, float *R) { /* m3Matrix.cpp:73-113 (S is computed there but never used) */
    float ATA[9], U[9], S1[9];
    mul3_tl(ATA, A, A);
    eigen_decomposition(ATA, U, 3);
    float l0 = M(ATA,0,0); if (l0 <= 0.0f) l0 = 0.0f; else l0 = (float)((double)1.0f / sqrt((double)l0));
    float l1 = M(ATA,1,1); if (l1 <= 0.0f) l1 = 0.0f; else l1 = (float)((double)1.0f / sqrt((double)l1));
    float l2 = M(ATA,2,2); if (l2 <= 0.0f) l2 = 0.0f; else l2 = (float)((double)1.0f / sqrt((double)l2));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            S1[i * 3 + j] = l0 * M(U,i,0) * M(U,j,0) + l1 * M(U,i,1) * M(U,j,1) + l2 * M(U,i,2) * M(U,j,2);
    mul3(R, A, S1);
}
static void invert9(float *m) { /* m9Matrix.cpp:80-102 */
    float A[81], R[81], d[9];
    memcpy(A, m, sizeof(A));
    eigen_decomposition(A, R, 9);
    for (int i = 0; i < 9; i++) { d[i] = A[i * 9 + i]; if (d[i] != 0.0f) d[i] = 1.0f / d[i]; }
    for (int i = 0; i < 9; i++)
        for (int j = 0; j < 9; j++) {
            float a = 0.0f;
            for (int k = 0; k < 9; k++) a += d[k] * R[i * 9 + k] * R[j * 9 + k];
            m[i * 9 + j] = a;
        }
}
void ora_polar3(const float *a9, float *r9) { polar3(a9, r9); }
int ora_invert3(float *a9) { return invert3(a9); }
void ora_invert9(float *a81) { invert9(a81); }

/* ---------------------------------------------------------------- stage 2: cpp:215-446, 653-667 */
static void apply_external_forces(ora_sim *s) { /* cpp:226-231; the force-array loop is dead (size 0) */
    for (int i = 0; i < s->n; i++) {
        ora_particle *p = &s->p[i];
        if (p->fixed) continue;
        p->predicted_vel[0] = p->vel[0] + (s->gravity.x * s->dt) / p->mass;
        p->predicted_vel[1] = p->vel[1] + (s->gravity.y * s->dt) / p->mass;
        p->predicted_vel[2] = p->vel[2] + (s->gravity.z * s->dt) / p->mass;
    }
}

static void project_positions(ora_sim *s) {
    int n = s->n;
    if (n <= 1) return; /* cpp:236 */
    float cm[3], ocm[3], Apq[9], Aqq[9];
    const int dbl = s->moments_in_double;

    /* centres, cpp:240-254 */
    if (!dbl) {
        float mass = 0.0f;
        cm[0] = cm[1] = cm[2] = ocm[0] = ocm[1] = ocm[2] = 0.0f;
        for (int i = 0; i < n; i++) {
            const ora_particle *p = &s->p[i];
            float m = p->mass;
            if (p->fixed) m *= 100.0f;
            mass += m;
            for (int a = 0; a < 3; a++) { cm[a] += p->pos[a] * m; ocm[a] += p->orig[a] * m; }
        }
        for (int a = 0; a < 3; a++) { cm[a] /= mass; ocm[a] /= mass; }
    } else {
        double mass = 0.0, c[3] = {0, 0, 0}, o[3] = {0, 0, 0};
        for (int i = 0; i < n; i++) {
            const ora_particle *p = &s->p[i];
            float m = p->mass;
            if (p->fixed) m *= 100.0f;
            mass += m;
            for (int a = 0; a < 3; a++) { c[a] += (double)(p->pos[a] * m); o[a] += (double)(p->orig[a] * m); }
        }
        float fm = (float)mass;
        for (int a = 0; a < 3; a++) { cm[a] = (float)c[a] / fm; ocm[a] = (float)o[a] / fm; }
    }

    /* moments, cpp:256-292 */
    if (!dbl) {
        for (int e = 0; e < 9; e++) Apq[e] = Aqq[e] = 0.0f;
        for (int i = 0; i < n; i++) {
            const ora_particle *pt = &s->p[i];
            float p[3], q[3];
            for (int a = 0; a < 3; a++) { p[a] = pt->pos[a] - cm[a]; q[a] = pt->orig[a] - ocm[a]; }
            float m = pt->mass;
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) { M(Apq,a,b) += m * p[a] * q[b]; M(Aqq,a,b) += m * q[a] * q[b]; }
        }
    } else {
        double dpq[9] = {0}, dqq[9] = {0};
        for (int i = 0; i < n; i++) {
            const ora_particle *pt = &s->p[i];
            float p[3], q[3];
            for (int a = 0; a < 3; a++) { p[a] = pt->pos[a] - cm[a]; q[a] = pt->orig[a] - ocm[a]; }
            float m = pt->mass;
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) { dpq[a * 3 + b] += (double)(m * p[a] * q[b]); dqq[a * 3 + b] += (double)(m * q[a] * q[b]); }
        }
        for (int e = 0; e < 9; e++) { Apq[e] = (float)dpq[e]; Aqq[e] = (float)dqq[e]; }
    }

    if (!s->allow_flip && det3(Apq) < 0.0f) { /* cpp:294-299 (sic: r01, r11, r22) */
        M(Apq,0,1) = -M(Apq,0,1); M(Apq,1,1) = -M(Apq,1,1); M(Apq,2,2) = -M(Apq,2,2);
    }

    float R[9];
    polar3(Apq, R); /* cpp:301-302 */

    for (int a = 0; a < 3; a++) { s->dbg_cm[a] = cm[a]; s->dbg_ocm[a] = ocm[a]; }
    memset(s->dbg_xform, 0, sizeof(s->dbg_xform));

    if (!s->quadratic) { /* cpp:304-330 */
        float A[9], T[9];
        memcpy(A, Aqq, sizeof(A));
        invert3(A);
        mul3(A, Apq, A);
        if (s->volume_conservation) {
            float det = det3(A);
            if (det != 0.0f) {
                det = 1.0f / sqrtf(fabsf(det)); /* float overloads in SPH_SM_monodomain.cpp (Q16) */
                if (det > 2.0f) det = 2.0f;
                for (int e = 0; e < 9; e++) A[e] *= det;
            }
        }
        float omb = 1.0f - s->beta;
        for (int e = 0; e < 9; e++) T[e] = R[e] * omb + A[e] * s->beta;
        memcpy(s->dbg_xform, T, sizeof(T));
        for (int i = 0; i < n; i++) {
            ora_particle *pt = &s->p[i];
            if (pt->fixed) continue;
            float q[3];
            for (int a = 0; a < 3; a++) q[a] = pt->orig[a] - ocm[a];
            for (int a = 0; a < 3; a++) pt->goal[a] = (M(T,a,0) * q[0] + M(T,a,1) * q[1] + M(T,a,2) * q[2]) + cm[a];
        }
    } else { /* cpp:331-445 */
        float A9pq[3][9], A9qq[81], A9[3][9];
        if (!dbl) {
            memset(A9pq, 0, sizeof(A9pq));
            memset(A9qq, 0, sizeof(A9qq));
            for (int i = 0; i < n; i++) {
                const ora_particle *pt = &s->p[i];
                float p[3], q[3], q9[9];
                for (int a = 0; a < 3; a++) { p[a] = pt->pos[a] - cm[a]; q[a] = pt->orig[a] - ocm[a]; }
                q9[0] = q[0]; q9[1] = q[1]; q9[2] = q[2]; q9[3] = q[0] * q[0]; q9[4] = q[1] * q[1]; q9[5] = q[2] * q[2];
                q9[6] = q[0] * q[1]; q9[7] = q[1] * q[2]; q9[8] = q[2] * q[0];
                float m = pt->mass;
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 9; b++) A9pq[a][b] += m * p[a] * q9[b];
                for (int j = 0; j < 9; j++)
                    for (int k = 0; k < 9; k++) A9qq[j * 9 + k] += m * q9[j] * q9[k];
            }
        } else {
            double dpq[27] = {0}, dqq[81] = {0};
            for (int i = 0; i < n; i++) {
                const ora_particle *pt = &s->p[i];
                float p[3], q[3], q9[9];
                for (int a = 0; a < 3; a++) { p[a] = pt->pos[a] - cm[a]; q[a] = pt->orig[a] - ocm[a]; }
                q9[0] = q[0]; q9[1] = q[1]; q9[2] = q[2]; q9[3] = q[0] * q[0]; q9[4] = q[1] * q[1]; q9[5] = q[2] * q[2];
                q9[6] = q[0] * q[1]; q9[7] = q[1] * q[2]; q9[8] = q[2] * q[0];
                float m = pt->mass;
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 9; b++) dpq[a * 9 + b] += (double)(m * p[a] * q9[b]);
                for (int j = 0; j < 9; j++)
                    for (int k = 0; k < 9; k++) dqq[j * 9 + k] += (double)(m * q9[j] * q9[k]);
            }
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 9; b++) A9pq[a][b] = (float)dpq[a * 9 + b];
            for (int e = 0; e < 81; e++) A9qq[e] = (float)dqq[e];
        }
        invert9(A9qq); /* cpp:388 */
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 9; j++) { /* cpp:391-403 */
                A9[i][j] = 0.0f;
                for (int k = 0; k < 9; k++) A9[i][j] += A9pq[i][k] * A9qq[k * 9 + j];
                A9[i][j] *= s->beta;
                if (j < 3) A9[i][j] += (1.0f - s->beta) * M(R,i,j);
            }
        float det = A9[0][0] * (A9[1][1] * A9[2][2] - A9[2][1] * A9[1][2]) - A9[0][1] * (A9[1][0] * A9[2][2] - A9[2][0] * A9[1][2]) +
                    A9[0][2] * (A9[1][0] * A9[2][1] - A9[1][1] * A9[2][0]); /* cpp:405-408 */
        if (!s->allow_flip && det < 0.0f) { A9[0][1] = -A9[0][1]; A9[1][1] = -A9[1][1]; A9[2][2] = -A9[2][2]; } /* cpp:410-414 */
        if (s->volume_conservation && det != 0.0f) { /* cpp:416-427 (pre-flip det) */
            det = 1.0f / sqrtf(fabsf(det));
            if (det > 2.0f) det = 2.0f;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 9; j++) A9[i][j] *= det;
        }
        memcpy(s->dbg_xform, A9, sizeof(A9));
        for (int i = 0; i < n; i++) { /* cpp:429-444 */
            ora_particle *pt = &s->p[i];
            if (pt->fixed) continue;
            float qx = pt->orig[0] - ocm[0], qy = pt->orig[1] - ocm[1], qz = pt->orig[2] - ocm[2];
            for (int a = 0; a < 3; a++) {
                const float *r = A9[a];
                float g = r[0] * qx + r[1] * qy + r[2] * qz + r[3] * qx * qx + r[4] * qy * qy + r[5] * qz * qz + r[6] * qx * qy +
                          r[7] * qy * qz + r[8] * qz * qx;
                pt->goal[a] = g + cm[a];
            }
        }
    }
}

static void corrected_velocity(ora_sim *s) { /* cpp:653-667 */
    apply_external_forces(s);
    project_positions(s);
    float t1 = 1.0f / s->dt;
    for (int i = 0; i < s->n; i++) {
        ora_particle *p = &s->p[i];
        for (int a = 0; a < 3; a++) p->corrected_vel[a] = p->predicted_vel[a] + (p->goal[a] - p->pos[a]) * t1 * s->alpha;
    }
}

/* ---------------------------------------------------------------- stage 3: cpp:669-701 */
static void intermediate_velocity(ora_sim *s) {
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        float pv[3] = {0.0f, 0.0f, 0.0f};
        FOR_CANDIDATES(s, p, j, {
            const ora_particle *np = &s->p[j];
            float d[3];
            float r2 = dist2(p, np, d);
            float w = poly6(s, r2);
            float vol = np->mass / np->dens;
            for (int a = 0; a < 3; a++) pv[a] += (np->corrected_vel[a] - p->corrected_vel[a]) * w * vol;
        });
        for (int a = 0; a < 3; a++) p->inter_vel[a] = p->corrected_vel[a] + pv[a] * s->velocity_mixing;
    }
}

/* ---------------------------------------------------------------- stage 4: cpp:448-513 */
static void density_pressure(ora_sim *s) {
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        p->dens = 0;
        p->pres = 0;
        FOR_CANDIDATES(s, p, j, {
            const ora_particle *np = &s->p[j];
            float d[3];
            float r2 = dist2(p, np, d);
            p->dens += np->mass * poly6(s, r2);
        });
        p->dens += p->mass * poly6(s, 0.0f); /* cpp:483: the extra self term (Q1) */
        p->pres = s->K * (p->dens - s->stand_density);
        p->pres -= (p->Vm * s->voltage_constant);
        if (p->stim > 0) {
            if (p->pres < -s->max_pressure) p->pres = -s->max_pressure;
            else if (p->pres > s->max_pressure) p->pres = s->max_pressure;
        } else {
            p->pres = -0.0f;
        }
    }
}

/* ---------------------------------------------------------------- stage 5: cpp:575-593 */
static void cell_model(ora_sim *s) {
    float denom = (s->FH_Vp - s->FH_Vr);
    float asd = (s->FH_Vt - s->FH_Vr) / denom;
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        float u = (p->Vm - s->FH_Vr) / denom;
        /* cpp:589: the `1.0` literal promotes the product and everything after it to double */
        p->Iion = (float)((double)p->Iion +
                          (double)s->dt * ((double)(s->C1 * u * (u - asd)) * ((double)u - 1.0) + (double)(s->C2 * p->w)) / (double)p->mass);
        p->w += s->dt * s->C3 * (u - s->C4 * p->w) / p->mass;
    }
}

/* ---------------------------------------------------------------- stage 6: cpp:515-573 */
static void compute_force(ora_sim *s) {
    for (int k = 0; k < s->n; k++) {
        ora_particle *p = &s->p[k];
        p->acc[0] = p->acc[1] = p->acc[2] = 0.0f;
        p->Inter_Vm = 0.0f;
        FOR_CANDIDATES(s, p, j, {
            const ora_particle *np = &s->p[j];
            float d[3];
            float r2 = dist2(p, np, d);
            if (r2 > ORA_INF) {
                float dis = sqrtf(r2);
                float vol = np->mass / np->dens;
                float fp = vol * (p->pres + np->pres) / 2 * spiky(s, dis);
                for (int a = 0; a < 3; a++) p->acc[a] -= d[a] * fp / dis;
                float fv = vol * s->mu * visco(s, dis);
                for (int a = 0; a < 3; a++) p->acc[a] += (np->inter_vel[a] - p->inter_vel[a]) * fv;
                p->Inter_Vm += (np->Vm - p->Vm) * vol * bspline2(s, dis);
            }
        });
        for (int a = 0; a < 3; a++) p->acc[a] = p->acc[a] / p->dens; /* cpp:568 */
        /* cpp:571 (the += form, Q9) */
        p->Inter_Vm += (s->sigma / (s->Beta * s->Cm)) * p->Inter_Vm - ((p->Iion - p->stim * s->dt / p->mass) / s->Cm);
    }
}

/* ---------------------------------------------------------------- stage 7: cpp:598-651 */
static void update_properties(ora_sim *s) {
    const float W[3] = {s->world.x, s->world.y, s->world.z};
    const float bmin[3] = {s->bmin.x, s->bmin.y, s->bmin.z}, bmax[3] = {s->bmax.x, s->bmax.y, s->bmax.z};
    for (int i = 0; i < s->n; i++) {
        ora_particle *p = &s->p[i];
        if (!p->fixed) {
            for (int a = 0; a < 3; a++) p->vel[a] = p->inter_vel[a] + (p->acc[a] * s->dt / p->mass);
            for (int a = 0; a < 3; a++) p->pos[a] = p->pos[a] + (p->vel[a] * s->dt);
        }
        p->Vm += p->Inter_Vm * s->dt / p->mass;
        if (p->Vm > s->max_voltage) p->Vm = s->max_voltage;
        else if (p->Vm < -s->max_voltage) p->Vm = -s->max_voltage;
        for (int a = 0; a < 3; a++) {
            if (p->pos[a] < 0.0f) { p->vel[a] = p->vel[a] * s->wall_hit; p->pos[a] = 0.0f; }
            if (p->pos[a] >= W[a]) { p->vel[a] = p->vel[a] * s->wall_hit; p->pos[a] = W[a] - 0.0001f; }
        }
        /* bounds.clamp, m3Bounds.h:84-88 */
        for (int a = 0; a < 3; a++) {
            if (bmin[a] > p->pos[a]) p->pos[a] = bmin[a];
            if (bmax[a] < p->pos[a]) p->pos[a] = bmax[a];
        }
    }
}

/* ---------------------------------------------------------------- step: cpp:794-829 */
void ora_stage(ora_sim *s, int stage) {
    switch (stage) {
        case 0:
            find_neighbors(s); corrected_velocity(s); intermediate_velocity(s); density_pressure(s);
            cell_model(s); compute_force(s); update_properties(s);
            break;
        case 1: find_neighbors(s); break;
        case 2: corrected_velocity(s); break;
        case 3: intermediate_velocity(s); break;
        case 4: density_pressure(s); break;
        case 5: cell_model(s); break;
        case 6: compute_force(s); break;
        case 7: update_properties(s); break;
        default: break;
    }
}
void ora_steps(ora_sim *s, int n) {
    for (int i = 0; i < n; i++) ora_stage(s, 0);
}

void ora_sm_debug(ora_sim *s, float *cm3, float *ocm3, float *xform27) {
    memcpy(cm3, s->dbg_cm, sizeof(s->dbg_cm));
    memcpy(ocm3, s->dbg_ocm, sizeof(s->dbg_ocm));
    memcpy(xform27, s->dbg_xform, sizeof(s->dbg_xform));
}
