// sphsm_pass2.cuh — the production (fast-path) neighbour passes.  Same arithmetic as sphsm_pass.cuh's FAST branches,
// restructured after the first ncu captures (profiles/r01_v1_*.json: both passes instruction-issue bound; the in-range
// work was if-converted and issued for every candidate at ~19 % lane activity; 9 LDC/LDCU parameter re-loads and a
// denormal-safe rsqrt sequence sat in the inner loop):
//   * two phases per lane: phase 1 sweeps the 27-cell stencil, tests every candidate (in pass B it also accumulates
//     the 2h-support Laplacian that ~94 % of the candidates contribute to) and appends the slots of the in-range ones
//     (r <= h, ~19 %) to a small per-lane list in shared memory; phase 2 walks that list and does the heavy in-range
//     work with every lane busy;
//   * loop-invariant parameters pinned in registers, rsqrt.approx.ftz (r2 > 1e-12 there).
// Shared memory is kept to 8 KB per block on purpose: the gathers live on L1 hits (a first version with 25 KB per
// block left 23 KB of L1 per SM and the hit rate fell from 64 % to 9-26 %).
// Neighbour-set membership stays bit-exact: r^2 is evaluated without FMA and compared with the same exact thresholds.
#pragma once
#include "sphsm_pass.cuh"
#include "sphsm_types.cuh"

namespace sphsm {

#ifndef SPHSM_P1_UNROLL
#define SPHSM_P1_UNROLL 2
#endif
constexpr int PT = 128;     // threads per block in the fast passes
constexpr int LIST_K = 16;  // in-range list entries per lane between drains

// Loop invariants are read from the GLOBAL-memory copy of the parameter block (`g`): ptxas re-materialises anything it
// can recompute from kernel parameters or special registers inside the candidate loop (LDC/LDCU, S2R+ULEA for the
// shared-memory base: 8 issue slots per candidate in the first builds), but it never re-issues a global load.
// Base pointers stay kernel parameters: LDC.64 + IMAD.WIDE per gather is what ptxas emits at best anyway.
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Sweep the stencil rows in the reference's order.  `body(j)` runs for every candidate slot and returns true when the
// slot must be appended to the in-range list; `drain()` consumes the list.  A row that could overflow the list takes
// the checked path (dense meshes); lattice-like inputs never do.
// The bounds of the three rows of a stencil plane are loaded together, one plane AHEAD of the sweep (r01_v3 profile:
// a third of the stall samples sat on the row-bounds load -> candidate load chain, nine times per particle).
struct PlaneRows {
    int s[3], e[3];
};
__device__ __forceinline__ void load_plane_rows(const DevParams &p, const int *__restrict__ cell_start, int a_lo, int a_hi, int cb, int c2, PlaneRows &r) {
    const bool plane_ok = c2 >= p.c_off && c2 < p.c_off + p.gcl;
    const int base = p.ga * (p.gb * (c2 - p.c_off));
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int b2 = cb + k - 1;
        const bool ok = plane_ok && b2 >= 0 && b2 < p.gb;
        const int row = base + p.ga * b2;
        r.s[k] = ok ? __ldg(cell_start + row + a_lo) : 0;
        r.e[k] = ok ? __ldg(cell_start + row + a_hi + 1) : 0;
    }
}

template <class Body, class Drain>
__device__ __forceinline__ void sweep_two_phase(const DevParams &p, const int *__restrict__ cell_start, int ca, int cb, int cc,
                                                int *my_list /* s_list + tid, stride PT; also read by drain(): no restrict */, int &cnt, Body &&body, Drain &&drain) {
    const int a_lo = max(ca - 1, 0), a_hi = min(ca + 1, p.ga - 1);
    PlaneRows cur, nxt;
    load_plane_rows(p, cell_start, a_lo, a_hi, cb, cc - 1, cur);
#pragma unroll 1
    for (int dc = -1; dc <= 1; dc++) {
        if (dc < 1) load_plane_rows(p, cell_start, a_lo, a_hi, cb, cc + dc + 1, nxt);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int j = cur.s[k];
            const int e = cur.e[k];
            if (cnt + (e - j) <= LIST_K) {
                // candidates two at a time; a row of odd length repeats its last slot with `live` false, so lanes whose
                // rows differ in parity stay in the same code (the per-row remainder loop was 8 % of the issued
                // instructions at 19 of 32 lanes, and rows are only 3-6 candidates long on lattice-like inputs)
#pragma unroll 1
                for (; j < e; j += 2) {
                    const bool two = j + 1 < e;
                    const int j1 = two ? j + 1 : j;
                    const bool in0 = body(j, true);
                    const bool in1 = body(j1, two);
                    if (in0) {
                        my_list[cnt * PT] = j;
                        cnt++;
                    }
                    if (in1) {
                        my_list[cnt * PT] = j1;
                        cnt++;
                    }
                }
            } else {
                if (cnt == LIST_K) drain();  // the unchecked path may have filled the list exactly
#pragma unroll 1
                for (; j < e; j++) {
                    if (body(j, true)) {
                        my_list[cnt * PT] = j;
                        if (++cnt == LIST_K) drain();
                    }
                }
            }
        }
        cur = nxt;
    }
    drain();
}

// ---------------------------------------------------------------------------------------------------
// pass A: density / pressure + XSPH intermediate velocity (reference cpp:448-513, 669-701)
__global__ void __launch_bounds__(PT) k_pass_a2(const __grid_constant__ DevParams p, const DevParams *__restrict__ g, Arrays a,
                                                const int *__restrict__ cell_start) {
    __shared__ int s_list[LIST_K * PT];
    const int i = p.own_begin + blockIdx.x * PT + threadIdx.x;
    if (i >= p.own_end) return;
    const float4 pi = a.P[i];
    const float4 ci = a.C[i];
    const float4 *__restrict__ P = a.P;
    const float4 *__restrict__ C = a.C;
    const float h2 = g->h2, c6 = g->c_poly6;
    const int z0 = g->zero;
    float dens = 0.0f, pvx = 0.0f, pvy = 0.0f, pvz = 0.0f;
    int cnt = 0;
    int *my_list = s_list + (threadIdx.x + z0);  // z0 == 0, loaded from global: keeps the base in a register
    int ca, cb, cc;
    if (cell_coords(p, pi.x, pi.y, pi.z, ca, cb, cc)) {
        sweep_two_phase(
            p, cell_start, ca, cb, cc, my_list, cnt,
            [&](int j, bool live) -> bool {
                const float4 pj = __ldg(P + j);
                return live && dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z) <= h2;  // Poly6 support, cpp:151
            },
            [&]() {
                for (int k = 0; k < cnt; k++) {
                    const int jj = my_list[k * PT];
                    const float4 pj = __ldg(P + jj);
                    const float4 cj = __ldg(C + jj);
                    const float x = h2 - dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                    const float w = c6 * x * x * x;  // Poly6, cpp:151 (float on the fast path)
                    dens = fmaf(pj.w, w, dens);
                    const float t = w * cj.w;
                    pvx = fmaf(cj.x - ci.x, t, pvx);
                    pvy = fmaf(cj.y - ci.y, t, pvy);
                    pvz = fmaf(cj.z - ci.z, t, pvz);
                }
                cnt = 0;
            });
    }
    const float4 e4 = a.E[i];
    dens = fmaf(pi.w, p.poly6_self, dens);                           // the extra self term, cpp:483 (Q1)
    float pres = p.K * (dens - p.rho0) - e4.x * p.voltage_constant;  // cpp:486-491
    if (e4.w > 0.0f) pres = fminf(fmaxf(pres, -p.max_pressure), p.max_pressure);
    else pres = -0.0f;  // cpp:493-503 (Q2)
    a.VEL[i].w = dens;
    a.S[i] = make_float2(pres, e4.x);
    const float vol = __fdiv_rn(pi.w, dens);  // np->mass / np->dens as pass B reads it, cpp:551
    a.V[i] = make_float4(fmaf(pvx, p.mix, ci.x), fmaf(pvy, p.mix, ci.y), fmaf(pvz, p.mix, ci.z), vol);
    a.VN[i] = vol;
}

// ---------------------------------------------------------------------------------------------------
// pass B: ionic cell model + pressure / viscosity force + SPH Laplacian of Vm + integration and walls
// (reference cpp:575-593, 515-573, 598-651).  PB = (pos.xyz, Vm) is the neighbour record of this pass.
template <bool DIAG>
__global__ void __launch_bounds__(PT) k_pass_b2(const __grid_constant__ DevParams p, const DevParams *__restrict__ g, Arrays a,
                                                float4 *__restrict__ Pout, const int *__restrict__ cell_start) {
    __shared__ int s_list[LIST_K * PT];
    const int i = p.own_begin + blockIdx.x * PT + threadIdx.x;
    if (i >= p.own_end) return;
    const float4 pi = a.P[i];
    const float4 vi = a.V[i];
    float4 e4 = a.E[i];
    const float pres_i = a.S[i].x;
    const float Vm_i = e4.x;
    cell_model<false>(p, e4.x, pi.w, e4.y, e4.z);

    const float4 *__restrict__ PB = a.PB;
    const float4 *__restrict__ V = a.V;
    const float2 *__restrict__ S = a.S;
    const float *__restrict__ VN = a.VN;
    const float q2 = g->r2_q2, q1 = g->r2_q1, sp2 = g->r2_spiky;
    const float a1 = g->bs_a1, b1 = g->bs_b1, a2 = g->bs_a2, b2 = g->bs_b2;
    const int z0 = g->zero;
    float ax = 0.0f, ay = 0.0f, az = 0.0f, L = 0.0f;
    int cnt = 0;
    int *my_list = s_list + (threadIdx.x + z0);  // z0 == 0, loaded from global: keeps the base in a register
    int ca, cb, cc;
    if (cell_coords(p, pi.x, pi.y, pi.z, ca, cb, cc)) {
        sweep_two_phase(
            p, cell_start, ca, cb, cc, my_list, cnt,
            [&](int j, bool live) -> bool {
                const float4 pj = __ldg(PB + j);
                const float vol = __ldg(VN + j);
                const float r2 = dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                const bool near0 = !(live && r2 > 1e-12f);  // INF, SPH_SM_monodomain.h:24, cpp:546 (also catches NaN)
                const bool valid = !near0 && r2 <= q2;    // B_spline_2 support (q < 2), cpp:193
                const float r2s = valid ? r2 : 1.0f;
                const float r = r2s * rsqrt_ftz(r2s);
                const bool inner = r2 <= q1;              // q < 1, cpp:191
                const float bs = fmaf(inner ? a1 : a2, r, inner ? b1 : b2);
                L = fmaf((pj.w - Vm_i) * (valid ? vol : 0.0f), bs, L);  // cpp:563
                return !near0 && r2 <= sp2;               // Spiky / Visco support (r <= h), cpp:157,163
            },
            [&]() {
                const float hh = g->h, cs = g->c_spiky, mu = g->mu;
                for (int k = 0; k < cnt; k++) {
                    const int jj = my_list[k * PT];
                    const float4 pj = __ldg(PB + jj);
                    const float4 vj = __ldg(V + jj);
                    const float pres_j = __ldg(&S[jj].x);
                    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    const float r2 = dist2_exact(dx, dy, dz);
                    const float inv_r = rsqrt_ftz(r2);
                    const float r = r2 * inv_r;
                    const float hr = hh - r;
                    const float t = vj.w * hr * cs;
                    const float fpr = t * (pres_i + pres_j) * (0.5f * hr) * inv_r;  // = -(Force_pressure / dis), cpp:553-554
                    const float fv = t * mu;                                        // Force_viscosity, cpp:559
                    ax = fmaf(dx, fpr, ax);
                    ay = fmaf(dy, fpr, ay);
                    az = fmaf(dz, fpr, az);
                    ax = fmaf(vj.x - vi.x, fv, ax);
                    ay = fmaf(vj.y - vi.y, fv, ay);
                    az = fmaf(vj.z - vi.z, fv, az);
                }
                cnt = 0;
            });
    }
    float4 v4 = a.VEL[i];
    const float dens = v4.w;
    ax = ax / dens;  // cpp:568
    ay = ay / dens;
    az = az / dens;
    // cpp:571: Inter_Vm += (sigma/(Beta*Cm))*Inter_Vm - ((Iion - stim*dt/mass)/Cm)   (the += form, Q9)
    const float ivm = L + (p.diff_coef * L - (e4.y - (e4.w * p.dt) / pi.w) / p.Cm);
    if (DIAG) a.ACC[i] = make_float4(ax, ay, az, ivm);
    const bool fixed = __float_as_int(a.O[i].w) != 0;
    float x = pi.x, y = pi.y, z = pi.z;
    integrate<false>(p, fixed, pi.w, vi.x, vi.y, vi.z, ax, ay, az, ivm, x, y, z, v4.x, v4.y, v4.z, e4.x);
    Pout[i] = make_float4(x, y, z, pi.w);
    a.VEL[i] = v4;
    a.E[i] = e4;
}

}  // namespace sphsm
