// sphsm_pass5.cuh — neighbour passes with TWO target particles per thread (fifth generation).
//
// Generation 4 (sphsm_pass4.cuh) left both passes tied between the issue slots and the L1 data path (ncu r01_v7: pass A 87 %
// L1 / 61 % issue, pass B 73 % / 66 %): every lane requests 16-20 B per candidate, and a warp-wide gather costs 4.6-5.0 L1 tag
// lookups per LDG.128 whatever the layout.  The only way to ask L1 for less is to use each loaded record more than once.
// Here a thread owns two CONSECUTIVE slots.  They sit in the same or in adjacent cells, so for every stencil row their
// candidate windows [s0, e0) and [s1, e1) overlap almost entirely (cell_start is monotone in the key, hence s0 <= s1 and
// e0 <= e1); the thread sweeps the union once and evaluates each loaded pair of candidates against both targets:
//   * loads, address arithmetic and loop control are shared (per candidate pair: 34 instructions for two targets instead of
//     2 x 29 in pass A), the union is ~25 % longer than one window, so L1 requests fall by ~38 % and instructions by ~30 %
//     in phase 1;
//   * a candidate j belongs to target 0 iff j < e0 and to target 1 iff j >= s1 (one compare each); windows that do not touch
//     (the two slots straddle a row end) run as two segments with the same code;
//   * each target keeps its own in-range list and its phase 2 is unchanged, so every sum visits the same candidates in the
//     same order as generation 4: results are bit-identical to it.
#pragma once
#include <limits.h>

#include "sphsm_pass4.cuh"

namespace sphsm {

constexpr int PT5 = 128;                // threads per block = 256 target particles
constexpr unsigned LSTEP5 = 4u * PT5;   // bytes between consecutive entries of one list
constexpr int LIST_K5 = 16;             // entries per list between drains

// rows cb-1, cb, cb+1 of one plane for one target; `none`: what an absent window reads as (0 for target 0, INT_MAX for target 1,
// so that the monotone-window logic of sweep5 sees "before everything" / "after everything")
__device__ __forceinline__ void load_rows3v(const int *__restrict__ mid, int ga, bool ok, int none, Rows3 &r) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int *q = mid + (k - 1) * ga;
        r.s[k] = ok ? __ldg(q) : none;
        r.e[k] = ok ? __ldg(q + 3) : none;
    }
}

// pair(j, two, e0, s1, l0, l1): candidates j, j+1 against both targets (target 0 takes j < e0, target 1 takes j >= s1);
// one(j, e0, s1, l0, l1): single-candidate form; drain0 / drain1 consume the two lists.
template <class Pair, class One, class Drain0, class Drain1>
__device__ __forceinline__ void sweep5(const DevParams &p, const int *__restrict__ cell_start, int ga, int gagb, bool ok0, int key0, int cc0, bool ok1,
                                       int key1, int cc1, unsigned lbase0, unsigned &lofs0, unsigned lbase1, unsigned &lofs1, Pair &&pair, One &&one,
                                       Drain0 &&drain0, Drain1 &&drain1) {
    const int *c0 = cell_start + (key0 - 1), *c1 = cell_start + (key1 - 1);
    const unsigned lmax0 = lbase0 + LIST_K5 * LSTEP5, lmax1 = lbase1 + LIST_K5 * LSTEP5;
    const int c_lo = p.c_off, c_hi = p.c_off + p.gcl;
    Rows3 cur0, cur1, nxt0, nxt1;
    load_rows3v(c0 - gagb, ga, ok0 && cc0 - 1 >= c_lo, 0, cur0);
    load_rows3v(c1 - gagb, ga, ok1 && cc1 - 1 >= c_lo, INT_MAX, cur1);
    auto segment = [&](int j, const int end, const int e0, const int s1) {
        if (lofs0 + (unsigned)(end - j) * LSTEP5 <= lmax0 && lofs1 + (unsigned)(end - j) * LSTEP5 <= lmax1) {
#pragma unroll 1
            for (; j < end; j += 2) pair(j, j + 1 < end, e0, s1, lofs0, lofs1);
        } else {
            if (lofs0 == lmax0) drain0(lofs0);  // the unchecked path may have filled a list exactly
            if (lofs1 == lmax1) drain1(lofs1);
#pragma unroll 1
            for (; j < end; j++) {
                one(j, e0, s1, lofs0, lofs1);
                if (lofs0 == lmax0) drain0(lofs0);
                if (lofs1 == lmax1) drain1(lofs1);
            }
        }
    };
#pragma unroll 1
    for (int dc = -1; dc <= 1; dc++) {
        if (dc < 1) {
            load_rows3v(c0 + (dc + 1) * gagb, ga, ok0 && cc0 + dc + 1 < c_hi, 0, nxt0);
            load_rows3v(c1 + (dc + 1) * gagb, ga, ok1 && cc1 + dc + 1 < c_hi, INT_MAX, nxt1);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int s0 = cur0.s[k], e0 = cur0.e[k], s1 = cur1.s[k], e1 = cur1.e[k];
            if (s1 <= e0) {
                segment(s0, max(e0, e1), e0, s1);  // the windows touch: one sweep over their union
            } else {
                segment(s0, e0, e0, s1);           // target 0 alone (every j < s1)
                if (e1 != INT_MAX) segment(s1, e1, e0, s1);  // target 1 alone (every j >= e0)
            }
        }
        cur0 = nxt0;
        cur1 = nxt1;
    }
    drain0(lofs0);
    drain1(lofs1);
}

// slot of the k-th particle of the launch: own range minus the hole (see DevParams::hole_begin)
__device__ __forceinline__ int launch_slot(const DevParams &p, int k) {
    int i = p.own_begin + k;
    if (i >= p.hole_begin) i += p.hole_len;
    return i;
}

// ---------------------------------------------------------------------------------------------------
// pass A: density / pressure + XSPH intermediate velocity (reference cpp:448-513, 669-701)
__device__ __forceinline__ void pass_a_finish(const DevParams &p, const Arrays &a, int i, const float4 pi, const float4 ci, float dens, float pvx,
                                              float pvy, float pvz) {
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

__global__ void __launch_bounds__(PT5, 6) k_pass_a5(const __grid_constant__ DevParams p, const DevParams *__restrict__ g, Arrays a,
                                                    const int *__restrict__ cell_start, int count) {
    __shared__ int s_list[2 * LIST_K5 * PT5];
    const int k0 = 2 * (blockIdx.x * PT5 + threadIdx.x);
    if (k0 >= count) return;
    const bool has1 = k0 + 1 < count;
    const int i0 = launch_slot(p, k0), i1 = has1 ? launch_slot(p, k0 + 1) : i0;
    const float4 pi0 = a.P[i0], pi1 = a.P[i1];
    const int z0 = g->zero;  // == 0, loaded from global: what is derived from it stays in registers (see list_put)
    const float4 *__restrict__ P = pinned(a.P, z0);
    const float4 *__restrict__ C = a.C;
    const float h2 = g->h2, c6 = g->c_poly6;
    const int ga = g->ga, gagb = g->ga * g->gb;
    const unsigned lbase0 = (unsigned)__cvta_generic_to_shared(s_list) + 4u * (unsigned)(threadIdx.x + z0);
    const unsigned lbase1 = lbase0 + LIST_K5 * LSTEP5;
    const float2 nxy0 = make_float2(-pi0.x, -pi0.y), nxy1 = make_float2(-pi1.x, -pi1.y);
    const float nz0 = -pi0.z, nz1 = -pi1.z;
    float dens0 = 0.0f, ux0 = 0.0f, uy0 = 0.0f, uz0 = 0.0f, dens1 = 0.0f, ux1 = 0.0f, uy1 = 0.0f, uz1 = 0.0f;
    unsigned lofs0 = lbase0, lofs1 = lbase1;
    int ca, cb, cc0 = 0, cc1 = 0, key0 = 1, key1 = 1;
    const bool ok0 = cell_coords(p, pi0.x, pi0.y, pi0.z, ca, cb, cc0);
    if (ok0) key0 = cell_key(p, ca, cb, cc0);
    const bool ok1 = has1 && cell_coords(p, pi1.x, pi1.y, pi1.z, ca, cb, cc1);
    if (ok1) key1 = cell_key(p, ca, cb, cc1);
    auto r2_of = [](const float4 pj, const float2 nxy, const float nz) { return dist2_packed(__fadd2_rn(make_float2(pj.x, pj.y), nxy), pj.z + nz); };
    auto drain = [&](const int i, const float2 nxy, const float nz, const unsigned lbase, unsigned &lo, float &dens, float &ux, float &uy, float &uz) {
        if (lo == lbase) return;
        const float4 ci = a.C[i];
        for (unsigned q = lbase; q < lo; q += LSTEP5) {
            const int jj = list_get(q);
            const float4 pj = __ldg(P + jj);
            const float4 cj = __ldg(C + jj);
            const float x = h2 - r2_of(pj, nxy, nz);
            const float w = c6 * x * x * x;  // Poly6, cpp:151 (float on the fast path)
            dens = fmaf(pj.w, w, dens);
            const float t = w * cj.w;
            ux = fmaf(cj.x - ci.x, t, ux);
            uy = fmaf(cj.y - ci.y, t, uy);
            uz = fmaf(cj.z - ci.z, t, uz);
        }
        lo = lbase;
    };
    if (ok0 || ok1) {
        sweep5(
            p, cell_start, ga, gagb, ok0, key0, cc0, ok1, key1, cc1, lbase0, lofs0, lbase1, lofs1,
            [&](int j, bool two, int e0, int s1, unsigned &l0, unsigned &l1) {
                const float4 qa = __ldg(P + j), qb = __ldg(P + j + 1);
                // all four distances first, then the predicated appends (no short-circuit: the tests stay branch-free)
                const float ra0 = r2_of(qa, nxy0, nz0), rb0 = r2_of(qb, nxy0, nz0);
                const float ra1 = r2_of(qa, nxy1, nz1), rb1 = r2_of(qb, nxy1, nz1);
                const bool a0 = (j < e0) & (ra0 <= h2);  // Poly6 support, cpp:151
                const bool b0 = two & (j + 1 < e0) & (rb0 <= h2);
                const bool a1 = (j >= s1) & (ra1 <= h2);
                const bool b1 = two & (j + 1 >= s1) & (rb1 <= h2);
                if (a0) {
                    list_put(l0, j);
                    l0 += LSTEP5;
                }
                if (b0) {
                    list_put(l0, j + 1);
                    l0 += LSTEP5;
                }
                if (a1) {
                    list_put(l1, j);
                    l1 += LSTEP5;
                }
                if (b1) {
                    list_put(l1, j + 1);
                    l1 += LSTEP5;
                }
            },
            [&](int j, int e0, int s1, unsigned &l0, unsigned &l1) {
                const float4 qa = __ldg(P + j);
                if (j < e0 && r2_of(qa, nxy0, nz0) <= h2) {
                    list_put(l0, j);
                    l0 += LSTEP5;
                }
                if (j >= s1 && r2_of(qa, nxy1, nz1) <= h2) {
                    list_put(l1, j);
                    l1 += LSTEP5;
                }
            },
            [&](unsigned &lo) { drain(i0, nxy0, nz0, lbase0, lo, dens0, ux0, uy0, uz0); },
            [&](unsigned &lo) { drain(i1, nxy1, nz1, lbase1, lo, dens1, ux1, uy1, uz1); });
    }
    pass_a_finish(p, a, i0, pi0, a.C[i0], dens0, ux0, uy0, uz0);
    if (has1) pass_a_finish(p, a, i1, pi1, a.C[i1], dens1, ux1, uy1, uz1);
}

}  // namespace sphsm
