// sphsm_pass3.cuh — the neighbour passes with WARP-STAGED candidate tiles (production fast path, sm_100a).
//
// Why (ncu, profiles/r01_v3_*, 8M lattice): with one thread per particle gathering its ~37 candidates straight from
// global memory, pass A saturates the L1 pipe (l1tex throughput 82 %: a warp-wide 16-byte gather touches 4-5 lines per
// candidate, and the three-cell runs of neighbouring lanes overlap, so every record crosses L1 about four times per warp)
// and pass B is issue / latency bound behind two dependent loads per stencil row (row bounds -> candidates).  Here a warp
// cooperates instead:
//   1. every lane loads its nine row windows [s_k, e_k) (18 independent loads, one batch); the warp-wide union of each
//      row, [min s_k, max e_k), is one contiguous slot run because consecutive lanes sit in the same or adjacent cells;
//   2. the unions are copied ONCE, coalesced, into shared memory as structure-of-arrays (x[], y[], z[] and for pass B
//      Vm[], vol[]); unions that do not fit (dense meshes, warps that straddle a cell row) are staged in batches;
//   3. each lane walks its own windows inside the staged runs TWO candidates at a time: LDS.64 yields (x_j, x_j+1) as
//      an aligned register pair, and the distance / B-spline / Laplacian arithmetic runs on Blackwell's packed FP32
//      pipe (FADD2 / FMUL2 / FFMA2 with the particle's own coordinates as broadcast scalar operands); every element is
//      rounded on its own, so r^2 is bit-identical to the reference's mul/add sequence and neighbour-set membership
//      stays exact;
//   4. trip counts are made warp-uniform (warp max, lanes beyond their window are masked), so the candidate loop has no
//      divergent branches and no per-row remainder code;
//   5. in-range candidates (r <= h, ~19 %) are listed per lane and get their heavy terms in phase 2 as before.
#pragma once
#include "sphsm_pass.cuh"
#include "sphsm_pass2.cuh"
#include "sphsm_types.cuh"

namespace sphsm {

constexpr int W3_CAP = 384;   // staged candidates per warp and batch (lattice: 9 rows x ~38 = ~350)
constexpr int W3_LIST = 16;   // in-range list entries per lane between drains
constexpr int W3_BLK = 4;     // pair iterations between list-overflow checks (2 * W3_BLK <= W3_LIST)
constexpr unsigned FULL = 0xffffffffu;

template <int NF>
struct WarpSmem {
    float f[NF][W3_CAP];      // SoA fields of the staged candidates
    int ws[9][32], we[9][32]; // per-lane row windows (global slots)
    int us[9], ue[9];         // warp-wide union of each row
    int4 seg[9];              // staged segments of the current batch: (row, first global slot, length, staged offset)
    int list[W3_LIST][32];    // in-range candidates (global slots) of each lane
};

__device__ __forceinline__ float2 ldpair(const float *p) { return *reinterpret_cast<const float2 *>(p); }

// Row windows of this lane for the 27-cell stencil, reference order (dc outer, db inner; cpp:462-464): 18 loads in one
// batch.  Lanes without a particle (or outside the grid) get empty windows.
template <int NF>
__device__ __forceinline__ void load_windows(const DevParams &p, const int *__restrict__ cell_start, bool ok, int ca, int cb, int cc,
                                             WarpSmem<NF> &sm, int lane) {
    const int a_lo = max(ca - 1, 0), a_hi = min(ca + 1, p.ga - 1);
    int s[9], e[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const int c2 = cc + k / 3 - 1, b2 = cb + k % 3 - 1;
        const bool rok = ok && c2 >= p.c_off && c2 < p.c_off + p.gcl && b2 >= 0 && b2 < p.gb;
        const int row = p.ga * (b2 + p.gb * (c2 - p.c_off));
        s[k] = rok ? __ldg(cell_start + row + a_lo) : 0;
        e[k] = rok ? __ldg(cell_start + row + a_hi + 1) : 0;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
        sm.ws[k][lane] = s[k];
        sm.we[k][lane] = e[k];
        const int lo = __reduce_min_sync(FULL, e[k] > s[k] ? s[k] : 0x7fffffff);
        const int hi = __reduce_max_sync(FULL, e[k] > s[k] ? e[k] : 0);
        if (lane == 0) {
            sm.us[k] = lo;
            sm.ue[k] = hi;  // hi == 0: no lane has candidates in this row
        }
    }
    __syncwarp();
}

// The staged sweep.  `stage(g, t)` copies global slot g into staged position t; `pair(pp, m0, m1, j0)` evaluates the
// candidates at staged positions pp, pp + 1 (masks m0, m1; global slot of pp is j0) and returns the in-range bits
// (bit 0 / bit 1); `drain()` consumes the lane's in-range list.
template <int NF, class Stage, class Pair, class Drain>
__device__ __forceinline__ void sweep_staged(WarpSmem<NF> &sm, int lane, int &cnt, Stage &&stage, Pair &&pair, Drain &&drain) {
    int k = 0;
    int c0 = sm.us[0];
    while (k < 9) {
        // ---- stage as many whole rows (or one chunk of an oversized row) as fit; all of this is warp-uniform ----
        int nseg = 0, fill = 0;
#pragma unroll 1
        for (int q = 0; q < 9; q++) {
            while (k < 9 && c0 >= sm.ue[k]) {
                k++;
                if (k < 9) c0 = sm.us[k];
            }
            if (k >= 9) break;
            const int want = sm.ue[k] - c0;
            const int take = min(want, W3_CAP - fill);
            if (take <= 0 || (take < want && fill > 0)) break;  // rows are kept whole unless one alone overflows a batch
            if (lane == 0) sm.seg[q] = make_int4(k, c0, take, fill);
            for (int t = lane; t < take; t += 32) stage(c0 + t, fill + t);
            fill += (take + 1) & ~1;  // every segment starts on an even staged position (LDS.64 pairs)
            c0 += take;
            nseg = q + 1;
        }
        __syncwarp();
        // ---- every lane walks its own window inside each staged segment, two candidates per iteration ----
#pragma unroll 1
        for (int q = 0; q < nseg; q++) {
            const int4 sg = sm.seg[q];  // (row, first global slot, length, staged offset)
            const int ws = sm.ws[sg.x][lane], we = sm.we[sg.x][lane];
            const int delta = sg.y - sg.w;  // global slot = staged position + delta
            const int a = max(ws, sg.y) - delta, b = min(we, sg.y + sg.z) - delta;  // staged positions [a, b) of this lane
            int pp = a & ~1;
            const int np = b > a ? (b - pp + 1) >> 1 : 0;
            const int NP = __reduce_max_sync(FULL, np);  // warp-uniform trip count; lanes past their window are masked
#pragma unroll 1
            for (int it0 = 0; it0 < NP; it0 += W3_BLK) {
                const int nb = min(W3_BLK, NP - it0);
                if (__any_sync(FULL, cnt + 2 * nb > W3_LIST)) drain();
#pragma unroll 1
                for (int it = 0; it < nb; it++) {
                    const int pc = min(pp, W3_CAP - 2);
                    const bool m0 = pp >= a && pp < b, m1 = pp + 1 >= a && pp + 1 < b;
                    const int bits = pair(pc, m0, m1);
                    if (bits & 1) { sm.list[cnt][lane] = pp + delta; cnt++; }
                    if (bits & 2) { sm.list[cnt][lane] = pp + 1 + delta; cnt++; }
                    pp += 2;
                }
            }
        }
        __syncwarp();  // the staged arrays are rewritten by the next batch
    }
    drain();
}

// ---------------------------------------------------------------------------------------------------
// pass A: density / pressure + XSPH intermediate velocity (reference cpp:448-513, 669-701)
__global__ void __launch_bounds__(PT) k_pass_a3(const __grid_constant__ DevParams p, const DevParams *__restrict__ g, Arrays a,
                                                const int *__restrict__ cell_start) {
    __shared__ WarpSmem<3> s_w[PT / 32];
    const int lane = threadIdx.x & 31;
    WarpSmem<3> &sm = s_w[threadIdx.x >> 5];
    const int i = p.own_begin + blockIdx.x * PT + threadIdx.x;
    const bool act = i < p.own_end;
    if (p.own_begin + blockIdx.x * PT + (threadIdx.x & ~31) >= p.own_end) return;  // whole warp past the end
    const float4 pi = act ? a.P[i] : make_float4(0.f, 0.f, 0.f, 1.f);
    const float4 ci = act ? a.C[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *__restrict__ P = a.P;
    const float4 *__restrict__ C = a.C;
    const float h2 = g->h2, c6 = g->c_poly6;
    float dens = 0.0f, pvx = 0.0f, pvy = 0.0f, pvz = 0.0f;
    int cnt = 0;
    int ca = 0, cb = 0, cc = 0;
    const bool ok = act && cell_coords(p, pi.x, pi.y, pi.z, ca, cb, cc);
    load_windows<3>(p, cell_start, ok, ca, cb, cc, sm, lane);
    sweep_staged<3>(
        sm, lane, cnt,
        [&](int gslot, int t) {
            const float4 q = __ldg(P + gslot);
            sm.f[0][t] = q.x; sm.f[1][t] = q.y; sm.f[2][t] = q.z;
        },
        [&](int pc, bool m0, bool m1) -> int {
            const float2 x2 = ldpair(&sm.f[0][pc]), y2 = ldpair(&sm.f[1][pc]), z2 = ldpair(&sm.f[2][pc]);
            const float2 dx = __fadd2_rn(make_float2(pi.x, pi.x), make_float2(-x2.x, -x2.y));
            const float2 dy = __fadd2_rn(make_float2(pi.y, pi.y), make_float2(-y2.x, -y2.y));
            const float2 dz = __fadd2_rn(make_float2(pi.z, pi.z), make_float2(-z2.x, -z2.y));
            const float2 r2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
            return (m0 && r2.x <= h2 ? 1 : 0) | (m1 && r2.y <= h2 ? 2 : 0);  // Poly6 support, cpp:151
        },
        [&]() {
            for (int k = 0; k < cnt; k++) {
                const int jj = sm.list[k][lane];
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
    if (!act) return;
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
// (reference cpp:575-593, 515-573, 598-651)
template <bool DIAG>
__global__ void __launch_bounds__(PT) k_pass_b3(const __grid_constant__ DevParams p, const DevParams *__restrict__ g, Arrays a,
                                                float4 *__restrict__ Pout, const int *__restrict__ cell_start) {
    __shared__ WarpSmem<5> s_w[PT / 32];
    const int lane = threadIdx.x & 31;
    WarpSmem<5> &sm = s_w[threadIdx.x >> 5];
    const int i = p.own_begin + blockIdx.x * PT + threadIdx.x;
    const bool act = i < p.own_end;
    if (p.own_begin + blockIdx.x * PT + (threadIdx.x & ~31) >= p.own_end) return;
    const float4 pi = act ? a.P[i] : make_float4(0.f, 0.f, 0.f, 1.f);
    const float4 vi = act ? a.V[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 e4 = act ? a.E[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float pres_i = act ? a.S[i].x : 0.f;
    const float Vm_i = e4.x;
    cell_model<false>(p, e4.x, pi.w, e4.y, e4.z);

    const float4 *__restrict__ PB = a.PB;
    const float4 *__restrict__ V = a.V;
    const float2 *__restrict__ S = a.S;
    const float *__restrict__ VN = a.VN;
    const float sp2 = g->r2_spiky;
    const float a1 = g->bs_a1, b1 = g->bs_b1, a2 = g->bs_a2, b2 = g->bs_b2;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    float2 L2 = make_float2(0.f, 0.f);
    int cnt = 0;
    int ca = 0, cb = 0, cc = 0;
    const bool ok = act && cell_coords(p, pi.x, pi.y, pi.z, ca, cb, cc);
    load_windows<5>(p, cell_start, ok, ca, cb, cc, sm, lane);
    sweep_staged<5>(
        sm, lane, cnt,
        [&](int gslot, int t) {
            const float4 q = __ldg(PB + gslot);
            sm.f[0][t] = q.x; sm.f[1][t] = q.y; sm.f[2][t] = q.z; sm.f[3][t] = q.w;
            sm.f[4][t] = __ldg(VN + gslot);
        },
        [&](int pc, bool m0, bool m1) -> int {
            const float2 x2 = ldpair(&sm.f[0][pc]), y2 = ldpair(&sm.f[1][pc]), z2 = ldpair(&sm.f[2][pc]);
            const float2 vm2 = ldpair(&sm.f[3][pc]), vol2 = ldpair(&sm.f[4][pc]);
            const float2 dx = __fadd2_rn(make_float2(pi.x, pi.x), make_float2(-x2.x, -x2.y));
            const float2 dy = __fadd2_rn(make_float2(pi.y, pi.y), make_float2(-y2.x, -y2.y));
            const float2 dz = __fadd2_rn(make_float2(pi.z, pi.z), make_float2(-z2.x, -z2.y));
            const float2 r2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
            const bool v0 = m0 && r2.x > 1e-12f, v1 = m1 && r2.y > 1e-12f;  // INF, SPH_SM_monodomain.h:24, cpp:546
            // masked slots may hold stale shared memory: give them r^2 = 1 and select their product away (no NaN reaches L)
            const float2 rc = make_float2(v0 ? r2.x : 1.0f, v1 ? r2.y : 1.0f);
            const float2 r = __fmul2_rn(rc, make_float2(rsqrt_ftz(rc.x), rsqrt_ftz(rc.y)));
            // B_spline_2 (cpp:188-197) is min(inner line, outer line) inside 2h and 0 outside: min(t1, max(t2, 0))
            const float2 t1 = __ffma2_rn(r, make_float2(a1, a1), make_float2(b1, b1));
            const float2 t2 = __ffma2_rn(r, make_float2(a2, a2), make_float2(b2, b2));
            const float2 bs = make_float2(fminf(t1.x, fmaxf(t2.x, 0.0f)), fminf(t1.y, fmaxf(t2.y, 0.0f)));
            const float2 dv = __fadd2_rn(vm2, make_float2(-Vm_i, -Vm_i));
            const float2 dw = __fmul2_rn(dv, vol2);
            L2 = __ffma2_rn(make_float2(v0 ? dw.x : 0.0f, v1 ? dw.y : 0.0f), bs, L2);  // cpp:563
            return (v0 && r2.x <= sp2 ? 1 : 0) | (v1 && r2.y <= sp2 ? 2 : 0);  // Spiky / Visco support (r <= h), cpp:157,163
        },
        [&]() {
            const float hh = g->h, cs = g->c_spiky, mu = g->mu;
            for (int k = 0; k < cnt; k++) {
                const int jj = sm.list[k][lane];
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
    if (!act) return;
    const float L = L2.x + L2.y;
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
