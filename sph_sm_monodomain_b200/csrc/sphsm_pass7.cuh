// sphsm_pass7.cuh — neighbour passes over MULTI-ROW TILES with bulk-copy staging (seventh generation).
//
// Generation 6 (sphsm_pass6.cuh) showed what staging buys and what it costs (ncu profiles/r02_gen6_pass_b.json): phase 1 out of
// shared memory takes 18 % of the warp samples for 48 % of the instructions, but one ROW of T targets needs nine spans of ~T
// records each, 280 B of shared memory per target, i.e. 20 resident warps per SM — too few to hide the block start-up chain
// (key -> cell table -> bulk copies), phase 2's L2 gathers and the epilogue, and the kernel ends up slower than the gathered
// generation 4 (866 us against 688 at 8M).
//
// The fix is geometric.  A tile here is R = 4 adjacent cell rows of one plane times a run of cells along the fast axis, cut so
// that it holds at most T = 128 targets.  Its stencil is (R + 2) rows x 3 planes = 18 spans of (run + 2) cells: ~4.9 staged
// records per target instead of ~13, because the four rows share their neighbours.  ~100 B of staging per target puts 7-8
// blocks (28-32 warps) on an SM again — the occupancy of generation 4 with the phase 1 of generation 6 — and halves the bytes
// the bulk copies move through L2.
//
//   k_build_tiles   one warp per (plane, group of R rows): walks the cell table along the fast axis, cuts tiles greedily at
//                   <= T targets (counts are differences of cell_start entries: no scan), and writes for each tile its four
//                   target slot ranges and the slot ranges of its 18 spans.  Runs beside the rest of the sort.
//   k_pass_a7/b7    persistent blocks pull tiles from the list (atomic ticket), stage the spans with one bulk copy each
//                   (cp.async.bulk + mbarrier), map thread t to the t-th target of the tile and run the sweep of generation 6;
//                   a tile whose spans do not fit the stage (dense meshes) runs the gathered sweep instead.
// Candidates, order and arithmetic per target are those of generation 4: results are bit-identical to it.
#pragma once
#include "sphsm_pass6.cuh"

namespace sphsm {

constexpr int TILE_R = 4;     // cell rows per tile
constexpr int TILE_T = 128;   // targets per tile = threads per block
constexpr int TILE_SPANS = 3 * (TILE_R + 2);
constexpr int TILE_INTS = 64;  // ints per tile record
// tile record: [0] plane (relative to the rank's window) [1] first row (border-relative) [2] x0 [3] x1 (border-relative cells, inclusive)
//              [4..7] first target slot per row   [8..12] prefix of the per-row target counts (12 = total)
//              [16..33] first slot of each span   [34..51] length of each span (slots)
constexpr int TILE_TS = 4, TILE_TP = 8, TILE_SLO = 16, TILE_LEN = 34;

struct TileList {
    int *rec;      // TILE_INTS per tile
    int *count;    // tiles written
    int *tickets;  // work counters of the pass launches (zeroed by the builder)
    int capacity;
};
constexpr int TILE_TICKETS = 16;  // pass A: 0..4, pass B: 8..12 (launch kinds: all, interior, boundary, inner2, outer2)

// ---- tile builder ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_tile(const TileList &tl, const int *__restrict__ cs, int ga, int gb, int gcl, int cc, int cb0, int x0, int x1,
                                          const int (&ts)[TILE_R], const int (&tn)[TILE_R], int lane) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(tl.count, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= tl.capacity) return;  // (sized for the worst case of the greedy cut; never reached)
    int *r = tl.rec + (size_t)idx * TILE_INTS;
    if (lane == 0) {
        r[0] = cc; r[1] = cb0; r[2] = x0; r[3] = x1;
        int p = 0;
#pragma unroll
        for (int k = 0; k < TILE_R; k++) {
            r[TILE_TS + k] = ts[k];
            r[TILE_TP + k] = p;
            p += tn[k];
        }
        r[TILE_TP + TILE_R] = p;
    }
    if (lane < TILE_SPANS) {
        const int plane = cc + lane / (TILE_R + 2) - 1, row = cb0 - 1 + lane % (TILE_R + 2);
        int slo = 0, len = 0;
        if (plane >= 0 && plane < gcl && row >= 0 && row < gb) {
            const int klo = (x0 - 1) + ga * (row + gb * plane);
            slo = __ldg(cs + klo);
            len = __ldg(cs + klo + (x1 - x0 + 3)) - slo;
        }
        r[TILE_SLO + lane] = slo;
        r[TILE_LEN + lane] = len;
    }
}

// limbo_end > 0 (one GPU): the particles outside the grid (limbo bucket, slots [cell_start[num_cells], limbo_end)) have no
// neighbours but are still integrated: they become target-only tiles of plane -1.  A slab rank's limbo bucket holds dead entries.
__global__ void __launch_bounds__(128) k_build_tiles(const int *__restrict__ cs, int ga, int gb, int gcl, int limbo_end, TileList tl) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w == 0 && lane < TILE_TICKETS) tl.tickets[lane] = 0;
    const int groups = (gb - 2 + TILE_R - 1) / TILE_R;
    if (w == groups * gcl) {
        const int s = __ldg(cs + ga * gb * gcl);
        for (int c = s; c < limbo_end; c += TILE_T) {
            int idx = 0;
            if (lane == 0) idx = atomicAdd(tl.count, 1);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx >= tl.capacity) break;
            int *r = tl.rec + (size_t)idx * TILE_INTS;
            for (int k = lane; k < TILE_INTS; k += 32) r[k] = 0;
            __syncwarp();
            if (lane == 0) {
                const int cnt = min(TILE_T, limbo_end - c);
                r[0] = -1;
                r[TILE_TS] = c;
                r[TILE_TP + 1] = r[TILE_TP + 2] = r[TILE_TP + 3] = r[TILE_TP + 4] = cnt;
            }
        }
        return;
    }
    if (w >= groups * gcl) return;
    const int cc = w / groups, cb0 = 1 + (w % groups) * TILE_R;
    int rowkey[TILE_R];
    bool rowok[TILE_R];
#pragma unroll
    for (int k = 0; k < TILE_R; k++) {
        rowok[k] = cb0 + k <= gb - 2;
        rowkey[k] = ga * (cb0 + k + gb * cc);
    }
    // F(x) = sum over the rows of cell_start[row, x]: the targets in columns [a, b] are F(b + 1) - F(a)
    auto F = [&](int x) {
        int s = 0;
#pragma unroll
        for (int k = 0; k < TILE_R; k++)
            if (rowok[k]) s += __ldg(cs + rowkey[k] + x);
        return s;
    };
    const int xend = ga - 2;  // last real column
    if (F(xend + 1) - F(1) == 0) return;  // nothing in this group of rows
    int x0 = 1;
    while (x0 <= xend) {
        const int f0 = F(x0);
        // largest x1 >= x0 with at most TILE_T targets in [x0, x1]: the lanes probe 32 columns at a time (counts are monotone)
        int x1 = x0 - 1;
        for (int base = x0; base <= xend; base += 32) {
            const int x = base + lane;
            const bool fits = x <= xend && F(x + 1) - f0 <= TILE_T;
            const unsigned m = __ballot_sync(0xffffffffu, fits);
            const int lead = m == 0xffffffffu ? 32 : __ffs(~m) - 1;  // leading run of fitting columns
            x1 = base + lead - 1;
            if (lead < 32) break;
        }
        if (x1 < x0) {
            // one column of the four rows holds more than TILE_T targets (dense meshes): a tile per row, cells cut into
            // chunks of TILE_T targets; their spans cover the same stencil (and will not fit the stage: gathered sweep)
#pragma unroll
            for (int k = 0; k < TILE_R; k++) {
                if (!rowok[k]) continue;
                const int s = __ldg(cs + rowkey[k] + x0), e = __ldg(cs + rowkey[k] + x0 + 1);
                for (int c = s; c < e; c += TILE_T) {
                    int ts[TILE_R] = {0, 0, 0, 0}, tn[TILE_R] = {0, 0, 0, 0};
                    ts[k] = c;
                    tn[k] = min(TILE_T, e - c);
                    emit_tile(tl, cs, ga, gb, gcl, cc, cb0, x0, x0, ts, tn, lane);
                }
            }
            x0 += 1;
            continue;
        }
        int ts[TILE_R], tn[TILE_R], total = 0;
#pragma unroll
        for (int k = 0; k < TILE_R; k++) {
            ts[k] = rowok[k] ? __ldg(cs + rowkey[k] + x0) : 0;
            tn[k] = rowok[k] ? __ldg(cs + rowkey[k] + x1 + 1) - ts[k] : 0;
            total += tn[k];
        }
        if (total > 0) emit_tile(tl, cs, ga, gb, gcl, cc, cb0, x0, x1, ts, tn, lane);
        x0 = x1 + 1;
    }
}

// ---- shared-memory layout of a tile block ------------------------------------------------------------------------------------
template <bool WITH4>
struct Lay7 {
    static constexpr int SLOTS = 1024;  // 16-byte records staged per tile: (R + 2) x 3 spans; a lattice tile needs 500-960
    static constexpr int SLOTS4 = SLOTS + TILE_SPANS * 8;
    static constexpr int LISTK = 12;  // in-range list entries per lane between drains (8 blocks of 128 threads per SM)
    static constexpr unsigned OFF_MBAR = 0, OFF_TICKET = 8, OFF_FLAG = 12, OFF_C16 = 16, OFF_C4 = OFF_C16 + 4u * TILE_SPANS, OFF_LIST = 256;
    static constexpr unsigned OFF_ST16 = OFF_LIST + 4u * LISTK * TILE_T;
    static constexpr unsigned OFF_ST4 = OFF_ST16 + 16u * SLOTS;
    static constexpr unsigned BYTES = WITH4 ? OFF_ST4 + 4u * SLOTS4 : OFF_ST4;
};
static_assert(Lay7<true>::OFF_C4 + 4u * TILE_SPANS <= Lay7<true>::OFF_LIST, "tables overlap the lists");

// warp 0: read the tile's spans, decide, issue the bulk copies, publish the address tables (as stage_spans6, 18 spans)
template <bool WITH4>
__device__ __forceinline__ void stage_tile7(const int *__restrict__ rec, const float4 *__restrict__ A16, const float *__restrict__ A4, const unsigned smem0,
                                            const bool enabled) {
    using L = Lay7<WITH4>;
    const int lane = threadIdx.x & 31;
    int slo = 0, len = 0;
    if (lane < TILE_SPANS) {
        slo = __ldg(rec + TILE_SLO + lane);
        len = __ldg(rec + TILE_LEN + lane);
    }
    const int n16 = len > 0 ? len + SPAN_TAIL : 0;
    const int lo4 = slo & ~3;
    const int n4 = (WITH4 && len > 0) ? ((slo + len + SPAN_TAIL + 3) & ~3) - lo4 : 0;
    int o16 = n16, o4 = n4;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t16 = __shfl_up_sync(0xffffffffu, o16, d), t4 = __shfl_up_sync(0xffffffffu, o4, d);
        if (lane >= d) { o16 += t16; o4 += t4; }
    }
    const int tot16 = __shfl_sync(0xffffffffu, o16, 31), tot4 = __shfl_sync(0xffffffffu, o4, 31);
    o16 -= n16;
    o4 -= n4;
    const bool staged = enabled && tot16 <= L::SLOTS && tot4 <= L::SLOTS4;
    const unsigned mbar = smem0 + L::OFF_MBAR;
    const unsigned st16 = smem0 + L::OFF_ST16 + 16u * (unsigned)o16, st4 = smem0 + L::OFF_ST4 + 4u * (unsigned)o4;
    if (lane < TILE_SPANS) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem0 + L::OFF_C16 + 4u * lane), "r"(st16 - 16u * (unsigned)slo) : "memory");
        if (WITH4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem0 + L::OFF_C4 + 4u * lane), "r"(st4 - 4u * (unsigned)lo4) : "memory");
    }
    if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem0 + L::OFF_FLAG), "r"(staged ? 1u : 0u) : "memory");
    if (staged) {
        if (lane == 0) mbar_arrive_expect_tx(mbar, 16u * (unsigned)tot16 + 4u * (unsigned)tot4);
        __syncwarp();
        if (n16 > 0) {
            bulk_g2s(st16, A16 + slo, 16u * (unsigned)n16, mbar);
            if (WITH4) bulk_g2s(st4, A4 + lo4, 4u * (unsigned)n4, mbar);
        }
    }
}

// thread t of the block -> the t-th target of the tile (row index in *row), or -1
__device__ __forceinline__ int tile_target(const int *__restrict__ rec, int t, int &row) {
    const int p1 = __ldg(rec + TILE_TP + 1), p2 = __ldg(rec + TILE_TP + 2), p3 = __ldg(rec + TILE_TP + 3), p4 = __ldg(rec + TILE_TP + 4);
    if (t >= p4) return -1;
    row = (t >= p1) + (t >= p2) + (t >= p3);
    const int base = row == 0 ? 0 : (row == 1 ? p1 : (row == 2 ? p2 : p3));
    return __ldg(rec + TILE_TS + row) + (t - base);
}

// planes a launch covers: [lo, hi) minus [hole_lo, hole_hi), relative to the rank's window (one GPU: everything)
struct PlaneSet {
    int lo, hi, hole_lo, hole_hi;
};
__device__ __forceinline__ bool plane_in(const PlaneSet &ps, int c) { return c >= ps.lo && c < ps.hi && !(c >= ps.hole_lo && c < ps.hole_hi); }

// ---------------------------------------------------------------------------------------------------
// pass A over tiles
__global__ void __launch_bounds__(TILE_T, 8) k_pass_a7(const __grid_constant__ DevParams p, const DevParams *__restrict__ g, Arrays a,
                                                        const int *__restrict__ cell_start, const uint32_t *__restrict__ skey, TileList tl, int ticket,
                                                        PlaneSet ps, int stage_on) {
    using L = Lay7<false>;
    extern __shared__ __align__(128) uint8_t smem7[];
    const unsigned smem0 = smem_u32(smem7);
    if (threadIdx.x == 0) {
        mbar_init(smem0 + L::OFF_MBAR, 1);
        mbar_fence_init();
    }
    const int ntiles = min(*tl.count, tl.capacity);
    unsigned parity = 0;
    while (true) {
        __syncthreads();  // every thread is done with the previous tile's shared memory (and sees the barrier initialised)
        if (threadIdx.x == 0) {
            const int t = atomicAdd(tl.tickets + ticket, 1);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem0 + L::OFF_TICKET), "r"(t) : "memory");
        }
        __syncthreads();
        const int t = (int)lds_u32(smem0 + L::OFF_TICKET);
        if (t >= ntiles) break;
        const int *rec = tl.rec + (size_t)t * TILE_INTS;
        if (!plane_in(ps, __ldg(rec))) continue;
        if (threadIdx.x < 32) stage_tile7<false>(rec, a.P, nullptr, smem0, stage_on != 0);
        int row = 0;
        const int i = tile_target(rec, threadIdx.x, row);
        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), ci = pi;
        int key = p.num_cells;
        if (i >= 0) {
            pi = a.P[i];
            ci = a.C[i];
            key = (int)skey[i];
        }
        __syncthreads();  // the tables of this tile are published
        const bool staged = lds_u32(smem0 + L::OFF_FLAG) != 0;
        float dens = 0.0f, pvx = 0.0f, pvy = 0.0f, pvz = 0.0f;
        if (staged) {
            mbar_wait(smem0 + L::OFF_MBAR, parity);
            parity ^= 1u;
            if (key < p.num_cells) pass_a6_neighbours<L, TILE_T, true>(g, a, cell_start, key, smem0, TILE_R + 2, row, pi, ci, dens, pvx, pvy, pvz);
        } else if (key < p.num_cells) {
            pass_a6_neighbours<L, TILE_T, false>(g, a, cell_start, key, smem0, TILE_R + 2, row, pi, ci, dens, pvx, pvy, pvz);
        }
        if (i >= 0) pass_a_finish(p, a, i, pi, ci, dens, pvx, pvy, pvz);
    }
}

// pass B over tiles (cell_count != nullptr: the next step's counting-sort input is filed as in generation 4)
template <int STEP, bool DIAG, int MINB>
__global__ void __launch_bounds__(TILE_T, MINB) k_pass_b7(const __grid_constant__ DevParams p, const DevParams *__restrict__ g, Arrays a,
                                                        float4 *__restrict__ Pout, const int *__restrict__ cell_start, const uint32_t *__restrict__ skey,
                                                        uint32_t *__restrict__ next_keys, uint32_t *__restrict__ next_rank,
                                                        uint32_t *__restrict__ cell_count, TileList tl, int ticket, PlaneSet ps, int stage_on) {
    using L = Lay7<true>;
    extern __shared__ __align__(128) uint8_t smem7[];
    const unsigned smem0 = smem_u32(smem7);
    if (threadIdx.x == 0) {
        mbar_init(smem0 + L::OFF_MBAR, 1);
        mbar_fence_init();
    }
    const int ntiles = min(*tl.count, tl.capacity);
    unsigned parity = 0;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const int t = atomicAdd(tl.tickets + ticket, 1);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem0 + L::OFF_TICKET), "r"(t) : "memory");
        }
        __syncthreads();
        const int t = (int)lds_u32(smem0 + L::OFF_TICKET);
        if (t >= ntiles) break;
        const int *rec = tl.rec + (size_t)t * TILE_INTS;
        if (!plane_in(ps, __ldg(rec))) continue;
        if (threadIdx.x < 32) stage_tile7<true>(rec, a.PB, a.VN, smem0, stage_on != 0);
        int row = 0;
        const int i = tile_target(rec, threadIdx.x, row);
        // the tile's own records and the ionic model run while the spans are in flight
        float4 pi = make_float4(0.f, 0.f, 0.f, 1.f), vi = make_float4(0.f, 0.f, 0.f, 0.f), e4 = vi;
        float2 si = make_float2(0.f, 1.f);
        bool fixed = false;
        int key = p.num_cells;
        if (i >= 0) {
            pi = a.P[i];
            vi = a.V[i];
            e4 = a.E[i];
            si = a.S[i];  // (pres, dens)
            fixed = __float_as_int(a.O[i].w) != 0;
            key = (int)skey[i];
        }
        const float Vm_i = e4.x;
        const float inv_mass = rcp_ftz(pi.w);
        if (i >= 0) {
            if (SPHSM_FAST_ODE) cell_model_fast(p, e4.x, inv_mass, e4.y, e4.z);
            else cell_model<false>(p, e4.x, pi.w, e4.y, e4.z);
        }
        __syncthreads();
        const bool staged = lds_u32(smem0 + L::OFF_FLAG) != 0;
        float ax = 0.0f, ay = 0.0f, az = 0.0f, Lsum = 0.0f;
        if (staged) {
            mbar_wait(smem0 + L::OFF_MBAR, parity);
            parity ^= 1u;
            if (key < p.num_cells) pass_b6_neighbours<L, TILE_T, STEP, true>(g, a, cell_start, key, smem0, TILE_R + 2, row, pi, vi, Vm_i, si.x, ax, ay, az, Lsum);
        } else if (key < p.num_cells) {
            pass_b6_neighbours<L, TILE_T, STEP, false>(g, a, cell_start, key, smem0, TILE_R + 2, row, pi, vi, Vm_i, si.x, ax, ay, az, Lsum);
        }
        if (i >= 0) pass_b_finish<DIAG>(p, a, Pout, i, pi, vi, e4, si.y, fixed, ax, ay, az, Lsum, inv_mass, next_keys, next_rank, cell_count);
    }
}

}  // namespace sphsm
