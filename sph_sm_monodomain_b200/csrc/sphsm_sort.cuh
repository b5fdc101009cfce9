// sphsm_sort.cuh — subsystem (1): uniform-grid neighbour search.
//   k_hash            cell key per particle (+ all radix-digit histograms in the same read)
//   k_radix_pass      CUB-free LSD radix sort pass, 8-bit digits: one kernel per pass, per-tile stable ranking
//                     with warp match/ballot, inter-tile digit prefixes by decoupled look-back, tile staged in
//                     shared memory so the scatter leaves the SM as coalesced runs
//   k_cell_order_fix  (strict mode) in-cell order = ascending ORIGINAL index, the reference's bucket order
//   k_cell_bounds     cell_start table (num_cells + 2 entries, prefix form: cell c = [start[c], start[c+1]))
//   k_cell_count / k_scan_onepass / k_cell_scatter / k_cell_sort_big
//                     the counting sort of the fast path: per-cell counts + provisional ranks (filed by pass B on one GPU), cell table
//                     by a single-pass scan, slots; in-cell order (ascending original index) is applied by the gather itself
//                     (ordered_source), cells of more than 8 entries by a warp each
//   k_reorder         gather the SoA state into the new slot order
// Replaces Find_neighbors / Calculate_Cell_Position / Calculate_Cell_Hash, reference cpp:127-146, 199-213.
#pragma once
#include <limits.h>

#include "sphsm_types.cuh"

namespace sphsm {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_IPT = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;  // 4096 keys per tile
constexpr int MAX_SORT_PASSES = 4;

constexpr uint32_t TS_FLAG_AGG = 1u << 30;   // tile aggregate published
constexpr uint32_t TS_FLAG_INCL = 2u << 30;  // inclusive prefix published
constexpr uint32_t TS_MASK = (1u << 30) - 1;

// ---- cell key + digit histograms ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hash(const __grid_constant__ DevParams p, const float4 *__restrict__ P,
                                              uint32_t *__restrict__ keys, uint32_t *__restrict__ ghist /*passes*256*/,
                                              int passes) {
    __shared__ uint32_t s_hist[MAX_SORT_PASSES * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool valid = i < p.n;
        const uint32_t active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            float4 q = P[i];
            int ca, cb, cc;
            uint32_t key = cell_coords(p, q.x, q.y, q.z, ca, cb, cc) ? (uint32_t)cell_key(p, ca, cb, cc) : (uint32_t)p.num_cells;
            keys[i] = key;
            // neighbouring slots share cells, hence digits: aggregate per warp before touching shared memory
            for (int k = 0; k < passes; k++) {
                uint32_t d = (key >> (k * RADIX_BITS)) & (RADIX - 1);
                uint32_t peers = __match_any_sync(active, d);
                if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[k * RADIX + d], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) {
        uint32_t v = s_hist[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive scan of one value per thread over a 256-thread block
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t *s_warp /*8*/, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
        uint32_t t = s_warp[w];
        if (w < warp) wbase += t;
        tot += t;
    }
    total = tot;
    __syncthreads();
    return wbase + inc - v;
}

// One LSD pass.  vin == nullptr means "values are the identity" (first pass).
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_pass(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
             uint32_t *__restrict__ vout, int n, int shift, const uint32_t *__restrict__ ghist /*256, this pass*/,
             uint32_t *tile_state /*tiles*256, zeroed*/, uint32_t *tile_counter /*zeroed*/) {
    __shared__ uint32_t s_warp_hist[SORT_WARPS][RADIX];
    __shared__ uint32_t s_local_start[RADIX];
    __shared__ uint32_t s_out_base[RADIX];
    __shared__ uint32_t s_keys[SORT_TILE];
    __shared__ uint32_t s_vals[SORT_TILE];
    __shared__ uint32_t s_scan[SORT_WARPS];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);  // ticket: lower tiles are always already running
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) s_warp_hist[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int base = (int)tile * SORT_TILE;
    const int tile_count = min(SORT_TILE, n - base);

    // warp-striped load: element e = warp*512 + r*32 + lane keeps tile order == (warp, round, lane) order
    uint32_t key[SORT_IPT], val[SORT_IPT];
    uint16_t rank[SORT_IPT];
#pragma unroll
    for (int r = 0; r < SORT_IPT; r++) {
        int e = warp * (32 * SORT_IPT) + r * 32 + lane;
        int idx = base + e;
        bool ok = e < tile_count;
        key[r] = ok ? kin[idx] : 0xffffffffu;  // pads rank last (stable) and are never written out
        val[r] = ok ? (vin ? vin[idx] : (uint32_t)idx) : 0u;
    }
    const uint32_t lanemask_lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_IPT; r++) {
        uint32_t d = (key[r] >> shift) & (RADIX - 1);
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (lane == leader) {
            prev = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[r] = (uint16_t)(prev + __popc(peers & lanemask_lt));
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: warp-exclusive bases and the tile total
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
        uint32_t t = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = tot;
        tot += t;
    }
    // decoupled look-back over earlier tiles for the global prefix of digit d
    uint32_t excl = 0;
    uint32_t *my = tile_state + (size_t)tile * RADIX + tid;
    if (tile == 0) {
        st_volatile(my, TS_FLAG_INCL | tot);
    } else {
        st_volatile(my, TS_FLAG_AGG | tot);
        int t = (int)tile - 1;
        while (true) {
            uint32_t s = ld_volatile(tile_state + (size_t)t * RADIX + tid);
            uint32_t flag = s & ~TS_MASK;
            if (flag == 0) continue;  // not published yet: spin
            excl += s & TS_MASK;
            if (flag == TS_FLAG_INCL) break;
            t--;
        }
        st_volatile(my, TS_FLAG_INCL | ((excl + tot) & TS_MASK));
    }
    // global start of digit d (exclusive scan of the global histogram) and local start inside the tile
    uint32_t dummy;
    uint32_t gstart = block_excl_scan_256(ghist[tid], s_scan, dummy);
    uint32_t lstart = block_excl_scan_256(tot, s_scan, dummy);
    s_local_start[tid] = lstart;
    s_out_base[tid] = gstart + excl - lstart;  // out index = s_out_base[d] + (position inside the sorted tile)
    __syncthreads();

#pragma unroll
    for (int r = 0; r < SORT_IPT; r++) {
        uint32_t d = (key[r] >> shift) & (RADIX - 1);
        uint32_t pos = s_local_start[d] + s_warp_hist[warp][d] + rank[r];
        s_keys[pos] = key[r];
        s_vals[pos] = val[r];
    }
    __syncthreads();
    for (int i = tid; i < tile_count; i += SORT_THREADS) {
        uint32_t k = s_keys[i];
        uint32_t d = (k >> shift) & (RADIX - 1);
        uint32_t o = s_out_base[d] + (uint32_t)i;
        kout[o] = k;
        vout[o] = s_vals[i];
    }
}

// strict mode: the reference's bucket order is ascending particle index (push_back order, cpp:207-212).
// One thread per sorted slot that starts a cell: insertion-sort that cell's source slots by original index.
// Slots [first, first + count) are examined (a cell that STARTS there is ordered completely, wherever it ends).
__global__ void k_cell_order_fix(const uint32_t *__restrict__ keys, uint32_t *vals, const int *__restrict__ id_src, int n, uint32_t num_cells,
                                 int first, int count) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    s += first;
    if (s >= n) return;
    uint32_t k = keys[s];
    if (k >= num_cells) return;  // the limbo bucket (outside the grid / dead entries) has no order to keep
    if (s > 0 && keys[s - 1] == k) return;
    int e = s + 1;
    while (e < n && keys[e] == k) e++;
    for (int i = s + 1; i < e; i++) {
        uint32_t v = vals[i];
        int vid = id_src[v];
        int j = i - 1;
        while (j >= s && id_src[vals[j]] > vid) {
            vals[j + 1] = vals[j];
            j--;
        }
        vals[j + 1] = v;
    }
}

// cell_start[c] = first sorted slot with key >= c, for c in [0, num_cells + 1]; cell_start[num_cells + 1] = n.
// Each boundary between different keys fills the gap of (possibly many) empty cells warp-cooperatively.
__global__ void __launch_bounds__(256) k_cell_bounds(const uint32_t *__restrict__ keys, int *__restrict__ cell_start, int n, int num_cells) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // slot s (0..n) closes the key gap (prev, cur]: prev = key[s-1] (or -1), cur = key[s] (or num_cells+1 at s == n)
    int prev = 0, cur = -1;
    if (s <= n) {
        prev = (s == 0) ? -1 : (int)keys[s - 1];
        cur = (s == n) ? num_cells + 1 : (int)keys[s];
    }
    int gap = cur - prev;  // number of cell_start entries this slot must write (0 for most slots)
    if (gap < 0) gap = 0;
    if (gap > 0 && gap <= 4) {
        for (int c = prev + 1; c <= cur; c++) cell_start[c] = s;
        gap = 0;
    }
    uint32_t pending = __ballot_sync(0xffffffffu, gap > 0);
    while (pending) {
        int src = __ffs(pending) - 1;
        pending &= pending - 1;
        int lo = __shfl_sync(0xffffffffu, prev, src) + 1;
        int hi = __shfl_sync(0xffffffffu, cur, src);
        int val = __shfl_sync(0xffffffffu, s, src);
        for (int c = lo + lane; c <= hi; c += 32) cell_start[c] = val;
    }
}

// ---- one-pass radix (counting) sort with the cell id as the single digit --------------------------------------------------
// Particles barely move between steps and a cell holds only a few of them, so instead of P LSD passes over 8-bit digits
// the production path sorts in ONE pass over the full key: count per cell (the atomic's return value is a provisional
// rank inside the cell), exclusive scan over the cells (which IS the cell_start table the neighbour passes need, so
// k_cell_bounds disappears), scatter, and a per-cell ordering by original index that makes the result deterministic and
// equal to the reference's bucket order (push_back order, cpp:207-212).  The LSD radix sort above remains the path for
// grids with far more cells than particles, where a pass over the cell table would cost more than sorting the keys.
// n_dev != nullptr: the entry count is *n_dev + n_add, read from device memory (slab step: the grid is sized for an upper bound)
__global__ void __launch_bounds__(256) k_cell_count(const __grid_constant__ DevParams p, const float4 *__restrict__ P, uint32_t *__restrict__ keys,
                                                    uint32_t *__restrict__ rank, uint32_t *__restrict__ cell_count, const int *__restrict__ n_dev,
                                                    int n_add) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < (n_dev ? *n_dev + n_add : p.n);
    uint32_t key = (uint32_t)p.num_cells;
    if (live) {
        const float4 q = P[i];
        int ca, cb, cc;
        if (cell_coords(p, q.x, q.y, q.z, ca, cb, cc)) key = (uint32_t)cell_key(p, ca, cb, cc);
    }
    // The limbo bucket (outside the grid; in slab mode every dead slot of the two message regions, tens of thousands of
    // CONSECUTIVE entries) is counted once per warp: one atomic per entry on that single address serialised the whole
    // kernel (78 us at 1M + 88k dead entries, against 52 us for 8M entries without them).
    const bool limbo = live && key == (uint32_t)p.num_cells;
    const unsigned lm = __ballot_sync(0xffffffffu, limbo);
    uint32_t r = 0;
    if (limbo) {
        const int lane = threadIdx.x & 31, leader = __ffs(lm) - 1;
        if (lane == leader) r = atomicAdd(&cell_count[key], (uint32_t)__popc(lm));
        r = __shfl_sync(lm, r, leader) + (uint32_t)__popc(lm & ((1u << lane) - 1u));
    } else if (live) {
        r = atomicAdd(&cell_count[key], 1u);
    }
    if (live) {
        keys[i] = key;
        rank[i] = r;
    }
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

// In-cell order = ascending ORIGINAL index.  Cells of up to BIG_CELL particles (a lattice holds 1-8 per cell) are put in that order
// by the gather itself (ordered_source); fuller cells (the reference's meshes reach 75) are listed by the scan, while their counts pass
// through, for k_cell_sort_big — the one-thread insertion sort of a 75-particle cell took 140 us of cfg2's 480 us step.
constexpr int BIG_CELL = 8;
// skey[slot] = the key again, in slot order (the in-cell ordering below permutes slots of ONE cell, so it stays valid): the
// neighbour passes read it instead of recomputing cell coordinates
__global__ void __launch_bounds__(256) k_cell_scatter(int n, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ rank,
                                                      const int *__restrict__ cell_start, uint32_t *__restrict__ vals, uint32_t *__restrict__ skey,
                                                      const int *__restrict__ n_dev, int n_add) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_dev ? *n_dev + n_add : n)) return;
    const uint32_t key = keys[i], slot = (uint32_t)cell_start[key] + rank[i];
    vals[slot] = (uint32_t)i;
    skey[slot] = key;
}
// cell_start[c] = exclusive prefix of the counts for c in [0, m], in ONE pass (decoupled look-back; rounds 1-2 ran tile sums / a
// single-block scan of them / apply as three launches: 16 us of the slab step at 8 GPUs): a block takes the next tile by ticket,
// publishes its tile sum, looks back over its predecessors' published sums / inclusive prefixes (a warp reads 32 of them at a time)
// and writes its slice of cell_start; the count table is zeroed for the next step; cells (not the limbo bucket m - 1) with more
// than BIG_CELL entries go on the worklist.  The published words carry the epoch of the launch — ctl[2], advanced by the last block to finish, which also rewinds
// the ticket — so nothing has to be cleared between launches and a captured graph can replay the kernel with frozen arguments.
//   ctl[0] ticket, ctl[1] blocks done, ctl[2] epoch;  state[tile] = epoch << 34 | status << 32 | value  (status 1: tile sum, 2: inclusive prefix)
//   big_count[2]: the worklist counter of this launch is big_count[epoch & 1]; the other one is cleared for the next launch
__device__ __forceinline__ unsigned long long scan_word(uint32_t epoch, uint32_t status, uint32_t value) {
    return ((unsigned long long)(epoch & 0x3fffffffu) << 34) | ((unsigned long long)status << 32) | value;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(uint32_t *__restrict__ counts, int m, int *__restrict__ cell_start,
                                                               unsigned long long *state, uint32_t *ctl, int *__restrict__ big_cells, int *big_count) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_tile, s_epoch, s_prefix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_epoch = *reinterpret_cast<volatile uint32_t *>(ctl + 2);
        s_tile = atomicAdd(ctl, 1u);
    }
    __syncthreads();
    const uint32_t epoch = s_epoch & 0x3fffffffu;
    const int tile = (int)s_tile;
    int *my_big = big_count + (epoch & 1u);
    if (tile == 0 && threadIdx.x == 0) big_count[(epoch + 1u) & 1u] = 0;
    const int base = tile * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t c[SCAN_IPT];
    const bool full = base + SCAN_IPT <= m;
    if (full) {
        const uint4 a = *reinterpret_cast<const uint4 *>(counts + base), b = *reinterpret_cast<const uint4 *>(counts + base + 4);
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) c[k] = base + k < m ? counts[base + k] : 0u;
    }
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) {
        tsum += c[k];
        if (c[k] > (uint32_t)BIG_CELL && base + k < m - 1) big_cells[atomicAdd(my_big, 1)] = base + k;
    }
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t total = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        uint32_t prefix = 0;
        if (tile > 0) {
            if (lane == 0) atomicExch(state + tile, scan_word(epoch, 1u, total));
            int j = tile - 1;
            unsigned spins = 0;
            while (true) {
                const int idx = j - lane;
                unsigned long long w = scan_word(epoch, 2u, 0u);  // (before tile 0: an inclusive prefix of nothing)
                bool ready;
                do {
                    if (idx >= 0) w = *reinterpret_cast<volatile unsigned long long *>(state + idx);
                    ready = (uint32_t)(w >> 34) == epoch && ((w >> 32) & 3u) != 0u;
                } while (__any_sync(0xffffffffu, !ready) && ++spins < (1u << 26));  // (the bound only guards against a hang: never reached)
                const uint32_t val = (uint32_t)w;
                const unsigned done = __ballot_sync(0xffffffffu, ((w >> 32) & 3u) == 2u);
                const int first = done ? __ffs(done) - 1 : 31;  // nearest predecessor that already holds an inclusive prefix
                uint32_t part = lane <= first ? val : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (done || spins >= (1u << 26)) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            atomicExch(state + tile, scan_word(epoch, 2u, prefix + total));
            s_prefix = prefix;
        }
    }
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++)
        if (w < warp) wbase += s_warp[w];
    uint32_t run = s_prefix + wbase + inc - tsum;
    if (full) {
        int o[SCAN_IPT];
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) {
            o[k] = (int)run;
            run += c[k];
        }
        *reinterpret_cast<int4 *>(cell_start + base) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4 *>(cell_start + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
        *reinterpret_cast<uint4 *>(counts + base) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(counts + base + 4) = make_uint4(0u, 0u, 0u, 0u);
        if (base + SCAN_IPT == m) cell_start[m] = (int)run;  // entry m (= total)
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) {
            if (base + k <= m) cell_start[base + k] = (int)run;  // entry m (= total) is written by the thread that owns it
            if (base + k < m) counts[base + k] = 0u;
            run += c[k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ctl + 1, 1u) == gridDim.x - 1) {  // the last block: rewind for the next launch
            ctl[0] = 0u;
            ctl[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(ctl + 2) = s_epoch + 1u;
        }
    }
}

// The source index that belongs in slot s once its cell is in canonical order: the member whose number of smaller original indices
// equals the slot's position in the cell (indices are unique).  Evaluated by the gather for its own slot — the separate pass over
// all cells (three dependent loads per cell, two launches) cost 21 us of the 306 us slab step at 8 GPUs and ~45 us at 8M on one —
// from vals / ids that neighbouring threads of the same cell load as well (L1 hits).  Big cells were ordered in place before.
__device__ __forceinline__ uint32_t ordered_source(const uint32_t *__restrict__ vals, const uint32_t *__restrict__ skey, const int *__restrict__ cell_start,
                                                   const int *__restrict__ id_src, int num_cells, int s) {
    uint32_t v = vals[s];
    const uint32_t key = skey[s];
    if (key >= (uint32_t)num_cells) return v;  // the limbo bucket (outside the grid / dead entries) has no order to keep
    const int cs = cell_start[key], cnt = cell_start[key + 1] - cs;
    if (cnt < 2 || cnt > BIG_CELL) return v;
    const int want = s - cs;
    if (cnt == 2) {
        const uint32_t u = vals[cs + 1 - want];  // the other member
        const bool v_first = id_src[v] < id_src[u];
        return (v_first == (want == 0)) ? v : u;  // in order already: keep; otherwise the two swap
    }
    uint32_t mv[BIG_CELL];
    int mid[BIG_CELL];
#pragma unroll
    for (int k = 0; k < BIG_CELL; k++) {
        mv[k] = k < cnt ? vals[cs + k] : 0u;
        mid[k] = k < cnt ? id_src[mv[k]] : INT_MAX;
    }
#pragma unroll
    for (int a = 0; a < BIG_CELL; a++) {
        int r = 0;
#pragma unroll
        for (int b = 0; b < BIG_CELL; b++) r += mid[b] < mid[a];
        if (a < cnt && r == want) v = mv[a];
    }
    return v;
}
// one warp per listed cell: every member's final position is the number of members with a smaller original index (indices are
// unique); ranks go to `tmp` first so that no lane overwrites an entry another lane still has to read.
__global__ void __launch_bounds__(256) k_cell_sort_big(const int *__restrict__ cell_start, uint32_t *vals, uint32_t *tmp, const int *__restrict__ id_src,
                                                       const int *__restrict__ big_cells, const int *__restrict__ big_count, const uint32_t *__restrict__ ctl) {
    const int lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = big_count[(ctl[2] - 1u) & 1u];  // the counter of the scan launch that has just finished (it advanced the epoch)
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < total; k += nwarps) {
        const int c = big_cells[k];
        const int s = cell_start[c], e = cell_start[c + 1];
        for (int m = s + lane; m < e; m += 32) {
            const uint32_t v = vals[m];
            const int vid = id_src[v];
            int rank = 0;
            for (int x = s; x < e; x++) rank += id_src[vals[x]] < vid;
            tmp[s + rank] = v;
        }
        __syncwarp();
        for (int m = s + lane; m < e; m += 32) vals[m] = tmp[m];
        __syncwarp();
    }
}
// (the other worklist counter is cleared by the scan kernel of the same step, for the next one)

// gather into the new slot order; `all` also permutes the intermediate / diagnostic arrays (after an upload, or
// in diagnostics mode, they are live across a re-sort)
// order_cells > 0: the slots of a cell are still in arrival order (counting sort); the canonical order is applied on the fly and the
// final permutation goes to vals_out
__global__ void __launch_bounds__(256) k_reorder(int n, const uint32_t *__restrict__ vals, Arrays src, Arrays dst, int all, const uint32_t *__restrict__ skey,
                                                 const int *__restrict__ cell_start, int order_cells, uint32_t *__restrict__ vals_out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t v;
    if (order_cells > 0) {
        v = ordered_source(vals, skey, cell_start, src.ID, order_cells, s);
        vals_out[s] = v;
    } else v = vals[s];
    const float4 p4 = src.P[v], e4 = src.E[v];
    dst.P[s] = p4;
    dst.VEL[s] = src.VEL[v];
    dst.O[s] = src.O[v];
    dst.E[s] = e4;
    dst.ID[s] = src.ID[v];
    dst.PB[s] = make_float4(p4.x, p4.y, p4.z, e4.x);
    if (all) {
        dst.C[s] = src.C[v];
        dst.V[s] = src.V[v];
        dst.S[s] = src.S[v];
        dst.ACC[s] = src.ACC[v];
        dst.GOAL[s] = src.GOAL[v];
        dst.PV[s] = src.PV[v];
    }
}

}  // namespace sphsm
