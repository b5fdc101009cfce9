// sphsm_comm.cuh — multi-GPU slab layer, device side (nothing in the reference corresponds to it).
//
// One process per GPU; rank r owns the cell planes [slab_lo, slab_hi) along the slab axis (perm[2], the slowest axis of
// the cell key) and keeps one halo plane on each interior side, so its sorted slot array is
//     [ left halo plane | owned planes ... | right halo plane | (dead) ]
// with every plane a contiguous slot range.  Per step:
//   k_mg_classify   stale halos are marked dead; owned particles whose NEW plane is the first / last owned plane or
//                   beyond are copied into the message for that neighbour (halo refresh and migration are the same
//                   message: the receiver's hash decides ownership) and stay here (as owned or as halo)
//   exchange 1      messages with the count in the header.  NCCL mode with CUDA-IPC neighbours: the packing kernel stores them
//                   straight into the neighbour's receive slot over NVLink and a flag word publishes them (push exchange, below);
//                   otherwise ncclSend / ncclRecv of the full capacity (or of a lagged estimate, x1_plan)
//   k_mg_unpack     (push exchange: waits for the neighbours' flag words, retires last step's halo copies) arrivals are appended
//                   behind the local particles; unused message slots become dead entries that the sort pushes into the limbo bucket
//   ... hash, sort (canonical in-cell order = ascending original index, so both sides of a slab face hold that plane in
//       the same order), cell table; k_mg_meta + one 32-byte read-back give the host the plane boundaries ...
//   exchange 2      after pass A the boundary planes' V / S records travel as contiguous slot ranges (ncclSend / ncclRecv)
//   k_p2p_allreduce the moment sums + error flag of the step, summed over the ranks through the same peer mappings
#pragma once
#include <limits.h>
#include <stddef.h>

#include "sphsm_types.cuh"

namespace sphsm {

// The slot ranges of a rank (live count, plane boundaries) live in DEVICE memory (SlabMeta, double-buffered: the sort of step
// t writes one, the kernels of step t+1 that still work on the pre-sort layout read it): every kernel of the slab step takes
// its ranges from there and its grid from a host-side upper bound, so the host never waits for a step to learn them.
struct SlabMeta {
    int n_live;        // live slots after the sort: [left halo | owned planes | right halo]
    int own_begin;     // first owned slot
    int b2;            // start of the 2nd owned plane
    int b3;            // start of the last owned plane
    int own_end;       // end of the owned slots
    int err0, err1;    // particles that crossed more than one plane / message overflows and face mismatches, so far
    int flag;          // sum over ranks of "this rank has an error", as it arrived with this step's moment allreduce
    int rng_all[4];    // launch ranges {begin, end, hole_begin, hole_len}: every owned slot
    int rng_int[4];    //   the interior planes [b2, b3)
    int rng_bnd[4];    //   the two boundary planes: [own_begin, own_end) minus the hole [b2, b3)
    int rng_int2[4];   //   the planes at least two planes away from both faces
    int rng_bnd2[4];   //   the two outermost owned planes on either side (everything an exchange-1 message can come from)
    int bound_viol;    // live slots exceeded the grid bound the host launched with
    int pad[3];        // (the ranges are read as int4: every one of them sits on a 16-byte boundary)
};
static_assert(offsetof(SlabMeta, rng_all) % 16 == 0 && offsetof(SlabMeta, rng_int2) % 16 == 0 && offsetof(SlabMeta, rng_bnd2) % 16 == 0, "int4 loads");
static_assert(sizeof(SlabMeta) == 128, "SlabMeta is 32 ints");

// one message = [count, pad x3] [P x cap] [VEL x cap] [O x cap] [E x cap] [ID x cap]
struct MsgView {
    int *count;
    float4 *P, *VEL, *O, *E;
    int *ID;
};
inline size_t msg_bytes(int cap) { return 16 + (size_t)cap * (4 * sizeof(float4) + sizeof(int)); }
inline MsgView msg_view(uint8_t *base, int cap) {
    MsgView v;
    v.count = reinterpret_cast<int *>(base);
    v.P = reinterpret_cast<float4 *>(base + 16);
    v.VEL = v.P + cap;
    v.O = v.VEL + cap;
    v.E = v.O + cap;
    v.ID = reinterpret_cast<int *>(v.E + cap);
    return v;
}

// ---- push exchange: the receive slots of a rank and the two words its neighbours publish into ---------------------------------
// block = [flag from left, flag from right, pad to 256 B][slot(side 0, parity 0)][slot(0, 1)][slot(1, 0)][slot(1, 1)], a slot is one
// message (header + arrays for the full halo capacity).  side 0 = written by the left neighbour, side 1 = by the right one.
// Behind the slots: the push allreduce's landing area, [parity][from rank][P2P_RED_MAX doubles]; its flags are words 16 + rank of
// the flag region.
constexpr size_t P2P_FLAGS_BYTES = 256;
constexpr int P2P_RED_MAX = 96, P2P_MAX_RANKS = 32, P2P_RED_FLAG0 = 16;
inline size_t p2p_slot_bytes(int cap) { return (msg_bytes(cap) + 255) / 256 * 256; }
inline size_t p2p_red_offset(int cap) { return P2P_FLAGS_BYTES + 4 * p2p_slot_bytes(cap); }
inline size_t p2p_block_bytes(int cap, int nranks) { return p2p_red_offset(cap) + (size_t)2 * nranks * P2P_RED_MAX * sizeof(double); }
inline uint8_t *p2p_slot(uint8_t *block, int cap, int side, int parity) { return block + P2P_FLAGS_BYTES + (size_t)(side * 2 + parity) * p2p_slot_bytes(cap); }
inline int *p2p_flag(uint8_t *block, int side) { return reinterpret_cast<int *>(block) + side; }

// after the packing kernel (same stream): the counts go into the headers of the neighbours' slots, then — behind a system-scope
// fence, so that whoever sees the sequence number also sees the counts and everything the packing kernel stored — the sequence
// number of this exchange into their flag words
__global__ void k_p2p_signal(const int *__restrict__ cntL, const int *__restrict__ cntR, int *hdrL, int *hdrR, int *flagL, int *flagR, int seq) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (hdrL) *reinterpret_cast<volatile int *>(hdrL) = *cntL;
    if (hdrR) *reinterpret_cast<volatile int *>(hdrR) = *cntR;
    __threadfence_system();
    if (flagL) *reinterpret_cast<volatile int *>(flagL) = seq;
    if (flagR) *reinterpret_cast<volatile int *>(flagR) = seq;
}
// (the receiver's side is in k_mg_unpack: its blocks poll this rank's own flag words until the neighbours have published exchange
// `seq`; a neighbour that never arrives — its process died, or the two sides lost count of the exchanges — ends the wait after a
// time-out with err[1] raised, so the step fails like any other exchange error instead of hanging the device)

// Push allreduce (sum of `count` <= P2P_RED_MAX doubles in place, every rank ends with the same bits): one block per rank stores its
// values into EVERY rank's landing area (its own included), publishes the sequence number in every rank's flag word, waits until
// all ranks have published theirs here, and adds the contributions in rank order.  One launch and one NVLink round trip instead
// of the ~45-60 us the 34-double ncclAllReduce of the moment sums took at 8 GPUs (SPHSM_TRACE) — that chain (sums -> allreduce ->
// solve) is what the gather waits for.  Two landing areas by parity: a rank can be at most one allreduce ahead of the slowest.
__global__ void __launch_bounds__(128) k_p2p_allreduce(double *totals, int count, int rank, int nranks, int seq, uint8_t *const *__restrict__ blocks,
                                                       size_t red_off, int *err, unsigned long long timeout_ns) {
    const int t = threadIdx.x, par = seq & 1;
    if (t < count) {
        const double v = totals[t];
        for (int p = 0; p < nranks; p++) {
            double *dst = reinterpret_cast<double *>(blocks[p] + red_off) + ((size_t)par * nranks + rank) * P2P_RED_MAX;
            *reinterpret_cast<volatile double *>(dst + t) = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t < nranks) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(reinterpret_cast<int *>(blocks[t]) + P2P_RED_FLAG0 + rank) = seq;
        const volatile int *f = reinterpret_cast<const int *>(blocks[rank]) + P2P_RED_FLAG0 + t;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        while (*f - seq < 0) {
            if ((++spins & 1023u) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now - t0 > timeout_ns) {  // a rank that never arrives: the step fails instead of hanging the device
                    atomicAdd(&err[1], 1);
                    break;
                }
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (t < count) {
        const volatile double *src = reinterpret_cast<const double *>(blocks[rank] + red_off) + (size_t)par * nranks * P2P_RED_MAX;
        double sum = 0.0;
        for (int p = 0; p < nranks; p++) sum += src[(size_t)p * P2P_RED_MAX + t];
        totals[t] = sum;
    }
}

// exchange 2 (pass A's records of a boundary plane) reuses the message buffers: [count, pad x3] [V x cap] [S x cap]
struct Msg2View {
    int *count;
    float4 *V;
    float2 *S;
};
inline size_t msg2_bytes(int cap) { return 16 + (size_t)cap * (sizeof(float4) + sizeof(float2)); }
inline Msg2View msg2_view(uint8_t *base, int cap) {
    Msg2View v;
    v.count = reinterpret_cast<int *>(base);
    v.V = reinterpret_cast<float4 *>(base + 16);
    v.S = reinterpret_cast<float2 *>(v.V + cap);
    return v;
}

// cell plane along the slab axis, evaluated exactly as the cell key does (float division, truncation); INT_MIN for a
// dead (NaN) entry
__device__ __forceinline__ int slab_plane(const DevParams &p, const float4 q) {
    const float x = p.perm[2] == 0 ? q.x : (p.perm[2] == 1 ? q.y : q.z);
    if (!(q.x == q.x) || !(x == x)) return INT_MIN;
    return __float2int_rz(__fdiv_rn(x, p.cell_size));
}
__device__ __forceinline__ bool slab_owned(const DevParams &p, const float4 q) {
    const int pl = slab_plane(p, q);
    return pl >= p.slab_lo && pl < p.slab_hi;
}

__device__ __forceinline__ void msg_put(const MsgView &m, int k, const Arrays &a, int s, const float4 q) {
    m.P[k] = q;
    m.VEL[k] = a.VEL[s];
    m.O[k] = a.O[s];
    m.E[k] = a.E[s];
    m.ID[k] = a.ID[s];
}

// err[0]: particles that crossed more than one plane in a step; err[1]: message overflow / face population mismatch
__global__ void __launch_bounds__(256) k_mg_classify(const __grid_constant__ DevParams p, Arrays a, int has_left, int has_right, MsgView L,
                                                     MsgView R, int capL, int capR, int *err, const SlabMeta *__restrict__ m) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m->n_live) return;
    if (s < m->own_begin || s >= m->own_end) {  // last step's halo copy: the owner sends a fresh one
        a.P[s].x = __int_as_float(0x7fc00000);
        return;
    }
    const float4 q = a.P[s];
    const int pl = slab_plane(p, q);
    if ((has_left && pl < p.slab_lo - 1) || (has_right && pl > p.slab_hi) || pl == INT_MIN) atomicAdd(&err[0], 1);
    if (has_left && pl <= p.slab_lo) {
        const int k = atomicAdd(L.count, 1);
        if (k < capL) msg_put(L, k, a, s, q);
        else atomicAdd(&err[1], 1);
    }
    if (has_right && pl >= p.slab_hi - 1) {
        const int k = atomicAdd(R.count, 1);
        if (k < capR) msg_put(R, k, a, s, q);
        else atomicAdd(&err[1], 1);
    }
}

// The same selection over a launch range (rng = {begin, end, hole_begin, hole_len} in device memory) with the positions given
// explicitly: run right after pass B on the two outermost planes of either side, on the freshly integrated positions `Pnew`,
// so that exchange 1 of the NEXT step travels while pass B still works on the interior planes.
__global__ void __launch_bounds__(256) k_mg_classify_rng(const __grid_constant__ DevParams p, Arrays a, const float4 *__restrict__ Pnew,
                                                         const int *__restrict__ rng, int has_left, int has_right, MsgView L, MsgView R, int capL,
                                                         int capR, int *err) {
    const int4 r = *reinterpret_cast<const int4 *>(rng);
    int s = r.x + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= r.z) s += r.w;
    if (s >= r.y) return;
    const float4 q = Pnew[s];
    const int pl = slab_plane(p, q);
    if ((has_left && pl < p.slab_lo - 1) || (has_right && pl > p.slab_hi) || pl == INT_MIN) atomicAdd(&err[0], 1);
    if (has_left && pl <= p.slab_lo) {
        const int k = atomicAdd(L.count, 1);
        if (k < capL) msg_put(L, k, a, s, q);
        else atomicAdd(&err[1], 1);
    }
    if (has_right && pl >= p.slab_hi - 1) {
        const int k = atomicAdd(R.count, 1);
        if (k < capR) msg_put(R, k, a, s, q);
        else atomicAdd(&err[1], 1);
    }
}
// What is left of k_mg_classify at the start of the next step when its exchange has already happened: last step's halo copies are
// dropped (inside k_mg_unpack), and an interior particle that now sits where it should have been sent (it crossed two planes) is
// an error
__global__ void __launch_bounds__(256) k_mg_check_interior(const __grid_constant__ DevParams p, const float4 *__restrict__ P, const SlabMeta *__restrict__ m,
                                                           int has_left, int has_right, int *err) {
    const int s = m->rng_int2[0] + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m->rng_int2[1]) return;
    const int pl = slab_plane(p, P[s]);
    if ((has_left && pl <= p.slab_lo) || (has_right && pl >= p.slab_hi - 1) || pl == INT_MIN) atomicAdd(&err[0], 1);
}

// arrivals -> slots [n0, n0 + 2*cap), n0 = the live count before this step's exchange: left message first; unused slots are dead.
// A message whose header count exceeds the capacity was truncated by its sender: the receiver flags it in the same step.
// capL / capR: the capacity this exchange's messages were laid out for (<= cap, see x1_plan).  Thread 0 files the populations of the
// exchange — what this rank sent (its send headers are still intact) and what it received — for the sizing of a later one.
// Push exchange: flagL / flagR != nullptr make every block poll this rank's flag words until the neighbours have published exchange
// `seq`, and drop != 0 retires last step's halo copies on the way (both were kernels of their own at first: three launches at the
// head of the slab step are one).
__global__ void __launch_bounds__(256) k_mg_unpack(const SlabMeta *__restrict__ prev, Arrays a, int has_left, int has_right, MsgView L, MsgView R,
                                                   int cap, int capL, int capR, int *err, const int *__restrict__ sentL, const int *__restrict__ sentR,
                                                   int *__restrict__ rec, const int *flagL, const int *flagR, int seq, unsigned long long timeout_ns,
                                                   int drop) {
    if (flagL || flagR) {
        if (threadIdx.x < 2) {
            const int *f = threadIdx.x == 0 ? flagL : flagR;
            if (f) {
                unsigned long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                unsigned spins = 0;
                while (*reinterpret_cast<const volatile int *>(f) - seq < 0) {
                    if ((++spins & 1023u) == 0) {
                        unsigned long long t;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                        if (t - t0 > timeout_ns) {
                            if (blockIdx.x == 0) atomicAdd(&err[1], 1);
                            break;
                        }
                    }
                }
                __threadfence_system();
            }
        }
        __syncthreads();
    }
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (drop && idx < 2 * cap) {
        const int side = idx >= cap, k = idx - side * cap;
        const int first = side ? prev->own_end : 0, cnt = side ? prev->n_live - prev->own_end : prev->own_begin;
        if (k < cnt) a.P[first + k].x = __int_as_float(0x7fc00000);
    }
    if (idx == 0) {
        rec[0] = has_left ? *sentL : 0;
        rec[1] = has_right ? *sentR : 0;
        rec[2] = has_left ? *L.count : 0;
        rec[3] = has_right ? *R.count : 0;
    }
    if (idx >= 2 * cap) return;
    const int n0 = prev->n_live;
    const int side = idx >= cap, k = idx - side * cap, slot = n0 + idx;
    const MsgView &m = side ? R : L;
    const int mcap = side ? capR : capL;
    const int sent = (side ? has_right : has_left) ? *m.count : 0;
    if (k == 0 && sent > mcap) atomicAdd(&err[1], 1);
    const int cnt = min(sent, mcap);
    if (k < cnt) {
        a.P[slot] = m.P[k];
        a.VEL[slot] = m.VEL[k];
        a.O[slot] = m.O[k];
        a.E[slot] = m.E[k];
        a.ID[slot] = m.ID[k];
    } else {
        a.P[slot] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 1.f);
        a.VEL[slot] = make_float4(0.f, 0.f, 0.f, 1.f);
        a.O[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        a.E[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
        a.ID[slot] = -1;
    }
}

// sphsm_comm_set_slab: everything outside the owned planes is dropped
__global__ void __launch_bounds__(256) k_mg_filter(const __grid_constant__ DevParams p, Arrays a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n) return;
    if (!slab_owned(p, a.P[s])) a.P[s].x = __int_as_float(0x7fc00000);
}

// the plane boundaries of the freshly sorted array.  flag_src: where this step's allreduce left the summed error flag (nullptr:
// none); n_bound: the slot count the host sized this step's grids for.
__global__ void k_mg_meta(const int *__restrict__ cell_start, int num_cells, int plane_cells, int gcl, const int *__restrict__ err,
                          const double *__restrict__ flag_src, int n_bound, SlabMeta *m) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = cell_start[num_cells], ob = cell_start[plane_cells], b2 = cell_start[plane_cells * 2];
    const int b3 = cell_start[plane_cells * (gcl - 2)], oe = cell_start[plane_cells * (gcl - 1)];
    m->n_live = n; m->own_begin = ob; m->b2 = b2; m->b3 = b3; m->own_end = oe;
    m->err0 = err[0]; m->err1 = err[1] + (n > n_bound ? 1 : 0);  // (more live slots than the grids were sized for: treated like an overflow)
    m->flag = flag_src ? (int)*flag_src : 0;
    m->rng_all[0] = ob; m->rng_all[1] = oe; m->rng_all[2] = 0; m->rng_all[3] = 0;
    const bool three = b3 > b2;  // at least three populated owned planes: an interior exists
    m->rng_int[0] = b2; m->rng_int[1] = three ? b3 : b2; m->rng_int[2] = 0; m->rng_int[3] = 0;
    m->rng_bnd[0] = ob; m->rng_bnd[1] = oe; m->rng_bnd[2] = three ? b2 : 0; m->rng_bnd[3] = three ? b3 - b2 : 0;
    m->bound_viol = n > n_bound ? 1 : 0;
    // two planes per side (a particle moves at most one plane per step, so whatever must be sent to a neighbour after the step
    // sits in one of them): [own_begin, p2) and [p3, own_end), p2 / p3 = start of the 3rd / of the last-but-one owned plane
    const int p2 = cell_start[plane_cells * min(3, gcl - 1)], p3 = cell_start[plane_cells * max(gcl - 3, 1)];
    const bool five = gcl - 2 >= 5 && p3 > p2;
    m->rng_int2[0] = p2; m->rng_int2[1] = five ? p3 : p2; m->rng_int2[2] = 0; m->rng_int2[3] = 0;
    m->rng_bnd2[0] = ob; m->rng_bnd2[1] = oe; m->rng_bnd2[2] = five ? p2 : 0; m->rng_bnd2[3] = five ? p3 - p2 : 0;
}

// exchange 2, sender side: pass A's records (V = inter_vel + m/dens, S = pres + dens) of the first owned plane go to the left
// neighbour, those of the last owned plane to the right one, with the plane population in the header
__global__ void __launch_bounds__(256) k_mg_pack2(const SlabMeta *__restrict__ m, Arrays a, int has_left, int has_right, Msg2View L, Msg2View R, int cap) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * cap) return;
    const int side = idx >= cap, k = idx - side * cap;
    if (side ? !has_right : !has_left) return;
    const int first = side ? m->b3 : m->own_begin, cnt = side ? m->own_end - m->b3 : m->b2 - m->own_begin;
    const Msg2View &o = side ? R : L;
    if (k == 0) *o.count = cnt;
    if (k < min(cnt, cap)) {
        o.V[k] = a.V[first + k];
        o.S[k] = a.S[first + k];
    }
}
// receiver side: the left halo plane is slots [0, own_begin), the right one [own_end, n_live); both sides of a face hold the
// shared plane in the same (canonical) order.  A population that differs from the sender's is an error (err[1]).
// zero0 / zero1: the headers of the two SEND buffers (free again: this kernel runs behind exchange 2), cleared for the early
// exchange-1 packing that may follow
__global__ void __launch_bounds__(256) k_mg_unpack2(const SlabMeta *__restrict__ m, Arrays a, int has_left, int has_right, Msg2View L, Msg2View R,
                                                    int cap, int *err, int *zero0, int *zero1) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) { *zero0 = 0; *zero1 = 0; }
    if (idx >= 2 * cap) return;
    const int side = idx >= cap, k = idx - side * cap;
    if (side ? !has_right : !has_left) return;
    const int first = side ? m->own_end : 0, cnt = side ? m->n_live - m->own_end : m->own_begin;
    const Msg2View &in = side ? R : L;
    if (k == 0 && *in.count != cnt) atomicAdd(&err[1], 1);
    if (k < min(cnt, cap)) {
        const float4 v = in.V[k];
        a.V[first + k] = v;
        a.S[first + k] = in.S[k];
        a.VN[first + k] = v.w;  // the dense copy of V.w that pass B's phase 1 gathers
    }
}

// compact (id, xyz) of the owned slots for sphsm_download_owned; the range comes from device memory (rng = {begin, end}) and the
// count is left in *count_out for the asynchronous form
__global__ void __launch_bounds__(256) k_mg_owned_out(const int *__restrict__ rng, int first_h, int count_h, int cap, Arrays a, int *__restrict__ ids,
                                                      float *__restrict__ xyz, int *__restrict__ count_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int first = rng ? rng[0] : first_h, count = rng ? rng[1] - rng[0] : count_h;
    if (k == 0 && count_out) *count_out = count;
    if (k >= count || k >= cap) return;
    const float4 q = a.P[first + k];
    ids[k] = a.ID[first + k];
    xyz[3 * (size_t)k] = q.x;
    xyz[3 * (size_t)k + 1] = q.y;
    xyz[3 * (size_t)k + 2] = q.z;
}

}  // namespace sphsm
