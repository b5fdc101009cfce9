// sphsm_comm.cuh — multi-GPU slab layer, device side (nothing in the reference corresponds to it).
//
// One process per GPU; rank r owns the cell planes [slab_lo, slab_hi) along the slab axis (perm[2], the slowest axis of
// the cell key) and keeps one halo plane on each interior side, so its sorted slot array is
//     [ left halo plane | owned planes ... | right halo plane | (dead) ]
// with every plane a contiguous slot range.  Per step:
//   k_mg_classify   stale halos are marked dead; owned particles whose NEW plane is the first / last owned plane or
//                   beyond are copied into the message for that neighbour (halo refresh and migration are the same
//                   message: the receiver's hash decides ownership) and stay here (as owned or as halo)
//   exchange 1      fixed-capacity messages (count in the header), ncclSend / ncclRecv
//   k_mg_unpack     arrivals are appended behind the local particles; unused message slots become dead entries that the
//                   sort pushes into the limbo bucket
//   ... hash, sort (canonical in-cell order = ascending original index, so both sides of a slab face hold that plane in
//       the same order), cell table; k_mg_meta + one 32-byte read-back give the host the plane boundaries ...
//   exchange 2      after pass A the boundary planes' V / S records travel as contiguous slot ranges
#pragma once
#include <limits.h>

#include "sphsm_types.cuh"

namespace sphsm {

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

// err[0]: particles that crossed more than one plane in a step; err[1]: message overflow
__global__ void __launch_bounds__(256) k_mg_classify(const __grid_constant__ DevParams p, Arrays a, int has_left, int has_right, MsgView L,
                                                     MsgView R, int cap, int *err) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n) return;
    if (s < p.own_begin || s >= p.own_end) {  // last step's halo copy: the owner sends a fresh one
        a.P[s].x = __int_as_float(0x7fc00000);
        return;
    }
    const float4 q = a.P[s];
    const int pl = slab_plane(p, q);
    if ((has_left && pl < p.slab_lo - 1) || (has_right && pl > p.slab_hi) || pl == INT_MIN) atomicAdd(&err[0], 1);
    if (has_left && pl <= p.slab_lo) {
        const int k = atomicAdd(L.count, 1);
        if (k < cap) msg_put(L, k, a, s, q);
        else atomicAdd(&err[1], 1);
    }
    if (has_right && pl >= p.slab_hi - 1) {
        const int k = atomicAdd(R.count, 1);
        if (k < cap) msg_put(R, k, a, s, q);
        else atomicAdd(&err[1], 1);
    }
}

// arrivals -> slots [n0, n0 + 2*cap): left message first; unused slots are dead
__global__ void __launch_bounds__(256) k_mg_unpack(int n0, Arrays a, int has_left, int has_right, MsgView L, MsgView R, int cap) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * cap) return;
    const int side = idx >= cap, k = idx - side * cap, slot = n0 + idx;
    const MsgView &m = side ? R : L;
    const int cnt = (side ? has_right : has_left) ? min(*m.count, cap) : 0;
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

// meta[0] live slots, [1] own_begin, [2] start of the 2nd owned plane, [3] start of the last owned plane, [4] own_end,
// [5] err0, [6] err1, [7] spare
__global__ void k_mg_meta(const int *__restrict__ cell_start, int num_cells, int plane_cells, int gcl, const int *__restrict__ err, int *meta) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    meta[0] = cell_start[num_cells];
    meta[1] = cell_start[plane_cells];
    meta[2] = cell_start[plane_cells * 2];
    meta[3] = cell_start[plane_cells * (gcl - 2)];
    meta[4] = cell_start[plane_cells * (gcl - 1)];
    meta[5] = err[0];
    meta[6] = err[1];
    meta[7] = 0;
}

// exchange 2 carries V = (inter_vel, m/dens) and S; the dense copy of V.w that pass B's phase 1 gathers is rebuilt here for
// the two halo ranges [0, n_left) and [right_begin, right_begin + n_right)
__global__ void __launch_bounds__(256) k_mg_halo_vn(int n_left, int right_begin, int n_right, const float4 *__restrict__ V, float *__restrict__ VN) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_left + n_right) return;
    const int s = k < n_left ? k : right_begin + (k - n_left);
    VN[s] = V[s].w;
}

__global__ void k_store_double(double *dst, double v) { *dst = v; }

// compact (id, xyz) of the owned slots for sphsm_download_owned
__global__ void __launch_bounds__(256) k_mg_owned_out(int first, int count, Arrays a, int *__restrict__ ids, float *__restrict__ xyz) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const float4 q = a.P[first + k];
    ids[k] = a.ID[first + k];
    xyz[3 * (size_t)k] = q.x;
    xyz[3 * (size_t)k + 1] = q.y;
    xyz[3 * (size_t)k + 2] = q.z;
}

}  // namespace sphsm
