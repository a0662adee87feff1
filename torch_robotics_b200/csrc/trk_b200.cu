// trk_b200.cu - runtime-tree ("generic") kernel family and the C ABI (include/trk_b200.h).
//
// One thread evaluates one configuration.  The kinematic tree and the cost tables are staged once
// per block in shared memory and read as warp-uniform broadcasts; per-thread link poses live in a
// thread-local array indexed by body (the tree is runtime data here - compiled-in robots use the
// register-resident family in trk_static.cuh).  Dense FK output / upstream-gradient tiles are
// transposed through a per-warp shared-memory stage so that every global access is a 16-byte
// vector covering whole 64-byte link matrices (8 configurations per instruction).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "trk_device.cuh"
#include "trk_static.cuh"

using namespace trk;

// (the per-warp transpose stage lives in trk_static.cuh and is shared by both families)

constexpr int kBlock = 128;
constexpr int kWarps = kBlock / 32;

// ------------------------------------------------------------------------------------------
// generic kernels
// ------------------------------------------------------------------------------------------
// Shared memory of the runtime-tree kernels: tree words | [cost words] | per warp { stage (2 KB) | live pose slots
// (n_live x 12 x 32) | joint columns (D x 12 x 32, reverse mode only) | pair-position slots (n_pslots x 3 x 32) }.
__host__ __device__ __forceinline__ int generic_warp_words(int stage_words, int n_live, int n_joint_cols, int n_pslots) {
    return stage_words + (n_live * 12 + n_joint_cols * 12 + n_pslots * 3) * 32;
}

__global__ void __launch_bounds__(kBlock) k_fk_forward_generic(const float* __restrict__ tree_words, int tree_nwords,
                                                               const __grid_constant__ TreePlan plan, const float* __restrict__ q, long long n,
                                                               const __grid_constant__ SlotMap slots, int n_sel, float* __restrict__ H) {
    extern __shared__ __align__(16) float smem[];
    float* tw = smem;
    block_copy_words(tw, tree_words, tree_nwords);
    __syncthreads();
    const TreeHeader* hdr = reinterpret_cast<const TreeHeader*>(tw);
    const LinkRec* recs = reinterpret_cast<const LinkRec*>(tw + 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthreads = blockDim.x, D = hdr->D;
    float* wbase = smem + tree_nwords + (size_t)warp * generic_warp_words(512, plan.n_live, 0, 0);
    float4* stage = reinterpret_cast<float4*>(wbase);
    float* wl = wbase + 512 + lane;
    for (long long base = (long long)blockIdx.x * nthreads; base < n; base += (long long)gridDim.x * nthreads) {
        const long long cfg = base + threadIdx.x;
        const long long warp_cfg0 = base + (threadIdx.x & ~31);
        if (warp_cfg0 >= n) continue;   // warp-uniform
        fk_live(hdr, recs, plan, q + (cfg < n ? cfg : n - 1) * D, wl, 32,
                [&](int, const LinkRec& r, const float* w, const float*, bool) {
                    const int s = slots.slot[r.body];
                    if (s >= 0)
                        warp_store_link(stage, lane, make_float4(w[0], w[1], w[2], w[9]), make_float4(w[3], w[4], w[5], w[10]),
                                        make_float4(w[6], w[7], w[8], w[11]), warp_cfg0, n, n_sel, s, H);
                });
    }
}

__global__ void __launch_bounds__(kBlock) k_fk_backward_generic(const float* __restrict__ tree_words, int tree_nwords,
                                                                const __grid_constant__ TreePlan plan, const float* __restrict__ q, long long n,
                                                                const __grid_constant__ SlotMap slots, int n_sel, const float* __restrict__ G,
                                                                float* __restrict__ grad_q) {
    extern __shared__ __align__(16) float smem[];
    float* tw = smem;
    block_copy_words(tw, tree_words, tree_nwords);
    __syncthreads();
    const TreeHeader* hdr = reinterpret_cast<const TreeHeader*>(tw);
    const LinkRec* recs = reinterpret_cast<const LinkRec*>(tw + 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthreads = blockDim.x, D = hdr->D;
    float* wbase = smem + tree_nwords + (size_t)warp * generic_warp_words(1024, plan.n_live, D, 0);
    float4* stage = reinterpret_cast<float4*>(wbase);      // two 2-KB buffers: the next body's G streams in while this one is used
    float* wl = wbase + 1024 + lane;
    float* JS = wbase + 1024 + plan.n_live * 12 * 32 + lane;
    const int L = hdr->L;
    for (long long base = (long long)blockIdx.x * nthreads; base < n; base += (long long)gridDim.x * nthreads) {
        const long long cfg = base + threadIdx.x;
        const long long warp_cfg0 = base + (threadIdx.x & ~31);
        if (warp_cfg0 >= n) continue;
        unsigned pass_mask = 0u;
        // a body's G is fetched only if it will be consumed (selected and below some joint): every stage buffer that
        // is written is also waited for before it is written again
        auto fetch = [&](int pos) {
            const int bn = recs[pos].body, sn = slots.slot[bn];
            if (sn >= 0 && plan.jm[bn] >= 0) warp_prefetch_link(stage + (pos & 1) * 128, lane, warp_cfg0, n, n_sel, sn, G);
            cp_async_commit();
        };
        fetch(0);
        fk_live(hdr, recs, plan, q + (cfg < n ? cfg : n - 1) * D, wl, 32,
                [&](int i, const LinkRec& r, const float* w, const float* wp, bool pass) {
                    if (i + 1 < L) fetch(i + 1); else cp_async_commit();
                    if (r.qidx >= 0) {
                        joint_store_axis(JS, r, w, wp);
                        if (pass) pass_mask |= 1u << r.qidx;
                    }
                    const int s = slots.slot[r.body], c = plan.jm[r.body];
                    if (s < 0 || c < 0) return;     // not selected, or no joint above: no gradient (warp-uniform)
                    cp_async_wait<1>();
                    __syncwarp();
                    const float4* st = stage + (i & 1) * 128;
                    const float4 m0 = st[stage_idx(lane, 0)], m1 = st[stage_idx(lane, 1)], m2 = st[stage_idx(lane, 2)];
                    __syncwarp();
                    const float gR[9] = {m0.x, m0.y, m0.z, m1.x, m1.y, m1.z, m2.x, m2.y, m2.z};
                    float ax, ay, az;
                    axial_gR_Rt(gR, w, ax, ay, az);
                    joint_fold(JS, c, w[9], w[10], w[11], m0.w, m1.w, m2.w, ax, ay, az);
                });
        cp_async_wait<0>();
        if (cfg < n) joint_reverse_sweep(JS, plan, D, pass_mask, grad_q + cfg * D);
    }
}

// ------------------------------------------------------------------------------------------
// Geometric Jacobian of one link (compute_forward_kinematics_and_geometric_jacobian, robot_tree.py:218-248):
//   lin[:, c] = z_c x (p_link - p_c),  ang[:, c] = z_c   for every q column c in `col_mask`, zero elsewhere,
// z_c = the joint-axis column of the world rotation of the body q column c drives (the reference takes the column of
// the axis' non-zero component, sign and joint type ignored: rigid_body.py:97-99), p_c that body's world position.
// One thread per configuration; z_c / p_c wait in per-lane shared-memory columns until the link's position is known.
// Outputs: pose [N][12] (R row-major | p), lin / ang [N][3][D].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_geometric_jacobian(const float* __restrict__ tree_words, int tree_nwords,
                                                               const __grid_constant__ TreePlan live, const float* __restrict__ q, long long n,
                                                               int link_body, unsigned col_mask, float* __restrict__ pose,
                                                               float* __restrict__ lin, float* __restrict__ ang) {
    extern __shared__ __align__(16) float smem[];
    float* tw = smem;
    block_copy_words(tw, tree_words, tree_nwords);
    __syncthreads();
    const TreeHeader* hdr = reinterpret_cast<const TreeHeader*>(tw);
    const LinkRec* recs = reinterpret_cast<const LinkRec*>(tw + 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, D = hdr->D;
    float* wbase = smem + tree_nwords + (size_t)warp * (live.n_live * 12 + D * 6) * 32;
    float* wl = wbase + lane;
    float* jc = wbase + live.n_live * 12 * 32 + lane;       // [(c * 6 + k) * 32]: k 0-2 axis column, 3-5 position
    for (long long cfg = (long long)blockIdx.x * blockDim.x + threadIdx.x; cfg < n; cfg += (long long)gridDim.x * blockDim.x) {
        float pe[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) pe[k] = 0.f;
        fk_live(hdr, recs, live, q + cfg * D, wl, 32, [&](int, const LinkRec& r, const float* w, const float*, bool) {
            if (r.qidx >= 0) {
                int k = (r.flags >> F_AXIS_SHIFT) & 3;
                if ((r.flags & F_KIND_MASK) == TRK_KIND_PRISMATIC) k = r.ax[0] != 0.f ? 0 : (r.ax[1] != 0.f ? 1 : 2);
                float* o = jc + r.qidx * 6 * 32;
                o[0] = k == 0 ? w[0] : (k == 1 ? w[1] : w[2]);
                o[32] = k == 0 ? w[3] : (k == 1 ? w[4] : w[5]);
                o[64] = k == 0 ? w[6] : (k == 1 ? w[7] : w[8]);
                o[96] = w[9]; o[128] = w[10]; o[160] = w[11];
            }
            if (r.body == link_body) {
#pragma unroll
                for (int k = 0; k < 12; ++k) pe[k] = w[k];
            }
        });
        float* po = pose + cfg * 12;
#pragma unroll
        for (int k = 0; k < 12; ++k) po[k] = pe[k];
        float* lo = lin + cfg * 3 * D;
        float* ao = ang + cfg * 3 * D;
        for (int c = 0; c < D; ++c) {
            float zx = 0.f, zy = 0.f, zz = 0.f, lx = 0.f, ly = 0.f, lz = 0.f;
            if ((col_mask >> c) & 1u) {
                const float* o = jc + c * 6 * 32;
                zx = o[0]; zy = o[32]; zz = o[64];
                const float dx = pe[9] - o[96], dy = pe[10] - o[128], dz = pe[11] - o[160];
                lx = zy * dz - zz * dy; ly = zz * dx - zx * dz; lz = zx * dy - zy * dx;
            }
            lo[c] = lx; lo[D + c] = ly; lo[2 * D + c] = lz;
            ao[c] = zx; ao[D + c] = zy; ao[2 * D + c] = zz;
        }
    }
}

template <bool WANT_GRAD>
__global__ void __launch_bounds__(kBlock) k_objective_generic(const float* __restrict__ tree_words, int tree_nwords,
                                                              const float* __restrict__ cost_words, int cost_nwords,
                                                              const __grid_constant__ TreePlan plan, const float* __restrict__ q, long long n,
                                                              const float* __restrict__ ee_target_cfg, float margin_override,
                                                              int use_override, int joint, float* __restrict__ cost_total,
                                                              float* __restrict__ cost_terms, float* __restrict__ grad_q,
                                                              uint8_t* __restrict__ occupancy, int* __restrict__ argmin) {
    // joint != 0: the scene is one equal-radius identity-pose sphere object without relu (decided on the host): every
    // referenced body's position is parked in a slot during the sweep and the collision points are evaluated afterwards
    // through the joint scan in groups of five; otherwise a point is evaluated while its body's pose is in registers
    // and only the bodies of self-collision pairs are parked.
    extern __shared__ __align__(16) float smem[];
    float* tw = smem;
    float* cw = smem + tree_nwords;
    block_copy_words(tw, tree_words, tree_nwords);
    block_copy_words(cw, cost_words, cost_nwords);
    __syncthreads();
    const TreeHeader* hdr = reinterpret_cast<const TreeHeader*>(tw);
    const LinkRec* recs = reinterpret_cast<const LinkRec*>(tw + 16);
    const CostView cv(cw);
    const CostHeader* h = cv.h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthreads = blockDim.x, D = hdr->D;
    const int njc = WANT_GRAD ? D : 0;
    const int n_pos = joint ? h->n_slots : h->n_pslots;
    int* slot_jm = reinterpret_cast<int*>(smem + tree_nwords + ((cost_nwords + 3) & ~3));   // [TRK_MAX_LINKS] joint column above a slot's body
    if (threadIdx.x == 0)
        for (int b = 0; b < hdr->L; ++b) {
            const int sl = joint ? cv.slot_of_body[b] : cv.pslot_of_body[b];
            if (sl >= 0) slot_jm[sl] = plan.jm[b];
        }
    __syncthreads();
    float* wbase = smem + tree_nwords + ((cost_nwords + 3) & ~3) + TRK_MAX_LINKS + (size_t)warp * generic_warp_words(0, plan.n_live, njc, n_pos);
    float* wl = wbase + lane;
    float* JS = wbase + plan.n_live * 12 * 32 + lane;
    float* PS = wbase + (plan.n_live + njc) * 12 * 32 + lane;           // pair-position slots [slot * 3 + k][lane]
    const int S = h->S;
    for (long long cfg = (long long)blockIdx.x * nthreads + threadIdx.x; cfg < n; cfg += (long long)gridDim.x * nthreads) {
        float terms[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned occ = 0u, pass_mask = 0u;
        int* argmin_row = argmin ? argmin + cfg * (long long)(h->nObj * S) : nullptr;
        // forward: every body is seen once with its pose in registers; its collision points are evaluated on the spot
        // and their forces go straight to the accumulator of the joint above
        fk_live(hdr, recs, plan, q + cfg * D, wl, 32, [&](int, const LinkRec& r, const float* w, const float* wp, bool pass) {
            const int b = r.body;
            if (WANT_GRAD && r.qidx >= 0) {
                joint_store_axis(JS, r, w, wp);
                if (pass) pass_mask |= 1u << r.qidx;
            }
            const float x = w[9], y = w[10], z = w[11];
            const int ps = joint ? cv.slot_of_body[b] : cv.pslot_of_body[b];
            if (ps >= 0) { PS[ps * 96] = x; PS[ps * 96 + 32] = y; PS[ps * 96 + 64] = z; }
            const int c = plan.jm[b];
            if (!joint) {
#pragma unroll 1
                for (int k = cv.bp_start[b]; k < cv.bp_start[b + 1]; ++k) {
                    const int s = cv.bp_list[k];
                    float gx = 0.f, gy = 0.f, gz = 0.f;
                    if (h->nObj > 0) {
                        const float m = cv.pt_mobj[s];
                        terms[1] = __fadd_rn(terms[1], object_point_term<WANT_GRAD, true>(cv, x, y, z, m, use_override ? margin_override : m,
                                                                                          occ, argmin_row, s, gx, gy, gz));
                    }
                    if (h->ws_enabled) {
                        const float m = cv.pt_mws[s];
                        terms[2] = __fadd_rn(terms[2], ws_point_term<WANT_GRAD, true>(h, x, y, z, m, use_override ? margin_override : m, occ,
                                                                                      gx, gy, gz));
                    }
                    if (WANT_GRAD && c >= 0) joint_fold(JS, c, x, y, z, gx, gy, gz, 0.f, 0.f, 0.f);
                }
            }
            if (h->ee_enabled && b == h->ee_link) {
                float Rt[9], pt[3], gR[9], gp[3];
                if (ee_target_cfg) {
                    const float* Ht = ee_target_cfg + cfg * 16;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        Rt[i * 3 + 0] = __ldg(Ht + i * 4 + 0); Rt[i * 3 + 1] = __ldg(Ht + i * 4 + 1);
                        Rt[i * 3 + 2] = __ldg(Ht + i * 4 + 2); pt[i] = __ldg(Ht + i * 4 + 3);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Rt[i] = h->ee_R[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) pt[i] = h->ee_t[i];
                }
                terms[3] = ee_term<WANT_GRAD>(h, w, w + 9, Rt, pt, gR, gp);
                if (WANT_GRAD && c >= 0) {
                    float ax, ay, az;
                    axial_gR_Rt(gR, w, ax, ay, az);
                    joint_fold(JS, c, x, y, z, gp[0], gp[1], gp[2], ax, ay, az);
                }
            }
        });
        // self-collision pairs need both bodies: positions come from the slots; forces go to the joint above each body
        auto slot_pos = [&](int sl, float& x, float& y, float& z) { x = PS[sl * 96]; y = PS[sl * 96 + 32]; z = PS[sl * 96 + 64]; };
        auto slot_acc = [&](int sl, float gx, float gy, float gz) {
            const int c = slot_jm[sl];
            if (c >= 0) joint_fold(JS, c, PS[sl * 96], PS[sl * 96 + 32], PS[sl * 96 + 64], gx, gy, gz, 0.f, 0.f, 0.f);
        };
        terms[0] = self_pairs_term<WANT_GRAD, true>(cv, joint ? cv.pair_si : cv.pair_pi, joint ? cv.pair_sj : cv.pair_pj, slot_pos, slot_acc,
                                                    margin_override, use_override != 0, occ);
        if (joint) {
#pragma unroll 1
            for (int s0 = 0; s0 < S; s0 += 5)
                joint_points_terms_rt<WANT_GRAD, true, 5>(
                    cv, s0, (S - s0) < 5 ? (S - s0) : 5,
                    [&](int s, float& x, float& y, float& z) { slot_pos(cv.pt_slot[s], x, y, z); },
                    [&](int s, float gx, float gy, float gz) { slot_acc(cv.pt_slot[s], gx, gy, gz); },
                    margin_override, use_override != 0, terms[1], terms[2], occ, argmin_row);
        }
        if (cost_total) cost_total[cfg] = weighted_total(h, terms);
        if (cost_terms) reinterpret_cast<float4*>(cost_terms)[cfg] = make_float4(terms[0], terms[1], terms[2], terms[3]);
        if (occupancy) occupancy[cfg] = (uint8_t)occ;
        if (WANT_GRAD) joint_reverse_sweep(JS, plan, D, pass_mask, grad_q + cfg * D);
    }
}

// ------------------------------------------------------------------------------------------
// Link Jacobians (DifferentiableTree.compute_analytical_jacobian_all_links, robot_tree.py:250-265):
// J[cfg][body][7][D] = d (position xyz, quaternion wxyz) / d q.  The reference differentiates FK +
// rotation_matrix_to_q (quaternion.py:135-166) by autograd; in closed form, for joint j on the path root -> body,
//   revolute : d p = a_j x (p - p_j),  d quat = 1/2 (0, a_j) (x) quat        prismatic: d p = a_j, d quat = 0
// with a_j the world joint axis times the clamp mask, zero elsewhere.  Output-bound (L*7*D*4 B per configuration):
// phase 1, one thread per configuration: FK, quaternions and masked axes into a per-warp shared-memory record;
// phase 2, the warp writes its 32 configurations' blocks cooperatively, consecutive lanes -> consecutive floats.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_link_jacobian(const float* __restrict__ tree_words, int tree_nwords,
                                                          const __grid_constant__ TreePlan live, const float* __restrict__ q, long long n,
                                                          int cpw, int rec_stride, int link_chunk, float* __restrict__ J) {
    extern __shared__ __align__(16) float smem[];
    float* tw = smem;
    block_copy_words(tw, tree_words, tree_nwords);
    unsigned* anc = reinterpret_cast<unsigned*>(smem + tree_nwords);          // [TRK_MAX_LINKS] q columns on the path root -> body
    int* jbody = reinterpret_cast<int*>(anc + TRK_MAX_LINKS);                 // [32] body driven by q column j
    int* jprism = jbody + 32;                                                  // [32] 1 = prismatic
    float* stage = reinterpret_cast<float*>(jprism + 32);
    __syncthreads();
    const TreeHeader* hdr = reinterpret_cast<const TreeHeader*>(tw);
    const LinkRec* recs = reinterpret_cast<const LinkRec*>(tw + 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, D = hdr->D, L = hdr->L;
    if (threadIdx.x == 0) {
        for (int i = 0; i < L; ++i) {
            const LinkRec& r = recs[i];
            anc[r.body] = (r.parent < 0 ? 0u : anc[r.parent]) | (r.qidx >= 0 ? 1u << r.qidx : 0u);
            if (r.qidx >= 0) { jbody[r.qidx] = r.body; jprism[r.qidx] = (r.flags & F_KIND_MASK) == TRK_KIND_PRISMATIC; }
        }
    }
    __syncthreads();
    // per warp: two output tiles | cpw records (p, quat per body; masked axis per q column) | cpw x n_live pose slots
    const int row = 7 * D, per_cfg = L * row;
    const int tile_words = (link_chunk * row + 4 + 3) & ~3;                   // + 4: room for the alignment phase
    const int warp_words = 2 * tile_words + cpw * rec_stride + cpw * 12 * live.n_live;
    float* tile = stage + (size_t)warp * warp_words;
    float* wst = tile + 2 * tile_words;
    float* wlive = wst + cpw * rec_stride;
    int buf = 0;
    const float invD = 1.f / (float)D, invL = 1.f / (float)L;
    for (long long base = ((long long)blockIdx.x * nwarps + warp) * cpw; base < n; base += (long long)gridDim.x * nwarps * cpw) {
        const int ncfg = (int)((n - base) < cpw ? (n - base) : cpw);
        if (lane < ncfg) {   // phase 1: one lane per configuration
            float* rec = wst + lane * rec_stride;
            float* axes = rec + L * 7;
            fk_live(hdr, recs, live, q + (base + lane) * D, wlive + lane, cpw,
                    [&](int, const LinkRec& r, const float* w, const float* wp, bool pass) {
                        float qt[4];
                        quat_of_rotation(w, qt);
                        float* o = rec + r.body * 7;
                        o[0] = w[9]; o[1] = w[10]; o[2] = w[11]; o[3] = qt[0]; o[4] = qt[1]; o[5] = qt[2]; o[6] = qt[3];
                        if (r.qidx >= 0) {
                            const float mk = pass ? 1.f : 0.f;
                            float ax, ay, az;
                            if ((r.flags & F_KIND_MASK) == TRK_KIND_REVOLUTE) {
                                const int k = (r.flags >> F_AXIS_SHIFT) & 3;
                                ax = r.s * w[k]; ay = r.s * w[3 + k]; az = r.s * w[6 + k];
                            } else {
                                ax = wp[0] * r.ax[0] + wp[1] * r.ax[1] + wp[2] * r.ax[2];
                                ay = wp[3] * r.ax[0] + wp[4] * r.ax[1] + wp[5] * r.ax[2];
                                az = wp[6] * r.ax[0] + wp[7] * r.ax[1] + wp[8] * r.ax[2];
                            }
                            float* a = axes + r.qidx * 3;
                            a[0] = mk * ax; a[1] = mk * ay; a[2] = mk * az;
                        }
                    });
        }
        __syncwarp();
        // phase 2: the warp's output is one contiguous run of ncfg * L link blocks (row = 7 * D floats each); handle
        // `link_chunk` blocks per pass, whatever configuration they belong to
        const int n_inst = ncfg * L;
        const long long out0 = base * (long long)per_cfg;
        for (int u0 = 0; u0 < n_inst; u0 += link_chunk, buf ^= 1) {
            const int nu = min(link_chunk, n_inst - u0);
            const long long f0 = out0 + (long long)u0 * row;
            const int ph = (int)(f0 & 3);                        // same 16-byte phase in shared and global memory
            float* tl = tile + buf * tile_words + ph;
            // this buffer left with the bulk store issued two passes ago: it must have been read before it is refilled
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            // 2a: lane <- (link block, q column); its seven values go to the tile at their output offsets.  Branch-free
            // (non-ancestor columns are computed and zeroed) so that four columns per lane are in flight at once.
            const int ncol = nu * D;
#pragma unroll 4
            for (int t = lane; t < ncol; t += 32) {
                const int ll = __float2int_rz(((float)t + 0.5f) * invD);
                const int j = t - ll * D, u = u0 + ll;
                const int c = __float2int_rz(((float)u + 0.5f) * invL);
                const int l = u - c * L;
                const float* rec = wst + c * rec_stride;
                const float* a = rec + L * 7 + j * 3;
                const float* pl = rec + l * 7;
                const float* pj = rec + jbody[j] * 7;
                const bool on = ((anc[l] >> j) & 1u) != 0, pris = jprism[j] != 0;
                const bool rev = on && !pris, pr = on && pris;     // selects, not products: NaN poses must not leak into zero columns
                const float ax = a[0], ay = a[1], az = a[2];
                const float dx = pl[0] - pj[0], dy = pl[1] - pj[1], dz = pl[2] - pj[2];
                const float qw = pl[3], qx = pl[4], qy = pl[5], qz = pl[6];
                float* o = tl + ll * row + j;
                o[0]     = rev ? ay * dz - az * dy : (pr ? ax : 0.f);
                o[D]     = rev ? az * dx - ax * dz : (pr ? ay : 0.f);
                o[2 * D] = rev ? ax * dy - ay * dx : (pr ? az : 0.f);
                o[3 * D] = rev ? -0.5f * (ax * qx + ay * qy + az * qz) : 0.f;
                o[4 * D] = rev ? 0.5f * (qw * ax + (ay * qz - az * qy)) : 0.f;
                o[5 * D] = rev ? 0.5f * (qw * ay + (az * qx - ax * qz)) : 0.f;
                o[6 * D] = rev ? 0.5f * (qw * az + (ax * qy - ay * qx)) : 0.f;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            // 2b: the tile is a contiguous span of J: its 16-byte aligned body leaves with ONE bulk store
            // (cp.async.bulk.global.shared::cta) issued by lane 0, a scalar head and tail cover the misaligned ends.
            const int nf = nu * row;
            const int head = min((4 - ph) & 3, nf);
            const int nbody = (nf - head) >> 2;
            const int tail0 = head + 4 * nbody;
            if (lane < head) __stcs(J + f0 + lane, tl[lane]);
            if (lane == 0) {
                if (nbody > 0)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(J + f0 + head), "r"((unsigned)__cvta_generic_to_shared(tl + head)), "r"(nbody * 16) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (lane < nf - tail0) __stcs(J + f0 + tail0 + lane, tl[tail0 + lane]);
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// Fused batched Adam inverse kinematics for the runtime-tree family: the loop of k_ik_adam_static (trk_static.cuh;
// DifferentiableTree.inverse_kinematics, robot_tree.py:303-442) with q, the Adam moments, the joint axes / origins and
// the live poses in per-lane shared-memory columns instead of registers.  One thread owns one seed for all iterations.
// Per-warp words: (n_live * 12 + D * 6 + D * 3) * 32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_ik_adam_generic(const float* __restrict__ tree_words, int tree_nwords,
                                                            const __grid_constant__ TreePlan plan, const __grid_constant__ trk::IKParams P,
                                                            const float* __restrict__ step_size, const float* __restrict__ bc2_sqrt,
                                                            const float* __restrict__ target_seed, long long n_seeds,
                                                            float* __restrict__ q_io, float* __restrict__ err_out,
                                                            uint8_t* __restrict__ valid_out, int* __restrict__ valid_count,
                                                            float* __restrict__ loss_sum, float* __restrict__ loss_sqsum) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_cnt;
    __shared__ float s_sum, s_sq;
    float* tw = smem;
    block_copy_words(tw, tree_words, tree_nwords);
    __syncthreads();
    const TreeHeader* hdr = reinterpret_cast<const TreeHeader*>(tw);
    const LinkRec* recs = reinterpret_cast<const LinkRec*>(tw + 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, D = hdr->D;
    float* wbase = smem + tree_nwords + (size_t)warp * (plan.n_live * 12 + D * 9) * 32;
    float* wl = wbase + lane;                                  // live poses
    float* JA = wbase + plan.n_live * 12 * 32 + lane;          // [c * 6 + k]: axis (3), origin (3) of q column c
    float* QS = JA + D * 6 * 32;                               // [c * 3 + k]: q, exp_avg, exp_avg_sq
    const long long seed = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = seed < n_seeds;
    for (int c = 0; c < D; ++c) { QS[(c * 3) * 32] = live ? q_io[seed * D + c] : 0.f; QS[(c * 3 + 1) * 32] = 0.f; QS[(c * 3 + 2) * 32] = 0.f; }
    float Rt[9], pt[3];
    {
        const float* Ht = (P.per_seed_target && live) ? target_seed + seed * 16 : P.target;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            Rt[i * 3 + 0] = Ht[i * 4 + 0]; Rt[i * 3 + 1] = Ht[i * 4 + 1]; Rt[i * 3 + 2] = Ht[i * 4 + 2]; pt[i] = Ht[i * 4 + 3];
        }
    }
    EEParams ee;
    ee.ee_w_rot = 1.f; ee.ee_w_pos = 1.f; ee.ee_square = 0; ee.w[3] = P.w_se3;
    float err = 0.f;
    bool valid = false;
#pragma unroll 1
    for (int it = 0; it < P.n_iters; ++it) {
        if (threadIdx.x == 0) { s_cnt = 0; s_sum = 0.f; s_sq = 0.f; }
        __syncthreads();
        float Wee[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) Wee[k] = 0.f;
        unsigned pass_mask = 0u;
        fk_live_q(hdr, recs, plan, [&](int c) { return QS[(c * 3) * 32]; }, wl, 32,
                  [&](int, const LinkRec& r, const float* w, const float* wp, bool pass) {
                      if (r.qidx >= 0) {
                          float ax, ay, az;
                          if ((r.flags & F_KIND_MASK) == TRK_KIND_REVOLUTE) {
                              const int k = (r.flags >> F_AXIS_SHIFT) & 3;
                              ax = r.s * (k == 0 ? w[0] : (k == 1 ? w[1] : w[2]));
                              ay = r.s * (k == 0 ? w[3] : (k == 1 ? w[4] : w[5]));
                              az = r.s * (k == 0 ? w[6] : (k == 1 ? w[7] : w[8]));
                          } else {
                              ax = wp[0] * r.ax[0] + wp[1] * r.ax[1] + wp[2] * r.ax[2];
                              ay = wp[3] * r.ax[0] + wp[4] * r.ax[1] + wp[5] * r.ax[2];
                              az = wp[6] * r.ax[0] + wp[7] * r.ax[1] + wp[8] * r.ax[2];
                          }
                          float* j = JA + r.qidx * 6 * 32;
                          j[0] = ax; j[32] = ay; j[64] = az; j[96] = w[9]; j[128] = w[10]; j[160] = w[11];
                          if (pass) pass_mask |= 1u << r.qidx;
                      }
                      if (r.body == P.ee_link) {
#pragma unroll
                          for (int k = 0; k < 12; ++k) Wee[k] = w[k];
                      }
                  });
        float gR[9], gp[3];
        err = ee_term<true>(&ee, Wee, &Wee[9], Rt, pt, gR, gp);       // SE3_distance, un-squared, weights 1
        bool in_lim = true;
        float e_lim = 0.f;
        for (int c = 0; c < D; ++c) {
            const float lo = P.lower[c], hi = P.upper[c], qc = QS[(c * 3) * 32];
            in_lim = in_lim && (qc >= lo) && (qc <= hi);
            const float dl = lo - qc, dh = hi - qc;
            e_lim += (qc < lo ? dl * dl : 0.f) + (qc > hi ? dh * dh : 0.f);
        }
        valid = in_lim && (err < P.se3_eps);
        const float loss = P.w_se3 * err + P.w_lim * e_lim;
        const unsigned ball = __ballot_sync(0xffffffffu, valid && live);
        float ls = live ? loss : 0.f, lq = live ? loss * loss : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ls += __shfl_xor_sync(0xffffffffu, ls, o); lq += __shfl_xor_sync(0xffffffffu, lq, o); }
        if (lane == 0) { atomicAdd(&s_cnt, __popc(ball)); atomicAdd(&s_sum, ls); atomicAdd(&s_sq, lq); }
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(valid_count + it, s_cnt);
            if (loss_sum) { atomicAdd(loss_sum + it, s_sum); atomicAdd(loss_sqsum + it, s_sq); }
        }
        if (it >= P.n_updates) continue;
        // the only wrench is the EE one: d loss / d q_c = axis_c . (T - p_c x F) for the joints above the EE body
        float tx, ty, tz;
        axial_gR_Rt(gR, Wee, tx, ty, tz);
        const float fx = gp[0], fy = gp[1], fz = gp[2];
        tx += Wee[10] * fz - Wee[11] * fy; ty += Wee[11] * fx - Wee[9] * fz; tz += Wee[9] * fy - Wee[10] * fx;
        const float ss = step_size[it], inv_b2s = 1.f / bc2_sqrt[it];
        const unsigned gmask = P.anc_mask & pass_mask;
        for (int c = 0; c < D; ++c) {
            const float* j = JA + c * 6 * 32;
            const float ax = j[0], ay = j[32], az = j[64], px = j[96], py = j[128], pz = j[160];
            float g = ((plan.prism_mask >> c) & 1u)
                          ? ax * fx + ay * fy + az * fz
                          : ax * (tx - (py * fz - pz * fy)) + ay * (ty - (pz * fx - px * fz)) + az * (tz - (px * fy - py * fx));
            g = ((gmask >> c) & 1u) ? g : 0.f;
            const float lo = P.lower[c], hi = P.upper[c];
            float* s = QS + c * 3 * 32;
            const float qc = s[0];
            g += P.w_lim * ((qc < lo ? -2.f * (lo - qc) : 0.f) + (qc > hi ? -2.f * (hi - qc) : 0.f));
            const float mom = fmaf(1.f - P.beta1, g - s[32], s[32]);                 // exp_avg.lerp_(grad, 1 - beta1)
            const float var = fmaf((1.f - P.beta2) * g, g, s[64] * P.beta2);          // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
            float rs_unused;
            const float denom = sqrt_rn_core(var, rs_unused) * inv_b2s + P.eps;
            s[0] = qc - ss * __fdividef(mom, denom);
            s[32] = mom; s[64] = var;
        }
    }
    if (live) {
        for (int c = 0; c < D; ++c) q_io[seed * D + c] = QS[(c * 3) * 32];
        if (err_out) err_out[seed] = err;
        if (valid_out) valid_out[seed] = valid ? 1 : 0;
    }
}

// cost stage on caller-supplied positions (DistanceField.compute_cost, distance_fields.py:26-55)
template <int MAXL, bool WANT_GRAD>
__global__ void __launch_bounds__(kBlock) k_cost_stage(const float* __restrict__ cost_words, int cost_nwords, int L,
                                                       const float* __restrict__ link_pos, const float* __restrict__ ee_pose,
                                                       long long n, const float* __restrict__ ee_target_cfg,
                                                       float margin_override, int use_override,
                                                       float* __restrict__ cost_total, float* __restrict__ cost_terms,
                                                       float* __restrict__ grad_pos, float* __restrict__ grad_ee,
                                                       uint8_t* __restrict__ occupancy, int* __restrict__ argmin) {
    extern __shared__ __align__(16) float smem[];
    float* cw = smem;
    block_copy_words(cw, cost_words, cost_nwords);
    __syncthreads();
    const CostView cv(cw);
    const CostHeader* h = cv.h;
    for (long long cfg = (long long)blockIdx.x * kBlock + threadIdx.x; cfg < n; cfg += (long long)gridDim.x * kBlock) {
        const float* P = link_pos + cfg * (long long)L * 3;
        float F[MAXL][3];
        if (WANT_GRAD)
            for (int b = 0; b < L; ++b) F[b][0] = F[b][1] = F[b][2] = 0.f;
        float terms[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned occ;
        collision_terms<WANT_GRAD, true>(
            cv, cv.pt_link, cv.pair_i, cv.pair_j, [&](int b, float& x, float& y, float& z) { x = __ldg(P + b * 3); y = __ldg(P + b * 3 + 1); z = __ldg(P + b * 3 + 2); },
            [&](int b, float gx, float gy, float gz) { F[b][0] += gx; F[b][1] += gy; F[b][2] += gz; }, margin_override,
            use_override != 0, terms, occ, argmin ? argmin + cfg * (long long)(h->nObj * h->S) : nullptr);
        if (h->ee_enabled && ee_pose) {
            float R[9], p[3], Rt[9], pt[3], gR[9], gp[3];
            const float* He = ee_pose + cfg * 16;
            const float* Ht = ee_target_cfg ? ee_target_cfg + cfg * 16 : nullptr;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    R[i * 3 + j] = __ldg(He + i * 4 + j);
                    Rt[i * 3 + j] = Ht ? __ldg(Ht + i * 4 + j) : h->ee_R[i * 3 + j];
                }
                p[i] = __ldg(He + i * 4 + 3);
                pt[i] = Ht ? __ldg(Ht + i * 4 + 3) : h->ee_t[i];
            }
            terms[3] = ee_term<WANT_GRAD>(h, R, p, Rt, pt, gR, gp);
            if (WANT_GRAD && grad_ee) {
                float* Ge = grad_ee + cfg * 16;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    Ge[i * 4 + 0] = gR[i * 3 + 0]; Ge[i * 4 + 1] = gR[i * 3 + 1]; Ge[i * 4 + 2] = gR[i * 3 + 2];
                    Ge[i * 4 + 3] = gp[i];
                }
                Ge[12] = Ge[13] = Ge[14] = Ge[15] = 0.f;
            }
        }
        if (cost_total) cost_total[cfg] = weighted_total(h, terms);
        if (cost_terms) reinterpret_cast<float4*>(cost_terms)[cfg] = make_float4(terms[0], terms[1], terms[2], terms[3]);
        if (occupancy) occupancy[cfg] = (uint8_t)occ;
        if (WANT_GRAD && grad_pos) {
            float* Gp = grad_pos + cfg * (long long)L * 3;
            for (int b = 0; b < L; ++b) { Gp[b * 3] = F[b][0]; Gp[b * 3 + 1] = F[b][1]; Gp[b * 3 + 2] = F[b][2]; }
        }
    }
}

// self-test: the branch-free square root used on the decision path against sqrt.rn.f32
__global__ void k_selftest_sqrt(const float* __restrict__ x, long long n, float* __restrict__ y_core, float* __restrict__ y_ref) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float rs;
        y_core[i] = sqrt_rn_core(x[i], rs);
        y_ref[i] = __fsqrt_rn(x[i]);
    }
}

// ------------------------------------------------------------------------------------------
// host side: handles, validation, launches
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

struct trk_model {
    int device;
    int L, D;
    int maxl_bucket;            // 16 / 32 / 64
    int n_sms;
    float* d_words;             // device tree buffer
    int nwords;
    int family;                 // trk::StaticFamily id, 0 = generic
    int ext_family;             // > 0: a compiled-in robot plus attached fixed bodies (trk_ext.cuh serves its objective)
    std::vector<float> h_words; // host copy of the tree buffer (trk_cost_create folds attached bodies onto their anchors)
    int32_t parent[TRK_MAX_LINKS];
    int32_t qidx[TRK_MAX_LINKS];
    TreePlan live;              // live pose slots + joint tables of the runtime-tree kernels
};

// Parents-first walk: a body's pose slot is released once its last child has been composed (the child already holds
// the parent pose in registers, so it may take over the very slot), leaves need none.
static void build_tree_plan(const int32_t* topo, const int32_t* parent, const int32_t* qidx, const int32_t* kind, const float* R_fixed, int L,
                            TreePlan& ls) {
    memset(&ls, 0, sizeof(ls));
    int nj = 0;
    for (int i = 0; i < L; ++i) {   // joint tables (parents first)
        const int b = topo[i], p = parent[b];
        const int up = p >= 0 ? ls.jm[p] : -1;
        ls.jm[b] = (int8_t)(qidx[b] >= 0 ? qidx[b] : up);
        if (qidx[b] >= 0) {
            ls.pjoint[qidx[b]] = (int8_t)up;
            ls.jorder[nj++] = (int8_t)qidx[b];
            if (kind[b] == TRK_KIND_PRISMATIC) ls.prism_mask |= 1u << qidx[b];
        }
    }
    int pos_of[TRK_MAX_LINKS], last_child[TRK_MAX_LINKS], slot_of[TRK_MAX_LINKS];
    bool busy[TRK_MAX_LINKS] = {false};
    for (int i = 0; i < L; ++i) { pos_of[topo[i]] = i; last_child[i] = -1; }
    for (int i = 0; i < L; ++i) { const int p = parent[topo[i]]; if (p >= 0) last_child[pos_of[p]] = i; }
    ls.n_live = 0;
    for (int i = 0; i < L; ++i) {
        const int p = parent[topo[i]];
        ls.par[i] = (int8_t)(p >= 0 ? slot_of[pos_of[p]] : -1);
        if (p >= 0 && last_child[pos_of[p]] == i) busy[slot_of[pos_of[p]]] = false;
        ls.own[i] = -1; slot_of[i] = -1;
        {
            const int b = topo[i];
            const float* Rf = R_fixed + 9 * b;
            const bool ident = Rf[0] == 1.f && Rf[4] == 1.f && Rf[8] == 1.f && Rf[1] == 0.f && Rf[2] == 0.f && Rf[3] == 0.f && Rf[5] == 0.f &&
                               Rf[6] == 0.f && Rf[7] == 0.f;
            if (p >= 0 && last_child[i] < 0 && kind[b] == TRK_KIND_FIXED && ident) ls.point_mask |= 1ull << i;
        }
        if (last_child[i] >= 0) {
            int s = 0;
            while (busy[s]) ++s;
            busy[s] = true; slot_of[i] = s; ls.own[i] = (int8_t)s;
            if (s + 1 > ls.n_live) ls.n_live = s + 1;
        }
    }
}

struct trk_cost {
    int device;
    int L;
    int S, nObj, K, P;
    int ee_enabled;
    int n_slots;                // bodies referenced by the tables
    int n_pslots;               // bodies referenced by self-collision pairs
    int cm_match;               // cost table equals the compile-time collision model of the model's family
    float* d_words;
    int nwords;
    std::vector<float> h_words;  // host mirror, for the setters
    // attached-body models (trk_ext.cuh): every referenced body as (movable anchor, offset); built when the model is one
    float* d_ext = nullptr;
    int ext_nwords = 0;
    std::vector<float> h_ext;    // host copy (the launcher reads the header)
    bool ext_ok = false;
    // host-buffer pipeline (trk_objective_host): staging buffers, streams and events, created on first use
    float* hp_q = nullptr; float* hp_cost = nullptr; float* hp_grad = nullptr;
    int64_t hp_cap = 0;
    cudaStream_t hp_s[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t hp_in[64], hp_k[64];
    int hp_events = 0;
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

static int grid_for(const trk_model* m, long long n, int blocks_per_sm) {
    long long need = (n + kBlock - 1) / kBlock;
    long long cap = (long long)m->n_sms * blocks_per_sm;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// Block size and dynamic shared memory of a runtime-tree kernel: per-warp words as generic_warp_words() lays them out;
// four warps per block unless that would leave fewer than two blocks per SM.
static int align4(int x) { return (x + 3) & ~3; }
struct GenericLaunch { int warps, grid; size_t smem; };
static GenericLaunch generic_launch(const trk_model* m, long long n, size_t fixed_words, int stage_words, int n_joint_cols, int n_pslots) {
    const size_t per_warp = (size_t)generic_warp_words(stage_words, m->live.n_live, n_joint_cols, n_pslots) * 4;
    GenericLaunch g;
    g.warps = kWarps;
    while (g.warps > 1 && fixed_words * 4 + g.warps * per_warp > 100 * 1024) --g.warps;
    g.smem = fixed_words * 4 + g.warps * per_warp;
    const long long need = (n + g.warps * 32 - 1) / (g.warps * 32), cap = (long long)m->n_sms * 16;
    g.grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
    return g;
}

static int build_slots(const trk_model* m, const int32_t* link_sel, int32_t n_sel, SlotMap& sm, int& n_out) {
    for (int i = 0; i < TRK_MAX_LINKS; ++i) sm.slot[i] = -1;
    if (!link_sel) {
        for (int i = 0; i < m->L; ++i) sm.slot[i] = (int8_t)i;
        n_out = m->L;
        return TRK_OK;
    }
    if (n_sel <= 0 || n_sel > m->L) return fail(TRK_E_RANGE, "n_sel=%d outside [1, %d]", n_sel, m->L);
    for (int s = 0; s < n_sel; ++s) {
        const int b = link_sel[s];
        if (b < 0 || b >= m->L) return fail(TRK_E_RANGE, "link_sel[%d]=%d outside the model", s, b);
        if (sm.slot[b] >= 0) return fail(TRK_E_UNSUPPORTED, "link_sel lists body %d twice", b);
        sm.slot[b] = (int8_t)s;
    }
    n_out = n_sel;
    return TRK_OK;
}

#define TRK_CHECK_PTR(p, name) \
    if (!(p)) return fail(TRK_E_NULL, "%s is null", name)
#define TRK_CHECK_ALIGN16(p, name) \
    if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) return fail(TRK_E_ALIGN, "%s must be 16-byte aligned", name)

template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Attached-body model: every body the cost tables reference as (anchor, offset).  Walking up from the body through FIXED
// joints composes its constant offset in the frame of the first movable body met (anchor 1 + q column) or of the root
// (anchor 0): pose_body = pose_parent o (Rf, t)  =>  offset_in_parent = Rf offset + t.
static int build_ext_tables(const trk_model* m, const trk_cost_desc* d, trk_cost* c) {
    const LinkRec* recs = reinterpret_cast<const LinkRec*>(m->h_words.data() + 16);
    int rec_of[TRK_MAX_LINKS];
    for (int i = 0; i < m->L; ++i) rec_of[recs[i].body] = i;
    auto fold = [&](int body, int& anchor, int& col, float t[3]) {
        t[0] = t[1] = t[2] = 0.f;
        int cur = body;
        while (true) {
            const LinkRec& r = recs[rec_of[cur]];
            if (r.qidx >= 0) { anchor = 1 + r.qidx; col = r.qidx; return; }
            if (r.parent < 0) { anchor = 0; col = -1; return; }
            const float x = t[0], y = t[1], z = t[2];
            t[0] = r.Rf[0] * x + r.Rf[1] * y + r.Rf[2] * z + r.t[0];
            t[1] = r.Rf[3] * x + r.Rf[4] * y + r.Rf[5] * z + r.t[1];
            t[2] = r.Rf[6] * x + r.Rf[7] * y + r.Rf[8] * z + r.t[2];
            cur = r.parent;
        }
    };
    const int S = d->n_points, P = d->n_pairs, NA = m->D + 1;
    struct Pt { int anchor, s; float t[3]; };
    std::vector<Pt> pts(S);
    for (int s = 0; s < S; ++s) { int col; pts[s].s = s; fold(d->pt_link[s], pts[s].anchor, col, pts[s].t); }
    std::stable_sort(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) { return a.anchor < b.anchor; });
    // pair-position slots: one per distinct pair body, grouped by anchor
    struct Ps { int anchor, col, body; float t[3]; };
    std::vector<Ps> ps;
    std::vector<int> slot_of_body(m->L, -1);
    for (int p = 0; p < P; ++p)
        for (int b : {d->pair_i[p], d->pair_j[p]})
            if (slot_of_body[b] < 0) { Ps e; e.body = b; fold(b, e.anchor, e.col, e.t); slot_of_body[b] = (int)ps.size(); ps.push_back(e); }
    std::vector<int> order(ps.size());
    for (size_t i = 0; i < ps.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ps[a].anchor < ps[b].anchor; });
    std::vector<int> new_slot(ps.size());
    for (size_t i = 0; i < order.size(); ++i) new_slot[order[i]] = (int)i;
    const int n_ps = (int)ps.size();
    // runs: pairs that share a hub body and whose other bodies hang on one anchor.  The hub of a pair is its end with the larger
    // number of pairs (the link a grasped object's points are all checked against), ties: the second body.
    struct Item { int hub, other, col_o, p; };
    std::vector<Item> its(P);
    {
        std::vector<int> degree(m->L, 0);
        for (int p = 0; p < P; ++p) { ++degree[d->pair_i[p]]; ++degree[d->pair_j[p]]; }
        for (int p = 0; p < P; ++p) {
            const int bi = d->pair_i[p], bj = d->pair_j[p];
            const bool hub_is_i = degree[bi] > degree[bj];
            its[p].hub = new_slot[slot_of_body[hub_is_i ? bi : bj]];
            its[p].other = new_slot[slot_of_body[hub_is_i ? bj : bi]];
            its[p].col_o = ps[order[its[p].other]].col;
            its[p].p = p;
        }
        std::stable_sort(its.begin(), its.end(), [](const Item& a, const Item& b) { return a.hub != b.hub ? a.hub < b.hub : a.col_o < b.col_o; });
    }
    int n_runs = 0;
    for (int k = 0; k < P; ++k) if (k == 0 || its[k].hub != its[k - 1].hub || its[k].col_o != its[k - 1].col_o) ++n_runs;
    int off = (int)(sizeof(trk::ExtHeader) / 4);
    auto take = [&](int nw) { int o = off; off = align4(off + nw); return o; };
    const int o_pts = take(NA + 1), o_pt = take(8 * S), o_pss = take(NA + 1), o_ps = take(4 * n_ps), o_run = take(4 * n_runs),
              o_item = take(4 * P);
    std::vector<float> w(off, 0.f);
    int* iw = reinterpret_cast<int*>(w.data());
    trk::ExtHeader* h = reinterpret_cast<trk::ExtHeader*>(w.data());
    h->n_pts = S; h->n_ps = n_ps; h->n_pairs = P; h->n_anchor = NA;
    h->off_pt_start = o_pts; h->off_pt = o_pt; h->off_ps_start = o_pss; h->off_ps = o_ps; h->off_run = o_run; h->off_item = o_item;
    h->n_runs = n_runs;
    h->total_words = off;
    {
        int k = 0;
        for (int a = 0; a <= NA; ++a) {
            while (k < S && pts[k].anchor < a) ++k;
            iw[o_pts + a] = k;
        }
    }
    for (int k2 = 0; k2 < S; ++k2) {
        float* r = w.data() + o_pt + 8 * k2;
        r[0] = pts[k2].t[0]; r[1] = pts[k2].t[1]; r[2] = pts[k2].t[2];
        r[3] = d->pt_margin_obj[pts[k2].s]; r[4] = d->pt_margin_ws[pts[k2].s];       // [5], [6]: occupancy margins, written by the kernel
        reinterpret_cast<int*>(r)[7] = pts[k2].s;
    }
    {
        int k = 0;
        for (int a = 0; a <= NA; ++a) {
            while (k < n_ps && ps[order[k]].anchor < a) ++k;
            iw[o_pss + a] = k;
        }
        for (int k2 = 0; k2 < n_ps; ++k2) {
            const Ps& e = ps[order[k2]];
            float* r = w.data() + o_ps + 4 * k2;
            r[0] = e.t[0]; r[1] = e.t[1]; r[2] = e.t[2];
        }
        int run = -1;
        for (int k2 = 0; k2 < P; ++k2) {
            if (k2 == 0 || its[k2].hub != its[k2 - 1].hub || its[k2].col_o != its[k2 - 1].col_o) {
                ++run;
                int* rr = iw + o_run + 4 * run;
                rr[0] = its[k2].hub; rr[1] = k2;
                rr[3] = (ps[order[its[k2].hub]].col & 0xff) | ((its[k2].col_o & 0xff) << 8);
                if (k2 == 0 || its[k2].hub != its[k2 - 1].hub) rr[3] |= 0x10000;
            }
            if (k2 + 1 == P || its[k2 + 1].hub != its[k2].hub) iw[o_run + 4 * run + 3] |= 0x20000;
            iw[o_run + 4 * run + 2] = k2 + 1;
            float* r = w.data() + o_item + 4 * k2;
            reinterpret_cast<int*>(r)[0] = its[k2].other;
            r[1] = d->pair_margin[its[k2].p];
        }
    }
    DeviceGuard guard(m->device);
    cudaError_t e = cudaMalloc(&c->d_ext, w.size() * sizeof(float));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(ext tables)");
    e = cudaMemcpy(c->d_ext, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(ext tables)");
    c->ext_nwords = (int)w.size(); c->h_ext = w; c->ext_ok = true;
    return TRK_OK;
}

extern "C" {

int trk_abi_version(void) { return TRK_ABI_VERSION; }
const char* trk_last_error(void) { return g_err; }

int trk_model_create(const trk_model_desc* d, int device, trk_model** out) {
    TRK_CHECK_PTR(d, "desc");
    TRK_CHECK_PTR(out, "out");
    *out = nullptr;
    if (d->n_links < 1 || d->n_links > TRK_MAX_LINKS) return fail(TRK_E_RANGE, "n_links=%d outside [1, %d]", d->n_links, TRK_MAX_LINKS);
    if (d->n_dofs < 0 || d->n_dofs > TRK_MAX_DOFS) return fail(TRK_E_RANGE, "n_dofs=%d outside [0, %d]", d->n_dofs, TRK_MAX_DOFS);
    TRK_CHECK_PTR(d->parent, "parent"); TRK_CHECK_PTR(d->kind, "kind"); TRK_CHECK_PTR(d->axis_k, "axis_k");
    TRK_CHECK_PTR(d->axis_s, "axis_s"); TRK_CHECK_PTR(d->axis, "axis"); TRK_CHECK_PTR(d->q_index, "q_index");
    TRK_CHECK_PTR(d->R_fixed, "R_fixed"); TRK_CHECK_PTR(d->trans, "trans"); TRK_CHECK_PTR(d->lo, "lo");
    TRK_CHECK_PTR(d->hi, "hi"); TRK_CHECK_PTR(d->has_lim, "has_lim"); TRK_CHECK_PTR(d->topo, "topo");
    const int L = d->n_links, D = d->n_dofs;
    // validate topology: topo is a permutation, parents first, exactly one root, q columns used once
    bool seen[TRK_MAX_LINKS] = {false};
    bool qseen[TRK_MAX_DOFS] = {false};
    int roots = 0;
    for (int i = 0; i < L; ++i) {
        const int b = d->topo[i];
        if (b < 0 || b >= L || seen[b]) return fail(TRK_E_RANGE, "topo is not a permutation of the bodies");
        const int p = d->parent[b];
        if (p < 0) {
            ++roots;
        } else if (p >= L || !seen[p]) {
            return fail(TRK_E_RANGE, "topo lists body %d before its parent %d", b, p);
        }
        seen[b] = true;
        const int k = d->kind[b];
        if (k < 0 || k > 2) return fail(TRK_E_RANGE, "kind[%d]=%d", b, k);
        const int qi = d->q_index[b];
        if (k == TRK_KIND_FIXED) {
            if (qi >= 0) return fail(TRK_E_RANGE, "fixed body %d has a q column", b);
        } else {
            if (p < 0) return fail(TRK_E_RANGE, "the root cannot be actuated");
            if (qi < 0 || qi >= D || qseen[qi]) return fail(TRK_E_RANGE, "q_index[%d]=%d invalid or duplicated", b, qi);
            qseen[qi] = true;
            if (d->axis_k[b] < 0 || d->axis_k[b] > 2) return fail(TRK_E_RANGE, "axis_k[%d]", b);
        }
    }
    if (roots != 1) return fail(TRK_E_RANGE, "expected exactly one root, found %d", roots);
    for (int c = 0; c < D; ++c)
        if (!qseen[c]) return fail(TRK_E_RANGE, "q column %d drives no body", c);

    const int nwords = 16 + 24 * L;
    std::vector<float> words(nwords, 0.f);
    TreeHeader* hdr = reinterpret_cast<TreeHeader*>(words.data());
    hdr->L = L; hdr->D = D; hdr->pad0 = hdr->pad1 = 0;
    const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    memcpy(hdr->base, d->base_pose ? d->base_pose : ident, sizeof(ident));
    LinkRec* recs = reinterpret_cast<LinkRec*>(words.data() + 16);
    for (int i = 0; i < L; ++i) {
        const int b = d->topo[i];
        LinkRec& r = recs[i];
        memcpy(r.Rf, d->R_fixed + 9 * b, 9 * sizeof(float));
        memcpy(r.t, d->trans + 3 * b, 3 * sizeof(float));
        memcpy(r.ax, d->axis + 3 * b, 3 * sizeof(float));
        r.s = d->axis_s[b];
        r.lo = d->lo[b]; r.hi = d->hi[b];
        r.flags = d->kind[b] | (d->axis_k[b] << F_AXIS_SHIFT) | (d->has_lim[b] ? F_HAS_LIM : 0);
        r.parent = d->parent[b];
        r.body = b;
        r.qidx = d->q_index[b];
        r.pad0 = r.pad1 = 0;
    }
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    trk_model* m = new (std::nothrow) trk_model();
    if (!m) return fail(TRK_E_RANGE, "out of host memory");
    m->device = device; m->L = L; m->D = D; m->nwords = nwords;
    m->maxl_bucket = L <= 16 ? 16 : (L <= 32 ? 32 : 64);
    memcpy(m->parent, d->parent, L * sizeof(int32_t));
    memcpy(m->qidx, d->q_index, L * sizeof(int32_t));
    build_tree_plan(d->topo, d->parent, d->q_index, d->kind, d->R_fixed, L, m->live);
    cudaError_t e = cudaDeviceGetAttribute(&m->n_sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaDeviceGetAttribute"); }
    e = cudaMalloc(&m->d_words, nwords * sizeof(float));
    if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaMalloc(tree)"); }
    e = cudaMemcpy(m->d_words, words.data(), nwords * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(m->d_words); delete m; return cuda_fail(e, "cudaMemcpy(tree)"); }
    m->family = trk::match_static_family(words.data(), nwords);
    m->ext_family = m->family ? 0 : trk::match_static_family_prefix(words.data(), nwords);
    m->h_words = words;
    if (getenv("TRK_B200_FORCE_GENERIC")) { m->family = 0; m->ext_family = 0; }   // diagnostic switch: runtime-tree kernels only
    *out = m;
    return TRK_OK;
}

void trk_model_destroy(trk_model* m) {
    if (!m) return;
    DeviceGuard guard(m->device);
    cudaFree(m->d_words);
    delete m;
}

const char* trk_model_kernel_family(const trk_model* m) { return m ? trk::static_family_name(m->family) : "null"; }

int trk_fk_forward(const trk_model* m, const float* q, int64_t n, const int32_t* link_sel, int32_t n_sel, float* H,
                   void* stream) {
    TRK_CHECK_PTR(m, "model");
    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(q, "q"); TRK_CHECK_PTR(H, "H_out");
    TRK_CHECK_ALIGN16(H, "H_out");
    SlotMap sm; int n_out;
    int rc = build_slots(m, link_sel, n_sel, sm, n_out);
    if (rc) return rc;
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->family) {
        cudaError_t e = trk::launch_static_fk_forward(m->family, m->n_sms, q, n, sm, n_out, H, st);
        return e == cudaSuccess ? TRK_OK : cuda_fail(e, "fk_forward(static)");
    }
    const GenericLaunch gl = generic_launch(m, n, m->nwords, 512, 0, 0);
    cudaError_t e = allow_smem(k_fk_forward_generic, gl.smem);
    if (e != cudaSuccess) return cuda_fail(e, "fk_forward smem");
    k_fk_forward_generic<<<gl.grid, gl.warps * 32, gl.smem, st>>>(m->d_words, m->nwords, m->live, q, n, sm, n_out, H);
    e = cudaGetLastError();
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "fk_forward launch");
}

int trk_fk_backward(const trk_model* m, const float* q, int64_t n, const int32_t* link_sel, int32_t n_sel, const float* G,
                    float* grad_q, void* stream) {
    TRK_CHECK_PTR(m, "model");
    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(q, "q"); TRK_CHECK_PTR(G, "G"); TRK_CHECK_PTR(grad_q, "grad_q");
    TRK_CHECK_ALIGN16(G, "G");
    SlotMap sm; int n_out;
    int rc = build_slots(m, link_sel, n_sel, sm, n_out);
    if (rc) return rc;
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->family) {
        cudaError_t e = trk::launch_static_fk_backward(m->family, m->n_sms, q, n, sm, n_out, G, grad_q, st);
        return e == cudaSuccess ? TRK_OK : cuda_fail(e, "fk_backward(static)");
    }
    if (m->D == 0) return TRK_OK;     // no movable joint: grad_q has no columns
    const GenericLaunch gl = generic_launch(m, n, m->nwords, 1024, m->D, 0);
    cudaError_t e = allow_smem(k_fk_backward_generic, gl.smem);
    if (e != cudaSuccess) return cuda_fail(e, "fk_backward smem");
    k_fk_backward_generic<<<gl.grid, gl.warps * 32, gl.smem, st>>>(m->d_words, m->nwords, m->live, q, n, sm, n_out, G, grad_q);
    e = cudaGetLastError();
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "fk_backward launch");
}


int trk_cost_create(const trk_model* m, const trk_cost_desc* d, trk_cost** out) {
    TRK_CHECK_PTR(m, "model"); TRK_CHECK_PTR(d, "desc"); TRK_CHECK_PTR(out, "out");
    *out = nullptr;
    const int S = d->n_points, nObj = d->n_objects, K = d->n_spheres, P = d->n_pairs, KB = d->n_boxes;
    if (S < 0 || S > TRK_MAX_POINTS) return fail(TRK_E_RANGE, "n_points=%d", S);
    if (nObj < 0 || nObj > TRK_MAX_OBJECTS) return fail(TRK_E_RANGE, "n_objects=%d", nObj);
    if (K < 0 || K > TRK_MAX_SPHERES) return fail(TRK_E_RANGE, "n_spheres=%d", K);
    if (P < 0 || P > TRK_MAX_PAIRS) return fail(TRK_E_RANGE, "n_pairs=%d", P);
    if (KB < 0 || KB > TRK_MAX_BOXES) return fail(TRK_E_RANGE, "n_boxes=%d", KB);
    if (KB) { TRK_CHECK_PTR(d->obj_box_start, "obj_box_start"); TRK_CHECK_PTR(d->box_center, "box_center"); TRK_CHECK_PTR(d->box_half, "box_half");
              TRK_CHECK_PTR(d->box_radius, "box_radius"); TRK_CHECK_PTR(d->box_rounded, "box_rounded"); }
    if (S) { TRK_CHECK_PTR(d->pt_link, "pt_link"); TRK_CHECK_PTR(d->pt_margin_obj, "pt_margin_obj"); TRK_CHECK_PTR(d->pt_margin_ws, "pt_margin_ws"); }
    if (nObj) { TRK_CHECK_PTR(d->obj_start, "obj_start"); TRK_CHECK_PTR(d->obj_identity, "obj_identity"); TRK_CHECK_PTR(d->obj_R, "obj_R"); TRK_CHECK_PTR(d->obj_pos, "obj_pos"); }
    if (K) { TRK_CHECK_PTR(d->sph_center, "sph_center"); TRK_CHECK_PTR(d->sph_radius, "sph_radius"); }
    if (P) { TRK_CHECK_PTR(d->pair_i, "pair_i"); TRK_CHECK_PTR(d->pair_j, "pair_j"); TRK_CHECK_PTR(d->pair_margin, "pair_margin"); }
    for (int s = 0; s < S; ++s)
        if (d->pt_link[s] < 0 || d->pt_link[s] >= m->L) return fail(TRK_E_RANGE, "pt_link[%d]=%d", s, d->pt_link[s]);
    for (int p = 0; p < P; ++p)
        if (d->pair_i[p] < 0 || d->pair_i[p] >= m->L || d->pair_j[p] < 0 || d->pair_j[p] >= m->L)
            return fail(TRK_E_RANGE, "self pair %d references a body outside the model", p);
    if (nObj) {
        if (d->obj_start[0] != 0 || d->obj_start[nObj] != K) return fail(TRK_E_RANGE, "obj_start must span [0, n_spheres]");
        for (int o = 0; o < nObj; ++o)
            if (d->obj_start[o + 1] < d->obj_start[o]) return fail(TRK_E_RANGE, "obj_start must be non-decreasing");
        if (KB && (d->obj_box_start[0] != 0 || d->obj_box_start[nObj] != KB)) return fail(TRK_E_RANGE, "obj_box_start must span [0, n_boxes]");
        for (int o = 0; o < nObj; ++o) {
            const int nb = KB ? d->obj_box_start[o + 1] - d->obj_box_start[o] : 0;
            if (nb < 0) return fail(TRK_E_RANGE, "obj_box_start must be non-decreasing");
            const int nprim = d->obj_start[o + 1] - d->obj_start[o] + nb;
            if (o == d->grid_obj ? nprim != 0 : nprim < 1)
                return fail(TRK_E_RANGE, o == d->grid_obj ? "object %d is the voxel SDF and must own no primitives" : "object %d has no primitive", o);
        }
    }
    if (d->ee_enabled && (d->ee_link < 0 || d->ee_link >= m->L)) return fail(TRK_E_RANGE, "ee_link=%d", d->ee_link);
    const int grid_obj = nObj ? d->grid_obj : -1;
    if (grid_obj >= nObj) return fail(TRK_E_RANGE, "grid_obj=%d", grid_obj);
    if (grid_obj >= 0) {
        TRK_CHECK_PTR(d->grid_data, "grid_data");
        if (reinterpret_cast<uintptr_t>(d->grid_data) % 16) return fail(TRK_E_ALIGN, "grid_data must be 16-byte aligned");
        if (!d->obj_identity[grid_obj]) return fail(TRK_E_RANGE, "the voxel-SDF object must have the identity pose");
        long long cells = 1;
        for (int i = 0; i < 3; ++i) {
            if (d->grid_n[i] < 1 || !(d->grid_extent[i] > 0.f)) return fail(TRK_E_RANGE, "grid axis %d: n=%d extent=%g", i, d->grid_n[i], d->grid_extent[i]);
            cells *= d->grid_n[i];
        }
        if (cells > 0x7fffffffLL) return fail(TRK_E_RANGE, "voxel SDF has %lld cells (limit 2^31 - 1)", cells);
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, d->grid_data) != cudaSuccess || attr.type != cudaMemoryTypeDevice || attr.device != m->device) {
            cudaGetLastError();
            return fail(TRK_E_RANGE, "grid_data must be device memory on the model's device");
        }
    }

    int KQ = 0;   // spheres padded per object to a multiple of two (scan table)
    for (int o = 0; o < nObj; ++o) KQ += (d->obj_start[o + 1] - d->obj_start[o] + 1) & ~1;
    int off = (int)(sizeof(CostHeader) / 4);
    auto take = [&](int nw) { int o = off; off = align4(off + nw); return o; };
    const int o_pt_link = take(S), o_mobj = take(S), o_mws = take(S), o_ostart = take(nObj + 1), o_oid = take(nObj),
              o_oR = take(nObj * 9), o_opos = take(nObj * 3), o_sph = take(K * 4), o_pi = take(P), o_pj = take(P), o_pm = take(P), o_ormax = take(nObj), o_oc2 = take(nObj), o_sphq = take(KQ * 4), o_oqs = take(nObj + 1), o_otol = take(nObj),
              o_sob = take(m->L), o_pts = take(S), o_psi = take(P), o_psj = take(P),
              o_obs = take(nObj + 1), o_box = take(KB * 8),
              o_psb = take(m->L), o_ppi = take(P), o_ppj = take(P), o_bps = take(m->L + 1), o_bpl = take(S);
    const int nwords = off;
    trk_cost* c = new (std::nothrow) trk_cost();
    if (!c) return fail(TRK_E_RANGE, "out of host memory");
    c->h_words.assign(nwords, 0.f);
    float* w = c->h_words.data();
    int* iw = reinterpret_cast<int*>(w);
    CostHeader* h = reinterpret_cast<CostHeader*>(w);
    h->S = S; h->nObj = nObj; h->K = K; h->P = P;
    h->ws_enabled = d->ws_enabled ? 1 : 0; h->ee_enabled = d->ee_enabled ? 1 : 0; h->ee_link = d->ee_link; h->ee_square = d->ee_square ? 1 : 0;
    h->clamp_self = d->clamp_self ? 1 : 0; h->clamp_obj = d->clamp_obj ? 1 : 0; h->clamp_ws = d->clamp_ws ? 1 : 0; h->total_words = nwords;
    for (int i = 0; i < 4; ++i) h->w[i] = d->weights[i];
    h->ee_w_pos = d->ee_w_pos; h->ee_w_rot = d->ee_w_rot;
    for (int i = 0; i < 3; ++i) { h->ws_min[i] = d->ws_min[i]; h->ws_max[i] = d->ws_max[i]; }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) h->ee_R[i * 3 + j] = d->ee_target[i * 4 + j];
        h->ee_t[i] = d->ee_target[i * 4 + 3];
    }
    h->off_pt_link = o_pt_link; h->off_pt_mobj = o_mobj; h->off_pt_mws = o_mws; h->off_obj_start = o_ostart;
    h->off_obj_ident = o_oid; h->off_obj_R = o_oR; h->off_obj_pos = o_opos; h->off_sph = o_sph;
    h->off_pair_i = o_pi; h->off_pair_j = o_pj; h->off_pair_m = o_pm; h->off_obj_rmax = o_ormax; h->off_obj_c2max = o_oc2; h->off_sphq = o_sphq; h->off_obj_qstart = o_oqs; h->off_obj_tol = o_otol;
    for (int s = 0; s < S; ++s) { iw[o_pt_link + s] = d->pt_link[s]; w[o_mobj + s] = d->pt_margin_obj[s]; w[o_mws + s] = d->pt_margin_ws[s]; }
    if (nObj) {
        for (int o = 0; o <= nObj; ++o) iw[o_ostart + o] = d->obj_start[o];
        int qpos = 0;
        for (int o = 0; o < nObj; ++o) {
            // per-object scan parameters: equal radii -> squared-distance keys; bits = index width packed in the key
            const int k0 = d->obj_start[o], k1 = d->obj_start[o + 1];
            bool uniform = true;
            float rmax = 0.f, c2max = 0.f;
            for (int k = k0; k < k1; ++k) {
                if (d->sph_radius[k] != d->sph_radius[k0]) uniform = false;
                rmax = fmaxf(rmax, fabsf(d->sph_radius[k]));
                const float* cc = d->sph_center + 3 * k;
                const float c2 = cc[0] * cc[0] + cc[1] * cc[1] + cc[2] * cc[2];
                c2max = fmaxf(c2max, c2);
                float* qw = w + o_sphq + 4 * (qpos + (k - k0));
                qw[0] = -2.f * cc[0]; qw[1] = -2.f * cc[1]; qw[2] = -2.f * cc[2]; qw[3] = c2;
            }
            iw[o_oqs + o] = qpos;
            const int padded = (k1 - k0 + 1) & ~1;
            for (int k = k1 - k0; k < padded; ++k) {   // dummies that can never be nearest
                float* qw = w + o_sphq + 4 * (qpos + k);
                qw[0] = qw[1] = qw[2] = 0.f; qw[3] = 1e30f;
            }
            qpos += padded;
            iw[o_oqs + o + 1] = qpos;
            w[o_oc2 + o] = c2max;
            int bits = 1;
            while ((1 << bits) < (k1 - k0)) ++bits;
            if (k1 == k0) uniform = false;
            iw[o_oid + o] = (d->obj_identity[o] ? OBJ_IDENTITY : 0) | (uniform ? OBJ_UNIFORM_RADIUS : 0) | (o == grid_obj ? OBJ_GRID : 0) |
                            (bits << OBJ_BITS_SHIFT);
            w[o_ormax + o] = rmax;
            w[o_otol + o] = ldexpf(1.f, bits - 19);
        }
        memcpy(w + o_oR, d->obj_R, nObj * 9 * sizeof(float));
        memcpy(w + o_opos, d->obj_pos, nObj * 3 * sizeof(float));
    } else {
        iw[o_ostart] = 0;
    }
    for (int k = 0; k < K; ++k) {
        w[o_sph + 4 * k + 0] = d->sph_center[3 * k + 0]; w[o_sph + 4 * k + 1] = d->sph_center[3 * k + 1];
        w[o_sph + 4 * k + 2] = d->sph_center[3 * k + 2]; w[o_sph + 4 * k + 3] = d->sph_radius[k];
    }
    for (int p = 0; p < P; ++p) { iw[o_pi + p] = d->pair_i[p]; iw[o_pj + p] = d->pair_j[p]; w[o_pm + p] = d->pair_margin[p]; }
    {   // dense slots for the bodies the tables reference (per-thread shared-memory columns exist only for these)
        int n_slots = 0;
        for (int b = 0; b < m->L; ++b) iw[o_sob + b] = -1;
        auto slot = [&](int b) { if (iw[o_sob + b] < 0) iw[o_sob + b] = n_slots++; return iw[o_sob + b]; };
        for (int p = 0; p < P; ++p) { iw[o_psi + p] = slot(d->pair_i[p]); iw[o_psj + p] = slot(d->pair_j[p]); }
        for (int s = 0; s < S; ++s) iw[o_pts + s] = slot(d->pt_link[s]);
        h->KB = KB; h->off_obj_box_start = o_obs; h->off_box = o_box;
        for (int o = 0; o <= nObj; ++o) iw[o_obs + o] = KB ? d->obj_box_start[o] : 0;
        for (int kb = 0; kb < KB; ++kb) {
            float* bw = w + o_box + 8 * kb;
            bw[0] = d->box_center[3 * kb]; bw[1] = d->box_center[3 * kb + 1]; bw[2] = d->box_center[3 * kb + 2]; bw[3] = d->box_radius[kb];
            bw[4] = d->box_half[3 * kb]; bw[5] = d->box_half[3 * kb + 1]; bw[6] = d->box_half[3 * kb + 2]; bw[7] = d->box_rounded[kb] ? 1.f : 0.f;
        }
        h->grid_obj = grid_obj;
        if (grid_obj >= 0) {
            for (int i = 0; i < 3; ++i) {
                h->grid_n[i] = d->grid_n[i]; h->grid_lo[i] = d->grid_lo[i]; h->grid_extent[i] = d->grid_extent[i];
                h->grid_nf[i] = (float)d->grid_n[i];
            }
            h->grid_data = reinterpret_cast<const float4*>(d->grid_data);
        }
        {   // pair-position slots and the body -> points lists of the runtime-tree objective kernel
            int n_ps = 0;
            for (int b = 0; b < m->L; ++b) iw[o_psb + b] = -1;
            auto pslot = [&](int b) { if (iw[o_psb + b] < 0) iw[o_psb + b] = n_ps++; return iw[o_psb + b]; };
            for (int p = 0; p < P; ++p) { iw[o_ppi + p] = pslot(d->pair_i[p]); iw[o_ppj + p] = pslot(d->pair_j[p]); }
            int pos = 0;
            for (int b = 0; b < m->L; ++b) {
                iw[o_bps + b] = pos;
                for (int s2 = 0; s2 < S; ++s2) if (d->pt_link[s2] == b) iw[o_bpl + pos++] = s2;
            }
            iw[o_bps + m->L] = pos;
            h->n_pslots = n_ps; h->off_pslot_of_body = o_psb; h->off_pair_pi = o_ppi; h->off_pair_pj = o_ppj;
            h->off_bp_start = o_bps; h->off_bp_list = o_bpl;
            c->n_pslots = n_ps;
        }
        h->n_slots = n_slots; h->off_slot_of_body = o_sob; h->off_pt_slot = o_pts; h->off_pair_si = o_psi; h->off_pair_sj = o_psj;
        c->n_slots = n_slots;
    }
    c->cm_match = trk::static_cm_match(m->family, w);
    c->device = m->device; c->L = m->L; c->S = S; c->nObj = nObj; c->K = K; c->P = P; c->ee_enabled = h->ee_enabled; c->nwords = nwords;
    DeviceGuard guard(m->device);
    cudaError_t e = cudaMalloc(&c->d_words, nwords * sizeof(float));
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaMalloc(cost)"); }
    e = cudaMemcpy(c->d_words, w, nwords * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(c->d_words); delete c; return cuda_fail(e, "cudaMemcpy(cost)"); }
    if (m->ext_family) {
        const int rc = build_ext_tables(m, d, c);
        if (rc != TRK_OK) { cudaFree(c->d_words); delete c; return rc; }
    }
    *out = c;
    return TRK_OK;
}

void trk_cost_destroy(trk_cost* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    cudaFree(c->hp_q); cudaFree(c->hp_cost); cudaFree(c->hp_grad);
    for (int i = 0; i < 3; ++i) if (c->hp_s[i]) cudaStreamDestroy(c->hp_s[i]);
    for (int i = 0; i < c->hp_events; ++i) { cudaEventDestroy(c->hp_in[i]); cudaEventDestroy(c->hp_k[i]); }
    cudaFree(c->d_words);
    cudaFree(c->d_ext);
    delete c;
}

static int cost_upload_header(trk_cost* c, void* stream) {
    DeviceGuard guard(c->device);
    // the header is 256 bytes; cudaMemcpyAsync from pageable memory stages the bytes before returning
    cudaError_t e = cudaMemcpyAsync(c->d_words, c->h_words.data(), sizeof(CostHeader), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "cudaMemcpyAsync(cost header)");
}

int trk_cost_set_ee_target(trk_cost* c, const float* H, void* stream) {
    TRK_CHECK_PTR(c, "cost"); TRK_CHECK_PTR(H, "H_target16");
    CostHeader* h = reinterpret_cast<CostHeader*>(c->h_words.data());
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) h->ee_R[i * 3 + j] = H[i * 4 + j];
        h->ee_t[i] = H[i * 4 + 3];
    }
    return cost_upload_header(c, stream);
}

int trk_cost_set_weights(trk_cost* c, const float* w4, void* stream) {
    TRK_CHECK_PTR(c, "cost"); TRK_CHECK_PTR(w4, "weights4");
    CostHeader* h = reinterpret_cast<CostHeader*>(c->h_words.data());
    for (int i = 0; i < 4; ++i) h->w[i] = w4[i];
    return cost_upload_header(c, stream);
}

static int objective_impl(const trk_model* m, const trk_cost* c, const float* q, int64_t n, const float* ee_target_cfg,
                          float margin_override, float* cost_total, float* cost_terms, float* grad_q, uint8_t* occupancy,
                          int32_t* argmin, float* block_sums, int* n_block_sums, void* stream);

int trk_objective(const trk_model* m, const trk_cost* c, const float* q, int64_t n, const float* ee_target_cfg,
                  float margin_override, float* cost_total, float* cost_terms, float* grad_q, uint8_t* occupancy,
                  int32_t* argmin, void* stream) {
    return objective_impl(m, c, q, n, ee_target_cfg, margin_override, cost_total, cost_terms, grad_q, occupancy, argmin, nullptr, nullptr, stream);
}

int trk_objective_sums(const trk_model* m, const trk_cost* c, const float* q, int64_t n, const float* ee_target_cfg,
                       float margin_override, float* cost_total, float* grad_q, float* block_sums, void* stream) {
    TRK_CHECK_PTR(block_sums, "block_sums");
    TRK_CHECK_PTR(m, "model");
    DeviceGuard guard(m->device);
    cudaError_t e = cudaMemsetAsync(block_sums, 0, sizeof(float) * TRK_MAX_BLOCK_SUMS, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(block_sums)");
    if (n == 0) return TRK_OK;
    int n_sums = 0;
    const int rc = objective_impl(m, c, q, n, ee_target_cfg, margin_override, cost_total, nullptr, grad_q, nullptr, nullptr, block_sums, &n_sums, stream);
    if (rc != TRK_OK) return rc;
    if (n_sums <= 0 || n_sums > TRK_MAX_BLOCK_SUMS)
        return fail(TRK_E_UNSUPPORTED, "this model / cost table is served by a kernel without the fused cost reduction (sum cost_total instead)");
    return TRK_OK;
}

static int objective_impl(const trk_model* m, const trk_cost* c, const float* q, int64_t n, const float* ee_target_cfg,
                          float margin_override, float* cost_total, float* cost_terms, float* grad_q, uint8_t* occupancy,
                          int32_t* argmin, float* block_sums, int* n_block_sums, void* stream) {
    TRK_CHECK_PTR(m, "model"); TRK_CHECK_PTR(c, "cost");
    if (n_block_sums) *n_block_sums = 0;
    if (c->device != m->device || c->L != m->L) return fail(TRK_E_RANGE, "cost tables were built for a different model");
    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(q, "q");
    if (cost_terms) TRK_CHECK_ALIGN16(cost_terms, "cost_terms");
    const int use_override = std::isnan(margin_override) ? 0 : 1;
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->family && trk::static_objective_supported(m->family)) {
        const CostHeader* hh = reinterpret_cast<const CostHeader*>(c->h_words.data());
        const int clamp_any = hh->clamp_self | hh->clamp_obj | hh->clamp_ws | (hh->KB > 0 ? 1 : 0) | (hh->grid_obj >= 0 ? 1 : 0);   // general-tables variant
        cudaError_t e = trk::launch_static_objective(m->family, m->n_sms, c->h_words.data(), c->d_words, c->nwords, c->n_slots, clamp_any, c->cm_match, q, n, ee_target_cfg,
                                                     margin_override, use_override, cost_total, cost_terms, grad_q,
                                                     occupancy, argmin, block_sums, n_block_sums, st);
        return e == cudaSuccess ? TRK_OK : cuda_fail(e, "objective(static)");
    }
    // one equal-radius identity-pose sphere object, no relu on the object term: collision points take the joint scan
    const CostHeader* gh = reinterpret_cast<const CostHeader*>(c->h_words.data());
    const int* giw = reinterpret_cast<const int*>(c->h_words.data());
    if (m->ext_family && c->ext_ok && !ee_target_cfg) {
        // a compiled-in robot with attached bodies in the common scene: unrolled FK + points evaluated at their anchors (trk_ext.cuh)
        static const bool no_hot = getenv("TRK_B200_NO_HOT") != nullptr;   // A/B switch
        const int need = OBJ_IDENTITY | OBJ_UNIFORM_RADIUS;
        const bool sphere_scene = gh->nObj == 1 && gh->KB == 0 && (giw[gh->off_obj_ident] & (need | OBJ_GRID)) == need &&
                                  giw[gh->off_obj_start + 1] > giw[gh->off_obj_start];
        const bool scene = (sphere_scene || trk::general_scene_ok(c->h_words.data())) && !gh->clamp_self && !gh->clamp_obj && !gh->clamp_ws &&
                           gh->ws_enabled &&
                           (!gh->ee_enabled || (gh->ee_w_pos > 0.f && gh->ee_w_rot > 0.f && gh->ee_link == trk::static_family_last_body(m->ext_family)));
        if (scene && !no_hot) {
            cudaError_t e = trk::launch_static_objective_ext(m->ext_family, m->n_sms, c->h_words.data(), c->d_words, c->d_ext, c->ext_nwords, c->h_ext.data(),
                                                             gh->S, q, n, margin_override, use_override, cost_total, cost_terms, grad_q, occupancy, argmin, st);
            if (e == cudaSuccess) return TRK_OK;
            if (e != cudaErrorInvalidValue) return cuda_fail(e, "objective(ext)");
            cudaGetLastError();      // tables too large for the per-thread shared-memory scratch: the runtime-tree kernel serves the call
        }
    }
    const int joint = gh->nObj == 1 && gh->KB == 0 && !gh->clamp_obj && gh->S > 0 &&
                      (giw[gh->off_obj_ident] & (OBJ_IDENTITY | OBJ_UNIFORM_RADIUS | OBJ_GRID)) == (OBJ_IDENTITY | OBJ_UNIFORM_RADIUS) &&
                      giw[gh->off_obj_start + 1] > giw[gh->off_obj_start];
    const GenericLaunch gl = generic_launch(m, n, (size_t)m->nwords + align4(c->nwords) + TRK_MAX_LINKS, 0, grad_q ? m->D : 0,
                                            joint ? c->n_slots : c->n_pslots);
    cudaError_t e = allow_smem(k_objective_generic<true>, gl.smem);
    if (e == cudaSuccess) e = allow_smem(k_objective_generic<false>, gl.smem);
    if (e != cudaSuccess) return cuda_fail(e, "objective smem");
    if (grad_q)
        k_objective_generic<true><<<gl.grid, gl.warps * 32, gl.smem, st>>>(m->d_words, m->nwords, c->d_words, c->nwords, m->live, q, n,
            ee_target_cfg, margin_override, use_override, joint, cost_total, cost_terms, grad_q, occupancy, argmin);
    else
        k_objective_generic<false><<<gl.grid, gl.warps * 32, gl.smem, st>>>(m->d_words, m->nwords, c->d_words, c->nwords, m->live, q, n,
            ee_target_cfg, margin_override, use_override, joint, cost_total, cost_terms, grad_q, occupancy, argmin);
    e = cudaGetLastError();
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "objective launch");
}

static int ik_impl(const trk_model* m, const trk_ik_params* p, const float* step_size, const float* bc2_sqrt,
                   const float* target_seed, int64_t n, float* q_io, float* err_out, uint8_t* valid_out, int32_t* valid_count,
                   float* loss_sum, float* loss_sqsum, int64_t* best_keys, int best_k, void* stream);

int trk_ik_adam(const trk_model* m, const trk_ik_params* p, const float* step_size, const float* bc2_sqrt,
                const float* target_seed, int64_t n, float* q_io, float* err_out, uint8_t* valid_out, int32_t* valid_count,
                float* loss_sum, float* loss_sqsum, void* stream) {
    return ik_impl(m, p, step_size, bc2_sqrt, target_seed, n, q_io, err_out, valid_out, valid_count, loss_sum, loss_sqsum, nullptr, 0, stream);
}

int trk_ik_adam_best(const trk_model* m, const trk_ik_params* p, const float* step_size, const float* bc2_sqrt,
                     const float* target_seed, int64_t n, float* q_io, float* err_out, uint8_t* valid_out, int32_t* valid_count,
                     int64_t* best_keys, int32_t best_k, void* stream) {
    TRK_CHECK_PTR(m, "model"); TRK_CHECK_PTR(p, "params"); TRK_CHECK_PTR(best_keys, "best_keys");
    if (best_k < 1 || best_k > TRK_IK_BLOCK_SEEDS) return fail(TRK_E_RANGE, "best_k=%d outside [1, %d]", best_k, TRK_IK_BLOCK_SEEDS);
    if (n >= (1ll << 32)) return fail(TRK_E_RANGE, "n=%lld: seed indices must fit 32 bits", (long long)n);
    static const bool no_hot = getenv("TRK_B200_NO_HOT") != nullptr;
    if (!m->family || no_hot || p->ee_link != trk::static_family_last_body(m->family))
        return fail(TRK_E_UNSUPPORTED, "the fused per-block top-k needs a compiled-in robot with the EE on its last body (use trk_ik_adam and select on the host side)");
    return ik_impl(m, p, step_size, bc2_sqrt, target_seed, n, q_io, err_out, valid_out, valid_count, nullptr, nullptr, best_keys, best_k, stream);
}

static int ik_impl(const trk_model* m, const trk_ik_params* p, const float* step_size, const float* bc2_sqrt,
                   const float* target_seed, int64_t n, float* q_io, float* err_out, uint8_t* valid_out, int32_t* valid_count,
                   float* loss_sum, float* loss_sqsum, int64_t* best_keys, int best_k, void* stream) {
    TRK_CHECK_PTR(m, "model"); TRK_CHECK_PTR(p, "params");

    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (p->n_iters < 1 || p->n_updates < 0 || p->n_updates > p->n_iters) return fail(TRK_E_RANGE, "n_iters=%d n_updates=%d", p->n_iters, p->n_updates);
    if (p->ee_link < 0 || p->ee_link >= m->L) return fail(TRK_E_RANGE, "ee_link=%d", p->ee_link);
    TRK_CHECK_PTR(p->lower, "lower"); TRK_CHECK_PTR(p->upper, "upper");
    if (!target_seed) TRK_CHECK_PTR(p->H_target, "H_target");
    TRK_CHECK_PTR(step_size, "step_size"); TRK_CHECK_PTR(bc2_sqrt, "bc2_sqrt"); TRK_CHECK_PTR(valid_count, "valid_count");
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(q_io, "q_io");
    IKParams P;
    memset(&P, 0, sizeof(P));
    P.ee_link = p->ee_link; P.n_iters = p->n_iters; P.n_updates = p->n_updates; P.per_seed_target = target_seed ? 1 : 0;
    P.lr = p->lr; P.beta1 = p->beta1; P.beta2 = p->beta2; P.eps = p->eps;
    P.w_se3 = p->w_se3; P.w_lim = p->w_joint_limits; P.se3_eps = p->se3_eps;
    for (int c = 0; c < m->D; ++c) { P.lower[c] = p->lower[c]; P.upper[c] = p->upper[c]; }
    if (p->H_target) memcpy(P.target, p->H_target, sizeof(P.target));
    P.best_keys = reinterpret_cast<unsigned long long*>(best_keys); P.best_k = best_k;
    for (int b = p->ee_link; b >= 0; b = m->parent[b])   // q columns of the joints between the root and the EE body
        if (m->qidx[b] >= 0) P.anc_mask |= 1u << m->qidx[b];
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(valid_count, 0, sizeof(int32_t) * p->n_iters, st);
    if (e == cudaSuccess && loss_sum) e = cudaMemsetAsync(loss_sum, 0, sizeof(float) * p->n_iters, st);
    if (e == cudaSuccess && loss_sqsum) e = cudaMemsetAsync(loss_sqsum, 0, sizeof(float) * p->n_iters, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(ik statistics)");
    if (loss_sum && !loss_sqsum) return fail(TRK_E_NULL, "loss_sqsum is required with loss_sum");
    if (m->family) {
        e = trk::launch_static_ik(m->family, P, step_size, bc2_sqrt, target_seed, n, q_io, err_out, valid_out, valid_count, loss_sum,
                                  loss_sqsum, st);
    } else {
        const size_t per_warp = (size_t)(m->live.n_live * 12 + m->D * 9) * 32 * 4;
        int warps = kWarps;
        while (warps > 1 && (size_t)m->nwords * 4 + warps * per_warp > 100 * 1024) --warps;
        const size_t smem = (size_t)m->nwords * 4 + warps * per_warp;
        const long long blocks = (n + warps * 32 - 1) / (warps * 32);
        e = allow_smem(k_ik_adam_generic, smem);
        if (e == cudaSuccess) {
            k_ik_adam_generic<<<(unsigned)blocks, warps * 32, smem, st>>>(m->d_words, m->nwords, m->live, P, step_size, bc2_sqrt, target_seed, n,
                                                                          q_io, err_out, valid_out, valid_count, loss_sum, loss_sqsum);
            e = cudaGetLastError();
        }
    }
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "ik_adam launch");
}

int trk_link_jacobian(const trk_model* m, const float* q, int64_t n, float* J, void* stream) {
    TRK_CHECK_PTR(m, "model");
    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (m->D < 1) return fail(TRK_E_RANGE, "the model has no movable joint");
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(q, "q"); TRK_CHECK_PTR(J, "J");
    TRK_CHECK_ALIGN16(J, "J");
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->family) {       // compiled-in robots whose tile fits: FK in registers, one bulk store per 32 configurations
        static const bool no_hot = getenv("TRK_B200_NO_HOT") != nullptr;
        bool served = false;
        if (!no_hot) {
            cudaError_t es = trk::launch_static_link_jacobian(m->family, m->n_sms, q, n, J, st, &served);
            if (es != cudaSuccess) return cuda_fail(es, "link_jacobian(static) launch");
            if (served) return TRK_OK;
        }
    }
    const int rec_stride = (m->L * 7 + m->D * 3) | 1;                       // odd: lanes hit distinct banks in phase 1
    int link_chunk = 512 / (7 * m->D);                                       // link blocks per output tile (~2 KB)
    link_chunk = link_chunk < 1 ? 1 : link_chunk;
    const size_t tile_words = ((size_t)link_chunk * 7 * m->D + 4 + 3) & ~(size_t)3;
    // configurations per warp pass: shared memory per warp decides how many warps an SM holds, and the kernel is
    // latency-bound, so trade phase-1 lane use for residency (TRK_B200_JAC_BUDGET: bytes per warp, default 24 KB)
    static const long budget = [] { const char* v = getenv("TRK_B200_JAC_BUDGET"); long b = v ? atol(v) : 0; return b > 0 ? b : 24576L; }();
    const size_t per_cfg_words = (size_t)rec_stride + 12 * (size_t)m->live.n_live;
    int cpw = 32;
    while (cpw > 4 && (2 * tile_words + cpw * per_cfg_words) * 4 > (size_t)budget) cpw >>= 1;
    const size_t per_warp = (2 * tile_words + cpw * per_cfg_words) * 4;
    const size_t fixed = (size_t)m->nwords * 4 + (size_t)(TRK_MAX_LINKS + 64) * 4;
    const int warps = kWarps;
    const size_t smem = fixed + warps * per_warp;
    long long tiles = (n + (long long)cpw * warps - 1) / ((long long)cpw * warps);
    const long long cap = (long long)m->n_sms * 16;
    const int grid = (int)(tiles < cap ? tiles : cap);
    cudaError_t e = allow_smem(k_link_jacobian, smem);
    if (e == cudaSuccess)
        k_link_jacobian<<<grid, warps * 32, smem, st>>>(m->d_words, m->nwords, m->live, q, n, cpw, rec_stride, link_chunk, J);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "link_jacobian launch");
}

int trk_geometric_jacobian(const trk_model* m, const float* q, int64_t n, int32_t link, uint32_t col_mask, float* pose12,
                           float* lin, float* ang, void* stream) {
    TRK_CHECK_PTR(m, "model");
    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (link < 0 || link >= m->L) return fail(TRK_E_RANGE, "link=%d outside the model", link);
    if (m->D < 1) return fail(TRK_E_RANGE, "the model has no movable joint");
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(q, "q"); TRK_CHECK_PTR(pose12, "pose12"); TRK_CHECK_PTR(lin, "lin"); TRK_CHECK_PTR(ang, "ang");
    DeviceGuard guard(m->device);
    const size_t per_warp = (size_t)(m->live.n_live * 12 + m->D * 6) * 32 * 4;
    int warps = kWarps;
    while (warps > 1 && (size_t)m->nwords * 4 + warps * per_warp > 100 * 1024) --warps;
    const size_t smem = (size_t)m->nwords * 4 + warps * per_warp;
    const long long need = (n + warps * 32 - 1) / (warps * 32), cap = (long long)m->n_sms * 16;
    cudaError_t e = allow_smem(k_geometric_jacobian, smem);
    if (e == cudaSuccess) {
        k_geometric_jacobian<<<(int)(need < cap ? need : cap), warps * 32, smem, (cudaStream_t)stream>>>(m->d_words, m->nwords, m->live, q, n, link,
                                                                                                  col_mask, pose12, lin, ang);
        e = cudaGetLastError();
    }
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "geometric_jacobian launch");
}

int trk_selftest_sqrt(const float* x, int64_t n, float* y_core, float* y_ref, void* stream) {
    if (n <= 0) return TRK_OK;
    TRK_CHECK_PTR(x, "x"); TRK_CHECK_PTR(y_core, "y_core"); TRK_CHECK_PTR(y_ref, "y_ref");
    k_selftest_sqrt<<<1184, 256, 0, (cudaStream_t)stream>>>(x, n, y_core, y_ref);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "selftest_sqrt launch");
}

int trk_objective_host(const trk_model* m, trk_cost* c, const float* q_host, int64_t n, int32_t n_chunks, float* cost_host,
                       float* grad_host) {
    TRK_CHECK_PTR(m, "model"); TRK_CHECK_PTR(c, "cost");
    if (c->device != m->device || c->L != m->L) return fail(TRK_E_RANGE, "cost tables were built for a different model");
    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(q_host, "q_host");
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > 64) n_chunks = 64;
    DeviceGuard guard(m->device);
    const int D = m->D;
    cudaError_t e = cudaSuccess;
    if (!c->hp_s[0]) {
        for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&c->hp_s[i], cudaStreamNonBlocking);
        for (; c->hp_events < 64 && e == cudaSuccess; ++c->hp_events) {
            e = cudaEventCreateWithFlags(&c->hp_in[c->hp_events], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->hp_k[c->hp_events], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) return cuda_fail(e, "objective_host: stream/event creation");
    }
    if (n > c->hp_cap) {
        cudaFree(c->hp_q); cudaFree(c->hp_cost); cudaFree(c->hp_grad);
        c->hp_q = c->hp_cost = c->hp_grad = nullptr; c->hp_cap = 0;
        e = cudaMalloc(&c->hp_q, (size_t)n * D * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&c->hp_cost, (size_t)n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&c->hp_grad, (size_t)n * D * sizeof(float));
        if (e != cudaSuccess) return cuda_fail(e, "objective_host: staging allocation");
        c->hp_cap = n;
    }
    const int64_t step = (n + n_chunks - 1) / n_chunks;
    int ci = 0;
    for (int64_t lo = 0; lo < n; lo += step, ++ci) {
        const int64_t cnt = (lo + step <= n) ? step : n - lo;
        e = cudaMemcpyAsync(c->hp_q + lo * D, q_host + lo * D, (size_t)cnt * D * sizeof(float), cudaMemcpyHostToDevice, c->hp_s[0]);
        if (e == cudaSuccess) e = cudaEventRecord(c->hp_in[ci], c->hp_s[0]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->hp_s[1], c->hp_in[ci], 0);
        if (e != cudaSuccess) return cuda_fail(e, "objective_host: H2D");
        const int rc = trk_objective(m, c, c->hp_q + lo * D, cnt, nullptr, NAN, c->hp_cost + lo, nullptr,
                                     grad_host ? c->hp_grad + lo * D : nullptr, nullptr, nullptr, c->hp_s[1]);
        if (rc) return rc;
        e = cudaEventRecord(c->hp_k[ci], c->hp_s[1]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->hp_s[2], c->hp_k[ci], 0);
        if (e == cudaSuccess && cost_host)
            e = cudaMemcpyAsync(cost_host + lo, c->hp_cost + lo, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, c->hp_s[2]);
        if (e == cudaSuccess && grad_host)
            e = cudaMemcpyAsync(grad_host + lo * D, c->hp_grad + lo * D, (size_t)cnt * D * sizeof(float), cudaMemcpyDeviceToHost, c->hp_s[2]);
        if (e != cudaSuccess) return cuda_fail(e, "objective_host: D2H");
    }
    e = cudaStreamSynchronize(c->hp_s[2]);
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "objective_host: synchronize");
}

int trk_cost_stage(const trk_model* m, const trk_cost* c, const float* link_pos, const float* ee_pose, int64_t n,
                   const float* ee_target_cfg, float margin_override, float* cost_total, float* cost_terms,
                   float* grad_pos, float* grad_ee, uint8_t* occupancy, int32_t* argmin, void* stream) {
    TRK_CHECK_PTR(m, "model"); TRK_CHECK_PTR(c, "cost");
    if (c->device != m->device || c->L != m->L) return fail(TRK_E_RANGE, "cost tables were built for a different model");
    if (n < 0) return fail(TRK_E_RANGE, "n=%lld", (long long)n);
    if (n == 0) return TRK_OK;
    TRK_CHECK_PTR(link_pos, "link_pos");
    if (c->ee_enabled && !ee_pose) return fail(TRK_E_NULL, "ee_pose is required when the EE term is enabled");
    if (cost_terms) TRK_CHECK_ALIGN16(cost_terms, "cost_terms");
    const int use_override = std::isnan(margin_override) ? 0 : 1;
    DeviceGuard guard(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)c->nwords * 4;
    const int grid = grid_for(m, n, 16);
    const bool want_grad = grad_pos || grad_ee;
#define LAUNCH_STAGE(MAXL)                                                                                              \
    do {                                                                                                                \
        if (want_grad)                                                                                                  \
            k_cost_stage<MAXL, true><<<grid, kBlock, smem, st>>>(c->d_words, c->nwords, m->L, link_pos, ee_pose, n,     \
                ee_target_cfg, margin_override, use_override, cost_total, cost_terms, grad_pos, grad_ee, occupancy, argmin); \
        else                                                                                                            \
            k_cost_stage<MAXL, false><<<grid, kBlock, smem, st>>>(c->d_words, c->nwords, m->L, link_pos, ee_pose, n,    \
                ee_target_cfg, margin_override, use_override, cost_total, cost_terms, grad_pos, grad_ee, occupancy, argmin); \
    } while (0)
    switch (m->maxl_bucket) {
        case 16: LAUNCH_STAGE(16); break;
        case 32: LAUNCH_STAGE(32); break;
        default: LAUNCH_STAGE(64); break;
    }
#undef LAUNCH_STAGE
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TRK_OK : cuda_fail(e, "cost_stage launch");
}

}  // extern "C"
