// Fused objective for a compiled-in robot that carries ATTACHED fixed bodies: grasped-object points
// (environments/objects.py:37-89, robots/robot_panda.py:155-168) and per-link multi-sphere collision models
// (data/configs/*_sphere_config.yaml).  DifferentiableTree.attach_points / attach_link append such bodies to the tree as
// fixed leaves, which used to move the whole robot to the runtime-tree kernels (45-sphere Panda: 4.6e8 configurations/s).
// Here the base robot keeps its unrolled FK (HTree of trk_hot.cuh); every body the cost tables reference - base or
// attached - is described to the kernel as (movable anchor body, constant offset in the anchor's frame), folded on the host
// from the fixed chain above it (trk_cost_create).  At each movable body the kernel
//   * parks the positions of the self-collision-pair bodies anchored there in per-thread shared-memory slots,
//   * evaluates the collision points anchored there in groups of five through the joint sphere scan of the headline kernel
//     (one table load serves five points, FFMA2 keys) and finishes them from registers, accumulating the body's wrench,
// in ONE non-inlined device function (the unrolled FK calls it once per movable body: inlined it would be 70 KB of code).
// Pairs run afterwards from the slots; the reverse sweep is the compile-time one, reading each body's wrench from shared
// memory.  Scene: one equal-radius identity-pose sphere object, workspace box, EE term on the robot's last body, no relu.
#pragma once

#include "trk_hot.cuh"

namespace trk {

#ifndef TRK_EXT_MINBLOCKS
#define TRK_EXT_MINBLOCKS 4
#endif

struct ExtHeader {            // device tables built by trk_cost_create for a model with attached bodies
    int n_pts, n_ps, n_pairs, n_anchor;         // collision points, pair-position slots, pairs, movable bodies + 1 (root)
    int off_pt_start, off_pt, off_ps_start, off_ps, off_ps_col, off_pair, total_words, pad;
};
// words after the header:
//   pt_start [n_anchor + 1]   points of anchor a (0 = root, 1 + c = the body q column c drives): pt[pt_start[a] .. pt_start[a+1])
//   pt       [n_pts][8]       offset xyz, margin (object), margin (ws), point index, 2 pad   (a margin override is a kernel parameter)
//   ps_start [n_anchor + 1], ps [n_ps][4] offset xyz + pad, ps_col [n_ps] q column of the slot's anchor (-1: root)
//   pair     [n_pairs][4]     slot i, slot j, margin, q columns of the two anchors (col_i | col_j << 8, 0xff = root: no joint above);
//                             sorted by slot i, so a run of pairs accumulates the force on body i in registers

struct ExtView {
    const ExtHeader* h;
    const int* pt_start; const float* pt; const int* ps_start; const float* ps; const int* ps_col; const float* pair;
    __device__ explicit ExtView(const float* w) {
        h = reinterpret_cast<const ExtHeader*>(w);
        const int* iw = reinterpret_cast<const int*>(w);
        pt_start = iw + h->off_pt_start; pt = w + h->off_pt; ps_start = iw + h->off_ps_start; ps = w + h->off_ps;
        ps_col = iw + h->off_ps_col; pair = w + h->off_pair;
    }
};

// per-thread shared-memory scratch, [slot][thread] layout (stride kSBlock)
struct ExtScratch {
    float* PS;      // [3 n_ps]  positions of the pair bodies
    float* WS;      // [6 D]     wrench (F, T about the world origin) of the bodies anchored at each movable body
    float* ACC;     // [3]       c_obj, c_ws, occupancy bits
};

// Everything anchored at one movable body (or the root, col = -1) whose world pose is (Rw, pw).
static __device__ __noinline__ void ext_at_anchor(float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7, float r8,
                                           float px, float py, float pz, int anchor, int col, const float* __restrict__ ext_words,
                                           const HotParams* __restrict__ hp, const float4* __restrict__ tq, const float4* __restrict__ tsph,
                                           ExtScratch sc, int* __restrict__ argmin_row, bool want_grad) {
    const ExtView ev(ext_words);
    // (1) pair bodies anchored here
    for (int k = ev.ps_start[anchor]; k < ev.ps_start[anchor + 1]; ++k) {
        const float tx = ev.ps[4 * k], ty = ev.ps[4 * k + 1], tz = ev.ps[4 * k + 2];
        sc.PS[(3 * k + 0) * kSBlock] = fmaf(r2, tz, fmaf(r1, ty, fmaf(r0, tx, px)));
        sc.PS[(3 * k + 1) * kSBlock] = fmaf(r5, tz, fmaf(r4, ty, fmaf(r3, tx, py)));
        sc.PS[(3 * k + 2) * kSBlock] = fmaf(r8, tz, fmaf(r7, ty, fmaf(r6, tx, pz)));
    }
    // (2) collision points anchored here, five at a time
    float Fx = 0.f, Fy = 0.f, Fz = 0.f, Tx = 0.f, Ty = 0.f, Tz = 0.f, c_obj = 0.f, c_ws = 0.f;
    bool occ_obj = false, occ_ws = false;
    const int k1 = ev.pt_start[anchor + 1];
    const float kInf = __int_as_float(0x7f800000);
    const unsigned keep = ~hp->lowmask;
#pragma unroll 1
    for (int k0 = ev.pt_start[anchor]; k0 < k1; k0 += 5) {
        const int np = (k1 - k0) < 5 ? (k1 - k0) : 5;
        float X[5][3];
#pragma unroll
        for (int p = 0; p < 5; ++p) {      // lanes of the group beyond np repeat the last valid point and are ignored
            const float* rec = ev.pt + 8 * (k0 + (p < np ? p : np - 1));
            const float tx = rec[0], ty = rec[1], tz = rec[2];
            X[p][0] = fmaf(r2, tz, fmaf(r1, ty, fmaf(r0, tx, px)));
            X[p][1] = fmaf(r5, tz, fmaf(r4, ty, fmaf(r3, tx, py)));
            X[p][2] = fmaf(r8, tz, fmaf(r7, ty, fmaf(r6, tx, pz)));
        }
        float m1[5], m2[5];
#pragma unroll
        for (int p = 0; p < 5; ++p) { m1[p] = kInf; m2[p] = kInf; }
        unsigned idx = 0u;
#pragma unroll 1
        for (int j = 0; j < hp->n_pairs2; ++j, idx += 2u) {
            const float4 t0 = tq[2 * j], t1 = tq[2 * j + 1];
            const float2 cx = make_float2(t0.x, t0.y), cy = make_float2(t0.z, t0.w), cz = make_float2(t1.x, t1.y), cw = make_float2(t1.z, t1.w);
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const float2 k2 = __ffma2_rn(make_float2(X[p][0], X[p][0]), cx,
                                             __ffma2_rn(make_float2(X[p][1], X[p][1]), cy, __ffma2_rn(make_float2(X[p][2], X[p][2]), cz, cw)));
                const float pa = __uint_as_float((__float_as_uint(k2.x) & keep) | idx);
                const float pb = __uint_as_float((__float_as_uint(k2.y) & keep) | (idx + 1u));
                const float lo = fminf(pa, pb), hi = fmaxf(pa, pb);
                m2[p] = fmin3(fmaxf(m1[p], lo), m2[p], hi);
                m1[p] = fminf(m1[p], lo);
            }
        }
#pragma unroll
        for (int p = 0; p < 5; ++p) {
            if (p < np) {
                const float* rec = ev.pt + 8 * (k0 + p);
                const float x = X[p][0], y = X[p][1], z = X[p][2];
                const float xx = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
                int widx = (int)(__float_as_uint(m1[p]) & hp->lowmask);
                if (!((m2[p] - m1[p]) > hp->tol * (xx + hp->c2max))) {      // near tie (or NaN): exact scan, first index wins (rare)
                    float best = kInf;
                    widx = 0;
#pragma unroll 1
                    for (int kk = 0; kk < hp->n_sph; ++kk) {
                        const float4 c = tsph[kk];
                        const float sd = __fsub_rn(norm3_rn(__fsub_rn(x, c.x), __fsub_rn(y, c.y), __fsub_rn(z, c.z)), c.w);
                        if (sd < best) { best = sd; widx = kk; }
                    }
                }
                const float4 c = tsph[widx];
                const float vx = __fsub_rn(x, c.x), vy = __fsub_rn(y, c.y), vz = __fsub_rn(z, c.z);
                float rs;
                const float best = __fsub_rn(norm3_clamped(vx, vy, vz, rs), c.w);
                if (argmin_row) argmin_row[__float_as_int(rec[5])] = widx;
                occ_obj = occ_obj || (best < (hp->use_override ? hp->margin_override : rec[3]));
                c_obj = __fadd_rn(c_obj, -__fsub_rn(best, rec[3]));
                const float e0 = __fsub_rn(x, hp->ws_min[0]), e1 = __fsub_rn(y, hp->ws_min[1]), e2 = __fsub_rn(z, hp->ws_min[2]);
                const float f0 = __fsub_rn(hp->ws_max[0], x), f1 = __fsub_rn(hp->ws_max[1], y), f2 = __fsub_rn(hp->ws_max[2], z);
                const float emin = fmin3(e0, e1, e2), fmn = fmin3(f0, f1, f2);
                const float dmin = fminf(emin, fmn);
                occ_ws = occ_ws || (dmin < (hp->use_override ? hp->margin_override : rec[4]));
                c_ws = __fadd_rn(c_ws, -__fsub_rn(dmin, rec[4]));
                if (want_grad) {
                    const float w1 = hp->w[1] * rs;
                    float gx = -(w1 * vx), gy = -(w1 * vy), gz = -(w1 * vz);
                    const bool upper = fmn < emin;                       // ties go to the lower index: the x - min side
                    float g = upper ? hp->w[2] : -hp->w[2];
                    g = (dmin != 0.f) ? g : 0.f;
                    const float a0 = upper ? f0 : e0, a1 = upper ? f1 : e1;
                    const bool p0 = a0 == dmin, p1 = a1 == dmin;
                    gx += p0 ? g : 0.f;
                    gy += (!p0 && p1) ? g : 0.f;
                    gz += (!p0 && !p1) ? g : 0.f;
                    Fx += gx; Fy += gy; Fz += gz;
                    Tx = fmaf(y, gz, fmaf(-z, gy, Tx));
                    Ty = fmaf(z, gx, fmaf(-x, gz, Ty));
                    Tz = fmaf(x, gy, fmaf(-y, gx, Tz));
                }
            }
        }
    }
    if (col >= 0) {
        float* w = sc.WS + 6 * col * kSBlock;
        w[0] = Fx; w[kSBlock] = Fy; w[2 * kSBlock] = Fz; w[3 * kSBlock] = Tx; w[4 * kSBlock] = Ty; w[5 * kSBlock] = Tz;
    }
    sc.ACC[0] += c_obj;
    sc.ACC[kSBlock] += c_ws;
    if (occ_obj || occ_ws)
        sc.ACC[2 * kSBlock] = __int_as_float(__float_as_int(sc.ACC[2 * kSBlock]) | (occ_obj ? (int)TRK_OCC_OBJECT : 0) | (occ_ws ? (int)TRK_OCC_WS : 0));
}

template <class R, int EE, bool WANT_GRAD>
__global__ void __launch_bounds__(kSBlock, TRK_EXT_MINBLOCKS) k_objective_ext(const __grid_constant__ HotParams hp_in, const float* __restrict__ cost_words,
                                                           const float* __restrict__ ext_words_g, int ext_nwords, int n_ps,
                                                           const float* __restrict__ q, long long n, float* __restrict__ cost_total,
                                                           float* __restrict__ cost_terms, float* __restrict__ grad_q,
                                                           uint8_t* __restrict__ occupancy, int* __restrict__ argmin, int n_points) {
    using T = HTree<R>;
    constexpr int L = R::L, D = R::D;
    extern __shared__ __align__(16) float esm[];
    // tables: scan table | sphere table | HotParams copy | ext words ; then the per-thread scratch columns
    float4* tq = reinterpret_cast<float4*>(esm);
    float4* tsph = tq + 2 * hp_in.n_pairs2;
    HotParams* hp = reinterpret_cast<HotParams*>(tsph + hp_in.n_sph);
    float* ew = reinterpret_cast<float*>(hp) + ((sizeof(HotParams) + 15) / 16) * 4;
    float* scratch = ew + ((ext_nwords + 3) & ~3);
    const int tid = threadIdx.x;
    {
        const float* sphq = cost_words + hp_in.off_sphq;
        for (int i = tid; i < hp_in.n_pairs2 * 8; i += kSBlock) {
            const int k = i >> 2, comp = i & 3;
            esm[(k >> 1) * 8 + comp * 2 + (k & 1)] = __ldg(sphq + i);
        }
        const float4* sph = reinterpret_cast<const float4*>(cost_words + hp_in.off_sph);
        for (int i = tid; i < hp_in.n_sph; i += kSBlock) tsph[i] = __ldg(sph + i);
        const float* hsrc = reinterpret_cast<const float*>(&hp_in);
        for (int i = tid; i < (int)(sizeof(HotParams) / 4); i += kSBlock) reinterpret_cast<float*>(hp)[i] = hsrc[i];
        for (int i = tid; i < ext_nwords; i += kSBlock) ew[i] = __ldg(ext_words_g + i);
    }
    __syncthreads();
    ExtScratch sc;
    sc.PS = scratch + tid;
    sc.WS = sc.PS + 3 * n_ps * kSBlock;
    sc.ACC = sc.WS + 6 * D * kSBlock;
    const ExtView ev(ew);
    for (long long cfg = (long long)blockIdx.x * kSBlock + tid; cfg < n; cfg += (long long)gridDim.x * kSBlock) {
        const float* qrow = q + cfg * D;
        float qv[D];
        sfor<D>([&](auto c) { qv[c] = __ldg(qrow + c); });
        sc.ACC[0] = 0.f; sc.ACC[kSBlock] = 0.f; sc.ACC[2 * kSBlock] = 0.f;
        int* amrow = argmin ? argmin + cfg * n_points : nullptr;
        // ---------------- forward: unrolled FK; everything anchored at a body is evaluated while its pose is in registers ----------------
        float W[L][12], JA[D > 0 ? D : 1][3], PJ[D > 0 ? D : 1][3], Ree[9], pee[3];
        unsigned pass = 0u;
        ext_at_anchor(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0, -1, ew, hp, tq, tsph, sc, amrow, false);     // root-anchored bodies
        sfor<L>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (T::pix(i) >= 0 && T::movable(i)) {
                constexpr int qi = R::qidx[i];
                const float qraw = qv[qi];
                float theta = qraw;
                if constexpr (R::has_lim[i] != 0) {
                    constexpr float lo = R::lo[i], hi = R::hi[i];
                    theta = fmin_nan(fmax_nan(qraw, lo), hi);
                    if (theta == qraw) pass |= (1u << qi);
                } else {
                    pass |= (1u << qi);
                }
                constexpr int a = T::ap(i);
                T::template compose<i, T::small_angle(i)>(W[a], theta, W[i]);
                T::template joint_axis<i>(W[a], W[i], JA[qi]);
                sfor<3>([&](auto rc) { PJ[qi][rc] = W[i][9 + rc]; });
                ext_at_anchor(W[i][0], W[i][1], W[i][2], W[i][3], W[i][4], W[i][5], W[i][6], W[i][7], W[i][8], W[i][9], W[i][10], W[i][11],
                              1 + qi, qi, ew, hp, tq, tsph, sc, amrow, WANT_GRAD);
            }
        });
        if constexpr (T::pix(EE) < 0) {
            sfor<9>([&](auto e) { Ree[e] = (e == 0 || e == 4 || e == 8) ? 1.f : 0.f; });
            sfor<3>([&](auto rc) { pee[rc] = 0.f; });
        } else if constexpr (T::movable(EE)) {
            sfor<9>([&](auto e) { Ree[e] = W[EE][e]; });
            sfor<3>([&](auto rc) { pee[rc] = W[EE][9 + rc]; });
        } else {
            T::template fixed_rotation<EE>(W[T::anchor(EE)], Ree);
            T::template fixed_position<EE>(W[T::anchor(EE)], pee);
        }
        // ---------------- EE SE(3) term ----------------
        float c_ee, eF[3], eT[3];
        {
            const float* Rt = hp->ee_R;
            float dg[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                dg[i] = __fmaf_rn(Ree[i * 3 + 2], Rt[i * 3 + 2], __fmaf_rn(Ree[i * 3 + 1], Rt[i * 3 + 1], __fmul_rn(Ree[i * 3 + 0], Rt[i * 3 + 0])));
            const float tr = __fadd_rn(__fadd_rn(dg[0], dg[1]), dg[2]);
            const float phi = __fmul_rn(__fsub_rn(tr, 1.f), 0.5f);
            float Dd = __fmul_rn(hp->ee_w_rot, __fsub_rn(1.f, phi));
            const float vx = __fsub_rn(pee[0], hp->ee_t[0]), vy = __fsub_rn(pee[1], hp->ee_t[1]), vz = __fsub_rn(pee[2], hp->ee_t[2]);
            float rs;
            const float dist = norm3_clamped(vx, vy, vz, rs);
            Dd = __fadd_rn(Dd, __fmul_rn(hp->ee_w_pos, dist));
            c_ee = hp->ee_enabled ? (hp->ee_square ? __fmul_rn(Dd, Dd) : Dd) : 0.f;
            const float scale = hp->ee_enabled ? hp->w[3] * (hp->ee_square ? 2.f * Dd : 1.f) : 0.f;
            const float kr = scale * (-0.5f * hp->ee_w_rot);
            const float kp = scale * hp->ee_w_pos * rs;
            eF[0] = kp * vx; eF[1] = kp * vy; eF[2] = kp * vz;
            const float m21 = Rt[6] * Ree[3] + Rt[7] * Ree[4] + Rt[8] * Ree[5];
            const float m12 = Rt[3] * Ree[6] + Rt[4] * Ree[7] + Rt[5] * Ree[8];
            const float m02 = Rt[0] * Ree[6] + Rt[1] * Ree[7] + Rt[2] * Ree[8];
            const float m20 = Rt[6] * Ree[0] + Rt[7] * Ree[1] + Rt[8] * Ree[2];
            const float m10 = Rt[3] * Ree[0] + Rt[4] * Ree[1] + Rt[5] * Ree[2];
            const float m01 = Rt[0] * Ree[3] + Rt[1] * Ree[4] + Rt[2] * Ree[5];
            eT[0] = fmaf(kr, m21 - m12, fmaf(pee[1], eF[2], -(pee[2] * eF[1])));
            eT[1] = fmaf(kr, m02 - m20, fmaf(pee[2], eF[0], -(pee[0] * eF[2])));
            eT[2] = fmaf(kr, m10 - m01, fmaf(pee[0], eF[1], -(pee[1] * eF[0])));
        }
        // ---------------- self-collision pairs from the slots (distance_fields.py:194-215) ----------------
        float c_self = 0.f;
        bool occ_self = false;
        {
            // A pair's force acts along the line through the two points, so its torque about the world origin is the same at
            // both ends: body j's anchor receives (+u, +tau), body i's (-u, -tau).  Pairs are sorted by slot i: the wrench on
            // body i's anchor is kept in registers over a run and flushed when the anchor changes.
            float aFx = 0.f, aFy = 0.f, aFz = 0.f, aTx = 0.f, aTy = 0.f, aTz = 0.f;
            int acol = -1;
            auto flush = [&]() {
                if (acol >= 0) {
                    float* w = sc.WS + 6 * acol * kSBlock;
                    w[0] -= aFx; w[kSBlock] -= aFy; w[2 * kSBlock] -= aFz; w[3 * kSBlock] -= aTx; w[4 * kSBlock] -= aTy; w[5 * kSBlock] -= aTz;
                }
                aFx = aFy = aFz = aTx = aTy = aTz = 0.f;
            };
#pragma unroll 1
            for (int p = 0; p < ev.h->n_pairs; ++p) {
                const float* rec = ev.pair + 4 * p;
                const int si = __float_as_int(rec[0]), sj = __float_as_int(rec[1]);
                const float xi = sc.PS[(3 * si) * kSBlock], yi = sc.PS[(3 * si + 1) * kSBlock], zi = sc.PS[(3 * si + 2) * kSBlock];
                const float vx = __fsub_rn(xi, sc.PS[(3 * sj) * kSBlock]);
                const float vy = __fsub_rn(yi, sc.PS[(3 * sj + 1) * kSBlock]);
                const float vz = __fsub_rn(zi, sc.PS[(3 * sj + 2) * kSBlock]);
                float rs;
                const float dist = norm3_clamped(vx, vy, vz, rs);
                occ_self = occ_self || (dist < (hp->use_override ? hp->margin_override : rec[2]));
                c_self = __fadd_rn(c_self, -__fsub_rn(dist, rec[2]));
                if (WANT_GRAD) {
                    const int cols = __float_as_int(rec[3]);
                    const int ci = (cols & 0xff) == 0xff ? -1 : (cols & 0xff), cj = ((cols >> 8) & 0xff) == 0xff ? -1 : ((cols >> 8) & 0xff);
                    if (ci != cj) {            // both ends on one anchor: an internal force, no joint sees it
                        const float inv = hp->w[0] * rs;
                        const float ux = vx * inv, uy = vy * inv, uz = vz * inv;
                        const float tx = yi * uz - zi * uy, ty = zi * ux - xi * uz, tz = xi * uy - yi * ux;
                        if (ci != acol) { flush(); acol = ci; }
                        aFx += ux; aFy += uy; aFz += uz; aTx += tx; aTy += ty; aTz += tz;
                        if (cj >= 0) {
                            float* w = sc.WS + 6 * cj * kSBlock;
                            w[0] += ux; w[kSBlock] += uy; w[2 * kSBlock] += uz; w[3 * kSBlock] += tx; w[4 * kSBlock] += ty; w[5 * kSBlock] += tz;
                        }
                    }
                }
            }
            if (WANT_GRAD) flush();
        }
        // ---------------- outputs ----------------
        {
            const float c_obj = sc.ACC[0], c_ws = sc.ACC[kSBlock];
            const float total = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(hp->w[0], c_self), __fmul_rn(hp->w[1], c_obj)), __fmul_rn(hp->w[2], c_ws)),
                                          __fmul_rn(hp->w[3], c_ee));
            if (cost_total) cost_total[cfg] = total;
            if (cost_terms) reinterpret_cast<float4*>(cost_terms)[cfg] = make_float4(c_self, c_obj, c_ws, c_ee);
            if (occupancy) occupancy[cfg] = (uint8_t)((unsigned)__float_as_int(sc.ACC[2 * kSBlock]) | (occ_self ? TRK_OCC_SELF : 0u));
        }
        // ---------------- reverse mode: leaf -> root over the movable bodies, wrenches from shared memory ----------------
        if constexpr (WANT_GRAD) {
            float F[L][3], Tq[L][3], gq[D > 0 ? D : 1];
            sfor<L>([&](auto ic) {
                constexpr int i = L - 1 - decltype(ic)::value;
                if constexpr (T::pix(i) >= 0 && T::movable(i)) {
                    constexpr int qi = R::qidx[i];
                    constexpr bool child_started = T::has_movable_child(i);
                    const float* w = sc.WS + 6 * qi * kSBlock;
                    sfor<3>([&](auto rc) {
                        F[i][rc] = child_started ? F[i][rc] + w[rc * kSBlock] : w[rc * kSBlock];
                        Tq[i][rc] = child_started ? Tq[i][rc] + w[(3 + rc) * kSBlock] : w[(3 + rc) * kSBlock];
                    });
                    if constexpr (T::anchor(EE) == i) sfor<3>([&](auto rc) { F[i][rc] += eF[rc]; Tq[i][rc] += eT[rc]; });
                    float g;
                    if constexpr (R::kind[i] == TRK_KIND_REVOLUTE) {
                        const float px = PJ[qi][0], py = PJ[qi][1], pz = PJ[qi][2];
                        const float mx = fmaf(pz, F[i][1], fmaf(-py, F[i][2], Tq[i][0]));
                        const float my = fmaf(px, F[i][2], fmaf(-pz, F[i][0], Tq[i][1]));
                        const float mz = fmaf(py, F[i][0], fmaf(-px, F[i][1], Tq[i][2]));
                        g = fmaf(JA[qi][2], mz, fmaf(JA[qi][1], my, JA[qi][0] * mx));
                    } else {
                        g = fmaf(JA[qi][2], F[i][2], fmaf(JA[qi][1], F[i][1], JA[qi][0] * F[i][0]));
                    }
                    gq[qi] = ((pass >> qi) & 1u) ? g : 0.f;
                    constexpr int pa = T::ap(i);
                    if constexpr (T::pix(pa) >= 0) {
                        constexpr bool st = T::has_later_movable_sibling(pa, i);
                        sfor<3>([&](auto rc) {
                            F[pa][rc] = st ? F[pa][rc] + F[i][rc] : F[i][rc];
                            Tq[pa][rc] = st ? Tq[pa][rc] + Tq[i][rc] : Tq[i][rc];
                        });
                    }
                }
            });
            float* grow = grad_q + cfg * D;
            sfor<D>([&](auto c) { grow[c] = gq[c]; });
        }
    }
}

}  // namespace trk
