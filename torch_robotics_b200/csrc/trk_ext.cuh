// Fused objective for a compiled-in robot that carries ATTACHED fixed bodies: grasped-object points
// (environments/objects.py:37-89, robots/robot_panda.py:155-168) and per-link multi-sphere collision models
// (data/configs/*_sphere_config.yaml).  DifferentiableTree.attach_points / attach_link append such bodies to the tree as
// fixed leaves, which used to move the whole robot to the runtime-tree kernels (45-sphere Panda: 4.6e8 configurations/s).
// Here the base robot keeps its unrolled FK (HTree of trk_hot.cuh); every body the cost tables reference - base or
// attached - is described to the kernel as (movable anchor body, constant offset in the anchor's frame), folded on the host
// from the fixed chain above it (trk_cost_create).  At each movable body the kernel
//   * parks the positions of the self-collision-pair bodies anchored there in per-thread shared-memory slots,
//   * evaluates the collision points anchored there through the joint sphere scan of the headline kernel in groups of four
//     (then at most one pair and one single point: no padding lanes; one table load serves the group, FFMA2 keys) and
//     finishes each group from registers in one straight-line block, accumulating the body's wrench,
// in ONE non-inlined device function (the unrolled FK calls it once per movable body: inlined it would be 70 KB of code)
// that reads its uniform scalars from a shared-memory copy of the kernel parameters once per call.
// Pairs run afterwards from the slots, grouped into runs that share a hub body (see ExtHeader); the reverse sweep is the
// compile-time one, reading each body's wrench from shared memory.
// Scene: one equal-radius identity-pose sphere object, workspace box, EE term on the robot's last body, no relu.
#pragma once

#include "trk_hot.cuh"

namespace trk {

#ifndef TRK_EXT_MINBLOCKS
#define TRK_EXT_MINBLOCKS 4
#endif

struct ExtHeader {            // device tables built by trk_cost_create for a model with attached bodies
    int n_pts, n_ps, n_pairs, n_anchor;         // collision points, pair-position slots, pairs, movable bodies + 1 (root)
    int off_pt_start, off_pt, off_ps_start, off_ps, off_run, off_item, n_runs, total_words, pad[4];
};
// words after the header:
//   pt_start  [n_anchor + 1]  points of anchor a (0 = root, 1 + c = the body q column c drives): pt[pt_start[a] .. pt_start[a+1])
//   pt        [n_pts][8]      offset xyz, margin (object) | margin (ws), occupancy margin (object), occupancy margin (ws), point index.
//                             The two occupancy margins are written by the kernel's prologue into its shared-memory copy (the per-call
//                             margin override resolved once per block, not per point)
//   ps_start  [n_anchor + 1], ps [n_ps][4] offset xyz + pad: positions of the self-collision-pair bodies
//   run       [n_runs][4]     hub slot, first item, end item, q columns of the anchors (hub | others << 8, 0xff = root: no joint above)
//                             | first run of its hub << 16 | last run of its hub << 17
//   item      [n_pairs][4]    other slot, margin, occupancy margin (prologue), pad
// A run is a set of pairs that share one body (the hub) and whose other bodies hang on one anchor: every pair force lies on a line
// through the hub, so the run's wrench about the world origin is (U, p_hub x U) with U the sum of the pair forces - three FMAs per
// pair and one cross product per run (grasped box: 56 of the 66 pairs are 14 object points against 4 links = 4 runs).

struct alignas(16) ExtParams {     // kernel parameter; the block's shared memory starts with a copy for ext_at_anchor
    float ws_min[3], w1;           // w1 = weight of the object term, w2 = of the workspace term
    float ws_max[3], w2;
    float tol, c2max; unsigned lowmask; int n_pairs2;
    int o_tq, o_tsph, o_pt_start, o_pt;          // word offsets into the dynamic shared memory
    int o_ps_start, o_ps, o_scratch, n_ps;
    int n_sph, D, pad0, pad1;
    // kernel body only
    float w0, w3, ee_w_pos, ee_w_rot;
    float ee_R[9], ee_t[3];
    int ee_square, ee_enabled, use_override; float margin_override;
    int off_sphq, off_sph, ext_nwords, o_ew;
    int o_run, o_item, n_runs, n_points;
    // general scene (scene == 1: the tables of general_scene_eval, trk_hot.cuh) instead of the single sphere object
    int scene, n_obj, n_box, n_sph_all;
    int n_sphq_all, off_box, off_sph_all, off_sphq_all;
    int off_obj_box_start, off_obj_ident, off_obj_R, off_obj_pos;
    int off_obj_start, off_obj_qstart, off_obj_tol, off_obj_c2max;
    int o_gbox, grid_n[3];
    float grid_lo[3], grid_extent[3], grid_nf[3], pad2[3];
    const float4* grid_data; long long pad3;
};
constexpr int kExtParamWords = (int)(sizeof(ExtParams) / 4);

// W collision points (records rec[0 .. 2 W)) of a body whose world pose is (Rw, pw): joint sphere scan (one table load serves the
// group, FFMA2 keys as in the headline kernel), then each point finished from registers in one straight-line block.
struct ExtAcc { float Fx, Fy, Fz, Tx, Ty, Tz, c_obj, c_ws; bool occ_obj, occ_ws; };
struct ExtConsts {
    float4 wmin, wmax;                 // workspace box; .w = weight of the object term / of the workspace term
    float tol, tol_c2; unsigned lowmask; int n_pairs2;
    const float4* tq; const float4* tsph;
    GeneralScene GS;                   // scene == 1 only
};
template <int W, bool GRAD, bool ARGMIN, int SCENE>
static __device__ __forceinline__ void ext_group(float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7, float r8,
                                                 float px, float py, float pz, const float4* __restrict__ rec, const ExtConsts& K,
                                                 const ExtParams* __restrict__ ep, int* __restrict__ argmin_row, ExtAcc& A) {
    const float kInf = __int_as_float(0x7f800000);
    // workspace box of one point (distance_fields.py:319-332) + the point's force (gx, gy, gz: the object term's share) into the wrench
    auto ws_and_wrench = [&](const float4 mr, float x, float y, float z, float gx, float gy, float gz) {
        const float e0 = __fsub_rn(x, K.wmin.x), e1 = __fsub_rn(y, K.wmin.y), e2 = __fsub_rn(z, K.wmin.z);
        const float f0 = __fsub_rn(K.wmax.x, x), f1 = __fsub_rn(K.wmax.y, y), f2 = __fsub_rn(K.wmax.z, z);
        const float emin = fmin3(e0, e1, e2), fmn = fmin3(f0, f1, f2);
        const float dmin = fminf(emin, fmn);
        A.occ_ws = A.occ_ws || (dmin < mr.z);
        A.c_ws = __fadd_rn(A.c_ws, -__fsub_rn(dmin, mr.x));
        if constexpr (GRAD) {
            const bool upper = fmn < emin;                       // ties go to the lower index: the x - min side
            float gg = upper ? K.wmax.w : -K.wmax.w;
            gg = (dmin != 0.f) ? gg : 0.f;                       // sign(d) |d| has zero gradient at d == 0
            const float a0 = upper ? f0 : e0, a1 = upper ? f1 : e1;
            const bool p0 = a0 == dmin, p1 = a1 == dmin;
            gx += p0 ? gg : 0.f;
            gy += (!p0 && p1) ? gg : 0.f;
            gz += (!p0 && !p1) ? gg : 0.f;
            A.Fx += gx; A.Fy += gy; A.Fz += gz;
            A.Tx = fmaf(y, gz, fmaf(-z, gy, A.Tx));
            A.Ty = fmaf(z, gx, fmaf(-x, gz, A.Ty));
            A.Tz = fmaf(x, gy, fmaf(-y, gx, A.Tz));
        }
    };
    if constexpr (SCENE == 1) {
        float X[W][3], mg[W], mog[W], vmax[W], ug[W][3];
        int pidx[W];
#pragma unroll
        for (int p = 0; p < W; ++p) {
            const float4 t = rec[2 * p], mr = rec[2 * p + 1];
            X[p][0] = fmaf(r2, t.z, fmaf(r1, t.y, fmaf(r0, t.x, px)));
            X[p][1] = fmaf(r5, t.z, fmaf(r4, t.y, fmaf(r3, t.x, py)));
            X[p][2] = fmaf(r8, t.z, fmaf(r7, t.y, fmaf(r6, t.x, pz)));
            mg[p] = t.w; mog[p] = mr.y; pidx[p] = __float_as_int(mr.w);
        }
        const int n_points = ep->n_points;
        general_scene_eval<W, GRAD>(K.GS, X, mg, mog, vmax, A.occ_obj, ug, [&](int o, int p, int idx) {
            if constexpr (ARGMIN) argmin_row[o * n_points + pidx[p]] = idx;
        });
#pragma unroll
        for (int p = 0; p < W; ++p) {
            A.c_obj = __fadd_rn(A.c_obj, vmax[p]);
            float gx = 0.f, gy = 0.f, gz = 0.f;
            if constexpr (GRAD) { gx = -(K.wmin.w * ug[p][0]); gy = -(K.wmin.w * ug[p][1]); gz = -(K.wmin.w * ug[p][2]); }
            ws_and_wrench(rec[2 * p + 1], X[p][0], X[p][1], X[p][2], gx, gy, gz);
        }
        return;
    }
    const unsigned keep = ~K.lowmask;
    float X[W][3], m1[W], m2[W];
#pragma unroll
    for (int p = 0; p < W; ++p) {
        const float4 t = rec[2 * p];
        X[p][0] = fmaf(r2, t.z, fmaf(r1, t.y, fmaf(r0, t.x, px)));
        X[p][1] = fmaf(r5, t.z, fmaf(r4, t.y, fmaf(r3, t.x, py)));
        X[p][2] = fmaf(r8, t.z, fmaf(r7, t.y, fmaf(r6, t.x, pz)));
        m1[p] = kInf; m2[p] = kInf;
    }
    unsigned idx = 0u;
    float4 t0 = K.tq[0], t1 = K.tq[1];
#pragma unroll 1
    for (int j = 0; j < K.n_pairs2; ++j, idx += 2u) {
        const float2 cx = make_float2(t0.x, t0.y), cy = make_float2(t0.z, t0.w), cz = make_float2(t1.x, t1.y), cw = make_float2(t1.z, t1.w);
        const int jn = (j + 1 < K.n_pairs2) ? j + 1 : j;          // the next pair's table entry is in flight during this one
        t0 = K.tq[2 * jn]; t1 = K.tq[2 * jn + 1];
#pragma unroll
        for (int p = 0; p < W; ++p) {
            const float2 k2 = __ffma2_rn(make_float2(X[p][0], X[p][0]), cx,
                                         __ffma2_rn(make_float2(X[p][1], X[p][1]), cy, __ffma2_rn(make_float2(X[p][2], X[p][2]), cz, cw)));
            const float pa = __uint_as_float((__float_as_uint(k2.x) & keep) | idx);
            const float pb = __uint_as_float((__float_as_uint(k2.y) & keep) | (idx + 1u));
            const float lo = fminf(pa, pb), hi = fmaxf(pa, pb);
            m2[p] = fmin3(fmaxf(m1[p], lo), m2[p], hi);
            m1[p] = fminf(m1[p], lo);
        }
    }
    // winners; a near tie (or NaN) on some point: exact scan of that point, first index wins (rare)
    int widx[W];
    bool all_certain = true;
#pragma unroll
    for (int p = 0; p < W; ++p) {
        const float xx = __fmaf_rn(X[p][2], X[p][2], __fmaf_rn(X[p][1], X[p][1], __fmul_rn(X[p][0], X[p][0])));
        all_certain = all_certain && ((m2[p] - m1[p]) > __fmaf_rn(K.tol, xx, K.tol_c2));
        widx[p] = (int)(__float_as_uint(m1[p]) & K.lowmask);
    }
    if (!all_certain) {
        const int n_sph = ep->n_sph;
#pragma unroll
        for (int p = 0; p < W; ++p) {
            const float x = X[p][0], y = X[p][1], z = X[p][2];
            const float xx = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
            if (!((m2[p] - m1[p]) > __fmaf_rn(K.tol, xx, K.tol_c2))) {
                float best = kInf;
                int bidx = 0;
#pragma unroll 1
                for (int kk = 0; kk < n_sph; ++kk) {
                    const float4 c = K.tsph[kk];
                    const float sd = __fsub_rn(norm3_rn(__fsub_rn(x, c.x), __fsub_rn(y, c.y), __fsub_rn(z, c.z)), c.w);
                    if (sd < best) { best = sd; bidx = kk; }
                }
                widx[p] = bidx;
            }
        }
    }
    // exact evaluation of each point's winner + workspace box (distance_fields.py:319-332)
#pragma unroll
    for (int p = 0; p < W; ++p) {
        const float4 mr = rec[2 * p + 1];                 // margin (ws), occupancy margins (object, ws), point index
        const float m_obj = rec[2 * p].w;
        const float x = X[p][0], y = X[p][1], z = X[p][2];
        const float4 c = K.tsph[widx[p]];
        const float vx = __fsub_rn(x, c.x), vy = __fsub_rn(y, c.y), vz = __fsub_rn(z, c.z);
        float rs;
        const float best = __fsub_rn(norm3_clamped(vx, vy, vz, rs), c.w);
        if constexpr (ARGMIN) argmin_row[__float_as_int(mr.w)] = widx[p];
        A.occ_obj = A.occ_obj || (best < mr.y);
        A.c_obj = __fadd_rn(A.c_obj, -__fsub_rn(best, m_obj));
        const float w1 = K.wmin.w * rs;
        ws_and_wrench(mr, x, y, z, -(w1 * vx), -(w1 * vy), -(w1 * vz));
    }
}

// points [l, l1) of one anchor: quads, then at most one pair and one single point - no padding lanes
template <bool GRAD, bool ARGMIN, int SCENE>
static __device__ __forceinline__ void ext_points(float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7, float r8,
                                                  float px, float py, float pz, int l, int l1, const ExtParams* __restrict__ ep,
                                                  int* __restrict__ argmin_row, ExtAcc& A) {
    extern __shared__ __align__(16) float esm[];
    ExtConsts K;
    K.wmin = *reinterpret_cast<const float4*>(ep->ws_min); K.wmax = *reinterpret_cast<const float4*>(ep->ws_max);
    K.tol = ep->tol; K.tol_c2 = ep->tol * ep->c2max;
    K.lowmask = ep->lowmask; K.n_pairs2 = ep->n_pairs2;
    K.tq = reinterpret_cast<const float4*>(esm + ep->o_tq);
    K.tsph = reinterpret_cast<const float4*>(esm + ep->o_tsph);
    if constexpr (SCENE == 1) {
        K.GS.box = reinterpret_cast<const float4*>(esm + ep->o_gbox);
        K.GS.sph = K.GS.box + 2 * ep->n_box; K.GS.sphq = K.GS.sph + ep->n_sph_all; K.GS.obj = K.GS.sphq + ep->n_sphq_all;
        K.GS.n_obj = ep->n_obj; K.GS.grid_data = ep->grid_data;
#pragma unroll
        for (int i = 0; i < 3; ++i) { K.GS.grid_n[i] = ep->grid_n[i]; K.GS.grid_lo[i] = ep->grid_lo[i]; K.GS.grid_extent[i] = ep->grid_extent[i]; K.GS.grid_nf[i] = ep->grid_nf[i]; }
    }
    const float4* pt = reinterpret_cast<const float4*>(esm + ep->o_pt);
    if constexpr (SCENE == 0)          // (the general evaluation holds more per point: pairs there)
#pragma unroll 1
    for (; l + 4 <= l1; l += 4) ext_group<4, GRAD, ARGMIN, SCENE>(r0, r1, r2, r3, r4, r5, r6, r7, r8, px, py, pz, pt + 2 * l, K, ep, argmin_row, A);
#pragma unroll 1
    for (; l + 2 <= l1; l += 2) ext_group<2, GRAD, ARGMIN, SCENE>(r0, r1, r2, r3, r4, r5, r6, r7, r8, px, py, pz, pt + 2 * l, K, ep, argmin_row, A);
    if (l < l1) ext_group<1, GRAD, ARGMIN, SCENE>(r0, r1, r2, r3, r4, r5, r6, r7, r8, px, py, pz, pt + 2 * l, K, ep, argmin_row, A);
}

// Everything anchored at one movable body (or the root, col = -1) whose world pose is (Rw, pw).  Per-thread scratch in shared
// memory, [slot][thread] layout (stride kSBlock): PS [3 n_ps] positions of the pair bodies | WS [6 D] wrench (F, T about the world
// origin) of the bodies anchored at each movable body | ACC [3] c_obj, c_ws, occupancy bits.  One non-inlined function: the unrolled
// FK calls it once per movable body (inlined it would be 70 KB of straight-line code).
template <int SCENE>
static __device__ __noinline__ void ext_at_anchor(float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7, float r8,
                                                  float px, float py, float pz, int anchor, int col, int* __restrict__ argmin_row, bool want_grad) {
    extern __shared__ __align__(16) float esm[];
    const ExtParams* ep = reinterpret_cast<const ExtParams*>(esm);
    const int* ism = reinterpret_cast<const int*>(esm);
    float* PS = esm + ep->o_scratch + threadIdx.x;
    // (1) pair bodies anchored here
    {
        const int k1 = ism[ep->o_ps_start + anchor + 1];
        const float4* ps = reinterpret_cast<const float4*>(esm + ep->o_ps);
#pragma unroll 1
        for (int k = ism[ep->o_ps_start + anchor]; k < k1; ++k) {
            const float4 t = ps[k];
            PS[(3 * k + 0) * kSBlock] = fmaf(r2, t.z, fmaf(r1, t.y, fmaf(r0, t.x, px)));
            PS[(3 * k + 1) * kSBlock] = fmaf(r5, t.z, fmaf(r4, t.y, fmaf(r3, t.x, py)));
            PS[(3 * k + 2) * kSBlock] = fmaf(r8, t.z, fmaf(r7, t.y, fmaf(r6, t.x, pz)));
        }
    }
    // (2) collision points anchored here
    const int g0 = ism[ep->o_pt_start + anchor], g1 = ism[ep->o_pt_start + anchor + 1];
    ExtAcc A{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, false, false};
    if (g0 < g1) {
#define TRK_EXT_POINTS(G_, A_, S_) ext_points<G_, A_, S_>(r0, r1, r2, r3, r4, r5, r6, r7, r8, px, py, pz, g0, g1, ep, argmin_row, A)
        if (want_grad) { if (argmin_row) TRK_EXT_POINTS(true, true, SCENE); else TRK_EXT_POINTS(true, false, SCENE); }
        else           { if (argmin_row) TRK_EXT_POINTS(false, true, SCENE); else TRK_EXT_POINTS(false, false, SCENE); }
#undef TRK_EXT_POINTS
    }
    float* WS = PS + 3 * ep->n_ps * kSBlock;
    if (col >= 0) {
        float* w = WS + 6 * col * kSBlock;
        w[0] = A.Fx; w[kSBlock] = A.Fy; w[2 * kSBlock] = A.Fz; w[3 * kSBlock] = A.Tx; w[4 * kSBlock] = A.Ty; w[5 * kSBlock] = A.Tz;
    }
    if (g0 < g1) {
        float* ACC = WS + 6 * ep->D * kSBlock;
        ACC[0] += A.c_obj;
        ACC[kSBlock] += A.c_ws;
        if (A.occ_obj || A.occ_ws)
            ACC[2 * kSBlock] = __int_as_float(__float_as_int(ACC[2 * kSBlock]) | (A.occ_obj ? (int)TRK_OCC_OBJECT : 0) | (A.occ_ws ? (int)TRK_OCC_WS : 0));
    }
}

// SCENE: 0 = one equal-radius identity-pose sphere object (lean scan), 1 = general scene (general_scene_eval of trk_hot.cuh); separate
// kernels and separate anchor functions, so that the general evaluation's register needs do not reach the sphere scene
template <class R, int EE, bool WANT_GRAD, int SCENE = 0>
__global__ void __launch_bounds__(kSBlock, TRK_EXT_MINBLOCKS) k_objective_ext(const __grid_constant__ ExtParams ep, const float* __restrict__ cost_words,
                                                           const float* __restrict__ ext_words_g, const float* __restrict__ q, long long n,
                                                           float* __restrict__ cost_total, float* __restrict__ cost_terms,
                                                           float* __restrict__ grad_q, uint8_t* __restrict__ occupancy, int* __restrict__ argmin) {
    using T = HTree<R>;
    constexpr int L = R::L, D = R::D;
    extern __shared__ __align__(16) float esm[];
    // ExtParams copy | ext words | scan table | sphere table ; then the per-thread scratch columns
    const int tid = threadIdx.x;
    {
        const float* psrc = reinterpret_cast<const float*>(&ep);
        for (int i = tid; i < kExtParamWords; i += kSBlock) esm[i] = psrc[i];
        float* ew = esm + ep.o_ew;
        for (int i = tid; i < ep.ext_nwords; i += kSBlock) ew[i] = __ldg(ext_words_g + i);
        if constexpr (SCENE == 1) {          // the tables of general_scene_eval
            float4* gbox = reinterpret_cast<float4*>(esm + ep.o_gbox);
            const float4* box = reinterpret_cast<const float4*>(cost_words + ep.off_box);
            for (int i = tid; i < 2 * ep.n_box; i += kSBlock) gbox[i] = __ldg(box + i);
            float4* gsph = gbox + 2 * ep.n_box;
            const float4* sph = reinterpret_cast<const float4*>(cost_words + ep.off_sph_all);
            for (int i = tid; i < ep.n_sph_all; i += kSBlock) gsph[i] = __ldg(sph + i);
            float* gq = reinterpret_cast<float*>(gsph + ep.n_sph_all);
            const float* sphq = cost_words + ep.off_sphq_all;
            for (int i = tid; i < ep.n_sphq_all * 4; i += kSBlock) {
                const int k = i >> 2, comp = i & 3;
                gq[(k >> 1) * 8 + comp * 2 + (k & 1)] = __ldg(sphq + i);
            }
            float4* tobj = reinterpret_cast<float4*>(gq) + ep.n_sphq_all;
            for (int o = tid; o < ep.n_obj; o += kSBlock)
                general_scene_object_record(tobj + 6 * o, cost_words, o, ep.off_obj_R, ep.off_obj_pos, ep.off_obj_ident, ep.off_obj_start,
                                            ep.off_obj_qstart, ep.off_obj_box_start, ep.off_obj_tol, ep.off_obj_c2max);
        } else {
            const float* sphq = cost_words + ep.off_sphq;
            float* tqw = esm + ep.o_tq;
            for (int i = tid; i < ep.n_pairs2 * 8; i += kSBlock) {      // two spheres interleaved: (ax, bx, ay, by) (az, bz, aw, bw)
                const int k = i >> 2, comp = i & 3;
                tqw[(k >> 1) * 8 + comp * 2 + (k & 1)] = __ldg(sphq + i);
            }
            const float4* sph = reinterpret_cast<const float4*>(cost_words + ep.off_sph);
            float4* tsph = reinterpret_cast<float4*>(esm + ep.o_tsph);
            for (int i = tid; i < ep.n_sph; i += kSBlock) tsph[i] = __ldg(sph + i);
        }
    }
    __syncthreads();
    const int4* runs = reinterpret_cast<const int4*>(esm + ep.o_run);
    float4* items = reinterpret_cast<float4*>(esm + ep.o_item);
    {   // occupancy margins: the call's override, or each record's own margin
        const ExtHeader* eh = reinterpret_cast<const ExtHeader*>(esm + ep.o_ew);
        float* pt = esm + ep.o_pt;
        for (int i = tid; i < eh->n_pts; i += kSBlock) {
            pt[8 * i + 5] = ep.use_override ? ep.margin_override : pt[8 * i + 3];
            pt[8 * i + 6] = ep.use_override ? ep.margin_override : pt[8 * i + 4];
        }
        for (int i = tid; i < eh->n_pairs; i += kSBlock) items[i].z = ep.use_override ? ep.margin_override : items[i].y;
    }
    __syncthreads();
    float* PS = esm + ep.o_scratch + tid;
    float* WS = PS + 3 * ep.n_ps * kSBlock;
    float* ACC = WS + 6 * D * kSBlock;
    for (long long cfg = (long long)blockIdx.x * kSBlock + tid; cfg < n; cfg += (long long)gridDim.x * kSBlock) {
        const float* qrow = q + cfg * D;
        float qv[D];
        sfor<D>([&](auto c) { qv[c] = __ldg(qrow + c); });
        ACC[0] = 0.f; ACC[kSBlock] = 0.f; ACC[2 * kSBlock] = 0.f;
        int* amrow = argmin ? argmin + cfg * (long long)(ep.n_points * (SCENE == 1 ? ep.n_obj : 1)) : nullptr;
        // ---------------- forward: unrolled FK; everything anchored at a body is evaluated while its pose is in registers ----------------
        float W[L][12], JA[D > 0 ? D : 1][3], PJ[D > 0 ? D : 1][3], Ree[9], pee[3];
        unsigned pass = 0u;
        ext_at_anchor<SCENE>(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0, -1, amrow, false);     // root-anchored bodies
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
                ext_at_anchor<SCENE>(W[i][0], W[i][1], W[i][2], W[i][3], W[i][4], W[i][5], W[i][6], W[i][7], W[i][8], W[i][9], W[i][10], W[i][11],
                              1 + qi, qi, amrow, WANT_GRAD);
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
            const float* Rt = ep.ee_R;
            float dg[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                dg[i] = __fmaf_rn(Ree[i * 3 + 2], Rt[i * 3 + 2], __fmaf_rn(Ree[i * 3 + 1], Rt[i * 3 + 1], __fmul_rn(Ree[i * 3 + 0], Rt[i * 3 + 0])));
            const float tr = __fadd_rn(__fadd_rn(dg[0], dg[1]), dg[2]);
            const float phi = __fmul_rn(__fsub_rn(tr, 1.f), 0.5f);
            float Dd = __fmul_rn(ep.ee_w_rot, __fsub_rn(1.f, phi));
            const float vx = __fsub_rn(pee[0], ep.ee_t[0]), vy = __fsub_rn(pee[1], ep.ee_t[1]), vz = __fsub_rn(pee[2], ep.ee_t[2]);
            float rs;
            const float dist = norm3_clamped(vx, vy, vz, rs);
            Dd = __fadd_rn(Dd, __fmul_rn(ep.ee_w_pos, dist));
            c_ee = ep.ee_enabled ? (ep.ee_square ? __fmul_rn(Dd, Dd) : Dd) : 0.f;
            const float scale = ep.ee_enabled ? ep.w3 * (ep.ee_square ? 2.f * Dd : 1.f) : 0.f;
            const float kr = scale * (-0.5f * ep.ee_w_rot);
            const float kp = scale * ep.ee_w_pos * rs;
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
        // ---------------- self-collision pairs, run by run (distance_fields.py:194-215) ----------------
        float c_self = 0.f;
        bool occ_self = false;
        {
            float hx = 0.f, hy = 0.f, hz = 0.f, Hx = 0.f, Hy = 0.f, Hz = 0.f;     // hub position, sum of the forces on the hub over its runs
#pragma unroll 1
            for (int r = 0; r < ep.n_runs; ++r) {
                const int4 hd = runs[r];
                if (hd.w & 0x10000) {                                             // first run of this hub
                    hx = PS[(3 * hd.x) * kSBlock]; hy = PS[(3 * hd.x + 1) * kSBlock]; hz = PS[(3 * hd.x + 2) * kSBlock];
                    Hx = Hy = Hz = 0.f;
                }
                float Ux = 0.f, Uy = 0.f, Uz = 0.f;
#pragma unroll 1
                for (int k = hd.y; k < hd.z; ++k) {
                    const float4 it = items[k];
                    const int so = __float_as_int(it.x);
                    const float vx = __fsub_rn(PS[(3 * so) * kSBlock], hx), vy = __fsub_rn(PS[(3 * so + 1) * kSBlock], hy),
                                vz = __fsub_rn(PS[(3 * so + 2) * kSBlock], hz);
                    float rs;
                    const float dist = norm3_clamped(vx, vy, vz, rs);
                    occ_self = occ_self || (dist < it.z);
                    c_self = __fadd_rn(c_self, -__fsub_rn(dist, it.y));
                    if constexpr (WANT_GRAD) {
                        const float inv = ep.w0 * rs;          // force on the hub: + v / dist (v = other - hub), on the other body: the opposite
                        Ux = fmaf(vx, inv, Ux); Uy = fmaf(vy, inv, Uy); Uz = fmaf(vz, inv, Uz);
                    }
                }
                if constexpr (WANT_GRAD) {
                    const int ch = hd.w & 0xff, co = (hd.w >> 8) & 0xff;
                    if (ch != co) {            // both ends on one anchor: internal forces, no joint sees them
                        Hx += Ux; Hy += Uy; Hz += Uz;
                        if (co != 0xff) {      // every force of the run lies on a line through the hub: torque about the world origin = p_hub x U
                            float* w = WS + 6 * co * kSBlock;
                            w[0] -= Ux; w[kSBlock] -= Uy; w[2 * kSBlock] -= Uz;
                            w[3 * kSBlock] -= hy * Uz - hz * Uy; w[4 * kSBlock] -= hz * Ux - hx * Uz; w[5 * kSBlock] -= hx * Uy - hy * Ux;
                        }
                    }
                    if ((hd.w & 0x20000) && ch != 0xff) {                          // last run of this hub
                        float* w = WS + 6 * ch * kSBlock;
                        w[0] += Hx; w[kSBlock] += Hy; w[2 * kSBlock] += Hz;
                        w[3 * kSBlock] += hy * Hz - hz * Hy; w[4 * kSBlock] += hz * Hx - hx * Hz; w[5 * kSBlock] += hx * Hy - hy * Hx;
                    }
                }
            }
        }
        // ---------------- outputs ----------------
        {
            const float c_obj = ACC[0], c_ws = ACC[kSBlock];
            const float total = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ep.w0, c_self), __fmul_rn(ep.w1, c_obj)), __fmul_rn(ep.w2, c_ws)),
                                          __fmul_rn(ep.w3, c_ee));
            if (cost_total) cost_total[cfg] = total;
            if (cost_terms) reinterpret_cast<float4*>(cost_terms)[cfg] = make_float4(c_self, c_obj, c_ws, c_ee);
            if (occupancy) occupancy[cfg] = (uint8_t)((unsigned)__float_as_int(ACC[2 * kSBlock]) | (occ_self ? TRK_OCC_SELF : 0u));
        }
        // ---------------- reverse mode: leaf -> root over the movable bodies, wrenches from shared memory ----------------
        if constexpr (WANT_GRAD) {
            float F[L][3], Tq[L][3], gq[D > 0 ? D : 1];
            sfor<L>([&](auto ic) {
                constexpr int i = L - 1 - decltype(ic)::value;
                if constexpr (T::pix(i) >= 0 && T::movable(i)) {
                    constexpr int qi = R::qidx[i];
                    constexpr bool child_started = T::has_movable_child(i);
                    const float* w = WS + 6 * qi * kSBlock;
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
