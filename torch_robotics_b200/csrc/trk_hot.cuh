// Headline kernel of the compiled-in family: fused objective (FK + self pairs + sphere scene + workspace box + EE SE(3)
// term + reverse mode) for a robot whose collision model is known at compile time, in the common planning scene
// (ONE equal-radius identity-pose sphere object, workspace box on, EE term on the model's own EE body, no relu).
//
// The kernel is issue-bound (one warp instruction per clock per SM sub-partition whatever the pipe,
// profiles/r01_microbench_fp32.md), so everything here is about instructions per configuration and about keeping the
// straight-line body inside the 32 KB instruction cache level:
//   * fixed bodies are never composed: at compile time every fixed chain is folded into one constant transform relative
//     to its nearest movable ancestor (HTree::origin); a fixed body costs three FMAs for its position, and only when the
//     collision model or the EE term needs it;
//   * joint clamp = max.NaN / min.NaN (torch.clamp semantics incl. NaN) + one compare for the gradient mask;
//   * sin/cos: two-constant Cody-Waite with a magic-number round (no conversion instructions, no large-argument path)
//     for joints whose limits bound the argument;
//   * every uniform scalar (margins, weights, box, EE target) is a kernel parameter, i.e. a constant-bank operand of the
//     arithmetic instruction itself: no table loads, no `use_override` selects (the host resolves the margins);
//   * sphere scan: two spheres per packed fma.rn.f32x2 (FFMA2: one issue slot for two FMAs; IEEE per lane, so keys
//     and decisions are bit-identical to the scalar scan);
//   * reverse mode: one running wrench per open branch, first writes are assignments (no zero-initialised accumulators),
//     cross products fused into the accumulation, bodies whose position cannot depend on any joint receive no force.
// Decision arithmetic (distances that feed `<`, arg-min, arg-max) is the canonical sequence of trk_device.cuh.
#pragma once

#include "trk_device.cuh"

namespace trk {

#ifndef TRK_HOT_MINBLOCKS
#define TRK_HOT_MINBLOCKS 4
#endif
#ifndef TRK_HOT_QSTAGE      // 0: direct loads, 1: next configuration's q prefetched into registers, 2: cp.async block tile + barrier,
                            // 3: cp.async per-warp tile (no block barrier)
#define TRK_HOT_QSTAGE 1
#endif

#ifndef TRK_HOT_SCAN_UNROLL   // sphere pairs per trip of the scan loop
#define TRK_HOT_SCAN_UNROLL 2
#endif
constexpr int kHotScanUnroll = TRK_HOT_SCAN_UNROLL;

constexpr int kHotMaxPairs = 16, kHotMaxPoints = 16;

struct HotParams {
    float m_pair[kHotMaxPairs], mo_pair[kHotMaxPairs];   // self pairs: cost margin / occupancy margin
    float m_obj[kHotMaxPoints], mo_obj[kHotMaxPoints];   // collision points vs the sphere object
    float m_ws[kHotMaxPoints], mo_ws[kHotMaxPoints];     // collision points vs the workspace box
    float w[4];
    float ws_min[3], ws_max[3];
    float ee_R[9], ee_t[3];
    float ee_w_pos, ee_w_rot;
    int   ee_square;
    float tol, c2max;          // gap test of the scan
    unsigned lowmask;          // index bits packed into the scan keys
    int   n_pairs2;            // padded spheres / 2
    int   n_sph;
    int   off_sphq, off_sph;   // word offsets into the cost buffer
    // collision points whose position is a compile-time constant (HTree::pos_is_const): evaluated once on the host with
    // the same arithmetic (hot_params_from_cost)
    float c_obj_const, c_ws_const;
    unsigned occ_const;
    int   widx_const[kHotMaxPoints];
    // k_objective_ext only (its margins live in device tables): the occupancy tests use this margin instead when set
    int   use_override;
    float margin_override;
    int   ee_enabled;          // 0 = the cost table has no EE term (PlanningTask.compute_collision_cost): c_ee = 0, no EE wrench
    // general scenes (SCENE == 1: up to kHotMaxObjects posed objects of spheres and / or boxes and rounded boxes, e.g.
    // EnvTableShelf, EnvMazeBoxes3D, EnvSpheres3DExtraObjects): offsets of the general tables in the cost buffer; arg-min of the
    // constant points per object
    int   n_obj, n_box, off_box, off_obj_box_start, off_obj_ident, off_obj_R, off_obj_pos;
    int   n_sph_all, n_sphq_all, off_sph_all, off_sphq_all, off_obj_start, off_obj_qstart, off_obj_tol, off_obj_c2max;
    int   widx_const_obj[4][kHotMaxPoints];
    // voxel-SDF object of a general scene (grid_map_sdf.py; identity pose, owns no primitives): cell = clamp(floor((x - lo) /
    // extent * nf)), one float4 (sdf, gradient) per cell in HBM.  The host cannot read the grid, so the constant points are
    // evaluated by the kernel there: vconst[s] = their value over the other objects (-inf when there are none)
    int   grid_obj, grid_n[3];           // grid_obj < 0: no voxel grid
    float grid_lo[3], grid_extent[3], grid_nf[3];
    const float4* grid_data;
    float vconst[kHotMaxPoints];
};
constexpr int kHotMaxObjects = 4;

__device__ __forceinline__ float fmax_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// sin/cos for |x| <= 64 (k = round(x 2/pi) < 2^6: the products k*c1, k*c2 need no third constant): same kernels and
// the same quadrant logic as trk_sincos; the round is a magic-number add, so no FRND / F2I.
__device__ __forceinline__ void trk_sincos_small(float x, float& s, float& c) {
    const float km = fmaf(x, 0.63661977236758134308f, 12582912.f);
    const int k = __float_as_int(km);
    const float kf = km - 12582912.f;
    float r = fmaf(kf, -0x1.921fb6p+0f, x);
    r = fmaf(kf, 0x1.777a5cp-25f, r);
    const float r2 = r * r;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    sp = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    cp = fmaf(cp, r2, -0.5f);
    cp = fmaf(cp, r2, 1.0f);
    const float ss = (k & 1) ? cp : sp, cc = (k & 1) ? sp : cp;
    s = __int_as_float(__float_as_int(ss) ^ ((k << 30) & 0x80000000));
    c = __int_as_float(__float_as_int(cc) ^ (((k << 30) ^ (k << 31)) & 0x80000000));
}

// sqrt_rn_core with the zero guard folded into a clamp of the reciprocal-root estimate: rs <= 2^60 (the estimate of any
// s >= 2^-100 is below 2^51), so s == 0 gives g = 0 * 2^60 = 0 without a select, and direction vectors v * (w * rs)
// vanish with v instead of becoming 0 * inf.  Same result as sqrt_rn_core for s == 0 and 2^-100 <= s < 2^126.
__device__ __forceinline__ float sqrt_rn_clamped(float s, float& rs) {
    rs = fminf(mufu_rsq(s), 0x1p60f);
    float g = __fmul_rn(s, rs);
    const float hh = __fmul_rn(rs, 0.5f);
    const float r = __fmaf_rn(-g, g, s);
    return __fmaf_rn(r, hh, g);
}
__device__ __forceinline__ float norm3_clamped(float vx, float vy, float vz, float& rs) {
    float s = __fmul_rn(vx, vx);
    s = __fmaf_rn(vy, vy, s);
    s = __fmaf_rn(vz, vz, s);
    return sqrt_rn_clamped(s, rs);
}

// acc (+)= w * c for an immediate c; unit coefficients are taken first by the callers so that the remaining terms fuse
__device__ __forceinline__ void hterm(float& acc, bool& started, float w, float c) {
    if (c == 0.f) return;
    if (!started) {
        started = true;
        acc = (c == 1.f) ? w : ((c == -1.f) ? -w : w * c);
        return;
    }
    acc = (c == 1.f) ? acc + w : ((c == -1.f) ? acc - w : fmaf(w, c, acc));
}
// sum_k w[k] * c[k] (+ init when has_init): +-1 coefficients first (they are moves / adds), then the rest as FMAs
__device__ __forceinline__ float hdot3(float w0, float w1, float w2, float c0, float c1, float c2, bool has_init, float init) {
    float acc = init;
    bool st = has_init;
    if (c0 == 1.f || c0 == -1.f) hterm(acc, st, w0, c0);
    if (c1 == 1.f || c1 == -1.f) hterm(acc, st, w1, c1);
    if (c2 == 1.f || c2 == -1.f) hterm(acc, st, w2, c2);
    if (!(c0 == 1.f || c0 == -1.f)) hterm(acc, st, w0, c0);
    if (!(c1 == 1.f || c1 == -1.f)) hterm(acc, st, w1, c1);
    if (!(c2 == 1.f || c2 == -1.f)) hterm(acc, st, w2, c2);
    return st ? acc : 0.f;
}

// ------------------------------------------------------------------------------------------
// General scene, shared by k_objective_hot (SCENE 1) and k_objective_ext: up to kHotMaxObjects objects - posed sets of
// spheres and / or boxes and rounded boxes, and the voxel-SDF object - against NP points at once (primitives.py:135-139
// spheres, :219-223 / :326-334 boxes, :397-398 object pose, grid_map_sdf.py:96-112 voxel lookup; distance_fields.py:
// value = margin - sdf, max over objects, the first maximal object keeps the gradient).  Tables in shared memory:
//   box [2 n_box] (c.xyz, radius | half.xyz, rounded flag) | sph [n_sph] (c, r) | sphq [2 pairs] scan table, two spheres
//   interleaved | obj [6 n_obj]: (pos.xyz, identity flag) (R rows; .w: box range, first sphere) (sphere count, first scan pair,
//   scan pairs, index mask) (tol, c2max, uniform-radius flag, voxel-grid flag)
// Per object the points are taken to the object frame once; an equal-radius sphere set goes through the FFMA2 key scan (a near
// tie or NaN on any point sends the object to the exact loop), unequal radii are evaluated exactly; every box exactly, all
// points per step of a rolled loop (the next box in flight); strict <: the earlier primitive wins ties.
// ------------------------------------------------------------------------------------------
struct GeneralScene {
    const float4* box; const float4* sph; const float4* sphq; const float4* obj;
    int n_obj;
    const float4* grid_data;
    int grid_n[3];
    float grid_lo[3], grid_extent[3], grid_nf[3];
};

__device__ __forceinline__ int general_grid_lookup(const GeneralScene& G, float x, float y, float z, float& sd) {
    const int i0 = grid_cell(x, G.grid_lo[0], G.grid_extent[0], G.grid_nf[0], G.grid_n[0]);
    const int i1 = grid_cell(y, G.grid_lo[1], G.grid_extent[1], G.grid_nf[1], G.grid_n[1]);
    const int i2 = grid_cell(z, G.grid_lo[2], G.grid_extent[2], G.grid_nf[2], G.grid_n[2]);
    const int idx = (i0 * G.grid_n[1] + i1) * G.grid_n[2] + i2;
    sd = __ldg(&G.grid_data[idx].x);
    return idx;
}

// the object tables of a general scene, written by one thread per object (prologue of both kernels)
__device__ __forceinline__ void general_scene_object_record(float4* rec, const float* __restrict__ cost_words, int o, int off_obj_R, int off_obj_pos,
                                                            int off_obj_ident, int off_obj_start, int off_obj_qstart, int off_obj_box_start,
                                                            int off_obj_tol, int off_obj_c2max) {
    const int* iw = reinterpret_cast<const int*>(cost_words);
    const float* Ro = cost_words + off_obj_R + 9 * o;
    const float* po = cost_words + off_obj_pos + 3 * o;
    const int flags = iw[off_obj_ident + o];
    const int k0 = iw[off_obj_start + o], k1 = iw[off_obj_start + o + 1];
    const int q0 = iw[off_obj_qstart + o], q1 = iw[off_obj_qstart + o + 1];
    rec[0] = make_float4(po[0], po[1], po[2], (flags & OBJ_IDENTITY) ? 1.f : 0.f);
    rec[1] = make_float4(Ro[0], Ro[1], Ro[2], __int_as_float(iw[off_obj_box_start + o]));
    rec[2] = make_float4(Ro[3], Ro[4], Ro[5], __int_as_float(iw[off_obj_box_start + o + 1]));
    rec[3] = make_float4(Ro[6], Ro[7], Ro[8], __int_as_float(k0));
    rec[4] = make_float4(__int_as_float(k1 - k0), __int_as_float(q0 / 2), __int_as_float((q1 - q0) / 2), __int_as_float((1 << (flags >> OBJ_BITS_SHIFT)) - 1));
    rec[5] = make_float4(cost_words[off_obj_tol + o], cost_words[off_obj_c2max + o], (flags & OBJ_UNIFORM_RADIUS) ? 1.f : 0.f, (flags & OBJ_GRID) ? 1.f : 0.f);
}

// X: world positions; m / mo: cost and occupancy margins; vmax: value of each point (max over objects of margin - sdf);
// u: world-frame d sdf / d x of the winner (zero when there is none); amin(o, p, index): arg-min of point p in object o
template <int NP, bool GRAD, class ArgminFn>
__device__ __forceinline__ void general_scene_eval(const GeneralScene& G, const float (&X)[NP > 0 ? NP : 1][3], const float (&m)[NP > 0 ? NP : 1],
                                                   const float (&mo)[NP > 0 ? NP : 1], float (&vmax)[NP > 0 ? NP : 1], bool& occ,
                                                   float (&u)[NP > 0 ? NP : 1][3], ArgminFn&& amin) {
    constexpr int N1 = NP > 0 ? NP : 1;
    const float kInf = __int_as_float(0x7f800000);
    int wobj[N1], wsel[N1];      // winner: object, primitive (spheres of the object first, then its boxes; the cell of a voxel grid)
    auto to_object = [&](const float4 o0, const float4 r0, const float4 r1, const float4 r2, float x, float y, float z, float& xo, float& yo,
                         float& zo) {      // x' = (x - pos) @ R for a posed object
        xo = x; yo = y; zo = z;
        if (o0.w == 0.f) {
            const float d0 = __fsub_rn(x, o0.x), d1 = __fsub_rn(y, o0.y), d2 = __fsub_rn(z, o0.z);
            xo = __fmaf_rn(d2, r2.x, __fmaf_rn(d1, r1.x, __fmul_rn(d0, r0.x)));
            yo = __fmaf_rn(d2, r2.y, __fmaf_rn(d1, r1.y, __fmul_rn(d0, r0.y)));
            zo = __fmaf_rn(d2, r2.z, __fmaf_rn(d1, r1.z, __fmul_rn(d0, r0.z)));
        }
    };
#pragma unroll 1
    for (int o = 0; o < G.n_obj; ++o) {
        const float4 o0 = G.obj[6 * o], r0 = G.obj[6 * o + 1], r1 = G.obj[6 * o + 2], r2 = G.obj[6 * o + 3], o4 = G.obj[6 * o + 4];
        const int kb0 = __float_as_int(r0.w), kb1 = __float_as_int(r1.w), k0 = __float_as_int(r2.w), ns = __float_as_int(o4.x);
        float Xo[N1][3], best[N1];
        int bsel[N1];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            to_object(o0, r0, r1, r2, X[k][0], X[k][1], X[k][2], Xo[k][0], Xo[k][1], Xo[k][2]);
            best[k] = kInf; bsel[k] = -1;
        }
        const float4 o5 = G.obj[6 * o + 5];
        if (o5.w != 0.f) {           // voxel SDF lookup: one 4-byte gather per point (the winner's gradient is gathered at the end)
#pragma unroll
            for (int k = 0; k < NP; ++k) bsel[k] = general_grid_lookup(G, Xo[k][0], Xo[k][1], Xo[k][2], best[k]);
        }
        if (ns > 0) {
            const float4* sphs = G.sph + k0;
            bool exact = o5.z == 0.f;                     // unequal radii: every sphere exactly (rare: the shipped scenes are uniform)
            if (!exact) {
                const float4* pq = G.sphq + 2 * __float_as_int(o4.y);
                const int np2 = __float_as_int(o4.z);
                const unsigned lowmask = (unsigned)__float_as_int(o4.w), keep = ~lowmask;
                float m1[N1], m2[N1];
#pragma unroll
                for (int k = 0; k < NP; ++k) { m1[k] = kInf; m2[k] = kInf; }
                unsigned idx = 0u;
                float4 t0 = pq[0], t1 = pq[1];
#pragma unroll 1
                for (int j = 0; j < np2; ++j, idx += 2u) {
                    const float2 cx = make_float2(t0.x, t0.y), cy = make_float2(t0.z, t0.w), cz = make_float2(t1.x, t1.y), cw = make_float2(t1.z, t1.w);
                    const int jn = (j + 1 < np2) ? j + 1 : j;
                    t0 = pq[2 * jn]; t1 = pq[2 * jn + 1];
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const float2 k2 = __ffma2_rn(make_float2(Xo[k][0], Xo[k][0]), cx,
                                                     __ffma2_rn(make_float2(Xo[k][1], Xo[k][1]), cy, __ffma2_rn(make_float2(Xo[k][2], Xo[k][2]), cz, cw)));
                        const float pa = __uint_as_float((__float_as_uint(k2.x) & keep) | idx);
                        const float pb = __uint_as_float((__float_as_uint(k2.y) & keep) | (idx + 1u));
                        const float lo = fminf(pa, pb), hi = fmaxf(pa, pb);
                        m2[k] = fmin3(fmaxf(m1[k], lo), m2[k], hi);
                        m1[k] = fminf(m1[k], lo);
                    }
                }
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    const float xx = __fmaf_rn(Xo[k][2], Xo[k][2], __fmaf_rn(Xo[k][1], Xo[k][1], __fmul_rn(Xo[k][0], Xo[k][0])));
                    exact = exact || !((m2[k] - m1[k]) > o5.x * (xx + o5.y));
                    bsel[k] = (int)(__float_as_uint(m1[k]) & lowmask);
                }
                if (!exact) {
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const float4 c = sphs[bsel[k]];
                        best[k] = __fsub_rn(norm3_rn(__fsub_rn(Xo[k][0], c.x), __fsub_rn(Xo[k][1], c.y), __fsub_rn(Xo[k][2], c.z)), c.w);
                    }
                }
            }
            if (exact) {                                   // first index wins ties (strict <)
#pragma unroll
                for (int k = 0; k < NP; ++k) { best[k] = kInf; bsel[k] = 0; }
#pragma unroll 1
                for (int kk = 0; kk < ns; ++kk) {
                    const float4 c = sphs[kk];
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const float sd = __fsub_rn(norm3_rn(__fsub_rn(Xo[k][0], c.x), __fsub_rn(Xo[k][1], c.y), __fsub_rn(Xo[k][2], c.z)), c.w);
                        const bool better = sd < best[k];
                        best[k] = better ? sd : best[k];
                        bsel[k] = better ? kk : bsel[k];
                    }
                }
            }
        }
        if (kb0 < kb1) {
            float4 bc = G.box[2 * kb0], bh = G.box[2 * kb0 + 1];
#pragma unroll 1
            for (int kb = kb0; kb < kb1; ++kb) {
                const float4 c = bc, hh = bh;
                const int kn = (kb + 1 < kb1) ? kb + 1 : kb;        // the next box is in flight during this one
                bc = G.box[2 * kn]; bh = G.box[2 * kn + 1];
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    const float sd = box_sdf(c, hh, Xo[k][0], Xo[k][1], Xo[k][2]);
                    const bool better = sd < best[k];                // strict: the earlier primitive wins ties
                    best[k] = better ? sd : best[k];
                    bsel[k] = better ? ns + (kb - kb0) : bsel[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            occ = occ || (best[k] < mo[k]);
            const float val = -__fsub_rn(best[k], m[k]);
            amin(o, k, bsel[k] >= 0 ? bsel[k] : 0);
            if (o == 0 || val > vmax[k]) { vmax[k] = val; wobj[k] = o; wsel[k] = bsel[k]; }   // the first maximal object keeps the gradient
        }
    }
    if constexpr (GRAD) {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            float ux = 0.f, uy = 0.f, uz = 0.f;
            if (wsel[k] >= 0) {
                const int o = wobj[k];
                const float4 o0 = G.obj[6 * o], r0 = G.obj[6 * o + 1], r1 = G.obj[6 * o + 2], r2 = G.obj[6 * o + 3];
                const int ns = __float_as_int(G.obj[6 * o + 4].x);
                float xo, yo, zo;
                to_object(o0, r0, r1, r2, X[k][0], X[k][1], X[k][2], xo, yo, zo);
                if (G.obj[6 * o + 5].w != 0.f) {      // voxel grid: the cell carries the gradient
                    const float4 cell = __ldg(G.grid_data + wsel[k]);
                    ux = cell.y; uy = cell.z; uz = cell.w;
                } else if (wsel[k] < ns) {             // unit vector from the sphere centre to the point (zero when they coincide)
                    const float4 c = G.sph[__float_as_int(r2.w) + wsel[k]];
                    const float vx = __fsub_rn(xo, c.x), vy = __fsub_rn(yo, c.y), vz = __fsub_rn(zo, c.z);
                    float rs;
                    norm3_clamped(vx, vy, vz, rs);
                    ux = vx * rs; uy = vy * rs; uz = vz * rs;
                } else {
                    const int kb = __float_as_int(r0.w) + (wsel[k] - ns);
                    box_sdf_grad(G.box[2 * kb], G.box[2 * kb + 1], xo, yo, zo, ux, uy, uz);
                }
                if (o0.w == 0.f) {                     // back to the world frame: g = R g'
                    const float wx = r0.x * ux + r0.y * uy + r0.z * uz, wy = r1.x * ux + r1.y * uy + r1.z * uz, wz = r2.x * ux + r2.y * uy + r2.z * uz;
                    ux = wx; uy = wy; uz = wz;
                }
            }
            u[k][0] = ux; u[k][1] = uy; u[k][2] = uz;
        }
    }
}

// ------------------------------------------------------------------------------------------
// compile-time view of a robot's tables with the fixed chains folded away
// ------------------------------------------------------------------------------------------
struct HXf { float R[9]; float t[3]; };

template <class R>
struct HTree {
    static constexpr int L = R::L, D = R::D;
    __host__ __device__ static constexpr int tix(int body) {
        for (int i = 0; i < L; ++i) if (R::body[i] == body) return i;
        return 0;
    }
    __host__ __device__ static constexpr int pix(int i) { return R::parent[i] < 0 ? -1 : tix(R::parent[i]); }   // topo index of the parent
    __host__ __device__ static constexpr bool movable(int i) { return R::qidx[i] >= 0 && R::kind[i] != TRK_KIND_FIXED; }
    __host__ __device__ static constexpr bool revolute(int i) { return movable(i) && R::kind[i] == TRK_KIND_REVOLUTE; }
    // nearest movable-or-root ancestor-or-self
    __host__ __device__ static constexpr int anchor(int i) {
        int a = i;
        while (pix(a) >= 0 && !movable(a)) a = pix(a);
        return a;
    }
    __host__ __device__ static constexpr int ap(int i) { return anchor(pix(i)); }   // i is not the root
    __host__ __device__ static constexpr HXf identity() { return HXf{{1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, {0.f, 0.f, 0.f}}; }
    // pose of body j in the frame of anchor(j): identity for movable bodies and the root, the folded chain otherwise
    __host__ __device__ static constexpr HXf rel(int j) {
        if (pix(j) < 0 || movable(j)) return identity();
        return origin(j);
    }
    // joint-origin frame of body i (before the joint motion) in the frame of ap(i): rel(parent) o (Rf_i, t_i)
    __host__ __device__ static constexpr HXf origin(int i) {
        const HXf p = rel(pix(i));
        HXf o = identity();
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) {
                float acc = 0.f;
                for (int k = 0; k < 3; ++k) acc += p.R[r * 3 + k] * R::Rf[i][k * 3 + c];
                o.R[r * 3 + c] = acc;
            }
            float acc = p.t[r];
            for (int k = 0; k < 3; ++k) acc += p.R[r * 3 + k] * R::t[i][k];
            o.t[r] = acc;
        }
        return o;
    }
    // prismatic axis in the frame of ap(i) (the reference adds axis * q to the joint-origin translation in the PARENT's
    // frame, rigid_body.py:176)
    __host__ __device__ static constexpr float paxis(int i, int r) {
        const HXf p = rel(pix(i));
        return p.R[r * 3 + 0] * R::ax[i][0] + p.R[r * 3 + 1] * R::ax[i][1] + p.R[r * 3 + 2] * R::ax[i][2];
    }
    __host__ __device__ static constexpr bool origin_t_zero(int i) {
        const HXf o = origin(i);
        return o.t[0] == 0.f && o.t[1] == 0.f && o.t[2] == 0.f;
    }
    // the origin of body b coincides structurally with the origin of movable body j (an ancestor-or-self anchor): every
    // translation on the way is exactly zero and no prismatic joint moves it
    __host__ __device__ static constexpr bool coincident(int b, int j) {
        int a = b;
        while (a != j) {
            if (!origin_t_zero(a)) return false;
            if (movable(a) && !revolute(a)) return false;
            a = ap(a);
            if (pix(a) < 0 && a != j) return false;
        }
        return true;
    }
    // a force applied at the origin of body b moves no joint: every movable ancestor-or-self is revolute about an axis
    // through that origin (or there is none)
    __host__ __device__ static constexpr bool force_is_useless(int b) {
        int j = anchor(b);
        while (pix(j) >= 0) {
            if (!revolute(j) || !coincident(b, j)) return false;
            j = ap(j);
        }
        return true;
    }

    // movable children in the folded tree (the reverse sweep's "first write is an assignment" bookkeeping)
    __host__ __device__ static constexpr bool has_movable_child(int i) {
        for (int c = i + 1; c < L; ++c) if (movable(c) && ap(c) == i) return true;
        return false;
    }
    __host__ __device__ static constexpr bool has_later_movable_sibling(int pa, int i) {
        for (int c = i + 1; c < L; ++c) if (movable(c) && ap(c) == pa) return true;
        return false;
    }
    // topo index of the body q column c drives
    __host__ __device__ static constexpr int body_of_column(int c) {
        for (int i = 0; i < L; ++i) if (R::qidx[i] == c && movable(i)) return i;
        return 0;
    }
    // movable body j is an ancestor-or-self anchor of body b (its joint moves b)
    __host__ __device__ static constexpr bool is_anchor_above(int j, int b) {
        int a = anchor(b);
        while (true) {
            if (a == j) return true;
            if (pix(a) < 0) return false;
            a = ap(a);
        }
    }
    // the world position of body b does not depend on q (same criterion), and its value
    __host__ __device__ static constexpr bool pos_is_const(int b) { return force_is_useless(b); }
    __host__ __device__ static constexpr HXf const_pose(int b) {      // .t = the constant position; valid when pos_is_const(b)
        const int a = anchor(b);
        if (pix(a) < 0) return rel(b);
        int top = a;
        while (pix(ap(top)) >= 0) top = ap(top);
        return origin(top);
    }

    // pose of movable body I from the pose of its anchor-parent.  ROOTP: the anchor-parent is the root (identity pose).
    template <int I, bool SMALL>
    static __device__ __forceinline__ void compose(const float (&Wp)[12], float theta, float (&Wc)[12]) {
        constexpr HXf O = origin(I);
        constexpr bool rootp = pix(ap(I)) < 0;
        float A[9];
        sfor<3>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            sfor<3>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                if constexpr (rootp) A[r * 3 + c] = O.R[r * 3 + c];
                else A[r * 3 + c] = hdot3(Wp[r * 3 + 0], Wp[r * 3 + 1], Wp[r * 3 + 2], O.R[0 + c], O.R[3 + c], O.R[6 + c], false, 0.f);
            });
            if constexpr (R::kind[I] == TRK_KIND_PRISMATIC) {
                constexpr float a0 = paxis(I, 0), a1 = paxis(I, 1), a2 = paxis(I, 2);
                const float t0 = (a0 != 0.f) ? fmaf(a0, theta, O.t[0]) : O.t[0];
                const float t1 = (a1 != 0.f) ? fmaf(a1, theta, O.t[1]) : O.t[1];
                const float t2 = (a2 != 0.f) ? fmaf(a2, theta, O.t[2]) : O.t[2];
                if constexpr (rootp) Wc[9 + r] = (r == 0) ? t0 : ((r == 1) ? t1 : t2);
                else Wc[9 + r] = fmaf(Wp[r * 3 + 2], t2, fmaf(Wp[r * 3 + 1], t1, fmaf(Wp[r * 3 + 0], t0, Wp[9 + r])));
            } else {
                if constexpr (rootp) Wc[9 + r] = O.t[r];
                else Wc[9 + r] = hdot3(Wp[r * 3 + 0], Wp[r * 3 + 1], Wp[r * 3 + 2], O.t[0], O.t[1], O.t[2], true, Wp[9 + r]);
            }
        });
        if constexpr (R::kind[I] == TRK_KIND_REVOLUTE && R::axis_s[I] != 0.f) {
            float sn, cs;
            if constexpr (SMALL) trk_sincos_small(theta, sn, cs); else trk_sincos(theta, sn, cs);
            if constexpr (R::axis_s[I] < 0.f) sn = -sn;
            constexpr int k = R::axis_k[I], ci = (k + 1) % 3, cj = (k + 2) % 3;
            sfor<3>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                if constexpr (rootp) {
                    constexpr float a = O.R[r * 3 + ci], b = O.R[r * 3 + cj];
                    Wc[r * 3 + ci] = hdot3(cs, sn, 0.f, a, b, 0.f, false, 0.f);
                    Wc[r * 3 + cj] = hdot3(cs, sn, 0.f, b, -a, 0.f, false, 0.f);
                } else {
                    const float a = A[r * 3 + ci], b = A[r * 3 + cj];
                    Wc[r * 3 + ci] = fmaf(cs, a, sn * b);
                    Wc[r * 3 + cj] = fmaf(cs, b, -(sn * a));
                }
                Wc[r * 3 + k] = A[r * 3 + k];
            });
        } else {
            sfor<9>([&](auto ic) { Wc[ic] = A[ic]; });
        }
    }

    // position of fixed body I from the pose of its anchor
    template <int I>
    static __device__ __forceinline__ void fixed_position(const float (&Wa)[12], float (&p)[3]) {
        constexpr HXf O = origin(I);
        constexpr bool rootp = pix(ap(I)) < 0;
        sfor<3>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            if constexpr (rootp) p[r] = O.t[r];
            else p[r] = hdot3(Wa[r * 3 + 0], Wa[r * 3 + 1], Wa[r * 3 + 2], O.t[0], O.t[1], O.t[2], true, Wa[9 + r]);
        });
    }
    // rotation of fixed body I from the pose of its anchor
    template <int I>
    static __device__ __forceinline__ void fixed_rotation(const float (&Wa)[12], float (&Rb)[9]) {
        constexpr HXf O = origin(I);
        constexpr bool rootp = pix(ap(I)) < 0;
        sfor<9>([&](auto ec) {
            constexpr int e = decltype(ec)::value, r = e / 3, c = e % 3;
            if constexpr (rootp) Rb[e] = O.R[e];
            else Rb[e] = hdot3(Wa[r * 3 + 0], Wa[r * 3 + 1], Wa[r * 3 + 2], O.R[0 + c], O.R[3 + c], O.R[6 + c], false, 0.f);
        });
    }
    // world joint axis of movable body I
    template <int I>
    static __device__ __forceinline__ void joint_axis(const float (&Wp)[12], const float (&Wb)[12], float (&ja)[3]) {
        if constexpr (R::kind[I] == TRK_KIND_REVOLUTE) {
            constexpr int k = R::axis_k[I];
            constexpr float sgn = R::axis_s[I];
            ja[0] = sgn * Wb[0 + k]; ja[1] = sgn * Wb[3 + k]; ja[2] = sgn * Wb[6 + k];
        } else {
            constexpr bool rootp = pix(ap(I)) < 0;
            constexpr float a0 = paxis(I, 0), a1 = paxis(I, 1), a2 = paxis(I, 2);
            sfor<3>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                if constexpr (rootp) ja[r] = (r == 0) ? a0 : ((r == 1) ? a1 : a2);
                else ja[r] = hdot3(Wp[r * 3 + 0], Wp[r * 3 + 1], Wp[r * 3 + 2], a0, a1, a2, false, 0.f);
            });
        }
    }
    // |theta| of joint I is bounded by its limits (the clamp), so the short sin/cos applies
    __host__ __device__ static constexpr bool small_angle(int i) {
        return R::has_lim[i] != 0 && R::lo[i] >= -64.f && R::hi[i] <= 64.f;
    }
};

// which bodies of the collision model CM (and the EE body, topo index EE) carry positions / forces
template <class R, class CM, int EE>
struct HotPlan {
    using T = HTree<R>;
    static constexpr int L = R::L;
    __host__ __device__ static constexpr bool in_cm(int i) {
        const int b = R::body[i];
        for (int s = 0; s < CM::S; ++s) if (CM::pt_link[s] == b) return true;
        for (int p = 0; p < CM::P; ++p) if (CM::pair_i[p] == b || CM::pair_j[p] == b) return true;
        return false;
    }
    __host__ __device__ static constexpr bool pos_needed(int i) { return in_cm(i) || i == EE; }
    __host__ __device__ static constexpr bool has_force(int i) { return in_cm(i) && !T::force_is_useless(i); }
    // collision points whose position moves with the joints ("dynamic"), in table order
    __host__ __device__ static constexpr bool point_is_const(int s) { return T::pos_is_const(T::tix(CM::pt_link[s])); }
    __host__ __device__ static constexpr int n_dyn_points() {
        int nd = 0;
        for (int s = 0; s < CM::S; ++s) if (!point_is_const(s)) ++nd;
        return nd;
    }
    __host__ __device__ static constexpr int dyn_point(int k) {       // k-th dynamic point -> point index
        int nd = 0;
        for (int s = 0; s < CM::S; ++s) if (!point_is_const(s)) { if (nd == k) return s; ++nd; }
        return 0;
    }
    __host__ __device__ static constexpr int dyn_rank(int s) {        // point index -> rank among the dynamic points, -1 = constant
        if (point_is_const(s)) return -1;
        int nd = 0;
        for (int s2 = 0; s2 < s; ++s2) if (!point_is_const(s2)) ++nd;
        return nd;
    }
    // the constant points as data (for the host: it evaluates them once per launch)
    struct ConstPoints { bool is_const[kHotMaxPoints]; float pos[kHotMaxPoints][3]; };
    __host__ __device__ static constexpr ConstPoints make_const_points() {
        ConstPoints c{};
        for (int s = 0; s < CM::S; ++s) {
            c.is_const[s] = point_is_const(s);
            const HXf pose = T::const_pose(T::tix(CM::pt_link[s]));
            for (int r = 0; r < 3; ++r) c.pos[s][r] = c.is_const[s] ? pose.t[r] : 0.f;
        }
        return c;
    }
    static constexpr ConstPoints kConstPoints = make_const_points();
    // movable body m: some body anchored at m carries a force (or the EE wrench)
    __host__ __device__ static constexpr bool own_wrench(int m) {
        for (int b = 0; b < L; ++b) if (T::anchor(b) == m && (has_force(b) || b == EE)) return true;
        return false;
    }
    __host__ __device__ static constexpr bool subtree_wrench(int m) {
        if (own_wrench(m)) return true;
        for (int c = m + 1; c < L; ++c) if (T::movable(c) && T::ap(c) == m && subtree_wrench(c)) return true;
        return false;
    }
    // accumulator of movable body m already written when the EE wrench (anchored at m) arrives: a child pushed
    __host__ __device__ static constexpr bool started_before_ee(int m) {
        for (int c = m + 1; c < L; ++c) if (T::movable(c) && T::ap(c) == m && subtree_wrench(c)) return true;
        return false;
    }
    // ... when body b (anchored at m) adds its force: a child pushed, the EE wrench came first, or a body with a larger
    // topo index did (the EE body itself may also carry a collision point)
    __host__ __device__ static constexpr bool started_before_body(int m, int b) {
        if (started_before_ee(m)) return true;
        if (T::anchor(EE) == m) return true;
        for (int b2 = b + 1; b2 < L; ++b2) if (T::anchor(b2) == m && has_force(b2)) return true;
        return false;
    }
    // accumulator of pa already written when movable child c pushes into it
    __host__ __device__ static constexpr bool started_before_child(int pa, int c) {
        for (int c2 = c + 1; c2 < L; ++c2) if (T::movable(c2) && T::ap(c2) == pa && subtree_wrench(c2)) return true;
        return false;
    }
};

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
// SCENE_EE: bit 0 = the scene is made of boxes (else: one equal-radius identity-pose sphere object), bit 1 = the cost table has
// no EE term (PlanningTask.compute_collision_cost)
template <class R, class CM, int EE, bool PER_CFG_TARGET, bool WANT_GRAD, int SCENE_EE = 0>
__global__ void __launch_bounds__(kSBlock, TRK_HOT_MINBLOCKS)
k_objective_hot(const __grid_constant__ HotParams hp, const float* __restrict__ cost_words, const float* __restrict__ q, long long n,
                const float* __restrict__ ee_target_cfg, float* __restrict__ cost_total, float* __restrict__ cost_terms,
                float* __restrict__ grad_q, uint8_t* __restrict__ occupancy, int* __restrict__ argmin, float* __restrict__ block_sums) {
    using T = HTree<R>;
    using Plan = HotPlan<R, CM, EE>;
    constexpr int L = R::L, D = R::D, S = CM::S, P = CM::P;
    constexpr int SCENE = SCENE_EE & 1;
    constexpr bool HAS_EE = (SCENE_EE & 2) == 0;
    static_assert(S <= kHotMaxPoints && P <= kHotMaxPairs, "collision model too large for HotParams");
    extern __shared__ __align__(16) float hsm[];
    // [n_pairs2 * 8] scan table, two spheres interleaved: (ax, bx, ay, by) (az, bz, aw, bw) | [n_sph * 4] (c, r) | q tiles
    float4* tq = reinterpret_cast<float4*>(hsm);
    float4* tsph = tq + 2 * hp.n_pairs2;
    const int tid = threadIdx.x;
    if constexpr (SCENE == 1) {
        // general scene: [2 n_box] boxes (c.xyz, radius | half.xyz, rounded flag) | [n_sph_all] spheres (c, r) | [2 n_sphq_all]
        // scan table, two spheres interleaved | [6 n_obj] objects: (pos.xyz, identity flag) (R rows; .w: box range, first sphere)
        // (sphere count, first scan pair, scan pairs, index mask) (tol, c2max, uniform-radius flag, voxel-grid flag)
        const float4* box = reinterpret_cast<const float4*>(cost_words + hp.off_box);
        for (int i = tid; i < 2 * hp.n_box; i += kSBlock) tq[i] = __ldg(box + i);
        float4* gsph = tq + 2 * hp.n_box;
        const float4* sph = reinterpret_cast<const float4*>(cost_words + hp.off_sph_all);
        for (int i = tid; i < hp.n_sph_all; i += kSBlock) gsph[i] = __ldg(sph + i);
        float* gq = reinterpret_cast<float*>(gsph + hp.n_sph_all);
        const float* sphq = cost_words + hp.off_sphq_all;
        for (int i = tid; i < hp.n_sphq_all * 4; i += kSBlock) {
            const int k = i >> 2, comp = i & 3;
            gq[(k >> 1) * 8 + comp * 2 + (k & 1)] = __ldg(sphq + i);
        }
        float4* tobj = reinterpret_cast<float4*>(gq) + hp.n_sphq_all;
        for (int o = tid; o < hp.n_obj; o += kSBlock)
            general_scene_object_record(tobj + 6 * o, cost_words, o, hp.off_obj_R, hp.off_obj_pos, hp.off_obj_ident, hp.off_obj_start,
                                        hp.off_obj_qstart, hp.off_obj_box_start, hp.off_obj_tol, hp.off_obj_c2max);
    } else {
        const float* sphq = cost_words + hp.off_sphq;
        for (int i = tid; i < hp.n_pairs2 * 8; i += kSBlock) {
            const int k = i >> 2, comp = i & 3;
            hsm[(k >> 1) * 8 + comp * 2 + (k & 1)] = __ldg(sphq + i);
        }
        const float4* sph = reinterpret_cast<const float4*>(cost_words + hp.off_sph);
        for (int i = tid; i < hp.n_sph; i += kSBlock) tsph[i] = __ldg(sph + i);
    }
#if TRK_HOT_QSTAGE >= 2
    // q staging: a group of kStage threads (the block, or one warp: no block barrier) copies the group's kStage x D
    // joint values with cp.async into its own double buffer while the previous tile is being evaluated
    constexpr int kStage = (TRK_HOT_QSTAGE == 2) ? kSBlock : 32;
    constexpr int kTileWords = kStage * D, kTileChunks = (kTileWords + 3) / 4;
    static_assert((kTileWords * 4) % 16 == 0, "a staged tile must be a whole number of 16-byte chunks");
    const int gt = tid % kStage;
    float* qs = reinterpret_cast<float*>(tsph + hp.n_sph) + (tid / kStage) * 2 * kTileWords;      // [group][2][kStage * D]
    auto stage_sync = [&]() { if (TRK_HOT_QSTAGE == 2) __syncthreads(); else __syncwarp(); };
    // a full tile is kTileChunks 16-byte chunks; the last (partial) tile zero-fills what lies beyond n
    auto prefetch_tile = [&](long long tile, int buf) {
        const long long cfg0 = tile * kSBlock + (tid / kStage) * kStage;      // first configuration of the group's tile
        const float* src0 = q + cfg0 * D;
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(qs) + (unsigned)(buf * kTileWords * 4);
        if (cfg0 + kStage <= n) {
#pragma unroll
            for (int c0 = 0; c0 < kTileChunks; c0 += kStage) {
                const int c = c0 + gt;
                if (c0 + kStage <= kTileChunks || c < kTileChunks)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16u * c), "l"(src0 + 4 * c) : "memory");
            }
        } else {
            const long long left_words = (n - cfg0) * D;
            for (int c = gt; c < kTileChunks; c += kStage) {
                const long long left = (left_words - 4 * c) * 4;
                const int bytes = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + 16u * c), "l"(bytes > 0 ? src0 + 4 * c : q), "r"(bytes) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#endif
    const long long n_tiles = (n + kSBlock - 1) / kSBlock;
    long long tile = blockIdx.x;
    float sum_acc = 0.f;          // this thread's share of sum(total): reduced per block at the end (block_sums)
#if TRK_HOT_QSTAGE >= 2
    int buf = 0;
    if (tile < n_tiles) prefetch_tile(tile, 0);
#elif TRK_HOT_QSTAGE == 1
    float qn[D];
    {
        const long long c0 = tile * kSBlock + tid;
        const float* qrow = q + (c0 < n ? c0 : n - 1) * D;
        sfor<D>([&](auto c) { qn[c] = __ldg(qrow + c); });
    }
#endif
    __syncthreads();
    for (; tile < n_tiles; tile += gridDim.x) {
        const long long cfg = tile * kSBlock + tid;
        const bool live = cfg < n;
        float qv[D];
#if TRK_HOT_QSTAGE >= 2
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        stage_sync();
        sfor<D>([&](auto c) { qv[c] = qs[buf * kTileWords + gt * D + c]; });
        if (tile + gridDim.x < n_tiles) prefetch_tile(tile + gridDim.x, buf ^ 1);
        buf ^= 1;
#elif TRK_HOT_QSTAGE == 1
        sfor<D>([&](auto c) { qv[c] = qn[c]; });
        {
            const long long c1 = cfg + (long long)gridDim.x * kSBlock;
            const float* qrow = q + (c1 < n ? c1 : n - 1) * D;
            sfor<D>([&](auto c) { qn[c] = __ldg(qrow + c); });
        }
#else
        {
            const float* qrow = q + (live ? cfg : n - 1) * D;
            sfor<D>([&](auto c) { qv[c] = __ldg(qrow + c); });
        }
#endif
        // ---------------- forward kinematics (movable bodies only; fixed bodies on demand) ----------------
        float W[L][12], Pb[L][3], JA[D > 0 ? D : 1][3], Ree[9];
        unsigned pass = 0u;
        sfor<L>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (T::pix(i) < 0) {
                sfor<3>([&](auto rc) { Pb[i][rc] = 0.f; });
                if constexpr (i == EE) sfor<9>([&](auto e) { Ree[e] = (e == 0 || e == 4 || e == 8) ? 1.f : 0.f; });
            } else if constexpr (T::movable(i)) {
                constexpr int qi = R::qidx[i];
                const float qraw = qv[qi];
                float theta = qraw;
                if constexpr (R::has_lim[i] != 0) {
                    constexpr float lo = R::lo[i], hi = R::hi[i];
                    theta = fmin_nan(fmax_nan(qraw, lo), hi);                  // torch.clamp: NaN propagates
                    if (theta == qraw) pass |= (1u << qi);                     // gradient passes inside [lo, hi], bounds included
                } else {
                    pass |= (1u << qi);
                }
                constexpr int a = T::ap(i);
                T::template compose<i, T::small_angle(i)>(W[a], theta, W[i]);
                sfor<3>([&](auto rc) { Pb[i][rc] = W[i][9 + rc]; });
                T::template joint_axis<i>(W[a], W[i], JA[qi]);
                if constexpr (i == EE) sfor<9>([&](auto e) { Ree[e] = W[i][e]; });
            } else {
                constexpr int a = T::anchor(i);
                if constexpr (Plan::pos_needed(i)) T::template fixed_position<i>(W[a], Pb[i]);
                if constexpr (i == EE) T::template fixed_rotation<i>(W[a], Ree);
            }
        });
        // ---------------- EE SE(3) term (geometrics/utils.py:130-154; distance_fields.py:347-356) ----------------
        float c_ee = 0.f, eF[3] = {0.f, 0.f, 0.f}, eT[3] = {0.f, 0.f, 0.f};
        if constexpr (HAS_EE) {
            float Rt[9], pt[3];
            if constexpr (PER_CFG_TARGET) {
                const float* Ht = ee_target_cfg + (live ? cfg : 0) * 16;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float4 row = __ldg(reinterpret_cast<const float4*>(Ht) + i);
                    Rt[i * 3 + 0] = row.x; Rt[i * 3 + 1] = row.y; Rt[i * 3 + 2] = row.z; pt[i] = row.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i) Rt[i] = hp.ee_R[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) pt[i] = hp.ee_t[i];
            }
            float dg[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                dg[i] = __fmaf_rn(Ree[i * 3 + 2], Rt[i * 3 + 2], __fmaf_rn(Ree[i * 3 + 1], Rt[i * 3 + 1], __fmul_rn(Ree[i * 3 + 0], Rt[i * 3 + 0])));
            const float tr = __fadd_rn(__fadd_rn(dg[0], dg[1]), dg[2]);
            const float phi = __fmul_rn(__fsub_rn(tr, 1.f), 0.5f);
            float Dd = __fmul_rn(hp.ee_w_rot, __fsub_rn(1.f, phi));
            const float vx = __fsub_rn(Pb[EE][0], pt[0]), vy = __fsub_rn(Pb[EE][1], pt[1]), vz = __fsub_rn(Pb[EE][2], pt[2]);
            float rs;
            const float dist = norm3_clamped(vx, vy, vz, rs);
            Dd = __fadd_rn(Dd, __fmul_rn(hp.ee_w_pos, dist));
            c_ee = hp.ee_square ? __fmul_rn(Dd, Dd) : Dd;
            if constexpr (WANT_GRAD) {
            const float scale = hp.w[3] * (hp.ee_square ? 2.f * Dd : 1.f);
            const float kr = scale * (-0.5f * hp.ee_w_rot);
            const float kp = scale * hp.ee_w_pos * rs;
            eF[0] = kp * vx; eF[1] = kp * vy; eF[2] = kp * vz;
            // torque about the world origin: p x F + kr * axial(Rt Ree^T)
            const float m21 = Rt[6] * Ree[3] + Rt[7] * Ree[4] + Rt[8] * Ree[5];
            const float m12 = Rt[3] * Ree[6] + Rt[4] * Ree[7] + Rt[5] * Ree[8];
            const float m02 = Rt[0] * Ree[6] + Rt[1] * Ree[7] + Rt[2] * Ree[8];
            const float m20 = Rt[6] * Ree[0] + Rt[7] * Ree[1] + Rt[8] * Ree[2];
            const float m10 = Rt[3] * Ree[0] + Rt[4] * Ree[1] + Rt[5] * Ree[2];
            const float m01 = Rt[0] * Ree[3] + Rt[1] * Ree[4] + Rt[2] * Ree[5];
            eT[0] = fmaf(kr, m21 - m12, fmaf(Pb[EE][1], eF[2], -(Pb[EE][2] * eF[1])));
            eT[1] = fmaf(kr, m02 - m20, fmaf(Pb[EE][2], eF[0], -(Pb[EE][0] * eF[2])));
            eT[2] = fmaf(kr, m10 - m01, fmaf(Pb[EE][0], eF[1], -(Pb[EE][1] * eF[0])));
            }
        }
        // ---------------- self-collision pairs (distance_fields.py:194-215) ----------------
        float G[L][3];
        bool gst[L];
        sfor<L>([&](auto ic) { gst[ic] = false; });
        auto add_force = [&](auto bc, float fx, float fy, float fz) {   // G[b] += f with a compile-time "first write"
            constexpr int b = decltype(bc)::value;
            if (gst[b]) { G[b][0] += fx; G[b][1] += fy; G[b][2] += fz; }
            else { G[b][0] = fx; G[b][1] = fy; G[b][2] = fz; gst[b] = true; }
        };
        bool occ_self = false, occ_obj = false, occ_ws = false;
        float c_self = 0.f;
        sfor<P>([&](auto pc) {
            constexpr int p = decltype(pc)::value, bi = T::tix(CM::pair_i[p]), bj = T::tix(CM::pair_j[p]);
            const float vx = __fsub_rn(Pb[bi][0], Pb[bj][0]), vy = __fsub_rn(Pb[bi][1], Pb[bj][1]), vz = __fsub_rn(Pb[bi][2], Pb[bj][2]);
            float rs;
            const float dist = norm3_clamped(vx, vy, vz, rs);
            occ_self = occ_self || (dist < hp.mo_pair[p]);
            c_self = __fadd_rn(c_self, -__fsub_rn(dist, hp.m_pair[p]));
            if constexpr (WANT_GRAD) {
                const float inv = hp.w[0] * rs;      // d(-dist)/d x_i = -v / dist (zero when the origins coincide: v = 0)
                if constexpr (Plan::has_force(bi)) add_force(IC<bi>{}, -(vx * inv), -(vy * inv), -(vz * inv));
                if constexpr (Plan::has_force(bj)) add_force(IC<bj>{}, vx * inv, vy * inv, vz * inv);
            }
        });
        constexpr int ND = Plan::n_dyn_points();
        // workspace box of one moving point (distance_fields.py:319-332): the six half-space distances in the reference's order
        // [x - min (3), max - x (3)]; value = margin - smallest distance (= the reference's max of margin - d_i: rounding is
        // monotone); the gradient goes to the first index that attains the smallest distance.  (gx, gy, gz): the object term's
        // force on the point, completed here and handed to the point's body.
        float c_obj = hp.c_obj_const, c_ws = hp.c_ws_const;
        auto ws_and_force = [&](auto kc, float x, float y, float z, float gx, float gy, float gz) {
            constexpr int k = decltype(kc)::value, s = Plan::dyn_point(k), b = T::tix(CM::pt_link[s]);
            const float e0 = __fsub_rn(x, hp.ws_min[0]), e1 = __fsub_rn(y, hp.ws_min[1]), e2 = __fsub_rn(z, hp.ws_min[2]);
            const float f0 = __fsub_rn(hp.ws_max[0], x), f1 = __fsub_rn(hp.ws_max[1], y), f2 = __fsub_rn(hp.ws_max[2], z);
            const float emin = fmin3(e0, e1, e2), fmn = fmin3(f0, f1, f2);
            const float dmin = fminf(emin, fmn);
            occ_ws = occ_ws || (dmin < hp.mo_ws[s]);
            c_ws = __fadd_rn(c_ws, -__fsub_rn(dmin, hp.m_ws[s]));
            if constexpr (WANT_GRAD) {
                const bool upper = fmn < emin;                       // ties go to the lower index: the x - min side
                float g = upper ? hp.w[2] : -hp.w[2];
                g = (dmin != 0.f) ? g : 0.f;                         // sign(d) |d| has zero gradient at d == 0
                const float a0 = upper ? f0 : e0, a1 = upper ? f1 : e1;
                const bool p0 = a0 == dmin, p1 = a1 == dmin;
                gx += p0 ? g : 0.f;
                gy += (!p0 && p1) ? g : 0.f;
                gz += (!p0 && !p1) ? g : 0.f;
                if constexpr (Plan::has_force(b)) add_force(IC<b>{}, gx, gy, gz);
            }
        };
        // ---------------- general scene: every object's spheres, boxes and voxel grid against all moving points ----------------
        [[maybe_unused]] int widx[ND > 0 ? ND : 1];       // sphere scene: index of each moving point's nearest sphere
        if constexpr (SCENE == 1) {
            GeneralScene GS;
            GS.box = tq; GS.sph = tq + 2 * hp.n_box; GS.sphq = GS.sph + hp.n_sph_all; GS.obj = GS.sphq + hp.n_sphq_all; GS.n_obj = hp.n_obj;
            GS.grid_data = hp.grid_data;
            sfor<3>([&](auto i) { GS.grid_n[i] = hp.grid_n[i]; GS.grid_lo[i] = hp.grid_lo[i]; GS.grid_extent[i] = hp.grid_extent[i]; GS.grid_nf[i] = hp.grid_nf[i]; });
            float X[ND > 0 ? ND : 1][3], mg[ND > 0 ? ND : 1], mog[ND > 0 ? ND : 1], vmax[ND > 0 ? ND : 1], ug[ND > 0 ? ND : 1][3];
            sfor<ND>([&](auto kc) {
                constexpr int s = Plan::dyn_point(decltype(kc)::value), b = T::tix(CM::pt_link[s]);
                sfor<3>([&](auto rc) { X[kc][rc] = Pb[b][rc]; });
                mg[kc] = hp.m_obj[s]; mog[kc] = hp.mo_obj[s];
            });
            int dyn_s[ND > 0 ? ND : 1];          // moving point -> point index (compile-time values; indexed by the unrolled loops)
            sfor<ND>([&](auto kc) { dyn_s[kc] = Plan::dyn_point(decltype(kc)::value); });
            general_scene_eval<ND, WANT_GRAD>(GS, X, mg, mog, vmax, occ_obj, ug, [&](int o, int k, int idx) {
                if (argmin && live) argmin[cfg * (long long)(hp.n_obj * S) + o * S + dyn_s[k]] = idx;
            });
            if (hp.grid_obj >= 0) {
                // the constant points against the voxel grid: the host cannot read it; the same cell for every configuration (an
                // L1 hit), combined with the host's value over the other objects
                sfor<S>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    if constexpr (Plan::dyn_rank(s) < 0) {
                        constexpr auto cp = Plan::kConstPoints;
                        float sd;
                        const int idx = general_grid_lookup(GS, cp.pos[s][0], cp.pos[s][1], cp.pos[s][2], sd);
                        occ_obj = occ_obj || (sd < hp.mo_obj[s]);
                        c_obj = __fadd_rn(c_obj, fmaxf(hp.vconst[s], -__fsub_rn(sd, hp.m_obj[s])));
                        if (argmin && live) argmin[cfg * (long long)(hp.n_obj * S) + hp.grid_obj * S + s] = idx;
                    }
                });
            }
            sfor<ND>([&](auto kc) {
                constexpr int k = decltype(kc)::value, s = Plan::dyn_point(k), b = T::tix(CM::pt_link[s]);
                c_obj = __fadd_rn(c_obj, vmax[k]);
                float gx = 0.f, gy = 0.f, gz = 0.f;
                if constexpr (WANT_GRAD) { gx = -(hp.w[1] * ug[k][0]); gy = -(hp.w[1] * ug[k][1]); gz = -(hp.w[1] * ug[k][2]); }
                ws_and_force(kc, Pb[b][0], Pb[b][1], Pb[b][2], gx, gy, gz);
            });
        } else {
        // ---------------- sphere object: joint scan of the moving collision points, two spheres per FFMA2 ----------------
        // (points whose position is a compile-time constant were evaluated on the host: HotParams::*_const)
        float m1[ND > 0 ? ND : 1], m2[ND > 0 ? ND : 1];
        {
            float2 X2[ND > 0 ? ND : 1][3];
            sfor<ND>([&](auto kc) {
                constexpr int b = T::tix(CM::pt_link[Plan::dyn_point(decltype(kc)::value)]);
                sfor<3>([&](auto rc) { X2[kc][rc] = make_float2(Pb[b][rc], Pb[b][rc]); });
            });
            const float kInf = __int_as_float(0x7f800000);
#pragma unroll
            for (int k = 0; k < ND; ++k) { m1[k] = kInf; m2[k] = kInf; }
            const unsigned keep = ~hp.lowmask;
            unsigned idx = 0u;
            float4 t0 = tq[0], t1 = tq[1];
#pragma unroll kHotScanUnroll
            for (int j = 0; j < hp.n_pairs2; ++j, idx += 2u) {
                const float2 cx = make_float2(t0.x, t0.y), cy = make_float2(t0.z, t0.w), cz = make_float2(t1.x, t1.y), cw = make_float2(t1.z, t1.w);
                const int jn = (j + 1 < hp.n_pairs2) ? j + 1 : j;      // next pair's table entry is in flight during this one
                t0 = tq[2 * jn]; t1 = tq[2 * jn + 1];
#pragma unroll
                for (int k = 0; k < ND; ++k) {
                    const float2 k2 = __ffma2_rn(X2[k][0], cx, __ffma2_rn(X2[k][1], cy, __ffma2_rn(X2[k][2], cz, cw)));
                    const float pa = __uint_as_float((__float_as_uint(k2.x) & keep) | idx);
                    const float pb = __uint_as_float((__float_as_uint(k2.y) & keep) | (idx + 1u));
                    const float lo = fminf(pa, pb), hi = fmaxf(pa, pb);
                    m2[k] = fmin3(fmaxf(m1[k], lo), m2[k], hi);
                    m1[k] = fminf(m1[k], lo);
                }
            }
        }
        bool all_certain = true;
        sfor<ND>([&](auto kc) {
            constexpr int k = decltype(kc)::value, b = T::tix(CM::pt_link[Plan::dyn_point(k)]);
            const float xx = __fmaf_rn(Pb[b][2], Pb[b][2], __fmaf_rn(Pb[b][1], Pb[b][1], __fmul_rn(Pb[b][0], Pb[b][0])));
            all_certain = all_certain && ((m2[k] - m1[k]) > hp.tol * (xx + hp.c2max));
            widx[k] = (int)(__float_as_uint(m1[k]) & hp.lowmask);
        });
        if (!all_certain) {   // near tie (or NaN) on some point: exact scan of the uncertain points, first index wins (rare)
            sfor<ND>([&](auto kc) {
                constexpr int k = decltype(kc)::value, b = T::tix(CM::pt_link[Plan::dyn_point(k)]);
                const float xx = __fmaf_rn(Pb[b][2], Pb[b][2], __fmaf_rn(Pb[b][1], Pb[b][1], __fmul_rn(Pb[b][0], Pb[b][0])));
                if (!((m2[k] - m1[k]) > hp.tol * (xx + hp.c2max))) {
                    float best = __int_as_float(0x7f800000);
                    int bidx = 0;
#pragma unroll 1
                    for (int kk = 0; kk < hp.n_sph; ++kk) {
                        const float4 c = tsph[kk];
                        const float sd = __fsub_rn(norm3_rn(__fsub_rn(Pb[b][0], c.x), __fsub_rn(Pb[b][1], c.y), __fsub_rn(Pb[b][2], c.z)), c.w);
                        if (sd < best) { best = sd; bidx = kk; }
                    }
                    widx[k] = bidx;
                }
            });
        }
        // ---------------- moving collision points: exact evaluation of the winner + workspace box ----------------
        sfor<ND>([&](auto kc) {
            constexpr int k = decltype(kc)::value, s = Plan::dyn_point(k), b = T::tix(CM::pt_link[s]);
            const float x = Pb[b][0], y = Pb[b][1], z = Pb[b][2];
            const float4 c = tsph[widx[k]];
            const float vx = __fsub_rn(x, c.x), vy = __fsub_rn(y, c.y), vz = __fsub_rn(z, c.z);
            float rs;
            const float best = __fsub_rn(norm3_clamped(vx, vy, vz, rs), c.w);
            occ_obj = occ_obj || (best < hp.mo_obj[s]);
            c_obj = __fadd_rn(c_obj, -__fsub_rn(best, hp.m_obj[s]));
            const float w1 = hp.w[1] * rs;
            ws_and_force(kc, x, y, z, -(w1 * vx), -(w1 * vy), -(w1 * vz));
        });
        }
        // ---------------- outputs ----------------
        if (live) {
            const float total = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(hp.w[0], c_self), __fmul_rn(hp.w[1], c_obj)), __fmul_rn(hp.w[2], c_ws)),
                                          __fmul_rn(hp.w[3], c_ee));
            if (cost_total) cost_total[cfg] = total;
            sum_acc += total;
            if (cost_terms) reinterpret_cast<float4*>(cost_terms)[cfg] = make_float4(c_self, c_obj, c_ws, c_ee);
            if (occupancy) occupancy[cfg] = (uint8_t)(hp.occ_const | (occ_self ? TRK_OCC_SELF : 0u) | (occ_obj ? TRK_OCC_OBJECT : 0u) | (occ_ws ? TRK_OCC_WS : 0u));
            if (argmin) {
                if constexpr (SCENE == 1) {      // the moving points' entries were written object by object; the constant points' here
                    for (int o = 0; o < hp.n_obj; ++o) {
                        if (o == hp.grid_obj) continue;       // written with the lookup
                        int* row = argmin + cfg * (long long)(hp.n_obj * S) + o * S;
                        sfor<S>([&](auto sc) {
                            constexpr int s = decltype(sc)::value;
                            if constexpr (Plan::dyn_rank(s) < 0) row[s] = hp.widx_const_obj[o][s];
                        });
                    }
                } else {
                    int* row = argmin + cfg * S;
                    sfor<S>([&](auto sc) {
                        constexpr int s = decltype(sc)::value, rk = Plan::dyn_rank(s);
                        if constexpr (rk >= 0) row[s] = widx[rk];
                        else row[s] = hp.widx_const[s];
                    });
                }
            }
        }
        // ---------------- reverse mode: leaf -> root over the movable bodies ----------------
        if constexpr (WANT_GRAD) {
        float F[L][3], Tq[L][3], gq[D > 0 ? D : 1];
        sfor<D>([&](auto c) { gq[c] = 0.f; });
        sfor<L>([&](auto ic) {
            constexpr int i = L - 1 - decltype(ic)::value;
            if constexpr (T::pix(i) >= 0 && T::movable(i) && Plan::subtree_wrench(i)) {
                // forces of the bodies anchored here (the EE wrench first, then bodies by decreasing topo index)
                if constexpr (T::anchor(EE) == i) {
                    constexpr bool st = Plan::started_before_ee(i);
                    sfor<3>([&](auto rc) {
                        F[i][rc] = st ? F[i][rc] + eF[rc] : eF[rc];
                        Tq[i][rc] = st ? Tq[i][rc] + eT[rc] : eT[rc];
                    });
                }
                sfor<L>([&](auto bc) {
                    constexpr int b = L - 1 - decltype(bc)::value;
                    if constexpr (T::anchor(b) == i && Plan::has_force(b)) {
                        constexpr bool st = Plan::started_before_body(i, b);
                        const float px = Pb[b][0], py = Pb[b][1], pz = Pb[b][2], fx = G[b][0], fy = G[b][1], fz = G[b][2];
                        if constexpr (st) {
                            F[i][0] += fx; F[i][1] += fy; F[i][2] += fz;
                            Tq[i][0] = fmaf(py, fz, fmaf(-pz, fy, Tq[i][0]));
                            Tq[i][1] = fmaf(pz, fx, fmaf(-px, fz, Tq[i][1]));
                            Tq[i][2] = fmaf(px, fy, fmaf(-py, fx, Tq[i][2]));
                        } else {
                            F[i][0] = fx; F[i][1] = fy; F[i][2] = fz;
                            Tq[i][0] = fmaf(py, fz, -(pz * fy));
                            Tq[i][1] = fmaf(pz, fx, -(px * fz));
                            Tq[i][2] = fmaf(px, fy, -(py * fx));
                        }
                    }
                });
                // project on the joint axis: revolute a . (T - p x F), prismatic a . F; torch.clamp's mask
                constexpr int qi = R::qidx[i];
                float g;
                if constexpr (R::kind[i] == TRK_KIND_REVOLUTE) {
                    const float px = Pb[i][0], py = Pb[i][1], pz = Pb[i][2];
                    const float mx = fmaf(pz, F[i][1], fmaf(-py, F[i][2], Tq[i][0]));
                    const float my = fmaf(px, F[i][2], fmaf(-pz, F[i][0], Tq[i][1]));
                    const float mz = fmaf(py, F[i][0], fmaf(-px, F[i][1], Tq[i][2]));
                    g = fmaf(JA[qi][2], mz, fmaf(JA[qi][1], my, JA[qi][0] * mx));
                } else {
                    g = fmaf(JA[qi][2], F[i][2], fmaf(JA[qi][1], F[i][1], JA[qi][0] * F[i][0]));
                }
                gq[qi] = ((pass >> qi) & 1u) ? g : 0.f;
                // hand the branch's wrench to the anchor-parent
                constexpr int pa = T::ap(i);
                if constexpr (T::pix(pa) >= 0) {
                    constexpr bool st = Plan::started_before_child(pa, i);
                    sfor<3>([&](auto rc) {
                        F[pa][rc] = st ? F[pa][rc] + F[i][rc] : F[i][rc];
                        Tq[pa][rc] = st ? Tq[pa][rc] + Tq[i][rc] : Tq[i][rc];
                    });
                }
            }
        });
        if (live) {
            float* grow = grad_q + cfg * D;
            sfor<D>([&](auto c) { grow[c] = gq[c]; });
        }
        }
    }
    // fused first stage of the cost reduction (distributed.total_cost): one partial sum per block, written (not
    // accumulated) in block order, so the caller's second stage over gridDim.x values is deterministic
    if (block_sums) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum_acc += __shfl_xor_sync(0xffffffffu, sum_acc, o);
        __syncthreads();                                   // the q tiles / tables in shared memory are no longer read
        if ((tid & 31) == 0) hsm[tid >> 5] = sum_acc;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kSBlock / 32; ++w) s += hsm[w];
            block_sums[blockIdx.x] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Link Jacobians of a compiled-in robot (compute_analytical_jacobian_all_links, robot_tree.py:250-265; closed form as in
// k_link_jacobian of trk_b200.cu): J[cfg][link][7][D] = d (position xyz, quaternion wxyz) / d q.
// Output-bound (L * 7 * D * 4 bytes per configuration).  One warp owns 32 consecutive configurations, whose blocks are ONE
// contiguous span of J.  A block of four warps owns a 32-configuration tile: lane = configuration; every warp runs the unrolled
// FK (redundantly: cheaper than exchanging poses), which leaves every
// joint's world axis and origin in per-lane shared-memory columns, then warp w produces the [7][D] blocks of bodies w, w + 4, ... by a rolled loop
// over the q columns straight into the tile at its output offsets (row stride L*7*D words: conflict-free when odd, e.g.
// Panda's 539), and one lane hands the whole tile to the TMA engine with a single linear bulk store.  While a block's tile
// drains, the other blocks of the SM compute (three tiles fit).  (Unrolling the column loop, or splitting a tile over four
// warps by body range, made 50-64 KB of straight-line code: 5.4 instruction-fetch stalls per issue, 58-62% of HBM.)
// ------------------------------------------------------------------------------------------
template <class R>
struct JacTile {
    using T = HTree<R>;
    static constexpr int L = R::L, D = R::D;
    static constexpr int kRow = L * 7 * (D > 0 ? D : 1);          // floats per configuration
    static constexpr int kCols = 6 * (D > 0 ? D : 1);             // per-lane joint columns: axis (3), origin (3) per q column
    static constexpr size_t smem_bytes = (size_t)32 * (kRow + kCols) * sizeof(float);
    static constexpr bool fits = smem_bytes <= 75 * 1024;         // three blocks per SM
    static constexpr int kWarps = 4;                               // warps per tile: all run the FK, each fills every kWarps-th body's block
    __host__ __device__ static constexpr unsigned anc_mask(int i) {       // q columns that move body i
        unsigned m = 0u;
        for (int c = 0; c < D; ++c) if (T::is_anchor_above(T::body_of_column(c), i)) m |= 1u << c;
        return m;
    }
    __host__ __device__ static constexpr unsigned prismatic_mask() {
        unsigned m = 0u;
        for (int c = 0; c < D; ++c) if (R::kind[T::body_of_column(c)] == TRK_KIND_PRISMATIC) m |= 1u << c;
        return m;
    }
};

template <class R>
__global__ void __launch_bounds__(32 * JacTile<R>::kWarps) k_link_jacobian_static(const float* __restrict__ q, long long n, float* __restrict__ J) {
    using T = HTree<R>;
    using JT = JacTile<R>;
    constexpr int L = R::L, D = R::D, kRow = JT::kRow;
    extern __shared__ __align__(128) float jtile[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(jtile);
    float* row = jtile + lane * kRow;
    float* jc = jtile + 32 * kRow + lane;            // joint columns [(c * 6 + k) * 32]: k 0-2 world axis, 3-5 joint origin
    float qn[D > 0 ? D : 1];                         // next tile's joint values: loaded while the current tile is computed
    {
        const long long c0 = (long long)blockIdx.x * 32 + lane;
        const float* qrow = q + (c0 < n ? c0 : n - 1) * D;
        sfor<D>([&](auto c) { qn[c] = __ldg(qrow + c); });
    }
    for (long long base = (long long)blockIdx.x * 32; base < n; base += (long long)gridDim.x * 32) {
        float qv[D > 0 ? D : 1];
        sfor<D>([&](auto c) { qv[c] = qn[c]; });
        {
            const long long c1 = base + (long long)gridDim.x * 32 + lane;
            const float* qrow = q + (c1 < n ? c1 : n - 1) * D;
            sfor<D>([&](auto c) { qn[c] = __ldg(qrow + c); });
        }
        // the previous bulk store must have finished READING the tile (and every warp its joint columns) before they are overwritten
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        float W[L][12];
        unsigned pass = 0u;
        // phase 1, every warp: the unrolled FK of the lane's configuration (warp 0 also leaves the joint columns in shared memory)
        sfor<L>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if constexpr (T::pix(i) < 0) {
                sfor<12>([&](auto k) { W[i][k] = (k == 0 || k == 4 || k == 8) ? 1.f : 0.f; });
            } else if constexpr (T::movable(i)) {
                constexpr int qi = R::qidx[i];
                const float qraw = qv[qi];
                float theta = qraw;
                if constexpr (R::has_lim[i] != 0) {
                    constexpr float lo = R::lo[i], hi = R::hi[i];
                    theta = fmin_nan(fmax_nan(qraw, lo), hi);
                    if (theta == qraw) pass |= 1u << qi;          // clamped joints: zero columns
                } else {
                    pass |= 1u << qi;
                }
                constexpr int a = T::ap(i);
                T::template compose<i, T::small_angle(i)>(W[a], theta, W[i]);
                float ja[3];
                T::template joint_axis<i>(W[a], W[i], ja);
                if (warp == 0) {
                    float* o = jc + qi * 6 * 32;
                    o[0] = ja[0]; o[32] = ja[1]; o[64] = ja[2]; o[96] = W[i][9]; o[128] = W[i][10]; o[160] = W[i][11];
                }
            } else {
                constexpr int a = T::anchor(i);
                float Rb[9], pb[3];
                T::template fixed_rotation<i>(W[a], Rb);
                T::template fixed_position<i>(W[a], pb);
                sfor<9>([&](auto k) { W[i][k] = Rb[k]; });
                sfor<3>([&](auto k) { W[i][9 + k] = pb[k]; });
            }
        });
        __syncthreads();
        // phase 2, warp w: the blocks of bodies w, w + kWarps, ...  A body's block is [7][D] at offset body * 7 * D of the row,
        // filled by a rolled loop over the q columns (unrolled it is 50 KB of straight-line code that stalls on instruction fetch)
        sfor<L>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            if ((i % JT::kWarps) != warp) return;
            float qt[4];
            quat_of_rotation(W[i], qt);
            const float px = W[i][9], py = W[i][10], pz = W[i][11];
            const unsigned on_mask = JT::anc_mask(i) & pass;
            float* o = row + R::body[i] * 7 * D;
#pragma unroll 1
            for (int c = 0; c < D; ++c) {
                const float* j = jc + c * 6 * 32;
                const float ax = j[0], ay = j[32], az = j[64];
                const bool on = (on_mask >> c) & 1u, pris = (JT::prismatic_mask() >> c) & 1u;
                const bool rev = on && !pris, pr = on && pris;      // selects, not products: a NaN pose must not leak into a zero column
                const float dx = px - j[96], dy = py - j[128], dz = pz - j[160];
                o[c]         = rev ? ay * dz - az * dy : (pr ? ax : 0.f);
                o[D + c]     = rev ? az * dx - ax * dz : (pr ? ay : 0.f);
                o[2 * D + c] = rev ? ax * dy - ay * dx : (pr ? az : 0.f);
                o[3 * D + c] = rev ? -0.5f * (ax * qt[1] + ay * qt[2] + az * qt[3]) : 0.f;
                o[4 * D + c] = rev ? 0.5f * (qt[0] * ax + (ay * qt[3] - az * qt[2])) : 0.f;
                o[5 * D + c] = rev ? 0.5f * (qt[0] * ay + (az * qt[1] - ax * qt[3])) : 0.f;
                o[6 * D + c] = rev ? 0.5f * (qt[0] * az + (ax * qt[2] - ay * qt[1])) : 0.f;
            }
        });
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const long long left = n - base;
        const int ncfg = left < 32 ? (int)left : 32;
        float* dst = J + base * kRow;
        if ((ncfg * kRow) % 4 == 0) {          // the span is a whole number of 16-byte chunks (always for 32 configurations)
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tile_s), "r"(ncfg * kRow * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {                                // ragged last tile
            for (int k = threadIdx.x; k < ncfg * kRow; k += blockDim.x) dst[k] = jtile[k];
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace trk
