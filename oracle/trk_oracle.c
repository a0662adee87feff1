/*
 * trk_oracle.c - plain-C restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this library; nothing under torch_robotics_b200/ does.  It exists (a) to check the CUDA kernels
 * at sizes where the numpy restatement (oracle/restatement.py) is slow and (b) as the CPU port
 * that bench.py times on the host cores (OpenMP over configurations).
 *
 * Restated reference code (paths relative to /root/reference/torch_robotics):
 *   torch_kinematics_tree/models/rigid_body.py:146-211      per-link joint transform + composition
 *   torch_kinematics_tree/geometrics/spatial_vector.py:8-47 axis rotations
 *   torch_kinematics_tree/geometrics/utils.py:11-17,130-154 multiply_transform, SE3_distance
 *   torch_planning_objectives/fields/distance_fields.py:107-130,194-215,283-291,307-332,347-356
 *   environments/primitives.py:108-112,387-405              sphere SDF, object pose
 *   tasks/tasks.py:139-232                                  composition of the three collision terms
 * Reverse mode: the reference uses autograd; the analytic wrench form is SURVEY.md section 8(a).
 *
 * Pinned against the tests/golden fixtures (recorded from the unmodified reference) through
 * tests/test_c_oracle.py, and cross-checked bit-for-bit against oracle/restatement.py on the
 * decision values.  Compile with -ffp-contract=off: fused operations appear only where fmaf()
 * is written, which is the canonical order documented in oracle/restatement.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXL 64

typedef struct {
    int32_t n_links, n_dofs;
    const int32_t *parent, *kind, *axis_k;
    const float *axis_s, *axis;
    const int32_t *q_index;
    const float *R_fixed, *trans, *lo, *hi;
    const int32_t *has_lim, *topo;
    const float *base_pose; /* [12] or NULL */
} orc_tree;

typedef struct {
    int32_t n_points;
    const int32_t *pt_link;
    const float *pt_margin_obj, *pt_margin_ws;
    int32_t n_objects;
    const int32_t *obj_start, *obj_identity;
    const float *obj_R, *obj_pos;
    int32_t n_spheres;
    const float *sph_center, *sph_radius;
    int32_t n_boxes;
    const int32_t *obj_box_start;
    const float *box_center, *box_half, *box_radius;
    const int32_t *box_rounded;
    int32_t grid_obj;                  /* voxel SDF object (grid_map_sdf.py), -1 = none */
    int32_t grid_n[3];
    float grid_lo[3], grid_extent[3];
    const float *grid_data;            /* HOST [n0][n1][n2][4] = (sdf, grad) */
    int32_t ws_enabled;
    float ws_min[3], ws_max[3];
    int32_t n_pairs;
    const int32_t *pair_i, *pair_j;
    const float *pair_margin;
    int32_t ee_enabled, ee_link;
    float ee_w_pos, ee_w_rot;
    int32_t ee_square;
    float ee_target[16];
    int32_t clamp_self, clamp_obj, clamp_ws;
    float weights[4];
} orc_cost;

static float norm3(float x, float y, float z) {
    float s = x * x;
    s = fmaf(y, y, s);
    s = fmaf(z, z, s);
    return sqrtf(s);
}

/* one box primitive (primitives.py:219-223 sharp, :326-334 rounded): signed distance, optionally its gradient */
/* cell index along one axis: clamp(int(floor((x - lo) / extent * n)), 0, n - 1) in fp32 (grid_map_sdf.py:96-101) */
static int grid_cell(float x, const orc_cost *k, int a) {
    const float t = floorf(((x - k->grid_lo[a]) / k->grid_extent[a]) * (float)k->grid_n[a]);
    if (!(t >= 0.f)) return 0;
    if (t >= (float)k->grid_n[a]) return k->grid_n[a] - 1;
    return (int)t;
}

static float box_sdf(const orc_cost *k, int kb, const float *x, float *g) {
    const float *c = k->box_center + 3 * kb, *hh = k->box_half + 3 * kb;
    const float r = k->box_radius[kb];
    const int rounded = k->box_rounded[kb];
    float v[3], q[3], w[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < 3; ++i) { v[i] = x[i] - c[i]; q[i] = fabsf(v[i]) - hh[i]; if (rounded) q[i] = q[i] + r; }
    const float mq = fmaxf(fmaxf(q[0], q[1]), q[2]);
    float sd = mq;
    int max_active = 1;
    if (rounded) {
        const float rq[3] = {fmaxf(q[0], 0.f), fmaxf(q[1], 0.f), fmaxf(q[2], 0.f)};
        const float nrm = norm3(rq[0], rq[1], rq[2]);
        sd = (fminf(mq, 0.f) + nrm) - r;
        if (nrm > 0.f) for (int i = 0; i < 3; ++i) w[i] = rq[i] / nrm;
        max_active = mq < 0.f;
    }
    if (g) {
        if (max_active) { if (q[0] == mq) w[0] += 1.f; else if (q[1] == mq) w[1] += 1.f; else w[2] += 1.f; }
        for (int i = 0; i < 3; ++i) g[i] = (v[i] > 0.f ? 1.f : (v[i] < 0.f ? -1.f : 0.f)) * w[i];
    }
    return sd;
}

/* rigid_body.py:146-211 for one configuration; W[b] = R (9) | p (3); pass[b] = clamp passes gradient */
static void fk_one(const orc_tree *t, const float *q, float W[][12], int *pass) {
    for (int i = 0; i < t->n_links; ++i) {
        const int b = t->topo[i], p = t->parent[b];
        pass[b] = 0;
        if (p < 0) {
            static const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
            memcpy(W[b], t->base_pose ? t->base_pose : ident, sizeof(ident));
            continue;
        }
        const float *Rf = t->R_fixed + 9 * b, *tr = t->trans + 3 * b;
        float theta = 0.f;
        if (t->q_index[b] >= 0) {
            theta = q[t->q_index[b]];
            pass[b] = 1;
            if (t->has_lim[b]) { /* torch.clamp: rigid_body.py:157-160 */
                pass[b] = (theta >= t->lo[b]) && (theta <= t->hi[b]);
                if (theta == theta) theta = fminf(fmaxf(theta, t->lo[b]), t->hi[b]);
            }
        }
        float Rl[9], tl[3] = {tr[0], tr[1], tr[2]};
        memcpy(Rl, Rf, sizeof(Rl));
        if (t->kind[b] == 1) { /* Rl = Rf @ Rot_k(s * theta): rigid_body.py:162-174 */
            const float a = t->axis_s[b] * theta, c = cosf(a), s = sinf(a);
            const int k = t->axis_k[b], ci = (k + 1) % 3, cj = (k + 2) % 3;
            for (int r = 0; r < 3; ++r) {
                const float u = Rf[r * 3 + ci], v = Rf[r * 3 + cj];
                Rl[r * 3 + ci] = c * u + s * v;
                Rl[r * 3 + cj] = c * v - s * u;
            }
        } else if (t->kind[b] == 2) { /* rigid_body.py:175-182: raw axis, not rotated */
            for (int r = 0; r < 3; ++r) tl[r] = tr[r] + t->axis[3 * b + r] * theta;
        }
        const float *Wp = W[p];
        for (int r = 0; r < 3; ++r) { /* geometrics/utils.py:14-15 */
            for (int c = 0; c < 3; ++c)
                W[b][r * 3 + c] = Wp[r * 3] * Rl[c] + Wp[r * 3 + 1] * Rl[3 + c] + Wp[r * 3 + 2] * Rl[6 + c];
            W[b][9 + r] = Wp[r * 3] * tl[0] + Wp[r * 3 + 1] * tl[1] + Wp[r * 3 + 2] * tl[2] + Wp[9 + r];
        }
    }
}

/* leaf->root wrench accumulation; F/T are consumed */
static void tree_backward_one(const orc_tree *t, float W[][12], float F[][3], float T[][3], const int *pass, float *grad) {
    for (int i = t->n_links - 1; i >= 0; --i) {
        const int b = t->topo[i], p = t->parent[b];
        if (p < 0) continue;
        const float *w = W[b];
        if (t->q_index[b] >= 0) {
            float g;
            if (t->kind[b] == 1) {
                const int k = t->axis_k[b];
                const float s = t->axis_s[b];
                const float mx = T[b][0] - (w[10] * F[b][2] - w[11] * F[b][1]);
                const float my = T[b][1] - (w[11] * F[b][0] - w[9] * F[b][2]);
                const float mz = T[b][2] - (w[9] * F[b][1] - w[10] * F[b][0]);
                g = s * w[k] * mx + s * w[3 + k] * my + s * w[6 + k] * mz;
            } else {
                const float *Wp = W[p], *a = t->axis + 3 * b;
                g = 0.f;
                for (int r = 0; r < 3; ++r) g += (Wp[r * 3] * a[0] + Wp[r * 3 + 1] * a[1] + Wp[r * 3 + 2] * a[2]) * F[b][r];
            }
            grad[t->q_index[b]] = pass[b] ? g : 0.f;
        }
        for (int r = 0; r < 3; ++r) { F[p][r] += F[b][r]; T[p][r] += T[b][r]; }
    }
}

static void axial_gR_Rt(const float *gR, const float *R, float *a) {
    float M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = gR[i * 3] * R[j * 3] + gR[i * 3 + 1] * R[j * 3 + 1] + gR[i * 3 + 2] * R[j * 3 + 2];
    a[0] = M[7] - M[5]; a[1] = M[2] - M[6]; a[2] = M[3] - M[1];
}

void orc_fk_forward(const orc_tree *t, const float *q, int64_t n, const int32_t *sel, int32_t n_sel, float *H) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n; ++c) {
        float W[MAXL][12];
        int pass[MAXL];
        fk_one(t, q + c * t->n_dofs, W, pass);
        for (int s = 0; s < n_sel; ++s) {
            const float *w = W[sel ? sel[s] : s];
            float *h = H + (c * n_sel + s) * 16;
            for (int r = 0; r < 3; ++r) { h[r * 4] = w[r * 3]; h[r * 4 + 1] = w[r * 3 + 1]; h[r * 4 + 2] = w[r * 3 + 2]; h[r * 4 + 3] = w[9 + r]; }
            h[12] = h[13] = h[14] = 0.f; h[15] = 1.f;
        }
    }
}

void orc_fk_backward(const orc_tree *t, const float *q, int64_t n, const int32_t *sel, int32_t n_sel, const float *G, float *grad_q) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n; ++c) {
        float W[MAXL][12], F[MAXL][3], T[MAXL][3];
        int pass[MAXL];
        fk_one(t, q + c * t->n_dofs, W, pass);
        memset(F, 0, sizeof(F));
        memset(T, 0, sizeof(T));
        for (int s = 0; s < n_sel; ++s) {
            const int b = sel ? sel[s] : s;
            const float *g = G + (c * n_sel + s) * 16, *w = W[b];
            const float gR[9] = {g[0], g[1], g[2], g[4], g[5], g[6], g[8], g[9], g[10]};
            const float gp[3] = {g[3], g[7], g[11]};
            float a[3];
            axial_gR_Rt(gR, w, a);
            F[b][0] += gp[0]; F[b][1] += gp[1]; F[b][2] += gp[2];
            T[b][0] += w[10] * gp[2] - w[11] * gp[1] + a[0];
            T[b][1] += w[11] * gp[0] - w[9] * gp[2] + a[1];
            T[b][2] += w[9] * gp[1] - w[10] * gp[0] + a[2];
        }
        tree_backward_one(t, W, F, T, pass, grad_q + c * t->n_dofs);
    }
}

/* one configuration of the objective; P = body positions accessor via W */
static void objective_one(const orc_tree *t, const orc_cost *k, const float *q, const float *Htgt, int use_mo, float mo,
                          float *terms, float *total, float *grad, uint8_t *occ_out, int32_t *argmin) {
    float W[MAXL][12], F[MAXL][3], T[MAXL][3];
    int pass[MAXL];
    fk_one(t, q, W, pass);
    memset(F, 0, sizeof(F));
    unsigned occ = 0;
    const float *wt = k->weights;
    /* self pairs: distance_fields.py:194-215 */
    float c_self = 0.f;
    for (int p = 0; p < k->n_pairs; ++p) {
        const int i = k->pair_i[p], j = k->pair_j[p];
        const float vx = W[i][9] - W[j][9], vy = W[i][10] - W[j][10], vz = W[i][11] - W[j][11];
        const float d = norm3(vx, vy, vz), m = k->pair_margin[p];
        if (d < (use_mo ? mo : m)) occ |= 1u;
        float val = -(d - m);
        int act = 1;
        if (k->clamp_self) { act = val > 0.f; val = fmaxf(val, 0.f); }
        c_self = c_self + val;
        if (act && d > 0.f) {
            const float s = wt[0] / d;
            F[i][0] -= vx * s; F[i][1] -= vy * s; F[i][2] -= vz * s;
            F[j][0] += vx * s; F[j][1] += vy * s; F[j][2] += vz * s;
        }
    }
    float c_obj = 0.f, c_ws = 0.f;
    const int S = k->n_points;
    for (int s = 0; s < S; ++s) {
        const int b = k->pt_link[s];
        const float x = W[b][9], y = W[b][10], z = W[b][11];
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (k->n_objects > 0) { /* distance_fields.py:307-316, primitives.py:387-405, :108-112 */
            const float m = k->pt_margin_obj[s];
            float vmax = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
            for (int o = 0; o < k->n_objects; ++o) {
                float xo = x, yo = y, zo = z;
                const float *Ro = k->obj_R + 9 * o;
                if (!k->obj_identity[o]) {
                    const float d0 = x - k->obj_pos[3 * o], d1 = y - k->obj_pos[3 * o + 1], d2 = z - k->obj_pos[3 * o + 2];
                    xo = fmaf(d2, Ro[6], fmaf(d1, Ro[3], d0 * Ro[0]));
                    yo = fmaf(d2, Ro[7], fmaf(d1, Ro[4], d0 * Ro[1]));
                    zo = fmaf(d2, Ro[8], fmaf(d1, Ro[5], d0 * Ro[2]));
                }
                float best = INFINITY, bvx = 0.f, bvy = 0.f, bvz = 0.f, bd = 1.f;
                int bidx = 0;
                for (int kk = k->obj_start[o]; kk < k->obj_start[o + 1]; ++kk) {
                    const float vx = xo - k->sph_center[3 * kk], vy = yo - k->sph_center[3 * kk + 1], vz = zo - k->sph_center[3 * kk + 2];
                    const float d = norm3(vx, vy, vz), sd = d - k->sph_radius[kk];
                    if (sd < best) { best = sd; bidx = kk - k->obj_start[o]; bvx = vx; bvy = vy; bvz = vz; bd = d; }
                }
                int bbox = -1;
                const float xob[3] = {xo, yo, zo};
                if (k->n_boxes > 0)
                    for (int kb = k->obj_box_start[o]; kb < k->obj_box_start[o + 1]; ++kb) {
                        const float sd = box_sdf(k, kb, xob, 0);
                        if (sd < best) { best = sd; bbox = kb; bidx = (k->obj_start[o + 1] - k->obj_start[o]) + (kb - k->obj_box_start[o]); }
                    }
                const float *cell = 0;
                if (o == k->grid_obj) { /* grid_map_sdf.py:96-112: cell lookup, stored gradient */
                    bidx = (grid_cell(xo, k, 0) * k->grid_n[1] + grid_cell(yo, k, 1)) * k->grid_n[2] + grid_cell(zo, k, 2);
                    cell = k->grid_data + 4 * (int64_t)bidx;
                    best = cell[0];
                }
                if (argmin) argmin[o * S + s] = bidx;
                if (best < (use_mo ? mo : m)) occ |= 2u;
                float val = -(best - m);
                int act = 1;
                if (k->clamp_obj) { act = val > 0.f; val = fmaxf(val, 0.f); }
                if (o == 0 || val > vmax) {
                    vmax = val;
                    if (act && (bd > 0.f || bbox >= 0 || cell)) {
                        float ux = 0.f, uy = 0.f, uz = 0.f;
                        if (cell) { ux = cell[1]; uy = cell[2]; uz = cell[3]; }
                        else if (bbox >= 0) { float gb[3]; box_sdf(k, bbox, xob, gb); ux = gb[0]; uy = gb[1]; uz = gb[2]; }
                        else { ux = bvx / bd; uy = bvy / bd; uz = bvz / bd; }
                        if (!k->obj_identity[o]) {
                            const float wx = Ro[0] * ux + Ro[1] * uy + Ro[2] * uz, wy = Ro[3] * ux + Ro[4] * uy + Ro[5] * uz,
                                        wz = Ro[6] * ux + Ro[7] * uy + Ro[8] * uz;
                            ux = wx; uy = wy; uz = wz;
                        }
                        dx = ux; dy = uy; dz = uz;
                    } else { dx = dy = dz = 0.f; }
                }
            }
            c_obj = c_obj + vmax;
            gx -= wt[1] * dx; gy -= wt[1] * dy; gz -= wt[1] * dz;
        }
        if (k->ws_enabled) { /* distance_fields.py:326-332 */
            const float m = k->pt_margin_ws[s];
            const float d[6] = {x - k->ws_min[0], y - k->ws_min[1], z - k->ws_min[2], k->ws_max[0] - x, k->ws_max[1] - y, k->ws_max[2] - z};
            float vmax = 0.f, dsel = 0.f;
            int imax = 0, amax = 1;
            for (int i = 0; i < 6; ++i) {
                if (d[i] < (use_mo ? mo : m)) occ |= 4u;
                float val = -(d[i] - m);
                int act = 1;
                if (k->clamp_ws) { act = val > 0.f; val = fmaxf(val, 0.f); }
                if (i == 0 || val > vmax) { vmax = val; imax = i; amax = act; dsel = d[i]; }
            }
            c_ws = c_ws + vmax;
            if (amax && dsel != 0.f) {
                const float g = imax < 3 ? -wt[2] : wt[2];
                const int a = imax < 3 ? imax : imax - 3;
                if (a == 0) gx += g; else if (a == 1) gy += g; else gz += g;
            }
        }
        F[b][0] += gx; F[b][1] += gy; F[b][2] += gz;
    }
    float c_ee = 0.f, gR[9] = {0}, gp[3] = {0};
    if (k->ee_enabled) { /* geometrics/utils.py:130-154; distance_fields.py:347-356 */
        const float *w = W[k->ee_link];
        float Rt[9], pt[3];
        for (int i = 0; i < 3; ++i) { Rt[i * 3] = Htgt[i * 4]; Rt[i * 3 + 1] = Htgt[i * 4 + 1]; Rt[i * 3 + 2] = Htgt[i * 4 + 2]; pt[i] = Htgt[i * 4 + 3]; }
        float D = 0.f;
        if (k->ee_w_rot > 0.f) {
            float dg[3];
            for (int i = 0; i < 3; ++i) dg[i] = fmaf(w[i * 3 + 2], Rt[i * 3 + 2], fmaf(w[i * 3 + 1], Rt[i * 3 + 1], w[i * 3] * Rt[i * 3]));
            const float tr = (dg[0] + dg[1]) + dg[2];
            D = k->ee_w_rot * (1.f - (tr - 1.f) * 0.5f);
        }
        float vx = 0.f, vy = 0.f, vz = 0.f, dist = 0.f;
        if (k->ee_w_pos > 0.f) {
            vx = w[9] - pt[0]; vy = w[10] - pt[1]; vz = w[11] - pt[2];
            dist = norm3(vx, vy, vz);
            D = D + k->ee_w_pos * dist;
        }
        c_ee = k->ee_square ? D * D : D;
        const float scale = wt[3] * (k->ee_square ? 2.f * D : 1.f);
        if (k->ee_w_rot > 0.f) for (int i = 0; i < 9; ++i) gR[i] = scale * (-0.5f * k->ee_w_rot) * Rt[i];
        if (k->ee_w_pos > 0.f && dist > 0.f) { const float kp = scale * k->ee_w_pos / dist; gp[0] = kp * vx; gp[1] = kp * vy; gp[2] = kp * vz; }
    }
    if (terms) { terms[0] = c_self; terms[1] = c_obj; terms[2] = c_ws; terms[3] = c_ee; }
    if (total) *total = ((wt[0] * c_self + wt[1] * c_obj) + wt[2] * c_ws) + wt[3] * c_ee;
    if (occ_out) *occ_out = (uint8_t)occ;
    if (grad) {
        for (int b = 0; b < t->n_links; ++b) {
            const float *w = W[b];
            T[b][0] = w[10] * F[b][2] - w[11] * F[b][1];
            T[b][1] = w[11] * F[b][0] - w[9] * F[b][2];
            T[b][2] = w[9] * F[b][1] - w[10] * F[b][0];
        }
        if (k->ee_enabled) {
            const int e = k->ee_link;
            const float *w = W[e];
            float a[3];
            axial_gR_Rt(gR, w, a);
            F[e][0] += gp[0]; F[e][1] += gp[1]; F[e][2] += gp[2];
            T[e][0] += w[10] * gp[2] - w[11] * gp[1] + a[0];
            T[e][1] += w[11] * gp[0] - w[9] * gp[2] + a[1];
            T[e][2] += w[9] * gp[1] - w[10] * gp[0] + a[2];
        }
        tree_backward_one(t, W, F, T, pass, grad);
    }
}

/* margin_override: NaN = default margins */
void orc_objective(const orc_tree *t, const orc_cost *k, const float *q, int64_t n, const float *ee_target_cfg,
                   float margin_override, float *total, float *terms, float *grad_q, uint8_t *occ, int32_t *argmin) {
    const int use_mo = !(margin_override != margin_override);
    const int64_t am_stride = (int64_t)k->n_objects * k->n_points;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n; ++c)
        objective_one(t, k, q + c * t->n_dofs, ee_target_cfg ? ee_target_cfg + c * 16 : k->ee_target, use_mo, margin_override,
                      terms ? terms + c * 4 : 0, total ? total + c : 0, grad_q ? grad_q + c * t->n_dofs : 0, occ ? occ + c : 0,
                      argmin ? argmin + c * am_stride : 0);
}

#ifdef _OPENMP
#include <omp.h>
int orc_max_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { omp_set_num_threads(n); }
#else
int orc_max_threads(void) { return 1; }
void orc_set_threads(int n) { (void)n; }
#endif
