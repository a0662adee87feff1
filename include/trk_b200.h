/*
 * trk_b200.h - C ABI of the B200-native FK + planning-objective hot path.
 *
 * The reference (anindex/torch_robotics) is pure Python on PyTorch and has no FFI of its own;
 * its boundary for this path is a set of Python methods.  Every entry point below names the
 * reference method it stands behind (paths relative to /root/reference/torch_robotics).  The
 * Python drop-in layer (torch_robotics_b200/) binds these symbols with ctypes and wraps them in
 * torch.autograd.Function; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross the boundary;
 *   - every data pointer is a DEVICE pointer to contiguous fp32/int32/uint8 memory owned by the
 *     caller, unless the parameter is documented as host memory;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*): no allocation, no
 *     synchronisation, no host read-back on the data path;
 *   - return value: 0 = OK, <0 = invalid argument (TRK_E_*), >0 = cudaError_t;
 *     trk_last_error() returns a thread-local message for the last non-zero status;
 *   - handles are immutable after creation (except the documented setters) and may be shared
 *     between host threads; one handle per (model, device).
 */
#ifndef TRK_B200_H
#define TRK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRK_ABI_VERSION 1

#define TRK_MAX_LINKS   64
#define TRK_MAX_DOFS    32
#define TRK_MAX_POINTS  64
#define TRK_MAX_SPHERES 256
#define TRK_MAX_OBJECTS 16
#define TRK_MAX_PAIRS   128
#define TRK_MAX_BOXES   64

#define TRK_OK              0
#define TRK_E_NULL         -1   /* required pointer is null */
#define TRK_E_RANGE        -2   /* count / index out of range */
#define TRK_E_ALIGN        -3   /* pointer not aligned as required */
#define TRK_E_UNSUPPORTED  -4   /* valid request the library does not implement */

#define TRK_KIND_FIXED      0
#define TRK_KIND_REVOLUTE   1   /* also 'continuous' (has_lim = 0) */
#define TRK_KIND_PRISMATIC  2

/* occupancy bits written by trk_objective / trk_cost_stage */
#define TRK_OCC_SELF   1u
#define TRK_OCC_OBJECT 2u
#define TRK_OCC_WS     4u

typedef struct trk_model trk_model;   /* opaque: kinematic tree resident on one GPU */
typedef struct trk_cost  trk_cost;    /* opaque: planning-objective tables resident on one GPU */

/*
 * Kinematic tree, one entry per body in URDF <link> order (HOST arrays, copied at creation).
 * Produced by torch_robotics_b200/descriptor.py:compile_urdf, which restates
 *   torch_kinematics_tree/models/utils.py:199-313   (URDFRobotModel.get_body_parameters)
 *   torch_kinematics_tree/models/rigid_body.py:74-123 (DifferentiableRigidBody.__init__)
 *   torch_kinematics_tree/models/robot_tree.py:77-126 (DifferentiableTree.__init__).
 */
typedef struct trk_model_desc {
    int32_t        n_links;   /* L <= TRK_MAX_LINKS */
    int32_t        n_dofs;    /* D <= TRK_MAX_DOFS  */
    const int32_t* parent;    /* [L] body index of the parent, -1 for the root (body 0) */
    const int32_t* kind;      /* [L] TRK_KIND_* */
    const int32_t* axis_k;    /* [L] revolute: rotation axis 0=x 1=y 2=z (rigid_body.py:163-168) */
    const float*   axis_s;    /* [L] revolute: sign multiplying q (-1, 0, +1) */
    const float*   axis;      /* [L*3] raw <axis xyz>; prismatic offset = axis * q (rigid_body.py:176) */
    const int32_t* q_index;   /* [L] column of q driving this body, -1 if fixed */
    const float*   R_fixed;   /* [L*9] row-major joint-origin rotation (rigid_body.py:93) */
    const float*   trans;     /* [L*3] joint-origin translation */
    const float*   lo;        /* [L] clamp bounds, fp32(lower/upper) (rigid_body.py:157-160) */
    const float*   hi;
    const int32_t* has_lim;   /* [L] 1 when FK clamps q to [lo, hi] */
    const int32_t* topo;      /* [L] processing order: every parent precedes its children */
    const float*   base_pose; /* [12] row-major [R | t] world pose of the root, NULL = identity
                                 (DifferentiableTree.update_base_pose, robot_tree.py:133) */
} trk_model_desc;

/*
 * Planning objective evaluated on the FK result (HOST arrays, copied at creation).
 * Restates what PlanningTask wires together, tasks/tasks.py:44-86, from
 *   robots/robot_base.py:57-141 + robots/robot_panda.py:47-127      (collision model)
 *   torch_planning_objectives/fields/distance_fields.py:107-130,180-215,269-356 (cost fields)
 *   environments/primitives.py:90-112,348-405                       (sphere SDF, object pose)
 */
typedef struct trk_cost_desc {
    int32_t        n_points;       /* S collision points, each at the origin of body pt_link[s] */
    const int32_t* pt_link;        /* [S] */
    const float*   pt_margin_obj;  /* [S] fp32(radius + cutoff) for the object term (distance_fields.py:113) */
    const float*   pt_margin_ws;   /* [S] fp32(radius + cutoff) for the workspace term */
    int32_t        n_objects;      /* ObjectFields; object o owns spheres [obj_start[o], obj_start[o+1]) */
    const int32_t* obj_start;      /* [n_objects + 1] */
    const int32_t* obj_identity;   /* [n_objects] 1 = identity pose: points are used untransformed */
    const float*   obj_R;          /* [n_objects*9] world<-object rotation (primitives.py:397) */
    const float*   obj_pos;        /* [n_objects*3] */
    int32_t        n_spheres;      /* K */
    const float*   sph_center;     /* [K*3] in the object frame */
    const float*   sph_radius;     /* [K] */
    /* boxes (primitives.py:198-334): object o also owns boxes [obj_box_start[o], obj_box_start[o+1]); the object's
       signed distance is the minimum over its spheres and boxes.  n_boxes may be 0 (all box pointers then unused). */
    int32_t        n_boxes;        /* KB <= TRK_MAX_BOXES */
    const int32_t* obj_box_start;  /* [n_objects + 1] */
    const float*   box_center;     /* [KB*3] in the object frame */
    const float*   box_half;       /* [KB*3] half sizes */
    const float*   box_radius;     /* [KB] rounding radius of MultiBoxField (:309-334); ignored for sharp boxes */
    const int32_t* box_rounded;    /* [KB] 1 = MultiBoxField (rounded), 0 = MultiSharpBoxField */
    /* precomputed voxel SDF (environments/grid_map_sdf.py:9-114; EnvBase(precompute_sdf_obj_fixed=True), env_base.py:59-65):
       at most one object may be a grid instead of primitives.  grid_obj = its object index (the object then owns no
       spheres or boxes and must have the identity pose), -1 = none.  Cell of a point x (grid_map_sdf.py:96-101):
       i = clamp(int(floor((x - grid_lo) / grid_extent * grid_n)), 0, grid_n - 1) per axis, in fp32.  grid_data is a
       DEVICE pointer to [n0, n1, n2, 4] fp32 = (sdf, d sdf/dx, d sdf/dy, d sdf/dz) per cell, 16-byte aligned, owned by
       the caller and kept alive as long as the cost handle; the value is sdf[cell], its gradient the stored one (the
       reference's surrogate, grid_map_sdf.py:103-112). */
    int32_t        grid_obj;
    int32_t        grid_n[3];
    float          grid_lo[3];
    float          grid_extent[3];
    const float*   grid_data;
    int32_t        ws_enabled;
    float          ws_min[3];      /* distance_fields.py:326-332 */
    float          ws_max[3];
    int32_t        n_pairs;        /* P self-collision pairs (body indices), robot_base.py:105-116 */
    const int32_t* pair_i;
    const int32_t* pair_j;
    const float*   pair_margin;    /* [P] */
    int32_t        ee_enabled;
    int32_t        ee_link;        /* body whose pose enters SE3_distance (distance_fields.py:347-350) */
    float          ee_w_pos, ee_w_rot;
    int32_t        ee_square;
    float          ee_target[16];  /* row-major 4x4 shared target (EESE3DistanceField.target_H) */
    int32_t        clamp_self, clamp_obj, clamp_ws;   /* relu(margin - sdf), distance_fields.py:117 */
    float          weights[4];     /* total = w0*self + w1*object + w2*ws + w3*ee */
} trk_cost_desc;

int         trk_abi_version(void);
const char* trk_last_error(void);

/* Replaces DifferentiableTree.__init__ (robot_tree.py:77-126) for the device side. */
int  trk_model_create(const trk_model_desc* desc, int device, trk_model** out);
void trk_model_destroy(trk_model* model);
/* Name of the kernel family serving this model: "generic" or a compiled-in robot such as "panda". */
const char* trk_model_kernel_family(const trk_model* model);

/*
 * DifferentiableTree.compute_forward_kinematics_all_links (robot_tree.py:267-301).
 *   q        [n, D]  device fp32, row-major
 *   link_sel [n_sel] HOST int32 body indices in output order, each body at most once; NULL = all bodies
 *   H_out    [n, n_sel, 4, 4] device fp32, 16-byte aligned
 */
int trk_fk_forward(const trk_model* model, const float* q, int64_t n,
                   const int32_t* link_sel, int32_t n_sel, float* H_out, void* stream);

/*
 * Reverse mode of the above (the reference uses autograd over rigid_body.py:146-211).
 * Recomputes FK from q; G [n, n_sel, 4, 4] is the upstream gradient (bottom rows ignored);
 * grad_q [n, D] is overwritten.
 */
int trk_fk_backward(const trk_model* model, const float* q, int64_t n,
                    const int32_t* link_sel, int32_t n_sel, const float* G, float* grad_q, void* stream);

/* Device tables for PlanningTask's fields (tasks.py:44-86) + EESE3DistanceField (distance_fields.py:335). */
int  trk_cost_create(const trk_model* model, const trk_cost_desc* desc, trk_cost** out);
void trk_cost_destroy(trk_cost* cost);
/* EESE3DistanceField.update_target (distance_fields.py:344): H_target16 is HOST memory, row-major 4x4. */
int  trk_cost_set_ee_target(trk_cost* cost, const float* H_target16, void* stream);
/* Replace the four term weights (HOST memory). */
int  trk_cost_set_weights(trk_cost* cost, const float* weights4, void* stream);

/*
 * Fused FK + objective + reverse mode:
 *   PlanningTask._compute_collision_or_cost (tasks.py:139-232) + EESE3DistanceField.compute_costs_impl
 *   (distance_fields.py:352-356) + cost.sum().backward().
 *   q             [n, D]
 *   ee_target_cfg [n, 4, 4] per-configuration EE targets, or NULL to use the shared target
 *   margin_override  occupancy threshold replacing every per-point/pair margin (the `margin=` kwarg,
 *                    tasks.py:251); pass NaN for the default margins
 *   cost_total [n]        weighted sum of the terms                      (nullable)
 *   cost_terms [n, 4]     self, object, ws, ee                           (nullable)
 *   grad_q     [n, D]     d cost_total / d q                             (nullable: forward only)
 *   occupancy  [n] uint8  TRK_OCC_* bits (compute_collision, tasks.py:131) (nullable)
 *   argmin     [n, n_objects, S] int32 nearest primitive per object and point, index within the object
 *                             (its spheres first, then its boxes) (nullable)
 */
int trk_objective(const trk_model* model, const trk_cost* cost, const float* q, int64_t n,
                  const float* ee_target_cfg, float margin_override,
                  float* cost_total, float* cost_terms, float* grad_q,
                  uint8_t* occupancy, int32_t* argmin, void* stream);

/*
 * trk_objective with the first stage of the cost reduction fused into the kernel (the caller of PlanningTask sums the
 * cost over the batch and, across GPUs, all-reduces the scalar: SURVEY.md section 8e).  cost_total (nullable) and grad_q as
 * above; block_sums [TRK_MAX_BLOCK_SUMS] fp32 DEVICE memory receives one partial sum of the total cost per thread block,
 * in block order, zero beyond the launch's grid: summing the array (fp64, any fixed order) gives sum(cost_total),
 * deterministic for a given batch size.  Returns TRK_E_UNSUPPORTED when the model / cost table is not served by the
 * headline kernel (trk_hot.cuh); sum cost_total instead.
 */
#define TRK_MAX_BLOCK_SUMS 8192
int trk_objective_sums(const trk_model* model, const trk_cost* cost, const float* q, int64_t n, const float* ee_target_cfg,
                       float margin_override, float* cost_total, float* grad_q, float* block_sums, void* stream);

/*
 * Host-buffer form of trk_objective: q, cost_total and grad_q are HOST pointers (page-locked memory gives
 * asynchronous, overlapped copies; pageable memory works but serialises).  The batch is cut into `n_chunks`
 * pieces that flow through three internal streams - H2D copy, kernel, D2H copy - so both PCIe directions and the
 * compute overlap.  Device staging buffers live in the cost handle and are grown on demand (the only
 * allocation this library performs after handle creation).  The call returns when the results are in host
 * memory.  Not re-entrant for one cost handle.
 *   q_host [n, D], cost_total_host [n] (nullable), grad_q_host [n, D] (nullable: forward only)
 */
int trk_objective_host(const trk_model* model, trk_cost* cost, const float* q_host, int64_t n, int32_t n_chunks,
                       float* cost_total_host, float* grad_q_host);

/*
 * Cost stage alone, on caller-supplied link positions - the DistanceField.compute_cost entry
 * (distance_fields.py:26-55) which receives FK output as a tensor.
 *   link_pos [n, L, 3]  world positions of all bodies (RobotBase.fk_map_collision, robot_base.py:171)
 *   ee_pose  [n, 4, 4]  pose of the EE body (nullable unless the EE term is enabled)
 *   grad_pos [n, L, 3]  d cost_total / d link_pos  (nullable)
 *   grad_ee  [n, 4, 4]  d (w3 * ee term) / d ee_pose (nullable)
 */
int trk_cost_stage(const trk_model* model, const trk_cost* cost, const float* link_pos, const float* ee_pose,
                   int64_t n, const float* ee_target_cfg, float margin_override,
                   float* cost_total, float* cost_terms, float* grad_pos, float* grad_ee,
                   uint8_t* occupancy, int32_t* argmin, void* stream);

/*
 * Fused batched Adam inverse kinematics: DifferentiableTree.inverse_kinematics (robot_tree.py:303-384) with
 * loss_fn_ik_per_q (:386-417) and ik_termination (:419-442) evaluated inside one kernel; q and the Adam
 * moments stay in registers for all iterations.  Iteration `it` (0-based) first tests termination on the
 * current q, then - if it < n_updates - applies one torch.optim.Adam step (default betas/eps semantics).
 * Calling with n_iters = n_updates = max_iters reproduces the reference loop that never breaks early; the
 * per-iteration count of valid seeds lets the host detect the iteration at which every seed was valid
 * (robot_tree.py:356) and re-run with n_iters = it + 1, n_updates = it.
 *   step_size [n_iters]  device fp32: lr / (1 - beta1^(it+1))
 *   bc2_sqrt  [n_iters]  device fp32: sqrt(1 - beta2^(it+1))
 *   target_seed [n, 4, 4] device fp32 per-seed targets, or NULL to use params->H_target
 *   q_io      [n, D]     device fp32: seeds in, solutions out
 *   err_out   [n]        SE3_distance at the last termination test            (nullable)
 *   valid_out [n] uint8  result of the last termination test                  (nullable)
 *   valid_count [n_iters] device int32: seeds passing the test at each iteration (zeroed by the call)
 *   loss_sum / loss_sqsum [n_iters] device fp32 sums of the loss and its square (nullable, zeroed by the call)
 * Compiled-in robots keep q and the Adam moments in registers; models served by the runtime-tree family keep them
 * in shared-memory columns (same arithmetic).
 */
typedef struct trk_ik_params {
    int32_t      ee_link;          /* body whose pose is driven to the target */
    int32_t      n_iters;          /* termination tests performed */
    int32_t      n_updates;        /* Adam steps performed (<= n_iters) */
    float        lr, beta1, beta2, eps;
    float        w_se3, w_joint_limits, se3_eps;
    const float* lower;            /* HOST [D] limits used by the loss / termination test */
    const float* upper;            /* HOST [D] */
    const float* H_target;         /* HOST 16 floats (row-major 4x4), may be NULL when target_seed is given */
} trk_ik_params;

int trk_ik_adam(const trk_model* model, const trk_ik_params* params, const float* step_size, const float* bc2_sqrt,
                const float* target_seed, int64_t n, float* q_io, float* err_out, uint8_t* valid_out,
                int32_t* valid_count, float* loss_sum, float* loss_sqsum, void* stream);

/*
 * trk_ik_adam with the first stage of the best-solution selection fused into the kernel (robot_tree.py:303-384 returns every
 * seed; a caller that wants the k best valid ones - BASELINE config 4: per-rank top-k, all-gather, top-k - otherwise runs a
 * select over all seeds).  Seeds are processed in blocks of TRK_IK_BLOCK_SEEDS; best_keys [ceil(n / TRK_IK_BLOCK_SEEDS) * best_k]
 * int64 DEVICE memory receives, per block and ascending, the best_k smallest keys
 *     (bits of the final SE(3) error, or of +inf for a seed that fails the termination test) << 32 | seed index,
 * 0x7fffffffffffffff where a block has fewer seeds.  The k smallest keys overall are the k best seeds, ties by index.
 * Compiled-in robots with the EE on their last body only; TRK_E_UNSUPPORTED otherwise.
 */
#define TRK_IK_BLOCK_SEEDS 128
int trk_ik_adam_best(const trk_model* model, const trk_ik_params* params, const float* step_size, const float* bc2_sqrt,
                     const float* target_seed, int64_t n, float* q_io, float* err_out, uint8_t* valid_out,
                     int32_t* valid_count, int64_t* best_keys, int32_t best_k, void* stream);

/*
 * Replaces DifferentiableTree.compute_analytical_jacobian_all_links (robot_tree.py:250-265), which differentiates
 * FK + rotation_matrix_to_q (geometrics/quaternion.py:135-166) with torch.autograd.functional.jacobian.
 * q [N, D] -> J [N, L, 7, D]: rows 0-2 d position / d q, rows 3-6 d quaternion (w, x, y, z) / d q, bodies in <link>
 * order; columns of clamped joints and of joints not on the path root -> body are zero.  The quaternion is the
 * representative rotation_matrix_to_q returns (its largest-magnitude component positive).
 */
int trk_link_jacobian(const trk_model* model, const float* q, int64_t n, float* J, void* stream);

/*
 * Replaces DifferentiableTree.compute_forward_kinematics_and_geometric_jacobian (robot_tree.py:218-248) for one link:
 * q [N, D] -> pose12 [N, 12] (world rotation of `link` row-major, then its position), lin [N, 3, D], ang [N, 3, D] with
 *   lin[:, c] = z_c x (p_link - p_c), ang[:, c] = z_c   for the q columns set in `col_mask`, zero elsewhere,
 * z_c / p_c = axis column / position of the body q column c drives (joint limits clamped as in trk_fk_forward).  The
 * reference includes a column when the joint's index in the URDF is not larger than the index of the joint above `link`
 * (robot_tree.py:237-240); the caller passes that selection as the mask.
 */
int trk_geometric_jacobian(const trk_model* model, const float* q, int64_t n, int32_t link, uint32_t col_mask, float* pose12,
                           float* lin, float* ang, void* stream);

/*
 * Self-test of the branch-free correctly-rounded square root used for every distance that feeds a
 * comparison (trk_device.cuh: sqrt_rn_core): y_core[i] from that routine, y_ref[i] from sqrt.rn.f32.
 * The two must be bit-identical for x == 0 and 2^-100 <= x < 2^126.  All pointers are device memory.
 */
int trk_selftest_sqrt(const float* x, int64_t n, float* y_core, float* y_ref, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRK_B200_H */
