/*
 * ikl.h -- C ABI of libikl.so: the B200 (sm_100a) implementation of the inverse-kinematics
 * Lagrangian hot path of miloknowles/lagrange-dual-learning:
 *
 *     batched planar-arm forward kinematics  ->  joint-limit / circular-obstacle penalties
 *     ->  per-constraint batch sums g[K]  ->  Lagrangian  sum_i lambda_i g_i  and its analytic
 *     gradient d/d(theta)  ->  dual-ascent multiplier update.
 *
 * The reference has no FFI of its own (it is pure Python on PyTorch eager ops); each entry point
 * below states the reference Python interface it replaces (file:line relative to the reference
 * repo).  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - Plain C: raw pointers, sizes and element strides.  No torch / C++ types cross this boundary.
 *   - Every data pointer is a DEVICE pointer valid on `stream` unless the name ends in `_host`.
 *   - Strides are in ELEMENTS (floats), may be 0 (broadcast) and need not be contiguous; the
 *     library picks a vectorised kernel when the strides describe one of the two native layouts
 *     (see IklBatch) and a generic strided kernel otherwise.  Results do not depend on the choice.
 *   - Ownership: the caller owns every buffer including the workspace.  The library never
 *     allocates, frees or synchronises; all work is ordered on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).  Calls are re-entrant across streams as long as each
 *     concurrent call uses its own workspace.
 *   - Errors: every function returns an IklStatus; nothing throws.  ikl_strerror() names it and
 *     ikl_last_cuda_error() returns the cudaError_t (as int) behind an IKL_ERR_CUDA on this thread.
 *   - There is no CPU implementation behind this ABI.  Without a CUDA device every compute entry
 *     point returns IKL_ERR_CUDA.
 */
#ifndef IKL_H_
#define IKL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IKL_ABI_VERSION 5
#define IKL_MAX_LINKS 3        /* reference trainer supports 3 (scripts/trainer.py:177-180); 2 is templated too */
#define IKL_MAX_OBSTACLES 4    /* reference hard-wires (B,4,4): scripts/trainer.py:218, kinematics/dataset.py:80 */
#define IKL_MAX_CONSTRAINTS (1 + IKL_MAX_LINKS + IKL_MAX_LINKS * IKL_MAX_OBSTACLES)   /* 16 */
#define IKL_MAX_RANKS 8        /* GPUs of one NVSwitch box that can take part in the in-kernel exchange of the K sums */

typedef enum IklStatus {
  IKL_OK = 0,
  IKL_ERR_INVALID_ARGUMENT = 1,   /* null pointer, negative size, K mismatch, negative slope ... */
  IKL_ERR_UNSUPPORTED = 2,        /* num_links / num_obstacles outside the templated range */
  IKL_ERR_CUDA = 3                /* a CUDA runtime call failed; see ikl_last_cuda_error() */
} IklStatus;

/* What a resources/cfg_*.json plus --penalty_good_slope/--penalty_bad_slope reduce to for this path
 * (scripts/trainer.py:71-99 constraint layout; :170 limits; :203-204, :237 slope assignment;
 * utils/constants.py:8-10 link lengths). */
typedef struct IklSpec {
  int32_t num_links;              /* J: 2 or 3 */
  int32_t enforce_joint_limits;   /* cfg["enforce_joint_limits"] */
  int32_t num_obstacles;          /* 0 when cfg["enforce_obstacles"] is false; else 1..4 */
  int32_t dynamic_obstacles;      /* 1: per-sample obstacles (IklBatch.obstacles); 0: static_obstacles below */
  float limit_lo, limit_hi;       /* cfg["static_constraints"]["joint_limits"] as float32 */
  float good_slope, bad_slope;    /* opt.penalty_good_slope / opt.penalty_bad_slope, both >= 0 */
  float link_length[IKL_MAX_LINKS];
  float static_obstacles[IKL_MAX_OBSTACLES][3];   /* (x, y, radius) rows, float32 of the JSON values */
} IklSpec;

/* Number of constraints / multipliers K = 1 + [JL] J + J * num_obstacles (scripts/trainer.py:72-97).
 * Order: [EE, JL_1..JL_J, (obstacle 0: j=0..J-1), (obstacle 1: ...), ...], where j indexes the FK
 * tuple, i.e. j = 0 is the END EFFECTOR tip and j = J-1 the first link's tip (trainer.py:231-240). */
int32_t ikl_num_constraints(const IklSpec* spec);

/* One batch of inputs to the Lagrangian, as strided views.
 *   theta    (B, J)        network output; element (b, j) at theta[b*theta_sb + j*theta_sj]
 *   target   (B, >=2)      q_ee_desired; only x (c=0) and y (c=1) are read (trainer.py:192)
 *   obstacles(B, n_obs, 3+) per-sample circles (x, y, radius) for dynamic configs, else NULL
 * Native layouts (vectorised, no re-layout pass):
 *   "rows"   the reference's own tensors: theta (B,J) contiguous, target (B,3) or (B,2) contiguous,
 *            obstacles (B,4,4) contiguous [x, y, r, pad]   (kinematics/dataset.py:80-85,143-155).
 *            B >= 512 with 16-byte aligned tensors selects the bulk-copy staged kernel (whole
 *            512-row tiles by cp.async.bulk); anything else the direct-load rows kernel.
 *   "planes" structure-of-arrays: every column its own contiguous plane of B floats (theta_sb == 1,
 *            theta_sj == plane pitch, ...).  Even B, 8-byte aligned planes and even pitches select the
 *            packed kernel; B % 4 == 0 with 16-byte aligned planes and pitches % 4 == 0 select the
 *            TMA-staged kernel (tiled tensor-map loads), the fastest path.
 * Environment switches for tests and measurements: IKL_FORCE_LAYOUT=strided|rows|planes,
 * IKL_PLANES_TMA=0 and IKL_ROWS_TMA=0 (direct-load kernels instead of the staged ones).
 */
typedef struct IklBatch {
  int64_t batch;                                   /* B >= 0 */
  const float* theta;      int64_t theta_sb, theta_sj;
  const float* target;     int64_t target_sb, target_sc;
  const float* obstacles;  int64_t obst_sb, obst_so, obst_sc;
} IklBatch;

/* Per-constraint weights of the Lagrangian.  Exactly one of the two pointers is non-NULL.
 *   host:   K floats in host memory, captured by value at launch (they ride in the kernel
 *           parameters / constant bank: cheapest; use when lambda is known on the host, which it is
 *           for the whole of train_with_relaxation, trainer.py:251-265);
 *   device: K floats in device memory, read by the kernel (no host sync needed). */
typedef struct IklWeights {
  const float* host;
  const float* device;
} IklWeights;

/* Bytes of device workspace any launch on a batch of this size may need.  The workspace must be
 * zero-filled once after allocation; the library leaves it consistent after every call. */
size_t ikl_workspace_bytes(int64_t batch);

/* Forward only: constraint sums and Lagrangian of one batch.
 * Replaces the FK/penalty part of IkLagrangeDualTrainer.process_batch (scripts/trainer.py:182-247)
 * as used under no_grad by update_multipliers (:276-278) and validate (:297-301).
 *   viol_out   [K] float   outputs["constraint_violations"]  (batch SUMS, trainer.py:195,204,239)
 *   loss_out   [1] float   outputs["lagrange_loss"] = sum_i w_i * viol_i (trainer.py:247)
 *   viol_accum [K] double  optional running sum over calls (the T.sum(axis=0) of trainer.py:274-280)
 *   q_ee_out   (B,3) float optional outputs["q_ee_actual"] = (x, y, phi) of the end effector
 *   err_sq_out (B,2) float optional outputs["position_err_sq"]  (trainer.py:192)
 */
IklStatus ikl_lagrangian_forward(const IklSpec* spec, const IklBatch* in, const IklWeights* w,
                                 void* workspace, float* viol_out, float* loss_out, double* viol_accum,
                                 float* q_ee_out, float* err_sq_out, void* stream);

/* One pass that produces the sums AND d(sum_i w_i g_i)/d(theta) analytically (no autograd tape).
 * Replaces process_batch (trainer.py:182-247) followed by lagrange_loss.backward() down to the
 * network output (trainer.py:264).  dtheta element (b, j) is written at dtheta[b*dtheta_sb + j*dtheta_sj].
 * Gradient conventions at the non-smooth points follow torch autograd (max ties split evenly,
 * sign(0) = 0); at a circle centre (d == 0) the direction term is 0 where torch gives NaN. */
IklStatus ikl_lagrangian_forward_backward(const IklSpec* spec, const IklBatch* in, const IklWeights* w,
                                          void* workspace, float* viol_out, float* loss_out, double* viol_accum,
                                          float* dtheta, int64_t dtheta_sb, int64_t dtheta_sj, void* stream);

/* ---- multi-GPU: the K sums exchanged INSIDE the kernel over NVLink peer memory ---------------------------------------
 * The reference has no distributed code; the path shards because every constraint value is a SUM over samples
 * (scripts/trainer.py:195,204,239).  With one process per GPU, each rank runs the fused kernel on its shard; the block that
 * finishes a rank's reduction stores that rank's K double-precision sums straight into every peer's mailbox (NVLink P2P
 * stores; every 8-byte word carries 32 bits of payload and the launch's 32-bit sequence number, so value and "valid" flag
 * arrive in one atomic store: no fence, no separate flag), polls its own mailbox until the K-vectors of all peers have landed
 * and adds them in rank order -- so viol_out / loss_out / viol_accum hold the GLOBAL sums, bit-identical on every rank,
 * when the launch completes.  One kernel does the compute and the collective: no NCCL launch, no second kernel.
 *   peers[r]  rank r's exchange buffer (ikl_exchange_bytes() bytes, zero-filled once by its owner before anyone's first
 *             launch) as mapped into THIS process -- CUDA IPC / symmetric memory; peers[rank] is the local buffer.
 * Every rank must issue the same sequence of *_allreduce launches (they pair up by a sequence counter kept in the local
 * buffer, so CUDA-graph replays work); a peer that never arrives makes the kernel trap after a few seconds instead of
 * hanging the GPU. */
typedef struct IklExchange {
  int32_t rank, world;            /* 0 <= rank < world <= IKL_MAX_RANKS */
  void* peers[IKL_MAX_RANKS];
} IklExchange;

size_t ikl_exchange_bytes(void);

/* ikl_lagrangian_forward / ikl_lagrangian_forward_backward on this rank's shard, sums exchanged as described above.
 * dtheta is local to the shard (the gradient of the GLOBAL Lagrangian w.r.t. this rank's samples). */
IklStatus ikl_lagrangian_forward_allreduce(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace,
                                           float* viol_out, float* loss_out, double* viol_accum, const IklExchange* ex,
                                           void* stream);
IklStatus ikl_lagrangian_forward_backward_allreduce(const IklSpec* spec, const IklBatch* in, const IklWeights* w,
                                                    void* workspace, float* viol_out, float* loss_out, double* viol_accum,
                                                    float* dtheta, int64_t dtheta_sb, int64_t dtheta_sj,
                                                    const IklExchange* ex, void* stream);

/* ---- the launch that CLOSES a dual step: compute + collective + multiplier update in one kernel -----------------------
 * update_multipliers (scripts/trainer.py:267-286) sums the K-vectors of every validation batch and then moves lambda ONCE.
 * So a data-parallel dual step needs the global sums once per epoch, not once per batch: the earlier batches of the epoch
 * are plain ikl_lagrangian_forward[_backward] launches that add into the rank's LOCAL viol_accum, and the epoch's last batch
 * is launched through one of the two entry points below.  The block that finishes that launch's reduction
 *   1. adds the launch's K sums to viol_accum (this rank's running sum of the epoch),
 *   2. with ex (world > 1) exchanges those K running sums over NVLink peer memory -- the tagged-word protocol described
 *      above, ONE exchange per epoch -- and adds the ranks' vectors in rank order,
 *   3. leaves the GLOBAL epoch sums in viol_accum (bit-identical on every rank) and,
 *   4. with dual, writes lam_out[i] = max(0, lam_in[i] + multiplier_lr * (float)viol_accum[i]) -- the arithmetic of
 *      ikl_dual_update, so lambda is bit-identical on every rank and equal to what the separate kernel would give.
 * viol_out / loss_out hold this launch's LOCAL sums and Lagrangian, evaluated with the multipliers of BEFORE the update
 * (w; w->device and dual->lam_out may be the same buffer).  viol_accum is required; ex may be NULL (single process: the
 * dual update is simply fused into the last forward launch); dual may be NULL (only close the epoch's sum).
 * Every rank must close the epoch with the same kind of launch (its batch may be empty). */
typedef struct IklDualStep {
  const float* lam_in;            /* [K] device */
  float* lam_out;                 /* [K] device; may alias lam_in */
  float multiplier_lr;            /* opt.multiplier_lr (scripts/options.py) */
} IklDualStep;

IklStatus ikl_lagrangian_forward_dual_step(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace,
                                           float* viol_out, float* loss_out, double* viol_accum, const IklExchange* ex,
                                           const IklDualStep* dual, void* stream);
IklStatus ikl_lagrangian_forward_backward_dual_step(const IklSpec* spec, const IklBatch* in, const IklWeights* w,
                                                    void* workspace, float* viol_out, float* loss_out, double* viol_accum,
                                                    float* dtheta, int64_t dtheta_sb, int64_t dtheta_sj,
                                                    const IklExchange* ex, const IklDualStep* dual, void* stream);

/* In-place sum over the ranks of a flat fp32 vector -- the one exchange data-parallel training adds to the reference's
 * step: the replicas' parameter gradients after lagrange_loss.backward() (scripts/trainer.py:263-265; the reference is
 * single-process).  ONE kernel over NVLink peer memory, same tagged-word protocol as the K-sum exchange above: every rank
 * pushes its elements (fp32 bits + the launch's 32-bit sequence number in one 8-byte store) into every peer's mailbox,
 * polls its own mailbox and adds the ranks' values in rank order, so the result is bit-identical everywhere.  No NCCL
 * launch, capturable in a CUDA graph (the sequence number lives in the mailbox and is advanced by the kernel).  Meant for
 * latency-bound sizes (the reference's networks have 2.5 k - 83 k parameters); every rank must call it with the same n in
 * the same order.
 *   ex->peers[r]  rank r's mailbox: ikl_allreduce_bytes(max_elems, world) bytes of symmetric memory, zero-filled once by
 *                 its owner before anyone's first launch.  n <= max_elems. */
size_t ikl_allreduce_bytes(int64_t max_elems, int32_t world);
IklStatus ikl_allreduce_sum_f32(float* data, int64_t n, const IklExchange* ex, int64_t max_elems, void* stream);

/* Gradient alone: d(sum_i w_i g_i)/d(theta) for arbitrary per-constraint weights w (for autograd:
 * w_i = grad(lagrange_loss) * lambda_i + grad(constraint_violations)_i).  Recomputes from the inputs -- nothing is
 * saved by the forward pass -- and replaces the autograd tape between lagrange_loss and the network output
 * (scripts/trainer.py:264).  This is also the launch of a TRAINING step: train_with_relaxation consumes nothing of
 * process_batch's outputs but lagrange_loss.backward() (scripts/trainer.py:262-265), so no constraint sum is needed there.
 * On the staged layouts (rows >= 512, TMA planes) it runs the GRADIENT-ONLY specialisations of the fused kernels: same
 * per-sample arithmetic as ikl_lagrangian_forward_backward (dtheta is bit-identical), but no running sums, no reduction,
 * no partials, nothing written but dtheta.  Other layouts run the full kernel and discard its sums (IKL_BACKWARD_SUMS=1
 * forces that everywhere). */
IklStatus ikl_lagrangian_backward(const IklSpec* spec, const IklBatch* in, const IklWeights* w, void* workspace,
                                  float* dtheta, int64_t dtheta_sb, int64_t dtheta_sj, void* stream);

/* lambda <- max(0, lambda + multiplier_lr * viol_sum)   (scripts/trainer.py:280-284, update_multipliers).
 * viol_sum is the double running sum filled through viol_accum (or NULL to use viol_sum_f32).
 * lam_out may alias lam_in. */
IklStatus ikl_dual_update(const float* lam_in, const double* viol_sum, const float* viol_sum_f32,
                          float multiplier_lr, int32_t num_constraints, float* lam_out, void* stream);

/* ---- function-level drop-ins (let an unchanged scripts/trainer.py reach CUDA op by op) ----------
 * float32 only.  The reference's Python functions are dtype-generic (torch evaluates a float64 tensor in float64); the
 * Python wrappers (ikl/ops.py) refuse any other dtype with a TypeError rather than narrowing it silently. */

/* kinematics/forward_kinematics.py:47-83 ForwardKinematicsThreeLinkTorch (num_links 3; also 2: the maths of :7-22, and
 * 8: what the reference's non-running eight-link function :86-110 spells out).
 * tips_out: J contiguous (B,3) blocks, END EFFECTOR FIRST: block k holds tip J-k as (x, y, cumulative angle). */
IklStatus ikl_fk_forward(const float* theta, int64_t theta_sb, int64_t theta_sj, int64_t batch,
                         int32_t num_links, const float* link_length_host, float* tips_out, void* stream);

/* Backward of the above: grad_tips has the layout of tips_out (NULL blocks are not supported; pass zeros);
 * writes dtheta (B,J) contiguous. */
IklStatus ikl_fk_backward(const float* theta, int64_t theta_sb, int64_t theta_sj, int64_t batch,
                          int32_t num_links, const float* link_length_host, const float* grad_tips,
                          float* dtheta, void* stream);

/* kinematics/penalties.py:5-17 piecewise_circle_penalty, elementwise over n broadcast elements.
 * Each operand is a 1-D strided view (stride 0 = broadcast).  Bit-exact with the reference's fp32
 * arithmetic (same roundings: no FMA contraction, IEEE sqrt). */
IklStatus ikl_circle_penalty_forward(const float* jx, int64_t s_jx, const float* jy, int64_t s_jy,
                                     const float* ox, int64_t s_ox, const float* oy, int64_t s_oy,
                                     const float* radius, int64_t s_r, int64_t n,
                                     float inside_slope, float outside_slope, float* out, void* stream);

/* Backward: grad_out (n elements, element stride s_g; s_g = 0: ONE value for every element, which is what the
 * `.sum()` the reference applies to every penalty hands back -- scripts/trainer.py:204, :239 -- so the expanded gradient
 * is never materialised) -> d/djx, d/djy (n, contiguous; either may be NULL), and optionally d/dox, d/doy, d/dradius
 * per element (caller reduces over broadcast dims). */
IklStatus ikl_circle_penalty_backward(const float* jx, int64_t s_jx, const float* jy, int64_t s_jy,
                                      const float* ox, int64_t s_ox, const float* oy, int64_t s_oy,
                                      const float* radius, int64_t s_r, int64_t n,
                                      float inside_slope, float outside_slope, const float* grad_out, int64_t s_g,
                                      float* d_jx, float* d_jy, float* d_ox, float* d_oy, float* d_radius,
                                      void* stream);

/* kinematics/penalties.py:44-56 piecewise_joint_limit_penalty, elementwise, same conventions. */
IklStatus ikl_joint_limit_penalty_forward(const float* theta, int64_t s_t, const float* lo, int64_t s_lo,
                                          const float* hi, int64_t s_hi, int64_t n,
                                          float inside_slope, float outside_slope, float* out, void* stream);

/* grad_out: n elements with element stride s_g (0 = one value for every element: the gradient of `.sum()`,
 * scripts/trainer.py:204). */
IklStatus ikl_joint_limit_penalty_backward(const float* theta, int64_t s_t, const float* lo, int64_t s_lo,
                                           const float* hi, int64_t s_hi, int64_t n,
                                           float inside_slope, float outside_slope, const float* grad_out, int64_t s_g,
                                           float* d_theta, float* d_lo, float* d_hi, void* stream);

/* ---- the callers either side of the path, on the device ------------------------------------------ */

/* kinematics/dataset.py:52-121 IkDataset.initialize, for 3-link arms: uniform joint angles inside the limits, FK
 * for the targets, rejection sampling of dynamic obstacle centres in [-1.2, 1.2]^2 (dataset.py:98-104) or of the
 * joint angles against static obstacles (dataset.py:109-121), collision rule of kinematics/penalties.py:75-92.
 * One Philox stream per example: the output depends only on (seed, example index), not on the launch geometry.
 * It has the reference's distribution, not its torch-CPU random stream (kinematics.dataset.IkDataset reproduces
 * that one exactly, on the host).
 *   theta_out (n,3), ee_out (n,3) = (x, y, phi); obstacles_out (n,4,4) rows [x, y, radius, 0] for dynamic configs
 *   (NULL otherwise); exhausted_out: device counter (zeroed by the caller) of examples that gave up after 100000
 *   re-draws (an obstacle covering the arm base loops forever in the reference). */
IklStatus ikl_dataset_generate(const IklSpec* spec, int64_t n, uint64_t seed, float dynamic_radius, float* theta_out,
                               float* ee_out, float* obstacles_out, uint32_t* exhausted_out, void* stream);

/* scripts/evaluation/visualize_ik_solutions.py:97-119 on the device: over the per-sample constraint terms of a batch,
 * counts_out[i] (i < K) += #(term_i > threshold) (the reference uses 1e-4); counts_out[16] += #samples with any
 * joint-limit term > 0; counts_out[17] += #samples with any obstacle term > 0; pos_err_sum_out += sum over samples of
 * sqrt(sum(position_err_sq)).  counts_out: 18 uint64 on the device, accumulated across calls (zero them first). */
IklStatus ikl_eval_stats(const IklSpec* spec, const IklBatch* in, float threshold, uint64_t* counts_out,
                         double* pos_err_sum_out, void* stream);

/* ---- housekeeping ------------------------------------------------------------------------------ */
int32_t ikl_abi_version(void);
const char* ikl_strerror(IklStatus status);
int32_t ikl_last_cuda_error(void);
/* Leave `n` SMs out of the persistent grids of the Lagrangian kernels (default 0).  A multi-GPU caller sets 1 so that
 * the tiny NCCL all-reduce of step k runs on the free SM while the kernel of step k+1 already occupies the others. */
void ikl_set_reserved_sms(int32_t n);
/* Number of kernel launches this library has issued in this process (bench.py's gpu_launches). */
int64_t ikl_launch_count(void);
/* Name of the kernel variant the last ikl_lagrangian_* call on this thread dispatched to. */
const char* ikl_last_kernel_name(void);

#ifdef __cplusplus
}
#endif
#endif  /* IKL_H_ */
