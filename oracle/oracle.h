/* oracle/oracle.h — TEST INFRASTRUCTURE. C ABI of the CPU oracle (liblfba_oracle.so).
 *
 * The oracle is a CPU restatement of the reference's LF-BA path:
 *   functors   /root/reference/src/BundleAdjustment/BundleAdjustment.h:25-279, src/CameraModel.h:87-264
 *   assembly   /root/reference/src/CameraCalibration.cpp:774-965
 *   solver     Ceres Solver 2.1.0 (external, un-vendored; pin: installation/Dockerfile:105,
 *              installation/playbooks/LiFCal_install_CeresSolver_2_1_0.yaml:32): trust-region
 *              Levenberg-Marquardt, Jacobi scaling, DENSE_SCHUR, Eigen LLT, CauchyLoss, SubsetManifold,
 *              box bounds + projected Armijo line search — restated from its published algorithm
 *              (SURVEY.md Appendix B).
 * Parity status: the FUNCTOR arithmetic is pinned against the reference's own headers compiled in place
 * (oracle/ref_bridge.cpp -> oracle/_ref/libref_functor.so, tests/test_oracle_ref_pin.py).  The LM LOOP is
 * "parity unpinned": Ceres is absent from /root/reference and from this image and the reference ships no
 * golden vectors, so the loop is a restatement checked only against Ceres' documented semantics.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * It shares the plain-C problem/option/summary structs of include/lfba.h (types only).
 */
#ifndef LFBA_ORACLE_H_
#define LFBA_ORACLE_H_
#include "../include/lfba.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Optional per-block evaluator override: lets a test plug the reference's own functor
 * (oracle/_ref/libref_functor.so: ref_block_eval) into the oracle's solver. Signature:
 *   config, obs(2), ml(2), spx, spy, scale, camera17, view6 (or NULL), point3 (or NULL),
 *   fixed_point3 (or NULL), fixed_view6 (or NULL) -> residual2, jac 2x26 row-major [cam17|view6|point3]
 *   (jac may be NULL).  Returns 1 on success. */
typedef int (*oracle_block_fn)(uint32_t config, const double* obs, const double* ml, double spx, double spy,
                               double scale, const double* camera, const double* view, const double* point,
                               const double* fixed_point, const double* fixed_view, double* residual,
                               double* jac);

/* Raw (loss-uncorrected) residuals and autodiff Jacobians of every reprojection block, in input order.
 * Same output layout as lfba_eval. `block_fn` may be NULL (use the restated functor). */
int oracle_eval(const lfba_problem* problem, const double* camera17, const double* views6F,
                const double* points3P, double* residuals, double* jac_camera, double* jac_view,
                double* jac_point, double* cost, lfba_reproj_stats* stats, double inlier_threshold,
                int num_threads, oracle_block_fn block_fn);

/* Full Ceres-equivalent solve; camera/views/points updated in place to the last accepted iterate.
 * max_seconds > 0 bounds the run for baseline timing (the loop stops after the iteration that
 * crosses it; summary->num_iterations tells how far it got). */
int oracle_solve(const lfba_problem* problem, const lfba_options* options, double* camera17, double* views6F,
                 double* points3P, lfba_summary* summary, int num_threads, double max_seconds,
                 oracle_block_fn block_fn);

/* The same solve without the O(N) Jacobian storage ("streaming", block-recompute mode): the block rows of one point's
 * observations are re-derived by autodiff wherever Ceres would read its stored Jacobian (evaluation, Schur
 * elimination, back-substitution), so the 1M-point x 1000-frame scene (41.6 GB of Jacobian in Ceres' layout) runs in
 * a few GB. Results equal oracle_solve to rounding (sums are taken in point order). On return
 * summary->num_lenses holds the number of passes over all blocks (3 per LM iteration instead of 1). */
int oracle_solve_streaming(const lfba_problem* problem, const lfba_options* options, double* camera17,
                           double* views6F, double* points3P, lfba_summary* summary, int num_threads,
                           double max_seconds, oracle_block_fn block_fn);

/* Wall-clock seconds of `reps` full residual+Jacobian evaluations (Jet<26> autodiff, Jacobian
 * materialised in Ceres' block layout, gradient formed) — the "M evals/s" CPU baseline. */
int oracle_time_eval(const lfba_problem* problem, const double* camera17, const double* views6F,
                     const double* points3P, int reps, int num_threads, double* seconds);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
