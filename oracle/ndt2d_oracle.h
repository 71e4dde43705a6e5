/*
 * SPEC ORACLE — test infrastructure, NOT product code, NOT the reference matcher.
 *
 * A plain-C CPU restatement of SPEC.md (this repo's frozen 2D NDT specification).
 * The upstream reference (sven-glory/GTSAM-NDT) is mounted with a single file,
 * /root/reference/README.md:1 ("# GTSAM-NDT"), and no source, tests or golden
 * vectors. PARITY IS THEREFORE UNPINNED: this oracle is checked only against
 * closed-form cases, an independent numpy f64 model and finite differences
 * (tests/test_oracle_*.py), never against the reference's own arithmetic.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library. The product (gtsam_ndt_b200) never does.
 */
#ifndef NDT2D_ORACLE_H
#define NDT2D_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_LEVELS 8

/* SPEC.md section 1. Same field order as ndt2d_params in include/ndt2d.h. */
typedef struct {
    double eig_ratio, eps_trans, eps_rot, max_step_trans, max_step_rot;
    double lambda_init, lambda_min, lambda_max, lambda_up, lambda_down, lambda_fail_up;
    int32_t min_points, max_iterations, overlap, reserved;
} oracle_params;

/* SPEC.md section 5 outputs. Same layout as ndt2d_result. */
typedef struct {
    double pose[3];
    double score;
    double grad[3];
    double hessian[9];
    int32_t iterations, status, count, reserved;
} oracle_result;

typedef struct oracle_matcher oracle_matcher;

oracle_matcher *oracle_create(void);
void oracle_destroy(oracle_matcher *m);
void oracle_default_params(oracle_params *p);
int oracle_set_params(oracle_matcher *m, const oracle_params *p);
/* pyramid of cell sizes, coarse to fine; one level = plain NDT */
int oracle_set_resolutions(oracle_matcher *m, const float *res, int nlevels);
/* explicit lattice: origin and extent in metres; extent <= 0 restores auto-fit */
int oracle_set_grid(oracle_matcher *m, float ox, float oy, float ex, float ey);
/* SPEC 3: build (replace) or extend (SPEC 7) the target */
int oracle_set_target(oracle_matcher *m, const float *xy, int n);
int oracle_add_target(oracle_matcher *m, const float *xy, int n);

/* geometry of a level: out = {res, st, inv_st, ox, oy}, dims = {nhx, nhy, njx, njy} */
int oracle_level_geometry(const oracle_matcher *m, int level, float out[5], int32_t dims[4]);
/* cell records (njx*njy*8 floats: mux muy | B00 B01 | B01 B11 | n valid) and raw integer sums (n: u32, sums: 5 x i64 per cell) */
int oracle_get_cells(const oracle_matcher *m, int level, float *cells);
int oracle_get_sums(const oracle_matcher *m, int level, uint32_t *n, int64_t *sums);
/* SPEC 2: lattice index of points transformed by pose (pose == NULL: identity, no transform) */
int oracle_cell_index(const oracle_matcher *m, int level, const float *xy, int n,
                      const double *pose, int32_t *idx);
/* SPEC 4: out10 = {S, g0,g1,g2, H00,H01,H02,H11,H12,H22} */
int oracle_evaluate(const oracle_matcher *m, int level, const float *xy, int n,
                    const double pose[3], double out10[10], int32_t *count);
/* SPEC 4 per-pair factors (e, c1..c9) for bit-exact checks: terms[n*K*10] f32 (zeros where skipped) */
int oracle_point_terms(const oracle_matcher *m, int level, const float *xy, int n,
                       const double pose[3], float *terms);
/* SPEC 5 */
int oracle_align(const oracle_matcher *m, const float *xy, int n, const double init[3],
                 oracle_result *res);
/* independent scans; offsets[b]..offsets[b+1] index points; nthreads <= 0: all cores */
int oracle_align_batch(const oracle_matcher *m, const float *xy, const int64_t *offsets, int nscans,
                       const double *init, oracle_result *res, int nthreads);
/* SPEC 6: scores[m] (may be NULL), best index / score */
int oracle_sweep(const oracle_matcher *m, int level, const float *xy, int n, const float *hyp,
                 int64_t nhyp, double *scores, int64_t *best_idx, double *best_score, int nthreads);
/* SPEC 8: polar to Cartesian; returns kept count. ranges_u16 or ranges_f32 (other NULL) */
int oracle_polar_to_points(const float *ranges_f32, const uint16_t *ranges_u16, int nbeams,
                           double angle_min, double angle_inc, float range_scale,
                           float range_min, float range_max, float *xy_out);
/* the bit-exact pieces, exposed for unit tests */
float oracle_expneg(float h);
void oracle_sincos(double th, double *sn, double *cs); /* SPEC 4.2 */
int oracle_solve(const double g[3], const double H6[6], double lambda, double d[3]);
int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
