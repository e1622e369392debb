/*
 * ik_oracle.c -- plain-C restatement of the FK + penalty Lagrangian and its closed-form gradient.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/ik_oracle.py for the rules): a second, independent CPU checker next to the
 * PyTorch restatement, fast enough (OpenMP over samples) to evaluate full benchmark-size batches in seconds, and the
 * "what a tuned CPU port would do" figure bench.py reports beside the reference-style PyTorch CPU baseline.
 *
 * Reference lines followed (miloknowles/lagrange-dual-learning):
 *   forward kinematics      kinematics/forward_kinematics.py:64-81
 *   joint-limit penalty     kinematics/penalties.py:51-56   (inside = good slope, outside = bad: scripts/trainer.py:203-204)
 *   circle penalty          kinematics/penalties.py:12-17   (inside = bad slope, outside = good: scripts/trainer.py:237)
 *   sums / Lagrangian       scripts/trainer.py:192-247      (constraint order: EE, JL1..3, then obstacle-major, end effector first)
 *   gradient                SURVEY appendix A.2 with torch's conventions (max ties split evenly, sign(0) = 0); at a
 *                           circle centre the direction term is 0 (torch: NaN) -- the documented deviation.
 *
 * PARITY PINNED: tests/test_oracle_pinned.py compares this file with oracle/ik_oracle.py, which is pinned bit for bit to
 * golden vectors produced by running the unmodified reference (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libik_oracle.so oracle/ik_oracle.c -lm   (oracle/build_oracle.py)
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXK 16

typedef struct OracleSpecC {
  int enforce_joint_limits;
  int num_obstacles;          /* 0..4 */
  int dynamic;                /* 1: obstacles per sample (B,4,4); 0: static table */
  float limit_lo, limit_hi;   /* float32 like torch.Tensor(json limits) */
  double good_slope, bad_slope;
  float static_obstacles[4][3];
} OracleSpecC;

static int num_constraints(const OracleSpecC* s) { return 1 + (s->enforce_joint_limits ? 3 : 0) + 3 * s->num_obstacles; }

int ik_oracle_num_constraints(const OracleSpecC* s) { return num_constraints(s); }
int ik_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* One sample in double precision.  terms[K] receives the K constraint terms, dth[3] the gradient (may be NULL). */
static void sample_f64(const OracleSpecC* s, const float* th_f, const float* tg_f, const float* ob_f, const double* w,
                       double* terms, double* dth) {
  const double a = s->good_slope, b = s->bad_slope;
  double th[3] = {th_f[0], th_f[1], th_f[2]};
  double phi[3], c[3], sn[3], px[3], py[3], gx[3] = {0, 0, 0}, gy[3] = {0, 0, 0}, dj[3] = {0, 0, 0};
  phi[0] = th[0]; phi[1] = phi[0] + th[1]; phi[2] = phi[1] + th[2];
  for (int m = 0; m < 3; ++m) { c[m] = cos(phi[m]); sn[m] = sin(phi[m]); }
  px[0] = c[0]; py[0] = sn[0];
  for (int m = 1; m < 3; ++m) { px[m] = px[m - 1] + c[m]; py[m] = py[m - 1] + sn[m]; }
  const double ex = (double)tg_f[0] - px[2], ey = (double)tg_f[1] - py[2];
  terms[0] = 0.5 * ex * ex + 0.5 * ey * ey;
  gx[2] = -w[0] * ex; gy[2] = -w[0] * ey;
  int idx = 1;
  if (s->enforce_joint_limits) {
    const double lo = s->limit_lo, hi = s->limit_hi;
    const double mid = (lo + hi) / 2.0, h = 0.5 * fabs(hi - lo);
    for (int j = 0; j < 3; ++j, ++idx) {
      const double u = th[j] - mid, au = fabs(u);
      const double vin = a * au - a * h, vout = b * au - b * h;
      terms[idx] = vin > vout ? vin : vout;
      const double slope = vin > vout ? a : (vin < vout ? b : 0.5 * (a + b));
      dj[j] = w[idx] * slope * (u > 0 ? 1.0 : (u < 0 ? -1.0 : 0.0));
    }
  }
  for (int o = 0; o < s->num_obstacles; ++o) {
    const double ox = s->dynamic ? (double)ob_f[4 * o] : (double)s->static_obstacles[o][0];
    const double oy = s->dynamic ? (double)ob_f[4 * o + 1] : (double)s->static_obstacles[o][1];
    const double r = s->dynamic ? (double)ob_f[4 * o + 2] : (double)s->static_obstacles[o][2];
    for (int j = 0; j < 3; ++j, ++idx) {
      const int m = 2 - j;                  /* FK tuple index j -> tip m (end effector first) */
      const double dx = px[m] - ox, dy = py[m] - oy;
      const double d = sqrt(dx * dx + dy * dy), t = d - r;
      const double vin = -b * t, vout = -a * t;
      terms[idx] = vin > vout ? vin : vout;
      const double slope = vin > vout ? -b : (vin < vout ? -a : -0.5 * (a + b));
      const double inv = d > 0 ? 1.0 / d : 0.0;
      gx[m] += w[idx] * slope * dx * inv;
      gy[m] += w[idx] * slope * dy * inv;
    }
  }
  if (dth) {
    double sx = 0, sy = 0, run = 0;
    for (int m = 2; m >= 0; --m) {
      sx += gx[m]; sy += gy[m];
      run += -sn[m] * sx + c[m] * sy;
      dth[m] = run + dj[m];
    }
  }
}

/* Same arithmetic in float (libm cosf/sinf/sqrtf), the way a straightforward CPU port of the fp32 path computes it. */
static void sample_f32(const OracleSpecC* s, const float* th, const float* tg, const float* ob, const float* w,
                       float* terms, float* dth) {
  const float a = (float)s->good_slope, b = (float)s->bad_slope;
  float phi[3], c[3], sn[3], px[3], py[3], gx[3] = {0, 0, 0}, gy[3] = {0, 0, 0}, dj[3] = {0, 0, 0};
  phi[0] = th[0]; phi[1] = phi[0] + th[1]; phi[2] = phi[1] + th[2];
  for (int m = 0; m < 3; ++m) { c[m] = cosf(phi[m]); sn[m] = sinf(phi[m]); }
  px[0] = c[0]; py[0] = sn[0];
  for (int m = 1; m < 3; ++m) { px[m] = px[m - 1] + c[m]; py[m] = py[m - 1] + sn[m]; }
  const float ex = tg[0] - px[2], ey = tg[1] - py[2];
  terms[0] = 0.5f * (ex * ex) + 0.5f * (ey * ey);
  gx[2] = -w[0] * ex; gy[2] = -w[0] * ey;
  int idx = 1;
  if (s->enforce_joint_limits) {
    const float mid = (s->limit_lo + s->limit_hi) / 2.0f, rng = fabsf(s->limit_hi - s->limit_lo);
    const float ca = (a * 0.5f) * rng, cb = (b * 0.5f) * rng;
    for (int j = 0; j < 3; ++j, ++idx) {
      const float u = th[j] - mid, au = fabsf(u);
      const float vin = a * au - ca, vout = b * au - cb;
      terms[idx] = vin > vout ? vin : vout;
      const float slope = vin > vout ? a : (vin < vout ? b : 0.5f * (a + b));
      dj[j] = w[idx] * slope * (u > 0 ? 1.0f : (u < 0 ? -1.0f : 0.0f));
    }
  }
  for (int o = 0; o < s->num_obstacles; ++o) {
    const float ox = s->dynamic ? ob[4 * o] : s->static_obstacles[o][0];
    const float oy = s->dynamic ? ob[4 * o + 1] : s->static_obstacles[o][1];
    const float r = s->dynamic ? ob[4 * o + 2] : s->static_obstacles[o][2];
    for (int j = 0; j < 3; ++j, ++idx) {
      const int m = 2 - j;
      const float dx = px[m] - ox, dy = py[m] - oy;
      const float d = sqrtf(dx * dx + dy * dy), t = d - r;
      const float vin = -b * t, vout = -a * t;
      terms[idx] = vin > vout ? vin : vout;
      const float slope = vin > vout ? -b : (vin < vout ? -a : -0.5f * (a + b));
      const float inv = d > 0 ? 1.0f / d : 0.0f;
      gx[m] += w[idx] * slope * dx * inv;
      gy[m] += w[idx] * slope * dy * inv;
    }
  }
  if (dth) {
    float sx = 0, sy = 0, run = 0;
    for (int m = 2; m >= 0; --m) {
      sx += gx[m]; sy += gy[m];
      run += -sn[m] * sx + c[m] * sy;
      dth[m] = run + dj[m];
    }
  }
}

/* Row-major inputs as the reference's DataLoader hands them: theta (B,3), target (B,3), obstacles (B,4,4) or NULL.
 * sums[K] = sum_b terms; abs_sums[K] = sum_b |terms| (NULL to skip); dtheta (B,3) doubles (NULL to skip). */
int ik_oracle_lagrangian_f64(const OracleSpecC* s, long long B, const float* theta, const float* target, const float* obst,
                             const double* w, double* sums, double* abs_sums, double* dtheta) {
  const int K = num_constraints(s);
  if (K > MAXK || (s->dynamic && s->num_obstacles > 0 && !obst)) return 1;
  double tot[MAXK] = {0}, tot_abs[MAXK] = {0};
#pragma omp parallel
  {
    double loc[MAXK] = {0}, loc_abs[MAXK] = {0}, terms[MAXK], d[3];
#pragma omp for schedule(static)
    for (long long i = 0; i < B; ++i) {
      sample_f64(s, theta + 3 * i, target + 3 * i, obst ? obst + 16 * i : NULL, w, terms, dtheta ? d : NULL);
      for (int k = 0; k < K; ++k) { loc[k] += terms[k]; loc_abs[k] += fabs(terms[k]); }
      if (dtheta) { dtheta[3 * i] = d[0]; dtheta[3 * i + 1] = d[1]; dtheta[3 * i + 2] = d[2]; }
    }
#pragma omp critical
    for (int k = 0; k < K; ++k) { tot[k] += loc[k]; tot_abs[k] += loc_abs[k]; }
  }
  for (int k = 0; k < K; ++k) { sums[k] = tot[k]; if (abs_sums) abs_sums[k] = tot_abs[k]; }
  return 0;
}

/* float arithmetic per sample, double accumulation of the sums; dtheta (B,3) floats. */
int ik_oracle_lagrangian_f32(const OracleSpecC* s, long long B, const float* theta, const float* target, const float* obst,
                             const float* w, double* sums, float* dtheta) {
  const int K = num_constraints(s);
  if (K > MAXK || (s->dynamic && s->num_obstacles > 0 && !obst)) return 1;
  double tot[MAXK] = {0};
#pragma omp parallel
  {
    double loc[MAXK] = {0};
    float terms[MAXK], d[3];
#pragma omp for schedule(static)
    for (long long i = 0; i < B; ++i) {
      sample_f32(s, theta + 3 * i, target + 3 * i, obst ? obst + 16 * i : NULL, w, terms, dtheta ? d : NULL);
      for (int k = 0; k < K; ++k) loc[k] += terms[k];
      if (dtheta) { dtheta[3 * i] = d[0]; dtheta[3 * i + 1] = d[1]; dtheta[3 * i + 2] = d[2]; }
    }
#pragma omp critical
    for (int k = 0; k < K; ++k) tot[k] += loc[k];
  }
  for (int k = 0; k < K; ++k) sums[k] = tot[k];
  return 0;
}

/* Per-sample diagnostics in double precision for tools/parity_report.py (any output may be NULL):
 *   kink_dist[b]  min over the penalty terms (everything but the end-effector term) of |term|: how close the sample sits to
 *                 a kink of a piecewise-linear penalty, where a 1-ulp difference legitimately flips the subgradient
 *                 (the GPU tests mask |term| < 1e-6); +infinity when the config has no penalty terms
 *   cond[b]       sum over the obstacle terms of |w_i| * max(slope) / d_i: the lever arm by which the unit vector (p - c)/d
 *                 amplifies the ~1e-7 round-off any fp32 forward kinematics carries into the gradient
 *   tips[b][6]    (p3x, p3y, p2x, p2y, p1x, p1y)  -- end effector first, like the FK tuple
 *   terms[b][K]   the K per-sample constraint terms */
int ik_oracle_diagnostics_f64(const OracleSpecC* s, long long B, const float* theta, const float* target, const float* obst,
                              const double* w, double* kink_dist, double* cond, double* tips, double* terms_out) {
  const int K = num_constraints(s);
  if (K > MAXK || (s->dynamic && s->num_obstacles > 0 && !obst)) return 1;
  const double hi = s->good_slope > s->bad_slope ? s->good_slope : s->bad_slope;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < B; ++i) {
    double terms[MAXK];
    sample_f64(s, theta + 3 * i, target + 3 * i, obst ? obst + 16 * i : NULL, w, terms, NULL);
    if (terms_out) for (int k = 0; k < K; ++k) terms_out[(size_t)i * K + k] = terms[k];
    if (kink_dist) {
      double m = INFINITY;
      for (int k = 1; k < K; ++k) m = fabs(terms[k]) < m ? fabs(terms[k]) : m;
      kink_dist[i] = m;
    }
    if (cond || tips) {
      const double t0 = theta[3 * i], t1 = t0 + theta[3 * i + 1], t2 = t1 + theta[3 * i + 2];
      double px[3], py[3];
      px[0] = cos(t0); py[0] = sin(t0);
      px[1] = px[0] + cos(t1); py[1] = py[0] + sin(t1);
      px[2] = px[1] + cos(t2); py[2] = py[1] + sin(t2);
      if (tips) for (int j = 0; j < 3; ++j) { tips[6 * i + 2 * j] = px[2 - j]; tips[6 * i + 2 * j + 1] = py[2 - j]; }
      if (cond) {
        double c = 0.0;
        int idx = 1 + (s->enforce_joint_limits ? 3 : 0);
        for (int o = 0; o < s->num_obstacles; ++o) {
          const float* ob = s->dynamic ? obst + 16 * i + 4 * o : s->static_obstacles[o];
          for (int j = 0; j < 3; ++j, ++idx) {
            const double dx = px[2 - j] - (double)ob[0], dy = py[2 - j] - (double)ob[1];
            const double d = sqrt(dx * dx + dy * dy);
            c += fabs(w[idx]) * hi / (d > 1e-12 ? d : 1e-12);
          }
        }
        cond[i] = c;
      }
    }
  }
  return 0;
}
