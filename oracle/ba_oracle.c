/*
 * oracle/ba_oracle.c -- CPU restatement (FP64) of OpenVSLAM's motion-only and local bundle
 * adjustment, i.e. optimize::pose_optimizer::optimize and
 * optimize::local_bundle_adjuster::optimize together with the parts of g2o they drive:
 * OptimizationAlgorithmLevenberg::solve, BlockSolver_6_3 (landmarks marginalised by the
 * Schur complement), RobustKernelHuber, and the edge / vertex types of optimize/g2o/se3/.
 *
 * TEST INFRASTRUCTURE ONLY (see orb_oracle.c).  PARITY STATUS: **parity unpinned**: neither
 * OpenVSLAM's source nor g2o is available here (SURVEY.md sections 0 and 8c); this file restates
 * the published algorithms as recalled (file names per SURVEY.md 8a: optimize/pose_optimizer.cc,
 * optimize/local_bundle_adjuster.cc, optimize/g2o/se3/{perspective,equirectangular}_reproj_edge.cc, shot_vertex.h; g2o's
 * optimization_algorithm_levenberg.cpp, block_solver.hpp, robust_kernel_impl.cpp, se3quat.h).
 * Independent check available here: tests/test_ba_oracle.py verifies the analytic Jacobians
 * against finite differences and that the optimisers reduce the cost on problems with known
 * ground truth.
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include "ba_oracle.h"

/* ------------------------------------------------------------------ small linear algebra */
static void mat3_mul_vec(const double* R, const double* v, double* o) {
    o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
static void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

/* g2o::SE3Quat::exp(update), update = [omega(3), upsilon(3)]; returns R (row-major) and t. */
void ob_se3_exp(const double* u, double* R, double* t) {
    const double wx = u[0], wy = u[1], wz = u[2];
    const double theta = sqrt(wx * wx + wy * wy + wz * wz);
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9];
    mat3_mul(O, O, O2);
    double a, b, c; /* R = I + a O + b O2 ; V = I + b O + c O2 */
    if (theta < 0.00001) { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
    else {
        a = sin(theta) / theta;
        b = (1 - cos(theta)) / (theta * theta);
        c = (theta - sin(theta)) / (theta * theta * theta);
    }
    double V[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = I + a * O[i] + b * O2[i];
        V[i] = I + b * O[i] + c * O2[i];
    }
    mat3_mul_vec(V, u + 3, t);
}

/* shot_vertex::oplusImpl: estimate <- exp(update) * estimate.  pose = {R (9, row-major), t (3)}. */
void ob_pose_oplus(const double* pose, const double* u, double* out) {
    double Rd[9], td[3], Rn[9], tn[3];
    ob_se3_exp(u, Rd, td);
    mat3_mul(Rd, pose, Rn);
    mat3_mul_vec(Rd, pose + 9, tn);
    for (int i = 0; i < 9; ++i) out[i] = Rn[i];
    for (int i = 0; i < 3; ++i) out[9 + i] = tn[i] + td[i];
}

/* ------------------------------------------------------------------------ edge models */
/* Residual e = obs - project(R pw + t) and Jacobians of e wrt the pose update (Jp, dim x 6,
 * [omega, upsilon]) and wrt the landmark (Jl, dim x 3).  Returns dim (2 or 3).
 * perspective_reproj_edge / stereo_perspective_reproj_edge / equirectangular_reproj_edge. */
int ob_edge_eval(const ob_camera* cam, const double* pose, const double* pw, const double* obs, int stereo,
                 double* e, double* Jp, double* Jl, double* pc_out) {
    double pc[3];
    mat3_mul_vec(pose, pw, pc);
    pc[0] += pose[9]; pc[1] += pose[10]; pc[2] += pose[11];
    if (pc_out) { pc_out[0] = pc[0]; pc_out[1] = pc[1]; pc_out[2] = pc[2]; }
    const double x = pc[0], y = pc[1], z = pc[2];
    const double* R = pose;
    if (cam->model == OB_CAM_EQUIRECTANGULAR) {
        const double L = sqrt(x * x + y * y + z * z);
        const double theta = atan2(x, z);
        const double phi = -asin(y / L);
        e[0] = obs[0] - cam->cols * (0.5 + theta / (2 * M_PI));
        e[1] = obs[1] - cam->rows * (0.5 - phi / M_PI);
        if (Jp || Jl) {
            /* d pc / d [rx ry rz tx ty tz pwx pwy pwz] */
            double dpc[3][9] = {
                {0, z, -y, 1, 0, 0, R[0], R[1], R[2]},
                {-z, 0, x, 0, 1, 0, R[3], R[4], R[5]},
                {y, -x, 0, 0, 0, 1, R[6], R[7], R[8]}};
            double dL[9];
            for (int k = 0; k < 9; ++k) dL[k] = (1.0 / L) * (x * dpc[0][k] + y * dpc[1][k] + z * dpc[2][k]);
            const double xz2 = x * x + z * z;
            const double c0 = -(cam->cols / (2 * M_PI)) * (1.0 / xz2);
            const double c1 = -(cam->rows / M_PI) * (1.0 / (L * sqrt(xz2)));
            for (int k = 0; k < 9; ++k) {
                const double j0 = c0 * (z * dpc[0][k] - x * dpc[2][k]);
                const double j1 = c1 * (L * dpc[1][k] - y * dL[k]);
                if (k < 6) { if (Jp) { Jp[k] = j0; Jp[6 + k] = j1; } }
                else if (Jl) { Jl[k - 6] = j0; Jl[3 + k - 6] = j1; }
            }
        }
        return 2;
    }
    const double fx = cam->fx, fy = cam->fy, fb = cam->focal_x_baseline;
    const double z_sq = z * z;
    const double reproj_x = fx * x / z + cam->cx;
    e[0] = obs[0] - reproj_x;
    e[1] = obs[1] - (fy * y / z + cam->cy);
    if (stereo) e[2] = obs[2] - (reproj_x - fb / z);
    if (Jl) {
        for (int k = 0; k < 3; ++k) {
            Jl[k] = -fx * R[k] / z + fx * x * R[6 + k] / z_sq;
            Jl[3 + k] = -fy * R[3 + k] / z + fy * y * R[6 + k] / z_sq;
            if (stereo) Jl[6 + k] = Jl[k] - fb * R[6 + k] / z_sq;
        }
    }
    if (Jp) {
        Jp[0] = x * y / z_sq * fx; Jp[1] = -(1.0 + (x * x / z_sq)) * fx; Jp[2] = y / z * fx;
        Jp[3] = -1.0 / z * fx; Jp[4] = 0.0; Jp[5] = x / z_sq * fx;
        Jp[6] = (1.0 + y * y / z_sq) * fy; Jp[7] = -x * y / z_sq * fy; Jp[8] = -x / z * fy;
        Jp[9] = 0.0; Jp[10] = -1.0 / z * fy; Jp[11] = y / z_sq * fy;
        if (stereo) {
            Jp[12] = Jp[0] - fb * y / z_sq; Jp[13] = Jp[1] + fb * x / z_sq; Jp[14] = Jp[2];
            Jp[15] = Jp[3]; Jp[16] = 0.0; Jp[17] = Jp[5] - fb / z_sq;
        }
    }
    return stereo ? 3 : 2;
}

/* g2o::RobustKernelHuber::robustify */
static void huber(double e2, double delta, double* rho) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
    else {
        const double sqrte = sqrt(e2);
        rho[0] = 2 * sqrte * delta - dsqr;
        rho[1] = delta / sqrte;
        rho[2] = -0.5 * rho[1] / e2;
    }
}

/* Dense Cholesky solve A x = b (A symmetric n x n, row-major, lower triangle used; destroyed).
 * Returns 0 on success, -1 if A is not positive definite. */
static int chol_solve(double* A, int n, const double* b, double* x) {
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0.0) || !isfinite(d)) return -1;
        d = sqrt(d);
        A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * x[k];
        x[i] = s / A[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * x[k];
        x[i] = s / A[(size_t)i * n + i];
    }
    return 0;
}

static int inv3_sym(const double* D, double* Di) { /* D: full 3x3 */
    const double a = D[0], b = D[1], c = D[2], d = D[4], e = D[5], f = D[8];
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (det == 0.0 || !isfinite(det)) return -1;
    const double id = 1.0 / det;
    Di[0] = c00 * id; Di[1] = c01 * id; Di[2] = c02 * id;
    Di[3] = Di[1]; Di[4] = (a * f - c * c) * id; Di[5] = (b * c - a * e) * id;
    Di[6] = Di[2]; Di[7] = Di[5]; Di[8] = (a * d - b * b) * id;
    return 0;
}

/* -------------------------------------------------------------------- the optimiser core */
typedef struct {
    const ob_camera* cam;
    int K, L, M;
    double* poses;            /* K x 12 (current estimate) */
    const uint8_t* fixed;     /* K */
    double* points;           /* L x 3 (NULL in pose-only mode: points are constants in pts_const) */
    const double* pts_const;  /* pose-only: M x 3 landmark positions */
    const int* obs_kf; const int* obs_lm;
    const float* obs_xy; const float* obs_xr; const float* inv_sigma_sq;
    uint8_t* level;           /* M: 0 = active, 1 = outlier (excluded) */
    int use_huber; double delta;
    /* edge state */
    double* err;              /* M x 3: edge->_error as of the last computeActiveErrors */
    int* free_idx;            /* K: index among non-fixed poses or -1 */
    int nfree;
    const volatile int* force_stop;
} ob_problem;

static int edge_is_stereo(const ob_problem* P, int i) { return P->obs_xr && P->obs_xr[i] >= 0.0f; }
static const double* edge_point(const ob_problem* P, const double* points, int i) {
    return P->points ? points + 3 * (size_t)P->obs_lm[i] : P->pts_const + 3 * (size_t)i;
}
static void edge_obs(const ob_problem* P, int i, double* o) {
    o[0] = (double)P->obs_xy[2 * i]; o[1] = (double)P->obs_xy[2 * i + 1];
    o[2] = P->obs_xr ? (double)P->obs_xr[i] : -1.0;
}
static double edge_chi2(const ob_problem* P, int i) {
    const double* e = P->err + 3 * (size_t)i;
    const double w = (double)P->inv_sigma_sq[i];
    double c = w * (e[0] * e[0] + e[1] * e[1]);
    if (edge_is_stereo(P, i)) c += w * e[2] * e[2];
    return c;
}

/* SparseOptimizer::computeActiveErrors + activeRobustChi2 at the given state. */
static double compute_active_errors(ob_problem* P, const double* poses, const double* points) {
    double total = 0;
    for (int i = 0; i < P->M; ++i) {
        if (P->level[i]) continue;
        double o[3]; edge_obs(P, i, o);
        double* e = P->err + 3 * (size_t)i;
        e[2] = 0;
        ob_edge_eval(P->cam, poses + 12 * (size_t)P->obs_kf[i], edge_point(P, points, i), o, edge_is_stereo(P, i), e, NULL, NULL, NULL);
        const double c = edge_chi2(P, i);
        if (P->use_huber) { double rho[3]; huber(c, P->delta, rho); total += rho[0]; }
        else total += c;
    }
    return total;
}

typedef struct {
    int n;            /* 6 * nfree */
    double* Hpp;      /* n x n (block diagonal part filled by build, Schur result in S) */
    double* bp;       /* n */
    double* Hll;      /* L x 9 */
    double* bl;       /* L x 3 */
    double* Hpl;      /* M x 18 (6x3 per active edge on a free pose) */
    double* S; double* bS; double* x; /* n */
    double* xl;       /* L x 3 */
    double* Dinv;     /* L x 9 */
} ob_system;

/* BlockSolver::buildSystem: H and b from the current linearisation (robustified). */
static void build_system(ob_problem* P, ob_system* Y) {
    const int n = Y->n;
    memset(Y->Hpp, 0, sizeof(double) * (size_t)n * n);
    memset(Y->bp, 0, sizeof(double) * n);
    if (P->points) { memset(Y->Hll, 0, sizeof(double) * 9 * (size_t)P->L); memset(Y->bl, 0, sizeof(double) * 3 * (size_t)P->L); }
    for (int i = 0; i < P->M; ++i) {
        if (P->level[i]) continue;
        const int kf = P->obs_kf[i];
        const int fi = P->free_idx[kf];
        const int stereo = edge_is_stereo(P, i);
        double o[3], e[3] = {0, 0, 0}, Jp[18], Jl[9];
        edge_obs(P, i, o);
        const int dim = ob_edge_eval(P->cam, P->poses + 12 * (size_t)kf, edge_point(P, P->points, i), o, stereo, e, Jp, Jl, NULL);
        const double w = (double)P->inv_sigma_sq[i];
        double chi = 0;
        for (int d = 0; d < dim; ++d) chi += w * e[d] * e[d];
        double rho1 = 1.0;
        if (P->use_huber) { double rho[3]; huber(chi, P->delta, rho); rho1 = rho[1]; }
        const double ww = rho1 * w; /* robustInformation = rho[1] * Omega ; omega_r = -rho[1] * Omega * e */
        if (fi >= 0) {
            for (int a = 0; a < 6; ++a) {
                double ba = 0;
                for (int d = 0; d < dim; ++d) ba -= Jp[6 * d + a] * ww * e[d];
                Y->bp[6 * fi + a] += ba;
                for (int b = 0; b < 6; ++b) {
                    double h = 0;
                    for (int d = 0; d < dim; ++d) h += Jp[6 * d + a] * ww * Jp[6 * d + b];
                    Y->Hpp[(size_t)(6 * fi + a) * n + 6 * fi + b] += h;
                }
            }
        }
        if (P->points) {
            const int lm = P->obs_lm[i];
            for (int a = 0; a < 3; ++a) {
                double ba = 0;
                for (int d = 0; d < dim; ++d) ba -= Jl[3 * d + a] * ww * e[d];
                Y->bl[3 * (size_t)lm + a] += ba;
                for (int b = 0; b < 3; ++b) {
                    double h = 0;
                    for (int d = 0; d < dim; ++d) h += Jl[3 * d + a] * ww * Jl[3 * d + b];
                    Y->Hll[9 * (size_t)lm + 3 * a + b] += h;
                }
            }
            if (fi >= 0) {
                double* W = Y->Hpl + 18 * (size_t)i;
                for (int a = 0; a < 6; ++a)
                    for (int b = 0; b < 3; ++b) {
                        double h = 0;
                        for (int d = 0; d < dim; ++d) h += Jp[6 * d + a] * ww * Jl[3 * d + b];
                        W[3 * a + b] = h;
                    }
            }
        }
    }
}

/* BlockSolver::solve with lambda on every diagonal: Schur complement, dense Cholesky,
 * back-substitution.  Returns 0 if ok. */
static int solve_system(ob_problem* P, ob_system* Y, double lambda, const int* lm_first, const int* lm_edges) {
    const int n = Y->n;
    memcpy(Y->S, Y->Hpp, sizeof(double) * (size_t)n * n);
    memcpy(Y->bS, Y->bp, sizeof(double) * n);
    for (int d = 0; d < n; ++d) Y->S[(size_t)d * n + d] += lambda;
    if (P->points) {
        for (int l = 0; l < P->L; ++l) {
            double D[9];
            memcpy(D, Y->Hll + 9 * (size_t)l, sizeof(D));
            D[0] += lambda; D[4] += lambda; D[8] += lambda;
            double* Di = Y->Dinv + 9 * (size_t)l;
            if (inv3_sym(D, Di) != 0) return -1;
            const double* bl = Y->bl + 3 * (size_t)l;
            double z[3];
            mat3_mul_vec(Di, bl, z);
            /* edges of this landmark on free poses */
            for (int p = lm_first[l]; p < lm_first[l + 1]; ++p) {
                const int i = lm_edges[p];
                const int fi = P->free_idx[P->obs_kf[i]];
                if (P->level[i] || fi < 0) continue;
                const double* Wi = Y->Hpl + 18 * (size_t)i;
                double Yi[18]; /* Wi * Dinv */
                for (int a = 0; a < 6; ++a)
                    for (int b = 0; b < 3; ++b) Yi[3 * a + b] = Wi[3 * a] * Di[b] + Wi[3 * a + 1] * Di[3 + b] + Wi[3 * a + 2] * Di[6 + b];
                for (int a = 0; a < 6; ++a) Y->bS[6 * fi + a] -= Wi[3 * a] * z[0] + Wi[3 * a + 1] * z[1] + Wi[3 * a + 2] * z[2];
                for (int q = lm_first[l]; q < lm_first[l + 1]; ++q) {
                    const int j = lm_edges[q];
                    const int fj = P->free_idx[P->obs_kf[j]];
                    if (P->level[j] || fj < 0) continue;
                    const double* Wj = Y->Hpl + 18 * (size_t)j;
                    for (int a = 0; a < 6; ++a)
                        for (int b = 0; b < 6; ++b)
                            Y->S[(size_t)(6 * fi + a) * n + 6 * fj + b] -= Yi[3 * a] * Wj[3 * b] + Yi[3 * a + 1] * Wj[3 * b + 1] + Yi[3 * a + 2] * Wj[3 * b + 2];
                }
            }
        }
    }
    if (n > 0 && chol_solve(Y->S, n, Y->bS, Y->x) != 0) return -1;
    if (P->points) {
        for (int l = 0; l < P->L; ++l) {
            double r[3] = {Y->bl[3 * (size_t)l], Y->bl[3 * (size_t)l + 1], Y->bl[3 * (size_t)l + 2]};
            for (int p = lm_first[l]; p < lm_first[l + 1]; ++p) {
                const int i = lm_edges[p];
                const int fi = P->free_idx[P->obs_kf[i]];
                if (P->level[i] || fi < 0) continue;
                const double* Wi = Y->Hpl + 18 * (size_t)i;
                for (int b = 0; b < 3; ++b)
                    for (int a = 0; a < 6; ++a) r[b] -= Wi[3 * a + b] * Y->x[6 * fi + a];
            }
            mat3_mul_vec(Y->Dinv + 9 * (size_t)l, r, Y->xl + 3 * (size_t)l);
        }
    }
    return 0;
}

/* SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg.  Returns the number
 * of iterations executed. */
static int lm_optimize(ob_problem* P, int iterations, ob_stats* st) {
    const int n = 6 * P->nfree;
    ob_system Y; memset(&Y, 0, sizeof(Y));
    Y.n = n;
    Y.Hpp = (double*)calloc((size_t)n * n + 1, sizeof(double)); Y.S = (double*)calloc((size_t)n * n + 1, sizeof(double));
    Y.bp = (double*)calloc(n + 1, sizeof(double)); Y.bS = (double*)calloc(n + 1, sizeof(double)); Y.x = (double*)calloc(n + 1, sizeof(double));
    int* lm_first = NULL; int* lm_edges = NULL;
    if (P->points) {
        Y.Hll = (double*)calloc(9 * (size_t)P->L + 1, sizeof(double)); Y.bl = (double*)calloc(3 * (size_t)P->L + 1, sizeof(double));
        Y.Hpl = (double*)calloc(18 * (size_t)P->M + 1, sizeof(double)); Y.xl = (double*)calloc(3 * (size_t)P->L + 1, sizeof(double));
        Y.Dinv = (double*)calloc(9 * (size_t)P->L + 1, sizeof(double));
        lm_first = (int*)calloc((size_t)P->L + 2, sizeof(int)); lm_edges = (int*)calloc((size_t)P->M + 1, sizeof(int));
        for (int i = 0; i < P->M; ++i) lm_first[P->obs_lm[i] + 1]++;
        for (int l = 0; l < P->L; ++l) lm_first[l + 1] += lm_first[l];
        int* pos = (int*)malloc(sizeof(int) * ((size_t)P->L + 1));
        memcpy(pos, lm_first, sizeof(int) * ((size_t)P->L + 1));
        for (int i = 0; i < P->M; ++i) lm_edges[pos[P->obs_lm[i]]++] = i;
        free(pos);
    }
    double* poses_bak = (double*)malloc(sizeof(double) * 12 * (size_t)P->K);
    double* points_bak = P->points ? (double*)malloc(sizeof(double) * 3 * (size_t)P->L) : NULL;

    double lambda = 0; double ni = 2;
    int it = 0, ok = 1;
    for (; it < iterations && ok; ++it) {
        if (P->force_stop && *P->force_stop) break;
        double currentChi = compute_active_errors(P, P->poses, P->points);
        double tempChi = currentChi;
        build_system(P, &Y);
        if (it == 0) {
            /* computeLambdaInit: tau * max |H_jj| over all vertices */
            double maxd = 0;
            for (int d = 0; d < n; ++d) maxd = fmax(fabs(Y.Hpp[(size_t)d * n + d]), maxd);
            if (P->points)
                for (int l = 0; l < P->L; ++l)
                    for (int d = 0; d < 3; ++d) maxd = fmax(fabs(Y.Hll[9 * (size_t)l + 4 * d]), maxd);
            lambda = 1e-5 * maxd;
            ni = 2;
            if (st && st->num_rounds < OB_MAX_ROUNDS) st->lambda_init[st->num_rounds] = lambda;
        }
        double rho = 0;
        int qmax = 0;
        do {
            memcpy(poses_bak, P->poses, sizeof(double) * 12 * (size_t)P->K);            /* push */
            if (P->points) memcpy(points_bak, P->points, sizeof(double) * 3 * (size_t)P->L);
            const int ok2 = solve_system(P, &Y, lambda, lm_first, lm_edges) == 0;
            if (ok2) {                                                                   /* update */
                for (int k = 0; k < P->K; ++k) {
                    const int fi = P->free_idx[k];
                    if (fi < 0) continue;
                    double np[12];
                    ob_pose_oplus(P->poses + 12 * (size_t)k, Y.x + 6 * fi, np);
                    memcpy(P->poses + 12 * (size_t)k, np, sizeof(np));
                }
                if (P->points) for (size_t j = 0; j < 3 * (size_t)P->L; ++j) P->points[j] += Y.xl[j];
            }
            tempChi = compute_active_errors(P, P->poses, P->points);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0;
            if (ok2) {
                for (int d = 0; d < n; ++d) scale += Y.x[d] * (lambda * Y.x[d] + Y.bp[d]);
                if (P->points) for (size_t j = 0; j < 3 * (size_t)P->L; ++j) scale += Y.xl[j] * (lambda * Y.xl[j] + Y.bl[j]);
            }
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                const double scaleFactor = fmax(1. / 3., alpha);
                lambda *= scaleFactor;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                memcpy(P->poses, poses_bak, sizeof(double) * 12 * (size_t)P->K);        /* pop */
                if (P->points) memcpy(P->points, points_bak, sizeof(double) * 3 * (size_t)P->L);
            }
            qmax++;
            if (st) st->num_trials++;
        } while (rho < 0 && qmax < 10 && !(P->force_stop && *P->force_stop));
        if (st) { st->last_chi2 = currentChi; st->last_lambda = lambda; }
        if (getenv("OB_TRACE_TRIALS")) fprintf(stderr, "ob_lm: iteration %d trials %d lambda %.3g chi2 %.6f\n", it, qmax, lambda, currentChi);
        if (qmax == 10 || rho == 0) ok = 0; /* Terminate */
    }
    if (st) { st->num_iterations += it; if (st->num_rounds < OB_MAX_ROUNDS) st->round_iterations[st->num_rounds] = it; st->num_rounds++; }
    free(Y.Hpp); free(Y.S); free(Y.bp); free(Y.bS); free(Y.x); free(Y.Hll); free(Y.bl); free(Y.Hpl); free(Y.xl); free(Y.Dinv);
    free(lm_first); free(lm_edges); free(poses_bak); free(points_bak);
    return it;
}

/* ------------------------------------------------------------------ pose_optimizer::optimize */
int ob_pose_optimize(const ob_camera* cam, int setup_is_mono, int n, const double* pts_w, const float* obs_xy,
                     const float* obs_xr, const float* inv_sigma_sq, double* pose_cw, uint8_t* outlier_flags,
                     int num_trials, int num_each_iter, ob_stats* st) {
    if (st) memset(st, 0, sizeof(*st));
    for (int i = 0; i < n; ++i) outlier_flags[i] = 0;
    if (n < 5) return 0;
    const float chi_sq_2D = 5.99146f, chi_sq_3D = 7.81473f;
    const float sqrt_chi_sq = setup_is_mono ? sqrtf(chi_sq_2D) : sqrtf(chi_sq_3D);
    ob_problem P; memset(&P, 0, sizeof(P));
    uint8_t fixed0 = 0; int free0 = 0;
    int* kf = (int*)calloc((size_t)n + 1, sizeof(int));
    P.cam = cam; P.K = 1; P.L = 0; P.M = n; P.poses = pose_cw; P.fixed = &fixed0; P.points = NULL; P.pts_const = pts_w;
    P.obs_kf = kf; P.obs_lm = NULL; P.obs_xy = obs_xy; P.obs_xr = obs_xr; P.inv_sigma_sq = inv_sigma_sq;
    P.level = (uint8_t*)calloc((size_t)n + 1, 1); P.err = (double*)calloc(3 * (size_t)n + 1, sizeof(double));
    P.free_idx = &free0; P.nfree = 1; P.use_huber = 1; P.delta = (double)sqrt_chi_sq; P.force_stop = NULL;
    /* per-edge robust kernel switch: the reference drops the kernel of every edge after the
     * round `num_trials - 2`, so it is a per-round switch */
    int num_bad = 0;
    for (int trial = 0; trial < num_trials; ++trial) {
        lm_optimize(&P, num_each_iter, st);
        num_bad = 0;
        for (int i = 0; i < n; ++i) {
            if (outlier_flags[i]) { /* edge->computeError() */
                double o[3]; edge_obs(&P, i, o);
                double* e = P.err + 3 * (size_t)i; e[2] = 0;
                ob_edge_eval(cam, pose_cw, pts_w + 3 * (size_t)i, o, edge_is_stereo(&P, i), e, NULL, NULL, NULL);
            }
            const double thr = edge_is_stereo(&P, i) ? (double)chi_sq_3D : (double)chi_sq_2D;
            if (thr < edge_chi2(&P, i)) { outlier_flags[i] = 1; P.level[i] = 1; ++num_bad; }
            else { outlier_flags[i] = 0; P.level[i] = 0; }
        }
        if (trial == num_trials - 2) P.use_huber = 0;
        if (n - num_bad < 5) break;
    }
    if (st) { st->final_chi2 = 0; for (int i = 0; i < n; ++i) if (!P.level[i]) st->final_chi2 += edge_chi2(&P, i); }
    free(kf); free(P.level); free(P.err);
    return n - num_bad;
}

/* --------------------------------------------------------- local_bundle_adjuster::optimize */
int ob_local_ba(const ob_camera* cam, int setup_is_mono, int K, double* poses, const uint8_t* fixed, int L, double* points,
                int M, const int* obs_kf, const int* obs_lm, const float* obs_xy, const float* obs_xr,
                const float* inv_sigma_sq, int num_first_iter, int num_second_iter, const volatile int* force_stop,
                uint8_t* outlier_out, ob_stats* st) {
    if (st) memset(st, 0, sizeof(*st));
    const float chi_sq_2D = 5.99146f, chi_sq_3D = 7.81473f;
    const float sqrt_chi_sq = setup_is_mono ? sqrtf(chi_sq_2D) : sqrtf(chi_sq_3D);
    ob_problem P; memset(&P, 0, sizeof(P));
    P.cam = cam; P.K = K; P.L = L; P.M = M; P.poses = poses; P.fixed = fixed; P.points = points; P.pts_const = NULL;
    P.obs_kf = obs_kf; P.obs_lm = obs_lm; P.obs_xy = obs_xy; P.obs_xr = obs_xr; P.inv_sigma_sq = inv_sigma_sq;
    P.level = (uint8_t*)calloc((size_t)M + 1, 1); P.err = (double*)calloc(3 * (size_t)M + 1, sizeof(double));
    P.free_idx = (int*)malloc(sizeof(int) * ((size_t)K + 1));
    P.nfree = 0;
    for (int k = 0; k < K; ++k) P.free_idx[k] = fixed[k] ? -1 : P.nfree++;
    P.use_huber = 1; P.delta = (double)sqrt_chi_sq; P.force_stop = force_stop;
    for (int i = 0; i < M; ++i) outlier_out[i] = 0;

    if (!(force_stop && *force_stop)) {
        lm_optimize(&P, num_first_iter, st);
        int run_robust_BA = 1;
        if (force_stop && *force_stop) run_robust_BA = 0;
        if (run_robust_BA) {
            for (int i = 0; i < M; ++i) {
                const int stereo = edge_is_stereo(&P, i);
                const double thr = stereo ? (double)chi_sq_3D : (double)chi_sq_2D;
                int depth_pos = 1;
                if (cam->model != OB_CAM_EQUIRECTANGULAR) {
                    const double* ps = poses + 12 * (size_t)obs_kf[i]; const double* pw = points + 3 * (size_t)obs_lm[i];
                    depth_pos = (ps[6] * pw[0] + ps[7] * pw[1] + ps[8] * pw[2] + ps[11]) > 0;
                }
                if (thr < edge_chi2(&P, i) || !depth_pos) P.level[i] = 1;
            }
            P.use_huber = 0;
            lm_optimize(&P, num_second_iter, st);
        }
        for (int i = 0; i < M; ++i) {
            const int stereo = edge_is_stereo(&P, i);
            const double thr = stereo ? (double)chi_sq_3D : (double)chi_sq_2D;
            int depth_pos = 1;
            if (cam->model != OB_CAM_EQUIRECTANGULAR) {
                const double* ps = poses + 12 * (size_t)obs_kf[i]; const double* pw = points + 3 * (size_t)obs_lm[i];
                depth_pos = (ps[6] * pw[0] + ps[7] * pw[1] + ps[8] * pw[2] + ps[11]) > 0;
            }
            if (thr < edge_chi2(&P, i) || !depth_pos) outlier_out[i] = 1;
        }
    }
    if (st) { st->final_chi2 = 0; for (int i = 0; i < M; ++i) if (!P.level[i]) st->final_chi2 += edge_chi2(&P, i); }
    free(P.level); free(P.err); free(P.free_idx);
    return 0;
}


/* optimize::global_bundle_adjuster::optimize (optimize/global_bundle_adjuster.cc, as recalled) -- SURVEY.md 8f rank 4, oracle
 * only so far: every keyframe and landmark of the map, the origin keyframe(s) fixed (`fixed`), ONE Levenberg round of
 * num_iter iterations (default 10), Huber kernel on every edge when use_huber_kernel (default true), no outlier
 * classification; the caller writes the result back as pose_cw_after_loop_BA_ / pos_w_after_global_BA_.  Same graph arrays
 * as ob_local_ba. */
int ob_global_ba(const ob_camera* cam, int setup_is_mono, int K, double* poses, const uint8_t* fixed, int L, double* points,
                 int M, const int* obs_kf, const int* obs_lm, const float* obs_xy, const float* obs_xr,
                 const float* inv_sigma_sq, int num_iter, int use_huber_kernel, const volatile int* force_stop, ob_stats* st) {
    if (st) memset(st, 0, sizeof(*st));
    const float chi_sq_2D = 5.99146f, chi_sq_3D = 7.81473f;
    const float sqrt_chi_sq = setup_is_mono ? sqrtf(chi_sq_2D) : sqrtf(chi_sq_3D);
    ob_problem P; memset(&P, 0, sizeof(P));
    P.cam = cam; P.K = K; P.L = L; P.M = M; P.poses = poses; P.fixed = fixed; P.points = points; P.pts_const = NULL;
    P.obs_kf = obs_kf; P.obs_lm = obs_lm; P.obs_xy = obs_xy; P.obs_xr = obs_xr; P.inv_sigma_sq = inv_sigma_sq;
    P.level = (uint8_t*)calloc((size_t)M + 1, 1); P.err = (double*)calloc(3 * (size_t)M + 1, sizeof(double));
    P.free_idx = (int*)malloc(sizeof(int) * ((size_t)K + 1));
    P.nfree = 0;
    for (int k = 0; k < K; ++k) P.free_idx[k] = fixed[k] ? -1 : P.nfree++;
    P.use_huber = use_huber_kernel ? 1 : 0; P.delta = (double)sqrt_chi_sq; P.force_stop = force_stop;
    if (!(force_stop && *force_stop)) lm_optimize(&P, num_iter, st);
    if (st) { st->final_chi2 = 0; for (int i = 0; i < M; ++i) st->final_chi2 += edge_chi2(&P, i); }
    free(P.level); free(P.err); free(P.free_idx);
    return 0;
}
