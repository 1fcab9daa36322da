// Consumers of the marker table, batched on the device: what MamriLogic.process does
// with "DetectedFiducials" after the segmentation (Mamri/Mamri.py:858-870):
//   joint_detection              :1343-1363  L-shape triplets per marker-bearing link, first matching
//                                            3-combination in node order wins, its points are consumed
//   _sort_l_shaped_markers       :1782-1792  corner, short arm, long arm
//   baseplate y-flatten          :1371-1373
//   _calculate_fiducial_alignment_matrix :1771-1780  vtkLandmarkTransform, rigid (Horn's unit quaternion),
//                                            landmarks rounded to float32 as vtkPoints stores them
//   _solve_full_chain_ik         :1410-1447  bounded least squares on the effector (+ weighted secondary) markers
//   _get_world_transform_for_joint / _full_chain_ik_error_function :1486-1536  forward kinematics, residuals
//
// One warp per scan.  Matching is warp-parallel over the 3-combinations (the lexicographically first match =
// itertools.combinations order is a min-reduction over the combination rank); registration and the IK run
// on lane 0: 4x4 Jacobi eigen-solve, then the reference's own algorithm -- Trust Region Reflective least squares
// with a 2-point finite-difference Jacobian, ftol = xtol = 1e-6 -- restated step by step (trf_solve), so that it
// takes SciPy's iterates and ends in SciPy's minimum (the objective has several; tests: >= 95 % of scenes within
// 1e-5 rad of SciPy, the rest are ill-conditioned fits where rounding moves the stopping point).
#include "common.cuh"

namespace {

struct M34 { double r[9]; double t[3]; };      // rigid transform: rotation row-major + translation

__device__ __forceinline__ M34 mul(const M34& a, const M34& b) {
    M34 o;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o.r[3 * i + j] = a.r[3 * i] * b.r[j] + a.r[3 * i + 1] * b.r[3 + j] + a.r[3 * i + 2] * b.r[6 + j];
        o.t[i] = a.r[3 * i] * b.t[0] + a.r[3 * i + 1] * b.t[1] + a.r[3 * i + 2] * b.t[2] + a.t[i];
    }
    return o;
}

// fixed_offset @ articulation of one link: translate, then RotateZ(a) / RotateY(-a) / RotateX(a)  (Mamri.py:1760-1769)
__device__ __forceinline__ M34 local_tf(const mamri_link& l, double ang) {
    M34 m;
    for (int i = 0; i < 9; ++i) m.r[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 3; ++i) m.t[i] = l.translate[i];
    double s, c;
    if (l.axis == MAMRI_AXIS_IS) { sincos(ang, &s, &c); m.r[0] = c; m.r[1] = -s; m.r[3] = s; m.r[4] = c; }
    else if (l.axis == MAMRI_AXIS_PA) { sincos(-ang, &s, &c); m.r[0] = c; m.r[2] = s; m.r[6] = -s; m.r[8] = c; }
    else if (l.axis == MAMRI_AXIS_LR) { sincos(ang, &s, &c); m.r[4] = c; m.r[5] = -s; m.r[7] = s; m.r[8] = c; }
    return m;
}

__device__ __forceinline__ double dist3(const double* a, const double* b) {
    const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

__device__ __forceinline__ void sort3(double& a, double& b, double& c) {
    double t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
}

// Cyclic Jacobi on a symmetric n x n matrix (n <= 4); eigenvectors in the columns of v.
template <int N>
__device__ void jacobi_eig(double (&a)[N][N], double (&v)[N][N], double (&w)[N]) {
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < N; ++i) for (int j = i + 1; j < N; ++j) off += a[i][j] * a[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < N; ++p)
            for (int q = p + 1; q < N; ++q) {
                if (fabs(a[p][q]) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
                for (int k = 0; k < N; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
                for (int k = 0; k < N; ++k) { const double x = v[k][p], y = v[k][q]; v[k][p] = c * x - s * y; v[k][q] = s * x + c * y; }
            }
    }
    for (int i = 0; i < N; ++i) w[i] = a[i][i];
}

// vtkLandmarkTransform, RigidBody mode, on 3 point pairs (float32-rounded landmarks).
__device__ M34 landmark_rigid(const double* src9, const double* tgt9) {
    double s[3][3], t[3][3], sc[3] = {0, 0, 0}, tc[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) { s[i][k] = double(float(src9[3 * i + k])); t[i][k] = double(float(tgt9[3 * i + k])); }
    for (int k = 0; k < 3; ++k) { sc[k] = (s[0][k] + s[1][k] + s[2][k]) / 3.0; tc[k] = (t[0][k] + t[1][k] + t[2][k]) / 3.0; }
    double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int p = 0; p < 3; ++p)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) m[i][j] += (s[p][i] - sc[i]) * (t[p][j] - tc[j]);
    double n[4][4] = {
        {m[0][0] + m[1][1] + m[2][2], m[1][2] - m[2][1], m[2][0] - m[0][2], m[0][1] - m[1][0]},
        {m[1][2] - m[2][1], m[0][0] - m[1][1] - m[2][2], m[0][1] + m[1][0], m[2][0] + m[0][2]},
        {m[2][0] - m[0][2], m[0][1] + m[1][0], -m[0][0] + m[1][1] - m[2][2], m[1][2] + m[2][1]},
        {m[0][1] - m[1][0], m[2][0] + m[0][2], m[1][2] + m[2][1], -m[0][0] - m[1][1] + m[2][2]}};
    double v[4][4], w[4];
    jacobi_eig<4>(n, v, w);
    int best = 0;
    for (int i = 1; i < 4; ++i) if (w[i] > w[best]) best = i;
    const double qw = v[0][best], qx = v[1][best], qy = v[2][best], qz = v[3][best];
    M34 o;
    o.r[0] = qw * qw + qx * qx - qy * qy - qz * qz; o.r[1] = 2 * (qx * qy - qw * qz); o.r[2] = 2 * (qx * qz + qw * qy);
    o.r[3] = 2 * (qx * qy + qw * qz); o.r[4] = qw * qw - qx * qx + qy * qy - qz * qz; o.r[5] = 2 * (qy * qz - qw * qx);
    o.r[6] = 2 * (qx * qz - qw * qy); o.r[7] = 2 * (qy * qz + qw * qx); o.r[8] = qw * qw - qx * qx - qy * qy + qz * qz;
    for (int i = 0; i < 3; ++i) o.t[i] = tc[i] - (o.r[3 * i] * sc[0] + o.r[3 * i + 1] * sc[1] + o.r[3 * i + 2] * sc[2]);
    return o;
}

struct IkProblem {
    const mamri_robot* rb;
    M34 base;
    int n_sets;                 // 1 (effector) or 2 (+ secondary)
    int link[2];
    double weight[2];
    double local[2][9];         // marker coordinates in the link frame (effector: optionally turned 180 deg about z)
    double target[2][9];
    int n_unknowns;
    double lo[MAMRI_MAX_CHAIN], hi[MAMRI_MAX_CHAIN];
};

// residuals r[9 * n_sets] and Jacobian J[row][unknown] at x
__device__ void ik_eval(const IkProblem& P, const double* x, double* r, double (*J)[MAMRI_MAX_CHAIN]) {
    const mamri_robot& rb = *P.rb;
    M34 world[MAMRI_MAX_LINKS];
    for (int l = 0; l < rb.n_links; ++l) {
        const mamri_link& L = rb.links[l];
        const double ang = (L.chain_index >= 0 && L.chain_index < P.n_unknowns) ? x[L.chain_index] : 0.0;
        world[l] = mul(L.parent >= 0 ? world[L.parent] : P.base, local_tf(L, ang));
    }
    for (int s = 0; s < P.n_sets; ++s) {
        const M34& T = world[P.link[s]];
        for (int i = 0; i < 3; ++i) {
            const double* lp = &P.local[s][3 * i];
            double p[3];
            for (int k = 0; k < 3; ++k) p[k] = T.r[3 * k] * lp[0] + T.r[3 * k + 1] * lp[1] + T.r[3 * k + 2] * lp[2] + T.t[k];
            const int row = 9 * s + 3 * i;
            for (int k = 0; k < 3; ++k) r[row + k] = P.weight[s] * (p[k] - P.target[s][3 * i + k]);
            if (!J) continue;
            for (int k = 0; k < 3; ++k) for (int u = 0; u < P.n_unknowns; ++u) J[row + k][u] = 0.0;
            for (int a = P.link[s]; a >= 0; a = rb.links[a].parent) {      // joints between the root and this link
                const mamri_link& A = rb.links[a];
                if (A.chain_index < 0 || A.chain_index >= P.n_unknowns) continue;
                double w[3];
                if (A.axis == MAMRI_AXIS_IS) for (int k = 0; k < 3; ++k) w[k] = world[a].r[3 * k + 2];
                else if (A.axis == MAMRI_AXIS_PA) for (int k = 0; k < 3; ++k) w[k] = -world[a].r[3 * k + 1];
                else if (A.axis == MAMRI_AXIS_LR) for (int k = 0; k < 3; ++k) w[k] = world[a].r[3 * k];
                else continue;
                const double d[3] = {p[0] - world[a].t[0], p[1] - world[a].t[1], p[2] - world[a].t[2]};
                J[row][A.chain_index] = P.weight[s] * (w[1] * d[2] - w[2] * d[1]);
                J[row + 1][A.chain_index] = P.weight[s] * (w[2] * d[0] - w[0] * d[2]);
                J[row + 2][A.chain_index] = P.weight[s] * (w[0] * d[1] - w[1] * d[0]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Trust Region Reflective least squares: the solver the reference calls (Mamri.py:1430-1433,
// scipy.optimize.least_squares(method='trf', bounds, ftol = xtol = 1e-6), defaults otherwise: gtol 1e-8, x_scale 1,
// '2-point' finite-difference Jacobian, exact trust-region sub-problem on the SVD of the augmented Jacobian).
// The objective has several local minima and, for three markers on a six-joint chain, several exact solutions of
// equal cost; which one a solver started at zero ends in is decided by its iterates, so this follows the published
// algorithm (Branch, Coleman & Li 1999; More 1977 for the sub-problem; scipy/optimize/_lsq/trf.py trf_bounds,
// common.py select_step / solve_lsq_trust_region / CL_scaling_vector / update_tr_radius / check_termination) step
// by step, including the finite-difference Jacobian, so that its iterates are SciPy's up to rounding.  The SVD is a
// one-sided Jacobi (Hestenes) on the (m + n) x n augmented matrix.  n <= MAMRI_MAX_CHAIN unknowns, m <= 18 residuals.
// ------------------------------------------------------------------------------------------------
constexpr int TN = MAMRI_MAX_CHAIN, TM = 18, TA = TM + TN;
constexpr double T_EPS = 2.220446049250313e-16;

__device__ double vnorm(const double* v, int n) { double s = 0.0; for (int i = 0; i < n; ++i) s += v[i] * v[i]; return sqrt(s); }

// 2-point finite differences as scipy.optimize._numdiff.approx_derivative does them: h = sqrt(eps) * sign(x) *
// max(1, |x|), turned around where it would leave the bounds, J[:, i] = (f(x + h e_i) - f0) / ((x_i + h) - x_i).
__device__ void trf_jac(const IkProblem& P, const double* x, const double* f0, int n, int m, double (*J)[TN]) {
    const double rel = 1.4901161193847656e-08;
    for (int i = 0; i < n; ++i) {
        double h = rel * (x[i] >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(x[i]));
        const double xh = x[i] + h;
        if (xh < P.lo[i] || xh > P.hi[i]) {
            const double lower = x[i] - P.lo[i], upper = P.hi[i] - x[i];
            if (x[i] - h >= P.lo[i] && x[i] - h <= P.hi[i]) h = -h;
            else if (upper >= lower) h = upper;
            else h = -lower;
        }
        double x1[TN], f1[TM];
        for (int u = 0; u < n; ++u) x1[u] = x[u];
        x1[i] = x[i] + h;
        const double dx = x1[i] - x[i];
        ik_eval(P, x1, f1, nullptr);
        for (int r = 0; r < m; ++r) J[r][i] = (f1[r] - f0[r]) / dx;
    }
}

// min over i of the step t > 0 that takes x + t s onto a bound (inf if s == 0); hits[i] = sign(s_i) where that
// minimum is attained (common.step_size_to_bound)
__device__ double trf_step_to_bound(const double* x, const double* s, const double* lo, const double* hi, int n, int* hits) {
    double steps[TN], ms = INFINITY;
    for (int i = 0; i < n; ++i) {
        steps[i] = INFINITY;
        if (s[i] != 0.0) steps[i] = fmax((lo[i] - x[i]) / s[i], (hi[i] - x[i]) / s[i]);
        ms = fmin(ms, steps[i]);
    }
    if (hits) for (int i = 0; i < n; ++i) hits[i] = steps[i] == ms ? (s[i] > 0.0 ? 1 : (s[i] < 0.0 ? -1 : 0)) : 0;
    return ms;
}

struct TrfQuad {                      // the quadratic model in the scaled ("hat") variables
    const double (*Jh)[TN];
    const double* gh;
    const double* diag;
    int n, m;
    __device__ void Js(const double* s, double* out) const {
        for (int r = 0; r < m; ++r) { double a = 0.0; for (int u = 0; u < n; ++u) a += Jh[r][u] * s[u]; out[r] = a; }
    }
    __device__ double value(const double* s) const {                                   // evaluate_quadratic
        double v[TM], q = 0.0, l = 0.0;
        Js(s, v);
        for (int r = 0; r < m; ++r) q += v[r] * v[r];
        for (int u = 0; u < n; ++u) { q += s[u] * diag[u] * s[u]; l += s[u] * gh[u]; }
        return 0.5 * q + l;
    }
    __device__ void line(const double* s, const double* s0, double& a, double& b, double& c) const {   // build_quadratic_1d
        double v[TM], u0[TM];
        Js(s, v);
        a = 0.0; b = 0.0; c = 0.0;
        for (int r = 0; r < m; ++r) a += v[r] * v[r];
        for (int u = 0; u < n; ++u) { a += s[u] * diag[u] * s[u]; b += gh[u] * s[u]; }
        a *= 0.5;
        if (s0) {
            Js(s0, u0);
            double uu = 0.0, uv = 0.0, g0 = 0.0, sd = 0.0, s0d = 0.0;
            for (int r = 0; r < m; ++r) { uu += u0[r] * u0[r]; uv += u0[r] * v[r]; }
            for (int u = 0; u < n; ++u) { g0 += gh[u] * s0[u]; sd += s0[u] * diag[u] * s[u]; s0d += s0[u] * diag[u] * s0[u]; }
            b += uv + sd;
            c = 0.5 * uu + g0 + 0.5 * s0d;
        }
    }
};

// minimize_quadratic_1d: a t^2 + b t + c on [lo, hi]; candidates in SciPy's order (lo, hi, interior extremum), first minimum wins
__device__ void trf_minq(double a, double b, double lo, double hi, double c, double& t_best, double& y_best) {
    t_best = lo; y_best = lo * (a * lo + b) + c;
    const double yh = hi * (a * hi + b) + c;
    if (yh < y_best) { t_best = hi; y_best = yh; }
    if (a != 0.0) {
        const double e = -0.5 * b / a;
        if (lo < e && e < hi) { const double ye = e * (a * e + b) + c; if (ye < y_best) { t_best = e; y_best = ye; } }
    }
}

// One-sided Jacobi SVD of A (rows x n): on return the columns of A are U diag(s) sorted by descending s, V the
// right singular vectors in the same order.
__device__ void trf_svd(double (*A)[TN], int rows, int n, double* sv, double (*V)[TN]) {
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rot = 0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int r = 0; r < rows; ++r) { al += A[r][p] * A[r][p]; be += A[r][q] * A[r][q]; ga += A[r][p] * A[r][q]; }
                if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
                ++rot;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int r = 0; r < rows; ++r) { const double x = A[r][p], y = A[r][q]; A[r][p] = c * x - sn * y; A[r][q] = sn * x + c * y; }
                for (int r = 0; r < n; ++r) { const double x = V[r][p], y = V[r][q]; V[r][p] = c * x - sn * y; V[r][q] = sn * x + c * y; }
            }
        if (!rot) break;
    }
    for (int j = 0; j < n; ++j) { double a = 0.0; for (int r = 0; r < rows; ++r) a += A[r][j] * A[r][j]; sv[j] = sqrt(a); }
    for (int i = 0; i < n - 1; ++i) {                    // selection sort, descending, stable enough for distinct values
        int best = i;
        for (int j = i + 1; j < n; ++j) if (sv[j] > sv[best]) best = j;
        if (best != i) {
            double t = sv[i]; sv[i] = sv[best]; sv[best] = t;
            for (int r = 0; r < rows; ++r) { t = A[r][i]; A[r][i] = A[r][best]; A[r][best] = t; }
            for (int r = 0; r < n; ++r) { t = V[r][i]; V[r][i] = V[r][best]; V[r][best] = t; }
        }
    }
}

// solve_lsq_trust_region with suf = s * (U^T f): min |J p + f| subject to |p| <= Delta
__device__ void trf_subproblem(int n, int m, const double* suf, const double* sv, const double (*V)[TN], double Delta,
                               double& alpha, double* p) {
    auto phi = [&](double a, double& dphi) {
        double pn = 0.0, d3 = 0.0;
        for (int i = 0; i < n; ++i) { const double den = sv[i] * sv[i] + a; pn += (suf[i] / den) * (suf[i] / den); d3 += suf[i] * suf[i] / (den * den * den); }
        pn = sqrt(pn);
        dphi = -d3 / pn;
        return pn - Delta;
    };
    auto step = [&](double a) {
        for (int r = 0; r < n; ++r) { double acc = 0.0; for (int i = 0; i < n; ++i) acc += V[r][i] * (suf[i] / (sv[i] * sv[i] + a)); p[r] = -acc; }
    };
    const bool full_rank = m >= n && sv[n - 1] > T_EPS * m * sv[0];
    if (full_rank) {
        step(0.0);
        if (vnorm(p, n) <= Delta) { alpha = 0.0; return; }
    }
    double au = vnorm(suf, n) / Delta, al = 0.0, dphi;
    if (full_rank) { const double ph = phi(0.0, dphi); al = -ph / dphi; }
    if (!full_rank && alpha == 0.0) alpha = fmax(0.001 * au, sqrt(al * au));
    for (int it = 0; it < 10; ++it) {
        if (alpha < al || alpha > au) alpha = fmax(0.001 * au, sqrt(al * au));
        const double ph = phi(alpha, dphi);
        if (ph < 0.0) au = alpha;
        const double ratio = ph / dphi;
        al = fmax(al, alpha - ratio);
        alpha -= (ph + Delta) * ratio / Delta;
        if (fabs(ph) < 0.01 * Delta) break;
    }
    step(alpha);
    const double pn = vnorm(p, n);
    for (int r = 0; r < n; ++r) p[r] *= Delta / pn;
}

// trf.select_step: the trust-region step, its reflection at the first bound it hits, or the constrained Cauchy step
__device__ double trf_select_step(const double* x, const TrfQuad& Q, const double* p_in, const double* ph_in, const double* d,
                                  double Delta, const double* lo, const double* hi, double theta, double* step, double* step_h) {
    const int n = Q.n;
    double p[TN], ph[TN], xp[TN];
    bool inside = true;
    for (int i = 0; i < n; ++i) { p[i] = p_in[i]; ph[i] = ph_in[i]; xp[i] = x[i] + p[i]; inside = inside && xp[i] >= lo[i] && xp[i] <= hi[i]; }
    if (inside) {
        for (int i = 0; i < n; ++i) { step[i] = p[i]; step_h[i] = ph[i]; }
        return -Q.value(ph);
    }
    int hits[TN];
    const double p_stride = trf_step_to_bound(x, p, lo, hi, n, hits);
    double rh[TN], r[TN], xb[TN];
    for (int i = 0; i < n; ++i) { rh[i] = hits[i] ? -ph[i] : ph[i]; r[i] = d[i] * rh[i]; }
    for (int i = 0; i < n; ++i) { p[i] *= p_stride; ph[i] *= p_stride; xb[i] = x[i] + p[i]; }
    double to_tr;
    {   // intersect_trust_region(p_h, r_h, Delta): positive root of |p_h + t r_h| = Delta
        double a = 0.0, b = 0.0, c = -Delta * Delta;
        for (int i = 0; i < n; ++i) { a += rh[i] * rh[i]; b += ph[i] * rh[i]; c += ph[i] * ph[i]; }
        const double dd = sqrt(b * b - a * c);
        const double q = -(b + copysign(dd, b));
        const double t1 = q / a, t2 = c / q;
        to_tr = fmax(t1, t2);
    }
    const double to_bound = trf_step_to_bound(xb, r, lo, hi, n, nullptr);
    double r_stride = fmin(to_bound, to_tr), rl, ru;
    if (r_stride > 0.0) { rl = (1.0 - theta) * p_stride / r_stride; ru = (r_stride == to_bound) ? theta * to_bound : to_tr; }
    else { rl = 0.0; ru = -1.0; }
    double r_value = INFINITY;
    if (rl <= ru) {
        double a, b, c;
        Q.line(rh, ph, a, b, c);
        trf_minq(a, b, rl, ru, c, r_stride, r_value);
        for (int i = 0; i < n; ++i) { rh[i] = rh[i] * r_stride + ph[i]; r[i] = rh[i] * d[i]; }
    }
    for (int i = 0; i < n; ++i) { p[i] *= theta; ph[i] *= theta; }
    const double p_value = Q.value(ph);
    double agh[TN], ag[TN];
    for (int i = 0; i < n; ++i) { agh[i] = -Q.gh[i]; ag[i] = d[i] * agh[i]; }
    const double ag_to_tr = Delta / vnorm(agh, n);
    const double ag_to_bound = trf_step_to_bound(x, ag, lo, hi, n, nullptr);
    double ag_stride = ag_to_bound < ag_to_tr ? theta * ag_to_bound : ag_to_tr, ag_value;
    {
        double a, b, c;
        Q.line(agh, nullptr, a, b, c);
        trf_minq(a, b, 0.0, ag_stride, 0.0, ag_stride, ag_value);
    }
    for (int i = 0; i < n; ++i) { agh[i] *= ag_stride; ag[i] *= ag_stride; }
    const double* sp; const double* sh; double val;
    if (p_value < r_value && p_value < ag_value) { sp = p; sh = ph; val = p_value; }
    else if (r_value < p_value && r_value < ag_value) { sp = r; sh = rh; val = r_value; }
    else { sp = ag; sh = agh; val = ag_value; }
    for (int i = 0; i < n; ++i) { step[i] = sp[i]; step_h[i] = sh[i]; }
    return -val;
}

// Returns SciPy's termination status (1 gtol, 2 ftol, 3 xtol, 4 both, 0 = evaluation budget spent: res.success is
// status > 0); x holds the solution, cost = 0.5 |f|^2, nfev the residual evaluations used (Jacobians not counted).
__device__ int trf_solve(const IkProblem& P, double* x, double& cost, int& nfev) {
    const int n = P.n_unknowns, m = 9 * P.n_sets;
    const double ftol = 1e-6, xtol = 1e-6, gtol = 1e-8;
    const double* lo = P.lo; const double* hi = P.hi;
    for (int i = 0; i < n; ++i) {                         // least_squares: make_strictly_feasible(x0, lb, ub) with rstep = 1e-10
        const double lt = 1e-10 * fmax(1.0, fabs(lo[i])), ut = 1e-10 * fmax(1.0, fabs(hi[i]));
        if (x[i] - lo[i] <= fmin(hi[i] - x[i], lt)) x[i] = lo[i] + lt;
        else if (hi[i] - x[i] <= fmin(x[i] - lo[i], ut)) x[i] = hi[i] - ut;
        if (x[i] < lo[i] || x[i] > hi[i]) x[i] = 0.5 * (lo[i] + hi[i]);
    }
    double f[TM], J[TM][TN], g[TN], v[TN], dv[TN];
    ik_eval(P, x, f, nullptr);
    nfev = 1;
    trf_jac(P, x, f, n, m, J);
    cost = 0.0;
    for (int r = 0; r < m; ++r) cost += f[r] * f[r];
    cost *= 0.5;
    auto grad = [&]() { for (int u = 0; u < n; ++u) { double a = 0.0; for (int r = 0; r < m; ++r) a += J[r][u] * f[r]; g[u] = a; } };
    auto cl_scaling = [&]() {                             // Coleman-Li scaling vector and its derivative
        for (int i = 0; i < n; ++i) {
            v[i] = 1.0; dv[i] = 0.0;
            if (g[i] < 0.0) { v[i] = hi[i] - x[i]; dv[i] = -1.0; }
            else if (g[i] > 0.0) { v[i] = x[i] - lo[i]; dv[i] = 1.0; }
        }
    };
    grad();
    cl_scaling();
    double Delta = 0.0;
    for (int i = 0; i < n; ++i) Delta += x[i] * x[i] / v[i];
    Delta = sqrt(Delta);
    if (Delta == 0.0) Delta = 1.0;
    double alpha = 0.0;
    int status = -1;
    const int max_nfev = 100 * n;
    while (true) {
        cl_scaling();
        double g_norm = 0.0;
        for (int i = 0; i < n; ++i) g_norm = fmax(g_norm, fabs(g[i] * v[i]));
        if (g_norm < gtol) status = 1;
        if (status >= 0 || nfev == max_nfev) break;
        double d[TN], diag[TN], gh[TN], Jh[TM][TN], Ja[TA][TN], V[TN][TN], sv[TN], suf[TN];
        for (int i = 0; i < n; ++i) { d[i] = sqrt(v[i]); diag[i] = g[i] * dv[i]; gh[i] = d[i] * g[i]; }
        for (int r = 0; r < m; ++r) for (int u = 0; u < n; ++u) { Jh[r][u] = J[r][u] * d[u]; Ja[r][u] = Jh[r][u]; }
        for (int r = 0; r < n; ++r) for (int u = 0; u < n; ++u) Ja[m + r][u] = r == u ? sqrt(diag[u]) : 0.0;
        trf_svd(Ja, m + n, n, sv, V);
        for (int u = 0; u < n; ++u) { double a = 0.0; for (int r = 0; r < m; ++r) a += Ja[r][u] * f[r]; suf[u] = a; }   // s * (U^T f_aug)
        const double theta = fmax(0.995, 1.0 - g_norm);
        TrfQuad Q{Jh, gh, diag, n, m};
        double actual = -1.0, x_new[TN], f_new[TM], cost_new = cost;
        while (actual <= 0.0 && nfev < max_nfev) {
            double ph[TN], p[TN], step[TN], step_h[TN];
            trf_subproblem(n, m, suf, sv, V, Delta, alpha, ph);
            for (int i = 0; i < n; ++i) p[i] = d[i] * ph[i];
            const double predicted = trf_select_step(x, Q, p, ph, d, Delta, lo, hi, theta, step, step_h);
            for (int i = 0; i < n; ++i) {                 // make_strictly_feasible(x + step, rstep = 0)
                double xn = x[i] + step[i];
                if (xn <= lo[i]) xn = nextafter(lo[i], hi[i]);
                else if (xn >= hi[i]) xn = nextafter(hi[i], lo[i]);
                if (xn < lo[i] || xn > hi[i]) xn = 0.5 * (lo[i] + hi[i]);
                x_new[i] = xn;
            }
            ik_eval(P, x_new, f_new, nullptr);
            ++nfev;
            const double sh_norm = vnorm(step_h, n);
            cost_new = 0.0;
            for (int r = 0; r < m; ++r) cost_new += f_new[r] * f_new[r];
            cost_new *= 0.5;
            if (!isfinite(cost_new)) { Delta = 0.25 * sh_norm; continue; }
            actual = cost - cost_new;
            double ratio;                                 // update_tr_radius
            if (predicted > 0.0) ratio = actual / predicted;
            else if (predicted == 0.0 && actual == 0.0) ratio = 1.0;
            else ratio = 0.0;
            double Delta_new = Delta;
            if (ratio < 0.25) Delta_new = 0.25 * sh_norm;
            else if (ratio > 0.75 && sh_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
            const double s_norm = vnorm(step, n);
            const bool ft = actual < ftol * cost && ratio > 0.25;          // check_termination
            const bool xt = s_norm < xtol * (xtol + vnorm(x, n));
            if (ft && xt) status = 4; else if (ft) status = 2; else if (xt) status = 3;
            if (status >= 0) break;
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
        }
        if (actual > 0.0) {
            for (int i = 0; i < n; ++i) x[i] = x_new[i];
            for (int r = 0; r < m; ++r) f[r] = f_new[r];
            cost = cost_new;
            trf_jac(P, x, f, n, m, J);
            grad();
        }
    }
    return status < 0 ? 0 : status;
}

// _solve_full_chain_ik (Mamri.py:1410-1447): the solver is run from every initial guess -- the current joint angles
// when the caller has them, then zeros (:1425; on a fresh scene both are zeros and one run suffices) -- and the
// lowest cost among the runs SciPy would call successful (status > 0) is kept.
__device__ void solve_ik(const IkProblem& P, const double* initial, mamri_pose* out) {
    const int n = P.n_unknowns, m = 9 * P.n_sets;
    double best_x[MAMRI_MAX_CHAIN], best_cost = INFINITY;
    int best_status = -1, evals = 0;
    for (int guess = 0; guess < 2; ++guess) {
        double x[MAMRI_MAX_CHAIN];
        bool zero = true;
        for (int u = 0; u < n; ++u) { x[u] = (guess == 0 && initial) ? initial[u] : 0.0; zero = zero && x[u] == 0.0; }
        if (guess == 0 && (!initial || zero)) continue;            // the first guess would repeat the second
        double cost;
        int nfev;
        const int st = trf_solve(P, x, cost, nfev);
        evals += nfev;
        if (best_status < 0 || (st > 0 && (best_status == 0 || cost < best_cost))) {
            best_status = st; best_cost = cost;
            for (int u = 0; u < n; ++u) best_x[u] = x[u];
        }
    }
    double r[18];
    ik_eval(P, best_x, r, nullptr);
    for (int u = 0; u < MAMRI_MAX_CHAIN; ++u) out->joint_angles[u] = u < n ? best_x[u] : 0.0;
    double c2 = 0.0;
    for (int i = 0; i < m; ++i) c2 += r[i] * r[i];
    out->ik_cost = 0.5 * c2;
    double ss = 0.0;
    for (int i = 0; i < 9; ++i) ss += r[i] * r[i];                // last_ik_error: effector residuals only (:1443-1444)
    out->ik_rms_error = sqrt(ss / 9.0);
    out->ik_iterations = evals;
    out->ik_termination = best_status;
    out->ik_status = best_status > 0 ? MAMRI_IK_CONVERGED : MAMRI_IK_MAX_ITER;
}

}  // namespace

// Points of scan s: either packed [n_scans][max_points][3] with counts[] (host tables), or -- counts == NULL -- the
// device-written marker tables [n_scans][max_points][8] (rows {label, count, volume, RAS x y z, n_labels, body};
// a row is in use while its label is non-zero), so the stage can be queued right behind the scans.
__global__ void __launch_bounds__(32) k_pose(const mamri_robot* __restrict__ robot, const double* __restrict__ points,
                                             const int32_t* __restrict__ counts, int max_points, const double* __restrict__ initial,
                                             const double* __restrict__ saved_base, int prefer_saved,
                                             mamri_pose* __restrict__ poses) {
    __shared__ mamri_robot rb;
    __shared__ double pt[MAMRI_POSE_MAX_POINTS][3];
    __shared__ int s_match[MAMRI_MAX_LINKS][3];
    const int scan = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < int(sizeof(mamri_robot) / 4); i += 32) reinterpret_cast<uint32_t*>(&rb)[i] = reinterpret_cast<const uint32_t*>(robot)[i];
    const bool tables = counts == nullptr;
    int n_in;
    if (tables) {
        const double* t = points + size_t(scan) * max_points * 8;
        int c = 0;
        for (int i = lane; i < max_points; i += 32) c += t[size_t(i) * 8] != 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        n_in = c;
    } else {
        n_in = counts[scan];
    }
    int n = n_in;
    mamri_pose* out = poses + scan;
    const bool too_many = n > MAMRI_POSE_MAX_POINTS || n > max_points;
    if (too_many) n = 0;
    if (tables) {
        const double* t = points + size_t(scan) * max_points * 8;
        for (int i = lane; i < n * 3; i += 32) pt[i / 3][i % 3] = t[size_t(i / 3) * 8 + 3 + i % 3];
    } else {
        for (int i = lane; i < n * 3; i += 32) pt[i / 3][i % 3] = points[(size_t(scan) * max_points) * 3 + i];
    }
    for (int i = lane; i < MAMRI_MAX_LINKS * 3; i += 32) s_match[i / 3][i % 3] = -1;
    __syncwarp();

    // ---- joint_detection (:1343-1363)
    unsigned long long used = 0ull;
    for (int l = 0; l < rb.n_links; ++l) {
        const mamri_link& L = rb.links[l];
        if (!L.has_markers) continue;
        int avail[MAMRI_POSE_MAX_POINTS];
        int na = 0;
        for (int i = 0; i < n; ++i) if (!((used >> i) & 1ull)) avail[na++] = i;
        if (n < 3 || na < 3) continue;
        const double l1 = L.arm_lengths[0], l2 = L.arm_lengths[1];
        double e0 = l1, e1 = l2, e2 = hypot(l1, l2);
        sort3(e0, e1, e2);
        const double tol = rb.distance_tolerance;
        unsigned best = 0xFFFFFFFFu;                              // rank (i*na + j)*na + k of the first matching combination
        for (int pr = lane; pr < na * na; pr += 32) {
            const int i = pr / na, j = pr - i * na;
            if (j <= i) continue;
            const double dij = dist3(pt[avail[i]], pt[avail[j]]);
            for (int k = j + 1; k < na; ++k) {
                double d0 = dij, d1 = dist3(pt[avail[i]], pt[avail[k]]), d2 = dist3(pt[avail[j]], pt[avail[k]]);
                sort3(d0, d1, d2);
                if (fabs(d0 - e0) <= tol && fabs(d1 - e1) <= tol && fabs(d2 - e2) <= tol) {
                    best = min(best, unsigned((i * na + j) * na + k));
                    break;                                        // later k only have higher ranks
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(FULL, best, o));
        if (best == 0xFFFFFFFFu) continue;
        const int k = int(best % unsigned(na)), j = int((best / unsigned(na)) % unsigned(na)), i = int(best / unsigned(na * na));
        int ids[3] = {avail[i], avail[j], avail[k]};
        used |= (1ull << ids[0]) | (1ull << ids[1]) | (1ull << ids[2]);
        // _sort_l_shaped_markers (:1782-1792): corner, short arm, long arm; unsorted combination if no corner qualifies
        const double ls = fmin(l1, l2), ll = fmax(l1, l2);
        int srt[3] = {ids[0], ids[1], ids[2]};
        for (int c = 0; c < 3; ++c) {
            const int p1 = (c + 1) % 3, p2 = (c + 2) % 3;
            const double d1 = dist3(pt[ids[c]], pt[ids[p1]]), d2 = dist3(pt[ids[c]], pt[ids[p2]]);
            if (fabs(d1 - ls) <= tol && fabs(d2 - ll) <= tol) { srt[0] = ids[c]; srt[1] = ids[p1]; srt[2] = ids[p2]; break; }
            if (fabs(d1 - ll) <= tol && fabs(d2 - ls) <= tol) { srt[0] = ids[c]; srt[1] = ids[p2]; srt[2] = ids[p1]; break; }
        }
        if (lane == 0) for (int q = 0; q < 3; ++q) s_match[l][q] = srt[q];
        __syncwarp();
    }
    __syncwarp();
    if (lane != 0) return;

    // ---- results, registration, IK (lane 0)
    out->n_points = n_in;
    out->status = too_many ? MAMRI_ERR_CAPACITY : MAMRI_OK;
    for (int l = 0; l < MAMRI_MAX_LINKS; ++l) for (int q = 0; q < 3; ++q) out->matched[l][q] = s_match[l][q];
    out->has_base = 0;
    out->ik_status = MAMRI_IK_NOT_RUN;
    out->ik_iterations = 0;
    out->ik_termination = 0;
    out->ik_cost = 0.0;
    out->ik_rms_error = 0.0;
    for (int i = 0; i < 16; ++i) out->base_matrix[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int u = 0; u < MAMRI_MAX_CHAIN; ++u) out->joint_angles[u] = 0.0;
    // _get_baseplate_transform (:1376-1408): the saved transform when the caller prefers it, else the registration of the
    // scan's baseplate markers, else the saved transform as a fall-back, else no pose
    const int bl = rb.base_link;
    const bool in_scan = bl >= 0 && bl < rb.n_links && s_match[bl][0] >= 0;
    IkProblem P;
    P.rb = &rb;
    if (saved_base && (prefer_saved || !in_scan)) {
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) P.base.r[3 * i + j] = saved_base[4 * i + j]; P.base.t[i] = saved_base[4 * i + 3]; }
        out->has_base = 2;
    } else if (in_scan) {
        double tgt[9];
        for (int q = 0; q < 3; ++q) for (int k = 0; k < 3; ++k) tgt[3 * q + k] = pt[s_match[bl][q]][k];
        const double avg_y = (tgt[1] + tgt[4] + tgt[7]) / 3.0;          // (:1371-1373)
        tgt[1] = tgt[4] = tgt[7] = avg_y;
        P.base = landmark_rigid(rb.links[bl].marker_coords, tgt);
        out->has_base = 1;
    } else {
        return;                                                         // neither: "Pose estimation failed" (:1406-1408)
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out->base_matrix[4 * i + j] = P.base.r[3 * i + j];
        out->base_matrix[4 * i + 3] = P.base.t[i];
    }
    const int ef = rb.effector_link;
    if (ef < 0 || ef >= rb.n_links || s_match[ef][0] < 0) return;   // "Joint6Fiducials" missing: no IK (:866-868)
    P.n_sets = 1;
    P.link[0] = ef; P.weight[0] = 1.0;
    for (int q = 0; q < 3; ++q) {
        const double* lc = &rb.links[ef].marker_coords[3 * q];
        const double sgn = rb.apply_correction ? -1.0 : 1.0;        // RotateZ(180) of the effector markers (:1511-1514)
        P.local[0][3 * q] = sgn * lc[0]; P.local[0][3 * q + 1] = sgn * lc[1]; P.local[0][3 * q + 2] = lc[2];
        for (int k = 0; k < 3; ++k) P.target[0][3 * q + k] = pt[s_match[ef][q]][k];
    }
    const int se = rb.secondary_link;
    if (se >= 0 && se < rb.n_links && s_match[se][0] >= 0) {        // "Joint4Fiducials" as a weighted secondary objective (:1419-1424)
        P.n_sets = 2;
        P.link[1] = se; P.weight[1] = rb.secondary_weight;
        for (int q = 0; q < 3; ++q)
            for (int k = 0; k < 3; ++k) { P.local[1][3 * q + k] = rb.links[se].marker_coords[3 * q + k]; P.target[1][3 * q + k] = pt[s_match[se][q]][k]; }
    }
    P.n_unknowns = 0;
    for (int l = 0; l < rb.n_links; ++l) {
        const int ci = rb.links[l].chain_index;
        if (ci >= 0 && ci < MAMRI_MAX_CHAIN) {
            P.lo[ci] = rb.links[l].limits_deg[0] * (M_PI / 180.0);
            P.hi[ci] = rb.links[l].limits_deg[1] * (M_PI / 180.0);
            if (ci + 1 > P.n_unknowns) P.n_unknowns = ci + 1;
        }
    }
    solve_ik(P, initial ? initial + size_t(scan) * MAMRI_MAX_CHAIN : nullptr, out);
}

// ------------------------------------------------------------------------------------------------
// robot-vs-body collision sampling: stands in for MamriLogic._check_collision
// (Mamri/Mamri.py:1555-1575), which runs vtkCollisionDetectionFilter between every link's collision mesh and
// the body mesh for one joint configuration at a time (called per configuration of a path, :976-982, and inside
// the trajectory IK's error function, :1541).  Here: a batch of configurations, one CTA each; every sample point
// of every link (link-local coordinates, e.g. the vertices of the *_collision.STL meshes) goes through the
// link's forward-kinematics transform and the RAS -> voxel affine, and hits if it lands on a non-zero voxel of
// the body labelmap.  A different geometric predicate than VTK's triangle test (points inside the volume, not
// surface intersection): decisions agree on clear / colliding poses, not on grazing contacts.
// ------------------------------------------------------------------------------------------------
struct CollisionArgs {
    double base[16];
    double m[12];              // RAS mm -> voxel index, row-major 3x4
    int nx, ny, nz;
    int n_links;
    int offsets[MAMRI_MAX_LINKS + 1];
};

__global__ void __launch_bounds__(256) k_collision(const mamri_robot* __restrict__ robot, CollisionArgs a,
                                                   const double* __restrict__ angles, const float* __restrict__ pts,
                                                   const uint8_t* __restrict__ mask, mamri_collision_result* __restrict__ out) {
    __shared__ M34 world[MAMRI_MAX_LINKS];
    __shared__ unsigned s_mask, s_count;
    __shared__ int s_first;
    const int cfg = blockIdx.x;
    if (threadIdx.x == 0) {
        M34 base;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) base.r[3 * i + j] = a.base[4 * i + j]; base.t[i] = a.base[4 * i + 3]; }
        for (int l = 0; l < robot->n_links; ++l) {
            const mamri_link& L = robot->links[l];
            const double ang = (L.chain_index >= 0 && L.chain_index < MAMRI_MAX_CHAIN) ? angles[size_t(cfg) * MAMRI_MAX_CHAIN + L.chain_index] : 0.0;
            world[l] = mul(L.parent >= 0 ? world[L.parent] : base, local_tf(L, ang));
        }
        s_mask = 0u; s_count = 0u; s_first = MAMRI_MAX_LINKS;
    }
    __syncthreads();
    unsigned hits = 0, lmask = 0;
    int first = MAMRI_MAX_LINKS;
    for (int l = 0; l < a.n_links; ++l) {
        const M34& T = world[l];
        for (int i = a.offsets[l] + int(threadIdx.x); i < a.offsets[l + 1]; i += blockDim.x) {
            const double lx = pts[3 * size_t(i)], ly = pts[3 * size_t(i) + 1], lz = pts[3 * size_t(i) + 2];
            double w[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                w[k] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[3 * k], lx), __dmul_rn(T.r[3 * k + 1], ly)), __dmul_rn(T.r[3 * k + 2], lz)), T.t[k]);
            long long idx[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                idx[k] = llrint(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.m[4 * k], w[0]), __dmul_rn(a.m[4 * k + 1], w[1])), __dmul_rn(a.m[4 * k + 2], w[2])), a.m[4 * k + 3]));
            if (idx[0] >= 0 && idx[0] < a.nx && idx[1] >= 0 && idx[1] < a.ny && idx[2] >= 0 && idx[2] < a.nz &&
                mask[(size_t(idx[2]) * a.ny + idx[1]) * a.nx + idx[0]] != 0) {
                ++hits;
                lmask |= 1u << l;
                if (l < first) first = l;
            }
        }
    }
    if (hits) { atomicAdd(&s_count, hits); atomicOr(&s_mask, lmask); atomicMin(&s_first, first); }
    __syncthreads();
    if (threadIdx.x == 0) {
        out[cfg].link_mask = s_mask;
        out[cfg].n_points_inside = s_count;
        out[cfg].first_link = s_mask ? s_first : -1;
        out[cfg].reserved = 0;
    }
}

cudaError_t launch_collision(const mamri_robot* d_robot, const double base[16], const double m[12], int nx, int ny, int nz,
                             const int* offsets, int n_links, const double* d_angles, int n_configs, const float* d_points,
                             const uint8_t* d_mask, mamri_collision_result* d_out, cudaStream_t s) {
    if (n_configs <= 0) return cudaSuccess;
    CollisionArgs a;
    for (int i = 0; i < 16; ++i) a.base[i] = base[i];
    for (int i = 0; i < 12; ++i) a.m[i] = m[i];
    a.nx = nx; a.ny = ny; a.nz = nz; a.n_links = n_links;
    for (int i = 0; i <= MAMRI_MAX_LINKS; ++i) a.offsets[i] = i <= n_links ? offsets[i] : offsets[n_links];
    k_collision<<<n_configs, 256, 0, s>>>(d_robot, a, d_angles, d_points, d_mask, d_out);
    return cudaGetLastError();
}

cudaError_t launch_pose(const mamri_robot* d_robot, const double* d_points, const int32_t* d_counts, int n_scans,
                        int max_points, const double* d_initial, const double* d_saved_base, int prefer_saved,
                        mamri_pose* d_poses, cudaStream_t s) {
    if (n_scans <= 0) return cudaSuccess;
    k_pose<<<n_scans, 32, 0, s>>>(d_robot, d_points, d_counts, max_points, d_initial, d_saved_base, prefer_saved, d_poses);
    return cudaGetLastError();
}
