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
// on lane 0: 4x4 Jacobi eigen-solve, then a projected Levenberg-Marquardt with the analytic Jacobian of the
// chain.  The reference's solver is SciPy's TRF with a finite-difference Jacobian stopped at ftol = xtol =
// 1e-6; this one iterates to the minimum itself, so the two agree to the reference's own stopping error
// (tests: 1e-5 rad), not bit for bit -- SciPy stays the parity path for the 1e-6 rad criterion.
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

// Solves (A + lambda * diag(A)) dx = -g for symmetric positive semi-definite A (n <= MAMRI_MAX_CHAIN) by Cholesky.
__device__ bool solve_damped(const double (*A)[MAMRI_MAX_CHAIN], const double* g, double lambda, int n, double* dx) {
    double L[MAMRI_MAX_CHAIN][MAMRI_MAX_CHAIN];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double sum = A[i][j] + (i == j ? lambda * (A[i][i] > 1e-12 ? A[i][i] : 1e-12) : 0.0);
            for (int k = 0; k < j; ++k) sum -= L[i][k] * L[j][k];
            if (i == j) { if (!(sum > 0.0)) return false; L[i][i] = sqrt(sum); }
            else L[i][j] = sum / L[j][j];
        }
    double y[MAMRI_MAX_CHAIN];
    for (int i = 0; i < n; ++i) { double sum = -g[i]; for (int k = 0; k < i; ++k) sum -= L[i][k] * y[k]; y[i] = sum / L[i][i]; }
    for (int i = n - 1; i >= 0; --i) { double sum = y[i]; for (int k = i + 1; k < n; ++k) sum -= L[k][i] * dx[k]; dx[i] = sum / L[i][i]; }
    return true;
}

__device__ void solve_ik(const IkProblem& P, mamri_pose* out) {
    const int n = P.n_unknowns, m = 9 * P.n_sets;
    double x[MAMRI_MAX_CHAIN], r[18], J[18][MAMRI_MAX_CHAIN];
    for (int u = 0; u < n; ++u) x[u] = fmin(fmax(0.0, P.lo[u]), P.hi[u]);   // both initial guesses of :1425 are zeros on a fresh scene
    ik_eval(P, x, r, J);
    double cost = 0.0;
    for (int i = 0; i < m; ++i) cost += r[i] * r[i];
    cost *= 0.5;
    double lambda = 1e-3;
    int it = 0, status = MAMRI_IK_MAX_ITER;
    for (; it < 500; ++it) {
        double A[MAMRI_MAX_CHAIN][MAMRI_MAX_CHAIN], g[MAMRI_MAX_CHAIN];
        for (int u = 0; u < n; ++u) {
            g[u] = 0.0;
            for (int i = 0; i < m; ++i) g[u] += J[i][u] * r[i];
            for (int v = 0; v <= u; ++v) { double sum = 0.0; for (int i = 0; i < m; ++i) sum += J[i][u] * J[i][v]; A[u][v] = A[v][u] = sum; }
        }
        // active set: unknowns sitting on a joint limit with the gradient pushing outwards stay fixed for this step
        int idx[MAMRI_MAX_CHAIN], nf = 0;
        double gmax = 0.0;
        for (int u = 0; u < n; ++u) {
            const bool blocked = (x[u] <= P.lo[u] && g[u] > 0.0) || (x[u] >= P.hi[u] && g[u] < 0.0);
            if (!blocked) { idx[nf++] = u; gmax = fmax(gmax, fabs(g[u])); }
        }
        if (nf == 0 || gmax < 1e-11) { status = MAMRI_IK_CONVERGED; break; }
        double Af[MAMRI_MAX_CHAIN][MAMRI_MAX_CHAIN], gf[MAMRI_MAX_CHAIN];
        for (int a = 0; a < nf; ++a) { gf[a] = g[idx[a]]; for (int b = 0; b < nf; ++b) Af[a][b] = A[idx[a]][idx[b]]; }
        bool accepted = false;
        for (int tries = 0; tries < 40 && !accepted; ++tries) {
            double df[MAMRI_MAX_CHAIN], xt[MAMRI_MAX_CHAIN], rt[18];
            if (!solve_damped(Af, gf, lambda, nf, df)) { lambda *= 10.0; continue; }
            double step = 0.0;
            for (int u = 0; u < n; ++u) xt[u] = x[u];
            for (int a = 0; a < nf; ++a) {
                const int u = idx[a];
                xt[u] = fmin(fmax(x[u] + df[a], P.lo[u]), P.hi[u]);
                step = fmax(step, fabs(xt[u] - x[u]));
            }
            ik_eval(P, xt, rt, nullptr);
            double ct = 0.0;
            for (int i = 0; i < m; ++i) ct += rt[i] * rt[i];
            ct *= 0.5;
            if (ct <= cost) {
                for (int u = 0; u < n; ++u) x[u] = xt[u];
                const bool tiny = step < 1e-14 || (cost - ct) <= 1e-16 * cost;
                cost = ct;
                lambda = fmax(lambda / 5.0, 1e-15);
                accepted = true;
                if (tiny) status = MAMRI_IK_CONVERGED;
            } else {
                lambda = fmin(lambda * 4.0, 1e12);
            }
        }
        if (!accepted) { status = MAMRI_IK_CONVERGED; break; }    // no downhill step left at any damping: at a minimum to rounding
        if (status == MAMRI_IK_CONVERGED) { ++it; break; }
        ik_eval(P, x, r, J);
    }
    ik_eval(P, x, r, nullptr);
    for (int u = 0; u < MAMRI_MAX_CHAIN; ++u) out->joint_angles[u] = u < n ? x[u] : 0.0;
    out->ik_cost = cost;
    double ss = 0.0;
    for (int i = 0; i < 9; ++i) ss += r[i] * r[i];                // last_ik_error: effector residuals only (:1443-1444)
    out->ik_rms_error = sqrt(ss / 9.0);
    out->ik_iterations = it;
    out->ik_status = status;
}

}  // namespace

// Points of scan s: either packed [n_scans][max_points][3] with counts[] (host tables), or -- counts == NULL -- the
// device-written marker tables [n_scans][max_points][8] (rows {label, count, volume, RAS x y z, n_labels, body};
// a row is in use while its label is non-zero), so the stage can be queued right behind the scans.
__global__ void __launch_bounds__(32) k_pose(const mamri_robot* __restrict__ robot, const double* __restrict__ points,
                                             const int32_t* __restrict__ counts, int max_points, mamri_pose* __restrict__ poses) {
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
    out->ik_cost = 0.0;
    out->ik_rms_error = 0.0;
    for (int i = 0; i < 16; ++i) out->base_matrix[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int u = 0; u < MAMRI_MAX_CHAIN; ++u) out->joint_angles[u] = 0.0;
    const int bl = rb.base_link;
    if (bl < 0 || bl >= rb.n_links || s_match[bl][0] < 0) return;   // no baseplate in the scan (:1392-1398 falls back to a saved transform)
    double tgt[9];
    for (int q = 0; q < 3; ++q) for (int k = 0; k < 3; ++k) tgt[3 * q + k] = pt[s_match[bl][q]][k];
    const double avg_y = (tgt[1] + tgt[4] + tgt[7]) / 3.0;          // (:1371-1373)
    tgt[1] = tgt[4] = tgt[7] = avg_y;
    IkProblem P;
    P.rb = &rb;
    P.base = landmark_rigid(rb.links[bl].marker_coords, tgt);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out->base_matrix[4 * i + j] = P.base.r[3 * i + j];
        out->base_matrix[4 * i + 3] = P.base.t[i];
    }
    out->has_base = 1;
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
    solve_ik(P, out);
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
                        int max_points, mamri_pose* d_poses, cudaStream_t s) {
    if (n_scans <= 0) return cudaSuccess;
    k_pose<<<n_scans, 32, 0, s>>>(d_robot, d_points, d_counts, max_points, d_poses);
    return cudaGetLastError();
}
