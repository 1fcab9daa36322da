/*
 * mamri_oracle.c -- plain-C restatement of the reference's fiducial-detection path, for volumes too
 * large for the NumPy oracle and as the timed CPU arm of bench.py.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into or called by the product.
 * PARITY UNPINNED: the reference does this work inside SimpleITK (Mamri/Mamri.py:1308-1310), which is
 * not available here; this file restates ITK 5.x semantics and is cross-checked bit for
 * bit against oracle/segmentation.py and oracle/bruteforce.py by tests/test_oracle_c.py.
 *
 * Layout: x fastest, linear index = x + nx*(y + ny*z).  Every stage is threaded with OpenMP: threshold and morphology
 * over rows, the labelling over slabs of slices that are merged along their faces (as ITK threads it), the statistics
 * over rows with atomic adds of per-run sums.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- sitk.BinaryThreshold(img, lo, hi)                     Mamri.py:1308 ------------------------------ */
/* bounds are static_cast to the pixel type (truncation; clamped when out of range), both ends inclusive */
static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

#define THRESH_INT(NAME, T, TMIN, TMAX)                                                               \
    static void NAME(const T* v, size_t n, double lo, double hi, uint8_t* out) {                      \
        T l = (T)clampd(trunc(lo), TMIN, TMAX), h = (T)clampd(trunc(hi), TMIN, TMAX);                  \
        _Pragma("omp parallel for schedule(static)")                                                  \
        for (size_t i = 0; i < n; ++i) out[i] = (v[i] >= l && v[i] <= h) ? 1 : 0;                      \
    }
THRESH_INT(thresh_u8, uint8_t, 0.0, 255.0)
THRESH_INT(thresh_i16, int16_t, -32768.0, 32767.0)
THRESH_INT(thresh_u16, uint16_t, 0.0, 65535.0)
THRESH_INT(thresh_i32, int32_t, -2147483648.0, 2147483647.0)
static void thresh_f32(const float* v, size_t n, double lo, double hi, uint8_t* out) {
    float l = (float)lo, h = (float)hi;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = (v[i] >= l && v[i] <= h) ? 1 : 0;   /* NaN -> 0 */
}

static void thresh_f64(const double* v, size_t n, double lo, double hi, uint8_t* out) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) out[i] = (v[i] >= lo && v[i] <= hi) ? 1 : 0;   /* NaN -> 0 */
}

API int oracle_threshold(const void* vol, int dtype, size_t n, double lo, double hi, uint8_t* out) {
    switch (dtype) {
        case 0: thresh_u8((const uint8_t*)vol, n, lo, hi, out); return 0;
        case 1: thresh_i16((const int16_t*)vol, n, lo, hi, out); return 0;
        case 2: thresh_u16((const uint16_t*)vol, n, lo, hi, out); return 0;
        case 3: thresh_i32((const int32_t*)vol, n, lo, hi, out); return 0;
        case 4: thresh_f32((const float*)vol, n, lo, hi, out); return 0;
        case 5: thresh_f64((const double*)vol, n, lo, hi, out); return 0;
        default: return -1;
    }
}

/* ---- sitk.BinaryMorphologicalClosing(binary, [r]*3, sitkBall)   Mamri.py:1308 -------------------------- */
/* ITK ball (FlatStructuringElement::Ball, radiusIsParametric=false): offsets with d.d <= r*r + r.
 * SafeBorder closing = closing of the zero-extended mask: work on a copy padded by P = 2r (+r more in x so
 * that shifted row reads stay inside), dilate with OR over the ball, erode with AND over the ball, crop. */
static void morph_pass(const uint8_t* src, uint8_t* dst, int px, int py, int pz, int r, int m, int erode) {
    /* src/dst: padded volumes px*py*pz; results are computed for voxels at least m from every face
     * (m >= r) and left untouched elsewhere */
    const int r2 = r * r + r;
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = m; z < pz - m; ++z)
        for (int y = m; y < py - m; ++y) {
            uint8_t* o = dst + ((size_t)z * py + y) * px;
            const int n = px - 2 * m;
            memset(o + m, erode ? 1 : 0, (size_t)n);
            for (int dz = -r; dz <= r; ++dz)
                for (int dy = -r; dy <= r; ++dy)
                    for (int dx = -r; dx <= r; ++dx) {
                        if (dx * dx + dy * dy + dz * dz > r2) continue;
                        const uint8_t* s = src + ((size_t)(z + dz) * py + (y + dy)) * px + dx;
                        if (erode)
                            for (int x = m; x < m + n; ++x) o[x] &= s[x];
                        else
                            for (int x = m; x < m + n; ++x) o[x] |= s[x];
                    }
        }
}

API int oracle_closing(const uint8_t* in, int nx, int ny, int nz, int r, uint8_t* out) {
    const size_t n = (size_t)nx * ny * nz;
    if (r <= 0) { memcpy(out, in, n); return 0; }
    const int P = 2 * r;
    const int px = nx + 2 * P, py = ny + 2 * P, pz = nz + 2 * P;
    const size_t pn = (size_t)px * py * pz;
    uint8_t* a = (uint8_t*)calloc(pn, 1);
    uint8_t* b = (uint8_t*)calloc(pn, 1);
    if (!a || !b) { free(a); free(b); return -2; }
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            memcpy(a + ((size_t)(z + P) * py + (y + P)) * px + P, in + ((size_t)z * ny + y) * nx, (size_t)nx);
    morph_pass(a, b, px, py, pz, r, r, 0);          /* dilation on the image grown by r (b is 0 further out) */
    morph_pass(b, a, px, py, pz, r, P, 1);          /* erosion on the image domain; samples only that r-apron */
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            memcpy(out + ((size_t)z * ny + y) * nx, a + ((size_t)(z + P) * py + (y + P)) * px + P, (size_t)nx);
    free(a);
    free(b);
    return 0;
}

/* ---- sitk.BinaryMorphologicalOpening(binary, [r]*3, sitkBall): north_star extension, not in the reference ----- */
/* itk::BinaryMorphologicalOpeningImageFilter: BinaryErode with BoundaryToForeground = true (outside the image counts
 * as foreground), then BinaryDilate (outside = background); no safe-border padding.  On a copy padded by r: the
 * apron is 1 for the erosion and 0 for the dilation. */
API int oracle_opening(const uint8_t* in, int nx, int ny, int nz, int r, uint8_t* out) {
    const size_t n = (size_t)nx * ny * nz;
    if (r <= 0) { memcpy(out, in, n); return 0; }
    const int P = r;
    const int px = nx + 2 * P, py = ny + 2 * P, pz = nz + 2 * P;
    const size_t pn = (size_t)px * py * pz;
    uint8_t* a = (uint8_t*)malloc(pn);
    uint8_t* b = (uint8_t*)calloc(pn, 1);
    if (!a || !b) { free(a); free(b); return -2; }
    memset(a, 1, pn);                               /* outside = foreground for the erosion */
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            memcpy(a + ((size_t)(z + P) * py + (y + P)) * px + P, in + ((size_t)z * ny + y) * nx, (size_t)nx);
    morph_pass(a, b, px, py, pz, r, P, 1);          /* erosion on the image domain; b stays 0 in the apron */
    memset(a, 0, pn);
    morph_pass(b, a, px, py, pz, r, P, 0);          /* dilation on the image domain, outside = background */
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            memcpy(out + ((size_t)z * ny + y) * nx, a + ((size_t)(z + P) * py + (y + P)) * px + P, (size_t)nx);
    free(a);
    free(b);
    return 0;
}

/* ---- sitk.ConnectedComponent(closed)                       Mamri.py:1309 ------------------------------ */
/* Raster scan; a voxel takes the smallest root among its already-visited neighbours (6: -x,-y,-z; 26: the 13
 * preceding neighbours), unions keep the smaller label; roots are then renumbered consecutively in increasing
 * order, i.e. by the component's first voxel in raster order (ITK's numbering). */
static uint32_t uf_find(uint32_t* p, uint32_t x) {
    uint32_t r = x;
    while (p[r] != r) r = p[r];
    while (p[x] != r) { uint32_t nx_ = p[x]; p[x] = r; x = nx_; }
    return r;
}

/* Offsets of the already-visited neighbours in raster order: 3 for face connectivity, 13 for 26-connectivity. */
static int ccl_offsets(int conn, int offs[13][3]) {
    int no = 0;
    if (conn == 6) {
        int t[3][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
        memcpy(offs, t, sizeof(t));
        no = 3;
    } else if (conn == 26) {
        for (int dz = -1; dz <= 0; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (dz == 0 && (dy > 0 || (dy == 0 && dx >= 0))) continue;
                    offs[no][0] = dx; offs[no][1] = dy; offs[no][2] = dz; ++no;
                }
    }
    return no;
}

/* Raster scan of the slices [z0, z1): provisional labels 1, 2, ... local to the slab (neighbours below z0 are not
 * looked at), union-find in *parent_io (grown as needed).  Returns the number of entries used (labels + 1), 0 on
 * allocation failure. */
static size_t ccl_scan_slab(const uint8_t* mask, int nx, int ny, int z0, int z1, int offs[13][3], int no, uint32_t* labels,
                            uint32_t** parent_io, size_t* cap_io) {
    uint32_t* parent = *parent_io;
    size_t cap = *cap_io, used = 1;
    parent[0] = 0;
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t i = ((size_t)z * ny + y) * nx + x;
                if (!mask[i]) { labels[i] = 0; continue; }
                uint32_t best = 0;
                for (int k = 0; k < no; ++k) {
                    const int X = x + offs[k][0], Y = y + offs[k][1], Z = z + offs[k][2];
                    if (X < 0 || X >= nx || Y < 0 || Y >= ny || Z < z0) continue;
                    const uint32_t l = labels[((size_t)Z * ny + Y) * nx + X];
                    if (!l) continue;
                    const uint32_t rt = uf_find(parent, l);
                    if (!best) best = rt;
                    else if (rt != best) {
                        if (rt < best) { parent[best] = rt; best = rt; } else parent[rt] = best;
                    }
                }
                if (!best) {
                    if (used == cap) {
                        cap *= 2;
                        uint32_t* np_ = (uint32_t*)realloc(parent, cap * sizeof(uint32_t));
                        if (!np_) { *parent_io = parent; *cap_io = cap / 2; return 0; }
                        parent = np_;
                    }
                    best = (uint32_t)used;
                    parent[used++] = best;
                }
                labels[i] = best;
            }
    *parent_io = parent;
    *cap_io = cap;
    return used;
}

/* One slab = the whole volume: the sequential scan (kept as the cross-check of the slab-parallel form below). */
API int oracle_ccl_serial(const uint8_t* mask, int nx, int ny, int nz, int conn, uint32_t* labels, uint32_t* n_labels) {
    const size_t n = (size_t)nx * ny * nz;
    int offs[13][3];
    const int no = ccl_offsets(conn, offs);
    if (!no) return -1;
    size_t cap = 1 << 16;
    uint32_t* parent = (uint32_t*)malloc(cap * sizeof(uint32_t));
    if (!parent) return -2;
    const size_t used = ccl_scan_slab(mask, nx, ny, 0, nz, offs, no, labels, &parent, &cap);
    if (!used) { free(parent); return -2; }
    uint32_t* final_ = (uint32_t*)calloc(used, sizeof(uint32_t));
    if (!final_) { free(parent); return -2; }
    uint32_t k = 0;
    for (size_t l = 1; l < used; ++l)
        if (parent[l] == l) final_[l] = ++k;          /* roots in increasing provisional order */
    for (size_t l = 1; l < used; ++l)
        if (parent[l] != l) final_[l] = final_[uf_find(parent, (uint32_t)l)];
    for (size_t i = 0; i < n; ++i)
        if (labels[i]) labels[i] = final_[labels[i]];
    *n_labels = k;
    free(final_);
    free(parent);
    return 0;
}

/* The same labelling threaded the way itk::ConnectedComponentImageFilter is (slabs of slices scanned independently, then
 * merged along the slab faces): provisional labels of slab t come after those of slab t-1, inside a slab in raster order of
 * creation, and unions keep the smaller label -- so a component's root is still the label created at its first voxel in
 * raster order and the consecutive renumbering gives ITK's numbering, identical to the sequential scan. */
#define CCL_MAX_SLABS 256
API int oracle_ccl(const uint8_t* mask, int nx, int ny, int nz, int conn, uint32_t* labels, uint32_t* n_labels) {
    const size_t n = (size_t)nx * ny * nz;
    int offs[13][3];
    const int no = ccl_offsets(conn, offs);
    if (!no) return -1;
    int T = oracle_num_threads();
    if (T > nz) T = nz;
    if (T > CCL_MAX_SLABS) T = CCL_MAX_SLABS;
    if (T <= 1 || n < ((size_t)1 << 16)) return oracle_ccl_serial(mask, nx, ny, nz, conn, labels, n_labels);
    uint32_t* par[CCL_MAX_SLABS];
    size_t cap[CCL_MAX_SLABS], used[CCL_MAX_SLABS], off[CCL_MAX_SLABS + 1];
    int zs[CCL_MAX_SLABS + 1];
    for (int t = 0; t <= T; ++t) zs[t] = (int)((long long)nz * t / T);
    int fail = 0;
    for (int t = 0; t < T; ++t) {
        cap[t] = 1 << 14;
        par[t] = (uint32_t*)malloc(cap[t] * sizeof(uint32_t));
        if (!par[t]) fail = 1;
    }
    if (!fail) {
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int t = 0; t < T; ++t) used[t] = ccl_scan_slab(mask, nx, ny, zs[t], zs[t + 1], offs, no, labels, &par[t], &cap[t]);
        for (int t = 0; t < T; ++t) if (!used[t]) fail = 1;
    }
    if (fail) { for (int t = 0; t < T; ++t) free(par[t]); return -2; }
    /* global provisional label of local label l of slab t: off[t] + l   (l >= 1; off[0] = 0) */
    off[0] = 0;
    for (int t = 0; t < T; ++t) off[t + 1] = off[t] + (used[t] - 1);
    const size_t total = off[T] + 1;
    if (total > 0xFFFFFFFFull) { for (int t = 0; t < T; ++t) free(par[t]); return -2; }
    uint32_t* parent = (uint32_t*)malloc(total * sizeof(uint32_t));
    uint32_t* final_ = (uint32_t*)calloc(total, sizeof(uint32_t));
    if (!parent || !final_) { free(parent); free(final_); for (int t = 0; t < T; ++t) free(par[t]); return -2; }
    parent[0] = 0;
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        for (size_t l = 1; l < used[t]; ++l) parent[off[t] + l] = (uint32_t)(off[t] + par[t][l]);
        free(par[t]);
    }
    /* merge along the slab faces: every voxel of a slab's first slice with its neighbours in the slice below */
    for (int t = 1; t < T; ++t) {
        const int z = zs[t];
        if (z >= nz || z == 0) continue;
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t i = ((size_t)z * ny + y) * nx + x;
                if (!labels[i]) continue;
                const uint32_t a0 = (uint32_t)(off[t] + labels[i]);
                for (int k = 0; k < no; ++k) {
                    if (offs[k][2] != -1) continue;
                    const int X = x + offs[k][0], Y = y + offs[k][1];
                    if (X < 0 || X >= nx || Y < 0 || Y >= ny) continue;
                    const uint32_t lb = labels[((size_t)(z - 1) * ny + Y) * nx + X];
                    if (!lb) continue;
                    uint32_t ra = uf_find(parent, a0), rb = uf_find(parent, (uint32_t)(off[t - 1] + lb));
                    if (ra == rb) continue;
                    if (ra < rb) parent[rb] = ra; else parent[ra] = rb;   /* the smaller label stays the root */
                }
            }
    }
    uint32_t k = 0;
    for (size_t l = 1; l < total; ++l)
        if (parent[l] == l) final_[l] = ++k;          /* roots in increasing provisional order */
    for (size_t l = 1; l < total; ++l)
        if (parent[l] != l) final_[l] = final_[uf_find(parent, (uint32_t)l)];
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        const size_t i0 = (size_t)zs[t] * ny * nx, i1 = (size_t)zs[t + 1] * ny * nx;
        for (size_t i = i0; i < i1; ++i)
            if (labels[i]) labels[i] = final_[off[t] + labels[i]];
    }
    *n_labels = k;
    free(final_);
    free(parent);
    return 0;
}

/* ---- LabelShapeStatisticsImageFilter: exact integer sums per label     Mamri.py:1309 -------------------- */
/* sums[l*10 + ..] = count, sx, sy, sz, sxx, syy, szz, sxy, sxz, syz for label l+1.  Threaded over rows; a thread sums
 * each x-run of one label privately and adds it to the table with atomics (integer sums: the order does not matter). */
API int oracle_label_sums(const uint32_t* labels, int nx, int ny, int nz, uint32_t n_labels, uint64_t* sums) {
    memset(sums, 0, (size_t)n_labels * 10 * sizeof(uint64_t));
    const long long rows = (long long)ny * nz;
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < rows; ++r) {
        const int z = (int)(r / ny), y = (int)(r - (long long)z * ny);
        const uint32_t* row = labels + (size_t)r * nx;
        const uint64_t Y = (uint64_t)y, Z = (uint64_t)z;
        int x = 0;
        while (x < nx) {
            const uint32_t l = row[x];
            if (!l) { ++x; continue; }
            uint64_t c = 0, sx = 0, sxx = 0;
            while (x < nx && row[x] == l) { const uint64_t X = (uint64_t)x; c += 1; sx += X; sxx += X * X; ++x; }
            uint64_t* s = sums + (size_t)(l - 1) * 10;
            const uint64_t v[10] = {c, sx, c * Y, c * Z, sxx, c * Y * Y, c * Z * Z, sx * Y, sx * Z, c * Y * Z};
            for (int j = 0; j < 10; ++j) {
#pragma omp atomic
                s[j] += v[j];
            }
        }
    }
    return 0;
}

/* ---- the whole of Mamri.py:1308-1309 on one volume ----------------------------------------------------- */
/* closed: uint8[n]; labels: uint32[n]; sums: caller passes capacity max_labels*10 (returns -3 if too small). */
API int oracle_detect2(const void* vol, int dtype, int nx, int ny, int nz, double lo, double hi, int open_radius, int radius,
                       int conn, uint8_t* closed, uint32_t* labels, uint32_t* n_labels, uint64_t* sums, uint32_t max_labels);

API int oracle_detect(const void* vol, int dtype, int nx, int ny, int nz, double lo, double hi, int radius, int conn,
                      uint8_t* closed, uint32_t* labels, uint32_t* n_labels, uint64_t* sums, uint32_t max_labels) {
    return oracle_detect2(vol, dtype, nx, ny, nz, lo, hi, 0, radius, conn, closed, labels, n_labels, sums, max_labels);
}

API int oracle_detect2(const void* vol, int dtype, int nx, int ny, int nz, double lo, double hi, int open_radius, int radius,
                       int conn, uint8_t* closed, uint32_t* labels, uint32_t* n_labels, uint64_t* sums, uint32_t max_labels) {
    const size_t n = (size_t)nx * ny * nz;
    uint8_t* bin = (uint8_t*)malloc(n);
    if (!bin) return -2;
    int rc = oracle_threshold(vol, dtype, n, lo, hi, bin);
    if (rc == 0 && open_radius > 0) {
        uint8_t* op = (uint8_t*)malloc(n);
        if (!op) { free(bin); return -2; }
        rc = oracle_opening(bin, nx, ny, nz, open_radius, op);
        free(bin);
        bin = op;
    }
    if (rc == 0) rc = oracle_closing(bin, nx, ny, nz, radius, closed);
    free(bin);
    if (rc == 0) rc = oracle_ccl(closed, nx, ny, nz, conn, labels, n_labels);
    if (rc == 0 && sums) {
        if (*n_labels > max_labels) return -3;
        rc = oracle_label_sums(labels, nx, ny, nz, *n_labels, sums);
    }
    return rc;
}
