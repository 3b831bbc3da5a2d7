/*
 * TEST INFRASTRUCTURE — CPU oracle for the nearest-cylinder hot path.
 *
 * A plain-C restatement of the reference algorithm, one separately rounded fp32 operation
 * per reference tensor op, in the reference's own order ("mirror order"):
 *
 *   variant A  PreProcessing/LabelGenerationCuda.py:20-111   (perp_atol 1e-6, norm_eps 0)
 *   variant B  Modules/Projection.py:19-115                  (perp_atol 1e-3, norm_eps 1e-8)
 *   prep       LabelGenerationCuda.py:117-123 / Projection.py:121-132
 *
 * Parity pinning: the reference ships no tests or golden vectors for this path, so this file
 * is pinned against outputs of the reference itself, executed on CPU in the build container
 * (tests/golden/make_golden.py writes the committed fixtures; tests/test_oracle_golden.py and
 * tests/test_oracle_vs_reference.py compare bit-for-bit).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (the CUDA library) never does.
 *
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off: no FMA contraction, IEEE div/sqrt).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TM_ORACLE_VERSION 1

typedef struct {
    float dist;
    float fx, fy, fz;          /* final_projection_points (reference line A:81 / B:84) */
    float nsx, nsy, nsz;       /* new_axis_start */
    float nex, ney, nez;       /* new_axis_end == surface_projection_points bit-for-bit */
    float px, py, pz;          /* projection_on_new_axis */
    int perp;
} pair_eval;

/* torch.norm over a length-3 axis.  Two CPU code paths exist in ATen and they round differently:
 *   fma == 0: reduced axis is strided (cylinder tensors built from DataFrame.values are
 *             Fortran-ordered, LabelGenerationCuda.py:117 → every (N,M,3) temporary has the xyz
 *             axis at stride M): sqrt((x*x + y*y) + z*z), three separately rounded products;
 *   fma == 1: reduced axis is contiguous (C-ordered tensors as QSMFittingDepthFirst.py:1039 builds
 *             them, or M == 1): sqrt(fma(z,z, fma(y,y, x*x))).
 * torch.sum over the same axis is (x + y) + z in both layouts.  Probed in the build container. */
static inline float norm3(float x, float y, float z, int fma)
{
    if (fma) return sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
    return sqrtf((x * x + y * y) + z * z);
}

/* torch.clamp(x, lo, hi) for finite lo/hi: NaN in x propagates (A:43, A:74). */
static inline float clampf_nanprop(float x, float lo, float hi)
{
    float y = x < lo ? lo : x;
    return y > hi ? hi : y;
}

/* One (point, cylinder) pair; every line is one reference tensor op.  Citations: A = variant A line. */
static inline void eval_pair(float px, float py, float pz,
                             float sx, float sy, float sz, float ux, float uy, float uz,
                             float len, float rad, float atol, float eps, int fma, pair_eval *o)
{
    /* A:36  point_vectors */
    float vx = px - sx, vy = py - sy, vz = pz - sz;
    /* A:39  projection_lengths = sum(point_vectors * axis_unit) */
    float t = (vx * ux + vy * uy) + vz * uz;
    /* A:42-43 clamp to [0, axis_length] */
    float tc = clampf_nanprop(t, 0.0f, len);
    /* A:44  projection_points_clamped */
    float qx = sx + tc * ux, qy = sy + tc * uy, qz = sz + tc * uz;
    /* A:47  projection_vectors */
    float wx = px - qx, wy = py - qy, wz = pz - qz;
    /* A:50  dot_products */
    float d = (wx * ux + wy * uy) + wz * uz;
    /* A:51  isclose(d, 0, atol): |d| <= atol, false for NaN/Inf */
    int perp = fabsf(d) <= atol;
    /* A:54-55 rejected_vectors */
    float rx = wx - d * ux, ry = wy - d * uy, rz = wz - d * uz;
    /* A:58  norm_rejected */
    float rho = norm3(rx, ry, rz, fma);
    /* B:60-62 safe_norm_rejected (variant A: eps == 0 → unguarded) */
    float rho_s = (eps > 0.0f && rho < eps) ? eps : rho;
    /* A:60  new_axis_unit (true division) */
    float nx = rx / rho_s, ny = ry / rho_s, nz = rz / rho_s;
    /* A:63  new_axis_scaled = n * (2 r) */
    float r2 = 2.0f * rad;
    float scx = nx * r2, scy = ny * r2, scz = nz * r2;
    /* A:66-67 new axis end points */
    float hx = 0.5f * scx, hy = 0.5f * scy, hz = 0.5f * scz;
    float nsx = qx - hx, nsy = qy - hy, nsz = qz - hz;
    float nex = qx + hx, ney = qy + hy, nez = qz + hz;
    /* A:70  projection_length */
    float pl = ((px - nsx) * nx + (py - nsy) * ny) + (pz - nsz) * nz;
    /* A:73-74 clamp to [0, 2r] */
    float plc = clampf_nanprop(pl, 0.0f, r2);
    /* A:75  projection_on_new_axis */
    float pox = nsx + plc * nx, poy = nsy + plc * ny, poz = nsz + plc * nz;
    /* A:78  surface_projection_points = q + (rej / rho) * r */
    float sfx = qx + (rx / rho_s) * rad, sfy = qy + (ry / rho_s) * rad, sfz = qz + (rz / rho_s) * rad;
    /* A:81  final_projection_points */
    float fx = perp ? sfx : pox, fy = perp ? sfy : poy, fz = perp ? sfz : poz;
    /* A:84  distances */
    float ex = px - fx, ey = py - fy, ez = pz - fz;
    o->dist = norm3(ex, ey, ez, fma);
    o->fx = fx; o->fy = fy; o->fz = fz;
    o->nsx = nsx; o->nsy = nsy; o->nsz = nsz;
    o->nex = nex; o->ney = ney; o->nez = nez;
    o->px = pox; o->py = poy; o->pz = poz;
    o->perp = perp;
    (void)sfx; (void)sfy; (void)sfz;
}

/* distance only — the hot inner loop, kept free of stores so the compiler can vectorise it */
static inline float eval_dist(float px, float py, float pz,
                              float sx, float sy, float sz, float ux, float uy, float uz,
                              float len, float rad, float atol, float eps, int fma)
{
    float vx = px - sx, vy = py - sy, vz = pz - sz;
    float t = (vx * ux + vy * uy) + vz * uz;
    float tc = clampf_nanprop(t, 0.0f, len);
    float qx = sx + tc * ux, qy = sy + tc * uy, qz = sz + tc * uz;
    float wx = px - qx, wy = py - qy, wz = pz - qz;
    float d = (wx * ux + wy * uy) + wz * uz;
    int perp = fabsf(d) <= atol;
    float rx = wx - d * ux, ry = wy - d * uy, rz = wz - d * uz;
    float rho = norm3(rx, ry, rz, fma);
    float rho_s = (eps > 0.0f && rho < eps) ? eps : rho;
    float nx = rx / rho_s, ny = ry / rho_s, nz = rz / rho_s;
    float r2 = 2.0f * rad;
    float hx = 0.5f * (nx * r2), hy = 0.5f * (ny * r2), hz = 0.5f * (nz * r2);
    float nsx = qx - hx, nsy = qy - hy, nsz = qz - hz;
    float pl = ((px - nsx) * nx + (py - nsy) * ny) + (pz - nsz) * nz;
    float plc = clampf_nanprop(pl, 0.0f, r2);
    float pox = nsx + plc * nx, poy = nsy + plc * ny, poz = nsz + plc * nz;
    float sfx = qx + nx * rad, sfy = qy + ny * rad, sfz = qz + nz * rad;
    float fx = perp ? sfx : pox, fy = perp ? sfy : poy, fz = perp ? sfz : poz;
    float ex = px - fx, ey = py - fy, ez = pz - fz;
    return norm3(ex, ey, ez, fma);
}

#if defined(__x86_64__) && defined(__GNUC__) && !defined(TM_ORACLE_NO_CLONES)
#define TM_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define TM_CLONES
#endif

/* all M distances of one point into `out` (SoA cylinder columns) */
TM_CLONES
static void dist_row(float px, float py, float pz, int64_t m,
                     const float *sx, const float *sy, const float *sz,
                     const float *ux, const float *uy, const float *uz,
                     const float *len, const float *rad, float atol, float eps, int fma, float *out)
{
#pragma omp simd
    for (int64_t j = 0; j < m; ++j)
        out[j] = eval_dist(px, py, pz, sx[j], sy[j], sz[j], ux[j], uy[j], uz[j], len[j], rad[j], atol, eps, fma);
}

/* torch.argmin over one row: NaN beats everything, lowest index among equals
 * (ATen/native/SharedReduceOps.h LessOrNan).  Also reports the runner-up distance. */
static void argmin_row(const float *d, int64_t m, int64_t *best_j, float *second)
{
    int64_t bj = 0;
    float b = d[0];
    float s = INFINITY;            /* smallest distance among the non-winners (NaN-free view) */
    for (int64_t j = 1; j < m; ++j) {
        float x = d[j];
        int better = isnan(b) ? 0 : (isnan(x) ? 1 : (x < b));
        if (better) {
            if (!isnan(b) && b < s) s = b;
            b = x; bj = j;
        } else if (!isnan(x) && x < s) {
            s = x;
        }
    }
    *best_j = bj;
    *second = s;
}

int tm_oracle_version(void) { return TM_ORACLE_VERSION; }

int tm_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Cylinder preparation (LabelGenerationCuda.py:121-123; Projection.py:126-132 with guard_eps=1e-8):
 *   axis = end - start; axis_length = sqrt((x^2 + y^2) + z^2); axis_unit = axis / max-guarded length.
 * start/end are (M,3) row-major.  out_len is (M,), out_unit (M,3).
 */
int tm_oracle_prepare(const float *start, const float *end, int64_t m, float guard_eps, int norm_fma,
                      float *out_len, float *out_unit)
{
    for (int64_t j = 0; j < m; ++j) {
        float ax = end[3 * j] - start[3 * j], ay = end[3 * j + 1] - start[3 * j + 1], az = end[3 * j + 2] - start[3 * j + 2];
        float l = norm3(ax, ay, az, norm_fma);
        float dv = (guard_eps > 0.0f && l < guard_eps) ? guard_eps : l;
        out_len[j] = l;
        out_unit[3 * j] = ax / dv; out_unit[3 * j + 1] = ay / dv; out_unit[3 * j + 2] = az / dv;
    }
    return 0;
}

/*
 * closest_cylinder_cuda_batch for N points against M cylinders.
 *   pts (N,3), start (M,3), unit (M,3) row-major fp32; length (M,), radius (M,), ids (M,) int32.
 *   out_index (N,) row index of the winner, out_id (N,) = ids[index], out_dist (N,), out_offset (N,3),
 *   out_second (N,) nullable: smallest distance among the other cylinders (for near-tie bookkeeping).
 *   norm_fma: 0 = DataFrame (Fortran-ordered) layout, 1 = contiguous layout, see norm3().
 *   move_to_mantle: 1 = reference behaviour (A always; B default). 0 = offset to the distance foot
 *   point (the reference's own False branch is shape-broken, Projection.py:110; this is the
 *   evident intent and is documented as an extension).
 * Returns 0, or 1 for M == 0 with N > 0 (the reference raises in argmin on an empty dim).
 */
int tm_oracle_label(const float *pts, int64_t n, const float *start, const float *unit,
                    const float *length, const float *radius, const int32_t *ids, int64_t m,
                    float perp_atol, float norm_eps, int move_to_mantle, int norm_fma,
                    int32_t *out_index, int32_t *out_id, float *out_dist, float *out_offset,
                    float *out_second, int nthreads)
{
    if (n <= 0) return 0;
    if (m <= 0) return 1;
    float *cols = (float *)malloc(sizeof(float) * 8 * (size_t)m);
    if (!cols) return 2;
    float *sx = cols, *sy = cols + m, *sz = cols + 2 * m, *ux = cols + 3 * m, *uy = cols + 4 * m,
          *uz = cols + 5 * m, *ln = cols + 6 * m, *rd = cols + 7 * m;
    for (int64_t j = 0; j < m; ++j) {
        sx[j] = start[3 * j]; sy[j] = start[3 * j + 1]; sz[j] = start[3 * j + 2];
        ux[j] = unit[3 * j]; uy[j] = unit[3 * j + 1]; uz[j] = unit[3 * j + 2];
        ln[j] = length[j]; rd[j] = radius[j];
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    int fail = 0;
#pragma omp parallel num_threads(nthreads)
    {
        float *row = (float *)malloc(sizeof(float) * (size_t)m);
        if (!row) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            if (!row) continue;
            float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
            dist_row(px, py, pz, m, sx, sy, sz, ux, uy, uz, ln, rd, perp_atol, norm_eps, norm_fma, row);
            int64_t bj; float second;
            argmin_row(row, m, &bj, &second);
            pair_eval e;
            eval_pair(px, py, pz, sx[bj], sy[bj], sz[bj], ux[bj], uy[bj], uz[bj], ln[bj], rd[bj],
                      perp_atol, norm_eps, norm_fma, &e);
            float mx, my, mz;
            if (move_to_mantle) {
                /* A:92-100: nearer end of the new axis, unless perpendicular → surface point */
                float ax = e.px - e.nsx, ay = e.py - e.nsy, az = e.pz - e.nsz;
                float bx = e.px - e.nex, by = e.py - e.ney, bz = e.pz - e.nez;
                float ds = norm3(ax, ay, az, norm_fma);
                float de = norm3(bx, by, bz, norm_fma);
                int to_start = ds < de;
                mx = e.perp ? e.nex : (to_start ? e.nsx : e.nex);
                my = e.perp ? e.ney : (to_start ? e.nsy : e.ney);
                mz = e.perp ? e.nez : (to_start ? e.nsz : e.nez);
            } else {
                mx = e.fx; my = e.fy; mz = e.fz;
            }
            out_index[i] = (int32_t)bj;
            out_id[i] = ids ? ids[bj] : (int32_t)bj;
            out_dist[i] = row[bj];
            /* A:106 closest_offsets = final_projection_points - points */
            out_offset[3 * i] = mx - px; out_offset[3 * i + 1] = my - py; out_offset[3 * i + 2] = mz - pz;
            if (out_second) out_second[i] = second;
        }
        free(row);
    }
    free(cols);
    return fail ? 2 : 0;
}

/* All M distances of a handful of points (N*M floats, row-major) — used by tests that need the
 * full distance matrix (tie / NaN semantics) at small sizes. */
int tm_oracle_distance_matrix(const float *pts, int64_t n, const float *start, const float *unit,
                              const float *length, const float *radius, int64_t m,
                              float perp_atol, float norm_eps, int norm_fma, float *out)
{
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < m; ++j)
            out[i * m + j] = eval_dist(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2],
                                       start[3 * j], start[3 * j + 1], start[3 * j + 2],
                                       unit[3 * j], unit[3 * j + 1], unit[3 * j + 2],
                                       length[j], radius[j], perp_atol, norm_eps, norm_fma);
    return 0;
}
