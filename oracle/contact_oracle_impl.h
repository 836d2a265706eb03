/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Included twice by contact_oracle.c with
 * REAL = float (suffix _f32) and REAL = double (suffix _f64).
 *
 * Scalar CPU restatement of the reference's dense tensor algebra for the self-contact path.
 * Every function names the reference lines it follows (paths relative to /root/reference).
 * Compiled with -ffp-contract=off so each elementwise op rounds like a separate torch op.
 */

/* tuch/utils/contact.py:79-109 -- one (point, triangle) entry of solid_angles(), WITHOUT the
 * final factor 2 (returns atan2(numerator, denominator)). */
static inline REAL FN(half_solid_angle)(const REAL *p, const REAL *t)
{
    REAL a[3], b[3], c[3];
    for (int k = 0; k < 3; ++k) {           /* contact.py:80 centered_tris */
        a[k] = t[k] - p[k];
        b[k] = t[3 + k] - p[k];
        c[k] = t[6 + k] - p[k];
    }
    REAL na = SQRT((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);   /* contact.py:83 norms */
    REAL nb = SQRT((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
    REAL nc = SQRT((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
    REAL cr[3];                                                  /* contact.py:86-87 b x c */
    cr[0] = b[1] * c[2] - b[2] * c[1];
    cr[1] = b[2] * c[0] - b[0] * c[2];
    cr[2] = b[0] * c[1] - b[1] * c[0];
    REAL num = (a[0] * cr[0] + a[1] * cr[1]) + a[2] * cr[2];     /* contact.py:89 */
    REAL d01 = (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];        /* contact.py:92-94 */
    REAL d12 = (b[0] * c[0] + b[1] * c[1]) + b[2] * c[2];
    REAL d02 = (a[0] * c[0] + a[1] * c[1]) + a[2] * c[2];
    REAL den = ((na * nb) * nc + d01 * nc) + d02 * nb;           /* contact.py:97-102 */
    den = den + d12 * na;
    return ATAN2(num, den);                                      /* contact.py:106 */
}

/* tuch/utils/contact.py:49-109: out[q*F+f] = 2*atan2(...) */
void FN(oracle_solid_angles)(const REAL *points, const REAL *tris, int Q, int F, REAL *out)
{
#pragma omp parallel for schedule(static)
    for (int q = 0; q < Q; ++q)
        for (int f = 0; f < F; ++f)
            out[(size_t)q * F + f] = (REAL)2 * FN(half_solid_angle)(points + 3 * q, tris + 9 * f);
}

/* tuch/utils/contact.py:112-147: out[q] = 1/(4 pi) * sum_f solid_angle.  The sum over F is
 * accumulated in double (torch's fp32 reduction order is not reproducible scalar-wise). */
void FN(oracle_winding_numbers)(const REAL *points, const REAL *tris, int Q, int F, REAL *out)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int q = 0; q < Q; ++q) {
        double acc = 0.0;
        for (int f = 0; f < F; ++f)
            acc += (double)((REAL)2 * FN(half_solid_angle)(points + 3 * q, tris + 9 * f));
        out[q] = (REAL)(acc * (1.0 / (4.0 * 3.14159265358979323846)));
    }
}

/* tuch/utils/contact.py:23-47 (one batch element): P[i,j] = (|x_i|^2 + |y_j|^2) - 2 x_i.y_j,
 * optionally sqrt'ed (:44-45). */
static inline REAL FN(dot3)(const REAL *u, const REAL *v)
{
    return (u[0] * v[0] + u[1] * v[1]) + u[2] * v[2];
}

void FN(oracle_pairwise_dist)(const REAL *x, const REAL *y, int nx, int ny, int squared, REAL *P)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx; ++i) {
        REAL rx = FN(dot3)(x + 3 * i, x + 3 * i);
        for (int j = 0; j < ny; ++j) {
            REAL ry = FN(dot3)(y + 3 * j, y + 3 * j);
            REAL zz = FN(dot3)(x + 3 * i, y + 3 * j);
            REAL v = (rx + ry) - (REAL)2 * zz;
            P[(size_t)i * ny + j] = squared ? v : SQRT(v);
        }
    }
}

/* tuch/smplify/losses.py:76,92-93 and tuch/train/loss.py:256-270 (one body):
 * P = pairwise(v, v); P[r,c] = inf where !geomask[r,c]; per column c the first row index
 * attaining the minimum (argmin over axis=1 of the [1,V,V] tensor) and the minimum itself.
 * A fully masked column yields index 0 and +inf. */
void FN(oracle_masked_nearest)(const REAL *v, const unsigned char *geomask, int V,
                               int *argmin_out, REAL *min_out)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < V; ++c) {
        REAL best = (REAL)INFINITY;
        int bi = 0;
        REAL rc = FN(dot3)(v + 3 * c, v + 3 * c);
        for (int r = 0; r < V; ++r) {
            if (!geomask[(size_t)r * V + c]) continue;
            REAL rr = FN(dot3)(v + 3 * r, v + 3 * r);
            REAL zz = FN(dot3)(v + 3 * r, v + 3 * c);
            REAL p = (rr + rc) - (REAL)2 * zz;
            if (p < best) { best = p; bi = r; }
        }
        argmin_out[c] = bi;
        min_out[c] = best;
    }
}

/* tuch/smplify/losses.py:113-116 (one annotated pair of one body): min over the sub-matrix
 * P[idsA, :][:, idsB] of the geomask-masked squared distances; first flat index on ties.
 * Also used unmasked (geomask == NULL) for tuch/train/train_module.py:83-90. */
void FN(oracle_region_min)(const REAL *v, const unsigned char *geomask, int V,
                           const int *idsA, int nA, const int *idsB, int nB,
                           REAL *min_out, int *ia_out, int *ib_out)
{
    REAL best = (REAL)INFINITY;
    int ba = 0, bb = 0;
    for (int a = 0; a < nA; ++a) {
        int i = idsA[a];
        REAL ri = FN(dot3)(v + 3 * i, v + 3 * i);
        for (int b = 0; b < nB; ++b) {
            int j = idsB[b];
            REAL p;
            if (geomask && !geomask[(size_t)i * V + j]) {
                p = (REAL)INFINITY;
            } else {
                REAL rj = FN(dot3)(v + 3 * j, v + 3 * j);
                REAL zz = FN(dot3)(v + 3 * i, v + 3 * j);
                p = (ri + rj) - (REAL)2 * zz;
            }
            if (p < best) { best = p; ba = a; bb = b; }
        }
    }
    *min_out = best;
    *ia_out = ba;
    *ib_out = bb;
}
