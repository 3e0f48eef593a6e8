/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's native RPC path.
 * Never linked into the product library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg load it (as the checker / the timed baseline).
 *
 * Restated from (reference file:line):
 *   cubic in 20 monomials, RPC00B order            c/rpc.c:279-298   eval_pol20
 *   projection  (lon,lat,alt) -> (col,row)         c/rpc.c:442-452   eval_rpci, :337-349 eval_nrpci
 *   localisation (col,row,alt) -> (lon,lat)        c/rpc.c:429-439   eval_rpc, :414-426 eval_nrpc,
 *                                                  c/rpc.c:378-411   eval_nrpc_iterative, :361-374 basis solve
 *   image-to-image transfer at a height            c/rpc.c:455-462   eval_rpc_pair
 *   two-view height                                c/rpc.c:480-514   rpc_height (<=100 its, |lambda|<1e-5)
 *   per-match triangulation loop                   c/disp_to_h.c:40-65 stereo_corresp_to_lonlatalt
 * The struct layout is the ABI of c/rpc.h:14-32 (what bundle_adjust/s2p/triangulation.py:22-36 mirrors).
 * Parity pin: compared against oracle/_ref/disp_to_h.so (the reference's own sources compiled in
 * place by oracle/Makefile) in tests/test_oracle_pin.py.
 */
#include <math.h>
#include <stddef.h>

typedef struct rpco_model {
    double numx[20], denx[20], numy[20], deny[20];      /* direct model (col,row,alt) -> (lon,lat); NaN if absent */
    double scale[3], offset[3];                         /* col, row, alt */
    double inumx[20], idenx[20], inumy[20], ideny[20];  /* inverse model (lon,lat,alt) -> (col,row) */
    double iscale[3], ioffset[3];                       /* lon, lat, alt */
    double dmval[4], imval[4];
    double delta;                                       /* first probe of the iterative localisation */
} rpco_model;

double rpco_poly20(const double c[20], double lon, double lat, double alt)
{
    const double m[20] = {
        1, lon, lat, alt, lon * lat,
        lon * alt, lat * alt, lon * lon, lat * lat, alt * alt,
        lat * lon * alt, lon * lon * lon, lon * lat * lat, lon * alt * alt, lon * lon * lat,
        lat * lat * lat, lat * alt * alt, lon * lon * alt, lat * lat * alt, alt * alt * alt};
    double acc = 0;
    for (int i = 0; i < 20; i++) acc += c[i] * m[i];
    return acc;
}

/* normalised (lon,lat,alt) -> normalised (col,row) */
static void nproject(double out[2], const rpco_model *p, double lon, double lat, double alt)
{
    out[0] = rpco_poly20(p->inumx, lon, lat, alt) / rpco_poly20(p->idenx, lon, lat, alt);
    out[1] = rpco_poly20(p->inumy, lon, lat, alt) / rpco_poly20(p->ideny, lon, lat, alt);
}

void rpco_project(double out[2], const rpco_model *p, double lon, double lat, double alt)
{
    double n[2];
    nproject(n, p, (lon - p->ioffset[0]) / p->iscale[0], (lat - p->ioffset[1]) / p->iscale[1],
             (alt - p->ioffset[2]) / p->iscale[2]);
    out[0] = n[0] * p->scale[0] + p->offset[0];
    out[1] = n[1] * p->scale[1] + p->offset[1];
}

/* normalised (col,row,alt) -> normalised (lon,lat) by repeated affine inversion of the projection */
static void nlocalize_iterative(double out[2], const rpco_model *p, double x, double y, double alt)
{
    double d = p->delta ? p->delta : 1.0;
    double lon = -d, lat = -d, eps = 2 * d;
    double q0[2], q1[2], q2[2];
    nproject(q0, p, lon, lat, alt);
    nproject(q1, p, lon + eps, lat, alt);
    nproject(q2, p, lon, lat + eps, alt);
    while ((q0[0] - x) * (q0[0] - x) + (q0[1] - y) * (q0[1] - y) > 1e-18) {
        double ux = x - q0[0], uy = y - q0[1];
        double ax = q1[0] - q0[0], ay = q1[1] - q0[1];
        double bx = q2[0] - q0[0], by = q2[1] - q0[1];
        double det = ax * by - ay * bx;
        double c0 = (by * ux - bx * uy) / det;
        double c1 = (-ay * ux + ax * uy) / det;
        lon += c0 * eps;
        lat += c1 * eps;
        eps = 0.1;
        nproject(q0, p, lon, lat, alt);
        nproject(q1, p, lon + eps, lat, alt);
        nproject(q2, p, lon, lat + eps, alt);
    }
    out[0] = lon;
    out[1] = lat;
}

void rpco_localize(double out[2], const rpco_model *p, double col, double row, double alt)
{
    double x = (col - p->offset[0]) / p->scale[0];
    double y = (row - p->offset[1]) / p->scale[1];
    double z = (alt - p->offset[2]) / p->scale[2];
    double n[2];
    if (isfinite(p->numx[0])) {
        n[0] = rpco_poly20(p->numx, x, y, z) / rpco_poly20(p->denx, x, y, z);
        n[1] = rpco_poly20(p->numy, x, y, z) / rpco_poly20(p->deny, x, y, z);
    } else {
        nlocalize_iterative(n, p, x, y, z);
    }
    out[0] = n[0] * p->iscale[0] + p->ioffset[0];
    out[1] = n[1] * p->iscale[1] + p->ioffset[1];
}

static void transfer(double out[2], const rpco_model *a, const rpco_model *b, double col, double row, double h)
{
    double ll[2];
    rpco_localize(ll, a, col, row, h);
    rpco_project(out, b, ll[0], ll[1], h);
}

double rpco_height(const rpco_model *a, const rpco_model *b, double xa, double ya, double xb, double yb,
                   double *outerr)
{
    double h = 0;
    for (int it = 0; it < 100; it++) {
        double p[2], q[2];
        transfer(p, a, b, xa, ya, h);
        transfer(q, a, b, xa, ya, h + 1);
        double dx = q[0] - p[0], dy = q[1] - p[1];
        double ex = xb - p[0], ey = yb - p[1];
        double lambda = (dx * ex + dy * ey) / (dx * dx + dy * dy);
        double zx = p[0] + lambda * dx, zy = p[1] + lambda * dy;
        if (outerr) *outerr = hypot(zx - xb, zy - yb);
        h += lambda;
        if (fabs(lambda) < 0.00001) break;
    }
    return h;
}

void rpco_stereo_corresp_to_lonlatalt(double *lonlatalt, float *err, const float *kp_a, const float *kp_b, int n_kp,
                                      const rpco_model *rpc_a, const rpco_model *rpc_b)
{
    for (int i = 0; i < n_kp; i++) {
        double e, ll[2];
        double z = rpco_height(rpc_a, rpc_b, kp_a[2 * i], kp_a[2 * i + 1], kp_b[2 * i], kp_b[2 * i + 1], &e);
        rpco_localize(ll, rpc_a, kp_a[2 * i], kp_a[2 * i + 1], z);
        lonlatalt[3 * i] = ll[0];
        lonlatalt[3 * i + 1] = ll[1];
        lonlatalt[3 * i + 2] = z;
        err[i] = (float)e;
    }
}

/* batched helpers so that Python can time / compare whole arrays without per-point call overhead */
void rpco_project_batch(double *colrow, const rpco_model *p, const double *lonlatalt, int n)
{
    for (int i = 0; i < n; i++) rpco_project(colrow + 2 * i, p, lonlatalt[3 * i], lonlatalt[3 * i + 1], lonlatalt[3 * i + 2]);
}

void rpco_localize_batch(double *lonlat, const rpco_model *p, const double *colrowalt, int n)
{
    for (int i = 0; i < n; i++) rpco_localize(lonlat + 2 * i, p, colrowalt[3 * i], colrowalt[3 * i + 1], colrowalt[3 * i + 2]);
}
