/* CPU oracle (plain C) for the capsule dynamic-routing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build / load this.
 * It is never linked into the product library.
 *
 * Closed-form restatement (explicit backward, no autograd) of
 *   reference models.py:64-67   squash
 *   reference models.py:70-79   u_hat = u.W, then R x (softmax_j, weighted sum, squash, agreement)
 *   reference models.py:117     scores = |v|
 *   reference loss_fns.py:12-17,23  margin loss (recon off)
 * One sample at a time (u_hat for one sample is N*C*D reals).  Threading is done by the caller
 * (oracle/routing_c.py runs fixed 8-sample chunks on a thread pool and adds the per-chunk dW
 * partials in chunk order, so results do not depend on the thread count).
 * Pinned by tests/test_oracle.py against tests/golden/*.npz, which the unmodified reference wrote.
 *
 * The file is compiled twice by oracle/Makefile: -DREAL=float (symbols *_f32) and
 * -DREAL=double (symbols *_f64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#endif
#ifndef SUFFIX
#define SUFFIX f64
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

static void squash_vec(const REAL *s, REAL *v, int D) {          /* models.py:64-67 */
    REAL n2 = 0;
    for (int d = 0; d < D; ++d) n2 += s[d] * s[d];
    REAL f = (n2 / (1 + n2)) / (REAL)sqrt((double)n2);
    for (int d = 0; d < D; ++d) v[d] = f * s[d];
}

static void squash_bwd_vec(const REAL *s, const REAL *dv, REAL *ds, int D) {
    REAL n2 = 0, sdv = 0;
    for (int d = 0; d < D; ++d) { n2 += s[d] * s[d]; sdv += s[d] * dv[d]; }
    REAL n = (REAL)sqrt((double)n2);
    REAL a = n / (1 + n2);
    REAL b = sdv * (1 - n2) / (n * (1 + n2) * (1 + n2));
    for (int d = 0; d < D; ++d) ds[d] = a * dv[d] + b * s[d];
}

/* One sample: forward, margin-loss gradient (+ optional external grad), backward.
 * Outputs may be NULL.  dW_acc (N*C*K*D) is ADDED to.  Returns the sample's margin loss
 * (already multiplied by inv_batch). */
static REAL one_sample(const REAL *u, const REAL *W, long y, const REAL *gext, REAL inv_batch,
                       int N, int C, int K, int D, int R,
                       REAL *v_out, REAL *c_out, REAL *du, REAL *dW_acc, int do_bwd) {
    size_t ncd = (size_t)N * C * D, nc = (size_t)N * C, cd = (size_t)C * D;
    REAL *uh = (REAL *)malloc(sizeof(REAL) * ncd);
    REAL *blog = (REAL *)calloc(nc, sizeof(REAL));                 /* models.py:72 */
    REAL *cs = (REAL *)malloc(sizeof(REAL) * nc * R);              /* c^r */
    REAL *ss = (REAL *)malloc(sizeof(REAL) * cd * R);              /* s^r */
    REAL *vs = (REAL *)malloc(sizeof(REAL) * cd * R);              /* v^r */
    /* models.py:71 */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < C; ++j) {
            REAL *o = uh + ((size_t)i * C + j) * D;
            const REAL *w = W + ((size_t)i * C + j) * K * D;
            for (int d = 0; d < D; ++d) o[d] = 0;
            for (int k = 0; k < K; ++k) {
                REAL uk = u[(size_t)i * K + k];
                for (int d = 0; d < D; ++d) o[d] += uk * w[k * D + d];
            }
        }
    for (int r = 0; r < R; ++r) {
        REAL *c = cs + nc * r, *s = ss + cd * r, *v = vs + cd * r;
        memset(s, 0, sizeof(REAL) * cd);
        for (int i = 0; i < N; ++i) {
            const REAL *b = blog + (size_t)i * C;
            REAL mx = b[0], z = 0;
            for (int j = 1; j < C; ++j) mx = b[j] > mx ? b[j] : mx;
            for (int j = 0; j < C; ++j) { c[(size_t)i * C + j] = (REAL)exp((double)(b[j] - mx)); z += c[(size_t)i * C + j]; }
            for (int j = 0; j < C; ++j) {
                REAL cij = c[(size_t)i * C + j] / z;               /* models.py:75 */
                c[(size_t)i * C + j] = cij;
                const REAL *o = uh + ((size_t)i * C + j) * D;
                for (int d = 0; d < D; ++d) s[j * D + d] += cij * o[d];   /* models.py:76 */
            }
        }
        for (int j = 0; j < C; ++j) squash_vec(s + j * D, v + j * D, D);
        if (r != R - 1)                                            /* models.py:77-79 */
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < C; ++j) {
                    const REAL *o = uh + ((size_t)i * C + j) * D;
                    REAL a = 0;
                    for (int d = 0; d < D; ++d) a += o[d] * v[j * D + d];
                    blog[(size_t)i * C + j] += a;
                }
    }
    const REAL *vL = vs + cd * (R - 1);
    if (v_out) memcpy(v_out, vL, sizeof(REAL) * cd);
    if (c_out) memcpy(c_out, cs + nc * (R - 1), sizeof(REAL) * nc);

    /* margin loss and its gradient (models.py:117, loss_fns.py:12-17,23) */
    REAL loss = 0;
    REAL *dv = (REAL *)calloc(cd, sizeof(REAL));
    if (y >= 0)
        for (int j = 0; j < C; ++j) {
            REAL m2 = 0;
            for (int d = 0; d < D; ++d) m2 += vL[j * D + d] * vL[j * D + d];
            REAL m = (REAL)sqrt((double)m2);
            REAL T = (j == y) ? 1 : 0;
            REAL l = (REAL)0.9 - m; l = l > 0 ? l : 0;
            REAL rr = m - (REAL)0.1; rr = rr > 0 ? rr : 0;
            loss += (T * l * l + (REAL)0.5 * (1 - T) * rr * rr) * inv_batch;
            REAL dm = (-2 * T * l + (1 - T) * rr) * inv_batch;
            for (int d = 0; d < D; ++d) dv[j * D + d] = dm * vL[j * D + d] / m;
        }
    if (gext) for (size_t t = 0; t < cd; ++t) dv[t] += gext[t];

    if (do_bwd) {
        REAL *G = (REAL *)calloc(ncd, sizeof(REAL));
        REAL *beta = (REAL *)calloc(nc, sizeof(REAL));
        REAL *ds = (REAL *)malloc(sizeof(REAL) * cd);
        REAL *dvn = (REAL *)malloc(sizeof(REAL) * cd);
        for (int r = R - 1; r >= 0; --r) {
            const REAL *c = cs + nc * r, *s = ss + cd * r;
            for (int j = 0; j < C; ++j) squash_bwd_vec(s + j * D, dv + j * D, ds + j * D, D);
            memset(dvn, 0, sizeof(REAL) * cd);
            for (int i = 0; i < N; ++i) {
                REAL t = 0;
                REAL dc[1024];
                for (int j = 0; j < C; ++j) {
                    const REAL *o = uh + ((size_t)i * C + j) * D;
                    REAL *g = G + ((size_t)i * C + j) * D;
                    REAL cij = c[(size_t)i * C + j], a = 0;
                    for (int d = 0; d < D; ++d) { g[d] += cij * ds[j * D + d]; a += o[d] * ds[j * D + d]; }
                    dc[j] = a; t += cij * a;
                }
                if (r > 0) {
                    const REAL *vp = vs + cd * (r - 1);
                    for (int j = 0; j < C; ++j) {
                        const REAL *o = uh + ((size_t)i * C + j) * D;
                        REAL *g = G + ((size_t)i * C + j) * D;
                        REAL bt = beta[(size_t)i * C + j] + c[(size_t)i * C + j] * (dc[j] - t);
                        beta[(size_t)i * C + j] = bt;
                        for (int d = 0; d < D; ++d) { g[d] += bt * vp[j * D + d]; dvn[j * D + d] += bt * o[d]; }
                    }
                }
            }
            memcpy(dv, dvn, sizeof(REAL) * cd);
        }
        for (int i = 0; i < N; ++i) {
            REAL acc[64];
            for (int k = 0; k < K; ++k) acc[k] = 0;
            for (int j = 0; j < C; ++j) {
                const REAL *g = G + ((size_t)i * C + j) * D;
                const REAL *w = W + ((size_t)i * C + j) * K * D;
                REAL *dw = dW_acc ? dW_acc + ((size_t)i * C + j) * K * D : NULL;
                for (int k = 0; k < K; ++k) {
                    REAL uk = u[(size_t)i * K + k], a = 0;
                    for (int d = 0; d < D; ++d) { a += w[k * D + d] * g[d]; if (dw) dw[k * D + d] += uk * g[d]; }
                    acc[k] += a;
                }
            }
            if (du) for (int k = 0; k < K; ++k) du[(size_t)i * K + k] = acc[k];
        }
        free(G); free(beta); free(ds); free(dvn);
    }
    free(uh); free(blog); free(cs); free(ss); free(vs); free(dv);
    return loss;
}

/* Samples [0,B) of the given arrays, sequentially.  y may be NULL (no margin loss), gext may be
 * NULL, outputs may be NULL.  C <= 1024, K <= 64.  dW is ADDED to (caller zeroes it); *loss is
 * overwritten with the sum over these samples.  Returns 0, or -1 on bad arguments. */
int FN(caps_oracle_step)(const REAL *u, const REAL *W, const int64_t *y, const REAL *gext,
                         REAL inv_batch, int B, int N, int C, int K, int D, int R,
                         REAL *v, REAL *c, REAL *loss, REAL *du, REAL *dW, int do_bwd) {
    if (C > 1024 || K > 64 || B < 0 || R < 1) return -1;
    REAL l = 0;
    for (int b = 0; b < B; ++b)
        l += one_sample(u + (size_t)b * N * K, W, y ? (long)y[b] : -1,
                        gext ? gext + (size_t)b * C * D : NULL, inv_batch, N, C, K, D, R,
                        v ? v + (size_t)b * C * D : NULL, c ? c + (size_t)b * N * C : NULL,
                        du ? du + (size_t)b * N * K : NULL, dW, do_bwd);
    if (loss) *loss = l;
    return 0;
}
