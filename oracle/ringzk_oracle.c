/*
 * ringzk_oracle.c -- CPU restatement of AlvinHon/ring-zk's R_q hot path.
 * TEST INFRASTRUCTURE ONLY; see ringzk_oracle.h for the parity status
 * ("parity unpinned" for ring products; pins listed there).
 *
 * Deliberately a different algorithm from the GPU engine: schoolbook O(N^2)
 * negacyclic products with exact wide accumulators, and the reference's own
 * operation order (including the multiplications by the identity / zero key
 * blocks that Mat::dot does not skip, mat.rs:106-113).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference/src).
 */
#include "ringzk_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <assert.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __int128 i128;
typedef unsigned __int128 u128;

static __thread uint64_t g_products = 0;   /* per calling thread */

uint64_t rzko_product_count(int reset)
{
    uint64_t v = g_products;
    if (reset) g_products = 0;
    return v;
}

int rzko_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* params.rs:121-138 */
rzko_params rzko_default_params(int32_t N)
{
    rzko_params P;
    P.q = 3515337053LL;
    P.b = 1;
    P.N = N;
    P.n = 1;
    P.k = 3;
    P.l = 1;
    P.kappa = 36;
    return P;
}

/* Canonical centred residue in [-(q-1)/2, (q-1)/2] (q odd).  SURVEY.md 8(c). */
int64_t rzko_center(int64_t v, int64_t q)
{
    int64_t r = v % q;
    int64_t half = (q - 1) / 2;
    if (r > half) r -= q;
    else if (r < -half) r += q;
    return r;
}

static int64_t center128(i128 v, int64_t q)
{
    int64_t r = (int64_t)(v % (i128)q);
    int64_t half = (q - 1) / 2;
    if (r > half) r -= q;
    else if (r < -half) r += q;
    return r;
}

static uint64_t isqrt_u64(uint64_t v)
{
    if (v == 0) return 0;
    uint64_t x = (uint64_t)__builtin_sqrtl((long double)v);
    while ((u128)x * x > v) --x;
    while ((u128)(x + 1) * (x + 1) <= v) ++x;
    return x;
}

static uint64_t isqrt_u128(u128 v)
{
    if ((v >> 64) == 0) return isqrt_u64((uint64_t)v);
    /* Newton from above */
    u128 x = (u128)1 << 64;
    for (;;) {
        u128 y = (x + v / x) >> 1;
        if (y >= x) break;
        x = y;
    }
    return (uint64_t)x;
}

/* ---- Polynomial<ZqI64<Q>, N> operators (crate poly-ring-xnp1, not in tree) ----
 * c_k = sum_{i+j=k} a_i b_j - sum_{i+j=k+N} a_i b_j, reduced to the centred residue. */
void rzko_poly_mul(const rzko_params *P, const int64_t *a, const int64_t *b, int64_t *out)
{
    const int N = P->N;
    const int64_t q = P->q;
    g_products++;
    int64_t ma = 0, mb = 0;
    for (int i = 0; i < N; ++i) {
        int64_t va = a[i] < 0 ? -a[i] : a[i];
        int64_t vb = b[i] < 0 ? -b[i] : b[i];
        if (va > ma) ma = va;
        if (vb > mb) mb = vb;
    }
    /* exact in int64 when N * ma * mb < 2^62 */
    int small = 0;
    if (ma == 0 || mb == 0) small = 1;
    else {
        u128 bound = (u128)ma * (u128)mb * (u128)N;
        small = bound < ((u128)1 << 62);
    }
    if (small) {
        int64_t *acc = (int64_t *)calloc((size_t)2 * N, sizeof(int64_t));
        for (int i = 0; i < N; ++i) {
            const int64_t ai = a[i];
            if (ai == 0) continue;
            int64_t *row = acc + i;
            for (int j = 0; j < N; ++j) row[j] += ai * b[j];
        }
        for (int k = 0; k < N; ++k) out[k] = rzko_center(acc[k] - acc[k + N], q);
        free(acc);
    } else {
        i128 *acc = (i128 *)calloc((size_t)2 * N, sizeof(i128));
        for (int i = 0; i < N; ++i) {
            const i128 ai = a[i];
            if (ai == 0) continue;
            i128 *row = acc + i;
            for (int j = 0; j < N; ++j) row[j] += ai * (i128)b[j];
        }
        for (int k = 0; k < N; ++k) out[k] = center128(acc[k] - acc[k + N], q);
        free(acc);
    }
}

void rzko_poly_add(const rzko_params *P, const int64_t *a, const int64_t *b, int64_t *out)
{
    for (int i = 0; i < P->N; ++i) out[i] = rzko_center(a[i] + b[i], P->q);
}

void rzko_poly_sub(const rzko_params *P, const int64_t *a, const int64_t *b, int64_t *out)
{
    for (int i = 0; i < P->N; ++i) out[i] = rzko_center(a[i] - b[i], P->q);
}

int rzko_poly_eq(const rzko_params *P, const int64_t *a, const int64_t *b)
{
    return memcmp(a, b, sizeof(int64_t) * (size_t)P->N) == 0;
}

#define POLY(M, cols, i, j) ((M) + ((size_t)(i) * (size_t)(cols) + (size_t)(j)) * (size_t)N)

/* mat.rs:95-115 -- (m x n) . (n x p); acc = acc + a*b with no skip for 0/1 blocks */
void rzko_mat_dot(const rzko_params *P, int m, int n, int p,
                  const int64_t *A, const int64_t *B, int64_t *out)
{
    const int N = P->N;
    int64_t *prod = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < p; ++j) {
            int64_t *o = POLY(out, p, i, j);
            memset(o, 0, sizeof(int64_t) * (size_t)N);        /* Polynomial::zero(), mat.rs:105 */
            for (int kk = 0; kk < n; ++kk) {                   /* mat.rs:108-111 */
                rzko_poly_mul(P, POLY(A, n, i, kk), POLY(B, p, kk, j), prod);
                rzko_poly_add(P, o, prod, o);
            }
        }
    free(prod);
}

/* mat.rs:122-140 */
void rzko_mat_add(const rzko_params *P, int m, int n, const int64_t *A, const int64_t *B, int64_t *out)
{
    const int N = P->N;
    for (int i = 0; i < m * n; ++i)
        rzko_poly_add(P, A + (size_t)i * N, B + (size_t)i * N, out + (size_t)i * N);
}

/* mat.rs:147-165 */
void rzko_mat_sub(const rzko_params *P, int m, int n, const int64_t *A, const int64_t *B, int64_t *out)
{
    const int N = P->N;
    for (int i = 0; i < m * n; ++i)
        rzko_poly_sub(P, A + (size_t)i * N, B + (size_t)i * N, out + (size_t)i * N);
}

/* mat.rs:168-178 */
void rzko_mat_cmul(const rzko_params *P, int m, int n, const int64_t *A, const int64_t *e, int64_t *out)
{
    const int N = P->N;
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
    for (int i = 0; i < m * n; ++i) {
        rzko_poly_mul(P, A + (size_t)i * N, e, tmp);
        memcpy(out + (size_t)i * N, tmp, sizeof(int64_t) * (size_t)N);
    }
    free(tmp);
}

/* params.rs:94-98: b * (11*kappa) * isqrt(k*N) */
uint64_t rzko_sigma(const rzko_params *P)
{
    return (uint64_t)P->b * (uint64_t)(11 * P->kappa) * isqrt_u64((uint64_t)P->k * (uint64_t)P->N);
}

/* params.rs:104 */
uint64_t rzko_commit_bound(const rzko_params *P)
{
    return 4 * rzko_sigma(P) * isqrt_u64((uint64_t)P->N);
}

/* params.rs:114 */
uint64_t rzko_verify_bound(const rzko_params *P)
{
    return 2 * rzko_sigma(P) * isqrt_u64((uint64_t)P->N);
}

/* polynomial.rs:60-73 */
uint64_t rzko_norm2(const rzko_params *P, const int64_t *poly)
{
    u128 s = 0;
    for (int i = 0; i < P->N; ++i) {
        i128 c = poly[i];
        s += (u128)(c * c);
    }
    return isqrt_u128(s);
}

static int check_constraint(const rzko_params *P, int rows, const int64_t *r, uint64_t bound)
{
    for (int i = 0; i < rows; ++i)
        if (rzko_norm2(P, r + (size_t)i * P->N) > bound) return 0;
    return 1;
}

/* params.rs:102-108 */
int rzko_check_commit_constraint(const rzko_params *P, int rows, const int64_t *r)
{
    return check_constraint(P, rows, r, rzko_commit_bound(P));
}

/* params.rs:112-118 */
int rzko_check_verify_constraint(const rzko_params *P, int rows, const int64_t *r)
{
    return check_constraint(P, rows, r, rzko_verify_bound(P));
}

/* commit.rs:33-60 */
void rzko_key_expand(const rzko_params *P, const int64_t *a1p, const int64_t *a2p,
                     int64_t *a1, int64_t *a2)
{
    const int N = P->N, n = P->n, k = P->k, l = P->l;
    memset(a1, 0, sizeof(int64_t) * (size_t)n * k * N);
    memset(a2, 0, sizeof(int64_t) * (size_t)l * k * N);
    for (int i = 0; i < n; ++i) {
        POLY(a1, k, i, i)[0] = 1;                                     /* diag(n, n, one) commit.rs:39 */
        for (int j = 0; j < k - n; ++j)                                /* extend_cols(a1') commit.rs:42 */
            memcpy(POLY(a1, k, i, n + j), POLY(a1p, k - n, i, j), sizeof(int64_t) * (size_t)N);
    }
    for (int i = 0; i < l; ++i) {
        POLY(a2, k, i, n + i)[0] = 1;                                 /* [0_{l x n} | I_l | a2'] commit.rs:50-55 */
        for (int j = 0; j < k - n - l; ++j)
            memcpy(POLY(a2, k, i, n + l + j), POLY(a2p, k - n - l, i, j), sizeof(int64_t) * (size_t)N);
    }
}

/* [a1; a2] . v + [0_n; x]   (commit.rs:109-125) */
static void key_dot_plus_x(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                           const int64_t *v, const int64_t *x, int64_t *c)
{
    const int N = P->N, n = P->n, k = P->k, l = P->l;
    int64_t *a = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + l) * k * N);
    memcpy(a, a1, sizeof(int64_t) * (size_t)n * k * N);                         /* commit.rs:111 */
    memcpy(a + (size_t)n * k * N, a2, sizeof(int64_t) * (size_t)l * k * N);     /* commit.rs:112 */
    int64_t *z = (int64_t *)calloc((size_t)(n + l) * N, sizeof(int64_t));       /* commit.rs:116-121 */
    memcpy(z + (size_t)n * N, x, sizeof(int64_t) * (size_t)l * N);
    int64_t *ar = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + l) * N);
    rzko_mat_dot(P, n + l, k, 1, a, v, ar);                                     /* commit.rs:125 */
    rzko_mat_add(P, n + l, 1, ar, z, c);
    free(a); free(z); free(ar);
}

/* commit.rs:88-128, r supplied by the caller */
int rzko_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                const int64_t *x, const int64_t *r, int64_t *c)
{
    int ok = rzko_check_commit_constraint(P, P->k, r);     /* commit.rs:102 */
    key_dot_plus_x(P, a1, a2, r, x, c);
    return ok;
}

/* commit.rs:173-210 */
int rzko_commitment_verify(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                           const int64_t *c, const int64_t *x, const int64_t *r,
                           const int64_t *f)
{
    const int N = P->N, n = P->n, l = P->l;
    if (!rzko_check_commit_constraint(P, P->k, r)) return 0;      /* commit.rs:182 */
    size_t sz = (size_t)(n + l) * N;
    int64_t *rhs = (int64_t *)malloc(sizeof(int64_t) * sz);
    int res;
    if (f) {                                                        /* commit.rs:203-207 */
        int64_t *fx = (int64_t *)malloc(sizeof(int64_t) * (size_t)l * N);
        int64_t *lhs = (int64_t *)malloc(sizeof(int64_t) * sz);
        rzko_mat_cmul(P, l, 1, x, f, fx);          /* z.cmul(f): zero rows stay zero */
        key_dot_plus_x(P, a1, a2, r, fx, rhs);
        rzko_mat_cmul(P, n + l, 1, c, f, lhs);
        res = memcmp(lhs, rhs, sizeof(int64_t) * sz) == 0;
        free(fx); free(lhs);
    } else {                                                        /* commit.rs:208 */
        key_dot_plus_x(P, a1, a2, r, x, rhs);
        res = memcmp(rhs, c, sizeof(int64_t) * sz) == 0;
    }
    free(rhs);
    return res;
}

/* ---------------- Open proof ---------------- */

/* open.rs:80-103 */
int rzko_open_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                     const int64_t *x, const int64_t *r, const int64_t *y,
                     int64_t *c, int64_t *t)
{
    int ok = rzko_commit(P, a1, a2, x, r, c);          /* open.rs:85 */
    rzko_mat_dot(P, P->n, P->k, 1, a1, y, t);          /* open.rs:97 */
    return ok;
}

/* z = y + r.cmul(d)  (open.rs:113-115) */
void rzko_open_respond(const rzko_params *P, const int64_t *y, const int64_t *r,
                       const int64_t *d, int64_t *z)
{
    const int N = P->N, k = P->k;
    int64_t *rd = (int64_t *)malloc(sizeof(int64_t) * (size_t)k * N);
    rzko_mat_cmul(P, k, 1, r, d, rd);
    rzko_mat_add(P, k, 1, y, rd, z);
    free(rd);
}

/* lhs = a1.z ; rhs = t + c1.cmul(d)  (open.rs:171-173, linear.rs:225-235, sum.rs:277-298) */
static int check_first_eq(const rzko_params *P, const int64_t *a1, const int64_t *z,
                          const int64_t *t, const int64_t *c1, const int64_t *d)
{
    const int N = P->N, n = P->n, l = P->l;
    assert(n == l);   /* Mat::add asserts equal dims (mat.rs:129); c1 has n+l-n = l rows (mat.rs:203-213) */
    int64_t *lhs = (int64_t *)malloc(sizeof(int64_t) * (size_t)n * N);
    int64_t *cd = (int64_t *)malloc(sizeof(int64_t) * (size_t)l * N);
    int64_t *rhs = (int64_t *)malloc(sizeof(int64_t) * (size_t)n * N);
    rzko_mat_dot(P, n, P->k, 1, a1, z, lhs);
    rzko_mat_cmul(P, l, 1, c1, d, cd);
    rzko_mat_add(P, n, 1, t, cd, rhs);
    int res = memcmp(lhs, rhs, sizeof(int64_t) * (size_t)n * N) == 0;
    free(lhs); free(cd); free(rhs);
    return res;
}

/* open.rs:162-174 */
int rzko_open_verify(const rzko_params *P, const int64_t *a1,
                     const int64_t *z, const int64_t *t, const int64_t *c1,
                     const int64_t *d)
{
    if (!rzko_check_verify_constraint(P, P->k, z)) return 0;     /* open.rs:167-169 */
    return check_first_eq(P, a1, z, t, c1, d);
}

/* ---------------- Linear proof ---------------- */

/* u-like term: a2.dot(v).cmul(g)   (linear.rs:124-128, sum.rs:157) */
static void a2_dot_cmul(const rzko_params *P, const int64_t *a2, const int64_t *v,
                        const int64_t *g, int64_t *out)
{
    const int N = P->N, l = P->l;
    int64_t *w = (int64_t *)malloc(sizeof(int64_t) * (size_t)l * N);
    rzko_mat_dot(P, l, P->k, 1, a2, v, w);
    rzko_mat_cmul(P, l, 1, w, g, out);
    free(w);
}

/* linear.rs:82-140 */
int rzko_linear_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                       const int64_t *g, const int64_t *x,
                       const int64_t *rp, const int64_t *r,
                       const int64_t *y, const int64_t *yp,
                       int64_t *gx, int64_t *cp, int64_t *c,
                       int64_t *t, int64_t *tp, int64_t *u)
{
    const int N = P->N, l = P->l;
    rzko_mat_cmul(P, l, 1, x, g, gx);                        /* linear.rs:91-95  xi.mul(g) */
    int ok = rzko_commit(P, a1, a2, gx, rp, cp);             /* linear.rs:96 */
    ok &= rzko_commit(P, a1, a2, x, r, c);                   /* linear.rs:97 */
    rzko_mat_dot(P, P->n, P->k, 1, a1, y, t);                /* linear.rs:118 */
    rzko_mat_dot(P, P->n, P->k, 1, a1, yp, tp);              /* linear.rs:121 */
    int64_t *gy = (int64_t *)malloc(sizeof(int64_t) * (size_t)l * N);
    int64_t *ayp = (int64_t *)malloc(sizeof(int64_t) * (size_t)l * N);
    a2_dot_cmul(P, a2, y, g, gy);                            /* linear.rs:124-128 */
    rzko_mat_dot(P, l, P->k, 1, a2, yp, ayp);                /* linear.rs:129 */
    rzko_mat_sub(P, l, 1, gy, ayp, u);
    free(gy); free(ayp);
    return ok;
}

/* linear.rs:144-158 */
void rzko_linear_respond(const rzko_params *P, const int64_t *y, const int64_t *yp,
                         const int64_t *r, const int64_t *rp, const int64_t *d,
                         int64_t *z, int64_t *zp)
{
    rzko_open_respond(P, y, r, d, z);       /* linear.rs:150-152 */
    rzko_open_respond(P, yp, rp, d, zp);    /* linear.rs:154-156 */
}

/* third equation: lhs = sum_i a2.z_i.cmul(g_i) - a2.zp ; rhs = (sum_i c2_i.cmul(g_i) - c2p).cmul(d) + u
 * (linear.rs:236-249 with T = 1, sum.rs:300-319) */
static int check_third_eq(const rzko_params *P, const int64_t *a2, int T,
                          const int64_t *zs, const int64_t *zp,
                          const int64_t *c2s, size_t c2_stride, const int64_t *c2p,
                          const int64_t *gs, const int64_t *u, const int64_t *d)
{
    const int N = P->N, l = P->l, n = P->n, k = P->k;
    assert(n == l);   /* c2 = last n rows (mat.rs:206); Mat::sub asserts equal dims (mat.rs:154) */
    size_t sz = (size_t)l * N;
    int64_t *acc = (int64_t *)malloc(sizeof(int64_t) * sz);
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * sz);
    int64_t *lhs = (int64_t *)malloc(sizeof(int64_t) * sz);
    int64_t *rhs = (int64_t *)malloc(sizeof(int64_t) * sz);
    for (int i = 0; i < T; ++i) {                         /* map + reduce(add): sum.rs:301-307 */
        a2_dot_cmul(P, a2, zs + (size_t)i * k * N, gs + (size_t)i * N, i == 0 ? acc : tmp);
        if (i) rzko_mat_add(P, l, 1, acc, tmp, acc);
    }
    rzko_mat_dot(P, l, k, 1, a2, zp, tmp);
    rzko_mat_sub(P, l, 1, acc, tmp, lhs);                 /* sum.rs:308 */
    for (int i = 0; i < T; ++i) {                         /* sum.rs:309-315 */
        rzko_mat_cmul(P, n, 1, c2s + (size_t)i * c2_stride, gs + (size_t)i * N, i == 0 ? acc : tmp);
        if (i) rzko_mat_add(P, l, 1, acc, tmp, acc);
    }
    rzko_mat_sub(P, l, 1, acc, c2p, tmp);                 /* sum.rs:316 */
    rzko_mat_cmul(P, l, 1, tmp, d, acc);                  /* sum.rs:317 */
    rzko_mat_add(P, l, 1, acc, u, rhs);                   /* sum.rs:318 */
    int res = memcmp(lhs, rhs, sizeof(int64_t) * sz) == 0;
    free(acc); free(tmp); free(lhs); free(rhs);
    return res;
}

/* linear.rs:213-250.  c, cp are full commitments [(n+l)][N]; c1 = first l rows, c2 = last n rows. */
int rzko_linear_verify(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                       const int64_t *z, const int64_t *zp,
                       const int64_t *c, const int64_t *cp, const int64_t *g,
                       const int64_t *t, const int64_t *tp, const int64_t *u,
                       const int64_t *d)
{
    const int N = P->N, l = P->l;
    if (!rzko_check_verify_constraint(P, P->k, z)) return 0;      /* linear.rs:218 */
    if (!rzko_check_verify_constraint(P, P->k, zp)) return 0;     /* linear.rs:221 */
    if (!check_first_eq(P, a1, z, t, c, d)) return 0;             /* linear.rs:225-229 */
    if (!check_first_eq(P, a1, zp, tp, cp, d)) return 0;          /* linear.rs:231-235 */
    return check_third_eq(P, a2, 1, z, zp, c + (size_t)l * N, 0, cp + (size_t)l * N, g, u, d);
}

/* ---------------- Sum proof ---------------- */

/* sum.rs:99-178 */
int rzko_sum_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2, int T,
                    const int64_t *gs, const int64_t *xs,
                    const int64_t *rp, const int64_t *rs,
                    const int64_t *ys, const int64_t *yp,
                    int64_t *xp, int64_t *cp, int64_t *cs,
                    int64_t *ts, int64_t *tp, int64_t *u)
{
    const int N = P->N, n = P->n, k = P->k, l = P->l;
    assert(T > 0);                                               /* sum.rs:105 */
    size_t lsz = (size_t)l * N;
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * lsz);
    for (int i = 0; i < T; ++i) {                                /* sum.rs:107-115 */
        rzko_mat_cmul(P, l, 1, xs + (size_t)i * lsz, gs + (size_t)i * N, i == 0 ? xp : tmp);
        if (i) rzko_mat_add(P, l, 1, xp, tmp, xp);
    }
    int ok = rzko_commit(P, a1, a2, xp, rp, cp);                 /* sum.rs:116 */
    for (int i = 0; i < T; ++i)                                  /* sum.rs:117-120 */
        ok &= rzko_commit(P, a1, a2, xs + (size_t)i * lsz, rs + (size_t)i * k * N,
                          cs + (size_t)i * (n + l) * N);
    for (int i = 0; i < T; ++i)                                  /* sum.rs:145-148 */
        rzko_mat_dot(P, n, k, 1, a1, ys + (size_t)i * k * N, ts + (size_t)i * n * N);
    rzko_mat_dot(P, n, k, 1, a1, yp, tp);                        /* sum.rs:151 */
    int64_t *acc = (int64_t *)malloc(sizeof(int64_t) * lsz);
    for (int i = 0; i < T; ++i) {                                /* sum.rs:154-159 */
        a2_dot_cmul(P, a2, ys + (size_t)i * k * N, gs + (size_t)i * N, i == 0 ? acc : tmp);
        if (i) rzko_mat_add(P, l, 1, acc, tmp, acc);
    }
    rzko_mat_dot(P, l, k, 1, a2, yp, tmp);                       /* sum.rs:160 */
    rzko_mat_sub(P, l, 1, acc, tmp, u);
    free(acc); free(tmp);
    return ok;
}

/* sum.rs:182-200 */
void rzko_sum_respond(const rzko_params *P, int T, const int64_t *ys, const int64_t *yp,
                      const int64_t *rs, const int64_t *rp, const int64_t *d,
                      int64_t *zs, int64_t *zp)
{
    size_t ksz = (size_t)P->k * P->N;
    for (int i = 0; i < T; ++i)
        rzko_open_respond(P, ys + i * ksz, rs + i * ksz, d, zs + i * ksz);
    rzko_open_respond(P, yp, rp, d, zp);
}

/* sum.rs:257-320.  cs [T][(n+l)][N]. */
int rzko_sum_verify(const rzko_params *P, const int64_t *a1, const int64_t *a2, int T,
                    const int64_t *zs, const int64_t *zp,
                    const int64_t *cs, const int64_t *cp, const int64_t *gs,
                    const int64_t *ts, const int64_t *tp, const int64_t *u,
                    const int64_t *d)
{
    const int N = P->N, n = P->n, k = P->k, l = P->l;
    size_t ksz = (size_t)k * N, csz = (size_t)(n + l) * N;
    for (int i = 0; i < T; ++i)                                              /* sum.rs:262-268 */
        if (!rzko_check_verify_constraint(P, k, zs + i * ksz)) return 0;
    if (!rzko_check_verify_constraint(P, k, zp)) return 0;                   /* sum.rs:269 */
    /* sum.rs:278-291: Vec equality is evaluated after all lhs/rhs are built */
    int all = 1;
    for (int i = 0; i < T; ++i)
        all &= check_first_eq(P, a1, zs + i * ksz, ts + (size_t)i * n * N, cs + i * csz, d);
    if (!all) return 0;
    if (!check_first_eq(P, a1, zp, tp, cp, d)) return 0;                     /* sum.rs:294-298 */
    return check_third_eq(P, a2, T, zs, zp, cs + (size_t)l * N, csz, cp + (size_t)l * N, gs, u, d);
}

/* ---------------- batch drivers ---------------- */

#ifdef _OPENMP
#define OMP_FOR(nt) _Pragma("omp parallel for schedule(dynamic, 1) num_threads(nt)")
#else
#define OMP_FOR(nt)
#endif

static int nthr(int nthreads)
{
    int m = rzko_max_threads();
    if (nthreads <= 0 || nthreads > m) return m;
    return nthreads;
}

void rzko_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                       const int64_t *x, const int64_t *r, int64_t *c, uint8_t *ok, int nthreads)
{
    const size_t N = P->N, xs = P->l * N, rs = P->k * N, cs = (P->n + P->l) * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        ok[i] = (uint8_t)rzko_commit(P, a1, a2, x + i * xs, r + i * rs, c + i * cs);
}

void rzko_open_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                            const int64_t *x, const int64_t *r, const int64_t *y,
                            int64_t *c, int64_t *t, uint8_t *ok, int nthreads)
{
    const size_t N = P->N, xs = P->l * N, ks = P->k * N, cs = (P->n + P->l) * N, ns = P->n * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        ok[i] = (uint8_t)rzko_open_commit(P, a1, a2, x + i * xs, r + i * ks, y + i * ks,
                                          c + i * cs, t + i * ns);
}

void rzko_open_respond_batch(const rzko_params *P, size_t B, const int64_t *y, const int64_t *r,
                             const int64_t *d, int64_t *z, int nthreads)
{
    const size_t N = P->N, ks = P->k * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        rzko_open_respond(P, y + i * ks, r + i * ks, d + i * N, z + i * ks);
}

void rzko_open_verify_batch(const rzko_params *P, const int64_t *a1, size_t B,
                            const int64_t *z, const int64_t *t, const int64_t *c1,
                            const int64_t *d, uint8_t *ok, int nthreads)
{
    const size_t N = P->N, ks = P->k * N, ns = P->n * N, ls = P->l * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        ok[i] = (uint8_t)rzko_open_verify(P, a1, z + i * ks, t + i * ns, c1 + i * ls, d + i * N);
}

void rzko_linear_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                              const int64_t *g, const int64_t *x,
                              const int64_t *rp, const int64_t *r,
                              const int64_t *y, const int64_t *yp,
                              int64_t *gx, int64_t *cp, int64_t *c,
                              int64_t *t, int64_t *tp, int64_t *u, uint8_t *ok, int nthreads)
{
    const size_t N = P->N, ls = P->l * N, ks = P->k * N, cs = (P->n + P->l) * N, ns = P->n * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        ok[i] = (uint8_t)rzko_linear_commit(P, a1, a2, g + i * N, x + i * ls, rp + i * ks, r + i * ks,
                                            y + i * ks, yp + i * ks, gx + i * ls, cp + i * cs,
                                            c + i * cs, t + i * ns, tp + i * ns, u + i * ls);
}

void rzko_linear_respond_batch(const rzko_params *P, size_t B, const int64_t *y, const int64_t *yp,
                               const int64_t *r, const int64_t *rp, const int64_t *d,
                               int64_t *z, int64_t *zp, int nthreads)
{
    const size_t N = P->N, ks = P->k * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        rzko_linear_respond(P, y + i * ks, yp + i * ks, r + i * ks, rp + i * ks, d + i * N,
                            z + i * ks, zp + i * ks);
}

void rzko_linear_verify_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                              const int64_t *z, const int64_t *zp,
                              const int64_t *c, const int64_t *cp, const int64_t *g,
                              const int64_t *t, const int64_t *tp, const int64_t *u,
                              const int64_t *d, uint8_t *ok, int nthreads)
{
    const size_t N = P->N, ls = P->l * N, ks = P->k * N, cs = (P->n + P->l) * N, ns = P->n * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        ok[i] = (uint8_t)rzko_linear_verify(P, a1, a2, z + i * ks, zp + i * ks, c + i * cs, cp + i * cs,
                                            g + i * N, t + i * ns, tp + i * ns, u + i * ls, d + i * N);
}

void rzko_sum_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B, int T,
                           const int64_t *gs, const int64_t *xs,
                           const int64_t *rp, const int64_t *rs,
                           const int64_t *ys, const int64_t *yp,
                           int64_t *xp, int64_t *cp, int64_t *cs,
                           int64_t *ts, int64_t *tp, int64_t *u, uint8_t *ok, int nthreads)
{
    const size_t N = P->N, ls = P->l * N, ks = P->k * N, csz = (P->n + P->l) * N, ns = P->n * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        ok[i] = (uint8_t)rzko_sum_commit(P, a1, a2, T, gs + i * T * N, xs + i * T * ls,
                                         rp + i * ks, rs + i * T * ks, ys + i * T * ks, yp + i * ks,
                                         xp + i * ls, cp + i * csz, cs + i * T * csz,
                                         ts + i * T * ns, tp + i * ns, u + i * ls);
}

void rzko_sum_respond_batch(const rzko_params *P, size_t B, int T, const int64_t *ys, const int64_t *yp,
                            const int64_t *rs, const int64_t *rp, const int64_t *d,
                            int64_t *zs, int64_t *zp, int nthreads)
{
    const size_t N = P->N, ks = P->k * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        rzko_sum_respond(P, T, ys + i * T * ks, yp + i * ks, rs + i * T * ks, rp + i * ks, d + i * N,
                         zs + i * T * ks, zp + i * ks);
}

void rzko_sum_verify_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B, int T,
                           const int64_t *zs, const int64_t *zp,
                           const int64_t *cs, const int64_t *cp, const int64_t *gs,
                           const int64_t *ts, const int64_t *tp, const int64_t *u,
                           const int64_t *d, uint8_t *ok, int nthreads)
{
    const size_t N = P->N, ls = P->l * N, ks = P->k * N, csz = (P->n + P->l) * N, ns = P->n * N;
    int nt = nthr(nthreads);
    (void)nt;
    OMP_FOR(nt)
    for (long long i = 0; i < (long long)B; ++i)
        ok[i] = (uint8_t)rzko_sum_verify(P, a1, a2, T, zs + i * T * ks, zp + i * ks,
                                         cs + i * T * csz, cp + i * csz, gs + i * T * N,
                                         ts + i * T * ns, tp + i * ns, u + i * ls, d + i * N);
}
