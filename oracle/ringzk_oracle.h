/*
 * ringzk_oracle.h -- CPU restatement of AlvinHon/ring-zk's R_q hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ring-zk_b200/,
 * include/) may include, link or call this file.  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker or the timed CPU arm.
 *
 * PARITY STATUS: "parity unpinned" for ring products.  The ring arithmetic of
 * the reference lives in the third-party crate poly-ring-xnp1 = "0.3"
 * (/root/reference/Cargo.toml:18) whose source is not under /root/reference,
 * and the reference holds no golden vectors for products on ZqI64.  What IS
 * pinned from the reference's own tests (tests/test_oracle_pins.py):
 *   sigma(1024) = 21780            params.rs:144-150
 *   norm_2([1,-2,3,-4]) = 5        polynomial.rs:111-115
 *   Mat dot/add/sub/componentwise composition  mat.rs:243-406
 *   honest Open/Linear/Sum transcripts verify  tests/test.rs:11-93
 *   swapped openings fail          commit.rs:165-170
 * The product itself is the mathematically unique negacyclic convolution in
 * Z_q[X]/(X^N+1); the representative is the canonical centred residue in
 * [-(q-1)/2, (q-1)/2] (SURVEY.md section 8c derives this from observable
 * behaviour: polynomial.rs:22-23, commit.rs:100-105, open.rs:171-173,
 * params.rs:123-126).
 *
 * Data layout: polynomials are int64_t[N], coefficient i at index i, already
 * canonical centred unless noted.  A matrix is row-major [rows][cols][N].
 */
#ifndef RINGZK_ORACLE_H
#define RINGZK_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Params<I> (params.rs:18-36) plus the const generic N and the modulus Q of
 * ZqI64<Q> (params.rs:121). */
typedef struct {
    int64_t q;     /* modulus Q of ZqI64<Q>; Params.q field is q/2 */
    int64_t b;     /* params.b */
    int32_t N;     /* ring degree (const generic N) */
    int32_t n, k, l;
    int32_t kappa;
} rzko_params;

/* Params::default() (params.rs:121-138) with ring degree N. */
rzko_params rzko_default_params(int32_t N);

/* canonical centred residue of v mod q */
int64_t rzko_center(int64_t v, int64_t q);

/* Polynomial ops (crate poly-ring-xnp1; call sites mat.rs:109-110,135-136,160-161,176). */
void rzko_poly_mul(const rzko_params *P, const int64_t *a, const int64_t *b, int64_t *out);
void rzko_poly_add(const rzko_params *P, const int64_t *a, const int64_t *b, int64_t *out);
void rzko_poly_sub(const rzko_params *P, const int64_t *a, const int64_t *b, int64_t *out);
int  rzko_poly_eq(const rzko_params *P, const int64_t *a, const int64_t *b);

/* Mat ops (mat.rs:95-178).  All matrices row-major [rows][cols][N]. */
void rzko_mat_dot(const rzko_params *P, int m, int n, int p,
                  const int64_t *A, const int64_t *B, int64_t *out);
void rzko_mat_add(const rzko_params *P, int m, int n, const int64_t *A, const int64_t *B, int64_t *out);
void rzko_mat_sub(const rzko_params *P, int m, int n, const int64_t *A, const int64_t *B, int64_t *out);
void rzko_mat_cmul(const rzko_params *P, int m, int n, const int64_t *A, const int64_t *e, int64_t *out);

/* params.rs:94-98 sigma; params.rs:102-118 bounds 4*sigma*isqrt(N), 2*sigma*isqrt(N) */
uint64_t rzko_sigma(const rzko_params *P);
uint64_t rzko_commit_bound(const rzko_params *P);
uint64_t rzko_verify_bound(const rzko_params *P);
/* polynomial.rs:60-73: floor(sqrt(sum c_i^2)) */
uint64_t rzko_norm2(const rzko_params *P, const int64_t *poly);
/* params.rs:102-108 / 112-118 applied to a (rows x 1) matrix */
int rzko_check_commit_constraint(const rzko_params *P, int rows, const int64_t *r);
int rzko_check_verify_constraint(const rzko_params *P, int rows, const int64_t *r);

/* commit.rs:33-60: expand the random blocks a1p [n][k-n][N], a2p [l][k-n-l][N]
 * into a1 = [I_n | a1p] ([n][k][N]) and a2 = [0 | I_l | a2p] ([l][k][N]). */
void rzko_key_expand(const rzko_params *P, const int64_t *a1p, const int64_t *a2p,
                     int64_t *a1, int64_t *a2);

/* commit.rs:88-128 with r supplied.  x [l][N], r [k][N], c [(n+l)][N].
 * Returns check_commit_constraint(r) (the reference would redraw on 0). */
int rzko_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                const int64_t *x, const int64_t *r, int64_t *c);
/* commit.rs:173-210; f may be NULL (None). */
int rzko_commitment_verify(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                           const int64_t *c, const int64_t *x, const int64_t *r,
                           const int64_t *f);

/* Open proof: open.rs:80-103, 107-117, 162-174.
 * c1 is the first (n+l-n) rows of c as commit.rs:213-218 / mat.rs:203-213 split it. */
int  rzko_open_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                      const int64_t *x, const int64_t *r, const int64_t *y,
                      int64_t *c, int64_t *t);
void rzko_open_respond(const rzko_params *P, const int64_t *y, const int64_t *r,
                       const int64_t *d, int64_t *z);
int  rzko_open_verify(const rzko_params *P, const int64_t *a1,
                      const int64_t *z, const int64_t *t, const int64_t *c1,
                      const int64_t *d);

/* Linear proof: linear.rs:82-140, 144-158, 213-250.
 * Outputs: gx [l][N] (= opening_p.x), cp, c [(n+l)][N], t, tp [n][N], u [l][N]. */
int  rzko_linear_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                        const int64_t *g, const int64_t *x,
                        const int64_t *rp, const int64_t *r,
                        const int64_t *y, const int64_t *yp,
                        int64_t *gx, int64_t *cp, int64_t *c,
                        int64_t *t, int64_t *tp, int64_t *u);
void rzko_linear_respond(const rzko_params *P, const int64_t *y, const int64_t *yp,
                         const int64_t *r, const int64_t *rp, const int64_t *d,
                         int64_t *z, int64_t *zp);
int  rzko_linear_verify(const rzko_params *P, const int64_t *a1, const int64_t *a2,
                        const int64_t *z, const int64_t *zp,
                        const int64_t *c, const int64_t *cp, const int64_t *g,
                        const int64_t *t, const int64_t *tp, const int64_t *u,
                        const int64_t *d);

/* Sum proof: sum.rs:99-178, 182-200, 257-320.  T terms.
 * gs [T][N], xs [T][l][N], rs [T][k][N], ys [T][k][N]; rp, yp [k][N].
 * Outputs: xp [l][N], cp [(n+l)][N], cs [T][(n+l)][N], ts [T][n][N], tp [n][N], u [l][N]. */
int  rzko_sum_commit(const rzko_params *P, const int64_t *a1, const int64_t *a2, int T,
                     const int64_t *gs, const int64_t *xs,
                     const int64_t *rp, const int64_t *rs,
                     const int64_t *ys, const int64_t *yp,
                     int64_t *xp, int64_t *cp, int64_t *cs,
                     int64_t *ts, int64_t *tp, int64_t *u);
void rzko_sum_respond(const rzko_params *P, int T, const int64_t *ys, const int64_t *yp,
                      const int64_t *rs, const int64_t *rp, const int64_t *d,
                      int64_t *zs, int64_t *zp);
int  rzko_sum_verify(const rzko_params *P, const int64_t *a1, const int64_t *a2, int T,
                     const int64_t *zs, const int64_t *zp,
                     const int64_t *cs, const int64_t *cp, const int64_t *gs,
                     const int64_t *ts, const int64_t *tp, const int64_t *u,
                     const int64_t *d);

/* ---- batch drivers (OpenMP over items) used by tests and the CPU baseline ----
 * Arrays are [B][...per item...]; nthreads <= 0 means omp default.
 * "ok" arrays are one byte per item. */
void rzko_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                       const int64_t *x, const int64_t *r, int64_t *c, uint8_t *ok, int nthreads);
void rzko_open_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                            const int64_t *x, const int64_t *r, const int64_t *y,
                            int64_t *c, int64_t *t, uint8_t *ok, int nthreads);
void rzko_open_respond_batch(const rzko_params *P, size_t B, const int64_t *y, const int64_t *r,
                             const int64_t *d, int64_t *z, int nthreads);
void rzko_open_verify_batch(const rzko_params *P, const int64_t *a1, size_t B,
                            const int64_t *z, const int64_t *t, const int64_t *c1,
                            const int64_t *d, uint8_t *ok, int nthreads);
void rzko_linear_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                              const int64_t *g, const int64_t *x,
                              const int64_t *rp, const int64_t *r,
                              const int64_t *y, const int64_t *yp,
                              int64_t *gx, int64_t *cp, int64_t *c,
                              int64_t *t, int64_t *tp, int64_t *u, uint8_t *ok, int nthreads);
void rzko_linear_respond_batch(const rzko_params *P, size_t B, const int64_t *y, const int64_t *yp,
                               const int64_t *r, const int64_t *rp, const int64_t *d,
                               int64_t *z, int64_t *zp, int nthreads);
void rzko_linear_verify_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B,
                              const int64_t *z, const int64_t *zp,
                              const int64_t *c, const int64_t *cp, const int64_t *g,
                              const int64_t *t, const int64_t *tp, const int64_t *u,
                              const int64_t *d, uint8_t *ok, int nthreads);
void rzko_sum_commit_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B, int T,
                           const int64_t *gs, const int64_t *xs,
                           const int64_t *rp, const int64_t *rs,
                           const int64_t *ys, const int64_t *yp,
                           int64_t *xp, int64_t *cp, int64_t *cs,
                           int64_t *ts, int64_t *tp, int64_t *u, uint8_t *ok, int nthreads);
void rzko_sum_respond_batch(const rzko_params *P, size_t B, int T, const int64_t *ys, const int64_t *yp,
                            const int64_t *rs, const int64_t *rp, const int64_t *d,
                            int64_t *zs, int64_t *zp, int nthreads);
void rzko_sum_verify_batch(const rzko_params *P, const int64_t *a1, const int64_t *a2, size_t B, int T,
                           const int64_t *zs, const int64_t *zp,
                           const int64_t *cs, const int64_t *cp, const int64_t *gs,
                           const int64_t *ts, const int64_t *tp, const int64_t *u,
                           const int64_t *d, uint8_t *ok, int nthreads);

int rzko_max_threads(void);
/* number of ring products executed since the last reset (single-thread use only) */
uint64_t rzko_product_count(int reset);

#ifdef __cplusplus
}
#endif
#endif
