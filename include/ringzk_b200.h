/*
 * ringzk_b200.h -- C ABI of the B200-native batched engine for ring-zk's R_q hot path.
 *
 * This is the drop-in boundary.  The reference (AlvinHon/ring-zk, pure Rust) has no FFI;
 * the entry points below are what a thin Rust shim binds so that the reference's public
 * surface (lib.rs:5-24) keeps working, with `*_batch` methods added alongside.  Each entry
 * point cites the reference function it replaces (paths relative to /root/reference/src).
 * INTEGRATION.md shows the Rust-side `extern "C"` block and the wrappers.
 *
 * Conventions
 *   - Every call returns an int status (RZK_OK == 0) and never unwinds; rzk_last_error()
 *     gives the message.  The Rust shim turns RZK_ERR_INVALID into the reference's
 *     assert!/panic behaviour (params.rs:71, commit.rs:95, sum.rs:105).
 *   - Polynomials are arrays of N coefficients, coefficient i at index i, zero padded,
 *     batch-major row-major: [B][polys per item][N].
 *   - Coefficients are the canonical centred residues mod q in [-(q-1)/2, (q-1)/2]
 *     (what ZqI64<Q> -> i64 yields).  Any i32/i64 representative is accepted on input
 *     and canonicalised.  Compact types: int32_t for full-size values, int8_t for the
 *     small randomness r in [-b, b] and the challenge d in {-1, 0, 1}.
 *   - Shapes are the reference's default (n, k, l) = (1, 3, 1), N = 512,
 *     q = 3515337053 (params.rs:121-138), with b * kappa <= 74 (Params::default(): 1 * 36): rzk_create derives the
 *     exactness limits of the two-prime products from the actual (b, kappa) -- the masking vectors N(0, sigma) must sit
 *     10 sigma inside rzk_small_limit(), and A1.z of a response at the norm bound inside the CRT range -- and returns
 *     RZK_ERR_UNSUPPORTED for every other set (the shim keeps the reference's CPU path for those).
 *   - "ok"/verify results are bitmaps: bit (i & 7) of byte (i >> 3) is item i.
 *   - Host entry points take host pointers (pinned memory recommended: rzk_host_alloc)
 *     and pipeline H2D / kernels / D2H over chunks.  `_dev` entry points take device
 *     pointers, enqueue on the given cudaStream_t (as void*) and do not synchronise.
 *     Device arrays must be 16-byte aligned (cudaMalloc and framework allocations are; every row of a batch-major array
 *     then is, since a polynomial is 2048 or 512 bytes): the kernels read rows with 128-bit loads.  Host pointers of the
 *     host entry points need no particular alignment.
 *   - One engine is bound to one CUDA device and is externally synchronised.
 *   - The randomness r, y, d is drawn by the caller (host side, seeded RNG) and passed in; the
 *     protocol entry points sample nothing.  (Optional, separate: rzk_sample_*_dev, below.)
 *   - Environment variables read by rzk_create: RZK_CHUNK_ITEMS (items per chunk of the host pipeline, default 8192);
 *     RZK_TEST_LOWERING (comma-separated: generic, nosparse, norot, nodimg, nofuse, nosegments -- alternative lowerings of
 *     the same phases for the differential tests, results are identical in every setting); RZK_TUNE (developer A/B timing).
 *   - There is no CPU fallback: without a CUDA device rzk_create fails with RZK_ERR_CUDA.
 */
#ifndef RINGZK_B200_H
#define RINGZK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RZK_OK 0
#define RZK_ERR_INVALID 1       /* null pointer / bad shape / T == 0 */
#define RZK_ERR_UNSUPPORTED 2   /* parameter set or key structure outside the accelerated instantiation */
#define RZK_ERR_CUDA 3          /* CUDA runtime failure (message in rzk_last_error) */
#define RZK_ERR_RANGE 4         /* a masking vector y exceeded the exactness bound rzk_small_limit(): the outputs of the call
                                   that depend on it are not exact (honest samples never get there: the limit is >= 10 sigma) */
#define RZK_ERR_NOKEY 5         /* rzk_set_key has not been called */

typedef struct rzk_engine rzk_engine;

/* Params<ZqI64<Q>> (params.rs:18-36) + const generic N.  q is the modulus Q. */
typedef struct {
    int64_t q;
    int64_t b;
    int32_t N, n, k, l;
    int32_t kappa;
} rzk_params;

/* Params::default() (params.rs:121-138) at ring degree N. */
rzk_params rzk_default_params(int32_t N);

/* Engine lifetime.  device < 0 selects the current CUDA device. */
int rzk_create(const rzk_params *params, int device, rzk_engine **out);
void rzk_destroy(rzk_engine *e);
const char *rzk_last_error(const rzk_engine *e);   /* e may be NULL: last creation error */
int rzk_device(const rzk_engine *e);

/* params.rs:94-98 sigma, params.rs:104 / 114 norm bounds, and the |y| bound under which the
 * two-prime products are exact (violations are reported as RZK_ERR_RANGE, never silently). */
uint64_t rzk_sigma(const rzk_engine *e);
uint64_t rzk_commit_bound(const rzk_engine *e);
uint64_t rzk_verify_bound(const rzk_engine *e);
uint32_t rzk_small_limit(const rzk_engine *e);

/* CommitmentKey (commit.rs:19-60): a1 [n][k][N], a2 [l][k][N] as the reference stores them.
 * The identity / zero blocks are verified (else RZK_ERR_UNSUPPORTED); the random blocks are
 * transformed once and stay resident on the device in NTT form. */
int rzk_set_key(rzk_engine *e, const int64_t *a1, const int64_t *a2);

/* Pinned host memory helpers for the host entry points. */
void *rzk_host_alloc(size_t bytes);
void rzk_host_free(void *p);
/* Blocks until everything enqueued by `_dev` calls on `stream` has finished. */
int rzk_sync(rzk_engine *e, void *stream);

/* ---------------------------------------------------------------- commitment
 * CommitmentKey::commit (commit.rs:88-128) with r supplied:  c = [a1;a2].r + [0;x].
 *   x [B][1][N] i32, r [B][3][N] i8, c [B][2][N] i32,
 *   ok bitmap: check_commit_constraint(r) (commit.rs:102; the reference redraws r on 0).
 * Exact for ANY int8 r: the fast program covers |r| <= 15 (one word per product), and the items outside that range
 * are redone by the two-prime program in a masked launch on the same stream (no second pass over the batch).  Engines
 * created with b > 15 run the two-prime program for every item. */
int rzk_commit_batch(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r,
                     int32_t *c, uint8_t *ok_bitmap);

/* The same commitment with the randomness packed at 2 bits per coefficient -- what Params::default() (b = 1, r in {-1, 0, 1},
 * commit.rs:96-101 via polynomial.rs:14-23) needs: 384 bytes per commitment cross the bus instead of 1536.
 *   r2 [B][3][N/4] bytes: coefficient i of a row is the two's-complement field (0, 1, -2 -> 2, -1 -> 3) in bits
 *   2(i & 3) .. 2(i & 3) + 1 of byte i >> 2.  Unpacked on the device (rzk_unpack_r2_dev), then exactly rzk_commit_batch.
 * rzk_pack_r2 is the host-side packer (no engine, plain CPU loop; RZK_ERR_RANGE for an entry outside [-2, 1],
 * count = number of coefficients, a multiple of 4); rzk_unpack_r2_dev expands count coefficients (a multiple of 16) between
 * device arrays on `stream`. */
int rzk_commit_batch_r2(rzk_engine *e, size_t B, const int32_t *x, const uint8_t *r2,
                        int32_t *c, uint8_t *ok_bitmap);
int rzk_pack_r2(size_t count, const int8_t *r, uint8_t *r2);
int rzk_unpack_r2_dev(rzk_engine *e, size_t count, const uint8_t *r2, int8_t *r, void *stream);

/* Commitment::verify (commit.rs:173-210): check_commit_constraint(r), then
 *   f == NULL (Opening.f = None):  [a1;a2].r + [0;x] == c
 *   f != NULL (Some(f)):           f*c == [a1;a2].r + f*[0;x]          f [B][N] i8 (challenge-space polynomial)
 *   c [B][2][N] i32, x [B][1][N] i32, r [B][3][N] i8; bit i of the bitmap = the reference's bool. */
int rzk_commitment_verify_batch(rzk_engine *e, size_t B, const int32_t *c, const int32_t *x,
                                const int8_t *r, const int8_t *f, uint8_t *verify_bitmap);

/* ---------------------------------------------------------------- Open proof
 * OpenProofProver::commit (open.rs:80-103): commitment c plus t = A1.y.   y [B][3][N] i32, t [B][1][N] */
int rzk_open_commit_batch(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r,
                          const int32_t *y, int32_t *c, int32_t *t, uint8_t *ok_bitmap);
/* OpenProofProver::create_response (open.rs:107-117): z = y + d*r.   d [B][N] i8, z [B][3][N] */
int rzk_open_respond_batch(rzk_engine *e, size_t B, const int32_t *y, const int8_t *r,
                           const int8_t *d, int32_t *z);
/* OpenProofVerifier::verify (open.rs:162-174): norm check on z, then A1.z == t + c1*d.
 *   c1 [B][1][N] (first rows of c, commit.rs:213-218). */
int rzk_open_verify_batch(rzk_engine *e, size_t B, const int32_t *z, const int32_t *t,
                          const int32_t *c1, const int8_t *d, uint8_t *verify_bitmap);

/* ---------------------------------------------------------------- Linear proof
 * LinearProofProver::commit (linear.rs:82-140).  g, x [B][N]; rp, r [B][3][N] i8; y, yp [B][3][N];
 * outputs gx (= opening_p.x), cp, c [B][2][N], t, tp, u [B][N]. */
int rzk_linear_commit_batch(rzk_engine *e, size_t B, const int32_t *g, const int32_t *x,
                            const int8_t *rp, const int8_t *r, const int32_t *y, const int32_t *yp,
                            int32_t *gx, int32_t *cp, int32_t *c, int32_t *t, int32_t *tp, int32_t *u,
                            uint8_t *ok_bitmap);
/* LinearProofProver::create_response (linear.rs:144-158) */
int rzk_linear_respond_batch(rzk_engine *e, size_t B, const int32_t *y, const int32_t *yp,
                             const int8_t *r, const int8_t *rp, const int8_t *d,
                             int32_t *z, int32_t *zp);
/* LinearProofVerifier::verify (linear.rs:213-250).  c, cp are the full commitments [B][2][N]. */
int rzk_linear_verify_batch(rzk_engine *e, size_t B, const int32_t *z, const int32_t *zp,
                            const int32_t *c, const int32_t *cp, const int32_t *g,
                            const int32_t *t, const int32_t *tp, const int32_t *u,
                            const int8_t *d, uint8_t *verify_bitmap);

/* ---------------------------------------------------------------- Sum proof, T terms
 * SumProofProver::commit (sum.rs:99-178).  gs, xs [B][T][N]; rs, ys [B][T][3][N]; rp, yp [B][3][N];
 * outputs xp [B][N], cp [B][2][N], cs [B][T][2][N], ts [B][T][N], tp, u [B][N]. */
int rzk_sum_commit_batch(rzk_engine *e, size_t B, uint32_t T, const int32_t *gs, const int32_t *xs,
                         const int8_t *rp, const int8_t *rs, const int32_t *ys, const int32_t *yp,
                         int32_t *xp, int32_t *cp, int32_t *cs, int32_t *ts, int32_t *tp, int32_t *u,
                         uint8_t *ok_bitmap);
/* SumProofProver::create_response (sum.rs:182-200) */
int rzk_sum_respond_batch(rzk_engine *e, size_t B, uint32_t T, const int32_t *ys, const int32_t *yp,
                          const int8_t *rs, const int8_t *rp, const int8_t *d,
                          int32_t *zs, int32_t *zp);
/* SumProofVerifier::verify (sum.rs:257-320) */
int rzk_sum_verify_batch(rzk_engine *e, size_t B, uint32_t T, const int32_t *zs, const int32_t *zp,
                         const int32_t *cs, const int32_t *cp, const int32_t *gs,
                         const int32_t *ts, const int32_t *tp, const int32_t *u,
                         const int8_t *d, uint8_t *verify_bitmap);

/* ---------------------------------------------------------------- device-resident variants
 * Same semantics; every pointer is a device pointer on the engine's device.  `flags` is one
 * uint32_t per item and is OR-ed into, so the caller zeroes it:
 *   bit 0 (check failed): the commit / verify constraint or a verification equation does not hold -- the reference's `false`;
 *   bit 1 (range error): an operand left the range in which the item's products are exact, and the item's OUTPUTS ARE NOT
 *     VALID: a masking coefficient |y| > rzk_small_limit(), or -- `_dev` commit entry points of an engine with b <= 15 only --
 *     a randomness coefficient |r| > 15, which is outside the reference's own contract |r| <= b (polynomial.rs:14-24; the
 *     host entry points redo such items exactly, the `_dev` ones report them).  A caller must test bit 1, per item or
 *     through the range_any word of rzk_flags_to_bitmap_dev: the bitmap alone carries bit 0 only.
 * rzk_flags_to_bitmap_dev packs "(flags & 1) == 0" into a bitmap.  c_stride is the number of
 * polynomials per item in the commitment array handed to verify (1: c1 only, 2: full c).
 * The Linear / Sum variants and the responses keep intermediates (w = A2.y, the residue stash of the
 * three-prime products, the hand-over words of the response kernels) in scratch owned by the engine:
 * enqueue the `_dev` calls of one engine on ONE stream at a time (use one engine per stream otherwise). */
int rzk_commit_batch_dev(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r,
                         int32_t *c, uint32_t *flags, void *stream);
int rzk_commitment_verify_batch_dev(rzk_engine *e, size_t B, const int32_t *c, const int32_t *x,
                                    const int8_t *r, const int8_t *f, uint32_t *flags, void *stream);
int rzk_open_commit_batch_dev(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r,
                              const int32_t *y, int32_t *c, int32_t *t, uint32_t *flags, void *stream);
int rzk_open_respond_batch_dev(rzk_engine *e, size_t B, const int32_t *y, const int8_t *r,
                               const int8_t *d, int32_t *z, void *stream);
int rzk_open_verify_batch_dev(rzk_engine *e, size_t B, const int32_t *z, const int32_t *t,
                              const int32_t *c, uint32_t c_stride, const int8_t *d,
                              uint32_t *flags, void *stream);
int rzk_linear_commit_batch_dev(rzk_engine *e, size_t B, const int32_t *g, const int32_t *x,
                                const int8_t *rp, const int8_t *r, const int32_t *y, const int32_t *yp,
                                int32_t *gx, int32_t *cp, int32_t *c, int32_t *t, int32_t *tp, int32_t *u,
                                uint32_t *flags, void *stream);
int rzk_linear_respond_batch_dev(rzk_engine *e, size_t B, const int32_t *y, const int32_t *yp,
                                 const int8_t *r, const int8_t *rp, const int8_t *d,
                                 int32_t *z, int32_t *zp, void *stream);
int rzk_linear_verify_batch_dev(rzk_engine *e, size_t B, const int32_t *z, const int32_t *zp,
                                const int32_t *c, const int32_t *cp, const int32_t *g,
                                const int32_t *t, const int32_t *tp, const int32_t *u,
                                const int8_t *d, uint32_t *flags, void *stream);
int rzk_sum_commit_batch_dev(rzk_engine *e, size_t B, uint32_t T, const int32_t *gs, const int32_t *xs,
                             const int8_t *rp, const int8_t *rs, const int32_t *ys, const int32_t *yp,
                             int32_t *xp, int32_t *cp, int32_t *cs, int32_t *ts, int32_t *tp, int32_t *u,
                             uint32_t *flags, void *stream);
int rzk_sum_respond_batch_dev(rzk_engine *e, size_t B, uint32_t T, const int32_t *ys, const int32_t *yp,
                              const int8_t *rs, const int8_t *rp, const int8_t *d,
                              int32_t *zs, int32_t *zp, void *stream);
int rzk_sum_verify_batch_dev(rzk_engine *e, size_t B, uint32_t T, const int32_t *zs, const int32_t *zp,
                             const int32_t *cs, const int32_t *cp, const int32_t *gs,
                             const int32_t *ts, const int32_t *tp, const int32_t *u,
                             const int8_t *d, uint32_t *flags, void *stream);
/* bitmap[i>>3] bit (i&7) = (flags[i] & 1) == 0;  *range_any (device word, may be NULL) |= any bit 1 */
int rzk_flags_to_bitmap_dev(rzk_engine *e, size_t B, const uint32_t *flags, uint8_t *bitmap,
                            uint32_t *range_any, void *stream);

/* ---------------------------------------------------------------- i64 staging (device)
 * The reference's coefficient type converts through Into<i64>/From<i64> (tests/test.rs:16,
 * params.rs:126); these convert whole arrays on the device: any i64 representative ->
 * canonical centred i32, and back.  Host pointers; count = number of coefficients. */
int rzk_pack_i64(rzk_engine *e, size_t count, const int64_t *src, int32_t *dst);
int rzk_unpack_i64(rzk_engine *e, size_t count, const int32_t *src, int64_t *dst);

/* ---------------------------------------------------------------- optional on-device samplers
 * TEST AND BENCHMARK USE ONLY -- NOT CRYPTOGRAPHICALLY SECURE.  Philox4x32-10 is a statistical generator keyed by a 64-bit
 * seed, not a CSPRNG: in this protocol y hides r and the unpredictability of d carries soundness, so production callers draw
 * r, y, d from their own CSPRNG (the reference: rand::rng(), ChaCha-based) and pass them in, as every protocol entry point
 * expects.  The samplers exist to measure the flow with r and y resident on the device.
 * NOT part of the reference's flow, where r, y, d are drawn host side by the caller's RNG and passed in.  They keep the
 * prover's r and y on the device between commit and create_response.  Counter-based (Philox4x32-10 keyed by `seed`;
 * `tag` < 2^24 separates streams), so a value depends only on (seed, tag, polynomial index, coefficient index) and is
 * reproducible on the host (tests/philox_ref.py); they do not reproduce the stream of Rust's `rand`.
 *   small:     n_polys polynomials, coefficients exactly uniform in [-b, b]            (polynomial.rs:14-24)
 *   gaussian:  coefficients trunc(N(0, sigma)) (Box-Muller in binary64)                 (polynomial.rs:28-44)
 *   challenge: per item min(kappa, N) entries +-1 at distinct uniform positions         (challenge_space.rs:12-33) */
int rzk_sample_small_dev(rzk_engine *e, size_t n_polys, int32_t b, uint64_t seed, uint32_t tag, int8_t *out, void *stream);
int rzk_sample_gaussian_dev(rzk_engine *e, size_t n_polys, double sigma, uint64_t seed, uint32_t tag, int32_t *out, void *stream);
int rzk_sample_challenge_dev(rzk_engine *e, size_t n_items, int32_t kappa, uint64_t seed, uint32_t tag, int8_t *out, void *stream);

/* ---------------------------------------------------------------- wire format of the messages
 * The reference's own encoding of a batch of messages, packed and parsed on the device: bincode 1.3 (the crate's serde test,
 * mat.rs:424-438: little endian, fixed-width integers, u64 sequence lengths, one tag byte per Option, struct fields in
 * declaration order) of the derive(Serialize) layouts of commit.rs:134-141,222-235, prove/open.rs:180-228,
 * prove/linear.rs:256-315 and prove/sum.rs:327-391.  Mat = Vec<Vec<Polynomial>> (mat.rs:11-17); a Polynomial is the
 * length-prefixed sequence of its coefficients -- mat.rs:434 pins 8 + 8 + (8 + 3*4) = 36 bytes for the 1 x 1 matrix
 * [1 + 2X + 3X^2] over i32.  The serde impls of Polynomial / ZqI64 live in the absent dependency poly-ring-xnp1, so two facts
 * are parameters, not assumptions: `trim` (1: trailing zero coefficients are not stored -- the reading that reproduces the
 * 36 bytes; 0: always N coefficients) and `elem_bytes` (8: ZqI64's i64; 4: i32 as in the pinned test).
 *
 * A message kind is a list of tokens (rzk_wire_layout); its polynomials come from / go to numbered streams, which are the
 * arrays of the protocol entry points above (device pointers, [B][polys_per_item][N]):
 *   RZK_MSG_COMMITMENT         0 = c [2]
 *   RZK_MSG_OPENING            0 = x [1], 1 = r [3] (i8)                  f = None      RZK_MSG_OPENING_F: + 2 = f [1] (i8)
 *   RZK_MSG_OPEN_COMMITMENT    0 = c [2], 1 = t [1]
 *   RZK_MSG_CHALLENGE          0 = d [1] (i8)                            (Open, Linear and Sum challenges are the same struct)
 *   RZK_MSG_OPEN_RESPONSE      0 = z [3]
 *   RZK_MSG_LINEAR_COMMITMENT  0 = c [2], 1 = cp [2], 2 = g [1], 3 = t [1], 4 = tp [1], 5 = u [1]
 *   RZK_MSG_LINEAR_RESPONSE    0 = z [3], 1 = zp [3]     (linear.rs:318 lacks the derive: the layout it WOULD have -- an extension)
 *   RZK_MSG_SUM_COMMITMENT     0 = cp [2], 1 = cs [T*2], 2 = gs [T], 3 = tp [1], 4 = ts [T], 5 = u [1]
 *   RZK_MSG_SUM_RESPONSE       0 = zp [3], 1 = zs [T*3]
 * The contexts (ResponseContext / VerificationContext) never cross the wire in the protocol and are not offered.
 * Pack: item i occupies bytes [offsets[i], offsets[i+1]) of `out`; offsets is a device array of B + 1 words that the call
 * fills, *total_bytes (host) = offsets[B].  out == NULL only computes offsets and the total.  Unpack: offsets is an input
 * (the transport knows the message boundaries; an item whose offsets leave the `in_bytes` of the buffer or run backwards is
 * malformed and is not read); every literal is checked and flags[i] |= 1 marks a malformed item (wrong length or tag, a polynomial longer than N, a coefficient that does not fit an int8 stream, bytes missing or left
 * over); coefficients are canonicalised like ZqI64::from.  Both calls synchronise `stream` before they return. */
enum { RZK_WIRE_END = 0, RZK_WIRE_LEN = 1 /* u64 `value` */, RZK_WIRE_POLY = 2 /* polynomial `poly` of `stream` */, RZK_WIRE_TAG = 3 /* byte `value` */ };
typedef struct { uint32_t kind, stream, poly, value; } rzk_wire_tok;
typedef struct { const void *base; uint32_t polys_per_item; uint32_t dtype; /* 0 = int32, 1 = int8 */ } rzk_wire_stream;
enum { RZK_MSG_COMMITMENT = 1, RZK_MSG_OPENING, RZK_MSG_OPENING_F, RZK_MSG_OPEN_COMMITMENT, RZK_MSG_CHALLENGE, RZK_MSG_OPEN_RESPONSE,
       RZK_MSG_LINEAR_COMMITMENT, RZK_MSG_LINEAR_RESPONSE, RZK_MSG_SUM_COMMITMENT, RZK_MSG_SUM_RESPONSE };
/* Token list of a message kind (T: terms of a Sum proof, ignored otherwise).  toks == NULL: *ntoks = the length needed.  No GPU involved. */
int rzk_wire_layout(int message_kind, uint32_t T, rzk_wire_tok *toks, size_t cap, size_t *ntoks);
int rzk_wire_pack_dev(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                      int elem_bytes, int trim, uint8_t *out, size_t out_capacity, uint64_t *offsets, uint64_t *total_bytes, void *stream);
int rzk_wire_unpack_dev(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                        int elem_bytes, const uint8_t *in, size_t in_bytes, const uint64_t *offsets, uint32_t *flags, void *stream);
/* Host forms (what the Rust shim binds): the streams, the bytes and the offsets are HOST arrays; ok_bitmap bit i = item i parsed. */
int rzk_wire_pack(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                  int elem_bytes, int trim, uint8_t *out, size_t out_capacity, uint64_t *offsets, uint64_t *total_bytes);
int rzk_wire_unpack(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                    int elem_bytes, const uint8_t *in, size_t in_bytes, const uint64_t *offsets, uint8_t *ok_bitmap);

/* ---------------------------------------------------------------- Fiat-Shamir challenges on the device
 * NOT in the reference, which is interactive (open.rs:143-158 draws d from the verifier's RNG; README.md:16 names the
 * transform as a possibility).  docs/FIAT_SHAMIR.md specifies the transcript:
 *     d_i = SampleInBall_kappa( SHAKE128( prefix || polynomials of item i's first message ) )
 * every polynomial absorbed as N little-endian int32 canonical centred coefficients (int8 streams are widened), the segments
 * in the order given -- for the three protocols, the field order of the commitment struct (open.rs:190-198: c, t;
 * linear.rs:271-285: c, cp, g, t, tp, u; sum.rs:342-355: cp, cs, gs, tp, ts, u).  `prefix` (host pointer, a multiple of
 * 8 bytes: domain tag, key digest, shape words -- built by the caller, see the document) separates protocols, keys and shapes.
 * kappa <= 64.  Bit-reproducible on the host with any SHAKE128 (oracle/fs_ref.py uses hashlib).
 *   rzk_fs_challenge_dev          d [B][N] i8 from the transcript segments (device arrays [B][polys_per_item][N])
 *   rzk_open_prove_fs_batch_dev   commit (open.rs:80-103) -> challenge -> response (open.rs:107-117) on one stream without a host
 *                                 round trip: outputs c [B][2][N], t [B][N], d [B][N] i8, z [B][3][N]; flags as for the `_dev` calls
 *   rzk_open_verify_fs_batch_dev  recomputes d from (c, t) into the caller's buffer, then open.rs:162-174 */
int rzk_fs_challenge_dev(rzk_engine *e, size_t B, const uint8_t *prefix, size_t prefix_len, const rzk_wire_stream *segs, int nsegs,
                         int8_t *d, void *stream);
int rzk_open_prove_fs_batch_dev(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, const int32_t *y, const uint8_t *prefix,
                                size_t prefix_len, int32_t *c, int32_t *t, int8_t *d, int32_t *z, uint32_t *flags, void *stream);
int rzk_open_verify_fs_batch_dev(rzk_engine *e, size_t B, const int32_t *c, const int32_t *t, const int32_t *z, const uint8_t *prefix,
                                 size_t prefix_len, int8_t *d, uint32_t *flags, void *stream);
/* Host forms (host pointers, chunked pipeline): prove = commit + challenge + response with r, y supplied (ok bitmap: the commit
 * constraint, as rzk_open_commit_batch); verify returns the verdict bitmap.  A proof is (c, t, z); d is returned to the prover
 * for inspection only and is recomputed by the verifier. */
int rzk_open_prove_fs_batch(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, const int32_t *y, const uint8_t *prefix,
                            size_t prefix_len, int32_t *c, int32_t *t, int8_t *d, int32_t *z, uint8_t *ok_bitmap);
int rzk_open_verify_fs_batch(rzk_engine *e, size_t B, const int32_t *c, const int32_t *t, const int32_t *z, const uint8_t *prefix,
                             size_t prefix_len, uint8_t *verify_bitmap);

/* Counters for the benchmark harness: kernels launched by this engine since creation. */
uint64_t rzk_kernel_launches(const rzk_engine *e);

/* ---------------------------------------------------------------- several GPUs behind one handle
 * For a single-process caller (the Rust crate): a group owns one engine per listed device and one host
 * worker thread per engine.  Every host entry point has a group form with the same arguments; the batch
 * is split into contiguous item ranges whose starts are multiples of 8 (so every range owns whole bytes
 * of the result bitmap), each range runs on its device, and there is no exchange between devices
 * (commit.rs:88-128 reads only self, params, x, r).  A device id may be listed more than once.
 * Status: RZK_OK, or the first non-zero status of any device (message in rzk_group_last_error). */
typedef struct rzk_group rzk_group;
int rzk_group_create(const rzk_params *params, const int *device_ids, int n_devices, rzk_group **out);
void rzk_group_destroy(rzk_group *g);
int rzk_group_size(const rzk_group *g);
const char *rzk_group_last_error(const rzk_group *g);
int rzk_group_set_key(rzk_group *g, const int64_t *a1, const int64_t *a2);
uint64_t rzk_group_kernel_launches(const rzk_group *g);
int rzk_group_commit_batch(rzk_group *g, size_t B, const int32_t *x, const int8_t *r, int32_t *c, uint8_t *ok_bitmap);
int rzk_group_commit_batch_r2(rzk_group *g, size_t B, const int32_t *x, const uint8_t *r2, int32_t *c, uint8_t *ok_bitmap);
int rzk_group_commitment_verify_batch(rzk_group *g, size_t B, const int32_t *c, const int32_t *x, const int8_t *r,
                                      const int8_t *f, uint8_t *verify_bitmap);
int rzk_group_open_commit_batch(rzk_group *g, size_t B, const int32_t *x, const int8_t *r, const int32_t *y,
                                int32_t *c, int32_t *t, uint8_t *ok_bitmap);
int rzk_group_open_respond_batch(rzk_group *g, size_t B, const int32_t *y, const int8_t *r, const int8_t *d, int32_t *z);
int rzk_group_open_verify_batch(rzk_group *g, size_t B, const int32_t *z, const int32_t *t, const int32_t *c1,
                                const int8_t *d, uint8_t *verify_bitmap);
int rzk_group_linear_commit_batch(rzk_group *g, size_t B, const int32_t *gg, const int32_t *x, const int8_t *rp, const int8_t *r,
                                  const int32_t *y, const int32_t *yp, int32_t *gx, int32_t *cp, int32_t *c, int32_t *t,
                                  int32_t *tp, int32_t *u, uint8_t *ok_bitmap);
int rzk_group_linear_respond_batch(rzk_group *g, size_t B, const int32_t *y, const int32_t *yp, const int8_t *r, const int8_t *rp,
                                   const int8_t *d, int32_t *z, int32_t *zp);
int rzk_group_linear_verify_batch(rzk_group *g, size_t B, const int32_t *z, const int32_t *zp, const int32_t *c, const int32_t *cp,
                                  const int32_t *gg, const int32_t *t, const int32_t *tp, const int32_t *u, const int8_t *d,
                                  uint8_t *verify_bitmap);
int rzk_group_sum_commit_batch(rzk_group *g, size_t B, uint32_t T, const int32_t *gs, const int32_t *xs, const int8_t *rp,
                               const int8_t *rs, const int32_t *ys, const int32_t *yp, int32_t *xp, int32_t *cp, int32_t *cs,
                               int32_t *ts, int32_t *tp, int32_t *u, uint8_t *ok_bitmap);
int rzk_group_sum_respond_batch(rzk_group *g, size_t B, uint32_t T, const int32_t *ys, const int32_t *yp, const int8_t *rs,
                                const int8_t *rp, const int8_t *d, int32_t *zs, int32_t *zp);
int rzk_group_sum_verify_batch(rzk_group *g, size_t B, uint32_t T, const int32_t *zs, const int32_t *zp, const int32_t *cs,
                               const int32_t *cp, const int32_t *gs, const int32_t *ts, const int32_t *tp, const int32_t *u,
                               const int8_t *d, uint8_t *verify_bitmap);

#ifdef __cplusplus
}
#endif
#endif
