//! Batched Open proof on the B200 engine (feature `b200`).  Child module of `prove::open` (add
//! `#[cfg(feature = "b200")] mod batch;` to `src/prove/open.rs`), so it can build the message structs whose fields are
//! private to that module.  The sequential methods (`open.rs:80-117`, `143-174`) are untouched.
//!
//! Every `*_batch` method is equivalent to calling its sequential twin once per element, in order, on the same `rng`
//! (RNG draw order per instance: `r`, then `y` -- `open.rs:85-94`; the verifier: kappa `random_bool` calls and one
//! `shuffle` per challenge -- `challenge_space.rs:23-31`).

use poly_ring_xnp1::Polynomial;
use rand::RngExt;

use super::{
    OpenProofChallenge, OpenProofCommitment, OpenProofProver, OpenProofResponse, OpenProofResponseContext,
    OpenProofVerificationContext, OpenProofVerifier,
};
use crate::b200::{self, ffi, B200Error, Backend, Z};
use crate::commit::batch::draw_commit_randomness;
use crate::{commit::Commitment, commit::Opening};

impl<const N: usize> OpenProofProver<Z, N> {
    /// `commit` (`open.rs:80-103`) for B messages: commitments `c` and `t = A1 . y` on the engine
    /// (`rzk_open_commit_batch`).
    pub fn commit_batch(
        &self,
        rng: &mut impl RngExt,
        xs: Vec<Vec<Polynomial<Z, N>>>,
        be: &mut Backend,
    ) -> Result<Vec<(OpenProofResponseContext<Z, N>, OpenProofCommitment<Z, N>)>, B200Error> {
        b200::assert_default_shape(&self.params);
        let b = xs.len();
        let (mut rs, mut ys) = (Vec::with_capacity(b), Vec::with_capacity(b));
        let (mut xf, mut rf, mut yf) = (Vec::with_capacity(b * N), Vec::with_capacity(b * 3 * N), Vec::with_capacity(b * 3 * N));
        for x in &xs {
            assert_eq!(self.params.l, x.len()); // commit.rs:95
            let r = draw_commit_randomness::<N>(rng, &self.params); // ck.commit draws r first (open.rs:85)
            let y = b200::draw_masking::<N>(rng, &self.params); // open.rs:88-94
            b200::push_poly(&mut xf, &x[0]);
            b200::push_mat_i8(&mut rf, &r);
            b200::push_mat(&mut yf, &y);
            rs.push(r);
            ys.push(y);
        }
        let (rows, n) = (self.params.n + self.params.l, self.params.n);
        let (mut c, mut t) = (vec![0i32; b * rows * N], vec![0i32; b * n * N]);
        let mut ok = vec![0u8; (b + 7) / 8];
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_open_commit_batch(e, b, xf.as_ptr(), rf.as_ptr(), yf.as_ptr(), c.as_mut_ptr(), t.as_mut_ptr(), ok.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_open_commit_batch(g, b, xf.as_ptr(), rf.as_ptr(), yf.as_ptr(), c.as_mut_ptr(), t.as_mut_ptr(), ok.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        Ok(xs
            .into_iter()
            .zip(rs)
            .zip(ys)
            .enumerate()
            .map(|(i, ((x, r), y))| {
                (
                    OpenProofResponseContext { opening: Opening { x, r, f: None }, y },
                    OpenProofCommitment {
                        c: Commitment { c: b200::mat_from::<N>(&c[i * rows * N..(i + 1) * rows * N], rows) },
                        t: b200::polys_from::<N>(&t[i * n * N..(i + 1) * n * N], n),
                    },
                )
            })
            .collect())
    }

    /// `create_response` (`open.rs:107-117`) for B instances: `z = y + d * r` (`rzk_open_respond_batch`).
    pub fn create_response_batch(
        &self,
        contexts: Vec<OpenProofResponseContext<Z, N>>,
        challenges: Vec<OpenProofChallenge<Z, N>>,
        be: &mut Backend,
    ) -> Result<Vec<OpenProofResponse<Z, N>>, B200Error> {
        assert_eq!(contexts.len(), challenges.len());
        let (b, k) = (contexts.len(), self.params.k);
        let (mut yf, mut rf, mut df) = (Vec::with_capacity(b * k * N), Vec::with_capacity(b * k * N), Vec::with_capacity(b * N));
        for (ctx, ch) in contexts.iter().zip(&challenges) {
            b200::push_mat(&mut yf, &ctx.y);
            b200::push_mat_i8(&mut rf, &ctx.opening.r);
            b200::push_poly_i8(&mut df, &ch.d);
        }
        let mut z = vec![0i32; b * k * N];
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_open_respond_batch(e, b, yf.as_ptr(), rf.as_ptr(), df.as_ptr(), z.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_open_respond_batch(g, b, yf.as_ptr(), rf.as_ptr(), df.as_ptr(), z.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        Ok((0..b).map(|i| OpenProofResponse { z: b200::mat_from::<N>(&z[i * k * N..(i + 1) * k * N], k) }).collect())
    }

    /// NOT in the reference (an extension; `README.md:16` names the transform as a possibility): non-interactive Open proofs
    /// for B messages.  The challenge of instance i is `SampleInBall(SHAKE128(prefix || c_i || t_i))` computed on the device
    /// between `commit` and `create_response` (docs/FIAT_SHAMIR.md of the engine's repository; `prefix` = domain tag, key
    /// digest, shape words and an optional session id, a multiple of 8 bytes -- `b200::fs_prefix`).  Randomness is drawn
    /// exactly as `commit_batch` draws it (`r`, then `y`, per instance).  A proof is `(OpenProofCommitment, OpenProofResponse)`.
    pub fn prove_batch_fs(
        &self,
        rng: &mut impl RngExt,
        xs: Vec<Vec<Polynomial<Z, N>>>,
        prefix: &[u8],
        be: &mut Backend,
    ) -> Result<Vec<(Opening<Z, N>, OpenProofCommitment<Z, N>, OpenProofResponse<Z, N>)>, B200Error> {
        b200::assert_default_shape(&self.params);
        assert_eq!(prefix.len() % 8, 0, "the transcript prefix is a multiple of 8 bytes");
        let e = match *be {
            Backend::Engine(e) => e,
            Backend::Group(_) => return Err(B200Error::Unsupported("the Fiat-Shamir entry points take a single engine".into())),
        };
        let b = xs.len();
        let mut rs = Vec::with_capacity(b);
        let (mut xf, mut rf, mut yf) = (Vec::with_capacity(b * N), Vec::with_capacity(b * 3 * N), Vec::with_capacity(b * 3 * N));
        for x in &xs {
            assert_eq!(self.params.l, x.len()); // commit.rs:95
            let r = draw_commit_randomness::<N>(rng, &self.params);
            let y = b200::draw_masking::<N>(rng, &self.params);
            b200::push_poly(&mut xf, &x[0]);
            b200::push_mat_i8(&mut rf, &r);
            b200::push_mat(&mut yf, &y);
            rs.push(r);
        }
        let (rows, n, k) = (self.params.n + self.params.l, self.params.n, self.params.k);
        let (mut c, mut t, mut z) = (vec![0i32; b * rows * N], vec![0i32; b * n * N], vec![0i32; b * k * N]);
        let mut d = vec![0i8; b * N];
        let mut ok = vec![0u8; (b + 7) / 8];
        let rc = unsafe {
            ffi::rzk_open_prove_fs_batch(e, b, xf.as_ptr(), rf.as_ptr(), yf.as_ptr(), prefix.as_ptr(), prefix.len(), c.as_mut_ptr(), t.as_mut_ptr(), d.as_mut_ptr(), z.as_mut_ptr(), ok.as_mut_ptr())
        };
        be.check_or_panic(rc)?;
        Ok(xs
            .into_iter()
            .zip(rs)
            .enumerate()
            .map(|(i, (x, r))| {
                (
                    Opening { x, r, f: None },
                    OpenProofCommitment {
                        c: Commitment { c: b200::mat_from::<N>(&c[i * rows * N..(i + 1) * rows * N], rows) },
                        t: b200::polys_from::<N>(&t[i * n * N..(i + 1) * n * N], n),
                    },
                    OpenProofResponse { z: b200::mat_from::<N>(&z[i * k * N..(i + 1) * k * N], k) },
                )
            })
            .collect())
    }
}

impl<const N: usize> OpenProofVerifier<Z, N> {
    /// Verifies non-interactive Open proofs made by `OpenProofProver::prove_batch_fs` with the same `prefix`: the challenge is
    /// recomputed from `(c, t)` on the device, then `open.rs:162-174`.  NOT in the reference (see `prove_batch_fs`).
    pub fn verify_batch_fs(
        &self,
        proofs: &[(OpenProofCommitment<Z, N>, OpenProofResponse<Z, N>)],
        prefix: &[u8],
        be: &mut Backend,
    ) -> Result<Vec<bool>, B200Error> {
        b200::assert_default_shape(&self.params);
        assert_eq!(prefix.len() % 8, 0, "the transcript prefix is a multiple of 8 bytes");
        let e = match *be {
            Backend::Engine(e) => e,
            Backend::Group(_) => return Err(B200Error::Unsupported("the Fiat-Shamir entry points take a single engine".into())),
        };
        let b = proofs.len();
        let (mut cf, mut tf, mut zf) = (Vec::with_capacity(b * 2 * N), Vec::with_capacity(b * N), Vec::with_capacity(b * 3 * N));
        for (com, resp) in proofs {
            b200::push_mat(&mut cf, &com.c.c);
            for t in &com.t {
                b200::push_poly(&mut tf, t);
            }
            b200::push_mat(&mut zf, &resp.z);
        }
        let mut bm = vec![0u8; (b + 7) / 8];
        let rc = unsafe { ffi::rzk_open_verify_fs_batch(e, b, cf.as_ptr(), tf.as_ptr(), zf.as_ptr(), prefix.as_ptr(), prefix.len(), bm.as_mut_ptr()) };
        be.check_or_panic(rc)?;
        Ok((0..b).map(|i| b200::bit(&bm, i)).collect())
    }

    /// `generate_challenge` (`open.rs:143-158`) for B commitments.  Host side only (no ring arithmetic): it is the
    /// sequential method in a loop, kept here so that a batch flow reads the same as the single-instance one.
    pub fn generate_challenge_batch(
        &self,
        rng: &mut impl RngExt,
        commitments: Vec<OpenProofCommitment<Z, N>>,
    ) -> Vec<(OpenProofVerificationContext<Z, N>, OpenProofChallenge<Z, N>)> {
        commitments.into_iter().map(|c| self.generate_challenge(rng, c)).collect()
    }

    /// `verify` (`open.rs:162-174`) for B (response, context) pairs: the norm check on `z` and
    /// `A1 . z == t + c1 * d` on the engine (`rzk_open_verify_batch`), one bool per instance.
    pub fn verify_batch(
        &self,
        responses: Vec<OpenProofResponse<Z, N>>,
        contexts: Vec<OpenProofVerificationContext<Z, N>>,
        be: &mut Backend,
    ) -> Result<Vec<bool>, B200Error> {
        assert_eq!(responses.len(), contexts.len());
        b200::assert_default_shape(&self.params);
        let b = responses.len();
        let (mut zf, mut tf, mut cf, mut df) = (Vec::with_capacity(b * 3 * N), Vec::with_capacity(b * N), Vec::with_capacity(b * N), Vec::with_capacity(b * N));
        for (resp, ctx) in responses.iter().zip(&contexts) {
            b200::push_mat(&mut zf, &resp.z);
            for t in &ctx.t {
                b200::push_poly(&mut tf, t);
            }
            b200::push_mat(&mut cf, &ctx.c1);
            b200::push_poly_i8(&mut df, &ctx.d);
        }
        let mut bm = vec![0u8; (b + 7) / 8];
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_open_verify_batch(e, b, zf.as_ptr(), tf.as_ptr(), cf.as_ptr(), df.as_ptr(), bm.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_open_verify_batch(g, b, zf.as_ptr(), tf.as_ptr(), cf.as_ptr(), df.as_ptr(), bm.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        Ok((0..b).map(|i| b200::bit(&bm, i)).collect())
    }
}

impl<const N: usize> OpenProofCommitment<Z, N> {
    /// The messages of a batch in the crate's own bincode encoding, packed by the engine (`b200::wire::pack`,
    /// `RZK_MSG_OPEN_COMMITMENT`): `&bytes[off[i]..off[i + 1]] == bincode::serialize(&coms[i])`.
    pub fn to_wire_batch(coms: &[Self], be: &mut Backend) -> Result<(Vec<u8>, Vec<u64>), B200Error> {
        let (mut cf, mut tf) = (Vec::with_capacity(coms.len() * 2 * N), Vec::with_capacity(coms.len() * N));
        for c in coms {
            b200::push_mat(&mut cf, &c.c.c);
            for t in &c.t {
                b200::push_poly(&mut tf, t);
            }
        }
        b200::wire::pack(be, ffi::RZK_MSG_OPEN_COMMITMENT, 0, coms.len(), &[b200::wire::Stream::I32(&cf, 2), b200::wire::Stream::I32(&tf, 1)])
    }
}
