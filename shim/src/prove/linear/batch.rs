//! Batched Linear proof (`x' = g * x`) on the B200 engine (feature `b200`).  Child module of `prove::linear` (add
//! `#[cfg(feature = "b200")] mod batch;` to `src/prove/linear.rs`).  The sequential methods (`linear.rs:82-158`,
//! `184-250`) are untouched; every `*_batch` method equals its sequential twin called once per element, in order, on the
//! same `rng` (draw order per instance: `r'`, `r`, `y`, `y'` -- `linear.rs:96-115`).

use poly_ring_xnp1::Polynomial;
use rand::RngExt;

use super::{
    LinearProofChallenge, LinearProofCommitment, LinearProofProver, LinearProofResponse, LinearProofResponseContext,
    LinearProofVerificationContext, LinearProofVerifier,
};
use crate::b200::{self, ffi, B200Error, Backend, Z};
use crate::commit::batch::draw_commit_randomness;
use crate::{commit::Commitment, commit::Opening, mat::Mat};

impl<const N: usize> LinearProofProver<Z, N> {
    /// `commit` (`linear.rs:82-140`) for B (g, x) pairs: `g * x`, both commitments, `t`, `t'` and
    /// `u = g * (A2 . y) - A2 . y'` on the engine (`rzk_linear_commit_batch`).
    pub fn commit_batch(
        &self,
        rng: &mut impl RngExt,
        gs: Vec<Polynomial<Z, N>>,
        xs: Vec<Vec<Polynomial<Z, N>>>,
        be: &mut Backend,
    ) -> Result<Vec<(LinearProofResponseContext<Z, N>, LinearProofCommitment<Z, N>)>, B200Error> {
        assert_eq!(gs.len(), xs.len());
        b200::assert_default_shape(&self.params);
        let b = xs.len();
        let (mut rps, mut rs, mut ys, mut yps) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let (mut gf, mut xf, mut rpf, mut rf, mut yf, mut ypf) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (g, x) in gs.iter().zip(&xs) {
            assert_eq!(self.params.l, x.len()); // commit.rs:95, through linear.rs:96-97
            let rp = draw_commit_randomness::<N>(rng, &self.params); // commit(g * x) first (linear.rs:96)
            let r = draw_commit_randomness::<N>(rng, &self.params); // then commit(x)      (linear.rs:97)
            let y = b200::draw_masking::<N>(rng, &self.params); // linear.rs:100-106
            let yp = b200::draw_masking::<N>(rng, &self.params); // linear.rs:109-115
            b200::push_poly(&mut gf, g);
            b200::push_poly(&mut xf, &x[0]);
            b200::push_mat_i8(&mut rpf, &rp);
            b200::push_mat_i8(&mut rf, &r);
            b200::push_mat(&mut yf, &y);
            b200::push_mat(&mut ypf, &yp);
            rps.push(rp);
            rs.push(r);
            ys.push(y);
            yps.push(yp);
        }
        let rows = self.params.n + self.params.l;
        let (mut gx, mut cp, mut c) = (vec![0i32; b * N], vec![0i32; b * rows * N], vec![0i32; b * rows * N]);
        let (mut t, mut tp, mut u) = (vec![0i32; b * N], vec![0i32; b * N], vec![0i32; b * N]);
        let mut ok = vec![0u8; (b + 7) / 8];
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_linear_commit_batch(e, b, gf.as_ptr(), xf.as_ptr(), rpf.as_ptr(), rf.as_ptr(), yf.as_ptr(), ypf.as_ptr(),
                                                                  gx.as_mut_ptr(), cp.as_mut_ptr(), c.as_mut_ptr(), t.as_mut_ptr(), tp.as_mut_ptr(), u.as_mut_ptr(), ok.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_linear_commit_batch(g, b, gf.as_ptr(), xf.as_ptr(), rpf.as_ptr(), rf.as_ptr(), yf.as_ptr(), ypf.as_ptr(),
                                                                        gx.as_mut_ptr(), cp.as_mut_ptr(), c.as_mut_ptr(), t.as_mut_ptr(), tp.as_mut_ptr(), u.as_mut_ptr(), ok.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        let mut out = Vec::with_capacity(b);
        let mut it = gs.into_iter().zip(xs).zip(rps).zip(rs).zip(ys).zip(yps).enumerate();
        while let Some((i, (((((g, x), rp), r), y), yp))) = it.next() {
            let row = |v: &[i32], per: usize| v[i * per * N..(i + 1) * per * N].to_vec();
            out.push((
                LinearProofResponseContext {
                    opening: Opening { x, r, f: None },
                    opening_p: Opening { x: b200::polys_from::<N>(&row(&gx, 1), 1), r: rp, f: None },
                    y,
                    yp,
                },
                LinearProofCommitment {
                    c: Commitment { c: b200::mat_from::<N>(&row(&c, rows), rows) },
                    cp: Commitment { c: b200::mat_from::<N>(&row(&cp, rows), rows) },
                    g,
                    t: b200::polys_from::<N>(&row(&t, 1), 1),
                    tp: b200::polys_from::<N>(&row(&tp, 1), 1),
                    u: b200::mat_from::<N>(&row(&u, 1), 1),
                },
            ));
        }
        Ok(out)
    }

    /// `create_response` (`linear.rs:144-158`) for B instances: `z = y + d * r`, `z' = y' + d * r'`.
    pub fn create_response_batch(
        &self,
        contexts: Vec<LinearProofResponseContext<Z, N>>,
        challenges: Vec<LinearProofChallenge<Z, N>>,
        be: &mut Backend,
    ) -> Result<Vec<LinearProofResponse<Z, N>>, B200Error> {
        assert_eq!(contexts.len(), challenges.len());
        let (b, k) = (contexts.len(), self.params.k);
        let (mut yf, mut ypf, mut rf, mut rpf, mut df) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (ctx, ch) in contexts.iter().zip(&challenges) {
            b200::push_mat(&mut yf, &ctx.y);
            b200::push_mat(&mut ypf, &ctx.yp);
            b200::push_mat_i8(&mut rf, &ctx.opening.r);
            b200::push_mat_i8(&mut rpf, &ctx.opening_p.r);
            b200::push_poly_i8(&mut df, &ch.d);
        }
        let (mut z, mut zp) = (vec![0i32; b * k * N], vec![0i32; b * k * N]);
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_linear_respond_batch(e, b, yf.as_ptr(), ypf.as_ptr(), rf.as_ptr(), rpf.as_ptr(), df.as_ptr(), z.as_mut_ptr(), zp.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_linear_respond_batch(g, b, yf.as_ptr(), ypf.as_ptr(), rf.as_ptr(), rpf.as_ptr(), df.as_ptr(), z.as_mut_ptr(), zp.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        Ok((0..b)
            .map(|i| LinearProofResponse {
                z: b200::mat_from::<N>(&z[i * k * N..(i + 1) * k * N], k),
                zp: b200::mat_from::<N>(&zp[i * k * N..(i + 1) * k * N], k),
            })
            .collect())
    }
}

impl<const N: usize> LinearProofVerifier<Z, N> {
    /// `generate_challenge` (`linear.rs:184-209`) for B commitments (host side only).
    pub fn generate_challenge_batch(
        &self,
        rng: &mut impl RngExt,
        commitments: Vec<LinearProofCommitment<Z, N>>,
    ) -> Vec<(LinearProofVerificationContext<Z, N>, LinearProofChallenge<Z, N>)> {
        commitments.into_iter().map(|c| self.generate_challenge(rng, c)).collect()
    }

    /// `verify` (`linear.rs:213-250`) for B (response, context) pairs: both norm checks and the three equations on the
    /// engine (`rzk_linear_verify_batch`), one bool per instance.
    pub fn verify_batch(
        &self,
        responses: Vec<LinearProofResponse<Z, N>>,
        contexts: Vec<LinearProofVerificationContext<Z, N>>,
        be: &mut Backend,
    ) -> Result<Vec<bool>, B200Error> {
        assert_eq!(responses.len(), contexts.len());
        b200::assert_default_shape(&self.params);
        let b = responses.len();
        let (mut zf, mut zpf, mut cf, mut cpf, mut gf) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let (mut tf, mut tpf, mut uf, mut df) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (resp, ctx) in responses.iter().zip(&contexts) {
            b200::push_mat(&mut zf, &resp.z);
            b200::push_mat(&mut zpf, &resp.zp);
            b200::push_mat(&mut cf, &ctx.c1); // the engine takes the full commitments [c1; c2] (commit.rs:213-218)
            b200::push_mat(&mut cf, &ctx.c2);
            b200::push_mat(&mut cpf, &ctx.c1p);
            b200::push_mat(&mut cpf, &ctx.c2p);
            b200::push_poly(&mut gf, &ctx.g);
            for t in &ctx.t {
                b200::push_poly(&mut tf, t);
            }
            for t in &ctx.tp {
                b200::push_poly(&mut tpf, t);
            }
            b200::push_mat(&mut uf, &ctx.u);
            b200::push_poly_i8(&mut df, &ctx.d);
        }
        let mut bm = vec![0u8; (b + 7) / 8];
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_linear_verify_batch(e, b, zf.as_ptr(), zpf.as_ptr(), cf.as_ptr(), cpf.as_ptr(), gf.as_ptr(), tf.as_ptr(), tpf.as_ptr(), uf.as_ptr(), df.as_ptr(), bm.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_linear_verify_batch(g, b, zf.as_ptr(), zpf.as_ptr(), cf.as_ptr(), cpf.as_ptr(), gf.as_ptr(), tf.as_ptr(), tpf.as_ptr(), uf.as_ptr(), df.as_ptr(), bm.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        Ok((0..b).map(|i| b200::bit(&bm, i)).collect())
    }
}
