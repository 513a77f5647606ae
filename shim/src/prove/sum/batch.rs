//! Batched Sum proof (`x' = sum g_i * x_i`, T terms) on the B200 engine (feature `b200`).  Child module of `prove::sum`
//! (add `#[cfg(feature = "b200")] mod batch;` to `src/prove/sum.rs`).  The sequential methods (`sum.rs:99-200`, `228-320`)
//! are untouched; every `*_batch` method equals its sequential twin called once per element, in order, on the same `rng`
//! (draw order per instance: `r'`, `r_0 .. r_{T-1}`, `y_0 .. y_{T-1}`, `y'` -- `sum.rs:116-142`).
//! All instances of one batch have the same number of terms T (the engine's arrays are `[B][T][..]`).

use poly_ring_xnp1::Polynomial;
use rand::RngExt;

use super::{
    SumProofChallenge, SumProofCommitment, SumProofProver, SumProofResponse, SumProofResponseContext,
    SumProofVerificationContext, SumProofVerifier,
};
use crate::b200::{self, ffi, B200Error, Backend, Z};
use crate::commit::batch::draw_commit_randomness;
use crate::{commit::Commitment, commit::Opening, mat::Mat};

impl<const N: usize> SumProofProver<Z, N> {
    /// `commit` (`sum.rs:99-178`) for B instances of T terms: `x' = sum g_i x_i`, the T + 1 commitments, `t_i`, `t'` and
    /// `u = sum g_i (A2 . y_i) - A2 . y'` on the engine (`rzk_sum_commit_batch`).
    ///
    /// ## Panics
    /// As `commit`: if some `gs` is empty or `gs.len() != xs.len()`; and if the instances differ in T.
    pub fn commit_batch(
        &self,
        rng: &mut impl RngExt,
        gss: Vec<Vec<Polynomial<Z, N>>>,
        xss: Vec<Vec<Vec<Polynomial<Z, N>>>>,
        be: &mut Backend,
    ) -> Result<Vec<(SumProofResponseContext<Z, N>, SumProofCommitment<Z, N>)>, B200Error> {
        assert_eq!(gss.len(), xss.len());
        b200::assert_default_shape(&self.params);
        let b = gss.len();
        if b == 0 {
            return Ok(Vec::new());
        }
        let t_terms = gss[0].len();
        let (mut rps, mut rss, mut yss, mut yps) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let (mut gf, mut xf, mut rpf, mut rf, mut yf, mut ypf) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (gs, xs) in gss.iter().zip(&xss) {
            assert!(!gs.is_empty() && gs.len() == xs.len()); // sum.rs:105
            assert_eq!(gs.len(), t_terms, "all instances of a batch have the same number of terms");
            let rp = draw_commit_randomness::<N>(rng, &self.params); // commit(x') first (sum.rs:116)
            let rs: Vec<Mat<Z, N>> = xs.iter().map(|x| { assert_eq!(self.params.l, x.len()); draw_commit_randomness::<N>(rng, &self.params) }).collect(); // sum.rs:117-120
            let ys: Vec<Mat<Z, N>> = (0..t_terms).map(|_| b200::draw_masking::<N>(rng, &self.params)).collect(); // sum.rs:123-134
            let yp = b200::draw_masking::<N>(rng, &self.params); // sum.rs:137-143
            for i in 0..t_terms {
                b200::push_poly(&mut gf, &gs[i]);
                b200::push_poly(&mut xf, &xs[i][0]);
                b200::push_mat_i8(&mut rf, &rs[i]);
                b200::push_mat(&mut yf, &ys[i]);
            }
            b200::push_mat_i8(&mut rpf, &rp);
            b200::push_mat(&mut ypf, &yp);
            rps.push(rp);
            rss.push(rs);
            yss.push(ys);
            yps.push(yp);
        }
        let rows = self.params.n + self.params.l;
        let (mut xp, mut cp, mut cs) = (vec![0i32; b * N], vec![0i32; b * rows * N], vec![0i32; b * t_terms * rows * N]);
        let (mut ts, mut tp, mut u) = (vec![0i32; b * t_terms * N], vec![0i32; b * N], vec![0i32; b * N]);
        let mut ok = vec![0u8; (b + 7) / 8];
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_sum_commit_batch(e, b, t_terms as u32, gf.as_ptr(), xf.as_ptr(), rpf.as_ptr(), rf.as_ptr(), yf.as_ptr(), ypf.as_ptr(),
                                                               xp.as_mut_ptr(), cp.as_mut_ptr(), cs.as_mut_ptr(), ts.as_mut_ptr(), tp.as_mut_ptr(), u.as_mut_ptr(), ok.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_sum_commit_batch(g, b, t_terms as u32, gf.as_ptr(), xf.as_ptr(), rpf.as_ptr(), rf.as_ptr(), yf.as_ptr(), ypf.as_ptr(),
                                                                     xp.as_mut_ptr(), cp.as_mut_ptr(), cs.as_mut_ptr(), ts.as_mut_ptr(), tp.as_mut_ptr(), u.as_mut_ptr(), ok.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        let mut out = Vec::with_capacity(b);
        let mut it = gss.into_iter().zip(xss).zip(rps).zip(rss).zip(yss).zip(yps).enumerate();
        while let Some((i, (((((gs, xs), rp), rs), ys), yp))) = it.next() {
            let openings: Vec<Opening<Z, N>> = xs.into_iter().zip(rs).map(|(x, r)| Opening { x, r, f: None }).collect();
            let cs_i: Vec<Commitment<Z, N>> = (0..t_terms)
                .map(|j| Commitment { c: b200::mat_from::<N>(&cs[(i * t_terms + j) * rows * N..(i * t_terms + j + 1) * rows * N], rows) })
                .collect();
            let ts_i: Vec<Vec<Polynomial<Z, N>>> = (0..t_terms).map(|j| b200::polys_from::<N>(&ts[(i * t_terms + j) * N..(i * t_terms + j + 1) * N], 1)).collect();
            out.push((
                SumProofResponseContext {
                    openings,
                    opening_p: Opening { x: b200::polys_from::<N>(&xp[i * N..(i + 1) * N], 1), r: rp, f: None },
                    yp,
                    ys,
                },
                SumProofCommitment {
                    cp: Commitment { c: b200::mat_from::<N>(&cp[i * rows * N..(i + 1) * rows * N], rows) },
                    cs: cs_i,
                    gs,
                    tp: b200::polys_from::<N>(&tp[i * N..(i + 1) * N], 1),
                    ts: ts_i,
                    u: b200::mat_from::<N>(&u[i * N..(i + 1) * N], 1),
                },
            ));
        }
        Ok(out)
    }

    /// `create_response` (`sum.rs:182-200`) for B instances: `z_i = y_i + d * r_i`, `z' = y' + d * r'`.
    pub fn create_response_batch(
        &self,
        contexts: Vec<SumProofResponseContext<Z, N>>,
        challenges: Vec<SumProofChallenge<Z, N>>,
        be: &mut Backend,
    ) -> Result<Vec<SumProofResponse<Z, N>>, B200Error> {
        assert_eq!(contexts.len(), challenges.len());
        let (b, k) = (contexts.len(), self.params.k);
        if b == 0 {
            return Ok(Vec::new());
        }
        let t_terms = contexts[0].ys.len();
        let (mut yf, mut ypf, mut rf, mut rpf, mut df) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (ctx, ch) in contexts.iter().zip(&challenges) {
            assert!(ctx.ys.len() == t_terms && ctx.openings.len() == t_terms, "all instances of a batch have the same number of terms");
            for (y, o) in ctx.ys.iter().zip(&ctx.openings) {
                b200::push_mat(&mut yf, y);
                b200::push_mat_i8(&mut rf, &o.r);
            }
            b200::push_mat(&mut ypf, &ctx.yp);
            b200::push_mat_i8(&mut rpf, &ctx.opening_p.r);
            b200::push_poly_i8(&mut df, &ch.d);
        }
        let (mut zs, mut zp) = (vec![0i32; b * t_terms * k * N], vec![0i32; b * k * N]);
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_sum_respond_batch(e, b, t_terms as u32, yf.as_ptr(), ypf.as_ptr(), rf.as_ptr(), rpf.as_ptr(), df.as_ptr(), zs.as_mut_ptr(), zp.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_sum_respond_batch(g, b, t_terms as u32, yf.as_ptr(), ypf.as_ptr(), rf.as_ptr(), rpf.as_ptr(), df.as_ptr(), zs.as_mut_ptr(), zp.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        Ok((0..b)
            .map(|i| SumProofResponse {
                zp: b200::mat_from::<N>(&zp[i * k * N..(i + 1) * k * N], k),
                zs: (0..t_terms).map(|j| b200::mat_from::<N>(&zs[(i * t_terms + j) * k * N..(i * t_terms + j + 1) * k * N], k)).collect(),
            })
            .collect())
    }
}

impl<const N: usize> SumProofVerifier<Z, N> {
    /// `generate_challenge` (`sum.rs:228-253`) for B commitments (host side only).
    pub fn generate_challenge_batch(
        &self,
        rng: &mut impl RngExt,
        commitments: Vec<SumProofCommitment<Z, N>>,
    ) -> Vec<(SumProofVerificationContext<Z, N>, SumProofChallenge<Z, N>)> {
        commitments.into_iter().map(|c| self.generate_challenge(rng, c)).collect()
    }

    /// `verify` (`sum.rs:257-320`) for B (response, context) pairs, one bool per instance (`rzk_sum_verify_batch`).
    /// An instance whose response and context disagree in the number of terms is `false`, as in the reference, where the
    /// `Vec` comparison of `sum.rs:289` fails for it (the length test of `sum.rs:273` uses `&&` and lets it through);
    /// the instances that go to the engine must share one T.
    pub fn verify_batch(
        &self,
        responses: Vec<SumProofResponse<Z, N>>,
        contexts: Vec<SumProofVerificationContext<Z, N>>,
        be: &mut Backend,
    ) -> Result<Vec<bool>, B200Error> {
        assert_eq!(responses.len(), contexts.len());
        b200::assert_default_shape(&self.params);
        let well_formed = |r: &SumProofResponse<Z, N>, c: &SumProofVerificationContext<Z, N>| {
            !r.zs.is_empty() && r.zs.len() == c.ts.len() && r.zs.len() == c.cs.len() && r.zs.len() == c.gs.len()
        };
        let idx: Vec<usize> = (0..responses.len()).filter(|&i| well_formed(&responses[i], &contexts[i])).collect();
        let mut out = vec![false; responses.len()];
        if idx.is_empty() {
            return Ok(out);
        }
        let (b, t_terms) = (idx.len(), responses[idx[0]].zs.len());
        let (mut zf, mut zpf, mut cf, mut cpf, mut gf) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let (mut tf, mut tpf, mut uf, mut df) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for &i in &idx {
            let (resp, ctx) = (&responses[i], &contexts[i]);
            assert_eq!(resp.zs.len(), t_terms, "all instances of a batch have the same number of terms");
            for j in 0..t_terms {
                b200::push_mat(&mut zf, &resp.zs[j]);
                b200::push_mat(&mut cf, &ctx.cs[j].0); // full commitments [c1; c2]
                b200::push_mat(&mut cf, &ctx.cs[j].1);
                b200::push_poly(&mut gf, &ctx.gs[j]);
                for t in &ctx.ts[j] {
                    b200::push_poly(&mut tf, t);
                }
            }
            b200::push_mat(&mut zpf, &resp.zp);
            b200::push_mat(&mut cpf, &ctx.c1p);
            b200::push_mat(&mut cpf, &ctx.c2p);
            for t in &ctx.tp {
                b200::push_poly(&mut tpf, t);
            }
            b200::push_mat(&mut uf, &ctx.u);
            b200::push_poly_i8(&mut df, &ctx.d);
        }
        let mut bm = vec![0u8; (b + 7) / 8];
        let rc = unsafe {
            match *be {
                Backend::Engine(e) => ffi::rzk_sum_verify_batch(e, b, t_terms as u32, zf.as_ptr(), zpf.as_ptr(), cf.as_ptr(), cpf.as_ptr(), gf.as_ptr(), tf.as_ptr(), tpf.as_ptr(), uf.as_ptr(), df.as_ptr(), bm.as_mut_ptr()),
                Backend::Group(g) => ffi::rzk_group_sum_verify_batch(g, b, t_terms as u32, zf.as_ptr(), zpf.as_ptr(), cf.as_ptr(), cpf.as_ptr(), gf.as_ptr(), tf.as_ptr(), tpf.as_ptr(), uf.as_ptr(), df.as_ptr(), bm.as_mut_ptr()),
            }
        };
        be.check_or_panic(rc)?;
        for (j, &i) in idx.iter().enumerate() {
            out[i] = b200::bit(&bm, j);
        }
        Ok(out)
    }
}
