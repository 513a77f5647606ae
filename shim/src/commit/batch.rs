//! Batched `CommitmentKey::commit` / `Commitment::verify` on the B200 engine (feature `b200`).
//! Child module of `commit` (add `#[cfg(feature = "b200")] mod batch;` at the top of `src/commit.rs`); the sequential
//! methods (`commit.rs:88-128`, `173-210`) are untouched.

use poly_ring_xnp1::Polynomial;
use rand::RngExt;

use super::{Commitment, CommitmentKey, Opening};
use crate::b200::{self, ffi, B200Error, Backend, Z};
use crate::{mat::Mat, params::Params, polynomial::random_polynomial_within};

/// The randomness of ONE `commit` call, drawn exactly as `commit.rs:98-107` draws it: `k` polynomials row by row
/// (`Mat::new_with`, `mat.rs:72-74`), the whole matrix redrawn until `check_commit_constraint` passes -- so a batch
/// consumes the caller's RNG stream in the order B sequential calls would.
/// The constraint `floor(sqrt(sum r^2)) <= 4 sigma floor(sqrt(N))` (`params.rs:102-108`, `polynomial.rs:60-73`) is
/// evaluated exactly as `sum r^2 < (bound + 1)^2` in u128 instead of through `BigUint` per coefficient.
pub(crate) fn draw_commit_randomness<const N: usize>(rng: &mut impl RngExt, params: &Params<Z>) -> Mat<Z, N> {
    let bound = (4 * params.standard_deviation(N) * num::integer::Roots::sqrt(&N)) as u128;
    loop {
        let tmp = Mat::<Z, N>::new_with(params.k, 1, || random_polynomial_within(rng, params.b.clone()));
        let ok = tmp.polynomials.iter().all(|row| {
            row.iter().all(|p| {
                let s: u128 = p.iter().map(|c| { let v: i64 = c.clone().into(); (v as i128 * v as i128) as u128 }).sum();
                s < (bound + 1) * (bound + 1)
            })
        });
        if ok {
            return tmp;
        }
    }
}

/// `c = [a1; a2] . r + [0; x]` for B (x, r) pairs on the engine: `rzk_commit_batch`.
pub(crate) fn commit_with<const N: usize>(
    be: &mut Backend,
    params: &Params<Z>,
    xs: &[Vec<Polynomial<Z, N>>],
    rs: &[Mat<Z, N>],
) -> Result<Vec<Commitment<Z, N>>, B200Error> {
    b200::assert_default_shape(params);
    let b = xs.len();
    let (mut xf, mut rf) = (Vec::with_capacity(b * N), Vec::with_capacity(b * 3 * N));
    for (x, r) in xs.iter().zip(rs) {
        assert_eq!(params.l, x.len()); // commit.rs:95
        b200::push_poly(&mut xf, &x[0]);
        b200::push_mat_i8(&mut rf, r);
    }
    let rows = params.n + params.l;
    let mut c = vec![0i32; b * rows * N];
    let mut ok = vec![0u8; (b + 7) / 8];
    // randomness in {-1, 0, 1} (b = 1, the default) crosses the bus at 2 bits per coefficient: 384 bytes per commitment
    // instead of 1536 (`rzk_commit_batch_r2`); anything wider goes as int8
    let bound: i64 = params.b.clone().into();
    let r2 = if bound == 1 { b200::pack_r2(&rf) } else { None };
    let rc = unsafe {
        match (&*be, &r2) {
            (Backend::Engine(e), Some(p)) => ffi::rzk_commit_batch_r2(*e, b, xf.as_ptr(), p.as_ptr(), c.as_mut_ptr(), ok.as_mut_ptr()),
            (Backend::Group(g), Some(p)) => ffi::rzk_group_commit_batch_r2(*g, b, xf.as_ptr(), p.as_ptr(), c.as_mut_ptr(), ok.as_mut_ptr()),
            (Backend::Engine(e), None) => ffi::rzk_commit_batch(*e, b, xf.as_ptr(), rf.as_ptr(), c.as_mut_ptr(), ok.as_mut_ptr()),
            (Backend::Group(g), None) => ffi::rzk_group_commit_batch(*g, b, xf.as_ptr(), rf.as_ptr(), c.as_mut_ptr(), ok.as_mut_ptr()),
        }
    };
    be.check_or_panic(rc)?;
    debug_assert!((0..b).all(|i| b200::bit(&ok, i)), "r passed the host-side constraint check");
    Ok((0..b).map(|i| Commitment { c: b200::mat_from::<N>(&c[i * rows * N..(i + 1) * rows * N], rows) }).collect())
}

impl<const N: usize> CommitmentKey<Z, N> {
    /// `commit` (`commit.rs:88-128`) for a batch of messages.  Equivalent to calling `commit` once per element of
    /// `xs`, in order, on the same `rng`: the randomness is drawn here on the host, item by item, and the ring
    /// arithmetic of the whole batch runs on the GPU.
    ///
    /// ## Panics
    /// Panics if some `x.len() != params.l` (as `commit` does).
    pub fn commit_batch(
        &self,
        rng: &mut impl RngExt,
        xs: Vec<Vec<Polynomial<Z, N>>>,
        params: &Params<Z>,
        be: &mut Backend,
    ) -> Result<Vec<(Opening<Z, N>, Commitment<Z, N>)>, B200Error> {
        for x in &xs {
            assert_eq!(params.l, x.len());
        }
        let rs: Vec<Mat<Z, N>> = xs.iter().map(|_| draw_commit_randomness::<N>(rng, params)).collect();
        let cs = commit_with(be, params, &xs, &rs)?;
        Ok(xs.into_iter().zip(rs).zip(cs).map(|((x, r), c)| (Opening { x, r, f: None }, c)).collect())
    }
}

impl<const N: usize> Commitment<Z, N> {
    /// `verify` (`commit.rs:173-210`) for B (commitment, opening) pairs: `rzk_commitment_verify_batch`.  Openings with
    /// `f = None` and with `f = Some(..)` go to the engine as two calls (the C entry point takes `f` for the whole
    /// batch or not at all); the result is in the input order.
    pub fn verify_batch(
        commitments: &[Commitment<Z, N>],
        openings: &[Opening<Z, N>],
        params: &Params<Z>,
        be: &mut Backend,
    ) -> Result<Vec<bool>, B200Error> {
        assert_eq!(commitments.len(), openings.len());
        b200::assert_default_shape(params);
        let mut out = vec![false; commitments.len()];
        for with_f in [false, true] {
            let idx: Vec<usize> = (0..openings.len()).filter(|&i| openings[i].f.is_some() == with_f).collect();
            if idx.is_empty() {
                continue;
            }
            let b = idx.len();
            let (mut cf, mut xf, mut rf, mut ff) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
            for &i in &idx {
                b200::push_mat(&mut cf, &commitments[i].c);
                b200::push_poly(&mut xf, &openings[i].x[0]);
                b200::push_mat_i8(&mut rf, &openings[i].r);
                if let Some(f) = &openings[i].f {
                    b200::push_poly_i8(&mut ff, f);
                }
            }
            let fp = if with_f { ff.as_ptr() } else { std::ptr::null() };
            let mut bm = vec![0u8; (b + 7) / 8];
            let rc = unsafe {
                match *be {
                    Backend::Engine(e) => ffi::rzk_commitment_verify_batch(e, b, cf.as_ptr(), xf.as_ptr(), rf.as_ptr(), fp, bm.as_mut_ptr()),
                    Backend::Group(g) => ffi::rzk_group_commitment_verify_batch(g, b, cf.as_ptr(), xf.as_ptr(), rf.as_ptr(), fp, bm.as_mut_ptr()),
                }
            };
            be.check_or_panic(rc)?;
            for (j, &i) in idx.iter().enumerate() {
                out[i] = b200::bit(&bm, j);
            }
        }
        Ok(out)
    }
}
