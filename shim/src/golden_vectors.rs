//! Golden-vector test: pins the crate's OWN arithmetic (poly-ring-xnp1's `Polynomial<ZqI64<Q>, N>` operators as composed
//! by `Mat::dot / add / sub / componentwise_mul`, `CommitmentKey::commit`'s equation, and the three protocols) against
//! the vectors of the B200 engine's repository (`tests/golden/ringzk_n512.bin`, exported by
//! `tests/golden/export_flat.py`).  The engine, its C oracle and its Python big-int oracle all reproduce those vectors
//! bit for bit; this test closes the loop to the real crate.  It needs NO GPU and NO feature flag:
//!
//!     RINGZK_GOLDEN=/path/to/ringzk_n512.bin cargo test golden_vectors
//!
//! Add `#[cfg(test)] mod golden_vectors;` to `src/lib.rs` (a unit-test module, so that it can reach the
//! `pub(crate)` `Mat` and the key's fields and feed the recorded randomness r, y, d instead of drawing it).
//! NOT COMPILED where it was written (no Rust toolchain there).
//!
//! What a failure means: if `representative_is_canonical_centred` fails, `ZqI64` does not keep the canonical centred
//! residue (SURVEY 8c of the engine's repository assumed it from `polynomial.rs:22-23`, `commit.rs:100-105`,
//! `open.rs:171-173`) and only the flatten / unflatten layer of the shim changes; if a product test fails, the engine
//! and the crate disagree on ring arithmetic and the engine must not be used.

use std::collections::HashMap;

use poly_ring_xnp1::{zq::ZqI64, Polynomial};

use crate::{commit::Commitment, commit::Opening, mat::Mat, CommitmentKey, Params};

const N: usize = 512;
const Q: i64 = 3515337053;
type Z = ZqI64<Q>;
type P = Polynomial<Z, N>;

/// One array of the flat file: dims + values widened to i64.
struct Arr {
    dims: Vec<usize>,
    v: Vec<i64>,
}

/// Parser of the flat format: magic "RZKGOLD1", u32 count, then per entry: u32 name length, name, u8 dtype
/// (0 = i8, 1 = i32, 2 = i64, 3 = u8), u32 ndim, u32 dims[ndim], little-endian data.
fn load() -> HashMap<String, Arr> {
    let path = std::env::var("RINGZK_GOLDEN").unwrap_or_else(|_| "tests/golden/ringzk_n512.bin".to_string());
    let buf = std::fs::read(&path).unwrap_or_else(|e| panic!("cannot read {}: {} (set RINGZK_GOLDEN)", path, e));
    assert_eq!(&buf[..8], b"RZKGOLD1");
    let u32_at = |o: usize| u32::from_le_bytes([buf[o], buf[o + 1], buf[o + 2], buf[o + 3]]) as usize;
    let (count, mut off) = (u32_at(8), 12);
    let mut out = HashMap::new();
    for _ in 0..count {
        let ln = u32_at(off);
        off += 4;
        let name = String::from_utf8(buf[off..off + ln].to_vec()).unwrap();
        off += ln;
        let dt = buf[off];
        let nd = u32_at(off + 1);
        off += 5;
        let dims: Vec<usize> = (0..nd).map(|i| u32_at(off + 4 * i)).collect();
        off += 4 * nd;
        let cnt: usize = dims.iter().product();
        let mut v = Vec::with_capacity(cnt);
        for i in 0..cnt {
            v.push(match dt {
                0 => buf[off + i] as i8 as i64,
                3 => buf[off + i] as i64,
                1 => i32::from_le_bytes([buf[off + 4 * i], buf[off + 4 * i + 1], buf[off + 4 * i + 2], buf[off + 4 * i + 3]]) as i64,
                2 => i64::from_le_bytes(buf[off + 8 * i..off + 8 * i + 8].try_into().unwrap()),
                _ => panic!("dtype {}", dt),
            });
        }
        off += cnt * [1, 4, 8, 1][dt as usize];
        out.insert(name, Arr { dims, v });
    }
    assert_eq!(off, buf.len());
    out
}

/// Polynomial number `idx` (row-major over all leading dims) of an array whose last dim is N.
fn poly(a: &Arr, idx: usize) -> P {
    assert_eq!(*a.dims.last().unwrap(), N);
    Polynomial::new(a.v[idx * N..(idx + 1) * N].iter().map(|&c| Z::from(c)).collect())
}
/// `rows` consecutive polynomials starting at `first` as a (rows x 1) matrix.
fn mat(a: &Arr, first: usize, rows: usize) -> Mat<Z, N> {
    Mat::from_vec((first..first + rows).map(|i| poly(a, i)).collect())
}
/// Coefficients of a polynomial as i64, zero padded to N (what the engine's arrays hold).
fn coeffs(p: &P) -> Vec<i64> {
    let mut c: Vec<i64> = p.iter().map(|v| v.clone().into()).collect();
    c.resize(N, 0);
    c
}
fn assert_poly(p: &P, a: &Arr, idx: usize, what: &str) {
    assert_eq!(coeffs(p), a.v[idx * N..(idx + 1) * N].to_vec(), "{} (polynomial {})", what, idx);
}
fn assert_mat(m: &Mat<Z, N>, a: &Arr, first: usize, what: &str) {
    let mut i = first;
    for row in &m.polynomials {
        for p in row {
            assert_poly(p, a, i, what);
            i += 1;
        }
    }
}

/// commit.rs:33-60 with the recorded random blocks: a1 = [1 | a11 a12], a2 = [0 | 1 | a22].
fn key(g: &HashMap<String, Arr>) -> CommitmentKey<Z, N> {
    let mut a1 = Mat::<Z, N>::diag(1, 1, P::one());
    a1.extend_cols(Mat { polynomials: vec![vec![poly(&g["a1p"], 0), poly(&g["a1p"], 1)]] });
    let mut a2 = Mat::<Z, N>::from_element(1, 1, P::zero());
    a2.extend_cols(Mat::<Z, N>::diag(1, 1, P::one()));
    a2.extend_cols(Mat { polynomials: vec![vec![poly(&g["a2p"], 0)]] });
    CommitmentKey { a1, a2 }
}

/// commit.rs:109-125 with r supplied: c = [a1; a2] . r + [0_n; x].
fn commit_with(ck: &CommitmentKey<Z, N>, x: &P, r: &Mat<Z, N>) -> Mat<Z, N> {
    let mut a = ck.a1.clone();
    a.extend_rows(ck.a2.clone());
    let mut z = Mat::<Z, N>::from_element(1, 1, P::zero());
    z.extend_rows(Mat::from_vec(vec![x.clone()]));
    a.dot(r).add(&z)
}

#[test]
fn representative_is_canonical_centred() {
    // the one-liner of SURVEY 8(c): ZqI64::<3515337053>::from(1757668527).into() == -1757668526
    let v: i64 = Z::from(1757668527_i64).into();
    assert_eq!(v, -1757668526);
    let g = load();
    for (i, o) in g["rep_in"].v.iter().zip(&g["rep_out"].v) {
        let got: i64 = Z::from(*i).into();
        assert_eq!(got, *o, "ZqI64::from({})", i);
    }
}

#[test]
fn ring_operators_match() {
    // Polynomial `*`, `+`, `-` (called from mat.rs:109-110, 135-136, 160-161, 176): large x large, small x large,
    // sparse challenge, extreme residues, and the negacyclic wrap x^(N-1) * x = -1
    let g = load();
    let (a, b) = (poly(&g["a1p"], 0), poly(&g["a2p"], 0));
    let (r, d, hi) = (poly(&g["r"], 1), poly(&g["d"], 0), poly(&g["p_hi"], 0));
    assert_poly(&(a.clone() * b.clone()), &g["p_ab"], 0, "a * b");
    assert_poly(&(a.clone() * r), &g["p_ar"], 0, "a * r");
    assert_poly(&(a.clone() * d), &g["p_ad"], 0, "a * d");
    assert_poly(&(hi.clone() * hi), &g["p_hh"], 0, "hi * hi");
    assert_poly(&(a.clone() + b.clone()), &g["p_a_plus_b"], 0, "a + b");
    assert_poly(&(a - b), &g["p_a_minus_b"], 0, "a - b");
    let mut xn1 = vec![Z::from(0_i64); N];
    xn1[N - 1] = Z::from(1_i64);
    let x1 = vec![Z::from(0_i64), Z::from(1_i64)];
    assert_poly(&(P::new(xn1) * P::new(x1)), &g["p_wrap"], 0, "x^(N-1) * x");
}

#[test]
fn commitment_and_open_proof_match() {
    let g = load();
    let (ck, params) = (key(&g), Params::default());
    let b = g["x"].dims[0];
    for i in 0..b {
        let (x, r, y, d) = (poly(&g["x"], i), mat(&g["r"], 3 * i, 3), mat(&g["y"], 3 * i, 3), poly(&g["d"], i));
        // CommitmentKey::commit with the recorded r (commit.rs:123-125) and Commitment::verify (commit.rs:173-210)
        let c = commit_with(&ck, &x, &r);
        assert_mat(&c, &g["c"], 2 * i, "commit c");
        let com = Commitment { c: c.clone() };
        assert!(com.verify(&Opening { x: vec![x.clone()], r: r.clone(), f: None }, &ck, &params));
        // OpenProofProver::commit: t = A1 . y (open.rs:97); create_response: z = y + d * r (open.rs:113-115)
        let t = ck.a1.dot(&y);
        assert_mat(&t, &g["t"], i, "open t");
        let z = y.add(&r.componentwise_mul(&d));
        assert_mat(&z, &g["z"], 3 * i, "open z");
        // OpenProofVerifier::verify (open.rs:162-174)
        let (c1, _) = com.c1_c2(&params);
        let verify = |z: &Mat<Z, N>| params.check_verify_constraint(z) && ck.a1.dot(z) == t.add(&c1.componentwise_mul(&d));
        assert_eq!(verify(&z), g["open_ok"].v[i] == 1);
        let mut zb = z.clone();
        let mut cz = coeffs(&zb.polynomials[1][0]);
        cz[7] += 1; // the tampering recorded as open_bad (tests/golden/make_golden.py)
        zb.polynomials[1][0] = Polynomial::new(cz.into_iter().map(Z::from).collect());
        assert_eq!(verify(&zb), g["open_bad"].v[i] == 1);
    }
}

#[test]
fn linear_proof_matches() {
    let g = load();
    let (ck, params) = (key(&g), Params::default());
    for i in 0..g["x"].dims[0] {
        let (gg, x, d) = (poly(&g["g"], i), poly(&g["x"], i), poly(&g["d"], i));
        let (rp, r, y, yp) = (mat(&g["rp"], 3 * i, 3), mat(&g["r"], 3 * i, 3), mat(&g["y"], 3 * i, 3), mat(&g["yp"], 3 * i, 3));
        // linear.rs:91-129
        let gx = x.clone() * gg.clone();
        assert_poly(&gx, &g["l_gx"], i, "linear g*x");
        let (cp, c) = (commit_with(&ck, &gx, &rp), commit_with(&ck, &x, &r));
        assert_mat(&cp, &g["l_cp"], 2 * i, "linear cp");
        assert_mat(&c, &g["l_c"], 2 * i, "linear c");
        let (t, tp) = (ck.a1.dot(&y), ck.a1.dot(&yp));
        assert_mat(&t, &g["l_t"], i, "linear t");
        assert_mat(&tp, &g["l_tp"], i, "linear tp");
        let u = ck.a2.dot(&y).componentwise_mul(&gg).sub(&ck.a2.dot(&yp));
        assert_mat(&u, &g["l_u"], i, "linear u");
        // linear.rs:150-156
        let (z, zp) = (y.add(&r.componentwise_mul(&d)), yp.add(&rp.componentwise_mul(&d)));
        assert_mat(&z, &g["l_z"], 3 * i, "linear z");
        assert_mat(&zp, &g["l_zp"], 3 * i, "linear zp");
        // linear.rs:213-250
        let (c1, c2) = Commitment { c }.c1_c2(&params);
        let (c1p, c2p) = Commitment { c: cp }.c1_c2(&params);
        let verify = |u: &Mat<Z, N>| {
            params.check_verify_constraint(&z)
                && params.check_verify_constraint(&zp)
                && ck.a1.dot(&z) == t.add(&c1.componentwise_mul(&d))
                && ck.a1.dot(&zp) == tp.add(&c1p.componentwise_mul(&d))
                && ck.a2.dot(&z).componentwise_mul(&gg).sub(&ck.a2.dot(&zp)) == c2.componentwise_mul(&gg).sub(&c2p).componentwise_mul(&d).add(u)
        };
        assert_eq!(verify(&u), g["l_ok"].v[i] == 1);
        let mut ub = u.clone();
        let mut cu = coeffs(&ub.polynomials[0][0]);
        cu[9] += 1; // recorded as l_bad
        ub.polynomials[0][0] = Polynomial::new(cu.into_iter().map(Z::from).collect());
        assert_eq!(verify(&ub), g["l_bad"].v[i] == 1);
    }
}

#[test]
fn sum_proof_matches() {
    let g = load();
    let (ck, params) = (key(&g), Params::default());
    let (b, t_terms) = (g["gs"].dims[0], g["gs"].dims[1]);
    for i in 0..b {
        let d = poly(&g["d"], i);
        let gs: Vec<P> = (0..t_terms).map(|j| poly(&g["gs"], i * t_terms + j)).collect();
        let xs: Vec<P> = (0..t_terms).map(|j| poly(&g["xs"], i * t_terms + j)).collect();
        let rs: Vec<Mat<Z, N>> = (0..t_terms).map(|j| mat(&g["rs"], 3 * (i * t_terms + j), 3)).collect();
        let ys: Vec<Mat<Z, N>> = (0..t_terms).map(|j| mat(&g["ys"], 3 * (i * t_terms + j), 3)).collect();
        let (rp, yp) = (mat(&g["rps"], 3 * i, 3), mat(&g["yps"], 3 * i, 3));
        // sum.rs:107-160
        let xp = xs.iter().cloned().map(|x| Mat::<Z, N>::from_vec(vec![x])).zip(gs.iter()).map(|(x, g)| x.componentwise_mul(g))
            .reduce(|acc, x| acc.add(&x)).unwrap();
        assert_mat(&xp, &g["s_xp"], i, "sum x'");
        let cp = commit_with(&ck, &xp.polynomials[0][0], &rp);
        assert_mat(&cp, &g["s_cp"], 2 * i, "sum cp");
        let cs: Vec<Mat<Z, N>> = xs.iter().zip(&rs).map(|(x, r)| commit_with(&ck, x, r)).collect();
        let ts: Vec<Mat<Z, N>> = ys.iter().map(|y| ck.a1.dot(y)).collect();
        for j in 0..t_terms {
            assert_mat(&cs[j], &g["s_cs"], 2 * (i * t_terms + j), "sum c_j");
            assert_mat(&ts[j], &g["s_ts"], i * t_terms + j, "sum t_j");
        }
        let tp = ck.a1.dot(&yp);
        assert_mat(&tp, &g["s_tp"], i, "sum t'");
        let u = gs.iter().zip(ys.iter()).map(|(g, y)| ck.a2.dot(y).componentwise_mul(g)).reduce(|acc, x| acc.add(&x)).unwrap().sub(&ck.a2.dot(&yp));
        assert_mat(&u, &g["s_u"], i, "sum u");
        // sum.rs:188-197
        let zs: Vec<Mat<Z, N>> = ys.iter().zip(&rs).map(|(y, r)| y.add(&r.componentwise_mul(&d))).collect();
        let zp = yp.add(&rp.componentwise_mul(&d));
        for j in 0..t_terms {
            assert_mat(&zs[j], &g["s_zs"], 3 * (i * t_terms + j), "sum z_j");
        }
        assert_mat(&zp, &g["s_zp"], 3 * i, "sum z'");
        // sum.rs:262-319
        let split: Vec<(Mat<Z, N>, Mat<Z, N>)> = cs.iter().map(|c| Commitment { c: c.clone() }.c1_c2(&params)).collect();
        let (c1p, c2p) = Commitment { c: cp }.c1_c2(&params);
        let verify = |gs: &Vec<P>| {
            zs.iter().all(|z| params.check_verify_constraint(z))
                && params.check_verify_constraint(&zp)
                && zs.iter().zip(&split).zip(&ts).all(|((z, (c1, _)), t)| ck.a1.dot(z) == t.add(&c1.componentwise_mul(&d)))
                && ck.a1.dot(&zp) == tp.add(&c1p.componentwise_mul(&d))
                && zs.iter().zip(gs.iter()).map(|(z, g)| ck.a2.dot(z).componentwise_mul(g)).reduce(|acc, x| acc.add(&x)).unwrap().sub(&ck.a2.dot(&zp))
                    == split.iter().zip(gs.iter()).map(|((_, c2), g)| c2.componentwise_mul(g)).reduce(|acc, x| acc.add(&x)).unwrap()
                        .sub(&c2p).componentwise_mul(&d).add(&u)
        };
        assert_eq!(verify(&gs), g["s_ok"].v[i] == 1);
        let mut gb = gs.clone();
        let mut cg = coeffs(&gb[1]);
        cg[3] += 1; // recorded as s_bad
        gb[1] = Polynomial::new(cg.into_iter().map(Z::from).collect());
        assert_eq!(verify(&gb), g["s_bad"].v[i] == 1);
    }
}
