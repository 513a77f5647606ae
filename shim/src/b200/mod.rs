//! B200 engine binding for ring-zk (feature `b200`).
//!
//! This module is ADDED to the crate; nothing of the existing public surface (`src/lib.rs:5-24`) changes.  It holds
//!   * `ffi`      -- the `extern "C"` block for every export of `libringzk_b200.so` (generated from
//!                   `include/ringzk_b200.h` by `tools/gen_rust_ffi.py`),
//!   * `Backend`  -- a safe owner of one engine (`rzk_create`) or of a device group (`rzk_group_create`),
//!   * the flatten / unflatten helpers between `Polynomial<ZqI64<Q>, N>` and the flat `i32` / `i8` arrays of the C ABI.
//! The batched protocol methods live next to the types they extend, as child modules that can reach the private fields
//! (`src/commit/batch.rs`, `src/prove/{open,linear,sum}/batch.rs`).
//!
//! Accelerated instantiation: `Params::default()`-shaped parameters (`(n, k, l) = (1, 3, 1)`, `q = 3515337053`,
//! `b * kappa <= 74`) at `N = 512`.  `Backend::new` returns `Err(B200Error::Unsupported)` for anything else and the caller
//! keeps using the sequential CPU methods -- in particular the crate's own integration tests (`tests/test.rs:8`, `N = 16`,
//! challenges with `min(kappa, N) = 16` non-zeros) and every doctest run on the generic CPU code, untouched.
//!
//! NOT COMPILED in the environment this was written in (no Rust toolchain there); it follows the header one to one.

pub mod ffi;
pub mod wire;

use std::ffi::CStr;
use std::os::raw::c_int;

use poly_ring_xnp1::{zq::ZqI64, Polynomial};

use rand::RngExt;

use crate::{mat::Mat, polynomial::random_polynomial_in_normal_distribution, CommitmentKey, Params};

/// The modulus of the accelerated coefficient type (`src/params.rs:121,126`).
pub const Q: i64 = 3515337053;
/// The accelerated coefficient type.
pub type Z = ZqI64<Q>;

/// A non-zero status of the C ABI (`include/ringzk_b200.h`, `RZK_ERR_*`) with the engine's message.
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum B200Error {
    /// `RZK_ERR_INVALID`: a shape the reference itself rejects by `assert!` (`params.rs:71`, `commit.rs:95`, `sum.rs:105`).
    Invalid(String),
    /// `RZK_ERR_UNSUPPORTED`: parameters, ring degree or key structure outside the accelerated instantiation: use the CPU path.
    Unsupported(String),
    /// `RZK_ERR_CUDA`: no device / CUDA runtime failure.  There is no CPU fallback inside the engine.
    Cuda(String),
    /// `RZK_ERR_RANGE`: a masking coefficient left `rzk_small_limit()` (>= 10 sigma): repeat the call on the CPU path.
    Range(String),
    /// `RZK_ERR_NOKEY`
    NoKey,
}

impl std::fmt::Display for B200Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{:?}", self)
    }
}
impl std::error::Error for B200Error {}

/// One engine on one device, or one engine per listed device behind a group handle (`INTEGRATION.md` section 5).
/// Externally synchronised (`&mut self` on every call), `Send` but not `Sync`.
pub enum Backend {
    Engine(*mut ffi::RzkEngine),
    Group(*mut ffi::RzkGroup),
}
unsafe impl Send for Backend {}

impl Drop for Backend {
    fn drop(&mut self) {
        unsafe {
            match *self {
                Backend::Engine(e) => ffi::rzk_destroy(e),
                Backend::Group(g) => ffi::rzk_group_destroy(g),
            }
        }
    }
}

fn rzk_params<const N: usize>(params: &Params<Z>) -> ffi::RzkParams {
    ffi::RzkParams {
        q: Q,
        b: params.b.clone().into(),
        n_ring: N as i32,
        n: params.n as i32,
        k: params.k as i32,
        l: params.l as i32,
        kappa: params.kappa as i32,
    }
}

impl Backend {
    /// `rzk_create` on `device` (`-1`: the current CUDA device) + `rzk_set_key`.
    pub fn new<const N: usize>(ck: &CommitmentKey<Z, N>, params: &Params<Z>, device: i32) -> Result<Backend, B200Error> {
        let p = rzk_params::<N>(params);
        let mut raw: *mut ffi::RzkEngine = std::ptr::null_mut();
        let rc = unsafe { ffi::rzk_create(&p, device as c_int, &mut raw) };
        if rc != ffi::RZK_OK {
            return Err(status(rc, unsafe { ffi::rzk_last_error(std::ptr::null()) }));
        }
        let mut be = Backend::Engine(raw);
        be.set_key(ck)?;
        Ok(be)
    }

    /// `rzk_group_create` over `devices` + `rzk_group_set_key`: the batch of every call is split into contiguous item
    /// ranges, one per device, with no exchange between devices.
    pub fn new_group<const N: usize>(ck: &CommitmentKey<Z, N>, params: &Params<Z>, devices: &[i32]) -> Result<Backend, B200Error> {
        let p = rzk_params::<N>(params);
        let ids: Vec<c_int> = devices.iter().map(|&d| d as c_int).collect();
        let mut raw: *mut ffi::RzkGroup = std::ptr::null_mut();
        let rc = unsafe { ffi::rzk_group_create(&p, ids.as_ptr(), ids.len() as c_int, &mut raw) };
        if rc != ffi::RZK_OK {
            return Err(status(rc, unsafe { ffi::rzk_group_last_error(std::ptr::null()) }));
        }
        let mut be = Backend::Group(raw);
        be.set_key(ck)?;
        Ok(be)
    }

    /// CommitmentKey (commit.rs:19-60) -> `a1 [n][k][N]`, `a2 [l][k][N]` as `i64`, through `Into<i64>` (never a transmute:
    /// the layout of `ZqI64` is not part of poly-ring-xnp1's contract).
    fn set_key<const N: usize>(&mut self, ck: &CommitmentKey<Z, N>) -> Result<(), B200Error> {
        let flat = |m: &Mat<Z, N>| -> Vec<i64> {
            let mut out = Vec::with_capacity(m.polynomials.len() * m.polynomials.get(0).map_or(0, |r| r.len()) * N);
            for row in &m.polynomials {
                for p in row {
                    let mut c: Vec<i64> = p.iter().map(|v| v.clone().into()).collect();
                    c.resize(N, 0); // Polynomial stores only the given coefficients (mat.rs:424-438): pad with zeros
                    out.extend_from_slice(&c);
                }
            }
            out
        };
        let (a1, a2) = (flat(&ck.a1), flat(&ck.a2));
        let rc = unsafe {
            match *self {
                Backend::Engine(e) => ffi::rzk_set_key(e, a1.as_ptr(), a2.as_ptr()),
                Backend::Group(g) => ffi::rzk_group_set_key(g, a1.as_ptr(), a2.as_ptr()),
            }
        };
        self.check(rc)
    }

    /// Turns a status into `Ok(())` / `Err(..)` with the engine's message.
    pub(crate) fn check(&self, rc: c_int) -> Result<(), B200Error> {
        if rc == ffi::RZK_OK {
            return Ok(());
        }
        let msg = unsafe {
            match *self {
                Backend::Engine(e) => ffi::rzk_last_error(e),
                Backend::Group(g) => ffi::rzk_group_last_error(g),
            }
        };
        Err(status(rc, msg))
    }

    /// Same, with the reference's error convention: the conditions the sequential methods `assert!` on panic here too;
    /// everything else is returned for the caller to fall back to the CPU path.
    pub(crate) fn check_or_panic(&self, rc: c_int) -> Result<(), B200Error> {
        match self.check(rc) {
            Err(B200Error::Invalid(m)) => panic!("ring-zk b200: {}", m),
            other => other,
        }
    }
}

fn status(rc: c_int, msg: *const std::os::raw::c_char) -> B200Error {
    let m = if msg.is_null() { String::new() } else { unsafe { CStr::from_ptr(msg) }.to_string_lossy().into_owned() };
    match rc {
        ffi::RZK_ERR_INVALID => B200Error::Invalid(m),
        ffi::RZK_ERR_UNSUPPORTED => B200Error::Unsupported(m),
        ffi::RZK_ERR_RANGE => B200Error::Range(m),
        ffi::RZK_ERR_NOKEY => B200Error::NoKey,
        _ => B200Error::Cuda(m),
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Polynomial <-> flat arrays.  Coefficient i at index i, zero padded to N; the canonical centred residue of ZqI64<Q>
// (|v| <= (Q - 1) / 2 < 2^31) fits i32.

/// Appends the N coefficients of `p` to `out` as i32.
pub(crate) fn push_poly<const N: usize>(out: &mut Vec<i32>, p: &Polynomial<Z, N>) {
    let start = out.len();
    out.extend(p.iter().map(|c| Into::<i64>::into(c.clone()) as i32));
    out.resize(start + N, 0);
}

/// Appends the N coefficients of a small polynomial (the randomness r, |r| <= b <= 127, or the challenge d) as i8.
/// Panics on a coefficient outside i8: such a polynomial is not something `random_polynomial_within(b)` or
/// `random_polynomial_from_challenge_set` can produce.
pub(crate) fn push_poly_i8<const N: usize>(out: &mut Vec<i8>, p: &Polynomial<Z, N>) {
    let start = out.len();
    out.extend(p.iter().map(|c| {
        let v: i64 = c.clone().into();
        assert!((-128..=127).contains(&v), "small polynomial with a coefficient outside i8");
        v as i8
    }));
    out.resize(start + N, 0);
}

/// Appends every polynomial of a (rows x 1) matrix.
pub(crate) fn push_mat<const N: usize>(out: &mut Vec<i32>, m: &Mat<Z, N>) {
    for row in &m.polynomials {
        for p in row {
            push_poly(out, p);
        }
    }
}
pub(crate) fn push_mat_i8<const N: usize>(out: &mut Vec<i8>, m: &Mat<Z, N>) {
    for row in &m.polynomials {
        for p in row {
            push_poly_i8(out, p);
        }
    }
}

/// The 2-bit packing of `rzk_commit_batch_r2` (two's-complement fields, four coefficients per byte, low bits first) for
/// randomness with every entry in [-2, 1] -- `Params::default()` draws r from {-1, 0, 1}.  `None` if an entry does not fit.
pub(crate) fn pack_r2(r: &[i8]) -> Option<Vec<u8>> {
    debug_assert_eq!(r.len() % 4, 0);
    let mut out = Vec::with_capacity(r.len() / 4);
    for quad in r.chunks_exact(4) {
        let mut b = 0u8;
        for (k, &v) in quad.iter().enumerate() {
            if !(-2..=1).contains(&v) {
                return None;
            }
            b |= ((v as u8) & 3) << (2 * k);
        }
        out.push(b);
    }
    Some(out)
}

/// N coefficients -> Polynomial (full length; `Polynomial ==` compares values, padding zeros included on both sides
/// of every comparison the protocols make, since every engine output is full length).
pub(crate) fn poly_from<const N: usize>(v: &[i32]) -> Polynomial<Z, N> {
    debug_assert_eq!(v.len(), N);
    Polynomial::new(v.iter().map(|&c| Z::from(c as i64)).collect())
}

/// `rows * N` coefficients -> (rows x 1) matrix.
pub(crate) fn mat_from<const N: usize>(v: &[i32], rows: usize) -> Mat<Z, N> {
    debug_assert_eq!(v.len(), rows * N);
    Mat::from_vec((0..rows).map(|i| poly_from::<N>(&v[i * N..(i + 1) * N])).collect())
}

/// `rows * N` coefficients -> Vec of polynomials.
pub(crate) fn polys_from<const N: usize>(v: &[i32], rows: usize) -> Vec<Polynomial<Z, N>> {
    (0..rows).map(|i| poly_from::<N>(&v[i * N..(i + 1) * N])).collect()
}

/// One masking vector `y <- N^k_sigma`, drawn exactly as the provers draw it (`open.rs:88-94`, `linear.rs:100-115`,
/// `sum.rs:123-142`): k polynomials, N normal samples each, truncated by `I::from_f64`.
pub(crate) fn draw_masking<const N: usize>(rng: &mut impl RngExt, params: &Params<Z>) -> Mat<Z, N> {
    let sigma = params.standard_deviation(N) as f64;
    Mat::<Z, N>::new_with(params.k, 1, || random_polynomial_in_normal_distribution::<Z, N>(rng, 0.0, sigma))
}

/// Bit i of an ok / verify bitmap (bit `i & 7` of byte `i >> 3`).
#[inline]
pub(crate) fn bit(bitmap: &[u8], i: usize) -> bool {
    (bitmap[i >> 3] >> (i & 7)) & 1 == 1
}

/// The engine's shapes: `(n, k, l) = (1, 3, 1)`.  The batch methods assert them so that a mismatch is a loud panic, as
/// the reference's own shape asserts are.
pub(crate) fn assert_default_shape(params: &Params<Z>) {
    assert!(params.n == 1 && params.k == 3 && params.l == 1, "the B200 engine accelerates (n, k, l) = (1, 3, 1) only");
}

/// Transcript prefix of the Fiat-Shamir entry points (docs/FIAT_SHAMIR.md of the engine's repository):
/// `tag` (ASCII, zero padded to 32 bytes) || key digest (32 bytes: SHAKE128-256 of a11 || a12 || a22 as int32 LE canonical
/// coefficients, computed by the caller with the hash crate of its choice) || q u64 || N u32 || kappa u32 || T u32 || b u32 ||
/// session (a multiple of 8 bytes), all little endian.
pub fn fs_prefix<const N: usize>(tag: &str, key_digest: &[u8; 32], params: &Params<Z>, terms: u32, session: &[u8]) -> Vec<u8> {
    assert!(tag.len() <= 32 && session.len() % 8 == 0);
    let mut p = vec![0u8; 32];
    p[..tag.len()].copy_from_slice(tag.as_bytes());
    p.extend_from_slice(key_digest);
    p.extend_from_slice(&(Q as u64).to_le_bytes());
    p.extend_from_slice(&(N as u32).to_le_bytes());
    p.extend_from_slice(&(params.kappa as u32).to_le_bytes());
    p.extend_from_slice(&terms.to_le_bytes());
    let b: i64 = params.b.clone().into();
    p.extend_from_slice(&(b as u32).to_le_bytes());
    p.extend_from_slice(session);
    p
}
