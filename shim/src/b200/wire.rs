//! The crate's own bincode encoding of its messages, produced and parsed for whole batches by the engine
//! (`rzk_wire_layout`, `rzk_wire_pack`, `rzk_wire_unpack`; `include/ringzk_b200.h`, "wire format of the messages").
//!
//! `pack(kind, terms, streams)` returns the concatenated messages and their offsets: `&bytes[off[i]..off[i + 1]]` is what
//! `bincode::serialize(&message_i)` returns for the struct the sequential code would have built, so the receiving side can
//! `bincode::deserialize::<OpenProofCommitment<_, N>>` it without the engine.  `unpack` is the inverse for a batch of received
//! messages (every length word and tag is checked; a malformed message clears its bit in the returned bitmap).
//! `TRIM` / `ELEM_BYTES` are the two facts about poly-ring-xnp1's serde impls that could not be checked where this was
//! written: trailing zero coefficients not stored (reproduces the 36 bytes of `mat.rs:434`) and 8-byte `ZqI64` coefficients.
//! Pin them once with `bincode::serialize(&Mat::from_vec(vec![poly]))` on a polynomial with a zero top coefficient.
//!
//! NOT COMPILED in the environment this was written in (no Rust toolchain there).

use std::os::raw::c_int;

use super::{ffi, B200Error, Backend};

pub const TRIM: c_int = 1;
pub const ELEM_BYTES: c_int = 8;

/// One stream of a message kind: a flat array `[B][polys_per_item][N]` of i32 (`dtype` 0) or i8 (`dtype` 1) coefficients,
/// numbered as the header lists them per message kind (e.g. `RZK_MSG_OPEN_COMMITMENT`: 0 = c [2], 1 = t [1]).
pub enum Stream<'a> {
    I32(&'a [i32], u32),
    I8(&'a [i8], u32),
}

fn layout(kind: c_int, terms: u32) -> Vec<ffi::RzkWireTok> {
    let mut n = 0usize;
    let rc = unsafe { ffi::rzk_wire_layout(kind, terms, std::ptr::null_mut(), 0, &mut n) };
    assert_eq!(rc, ffi::RZK_OK, "unknown message kind / number of terms");
    let mut toks = vec![ffi::RzkWireTok::default(); n];
    let rc = unsafe { ffi::rzk_wire_layout(kind, terms, toks.as_mut_ptr(), n, &mut n) };
    assert_eq!(rc, ffi::RZK_OK);
    toks
}

/// Messages of `kind` for the `b` items of `streams`: `(bytes, offsets)` with `offsets.len() == b + 1`.
pub fn pack(be: &mut Backend, kind: c_int, terms: u32, b: usize, streams: &[Stream]) -> Result<(Vec<u8>, Vec<u64>), B200Error> {
    let e = match *be {
        Backend::Engine(e) => e,
        Backend::Group(_) => return Err(B200Error::Unsupported("the wire-format entry points take a single engine".into())),
    };
    let toks = layout(kind, terms);
    let cs: Vec<ffi::RzkWireStream> = streams
        .iter()
        .map(|s| match s {
            Stream::I32(v, polys) => ffi::RzkWireStream { base: v.as_ptr() as *const _, polys_per_item: *polys, dtype: 0 },
            Stream::I8(v, polys) => ffi::RzkWireStream { base: v.as_ptr() as *const _, polys_per_item: *polys, dtype: 1 },
        })
        .collect();
    let mut off = vec![0u64; b + 1];
    let mut total = 0u64;
    let rc = unsafe {
        ffi::rzk_wire_pack(e, b, toks.as_ptr(), toks.len(), cs.as_ptr(), cs.len() as c_int, ELEM_BYTES, TRIM, std::ptr::null_mut(), 0, off.as_mut_ptr(), &mut total)
    };
    be.check_or_panic(rc)?;
    let mut out = vec![0u8; total as usize];
    let rc = unsafe {
        ffi::rzk_wire_pack(e, b, toks.as_ptr(), toks.len(), cs.as_ptr(), cs.len() as c_int, ELEM_BYTES, TRIM, out.as_mut_ptr(), out.len(), off.as_mut_ptr(), &mut total)
    };
    be.check_or_panic(rc)?;
    Ok((out, off))
}

/// Parses `b` messages of `kind` into the flat arrays `outs` (same numbering and shapes as for `pack`; every array is
/// overwritten, polynomials zero padded to N, coefficients canonicalised like `ZqI64::from`).  Bit i of the result is set
/// when message i was well formed.
pub fn unpack(
    be: &mut Backend,
    kind: c_int,
    terms: u32,
    bytes: &[u8],
    offsets: &[u64],
    outs: &mut [(&mut [i32], u32)],
    outs_i8: &mut [(&mut [i8], u32)],
    order: &[bool], // order[s] == true: stream s is the next i8 array, false: the next i32 array
) -> Result<Vec<u8>, B200Error> {
    let e = match *be {
        Backend::Engine(e) => e,
        Backend::Group(_) => return Err(B200Error::Unsupported("the wire-format entry points take a single engine".into())),
    };
    let b = offsets.len() - 1;
    let toks = layout(kind, terms);
    let (mut i32s, mut i8s) = (outs.iter_mut(), outs_i8.iter_mut());
    let cs: Vec<ffi::RzkWireStream> = order
        .iter()
        .map(|&small| {
            if small {
                let (v, polys) = i8s.next().expect("fewer i8 arrays than `order` names");
                ffi::RzkWireStream { base: v.as_mut_ptr() as *const _, polys_per_item: *polys, dtype: 1 }
            } else {
                let (v, polys) = i32s.next().expect("fewer i32 arrays than `order` names");
                ffi::RzkWireStream { base: v.as_mut_ptr() as *const _, polys_per_item: *polys, dtype: 0 }
            }
        })
        .collect();
    let mut ok = vec![0u8; (b + 7) / 8];
    let rc = unsafe {
        ffi::rzk_wire_unpack(e, b, toks.as_ptr(), toks.len(), cs.as_ptr(), cs.len() as c_int, ELEM_BYTES, bytes.as_ptr(), bytes.len(), offsets.as_ptr(), ok.as_mut_ptr())
    };
    be.check_or_panic(rc)?;
    Ok(ok)
}
