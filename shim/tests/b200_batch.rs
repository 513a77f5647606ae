//! Integration test of the batched entry points (feature `b200`, needs a B200 and `libringzk_b200.so`):
//!
//!     RINGZK_B200_LIB_DIR=/path/to/ring-zk_b200/_build cargo test --features b200 --test b200_batch
//!
//! Every `*_batch` method must equal its sequential twin called once per element on a clone of the same seeded RNG
//! (same draw order, SURVEY section 3 "RNG draw order"), message for message -- `PartialEq` on the message structs
//! compares every polynomial -- and the verifier must accept the honest batch and reject a tampered element.
//! NOT COMPILED where it was written (no Rust toolchain there).
#![cfg(feature = "b200")]

use rand::{rngs::StdRng, SeedableRng};
use ring_zk::b200::{Backend, Z};
use ring_zk::{LinearProofProver, LinearProofVerifier, OpenProofProver, OpenProofVerifier, Params, SumProofProver, SumProofVerifier};

const N: usize = 512;
const B: usize = 9;

fn values(params: &Params<Z>, seed: i64, b: usize) -> Vec<Vec<poly_ring_xnp1::Polynomial<Z, N>>> {
    (0..b).map(|i| params.prepare_value::<N>(vec![vec![seed + i as i64, 2, 3, -(i as i64)]])).collect()
}

#[test]
fn open_proof_batch_equals_sequential() {
    let params = Params::default();
    let ck = params.generate_commitment_key::<N>(&mut StdRng::seed_from_u64(1));
    let mut be = Backend::new(&ck, &params, -1).expect("a B200 and libringzk_b200.so");
    let (prover, verifier) = (OpenProofProver::new(ck.clone(), params.clone()), OpenProofVerifier::new(ck.clone(), params.clone()));
    let xs = values(&params, 7, B);

    let (mut rng_a, mut rng_b) = (StdRng::seed_from_u64(2), StdRng::seed_from_u64(2));
    let seq: Vec<_> = xs.iter().cloned().map(|x| prover.commit(&mut rng_a, x)).collect();
    let bat = prover.commit_batch(&mut rng_b, xs, &mut be).unwrap();
    assert_eq!(seq.len(), bat.len());
    for ((ctx_s, com_s), (ctx_b, com_b)) in seq.iter().zip(&bat) {
        assert_eq!(ctx_s, ctx_b); // opening (x, r) and y: same RNG stream, same order
        assert_eq!(com_s, com_b); // c and t: the engine's ring arithmetic equals the crate's
        assert!(com_b.c.verify(&ctx_b.opening, &ck, &params));
    }
    let (ctxs, coms): (Vec<_>, Vec<_>) = bat.into_iter().unzip();
    let (mut rng_a, mut rng_b) = (StdRng::seed_from_u64(3), StdRng::seed_from_u64(3));
    let ch_seq: Vec<_> = coms.iter().cloned().map(|c| verifier.generate_challenge(&mut rng_a, c)).collect();
    let ch_bat = verifier.generate_challenge_batch(&mut rng_b, coms);
    assert_eq!(ch_seq, ch_bat);
    let (vctxs, challenges): (Vec<_>, Vec<_>) = ch_bat.into_iter().unzip();
    let resp_seq: Vec<_> = ctxs.iter().cloned().zip(challenges.iter().cloned()).map(|(c, d)| prover.create_response(c, d)).collect();
    let resp_bat = prover.create_response_batch(ctxs, challenges, &mut be).unwrap();
    assert_eq!(resp_seq, resp_bat);
    let ok = verifier.verify_batch(resp_bat.clone(), vctxs.clone(), &mut be).unwrap();
    assert!(ok.iter().all(|&v| v));
    // responses shifted by one instance: every equation fails
    let mut shifted = resp_bat;
    shifted.rotate_left(1);
    let bad = verifier.verify_batch(shifted.clone(), vctxs.clone(), &mut be).unwrap();
    let bad_seq: Vec<bool> = shifted.into_iter().zip(vctxs).map(|(r, c)| verifier.verify(r, c)).collect();
    assert_eq!(bad, bad_seq);
    assert!(bad.iter().all(|&v| !v));
}

#[test]
fn linear_proof_batch_equals_sequential() {
    let params = Params::default();
    let ck = params.generate_commitment_key::<N>(&mut StdRng::seed_from_u64(11));
    let mut be = Backend::new(&ck, &params, -1).expect("a B200 and libringzk_b200.so");
    let (prover, verifier) = (LinearProofProver::new(ck.clone(), params.clone()), LinearProofVerifier::new(ck.clone(), params.clone()));
    let xs = values(&params, 100, B);
    let gs: Vec<_> = (0..B).map(|i| params.prepare_scalar::<N>(vec![5 + i as i64, -1757668526, 1757668526, 9])).collect();

    let (mut rng_a, mut rng_b) = (StdRng::seed_from_u64(12), StdRng::seed_from_u64(12));
    let seq: Vec<_> = gs.iter().cloned().zip(xs.iter().cloned()).map(|(g, x)| prover.commit(&mut rng_a, g, x)).collect();
    let bat = prover.commit_batch(&mut rng_b, gs, xs, &mut be).unwrap();
    for (s, b) in seq.iter().zip(&bat) {
        assert_eq!(s, b);
    }
    let (ctxs, coms): (Vec<_>, Vec<_>) = bat.into_iter().unzip();
    let ch = verifier.generate_challenge_batch(&mut StdRng::seed_from_u64(13), coms);
    let (vctxs, challenges): (Vec<_>, Vec<_>) = ch.into_iter().unzip();
    let resp_seq: Vec<_> = ctxs.iter().cloned().zip(challenges.iter().cloned()).map(|(c, d)| prover.create_response(c, d)).collect();
    let resp_bat = prover.create_response_batch(ctxs, challenges, &mut be).unwrap();
    assert_eq!(resp_seq, resp_bat);
    assert!(verifier.verify_batch(resp_bat.clone(), vctxs.clone(), &mut be).unwrap().iter().all(|&v| v));
    let mut shifted = resp_bat;
    shifted.rotate_left(1);
    assert!(verifier.verify_batch(shifted, vctxs, &mut be).unwrap().iter().all(|&v| !v));
}

#[test]
fn sum_proof_batch_equals_sequential() {
    const T: usize = 4; // tests/test.rs:65, benches/bench.rs:200
    let params = Params::default();
    let ck = params.generate_commitment_key::<N>(&mut StdRng::seed_from_u64(21));
    let mut be = Backend::new(&ck, &params, -1).expect("a B200 and libringzk_b200.so");
    let (prover, verifier) = (SumProofProver::new(ck.clone(), params.clone()), SumProofVerifier::new(ck.clone(), params.clone()));
    let xss: Vec<Vec<_>> = (0..B).map(|i| values(&params, 1000 * i as i64, T)).collect();
    let gss: Vec<Vec<_>> = (0..B).map(|i| (0..T).map(|j| params.prepare_scalar::<N>(vec![(i * T + j) as i64 + 1, 1757668526])).collect()).collect();

    let (mut rng_a, mut rng_b) = (StdRng::seed_from_u64(22), StdRng::seed_from_u64(22));
    let seq: Vec<_> = gss.iter().cloned().zip(xss.iter().cloned()).map(|(g, x)| prover.commit(&mut rng_a, g, x)).collect();
    let bat = prover.commit_batch(&mut rng_b, gss, xss, &mut be).unwrap();
    for (s, b) in seq.iter().zip(&bat) {
        assert_eq!(s, b);
    }
    let (ctxs, coms): (Vec<_>, Vec<_>) = bat.into_iter().unzip();
    let ch = verifier.generate_challenge_batch(&mut StdRng::seed_from_u64(23), coms);
    let (vctxs, challenges): (Vec<_>, Vec<_>) = ch.into_iter().unzip();
    let resp_seq: Vec<_> = ctxs.iter().cloned().zip(challenges.iter().cloned()).map(|(c, d)| prover.create_response(c, d)).collect();
    let resp_bat = prover.create_response_batch(ctxs, challenges, &mut be).unwrap();
    assert_eq!(resp_seq, resp_bat);
    assert!(verifier.verify_batch(resp_bat, vctxs, &mut be).unwrap().iter().all(|&v| v));
}

#[test]
fn unsupported_parameters_keep_the_cpu_path() {
    // N = 16 (the crate's own integration tests, tests/test.rs:8): Backend::new declines and the caller keeps using the
    // sequential methods, which run the generic CPU code exactly as before
    let params = Params::default();
    let ck = params.generate_commitment_key::<16>(&mut StdRng::seed_from_u64(5));
    assert!(matches!(Backend::new(&ck, &params, -1), Err(ring_zk::b200::B200Error::Unsupported(_))));
}

#[test]
fn non_interactive_open_proofs_and_the_wire_format() {
    // the extension of docs/FIAT_SHAMIR.md: a proof is (commitment, response); the challenge comes from the transcript.
    // Cross-checked against the interactive methods: with the challenge the engine derived, the sequential verifier accepts.
    use ring_zk::b200::fs_prefix;
    let params = Params::default();
    let ck = params.generate_commitment_key::<N>(&mut StdRng::seed_from_u64(31));
    let mut be = Backend::new(&ck, &params, -1).expect("a B200 and libringzk_b200.so");
    let (prover, verifier) = (OpenProofProver::new(ck.clone(), params.clone()), OpenProofVerifier::new(ck.clone(), params.clone()));
    let digest = [7u8; 32]; // a real caller hashes the key (SHAKE128-256 of a11 || a12 || a22, int32 LE); any fixed value binds this test
    let prefix = fs_prefix::<N>("ring-zk/fs/open/v1", &digest, &params, 0, b"session1");
    let proofs = prover.prove_batch_fs(&mut StdRng::seed_from_u64(32), values(&params, 3, B), &prefix, &mut be).unwrap();
    let pairs: Vec<_> = proofs.iter().map(|(_, c, z)| (c.clone(), z.clone())).collect();
    assert!(verifier.verify_batch_fs(&pairs, &prefix, &mut be).unwrap().iter().all(|&v| v));
    let other = fs_prefix::<N>("ring-zk/fs/open/v1", &digest, &params, 0, b"session2");
    assert!(verifier.verify_batch_fs(&pairs, &other, &mut be).unwrap().iter().all(|&v| !v));
    // the commitments cross the wire in the crate's own encoding: what the engine packs is what bincode produces
    let coms: Vec<_> = pairs.iter().map(|(c, _)| c.clone()).collect();
    let (bytes, off) = ring_zk::OpenProofCommitment::to_wire_batch(&coms, &mut be).unwrap();
    for (i, c) in coms.iter().enumerate() {
        assert_eq!(&bytes[off[i] as usize..off[i + 1] as usize], &bincode::serialize(c).unwrap()[..]);
    }
}
