// rzk_tables.h -- host-side construction of NTT twiddle tables, CRT constants and the
// NTT-domain image of the commitment key (CommitmentKey, /root/reference/src/commit.rs:19-60).
#pragma once
#include <stdint.h>
#include <vector>
#include "rzk_vm.h"

namespace rzk {

// Global static list of auxiliary NTT primes, all == 1 (mod 4096):
//   slots 0..2 : the three largest below 2^30   (Harvey lazy range [0,4p) fits 32 bits)
//   slots 3..5 : the three smallest primes above 2^26 for the signed lazy arithmetic of rzk_arith.cuh; their kernel-side tables (g1, g2,
//                key images) hold centred values with signed Shoup companions round(w * 2^32 / p).  Slot 3 is the compile-time
//                prime of the b = 1 split-key commitment program (kStaticPrimeS, MODE_SPLITKEY_S)
extern const uint32_t kPrimeList[kNumPrimeSlots];
bool slot_is_signed(int slot);
void signed_shoup_pair(uint32_t w, uint32_t p, uint32_t &w_centred, uint32_t &w_companion);

uint32_t mod_pow(uint32_t b, uint64_t e, uint32_t p);
uint32_t mod_inv(uint32_t a, uint32_t p);
uint32_t shoup_companion(uint32_t w, uint32_t p);

// Builds (once, lazily, thread-safe) and returns the tables of prime slot s.
const PrimeTables &prime_tables(int slot);

// Exact reference transforms on fully reduced residues (key setup and tests).
void ntt_forward_ref(const PrimeTables &T, uint32_t a[kN]);
void ntt_inverse_ref(const PrimeTables &T, uint32_t a[kN]);   // includes the N^-1 scaling

// Fills PrimeC for a prime slot.
PrimeC make_prime_consts(int slot);
// CRT / Garner constants for an ordered prime set (np = 1..3) and modulus q.
CrtC make_crt_consts(const int *slots, int np, uint64_t q);

// Key image for one prime: N^-1 * NTT_p(poly) with Shoup companions, in the padded
// lane layout read by OP_MACK.  out: [2][kPadWords] (w row then w' row).
void key_image(const PrimeTables &T, const int64_t *poly_centered, uint32_t *out);

// Split-key images (MODE_SPLITKEY): poly = lo + 2^16 * hi with lo in [-2^15, 2^15); writes the image of
// lo to out[0 .. 2*kPadWords) and of hi to out[2*kPadWords .. 4*kPadWords).
void key_image_split(const PrimeTables &T, const int64_t *poly_centered, uint32_t *out);

// The same for a signed slot (MODE_SPLITKEY_S): centred residues, signed companions.
void key_image_split_signed(const PrimeTables &T, const int64_t *poly_centered, uint32_t *out);

inline int pad_index(int i) { return i + ((i >> 5) << 2); }

}  // namespace rzk
