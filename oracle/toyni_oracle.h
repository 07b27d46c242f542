/*
 * toyni_oracle.h — CPU restatement of the jonas089/toyni prover hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under toyni_b200/ (the product) may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker and CPU baseline.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference tree).  All field elements use the reference's storage: one canonical
 * value in [0,p) per uint64_t (src/babybear.rs:10-14); an Ext is four of them in
 * limb order [a0,a1,a2,a3] (src/ext.rs:22-26).
 *
 * Parity pinning: checked against every value-level known answer the reference's
 * own tests hold for this path (src/babybear.rs:221-254, src/ntt.rs:339-357) and
 * all of its identity tests (tests/test_oracle_reference_tests.py), against an
 * independent pure-Python restatement (oracle/pyref.py), against SURVEY Appendix A,
 * and — on the GPU box — against the reference's own CUDA implementation compiled
 * from /root/reference/cuda/ntt_kernel.cu into oracle/_ref/ (NTT / INTT outputs).
 * SHA-256 is FIPS 180-4 (the `sha2 0.10.8` crate, Cargo.lock:133-142, not vendored);
 * it is pinned against Python hashlib.  fri_fold_ext has no test and no caller in
 * the reference: its parity is UNPINNED by reference fixtures and follows
 * src/math/fri.rs:7-25 literally.
 */
#ifndef TOYNI_ORACLE_H
#define TOYNI_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TO_P 2013265921ULL /* src/babybear.rs:8 */

/* ---- BabyBear (src/babybear.rs) ---- */
uint64_t to_bb_new(uint64_t v);                 /* :26-30 */
uint64_t to_bb_add(uint64_t a, uint64_t b);     /* :80-89,133-138 */
uint64_t to_bb_sub(uint64_t a, uint64_t b);     /* :151-158 */
uint64_t to_bb_mul(uint64_t a, uint64_t b);     /* :173-177 */
uint64_t to_bb_neg(uint64_t a);                 /* :197-205 */
uint64_t to_bb_pow(uint64_t a, uint64_t e);     /* :91-108 */
uint64_t to_bb_inverse(uint64_t a);             /* :111-114 (returns 0 for 0; the reference panics) */
uint64_t to_bb_root_of_unity(uint32_t log_n);   /* :118-126 */

/* ---- Ext = F_p[X]/(X^4-11) (src/ext.rs) ---- */
void to_ext_add(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);       /* :138-146 */
void to_ext_sub(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);       /* :154-162 */
void to_ext_mul(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);       /* :178-192 */
void to_ext_mul_base(const uint64_t a[4], uint64_t s, uint64_t r[4]);           /* :76-78 */
void to_ext_inverse(const uint64_t a[4], uint64_t r[4]);                        /* :107-127 */

/* ---- NTT (src/ntt.rs) ---- */
void to_ntt(uint64_t* values, size_t n, uint64_t omega);      /* :24-53 */
void to_intt(uint64_t* values, size_t n, uint64_t omega);     /* :56-66 */
void to_roots_of_unity_domain(uint64_t* out, size_t n);       /* :69-81 */
/* Same arithmetic and the same stage order as to_ntt, with the independent butterfly
 * groups of each stage spread over `threads` OpenMP threads (bit-identical output).
 * The reference itself is single-threaded; this is the all-cores CPU baseline. */
void to_ntt_mt(uint64_t* values, size_t n, uint64_t omega, int threads);
void to_intt_mt(uint64_t* values, size_t n, uint64_t omega, int threads);

/* ---- Evaluation domain (src/math/domain.rs) ---- */
void to_domain_elements(uint64_t* out, size_t size, uint64_t shift);                          /* :61-69 */
/* fft: zero-pad / truncate coeffs to `size`, coset shift, forward NTT (:107-123,154-162) */
/* thread count of the element-wise loops and of the transforms inside to_domain_fft / ifft (1 = the reference's serial
 * loops verbatim; more = the same arithmetic cut into chunks, same bits) */
void to_set_threads(int threads);
int to_get_threads(void);
void to_domain_fft(const uint64_t* coeffs, size_t ncoeffs, size_t size, uint64_t shift, uint64_t* out);
/* ifft: INTT then undo the coset shift (:85-102,165-174); evals has exactly `size` entries */
void to_domain_ifft(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* out);
/* Ext transforms = four base transforms (:129-151); AoS, 4 limbs per element */
void to_domain_fft_ext(const uint64_t* coeffs, size_t ncoeffs, size_t size, uint64_t shift, uint64_t* out);
void to_domain_ifft_ext(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* out);

/* ---- FRI fold (src/math/fri.rs) ---- */
void to_fri_fold(const uint64_t* evals, size_t m, const uint64_t* xs, uint64_t beta, uint64_t* out);          /* :27-48 */
void to_fri_fold_ext(const uint64_t* evals, size_t m, const uint64_t* xs, const uint64_t beta[4], uint64_t* out); /* :7-25 */

/* ---- SHA-256 (FIPS 180-4; the sha2 crate behind src/lib.rs:14-18, src/merkle.rs:109-123) ---- */
void to_sha256(const uint8_t* data, size_t len, uint8_t out[32]);

/* ---- Merkle tree (src/merkle.rs) ---- */
void to_hash_leaf(const uint8_t* leaf, size_t len, uint8_t out[32]);            /* :109-114 */
void to_hash_node(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]);   /* :117-123 */
/* Total digests over all levels for `nleaves` leaves (odd levels duplicate the last node, :36-43). */
size_t to_merkle_node_count(size_t nleaves);
/* build_tree (:25-48): leaves are `nleaves` byte strings of `leaf_len` bytes each, back to back.
 * nodes_out receives every level, leaf level first, 32 bytes per digest; root_out the root (:82-84). */
void to_merkle_build(const uint8_t* leaves, size_t nleaves, size_t leaf_len, uint8_t* nodes_out, uint8_t root_out[32]);
/* Prover leaf encodings (src/fibonacci.rs:340-363): salted = salt[16] || LE-u64(value) (24 B);
 * unsalted = LE-u64(value) (8 B).  salts == NULL selects the unsalted form.  `limbs` is 1 for
 * base-field values and 4 for the Ext analogue (salt || 4 x LE-u64, src/ext.rs:83-89). */
void to_commit_values(const uint64_t* values, size_t n, int limbs, const uint8_t* salts, uint8_t* nodes_out,
                      uint8_t root_out[32]);
/* get_proof (:50-80) over a nodes array made by to_merkle_build: writes depth 32-byte siblings and
 * depth position flags, returns depth. */
size_t to_merkle_open(const uint8_t* nodes, size_t nleaves, size_t index, uint8_t* path_out, uint8_t* pos_out);
int to_merkle_verify(const uint8_t* leaf, size_t leaf_len, const uint8_t* path, const uint8_t* pos, size_t depth,
                     const uint8_t root[32]); /* :86-101 */

/* ---- Fiat-Shamir transcript (src/transcript.rs) ---- */
typedef struct {
    uint8_t* state;
    size_t len, cap;
} to_transcript;
void to_transcript_init(to_transcript* t);                                   /* :12-16 */
void to_transcript_free(to_transcript* t);
void to_transcript_absorb(to_transcript* t, const uint8_t* data, size_t n);  /* :19-21 */
uint64_t to_transcript_squeeze(to_transcript* t);                            /* :34-39 */
void to_transcript_squeeze_ext(to_transcript* t, uint64_t out[4]);           /* :43-50 */
void to_transcript_squeeze_indices(to_transcript* t, size_t count, size_t max, uint64_t* out); /* :58-72 */

/* ---- FRI commit loop of the prover (src/fibonacci.rs:200-247), randomness explicit ----
 * layer0: `n` base-field evaluations on the coset {shift*omega_n^i}.  Folds until the layer has
 * `final_size` entries.  salts: 16 bytes per leaf for every salted layer, layers back to back
 * (layer 0 first; the final layer is unsalted).  The transcript is continued in place.
 * layers_out: all layers back to back (n + n/2 + ... + final_size values).
 * roots_out: 32 bytes per layer.  betas_out: one per fold.  Returns the number of folds. */
size_t to_fri_commit(const uint64_t* layer0, size_t n, uint64_t shift, size_t final_size, const uint8_t* salts,
                     to_transcript* t, uint64_t* layers_out, uint8_t* roots_out, uint64_t* betas_out);
/* Same loop over the extension field (fri_fold_ext + squeeze_ext_challenge); Ext leaves are
 * salt || 32-byte Ext::to_bytes.  No reference caller exists: shape mirrors the base loop. */
size_t to_fri_commit_ext(const uint64_t* layer0, size_t n, uint64_t shift, size_t final_size, const uint8_t* salts,
                         to_transcript* t, uint64_t* layers_out, uint8_t* roots_out, uint64_t* betas_out);

/* ---- synthetic data: SplitMix64 stream, value = next_u64 % p ---- */
void to_fill_random(uint64_t* out, size_t n, uint64_t seed);
void to_fill_random_bytes(uint8_t* out, size_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
