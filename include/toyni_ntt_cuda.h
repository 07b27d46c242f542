/*
 * toyni_ntt_cuda.h — C ABI of the B200-native (sm_100a) BabyBear prover hot path.
 *
 * The library is a drop-in for the static library `ntt_cuda` that the reference links
 * (`#[link(name = "ntt_cuda", kind = "static")]`, src/ntt.rs:95; built by build.rs:75-116 from
 * cuda/ntt_kernel.cu).  Section 1 keeps, symbol for symbol, what src/ntt.rs:96-110 imports.
 * Section 2 adds device-resident, stream-ordered entry points for the rest of the path
 * (coset LDE, FRI fold, Merkle commit, FRI commit loop); section 3 the host-pointer forms
 * that the reference call sites in src/math/domain.rs, src/math/fri.rs and src/fibonacci.rs
 * would bind (see INTEGRATION.md for the Rust side).
 *
 * Conventions
 *   - Host field arrays use the reference's storage: one canonical value in [0,p) per uint64_t
 *     (`#[repr(C)] struct BabyBear { value: u64 }`, src/babybear.rs:10-14); an Ext is four of
 *     them, limb order [a0,a1,a2,a3] (src/ext.rs:22-26).  p = 2013265921.
 *   - Device field arrays are uint32_t (4 bytes per base element, 16 per Ext, same limb order).
 *   - All `bb_*` functions return 0 on success or a cudaError_t value (also kept as a sticky
 *     last error, bb_last_error()).  The void functions of section 1 record errors the same way.
 *   - There is no CPU fallback: without a usable sm_100 device every compute entry point fails.
 */
#ifndef TOYNI_NTT_CUDA_H
#define TOYNI_NTT_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * 1. Symbols the reference's Rust side already imports (src/ntt.rs:96-110).
 *    `count` is in uint64_t elements (cuda/ntt_kernel.cu:298-312).
 * ---------------------------------------------------------------------------------------- */
int cuda_copy_to_device(uint64_t* d_dest, const uint64_t* h_src, size_t count);   /* src/ntt.rs:97  */
int cuda_copy_from_device(uint64_t* h_dest, const uint64_t* d_src, size_t count); /* src/ntt.rs:98  */
int cuda_malloc(uint64_t** d_ptr, size_t count);                                  /* src/ntt.rs:99  */
int cuda_free(uint64_t* d_ptr);                                                   /* src/ntt.rs:100 */
const char* cuda_get_error_string(int error);                                     /* src/ntt.rs:101 */
/* cudaGetDeviceCount (src/ntt.rs:102) is resolved from libcudart, as in the reference. */

/* Per-size context (src/ntt.rs:107; cuda/ntt_kernel.cu:213-234).  n must be a power of two with
 * log2(n) <= 27, otherwise NULL (cuda/ntt_kernel.cu:220).  Contexts share one per-device twiddle
 * cache generated on the device; a context owns its staging buffers and a stream and serialises
 * concurrent callers itself. */
void* ntt_ctx_create(uint32_t n);
void ntt_ctx_destroy(void* ctx);                                                  /* cuda/ntt_kernel.cu:236 */
/* Forward / inverse NTT, in place on HOST data, natural order in and out, canonical outputs
 * (src/ntt.rs:108-109; cuda/ntt_kernel.cu:249-292).  Bit-exact with ntt()/intt() of
 * src/ntt.rs:24-66 for the canonical root BabyBear::get_root_of_unity(log2 n). */
void ntt_run_inplace(void* ctx, uint64_t* h_data);
void intt_run_inplace(void* ctx, uint64_t* h_data);
/* The same transforms with the outcome as the return value (0 or a cudaError_t): what a binding should call — the
 * reference's void signatures cannot report a failed copy or launch (cuda/ntt_kernel.cu:249-292 ignores them). */
int ntt_run_inplace_rc(void* ctx, uint64_t* h_data);
int intt_run_inplace_rc(void* ctx, uint64_t* h_data);
/* Not part of the drop-in surface (src/ntt.rs keeps a BabyBear in a u64): the same in-place host transform on canonical
 * u32 values, i.e. half the PCIe bytes.  dir: 0 forward, 1 inverse. */
int bb_ntt_host_u32(void* ctx, uint32_t* h_data, int dir);

/* ------------------------------------------------------------------------------------------
 * 2. Device-resident, stream-ordered API.
 * ---------------------------------------------------------------------------------------- */
int bb_last_error(void);                 /* sticky, per host thread: first error since that thread's last bb_clear_error() */
const char* bb_last_error_string(void);
void bb_clear_error(void);
int bb_device_ok(void);                  /* 1 iff the current device is compute capability 10.0 and the sm_100a image loads */
/* Stream used by every later call on this host thread's library state (NULL = legacy default stream). */
void bb_set_stream(void* cuda_stream);
int bb_sync(void);                       /* synchronise that stream */

int bb_dev_alloc(void** d_ptr, size_t bytes);
int bb_dev_free(void* d_ptr);
/* stream-ordered allocation on the bound stream from a pool that keeps freed memory (cudaMallocAsync): repeated
 * proofs allocate their LDE-sized arrays without driver calls; bb_pool_trim synchronises and returns the memory */
int bb_pool_alloc(void** d_ptr, size_t bytes);
int bb_pool_free(void* d_ptr);
int bb_pool_trim(void);
int bb_h2d(void* d_dst, const void* h_src, size_t bytes);            /* async on the stream */
int bb_d2h(void* h_dst, const void* d_src, size_t bytes);            /* async on the stream */
int bb_d2d(void* d_dst, const void* d_src, size_t bytes);            /* async on the stream */
/* width conversion between the reference's u64 storage and device u32 (values are reduced mod p) */
int bb_narrow_u64_to_u32(const uint64_t* d_src, uint32_t* d_dst, size_t count);
int bb_widen_u32_to_u64(const uint32_t* d_src, uint64_t* d_dst, size_t count);

/* dir: 0 forward (src/ntt.rs:24-53), 1 inverse incl. the n^-1 scale (src/ntt.rs:56-66). In place. */
int bb_ntt_device(uint32_t* d_data, uint32_t log_n, int dir);
/* `batch` independent vectors of 2^log_n elements, vector b at d_data + b * 2^log_n. */
int bb_ntt_batch_device(uint32_t* d_data, uint32_t log_n, size_t batch, int dir);
/* AoS extension-field array (4 limbs per element): the four coordinate transforms of
 * transform_ext (src/math/domain.rs:140-151) in one go. */
int bb_ntt_ext_device(uint32_t* d_data, uint32_t log_n, int dir);

/* Building blocks of the sharded four-step NTT (n = n1*n2, input index j1*n2 + j2, output k1 + n1*k2; one rank
 * owns a block of `cols` columns j2 of the n1 x n2 matrix):
 *   bb_ntt_columns_device  - n1-point transforms down the columns of a row-major [2^log_n1][cols] block
 *   bb_fourstep_twiddle_device - multiply element (k1, c) by w_n^((col_offset + c) * k1)
 * The transpose between the two halves is an all-to-all owned by the caller (toyni_b200/multigpu.py, NCCL). */
int bb_ntt_columns_device(uint32_t* d_block, uint32_t log_n1, size_t cols, int dir);
int bb_fourstep_twiddle_device(uint32_t* d_block, uint32_t log_n, uint32_t log_n1, size_t cols, size_t col_offset, int dir);
/* Fused form: the column transforms of this rank's block, with the twiddle AND the transpose folded into the last
 * NTT pass.  Row k1 is stored, already multiplied by w_n^(j2*k1), straight into peer_bufs[k1 / (n1/nranks)] at
 * [k1 mod (n1/nranks)][rank*cols + c] (row stride n2): peer_bufs[r] is rank r's receive buffer of (n1/nranks) x n2
 * words — this rank's own allocation for r == rank, a CUDA-IPC mapping (stores over NVLink) otherwise.  The caller
 * synchronises the ranks before (buffers free) and after (stores landed); no NCCL call, no re-layout pass. */
int bb_ntt_columns_scatter_device(uint32_t* d_block, uint32_t log_n, uint32_t log_n1, size_t cols, int dir,
                                  void* const* peer_bufs, uint32_t nranks, uint32_t rank);
/* Device-side rendezvous for the fused path (no host synchronisation, no NCCL): every rank owns 8 flag words
 * (one 128-byte line per writer).  signal: after this rank's scatter kernels (stream order), write `epoch` into
 * every rank's flag line for this rank (d_peer_flags: DEVICE array of nranks pointers to the ranks' flag blocks).
 * wait: spin until all nranks lines of this rank's own flag block reach `epoch`; a lost peer sets *d_err (device
 * word) after ~2 s instead of hanging. */
int bb_peer_signal_device(void* const* d_peer_flags, uint32_t nranks, uint32_t rank, uint32_t epoch);
int bb_peer_wait_device(void* d_flags, uint32_t nranks, uint32_t epoch, void* d_err);
/* CUDA IPC plumbing for the peer buffers (one process per GPU): 64-byte handles, exchanged by the caller. */
int bb_ipc_get_handle(const void* d_ptr, uint8_t handle_out[64]);
int bb_ipc_open_handle(const uint8_t handle[64], void** d_ptr_out);
int bb_ipc_close_handle(void* d_ptr);

/* BabyBearDomain::elements (src/math/domain.rs:61-69): d_out[i] = shift * w^i, i < 2^log_n; shift = 1 is
 * roots_of_unity_domain (src/ntt.rs:69-81).  From the twiddle cache, no sequential running product. */
int bb_domain_elements_device(uint32_t log_n, uint32_t shift, uint32_t* d_out);

/* BabyBearDomain::fft on a coset (src/math/domain.rs:107-123,154-162): zero-pad/truncate the
 * n_coeffs coefficients to 2^log_size, multiply by shift^i, forward NTT.  Fused into the first
 * NTT pass: only the n_coeffs inputs are read.  shift == 1 is the plain domain.  d_out must not
 * alias d_coeffs unless n_coeffs == 2^log_size.  limbs: 1 or 4 (fft_ext, :135-137). */
int bb_coset_fft_device(const uint32_t* d_coeffs, size_t n_coeffs, uint32_t log_size, uint32_t shift, int limbs,
                        uint32_t* d_out);
/* BabyBearDomain::ifft (src/math/domain.rs:85-102,165-174): INTT then multiply by shift^-i
 * (fused into the last pass together with n^-1).  In place.  limbs: 1 or 4 (ifft_ext, :130-132). */
int bb_coset_ifft_device(uint32_t* d_evals, uint32_t log_size, uint32_t shift, int limbs);

/* fri_fold / fri_fold_ext (src/math/fri.rs:27-48, :7-25) for evaluation points x_i = x0 * w_m^i,
 * w_m = get_root_of_unity(log2 m): the layout the prover uses (src/fibonacci.rs:214,228-231 gives
 * x0 = shift^(2^k) for layer k).  m values in, m/2 out; x^-1 is generated on chip.
 * limbs: 1 (beta[0] used) or 4. */
int bb_fri_fold_device(const uint32_t* d_evals, size_t m, uint32_t x0, const uint32_t beta[4], int limbs,
                       uint32_t* d_out);
/* Same for one shard of a cyclic multi-GPU layout: the shard holds global indices
 * rank, rank+nranks, ...; m_local = m/nranks values in, m_local/2 out, no communication. */
int bb_fri_fold_shard_device(const uint32_t* d_evals, size_t m_local, uint32_t log_m, uint32_t x0,
                             const uint32_t beta[4], int limbs, uint32_t nranks, uint32_t rank, uint32_t* d_out);
/* A whole fold chain on one cyclic shard in ONE call (the prover's loop, src/fibonacci.rs:213-231, without the commits):
 * layer k+1 = fold(layer k) with beta = betas[k * limbs ..], x0 squared from layer to layer, while the layer has more
 * than `until` values and at least 2 * nranks.  The folded layers land back to back in d_layers
 * (m_local/2 + m_local/4 + ... values); *folds_out = number of folds.  No host work between the launches. */
int bb_fri_fold_chain_shard_device(const uint32_t* d_layer0, size_t m_local, uint32_t log_m, uint32_t shift, const uint32_t* betas,
                                   size_t nbetas, int limbs, uint32_t nranks, uint32_t rank, size_t until, uint32_t* d_layers,
                                   size_t* folds_out);
/* Reference signature with an explicit xs array (only xs[0..m/2) is read, src/math/fri.rs:13-16). */
int bb_fri_fold_xs_device(const uint32_t* d_evals, size_t m, const uint32_t* d_xs, const uint32_t beta[4], int limbs,
                          uint32_t* d_out);

/* Merkle commitment (src/merkle.rs:16-48 with the prover's leaf encoding, src/fibonacci.rs:340-363):
 * leaf i = SHA256(0x00 || salt_i[16] || LE-u64 of each limb), or without the salt when d_salts is
 * NULL (build_unsalted_tree).  d_nodes receives every level, leaf level first, 32 bytes per node:
 * bb_merkle_node_count(n) * 32 bytes.  root_out (host, 32 bytes) may be NULL; if given the call
 * synchronises the stream. */
size_t bb_merkle_node_count(size_t nleaves);
int bb_merkle_commit_device(const uint32_t* d_vals, int limbs, size_t n, const uint8_t* d_salts, uint8_t* d_nodes,
                            uint8_t* root_out);
/* MerkleTree::new over arbitrary equal-length byte-string leaves already on the device. */
int bb_merkle_build_bytes_device(const uint8_t* d_leaves, size_t n, size_t leaf_len, uint8_t* d_nodes, uint8_t* root_out);
/* MerkleTree::get_proof (src/merkle.rs:50-80): path_out = depth*32 bytes (host), pos_out = depth flags. */
int bb_merkle_open_device(const uint8_t* d_nodes, size_t nleaves, size_t index, uint8_t* path_out, uint8_t* pos_out,
                          size_t* depth_out);

/* The prover's FRI commit loop (src/fibonacci.rs:200-247) on device-resident data.
 *   d_layer0   : n values (limbs u32 each) on the coset {shift * w_n^i}
 *   final_size : stop when the layer has this many values (src/fibonacci.rs:220-222)
 *   d_salts    : 16 bytes per leaf for every salted layer, layers back to back, layer 0 first
 *                (the final layer is committed unsalted, :234-238)
 *   challenge  : called once per fold with the root just committed (absorb) and must return
 *                beta (squeeze) — limbs values; it is the host transcript (src/transcript.rs).
 *                When NULL, betas_in supplies limbs values per fold (fold-only benchmarking).
 *   d_layers   : receives the folded layers 1, 2, ... back to back (n/2 + n/4 + ... + final_size values);
 *                layer 0 stays in d_layer0
 *   d_nodes    : receives the trees back to back (bb_merkle_node_count per layer); NULL skips hashing
 *   roots_out  : 32 bytes per layer (host); NULL if d_nodes is NULL
 * Returns the number of folds through *folds_out. */
typedef void (*bb_challenge_fn)(void* user, const uint8_t root[32], uint32_t layer, uint32_t* beta_out);
int bb_fri_commit_device(const uint32_t* d_layer0, size_t n, uint32_t shift, size_t final_size, int limbs,
                         const uint8_t* d_salts, bb_challenge_fn challenge, void* user, const uint32_t* betas_in,
                         uint32_t* d_layers, uint8_t* d_nodes, uint8_t* roots_out, size_t* folds_out);

/* Openings of a whole query set in one launch (src/fibonacci.rs:250-295 over src/merkle.rs:50-80): `indices` (host)
 * names nq leaves; paths_out (host) receives nq * depth * 32 bytes, pos_out nq * depth flags, *depth_out the depth.
 * bb_gather_device fetches the opened values / salts: out[q] = src[indices[q]] for elements of elem_bytes bytes. */
int bb_merkle_open_batch_device(const uint8_t* d_nodes, size_t nleaves, const uint64_t* indices, size_t nq, uint8_t* paths_out,
                                uint8_t* pos_out, size_t* depth_out);
int bb_gather_device(const void* d_src, size_t elem_bytes, const uint64_t* indices, size_t nq, void* out);
/* The openings of a whole proof in ONE launch, one copy and one synchronisation: request r opens leaves
 * indices[first .. first + count) of its own tree (requests cover indices[0 .. nq) in order).  paths_out receives, query
 * after query, depth(tree) * 32 bytes (paths_bytes = their total, checked), pos_out the matching position flags,
 * vals_out nq * elem_bytes bytes (elem_bytes = 4 for base-field leaves, 16 for Ext), salts_out nq * 16 bytes (zeros for
 * unsalted trees).  A 2^20-row proof has 24 trees and ~2100 openings: 72 round trips through the per-tree calls. */
typedef struct {
    const uint8_t* d_nodes; /* every level of the tree, leaf level first (bb_merkle_commit_device) */
    size_t nleaves;
    const void* d_vals;     /* the committed values */
    const uint8_t* d_salts; /* 16 bytes per leaf, NULL for an unsalted tree */
    size_t first, count;
} bb_open_request;
int bb_merkle_open_multi_device(const bb_open_request* reqs, size_t nreq, const uint64_t* indices, size_t nq, size_t elem_bytes,
                                uint8_t* paths_out, size_t paths_bytes, uint8_t* pos_out, uint8_t* vals_out, uint8_t* salts_out);

/* Re-layout after the cyclic -> block exchange of the sharded FRI commit (one process per GPU): d_src holds `groups`
 * runs of `chunk` elements (limbs = 1 or 4 words each), run r supplying destination positions r, r + groups, ... */
int bb_interleave_device(const uint32_t* d_src, uint32_t groups, size_t chunk, int limbs, uint32_t* d_dst);

/* Element-wise stages of StarkProver::generate_proof between the LDE and the FRI commit loop, device-resident
 * (SURVEY 8f rank 1).  x_i = shift * w_N^i (N = 2^log_n) comes from the twiddle cache; `step` is the blowup, so that
 * T(g x_i) = d_trace_lde[(i + step) mod N] (src/verifier.rs:128-129).
 *   constraint: c[i] = (T(g^2 x) - T(g x) - T(x)) (x - b1) (x - b2)                     src/fibonacci.rs:133-143
 *   scale_periodic: v[i] *= table[i mod period] (1 / Z_H(x_i) takes `blowup` values)      src/fibonacci.rs:147-150
 *   deep: d[i] = ((Q - q_z) + (T(g^2 x) - t_ggz) + (T(g x) - t_gz) + (T(x) - t_z)) / (x_i - z)   src/fibonacci.rs:186-198
 *   poly_eval: sum_k coeffs[k] z^k                                                       src/math/polynomial.rs:134-144 */
int bb_fib_constraint_device(const uint32_t* d_trace_lde, uint32_t log_n, uint32_t step, uint32_t shift, uint32_t b1, uint32_t b2,
                             uint32_t* d_out);
int bb_scale_periodic_device(uint32_t* d_vals, size_t n, const uint32_t* table, uint32_t period);
int bb_fib_deep_device(const uint32_t* d_quotient, const uint32_t* d_trace_lde, uint32_t log_n, uint32_t step, uint32_t shift, uint32_t z,
                       uint32_t q_z, uint32_t t_z, uint32_t t_gz, uint32_t t_ggz, uint32_t* d_out);
int bb_poly_eval_device(const uint32_t* d_coeffs, size_t n, uint32_t z, uint32_t* value_out);

/* Tuning / introspection */
int bb_ntt_set_plan(uint32_t log_n, int npass, const int* log_rows, const int* log_cols); /* npass 0 = default */
int bb_ntt_get_plan(uint32_t log_n, int* log_rows, int* log_cols);                        /* returns npass */
void bb_ntt_set_kernel(int kernel);                     /* 1 (default): TMA-staged two-pass kernel for plain 2^24-point vectors; 0: tile kernel everywhere */
int bb_ntt_launches(uint32_t log_n);                    /* kernels launched per device-resident transform */
unsigned long long bb_kernel_launch_count(void);        /* total launches of this library's kernels so far */
int bb_warmup(uint32_t log_n);                          /* build tables and scratch for this size */
int bb_ntt_diag(uint32_t words_out[16]);                /* diagnostic words of the TMA-staged kernel: [0] time-out code (0 = none) */
void bb_release(void);                                  /* free all cached device memory on this device */

/* ------------------------------------------------------------------------------------------
 * 3. Host-pointer forms of the reference call sites (u64 storage in and out).
 *    Each does H2D + kernels + D2H on the library stream and returns after synchronising.
 * ---------------------------------------------------------------------------------------- */
/* BabyBearDomain::fft / ifft (src/math/domain.rs:107-123, :85-102) */
int toyni_domain_fft(const uint64_t* coeffs, size_t n_coeffs, size_t size, uint64_t shift, uint64_t* evals_out);
int toyni_domain_ifft(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* coeffs_out);
/* roots_of_unity_domain(n) (src/ntt.rs:69-81) and BabyBearDomain::elements() (src/math/domain.rs:61-69) */
int toyni_roots_of_unity_domain(size_t n, uint64_t* out);
int toyni_domain_elements(size_t size, uint64_t shift, uint64_t* out);
/* fft_ext / ifft_ext (src/math/domain.rs:129-137): AoS Ext arrays, 4 x u64 per element */
int toyni_domain_fft_ext(const uint64_t* coeffs, size_t n_coeffs, size_t size, uint64_t shift, uint64_t* evals_out);
int toyni_domain_ifft_ext(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* coeffs_out);
/* fri_fold / fri_fold_ext (src/math/fri.rs:27, :7): xs has at least m/2 entries */
int toyni_fri_fold(const uint64_t* evals, size_t m, const uint64_t* xs, uint64_t beta, uint64_t* out);
int toyni_fri_fold_ext(const uint64_t* evals, size_t m, const uint64_t* xs, const uint64_t beta[4], uint64_t* out);
/* build_merkle_tree / build_unsalted_tree (src/fibonacci.rs:340-363): salts NULL = unsalted.
 * nodes_out (host, bb_merkle_node_count(n)*32 bytes) may be NULL when only the root is wanted. */
int toyni_merkle_commit(const uint64_t* values, size_t n, int limbs, const uint8_t* salts, uint8_t* nodes_out,
                        uint8_t root_out[32]);

/* ------------------------------------------------------------------------------------------
 * 4. The sharded paths from ONE host process over the G GPUs of a box (devices 0 .. G-1, G a power of two <= 8).
 *    The reference has no multi-GPU code; these are the entry points a Rust host binds for SURVEY 8e.  bb_mg_init
 *    enables peer access between all pairs: the kernels of one GPU store straight into buffers of another
 *    (NVLink / NVSwitch), devices are ordered by CUDA events on one stream per device — no NCCL, no host
 *    synchronisation inside a transform.  Calls return once the work is queued; bb_mg_sync waits for it.
 * ---------------------------------------------------------------------------------------- */
int bb_mg_init(int ngpus, void** mg_out);
void bb_mg_destroy(void* mg);
int bb_mg_ngpus(void* mg);
int bb_mg_sync(void* mg);
void* bb_mg_stream(void* mg, int device);   /* the cudaStream_t the layer uses on that device */
/* ONE 2^log_n transform over the G devices (four-step, n = n1 n2 with log2 n1 = ceil(log_n / 2)).  d_blocks[r] (on
 * device r) holds column block r of the row-major n1 x n2 input matrix, i.e. (n1, n2/G) words, and is destroyed;
 * d_outs[r] (on device r, (n1/G) x n2 words) receives out[k1_local][k2] = X[k1 + n1 k2], k1 = r n1/G + k1_local.  The
 * inter-half twiddle and the transpose run inside the last pass of the column transforms, which stores each row into
 * d_outs[owner] directly. */
int bb_mg_ntt_fourstep(void* mg, uint32_t log_n, int dir, uint32_t* const* d_blocks, uint32_t* const* d_outs);
/* Independent 2^log_n-point columns, ncols[r] of them back to back at d_cols[r] on device r, in place; no exchange. */
int bb_mg_ntt_batch(void* mg, uint32_t log_n, int dir, uint32_t* const* d_cols, const size_t* ncols);
/* FRI fold chain (src/math/fri.rs:7-48 layer after layer, x squared per layer as src/fibonacci.rs:228-231) on cyclic
 * shards: d_shards[r] holds the values i = r (mod G) of the 2^log_m codeword on shift * <w>; folds while the layer has
 * more than final_size values and its fold partner is still on the same device (m/2 >= G).  betas: limbs words per
 * fold.  d_layers_out[r] receives the local parts of layers 1, 2, ... back to back; *folds_out the number of folds. */
int bb_mg_fri_chain(void* mg, uint32_t log_m, uint32_t shift, int limbs, size_t final_size, const uint32_t* betas,
                    const uint32_t* const* d_shards, uint32_t* const* d_layers_out, size_t* folds_out);
/* Host-pointer form of the sharded transform: 2^log_n canonical u64 values in natural order, in place (what
 * ntt_run_inplace does on one device, src/ntt.rs:108): scatter, four-step, gather.  Synchronous. */
int bb_mg_ntt_host(void* mg, uint64_t* h_data, uint32_t log_n, int dir);

/* ------------------------------------------------------------------------------------------
 * 5. The prover loop behind one call: StarkProver::generate_proof (src/fibonacci.rs:99-310) for the Fibonacci AIR, every
 *    LDE-sized array on the device, the Fiat-Shamir transcript (src/transcript.rs) on the host inside the library
 *    (toyni_b200/host/toyni_prover.hpp is the loop).  The reference draws its randomness from thread_rng(); here it is an
 *    argument, so that a proof is reproducible:
 *      trace        trace_len canonical values (a power of two, 32 * trace_len <= 2^27)
 *      mask         140 blinding coefficients (MASK_DEGREE, src/fibonacci.rs:19,117-120)
 *      salts_trace, salts_quot   16 bytes per LDE point (32 * trace_len of them each)
 *      salts_fri    toyni_fri_salt_bytes(trace_len) bytes: 16 per leaf of every salted FRI layer, layer 0 first
 *      salts_on_device  non-zero: the three salt pointers are device pointers (nothing LDE-sized crosses PCIe)
 *    The canonical proof bytes (the format of toyni_b200/proof.py / toyni::serialize_proof; the reference has no
 *    serialization) are written to proof_out; *proof_len receives their length also when proof_cap is too small
 *    (cudaErrorInvalidValue then).  toyni_prover_error() describes the last failure of this host thread.
 * ---------------------------------------------------------------------------------------- */
size_t toyni_fri_salt_bytes(size_t trace_len);
int toyni_prove_fibonacci(const uint64_t* trace, size_t trace_len, const uint64_t* mask, const uint8_t* salts_trace,
                          const uint8_t* salts_quot, const uint8_t* salts_fri, size_t salts_fri_bytes, int salts_on_device,
                          uint8_t* proof_out, size_t proof_cap, size_t* proof_len);
const char* toyni_prover_error(void);

#ifdef __cplusplus
}
#endif
#endif
