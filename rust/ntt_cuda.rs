// rust/ntt_cuda.rs — the `mod cuda` block of src/ntt.rs (jonas089/toyni), re-written against the B200-native library of
// this repository (include/toyni_ntt_cuda.h).  It REPLACES lines 83-315 of the reference's src/ntt.rs (everything from
// `#[cfg(feature = "cuda")] mod cuda {` to the end of the file); lines 1-81 of that file (the CPU `ntt`, `intt`,
// `roots_of_unity_domain`) stay exactly as they are and are not carried here.  NOT COMPILED HERE: this image has no Rust
// toolchain; the same C ABI is exercised from Python (toyni_b200/ntt.py), C++ (toyni_b200/host/toyni.hpp) and by tests/.
//
// Public surface kept exactly (`--features cuda`): `cuda_available`, `ntt_cuda`, `intt_cuda`, `CudaBuffer`.  Added (same
// feature): `coset_fft_cuda`, `coset_ifft_cuda`, `fri_fold_cuda`, `fri_fold_ext_cuda`, `merkle_commit_cuda`, the
// device-resident prover stages, and the multi-GPU layer `MultiGpu` (header section 4), which the call sites in
// src/math/domain.rs, src/math/fri.rs and src/fibonacci.rs use behind their existing `use_gpu` flag.

// ── CUDA path (feature = "cuda") ─────────────────────────────────────────────────────────────
#[cfg(feature = "cuda")]
mod cuda {
    use super::BabyBear;
    use crate::ext::Ext;
    use std::ffi::{c_void, CStr};
    use std::os::raw::c_char;

    type CudaError = i32;
    const CUDA_SUCCESS: CudaError = 0;

    /// bb_open_request of the header: one tree and the slice of the index list that opens it
    #[repr(C)]
    pub struct OpenRequest {
        pub d_nodes: *const u8,
        pub nleaves: usize,
        pub d_vals: *const c_void,
        pub d_salts: *const u8, // null: unsalted tree
        pub first: usize,
        pub count: usize,
    }

    #[link(name = "ntt_cuda", kind = "static")]
    unsafe extern "C" {
        // unchanged from the reference (src/ntt.rs:96-110)
        fn cuda_copy_to_device(d_dest: *mut u64, h_src: *const u64, count: usize) -> CudaError;
        fn cuda_copy_from_device(h_dest: *mut u64, d_src: *const u64, count: usize) -> CudaError;
        fn cuda_malloc(d_ptr: *mut *mut u64, count: usize) -> CudaError;
        fn cuda_free(d_ptr: *mut u64) -> CudaError;
        fn cuda_get_error_string(error: CudaError) -> *const c_char;
        fn cudaGetDeviceCount(count: *mut i32) -> CudaError;
        fn ntt_ctx_create(n: u32) -> *mut c_void;
        fn ntt_run_inplace(ctx: *mut c_void, h_data: *mut u64);
        fn intt_run_inplace(ctx: *mut c_void, h_data: *mut u64);
        // the same transforms with the outcome as a return value (the void forms above cannot report a failed launch)
        fn ntt_run_inplace_rc(ctx: *mut c_void, h_data: *mut u64) -> CudaError;
        fn intt_run_inplace_rc(ctx: *mut c_void, h_data: *mut u64) -> CudaError;
        // new in the B200 library (include/toyni_ntt_cuda.h, sections 2-3)
        fn bb_last_error() -> CudaError;
        fn bb_clear_error();
        fn bb_device_ok() -> i32;
        fn toyni_domain_fft(coeffs: *const u64, n_coeffs: usize, size: usize, shift: u64, out: *mut u64) -> CudaError;
        fn toyni_domain_ifft(evals: *const u64, size: usize, shift: u64, out: *mut u64) -> CudaError;
        fn toyni_domain_fft_ext(coeffs: *const u64, n_coeffs: usize, size: usize, shift: u64, out: *mut u64) -> CudaError;
        fn toyni_domain_ifft_ext(evals: *const u64, size: usize, shift: u64, out: *mut u64) -> CudaError;
        fn toyni_fri_fold(evals: *const u64, m: usize, xs: *const u64, beta: u64, out: *mut u64) -> CudaError;
        fn toyni_fri_fold_ext(evals: *const u64, m: usize, xs: *const u64, beta: *const u64, out: *mut u64) -> CudaError;
        fn toyni_merkle_commit(values: *const u64, n: usize, limbs: i32, salts: *const u8, nodes_out: *mut u8,
                               root_out: *mut u8) -> CudaError;
        fn bb_merkle_node_count(nleaves: usize) -> usize;
        // device-resident prover stages and query-set openings (header section 2; device pointers are *u32 / *u8)
        fn bb_fib_constraint_device(d_trace_lde: *const u32, log_n: u32, step: u32, shift: u32, b1: u32, b2: u32, d_out: *mut u32) -> CudaError;
        fn bb_scale_periodic_device(d_vals: *mut u32, n: usize, table: *const u32, period: u32) -> CudaError;
        fn bb_fib_deep_device(d_quotient: *const u32, d_trace_lde: *const u32, log_n: u32, step: u32, shift: u32, z: u32,
                              q_z: u32, t_z: u32, t_gz: u32, t_ggz: u32, d_out: *mut u32) -> CudaError;
        fn bb_poly_eval_device(d_coeffs: *const u32, n: usize, z: u32, value_out: *mut u32) -> CudaError;
        fn bb_merkle_open_batch_device(d_nodes: *const u8, nleaves: usize, indices: *const u64, nq: usize, paths_out: *mut u8,
                                       pos_out: *mut u8, depth_out: *mut usize) -> CudaError;
        fn bb_gather_device(d_src: *const c_void, elem_bytes: usize, indices: *const u64, nq: usize, out: *mut c_void) -> CudaError;
        // every opening of a proof in one launch (24 trees, ~2100 leaves at 2^20 rows): requests cover indices[0..nq) in order
        fn bb_merkle_open_multi_device(reqs: *const OpenRequest, nreq: usize, indices: *const u64, nq: usize, elem_bytes: usize,
                                       paths_out: *mut u8, paths_bytes: usize, pos_out: *mut u8, vals_out: *mut u8, salts_out: *mut u8) -> CudaError;
        // the rest of what a device-resident StarkProver::generate_proof calls (toyni_b200/host/toyni_prover.hpp is that loop
        // in C++)
        fn bb_pool_alloc(d_ptr: *mut *mut c_void, bytes: usize) -> CudaError;
        fn bb_pool_free(d_ptr: *mut c_void) -> CudaError;
        fn bb_h2d(d_dst: *mut c_void, h_src: *const c_void, bytes: usize) -> CudaError;
        fn bb_d2h(h_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> CudaError;
        fn bb_d2d(d_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> CudaError;
        fn bb_sync() -> CudaError;
        fn bb_coset_fft_device(d_coeffs: *const u32, n_coeffs: usize, log_size: u32, shift: u32, limbs: i32, d_out: *mut u32) -> CudaError;
        fn bb_coset_ifft_device(d_evals: *mut u32, log_size: u32, shift: u32, limbs: i32) -> CudaError;
        fn bb_merkle_commit_device(d_vals: *const u32, limbs: i32, n: usize, d_salts: *const u8, d_nodes: *mut u8, root_out: *mut u8) -> CudaError;
        fn bb_fri_commit_device(d_layer0: *const u32, n: usize, shift: u32, final_size: usize, limbs: i32, d_salts: *const u8,
                                challenge: Option<extern "C" fn(user: *mut c_void, root: *const u8, layer: u32, beta_out: *mut u32)>,
                                user: *mut c_void, betas_in: *const u32, d_layers: *mut u32, d_nodes: *mut u8, roots_out: *mut u8,
                                folds_out: *mut usize) -> CudaError;
        // the whole prover loop behind one call (header section 5): canonical proof bytes out
        fn toyni_fri_salt_bytes(trace_len: usize) -> usize;
        fn toyni_prove_fibonacci(trace: *const u64, trace_len: usize, mask: *const u64, salts_trace: *const u8, salts_quot: *const u8,
                                 salts_fri: *const u8, salts_fri_bytes: usize, salts_on_device: i32, proof_out: *mut u8, proof_cap: usize,
                                 proof_len: *mut usize) -> CudaError;
        fn toyni_prover_error() -> *const c_char;
        // one process, G devices (header section 4)
        fn bb_mg_init(ngpus: i32, mg_out: *mut *mut c_void) -> CudaError;
        fn bb_mg_destroy(mg: *mut c_void);
        fn bb_mg_sync(mg: *mut c_void) -> CudaError;
        fn bb_mg_ntt_fourstep(mg: *mut c_void, log_n: u32, dir: i32, d_blocks: *const *mut u32, d_outs: *const *mut u32) -> CudaError;
        fn bb_mg_ntt_batch(mg: *mut c_void, log_n: u32, dir: i32, d_cols: *const *mut u32, ncols: *const usize) -> CudaError;
        fn bb_mg_fri_chain(mg: *mut c_void, log_m: u32, shift: u32, limbs: i32, final_size: usize, betas: *const u32,
                           d_shards: *const *const u32, d_layers_out: *const *mut u32, folds_out: *mut usize) -> CudaError;
        fn bb_mg_ntt_host(mg: *mut c_void, h_data: *mut u64, log_n: u32, dir: i32) -> CudaError;
    }

    // BabyBear is `#[repr(C)] { value: u64 }` and Ext is `#[repr(C)] { c: [BabyBear; 4] }`
    const _: () = assert!(std::mem::size_of::<BabyBear>() == 8 && std::mem::align_of::<BabyBear>() == 8);
    const _: () = assert!(std::mem::size_of::<Ext>() == 32);

    struct CtxPtr(*mut c_void);
    unsafe impl Send for CtxPtr {}
    unsafe impl Sync for CtxPtr {}

    fn err_string(e: CudaError) -> String {
        unsafe { CStr::from_ptr(cuda_get_error_string(e)).to_string_lossy().into_owned() }
    }

    fn get_or_create_ctx(n: usize) -> Result<*mut c_void, String> {
        use std::collections::HashMap;
        use std::sync::{Mutex, OnceLock};
        static CACHE: OnceLock<Mutex<HashMap<usize, CtxPtr>>> = OnceLock::new();
        let mut guard = CACHE.get_or_init(|| Mutex::new(HashMap::new())).lock().unwrap();
        if let Some(p) = guard.get(&n) {
            return Ok(p.0);
        }
        let ptr = unsafe { ntt_ctx_create(n as u32) };
        if ptr.is_null() {
            // the reference caches a null context unchecked (src/ntt.rs:138-139); this one reports it
            return Err(format!("ntt_ctx_create({n}) failed: {}", err_string(unsafe { bb_last_error() })));
        }
        guard.insert(n, CtxPtr(ptr));
        Ok(ptr)
    }

    /// True iff a CUDA device is present AND it can run this library (sm_100 only, no PTX fallback).
    pub fn cuda_available() -> bool {
        unsafe {
            let mut count = 0;
            cudaGetDeviceCount(&mut count) == CUDA_SUCCESS && count > 0 && bb_device_ok() == 1
        }
    }

    pub struct CudaBuffer {
        ptr: *mut u64,
        size: usize,
    }

    impl CudaBuffer {
        pub fn new(size: usize) -> Result<Self, String> {
            let mut ptr: *mut u64 = std::ptr::null_mut();
            let err = unsafe { cuda_malloc(&mut ptr, size) };
            if err != CUDA_SUCCESS {
                return Err(format!("CUDA malloc failed: {}", err_string(err)));
            }
            Ok(Self { ptr, size })
        }
        pub fn copy_from_host(&mut self, data: &[u64]) -> Result<(), String> {
            assert_eq!(data.len(), self.size, "Size mismatch");
            let err = unsafe { cuda_copy_to_device(self.ptr, data.as_ptr(), self.size) };
            if err != CUDA_SUCCESS {
                return Err(format!("CUDA copy to device failed: {}", err_string(err)));
            }
            Ok(())
        }
        pub fn copy_to_host(&self, data: &mut [u64]) -> Result<(), String> {
            assert_eq!(data.len(), self.size, "Size mismatch");
            let err = unsafe { cuda_copy_from_device(data.as_mut_ptr(), self.ptr, self.size) };
            if err != CUDA_SUCCESS {
                return Err(format!("CUDA copy from device failed: {}", err_string(err)));
            }
            Ok(())
        }
        pub fn as_ptr(&self) -> *mut u64 {
            self.ptr
        }
    }
    impl Drop for CudaBuffer {
        fn drop(&mut self) {
            unsafe {
                cuda_free(self.ptr);
            }
        }
    }
    unsafe impl Send for CudaBuffer {}
    unsafe impl Sync for CudaBuffer {}

    fn run(values: &mut [BabyBear], inverse: bool) -> Result<(), String> {
        if !cuda_available() {
            return Err("CUDA not available".to_string());
        }
        let n = values.len();
        assert!(n.is_power_of_two(), "NTT size must be power of 2");
        assert!(n.trailing_zeros() <= 27, "BabyBear only supports NTT up to 2^27");
        let ctx = get_or_create_ctx(n)?;
        let raw = values.as_mut_ptr() as *mut u64;
        // the reference calls the void `ntt_run_inplace` / `intt_run_inplace` (src/ntt.rs:233,248), which cannot fail
        // visibly; the `_rc` forms return the first CUDA error of the call
        match unsafe { if inverse { intt_run_inplace_rc(ctx, raw) } else { ntt_run_inplace_rc(ctx, raw) } } {
            CUDA_SUCCESS => Ok(()),
            e => Err(format!("CUDA NTT failed: {}", err_string(e))),
        }
    }

    pub fn ntt_cuda(values: &mut [BabyBear]) -> Result<(), String> { run(values, false) }
    pub fn intt_cuda(values: &mut [BabyBear]) -> Result<(), String> { run(values, true) }

    fn ck(e: CudaError, what: &str) -> Result<(), String> {
        if e == CUDA_SUCCESS { Ok(()) } else { Err(format!("{what} failed: {}", err_string(e))) }
    }

    /// BabyBearDomain::fft with the coset shift fused on the device (src/math/domain.rs:107-123).
    pub fn coset_fft_cuda(coeffs: &[BabyBear], size: usize, shift: BabyBear) -> Result<Vec<BabyBear>, String> {
        let mut out = vec![BabyBear::zero(); size];
        ck(unsafe { toyni_domain_fft(coeffs.as_ptr() as *const u64, coeffs.len(), size, shift.value, out.as_mut_ptr() as *mut u64) }, "CUDA NTT")?;
        Ok(out)
    }
    /// BabyBearDomain::ifft (src/math/domain.rs:85-102).
    pub fn coset_ifft_cuda(evals: &[BabyBear], shift: BabyBear) -> Result<Vec<BabyBear>, String> {
        let mut out = vec![BabyBear::zero(); evals.len()];
        ck(unsafe { toyni_domain_ifft(evals.as_ptr() as *const u64, evals.len(), shift.value, out.as_mut_ptr() as *mut u64) }, "CUDA INTT")?;
        Ok(out)
    }
    pub fn coset_fft_ext_cuda(coeffs: &[Ext], size: usize, shift: BabyBear) -> Result<Vec<Ext>, String> {
        let mut out = vec![Ext::zero(); size];
        ck(unsafe { toyni_domain_fft_ext(coeffs.as_ptr() as *const u64, coeffs.len(), size, shift.value, out.as_mut_ptr() as *mut u64) }, "CUDA NTT")?;
        Ok(out)
    }
    pub fn coset_ifft_ext_cuda(evals: &[Ext], shift: BabyBear) -> Result<Vec<Ext>, String> {
        let mut out = vec![Ext::zero(); evals.len()];
        ck(unsafe { toyni_domain_ifft_ext(evals.as_ptr() as *const u64, evals.len(), shift.value, out.as_mut_ptr() as *mut u64) }, "CUDA INTT")?;
        Ok(out)
    }
    /// fri_fold / fri_fold_ext (src/math/fri.rs:27, :7)
    pub fn fri_fold_cuda(evals: &[BabyBear], xs: &[BabyBear], beta: BabyBear) -> Result<Vec<BabyBear>, String> {
        assert!(evals.len() % 2 == 0, "Evaluations length must be even");
        let mut out = vec![BabyBear::zero(); evals.len() / 2];
        ck(unsafe { toyni_fri_fold(evals.as_ptr() as *const u64, evals.len(), xs.as_ptr() as *const u64, beta.value, out.as_mut_ptr() as *mut u64) }, "CUDA fold")?;
        Ok(out)
    }
    pub fn fri_fold_ext_cuda(evals: &[Ext], xs: &[BabyBear], beta: Ext) -> Result<Vec<Ext>, String> {
        assert!(evals.len() % 2 == 0, "Evaluations length must be even");
        let mut out = vec![Ext::zero(); evals.len() / 2];
        ck(unsafe { toyni_fri_fold_ext(evals.as_ptr() as *const u64, evals.len(), xs.as_ptr() as *const u64, beta.c.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64) }, "CUDA fold")?;
        Ok(out)
    }
    /// build_merkle_tree / build_unsalted_tree (src/fibonacci.rs:340-363).  `want_nodes` = false returns only the root
    /// (no 2 GiB download at 2^25 leaves; openings then come from `merkle_open_batch_cuda` on the device-resident tree);
    /// true returns every level, leaf level first, 32 B each.
    pub fn merkle_commit_cuda(evals: &[BabyBear], salts: Option<&[[u8; 16]]>, want_nodes: bool) -> Result<(Vec<u8>, [u8; 32]), String> {
        let n = evals.len();
        let mut nodes = if want_nodes { vec![0u8; unsafe { bb_merkle_node_count(n) } * 32] } else { Vec::new() };
        let mut root = [0u8; 32];
        let sp = salts.map(|s| s.as_ptr() as *const u8).unwrap_or(std::ptr::null());
        let np = if want_nodes { nodes.as_mut_ptr() } else { std::ptr::null_mut() };
        ck(unsafe { toyni_merkle_commit(evals.as_ptr() as *const u64, n, 1, sp, np, root.as_mut_ptr()) }, "CUDA Merkle commit")?;
        Ok((nodes, root))
    }

    /// The sharded paths over the G GPUs of one box, driven from this process (header section 4).
    pub struct MultiGpu {
        h: *mut c_void,
    }
    unsafe impl Send for MultiGpu {}
    impl MultiGpu {
        pub fn new(ngpus: usize) -> Result<Self, String> {
            let mut h = std::ptr::null_mut();
            ck(unsafe { bb_mg_init(ngpus as i32, &mut h) }, "bb_mg_init")?;
            Ok(Self { h })
        }
        /// One 2^log_n NTT over all devices, natural order in and out on the host (scatter, four-step, gather).
        pub fn ntt(&self, values: &mut [BabyBear], inverse: bool) -> Result<(), String> {
            let n = values.len();
            assert!(n.is_power_of_two() && n.trailing_zeros() <= 27, "NTT size must be a power of 2 up to 2^27");
            ck(unsafe { bb_mg_ntt_host(self.h, values.as_mut_ptr() as *mut u64, n.trailing_zeros(), inverse as i32) }, "multi-GPU NTT")
        }
        /// Device-resident forms: column blocks in, result slabs out (see the header for the layouts).
        pub fn ntt_fourstep(&self, log_n: u32, inverse: bool, d_blocks: &[*mut u32], d_outs: &[*mut u32]) -> Result<(), String> {
            ck(unsafe { bb_mg_ntt_fourstep(self.h, log_n, inverse as i32, d_blocks.as_ptr(), d_outs.as_ptr()) }, "multi-GPU four-step NTT")
        }
        pub fn ntt_batch(&self, log_n: u32, inverse: bool, d_cols: &[*mut u32], ncols: &[usize]) -> Result<(), String> {
            ck(unsafe { bb_mg_ntt_batch(self.h, log_n, inverse as i32, d_cols.as_ptr(), ncols.as_ptr()) }, "multi-GPU batched NTT")
        }
        /// Fold chain on cyclic shards (betas: `limbs` words per fold); returns the number of folds done.
        pub fn fri_chain(&self, log_m: u32, shift: BabyBear, limbs: usize, final_size: usize, betas: &[u32], d_shards: &[*const u32],
                         d_layers_out: &[*mut u32]) -> Result<usize, String> {
            let mut folds = 0usize;
            ck(unsafe { bb_mg_fri_chain(self.h, log_m, shift.value as u32, limbs as i32, final_size, betas.as_ptr(), d_shards.as_ptr(),
                                        d_layers_out.as_ptr(), &mut folds) }, "multi-GPU fold chain")?;
            Ok(folds)
        }
        pub fn sync(&self) -> Result<(), String> { ck(unsafe { bb_mg_sync(self.h) }, "bb_mg_sync") }
    }
    impl Drop for MultiGpu {
        fn drop(&mut self) {
            unsafe { bb_mg_destroy(self.h) }
        }
    }

    /// Device-resident column of the shifted domain (u32 per element): what `generate_proof` keeps in HBM between the
    /// LDE and the query phase when the prover runs on the GPU (src/fibonacci.rs:124-198).
    pub struct DeviceColumn {
        pub ptr: *mut u32,
        pub log_n: u32,
    }
    /// c_evals of src/fibonacci.rs:133-143 from the device-resident trace LDE.
    pub fn fib_constraint_cuda(trace_lde: &DeviceColumn, out: &mut DeviceColumn, blowup: u32, shift: BabyBear, b1: BabyBear, b2: BabyBear) -> Result<(), String> {
        ck(unsafe { bb_fib_constraint_device(trace_lde.ptr, trace_lde.log_n, blowup, shift.value as u32, b1.value as u32, b2.value as u32, out.ptr) }, "CUDA constraint evaluation")
    }
    /// q_evals of :147-150: multiply by the `blowup` distinct values of 1 / Z_H(x_i).
    pub fn fib_quotient_cuda(c_evals: &mut DeviceColumn, inv_zh: &[BabyBear]) -> Result<(), String> {
        let tab: Vec<u32> = inv_zh.iter().map(|v| v.value as u32).collect();
        ck(unsafe { bb_scale_periodic_device(c_evals.ptr, 1usize << c_evals.log_n, tab.as_ptr(), tab.len() as u32) }, "CUDA quotient")
    }
    /// d_evals of :186-198.
    pub fn fib_deep_cuda(q: &DeviceColumn, trace_lde: &DeviceColumn, out: &mut DeviceColumn, blowup: u32, shift: BabyBear, z: BabyBear,
                         q_z: BabyBear, t_z: BabyBear, t_gz: BabyBear, t_ggz: BabyBear) -> Result<(), String> {
        ck(unsafe { bb_fib_deep_device(q.ptr, trace_lde.ptr, trace_lde.log_n, blowup, shift.value as u32, z.value as u32, q_z.value as u32,
                                       t_z.value as u32, t_gz.value as u32, t_ggz.value as u32, out.ptr) }, "CUDA DEEP composition")
    }
    /// Polynomial::evaluate (src/math/polynomial.rs:134-144) of device-resident coefficients.
    pub fn poly_eval_cuda(d_coeffs: *const u32, n: usize, z: BabyBear) -> Result<BabyBear, String> {
        let mut v = 0u32;
        ck(unsafe { bb_poly_eval_device(d_coeffs, n, z.value as u32, &mut v) }, "CUDA polynomial evaluation")?;
        Ok(BabyBear::new(v as u64))
    }
    /// MerkleTree::get_proof (src/merkle.rs:50-80) for a whole query set: (paths, positions), depth entries per query.
    pub fn merkle_open_batch_cuda(d_nodes: *const u8, nleaves: usize, indices: &[u64]) -> Result<(Vec<Vec<[u8; 32]>>, Vec<Vec<bool>>), String> {
        let mut depth = 0usize;
        let mut m = nleaves;
        while m > 1 { m = (m + 1) / 2; depth += 1; }
        let mut paths = vec![0u8; indices.len() * depth * 32];
        let mut pos = vec![0u8; indices.len() * depth];
        ck(unsafe { bb_merkle_open_batch_device(d_nodes, nleaves, indices.as_ptr(), indices.len(), paths.as_mut_ptr(), pos.as_mut_ptr(), &mut depth) },
           "CUDA Merkle openings")?;
        let p = (0..indices.len()).map(|q| (0..depth).map(|d| { let mut h = [0u8; 32]; h.copy_from_slice(&paths[(q * depth + d) * 32..][..32]); h }).collect()).collect();
        let b = (0..indices.len()).map(|q| (0..depth).map(|d| pos[q * depth + d] != 0).collect()).collect();
        Ok((p, b))
    }
    /// Opened values / salts: out[q] = src[indices[q]] for elements of `elem_bytes` bytes.
    pub fn gather_cuda(d_src: *const c_void, elem_bytes: usize, indices: &[u64]) -> Result<Vec<u8>, String> {
        let mut out = vec![0u8; indices.len() * elem_bytes];
        ck(unsafe { bb_gather_device(d_src, elem_bytes, indices.as_ptr(), indices.len(), out.as_mut_ptr() as *mut c_void) }, "CUDA gather")?;
        Ok(out)
    }
}

#[cfg(feature = "cuda")]
pub use cuda::{
    coset_fft_cuda, coset_fft_ext_cuda, coset_ifft_cuda, coset_ifft_ext_cuda, cuda_available, fri_fold_cuda,
    fib_constraint_cuda, fib_deep_cuda, fib_quotient_cuda, fri_fold_ext_cuda, gather_cuda, intt_cuda, merkle_commit_cuda,
    merkle_open_batch_cuda, ntt_cuda, poly_eval_cuda, CudaBuffer, DeviceColumn, MultiGpu,
};
