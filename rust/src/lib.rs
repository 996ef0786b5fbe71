//! Drop-in facade for `halo2_dynamic_sha256::Sha256DynamicConfig` (reference src/lib.rs:38-369) whose witness values
//! come from the B200 engine.  SOURCE ONLY (never compiled in the build image: no cargo).
//!
//! * keygen / MockProver shape pass: `digest` replays the static plan (selectors, fixed cells, copy constraints from
//!   `h2sha_get_shape`) and assigns `Value::unknown()`.
//! * proving: `digest_batch` generates the advice columns of many instances on the GPU; `assign_instance` bulk-loads
//!   one instance's columns into a region with the same (column,row) the reference would have used.
#![allow(non_camel_case_types, dead_code)]
include!(concat!(env!("OUT_DIR"), "/bindings.rs"));

use std::ffi::CStr;

#[derive(Debug)]
pub enum Error {
    /// the reference would `assert!`-panic (lib.rs:89-90)
    ReferencePanic(String),
    Engine(i32, String),
}

fn check(rc: i32) -> Result<(), Error> {
    if rc == H2SHA_OK as i32 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(h2sha_last_error()) }.to_string_lossy().into_owned();
    if rc == H2SHA_EPANIC { Err(Error::ReferencePanic(msg)) } else { Err(Error::Engine(rc, msg)) }
}

/// `AssignedHashResult` (lib.rs:31-36) as gate-stream indices; map to cells with `Sha256DynamicConfig::cell_position`.
pub struct AssignedHashResult {
    pub input_len: u32,
    pub input_bytes: Vec<u32>,
    pub output_bytes: Vec<u32>,
}

pub struct Sha256DynamicConfig {
    engine: *mut h2sha_engine_t,
    pub max_variable_byte_sizes: Vec<usize>,
    pub cur_hash_idx: usize,
    layout: h2sha_layout_t,
}

impl Sha256DynamicConfig {
    /// lib.rs:49-69.  `max_rows` = `range.gate.max_rows`, `lookup_bits` = the RangeConfig's (lib.rs:409-418).
    pub fn configure(
        max_variable_byte_sizes: Vec<usize>,
        max_rows: usize,
        lookup_bits: usize,
        num_bits_lookup: usize,
        num_advice_columns: usize,
        is_input_range_check: bool,
        device: i32,
    ) -> Result<Self, Error> {
        let sizes: Vec<u32> = max_variable_byte_sizes.iter().map(|&x| x as u32).collect();
        let cfg = h2sha_config_t {
            n_digests: sizes.len() as u32,
            max_variable_byte_sizes: sizes.as_ptr(),
            max_rows: max_rows as u32,
            lookup_bits: lookup_bits as u32,
            num_bits_lookup: num_bits_lookup as u32,
            num_advice_columns: num_advice_columns as u32,
            is_input_range_check: is_input_range_check as u32,
            gate_col_rows: 0,
            lookup_col_rows: 0,
            spread_rows: 0,
            device,
            build_shape: 1,
            block_parts: 0,
        };
        let mut engine = std::ptr::null_mut();
        check(unsafe { h2sha_create(&cfg, &mut engine) })?;
        let mut layout: h2sha_layout_t = unsafe { std::mem::zeroed() };
        check(unsafe { h2sha_get_layout(engine, &mut layout) })?;
        Ok(Self { engine, max_variable_byte_sizes, cur_hash_idx: 0, layout })
    }

    /// `digest` (lib.rs:71-76) for a whole batch.  `msgs[i][d]` is the input of the d-th digest call of instance i.
    /// Device buffers (`gate`, `lookup`, `spread`: `cudarc::driver::CudaSlice<u64>` raw pointers) must hold
    /// `n * layout.{gate,lookup,spread}_bytes`.  Digests and checksums come back on the host.
    pub fn digest_batch(
        &mut self,
        msgs: &[Vec<&[u8]>],
        precomputed_input_lens: Option<&[Vec<usize>]>,
        gate: u64,
        lookup: u64,
        spread: u64,
        stream: *mut std::ffi::c_void,
    ) -> Result<(Vec<[u8; 32]>, Vec<[u64; 4]>), Error> {
        let d = self.max_variable_byte_sizes.len();
        let n = msgs.len();
        let mut blob = Vec::new();
        let (mut offs, mut lens, mut pre) = (Vec::new(), Vec::new(), Vec::new());
        for (i, inst) in msgs.iter().enumerate() {
            assert_eq!(inst.len(), d);
            for (k, m) in inst.iter().enumerate() {
                offs.push(blob.len() as u64);
                lens.push(m.len() as u32);
                pre.push(precomputed_input_lens.map(|p| p[i][k] as u32).unwrap_or(0));
                blob.extend_from_slice(m);
            }
        }
        let mut digests = vec![[0u8; 32]; n * d];
        let mut cks = vec![[0u64; 4]; n];
        let batch = h2sha_batch_t {
            n_instances: n as u64,
            msgs: blob.as_ptr(),
            msgs_on_device: 0,
            msgs_bytes: blob.len() as u64,
            offsets: offs.as_ptr(),
            lens: lens.as_ptr(),
            precomputed_lens: if precomputed_input_lens.is_some() { pre.as_ptr() } else { std::ptr::null() },
            gate: gate as *mut _,
            lookup: lookup as *mut _,
            spread: spread as *mut _,
            digests_dev: std::ptr::null_mut(),
            checksums_dev: std::ptr::null_mut(),
            digests_host: digests.as_mut_ptr() as *mut u8,
            checksums_host: cks.as_mut_ptr() as *mut u64,
            stream,
            reuse_inputs: 0,
            time_kernels: 0,
        };
        check(unsafe { h2sha_digest_batch(self.engine, &batch) })?;
        // caller synchronises `stream` before reading digests / cks
        Ok((digests, cks))
    }

    /// Handles of the `cur_hash_idx`-th digest (lib.rs:342-347); advances `cur_hash_idx` like the reference.
    pub fn handles(&mut self) -> Result<AssignedHashResult, Error> {
        let d = self.cur_hash_idx;
        let mut input_len = 0u32;
        let mut input_bytes = vec![0u32; self.max_variable_byte_sizes[d]];
        let mut output_bytes = vec![0u32; 32];
        check(unsafe { h2sha_get_handles(self.engine, d as u32, &mut input_len, input_bytes.as_mut_ptr(), output_bytes.as_mut_ptr()) })?;
        self.cur_hash_idx += 1;
        Ok(AssignedHashResult { input_len, input_bytes, output_bytes })
    }

    /// gate-stream index -> (advice column, row)
    pub fn cell_position(&self, breaks: &[u32], idx: u32) -> (usize, usize) {
        let col = breaks.partition_point(|&b| b <= idx) - 1;
        (col, (idx - breaks[col]) as usize)
    }

    pub fn layout(&self) -> &h2sha_layout_t { &self.layout }

    /// Lookup-argument pre-work on the batch that is in HBM: table-row multiplicities of the range lookup(s)
    /// (halo2-base RangeConfig, lib.rs:409-418,469) and the spread lookups (spread.rs:53-62).  `mult` is a device
    /// buffer of `n * lookup_info().mult_words_per_instance` u32; `not_in_table` a device u32 (or 0).
    pub fn lookup_multiplicities(&mut self, n: usize, lookup: u64, spread: u64, usable_rows: u32, mult: u64, not_in_table: u64,
                                 stream: *mut std::ffi::c_void) -> Result<(), Error> {
        check(unsafe {
            h2sha_lookup_multiplicities(self.engine, n as u64, lookup as *const _, spread as *const _, usable_rows, mult as *mut u32,
                                        not_in_table as *mut u32, stream)
        })
    }

    /// The permuted pair (A', S') halo2's `permute_expression_pair` would build for lookup `lookup_idx` of every
    /// instance (`theta`: the transcript challenge in Montgomery limbs, only for the two-expression spread lookups).
    /// Outputs: device buffers of `n * usable_rows` Fr each; the blinding rows stay with the prover.
    pub fn permute_lookup(&mut self, n: usize, lookup_idx: u32, mult: u64, usable_rows: u32, theta: Option<&[u64; 4]>, permuted_input: u64,
                          permuted_table: u64, errors: u64, stream: *mut std::ffi::c_void) -> Result<(), Error> {
        check(unsafe {
            h2sha_permute_lookup(self.engine, n as u64, lookup_idx, mult as *const u32, usable_rows,
                                 theta.map(|t| t.as_ptr()).unwrap_or(std::ptr::null()), permuted_input as *mut _, permuted_table as *mut _,
                                 errors as *mut u32, stream)
        })
    }

    /// `MockProver::run(..).verify()` (lib.rs:525-526) for every instance of a batch, on the device: violation counts
    /// [gates, copies, range lookups, spread lookups, digest bytes]; all zero = the witness the reference's tests accept.
    pub fn check_batch(&mut self, n: usize, gate: u64, lookup: u64, spread: u64, digests_dev: u64, stream: *mut std::ffi::c_void) -> Result<[u64; 5], Error> {
        let mut v = [0u64; 5];
        check(unsafe { h2sha_check_batch(self.engine, n as u64, gate as *const _, lookup as *const _, spread as *const _, digests_dev as *const u8, v.as_mut_ptr(), stream) })?;
        Ok(v)
    }

    pub fn lookup_info(&self) -> Result<h2sha_lookup_info_t, Error> {
        let mut li: h2sha_lookup_info_t = unsafe { std::mem::zeroed() };
        check(unsafe { h2sha_get_lookup_info(self.engine, &mut li) })?;
        Ok(li)
    }
}

/// The path's only collective: all-gather of digests and per-instance checksums over the caller's `ncclComm_t`
/// (one per GPU).  Receive buffers are rank-major and hold `n_ranks` times the send size.
pub fn gather(comm: *mut std::ffi::c_void, n_per_rank: usize, n_digests: u32, digests: u64, checksums: u64, all_digests: u64, all_checksums: u64,
              stream: *mut std::ffi::c_void) -> Result<(), Error> {
    check(unsafe { h2sha_gather(comm, n_per_rank as u64, n_digests, digests as *const u8, checksums as *const u64, all_digests as *mut u8, all_checksums as *mut u64, stream) })
}

impl Drop for Sha256DynamicConfig {
    fn drop(&mut self) {
        unsafe { h2sha_destroy(self.engine) }
    }
}
