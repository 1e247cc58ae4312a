//! Drop-in `SequenceAligner` for biogarden backed by the B200 engine (libbgalign.so).
//!
//! Same public surface as `biogarden::alignment::aligner::SequenceAligner`
//! (reference src/alignment/aligner.rs:44,84,150,216,290,351) plus the batched entry points.
//! NOT COMPILED in this repository's image (no Rust toolchain); kept as the binding a maintainer
//! would add.  The C++ and Python mirrors in this repo exercise the same C ABI.
mod ffi;

use biogarden::ds::sequence::Sequence;
use biogarden::ds::tile::Tile;
use biogarden::error::{BioError, Result};
use std::ptr;

#[derive(Clone, Copy)]
pub enum Mode { Global, Local, Semiglobal, Fitting, Overlap }

pub struct SequenceAligner { ctx: *mut ffi::bg_ctx }

// `&mut self` on every method mirrors the reference: one call at a time per aligner.
unsafe impl Send for SequenceAligner {}

impl SequenceAligner {
    /// aligner.rs:44 -- engine on the current CUDA device.
    pub fn new() -> SequenceAligner { Self::with_devices(&[]) }

    /// Shard batches over several GPUs of the box (no collective: pairs are independent).
    pub fn with_devices(devices: &[i32]) -> SequenceAligner {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { ffi::bg_create(if devices.is_empty() { ptr::null() } else { devices.as_ptr() }, devices.len() as i32, &mut ctx) };
        assert!(rc == ffi::BG_OK, "bg_create failed: no usable CUDA device (the engine has no CPU path)");
        SequenceAligner { ctx }
    }

    pub fn global_alignment(&mut self, seq1: &Sequence, seq2: &Sequence, score: &dyn Fn(&u8, &u8) -> i32, a: i32, b: i32)
        -> Result<(i32, Sequence, Sequence)> { self.single(Mode::Global, seq1, seq2, score, a, b) }
    pub fn local_alignment(&mut self, seq1: &Sequence, seq2: &Sequence, score: &dyn Fn(&u8, &u8) -> i32, a: i32, b: i32)
        -> Result<(i32, Sequence, Sequence)> { self.single(Mode::Local, seq1, seq2, score, a, b) }
    pub fn fitting_alignment(&mut self, seq1: &Sequence, seq2: &Sequence, score: &dyn Fn(&u8, &u8) -> i32, a: i32, b: i32)
        -> Result<(i32, Sequence, Sequence)> { self.single(Mode::Fitting, seq1, seq2, score, a, b) }
    pub fn overlap_alignment(&mut self, seq1: &Sequence, seq2: &Sequence, score: &dyn Fn(&u8, &u8) -> i32, a: i32, b: i32)
        -> Result<(i32, Sequence, Sequence)> { self.single(Mode::Overlap, seq1, seq2, score, a, b) }
    pub fn semiglobal_alignment(&mut self, seq1: &Sequence, seq2: &Sequence, score: &dyn Fn(&u8, &u8) -> i32, a: i32, b: i32)
        -> Result<(i32, Sequence, Sequence)> { self.single(Mode::Semiglobal, seq1, seq2, score, a, b) }

    /// New: every pair (tile[2p], tile[2p+1]) in one call.
    pub fn align_batch(&mut self, pairs: &Tile, mode: Mode, score: &dyn Fn(&u8, &u8) -> i32, a: i32, b: i32)
        -> Result<Vec<(i32, Sequence, Sequence)>> {
        if pairs.len() % 2 != 0 { return Err(BioError::InvalidInputSize); }
        // reference order of checks: sign (aligner.rs:87-89,153-155,219-221), then fitting size (aligner.rs:223-225)
        if matches!(mode, Mode::Global | Mode::Local | Mode::Fitting) && (a > 0 || b > 0) { return Err(BioError::InvalidArgumentRange); }
        let n_pairs = pairs.len() / 2;
        if matches!(mode, Mode::Fitting) {
            for p in 0..n_pairs { if pairs[2 * p].len() < pairs[2 * p + 1].len() { return Err(BioError::InvalidInputSize); } }
        }
        // Tile -> one residue arena + offsets
        let mut residues: Vec<u8> = Vec::new();
        let mut off: Vec<u64> = vec![0];
        for s in pairs { residues.extend_from_slice(&s.chain); off.push(residues.len() as u64); }
        // score callback -> dense table over the residues present (never called on absent bytes)
        let (mut in_a, mut in_b) = ([false; 256], [false; 256]);
        for p in 0..n_pairs { for x in &pairs[2 * p].chain { in_a[*x as usize] = true; } for y in &pairs[2 * p + 1].chain { in_b[*y as usize] = true; } }
        let rows: Vec<u8> = (0..=255u8).filter(|x| in_a[*x as usize]).collect();
        let cols: Vec<u8> = (0..=255u8).filter(|y| in_b[*y as usize]).collect();
        let (mut row_code, mut col_code) = ([0xFFu8; 256], [0xFFu8; 256]);
        for (i, x) in rows.iter().enumerate() { row_code[*x as usize] = i as u8; }
        for (j, y) in cols.iter().enumerate() { col_code[*y as usize] = j as u8; }
        let (nr, nc) = (rows.len().max(1), cols.len().max(1));
        let mut table = vec![0i32; nr * nc];
        for (i, x) in rows.iter().enumerate() { for (j, y) in cols.iter().enumerate() { table[i * nc + j] = score(x, y); } }

        let batch = ffi::bg_batch { n_pairs: n_pairs as u64, residues: residues.as_ptr(), seq_off: off.as_ptr(), packing: 0, reserved_: 0, alphabet: std::ptr::null() };
        let params = ffi::bg_params { mode: mode as i32, gap_open: a, gap_extend: b, flags: 0, table: table.as_ptr(),
                                      n_rows: nr as i32, n_cols: nc as i32, row_code: row_code.as_ptr(), col_code: col_code.as_ptr() };
        // Compact results: the engine returns WHAT backtrack() emitted (one 2-bit op per column); every pair is expanded
        // straight into the two Vec<u8> its Sequences own -- no intermediate arena, 0.3 instead of 2 bytes per column over
        // the host link.  (With rayon the loop below is `into_par_iter()`; bg_expand_ops is thread safe.)
        let mut res: ffi::bg_ops_result = unsafe { std::mem::zeroed() };
        let rc = unsafe { ffi::bg_align_batch_ops(self.ctx, &batch, &params, &mut res) };
        match rc {
            ffi::BG_OK => {}
            ffi::BG_EINVAL_RANGE => return Err(BioError::InvalidArgumentRange),
            ffi::BG_EINVAL_SIZE => return Err(BioError::InvalidInputSize),
            _ => panic!("bgalign: {}", unsafe { std::ffi::CStr::from_ptr(ffi::bg_strerror(rc)) }.to_string_lossy()),
        }
        let mut out = Vec::with_capacity(n_pairs);
        unsafe {
            for p in 0..n_pairs {
                // where the reference itself panics / hangs, panic like it does (status != OK)
                assert!(*res.status.add(p) == ffi::BG_ST_OK, "alignment undefined in the reference for pair {}", p);
                let len = *res.len.add(p) as usize;
                let (mut a_al, mut b_al) = (vec![0u8; len], vec![0u8; len]);
                let s1 = residues.as_ptr().add(off[2 * p] as usize + *res.first.add(2 * p) as usize);
                let s2 = residues.as_ptr().add(off[2 * p + 1] as usize + *res.first.add(2 * p + 1) as usize);
                ffi::bg_expand_ops(s1, s2, res.ops.add(*res.ops_off.add(p) as usize), len as u64, a_al.as_mut_ptr(), b_al.as_mut_ptr());
                out.push((*res.score.add(p), Sequence::from(a_al), Sequence::from(b_al)));
            }
            ffi::bg_ops_result_free(&mut res);
        }
        Ok(out)
    }

    /// analysis::seq::edit_distance for every pair (seq.rs:105-130).
    pub fn edit_distance_batch(&mut self, pairs: &Tile) -> Result<Vec<usize>> {
        if pairs.len() % 2 != 0 { return Err(BioError::InvalidInputSize); }
        let mut residues: Vec<u8> = Vec::new();
        let mut off: Vec<u64> = vec![0];
        for s in pairs { residues.extend_from_slice(&s.chain); off.push(residues.len() as u64); }
        let batch = ffi::bg_batch { n_pairs: (pairs.len() / 2) as u64, residues: residues.as_ptr(), seq_off: off.as_ptr(), packing: 0, reserved_: 0, alphabet: std::ptr::null() };
        let mut out = vec![0u64; pairs.len() / 2];
        let rc = unsafe { ffi::bg_edit_distance_batch(self.ctx, &batch, out.as_mut_ptr()) };
        assert!(rc == ffi::BG_OK);
        Ok(out.into_iter().map(|d| d as usize).collect())
    }

    fn single(&mut self, mode: Mode, seq1: &Sequence, seq2: &Sequence, score: &dyn Fn(&u8, &u8) -> i32, a: i32, b: i32)
        -> Result<(i32, Sequence, Sequence)> {
        let mut t = Tile::new();
        t.push(seq1.clone());
        t.push(seq2.clone());
        Ok(self.align_batch(&t, mode, score, a, b)?.remove(0))
    }
}

impl Default for SequenceAligner { fn default() -> Self { Self::new() } }
impl Drop for SequenceAligner { fn drop(&mut self) { unsafe { ffi::bg_destroy(self.ctx) } } }
