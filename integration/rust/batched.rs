// bioshell-seq/src/alignment/batched.rs -- SOURCE ONLY (no rustc in the build image).
//
// The new batched entry points next to `align_all_pairs`
// (bioshell-seq/src/alignment/alignment_protocols.rs:83-115), behind the existing module:
// add `mod batched; pub use batched::*;` to bioshell-seq/src/alignment/mod.rs:15-20.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

use crate::alignment::{aligned_sequences, AlignmentPath, AlignmentReporter};
use crate::scoring::{SubstitutionMatrix, SubstitutionMatrixList};
use crate::sequence::Sequence;

#[repr(C)]
pub struct BsaCtx { _private: [u8; 0] }

extern "C" {
    fn bsa_create(device_id: c_int) -> *mut BsaCtx;
    fn bsa_create_multi(device_ids: *const c_int, n_dev: c_int) -> *mut BsaCtx;
    fn bsa_gather_sequences(ctx: *mut BsaCtx, src_set: c_int, dst_set: c_int, idx: *const u32, n: u32) -> c_int;
    fn bsa_destroy(ctx: *mut BsaCtx);
    fn bsa_last_error(ctx: *const BsaCtx) -> *const c_char;
    fn bsa_set_scoring(ctx: *mut BsaCtx, score: *const i32, aa_index: *const u8, gap_open: i32, gap_extend: i32) -> c_int;
    fn bsa_load_sequences(ctx: *mut BsaCtx, set_id: c_int, residues: *const u8, offsets: *const u64, n: u32) -> c_int;
    fn bsa_align_all_pairs(ctx: *mut BsaCtx, q_set: c_int, t_set: c_int, q_counts: *const u32, t_begin: u32,
                           t_end: u32, flags: u32, scores: *mut i32, n_identical: *mut u32, n_results: *mut u64) -> c_int;
    fn bsa_align_pairs_paths(ctx: *mut BsaCtx, q_set: c_int, t_set: c_int, q_idx: *const u32, t_idx: *const u32,
                             n_pairs: u64, scores: *mut i32, n_identical: *mut u32, path_buf: *mut u8,
                             path_off: *mut u64) -> c_int;
    fn bsa_all_vs_all(ctx: *mut BsaCtx, set_id: c_int, flags: u32, scores: *mut i32, n_identical: *mut u32) -> c_int;
    fn bsa_one_vs_many(ctx: *mut BsaCtx, q_set: c_int, db_set: c_int, flags: u32, scores: *mut i32,
                       n_identical: *mut u32) -> c_int;
    fn bsa_plan_shards(ctx: *mut BsaCtx, q_set: c_int, t_set: c_int, q_counts: *const u32, n_shards: u32,
                       bounds: *mut u32) -> c_int;
    fn bsa_local_align_pairs(ctx: *mut BsaCtx, q_set: c_int, t_set: c_int, q_idx: *const u32, t_idx: *const u32,
                             n_pairs: u64, scores: *mut i32, end_q: *mut u32, end_t: *mut u32, start_q: *mut u32,
                             start_t: *mut u32, path_buf: *mut u8, path_off: *mut u64) -> c_int;
    fn bsa_hclust(ctx: *mut BsaCtx, n: u32, dist: *const f32, linkage: c_int, flags: u32, mat_i: *mut u32,
                  mat_j: *mut u32, merge_dist: *mut f32) -> c_int;
}

const BSA_WANT_SCORE: u32 = 1;
const BSA_WANT_IDENTICAL: u32 = 2;

/// Scores and identical-residue counts in the report order of `align_all_pairs`
/// (template-major: for t { for q { .. } }, alignment_protocols.rs:94-102).
pub struct PairResults {
    pub scores: Vec<i32>,
    pub n_identical: Vec<u32>,
    /// number of queries aligned against template t (the position of the triangle `break`)
    pub q_counts: Vec<u32>,
}

pub struct GpuAligner { ctx: *mut BsaCtx }

impl GpuAligner {
    pub fn new(device: i32) -> Result<Self, String> {
        let ctx = unsafe { bsa_create(device) };
        if ctx.is_null() {
            return Err(unsafe { CStr::from_ptr(bsa_last_error(std::ptr::null())) }.to_string_lossy().into());
        }
        Ok(GpuAligner { ctx })
    }

    /// One context over every visible GPU (n_dev = 0) or the listed ones: the caller stays one process
    /// (bin/cluster_sequences.rs:173-177) and every entry point returns the same bytes as on one device.
    pub fn new_multi(devices: &[i32]) -> Result<Self, String> {
        let ctx = unsafe { bsa_create_multi(devices.as_ptr(), devices.len() as c_int) };
        if ctx.is_null() {
            return Err(unsafe { CStr::from_ptr(bsa_last_error(std::ptr::null())) }.to_string_lossy().into());
        }
        Ok(GpuAligner { ctx })
    }

    /// Sequence idx[i] of `src_set` becomes sequence i of `dst_set`, on the device (bucket clustering's
    /// candidate-vs-representatives batches, bucket_clustering.rs:272-309).
    pub fn gather(&self, src_set: i32, dst_set: i32, idx: &[u32]) -> Result<(), String> {
        self.check(unsafe { bsa_gather_sequences(self.ctx, src_set, dst_set, idx.as_ptr(), idx.len() as u32) })
    }

    fn check(&self, rc: c_int) -> Result<(), String> {
        if rc == 0 { Ok(()) } else {
            Err(unsafe { CStr::from_ptr(bsa_last_error(self.ctx)) }.to_string_lossy().into())
        }
    }

    fn load(&self, set_id: i32, seqs: &[Sequence]) -> Result<(), String> {
        let mut residues: Vec<u8> = Vec::new();
        let mut offsets: Vec<u64> = vec![0];
        for s in seqs {
            residues.extend_from_slice(s.as_u8());
            offsets.push(residues.len() as u64);
        }
        self.check(unsafe { bsa_load_sequences(self.ctx, set_id, residues.as_ptr(), offsets.as_ptr(), seqs.len() as u32) })
    }

    /// The inner-loop trip count of alignment_protocols.rs:96-97 for every template.
    /// `template == query` is `#[derive(PartialEq)]` on Sequence (description AND residues, sequence.rs:8);
    /// the first query equal to each template is looked up in a hash map keyed by (description, residues)
    /// -- one pass over the queries instead of a scan per template (5e9 compares at 100k sequences).
    fn triangle_counts(queries: &[Sequence], templates: &[Sequence], triangle: bool) -> Vec<u32> {
        if !triangle { return vec![queries.len() as u32; templates.len()]; }
        let mut first: std::collections::HashMap<(&str, &[u8]), u32> = std::collections::HashMap::with_capacity(queries.len());
        for (i, q) in queries.iter().enumerate() {
            first.entry((q.description(), q.as_u8())).or_insert(i as u32);
        }
        templates.iter().map(|t| *first.get(&(t.description(), t.as_u8())).unwrap_or(&(queries.len() as u32))).collect()
    }

    /// Batched replacement of the `align_all_pairs` double loop: no strings are built.
    pub fn align_pairs_batched(&self, queries: &[Sequence], templates: &[Sequence], matrix: SubstitutionMatrixList,
                               gap_open: i32, gap_extend: i32, if_triangle_only: bool) -> Result<PairResults, String> {
        let m = SubstitutionMatrix::load(matrix);
        let mut aa = [0u8; 256];
        aa[..255].copy_from_slice(&m.aa_indexes);          // pub(crate) fields, substitution_matrix.rs:31-32
        self.check(unsafe { bsa_set_scoring(self.ctx, m.score.as_ptr(), aa.as_ptr(), gap_open, gap_extend) })?;
        self.load(0, queries)?;
        self.load(1, templates)?;
        let q_counts = Self::triangle_counts(queries, templates, if_triangle_only);
        let n: u64 = q_counts.iter().map(|&c| c as u64).sum();
        let mut scores = vec![0i32; n as usize];
        let mut nid = vec![0u32; n as usize];
        let mut n_res = 0u64;
        self.check(unsafe {
            bsa_align_all_pairs(self.ctx, 0, 1, q_counts.as_ptr(), 0, templates.len() as u32,
                                BSA_WANT_SCORE | BSA_WANT_IDENTICAL, scores.as_mut_ptr(), nid.as_mut_ptr(), &mut n_res)
        })?;
        Ok(PairResults { scores, n_identical: nid, q_counts })
    }

    /// Drop-in for `align_all_pairs` with a reporter: same pairs, same (t-major) order, the
    /// aligned `Sequence`s built from the GPU traceback paths.
    pub fn align_all_pairs<R: AlignmentReporter>(&self, queries: &Vec<Sequence>, templates: &Vec<Sequence>,
            matrix: SubstitutionMatrixList, gap_open: i32, gap_extend: i32, if_triangle_only: bool,
            reporter: &mut R) -> Result<(), String> {
        let m = SubstitutionMatrix::load(matrix);
        let mut aa = [0u8; 256];
        aa[..255].copy_from_slice(&m.aa_indexes);
        self.check(unsafe { bsa_set_scoring(self.ctx, m.score.as_ptr(), aa.as_ptr(), gap_open, gap_extend) })?;
        self.load(0, queries)?;
        self.load(1, templates)?;
        let counts = Self::triangle_counts(queries, templates, if_triangle_only);
        for (t, &cnt) in counts.iter().enumerate() {
            if cnt == 0 { continue; }
            let q_idx: Vec<u32> = (0..cnt).collect();
            let t_idx: Vec<u32> = vec![t as u32; cnt as usize];
            let cap: usize = q_idx.iter().map(|&q| queries[q as usize].len() + templates[t].len()).sum();
            let mut path_buf = vec![0u8; cap];
            let mut path_off = vec![0u64; cnt as usize + 1];
            self.check(unsafe {
                bsa_align_pairs_paths(self.ctx, 0, 1, q_idx.as_ptr(), t_idx.as_ptr(), cnt as u64, std::ptr::null_mut(),
                                      std::ptr::null_mut(), path_buf.as_mut_ptr(), path_off.as_mut_ptr())
            })?;
            for q in 0..cnt as usize {
                let glyphs = std::str::from_utf8(&path_buf[path_off[q] as usize..path_off[q + 1] as usize]).unwrap();
                let path = AlignmentPath::try_from(glyphs).unwrap();      // alignment_path.rs:88-105
                let (ali_q, ali_t) = aligned_sequences(&path, &queries[q], &templates[t], '-');
                reporter.report(&ali_q, &ali_t);                          // alignment_protocols.rs:102
            }
        }
        Ok(())
    }
}

/// One local alignment as `LocalAlignment::{recent_score, recent_end_point, backtrace}` report it
/// (bioshell-seq/src/alignment/local.rs:205-284).
pub struct LocalHit {
    pub score: i32,
    pub end: (u32, u32),
    pub start: (u32, u32),
    pub path: AlignmentPath,
}

/// Linkage codes of `bsa_hclust`, in the order of bioshell-clustering/src/hierarchical/strategies/mod.rs:25-92.
#[derive(Clone, Copy)]
pub enum Linkage { Single = 0, Complete = 1, Average = 2, Median = 3, Centroid = 4, Ward = 5 }

impl GpuAligner {
    /// Score-only search of every query against every database sequence (needleman_wunsh.rs:108-109
    /// without the strings): result k = t * |queries| + q.
    pub fn align_one_vs_many(&self, queries: &[Sequence], database: &[Sequence], matrix: SubstitutionMatrixList,
                             gap_open: i32, gap_extend: i32) -> Result<Vec<i32>, String> {
        let m = SubstitutionMatrix::load(matrix);
        let mut aa = [0u8; 256];
        aa[..255].copy_from_slice(&m.aa_indexes);
        self.check(unsafe { bsa_set_scoring(self.ctx, m.score.as_ptr(), aa.as_ptr(), gap_open, gap_extend) })?;
        self.load(0, queries)?;
        self.load(1, database)?;
        let mut scores = vec![0i32; queries.len() * database.len()];
        self.check(unsafe { bsa_one_vs_many(self.ctx, 0, 1, BSA_WANT_SCORE, scores.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok(scores)
    }

    /// Strict upper triangle of one loaded set (the `cluster_sequences` call): k = t (t - 1) / 2 + q.
    pub fn all_vs_all_loaded(&self, n: usize) -> Result<(Vec<i32>, Vec<u32>), String> {
        let pairs = n * (n - 1) / 2;
        let (mut scores, mut nid) = (vec![0i32; pairs], vec![0u32; pairs]);
        self.check(unsafe {
            bsa_all_vs_all(self.ctx, 0, BSA_WANT_SCORE | BSA_WANT_IDENTICAL, scores.as_mut_ptr(), nid.as_mut_ptr())
        })?;
        Ok((scores, nid))
    }

    /// Cell-balanced template ranges for `n_shards` processes, one GPU each; shard r then calls
    /// `bsa_align_all_pairs(.., bounds[r], bounds[r + 1], ..)`.  No collective is involved.
    pub fn plan_shards(&self, q_counts: &[u32], n_shards: u32) -> Result<Vec<u32>, String> {
        let mut bounds = vec![0u32; n_shards as usize + 1];
        self.check(unsafe { bsa_plan_shards(self.ctx, 0, 1, q_counts.as_ptr(), n_shards, bounds.as_mut_ptr()) })?;
        Ok(bounds)
    }

    /// `LocalAlignment::align` + `backtrace` + `recent_end_point` for a list of (query, template) index pairs
    /// of the loaded sets 0 and 1.
    pub fn local_align_pairs(&self, queries: &[Sequence], templates: &[Sequence], pairs: &[(u32, u32)])
            -> Result<Vec<LocalHit>, String> {
        let n = pairs.len();
        let q_idx: Vec<u32> = pairs.iter().map(|p| p.0).collect();
        let t_idx: Vec<u32> = pairs.iter().map(|p| p.1).collect();
        let cap: usize = pairs.iter().map(|p| queries[p.0 as usize].len() + templates[p.1 as usize].len()).sum();
        let (mut scores, mut path_buf, mut path_off) = (vec![0i32; n], vec![0u8; cap.max(1)], vec![0u64; n + 1]);
        let (mut eq, mut et, mut sq, mut st) = (vec![0u32; n], vec![0u32; n], vec![0u32; n], vec![0u32; n]);
        self.check(unsafe {
            bsa_local_align_pairs(self.ctx, 0, 1, q_idx.as_ptr(), t_idx.as_ptr(), n as u64, scores.as_mut_ptr(),
                                  eq.as_mut_ptr(), et.as_mut_ptr(), sq.as_mut_ptr(), st.as_mut_ptr(),
                                  path_buf.as_mut_ptr(), path_off.as_mut_ptr())
        })?;
        Ok((0..n).map(|k| {
            let glyphs = std::str::from_utf8(&path_buf[path_off[k] as usize..path_off[k + 1] as usize]).unwrap();
            LocalHit { score: scores[k], end: (eq[k], et[k]), start: (sq[k], st[k]),
                       path: AlignmentPath::try_from(glyphs).unwrap() }
        }).collect())
    }

    /// The merge log of `hierarchical_clustering` (bioshell-clustering/src/hierarchical/hierarchical.rs:22-80):
    /// per step the two matrix indices `closest_elements` returned and their distance; the caller replays
    /// hierarchical.rs:44-75 on it to build the `BinaryTreeNode<HierarchicalCluster>` tree.
    /// `dist` is n x n row-major, only dist[i * n + j] with i > j is read (clustering_matrix.rs:14-19).
    pub fn hclust_merge_log(&self, n: usize, dist: &[f32], linkage: Linkage) -> Result<(Vec<u32>, Vec<u32>, Vec<f32>), String> {
        assert_eq!(dist.len(), n * n);
        let k = n.saturating_sub(1).max(1);
        let (mut mi, mut mj, mut md) = (vec![0u32; k], vec![0u32; k], vec![0f32; k]);
        self.check(unsafe {
            bsa_hclust(self.ctx, n as u32, dist.as_ptr(), linkage as c_int, 0, mi.as_mut_ptr(), mj.as_mut_ptr(),
                       md.as_mut_ptr())
        })?;
        mi.truncate(n.saturating_sub(1));
        mj.truncate(n.saturating_sub(1));
        md.truncate(n.saturating_sub(1));
        Ok((mi, mj, md))
    }
}

impl Drop for GpuAligner {
    fn drop(&mut self) { unsafe { bsa_destroy(self.ctx) } }
}
