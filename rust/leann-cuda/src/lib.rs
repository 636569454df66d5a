//! leann-cuda — `BackendSearcher` implementations backed by libleann_cuda (sm_100a).
//!
//! SOURCE ONLY (no Rust toolchain in the build image). Mirrors include/leann_cuda.h one to one.
//! Drop-in points in leann-rs:
//!   * `src/backend/mod.rs:23-45`  `BackendType::load_searcher` -> `CudaSearcher::load`
//!   * `src/backend/traits.rs:11-30` `trait BackendSearcher`     -> `impl BackendSearcher for CudaSearcher`
//!   * `src/index/searcher.rs:123-210` hybrid path               -> `CudaSearcher::hybrid_search`
use std::ffi::{c_char, c_float, c_int, c_void, CString};
use std::path::Path;

#[repr(C)]
pub struct LeannCudaIndex {
    _p: [u8; 0],
}
#[repr(C)]
pub struct LeannCudaBm25 {
    _p: [u8; 0],
}
#[repr(C)]
pub struct LeannCudaFilter {
    _p: [u8; 0],
}
#[repr(C)]
pub struct LeannCudaMetacols {
    _p: [u8; 0],
}
#[repr(C)]
pub struct LeannCudaShards {
    _p: [u8; 0],
}
#[repr(C)]
pub struct LeannCudaSearcher {
    _p: [u8; 0],
}

pub const BACKEND_HNSW: c_int = 0;
pub const BACKEND_VAMANA: c_int = 1;
pub const BACKEND_FLAT: c_int = 2;
pub const METRIC_DEFAULT: c_int = -1;

extern "C" {
    fn leann_cuda_open(base_path: *const c_char, backend: c_int, dims: usize, metric: c_int, device: c_int,
                       out: *mut *mut LeannCudaIndex, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_len(index: *const LeannCudaIndex) -> usize;
    fn leann_cuda_search(index: *const LeannCudaIndex, queries: *const c_float, nq: usize, k: usize, ef: usize,
                         mask_bits: *const u64, mask_mode: c_int, keys: *mut u64, dists: *mut c_float,
                         counts: *mut u32, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_close(index: *mut LeannCudaIndex);
    fn leann_cuda_hnsw_build(vectors: *const c_float, vectors_on_device: c_int, n: usize, dims: usize, graph_degree: usize,
                             complexity: usize, metric: c_int, seed: u64, device: c_int, out: *mut *mut LeannCudaIndex,
                             err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_hnsw_add(index: *mut LeannCudaIndex, vectors: *const c_float, vectors_on_device: c_int, m: usize,
                           start_id: u64, complexity: usize, seed: u64, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_save(index: *const LeannCudaIndex, base_path: *const c_char, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_bm25_build(docs: *const *const c_char, doc_bytes: *const usize, n_docs: usize, device: c_int,
                             out: *mut *mut LeannCudaBm25, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_bm25_free(b: *mut LeannCudaBm25);
    fn leann_cuda_bm25_dense_rows(b: *const LeannCudaBm25) -> usize;
    fn leann_cuda_filter_parse(expr: *const c_char, out: *mut *mut LeannCudaFilter, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_filter_free(f: *mut LeannCudaFilter);
    fn leann_cuda_metacols_build(metadata_json: *const *const c_char, bytes: *const usize, n: usize,
                                 out: *mut *mut LeannCudaMetacols, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_metacols_mask(cols: *const LeannCudaMetacols, f: *const LeannCudaFilter, mask_bits: *mut u64,
                                err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_metacols_free(cols: *mut LeannCudaMetacols);
    fn leann_cuda_set_coalescing(index: *mut LeannCudaIndex, max_batch: usize, max_wait_us: u32) -> c_int;
    fn leann_cuda_set_visited_hash(index: *mut LeannCudaIndex, capacity: usize) -> c_int;
    fn leann_cuda_write_layout_cache(index: *const LeannCudaIndex, base_path: *const c_char, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_vamana_build(vectors: *const c_float, vectors_on_device: c_int, n: usize, dims: usize, graph_degree: usize,
                               complexity: usize, alpha: c_float, metric: c_int, seed: u64, device: c_int,
                               out: *mut *mut LeannCudaIndex, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_flat_from_host(vectors: *const c_float, n: usize, dims: usize, metric: c_int, device: c_int,
                                 out: *mut *mut LeannCudaIndex, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_shards_open(base_paths: *const *const c_char, n_shards: usize, backend: c_int, dims: usize, metric: c_int,
                              devices: *const c_int, key_offsets: *const u64, exchange: c_int, out: *mut *mut LeannCudaShards,
                              err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_shards_len(shards: *const LeannCudaShards) -> usize;
    fn leann_cuda_shards_search(shards: *mut LeannCudaShards, queries: *const c_float, nq: usize, k: usize, ef: usize,
                                shard_masks: *const *const u64, keys: *mut u64, dists: *mut c_float, counts: *mut u32,
                                err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_shards_close(shards: *mut LeannCudaShards);
    fn leann_cuda_searcher_load(base_path: *const c_char, backend_name: *const c_char, dims: usize, device: c_int,
                                out: *mut *mut LeannCudaSearcher, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_searcher_len(s: *const LeannCudaSearcher) -> usize;
    fn leann_cuda_searcher_id(s: *const LeannCudaSearcher, idx: u64, out: *mut c_char, cap: usize) -> usize;
    fn leann_cuda_searcher_search(s: *const LeannCudaSearcher, queries: *const c_float, query_texts: *const *const c_char,
                                  query_text_bytes: *const usize, nq: usize, top_k: usize, complexity: usize,
                                  filter_expr: *const c_char, hybrid: c_int, alpha: c_float, idx: *mut u64,
                                  scores: *mut c_float, counts: *mut u32, err: *mut c_char, errlen: usize) -> c_int;
    fn leann_cuda_searcher_bm25_search(s: *const LeannCudaSearcher, query: *const c_char, query_bytes: usize, top_k: usize,
                                       idx: *mut u64, scores: *mut c_float, count: *mut u32, err: *mut c_char,
                                       errlen: usize) -> c_int;
    fn leann_cuda_searcher_close(s: *mut LeannCudaSearcher);
    fn leann_cuda_hybrid_search(index: *const LeannCudaIndex, bm25: *const LeannCudaBm25, queries: *const c_float,
                                query_texts: *const *const c_char, query_text_bytes: *const usize, nq: usize,
                                top_k: usize, ef: usize, hybrid: c_int, alpha: c_float, filter_mask: *const u64,
                                idx: *mut u64, scores: *mut c_float, counts: *mut u32, err: *mut c_char,
                                errlen: usize) -> c_int;
}

/// The trait of leann-rs (`src/backend/traits.rs:11-30`), repeated so the crate is self-contained.
pub trait BackendSearcher: Send + Sync {
    fn search(&self, query: &[f32], top_k: usize, complexity: usize) -> anyhow::Result<(Vec<u64>, Vec<f32>)>;
    fn len(&self) -> usize;
    fn is_empty(&self) -> bool {
        self.len() == 0
    }
}

pub struct CudaSearcher {
    handle: *mut LeannCudaIndex,
    dims: usize,
    /// HnswSearcher ignores `complexity` and always runs expansion_search = 64 (hnsw.rs:49,83).
    fixed_ef: Option<usize>,
}
// The library serialises concurrent calls on one handle internally (see include/leann_cuda.h).
unsafe impl Send for CudaSearcher {}
unsafe impl Sync for CudaSearcher {}

fn check(rc: c_int, err: &[u8]) -> anyhow::Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let end = err.iter().position(|&b| b == 0).unwrap_or(err.len());
    anyhow::bail!("{}", String::from_utf8_lossy(&err[..end]))
}

impl CudaSearcher {
    /// `index_path` is the extension-less base, exactly what `load_searcher` receives.
    pub fn load(index_path: &Path, dimensions: usize, backend_name: &str, device: i32) -> anyhow::Result<Self> {
        let (backend, fixed_ef) = match backend_name {
            "hnsw" => (BACKEND_HNSW, Some(64)),
            "diskann" => (BACKEND_VAMANA, None),
            "flat" => (BACKEND_FLAT, None),
            other => anyhow::bail!("Unknown backend: {}", other),
        };
        let base = CString::new(index_path.to_string_lossy().as_bytes())?;
        let mut handle = std::ptr::null_mut();
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_open(base.as_ptr(), backend, dimensions, METRIC_DEFAULT, device, &mut handle,
                            err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok(Self { handle, dims: dimensions, fixed_ef })
    }

    /// Batched form: `queries` is nq x dims row-major. Returns (keys, dists, counts), nq x k each.
    pub fn search_batch(&self, queries: &[f32], top_k: usize, ef: usize) -> anyhow::Result<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        let nq = queries.len() / self.dims;
        let mut keys = vec![u64::MAX; nq * top_k];
        let mut dists = vec![f32::INFINITY; nq * top_k];
        let mut counts = vec![0u32; nq];
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_search(self.handle, queries.as_ptr(), nq, top_k, ef, std::ptr::null(), 0, keys.as_mut_ptr(),
                              dists.as_mut_ptr(), counts.as_mut_ptr(), err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok((keys, dists, counts))
    }
}

impl CudaSearcher {
    /// Request coalescing for `leann serve` (src/cli/serve.rs:260-311 issues one `search(&self)` per HTTP request on a shared
    /// searcher): concurrent single-query calls are merged into batched launches. On by default (max_batch 256, no waiting);
    /// `max_wait_us > 0` lets a leader wait for company, `max_batch <= 1` turns it off.
    pub fn set_coalescing(&mut self, max_batch: usize, max_wait_us: u32) {
        unsafe { leann_cuda_set_coalescing(self.handle, max_batch, max_wait_us) };
    }

    /// Visited-set representation of the traversal (a tuning hook, results never depend on it): 0 = automatic, 1 = byte maps
    /// only, 2..=5 = the shared-memory table forms, >= 1024 = per-warp hash tables of that capacity. Returns false for a
    /// capacity the library rejects (6..=1023).
    pub fn set_visited_hash(&mut self, capacity: usize) -> bool {
        unsafe { leann_cuda_set_visited_hash(self.handle, capacity) == 0 }
    }

    /// Writes `<index_path>.cuda-layout` (the parsed adjacency, bound to the `.index` file by size, mtime and header hash):
    /// the next `load` of the same index streams it instead of parsing usearch's node records (`leann build` would call this
    /// once after `index.save`, src/backend/hnsw.rs:134).
    pub fn write_layout_cache(&self, index_path: &Path) -> anyhow::Result<()> {
        let base = CString::new(index_path.to_string_lossy().as_bytes())?;
        let mut err = [0u8; 1024];
        let rc = unsafe { leann_cuda_write_layout_cache(self.handle, base.as_ptr(), err.as_mut_ptr() as *mut c_char, err.len()) };
        check(rc, &err)
    }

    /// Exact scan over raw embeddings: the scoring + sort + take(k) of `RecomputeSearcher::search`
    /// (src/index/recompute.rs:96-110) with the database resident in HBM (`metric` 3 = raw dot, descending).
    pub fn flat_from_embeddings(embeddings: &[f32], dimensions: usize, device: i32) -> anyhow::Result<Self> {
        let mut handle = std::ptr::null_mut();
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_flat_from_host(embeddings.as_ptr(), embeddings.len() / dimensions, dimensions, 3, device, &mut handle,
                                      err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok(Self { handle, dims: dimensions, fixed_ef: None })
    }
}

impl BackendSearcher for CudaSearcher {
    fn search(&self, query: &[f32], top_k: usize, complexity: usize) -> anyhow::Result<(Vec<u64>, Vec<f32>)> {
        let ef = self.fixed_ef.unwrap_or(complexity);
        let (mut keys, mut dists, counts) = self.search_batch(query, top_k, ef)?;
        keys.truncate(counts[0] as usize);
        dists.truncate(counts[0] as usize);
        Ok((keys, dists))
    }
    fn len(&self) -> usize {
        unsafe { leann_cuda_len(self.handle) }
    }
}

impl Drop for CudaSearcher {
    fn drop(&mut self) {
        unsafe { leann_cuda_close(self.handle) }
    }
}

/// The index split into sub-indexes on several GPUs of one box (`<base>` per shard): `BackendType::load_searcher`
/// (src/backend/mod.rs:23-45) returns this when the index metadata lists shards. Keys are global passage ordinals.
pub struct ShardedCudaSearcher {
    handle: *mut LeannCudaShards,
    dims: usize,
    fixed_ef: Option<usize>,
}
unsafe impl Send for ShardedCudaSearcher {}
unsafe impl Sync for ShardedCudaSearcher {}

impl ShardedCudaSearcher {
    pub fn load(shard_paths: &[&Path], dimensions: usize, backend_name: &str, devices: &[i32]) -> anyhow::Result<Self> {
        anyhow::ensure!(shard_paths.len() == devices.len(), "one device per shard");
        let (backend, fixed_ef) = match backend_name {
            "hnsw" => (BACKEND_HNSW, Some(64)),
            "diskann" => (BACKEND_VAMANA, None),
            "flat" => (BACKEND_FLAT, None),
            other => anyhow::bail!("Unknown backend: {}", other),
        };
        let bases: Vec<CString> = shard_paths.iter().map(|p| CString::new(p.to_string_lossy().as_bytes())).collect::<Result<_, _>>()?;
        let ptrs: Vec<*const c_char> = bases.iter().map(|b| b.as_ptr()).collect();
        let mut handle = std::ptr::null_mut();
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_shards_open(ptrs.as_ptr(), ptrs.len(), backend, dimensions, METRIC_DEFAULT, devices.as_ptr(),
                                   std::ptr::null(), 0, &mut handle, err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok(Self { handle, dims: dimensions, fixed_ef })
    }
    pub fn search_batch(&self, queries: &[f32], top_k: usize, ef: usize) -> anyhow::Result<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        let nq = queries.len() / self.dims;
        let mut keys = vec![u64::MAX; nq * top_k];
        let mut dists = vec![f32::INFINITY; nq * top_k];
        let mut counts = vec![0u32; nq];
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_shards_search(self.handle, queries.as_ptr(), nq, top_k, ef, std::ptr::null(), keys.as_mut_ptr(),
                                     dists.as_mut_ptr(), counts.as_mut_ptr(), err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok((keys, dists, counts))
    }
}
impl BackendSearcher for ShardedCudaSearcher {
    fn search(&self, query: &[f32], top_k: usize, complexity: usize) -> anyhow::Result<(Vec<u64>, Vec<f32>)> {
        let (mut keys, mut dists, counts) = self.search_batch(query, top_k, self.fixed_ef.unwrap_or(complexity))?;
        keys.truncate(counts[0] as usize);
        dists.truncate(counts[0] as usize);
        Ok((keys, dists))
    }
    fn len(&self) -> usize {
        unsafe { leann_cuda_shards_len(self.handle) }
    }
}
impl Drop for ShardedCudaSearcher {
    fn drop(&mut self) {
        unsafe { leann_cuda_shards_close(self.handle) }
    }
}

/// `IndexSearcher` (src/index/searcher.rs:66-257) over the library: opens `<base>.passages.*`, `<base>.ids.txt` and
/// the backend; `search` is `search_with_options` without the passage fetch (ids + scores; the caller's PassageStore
/// resolves texts as before), `bm25_search` is searcher.rs:228-246.
pub struct CudaIndexSearcher(*mut LeannCudaSearcher);
unsafe impl Send for CudaIndexSearcher {}
unsafe impl Sync for CudaIndexSearcher {}

impl CudaIndexSearcher {
    pub fn load(index_path: &Path, backend_name: &str, dimensions: usize, device: i32) -> anyhow::Result<Self> {
        let base = CString::new(index_path.to_string_lossy().as_bytes())?;
        let name = CString::new(backend_name)?;
        let mut h = std::ptr::null_mut();
        let mut err = [0u8; 1024];
        let rc = unsafe { leann_cuda_searcher_load(base.as_ptr(), name.as_ptr(), dimensions, device, &mut h, err.as_mut_ptr() as *mut c_char, err.len()) };
        check(rc, &err)?;
        Ok(Self(h))
    }
    pub fn len(&self) -> usize {
        unsafe { leann_cuda_searcher_len(self.0) }
    }
    /// `id_map[idx]` or `idx.to_string()` (searcher.rs:180-184).
    pub fn passage_id(&self, idx: u64) -> String {
        let mut buf = vec![0u8; 256];
        let n = unsafe { leann_cuda_searcher_id(self.0, idx, buf.as_mut_ptr() as *mut c_char, buf.len()) };
        String::from_utf8_lossy(&buf[..n.min(buf.len())]).into_owned()
    }
    /// One query through `search_with_options` (searcher.rs:123-210): `filter` is the DSL string, `hybrid` = Some((text, alpha)).
    pub fn search(&self, query: &[f32], top_k: usize, complexity: usize, filter: Option<&str>, hybrid: Option<(&str, f32)>)
                  -> anyhow::Result<Vec<(String, f32)>> {
        let fexpr = filter.map(CString::new).transpose()?;
        let tptr = hybrid.map(|(t, _)| t.as_ptr() as *const c_char);
        let tlen = hybrid.map(|(t, _)| t.len());
        let mut idx = vec![u64::MAX; top_k];
        let mut scores = vec![0f32; top_k];
        let mut count = 0u32;
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_searcher_search(self.0, query.as_ptr(), tptr.as_ref().map_or(std::ptr::null(), |p| p as *const _),
                                       tlen.as_ref().map_or(std::ptr::null(), |l| l as *const _), 1, top_k, complexity,
                                       fexpr.as_ref().map_or(std::ptr::null(), |f| f.as_ptr()), hybrid.is_some() as c_int,
                                       hybrid.map_or(0.7, |(_, a)| a), idx.as_mut_ptr(), scores.as_mut_ptr(), &mut count,
                                       err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok((0..count as usize).map(|i| (self.passage_id(idx[i]), scores[i])).collect())
    }
    pub fn bm25_search(&self, query: &str, top_k: usize) -> anyhow::Result<Vec<(String, f32)>> {
        let mut idx = vec![u64::MAX; top_k];
        let mut scores = vec![0f32; top_k];
        let mut count = 0u32;
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_searcher_bm25_search(self.0, query.as_ptr() as *const c_char, query.len(), top_k, idx.as_mut_ptr(),
                                            scores.as_mut_ptr(), &mut count, err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok((0..count as usize).map(|i| (self.passage_id(idx[i]), scores[i])).collect())
    }
}
impl Drop for CudaIndexSearcher {
    fn drop(&mut self) {
        unsafe { leann_cuda_searcher_close(self.0) }
    }
}

/// BM25 index built once (the reference rebuilds it per query, searcher.rs:149-150).
pub struct CudaBm25(*mut LeannCudaBm25);
unsafe impl Send for CudaBm25 {}
unsafe impl Sync for CudaBm25 {}

impl CudaBm25 {
    pub fn build(documents: &[String], device: i32) -> anyhow::Result<Self> {
        let ptrs: Vec<*const c_char> = documents.iter().map(|d| d.as_ptr() as *const c_char).collect();
        let lens: Vec<usize> = documents.iter().map(|d| d.len()).collect();
        let mut h = std::ptr::null_mut();
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_bm25_build(ptrs.as_ptr(), lens.as_ptr(), documents.len(), device, &mut h,
                                  err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok(Self(h))
    }
    /// Terms the handle also keeps as dense score rows (LEANN_CUDA_BM25_DENSE_FRAC / _MAX at build time; results do not depend on it).
    pub fn dense_rows(&self) -> usize {
        unsafe { leann_cuda_bm25_dense_rows(self.0) }
    }
}
impl Drop for CudaBm25 {
    fn drop(&mut self) {
        unsafe { leann_cuda_bm25_free(self.0) }
    }
}

impl CudaSearcher {
    /// The arithmetic of `IndexSearcher::search_with_options` (searcher.rs:123-210) for a batch.
    /// `filter_mask`: bit i set = passage i loads and passes `MetadataFilter::matches`.
    #[allow(clippy::too_many_arguments)]
    pub fn hybrid_search(&self, bm25: Option<&CudaBm25>, queries: &[f32], texts: Option<&[String]>, top_k: usize,
                         complexity: usize, alpha: f32, filter_mask: Option<&[u64]>)
                         -> anyhow::Result<(Vec<u64>, Vec<f32>, Vec<u32>)> {
        let nq = queries.len() / self.dims;
        let tptr: Vec<*const c_char> = texts.map(|t| t.iter().map(|s| s.as_ptr() as *const c_char).collect()).unwrap_or_default();
        let tlen: Vec<usize> = texts.map(|t| t.iter().map(|s| s.len()).collect()).unwrap_or_default();
        let mut idx = vec![u64::MAX; nq * top_k];
        let mut scores = vec![0f32; nq * top_k];
        let mut counts = vec![0u32; nq];
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_hybrid_search(
                self.handle, bm25.map_or(std::ptr::null(), |b| b.0 as *const _), queries.as_ptr(),
                if texts.is_some() { tptr.as_ptr() } else { std::ptr::null() },
                if texts.is_some() { tlen.as_ptr() } else { std::ptr::null() },
                nq, top_k, self.fixed_ef.unwrap_or(complexity), texts.is_some() as c_int, alpha,
                filter_mask.map_or(std::ptr::null(), |m| m.as_ptr()), idx.as_mut_ptr(), scores.as_mut_ptr(),
                counts.as_mut_ptr(), err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok((idx, scores, counts))
    }
}

/// Typed columns of every passage's metadata, built once when the index is opened; `mask(filter)` evaluates a filter
/// string (`MetadataFilter::parse` grammar, filter.rs:52-134) into the N-bit mask `hybrid_search` takes, replacing the
/// per-candidate `passages.get` + `filter.matches` of searcher.rs:186-194. `Ok(None)` where `parse` returns `None`.
pub struct CudaMetadataColumns {
    handle: *mut LeannCudaMetacols,
    n: usize,
}
unsafe impl Send for CudaMetadataColumns {}
unsafe impl Sync for CudaMetadataColumns {}

impl CudaMetadataColumns {
    pub fn build(metadata_json: &[String]) -> anyhow::Result<Self> {
        let ptrs: Vec<*const c_char> = metadata_json.iter().map(|d| d.as_ptr() as *const c_char).collect();
        let lens: Vec<usize> = metadata_json.iter().map(|d| d.len()).collect();
        let mut h = std::ptr::null_mut();
        let mut err = [0u8; 1024];
        let rc = unsafe {
            leann_cuda_metacols_build(ptrs.as_ptr(), lens.as_ptr(), metadata_json.len(), &mut h, err.as_mut_ptr() as *mut c_char, err.len())
        };
        check(rc, &err)?;
        Ok(Self { handle: h, n: metadata_json.len() })
    }
    pub fn mask(&self, filter_str: &str) -> anyhow::Result<Option<Vec<u64>>> {
        const LEANN_ERR_PARSE: c_int = -8;
        let expr = CString::new(filter_str)?;
        let mut f = std::ptr::null_mut();
        let mut err = [0u8; 1024];
        let rc = unsafe { leann_cuda_filter_parse(expr.as_ptr(), &mut f, err.as_mut_ptr() as *mut c_char, err.len()) };
        if rc == LEANN_ERR_PARSE {
            return Ok(None);
        }
        check(rc, &err)?;
        let mut bits = vec![0u64; (self.n + 63) / 64];
        let rc = unsafe { leann_cuda_metacols_mask(self.handle, f, bits.as_mut_ptr(), err.as_mut_ptr() as *mut c_char, err.len()) };
        unsafe { leann_cuda_filter_free(f) };
        check(rc, &err)?;
        Ok(Some(bits))
    }
}
impl Drop for CudaMetadataColumns {
    fn drop(&mut self) {
        unsafe { leann_cuda_metacols_free(self.handle) }
    }
}

#[allow(dead_code)]
fn _unused(_: *const c_void) {}

fn flatten(embeddings: &[Vec<f32>], dimensions: usize) -> anyhow::Result<Vec<f32>> {
    let mut flat = Vec::with_capacity(embeddings.len() * dimensions);
    for e in embeddings {
        anyhow::ensure!(e.len() == dimensions, "Dimension mismatch: expected {}, got {}", dimensions, e.len());
        flat.extend_from_slice(e);
    }
    Ok(flat)
}

/// Drop-in for `hnsw::build_index` (`src/backend/hnsw.rs:96-139`): same arguments, same `<base>.index` output.
pub fn build_index(embeddings: &[Vec<f32>], _ids: &[String], index_path: &Path, dimensions: usize, graph_degree: usize,
                   complexity: usize) -> anyhow::Result<()> {
    let flat = flatten(embeddings, dimensions)?;
    let base = CString::new(index_path.to_string_lossy().as_bytes())?;
    let mut err = [0u8; 1024];
    let mut h = std::ptr::null_mut();
    unsafe {
        check(leann_cuda_hnsw_build(flat.as_ptr(), 0, embeddings.len(), dimensions, graph_degree, complexity, METRIC_DEFAULT, 1, 0,
                                    &mut h, err.as_mut_ptr() as *mut c_char, err.len()), &err)?;
        let rc = leann_cuda_save(h, base.as_ptr(), err.as_mut_ptr() as *mut c_char, err.len());
        leann_cuda_close(h);
        check(rc, &err)
    }
}

/// Drop-in for `hnsw::add_to_index` (`src/backend/hnsw.rs:142-191`): load, append with keys from `start_id`, save.
pub fn add_to_index(embeddings: &[Vec<f32>], index_path: &Path, dimensions: usize, start_id: usize) -> anyhow::Result<()> {
    let flat = flatten(embeddings, dimensions)?;
    let base = CString::new(index_path.to_string_lossy().as_bytes())?;
    let mut err = [0u8; 1024];
    let mut h = std::ptr::null_mut();
    unsafe {
        check(leann_cuda_open(base.as_ptr(), BACKEND_HNSW, dimensions, METRIC_DEFAULT, 0, &mut h,
                              err.as_mut_ptr() as *mut c_char, err.len()), &err)?;
        let mut rc = leann_cuda_hnsw_add(h, flat.as_ptr(), 0, embeddings.len(), start_id as u64, 64, 1,
                                         err.as_mut_ptr() as *mut c_char, err.len());
        if rc == 0 {
            rc = leann_cuda_save(h, base.as_ptr(), err.as_mut_ptr() as *mut c_char, err.len());
        }
        leann_cuda_close(h);
        check(rc, &err)
    }
}

/// Drop-in for `diskann::build_index` (`src/backend/diskann.rs:70-105`): Vamana graph built on the GPU with the reference's
/// parameters (max_degree = graph_degree, build beam = complexity, alpha = 1.2, DistDot), written as `<base>.diskann`.
pub fn build_diskann_index(embeddings: &[Vec<f32>], index_path: &Path, dimensions: usize, graph_degree: usize,
                           complexity: usize) -> anyhow::Result<()> {
    let flat = flatten(embeddings, dimensions)?;
    let base = CString::new(index_path.to_string_lossy().as_bytes())?;
    let mut err = [0u8; 1024];
    let mut h = std::ptr::null_mut();
    unsafe {
        check(leann_cuda_vamana_build(flat.as_ptr(), 0, embeddings.len(), dimensions, graph_degree, complexity, 1.2, METRIC_DEFAULT,
                                      1, 0, &mut h, err.as_mut_ptr() as *mut c_char, err.len()), &err)?;
        let rc = leann_cuda_save(h, base.as_ptr(), err.as_mut_ptr() as *mut c_char, err.len());
        leann_cuda_close(h);
        check(rc, &err)
    }
}
