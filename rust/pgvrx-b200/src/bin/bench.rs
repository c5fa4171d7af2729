//! pgvrx-bench -- the standalone bench binary of the north star, over the safe wrapper crate.
//!
//!   cargo run --release --bin pgvrx-bench -- --n 1000000 --dim 128 --m 16 --efc 64 --ef 40 --k 10 --nq 10000 \
//!       --steps 50 --warmup 5 --gpus 1
//!
//! Same program as tools/pgvrx_bench.cpp (which is what this repository compiles and runs in its tests, there being no
//! Rust toolchain in its image): CREATE INDEX on GPU 0 through `Index::build`, replicas by export -> load, one scanning
//! host thread per GPU, recall@k against the exact kNN, one JSON line on stdout.  No CPU fallback: without a device every
//! call fails and the program exits 3.
use pgvrx_b200::{device_count, level_from_uniform, Elem, Index, Metric};
use std::sync::{Arc, Barrier};
use std::time::Instant;

struct Args {
    n: usize,
    nq: usize,
    dim: i32,
    m: i32,
    efc: i32,
    ef: i32,
    k: i32,
    steps: usize,
    warmup: usize,
    gpus: usize,
    batches: usize,
    recall_queries: usize,
}

fn parse() -> Option<Args> {
    let mut a = Args { n: 100_000, nq: 10_000, dim: 128, m: 16, efc: 64, ef: 40, k: 10, steps: 20, warmup: 3, gpus: 1, batches: 4, recall_queries: 1000 };
    let mut it = std::env::args().skip(1);
    while let Some(f) = it.next() {
        let v = it.next()?;
        match f.as_str() {
            "--n" => a.n = v.parse().ok()?,
            "--nq" => a.nq = v.parse().ok()?,
            "--dim" => a.dim = v.parse().ok()?,
            "--m" => a.m = v.parse().ok()?,
            "--efc" => a.efc = v.parse().ok()?,
            "--ef" => a.ef = v.parse().ok()?,
            "--k" => a.k = v.parse().ok()?,
            "--steps" => a.steps = v.parse().ok()?,
            "--warmup" => a.warmup = v.parse().ok()?,
            "--gpus" => a.gpus = v.parse().ok()?,
            "--batches" => a.batches = v.parse().ok()?,
            "--recall-queries" => a.recall_queries = v.parse().ok()?,
            _ => return None,
        }
    }
    Some(a)
}

// counter-based generator: value (stream, i) does not depend on how the work is split
fn mix64(mut x: u64) -> u64 {
    x = x.wrapping_add(0x9e37_79b9_7f4a_7c15);
    x = (x ^ (x >> 30)).wrapping_mul(0xbf58_476d_1ce4_e5b9);
    x = (x ^ (x >> 27)).wrapping_mul(0x94d0_49bb_1331_11eb);
    x ^ (x >> 31)
}
fn uniform(stream: u64, i: u64) -> f64 {
    ((mix64(stream.wrapping_mul(0x1000_0000_01b3) ^ i) >> 11) as f64 + 0.5) / (1u64 << 53) as f64
}
fn gaussian_rows(n: usize, dim: usize, stream: u64) -> Vec<u8> {
    let mut out = Vec::with_capacity(n * dim * 4);
    for i in 0..(n * dim / 2 + 1) {
        // Box-Muller, two values per pair of draws
        let (u1, u2) = (uniform(stream, 2 * i as u64), uniform(stream, 2 * i as u64 + 1));
        let r = (-2.0 * u1.ln()).sqrt();
        for v in [r * (2.0 * std::f64::consts::PI * u2).cos(), r * (2.0 * std::f64::consts::PI * u2).sin()] {
            if out.len() < n * dim * 4 {
                out.extend_from_slice(&(v as f32).to_le_bytes());
            }
        }
    }
    out
}

fn main() {
    let a = match parse() {
        Some(a) => a,
        None => {
            eprintln!("usage: pgvrx-bench [--n N] [--dim D] [--m M] [--efc E] [--ef EF] [--k K] [--nq Q] [--steps S] [--warmup W] [--gpus G] [--batches B] [--recall-queries R]");
            std::process::exit(2);
        }
    };
    if device_count() < 1 {
        eprintln!("pgvrx-bench: no CUDA device (the library has no CPU path)");
        std::process::exit(3);
    }
    let die = |e: pgvrx_b200::Error| -> ! {
        eprintln!("pgvrx-bench: {}", e);
        std::process::exit(3)
    };

    // ---- CREATE INDEX on GPU 0 ----
    let rows = gaussian_rows(a.n, a.dim as usize, 2);
    let levels: Vec<i32> = (0..a.n).map(|i| level_from_uniform(uniform(4, i as u64), a.m)).collect();
    let mut ix0 = Index::create(Elem::F32, Metric::L2Squared, a.dim, a.m, a.efc, 0, a.n as i64).unwrap_or_else(|e| die(e));
    let t0 = Instant::now();
    let bst = ix0.build(&rows, &levels, None).unwrap_or_else(|e| die(e));
    let build_s = t0.elapsed().as_secs_f64();

    // ---- replicas: the build does not shard (insert i depends on every insert before it) ----
    let g = ix0.export_graph().unwrap_or_else(|e| die(e));
    let mut replicas = vec![ix0];
    for d in 1..a.gpus {
        let mut ix = Index::create(Elem::F32, Metric::L2Squared, a.dim, a.m, a.efc, d as i32, a.n as i64).unwrap_or_else(|e| die(e));
        ix.load_graph(&g).unwrap_or_else(|e| die(e));
        replicas.push(ix);
    }

    // ---- index scans: one host thread per GPU, blocking host-buffer calls (the e2e figure) ----
    let barrier = Arc::new(Barrier::new(a.gpus + 1));
    let mut handles = Vec::new();
    for (d, mut ix) in replicas.into_iter().enumerate() {
        let b = barrier.clone();
        let (nq, dim, ef, k, steps, warmup, batches) = (a.nq, a.dim as usize, a.ef, a.k, a.steps, a.warmup, a.batches);
        handles.push(std::thread::spawn(move || {
            let qs: Vec<Vec<u8>> = (0..batches).map(|i| gaussian_rows(nq, dim, 1000 + 100 * d as u64 + i as u64)).collect();
            let mut kernel_ms = 0f64;
            let mut last = None;
            for i in 0..warmup {
                ix.search(&qs[i % batches], nq, ef, k).expect("warm-up scan");
            }
            b.wait();
            for i in 0..steps {
                let (r, st) = ix.search(&qs[i % batches], nq, ef, k).expect("scan");
                kernel_ms += st.kernel_ms as f64;
                last = Some((r, (steps - 1) % batches));
            }
            b.wait();
            (ix, qs, last, kernel_ms)
        }));
    }
    barrier.wait();
    let t1 = Instant::now();
    barrier.wait();
    let scan_s = t1.elapsed().as_secs_f64();
    let mut results: Vec<_> = handles.into_iter().map(|h| h.join().expect("scan thread")).collect();

    // ---- recall@k of GPU 0's last batch against the exact kNN ----
    let (ix0, qs0, last0, kernel_ms0) = &mut results[0];
    let (res, bi) = last0.as_ref().expect("at least one step");
    let nr = a.recall_queries.min(a.nq);
    let (gt, _, _) = ix0.flat_topk(&qs0[*bi], nr, a.k).unwrap_or_else(|e| die(e));
    let ku = a.k as usize;
    let mut hit = 0usize;
    for q in 0..nr {
        let (ids, _) = res.row(q);
        hit += ids.iter().filter(|id| gt[q * ku..(q + 1) * ku].contains(id)).count();
    }
    println!(
        "{{\"metric\": \"QPS@recall@{k}\", \"value\": {qps:.1}, \"unit\": \"queries/s\", \"n_gpus\": {g}, \"steps\": {s}, \"warmup\": {w}, \
         \"recall\": {rec:.5}, \"kernel_qps_gpu0\": {kq:.1}, \"build_vectors_per_s\": {bv:.1}, \"build_rounds\": {br}, \
         \"config\": {{\"n\": {n}, \"dim\": {d}, \"m\": {m}, \"ef_construction\": {efc}, \"ef_search\": {ef}, \"nq\": {nq}}}, \"host\": \"rust\"}}",
        k = a.k,
        qps = (a.gpus * a.nq * a.steps) as f64 / scan_s,
        g = a.gpus,
        s = a.steps,
        w = a.warmup,
        rec = hit as f64 / (nr * ku) as f64,
        kq = (a.nq * a.steps) as f64 / (*kernel_ms0 / 1e3),
        bv = a.n as f64 / build_s,
        br = bst.n_rounds,
        n = a.n,
        d = a.dim,
        m = a.m,
        efc = a.efc,
        ef = a.ef,
        nq = a.nq
    );
}
