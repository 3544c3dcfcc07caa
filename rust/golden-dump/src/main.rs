//! Dump golden vectors of the batched polynomial commitment from the reference's own proof system.
//!
//! What the reference reaches only through `prove()` (wormhole/prover/src/lib.rs:171-175) is called here directly on
//! seeded inputs -- `PolynomialBatch::from_values`, `MerkleTree` caps / leaves / paths, `Challenger`,
//! `PolynomialBatch::prove_openings` -- and written as JSON in the schema `tests/test_reference_vectors.py` consumes.
//! Inputs follow SURVEY.md 8d: SplitMix64 stream seeded `0x51575a4b00000000 + tag`, rejection-sampled below p, filled
//! column-major -- the same generator as tests/conftest.py::splitmix_columns and bench.py.
//!
//!     cargo run --release -- ../../tests/golden/reference_vectors.json
//!
//! NEVER COMPILED (no Rust toolchain in the build image): API names are upstream plonky2's as the fork keeps them
//! (`plonky2::fri::oracle::PolynomialBatch`, `plonky2::hash::merkle_tree::MerkleTree`, `plonky2::iop::challenger::Challenger`);
//! expect to fix imports on first build.
use plonky2::field::extension::Extendable;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::polynomial::PolynomialValues;
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::fri::reduction_strategies::FriReductionStrategy;
use plonky2::fri::structure::{FriBatchInfo, FriInstanceInfo, FriOracleInfo, FriPolynomialInfo};
use plonky2::fri::FriConfig;
use plonky2::hash::hash_types::RichField;
use plonky2::hash::poseidon::PoseidonHash;
use plonky2::hash::poseidon2::Poseidon2Hash;
use plonky2::iop::challenger::Challenger;
use plonky2::plonk::config::{GenericConfig, Hasher, PoseidonGoldilocksConfig};
use plonky2::util::timing::TimingTree;
use serde_json::{json, Value};
use sha2::{Digest, Sha256};

type F = GoldilocksField;
type C = PoseidonGoldilocksConfig;
const D: usize = 2;
const P: u64 = 0xFFFF_FFFF_0000_0001;

struct SplitMix(u64);
impl SplitMix {
    fn next(&mut self) -> u64 {
        loop {
            self.0 = self.0.wrapping_add(0x9E37_79B9_7F4A_7C15);
            let mut z = self.0;
            z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
            z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
            z ^= z >> 31;
            if z < P {
                return z;
            }
        }
    }
}

fn columns(seed: u64, cols: usize, n: usize) -> Vec<Vec<F>> {
    let mut g = SplitMix(seed);
    (0..cols).map(|_| (0..n).map(|_| F::from_canonical_u64(g.next())).collect()).collect()
}

fn u(x: &F) -> u64 {
    x.to_canonical_u64()
}

fn sha_words<'a>(it: impl Iterator<Item = &'a F>) -> String {
    let mut h = Sha256::new();
    for x in it {
        h.update(u(x).to_le_bytes());
    }
    h.finalize().iter().map(|b| format!("{:02x}", b)).collect()
}

fn commit_case(tag: u64, log_n: usize, cols: usize, rate_bits: usize, cap_height: usize) -> (Value, PolynomialBatch<F, C, D>) {
    let seed = 0x5157_5a4b_0000_0000u64 + tag;
    let n = 1usize << log_n;
    let values: Vec<PolynomialValues<F>> = columns(seed, cols, n).into_iter().map(PolynomialValues::new).collect();
    let mut timing = TimingTree::default();
    let batch = PolynomialBatch::<F, C, D>::from_values(values, rate_bits, false, cap_height, &mut timing, None);
    let leaves = &batch.merkle_tree.leaves;
    let num_leaves = leaves.len();
    let sample: Vec<usize> = vec![0, 1, num_leaves / 3, num_leaves / 2 + 5, num_leaves - 1];
    let rows: Vec<Value> = sample
        .iter()
        .map(|&i| {
            let proof = batch.merkle_tree.prove(i);
            json!({
                "index": i,
                "row": leaves[i].iter().map(u).collect::<Vec<_>>(),
                "siblings": proof.siblings.iter().map(|h| h.elements.iter().map(u).collect::<Vec<_>>()).collect::<Vec<_>>(),
            })
        })
        .collect();
    let v = json!({
        "seed": seed, "log_n": log_n, "cols": cols, "rate_bits": rate_bits, "cap_height": cap_height,
        "cap": batch.merkle_tree.cap.0.iter().flat_map(|h| h.elements.iter().map(u)).collect::<Vec<_>>(),
        "coeffs_sha256": sha_words(batch.polynomials.iter().flat_map(|p| p.coeffs.iter())),
        "leaves_sha256": sha_words(leaves.iter().flat_map(|r| r.iter())),
        "rows": rows,
    });
    (v, batch)
}

fn fri_case() -> Value {
    // four oracles of the leaf circuit's column counts at a small degree; one opening batch at zeta
    let log_n = 8usize;
    let cols = [6usize, 10, 3, 2];
    let (rate_bits, cap_height) = (3usize, 4usize);
    let batches: Vec<PolynomialBatch<F, C, D>> = cols.iter().enumerate().map(|(k, &c)| commit_case(100 + k as u64, log_n, c, rate_bits, cap_height).1).collect();
    let mut challenger = Challenger::<F, <C as GenericConfig<D>>::Hasher>::new();
    for b in &batches {
        challenger.observe_cap(&b.merkle_tree.cap);
    }
    let zeta = challenger.get_extension_challenge::<D>();
    let instance = FriInstanceInfo {
        oracles: cols.iter().map(|&c| FriOracleInfo { num_polys: c, blinding: false }).collect(),
        batches: vec![FriBatchInfo {
            point: zeta,
            polynomials: cols.iter().enumerate().flat_map(|(o, &c)| (0..c).map(move |i| FriPolynomialInfo { oracle_index: o, polynomial_index: i })).collect(),
        }],
    };
    let openings: Vec<<F as Extendable<D>>::Extension> =
        batches.iter().flat_map(|b| b.polynomials.iter().map(|p| p.to_extension::<D>().eval(zeta)).collect::<Vec<_>>()).collect();
    challenger.observe_extension_elements(&openings);
    let fri_config = FriConfig {
        rate_bits, cap_height, proof_of_work_bits: 10,
        reduction_strategy: FriReductionStrategy::ConstantArityBits(4, 3),
        num_query_rounds: 8,
    };
    let fri_params = fri_config.fri_params(log_n, false);
    let mut timing = TimingTree::default();
    let refs: Vec<&PolynomialBatch<F, C, D>> = batches.iter().collect();
    let proof = PolynomialBatch::<F, C, D>::prove_openings(&instance, &refs, &mut challenger, &fri_params, &mut timing);
    use plonky2::field::extension::FieldExtension;
    let ext = |e: &<F as Extendable<D>>::Extension| -> Vec<u64> { e.to_basefield_array().iter().map(u).collect() };
    json!({
        "log_n": log_n, "cols": cols, "rate_bits": rate_bits, "cap_height": cap_height, "pow_bits": 10, "queries": 8,
        "arity_bits": 4, "final_poly_bits": 3, "seed_tags": [100, 101, 102, 103],
        "zeta": ext(&zeta),
        "openings": openings.iter().map(ext).collect::<Vec<_>>(),
        // upstream's witness search is parallel `find_any`: hand this value to qpzk_fri_prove as pow_witness_in
        "pow_witness": u(&proof.pow_witness),
        "commit_phase_caps": proof.commit_phase_merkle_caps.iter().map(|c| c.0.iter().flat_map(|h| h.elements.iter().map(u)).collect::<Vec<_>>()).collect::<Vec<_>>(),
        "final_poly": proof.final_poly.coeffs.iter().map(ext).collect::<Vec<_>>(),
        "first_query_initial_rows": proof.query_round_proofs[0].initial_trees_proof.evals_proofs.iter().map(|(row, _)| row.iter().map(u).collect::<Vec<_>>()).collect::<Vec<_>>(),
    })
}

fn main() -> anyhow::Result<()> {
    let out = std::env::args().nth(1).unwrap_or_else(|| "reference_vectors.json".to_string());
    let poseidon: Vec<Value> = [vec![0u64; 12], (0u64..12).collect(), vec![P - 1; 12]]
        .iter()
        .map(|inp| {
            let st: [F; 12] = core::array::from_fn(|i| F::from_canonical_u64(inp[i]));
            let perm = <F as plonky2::hash::poseidon::Poseidon>::poseidon(st);
            json!({"input": inp, "output": perm.iter().map(u).collect::<Vec<_>>()})
        })
        .collect();
    let msg: Vec<F> = (1u64..20).map(F::from_canonical_u64).collect();
    let commits: Vec<Value> = [(5usize, 7usize, 3usize, 2usize), (8, 135, 3, 4), (13, 20, 3, 4), (16, 9, 3, 4)]
        .iter()
        .map(|&(log_n, cols, rb, h)| commit_case(log_n as u64, log_n, cols, rb, h).0)
        .collect();
    let v = json!({
        "source": "qp-plonky2 =1.5.5 (rust/golden-dump)",
        "poseidon_kats": poseidon,
        "poseidon_hash_no_pad_1_to_19": PoseidonHash::hash_no_pad(&msg).elements.iter().map(u).collect::<Vec<_>>(),
        "poseidon2_hash_no_pad_1_to_19": Poseidon2Hash::hash_no_pad(&msg).elements.iter().map(u).collect::<Vec<_>>(),
        "commits": commits,
        "fri": fri_case(),
    });
    std::fs::write(&out, serde_json::to_string_pretty(&v)?)?;
    println!("wrote {}", out);
    Ok(())
}
