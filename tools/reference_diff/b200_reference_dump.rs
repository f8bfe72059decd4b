//! Dumps the REFERENCE prover's proof of the synthetic workload this repository benchmarks and tests, so that
//! `tools/reference_diff/compare.py` can diff it against the oracle / the CUDA backend stream by stream.
//!
//! NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT (no Rust toolchain, un-vendored Plonky3): written against the
//! reference sources as read (file:line below); expect to fix an import path or two.
//!
//! Install (reference checkout = 0xMiden/miden-vm):
//!   cp b200_reference_dump.rs  benches/miden-bench/src/bin/b200_reference_dump.rs
//!   benches/miden-bench/Cargo.toml [dependencies]:  miden-air.workspace = true, miden-core.workspace = true,
//!                                                   serde_json = { workspace = true, features = ["std"] }
//!   cargo run --profile optimized -p miden-bench --bin b200_reference_dump -- 10 9 8 > ref_10_9_8.json
//!   python tools/reference_diff/compare.py --reference ref_10_9_8.json --log-heights 10 9 8          (this repository)
//!
//! What it reproduces, item by item:
//!   * configuration: `miden_air::config::poseidon2_config(pcs_params(), RELATION_DIGEST)` (air/src/config.rs:57-81,
//!     241-273) and `observe_protocol_params` on `config.challenger()` (:188-198) -- what `prove_stark` does
//!     (prover/src/lib.rs:317-355) and what `miden-vm_b200/workload.py::initial_challenger` restates;
//!   * AIRs: `DummyMidenAir::new(width, aux_width)` for (51,4), (22,3), (16,1) with all-zero aux traces, exactly
//!     like `miden-bench miden:H:51:4 miden:H:22:3 miden:H:16:1` (benches/miden-bench/src/lifted.rs:23-92);
//!   * statement: `Statement::new(multi_air, vec![], vec![])`, default `MultiAir::observe`
//!     (crates/lifted-air/src/air.rs:307-324);
//!   * traces: NOT miden-bench's `rand` stream (unreproducible outside Rust) but this repository's generator,
//!     `workload.py::synthetic_trace`: cell i of AIR a = splitmix64(i ^ (2025 ^ (a << 56))) reduced once mod p,
//!     column 0 zero (the DummyMidenAir constraint, testing/airs/miden.rs:49-56,101-123).
//! Output: serde_json of `StarkProofData` (log_trace_heights + transcript{fields, commitments}; proof.rs:55-63).

use miden_air::config::{RELATION_DIGEST, observe_protocol_params, pcs_params, poseidon2_config};
use miden_core::{Felt, field::QuadFelt};
use miden_lifted_stark::{
    ProverInstance, StarkConfig,
    air::{BaseAir, LiftedAir, LiftedAirBuilder, MultiAir, ProverStatement, Statement},
    testing::airs::miden::DummyMidenAir,
};
use p3_field::Field;
use p3_matrix::{Matrix, dense::RowMajorMatrix};

const P: u64 = 0xFFFF_FFFF_0000_0001;
const WIDTHS: [usize; 3] = [51, 22, 16];
const AUX_WIDTHS: [usize; 3] = [4, 3, 1];

fn splitmix64(x: u64) -> u64 {
    let x = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
    let mut z = x;
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
    z ^ (z >> 31)
}

fn synthetic_trace(air: u64, log_height: u8, width: usize) -> RowMajorMatrix<Felt> {
    let height = 1usize << log_height;
    let mut values = Vec::with_capacity(height * width);
    for i in 0..(height * width) as u64 {
        let v = splitmix64(i ^ (2025u64 ^ (air << 56)));
        let v = if v >= P { v - P } else { v };
        values.push(if (i as usize) % width == 0 { Felt::ZERO } else { Felt::new_unchecked(v) });
    }
    RowMajorMatrix::new(values, width)
}

/// DummyMidenAir with the all-zero aux trace miden-bench gives it (lifted.rs:65-79).
struct ZeroAuxMidenAir(DummyMidenAir);

impl BaseAir<Felt> for ZeroAuxMidenAir {
    fn width(&self) -> usize {
        BaseAir::<Felt>::width(&self.0)
    }
}

impl<EF: Field> LiftedAir<Felt, EF> for ZeroAuxMidenAir {
    fn num_randomness(&self) -> usize {
        LiftedAir::<Felt, EF>::num_randomness(&self.0)
    }
    fn aux_width(&self) -> usize {
        LiftedAir::<Felt, EF>::aux_width(&self.0)
    }
    fn num_aux_values(&self) -> usize {
        LiftedAir::<Felt, EF>::num_aux_values(&self.0)
    }
    fn build_aux_trace(
        &self,
        main: &RowMajorMatrix<Felt>,
        _air_inputs: &[Felt],
        _aux_inputs: &[Felt],
        _challenges: &[EF],
    ) -> (RowMajorMatrix<EF>, Vec<EF>) {
        let aux_width = LiftedAir::<Felt, EF>::aux_width(self);
        let num_aux_values = LiftedAir::<Felt, EF>::num_aux_values(self);
        (RowMajorMatrix::new(EF::zero_vec(main.height() * aux_width), aux_width), EF::zero_vec(num_aux_values))
    }
    fn eval<AB: LiftedAirBuilder<F = Felt>>(&self, builder: &mut AB) {
        LiftedAir::<Felt, EF>::eval(&self.0, builder)
    }
}

struct Airs(Vec<ZeroAuxMidenAir>);

impl MultiAir<Felt, QuadFelt> for Airs {
    type Air = ZeroAuxMidenAir;
    fn airs(&self) -> &[Self::Air] {
        &self.0
    }
}

fn main() {
    let log_heights: Vec<u8> = std::env::args().skip(1).map(|a| a.parse().expect("log height")).collect();
    assert!(!log_heights.is_empty() && log_heights.len() <= 3, "usage: b200_reference_dump LOG_H0 [LOG_H1 [LOG_H2]]");
    let k = log_heights.len();

    let config = poseidon2_config(pcs_params(), RELATION_DIGEST);
    let airs: Vec<ZeroAuxMidenAir> =
        (0..k).map(|i| ZeroAuxMidenAir(DummyMidenAir::new(WIDTHS[i], AUX_WIDTHS[i]))).collect();
    let traces: Vec<RowMajorMatrix<Felt>> =
        (0..k).map(|i| synthetic_trace(i as u64, log_heights[i], WIDTHS[i])).collect();

    let statement = Statement::new(Airs(airs), Vec::new(), Vec::new()).expect("statement");
    let prover_statement = ProverStatement::new(statement, traces).expect("prover statement");
    let instance = ProverInstance::new(&config, &prover_statement, None).expect("no preprocessed columns");

    let mut challenger = config.challenger();
    observe_protocol_params(&mut challenger);
    let output = instance.prove(challenger).expect("proving failed");

    println!("{}", serde_json::to_string(&output.proof).expect("serialise proof"));
}
