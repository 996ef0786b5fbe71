//! dump_witness <out.bin> <hex msg 0> <pre len 0> <hex msg 1> <pre len 1>
//!
//! Rebuilds the reference's TestCircuit (halo2-dynamic-sha256 src/lib.rs:393-494: two digest() calls in one Context,
//! MAX_BYTE_SIZE 128/128, NUM_ADVICE 3, NUM_LOOKUP_ADVICE 1, LOOKUP_BITS 16, k = 17), runs MockProver and writes every
//! advice column as canonical little-endian 32-byte field elements:
//!
//!   magic "H2SHADMP" | u32 n_columns | u32 n_rows | n_columns x n_rows x [u8; 32]   (unassigned cells = zero)
//!
//! Column order = allocation order in `configure`: gate advice [0, NUM_ADVICE), lookup advice, then SpreadConfig's
//! denses[0..2), spreads[0..2) (spread.rs:39-52).  Compare with `python tools/compare_rust_dump.py out.bin ...`.
//!
//! SOURCE ONLY: written against the PSE halo2 fork's `MockProver::advice()` accessor; it has never been compiled here.
use halo2_base::gates::range::{RangeConfig, RangeStrategy::Vertical};
use halo2_base::halo2_proofs::{
    circuit::{Layouter, SimpleFloorPlanner},
    dev::{CellValue, MockProver},
    halo2curves::bn256::Fr,
    plonk::{Circuit, ConstraintSystem, Error},
};
use halo2_base::halo2_proofs::halo2curves::group::ff::PrimeField;
use halo2_base::SKIP_FIRST_PASS;
use halo2_dynamic_sha256::Sha256DynamicConfig;
use std::io::Write;

#[derive(Clone)]
struct DumpCircuit {
    inputs: Vec<Vec<u8>>,
    pre: Vec<usize>,
}

impl Circuit<Fr> for DumpCircuit {
    type Config = Sha256DynamicConfig<Fr>;
    type FloorPlanner = SimpleFloorPlanner;
    fn without_witnesses(&self) -> Self {
        self.clone()
    }
    fn configure(meta: &mut ConstraintSystem<Fr>) -> Self::Config {
        let range = RangeConfig::configure(meta, Vertical, &[3], &[1], 1, 16, 0, 17);
        Sha256DynamicConfig::configure(meta, vec![128, 128], range, 8, 2, true)
    }
    fn synthesize(&self, config: Self::Config, mut layouter: impl Layouter<Fr>) -> Result<(), Error> {
        let mut sha256 = config.clone();
        let range = sha256.range().clone();
        sha256.range().load_lookup_table(&mut layouter)?;
        sha256.load(&mut layouter)?;
        let mut first_pass = SKIP_FIRST_PASS;
        layouter.assign_region(
            || "dump",
            |region| {
                if first_pass {
                    first_pass = false;
                    return Ok(());
                }
                let ctx = &mut sha256.new_context(region);
                for (m, p) in self.inputs.iter().zip(self.pre.iter()) {
                    sha256.digest(ctx, m, Some(*p))?;
                }
                range.finalize(ctx);
                Ok(())
            },
        )
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    assert!(args.len() == 6, "usage: dump_witness out.bin hex0 pre0 hex1 pre1");
    let circuit = DumpCircuit {
        inputs: vec![hex::decode(&args[2]).unwrap(), hex::decode(&args[4]).unwrap()],
        pre: vec![args[3].parse().unwrap(), args[5].parse().unwrap()],
    };
    let prover = MockProver::run(17, &circuit, vec![]).unwrap();
    assert_eq!(prover.verify(), Ok(()));
    let advice = prover.advice();
    let mut f = std::fs::File::create(&args[1]).unwrap();
    f.write_all(b"H2SHADMP").unwrap();
    f.write_all(&(advice.len() as u32).to_le_bytes()).unwrap();
    f.write_all(&(advice[0].len() as u32).to_le_bytes()).unwrap();
    for col in advice.iter() {
        for cell in col.iter() {
            let v = match cell {
                CellValue::Assigned(v) => *v,
                _ => Fr::zero(),
            };
            f.write_all(v.to_repr().as_ref()).unwrap();
        }
    }
}
