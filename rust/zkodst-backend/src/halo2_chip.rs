//! `Table16Chip::configure` over halo2's own `ConstraintSystem`, and a `Circuit` whose `synthesize` assigns
//! the cells the CUDA library computes.
//!
//! Mirrors (paths in the reference repository):
//!   * `Table16Chip::configure(meta: &mut ConstraintSystem<pallas::Base>) -> Table16Config`
//!                                       blake2f-circuit/src/blake2f/table16.rs:277-327
//!   * `SpreadTableChip::configure`     blake2f-circuit/src/blake2f/table16/spread_table.rs:425-467
//!   * `CompressionConfig::configure`   blake2f-circuit/src/blake2f/table16/compression.rs:555-1074
//!   * `Circuit::{without_witnesses, configure, synthesize}`  blake2f-circuit/src/blake2f.rs:257-277
//! for the circuit frozen in docs/CIRCUIT.md: 12 advice columns in the reference's allocation order, the
//! selector-less (tag, dense, spread) lookup, equality on a_1..a_8, the reference's 12 selectors plus
//! `s_const` / `s_fmask`, one constants fixed column, and the 26 gate polynomials in declaration order.
//!
//! Every expression below is built with the same operators, in the same order, as `blake2f_configure` of the
//! C++ oracle (oracle/blake2f_circuit.hpp) and the string builder of the library
//! (zk-odst_b200/csrc/vk_repr.cpp): `format!("{:?}", vk.pinned())` of this circuit is what
//! `zk_vk_pinned_debug` is meant to reproduce, and rust/xcheck compares the two character by character.
//! The order of `query_advice` calls fixes `advice_queries` (docs/CIRCUIT.md "Queries"); do not reorder `let`s.
//!
//! `synthesize` is data-driven: the region's copy constraints, selector activations and constants come from
//! `zk_blake2f_layout_tables` (the tables the library's keygen uses), the cell values from
//! `zk_blake2f_witness_batch`.  One `assign_region` per compression, regions in order, as in the reference's
//! `Table16Chip::compress` (table16.rs:361-373).
//!
//! NOT COMPILED in the build image (no Rust toolchain): shipped as source, see INTEGRATION.md.
use halo2_proofs::{
    circuit::{AssignedCell, Layouter, SimpleFloorPlanner, Value},
    pasta::{group::ff::{Field, PrimeField}, pallas},
    plonk::{Advice, Circuit, Column, ConstraintSystem, Error, Expression, Fixed, Selector, TableColumn},
    poly::Rotation,
};

use crate::{Blake2fWitness, GpuProver};
use zkodst_sys as sys;

type F = pallas::Base;

/// a-number (the reference's gate comments a_0..a_9) -> halo2 advice column index (table16.rs:281-310).
pub const A_COLUMN: [usize; 10] = [7, 8, 9, 1, 2, 0, 3, 4, 5, 6];
/// Selector declaration order: the reference's (compression.rs:561-577), then the two pinned-input selectors.
pub const SEL_NAMES: [&str; 14] = [
    "s_spread_a1", "s_spread_b1", "s_spread_c1", "s_spread_d1", "s_spread_a2", "s_spread_b2", "s_spread_c2",
    "s_spread_d2", "s_decompose_abcd", "s_decompose_efgh", "s_decompose_ijkl", "s_digest", "s_const", "s_fmask",
];
const A1: usize = 0; const B1: usize = 1; const C1: usize = 2; const D1: usize = 3;
const A2: usize = 4; const B2: usize = 5; const C2: usize = 6; const D2: usize = 7;
const ABCD: usize = 8; const EFGH: usize = 9; const IJKL: usize = 10; const DIGEST: usize = 11;
const CONST: usize = 12; const FMASK: usize = 13;

#[derive(Clone, Debug)]
pub struct Table16Config {
    /// by halo2 advice column index 0..11 (10 and 11 are allocated by the spread table and never assigned)
    pub advice: [Column<Advice>; 12],
    pub table: [TableColumn; 3], // tag, dense, spread
    pub constants: Column<Fixed>,
    pub selectors: [Selector; 14],
}

fn pow2(e: u32) -> F {
    F::from_u128(1u128 << e)
}

#[derive(Clone, Debug)]
pub struct Table16Chip {
    config: Table16Config,
}

impl Table16Chip {
    pub fn construct(config: Table16Config) -> Self {
        Table16Chip { config }
    }
    pub fn config(&self) -> &Table16Config {
        &self.config
    }

    /// table16.rs:277-327
    pub fn configure(meta: &mut ConstraintSystem<F>) -> Table16Config {
        // Columns required by this chip: message_schedule (idx 0 = a_5), six extras (idx 1..6 =
        // a_3, a_4, a_6, a_7, a_8, a_9), then the lookup inputs (idx 7..9 = a_0, a_1, a_2).
        let message_schedule = meta.advice_column();
        let extras = [
            meta.advice_column(), meta.advice_column(), meta.advice_column(),
            meta.advice_column(), meta.advice_column(), meta.advice_column(),
        ];
        let input_tag = meta.advice_column();
        let input_dense = meta.advice_column();
        let input_spread = meta.advice_column();
        // SpreadTableChip::configure (spread_table.rs:425-467)
        let table_tag = meta.lookup_table_column();
        let table_dense = meta.lookup_table_column();
        let table_spread = meta.lookup_table_column();
        let unused_a = meta.advice_column(); // idx 10 (spread_table.rs:435-441)
        let unused_b = meta.advice_column(); // idx 11
        meta.lookup(|meta| {
            let tag_cur = meta.query_advice(input_tag, Rotation::cur());
            let dense_cur = meta.query_advice(input_dense, Rotation::cur());
            let spread_cur = meta.query_advice(input_spread, Rotation::cur());
            vec![(tag_cur, table_tag), (dense_cur, table_dense), (spread_cur, table_spread)]
        });
        let a: [Column<Advice>; 10] = [
            input_tag, input_dense, input_spread, extras[0], extras[1], message_schedule, extras[2], extras[3],
            extras[4], extras[5],
        ];
        for col in a.iter().take(9).skip(1) {
            meta.enable_equality(*col); // a_1..a_8 (table16.rs:312-314)
        }
        let sel: [Selector; 14] = [
            meta.selector(), meta.selector(), meta.selector(), meta.selector(), meta.selector(), meta.selector(),
            meta.selector(), meta.selector(), meta.selector(), meta.selector(), meta.selector(), meta.selector(),
            meta.selector(), meta.selector(),
        ];
        let constants = meta.fixed_column(); // fixed column 3

        let prev = Rotation::prev();
        let cur = Rotation::cur();
        let next = Rotation::next();
        let one = || Expression::Constant(F::ONE);

        meta.create_gate("decompose ABCD", |meta| {
            let s = meta.query_selector(sel[ABCD]);
            let word = meta.query_advice(a[3], cur);
            let l0 = meta.query_advice(a[1], prev);
            let l1 = meta.query_advice(a[1], cur);
            let l2 = meta.query_advice(a[1], next);
            let l3 = meta.query_advice(a[4], cur);
            vec![("dense", s * (word - l0 - l1 * pow2(16) - l2 * pow2(32) - l3 * pow2(48)))]
        });
        meta.create_gate("Decompose EFGH", |meta| {
            let s = meta.query_selector(sel[EFGH]);
            let tag_lo = meta.query_advice(a[0], cur);
            let tag_hi = meta.query_advice(a[0], next);
            let r_dense = meta.query_advice(a[3], cur);
            let p_hi_d = meta.query_advice(a[1], next);
            let p_lo_d = meta.query_advice(a[1], cur);
            let r_spread = meta.query_advice(a[4], cur);
            let p_hi_s = meta.query_advice(a[2], next);
            let p_lo_s = meta.query_advice(a[2], cur);
            vec![
                ("tag_p0", s.clone() * tag_lo),
                ("tag_p4", s.clone() * tag_hi),
                ("dense", s.clone() * (r_dense - p_hi_d - p_lo_d * pow2(8))),
                ("spread", s * (r_spread - p_hi_s - p_lo_s * pow2(16))),
            ]
        });
        meta.create_gate("Decompose IJKL", |meta| {
            let s = meta.query_selector(sel[IJKL]);
            let tag = meta.query_advice(a[0], cur);
            let bit = meta.query_advice(a[5], cur);
            let r_dense = meta.query_advice(a[3], cur);
            let q_d = meta.query_advice(a[1], cur);
            let r_spread = meta.query_advice(a[4], cur);
            let q_s = meta.query_advice(a[2], cur);
            vec![
                ("tag_q0", s.clone() * (tag.clone() * (tag - one()))),
                ("bit", s.clone() * (bit.clone() * (bit.clone() - one()))),
                ("dense", s.clone() * (r_dense - bit.clone() - q_d * pow2(1))),
                ("spread", s * (r_spread - bit - q_s * pow2(2))),
            ]
        });

        // The 8 (or 12) window inputs: X0..X3 = a3..a6[prev], Y0, Y1 = a7, a8[prev], Y2, Y3 = a3, a4[cur]
        // (then a5..a8[cur] for the third addend), accumulated with limb weights 2^(w i).
        let window = |meta: &mut halo2_proofs::plonk::VirtualCells<'_, F>, w: u32, three: bool| -> Expression<F> {
            let x0 = meta.query_advice(a[3], prev);
            let x1 = meta.query_advice(a[4], prev);
            let x2 = meta.query_advice(a[5], prev);
            let x3 = meta.query_advice(a[6], prev);
            let y0 = meta.query_advice(a[7], prev);
            let y1 = meta.query_advice(a[8], prev);
            let y2 = meta.query_advice(a[3], cur);
            let y3 = meta.query_advice(a[4], cur);
            let xs = [x0, x1, x2, x3];
            let ys = [y0, y1, y2, y3];
            if three {
                let z0 = meta.query_advice(a[5], cur);
                let z1 = meta.query_advice(a[6], cur);
                let z2 = meta.query_advice(a[7], cur);
                let z3 = meta.query_advice(a[8], cur);
                let zs = [z0, z1, z2, z3];
                let mut acc = xs[0].clone() + ys[0].clone() + zs[0].clone();
                for i in 1..4 {
                    acc = acc + (xs[i].clone() + ys[i].clone() + zs[i].clone()) * pow2(w * i as u32);
                }
                acc
            } else {
                let mut acc = xs[0].clone() + ys[0].clone();
                for i in 1..4 {
                    acc = acc + (xs[i].clone() + ys[i].clone()) * pow2(w * i as u32);
                }
                acc
            }
        };
        let add_gate = |meta: &mut ConstraintSystem<F>, name: &'static str, s_idx: usize, three: bool| {
            meta.create_gate(name, |meta| {
                let s = meta.query_selector(sel[s_idx]);
                if three {
                    // the oracle queries `sum = A(3, PREV)` first; it is already the first window input
                    let _ = meta.query_advice(a[3], prev);
                }
                let acc = window(meta, 16, three);
                let z0 = meta.query_advice(a[1], prev);
                let z1 = meta.query_advice(a[1], cur);
                let z2 = meta.query_advice(a[1], next);
                let z3 = meta.query_advice(a[3], next);
                let carry = meta.query_advice(a[9], cur);
                let lin = acc - z0 - z1 * pow2(16) - z2 * pow2(32) - z3 * pow2(48) - carry.clone() * pow2(64);
                let rng = if three {
                    carry.clone() * (carry.clone() - one()) * (carry - Expression::Constant(F::from(2)))
                } else {
                    carry.clone() * (carry - one())
                };
                vec![("sum", s.clone() * lin), ("carry", s * rng)]
            });
        };
        let xor_limb_gate = |meta: &mut ConstraintSystem<F>, name: &'static str, s_idx: usize| {
            meta.create_gate(name, |meta| {
                let s = meta.query_selector(sel[s_idx]);
                let x = meta.query_advice(a[3], cur);
                let y = meta.query_advice(a[4], cur);
                let even = meta.query_advice(a[2], cur);
                let odd = meta.query_advice(a[2], next);
                vec![("xor", s * (x + y - even - odd * pow2(1)))]
            });
        };
        let xor_word_gate = |meta: &mut ConstraintSystem<F>, name: &'static str, s_idx: usize, offs: [u32; 5]| {
            meta.create_gate(name, |meta| {
                let s = meta.query_selector(sel[s_idx]);
                let acc = window(meta, 32, false);
                let p0 = meta.query_advice(a[5], cur);
                let p1 = meta.query_advice(a[6], cur);
                let p2 = meta.query_advice(a[7], cur);
                let p3 = meta.query_advice(a[8], cur);
                let p4 = meta.query_advice(a[3], next);
                let ps = [p0, p1, p2, p3, p4];
                let mut even = ps[0].clone() * pow2(2 * offs[0]);
                for i in 1..5 {
                    even = even + ps[i].clone() * pow2(2 * offs[i]);
                }
                let o0 = meta.query_advice(a[2], prev);
                let o1 = meta.query_advice(a[2], cur);
                let o2 = meta.query_advice(a[2], next);
                let o3 = meta.query_advice(a[4], next);
                let odd = o0 + o1 * pow2(32) + o2 * pow2(64) + o3 * pow2(96);
                vec![("xor", s * (acc - even - odd * pow2(1)))]
            });
        };
        add_gate(meta, "s_spread_a1", A1, true);
        xor_limb_gate(meta, "s_spread_d1", D1);
        add_gate(meta, "s_spread_c1", C1, false);
        xor_word_gate(meta, "s_spread_b1", B1, [0, 8, 24, 40, 56]);
        add_gate(meta, "s_spread_a2", A2, true);
        xor_limb_gate(meta, "s_spread_d2", D2);
        add_gate(meta, "s_spread_c2", C2, false);
        xor_word_gate(meta, "s_spread_b2", B2, [0, 15, 31, 47, 63]);
        meta.create_gate("s_digest", |meta| {
            let s = meta.query_selector(sel[DIGEST]);
            let acc = window(meta, 32, false);
            let e0 = meta.query_advice(a[2], prev);
            let e1 = meta.query_advice(a[2], cur);
            let e2 = meta.query_advice(a[2], next);
            let e3 = meta.query_advice(a[5], cur);
            let even = e0 + e1 * pow2(32) + e2 * pow2(64) + e3 * pow2(96);
            let o0 = meta.query_advice(a[6], cur);
            let o1 = meta.query_advice(a[7], cur);
            let o2 = meta.query_advice(a[8], cur);
            let o3 = meta.query_advice(a[3], next);
            let odd = o0 + o1 * pow2(32) + o2 * pow2(64) + o3 * pow2(96);
            let d0 = meta.query_advice(a[1], prev);
            let d1 = meta.query_advice(a[1], cur);
            let d2 = meta.query_advice(a[1], next);
            let d3 = meta.query_advice(a[4], next);
            let out = meta.query_advice(a[5], next);
            vec![
                ("xor", s.clone() * (acc - even - odd * pow2(1))),
                ("word", s * (out - d0 - d1 * pow2(16) - d2 * pow2(32) - d3 * pow2(48))),
            ]
        });
        // Pinned inputs (docs/CIRCUIT.md): the reference witnesses the IV freely (subregion_initial.rs:11-52).
        meta.create_gate("pin constant", |meta| {
            let s = meta.query_selector(sel[CONST]);
            let word = meta.query_advice(a[3], cur);
            let c = meta.query_fixed(constants, cur);
            vec![("word", s * (word - c))]
        });
        meta.create_gate("final flag", |meta| {
            let s = meta.query_selector(sel[FMASK]);
            let word = meta.query_advice(a[3], cur);
            let bit = meta.query_advice(a[9], cur);
            vec![
                ("mask", s.clone() * (word - bit.clone() * (pow2(64) - F::ONE))),
                ("bit", s * (bit.clone() * (bit - one()))),
            ]
        });

        let mut advice = [message_schedule; 12];
        for (an, col) in a.iter().enumerate() {
            advice[A_COLUMN[an]] = *col;
        }
        advice[10] = unused_a;
        advice[11] = unused_b;
        Table16Config { advice, table: [table_tag, table_dense, table_spread], constants, selectors: sel }
    }

    /// `SpreadTableChip::load` (spread_table.rs:470-508): row i = (tag(i), i, spread(i)) for i < 2^16.
    pub fn load(config: Table16Config, layouter: &mut impl Layouter<F>) -> Result<(), Error> {
        layouter.assign_table(
            || "spread table",
            |mut table| {
                for i in 0..(1usize << 16) {
                    let dense = i as u64;
                    let tag: u64 = if dense < (1 << 8) { 0 } else if dense < (1 << 15) { 1 } else { 2 };
                    let mut spread = 0u64;
                    for b in 0..16 {
                        spread |= ((dense >> b) & 1) << (2 * b);
                    }
                    table.assign_cell(|| "tag", config.table[0], i, || Value::known(F::from(tag)))?;
                    table.assign_cell(|| "dense", config.table[1], i, || Value::known(F::from(dense)))?;
                    table.assign_cell(|| "spread", config.table[2], i, || Value::known(F::from(spread)))?;
                }
                Ok(())
            },
        )
    }
}

/// The layout tables of one region (zk_blake2f_layout_tables).
pub struct RegionTables {
    pub rows: usize,
    pub copies: Vec<[u32; 4]>,
    pub selectors: Vec<u8>, // [14][rows]
    pub constants: Vec<u64>,
    pub chain_rows: [u32; 16],
}

impl RegionTables {
    pub fn new(rounds: u32) -> Self {
        let mut rows = 0u64;
        unsafe { sys::zk_blake2f_rows_per_compression(rounds, &mut rows) };
        let rows = rows as usize;
        let mut n = 0u64;
        unsafe {
            sys::zk_blake2f_layout_tables(rounds, std::ptr::null_mut(), &mut n, std::ptr::null_mut(),
                                          std::ptr::null_mut(), std::ptr::null_mut())
        };
        let mut flat = vec![0u32; 4 * n as usize];
        let mut selectors = vec![0u8; sys::ZK_NUM_SELECTORS * rows];
        let mut constants = vec![0u64; rows];
        let mut chain_rows = [0u32; 16];
        let rc = unsafe {
            sys::zk_blake2f_layout_tables(rounds, flat.as_mut_ptr(), &mut n, selectors.as_mut_ptr(),
                                          constants.as_mut_ptr(), chain_rows.as_mut_ptr())
        };
        assert_eq!(rc, sys::ZK_OK);
        let copies = flat.chunks(4).map(|c| [c[0], c[1], c[2], c[3]]).collect();
        RegionTables { rows, copies, selectors, constants, chain_rows }
    }
}

/// The batch circuit: `n` compressions of `rounds` rounds, region j on rows [j R, (j + 1) R).
/// `advice` (12 x 2^k Montgomery cells from `GpuProver::witness`) is `None` for key generation
/// (`without_witnesses`), where every cell is `Value::unknown()`.
#[derive(Clone)]
pub struct Blake2fCircuit {
    pub k: u32,
    pub rounds: u32,
    pub n_compressions: usize,
    /// chain[j]: compression j continues compression j - 1 (zk_blake2f_keygen_chained)
    pub chain: Vec<bool>,
    pub advice: Option<std::sync::Arc<Vec<[u64; 4]>>>,
}

impl Blake2fCircuit {
    pub fn new(gpu: &mut GpuProver, k: u32, rounds: u32, inputs: &[Blake2fWitness], chain: Vec<bool>)
               -> Result<Self, crate::Error> {
        let mut advice = vec![[0u64; 4]; sys::ZK_NUM_ADVICE << k];
        gpu.witness(k, rounds, inputs, &mut advice)?;
        Ok(Blake2fCircuit { k, rounds, n_compressions: inputs.len(), chain, advice: Some(std::sync::Arc::new(advice)) })
    }
}

/// Montgomery limbs (the library's cell format = pasta_curves' in-memory form) -> field element.
fn from_montgomery(limbs: [u64; 4], r_inv: F) -> F {
    // from_raw interprets the limbs as a canonical integer; the limbs hold value * R mod p
    F::from_raw(limbs) * r_inv
}

impl Circuit<F> for Blake2fCircuit {
    type Config = Table16Config;
    type FloorPlanner = SimpleFloorPlanner;

    fn without_witnesses(&self) -> Self {
        Blake2fCircuit { advice: None, ..self.clone() }
    }

    fn configure(meta: &mut ConstraintSystem<F>) -> Self::Config {
        Table16Chip::configure(meta)
    }

    fn synthesize(&self, config: Self::Config, mut layouter: impl Layouter<F>) -> Result<(), Error> {
        Table16Chip::load(config.clone(), &mut layouter)?;
        let t = RegionTables::new(self.rounds);
        let n = 1usize << self.k;
        // R^-1 where R = 2^256 mod p: from_raw(R mod p) is the element R
        let r_limbs: [u64; 4] = [0x34786d38fffffffd, 0x992c350be41914ad, 0xffffffffffffffff, 0x3fffffffffffffff];
        let r_inv = F::from_raw(r_limbs).invert().unwrap();
        // the output word cells of the previous region, for chaining
        let mut prev_out: Vec<AssignedCell<F, F>> = Vec::new();
        for j in 0..self.n_compressions {
            let base = j * t.rows;
            let chained = j > 0 && self.chain.get(j).copied().unwrap_or(false);
            let outs = layouter.assign_region(
                || format!("compression {}", j),
                |mut region| {
                    // every cell of the ten used columns, by halo2 column index
                    let mut cells: Vec<Vec<AssignedCell<F, F>>> = Vec::with_capacity(10);
                    for col in 0..10 {
                        let mut v = Vec::with_capacity(t.rows);
                        for row in 0..t.rows {
                            let value = match &self.advice {
                                Some(a) => Value::known(from_montgomery(a[col * n + base + row], r_inv)),
                                None => Value::unknown(),
                            };
                            v.push(region.assign_advice(|| "cell", config.advice[col], row, || value)?);
                        }
                        cells.push(v);
                    }
                    for s in 0..sys::ZK_NUM_SELECTORS {
                        for row in 0..t.rows {
                            if t.selectors[s * t.rows + row] != 0 {
                                config.selectors[s].enable(&mut region, row)?;
                            }
                        }
                    }
                    for row in 0..t.rows {
                        if t.constants[row] != 0 {
                            region.assign_fixed(|| "constant", config.constants, row,
                                                || Value::known(F::from(t.constants[row])))?;
                        }
                    }
                    // chaining copies come first in a continuing region (they belong to its h input slots)
                    if chained {
                        for i in 0..8 {
                            region.constrain_equal(prev_out[i].cell(), cells[1][t.chain_rows[i] as usize].cell())?;
                        }
                    }
                    // copy_advice(src -> dst) = constrain_equal(src, dst), in the layout's call order
                    for c in &t.copies {
                        region.constrain_equal(cells[c[0] as usize][c[1] as usize].cell(),
                                               cells[c[2] as usize][c[3] as usize].cell())?;
                    }
                    Ok((0..8).map(|i| cells[0][t.chain_rows[8 + i] as usize].clone()).collect::<Vec<_>>())
                },
            )?;
            prev_out = outs;
        }
        Ok(())
    }
}
