//! Lowering of `LiftedAir::eval` to the flat op-list the backend's constraint kernel interprets (`mdn_air.program`,
//! include/miden_b200.h).
//!
//! The capture is the one the reference already performs for its ACE circuits and for degree analysis: run `eval` on
//! the `SymbolicAirBuilder` (crates/lifted-air/src/air.rs:151-165 `ConstraintDegrees::from_air`;
//! crates/ace-codegen/src/pipeline.rs:96-111) and read back the base / extension constraint expressions together
//! with the `ConstraintLayout` that records their GLOBAL emission positions.  The walk over the expression trees
//! follows `lower_base_expr` / `lower_ext_expr` (crates/ace-codegen/src/dag/lower.rs:105-210) with three differences
//! that the device format wants: aux (permutation) columns stay extension-typed leaves (`AUX`) instead of being
//! rebuilt from coordinates, periodic columns stay leaves (`PERIODIC`; the backend evaluates them on the LDE coset
//! like `PeriodicLde`, prover/periodic.rs:49-98), and preprocessed columns are supported (`PREPROCESSED`).
//!
//! Program format (little-endian u32 words):
//!   [0x5249414D "MAIR", 1, n_nodes, n_constraints, n_consts | nodes (op, a, b) | constraint node ids in emission
//!    order | consts as (lo, hi) word pairs]
//! Constraints are folded on the device as `acc <- acc * alpha + C_k` in emission order, base and extension
//! constraints interleaved exactly as `eval` emitted them -- the orientation of the verifier
//! (crates/lifted-stark/src/verifier/constraints.rs:83,108).
//!
//! NOT COMPILED in this repository (no Rust toolchain in the build image); written against the sources as read.

use std::{collections::HashMap, sync::Arc};

use miden_core::{Felt, field::{BasedVectorSpace, PrimeField64, QuadFelt}};
use miden_crypto::stark::air::{
    BaseAir, LiftedAir,
    symbolic::{
        BaseEntry, BaseLeaf, ConstraintLayout, ExtEntry, ExtLeaf, SymbolicAirBuilder, SymbolicExpression,
        SymbolicExpressionExt,
    },
};

pub const MAGIC_AIR: u32 = 0x5249_414D; // "MAIR"
pub const MAGIC_LOOKUP: u32 = 0x504B_4C4D; // "MLKP"

/// Node opcodes of the device format (include/miden_b200.h, `mdn_air.program`).
#[repr(u32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq, Hash)]
pub enum Op {
    Main = 0,         // a = row offset (0 | 1), b = column
    Aux = 1,          // a = row offset, b = EF column
    Public = 2,       // a = index into the public values
    Challenge = 3,    // a = index into the shared randomness pool
    AuxValue = 4,     // a = index into this AIR's aux (permutation) values
    IsFirstRow = 5,
    IsLastRow = 6,
    IsTransition = 7,
    Const = 8,        // a = constant slot
    ExtConst = 9,     // a = constant slot of c0; c1 is slot a + 1
    Add = 10,
    Sub = 11,
    Mul = 12,
    Neg = 13,
    Periodic = 14,    // a = periodic column
    Preprocessed = 15, // a = row offset, b = column
}

/// Hash-consing builder of one op-list: identical sub-expressions -- the symbolic builder shares them through `Arc`,
/// and AIR code re-derives many of them -- become ONE node, which is what keeps the Miden AIRs (~5.2 k gates in the
/// reference's ACE circuit, air/src/lib.rs:567) near that size here too.
#[derive(Default)]
pub struct OpList {
    nodes: Vec<[u32; 3]>,
    dedup: HashMap<[u32; 3], u32>,
    consts: Vec<u64>,
    const_slot: HashMap<u64, u32>,
    ext_const_slot: HashMap<(u64, u64), u32>,
}

impl OpList {
    pub fn node(&mut self, op: Op, a: u32, b: u32) -> u32 {
        let key = [op as u32, a, b];
        if let Some(&id) = self.dedup.get(&key) {
            return id;
        }
        let id = self.nodes.len() as u32;
        self.nodes.push(key);
        self.dedup.insert(key, id);
        id
    }
    pub fn constant(&mut self, c: Felt) -> u32 {
        let v = c.as_canonical_u64();
        let slot = *self.const_slot.entry(v).or_insert_with(|| {
            self.consts.push(v);
            (self.consts.len() - 1) as u32
        });
        self.node(Op::Const, slot, 0)
    }
    pub fn ext_constant(&mut self, c: QuadFelt) -> u32 {
        let co: &[Felt] = c.as_basis_coefficients_slice();
        let (c0, c1) = (co[0].as_canonical_u64(), co[1].as_canonical_u64());
        if c1 == 0 {
            return self.constant(co[0]); // embeds the base field: keeps the node base-typed on the device
        }
        let slot = *self.ext_const_slot.entry((c0, c1)).or_insert_with(|| {
            self.consts.push(c0);
            self.consts.push(c1);
            (self.consts.len() - 2) as u32
        });
        self.node(Op::ExtConst, slot, 0)
    }
    pub fn num_nodes(&self) -> usize {
        self.nodes.len()
    }
    /// Serialise with the given trailer items (`item_words` words each): constraint ids for an AIR program,
    /// `{column, flag | !0, multiplicity, denominator}` records for a lookup program.
    pub fn serialize(&self, magic: u32, items: &[u32], item_words: usize) -> Vec<u32> {
        assert!(items.len() % item_words == 0);
        let mut w = Vec::with_capacity(5 + 3 * self.nodes.len() + items.len() + 2 * self.consts.len());
        w.extend_from_slice(&[magic, 1, self.nodes.len() as u32, (items.len() / item_words) as u32, self.consts.len() as u32]);
        for n in &self.nodes {
            w.extend_from_slice(n);
        }
        w.extend_from_slice(items);
        for &c in &self.consts {
            w.push(c as u32);
            w.push((c >> 32) as u32);
        }
        w
    }
}

/// Expression walker with `Arc`-identity memoisation (the symbolic trees are DAGs).
pub struct Lowerer<'a> {
    pub ops: &'a mut OpList,
    base_memo: HashMap<*const SymbolicExpression<Felt>, u32>,
    ext_memo: HashMap<*const SymbolicExpressionExt<Felt, QuadFelt>, u32>,
}

impl<'a> Lowerer<'a> {
    pub fn new(ops: &'a mut OpList) -> Self {
        Self { ops, base_memo: HashMap::new(), ext_memo: HashMap::new() }
    }

    fn base_arc(&mut self, e: &Arc<SymbolicExpression<Felt>>) -> u32 {
        let key = Arc::as_ptr(e);
        if let Some(&id) = self.base_memo.get(&key) {
            return id;
        }
        let id = self.base(e);
        self.base_memo.insert(key, id);
        id
    }
    fn ext_arc(&mut self, e: &Arc<SymbolicExpressionExt<Felt, QuadFelt>>) -> u32 {
        let key = Arc::as_ptr(e);
        if let Some(&id) = self.ext_memo.get(&key) {
            return id;
        }
        let id = self.ext(e);
        self.ext_memo.insert(key, id);
        id
    }

    /// crates/ace-codegen/src/dag/lower.rs:105-155 (`lower_base_expr`), device vocabulary.
    pub fn base(&mut self, expr: &SymbolicExpression<Felt>) -> u32 {
        match expr {
            SymbolicExpression::Leaf(leaf) => match leaf {
                BaseLeaf::Variable(v) => match v.entry {
                    BaseEntry::Main { offset } => self.ops.node(Op::Main, offset as u32, v.index as u32),
                    BaseEntry::Public => self.ops.node(Op::Public, v.index as u32, 0),
                    BaseEntry::Periodic => self.ops.node(Op::Periodic, v.index as u32, 0),
                    BaseEntry::Preprocessed { offset } => self.ops.node(Op::Preprocessed, offset as u32, v.index as u32),
                },
                BaseLeaf::IsFirstRow => self.ops.node(Op::IsFirstRow, 0, 0),
                BaseLeaf::IsLastRow => self.ops.node(Op::IsLastRow, 0, 0),
                BaseLeaf::IsTransition => self.ops.node(Op::IsTransition, 0, 0),
                BaseLeaf::Constant(c) => self.ops.constant(*c),
            },
            SymbolicExpression::Add { x, y, .. } => {
                let (a, b) = (self.base_arc(x), self.base_arc(y));
                self.ops.node(Op::Add, a, b)
            },
            SymbolicExpression::Sub { x, y, .. } => {
                let (a, b) = (self.base_arc(x), self.base_arc(y));
                self.ops.node(Op::Sub, a, b)
            },
            SymbolicExpression::Mul { x, y, .. } => {
                let (a, b) = (self.base_arc(x), self.base_arc(y));
                self.ops.node(Op::Mul, a, b)
            },
            SymbolicExpression::Neg { x, .. } => {
                let a = self.base_arc(x);
                self.ops.node(Op::Neg, a, 0)
            },
        }
    }

    /// crates/ace-codegen/src/dag/lower.rs:158-210 (`lower_ext_expr`), device vocabulary.
    pub fn ext(&mut self, expr: &SymbolicExpressionExt<Felt, QuadFelt>) -> u32 {
        match expr {
            SymbolicExpressionExt::Leaf(leaf) => match leaf {
                ExtLeaf::Base(b) => self.base(b),
                ExtLeaf::ExtVariable(v) => match v.entry {
                    ExtEntry::Permutation { offset } => self.ops.node(Op::Aux, offset as u32, v.index as u32),
                    ExtEntry::Challenge => self.ops.node(Op::Challenge, v.index as u32, 0),
                    ExtEntry::PermutationValue => self.ops.node(Op::AuxValue, v.index as u32, 0),
                },
                ExtLeaf::ExtConstant(c) => self.ops.ext_constant(*c),
            },
            SymbolicExpressionExt::Add { x, y, .. } => {
                let (a, b) = (self.ext_arc(x), self.ext_arc(y));
                self.ops.node(Op::Add, a, b)
            },
            SymbolicExpressionExt::Sub { x, y, .. } => {
                let (a, b) = (self.ext_arc(x), self.ext_arc(y));
                self.ops.node(Op::Sub, a, b)
            },
            SymbolicExpressionExt::Mul { x, y, .. } => {
                let (a, b) = (self.ext_arc(x), self.ext_arc(y));
                self.ops.node(Op::Mul, a, b)
            },
            SymbolicExpressionExt::Neg { x, .. } => {
                let a = self.ext_arc(x);
                self.ops.node(Op::Neg, a, 0)
            },
        }
    }
}

/// Everything `mdn_air` points at, owned.
pub struct LoweredAir {
    pub width: u32,
    pub aux_width: u32,
    pub num_aux_values: u32,
    pub num_randomness: u32,
    pub log_quotient_degree: u32,
    pub preprocessed_width: u32,
    pub program: Vec<u32>,
    /// `BaseAir::periodic_columns_matrix()` row-major (max_period x n_cols), every column repeated to the max period
    pub periodic_values: Vec<u64>,
    pub num_periodic_columns: u32,
    pub log_max_period: u32,
    pub num_constraints: usize,
    pub num_nodes: usize,
}

/// `ceil(log2(n))` for n >= 1 (crates/lifted-air/src/util.rs `log2_ceil_u8`).
fn log2_ceil(n: usize) -> u32 {
    if n <= 1 { 0 } else { usize::BITS - (n - 1).leading_zeros() }
}

/// Lower one AIR.  Run once per AIR type and cache: the result depends only on the AIR definition.
pub fn lower_air<A>(air: &A) -> LoweredAir
where
    A: LiftedAir<Felt, QuadFelt>,
{
    // symbolic capture: crates/lifted-air/src/air.rs:151-165, crates/ace-codegen/src/pipeline.rs:96-111
    let mut builder = SymbolicAirBuilder::<Felt, QuadFelt>::new(air.air_layout());
    air.eval(&mut builder);
    let layout: ConstraintLayout = builder.constraint_layout();
    let base = builder.base_constraints();
    let ext = builder.extension_constraints();

    // global emission order: base and extension constraints interleaved by their recorded positions
    // (crates/ace-codegen/src/dag/lower.rs:236-246)
    let mut ordered: Vec<(usize, bool, usize)> = Vec::with_capacity(base.len() + ext.len());
    ordered.extend(layout.base_indices.iter().enumerate().map(|(i, &pos)| (pos, false, i)));
    ordered.extend(layout.ext_indices.iter().enumerate().map(|(i, &pos)| (pos, true, i)));
    ordered.sort_by_key(|&(pos, ..)| pos);

    let mut ops = OpList::default();
    let mut constraint_ids = Vec::with_capacity(ordered.len());
    {
        let mut lw = Lowerer::new(&mut ops);
        for &(_, is_ext, idx) in &ordered {
            constraint_ids.push(if is_ext { lw.ext(&ext[idx]) } else { lw.base(&base[idx]) });
        }
    }

    // log_quotient_degree: crates/lifted-stark/src/domain.rs:585-598 on the same symbolic pass
    let max_degree = base
        .iter()
        .map(SymbolicExpression::degree_multiple)
        .chain(ext.iter().map(SymbolicExpressionExt::degree_multiple))
        .max()
        .unwrap_or(0);
    let log_quotient_degree = log2_ceil(max_degree.saturating_sub(1).max(1));

    // periodic columns: the repeated matrix the reference hands to PeriodicLde (prover/periodic.rs:38-47)
    let cols = air.periodic_columns();
    let max_period = cols.iter().map(Vec::len).max().unwrap_or(0);
    let mut periodic_values = Vec::with_capacity(max_period * cols.len());
    for r in 0..max_period {
        for c in &cols {
            periodic_values.push(c[r % c.len()].as_canonical_u64());
        }
    }

    LoweredAir {
        width: air.width() as u32,
        aux_width: air.aux_width() as u32,
        num_aux_values: air.num_aux_values() as u32,
        num_randomness: air.num_randomness() as u32,
        log_quotient_degree,
        preprocessed_width: air.preprocessed_width() as u32,
        program: ops.serialize(MAGIC_AIR, &constraint_ids, 1),
        periodic_values,
        num_periodic_columns: cols.len() as u32,
        log_max_period: if max_period > 1 { max_period.trailing_zeros() } else { 0 },
        num_constraints: constraint_ids.len(),
        num_nodes: ops.num_nodes(),
    }
}
