"""Symbolic AIR builder: the declarative source of the constraint programs and lookup tables that the
CUDA quotient / LogUp kernels are generated from (codegen.py) and that the host-side checks evaluate.

It plays the role of the reference's builder family: `BfAirBuilder` + `FilteredAirBuilder` semantics
(reference crates/stark/src/air/builder.rs:20-271), `LookupBuilder` / `symbolic_to_virtual_pair`
(crates/stark/src/lookup/builder.rs:10-172), `Lookup` / `LookupKind` (crates/stark/src/lookup/lookup.rs:8-66)
and the constraint folding order of `ProverConstraintFolder::assert_zero` (crates/stark/src/folder.rs:68-72):
constraints are recorded in call order, which is the alpha-folding order.
"""
P = 2130706433

# LookupKind (lookup.rs:19-41)
MEMORY, PROGRAM, ALU, JUMP, MEMINSTR, IO, BYTE = 1, 2, 3, 4, 5, 6, 7
KIND_NAMES = {1: "Memory", 2: "Program", 3: "Alu", 4: "Jump", 5: "MemInstr", 6: "IO", 7: "Byte"}


class Expr:
    """Node of a hash-consed expression DAG over the base field."""
    __slots__ = ("op", "args", "degree", "_id")
    _table = {}
    _next = 0

    def __new__(cls, op, args, degree):
        key = (op, args)
        e = cls._table.get(key)
        if e is None:
            e = object.__new__(cls)
            e.op, e.args, e.degree = op, args, degree
            e._id = Expr._next
            Expr._next += 1
            cls._table[key] = e
        return e

    # -- constructors ---------------------------------------------------------------------------
    @staticmethod
    def const(v):
        return Expr("const", (int(v) % P,), 0)

    @staticmethod
    def var(kind, offset, index):
        """kind in {'prep', 'main'}; offset 0 = local row, 1 = next row."""
        return Expr("var", (kind, offset, index), 1)

    @staticmethod
    def selector(name):
        """'is_first_row' | 'is_last_row' | 'is_transition'.  Degree multiples as Plonky3's symbolic builder
        counts them (is_first/is_last: 1, is_transition: 0), which is what `log_quotient_degree` is derived from."""
        return Expr("sel", (name,), 0 if name == "is_transition" else 1)

    # -- arithmetic --------------------------------------------------------------------------------
    @staticmethod
    def wrap(x):
        return x if isinstance(x, Expr) else Expr.const(x)

    def __add__(self, o):
        o = Expr.wrap(o)
        if self.op == "const" and o.op == "const":
            return Expr.const(self.args[0] + o.args[0])
        if o.op == "const" and o.args[0] == 0:
            return self
        if self.op == "const" and self.args[0] == 0:
            return o
        return Expr("add", (self, o), max(self.degree, o.degree))

    __radd__ = lambda self, o: Expr.wrap(o) + self

    def __sub__(self, o):
        o = Expr.wrap(o)
        if self.op == "const" and o.op == "const":
            return Expr.const(self.args[0] - o.args[0])
        if o.op == "const" and o.args[0] == 0:
            return self
        return Expr("sub", (self, o), max(self.degree, o.degree))

    __rsub__ = lambda self, o: Expr.wrap(o) - self

    def __mul__(self, o):
        o = Expr.wrap(o)
        if self.op == "const" and o.op == "const":
            return Expr.const(self.args[0] * o.args[0])
        for a, b in ((self, o), (o, self)):
            if a.op == "const" and a.args[0] == 1:
                return b
            if a.op == "const" and a.args[0] == 0:
                return Expr.const(0)
        return Expr("mul", (self, o), self.degree + o.degree)

    __rmul__ = lambda self, o: Expr.wrap(o) * self

    def __neg__(self):
        return Expr.const(0) - self

    def __repr__(self):
        if self.op == "const":
            return str(self.args[0])
        if self.op == "var":
            return f"{self.args[0]}[{self.args[1]}][{self.args[2]}]"
        if self.op == "sel":
            return self.args[0]
        sym = {"add": "+", "sub": "-", "mul": "*"}[self.op]
        return f"({self.args[0]} {sym} {self.args[1]})"


def topo_order(roots):
    """Nodes reachable from `roots`, children before parents (iterative DFS)."""
    seen, order = set(), []
    stack = [(r, False) for r in reversed(roots)]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if node._id in seen:
            continue
        seen.add(node._id)
        stack.append((node, True))
        if node.op in ("add", "sub", "mul"):
            for a in reversed(node.args):
                if a._id not in seen:
                    stack.append((a, False))
    return order


class Affine:
    """VirtualPairCol: sum_i w_i * column_i + constant over the LOCAL row (prep columns, then main)."""

    def __init__(self, terms, const):
        acc = {}
        for key, w in terms:
            acc[key] = (acc.get(key, 0) + w) % P
        self.terms = [(k, w) for k, w in acc.items() if w]  # [(('prep'|'main', idx), weight)]
        self.const = const % P

    @staticmethod
    def from_expr(e):
        """symbolic_to_virtual_pair (lookup/builder.rs:109-172): panics on non-affine expressions."""
        e = Expr.wrap(e)
        if e.op == "const":
            return Affine([], e.args[0])
        if e.op == "var":
            kind, offset, idx = e.args
            if offset != 0:
                raise ValueError("lookup value is not an expression of the current row")
            return Affine([((kind, idx), 1)], 0)
        if e.op == "sel":
            raise ValueError("lookup value depends on a row selector")
        a, b = Affine.from_expr(e.args[0]), Affine.from_expr(e.args[1])
        if e.op == "add":
            return Affine(a.terms + b.terms, a.const + b.const)
        if e.op == "sub":
            return Affine(a.terms + [(k, -w) for k, w in b.terms], a.const - b.const)
        if a.terms and b.terms:
            raise ValueError("lookup value is not affine")
        return Affine([(k, w * b.const) for k, w in a.terms] + [(k, w * a.const) for k, w in b.terms], a.const * b.const)


class Lookup:
    def __init__(self, kind, values, multiplicity):
        self.kind = kind
        self.values = [Affine.from_expr(v) for v in values]
        self.multiplicity = Affine.from_expr(multiplicity)


class Row:
    """Named access to the columns of one row (the reference's `#[repr(C)]` cols structs)."""

    def __init__(self, kind, offset, layout):
        object.__setattr__(self, "_kind", kind)
        object.__setattr__(self, "_offset", offset)
        object.__setattr__(self, "_layout", layout)

    def __getattr__(self, name):
        spec = self._layout[name]
        if isinstance(spec, int):
            return Expr.var(self._kind, self._offset, spec)
        return [Expr.var(self._kind, self._offset, i) for i in spec]


def layout(fields):
    """[(name, width)] -> ({name: index | [indices]}, total width); width 1 gives a scalar column."""
    out, pos = {}, 0
    for name, w in fields:
        if w == 1:
            out[name] = pos
        else:
            out[name] = list(range(pos, pos + w))
        pos += w
    return out, pos


class Builder:
    """Records constraints (already multiplied by their `when` conditions) and lookups."""

    def __init__(self, main_layout, prep_layout=None, _shared=None, _cond=None):
        self._main_layout, self._prep_layout = main_layout, prep_layout
        self._s = _shared if _shared is not None else {"constraints": [], "sends": [], "receives": []}
        self._cond = _cond

    # rows
    def main(self, offset=0):
        return Row("main", offset, self._main_layout)

    def preprocessed(self, offset=0):
        return Row("prep", offset, self._prep_layout)

    # p3_air::AirBuilder
    def when(self, cond):
        cond = Expr.wrap(cond)
        return Builder(self._main_layout, self._prep_layout, self._s, cond if self._cond is None else self._cond * cond)

    def when_ne(self, x, y):
        return self.when(Expr.wrap(x) - Expr.wrap(y))

    def when_not(self, cond):  # air/builder.rs:31-33: when_ne(condition, ONE)
        return self.when_ne(cond, 1)

    def when_first_row(self):
        return self.when(Expr.selector("is_first_row"))

    def when_last_row(self):
        return self.when(Expr.selector("is_last_row"))

    def when_transition(self):
        return self.when(Expr.selector("is_transition"))

    def assert_zero(self, x):
        x = Expr.wrap(x)
        self._s["constraints"].append(x if self._cond is None else self._cond * x)

    def assert_eq(self, x, y):
        self.assert_zero(Expr.wrap(x) - Expr.wrap(y))

    def assert_one(self, x):
        self.assert_zero(Expr.wrap(x) - 1)

    def assert_bool(self, x):
        x = Expr.wrap(x)
        self.assert_zero(x * (x - 1))

    # MessageBuilder (conditions do not apply to lookups: FilteredAirBuilder forwards them unchanged)
    def send(self, kind, values, multiplicity):
        self._s["sends"].append(Lookup(kind, values, multiplicity))

    def receive(self, kind, values, multiplicity):
        self._s["receives"].append(Lookup(kind, values, multiplicity))

    # ByteAirBuilder / InstructionAirBuilder (air/builder.rs:51-229)
    def send_byte(self, opcode, a, b, mult):
        self.send(BYTE, [opcode, a, b], mult)

    def receive_byte(self, opcode, a, b, mult):
        self.receive(BYTE, [opcode, a, b], mult)

    def send_alu(self, pc, opcode, next_mv, mv, mult):
        self.send(ALU, [pc, opcode, next_mv, mv], mult)

    def receive_alu(self, pc, opcode, next_mv, mv, mult):
        self.receive(ALU, [pc, opcode, next_mv, mv], mult)

    def send_jump(self, pc, next_pc, opcode, mv, mult):
        self.send(JUMP, [pc, next_pc, opcode, mv], mult)

    def receive_jump(self, pc, next_pc, opcode, mv, mult):
        self.receive(JUMP, [pc, next_pc, opcode, mv], mult)

    def send_memory_instr(self, clk, pc, opcode, mp, next_mp, mult):
        self.send(MEMINSTR, [clk, pc, opcode, mp, next_mp], mult)

    def receive_memory_instr(self, clk, pc, opcode, mp, next_mp, mult):
        self.receive(MEMINSTR, [clk, pc, opcode, mp, next_mp], mult)

    def send_io(self, pc, opcode, mp, mv, mult):
        self.send(IO, [pc, opcode, mp, mv], mult)

    def receive_io(self, pc, opcode, mp, mv, mult):
        self.receive(IO, [pc, opcode, mp, mv], mult)

    @property
    def constraints(self):
        return self._s["constraints"]

    @property
    def sends(self):
        return self._s["sends"]

    @property
    def receives(self):
        return self._s["receives"]


class Chip:
    """`Chip::new` (crates/stark/src/chip.rs:65-90): captured lookups, constraint list, quotient degree."""

    def __init__(self, name, main_width, prep_width, local_only, eval_fn, main_layout, prep_layout=None):
        self.name, self.main_width, self.prep_width, self.local_only = name, main_width, prep_width, local_only
        b = Builder(main_layout, prep_layout)
        eval_fn(b)
        self.constraints, self.sends, self.receives = b.constraints, b.sends, b.receives
        max_deg = max([c.degree for c in self.constraints], default=0)
        if self.sends or self.receives:
            max_deg = max(max_deg, 3)
        self.max_constraint_degree = max_deg
        self.log_quotient_degree = max(0, (max_deg - 1 - 1).bit_length()) if max_deg > 1 else 0  # log2_ceil(max_deg - 1)
        self.batch_size = 1 << self.log_quotient_degree  # logup_batch_size (chip.rs:157-160)
        n_lookups = len(self.sends) + len(self.receives)
        # permutation_trace_width (permutation.rs:15-21), in extension columns
        self.perm_width = 0 if n_lookups == 0 else -(-n_lookups // self.batch_size) + 1

    @property
    def lookups(self):
        """sends then receives, with their sign: the chunking order of populate_permutation_row."""
        return [(l, True) for l in self.sends] + [(l, False) for l in self.receives]

    def uses_next_row(self):
        return any(n.op == "var" and n.args[1] == 1 for n in topo_order(self.constraints))
