"""Minimal PQL call tree + parser for the hot-path calls (mirror of pql.Call, reference pql/ast.go), so that parity
tests read like the reference's executor tests.  The full PEG grammar stays in Go; only this subset is needed:
Row, Intersect, Union, Difference, Xor, Not, All, Count, TopN, TopK, Rows, GroupBy, Sum, Min, Max, Percentile."""
import re


class Condition:
    """pql.Condition: op in {==, !=, <, <=, >, >=, ><}; value int or [lo, hi] (><) or None (null)"""

    def __init__(self, op, value):
        self.op, self.value = op, value

    def __repr__(self):
        return f"Condition({self.op!r}, {self.value!r})"


class Call:
    def __init__(self, name, args=None, children=None):
        self.name, self.args, self.children = name, dict(args or {}), list(children or [])

    def __repr__(self):
        parts = [repr(c) for c in self.children] + [f"{k}={v!r}" for k, v in self.args.items()]
        return f"{self.name}({', '.join(parts)})"


_TOK = re.compile(r"\s*(?:(?P<str>\"[^\"]*\"|'[^']*')|(?P<ts>\d{4}-\d{2}-\d{2}T\d{2}:\d{2})|(?P<flt>-?\d+\.\d+)|(?P<num>-?\d+)|(?P<id>[A-Za-z_][A-Za-z0-9_\-]*)|(?P<op>><|<=|>=|==|!=|[(),=<>\[\]]))")


def _tokens(s):
    pos, out = 0, []
    while pos < len(s):
        if s[pos:].strip() == "":
            break
        m = _TOK.match(s, pos)
        if not m:
            raise ValueError(f"PQL syntax error at {pos}: {s[pos:pos + 20]!r}")
        pos = m.end()
        if m.group("str") is not None:
            out.append(("str", m.group("str")[1:-1]))
        elif m.group("ts") is not None:
            out.append(("str", m.group("ts")))                    # timestamp literal 2006-01-02T15:04 (from= / to=)
        elif m.group("flt") is not None:
            out.append(("flt", float(m.group("flt"))))
        elif m.group("num") is not None:
            out.append(("num", int(m.group("num"))))
        elif m.group("id") is not None:
            out.append(("id", m.group("id")))
        else:
            out.append(("op", m.group("op")))
    return out


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, kind, val=None):
        tok = self.next()
        if tok[0] != kind or (val is not None and tok[1] != val):
            raise ValueError(f"PQL: expected {val or kind}, got {tok}")
        return tok[1]

    def value(self):
        kind, v = self.next()
        if kind in ("num", "str", "flt"):
            return v
        if kind == "id":
            if v == "null":
                return None
            if v in ("true", "false"):
                return v == "true"
            return v
        if (kind, v) == ("op", "["):
            vals = []
            while self.peek() != ("op", "]"):
                vals.append(self.value())
                if self.peek() == ("op", ","):
                    self.next()
            self.next()
            return vals
        raise ValueError(f"PQL: unexpected token {v!r}")

    def call(self):
        name = self.expect("id")
        self.expect("op", "(")
        c = Call(name)
        while self.peek() != ("op", ")"):
            k0, k1 = self.peek(), self.peek(1)
            if k0[0] == "id" and k1 == ("op", "("):
                c.children.append(self.call())
            elif k0[0] == "id" and k1 == ("op", "="):
                key = self.next()[1]
                self.next()
                if self.peek()[0] == "id" and self.peek(1) == ("op", "("):
                    c.args[key] = self.call()
                else:
                    c.args[key] = self.value()
            elif k0[0] == "id" and k1[0] == "op" and k1[1] in ("<", "<=", ">", ">=", "==", "!=", "><"):
                key = self.next()[1]
                op = self.next()[1]
                c.args[key] = Condition(op, self.value())
            elif k0[0] == "num" and k1[0] == "op" and k1[1] in ("<", "<="):
                # lo < f < hi  (pql BTWN_LT_LT etc.; normalised to inclusive ><)
                lo = self.next()[1]
                op1 = self.next()[1]
                key = self.expect("id")
                op2 = self.next()[1]
                hi = self.value()
                c.args[key] = Condition("><", [lo + (1 if op1 == "<" else 0), hi - (1 if op2 == "<" else 0)])
            elif k0[0] == "id":
                # bare positional field name (TopN(f, ...), Rows(f))
                c.args["_field"] = self.next()[1]
            else:
                raise ValueError(f"PQL: unexpected token {k0}")
            if self.peek() == ("op", ","):
                self.next()
        self.next()
        return c


def parse(s):
    p = _Parser(_tokens(s))
    calls = []
    while p.peek()[0] is not None:
        calls.append(p.call())
    return calls
