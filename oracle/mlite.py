"""TEST INFRASTRUCTURE ONLY: a small interpreter for the subset of the Matlab language that the reference's sampler helpers
are written in, so that their SOURCE FILES can be executed here (this image has neither Matlab nor Octave).

It runs the unmodified text of
    matlab/samplers/{qmcnodes, randref, essinv, hellinger, iw_prune, mcmc_prune, tt_irt_sqr}.m
(tests/golden/make_golden_matlab.py) and so pins the numpy restatements under oracle/ against the reference's own code instead
of against a second reading of it.  What it is NOT: Matlab's numerical library.  Elementary functions, sums, matrix products and
the QR factorisation come from numpy / LAPACK, so results agree with a real Matlab run to rounding (summation order, libm),
not bit for bit; the comparisons that use it carry tolerances accordingly.

Supported: function files (one or more functions), assignments with () / {} indexing on the left (nested: h{k}(2:n) = ...,
automatic growth), [a, b] = f(...) and [~, b] = f(...), if / elseif / else, for, while, break, continue, return, the operators
+ - * / ^ .* ./ .^ ' .' < <= > >= == ~= ~ & | && || and ranges a:b, a:s:b, matrices [..., ...; ...], cell arrays, strings,
anonymous functions, read access to struct fields (a Python dict), c{:} argument expansion, `end` inside indices, implicit expansion of singleton dimensions, N-d arrays in
column-major order, and the builtins in Interp.builtins.  Anything else raises MatlabError.
"""
import math
import re

import numpy as np


class MatlabError(Exception):
    pass


class MStr(str):
    """a Matlab char row vector"""


class MCell(object):
    def __init__(self, shape):
        self.a = np.empty(shape, dtype=object)
        for i in np.ndindex(*shape):
            self.a[i] = zeros((0, 0))


class MFunc(object):
    def __init__(self, params, body, env):
        self.params, self.body, self.env = params, body, env


class CSList(list):
    """comma-separated list (c{:})"""


def zeros(shape):
    return np.zeros(shape, dtype=np.float64, order="F")


def mat(x):
    """anything numeric -> float64 / bool ndarray with ndim >= 2"""
    if isinstance(x, np.ndarray):
        if x.ndim >= 2:
            return x
        return x.reshape((1, -1) if x.ndim == 1 else (1, 1), order="F")
    if isinstance(x, (bool, np.bool_)):
        return np.array([[bool(x)]])
    if isinstance(x, (int, float, np.integer, np.floating)):
        return np.array([[float(x)]])
    if isinstance(x, MStr):
        return np.array([[float(ord(c)) for c in x]]).reshape(1, len(x))
    raise MatlabError("not numeric: %r" % (type(x),))


def scalar(x):
    a = mat(x)
    if a.size != 1:
        raise MatlabError("scalar expected, got size %s" % (a.shape,))
    return float(a.reshape(-1)[0])


def truth(x):
    a = mat(x)
    return a.size > 0 and bool(np.all(a != 0))


def squeeze_trailing(a):
    while a.ndim > 2 and a.shape[-1] == 1:
        a = a.reshape(a.shape[:-1], order="F")
    return a


# ------------------------------------------------------------------------------------------------------------------
# tokens
# ------------------------------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<num>(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<field>\.[A-Za-z_]\w*)
  | (?P<op>\.\*|\./|\.\^|\.'|==|~=|<=|>=|&&|\|\||[-+*/\\^<>=&|~(){}\[\],;:'@])
  | (?P<nl>\n)
  | (?P<ws>[ \t\r]+)
""", re.X)

KEYWORDS = {"function", "if", "elseif", "else", "end", "for", "parfor", "while", "break", "continue", "return"}


def tokenize(src):
    toks, i, n = [], 0, len(src)
    depth = 0                                   # inside ( [ {: newlines do not end statements in ( ), `end` is an index
    while i < n:
        c = src[i]
        if c == "%":                            # comment to end of line
            while i < n and src[i] != "\n":
                i += 1
            continue
        if src.startswith("...", i):            # continuation
            while i < n and src[i] != "\n":
                i += 1
            i += 1
            continue
        if c == "'":
            prev = toks[-1] if toks else None
            # a quote is the transpose operator right behind a value (number, name, closing bracket, another transpose)
            is_transpose = prev is not None and ((prev[0] in ("num", "field")) or (prev[0] == "id" and (prev[1] not in KEYWORDS or prev[1] == "end")) or
                                                 (prev[0] == "op" and prev[1] in (")", "]", "}", "'", ".'")))
            if is_transpose:
                toks.append(("op", "'", False)); i += 1
                continue
            j, chars = i + 1, []
            while True:
                if j >= n:
                    raise MatlabError("unterminated string")
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        chars.append("'"); j += 2
                        continue
                    break
                chars.append(src[j]); j += 1
            toks.append(("str", "".join(chars), False)); i = j + 1
            continue
        m = TOKEN_RE.match(src, i)
        if not m:
            raise MatlabError("cannot tokenize at %r" % src[i:i + 20])
        i = m.end()
        if m.lastgroup == "ws":
            if toks:
                toks[-1] = (toks[-1][0], toks[-1][1], True)      # followed by whitespace
            continue
        kind = m.lastgroup
        text = m.group(kind)
        if kind == "op":
            if text in "([{":
                depth += 1
            elif text in ")]}":
                depth -= 1
        if kind == "nl":
            if depth > 0:
                continue
            toks.append(("nl", "\n", False))
            continue
        toks.append((kind, text, False))
    toks.append(("eof", "", False))
    return toks


# ------------------------------------------------------------------------------------------------------------------
# parser -> nested tuples
# ------------------------------------------------------------------------------------------------------------------
class Parser(object):
    def __init__(self, src):
        self.t = tokenize(src)
        self.p = 0
        self.idx_depth = 0

    def peek(self):
        return self.t[self.p]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def at(self, kind, text=None):
        k, v, _ = self.t[self.p]
        return k == kind and (text is None or v == text)

    def at_op(self, *texts):
        k, v, _ = self.t[self.p]
        return k == "op" and v in texts

    def expect_op(self, text):
        if not self.at_op(text):
            raise MatlabError("expected %r, got %r" % (text, self.peek()[:2]))
        self.next()

    def skip_seps(self):
        while self.at("nl") or self.at_op(";", ","):
            self.next()

    # ---- file / statements ----
    def parse_file(self):
        funcs = {}
        self.skip_seps()
        while not self.at("eof"):
            f = self.parse_function()
            funcs.setdefault("__first__", f[0])
            funcs[f[0]] = f
            self.skip_seps()
        return funcs

    def parse_function(self):
        if not self.at("id", "function"):
            raise MatlabError("function expected, got %r" % (self.peek()[:2],))
        self.next()
        outs = []
        if self.at_op("["):
            self.next()
            while not self.at_op("]"):
                if self.at_op(","):
                    self.next()
                    continue
                outs.append(self.next()[1])
            self.next()
            self.expect_op("=")
            name = self.next()[1]
        else:
            name = self.next()[1]
            if self.at_op("="):
                self.next()
                outs = [name]
                name = self.next()[1]
        params = []
        if self.at_op("("):
            self.next()
            while not self.at_op(")"):
                if self.at_op(","):
                    self.next()
                    continue
                params.append(self.next()[1])
            self.next()
        body = self.parse_block(("end", "function"))
        if self.at("id", "end"):
            self.next()
        return (name, params, outs, body)

    def parse_block(self, terminators):
        stmts = []
        while True:
            self.skip_seps()
            if self.at("eof"):
                break
            if self.peek()[0] == "id" and self.peek()[1] in terminators:
                break
            stmts.append(self.parse_statement())
        return stmts

    def parse_statement(self):
        k, v, _ = self.peek()
        if k == "id" and v == "if":
            self.next()
            clauses, other = [], None
            cond = self.parse_expr()
            body = self.parse_block(("elseif", "else", "end"))
            clauses.append((cond, body))
            while True:
                if self.at("id", "elseif"):
                    self.next()
                    cond = self.parse_expr()
                    clauses.append((cond, self.parse_block(("elseif", "else", "end"))))
                elif self.at("id", "else"):
                    self.next()
                    other = self.parse_block(("end",))
                else:
                    break
            self.next()  # end
            return ("if", clauses, other)
        if k == "id" and v in ("for", "parfor"):          # (parfor runs as a plain loop)
            self.next()
            paren = self.at_op("(")
            if paren:
                self.next()
            var = self.next()[1]
            self.expect_op("=")
            rng = self.parse_expr()
            if paren:
                self.expect_op(")")
            body = self.parse_block(("end",))
            self.next()
            return ("for", var, rng, body)
        if k == "id" and v == "while":
            self.next()
            cond = self.parse_expr()
            body = self.parse_block(("end",))
            self.next()
            return ("while", cond, body)
        if k == "id" and v in ("break", "continue", "return"):
            self.next()
            return (v,)
        # [a, b] = f(...)
        if k == "op" and v == "[":
            save = self.p
            try:
                lhs = self.try_multi_lhs()
            except MatlabError:
                lhs = None
            if lhs is not None and self.at_op("="):
                self.next()
                return ("massign", lhs, self.parse_expr())
            self.p = save
        e = self.parse_expr()
        if self.at_op("="):
            self.next()
            return ("assign", e, self.parse_expr())
        return ("expr", e)

    def try_multi_lhs(self):
        self.expect_op("[")
        lhs = []
        while not self.at_op("]"):
            if self.at_op(","):
                self.next()
                continue
            if self.at_op("~"):
                self.next()
                lhs.append(None)
                continue
            lhs.append(self.parse_postfix())
        self.next()
        return lhs

    # ---- expressions ----
    def parse_expr(self):
        return self.parse_binary(0)

    LEVELS = [("||",), ("&&",), ("|",), ("&",), ("<", "<=", ">", ">=", "==", "~=")]

    def parse_binary(self, lvl):
        if lvl == len(self.LEVELS):
            return self.parse_range()
        left = self.parse_binary(lvl + 1)
        while self.at_op(*self.LEVELS[lvl]):
            op = self.next()[1]
            left = ("bin", op, left, self.parse_binary(lvl + 1))
        return left

    def parse_range(self):
        first = self.parse_additive()
        if self.at_op(":") and not self.colon_is_bare():
            self.next()
            second = self.parse_additive()
            if self.at_op(":") and not self.colon_is_bare():
                self.next()
                third = self.parse_additive()
                return ("range", first, second, third)
            return ("range", first, None, second)
        return first

    def colon_is_bare(self):
        k, v, _ = self.t[self.p + 1]
        return k == "op" and v in (")", ",", "}")

    def parse_additive(self):
        left = self.parse_mul()
        while self.at_op("+", "-"):
            op = self.next()[1]
            left = ("bin", op, left, self.parse_mul())
        return left

    def parse_mul(self):
        left = self.parse_unary()
        while self.at_op("*", "/", ".*", "./", "\\"):
            op = self.next()[1]
            left = ("bin", op, left, self.parse_unary())
        return left

    def parse_unary(self):
        if self.at_op("-", "+", "~"):
            op = self.next()[1]
            return ("un", op, self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        base = self.parse_postfix()
        while self.at_op("^", ".^"):
            op = self.next()[1]
            if self.at_op("-", "+", "~"):
                uop = self.next()[1]
                expo = ("un", uop, self.parse_postfix())
            else:
                expo = self.parse_postfix()
            base = ("bin", op, base, expo)
        return base

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            if self.at_op("("):
                self.next()
                e = ("index", e, self.parse_args(")"))
            elif self.at_op("{"):
                self.next()
                e = ("cindex", e, self.parse_args("}"))
            elif self.at_op("'", ".'"):
                self.next()
                e = ("transpose", e)
            elif self.at("field"):
                e = ("field", e, self.next()[1][1:])
            else:
                return e

    def parse_args(self, close):
        args = []
        self.idx_depth += 1
        while not self.at_op(close):
            if self.at_op(","):
                self.next()
                continue
            if self.at_op(":") and self.colon_is_bare_here(close):
                self.next()
                args.append(("colon",))
                continue
            args.append(self.parse_expr())
        self.next()
        self.idx_depth -= 1
        return args

    def colon_is_bare_here(self, close):
        k, v, _ = self.t[self.p + 1]
        return k == "op" and v in (close, ",")

    def parse_primary(self):
        k, v, _ = self.next()
        if k == "num":
            return ("num", float(v))
        if k == "str":
            return ("str", v)
        if k == "id":
            if v == "end" and self.idx_depth > 0:
                return ("endidx",)
            if v in KEYWORDS:
                raise MatlabError("unexpected keyword %r" % v)
            return ("id", v)
        if k == "op" and v == "(":
            self.idx_depth, save = 0, self.idx_depth
            e = self.parse_expr()
            self.idx_depth = save
            self.expect_op(")")
            return ("paren", e)
        if k == "op" and v == "[":
            rows, row = [], []
            while True:
                if self.at_op("]"):
                    self.next()
                    break
                if self.at_op(","):
                    self.next()
                    continue
                if self.at_op(";") or self.at("nl"):
                    self.next()
                    rows.append(row)
                    row = []
                    continue
                row.append(self.parse_expr())
            rows.append(row)
            return ("matrix", [r for r in rows if r])
        if k == "op" and v == "{":
            items = []
            while not self.at_op("}"):
                if self.at_op(",") or self.at_op(";"):
                    self.next()
                    continue
                items.append(self.parse_expr())
            self.next()
            return ("cellrow", items)
        if k == "op" and v == "@":
            if self.at_op("("):
                self.next()
                params = []
                while not self.at_op(")"):
                    if self.at_op(","):
                        self.next()
                        continue
                    params.append(self.next()[1])
                self.next()
                return ("anon", params, self.parse_expr())
            return ("fhandle", self.next()[1])
        raise MatlabError("unexpected token %r" % ((k, v),))


# ------------------------------------------------------------------------------------------------------------------
# evaluation
# ------------------------------------------------------------------------------------------------------------------
class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    pass


def _index_vector(ix, dimlen):
    """one subscript -> 0-based integer index array (flat) and its shape as written"""
    if isinstance(ix, tuple) and ix == ("colon",):
        return np.arange(dimlen), None
    a = mat(ix)
    if a.dtype == np.bool_:
        return np.flatnonzero(a.reshape(-1, order="F")), (-1,)
    v = a.reshape(-1, order="F")
    iv = np.rint(v).astype(np.int64)
    if v.size and (np.any(np.abs(v - iv) > 0) or np.any(iv < 1)):
        raise MatlabError("bad subscript")
    return iv - 1, a.shape


class Interp(object):
    def __init__(self, rand_stream=None, files=None, externals=None):
        """rand_stream: callable(shape) -> array of uniforms (Matlab's rand); files: name -> array for load();
        externals: name -> python callable(interp, args, nargout) for functions implemented outside (MEX files)."""
        self.funcs = {}
        self.rand_stream = rand_stream
        self.files = files or {}
        self.externals = externals or {}
        self.printed = []

    def load_source(self, src):
        fs = Parser(src).parse_file()
        first = fs.pop("__first__")
        self.funcs.update(fs)
        return first

    def load_file(self, path):
        with open(path) as fh:
            return self.load_source(fh.read())

    # ---- calls ----
    def call(self, name, args, nargout=1):
        if name in self.externals:
            out = self.externals[name](self, args, nargout)
            return list(out) if isinstance(out, (list, tuple)) else [out]
        if name in self.funcs:
            _, params, outs, body = self.funcs[name]
            env = {}
            nfixed = len(params) - (1 if params and params[-1] == "varargin" else 0)
            for p_, a_ in zip(params[:nfixed], args[:nfixed]):
                env[p_] = a_
            if params and params[-1] == "varargin":
                c = MCell((1, max(0, len(args) - nfixed)))
                for i, a_ in enumerate(args[nfixed:]):
                    c.a[0, i] = a_
                env["varargin"] = c
            env["__nargin__"] = len(args)
            env["__nargout__"] = nargout
            try:
                self.exec_block(body, env)
            except _Return:
                pass
            res = []
            for o in outs[:max(1, nargout)]:
                if o not in env:
                    if len(res) >= nargout:
                        break
                    raise MatlabError("output %s of %s not assigned" % (o, name))
                res.append(env[o])
            return res
        if name in self.builtins:
            out = self.builtins[name](self, args, nargout)
            return list(out) if isinstance(out, (list, tuple)) else [out]
        raise MatlabError("unknown function %r" % name)

    # ---- statements ----
    def exec_block(self, stmts, env):
        for s in stmts:
            self.exec_stmt(s, env)

    def exec_stmt(self, s, env):
        kind = s[0]
        if kind == "expr":
            e = s[1]
            if e[0] == "id" and e[1] not in env:
                self.call(e[1], [], 0)
            elif e[0] == "index" and e[1][0] == "id" and e[1][1] not in env:
                self.call(e[1][1], self.eval_args(e[2], env, None, None), 0)
            else:
                env["ans"] = self.eval(e, env)
        elif kind == "assign":
            self.assign(s[1], self.eval(s[2], env), env)
        elif kind == "massign":
            lhs, rhs = s[1], s[2]
            if rhs[0] == "index" and rhs[1][0] == "id" and rhs[1][1] not in env:
                vals = self.call(rhs[1][1], self.eval_args(rhs[2], env, None, None), len(lhs))
            elif rhs[0] == "id" and rhs[1] not in env:
                vals = self.call(rhs[1], [], len(lhs))
            else:
                raise MatlabError("multiple assignment needs a function call")
            if len(vals) < len([x for x in lhs]):
                raise MatlabError("not enough outputs")
            for l, v in zip(lhs, vals):
                if l is not None:
                    self.assign(l, v, env)
        elif kind == "if":
            for cond, body in s[1]:
                if truth(self.eval(cond, env)):
                    self.exec_block(body, env)
                    return
            if s[2] is not None:
                self.exec_block(s[2], env)
        elif kind == "for":
            rng = self.eval(s[2], env)
            cols = mat(rng)
            for j in range(cols.shape[1] if cols.size else 0):
                env[s[1]] = cols[:, j:j + 1].copy() if cols.shape[0] > 1 else np.array([[cols[0, j]]], dtype=np.float64)
                try:
                    self.exec_block(s[3], env)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == "while":
            while truth(self.eval(s[1], env)):
                try:
                    self.exec_block(s[2], env)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == "break":
            raise _Break()
        elif kind == "continue":
            raise _Continue()
        elif kind == "return":
            raise _Return()
        else:
            raise MatlabError("statement %r" % kind)

    # ---- assignment ----
    def assign(self, target, value, env):
        if target[0] == "id":
            env[target[1]] = value
            return
        if target[0] in ("index", "cindex"):
            base = target[1]
            cur = self.lookup_for_assign(base, env)
            new = self.assign_into(cur, target[0], target[2], value, env)
            self.assign(base, new, env)
            return
        raise MatlabError("cannot assign to %r" % (target[0],))

    def lookup_for_assign(self, e, env):
        if e[0] == "id":
            return env.get(e[1])
        if e[0] == "cindex":
            c = self.lookup_for_assign(e[1], env)
            if c is None:
                return None
            idx = self.eval_args(e[2], env, c, "cell")
            try:
                return self.cell_get(c, idx)[0]
            except (IndexError, MatlabError):
                return None
        if e[0] == "index":
            return self.eval(e, env)
        raise MatlabError("bad assignment target")

    def assign_into(self, cur, kind, arg_asts, value, env):
        if kind == "cindex":
            c = cur if isinstance(cur, MCell) else MCell((0, 0))
            idx = self.eval_args(arg_asts, env, c, "cell")
            return self.cell_set(c, idx, value)
        a = zeros((0, 0)) if cur is None else cur
        if isinstance(a, MCell):
            raise MatlabError("() assignment into a cell is not supported")
        a = mat(a)
        idx = self.eval_args(arg_asts, env, a, "array")
        v = mat(value)
        return self.array_set(a, idx, v)

    # ---- indexing helpers ----
    def eval_args(self, asts, env, container, ckind):
        out = []
        n = len(asts)
        for pos, a in enumerate(asts):
            if a == ("colon",):
                out.append(("colon",))
                continue
            env2 = env
            if container is not None:
                env2 = dict(env)
                env2["__end__"] = self.end_value(container, pos, n)
            v = self.eval(a, env2)
            if isinstance(v, CSList):
                out.extend(v)
            else:
                out.append(v)
        return out

    @staticmethod
    def end_value(container, pos, n):
        shp = container.a.shape if isinstance(container, MCell) else (mat(container).shape if not isinstance(container, MStr) else (1, len(container)))
        if n == 1:
            return float(int(np.prod(shp)))
        if pos < n - 1:
            return float(shp[pos]) if pos < len(shp) else 1.0
        return float(int(np.prod(shp[pos:]))) if pos < len(shp) else 1.0

    def array_get(self, a, idx):
        if isinstance(a, MStr):
            s = mat(a)
            r = self.array_get(s, idx)
            return MStr("".join(chr(int(c)) for c in r.reshape(-1, order="F")))
        a = mat(a)
        if len(idx) == 1:
            iv, shp = _index_vector(idx[0], a.size)
            flat = a.reshape(-1, order="F")
            if iv.size and iv.max() >= flat.size:
                raise MatlabError("index exceeds array size")
            r = flat[iv]
            if shp is None:                      # A(:) -> column
                return r.reshape(-1, 1, order="F")
            if shp == (-1,):                     # logical mask: column for matrices, orientation of a vector is kept
                return r.reshape((1, -1) if (a.ndim == 2 and a.shape[0] == 1) else (-1, 1), order="F")
            if (len(shp) == 2 and min(shp) == 1) and a.ndim == 2 and min(a.shape) == 1 and a.size != 1:
                return r.reshape((1, -1) if a.shape[0] == 1 else (-1, 1), order="F")   # vector indexed by vector: orientation of A
            return r.reshape(shp, order="F")
        nd = len(idx)
        shape = list(a.shape) + [1] * max(0, nd - a.ndim)
        if nd < len(shape):                      # fewer subscripts than dimensions: the last one runs over the rest
            shape = shape[:nd - 1] + [int(np.prod(shape[nd - 1:]))]
        b = a.reshape(shape, order="F")
        ivs = []
        for k, ix in enumerate(idx):
            iv, _ = _index_vector(ix, shape[k])
            if iv.size and iv.max() >= shape[k]:
                raise MatlabError("index exceeds array size")
            ivs.append(iv)
        r = b[np.ix_(*ivs)]
        return squeeze_trailing(np.asfortranarray(r))

    def array_set(self, a, idx, v):
        if len(idx) == 1:
            iv, shp = _index_vector(idx[0], a.size)
            need = int(iv.max()) + 1 if iv.size else 0
            if need > a.size:
                if a.ndim == 2 and a.shape[0] <= 1:
                    b = zeros((1, need)); b[0, :a.size] = a.reshape(-1, order="F"); a = b
                elif a.ndim == 2 and a.shape[1] == 1:
                    b = zeros((need, 1)); b[:a.size, 0] = a.reshape(-1, order="F"); a = b
                else:
                    raise MatlabError("cannot grow a matrix with a linear index")
            out = np.array(a, dtype=np.bool_ if (a.dtype == np.bool_ and v.dtype == np.bool_) else np.float64, order="F", copy=True)
            flat = out.reshape(-1, order="F")
            vals = v.reshape(-1, order="F")
            if vals.size == 1:
                flat[iv] = vals[0]
            elif vals.size == iv.size:
                flat[iv] = vals
            else:
                raise MatlabError("assignment size mismatch (%d values into %d places)" % (vals.size, iv.size))
            return flat.reshape(out.shape, order="F")
        nd = len(idx)
        shape = list(a.shape) + [1] * max(0, nd - a.ndim)
        if nd < len(shape):
            raise MatlabError("assignment with fewer subscripts than dimensions is not supported")
        ivs, newshape = [], list(shape)
        vshape = list(v.shape)
        free = 0
        for k, ix in enumerate(idx):
            if isinstance(ix, tuple) and shape[k] == 0 and a.size == 0:
                # A(:, j) = v on an empty A: the colon takes the extent of v's matching dimension
                ext = [s_ for s_ in vshape if s_ != 1]
                iv = np.arange(ext[free] if free < len(ext) else 1)
                free += 1
            else:
                iv, _ = _index_vector(ix, shape[k])
            ivs.append(iv)
            if iv.size:
                newshape[k] = max(newshape[k], int(iv.max()) + 1)
        out = zeros(newshape)
        if a.size:
            out[tuple(slice(0, s_) for s_ in shape)] = a.reshape(shape, order="F")
        target = tuple(len(iv) for iv in ivs)
        if v.size == 1:
            out[np.ix_(*ivs)] = v.reshape(-1)[0]
        else:
            if [s_ for s_ in v.shape if s_ != 1] != [s_ for s_ in target if s_ != 1]:
                raise MatlabError("assignment size mismatch %s into %s" % (v.shape, target))
            out[np.ix_(*ivs)] = v.reshape(target, order="F")
        return squeeze_trailing(out)

    def cell_get(self, c, idx):
        if not isinstance(c, MCell):
            raise MatlabError("{} on a non-cell")
        if len(idx) == 1:
            iv, _ = _index_vector(idx[0], c.a.size)
            flat = c.a.reshape(-1, order="F")
            return [flat[i] for i in iv]
        ivs = [_index_vector(ix, c.a.shape[k])[0] for k, ix in enumerate(idx)]
        return [c.a[i, j] for j in ivs[1] for i in ivs[0]]

    def cell_set(self, c, idx, value):
        if len(idx) == 1:
            iv, _ = _index_vector(idx[0], c.a.size)
            if iv.size != 1:
                raise MatlabError("cell assignment needs one element")
            i = int(iv[0])
            shape = c.a.shape
            if i >= c.a.size:                     # grow a vector cell (a column stays a column, anything else becomes a row)
                shape = (i + 1, 1) if (c.a.shape[1] == 1 and c.a.shape[0] > 1) else (1, i + 1)
            out = MCell(shape)
            f = out.a.reshape(-1, order="F")
            f[:c.a.size] = c.a.reshape(-1, order="F")
            f[i] = value
            out.a = f.reshape(shape, order="F")
            return out
        i, j = (int(_index_vector(ix, c.a.shape[k])[0][0]) for k, ix in enumerate(idx))
        shape = (max(c.a.shape[0], i + 1), max(c.a.shape[1], j + 1))
        out = MCell(shape)
        out.a[:c.a.shape[0], :c.a.shape[1]] = c.a
        out.a[i, j] = value
        return out

    # ---- expressions ----
    def eval(self, e, env):
        k = e[0]
        if k == "num":
            return np.array([[e[1]]])
        if k == "str":
            return MStr(e[1])
        if k == "paren":
            v = self.eval(e[1], env)
            return v[0] if isinstance(v, CSList) else v
        if k == "endidx":
            return np.array([[env["__end__"]]])
        if k == "id":
            name = e[1]
            if name in env:
                return env[name]
            if name == "nargin":
                return np.array([[float(env["__nargin__"])]])
            if name == "nargout":
                return np.array([[float(env["__nargout__"])]])
            if name == "pi":
                return np.array([[math.pi]])
            vals = self.call(name, [], 1)
            return vals[0] if vals else zeros((0, 0))
        if k == "anon":
            return MFunc(e[1], e[2], dict(env))
        if k == "fhandle":
            name = e[1]
            return MFunc(None, name, None)
        if k == "index":
            base = e[1]
            if base[0] == "id" and base[1] not in env:
                args = self.eval_args(e[2], env, None, None)
                vals = self.call(base[1], args, 1)
                return vals[0] if vals else zeros((0, 0))
            b = self.eval(base, env)
            if isinstance(b, MFunc) or (callable(b) and not isinstance(b, (np.ndarray, MCell, str, dict))):
                return self.call_handle(b, self.eval_args(e[2], env, None, None))
            if isinstance(b, MCell):
                idx = self.eval_args(e[2], env, b, "cell")
                items = self.cell_get(b, idx)
                out = MCell((1, len(items)))
                for i, it in enumerate(items):
                    out.a[0, i] = it
                return out
            idx = self.eval_args(e[2], env, b, "array")
            return self.array_get(b, idx)
        if k == "cindex":
            b = self.eval(e[1], env)
            idx = self.eval_args(e[2], env, b, "cell")
            items = self.cell_get(b, idx)
            if len(items) == 1 and not (len(e[2]) >= 1 and any(a == ("colon",) for a in e[2])):
                return items[0]
            return CSList(items)
        if k == "field":
            st = self.eval(e[1], env)
            if not isinstance(st, dict) or e[2] not in st:
                raise MatlabError("no field %r" % (e[2],))
            return st[e[2]]
        if k == "transpose":
            v = self.eval(e[1], env)
            a = mat(v)
            if a.ndim != 2:
                raise MatlabError("transpose of an N-d array")
            return np.asfortranarray(a.T)
        if k == "un":
            v = mat(self.eval(e[2], env))
            if e[1] == "-":
                return -v.astype(np.float64)
            if e[1] == "+":
                return v.astype(np.float64)
            return v == 0
        if k == "range":
            a = scalar(self.eval(e[1], env))
            b = scalar(self.eval(e[3], env))
            s = 1.0 if e[2] is None else scalar(self.eval(e[2], env))
            if s == 0 or (s > 0 and a > b) or (s < 0 and a < b):
                return zeros((1, 0))
            n = int(math.floor((b - a) / s * (1 + 4e-16))) + 1
            return (a + s * np.arange(n, dtype=np.float64)).reshape(1, n)
        if k == "matrix":
            rows = []
            for r in e[1]:
                items = []
                for x in r:
                    v = self.eval(x, env)
                    items.extend(v if isinstance(v, CSList) else [v])
                if all(isinstance(i, MStr) for i in items):
                    rows.append(MStr("".join(items)))
                    continue
                arrs = [mat(i) for i in items if mat(i).size > 0 or len(items) == 1]
                if not arrs:
                    continue
                nd = max(x.ndim for x in arrs)
                arrs = [x.reshape(x.shape + (1,) * (nd - x.ndim), order="F") for x in arrs]
                rows.append(np.concatenate(arrs, axis=1))
            if not rows:
                return zeros((0, 0))
            if len(rows) == 1:
                return rows[0]
            rows = [mat(r) for r in rows]
            return np.asfortranarray(np.concatenate(rows, axis=0))
        if k == "cellrow":
            c = MCell((1, len(e[1])))
            for i, x in enumerate(e[1]):
                c.a[0, i] = self.eval(x, env)
            return c
        if k == "bin":
            op = e[1]
            if op == "||":
                return np.array([[truth(self.eval(e[2], env)) or truth(self.eval(e[3], env))]])
            if op == "&&":
                return np.array([[truth(self.eval(e[2], env)) and truth(self.eval(e[3], env))]])
            a, b = mat(self.eval(e[2], env)), mat(self.eval(e[3], env))
            return self.binop(op, a, b)
        raise MatlabError("expression %r" % (k,))

    def call_handle(self, h, args):
        if not isinstance(h, MFunc):                 # a Python callable handed in as a function handle
            return h(*args)
        if h.params is None:
            return self.call(h.body, args, 1)[0]
        env = dict(h.env)
        for p_, a_ in zip(h.params, args):
            env[p_] = a_
        return self.eval(h.body, env)

    @staticmethod
    def _bcast(a, b):
        nd = max(a.ndim, b.ndim)
        a = a.reshape(a.shape + (1,) * (nd - a.ndim), order="F")
        b = b.reshape(b.shape + (1,) * (nd - b.ndim), order="F")
        for x, y in zip(a.shape, b.shape):
            if x != y and x != 1 and y != 1:
                raise MatlabError("array sizes do not match: %s and %s" % (a.shape, b.shape))
        return a, b

    def binop(self, op, a, b):
        if op in ("*", "/", "^", "\\") and (a.size == 1 or b.size == 1):
            op = {"*": ".*", "/": "./", "^": ".^", "\\": ".\\"}[op]
        if op == "*":
            if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[0]:
                raise MatlabError("inner matrix dimensions must agree: %s * %s" % (a.shape, b.shape))
            return np.asfortranarray(a.astype(np.float64) @ b.astype(np.float64))
        if op in ("/", "^", "\\"):
            raise MatlabError("matrix %s is not supported" % op)
        a, b = self._bcast(a, b)
        if op in ("<", "<=", ">", ">=", "==", "~="):
            f = {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal, "==": np.equal, "~=": np.not_equal}[op]
            return squeeze_trailing(np.asfortranarray(f(a.astype(np.float64), b.astype(np.float64))))
        if op in ("&", "|"):
            f = np.logical_and if op == "&" else np.logical_or
            return squeeze_trailing(np.asfortranarray(f(a != 0, b != 0)))
        a = a.astype(np.float64)
        b = b.astype(np.float64)
        with np.errstate(all="ignore"):
            if op == "+":
                r = a + b
            elif op == "-":
                r = a - b
            elif op == ".*":
                r = a * b
            elif op == "./":
                r = a / b
            elif op == ".\\":
                r = b / a
            elif op == ".^":
                r = np.power(a, b)
            else:
                raise MatlabError("operator %r" % op)
        return squeeze_trailing(np.asfortranarray(r))

    builtins = {}


# ------------------------------------------------------------------------------------------------------------------
# builtins
# ------------------------------------------------------------------------------------------------------------------
def _dims(args):
    d = [int(scalar(a)) for a in args] if len(args) != 1 or mat(args[0]).size == 1 else [int(x) for x in mat(args[0]).reshape(-1)]
    if len(d) == 1:
        d = [d[0], d[0]]
    return tuple(max(0, x) for x in d)


def _first_dim(a):
    for k, s in enumerate(a.shape):
        if s != 1:
            return k
    return 0


def _reduce(f, a, dim, keep=True):
    r = f(a, axis=dim)
    return squeeze_trailing(np.asfortranarray(np.expand_dims(r, dim)))


def _elementwise(f):
    def g(ip, args, nargout):
        with np.errstate(all="ignore"):
            return np.asfortranarray(f(mat(args[0]).astype(np.float64)))
    return g


def _b_sum(ip, args, nargout):
    a = mat(args[0]).astype(np.float64)
    if a.size == 0:
        return np.array([[0.0]])
    dim = int(scalar(args[1])) - 1 if len(args) > 1 else _first_dim(a)
    return _reduce(np.sum, a, dim)


def _b_mean(ip, args, nargout):
    a = mat(args[0]).astype(np.float64)
    dim = int(scalar(args[1])) - 1 if len(args) > 1 else _first_dim(a)
    return _reduce(np.sum, a, dim) / float(a.shape[dim])          # Matlab: sum(x) / n


def _b_cumsum(ip, args, nargout):
    a = mat(args[0]).astype(np.float64)
    dim = int(scalar(args[1])) - 1 if len(args) > 1 else _first_dim(a)
    return np.asfortranarray(np.cumsum(a, axis=dim))


def _b_minmax(f, fe):
    def g(ip, args, nargout):
        a = mat(args[0]).astype(np.float64)
        if len(args) >= 2 and mat(args[1]).size > 0:
            x, y = Interp._bcast(a, mat(args[1]).astype(np.float64))
            return squeeze_trailing(np.asfortranarray(fe(x, y)))
        dim = int(scalar(args[2])) - 1 if len(args) > 2 else _first_dim(a)
        return _reduce(f, a, dim)
    return g


def _b_size(ip, args, nargout):
    x = args[0]
    shp = x.a.shape if isinstance(x, MCell) else ((1, len(x)) if isinstance(x, MStr) else mat(x).shape)
    if len(args) > 1:
        k = int(scalar(args[1])) - 1
        return np.array([[float(shp[k]) if k < len(shp) else 1.0]])
    if nargout <= 1:
        return np.array([[float(s) for s in shp]])
    out = [float(s) for s in shp[:nargout - 1]] + [float(int(np.prod(shp[nargout - 1:])))]
    return [np.array([[o]]) for o in out]


def _numel(x):
    return x.a.size if isinstance(x, MCell) else (len(x) if isinstance(x, MStr) else mat(x).size)


def _b_length(ip, args, nargout):
    x = args[0]
    shp = x.a.shape if isinstance(x, MCell) else ((1, len(x)) if isinstance(x, MStr) else mat(x).shape)
    return np.array([[float(0 if 0 in shp else max(shp))]])


def _b_repmat(ip, args, nargout):
    a = mat(args[0])
    reps = _dims(args[1:])
    nd = max(a.ndim, len(reps))
    a = a.reshape(a.shape + (1,) * (nd - a.ndim), order="F")
    reps = tuple(reps) + (1,) * (nd - len(reps))
    return squeeze_trailing(np.asfortranarray(np.tile(a, reps)))


def _b_reshape(ip, args, nargout):
    a = mat(args[0])
    if len(args) == 2:
        dims = [int(x) for x in mat(args[1]).reshape(-1)]
    else:
        dims = [None if mat(x).size == 0 else int(scalar(x)) for x in args[1:]]
        known = int(np.prod([d_ for d_ in dims if d_ is not None])) if any(d_ is not None for d_ in dims) else 1
        dims = [(a.size // known if known else 0) if d_ is None else d_ for d_ in dims]
    if int(np.prod(dims)) != a.size:
        raise MatlabError("reshape: number of elements must not change (%s -> %s)" % (a.shape, dims))
    if len(dims) == 1:
        dims = [dims[0], 1]
    return squeeze_trailing(np.asfortranarray(a.reshape(dims, order="F")))


def _b_permute(ip, args, nargout):
    a = mat(args[0])
    order = [int(x) - 1 for x in mat(args[1]).reshape(-1)]
    a = a.reshape(a.shape + (1,) * (len(order) - a.ndim), order="F")
    return squeeze_trailing(np.asfortranarray(np.transpose(a, order)))


def _b_qr(ip, args, nargout):
    a = mat(args[0]).astype(np.float64)
    econ = len(args) > 1
    q, r = np.linalg.qr(a, mode="reduced" if econ else "complete")     # LAPACK dgeqrf / dorgqr, as Matlab's qr
    if nargout <= 1:
        return np.asfortranarray(r)
    return [np.asfortranarray(q), np.asfortranarray(r)]


def _b_find(ip, args, nargout):
    a = mat(args[0])
    iv = np.flatnonzero(a.reshape(-1, order="F") != 0).astype(np.float64) + 1.0
    return iv.reshape((1, -1) if (a.ndim == 2 and a.shape[0] == 1 and a.shape[1] != 1) else (-1, 1), order="F")


def _b_spdiags(ip, args, nargout):
    b, d = mat(args[0]).astype(np.float64), [int(x) for x in mat(args[1]).reshape(-1)]
    m, n = int(scalar(args[2])), int(scalar(args[3]))
    out = zeros((m, n))
    for c, k in enumerate(d):                  # Matlab: for m >= n diagonal k takes B(j, c) at column j (super-diagonals use the upper rows' columns)
        for j in range(n):
            i = j - k
            if 0 <= i < m:
                out[i, j] = b[j if m >= n else i, c]
    return out


def _b_cellfun(ip, args, nargout):
    h, c = args[0], args[1]
    vals = [scalar(ip.call_handle(h, [x])) for x in c.a.reshape(-1, order="F")]
    return np.array(vals, dtype=np.float64).reshape(c.a.shape, order="F")


def _b_isa(ip, args, nargout):
    x, cls = args[0], str(args[1])
    if isinstance(x, dict):                      # a struct, or an object of a class the caller names in the field __class__
        return np.array([[x.get("__class__", "struct") == cls]])
    actual = "cell" if isinstance(x, MCell) else ("char" if isinstance(x, MStr) else ("function_handle" if isinstance(x, MFunc) else ("logical" if mat(x).dtype == np.bool_ else "double")))
    return np.array([[actual == cls or (cls in ("numeric", "float") and actual == "double")]])


def _b_rand(ip, args, nargout):
    if ip.rand_stream is None:
        raise MatlabError("rand called but no stream was supplied")
    shape = _dims(args) if args else (1, 1)
    return np.asfortranarray(np.asarray(ip.rand_stream(shape), dtype=np.float64).reshape(shape, order="F"))


def _b_load(ip, args, nargout):
    name = str(args[0])
    if name not in ip.files:
        raise MatlabError("load(%r): no such table supplied" % name)
    return np.asfortranarray(np.asarray(ip.files[name], dtype=np.float64))


def _b_str2double(ip, args, nargout):
    try:
        return np.array([[float(str(args[0]))]])
    except ValueError:
        return np.array([[float("nan")]])


def _b_fprintf(ip, args, nargout):
    ip.printed.append(args)
    return []


def _b_error(ip, args, nargout):
    raise MatlabError("error(): %s" % (args[0] if args else ""))


def _b_keyboard(ip, args, nargout):
    raise MatlabError("keyboard reached (the reference stops in the debugger here)")


def _b_cell2mat(ip, args, nargout):
    c = args[0]
    if c.a.shape[1] == 1:
        return np.asfortranarray(np.concatenate([mat(x) for x in c.a[:, 0]], axis=0))
    if c.a.shape[0] == 1:
        return np.asfortranarray(np.concatenate([mat(x) for x in c.a[0, :]], axis=1))
    return np.asfortranarray(np.concatenate([np.concatenate([mat(x) for x in row], axis=1) for row in c.a], axis=0))


def _special():
    from scipy import special
    return special


Interp.builtins = {
    "exp": _elementwise(np.exp), "log": _elementwise(np.log), "sqrt": _elementwise(np.sqrt), "abs": _elementwise(np.abs),
    "floor": _elementwise(np.floor), "ceil": _elementwise(np.ceil),
    "erf": _elementwise(lambda x: _special().erf(x)), "erfinv": _elementwise(lambda x: _special().erfinv(x)),
    "isnan": lambda ip, a, n: np.isnan(mat(a[0]).astype(np.float64)),
    "sum": _b_sum, "mean": _b_mean, "cumsum": _b_cumsum, "max": _b_minmax(np.max, np.maximum), "min": _b_minmax(np.min, np.minimum),
    "any": lambda ip, a, n: np.array([[bool(np.any(mat(a[0]) != 0))]]) if min(mat(a[0]).shape) <= 1 else _reduce(np.any, mat(a[0]) != 0, 0),
    "numel": lambda ip, a, n: np.array([[float(_numel(a[0]))]]), "length": _b_length, "size": _b_size,
    "isempty": lambda ip, a, n: np.array([[_numel(a[0]) == 0]]),
    "isscalar": lambda ip, a, n: np.array([[_numel(a[0]) == 1]]),
    "zeros": lambda ip, a, n: zeros(_dims(a) if a else (1, 1)), "ones": lambda ip, a, n: zeros(_dims(a) if a else (1, 1)) + 1.0,
    "cell": lambda ip, a, n: MCell(_dims(a)),
    "repmat": _b_repmat, "reshape": _b_reshape, "permute": _b_permute, "qr": _b_qr, "find": _b_find, "spdiags": _b_spdiags,
    "cellfun": _b_cellfun, "isa": _b_isa, "rand": _b_rand, "load": _b_load,
    "lower": lambda ip, a, n: MStr(str(a[0]).lower()), "double": lambda ip, a, n: mat(a[0]).astype(np.float64),
    "char": lambda ip, a, n: MStr("".join(chr(int(c)) for c in mat(a[0]).reshape(-1, order="F"))),
    "true": lambda ip, a, n: np.array([[True]]), "false": lambda ip, a, n: np.array([[False]]),
    "tic": lambda ip, a, n: np.array([[0.0]]), "toc": lambda ip, a, n: np.array([[0.0]]), "nan": lambda ip, a, n: np.array([[float("nan")]]),
    "strcmpi": lambda ip, a, n: np.array([[isinstance(a[0], str) and isinstance(a[1], str) and str(a[0]).lower() == str(a[1]).lower()]]),
    "cell2mat": _b_cell2mat,
    "strcmp": lambda ip, a, n: np.array([[isinstance(a[0], str) and isinstance(a[1], str) and str(a[0]) == str(a[1])]]),
    "warning": _b_fprintf,
    "str2double": _b_str2double, "fprintf": _b_fprintf, "error": _b_error, "keyboard": _b_keyboard,
}
