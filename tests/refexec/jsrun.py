"""Evaluate the `cparse` AST with JavaScript semantics (numbers are IEEE float64).

TEST INFRASTRUCTURE ONLY -- see cparse.py.  Supports exactly what the reference's geometry,
initialisation, statistics and force functions use.  Math.* maps to Python's math (glibc), the
same stand-in for V8's fdlibm that the oracle documents.
"""
from __future__ import annotations

import functools
import math

import numpy as np

from .cparse import parse


class JSReturn(Exception):
    def __init__(self, v):
        self.v = v


class JSBreak(Exception):
    pass


class JSContinue(Exception):
    pass


UNDEF = object()


class Env:
    def __init__(self, parent=None):
        self.v = {}
        self.parent = parent

    def find(self, name):
        e = self
        while e is not None:
            if name in e.v:
                return e
            e = e.parent
        return None

    def get(self, name):
        e = self.find(name)
        if e is None:
            raise NameError(f"JS: {name} is not defined")
        return e.v[name]

    def set(self, name, val):
        e = self.find(name)
        if e is None:
            raise NameError(f"JS: assignment to undeclared {name}")
        e.v[name] = val

    def declare(self, name, val):
        self.v[name] = val


class TypedArray:
    """Uint8Array / Float32Array: element stores convert like the real thing."""

    def __init__(self, kind, n):
        self.kind = kind
        self.a = np.zeros(int(n), np.uint8 if kind == "Uint8Array" else np.float32)

    def get(self, i):
        v = self.a[int(i)]
        return float(v)

    def set(self, i, v):
        if self.kind == "Uint8Array":
            self.a[int(i)] = int(v) & 0xFF
        else:
            self.a[int(i)] = np.float32(v)


class Function:
    def __init__(self, params, body, env, is_expr, interp):
        self.params, self.body, self.env, self.is_expr, self.interp = params, body, env, is_expr, interp

    def __call__(self, *args):
        env = Env(self.env)
        for k, p in enumerate(self.params):
            self.interp.bind(env, p, args[k] if k < len(args) else UNDEF)
        if self.is_expr:
            return self.interp.ev(self.body, env)
        try:
            self.interp.exec_block(self.body, env)
        except JSReturn as r:
            return r.v
        return UNDEF


def truthy(v):
    if v is None or v is UNDEF or v is False:
        return False
    if isinstance(v, (int, float)):
        return not (v == 0 or v != v)
    if isinstance(v, str):
        return v != ""
    return True


def js_div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


MATH = {
    "PI": math.pi, "cos": math.cos, "sin": math.sin, "sqrt": lambda x: math.sqrt(x) if x >= 0 else math.nan,
    "atan": math.atan, "hypot": math.hypot, "ceil": lambda x: float(math.ceil(x)),
    "floor": lambda x: float(math.floor(x)), "max": lambda *a: float(max(a)), "min": lambda *a: float(min(a)),
    "abs": abs,
    # Math.round: nearest integer, ties towards +infinity -- NOT floor(x + 0.5), whose sum rounds up for
    # 0.49999999999999994 (x - floor(x) is exact)
    "round": lambda x: float(math.floor(x) + (1.0 if x - math.floor(x) >= 0.5 else 0.0)),
}


class Interp:
    def __init__(self):
        self.globals = Env()
        g = self.globals
        g.declare("Math", MATH)
        g.declare("NaN", math.nan)
        g.declare("Infinity", math.inf)
        g.declare("undefined", UNDEF)
        g.declare("isFinite", lambda x: isinstance(x, (int, float)) and math.isfinite(x))

    def run(self, src):
        self.exec_block(parse(src, "js"), self.globals, new_scope=False)

    def call(self, name, *args):
        return self.globals.get(name)(*args)

    # -- binding patterns --------------------------------------------------------------------
    def bind(self, env, pat, val):
        if pat[0] == "name":
            env.declare(pat[1], val)
        elif pat[0] == "arraypat":
            for k, sub in enumerate(pat[1]):
                self.bind(env, sub, val[k])
        elif pat[0] == "objpat":
            for name in pat[1]:
                env.declare(name, val[name])

    # -- statements ----------------------------------------------------------------------------
    def exec_block(self, node, env, new_scope=True):
        scope = Env(env) if new_scope else env
        for st in node[1]:                       # hoist function declarations
            if st[0] == "funcdecl":
                scope.declare(st[1], Function(st[2], st[3], scope, False, self))
        for st in node[1]:
            self.ex(st, scope)

    def ex(self, st, env):
        k = st[0]
        if k == "expr":
            self.ev(st[1], env)
        elif k == "decl":
            for pat, init in st[1]:
                self.bind(env, pat, self.ev(init, env) if init is not None else UNDEF)
        elif k == "funcdecl" or k == "empty":
            pass
        elif k == "block":
            self.exec_block(st, env)
        elif k == "if":
            if truthy(self.ev(st[1], env)):
                self.ex(st[2], env)
            elif st[3] is not None:
                self.ex(st[3], env)
        elif k == "for":
            scope = Env(env)
            if st[1] is not None:
                self.ex(st[1], scope)
            while st[2] is None or truthy(self.ev(st[2], scope)):
                try:
                    self.ex(st[4], Env(scope))
                except JSBreak:
                    break
                except JSContinue:
                    pass
                if st[3] is not None:
                    self.ev(st[3], scope)
        elif k == "while":
            while truthy(self.ev(st[1], env)):
                try:
                    self.ex(st[2], env)
                except JSBreak:
                    break
                except JSContinue:
                    pass
        elif k == "return":
            raise JSReturn(self.ev(st[1], env) if st[1] is not None else UNDEF)
        elif k == "continue":
            raise JSContinue()
        elif k == "break":
            raise JSBreak()
        else:
            raise SyntaxError(f"JS: statement {k}")

    # -- expressions -----------------------------------------------------------------------------
    def ev(self, e, env):
        k = e[0]
        if k == "num":
            return float(e[1])
        if k == "str":
            return e[1]
        if k == "id":
            n = e[1]
            if n == "null":
                return None
            if n == "true":
                return True
            if n == "false":
                return False
            return env.get(n)
        if k == "bin":
            return self.binop(e, env)
        if k == "un":
            v = self.ev(e[2], env)
            if e[1] == "-":
                return -v
            if e[1] == "+":
                return +v
            return not truthy(v)
        if k == "cond":
            return self.ev(e[2], env) if truthy(self.ev(e[1], env)) else self.ev(e[3], env)
        if k == "assign":
            return self.assign(e, env)
        if k in ("postinc", "preinc"):
            old = self.ev(e[2], env)
            new = old + 1 if e[1] == "++" else old - 1
            self.store(e[2], new, env)
            return old if k == "postinc" else new
        if k == "array":
            return [self.ev(x, env) for x in e[1]]
        if k == "object":
            return {key: self.ev(v, env) for key, v in e[1]}
        if k == "arrow":
            return Function(e[1], e[2], env, e[3], self)
        if k == "index":
            obj = self.ev(e[1], env)
            idx = self.ev(e[2], env)
            if isinstance(obj, TypedArray):
                return obj.get(idx)
            if isinstance(obj, dict):
                return obj[idx]
            return obj[int(idx)]
        if k == "member":
            return self.member(self.ev(e[1], env), e[2])
        if k == "call":
            return self.call_expr(e, env)
        if k == "new":
            name = e[1][1]
            args = [self.ev(a, env) for a in e[2]]
            if name in ("Uint8Array", "Float32Array"):
                return TypedArray(name, args[0])
            raise SyntaxError(f"JS: new {name}")
        if k == "typeof":
            try:
                v = self.ev(e[1], env)
            except NameError:
                return "undefined"
            return "undefined" if v is UNDEF else "number" if isinstance(v, (int, float)) else "object"
        if k == "comma":
            self.ev(e[1], env)
            return self.ev(e[2], env)
        raise SyntaxError(f"JS: expression {k}")

    def binop(self, e, env):
        op = e[1]
        if op == "&&":
            a = self.ev(e[2], env)
            return self.ev(e[3], env) if truthy(a) else a
        if op == "||":
            a = self.ev(e[2], env)
            return a if truthy(a) else self.ev(e[3], env)
        a = self.ev(e[2], env)
        b = self.ev(e[3], env)
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            return js_div(a, b)
        if op == "%":
            return math.fmod(a, b)
        if op == "**":
            return math.pow(a, b)
        if op == "<":
            return a < b
        if op == ">":
            return a > b
        if op == "<=":
            return a <= b
        if op == ">=":
            return a >= b
        if op in ("===", "=="):
            return self.equal(a, b)
        if op in ("!==", "!="):
            return not self.equal(a, b)
        raise SyntaxError(f"JS: operator {op}")

    @staticmethod
    def equal(a, b):
        if a is None or b is None or a is UNDEF or b is UNDEF:
            return a is b
        if isinstance(a, bool) or isinstance(b, bool):
            return a is b if isinstance(a, bool) and isinstance(b, bool) else float(a) == float(b)
        return a == b

    def store(self, target, val, env):
        if target[0] == "id":
            env.set(target[1], val)
        elif target[0] == "index":
            obj = self.ev(target[1], env)
            idx = self.ev(target[2], env)
            if isinstance(obj, TypedArray):
                obj.set(idx, val)
            elif isinstance(obj, dict):
                obj[idx] = val
            else:
                i = int(idx)
                while len(obj) <= i:
                    obj.append(UNDEF)
                obj[i] = val
        elif target[0] == "member":
            self.ev(target[1], env)[target[2]] = val
        else:
            raise SyntaxError("JS: bad assignment target")

    def assign(self, e, env):
        op, target = e[1], e[2]
        val = self.ev(e[3], env)
        if op != "=":
            cur = self.ev(target, env)
            val = {"+=": lambda: cur + val, "-=": lambda: cur - val, "*=": lambda: cur * val,
                   "/=": lambda: js_div(cur, val)}[op]()
        self.store(target, val, env)
        return val

    def member(self, obj, name):
        if isinstance(obj, dict):
            return obj[name]
        if name == "length":
            return float(len(obj.a) if isinstance(obj, TypedArray) else len(obj))
        raise SyntaxError(f"JS: member {name}")

    def call_expr(self, e, env):
        callee = e[1]
        args = [self.ev(a, env) for a in e[2]]
        if callee[0] == "member":
            obj = self.ev(callee[1], env)
            name = callee[2]
            if isinstance(obj, dict):
                return obj[name](*args)
            if isinstance(obj, list):
                return self.array_method(obj, name, args)
            raise SyntaxError(f"JS: method {name}")
        fn = self.ev(callee, env)
        return fn(*args)

    @staticmethod
    def array_method(a, name, args):
        if name == "push":
            a.extend(args)
            return float(len(a))
        if name == "map":
            return [args[0](v, float(i)) for i, v in enumerate(a)]
        if name == "reverse":
            a.reverse()
            return a
        if name == "concat":
            return a + list(args[0])
        if name == "slice":
            return a[int(args[0]):] if len(args) == 1 else a[int(args[0]):int(args[1])]
        if name == "sort":
            cmp = args[0]
            a.sort(key=functools.cmp_to_key(lambda x, y: (lambda r: -1 if r < 0 else (1 if r > 0 else 0))(cmp(x, y))))
            return a
        if name == "pop":
            return a.pop()
        raise SyntaxError(f"JS: Array.prototype.{name}")
