"""Evaluate the `cparse` AST of a GLSL ES 3.00 fragment shader with strict fp32 semantics.

TEST INFRASTRUCTURE ONLY -- see cparse.py.  Every `highp float` operation is one correctly
rounded IEEE binary32 operation (NumPy float32 scalars), evaluated in source order with no
contraction -- the strictest reading of the shader, and the one the CPU oracle restates.  Real
GPUs may deviate (GLSL ES does not require correctly rounded division, and allows FMA
contraction); that caveat is the oracle's, stated in DESIGN.md.

Textures: NEAREST filtering, CLAMP_TO_EDGE (HTML:438-458).  A fragment at pixel (x, y) receives
vUV = ((x+0.5)/W, (y+0.5)/H) rounded to fp32 (VS_SRC draws one full-screen triangle, HTML:213-220).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from .cparse import parse

f32 = np.float32
_libc = ctypes.CDLL(None)
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


def lit(text: str):
    """Decimal literal -> fp32, rounded once (like a C/GLSL compiler), not via float64."""
    return f32(_libc.strtof(text.encode(), None))


class Vec:
    __slots__ = ("v", "kind")

    def __init__(self, v, kind="f"):
        self.v = list(v)
        self.kind = kind

    def __len__(self):
        return len(self.v)


class Sampler:
    def __init__(self, arr):
        self.a = arr              # [H][W][C] float32
        self.h, self.w = arr.shape[:2]

    def fetch(self, uv):
        ix = int(math.floor(float(uv.v[0]) * self.w))
        iy = int(math.floor(float(uv.v[1]) * self.h))
        ix = min(max(ix, 0), self.w - 1)
        iy = min(max(iy, 0), self.h - 1)
        t = self.a[iy, ix]
        if t.shape[0] == 1:
            return Vec([f32(t[0]), f32(0), f32(0), f32(1)])
        return Vec([f32(t[0]), f32(t[1]), f32(t[2]), f32(t[3])])


class GReturn(Exception):
    def __init__(self, v):
        self.v = v


SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}


def _arith(op, a, b):
    if op == "+":
        return a + b
    if op == "-":
        return a - b
    if op == "*":
        return a * b
    if op == "/":
        if isinstance(a, int):
            return int(a / b)
        return a / b
    raise SyntaxError(op)


def binop(op, a, b):
    if isinstance(a, Vec) or isinstance(b, Vec):
        n = len(a) if isinstance(a, Vec) else len(b)
        av = a.v if isinstance(a, Vec) else [a] * n
        bv = b.v if isinstance(b, Vec) else [b] * n
        kind = a.kind if isinstance(a, Vec) else b.kind
        return Vec([_arith(op, x, y) for x, y in zip(av, bv)], kind)
    if type(a) is not type(b) and not (isinstance(a, (int, bool)) and isinstance(b, (int, bool))):
        raise TypeError(f"GLSL: no implicit conversion in {type(a).__name__} {op} {type(b).__name__}")
    return _arith(op, a, b)


class Shader:
    def __init__(self, src: str):
        self.ast = parse(src, "glsl")
        self.funcs = {}
        self.globals = {}
        self.uniform_names = []
        self.out_names = []
        self.frame_globals = {}
        with np.errstate(all="ignore"):
            for st in self.ast[1]:
                if st[0] == "gfunc":
                    self.funcs[st[2]] = st
                elif st[0] == "gdecl":
                    quals, typ, decls = st[1], st[2], st[3]
                    for name, size, init in decls:
                        if "uniform" in quals or "in" in quals:
                            self.uniform_names.append(name)
                        elif "out" in quals:
                            self.out_names.append(name)
                        elif init is not None:
                            self.globals[name] = self.ev(init, {})

    # -- running --------------------------------------------------------------------------------
    def run_fragment(self, inputs: dict) -> dict:
        """inputs: uniforms and `in` variables.  Returns the `out` variables after main()."""
        env = dict(inputs)
        self.frame_globals = env
        with np.errstate(all="ignore"):
            self.call("main", [])
        return {k: env.get(k) for k in self.out_names}

    def call(self, name, args):
        fn = self.funcs[name]
        local = {}
        for (ptype, pname), a in zip(fn[3], args):
            local[pname] = a
        try:
            self.exec_block(fn[4], local)
        except GReturn as r:
            return r.v
        return None

    # -- variable access: locals, then per-fragment globals (uniforms/outs), then constants --
    def lookup(self, name, local):
        if name in local:
            return local[name]
        if name in self.frame_globals:
            return self.frame_globals[name]
        if name in self.globals:
            return self.globals[name]
        raise NameError(f"GLSL: {name}")

    def assign_var(self, name, val, local):
        if name in local:
            local[name] = val
        elif name in self.out_names or name in self.frame_globals:
            self.frame_globals[name] = val
        else:
            raise NameError(f"GLSL: assignment to undeclared {name}")

    # -- statements -----------------------------------------------------------------------------
    def exec_block(self, node, local):
        for st in node[1]:
            self.ex(st, local)

    def ex(self, st, local):
        k = st[0]
        if k == "expr":
            self.ev(st[1], local)
        elif k == "gdecl":
            typ = st[2]
            for name, size, init in st[3]:
                if size is not None:
                    n = self.ev(size, local)
                    local[name] = [self.zero(typ) for _ in range(n)]
                elif init is not None:
                    local[name] = self.ev(init, local)
                else:
                    local[name] = self.zero(typ)
        elif k == "block":
            self.exec_block(st, local)
        elif k == "if":
            if self.ev(st[1], local):
                self.ex(st[2], local)
            elif st[3] is not None:
                self.ex(st[3], local)
        elif k == "for":
            if st[1] is not None:
                self.ex(st[1], local)
            while st[2] is None or self.ev(st[2], local):
                self.ex(st[4], local)
                if st[3] is not None:
                    self.ev(st[3], local)
        elif k == "return":
            raise GReturn(self.ev(st[1], local) if st[1] is not None else None)
        elif k == "empty":
            pass
        else:
            raise SyntaxError(f"GLSL: statement {k}")

    @staticmethod
    def zero(typ):
        if typ == "float":
            return f32(0)
        if typ == "int":
            return 0
        if typ == "bool":
            return False
        if typ in ("vec2", "vec3", "vec4"):
            return Vec([f32(0)] * int(typ[3]))
        raise SyntaxError(f"GLSL: zero of {typ}")

    # -- expressions ----------------------------------------------------------------------------
    def ev(self, e, local):
        k = e[0]
        if k == "num":
            return lit(e[1]) if e[2] else int(e[1])
        if k == "id":
            if e[1] == "true":
                return True
            if e[1] == "false":
                return False
            return self.lookup(e[1], local)
        if k == "bin":
            op = e[1]
            if op == "&&":
                return bool(self.ev(e[2], local)) and bool(self.ev(e[3], local))
            if op == "||":
                return bool(self.ev(e[2], local)) or bool(self.ev(e[3], local))
            a = self.ev(e[2], local)
            b = self.ev(e[3], local)
            if op in ("+", "-", "*", "/"):
                return binop(op, a, b)
            if type(a) is not type(b) and not (isinstance(a, int) and isinstance(b, int)):
                raise TypeError(f"GLSL: comparing {type(a).__name__} with {type(b).__name__}")
            if op == "<":
                return bool(a < b)
            if op == ">":
                return bool(a > b)
            if op == "<=":
                return bool(a <= b)
            if op == ">=":
                return bool(a >= b)
            if op == "==":
                return bool(a == b)
            if op == "!=":
                return bool(a != b)
            raise SyntaxError(f"GLSL: operator {op}")
        if k == "un":
            v = self.ev(e[2], local)
            if e[1] == "-":
                return Vec([-x for x in v.v], v.kind) if isinstance(v, Vec) else -v
            if e[1] == "!":
                return not v
            return v
        if k == "cond":
            return self.ev(e[2], local) if self.ev(e[1], local) else self.ev(e[3], local)
        if k == "assign":
            op, target = e[1], e[2]
            val = self.ev(e[3], local)
            if op != "=":
                val = binop(op[0], self.ev(target, local), val)
            self.store(target, val, local)
            return val
        if k in ("postinc", "preinc"):
            old = self.ev(e[2], local)
            new = old + 1 if e[1] == "++" else old - 1
            self.store(e[2], new, local)
            return old if k == "postinc" else new
        if k == "index":
            return self.ev(e[1], local)[self.ev(e[2], local)]
        if k == "member":
            obj = self.ev(e[1], local)
            if len(e[2]) != 1:
                raise SyntaxError("GLSL: multi-component swizzle")
            return obj.v[SWZ[e[2]]]
        if k == "call":
            return self.call_expr(e, local)
        raise SyntaxError(f"GLSL: expression {k}")

    def store(self, target, val, local):
        if target[0] == "id":
            self.assign_var(target[1], val, local)
        elif target[0] == "index":
            self.ev(target[1], local)[self.ev(target[2], local)] = val
        else:
            raise SyntaxError("GLSL: bad assignment target")

    def call_expr(self, e, local):
        name = e[1][1]
        args = [self.ev(a, local) for a in e[2]]
        if name in self.funcs:
            return self.call(name, args)
        if name in ("vec2", "vec3", "vec4"):
            n = int(name[3])
            flat = []
            for a in args:
                if isinstance(a, Vec):
                    flat.extend(f32(x) for x in a.v)
                else:
                    flat.append(f32(a))
            if len(flat) == 1:
                flat = flat * n
            assert len(flat) == n, name
            return Vec(flat)
        if name == "ivec2":
            a = args[0]
            return Vec([int(a.v[0]), int(a.v[1])], "i")       # float -> int truncates toward zero
        if name == "float":
            return f32(args[0])
        if name == "int":
            return int(args[0])
        if name == "texture":
            return args[0].fetch(args[1])
        if name == "clamp":
            return min(max(args[0], args[1]), args[2])
        if name == "sqrt":
            return np.sqrt(args[0])
        if name == "floor":
            return f32(math.floor(float(args[0])))
        if name == "max":
            return max(args[0], args[1])
        if name == "min":
            return min(args[0], args[1])
        if name == "length":
            v = args[0].v
            s = v[0] * v[0]
            for x in v[1:]:
                s = s + x * x
            return np.sqrt(s)
        if name == "mix":
            x, y, a = args
            one = f32(1)
            return Vec([xi * (one - a) + yi * a for xi, yi in zip(x.v, y.v)])
        raise SyntaxError(f"GLSL: function {name}")


def run_pass(shader: Shader, w: int, h: int, uniforms: dict):
    """Draw the full-screen triangle: one fragment per pixel.  Returns {out_name: [h][w][4] fp32}."""
    outs = {name: np.zeros((h, w, 4), np.float32) for name in shader.out_names}
    for y in range(h):
        for x in range(w):
            inputs = dict(uniforms)
            inputs["vUV"] = Vec([f32((x + 0.5) / w), f32((y + 0.5) / h)])
            res = shader.run_fragment(inputs)
            for name, v in res.items():
                outs[name][y, x] = [float(c) for c in v.v]
    return outs
