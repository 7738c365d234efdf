"""Tokenizer and parser for the C-like subset shared by the reference's JavaScript and GLSL.

TEST INFRASTRUCTURE ONLY.  Purpose: execute the reference's OWN source text
(pages/airfoil_flow_lbm_aerolab.html: the geometry / statistics / force functions in JavaScript
and the step / render fragment shaders in GLSL ES 3.00) so that the CPU oracle can be pinned
against outputs of the reference itself.  No JavaScript engine or browser exists in the build
image, hence this purpose-built interpreter: `cparse` produces a small AST, `jsrun` evaluates it
with JavaScript (float64) semantics, `glslrun` with GLSL highp (fp32) semantics.

Only the constructs that those functions use are supported; anything else raises SyntaxError so
that a silent misreading is impossible.

AST nodes are tuples: (kind, ...).
"""
from __future__ import annotations

import re

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>'(?:[^'\\]|\\.)*'|"(?:[^"\\]|\\.)*")
  | (?P<op>===|!==|\*\*|=>|\+\+|--|&&|\|\||==|!=|<=|>=|\+=|-=|\*=|/=|[-+*/%<>=!?:;,.(){}\[\]])
""", re.X | re.S)

GLSL_TYPES = {"float", "int", "bool", "vec2", "vec3", "vec4", "ivec2", "void", "sampler2D"}
GLSL_QUALIFIERS = {"const", "uniform", "in", "out", "highp", "mediump", "lowp"}


def tokenize(src: str):
    toks = []
    pos = 0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {src[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        toks.append((kind, m.group(kind)))
    toks.append(("eof", ""))
    return toks


class Parser:
    def __init__(self, src: str, lang: str):
        assert lang in ("js", "glsl")
        self.lang = lang
        self.t = tokenize(src)
        self.i = 0

    # -- helpers -----------------------------------------------------------------------------
    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, text, k=0):
        tok = self.t[self.i + k]
        return tok[0] in ("op", "id") and tok[1] == text

    def eat(self, text=None):
        tok = self.t[self.i]
        if text is not None and tok[1] != text:
            raise SyntaxError(f"expected {text!r}, got {tok[1]!r} (token {self.i})")
        self.i += 1
        return tok

    def accept(self, text):
        if self.at(text):
            self.i += 1
            return True
        return False

    # -- program / statements -------------------------------------------------------------------
    def program(self):
        body = []
        while self.peek()[0] != "eof":
            body.append(self.statement())
        return ("block", body)

    def block(self):
        self.eat("{")
        body = []
        while not self.at("}"):
            body.append(self.statement())
        self.eat("}")
        return ("block", body)

    def statement(self):
        tok = self.peek()
        if self.at("{"):
            return self.block()
        if self.at(";"):
            self.eat()
            return ("empty",)
        if tok[0] == "id":
            w = tok[1]
            if w == "function" and self.lang == "js":
                return self.js_function_decl()
            if w in ("const", "let", "var") and self.lang == "js":
                d = self.js_decl()
                self.accept(";")
                return d
            if w == "if":
                return self.if_stmt()
            if w == "for":
                return self.for_stmt()
            if w == "while":
                self.eat()
                self.eat("(")
                c = self.expr()
                self.eat(")")
                return ("while", c, self.statement())
            if w == "return":
                self.eat()
                if self.at(";") or self.at("}"):
                    self.accept(";")
                    return ("return", None)
                e = self.expr()
                self.accept(";")
                return ("return", e)
            if w == "continue":
                self.eat()
                self.accept(";")
                return ("continue",)
            if w == "break":
                self.eat()
                self.accept(";")
                return ("break",)
            if self.lang == "glsl":
                if w == "precision":
                    while not self.at(";"):
                        self.eat()
                    self.eat(";")
                    return ("empty",)
                if w == "layout":
                    self.eat()
                    self.eat("(")
                    while not self.at(")"):
                        self.eat()
                    self.eat(")")
                    return self.statement()
                if w in GLSL_TYPES or w in GLSL_QUALIFIERS:
                    return self.glsl_decl_or_function()
        e = self.expr()
        self.accept(";")
        return ("expr", e)

    def if_stmt(self):
        self.eat("if")
        self.eat("(")
        c = self.expr()
        self.eat(")")
        then = self.statement()
        other = None
        if self.accept("else"):
            other = self.statement()
        return ("if", c, then, other)

    def for_stmt(self):
        self.eat("for")
        self.eat("(")
        init = None
        if not self.at(";"):
            if self.lang == "js" and self.peek()[1] in ("let", "const", "var"):
                init = self.js_decl()
            elif self.lang == "glsl" and self.peek()[1] in GLSL_TYPES:
                init = self.glsl_decl_or_function(in_for=True)
            else:
                init = ("expr", self.expr())
        self.eat(";")
        cond = None if self.at(";") else self.expr()
        self.eat(";")
        step = None if self.at(")") else self.expr()
        self.eat(")")
        return ("for", init, cond, step, self.statement())

    # -- JavaScript declarations -------------------------------------------------------------------
    def js_function_decl(self):
        self.eat("function")
        name = self.eat()[1]
        params = self.js_params()
        return ("funcdecl", name, params, self.block())

    def js_params(self):
        self.eat("(")
        params = []
        while not self.at(")"):
            params.append(self.js_pattern())
            self.accept(",")
        self.eat(")")
        return params

    def js_pattern(self):
        if self.at("["):
            self.eat()
            names = []
            while not self.at("]"):
                names.append(self.js_pattern())
                self.accept(",")
            self.eat("]")
            return ("arraypat", names)
        if self.at("{"):
            self.eat()
            names = []
            while not self.at("}"):
                names.append(self.eat()[1])
                self.accept(",")
            self.eat("}")
            return ("objpat", names)
        return ("name", self.eat()[1])

    def js_decl(self):
        self.eat()   # const / let / var
        decls = []
        while True:
            pat = self.js_pattern()
            init = None
            if self.accept("="):
                init = self.assign()
            decls.append((pat, init))
            if not self.accept(","):
                break
        return ("decl", decls)

    # -- GLSL declarations -------------------------------------------------------------------------
    def glsl_decl_or_function(self, in_for=False):
        quals = []
        while self.peek()[1] in GLSL_QUALIFIERS:
            quals.append(self.eat()[1])
        typ = self.eat()[1]
        if typ not in GLSL_TYPES:
            raise SyntaxError(f"unknown GLSL type {typ!r}")
        name = self.eat()[1]
        if self.at("(") and not in_for:
            # function definition
            self.eat("(")
            params = []
            while not self.at(")"):
                while self.peek()[1] in GLSL_QUALIFIERS:
                    self.eat()
                ptype = self.eat()[1]
                pname = self.eat()[1]
                params.append((ptype, pname))
                self.accept(",")
            self.eat(")")
            return ("gfunc", typ, name, params, self.block())
        decls = []
        while True:
            size = None
            if self.accept("["):
                size = self.expr()
                self.eat("]")
            init = None
            if self.accept("="):
                init = self.assign()
            decls.append((name, size, init))
            if not self.accept(","):
                break
            name = self.eat()[1]
        if not in_for:
            self.eat(";")
        return ("gdecl", quals, typ, decls)

    # -- expressions -----------------------------------------------------------------------------
    def expr(self):
        e = self.assign()
        while self.at(","):      # comma operator (only in for-steps of the reference)
            self.eat()
            e = ("comma", e, self.assign())
        return e

    def is_arrow_ahead(self):
        if self.peek()[0] == "id" and self.at("=>", 1):
            return True
        if not self.at("("):
            return False
        depth = 0
        k = 0
        while True:
            tok = self.peek(k)
            if tok[0] == "eof":
                return False
            if tok[1] in "([{" and tok[0] == "op":
                depth += 1
            elif tok[1] in ")]}" and tok[0] == "op":
                depth -= 1
                if depth == 0:
                    return self.at("=>", k + 1)
            k += 1

    def arrow(self):
        if self.peek()[0] == "id":
            params = [("name", self.eat()[1])]
        else:
            params = self.js_params()
        self.eat("=>")
        if self.at("{"):
            return ("arrow", params, self.block(), False)
        return ("arrow", params, self.assign(), True)

    def assign(self):
        if self.lang == "js" and self.is_arrow_ahead():
            return self.arrow()
        left = self.ternary()
        if self.peek()[0] == "op" and self.peek()[1] in ("=", "+=", "-=", "*=", "/="):
            op = self.eat()[1]
            right = self.assign()
            return ("assign", op, left, right)
        return left

    def ternary(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.assign()
            self.eat(":")
            b = self.assign()
            return ("cond", c, a, b)
        return c

    LEVELS = [["||"], ["&&"], ["==", "!=", "===", "!=="], ["<", ">", "<=", ">="], ["+", "-"], ["*", "/", "%"]]

    def binary(self, level):
        if level == len(self.LEVELS):
            return self.unary()
        left = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.LEVELS[level]:
            op = self.eat()[1]
            right = self.binary(level + 1)
            left = ("bin", op, left, right)
        return left

    def unary(self):
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("-", "+", "!"):
            self.eat()
            return ("un", tok[1], self.unary())
        if tok[0] == "op" and tok[1] in ("++", "--"):
            self.eat()
            return ("preinc", tok[1], self.unary())
        if tok == ("id", "typeof"):
            self.eat()
            return ("typeof", self.unary())
        if tok == ("id", "new"):
            self.eat()
            callee = ("id", self.eat()[1])
            args = self.call_args()
            return ("new", callee, args)
        return self.power()

    def power(self):
        base = self.postfix()
        if self.at("**"):
            self.eat()
            return ("bin", "**", base, self.unary())   # right associative
        return base

    def call_args(self):
        self.eat("(")
        args = []
        while not self.at(")"):
            args.append(self.assign())
            self.accept(",")
        self.eat(")")
        return args

    def postfix(self):
        e = self.primary()
        while True:
            if self.at("("):
                e = ("call", e, self.call_args())
            elif self.at("["):
                self.eat()
                idx = self.expr()
                self.eat("]")
                e = ("index", e, idx)
            elif self.at("."):
                self.eat()
                e = ("member", e, self.eat()[1])
            elif self.peek()[0] == "op" and self.peek()[1] in ("++", "--"):
                e = ("postinc", self.eat()[1], e)
            else:
                return e

    def primary(self):
        kind, text = self.peek()
        if kind == "num":
            self.eat()
            is_float = any(c in text for c in ".eE")
            return ("num", text, is_float)
        if kind == "str":
            self.eat()
            return ("str", text[1:-1])
        if kind == "id":
            self.eat()
            return ("id", text)
        if self.at("("):
            self.eat()
            e = self.expr()
            self.eat(")")
            return e
        if self.at("[") and self.lang == "js":
            self.eat()
            items = []
            while not self.at("]"):
                items.append(self.assign())
                self.accept(",")
            self.eat("]")
            return ("array", items)
        if self.at("{") and self.lang == "js":
            self.eat()
            props = []
            while not self.at("}"):
                key = self.eat()[1]
                if self.accept(":"):
                    props.append((key, self.assign()))
                else:
                    props.append((key, ("id", key)))      # shorthand {xp,yp}
                self.accept(",")
            self.eat("}")
            return ("object", props)
        raise SyntaxError(f"unexpected token {text!r} (token {self.i})")


def parse(src: str, lang: str):
    return Parser(src, lang).program()
