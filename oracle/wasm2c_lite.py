#!/usr/bin/env python
"""wasm2c_lite -- translate the reference's SHIPPED WebAssembly module to C so that the oracle can
be pinned against the artefact users actually run, not only against a native build of its source.

TEST INFRASTRUCTURE ONLY. Nothing in the product imports or executes this.

Input : the Emscripten glue of the reference (src/speex_wasm.js), which embeds the module as a
        base64 data URI (26 447 bytes, MVP opcodes only, 31 functions).
Output: one C file (written under oracle/_ref/, git-ignored -- it is derived from the reference's
        binary and is never committed) that oracle/Makefile compiles to oracle/_ref/libspeex_wasm.so.

The translation is mechanical and keeps WebAssembly's semantics: a function's operand stack
becomes C locals named by stack depth and type (validation guarantees both are static), structured
control flow becomes labels and gotos, linear memory is one byte array accessed through memcpy,
integer division / shifts / float-to-int conversions follow the spec, and every f32/f64 operation
is a single C operation compiled with -ffp-contract=off (IEEE-754, no fusion) -- the same
arithmetic a WebAssembly engine performs. The two imports are implemented as the glue implements
them (emscripten_memcpy_big = copy inside linear memory, emscripten_resize_heap = grow the memory
by the glue's own rule, up to a fixed reservation).

usage: python oracle/wasm2c_lite.py /root/reference/src/speex_wasm.js oracle/_ref/speex_wasm.c
"""
import base64
import re
import struct
import sys

I32, I64, F32, F64 = 0x7F, 0x7E, 0x7D, 0x7C
CT = {I32: "u32", I64: "u64", F32: "f32", F64: "f64"}
SUF = {I32: "i", I64: "l", F32: "f", F64: "d"}


class Reader:
    def __init__(self, data, pos=0, end=None):
        self.d, self.p, self.end = data, pos, len(data) if end is None else end

    def byte(self):
        b = self.d[self.p]
        self.p += 1
        return b

    def u(self):  # unsigned LEB128
        r = s = 0
        while True:
            b = self.byte()
            r |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                return r

    def s(self, bits):  # signed LEB128
        r = s = 0
        while True:
            b = self.byte()
            r |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                if b & 0x40:
                    r -= 1 << s
                return r

    def bytes(self, n):
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def name(self):
        return self.bytes(self.u()).decode()

    def eof(self):
        return self.p >= self.end


def extract_module(js_path):
    js = open(js_path).read()
    m = re.search(r"data:application/octet-stream;base64,([A-Za-z0-9+/=]+)", js)
    if not m:
        raise SystemExit("no embedded wasm module in " + js_path)
    exports = {}
    for mm in re.finditer(r'Module\["(_[A-Za-z0-9_]+)"\]\s*=\s*function\s*\(\)\s*\{\s*return\s*\(\s*_[A-Za-z0-9_]+\s*='
                          r'\s*Module\["\1"\]\s*=\s*Module\["asm"\]\["([A-Za-z0-9_$]+)"\]', js):
        exports[mm.group(2)] = mm.group(1)
    # the glue's own memory initialisation: HEAP32[DYNAMICTOP_PTR >> 2] = DYNAMIC_BASE (where sbrk starts)
    glue = {}
    for key in ("DYNAMICTOP_PTR", "DYNAMIC_BASE"):
        mm = re.search(key + r"\s*=\s*(\d+)", js)
        if mm:
            glue[key] = int(mm.group(1))
    if re.search(r"HEAP32\[DYNAMICTOP_PTR\s*>>\s*2\]\s*=\s*DYNAMIC_BASE", js) is None:
        glue = {}
    exports["__glue__"] = glue
    return base64.b64decode(m.group(1)), exports


class Module:
    def __init__(self, data):
        r = Reader(data)
        assert r.bytes(4) == b"\0asm" and r.bytes(4) == b"\1\0\0\0"
        self.types, self.imports, self.funcs, self.globals = [], [], [], []
        self.exports, self.elems, self.codes, self.datas = {}, [], [], []
        self.mem_min = 0
        self.table_size = 0
        while not r.eof():
            sid, size = r.byte(), r.u()
            sec = Reader(data, r.p, r.p + size)
            r.p += size
            getattr(self, f"sec{sid}", lambda s: None)(sec)
        self.n_imp_funcs = sum(1 for i in self.imports if i[2] == 0)

    def sec1(self, s):
        for _ in range(s.u()):
            assert s.byte() == 0x60
            params = [s.byte() for _ in range(s.u())]
            results = [s.byte() for _ in range(s.u())]
            self.types.append((params, results))

    def sec2(self, s):
        for _ in range(s.u()):
            mod, name, kind = s.name(), s.name(), s.byte()
            if kind == 0:
                self.imports.append((mod, name, 0, s.u()))
            elif kind == 1:
                s.byte()
                flags = s.u()
                self.table_size = s.u()
                if flags & 1:
                    s.u()
                self.imports.append((mod, name, 1, None))
            elif kind == 2:
                flags = s.u()
                self.mem_min = s.u()
                if flags & 1:
                    s.u()
                self.imports.append((mod, name, 2, None))
            else:
                raise SystemExit("imported globals are not supported")

    def sec3(self, s):
        self.funcs = [s.u() for _ in range(s.u())]

    def sec4(self, s):
        for _ in range(s.u()):
            s.byte()
            flags = s.u()
            self.table_size = s.u()
            if flags & 1:
                s.u()

    def sec5(self, s):
        for _ in range(s.u()):
            flags = s.u()
            self.mem_min = s.u()
            if flags & 1:
                s.u()

    def const_expr(self, s):
        op = s.byte()
        if op == 0x41:
            v = s.s(32) & 0xFFFFFFFF
        elif op == 0x42:
            v = s.s(64) & 0xFFFFFFFFFFFFFFFF
        elif op == 0x43:
            v = struct.unpack("<I", s.bytes(4))[0]
        elif op == 0x44:
            v = struct.unpack("<Q", s.bytes(8))[0]
        else:
            raise SystemExit("unsupported constant expression")
        assert s.byte() == 0x0B
        return v

    def sec6(self, s):
        for _ in range(s.u()):
            t, mut = s.byte(), s.byte()
            self.globals.append((t, self.const_expr(s)))

    def sec7(self, s):
        for _ in range(s.u()):
            name, kind, idx = s.name(), s.byte(), s.u()
            if kind == 0:
                self.exports[name] = idx

    def sec9(self, s):
        for _ in range(s.u()):
            assert s.u() == 0
            off = self.const_expr(s)
            self.elems.append((off, [s.u() for _ in range(s.u())]))

    def sec10(self, s):
        for _ in range(s.u()):
            size = s.u()
            body = Reader(s.d, s.p, s.p + size)
            s.p += size
            local_types = []
            for _ in range(body.u()):
                n, t = body.u(), body.byte()
                local_types += [t] * n
            self.codes.append((local_types, body))

    def sec11(self, s):
        for _ in range(s.u()):
            assert s.u() == 0
            off = self.const_expr(s)
            self.datas.append((off, s.bytes(s.u())))

    def func_type(self, fidx):
        if fidx < self.n_imp_funcs:
            return self.types[[i for i in self.imports if i[2] == 0][fidx][3]]
        return self.types[self.funcs[fidx - self.n_imp_funcs]]


# (mnemonic-free) numeric tables: opcode -> (operand type, result type, C template)
BIN = {}
UN = {}
CMP = {}


def _fill():
    for base, t, sgn in ((0x6A, I32, "s32"), (0x7C, I64, "s64")):
        w = 32 if t == I32 else 64
        c = CT[t]
        ops = ["({a} + {b})", "({a} - {b})", "({a} * {b})", f"div_s{w}({{a}}, {{b}})", f"div_u{w}({{a}}, {{b}})",
               f"rem_s{w}({{a}}, {{b}})", f"rem_u{w}({{a}}, {{b}})", "({a} & {b})", "({a} | {b})", "({a} ^ {b})",
               f"({{a}} << ({{b}} & {w - 1}))", f"(({c})(({sgn}){{a}} >> ({{b}} & {w - 1})))", f"({{a}} >> ({{b}} & {w - 1}))",
               f"rotl{w}({{a}}, {{b}})", f"rotr{w}({{a}}, {{b}})"]
        for k, e in enumerate(ops):
            BIN[base + k] = (t, t, e)
    for base, t in ((0x67, I32), (0x79, I64)):
        w = 32 if t == I32 else 64
        for k, e in enumerate([f"clz{w}({{a}})", f"ctz{w}({{a}})", f"popcnt{w}({{a}})"]):
            UN[base + k] = (t, t, e)
    for base, t, sgn in ((0x46, I32, "s32"), (0x51, I64, "s64")):
        ops = ["({a} == {b})", "({a} != {b})", f"(({sgn}){{a}} < ({sgn}){{b}})", "({a} < {b})",
               f"(({sgn}){{a}} > ({sgn}){{b}})", "({a} > {b})", f"(({sgn}){{a}} <= ({sgn}){{b}})", "({a} <= {b})",
               f"(({sgn}){{a}} >= ({sgn}){{b}})", "({a} >= {b})"]
        for k, e in enumerate(ops):
            CMP[base + k] = (t, e)
    for base, t in ((0x5B, F32), (0x61, F64)):
        for k, e in enumerate(["({a} == {b})", "({a} != {b})", "({a} < {b})", "({a} > {b})", "({a} <= {b})", "({a} >= {b})"]):
            CMP[base + k] = (t, e)
    for base, t, sfx in ((0x8B, F32, "f"), (0x99, F64, "")):
        un = [f"fabs{sfx}({{a}})", "(-{a})", f"ceil{sfx}({{a}})", f"floor{sfx}({{a}})", f"trunc{sfx}({{a}})",
              f"nearbyint{sfx}({{a}})", f"sqrt{sfx}({{a}})"]
        for k, e in enumerate(un):
            UN[base + k] = (t, t, e)
        w = "32" if t == F32 else "64"
        bi = ["({a} + {b})", "({a} - {b})", "({a} * {b})", "({a} / {b})", f"fmin{w}({{a}}, {{b}})", f"fmax{w}({{a}}, {{b}})",
              f"copysign{sfx}({{a}}, {{b}})"]
        for k, e in enumerate(bi):
            BIN[base + 7 + k] = (t, t, e)
    conv = {
        0xA7: (I64, I32, "(u32){a}"), 0xA8: (F32, I32, "(u32)(s32){a}"), 0xA9: (F32, I32, "(u32){a}"),
        0xAA: (F64, I32, "(u32)(s32){a}"), 0xAB: (F64, I32, "(u32){a}"),
        0xAC: (I32, I64, "(u64)(s64)(s32){a}"), 0xAD: (I32, I64, "(u64){a}"),
        0xAE: (F32, I64, "(u64)(s64){a}"), 0xAF: (F32, I64, "(u64){a}"), 0xB0: (F64, I64, "(u64)(s64){a}"),
        0xB1: (F64, I64, "(u64){a}"),
        0xB2: (I32, F32, "(f32)(s32){a}"), 0xB3: (I32, F32, "(f32){a}"), 0xB4: (I64, F32, "(f32)(s64){a}"),
        0xB5: (I64, F32, "(f32){a}"), 0xB6: (F64, F32, "(f32){a}"),
        0xB7: (I32, F64, "(f64)(s32){a}"), 0xB8: (I32, F64, "(f64){a}"), 0xB9: (I64, F64, "(f64)(s64){a}"),
        0xBA: (I64, F64, "(f64){a}"), 0xBB: (F32, F64, "(f64){a}"),
        0xBC: (F32, I32, "bits_f32({a})"), 0xBD: (F64, I64, "bits_f64({a})"),
        0xBE: (I32, F32, "f32_bits({a})"), 0xBF: (I64, F64, "f64_bits({a})"),
    }
    UN.update(conv)


_fill()

LOADS = {0x28: (I32, "u32", 4, ""), 0x29: (I64, "u64", 8, ""), 0x2A: (F32, "f32", 4, ""), 0x2B: (F64, "f64", 8, ""),
         0x2C: (I32, "int8_t", 1, "(u32)(s32)"), 0x2D: (I32, "uint8_t", 1, "(u32)"),
         0x2E: (I32, "int16_t", 2, "(u32)(s32)"), 0x2F: (I32, "uint16_t", 2, "(u32)"),
         0x30: (I64, "int8_t", 1, "(u64)(s64)"), 0x31: (I64, "uint8_t", 1, "(u64)"),
         0x32: (I64, "int16_t", 2, "(u64)(s64)"), 0x33: (I64, "uint16_t", 2, "(u64)"),
         0x34: (I64, "int32_t", 4, "(u64)(s64)"), 0x35: (I64, "uint32_t", 4, "(u64)")}
STORES = {0x36: (I32, "u32"), 0x37: (I64, "u64"), 0x38: (F32, "f32"), 0x39: (F64, "f64"),
          0x3A: (I32, "uint8_t"), 0x3B: (I32, "uint16_t"), 0x3C: (I64, "uint8_t"), 0x3D: (I64, "uint16_t"),
          0x3E: (I64, "uint32_t")}

PRELUDE = r"""/* GENERATED by oracle/wasm2c_lite.py from the reference's shipped WebAssembly module.
 * Derived from the reference's binary: lives under oracle/_ref/ (git-ignored), never committed. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef uint32_t u32; typedef int32_t s32; typedef uint64_t u64; typedef int64_t s64;
typedef float f32; typedef double f64;
static uint8_t *MEM; static u32 MEM_PAGES; static const u32 MEM_MAX_PAGES = %(max_pages)du;
static void trap(void) { abort(); }
#define LD(T, a) ({ T v_; memcpy(&v_, MEM + (a), sizeof(T)); v_; })
#define ST(T, a, v) do { T v_ = (T)(v); memcpy(MEM + (a), &v_, sizeof(T)); } while (0)
static inline u32 div_s32(u32 a, u32 b) { if (!b || (a == 0x80000000u && b == 0xffffffffu)) trap(); return (u32)((s32)a / (s32)b); }
static inline u32 div_u32(u32 a, u32 b) { if (!b) trap(); return a / b; }
static inline u32 rem_s32(u32 a, u32 b) { if (!b) trap(); if (b == 0xffffffffu) return 0; return (u32)((s32)a %% (s32)b); }
static inline u32 rem_u32(u32 a, u32 b) { if (!b) trap(); return a %% b; }
static inline u64 div_s64(u64 a, u64 b) { if (!b || (a == 0x8000000000000000ull && b == ~0ull)) trap(); return (u64)((s64)a / (s64)b); }
static inline u64 div_u64(u64 a, u64 b) { if (!b) trap(); return a / b; }
static inline u64 rem_s64(u64 a, u64 b) { if (!b) trap(); if (b == ~0ull) return 0; return (u64)((s64)a %% (s64)b); }
static inline u64 rem_u64(u64 a, u64 b) { if (!b) trap(); return a %% b; }
static inline u32 rotl32(u32 a, u32 b) { b &= 31; return b ? (a << b) | (a >> (32 - b)) : a; }
static inline u32 rotr32(u32 a, u32 b) { b &= 31; return b ? (a >> b) | (a << (32 - b)) : a; }
static inline u64 rotl64(u64 a, u64 b) { b &= 63; return b ? (a << b) | (a >> (64 - b)) : a; }
static inline u64 rotr64(u64 a, u64 b) { b &= 63; return b ? (a >> b) | (a << (64 - b)) : a; }
static inline u32 clz32(u32 a) { return a ? (u32)__builtin_clz(a) : 32; }
static inline u32 ctz32(u32 a) { return a ? (u32)__builtin_ctz(a) : 32; }
static inline u32 popcnt32(u32 a) { return (u32)__builtin_popcount(a); }
static inline u64 clz64(u64 a) { return a ? (u64)__builtin_clzll(a) : 64; }
static inline u64 ctz64(u64 a) { return a ? (u64)__builtin_ctzll(a) : 64; }
static inline u64 popcnt64(u64 a) { return (u64)__builtin_popcountll(a); }
static inline u32 bits_f32(f32 a) { u32 v; memcpy(&v, &a, 4); return v; }
static inline u64 bits_f64(f64 a) { u64 v; memcpy(&v, &a, 8); return v; }
static inline f32 f32_bits(u32 a) { f32 v; memcpy(&v, &a, 4); return v; }
static inline f64 f64_bits(u64 a) { f64 v; memcpy(&v, &a, 8); return v; }
/* wasm min/max: NaN if either is NaN, -0 < +0 */
static inline f32 fmin32(f32 a, f32 b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? a : b; return a < b ? a : b; }
static inline f32 fmax32(f32 a, f32 b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? b : a; return a > b ? a : b; }
static inline f64 fmin64(f64 a, f64 b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? a : b; return a < b ? a : b; }
static inline f64 fmax64(f64 a, f64 b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? b : a; return a > b ? a : b; }
"""


class FuncGen:
    def __init__(self, mod, fidx, out):
        self.m, self.fidx, self.out = mod, fidx, out
        params, results = mod.func_type(fidx)
        local_types, body = mod.codes[fidx - mod.n_imp_funcs]
        self.params, self.results = params, results
        self.ltypes = list(params) + local_types
        self.body = body
        self.stack = []  # types
        self.vars = set()
        self.lines = []
        self.ctrl = []
        self.nlabel = 0

    def var(self, depth, t):
        name = f"s{depth}{SUF[t]}"
        self.vars.add((name, t))
        return name

    def push(self, t):
        self.stack.append(t)
        return self.var(len(self.stack) - 1, t)

    def pop(self, t=None):
        tt = self.stack.pop()
        if t is not None and tt != t:
            raise SystemExit(f"func {self.fidx}: type mismatch at {self.body.p}: got {tt:#x} want {t:#x}")
        return self.var(len(self.stack), tt)

    def top(self):
        return self.var(len(self.stack) - 1, self.stack[-1])

    def emit(self, line):
        self.lines.append("  " + line)

    def label(self):
        self.nlabel += 1
        return self.nlabel

    def blocktype(self):
        b = self.body.byte()
        return None if b == 0x40 else b

    def branch_code(self, depth):
        """statements that carry the branch operand to the target's result slot and jump"""
        fr = self.ctrl[-1 - depth]
        code = ""
        if fr["kind"] != "loop" and fr["result"] is not None:
            src = self.var(len(self.stack) - 1, fr["result"])
            dst = self.var(fr["height"], fr["result"])
            if src != dst:
                code += f"{dst} = {src}; "
        if fr["kind"] == "func":
            if fr["result"] is not None:
                return code + f"return {self.var(fr['height'], fr['result'])};"
            return code + "return;"
        return code + f"goto L{fr['label']};"

    def skip_dead(self):
        """after an unconditional transfer: skip to the matching else/end of the current frame"""
        depth = 0
        b = self.body
        while True:
            op = b.byte()
            if op in (0x02, 0x03, 0x04):
                b.byte()
                depth += 1
            elif op == 0x05 and depth == 0:
                return 0x05
            elif op == 0x0B:
                if depth == 0:
                    return 0x0B
                depth -= 1
            elif op in (0x0C, 0x0D, 0x10, 0x20, 0x21, 0x22, 0x23, 0x24):
                b.u()
            elif op == 0x0E:
                for _ in range(b.u() + 1):
                    b.u()
            elif op == 0x11:
                b.u()
                b.byte()
            elif 0x28 <= op <= 0x3E:
                b.u()
                b.u()
            elif op in (0x3F, 0x40):
                b.byte()
            elif op == 0x41:
                b.s(32)
            elif op == 0x42:
                b.s(64)
            elif op == 0x43:
                b.bytes(4)
            elif op == 0x44:
                b.bytes(8)

    def end_frame(self, op):
        """handle `else` / `end` for the innermost frame; returns False when the function ended"""
        fr = self.ctrl[-1]
        if op == 0x05:
            assert fr["kind"] == "if"
            if not fr["dead"] and fr["result"] is not None:
                src, dst = self.pop(fr["result"]), self.var(fr["height"], fr["result"])
                if src != dst:
                    self.emit(f"{dst} = {src};")
            self.emit(f"goto L{fr['label']};")
            self.emit(f"E{fr['label']}:;")
            fr["has_else"] = True
            fr["dead"] = False
            del self.stack[fr["height"]:]
            return True
        # end
        if not fr["dead"] and fr["result"] is not None and fr["kind"] != "func":
            src, dst = self.pop(fr["result"]), self.var(fr["height"], fr["result"])
            if src != dst:
                self.emit(f"{dst} = {src};")
        self.ctrl.pop()
        if fr["kind"] == "func":
            if not fr["dead"]:
                if fr["result"] is not None:
                    self.emit(f"return {self.pop(fr['result'])};")
                else:
                    self.emit("return;")
            return False
        del self.stack[fr["height"]:]
        if fr["kind"] == "if" and not fr["has_else"]:
            self.emit(f"E{fr['label']}:;")
        if fr["kind"] != "loop":
            self.emit(f"L{fr['label']}:;")
        if fr["result"] is not None:
            self.push(fr["result"])
        return True

    def dead(self):
        self.ctrl[-1]["dead"] = True
        op = self.skip_dead()
        return self.end_frame(op)

    def run(self):
        m, b = self.m, self.body
        res = self.results[0] if self.results else None
        self.ctrl.append({"kind": "func", "label": 0, "result": res, "height": 0, "dead": False, "has_else": False})
        alive = True
        while alive:
            op = b.byte()
            if op == 0x00:
                self.emit("trap();")
                alive = self.dead()
            elif op == 0x01:
                pass
            elif op in (0x02, 0x03):
                bt = self.blocktype()
                lab = self.label()
                kind = "block" if op == 0x02 else "loop"
                self.ctrl.append({"kind": kind, "label": lab, "result": bt, "height": len(self.stack), "dead": False,
                                  "has_else": False})
                if kind == "loop":
                    self.emit(f"L{lab}:;")
            elif op == 0x04:
                bt = self.blocktype()
                cond = self.pop(I32)
                lab = self.label()
                self.ctrl.append({"kind": "if", "label": lab, "result": bt, "height": len(self.stack), "dead": False,
                                  "has_else": False})
                self.emit(f"if (!{cond}) goto E{lab};")
            elif op in (0x05, 0x0B):
                alive = self.end_frame(op)
            elif op == 0x0C:
                self.emit(self.branch_code(b.u()))
                alive = self.dead()
            elif op == 0x0D:
                d = b.u()
                cond = self.pop(I32)
                self.emit(f"if ({cond}) {{ {self.branch_code(d)} }}")
            elif op == 0x0E:
                targets = [b.u() for _ in range(b.u())]
                default = b.u()
                idx = self.pop(I32)
                self.emit(f"switch ({idx}) {{")
                for k, d in enumerate(targets):
                    self.emit(f"  case {k}: {self.branch_code(d)}")
                self.emit(f"  default: {self.branch_code(default)}")
                self.emit("}")
                alive = self.dead()
            elif op == 0x0F:
                self.emit(self.branch_code(len(self.ctrl) - 1))
                alive = self.dead()
            elif op == 0x10:
                f = b.u()
                params, results = m.func_type(f)
                args = [self.pop(t) for t in reversed(params)][::-1]
                call = f"fn{f}({', '.join(args)})"
                self.emit(f"{self.push(results[0])} = {call};" if results else f"{call};")
            elif op == 0x11:
                ti = b.u()
                b.byte()
                params, results = m.types[ti]
                idx = self.pop(I32)
                args = [self.pop(t) for t in reversed(params)][::-1]
                call = f"ci{ti}({', '.join([idx] + args)})"
                self.emit(f"{self.push(results[0])} = {call};" if results else f"{call};")
            elif op == 0x1A:
                self.pop()
            elif op == 0x1B:
                c = self.pop(I32)
                v2 = self.pop()
                t = self.stack[-1]
                v1 = self.pop(t)
                self.emit(f"{self.push(t)} = {c} ? {v1} : {v2};")
            elif op == 0x20:
                i = b.u()
                self.emit(f"{self.push(self.ltypes[i])} = l{i};")
            elif op == 0x21:
                i = b.u()
                self.emit(f"l{i} = {self.pop(self.ltypes[i])};")
            elif op == 0x22:
                i = b.u()
                self.emit(f"l{i} = {self.top()};")
            elif op == 0x23:
                i = b.u()
                self.emit(f"{self.push(m.globals[i][0])} = g{i};")
            elif op == 0x24:
                i = b.u()
                self.emit(f"g{i} = {self.pop(m.globals[i][0])};")
            elif op in LOADS:
                b.u()
                off = b.u()
                t, ct, _, cast = LOADS[op]
                a = self.pop(I32)
                self.emit(f"{self.push(t)} = {cast}LD({ct}, (u64){a} + {off}u);")
            elif op in STORES:
                b.u()
                off = b.u()
                t, ct = STORES[op]
                v = self.pop(t)
                a = self.pop(I32)
                self.emit(f"ST({ct}, (u64){a} + {off}u, {v});")
            elif op == 0x3F:
                b.byte()
                self.emit(f"{self.push(I32)} = MEM_PAGES;")
            elif op == 0x40:
                b.byte()
                n = self.pop(I32)
                self.emit(f"{self.push(I32)} = mem_grow({n});")
            elif op == 0x41:
                self.emit(f"{self.push(I32)} = {b.s(32) & 0xFFFFFFFF}u;")
            elif op == 0x42:
                self.emit(f"{self.push(I64)} = {b.s(64) & 0xFFFFFFFFFFFFFFFF}ull;")
            elif op == 0x43:
                self.emit(f"{self.push(F32)} = f32_bits({struct.unpack('<I', b.bytes(4))[0]}u);")
            elif op == 0x44:
                self.emit(f"{self.push(F64)} = f64_bits({struct.unpack('<Q', b.bytes(8))[0]}ull);")
            elif op == 0x45 or op == 0x50:
                t = I32 if op == 0x45 else I64
                a = self.pop(t)
                self.emit(f"{self.push(I32)} = ({a} == 0);")
            elif op in CMP:
                t, e = CMP[op]
                y = self.pop(t)
                x = self.pop(t)
                self.emit(f"{self.push(I32)} = {e.format(a=x, b=y)};")
            elif op in BIN:
                t, rt, e = BIN[op]
                y = self.pop(t)
                x = self.pop(t)
                self.emit(f"{self.push(rt)} = {e.format(a=x, b=y)};")
            elif op in UN:
                t, rt, e = UN[op]
                x = self.pop(t)
                self.emit(f"{self.push(rt)} = {e.format(a=x)};")
            else:
                raise SystemExit(f"func {self.fidx}: unsupported opcode {op:#x} at {b.p}")
        rt = CT[self.results[0]] if self.results else "void"
        sig = ", ".join(f"{CT[t]} l{i}" for i, t in enumerate(self.params)) or "void"
        o = self.out
        o.append(f"static {rt} fn{self.fidx}({sig}) {{")
        for i in range(len(self.params), len(self.ltypes)):
            o.append(f"  {CT[self.ltypes[i]]} l{i} = 0;")
        for name, t in sorted(self.vars):
            o.append(f"  {CT[t]} {name} = 0; (void){name};")
        o.extend(self.lines)
        o.append("}")


def translate(js_path, c_path):
    data, export_names = extract_module(js_path)
    m = Module(data)
    max_pages = 8192  # 512 MiB of lazily committed address space; growth beyond is refused like a full heap
    out = [PRELUDE % {"max_pages": max_pages}]
    for i, (t, v) in enumerate(m.globals):
        init = {I32: f"{v}u", I64: f"{v}ull", F32: f"0; /* bits {v} */", F64: f"0; /* bits {v} */"}[t]
        out.append(f"static {CT[t]} g{i} = {init};")
    nf = m.n_imp_funcs + len(m.funcs)
    for f in range(nf):
        params, results = m.func_type(f)
        rt = CT[results[0]] if results else "void"
        out.append(f"static {rt} fn{f}({', '.join(CT[t] for t in params) or 'void'});")
    out.append("static u32 mem_grow(u32 n) { u32 old = MEM_PAGES; if ((u64)old + n > MEM_MAX_PAGES) return 0xffffffffu; "
               "MEM_PAGES += n; return old; }")
    # imports, implemented as the Emscripten glue implements them
    fimps = [i for i in m.imports if i[2] == 0]
    for f, (mod, name, _, ti) in enumerate(fimps):
        params, results = m.types[ti]
        sig = ", ".join(f"{CT[t]} a{k}" for k, t in enumerate(params))
        if len(params) == 3:  # emscripten_memcpy_big(dest, src, num): HEAPU8.copyWithin
            out.append(f"static {CT[results[0]] if results else 'void'} fn{f}({sig}) {{ memmove(MEM + a0, MEM + a1, a2); "
                       + ("return a0; }" if results else "}"))
        elif len(params) == 1:  # emscripten_resize_heap(requestedSize): grow like the glue (ALLOW_MEMORY_GROWTH)
            out.append(f"static u32 fn{f}({sig}) {{ u64 old = (u64)MEM_PAGES << 16, over = old + old / 5, want = a0; "
                       "if (over > (u64)a0 + 100663296ull) over = (u64)a0 + 100663296ull; if (want < over) want = over; "
                       "if (want < 16777216ull) want = 16777216ull; u64 pages = (want + 65535) >> 16; "
                       "if (pages > MEM_MAX_PAGES) return 0; if (pages > MEM_PAGES) MEM_PAGES = (u32)pages; return 1; }")
        else:
            raise SystemExit(f"unexpected import {mod}.{name}")
    # function table for call_indirect
    table = [0xFFFFFFFF] * max(m.table_size, 1)
    for off, fs in m.elems:
        for k, f in enumerate(fs):
            table[off + k] = f
    out.append(f"static const u32 TABLE[{len(table)}] = {{{', '.join(str(x) + 'u' for x in table)}}};")
    for ti, (params, results) in enumerate(m.types):
        rt = CT[results[0]] if results else "void"
        sig = ", ".join(["u32 idx"] + [f"{CT[t]} a{k}" for k, t in enumerate(params)])
        args = ", ".join(f"a{k}" for k in range(len(params)))
        out.append(f"static {rt} ci{ti}({sig}) {{")
        out.append(f"  if (idx >= {len(table)}u) trap();")
        out.append("  switch (TABLE[idx]) {")
        for f in sorted(set(x for x in table if x != 0xFFFFFFFF)):
            if m.func_type(f) == (params, results):
                out.append(f"    case {f}u: {'return ' if results else ''}fn{f}({args});{'' if results else ' return;'}")
        out.append("    default: trap();")
        out.append("  }")
        if results:
            out.append("  return 0;")
        out.append("}")
    for f in range(m.n_imp_funcs, nf):
        FuncGen(m, f, out).run()
    # instantiate: memory, data segments, constructors
    out.append("static int READY;")
    out.append("static void instantiate(void) {")
    out.append("  if (READY) return;")
    out.append(f"  MEM = (uint8_t *)calloc((size_t)MEM_MAX_PAGES, 65536); MEM_PAGES = {max(m.mem_min, 1)}u;")
    for k, (off, blob) in enumerate(m.datas):
        out.append(f"  {{ static const uint8_t d{k}[] = {{{','.join(str(x) for x in blob)}}}; memcpy(MEM + {off}u, d{k}, sizeof d{k}); }}")
    glue = export_names.pop("__glue__", {})
    if "DYNAMICTOP_PTR" in glue and "DYNAMIC_BASE" in glue:  # done by the JS glue before any call
        out.append(f"  ST(u32, {glue['DYNAMICTOP_PTR']}u, {glue['DYNAMIC_BASE']}u);")
    if "___wasm_call_ctors" in export_names.values():
        key = [k for k, v in export_names.items() if v == "___wasm_call_ctors"][0]
        out.append(f"  fn{m.exports[key]}();")
    out.append("  READY = 1;")
    out.append("}")
    # exported entry points under their Emscripten names (wasm_ prefix), plus memory access
    out.append("uint8_t *wasm_memory(void) { instantiate(); return MEM; }")
    out.append("u32 wasm_memory_bytes(void) { instantiate(); return MEM_PAGES * 65536u; }")
    for key, fidx in sorted(m.exports.items()):
        name = export_names.get(key)
        if not name or name == "___wasm_call_ctors":
            continue
        params, results = m.func_type(fidx)
        rt = CT[results[0]] if results else "void"
        sig = ", ".join(f"{CT[t]} a{k}" for k, t in enumerate(params)) or "void"
        args = ", ".join(f"a{k}" for k in range(len(params)))
        out.append(f"{rt} wasm{name}({sig}) {{ instantiate(); {'return ' if results else ''}fn{fidx}({args}); }}")
    open(c_path, "w").write("\n".join(out) + "\n")
    return m, export_names


if __name__ == "__main__":
    mod, names = translate(sys.argv[1], sys.argv[2])
    print(f"translated {len(mod.funcs)} functions, {len(mod.datas)} data segments, exports: "
          + ", ".join(sorted(v for v in names.values())))
