#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle/_ref): rewrites the reference's GLSL 3.30 hot-path shaders into C++ that compiles against
glsl_types.hpp, so that the reference's own shader text runs on the CPU and pins the hand-written oracle (oracle/*.c).

The shaders are read where they lie (/root/reference/src/shader); the generated C++ goes to oracle/_ref/gen/ (git-ignored).
Nothing of the reference is copied into the repository.

The rewrite is mechanical and keeps every expression as written:
  * #version / #pragma dropped, #include "shader/x.glsl" inlined, comments removed
  * layout(...) / flat / uniform / in / out qualifiers dropped at file scope; globals become members of one struct per
    stage (`struct Shader : glsl::StageBase`), functions become its methods; `out` names are recorded so that
    EmitVertex() can snapshot them (geometry stages)
  * interface blocks `out SURFEL {...} vs_out;` / `in SURFEL {...} gs_in[];` become a nested struct + a member (array of 1)
  * floating literals get an `f` suffix (GLSL literals are fp32; C++ ones would silently promote expressions to double)
  * `discard;` sets a flag and returns

usage: glsl2cpp.py <reference_src_dir> <out_dir> shader/a.vert shader/b.geom ...
"""
import hashlib
import os
import re
import sys

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)[fF]?(?![\w.])")


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def load(src_dir, rel, seen=None):
    seen = seen or []
    text = open(os.path.join(src_dir, rel)).read()
    digest = [(rel, hashlib.sha256(text.encode()).hexdigest()[:16])]
    out = []
    for line in strip_comments(text).splitlines():
        m = re.match(r'\s*#\s*include\s+"([^"]+)"', line)
        if m:
            if m.group(1) not in seen:
                seen.append(m.group(1))
                sub, d = load(src_dir, m.group(1), seen)
                out.append(sub)
                digest += d
            continue
        if re.match(r"\s*#\s*(version|pragma)", line):
            continue
        out.append(line)
    return "\n".join(out), digest


def chunks(text):
    """file-scope pieces: declarations (ending in ';' at depth 0) and function definitions (ending in their '}')"""
    depth, start, i, n = 0, 0, 0, len(text)
    while i < n:
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                head = text[start:text.index("{", start)]
                if "(" in head:                       # a function body just closed
                    yield text[start:i + 1].strip()
                    start = i + 1
        elif c == ";" and depth == 0:
            piece = text[start:i + 1].strip()
            if piece != ";":
                yield piece
            start = i + 1
        i += 1
    rest = text[start:].strip()
    if rest:
        raise SystemExit("glsl2cpp: trailing text not understood: %r" % rest[:80])


def rewrite_decl(piece, outs):
    """one file-scope declaration -> member declaration text"""
    p = re.sub(r"layout\s*\([^)]*\)\s*", "", piece).strip()
    if re.fullmatch(r"(in|out)\s*;", p):
        return None                                    # layout(points) in;  /  layout(..., max_vertices = N) out;
    p = re.sub(r"^\s*flat\s+", "", p)
    m = re.fullmatch(r"(in|out)\s+(\w+)\s*\{(.*)\}\s*(\w+)\s*(\[\s*\])?\s*;", p, flags=re.S)
    if m:
        qual, block, body, inst, arr = m.groups()
        if qual == "out":
            outs.append((block, inst))
        return "struct %s {%s};\n  %s %s%s;" % (block, body, block, inst, "[1]" if arr else "")
    m = re.fullmatch(r"(in|out)\s+(\w+)\s+(\w+)\s*;", p)
    if m:
        qual, typ, name = m.groups()
        if qual == "out":
            outs.append((typ, name))
        return "%s %s%s;" % (typ, name, " = %s()" % typ if typ in ("int", "float", "bool") else "")
    m = re.fullmatch(r"uniform\s+(\w+)\s+(\w+)\s*;", p)
    if m:
        typ, name = m.groups()
        return "%s %s%s;" % (typ, name, " = %s()" % typ if typ in ("int", "float", "bool") else "")
    return p                                           # const / plain global with or without initialiser


def transpile(src_dir, rel):
    text, digest = load(src_dir, rel)
    text = FLOAT_LIT.sub(lambda m: m.group(1) + "f", text)
    text = re.sub(r"\bdiscard\s*;", "{ discarded_ = true; return; }", text)
    outs, body = [], []
    for piece in chunks(text):
        head = piece.split("{", 1)[0]
        if "(" in head and piece.endswith("}"):
            body.append(piece)                         # function -> method, text unchanged
        else:
            d = rewrite_decl(piece, outs)
            if d:
                body.append(d)
    name = "ref_" + os.path.basename(rel).replace(".", "_")
    snap = "".join(" e.%s = %s;" % (n, n) for _, n in outs)
    fields = "".join(" %s %s;" % (t, n) for t, n in outs)
    lines = ["// GENERATED by oracle/ref_harness/glsl2cpp.py -- derived from the reference's source, never committed.",
             "// sources: " + ", ".join("%s sha256:%s" % d for d in digest),
             "#pragma once", '#include "glsl_types.hpp"', "namespace glsl { namespace %s {" % name,
             "struct Shader : StageBase {"]
    lines += ["  " + b for b in body]
    lines += ["  struct Emitted { vec4 gl_Position;%s };" % fields,
              "  std::vector<Emitted> emitted_; std::vector<int> prim_end_;",
              "  void EmitVertex() { Emitted e; e.gl_Position = gl_Position;%s emitted_.push_back(e); }" % snap,
              "  void EndPrimitive() { prim_end_.push_back((int)emitted_.size()); }",
              "};", "}}  // namespace", ""]
    return name, "\n".join(lines), digest


def swizzle_inc(n):
    names = ["xyzw"[:n], "rgba"[:n]]
    out = []
    for comps in names:
        for k in (2, 3, 4):
            idx = [0] * k
            while True:
                nm = "".join(comps[i] for i in idx)
                out.append("Swz<%d, %s> %s;" % (n, ", ".join(map(str, idx)), nm))
                j = k - 1
                while j >= 0:
                    idx[j] += 1
                    if idx[j] < n:
                        break
                    idx[j] = 0
                    j -= 1
                if j < 0:
                    break
    return "\n".join(out) + "\n"


def main():
    src_dir, out_dir, rels = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(out_dir, exist_ok=True)
    for n in (2, 3, 4):
        open(os.path.join(out_dir, "swizzles%d.inc" % n), "w").write(swizzle_inc(n))
    manifest = []
    for rel in rels:
        name, cpp, digest = transpile(src_dir, rel)
        open(os.path.join(out_dir, name + ".hpp"), "w").write(cpp)
        manifest.append("%s <- %s" % (name, ", ".join("%s@%s" % d for d in digest)))
    open(os.path.join(out_dir, "MANIFEST.txt"), "w").write("\n".join(manifest) + "\n")


if __name__ == "__main__":
    main()
